"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): the nnz-balanced row partition, the
column shards, and the two exchange steps of a PDHG pass in their segmented layout --
reduce-scatter of the partial A_g' y_g with |dy|^2 and the row-side interaction riding in the two tail slots
of every segment, and all-gather of the x shards with |dx|^2 in the tail -- reproduce the single-rank
quantities, including the identity (x-x').(A'y - A'y') == (Ax - Ax').(y - y') that lets the interaction be
taken on the row side.  The SpMV itself is stood in for by scipy on the oracle's scaled matrix; the CUDA
kernels are exercised by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp
        from highs_b200 import engine
        from highs_b200.lp import synthetic_lp
        from oracle import binding as ob
        lp = synthetic_lp(4000, 3000, 6, 21, dense_col_nnz=1500)
        bounds = engine.partition_rows(lp, world)
        # every rank must derive the same partition
        t = torch.tensor(bounds, dtype=torch.int64)
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref)
        f = ob.formulate_and_scale(lp)
        A = sp.csr_matrix((f["rval"], f["ridx"], f["rbeg"]), shape=(f["m"], f["n"]))
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        Ag = A[r0:r1]
        rng = np.random.default_rng(5)          # same seed everywhere: replicated x, full y then sliced
        x, y, yold = rng.standard_normal(f["n"]), rng.standard_normal(f["m"]), rng.standard_normal(f["m"])
        # unique-id style plumbing: rank 0 creates 128 opaque bytes, everyone receives the same
        ids = [bytes(rng.integers(0, 256, 128, dtype=np.uint8)) if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        assert isinstance(ids[0], bytes) and len(ids[0]) == 128
        # Ax is local
        ax_local = Ag @ x
        assert np.allclose(ax_local, (A @ x)[r0:r1], rtol=1e-13, atol=1e-13)
        # A'y needs ONE all-reduce; |dy|^2 travels in the tail slot of the same buffer
        buf = np.zeros(f["n"] + 1)
        buf[: f["n"]] = Ag.T @ y[r0:r1]
        buf[f["n"]] = np.sum((y[r0:r1] - yold[r0:r1]) ** 2)
        tb = torch.from_numpy(buf)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        assert np.allclose(buf[: f["n"]], A.T @ y, rtol=1e-12, atol=1e-12)
        assert np.isclose(buf[f["n"]], np.sum((y - yold) ** 2), rtol=1e-12)
        # ---- the segmented exchange of one PDHG pass
        n, m = f["n"], f["m"]
        shard_len = ((n + world - 1) // world + 1) & ~1
        seg_len = shard_len + 2
        pos = lambda j: (j // shard_len) * seg_len + (j % shard_len)
        c0 = rank * shard_len
        nl = max(0, min(shard_len, n - c0))
        rng2 = np.random.default_rng(77)        # identical on every rank (rank 0 consumed extra numbers above)
        xnew = rng2.standard_normal(n)
        ynew = rng2.standard_normal(m)
        # all-gather: my x' shard (+ |dx|^2 partial in the tail) -> xfull on every rank
        send = np.zeros(seg_len)
        send[:nl] = xnew[c0:c0 + nl]
        send[shard_len] = np.sum((x[c0:c0 + nl] - xnew[c0:c0 + nl]) ** 2)
        gathered = [torch.zeros(seg_len, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(send))
        xfull = torch.cat(gathered).numpy()
        assert np.array_equal(xfull[[pos(j) for j in range(0, n, 97)]], xnew[::97])
        dx2 = sum(xfull[g * seg_len + shard_len] for g in range(world))
        assert np.isclose(dx2, np.sum((x - xnew) ** 2), rtol=1e-12)
        # reduce-scatter (emulated with all-reduce + slice): partial A_g' y' with the scalars in every tail
        part = np.zeros(world * seg_len)
        pg = Ag.T @ ynew[r0:r1]
        idx = np.array([pos(j) for j in range(n)])
        part[idx] = pg
        dy2_g = np.sum((y[r0:r1] - ynew[r0:r1]) ** 2)
        inter_g = np.dot(Ag @ x - Ag @ xnew, y[r0:r1] - ynew[r0:r1])
        for g in range(world):
            part[g * seg_len + shard_len] = dy2_g
            part[g * seg_len + shard_len + 1] = inter_g
        tp = torch.from_numpy(part)
        dist.all_reduce(tp)
        red = part[rank * seg_len:(rank + 1) * seg_len]
        assert np.allclose(red[:nl], (A.T @ ynew)[c0:c0 + nl], rtol=1e-12, atol=1e-12)
        assert np.isclose(red[shard_len], np.sum((y - ynew) ** 2), rtol=1e-12)
        inter_col = np.dot(x - xnew, A.T @ y - A.T @ ynew)      # the reference's column-side form
        assert np.isclose(red[shard_len + 1], inter_col, rtol=1e-9, atol=1e-9)
        # gather of a row-partitioned vector at the end of the solve = zero-padded sum
        full = np.zeros(f["m"])
        full[r0:r1] = y[r0:r1]
        tf = torch.from_numpy(full)
        dist.all_reduce(tf)
        assert np.array_equal(full, y)
        # all ranks take the same decisions from the same reduced scalars (max over ranks of timings)
        tm = torch.tensor([float(rank + 1)])
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        assert tm.item() == world
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_row_partition_exchange_world2(engine_lib, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
