"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): the row partition, the shard's
standard-form data and the one exchange step per iteration -- all-reduce(sum) of the partial A_g' y_g
with the scalar |dy|^2 riding in the tail slot (DESIGN.md "Multi-GPU") -- reproduce the single-rank
quantities.  The SpMV itself is stood in for by scipy on the oracle's scaled matrix; the CUDA kernels
are exercised by the -m gpu tests."""
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
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
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
