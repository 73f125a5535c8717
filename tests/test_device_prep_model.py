"""CPU model of the ALGORITHMS the device prologue uses (highs_b200/csrc/device_prep.cu), against the host twin
(host_prep.cpp, through the b200pdlp_form_* accessors): the scan-based stable partition of every column, the radix-sort
transposition and its second pass for unsorted columns, the window-key length sort, the sliced-ELL plan.  numpy stands in
for the kernels and CUB (exclusive scans, stable sorts), so this checks the index arithmetic and the equivalences the
kernels rely on -- without a GPU; the kernels themselves are compared bit for bit on hardware (tests/test_gpu_device_prep.py)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

FILES = sorted(glob.glob(os.path.join(GOLDEN, "instances", "*.b2lp")))


def host_form(lp, scaling=0):
    from highs_b200 import engine
    L = engine.lib()
    clp, keep = engine.make_clp(lp)
    h = C.c_void_p()
    assert L.b200pdlp_form_create(C.byref(clp), scaling, C.byref(h)) == 0
    dims, sc = (C.c_int32 * 5)(), (C.c_double * 3)()
    L.b200pdlp_form_dims(h, dims, sc)
    n, m, nnz, neq, n0 = list(dims)
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    cbeg, cidx, cval = np.zeros(n + 1, np.int32), np.zeros(max(nnz, 1), np.int32), np.zeros(max(nnz, 1))
    L.b200pdlp_form_get_csc(h, cbeg.ctypes.data_as(ip), cidx.ctypes.data_as(ip), cval.ctypes.data_as(dp))
    rptr, rcol, rval = np.zeros(m + 1, np.int32), np.zeros(max(nnz, 1), np.int32), np.zeros(max(nnz, 1))
    L.b200pdlp_form_get_csr(h, rptr.ctypes.data_as(ip), rcol.ctypes.data_as(ip), rval.ctypes.data_as(dp))
    new_idx, cls = np.zeros(max(m, 1), np.int32), np.zeros(max(m, 1), np.int32)
    L.b200pdlp_form_get_row_map(h, new_idx.ctypes.data_as(ip), cls.ctypes.data_as(ip))
    L.b200pdlp_form_destroy(h)
    return dict(n=n, m=m, nnz=nnz, neq=neq, n0=n0, cbeg=cbeg, cidx=cidx[:nnz], cval=cval[:nnz], rptr=rptr, rcol=rcol[:nnz],
                rval=rval[:nnz], new_idx=new_idx[:m], cls=cls[:m])


def model_formulate(lp):
    """classify_rows_kernel, row_maps_kernel, nnz_flags_kernel + scan, scatter_entries_kernel"""
    a = lp.a_matrix_
    n0, m = lp.num_col_, lp.num_row_
    nnz0 = a.numNz()
    rl, ru = lp.row_lower_, lp.row_upper_
    lo, up = rl > -1e20, ru < 1e20
    cls = np.where(lo & up & (rl == ru), 0, np.where(lo & ~up, 2, np.where(~lo & up, 1, 3))).astype(np.int32)
    eql = ((cls == 0) | (cls == 3)).astype(np.int64)
    eq_ex = np.concatenate([[0], np.cumsum(eql)])           # exclusive scan with the total at [m]
    bd_ex = np.concatenate([[0], np.cumsum(cls == 3)])
    neq, nbound = int(eq_ex[m]), int(bd_ex[m])
    i = np.arange(m)
    new_idx = np.where(eql == 1, eq_ex[:m], neq + (i - eq_ex[:m])).astype(np.int32)
    start = a.start_.astype(np.int64)
    colof = np.repeat(np.arange(n0), np.diff(start))
    idx, val = a.index_[:nnz0].astype(np.int64), a.value_[:nnz0]
    fl = eql[idx]
    fl_ex = np.concatenate([[0], np.cumsum(fl)])
    p = np.arange(nnz0)
    s, e = start[colof], start[colof + 1]
    ne = fl_ex[e] - fl_ex[s]
    k = fl_ex[p] - fl_ex[s]
    q = np.where(fl == 1, s + k, s + ne + (p - s - k))
    assert len(np.unique(q)) == nnz0                        # a permutation
    cidx = np.zeros(nnz0 + nbound, np.int32)
    cval = np.zeros(nnz0 + nbound)
    cidx[q] = new_idx[idx]
    cval[q] = np.where(cls[idx] == 1, -val, val)
    b = np.flatnonzero(cls == 3)
    cidx[nnz0 + bd_ex[b]] = new_idx[b]
    cval[nnz0 + bd_ex[b]] = -1.0
    cbeg = np.concatenate([start[:n0], nnz0 + np.arange(nbound), [nnz0 + nbound]]).astype(np.int32)
    return dict(n=n0 + nbound, m=m, nnz=nnz0 + nbound, neq=neq, cls=cls, new_idx=new_idx, cbeg=cbeg, cidx=cidx, cval=cval)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-5] for f in FILES])
def test_scan_partition_formulation_equals_host(engine_lib, path):
    from highs_b200.lp import read_b2lp
    lp = read_b2lp(path)
    if lp.a_matrix_.numNz() == 0:
        pytest.skip("empty matrix")
    h, d = host_form(lp), model_formulate(lp)
    for k in ("n", "m", "nnz", "neq"):
        assert h[k] == d[k], k
    assert np.array_equal(h["cls"], d["cls"]) and np.array_equal(h["new_idx"], d["new_idx"])
    assert np.array_equal(h["cbeg"], d["cbeg"]) and np.array_equal(h["cidx"], d["cidx"]) and np.array_equal(h["cval"], d["cval"])
    # transposition: stable sort of the positions by row == the host's counting-sort CSR
    rpos = np.argsort(d["cidx"], kind="stable")
    colof = np.repeat(np.arange(d["n"]), np.diff(d["cbeg"]))
    assert np.array_equal(colof[rpos], h["rcol"]) and np.array_equal(d["cval"][rpos], h["rval"])
    assert np.array_equal(np.concatenate([[0], np.cumsum(np.bincount(d["cidx"], minlength=d["m"]))]), h["rptr"])


def test_second_sort_pass_orders_unsorted_columns_by_row():
    """positions sorted by row (stable), then by column (stable) == every column's entries in ascending-row order, storage
    order among equal rows: what build_col_major's per-column stable_sort produces on the host"""
    rng = np.random.default_rng(0)
    n, m = 300, 200
    counts = rng.integers(0, 12, n)
    cbeg = np.concatenate([[0], np.cumsum(counts)])
    cidx = rng.integers(0, m, cbeg[-1])                      # unsorted, with duplicates
    colof = np.repeat(np.arange(n), counts)
    rpos = np.argsort(cidx, kind="stable")
    cpos = rpos[np.argsort(colof[rpos], kind="stable")]
    for j in range(n):
        seg = np.arange(cbeg[j], cbeg[j + 1])
        want = seg[np.argsort(cidx[seg], kind="stable")]
        assert np.array_equal(cpos[cbeg[j]:cbeg[j + 1]], want)


def host_make_perm(lens, boundary, window=8192):
    """host_prep.cpp::make_perm: stable sort by DESCENDING length inside windows that do not straddle `boundary`"""
    perm = np.arange(len(lens))
    for b, e in ((0, boundary), (boundary, len(lens))):
        for w in range(b, e, window):
            we = min(w + window, e)
            perm[w:we] = w + np.argsort(-lens[w:we], kind="stable")
    return perm


@pytest.mark.parametrize("nrows,boundary", [(30000, 0), (30000, 30000), (50000, 12345), (100, 50), (8192, 8192), (8193, 1)])
def test_window_key_sort_equals_make_perm(nrows, boundary):
    rng = np.random.default_rng(nrows + boundary)
    lens = rng.integers(0, 40, nrows)
    lens[rng.integers(0, nrows, 5)] = 70000                  # a few very long rows
    W = 8192
    i = np.arange(nrows)
    w0 = (boundary + W - 1) // W
    wid = np.where(i < boundary, i // W, w0 + (i - boundary) // W)
    keys = (wid.astype(np.uint64) << np.uint64(32)) | (np.uint64(0x7fffffff) - lens.astype(np.uint64))
    perm = np.argsort(keys, kind="stable")                   # = stable radix sort of (key, index) pairs
    assert np.array_equal(perm, host_make_perm(lens, boundary, W))


def test_slice_plan_model_matches_host_layout(engine_lib):
    """slice_plan_kernel / long_plan_kernel arithmetic (numpy) against the host plan through b200pdlp_form_layout_eval's stats"""
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(20000, 15000, 6, seed=3, dense_col_nnz=5000)
    h = host_form(lp, scaling=1)
    L = engine.lib()
    clp, keep = engine.make_clp(lp)
    fh = C.c_void_p()
    assert L.b200pdlp_form_create(C.byref(clp), 1, C.byref(fh)) == 0
    x, y = np.ones(h["n"]), np.ones(h["m"])
    ax, aty, stats = np.zeros(h["m"]), np.zeros(h["n"]), np.zeros(12)
    dp = C.POINTER(C.c_double)
    assert L.b200pdlp_form_layout_eval(fh, 0, 1, -1, x.ctypes.data_as(dp), y.ctypes.data_as(dp), ax.ctypes.data_as(dp),
                                       aty.ctypes.data_as(dp), stats.ctypes.data_as(dp)) == 0
    L.b200pdlp_form_destroy(fh)

    def plan(lens, boundary):
        perm = host_make_perm(lens, boundary)
        ln = lens[perm]
        nsl = (len(ln) + 31) // 32
        pad = np.zeros(nsl * 32, np.int64)
        pad[:len(ln)] = np.where(ln > 512, 0, ln)            # long rows and lanes beyond the last row are masked
        padded = int((pad.reshape(nsl, 32).max(axis=1) * 32).sum())
        long_l = ln[ln > 512]
        nseg = int(((long_l + 2047) // 2048).sum())
        return padded, len(long_l), nseg
    a_pad, a_long, a_seg = plan(np.diff(h["rptr"]).astype(np.int64), h["neq"])
    t_pad, t_long, t_seg = plan(np.diff(h["cbeg"]).astype(np.int64), h["n"])
    assert (a_pad, a_long, a_seg) == (stats[3], stats[4], stats[5])
    assert (t_pad, t_long, t_seg) == (stats[6], stats[7], stats[8])


def model_tile_window(cols, vals, ncols, max_window=24576):
    """tile_window_kernel (device_prep.cu) in numpy: the window [lo, lo + w) of the input vector that a tile's real entries touch --
    lo even and w even (16-byte granules of the bulk copy), inside the vector, w = 0 when it does not fit the staging buffer."""
    real = (cols != 0) | (vals != 0.0)          # padding entries of the layout are (column 0, value 0)
    if not real.any():
        return 0, 0
    mn, mx = int(cols[real].min()), int(cols[real].max())
    lo = mn & ~1
    w = (mx - lo + 2) & ~1
    if lo + w > ncols:
        w = 0 if (ncols & 1) else ncols - lo
    if w > max_window or w <= 0:
        return 0, 0
    return lo, w


@pytest.mark.parametrize("seed", range(6))
def test_tile_window_model_covers_every_real_entry(seed):
    """every real entry of a tile lies inside [lo, lo + w), the window is 16-byte aligned in position and size and never runs
    past the vector; tiles that touch both ends of the vector (wrap-around patterns) or more than the buffer get w = 0"""
    rng = np.random.default_rng(seed)
    ncols = int(rng.integers(9000, 40000)) | (seed & 1)          # odd and even vector lengths
    centre = int(rng.integers(0, ncols))
    half = int(rng.integers(1, 14000))
    cols = np.clip(centre + rng.integers(-half, half + 1, size=4096), 0, ncols - 1).astype(np.int64)
    vals = rng.standard_normal(4096)
    pad = rng.random(4096) < 0.2
    cols[pad], vals[pad] = 0, 0.0
    lo, w = model_tile_window(cols, vals, ncols)
    real = (cols != 0) | (vals != 0.0)
    if w == 0:
        span = cols[real].max() - (cols[real].min() & ~1) + 1
        assert span > 24576 - 2 or ((ncols & 1) and (cols[real].max() >= ncols - 1)), (span, ncols)
        return
    assert lo % 2 == 0 and w % 2 == 0 and 0 < w <= 24576 and lo + w <= ncols
    assert ((cols[real] >= lo) & (cols[real] < lo + w)).all()
    # a padded lane reads outside the window or inside it: either way it multiplies a zero value
    assert (vals[~real] == 0.0).all()
