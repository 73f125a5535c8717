"""HighsLp mirror + LP interchange + the synthetic LP generator of SURVEY.md section 8(d).

`HighsLp` carries exactly the members of the reference's `HighsLp` that
`solveLpCupdlp` reads (/root/reference/highs/lp_data/HighsLp.h:23-35 and
highs/pdlp/CupdlpWrapper.cpp:280-300): num_col_, num_row_, col_cost_, col_lower_,
col_upper_, row_lower_, row_upper_, a_matrix_ (column-wise start_/index_/value_),
sense_ and offset_.  Names keep the reference's trailing underscore so that code
and tests read like the reference's own.
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

kHighsInf = float("inf")
B2LP_MAGIC = 0x504C3242  # "B2LP"


@dataclasses.dataclass
class HighsSparseMatrix:
    """Column-wise sparse matrix (highs/util/HighsSparseMatrix.h:29-37)."""
    num_col_: int
    num_row_: int
    start_: np.ndarray  # int32 [num_col_+1]
    index_: np.ndarray  # int32 [nnz]
    value_: np.ndarray  # float64 [nnz]

    def numNz(self) -> int:
        return int(self.start_[self.num_col_])

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.value_, self.index_, self.start_), shape=(self.num_row_, self.num_col_))


@dataclasses.dataclass
class HighsLp:
    num_col_: int
    num_row_: int
    col_cost_: np.ndarray
    col_lower_: np.ndarray
    col_upper_: np.ndarray
    row_lower_: np.ndarray
    row_upper_: np.ndarray
    a_matrix_: HighsSparseMatrix
    sense_: int = 1          # ObjSense::kMinimize = 1, kMaximize = -1 (HConst.h)
    offset_: float = 0.0
    model_name_: str = ""

    def __post_init__(self):
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.col_cost_, self.col_lower_, self.col_upper_ = f64(self.col_cost_), f64(self.col_lower_), f64(self.col_upper_)
        self.row_lower_, self.row_upper_ = f64(self.row_lower_), f64(self.row_upper_)
        a = self.a_matrix_
        a.start_ = np.ascontiguousarray(a.start_, dtype=np.int32)
        a.index_ = np.ascontiguousarray(a.index_, dtype=np.int32)
        a.value_ = np.ascontiguousarray(a.value_, dtype=np.float64)
        assert a.start_.shape == (self.num_col_ + 1,)
        assert self.col_cost_.shape == (self.num_col_,) and self.row_lower_.shape == (self.num_row_,)

    def objectiveValue(self, col_value: np.ndarray) -> float:
        """HighsLp::objectiveValue (highs/lp_data/HighsLp.cpp) = offset + c.x"""
        return float(self.offset_ + np.dot(self.col_cost_, col_value))


def write_b2lp(path: str, lp: HighsLp) -> None:
    """Binary interchange file shared with oracle/ref_driver.cpp."""
    a = lp.a_matrix_
    with open(path, "wb") as f:
        f.write(struct.pack("<4q2d", B2LP_MAGIC, lp.num_col_, lp.num_row_, a.numNz(), float(lp.sense_), float(lp.offset_)))
        for v in (lp.col_cost_, lp.col_lower_, lp.col_upper_, lp.row_lower_, lp.row_upper_):
            f.write(v.astype("<f8").tobytes())
        f.write(a.start_.astype("<i4").tobytes())
        f.write(a.index_[: a.numNz()].astype("<i4").tobytes())
        f.write(a.value_[: a.numNz()].astype("<f8").tobytes())


def read_b2lp(path: str) -> HighsLp:
    with open(path, "rb") as f:
        magic, n, m, nnz, sense, offset = struct.unpack("<4q2d", f.read(48))
        if magic != B2LP_MAGIC:
            raise ValueError(f"{path}: not a .b2lp file")
        rd = lambda k, dt: np.frombuffer(f.read(k * np.dtype(dt).itemsize), dtype=dt).copy()
        c, lo, up = rd(n, "<f8"), rd(n, "<f8"), rd(n, "<f8")
        rl, ru = rd(m, "<f8"), rd(m, "<f8")
        start, index, value = rd(n + 1, "<i4"), rd(nnz, "<i4"), rd(nnz, "<f8")
    return HighsLp(n, m, c, lo, up, rl, ru, HighsSparseMatrix(n, m, start, index, value), -1 if sense < 0 else 1, offset)


def synthetic_lp(m: int, n: int, nnz_per_col: int, seed: int = 12345, dense_col_nnz: int = 0, band: int = 0) -> HighsLp:
    """Random sparse LP with a planted strictly-complementary optimal pair.

    min c.x  s.t.  A x >= b, x >= 0.   Each column draws `nnz_per_col` row
    indices uniformly (sorted, de-duplicated), values N(0,1).  x*_j = 0 w.p. 1/2
    else U(0,1); y*_i likewise; b = A x* - r with r_i = 0 where y*_i > 0 else
    U(0,1); c = A'y* + z with z_j = 0 where x*_j > 0 else U(0,1).  Every row is a
    GEQ row, so the cuPDLP standard form adds no slack columns and keeps (m, n).
    `dense_col_nnz` > 0 replaces column 0 by that many distinct random rows (the
    "pathological" configuration S5).  `band` > 0: banded structure instead of uniformly random rows; `band` < 0: the same
    |band|-bounded offsets in every column (multi-diagonal).
    numpy's PCG64 replaces the survey's mt19937_64: the generator defines the
    workload, it is not part of the parity contract.
    """
    rng = np.random.default_rng(seed)
    if band < 0:
        # multi-diagonal variant (bench workload S3D): the same `nnz_per_col` offsets (drawn once from [band, -band]) for every
        # column, like a stencil / staircase LP -- consecutive rows hold consecutive columns, so a warp's 32 gathers fall into
        # a few sectors.  This is the upper end of what sector sharing can give the SpMV kernels.
        offs = rng.integers(band, -band + 1, size=nnz_per_col, dtype=np.int64)
        centre = (np.arange(n, dtype=np.int64) * m) // max(n, 1)
        rows = (centre[:, None] + offs[None, :]) % m
    elif band > 0:
        # structured variant (bench workload S3B): column j's rows lie within `band` of the diagonal position j m / n,
        # like the staircase / block-angular matrices of real LPs -- neighbouring rows and columns share vector entries
        centre = (np.arange(n, dtype=np.int64) * m) // max(n, 1)
        rows = centre[:, None] + rng.integers(-band, band + 1, size=(n, nnz_per_col), dtype=np.int64)
        rows = np.clip(rows, 0, m - 1)
    else:
        rows = rng.integers(0, m, size=(n, nnz_per_col), dtype=np.int64)
    rows.sort(axis=1)
    keep = np.ones_like(rows, dtype=bool)
    keep[:, 1:] = rows[:, 1:] != rows[:, :-1]
    counts = keep.sum(axis=1)
    if dense_col_nnz > 0:
        counts[0] = 0
        keep[0, :] = False
    index = rows[keep].astype(np.int32)
    if dense_col_nnz > 0:
        dense_rows = np.sort(rng.choice(m, size=dense_col_nnz, replace=False)).astype(np.int32)
        index = np.concatenate([dense_rows, index])
        counts[0] = dense_col_nnz
    start = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=start[1:])
    nnz = int(start[n])
    value = rng.standard_normal(nnz)
    xs = np.where(rng.random(n) < 0.5, 0.0, rng.random(n))
    ys = np.where(rng.random(m) < 0.5, 0.0, rng.random(m))
    import scipy.sparse as sp
    A = sp.csc_matrix((value, index, start.astype(np.int32)), shape=(m, n))
    r = np.where(ys > 0, 0.0, rng.random(m))
    z = np.where(xs > 0, 0.0, rng.random(n))
    b = A @ xs - r
    c = A.T @ ys + z
    return HighsLp(n, m, c, np.zeros(n), np.full(n, kHighsInf), b, np.full(m, kHighsInf),
                   HighsSparseMatrix(n, m, start.astype(np.int32), index, value), 1, 0.0,
                   f"synthetic_m{m}_n{n}_k{nnz_per_col}_s{seed}" + (f"_dense{dense_col_nnz}" if dense_col_nnz else "")
                   + (f"_band{band}" if band else ""))
