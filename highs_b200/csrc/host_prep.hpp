// highs_b200/csrc/host_prep.hpp -- host-side preparation of the PDLP standard form.
//
// Product code (C++), no CUDA.  Replaces, for the B200 engine, what the reference
// does on the host before its solver starts:
//   formulateLP_highs            /root/reference/highs/pdlp/CupdlpWrapper.cpp:280-448
//   Init_Scaling/PDHG_Scale_Data highs/pdlp/cupdlp/cupdlp_scaling.c:233-425
//   csc2csr                      highs/pdlp/cupdlp/cupdlp_utils.c:1222-1254
// and then builds the row-blocked layouts the SpMV kernels stream.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/b200pdlp.h"

namespace b200 {

enum RowClass : int { kEq = 0, kLeq = 1, kGeq = 2, kBound = 3 };  // cupdlp_defs.h numbering

// min c'x  s.t.  A x = b (rows < neq),  A x >= b (rows >= neq),  l <= x <= u
struct StdForm {
  int n = 0, m = 0, nnz = 0, neq = 0, n_orig = 0;
  std::vector<double> cost, lower, upper, rhs;
  std::vector<int> cbeg, cidx;   // column-wise copy (scaled in place)
  std::vector<double> cval;
  std::vector<double> col_scale, row_scale;
  std::vector<int> row_new_idx, row_class;  // by ORIGINAL row
  double sense = 1.0, offset = 0.0;
  double norm_cost = 0.0, norm_rhs = 0.0;   // 2-norms of the unscaled cost / rhs
  double amax = 0.0;                        // max |a_ij| after scaling
  // row-major index of the nonzeros (built once, by scale() or on demand): row i owns positions
  // rpos[rptr[i] .. rptr[i+1]) of cidx/cval, columns ascending
  std::vector<int> rptr, rpos;
};

// plain row-major matrix (CSR); entries of a row keep the order the reference's scatter SpMV adds them in
struct Csr {
  int nrows = 0, ncols = 0, nnz = 0;
  std::vector<int> rowptr, col;
  std::vector<double> val;
};

// Device layout of one matrix: sliced ELL body + split long rows.
//  * body: rows (in `perm` order) are grouped in slices of 32; slice s stores its entries k-major /
//    lane-minor (element (k, lane) at ptr + 32 k + lane), padded with (col 0, val 0) to the longest
//    row of the slice, so a warp reads col/val fully coalesced and each lane accumulates ITS row in
//    registers, in the row's own entry order.  Rows are sorted by length inside windows of
//    kSortWindow rows (never across the equality / inequality boundary) to keep the padding small.
//  * rows longer than `long_threshold` are left out of the body (their lanes are masked) and cut into
//    segments of kNnzPerBlock nonzeros; the last segment CTA to finish combines the partial sums.
struct SellMatrix {
  int nrows = 0, ncols = 0;
  long long nnz = 0, padded = 0;
  struct Slice { int ptr, len; unsigned skipmask; int pad; };
  std::vector<Slice> slices;
  std::vector<int> col;
  std::vector<double> val;
  struct Seg { int row, nnz_begin, nnz_end, long_id; };
  struct LongRow { int row, first_seg, nseg, partial_offset; };
  std::vector<Seg> segs;
  std::vector<LongRow> long_rows;
  std::vector<int> lcol;
  std::vector<double> lval;
  int n_partials = 0;
};

constexpr int kNnzPerBlock = 2048;   // long-row segment size (== kernels.cuh kNnzBlk)
constexpr int kSortWindow = 8192;

void formulate(const b200pdlp_lp& lp, StdForm& f);
void build_row_index(StdForm& f);   // fills f.rptr / f.rpos (parallel counting sort)
void scale(StdForm& f, bool do_scale);
// nnz-balanced contiguous partition of the m rows into `world` parts
std::vector<int> partition_rows(const StdForm& f, int world);
// rows [r0,r1) of A, row-major (columns ascending within a row)
void build_row_major(const StdForm& f, int r0, int r1, Csr& a);
// transpose of rows [r0,r1): n rows, (r1-r0) columns, local row ids ascending
void build_col_major(const StdForm& f, int r0, int r1, Csr& at);
// new -> old ordering of `count` rows: identity, or (sort) by descending length inside windows that
// do not straddle `boundary`
std::vector<int> make_perm(const std::vector<int>& rowptr, int boundary, bool sort);
std::vector<int> invert_perm(const std::vector<int>& perm);
// rows of `a` taken in `perm` order, column ids mapped through `colmap` (old -> new)
void build_sell(const Csr& a, const std::vector<int>& perm, const std::vector<int>& colmap, int long_threshold,
                SellMatrix& out);

}  // namespace b200
