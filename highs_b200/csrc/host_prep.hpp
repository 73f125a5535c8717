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
};

// Row-major matrix cut into blocks of consecutive rows holding <= kNnzPerBlock
// nonzeros; a row longer than that is split into segment blocks whose partial
// sums are combined by the last segment to finish.
struct BlockedCsr {
  int nrows = 0, ncols = 0, nnz = 0;
  std::vector<int> rowptr;       // [nrows+1]
  std::vector<int> col;          // [nnz padded to a multiple of 4, +4]
  std::vector<double> val;
  struct Block { int row_begin, row_end, nnz_begin, nnz_end; };
  std::vector<Block> blocks;
  std::vector<int> block_long;   // per block: long-row id or -1
  struct LongRow { int row, first_block, nseg, partial_offset; };
  std::vector<LongRow> long_rows;
  int n_partials = 0;
};

constexpr int kNnzPerBlock = 2048;
constexpr int kMaxRowsPerBlock = 1024;

void formulate(const b200pdlp_lp& lp, StdForm& f);
void scale(StdForm& f, bool do_scale);
// nnz-balanced contiguous partition of the m rows into `world` parts
std::vector<int> partition_rows(const StdForm& f, int world);
// rows [r0,r1) of A, row-major (columns ascending within a row)
void build_row_major(const StdForm& f, int r0, int r1, BlockedCsr& a);
// transpose of rows [r0,r1): n rows, (r1-r0) columns, local row ids ascending
void build_col_major(const StdForm& f, int r0, int r1, BlockedCsr& at);

}  // namespace b200
