// highs_b200/csrc/setup_kernels.hpp -- device-side PDHG_Scale_Data (see setup_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace b200 {

// the standard form's column-major matrix and vectors, resident on the device (standard-form order)
struct DevForm {
  int n, m, nnz;
  const int* cbeg;   // [n+1]
  const int* cidx;   // [nnz] row of each nonzero
  int* colof;        // [nnz] column of each nonzero (filled by device_scale_ruiz)
  double* cval;      // [nnz] scaled in place
  double *cost, *lower, *upper, *colscale;   // [n]
  double *rhs, *rowscale;                    // [m]
};
struct DevScaleScratch {
  double *cs, *cnorm;   // [n]
  double *rs, *rnorm;   // [m]
  double* amax;         // [1] max |a_ij| of the scaled matrix
};

// the 10 Ruiz passes (cupdlp_scaling.c:47-120); needs nothing but the column-major matrix
// have_colof: F.colof is filled already
void device_scale_ruiz(cudaStream_t s, const DevForm& F, DevScaleScratch& w, bool have_colof = false);
// *amax = max |val[p]|
void device_abs_max(cudaStream_t s, int nnz, const double* val, double* amax);
// the Pock-Chambolle pass (:174-231); rptr[m+1] / rpos[nnz] = device copy of the row-major index of the nonzeros
void device_scale_pock_chambolle(cudaStream_t s, const DevForm& F, DevScaleScratch& w, const int* rptr, const int* rpos);

// where a sliced-ELL fill reads a row's entries from: entries [beg[r], end[r]) of the OLD row r, optionally through an
// index (pos), column id = colmap[idx[.] - idx_offset], value = val[.]
struct SellSource {
  const int* beg;
  const int* end;
  const int* pos;      // nullptr: direct
  const int* idx;
  const double* val;
  int idx_offset;
  const int* colmap;
};
// fills col/val[0 .. padded + 32) and, for n_long > 0, lcol/lval (which must be zero-initialised) from the plan's
// descriptors (all pointers are device pointers; perm[new row] = old row)
void device_fill_sell(cudaStream_t s, int nrows, int nslices, const int4* slices, const int* perm, const SellSource& S,
                      int* col, double* val, long long padded, int n_long, const int4* long_rows, const int4* segs,
                      int* lcol, double* lval);

}  // namespace b200
