// tools/dsmem_gather_bench.cu -- experiment for the SpMV gather ceiling (DESIGN.md section 6 / VERDICT r1 weak item 5).
// The fused SpMV kernels are bound by L2->SM sector traffic: every fp64 gather of x moves a 32-byte sector for 8 useful
// bytes (tools/gather_bench.cu: ~245 G gathers/s whatever the load path).  Question: is a random 8-byte gather from
// DISTRIBUTED SHARED MEMORY (x spread over the shared memory of a thread-block cluster, ld.shared::cluster) faster?
// Prints G gathers/s for: global/L2 gathers, DSMEM gathers with cluster sizes 8 and 16, and a 50/50 mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dsmem_gather_bench tools/dsmem_gather_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

namespace cg = cooperative_groups;

#define OK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) { printf("%s failed: %s\n", #call, cudaGetErrorString(e_)); exit(1); } \
  } while (0)

constexpr int kThreads = 1024;

// mode 0: every gather from global memory (L2-resident x)
// mode 1: every gather from the cluster's distributed shared memory
// mode 2: even gathers DSMEM, odd gathers global
template <int MODE>
__global__ void __launch_bounds__(kThreads)
gather_kernel(const double* __restrict__ x, int per_cta, const int* __restrict__ idx, long long per_cluster, double* __restrict__ out) {
  extern __shared__ double sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank(), csize = cluster.num_blocks();
  const long long cluster_id = blockIdx.x / csize;
  if (MODE != 0) {
    for (int i = threadIdx.x; i < per_cta; i += kThreads) sm[i] = x[(size_t)rank * per_cta + i];
    cluster.sync();
  }
  const int* my = idx + cluster_id * per_cluster;
  double acc = 0.0;
  const long long stride = (long long)csize * kThreads;
  for (long long g = (long long)rank * kThreads + threadIdx.x; g + 7 * stride < per_cluster; g += 8 * stride) {
    int c[8];
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) c[u] = my[g + u * stride];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const bool dsm = MODE == 1 || (MODE == 2 && (u & 1) == 0);
      if (dsm) {
        const unsigned r = (unsigned)c[u] / (unsigned)per_cta, off = (unsigned)c[u] - r * (unsigned)per_cta;
        const double* remote = cluster.map_shared_rank(sm, r);
        v[u] = remote[off];
      } else {
        v[u] = x[c[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < 8; u++) acc += v[u];
  }
  if (MODE != 0) cluster.sync();   // nobody leaves while its shared memory may still be read
  if (acc == 123.456) out[blockIdx.x] = acc;
}

template <int MODE>
static double run(int csize, int nclusters, const double* x, int per_cta, const int* idx, long long per_cluster, double* out, int reps) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(csize * nclusters));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (size_t)per_cta * sizeof(double);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  OK(cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));
  OK(cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  int max_clusters = 0;
  cudaOccupancyMaxActiveClusters(&max_clusters, gather_kernel<MODE>, &cfg);
  cudaEvent_t e0, e1;
  OK(cudaEventCreate(&e0)); OK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; w++) OK(cudaLaunchKernelEx(&cfg, gather_kernel<MODE>, x, per_cta, idx, per_cluster, out));
  OK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) OK(cudaLaunchKernelEx(&cfg, gather_kernel<MODE>, x, per_cta, idx, per_cluster, out));
  OK(cudaEventRecord(e1));
  OK(cudaEventSynchronize(e1));
  float ms = 0;
  OK(cudaEventElapsedTime(&ms, e0, e1));
  const double gathers = (double)per_cluster * nclusters * reps;
  printf("  mode %d cluster %2d x %2d clusters (max co-resident %d): %8.1f us per launch, %7.1f G gathers/s\n", MODE, csize, nclusters,
         max_clusters, 1e3 * ms / reps, gathers / (ms * 1e6));
  return gathers / (ms * 1e6);
}

int main() {
  const long long total = 8LL << 20;   // gathers per launch (~ nnz of S3)
  for (int csize : {8, 16}) {
    const int per_cta = 24 * 1024;                  // 192 KiB of x per CTA
    const int nx = per_cta * csize;                 // 1.5 M / 3 M doubles = 12 / 24 MiB... (x of S3 is 8 MiB)
    const int nclusters = 148 / csize;
    const long long per_cluster = total / nclusters / (8LL * csize * kThreads) * (8LL * csize * kThreads);
    std::vector<double> hx(nx, 1.0);
    std::vector<int> hidx((size_t)per_cluster * nclusters);
    std::mt19937 rng(7);
    for (auto& v : hidx) v = (int)(rng() % (unsigned)nx);
    double *dx, *dout;
    int* didx;
    OK(cudaMalloc(&dx, nx * sizeof(double))); OK(cudaMalloc(&dout, 4096 * sizeof(double)));
    OK(cudaMalloc(&didx, hidx.size() * sizeof(int)));
    OK(cudaMemcpy(dx, hx.data(), nx * sizeof(double), cudaMemcpyHostToDevice));
    OK(cudaMemcpy(didx, hidx.data(), hidx.size() * sizeof(int), cudaMemcpyHostToDevice));
    printf("cluster size %d: x = %d doubles (%.1f MiB) per cluster, %lld gathers per launch\n", csize, nx, nx * 8.0 / (1 << 20),
           per_cluster * nclusters);
    run<0>(csize, nclusters, dx, per_cta, didx, per_cluster, dout, 20);
    run<1>(csize, nclusters, dx, per_cta, didx, per_cluster, dout, 20);
    run<2>(csize, nclusters, dx, per_cta, didx, per_cluster, dout, 20);
    cudaFree(dx); cudaFree(dout); cudaFree(didx);
  }
  return 0;
}
