// tools/gather_bench.cu -- microbenchmark: how fast can one B200 gather 8-byte elements at random
// from an L2-resident 8 MB vector?  (the SpMV bottleneck: L1TEX t-stage, ~2 cycles per distinct line)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

enum Mode { LD = 0, LDG_NC, LD_CG, LD_CV, NC_NOALLOC, TEX, ONELANE, LDGSTS, BULK16, NMODES };
const char* names[] = {"ld.global", "ld.global.nc", "ld.global.cg", "ld.global.cv", "ld.nc.L1::no_allocate", "tex1Dfetch<int2>", "one-lane-per-load", "cp.async 8B->smem", "cp.async.bulk 16B->smem"};

template <int MODE>
__global__ void __launch_bounds__(256) gather_kernel(const int* __restrict__ idx, const double* x, cudaTextureObject_t tex, double* out, int n) {
  __shared__ __align__(16) double stage[256 * 4 * 2];
  __shared__ __align__(8) unsigned long long mbar;
  double acc = 0.0;
  const int stride = gridDim.x * blockDim.x * 4;
  if (MODE == BULK16) {
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&mbar))); }
    __syncthreads();
  }
  unsigned phase = 0;
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) * 4; base < n; base += stride) {
    const int4 c = *reinterpret_cast<const int4*>(idx + base);
    const int cc[4] = {c.x, c.y, c.z, c.w};
    if (MODE == LD) {
#pragma unroll
      for (int k = 0; k < 4; k++) acc += x[cc[k]];
    } else if (MODE == LDG_NC) {
#pragma unroll
      for (int k = 0; k < 4; k++) acc += __ldg(x + cc[k]);
    } else if (MODE == LD_CG) {
#pragma unroll
      for (int k = 0; k < 4; k++) acc += __ldcg(x + cc[k]);
    } else if (MODE == LD_CV) {
#pragma unroll
      for (int k = 0; k < 4; k++) acc += __ldcv(x + cc[k]);
    } else if (MODE == NC_NOALLOC) {
#pragma unroll
      for (int k = 0; k < 4; k++) { double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(x + cc[k])); acc += v; }
    } else if (MODE == TEX) {
#pragma unroll
      for (int k = 0; k < 4; k++) { int2 t = tex1Dfetch<int2>(tex, cc[k]); acc += __hiloint2double(t.y, t.x); }
    } else if (MODE == ONELANE) {
      // each lane's load issued as its own instruction (others predicated off)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        double v = 0.0;
        for (int l = 0; l < 32; l++) if ((threadIdx.x & 31) == l) v = x[cc[k]];
        acc += v;
      }
    } else if (MODE == LDGSTS) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        unsigned d = (unsigned)__cvta_generic_to_shared(&stage[threadIdx.x * 4 + k]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(x + cc[k]));
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int k = 0; k < 4; k++) acc += stage[threadIdx.x * 4 + k];
    } else if (MODE == BULK16) {
      unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
      if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mb), "r"(256 * 4 * 16));
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; k++) {
        unsigned d = (unsigned)__cvta_generic_to_shared(&stage[(threadIdx.x * 4 + k) * 2]);
        const double* src = x + (cc[k] & ~1);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(d), "l"(src), "r"(mb) : "memory");
      }
      // wait
      asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"(mb), "r"(phase) : "memory");
      phase ^= 1;
#pragma unroll
      for (int k = 0; k < 4; k++) acc += stage[(threadIdx.x * 4 + k) * 2 + (cc[k] & 1)];
      __syncthreads();
    }
  }
  if (acc == 123.456) out[0] = acc;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
float run(const int* idx, const double* x, cudaTextureObject_t tex, double* out, int n, int grid) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; i++) gather_kernel<MODE><<<grid, 256>>>(idx, x, tex, out, n);
  CK(cudaEventRecord(a));
  const int reps = 20;
  for (int i = 0; i < reps; i++) gather_kernel<MODE><<<grid, 256>>>(idx, x, tex, out, n);
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main() {
  const int nx = 1 << 20, n = 8 << 20;
  std::vector<int> h(n); std::mt19937 g(1); for (auto& v : h) v = g() % nx;
  std::vector<double> hx(nx, 1.0);
  int* idx; double *x, *out;
  CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&x, nx * 8)); CK(cudaMalloc(&out, 148 * 16 * 256 * 8));
  CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(x, hx.data(), nx * 8, cudaMemcpyHostToDevice));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = x; rd.res.linear.desc = cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes = (size_t)nx * 8;
  cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  for (int grid : {148 * 4, 148 * 8}) {
    printf("grid %d (8M random 8-byte gathers from an 8 MB vector):\n", grid);
    float t;
#define R(M) t = run<M>(idx, x, tex, out, n, grid); printf("  %-28s %8.1f us  %6.1f Gelem/s  %.2f cyc/elem/SM\n", names[M], t * 1e3, n / t / 1e6, t * 1e-3 * 1.965e9 * 148 / n);
    R(LD) R(LDG_NC) R(LD_CG) R(LD_CV) R(NC_NOALLOC) R(TEX) R(ONELANE) R(LDGSTS) R(BULK16)
  }
  // sequential-index control (perfectly coalesced)
  for (int i = 0; i < n; i++) h[i] = i % nx;
  CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
  float t = run<LD>(idx, x, tex, out, n, 148 * 8);
  printf("coalesced control: %8.1f us\n", t * 1e3);
  return 0;
}
