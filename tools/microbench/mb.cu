// Dev microbenchmarks (not product): grid-barrier variants and packed-fp32 pipe throughput on B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void bar_fenced(unsigned* c, unsigned& target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c), "r"(1u) : "memory");
    unsigned seen;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(c) : "memory"); } while ((int)(seen - target) < 0);
    __threadfence();
  }
  __syncthreads();
}
__device__ __forceinline__ void bar_lean(unsigned* c, unsigned& target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c), "r"(1u) : "memory");
    unsigned seen;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(c) : "memory"); } while ((int)(seen - target) < 0);
  }
  __syncthreads();
}
// hierarchical: CTAs arrive on one of 8 group counters, last arriver of a group bumps the root
__device__ __forceinline__ void bar_relaxed_poll(unsigned* c, unsigned& target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c), "r"(1u) : "memory");
    unsigned seen;
    do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(c) : "memory"); } while ((int)(seen - target) < 0);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}
template <int V>
__global__ void k_bar(unsigned* c, int n, float* sink) {
  unsigned target = 0;
  cg::grid_group g = cg::this_grid();
  float v = threadIdx.x;
  for (int i = 0; i < n; ++i) {
    if (V == 0) bar_fenced(c, target);
    if (V == 1) bar_lean(c, target);
    if (V == 2) g.sync();
    if (V == 3) bar_relaxed_poll(c, target);
    v = v * 1.0001f + 1.0f;
  }
  if (v == 12345.f) *sink = v;
}

// pipe throughput: R independent chains per thread
template <int MODE>
__global__ void k_pipe(float* out, int n, float a, float b) {
  float x[8]; unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x[i]), "f"(x[i] + 0.5f)); }
  unsigned long long pa, pb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) x[i] = fmaf(x[i], a, b);                                                   // FFMA
      if (MODE == 1) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));       // FFMA2
      if (MODE == 2) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));                    // FADD2
      if (MODE == 3) x[i] = fminf(fmaxf(x[i], a), b);                                           // 2 x FMNMX
      if (MODE == 4) x[i] = __fadd_rn(x[i], b);                                                 // FADD
      if (MODE == 5) { x[i] = fmaf(x[i], a, b); x[i] = fmaxf(x[i], a); }                        // FFMA + FMNMX (dual pipe)
      if (MODE == 6) { asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); x[i] = fmaxf(x[i], a); }  // FFMA2 + FMNMX
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += x[i] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
int run_bar(unsigned* c, float* sink, int grid, int threads, int n) {
  CK(cudaMemset(c, 0, 4));
  void* args[] = {&c, &n, &sink};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaMemset(c, 0, 4));
    cudaEventRecord(e0);
    CK(cudaLaunchCooperativeKernel((const void*)k_bar<V>, dim3(grid), dim3(threads), args, 0, 0));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("barrier variant %d grid %3d threads %3d: %.3f us per barrier\n", V, grid, threads, ms * 1e3 / n);
  return 0;
}
template <int MODE>
int run_pipe(float* out, const char* name, int ops_per_inst, int inst_per_iter) {
  int n = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k_pipe<MODE><<<148, 512>>>(out, n, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst = 148.0 * 16 * 8.0 * inst_per_iter * n;  // warp-instructions
  double per_sm_clk = inst / 148 / (ms * 1e-3 * 1.965e9);
  printf("%-16s %.3f ms: %.2f warp-inst/clk/SM (%.1f lane-ops/clk/SM)\n", name, ms, per_sm_clk, per_sm_clk * 32 * ops_per_inst);
  return 0;
}
int main() {
  unsigned* c; float* sink; float* out;
  CK(cudaMalloc(&c, 256)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&out, 148 * 512 * 4));
  const int n = 2000;
  for (int grid : {148, 64, 16, 4}) {
    run_bar<0>(c, sink, grid, 512, n); run_bar<1>(c, sink, grid, 512, n); run_bar<2>(c, sink, grid, 512, n); run_bar<3>(c, sink, grid, 512, n);
  }
  run_pipe<0>(out, "FFMA", 1, 1); run_pipe<1>(out, "FFMA2", 2, 1); run_pipe<2>(out, "FADD2", 2, 1); run_pipe<3>(out, "FMNMX x2", 1, 2);
  run_pipe<4>(out, "FADD", 1, 1); run_pipe<5>(out, "FFMA+FMNMX", 1, 2); run_pipe<6>(out, "FFMA2+FMNMX", 1, 2);
  return 0;
}
