// Dev tool: per-SM L2 -> shared memory bandwidth of TMA 2-D tile loads (32 KB boxes, 4-deep ring, data L2 resident).
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, int rows_total, int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long full[4];
  const unsigned base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    const int row_blocks = rows_total / 128;
    for (int it = 0; it < iters + 4; ++it) {
      const int d = it & 3;
      if (it >= 4) {  // wait for the previous fill of this slot
        const unsigned par = ((it >> 2) - 1) & 1u;
        unsigned ok;
        do {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&full[d])), "r"(par) : "memory");
        } while (!ok);
      }
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[d])), "r"(32768u) : "memory");
        const int rb = (blockIdx.x * 7 + it) % row_blocks;
        for (int a = 0; a < 2; ++a)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(base + d * 32768 + a * 16384), "l"(&tm), "r"(((it * 2 + a) % 16) * 32), "r"(rb * 128), "r"(smem_u32(&full[d])) : "memory");
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
int main() {
  const int rows = 4096, cols = 1152;
  float* A; cudaMalloc(&A, (size_t)rows * cols * 4); cudaMemset(A, 0, (size_t)rows * cols * 4);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tm; cuuint64_t gd[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t gs[1] = {(cuuint64_t)cols * 4}; cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
  ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 1024);
  for (int grid : {1, 8, 32, 148}) {
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) k<<<grid, 128, 4 * 32768 + 1024>>>(tm, rows, iters, out);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("grid %3d: %.1f B/clk per SM (%s), aggregate %.1f TB/s at 1.965 GHz\n", grid, 32768.0 * iters / mx, cudaGetErrorString(cudaGetLastError()), 32768.0 * iters / mx * grid * 1.965e9 / 1e12);
  }
  return 0;
}
