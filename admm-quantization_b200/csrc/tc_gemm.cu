// tc_gemm.cu - C = A . B^T in 3xTF32 on tcgen05 (admmq_gemm_nt), the stand-alone form of the tile product that
// the persistent ADMM loop uses for its ridge product and that the tensor-core MTTKRP builds on.
#include <algorithm>
#include "tc_gemm.cuh"

namespace admmq {

template <int BN>
__global__ void __launch_bounds__(tc::kThreadsTC, 1)
k_gemm_nt_tc(const float* __restrict__ A, int lda, int M, const float* __restrict__ B, int ldb, int N, int K,
             float* __restrict__ C, int ldc) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  __shared__ tc::Pipe pipe;
  tc::PipeState st;
  constexpr unsigned int kCols = BN < 32 ? 32 : BN;
  tc::pipe_setup(pipe, st, kCols);
  const int tilesM = (M + tc::kTileM - 1) / tc::kTileM, tilesN = (N + BN - 1) / BN;
  for (int tile = blockIdx.x; tile < tilesM * tilesN; tile += gridDim.x) {
    const int i0 = (tile / tilesN) * tc::kTileM, n0 = (tile % tilesN) * BN;
    tc::tile_3xtf32<BN>(A, lda, i0, M, B, ldb, n0, N, K, smem_dyn, pipe, st);
    float v[BN / 4];
    int row, col0;
    tc::load_acc<BN>(pipe, v, row, col0);
    if (i0 + row < M) {
#pragma unroll
      for (int i = 0; i < BN / 4; ++i)
        if (n0 + col0 + i < N) C[(size_t)(i0 + row) * ldc + n0 + col0 + i] = v[i];
    }
    tc::release_acc();
  }
  tc::pipe_teardown(pipe, kCols);
}

}  // namespace admmq

using namespace admmq;

extern "C" int admmq_gemm_nt(const float* A, int lda, int M, const float* B, int ldb, int N, int K, float* C, int ldc,
                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (A == nullptr || B == nullptr || C == nullptr || M <= 0 || N <= 0 || K <= 0)
    return fail(ADMMQ_E_BADARG, "admmq_gemm_nt: null pointer or empty shape");
  if ((lda & 3) || (ldb & 3) || lda < K || ldb < K || ldc < N || ((uintptr_t)A & 15) || ((uintptr_t)B & 15))
    return fail(ADMMQ_E_BADARG, "admmq_gemm_nt: lda/ldb must be multiples of 4 and >= K, A/B 16-byte aligned, ldc >= N");
  DeviceProps dp;
  if (int e = device_props(&dp)) return e;
  if (dp.cc_major != 10) return fail(ADMMQ_E_UNSUPPORTED, "admmq_gemm_nt needs an sm_100 device (tcgen05)");
  const int tilesM = (M + tc::kTileM - 1) / tc::kTileM;
  // widest tile that still gives every SM work
  const int bn = ((long long)tilesM * ((N + 63) / 64) >= dp.sm_count) ? 64 : ((long long)tilesM * ((N + 31) / 32) >= dp.sm_count ? 32 : 16);
  const int tiles = tilesM * ((N + bn - 1) / bn);
  const int grid = std::min(tiles, dp.sm_count);
  if (bn == 64) {
    const int smem = tc::TileSmem<64>::kBytes;
    ADMMQ_CUDA_OK(cudaFuncSetAttribute(k_gemm_nt_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_gemm_nt_tc<64><<<grid, tc::kThreadsTC, smem, stream>>>(A, lda, M, B, ldb, N, K, C, ldc);
  } else if (bn == 32) {
    const int smem = tc::TileSmem<32>::kBytes;
    ADMMQ_CUDA_OK(cudaFuncSetAttribute(k_gemm_nt_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_gemm_nt_tc<32><<<grid, tc::kThreadsTC, smem, stream>>>(A, lda, M, B, ldb, N, K, C, ldc);
  } else {
    const int smem = tc::TileSmem<16>::kBytes;
    ADMMQ_CUDA_OK(cudaFuncSetAttribute(k_gemm_nt_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_gemm_nt_tc<16><<<grid, tc::kThreadsTC, smem, stream>>>(A, lda, M, B, ldb, N, K, C, ldc);
  }
  ADMMQ_CUDA_OK(cudaGetLastError());
  count_launches(1);
  return ADMMQ_OK;
}
