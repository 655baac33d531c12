// Dev tool: cycle breakdown of the tcgen05 tile pipeline (compiled with -DADMMQ_TC_PROFILE).
#include <cstdio>
#include <cstdlib>
#include <vector>
#define ADMMQ_TC_PROFILE 1
#include "../../admm-quantization_b200/csrc/tc_gemm.cu"
namespace admmq { char* error_buffer() { static char b[8]; return b; } int fail(int c, const char*, ...) { return c; } void count_launches(int) {} int device_props(DeviceProps*) { return 0; } }
struct Maps3 { CUtensorMap a, b, blo; };
template <int BN, bool PS>
__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ Maps3 maps, int M, int N, int K, int bn, float* C, int ldc, long long* dbg, float nzv) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  __shared__ tc::Pipe pipe;
  tc::PipeState st;
  tc::pipe_setup(pipe, st, nzv);
  const long long t_begin = clock64();
  const int tilesM = (M + 127) / 128, tilesN = (N + bn - 1) / bn;
  long long epi = 0;
  for (int tile = blockIdx.x; tile < tilesM * tilesN; tile += gridDim.x) {
    const int i0 = (tile / tilesN) * 128, n0 = (tile % tilesN) * bn;
    const int next = tile + (int)gridDim.x; const bool has_next = next < tilesM * tilesN;
    tc::tile_3xtf32<BN, PS>(&maps.a, i0, &maps.b, &maps.blo, n0, bn, K, smem_dyn, pipe, st, has_next ? (next / tilesN) * 128 : -1, has_next ? (next % tilesN) * bn : -1);
    const long long e0 = clock64();
    const float* tile_c = tc::acc_to_smem<BN, PS>(pipe, smem_dyn);
    using ET = tc::EpiTile<BN>;
    for (int g0 = 0; g0 < ET::kGroups; g0 += 512) {
      const int g = g0 + (int)threadIdx.x;
      const int row = g / ET::kGroupsPerRow, c4 = (g - row * ET::kGroupsPerRow) * 4;
      if (g < ET::kGroups && c4 < bn && i0 + row < M && n0 + c4 < N) {
        const float4 h4 = *reinterpret_cast<const float4*>(tile_c + ET::offset(row, c4 >> 2));
        float* dst = C + (size_t)(i0 + row) * ldc + n0 + c4;
        const float h[4] = {h4.x, h4.y, h4.z, h4.w};
        for (int q = 0; q < 4; ++q) if (n0 + c4 + q < N) dst[q] = h[q];
      }
    }
    __syncthreads();
    epi += clock64() - e0;
  }
  const long long total = clock64() - t_begin;
  if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 480)) {
    long long* d = dbg + (threadIdx.x == 0 ? 0 : 8);
    for (int i = 0; i < 6; ++i) d[i] = st.cyc[i];
    d[6] = epi; d[7] = total;
  }
  tc::pipe_teardown(pipe);
}
template <int BN, bool PS> void run(int M, int N, int K, int gmax = 148, int bn = BN) {
  float *A, *B, *C; long long* dbg;
  cudaMalloc(&A, (size_t)M * K * 4); cudaMalloc(&B, (size_t)N * K * 4); cudaMalloc(&C, (size_t)M * N * 4); cudaMalloc(&dbg, 128);
  cudaMemset(A, 0, (size_t)M * K * 4); cudaMemset(B, 0, (size_t)N * K * 4);
  const int smem = tc::TileSmem<BN, PS>::kBytes;
  cudaFuncSetAttribute(k<BN, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int tiles = ((M + 127) / 128) * ((N + bn - 1) / bn);
  const int grid = tiles < gmax ? tiles : gmax;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  Maps3 maps; tc::make_operand_tmap(&maps.a, A, M, K, K, 128); tc::make_operand_tmap(&maps.b, B, N, K, K, bn); tc::make_operand_tmap(&maps.blo, B, N, K, K, bn);
  for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); k<BN, PS><<<grid, 512, smem>>>(maps, M, N, K, bn, C, N, dbg, -0.0f); cudaEventRecord(e1); cudaDeviceSynchronize(); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[16]; cudaMemcpy(h, dbg, 128, cudaMemcpyDeviceToHost);
  const int nkb = (K + 63) / 64; const int tiles_cta0 = (tiles + grid - 1) / grid;
  printf("BN=%d bn=%d PS=%d grid=%d M=%d N=%d K=%d: %.1f us (%s), CTA0: %d tiles x %d K-blocks, total %lld cyc\n", BN, bn, (int)PS, grid, M, N, K, ms * 1e3, cudaGetErrorString(cudaGetLastError()), tiles_cta0, nkb, h[7]);
  printf("   producer(thread 0): cp.async wait %lld, stage_free wait %lld, convert+issue %lld, tile_done wait %lld, epilogue %lld\n", h[0], h[1], h[2], h[5], h[6]);
  printf("   mma warp (lane 0) : A-full wait %lld, B-full wait %lld, issue %lld, tile_done wait %lld   => per K-block: full-wait %.0f issue %.0f\n", h[8 + 3], h[8 + 0], h[8 + 4], h[8 + 5], (double)h[11] / (tiles_cta0 * nkb), (double)h[12] / (tiles_cta0 * nkb));
}
int main(int argc, char** argv) { int c = argc > 1 ? atoi(argv[1]) : 0; if (c == 0) { run<64, true>(512, 1141, 1144, 32, 48); run<128, true>(512, 1141, 1144, 36, 128); run<128, true>(512, 1141, 1144, 32, 80); run<128, true>(512, 1141, 1144, 24, 96); } if (c == 1) run<128, true>(256, 566, 568, 7, 96); if (c == 2) run<64, true>(64, 134, 136, 1, 48); if (c == 3) run<128, true>(4096, 4096, 4096, 148, 128); if (c == 4) run<64, false>(4096, 4096, 4096); if (c == 5) run<64, true>(4608, 1141, 512); if (c == 6) run<128, true>(4608, 1141, 512, 148, 128); return 0; }
