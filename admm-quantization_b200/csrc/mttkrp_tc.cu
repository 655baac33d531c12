// mttkrp_tc.cu - MTTKRP in 3xTF32 on the tensor cores (admmq_mttkrp precision 1, admmq_mttkrp_tc).
//
//   F[m, r] = sum_{x, y} Wn[m, x * ny + y] * X[x, r] * Y[y, r]        scripts/factorize.py:217,227,237
//
// The Khatri-Rao operand is never formed.  The contraction over the LARGE index x is a dense GEMM against the factor X
// alone, T[(m, y), r] = sum_x V[(m, y), x] * X[x, r] with V the (m, y, x) permutation of the tensor (constant per layer,
// made once by admmq_permute_myx), run by the tcgen05 tile kernel of tc_gemm.cu; the small index y (the 9 kernel taps of
// a 3x3 convolution) is then folded with a float64-accumulating weighted reduction F[m, r] = sum_y Y[y, r] T[(m, y), r].
// Same flops as the direct form (2 M nx ny R), operands K-major without any on-chip Khatri-Rao generation, and the
// intermediate T is M * ny * R floats (21 MB for 512 x 512 x 9 at R = 1141 instead of the 1.2 GB einsum temporary).
// For a matrix (Y == NULL, ny = 1) it is the plain product W . X (scripts/factorize.py:277,287).
#include <algorithm>
#include "common.cuh"

namespace admmq {

int gemm_nt(const float* A, int lda, int M, const float* B, const float* Blo, int ldb, int N, int K, float* C, int ldc,
            cudaStream_t stream);  // tc_gemm.cu
int mttkrp_fold_gemm(const float* V, int ldv, int Mout, int ny, const float* B, const float* Blo, int ldb, int N, int K,
                     const float* Y, float* F, cudaStream_t stream);  // tc_gemm.cu
size_t mttkrp_foldlong_partial_bytes(int Mout, int ny, int N);
int mttkrp_foldlong_gemm(const float* V, int ldv, int Mout, int ny, const float* B, const float* Blo, int ldb, int N, int K,
                         const float* Y, float* F, double* partial, cudaStream_t stream);  // tc_gemm.cu

constexpr int kTT = 256;

// V[(m, y), x] = Wn[m, x * ny + y]: per m a (nx x ny) -> (ny x nx) transpose, staged through shared memory
__global__ void __launch_bounds__(kTT) k_permute_myx(const float* __restrict__ Wn, int M, int nx, int ny,
                                                    float* __restrict__ V, int ldv) {
  __shared__ float tile[32][33];
  const int m = blockIdx.z;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* src = Wn + (size_t)m * nx * ny;
  for (int r = ty; r < 32; r += 8) {
    const int x = x0 + r, y = y0 + tx;
    tile[r][tx] = (x < nx && y < ny) ? src[(size_t)x * ny + y] : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int y = y0 + r, x = x0 + tx;
    if (y < ny && x < nx) V[((size_t)m * ny + y) * ldv + x] = tile[tx][r];
  }
}

// out[c, r] = in[r, c]  (rows x cols -> cols x rows, leading dimension ldo), zero-filling columns [rows, ldo), split
// into the tf32 hi / lo parts the tensor-core product takes as its pre-split B operand (Veltkamp split as in
// tc_gemm.cuh::split2; neg_zero = -0.0f as a run-time value keeps the product from being contracted away)
__global__ void __launch_bounds__(kTT) k_transpose_split(const float* __restrict__ in, int rows, int cols,
                                                        float* __restrict__ out, float* __restrict__ out_lo, int ldo,
                                                        float neg_zero) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int rr = r0 + r, cc = c0 + tx;
    tile[r][tx] = (rr < rows && cc < cols) ? in[(size_t)rr * cols + cc] : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int cc = c0 + r, rr = r0 + tx;
    if (cc < cols && rr < ldo) {
      const float x = (rr < rows) ? tile[tx][r] : 0.0f;
      const float tt = __fmaf_rn(x, 8193.0f, neg_zero);
      const float hi = __fsub_rn(tt, __fsub_rn(tt, x));
      out[(size_t)cc * ldo + rr] = hi;
      out_lo[(size_t)cc * ldo + rr] = __fsub_rn(x, hi);
    }
  }
}

static size_t tc_layout(int M, int nx, int ny, int R, size_t* xt_off, size_t* t_off) {
  const size_t ldx = (size_t)(nx + 3) / 4 * 4;
  size_t off = 0;
  *xt_off = off;
  off += 2 * align_up((size_t)R * ldx * sizeof(float), 256);  // hi and lo parts of X^T
  *t_off = off;
  // ny > 128: float64 partial sums of the long fold (two per row tile and column); the intermediate T itself is never stored
  if (ny > 128) off += align_up(mttkrp_foldlong_partial_bytes(M, ny, R), 256);
  return off;
}

size_t mttkrp_tc_workspace_bytes(int M, int nx, int ny, int R) {
  size_t a, b;
  return tc_layout(M, nx, std::max(ny, 1), R, &a, &b);
}

// V: (M * ny) x ldv float32, rows (m, y), columns x (ldv = nx rounded up to 4, pad columns zero)
int mttkrp_tc(const float* V, int M, const float* X, int nx, const float* Y, int ny, int R, float* F, void* workspace,
              size_t workspace_bytes, cudaStream_t stream) {
  if (Y == nullptr) ny = 1;
  size_t xt_off, t_off;
  const size_t need = tc_layout(M, nx, ny, R, &xt_off, &t_off);
  if (workspace == nullptr || workspace_bytes < need || ((uintptr_t)workspace & 255) != 0)
    return fail(ADMMQ_E_WORKSPACE, "admmq_mttkrp_tc: workspace needs %zu bytes, 256-byte aligned", need);
  const int ldx = (nx + 3) / 4 * 4;
  float* Xt = (float*)((char*)workspace + xt_off);
  float* XtLo = (float*)((char*)workspace + xt_off + align_up((size_t)R * ldx * sizeof(float), 256));
  double* partial = (double*)((char*)workspace + t_off);
  dim3 tg((R + 31) / 32, (ldx + 31) / 32);
  k_transpose_split<<<tg, kTT, 0, stream>>>(X, nx, R, Xt, XtLo, ldx, -0.0f);  // (R x ldx) = X^T, pad columns zero
  ADMMQ_CUDA_OK(cudaGetLastError());
  count_launches(1);
  if (ny > 1 && ny <= 128) {
    // the fold over the small index happens in the GEMM's epilogue: T is never materialised
    return mttkrp_fold_gemm(V, ldx, M, ny, Xt, XtLo, ldx, R, nx, Y, F, stream);
  }
  if (ny > 128) return mttkrp_foldlong_gemm(V, ldx, M, ny, Xt, XtLo, ldx, R, nx, Y, F, partial, stream);
  return gemm_nt(V, ldx, M, Xt, XtLo, ldx, R, nx, F, R, stream);   // matrix case: F = W . X
}

}  // namespace admmq

using namespace admmq;

extern "C" int admmq_permute_myx(const float* Wn, int M, int nx, int ny, float* V, void* stream_) {
  if (Wn == nullptr || V == nullptr || M <= 0 || nx <= 0 || ny <= 0) return fail(ADMMQ_E_BADARG, "admmq_permute_myx: bad argument");
  const int ldv = (nx + 3) / 4 * 4;
  if (ldv != nx) ADMMQ_CUDA_OK(cudaMemsetAsync(V, 0, (size_t)M * ny * ldv * sizeof(float), (cudaStream_t)stream_));
  dim3 grid((nx + 31) / 32, (ny + 31) / 32, M);
  k_permute_myx<<<grid, kTT, 0, (cudaStream_t)stream_>>>(Wn, M, nx, ny, V, ldv);
  ADMMQ_CUDA_OK(cudaGetLastError());
  count_launches(1);
  return ADMMQ_OK;
}

extern "C" size_t admmq_mttkrp_tc_workspace_bytes(int M, int nx, int ny, int R) { return mttkrp_tc_workspace_bytes(M, nx, ny, R); }

extern "C" int admmq_mttkrp_tc(const float* V, int M, const float* X, int nx, const float* Y, int ny, int R, float* F,
                               void* workspace, size_t workspace_bytes, void* stream_) {
  if (V == nullptr || X == nullptr || F == nullptr || M <= 0 || nx <= 0 || R <= 0 || (Y != nullptr && ny <= 0))
    return fail(ADMMQ_E_BADARG, "admmq_mttkrp_tc: bad argument");
  if (((uintptr_t)V & 15) != 0) return fail(ADMMQ_E_BADARG, "admmq_mttkrp_tc: V must be 16-byte aligned");
  return mttkrp_tc(V, M, X, nx, Y, ny, R, F, workspace, workspace_bytes, (cudaStream_t)stream_);
}
