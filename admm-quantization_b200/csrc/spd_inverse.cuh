// spd_inverse.cuh - rho = trace(G)/R and Minv = (G + rho I)^-1 in float64 by one cooperative
// grid (replaces torch.linalg.cholesky(G + rho*eye) of source/admm.py:52-54; the per-iteration
// torch.cholesky_solve of :56 becomes a product with Minv, see admm_loop.cu).
//
// R exceeds what fits in shared memory for most layers (R = 134 ... 1522), so the matrix lives
// in global memory (L2 resident) as 32x32 float64 tiles and the grid walks a blocked algorithm
// separated by device-wide barriers:
//   1. left-looking blocked Cholesky  A = L L^T           (2 barriers per 32-wide panel)
//   2. blocked triangular inverse     X = L^-1            (1 barrier per block diagonal)
//   3. Minv = X^T X, rounded to float32                   (1 barrier)
// cond(G + rho I) <= R + 1 by construction of rho, and everything is float64, so Minv is the
// correctly rounded inverse for all practical purposes.
#pragma once
#include "common.cuh"
#include "numerics.cuh"

namespace admmq {

constexpr int kNB = 32;          // tile edge
constexpr int kInvThreads = 256; // CTA size

struct InvSmem {
  double a[kNB][kNB + 1];
  double b[kNB][kNB + 1];
  double c[kNB][kNB + 1];
  double a2[kNB][kNB + 1];  // second buffers of the software-pipelined inner loops: step q computes from one pair
  double b2[kNB][kNB + 1];  // while step q+1 is stashed into the other, so one barrier per step is enough
  double red[kInvThreads];
  int flag;
};

struct InvParams {
  const float* G;       // R x R
  int R, nb, Rb;        // Rb = nb * 32
  int ldm;              // leading dimension of Minv
  double* Lw;           // Rb x Rb  (in: A, out: L in the lower tiles)
  double* Xw;           // Rb x Rb  (L^-1 in the lower tiles)
  float* Minv;          // R x ldm
  double* Minv64;       // R x ldm or nullptr: the same inverse before its rounding to float32 (parity mode of the loop)
  float* rho_out;       // device float
  int* status;          // device int: 0 or ADMMQ_E_NOT_PD
  unsigned int* barrier;
};

__device__ __forceinline__ void inv_load_tile(double (*dst)[kNB + 1], const double* src, int ld) {
  for (int i = threadIdx.x; i < kNB * kNB; i += kInvThreads) dst[i >> 5][i & 31] = __ldcg(src + (size_t)(i >> 5) * ld + (i & 31));
}

// The same tile load split in two halves, so that the L2 round trip of step q+1 overlaps the FMA loop of step q:
// fetch (global -> registers) before the compute, stash (registers -> shared memory) after the next barrier.
__device__ __forceinline__ void inv_fetch_tile(double r[4], const double* src, int ld) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int i = threadIdx.x + m * kInvThreads;
    r[m] = __ldcg(src + (size_t)(i >> 5) * ld + (i & 31));
  }
}
__device__ __forceinline__ void inv_stash_tile(double (*dst)[kNB + 1], const double r[4]) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int i = threadIdx.x + m * kInvThreads;
    dst[i >> 5][i & 31] = r[m];
  }
}

// acc[m] (+/-)= sum_q A[row_m][q] * B[c][q]      (B used transposed)
__device__ __forceinline__ void inv_mma_nt(double acc[4], const double (*A)[kNB + 1], const double (*B)[kNB + 1], double sign) {
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
#pragma unroll 8
  for (int q = 0; q < kNB; ++q) {
    const double bv = sign * B[c][q];
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[m] = fma(A[rg + 8 * m][q], bv, acc[m]);
  }
}
// acc[m] += sign * sum_q A[row_m][q] * B[q][c]
__device__ __forceinline__ void inv_mma_nn(double acc[4], const double (*A)[kNB + 1], const double (*B)[kNB + 1], double sign) {
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
#pragma unroll 8
  for (int q = 0; q < kNB; ++q) {
    const double bv = sign * B[q][c];
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[m] = fma(A[rg + 8 * m][q], bv, acc[m]);
  }
}
// acc[m] += sum_q A[q][row_m] * B[q][c]           (A used transposed)
__device__ __forceinline__ void inv_mma_tn(double acc[4], const double (*A)[kNB + 1], const double (*B)[kNB + 1]) {
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
#pragma unroll 8
  for (int q = 0; q < kNB; ++q) {
    const double bv = B[q][c];
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[m] = fma(A[q][rg + 8 * m], bv, acc[m]);
  }
}

// In-place Cholesky of the 32x32 tile in sm.a (lower), then its inverse into sm.b (lower, upper zeroed).
__device__ inline void inv_factor_diag(InvSmem& sm) {
  const int t = threadIdx.x;
  for (int j = 0; j < kNB; ++j) {
    if (t == 0) {
      const double d = sm.a[j][j];
      if (!(d > 0.0) || !(d < 1.0e300)) sm.flag = 1;
      sm.a[j][j] = sqrt(d);
    }
    __syncthreads();
    if (t > j && t < kNB) sm.a[t][j] /= sm.a[j][j];
    __syncthreads();
    for (int i = t; i < kNB * kNB; i += kInvThreads) {
      const int r = i >> 5, c = i & 31;
      if (c > j && r >= c) sm.a[r][c] -= sm.a[r][j] * sm.a[c][j];
    }
    __syncthreads();
  }
  if (t < kNB) {
    const int c = t;
    for (int r = 0; r < c; ++r) sm.b[r][c] = 0.0;
    sm.b[c][c] = 1.0 / sm.a[c][c];
    for (int r = c + 1; r < kNB; ++r) {
      double s = 0.0;
      for (int q = c; q < r; ++q) s = fma(sm.a[r][q], sm.b[q][c], s);
      sm.b[r][c] = -s / sm.a[r][r];
    }
  }
  __syncthreads();
}

// Whole procedure; every CTA of the cooperative grid calls it.  Returns 0 or ADMMQ_E_NOT_PD (uniform).
__device__ inline int spd_inverse_body(const InvParams& p, InvSmem& sm, GridBarrier& bar) {
  const int t = threadIdx.x, c = t & 31, rg = t >> 5;
  const int R = p.R, nb = p.nb, Rb = p.Rb;
  // ---- rho = trace(G) / R: ATen's CPU trace accumulates in double, result float32 (source/admm.py:53)
  double part = 0.0;
  for (int i = t; i < R; i += kInvThreads) part += (double)p.G[(size_t)i * R + i];
  sm.red[t] = part;
  if (t == 0) sm.flag = 0;
  __syncthreads();
  for (int s = kInvThreads / 2; s > 0; s >>= 1) {
    if (t < s) sm.red[t] += sm.red[t + s];
    __syncthreads();
  }
  const float rho = div_rn((float)sm.red[0], (float)R);
  if (blockIdx.x == 0 && t == 0) *p.rho_out = rho;
  // ---- A = G + rho*I as float32 (what the reference factors, :54), widened to float64; pad = identity
  for (size_t idx = (size_t)blockIdx.x * kInvThreads + t; idx < (size_t)Rb * Rb; idx += (size_t)gridDim.x * kInvThreads) {
    const int i = (int)(idx / Rb), j = (int)(idx % Rb);
    double v = 0.0;
    if (i < R && j < R) v = (i == j) ? (double)add_rn(p.G[(size_t)i * R + j], rho) : (double)p.G[(size_t)i * R + j];
    else if (i == j) v = 1.0;
    p.Lw[idx] = v;
  }
  bar.sync();
  // ---- 1. blocked Cholesky
  for (int k = 0; k < nb; ++k) {
    if (k > 0) {
      for (int ti = blockIdx.x; ti < nb - k; ti += gridDim.x) {  // S_ik = A_ik - sum_{p<k} L_ip L_kp^T
        const int i = k + ti;
        double* tile = p.Lw + ((size_t)i * kNB) * Rb + (size_t)k * kNB;
        double acc[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) acc[m] = __ldcg(tile + (size_t)(rg + 8 * m) * Rb + c);
        double ra[4], rb[4];
        inv_fetch_tile(ra, p.Lw + ((size_t)i * kNB) * Rb, Rb);
        inv_fetch_tile(rb, p.Lw + ((size_t)k * kNB) * Rb, Rb);
        __syncthreads();
        for (int q = 0; q < k; ++q) {
          double (*ta)[kNB + 1] = (q & 1) ? sm.a2 : sm.a;
          double (*tb)[kNB + 1] = (q & 1) ? sm.b2 : sm.b;
          inv_stash_tile(ta, ra);
          inv_stash_tile(tb, rb);
          __syncthreads();
          if (q + 1 < k) {
            inv_fetch_tile(ra, p.Lw + ((size_t)i * kNB) * Rb + (size_t)(q + 1) * kNB, Rb);
            inv_fetch_tile(rb, p.Lw + ((size_t)k * kNB) * Rb + (size_t)(q + 1) * kNB, Rb);
          }
          inv_mma_nt(acc, ta, tb, -1.0);
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 4; ++m) tile[(size_t)(rg + 8 * m) * Rb + c] = acc[m];
      }
      bar.sync();
    }
    if ((int)blockIdx.x < nb - k) {  // this CTA owns at least one tile of panel k: factor S_kk redundantly
      __syncthreads();
      inv_load_tile(sm.a, p.Lw + ((size_t)k * kNB) * Rb + (size_t)k * kNB, Rb);
      __syncthreads();
      inv_factor_diag(sm);  // sm.a = L_kk, sm.b = L_kk^-1
      for (int ti = blockIdx.x; ti < nb - k; ti += gridDim.x) {
        const int i = k + ti;
        double* tile = p.Lw + ((size_t)i * kNB) * Rb + (size_t)k * kNB;
        if (ti == 0) {
          // Only L_kk^-1 is kept (diagonal tile of Xw).  L_kk itself is never read again, and the
          // tile in Lw must keep S_kk because other CTAs may still be loading it.
          double* xt = p.Xw + ((size_t)k * kNB) * Rb + (size_t)k * kNB;
          for (int e = t; e < kNB * kNB; e += kInvThreads) xt[(size_t)(e >> 5) * Rb + (e & 31)] = sm.b[e >> 5][e & 31];
        } else {  // L_ik = S_ik L_kk^-T
          __syncthreads();
          inv_load_tile(sm.c, tile, Rb);
          __syncthreads();
          double acc[4] = {0.0, 0.0, 0.0, 0.0};
          inv_mma_nt(acc, sm.c, sm.b, 1.0);
#pragma unroll
          for (int m = 0; m < 4; ++m) tile[(size_t)(rg + 8 * m) * Rb + c] = acc[m];
        }
      }
      if (sm.flag != 0 && t == 0) atomicExch(p.status, ADMMQ_E_NOT_PD);
    }
    bar.sync();
    if (__ldcg(p.status) != 0) return ADMMQ_E_NOT_PD;  // uniform: every CTA reads the same word after the barrier
  }
  // ---- 2. X = L^-1 by block diagonals: X_ik = -X_ii * sum_{j=k}^{i-1} L_ij X_jk
  for (int d = 1; d < nb; ++d) {
    for (int k = blockIdx.x; k < nb - d; k += gridDim.x) {
      const int i = k + d;
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      double ra[4], rb[4];
      inv_fetch_tile(ra, p.Lw + ((size_t)i * kNB) * Rb + (size_t)k * kNB, Rb);
      inv_fetch_tile(rb, p.Xw + ((size_t)k * kNB) * Rb + (size_t)k * kNB, Rb);
      __syncthreads();
      for (int j = k; j < i; ++j) {
        double (*ta)[kNB + 1] = ((j - k) & 1) ? sm.a2 : sm.a;
        double (*tb)[kNB + 1] = ((j - k) & 1) ? sm.b2 : sm.b;
        inv_stash_tile(ta, ra);
        inv_stash_tile(tb, rb);
        __syncthreads();
        if (j + 1 < i) {
          inv_fetch_tile(ra, p.Lw + ((size_t)i * kNB) * Rb + (size_t)(j + 1) * kNB, Rb);
          inv_fetch_tile(rb, p.Xw + ((size_t)(j + 1) * kNB) * Rb + (size_t)k * kNB, Rb);
        }
        inv_mma_nn(acc, ta, tb, 1.0);
      }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < 4; ++m) sm.c[rg + 8 * m][c] = acc[m];
      inv_load_tile(sm.a, p.Xw + ((size_t)i * kNB) * Rb + (size_t)i * kNB, Rb);
      __syncthreads();
      double out[4] = {0.0, 0.0, 0.0, 0.0};
      inv_mma_nn(out, sm.a, sm.c, -1.0);
      double* xt = p.Xw + ((size_t)i * kNB) * Rb + (size_t)k * kNB;
#pragma unroll
      for (int m = 0; m < 4; ++m) xt[(size_t)(rg + 8 * m) * Rb + c] = out[m];
    }
    bar.sync();
  }
  // ---- 3. Minv = X^T X (lower block triangle computed, mirrored on store), float32
  const int npairs = nb * (nb + 1) / 2;
  for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
    int a = 0, rem = pr;
    while (rem >= a + 1) {  // pr -> (a, b) with b <= a
      rem -= a + 1;
      ++a;
    }
    const int b = rem;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    double ra[4], rb[4];
    inv_fetch_tile(ra, p.Xw + ((size_t)a * kNB) * Rb + (size_t)a * kNB, Rb);
    inv_fetch_tile(rb, p.Xw + ((size_t)a * kNB) * Rb + (size_t)b * kNB, Rb);
    __syncthreads();
    for (int q = a; q < nb; ++q) {
      double (*ta)[kNB + 1] = ((q - a) & 1) ? sm.a2 : sm.a;
      double (*tb)[kNB + 1] = ((q - a) & 1) ? sm.b2 : sm.b;
      inv_stash_tile(ta, ra);
      inv_stash_tile(tb, rb);
      __syncthreads();
      if (q + 1 < nb) {
        inv_fetch_tile(ra, p.Xw + ((size_t)(q + 1) * kNB) * Rb + (size_t)a * kNB, Rb);
        inv_fetch_tile(rb, p.Xw + ((size_t)(q + 1) * kNB) * Rb + (size_t)b * kNB, Rb);
      }
      inv_mma_tn(acc, ta, tb);
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int row = a * kNB + rg + 8 * m, col = b * kNB + c;
      if (row < R && col < p.ldm) p.Minv[(size_t)row * p.ldm + col] = (col < R) ? (float)acc[m] : 0.0f;
      if (a != b && col < R && row < p.ldm) p.Minv[(size_t)col * p.ldm + row] = (row < R) ? (float)acc[m] : 0.0f;
      if (p.Minv64 != nullptr) {
        if (row < R && col < p.ldm) p.Minv64[(size_t)row * p.ldm + col] = (col < R) ? acc[m] : 0.0;
        if (a != b && col < R && row < p.ldm) p.Minv64[(size_t)col * p.ldm + row] = (row < R) ? acc[m] : 0.0;
      }
    }
  }
  bar.sync();
  return 0;
}

}  // namespace admmq
