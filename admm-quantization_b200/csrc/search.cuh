// search.cuh - CTA-level building blocks of the clip-search projection
// (quantize_tensor_mse, source/quantization.py:118-144), shared by the standalone
// projection kernels (project.cu) and the persistent ADMM loop (admm_loop.cu).
//
// Work split: the CTA's elements are staged through shared memory; every warp walks its
// slice of the stage reading one element at a time as a warp-wide broadcast, and each LANE
// owns CPL candidates (scale and 1/scale in registers) -> no per-candidate warp reduction.
// Squared errors are added in float32 over aligned groups of 8 consecutive elements and the
// group sums go into float64; the CTA total is converted to fixed point and added to the
// global per-candidate accumulators with integer atomics (order independent).
#pragma once
#include "common.cuh"
#include "numerics.cuh"

namespace admmq {

constexpr int kThreads = 256;               // CTA size of every kernel that uses these helpers
constexpr int kWarps = kThreads / 32;
constexpr int kCPL = 7;                     // candidates per lane per pass (200 = 6.25 * 32)
constexpr int kCandPerPass = 32 * kCPL;     // 224
constexpr int kStage = 2048;                // elements staged in shared memory at a time
constexpr int kGroup = 8;                   // float32 accumulation group (aligned, see header)
constexpr int kMaxCandidates = 1024;        // num_attempts limit (custom_benchmark.py uses 1000)
constexpr int kChunkAlign = 64;             // CTA chunks are multiples of this many elements

struct SearchSmem {
  float stage[kStage];
  double red[kWarps * kCandPerPass];
  unsigned long long key[kWarps];
};

// Elements [e0, e1) of the flattened tensor belong to this CTA.
__host__ __device__ inline long long chunk_size(long long n, int ctas) {
  long long c = (n + ctas - 1) / ctas;
  return (c + kChunkAlign - 1) / kChunkAlign * kChunkAlign;
}

// Adds this CTA's share of sum_e (x_e - Q_c(x_e))^2 for every candidate c < Nc into
// cand_sums[c] (fixed point, see numerics.cuh).  loadv(e) returns element e.
template <class LoadV>
__device__ void cta_candidate_sums(LoadV loadv, long long e0, long long e1, float absmax, int Nc,
                                   const Levels L, double n_total, unsigned long long* cand_sums,
                                   SearchSmem& sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (e0 >= e1) return;  // uniform per CTA
  const ClipGrid g = make_clip_grid(absmax, Nc);
  const double unit_inv = fixed_point_unit_inv(n_total, absmax);
  for (int c0 = 0; c0 < Nc; c0 += kCandPerPass) {
    float sc[kCPL], rc[kCPL];
    double dacc[kCPL];
#pragma unroll
    for (int j = 0; j < kCPL; ++j) {
      const int c = c0 + j * 32 + lane;
      const float s = (c < Nc) ? scale_of(clip_candidate(g, c), L) : 1.0f;
      sc[j] = s;
      rc[j] = (c < Nc) ? div_rn(1.0f, s) : 0.0f;
      dacc[j] = 0.0;
    }
    for (long long base = e0; base < e1; base += kStage) {
      const int cnt = (int)min((long long)kStage, e1 - base);
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += kThreads) sm.stage[i] = loadv(base + i);
      __syncthreads();
      // warp slice, multiple of kGroup (base and e0 are multiples of kChunkAlign)
      int per = (cnt + kWarps - 1) / kWarps;
      per = (per + kGroup - 1) / kGroup * kGroup;
      const int wb = min(warp * per, cnt), we = min(wb + per, cnt);
      for (int gb = wb; gb < we; gb += kGroup) {
        float acc[kCPL];
#pragma unroll
        for (int j = 0; j < kCPL; ++j) acc[j] = 0.0f;
        const int ge = min(gb + kGroup, we);
        for (int i = gb; i < ge; ++i) {
          const float x = sm.stage[i];
          float c[kCPL];
          float worst = 0.0f;
#pragma unroll
          for (int j = 0; j < kCPL; ++j) {
            float frac;
            c[j] = sqerr_fast(x, sc[j], rc[j], L, frac);
            worst = fmaxf(worst, frac);
          }
          if (!(worst <= L.fast_thr)) {  // rare: quotient too close to a rounding boundary (or non-finite)
#pragma unroll
            for (int j = 0; j < kCPL; ++j) c[j] = sqerr_exact(x, sc[j], L);
          }
#pragma unroll
          for (int j = 0; j < kCPL; ++j) acc[j] = add_rn(acc[j], c[j]);
        }
#pragma unroll
        for (int j = 0; j < kCPL; ++j) dacc[j] += (double)acc[j];
      }
    }
    // fixed-order reduction over the CTA's warps, then one integer atomic per candidate
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kCPL; ++j) sm.red[warp * kCandPerPass + j * 32 + lane] = dacc[j];
    __syncthreads();
    if (threadIdx.x < kCandPerPass && c0 + (int)threadIdx.x < Nc) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) tot += sm.red[w * kCandPerPass + threadIdx.x];
      const long long fx = __double2ll_rn(tot * unit_inv);
      atomicAdd(cand_sums + c0 + threadIdx.x, (unsigned long long)fx);
    }
  }
}

// First index of the smallest MSE (torch.argmin, source/quantization.py:141), evaluated
// redundantly by every CTA from the global fixed-point sums.  Returns the index to all threads.
__device__ inline int cta_best_candidate(const unsigned long long* cand_sums, int Nc, float absmax,
                                         double n_total, SearchSmem& sm) {
  const double unit = fixed_point_unit(n_total, absmax);
  const float nf = (float)n_total;
  unsigned long long best = ~0ull;
  for (int c = threadIdx.x; c < Nc; c += kThreads) {
    const long long fx = (long long)__ldcg(cand_sums + c);
    const float mse = mse_from_fixed(fx, unit, nf);
    const unsigned long long key = ((unsigned long long)float_key(mse) << 32) | (unsigned int)c;
    best = min(best, key);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm.key[threadIdx.x >> 5] = best;
  __syncthreads();
  unsigned long long b = sm.key[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) b = min(b, sm.key[w]);
  return (int)(b & 0xffffffffu);
}

// Parameters of the non-search schemes from the tensor's min / max (source/quantization.py:48-66, 91-106)
__device__ inline QParams params_from_minmax(int scheme, int bits, float tmin, float tmax, const Levels& L) {
  QParams p;
  p.scheme = scheme;
  p.bits = bits;
  p.n = (float)((1u << bits) - 1u);
  p.aux = 0.0f;
  p.scale = 0.0f;
  if (scheme == ADMMQ_Q_MINMAX) {
    p.scale = sub_rn(tmax, tmin);
    p.aux = tmin;
  } else if (scheme == ADMMQ_Q_SYMMETRIC) {
    const float a = fabsf(tmin);
    p.scale = div_rn(mul_rn(2.0f, (a > tmax) ? a : tmax), L.denom);
  } else if (scheme == ADMMQ_Q_AFFINE) {
    p.scale = div_rn(sub_rn(tmax, tmin), L.denom);
    p.aux = affine_zero_point(tmin, p.scale, L);
  }
  return p;
}

// value on the grid and its integer code for any scheme
__device__ __forceinline__ float quantize_value(float x, const QParams& p, const Levels& L, float& code) {
  if (p.scheme == ADMMQ_Q_MINMAX) {
    float level;
    const float v = minmax_value(x, p, level);
    code = (p.bits == 1) ? level : level - (float)(1 << (p.bits - 1));
    return v;
  }
  if (p.scheme == ADMMQ_Q_AFFINE) return affine_value(x, p, L, code);
  code = code_exact(x, p.scale, L);
  // tensor_symmetric multiplies the INTEGER code by the scale (source/quantization.py:95 `.to(int)`),
  // which has no negative zero; the clip search (:127, :144) stays in float and keeps -0.
  if (p.scheme == ADMMQ_Q_SYMMETRIC) code = add_rn(code, 0.0f);
  return mul_rn(code, p.scale);
}

}  // namespace admmq
