// hostcheck.cpp - compiles the product's numerics header (csrc/numerics.cuh) for the HOST so that
// the CPU test-suite can pin the exact float32 recipe the kernels use (candidate grid, scale,
// fast-path rounding + fallback, fixed-point MSE) against the oracle without a GPU.
// TEST INFRASTRUCTURE: never linked into libadmmq.so, never used by the product.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/test_host_numerics.py does it).
#include <cstdint>
#include <cmath>
#include <algorithm>
#include <vector>
#include "../../admm-quantization_b200/csrc/numerics.cuh"

using namespace admmq;

extern "C" {

void hc_candidates(float absmax, int n, int bits, float* clip, float* scale) {
  const Levels L = make_levels(bits);
  const ClipGrid g = make_clip_grid(absmax, n);
  for (int i = 0; i < n; ++i) {
    clip[i] = clip_candidate(g, i);
    scale[i] = scale_of(clip[i], L);
  }
}

// per-candidate MSE exactly as the kernels form it: fast path with exact fallback per group of 8,
// the group_sum8 recipe of numerics.cuh, float32 over the 8 groups of a 64-element block, float64 across blocks,
// fixed point, float32 mean.
// Returns the number of groups that needed the exact-division fallback.
long long hc_mse(const float* x, long long n, float absmax, int bits, int nc, float* mse, int* best, int force_exact) {
  const Levels L = make_levels(bits);
  const ClipGrid g = make_clip_grid(absmax, nc);
  const double unit_inv = fixed_point_unit_inv((double)n, absmax), unit = fixed_point_unit((double)n, absmax);
  long long slow = 0;
  int arg = 0;
  for (int c = 0; c < nc; ++c) {
    const float s = scale_of(clip_candidate(g, c), L);
    const float r = div_rn(1.0f, s);
    double tot = 0.0;
    float block = 0.0f;
    for (long long b = 0; b < n; b += kSumGroup) {
      float d[kSumGroup];
      float worst = 0.0f;
      for (int i = 0; i < kSumGroup; ++i) {
        const float xv = (b + i < n) ? x[b + i] : 0.0f;
        float frac;
        d[i] = dev_fast(xv, s, r, L, frac);
        worst = fmaxf(worst, frac);
      }
      if (force_exact || !(worst <= L.fast_thr)) {
        for (int i = 0; i < kSumGroup; ++i) d[i] = dev_exact((b + i < n) ? x[b + i] : 0.0f, s, L);
        ++slow;
      }
      block = add_rn(block, group_sum8(d));
      if ((b + kSumGroup) % kSumBlock == 0 || b + kSumGroup >= n) {
        tot += (double)block;
        block = 0.0f;
      }
    }
    const long long fx = llrint(tot * unit_inv);
    mse[c] = mse_from_fixed(fx, unit, (float)n);
    if (mse[c] < mse[arg]) arg = c;
  }
  *best = arg;
  return slow;
}

// per-candidate MSE in the THRESHOLD form the kernels use (numerics.cuh): exact thresholds, counts and fixed-point
// prefix sums below each threshold (here from a sorted copy), float64 terms converted one by one to the accumulator's
// fixed point, float32 mean.  Also returns, through `sums`, the float64 value of each sum.
void hc_mse_thresholds(const float* x, long long n, float absmax, int bits, int nc, float* mse, double* sums, int* best) {
  const Levels L = make_levels(bits);
  const ClipGrid g = make_clip_grid(absmax, nc);
  const double unit_inv = fixed_point_unit_inv((double)n, absmax), unit = fixed_point_unit((double)n, absmax);
  const FixX fx = make_fix_x(absmax);
  std::vector<float> xs(x, x + n);
  std::sort(xs.begin(), xs.end());
  std::vector<long long> pre(n + 1, 0);
  for (long long i = 0; i < n; ++i) pre[i + 1] = pre[i] + fix_x(xs[i], fx);
  double x2 = 0.0;
  for (long long i = 0; i < n; ++i) x2 += (double)x[i] * (double)x[i];
  const long long x2f = llrint(x2 * unit_inv);
  const int nthr = (1 << bits) - 1;
  int arg = 0;
  for (int c = 0; c < nc; ++c) {
    const float s = scale_of(clip_candidate(g, c), L);
    long long acc = 0;
    for (int j = 0; j < nthr; ++j) {
      const float level = L.lo + (float)j;
      const float theta = code_threshold(s, level);
      const long long cnt = std::lower_bound(xs.begin(), xs.end(), theta) - xs.begin();  // #{x < theta}
      double term = threshold_term(s, level, cnt, pre[cnt], fx.unit);
      if (j == nthr - 1) term += closing_term(s, L.hi, n, pre[n], fx.unit);
      acc += llrint(term * unit_inv);
    }
    const long long f = std::max(acc + x2f, 0ll);
    sums[c] = (double)f * unit;
    mse[c] = mse_from_fixed(f, unit, (float)n);
    if (mse[c] < mse[arg]) arg = c;
  }
  *best = arg;
}

// code_threshold(): theta is the smallest float whose code reaches level + 1.  Returns the number of violations among
// the thresholds of every level for this scale (checks theta itself and its predecessor with the exact division).
long long hc_threshold_violations(float scale, int bits) {
  const Levels L = make_levels(bits);
  long long bad = 0;
  for (float level = L.lo; level < L.hi; level += 1.0f) {
    const float theta = code_threshold(scale, level);
    const float below = ordered_float(ordered_key(theta) - 1u);
    if (!(rint_rn(div_rn(theta, scale)) >= level + 1.0f)) ++bad;
    if (rint_rn(div_rn(below, scale)) >= level + 1.0f) ++bad;
    if (theta != code_threshold_search(scale, level)) ++bad;   // closed form == search over neighbouring floats
  }
  return bad;
}

// fast path alone vs exact path: count of elements where they disagree although the fast path was accepted
long long hc_fastpath_violations(const float* x, long long n, float scale, int bits) {
  const Levels L = make_levels(bits);
  const float r = div_rn(1.0f, scale);
  long long bad = 0;
  for (long long i = 0; i < n; ++i) {
    float frac;
    const float f = sqerr_fast(x[i], scale, r, L, frac);
    if (frac <= L.fast_thr && f != sqerr_exact(x[i], scale, L)) ++bad;
    float frac2;
    const float d2 = dev_fast2(x[i], scale, r, L, frac2);
    if (frac2 <= L.fast_thr && d2 != dev_exact(x[i], scale, L)) ++bad;
  }
  return bad;
}

void hc_quantize(const float* x, long long n, float scale, int bits, float* xq, signed char* codes) {
  const Levels L = make_levels(bits);
  for (long long i = 0; i < n; ++i) {
    const float k = code_exact(x[i], scale, L);
    codes[i] = (signed char)k;
    xq[i] = mul_rn(k, scale);
  }
}

void hc_minmax(const float* x, long long n, float tmin, float tmax, int bits, float* out) {
  QParams p;
  p.scheme = 1;
  p.bits = bits;
  p.scale = sub_rn(tmax, tmin);
  p.aux = tmin;
  p.n = (float)((1u << bits) - 1u);
  for (long long i = 0; i < n; ++i) {
    float lvl;
    out[i] = minmax_value(x[i], p, lvl);
  }
}

void hc_affine(const float* x, long long n, float tmin, float tmax, int bits, float* out) {
  const Levels L = make_levels(bits);
  QParams p;
  p.scheme = 3;
  p.bits = bits;
  p.scale = div_rn(sub_rn(tmax, tmin), L.denom);
  p.aux = affine_zero_point(tmin, p.scale, L);
  p.n = 0.0f;
  for (long long i = 0; i < n; ++i) {
    float code;
    out[i] = affine_value(x[i], p, L, code);
  }
}
}
