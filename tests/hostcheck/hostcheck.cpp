// hostcheck.cpp - compiles the product's numerics header (csrc/numerics.cuh) for the HOST so that
// the CPU test-suite can pin the exact float32 recipe the kernels use (candidate grid, scale,
// fast-path rounding + fallback, fixed-point MSE) against the oracle without a GPU.
// TEST INFRASTRUCTURE: never linked into libadmmq.so, never used by the product.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/test_host_numerics.py does it).
#include <cstdint>
#include <cmath>
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
