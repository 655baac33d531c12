// numerics.cuh - the float32 recipe of the projection, shared by every kernel that
// quantizes (and compiled for the host by tests/hostcheck to pin it against the oracle).
//
// Follows source/quantization.py:69-144 of the reference operation by operation:
// IEEE float32 multiply / divide / subtract with round-to-nearest-even and no FMA
// contraction, `torch.round` = rint (half to even), `torch.linspace` = FMA form
// (SURVEY App. A.3).  Nothing here may be compiled with --use_fast_math.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define ADMMQ_HD __host__ __device__ __forceinline__
#else
#define ADMMQ_HD inline
#endif

namespace admmq {

// --- explicitly rounded float32 primitives (no contraction on either side) ------------
ADMMQ_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
ADMMQ_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
ADMMQ_HD float sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
ADMMQ_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
ADMMQ_HD float fma_rn(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
ADMMQ_HD float rint_rn(float a) {
#if defined(__CUDA_ARCH__)
  return rintf(a);
#else
  return nearbyintf(a);
#endif
}

// --- quantizer description -------------------------------------------------------------
struct Levels {
  float lo;     // -q            (source/quantization.py:88: q = 2**(bits-1))
  float hi;     //  q-1
  float denom;  //  2q-1         (:89 scale_denom)
  float fast_lo, fast_hi, fast_thr;  // fast-path clamp window and "too close to .5" threshold
};

ADMMQ_HD Levels make_levels(int bits) {
  Levels L;
  const float q = (float)(1 << (bits - 1));
  L.lo = -q;
  L.hi = q - 1.0f;
  L.denom = 2.0f * q - 1.0f;
  L.fast_lo = -q - 0.25f;
  L.fast_hi = q - 0.75f;
  // |x*rcp(s) - x/s| <= |t| * 2^-23 (two roundings) and |fl(x/s) - x/s| <= |t| * 2^-24;
  // (q + 2) * 2^-21 leaves a >2.5x margin for every quotient |t| <= q + 1.5 whose rounding can matter (larger ones
  // clamp to the same code either way).
  L.fast_thr = 0.5f - (q + 2.0f) * 4.76837158203125e-07f;
  return L;
}

// --- candidate grid of quantize_tensor_mse (source/quantization.py:129-131) ----------
struct ClipGrid {
  float start, end, step;
  int n;
};

ADMMQ_HD ClipGrid make_clip_grid(float absmax, int n) {
  ClipGrid g;
  g.n = n;
  g.start = (float)(0.2 * (double)absmax);
  g.end = (float)(1.2 * (double)absmax);
  g.step = n > 1 ? div_rn(sub_rn(g.end, g.start), (float)(n - 1)) : 0.0f;
  return g;
}

ADMMQ_HD float clip_candidate(const ClipGrid& g, int i) {
  if (g.n == 1) return g.start;
  return (i < g.n / 2) ? fma_rn(g.step, (float)i, g.start) : fma_rn(-g.step, (float)(g.n - 1 - i), g.end);
}

// scale = 2 * tmax / scale_denom  (source/quantization.py:125)
ADMMQ_HD float scale_of(float clip, const Levels& L) { return div_rn(mul_rn(2.0f, clip), L.denom); }

// clamp(round(x / scale), -q, q-1)  (source/quantization.py:127) - the exact form
ADMMQ_HD float code_exact(float x, float scale, const Levels& L) {
  float k = rint_rn(div_rn(x, scale));
  k = fminf(fmaxf(k, L.lo), L.hi);
  return k;
}

// (x - code*scale)^2 with every intermediate rounded to float32 (:127, :138)
ADMMQ_HD float sqerr_exact(float x, float scale, const Levels& L) {
  const float d = sub_rn(x, mul_rn(code_exact(x, scale, L), scale));
  return mul_rn(d, d);
}

// Deviation (x - code*scale) with every intermediate rounded to float32 (:127, :138) - exact form.
ADMMQ_HD float dev_exact(float x, float scale, const Levels& L) {
  return sub_rn(x, mul_rn(code_exact(x, scale, L), scale));
}

// Fast path used inside the 200-candidate search: the code is obtained from x * (1/scale)
// (no division); `frac` returns |t - k| so that the caller can detect the rare inputs whose
// quotient lies too close to a rounding boundary for the shortcut to be provably equal to
// code_exact() and redo them with dev_exact().  MAGIC rounding = round-half-even for |t| < 2^22
// (|t| <= 37.5 here: |x| <= absmax and scale >= 0.2 * absmax * 2 / denom).
ADMMQ_HD float dev_fast(float x, float scale, float rcp_scale, const Levels& L, float& frac) {
  const float MAGIC = 12582912.0f;  // 1.5 * 2^23
  float t = mul_rn(x, rcp_scale);
  t = fminf(fmaxf(t, L.fast_lo), L.fast_hi);
  const float k = sub_rn(add_rn(t, MAGIC), MAGIC);
  frac = fabsf(sub_rn(t, k));
  return sub_rn(x, mul_rn(k, scale));
}
// An alternative form (not used by the kernels, see search.cuh eval_pair; kept pinned by the host tests): round the EXACT product x * rcp_scale (one FMA with the magic
// constant), take its rounding residual r with a second FMA, clamp the integer afterwards.  rint(x / scale) can differ
// from the rounded product only when the quotient lies within |x/s| * 2^-23 of a half integer, and only matters while
// the integer is inside [-q - 1, q]; |r| <= fast_thr = 0.5 - q * 2^-21 excludes both, everything else is redone exactly.
ADMMQ_HD float dev_fast2(float x, float scale, float rcp_scale, const Levels& L, float& frac) {
  const float MAGIC = 12582912.0f;  // 1.5 * 2^23
  const float ku = sub_rn(fma_rn(x, rcp_scale, MAGIC), MAGIC);
  frac = fabsf(fma_rn(x, rcp_scale, -ku));
  const float k = fminf(fmaxf(ku, L.lo), L.hi);
  return sub_rn(x, mul_rn(k, scale));
}
ADMMQ_HD float sqerr_fast(float x, float scale, float rcp_scale, const Levels& L, float& frac) {
  const float d = dev_fast(x, scale, rcp_scale, L, frac);
  return mul_rn(d, d);
}

// --- the summation recipe of the clip search ---------------------------------------------
// The reference takes `mean((x - xq)**2)` in float32 with ATen's reduction order, which no other
// implementation can reproduce bit for bit.  The kernels define the per-candidate sum as follows
// (tests/hostcheck pins this definition against the reference's argmin on the golden vectors):
//   * elements are taken in aligned groups of 8 (kSumGroup); missing tail elements count as x = 0,
//     which contributes exactly 0;
//   * within a group two float32 FMA chains accumulate d*d, one over the even and one over the odd
//     positions (the two lanes of a packed f32x2 FMA), and are added in float32;
//   * the 8 group sums of an aligned block of 64 elements (kSumBlock) are added in float32 in order (the float32 ->
//     float64 conversion runs on the quarter-rate conversion pipe: once per 64 elements instead of once per 8 it
//     costs ~2 % instead of ~14 % of the search);
//   * block sums are added in float64 in element order, the CTA/chunk totals in 64-bit fixed point.
// Blocks are aligned to the ABSOLUTE element index (chunk starts, stage sizes and warp slices are multiples of 64), so
// the result does not depend on the grid size or on how the elements are split over CTAs and warps.
constexpr int kSumGroup = 8;
constexpr int kSumBlock = 64;
ADMMQ_HD float group_sum8(const float d[kSumGroup]) {
  float even = 0.0f, odd = 0.0f;
  for (int i = 0; i < kSumGroup; i += 2) {
    even = fma_rn(d[i], d[i], even);
    odd = fma_rn(d[i + 1], d[i + 1], odd);
  }
  return add_rn(even, odd);
}

// --- fixed-point accumulation of the per-candidate squared-error sums --------------------
// Every element contributes d^2 <= 16 * absmax^2, so sum <= 16 * N * absmax^2 which is mapped to
// 2^62.  Integer addition is associative: the total does not depend on how many CTAs or in which
// order the partial sums arrive, and it resolves ~3e-15 of a typical total.
ADMMQ_HD double fixed_point_unit_inv(double n_elems, float absmax) {
  const double bound = n_elems * (double)absmax * (double)absmax;  // * 16 folded into 2^58
  return 288230376151711744.0 /* 2^58 */ / bound;
}
ADMMQ_HD double fixed_point_unit(double n_elems, float absmax) {
  const double bound = n_elems * (double)absmax * (double)absmax;
  return bound / 288230376151711744.0;
}

// mean as the reference forms it (:138 `.mean`): float32 total divided by float32 count, where
// the total is the correctly rounded sum of the float32 squares.
ADMMQ_HD float mse_from_fixed(long long fixed_sum, double unit, float n_elems_f) {
  const float total = (float)((double)fixed_sum * unit);
  return div_rn(total, n_elems_f);
}

// --- the threshold ("binned") form of the clip search ----------------------------------------
// For one candidate scale s the quantizer k(x) = clamp(rint(x / s), -q, q-1) is a monotone step function of x, so
// it is fully described by its 2q-1 thresholds: theta_j = the SMALLEST float32 x with rint(fl(x / s)) >= -q + j + 1
// (found by stepping through neighbouring floats with the exact division, so ties-to-even and the rounding of the
// quotient are captured exactly).  With C_j = #{x < theta_j}, P_j = sum{x < theta_j} x, the grid values
// y_j = fl32((-q + j) * s) (source/quantization.py:127 rounds `codes * scale` to float32) and summation by parts,
//     sum_e (x_e - y_k(e))^2 = sum_e x_e^2 + sum_{j=0}^{2q-2} [ C_j (y_j^2 - y_{j+1}^2) - 2 P_j (y_j - y_{j+1}) ]
//                                           + n y_top^2 - 2 P_tot y_top,
// which needs C and P at (2q-1) * num_attempts thresholds instead of num_attempts evaluations per element.  C is an
// integer, P is accumulated in 64-bit fixed point (x rounded to 2^-41 of the power of two above absmax), the bracket is
// evaluated in float64 and converted to the fixed-point unit of the per-candidate accumulators, so the result is
// independent of arrival order (bit-reproducible) and depends on the chunking only through the fixed-point rounding
// of the per-chunk terms (~1e-13 relative).  It is the exact real-number value of the reference's
// sum((x - xq)**2) up to ~1e-12 relative; the reference's own float32 evaluation of that sum carries ~1e-7.
ADMMQ_HD unsigned int f32_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  unsigned int u;
  std::memcpy(&u, &f, 4);
  return u;
#endif
}
ADMMQ_HD float bits_f32(unsigned int u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  std::memcpy(&f, &u, 4);
  return f;
#endif
}
// order-preserving map float -> uint32 (larger float <=> larger key); NaN maps above +inf
ADMMQ_HD unsigned int ordered_key(float f) {
  const unsigned int b = f32_bits(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
ADMMQ_HD float ordered_float(unsigned int k) { return bits_f32((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// smallest float32 x with rint(fl(x / scale)) >= level + 1   (level = -q .. q-2), in closed form:
//   rint(y) >= T  <=>  y >= y*,  y* = T - 0.5 if T is even (the tie rounds up to T), else the next float above it;
//   fl(z) >= y*   <=>  z > mid or (z == mid and the mantissa of y* is even),  mid = midpoint of pred(y*) and y*;
//   mid * scale is exact in float64 (25 x 24 significant bits), so theta is the first float above that product
//   (or the product itself when it is a float and the tie goes to y*).
// tests/hostcheck pins theta and its predecessor against the exact division for every level of random scales.
// The level-only part (mid, parity of y*) and the scale-dependent part are separate so that the clip search forms the
// former once per level instead of once per (candidate, level) pair.
struct ThresholdMid {
  double mid;   // midpoint of pred(y*) and y* (25 significant bits, never zero)
  bool odd;     // mantissa of y* is odd: a product that lands exactly on mid rounds away from y*
};
ADMMQ_HD ThresholdMid threshold_mid(float level) {
  const float target = level + 1.0f;
  const float b = target - 0.5f;
  const bool t_even = (((int)target) & 1) == 0;
  const float ystar = t_even ? b : ordered_float(ordered_key(b) + 1u);
  const float pred = ordered_float(ordered_key(ystar) - 1u);
  ThresholdMid m;
  m.mid = 0.5 * ((double)pred + (double)ystar);
  m.odd = (f32_bits(ystar) & 1u) != 0u;
  return m;
}
ADMMQ_HD float code_threshold_from_mid(float scale, double mid, bool odd) {
  const double P = mid * (double)scale;                 // exact; scale > 0, so P != 0 and has the sign of mid
  float x = (float)P;                                   // round to nearest
  const double xd = (double)x;
  // first float above P, or P itself when it is a float and the tie goes to y*: one step towards +inf
  if (xd < P || (xd == P && odd)) x = bits_f32(f32_bits(x) + (x > 0.0f ? 1u : 0xffffffffu));
  return x;
}
ADMMQ_HD float code_threshold(float scale, float level) {
  const ThresholdMid m = threshold_mid(level);
  return code_threshold_from_mid(scale, m.mid, m.odd);
}

// the same threshold by stepping through neighbouring floats with the exact division (reference implementation of the
// definition; used by the host tests)
ADMMQ_HD float code_threshold_search(float scale, float level) {
  const float target = level + 1.0f;
  unsigned int key = ordered_key(mul_rn(level + 0.5f, scale));
  if (rint_rn(div_rn(ordered_float(key), scale)) >= target) {
    for (int it = 0; it < 64 && rint_rn(div_rn(ordered_float(key - 1u), scale)) >= target; ++it) --key;
  } else {
    for (int it = 0; it < 64; ++it) {
      ++key;
      if (rint_rn(div_rn(ordered_float(key), scale)) >= target) break;
    }
  }
  return ordered_float(key);
}

// 64-bit fixed point of the elements: unit = 2^-41 of the power of two above absmax.  The threshold form is used for
// absmax in [2^-40, 2^40] (float32 squares neither overflow nor vanish); outside, the kernels fall back to the direct
// per-element evaluation.
struct FixX {
  float p2a;     // 2^a with 2^20 <= absmax * 2^a < 2^21
  double unit;   // value of one fixed-point step = 2^-(a + 20)
};
constexpr float kBinnedMinAbs = 9.094947017729282e-13f;  // 2^-40
constexpr float kBinnedMaxAbs = 1099511627776.0f;        // 2^40
ADMMQ_HD bool binned_range_ok(float absmax) { return absmax >= kBinnedMinAbs && absmax <= kBinnedMaxAbs; }
ADMMQ_HD FixX make_fix_x(float absmax) {
  const int e = (int)((f32_bits(absmax) >> 23) & 0xffu) - 127;  // absmax = 1.m * 2^e
  FixX f;
  f.p2a = bits_f32((unsigned int)(20 - e + 127) << 23);
  // 2^-(a + 20) = 2^(e - 40) as a double
  unsigned long long db = (unsigned long long)(e - 40 + 1023) << 52;
#if defined(__CUDA_ARCH__)
  f.unit = __longlong_as_double((long long)db);
#else
  std::memcpy(&f.unit, &db, 8);
#endif
  return f;
}
// x -> rne(x * 2^(a + 20)) without the conversion pipe: two magic-number roundings (|x * 2^a| < 2^21)
ADMMQ_HD long long fix_x(float x, const FixX& f) {
  const float MAGIC = 12582912.0f;  // 1.5 * 2^23, bits 0x4B400000
  const float xs = mul_rn(x, f.p2a);                 // exact (power of two)
  const float hb = add_rn(xs, MAGIC);                // integer part (rne) in the low mantissa bits
  const float r = sub_rn(xs, sub_rn(hb, MAGIC));     // exact, |r| <= 0.5
  const float lb = fma_rn(r, 1048576.0f, MAGIC);     // rne(r * 2^20)
  const long long th = (long long)((int)f32_bits(hb) - 0x4B400000);
  const long long tl = (long long)((int)f32_bits(lb) - 0x4B400000);
  return th * 1048576ll + tl;
}

// sum of squares of four consecutive elements as the kernels form it (float32 FMA chain, then float64)
ADMMQ_HD float sumsq4(float a, float b, float c, float d) {
  return fma_rn(d, d, fma_rn(c, c, fma_rn(b, b, mul_rn(a, a))));
}

// contribution of threshold j of one candidate to the per-candidate sum, in float64:
//   C (y_lo^2 - y_hi^2) - 2 P unit (y_lo - y_hi),   y_lo = fl32(level * s), y_hi = fl32((level + 1) * s)
ADMMQ_HD double threshold_term(float scale, float level, long long count, long long psum, double xunit) {
  const double ylo = (double)mul_rn(level, scale), yhi = (double)mul_rn(level + 1.0f, scale);
  const double d1 = ylo - yhi;
  return (double)count * (d1 * (ylo + yhi)) - 2.0 * ((double)psum * xunit) * d1;
}
// closing term n y_top^2 - 2 P_tot unit y_top, y_top = fl32((q - 1) * s)
ADMMQ_HD double closing_term(float scale, float top_level, long long n, long long ptot, double xunit) {
  const double y = (double)mul_rn(top_level, scale);
  return (double)n * (y * y) - 2.0 * ((double)ptot * xunit) * y;
}

// --- the other tensor_* schemes ---------------------------------------------------------
struct QParams {
  int scheme;   // ADMMQ_Q_*
  float scale;  // mse/symmetric/affine: grid scale; minmax: (max - min)
  float aux;    // affine: zero point; minmax: min
  float n;      // minmax: 2^bits - 1
  int bits;
};

// min_max_quantize, source/quantization.py:48-66
ADMMQ_HD float minmax_value(float x, const QParams& p, float& level) {
  if (p.bits == 1) {
    const float sgn = (x > 0.0f ? 1.0f : 0.0f) - (x < 0.0f ? 1.0f : 0.0f);  // torch.sign = (0<x)-(x<0)
    level = sgn;
    return sub_rn(sgn, 1.0f);
  }
  const float unit = div_rn(sub_rn(x, p.aux), p.scale);
  const float k = floorf(add_rn(mul_rn(unit, p.n), 0.5f));
  level = k;
  return add_rn(div_rn(mul_rn(k, p.scale), p.n), p.aux);
}

// tensor_affine, source/quantization.py:97-106
ADMMQ_HD float affine_value(float x, const QParams& p, const Levels& L, float& code) {
  float k = add_rn(rint_rn(div_rn(x, p.scale)), p.aux);
  k = add_rn(fminf(fmaxf(k, L.lo), L.hi), 0.0f);  // `.to(int)` (:105) has no negative zero
  code = k;
  return mul_rn(sub_rn(k, p.aux), p.scale);
}

// zero point of tensor_affine (:102-103): clamp(int(-q - int(tmin/scale)), -q, q-1), `.int()` truncates
ADMMQ_HD float affine_zero_point(float tmin, float scale, const Levels& L) {
  const float t = truncf(div_rn(tmin, scale));
  float zp = truncf(sub_rn(L.lo, t));
  zp = fminf(fmaxf(zp, L.lo), L.hi);
  return zp;
}

}  // namespace admmq
