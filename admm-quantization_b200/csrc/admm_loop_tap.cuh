// admm_loop_tap.cuh - the ADMM inner loop (source/admm.py:55-65) for the TAP FACTOR of a wide convolution: few rows
// (the 3 x 3 = 9 kernel taps, I <= 9) and a large rank (R = 278 ... 1365), on ONE THREAD-BLOCK CLUSTER of 8 CTAs.
//
// In the general kernel (admm_loop.cu) such a factor is pure latency: ~10 k elements spread over the layer's 30+ CTAs
// cost 27 us per iteration - three device-wide barriers, a few hundred elements per CTA, every CTA repeating the 3000
// thresholds of the clip search.  Here the COLUMNS of the factor and the CANDIDATES of the clip search are split over
// the CTAs of a cluster (admm_loop_cluster.cuh splits rows; with 9 rows that would leave CTAs idle and make every CTA
// stream the whole inverse):
//   P1  every CTA holds the whole right-hand side (I x Rp floats) in shared memory and forms H_ls = RHS . Minv for ITS
//       column strip: thread (column group, k-slice) accumulates 4 columns x I rows over its k-slice with float32 FMAs,
//       rows of Minv read as float4 from L2 (the strip of the inverse, 0.65 MB at R = 1141, is the only global traffic
//       of the iteration); k-slices summed in fixed order; V = H_ls - U and the min / max keys go to EVERY CTA's shared
//       memory                                                              -- cluster barrier 1
//   P2  every CTA histograms and sorts ALL elements (replicated) and evaluates the thresholds of ITS candidates only;
//       a candidate's sum is broadcast                                      -- cluster barrier 2
//   P3  argmin, H = Q(V), U += H - H_ls, residual sums for the own columns; the new right-hand side of the own columns
//       and the residual partial sums go to every CTA                       -- cluster barrier 3, exit test
// Same float32 recipe as the general kernel's column-strip product up to the order of the k-slices.
#pragma once
#include <cstddef>
#include "admm_loop_cluster.cuh"

namespace admmq {

constexpr int kTapMaxElems = 12288;   // I * Rp (right-hand side, V, sorted elements: 48 KB each)
constexpr int kTapMaxRows = 9;      // the 3 x 3 taps (the k-slice scratch of P1 is sized for 9 rows x 512 threads)
constexpr int kTapMaxOwn = 3072;      // I * (columns of one CTA)
constexpr int kTapCtas = 8;

struct __align__(16) TapSmem {
  float rhs[kTapMaxElems];      // the whole right-hand side, row pitch Rp (pad columns zero)
  float Vall[kTapMaxElems];     // all elements of V, dense (row pitch R), written by every CTA of the cluster
  float X[kTapMaxElems];        // elements grouped by bin during P2; k-slice partial sums during P1 (with cnt .. shi)
  unsigned int cnt[kResBins];
  unsigned int slo[kResBins];
  unsigned int shi[kResBins];
  float U[kTapMaxOwn];          // own columns: [row][own column]
  float Hls[kTapMaxOwn];
  unsigned long long acc[kMaxCandidates];
  unsigned long long cand[kMaxCandidates];
  float scale[kMaxCandidates];
  double slots[kTapCtas][4];
  unsigned int keys[kTapCtas][2];
  unsigned long long wsum[kWarps];
  unsigned int wcnt[kWarps];
  unsigned int wkey[2][kWarps];
  double red[4][kWarps];
  unsigned long long best[kWarps];
};
static_assert(sizeof(TapSmem) <= 227 * 1024, "tap-factor loop state must fit the shared memory of one SM");
// the k-slice partial sums of P1 (kThreads x MI x 4 floats) live in X + cnt + slo + shi, which are dead during P1
static_assert((size_t)kThreads * kTapMaxRows * 4 * sizeof(float) <= sizeof(float) * kTapMaxElems + 3 * sizeof(unsigned int) * kResBins,
              "P1 scratch must fit the arrays it aliases");
static_assert(offsetof(TapSmem, cnt) == offsetof(TapSmem, X) + sizeof(float) * kTapMaxElems &&
              offsetof(TapSmem, shi) == offsetof(TapSmem, cnt) + 2 * sizeof(unsigned int) * kResBins, "X, cnt, slo, shi are contiguous");

template <int MI>
__global__ void __launch_bounds__(kThreads, 1) k_admm_loop_tap(const ResidentParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TapSmem& sm = *reinterpret_cast<TapSmem*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int I = p.I, R = p.R, Rp = p.Rp, n = I * R;
  const int rank = (int)cl_rank(), C = (int)cl_size();
  admmq_loop_report rep;
  rep.iterations = 0;
  rep.status = 0;
  rep.rho = *p.rho;
  rep.scale = 0.0f;
  rep.r = 0.0f;
  rep.s = 0.0f;
  rep.best_index = -1;
  rep.absmax = 0.0f;
  rep.phase_ns[0] = rep.phase_ns[1] = rep.phase_ns[2] = rep.phase_ns[3] = 0ull;
  if (p.inv_status != nullptr && *p.inv_status != 0) {  // uniform over the cluster: nobody reaches a barrier
    rep.status = *p.inv_status;
    if (rank == 0 && t == 0) *p.report = rep;
    return;
  }
  const unsigned long long t_begin = global_ns();
  unsigned long long t_mark = t_begin;
  auto lap = [&](int phase) {
    const unsigned long long now = global_ns();
    rep.phase_ns[phase] += now - t_mark;
    t_mark = now;
  };
  const float rho = rep.rho;
  const Levels L = make_levels(p.bits);
  const float qnan = __int_as_float(0x7fc00000);
  // own columns [n0, n1) (a multiple of four per CTA) and own candidates [c0, c1)
  const int cps = ((R + C - 1) / C + 3) / 4 * 4;
  const int n0 = min(R, rank * cps), n1 = min(R, n0 + cps), w = n1 - n0;
  const int own = I * w;
  const int cand_per = (p.Nc + C - 1) / C;
  const int c0 = min(p.Nc, rank * cand_per), c1 = min(p.Nc, c0 + cand_per);
  // ---- the whole right-hand side F + rho (H + U) for the first iteration (:56), own U
  for (int e = t; e < I * Rp; e += kThreads) {
    const int i = e / Rp, c = e - i * Rp;
    sm.rhs[e] = (c < R) ? add_rn(p.F[i * R + c], mul_rn(rho, add_rn(p.H[i * R + c], p.U[i * R + c]))) : 0.0f;
  }
  for (int e = t; e < own; e += kThreads) {
    const int i = e / w, c = e - i * w;
    sm.U[e] = p.U[i * R + n0 + c];
  }
  cl_sync();   // every CTA of the cluster is running (remote shared memory may be written from here on)
  float* part = sm.X;   // k-slice partial sums: [thread][MI][4]
  const int cg = (w + 3) >> 2;                 // float4 column groups of the own strip (n0 + 4 cg <= Rp)
  const int ks = cg > 0 ? kThreads / cg : 0;   // k-slices
  const int s_id = cg > 0 ? t / cg : 0, c_id = cg > 0 ? t - s_id * cg : 0;
  int done = 0;
  for (int j = 1; j < p.max_iter; ++j) {  // range(1, max_iter), :55
    // ---------------- P1: H_ls = RHS . Minv for the own column strip
    unsigned int kmax = 0u, kinv = 0u;
    if (cg > 0) {
      float acc[MI][4];
#pragma unroll
      for (int i = 0; i < MI; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
      if (s_id < ks) {
        const float* mp = p.Minv + n0 + 4 * c_id;
        // groups of four consecutive rows of Minv (k = 4 g .. 4 g + 3, g = s, s + ks, ...): the four RHS values of a factor
        // row are one 16-byte shared-memory load for 16 FMAs; two groups in flight
        constexpr int kG = 2;
        const int ngroups = (R + 3) >> 2;
        for (int gq = s_id; gq < ngroups; gq += ks * kG) {
          float4 m[kG][4];
#pragma unroll
          for (int u = 0; u < kG; ++u) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int k = 4 * (gq + u * ks) + q;
              m[u][q] = (k < R) ? __ldg(reinterpret_cast<const float4*>(mp + (size_t)k * Rp)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < kG; ++u) {
            const int g = gq + u * ks;
            if (g < ngroups) {
#pragma unroll
              for (int i = 0; i < MI; ++i) {
                if (i < I) {
                  const float4 r4 = *reinterpret_cast<const float4*>(&sm.rhs[i * Rp + 4 * g]);
                  const float r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    acc[i][0] = fmaf(r[q], m[u][q].x, acc[i][0]);
                    acc[i][1] = fmaf(r[q], m[u][q].y, acc[i][1]);
                    acc[i][2] = fmaf(r[q], m[u][q].z, acc[i][2]);
                    acc[i][3] = fmaf(r[q], m[u][q].w, acc[i][3]);
                  }
                }
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
          *reinterpret_cast<float4*>(&part[(size_t)t * (MI * 4) + i * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      }
    }
    __syncthreads();
    for (int o = t; o < own; o += kThreads) {
      const int i = o / w, nn = o - i * w;
      const int cc = nn >> 2, q = nn & 3;
      // fixed order: four interleaved partial sums over the k-slices, then ((h0 + h1) + (h2 + h3))
      float h4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      const float* rp = &part[(size_t)cc * (MI * 4) + i * 4 + q];
      int sl = 0;
      for (; sl + 3 < ks; sl += 4) {
#pragma unroll
        for (int a = 0; a < 4; ++a) h4[a] = add_rn(h4[a], rp[(size_t)((sl + a) * cg) * (MI * 4)]);
      }
      for (; sl < ks; ++sl) h4[sl & 3] = add_rn(h4[sl & 3], rp[(size_t)(sl * cg) * (MI * 4)]);
      const float h = add_rn(add_rn(h4[0], h4[1]), add_rn(h4[2], h4[3]));
      sm.Hls[o] = h;
      const float v = sub_rn(h, sm.U[o]);   // V = H_ls - U (:59)
      const unsigned int key = float_key(v);
      kmax = max(kmax, key);
      kinv = max(kinv, ~key);
      const int e = i * R + n0 + nn;
      for (int rk = 0; rk < C; ++rk) cl_store(cl_map(&sm.Vall[e], (unsigned int)rk), v);
    }
    kmax = warp_max_u32(kmax);
    kinv = warp_max_u32(kinv);
    if (lane == 0) {
      sm.wkey[0][warp] = kmax;
      sm.wkey[1][warp] = kinv;
    }
    __syncthreads();
    if (t < 2 * C) {   // this CTA's keys into slot `rank` of every CTA
      unsigned int k = 0u;
      for (int wq = 0; wq < kWarps; ++wq) k = max(k, sm.wkey[t & 1][wq]);
      cl_store(cl_map(&sm.keys[rank][t & 1], (unsigned int)(t >> 1)), k);
    }
    cl_sync();   // ---- barrier 1: V and the keys are everywhere
    kmax = 0u;
    kinv = 0u;
    for (int rk = 0; rk < C; ++rk) {
      kmax = max(kmax, sm.keys[rk][0]);
      kinv = max(kinv, sm.keys[rk][1]);
    }
    lap(0);
    // ---------------- P2
    const float tmax = key_float(kmax), tmin = key_float(~kinv);
    float absmax = fmaxf(fabsf(tmin), fabsf(tmax));
    if (tmin != tmin || tmax != tmax) absmax = qnan;
    rep.absmax = absmax;
    rep.iterations = j;
    done = j;
    QParams qp;
    qp.scheme = p.scheme;
    qp.bits = p.bits;
    qp.aux = 0.0f;
    qp.n = 0.0f;
    qp.scale = 0.0f;
    bool degenerate = false;
    if (p.scheme == ADMMQ_Q_MSEMINMAX_SYMMETRIC) {
      degenerate = !(absmax > 0.0f) || isinf(absmax);
      if (!degenerate) {
        if (binned_range_ok(absmax)) {
          cl_candidate_sums(sm, n, absmax, p.Nc, c0, c1, L, p.bits);
        } else {
          const ClipGrid g = make_clip_grid(absmax, p.Nc);
          const double unit_inv = fixed_point_unit_inv((double)n, absmax);
          for (int c = c0 + t; c < c1; c += kThreads) {
            const float s = scale_of(clip_candidate(g, c), L);
            double tot = 0.0;
            for (int e = 0; e < n; ++e) tot += (double)sqerr_exact(sm.Vall[e], s, L);
            sm.acc[c] = (unsigned long long)__double2ll_rn(tot * unit_inv);
          }
          __syncthreads();
        }
        for (int c = c0 + t; c < c1; c += kThreads) {
          const unsigned long long v = sm.acc[c];
          for (int rk = 0; rk < C; ++rk) cl_store(cl_map(&sm.cand[c], (unsigned int)rk), v);
        }
      }
    } else {
      degenerate = (absmax != absmax) || isinf(absmax);
      qp = params_from_minmax(p.scheme, p.bits, tmin, tmax, L);
    }
    cl_sync();   // ---- barrier 2: every candidate's sum is everywhere (nobody reads V / the sorted array any more)
    if (p.scheme == ADMMQ_Q_MSEMINMAX_SYMMETRIC && !degenerate) {
      const double unit = fixed_point_unit((double)n, absmax);
      const float nf = (float)n;
      unsigned long long best = ~0ull;
      for (int c = t; c < p.Nc; c += kThreads) {
        const float mse = mse_from_fixed((long long)sm.cand[c], unit, nf);
        best = min(best, ((unsigned long long)float_key(mse) << 32) | (unsigned int)c);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (lane == 0) sm.best[warp] = best;
      __syncthreads();
      unsigned long long b = sm.best[0];
#pragma unroll
      for (int wq = 1; wq < kWarps; ++wq) b = min(b, sm.best[wq]);
      rep.best_index = (int)(b & 0xffffffffu);
      qp.scale = scale_of(clip_candidate(make_clip_grid(absmax, p.Nc), rep.best_index), L);
    }
    rep.scale = qp.scale;
    lap(1);
    // ---------------- P3: H = Q(V), U += H - H_ls, residual sums, next RHS (to every CTA) - own columns
    float f0 = 0.0f, f1 = 0.0f, f2 = 0.0f, f3 = 0.0f;
    double sums[4] = {0.0, 0.0, 0.0, 0.0};
    int cnt = 0;
    for (int o = t; o < own; o += kThreads) {
      const int i = o / w, nn = o - i * w;
      const int e = i * R + n0 + nn;
      const float hp = p.H[e];
      const float fv = __ldg(p.F + e);
      const float hls = sm.Hls[o], u = sm.U[o];
      const float v = sub_rn(hls, u);
      float code = 0.0f;
      const float hq = degenerate ? qnan : quantize_value(v, qp, L, code);   // H = Q(H_ls - U)   (:59)
      const float d1 = sub_rn(hq, hls);
      const float un = add_rn(u, d1);                                        // U += H - H_ls     (:60)
      const float d2 = sub_rn(hq, hp);
      f0 = fmaf(d1, d1, f0);  // sum (H - H_ls)^2     (:62)
      f1 = fmaf(hq, hq, f1);  // sum H^2
      f2 = fmaf(d2, d2, f2);  // sum (H - H_prev)^2   (:63)
      f3 = fmaf(un, un, f3);  // sum U^2
      p.H[e] = hq;
      sm.U[o] = un;
      const float rnew = add_rn(fv, mul_rn(rho, add_rn(hq, un)));
      for (int rk = 0; rk < C; ++rk) cl_store(cl_map(&sm.rhs[i * Rp + n0 + nn], (unsigned int)rk), rnew);
      if (p.codes != nullptr) p.codes[e] = (int8_t)code;
      if (++cnt == 16) {
        sums[0] += (double)f0;
        sums[1] += (double)f1;
        sums[2] += (double)f2;
        sums[3] += (double)f3;
        f0 = f1 = f2 = f3 = 0.0f;
        cnt = 0;
      }
    }
    sums[0] += (double)f0;
    sums[1] += (double)f1;
    sums[2] += (double)f2;
    sums[3] += (double)f3;
    if (degenerate) {  // uniform: the reference would carry NaN through every remaining iteration
      rep.status |= ADMMQ_ST_NONFINITE;
      rep.r = qnan;
      rep.s = qnan;
      break;
    }
    cl_sum4(sums, sm);
    if (t < 4 * C) cl_store(cl_map(&sm.slots[rank][t & 3], (unsigned int)(t >> 2)), sums[t & 3]);
    cl_sync();   // ---- barrier 3: the next right-hand side and the residual sums are everywhere
    double tot[4] = {0.0, 0.0, 0.0, 0.0};
    for (int rk = 0; rk < C; ++rk)
#pragma unroll
      for (int q = 0; q < 4; ++q) tot[q] += sm.slots[rk][q];
    rep.r = div_rn((float)tot[0], (float)tot[1]);
    rep.s = div_rn((float)tot[2], (float)tot[3]);
    lap(2);
    if (rep.r < p.eps && rep.s < p.eps) {   // exit test (:62-65), evaluated identically by every CTA
      rep.status |= ADMMQ_ST_CONVERGED;
      break;
    }
  }
  cl_sync();   // no CTA leaves while its shared memory may still be written
  rep.iterations = done;
  for (int o = t; o < own; o += kThreads) {
    const int i = o / w, nn = o - i * w;
    p.U[i * R + n0 + nn] = sm.U[o];
  }
  rep.phase_ns[3] = global_ns() - t_begin;
  if (rank == 0 && t == 0) *p.report = rep;
}

// Eligible: few rows, everything fits the shared-memory arrays, a budget of at least one cluster - and a rank up to
// kTapMaxRank: the product streams the CTA's strip of the inverse with float32 FMAs, 1.5 MFMA per CTA at R = 1141, which
// eight SMs do slower than the general kernel's thirty-odd.  Measured per iteration (9 x R; cluster of 8 against the
// general kernel on 7 / 33 CTAs): R = 278: 13.0 us against 23.1, R = 566: 19.7 against 28.1 / 26.9, R = 1141: 38.0
// against 49.3 / 28.0.
constexpr int kTapMaxRank = 640;
inline bool tap_cluster_fits(int I, int R, int Rp, int num_attempts, int budget) {
  if (I > kTapMaxRows || R > kTapMaxRank || budget < kTapCtas || num_attempts > kMaxCandidates) return false;
  const int cps = ((R + kTapCtas - 1) / kTapCtas + 3) / 4 * 4;
  return (long long)I * Rp <= kTapMaxElems && (long long)I * cps <= kTapMaxOwn && cps >= 4;
}

}  // namespace admmq
