// tc_gemm.cuh - 3xTF32 tile product on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   D[128 x BN] (TMEM, float32) = A[128 x K] . B[BN x K]^T          A, B: float32, K-major
//
// Every float32 operand x is split as x = hi + lo + O(2^-22 |x|) with hi = tf32_rn(x), lo = tf32_rn(x - hi), and
// the tile is accumulated as lo.hi + hi.lo + hi.hi (small terms first) by three tcgen05.mma.kind::tf32
// per 8-deep k-step into one TMEM accumulator, which keeps float32-level accuracy (the dropped lo.lo term
// and the split residuals are 2^-22 relative) at 1/3 of the TF32 rate.
//
// Staging (warp specialised, 512 threads): warps 0-14 are producers, warp 15 issues the MMAs.  Operands are read
// from global/L2 ONCE as float32 with cp.async (16-byte chunks, 3-4 K-blocks in flight) into a raw ring in
// shared memory; each producer thread then splits ITS OWN chunks into hi/lo in registers and writes them to the
// operand stage in the canonical K-major SWIZZLE_128B layout the UMMA descriptors expect (row r of a K-block =
// 128 bytes = 32 floats; 16-byte chunk c of row r is stored at chunk c ^ (r & 7); 8-row groups are 1024 bytes apart).
// full[s] (one arrival per producer warp) hands a stage to the MMA warp; tcgen05.commit on free[s] hands it back
// when the tensor pipe has read it, so the MMAs of K-block k overlap the staging of k+1, k+2.
#pragma once
#include <cstdint>
#include "common.cuh"

namespace admmq {
namespace tc {

constexpr int kThreadsTC = 512;
constexpr int kProducers = 480;  // warps 0..14; warp 15 issues the MMAs
constexpr int kMmaWarp = 15;
constexpr int kBlockK = 32;      // floats per K-block (one 128-byte swizzle row)
constexpr int kUmmaK = 8;        // k per tcgen05.mma.kind::tf32
constexpr int kStages = 3;       // operand stages (hi/lo, swizzled)
constexpr int kTileM = 128;

template <int BN>
struct TileSmem {
  static constexpr int kABytes = kTileM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;  // A_hi, A_lo, B_hi, B_lo
  static constexpr int kRawBytes = kABytes + kBBytes;            // one K-block of raw float32
  static constexpr int kChunks = (kABytes + kBBytes) / 16;       // 16-byte chunks per K-block
  static constexpr int kPerThread = (kChunks + kProducers - 1) / kProducers;
  static constexpr int kRawDepth = (BN >= 64) ? 3 : 4;           // K-blocks of raw float32 in flight (cp.async)
  static constexpr int kBytes = kStages * kStageBytes + kRawDepth * kRawBytes + 1024;  // + slack for 1024-byte alignment
  static_assert(kBytes <= 227 * 1024, "tile does not fit in shared memory");
};

struct Pipe {  // lives in shared memory (static), one per CTA
  unsigned long long stage_full[kStages];
  unsigned long long stage_free[kStages];
  unsigned long long tile_done;
  unsigned int tmem_base;
  unsigned int pad;
};

struct PipeState {  // per-thread copy, uniform across the CTA
  unsigned int uses[kStages];  // how often each stage has been filled / consumed so far
  unsigned int tiles;          // commits issued so far on tile_done
#ifdef ADMMQ_TC_PROFILE
  long long cyc[6];            // producer: cp.async wait, stage_free wait, convert; mma: full wait, issue; all: tile_done wait
#endif
};
#ifdef ADMMQ_TC_PROFILE
#define TC_T0() const long long _t0 = clock64()
#define TC_ACC(i) st.cyc[i] += clock64() - _t0
#else
#define TC_T0()
#define TC_ACC(i)
#endif

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
  const unsigned int addr = smem_u32(bar);
  unsigned int ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (ok == 0u);
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned int dst_smem, const void* src, bool valid) {
  const unsigned int sz = valid ? 16u : 0u;  // 0 source bytes => the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(unsigned int* slot, unsigned int cols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(unsigned int base, unsigned int cols) {  // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

__device__ __forceinline__ void umma_tf32(unsigned int d_tmem, unsigned long long a_desc, unsigned long long b_desc,
                                          unsigned int idesc, unsigned int accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = BN
template <int BN>
__device__ __forceinline__ constexpr unsigned int make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned int)(BN >> 3) << 17) | ((unsigned int)(kTileM >> 4) << 24);
}
// shared-memory matrix descriptor, K-major SWIZZLE_128B: LBO = 1 (unused), SBO = 1024 B between 8-row groups,
// version 1 (sm_100), layout type 2
__device__ __forceinline__ unsigned long long make_desc(unsigned int smem_addr) {
  return (unsigned long long)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ float tf32_rn(float x) {
  unsigned int r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi.x = tf32_rn(v.x); lo.x = tf32_rn(v.x - hi.x);
  hi.y = tf32_rn(v.y); lo.y = tf32_rn(v.y - hi.y);
  hi.z = tf32_rn(v.z); lo.z = tf32_rn(v.z - hi.z);
  hi.w = tf32_rn(v.w); lo.w = tf32_rn(v.w - hi.w);
}

// byte offset of 16-byte chunk c (0..7) of row r inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ unsigned int sw128(int r, int c) { return (unsigned int)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void pipe_setup(Pipe& pipe, PipeState& st, unsigned int tmem_cols) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&pipe.stage_full[s], kProducers / 32);
      mbar_init(&pipe.stage_free[s], 1);
    }
    mbar_init(&pipe.tile_done, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&pipe.tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int s = 0; s < kStages; ++s) st.uses[s] = 0u;
  st.tiles = 0u;
#ifdef ADMMQ_TC_PROFILE
  for (int i = 0; i < 6; ++i) st.cyc[i] = 0;
#endif
}
__device__ __forceinline__ void pipe_teardown(Pipe& pipe, unsigned int tmem_cols) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_free(pipe.tmem_base, tmem_cols);
}

// One 128 x BN tile: rows [a_row0, a_row0+128) of A (valid while < a_rows) against rows [b_row0, b_row0+BN) of B
// (valid while < b_rows); A and B are float32 row-major with leading dimensions lda/ldb (multiples of 4, 16-byte aligned
// rows).  K valid columns; columns in [K, round_up(K, 4)) must be readable zeros, nothing beyond is touched.
// On return the accumulator is complete in TMEM (pipe.tmem_base) and visible to every thread.
template <int BN>
__device__ void tile_3xtf32(const float* __restrict__ A, int lda, int a_row0, int a_rows, const float* __restrict__ B,
                            int ldb, int b_row0, int b_rows, int K, unsigned char* smem_tiles, Pipe& pipe, PipeState& st) {
  using TS = TileSmem<BN>;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const unsigned int tiles_base = (smem_u32(smem_tiles) + 1023u) & ~1023u;
  unsigned char* tiles_ptr = smem_tiles + (tiles_base - smem_u32(smem_tiles));
  const unsigned int raw_base = tiles_base + (unsigned int)(kStages * TS::kStageBytes);
  unsigned char* raw_ptr = tiles_ptr + (size_t)kStages * TS::kStageBytes;
  const int nkb = (K + kBlockK - 1) / kBlockK;
  if (warp != kMmaWarp) {
    // ---------------- producers: chunk ids t, t + 480, ...; ids < 1024 belong to A (row = id / 8), the rest to B
    auto issue = [&](int kb) {
      if (kb < nkb) {
        const unsigned int slot = raw_base + (unsigned int)((kb % TS::kRawDepth) * TS::kRawBytes);
#pragma unroll
        for (int i = 0; i < TS::kPerThread; ++i) {
          const int id = t + i * kProducers;
          if (id < TS::kChunks) {
            const int k4 = kb * kBlockK + (id & 7) * 4;
            if (id < kTileM * 8) {
              const int row = a_row0 + (id >> 3);
              const bool ok = (k4 < K) && (row < a_rows);
              cp_async16(slot + (unsigned int)id * 16u, ok ? (const void*)(A + (size_t)row * lda + k4) : (const void*)A, ok);
            } else {
              const int row = b_row0 + ((id - kTileM * 8) >> 3);
              const bool ok = (k4 < K) && (row < b_rows);
              cp_async16(slot + (unsigned int)id * 16u, ok ? (const void*)(B + (size_t)row * ldb + k4) : (const void*)B, ok);
            }
          }
        }
      }
      cp_async_commit();  // always: keeps the group count uniform
    };
#pragma unroll
    for (int d = 0; d < TS::kRawDepth; ++d) issue(d);
    for (int kb = 0; kb < nkb; ++kb) {
      { TC_T0(); cp_async_wait<TS::kRawDepth - 1>(); TC_ACC(0); }  // this thread's chunks of K-block kb have landed
      const int s = kb % kStages;
      { TC_T0(); if (st.uses[s] > 0u) mbar_wait(&pipe.stage_free[s], (st.uses[s] - 1u) & 1u); TC_ACC(1); }  // tensor pipe is done with stage s
      TC_T0();
      unsigned char* stage = tiles_ptr + (size_t)s * TS::kStageBytes;
      const unsigned char* raw = raw_ptr + (size_t)(kb % TS::kRawDepth) * TS::kRawBytes;
#pragma unroll
      for (int i = 0; i < TS::kPerThread; ++i) {
        const int id = t + i * kProducers;
        if (id < TS::kChunks) {
          const float4 v = *reinterpret_cast<const float4*>(raw + (size_t)id * 16);
          float4 hi, lo;
          split4(v, hi, lo);
          if (id < kTileM * 8) {
            const unsigned int off = sw128(id >> 3, id & 7);
            *reinterpret_cast<float4*>(stage + off) = hi;
            *reinterpret_cast<float4*>(stage + TS::kABytes + off) = lo;
          } else {
            const int bid = id - kTileM * 8;
            const unsigned int off = sw128(bid >> 3, bid & 7);
            *reinterpret_cast<float4*>(stage + 2 * TS::kABytes + off) = hi;
            *reinterpret_cast<float4*>(stage + 2 * TS::kABytes + TS::kBBytes + off) = lo;
          }
        }
      }
      issue(kb + TS::kRawDepth);     // refill the raw slot just consumed (only this thread reads these chunks)
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&pipe.stage_full[s]);
      st.uses[s] += 1u;
      TC_ACC(2);
    }
    cp_async_wait<0>();
  } else {
    // ---------------- MMA warp: one elected lane issues, the tensor pipe runs asynchronously
    const unsigned int idesc = make_idesc<BN>();
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages;
      { TC_T0(); mbar_wait(&pipe.stage_full[s], st.uses[s] & 1u); TC_ACC(3); }
      tc_fence_after();
      TC_T0();
      if (lane == 0) {
        const unsigned int sa = tiles_base + (unsigned int)(s * TS::kStageBytes);
        const unsigned long long a_hi = make_desc(sa), a_lo = make_desc(sa + TS::kABytes);
        const unsigned long long b_hi = make_desc(sa + 2 * TS::kABytes), b_lo = make_desc(sa + 2 * TS::kABytes + TS::kBBytes);
#pragma unroll
        for (int ks = 0; ks < kBlockK / kUmmaK; ++ks) {
          const unsigned long long adv = (unsigned long long)((ks * kUmmaK * 4) >> 4);  // +32 bytes per k-step
          umma_tf32(pipe.tmem_base, a_lo + adv, b_hi + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
          umma_tf32(pipe.tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
          umma_tf32(pipe.tmem_base, a_hi + adv, b_hi + adv, idesc, 1u);
        }
        umma_commit(&pipe.stage_free[s]);
        if (kb + 1 == nkb) umma_commit(&pipe.tile_done);
      }
      __syncwarp();
      st.uses[s] += 1u;
      TC_ACC(4);
    }
  }
  { TC_T0(); mbar_wait(&pipe.tile_done, st.tiles & 1u); TC_ACC(5); }
  st.tiles += 1u;
  tc_fence_after();
}

// Accumulator read-back for the calling warp: TMEM lanes 32 * (warp % 4) .. +31 (= tile rows), BN / 4 columns
// starting at (warp / 4) * BN / 4.  v[i] = D[row = 32 * (warp % 4) + lane][col0 + i].
template <int BN>
__device__ __forceinline__ void load_acc(const Pipe& pipe, float v[BN / 4], int& row, int& col0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, cgp = warp >> 2;
  row = q * 32 + lane;
  col0 = cgp * (BN / 4);
  const unsigned int taddr = pipe.tmem_base + ((unsigned int)(q * 32) << 16) + (unsigned int)col0;
  unsigned int r[BN / 4];
  if constexpr (BN == 32) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
  } else if constexpr (BN == 16) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
  } else {
    static_assert(BN == 64, "unsupported BN");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
  }
#pragma unroll
  for (int i = 0; i < BN / 4; ++i) v[i] = __uint_as_float(r[i]);
}

// all accumulator reads of this tile are done: the next tile may overwrite TMEM
__device__ __forceinline__ void release_acc() {
  tc_fence_before();
  __syncthreads();
}

}  // namespace tc
}  // namespace admmq
