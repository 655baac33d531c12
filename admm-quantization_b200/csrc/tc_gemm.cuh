// tc_gemm.cuh - 3xTF32 tile product on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   D[128 x BN] (TMEM, float32) = A[128 x K] . B[BN x K]^T          A, B: float32, K-major
//
// Every float32 operand x is split as x = hi + lo + O(2^-22 |x|) with hi = tf32_rn(x), lo = tf32_rn(x - hi), and
// the tile is accumulated as lo.hi + hi.lo + hi.hi (small terms first) by three tcgen05.mma.kind::tf32
// per 8-deep k-step into one TMEM accumulator, which keeps float32-level accuracy (the dropped lo.lo term
// and the split residuals are 2^-22 relative) at 1/3 of the TF32 rate.
//
// Operand paths.  Measured on B200: with BOTH operands in shared memory one tcgen05.mma of 128 x N x 8 tf32 costs
// ~90 cycles for every N <= 64 - the tensor core re-reads the 128-row A slab (32 bytes from each of 128 rows) for
// every instruction - so narrow tiles, which this path needs to keep 148 SMs busy on 512 x 1141 outputs, were
// MMA-issue bound.  The A operand therefore lives in TENSOR MEMORY (TS form of tcgen05.mma): lane = tile row,
// column = k, written there by the producers with tcgen05.st, and only the small B slab (N rows x 32 bytes) is read
// from shared memory per instruction.
//
// Staging (512 threads, all 16 warps produce): warp w owns TMEM lanes 32 * (w % 4) .. +31 (the hardware restricts a
// warp to its lane quarter) and a share of the k-columns of each 32-deep K-block.  Operands are fetched from
// global/L2 ONCE as float32 by TMA (cp.async.bulk.tensor, 128B swizzle, out-of-range rows/columns zero-filled by the
// hardware; 4 K-blocks in flight; LDGSTS-based fetching measured ~1000 cycles per 20 KB K-block and was the
// bottleneck) into a raw ring in shared memory whose swizzle makes the row-per-lane read-back conflict free;
// raw_full[d] (expect_tx / complete_tx) publishes a K-block to the CTA.  Each producer thread then splits the chunks of
// ITS row into hi/lo in registers and
// stores A to TMEM (tcgen05.st) and B to shared memory in the canonical K-major SWIZZLE_128B layout (row r of a
// K-block = 128 bytes; 16-byte chunk c of row r is stored at chunk c ^ (r & 7); 8-row groups 1024 bytes apart).
// full[s] (one arrival per producer warp) hands a stage to warp 15, whose elected lane issues the 12 MMAs of the
// K-block, commits them to free[s] (which hands the stage back once the tensor pipe has consumed it) and then
// re-arms the raw slot the producers have just finished with by issuing the TMA loads of K-block k + 4.
#pragma once
#include <cuda.h>
#include <cstdint>
#include "common.cuh"

namespace admmq {
namespace tc {

constexpr int kThreadsTC = 512;
constexpr int kMmaWarp = 15;     // its lane 0 also issues the MMAs
constexpr int kAtomK = 32;       // floats per 128-byte swizzle row
constexpr int kAtoms = 2;        // swizzle atoms along K per K-block
constexpr int kBlockK = kAtomK * kAtoms;  // 64: halves the per-block synchronisation cost of a 32-deep block
constexpr int kUmmaK = 8;        // k per tcgen05.mma.kind::tf32
constexpr int kMaxStages = 3;    // operand stages (hi/lo, swizzled): 3 for BN <= 32, 2 for BN = 64 (TMEM budget)
constexpr int kTileM = 128;

constexpr int kRawDepth = 3;     // K-blocks of raw float32 in flight (TMA)
constexpr int kStageCols = 2 * kBlockK;  // TMEM columns of one A stage: hi [0, 64), lo [64, 128)
constexpr int kTmemCols = 512;   // 3 accumulators + stages x 128: 96 + 3 x 128 = 480 (BN <= 32), 192 + 2 x 128 = 448 (BN = 64)

template <int BN>
struct TileSmem {
  static_assert(BN == 16 || BN == 32 || BN == 64, "BN must be 16, 32 or 64");
  static constexpr int kStages = (BN == 64) ? 2 : 3;
  static constexpr int kAccStride = (BN == 64) ? 64 : 32;        // TMEM columns between the three accumulators
  static constexpr int kAccCols = 3 * kAccStride;                // hi.hi at +0, hi.lo at +stride, lo.hi at +2 stride
  static_assert(kAccCols + kStages * kStageCols <= kTmemCols, "tensor memory budget");
  static constexpr int kAAtomBytes = kTileM * 128;               // one 32-deep atom of A as raw float32
  static constexpr int kBAtomBytes = BN * 128;
  static constexpr int kABytes = kAtoms * kAAtomBytes;           // one K-block of A
  static constexpr int kBBytes = kAtoms * kBAtomBytes;
  static constexpr int kStageBytes = 2 * kBBytes;                // B_hi, B_lo (A goes to tensor memory)
  static constexpr int kRawBytes = kABytes + kBBytes;            // one K-block of raw float32
  static constexpr int kBytes = kStages * kStageBytes + kRawDepth * kRawBytes + 1024;  // + slack for 1024-byte alignment
  static_assert(kBytes <= 227 * 1024, "tile does not fit in shared memory");
};

struct Pipe {  // lives in shared memory (static), one per CTA
  unsigned long long raw_full[kRawDepth];  // cp.async data of a K-block has landed (one arrival per thread)
  unsigned long long stage_full[kMaxStages];
  unsigned long long stage_free[kMaxStages];
  unsigned long long tile_done;
  unsigned int tmem_base;
  unsigned int pad;
};

struct PipeState {  // per-thread copy, uniform across the CTA
  unsigned int raw_uses[kRawDepth];  // how often each raw slot has been filled so far
  unsigned int uses[kMaxStages];  // how often each stage has been filled / consumed so far
  unsigned int tiles;          // commits issued so far on tile_done
  unsigned long long nz2;      // (-0.0f, -0.0f) as a run-time value (see split2)
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
// arrive on `bar` once every cp.async issued so far by this thread has landed (the barrier's count includes it)
__device__ __forceinline__ void cp_async_arrive(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA: one arrival + `bytes` expected on `bar`, then a 2-D tile load (coordinates: c0 = innermost = column, c1 = row)
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned int dst_smem, const CUtensorMap* tmap, int c0, int c1,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst_smem), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// true in exactly one (always the same) lane of a fully active warp; the surrounding code stays warp-uniform,
// which lets ptxas keep MMA descriptors in uniform registers instead of broadcasting them lane by lane
__device__ __forceinline__ bool elect_one() {
  unsigned int pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0u;
}
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
// TS form: A from tensor memory (lane = row, column = k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(unsigned int d_tmem, unsigned int a_tmem, unsigned long long b_desc,
                                             unsigned int idesc, unsigned int accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 8 consecutive columns of the calling thread's TMEM lane
__device__ __forceinline__ void tmem_store8(unsigned int taddr, const float v[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_store_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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

// hi/lo split of two packed float32 values in four FMA-pipe instructions (the conversion is issue bound: the integer
// form costs five half-rate ALU operations per element, cvt.rna.tf32 runs on the quarter-rate conversion pipe):
//   Veltkamp/Dekker: t = x * (2^13 + 1); hi = t - (t - x) is x rounded to nearest to 11 significant bits, i.e.
//   exactly a tf32 number; lo = x - hi is exact in float32 and |lo| <= 2^-11 |x|.
// lo keeps up to 13 significant bits; the tensor core reads the upper 19 bits of each 32-bit operand, so lo enters the
// product truncated to tf32 (error <= 2^-21 |x|, same order as the dropped lo.lo term).
// ptxas contracts a packed multiply into each packed add/sub that consumes it (it even duplicates the multiply), which
// turns hi = t - (t - x) into fma(x, c, -(fma(x, c, -x))) = x.  The product is therefore formed as fma(x, c, nz) with
// nz = (-0.0f, -0.0f) taken from a kernel parameter, which cannot be folded away.
__device__ __forceinline__ void split2(unsigned long long x, unsigned long long nz, unsigned long long& hi,
                                       unsigned long long& lo) {
  unsigned long long t, u;
  const unsigned long long c = 0x4600040046000400ull;  // (8193.0f, 8193.0f)
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(x), "l"(c), "l"(nz));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(t), "l"(x));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(hi) : "l"(t), "l"(u));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(lo) : "l"(x), "l"(hi));
}
__device__ __forceinline__ void split4(const float4 v, unsigned long long nz, float4& hi, float4& lo) {
  const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(&v);
  ulonglong2 h, l;
  split2(x.x, nz, h.x, l.x);
  split2(x.y, nz, h.y, l.y);
  hi = *reinterpret_cast<const float4*>(&h);
  lo = *reinterpret_cast<const float4*>(&l);
}

// byte offset of 16-byte chunk c (0..7) of row r inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ unsigned int sw128(int r, int c) { return (unsigned int)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void pipe_setup(Pipe& pipe, PipeState& st, float neg_zero) {
  const unsigned int tmem_cols = kTmemCols;
  asm("mov.b64 %0, {%1, %1};" : "=l"(st.nz2) : "f"(neg_zero));
  if (threadIdx.x == 0) {
    for (int d = 0; d < kRawDepth; ++d) mbar_init(&pipe.raw_full[d], 1);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&pipe.stage_full[s], kThreadsTC / 32 - 1);
      mbar_init(&pipe.stage_free[s], 1);
    }
    mbar_init(&pipe.tile_done, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&pipe.tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int s = 0; s < kMaxStages; ++s) st.uses[s] = 0u;
  for (int d = 0; d < kRawDepth; ++d) st.raw_uses[d] = 0u;
  st.tiles = 0u;
#ifdef ADMMQ_TC_PROFILE
  for (int i = 0; i < 6; ++i) st.cyc[i] = 0;
#endif
}
__device__ __forceinline__ void pipe_teardown(Pipe& pipe) {
  const unsigned int tmem_cols = kTmemCols;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_free(pipe.tmem_base, tmem_cols);
}

// 4 consecutive columns of the calling thread's TMEM lane
__device__ __forceinline__ void tmem_store4(unsigned int taddr, const float4 v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
               "r"(__float_as_uint(v.w))
               : "memory");
}

// One 128 x BN tile: the 128 rows of A starting at a_row0 against the BN rows of B starting at b_row0, over K
// columns.  tmA / tmB are TMA descriptors of the two row-major float32 matrices with boxes {32, 128} and {32, BN}
// (make_operand_tmap); rows and columns outside the matrices read as zero.
// On return the accumulator is complete in TMEM (columns [0, BN) at pipe.tmem_base) and visible to every thread.
//
// Roles: warp 15 issues MMAs and TMA loads.  Warps 0..14 convert: warp w writes TMEM lanes 32 * (w % 4) .. +31; the
// 8 chunks (4 k-columns each) of a K-block row are split 2/2/2/2 over the four warps of quarters 0-2 and 3/3/2 over
// warps 3, 7, 11 of quarter 3; the B chunks go to the twelve warps of quarters 0-2.
template <int BN>
__device__ void tile_3xtf32(const CUtensorMap* tmA, int a_row0, const CUtensorMap* tmB, int b_row0, int K,
                            unsigned char* smem_tiles, Pipe& pipe, PipeState& st) {
  using TS = TileSmem<BN>;
  constexpr int kChunksB = BN * 8 * kAtoms;  // 16-byte chunks of B per K-block
  constexpr int kRowChunks = 8 * kAtoms;     // 16-byte chunks per row per K-block
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const unsigned int tiles_base = (smem_u32(smem_tiles) + 1023u) & ~1023u;
  unsigned char* tiles_ptr = smem_tiles + (tiles_base - smem_u32(smem_tiles));
  const unsigned int raw_base = tiles_base + (unsigned int)(TS::kStages * TS::kStageBytes);
  unsigned char* raw_ptr = tiles_ptr + (size_t)TS::kStages * TS::kStageBytes;
  const int nkb = (K + kBlockK - 1) / kBlockK;
  const unsigned int tmem = pipe.tmem_base;  // keep in a register: the asm memory clobbers would re-read it

  if (warp == kMmaWarp) {
    // ---------------- MMA + TMA warp: the whole warp runs the loop, one elected lane issues
    const unsigned int idesc = make_idesc<BN>();
    auto fetch = [&](int kb) {  // elected lane only
      const int d = kb % kRawDepth;
      const unsigned int slot = raw_base + (unsigned int)(d * TS::kRawBytes);
      mbar_expect_tx(&pipe.raw_full[d], (unsigned int)TS::kRawBytes);
#pragma unroll
      for (int a = 0; a < kAtoms; ++a) {
        tma_load_2d(slot + (unsigned int)(a * TS::kAAtomBytes), tmA, kb * kBlockK + a * kAtomK, a_row0, &pipe.raw_full[d]);
        tma_load_2d(slot + (unsigned int)(TS::kABytes + a * TS::kBAtomBytes), tmB, kb * kBlockK + a * kAtomK, b_row0,
                    &pipe.raw_full[d]);
      }
    };
    if (elect_one()) {
      for (int d = 0; d < kRawDepth; ++d)
        if (d < nkb) fetch(d);
    }
    __syncwarp();
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TS::kStages;
      { TC_T0(); mbar_wait(&pipe.stage_full[s], st.uses[s] & 1u); TC_ACC(3); }
      st.uses[s] += 1u;
      tc_fence_after();
      TC_T0();
      const unsigned int a_col = tmem + (unsigned int)(TS::kAccCols + s * kStageCols);
      const unsigned int sb = tiles_base + (unsigned int)(s * TS::kStageBytes);
      const unsigned int first = kb != 0 ? 1u : 0u;
      unsigned long long* free_bar = &pipe.stage_free[s];
      unsigned long long* done_bar = &pipe.tile_done;
      const bool last = kb + 1 == nkb;
      if (elect_one()) {
#pragma unroll
        for (int a = 0; a < kAtoms; ++a) {
          const unsigned long long b_hi = make_desc(sb + (unsigned int)(a * TS::kBAtomBytes));
          const unsigned long long b_lo = make_desc(sb + (unsigned int)(TS::kBBytes + a * TS::kBAtomBytes));
#pragma unroll
          for (int ks = 0; ks < kAtomK / kUmmaK; ++ks) {
            const unsigned long long adv = (unsigned long long)((ks * kUmmaK * 4) >> 4);  // +32 bytes per k-step
            const unsigned int a_hi = a_col + (unsigned int)(a * kAtomK + ks * kUmmaK), a_lo = a_hi + (unsigned int)kBlockK;
            // three independent accumulators, added small terms first by the epilogue: besides shortening the
            // dependent chain this keeps the lo terms from being absorbed into the large hi.hi sums
            const unsigned int acc = (a | ks) != 0 ? 1u : first;
            umma_tf32_ts(tmem + 2u * TS::kAccStride, a_lo, b_hi + adv, idesc, acc);
            umma_tf32_ts(tmem + 1u * TS::kAccStride, a_hi, b_lo + adv, idesc, acc);
            umma_tf32_ts(tmem, a_hi, b_hi + adv, idesc, acc);
          }
        }
        umma_commit(free_bar);
        if (last) umma_commit(done_bar);
        // every producer has finished reading raw slot kb % kRawDepth (stage_full completed): refill it
        if (kb + kRawDepth < nkb) fetch(kb + kRawDepth);
      }
      __syncwarp();
      TC_ACC(4);
    }
  } else {
    // ---------------- producers
    const int q = warp & 3, kgi = warp >> 2;
    const int ar = q * 32 + lane;  // tile row = TMEM lane of this thread
    // chunks (4 k-columns each) of the row this thread converts: 4/4/4/4 in quarters 0-2, 6/5/5 over warps 3, 7, 11
    const int c_begin = (q == 3) ? (kgi == 0 ? 0 : 1 + 5 * kgi) : kgi * (kRowChunks / 4);
    const int c_count = (q == 3) ? (kgi == 0 ? 6 : 5) : kRowChunks / 4;
    const unsigned int lane_addr = tmem + ((unsigned int)(q * 32) << 16);
    // B chunk ids handled by this thread: bi0, bi0 + 384, ... (warps of quarters 0-2)
    const int bi0 = (q == 3) ? kChunksB : (kgi * 3 + q) * 32 + lane;
    for (int kb = 0; kb < nkb; ++kb) {
      const int d = kb % kRawDepth;
      { TC_T0(); mbar_wait(&pipe.raw_full[d], (st.raw_uses[d] + (unsigned int)(kb / kRawDepth)) & 1u); TC_ACC(0); }  // K-block kb has landed
      const int s = kb % TS::kStages;
      { TC_T0(); if (st.uses[s] > 0u) { mbar_wait(&pipe.stage_free[s], (st.uses[s] - 1u) & 1u); tc_fence_after(); } TC_ACC(1); }
      TC_T0();
      const unsigned char* raw = raw_ptr + (size_t)d * TS::kRawBytes;
      const unsigned int col0 = lane_addr + (unsigned int)(TS::kAccCols + s * kStageCols);
#ifndef TC_EXP_NO_A
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (i < c_count) {
          const int c = c_begin + i;  // chunk c of the row: atom c / 8, chunk c % 8 within the atom's 128-byte row
          const float4 v = *reinterpret_cast<const float4*>(raw + (c >> 3) * TS::kAAtomBytes + sw128(ar, c & 7));
          float4 hi, lo;
          split4(v, st.nz2, hi, lo);
          tmem_store4(col0 + (unsigned int)(c * 4), hi);
          tmem_store4(col0 + (unsigned int)(kBlockK + c * 4), lo);
        }
      }
#endif
      unsigned char* stage = tiles_ptr + (size_t)s * TS::kStageBytes;
#ifndef TC_EXP_NO_B
      for (int bi = bi0; bi < kChunksB; bi += 384) {
        const int atom = bi / (BN * 8), lid = bi % (BN * 8);
        const unsigned int off = (unsigned int)(atom * TS::kBAtomBytes) + sw128(lid >> 3, lid & 7);
        const float4 v = *reinterpret_cast<const float4*>(raw + TS::kABytes + off);
        float4 hi, lo;
        split4(v, st.nz2, hi, lo);
        *reinterpret_cast<float4*>(stage + off) = hi;
        *reinterpret_cast<float4*>(stage + TS::kBBytes + off) = lo;
      }
#endif
#ifndef TC_EXP_NO_A
      tmem_store_wait();
#endif
      tc_fence_before();
#ifndef TC_EXP_NO_FENCE
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
#endif
      __syncwarp();
      if (lane == 0) mbar_arrive(&pipe.stage_full[s]);
      TC_ACC(2);
      st.uses[s] += 1u;
    }
  }
  // raw slot use counts advance identically in both roles (fills == consumptions)
  for (int kb = 0; kb < nkb; ++kb) st.raw_uses[kb % kRawDepth] += 1u;
  { TC_T0(); mbar_wait(&pipe.tile_done, st.tiles & 1u); TC_ACC(5); }
  st.tiles += 1u;
  tc_fence_after();
}

// Accumulator read-back for the calling warp: TMEM lanes 32 * (warp % 4) .. +31 (= tile rows), BN / 4 columns
// starting at (warp / 4) * BN / 4.  v[i] = D[row = 32 * (warp % 4) + lane][col0 + i] = (lo.hi + hi.lo) + hi.hi.
template <int N>
__device__ __forceinline__ void tmem_load(unsigned int taddr, unsigned int r[N]) {
  if constexpr (N == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
  } else if constexpr (N == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
  } else {
    static_assert(N == 16, "unsupported width");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
  }
}
template <int BN>
__device__ __forceinline__ void load_acc(const Pipe& pipe, float v[BN / 4], int& row, int& col0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, cgp = warp >> 2;
  row = q * 32 + lane;
  col0 = cgp * (BN / 4);
  const unsigned int taddr = pipe.tmem_base + ((unsigned int)(q * 32) << 16) + (unsigned int)col0;
  unsigned int hh[BN / 4], hl[BN / 4], lh[BN / 4];
  tmem_load<BN / 4>(taddr, hh);
  tmem_load<BN / 4>(taddr + 1u * TileSmem<BN>::kAccStride, hl);
  tmem_load<BN / 4>(taddr + 2u * TileSmem<BN>::kAccStride, lh);
#pragma unroll
  for (int i = 0; i < BN / 4; ++i) v[i] = (__uint_as_float(lh[i]) + __uint_as_float(hl[i])) + __uint_as_float(hh[i]);
}

// all accumulator reads of this tile are done: the next tile may overwrite TMEM
__device__ __forceinline__ void release_acc() {
  tc_fence_before();
  __syncthreads();
}

// Epilogue staging.  A thread holds one ROW of the accumulator (TMEM lane = row), so writing results straight to
// row-major global memory touches 32 different rows per warp instruction (4 useful bytes per 32-byte sector, measured
// ~6-8 k cycles per 128 x 64 tile).  acc_to_smem() parks the tile in shared memory instead - rows of BN + 4 floats
// (the pad keeps the per-row float4 stores conflict free) in the operand stage buffers, which are idle once tile_done
// has completed - releases the accumulator and returns the tile, so that the caller can walk it in row-contiguous
// float4 groups (thread t -> group t, t + 512, ...; group g = row g / (BN / 4), columns 4 (g % (BN / 4)) .. +3).  The
// caller must __syncthreads() after its last read and before the next tile_3xtf32().
template <int BN>
struct EpiTile {
  static constexpr int kLd = BN + 4;
  static constexpr int kGroupsPerRow = BN / 4;
  static constexpr int kGroups = kTileM * kGroupsPerRow;
  static_assert(kTileM * kLd * 4 <= TileSmem<BN>::kStages * TileSmem<BN>::kStageBytes, "staging tile must fit in the operand stages");
};
template <int BN>
__device__ __forceinline__ const float* acc_to_smem(const Pipe& pipe, unsigned char* smem_tiles) {
  const unsigned int tiles_base = (smem_u32(smem_tiles) + 1023u) & ~1023u;
  float* tile = reinterpret_cast<float*>(smem_tiles + (tiles_base - smem_u32(smem_tiles)));
  float v[BN / 4];
  int row, col0;
  load_acc<BN>(pipe, v, row, col0);
  float* dst = tile + row * EpiTile<BN>::kLd + col0;
#pragma unroll
  for (int c = 0; c < BN / 4; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
  release_acc();
  return tile;
}

// Host: TMA descriptor of a row-major float32 matrix (rows x cols, leading dimension ld floats, ld % 4 == 0, base
// 16-byte aligned) with box {32 columns, box_rows rows} and 128-byte swizzle.  Returns 0 or ADMMQ_E_CUDA.
int make_operand_tmap(CUtensorMap* out, const float* base, int rows, int cols, int ld, int box_rows);

}  // namespace tc
}  // namespace admmq
