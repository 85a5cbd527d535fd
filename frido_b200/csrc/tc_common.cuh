// Shared pieces of the tcgen05 engines (conv_tc.cu: per-tap operand tiles; conv_nf.cu: halo-resident operand with the
// GroupNorm / SPADE / SiLU of the input applied on load): tile constants, launch parameters, the work iterator, PTX
// wrappers and the epilogue warps' role.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace frido {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;             // fp32 elements per stage row = 128 bytes = one swizzle row
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_MAX_BN = 256;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;       // 16 KB
constexpr int TC_SMEM_BUDGET = 200 * 1024;          // operand ring
constexpr int TC_STG_BYTES = 4 * 4096;              // epilogue transpose staging, 4 KB per epilogue warp
constexpr int TC_CSUM_BYTES = 4 * 256 * 2 * 4;          // per-tile channel sum / sum-of-squares accumulators [img<=4][BN<=256][2]
constexpr int TC_SMEM_BYTES = TC_SMEM_BUDGET + TC_STG_BYTES + TC_CSUM_BYTES + 512 /*barriers*/ + 1024 /*align slack*/;  // < 227 KB
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // TMA, MMA, 8 epilogue warps
constexpr int TC_BF_ACC_STRIDE = 192;  // BF16x3: accumulators at TMEM columns 0 / 192 (BN <= 192) ...
constexpr int TC_BF_A_COL = 384;       // ... and the split A operand ring behind them: 32 columns (hi 16 | lo 16) per stage
constexpr int TC_BF_MAX_STAGES = 4;    // (512 - 384) / 32
constexpr int TC_SPLIT_WARPS = 4;
constexpr int TC_SPLIT_THREADS = TC_SPLIT_WARPS * 32;
constexpr int TC_THREADS_X3 = TC_THREADS + TC_SPLIT_THREADS;  // + splitter warps (error-compensated modes)

struct TcParams {
  int B, Hout, Wout, Cout;
  int c0, c1;           // channels per source (multiples of 32)
  int cx0, cx1;         // fused 1x1 side input (ResBlock skip_connection): extra K steps after the taps, from maps x0 | x1
  int ksize, pad, stride;
  int TW, TH, TB;       // tile = TW*TH*TB = 128 pixels (powers of two)
  int lTW, lTH;         // log2
  int tiles_x, tiles_y, tiles_b, tiles_n;
  int BN;
  int stages;           // depth of the smem ring (decoupled BF16x3 kernel: of the W ring)
  int a_stages;         // decoupled BF16x3 kernel: depth of the raw A-tile ring
  int w_batched;        // weights have a per-image leading dim
  const float* bias;
  const float* rowvec; long long rowvec_sb;
  const float* res;
  float alpha;
  int act;
  float* out; long long o_sb, o_sp, o_sn;
  int round_tf32;
  uint16_t* out_hi; uint16_t* out_lo;  // optional bf16 pair copy of the outputs
  double* csum;         // optional [B][Cout][2] per-channel sum / sum of squares of the stored outputs (GroupNorm fusion)
  // stream-K (under-filled launches): the (tile, k-step) iteration space is cut into equal contiguous ranges, one per CTA;
  // a range that covers only part of a tile's K loop parks its partial accumulator in sk_ws and the LAST contributor to
  // arrive (sk_cnt[tile]) sums the partials in CTA order (deterministic) and runs the normal epilogue.
  int sk;               // 1 = stream-K schedule, 0 = one whole tile per CTA at a time
  int sk_per;           // iterations (k-steps) per CTA
  float4* sk_ws;        // [2 * grid][BN/16][4][128] float4 partial-accumulator slots
  int* sk_cnt;          // [tiles] arrival counters, zero between launches
  // CTA-pair schedule (conv_tc2.cu, cta_group::2): a cluster of two CTAs takes the M tiles (2q, 2q+1) of one N tile; work items
  // are (q, n-tile) pairs, one per cluster at a time
  int pair;             // 1 = CTA-pair kernel
  int dbg;              // profiling aid (FRIDO_TC_DBG bit mask, results are WRONG with any bit set): 1 = no A-tile TMA, 2 = splitter
                        // does no arithmetic, 4 = splitter does not touch shared memory either
  const void* dbg_w;    // profiling aid (FRIDO_TC_DBG_BULKW=1, results are WRONG): W stages come as contiguous bulk copies from here
};

// Work iterator shared by all warp roles: yields (tile, [k0, k1)) segments in the same order everywhere.
struct SegIter {
  int sk, ksteps, total_tiles, stride, cur, end, pair, rank, tiles_n;
  __device__ __forceinline__ SegIter(const TcParams& p, int ksteps_, int total_tiles_)
      : sk(p.sk), ksteps(ksteps_), total_tiles(total_tiles_), stride((int)gridDim.x), pair(p.pair), rank(0), tiles_n(p.tiles_n) {
    if (pair) {  // items = (M-tile pair, N tile); cluster c starts at item c and strides by the number of clusters
      rank = (int)(blockIdx.x & 1);
      cur = (int)(blockIdx.x >> 1);
      stride = (int)(gridDim.x >> 1);
      total_tiles = total_tiles_ >> 1;  // items
      end = 0;
    } else if (sk) {
      cur = (int)blockIdx.x * p.sk_per;
      end = min(cur + p.sk_per, total_tiles_ * ksteps_);
    } else {
      cur = (int)blockIdx.x;
      end = 0;
    }
  }
  __device__ __forceinline__ bool next(int& tile, int& k0, int& k1) {
    if (pair) {
      if (cur >= total_tiles) return false;
      const int q = cur / tiles_n, nt = cur - q * tiles_n;
      tile = (2 * q + rank) * tiles_n + nt; k0 = 0; k1 = ksteps; cur += stride;
      return true;
    }
    if (!sk) {
      if (cur >= total_tiles) return false;
      tile = cur; k0 = 0; k1 = ksteps; cur += stride;
      return true;
    }
    if (cur >= end) return false;
    tile = cur / ksteps;
    k0 = cur - tile * ksteps;
    k1 = min(ksteps, k0 + (end - cur));
    cur += k1 - k0;
    return true;
  }
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// ---- thread-block cluster helpers (CTA-pair kernel)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// Remote arrive with the default (.release.cta) semantics, as CUTLASS's ClusterBarrier::arrive(cta_id): what the waiter consumes
// was produced through tcgen05 / TMA and is ordered by the tcgen05 fences, not by this arrive.  (`.release.cluster` compiles
// to MEMBAR.ALL.GPU + ERRBAR per arrive - measured: it alone held the pair kernel at 0.9 us per k-step.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 1D bulk copy global -> shared, completion on an mbarrier (size and both addresses multiples of 16 bytes)
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// elected-lane variants for issue loops that the whole warp runs converged (see umma_bf16_ts_elect)
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}\n" ::"r"(bar), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
      "@e cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n}\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
      "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n}\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d_elect(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
      "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n}\n" ::"r"(dst), "l"(src), "r"(bytes),
      "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=2 (SW128) [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// K-major, SWIZZLE_64B descriptor (64-byte rows: 32 bf16 of K per row; 8-row atoms of 512 B)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;        // SBO: 8 rows * 64 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
  return d;
}
// kind::f16 instruction descriptor with BF16 operands, F32 accumulate
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (lane = row, 2 bf16 of K per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-converged issue: ALL lanes of the issuing warp execute these with warp-uniform operands and one elected lane issues.
// (Inside `if (lane == 0)` the compiler cannot prove the operands uniform: every tcgen05 instruction then gets R2UR moves and an
// ELECT / BRA.U.ANY waterfall, and the 138-instruction k-step loop of the single issuing thread - not the tensor core, not the
// operand feed - set the pace of the BF16x3 engine at ~700 clocks per k-step.)
__device__ __forceinline__ void umma_bf16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_2cta_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta_elect(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
      "}\n" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
// uniform warp index (CUTLASS's canonical_warp_idx_sync): role branches on it are uniform control flow
__device__ __forceinline__ int warp_idx_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));  // first source -> upper half
  return r;
}
__device__ __forceinline__ float bf16_lo_to_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_to_f32(uint32_t packed) { return __uint_as_float(packed & 0xFFFF0000u); }

// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, K-major both
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// cta_group::2: one instruction drives the tensor cores of both CTAs of the pair (M = 256: 128 rows per CTA, each CTA holds its
// own A rows in tensor memory and HALF of the B tile's rows in shared memory at the same offset); issued by the leader CTA only
__device__ __forceinline__ void umma_bf16_ts_2cta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// SS form of the pair instruction: each CTA's A rows and its half of the B rows come from its own shared memory (same offsets)
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of the pair -> one arrive on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

enum { EPI_GENERIC = 0, EPI_BIAS = 1, EPI_BIAS_RES = 2, EPI_BIAS_RV_CS = 3, EPI_BIAS_RES_CS = 4, EPI_BIAS_GEGLU = 5, EPI_BIAS_CS = 6, EPI_BIAS_PAIR = 7, EPI_COUNT = 8 };

// barrier slots shared by both kernels (8 bytes each, from bar_base): tmem_full[2] | tmem_empty[2] | tmem_ptr | sk_flag
constexpr int TC_BAR_TFULL = 3 * TC_MAX_STAGES, TC_BAR_TEMPTY = 3 * TC_MAX_STAGES + 2, TC_BAR_TMEM_SLOT = 3 * TC_MAX_STAGES + 4,
              TC_BAR_SK_FLAG = 3 * TC_MAX_STAGES + 5;

// ===================== epilogue role (8 warps starting at warp EW0) =====================
// `ksteps` is the length of a tile's K loop in the units the work iterator counts (k-steps in conv_tc.cu, operand units in
// conv_nf.cu); the stream-K bookkeeping only needs it to be the same everywhere.
// EW0 = index of the first of the eight epilogue warps (a multiple of 2 so that warp & 3 walks the TMEM lane quarters).
// NE = number of epilogue warps (8, or 16 for the short-K variant of the decoupled kernel: warp w then handles a column QUARTER of
// its TMEM lane quarter; the staging tiles of the extra warps grow downwards into the operand budget, which that kernel leaves free)
template <int EPI, int EW0 = 2, int NE = TC_EPI_WARPS>
__device__ __forceinline__ void tc_epilogue_role(const TcParams& p, uint8_t* smem_raw, uint32_t smem_base, uint32_t bar_base,
                                                 uint32_t tmem_base, uint32_t acc_stride, int ksteps, int total_tiles) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  auto tfull_bar = [&](int a) { return bar_base + 8u * (TC_BAR_TFULL + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (TC_BAR_TEMPTY + a); };
  // Eight warps: warp w reads TMEM lane quarter (w & 3) and the column half ((w - 2) >> 2) of the accumulator, in
  // chunks of 16 columns.  tcgen05.ld gives thread = accumulator row; for NHWC outputs (o_sn == 1) each 32x16 chunk
  // is transposed through a 2 KB per-warp swizzled staging tile so that every global instruction covers 8 rows x
  // 64 contiguous bytes, and bias / timestep row / residual / activation are applied in that arrangement.
  // Transposed outputs (V^T: o_sp == 1) are already coalesced across lanes and go out directly.
  const int ew = warp - EW0;
  const int q = warp & 3;            // TMEM lane quarter this warp may access
  const int half = ew >> 2;            // column part of this warp: 0..NE/4-1
  const int row = q * 32 + lane;     // tile row owned by this thread (direct path)
  float4* stg = reinterpret_cast<float4*>(smem_raw + (smem_base - smem_u32(smem_raw)) + TC_SMEM_BUDGET - (NE - TC_EPI_WARPS) * 2048) + ew * 128;
  float* cacc = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + TC_SMEM_BUDGET + TC_STG_BYTES);
  const int et = threadIdx.x - 32 * EW0;   // 0..255 among the epilogue warps
  const int img_q = (q * 32) >> (p.lTW + p.lTH);  // image slot of this warp's rows inside the tile (all 32 rows share it)
  const int sub = lane >> 2, c4 = lane & 3;
  const int hcols = p.BN / (NE / 4);   // columns per warp (a multiple of 16 for every tile width the launchers pick)
  const int col_lo = half * hcols;
  constexpr bool GEN = EPI == EPI_GENERIC;
  const bool geglu = GEN ? (p.act == FRIDO_ACT_GEGLU || p.act == FRIDO_ACT_GEGLU_FAST) : (EPI == EPI_BIAS_GEGLU);
  const bool has_res = GEN ? (p.res != nullptr) : (EPI == EPI_BIAS_RES || EPI == EPI_BIAS_RES_CS);
  const bool has_rv = GEN ? (p.rowvec != nullptr) : (EPI == EPI_BIAS_RV_CS);
  const bool has_bias = p.bias != nullptr;
  const bool has_cs = GEN ? (p.csum != nullptr) : (EPI == EPI_BIAS_RV_CS || EPI == EPI_BIAS_RES_CS || EPI == EPI_BIAS_CS);
  const bool has_pair = GEN ? (p.out_hi != nullptr) : (EPI == EPI_BIAS_PAIR);
  const bool rnd = GEN ? (p.round_tf32 != 0) : false;
  const float alpha = GEN ? p.alpha : 1.0f;
  const int act = GEN ? p.act : (EPI == EPI_BIAS_GEGLU ? FRIDO_ACT_GEGLU_FAST : FRIDO_ACT_NONE);
  const bool nhwc = GEN ? (p.o_sn == 1) : true;
  int acc = 0;
  uint32_t acc_phase = 0;
  volatile uint32_t* sk_flag = reinterpret_cast<volatile uint32_t*>(smem_raw + (bar_base + 8u * TC_BAR_SK_FLAG - smem_u32(smem_raw)));
  SegIter it(p, ksteps, total_tiles);
  int tile, k0, k1;
  bool first_seg = true;
  while (it.next(tile, k0, k1)) {
    const int nt = tile % p.tiles_n;
    int mt = tile / p.tiles_n;
    const int tx = mt % p.tiles_x; mt /= p.tiles_x;
    const int ty = mt % p.tiles_y;
    const int tb = mt / p.tiles_y;
    const int n0 = nt * p.BN + col_lo;
    mbar_wait(tfull_bar(acc), acc_phase);
    tc_fence_after();
    const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_stride + (uint32_t)col_lo;
    // ---- stream-K: a segment that covers only part of the K loop parks its partial sums; the last contributor of the
    // tile to arrive adds them up in CTA order and runs the epilogue below from the workspace instead of TMEM
    bool from_ws = false;
    int c_first = 0, c_last = 0;
    if (k1 - k0 != ksteps) {
      float4* slot = p.sk_ws + (size_t)(2 * blockIdx.x + (first_seg ? 0 : 1)) * (size_t)(32 * p.BN);
      for (int c = 0; c < hcols; c += 16) {
        uint32_t r[16];
        tmem_ld16(t_base + c, r);
        const int ch = (col_lo + c) >> 4;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          slot[(ch * 4 + k) * 128 + row] = make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]),
                                                       __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));  // the accumulator is free again (stream-K is never a pair launch)
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      first_seg = false;
      c_first = (tile * ksteps) / p.sk_per;
      c_last = ((tile + 1) * ksteps - 1) / p.sk_per;
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NE) : "memory");
      if (et == 0) {
        const int old = atomicAdd(p.sk_cnt + tile, 1);
        const bool last = old == c_last - c_first;
        if (last) p.sk_cnt[tile] = 0;  // every contributor has arrived: leave the counter ready for the next launch
        *sk_flag = last ? 1u : 0u;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NE) : "memory");
      if (*sk_flag == 0u) continue;
      __threadfence();
      from_ws = true;
    }
    first_seg = false;
    auto fetch16 = [&](int c, uint32_t (&r)[16]) {
      if (!from_ws) { tmem_ld16(t_base + c, r); return; }
      float4 a[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int ch = (col_lo + c) >> 4;
      // two contributors' partials in flight at a time (L2 latency, not bandwidth, is what this costs); the sums are
      // still taken in CTA order
      for (int cb = c_first; cb <= c_last; cb += 2) {
        float4 v[2][4];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int cc = cb + g;
          if (cc <= c_last) {
            const int sl = 2 * cc + ((cc * p.sk_per) / ksteps == tile ? 0 : 1);
            const float4* src = p.sk_ws + (size_t)sl * (size_t)(32 * p.BN) + (ch * 4) * 128 + row;
#pragma unroll
            for (int k = 0; k < 4; ++k) v[g][k] = __ldcg(src + k * 128);
          }
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (cb + g <= c_last) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { a[k].x += v[g][k].x; a[k].y += v[g][k].y; a[k].z += v[g][k].z; a[k].w += v[g][k].w; }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        r[4 * k] = __float_as_uint(a[k].x); r[4 * k + 1] = __float_as_uint(a[k].y);
        r[4 * k + 2] = __float_as_uint(a[k].z); r[4 * k + 3] = __float_as_uint(a[k].w);
      }
    };
    if (nhwc) {
      // the 4 output rows this lane serves in the coalesced arrangement (rl = 8j + sub) are the same for every chunk
      long long obase[4];
      int bimg[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = q * 32 + 8 * j + sub;
        const int ox = tx * p.TW + (rr & (p.TW - 1));
        const int oy = ty * p.TH + ((rr >> p.lTW) & (p.TH - 1));
        const int b = tb * p.TB + (rr >> (p.lTW + p.lTH));
        bimg[j] = b;
        obase[j] = (ox < p.Wout && oy < p.Hout && b < p.B) ? (long long)b * p.o_sb + ((long long)oy * p.Wout + ox) * p.o_sp : -1;
      }
      if (has_cs) {
        for (int i = et; i < p.TB * p.BN * 2; i += 32 * NE) cacc[i] = 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NE) : "memory");
      }
      // the bias slice of a chunk is fetched one chunk ahead: an L2 round trip is longer than a whole chunk (ncu: the first
      // use of the bias was the epilogue's top stall)
      float4 bias_nx = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_bias) bias_nx = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 4 * c4));
      // the residual tile of a chunk is fetched ONE CHUNK AHEAD (like the bias): a DRAM / L2 round trip is longer than a whole
      // chunk of epilogue work, and with the loads issued at the top of their own chunk every chunk of a short-K GEMM's
      // epilogue stalled for a full memory latency
      float4 rres_nx[4];
      auto load_res = [&](int cc, float4 (&dst)[4]) {
        const int nn = n0 + cc + 4 * c4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (obase[j] >= 0) {
            if (geglu) { const float2 t2 = *reinterpret_cast<const float2*>(p.res + obase[j] + (nn >> 1)); dst[j].x = t2.x; dst[j].y = t2.y; }
            else dst[j] = *reinterpret_cast<const float4*>(p.res + obase[j] + nn);
          }
        }
      };
      if (has_res) load_res(0, rres_nx);
      for (int c = 0; c < hcols; c += 16) {
        const int n = n0 + c + 4 * c4;
        float4 rres[4];
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rres[j] = rres_nx[j];
          if (c + 16 < hcols) load_res(c + 16, rres_nx);
        }
        const float4 bias4 = bias_nx;
        if (has_bias && c + 16 < hcols) bias_nx = __ldg(reinterpret_cast<const float4*>(p.bias + n + 16));
        uint32_t r[16];
        fetch16(c, r);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          stg[lane * 4 + (k ^ ((lane >> 1) & 3))] = make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]),
                                                                __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
        __syncwarp();
        float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rl = 8 * j + sub;
          float4 v = stg[rl * 4 + (c4 ^ ((rl >> 1) & 3))];
          if (obase[j] >= 0) {
            v.x = fmaf(v.x, alpha, bias4.x); v.y = fmaf(v.y, alpha, bias4.y); v.z = fmaf(v.z, alpha, bias4.z); v.w = fmaf(v.w, alpha, bias4.w);
            if (has_rv) {
              const float4 e = __ldg(reinterpret_cast<const float4*>(p.rowvec + (long long)bimg[j] * p.rowvec_sb + n));
              v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
            }
            if (geglu) {
              float2 t = act == FRIDO_ACT_GEGLU ? make_float2(v.x * gelu_erf(v.y), v.z * gelu_erf(v.w))
                                                : make_float2(v.x * gelu_erf_fast(v.y), v.z * gelu_erf_fast(v.w));
              if (has_res) { t.x += rres[j].x; t.y += rres[j].y; }
              if (rnd) { t.x = round_tf32(t.x); t.y = round_tf32(t.y); }
              *reinterpret_cast<float2*>(p.out + obase[j] + (n >> 1)) = t;
            } else {
              if (has_res) { v.x += rres[j].x; v.y += rres[j].y; v.z += rres[j].z; v.w += rres[j].w; }
              if (act == FRIDO_ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              else if (act == FRIDO_ACT_SILU) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
              else if (act == FRIDO_ACT_GELU) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
              if (rnd) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
              const long long o = obase[j] + n;
              *reinterpret_cast<float4*>(p.out + o) = v;
              if (has_pair) {
                uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
                split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
                *reinterpret_cast<uint2*>(p.out_hi + o) = make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
                *reinterpret_cast<uint2*>(p.out_lo + o) = make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
              }
              if (has_cs) {
                cs[0] += v.x; cs[1] += v.y; cs[2] += v.z; cs[3] += v.w;
                cq[0] += v.x * v.x; cq[1] += v.y * v.y; cq[2] += v.z * v.z; cq[3] += v.w * v.w;
              }
            }
          }
        }
        if (has_cs) {  // reduce over the 8 row-subsets (lanes with equal c4), then one shared atomic per column
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 4); cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8); cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
            cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 4); cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 8); cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 16);
          }
          if (sub == 0) {
            float* a = cacc + ((img_q * p.BN) + col_lo + c + 4 * c4) * 2;
#pragma unroll
            for (int e = 0; e < 4; ++e) { atomicAdd(a + 2 * e, cs[e]); atomicAdd(a + 2 * e + 1, cq[e]); }
          }
        }
        __syncwarp();
      }
      if (has_cs) {
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NE) : "memory");
        for (int i = et; i < p.TB * p.BN; i += 32 * NE) {
          const int im = i / p.BN, col = i - im * p.BN;
          const int b = tb * p.TB + im;
          if (b < p.B) {
            double* d = p.csum + ((long long)b * p.Cout + nt * p.BN + col) * 2;
            atomicAdd(d, (double)cacc[2 * i]);
            atomicAdd(d + 1, (double)cacc[2 * i + 1]);
          }
        }
      }
    } else {
      const int ox = tx * p.TW + (row & (p.TW - 1));
      const int oy = ty * p.TH + ((row >> p.lTW) & (p.TH - 1));
      const int b = tb * p.TB + (row >> (p.lTW + p.lTH));
      const bool valid = ox < p.Wout && oy < p.Hout && b < p.B;
      const long long pix = (long long)oy * p.Wout + ox;
      const long long rowoff = (long long)b * p.o_sb + pix * p.o_sp;
      const float* __restrict__ rv = has_rv ? p.rowvec + (long long)b * p.rowvec_sb : nullptr;
      for (int c = 0; c < hcols; c += 16) {
        uint32_t r[16];
        fetch16(c, r);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c + j;
            float t = __uint_as_float(r[j]) * alpha;
            if (has_bias) t += __ldg(p.bias + n);
            if (rv) t += __ldg(rv + n);
            const long long o = rowoff + (long long)n * p.o_sn;
            if (has_res) t += p.res[o];
            if (act == FRIDO_ACT_RELU) t = fmaxf(t, 0.f);
            else if (act == FRIDO_ACT_SILU) t = silu_f(t);
            else if (act == FRIDO_ACT_GELU) t = gelu_erf(t);
            t = rnd ? round_tf32(t) : t;
            if (p.out) p.out[o] = t;
            if (has_pair) {
              uint16_t hh, ll;
              split_bf16(t, hh, ll);
              p.out_hi[o] = hh; p.out_lo[o] = ll;
            }
          }
        }
      }
    }
    if (!from_ws) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (p.pair) mbar_arrive_cluster(mapa_rank(tempty_bar(acc), 0));  // the leader CTA's MMA thread waits for both epilogues
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// Clocks per 32-wide k-step of the BF16x3 kernels as a function of the tile width (fit to tools/prof/conv_bench.py sweeps):
// FRIDO_TC_CLK="A,B" overrides intercept and slope-per-2-columns (tuning aid).
inline int bf_stage_clk(int n) {
  static int A = -1, B = 0;
  if (A < 0) {
    A = 450; B = 4;   // decoupled, uniform-issue kernels: ~520 clocks at BN = 64, ~810 at BN = 192 (was 256 + 2.5 n for the round-1 kernel)
    if (const char* e = getenv("FRIDO_TC_CLK")) { int a = 0, b = 0; if (sscanf(e, "%d,%d", &a, &b) == 2 && a > 0 && b > 0) { A = a; B = b; } }
  }
  return A + B * n / 2;
}

inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// rank-4 fp32 map (C, W, H, B) with a {32, bw, bh, bb} box, SWIZZLE_128B
inline bool make_map4(CUtensorMap* m, const float* base, uint64_t C, uint64_t W, uint64_t H, uint64_t Bn, int64_t sx, int64_t sy,
                      int64_t sb, uint32_t bw, uint32_t bh, uint32_t bb, uint32_t es_xy) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {C, W, H, Bn};
  // strides of dims 1..3 in bytes; size-1 dims get a harmless natural stride
  const int64_t s1 = sx, s2 = (H > 1 || sy) ? sy : sx * (int64_t)W, s3 = (Bn > 1 || sb) ? sb : (s2 ? s2 : sx * (int64_t)W) * (int64_t)H;
  cuuint64_t strides[3] = {(cuuint64_t)s1 * 4, (cuuint64_t)(s2 ? s2 : s1 * (int64_t)W) * 4, (cuuint64_t)(s3 ? s3 : s1 * (int64_t)W * (int64_t)H) * 4};
  // traversal stride es_xy (stride-2 convs): the box spans bw*es input columns and yields bw of them
  cuuint32_t box[4] = {TC_BK, bw * es_xy, bh * es_xy, bb};
  cuuint32_t es[4] = {1, es_xy, es_xy, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool make_map3(CUtensorMap* m, const void* base, uint64_t K, uint64_t N, uint64_t Bn, int64_t ld, int64_t sb, uint32_t bn,
                      bool bf16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  const int es_bytes = bf16 ? 2 : 4;
  cuuint64_t dims[3] = {K, N, Bn};
  cuuint64_t strides[2] = {(cuuint64_t)ld * es_bytes, (cuuint64_t)(sb ? sb : ld * (int64_t)N) * es_bytes};
  cuuint32_t box[3] = {TC_BK, bn, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides,
             box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace frido
