// tcgen05 TF32 implicit-GEMM engine (sm_100a): conv3x3 / conv1x1 / linear /
// batched matmul on the 5th-gen tensor cores.
//
//   D[128 x BN] (fp32, TMEM) += A[128 x 32] (fp32 read as TF32, smem) * W[BN x 32]^T
//
// * persistent CTAs (grid = min(tiles, #SM)), 192 threads:
//     warp 0   : TMA producer   (cp.async.bulk.tensor, 4-stage mbarrier ring)
//     warp 1   : TMEM allocator + single-thread tcgen05.mma issuer
//     warps 2-5: epilogue       (tcgen05.ld -> bias/emb/residual/act -> global)
// * A tile = one TMA box of an NHWC tensor: {32 channels, TW, TH, TB} pixels
//   (128 rows), loaded per filter tap at shifted coordinates; TMA zero-fills the
//   out-of-bounds halo, so the conv padding costs nothing.  Two A sources give
//   the channel concat (skip connections) for free.
// * W tile = TMA box {32 k, BN rows} of the K-major weight matrix (optionally
//   one matrix per image: attention QK^T / PV).
// * both tiles land in the canonical K-major SWIZZLE_128B layout, so the UMMA
//   shared-memory descriptors are (start, SBO=1024B, swizzle=128B) and K-steps
//   of 8 TF32 advance the start address by 32 bytes.
// * accumulators are double-buffered in TMEM (2 x 256 columns): the epilogue of
//   tile i overlaps the main loop of tile i+1.
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
constexpr int TC_SMEM_BYTES = TC_SMEM_BUDGET + TC_STG_BYTES + TC_CSUM_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;  // < 227 KB
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
  int stages;           // depth of the smem ring
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
};

// Work iterator shared by all warp roles: yields (tile, [k0, k1)) segments in the same order everywhere.
struct SegIter {
  int sk, ksteps, total_tiles, stride, cur, end;
  __device__ __forceinline__ SegIter(const TcParams& p, int ksteps_, int total_tiles_)
      : sk(p.sk), ksteps(ksteps_), total_tiles(total_tiles_), stride((int)gridDim.x) {
    if (sk) {
      cur = (int)blockIdx.x * p.sk_per;
      end = min(cur + p.sk_per, total_tiles_ * ksteps_);
    } else {
      cur = (int)blockIdx.x;
      end = 0;
    }
  }
  __device__ __forceinline__ bool next(int& tile, int& k0, int& k1) {
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

// ----------------------------------------------------------------------------
// MODE 0: single-pass TF32.
// MODE 1: error-compensated 3xTF32: every operand tile is split in shared memory into hi = rna_tf32(v) and
//         lo = v - hi by four extra warps, and each K-step issues hi*hi + lo*hi + hi*lo (products ~2^-21).
// MODE 2: error-compensated BF16x3 at twice the TF32 issue rate: the fp32 A tile (TMA, shared memory) is split by four
//         warps into bf16 hi/lo halves that go straight into TENSOR MEMORY (tcgen05.st) and feed the MMA as its TMEM
//         A operand; the weights arrive PRE-SPLIT from HBM as two bf16 matrices (two TMA maps, SWIZZLE_64B);
//         hi*hi + lo*hi + hi*lo with fp32 accumulation (products ~2^-16).  Shared-memory bandwidth (128 B/clk/SM) is
//         what bounds this kernel: per 32-wide K step it carries the TMA writes (16 KB A + 128*BN B of W), the
//         splitter's read of A (16 KB) and the tensor core's reads of W (3 MMAs x 2 x 32*BN B); keeping the split A
//         tiles out of shared memory removes 16 KB of writes and 32 KB of MMA operand reads per step.
// Epilogue specialisations: compile-time feature sets for the dense NHWC store path (alpha = 1, no TF32 rounding, no bf16
// pair copy), so the per-chunk loop carries no dead branches.  Measured with ncu on a K=384 GEMM: the generic epilogue
// executes ~300 instructions per 32x16 chunk at ~12 clocks each (instruction-cache misses and branch resolution on the
// uniform feature tests) and, at 22 k clocks per tile, outlasts the 12 k-step main loop it is supposed to hide behind.
enum { EPI_GENERIC = 0, EPI_BIAS = 1, EPI_BIAS_RES = 2, EPI_BIAS_RV_CS = 3, EPI_BIAS_RES_CS = 4, EPI_BIAS_GEGLU = 5, EPI_BIAS_CS = 6, EPI_BIAS_PAIR = 7, EPI_COUNT = 8 };

template <int MODE, int EPI>
__global__ void __launch_bounds__(MODE ? TC_THREADS_X3 : TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_wlo,
               const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1, const TcParams p) {
  constexpr bool X3 = MODE == 1;
  constexpr bool BF = MODE == 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  // stage layout  MODE 0: A | W      MODE 1: A_hi | A_lo | W_hi | W_lo      MODE 2: A_raw(fp32) | W_hi | W_lo (bf16)
  const uint32_t b_bytes = BF ? (uint32_t)p.BN * TC_BK * 2 : (uint32_t)p.BN * TC_BK * 4;
  const uint32_t stage_bytes = BF ? (TC_A_BYTES + 2u * b_bytes) : (TC_A_BYTES + b_bytes) * (X3 ? 2u : 1u);
  const uint32_t off_alo = (uint32_t)TC_A_BYTES;
  const uint32_t off_w = X3 ? 2u * TC_A_BYTES : (uint32_t)TC_A_BYTES;
  const uint32_t off_wlo = off_w + b_bytes;
  const uint32_t acc_stride = BF ? TC_BF_ACC_STRIDE : TC_MAX_BN;
  const uint32_t bar_base = smem_base + TC_SMEM_BUDGET + TC_STG_BYTES + TC_CSUM_BYTES;
  // barrier layout: full[6] | empty[6] | split[6] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
  auto split_bar = [&](int s) { return bar_base + 8u * (2 * TC_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * TC_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * TC_MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * TC_MAX_STAGES + 4);
  const int NS = p.stages;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  const int Cin = p.c0 + p.c1;
  const int kchunks = Cin / TC_BK;
  const int taps = p.ksize * p.ksize;
  const int ksteps_main = taps * kchunks;
  const int ksteps = ksteps_main + (p.cx0 + p.cx1) / TC_BK;  // + the 1x1 side input's channels
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const int total_tiles = m_tiles * p.tiles_n;
  const uint32_t stage_tx = TC_A_BYTES + b_bytes * (BF ? 2u : 1u);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a0);
    if (p.c1) prefetch_tmap(&map_a1);
    if (p.cx0) prefetch_tmap(&map_x0);
    if (p.cx1) prefetch_tmap(&map_x1);
    prefetch_tmap(&map_w);
    if (BF) prefetch_tmap(&map_wlo);
    for (int s = 0; s < NS; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(split_bar(s), TC_SPLIT_WARPS);  // one arrive per splitter warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), TC_EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();  // set-up above overlapped the predecessor's tail; from here on we touch its outputs

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      SegIter it(p, ksteps, total_tiles);
      int tile, k0, k1;
      while (it.next(tile, k0, k1)) {
        const int nt = tile % p.tiles_n;
        int mt = tile / p.tiles_n;
        const int tx = mt % p.tiles_x; mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tb = mt / p.tiles_y;
        const int ox0 = tx * p.TW, oy0 = ty * p.TH, b0 = tb * p.TB;
        const int n0 = nt * p.BN;
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + off_w;
          mbar_expect_tx(full_bar(stage), stage_tx);
          if (ks < ksteps_main) {
            const int tap = ks / kchunks;
            const int kc = ks - tap * kchunks;
            const int dy = tap / p.ksize, dx = tap - dy * p.ksize;
            const int ch = kc * TC_BK;
            const int cx = ox0 * p.stride + dx - p.pad, cy = oy0 * p.stride + dy - p.pad;
            if (ch < p.c0) tma_load_4d(sa, &map_a0, full_bar(stage), ch, cx, cy, b0);
            else           tma_load_4d(sa, &map_a1, full_bar(stage), ch - p.c0, cx, cy, b0);
          } else {  // side input: the pixel itself (a 1x1 tap), channels of x0 then x1
            const int ch = (ks - ksteps_main) * TC_BK;
            if (ch < p.cx0) tma_load_4d(sa, &map_x0, full_bar(stage), ch, ox0 * p.stride, oy0 * p.stride, b0);
            else            tma_load_4d(sa, &map_x1, full_bar(stage), ch - p.cx0, ox0 * p.stride, oy0 * p.stride, b0);
          }
          // weight columns run [tap][channel] then the side input's channels: k-step ks starts at column 32 * ks
          tma_load_3d(sb, &map_w, full_bar(stage), ks * TC_BK, n0, p.w_batched ? b0 : 0);
          if (BF) tma_load_3d(sa + off_wlo, &map_wlo, full_bar(stage), ks * TC_BK, n0, p.w_batched ? b0 : 0);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = BF ? umma_idesc_bf16(TC_BM, p.BN) : umma_idesc_tf32(TC_BM, p.BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      SegIter it(p, ksteps, total_tiles);
      int tile, k0, k1;
      while (it.next(tile, k0, k1)) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(MODE ? split_bar(stage) : full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + off_w;
          if (BF) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {  // 2 K-steps of 16 bf16: 8 TMEM columns of A, 32 B inside the 64 B swizzle row of W
              const uint32_t ah = tmem_base + (uint32_t)(TC_BF_A_COL + stage * 32 + k * 8), al = ah + 16;
              const uint64_t bh = umma_desc_sw64(sb + k * 32), bl = umma_desc_sw64(sa + off_wlo + k * 32);
              umma_bf16_ts(d_tmem, ah, bh, idesc, ((ks - k0) | k) ? 1u : 0u);
              umma_bf16_ts(d_tmem, al, bh, idesc, 1u);
              umma_bf16_ts(d_tmem, ah, bl, idesc, 1u);
            }
          } else
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t ad = umma_desc_sw128(sa + k * 32);
            const uint64_t bd = umma_desc_sw128(sb + k * 32);
            umma_tf32(d_tmem, ad, bd, idesc, ((ks - k0) | k) ? 1u : 0u);
            if (X3) {
              const uint64_t al = umma_desc_sw128(sa + off_alo + k * 32);
              const uint64_t bl = umma_desc_sw128(sa + off_wlo + k * 32);
              umma_tf32(d_tmem, al, bd, idesc, 1u);
              umma_tf32(d_tmem, ad, bl, idesc, 1u);
            }
          }
          umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));  // accumulator ready for the epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 2 + TC_EPI_WARPS) {
    // ===================== epilogue (warps 2..9) =====================
    // Eight warps: warp w reads TMEM lane quarter (w & 3) and the column half ((w - 2) >> 2) of the accumulator, in
    // chunks of 16 columns.  tcgen05.ld gives thread = accumulator row; for NHWC outputs (o_sn == 1) each 32x16 chunk
    // is transposed through a 2 KB per-warp swizzled staging tile so that every global instruction covers 8 rows x
    // 64 contiguous bytes, and bias / timestep row / residual / activation are applied in that arrangement.
    // Transposed outputs (V^T: o_sp == 1) are already coalesced across lanes and go out directly.
    const int ew = warp - 2;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = ew >> 2;
    const int row = q * 32 + lane;     // tile row owned by this thread (direct path)
    float4* stg = reinterpret_cast<float4*>(smem_raw + (smem_base - smem_u32(smem_raw)) + TC_SMEM_BUDGET) + ew * 128;
    float* cacc = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + TC_SMEM_BUDGET + TC_STG_BYTES);
    const int et = threadIdx.x - 64;   // 0..255 among the epilogue warps
    const int img_q = (q * 32) >> (p.lTW + p.lTH);  // image slot of this warp's rows inside the tile (all 32 rows share it)
    const int sub = lane >> 2, c4 = lane & 3;
    const int hcols = p.BN >> 1;
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
    volatile uint32_t* sk_flag = reinterpret_cast<volatile uint32_t*>(smem_raw + (bar_base + 8u * 23 - smem_u32(smem_raw)));
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
        if (lane == 0) mbar_arrive(tempty_bar(acc));  // the accumulator is free again
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        first_seg = false;
        c_first = (tile * ksteps) / p.sk_per;
        c_last = ((tile + 1) * ksteps - 1) / p.sk_per;
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          const int old = atomicAdd(p.sk_cnt + tile, 1);
          const bool last = old == c_last - c_first;
          if (last) p.sk_cnt[tile] = 0;  // every contributor has arrived: leave the counter ready for the next launch
          *sk_flag = last ? 1u : 0u;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
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
          for (int i = et; i < p.TB * p.BN * 2; i += 32 * TC_EPI_WARPS) cacc[i] = 0.f;
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        // the bias slice of a chunk is fetched one chunk ahead: an L2 round trip is longer than a whole chunk (ncu: the first
        // use of the bias was the epilogue's top stall)
        float4 bias_nx = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_bias) bias_nx = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 4 * c4));
        for (int c = 0; c < hcols; c += 16) {
          const int n = n0 + c + 4 * c4;
          // residual loads of this chunk go out first: their latency overlaps the TMEM load + transpose
          float4 rres[4];
          if (has_res) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              rres[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (obase[j] >= 0) {
                if (geglu) { const float2 t2 = *reinterpret_cast<const float2*>(p.res + obase[j] + (n >> 1)); rres[j].x = t2.x; rres[j].y = t2.y; }
                else rres[j] = *reinterpret_cast<const float4*>(p.res + obase[j] + n);
              }
            }
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
          asm volatile("bar.sync 1, 256;" ::: "memory");
          for (int i = et; i < p.TB * p.BN; i += 32 * TC_EPI_WARPS) {
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
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (BF) {
    // ===================== splitter (warps 10..13), BF16x3: fp32 A tile (smem) -> bf16 hi / lo halves in TMEM ==========
    // source: 128 rows x 128 B, SWIZZLE_128B (16-B chunk c of row r sits at chunk c ^ (r & 7)).  Thread = row (the warp
    // owns TMEM lanes 32*(warp & 3) ..+31); destination columns: hi k0..31 -> 16 columns, lo -> the next 16.
    const int r = (warp & 3) * 32 + lane;
    const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)TC_BF_A_COL;
    int stage = 0;
    uint32_t phase = 0;
    SegIter it(p, ksteps, total_tiles);
    int tile, k0, k1;
    while (it.next(tile, k0, k1)) {
      for (int ks = k0; ks < k1; ++ks) {
        mbar_wait(full_bar(stage), phase);
        const uint8_t* srow = smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)stage * stage_bytes + r * 128;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(srow + ((c ^ (r & 7)) << 4));
          const uint32_t h0 = pack_bf16x2(v.x, v.y), h1 = pack_bf16x2(v.z, v.w);
          hi[2 * c] = h0; hi[2 * c + 1] = h1;
          lo[2 * c] = pack_bf16x2(v.x - bf16_lo_to_f32(h0), v.y - bf16_hi_to_f32(h0));
          lo[2 * c + 1] = pack_bf16x2(v.z - bf16_lo_to_f32(h1), v.w - bf16_hi_to_f32(h1));
        }
        tc_fence_after();  // the MMAs that last read this TMEM slot retired before the TMA refilled the stage
        tmem_st16(a_lane + (uint32_t)(stage * 32), hi);
        tmem_st16(a_lane + (uint32_t)(stage * 32 + 16), lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(stage));
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (X3) {
    // ===================== splitter (warps 6..9, 3xTF32 only) =====================
    // In place: v -> hi = rna_tf32(v); lo = v - hi goes to the twin buffer at the same (swizzled) offset.
    const int t = threadIdx.x - TC_THREADS;
    int stage = 0;
    uint32_t phase = 0;
    const int a_vec = TC_A_BYTES / 16, w_vec = (int)(b_bytes / 16);
    SegIter it(p, ksteps, total_tiles);
    int tile, k0, k1;
    while (it.next(tile, k0, k1)) {
      for (int ks = k0; ks < k1; ++ks) {
        mbar_wait(full_bar(stage), phase);
        uint8_t* sbase = smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)stage * stage_bytes;
        float4* a_hi = reinterpret_cast<float4*>(sbase);
        float4* a_lo = reinterpret_cast<float4*>(sbase + off_alo);
        float4* w_hi = reinterpret_cast<float4*>(sbase + off_w);
        float4* w_lo = reinterpret_cast<float4*>(sbase + off_wlo);
#pragma unroll 4
        for (int i = t; i < a_vec; i += TC_SPLIT_THREADS) {
          const float4 v = a_hi[i];
          const float4 h = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
          a_hi[i] = h;
          a_lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
#pragma unroll 4
        for (int i = t; i < w_vec; i += TC_SPLIT_THREADS) {
          const float4 v = w_hi[i];
          const float4 h = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
          w_hi[i] = h;
          w_lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(stage));
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
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

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// rank-4 fp32 map (C, W, H, B) with a {32, bw, bh, bb} box, SWIZZLE_128B
static bool make_map4(CUtensorMap* m, const float* base, uint64_t C, uint64_t W, uint64_t H, uint64_t Bn, int64_t sx, int64_t sy,
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

static bool make_map3(CUtensorMap* m, const void* base, uint64_t K, uint64_t N, uint64_t Bn, int64_t ld, int64_t sb, uint32_t bn,
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

static bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int conv2d_tc(const FridoConvParams* p, cudaStream_t s) {
  if (!p->a0 || !p->w || (!p->out && !(p->out_hi && p->o_sn != 1))) return set_error(FRIDO_E_ARG, "conv2d_tc: null pointer");
  if (p->ups != 1) return set_error(FRIDO_E_ARG, "conv2d_tc: ups must be 1 (materialise the upsample first)");
  if (p->out_u8) return set_error(FRIDO_E_ARG, "conv2d_tc: out_u8 is a feature of the small-Cout head kernels (engine 0)");
  if (p->stride != 1 && !(p->stride == 2 && p->ksize == 3)) return set_error(FRIDO_E_ARG, "conv2d_tc: stride must be 1, or 2 for 3x3");
  if (p->ksize != 1 && p->ksize != 3) return set_error(FRIDO_E_ARG, "conv2d_tc: ksize must be 1 or 3");
  // pad = ksize/2, or 0 for the encoder's asymmetric stride-2 conv (taming model.py:68-72: pad right/bottom by 1 = TMA zero fill)
  if (p->pad != p->ksize / 2 && !(p->pad == 0 && p->stride == 2 && p->ksize == 3))
    return set_error(FRIDO_E_ARG, "conv2d_tc: pad must be ksize/2 (or 0 for 3x3 stride 2)");
  if (p->c0 % TC_BK || p->c1 % TC_BK || p->c0 <= 0) return set_error(FRIDO_E_ARG, "conv2d_tc: channels must be multiples of 32");
  if ((p->c1 > 0) != (p->a1 != nullptr)) return set_error(FRIDO_E_ARG, "conv2d_tc: a1/c1 mismatch");
  if (p->Cout % 64) return set_error(FRIDO_E_ARG, "conv2d_tc: Cout must be a multiple of 64");
  if (p->a0_sc != 1 || (p->a1 && p->a1_sc != 1)) return set_error(FRIDO_E_ARG, "conv2d_tc: channel stride must be 1");
  if (p->pad == 0 && p->ksize == 3) {
    if (p->Hout != (p->Hin - 2) / 2 + 1 || p->Wout != (p->Win - 2) / 2 + 1) return set_error(FRIDO_E_ARG, "conv2d_tc: bad output size");
  } else if (p->Hout != (p->Hin + p->stride - 1) / p->stride || p->Wout != (p->Win + p->stride - 1) / p->stride)
    return set_error(FRIDO_E_ARG, "conv2d_tc: output size must be ceil(in/stride)");
  if (!a16(p->a0) || !a16(p->w) || (p->a1 && !a16(p->a1)) || (p->out && !a16(p->out)) || (p->res && !a16(p->res)))
    return set_error(FRIDO_E_ARG, "conv2d_tc: pointers must be 16-byte aligned");
  if (p->a0_sx % 4 || p->a0_sy % 4 || p->a0_sb % 4 || (p->a1 && (p->a1_sx % 4 || p->a1_sy % 4 || p->a1_sb % 4)))
    return set_error(FRIDO_E_ARG, "conv2d_tc: strides must be multiples of 16 bytes");
  const int Cin = p->c0 + p->c1;
  if (p->cx0 < 0 || p->cx1 < 0 || p->cx0 % TC_BK || p->cx1 % TC_BK || (p->cx0 > 0) != (p->x0 != nullptr) || (p->cx1 > 0) != (p->x1 != nullptr) ||
      (p->cx1 > 0 && p->cx0 == 0))
    return set_error(FRIDO_E_ARG, "conv2d_tc: side input channels must be multiples of 32 and match x0/x1");
  if (p->cx0 && (p->stride != 1 || p->w_sb || !a16(p->x0) || (p->x1 && !a16(p->x1)) || p->x0_sx % 4 || p->x0_sy % 4 || p->x0_sb % 4 ||
                 (p->x1 && (p->x1_sx % 4 || p->x1_sy % 4 || p->x1_sb % 4))))
    return set_error(FRIDO_E_ARG, "conv2d_tc: side input needs stride 1, shared weights and 16-byte aligned strides");
  const int64_t Ktot = (int64_t)p->ksize * p->ksize * Cin + p->cx0 + p->cx1;
  const int64_t w_ld = p->w_ld ? p->w_ld : Ktot;
  const bool bf = p->engine == 3;
  if (bf && !p->w_lo) return set_error(FRIDO_E_ARG, "conv2d_tc: engine 3 (bf16x3) needs pre-split weights (w = bf16 hi, w_lo = bf16 lo)");
  if (bf && (w_ld % 8 || p->w_sb % 8 || !a16(p->w_lo)))
    return set_error(FRIDO_E_ARG, "conv2d_tc: bf16 weight strides must be multiples of 16 bytes");
  if (w_ld % 4 || p->w_sb % 4) return set_error(FRIDO_E_ARG, "conv2d_tc: weight strides must be multiples of 16 bytes");
  if ((p->act == FRIDO_ACT_GEGLU || p->act == FRIDO_ACT_GEGLU_FAST) && p->o_sn != 1) return set_error(FRIDO_E_ARG, "conv2d_tc: GEGLU needs a dense output");
  if (p->o_sn == 1 && (p->o_sp % 4 || p->o_sb % 4)) return set_error(FRIDO_E_ARG, "conv2d_tc: output rows must be 16-byte aligned");
  if (p->o_sn == 1 && ((p->bias && !a16(p->bias)) || (p->rowvec && (!a16(p->rowvec) || p->rowvec_sb % 4))))
    return set_error(FRIDO_E_ARG, "conv2d_tc: bias / rowvec must be 16-byte aligned");

  TcParams t;
  t.B = p->B; t.Hout = p->Hout; t.Wout = p->Wout; t.Cout = p->Cout;
  t.c0 = p->c0; t.c1 = p->c1; t.cx0 = p->cx0; t.cx1 = p->cx1; t.ksize = p->ksize; t.pad = p->pad; t.stride = p->stride;
  t.TW = next_pow2(p->Wout) < TC_BM ? next_pow2(p->Wout) : TC_BM;
  t.TH = next_pow2(p->Hout) < TC_BM / t.TW ? next_pow2(p->Hout) : TC_BM / t.TW;
  t.TB = TC_BM / (t.TW * t.TH);
  t.lTW = 0; while ((1 << t.lTW) < t.TW) ++t.lTW;
  t.lTH = 0; while ((1 << t.lTH) < t.TH) ++t.lTH;
  t.tiles_x = (p->Wout + t.TW - 1) / t.TW;
  t.tiles_y = (p->Hout + t.TH - 1) / t.TH;
  t.tiles_b = (p->B + t.TB - 1) / t.TB;
  if (p->w_sb && t.TB != 1) return set_error(FRIDO_E_ARG, "conv2d_tc: per-image weights need >= 128 rows per image");
  const int m_tiles = t.tiles_x * t.tiles_y * t.tiles_b;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // Tile width: minimise waves x per-stage time.  Per stage the tensor pipe needs (#MMA per stage) * BN/2 cycles and the
  // operand splitter / TMA issue put a floor under it, so narrow tiles only pay when they remove whole waves.
  const int mode = p->engine == 3 ? 2 : (p->engine == 2 ? 1 : 0);
  const int mma_per_stage[3] = {4, 12, 6};
  const int floor_clk[3] = {260, 1000, 450};
  int bn = 64;
  double best_cost = 1e30;
  const int cands[4] = {256, 192, 128, 64};
  for (int i = 0; i < 4; ++i) {
    if (p->Cout % cands[i]) continue;
    if (mode == 2 && cands[i] > TC_BF_ACC_STRIDE) continue;  // BF16x3 keeps the A operand ring in TMEM next to 2 x 192 accumulator columns
    const int64_t tiles = (int64_t)m_tiles * (p->Cout / cands[i]);
    const int64_t waves = (tiles + sms - 1) / sms;
    int clk = mma_per_stage[mode] * cands[i] / 2;
    if (clk < floor_clk[mode]) clk = floor_clk[mode];
    // BF16x3 is bound by shared-memory bandwidth (128 B/clk): TMA writes 16 KB + 128*BN, splitter reads 16 KB, MMAs read 192*BN
    if (mode == 2) clk = 256 + 5 * cands[i] / 2;
    const double cost = (double)waves * clk;
    if (cost < best_cost) { best_cost = cost; bn = cands[i]; }
  }
  if (const char* f = getenv("FRIDO_TC_FORCE_BN")) {  // profiling aid only (tools/prof/conv_bench.py)
    const int v = atoi(f);
    if (v >= 32 && v <= (mode == 2 ? TC_BF_ACC_STRIDE : TC_MAX_BN) && v % 32 == 0 && p->Cout % v == 0) bn = v;
  }
  // Stream-K for launches that cannot fill the machine with whole tiles (8x8 / 16x16 levels, long K): compare the
  // data-parallel schedule chosen above with an even split of all (tile, k-step) iterations over the SMs, in clocks.
  t.sk = 0; t.sk_per = 0; t.sk_ws = nullptr; t.sk_cnt = nullptr;
  int sk_grid = 0;
  {
    const int ksteps = p->ksize * p->ksize * (Cin / TC_BK) + (p->cx0 + p->cx1) / TC_BK;
    const char* sk_e = getenv("FRIDO_SK");  // 0 = off, 1 = cost model (default), 2 = whenever legal (tests)
    const int sk_env = sk_e ? atoi(sk_e) : 1;
    auto stage_clk = [&](int n) {
      int clk = mma_per_stage[mode] * n / 2;
      if (clk < floor_clk[mode]) clk = floor_clk[mode];
      if (mode == 2) clk = 256 + 5 * n / 2;
      return clk;
    };
    const int64_t dp_tiles = (int64_t)m_tiles * (p->Cout / bn);
    const double dp_cost = (double)((dp_tiles + sms - 1) / sms) * ksteps * stage_clk(bn);
    double thresh = 0.95;
    if (const char* e = getenv("FRIDO_SK_THRESH")) thresh = atof(e);  // tuning aid
    double best = sk_env == 2 ? 1e30 : thresh * dp_cost;
    const bool forced_bn = getenv("FRIDO_TC_FORCE_BN") != nullptr;
    if (sk_env && p->sk_ws && (reinterpret_cast<uintptr_t>(p->sk_ws) & 15) == 0 && ksteps >= 8) {
      for (int i = 0; i < 4; ++i) {
        const int n = cands[i];
        if (p->Cout % n || (mode == 2 && n > TC_BF_ACC_STRIDE) || (forced_bn && n != bn)) continue;
        const int64_t tiles = (int64_t)m_tiles * (p->Cout / n);
        const int64_t iters = tiles * ksteps;
        if (tiles > 1024 || iters > (1 << 28)) continue;
        int64_t g = iters / 4 < sms ? iters / 4 : sms;   // at least 4 k-steps per CTA
        if (g < 1) g = 1;
        int64_t per = (iters + g - 1) / g;
        const int64_t min_per = (ksteps + 5) / 6;         // at most 7 contributors per tile
        if (per < min_per) per = min_per;
        g = (iters + per - 1) / per;
        if (per % ksteps == 0 && sk_env != 2) continue;   // whole tiles only: that is the data-parallel schedule
        const int64_t need = 4096 + 2 * g * 128 * n * 4;
        if (need > p->sk_ws_bytes) continue;
        const int contrib = (int)((ksteps + per - 1) / per) + 1;
        const double cost = (double)per * stage_clk(n) + 3000.0 + 8.0 * n * (1 + contrib);
        if (cost < best) { best = cost; bn = n; t.sk = 1; t.sk_per = (int)per; sk_grid = (int)g; }
      }
    }
    if (t.sk) {
      t.sk_cnt = reinterpret_cast<int*>(p->sk_ws);
      t.sk_ws = reinterpret_cast<float4*>(reinterpret_cast<char*>(p->sk_ws) + 4096);
    }
  }
  t.BN = bn;
  t.tiles_n = p->Cout / bn;
  t.w_batched = p->w_sb != 0;
  t.bias = p->bias; t.rowvec = p->rowvec; t.rowvec_sb = p->rowvec_sb; t.res = p->res;
  t.alpha = p->alpha; t.act = p->act; t.out = p->out; t.o_sb = p->o_sb; t.o_sp = p->o_sp; t.o_sn = p->o_sn;
  t.round_tf32 = p->round_tf32;
  t.csum = p->chan_sums;
  t.out_hi = (uint16_t*)p->out_hi; t.out_lo = (uint16_t*)p->out_lo;
  if ((p->out_hi != nullptr) != (p->out_lo != nullptr) || (p->out_hi && (p->act == FRIDO_ACT_GEGLU || p->act == FRIDO_ACT_GEGLU_FAST)))
    return set_error(FRIDO_E_ARG, "conv2d_tc: out_hi/out_lo must come together and not with GEGLU");
  if (p->chan_sums && (p->o_sn != 1 || (p->act == FRIDO_ACT_GEGLU || p->act == FRIDO_ACT_GEGLU_FAST) || t.TW * t.TH < 32 || t.TB > 4))
    return set_error(FRIDO_E_ARG, "conv2d_tc: chan_sums needs a dense NHWC output, no GEGLU and >= 32 pixels per image");

  CUtensorMap ma0, ma1, mw, mwlo, mx0, mx1;
  if (!make_map4(&ma0, p->a0, p->c0, p->Win, p->Hin, p->B, p->a0_sx, p->a0_sy, p->a0_sb, t.TW, t.TH, t.TB, (uint32_t)p->stride))
    return set_error(FRIDO_E_ARG, "conv2d_tc: cuTensorMapEncodeTiled(a0) failed");
  if (p->a1) {
    if (!make_map4(&ma1, p->a1, p->c1, p->Win, p->Hin, p->B, p->a1_sx, p->a1_sy, p->a1_sb, t.TW, t.TH, t.TB, (uint32_t)p->stride))
      return set_error(FRIDO_E_ARG, "conv2d_tc: cuTensorMapEncodeTiled(a1) failed");
  } else {
    ma1 = ma0;
  }
  mx0 = ma0; mx1 = ma0;
  if (p->x0 && !make_map4(&mx0, p->x0, p->cx0, p->Wout, p->Hout, p->B, p->x0_sx, p->x0_sy, p->x0_sb, t.TW, t.TH, t.TB, 1u))
    return set_error(FRIDO_E_ARG, "conv2d_tc: cuTensorMapEncodeTiled(x0) failed");
  if (p->x1 && !make_map4(&mx1, p->x1, p->cx1, p->Wout, p->Hout, p->B, p->x1_sx, p->x1_sy, p->x1_sb, t.TW, t.TH, t.TB, 1u))
    return set_error(FRIDO_E_ARG, "conv2d_tc: cuTensorMapEncodeTiled(x1) failed");
  if (!make_map3(&mw, p->w, Ktot, p->Cout, p->w_sb ? p->B : 1, w_ld, p->w_sb, bn, bf))
    return set_error(FRIDO_E_ARG, "conv2d_tc: cuTensorMapEncodeTiled(w) failed");
  if (bf) {
    if (!make_map3(&mwlo, p->w_lo, Ktot, p->Cout, p->w_sb ? p->B : 1, w_ld, p->w_sb, bn, true))
      return set_error(FRIDO_E_ARG, "conv2d_tc: cuTensorMapEncodeTiled(w_lo) failed");
  } else {
    mwlo = mw;
  }

  using KernelFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, TcParams);
  static const KernelFn bf_kernels[EPI_COUNT] = {conv_tc_kernel<2, EPI_GENERIC>, conv_tc_kernel<2, EPI_BIAS>, conv_tc_kernel<2, EPI_BIAS_RES>,
                                                conv_tc_kernel<2, EPI_BIAS_RV_CS>, conv_tc_kernel<2, EPI_BIAS_RES_CS>,
                                                conv_tc_kernel<2, EPI_BIAS_GEGLU>, conv_tc_kernel<2, EPI_BIAS_CS>,
                                                conv_tc_kernel<2, EPI_BIAS_PAIR>};
  static bool attr = false;
  if (!attr) {
    bool ok = cudaFuncSetAttribute(conv_tc_kernel<0, EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) == cudaSuccess &&
              cudaFuncSetAttribute(conv_tc_kernel<1, EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) == cudaSuccess;
    for (int i = 0; i < EPI_COUNT && ok; ++i)
      ok = cudaFuncSetAttribute(bf_kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) == cudaSuccess;
    if (!ok) return set_error(FRIDO_E_LAUNCH, "conv2d_tc: cannot opt in to dynamic shared memory");
    attr = true;
  }
  // epilogue variant (BF16x3 only): the feature set of this launch, if one of the specialised kernels covers it
  int epi = EPI_GENERIC;
  {
    const char* e = getenv("FRIDO_EPI_SPEC");  // 0 = always the generic epilogue (A/B aid)
    const bool spec = !e || atoi(e) != 0;
    if (bf && spec && p->o_sn == 1 && p->alpha == 1.0f && !p->round_tf32 && p->out_hi) {
      if (p->out && p->act == FRIDO_ACT_NONE && !p->res && !p->rowvec && !p->chan_sums) epi = EPI_BIAS_PAIR;
    } else if (bf && spec && p->o_sn == 1 && p->alpha == 1.0f && !p->round_tf32) {
      const bool res = p->res != nullptr, rv = p->rowvec != nullptr, cs = p->chan_sums != nullptr;
      if (p->act == FRIDO_ACT_NONE) {
        if (!res && !rv && !cs) epi = EPI_BIAS;
        else if (res && !rv && !cs) epi = EPI_BIAS_RES;
        else if (!res && rv && cs) epi = EPI_BIAS_RV_CS;
        else if (res && !rv && cs) epi = EPI_BIAS_RES_CS;
        else if (!res && !rv && cs) epi = EPI_BIAS_CS;
      } else if (p->act == FRIDO_ACT_GEGLU_FAST && !res && !rv && !cs) {
        epi = EPI_BIAS_GEGLU;
      }
    }
  }
  const bool x3 = p->engine == 2;
  const int stage_bytes = bf ? (TC_A_BYTES + 2 * bn * TC_BK * 2) : (TC_A_BYTES + bn * TC_BK * 4) * (x3 ? 2 : 1);
  t.stages = TC_SMEM_BUDGET / stage_bytes;
  if (t.stages > TC_MAX_STAGES) t.stages = TC_MAX_STAGES;
  if (bf && t.stages > TC_BF_MAX_STAGES) t.stages = TC_BF_MAX_STAGES;
  const int total = m_tiles * t.tiles_n;
  const int grid = t.sk ? sk_grid : (total < sms ? total : sms);
  if (bf) launch_pdl(bf_kernels[epi], dim3(grid), dim3(TC_THREADS_X3), TC_SMEM_BYTES, s, ma0, ma1, mw, mwlo, mx0, mx1, t);
  else if (x3) launch_pdl(conv_tc_kernel<1, EPI_GENERIC>, dim3(grid), dim3(TC_THREADS_X3), TC_SMEM_BYTES, s, ma0, ma1, mw, mwlo, mx0, mx1, t);
  else launch_pdl(conv_tc_kernel<0, EPI_GENERIC>, dim3(grid), dim3(TC_THREADS), TC_SMEM_BYTES, s, ma0, ma1, mw, mwlo, mx0, mx1, t);
  return check_launch(bf ? "conv2d_tc(bf16x3)" : x3 ? "conv2d_tc(3xTF32)" : "conv2d_tc");
}

}  // namespace frido
