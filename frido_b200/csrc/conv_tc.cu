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
#include "tc_common.cuh"

#ifndef FRIDO_TC_EPI16_DEFAULT
#define FRIDO_TC_EPI16_DEFAULT 0
#endif
#ifndef FRIDO_TC_PAIR_DEFAULT
#define FRIDO_TC_PAIR_DEFAULT 0
#endif

namespace frido {

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

// MODE 2 runs one more warp (warp 14): the W tiles get their own TMA issuer.  Measured (tools/prof/conv_bench.py, FRIDO_TC_DBG /
// FRIDO_TC_STAGES): the k-step of this engine was 725 + 1.0 x BN clocks - a fixed cost that did not move with the MMA width, the W
// bytes or the ring depth beyond 4, and dropped 23 % without the A-tile TMA: ONE thread issuing three TMA instructions (plus the
// tap / channel coordinate arithmetic) per k-step set the pace of the whole pipeline.
constexpr int TC_THREADS_BF = TC_THREADS_X3 + 32;

template <int MODE, int EPI>
__global__ void __launch_bounds__(MODE == 2 ? TC_THREADS_BF : (MODE ? TC_THREADS_X3 : TC_THREADS), 1)
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
      mbar_init(full_bar(s), BF ? 2 : 1);       // BF16x3: the A and the W issuer each arrive (with their byte counts)
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
          mbar_expect_tx(full_bar(stage), BF ? ((p.dbg & 1) ? 0u : (uint32_t)TC_A_BYTES) : stage_tx);
          if (BF && (p.dbg & 1)) {
            // timing experiment: the A tile is not fetched at all
          } else if (ks < ksteps_main) {
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
          if (!BF) tma_load_3d(sb, &map_w, full_bar(stage), ks * TC_BK, n0, p.w_batched ? b0 : 0);   // BF16x3: warp 14 issues W
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
    tc_epilogue_role<EPI>(p, smem_raw, smem_base, bar_base, tmem_base, acc_stride, ksteps, total_tiles);
  } else if (BF && warp == 14) {
    // ===================== W-tile TMA issuer (BF16x3): hi and lo tiles of every k-step =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      SegIter it(p, ksteps, total_tiles);
      int tile, k0, k1;
      while (it.next(tile, k0, k1)) {
        const int n0 = (tile % p.tiles_n) * p.BN;
        const int wb = p.w_batched ? (tile / p.tiles_n / (p.tiles_x * p.tiles_y)) * p.TB : 0;
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * stage_bytes;
          mbar_expect_tx(full_bar(stage), 2u * b_bytes);
          tma_load_3d(sa + off_w, &map_w, full_bar(stage), ks * TC_BK, n0, wb);
          tma_load_3d(sa + off_wlo, &map_wlo, full_bar(stage), ks * TC_BK, n0, wb);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
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
        if (p.dbg & 4) {
#pragma unroll
          for (int c = 0; c < 16; ++c) { hi[c] = (uint32_t)(ks + c); lo[c] = (uint32_t)r; }
        } else if (p.dbg & 2) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(srow + ((c ^ (r & 7)) << 4));
            hi[2 * c] = v.x; hi[2 * c + 1] = v.y; lo[2 * c] = v.z; lo[2 * c + 1] = v.w;
          }
        } else
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

// ----------------------------------------------------------------------------
// BF16x3 with DECOUPLED operand rings (default for engine 3).  The stage loop of the kernel above - MMA retires -> slot free ->
// TMA issued -> tile lands -> split -> MMA - is ~1.9 us long, and with the four stages that the tensor-memory A slots allow a
// k-step could not go below a quarter of that (measured: 2 / 3 / 4 stages = 600 / 397 / 336 us for the same conv; MMA-bound
// would be 0.31 us per k-step).  Here every resource has its own ring and is released by its LAST reader:
//   raw A tiles   a_stages (6-8) x 16 KB   filled by the A issuer (warp 0), released by the splitter as soon as it has read them
//   W tiles       stages (3-4) x 2 tiles   filled by the W issuer (warp 14), released by the MMAs' commit
//   split A       4 tensor-memory slots    filled by the splitter, released by the MMAs' commit
// so the long TMA latency of the A tile is covered by six to eight tiles in flight instead of four.
// ----------------------------------------------------------------------------
constexpr int TCB_MAX_A = 8, TCB_MAX_W = 6;
// barrier slots (8 bytes each): a_full[8] 0.. | a_empty[8] 8.. | (18..23: accumulator / stream-K slots of the epilogue role)
//                               | w_full[6] 24.. | w_empty[6] 30.. | split[4] 36.. | t_empty[4] 40..
// NE = 8: warps 0 A issuer | 1 MMA | 2-9 epilogue | 10-13 splitter | 14 W issuer (480 threads)
// NE = 16 (short-K launches, whose pace the epilogue sets): 0 A issuer | 1 MMA | 2 W issuer | 3 idle | 4-19 epilogue | 20-23
//          splitter (768 threads, 80 registers each; 16 KB of the operand budget become staging tiles of the extra warps)
template <int EPI, int NE>
__global__ void __launch_bounds__(NE == 16 ? 768 : TC_THREADS_BF, 1)
conv_tc_bf_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_wlo,
                  const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)p.BN * TC_BK * 2;
  const uint32_t w_stage_bytes = 2u * b_bytes;
  const int NA = p.a_stages, NW = p.stages;
  const uint32_t w_ring = smem_base + (uint32_t)NA * TC_A_BYTES;
  const uint32_t bar_base = smem_base + TC_SMEM_BUDGET + TC_STG_BYTES + TC_CSUM_BYTES;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (8 + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (24 + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (30 + s); };
  auto split_bar = [&](int s) { return bar_base + 8u * (36 + s); };
  auto t_empty = [&](int s) { return bar_base + 8u * (40 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (TC_BAR_TFULL + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (TC_BAR_TEMPTY + a); };
  const uint32_t tmem_slot = bar_base + 8u * TC_BAR_TMEM_SLOT;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  // warp index as a warp-UNIFORM value (shuffle from lane 0): the role branches below are then uniform control flow and the
  // single-thread issue loops (TMA, MMA) compile to the uniform datapath instead of per-thread registers + R2UR moves
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  constexpr int EW0 = NE == 16 ? 4 : 2, W_ISSUER = NE == 16 ? 2 : 14;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  const int Cin = p.c0 + p.c1;
  const int kchunks = Cin / TC_BK;
  const int taps = p.ksize * p.ksize;
  const int ksteps_main = taps * kchunks;
  const int ksteps = ksteps_main + (p.cx0 + p.cx1) / TC_BK;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const int total_tiles = m_tiles * p.tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a0);
    if (p.c1) prefetch_tmap(&map_a1);
    if (p.cx0) prefetch_tmap(&map_x0);
    if (p.cx1) prefetch_tmap(&map_x1);
    prefetch_tmap(&map_w);
    prefetch_tmap(&map_wlo);
    for (int s = 0; s < TCB_MAX_A; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), TC_SPLIT_WARPS); }
    for (int s = 0; s < TCB_MAX_W; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < TC_BF_MAX_STAGES; ++s) { mbar_init(split_bar(s), TC_SPLIT_WARPS); mbar_init(t_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), NE); }
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
  pdl_wait();

  if (warp == 0) {
    // ===================== A-tile TMA issuer (whole warp converged, elected lane issues) =====================
    {
      int sa = 0;
      uint32_t pha = 0;
      SegIter it(p, ksteps, total_tiles);
      int tile, k0, k1;
      while (it.next(tile, k0, k1)) {
        int mt = tile / p.tiles_n;
        const int tx = mt % p.tiles_x; mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tb = mt / p.tiles_y;
        const int ox0 = tx * p.TW, oy0 = ty * p.TH, b0 = tb * p.TB;
        // (tap, channel chunk) of k-step ks advance incrementally: no divisions in the loop
        int tap = k0 < ksteps_main ? k0 / kchunks : 0, kc = k0 < ksteps_main ? k0 - tap * kchunks : 0;
        int dy = tap / p.ksize, dx = tap - dy * p.ksize;
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(a_empty(sa), pha ^ 1);
          const uint32_t dst = smem_base + (uint32_t)sa * TC_A_BYTES;
          mbar_expect_tx_elect(a_full(sa), (p.dbg & 1) ? 0u : (uint32_t)TC_A_BYTES);
          if (p.dbg & 1) {
            // timing experiment: no A tile
          } else if (ks < ksteps_main) {
            const int ch = kc * TC_BK;
            const int cx = ox0 * p.stride + dx - p.pad, cy = oy0 * p.stride + dy - p.pad;
            if (ch < p.c0) tma_load_4d_elect(dst, &map_a0, a_full(sa), ch, cx, cy, b0);
            else           tma_load_4d_elect(dst, &map_a1, a_full(sa), ch - p.c0, cx, cy, b0);
            if (++kc == kchunks) { kc = 0; ++tap; if (++dx == p.ksize) { dx = 0; ++dy; } }
          } else {  // side input: the pixel itself (a 1x1 tap), channels of x0 then x1
            const int ch = (ks - ksteps_main) * TC_BK;
            if (ch < p.cx0) tma_load_4d_elect(dst, &map_x0, a_full(sa), ch, ox0 * p.stride, oy0 * p.stride, b0);
            else            tma_load_4d_elect(dst, &map_x1, a_full(sa), ch - p.cx0, ox0 * p.stride, oy0 * p.stride, b0);
          }
          if (++sa == NA) { sa = 0; pha ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp runs the loop converged, one elected lane issues =====================
    {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, p.BN);
      // shared-memory descriptor of a W tile = constant high word | (address >> 4) in the low 14 bits
      const uint64_t desc_hi = umma_desc_sw64(0);
      const uint32_t w_ring4 = w_ring >> 4, wst4 = w_stage_bytes >> 4, bb4 = b_bytes >> 4;
      int sw = 0, ts = 0, acc = 0;
      uint32_t phw = 0, pht = 0, acc_phase = 0;
      SegIter it(p, ksteps, total_tiles);
      int tile, k0, k1;
      while (it.next(tile, k0, k1)) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * TC_BF_ACC_STRIDE;
        uint32_t accum = 0;
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(w_full(sw), phw);
          mbar_wait(split_bar(ts), pht);
          tc_fence_after();
          const uint32_t a0 = tmem_base + (uint32_t)(TC_BF_A_COL + ts * 32);
          const uint64_t b0 = desc_hi | (uint64_t)(w_ring4 + (uint32_t)sw * wst4);   // (all W addresses stay below 256 KB: no carry)
          umma_bf16_ts_elect(d_tmem, a0, b0, idesc, accum);               // k16 = 0: hi hi, lo hi, hi lo
          umma_bf16_ts_elect(d_tmem, a0 + 16, b0, idesc, 1u);
          umma_bf16_ts_elect(d_tmem, a0, b0 + bb4, idesc, 1u);
          umma_bf16_ts_elect(d_tmem, a0 + 8, b0 + 2, idesc, 1u);          // k16 = 1: +8 columns of A, +32 bytes inside the W rows
          umma_bf16_ts_elect(d_tmem, a0 + 24, b0 + 2, idesc, 1u);
          umma_bf16_ts_elect(d_tmem, a0 + 8, b0 + bb4 + 2, idesc, 1u);
          accum = 1u;
          umma_commit_elect(w_empty(sw));   // frees the W stage ...
          umma_commit_elect(t_empty(ts));   // ... and the tensor-memory operand slot when these MMAs retire
          if (++sw == NW) { sw = 0; phw ^= 1; }
          if (++ts == TC_BF_MAX_STAGES) { ts = 0; pht ^= 1; }
        }
        umma_commit_elect(tfull_bar(acc));  // accumulator ready for the epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= EW0 && warp < EW0 + NE) {
    tc_epilogue_role<EPI, EW0, NE>(p, smem_raw, smem_base, bar_base, tmem_base, (uint32_t)TC_BF_ACC_STRIDE, ksteps, total_tiles);
  } else if (warp == W_ISSUER) {
    // ===================== W-tile TMA issuer: hi and lo tiles of every k-step (converged warp, elected issue) ==========
    {
      int sw = 0;
      uint32_t phw = 0;
      SegIter it(p, ksteps, total_tiles);
      int tile, k0, k1;
      while (it.next(tile, k0, k1)) {
        const int n0 = (tile % p.tiles_n) * p.BN;
        const int wb = p.w_batched ? (tile / p.tiles_n / (p.tiles_x * p.tiles_y)) * p.TB : 0;
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(w_empty(sw), phw ^ 1);
          const uint32_t sb = w_ring + (uint32_t)sw * w_stage_bytes;
          mbar_expect_tx_elect(w_full(sw), w_stage_bytes);
          tma_load_3d_elect(sb, &map_w, w_full(sw), ks * TC_BK, n0, wb);
          tma_load_3d_elect(sb + b_bytes, &map_wlo, w_full(sw), ks * TC_BK, n0, wb);
          if (++sw == NW) { sw = 0; phw ^= 1; }
        }
      }
    }
  } else if (warp >= EW0 + NE) {
    // ===================== splitter (4 warps): fp32 A tile (shared memory) -> bf16 hi / lo halves in tensor memory ==========
    const int r = (warp & 3) * 32 + lane;
    const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)TC_BF_A_COL;
    int sa = 0, ts = 0;
    uint32_t pha = 0, pht = 0;
    SegIter it(p, ksteps, total_tiles);
    int tile, k0, k1;
    while (it.next(tile, k0, k1)) {
      for (int ks = k0; ks < k1; ++ks) {
        mbar_wait(a_full(sa), pha);
        const uint8_t* srow = smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)sa * TC_A_BYTES + r * 128;
        uint32_t hi[16], lo[16];
        if (p.dbg & 4) {
#pragma unroll
          for (int c = 0; c < 16; ++c) { hi[c] = (uint32_t)(ks + c); lo[c] = (uint32_t)r; }
        } else
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(srow + ((c ^ (r & 7)) << 4));
          const uint32_t h0 = pack_bf16x2(v.x, v.y), h1 = pack_bf16x2(v.z, v.w);
          hi[2 * c] = h0; hi[2 * c + 1] = h1;
          lo[2 * c] = pack_bf16x2(v.x - bf16_lo_to_f32(h0), v.y - bf16_hi_to_f32(h0));
          lo[2 * c + 1] = pack_bf16x2(v.z - bf16_lo_to_f32(h1), v.w - bf16_hi_to_f32(h1));
        }
        // the raw tile is in registers: hand the shared-memory slot back to the A issuer right away
        __syncwarp();
        if (lane == 0) mbar_arrive(a_empty(sa));
        mbar_wait(t_empty(ts), pht ^ 1);  // the MMAs that last read this tensor-memory slot have retired
        tc_fence_after();
        tmem_st16(a_lane + (uint32_t)(ts * 32), hi);
        tmem_st16(a_lane + (uint32_t)(ts * 32 + 16), lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(ts));
        if (++sa == NA) { sa = 0; pha ^= 1; }
        if (++ts == TC_BF_MAX_STAGES) { ts = 0; pht ^= 1; }
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


int conv2d_tc_pair_launch(const TcParams& t, int epi, int clusters, const CUtensorMap& ma0, const CUtensorMap& ma1,
                          const CUtensorMap& mw, const CUtensorMap& mwlo, const CUtensorMap& mx0, const CUtensorMap& mx1,
                          cudaStream_t s);

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
    if (mode == 2) clk = bf_stage_clk(cands[i]);
    const double cost = (double)waves * clk;
    if (cost < best_cost) { best_cost = cost; bn = cands[i]; }
  }
  if (const char* f = getenv("FRIDO_TC_FORCE_BN")) {  // profiling aid only (tools/prof/conv_bench.py)
    const int v = atoi(f);
    if (v >= 32 && v <= (mode == 2 ? TC_BF_ACC_STRIDE : TC_MAX_BN) && v % 32 == 0 && p->Cout % v == 0) bn = v;
  }
  // Stream-K for launches that cannot fill the machine with whole tiles (8x8 / 16x16 levels, long K): compare the
  // data-parallel schedule chosen above with an even split of all (tile, k-step) iterations over the SMs, in clocks.
  t.sk = 0; t.sk_per = 0; t.sk_ws = nullptr; t.sk_cnt = nullptr; t.pair = 0; t.dbg_w = nullptr; t.dbg = 0; t.a_stages = 0;
  if (const char* e = getenv("FRIDO_TC_DBG")) t.dbg = atoi(e);
  int sk_grid = 0;
  {
    const int ksteps = p->ksize * p->ksize * (Cin / TC_BK) + (p->cx0 + p->cx1) / TC_BK;
    const char* sk_e = getenv("FRIDO_SK");  // 0 = off, 1 = cost model (default), 2 = whenever legal (tests)
    const int sk_env = sk_e ? atoi(sk_e) : 1;
    auto stage_clk = [&](int n) {
      int clk = mma_per_stage[mode] * n / 2;
      if (clk < floor_clk[mode]) clk = floor_clk[mode];
      if (mode == 2) clk = bf_stage_clk(n);
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
  static const KernelFn bfd_kernels[2 * EPI_COUNT] = {
      conv_tc_bf_kernel<EPI_GENERIC, 8>, conv_tc_bf_kernel<EPI_BIAS, 8>, conv_tc_bf_kernel<EPI_BIAS_RES, 8>, conv_tc_bf_kernel<EPI_BIAS_RV_CS, 8>,
      conv_tc_bf_kernel<EPI_BIAS_RES_CS, 8>, conv_tc_bf_kernel<EPI_BIAS_GEGLU, 8>, conv_tc_bf_kernel<EPI_BIAS_CS, 8>, conv_tc_bf_kernel<EPI_BIAS_PAIR, 8>,
      conv_tc_bf_kernel<EPI_GENERIC, 16>, conv_tc_bf_kernel<EPI_BIAS, 16>, conv_tc_bf_kernel<EPI_BIAS_RES, 16>, conv_tc_bf_kernel<EPI_BIAS_RV_CS, 16>,
      conv_tc_bf_kernel<EPI_BIAS_RES_CS, 16>, conv_tc_bf_kernel<EPI_BIAS_GEGLU, 16>, conv_tc_bf_kernel<EPI_BIAS_CS, 16>, conv_tc_bf_kernel<EPI_BIAS_PAIR, 16>};
  static DevOnce attr;
  if (attr.need()) {
    bool ok = cudaFuncSetAttribute(conv_tc_kernel<0, EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) == cudaSuccess &&
              cudaFuncSetAttribute(conv_tc_kernel<1, EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) == cudaSuccess;
    for (int i = 0; i < EPI_COUNT && ok; ++i)
      ok = cudaFuncSetAttribute(bf_kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) == cudaSuccess &&
           cudaFuncSetAttribute(bfd_kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) == cudaSuccess &&
           cudaFuncSetAttribute(bfd_kernels[EPI_COUNT + i], cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) == cudaSuccess;
    if (!ok) return set_error(FRIDO_E_LAUNCH, "conv2d_tc: cannot opt in to dynamic shared memory");
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
  // CTA-pair schedule (conv_tc2.cu, cta_group::2): two M tiles of one N tile per cluster, each CTA holds half of the W tile.
  // FRIDO_TC_PAIR: 0 = never, 1 = launches with at least one full wave of pairs (default), 2 = whenever legal (tests)
  {
    const char* e = getenv("FRIDO_TC_PAIR");
    const int pair_env = e ? atoi(e) : FRIDO_TC_PAIR_DEFAULT;
    const long long items = (long long)(m_tiles / 2) * t.tiles_n;
    if (bf && pair_env && !t.sk && !p->w_sb && m_tiles % 2 == 0 && bn % 32 == 0 && (pair_env == 2 || items >= sms / 2)) {
      t.pair = 1;
      if (getenv("FRIDO_TC_DBG_BULKW")) t.dbg_w = p->w;  // timing experiment (wrong results)
      CUtensorMap mwh, mwl;
      if (!make_map3(&mwh, p->w, Ktot, p->Cout, 1, w_ld, 0, bn / 2, true) || !make_map3(&mwl, p->w_lo, Ktot, p->Cout, 1, w_ld, 0, bn / 2, true))
        return set_error(FRIDO_E_ARG, "conv2d_tc: cuTensorMapEncodeTiled(w, pair) failed");
      const int sb = TC_A_BYTES + 2 * (bn / 2) * TC_BK * 2;
      t.stages = TC_SMEM_BUDGET / sb;   // shared-memory ring; the tensor-memory operand ring behind it has TC_BF_MAX_STAGES slots
      if (t.stages > TC_MAX_STAGES) t.stages = TC_MAX_STAGES;
      const int clusters = (int)(items < sms / 2 ? items : sms / 2);
      return conv2d_tc_pair_launch(t, epi, clusters, ma0, ma1, mwh, mwl, mx0, mx1, s);
    }
  }
  const bool x3 = p->engine == 2;
  const int stage_bytes = bf ? (TC_A_BYTES + 2 * bn * TC_BK * 2) : (TC_A_BYTES + bn * TC_BK * 4) * (x3 ? 2 : 1);
  t.stages = TC_SMEM_BUDGET / stage_bytes;
  if (t.stages > TC_MAX_STAGES) t.stages = TC_MAX_STAGES;
  if (bf && t.stages > TC_BF_MAX_STAGES) t.stages = TC_BF_MAX_STAGES;
  if (const char* e = getenv("FRIDO_TC_STAGES")) { const int v = atoi(e); if (v >= 2 && v < t.stages) t.stages = v; }  // profiling aid
  const int total = m_tiles * t.tiles_n;
  const int grid = t.sk ? sk_grid : (total < sms ? total : sms);
  // decoupled operand rings (default; FRIDO_TC_DECOUPLE=0 = the stage-coupled kernel): W ring of 3-4 stages, raw A ring of
  // whatever is left of the operand budget (6 tiles at BN = 192, 8 below)
  bool decouple = bf;
  if (const char* e = getenv("FRIDO_TC_DECOUPLE")) decouple = decouple && atoi(e) != 0;
  if (decouple) {
    // 16 epilogue warps for short K loops (the epilogue of a 128 x BN tile outlasts a main loop of fewer than ~48 k-steps);
    // FRIDO_TC_EPI16 = 0 never | 1 by that rule (default) | 2 always
    const int ksteps_all = p->ksize * p->ksize * (Cin / TC_BK) + (p->cx0 + p->cx1) / TC_BK;
    int e16 = FRIDO_TC_EPI16_DEFAULT;
    if (const char* e = getenv("FRIDO_TC_EPI16")) e16 = atoi(e);
    const bool wide_epi = e16 == 2 || (e16 == 1 && ksteps_all <= 48);
    const int budget = TC_SMEM_BUDGET - (wide_epi ? 8 * 2048 : 0);
    const int wsb = 2 * bn * TC_BK * 2;
    int nw = 4;
    int na = (budget - nw * wsb) / TC_A_BYTES;
    if (na < 4) { nw = 3; na = (budget - nw * wsb) / TC_A_BYTES; }
    if (na > TCB_MAX_A) na = TCB_MAX_A;
    if (const char* e = getenv("FRIDO_TC_STAGES")) { const int v = atoi(e); if (v >= 2 && v < na) na = v; }  // profiling aid
    t.stages = nw; t.a_stages = na;
    launch_pdl(bfd_kernels[(wide_epi ? EPI_COUNT : 0) + epi], dim3(grid), dim3(wide_epi ? 768 : TC_THREADS_BF), TC_SMEM_BYTES, s, ma0, ma1, mw, mwlo,
               mx0, mx1, t);
    return check_launch("conv2d_tc(bf16x3)");
  }
  if (bf) launch_pdl(bf_kernels[epi], dim3(grid), dim3(TC_THREADS_BF), TC_SMEM_BYTES, s, ma0, ma1, mw, mwlo, mx0, mx1, t);
  else if (x3) launch_pdl(conv_tc_kernel<1, EPI_GENERIC>, dim3(grid), dim3(TC_THREADS_X3), TC_SMEM_BYTES, s, ma0, ma1, mw, mwlo, mx0, mx1, t);
  else launch_pdl(conv_tc_kernel<0, EPI_GENERIC>, dim3(grid), dim3(TC_THREADS), TC_SMEM_BYTES, s, ma0, ma1, mw, mwlo, mx0, mx1, t);
  return check_launch(bf ? "conv2d_tc(bf16x3)" : x3 ? "conv2d_tc(3xTF32)" : "conv2d_tc");
}

}  // namespace frido
