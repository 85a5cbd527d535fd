// tcgen05 BF16x3 implicit-GEMM engine with the INPUT NORMALISATION APPLIED ON LOAD ("normalise-on-load", sm_100a):
//
//   out = conv3x3 / conv1x1 ( act( GroupNorm(x) [* (1 + gamma_map) + beta_map] ) )  [+ fused 1x1 side input, bias, ...]
//
// i.e. a ResBlock's `in_layers` / `out_layers` (pyunet.py:209-240: normalization -> SiLU -> conv, SPADE modulation
// spade_norm.py:44-60) and a SpatialTransformer's `norm -> proj_in` (attention.py:254-262,296-298) as ONE launch: the
// normalised activation never exists in HBM (the separate norm_act pass was 14 % of a UNet step and ~4.8 GB of HBM traffic).
//
// Operand path (differs from conv_tc.cu, whose A tile is re-fetched per filter tap):
//   * the A operand of an output tile (TW x TH x TB = 128 pixels) is HALO-RESIDENT: per 32-channel chunk ONE TMA box
//     {32 ch, TW+2, TH+2, TB} lands in shared memory (TMA zero-fills outside the image), [for SPADE two more boxes with the
//     gamma / beta maps of the same pixels], and the four operand warps
//       pass 1: apply y = x*a[b,c] + b[b,c]  (a = rstd*gamma_c, b = beta_c - mean*rstd*gamma_c from the producer's channel
//               sums, frido_gn_finalize), the SPADE modulation and SiLU, force the halo outside the image back to exactly 0
//               (conv padding pads the ACTIVATED tensor), split into bf16 hi / lo and write the pair back in place;
//       pass 2: for each of the 9 taps copy the shifted 128 rows (thread = output pixel) into a tensor-memory slot
//               (tcgen05.st) from where the TS-form MMAs take them.
//     The normalisation, the activation and the operand split run ONCE per element instead of once per tap, and the
//     L2 -> shared-memory traffic of the A operand drops ~6x.
//   * weights: as conv_tc.cu (pre-split bf16 hi / lo, TMA, SWIZZLE_64B), K order = chunk-major: column (tap*C + chunk*32).
//   * extra K units after the chunks: the fused 1x1 side input (ResBlock skip_connection) reads RAW activations at the
//     output pixel, exactly as in conv_tc.cu.
// Warp roles (480 threads): 0 = weight TMA producer, 1 = TMEM allocator + MMA issuer, 2-9 = epilogue (shared with
// conv_tc.cu), 10-13 = operand warps, 14 = halo TMA producer.
#include "tc_common.cuh"

namespace frido {

constexpr int NF_THREADS = TC_THREADS_X3 + 32;
constexpr int NF_TSLOTS = TC_BF_MAX_STAGES;     // tensor-memory operand slots (32 columns each)
constexpr int NF_AB_BYTES = 2048;               // [2][TB <= 4][32][2] fp32 scale/shift table of the current chunk
constexpr int NF_MAX_HALO_ROWS = 208;

struct NfParams {
  int n_chunks;     // 32-channel chunks of the normalised input
  int taps;         // 9 (3x3) or 1 (1x1)
  int hpad;         // 1 or 0
  int HW2, HH2;     // halo box: TW + 2*hpad, TH + 2*hpad
  int rows_h;       // HW2 * HH2 * TB
  int side_units;   // (cx0 + cx1) / 32
  int slot_bytes;   // one halo tile in shared memory (1 KB multiple)
  int w_stages;     // depth of the weight ring
  int has_gb;       // SPADE maps present
  int Cn;           // c0 + c1
  int silu;
  int Hin, Win;
  const float* ab;  // [B][Cn][2]
};

__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

template <int EPI>
__global__ void __launch_bounds__(NF_THREADS, 1)
conv_nf_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_gb, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_wlo, const __grid_constant__ CUtensorMap map_x0,
               const __grid_constant__ CUtensorMap map_x1, const TcParams p, const NfParams q) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // operand region (TC_SMEM_BUDGET): x slot 0 | x slot 1 | [gamma | beta] | weight ring ... | ab table (last 2 KB)
  const uint32_t xs_off = 0;
  const uint32_t gb_off = 2u * q.slot_bytes;
  const uint32_t w_off = gb_off + (q.has_gb ? 2u * q.slot_bytes : 0u);
  const uint32_t b_bytes = (uint32_t)p.BN * TC_BK * 2;
  const uint32_t w_stage_bytes = 2u * b_bytes;
  const uint32_t ab_off = TC_SMEM_BUDGET - NF_AB_BYTES;
  const uint32_t bar_base = smem_base + TC_SMEM_BUDGET + TC_STG_BYTES + TC_CSUM_BYTES;
  // barrier slots: w_full[6] 0.. | w_empty[6] 6.. | t_full[4] 12.. | a_full[2] 16,17 | (18..23 shared with the epilogue role)
  //                | t_empty[4] 24.. | a_empty[2] 28,29 | gb_full 30 | gb_empty 31
  auto w_full = [&](int s) { return bar_base + 8u * s; };
  auto w_empty = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
  auto t_full = [&](int s) { return bar_base + 8u * (12 + s); };
  auto a_full = [&](int s) { return bar_base + 8u * (16 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (TC_BAR_TFULL + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (TC_BAR_TEMPTY + a); };
  auto t_empty = [&](int s) { return bar_base + 8u * (24 + s); };
  auto a_empty = [&](int s) { return bar_base + 8u * (28 + s); };
  const uint32_t gb_full = bar_base + 8u * 30, gb_empty = bar_base + 8u * 31;
  const uint32_t tmem_slot = bar_base + 8u * TC_BAR_TMEM_SLOT;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  const int n_units = q.n_chunks + q.side_units;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const int total_tiles = m_tiles * p.tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a0);
    if (p.c1) prefetch_tmap(&map_a1);
    if (q.has_gb) prefetch_tmap(&map_gb);
    if (p.cx0) prefetch_tmap(&map_x0);
    if (p.cx1) prefetch_tmap(&map_x1);
    prefetch_tmap(&map_w);
    prefetch_tmap(&map_wlo);
    for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < NF_TSLOTS; ++s) { mbar_init(t_full(s), TC_SPLIT_WARPS); mbar_init(t_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), TC_SPLIT_WARPS); }
    mbar_init(gb_full, 1);
    mbar_init(gb_empty, TC_SPLIT_WARPS);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), TC_EPI_WARPS); }
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
    // ===================== weight TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      SegIter it(p, n_units, total_tiles);
      int tile, u0, u1;
      while (it.next(tile, u0, u1)) {
        const int n0 = (tile % p.tiles_n) * p.BN;
        for (int u = u0; u < u1; ++u) {
          const bool nf = u < q.n_chunks;
          const int nk = nf ? q.taps : 1;
          for (int t = 0; t < nk; ++t) {
            // weight columns run [tap][channel] then the side input's channels
            const int col = nf ? t * q.Cn + u * TC_BK : q.taps * q.Cn + (u - q.n_chunks) * TC_BK;
            mbar_wait(w_empty(stage), phase ^ 1);
            const uint32_t sb = smem_base + w_off + stage * w_stage_bytes;
            mbar_expect_tx(w_full(stage), w_stage_bytes);
            tma_load_3d(sb, &map_w, w_full(stage), col, n0, 0);
            tma_load_3d(sb + b_bytes, &map_wlo, w_full(stage), col, n0, 0);
            if (++stage == q.w_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 14) {
    // ===================== halo TMA producer =====================
    if (lane == 0) {
      int aslot = 0;
      uint32_t aphase = 0, gphase = 0;
      SegIter it(p, n_units, total_tiles);
      int tile, u0, u1;
      while (it.next(tile, u0, u1)) {
        int mt = tile / p.tiles_n;
        const int tx = mt % p.tiles_x; mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tb = mt / p.tiles_y;
        const int ox0 = tx * p.TW, oy0 = ty * p.TH, b0 = tb * p.TB;
        for (int u = u0; u < u1; ++u) {
          mbar_wait(a_empty(aslot), aphase ^ 1);
          const uint32_t dst = smem_base + xs_off + aslot * q.slot_bytes;
          if (u < q.n_chunks) {
            const int ch = u * TC_BK;
            if (q.has_gb) {
              mbar_wait(gb_empty, gphase ^ 1);
              gphase ^= 1;
              mbar_expect_tx(gb_full, 2u * q.rows_h * 128u);
              tma_load_4d(smem_base + gb_off, &map_gb, gb_full, ch, ox0 - q.hpad, oy0 - q.hpad, b0);
              tma_load_4d(smem_base + gb_off + q.slot_bytes, &map_gb, gb_full, q.Cn + ch, ox0 - q.hpad, oy0 - q.hpad, b0);
            }
            mbar_expect_tx(a_full(aslot), (uint32_t)q.rows_h * 128u);
            if (ch < p.c0) tma_load_4d(dst, &map_a0, a_full(aslot), ch, ox0 - q.hpad, oy0 - q.hpad, b0);
            else           tma_load_4d(dst, &map_a1, a_full(aslot), ch - p.c0, ox0 - q.hpad, oy0 - q.hpad, b0);
          } else {  // side input: the output pixel itself, raw
            const int ch = (u - q.n_chunks) * TC_BK;
            mbar_expect_tx(a_full(aslot), (uint32_t)TC_A_BYTES);
            if (ch < p.cx0) tma_load_4d(dst, &map_x0, a_full(aslot), ch, ox0, oy0, b0);
            else            tma_load_4d(dst, &map_x1, a_full(aslot), ch - p.cx0, ox0, oy0, b0);
          }
          if (++aslot == 2) { aslot = 0; aphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, p.BN);
      int stage = 0, tslot = 0, acc = 0;
      uint32_t phase = 0, tphase = 0, acc_phase = 0;
      SegIter it(p, n_units, total_tiles);
      int tile, u0, u1;
      while (it.next(tile, u0, u1)) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * TC_BF_ACC_STRIDE;
        bool first = true;
        for (int u = u0; u < u1; ++u) {
          const int nk = u < q.n_chunks ? q.taps : 1;
          for (int t = 0; t < nk; ++t) {
            mbar_wait(w_full(stage), phase);
            mbar_wait(t_full(tslot), tphase);
            tc_fence_after();
            const uint32_t sb = smem_base + w_off + stage * w_stage_bytes;
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint32_t ah = tmem_base + (uint32_t)(TC_BF_A_COL + tslot * 32 + k * 8), al = ah + 16;
              const uint64_t bh = umma_desc_sw64(sb + k * 32), bl = umma_desc_sw64(sb + b_bytes + k * 32);
              umma_bf16_ts(d_tmem, ah, bh, idesc, (first && k == 0) ? 0u : 1u);
              umma_bf16_ts(d_tmem, al, bh, idesc, 1u);
              umma_bf16_ts(d_tmem, ah, bl, idesc, 1u);
            }
            first = false;
            umma_commit(w_empty(stage));   // frees the weight stage ...
            umma_commit(t_empty(tslot));   // ... and the tensor-memory operand slot when these MMAs retire
            if (++stage == q.w_stages) { stage = 0; phase ^= 1; }
            if (++tslot == NF_TSLOTS) { tslot = 0; tphase ^= 1; }
          }
        }
        umma_commit(tfull_bar(acc));  // accumulator ready for the epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 2 + TC_EPI_WARPS) {
    tc_epilogue_role<EPI>(p, smem_raw, smem_base, bar_base, tmem_base, (uint32_t)TC_BF_ACC_STRIDE, n_units, total_tiles);
  } else {
    // ===================== operand warps (10..13): normalise-on-load, bf16 hi / lo split, tensor-memory feed =====================
    const int m = (warp & 3) * 32 + lane;  // tile row = TMEM lane owned by this thread
    const int t = threadIdx.x - TC_THREADS;  // 0..127
    const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)TC_BF_A_COL;
    float* abtab = reinterpret_cast<float*>(smem_gen + ab_off);  // [2][TB*32][2]
    const int px = m & (p.TW - 1), py = (m >> p.lTW) & (p.TH - 1), pb = m >> (p.lTW + p.lTH);
    const int r0 = (pb * q.HH2 + py) * q.HW2 + px;  // halo row of tap (0, 0) for this output pixel
    const int hplane = q.HW2 * q.HH2;
    int aslot = 0, tslot = 0, abuf = 0;
    uint32_t aphase = 0, tphase = 0, gphase = 0;
    SegIter it(p, n_units, total_tiles);
    int tile, u0, u1;
    while (it.next(tile, u0, u1)) {
      int mt = tile / p.tiles_n;
      const int tx = mt % p.tiles_x; mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int tb = mt / p.tiles_y;
      const int ox0 = tx * p.TW, oy0 = ty * p.TH, b0 = tb * p.TB;
      for (int u = u0; u < u1; ++u) {
        const bool nf = u < q.n_chunks;
        float2 abv = make_float2(0.f, 0.f);
        if (nf && t < p.TB * 32) {  // scale / shift of this chunk's channels for the images of the tile (global, L2-resident)
          const int b = b0 + (t >> 5);
          if (b < p.B) abv = __ldg(reinterpret_cast<const float2*>(q.ab) + (size_t)b * q.Cn + u * TC_BK + (t & 31));
        }
        mbar_wait(a_full(aslot), aphase);
        const uint8_t* xs = smem_gen + xs_off + (size_t)aslot * q.slot_bytes;
        float* ab = abtab + abuf * (NF_AB_BYTES / 8);
        if (nf) {
          if (t < p.TB * 32) reinterpret_cast<float2*>(ab)[t] = abv;
          if (q.has_gb) mbar_wait(gb_full, gphase);
          asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        const uint8_t* gs = smem_gen + gb_off;
        if (nf && q.taps == 9) {
          // ---- pass 1: normalise (+SPADE) (+SiLU), zero the out-of-image halo, split, write the bf16 pair back in place ----
          for (int r = t; r < q.rows_h; r += TC_SPLIT_THREADS) {
            const int hb = r / hplane;
            const int rem = r - hb * hplane;
            const int hy = rem / q.HW2, hx = rem - hy * q.HW2;
            const int ix = ox0 - q.hpad + hx, iy = oy0 - q.hpad + hy;
            const bool valid = ix >= 0 && ix < q.Win && iy >= 0 && iy < q.Hin && (b0 + hb) < p.B;
            uint8_t* row = const_cast<uint8_t*>(xs) + r * 128;
            const float4* abr = reinterpret_cast<const float4*>(ab + hb * 64);
            const int sw = r & 7;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float4 v = *reinterpret_cast<const float4*>(row + ((c ^ sw) << 4));
              const float4 s0 = abr[2 * c], s1 = abr[2 * c + 1];  // (a, b) of channels 4c, 4c+1 | 4c+2, 4c+3
              v.x = fmaf(v.x, s0.x, s0.y); v.y = fmaf(v.y, s0.z, s0.w); v.z = fmaf(v.z, s1.x, s1.y); v.w = fmaf(v.w, s1.z, s1.w);
              if (q.has_gb) {
                const float4 g = *reinterpret_cast<const float4*>(gs + r * 128 + ((c ^ sw) << 4));
                const float4 e = *reinterpret_cast<const float4*>(gs + q.slot_bytes + r * 128 + ((c ^ sw) << 4));
                v.x = fmaf(v.x, 1.0f + g.x, e.x); v.y = fmaf(v.y, 1.0f + g.y, e.y);
                v.z = fmaf(v.z, 1.0f + g.z, e.z); v.w = fmaf(v.w, 1.0f + g.w, e.w);
              }
              if (q.silu) { v.x = silu_fast(v.x); v.y = silu_fast(v.y); v.z = silu_fast(v.z); v.w = silu_fast(v.w); }
              if (!valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
              const uint32_t h0 = pack_bf16x2(v.x, v.y), h1 = pack_bf16x2(v.z, v.w);
              hi[2 * c] = h0; hi[2 * c + 1] = h1;
              lo[2 * c] = pack_bf16x2(v.x - bf16_lo_to_f32(h0), v.y - bf16_hi_to_f32(h0));
              lo[2 * c + 1] = pack_bf16x2(v.z - bf16_lo_to_f32(h1), v.w - bf16_hi_to_f32(h1));
            }
            // row layout after the pass: 16-byte chunks 0..3 = hi words, 4..7 = lo words (same swizzle)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              *reinterpret_cast<uint4*>(row + ((j ^ sw) << 4)) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
              *reinterpret_cast<uint4*>(row + (((j + 4) ^ sw) << 4)) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
          }
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (q.has_gb) {  // the gamma / beta staging is free for the next chunk
            gphase ^= 1;
            if (lane == 0) mbar_arrive(gb_empty);
          }
          // ---- pass 2: one tensor-memory slot per tap, thread = output pixel, rows shifted inside the halo tile ----
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - 3 * dy;
            const int rr = r0 + dy * q.HW2 + dx;
            const uint8_t* row = xs + rr * 128;
            const int sw = rr & 7;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 a = *reinterpret_cast<const uint4*>(row + ((j ^ sw) << 4));
              const uint4 b = *reinterpret_cast<const uint4*>(row + (((j + 4) ^ sw) << 4));
              hi[4 * j] = a.x; hi[4 * j + 1] = a.y; hi[4 * j + 2] = a.z; hi[4 * j + 3] = a.w;
              lo[4 * j] = b.x; lo[4 * j + 1] = b.y; lo[4 * j + 2] = b.z; lo[4 * j + 3] = b.w;
            }
            mbar_wait(t_empty(tslot), tphase ^ 1);  // the MMAs that last read this slot have retired
            tc_fence_after();
            tmem_st16(a_lane + (uint32_t)(tslot * 32), hi);
            tmem_st16(a_lane + (uint32_t)(tslot * 32 + 16), lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_full(tslot));
            if (++tslot == NF_TSLOTS) { tslot = 0; tphase ^= 1; }
          }
        } else {
          // ---- one k-step straight from the 128-row tile: the 1x1 normalised conv (proj_in) or the raw side input ----
          const uint8_t* row = xs + m * 128;
          const float4* abr = reinterpret_cast<const float4*>(ab + pb * 64);
          const int sw = m & 7;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 v = *reinterpret_cast<const float4*>(row + ((c ^ sw) << 4));
            if (nf) {
              const float4 s0 = abr[2 * c], s1 = abr[2 * c + 1];
              v.x = fmaf(v.x, s0.x, s0.y); v.y = fmaf(v.y, s0.z, s0.w); v.z = fmaf(v.z, s1.x, s1.y); v.w = fmaf(v.w, s1.z, s1.w);
              if (q.has_gb) {
                const float4 g = *reinterpret_cast<const float4*>(gs + m * 128 + ((c ^ sw) << 4));
                const float4 e = *reinterpret_cast<const float4*>(gs + q.slot_bytes + m * 128 + ((c ^ sw) << 4));
                v.x = fmaf(v.x, 1.0f + g.x, e.x); v.y = fmaf(v.y, 1.0f + g.y, e.y);
                v.z = fmaf(v.z, 1.0f + g.z, e.z); v.w = fmaf(v.w, 1.0f + g.w, e.w);
              }
              if (q.silu) { v.x = silu_fast(v.x); v.y = silu_fast(v.y); v.z = silu_fast(v.z); v.w = silu_fast(v.w); }
            }
            const uint32_t h0 = pack_bf16x2(v.x, v.y), h1 = pack_bf16x2(v.z, v.w);
            hi[2 * c] = h0; hi[2 * c + 1] = h1;
            lo[2 * c] = pack_bf16x2(v.x - bf16_lo_to_f32(h0), v.y - bf16_hi_to_f32(h0));
            lo[2 * c + 1] = pack_bf16x2(v.z - bf16_lo_to_f32(h1), v.w - bf16_hi_to_f32(h1));
          }
          if (nf && q.has_gb) {
            gphase ^= 1;
            __syncwarp();
            if (lane == 0) mbar_arrive(gb_empty);
          }
          mbar_wait(t_empty(tslot), tphase ^ 1);
          tc_fence_after();
          tmem_st16(a_lane + (uint32_t)(tslot * 32), hi);
          tmem_st16(a_lane + (uint32_t)(tslot * 32 + 16), lo);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_full(tslot));
          if (++tslot == NF_TSLOTS) { tslot = 0; tphase ^= 1; }
        }
        // this warp is done with the halo slot: hand it back to the TMA producer (generic-proxy accesses ordered before the
        // async-proxy refill)
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_empty(aslot));
        if (++aslot == 2) { aslot = 0; aphase ^= 1; }
        if (nf) abuf ^= 1;
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
static bool make_map4_box(CUtensorMap* m, const float* base, uint64_t C, uint64_t W, uint64_t H, uint64_t Bn, int64_t sx, int64_t sy,
                          int64_t sb, uint32_t bw, uint32_t bh, uint32_t bb) {
  return make_map4(m, base, C, W, H, Bn, sx, sy, sb, bw, bh, bb, 1u);
}

bool conv2d_nf_eligible(const FridoConvParams* p) {
  if (p->engine != 3 || !p->nrm_ab || !p->w_lo) return false;
  if (p->ups != 1 || p->stride != 1 || (p->ksize != 1 && p->ksize != 3) || p->pad != p->ksize / 2) return false;
  if (p->w_sb || p->o_sn != 1 || !p->out) return false;
  if (p->Hout != p->Hin || p->Wout != p->Win) return false;
  const int TW = next_pow2(p->Wout) < 16 ? next_pow2(p->Wout) : 16;
  const int TH = next_pow2(p->Hout) < TC_BM / TW ? next_pow2(p->Hout) : TC_BM / TW;
  const int TB = TC_BM / (TW * TH);
  const int hp = p->ksize / 2;
  if (TB > 4 || (TW + 2 * hp) * (TH + 2 * hp) * TB > NF_MAX_HALO_ROWS) return false;
  return true;
}

int conv2d_nf(const FridoConvParams* p, cudaStream_t s) {
  if (!conv2d_nf_eligible(p))
    return set_error(FRIDO_E_ARG, "conv2d_nf: normalise-on-load needs engine 3, a 3x3/1x1 stride-1 conv with shared pre-split weights, "
                                  "a dense NHWC output and at most 4 images per 128-pixel tile");
  if (!p->a0 || !p->w) return set_error(FRIDO_E_ARG, "conv2d_nf: null pointer");
  if (p->c0 % TC_BK || p->c1 % TC_BK || p->c0 <= 0) return set_error(FRIDO_E_ARG, "conv2d_nf: channels must be multiples of 32");
  if ((p->c1 > 0) != (p->a1 != nullptr)) return set_error(FRIDO_E_ARG, "conv2d_nf: a1/c1 mismatch");
  if (p->Cout % 64) return set_error(FRIDO_E_ARG, "conv2d_nf: Cout must be a multiple of 64");
  if (p->a0_sc != 1 || (p->a1 && p->a1_sc != 1)) return set_error(FRIDO_E_ARG, "conv2d_nf: channel stride must be 1");
  if (!a16(p->a0) || !a16(p->w) || !a16(p->w_lo) || (p->a1 && !a16(p->a1)) || !a16(p->out) || (p->res && !a16(p->res)) ||
      (reinterpret_cast<uintptr_t>(p->nrm_ab) & 7) || (p->nrm_gb && !a16(p->nrm_gb)))
    return set_error(FRIDO_E_ARG, "conv2d_nf: pointers must be 16-byte aligned");
  if (p->a0_sx % 4 || p->a0_sy % 4 || p->a0_sb % 4 || (p->a1 && (p->a1_sx % 4 || p->a1_sy % 4 || p->a1_sb % 4)))
    return set_error(FRIDO_E_ARG, "conv2d_nf: strides must be multiples of 16 bytes");
  if (p->cx0 < 0 || p->cx1 < 0 || p->cx0 % TC_BK || p->cx1 % TC_BK || (p->cx0 > 0) != (p->x0 != nullptr) || (p->cx1 > 0) != (p->x1 != nullptr) ||
      (p->cx1 > 0 && p->cx0 == 0))
    return set_error(FRIDO_E_ARG, "conv2d_nf: side input channels must be multiples of 32 and match x0/x1");
  if (p->cx0 && (!a16(p->x0) || (p->x1 && !a16(p->x1)) || p->x0_sx % 4 || p->x0_sy % 4 || p->x0_sb % 4 ||
                 (p->x1 && (p->x1_sx % 4 || p->x1_sy % 4 || p->x1_sb % 4))))
    return set_error(FRIDO_E_ARG, "conv2d_nf: side input needs 16-byte aligned strides");
  const int Cn = p->c0 + p->c1;
  const int taps = p->ksize * p->ksize;
  const int64_t Ktot = (int64_t)taps * Cn + p->cx0 + p->cx1;
  const int64_t w_ld = p->w_ld ? p->w_ld : Ktot;
  if (w_ld % 8) return set_error(FRIDO_E_ARG, "conv2d_nf: bf16 weight rows must be multiples of 16 bytes");
  if (p->act == FRIDO_ACT_GEGLU || p->act == FRIDO_ACT_GEGLU_FAST) return set_error(FRIDO_E_ARG, "conv2d_nf: no GEGLU");
  if (p->o_sp % 4 || p->o_sb % 4) return set_error(FRIDO_E_ARG, "conv2d_nf: output rows must be 16-byte aligned");
  if ((p->bias && !a16(p->bias)) || (p->rowvec && (!a16(p->rowvec) || p->rowvec_sb % 4)))
    return set_error(FRIDO_E_ARG, "conv2d_nf: bias / rowvec must be 16-byte aligned");
  if ((p->out_hi != nullptr) != (p->out_lo != nullptr)) return set_error(FRIDO_E_ARG, "conv2d_nf: out_hi/out_lo must come together");

  TcParams t;
  NfParams q;
  t.B = p->B; t.Hout = p->Hout; t.Wout = p->Wout; t.Cout = p->Cout;
  t.c0 = p->c0; t.c1 = p->c1; t.cx0 = p->cx0; t.cx1 = p->cx1; t.ksize = p->ksize; t.pad = p->pad; t.stride = 1;
  t.TW = next_pow2(p->Wout) < 16 ? next_pow2(p->Wout) : 16;   // squarer tiles than conv_tc.cu: the halo overhead is (TW+2)(TH+2)/(TW TH)
  t.TH = next_pow2(p->Hout) < TC_BM / t.TW ? next_pow2(p->Hout) : TC_BM / t.TW;
  t.TB = TC_BM / (t.TW * t.TH);
  t.lTW = 0; while ((1 << t.lTW) < t.TW) ++t.lTW;
  t.lTH = 0; while ((1 << t.lTH) < t.TH) ++t.lTH;
  t.tiles_x = (p->Wout + t.TW - 1) / t.TW;
  t.tiles_y = (p->Hout + t.TH - 1) / t.TH;
  t.tiles_b = (p->B + t.TB - 1) / t.TB;
  const int m_tiles = t.tiles_x * t.tiles_y * t.tiles_b;
  q.n_chunks = Cn / TC_BK; q.taps = taps; q.hpad = p->ksize / 2;
  q.HW2 = t.TW + 2 * q.hpad; q.HH2 = t.TH + 2 * q.hpad; q.rows_h = q.HW2 * q.HH2 * t.TB;
  q.side_units = (p->cx0 + p->cx1) / TC_BK;
  q.slot_bytes = (q.rows_h * 128 + 1023) / 1024 * 1024;
  q.has_gb = p->nrm_gb != nullptr; q.Cn = Cn; q.silu = p->nrm_silu; q.Hin = p->Hin; q.Win = p->Win; q.ab = p->nrm_ab;
  const int n_units = q.n_chunks + q.side_units;
  const int ksteps = q.n_chunks * taps + q.side_units;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ring_budget = TC_SMEM_BUDGET - NF_AB_BYTES - (q.has_gb ? 4 : 2) * q.slot_bytes;
  auto stage_clk = [&](int n) { return 256 + 5 * n / 2; };  // per k-step, as conv_tc.cu's BF16x3 model
  const int cands[3] = {192, 128, 64};
  int bn = 64;
  double best_cost = 1e30;
  for (int i = 0; i < 3; ++i) {
    if (p->Cout % cands[i] || ring_budget / (cands[i] * 128) < 2) continue;
    const int64_t tiles = (int64_t)m_tiles * (p->Cout / cands[i]);
    const double cost = (double)((tiles + sms - 1) / sms) * stage_clk(cands[i]);
    if (cost < best_cost) { best_cost = cost; bn = cands[i]; }
  }
  if (const char* f = getenv("FRIDO_TC_FORCE_BN")) {
    const int v = atoi(f);
    if (v >= 64 && v <= TC_BF_ACC_STRIDE && v % 64 == 0 && p->Cout % v == 0 && ring_budget / (v * 128) >= 2) bn = v;
  }
  // stream-K over operand UNITS (a chunk's 9 taps stay together: the halo tile is normalised once)
  t.sk = 0; t.sk_per = 0; t.sk_ws = nullptr; t.sk_cnt = nullptr;
  int sk_grid = 0;
  {
    const char* sk_e = getenv("FRIDO_SK");
    const int sk_env = sk_e ? atoi(sk_e) : 1;
    const double k_per_unit = (double)ksteps / n_units;
    const int64_t dp_tiles = (int64_t)m_tiles * (p->Cout / bn);
    const double dp_cost = (double)((dp_tiles + sms - 1) / sms) * ksteps * stage_clk(bn);
    double thresh = 0.95;
    if (const char* e = getenv("FRIDO_SK_THRESH")) thresh = atof(e);
    double best = sk_env == 2 ? 1e30 : thresh * dp_cost;
    const bool forced_bn = getenv("FRIDO_TC_FORCE_BN") != nullptr;
    if (sk_env && p->sk_ws && (reinterpret_cast<uintptr_t>(p->sk_ws) & 15) == 0 && n_units >= 4) {
      for (int i = 0; i < 3; ++i) {
        const int n = cands[i];
        if (p->Cout % n || ring_budget / (n * 128) < 2 || (forced_bn && n != bn)) continue;
        const int64_t tiles = (int64_t)m_tiles * (p->Cout / n);
        const int64_t iters = tiles * n_units;
        if (tiles > 1024 || iters > (1 << 28)) continue;
        int64_t g = iters / 2 < sms ? iters / 2 : sms;   // at least 2 units per CTA
        if (g < 1) g = 1;
        int64_t per = (iters + g - 1) / g;
        const int64_t min_per = (n_units + 5) / 6;         // at most 7 contributors per tile
        if (per < min_per) per = min_per;
        g = (iters + per - 1) / per;
        if (per % n_units == 0 && sk_env != 2) continue;   // whole tiles only: that is the data-parallel schedule
        const int64_t need = 4096 + 2 * g * 128 * n * 4;
        if (need > p->sk_ws_bytes) continue;
        const int contrib = (int)((n_units + per - 1) / per) + 1;
        const double cost = (double)per * k_per_unit * stage_clk(n) + 3000.0 + 8.0 * n * (1 + contrib);
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
  t.w_batched = 0;
  t.bias = p->bias; t.rowvec = p->rowvec; t.rowvec_sb = p->rowvec_sb; t.res = p->res;
  t.alpha = p->alpha; t.act = p->act; t.out = p->out; t.o_sb = p->o_sb; t.o_sp = p->o_sp; t.o_sn = p->o_sn;
  t.round_tf32 = p->round_tf32;
  t.csum = p->chan_sums;
  t.out_hi = (uint16_t*)p->out_hi; t.out_lo = (uint16_t*)p->out_lo;
  t.stages = 0;
  if (p->chan_sums && (t.TW * t.TH < 32 || t.TB > 4)) return set_error(FRIDO_E_ARG, "conv2d_nf: chan_sums needs >= 32 pixels per image");
  q.w_stages = ring_budget / (bn * 128);
  if (q.w_stages > TC_MAX_STAGES) q.w_stages = TC_MAX_STAGES;

  CUtensorMap ma0, ma1, mgb, mw, mwlo, mx0, mx1;
  if (!make_map4_box(&ma0, p->a0, p->c0, p->Win, p->Hin, p->B, p->a0_sx, p->a0_sy, p->a0_sb, q.HW2, q.HH2, t.TB))
    return set_error(FRIDO_E_ARG, "conv2d_nf: cuTensorMapEncodeTiled(a0) failed");
  ma1 = ma0;
  if (p->a1 && !make_map4_box(&ma1, p->a1, p->c1, p->Win, p->Hin, p->B, p->a1_sx, p->a1_sy, p->a1_sb, q.HW2, q.HH2, t.TB))
    return set_error(FRIDO_E_ARG, "conv2d_nf: cuTensorMapEncodeTiled(a1) failed");
  mgb = ma0;
  if (p->nrm_gb && !make_map4_box(&mgb, p->nrm_gb, 2 * Cn, p->Win, p->Hin, p->B, 2 * Cn, (int64_t)p->Win * 2 * Cn,
                                  (int64_t)p->Hin * p->Win * 2 * Cn, q.HW2, q.HH2, t.TB))
    return set_error(FRIDO_E_ARG, "conv2d_nf: cuTensorMapEncodeTiled(gb) failed");
  mx0 = ma0; mx1 = ma0;
  if (p->x0 && !make_map4_box(&mx0, p->x0, p->cx0, p->Wout, p->Hout, p->B, p->x0_sx, p->x0_sy, p->x0_sb, t.TW, t.TH, t.TB))
    return set_error(FRIDO_E_ARG, "conv2d_nf: cuTensorMapEncodeTiled(x0) failed");
  if (p->x1 && !make_map4_box(&mx1, p->x1, p->cx1, p->Wout, p->Hout, p->B, p->x1_sx, p->x1_sy, p->x1_sb, t.TW, t.TH, t.TB))
    return set_error(FRIDO_E_ARG, "conv2d_nf: cuTensorMapEncodeTiled(x1) failed");
  if (!make_map3(&mw, p->w, Ktot, p->Cout, 1, w_ld, 0, bn, true) || !make_map3(&mwlo, p->w_lo, Ktot, p->Cout, 1, w_ld, 0, bn, true))
    return set_error(FRIDO_E_ARG, "conv2d_nf: cuTensorMapEncodeTiled(w) failed");

  using KernelFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, TcParams, NfParams);
  static const KernelFn kernels[EPI_COUNT] = {conv_nf_kernel<EPI_GENERIC>, conv_nf_kernel<EPI_BIAS>, conv_nf_kernel<EPI_BIAS_RES>,
                                              conv_nf_kernel<EPI_BIAS_RV_CS>, conv_nf_kernel<EPI_BIAS_RES_CS>, conv_nf_kernel<EPI_GENERIC>,
                                              conv_nf_kernel<EPI_BIAS_CS>, conv_nf_kernel<EPI_BIAS_PAIR>};
  static bool attr[64] = {};
  if (dev >= 0 && dev < 64 && !attr[dev]) {  // the opt-in is per device
    for (int i = 0; i < EPI_COUNT; ++i)
      if (cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess)
        return set_error(FRIDO_E_LAUNCH, "conv2d_nf: cannot opt in to dynamic shared memory");
    attr[dev] = true;
  }
  int epi = EPI_GENERIC;
  {
    const char* e = getenv("FRIDO_EPI_SPEC");
    const bool spec = !e || atoi(e) != 0;
    if (spec && p->alpha == 1.0f && !p->round_tf32 && p->act == FRIDO_ACT_NONE) {
      const bool res = p->res != nullptr, rv = p->rowvec != nullptr, cs = p->chan_sums != nullptr;
      if (p->out_hi) {
        if (!res && !rv && !cs) epi = EPI_BIAS_PAIR;
      } else if (!res && !rv && !cs) epi = EPI_BIAS;
      else if (res && !rv && !cs) epi = EPI_BIAS_RES;
      else if (!res && rv && cs) epi = EPI_BIAS_RV_CS;
      else if (res && !rv && cs) epi = EPI_BIAS_RES_CS;
      else if (!res && !rv && cs) epi = EPI_BIAS_CS;
    }
  }
  const int total = m_tiles * t.tiles_n;
  const int grid = t.sk ? sk_grid : (total < sms ? total : sms);
  launch_pdl(kernels[epi], dim3(grid), dim3(NF_THREADS), TC_SMEM_BYTES, s, ma0, ma1, mgb, mw, mwlo, mx0, mx1, t, q);
  return check_launch("conv2d_nf");
}

}  // namespace frido
