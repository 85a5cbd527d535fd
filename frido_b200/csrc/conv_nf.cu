// tcgen05 BF16x3 implicit-GEMM engine with the INPUT NORMALISATION APPLIED ON LOAD ("normalise-on-load", sm_100a):
//
//   out = conv3x3 / conv1x1 ( act( GroupNorm(x) [* (1 + gamma_map) + beta_map] ) )  [+ fused 1x1 side input, bias, ...]
//
// i.e. a ResBlock's `in_layers` / `out_layers` (pyunet.py:209-240: normalization -> SiLU -> conv, SPADE modulation
// spade_norm.py:44-60) and a SpatialTransformer's `norm -> proj_in` (attention.py:254-262,296-298) as ONE launch: the
// normalised activation never exists in HBM (the separate norm_act pass was 14 % of a UNet step and ~4.8 GB of HBM traffic).
//
// Operand path (differs from conv_tc.cu, whose A tile is re-fetched per filter tap):
//   * the A operand of an output tile (TW x TH x TB = 128 pixels, TW <= 16) is HALO-RESIDENT: per 32-channel chunk ONE TMA
//     box {32 ch, TW+2, TH+2, TB} lands in a ring of shared-memory slots (TMA zero-fills outside the image), together with
//     the chunk's (a, b) table, and the four operand warps
//       PREP: apply y = x*a[b,c] + b[b,c]  (a = rstd*gamma_c, b = beta_c - mean*rstd*gamma_c from the producer's channel
//             sums, frido_gn_finalize), the SPADE modulation y*(1+gamma)+beta (gamma / beta fetched with cp.async two steps
//             ahead by the thread that uses them), SiLU, force the halo outside the image back to exactly 0 (conv padding
//             pads the ACTIVATED tensor), split into bf16 hi / lo and write the pair back in place;
//       FEED: for each of the 9 taps copy the shifted 128 rows (thread = output pixel) into a tensor-memory slot
//             (tcgen05.st) from where the TS-form MMAs take them.
//     PREP of chunk k+1 is software-pipelined into the FEED of chunk k (one 256-item step between consecutive taps), i.e.
//     it runs in the time the feed would wait for the tensor core anyway.  The normalisation, the activation and the
//     operand split run ONCE per element instead of once per tap, and the L2 -> shared-memory traffic of the A operand
//     drops ~6x.
//   * weights: as conv_tc.cu (pre-split bf16 hi / lo, TMA, SWIZZLE_64B); K order = chunk-major: column (tap*C + chunk*32).
//   * extra K units after the chunks: the fused 1x1 side input (ResBlock skip_connection) reads RAW activations at the
//     output pixel, exactly as in conv_tc.cu.
// Warp roles (640 threads = 5 warpgroups, registers rebalanced with setmaxnreg): 0 = weight TMA producer, 1 = TMEM allocator +
// MMA issuer, 2 = halo TMA producer, 4-11 = epilogue (shared with conv_tc.cu), 12-15 = feed (shared memory -> tensor memory,
// nothing else), 16-19 = prep (normalise / activate / split in place, running ahead of the feed through the slot ring).
#include "tc_common.cuh"

#ifndef FRIDO_NF_PAIR_DEFAULT
#define FRIDO_NF_PAIR_DEFAULT 0   // measured neutral on this engine (its feed / prep warps, not the W operand, set its pace)
#endif

namespace frido {

constexpr int NF_THREADS = 640;                 // 5 warpgroups: control | epilogue x2 | feed | prep
constexpr int NF_TSLOTS = TC_BF_MAX_STAGES;     // tensor-memory operand slots (32 columns each)
constexpr int NF_MAX_ASLOTS = 6;
constexpr int NF_MAX_HALO_ROWS = 208;
constexpr int NF_BAR_EXTRA = 0;                 // the shared barrier area is 512 B (46 slots used here)
constexpr int NF_SMEM_BYTES = TC_SMEM_BYTES + NF_BAR_EXTRA;
constexpr int NF_GB_DEPTH = 4;                  // cp.async ring of SPADE gamma / beta: three steps (36 KB per SM) in flight ahead of
                                                // the step in use - the maps stream from HBM and the latency needs the bytes


struct NfParams {
  int n_chunks;     // 32-channel chunks of the normalised input
  int taps;         // 9 (3x3) or 1 (1x1)
  int hpad;         // 1 or 0
  int HW2, HH2;     // halo box: TW + 2*hpad, TH + 2*hpad
  int rows_h;       // HW2 * HH2 * TB
  int m_plane, m_hw2;  // 16-bit reciprocal multipliers: r / (HW2*HH2) = (r * m_plane) >> 16, r / HW2 = (r * m_hw2) >> 16 (r < 4096)
  int side_units;   // (cx0 + cx1) / 32
  int slot_bytes;   // one halo tile + the (a, b) table of its channels in shared memory (1 KB multiple)
  int ab_off;       // byte offset of the table [TB][32][2] fp32 inside a slot (= rows_h * 128)
  int a_slots;      // depth of the halo ring
  int w_stages;     // depth of the weight ring
  int has_gb;       // SPADE maps present
  int has_norm;     // 0: plain conv through the halo-resident operand path (no normalisation)
  int presplit;     // the chunks arrive as finished operand rows (FridoConvParams.a_presplit): nothing to prepare
  int Cn;           // c0 + c1
  int silu;
  int Hin, Win;
  const float* ab;  // [B][Cn][2]
  const float* gb;  // [B][Hin*Win][2*Cn] or NULL
  int dbg;          // profiling aid (FRIDO_NF_DBG bit mask; results are WRONG with any bit set): 1 = prep does no arithmetic,
                    // 2 = no prep at all, 4 = no SiLU
};

// 1D bulk copy global -> shared, completion on an mbarrier (size and both addresses multiples of 16 bytes)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// SiLU on the two approximate MUFU ops (1-2 ulp each): 5 instructions per element
__device__ __forceinline__ float silu_fast(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return v * r;
}

constexpr int NF_IPT = 3;                       // prep items per thread per step: 12 independent element chains in flight
constexpr int NF_STEP_ITEMS = NF_IPT * 128;
constexpr int NF_MAX_STEPS = 5;                 // ceil(208 rows * 8 / 384)
constexpr int NF_DESC_BYTES = 8192;             // >= NF_MAX_STEPS * NF_STEP_ITEMS * 4
constexpr int NF_GB_STEP_BYTES = NF_STEP_ITEMS * 32;

// One prep step of one thread: NF_IPT items, straight-line (a missing item computes on a valid dummy address and only its
// store is predicated).  Item descriptor: [0,11) 16-byte index inside the slot (swizzled), [11,16) halo x, [16,21) halo y,
// [21,23) image, bit 31 = the item exists in the halo tile.  NRM / GB / SILU are compile-time so the hot variants carry no
// selects: the role is bound by instruction issue (one warp per scheduler), every removed instruction counts.
template <bool NRM, bool GB, bool SILU>
__device__ __forceinline__ void nf_prep_items(uint8_t* xs, const uint32_t* dsc, int first_row, int row_limit, bool halo, int ab_off,
                                              int cq, const uint8_t* gbs, int ox, int oy, int b0, int Win, int Hin, int B, int TB) {
  uint32_t d[NF_IPT];
  float4 v[NF_IPT];
#pragma unroll
  for (int j = 0; j < NF_IPT; ++j) {
    d[j] = dsc[j * 128];
    if (first_row + 16 * j >= row_limit) d[j] &= 0x7FFFFFFFu;   // side tiles have 128 rows
    v[j] = *reinterpret_cast<const float4*>(xs + ((d[j] >> 31) ? (d[j] & 0x7FFu) << 4 : 0u));
  }
#pragma unroll
  for (int j = 0; j < NF_IPT; ++j) {
    const int hx = (d[j] >> 11) & 31, hy = (d[j] >> 16) & 31, hb = (d[j] >> 21) & 3;
    float4 w = v[j];
    if (NRM) {
      const float4* abp = reinterpret_cast<const float4*>(xs + ab_off + min(hb, TB - 1) * 256 + cq * 32);
      const float4 s0 = abp[0], s1 = abp[1];  // (a, b) of channels 4cq, 4cq+1 | 4cq+2, 4cq+3 of the item's image
      w.x = fmaf(w.x, s0.x, s0.y); w.y = fmaf(w.y, s0.z, s0.w); w.z = fmaf(w.z, s1.x, s1.y); w.w = fmaf(w.w, s1.z, s1.w);
      if (GB) {
        const float4* gp = reinterpret_cast<const float4*>(gbs + j * (128 * 32));
        const float4 g = gp[0], e = gp[1];
        w.x = fmaf(w.x, 1.0f + g.x, e.x); w.y = fmaf(w.y, 1.0f + g.y, e.y); w.z = fmaf(w.z, 1.0f + g.z, e.z); w.w = fmaf(w.w, 1.0f + g.w, e.w);
      }
      if (SILU) { w.x = silu_fast(w.x); w.y = silu_fast(w.y); w.z = silu_fast(w.z); w.w = silu_fast(w.w); }
    }
    // conv padding pads the ACTIVATED tensor with zeros: halo pixels outside the image (TMA filled them with 0 BEFORE the
    // normalisation) go back to exactly 0
    const bool ok = !halo || ((unsigned)(ox + hx) < (unsigned)Win && (unsigned)(oy + hy) < (unsigned)Hin && b0 + hb < B);
    if (!ok) w = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t h0 = pack_bf16x2(w.x, w.y), h1 = pack_bf16x2(w.z, w.w);
    const uint32_t l0 = pack_bf16x2(w.x - bf16_lo_to_f32(h0), w.y - bf16_hi_to_f32(h0));
    const uint32_t l1 = pack_bf16x2(w.z - bf16_lo_to_f32(h1), w.w - bf16_hi_to_f32(h1));
    // the item's 16 bytes now hold hi words 2c, 2c+1 | lo words 2c, 2c+1 of its row
    if (d[j] >> 31) *reinterpret_cast<uint4*>(xs + ((d[j] & 0x7FFu) << 4)) = make_uint4(h0, h1, l0, l1);
  }
}

// K units of a tile, in the order every warp role walks them: the 32-channel chunks of the normalised input (9 k-steps
// each for a 3x3 conv) spread evenly among the 32-channel units of the raw side input (1 k-step each): chunk i sits at
// position floor(i * (nc + ns) / nc), e.g. c s c s ... for nc = ns and c s s s c s s s ... for ns = 3 nc.  A run of short
// units is then never longer than the slot ring can prefetch behind the long unit in front of it; a dozen short units in
// a row would expose the TMA + prep latency of each.
__device__ __forceinline__ void nf_decode(const NfParams& q, int u, bool& chunk, int& idx) {
  const int nc = q.n_chunks, total = q.n_chunks + q.side_units;
  const int i = (u * nc + total - 1) / total;  // number of chunks placed before position u (or at it)
  chunk = i < nc && (i * total) / nc == u;
  idx = chunk ? i : u - min(i, nc);
}

// Position in the stream of (tile, unit) pairs a CTA processes.
struct NfCursor {
  SegIter it;
  int tile, u0, u1, u, seq, idx, ox0, oy0, b0, slot;
  uint32_t phase;  // halo-ring slot of this unit and the parity of its `full` barrier
  bool valid, chunk;
  __device__ __forceinline__ NfCursor(const TcParams& p, int n_units, int total_tiles)
      : it(p, n_units, total_tiles), tile(0), u0(0), u1(0), u(-1), seq(-1), idx(0), ox0(0), oy0(0), b0(0), slot(-1), phase(0),
        valid(true), chunk(false) {}
  __device__ __forceinline__ void advance(const TcParams& p, const NfParams& q) {
    ++seq; ++u;
    if (++slot == q.a_slots) { slot = 0; phase ^= 1; }
    if (u >= u1) {
      if (!it.next(tile, u0, u1)) { valid = false; return; }
      u = u0;
      int mt = tile / p.tiles_n;
      const int tx = mt % p.tiles_x; mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int tb = mt / p.tiles_y;
      ox0 = tx * p.TW; oy0 = ty * p.TH; b0 = tb * p.TB;
    }
    nf_decode(q, u, chunk, idx);
  }
};

// register budgets per warpgroup (setmaxnreg).  The pool is what the CTA was launched with - 640 threads x 96 registers =
// 61440 - not the SM's register file: 128 x 48 + 256 x 128 + 128 x 64 + 128 x 112 = 61440 (asking for more blocks forever).
constexpr int NF_REGS_LAUNCH = 96, NF_REGS_CTRL = 48, NF_REGS_EPI = 128, NF_REGS_FEED = 64, NF_REGS_PREP = 112;
static_assert(128 * NF_REGS_CTRL + 256 * NF_REGS_EPI + 128 * NF_REGS_FEED + 128 * NF_REGS_PREP <= NF_THREADS * NF_REGS_LAUNCH,
              "setmaxnreg budgets exceed the CTA's register pool");
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

template <int EPI, bool PAIR>
__global__ void __launch_bounds__(NF_THREADS, 1)
conv_nf_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_wlo,
               const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1, const TcParams p,
               const NfParams q) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // operand region (TC_SMEM_BUDGET): halo ring (a_slots x slot_bytes) | [SPADE gamma / beta ring] | weight ring
  const uint32_t gb_ring = (uint32_t)q.a_slots * q.slot_bytes;
  const uint32_t w_off = gb_ring + (q.has_gb ? (uint32_t)(NF_GB_DEPTH * NF_GB_STEP_BYTES) : 0u);
  // CTA-pair launches (PAIR, tcgen05.mma.cta_group::2 - see conv_tc2.cu): each CTA of the cluster owns one M tile (its own halo,
  // prep and feed) and HALF of the W tile's rows; the leader CTA issues the MMAs for both
  const uint32_t rank = PAIR ? (blockIdx.x & 1u) : 0u;
  const int HBN = PAIR ? (p.BN >> 1) : p.BN;
  const uint32_t b_bytes = (uint32_t)HBN * TC_BK * 2;
  const uint32_t w_stage_bytes = 2u * b_bytes;
  const uint32_t bar_base = smem_base + TC_SMEM_BUDGET + TC_STG_BYTES + TC_CSUM_BYTES;
  // barrier slots: w_full[6] 0.. | w_empty[6] 6.. | t_full[4] 12.. | (18..23 shared with the epilogue role) | t_empty[4] 24..
  //                | a_full[6] 28.. | a_empty[6] 34.. | a_ready[6] 40..
  auto w_full = [&](int s) { return bar_base + 8u * s; };
  auto w_empty = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
  auto t_full = [&](int s) { return bar_base + 8u * (12 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (TC_BAR_TFULL + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (TC_BAR_TEMPTY + a); };
  auto t_empty = [&](int s) { return bar_base + 8u * (24 + s); };
  auto a_full = [&](int s) { return bar_base + 8u * (28 + s); };    // TMA landed (halo tile + scale / shift table)
  auto a_empty = [&](int s) { return bar_base + 8u * (34 + s); };   // the feed warps have read the slot for the last time
  auto a_ready = [&](int s) { return bar_base + 8u * (40 + s); };   // the prep warps have turned the slot into bf16 hi | lo rows
  const uint32_t tmem_slot = bar_base + 8u * TC_BAR_TMEM_SLOT;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = warp_idx_uniform();   // uniform: the role branches below are uniform control flow (see tc_common.cuh)
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  const int n_units = q.n_chunks + q.side_units;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const int total_tiles = m_tiles * p.tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a0);
    if (p.c1) prefetch_tmap(&map_a1);
    if (p.cx0) prefetch_tmap(&map_x0);
    if (p.cx1) prefetch_tmap(&map_x1);
    prefetch_tmap(&map_w);
    prefetch_tmap(&map_wlo);
    for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    const uint32_t ncta = PAIR ? 2u : 1u;  // the leader's t_full / tempty barriers collect both CTAs of a pair
    for (int s = 0; s < NF_TSLOTS; ++s) { mbar_init(t_full(s), ncta * TC_SPLIT_WARPS); mbar_init(t_empty(s), 1); }
    for (int s = 0; s < NF_MAX_ASLOTS; ++s) {
      mbar_init(a_full(s), 1); mbar_init(a_empty(s), TC_SPLIT_WARPS); mbar_init(a_ready(s), TC_SPLIT_WARPS);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), ncta * TC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers exist before anyone arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp < 4) {
    // ===================== control warpgroup: 0 = weight TMA, 1 = MMA issuer, 2 = halo TMA, 3 = idle =====================
    reg_dec<NF_REGS_CTRL>();
    // (nested so that the branches on the uniform warp index stay uniform control flow: `warp == 0 && lane == 0` is not)
    if (warp == 0) {
      {  // W-tile issuer: the whole warp runs the loop converged, one elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      NfCursor c(p, n_units, total_tiles);
      for (c.advance(p, q); c.valid; c.advance(p, q)) {
        const int n0 = (c.tile % p.tiles_n) * p.BN + (int)rank * HBN;
        const int nk = c.chunk ? q.taps : 1;
        for (int t = 0; t < nk; ++t) {
          // weight columns run [tap][channel] then the side input's channels
          const int col = c.chunk ? t * q.Cn + c.idx * TC_BK : q.taps * q.Cn + c.idx * TC_BK;
          mbar_wait(w_empty(stage), phase ^ 1);
          const uint32_t sb = smem_base + w_off + stage * w_stage_bytes;
          mbar_expect_tx_elect(w_full(stage), w_stage_bytes);
          tma_load_3d_elect(sb, &map_w, w_full(stage), col, n0, 0);
          tma_load_3d_elect(sb + b_bytes, &map_wlo, w_full(stage), col, n0, 0);
          if (++stage == q.w_stages) { stage = 0; phase ^= 1; }
        }
      }
      }
    } else if (warp == 2) {
      {  // halo-tile issuer: converged warp, elected issue
      NfCursor c(p, n_units, total_tiles);
      for (c.advance(p, q); c.valid; c.advance(p, q)) {
        mbar_wait(a_empty(c.slot), c.phase ^ 1);
        const uint32_t dst = smem_base + c.slot * q.slot_bytes;
        const int ch = c.idx * TC_BK;
        if (c.chunk) {
          // the (a, b) pairs of this chunk's 32 channels for the images of the tile ride on the same barrier
          const int nimg = q.has_norm ? min(p.TB, p.B - c.b0) : 0;
          mbar_expect_tx_elect(a_full(c.slot), (uint32_t)q.rows_h * 128u + (uint32_t)nimg * 256u);
          if (ch < p.c0) tma_load_4d_elect(dst, &map_a0, a_full(c.slot), ch, c.ox0 - q.hpad, c.oy0 - q.hpad, c.b0);
          else           tma_load_4d_elect(dst, &map_a1, a_full(c.slot), ch - p.c0, c.ox0 - q.hpad, c.oy0 - q.hpad, c.b0);
          for (int i = 0; i < nimg; ++i)
            bulk_load_1d_elect(dst + q.ab_off + i * 256, q.ab + ((size_t)(c.b0 + i) * q.Cn + ch) * 2, 256u, a_full(c.slot));
        } else {  // side input: the output pixel itself, raw
          mbar_expect_tx_elect(a_full(c.slot), (uint32_t)TC_A_BYTES);
          if (ch < p.cx0) tma_load_4d_elect(dst, &map_x0, a_full(c.slot), ch, c.ox0, c.oy0, c.b0);
          else            tma_load_4d_elect(dst, &map_x1, a_full(c.slot), ch - p.cx0, c.ox0, c.oy0, c.b0);
        }
      }
      }
    } else if (warp == 1 && rank == 0) {
      // MMA issuer: the whole warp runs the loop converged, one elected lane issues (uniform datapath, no R2UR waterfalls)
      const bool pr = PAIR;
      const uint32_t idesc = umma_idesc_bf16(pr ? 2 * TC_BM : TC_BM, p.BN);
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
          bool chunk; int idx;
          nf_decode(q, u, chunk, idx);
          const int nk = chunk ? q.taps : 1;
          for (int t = 0; t < nk; ++t) {
            mbar_wait(w_full(stage), phase);
            mbar_wait(t_full(tslot), tphase);
            tc_fence_after();
            const uint32_t sb = smem_base + w_off + stage * w_stage_bytes;
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint32_t ah = tmem_base + (uint32_t)(TC_BF_A_COL + tslot * 32 + k * 8), al = ah + 16;
              const uint64_t bh = umma_desc_sw64(sb + k * 32), bl = umma_desc_sw64(sb + b_bytes + k * 32);
              if (pr) {
                umma_bf16_ts_2cta_elect(d_tmem, ah, bh, idesc, (first && k == 0) ? 0u : 1u);
                umma_bf16_ts_2cta_elect(d_tmem, al, bh, idesc, 1u);
                umma_bf16_ts_2cta_elect(d_tmem, ah, bl, idesc, 1u);
              } else {
                umma_bf16_ts_elect(d_tmem, ah, bh, idesc, (first && k == 0) ? 0u : 1u);
                umma_bf16_ts_elect(d_tmem, al, bh, idesc, 1u);
                umma_bf16_ts_elect(d_tmem, ah, bl, idesc, 1u);
              }
            }
            first = false;
            // frees the weight stage and the tensor-memory operand slot when these MMAs retire (in both CTAs of a pair)
            if (pr) { umma_commit_2cta_elect(w_empty(stage)); umma_commit_2cta_elect(t_empty(tslot)); }
            else    { umma_commit_elect(w_empty(stage)); umma_commit_elect(t_empty(tslot)); }
            if (++stage == q.w_stages) { stage = 0; phase ^= 1; }
            if (++tslot == NF_TSLOTS) { tslot = 0; tphase ^= 1; }
          }
        }
        if (pr) umma_commit_2cta_elect(tfull_bar(acc)); else umma_commit_elect(tfull_bar(acc));  // accumulator ready for the epilogue(s)
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 4 + TC_EPI_WARPS) {
    reg_inc<NF_REGS_EPI>();
    tc_epilogue_role<EPI, 4>(p, smem_raw, smem_base, bar_base, tmem_base, (uint32_t)TC_BF_ACC_STRIDE, n_units, total_tiles);
  } else if (warp < 16) {
    // ===================== feed warpgroup (12..15): shared memory -> tensor memory, nothing else =====================
    // Every slot it sees already holds finished operand rows (128 B = per 16-byte chunk c: bf16x2 hi words 2c, 2c+1 | lo
    // words 2c, 2c+1), written by the prep warpgroup.  Per k-step thread = output pixel reads its row - for a 3x3 chunk
    // the row of tap (dy, dx) inside the halo tile - and stores it into a tensor-memory slot (tcgen05.st) for the TS-form
    // MMAs.  The store of k-step k is only waited for at the start of k-step k+1 (after its shared-memory reads).
    reg_dec<NF_REGS_FEED>();
    const int m = (warp & 3) * 32 + lane;    // tile row = TMEM lane owned by this thread
    const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)TC_BF_A_COL;
    const int px = m & (p.TW - 1), py = (m >> p.lTW) & (p.TH - 1), pb = m >> (p.lTW + p.lTH);
    const int r0 = (pb * q.HH2 + py) * q.HW2 + px;  // halo row of tap (0, 0) for this output pixel
    int tseq = 0, pend = -1;
    // pair launches: a k-step is only announced (to the LEADER's t_full) once this CTA's half of the W tile has landed too -
    // the weight ring advances in lockstep with the k-steps, so the feed can follow it
    int wst = 0;
    uint32_t wph = 0;
    auto announce = [&](int slot_) {
      if (PAIR) {
        mbar_wait(w_full(wst), wph);
        if (++wst == q.w_stages) { wst = 0; wph ^= 1; }
        if (lane == 0) mbar_arrive_cluster(mapa_rank(t_full(slot_), 0));
      } else if (lane == 0) {
        mbar_arrive(t_full(slot_));
      }
    };
    NfCursor F(p, n_units, total_tiles);
    for (F.advance(p, q); F.valid; F.advance(p, q)) {
      const uint8_t* xs = smem_gen + (size_t)F.slot * q.slot_bytes;
      const bool halo = F.chunk && q.taps == 9;
      const int nk = F.chunk ? q.taps : 1;
      mbar_wait(a_ready(F.slot), F.phase);
      for (int tap = 0; tap < nk; ++tap) {
        int rr = m;
        if (halo) { const int dy = tap / 3, dx = tap - 3 * dy; rr = r0 + dy * q.HW2 + dx; }
        const uint8_t* row = xs + rr * 128;
        const int sw = rr & 7;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 w = *reinterpret_cast<const uint4*>(row + ((c ^ sw) << 4));
          hi[2 * c] = w.x; hi[2 * c + 1] = w.y; lo[2 * c] = w.z; lo[2 * c + 1] = w.w;
        }
        if (pend >= 0) {  // complete the previous k-step: its tcgen05.st had this k-step's loads to land
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          announce(pend);
        }
        const int tslot = tseq & (NF_TSLOTS - 1);
        mbar_wait(t_empty(tslot), (uint32_t)(((tseq >> 2) & 1) ^ 1));  // the MMAs that last read this slot have retired
        tc_fence_after();
        tmem_st16(a_lane + (uint32_t)(tslot * 32), hi);
        tmem_st16(a_lane + (uint32_t)(tslot * 32 + 16), lo);
        pend = tslot;
        ++tseq;
      }
      // done with the slot (its rows are in registers / tensor memory): hand it back to the TMA producer
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_empty(F.slot));
    }
    if (pend >= 0) {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      announce(pend);
    }
    static_assert(NF_TSLOTS == 4, "the feed assumes four tensor-memory operand slots");
  } else {
    // ===================== prep warpgroup (16..19): turn a landed slot into operand rows, in place =====================
    // For a chunk of the normalised input: y = x*a[b,c] + b[b,c], SPADE y*(1+gamma)+beta, SiLU, zero outside the image
    // (conv padding pads the ACTIVATED tensor), bf16 hi / lo split.  For a raw side-input tile: the split only.
    // Item = (row, 16-byte chunk = 4 channels); its hi / lo words go back into the 16 bytes it read, so items are
    // independent.  Item k of thread t is (row t/8 + 16 k, chunk t%8): the thread's channels never change, 8 consecutive
    // threads cover one 128-byte row (conflict-free), halo coordinates advance incrementally.  The warpgroup runs ahead of
    // the feed through the slot ring (bounded by `a_full`), so its latency never sits on the tensor core's critical path.
    reg_inc<NF_REGS_PREP>();
    const int t = threadIdx.x - 512;  // 0..127
    const int cq = t & 7, rq = t >> 3;
    const int hplane = q.HW2 * q.HH2;
    const int n_steps = (q.rows_h * 8 + NF_STEP_ITEMS - 1) / NF_STEP_ITEMS;  // per chunk (side tiles: 128 rows)
    const int n_steps_side = (TC_BM * 8 + NF_STEP_ITEMS - 1) / NF_STEP_ITEMS;
    // item descriptors, worked out once (they depend on neither the chunk nor the tile); each thread reads back only the
    // entries it wrote
    uint32_t* desc = reinterpret_cast<uint32_t*>(smem_gen + TC_SMEM_BUDGET - NF_DESC_BYTES);
    for (int k = 0; k < n_steps * NF_IPT; ++k) {
      const int r = rq + 16 * k;
      uint32_t d = 0;
      if (r < q.rows_h) {
        const int hb = r / hplane, rem = r - hb * hplane, hy = rem / q.HW2, hx = rem - hy * q.HW2;
        d = (uint32_t)(r * 8 + (cq ^ (r & 7))) | ((uint32_t)hx << 11) | ((uint32_t)hy << 16) | ((uint32_t)hb << 21) | 0x80000000u;
      }
      desc[k * 128 + t] = d;
    }
    const bool do_gb = q.has_gb != 0, do_silu = q.silu && !(q.dbg & 4);
    NfCursor P(p, n_units, total_tiles), G(p, n_units, total_tiles);
    auto next_chunk = [&](NfCursor& c) { do { c.advance(p, q); } while (c.valid && !c.chunk); };
    int g_step = 0, g_ring = 0, p_ring = 0;
    // cp.async group of one prep step: gamma | beta (16 B each) of this thread's items; always commits (possibly empty)
    auto gb_issue = [&]() {
      if (G.valid) {
#pragma unroll
        for (int j = 0; j < NF_IPT; ++j) {
          const uint32_t d = desc[(g_step * NF_IPT + j) * 128 + t];
          if (d >> 31) {
            const int ix = G.ox0 - q.hpad + (int)((d >> 11) & 31), iy = G.oy0 - q.hpad + (int)((d >> 16) & 31), b = G.b0 + (int)((d >> 21) & 3);
            const bool ok = (unsigned)ix < (unsigned)q.Win && (unsigned)iy < (unsigned)q.Hin && b < p.B;
            const float* src = q.gb + (ok ? (((size_t)b * q.Hin + iy) * q.Win + ix) * (size_t)(2 * q.Cn) + G.idx * TC_BK + 4 * cq : 0);
            const uint32_t dst = smem_base + gb_ring + g_ring * NF_GB_STEP_BYTES + (j * 128 + t) * 32;
            const int nbytes = ok ? 16 : 0;  // src-size 0: zero fill, nothing is read
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16), "l"(src + (ok ? q.Cn : 0)), "r"(nbytes) : "memory");
          }
        }
        if (++g_ring == NF_GB_DEPTH) g_ring = 0;
        if (++g_step == n_steps) { g_step = 0; next_chunk(G); }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (do_gb) {  // NF_GB_DEPTH - 1 steps of SPADE maps are always in flight ahead of the step being computed
      next_chunk(G);
      for (int i = 0; i < NF_GB_DEPTH - 1; ++i) gb_issue();
    }
    for (P.advance(p, q); P.valid; P.advance(p, q)) {
      mbar_wait(a_full(P.slot), P.phase);
      uint8_t* xs = smem_gen + (size_t)P.slot * q.slot_bytes;
      const bool nrm = P.chunk && q.has_norm && !(q.dbg & 1);
      const bool gbs = nrm && do_gb;
      const int rows = P.chunk ? q.rows_h : TC_BM;
      const int steps = ((q.dbg & 2) || (P.chunk && q.presplit)) ? 0 : (P.chunk ? n_steps : n_steps_side);
      const int ox = P.ox0 - q.hpad, oy = P.oy0 - q.hpad;
      for (int st = 0; st < steps; ++st) {
        const uint32_t* dsc = desc + st * NF_STEP_ITEMS + t;
        const uint8_t* gsm = smem_gen + gb_ring + p_ring * NF_GB_STEP_BYTES + t * 32;
        const int first_row = rq + 16 * NF_IPT * st;
        if (gbs) {
          gb_issue();  // step n + DEPTH - 1 goes out; all but the newest DEPTH - 1 groups (i.e. step n) must have landed
          asm volatile("cp.async.wait_group %0;" ::"n"(NF_GB_DEPTH - 1) : "memory");
          if (do_silu) nf_prep_items<true, true, true>(xs, dsc, first_row, rows, true, q.ab_off, cq, gsm, ox, oy, P.b0, q.Win, q.Hin, p.B, p.TB);
          else         nf_prep_items<true, true, false>(xs, dsc, first_row, rows, true, q.ab_off, cq, gsm, ox, oy, P.b0, q.Win, q.Hin, p.B, p.TB);
          if (++p_ring == NF_GB_DEPTH) p_ring = 0;
        } else if (nrm) {
          if (do_silu) nf_prep_items<true, false, true>(xs, dsc, first_row, rows, true, q.ab_off, cq, gsm, ox, oy, P.b0, q.Win, q.Hin, p.B, p.TB);
          else         nf_prep_items<true, false, false>(xs, dsc, first_row, rows, true, q.ab_off, cq, gsm, ox, oy, P.b0, q.Win, q.Hin, p.B, p.TB);
        } else {  // raw tile (side input, or a plain conv through the halo path): mask + split only
          nf_prep_items<false, false, false>(xs, dsc, first_row, rows, P.chunk, q.ab_off, cq, gsm, ox, oy, P.b0, q.Win, q.Hin, p.B, p.TB);
        }
      }
      // rows are final: release them to the feed warps (and order these generic-proxy writes before the TMA refill that
      // follows the feed's release of the slot)
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready(P.slot));
    }
    if (do_gb) asm volatile("cp.async.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // nobody frees tensor memory or exits while the pair's MMAs / remote arrives may still touch it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ host side
static bool make_map4_box(CUtensorMap* m, const float* base, uint64_t C, uint64_t W, uint64_t H, uint64_t Bn, int64_t sx, int64_t sy,
                          int64_t sb, uint32_t bw, uint32_t bh, uint32_t bb) {
  return make_map4(m, base, C, W, H, Bn, sx, sy, sb, bw, bh, bb, 1u);
}

bool conv2d_nf_eligible(const FridoConvParams* p) {
  if (p->engine != 3 || (!p->nrm_ab && !p->a_presplit) || !p->w_lo) return false;
  if (p->a_presplit && (p->nrm_ab || p->ksize != 3)) return false;
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
      (p->nrm_ab && !a16(p->nrm_ab)) || (p->nrm_gb && !a16(p->nrm_gb)))
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
  q.ab_off = q.rows_h * 128;
  q.slot_bytes = (q.ab_off + t.TB * 256 + 1023) / 1024 * 1024;
  q.m_plane = 65536 / (q.HW2 * q.HH2) + 1; q.m_hw2 = 65536 / q.HW2 + 1;  // exact for r < 4096 (r <= 207 here)
  q.has_gb = p->nrm_gb != nullptr; q.has_norm = p->nrm_ab != nullptr; q.presplit = p->a_presplit != 0; q.Cn = Cn; q.silu = p->nrm_silu; q.Hin = p->Hin; q.Win = p->Win;
  q.ab = p->nrm_ab; q.gb = p->nrm_gb;
  q.dbg = 0;
  if (const char* e = getenv("FRIDO_NF_DBG")) q.dbg = atoi(e);
  // halo ring: about half of the operand budget (4 slots of a 3x3 halo tile, 6 of a 1x1 tile), the rest is the weight ring
  q.a_slots = p->ksize == 3 ? (q.has_gb ? 3 : 4) : NF_MAX_ASLOTS;
  const int n_units = q.n_chunks + q.side_units;
  const int ksteps = q.n_chunks * taps + q.side_units;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ring_budget = TC_SMEM_BUDGET - q.a_slots * q.slot_bytes - (q.has_gb ? NF_GB_DEPTH * NF_GB_STEP_BYTES : 0) -
                          NF_DESC_BYTES;
  auto stage_clk = [&](int n) { return bf_stage_clk(n); };  // per k-step, as conv_tc.cu's BF16x3 model
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
  t.sk = 0; t.sk_per = 0; t.sk_ws = nullptr; t.sk_cnt = nullptr; t.pair = 0; t.dbg_w = nullptr; t.dbg = 0; t.a_stages = 0;
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
  // CTA-pair schedule (cta_group::2): FRIDO_NF_PAIR = 0 never (default) | 1 launches with a full wave of pairs | 2 whenever legal
  {
    const char* e = getenv("FRIDO_NF_PAIR");
    const int pair_env = e ? atoi(e) : FRIDO_NF_PAIR_DEFAULT;
    const long long items = (long long)(m_tiles / 2) * t.tiles_n;
    if (pair_env && !t.sk && m_tiles % 2 == 0 && (pair_env == 2 || items >= sms / 2)) t.pair = 1;
  }
  const int w_rows = t.pair ? bn / 2 : bn;   // W rows per CTA and stage
  q.w_stages = ring_budget / (w_rows * 128);
  if (q.w_stages > TC_MAX_STAGES) q.w_stages = TC_MAX_STAGES;

  CUtensorMap ma0, ma1, mw, mwlo, mx0, mx1;
  if (!make_map4_box(&ma0, p->a0, p->c0, p->Win, p->Hin, p->B, p->a0_sx, p->a0_sy, p->a0_sb, q.HW2, q.HH2, t.TB))
    return set_error(FRIDO_E_ARG, "conv2d_nf: cuTensorMapEncodeTiled(a0) failed");
  ma1 = ma0;
  if (p->a1 && !make_map4_box(&ma1, p->a1, p->c1, p->Win, p->Hin, p->B, p->a1_sx, p->a1_sy, p->a1_sb, q.HW2, q.HH2, t.TB))
    return set_error(FRIDO_E_ARG, "conv2d_nf: cuTensorMapEncodeTiled(a1) failed");
  mx0 = ma0; mx1 = ma0;
  if (p->x0 && !make_map4_box(&mx0, p->x0, p->cx0, p->Wout, p->Hout, p->B, p->x0_sx, p->x0_sy, p->x0_sb, t.TW, t.TH, t.TB))
    return set_error(FRIDO_E_ARG, "conv2d_nf: cuTensorMapEncodeTiled(x0) failed");
  if (p->x1 && !make_map4_box(&mx1, p->x1, p->cx1, p->Wout, p->Hout, p->B, p->x1_sx, p->x1_sy, p->x1_sb, t.TW, t.TH, t.TB))
    return set_error(FRIDO_E_ARG, "conv2d_nf: cuTensorMapEncodeTiled(x1) failed");
  if (!make_map3(&mw, p->w, Ktot, p->Cout, 1, w_ld, 0, w_rows, true) || !make_map3(&mwlo, p->w_lo, Ktot, p->Cout, 1, w_ld, 0, w_rows, true))
    return set_error(FRIDO_E_ARG, "conv2d_nf: cuTensorMapEncodeTiled(w) failed");

  using KernelFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, TcParams, NfParams);
  // [0..7]: single-CTA kernels, [8..15]: CTA-pair kernels (a kernel that contains cta_group::2 / cluster-barrier instructions
  // cannot be launched without a cluster: "cluster misconfiguration")
  static const KernelFn kernels_all[2 * EPI_COUNT] = {
      conv_nf_kernel<EPI_GENERIC, false>, conv_nf_kernel<EPI_BIAS, false>, conv_nf_kernel<EPI_BIAS_RES, false>,
      conv_nf_kernel<EPI_BIAS_RV_CS, false>, conv_nf_kernel<EPI_BIAS_RES_CS, false>, conv_nf_kernel<EPI_GENERIC, false>,
      conv_nf_kernel<EPI_BIAS_CS, false>, conv_nf_kernel<EPI_BIAS_PAIR, false>,
      conv_nf_kernel<EPI_GENERIC, true>, conv_nf_kernel<EPI_BIAS, true>, conv_nf_kernel<EPI_BIAS_RES, true>,
      conv_nf_kernel<EPI_BIAS_RV_CS, true>, conv_nf_kernel<EPI_BIAS_RES_CS, true>, conv_nf_kernel<EPI_GENERIC, true>,
      conv_nf_kernel<EPI_BIAS_CS, true>, conv_nf_kernel<EPI_BIAS_PAIR, true>};
  const KernelFn* kernels = kernels_all + (t.pair ? EPI_COUNT : 0);
  static bool attr[64] = {};
  if (dev >= 0 && dev < 64 && !attr[dev]) {  // the opt-in is per device
    for (int i = 0; i < 2 * EPI_COUNT; ++i)
      if (cudaFuncSetAttribute(kernels_all[i], cudaFuncAttributeMaxDynamicSharedMemorySize, NF_SMEM_BYTES) != cudaSuccess)
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
  if (t.pair) {  // clusters of two CTAs, one (M-tile pair, N tile) item per cluster at a time
    const int items = total / 2;
    const int clusters = items < sms / 2 ? items : sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(NF_THREADS); cfg.dynamicSmemBytes = NF_SMEM_BYTES; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernels[epi], ma0, ma1, mw, mwlo, mx0, mx1, t, q);
    const int rc = check_launch("conv2d_nf(cta pair)");
    g_prev_kernel = false;
    return rc;
  }
  const int grid = t.sk ? sk_grid : (total < sms ? total : sms);
  launch_pdl(kernels[epi], dim3(grid), dim3(NF_THREADS), NF_SMEM_BYTES, s, ma0, ma1, mw, mwlo, mx0, mx1, t, q);
  return check_launch("conv2d_nf");
}

}  // namespace frido
