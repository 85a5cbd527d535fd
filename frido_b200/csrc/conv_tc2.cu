// CTA-pair variant of the BF16x3 tcgen05 implicit-GEMM engine (conv_tc.cu MODE 2): `tcgen05.mma.cta_group::2`.
//
// A cluster of two CTAs (one TPC) takes the M tiles (2q, 2q+1) of one N tile:  D[256 x BN] += A[256 x 32] W[BN x 32]^T  with
//   * each CTA loading and splitting ITS OWN 128-row A tile (TMA -> shared memory -> splitter warps -> tensor-memory slots,
//     exactly as in conv_tc.cu), and
//   * each CTA loading only HALF of the W tile (BN/2 rows, hi and lo): the pair instruction reads both halves, so the
//     W share of the shared-memory traffic that bounds conv_tc.cu (TMA fill + MMA operand reads) halves per SM.
// The leader CTA (cluster rank 0) issues every MMA; it waits for BOTH CTAs' splitter warps (the peer's arrive remotely on the
// leader's barrier), and its commits are multicast to the ring / accumulator barriers of both CTAs.  Both epilogues arrive on
// the leader's accumulator-empty barrier.  Data-parallel schedule only (no stream-K), specialised epilogues shared with
// conv_tc.cu through tc_common.cuh.
#include "tc_common.cuh"

namespace frido {

constexpr int TC2_THREADS = TC_THREADS_X3 + 32;   // + warp 14: the W tiles' own TMA issuer (see conv_tc.cu)

template <int EPI>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                    const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_wlo,
                    const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int HBN = p.BN >> 1;                               // W rows this CTA holds
  const uint32_t b_bytes = (uint32_t)HBN * TC_BK * 2;      // one half tile, bf16
  const uint32_t stage_bytes = TC_A_BYTES + 2u * b_bytes;  // A_raw(fp32) | W_hi half | W_lo half
  const uint32_t off_w = (uint32_t)TC_A_BYTES, off_wlo = off_w + b_bytes;
  const uint32_t acc_stride = TC_BF_ACC_STRIDE;
  const uint32_t bar_base = smem_base + TC_SMEM_BUDGET + TC_STG_BYTES + TC_CSUM_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
  auto split_bar = [&](int s) { return bar_base + 8u * (2 * TC_MAX_STAGES + s); };
  auto tfree_bar = [&](int t) { return bar_base + 8u * (TC_BAR_SK_FLAG + 1 + t); };  // tensor-memory A slot t is free again
  auto tfull_bar = [&](int a) { return bar_base + 8u * (TC_BAR_TFULL + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (TC_BAR_TEMPTY + a); };
  const uint32_t tmem_slot = bar_base + 8u * TC_BAR_TMEM_SLOT;
  const int NS = p.stages;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = warp_idx_uniform();   // uniform role branches (see tc_common.cuh)
  const int lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1;  // cluster dims (2,1,1): rank in the pair
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
    for (int s = 0; s < NS; ++s) {
      mbar_init(full_bar(s), 2);                       // the A issuer and the W issuer each arrive with their byte counts
      mbar_init(empty_bar(s), 1);                      // one multicast commit from the leader
    }
    for (int t = 0; t < TC_BF_MAX_STAGES; ++t) {
      mbar_init(split_bar(t), 2 * TC_SPLIT_WARPS);     // (leader's copy is the one in use) both CTAs' splitter warps
      mbar_init(tfree_bar(t), 1);                      // one multicast commit from the leader
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * TC_EPI_WARPS);      // (leader's copy) both CTAs' epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anyone arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own A tile, own half of W) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      SegIter it(p, ksteps, total_tiles);
      int tile, k0, k1;
      while (it.next(tile, k0, k1)) {
        int mt = tile / p.tiles_n;
        const int tx = mt % p.tiles_x; mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tb = mt / p.tiles_y;
        const int ox0 = tx * p.TW, oy0 = ty * p.TH, b0 = tb * p.TB;
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * stage_bytes;
          mbar_expect_tx(full_bar(stage), (uint32_t)TC_A_BYTES);
          if (ks < ksteps_main) {
            const int tap = ks / kchunks;
            const int kc = ks - tap * kchunks;
            const int dy = tap / p.ksize, dx = tap - dy * p.ksize;
            const int ch = kc * TC_BK;
            const int cx = ox0 * p.stride + dx - p.pad, cy = oy0 * p.stride + dy - p.pad;
            if (ch < p.c0) tma_load_4d(sa, &map_a0, full_bar(stage), ch, cx, cy, b0);
            else           tma_load_4d(sa, &map_a1, full_bar(stage), ch - p.c0, cx, cy, b0);
          } else {
            const int ch = (ks - ksteps_main) * TC_BK;
            if (ch < p.cx0) tma_load_4d(sa, &map_x0, full_bar(stage), ch, ox0 * p.stride, oy0 * p.stride, b0);
            else            tma_load_4d(sa, &map_x1, full_bar(stage), ch - p.cx0, ox0 * p.stride, oy0 * p.stride, b0);
          }
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {  // the whole warp runs the issue loop converged, one elected lane issues
      const uint32_t idesc = umma_idesc_bf16(2 * TC_BM, p.BN);
      int stage = 0, slot = 0;
      uint32_t phase = 0, sphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      SegIter it(p, ksteps, total_tiles);
      int tile, k0, k1;
      while (it.next(tile, k0, k1)) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(split_bar(slot), sphase);       // both CTAs: A pair in tensor-memory slot `slot` (hence W halves landed too)
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint32_t ah = tmem_base + (uint32_t)(TC_BF_A_COL + slot * 32 + k * 8), al = ah + 16;
            const uint64_t bh = umma_desc_sw64(sa + off_w + k * 32), bl = umma_desc_sw64(sa + off_wlo + k * 32);
            umma_bf16_ts_2cta_elect(d_tmem, ah, bh, idesc, ((ks - k0) | k) ? 1u : 0u);
            umma_bf16_ts_2cta_elect(d_tmem, al, bh, idesc, 1u);
            umma_bf16_ts_2cta_elect(d_tmem, ah, bl, idesc, 1u);
          }
          umma_commit_2cta_elect(empty_bar(stage));  // frees the shared-memory ring slot in both CTAs
          umma_commit_2cta_elect(tfree_bar(slot));   // ... and the tensor-memory operand slot
          if (++stage == NS) { stage = 0; phase ^= 1; }
          if (++slot == TC_BF_MAX_STAGES) { slot = 0; sphase ^= 1; }
        }
        umma_commit_2cta_elect(tfull_bar(acc));      // accumulator halves ready for both epilogues
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 2 + TC_EPI_WARPS) {
    tc_epilogue_role<EPI>(p, smem_raw, smem_base, bar_base, tmem_base, acc_stride, ksteps, total_tiles);
  } else if (warp == 14) {
    // ===================== W-tile TMA issuer: this CTA's half of the hi and lo tiles of every k-step =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      SegIter it(p, ksteps, total_tiles);
      int tile, k0, k1;
      while (it.next(tile, k0, k1)) {
        const int n0 = (tile % p.tiles_n) * p.BN + (int)rank * HBN;
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * stage_bytes;
          mbar_expect_tx(full_bar(stage), 2u * b_bytes);
          if (p.dbg_w) {  // timing experiment only: the same bytes as two contiguous bulk copies (one request each)
            const char* src = reinterpret_cast<const char*>(p.dbg_w) + (size_t)((ks * 2 + (int)rank) & 63) * 2 * b_bytes;
            bulk_load_1d(sa + off_w, src, b_bytes, full_bar(stage));
            bulk_load_1d(sa + off_wlo, src + b_bytes, b_bytes, full_bar(stage));
          } else {
            tma_load_3d(sa + off_w, &map_w, full_bar(stage), ks * TC_BK, n0, 0);
            tma_load_3d(sa + off_wlo, &map_wlo, full_bar(stage), ks * TC_BK, n0, 0);
          }
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== splitter (warps 10..13): fp32 A tile -> bf16 hi / lo halves in this CTA's tensor memory ==========
    const int r = (warp & 3) * 32 + lane;
    const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)TC_BF_A_COL;
    int stage = 0, slot = 0;
    uint32_t phase = 0, sphase = 0;
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
        mbar_wait(tfree_bar(slot), sphase ^ 1);  // the MMAs that last read this tensor-memory slot have retired
        tc_fence_after();
        tmem_st16(a_lane + (uint32_t)(slot * 32), hi);
        tmem_st16(a_lane + (uint32_t)(slot * 32 + 16), lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_rank(split_bar(slot), 0));  // the leader's barrier collects both CTAs
        if (++stage == NS) { stage = 0; phase ^= 1; }
        if (++slot == TC_BF_MAX_STAGES) { slot = 0; sphase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA frees tensor memory or exits while the pair's MMAs / remote arrives may still touch it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Launch (called by conv2d_tc once it has chosen the tile width and built the A / side maps).  `mw`, `mwlo`: W maps whose box
// is BN/2 rows.  grid = 2 * clusters.
int conv2d_tc_pair_launch(const TcParams& t, int epi, int clusters, const CUtensorMap& ma0, const CUtensorMap& ma1,
                          const CUtensorMap& mw, const CUtensorMap& mwlo, const CUtensorMap& mx0, const CUtensorMap& mx1,
                          cudaStream_t s) {
  using KernelFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, TcParams);
  static const KernelFn kernels[EPI_COUNT] = {conv_tc_pair_kernel<EPI_GENERIC>, conv_tc_pair_kernel<EPI_BIAS>, conv_tc_pair_kernel<EPI_BIAS_RES>,
                                             conv_tc_pair_kernel<EPI_BIAS_RV_CS>, conv_tc_pair_kernel<EPI_BIAS_RES_CS>,
                                             conv_tc_pair_kernel<EPI_BIAS_GEGLU>, conv_tc_pair_kernel<EPI_BIAS_CS>,
                                             conv_tc_pair_kernel<EPI_BIAS_PAIR>};
  static DevOnce attr;
  if (attr.need()) {
    for (int i = 0; i < EPI_COUNT; ++i)
      if (cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess)
        return set_error(FRIDO_E_LAUNCH, "conv2d_tc(pair): cannot opt in to dynamic shared memory");
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(TC2_THREADS); cfg.dynamicSmemBytes = TC_SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernels[epi], ma0, ma1, mw, mwlo, mx0, mx1, t);
  const int rc = check_launch("conv2d_tc(bf16x3, cta pair)");
  g_prev_kernel = false;  // launched without the programmatic-serialisation attribute: the next kernel takes a full dependency
  return rc;
}

}  // namespace frido
