// Streaming-softmax single-head attention on the tcgen05 tensor cores (sm_100a), BF16x3 error-compensated:
//
//   out = res + bias + softmax(scale * Q K^T) V          (attention.py:170-193 with the folded weights of unet.py)
//
// One launch replaces QK^T -> softmax -> PV and the [B,N,N] score tensor never exists.  The heads here are wide (one head of
// C = 384 .. 960 channels, N = 256 .. 4096 tokens), so nothing but the accumulators is resident:
//
//   CTA = (image b, 128-query tile, output-column slice of DV <= 384 channels); 320 threads:
//     warp 0      TMA producer: streams Q|K chunks (32 channels, bf16 hi|lo pairs) and V^T chunks (32 keys) through one
//                 ring of 32 KB slots, in exactly the order the MMA warp consumes them
//     warp 1      TMEM allocator + single-thread tcgen05.mma issuer:
//                   S[128 x 128] (TMEM cols 384..511) = Q_hi K_hi^T + Q_lo K_hi^T + Q_hi K_lo^T over the C/32 chunks
//                   O[128 x DV]  (TMEM cols 0..DV)   += P_hi V_hi + P_lo V_hi + P_hi V_lo
//                 issue order S(0) S(1) PV(0) S(2) PV(1) ... so the tensor pipe works on S(j+1) while the softmax warps
//                 turn S(j) into P(j)
//     warps 2-9   softmax + epilogue, two threads per query row (64 keys each): tcgen05.ld the scores, running row maximum
//                 (the accumulator is only rescaled when the maximum grows by more than 2^8 - exact, and rare), p = exp2,
//                 bf16 hi/lo split written to shared memory in the K-major SWIZZLE_128B operand layout, row sums in fp32;
//                 finally O / l + bias + residual -> global
//   Output slices (DV < C) recompute S; the launcher picks the split that minimises waves x work.
#include "tc_common.cuh"

#ifndef FRIDO_FLASH_PAIR_DEFAULT
#define FRIDO_FLASH_PAIR_DEFAULT 0
#endif

namespace frido {

constexpr int FA_SLOT = 32768;
constexpr int FA_SLOTS = 5;                          // ring depth when the 1024-B alignment pad leaves room, else one less
constexpr int FA_SLOT_PAIR = 24576;                  // CTA-pair kernel: a CTA holds only half of the K / V^T rows of a stage ...
constexpr int FA_SLOTS_PAIR = 6;                     // ... so six slots fit (always: 144 KB + 66 KB)
constexpr int FA_P_OFF = 65536 + 2048 + 256;         // P_hi (2 K-atoms x 16 KB) | P_lo, counted back from the END of the window
constexpr int FA_SMEM_BYTES = 232448;                // everything an SM has (227 KB)
constexpr int FA_S_COL = 384;
constexpr int FA_SM_WARPS = 8;
constexpr int FA_THREADS = 64 + 32 * FA_SM_WARPS;
constexpr float FA_RESCALE_LOG2 = 8.0f;

struct FaParams {
  int B, N, C, DV, DN, ND, nsplit, m_tiles, nkv;
  float sl2;                       // scale * log2(e)
  const float* bias;
  const float* res; long long r_sb, r_ld;
  float* out; long long o_sb, o_ld;
};

// PAIR: a cluster of two CTAs (tcgen05.mma.cta_group::2) takes two adjacent query tiles of one image.  Both tiles need the
// same K / V^T chunks, so each CTA loads HALF of a chunk's rows (64 of the 128 keys of an S stage, DN/2 of the DN channels of a
// PV stage) and the pair instruction reads both halves: the B-operand share of the shared-memory traffic that bounds the
// single-CTA kernel halves.  The leader CTA issues all MMAs; the peer's MMA warp only relays "my half has landed" to the leader,
// the peer's softmax warps arrive on the leader's barriers, and the leader's commits are multicast to both CTAs.
template <bool PAIR>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_flash_kernel(const __grid_constant__ CUtensorMap map_qh, const __grid_constant__ CUtensorMap map_ql,
                  const __grid_constant__ CUtensorMap map_kh, const __grid_constant__ CUtensorMap map_kl,
                  const __grid_constant__ CUtensorMap map_vh, const __grid_constant__ CUtensorMap map_vl, const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: ring (NS x 32 KB) | P_hi | P_lo | max exchange (2 KB) | barriers (128 B).  Five slots fit when the 1024-B
  // alignment pad of the window is <= 896 B; a worse-aligned window gets four.
  constexpr int SLOT = PAIR ? FA_SLOT_PAIR : FA_SLOT;
  constexpr int MAXS = PAIR ? FA_SLOTS_PAIR : FA_SLOTS;
  const int NS = PAIR ? FA_SLOTS_PAIR
                      : ((int)(smem_base - smem_u32(smem_raw)) + FA_SLOTS * FA_SLOT + FA_P_OFF <= FA_SMEM_BYTES ? FA_SLOTS : FA_SLOTS - 1);
  const uint32_t ring = smem_base;
  const int p_off = NS * SLOT, xch_off = p_off + 65536, bar_off = xch_off + 2048;
  const uint32_t bar_base = smem_base + bar_off;
  // barriers: full[MAXS] | empty[MAXS] | peer_full[MAXS] (leader: the peer's half of the stage has landed) | s_full s_empty p_full
  // p_empty o_full | tmem slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAXS + s); };
  auto peer_full = [&](int s) { return bar_base + 8u * (2 * MAXS + s); };
  const uint32_t s_full = bar_base + 8u * (3 * MAXS), s_empty = s_full + 8, p_full = s_full + 16, p_empty = s_full + 24,
                 o_full = s_full + 32, tmem_slot = s_full + 40;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8 * (3 * MAXS) + 40);
  const uint32_t rank = PAIR ? (blockIdx.x & 1u) : 0u;
  // remote (leader) copies of the barriers the softmax warps arrive on
  auto arrive_on_leader = [&](uint32_t bar) {
    if (PAIR) mbar_arrive_cluster(mapa_rank(bar, 0)); else mbar_arrive(bar);
  };

  const int warp = warp_idx_uniform();   // uniform role branches (see tc_common.cuh)
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  int x = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ds = x % p.nsplit; x /= p.nsplit;
  const int mtn = PAIR ? (p.m_tiles >> 1) : p.m_tiles;
  const int mt = PAIR ? 2 * (x % mtn) + (int)rank : x % mtn;
  const int b = x / mtn;
  const int m0 = mt * 128;
  const int dv0 = ds * p.DV;
  const int kchunks = p.C / 32;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_qh); prefetch_tmap(&map_ql); prefetch_tmap(&map_kh); prefetch_tmap(&map_kl);
    prefetch_tmap(&map_vh); prefetch_tmap(&map_vl);
    for (int s = 0; s < MAXS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(peer_full(s), 1); }
    mbar_init(s_full, 1);
    mbar_init(s_empty, (PAIR ? 2 : 1) * FA_SM_WARPS);   // leader's copy collects both CTAs' softmax warps
    mbar_init(p_full, (PAIR ? 2 : 1) * FA_SM_WARPS);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
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
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      auto load_s = [&](int j) {
        for (int c = 0; c < kchunks; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sl = ring + stage * SLOT;
          if (PAIR) {  // own Q rows, this CTA's 64 of the 128 keys
            mbar_expect_tx_elect(full_bar(stage), 2u * 8192u + 2u * 4096u);
            tma_load_3d_elect(sl, &map_qh, full_bar(stage), c * 32, m0, b);
            tma_load_3d_elect(sl + 8192, &map_ql, full_bar(stage), c * 32, m0, b);
            tma_load_3d_elect(sl + 16384, &map_kh, full_bar(stage), c * 32, j * 128 + (int)rank * 64, b);
            tma_load_3d_elect(sl + 20480, &map_kl, full_bar(stage), c * 32, j * 128 + (int)rank * 64, b);
          } else {
            mbar_expect_tx_elect(full_bar(stage), 4u * 8192u);
            tma_load_3d_elect(sl, &map_qh, full_bar(stage), c * 32, m0, b);
            tma_load_3d_elect(sl + 8192, &map_ql, full_bar(stage), c * 32, m0, b);
            tma_load_3d_elect(sl + 16384, &map_kh, full_bar(stage), c * 32, j * 128, b);
            tma_load_3d_elect(sl + 24576, &map_kl, full_bar(stage), c * 32, j * 128, b);
          }
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      };
      auto load_v = [&](int j) {
        for (int ks = 0; ks < 4; ++ks)
          for (int h = 0; h < p.ND; ++h) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sl = ring + stage * SLOT;
            if (PAIR) {  // this CTA's DN/2 of the DN channels
              const int hdn = p.DN >> 1;
              mbar_expect_tx_elect(full_bar(stage), 2u * (uint32_t)hdn * 64u);
              tma_load_3d_elect(sl, &map_vh, full_bar(stage), j * 128 + ks * 32, dv0 + h * p.DN + (int)rank * hdn, b);
              tma_load_3d_elect(sl + 8192, &map_vl, full_bar(stage), j * 128 + ks * 32, dv0 + h * p.DN + (int)rank * hdn, b);
            } else {
              mbar_expect_tx_elect(full_bar(stage), 2u * (uint32_t)p.DN * 64u);
              tma_load_3d_elect(sl, &map_vh, full_bar(stage), j * 128 + ks * 32, dv0 + h * p.DN, b);
              tma_load_3d_elect(sl + 16384, &map_vl, full_bar(stage), j * 128 + ks * 32, dv0 + h * p.DN, b);
            }
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
      };
      load_s(0);
      for (int j = 1; j < p.nkv; ++j) { load_s(j); load_v(j - 1); }
      load_v(p.nkv - 1);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (PAIR: leader CTA only; the peer's warp relays its `full` barriers) =====================
    if (rank == 0) {  // the whole warp runs the issue loop converged, one elected lane issues
      const uint32_t idesc_s = umma_idesc_bf16(PAIR ? 256 : 128, 128), idesc_o = umma_idesc_bf16(PAIR ? 256 : 128, p.DN);
      const uint32_t s_tmem = tmem_base + FA_S_COL;
      const uint32_t p_hi = smem_base + p_off, p_lo = p_hi + 32768;
      constexpr uint32_t KLO = PAIR ? 20480u : 24576u;   // offset of the K_lo rows inside an S stage
      constexpr uint32_t VLO = PAIR ? 8192u : 16384u;    // offset of the V_lo rows inside a PV stage
      int stage = 0;
      uint32_t phase = 0;
      auto mma = [&](uint32_t d, uint64_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
        if (PAIR) umma_bf16_2cta_elect(d, a, bdesc, idesc, acc); else umma_bf16_elect(d, a, bdesc, idesc, acc);
      };
      auto commit = [&](uint32_t bar) { if (PAIR) umma_commit_2cta_elect(bar); else umma_commit_elect(bar); };
      auto wait_stage = [&]() {
        mbar_wait(full_bar(stage), phase);
        if (PAIR) mbar_wait(peer_full(stage), phase);
        tc_fence_after();
      };
      auto issue_s = [&](int j) {
        if (j > 0) { mbar_wait(s_empty, (uint32_t)((j - 1) & 1)); tc_fence_after(); }  // the softmax warps hold S(j-1) in registers
        for (int c = 0; c < kchunks; ++c) {
          wait_stage();
          const uint32_t sl = ring + stage * SLOT;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t ah = umma_desc_sw64(sl + k * 32), al = umma_desc_sw64(sl + 8192 + k * 32);
            const uint64_t bh = umma_desc_sw64(sl + 16384 + k * 32), bl = umma_desc_sw64(sl + KLO + k * 32);
            mma(s_tmem, ah, bh, idesc_s, (c | k) ? 1u : 0u);
            mma(s_tmem, al, bh, idesc_s, 1u);
            mma(s_tmem, ah, bl, idesc_s, 1u);
          }
          commit(empty_bar(stage));
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        commit(s_full);
      };
      auto issue_pv = [&](int j) {
        mbar_wait(p_full, (uint32_t)(j & 1));
        tc_fence_after();
        for (int ks = 0; ks < 4; ++ks)
          for (int h = 0; h < p.ND; ++h) {
            wait_stage();
            const uint32_t sl = ring + stage * SLOT;
            const uint32_t d = tmem_base + (uint32_t)(h * p.DN);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int k16 = ks * 2 + k;
              const uint32_t poff = (uint32_t)((k16 >> 2) * 16384 + (k16 & 3) * 32);
              const uint64_t ph = umma_desc_sw128(p_hi + poff), pl = umma_desc_sw128(p_lo + poff);
              const uint64_t vh = umma_desc_sw64(sl + k * 32), vl = umma_desc_sw64(sl + VLO + k * 32);
              mma(d, ph, vh, idesc_o, (j | k16) ? 1u : 0u);
              mma(d, pl, vh, idesc_o, 1u);
              mma(d, ph, vl, idesc_o, 1u);
            }
            commit(empty_bar(stage));
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        commit(p_empty);  // P(j) consumed, O holds blocks 0..j
      };
      issue_s(0);
      for (int j = 1; j < p.nkv; ++j) { issue_s(j); issue_pv(j - 1); }
      issue_pv(p.nkv - 1);
      commit(o_full);
    } else if (PAIR && lane == 0) {
      // peer CTA: tell the leader when this CTA's half of each stage has landed (same stage sequence as the MMA issuer)
      int stage = 0;
      uint32_t phase = 0;
      const int total = p.nkv * (kchunks + 4 * p.ND);
      for (int i = 0; i < total; ++i) {
        mbar_wait(full_bar(stage), phase);
        mbar_arrive_cluster(mapa_rank(peer_full(stage), 0));
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== softmax + epilogue (warps 2..9) =====================
    const int q = warp & 3;               // TMEM lane quarter
    const int half = (warp - 2) >> 2;     // which 64 keys of the block / which half of the output columns
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* xch = reinterpret_cast<float*>(smem + xch_off);
    uint8_t* prow_hi = smem + p_off + half * 16384 + row * 128;
    uint8_t* prow_lo = prow_hi + 32768;
    const int ocols = p.DV >> 1;          // output columns per thread
    const int oc0 = half * ocols;
    float m_ref = 0.f, l = 0.f;
    for (int j = 0; j < p.nkv; ++j) {
      mbar_wait(s_full, (uint32_t)(j & 1));
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld32(lane_base + FA_S_COL + half * 64, s0);
      tmem_ld32(lane_base + FA_S_COL + half * 64 + 32, s1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_on_leader(s_empty);
      float mb = __uint_as_float(s0[0]);
#pragma unroll
      for (int i = 1; i < 32; ++i) mb = fmaxf(mb, __uint_as_float(s0[i]));
#pragma unroll
      for (int i = 0; i < 32; ++i) mb = fmaxf(mb, __uint_as_float(s1[i]));
      float* xb = xch + (j & 1) * 256;
      xb[half * 128 + row] = mb;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mb = fmaxf(mb, xb[(half ^ 1) * 128 + row]);
      float factor = 1.0f;
      bool need = false;
      if (j == 0) {
        m_ref = mb;
      } else if ((mb - m_ref) * p.sl2 > FA_RESCALE_LOG2) {
        need = true;
        factor = exp2f((m_ref - mb) * p.sl2);
        l *= factor;
        m_ref = mb;
      }
      const bool any = __any_sync(0xffffffffu, need);
      // p = exp2((s - m_ref) * scale * log2 e), in place, while PV(j-1) is still running
      const float mneg = -m_ref * p.sl2;
      float lsum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e0 = exp2f(fmaf(__uint_as_float(s0[i]), p.sl2, mneg)), e1 = exp2f(fmaf(__uint_as_float(s1[i]), p.sl2, mneg));
        lsum += e0 + e1;
        s0[i] = __float_as_uint(e0); s1[i] = __float_as_uint(e1);
      }
      if (j > 0) { mbar_wait(p_empty, (uint32_t)((j - 1) & 1)); tc_fence_after(); }  // PV(j-1) retired: P is free, O is stable
      if (any) {
        for (int c = 0; c < ocols; c += 16) {
          uint32_t r[16];
          tmem_ld16(lane_base + oc0 + c, r);
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * factor);
          tmem_st16(lane_base + oc0 + c, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t h[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = __uint_as_float((c < 4) ? s0[c * 8 + 2 * e] : s1[(c - 4) * 8 + 2 * e]);
          const float b2 = __uint_as_float((c < 4) ? s0[c * 8 + 2 * e + 1] : s1[(c - 4) * 8 + 2 * e + 1]);
          h[e] = pack_bf16x2(a, b2);
          lo[e] = pack_bf16x2(a - bf16_lo_to_f32(h[e]), b2 - bf16_hi_to_f32(h[e]));
        }
        const int off = (c ^ (row & 7)) << 4;
        *reinterpret_cast<uint4*>(prow_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(prow_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      l += lsum;
      fence_proxy_async();  // generic-proxy writes of P -> visible to the tensor core
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_on_leader(p_full);
    }
    // ---- epilogue: out = res + bias + O / l
    {
      float* xb = xch + (p.nkv & 1) * 256;
      xb[half * 128 + row] = l;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      l += xb[(half ^ 1) * 128 + row];
    }
    const float inv = 1.0f / l;
    mbar_wait(o_full, 0u);
    tc_fence_after();
    const long long go = (long long)b * p.o_sb + (long long)(m0 + row) * p.o_ld + dv0 + oc0;
    const long long gr = (long long)b * p.r_sb + (long long)(m0 + row) * p.r_ld + dv0 + oc0;
    for (int c = 0; c < ocols; c += 16) {
      uint32_t r[16];
      float4 rr[4];
      if (p.res) {
#pragma unroll
        for (int k = 0; k < 4; ++k) rr[k] = *reinterpret_cast<const float4*>(p.res + gr + c + 4 * k);
      }
      tmem_ld16(lane_base + oc0 + c, r);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 v = make_float4(__uint_as_float(r[4 * k]) * inv, __uint_as_float(r[4 * k + 1]) * inv,
                               __uint_as_float(r[4 * k + 2]) * inv, __uint_as_float(r[4 * k + 3]) * inv);
        if (p.bias) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + dv0 + oc0 + c + 4 * k));
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        if (p.res) { v.x += rr[k].x; v.y += rr[k].y; v.z += rr[k].z; v.w += rr[k].w; }
        *reinterpret_cast<float4*>(p.out + go + c + 4 * k) = v;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace frido

using namespace frido;

extern "C" int frido_attn_flash_eligible(int32_t B, int32_t N, int32_t C) {
  return B >= 1 && N >= 128 && N % 128 == 0 && C >= 64 && C % 32 == 0 && C <= 384 * 8;
}

extern "C" int frido_attn_flash(const FridoFlashParams* p, void* stream) {
  if (!p || !p->q_hi || !p->q_lo || !p->k_hi || !p->k_lo || !p->vt_hi || !p->vt_lo || !p->out)
    return set_error(FRIDO_E_ARG, "attn_flash: null pointer");
  if (!frido_attn_flash_eligible(p->B, p->N, p->C))
    return set_error(FRIDO_E_ARG, "attn_flash: needs N % 128 == 0 and C % 32 == 0");
  if (p->q_ld % 8 || p->q_sb % 8 || p->k_ld % 8 || p->k_sb % 8 || p->vt_ld % 8 || p->vt_sb % 8 || !a16(p->q_hi) || !a16(p->q_lo) ||
      !a16(p->k_hi) || !a16(p->k_lo) || !a16(p->vt_hi) || !a16(p->vt_lo))
    return set_error(FRIDO_E_ARG, "attn_flash: bf16 operand strides / pointers must be multiples of 16 bytes");
  if (!a16(p->out) || p->o_ld % 4 || p->o_sb % 4 || (p->res && (!a16(p->res) || p->r_ld % 4 || p->r_sb % 4)) || (p->bias && !a16(p->bias)))
    return set_error(FRIDO_E_ARG, "attn_flash: fp32 rows must be 16-byte aligned");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  FaParams t;
  t.B = p->B; t.N = p->N; t.C = p->C;
  t.m_tiles = p->N / 128; t.nkv = p->N / 128;
  // output-column split: every slice recomputes S (cost ~ C per key block) and does its share of PV (cost ~ DV)
  int best_n = 0;
  double best_cost = 1e30;
  for (int n = 1; n <= 16; ++n) {
    if (p->C % (32 * n)) continue;
    const int dv = p->C / n;
    if (dv > 384) continue;
    const long long ctas = (long long)p->B * t.m_tiles * n;
    const long long waves = (ctas + sms - 1) / sms;
    const double cost = (double)waves * ((double)t.nkv * (p->C + dv) * 12.0 + 6000.0);
    if (cost < best_cost) { best_cost = cost; best_n = n; }
  }
  if (const char* e = getenv("FRIDO_FLASH_SPLIT")) {  // tests: force a split
    const int n = atoi(e);
    if (n >= 1 && p->C % (32 * n) == 0 && p->C / n <= 384) best_n = n;
  }
  if (!best_n) return set_error(FRIDO_E_ARG, "attn_flash: no legal output split");
  t.nsplit = best_n;
  t.DV = p->C / best_n;
  t.ND = t.DV > 256 ? 2 : 1;
  t.DN = t.DV / t.ND;
  t.sl2 = p->scale * 1.4426950408889634f;
  t.bias = p->bias; t.res = p->res; t.r_sb = p->r_sb; t.r_ld = p->r_ld;
  t.out = p->out; t.o_sb = p->o_sb; t.o_ld = p->o_ld;
  // CTA-pair kernel (cta_group::2): two adjacent query tiles per cluster.  FRIDO_FLASH_PAIR = 1 (default) | 0
  bool pair = t.m_tiles % 2 == 0 && t.DN % 16 == 0;
  if (const char* e = getenv("FRIDO_FLASH_PAIR")) pair = pair && atoi(e) != 0;
  else pair = pair && FRIDO_FLASH_PAIR_DEFAULT;
  const uint32_t krows = pair ? 64u : 128u, vrows = pair ? (uint32_t)t.DN / 2 : (uint32_t)t.DN;
  CUtensorMap mqh, mql, mkh, mkl, mvh, mvl;
  if (!make_map3(&mqh, p->q_hi, p->C, p->N, p->B, p->q_ld, p->q_sb, 128, true) ||
      !make_map3(&mql, p->q_lo, p->C, p->N, p->B, p->q_ld, p->q_sb, 128, true) ||
      !make_map3(&mkh, p->k_hi, p->C, p->N, p->B, p->k_ld, p->k_sb, krows, true) ||
      !make_map3(&mkl, p->k_lo, p->C, p->N, p->B, p->k_ld, p->k_sb, krows, true) ||
      !make_map3(&mvh, p->vt_hi, p->N, p->C, p->B, p->vt_ld, p->vt_sb, vrows, true) ||
      !make_map3(&mvl, p->vt_lo, p->N, p->C, p->B, p->vt_ld, p->vt_sb, vrows, true))
    return set_error(FRIDO_E_ARG, "attn_flash: cuTensorMapEncodeTiled failed");
  static DevOnce attr;
  if (attr.need()) {
    if (cudaFuncSetAttribute(attn_flash_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(attn_flash_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES) != cudaSuccess)
      return set_error(FRIDO_E_LAUNCH, "attn_flash: cannot opt in to dynamic shared memory");
  }
  const int grid = p->B * t.m_tiles * t.nsplit;
  if (pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(FA_THREADS); cfg.dynamicSmemBytes = FA_SMEM_BYTES; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, attn_flash_kernel<true>, mqh, mql, mkh, mkl, mvh, mvl, t);
    const int rc = check_launch("attn_flash(cta pair)");
    g_prev_kernel = false;
    return rc;
  }
  launch_pdl(attn_flash_kernel<false>, dim3(grid), dim3(FA_THREADS), FA_SMEM_BYTES, (cudaStream_t)stream, mqh, mql, mkh, mkl, mvh, mvl, t);
  return check_launch("attn_flash");
}
