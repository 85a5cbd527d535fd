// SIMT fp32 implicit-GEMM engine: every conv / linear / batched matmul shape the
// path has, exact fp32 FFMA accumulation.  It carries the odd shapes the tcgen05
// engine does not take (C_in = 3/6, C_out = 3, K tails, NCHW latents in/out)
// and is the numerical yardstick for the tensor-core engine.
//
// out[b,p,n] = act(alpha * sum_k A(b,p,k) W[n,k] + bias[n] + rowvec[b,n] + res[b,p,n])
// k = tap * Cin + c ; tap = dy*ksize+dx ; A gathers from up to two channel-
// concatenated NHWC (or arbitrarily strided) sources.
#include "common.cuh"

namespace frido {

constexpr int BM = 64, BN = 64, BK = 32, NT = 256;

struct ALoad {
  float v[4];
};

template <bool VEC_A, bool VEC_W>
__global__ void __launch_bounds__(NT) conv_simt_kernel(const FridoConvParams p) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Ws[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int HWout = p.Hout * p.Wout;
  const int Cin = p.c0 + p.c1;
  const int taps = p.ksize * p.ksize;
  const int Ktot = taps * Cin;

  // loader role: one (row, 2 x 4 consecutive k) per thread for A and for W
  const int lrow = tid >> 2;        // 0..63
  const int lk = (tid & 3) * 4;     // 0,4,8,12 (+16 for the second quad)
  const int pm = m0 + lrow;
  const bool m_ok = pm < HWout;
  const int oy = m_ok ? pm / p.Wout : 0;
  const int ox = m_ok ? pm - oy * p.Wout : 0;
  const int Hl = p.Hin * p.ups, Wl = p.Win * p.ups;  // logical (post-upsample) input grid
  const int ush = p.ups == 2 ? 1 : 0;
  const float* __restrict__ a0 = p.a0 + (int64_t)b * p.a0_sb;
  const float* __restrict__ a1 = p.a1 ? p.a1 + (int64_t)b * p.a1_sb : nullptr;
  const int wn = n0 + lrow;
  const bool n_ok = wn < p.Cout;
  const float* __restrict__ wrow = p.w + (int64_t)b * p.w_sb + (int64_t)(n_ok ? wn : 0) * (p.w_ld ? p.w_ld : (int64_t)Ktot);

  auto load_a = [&](int k0, float (&v)[4]) {
    const int k = k0 + lk;
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (!m_ok || k >= Ktot) return;
    if (VEC_A) {
      const int tap = k / Cin;
      const int c = k - tap * Cin;
      const int dy = tap / p.ksize, dx = tap - dy * p.ksize;
      const int iy = oy * p.stride + dy - p.pad, ix = ox * p.stride + dx - p.pad;
      if (iy < 0 || iy >= Hl || ix < 0 || ix >= Wl) return;
      const int sy = iy >> ush, sx = ix >> ush;
      const float* src = (c < p.c0) ? a0 + (int64_t)sy * p.a0_sy + (int64_t)sx * p.a0_sx + c
                                    : a1 + (int64_t)sy * p.a1_sy + (int64_t)sx * p.a1_sx + (c - p.c0);
      const float4 t = __ldg(reinterpret_cast<const float4*>(src));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k + j;
        if (kk >= Ktot) break;
        const int tap = kk / Cin;
        const int c = kk - tap * Cin;
        const int dy = tap / p.ksize, dx = tap - dy * p.ksize;
        const int iy = oy * p.stride + dy - p.pad, ix = ox * p.stride + dx - p.pad;
        if (iy < 0 || iy >= Hl || ix < 0 || ix >= Wl) continue;
        const int sy = iy >> ush, sx = ix >> ush;
        v[j] = (c < p.c0) ? __ldg(a0 + (int64_t)sy * p.a0_sy + (int64_t)sx * p.a0_sx + (int64_t)c * p.a0_sc)
                          : __ldg(a1 + (int64_t)sy * p.a1_sy + (int64_t)sx * p.a1_sx + (int64_t)(c - p.c0) * p.a1_sc);
      }
    }
  };
  auto load_w = [&](int k0, float (&v)[4]) {
    const int k = k0 + lk;
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (!n_ok || k >= Ktot) return;
    if (VEC_W) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(wrow + k));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k + j < Ktot) v[j] = __ldg(wrow + k + j);
    }
  };

  // compute role: 4x4 micro-tile
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rw[4], ra2[4], rw2[4];
  load_a(0, ra);
  load_w(0, rw);
  load_a(16, ra2);
  load_w(16, rw2);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    As[0][lk + j][lrow] = ra[j];
    Ws[0][lk + j][lrow] = rw[j];
    As[0][16 + lk + j][lrow] = ra2[j];
    Ws[0][16 + lk + j][lrow] = rw2[j];
  }
  __syncthreads();

  const int nk = (Ktot + BK - 1) / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_a((kt + 1) * BK, ra);
      load_w((kt + 1) * BK, rw);
      load_a((kt + 1) * BK + 16, ra2);
      load_w((kt + 1) * BK + 16, rw2);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[cur][k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      const float w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        As[cur ^ 1][lk + j][lrow] = ra[j];
        Ws[cur ^ 1][lk + j][lrow] = rw[j];
        As[cur ^ 1][16 + lk + j][lrow] = ra2[j];
        Ws[cur ^ 1][16 + lk + j][lrow] = rw2[j];
      }
    }
    __syncthreads();
  }

  // epilogue
  float* __restrict__ out = p.out + (int64_t)b * p.o_sb;
  const float* __restrict__ res = p.res ? p.res + (int64_t)b * p.o_sb : nullptr;
  const float* __restrict__ rv = p.rowvec ? p.rowvec + (int64_t)b * p.rowvec_sb : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pmo = m0 + ty * 4 + i;
    if (pmo >= HWout) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      float t = acc[i][j] * p.alpha;
      if (n < p.Cout) {
        if (p.bias) t += __ldg(p.bias + n);
        if (rv) t += __ldg(rv + n);
      }
      v[j] = t;
    }
    if (p.act == FRIDO_ACT_GEGLU) {
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const int n = n0 + tx * 4 + j;
        if (n + 1 < p.Cout) {
          const int no = n >> 1;
          float t = v[j] * gelu_erf(v[j + 1]);
          const int64_t o = (int64_t)pmo * p.o_sp + (int64_t)no * p.o_sn;
          if (res) t += res[o];
          out[o] = p.round_tf32 ? round_tf32(t) : t;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < p.Cout) {
          float t = v[j];
          const int64_t o = (int64_t)pmo * p.o_sp + (int64_t)n * p.o_sn;
          if (res) t += res[o];
          if (p.act == FRIDO_ACT_RELU) t = fmaxf(t, 0.f);
          else if (p.act == FRIDO_ACT_SILU) t = silu_f(t);
          else if (p.act == FRIDO_ACT_GELU) t = gelu_erf(t);
          t = p.round_tf32 ? round_tf32(t) : t;
          out[o] = t;
          if (p.out_hi) {
            uint16_t hh, ll;
            split_bf16(t, hh, ll);
            const int64_t oo = (int64_t)b * p.o_sb + o;
            ((uint16_t*)p.out_hi)[oo] = hh; ((uint16_t*)p.out_lo)[oo] = ll;
          }
        }
      }
    }
  }
}

// Few output channels (UNet / decoder heads: C_out = 3): one thread per output pixel, all C_out accumulators in
// registers, weights in shared memory.  The generic 64x64 tile would waste 61 of 64 columns here.
constexpr int SC_MAXCOUT = 4;
__global__ void __launch_bounds__(128) conv_smallcout_kernel(const FridoConvParams p) {
  extern __shared__ float wsm[];  // [Cout][Ktot]
  const int Cin = p.c0 + p.c1;
  const int Ktot = p.ksize * p.ksize * Cin;
  const int b = blockIdx.y;
  const int64_t wld = p.w_ld ? p.w_ld : Ktot;
  for (int i = threadIdx.x; i < p.Cout * Ktot; i += blockDim.x) wsm[i] = p.w[(int64_t)b * p.w_sb + (int64_t)(i / Ktot) * wld + (i % Ktot)];
  __syncthreads();
  const int HWout = p.Hout * p.Wout;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HWout) return;
  const int oy = pix / p.Wout, ox = pix - oy * p.Wout;
  const int Hl = p.Hin * p.ups, Wl = p.Win * p.ups, ush = p.ups == 2 ? 1 : 0;
  const float* a0 = p.a0 + (int64_t)b * p.a0_sb;
  const float* a1 = p.a1 ? p.a1 + (int64_t)b * p.a1_sb : nullptr;
  float acc[SC_MAXCOUT] = {0.f, 0.f, 0.f, 0.f};
  for (int tap = 0; tap < p.ksize * p.ksize; ++tap) {
    const int dy = tap / p.ksize, dx = tap - dy * p.ksize;
    const int iy = oy * p.stride + dy - p.pad, ix = ox * p.stride + dx - p.pad;
    if (iy < 0 || iy >= Hl || ix < 0 || ix >= Wl) continue;
    const int sy = iy >> ush, sx = ix >> ush;
    const float4* s0 = reinterpret_cast<const float4*>(a0 + (int64_t)sy * p.a0_sy + (int64_t)sx * p.a0_sx);
    const float* wt = wsm + tap * Cin;
    for (int c4 = 0; c4 < p.c0 / 4; ++c4) {
      const float4 v = __ldg(s0 + c4);
#pragma unroll
      for (int n = 0; n < SC_MAXCOUT; ++n)
        if (n < p.Cout) {
          const float4 w = *reinterpret_cast<const float4*>(wt + n * Ktot + 4 * c4);
          acc[n] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[n]))));
        }
    }
    if (a1) {
      const float4* s1 = reinterpret_cast<const float4*>(a1 + (int64_t)sy * p.a1_sy + (int64_t)sx * p.a1_sx);
      for (int c4 = 0; c4 < p.c1 / 4; ++c4) {
        const float4 v = __ldg(s1 + c4);
#pragma unroll
        for (int n = 0; n < SC_MAXCOUT; ++n)
          if (n < p.Cout) {
            const float4 w = *reinterpret_cast<const float4*>(wt + n * Ktot + p.c0 + 4 * c4);
            acc[n] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[n]))));
          }
      }
    }
  }
  float* out = p.out + (int64_t)b * p.o_sb + (int64_t)pix * p.o_sp;
  const float* res = p.res ? p.res + (int64_t)b * p.o_sb + (int64_t)pix * p.o_sp : nullptr;
#pragma unroll
  for (int n = 0; n < SC_MAXCOUT; ++n)
    if (n < p.Cout) {
      float t = acc[n] * p.alpha;
      if (p.bias) t += __ldg(p.bias + n);
      if (p.rowvec) t += __ldg(p.rowvec + (int64_t)b * p.rowvec_sb + n);
      if (res) t += res[(int64_t)n * p.o_sn];
      if (p.act == FRIDO_ACT_RELU) t = fmaxf(t, 0.f);
      else if (p.act == FRIDO_ACT_SILU) t = silu_f(t);
      else if (p.act == FRIDO_ACT_GELU) t = gelu_erf(t);
      t = p.round_tf32 ? round_tf32(t) : t;
      out[(int64_t)n * p.o_sn] = t;
      if (p.out_u8) p.out_u8[((int64_t)b * HWout + pix) * p.Cout + n] = format_u8(t, p.u8_mode);
    }
}

// Same job for the common case (3x3, stride 1, one dense NHWC source: the UNet and decoder output heads), tiled: a CTA owns
// 8 x 16 output pixels, stages their 10 x 18 input halo tile in shared memory with cp.async (zero fill outside the
// image) and then every thread reads its 9 x Cin window from there instead of re-fetching it through L1 with 768-byte
// lane strides.  Pixel stride in shared memory is Cin + 4 floats, so the float4 reads of a warp are conflict-free.
// Four threads share a pixel (a quarter of the input channels each, partial sums combined through shared memory): with
// 160 KB of staging only one CTA fits an SM, so the 512 threads are what hides the shared-memory latency.
constexpr int SCT_TH = 8, SCT_TW = 16, SCT_PARTS = 4;
__global__ void __launch_bounds__(SCT_TH * SCT_TW * SCT_PARTS) conv_smallcout_tiled_kernel(const FridoConvParams p) {
  extern __shared__ float4 sct_sm[];
  const int Cin = p.c0;
  const int Q = Cin >> 2, QS = Q + 1;       // quads per pixel, padded stride
  const int Ktot = 9 * Cin;
  float4* xs = sct_sm;                                                  // [(TH+2)*(TW+2)][QS]
  float* wsm = reinterpret_cast<float*>(sct_sm + (SCT_TH + 2) * (SCT_TW + 2) * QS);  // [Cout][Ktot]
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * SCT_TH, ox0 = blockIdx.x * SCT_TW;
  const float* a0 = p.a0 + (int64_t)b * p.a0_sb;
  const int64_t wld = p.w_ld ? p.w_ld : Ktot;
  for (int i = threadIdx.x; i < (SCT_TH + 2) * (SCT_TW + 2) * Q; i += blockDim.x) {
    const int pix = i / Q, quad = i - pix * Q;
    const int ly = pix / (SCT_TW + 2), lx = pix - ly * (SCT_TW + 2);
    const int iy = oy0 + ly - 1, ix = ox0 + lx - 1;
    const bool in = iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
    const float* src = a0 + (in ? (int64_t)iy * p.a0_sy + (int64_t)ix * p.a0_sx + 4 * quad : 0);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(xs + pix * QS + quad);
    const int nbytes = in ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled (conv padding)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
  }
  for (int i = threadIdx.x; i < p.Cout * Ktot; i += blockDim.x) wsm[i] = p.w[(int64_t)b * p.w_sb + (int64_t)(i / Ktot) * wld + (i % Ktot)];
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int part = threadIdx.x / (SCT_TH * SCT_TW), tp = threadIdx.x - part * (SCT_TH * SCT_TW);
  const int ty = tp / SCT_TW, tx = tp - ty * SCT_TW;
  const int oy = oy0 + ty, ox = ox0 + tx;
  const int c_lo = part * Q / SCT_PARTS, c_hi = (part + 1) * Q / SCT_PARTS;
  float acc[SC_MAXCOUT] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap - dy * 3;
    const float4* xr = xs + ((ty + dy) * (SCT_TW + 2) + tx + dx) * QS;
    const float* wt = wsm + tap * Cin;
    for (int c4 = c_lo; c4 < c_hi; ++c4) {
      const float4 v = xr[c4];
#pragma unroll
      for (int n = 0; n < SC_MAXCOUT; ++n)
        if (n < p.Cout) {
          const float4 w = *reinterpret_cast<const float4*>(wt + n * Ktot + 4 * c4);
          acc[n] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[n]))));
        }
    }
  }
  // combine the channel quarters in a fixed order (the staged input tile is dead by now: reuse its memory)
  __syncthreads();
  float* red = reinterpret_cast<float*>(xs);  // [PARTS][SC_MAXCOUT][TH*TW]
#pragma unroll
  for (int n = 0; n < SC_MAXCOUT; ++n) red[(part * SC_MAXCOUT + n) * (SCT_TH * SCT_TW) + tp] = acc[n];
  __syncthreads();
  if (part != 0 || oy >= p.Hout || ox >= p.Wout) return;
  const int pix = oy * p.Wout + ox;
  float* out = p.out + (int64_t)b * p.o_sb + (int64_t)pix * p.o_sp;
  const float* res = p.res ? p.res + (int64_t)b * p.o_sb + (int64_t)pix * p.o_sp : nullptr;
#pragma unroll
  for (int n = 0; n < SC_MAXCOUT; ++n)
    if (n < p.Cout) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < SCT_PARTS; ++q) t += red[(q * SC_MAXCOUT + n) * (SCT_TH * SCT_TW) + tp];
      t *= p.alpha;
      if (p.bias) t += __ldg(p.bias + n);
      if (p.rowvec) t += __ldg(p.rowvec + (int64_t)b * p.rowvec_sb + n);
      if (res) t += res[(int64_t)n * p.o_sn];
      if (p.act == FRIDO_ACT_RELU) t = fmaxf(t, 0.f);
      else if (p.act == FRIDO_ACT_SILU) t = silu_f(t);
      else if (p.act == FRIDO_ACT_GELU) t = gelu_erf(t);
      t = p.round_tf32 ? round_tf32(t) : t;
      out[(int64_t)n * p.o_sn] = t;
      // output formatting fused into the head (sample_diffusion.py:115-121): the byte comes from the very fp32 value stored
      if (p.out_u8) p.out_u8[((int64_t)b * p.Hout * p.Wout + pix) * p.Cout + n] = format_u8(t, p.u8_mode);
    }
}

// Linear layer with at most 16 rows (the timestep-embedding MLP and the 22 stacked ResBlock embedding projections,
// pyunet.py:561-565,225-231: M = batch): weight streaming is the whole cost, so one warp owns an output column, reads its
// weight row once (coalesced float4) and accumulates all rows against the input held in shared memory.
constexpr int SM_MAXROWS = 16, SM_COLS_PER_WARP = 4;
__global__ void __launch_bounds__(256) linear_smallm_kernel(const FridoConvParams p) {
  extern __shared__ float4 lsm_x[];  // [M][K/4]
  const int M = p.Wout, K = p.c0, Q = K >> 2;
  for (int i = threadIdx.x; i < M * Q; i += blockDim.x) {
    const int m = i / Q, q = i - m * Q;
    lsm_x[i] = __ldg(reinterpret_cast<const float4*>(p.a0 + (int64_t)m * p.a0_sx) + q);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t wld = p.w_ld ? p.w_ld : K;
  for (int c = 0; c < SM_COLS_PER_WARP; ++c) {
    const int n = (blockIdx.x * 8 + warp) * SM_COLS_PER_WARP + c;
    if (n >= p.Cout) return;
    const float4* wr = reinterpret_cast<const float4*>(p.w + (int64_t)n * wld);
    float acc[SM_MAXROWS];
#pragma unroll
    for (int m = 0; m < SM_MAXROWS; ++m) acc[m] = 0.f;
    for (int q = lane; q < Q; q += 32) {
      const float4 w = __ldg(wr + q);
#pragma unroll
      for (int m = 0; m < SM_MAXROWS; ++m)
        if (m < M) {
          const float4 x = lsm_x[m * Q + q];
          acc[m] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[m]))));
        }
    }
#pragma unroll
    for (int m = 0; m < SM_MAXROWS; ++m)
      if (m < M) {
        float t = warp_sum(acc[m]) * p.alpha;
        if (lane == 0) {
          if (p.bias) t += __ldg(p.bias + n);
          if (p.rowvec) t += __ldg(p.rowvec + n);
          if (p.res) t += p.res[(int64_t)m * p.o_sp + n];
          if (p.act == FRIDO_ACT_RELU) t = fmaxf(t, 0.f);
          else if (p.act == FRIDO_ACT_SILU) t = silu_f(t);
          else if (p.act == FRIDO_ACT_GELU) t = gelu_erf(t);
          p.out[(int64_t)m * p.o_sp + n] = p.round_tf32 ? round_tf32(t) : t;
        }
      }
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int conv2d_simt(const FridoConvParams* p, cudaStream_t s) {
  if (!p || !p->a0 || !p->w || !p->out) return set_error(FRIDO_E_ARG, "conv2d: null pointer");
  if (p->B <= 0 || p->Hout <= 0 || p->Wout <= 0 || p->Cout <= 0 || p->c0 <= 0 || p->c1 < 0)
    return set_error(FRIDO_E_ARG, "conv2d: bad shape");
  if ((p->c1 > 0) != (p->a1 != nullptr)) return set_error(FRIDO_E_ARG, "conv2d: a1/c1 mismatch");
  if (p->x0 || p->x1 || p->cx0 || p->cx1) return set_error(FRIDO_E_ARG, "conv2d: the fused 1x1 side input is a tcgen05-engine feature");
  if (p->ups != 1 && p->ups != 2) return set_error(FRIDO_E_ARG, "conv2d: ups must be 1 or 2");
  if (p->ksize != 1 && p->ksize != 3) return set_error(FRIDO_E_ARG, "conv2d: ksize must be 1 or 3");
  if (p->act == FRIDO_ACT_GEGLU && (p->Cout & 1)) return set_error(FRIDO_E_ARG, "conv2d: GEGLU needs even Cout");
  if (p->out_u8 && (p->u8_mode != 0 && p->u8_mode != 1)) return set_error(FRIDO_E_ARG, "conv2d: u8_mode must be 0 or 1");
  const int Cin = p->c0 + p->c1;
  const int64_t Ktot = (int64_t)p->ksize * p->ksize * Cin;
  bool vecA = (Cin % 4 == 0) && (p->c0 % 4 == 0) && p->a0_sc == 1 && aligned16(p->a0) && p->a0_sb % 4 == 0 &&
              p->a0_sy % 4 == 0 && p->a0_sx % 4 == 0;
  if (p->a1)
    vecA = vecA && p->a1_sc == 1 && aligned16(p->a1) && p->a1_sb % 4 == 0 && p->a1_sy % 4 == 0 && p->a1_sx % 4 == 0;
  const bool vecW = (Ktot % 4 == 0) && aligned16(p->w) && (p->w_sb % 4 == 0) && (p->w_ld % 4 == 0);
  if ((p->out_hi != nullptr) != (p->out_lo != nullptr) || (p->out_hi && p->act == FRIDO_ACT_GEGLU))
    return set_error(FRIDO_E_ARG, "conv2d: out_hi/out_lo must come together and not with GEGLU");
  if (vecA && vecW && !p->out_hi && !p->a1 && p->ksize == 1 && p->stride == 1 && p->ups == 1 && p->B == 1 && p->Hout == 1 && p->Hin == 1 &&
      p->Wout == p->Win && p->Wout <= SM_MAXROWS && p->o_sn == 1 && p->w_sb == 0 && p->act != FRIDO_ACT_GEGLU && p->act != FRIDO_ACT_GEGLU_FAST &&
      p->Cout >= 64 && (size_t)p->Wout * Cin * 4 <= 96 * 1024 && !p->out_u8) {
    static DevOnce attr_l;
    if (attr_l.need()) cudaFuncSetAttribute(linear_smallm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const int cols_per_cta = 8 * SM_COLS_PER_WARP;
    linear_smallm_kernel<<<(p->Cout + cols_per_cta - 1) / cols_per_cta, 256, (size_t)p->Wout * Cin * 4, s>>>(*p);
    return check_launch("linear_smallm");
  }
  if (vecA && !p->out_hi && p->Cout <= SC_MAXCOUT && p->act != FRIDO_ACT_GEGLU && (size_t)p->Cout * Ktot * 4 <= 96 * 1024 && Ktot % 4 == 0) {
    const size_t smem = (size_t)p->Cout * Ktot * 4;
    static DevOnce attr;
    if (attr.need()) {
      cudaFuncSetAttribute(conv_smallcout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      cudaFuncSetAttribute(conv_smallcout_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    const size_t smem_t = (size_t)(SCT_TH + 2) * (SCT_TW + 2) * (Cin / 4 + 1) * 16 + smem;
    if (p->ksize == 3 && p->stride == 1 && p->pad == 1 && p->ups == 1 && !p->a1 && p->Hout == p->Hin && p->Wout == p->Win &&
        smem_t <= 200 * 1024 && p->Hout * p->Wout >= 256) {
      dim3 g((p->Wout + SCT_TW - 1) / SCT_TW, (p->Hout + SCT_TH - 1) / SCT_TH, p->B);
      conv_smallcout_tiled_kernel<<<g, SCT_TH * SCT_TW * SCT_PARTS, smem_t, s>>>(*p);
      return check_launch("conv2d_smallcout_tiled");
    }
    dim3 g((p->Hout * p->Wout + 127) / 128, p->B);
    conv_smallcout_kernel<<<g, 128, smem, s>>>(*p);
    return check_launch("conv2d_smallcout");
  }
  if (p->out_u8) return set_error(FRIDO_E_ARG, "conv2d: out_u8 needs Cout <= 4 and a 16-byte aligned dense NHWC source (the small-Cout head kernels)");
  dim3 grid((p->Hout * p->Wout + BM - 1) / BM, (p->Cout + BN - 1) / BN, p->B);
  if (grid.y > 65535 || grid.z > 65535) return set_error(FRIDO_E_ARG, "conv2d: grid too large");
  if (vecA && vecW) conv_simt_kernel<true, true><<<grid, NT, 0, s>>>(*p);
  else if (vecA) conv_simt_kernel<true, false><<<grid, NT, 0, s>>>(*p);
  else if (vecW) conv_simt_kernel<false, true><<<grid, NT, 0, s>>>(*p);
  else conv_simt_kernel<false, false><<<grid, NT, 0, s>>>(*p);
  return check_launch("conv2d_simt");
}

}  // namespace frido
