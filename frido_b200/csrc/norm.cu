// HBM-bound row/group kernels of the path: GroupNorm statistics, the fused
// normalise(+SPADE)(+SiLU) pass, LayerNorm, row softmax and the sinusoidal
// timestep embedding.  All NHWC fp32, float4 accesses, one pass over the data
// per kernel (softmax: three cached passes over one row).
#include "common.cuh"

namespace frido {

// ----------------------------------------------------------------------------
// GroupNorm statistics: fp32 per-thread partials over a pixel chunk, fp64
// shared + global atomics for the cross-thread reduction.
// ----------------------------------------------------------------------------
constexpr int GN_THREADS = 256;
constexpr int GN_MAXQPT = 4;  // quads (float4) of one pixel owned by one thread: C <= 4096

__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const FridoGnStatsParams p, int pix_per_cta) {
  __shared__ double s_sum[64], s_sq[64];
  const int C = p.c0 + p.c1;
  const int Q = C >> 2;
  const int cg = C / p.groups;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  if (tid < p.groups) { s_sum[tid] = 0.0; s_sq[tid] = 0.0; }
  __syncthreads();
  const int Qe = Q < GN_THREADS ? Q : GN_THREADS;  // quads covered per pixel pass
  const int PL = GN_THREADS / Qe;                  // pixel lanes
  const int pl = tid / Qe, ql = tid - pl * Qe;
  const int pix0 = blockIdx.x * pix_per_cta;
  const int pix1 = min(pix0 + pix_per_cta, p.HW);
  float s[GN_MAXQPT][4], ss[GN_MAXQPT][4];
#pragma unroll
  for (int j = 0; j < GN_MAXQPT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[j][e] = ss[j][e] = 0.f;
  if (pl < PL) {
    const float* a0 = p.a0 + (int64_t)b * p.HW * p.c0;
    const float* a1 = p.a1 ? p.a1 + (int64_t)b * p.HW * p.c1 : nullptr;
    for (int pix = pix0 + pl; pix < pix1; pix += PL) {
#pragma unroll
      for (int j = 0; j < GN_MAXQPT; ++j) {
        const int q = ql + j * GN_THREADS;
        if (q < Q) {
          const int c = q << 2;
          const float4 v = (c < p.c0) ? __ldg(reinterpret_cast<const float4*>(a0 + (int64_t)pix * p.c0 + c))
                                      : __ldg(reinterpret_cast<const float4*>(a1 + (int64_t)pix * p.c1 + (c - p.c0)));
          s[j][0] += v.x; ss[j][0] += v.x * v.x;
          s[j][1] += v.y; ss[j][1] += v.y * v.y;
          s[j][2] += v.z; ss[j][2] += v.z * v.z;
          s[j][3] += v.w; ss[j][3] += v.w * v.w;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < GN_MAXQPT; ++j) {
      const int q = ql + j * GN_THREADS;
      if (q < Q) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int g = ((q << 2) + e) / cg;
          atomicAdd(&s_sum[g], (double)s[j][e]);
          atomicAdd(&s_sq[g], (double)ss[j][e]);
        }
      }
    }
  }
  __syncthreads();
  if (tid < p.groups) {
    double* o = p.sums + ((int64_t)b * p.groups + tid) * 2;
    atomicAdd(o, s_sum[tid]);
    atomicAdd(o + 1, s_sq[tid]);
  }
}


// Per-channel statistics of one NHWC tensor: sums[b][c] = (sum, sum of squares) over the pixels, fp64 - the same format the
// tcgen05 conv epilogues accumulate (FridoConvParams.chan_sums), for tensors produced by the SIMT engine (the UNet's
// pre_input conv): computed ONCE, every GroupNorm that reads the tensor - alone or as half of a skip concat - then takes the
// producer-statistics path of norm_act instead of re-reducing the (concatenated) tensor.  groups = 0 selects this mode.
__global__ void __launch_bounds__(256) chan_stats_kernel(const FridoGnStatsParams p, int pix_per_cta, int Qe, int PL, int nj) {
  extern __shared__ float cst_sm[];  // [PL][C][2] per-thread partial sums, combined in a fixed order (replays stay bit-identical)
  const int C = p.c0, Q = C >> 2;
  const int b = blockIdx.y, tid = threadIdx.x;
  if (tid < Qe * PL) {
    const int pl = tid / Qe, ql = tid - pl * Qe;
    const int pix0 = blockIdx.x * pix_per_cta, pix1 = min(pix0 + pix_per_cta, p.HW);
    const float* a0 = p.a0 + (int64_t)b * p.HW * C;
    for (int j = 0; j < nj; ++j) {
      const int quad = ql + j * Qe;
      if (quad >= Q) break;
      float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
      for (int pix = pix0 + pl; pix < pix1; pix += PL) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(a0 + (int64_t)pix * C) + quad);
        s[0] += v.x; ss[0] += v.x * v.x; s[1] += v.y; ss[1] += v.y * v.y;
        s[2] += v.z; ss[2] += v.z * v.z; s[3] += v.w; ss[3] += v.w * v.w;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        cst_sm[2 * (pl * C + 4 * quad + e)] = s[e];
        cst_sm[2 * (pl * C + 4 * quad + e) + 1] = ss[e];
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += blockDim.x) {
    float t = 0.f;
    for (int pl = 0; pl < PL; ++pl) t += cst_sm[2 * pl * C + i];
    atomicAdd(p.sums + (int64_t)b * 2 * C + i, (double)t);
  }
}

static int gn_chunk(int B, int HW) {
  // aim for >= ~4 CTAs per SM overall, at least 16 pixels per CTA
  int target = (148 * 4 + B - 1) / B;
  int ppc = (HW + target - 1) / target;
  if (ppc < 16) ppc = 16;
  return ppc;
}

// ----------------------------------------------------------------------------
// normalise + affine (+SPADE) (+SiLU)
// Thread layout: tid -> (pixel lane pl, quad column ql); a thread keeps its channel quad for the whole CTA chunk, so
// gamma / beta / mean / rstd are loop invariants in registers and the pixel loop is pure streaming: 1 (or 3 with SPADE)
// float4 loads and one float4 store per iteration, four pixels in flight per thread.  Channels beyond 4*Qe are covered
// by nj passes (quad = ql + j*Qe).
// ----------------------------------------------------------------------------
constexpr int NA_MAX_THREADS = 384;
constexpr int NA_UNROLL = 4;

__global__ void __launch_bounds__(NA_MAX_THREADS) norm_act_kernel(const FridoNormActParams p, int pix_per_cta, int Qe, int PL, int nj) {
  __shared__ float s_mean[64], s_rstd[64];
  pdl_trigger();
  pdl_wait();
  const int C = p.c0 + p.c1;
  const int Q = C >> 2;
  const int cg = C / p.groups;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  // group statistics: EIGHT lanes per group (all groups of the CTA in one or two rounds instead of one warp per group in
  // sequence: the dependent load -> reduce -> sqrt chain of this prologue was ~5 us of every launch, fully exposed in the step)
  for (int g0 = 0; g0 < p.groups; g0 += (int)(blockDim.x >> 3)) {
    const int g = g0 + (tid >> 3), sub = tid & 7;
    double s0 = 0.0, s1 = 0.0;
    if (g < p.groups) {
      if (p.csum0) {  // group sums from the producers' per-channel sums (the group may straddle the two sources)
        for (int c = g * cg + sub; c < (g + 1) * cg; c += 8) {
          const double* cs = (c < p.c0) ? p.csum0 + ((int64_t)b * p.c0 + c) * 2 : p.csum1 + ((int64_t)b * p.c1 + (c - p.c0)) * 2;
          s0 += cs[0]; s1 += cs[1];
        }
      } else if (sub == 0) {
        const double* sm = p.sums + ((int64_t)b * p.groups + g) * 2;
        s0 = sm[0]; s1 = sm[1];
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    if (g < p.groups && sub == 0) {
      const double cnt = (double)cg * (double)p.HW;
      const double mean = s0 / cnt;
      double var = s1 / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = (float)mean;
      s_rstd[g] = (float)(1.0 / sqrt(var + (double)p.eps));
    }
  }
  __syncthreads();
  if (tid >= Qe * PL) return;
  const int pl = tid / Qe, ql = tid - pl * Qe;
  const int pix0 = blockIdx.x * pix_per_cta;
  const int pix1 = min(pix0 + pix_per_cta, p.HW);
  const float* gb = p.gb ? p.gb + (int64_t)b * p.HW * 2 * C : nullptr;
  float* out = p.out + (int64_t)b * p.HW * C;
  const bool silu = p.silu != 0, rnd = p.round_tf32 != 0;
  for (int j = 0; j < nj; ++j) {
    const int quad = ql + j * Qe;
    if (quad >= Q) break;
    const int c = quad << 2;
    const bool first = c < p.c0;
    const int cs = first ? p.c0 : p.c1;  // row stride of the source this quad lives in
    const float* src = first ? p.a0 + (int64_t)b * p.HW * p.c0 + c : p.a1 + (int64_t)b * p.HW * p.c1 + (c - p.c0);
    const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + c));
    const float4 bt = __ldg(reinterpret_cast<const float4*>(p.beta + c));
    const float g4[4] = {gm.x, gm.y, gm.z, gm.w};
    const float b4[4] = {bt.x, bt.y, bt.z, bt.w};
    float mean[4], rstd[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { const int g = (c + e) / cg; mean[e] = s_mean[g]; rstd[e] = s_rstd[g]; }
    for (int pix = pix0 + pl; pix < pix1; pix += PL * NA_UNROLL) {
      float4 xv[NA_UNROLL], ga[NA_UNROLL], be[NA_UNROLL];
#pragma unroll
      for (int u = 0; u < NA_UNROLL; ++u) {
        const int px = pix + u * PL;
        if (px < pix1) {
          xv[u] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)px * cs));
          if (gb) {
            // the SPADE maps are read once per step and are far larger than L2: streaming (evict-first) loads, so they do
            // not push the activation just written by the previous conv - and the output the next one reads - out of L2
            ga[u] = __ldcs(reinterpret_cast<const float4*>(gb + (int64_t)px * 2 * C + c));
            be[u] = __ldcs(reinterpret_cast<const float4*>(gb + (int64_t)px * 2 * C + C + c));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < NA_UNROLL; ++u) {
        const int px = pix + u * PL;
        if (px < pix1) {
          float x[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
          const float sg[4] = {ga[u].x, ga[u].y, ga[u].z, ga[u].w};
          const float sb[4] = {be[u].x, be[u].y, be[u].z, be[u].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float y = (x[e] - mean[e]) * rstd[e] * g4[e] + b4[e];
            if (gb) y = y * (1.0f + sg[e]) + sb[e];
            if (silu) y = silu_f(y);
            x[e] = rnd ? round_tf32(y) : y;
          }
          if (p.out_split) {  // the BF16x3 engine's operand form: bf16x2 hi(c0,c1), hi(c2,c3), lo(c0,c1), lo(c2,c3)
            uint16_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16(x[e], h[e], l[e]);
            *reinterpret_cast<uint4*>(out + (int64_t)px * C + c) =
                make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16),
                           (uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
          } else
          *reinterpret_cast<float4*>(out + (int64_t)px * C + c) = make_float4(x[0], x[1], x[2], x[3]);
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------
// GroupNorm statistics -> per-(image, channel) scale / shift for the convs that normalise on load (conv_nf.cu).  One CTA per
// image; the group mean / rstd are formed exactly as in norm_act_kernel (fp64 sums, fp32 mean and rstd).
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_finalize_kernel(const FridoGnFinalizeParams p) {
  __shared__ float s_mean[64], s_rstd[64];
  pdl_trigger();
  pdl_wait();
  const int C = p.c0 + p.c1;
  const int cg = C / p.groups;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  // eight lanes per group, all 32 groups in one round (same scheme and summation order as norm_act_kernel)
  for (int g0 = 0; g0 < p.groups; g0 += (int)(blockDim.x >> 3)) {
    const int g = g0 + (tid >> 3), sub = tid & 7;
    double s0 = 0.0, s1 = 0.0;
    if (g < p.groups) {
      if (p.csum0) {
        for (int c = g * cg + sub; c < (g + 1) * cg; c += 8) {
          const double* cs = (c < p.c0) ? p.csum0 + ((int64_t)b * p.c0 + c) * 2 : p.csum1 + ((int64_t)b * p.c1 + (c - p.c0)) * 2;
          s0 += cs[0]; s1 += cs[1];
        }
      } else if (sub == 0) {
        const double* sm = p.sums + ((int64_t)b * p.groups + g) * 2;
        s0 = sm[0]; s1 = sm[1];
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    if (g < p.groups && sub == 0) {
      const double cnt = (double)cg * (double)p.HW;
      const double mean = s0 / cnt;
      double var = s1 / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = (float)mean;
      s_rstd[g] = (float)(1.0 / sqrt(var + (double)p.eps));
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    const int g = c / cg;
    const float a = s_rstd[g] * __ldg(p.gamma + c);
    reinterpret_cast<float2*>(p.ab)[(int64_t)b * C + c] = make_float2(a, fmaf(-s_mean[g], a, __ldg(p.beta + c)));
  }
}

// ----------------------------------------------------------------------------
// LayerNorm: one warp per row, row cached in registers (C <= 1024), two-pass var
// ----------------------------------------------------------------------------
constexpr int LN_MAXQ = 8;
__global__ void __launch_bounds__(256) layernorm_kernel(const FridoLayerNormParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();
  if (warp >= p.rows) return;
  const int Q = p.C >> 2;
  const float* x = p.x + (int64_t)warp * p.C;
  float4 v[LN_MAXQ];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAXQ; ++j) {
    const int q = lane + j * 32;
    if (q < Q) {
      v[j] = __ldg(reinterpret_cast<const float4*>(x) + q);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  }
  const float mean = warp_sum(s) / (float)p.C;
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAXQ; ++j) {
    const int q = lane + j * 32;
    if (q < Q) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)p.C + p.eps);
  float* out = p.out + (int64_t)warp * p.C;
#pragma unroll
  for (int j = 0; j < LN_MAXQ; ++j) {
    const int q = lane + j * 32;
    if (q < Q) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + q);
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta) + q);
      float4 o;
      o.x = (v[j].x - mean) * rstd * g.x + b.x;
      o.y = (v[j].y - mean) * rstd * g.y + b.y;
      o.z = (v[j].z - mean) * rstd * g.z + b.z;
      o.w = (v[j].w - mean) * rstd * g.w + b.w;
      if (p.round_tf32) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
      reinterpret_cast<float4*>(out)[q] = o;
      if (p.out_hi) {
        uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
        split_bf16(o.x, h0, l0); split_bf16(o.y, h1, l1); split_bf16(o.z, h2, l2); split_bf16(o.w, h3, l3);
        reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out_hi) + (int64_t)warp * p.C)[q] =
            make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
        reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out_lo) + (int64_t)warp * p.C)[q] =
            make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
      }
    }
  }
}

// ----------------------------------------------------------------------------
// Row softmax.  n <= 1024: one warp per row; else one CTA per row.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_warp_kernel(const FridoSoftmaxParams p) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();
  if (row >= p.rows) return;
  const float* s = p.s + row * p.ld;
  float* o = p.out + row * p.ld;
  float m = -INFINITY;
  for (int j = lane; j < p.n; j += 32) m = fmaxf(m, s[j] * p.scale);
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < p.n; j += 32) sum += __expf(s[j] * p.scale - m);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < p.n; j += 32) {
    const float v = __expf(s[j] * p.scale - m) * inv;
    o[j] = p.round_tf32 ? round_tf32(v) : v;
  }
}

__global__ void __launch_bounds__(256) softmax_cta_kernel(const FridoSoftmaxParams p) {
  __shared__ float red[8];
  __shared__ float bc;
  const int64_t row = blockIdx.x;
  const float* s = p.s + row * p.ld;
  float* o = p.out + row * p.ld;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int n4 = p.n >> 2;
  float m = -INFINITY;
  for (int j = tid; j < n4; j += 256) {
    const float4 v = reinterpret_cast<const float4*>(s)[j];
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) * p.scale);
  }
  for (int j = (n4 << 2) + tid; j < p.n; j += 256) m = fmaxf(m, s[j] * p.scale);
  m = warp_max(m);
  if (lane == 0) red[w] = m;
  __syncthreads();
  if (tid == 0) { float t = red[0]; for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]); bc = t; }
  __syncthreads();
  m = bc;
  float sum = 0.f;
  for (int j = tid; j < n4; j += 256) {
    const float4 v = reinterpret_cast<const float4*>(s)[j];
    sum += (__expf(v.x * p.scale - m) + __expf(v.y * p.scale - m)) + (__expf(v.z * p.scale - m) + __expf(v.w * p.scale - m));
  }
  for (int j = (n4 << 2) + tid; j < p.n; j += 256) sum += __expf(s[j] * p.scale - m);
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[w] = sum;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; bc = t; }
  __syncthreads();
  const float inv = 1.0f / bc;
  for (int j = tid; j < n4; j += 256) {
    float4 v = reinterpret_cast<const float4*>(s)[j];
    v.x = __expf(v.x * p.scale - m) * inv; v.y = __expf(v.y * p.scale - m) * inv;
    v.z = __expf(v.z * p.scale - m) * inv; v.w = __expf(v.w * p.scale - m) * inv;
    if (p.round_tf32) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    reinterpret_cast<float4*>(o)[j] = v;
  }
  for (int j = (n4 << 2) + tid; j < p.n; j += 256) {
    const float v = __expf(s[j] * p.scale - m) * inv;
    o[j] = p.round_tf32 ? round_tf32(v) : v;
  }
}

// ----------------------------------------------------------------------------
// timestep embedding
// ----------------------------------------------------------------------------
__global__ void time_embed_kernel(const FridoTimeEmbedParams p) {
  const int half = p.dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B * half) return;
  const int b = i / half, j = i - b * half;
  // util.py:160-166: freqs = exp(-ln(max_period) * j / half) in fp32; args = t.float() * freqs
  const float nl = (float)(-log((double)p.max_period));  // python float -> fp32 scalar
  const float f = p.freqs ? p.freqs[j] : expf(nl * (float)j / (float)half);
  const float a = (float)p.t[b] * f;
  p.out[(int64_t)b * p.dim + j] = cosf(a);
  p.out[(int64_t)b * p.dim + half + j] = sinf(a);
  if ((p.dim & 1) && j == 0) p.out[(int64_t)b * p.dim + p.dim - 1] = 0.f;
}

// nearest x2 upsample, NHWC, one float4 per thread
__global__ void __launch_bounds__(256) upsample2x_kernel(const FridoUpsampleParams p) {
  const int Q = p.C >> 2;
  const int64_t total = (int64_t)p.B * 4 * p.H * p.W * Q;
  pdl_trigger();
  pdl_wait();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(e % Q);
    int64_t r = e / Q;
    const int ox = (int)(r % (2 * p.W)); r /= (2 * p.W);
    const int oy = (int)(r % (2 * p.H));
    const int b = (int)(r / (2 * p.H));
    float4 v = __ldg(reinterpret_cast<const float4*>(p.x + (((int64_t)b * p.H + (oy >> 1)) * p.W + (ox >> 1)) * p.C) + q);
    if (p.round_tf32) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    if (p.out_split) {
      uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
      split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
      reinterpret_cast<uint4*>(p.out + (((int64_t)b * 2 * p.H + oy) * 2 * p.W + ox) * p.C)[q] =
          make_uint4((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16),
                     (uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
    } else
    reinterpret_cast<float4*>(p.out + (((int64_t)b * 2 * p.H + oy) * 2 * p.W + ox) * p.C)[q] = v;
  }
}

// token + position embedding gather (condition encoder)
__global__ void __launch_bounds__(256) embed_kernel(const FridoEmbedParams p) {
  const int Q = p.D >> 2;
  const int64_t total = (int64_t)p.B * p.L * Q;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(e % Q);
    const int64_t bl = e / Q;
    const int l = (int)(bl % p.L);
    int64_t tok = p.tokens[bl];
    tok = tok < 0 ? 0 : (tok >= p.vocab ? p.vocab - 1 : tok);
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.tok_emb + tok * p.D) + q);
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.pos_emb + (int64_t)l * p.D) + q);
    reinterpret_cast<float4*>(p.out + bl * p.D)[q] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// short-sequence multi-head attention: one CTA per (b, h); K, V in shared memory; one warp per query row
__global__ void __launch_bounds__(256) mha_small_kernel(const FridoMhaParams p) {
  extern __shared__ float sm[];
  const int L = p.L, Dh = p.Dh, HD = p.H * p.Dh;
  float* ks = sm;                 // [L][Dh+1]
  float* vs = sm + (size_t)L * (Dh + 1);  // [L][Dh]
  float* ps = vs + (size_t)L * Dh;        // [8 warps][L] probabilities
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const float* base = p.qkv + (int64_t)b * L * 3 * HD + h * Dh;
  for (int i = threadIdx.x; i < L * Dh; i += blockDim.x) {
    const int j = i / Dh, d = i - j * Dh;
    ks[j * (Dh + 1) + d] = base[(int64_t)j * 3 * HD + HD + d];
    vs[j * Dh + d] = base[(int64_t)j * 3 * HD + 2 * HD + d];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pw = ps + warp * L;
  for (int i = warp; i < L; i += 8) {
    const float* qrow = base + (int64_t)i * 3 * HD;
    float m = -INFINITY;
    for (int j = lane; j < L; j += 32) {
      float s = 0.f;
      for (int d = 0; d < Dh; ++d) s = fmaf(__ldg(qrow + d), ks[j * (Dh + 1) + d], s);
      s *= p.scale;
      pw[j] = s;
      m = fmaxf(m, s);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float e = __expf(pw[j] - m);
      pw[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int d = lane; d < Dh; d += 32) {
      float o = 0.f;
      for (int j = 0; j < L; ++j) o = fmaf(pw[j], vs[j * Dh + d], o);
      p.out[((int64_t)b * L + i) * HD + h * Dh + d] = o * inv;
    }
    __syncwarp();
  }
}

}  // namespace frido

using namespace frido;

extern "C" int frido_embed_tokens(const FridoEmbedParams* p, void* stream) {
  if (!p || !p->tokens || !p->tok_emb || !p->pos_emb || !p->out || (p->D & 3) || p->vocab <= 0)
    return set_error(FRIDO_E_ARG, "embed_tokens: bad argument");
  const int64_t total = (int64_t)p->B * p->L * (p->D >> 2);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  embed_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("embed_tokens");
}

extern "C" int frido_mha_small(const FridoMhaParams* p, void* stream) {
  if (!p || !p->qkv || !p->out || p->B <= 0 || p->L <= 0 || p->H <= 0 || p->Dh <= 0) return set_error(FRIDO_E_ARG, "mha_small: bad argument");
  const size_t smem = ((size_t)p->L * (p->Dh + 1) + (size_t)p->L * p->Dh + 8 * (size_t)p->L) * sizeof(float);
  if (smem > 200 * 1024) return set_error(FRIDO_E_ARG, "mha_small: sequence too long for shared memory");
  static DevOnce attr;
  if (attr.need()) cudaFuncSetAttribute(mha_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  mha_small_kernel<<<p->B * p->H, 256, smem, (cudaStream_t)stream>>>(*p);
  return check_launch("mha_small");
}

extern "C" int frido_upsample2x(const FridoUpsampleParams* p, void* stream) {
  if (!p || !p->x || !p->out || (p->C & 3)) return set_error(FRIDO_E_ARG, "upsample2x: bad argument");
  const int64_t total = (int64_t)p->B * 4 * p->H * p->W * (p->C >> 2);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(upsample2x_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, *p);
  return check_launch("upsample2x");
}

extern "C" int frido_gn_stats(const FridoGnStatsParams* p, void* stream) {
  if (!p || !p->a0 || !p->sums) return set_error(FRIDO_E_ARG, "gn_stats: null pointer");
  if (p->groups == 0) {  // per-channel mode
    if (p->a1 || p->c1 || (p->c0 & 3) || p->c0 <= 0 || p->c0 > 4096) return set_error(FRIDO_E_ARG, "gn_stats: per-channel mode takes one source, C % 4 == 0");
    int ppc = gn_chunk(p->B, p->HW);
    if (ppc < 64) ppc = 64;  // fewer, fatter CTAs: each one ends with 2C global fp64 atomics
    const int Q = p->c0 >> 2, nj = (Q + 255) / 256, Qe = (Q + nj - 1) / nj;
    int PL = 256 / Qe;
    if (PL < 1) PL = 1;
    dim3 grid((p->HW + ppc - 1) / ppc, p->B);
    const size_t smem = (size_t)PL * 2 * p->c0 * sizeof(float);
    if (smem > 48 * 1024) return set_error(FRIDO_E_ARG, "gn_stats: per-channel mode: too many channels");
    chan_stats_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(*p, ppc, Qe, PL, nj);
    return check_launch("chan_stats");
  }
  const int C = p->c0 + p->c1;
  if (p->groups <= 0 || p->groups > 64 || C % p->groups || (C & 3) || (p->c0 & 3) || C > GN_THREADS * 4 * GN_MAXQPT)
    return set_error(FRIDO_E_ARG, "gn_stats: unsupported channel count");
  if ((p->c1 > 0) != (p->a1 != nullptr)) return set_error(FRIDO_E_ARG, "gn_stats: a1/c1 mismatch");
  const int ppc = gn_chunk(p->B, p->HW);
  dim3 grid((p->HW + ppc - 1) / ppc, p->B);
  gn_stats_kernel<<<grid, GN_THREADS, 0, (cudaStream_t)stream>>>(*p, ppc);
  return check_launch("gn_stats");
}

extern "C" int frido_norm_act(const FridoNormActParams* p, void* stream) {
  if (!p || !p->a0 || (!p->sums && !p->csum0) || !p->out || !p->gamma || !p->beta) return set_error(FRIDO_E_ARG, "norm_act: null pointer");
  if (p->csum0 && p->c1 > 0 && !p->csum1) return set_error(FRIDO_E_ARG, "norm_act: csum1 missing for the second source");
  const int C = p->c0 + p->c1;
  if (p->groups <= 0 || p->groups > 64 || C % p->groups || (C & 3) || (p->c0 & 3))
    return set_error(FRIDO_E_ARG, "norm_act: unsupported channel count");
  if ((p->c1 > 0) != (p->a1 != nullptr)) return set_error(FRIDO_E_ARG, "norm_act: a1/c1 mismatch");
  int ppc = gn_chunk(p->B, p->HW);
  if (ppc == 16 && (long long)p->B * ((p->HW + 15) / 16) < 148) ppc = 8;  // 8x8 / 4x4 levels: at least ~one CTA per SM
  dim3 grid((p->HW + ppc - 1) / ppc, p->B);
  const int Q = C >> 2;
  const int nj = (Q + 255) / 256;
  const int Qe = (Q + nj - 1) / nj;
  int PL = (256 + Qe / 2) / Qe;  // pixel lanes: about 256 threads per CTA
  if (PL < 1) PL = 1;
  if (PL > ppc) PL = ppc;
  while (PL > 1 && Qe * PL > NA_MAX_THREADS) --PL;
  const int threads = (Qe * PL + 31) / 32 * 32;
  launch_pdl(norm_act_kernel, grid, dim3(threads), 0, (cudaStream_t)stream, *p, ppc, Qe, PL, nj);
  return check_launch("norm_act");
}

extern "C" int frido_gn_finalize(const FridoGnFinalizeParams* p, void* stream) {
  if (!p || (!p->sums && !p->csum0) || !p->gamma || !p->beta || !p->ab) return set_error(FRIDO_E_ARG, "gn_finalize: null pointer");
  if (p->csum0 && p->c1 > 0 && !p->csum1) return set_error(FRIDO_E_ARG, "gn_finalize: csum1 missing for the second source");
  const int C = p->c0 + p->c1;
  if (p->B <= 0 || p->HW <= 0 || p->groups <= 0 || p->groups > 64 || C <= 0 || C % p->groups || (reinterpret_cast<uintptr_t>(p->ab) & 7))
    return set_error(FRIDO_E_ARG, "gn_finalize: unsupported shape");
  launch_pdl(gn_finalize_kernel, dim3(p->B), dim3(256), 0, (cudaStream_t)stream, *p);
  return check_launch("gn_finalize");
}

extern "C" int frido_layernorm(const FridoLayerNormParams* p, void* stream) {
  if (!p || !p->x || !p->out || !p->gamma || !p->beta) return set_error(FRIDO_E_ARG, "layernorm: null pointer");
  if ((p->C & 3) || p->C > LN_MAXQ * 128 || p->rows <= 0) return set_error(FRIDO_E_ARG, "layernorm: unsupported C");
  if ((p->out_hi != nullptr) != (p->out_lo != nullptr)) return set_error(FRIDO_E_ARG, "layernorm: out_hi / out_lo come together");
  const int64_t blocks = (p->rows + 7) / 8;
  launch_pdl(layernorm_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, *p);
  return check_launch("layernorm");
}

extern "C" int frido_softmax(const FridoSoftmaxParams* p, void* stream) {
  if (!p || !p->s || !p->out || p->rows <= 0 || p->n <= 0) return set_error(FRIDO_E_ARG, "softmax: bad argument");
  if (p->n <= 1024) {
    const int64_t blocks = (p->rows + 7) / 8;
    launch_pdl(softmax_warp_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, *p);
  } else {
    if ((p->ld & 3) || (reinterpret_cast<uintptr_t>(p->s) & 15) || (reinterpret_cast<uintptr_t>(p->out) & 15))
      return set_error(FRIDO_E_ARG, "softmax: wide rows need 16B-aligned rows");
    softmax_cta_kernel<<<(unsigned)p->rows, 256, 0, (cudaStream_t)stream>>>(*p);
  }
  return check_launch("softmax");
}

extern "C" int frido_time_embed(const FridoTimeEmbedParams* p, void* stream) {
  if (!p || !p->t || !p->out || p->dim < 2) return set_error(FRIDO_E_ARG, "time_embed: bad argument");
  const int n = p->B * (p->dim / 2);
  time_embed_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("time_embed");
}
