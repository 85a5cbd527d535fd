// Fused single-head attention for SHORT key sequences (fp32 SIMT, flash-style online softmax):
//   q'        = LayerNorm(q[b,n,:])                       (optional, attention.py:203-205 / :323-325)
//   out[b,n,:] = softmax_j(scale * q' . k[b,j,:]) @ v[b,j,:] + bias + res[b,n,:]      (bias / res optional)
// Replaces einsum -> *scale -> softmax -> einsum of CrossAttention.forward (attention.py:178-191) where the tensor-core
// engine has nothing to chew on: the cross-attention to a 26-token layout condition (QK^T is a [N x 26] product, PV a
// K=26 product) and the self-attention of the 8x8 level (64 tokens per image).  The scores never leave the SM.
// With k = K Wq and v = V Wo^T folded on the host side (both step-invariant for the condition) and the LayerNorm, output
// bias and residual fused here, the whole cross-attention sub-block x + to_out(attn(LN(x), ctx)) is this one kernel:
// one read of x, one write of the result.
//
// One warp owns R query rows; the C channels are spread over the lanes as float4 quads (lane, lane+32, ...), so q and the
// output accumulator live in registers.  Keys/values are staged in shared memory in chunks of KC <= 32 keys, shared by
// the 8 warps of the CTA.  Scores: every lane forms its partial dot products for a block of 32 (key,row) pairs, and one
// transposing butterfly (31 shuffles instead of 32 x 5) leaves pair L fully reduced in lane L; running (max, sum) per row
// rescale the accumulator between chunks.
#include "common.cuh"

namespace frido {

constexpr int ATTN_WARPS = 8;
constexpr int ATTN_SMEM_MAX = 200 * 1024;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// reduce-scatter over the warp: on return v[0] of lane L holds the sum over all lanes of the caller's v[L]
__device__ __forceinline__ void warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

// reductions over the lanes that share (lane % R): xor offsets R, 2R, ..., 16
template <int R>
__device__ __forceinline__ float row_max(float v) {
#pragma unroll
  for (int o = 16; o >= R; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int R>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int o = 16; o >= R; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// SINGLE: all keys fit one chunk (Nk <= KC): straight-line code, q is dead before the accumulator comes alive.
template <int NJ, int R, bool SINGLE>
__global__ void __launch_bounds__(ATTN_WARPS * 32, SINGLE ? 2 : 1) attn_small_kernel(const FridoAttnParams p, int KC, int iters) {
  constexpr int KB = 32 / R;  // keys per reduction block (32 (key,row) pairs)
  constexpr int NB = R;       // blocks per chunk (KC <= 32)
  extern __shared__ float4 attn_sm[];
  const int Q = p.C >> 2;
  float4* Ks = attn_sm;
  float4* Vs = attn_sm + (size_t)KC * Q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  pdl_trigger();
  pdl_wait();

  const float* kb = p.k + (int64_t)b * p.k_sb;
  const float* vb = p.v + (int64_t)b * p.v_sb;
  // K/V chunk -> shared memory with cp.async: every thread fires all of its 16-byte copies without waiting, so the
  // staging runs at L2 bandwidth instead of one load latency per float4 (and overlaps the q load / LayerNorm below)
  auto stage = [&](int j0, int kc) {
    for (int kk = warp; kk < kc; kk += ATTN_WARPS) {
      const float4* ks = reinterpret_cast<const float4*>(kb + (int64_t)(j0 + kk) * p.k_ld);
      const float4* vs = reinterpret_cast<const float4*>(vb + (int64_t)(j0 + kk) * p.v_ld);
      for (int quad = lane; quad < Q; quad += 32) {
        cp_async16(Ks + kk * Q + quad, ks + quad);
        cp_async16(Vs + kk * Q + quad, vs + quad);
      }
    }
  };
  stage(0, min(KC, p.Nk));

  // `iters` row groups per CTA (SINGLE only): the staged K / V chunk is reused, so launches whose CTAs would come in more
  // than one wave pay the staging once per CTA instead of once per 8*R rows
  for (int it = 0; it < iters; ++it) {
  const int row0 = ((blockIdx.x * iters + it) * ATTN_WARPS + warp) * R;
  float4 q[R][NJ], o[R][NJ];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int n = row0 + r;
    const float4* qp = reinterpret_cast<const float4*>(p.q + (int64_t)b * p.q_sb + (int64_t)(n < p.N ? n : 0) * p.q_ld);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int quad = lane + 32 * j;
      q[r][j] = (quad < Q && n < p.N) ? __ldg(qp + quad) : make_float4(0.f, 0.f, 0.f, 0.f);
      o[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (p.ln_gamma) {  // LayerNorm of the query rows, two-pass variance like layernorm_kernel (norm.cu)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) s += (q[r][j].x + q[r][j].y) + (q[r][j].z + q[r][j].w);  // lanes beyond Q hold zeros
      const float mean = warp_sum(s) / (float)p.C;
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (lane + 32 * j < Q) {
          const float a = q[r][j].x - mean, c = q[r][j].y - mean, d = q[r][j].z - mean, e = q[r][j].w - mean;
          ss += (a * a + c * c) + (d * d + e * e);
        }
      }
      const float rstd = rsqrtf(warp_sum(ss) / (float)p.C + p.ln_eps);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int quad = lane + 32 * j;
        if (quad < Q) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_gamma) + quad);
          const float4 be = __ldg(reinterpret_cast<const float4*>(p.ln_beta) + quad);
          q[r][j].x = (q[r][j].x - mean) * rstd * g.x + be.x;
          q[r][j].y = (q[r][j].y - mean) * rstd * g.y + be.y;
          q[r][j].z = (q[r][j].z - mean) * rstd * g.z + be.z;
          q[r][j].w = (q[r][j].w - mean) * rstd * g.w + be.w;
        }
      }
    }
  }
  float m_run = -INFINITY, l_run = 0.f;  // of row lane % R

  int j0 = 0;
  do {
    const int kc = min(KC, p.Nk - j0);
    if (!SINGLE && j0 > 0) {
      __syncthreads();  // every warp is done with the previous chunk
      stage(j0, kc);
    }
    cp_async_commit_wait_all();
    __syncthreads();
    // ---- scores: block blk covers keys [blk*KB, blk*KB + KB); lane L ends up with pair (key blk*KB + L/R, row L%R).
    // The block loop and the key loop of P V below stay ROLLED: fully unrolled this kernel was 124 KB of straight-line
    // code that every warp streamed through once, and instruction-cache misses were its largest stall (ncu).
    const int nblk = (kc + KB - 1) / KB;
    float sv[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) sv[i] = -INFINITY;
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
      float part[32];
#pragma unroll
      for (int kl = 0; kl < KB; ++kl) {
        const int kk = min(blk * KB + kl, kc - 1);  // keys beyond kc recompute the last key; masked below
        const float4* kr = Ks + kk * Q;
#pragma unroll
        for (int r = 0; r < R; ++r) part[kl * R + r] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int quad = lane + 32 * j;
          if (quad < Q) {
            const float4 kv = kr[quad];
#pragma unroll
            for (int r = 0; r < R; ++r)
              part[kl * R + r] = fmaf(q[r][j].x, kv.x, fmaf(q[r][j].y, kv.y, fmaf(q[r][j].z, kv.z, fmaf(q[r][j].w, kv.w, part[kl * R + r]))));
          }
        }
      }
      warp_transpose_sum(part, lane);
      const float val = (blk * KB + lane / R < kc) ? part[0] * p.scale : -INFINITY;
#pragma unroll
      for (int i = 0; i < NB; ++i)
        if (i == blk) sv[i] = val;  // predicated moves: sv stays in registers
    }
    // ---- softmax state of row lane % R (replicated over the lanes that share it)
    float mloc = sv[0];
#pragma unroll
    for (int i = 1; i < NB; ++i) mloc = fmaxf(mloc, sv[i]);
    const float m_new = fmaxf(m_run, row_max<R>(mloc));
    float pv[NB], psum = 0.f;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      pv[i] = __expf(sv[i] - m_new);  // exp(-inf) = 0 for pairs beyond kc
      psum += pv[i];
    }
    const float corr = __expf(m_run - m_new);  // first chunk: exp(-inf) = 0
    l_run = l_run * corr + row_sum<R>(psum);
    m_run = m_new;
    if (!SINGLE) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float cr = __shfl_sync(0xffffffffu, corr, r);
#pragma unroll
        for (int j = 0; j < NJ; ++j) { o[r][j].x *= cr; o[r][j].y *= cr; o[r][j].z *= cr; o[r][j].w *= cr; }
      }
    }
    // ---- accumulate P V
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
      float pvb = pv[0];
#pragma unroll
      for (int i = 1; i < NB; ++i)
        if (i == blk) pvb = pv[i];
      const int kend = min(KB, kc - blk * KB);
#pragma unroll 2
      for (int kl = 0; kl < kend; ++kl) {
        const float4* vr = Vs + (blk * KB + kl) * Q;
        float pk[R];
#pragma unroll
        for (int r = 0; r < R; ++r) pk[r] = __shfl_sync(0xffffffffu, pvb, kl * R + r);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int quad = lane + 32 * j;
          if (quad < Q) {
            const float4 vv = vr[quad];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              o[r][j].x = fmaf(pk[r], vv.x, o[r][j].x); o[r][j].y = fmaf(pk[r], vv.y, o[r][j].y);
              o[r][j].z = fmaf(pk[r], vv.z, o[r][j].z); o[r][j].w = fmaf(pk[r], vv.w, o[r][j].w);
            }
          }
        }
      }
    }
  } while (!SINGLE && (j0 += KC) < p.Nk);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int n = row0 + r;
    const float inv = 1.0f / __shfl_sync(0xffffffffu, l_run, r);
    const bool live = n < p.N;
    float4* op = reinterpret_cast<float4*>(p.out + (int64_t)b * p.o_sb + (int64_t)(live ? n : 0) * p.o_ld);
    const float4* rp = p.res ? reinterpret_cast<const float4*>(p.res + (int64_t)b * p.r_sb + (int64_t)(live ? n : 0) * p.r_ld) : nullptr;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int quad = lane + 32 * j;
      float4 v = make_float4(o[r][j].x * inv, o[r][j].y * inv, o[r][j].z * inv, o[r][j].w * inv);
      if (quad < Q && live) {
        if (p.bias) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias) + quad);
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        if (rp) {
          const float4 rr = __ldg(rp + quad);
          v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
        }
        op[quad] = v;
        s += (v.x + v.y) + (v.z + v.w);
      } else {
        v = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      o[r][j] = v;
    }
    if (p.ln2_gamma) {  // LayerNorm of the row just produced (the block's next norm, attention.py:325), second output
      const float mean = warp_sum(s) / (float)p.C;
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (lane + 32 * j < Q) {
          const float a = o[r][j].x - mean, c = o[r][j].y - mean, d = o[r][j].z - mean, e = o[r][j].w - mean;
          ss += (a * a + c * c) + (d * d + e * e);
        }
      }
      const float rstd = rsqrtf(warp_sum(ss) / (float)p.C + p.ln2_eps);
      float4* o2 = reinterpret_cast<float4*>(p.out2 + (int64_t)b * p.o_sb + (int64_t)(live ? n : 0) * p.o_ld);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int quad = lane + 32 * j;
        if (quad < Q && live) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln2_gamma) + quad);
          const float4 be = __ldg(reinterpret_cast<const float4*>(p.ln2_beta) + quad);
          o2[quad] = make_float4((o[r][j].x - mean) * rstd * g.x + be.x, (o[r][j].y - mean) * rstd * g.y + be.y,
                                 (o[r][j].z - mean) * rstd * g.z + be.z, (o[r][j].w - mean) * rstd * g.w + be.w);
        }
      }
    }
  }
  }  // row groups
}

template <int NJ, int R, bool SINGLE>
static int launch_attn2(const FridoAttnParams* p, int KC, cudaStream_t s) {
  static DevOnce attr;
  if (attr.need()) {
    if (cudaFuncSetAttribute(attn_small_kernel<NJ, R, SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM_MAX) != cudaSuccess)
      return set_error(FRIDO_E_LAUNCH, "attn_small: cannot opt in to dynamic shared memory");
  }
  const size_t smem = (size_t)2 * KC * p->C * sizeof(float);
  const int groups = (p->N + ATTN_WARPS * R - 1) / (ATTN_WARPS * R);
  // row groups per CTA: grow until the launch fits one wave of resident CTAs (at most 4)
  int iters = 1;
  if (SINGLE) {
    const int per_sm = smem > 110 * 1024 ? 1 : 2;
    while (iters < 4 && (long long)((groups + iters - 1) / iters) * p->B > 148LL * per_sm) ++iters;
  }
  dim3 grid((groups + iters - 1) / iters, p->B);
  launch_pdl(attn_small_kernel<NJ, R, SINGLE>, grid, dim3(ATTN_WARPS * 32), smem, s, *p, KC, iters);
  return check_launch("attn_small");
}

template <int NJ, int R>
static int launch_attn(const FridoAttnParams* p, int KC, cudaStream_t s) {
  return p->Nk <= KC ? launch_attn2<NJ, R, true>(p, KC, s) : launch_attn2<NJ, R, false>(p, KC, s);
}

}  // namespace frido

using namespace frido;

extern "C" int frido_attn_small(const FridoAttnParams* p, void* stream) {
  if (!p || !p->q || !p->k || !p->v || !p->out || p->B <= 0 || p->N <= 0 || p->Nk <= 0 || p->C <= 0)
    return set_error(FRIDO_E_ARG, "attn_small: bad argument");
  if ((p->C & 3) || p->C > 1024) return set_error(FRIDO_E_ARG, "attn_small: C must be a multiple of 4, at most 1024");
  if ((p->ln_gamma != nullptr) != (p->ln_beta != nullptr)) return set_error(FRIDO_E_ARG, "attn_small: ln_gamma / ln_beta come together");
  if ((p->ln2_gamma != nullptr) != (p->ln2_beta != nullptr) || (p->ln2_gamma != nullptr) != (p->out2 != nullptr))
    return set_error(FRIDO_E_ARG, "attn_small: ln2_gamma / ln2_beta / out2 come together");
  const int64_t strides[10] = {p->q_sb, p->q_ld, p->k_sb, p->k_ld, p->v_sb, p->v_ld, p->o_sb, p->o_ld, p->r_sb, p->r_ld};
  for (int i = 0; i < 10; ++i)
    if (strides[i] & 3) return set_error(FRIDO_E_ARG, "attn_small: strides must be multiples of 4 floats");
  const void* ptrs[11] = {p->q, p->k, p->v, p->out, p->ln_gamma, p->ln_beta, p->bias, p->res, p->ln2_gamma, p->ln2_beta, p->out2};
  for (int i = 0; i < 11; ++i)
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) return set_error(FRIDO_E_ARG, "attn_small: pointers must be 16-byte aligned");
  int KC = ATTN_SMEM_MAX / (8 * p->C);  // K and V chunk, fp32
  if (KC > 32) KC = 32;
  if (KC > p->Nk) KC = p->Nk;
  if (KC < 1) return set_error(FRIDO_E_ARG, "attn_small: C too large for shared memory");
  const int nj = ((p->C >> 2) + 31) / 32;
  cudaStream_t s = (cudaStream_t)stream;
  // rows per warp by register budget: q and the accumulator take 8 * NJ * R registers
  if (nj <= 3) return launch_attn<3, 4>(p, KC, s);
  if (nj <= 5) return launch_attn<5, 2>(p, KC, s);
  return launch_attn<8, 1>(p, KC, s);
}
