// Sampler-level kernels (NCHW fp32 latents, all HBM/latency-bound and tiny):
// per-step timestep broadcast, the fused DDIM/PLMS x_{t-1} update, the
// inter-stage pool/upsample snap, and the VQ codebook lookup.
//
// The update reproduces the reference's fp32 evaluation order exactly
// (ddim.py:243-268, plms.py:285-299): explicit *_rn intrinsics so nvcc cannot
// contract multiplies and adds into FMAs.
#include "common.cuh"

namespace frido {

__global__ void step_begin_kernel(const FridoStepBeginParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  int i = *p.step;
  if (p.use_next) i = min(i + 1, p.T - 1);  // plms.py:160
  if (i > p.T - 1) i = p.T - 1;
  p.ts[b] = p.t_table[i];
}

// Philox4x32-10 + Box-Muller (used only when eta > 0 and no noise is injected)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ float philox_normal(uint64_t seed, uint64_t stream, uint64_t idx) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float u1 = ((float)c[0] + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)c[1] + 0.5f) * 2.3283064365386963e-10f;
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

__device__ __forceinline__ float cfg_mix(float ec, float eu, float s) {
  return __fadd_rn(eu, __fmul_rn(s, __fsub_rn(ec, eu)));  // ddim.py:226
}

__global__ void __launch_bounds__(256) update_kernel(const FridoUpdateParams p) {
  const int i = *p.step;  // step counter within the stage; index = T-1-i is baked into coef order
  const float a_t = p.coef[4 * i + 0], a_prev = p.coef[4 * i + 1], sigma = p.coef[4 * i + 2], s1m = p.coef[4 * i + 3];
  const float sqrt_at = __fsqrt_rn(a_t);
  const float sqrt_ap = __fsqrt_rn(a_prev);
  const float dir_c = __fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, a_prev), __fmul_rn(sigma, sigma)));
  const int c_act = p.c_end - p.c_start;
  const int64_t per_img = (int64_t)p.c_end * p.HW;
  const int64_t total = (int64_t)p.B * per_img;
  const int64_t act_n = (int64_t)p.B * c_act * p.HW;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / per_img);
    const int64_t r = e - (int64_t)b * per_img;
    const int c = (int)(r / p.HW);
    const int hw = (int)(r - (int64_t)c * p.HW);
    const float x = p.x[e];
    float xp, p0;
    if (c < p.c_start) {
      p0 = x;  // ddim.py:246
      xp = x;  // ddim.py:266
    } else {
      const int64_t ea = ((int64_t)b * c_act + (c - p.c_start)) * p.HW + hw;
      float et = p.eps[ea];
      if (p.eps_uncond) et = cfg_mix(et, p.eps_uncond[ea], p.cfg_scale);
      float ep = et;
      if (p.plms_order > 0) {
        if (p.plms_mode == 1) {
          p.eps_save[ea] = et;  // e_t parked; provisional x_prev from e_t alone (plms.py:288)
        } else if (p.plms_mode == 2) {
          const float e0 = p.eps_save[ea];  // e_t of this step; `et` is e(x_prev, t_next)
          ep = __fdiv_rn(__fadd_rn(e0, et), 2.0f);  // plms.py:290
          p.hist[ea] = e0;                          // slot 0 (step 0)
        } else {
          const int nold = min(i, 3);
          const float* h1 = p.hist + (int64_t)((i + 2) % 3) * act_n;  // e_{i-1}
          const float* h2 = p.hist + (int64_t)((i + 1) % 3) * act_n;  // e_{i-2}
          float* h3 = p.hist + (int64_t)(i % 3) * act_n;              // e_{i-3}; receives e_i
          if (nold == 1) {
            ep = __fdiv_rn(__fsub_rn(__fmul_rn(3.0f, et), h1[ea]), 2.0f);
          } else if (nold == 2) {
            ep = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(23.0f, et), __fmul_rn(16.0f, h1[ea])), __fmul_rn(5.0f, h2[ea])), 12.0f);
          } else if (nold >= 3) {
            ep = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(55.0f, et), __fmul_rn(59.0f, h1[ea])),
                                               __fmul_rn(37.0f, h2[ea])), __fmul_rn(9.0f, h3[ea])), 24.0f);
          }
          h3[ea] = et;
        }
      }
      p0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(s1m, ep)), sqrt_at);  // ddim.py:243
      const float dir = __fmul_rn(dir_c, ep);                     // :258
      float nz = 0.f;
      if (p.noise) nz = p.noise[e];
      else if (sigma != 0.f) nz = philox_normal(p.seed ^ (p.seed_dev ? *p.seed_dev : 0ull), (uint64_t)i, (uint64_t)e);
      nz = __fmul_rn(__fmul_rn(sigma, nz), p.temperature);        // :260
      xp = __fadd_rn(__fadd_rn(__fmul_rn(sqrt_ap, p0), dir), nz); // :263
    }
    p.x_prev[e] = xp;
    if (p.x_dup) p.x_dup[e] = xp;
    if (p.pred_x0) p.pred_x0[e] = p0;
  }
}

// mask / x0 blend in front of a step (ddim.py:158-161): q_sample of the clean latent at this step's t, mixed in under the mask
__global__ void __launch_bounds__(256) blend_kernel(const FridoBlendParams p) {
  int i = *p.step;
  if (i > p.T - 1) i = p.T - 1;
  const int64_t t = p.t_table[i];
  const float a = p.sqrt_acp[t], b = p.sqrt_1m_acp[t];  // extract_into_tensor(...) (frido.py:306-307)
  const int64_t total = (int64_t)p.B * p.C * p.HW;
  const uint64_t seed = p.seed ^ (p.seed_dev ? *p.seed_dev : 0ull);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const float nz = p.noise ? p.noise[e] : philox_normal(seed, (1ull << 40) | (uint64_t)i, (uint64_t)e);
    const float orig = __fadd_rn(__fmul_rn(a, p.x0[e]), __fmul_rn(b, nz));
    const float m = p.mask[e];
    const float v = __fadd_rn(__fmul_rn(orig, m), __fmul_rn(__fsub_rn(1.0f, m), p.x[e]));
    p.x[e] = v;
    if (p.x_dup) p.x_dup[e] = v;
  }
}

__global__ void step_advance_kernel(int32_t* step) { *step += 1; }

// avg_pool2d(2) n times == mean over 2^n x 2^n blocks only up to rounding; the
// reference rounds after every level, so pool level by level in registers.
__global__ void snap_kernel(const FridoSnapParams p) {
  const int f = 1 << p.n;
  const int Hb = p.H / f, Wb = p.W / f;
  const int cn = p.c_end - p.c_start;
  const int64_t total = (int64_t)p.B * cn * Hb * Wb;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int wb = (int)(t % Wb);
  const int hb = (int)((t / Wb) % Hb);
  const int c = (int)((t / ((int64_t)Wb * Hb)) % cn) + p.c_start;
  const int b = (int)(t / ((int64_t)Wb * Hb * cn));
  float* base = p.x + (((int64_t)b * p.C + c) * p.H + (int64_t)hb * f) * p.W + (int64_t)wb * f;
  // hierarchical pooling, f <= 8 (n <= 3)
  float buf[64];
  for (int y = 0; y < f; ++y)
    for (int x = 0; x < f; ++x) buf[y * f + x] = base[(int64_t)y * p.W + x];
  for (int s = f; s > 1; s >>= 1) {
    const int h = s >> 1;
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < h; ++x) {
        // F.avg_pool2d sums the window then divides: ((a+b)+(c+d)) row-major order a,b,c,d
        const float a = buf[(2 * y) * f + 2 * x], bq = buf[(2 * y) * f + 2 * x + 1];
        const float cq = buf[(2 * y + 1) * f + 2 * x], d = buf[(2 * y + 1) * f + 2 * x + 1];
        const float sum = __fadd_rn(__fadd_rn(__fadd_rn(a, bq), cq), d);
        buf[y * f + x] = __fdiv_rn(sum, 4.0f);
      }
  }
  const float v = buf[0];
  for (int y = 0; y < f; ++y)
    for (int x = 0; x < f; ++x) base[(int64_t)y * p.W + x] = v;
}

// VQ: one thread per latent position, codebook (+ |e|^2) staged in shared memory.
__global__ void __launch_bounds__(256) vq_kernel(const FridoVqParams p) {
  extern __shared__ float sm[];
  float* cb = sm;                       // [n_e][e_dim]
  float* ee = sm + (size_t)p.n_e * p.e_dim;  // [n_e]
  for (int i = threadIdx.x; i < p.n_e * p.e_dim; i += blockDim.x) cb[i] = p.codebook[i];
  __syncthreads();
  for (int j = threadIdx.x; j < p.n_e; j += blockDim.x) {
    float s = 0.f;  // torch.sum(e**2, dim=1): sequential fp32 adds of rounded squares
    for (int d = 0; d < p.e_dim; ++d) s = __fadd_rn(s, __fmul_rn(cb[j * p.e_dim + d], cb[j * p.e_dim + d]));
    ee[j] = s;
  }
  __syncthreads();
  const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= (int64_t)p.B * p.HW) return;
  const int b = (int)(pos / p.HW);
  const int hw = (int)(pos - (int64_t)b * p.HW);
  float z[8];
  const float inv = __fdiv_rn(1.0f, p.scale_factor);  // frido.py:836 `1. / self.scale_factor[i]`
  float zz = 0.f;
  for (int d = 0; d < p.e_dim; ++d) {
    const float zin = p.z_nhwc ? p.z[pos * p.C_total + p.c_start + d] : p.z[((int64_t)b * p.C_total + p.c_start + d) * p.HW + hw];
    z[d] = __fmul_rn(zin, inv);
    zz = __fadd_rn(zz, __fmul_rn(z[d], z[d]));
  }
  float best = INFINITY;
  int bi = 0;
  for (int j = 0; j < p.n_e; ++j) {
    float dot = 0.f;  // K-ordered FMA chain, as an sgemm micro-kernel accumulates
    for (int d = 0; d < p.e_dim; ++d) dot = __fmaf_rn(z[d], cb[j * p.e_dim + d], dot);
    const float dist = __fsub_rn(__fadd_rn(zz, ee[j]), __fmul_rn(2.0f, dot));  // quantize.py:276-278
    if (dist < best) { best = dist; bi = j; }  // strict < keeps the first minimum (torch.argmin)
  }
  p.indices[pos] = bi;
  for (int d = 0; d < p.e_dim; ++d) {
    const float q = cb[bi * p.e_dim + d];
    p.out[pos * p.out_C + p.out_coff + d] = __fadd_rn(z[d], __fsub_rn(q, z[d]));  // quantize.py:294
  }
}

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = round_tf32(src[i]);
}

__global__ void split_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    uint32_t h;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(0.f), "f"(v));
    const float hf = __uint_as_float(h << 16);
    uint32_t l;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(0.f), "f"(v - hf));
    hi[i] = (uint16_t)(h & 0xFFFF);
    lo[i] = (uint16_t)(l & 0xFFFF);
  }
}

// ConvTranspose2d(k=4, s=2, p=1): out[oy,ox] += x[iy,ix] * w[ky,kx] with oy = 2*iy - 1 + ky  (tiny channel counts)
__global__ void convt_kernel(const FridoConvT2dParams p) {
  const int Ho = 2 * p.H, Wo = 2 * p.W;
  const int64_t total = (int64_t)p.B * Ho * Wo * p.Cout;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int co = (int)(e % p.Cout);
  int64_t r = e / p.Cout;
  const int ox = (int)(r % Wo); r /= Wo;
  const int oy = (int)(r % Ho);
  const int b = (int)(r / Ho);
  float acc = p.bias ? p.bias[co] : 0.f;
  const int ld = p.x_ld ? p.x_ld : p.Cin;
  // torch accumulates over ci, then ky, kx in its col2im GEMM; plain fp32 FMA chain here (parity ~1e-7)
  for (int ky = 0; ky < 4; ++ky) {
    const int ty = oy + 1 - ky;
    if (ty < 0 || (ty & 1)) continue;
    const int iy = ty >> 1;
    if (iy >= p.H) continue;
    for (int kx = 0; kx < 4; ++kx) {
      const int tx = ox + 1 - kx;
      if (tx < 0 || (tx & 1)) continue;
      const int ix = tx >> 1;
      if (ix >= p.W) continue;
      const float* xin = p.x + (((int64_t)b * p.H + iy) * p.W + ix) * ld;
      for (int ci = 0; ci < p.Cin; ++ci) acc = fmaf(xin[ci], p.w[((ci * p.Cout + co) * 4 + ky) * 4 + kx], acc);
    }
  }
  p.out[e] = acc;
}

__global__ void assemble_kernel(const FridoAssembleParams p) {
  const int64_t total = (int64_t)p.B * p.e * p.H * p.W;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int x = (int)(t % p.W);
  const int y = (int)((t / p.W) % p.H);
  const int d = (int)((t / ((int64_t)p.W * p.H)) % p.e);
  const int b = (int)(t / ((int64_t)p.W * p.H * p.e));
  const int hs = p.H >> p.sh, ws = p.W >> p.sh;
  const float v = p.h[(((int64_t)b * hs + (y >> p.sh)) * ws + (x >> p.sh)) * p.e + d];
  p.out[(((int64_t)b * p.C_total + p.c_off + d) * p.H + y) * p.W + x] = __fmul_rn(v, p.scale);
}

__global__ void to_uint8_kernel(const FridoToU8Params p) {
  const int64_t total = (int64_t)p.B * p.HW * p.C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % p.C);
    const int64_t r = e / p.C;
    const int hw = (int)(r % p.HW);
    const int b = (int)(r / p.HW);
    const float x = p.x[((int64_t)b * p.C + c) * p.HW + hw];
    p.out[e] = format_u8(x, p.mode);
  }
}

}  // namespace frido

using namespace frido;

extern "C" int frido_conv_transpose2d(const FridoConvT2dParams* p, void* stream) {
  if (!p || !p->x || !p->w || !p->out || p->B <= 0 || p->Cin <= 0 || p->Cout <= 0) return set_error(FRIDO_E_ARG, "conv_transpose2d: bad argument");
  const int64_t total = (int64_t)p->B * 4 * p->H * p->W * p->Cout;
  convt_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("conv_transpose2d");
}

extern "C" int frido_assemble_latent(const FridoAssembleParams* p, void* stream) {
  if (!p || !p->h || !p->out || p->sh < 0 || (p->H & ((1 << p->sh) - 1)) || (p->W & ((1 << p->sh) - 1)))
    return set_error(FRIDO_E_ARG, "assemble_latent: bad argument");
  const int64_t total = (int64_t)p->B * p->e * p->H * p->W;
  assemble_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("assemble_latent");
}

extern "C" int frido_to_uint8(const FridoToU8Params* p, void* stream) {
  if (!p || !p->x || !p->out || p->B <= 0 || p->C <= 0 || p->HW <= 0 || (p->mode != 0 && p->mode != 1))
    return set_error(FRIDO_E_ARG, "to_uint8: bad argument");
  const int64_t total = (int64_t)p->B * p->HW * p->C;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  to_uint8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("to_uint8");
}

extern "C" int frido_split_bf16(const float* src, void* hi, void* lo, int64_t n, void* stream) {
  if (!src || !hi || !lo || n < 0) return set_error(FRIDO_E_ARG, "split_bf16: bad argument");
  if (n == 0) return FRIDO_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (uint16_t*)hi, (uint16_t*)lo, n);
  return check_launch("split_bf16");
}

extern "C" int frido_step_begin(const FridoStepBeginParams* p, void* stream) {
  if (!p || !p->step || !p->t_table || !p->ts || p->B <= 0 || p->T <= 0) return set_error(FRIDO_E_ARG, "step_begin: bad argument");
  step_begin_kernel<<<(p->B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("step_begin");
}

extern "C" int frido_sampler_update(const FridoUpdateParams* p, void* stream) {
  if (!p || !p->x || !p->eps || !p->coef || !p->step || !p->x_prev) return set_error(FRIDO_E_ARG, "update: null pointer");
  if (p->c_start < 0 || p->c_end <= p->c_start || p->B <= 0 || p->HW <= 0) return set_error(FRIDO_E_ARG, "update: bad shape");
  if (p->plms_order > 0 && !p->hist) return set_error(FRIDO_E_ARG, "update: PLMS needs hist");
  if (p->plms_order > 0 && p->plms_mode != 0 && !p->eps_save) return set_error(FRIDO_E_ARG, "update: PLMS first step needs eps_save");
  const int64_t total = (int64_t)p->B * p->c_end * p->HW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  update_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  int rc = check_launch("sampler_update");
  if (rc) return rc;
  if (p->advance) {
    step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p->step);
    rc = check_launch("step_advance");
  }
  return rc;
}

extern "C" int frido_mask_blend(const FridoBlendParams* p, void* stream) {
  if (!p || !p->x || !p->x0 || !p->mask || !p->sqrt_acp || !p->sqrt_1m_acp || !p->step || !p->t_table)
    return set_error(FRIDO_E_ARG, "mask_blend: null pointer");
  if (p->B <= 0 || p->C <= 0 || p->HW <= 0 || p->T <= 0) return set_error(FRIDO_E_ARG, "mask_blend: bad shape");
  const int64_t total = (int64_t)p->B * p->C * p->HW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  blend_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("mask_blend");
}

extern "C" int frido_stage_snap(const FridoSnapParams* p, void* stream) {
  if (!p || !p->x || p->n < 0 || p->n > 3) return set_error(FRIDO_E_ARG, "snap: bad argument (n<=3)");
  if (p->n == 0) return FRIDO_OK;
  const int f = 1 << p->n;
  if (p->H % f || p->W % f) return set_error(FRIDO_E_ARG, "snap: H,W must be divisible by 2^n");
  const int64_t total = (int64_t)p->B * (p->c_end - p->c_start) * (p->H / f) * (p->W / f);
  snap_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("stage_snap");
}

extern "C" int frido_vq_lookup(const FridoVqParams* p, void* stream) {
  if (!p || !p->z || !p->codebook || !p->out || !p->indices) return set_error(FRIDO_E_ARG, "vq: null pointer");
  if (p->e_dim <= 0 || p->e_dim > 8 || p->n_e <= 0) return set_error(FRIDO_E_ARG, "vq: e_dim must be 1..8");
  const size_t smem = ((size_t)p->n_e * p->e_dim + p->n_e) * sizeof(float);
  if (smem > 200 * 1024) return set_error(FRIDO_E_ARG, "vq: codebook exceeds shared memory");
  static DevOnce attr_set;
  if (attr_set.need()) cudaFuncSetAttribute(vq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int64_t n = (int64_t)p->B * p->HW;
  vq_kernel<<<(unsigned)((n + 255) / 256), 256, smem, (cudaStream_t)stream>>>(*p);
  return check_launch("vq_lookup");
}

extern "C" int frido_round_tf32(const float* src, float* dst, int64_t n, void* stream) {
  if (!src || !dst || n < 0) return set_error(FRIDO_E_ARG, "round_tf32: bad argument");
  if (n == 0) return FRIDO_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  round_tf32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  return check_launch("round_tf32");
}
