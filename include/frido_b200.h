/* frido_b200 — C ABI of the B200-native Frido sampling hot path.
 *
 * The reference (davidhalladay/Frido) has no FFI: its "operators" are stock
 * PyTorch calls made from Python classes resolved out of YAML `target:` strings
 * (SURVEY.md §8b).  This header is the boundary a maintainer binds instead: one
 * launcher per fused device op of the path, plain pointers + sizes, no torch
 * types.  Every launcher enqueues on `stream` (a cudaStream_t passed as void*),
 * never synchronises, never allocates, and returns 0 or a negative FRIDO_E_*.
 * All pointers are DEVICE pointers unless stated.  Activations are fp32 NHWC
 * ("[B,H,W,C]"; tokens "[B,N,C]" are the same memory); latents at the sampler
 * level are fp32 NCHW exactly as the reference's tensors.
 *
 * Each entry cites the reference code it replaces (paths into the reference).
 */
#ifndef FRIDO_B200_H
#define FRIDO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRIDO_ABI_VERSION 7
#define FRIDO_SK_WS_BYTES (40ll << 20)

#define FRIDO_OK 0
#define FRIDO_E_ARG (-1)     /* bad argument / unsupported shape */
#define FRIDO_E_LAUNCH (-2)  /* CUDA launch error (see frido_last_error) */
#define FRIDO_E_ARCH (-3)    /* device is not sm_100 */

/* activation codes for FridoConvParams.act */
#define FRIDO_ACT_NONE 0
#define FRIDO_ACT_RELU 1
#define FRIDO_ACT_SILU 2
#define FRIDO_ACT_GEGLU 3 /* columns (2j,2j+1) = (value,gate) -> out[j] = value*gelu_erf(gate) */
#define FRIDO_ACT_GELU 5  /* exact-erf GELU (x_transformer.py:199-202 FeedForward of the condition encoder) */
#define FRIDO_ACT_GEGLU_FAST 4 /* same with erf by Abramowitz-Stegun 7.1.26 (|err| < 5e-7); tcgen05 engines only */

/* ---------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear / batched matmul.
 *   out[b,p,n] = act( alpha * sum_{tap,c} A[b, pix(p,tap), c] * W[n, tap, c]
 *                     + bias[n] + rowvec[b,n] + res[b,p,n] )
 * Replaces nn.Conv2d 3x3/1x1 (pyunet.py:211,237,248,154,110; spade_norm.py:37-42;
 * taming/modules/diffusionmodules/model.py:43,88,98,570,612; msvqgan.py:75),
 * nn.Linear (pyunet.py:561-565,225-231; attention.py:40,60,161-168), the
 * channel concat th.cat([h, hs.pop()],1) (pyunet.py:939, via a0|a1), nearest x2
 * upsampling (pyunet.py:119, via ups=2), SPADE's nearest down-resize
 * (spade_norm.py:52, via scaled strides) and the attention einsums
 * (attention.py:180-191; model.py:176-188, via per-image weights w_sb != 0).
 * ------------------------------------------------------------------------- */
typedef struct FridoConvParams {
  const float* a0;          /* source 0 */
  const float* a1;          /* source 1 (channel-concatenated after source 0) or NULL */
  int32_t c0, c1;           /* channels taken from each source; K per tap = c0 + c1 */
  int64_t a0_sb, a0_sy, a0_sx, a0_sc; /* element strides: image, row, col, channel */
  int64_t a1_sb, a1_sy, a1_sx, a1_sc;
  int32_t B, Hin, Win;      /* source grid (before the optional x2 upsample) */
  int32_t ups;              /* 1, or 2 = nearest-neighbour x2 folded into addressing */
  int32_t ksize, stride, pad;
  int32_t Hout, Wout;
  const float* w;           /* [Cout][ksize*ksize][c0+c1], K contiguous */
  int64_t w_sb;             /* 0 = shared weights; else per-image weight stride (batched matmul) */
  int64_t w_ld;             /* element stride between weight rows; 0 = dense (ksize*ksize*(c0+c1)) */
  const void* w_lo;         /* engine 3 only: `w` is then the bf16 HIGH part and w_lo the bf16 LOW part of the
                               weights (W = hi + lo, same [Cout][K] layout, see frido_split_bf16); else NULL */
  int32_t Cout;
  const float* bias;        /* [Cout] or NULL */
  const float* rowvec;      /* per-image vector (timestep embedding) or NULL */
  int64_t rowvec_sb;
  const float* res;         /* residual, addressed like out, or NULL */
  float alpha;
  int32_t act;
  float* out;
  int64_t o_sb, o_sp, o_sn; /* out[b*o_sb + p*o_sp + n*o_sn], p = oy*Wout + ox */
  int32_t round_tf32;       /* round stored values to TF32 (rna) */
  void* out_hi; void* out_lo; /* optional: also store the outputs as a bf16 pair (hi = bf16(v), lo = bf16(v - hi)), indexed like
                               `out`; this is how activations that later act as the W operand of a BF16x3 matmul
                               (attention K and V^T) get their pre-split copy.  Not with GEGLU. */
  double* chan_sums;        /* optional (tcgen05 engines only): [B][Cout][2] += per-channel (sum, sum of squares) of the
                               stored outputs, i.e. the GroupNorm statistics of the tensor being produced; zero on entry */
  const float* x0;          /* optional (tcgen05 engines only) fused 1x1 side input, e.g. a ResBlock's skip_connection conv */
  const float* x1;          /*   (pyunet.py:248,299): out += sum_c X[b,p,c] * W[n, ksize*ksize*(c0+c1) + c], X = x0|x1 channel- */
  int32_t cx0, cx1;         /*   concatenated, sampled at the output pixel (stride 1 only); the weight rows carry the extra */
  int64_t x0_sb, x0_sy, x0_sx; /* cx0+cx1 columns after the taps.  Element strides: image, row, col (channel stride 1). */
  int64_t x1_sb, x1_sy, x1_sx;
  const float* nrm_ab;      /* optional (engine 3, 3x3 / 1x1 stride-1 convs with shared weights): NORMALISE-ON-LOAD.  The conv reads */
  const float* nrm_gb;      /*   act(x * a[b,c] + b[b,c]) instead of x, where nrm_ab = [B][c0+c1][2] holds (a, b) = (rstd*gamma_c, */
  int32_t nrm_silu;         /*   beta_c - mean*rstd*gamma_c) per image and input channel (frido_gn_finalize), nrm_gb = optional SPADE
                               maps [B,Hin*Win,2(c0+c1)] (gamma | beta: y = y*(1+gamma)+beta, spade_norm.py:58) and nrm_silu
                               selects SiLU: GroupNorm -> [SPADE] -> SiLU -> conv (pyunet.py:209-240) in one launch, the
                               normalised tensor never written to memory.  Zero padding applies to the ACTIVATED tensor, as
                               in the reference.  The fused side input x0 | x1 stays raw. */
  int32_t a_presplit;       /* engine 3, 3x3 stride-1 convs: a0 | a1 already hold the operand in the engine's split form (per 16 bytes =
                               4 channels: bf16x2 hi(c0,c1), hi(c2,c3), lo(c0,c1), lo(c2,c3), what frido_norm_act /
                               frido_upsample2x write with out_split = 1; same size as the fp32 tensor).  The conv then runs
                               the halo-resident operand path of csrc/conv_nf.cu with no per-element work at all. */
  uint8_t* out_u8;          /* optional (SIMT engine, Cout <= 4: the decoder's conv_out head): also store the outputs as uint8 NHWC */
  int32_t u8_mode;          /*   [B,Hout,Wout,Cout], formatted like frido_to_uint8 (mode 0 = custom_to_np, 1 = custom_to_pil,
                               scripts/sample_diffusion.py:103-121) from the same fp32 value that goes to `out` */
  void* sk_ws;              /* optional (tcgen05 engines only) stream-K workspace: launches with too few output tiles for
                               the 148 SMs split their K loops across CTAs and combine the partial sums here (in a fixed
                               order: results stay deterministic).  Zero it once; launches on one stream may share it. */
  int64_t sk_ws_bytes;      /* 40 MiB covers every shape (FRIDO_SK_WS_BYTES) */
  int32_t engine;           /* 0 = SIMT fp32; 1 = tcgen05 TF32; 2 = tcgen05 3xTF32 (error-compensated, ~2^-21 products);
                               3 = tcgen05 BF16x3 (error-compensated, ~2^-16 products, pre-split weights, 2x the TF32
                               issue rate); 1-3 take aligned shapes only (see csrc/conv_tc.cu) */
} FridoConvParams;

int frido_conv2d(const FridoConvParams* p, void* stream);

/* GroupNorm statistics over an NHWC tensor that may be the concat of two
 * sources: per (image, group) sum and sum of squares in fp64.
 * Replaces the reduction half of nn.GroupNorm(32) (util.py:214; attention.py:76;
 * taming model.py:34).  `sums` [B][groups][2] must be zero on entry.
 * groups = 0: per-CHANNEL sums of a single source instead, `sums` [B][c0][2] - the format FridoConvParams.chan_sums
 * produces and FridoNormActParams.csum0/csum1 consume (for tensors whose producer ran on the SIMT engine). */
typedef struct FridoGnStatsParams {
  const float* a0; const float* a1; int32_t c0, c1;
  int32_t B, HW, groups;
  double* sums;
} FridoGnStatsParams;
int frido_gn_stats(const FridoGnStatsParams* p, void* stream);

/* Normalise + affine [+ SPADE modulation] [+ SiLU], writing one NHWC tensor.
 *   y = ((x-mean)*rstd*gamma[c]+beta[c]);  if gb: y = y*(1+gb[b,p,c]) + gb[b,p,C+c];
 *   if silu: y = y*sigmoid(y)
 * Replaces the apply half of GroupNorm, SPADE.forward's modulation
 * (spade_norm.py:58) and nn.SiLU / swish (pyunet.py:210,234; model.py:29). */
typedef struct FridoNormActParams {
  const float* a0; const float* a1; int32_t c0, c1;
  int32_t B, HW, groups;
  const double* sums; float eps;          /* per-(image,group) sums from frido_gn_stats, or NULL when csum0 is given */
  const double* csum0; const double* csum1; /* alternative: per-channel sums [B][c0][2] / [B][c1][2] accumulated by the
                                               producing conv (FridoConvParams.chan_sums) */
  const float* gamma; const float* beta;  /* [C] */
  const float* gb;                        /* [B,HW,2C] SPADE (gamma|beta) or NULL */
  int32_t silu;
  int32_t round_tf32;
  float* out;                             /* [B,HW,C] */
  int32_t out_split;                      /* 1: write each group of 4 channels as the BF16x3 engine's operand (bf16x2 hi, hi, lo,
                                             lo; see FridoConvParams.a_presplit) instead of 4 floats: the consuming conv then
                                             needs no operand split */
} FridoNormActParams;
int frido_norm_act(const FridoNormActParams* p, void* stream);

/* GroupNorm statistics -> per-(image, channel) scale and shift for normalise-on-load (FridoConvParams.nrm_ab):
 *   mean_g, rstd_g from `sums` (per group, frido_gn_stats) or from the producers' per-channel sums csum0 | csum1
 *   (FridoConvParams.chan_sums), in fp64 exactly as frido_norm_act does;  ab[b][c] = (rstd_g*gamma_c, beta_c - mean_g*rstd_g*gamma_c).
 * Replaces the statistics half of nn.GroupNorm(32) (util.py:214) for convs that apply the normalisation themselves. */
typedef struct FridoGnFinalizeParams {
  int32_t c0, c1, B, HW, groups; float eps;
  const double* sums; const double* csum0; const double* csum1;
  const float* gamma; const float* beta;
  float* ab;                /* [B][c0+c1][2] */
} FridoGnFinalizeParams;
int frido_gn_finalize(const FridoGnFinalizeParams* p, void* stream);

/* nn.LayerNorm over the last dim (attention.py:203-205), eps 1e-5. */
typedef struct FridoLayerNormParams {
  const float* x; int64_t rows; int32_t C; float eps;
  const float* gamma; const float* beta; int32_t round_tf32; float* out;
  void* out_hi; void* out_lo; /* optional bf16 pair copy of the output (see FridoConvParams.out_hi): the LayerNorm output is
                                 the K operand of the folded self-attention score matmul */
} FridoLayerNormParams;
int frido_layernorm(const FridoLayerNormParams* p, void* stream);

/* Row softmax of scale*S (attention.py:180,189; model.py:180-181), in place allowed. */
typedef struct FridoSoftmaxParams {
  const float* s; int64_t rows; int32_t n; int64_t ld; float scale; int32_t round_tf32; float* out;
} FridoSoftmaxParams;
int frido_softmax(const FridoSoftmaxParams* p, void* stream);

/* timestep_embedding (util.py:151-171): out[b] = [cos(t_b f_j), sin(t_b f_j)], j < dim/2. */
typedef struct FridoTimeEmbedParams {
  const int64_t* t; int32_t B; int32_t dim; float max_period;
  const float* freqs; /* optional [dim/2] host-computed f_j (bit-identical to the reference's torch.exp); NULL = compute */
  float* out;
} FridoTimeEmbedParams;
int frido_time_embed(const FridoTimeEmbedParams* p, void* stream);

/* Sampler state kept on the device so that one captured step can be replayed:
 *   step[0] = running step counter i (0..T-1) of the current stage.
 * frido_step_begin writes ts[b] = t_table[i] (ddim.py:157). */
typedef struct FridoStepBeginParams {
  const int32_t* step; const int64_t* t_table; int32_t use_next; int32_t T; int64_t* ts; int32_t B;
} FridoStepBeginParams;
int frido_step_begin(const FridoStepBeginParams* p, void* stream);

/* Fused DDIM / PLMS x_{t-1} update (ddim.py:200-268; plms.py:247-301):
 * zero-pad eps to the frozen groups, optional classifier-free guidance
 * (ddim.py:226), optional Adams-Bashforth combination of the eps history
 * (plms.py:285-299), x0 prediction, direction, sigma*noise, group masking, and
 * (single thread) step counter increment.  NCHW fp32.
 *   coef[i] = {a_t, a_prev, sigma_t, sqrt(1-a_t)} for step i (index = T-1-i). */
typedef struct FridoUpdateParams {
  const float* x;            /* [B, c_end, H, W] */
  const float* eps;          /* [B, c_act, H, W] conditional (or only) model output */
  const float* eps_uncond;   /* same shape or NULL */
  float cfg_scale;
  int32_t B, c_start, c_end, HW;
  const float* coef;         /* [T][4] */
  int32_t* step;             /* device step counter */
  int32_t advance;           /* 1: increment *step after the update */
  int32_t plms_order;        /* 0 = DDIM; 1..4 = multistep order cap (uses min(order, i+1) unless forced) */
  int32_t plms_mode;         /* 0 plain, 1 = "first half" (x_prev from e_t only, history untouched),
                                2 = "second half" (e' = (hist_new + eps)/2 where eps is e(t_next)) */
  float* hist;               /* [3][B, c_act, H, W] eps ring (PLMS) or NULL */
  float* eps_save;           /* PLMS mode 1: where e_t is parked; mode 2: read back */
  const float* noise;        /* injected N(0,1) [B, c_end, H, W] or NULL */
  uint64_t seed;             /* Philox seed used when noise == NULL and sigma != 0 */
  const uint64_t* seed_dev;  /* optional device word XORed into seed (fresh per sample() call; graph-safe) */
  float temperature;
  float* x_prev;             /* [B, c_end, H, W] (may alias x) */
  float* x_dup;              /* optional second copy of x_prev (CFG 2B batch) or NULL */
  float* pred_x0;            /* optional or NULL */
} FridoUpdateParams;
int frido_sampler_update(const FridoUpdateParams* p, void* stream);

/* Inpainting / img2img blend in front of a sampler step (ddim.py:158-161; plms.py:162-165):
 *   t = t_table[*step];  img_orig = sqrt_acp[t]*x0 + sqrt_1m_acp[t]*noise   (q_sample, frido.py:302-307)
 *   x = img_orig*mask + (1 - mask)*x          -- same fp32 operation order as the reference, no FMA contraction.
 * NCHW fp32; `mask` is already expanded to x's shape.  noise == NULL draws N(0,1) from Philox (the reference's
 * torch.randn_like cannot be matched bit for bit; parity runs inject the noise on both sides). */
typedef struct FridoBlendParams {
  float* x;                  /* [B, C, H, W], blended in place */
  float* x_dup;              /* optional second copy of the result (CFG 2B batch) or NULL */
  const float* x0;           /* [B, C, H, W] clean latent */
  const float* mask;         /* [B, C, H, W], 1 = keep x0 */
  const float* noise;        /* injected N(0,1) [B, C, H, W] or NULL */
  const float* sqrt_acp;     /* [num_timesteps] sqrt_alphas_cumprod (frido.py:150) */
  const float* sqrt_1m_acp;  /* [num_timesteps] sqrt_one_minus_alphas_cumprod (frido.py:151) */
  const int32_t* step; const int64_t* t_table; int32_t T;
  int32_t B, C, HW;
  uint64_t seed; const uint64_t* seed_dev;
} FridoBlendParams;
int frido_mask_blend(const FridoBlendParams* p, void* stream);

/* Inter-stage snap (ddim.py:177-185): avg_pool2d(2) n times then nearest x2 n times
 * on channels [c_start,c_end) of an NCHW tensor, in place. */
typedef struct FridoSnapParams {
  float* x; int32_t B, C, H, W, c_start, c_end, n;
} FridoSnapParams;
int frido_stage_snap(const FridoSnapParams* p, void* stream);

/* decode_first_stage rescale + VectorQuantizer2.forward for one scale
 * (frido.py:832-838; quantize.py:272-297): z*(1/scale) -> argmin_j (|z|^2+|e_j|^2-2 z.e_j)
 * (first minimum) -> z + (e_idx - z), written NHWC at channel offset out_coff of
 * a [B,HW,out_C] tensor (msvqgan.py:392-393 reverses the group order). */
typedef struct FridoVqParams {
  const float* z;            /* NCHW [B, C_total, H, W]  (z_nhwc = 0)  or NHWC [B, HW, C_total] (z_nhwc = 1) */
  int32_t B, C_total, HW, c_start, e_dim;
  float scale_factor;
  const float* codebook;     /* [n_e][e_dim] */
  int32_t n_e;
  float* out; int32_t out_C, out_coff;
  int64_t* indices;          /* [B*HW] */
  int32_t z_nhwc;            /* input layout, see `z` (the encoder's top-down path quantises NHWC conv outputs) */
} FridoVqParams;
int frido_vq_lookup(const FridoVqParams* p, void* stream);

/* F.interpolate(scale_factor=2, mode="nearest") on an NHWC tensor (pyunet.py:119; taming
 * model.py:50) — used in front of the tcgen05 engine (the SIMT engine folds it into addressing). */
typedef struct FridoUpsampleParams { const float* x; int32_t B, H, W, C; int32_t round_tf32; float* out; int32_t out_split; } FridoUpsampleParams;
int frido_upsample2x(const FridoUpsampleParams* p, void* stream);

/* Condition encoder (BERTEmbedder = x-transformer encoder, frido/modules/x_transformer.py).
 * Token + absolute position embedding (x_transformer.py:34-36,608-609): out[b,l,:] = tok_emb[tokens[b,l]] + pos_emb[l]. */
typedef struct FridoEmbedParams {
  const int64_t* tokens; int32_t B, L, D; int32_t vocab; const float* tok_emb; const float* pos_emb; float* out;
} FridoEmbedParams;
int frido_embed_tokens(const FridoEmbedParams* p, void* stream);

/* Multi-head self-attention for short sequences (x_transformer.py:268-366, no mask, no memory): qkv [B,L,3*H*Dh] holds
 * q | k | v (each H*Dh wide, head-major); out[b,l,h*Dh+d] = sum_j softmax_j(scale * q.k_j) v_j.  One CTA per (b,h),
 * K and V staged in shared memory; L*Dh*8 bytes must fit in 200 KB. */
typedef struct FridoMhaParams {
  const float* qkv; int32_t B, L, H, Dh; float scale; float* out;
} FridoMhaParams;
int frido_mha_small(const FridoMhaParams* p, void* stream);

/* Fused single-head attention for short key sequences (attention.py:178-191: einsum -> *scale -> softmax -> einsum):
 *   q' = LayerNorm(q[b,n,:]) if ln_gamma else q[b,n,:]                                  (attention.py:323-325)
 *   out[b,n,:] = softmax_j(scale * q'.k[b,j,:]) @ v[b,j,:] + bias + res[b,n,:],  j < Nk, fp32 throughout, scores on chip.
 * Used for the cross-attention to a short condition (26 layout tokens) and for self-attention with < 128 tokens per
 * image (the 8x8 level), where the tcgen05 engine has no 128-row tile to work on.  For the cross-attention the host
 * folds the step-invariant projections into the operands (k = (ctx Wk^T) Wq, v = (ctx Wv^T) Wo^T), so that with the
 * LayerNorm, to_out bias and residual fused here  x + to_out(attn(LN(x), ctx))  (attention.py:324) is this one launch.
 * Row strides in floats; C <= 1024. */
typedef struct FridoAttnParams {
  const float* q; int64_t q_sb, q_ld;
  const float* k; int64_t k_sb, k_ld;
  const float* v; int64_t v_sb, v_ld;
  int32_t B, N, Nk, C; float scale;
  float* out; int64_t o_sb, o_ld;
  const float* ln_gamma; const float* ln_beta; float ln_eps; /* optional LayerNorm of the query rows ([C] each) */
  const float* bias;                                          /* optional [C] added to every output row */
  const float* res; int64_t r_sb, r_ld;                       /* optional residual, addressed like out */
  const float* ln2_gamma; const float* ln2_beta; float ln2_eps; /* optional: also write LayerNorm(out row) ... */
  float* out2;                                                /* ... here, addressed like out (the block's next norm) */
} FridoAttnParams;
int frido_attn_small(const FridoAttnParams* p, void* stream);

/* Streaming-softmax attention on the tcgen05 tensor cores (csrc/attn_flash.cu), one head of C channels:
 *   out = res + bias + softmax(scale * Q K^T) V        (attention.py:170-193; taming model.py:166-192)
 * in ONE launch - scores in tensor memory, running row maximum / sum in registers, P handed back to the tensor core
 * through shared memory; the [B,N,N] score tensor never exists.  All three matmul operands arrive in the BF16x3 engine's
 * pre-split form (bf16 hi / lo pairs, what frido_conv2d's out_hi / out_lo and frido_layernorm's out_hi / out_lo write):
 *   q, k  : [B][N][C]   rows of C bf16, row stride q_ld / k_ld, image stride q_sb / k_sb (elements)
 *   vt    : V transposed, channel-major: element (b, c, key) at  b * vt_sb + c * vt_ld + key
 * N % 128 == 0, C % 32 == 0.  out / res: fp32 [B][N][C] with row stride o_ld / r_ld. */
typedef struct FridoFlashParams {
  const void* q_hi; const void* q_lo; int64_t q_sb, q_ld;
  const void* k_hi; const void* k_lo; int64_t k_sb, k_ld;
  const void* vt_hi; const void* vt_lo; int64_t vt_sb, vt_ld;
  int32_t B, N, C; float scale;
  const float* bias;                                          /* optional [C] */
  const float* res; int64_t r_sb, r_ld;                       /* optional residual */
  float* out; int64_t o_sb, o_ld;
} FridoFlashParams;
int frido_attn_flash(const FridoFlashParams* p, void* stream);
int frido_attn_flash_eligible(int32_t B, int32_t N, int32_t C);

/* MS-VQGAN encode side.
 * nn.ConvTranspose2d(Cin, Cout, 4, stride=2, padding=1) on NHWC (msvqgan.py:82-84): out [B,2H,2W,Cout] dense;
 * w is the PyTorch layout [Cin][Cout][4][4]. */
typedef struct FridoConvT2dParams {
  const float* x; int32_t B, H, W, Cin, Cout; const float* w; const float* bias; float* out;
  int32_t x_ld;              /* floats between consecutive input pixels (>= Cin; 0 = Cin): lets x be a channel slice */
} FridoConvT2dParams;
int frido_conv_transpose2d(const FridoConvT2dParams* p, void* stream);

/* Assemble one scale into the latent (msvqgan.py:355-372 + frido.py:646-662): nearest-upsample an NHWC map
 * [B,H>>sh,W>>sh,e] by 2^sh, multiply by `scale`, write channels [c_off, c_off+e) of the NCHW latent [B,C_total,H,W]. */
typedef struct FridoAssembleParams {
  const float* h; int32_t B, H, W, e, sh; float scale; float* out; int32_t C_total, c_off;
} FridoAssembleParams;
int frido_assemble_latent(const FridoAssembleParams* p, void* stream);

/* Output formatting of the sampling script, fused into one pass: fp32 NCHW image in [-1,1] -> uint8 NHWC.
 *   mode 0 = custom_to_np  (scripts/sample_diffusion.py:115-121): ((x + 1) * 127.5).clamp(0, 255) -> uint8 (truncate)
 *   mode 1 = custom_to_pil (scripts/sample_diffusion.py:103-108): 255 * ((clamp(x,-1,1) + 1) / 2)  -> uint8 (truncate)
 * Same fp32 operation order as the reference, so the bytes are identical. */
typedef struct FridoToU8Params { const float* x; int32_t B, C, HW; int32_t mode; uint8_t* out; } FridoToU8Params;
int frido_to_uint8(const FridoToU8Params* p, void* stream);

/* Fill `n` bytes with zero (graph-capturable helper for the GN sums). */
int frido_zero(void* ptr, int64_t nbytes, void* stream);

/* BF16x3 weight packing: hi = bf16_rn(w), lo = bf16_rn(w - hi), element-wise (device pointers). */
int frido_split_bf16(const float* src, void* hi, void* lo, int64_t n, void* stream);

/* tcgen05 engine weight packing: W[Cout][K] fp32 -> TF32-rounded (rna) copy.
 * (The layout itself is unchanged; TMA tiles it.) */
int frido_round_tf32(const float* src, float* dst, int64_t n, void* stream);

/* ---------------------------------------------------------------------------
 * Weight packing (pack_weights / workspace_bytes).  Checkpoint tensors keep the reference's layouts
 * (Conv2d OIHW, Linear [out,in]); the engines read K-major rows [C_out][tap][C_in] with fused operands concatenated.
 * Every transformation the host runtime applies is one of these launches (csrc/pack.cu), all on device pointers.
 * ------------------------------------------------------------------------- */
/* dst[a*d0 + b*d1 + c*d2] = src[a*s0 + b*s1 + c*s2] for a < n0, b < n1, c < n2 (strides in floats): permutes,
 * concatenation along rows / columns (offset dst), transposes, GEGLU (value_j, gate_j) row interleave. */
int frido_pack_permute3(const float* src, int64_t s0, int64_t s1, int64_t s2, float* dst, int64_t d0, int64_t d1, int64_t d2,
                        int32_t n0, int32_t n1, int32_t n2, void* stream);
/* Conv2d weight OIHW -> [O][KH*KW][I] rows of length dst_ld >= KH*KW*I (pyunet.py / taming model.py convs). */
int frido_pack_conv_weight(const float* w, int32_t O, int32_t I, int32_t KH, int32_t KW, float* dst, int64_t dst_ld, void* stream);
/* out[m][n] = sum_k a[m*a_rs + k*a_cs] * b[k*b_rs + n*b_cs], fp64 products and sums in k order, rounded to fp32 once. */
int frido_matmul_f64acc(const float* a, int64_t a_rs, int64_t a_cs, const float* b, int64_t b_rs, int64_t b_cs, int32_t M, int32_t N,
                        int32_t K, float* out, int64_t o_ld, void* stream);
/* Self-attention weight folds of one CrossAttention module (attention.py:172-191 re-associated; all [C][C] row-major):
 *   a_out = Wk^T Wq   (sim = (x a_out^T) x^T: the keys are the tokens themselves)
 *   wv_out = Wo Wv    (the values carry to_out) */
int frido_fold_self_attention(const float* wq, const float* wk, const float* wv, const float* wo, int32_t C, float* a_out,
                              float* wv_out, void* stream);
int frido_vec_add(const float* a, const float* b, float* out, int64_t n, void* stream);

/* ---------------------------------------------------------------------------
 * Native op-program executor: the host runtime builds a flat array of ops once
 * per (stage, batch) and replays it (inside a CUDA graph) every step.
 * ------------------------------------------------------------------------- */
enum FridoOpKind {
  FRIDO_OP_CONV = 1, FRIDO_OP_GN_STATS = 2, FRIDO_OP_NORM_ACT = 3, FRIDO_OP_LAYERNORM = 4,
  FRIDO_OP_SOFTMAX = 5, FRIDO_OP_TIME_EMBED = 6, FRIDO_OP_STEP_BEGIN = 7, FRIDO_OP_UPDATE = 8,
  FRIDO_OP_SNAP = 9, FRIDO_OP_VQ = 10, FRIDO_OP_ZERO = 11, FRIDO_OP_UPSAMPLE = 12,
  FRIDO_OP_EMBED = 13, FRIDO_OP_MHA = 14, FRIDO_OP_CONVT = 15, FRIDO_OP_ASSEMBLE = 16, FRIDO_OP_ATTN = 17,
  FRIDO_OP_BLEND = 18, FRIDO_OP_GN_FINALIZE = 19, FRIDO_OP_FLASH = 20
};
typedef struct FridoZeroParams { void* ptr; int64_t nbytes; } FridoZeroParams;
typedef struct FridoOp {
  int32_t kind;
  int32_t tag; /* free for the host (profiling label index) */
  union {
    FridoConvParams conv; FridoGnStatsParams gn_stats; FridoNormActParams norm_act;
    FridoLayerNormParams layernorm; FridoSoftmaxParams softmax; FridoTimeEmbedParams time_embed;
    FridoStepBeginParams step_begin; FridoUpdateParams update; FridoSnapParams snap; FridoVqParams vq;
    FridoZeroParams zero; FridoUpsampleParams upsample; FridoEmbedParams embed; FridoMhaParams mha;
    FridoConvT2dParams convt; FridoAssembleParams assemble; FridoAttnParams attn; FridoBlendParams blend; FridoGnFinalizeParams gn_finalize; FridoFlashParams flash;
  } u;
} FridoOp;
/* Launches ops[0..n) in order on `stream`; returns 0 or (-(1000+i)) if op i failed. */
int frido_run_program(const FridoOp* ops, int32_t n, void* stream);

/* Bytes of scratch (FridoConvParams.sk_ws) the ops of a program need; the caller allocates it (a torch tensor in the
 * Python host) and passes it in every tcgen05 conv.  ops == NULL: the upper bound for any program. */
int64_t frido_workspace_bytes(const FridoOp* ops, int32_t n);

int frido_abi_version(void);
int frido_sizeof_op(void);
const char* frido_last_error(void);
/* number of kernel launches issued through this library since load (bench's gpu_launches) */
int64_t frido_launch_count(void);
/* 0 if the current device is sm_100 and the tcgen05 engine is usable */
int frido_check_device(void);

#ifdef __cplusplus
}
#endif
#endif
