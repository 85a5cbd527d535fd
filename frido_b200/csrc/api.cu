// C-ABI glue: conv engine selection, the op-program executor, misc entry points.
#include "common.cuh"

namespace frido {
char g_last_error[512] = "";
long long g_launch_count = 0;
bool g_prev_kernel = false;
}  // namespace frido

using namespace frido;

extern "C" int frido_conv2d(const FridoConvParams* p, void* stream) {
  if (!p) return set_error(FRIDO_E_ARG, "conv2d: null params");
  if (p->nrm_ab || p->a_presplit) return conv2d_nf(p, (cudaStream_t)stream);  // halo-resident operand path (engine 3; conv_nf.cu)
  if (p->engine >= 1 && p->engine <= 3) return conv2d_tc(p, (cudaStream_t)stream);
  return conv2d_simt(p, (cudaStream_t)stream);
}

extern "C" int frido_zero(void* ptr, int64_t nbytes, void* stream) {
  if (!ptr || nbytes < 0) return set_error(FRIDO_E_ARG, "zero: bad argument");
  ++g_launch_count;
  g_prev_kernel = false;  // a memset node: the next kernel takes a plain (full) dependency
  cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)nbytes, (cudaStream_t)stream);
  if (e != cudaSuccess) return set_error(FRIDO_E_LAUNCH, cudaGetErrorString(e));
  return FRIDO_OK;
}

extern "C" int frido_run_program(const FridoOp* ops, int32_t n, void* stream) {
  if (!ops || n < 0) return set_error(FRIDO_E_ARG, "run_program: bad argument");
  g_prev_kernel = false;  // whatever ran before this program is not ours: first kernel launches with a full dependency
  for (int i = 0; i < n; ++i) {
    const FridoOp& op = ops[i];
    int rc;
    switch (op.kind) {
      case FRIDO_OP_CONV: rc = frido_conv2d(&op.u.conv, stream); break;
      case FRIDO_OP_GN_STATS: rc = frido_gn_stats(&op.u.gn_stats, stream); break;
      case FRIDO_OP_NORM_ACT: rc = frido_norm_act(&op.u.norm_act, stream); break;
      case FRIDO_OP_LAYERNORM: rc = frido_layernorm(&op.u.layernorm, stream); break;
      case FRIDO_OP_SOFTMAX: rc = frido_softmax(&op.u.softmax, stream); break;
      case FRIDO_OP_TIME_EMBED: rc = frido_time_embed(&op.u.time_embed, stream); break;
      case FRIDO_OP_STEP_BEGIN: rc = frido_step_begin(&op.u.step_begin, stream); break;
      case FRIDO_OP_UPDATE: rc = frido_sampler_update(&op.u.update, stream); break;
      case FRIDO_OP_SNAP: rc = frido_stage_snap(&op.u.snap, stream); break;
      case FRIDO_OP_VQ: rc = frido_vq_lookup(&op.u.vq, stream); break;
      case FRIDO_OP_EMBED: rc = frido_embed_tokens(&op.u.embed, stream); break;
      case FRIDO_OP_ATTN: rc = frido_attn_small(&op.u.attn, stream); break;
      case FRIDO_OP_MHA: rc = frido_mha_small(&op.u.mha, stream); break;
      case FRIDO_OP_CONVT: rc = frido_conv_transpose2d(&op.u.convt, stream); break;
      case FRIDO_OP_ASSEMBLE: rc = frido_assemble_latent(&op.u.assemble, stream); break;
      case FRIDO_OP_UPSAMPLE: rc = frido_upsample2x(&op.u.upsample, stream); break;
      case FRIDO_OP_GN_FINALIZE: rc = frido_gn_finalize(&op.u.gn_finalize, stream); break;
      case FRIDO_OP_FLASH: rc = frido_attn_flash(&op.u.flash, stream); break;
      case FRIDO_OP_BLEND: rc = frido_mask_blend(&op.u.blend, stream); break;
      case FRIDO_OP_ZERO: rc = frido_zero(op.u.zero.ptr, op.u.zero.nbytes, stream); break;
      default: rc = set_error(FRIDO_E_ARG, "run_program: unknown op kind");
    }
    if (rc != FRIDO_OK) {
      char buf[600];
      snprintf(buf, sizeof(buf), "op %d (kind %d, tag %d): %s", i, op.kind, op.tag, g_last_error);
      set_error(rc, buf);
      return -(1000 + i);
    }
  }
  return FRIDO_OK;
}

extern "C" int frido_abi_version(void) { return FRIDO_ABI_VERSION; }
extern "C" int frido_sizeof_op(void) { return (int)sizeof(FridoOp); }
extern "C" const char* frido_last_error(void) { return g_last_error; }
extern "C" int64_t frido_launch_count(void) { return g_launch_count; }

extern "C" int frido_check_device(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
    return set_error(FRIDO_E_LAUNCH, "no CUDA device");
  if (prop.major != 10) {
    char buf[128];
    snprintf(buf, sizeof(buf), "device is sm_%d%d, this library is built for sm_100a only", prop.major, prop.minor);
    return set_error(FRIDO_E_ARCH, buf);
  }
  return FRIDO_OK;
}
