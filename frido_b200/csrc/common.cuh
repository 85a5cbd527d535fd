// Shared helpers for the frido_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <utility>

#include "../../include/frido_b200.h"

#ifndef FRIDO_PDL_DEFAULT
#define FRIDO_PDL_DEFAULT false
#endif

namespace frido {

extern char g_last_error[512];
extern long long g_launch_count;

inline int set_error(int code, const char* msg) {
  snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
  return code;
}

// Programmatic dependent launch (PDL): a kernel launched with the attribute may start while its predecessor on the stream
// is still draining; it must execute pdl_wait() before it touches global memory the predecessor may write (or still
// read).  The hot kernels call pdl_trigger() at their top, so the successor's launch latency and set-up (barrier init,
// TMEM allocation, tensor-map prefetch) overlap the predecessor's tail.  Only kernel -> kernel edges inside one op
// program use it (g_prev_kernel); FRIDO_PDL=0 turns it off.
extern bool g_prev_kernel;
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("FRIDO_PDL"); return e ? atoi(e) != 0 : FRIDO_PDL_DEFAULT; }();
  return on;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef FRIDO_PDL_TRIGGER
#define FRIDO_PDL_TRIGGER 1  // 0: no explicit trigger - a dependent launch then starts when this grid completes
#endif
__device__ __forceinline__ void pdl_trigger() {
#if FRIDO_PDL_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && g_prev_kernel) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);  // errors surface in check_launch
}

// cudaFuncSetAttribute (the > 48 KB dynamic shared memory opt-in) is PER DEVICE: one flag per (call site, device), so a
// process that drives several GPUs opts in on each of them.
struct DevOnce {
  bool done[64] = {};
  bool need() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

inline int check_launch(const char* what) {
  ++g_launch_count;
  g_prev_kernel = true;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
    return FRIDO_E_LAUNCH;
  }
  return FRIDO_OK;
}

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// bf16 hi/lo split of one value: hi = bf16_rn(v), lo = bf16_rn(v - hi)
__device__ __forceinline__ void split_bf16(float v, uint16_t& hi, uint16_t& lo) {
  uint32_t h, l;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(0.f), "f"(v));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(0.f), "f"(v - __uint_as_float(h << 16)));
  hi = (uint16_t)h; lo = (uint16_t)l;
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }
// exact-erf GELU (attention.py:44 F.gelu default)
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }
// Same function for the tensor-core GEGLU epilogue, where 2 erf per output make the epilogue the bottleneck:
// erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. fp32 rounding level) = 1 rcp + 1 ex2 + 7 FMA, branch-free.
__device__ __forceinline__ float gelu_erf_fast(float v) {
  const float x = fabsf(v) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * __expf(-x * x);   // erf(|v|/sqrt2)
  return 0.5f * v * (1.0f + copysignf(e, v));
}

// Output formatting of the sampling script, same fp32 operation order as the reference so the bytes are identical:
//   mode 0 = custom_to_np  (scripts/sample_diffusion.py:115-121): ((x + 1) * 127.5).clamp(0, 255) -> uint8 (truncate)
//   mode 1 = custom_to_pil (scripts/sample_diffusion.py:103-108): 255 * ((clamp(x,-1,1) + 1) / 2)  -> uint8 (truncate)
__device__ __forceinline__ uint8_t format_u8(float x, int mode) {
  float t;
  if (mode == 0) {
    t = __fmul_rn(__fadd_rn(x, 1.0f), 127.5f);
    t = fminf(fmaxf(t, 0.f), 255.f);
  } else {
    const float cl = fminf(fmaxf(x, -1.f), 1.f);
    t = __fmul_rn(255.0f, __fdiv_rn(__fadd_rn(cl, 1.0f), 2.0f));
  }
  return (uint8_t)(int)t;  // truncation, as torch .to(uint8) / numpy astype(uint8) on non-negative values
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

int conv2d_simt(const FridoConvParams* p, cudaStream_t s);
int conv2d_tc(const FridoConvParams* p, cudaStream_t s);
int conv2d_nf(const FridoConvParams* p, cudaStream_t s);
bool conv2d_nf_eligible(const FridoConvParams* p);

}  // namespace frido
