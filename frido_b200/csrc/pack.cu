// Weight packing behind the C ABI (SURVEY.md §8b "pack_weights / workspace_bytes"): everything the host runtime does to a
// checkpoint tensor before the engines can read it is one of these launches, so a non-Python host can drive the library.
//   frido_pack_permute3      strided 3-D gather -> strided 3-D scatter (OIHW -> [O][tap][I], concatenation along rows or
//                            columns, transposes, GEGLU row interleave: all are strides)
//   frido_pack_conv_weight   the conv case of it, by name
//   frido_matmul_f64acc      C = A B with fp64 products and sums, rounded once (attention weight folds, unet.py)
//   frido_fold_self_attention  A = Wk^T Wq and Wv' = Wo Wv of one CrossAttention module (attention.py:172-191 re-associated)
//   frido_vec_add            bias of a fused conv pair
//   frido_workspace_bytes    stream-K workspace an op program needs
#include "common.cuh"

namespace frido {

__global__ void __launch_bounds__(256) permute3_kernel(const float* __restrict__ src, long long s0, long long s1, long long s2,
                                                       float* __restrict__ dst, long long d0, long long d1, long long d2, int n0,
                                                       int n1, int n2) {
  const long long total = (long long)n0 * n1 * n2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % n2);
    const long long r = i / n2;
    const int b = (int)(r % n1);
    const int a = (int)(r / n1);
    dst[a * d0 + b * d1 + c * d2] = src[a * s0 + b * s1 + c * s2];
  }
}

// out[m][n] = sum_k a[m*a_rs + k*a_cs] * b[k*b_rs + n*b_cs], fp64 accumulate in k order (deterministic), 16x16 tiles
__global__ void __launch_bounds__(256) matmul_f64_kernel(const float* __restrict__ a, long long a_rs, long long a_cs,
                                                         const float* __restrict__ b, long long b_rs, long long b_cs, int M, int N,
                                                         int K, float* __restrict__ out, long long o_ld) {
  __shared__ double sa[16][17], sb[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  double acc = 0.0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int ka = k0 + tx, kb = k0 + ty;
    sa[ty][tx] = (m < M && ka < K) ? (double)a[m * a_rs + ka * a_cs] : 0.0;
    sb[ty][tx] = (kb < K && n < N) ? (double)b[kb * b_rs + n * b_cs] : 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fma(sa[ty][k], sb[k][tx], acc);
    __syncthreads();
  }
  if (m < M && n < N) out[m * o_ld + n] = (float)acc;
}

__global__ void __launch_bounds__(256) vec_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                      long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = a[i] + b[i];
}

static int plain_launch_done(const char* what) {
  const int rc = check_launch(what);
  g_prev_kernel = false;  // these kernels carry no griddepcontrol: the next launch takes a full dependency
  return rc;
}

}  // namespace frido

using namespace frido;

extern "C" int frido_pack_permute3(const float* src, int64_t s0, int64_t s1, int64_t s2, float* dst, int64_t d0, int64_t d1, int64_t d2,
                                   int32_t n0, int32_t n1, int32_t n2, void* stream) {
  if (!src || !dst || n0 < 0 || n1 < 0 || n2 < 0) return set_error(FRIDO_E_ARG, "pack_permute3: bad argument");
  const long long total = (long long)n0 * n1 * n2;
  if (total == 0) return FRIDO_OK;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  permute3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, s0, s1, s2, dst, d0, d1, d2, n0, n1, n2);
  return plain_launch_done("pack_permute3");
}

extern "C" int frido_pack_conv_weight(const float* w, int32_t O, int32_t I, int32_t KH, int32_t KW, float* dst, int64_t dst_ld,
                                      void* stream) {
  if (O <= 0 || I <= 0 || KH <= 0 || KW <= 0 || dst_ld < (int64_t)I * KH * KW) return set_error(FRIDO_E_ARG, "pack_conv_weight: bad shape");
  const int T = KH * KW;
  // src [O][I][T] -> dst [O][T][I]
  return frido_pack_permute3(w, (int64_t)I * T, 1, T, dst, dst_ld, I, 1, O, T, I, stream);
}

extern "C" int frido_matmul_f64acc(const float* a, int64_t a_rs, int64_t a_cs, const float* b, int64_t b_rs, int64_t b_cs, int32_t M,
                                   int32_t N, int32_t K, float* out, int64_t o_ld, void* stream) {
  if (!a || !b || !out || M <= 0 || N <= 0 || K <= 0 || o_ld < N) return set_error(FRIDO_E_ARG, "matmul_f64acc: bad argument");
  matmul_f64_kernel<<<dim3((N + 15) / 16, (M + 15) / 16), 256, 0, (cudaStream_t)stream>>>(a, a_rs, a_cs, b, b_rs, b_cs, M, N, K, out, o_ld);
  return plain_launch_done("matmul_f64acc");
}

extern "C" int frido_fold_self_attention(const float* wq, const float* wk, const float* wv, const float* wo, int32_t C, float* a_out,
                                         float* wv_out, void* stream) {
  if (!wq || !wk || !wv || !wo || !a_out || !wv_out || C <= 0) return set_error(FRIDO_E_ARG, "fold_self_attention: bad argument");
  // A = Wk^T Wq : A[m][n] = sum_k Wk[k][m] Wq[k][n]
  int rc = frido_matmul_f64acc(wk, 1, C, wq, C, 1, C, C, C, a_out, C, stream);
  if (rc != FRIDO_OK) return rc;
  // Wv' = Wo Wv
  return frido_matmul_f64acc(wo, C, 1, wv, C, 1, C, C, C, wv_out, C, stream);
}

extern "C" int frido_vec_add(const float* a, const float* b, float* out, int64_t n, void* stream) {
  if (!a || !b || !out || n < 0) return set_error(FRIDO_E_ARG, "vec_add: bad argument");
  if (n == 0) return FRIDO_OK;
  const int grid = (int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  vec_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
  return plain_launch_done("vec_add");
}

extern "C" int64_t frido_workspace_bytes(const FridoOp* ops, int32_t n) {
  // The only scratch the engines need beyond their operands is the stream-K workspace of the tcgen05 convs (partial
  // accumulators + arrival counters, FridoConvParams.sk_ws): one per stream is enough, launches are stream-ordered.
  if (!ops) return FRIDO_SK_WS_BYTES;
  for (int i = 0; i < n; ++i)
    if (ops[i].kind == FRIDO_OP_CONV && ops[i].u.conv.engine >= 1 && ops[i].u.conv.engine <= 3) return FRIDO_SK_WS_BYTES;
  return 0;
}
