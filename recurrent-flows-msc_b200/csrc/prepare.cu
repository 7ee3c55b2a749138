// Batched refresh of everything DERIVED FROM PARAMETERS, one launch per kind for a whole model (sm_100a).
//
// After an optimizer step every convolution weight has to be repacked (bf16, K-major, tap-major), every ActNorm /
// Conv2dZeros needs its (scale, shift) = (exp(f*logs), bias*exp(f*logs)), and every GlowStep its folded ActNorm . InvConv
// matrix  Wf = P (L o mask + I)(U o mask^T + diag(sign_s e^{log_s})) diag(e^{logs}),  bf = Wf bias  and its log-determinant
// (Flow/glow_modules.py:33-54, 188-205 of the reference rebuilds these with ~10 ATen launches per module per call).
// Per module that was ~25 tiny launches (ours + torch), ~1900 per training step; here each kind is ONE launch driven by a
// table of pointers in device memory (the parameters live in one flat buffer and the outputs in persistent tensors, so
// the tables are built once).  The reverse-mode chain from (dWf, dbf, d logdet) back to the ActNorm / LU parameters is
// batched the same way.
//
// Table entries are arrays of 64-bit words (pointers and integers alike) so that the host side can build them as a
// plain int64 tensor.
#include <algorithm>

#include "common.cuh"

namespace rfk {

constexpr int kEntryWords = 24;

template <typename T>
__device__ __forceinline__ T* eptr(const long long* e, int i) { return reinterpret_cast<T*>(static_cast<uintptr_t>(e[i])); }

// ---- weights -------------------------------------------------------------------------------------------------------
// entry: 0 src f32, 1 dst bf16, 2 perm i32 (0 = none), 3 N, 4 Cin, 5 taps, 6 mode, 7 rows, 8 kp, 9 rows_pad, 10 ktot,
//        11 logs, 12 bias (modes 4 / 5 only: weights with an ActNorm folded in, see rfk_pack_weight_folded)
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const long long* __restrict__ table) {
  // no pdl_trigger(): conv kernels prefetch weights BEFORE their dependency wait, so dependents must not start early
  pdl_wait();
  const long long* e = table + (size_t)blockIdx.y * kEntryWords;
  const float* __restrict__ src = eptr<const float>(e, 0);
  __nv_bfloat16* __restrict__ dst = eptr<__nv_bfloat16>(e, 1);
  const int* __restrict__ perm = eptr<const int>(e, 2);
  const int N = (int)e[3], Cin = (int)e[4], taps = (int)e[5], mode = (int)e[6], rows = (int)e[7], kp = (int)e[8];
  const int rows_pad = (int)e[9], ktot = (int)e[10];
  const float* __restrict__ logs = eptr<const float>(e, 11);
  const float* __restrict__ bias = eptr<const float>(e, 12);
  const long long total = (long long)rows_pad * ktot;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ktot), k = (int)(i % ktot);
    float v = 0.0f;
    if (r < rows) {
      if (mode == 4) {
        const int kmain = taps * kp;
        const float sc = expf(logs[r]);
        if (k < kmain) {
          const int t = k / kp, j = k % kp;
          if (j < Cin) v = src[((long long)r * Cin + (perm ? perm[j] : j)) * taps + t] * sc;
        } else if (k < kmain + 2) {
          const float sh = bias[r] * sc;
          const float hi = __bfloat162float(__float2bfloat16(sh));
          v = k == kmain ? hi : sh - hi;
        }
      } else if (mode == 5) {
        const int t = k / kp, co = k % kp;
        if (co < N) v = src[((long long)co * Cin + (perm ? perm[r] : r)) * taps + (taps - 1 - t)] * expf(logs[r]);
      } else if (mode == 0) {
        const int t = k / kp, j = k % kp;
        if (j < Cin) v = src[((long long)r * Cin + (perm ? perm[j] : j)) * taps + t];
      } else if (mode == 1) {
        const int t = k / kp, co = k % kp;
        if (co < N) v = src[((long long)co * Cin + (perm ? perm[r] : r)) * taps + (taps - 1 - t)];
      } else if (mode == 2) {
        const int t = r / N, c = r % N;
        if (k < Cin) v = src[((long long)c * Cin + k) * taps + t];
      } else {
        const int R = rows / taps, t = r / R, j = r % R;
        const int ci = perm ? perm[j] : j;
        if (k < N && ci >= 0 && ci < Cin) v = src[((long long)k * Cin + ci) * taps + (taps - 1 - t)];
      }
    }
    dst[i] = __float2bfloat16(v);
  }
}

// ---- per-channel affines ---------------------------------------------------------------------------------------------
// entry: 0 logs, 1 bias (nullable), 2 out scale, 3 out shift, 4 n, 5 factor (integer multiplier of logs)
__global__ void __launch_bounds__(128) affine_prepare_batched_kernel(const long long* __restrict__ table) {
  // no pdl_trigger(): conv kernels prefetch weights BEFORE their dependency wait, so dependents must not start early
  pdl_wait();
  const long long* e = table + (size_t)blockIdx.x * kEntryWords;
  const float* logs = eptr<const float>(e, 0);
  const float* bias = eptr<const float>(e, 1);
  float* scale = eptr<float>(e, 2);
  float* shift = eptr<float>(e, 3);
  const int n = (int)e[4];
  const float f = (float)e[5];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float s = expf(logs[i] * f);
    scale[i] = s;
    shift[i] = bias ? bias[i] * s : 0.0f;
  }
}

// ---- GlowStep fold: ActNorm . InvConv (LU form) ----------------------------------------------------------------------------
// entry: 0 bias, 1 logs, 2 lower, 3 upper, 4 log_s, 5 sign_s, 6 perm i32 (row i of P has its one at column perm[i]), 7 C, 8 HW,
//        9 Wf [C,C], 10 WfT [C,C], 11 bf [C], 12 ld float[2] = {per-pixel log-det, HW * per-pixel log-det}
__device__ __forceinline__ float lu_lower(const float* lower, int C, int r, int k) {   // (L o mask + I)[r,k]
  return k < r ? lower[r * C + k] : (k == r ? 1.0f : 0.0f);
}
__device__ __forceinline__ float lu_upper(const float* upper, const float* log_s, const float* sign_s, int C, int k, int j) {
  return j > k ? upper[k * C + j] : (j == k ? sign_s[k] * expf(log_s[k]) : 0.0f);   // (U o mask^T + diag)[k,j]
}

__global__ void __launch_bounds__(256) fold_prepare_batched_kernel(const long long* __restrict__ table) {
  // no pdl_trigger(): conv kernels prefetch weights BEFORE their dependency wait, so dependents must not start early
  pdl_wait();
  const long long* e = table + (size_t)blockIdx.x * kEntryWords;
  const float* bias = eptr<const float>(e, 0);
  const float* logs = eptr<const float>(e, 1);
  const float* lower = eptr<const float>(e, 2);
  const float* upper = eptr<const float>(e, 3);
  const float* log_s = eptr<const float>(e, 4);
  const float* sign_s = eptr<const float>(e, 5);
  const int* perm = eptr<const int>(e, 6);
  const int C = (int)e[7];
  const float HW = (float)e[8];
  float* Wf = eptr<float>(e, 9);
  float* WfT = eptr<float>(e, 10);
  float* bf = eptr<float>(e, 11);
  float* ld = eptr<float>(e, 12);
  for (int idx = threadIdx.x; idx < C * C; idx += blockDim.x) {
    const int i = idx / C, j = idx - i * C, r = perm[i];
    const int kmax = min(r, j);
    float acc = 0.0f;
    for (int k = 0; k <= kmax; ++k) acc += lu_lower(lower, C, r, k) * lu_upper(upper, log_s, sign_s, C, k, j);
    const float w = acc * expf(logs[j]);
    Wf[idx] = w;
    WfT[j * C + i] = w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float acc = 0.0f;
    for (int j = 0; j < C; ++j) acc += Wf[i * C + j] * bias[j];
    bf[i] = acc;
  }
  if (threadIdx.x < 32) {
    float s = 0.0f;
    for (int i = threadIdx.x; i < C; i += 32) s += logs[i] + log_s[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) { ld[0] = s; ld[1] = s * HW; }
  }
}

// ---- reverse mode of the fold ------------------------------------------------------------------------------------------------
// entry: 0..8 as above, 9 Wf, 10 dWf [C,C], 11 dbf [C], 12 scratch A [C,C], 13 d bias, 14 d logs, 15 d lower, 16 d upper,
//        17 d log_s   (all outputs are ACCUMULATED into: they may be the parameters' .grad storage)
// G = sum over the batch of d loss / d logdet[b]: every GlowStep adds HW * (sum logs + sum log_s) to every sample's log-det.
__global__ void __launch_bounds__(256) fold_backward_batched_kernel(const long long* __restrict__ table, const float* __restrict__ Gp) {
  pdl_trigger();
  pdl_wait();
  const long long* e = table + (size_t)blockIdx.x * kEntryWords;
  const float* bias = eptr<const float>(e, 0);
  const float* logs = eptr<const float>(e, 1);
  const float* lower = eptr<const float>(e, 2);
  const float* upper = eptr<const float>(e, 3);
  const float* log_s = eptr<const float>(e, 4);
  const float* sign_s = eptr<const float>(e, 5);
  const int* perm = eptr<const int>(e, 6);
  const int C = (int)e[7];
  const float HW = (float)e[8];
  const float* Wf = eptr<const float>(e, 9);
  const float* dWf = eptr<const float>(e, 10);
  const float* dbf = eptr<const float>(e, 11);
  float* A = eptr<float>(e, 12);
  float* d_bias = eptr<float>(e, 13);
  float* d_logs = eptr<float>(e, 14);
  float* d_lower = eptr<float>(e, 15);
  float* d_upper = eptr<float>(e, 16);
  float* d_log_s = eptr<float>(e, 17);
  const float G = Gp ? *Gp : 0.0f;
  // d bias_j = sum_i dbf_i Wf_ij ;  d logs_j = sum_i (dWf_ij + dbf_i b_j) Wf_ij + HW G
  for (int j = threadIdx.x; j < C; j += blockDim.x) {
    float db = 0.0f, dl = 0.0f;
    const float bj = bias[j];
    for (int i = 0; i < C; ++i) {
      const float w = Wf[i * C + j];
      db += dbf[i] * w;
      dl += (dWf[i * C + j] + dbf[i] * bj) * w;
    }
    d_bias[j] += db;
    d_logs[j] += dl + HW * G;
  }
  // A = P^T dW with dW_ij = (dWf_ij + dbf_i b_j) e^{logs_j}:  A[perm[i], j] = dW[i, j]
  for (int idx = threadIdx.x; idx < C * C; idx += blockDim.x) {
    const int i = idx / C, j = idx - i * C;
    A[perm[i] * C + j] = (dWf[idx] + dbf[i] * bias[j]) * expf(logs[j]);
  }
  __syncthreads();
  // d L' = A U'^T (strictly lower part), d U' = L'^T A (upper part incl. the diagonal -> log_s)
  for (int idx = threadIdx.x; idx < C * C; idx += blockDim.x) {
    const int r = idx / C, c = idx - r * C;
    if (c < r) {          // d lower[r, c] = sum_j A[r, j] U'[c, j], j >= c
      float acc = 0.0f;
      for (int j = c; j < C; ++j) acc += A[r * C + j] * lu_upper(upper, log_s, sign_s, C, c, j);
      d_lower[idx] += acc;
    } else {              // d U'[r, c] = sum_q L'[q, r] A[q, c], q >= r
      float acc = 0.0f;
      for (int q = r; q < C; ++q) acc += lu_lower(lower, C, q, r) * A[q * C + c];
      if (c > r) d_upper[idx] += acc;
      else d_log_s[r] += acc * sign_s[r] * expf(log_s[r]) + HW * G;
    }
  }
}

// dst[b, dst_off + j, p] += src[b, src_off + j, p]   (fp32 NCHW, j < n): gradient accumulation into a channel window
__global__ void __launch_bounds__(256) add_channels_kernel(float* __restrict__ dst, int dst_C, int dst_off, const float* __restrict__ src,
                                                           int src_C, int src_off, int n, int HW, long long total) {
  pdl_trigger();
  pdl_wait();
  const long long per = (long long)n * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per, r = i - b * per;
    dst[(b * dst_C + dst_off) * HW + r] += src[(b * src_C + src_off) * HW + r];
  }
}
__global__ void __launch_bounds__(256) add_channels_v4_kernel(float4* __restrict__ dst, int dst_C, int dst_off, const float4* __restrict__ src,
                                                              int src_C, int src_off, int n, int HW4, long long total4) {
  pdl_trigger();
  pdl_wait();
  const long long per = (long long)n * HW4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per, r = i - b * per;
    float4* d = dst + (b * dst_C + dst_off) * HW4 + r;
    const float4 s = ld_stream(src + (b * src_C + src_off) * HW4 + r);
    float4 v = *d;
    v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    *d = v;
  }
}

// out[b, c, p] = ws[(b*HW + p)*ld + c] + bias[c]  (fp32; ws = the pixel-major accumulator of a split-K GEMM); optionally
// clears ws so that the next split-K launch finds zeros.  Small tensors (a few hundred pixels): one thread per element.
__global__ void __launch_bounds__(256) ws_to_nchw_kernel(float* __restrict__ ws, int ld, const float* __restrict__ bias,
                                                         float* __restrict__ out, int C, int HW, long long total, int zero) {
  pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long pix = i / C;
    const long long b = pix / HW;
    const int p = (int)(pix - b * HW);
    float* src = ws + pix * ld + c;
    out[(b * C + c) * HW + p] = *src + (bias ? bias[c] : 0.0f);
    if (zero) *src = 0.0f;
  }
}

// ActNorm parameter gradients of a Conv2dNorm layer WITHOUT a pass over the activations.  With v = (a + b) e^{logs} the
// pre-activation, a = conv(x, W) and da = dv e^{logs} the gradient w.r.t. the conv output:
//   d b[c]    = sum_p da[p,c]                                                   (= the column sums the fused data-gradient
//                                                                                  epilogue of the NEXT layer accumulated)
//   d logs[c] = sum_p dv v = sum_p da (a + b) = sum_k W[c,k] dW[c,k] + b[c] d b[c]   because dW[c,k] = sum_p da[p,c] x[p,k]
// One CTA per output channel c; K = Cin * taps.  dW is this backward's own weight gradient (a zeroed scratch the
// weight-gradient kernel just filled); it is also accumulated into the parameter's gradient buffer here.
__global__ void __launch_bounds__(256) actnorm_param_bwd_kernel(const float* __restrict__ W, const float* __restrict__ dW, long long K,
                                                                const float* __restrict__ colsum, const float* __restrict__ bias,
                                                                float* __restrict__ grad_W, float* __restrict__ d_logs,
                                                                float* __restrict__ d_bias) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x;
  const float* w = W + (long long)c * K;
  const float* g = dW + (long long)c * K;
  float* gw = grad_W + (long long)c * K;
  float acc = 0.0f;
  for (long long k = threadIdx.x; k < K; k += blockDim.x) {
    const float gv = g[k];
    acc += w[k] * gv;
    gw[k] += gv;
  }
  __shared__ float part[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int i = 0; i < 8; ++i) s += part[i];
    const float db = colsum[c];
    d_bias[c] += db;
    d_logs[c] += s + bias[c] * db;
  }
}

}  // namespace rfk

using namespace rfk;

extern "C" int rfk_actnorm_param_bwd(const float* W, const float* dW, long long K, const float* colsum, const float* bias,
                                     float* grad_W, float* d_logs, float* d_bias, int n, void* stream) {
  RFK_REQUIRE(W && dW && colsum && bias && grad_W && d_logs && d_bias && n > 0 && K > 0, "rfk_actnorm_param_bwd: null pointer or empty shape");
  RFK_LAUNCH(actnorm_param_bwd_kernel, n, 256, 0, (cudaStream_t)stream, W, dW, K, colsum, bias, grad_W, d_logs, d_bias);
  return check_launch("rfk_actnorm_param_bwd");
}

extern "C" int rfk_ws_to_nchw(float* ws, int ld, const float* bias, float* out, int B, int C, int HW, int zero, void* stream) {
  RFK_REQUIRE(ws && out && B > 0 && C > 0 && HW > 0 && ld >= C, "rfk_ws_to_nchw: null pointer or bad shape");
  const long long total = (long long)B * HW * C;
  RFK_LAUNCH(ws_to_nchw_kernel, stream_grid(total, 256, 8), 256, 0, (cudaStream_t)stream, ws, ld, bias, out, C, HW, total, zero);
  return check_launch("rfk_ws_to_nchw");
}

extern "C" int rfk_pack_weights_batched(const long long* table, int n_entries, long long max_elements, void* stream) {
  RFK_REQUIRE(table && n_entries > 0 && max_elements > 0, "rfk_pack_weights_batched: null table or no entries");
  const int gx = (int)std::min<long long>(64, (max_elements + 255) / 256);
  RFK_LAUNCH(pack_weights_batched_kernel, dim3(gx, n_entries), 256, 0, (cudaStream_t)stream, table);
  return check_launch("rfk_pack_weights_batched");
}

extern "C" int rfk_affine_prepare_batched(const long long* table, int n_entries, void* stream) {
  RFK_REQUIRE(table && n_entries > 0, "rfk_affine_prepare_batched: null table or no entries");
  RFK_LAUNCH(affine_prepare_batched_kernel, n_entries, 128, 0, (cudaStream_t)stream, table);
  return check_launch("rfk_affine_prepare_batched");
}

extern "C" int rfk_fold_prepare_batched(const long long* table, int n_entries, void* stream) {
  RFK_REQUIRE(table && n_entries > 0, "rfk_fold_prepare_batched: null table or no entries");
  RFK_LAUNCH(fold_prepare_batched_kernel, n_entries, 256, 0, (cudaStream_t)stream, table);
  return check_launch("rfk_fold_prepare_batched");
}

extern "C" int rfk_fold_backward_batched(const long long* table, int n_entries, const float* g_sum, void* stream) {
  RFK_REQUIRE(table && n_entries > 0, "rfk_fold_backward_batched: null table or no entries");
  RFK_LAUNCH(fold_backward_batched_kernel, n_entries, 256, 0, (cudaStream_t)stream, table, g_sum);
  return check_launch("rfk_fold_backward_batched");
}

extern "C" int rfk_add_channels(float* dst, int dst_C, int dst_off, const float* src, int src_C, int src_off, int n, int B, int HW,
                                void* stream) {
  RFK_REQUIRE(dst && src && n > 0 && B > 0 && HW > 0 && dst_off >= 0 && dst_off + n <= dst_C && src_off >= 0 && src_off + n <= src_C,
              "rfk_add_channels: null pointer or channel window out of range");
  const long long total = (long long)B * n * HW;
  if (HW % 4 == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    RFK_LAUNCH(add_channels_v4_kernel, stream_grid(total / 4, 256, 8), 256, 0, (cudaStream_t)stream, (float4*)dst, dst_C, dst_off,
               (const float4*)src, src_C, src_off, n, HW / 4, total / 4);
  } else {
    RFK_LAUNCH(add_channels_kernel, stream_grid(total, 256, 8), 256, 0, (cudaStream_t)stream, dst, dst_C, dst_off, src, src_C,
               src_off, n, HW, total);
  }
  return check_launch("rfk_add_channels");
}
