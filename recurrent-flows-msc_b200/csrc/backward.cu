// Backward building blocks of the coupling network (sm_100a).  SURVEY.md 7.2 lists the formulas.
//
//   * rfk_act_affine_bwd   : backward of  h = act(conv*scale + shift)  w.r.t. the conv output, plus the two per-channel
//                            reductions that give d(logs) and d(bias) of the ActNorm (needs only h and dh)
//   * rfk_conv_wgrad       : weight gradient  dW[n, tap, c] = sum_p dY[p, n] * X[p + off(tap), c]  on the tensor cores
//   * data gradient        : no kernel of its own -- it is rfk_conv_gemm on the tap-flipped, transposed weights
//
// Round-1 status: building blocks with parity tests; the autograd wiring of the modules comes next.
#include <mma.h>

#include <algorithm>

#include <cstdlib>

#include "common.cuh"

namespace rfk {

int conv_wgrad_tc(const void* x, int x_ld, int cin, const void* dy, int dy_ld, int cout, int B, int H, int W, int taps,
                  float* dw, int dw_ld, int layout, const int* perm, void* ws, long long ws_bytes,
                  cudaStream_t stream);   // wgrad_tc.cu

// ------------------------------------------------------------------------------------------
// h = act(v), v = a*scale + shift (a = raw conv output, scale = e^{logs}, shift = bias*e^{logs}).
//   dv = dh * act'(v);  da = dv * scale  (written as bf16 NHWC);
//   d(logs) = sum_p dv*v = sum_p dh*h   (ReLU and LeakyReLU: dv*v == dh*h),   d(bias) = scale * sum_p dv.
// Rows = pixels, 8 channels per thread (one 16-byte load of dh and of h); a CTA walks a slice of rows and adds its
// per-channel partial sums with one atomicAdd per channel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) act_affine_bwd_kernel(const __nv_bfloat16* __restrict__ dh,
                                                             const __nv_bfloat16* __restrict__ h, int ld, int n,
                                                             const float* __restrict__ scale, int act_fn,
                                                             __nv_bfloat16* __restrict__ da, int da_ld,
                                                             float* __restrict__ r_dv, float* __restrict__ r_dvv,
                                                             float dvv_factor, int dv_scaled, long long rows,
                                                             long long rows_per_cta) {
  pdl_trigger();
  pdl_wait();
  const int groups = (n + 7) >> 3;              // channel groups of 8
  const int lanes_per_row = groups;             // threads that cover one row
  const int rows_par = blockDim.x / lanes_per_row;
  const int g = threadIdx.x % lanes_per_row, rl = threadIdx.x / lanes_per_row;
  float s_dv[8], s_dvv[8], sc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s_dv[k] = 0.0f; s_dvv[k] = 0.0f;
    sc[k] = (8 * g + k < n) ? scale[8 * g + k] : 0.0f;
  }
  const long long r0 = blockIdx.x * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  if (rl < rows_par) {
    for (long long r = r0 + rl; r < r1; r += rows_par) {
      const uint4 dq = *reinterpret_cast<const uint4*>(dh + r * ld + 8 * g);
      const uint4 hq = *reinterpret_cast<const uint4*>(h + r * ld + 8 * g);
      const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&dq);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hq);
      __nv_bfloat162 o2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 dv2 = __bfloat1622float2(d2[k]), hv2 = __bfloat1622float2(h2[k]);
        float dv[2] = {dv2.x, dv2.y};
        const float hv[2] = {hv2.x, hv2.y};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float slope = act_fn == RFK_ACT_NONE ? 1.0f : (hv[e] > 0.0f ? 1.0f : (act_fn == RFK_ACT_LEAKY ? 0.2f : 0.0f));
          const float v = act_fn == RFK_ACT_LEAKY && hv[e] < 0.0f ? hv[e] * 5.0f : hv[e];   // pre-activation value
          dv[e] *= slope;
          s_dv[2 * k + e] += dv[e];
          s_dvv[2 * k + e] += dv[e] * v;
          dv[e] *= sc[2 * k + e];
        }
        o2[k] = __floats2bfloat162_rn(dv[0], dv[1]);
      }
      *reinterpret_cast<uint4*>(da + r * da_ld + 8 * g) = *reinterpret_cast<uint4*>(o2);
    }
  }
  // reduce over the rows_par row-lanes of the CTA, then one atomic per channel
  __shared__ float red[2][2048];  // [2][256 threads * 8 channels]
  float* a0 = &red[0][0];
  float* a1 = &red[1][0];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a0[threadIdx.x * 8 + k] = s_dv[k];
    a1[threadIdx.x * 8 + k] = s_dvv[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < groups * 8; c += blockDim.x) {
    const int gg = c >> 3, k = c & 7;
    float t0 = 0.0f, t1 = 0.0f;
    for (int rr = 0; rr < rows_par; ++rr) {
      t0 += a0[(rr * lanes_per_row + gg) * 8 + k];
      t1 += a1[(rr * lanes_per_row + gg) * 8 + k];
    }
    if (c < n) {   // both transforms are linear, so every CTA may apply them to its partial sums
      atomicAdd(r_dv + c, dv_scaled ? t0 * scale[c] : t0);
      atomicAdd(r_dvv + c, t1 * dvv_factor);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores (warp-level mma.sync through the WMMA API; a tcgen05 version with MN-major
// operand descriptors is the planned replacement -- DESIGN.md 7).
//
//   dW[tap][n][c] = sum_p dY[p, n] * X[shift_tap(p), c]       X, dY: NHWC bf16; out-of-image taps contribute zero
//
// CTA tile = 64 output channels (n) x 64 input channels (c) for ONE tap, accumulated over a slice of pixels
// (grid.z = pixel slices, fp32 atomicAdd of the 64x64 tile at the end).  Per 64-pixel chunk the dY tile [64 px x 64 n]
// and the tap-shifted X tile [64 px x 64 c] (zero rows where the tap leaves the image) are staged in shared memory;
// dY^T is consumed as a col-major matrix_a fragment, X as a row-major matrix_b fragment, so no transpose is materialised.
// 8 warps: warp w owns the 16x32 sub-tile (n-block w/2, c-blocks 2*(w%2), 2*(w%2)+1).
// ------------------------------------------------------------------------------------------
constexpr int WG_N = 64, WG_C = 64, WG_P = 64, WG_LD = 72;   // padded leading dimension (bank conflicts)

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int cin,
                                                         const __nv_bfloat16* __restrict__ dy, int dy_ld, int cout,
                                                         int B, int H, int W, int taps, float* __restrict__ dw,
                                                         int dw_ld, int layout, const int* __restrict__ perm,
                                                         long long pix_per_slice) {
  using namespace nvcuda;
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(32) __nv_bfloat16 sy[WG_P * WG_LD];
  __shared__ __align__(32) __nv_bfloat16 sx[WG_P * WG_LD];
  const int c_tiles = (cin + WG_C - 1) / WG_C;
  const int tap = blockIdx.x / c_tiles, c0 = (blockIdx.x % c_tiles) * WG_C;
  const int n0 = blockIdx.y * WG_N;
  const int dyo = taps == 9 ? tap / 3 - 1 : 0, dxo = taps == 9 ? tap % 3 - 1 : 0;
  const long long npix = (long long)B * H * W;
  const long long p_begin = blockIdx.z * pix_per_slice, p_end = min(npix, p_begin + pix_per_slice);
  const int warp = threadIdx.x >> 5;
  const int nb = warp >> 1, cb = (warp & 1) * 2;
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2];
  wmma::fill_fragment(acc[0], 0.0f);
  wmma::fill_fragment(acc[1], 0.0f);
  for (long long pc = p_begin; pc < p_end; pc += WG_P) {
    // stage: 64 rows x 64 channels = 512 x 16-byte vectors per tile, 2 per thread
#pragma unroll
    for (int v = threadIdx.x; v < WG_P * 8; v += 256) {
      const int r = v >> 3, q = (v & 7) * 8;
      const long long p = pc + r;
      uint4 yv = make_uint4(0, 0, 0, 0), xv = make_uint4(0, 0, 0, 0);
      if (p < p_end) {
        if (n0 + q < cout) yv = *reinterpret_cast<const uint4*>(dy + p * dy_ld + n0 + q);
        const int xx = (int)(p % W), yy = (int)((p / W) % H);
        const int sxx = xx + dxo, syy = yy + dyo;
        if (sxx >= 0 && sxx < W && syy >= 0 && syy < H && c0 + q < x_ld)
          xv = *reinterpret_cast<const uint4*>(x + (p + (long long)dyo * W + dxo) * x_ld + c0 + q);
      }
      *reinterpret_cast<uint4*>(sy + r * WG_LD + q) = yv;
      *reinterpret_cast<uint4*>(sx + r * WG_LD + q) = xv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WG_P; k += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::col_major> fa;   // dY^T[n, p]
      wmma::load_matrix_sync(fa, sy + k * WG_LD + nb * 16, WG_LD);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fb;  // X[p, c]
        wmma::load_matrix_sync(fb, sx + k * WG_LD + (cb + j) * 16, WG_LD);
        wmma::mma_sync(acc[j], fa, fb, acc[j]);
      }
    }
    __syncthreads();
  }
  // accumulate the 64x64 tile: dw[(tap*cout_rows + n) * dw_ld + c]  (layout [taps][cout][dw_ld])
  __shared__ float so[8][16 * 16];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    wmma::store_matrix_sync(&so[warp][0], acc[j], 16, wmma::mem_row_major);
    __syncwarp();
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < 256; e += 32) {
      const int rn = e >> 4, rc = e & 15;
      const int n = n0 + nb * 16 + rn, c = c0 + (cb + j) * 16 + rc;
      if (n < cout && c < cin)
        atomicAdd(dw + (layout ? ((long long)n * dw_ld + (perm ? perm[c] : c)) * taps + tap
                               : ((long long)tap * cout + n) * dw_ld + c), so[warp][e]);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// Affine coupling tail, backward (tap-split form; forward = coupling_taps_kernel in elementwise.cu).
//   forward:  t = S*sc[2j] + sh[2j],  raw = R*sc[2j+1] + sh[2j+1]  (S, R = nine-tap sums),  ls = clamp(raw),
//             z2' = (z2 + t) * e^{ls},  logdet[b] += sum ls
//   given dz (gradient w.r.t. the coupling output, [B,C,H,W], z2 half overwritten with the gradient w.r.t. z2),
//   z_out (the coupling output) and g_ld[b] (gradient w.r.t. logdet[b]):
//     dls = dz2'*z2' + g_ld,  dt = dz2'*e^{ls},  dz2 = dz2'*e^{ls},  draw = dls*clamp'(raw)
//     dsum[2j] = dt*sc[2j],  dsum[2j+1] = draw*sc[2j+1]                    (gradient w.r.t. S and R, fp32 NCHW)
//     d sc[2j] += dt*S, d sh[2j] += dt, d sc[2j+1] += draw*R, d sh[2j+1] += draw   (Conv2dZeros affine)
//     realnvp clamp ls = a*tanh(raw)+b:  d a += dls*tanh(raw),  d b += dls
//   grid = (chunks, B); the eight per-channel sums are reduced per CTA (one (j) per blockIdx.z) then atomically added.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = 0.0f;
  if (threadIdx.x < 32) {
    r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0f;
    r = warp_sum(r);
  }
  return r;  // valid in thread 0
}

__global__ void __launch_bounds__(256) coupling_taps_bwd_kernel(const float* __restrict__ taps, const float* __restrict__ z_out,
                                                                float* __restrict__ dz, float* __restrict__ dsum, int B, int C,
                                                                int H, int W, const float* __restrict__ scale,
                                                                const float* __restrict__ shift, int clamp_type,
                                                                const float* __restrict__ cs, const float* __restrict__ csh,
                                                                const float* __restrict__ g_ld, float* __restrict__ d_scale,
                                                                float* __restrict__ d_shift, float* __restrict__ d_cs,
                                                                float* __restrict__ d_csh, float logs_factor) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int j = blockIdx.y, half = C >> 1, HW = H * W;
  const float sc_s = scale[2 * j], sc_r = scale[2 * j + 1], sh_r = shift[2 * j + 1];
  float a = 0.0f, bb = 0.0f;
  if (clamp_type == RFK_CLAMP_REALNVP) { a = cs[j]; bb = csh[j]; }
  float acc[6] = {0, 0, 0, 0, 0, 0};   // d sc_s, d sh_s, d sc_r, d sh_r, d a, d b
  const long long n = (long long)B * HW;
  // threads run over (sample, pixel) jointly: deep levels have only a few pixels per sample
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / HW), p = (int)(idx % HW);
    const float* tb = taps + (long long)b * 9 * C * HW;
    const float* zo = z_out + ((long long)b * C + half + j) * HW;
    float* dzp = dz + ((long long)b * C + half + j) * HW;
    float* dsp = dsum + ((long long)b * C + 2 * j) * HW;
    const float gl = g_ld ? g_ld[b] : 0.0f;
    const int y = p / W, x = p - y * W;
    float S = 0.0f, R = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const float* q = tb + ((long long)((3 * ky + kx) * C + 2 * j) * HW) + yy * W + xx;
        S += __ldg(q);
        R += __ldg(q + HW);
      }
    }
    const float raw = fmaf(R, sc_r, sh_r);
    float ls, dclamp, th = 0.0f;
    switch (clamp_type) {
      case RFK_CLAMP_REALNVP: th = tanhf(raw); ls = a * th + bb; dclamp = a * (1.0f - th * th); break;
      case RFK_CLAMP_GLOW: { const float t = -(raw + 2.0f); ls = -(t > 15.0f ? t : log1pf(expf(t))); dclamp = 1.0f / (1.0f + expf(raw + 2.0f)); break; }
      case RFK_CLAMP_SOFT: { const float u = raw * (1.0f / 2.5f); ls = 2.5f * 0.636f * atanf(u); dclamp = 0.636f / (1.0f + u * u); break; }
      default: ls = raw; dclamp = 1.0f;
    }
    const float e = expf(ls);
    const float dzo = dzp[p];
    const float dls = dzo * zo[p] + gl;
    const float dt = dzo * e;
    const float draw = dls * dclamp;
    dzp[p] = dt;                 // gradient w.r.t. z2 (the coupling's input half)
    dsp[p] = dt * sc_s;          // gradient w.r.t. S
    dsp[HW + p] = draw * sc_r;   // gradient w.r.t. R
    acc[0] += dt * S; acc[1] += dt; acc[2] += draw * R; acc[3] += draw; acc[4] += dls * th; acc[5] += dls;
  }
  float red[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) red[k] = block_sum_256(acc[k], sh);
  if (threadIdx.x == 0) {
    if (logs_factor != 0.0f) {
      // parameter form (linear in the partial sums): out = (conv + bias) * exp(f*logs) -> d logs = f*(d_sc*scale + d_sh*shift),
      // d bias = d_sh*scale; written to d_scale (d logs) and d_shift (d bias)
      const float sh_s = shift[2 * j];
      atomicAdd(d_scale + 2 * j, logs_factor * (red[0] * sc_s + red[1] * sh_s));
      atomicAdd(d_shift + 2 * j, red[1] * sc_s);
      atomicAdd(d_scale + 2 * j + 1, logs_factor * (red[2] * sc_r + red[3] * sh_r));
      atomicAdd(d_shift + 2 * j + 1, red[3] * sc_r);
    } else {
      atomicAdd(d_scale + 2 * j, red[0]);
      atomicAdd(d_shift + 2 * j, red[1]);
      atomicAdd(d_scale + 2 * j + 1, red[2]);
      atomicAdd(d_shift + 2 * j + 1, red[3]);
    }
    if (clamp_type == RFK_CLAMP_REALNVP) {
      atomicAdd(d_cs + j, red[4]);
      atomicAdd(d_csh + j, red[5]);
    }
  }
}

// Same, four consecutive pixels of a row per thread (W % 4 == 0, 16-byte aligned tensors): per tap plane one 128-bit load
// plus, for the horizontally shifted taps, one scalar edge element (the layout of coupling_taps_v4_kernel).
__global__ void __launch_bounds__(256, 4) coupling_taps_bwd_v4_kernel(const float* __restrict__ taps, const float* __restrict__ z_out,
                                                                   float* __restrict__ dz, float* __restrict__ dsum, int B, int C,
                                                                   int H, int W, const float* __restrict__ scale,
                                                                   const float* __restrict__ shift, int clamp_type,
                                                                   const float* __restrict__ cs, const float* __restrict__ csh,
                                                                   const float* __restrict__ g_ld, float* __restrict__ d_scale,
                                                                   float* __restrict__ d_shift, float* __restrict__ d_cs,
                                                                   float* __restrict__ d_csh, float logs_factor) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int j = blockIdx.y, half = C >> 1, HW = H * W, W4 = W >> 2, HW4 = H * W4;
  const float sc_s = scale[2 * j], sc_r = scale[2 * j + 1], sh_r = shift[2 * j + 1];
  float a = 0.0f, bb = 0.0f;
  if (clamp_type == RFK_CLAMP_REALNVP) { a = cs[j]; bb = csh[j]; }
  float acc[6] = {0, 0, 0, 0, 0, 0};
  const long long n = (long long)B * HW4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / HW4), q = (int)(idx - (long long)b * HW4);
    const int y = q / W4, x = (q - y * W4) << 2;
    const float* tb = taps + (long long)b * 9 * C * HW;
    float S[4] = {0, 0, 0, 0}, R[4] = {0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* qq0 = tb + ((long long)((3 * ky + kx) * C + 2 * j) * HW) + yy * W + x;
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const float* qq = qq0 + pr * HW;
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(qq));
          float v0, v1, v2, v3;
          if (kx == 1) { v0 = c4.x; v1 = c4.y; v2 = c4.z; v3 = c4.w; }
          else if (kx == 0) { v0 = x > 0 ? __ldg(qq - 1) : 0.0f; v1 = c4.x; v2 = c4.y; v3 = c4.z; }
          else { v0 = c4.y; v1 = c4.z; v2 = c4.w; v3 = x + 4 < W ? __ldg(qq + 4) : 0.0f; }
          float* dst = pr ? R : S;
          dst[0] += v0; dst[1] += v1; dst[2] += v2; dst[3] += v3;
        }
      }
    }
    const long long off = ((long long)b * C + half + j) * HW + y * W + x;
    const float4 zo4 = ld_stream(reinterpret_cast<const float4*>(z_out + off));
    const float4 dz4 = *reinterpret_cast<const float4*>(dz + off);
    const float zo[4] = {zo4.x, zo4.y, zo4.z, zo4.w}, dzo[4] = {dz4.x, dz4.y, dz4.z, dz4.w};
    const float gl = g_ld ? g_ld[b] : 0.0f;
    float o_dz[4], o_s[4], o_r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float raw = fmaf(R[k], sc_r, sh_r);
      float ls, dclamp, th = 0.0f;
      switch (clamp_type) {
        case RFK_CLAMP_REALNVP: th = tanhf(raw); ls = a * th + bb; dclamp = a * (1.0f - th * th); break;
        case RFK_CLAMP_GLOW: { const float t = -(raw + 2.0f); ls = -(t > 15.0f ? t : log1pf(expf(t))); dclamp = 1.0f / (1.0f + expf(raw + 2.0f)); break; }
        case RFK_CLAMP_SOFT: { const float u = raw * (1.0f / 2.5f); ls = 2.5f * 0.636f * atanf(u); dclamp = 0.636f / (1.0f + u * u); break; }
        default: ls = raw; dclamp = 1.0f;
      }
      const float e = expf(ls);
      const float dls = dzo[k] * zo[k] + gl;
      const float dt = dzo[k] * e;
      const float draw = dls * dclamp;
      o_dz[k] = dt; o_s[k] = dt * sc_s; o_r[k] = draw * sc_r;
      acc[0] += dt * S[k]; acc[1] += dt; acc[2] += draw * R[k]; acc[3] += draw; acc[4] += dls * th; acc[5] += dls;
    }
    *reinterpret_cast<float4*>(dz + off) = make_float4(o_dz[0], o_dz[1], o_dz[2], o_dz[3]);
    const long long so = ((long long)b * C + 2 * j) * HW + y * W + x;
    st_stream(reinterpret_cast<float4*>(dsum + so), make_float4(o_s[0], o_s[1], o_s[2], o_s[3]));
    st_stream(reinterpret_cast<float4*>(dsum + so + HW), make_float4(o_r[0], o_r[1], o_r[2], o_r[3]));
  }
  float red[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) red[k] = block_sum_256(acc[k], sh);
  if (threadIdx.x == 0) {
    if (logs_factor != 0.0f) {
      const float sh_s = shift[2 * j];
      atomicAdd(d_scale + 2 * j, logs_factor * (red[0] * sc_s + red[1] * sh_s));
      atomicAdd(d_shift + 2 * j, red[1] * sc_s);
      atomicAdd(d_scale + 2 * j + 1, logs_factor * (red[2] * sc_r + red[3] * sh_r));
      atomicAdd(d_shift + 2 * j + 1, red[3] * sc_r);
    } else {
      atomicAdd(d_scale + 2 * j, red[0]);
      atomicAdd(d_shift + 2 * j, red[1]);
      atomicAdd(d_scale + 2 * j + 1, red[2]);
      atomicAdd(d_shift + 2 * j + 1, red[3]);
    }
    if (clamp_type == RFK_CLAMP_REALNVP) {
      atomicAdd(d_cs + j, red[4]);
      atomicAdd(d_csh + j, red[5]);
    }
  }
}

// Gradient w.r.t. the nine tap planes, written directly as the NHWC bf16 operand of the tap GEMM's backward:
//   dtaps[p, t*C + c] = dsum[c](p - off(t))  (zero when that pixel is outside the image)
__global__ void __launch_bounds__(256) taps_scatter_kernel(const float* __restrict__ dsum, __nv_bfloat16* __restrict__ dtaps,
                                                           int ld, int C, int H, int W, long long total) {
  pdl_trigger();
  pdl_wait();
  const int n9 = 9 * C, HW = H * W;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(t % n9);
    const long long pix = t / n9;
    const int tap = col / C, c = col - tap * C;
    const int p = (int)(pix % HW);
    const long long b = pix / HW;
    const int y = p / W - (tap / 3 - 1), x = p % W - (tap % 3 - 1);
    float v = 0.0f;
    if (y >= 0 && y < H && x >= 0 && x < W) v = dsum[(b * C + c) * HW + y * W + x];
    dtaps[pix * ld + col] = __float2bfloat16(v);
  }
}

// 1x1 mix backward, parameter part: dW[o,i] += sum_{b,p} dy[b,o,p]*x[b,i,p],  db[o] += sum dy[b,o,p]
// (the data part dx = W^T dy is rfk_mix1x1 with the transposed matrix).  A CTA walks a slice of pixels, 128 at a time
// through shared memory.  Thread t owns output (t mod n_eff) and pixel segment (t / n_eff): with few channels (C = 4:
// 16 outputs) the 256 threads split each chunk 16 ways instead of idling; with many (C = 64: 4096 outputs) each
// thread owns 16 outputs over the whole chunk.
// Few channels (C = 4, 8: the two largest flow levels): no shared-memory staging at all -- a thread walks pixels (coalesced
// along the pixel index for every channel plane), keeps the C x C outer-product sums in registers, and the sums are
// reduced warp -> CTA (shared-memory atomics) -> global (C*C + C atomics per CTA).
template <int C, int OB>   // OB output rows per pass (C*OB accumulators in registers); C % OB == 0
__global__ void __launch_bounds__(256) mix1x1_wgrad_small_kernel(const float* __restrict__ x, const float* __restrict__ dy, int HW,
                                                                 long long npix, float* __restrict__ dW, float* __restrict__ db) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[C * C + C];
  for (int i = threadIdx.x; i < C * C + C; i += blockDim.x) red[i] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int o0 = 0; o0 < C; o0 += OB) {
    float acc[OB][C], accb[OB];
#pragma unroll
    for (int o = 0; o < OB; ++o) {
      accb[o] = 0.0f;
#pragma unroll
      for (int i = 0; i < C; ++i) acc[o][i] = 0.0f;
    }
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
      const long long b = p / HW;
      const long long base = b * C * HW + (p - b * HW);
      float xv[C], dv[OB];
#pragma unroll
      for (int c = 0; c < C; ++c) xv[c] = __ldg(x + base + (long long)c * HW);
#pragma unroll
      for (int o = 0; o < OB; ++o) dv[o] = __ldg(dy + base + (long long)(o0 + o) * HW);
#pragma unroll
      for (int o = 0; o < OB; ++o) {
        accb[o] += dv[o];
#pragma unroll
        for (int i = 0; i < C; ++i) acc[o][i] = fmaf(dv[o], xv[i], acc[o][i]);
      }
    }
#pragma unroll
    for (int o = 0; o < OB; ++o) {
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const float v = warp_sum(acc[o][i]);
        if (lane == 0) atomicAdd(&red[(o0 + o) * C + i], v);
      }
      const float vb = warp_sum(accb[o]);
      if (lane == 0) atomicAdd(&red[C * C + o0 + o], vb);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C + C; i += blockDim.x) atomicAdd(i < C * C ? dW + i : db + (i - C * C), red[i]);
}

constexpr int MW_CHUNK = 128;
__global__ void __launch_bounds__(256) mix1x1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int C,
                                                           int HW, long long npix, long long pix_per_cta,
                                                           float* __restrict__ dW, float* __restrict__ db) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* xs = sm;                    // [C][MW_CHUNK + 1]
  float* ds = sm + C * (MW_CHUNK + 1);
  constexpr int LD = MW_CHUNK + 1;   // odd stride: rows of different channels fall in different banks
  const long long p0 = blockIdx.x * pix_per_cta, p1 = min(npix, p0 + pix_per_cta);
  const int n_out = C * C;
  const int n_eff = n_out < 256 ? n_out : 256;
  const int nseg = 256 / n_eff;                       // pixel segments per chunk (1 when n_out >= 256)
  const int seg = threadIdx.x / n_eff, oid = threadIdx.x % n_eff;
  const bool active = seg < nseg;
  constexpr int kMax = 16;                            // outputs per thread when n_out > 256 (C*C <= 4096)
  float acc[kMax];
#pragma unroll
  for (int k = 0; k < kMax; ++k) acc[k] = 0.0f;
  float accb = 0.0f;
  for (long long pc = p0; pc < p1; pc += MW_CHUNK) {
    for (int e = threadIdx.x; e < C * MW_CHUNK; e += blockDim.x) {
      const int c = e / MW_CHUNK, r = e % MW_CHUNK;
      const long long p = pc + r;
      float xv = 0.0f, dv = 0.0f;
      if (p < p1) {
        const long long b = p / HW;
        const int q = (int)(p % HW);
        xv = x[(b * C + c) * HW + q];
        dv = dy[(b * C + c) * HW + q];
      }
      xs[c * LD + r] = xv;
      ds[c * LD + r] = dv;
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int k = 0; k < kMax; ++k) {
        const int idx = oid + k * 256;
        if (idx < n_out) {
          const int o = idx / C, i = idx - o * C;
          float s = 0.0f;
          for (int r = seg; r < MW_CHUNK; r += nseg) s = fmaf(ds[o * LD + r], xs[i * LD + r], s);
          acc[k] += s;
        }
      }
    }
    {  // bias: thread t -> channel t mod C, pixel lane t / C
      const int c = threadIdx.x % C, l = threadIdx.x / C, nl = 256 / C;
      if (l < nl) {
        float s = 0.0f;
        for (int r = l; r < MW_CHUNK; r += nl) s += ds[c * LD + r];
        accb += s;
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < kMax; ++k) {
      const int idx = oid + k * 256;
      if (idx < n_out) atomicAdd(dW + idx, acc[k]);
    }
  }
  if (threadIdx.x / C < 256 / C) atomicAdd(db + threadIdx.x % C, accb);
}

// Gaussian log-density backward (forward = gauss_logp_kernel): upstream g[b] on logdet[b] += sum log N(z; mean, std(raw)).
//   dz += g*(-(z-mean)/std^2);  dmean = g*(z-mean)/std^2;  draw = g*((z-mean)^2/std^3 - 1/std)*dstd/draw
__global__ void __launch_bounds__(256) gauss_logp_bwd_kernel(const float* __restrict__ z, int z_C, int z_off,
                                                             const float* __restrict__ params, int n, int HW, int pairing,
                                                             int std_kind, const float* __restrict__ g,
                                                             float* __restrict__ dz, float* __restrict__ dparams) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const long long per = (long long)n * HW;
  const float gb = g[b];
  const float* zb = z + ((long long)b * z_C + z_off) * HW;
  float* dzb = dz + ((long long)b * z_C + z_off) * HW;
  const float* pb = params ? params + (long long)b * 2 * n * HW : nullptr;
  float* dpb = dparams ? dparams + (long long)b * 2 * n * HW : nullptr;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < per; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t / HW), p = (int)(t % HW);
    const int cm = pairing == RFK_PAIR_CROSS ? 2 * j : j, cr = pairing == RFK_PAIR_CROSS ? 2 * j + 1 : n + j;
    float mean = 0.0f, raw = 0.0f;
    if (pb) { mean = pb[(long long)cm * HW + p]; raw = pb[(long long)cr * HW + p]; }
    const float sd = std_from_raw(raw, std_kind);
    const float dstd = std_kind == RFK_STD_EXP ? sd : 1.0f / (1.0f + expf(-raw));   // d std / d raw
    const float d = zb[t] - mean, inv = 1.0f / sd;
    const float gm = gb * d * inv * inv;
    dzb[t] += -gm;
    if (dpb) {
      dpb[(long long)cm * HW + p] = gm;
      dpb[(long long)cr * HW + p] = gb * (d * d * inv * inv * inv - inv) * dstd;
    }
  }
}

}  // namespace rfk

using namespace rfk;

extern "C" int rfk_coupling_taps_bwd(const float* taps, const float* z_out, float* dz, float* dsum, int B, int C, int H,
                                     int W, const float* scale, const float* shift, int clamp_type, const float* clamp_scale,
                                     const float* clamp_shift, const float* g_ld, float* d_scale, float* d_shift,
                                     float* d_clamp_scale, float* d_clamp_shift, float logs_factor, void* stream) {
  RFK_REQUIRE(taps && z_out && dz && dsum && scale && shift && d_scale && d_shift && B > 0 && C > 0 && C % 2 == 0 && H > 0 && W > 0,
              "rfk_coupling_taps_bwd: null pointer or bad shape");
  RFK_REQUIRE(clamp_type >= 0 && clamp_type <= 3, "rfk_coupling_taps_bwd: unknown clamp_type %d", clamp_type);
  RFK_REQUIRE(clamp_type != RFK_CLAMP_REALNVP || (clamp_scale && clamp_shift && d_clamp_scale && d_clamp_shift),
              "rfk_coupling_taps_bwd: realnvp clamp needs scale/scale_shift and their gradient buffers");
  RFK_REQUIRE(B <= 65535 && C / 2 <= 65535, "rfk_coupling_taps_bwd: B or C too large for the grid");
  const int cap = std::max(1, ceil_div((long long)sm_count() * 8, C / 2));
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (W % 4 == 0 && al16(taps) && al16(z_out) && al16(dz) && al16(dsum)) {
    const int chunks4 = std::min(cap, ceil_div((long long)B * H * W / 4, 256));
    RFK_LAUNCH(coupling_taps_bwd_v4_kernel, dim3(chunks4, C / 2), 256, 0, (cudaStream_t)stream, taps, z_out, dz, dsum, B, C, H, W,
               scale, shift, clamp_type, clamp_scale, clamp_shift, g_ld, d_scale, d_shift, d_clamp_scale, d_clamp_shift, logs_factor);
    return check_launch("rfk_coupling_taps_bwd");
  }
  int chunks = ceil_div((long long)B * H * W, 256);
  if (chunks > cap) chunks = cap;
  RFK_LAUNCH(coupling_taps_bwd_kernel, dim3(chunks, C / 2), 256, 0, (cudaStream_t)stream, taps, z_out, dz, dsum, B, C, H, W,
             scale, shift, clamp_type, clamp_scale, clamp_shift, g_ld, d_scale, d_shift, d_clamp_scale, d_clamp_shift, logs_factor);
  return check_launch("rfk_coupling_taps_bwd");
}

extern "C" int rfk_taps_scatter(const float* dsum, void* dtaps, int ld, int B, int C, int H, int W, void* stream) {
  RFK_REQUIRE(dsum && dtaps && B > 0 && C > 0 && H > 0 && W > 0 && ld >= 9 * C, "rfk_taps_scatter: null pointer or bad shape");
  const long long total = (long long)B * H * W * 9 * C;
  RFK_LAUNCH(taps_scatter_kernel, stream_grid(total, 256, 8), 256, 0, (cudaStream_t)stream, dsum, (__nv_bfloat16*)dtaps, ld, C,
             H, W, total);
  return check_launch("rfk_taps_scatter");
}

// Register-blocked variant for C % 4 == 0 (C = 12 ... 64): a thread owns a 4 x 4 block of dW and a share of the chunk's pixels,
// so every pixel costs it 8 shared-memory loads for 16 FMAs (the kernel above does 2 loads per FMA and is LSU-bound on the
// deep levels: 60-80 us for a few thousand pixels).
__global__ void __launch_bounds__(256) mix1x1_wgrad_blocked_kernel(const float* __restrict__ x, const float* __restrict__ dy, int C,
                                                                   int HW, long long npix, long long pix_per_cta,
                                                                   float* __restrict__ dW, float* __restrict__ db) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  constexpr int LD = MW_CHUNK + 1;
  float* xs = sm;
  float* ds = sm + C * LD;
  const long long p0 = blockIdx.x * pix_per_cta, p1 = min(npix, p0 + pix_per_cta);
  const int Cq = C >> 2, nb = Cq * Cq;                 // 4 x 4 output blocks
  const int nseg = max(1, 256 / nb);                   // pixel segments per chunk
  const int blk = threadIdx.x % nb, seg = threadIdx.x / nb;
  const bool active = threadIdx.x < nb * nseg;
  const int bo = blk / Cq, bi = blk - bo * Cq;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
  float accb = 0.0f;
  for (long long pc = p0; pc < p1; pc += MW_CHUNK) {
    for (int e = threadIdx.x; e < C * MW_CHUNK; e += blockDim.x) {
      const int c = e / MW_CHUNK, r = e % MW_CHUNK;
      const long long p = pc + r;
      float xv = 0.0f, dv = 0.0f;
      if (p < p1) {
        const long long b = p / HW;
        const int q = (int)(p - b * HW);
        xv = x[(b * C + c) * HW + q];
        dv = dy[(b * C + c) * HW + q];
      }
      xs[c * LD + r] = xv;
      ds[c * LD + r] = dv;
    }
    __syncthreads();
    if (active) {
      const float* dr = ds + 4 * bo * LD;
      const float* xr = xs + 4 * bi * LD;
      for (int r = seg; r < MW_CHUNK; r += nseg) {
        const float d0 = dr[r], d1 = dr[LD + r], d2 = dr[2 * LD + r], d3 = dr[3 * LD + r];
        const float x0 = xr[r], x1 = xr[LD + r], x2 = xr[2 * LD + r], x3 = xr[3 * LD + r];
        acc[0][0] = fmaf(d0, x0, acc[0][0]); acc[0][1] = fmaf(d0, x1, acc[0][1]); acc[0][2] = fmaf(d0, x2, acc[0][2]); acc[0][3] = fmaf(d0, x3, acc[0][3]);
        acc[1][0] = fmaf(d1, x0, acc[1][0]); acc[1][1] = fmaf(d1, x1, acc[1][1]); acc[1][2] = fmaf(d1, x2, acc[1][2]); acc[1][3] = fmaf(d1, x3, acc[1][3]);
        acc[2][0] = fmaf(d2, x0, acc[2][0]); acc[2][1] = fmaf(d2, x1, acc[2][1]); acc[2][2] = fmaf(d2, x2, acc[2][2]); acc[2][3] = fmaf(d2, x3, acc[2][3]);
        acc[3][0] = fmaf(d3, x0, acc[3][0]); acc[3][1] = fmaf(d3, x1, acc[3][1]); acc[3][2] = fmaf(d3, x2, acc[3][2]); acc[3][3] = fmaf(d3, x3, acc[3][3]);
      }
    }
    {
      const int c = threadIdx.x % C, l = threadIdx.x / C, nl = 256 / C;
      if (l < nl) {
        float sb = 0.0f;
        for (int r = l; r < MW_CHUNK; r += nl) sb += ds[c * LD + r];
        accb += sb;
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) atomicAdd(dW + (4 * bo + a) * C + 4 * bi + b, acc[a][b]);
  }
  if (threadIdx.x / C < 256 / C) atomicAdd(db + threadIdx.x % C, accb);
}

extern "C" int rfk_mix1x1_wgrad(const float* x, const float* dy, int B, int C, int HW, float* dW, float* db, void* stream) {
  RFK_REQUIRE(x && dy && dW && db && B > 0 && C > 0 && C <= 64 && HW > 0, "rfk_mix1x1_wgrad: null pointer or bad shape (C <= 64)");
  const long long npix = (long long)B * HW;
  if (C == 4 || C == 8 || (C == 12 && npix >= 65536)) {
    const int grid = (int)std::min<long long>((long long)sm_count() * 2, (npix + 255) / 256);
    if (C == 4) RFK_LAUNCH((mix1x1_wgrad_small_kernel<4, 4>), grid, 256, 0, (cudaStream_t)stream, x, dy, HW, npix, dW, db);
    else if (C == 8) RFK_LAUNCH((mix1x1_wgrad_small_kernel<8, 8>), grid, 256, 0, (cudaStream_t)stream, x, dy, HW, npix, dW, db);
    else RFK_LAUNCH((mix1x1_wgrad_small_kernel<12, 6>), grid, 256, 0, (cudaStream_t)stream, x, dy, HW, npix, dW, db);
    return check_launch("rfk_mix1x1_wgrad");
  }
  long long ctas = std::min<long long>((long long)sm_count() * 4, (npix + MW_CHUNK - 1) / MW_CHUNK);
  if (ctas < 1) ctas = 1;
  long long ppc = (npix + ctas - 1) / ctas;
  ppc = (ppc + MW_CHUNK - 1) / MW_CHUNK * MW_CHUNK;
  ctas = (npix + ppc - 1) / ppc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mix1x1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * (MW_CHUNK + 1) * 4);
    attr_set = true;
  }
  static const bool blocked = [] { const char* e = getenv("RFK_MIXW_BLOCKED"); return !(e && e[0] == '0'); }();
  if (C % 4 == 0 && C >= 32 && blocked) {   // fewer 4x4 blocks than threads would only multiply the atomics (C=16: 185 vs 17 us)
    static bool battr = false;
    if (!battr) {
      cudaFuncSetAttribute(mix1x1_wgrad_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * (MW_CHUNK + 1) * 4);
      battr = true;
    }
    RFK_LAUNCH(mix1x1_wgrad_blocked_kernel, (int)ctas, 256, (size_t)2 * C * (MW_CHUNK + 1) * sizeof(float), (cudaStream_t)stream, x,
               dy, C, HW, npix, ppc, dW, db);
    return check_launch("rfk_mix1x1_wgrad");
  }
  RFK_LAUNCH(mix1x1_wgrad_kernel, (int)ctas, 256, (size_t)2 * C * (MW_CHUNK + 1) * sizeof(float), (cudaStream_t)stream, x, dy, C,
             HW, npix, ppc, dW, db);
  return check_launch("rfk_mix1x1_wgrad");
}

extern "C" int rfk_gauss_logp_bwd(const float* z, int z_C, int z_off, const float* params, int n, int B, int HW, int pairing,
                                  int std_kind, const float* g, float* dz, float* dparams, void* stream) {
  RFK_REQUIRE(z && g && dz && B > 0 && n > 0 && HW > 0 && z_off >= 0 && z_off + n <= z_C, "rfk_gauss_logp_bwd: bad arguments");
  RFK_REQUIRE((params == nullptr) == (dparams == nullptr), "rfk_gauss_logp_bwd: params and dparams go together");
  long long per = (long long)n * HW;
  int chunks = ceil_div(per, 256);
  int cap = ceil_div((long long)sm_count() * 8, B);
  if (chunks > cap) chunks = cap;
  RFK_LAUNCH(gauss_logp_bwd_kernel, dim3(chunks, B), 256, 0, (cudaStream_t)stream, z, z_C, z_off, params, n, HW, pairing,
             std_kind, g, dz, dparams);
  return check_launch("rfk_gauss_logp_bwd");
}


extern "C" int rfk_act_affine_bwd(const void* dh, const void* h, int ld, int n, const float* scale, int act_fn, void* da,
                                  int da_ld, float* r_dv, float* r_dvv, float dvv_factor, int dv_scaled, long long rows,
                                  void* stream) {
  RFK_REQUIRE(dh && h && da && scale && r_dv && r_dvv && rows > 0 && n > 0, "rfk_act_affine_bwd: null pointer or empty shape");
  const int n8 = (n + 7) / 8 * 8;   // channels are handled in groups of 8; the tail group reads/writes pad columns
  RFK_REQUIRE(n <= 2048 && ld % 8 == 0 && da_ld % 8 == 0 && n8 <= ld && n8 <= da_ld,
              "rfk_act_affine_bwd: n=%d (<= 2048) rounded up to 8 must fit the 8-aligned row strides", n);
  RFK_REQUIRE(((reinterpret_cast<uintptr_t>(dh) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(da)) & 15) == 0,
              "rfk_act_affine_bwd: tensors must be 16-byte aligned");
  RFK_REQUIRE(act_fn >= 0 && act_fn <= 2, "rfk_act_affine_bwd: bad act_fn %d", act_fn);
  RFK_REQUIRE((n8 / 8) <= 256, "rfk_act_affine_bwd: too many channels per row for one CTA");
  const int ctas = (int)std::min<long long>((long long)sm_count() * 4, (rows + 63) / 64);
  const long long rows_per_cta = (rows + ctas - 1) / ctas;
  RFK_LAUNCH(act_affine_bwd_kernel, ctas, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)dh, (const __nv_bfloat16*)h,
             ld, n, scale, act_fn, (__nv_bfloat16*)da, da_ld, r_dv, r_dvv, dvv_factor, dv_scaled, rows, rows_per_cta);
  return check_launch("rfk_act_affine_bwd");
}

extern "C" int rfk_conv_wgrad(const void* x, int x_ld, int cin, const void* dy, int dy_ld, int cout, int B, int H, int W,
                              int taps, float* dw, int dw_ld, int layout, const int* perm, void* ws, long long ws_bytes,
                              void* stream) {
  RFK_REQUIRE(layout == 0 || layout == 1, "rfk_conv_wgrad: layout must be 0 ([taps][cout][ld]) or 1 ([cout][ld][taps])");
  RFK_REQUIRE(layout == 1 || perm == nullptr, "rfk_conv_wgrad: a channel permutation needs layout 1");
  RFK_REQUIRE(x && dy && dw && B > 0 && H > 0 && W > 0 && cin > 0 && cout > 0, "rfk_conv_wgrad: null pointer or empty shape");
  RFK_REQUIRE(taps == 1 || taps == 9, "rfk_conv_wgrad: taps=%d (only 1x1 and 3x3 kernels)", taps);
  RFK_REQUIRE(x_ld % 8 == 0 && dy_ld % 8 == 0 && cin <= x_ld && cout <= dy_ld && dw_ld >= cin,
              "rfk_conv_wgrad: row strides must be multiples of 8 and cover the channels");
  RFK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0,
              "rfk_conv_wgrad: x / dy must be 16-byte aligned");
  {
    // tensor-core (tcgen05) path; RFK_WGRAD_WMMA=1 keeps the warp-level mma.sync kernel for A/B comparisons
    static const bool force_wmma = [] { const char* e = getenv("RFK_WGRAD_WMMA"); return e && e[0] == '1'; }();
    if (!force_wmma) {
      const int rc = conv_wgrad_tc(x, x_ld, cin, dy, dy_ld, cout, B, H, W, taps, dw, dw_ld, layout, perm, ws, ws_bytes,
                                   (cudaStream_t)stream);
      if (rc <= 0) return rc;   // ran (0) or failed (<0); positive = shape not covered, fall through
    }
  }
  const long long npix = (long long)B * H * W;
  const int c_tiles = (cin + WG_C - 1) / WG_C, n_tiles = (cout + WG_N - 1) / WG_N;
  const long long base_ctas = (long long)taps * c_tiles * n_tiles;
  long long slices = std::max<long long>(1, ((long long)sm_count() * 4) / base_ctas);
  const long long max_slices = (npix + 4 * WG_P - 1) / (4 * WG_P);
  if (slices > max_slices) slices = max_slices;
  if (slices > 65535) slices = 65535;
  long long pps = (npix + slices - 1) / slices;
  pps = (pps + WG_P - 1) / WG_P * WG_P;
  slices = (npix + pps - 1) / pps;
  dim3 grid((unsigned)(taps * c_tiles), (unsigned)n_tiles, (unsigned)slices);
  RFK_LAUNCH(conv_wgrad_kernel, grid, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, x_ld, cin,
             (const __nv_bfloat16*)dy, dy_ld, cout, B, H, W, taps, dw, dw_ld, layout, perm, pps);
  return check_launch("rfk_conv_wgrad");
}


// ------------------------------------------------------------------------------------------
// Adam over ALL parameters of the model in one launch (torch.optim.Adam semantics without weight decay / amsgrad:
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)).
// p, g, m, v are flat fp32 buffers (the parameters are views into p); g is scaled by grad_scale first (1/world after a
// sum-allreduce); the step count t lives on the device so that the launch can be replayed from a CUDA graph.
// HBM-bound: 4 reads + 3 writes of 4 bytes per parameter.
// ------------------------------------------------------------------------------------------
namespace rfk {
__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                   float4* __restrict__ v, long long n4, float lr, float b1, float b2, float eps,
                                                   float grad_scale, const float* __restrict__ step) {
  pdl_trigger();
  pdl_wait();
  const float t = *step;
  const float step_size = lr / (1.0f - powf(b1, t));
  const float inv_bc2 = rsqrtf(1.0f - powf(b2, t));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = p[i], gv = g[i], mv = m[i], vv = v[i];
    float* pp = reinterpret_cast<float*>(&pv);
    float* gp = reinterpret_cast<float*>(&gv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gg = gp[k] * grad_scale;
      mp[k] = b1 * mp[k] + (1.0f - b1) * gg;
      vp[k] = b2 * vp[k] + (1.0f - b2) * gg * gg;
      pp[k] -= step_size * mp[k] / (sqrtf(vp[k]) * inv_bc2 + eps);
    }
    p[i] = pv; m[i] = mv; v[i] = vv;
  }
}
}  // namespace rfk

extern "C" int rfk_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                             float eps, float grad_scale, const float* step, void* stream) {
  using namespace rfk;
  RFK_REQUIRE(p && g && m && v && step && n > 0 && n % 4 == 0, "rfk_adam_step: null pointer, or n=%lld is not a multiple of 4", n);
  RFK_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                reinterpret_cast<uintptr_t>(v)) & 15) == 0, "rfk_adam_step: buffers must be 16-byte aligned");
  RFK_LAUNCH(adam_kernel, stream_grid(n / 4, 256, 8), 256, 0, (cudaStream_t)stream, (float4*)p, (const float4*)g, (float4*)m,
             (float4*)v, n / 4, lr, beta1, beta2, eps, grad_scale, step);
  return check_launch("rfk_adam_step");
}


// ------------------------------------------------------------------------------------------
// ConvLSTM cell update, backward (forward = lstm_point in elementwise.cu; Utils/modules.py:369-377).
//   i = s(ai + wi c_prev), f = s(af + wf c_prev), g = tanh(ag), c = f c_prev + i g, o = s(ao + wo c), h = o tanh(c)
// Gates are recomputed from the saved pre-activations cc [B,4Hc,H,W] (bias included) and c_prev.  Given dh (gradient
// w.r.t. h, batch-strided) and dc_in (w.r.t. c, nullable):
//   dcc [B,4Hc,H,W] (w.r.t. the pre-activations, order i,f,o,g), dc_prev [B,Hc,H,W], dbias[4Hc] += sum_{b,p} dcc.
// One CTA row (blockIdx.y) per hidden channel, so the bias reduction is a block reduction + 4 atomics.
// The peephole tensors are constants in the reference (never registered as trainable parameters).
// ------------------------------------------------------------------------------------------
namespace rfk {
__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + __expf(-x)); }

struct LstmBwdOut { float dai, daf, dao, dag, dcp; };
__device__ __forceinline__ LstmBwdOut lstm_point_bwd(float ai, float af, float ao, float ag, float cp, float wi, float wf,
                                                     float wo, float dhv, float dci) {
  const float i = sigm(ai + wi * cp), f = sigm(af + wf * cp), g = tanhf(ag);
  const float c = f * cp + i * g;
  const float o = sigm(ao + wo * c), tc = tanhf(c);
  LstmBwdOut r;
  r.dao = dhv * tc * o * (1.0f - o);
  const float dct = dci + dhv * o * (1.0f - tc * tc) + r.dao * wo;
  r.dai = dct * g * i * (1.0f - i);
  r.daf = dct * cp * f * (1.0f - f);
  r.dag = dct * i * (1.0f - g * g);
  r.dcp = dct * f + r.dai * wi + r.daf * wf;
  return r;
}

// kVec: HW % 4 == 0 and 16-byte aligned tensors -> four consecutive positions per thread, 128-bit accesses
template <bool kVec>
__global__ void __launch_bounds__(256) lstm_pointwise_bwd_kernel(const float* __restrict__ cc, const float* __restrict__ c_prev,
                                                                 const float* __restrict__ peep, const float* __restrict__ dh,
                                                                 long long dh_bs, const float* __restrict__ dc_in,
                                                                 float* __restrict__ dcc, float* __restrict__ dc_prev,
                                                                 float* __restrict__ dbias, int B, int Hc, int HW) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int ch = blockIdx.y;
  const long long per = (long long)Hc * HW;
  constexpr int V = kVec ? 4 : 1;
  const int HWv = HW / V;
  const long long n = (long long)B * HWv;
  float acc[4] = {0, 0, 0, 0};
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / HWv;
    const int p = (int)(idx - b * HWv) * V;
    const long long r = (long long)ch * HW + p, e = b * per + r;
    const float* g0 = cc + b * 4 * per + r;
    float* d0 = dcc + b * 4 * per + r;
    if (kVec) {
      const float4 ai = ld_stream(reinterpret_cast<const float4*>(g0)), af = ld_stream(reinterpret_cast<const float4*>(g0 + per));
      const float4 ao = ld_stream(reinterpret_cast<const float4*>(g0 + 2 * per)), ag = ld_stream(reinterpret_cast<const float4*>(g0 + 3 * per));
      const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      const float4 cp = c_prev ? ld_stream(reinterpret_cast<const float4*>(c_prev + e)) : z4;
      const float4 dci = dc_in ? ld_stream(reinterpret_cast<const float4*>(dc_in + e)) : z4;
      const float4 dhv = ld_stream(reinterpret_cast<const float4*>(dh + b * dh_bs + r));
      float4 wi = z4, wf = z4, wo = z4;
      if (peep) {
        wi = *reinterpret_cast<const float4*>(peep + r);
        wf = *reinterpret_cast<const float4*>(peep + per + r);
        wo = *reinterpret_cast<const float4*>(peep + 2 * per + r);
      }
      const LstmBwdOut o0 = lstm_point_bwd(ai.x, af.x, ao.x, ag.x, cp.x, wi.x, wf.x, wo.x, dhv.x, dci.x);
      const LstmBwdOut o1 = lstm_point_bwd(ai.y, af.y, ao.y, ag.y, cp.y, wi.y, wf.y, wo.y, dhv.y, dci.y);
      const LstmBwdOut o2 = lstm_point_bwd(ai.z, af.z, ao.z, ag.z, cp.z, wi.z, wf.z, wo.z, dhv.z, dci.z);
      const LstmBwdOut o3 = lstm_point_bwd(ai.w, af.w, ao.w, ag.w, cp.w, wi.w, wf.w, wo.w, dhv.w, dci.w);
      st_stream(reinterpret_cast<float4*>(d0), make_float4(o0.dai, o1.dai, o2.dai, o3.dai));
      st_stream(reinterpret_cast<float4*>(d0 + per), make_float4(o0.daf, o1.daf, o2.daf, o3.daf));
      st_stream(reinterpret_cast<float4*>(d0 + 2 * per), make_float4(o0.dao, o1.dao, o2.dao, o3.dao));
      st_stream(reinterpret_cast<float4*>(d0 + 3 * per), make_float4(o0.dag, o1.dag, o2.dag, o3.dag));
      st_stream(reinterpret_cast<float4*>(dc_prev + e), make_float4(o0.dcp, o1.dcp, o2.dcp, o3.dcp));
      acc[0] += (o0.dai + o1.dai) + (o2.dai + o3.dai);
      acc[1] += (o0.daf + o1.daf) + (o2.daf + o3.daf);
      acc[2] += (o0.dao + o1.dao) + (o2.dao + o3.dao);
      acc[3] += (o0.dag + o1.dag) + (o2.dag + o3.dag);
    } else {
      float wi = 0.0f, wf = 0.0f, wo = 0.0f;
      if (peep) { wi = peep[r]; wf = peep[per + r]; wo = peep[2 * per + r]; }
      const LstmBwdOut o = lstm_point_bwd(g0[0], g0[per], g0[2 * per], g0[3 * per], c_prev ? c_prev[e] : 0.0f, wi, wf, wo,
                                          dh[b * dh_bs + r], dc_in ? dc_in[e] : 0.0f);
      d0[0] = o.dai; d0[per] = o.daf; d0[2 * per] = o.dao; d0[3 * per] = o.dag;
      dc_prev[e] = o.dcp;
      acc[0] += o.dai; acc[1] += o.daf; acc[2] += o.dao; acc[3] += o.dag;
    }
  }
  if (dbias) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float r = block_sum_256(acc[k], sh);
      if (threadIdx.x == 0) atomicAdd(dbias + k * Hc + ch, r);
    }
  }
}
}  // namespace rfk

extern "C" int rfk_convlstm_pointwise_bwd(const float* cc, const float* c_prev, const float* peep, const float* dh,
                                          long long dh_bstride, const float* dc_in, float* dcc, float* dc_prev, float* dbias,
                                          int B, int Hc, int HW, void* stream) {
  using namespace rfk;
  RFK_REQUIRE(cc && dh && dcc && dc_prev && B > 0 && Hc > 0 && HW > 0 && Hc <= 65535,
              "rfk_convlstm_pointwise_bwd: null pointer or bad shape");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec = HW % 4 == 0 && dh_bstride % 4 == 0 && al16(cc) && al16(dh) && al16(dcc) && al16(dc_prev) &&
                   (!c_prev || al16(c_prev)) && (!dc_in || al16(dc_in)) && (!peep || al16(peep));
  int chunks = ceil_div((long long)B * HW / (vec ? 4 : 1), 256);
  const int cap = std::max(1, ceil_div((long long)sm_count() * 8, Hc));
  if (chunks > cap) chunks = cap;
  if (vec)
    RFK_LAUNCH((lstm_pointwise_bwd_kernel<true>), dim3(chunks, Hc), 256, 0, (cudaStream_t)stream, cc, c_prev, peep, dh, dh_bstride,
               dc_in, dcc, dc_prev, dbias, B, Hc, HW);
  else
    RFK_LAUNCH((lstm_pointwise_bwd_kernel<false>), dim3(chunks, Hc), 256, 0, (cudaStream_t)stream, cc, c_prev, peep, dh, dh_bstride,
               dc_in, dcc, dc_prev, dbias, B, Hc, HW);
  return check_launch("rfk_convlstm_pointwise_bwd");
}
