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

#include "common.cuh"

namespace rfk {

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
                                                             long long rows, long long rows_per_cta) {
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
    if (c < n) {
      atomicAdd(r_dv + c, t0);
      atomicAdd(r_dvv + c, t1);
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
                                                         int dw_ld, long long pix_per_slice) {
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
      if (n < cout && c < cin) atomicAdd(dw + ((long long)tap * cout + n) * dw_ld + c, so[warp][e]);
    }
    __syncwarp();
  }
}

}  // namespace rfk

using namespace rfk;

extern "C" int rfk_act_affine_bwd(const void* dh, const void* h, int ld, int n, const float* scale, int act_fn, void* da,
                                  int da_ld, float* r_dv, float* r_dvv, long long rows, void* stream) {
  RFK_REQUIRE(dh && h && da && scale && r_dv && r_dvv && rows > 0 && n > 0, "rfk_act_affine_bwd: null pointer or empty shape");
  RFK_REQUIRE(n % 8 == 0 && n <= 2048 && ld % 8 == 0 && da_ld % 8 == 0 && n <= ld && n <= da_ld,
              "rfk_act_affine_bwd: n=%d must be a multiple of 8 (<= 2048) within 8-aligned row strides", n);
  RFK_REQUIRE(((reinterpret_cast<uintptr_t>(dh) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(da)) & 15) == 0,
              "rfk_act_affine_bwd: tensors must be 16-byte aligned");
  RFK_REQUIRE(act_fn >= 0 && act_fn <= 2, "rfk_act_affine_bwd: bad act_fn %d", act_fn);
  RFK_REQUIRE((n / 8) <= 256, "rfk_act_affine_bwd: too many channels per row for one CTA");
  const int ctas = (int)std::min<long long>((long long)sm_count() * 4, (rows + 63) / 64);
  const long long rows_per_cta = (rows + ctas - 1) / ctas;
  RFK_LAUNCH(act_affine_bwd_kernel, ctas, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)dh, (const __nv_bfloat16*)h,
             ld, n, scale, act_fn, (__nv_bfloat16*)da, da_ld, r_dv, r_dvv, rows, rows_per_cta);
  return check_launch("rfk_act_affine_bwd");
}

extern "C" int rfk_conv_wgrad(const void* x, int x_ld, int cin, const void* dy, int dy_ld, int cout, int B, int H, int W,
                              int taps, float* dw, int dw_ld, void* stream) {
  RFK_REQUIRE(x && dy && dw && B > 0 && H > 0 && W > 0 && cin > 0 && cout > 0, "rfk_conv_wgrad: null pointer or empty shape");
  RFK_REQUIRE(taps == 1 || taps == 9, "rfk_conv_wgrad: taps=%d (only 1x1 and 3x3 kernels)", taps);
  RFK_REQUIRE(x_ld % 8 == 0 && dy_ld % 8 == 0 && cin <= x_ld && cout <= dy_ld && dw_ld >= cin,
              "rfk_conv_wgrad: row strides must be multiples of 8 and cover the channels");
  RFK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0,
              "rfk_conv_wgrad: x / dy must be 16-byte aligned");
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
             (const __nv_bfloat16*)dy, dy_ld, cout, B, H, W, taps, dw, dw_ld, pps);
  return check_launch("rfk_conv_wgrad");
}
