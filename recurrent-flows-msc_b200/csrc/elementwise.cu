// Bandwidth-bound kernels of the Glow step and the ConvLSTM cell (sm_100a).
// Every kernel streams its tensors once with 128-bit accesses where the shape allows it and
// reduces per-sample sums with warp shuffles + one atomic per CTA.
#include <cstdlib>

#include "common.cuh"

namespace rfk {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------
// a4  Squeeze2d  (Flow/glow_modules.py:298-310)
// ------------------------------------------------------------------------------------------
// forward, W % 8 == 0: one thread reads 8 consecutive pixels of one input row (2 x 128 bit) and
// writes 4 pixels to each of the two dx planes (1 x 128 bit each).
__global__ void __launch_bounds__(kThreads) squeeze_fwd_v8(const float* __restrict__ x, float* __restrict__ y,
                                                           int C, int H, int W, long long total8) {
  pdl_trigger();
  pdl_wait();
  const int W8 = W >> 3, Ho = H >> 1, Wo = W >> 1;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total8;
       t += (long long)gridDim.x * blockDim.x) {
    int x8 = (int)(t % W8);
    long long r = t / W8;
    int row = (int)(r % H);
    long long bc = r / H;  // b*C + c
    const float4* src = reinterpret_cast<const float4*>(x + (bc * H + row) * W + x8 * 8);
    float4 a = ld_stream(src), b = ld_stream(src + 1);
    int dy = row & 1, i = row >> 1;
    long long b_ = bc / C;
    int c = (int)(bc % C);
    float* d0 = y + (((b_ * 4 * C + 4 * c + 2 * dy) * Ho + i) * (long long)Wo) + x8 * 4;
    st_stream(reinterpret_cast<float4*>(d0), make_float4(a.x, a.z, b.x, b.z));
    st_stream(reinterpret_cast<float4*>(d0 + (long long)Ho * Wo), make_float4(a.y, a.w, b.y, b.w));
  }
}

// undo, Wout % 8 == 0: inverse of the above (reads 2 x 128 bit from the two dx planes, writes 2 x 128 bit)
__global__ void __launch_bounds__(kThreads) squeeze_undo_v8(const float* __restrict__ x, float* __restrict__ y,
                                                            int Co, int Hout, int Wout, long long total8) {
  pdl_trigger();
  pdl_wait();
  const int W8 = Wout >> 3, Hi = Hout >> 1, Wi = Wout >> 1;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total8;
       t += (long long)gridDim.x * blockDim.x) {
    int x8 = (int)(t % W8);
    long long r = t / W8;
    int row = (int)(r % Hout);
    long long bc = r / Hout;  // b*Co + c  (output channel)
    int dy = row & 1, i = row >> 1;
    long long b_ = bc / Co;
    int c = (int)(bc % Co);
    const float* s0 = x + (((b_ * 4 * Co + 4 * c + 2 * dy) * Hi + i) * (long long)Wi) + x8 * 4;
    float4 e = ld_stream(reinterpret_cast<const float4*>(s0));
    float4 o = ld_stream(reinterpret_cast<const float4*>(s0 + (long long)Hi * Wi));
    float4* dst = reinterpret_cast<float4*>(y + (bc * Hout + row) * Wout + x8 * 8);
    st_stream(dst, make_float4(e.x, o.x, e.y, o.y));
    st_stream(dst + 1, make_float4(e.z, o.z, e.w, o.w));
  }
}

// generic scalar version for narrow maps (W in {2,4,6,...}); indexed by OUTPUT element
__global__ void __launch_bounds__(kThreads) squeeze_scalar(const float* __restrict__ x, float* __restrict__ y,
                                                           int C, int H, int W, int undo, long long total) {
  pdl_trigger();
  pdl_wait();
  // C,H,W describe the un-squeezed ("big") tensor [B,C,H,W]; the squeezed one is [B,4C,H/2,W/2]
  const int Ho = H >> 1, Wo = W >> 1;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    // decode t as an index into the squeezed tensor
    int j = (int)(t % Wo);
    long long r = t / Wo;
    int i = (int)(r % Ho);
    r /= Ho;
    int cs = (int)(r % (4 * C));
    long long b = r / (4 * C);
    int c = cs >> 2, dy = (cs >> 1) & 1, dx = cs & 1;
    long long big = ((b * C + c) * H + 2 * i + dy) * (long long)W + 2 * j + dx;
    if (undo) y[big] = x[t]; else y[t] = x[big];
  }
}

// ------------------------------------------------------------------------------------------
// a1  ActNorm apply  (Flow/glow_modules.py:38-54)
// ------------------------------------------------------------------------------------------
template <bool kVec>
__global__ void __launch_bounds__(kThreads) actnorm_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           const float* __restrict__ bias,
                                                           const float* __restrict__ logs, int C, int HW,
                                                           int reverse, long long total) {
  pdl_trigger();
  pdl_wait();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long e = kVec ? t * 4 : t;
    int c = (int)((e / HW) % C);
    float b = __ldg(bias + c), l = __ldg(logs + c);
    if (kVec) {
      float4 v = ld_stream(reinterpret_cast<const float4*>(x) + t);
      if (!reverse) {
        float s = expf(l);
        v.x = (v.x + b) * s; v.y = (v.y + b) * s; v.z = (v.z + b) * s; v.w = (v.w + b) * s;
      } else {
        float s = expf(-l);
        v.x = v.x * s - b; v.y = v.y * s - b; v.z = v.z * s - b; v.w = v.w * s - b;
      }
      st_stream(reinterpret_cast<float4*>(y) + t, v);
    } else {
      float v = x[t];
      y[t] = reverse ? v * expf(-l) - b : (v + b) * expf(l);
    }
  }
}

// ------------------------------------------------------------------------------------------
// a1  ActNorm data-dependent init  (Flow/glow_modules.py:26-31): one CTA per channel,
// two passes (mean, then centred sum of squares) accumulated in double.
// ------------------------------------------------------------------------------------------
__device__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  int nw = blockDim.x >> 5;
  if (threadIdx.x < 32) {
    r = lane < nw ? sh[lane] : 0.0;
    r = warp_sum(r);
    if (lane == 0) sh[32] = r;
  }
  __syncthreads();
  return sh[32];
}

__global__ void __launch_bounds__(512) actnorm_init_kernel(const float* __restrict__ x, float* __restrict__ bias,
                                                           float* __restrict__ logs, float* __restrict__ mean_out,
                                                           float* __restrict__ std_out, int B, int C, int HW) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sh[33];
  const int c = blockIdx.x;
  const long long n = (long long)B * HW;
  double s = 0.0;
  for (long long t = threadIdx.x; t < n; t += blockDim.x) {
    long long b = t / HW;
    int p = (int)(t % HW);
    s += (double)x[(b * C + c) * HW + p];
  }
  const double mean = block_sum(s, sh) / (double)n;
  double q = 0.0;
  for (long long t = threadIdx.x; t < n; t += blockDim.x) {
    long long b = t / HW;
    int p = (int)(t % HW);
    double d = (double)x[(b * C + c) * HW + p] - mean;
    q += d * d;
  }
  const double var = block_sum(q, sh) / (double)(n - 1);
  if (threadIdx.x == 0) {
    float sd = (float)sqrt(var);
    if (bias) bias[c] = -(float)mean;
    if (logs) logs[c] = logf(1.0f / (sd + 1e-6f));
    if (mean_out) mean_out[c] = (float)mean;
    if (std_out) std_out[c] = sd;
  }
}

// ------------------------------------------------------------------------------------------
// a2  1x1 channel mix  y[b,o,p] = sum_i Wm[o,i] x[b,i,p] + bvec[o]   (Flow/glow_modules.py:213)
// HBM-bound (AI = C/4 flop/B): one thread per pixel, x tile and Wm staged in shared memory,
// 8 output channels register-blocked.  Optional bf16 NHWC side output of the first side_n channels.
// ------------------------------------------------------------------------------------------
// optional producer for mix1x1_kernel's input: a pending tap-split coupling (rfk_coupling_taps_mix)
struct CouplingSrc {
  const float* taps;    // [B, 9C, H, W] or null = plain mix
  const float* scale;   // Conv2dZeros affine [C]
  const float* shift;
  int clamp_type;
  const float* cs;
  const float* csh;
  float* logdet;        // += (-=) sum of the coupling's log-scales, or null
  int reverse;
  int H, W;
};

// CTA = PT pixels x G output groups: thread (tx, ty) computes outputs [8*ty, 8*ty+8) of pixel tx, reading the
// x tile (staged once in shared memory, conflict-free) and Wm (shared memory broadcast).  Optionally also adds
// alpha * (*addend) to logdet[0..B) -- the parameter-only log-det term of ActNorm + InvConv -- so that a GlowStep
// needs no separate launch for it.
__global__ void __launch_bounds__(1024) mix1x1_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                      const float* __restrict__ Wm, const float* __restrict__ bvec,
                                                      int C, int HW, long long npix, __nv_bfloat16* __restrict__ side,
                                                      int side_n, int side_off, int side_ld, int w_smem,
                                                      float* __restrict__ logdet, const float* __restrict__ addend,
                                                      float alpha, int B, const CouplingSrc cp) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float smem[];
  const int PT = blockDim.x, G = blockDim.y;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * PT + tx, nthr = PT * G;
  float* bs = smem;                 // [C]
  float* xs = bs + C;               // [C][PT]
  const float* ws = Wm;             // [C][C]: shared memory when it fits, else L1-cached broadcast loads
  // issue this thread's x loads first (independent of the weight staging below) so their latency overlaps it
  const long long pix = blockIdx.x * (long long)PT + tx;
  const bool ok = pix < npix;
  const long long b = ok ? pix / HW : 0;
  const int p = ok ? (int)(pix % HW) : 0;
  const float* xp = x + b * C * HW + p;
  float xr[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = ty + k * G;
    xr[k] = (ok && i < C) ? xp[(long long)i * HW] : 0.0f;
  }
  if (cp.taps) {
    // the input is the PREVIOUS coupling's un-finished state: channels >= C/2 still need (z2 + t)*e^{ls} (or the inverse),
    // with (t, ls) gathered from that coupling's nine tap planes -- see coupling_taps_kernel
    const int half = C >> 1, W_ = cp.W, H_ = cp.H;
    const int yy0 = p / W_, xx0 = p - yy0 * W_;
    const float* tb = cp.taps + b * 9 * C * HW;
    float ld_acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = ty + k * G;
      if (ok && i >= half && i < C) {
        const int j = i - half;
        float s_sum = 0.0f, r_sum = 0.0f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = yy0 + ky - 1;
          if (yy < 0 || yy >= H_) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int xx = xx0 + kx - 1;
            if (xx < 0 || xx >= W_) continue;
            const float* q = tb + ((long long)((3 * ky + kx) * C + 2 * j) * HW) + yy * W_ + xx;
            s_sum += __ldg(q);
            r_sum += __ldg(q + HW);
          }
        }
        const float sft = fmaf(s_sum, __ldg(cp.scale + 2 * j), __ldg(cp.shift + 2 * j));
        const float raw = fmaf(r_sum, __ldg(cp.scale + 2 * j + 1), __ldg(cp.shift + 2 * j + 1));
        float a = 0.0f, bb = 0.0f;
        if (cp.clamp_type == RFK_CLAMP_REALNVP) { a = __ldg(cp.cs + j); bb = __ldg(cp.csh + j); }
        const float ls = clamp_ls(raw, cp.clamp_type, a, bb);
        ld_acc += ls;
        xr[k] = cp.reverse ? xr[k] * expf(-ls) - sft : (xr[k] + sft) * expf(ls);
      }
    }
    if (cp.logdet) {
      const float v = cp.reverse ? -ld_acc : ld_acc;
      if ((HW & 31) == 0) {   // a warp's 32 pixels lie in one sample: one atomic per warp
        const float r = warp_sum(ok ? v : 0.0f);
        if ((tx & 31) == 0 && ok) atomicAdd(cp.logdet + b, r);
      } else if (ok && v != 0.0f) {
        atomicAdd(cp.logdet + b, v);
      }
    }
  }
  if (w_smem) {
    float* wsm = smem + (((C + C * PT) + 3) & ~3);   // 16-byte aligned for the 128-bit staging stores
    const int n = C * C;
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(Wm) & 15) == 0) {
      const float4* src = reinterpret_cast<const float4*>(Wm);
      float4* dst = reinterpret_cast<float4*>(wsm);
#pragma unroll 4
      for (int i = tid; i < (n >> 2); i += nthr) dst[i] = __ldg(src + i);
    } else {
#pragma unroll 4
      for (int i = tid; i < n; i += nthr) wsm[i] = Wm[i];
    }
    ws = wsm;
  }
  for (int i = tid; i < C; i += nthr) bs[i] = bvec ? bvec[i] : 0.0f;
  if (logdet && blockIdx.x == 0) {
    const float add = alpha * (*addend);
    for (int i = tid; i < B; i += nthr) atomicAdd(logdet + i, add);   // other CTAs may be adding coupling terms to the same words
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = ty + k * G;
    if (i < C) xs[i * PT + tx] = xr[k];
  }
  __syncthreads();
  const int o0 = ty * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = (o0 + j < C) ? bs[o0 + j] : 0.0f;
  for (int i = 0; i < C; ++i) {
    const float xv = xs[i * PT + tx];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (o0 + j < C) acc[j] = fmaf(ws[(o0 + j) * C + i], xv, acc[j]);
  }
  if (ok) {
    float* yp = y + b * C * HW + p;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = o0 + j;
      if (o < C) {
        yp[(long long)o * HW] = acc[j];
        if (side && o < side_n) side[pix * side_ld + side_off + o] = __float2bfloat16(acc[j]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// layout helpers
// ------------------------------------------------------------------------------------------
// NCHW f32 channel slice -> NHWC bf16.  One thread per (pixel, group of 8 channels): consecutive threads take
// consecutive pixels (coalesced reads of each channel plane), each writes one 16-byte group.
__global__ void __launch_bounds__(kThreads) pack_nhwc_kernel(const float* __restrict__ src, long long src_bs,
                                                             int HW, int c_lo, int n, __nv_bfloat16* __restrict__ dst,
                                                             int dst_off, int dst_ld, long long npix, int vec_ok, int residual) {
  pdl_trigger();
  pdl_wait();
  const int groups = (n + 7) >> 3;
  const long long total = npix * groups;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long pix = t % npix;
    const int j = (int)(t / npix) * 8;
    const long long b = pix / HW;
    const int p = (int)(pix % HW);
    const float* sp = src + b * src_bs + (long long)(c_lo + j) * HW + p;
    __nv_bfloat16* dp = dst + pix * dst_ld + dst_off + j;
    if (residual) {   // split precision: what the bf16 rounding of the value discarded
      for (int k = 0; k < 8 && j + k < n; ++k) {
        const float v = sp[(long long)k * HW];
        dp[k] = __float2bfloat16(v - __bfloat162float(__float2bfloat16(v)));
      }
    } else if (vec_ok && j + 8 <= n) {
      __nv_bfloat162 h[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        h[k] = __floats2bfloat162_rn(sp[(long long)(2 * k) * HW], sp[(long long)(2 * k + 1) * HW]);
      *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<uint4*>(h);
    } else {
      for (int k = 0; k < 8 && j + k < n; ++k) dp[k] = __float2bfloat16(sp[(long long)k * HW]);
    }
  }
}

// Same, through a shared-memory transpose: a CTA takes 64 consecutive pixels x up to 64 channels (blockIdx.y = channel
// chunk); reads are coalesced along the pixels of each channel plane, writes are coalesced along the channels of each
// pixel row (a warp writes whole 128-byte rows instead of thirty-two 16-byte pieces of different rows).
constexpr int PK_PIX = 128;
__global__ void __launch_bounds__(256) pack_nhwc_tiled_kernel(const float* __restrict__ src, long long src_bs, int HW, int c_lo,
                                                              int n, __nv_bfloat16* __restrict__ dst, int dst_off, int dst_ld,
                                                              long long npix, int vec4) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[64][PK_PIX + 1];
  const int c0 = blockIdx.y * 64, nc = min(64, n - c0), G = nc >> 3;   // nc is a multiple of 8 on this path
  for (long long g0 = (long long)blockIdx.x * PK_PIX; g0 < npix; g0 += (long long)gridDim.x * PK_PIX) {
    const long long b0 = g0 / HW;             // one 64-bit division per tile; the rest is 32-bit
    const int p0 = (int)(g0 - b0 * HW);
    const int valid = (int)min((long long)PK_PIX, npix - g0);
    if (vec4) {   // HW % 4 == 0: four consecutive pixels of a plane never straddle a sample, 128-bit loads
      for (int e = threadIdx.x; e < nc * (PK_PIX / 4); e += blockDim.x) {
        const int c = e / (PK_PIX / 4), pl = (e - c * (PK_PIX / 4)) * 4;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (pl < valid) {
          const int t = p0 + pl, bq = t / HW;
          v = ld_stream(reinterpret_cast<const float4*>(src + (b0 + bq) * src_bs + (long long)(c_lo + c0 + c) * HW + (t - bq * HW)));
        }
        tile[c][pl] = v.x; tile[c][pl + 1] = v.y; tile[c][pl + 2] = v.z; tile[c][pl + 3] = v.w;
      }
    } else {
      for (int e = threadIdx.x; e < nc * PK_PIX; e += blockDim.x) {
        const int c = e / PK_PIX, pl = e - c * PK_PIX;
        float v = 0.0f;
        if (pl < valid) {
          const int t = p0 + pl, bq = t / HW;
          v = __ldg(src + (b0 + bq) * src_bs + (long long)(c_lo + c0 + c) * HW + (t - bq * HW));
        }
        tile[c][pl] = v;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < PK_PIX * G; e += blockDim.x) {
      const int pl = e / G, gq = e - pl * G;
      if (pl < valid) {
        __nv_bfloat162 h[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(tile[8 * gq + 2 * k][pl], tile[8 * gq + 2 * k + 1][pl]);
        *reinterpret_cast<uint4*>(dst + (g0 + pl) * dst_ld + dst_off + c0 + 8 * gq) = *reinterpret_cast<uint4*>(h);
      }
    }
    __syncthreads();
  }
}

// Register-only variant with full-sector accesses on both sides: a warp takes 8 consecutive pixels x 32 channels
// (lane = pixel + 8 * channel-group): every load instruction reads four 32-byte runs (8 pixels of one plane each), every
// store instruction writes eight 64-byte runs (4 groups x 16 bytes of one pixel row).  No shared memory, no barriers.
__global__ void __launch_bounds__(256) pack_nhwc_oct_kernel(const float* __restrict__ src, long long src_bs, int HW, int c_lo,
                                                            int n, __nv_bfloat16* __restrict__ dst, int dst_off, int dst_ld,
                                                            long long npix) {
  pdl_trigger();
  pdl_wait();
  const int G = n >> 3, Q = (G + 3) >> 2;                 // 8-channel groups, quads of groups (one warp each)
  const long long octs = (npix + 7) >> 3, units = octs * Q;
  const int lane = threadIdx.x & 31, pl = lane & 7, g4 = lane >> 3;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long u = warp0; u < units; u += nwarps) {
    const long long o = u / Q;
    const int q = (int)(u - o * Q), gq = 4 * q + g4;
    const long long pix = 8 * o + pl;
    if (gq >= G || pix >= npix) continue;
    const long long b = pix / HW;
    const float* sp = src + b * src_bs + (long long)(c_lo + 8 * gq) * HW + (pix - b * HW);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(sp + (long long)k * HW);
    __nv_bfloat162 h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    *reinterpret_cast<uint4*>(dst + pix * dst_ld + dst_off + 8 * gq) = *reinterpret_cast<uint4*>(h);
  }
}

template <bool kVec>
__global__ void __launch_bounds__(kThreads) copy_channels_kernel(const float* __restrict__ src, int src_C,
                                                                 int src_off, float* __restrict__ dst, int dst_C,
                                                                 int dst_off, int n, int HW, long long total) {
  pdl_trigger();
  pdl_wait();
  const long long per = (long long)n * HW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long e = kVec ? t * 4 : t;
    long long b = e / per, r = e % per;
    const float* s = src + (b * src_C + src_off) * HW + r;
    float* d = dst + (b * dst_C + dst_off) * HW + r;
    if (kVec) st_stream(reinterpret_cast<float4*>(d), ld_stream(reinterpret_cast<const float4*>(s)));
    else *d = *s;
  }
}

// ------------------------------------------------------------------------------------------
// a3  coupling tail  (Flow/glow_modules.py:275-290).  grid = (chunks, B)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cta_atomic_add(float v, float* dst, float* sh) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0f;
    r = warp_sum(r);
    if (lane == 0) atomicAdd(dst, r);
  }
}

template <bool kVec>
__global__ void __launch_bounds__(kThreads) coupling_tail_kernel(const float* __restrict__ nn, float* __restrict__ z,
                                                                 int C, int HW, int clamp_type,
                                                                 const float* __restrict__ cs,
                                                                 const float* __restrict__ csh,
                                                                 float* __restrict__ logdet, int reverse) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int b = blockIdx.y, half = C >> 1;
  const long long per = (long long)half * HW;   // elements of z2 per sample
  const long long units = kVec ? per >> 2 : per;
  const float* nb = nn + (long long)b * C * HW;
  float* zb = z + ((long long)b * C + half) * HW;
  float acc = 0.0f;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < units;
       t += (long long)gridDim.x * blockDim.x) {
    long long e = kVec ? t * 4 : t;
    int j = (int)(e / HW), p = (int)(e % HW);
    float a = 0.0f, sft = 0.0f;
    if (clamp_type == RFK_CLAMP_REALNVP) { a = __ldg(cs + j); sft = __ldg(csh + j); }
    const float* shp = nb + (long long)(2 * j) * HW + p;
    if (kVec) {
      float4 s4 = ld_stream(reinterpret_cast<const float4*>(shp));
      float4 r4 = ld_stream(reinterpret_cast<const float4*>(shp + HW));
      float4 v = *reinterpret_cast<const float4*>(zb + e);
      float sv[4] = {s4.x, s4.y, s4.z, s4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w}, zv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float ls = clamp_ls(rv[k], clamp_type, a, sft);
        acc += ls;
        zv[k] = reverse ? zv[k] * expf(-ls) - sv[k] : (zv[k] + sv[k]) * expf(ls);
      }
      *reinterpret_cast<float4*>(zb + e) = make_float4(zv[0], zv[1], zv[2], zv[3]);
    } else {
      float ls = clamp_ls(shp[HW], clamp_type, a, sft);
      acc += ls;
      float v = zb[e];
      zb[e] = reverse ? v * expf(-ls) - shp[0] : (v + shp[0]) * expf(ls);
    }
  }
  if (logdet) cta_atomic_add(reverse ? -acc : acc, logdet + b, sh);
}

// Coupling tail fed by a tap-split convolution: taps [B, 9*C, H, W] holds, for every filter tap t = 3*ky+kx,
// the 1x1 product W[:, :, ky, kx] * h at each pixel, so the 3x3 'same' convolution output is
//   nn[c](y, x) = sum_t taps[t*C + c](y + ky - 1, x + kx - 1)        (zero outside the image)
// followed by Conv2dZeros' per-channel affine (scale, shift) and the affine-coupling update.
// Every element of `taps` is read exactly once (coalesced, shifted by at most one pixel).  grid = (chunks, B)
__global__ void __launch_bounds__(kThreads) coupling_taps_kernel(const float* __restrict__ taps, float* __restrict__ z,
                                                                int C, int H, int W, const float* __restrict__ scale,
                                                                const float* __restrict__ shift, int clamp_type,
                                                                const float* __restrict__ cs,
                                                                const float* __restrict__ csh,
                                                                float* __restrict__ logdet, int reverse) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int b = blockIdx.y, half = C >> 1, HW = H * W;
  const long long per = (long long)half * HW;
  const float* tb = taps + (long long)b * 9 * C * HW;
  float* zb = z + ((long long)b * C + half) * HW;
  float acc = 0.0f;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < per;
       t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t / HW), p = (int)(t % HW);
    const int y = p / W, x = p - y * W;
    float s_sum = 0.0f, r_sum = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const float* q = tb + ((long long)((3 * ky + kx) * C + 2 * j) * HW) + yy * W + xx;
        s_sum += __ldg(q);
        r_sum += __ldg(q + HW);
      }
    }
    const float sft = fmaf(s_sum, __ldg(scale + 2 * j), __ldg(shift + 2 * j));
    const float raw = fmaf(r_sum, __ldg(scale + 2 * j + 1), __ldg(shift + 2 * j + 1));
    float a = 0.0f, bb = 0.0f;
    if (clamp_type == RFK_CLAMP_REALNVP) { a = __ldg(cs + j); bb = __ldg(csh + j); }
    const float ls = clamp_ls(raw, clamp_type, a, bb);
    acc += ls;
    const float v = zb[t];
    zb[t] = reverse ? v * expf(-ls) - sft : (v + sft) * expf(ls);
  }
  if (logdet) cta_atomic_add(reverse ? -acc : acc, logdet + b, sh);
}

// Same, four consecutive pixels of a row per thread (W % 4 == 0): per tap plane one aligned 128-bit load plus, for the
// horizontally shifted taps, one scalar edge element.
__global__ void __launch_bounds__(kThreads) coupling_taps_v4_kernel(const float* __restrict__ taps, float* __restrict__ z,
                                                                   int C, int H, int W, const float* __restrict__ scale,
                                                                   const float* __restrict__ shift, int clamp_type,
                                                                   const float* __restrict__ cs,
                                                                   const float* __restrict__ csh,
                                                                   float* __restrict__ logdet, int reverse) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int b = blockIdx.y, half = C >> 1, HW = H * W, W4 = W >> 2;
  const long long per4 = (long long)half * H * W4;
  const float* tb = taps + (long long)b * 9 * C * HW;
  float* zb = z + ((long long)b * C + half) * HW;
  float acc = 0.0f;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < per4;
       t += (long long)gridDim.x * blockDim.x) {
    const int x4 = (int)(t % W4);
    const long long r = t / W4;
    const int y = (int)(r % H), j = (int)(r / H);
    const int x = x4 << 2;
    float s_sum[4] = {0, 0, 0, 0}, r_sum[4] = {0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* q = tb + ((long long)((3 * ky + kx) * C + 2 * j) * HW) + yy * W + x;
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {  // shift plane, raw plane
          const float* qq = q + pr * HW;
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(qq));
          float v0, v1, v2, v3;
          if (kx == 1) { v0 = c4.x; v1 = c4.y; v2 = c4.z; v3 = c4.w; }
          else if (kx == 0) { v0 = x > 0 ? __ldg(qq - 1) : 0.0f; v1 = c4.x; v2 = c4.y; v3 = c4.z; }
          else { v0 = c4.y; v1 = c4.z; v2 = c4.w; v3 = x + 4 < W ? __ldg(qq + 4) : 0.0f; }
          float* dst = pr ? r_sum : s_sum;
          dst[0] += v0; dst[1] += v1; dst[2] += v2; dst[3] += v3;
        }
      }
    }
    const float sc_s = __ldg(scale + 2 * j), sh_s = __ldg(shift + 2 * j);
    const float sc_r = __ldg(scale + 2 * j + 1), sh_r = __ldg(shift + 2 * j + 1);
    float a = 0.0f, bb = 0.0f;
    if (clamp_type == RFK_CLAMP_REALNVP) { a = __ldg(cs + j); bb = __ldg(csh + j); }
    float4* zp = reinterpret_cast<float4*>(zb + ((long long)j * H + y) * W + x);
    const float4 zv = *zp;
    float zi[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sft = fmaf(s_sum[k], sc_s, sh_s);
      const float ls = clamp_ls(fmaf(r_sum[k], sc_r, sh_r), clamp_type, a, bb);
      acc += ls;
      zi[k] = reverse ? zi[k] * expf(-ls) - sft : (zi[k] + sft) * expf(ls);
    }
    *zp = make_float4(zi[0], zi[1], zi[2], zi[3]);
  }
  if (logdet) cta_atomic_add(reverse ? -acc : acc, logdet + b, sh);
}

// ------------------------------------------------------------------------------------------
// a5/a7  Gaussian log-density / sampling.  grid = (chunks, B)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gauss_logp_kernel(const float* __restrict__ z, int z_C, int z_off,
                                                              const float* __restrict__ params, int n, int HW,
                                                              int pairing, int std_kind,
                                                              float* __restrict__ logdet) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int b = blockIdx.y;
  const long long per = (long long)n * HW;
  const float* zb = z + ((long long)b * z_C + z_off) * HW;
  const float* pb = params ? params + (long long)b * 2 * n * HW : nullptr;
  float acc = 0.0f;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < per;
       t += (long long)gridDim.x * blockDim.x) {
    int j = (int)(t / HW), p = (int)(t % HW);
    float mean = 0.0f, raw = 0.0f;
    if (pb) {
      int cm = pairing == RFK_PAIR_CROSS ? 2 * j : j;
      int cr = pairing == RFK_PAIR_CROSS ? 2 * j + 1 : n + j;
      mean = pb[(long long)cm * HW + p];
      raw = pb[(long long)cr * HW + p];
    }
    float sd = std_from_raw(raw, std_kind);
    // exp-parameterised std: log(std) is raw itself (avoids exp/log round trip)
    float lsd = std_kind == RFK_STD_EXP ? raw : logf(sd);
    float d = zb[t] - mean;
    acc += -(d * d) / (2.0f * sd * sd) - lsd - 0.91893853320467274178f;
  }
  cta_atomic_add(acc, logdet + b, sh);
}

// 128-bit variant (HW % 4 == 0, 16-byte aligned planes): four consecutive positions of one channel per thread, 32-bit
// index arithmetic, no division in the density itself.
__global__ void __launch_bounds__(kThreads) gauss_logp_v4_kernel(const float* __restrict__ z, int z_C, int z_off,
                                                                 const float* __restrict__ params, int n, int HW,
                                                                 int pairing, int std_kind, float* __restrict__ logdet) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int b = blockIdx.y;
  const int per4 = (int)(((long long)n * HW) >> 2), HW4 = HW >> 2;
  const float4* zb = reinterpret_cast<const float4*>(z + ((long long)b * z_C + z_off) * HW);
  const float4* pb = params ? reinterpret_cast<const float4*>(params + (long long)b * 2 * n * HW) : nullptr;
  float acc = 0.0f;
  if (!pb) {   // N(0, 1): read-only stream, four independent 128-bit loads in flight per thread
    const int stride = gridDim.x * blockDim.x;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    for (; t + 3 * stride < per4; t += 4 * stride) {
      const float4 v0 = ld_stream(zb + t), v1 = ld_stream(zb + t + stride), v2 = ld_stream(zb + t + 2 * stride),
                   v3 = ld_stream(zb + t + 3 * stride);
      a0 += v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w;
      a1 += v1.x * v1.x + v1.y * v1.y + v1.z * v1.z + v1.w * v1.w;
      a2 += v2.x * v2.x + v2.y * v2.y + v2.z * v2.z + v2.w * v2.w;
      a3 += v3.x * v3.x + v3.y * v3.y + v3.z * v3.z + v3.w * v3.w;
      acc -= 16.0f * 0.91893853320467274178f;
    }
    for (; t < per4; t += stride) {
      const float4 v0 = ld_stream(zb + t);
      a0 += v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w;
      acc -= 4.0f * 0.91893853320467274178f;
    }
    acc -= 0.5f * ((a0 + a1) + (a2 + a3));
    cta_atomic_add(acc, logdet + b, sh);
    return;
  }
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per4; t += gridDim.x * blockDim.x) {
    const float4 zv = ld_stream(zb + t);
    const int j = t / HW4, p4 = t - j * HW4;
    const int cm = pairing == RFK_PAIR_CROSS ? 2 * j : j, cr = pairing == RFK_PAIR_CROSS ? 2 * j + 1 : n + j;
    const float4 mv = ld_stream(pb + (long long)cm * HW4 + p4), rv = ld_stream(pb + (long long)cr * HW4 + p4);
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w}, rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = zz[k] - mm[k];
      if (std_kind == RFK_STD_EXP) {   // log(std) = raw, 1/std^2 = exp(-2 raw)
        acc += -0.5f * d * d * expf(-2.0f * rr[k]) - rr[k] - 0.91893853320467274178f;
      } else {
        const float sd = std_from_raw(rr[k], std_kind), inv = 1.0f / sd;
        acc += -0.5f * d * d * inv * inv - logf(sd) - 0.91893853320467274178f;
      }
    }
  }
  cta_atomic_add(acc, logdet + b, sh);
}

__global__ void __launch_bounds__(kThreads) gauss_sample_kernel(const float* __restrict__ eps,
                                                                const float* __restrict__ params, int n, int HW,
                                                                int pairing, int std_kind, float temperature,
                                                                float* __restrict__ out, int out_C, int out_off) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const long long per = (long long)n * HW;
  const float* pb = params ? params + (long long)b * 2 * n * HW : nullptr;
  const float* eb = eps + (long long)b * per;
  float* ob = out + ((long long)b * out_C + out_off) * HW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < per;
       t += (long long)gridDim.x * blockDim.x) {
    int j = (int)(t / HW), p = (int)(t % HW);
    float mean = 0.0f, raw = 0.0f;
    if (pb) {
      int cm = pairing == RFK_PAIR_CROSS ? 2 * j : j;
      int cr = pairing == RFK_PAIR_CROSS ? 2 * j + 1 : n + j;
      mean = pb[(long long)cm * HW + p];
      raw = pb[(long long)cr * HW + p];
    }
    ob[t] = mean + std_from_raw(raw, std_kind) * temperature * eb[t];
  }
}

// ------------------------------------------------------------------------------------------
// a9  ConvLSTM cell update  (Utils/modules.py:369-377), gate order i,f,o,g; o peeps at c_next
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void lstm_point(float ci, float cf, float co, float cg, float c, float wi, float wf,
                                           float wo, float& h, float& cn) {
  float i = sigmoidf_(ci + wi * c);
  float f = sigmoidf_(cf + wf * c);
  float g = tanhf(cg);
  cn = f * c + i * g;
  float o = sigmoidf_(co + wo * cn);
  h = o * tanhf(cn);
}

template <bool kVec>
__global__ void __launch_bounds__(kThreads) lstm_pointwise_kernel(const float* __restrict__ cc,
                                                                  const float* __restrict__ c_prev,
                                                                  const float* __restrict__ peep,
                                                                  float* __restrict__ h_out,
                                                                  float* __restrict__ c_next, int Hc, int HW,
                                                                  long long total) {
  pdl_trigger();
  pdl_wait();
  const long long per = (long long)Hc * HW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long e = kVec ? t * 4 : t;
    long long b = e / per, r = e % per;  // r = ch*HW + p
    const float* g0 = cc + b * 4 * per + r;
    if (kVec) {
      float4 vi = ld_stream(reinterpret_cast<const float4*>(g0));
      float4 vf = ld_stream(reinterpret_cast<const float4*>(g0 + per));
      float4 vo = ld_stream(reinterpret_cast<const float4*>(g0 + 2 * per));
      float4 vg = ld_stream(reinterpret_cast<const float4*>(g0 + 3 * per));
      float4 vc = ld_stream(reinterpret_cast<const float4*>(c_prev + e));
      float4 wi = make_float4(0, 0, 0, 0), wf = wi, wo = wi;
      if (peep) {
        wi = *reinterpret_cast<const float4*>(peep + r);
        wf = *reinterpret_cast<const float4*>(peep + per + r);
        wo = *reinterpret_cast<const float4*>(peep + 2 * per + r);
      }
      float4 h, cn;
      lstm_point(vi.x, vf.x, vo.x, vg.x, vc.x, wi.x, wf.x, wo.x, h.x, cn.x);
      lstm_point(vi.y, vf.y, vo.y, vg.y, vc.y, wi.y, wf.y, wo.y, h.y, cn.y);
      lstm_point(vi.z, vf.z, vo.z, vg.z, vc.z, wi.z, wf.z, wo.z, h.z, cn.z);
      lstm_point(vi.w, vf.w, vo.w, vg.w, vc.w, wi.w, wf.w, wo.w, h.w, cn.w);
      st_stream(reinterpret_cast<float4*>(h_out + e), h);
      st_stream(reinterpret_cast<float4*>(c_next + e), cn);
    } else {
      float wi = 0, wf = 0, wo = 0;
      if (peep) { wi = peep[r]; wf = peep[per + r]; wo = peep[2 * per + r]; }
      float h, cn;
      lstm_point(g0[0], g0[per], g0[2 * per], g0[3 * per], c_prev[e], wi, wf, wo, h, cn);
      h_out[e] = h;
      c_next[e] = cn;
    }
  }
}

// ConvLSTM cell update from a pixel-major fp32 gate buffer (the split-K GEMM's workspace): cc[(b*HW+p)*ld + g*Hc + ch],
// gate order i,f,o,g, + bias; optionally clears the workspace behind itself so the next step starts from zero, and
// hands h to the next step as bf16 NHWC.
__global__ void __launch_bounds__(kThreads) lstm_pointwise_ws_kernel(float* __restrict__ cc, int cc_ld,
                                                                     const float* __restrict__ bias,
                                                                     const float* __restrict__ c_prev, long long c_prev_bs,
                                                                     const float* __restrict__ peep,
                                                                     float* __restrict__ h_out, long long h_bs,
                                                                     float* __restrict__ c_next, long long c_next_bs,
                                                                     __nv_bfloat16* __restrict__ h_nhwc, int h_off, int h_ld,
                                                                     int Hc, int HW, int zero_cc, long long total) {
  pdl_trigger();
  pdl_wait();
  const long long per = (long long)Hc * HW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / per, r = t % per;
    const int ch = (int)(r / HW), p = (int)(r % HW);
    float* g0 = cc + (b * HW + p) * cc_ld + ch;
    float ci = g0[0], cf = g0[Hc], co = g0[2 * Hc], cg = g0[3 * Hc];
    if (zero_cc) { g0[0] = 0.0f; g0[Hc] = 0.0f; g0[2 * Hc] = 0.0f; g0[3 * Hc] = 0.0f; }
    if (bias) { ci += bias[ch]; cf += bias[Hc + ch]; co += bias[2 * Hc + ch]; cg += bias[3 * Hc + ch]; }
    const float c = c_prev ? c_prev[b * c_prev_bs + r] : 0.0f;
    float wi = 0, wf = 0, wo = 0;
    if (peep) { wi = peep[r]; wf = peep[per + r]; wo = peep[2 * per + r]; }
    float h, cn;
    lstm_point(ci, cf, co, cg, c, wi, wf, wo, h, cn);
    h_out[b * h_bs + r] = h;
    c_next[b * c_next_bs + r] = cn;
    if (h_nhwc) h_nhwc[(b * HW + p) * h_ld + h_off + ch] = __float2bfloat16(h);
  }
}

// BatchNormFlow (Flow/glow_modules.py:56-104): per-position affine y[b,i] = x[b,i]*a[i] + c[i], i over (C,H,W)
__global__ void __launch_bounds__(kThreads) affine_pos_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              const float* __restrict__ a, const float* __restrict__ c,
                                                              long long n4, long long total4) {
  pdl_trigger();
  pdl_wait();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total4;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t % n4;
    const float4 v = ld_stream(reinterpret_cast<const float4*>(x) + t);
    const float4 av = __ldg(reinterpret_cast<const float4*>(a) + i), cv = __ldg(reinterpret_cast<const float4*>(c) + i);
    st_stream(reinterpret_cast<float4*>(y) + t,
              make_float4(fmaf(v.x, av.x, cv.x), fmaf(v.y, av.y, cv.y), fmaf(v.z, av.z, cv.z), fmaf(v.w, av.w, cv.w)));
  }
}
__global__ void __launch_bounds__(kThreads) affine_pos_scalar_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                     const float* __restrict__ a,
                                                                     const float* __restrict__ c, long long n,
                                                                     long long total) {
  pdl_trigger();
  pdl_wait();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x)
    y[t] = fmaf(x[t], a[t % n], c[t % n]);
}

// per-position mean and biased variance (+eps) over the batch dimension: one thread per position, coalesced over i
__global__ void __launch_bounds__(kThreads) batch_stats_pos_kernel(const float* __restrict__ x, float* __restrict__ mean,
                                                                   float* __restrict__ var, int B, long long n,
                                                                   float eps) {
  pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int b = 0; b < B; ++b) s += x[b * n + i];
    const float m = s / (float)B;
    float q = 0.0f;
    for (int b = 0; b < B; ++b) { const float d = x[b * n + i] - m; q = fmaf(d, d, q); }
    mean[i] = m;
    var[i] = q / (float)B + eps;
  }
}

__global__ void add_scalar_kernel(float* __restrict__ logdet, const float* __restrict__ addend, float alpha, int B) {
  pdl_trigger();
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) logdet[i] += alpha * (*addend);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Medium channel counts (16 < C <= 64, HW % 4 == 0): the weight matrix is staged ONCE per CTA (transposed, so the four
// outputs of a thread come from one 128-bit load) and the CTA walks pixel tiles; a thread owns 4 outputs x 8 pixels
// (32 FMAs per three 128-bit shared-memory loads).  The generic kernel above re-stages W for every 32 pixels and does
// nine shared-memory loads per eight FMAs, which makes it fp32-issue-bound long before HBM (C = 48: 0.6 TB/s).
__global__ void __launch_bounds__(256) mix1x1_mid_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                         const float* __restrict__ Wm, const float* __restrict__ bvec, int C,
                                                         int HW, long long npix, __nv_bfloat16* __restrict__ side, int side_n,
                                                         int side_off, int side_ld, float* __restrict__ logdet,
                                                         const float* __restrict__ addend, float alpha, int B, int PO) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  const int C4 = (C + 3) & ~3, OQ = C4 >> 2, TP = PO * 8;
  float* wT = smem;                 // [C][C4]: wT[i][o] = Wm[o][i]
  float* bs = wT + C * C4;          // [C4]
  float* xs = bs + C4;              // [C][TP]
  for (int e = threadIdx.x; e < C * C4; e += blockDim.x) {
    const int i = e / C4, o = e - i * C4;
    wT[e] = o < C ? Wm[o * C + i] : 0.0f;
  }
  for (int e = threadIdx.x; e < C4; e += blockDim.x) bs[e] = (bvec && e < C) ? bvec[e] : 0.0f;
  if (logdet && blockIdx.x == 0) {
    const float add = alpha * (*addend);
    for (int i = threadIdx.x; i < B; i += blockDim.x) atomicAdd(logdet + i, add);
  }
  const int po = threadIdx.x % PO, oq = threadIdx.x / PO;   // pixel group, output quad (oq >= OQ: helper thread, loads only)
  const int TPh = TP >> 1;   // a thread owns pixels [4po, 4po+4) of BOTH tile halves: conflict-free 128-bit shared loads
  for (long long g0 = (long long)blockIdx.x * TP; g0 < npix; g0 += (long long)gridDim.x * TP) {
    const long long b0 = g0 / HW;               // one 64-bit division per tile, 32-bit arithmetic below
    const int p0 = (int)(g0 - b0 * HW);
    const int valid = (int)min((long long)TP, npix - g0);
    __syncthreads();   // weights staged / previous tile consumed
    {
      // all of a thread's loads are issued before the first shared-memory store (a store right behind its load would
      // serialise the HBM latencies); C * TP / 4 <= 8 * 256 + a few, so at most nine rounds of 256 threads
      const int total = C * (TP / 4);
      float4 v[9];
      int dsto[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int e = threadIdx.x + k * 256;
        v[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        dsto[k] = -1;
        if (e < total) {
          const int i = e / (TP / 4), q4 = (e - i * (TP / 4)) * 4;
          dsto[k] = i * TP + q4;
          if (q4 < valid) {
            const int t = p0 + q4, bq = t / HW;
            v[k] = ld_stream(reinterpret_cast<const float4*>(x + ((b0 + bq) * C + i) * HW + (t - bq * HW)));
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 9; ++k)
        if (dsto[k] >= 0) *reinterpret_cast<float4*>(xs + dsto[k]) = v[k];
    }
    __syncthreads();
    if (oq < OQ && 4 * po < valid) {
      const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * oq);
      float acc[4][8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc[0][k] = b4.x; acc[1][k] = b4.y; acc[2][k] = b4.z; acc[3][k] = b4.w; }
#pragma unroll 2
      for (int i = 0; i < C; ++i) {
        const float4 w4 = *reinterpret_cast<const float4*>(wT + i * C4 + 4 * oq);
        const float4 xa = *reinterpret_cast<const float4*>(xs + i * TP + 4 * po);
        const float4 xb = *reinterpret_cast<const float4*>(xs + i * TP + TPh + 4 * po);
        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc[0][k] = fmaf(w4.x, xv[k], acc[0][k]);
          acc[1][k] = fmaf(w4.y, xv[k], acc[1][k]);
          acc[2][k] = fmaf(w4.z, xv[k], acc[2][k]);
          acc[3][k] = fmaf(w4.w, xv[k], acc[3][k]);
        }
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int q4 = hh * TPh + 4 * po;
        if (q4 >= valid) continue;
        const int t = p0 + q4, bq = t / HW, p = t - bq * HW;
        const long long b = b0 + bq, pix = g0 + q4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int o = 4 * oq + j;
          if (o < C) {
            st_stream(reinterpret_cast<float4*>(y + (b * C + o) * HW + p),
                      make_float4(acc[j][4 * hh], acc[j][4 * hh + 1], acc[j][4 * hh + 2], acc[j][4 * hh + 3]));
            if (side && o < side_n) {
#pragma unroll
              for (int k = 0; k < 4; ++k) side[(pix + k) * side_ld + side_off + o] = __float2bfloat16(acc[j][4 * hh + k]);
            }
          }
        }
      }
    }
  }
}

// Small channel counts (C <= 16): no shared-memory staging, one thread per 4 consecutive pixels, x and y in registers.
template <int CT>
__global__ void __launch_bounds__(kThreads) mix1x1_small_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                const float* __restrict__ Wm,
                                                                const float* __restrict__ bvec, int C, int HW,
                                                                long long nquad, __nv_bfloat16* __restrict__ side,
                                                                int side_n, int side_off, int side_ld,
                                                                float* __restrict__ logdet,
                                                                const float* __restrict__ addend, float alpha, int B) {
  pdl_trigger();
  pdl_wait();
  __shared__ float ws[CT * CT + CT];
  for (int i = threadIdx.x; i < CT * CT + CT; i += blockDim.x) {
    float v = 0.0f;
    if (i < CT * CT) { int o = i / CT, c = i % CT; if (o < C && c < C) v = Wm[o * C + c]; }
    else { int o = i - CT * CT; if (bvec && o < C) v = bvec[o]; }
    ws[i] = v;
  }
  if (logdet && blockIdx.x == 0) {
    const float add = alpha * (*addend);
    for (int i = threadIdx.x; i < B; i += blockDim.x) logdet[i] += add;
  }
  __syncthreads();
  const int HW4 = HW >> 2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nquad;
       t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / HW4;
    const int p = (int)(t % HW4) << 2;
    const float* xp = x + b * C * HW + p;
    float4 xi[CT];
#pragma unroll
    for (int i = 0; i < CT; ++i) xi[i] = i < C ? ld_stream(reinterpret_cast<const float4*>(xp + (long long)i * HW)) : make_float4(0, 0, 0, 0);
    float* yp = y + b * C * HW + p;
    const long long pix = b * HW + p;
    __align__(16) __nv_bfloat16 sv[4][CT / 2];   // side channels (the first half of the outputs) of the 4 pixels
#pragma unroll
    for (int o = 0; o < CT; ++o) {
      if (o < C) {
        const float bo = ws[CT * CT + o];
        float4 a = make_float4(bo, bo, bo, bo);
#pragma unroll
        for (int i = 0; i < CT; ++i) {
          const float w = ws[o * CT + i];
          a.x = fmaf(w, xi[i].x, a.x); a.y = fmaf(w, xi[i].y, a.y); a.z = fmaf(w, xi[i].z, a.z); a.w = fmaf(w, xi[i].w, a.w);
        }
        st_stream(reinterpret_cast<float4*>(yp + (long long)o * HW), a);
        if (o < CT / 2) {
          sv[0][o] = __float2bfloat16(a.x); sv[1][o] = __float2bfloat16(a.y);
          sv[2][o] = __float2bfloat16(a.z); sv[3][o] = __float2bfloat16(a.w);
        }
      }
    }
    if (side) {
      // one packed store per pixel when the side window is exactly the first C/2 channels and suitably aligned
      const bool packed = side_n == CT / 2 && ((side_off * 2) % (CT) == 0) && ((side_ld * 2) % (CT) == 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __nv_bfloat16* sp = side + (pix + k) * side_ld + side_off;
        if (packed) {
          if (CT == 4) *reinterpret_cast<uint32_t*>(sp) = *reinterpret_cast<const uint32_t*>(sv[k]);
          else if (CT == 8) *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(sv[k]);
          else *reinterpret_cast<uint4*>(sp) = *reinterpret_cast<const uint4*>(sv[k]);
        } else {
#pragma unroll
          for (int o = 0; o < CT / 2; ++o)
            if (o < side_n) sp[o] = sv[k][o];
        }
      }
    }
  }
}
}  // namespace rfk

using namespace rfk;

extern "C" int rfk_squeeze2d(const float* x, float* y, int B, int C, int H, int W, int undo, void* stream) {
  RFK_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0, "rfk_squeeze2d: null pointer or empty shape");
  cudaStream_t st = (cudaStream_t)stream;
  long long total = (long long)B * C * H * W;
  if (!undo) {
    RFK_REQUIRE(H % 2 == 0 && W % 2 == 0, "rfk_squeeze2d: H=%d, W=%d must be even", H, W);
    if (W % 8 == 0 && aligned16(x) && aligned16(y)) {
      long long t8 = total / 8;
      RFK_LAUNCH(squeeze_fwd_v8, stream_grid(t8, kThreads, 8), kThreads, 0, st, x, y, C, H, W, t8);
    } else {
      RFK_LAUNCH(squeeze_scalar, stream_grid(total, kThreads, 8), kThreads, 0, st, x, y, C, H, W, 0, total);
    }
  } else {
    RFK_REQUIRE(C % 4 == 0, "rfk_squeeze2d(undo): C=%d must be a multiple of 4", C);
    int Co = C / 4, Hout = 2 * H, Wout = 2 * W;
    if (Wout % 8 == 0 && aligned16(x) && aligned16(y)) {
      long long t8 = total / 8;
      RFK_LAUNCH(squeeze_undo_v8, stream_grid(t8, kThreads, 8), kThreads, 0, st, x, y, Co, Hout, Wout, t8);
    } else {
      RFK_LAUNCH(squeeze_scalar, stream_grid(total, kThreads, 8), kThreads, 0, st, x, y, Co, Hout, Wout, 1, total);
    }
  }
  return check_launch("rfk_squeeze2d");
}

extern "C" int rfk_actnorm(const float* x, float* y, const float* bias, const float* logs, int B, int C, int HW,
                           int reverse, void* stream) {
  RFK_REQUIRE(x && y && bias && logs && B > 0 && C > 0 && HW > 0, "rfk_actnorm: null pointer or empty shape");
  cudaStream_t st = (cudaStream_t)stream;
  long long total = (long long)B * C * HW;
  if (HW % 4 == 0 && aligned16(x) && aligned16(y)) {
    long long t4 = total / 4;
    RFK_LAUNCH((actnorm_kernel<true>), stream_grid(t4, kThreads, 8), kThreads, 0, st, x, y, bias, logs, C, HW, reverse, t4);
  } else {
    RFK_LAUNCH((actnorm_kernel<false>), stream_grid(total, kThreads, 8), kThreads, 0, st, x, y, bias, logs, C, HW, reverse,
                                                                                 total);
  }
  return check_launch("rfk_actnorm");
}

extern "C" int rfk_actnorm_init(const float* x, float* bias, float* logs, float* mean_out, float* std_out, int B,
                                int C, int HW, void* stream) {
  RFK_REQUIRE(x && B > 0 && C > 0 && HW > 0, "rfk_actnorm_init: null pointer or empty shape");
  RFK_REQUIRE((long long)B * HW > 1, "rfk_actnorm_init: unbiased std needs more than one element per channel");
  RFK_LAUNCH(actnorm_init_kernel, C, 512, 0, (cudaStream_t)stream, x, bias, logs, mean_out, std_out, B, C, HW);
  return check_launch("rfk_actnorm_init");
}

static int launch_mix_generic(const char* who, const float* x, float* y, const float* Wm, const float* bvec, int B, int C,
                              int HW, void* side, int side_n, int side_off, int side_ld, float* logdet,
                              const float* addend, float alpha, const rfk::CouplingSrc& cp, void* stream) {
  const int G = (C + 7) / 8;
  RFK_REQUIRE(G <= 32, "%s: C=%d is too large (max 256)", who, C);
  int PT = (256 / G) / 32 * 32;
  if (PT < 32) PT = 32;
  size_t smem = ((((size_t)C + (size_t)C * PT) + 3) & ~(size_t)3) * sizeof(float);
  const int w_smem = smem + (size_t)C * C * sizeof(float) <= 160 * 1024;
  if (w_smem) smem += (size_t)C * C * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(mix1x1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return RFK_ECUDA; }
    configured = smem;
  }
  long long npix = (long long)B * HW;
  RFK_LAUNCH(mix1x1_kernel, ceil_div(npix, PT), dim3(PT, G), smem, (cudaStream_t)stream,
             x, y, Wm, bvec, C, HW, npix, (__nv_bfloat16*)side, side ? side_n : 0, side_off, side_ld, w_smem, logdet, addend,
             alpha, B, cp);
  return check_launch(who);
}

extern "C" int rfk_mix1x1(const float* x, float* y, const float* Wm, const float* bvec, int B, int C, int HW,
                          void* side, int side_n, int side_off, int side_ld, float* logdet, const float* addend,
                          float alpha, void* stream) {
  RFK_REQUIRE(x && y && Wm && B > 0 && C > 0 && HW > 0, "rfk_mix1x1: null pointer or empty shape");
  RFK_REQUIRE(x != y, "rfk_mix1x1: in-place is not supported");
  RFK_REQUIRE(!logdet || addend, "rfk_mix1x1: logdet given without an addend");
  if (side) RFK_REQUIRE(side_n >= 0 && side_n <= C && side_off >= 0 && side_off + side_n <= side_ld,
                        "rfk_mix1x1: bad side-output window");
  if (C <= 16 && HW % 4 == 0 && aligned16(x) && aligned16(y) && (!side || side_n <= (C <= 4 ? 2 : C <= 8 ? 4 : 8))) {
    const long long nquad = (long long)B * HW / 4;
    const int grid = stream_grid(nquad, kThreads, 8);
    if (C > 8)
      RFK_LAUNCH((mix1x1_small_kernel<16>), grid, kThreads, 0, (cudaStream_t)stream, x, y, Wm, bvec, C, HW, nquad, (__nv_bfloat16*)side,
                                                                           side ? side_n : 0, side_off, side_ld, logdet, addend, alpha, B);
    else if (C <= 4)
      RFK_LAUNCH((mix1x1_small_kernel<4>), grid, kThreads, 0, (cudaStream_t)stream, x, y, Wm, bvec, C, HW, nquad, (__nv_bfloat16*)side,
                                                                          side ? side_n : 0, side_off, side_ld, logdet, addend, alpha, B);
    else
      RFK_LAUNCH((mix1x1_small_kernel<8>), grid, kThreads, 0, (cudaStream_t)stream, x, y, Wm, bvec, C, HW, nquad, (__nv_bfloat16*)side,
                                                                          side ? side_n : 0, side_off, side_ld, logdet, addend, alpha, B);
    return check_launch("rfk_mix1x1");
  }
  if (C > 16 && C <= 64 && HW % 4 == 0 && aligned16(x) && aligned16(y) && (long long)B * HW >= 32768) {   // small launches: the generic kernel's many small CTAs have the lower latency
    const int C4 = (C + 3) & ~3, OQ = C4 / 4, PO = 256 / OQ, TP = PO * 8;
    const size_t smem = ((size_t)C * C4 + C4 + (size_t)C * TP) * sizeof(float);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      cudaError_t e = cudaFuncSetAttribute(mix1x1_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_error("rfk_mix1x1: %s", cudaGetErrorString(e)); return RFK_ECUDA; }
      configured = smem;
    }
    const long long npix = (long long)B * HW;
    const int grid = (int)std::min<long long>((npix + TP - 1) / TP, (long long)sm_count() * 4);
    RFK_LAUNCH(mix1x1_mid_kernel, grid, 256, smem, (cudaStream_t)stream, x, y, Wm, bvec, C, HW, npix, (__nv_bfloat16*)side,
               side ? side_n : 0, side_off, side_ld, logdet, addend, alpha, B, PO);
    return check_launch("rfk_mix1x1");
  }
  rfk::CouplingSrc none;
  none.taps = nullptr; none.scale = nullptr; none.shift = nullptr; none.clamp_type = 0; none.cs = nullptr; none.csh = nullptr;
  none.logdet = nullptr; none.reverse = 0; none.H = 1; none.W = HW;
  return launch_mix_generic("rfk_mix1x1", x, y, Wm, bvec, B, C, HW, side, side_n, side_off, side_ld, logdet, addend, alpha,
                            none, stream);
}

extern "C" int rfk_coupling_taps_mix(const float* taps, const float* z, float* y, int B, int C, int H, int W,
                                     const float* scale, const float* shift, int clamp_type, const float* clamp_scale,
                                     const float* clamp_shift, float* cpl_logdet, int reverse, const float* Wm,
                                     const float* bvec, void* side, int side_n, int side_off, int side_ld,
                                     float* logdet, const float* addend, float alpha, void* stream) {
  RFK_REQUIRE(taps && z && y && Wm && scale && shift && B > 0 && C > 0 && C % 2 == 0 && H > 0 && W > 0,
              "rfk_coupling_taps_mix: null pointer or bad shape (C must be even)");
  RFK_REQUIRE(z != y, "rfk_coupling_taps_mix: in-place is not supported");
  RFK_REQUIRE(clamp_type >= 0 && clamp_type <= 3, "rfk_coupling_taps_mix: unknown clamp_type %d", clamp_type);
  RFK_REQUIRE(clamp_type != RFK_CLAMP_REALNVP || (clamp_scale && clamp_shift),
              "rfk_coupling_taps_mix: realnvp clamp needs scale and scale_shift");
  RFK_REQUIRE(!logdet || addend, "rfk_coupling_taps_mix: logdet given without an addend");
  if (side) RFK_REQUIRE(side_n >= 0 && side_n <= C && side_off >= 0 && side_off + side_n <= side_ld,
                        "rfk_coupling_taps_mix: bad side-output window");
  rfk::CouplingSrc cp;
  cp.taps = taps; cp.scale = scale; cp.shift = shift; cp.clamp_type = clamp_type; cp.cs = clamp_scale; cp.csh = clamp_shift;
  cp.logdet = cpl_logdet; cp.reverse = reverse; cp.H = H; cp.W = W;
  return launch_mix_generic("rfk_coupling_taps_mix", z, y, Wm, bvec, B, C, H * W, side, side_n, side_off, side_ld, logdet,
                            addend, alpha, cp, stream);
}

extern "C" int rfk_pack_nhwc_bf16(const float* src, long long src_bstride, int B, int Csrc, int HW, int c_lo, int n,
                                  void* dst, int dst_off, int dst_ld, void* stream) {
  RFK_REQUIRE(src && dst && B > 0 && HW > 0, "rfk_pack_nhwc_bf16: null pointer or empty shape");
  RFK_REQUIRE(c_lo >= 0 && n >= 0 && c_lo + n <= Csrc && dst_off >= 0 && dst_off + n <= dst_ld,
              "rfk_pack_nhwc_bf16: bad channel window");
  if (n == 0) return RFK_OK;
  long long npix = (long long)B * HW;
  int vec_ok = (dst_ld % 8 == 0) && (dst_off % 8 == 0) && aligned16(dst);
  static const int pack_mode = [] { const char* e = getenv("RFK_PACK_MODE"); return e ? atoi(e) : 2; }();   // 0 flat, 1 tiled, 2 octets
  if (pack_mode == 2 && vec_ok && n % 8 == 0 && npix >= 4096) {
    const int Q = (n / 8 + 3) / 4;
    const long long units = ((npix + 7) / 8) * Q;
    const int grid = (int)std::min<long long>((units + 7) / 8, (long long)sm_count() * 16);
    RFK_LAUNCH(pack_nhwc_oct_kernel, grid, 256, 0, (cudaStream_t)stream, src,
               src_bstride > 0 ? src_bstride : (long long)Csrc * HW, HW, c_lo, n, (__nv_bfloat16*)dst, dst_off, dst_ld, npix);
    return check_launch("rfk_pack_nhwc_bf16");
  }
  if (pack_mode == 1 && vec_ok && n % 8 == 0 && n >= 16 && npix >= 4096) {
    const int chunks = (n + 63) / 64;
    const long long tiles = (npix + PK_PIX - 1) / PK_PIX;
    const int gx = (int)std::min<long long>(tiles, std::max(1, sm_count() * 8 / chunks));
    const long long bs = src_bstride > 0 ? src_bstride : (long long)Csrc * HW;
    const int vec4 = (HW % 4 == 0) && (bs % 4 == 0) && aligned16(src);
    RFK_LAUNCH(pack_nhwc_tiled_kernel, dim3(gx, chunks), 256, 0, (cudaStream_t)stream, src, bs, HW, c_lo, n,
               (__nv_bfloat16*)dst, dst_off, dst_ld, npix, vec4);
    return check_launch("rfk_pack_nhwc_bf16");
  }
  RFK_LAUNCH(pack_nhwc_kernel, stream_grid(npix * ((n + 7) / 8), kThreads, 8), kThreads, 0, (cudaStream_t)stream, 
      src, src_bstride > 0 ? src_bstride : (long long)Csrc * HW, HW, c_lo, n, (__nv_bfloat16*)dst, dst_off, dst_ld,
      npix, vec_ok, 0);
  return check_launch("rfk_pack_nhwc_bf16");
}

extern "C" int rfk_pack_nhwc_bf16_lo(const float* src, long long src_bstride, int B, int Csrc, int HW, int c_lo, int n,
                                     void* dst, int dst_off, int dst_ld, void* stream) {
  RFK_REQUIRE(src && dst && B > 0 && HW > 0, "rfk_pack_nhwc_bf16_lo: null pointer or empty shape");
  RFK_REQUIRE(c_lo >= 0 && n >= 0 && c_lo + n <= Csrc && dst_off >= 0 && dst_off + n <= dst_ld,
              "rfk_pack_nhwc_bf16_lo: bad channel window");
  if (n == 0) return RFK_OK;
  long long npix = (long long)B * HW;
  RFK_LAUNCH(pack_nhwc_kernel, stream_grid(npix * ((n + 7) / 8), kThreads, 8), kThreads, 0, (cudaStream_t)stream,
      src, src_bstride > 0 ? src_bstride : (long long)Csrc * HW, HW, c_lo, n, (__nv_bfloat16*)dst, dst_off, dst_ld,
      npix, 0, 1);
  return check_launch("rfk_pack_nhwc_bf16_lo");
}

extern "C" int rfk_copy_channels(const float* src, int src_C, int src_off, float* dst, int dst_C, int dst_off,
                                 int n, int B, int HW, void* stream) {
  RFK_REQUIRE(src && dst && B > 0 && HW > 0 && n >= 0, "rfk_copy_channels: null pointer or empty shape");
  RFK_REQUIRE(src_off >= 0 && src_off + n <= src_C && dst_off >= 0 && dst_off + n <= dst_C,
              "rfk_copy_channels: bad channel window");
  if (n == 0) return RFK_OK;
  long long total = (long long)B * n * HW;
  cudaStream_t st = (cudaStream_t)stream;
  if (HW % 4 == 0 && aligned16(src) && aligned16(dst)) {
    RFK_LAUNCH((copy_channels_kernel<true>), stream_grid(total / 4, kThreads, 8), kThreads, 0, st, 
        src, src_C, src_off, dst, dst_C, dst_off, n, HW, total / 4);
  } else {
    RFK_LAUNCH((copy_channels_kernel<false>), stream_grid(total, kThreads, 8), kThreads, 0, st, 
        src, src_C, src_off, dst, dst_C, dst_off, n, HW, total);
  }
  return check_launch("rfk_copy_channels");
}

extern "C" int rfk_coupling_tail(const float* nn_out, float* z, int B, int C, int HW, int clamp_type,
                                 const float* clamp_scale, const float* clamp_shift, float* logdet, int reverse,
                                 void* stream) {
  RFK_REQUIRE(nn_out && z && B > 0 && C > 0 && C % 2 == 0 && HW > 0, "rfk_coupling_tail: bad shape (C must be even)");
  RFK_REQUIRE(clamp_type >= 0 && clamp_type <= 3, "rfk_coupling_tail: unknown clamp_type %d", clamp_type);
  RFK_REQUIRE(clamp_type != RFK_CLAMP_REALNVP || (clamp_scale && clamp_shift),
              "rfk_coupling_tail: realnvp clamp needs scale and scale_shift");
  long long per = (long long)(C / 2) * HW;
  bool vec = HW % 4 == 0 && aligned16(nn_out) && aligned16(z);
  int chunks = ceil_div(vec ? per / 4 : per, kThreads);
  int cap = ceil_div((long long)sm_count() * 8, B);
  if (chunks > cap) chunks = cap;
  dim3 grid(chunks, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec)
    RFK_LAUNCH((coupling_tail_kernel<true>), grid, kThreads, 0, st, nn_out, z, C, HW, clamp_type, clamp_scale, clamp_shift,
                                                          logdet, reverse);
  else
    RFK_LAUNCH((coupling_tail_kernel<false>), grid, kThreads, 0, st, nn_out, z, C, HW, clamp_type, clamp_scale, clamp_shift,
                                                           logdet, reverse);
  return check_launch("rfk_coupling_tail");
}

extern "C" int rfk_coupling_tail_taps(const float* taps, float* z, int B, int C, int H, int W, const float* scale,
                                      const float* shift, int clamp_type, const float* clamp_scale,
                                      const float* clamp_shift, float* logdet, int reverse, void* stream) {
  RFK_REQUIRE(taps && z && scale && shift && B > 0 && C > 0 && C % 2 == 0 && H > 0 && W > 0,
              "rfk_coupling_tail_taps: null pointer or bad shape (C must be even)");
  RFK_REQUIRE(clamp_type >= 0 && clamp_type <= 3, "rfk_coupling_tail_taps: unknown clamp_type %d", clamp_type);
  RFK_REQUIRE(clamp_type != RFK_CLAMP_REALNVP || (clamp_scale && clamp_shift),
              "rfk_coupling_tail_taps: realnvp clamp needs scale and scale_shift");
  long long per = (long long)(C / 2) * H * W;
  const bool v4 = W % 4 == 0 && aligned16(taps) && aligned16(z);
  int chunks = ceil_div(v4 ? per / 4 : per, kThreads);
  int cap = ceil_div((long long)sm_count() * 8, B);
  if (chunks > cap) chunks = cap;
  if (v4)
    RFK_LAUNCH(coupling_taps_v4_kernel, dim3(chunks, B), kThreads, 0, (cudaStream_t)stream, 
        taps, z, C, H, W, scale, shift, clamp_type, clamp_scale, clamp_shift, logdet, reverse);
  else
    RFK_LAUNCH(coupling_taps_kernel, dim3(chunks, B), kThreads, 0, (cudaStream_t)stream, 
        taps, z, C, H, W, scale, shift, clamp_type, clamp_scale, clamp_shift, logdet, reverse);
  return check_launch("rfk_coupling_tail_taps");
}

extern "C" int rfk_gauss_logp(const float* z, int z_C, int z_off, const float* params, int n, int B, int HW,
                              int pairing, int std_kind, float* logdet, void* stream) {
  RFK_REQUIRE(z && logdet && B > 0 && n > 0 && HW > 0, "rfk_gauss_logp: null pointer or empty shape");
  RFK_REQUIRE(z_off >= 0 && z_off + n <= z_C, "rfk_gauss_logp: bad z channel window");
  long long per = (long long)n * HW;
  int chunks = ceil_div(per, kThreads);
  int cap = ceil_div((long long)sm_count() * 8, B);
  if (chunks > cap) chunks = cap;
  if (HW % 4 == 0 && aligned16(z) && (!params || aligned16(params)) && per < (1LL << 31)) {
    chunks = std::min(cap, ceil_div(per / 4, kThreads));
    RFK_LAUNCH(gauss_logp_v4_kernel, dim3(chunks, B), kThreads, 0, (cudaStream_t)stream, z, z_C, z_off, params, n, HW, pairing,
               std_kind, logdet);
    return check_launch("rfk_gauss_logp");
  }
  RFK_LAUNCH(gauss_logp_kernel, dim3(chunks, B), kThreads, 0, (cudaStream_t)stream, z, z_C, z_off, params, n, HW, pairing,
                                                                            std_kind, logdet);
  return check_launch("rfk_gauss_logp");
}

extern "C" int rfk_gauss_sample(const float* eps, const float* params, int n, int B, int HW, int pairing,
                                int std_kind, float temperature, float* out, int out_C, int out_off,
                                void* stream) {
  RFK_REQUIRE(eps && out && B > 0 && n > 0 && HW > 0, "rfk_gauss_sample: null pointer or empty shape");
  RFK_REQUIRE(out_off >= 0 && out_off + n <= out_C, "rfk_gauss_sample: bad output channel window");
  long long per = (long long)n * HW;
  int chunks = ceil_div(per, kThreads);
  int cap = ceil_div((long long)sm_count() * 8, B);
  if (chunks > cap) chunks = cap;
  RFK_LAUNCH(gauss_sample_kernel, dim3(chunks, B), kThreads, 0, (cudaStream_t)stream, eps, params, n, HW, pairing, std_kind,
                                                                              temperature, out, out_C, out_off);
  return check_launch("rfk_gauss_sample");
}

extern "C" int rfk_convlstm_pointwise(const float* cc, const float* c_prev, const float* peep, float* h_out,
                                      float* c_next, int B, int Hc, int HW, void* stream) {
  RFK_REQUIRE(cc && c_prev && h_out && c_next && B > 0 && Hc > 0 && HW > 0,
              "rfk_convlstm_pointwise: null pointer or empty shape");
  long long total = (long long)B * Hc * HW;
  cudaStream_t st = (cudaStream_t)stream;
  bool vec = HW % 4 == 0 && aligned16(cc) && aligned16(c_prev) && aligned16(h_out) && aligned16(c_next) &&
             (!peep || aligned16(peep));
  if (vec)
    RFK_LAUNCH((lstm_pointwise_kernel<true>), stream_grid(total / 4, kThreads, 8), kThreads, 0, st, cc, c_prev, peep, h_out,
                                                                                          c_next, Hc, HW, total / 4);
  else
    RFK_LAUNCH((lstm_pointwise_kernel<false>), stream_grid(total, kThreads, 8), kThreads, 0, st, cc, c_prev, peep, h_out,
                                                                                       c_next, Hc, HW, total);
  return check_launch("rfk_convlstm_pointwise");
}

extern "C" int rfk_convlstm_pointwise_ws(float* cc, int cc_ld, const float* bias, const float* c_prev,
                                         long long c_prev_bstride, const float* peep, float* h_out, long long h_bstride,
                                         float* c_next, long long c_next_bstride, void* h_nhwc, int h_off, int h_ld,
                                         int B, int Hc, int HW, int zero_cc, void* stream) {
  RFK_REQUIRE(cc && h_out && c_next && B > 0 && Hc > 0 && HW > 0 && cc_ld >= 4 * Hc,
              "rfk_convlstm_pointwise_ws: null pointer, empty shape or cc_ld < 4*Hc");
  if (h_nhwc) RFK_REQUIRE(h_off >= 0 && h_off + Hc <= h_ld, "rfk_convlstm_pointwise_ws: h window exceeds h_ld");
  long long total = (long long)B * Hc * HW;
  RFK_LAUNCH(lstm_pointwise_ws_kernel, stream_grid(total, kThreads, 8), kThreads, 0, (cudaStream_t)stream, 
      cc, cc_ld, bias, c_prev, c_prev_bstride, peep, h_out, h_bstride, c_next, c_next_bstride, (__nv_bfloat16*)h_nhwc, h_off,
      h_ld, Hc, HW, zero_cc, total);
  return check_launch("rfk_convlstm_pointwise_ws");
}

extern "C" int rfk_affine_pos(const float* x, float* y, const float* a, const float* c, int B, long long n,
                              void* stream) {
  RFK_REQUIRE(x && y && a && c && B > 0 && n > 0, "rfk_affine_pos: null pointer or empty shape");
  const long long total = (long long)B * n;
  cudaStream_t st = (cudaStream_t)stream;
  if (n % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(a) && aligned16(c))
    RFK_LAUNCH(affine_pos_kernel, stream_grid(total / 4, kThreads, 8), kThreads, 0, st, x, y, a, c, n / 4, total / 4);
  else
    RFK_LAUNCH(affine_pos_scalar_kernel, stream_grid(total, kThreads, 8), kThreads, 0, st, x, y, a, c, n, total);
  return check_launch("rfk_affine_pos");
}

extern "C" int rfk_batch_stats_pos(const float* x, float* mean, float* var, int B, long long n, float eps,
                                   void* stream) {
  RFK_REQUIRE(x && mean && var && B > 0 && n > 0, "rfk_batch_stats_pos: null pointer or empty shape");
  RFK_LAUNCH(batch_stats_pos_kernel, stream_grid(n, kThreads, 8), kThreads, 0, (cudaStream_t)stream, x, mean, var, B, n,
             eps);
  return check_launch("rfk_batch_stats_pos");
}

extern "C" int rfk_add_scalar(float* logdet, const float* addend, float alpha, int B, void* stream) {
  RFK_REQUIRE(logdet && addend && B > 0, "rfk_add_scalar: null pointer or empty shape");
  RFK_LAUNCH(add_scalar_kernel, ceil_div(B, 256), 256, 0, (cudaStream_t)stream, logdet, addend, alpha, B);
  return check_launch("rfk_add_scalar");
}


// ------------------------------------------------------------------------------------------
// Weight repacking: fp32 [N, Cin, taps] conv weights -> the bf16 K-major GEMM operand [rows_pad, ktot] of
//   mode 0  the forward conv:        row n,        k = t*kp + j   <- W[n, perm[j], t]
//   mode 1  the data-gradient conv:  row r,        k = t*kp + co  <- W[co, perm[r], taps-1-t]   (flipped, in/out swapped)
//   mode 2  the tap-split 1x1 form:  row t*N + c,  k = j          <- W[c, j, t]
//   mode 3  tap-split data gradient: row t*R + j,  k = co         <- W[co, perm[j], taps-1-t]   (R = rows / taps)
// One launch per weight (the optimizer changes every weight every step, so this runs once per conv per step).
// ------------------------------------------------------------------------------------------
namespace rfk {
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ src, int N, int Cin, int taps, int mode,
                                                          const int* __restrict__ perm, int rows, int kp,
                                                          __nv_bfloat16* __restrict__ dst, int rows_pad, int ktot) {
  // no pdl_trigger(): the next conv kernel prefetches these weights BEFORE its dependency wait
  pdl_wait();
  const long long total = (long long)rows_pad * ktot;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ktot), k = (int)(i % ktot);
    float v = 0.0f;
    if (r < rows) {
      if (mode == 0) {
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
        const int ci = perm ? perm[j] : j;       // negative / out-of-range = a padding row of the tap segment
        if (k < N && ci >= 0 && ci < Cin) v = src[((long long)k * Cin + ci) * taps + (taps - 1 - t)];
      }
    }
    dst[i] = __float2bfloat16(v);
  }
}
}  // namespace rfk

extern "C" int rfk_pack_weight(const float* src, int N, int Cin, int taps, int mode, const int* perm, int rows, int kp,
                               void* dst, int rows_pad, int ktot, void* stream) {
  using namespace rfk;
  RFK_REQUIRE(src && dst && N > 0 && Cin > 0 && taps > 0 && rows > 0 && rows <= rows_pad && kp > 0 && ktot > 0,
              "rfk_pack_weight: null pointer or bad shape");
  RFK_REQUIRE(mode >= 0 && mode <= 3, "rfk_pack_weight: unknown mode %d", mode);
  RFK_REQUIRE(mode == 2 ? (ktot == kp && kp >= Cin && rows == taps * N)
              : mode == 3 ? (ktot == kp && kp >= N && rows % taps == 0)
                          : (ktot == taps * kp && kp >= (mode == 0 ? Cin : N)),
              "rfk_pack_weight: ktot=%d / kp=%d do not match mode %d", ktot, kp, mode);
  const long long total = (long long)rows_pad * ktot;
  RFK_LAUNCH(pack_weight_kernel, stream_grid(total, 256, 8), 256, 0, (cudaStream_t)stream, src, N, Cin, taps, mode, perm, rows,
             kp, (__nv_bfloat16*)dst, rows_pad, ktot);
  return check_launch("rfk_pack_weight");
}


// Conv weights with an ActNorm FOLDED IN.
//   mode 4 (forward, operand of rfk_coupling_nn_fused): the ActNorm that FOLLOWS the conv: row n is scaled by
//     s = exp(logs[n]) and 16 extra K columns carry the shift t = bias[n] * s as two bf16 words (t_hi, t_lo, 0 ...), which the
//     kernel multiplies by a constant-one operand: act(ActNorm(conv(x)))[n] = act(sum_k W'[n,k] x[k] + t_hi + t_lo).
//   mode 5 (data gradient, operand of rfk_conv_gemm_actbwd with scale = NULL): the ActNorm that PRODUCED the conv's input:
//     row r (the dgrad's output channel = the forward conv's input channel) is scaled by exp(logs[r]), so that the
//     epilogue's da = dh * act'(h) * e^{logs} needs no per-channel factor.
namespace rfk {
__global__ void __launch_bounds__(256) pack_weight_folded_kernel(const float* __restrict__ src, int N, int Cin, int taps, int mode,
                                                                 const int* __restrict__ perm, int rows, int kp,
                                                                 const float* __restrict__ logs, const float* __restrict__ bias,
                                                                 __nv_bfloat16* __restrict__ dst, int rows_pad, int ktot) {
  // no pdl_trigger(): the next conv kernel prefetches these weights BEFORE its dependency wait
  pdl_wait();
  const long long total = (long long)rows_pad * ktot;
  const int kmain = taps * kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ktot), k = (int)(i % ktot);
    float v = 0.0f;
    if (r < rows) {
      const float sc = expf(logs[r]);
      if (mode == 5) {
        const int t = k / kp, co = k % kp;
        if (co < N) v = src[((long long)co * Cin + (perm ? perm[r] : r)) * taps + (taps - 1 - t)] * sc;
      } else if (k < kmain) {
        const int t = k / kp, j = k % kp;
        if (j < Cin) v = src[((long long)r * Cin + (perm ? perm[j] : j)) * taps + t] * sc;
      } else if (k < kmain + 2) {
        const float sh = bias[r] * sc;
        const float hi = __bfloat162float(__float2bfloat16(sh));
        v = k == kmain ? hi : sh - hi;
      }
    }
    dst[i] = __float2bfloat16(v);
  }
}
}  // namespace rfk

extern "C" int rfk_pack_weight_folded(const float* src, int N, int Cin, int taps, int mode, const int* perm, int rows, int kp,
                                      const float* logs, const float* bias, void* dst, int rows_pad, int ktot, void* stream) {
  using namespace rfk;
  RFK_REQUIRE(src && dst && logs && N > 0 && Cin > 0 && taps > 0 && rows > 0 && rows <= rows_pad,
              "rfk_pack_weight_folded: null pointer or bad shape");
  RFK_REQUIRE(mode == 4 || mode == 5, "rfk_pack_weight_folded: mode %d (4 = forward + following ActNorm, 5 = data gradient + preceding ActNorm)", mode);
  if (mode == 4)
    RFK_REQUIRE(bias && rows == N && kp >= Cin && ktot == taps * kp + 16,
                "rfk_pack_weight_folded: mode 4 needs bias, rows = N, kp >= Cin and ktot = taps*kp + 16 (got rows=%d ktot=%d)", rows, ktot);
  else
    RFK_REQUIRE(perm == nullptr && rows == Cin && kp >= N && ktot == taps * kp,
                "rfk_pack_weight_folded: mode 5 needs perm = NULL, rows = Cin, kp >= N and ktot = taps*kp (got rows=%d ktot=%d)", rows, ktot);
  const long long total = (long long)rows_pad * ktot;
  RFK_LAUNCH(pack_weight_folded_kernel, stream_grid(total, 256, 8), 256, 0, (cudaStream_t)stream, src, N, Cin, taps, mode, perm, rows,
             kp, logs, bias, (__nv_bfloat16*)dst, rows_pad, ktot);
  return check_launch("rfk_pack_weight_folded");
}


// ------------------------------------------------------------------------------------------
// Gather of nine tap planes stored NHWC bf16 (the output of a tap-split 1x1 GEMM with N = 9*ns):
//   out[b, j, y, x] = sum_t T[b, y+ky-1, x+kx-1, t*ns + j]      (t = 3*ky + kx; zero outside the image), fp32 NCHW,
// ns = the per-tap channel stride (n rounded up to 8, so every tap segment is a whole number of 16-byte chunks).
// Used by the data gradient of a 3x3 conv with few input channels: the gradient tensor is read once by one GEMM instead
// of once per tap.  Every element of T is needed exactly once.  A CTA handles 64 consecutive pixels: in phase 1 a thread
// owns (pixel, 8-channel group) and issues nine 128-bit loads (channel-fastest = coalesced NHWC rows), phase 2 writes
// pixel-fastest (coalesced NCHW rows) through a shared-memory transpose.
// ------------------------------------------------------------------------------------------
namespace rfk {
constexpr int TG_PIX = 64;
__global__ void __launch_bounds__(256) taps_gather_nhwc_kernel(const __nv_bfloat16* __restrict__ T, int ld, int n, int ns, int B,
                                                               int H, int W, float* __restrict__ out, float* __restrict__ acc0,
                                                               int acc0_C, int n0, float* __restrict__ acc1, int acc1_C) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float tile[];   // [ns][TG_PIX + 1]
  const int HW = H * W, G = ns >> 3;
  const long long npix = (long long)B * HW;
  for (long long g0 = (long long)blockIdx.x * TG_PIX; g0 < npix; g0 += (long long)gridDim.x * TG_PIX) {
    for (int e = threadIdx.x; e < TG_PIX * G; e += blockDim.x) {
      const int pl = e / G, gq = e - pl * G;
      const long long g = g0 + pl;
      float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (g < npix) {
        const long long b = g / HW;
        const int p = (int)(g % HW), y = p / W, x = p - y * W;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = y + ky - 1;
          if (yy < 0 || yy >= H) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int xx = x + kx - 1;
            if (xx < 0 || xx >= W) continue;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(T + ((b * H + yy) * W + xx) * (long long)ld + (3 * ky + kx) * ns + 8 * gq));
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __bfloat1622float2(h2[k]);
              s[2 * k] += f.x;
              s[2 * k + 1] += f.y;
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) tile[(8 * gq + k) * (TG_PIX + 1) + pl] = s[k];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < TG_PIX * n; e += blockDim.x) {
      const int j = e / TG_PIX, pl = e - j * TG_PIX;
      const long long g = g0 + pl;
      if (g < npix) {
        const long long b = g / HW;
        const float v = tile[j * (TG_PIX + 1) + pl];
        if (out) out[(b * n + j) * HW + (g % HW)] = v;
        else if (j < n0) acc0[(b * acc0_C + j) * HW + (g % HW)] += v;          // accumulate form: channels [0, n0) -> acc0,
        else acc1[(b * acc1_C + (j - n0)) * HW + (g % HW)] += v;               // the rest -> the first channels of acc1
      }
    }
    __syncthreads();
  }
}
}  // namespace rfk

extern "C" int rfk_taps_gather_nhwc(const void* T, int ld, int n, int n_stride, int B, int H, int W, float* out, void* stream) {
  using namespace rfk;
  RFK_REQUIRE(T && out && n > 0 && n <= n_stride && n_stride <= 128 && n_stride % 8 == 0 && B > 0 && H > 0 && W > 0 &&
              ld >= 9 * n_stride && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(T) & 15) == 0,
              "rfk_taps_gather_nhwc: bad arguments (n <= n_stride <= 128, n_stride %% 8 == 0, ld >= 9*n_stride, ld %% 8 == 0)");
  const long long npix = (long long)B * H * W;
  long long ctas = std::min<long long>((npix + TG_PIX - 1) / TG_PIX, (long long)sm_count() * 16);
  RFK_LAUNCH(taps_gather_nhwc_kernel, (int)ctas, 256, (size_t)n_stride * (TG_PIX + 1) * sizeof(float), (cudaStream_t)stream,
             (const __nv_bfloat16*)T, ld, n, n_stride, B, H, W, out, (float*)nullptr, 0, 0, (float*)nullptr, 0);
  return check_launch("rfk_taps_gather_nhwc");
}

extern "C" int rfk_taps_gather_nhwc_acc(const void* T, int ld, int n, int n_stride, int B, int H, int W, float* acc0, int acc0_C,
                                        int n0, float* acc1, int acc1_C, void* stream) {
  using namespace rfk;
  RFK_REQUIRE(T && n > 0 && n <= n_stride && n_stride <= 128 && n_stride % 8 == 0 && B > 0 && H > 0 && W > 0 &&
              ld >= 9 * n_stride && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(T) & 15) == 0,
              "rfk_taps_gather_nhwc_acc: bad arguments (n <= n_stride <= 128, n_stride %% 8 == 0, ld >= 9*n_stride, ld %% 8 == 0)");
  RFK_REQUIRE(n0 >= 0 && n0 <= n && (n0 == 0 || (acc0 && acc0_C >= n0)) && (n0 == n || (acc1 && acc1_C >= n - n0)),
              "rfk_taps_gather_nhwc_acc: accumulation targets do not cover the %d channels (n0=%d, acc0_C=%d, acc1_C=%d)", n, n0,
              acc0_C, acc1_C);
  const long long npix = (long long)B * H * W;
  long long ctas = std::min<long long>((npix + TG_PIX - 1) / TG_PIX, (long long)sm_count() * 16);
  RFK_LAUNCH(taps_gather_nhwc_kernel, (int)ctas, 256, (size_t)n_stride * (TG_PIX + 1) * sizeof(float), (cudaStream_t)stream,
             (const __nv_bfloat16*)T, ld, n, n_stride, B, H, W, (float*)nullptr, acc0, acc0_C, n0, acc1, acc1_C);
  return check_launch("rfk_taps_gather_nhwc_acc");
}
