// Convolution (1x1 / 3x3 'same') as an implicit GEMM on the 5th-generation tensor cores (sm_100a).
//
//   D[pixel, out_ch] = sum_{tap, c} A[pixel + tap_offset, c] * Wt[out_ch, tap, c]
//
// * activations are NHWC bf16, so for one filter tap the A operand of a 128-pixel tile is a dense
//   [128 x 64] K-major block.  It is fetched by ONE 4-D TMA box {64 ch, TW, TH, NIMG} whose (x,y)
//   origin is shifted by the tap; out-of-image pixels are zero-filled by the TMA unit, which is the
//   convolution's zero padding -- no im2col buffer and no halo code.
// * weights are bf16 [n_pad, taps*cin_pad] (K-major), fetched by a 2-D TMA box {64, BN}.
// * both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma reads directly.
// * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into an fp32 accumulator in TMEM;
//   completion is tracked with tcgen05.commit -> mbarrier.  A multi-stage full/empty mbarrier ring
//   decouples the TMA producer warp from the MMA warp.
// * four epilogue warps read the accumulator with tcgen05.ld (one pixel per thread) and apply a fused
//   epilogue: per-channel affine (+ReLU) to bf16 NHWC / f32 NCHW, the affine-coupling tail with the
//   per-sample log-det reduction, or the ConvLSTM cell update.
//
// Two CTAs are resident per SM (<=113 KB shared memory and <=256 TMEM columns each), so one CTA's
// epilogue overlaps the other's main loop.
#include <cuda.h>

#include "common.cuh"

namespace rfk {

constexpr int BM = 128;          // pixels per tile = UMMA M
constexpr int BK = 64;           // bf16 channels per pipeline stage (= one 128 B swizzle row)
constexpr int UMMA_K = 16;       // K of one tcgen05.mma.kind::f16
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int kGemmThreads = 192;  // warp 0: TMA producer, warp 1: TMEM alloc + MMA issue, warps 2-5: epilogue

struct GemmArgs {
  int B, H, W;
  int n;                 // real output channels
  int BN;                // tile width in output channels (UMMA N), multiple of 16, <= 256
  int taps, kchunks;     // kchunks = cin_pad / 64
  int tw_log2, th_log2;  // tile = NIMG x TH x TW pixels, TW*TH*NIMG = 128
  int tiles_x, tiles_y;
  int stages;
  int tmem_cols;
};

struct PlainEpi {
  const float* scale;
  const float* shift;
  int act_fn;
  int out_kind;
  void* out;
  int out_ld, out_off;
  int vec_ok;
};

struct CouplingEpi {
  const float* scale;
  const float* shift;
  float* z;
  int clamp_type;
  const float* cs;
  const float* csh;
  float* logdet;
  int reverse;
};

struct LstmEpi {
  const float* bias;  // permuted like the weight rows
  int hidden, ht, ht_pad;
  const float* c_prev;
  long long c_prev_bs;
  const float* peep;
  float* c_next;
  long long c_next_bs;
  float* h_out;
  long long h_bs;
  __nv_bfloat16* h_nhwc;
  int h_off, h_ld, h_vec_ok;
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                    // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == RFK_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == RFK_ACT_LEAKY) return v >= 0.0f ? v : 0.2f * v;
  return v;
}

// ------------------------------------------------------------------------------------------
// epilogues: each thread owns accumulator row `row` (= one pixel), columns [0, BN) of the tile
// ------------------------------------------------------------------------------------------
struct PixelCoord {
  int b, y, x;
  bool valid;
};

__device__ __forceinline__ void epilogue(const GemmArgs& g, const PlainEpi& e, uint32_t taddr, PixelCoord pc,
                                         int n_tile) {
  const long long pix = ((long long)pc.b * g.H + pc.y) * g.W + pc.x;
  for (int c0 = 0; c0 < g.BN; c0 += 16) {
    const int col0 = n_tile * g.BN + c0;
    if (col0 >= g.n) break;  // warp-uniform
    float v[16];
    tmem_ld16(taddr + c0, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int col = col0 + j;
      if (col < g.n) {
        float s = e.scale ? __ldg(e.scale + col) : 1.0f;
        float t = e.shift ? __ldg(e.shift + col) : 0.0f;
        v[j] = apply_act(fmaf(v[j], s, t), e.act_fn);
      }
    }
    if (!pc.valid) continue;
    if (e.out_kind == RFK_OUT_NHWC_BF16) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.out) + pix * e.out_ld + e.out_off + col0;
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        if (e.vec_ok && col0 + 8 * h8 + 8 <= g.n) {
          __nv_bfloat162 pk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) pk[k] = __floats2bfloat162_rn(v[8 * h8 + 2 * k], v[8 * h8 + 2 * k + 1]);
          *reinterpret_cast<uint4*>(dst + 8 * h8) = *reinterpret_cast<uint4*>(pk);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (col0 + 8 * h8 + k < g.n) dst[8 * h8 + k] = __float2bfloat16(v[8 * h8 + k]);
        }
      }
    } else {
      float* dst = reinterpret_cast<float*>(e.out) + (((long long)pc.b * g.n + col0) * g.H + pc.y) * g.W + pc.x;
      const long long plane = (long long)g.H * g.W;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (col0 + j < g.n) dst[j * plane] = v[j];
    }
  }
}

__device__ __forceinline__ void epilogue(const GemmArgs& g, const CouplingEpi& e, uint32_t taddr, PixelCoord pc,
                                         int /*n_tile*/) {
  const int half = g.n >> 1;
  const long long plane = (long long)g.H * g.W;
  float* zp = e.z + (((long long)pc.b * g.n + half) * g.H + pc.y) * g.W + pc.x;
  float acc = 0.0f;
  for (int c0 = 0; c0 < g.BN; c0 += 16) {
    if (c0 >= g.n) break;
    float v[16];
    tmem_ld16(taddr + c0, v);
    if (pc.valid) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = (c0 >> 1) + k;
        if (j < half) {
          const int cs_ = 2 * j, cr_ = 2 * j + 1;
          float sh = fmaf(v[2 * k], e.scale ? __ldg(e.scale + cs_) : 1.0f, e.shift ? __ldg(e.shift + cs_) : 0.0f);
          float raw = fmaf(v[2 * k + 1], e.scale ? __ldg(e.scale + cr_) : 1.0f,
                           e.shift ? __ldg(e.shift + cr_) : 0.0f);
          float a = 0.0f, bsh = 0.0f;
          if (e.clamp_type == RFK_CLAMP_REALNVP) { a = __ldg(e.cs + j); bsh = __ldg(e.csh + j); }
          float ls = clamp_ls(raw, e.clamp_type, a, bsh);
          acc += ls;
          float zv = zp[j * plane];
          zp[j * plane] = e.reverse ? zv * expf(-ls) - sh : (zv + sh) * expf(ls);
        }
      }
    }
  }
  if (e.logdet) {
    // rows of one image are contiguous in the tile: reduce inside aligned lane groups of min(32, TW*TH)
    const int ppi_log2 = g.tw_log2 + g.th_log2;
    const int seg = ppi_log2 >= 5 ? 32 : (1 << ppi_log2);
    for (int o = seg >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const int lane = threadIdx.x & 31;
    if ((lane & (seg - 1)) == 0 && pc.b < g.B) atomicAdd(e.logdet + pc.b, e.reverse ? -acc : acc);
  }
}

__device__ __forceinline__ void epilogue(const GemmArgs& g, const LstmEpi& e, uint32_t taddr, PixelCoord pc,
                                         int n_tile) {
  const long long plane = (long long)g.H * g.W;
  const long long pofs = (long long)pc.y * g.W + pc.x;
  const long long pix = ((long long)pc.b * g.H + pc.y) * g.W + pc.x;
  for (int j0 = 0; j0 < e.ht_pad; j0 += 8) {
    if (j0 >= e.ht) break;  // warp-uniform
    uint32_t ri[8], rf[8], ro[8], rg[8];
    tmem_ld8_nowait(taddr + 0 * e.ht_pad + j0, ri);
    tmem_ld8_nowait(taddr + 1 * e.ht_pad + j0, rf);
    tmem_ld8_nowait(taddr + 2 * e.ht_pad + j0, ro);
    tmem_ld8_nowait(taddr + 3 * e.ht_pad + j0, rg);
    tmem_wait_ld();
    if (!pc.valid) continue;
    const float* bb = e.bias ? e.bias + (long long)n_tile * g.BN + j0 : nullptr;
    float hv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      hv[k] = 0.0f;
      if (j0 + k < e.ht) {
        const int ch = n_tile * e.ht + j0 + k;
        float bi = 0, bf = 0, bo = 0, bg = 0;
        if (bb) {
          bi = __ldg(bb + k); bf = __ldg(bb + e.ht_pad + k); bo = __ldg(bb + 2 * e.ht_pad + k);
          bg = __ldg(bb + 3 * e.ht_pad + k);
        }
        const long long co = ch * plane + pofs;
        float c = e.c_prev ? e.c_prev[pc.b * e.c_prev_bs + co] : 0.0f;
        float wi = 0, wf = 0, wo = 0;
        if (e.peep) {
          const long long hp = (long long)e.hidden * plane;
          wi = __ldg(e.peep + co); wf = __ldg(e.peep + hp + co); wo = __ldg(e.peep + 2 * hp + co);
        }
        float ig = sigmoidf_(__uint_as_float(ri[k]) + bi + wi * c);
        float fg = sigmoidf_(__uint_as_float(rf[k]) + bf + wf * c);
        float gg = tanhf(__uint_as_float(rg[k]) + bg);
        float cn = fg * c + ig * gg;
        float og = sigmoidf_(__uint_as_float(ro[k]) + bo + wo * cn);
        float h = og * tanhf(cn);
        e.c_next[pc.b * e.c_next_bs + co] = cn;
        e.h_out[pc.b * e.h_bs + co] = h;
        hv[k] = h;
      }
    }
    if (e.h_nhwc) {
      __nv_bfloat16* dst = e.h_nhwc + pix * e.h_ld + e.h_off + n_tile * e.ht + j0;
      if (e.h_vec_ok && j0 + 8 <= e.ht) {
        __nv_bfloat162 pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) pk[k] = __floats2bfloat162_rn(hv[2 * k], hv[2 * k + 1]);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(pk);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (j0 + k < e.ht) dst[k] = __float2bfloat16(hv[k]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <class Epi>
__global__ void __launch_bounds__(kGemmThreads, 2)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmArgs g, const Epi ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t b_stage_bytes = (uint32_t)g.BN * BK * 2;
  const uint32_t stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  const uint32_t bar_base = base + g.stages * stage_bytes;  // full[s], empty[s], tmem_full, then the TMEM slot
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + g.stages * stage_bytes + 8 * (2 * g.stages + 1));
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (g.stages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * g.stages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> pixel origin
  const int nimg_log2 = 7 - g.tw_log2 - g.th_log2;
  int mt = blockIdx.x;
  const int tx = mt % g.tiles_x;
  mt /= g.tiles_x;
  const int ty = mt % g.tiles_y;
  const int tn = mt / g.tiles_y;
  const int x0 = tx << g.tw_log2, y0 = ty << g.th_log2, n0 = tn << nimg_log2;
  const int n_tile = blockIdx.y;
  const int k_iters = g.taps * g.kchunks;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < k_iters; ++it) {
        const int s = it % g.stages;
        const uint32_t ph = (uint32_t)(it / g.stages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const int tap = it / g.kchunks, kc = it - tap * g.kchunks;
        const int dy = g.taps == 9 ? tap / 3 - 1 : 0;
        const int dx = g.taps == 9 ? tap % 3 - 1 : 0;
        const uint32_t a_dst = base + s * stage_bytes;
        mbar_expect_tx(full_bar(s), stage_bytes);
        tma_load_4d(a_dst, &tmA, full_bar(s), kc * BK, x0 + dx, y0 + dy, n0);
        tma_load_2d(a_dst + A_STAGE_BYTES, &tmB, full_bar(s), it * BK, n_tile * g.BN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int it = 0; it < k_iters; ++it) {
        const int s = it % g.stages;
        const uint32_t ph = (uint32_t)(it / g.stages) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_addr = base + s * stage_bytes;
        const uint64_t adesc = umma_desc_sw128(a_addr);
        const uint64_t bdesc = umma_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 32 B (16 bf16) along K inside the 128 B swizzle row: +2 in 16-byte units
          umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0);
        }
        umma_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);   // accumulator complete
    }
    __syncwarp();
  } else {
    // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ppi_log2 = g.tw_log2 + g.th_log2;
    PixelCoord pc;
    pc.b = n0 + (row >> ppi_log2);
    pc.y = y0 + ((row >> g.tw_log2) & ((1 << g.th_log2) - 1));
    pc.x = x0 + (row & ((1 << g.tw_log2) - 1));
    pc.valid = pc.b < g.B && pc.y < g.H && pc.x < g.W;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    epilogue(g, ep, tmem_base + ((uint32_t)(q * 32) << 16), pc, n_tile);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int ilog2_ceil(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

struct Plan {
  GemmArgs g;
  CUtensorMap tmA, tmB;
  dim3 grid;
  size_t smem;
};

static int make_plan(Plan& p, const char* who, const void* act, int B, int H, int W, int act_ld, int cin_pad,
                     const void* wgt, int n, int n_pad, int taps, int BN) {
  RFK_REQUIRE(act && wgt && B > 0 && H > 0 && W > 0, "%s: null pointer or empty shape", who);
  RFK_REQUIRE(cin_pad > 0 && cin_pad % BK == 0 && cin_pad <= act_ld, "%s: cin_pad=%d must be a multiple of %d and <= act_ld=%d",
              who, cin_pad, BK, act_ld);
  RFK_REQUIRE(act_ld % 8 == 0, "%s: act_ld=%d must be a multiple of 8 (16-byte TMA strides)", who, act_ld);
  RFK_REQUIRE(taps == 1 || taps == 9, "%s: taps=%d (only 1x1 and 3x3 kernels)", who, taps);
  RFK_REQUIRE(n > 0 && n <= n_pad && n_pad % 16 == 0, "%s: n=%d n_pad=%d (n_pad must be a multiple of 16)", who, n, n_pad);
  RFK_REQUIRE(BN % 16 == 0 && BN >= 16 && BN <= 256 && n_pad % BN == 0, "%s: bad N tile %d for n_pad=%d", who, BN, n_pad);
  RFK_REQUIRE((reinterpret_cast<uintptr_t>(act) & 15) == 0 && (reinterpret_cast<uintptr_t>(wgt) & 15) == 0,
              "%s: act / wgt must be 16-byte aligned", who);
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("%s: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)", who);
    return RFK_ECUDA;
  }
  GemmArgs& g = p.g;
  g.B = B; g.H = H; g.W = W; g.n = n; g.BN = BN; g.taps = taps; g.kchunks = cin_pad / BK;
  int twl = ilog2_ceil(W);
  if (twl > 7) twl = 7;
  int thl = ilog2_ceil(H);
  if (thl > 7 - twl) thl = 7 - twl;
  g.tw_log2 = twl; g.th_log2 = thl;
  const int TW = 1 << twl, TH = 1 << thl, NIMG = BM / (TW * TH);
  g.tiles_x = ceil_div(W, TW);
  g.tiles_y = ceil_div(H, TH);
  const int tiles_n = ceil_div(B, NIMG);
  const int stage_bytes = A_STAGE_BYTES + BN * BK * 2;
  int stages = (112 * 1024 - 1024 - 256) / stage_bytes;  // two CTAs per SM
  if (const char* s = getenv("RFK_GEMM_STAGES")) stages = atoi(s);
  if (stages > 8) stages = 8;
  if (stages > taps * g.kchunks) stages = taps * g.kchunks;
  if (stages < 1) stages = 1;
  g.stages = stages;
  int cols = 32;
  while (cols < BN) cols <<= 1;
  g.tmem_cols = cols;
  p.smem = (size_t)stages * stage_bytes + 1024 + 8 * (2 * stages + 1) + 16;
  p.grid = dim3((unsigned)(g.tiles_x * g.tiles_y * tiles_n), (unsigned)(n_pad / BN));

  // A: NHWC bf16 viewed as 4-D {C, W, H, B}; box {64, TW, TH, NIMG}; OOB -> zeros (the conv padding)
  cuuint64_t dimsA[4] = {(cuuint64_t)cin_pad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strA[3] = {(cuuint64_t)act_ld * 2, (cuuint64_t)W * act_ld * 2, (cuuint64_t)H * W * act_ld * 2};
  cuuint32_t boxA[4] = {BK, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)NIMG};
  cuuint32_t ones[4] = {1, 1, 1, 1};
  CUresult r = enc(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(act), dimsA, strA, boxA, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled(A) failed with CUresult %d (B=%d H=%d W=%d ld=%d cin_pad=%d)", who, (int)r, B, H, W,
              act_ld, cin_pad);
    return RFK_ECUDA;
  }
  // B: weights [n_pad, taps*cin_pad] viewed as 2-D {K, N}; box {64, BN}
  const cuuint64_t ktot = (cuuint64_t)taps * cin_pad;
  cuuint64_t dimsB[2] = {ktot, (cuuint64_t)n_pad};
  cuuint64_t strB[1] = {ktot * 2};
  cuuint32_t boxB[2] = {BK, (cuuint32_t)BN};
  r = enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wgt), dimsB, strB, boxB, ones,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled(B) failed with CUresult %d (n_pad=%d ktot=%llu BN=%d)", who, (int)r, n_pad,
              (unsigned long long)ktot, BN);
    return RFK_ECUDA;
  }
  return RFK_OK;
}

template <class Epi>
static int launch(const Plan& p, const Epi& ep, cudaStream_t st, const char* who) {
  static size_t configured = 0;
  if (p.smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", who, p.smem, cudaGetErrorString(e));
      return RFK_ECUDA;
    }
    configured = p.smem;
  }
  conv_gemm_kernel<Epi><<<p.grid, kGemmThreads, p.smem, st>>>(p.tmA, p.tmB, p.g, ep);
  return check_launch(who);
}

static int pick_bn(int n_pad) {
  for (int bn = 256; bn >= 16; bn -= 16)
    if (n_pad % bn == 0) return bn;
  return 16;
}

}  // namespace rfk

using namespace rfk;

extern "C" int rfk_conv_gemm(const void* act, int B, int H, int W, int act_ld, int cin_pad, const void* wgt, int n,
                             int n_pad, int taps, const float* scale, const float* shift, int act_fn, int out_kind,
                             void* out, int out_ld, int out_off, void* stream) {
  RFK_REQUIRE(out, "rfk_conv_gemm: null output");
  RFK_REQUIRE(out_kind == RFK_OUT_NHWC_BF16 || out_kind == RFK_OUT_NCHW_F32, "rfk_conv_gemm: bad out_kind %d", out_kind);
  RFK_REQUIRE(act_fn >= 0 && act_fn <= 2, "rfk_conv_gemm: bad act_fn %d", act_fn);
  Plan p;
  int rc = make_plan(p, "rfk_conv_gemm", act, B, H, W, act_ld, cin_pad, wgt, n, n_pad, taps, pick_bn(n_pad));
  if (rc) return rc;
  PlainEpi e;
  e.scale = scale; e.shift = shift; e.act_fn = act_fn; e.out_kind = out_kind; e.out = out;
  e.out_ld = out_ld; e.out_off = out_off; e.vec_ok = 0;
  if (out_kind == RFK_OUT_NHWC_BF16) {
    RFK_REQUIRE(out_off >= 0 && out_off + n <= out_ld, "rfk_conv_gemm: output window [%d,%d) exceeds out_ld=%d", out_off,
                out_off + n, out_ld);
    e.vec_ok = out_ld % 8 == 0 && out_off % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  }
  return launch(p, e, (cudaStream_t)stream, "rfk_conv_gemm");
}

extern "C" int rfk_conv_gemm_coupling(const void* act, int B, int H, int W, int act_ld, int cin_pad, const void* wgt,
                                      int n, int n_pad, int taps, const float* scale, const float* shift, float* z,
                                      int clamp_type, const float* clamp_scale, const float* clamp_shift,
                                      float* logdet, int reverse, void* stream) {
  RFK_REQUIRE(z && n % 2 == 0, "rfk_conv_gemm_coupling: null z or odd channel count %d", n);
  RFK_REQUIRE(n_pad <= 256, "rfk_conv_gemm_coupling: C=%d does not fit one N tile", n);
  RFK_REQUIRE(clamp_type >= 0 && clamp_type <= 3, "rfk_conv_gemm_coupling: unknown clamp_type %d", clamp_type);
  RFK_REQUIRE(clamp_type != RFK_CLAMP_REALNVP || (clamp_scale && clamp_shift),
              "rfk_conv_gemm_coupling: realnvp clamp needs scale and scale_shift");
  Plan p;
  int rc = make_plan(p, "rfk_conv_gemm_coupling", act, B, H, W, act_ld, cin_pad, wgt, n, n_pad, taps, n_pad);
  if (rc) return rc;
  CouplingEpi e;
  e.scale = scale; e.shift = shift; e.z = z; e.clamp_type = clamp_type; e.cs = clamp_scale; e.csh = clamp_shift;
  e.logdet = logdet; e.reverse = reverse;
  return launch(p, e, (cudaStream_t)stream, "rfk_conv_gemm_coupling");
}

extern "C" int rfk_conv_gemm_lstm(const void* act, int B, int H, int W, int act_ld, int cin_pad, const void* wgt,
                                  int hidden, int ht, int ht_pad, int taps, const float* bias, const float* c_prev,
                                  long long c_prev_bstride, const float* peep, float* c_next,
                                  long long c_next_bstride, float* h_out, long long h_bstride, void* h_nhwc,
                                  int h_off, int h_ld, void* stream) {
  RFK_REQUIRE(c_next && h_out, "rfk_conv_gemm_lstm: null output");
  RFK_REQUIRE(hidden > 0 && ht > 0 && hidden % ht == 0 && ht_pad >= ht && ht_pad % 8 == 0 && 4 * ht_pad <= 256,
              "rfk_conv_gemm_lstm: bad hidden tiling hidden=%d ht=%d ht_pad=%d", hidden, ht, ht_pad);
  const int n_tiles = hidden / ht, BN = 4 * ht_pad, n_pad = n_tiles * BN;
  Plan p;
  int rc = make_plan(p, "rfk_conv_gemm_lstm", act, B, H, W, act_ld, cin_pad, wgt, n_pad, n_pad, taps, BN);
  if (rc) return rc;
  LstmEpi e;
  e.bias = bias; e.hidden = hidden; e.ht = ht; e.ht_pad = ht_pad; e.c_prev = c_prev; e.c_prev_bs = c_prev_bstride;
  e.peep = peep; e.c_next = c_next; e.c_next_bs = c_next_bstride; e.h_out = h_out; e.h_bs = h_bstride;
  e.h_nhwc = (__nv_bfloat16*)h_nhwc; e.h_off = h_off; e.h_ld = h_ld;
  e.h_vec_ok = h_nhwc && h_ld % 8 == 0 && h_off % 8 == 0 && ht % 8 == 0 && (reinterpret_cast<uintptr_t>(h_nhwc) & 15) == 0;
  if (h_nhwc) RFK_REQUIRE(h_off >= 0 && h_off + hidden <= h_ld, "rfk_conv_gemm_lstm: h window exceeds h_ld");
  return launch(p, e, (cudaStream_t)stream, "rfk_conv_gemm_lstm");
}
