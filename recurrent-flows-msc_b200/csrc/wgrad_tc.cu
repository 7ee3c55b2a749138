// Convolution weight gradient on the 5th-generation tensor cores (sm_100a): tcgen05.mma with BOTH operands MN-major.
//
//   dW[tap][n][c] = sum_p dY[p, n] * X[p + off(tap), c]        X, dY: NHWC bf16; taps outside the image contribute zero
//
// The reduction dimension is the PIXEL index, which is the slow dimension of both NHWC operands, so each operand tile is
// "MN-major" for the MMA: a 4-D TMA box {64 channels, TW, TH, NIMG} with SWIZZLE_128B lands in shared memory as 128 pixel
// rows of 128 bytes -- exactly the canonical MN-major SWIZZLE_128B layout (8 K-rows x 64 MN-elements per swizzle atom,
// 8-row groups 1024 B apart, 64-channel atoms one box = 16 KB apart).  The tap shift is a coordinate offset of the box of
// one operand; out-of-image rows are zero-filled by TMA, which is the convolution's zero padding.
//
// Roles: the operand with more channels is the M side (128 channels per CTA = two boxes); the other one is the N side
// (up to 256 columns per tap).  All taps of a CTA accumulate in tensor memory at the same time (taps x n_cols <= 512
// columns) over the CTA's whole slice of pixel tiles, so the M-side tile is loaded once per pixel tile and reused by every
// tap; the accumulators are read once at the end and added to dW with fp32 atomics (the pixel range is split over CTAs).
//
//   warp 0   TMA producer: M-side ring (2 stages x 32 KB), N-side ring (one stage per tap, nbox x 16 KB each)
//   warp 1   MMA issuer:   8 x tcgen05.mma (K = 16 pixels each) per (pixel tile, tap)
//   warp 2   TMEM allocation
//   warps 4-7 epilogue:    tcgen05.ld -> atomicAdd
#include <algorithm>
#include <cstdlib>

#include "tc_ptx.cuh"

namespace rfk {

namespace {

constexpr int WT_THREADS = 256;
constexpr int WT_BOX_BYTES = 128 * 128;     // 128 pixel rows x 64 bf16 channels
constexpr int WT_A_STAGES = 2;
constexpr int WT_SMEM_LIMIT = 232448;

struct WgradArgs {
  int tiles_x, tiles_y, tiles;   // pixel tiles (128 pixels each)
  int tw_log2, th_log2;
  int taps;                      // 1 or 9
  int sign;                      // +1: the N-side box is shifted by +off(tap); -1: by -off(tap)
  int m_total, n_total;          // real channel counts of the M / N side
  int m_tiles, n_chunks, tap_groups;
  int n_cols;                    // MMA N per tap (multiple of 16, <= 256)
  int nbox;                      // 64-channel boxes per N-side stage
  int tpc;                       // taps per CTA
  int b_stages;
  int tmem_cols;
  int transpose_out;             // 0: M side = output channels (cout), N side = input channels   1: the other way round
  int dw_ld;                     // row stride of dw
  int dw_rows;                   // rows per tap of dw (cout)
  int layout;                    // 0: dw[tap][cout][dw_ld]   1: dw[cout][dw_ld][tap] with the input channel mapped through perm
  const int* perm;
  float* dw;
  float* ws;                     // non-null: partial sums [slice][work item][128 rows][tpc*n_cols] with plain stores
};

// MN-major SWIZZLE_128B shared-memory matrix descriptor: 64-element atoms 16 KB apart (LBO), 8-row K groups 1 KB apart (SBO)
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(WT_BOX_BYTES >> 4) << 16;   // leading byte offset: next 64-channel atom
  d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: next group of 8 pixel rows
  d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                     // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(WT_THREADS, 1) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                     const __grid_constant__ CUtensorMap tmB,
                                                                     const WgradArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_bytes = 2u * WT_BOX_BYTES;
  const uint32_t b_bytes = (uint32_t)g.nbox * WT_BOX_BYTES;
  const uint32_t a_off = 0, b_off = WT_A_STAGES * a_bytes;
  const uint32_t bar_off = b_off + (uint32_t)g.b_stages * b_bytes;
  const uint32_t bar_base = base + bar_off;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (WT_A_STAGES + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * WT_A_STAGES + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * WT_A_STAGES + g.b_stages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * WT_A_STAGES + 2 * g.b_stages);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8u * (2 * WT_A_STAGES + 2 * g.b_stages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.y -> (m tile, n chunk, tap group)
  int w = blockIdx.y;
  const int tg = w % g.tap_groups;
  w /= g.tap_groups;
  const int nc = w % g.n_chunks;
  const int mt = w / g.n_chunks;
  const int tap0 = tg * g.tpc;
  const int ntap = min(g.tpc, g.taps - tap0);
  const int n0 = nc * 256;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < WT_A_STAGES; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < g.b_stages; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nimg_log2 = 7 - g.tw_log2 - g.th_log2;
  auto tile_origin = [&](int t, int& x0, int& y0, int& i0) {
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    x0 = tx << g.tw_log2;
    y0 = ty << g.th_log2;
    i0 = (t / g.tiles_y) << nimg_log2;
  };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      for (int t = blockIdx.x; t < g.tiles; t += gridDim.x) {
        int x0, y0, i0;
        tile_origin(t, x0, y0, i0);
        mbar_wait(a_empty(sa), pha ^ 1u);
        mbar_expect_tx(a_full(sa), a_bytes);
        const uint32_t a_dst = base + a_off + sa * a_bytes;
        tma_load_4d(a_dst, &tmA, a_full(sa), mt * 128, x0, y0, i0);
        tma_load_4d(a_dst + WT_BOX_BYTES, &tmA, a_full(sa), mt * 128 + 64, x0, y0, i0);
        if (++sa == WT_A_STAGES) { sa = 0; pha ^= 1u; }
        for (int tp = 0; tp < ntap; ++tp) {
          const int tap = tap0 + tp;
          const int dy = g.taps == 9 ? g.sign * (tap / 3 - 1) : 0, dx = g.taps == 9 ? g.sign * (tap % 3 - 1) : 0;
          mbar_wait(b_empty(sb), phb ^ 1u);
          mbar_expect_tx(b_full(sb), b_bytes);
          const uint32_t b_dst = base + b_off + sb * b_bytes;
          for (int bx = 0; bx < g.nbox; ++bx)
            tma_load_4d(b_dst + bx * WT_BOX_BYTES, &tmB, b_full(sb), n0 + bx * 64, x0 + dx, y0 + dy, i0);
          if (++sb == g.b_stages) { sb = 0; phb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // kind::f16, fp32 accumulate, bf16 x bf16, A and B MN-major, N = n_cols, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(g.n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      bool first = true;
      for (int t = blockIdx.x; t < g.tiles; t += gridDim.x) {
        mbar_wait(a_full(sa), pha);
        tc_fence_after();
        const uint64_t adesc = umma_desc_mnmajor(base + a_off + sa * a_bytes);
        for (int tp = 0; tp < ntap; ++tp) {
          mbar_wait(b_full(sb), phb);
          tc_fence_after();
          const uint64_t bdesc = umma_desc_mnmajor(base + b_off + sb * b_bytes);
          const uint32_t d_tmem = tmem_base + (uint32_t)(tp * g.n_cols);
#pragma unroll
          for (int k = 0; k < 8; ++k)   // 16 pixel rows = 2048 bytes per K step (descriptor address in 16-byte units)
            umma_bf16(d_tmem, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, (first && k == 0) ? 0u : 1u);
          umma_commit(b_empty(sb));
          if (++sb == g.b_stages) { sb = 0; phb ^= 1u; }
        }
        umma_commit(a_empty(sa));
        if (++sa == WT_A_STAGES) { sa = 0; pha ^= 1u; }
        first = false;
      }
      umma_commit(done_bar);
    }
  } else if (warp >= 4) {
    // ===== epilogue: accumulators -> dW (fp32 atomics; the pixel range is split over blockIdx.x) =====
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int m = mt * 128 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if (g.ws) {
      // partial tile of this pixel slice: row-major [128][tpc*n_cols], 64 contiguous bytes per tcgen05.ld
      const int cols_cta = g.tpc * g.n_cols;
      float* row = g.ws + (((long long)blockIdx.x * gridDim.y + blockIdx.y) * 128 + q * 32 + lane) * cols_cta;
      for (int c0 = 0; c0 < ntap * g.n_cols; c0 += 16) {
        uint32_t r[16];
        tmem_ld16_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(row + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                 __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    } else if (blockIdx.x < g.tiles) {
      for (int tp = 0; tp < ntap; ++tp) {
        const int tap = tap0 + tp;
        for (int c0 = 0; c0 < g.n_cols; c0 += 16) {
          uint32_t r[16];
          tmem_ld16_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tp * g.n_cols + c0), r);
          tmem_wait_ld();
          if (m < g.m_total) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = n0 + c0 + j;
              if (n < g.n_total) {
                const int co = g.transpose_out ? n : m, ci = g.transpose_out ? m : n;
                float* dst = g.layout ? g.dw + ((long long)co * g.dw_ld + (g.perm ? g.perm[ci] : ci)) * g.taps + tap
                                      : g.dw + ((long long)tap * g.dw_rows + co) * g.dw_ld + ci;
                atomicAdd(dst, __uint_as_float(r[j]));
              }
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols)
                 : "memory");
  }
}

// Sum of the per-slice partial tiles -> dW (+=), one thread per output element of the work-item tile space.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WgradArgs g, int slices, int work) {
  pdl_trigger();
  pdl_wait();
  const int cols_cta = g.tpc * g.n_cols;
  const long long per_slice = (long long)work * 128 * cols_cta;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_slice; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % cols_cta);
    const long long t = i / cols_cta;
    const int rowi = (int)(t % 128);
    int w = (int)(t / 128);
    const int tg = w % g.tap_groups;
    w /= g.tap_groups;
    const int nc = w % g.n_chunks, mt = w / g.n_chunks;
    const int tp = col / g.n_cols, tap = tg * g.tpc + tp;
    const int m = mt * 128 + rowi, n = nc * 256 + (col - tp * g.n_cols);
    if (tap >= g.taps || m >= g.m_total || n >= g.n_total) continue;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;   // independent chains: the loads of several slices are in flight
    int sl = 0;
    for (; sl + 4 <= slices; sl += 4) {
      s0 += __ldcs(g.ws + (sl + 0) * per_slice + i);
      s1 += __ldcs(g.ws + (sl + 1) * per_slice + i);
      s2 += __ldcs(g.ws + (sl + 2) * per_slice + i);
      s3 += __ldcs(g.ws + (sl + 3) * per_slice + i);
    }
    for (; sl < slices; ++sl) s0 += __ldcs(g.ws + sl * per_slice + i);
    const float s = (s0 + s1) + (s2 + s3);
    const int co = g.transpose_out ? n : m, ci = g.transpose_out ? m : n;
    float* dst = g.layout ? g.dw + ((long long)co * g.dw_ld + (g.perm ? g.perm[ci] : ci)) * g.taps + tap
                          : g.dw + ((long long)tap * g.dw_rows + co) * g.dw_ld + ci;
    *dst += s;
  }
}

// ------------------------------------------------------------------------------------------
// "Tap-pair" variant for 3x3 convs where one side has at most 64 channels (the coupling network's first and last conv).
// With the roles above such a layer issues MMAs with N = 16..64 columns, and tcgen05.mma costs ~128 cycles whatever N is
// (it is bound by reading the 128 x 16 A tile from shared memory), so the tensor pipe idles.  Here the NARROW operand is
// the M side and the two 64-channel atoms of one M = 128 tile hold TWO DIFFERENT TAPS (two boxes of the same tensor with
// different shifts); the wide operand is the N side with up to 256 columns.  Nine taps = five pairs -> 5 x 8 MMAs of
// N = 256 per pixel tile instead of 2 x 72 of N = 16/32.  Accumulators: pairs x n_cols <= 512 TMEM columns per CTA, so the
// pairs are spread over blockIdx.y; partial tiles always go through the workspace.
// ------------------------------------------------------------------------------------------
struct PairArgs {
  int tiles_x, tiles_y, tiles, tw_log2, th_log2;
  int sign;                      // shift sign applied to the narrow operand's boxes
  int small_total, big_total;
  int n_chunks, n_cols, nbox;    // wide operand: columns per MMA, 64-channel boxes per stage
  int ppc, pair_groups;          // tap pairs per CTA, groups over blockIdx.y
  int inner_stages;              // pair-tile ring depth
  int tmem_cols;
  int small_is_x;                // 1: narrow = X (input channels), wide = dY;  0: narrow = dY (output channels), wide = X
  int tpt;                       // taps per M = 128 tile: 2 (64-channel atoms, SWIZZLE_128B) or 4 (32-channel atoms, SWIZZLE_64B)
  int dw_ld, dw_rows, layout;
  const int* perm;
  float* dw;
  float* ws;
};

// MN-major SWIZZLE_64B descriptor: 32-element atoms (64-byte rows) one 8 KB box apart, 8-row K groups 512 B apart
__device__ __forceinline__ uint64_t umma_desc_mnmajor64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((128 * 64) >> 4) << 16;     // leading byte offset: next 32-channel atom (= next tap's box)
  d |= (uint64_t)(512 >> 4) << 32;            // stride byte offset: next group of 8 pixel rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                     // SWIZZLE_64B
  return d;
}

__global__ void __launch_bounds__(WT_THREADS, 1) conv_wgrad_pairs_kernel(const __grid_constant__ CUtensorMap tmS,
                                                                        const __grid_constant__ CUtensorMap tmW,
                                                                        const PairArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t o_bytes = (uint32_t)g.nbox * WT_BOX_BYTES, i_bytes = 2u * WT_BOX_BYTES;   // 2 x 16 KB or 4 x 8 KB
  const uint32_t s_box = i_bytes / (uint32_t)g.tpt;          // one tap's box of the narrow operand
  const int ngrp = (9 + g.tpt - 1) / g.tpt;                  // tap groups (one accumulator each): 5 pairs or 3 quads
  const uint32_t o_off = 0, i_off = 2u * o_bytes;
  const uint32_t bar_off = i_off + (uint32_t)g.inner_stages * i_bytes;
  const uint32_t bar_base = base + bar_off;
  auto o_full = [&](int s) { return bar_base + 8u * s; };
  auto o_empty = [&](int s) { return bar_base + 8u * (2 + s); };
  auto i_full = [&](int s) { return bar_base + 8u * (4 + s); };
  auto i_empty = [&](int s) { return bar_base + 8u * (4 + g.inner_stages + s); };
  const uint32_t done_bar = bar_base + 8u * (4 + 2 * g.inner_stages);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8u * (4 + 2 * g.inner_stages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pg = blockIdx.y % g.pair_groups, nc = blockIdx.y / g.pair_groups;
  const int pair0 = pg * g.ppc;
  const int npair = min(g.ppc, ngrp - pair0);
  const int n0 = nc * g.n_cols;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < 2; ++s) {
      mbar_init(o_full(s), 1);
      mbar_init(o_empty(s), 1);
    }
    for (int s = 0; s < g.inner_stages; ++s) {
      mbar_init(i_full(s), 1);
      mbar_init(i_empty(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nimg_log2 = 7 - g.tw_log2 - g.th_log2;
  if (warp == 0) {
    if (lane == 0) {   // ===== TMA producer =====
      int so = 0, si = 0;
      uint32_t pho = 0, phi = 0;
      for (int t = blockIdx.x; t < g.tiles; t += gridDim.x) {
        int tt = t;
        const int x0 = (tt % g.tiles_x) << g.tw_log2;
        tt /= g.tiles_x;
        const int y0 = (tt % g.tiles_y) << g.th_log2;
        const int i0 = (tt / g.tiles_y) << nimg_log2;
        mbar_wait(o_empty(so), pho ^ 1u);
        mbar_expect_tx(o_full(so), o_bytes);
        for (int bx = 0; bx < g.nbox; ++bx)
          tma_load_4d(base + o_off + so * o_bytes + bx * WT_BOX_BYTES, &tmW, o_full(so), n0 + bx * 64, x0, y0, i0);
        if (++so == 2) { so = 0; pho ^= 1u; }
        for (int pl = 0; pl < npair; ++pl) {
          const int ta = g.tpt * (pair0 + pl), nt = min(g.tpt, 9 - ta);   // this group's taps (the last group is short)
          mbar_wait(i_empty(si), phi ^ 1u);
          mbar_expect_tx(i_full(si), (uint32_t)nt * s_box);
          const uint32_t dst = base + i_off + si * i_bytes;
          for (int k = 0; k < nt; ++k) {
            const int tp = ta + k;
            tma_load_4d(dst + k * s_box, &tmS, i_full(si), 0, x0 + g.sign * (tp % 3 - 1), y0 + g.sign * (tp / 3 - 1), i0);
          }
          if (++si == g.inner_stages) { si = 0; phi ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ===== MMA issuer =====
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(g.n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int so = 0, si = 0;
      uint32_t pho = 0, phi = 0;
      bool first = true;
      for (int t = blockIdx.x; t < g.tiles; t += gridDim.x) {
        mbar_wait(o_full(so), pho);
        tc_fence_after();
        const uint64_t bdesc = umma_desc_mnmajor(base + o_off + so * o_bytes);
        for (int pl = 0; pl < npair; ++pl) {
          mbar_wait(i_full(si), phi);
          tc_fence_after();
          // narrow side: 64-channel atoms (128-byte rows) or 32-channel atoms (64-byte rows, SWIZZLE_64B: atoms 8 KB apart,
          // 8-row groups 512 B apart, 16 pixel rows = 1024 B per K step)
          const uint64_t adesc = g.tpt == 2 ? umma_desc_mnmajor(base + i_off + si * i_bytes)
                                            : umma_desc_mnmajor64(base + i_off + si * i_bytes);
          const uint64_t a_step = g.tpt == 2 ? 128u : 64u;
          const uint32_t d_tmem = tmem_base + (uint32_t)(pl * g.n_cols);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(d_tmem, adesc + a_step * k, bdesc + (uint64_t)(128 * k), idesc, (first && k == 0) ? 0u : 1u);
          umma_commit(i_empty(si));
          if (++si == g.inner_stages) { si = 0; phi ^= 1u; }
        }
        umma_commit(o_empty(so));
        if (++so == 2) { so = 0; pho ^= 1u; }
        first = false;
      }
      umma_commit(done_bar);
    }
  } else if (warp >= 4) {
    // ===== epilogue: partial tile [128][ppc*n_cols] of this pixel slice -> workspace =====
    const int q = warp & 3;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int cols_cta = g.ppc * g.n_cols;
    float* row = g.ws + (((long long)blockIdx.x * gridDim.y + blockIdx.y) * 128 + q * 32 + lane) * cols_cta;
    for (int c0 = 0; c0 < npair * g.n_cols; c0 += 16) {
      uint32_t r[16];
      tmem_ld16_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(row + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                               __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols)
                 : "memory");
  }
}

__global__ void __launch_bounds__(256) wgrad_pairs_reduce_kernel(const PairArgs g, int slices, int work) {
  pdl_trigger();
  pdl_wait();
  const int cols_cta = g.ppc * g.n_cols;
  const long long per_slice = (long long)work * 128 * cols_cta;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_slice; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % cols_cta);
    const long long t = i / cols_cta;
    const int rowi = (int)(t % 128);
    const int w = (int)(t / 128);
    const int pg = w % g.pair_groups, nc = w / g.pair_groups;
    const int pl = col / g.n_cols;
    const int atom = 128 / g.tpt;                                   // channels per tap inside the M = 128 tile
    const int tap = g.tpt * (pg * g.ppc + pl) + rowi / atom;
    const int c = rowi % atom, n = nc * g.n_cols + (col - pl * g.n_cols);
    if (tap >= 9 || c >= g.small_total || n >= g.big_total) continue;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;   // independent chains: the loads of several slices are in flight
    int sl = 0;
    for (; sl + 4 <= slices; sl += 4) {
      s0 += __ldcs(g.ws + (sl + 0) * per_slice + i);
      s1 += __ldcs(g.ws + (sl + 1) * per_slice + i);
      s2 += __ldcs(g.ws + (sl + 2) * per_slice + i);
      s3 += __ldcs(g.ws + (sl + 3) * per_slice + i);
    }
    for (; sl < slices; ++sl) s0 += __ldcs(g.ws + sl * per_slice + i);
    const float s = (s0 + s1) + (s2 + s3);
    const int co = g.small_is_x ? n : c, ci = g.small_is_x ? c : n;
    float* dst = g.layout ? g.dw + ((long long)co * g.dw_ld + (g.perm ? g.perm[ci] : ci)) * 9 + tap
                          : g.dw + ((long long)tap * g.dw_rows + co) * g.dw_ld + ci;
    *dst += s;
  }
}

}  // namespace

// Returns RFK_OK when the tensor-core path ran, a positive value when the shape is not covered (caller falls back).
int conv_wgrad_tc(const void* x, int x_ld, int cin, const void* dy, int dy_ld, int cout, int B, int H, int W, int taps,
                  float* dw, int dw_ld, int layout, const int* perm, void* ws, long long ws_bytes, cudaStream_t stream) {
  const char* who = "rfk_conv_wgrad";
  if (taps != 1 && taps != 9) return 1;
  static const bool no_ws = [] { const char* e = getenv("RFK_WGRAD_ATOMICS"); return e && e[0] == '1'; }();
  static const bool no_pairs = [] { const char* e = getenv("RFK_WGRAD_NO_PAIRS"); return e && e[0] == '1'; }();
  if (taps == 9 && std::min(cin, cout) <= 64 && ws && !no_ws && !no_pairs && (reinterpret_cast<uintptr_t>(ws) & 15) == 0) {
    PairArgs p{};
    p.small_is_x = cin <= cout ? 1 : 0;
    p.small_total = p.small_is_x ? cin : cout;
    p.big_total = p.small_is_x ? cout : cin;
    p.sign = p.small_is_x ? 1 : -1;
    p.dw = dw; p.dw_ld = dw_ld; p.dw_rows = cout; p.layout = layout; p.perm = perm; p.ws = (float*)ws;
    int twl = ilog2_ceil(W);
    if (twl > 7) twl = 7;
    int thl = ilog2_ceil(H);
    if (thl > 7 - twl) thl = 7 - twl;
    p.tw_log2 = twl; p.th_log2 = thl;
    const int TW = 1 << twl, TH = 1 << thl, NIMG = 128 / (TW * TH);
    p.tiles_x = ceil_div(W, TW);
    p.tiles_y = ceil_div(H, TH);
    p.tiles = p.tiles_x * p.tiles_y * ceil_div(B, NIMG);
    const int n_cols_total = (p.big_total + 15) / 16 * 16;
    // narrow side <= 32 channels: FOUR taps per M = 128 tile (32-channel SWIZZLE_64B atoms) and the wide side in chunks of
    // 128 columns -> three accumulators x 128 columns fit one CTA, so the wide operand is read once per chunk instead of
    // once per tap-pair group (the pair form moves 3x the wide operand through L2 -> SM and is bound by that)
    static const bool no_quads = [] { const char* e = getenv("RFK_WGRAD_NO_QUADS"); return e && e[0] == '1'; }();
    p.tpt = (p.small_total <= 32 && !no_quads) ? 4 : 2;
    const int chunk = p.tpt == 4 ? 128 : 256;
    const int ngrp = (9 + p.tpt - 1) / p.tpt;
    p.n_chunks = ceil_div(n_cols_total, chunk);
    p.n_cols = p.n_chunks == 1 ? n_cols_total : chunk;
    p.nbox = ceil_div(p.n_cols, 64);
    p.pair_groups = ceil_div(ngrp, std::min(ngrp, 512 / p.n_cols));
    p.ppc = ceil_div(ngrp, p.pair_groups);
    int cols = 32;
    while (cols < p.ppc * p.n_cols) cols <<= 1;
    p.tmem_cols = cols;
    const int outer = 2 * p.nbox * WT_BOX_BYTES;
    p.inner_stages = std::min(8, (WT_SMEM_LIMIT - 1024 - outer - 512) / (2 * WT_BOX_BYTES));
    const int work = p.n_chunks * p.pair_groups;
    int slices = std::max(1, sm_count() / work);
    if (slices > p.tiles) slices = p.tiles;
    const long long per_slice = (long long)work * 128 * p.ppc * p.n_cols;
    if (p.inner_stages >= 2 && per_slice * slices * 4 <= ws_bytes) {
      const size_t smem = 1024 + (size_t)outer + (size_t)p.inner_stages * 2 * WT_BOX_BYTES + 8 * (4 + 2 * p.inner_stages + 2) + 16;
      CUtensorMap tmS, tmW;
      // the narrow side's map covers whole box rows when the row stride allows it (a 4-channel tensor declared as 4 channels
      // makes every box row an 8-byte request, which TMA handles far slower than one 64-byte row); the extra channels are
      // rows of the M tile that the reduce kernel never reads, so their contents do not matter
      const int s_ld = p.small_is_x ? x_ld : dy_ld, s_box = p.tpt == 4 ? 32 : 64;
      int rc = encode_act_map(&tmS, who, "narrow side", p.small_is_x ? x : dy, std::max(p.small_total, std::min(s_ld, s_box)),
                              s_ld, B, H, W, TW, TH, NIMG, s_box);
      if (rc != RFK_OK) return rc;
      rc = encode_act_map(&tmW, who, "wide side", p.small_is_x ? dy : x, p.big_total, p.small_is_x ? dy_ld : x_ld, B, H, W, TW,
                          TH, NIMG, 64);
      if (rc != RFK_OK) return rc;
      static bool pattr = false;
      if (!pattr) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM_LIMIT);
        if (e != cudaSuccess) {
          set_error("%s: cudaFuncSetAttribute failed: %s", who, cudaGetErrorString(e));
          return RFK_ECUDA;
        }
        pattr = true;
      }
      RFK_LAUNCH(conv_wgrad_pairs_kernel, dim3(slices, work), WT_THREADS, smem, stream, tmS, tmW, p);
      rc = check_launch(who);
      if (rc != RFK_OK) return rc;
      RFK_LAUNCH(wgrad_pairs_reduce_kernel, stream_grid(per_slice, 256, 8), 256, 0, stream, p, slices, work);
      return check_launch(who);
    }
  }
  WgradArgs g{};
  const bool swap = cin > cout;   // M side = the operand with more channels
  const void* a_ptr = swap ? x : dy;
  const void* b_ptr = swap ? dy : x;
  const int a_ld = swap ? x_ld : dy_ld, b_ld = swap ? dy_ld : x_ld;
  g.m_total = swap ? cin : cout;
  g.n_total = swap ? cout : cin;
  g.sign = swap ? -1 : 1;
  g.transpose_out = swap ? 1 : 0;
  g.taps = taps;
  g.dw = dw;
  g.dw_ld = dw_ld;
  g.dw_rows = cout;
  g.layout = layout;
  g.perm = perm;
  int twl = ilog2_ceil(W);
  if (twl > 7) twl = 7;
  int thl = ilog2_ceil(H);
  if (thl > 7 - twl) thl = 7 - twl;
  g.tw_log2 = twl;
  g.th_log2 = thl;
  const int TW = 1 << twl, TH = 1 << thl, NIMG = 128 / (TW * TH);
  g.tiles_x = ceil_div(W, TW);
  g.tiles_y = ceil_div(H, TH);
  g.tiles = g.tiles_x * g.tiles_y * ceil_div(B, NIMG);
  g.m_tiles = ceil_div(g.m_total, 128);
  const int n_cols_total = (g.n_total + 15) / 16 * 16;
  g.n_chunks = ceil_div(n_cols_total, 256);
  g.n_cols = g.n_chunks == 1 ? n_cols_total : 256;
  g.nbox = ceil_div(g.n_cols, 64);
  const int max_tpc = 512 / g.n_cols;
  g.tap_groups = ceil_div(taps, max_tpc);
  g.tpc = ceil_div(taps, g.tap_groups);
  int cols = 32;
  while (cols < g.tpc * g.n_cols) cols <<= 1;
  g.tmem_cols = cols;
  const int fixed = 1024 + WT_A_STAGES * 2 * WT_BOX_BYTES + 8 * (2 * WT_A_STAGES + 2 * 16 + 2) + 16;
  g.b_stages = std::min(16, (WT_SMEM_LIMIT - fixed) / (g.nbox * WT_BOX_BYTES));
  if (g.b_stages < 2) return 1;
  const size_t smem = 1024 + (size_t)WT_A_STAGES * 2 * WT_BOX_BYTES + (size_t)g.b_stages * g.nbox * WT_BOX_BYTES +
                      8 * (2 * WT_A_STAGES + 2 * g.b_stages + 2) + 16;

  CUtensorMap tmA, tmB;
  // maps cover whole box rows where the row stride allows it (short inner extents make TMA slow; the extra channels land in
  // accumulator rows / columns that are never written back)
  int rc = encode_act_map(&tmA, who, "M-side", a_ptr, std::max(g.m_total, std::min(a_ld, g.m_tiles * 128)), a_ld, B, H, W, TW,
                          TH, NIMG, 64);
  if (rc != RFK_OK) return rc;
  rc = encode_act_map(&tmB, who, "N-side", b_ptr, std::max(g.n_total, std::min(b_ld, g.n_chunks * g.nbox * 64)), b_ld, B, H, W,
                      TW, TH, NIMG, 64);
  if (rc != RFK_OK) return rc;

  const int work = g.m_tiles * g.n_chunks * g.tap_groups;
  int slices = std::max(1, sm_count() / work);
  if (slices > g.tiles) slices = g.tiles;
  {
    // every pixel slice flushes its accumulators with one atomic per output element: bound the total
    static const long long budget = [] { const char* e = getenv("RFK_WGRAD_ATOMIC_BUDGET"); return e ? atoll(e) : (1LL << 62); }();
    const long long out_elems = (long long)taps * cout * cin;
    const long long cap = std::max<long long>(1, budget / out_elems);
    if (slices > cap) slices = (int)cap;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM_LIMIT);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute failed: %s", who, cudaGetErrorString(e));
      return RFK_ECUDA;
    }
    attr_set = true;
  }
  // several pixel slices: partial tiles through the workspace + a reduce kernel when it fits (no atomics: a flush of
  // 128 x 256 scattered fp32 atomics per CTA costs more than the MMAs of the deep levels), else atomics straight into dw
  const long long need = (long long)slices * work * 128 * g.tpc * g.n_cols * 4;
  g.ws = (ws && !no_ws && need <= ws_bytes && (reinterpret_cast<uintptr_t>(ws) & 15) == 0) ? (float*)ws : nullptr;
  RFK_LAUNCH(conv_wgrad_tc_kernel, dim3(slices, work), WT_THREADS, smem, stream, tmA, tmB, g);
  int rc2 = check_launch(who);
  if (rc2 != RFK_OK || !g.ws) return rc2;
  const long long per_slice = (long long)work * 128 * g.tpc * g.n_cols;
  RFK_LAUNCH(wgrad_reduce_kernel, stream_grid(per_slice, 256, 8), 256, 0, stream, g, slices, work);
  return check_launch(who);
}

}  // namespace rfk
