// Back-to-back 1x1 GEMMs of the affine coupling network, fused so that the 256-channel hidden tensor h2 never
// leaves the SM (sm_100a):
//
//   h2   = act(ActNorm(conv1x1(h1)))          GEMM2: [128 px x K1] . W2^T -> fp32 accumulator in TMEM
//   taps = tap-split form of conv3x3(h2)      GEMM3: [128 px x hid] . W9^T -> fp32 accumulator in TMEM
//
// (Flow/glow_modules.py:232-238: net.2 = Conv2dNorm 1x1 + activation, net.4 = Conv2dZeros 3x3; the tap-split form of
// the latter is described at rfk_coupling_tail_taps in rfk.h.)
//
// Per 128-pixel tile: TMA streams the h1 tile (K-major, 128-byte swizzle); tcgen05.mma (SS form) accumulates GEMM2
// into TMEM against the shared-memory-resident W2; the eight epilogue warps read that accumulator (tcgen05.ld), apply
// the ActNorm affine + activation, round to bf16 and write the result BACK INTO TENSOR MEMORY (tcgen05.st, two bf16
// per 32-bit column) as the A operand of GEMM3; tcgen05.mma (TS form: A from TMEM, B = resident W9 in shared memory)
// accumulates the 9*C tap planes; the epilogue warps store them as fp32 NCHW.  HBM traffic per pixel drops from
// 512 B (h1 in) + 512 B (h2 out) + 512 B (h2 in) + 4*9C B to 512 B + 4*9C B.
//
// Persistent, one CTA per SM.  Hand-offs (all mbarriers): smem stage full/empty (TMA <-> MMA), acc2_full (GEMM2 done),
// a3_ready[kc] (bf16 h2 channels [64kc, 64kc+64) in TMEM, all 8 epilogue warps), acc3_full (GEMM3 done).  GEMM2 of tile i+1 overlaps the
// tap-plane stores of tile i; program order of the two roles makes every TMEM region single-writer at any time.
#include <algorithm>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rfk {

constexpr int kB2BThreads = 640;  // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-19: epilogue (4 per scheduler)
constexpr int kB2BEpiWarp0 = 4;
constexpr int kB2BEpiWarps = 16;
constexpr int kB2BChunkArrivals = 8;  // warps that write one 64-channel chunk of h2: 4 quadrants x 2 column parts
constexpr int kB2BSmemLimit = 232448;

struct B2BArgs {
  int B, H, W;
  int tw_log2, th_log2, tiles_x, tiles_y, m_tiles;
  int k1chunks;        // input channels of GEMM2 / 64
  int hid;             // N of GEMM2 = K of GEMM3 (multiple of 64, <= 256)
  int n3, n3_pad;      // tap planes 9*C, padded to a multiple of 16 (<= 128)
  int stages;
  int act_fn;
  const float* scale2;
  const float* shift2;
  float* taps;         // fp32 NCHW [B, n3, H, W]
};

__global__ void __launch_bounds__(kB2BThreads, 1)
conv1x1_taps_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW2,
                    const __grid_constant__ CUtensorMap tmW9, const B2BArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  // shared memory: [W2: k1chunks x (hid x 128 B)] [W9: hid/64 x (n3_pad x 128 B)] [stages x 16 KB] [scale|shift] [barriers]
  const int k3chunks = g.hid >> 6;
  const uint32_t w2_chunk = (uint32_t)g.hid * 128u, w9_chunk = (uint32_t)g.n3_pad * 128u;
  const uint32_t w2_base = base, w9_base = base + g.k1chunks * w2_chunk;
  const uint32_t stage_base = w9_base + k3chunks * w9_chunk;
  constexpr uint32_t kStage = 128u * 128u;
  const uint32_t ss_off = g.k1chunks * w2_chunk + k3chunks * w9_chunk + g.stages * kStage;
  float* ss = reinterpret_cast<float*>(smem + ss_off);
  const uint32_t bar_off = ss_off + 2u * g.hid * 4u;
  const uint32_t bar_base = base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (g.stages + s); };
  const uint32_t acc2_full = bar_base + 8u * (2 * g.stages);
  const uint32_t acc3_full = bar_base + 8u * (2 * g.stages + 1);
  const uint32_t w_full = bar_base + 8u * (2 * g.stages + 2);
  auto a3_ready = [&](int kc) { return bar_base + 8u * (2 * g.stages + 3 + kc); };  // one per 64-channel chunk of h2
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8u * (2 * g.stages + 7));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmW9);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc2_full, 1);
    for (int kc = 0; kc < 4; ++kc) mbar_init(a3_ready(kc), kB2BChunkArrivals);
    mbar_init(acc3_full, 1);
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // both weight matrices become shared-memory resident (parameters: safe before the dependency wait)
    mbar_expect_tx(w_full, g.k1chunks * w2_chunk + k3chunks * w9_chunk);
    for (int kc = 0; kc < g.k1chunks; ++kc) tma_load_2d(w2_base + kc * w2_chunk, &tmW2, w_full, kc * 64, 0);
    for (int kc = 0; kc < k3chunks; ++kc) tma_load_2d(w9_base + kc * w9_chunk, &tmW9, w_full, kc * 64, 0);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x; i < g.hid; i += blockDim.x) {
    ss[i] = g.scale2 ? g.scale2[i] : 1.0f;
    ss[g.hid + i] = g.shift2 ? g.shift2[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // tensor memory columns: [0, hid) GEMM2 accumulator | [hid, hid + hid/2) bf16 h2 (A of GEMM3) | then GEMM3 accumulator
  const uint32_t t_acc2 = tmem_base, t_a3 = tmem_base + g.hid, t_acc3 = tmem_base + g.hid + (g.hid >> 1);

  const int nimg_log2 = 7 - g.tw_log2 - g.th_log2;
  auto tile_origin = [&](int mt, int& x0, int& y0, int& n0) {
    const int tx = mt % g.tiles_x;
    mt /= g.tiles_x;
    const int ty = mt % g.tiles_y;
    const int tn = mt / g.tiles_y;
    x0 = tx << g.tw_log2;
    y0 = ty << g.th_log2;
    n0 = tn << nimg_log2;
  };

  if (warp == 0) {
    // ===== TMA producer: the h1 tile, one 64-channel chunk per stage =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int mt = blockIdx.x; mt < g.m_tiles; mt += gridDim.x) {
        int x0, y0, n0;
        tile_origin(mt, x0, y0, n0);
        for (int kc = 0; kc < g.k1chunks; ++kc) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), kStage);
          tma_load_4d(stage_base + s * kStage, &tmA, full_bar(s), kc * 64, x0, y0, n0);
          if (++s == g.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.hid >> 3) << 17) | (8u << 24);
      const uint32_t idesc3 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.n3_pad >> 3) << 17) | (8u << 24);
      const uint64_t desc_hi = umma_desc_kmajor(0, 64);
      mbar_wait(w_full, 0);
      int s = 0;
      uint32_t ph = 0, tl = 0;
      for (int mt = blockIdx.x; mt < g.m_tiles; mt += gridDim.x, ++tl) {
        // GEMM2: safe to overwrite acc2 -- this thread has already waited for a3_ready of the previous tile,
        // i.e. every epilogue warp has finished reading the previous accumulator
        uint32_t accumulate = 0;
        for (int kc = 0; kc < g.k1chunks; ++kc) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint64_t adesc = desc_hi | (uint64_t)(((stage_base + s * kStage) & 0x3FFFFu) >> 4);
          const uint64_t bdesc = desc_hi | (uint64_t)(((w2_base + kc * w2_chunk) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(t_acc2, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc2, accumulate);
            accumulate = 1;
          }
          umma_commit(empty_bar(s));
          if (++s == g.stages) { s = 0; ph ^= 1u; }
        }
        umma_commit(acc2_full);
        // GEMM3: A = bf16 h2 tile in tensor memory (8 columns per K=16 step), B = resident W9; chunk kc starts as soon
        // as the epilogue warps have written h2 channels [64 kc, 64 kc + 64), overlapping the rest of epilogue 2
        accumulate = 0;
        for (int kc = 0; kc < k3chunks; ++kc) {
          mbar_wait(a3_ready(kc), tl & 1u);
          tc_fence_after();
          const uint64_t bdesc = desc_hi | (uint64_t)(((w9_base + kc * w9_chunk) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16_ts(t_acc3, t_a3 + (uint32_t)(kc * 32 + k * 8), bdesc + (uint64_t)(2 * k), idesc3, accumulate);
            accumulate = 1;
          }
        }
        umma_commit(acc3_full);
      }
    }
    __syncwarp();
  } else if (warp >= kB2BEpiWarp0) {
    // ===== epilogue warps: quadrant q owns TMEM lanes [32q, 32q+32) = tile rows; the two warps of a quadrant split columns
    const int q = warp & 3, part = (warp - kB2BEpiWarp0) >> 2;   // part 0..3: which 32-column slice of every 128
    const int row = q * 32 + lane;
    const int ppi_log2 = g.tw_log2 + g.th_log2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const long long plane = (long long)g.H * g.W;
    uint32_t tl = 0;
    for (int mt = blockIdx.x; mt < g.m_tiles; mt += gridDim.x, ++tl) {
      int x0, y0, n0;
      tile_origin(mt, x0, y0, n0);
      const int b = n0 + (row >> ppi_log2);
      const int y = y0 + ((row >> g.tw_log2) & ((1 << g.th_log2) - 1));
      const int x = x0 + (row & ((1 << g.tw_log2) - 1));
      const bool valid = b < g.B && y < g.H && x < g.W;
      // ---- epilogue 2: fp32 accumulator -> ActNorm affine + activation -> bf16 -> tensor memory (A operand of GEMM3)
      mbar_wait(acc2_full, tl & 1u);
      tc_fence_after();
      for (int c0 = 32 * part; c0 < g.hid; c0 += 128) {
        uint32_t v[32];
        tmem_ld16_nowait(t_acc2 + lane_off + c0, v);
        tmem_ld16_nowait(t_acc2 + lane_off + c0 + 16, v + 16);
        tmem_wait_ld();
        uint32_t pk[16];
        const float4* sc = reinterpret_cast<const float4*>(ss + c0);
        const float4* sh = reinterpret_cast<const float4*>(ss + g.hid + c0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 s4 = sc[k], h4 = sh[k];
          float a0 = fmaf(__uint_as_float(v[4 * k + 0]), s4.x, h4.x);
          float a1 = fmaf(__uint_as_float(v[4 * k + 1]), s4.y, h4.y);
          float a2 = fmaf(__uint_as_float(v[4 * k + 2]), s4.z, h4.z);
          float a3 = fmaf(__uint_as_float(v[4 * k + 3]), s4.w, h4.w);
          if (g.act_fn == RFK_ACT_LEAKY) {
            a0 = apply_act(a0, RFK_ACT_LEAKY); a1 = apply_act(a1, RFK_ACT_LEAKY);
            a2 = apply_act(a2, RFK_ACT_LEAKY); a3 = apply_act(a3, RFK_ACT_LEAKY);
          }
          uint32_t p0 = pack_bf16(a0, a1), p1 = pack_bf16(a2, a3);
          if (g.act_fn == RFK_ACT_RELU) { p0 = relu_bf16x2(p0); p1 = relu_bf16x2(p1); }
          pk[2 * k] = p0;
          pk[2 * k + 1] = p1;
        }
        tmem_st16(t_a3 + lane_off + (uint32_t)(c0 >> 1), pk);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a3_ready(c0 >> 6));   // this warp's share of 64-channel chunk c0/64 is in place
      }
      // ---- epilogue 3: tap planes -> fp32 NCHW
      mbar_wait(acc3_full, tl & 1u);
      tc_fence_after();
      for (int c0 = 16 * part; c0 < g.n3_pad; c0 += 64) {
        if (c0 >= g.n3) break;  // warp-uniform
        uint32_t r[16];
        tmem_ld16_nowait(t_acc3 + lane_off + c0, r);
        tmem_wait_ld();
        if (valid) {
          float* dst = g.taps + (((long long)b * g.n3 + c0) * g.H + y) * g.W + x;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < g.n3) dst[j * plane] = __uint_as_float(r[j]);
        }
      }
      // the next tile's epilogue 2 overwrites the bf16 tile only after acc3_full of THIS tile (GEMM3 has consumed it),
      // and GEMM3 of the next tile starts only after this warp's next a3_ready arrive, i.e. after the loads above
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace rfk

using namespace rfk;

extern "C" int rfk_conv1x1_taps_fused(const void* act, int B, int H, int W, int act_ld, int cin_pad, const void* w2,
                                      int hid, const float* scale2, const float* shift2, int act_fn, const void* w9,
                                      int n3, int n3_pad, float* taps, void* stream) {
  RFK_REQUIRE(act && w2 && w9 && taps && B > 0 && H > 0 && W > 0, "rfk_conv1x1_taps_fused: null pointer or empty shape");
  RFK_REQUIRE(!conv_split_mode(), "rfk_conv1x1_taps_fused: the hidden tile is bf16 in tensor memory; not available in split-precision mode");
  RFK_REQUIRE(cin_pad > 0 && cin_pad % 64 == 0 && cin_pad <= act_ld && act_ld % 8 == 0,
              "rfk_conv1x1_taps_fused: cin_pad=%d must be a multiple of 64 and <= act_ld=%d (multiple of 8)", cin_pad, act_ld);
  RFK_REQUIRE(hid >= 64 && hid % 64 == 0 && hid <= 256, "rfk_conv1x1_taps_fused: hidden=%d must be 64, 128, 192 or 256", hid);
  RFK_REQUIRE(n3 > 0 && n3 <= n3_pad && n3_pad % 16 == 0 && n3_pad <= 128,
              "rfk_conv1x1_taps_fused: n3=%d n3_pad=%d (multiple of 16, at most 128)", n3, n3_pad);
  RFK_REQUIRE(act_fn >= 0 && act_fn <= 2, "rfk_conv1x1_taps_fused: bad act_fn %d", act_fn);
  RFK_REQUIRE(hid + hid / 2 + n3_pad <= 512, "rfk_conv1x1_taps_fused: tensor memory budget exceeded");
  B2BArgs g;
  g.B = B; g.H = H; g.W = W; g.k1chunks = cin_pad / 64; g.hid = hid; g.n3 = n3; g.n3_pad = n3_pad; g.act_fn = act_fn;
  g.scale2 = scale2; g.shift2 = shift2; g.taps = taps;
  int twl = ilog2_ceil(W);
  if (twl > 7) twl = 7;
  int thl = ilog2_ceil(H);
  if (thl > 7 - twl) thl = 7 - twl;
  g.tw_log2 = twl; g.th_log2 = thl;
  const int TW = 1 << twl, TH = 1 << thl, NIMG = 128 / (TW * TH);
  g.tiles_x = ceil_div(W, TW);
  g.tiles_y = ceil_div(H, TH);
  g.m_tiles = g.tiles_x * g.tiles_y * ceil_div(B, NIMG);
  const long long resident = (long long)g.k1chunks * hid * 128 + (long long)(hid / 64) * n3_pad * 128;
  const long long fixed = 1024 + 2LL * hid * 4 + 8 * (2 * 8 + 7) + 16;
  int stages = (int)((kB2BSmemLimit - fixed - resident) / (128 * 128));
  if (stages > 8) stages = 8;
  RFK_REQUIRE(stages >= 2, "rfk_conv1x1_taps_fused: weights (%lld B) leave no room for the activation pipeline", resident);
  g.stages = stages;
  const size_t smem = (size_t)1024 + resident + (size_t)stages * 128 * 128 + 2 * hid * 4 + 8 * (2 * stages + 7) + 16;
  CUtensorMap tmA, tmW2, tmW9;
  int rc = encode_act_map(&tmA, "rfk_conv1x1_taps_fused", "A", act, cin_pad, act_ld, B, H, W, TW, TH, NIMG, 64);
  if (rc) return rc;
  rc = encode_weight_map(&tmW2, "rfk_conv1x1_taps_fused", w2, cin_pad, hid, hid, 64);
  if (rc) return rc;
  rc = encode_weight_map(&tmW9, "rfk_conv1x1_taps_fused", w9, hid, n3_pad, n3_pad, 64);
  if (rc) return rc;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv1x1_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("rfk_conv1x1_taps_fused: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e));
      return RFK_ECUDA;
    }
    configured = smem;
  }
  int ctas = sm_count();
  if (ctas > g.m_tiles) ctas = g.m_tiles;
  launch_kernel(conv1x1_taps_kernel, dim3(ctas), dim3(kB2BThreads), smem, (cudaStream_t)stream, tmA, tmW2, tmW9, g);
  return check_launch("rfk_conv1x1_taps_fused");
}
