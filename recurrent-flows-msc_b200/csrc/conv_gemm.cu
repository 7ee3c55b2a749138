// Convolution (1x1 / 3x3 'same') as an implicit GEMM on the 5th-generation tensor cores (sm_100a).
//
//   D[pixel, out_ch] = sum_{tap, c} A[pixel + tap_offset, c] * Wt[out_ch, tap, c]
//
// * activations are NHWC bf16, so for one filter tap the A operand of a 128-pixel tile is a dense
//   [128 x 64] K-major block.  It is fetched by ONE 4-D TMA box {64 ch, TW, TH, NIMG} whose (x,y)
//   origin is shifted by the tap; out-of-image pixels are zero-filled by the TMA unit, which is the
//   convolution's zero padding -- no im2col buffer and no halo code.
// * weights are bf16 [n_pad, taps*cin_pad] (K-major).  When the whole [BN x K] slice fits in shared
//   memory next to the pipeline it is loaded ONCE per CTA (weight-stationary) and only activations
//   stream; otherwise a {64, BN} weight box travels with every activation stage.
// * both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma reads directly.
// * PERSISTENT: one CTA per SM walks the pixel tiles with a static stride.  One elected thread issues
//   tcgen05.mma (M=128, N=BN, K=16) into one of TWO fp32 accumulators in TMEM, so the epilogue of tile
//   i overlaps the main loop of tile i+1.  Producer -> MMA -> epilogue hand-offs are mbarriers
//   (TMA complete_tx, tcgen05.commit, and explicit arrives).
// * eight epilogue warps (two per TMEM lane quadrant) read the accumulator with tcgen05.ld and apply a
//   fused epilogue from shared-memory-staged per-channel scale/shift:
//     - affine (+ReLU) -> bf16, staged in swizzled shared memory and written with TMA stores (NHWC),
//       or f32 NCHW with direct coalesced stores;
//     - the affine-coupling tail with the per-sample log-det reduction (z2 updated in place);
//     - the ConvLSTM cell update.
#include <cuda.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rfk {

constexpr int BM = 128;          // pixels per tile = UMMA M
constexpr int BK = 64;           // bf16 channels per pipeline stage: 64 (128 B swizzle rows) or 32 (64 B rows, g.bk)
constexpr int UMMA_K = 16;       // K of one tcgen05.mma.kind::f16
constexpr int STG_BYTES = BM * 128;  // one 64-channel bf16 output block of a tile
constexpr int kGemmThreads = 384;    // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-11: epilogue
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int SMEM_LIMIT = 232448;   // 227 KB opt-in maximum per CTA

struct GemmArgs {
  int B, H, W;
  int n;                 // real output channels
  int BN;                // tile width in output channels (UMMA N), multiple of 16, <= 256
  int taps, kchunks;     // kchunks = cin_pad / bk (3 * cin_pad / bk in split-precision mode)
  int kch, lo_coord;     // split mode: the K loop of a tap has 3 parts of kch chunks that read the activation row's hi half,
                         // its lo half (channel lo_coord + ...) and the hi half again, against weights [hi | hi | lo];
                         // otherwise kch = kchunks (one part)
  int bk;                // K elements per chunk: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B, for Cin <= 32)
  int kgroup;            // K chunks per pipeline stage (one mbarrier round trip per stage)
  int kg_per_split;      // pipeline stages of K per CTA: all of them, or a 1/gridDim.z slice (split-K)
  int tw_log2, th_log2;  // tile = NIMG x TH x TW pixels, TW*TH*NIMG = 128
  int pair;              // 1: CTA pairs (cluster of 2, cta_group::2): M = 256 per MMA, the resident weights split over the pair
  int tiles_x, tiles_y, m_tiles;
  int stages;
  int tmem_cols;
  int b_resident;        // 1: the CTA's whole weight slice lives in shared memory
  int use_stg;           // 1: shared-memory staging buffers for TMA stores are allocated
  const float* scale;    // per output channel (length n_ss), nullable = 1
  const float* shift;    // per output channel (length n_ss), nullable = 0
  int n_ss;
  unsigned long long* timeline;  // debug: 8 globaltimer stamps per CTA (rfk_debug_set_timeline), else null
};

template <int ACT>
struct PlainEpi {
  static constexpr int kAct = ACT;
  int act_fn;
  int out_kind;
  void* out;
  int out_ld, out_off;
  int vec_ok;
  int tma_store;
  int lo_off;            // split-precision mode: the bf16 residual v - bf16(v) goes to channel + lo_off (0 = off)
};

// Data gradient of a conv whose INPUT was h = act(ActNorm(previous conv)): the epilogue turns the accumulator dh into
// da = dh * act'(h) * scale (bf16 NHWC through the TMA-store path) and accumulates the per-channel column sums of da.
// Replaces the separate rfk_act_affine_bwd pass (6 B/element) -- see rfk_conv_gemm_actbwd in rfk.h.
struct ActBwdEpi {
  const __nv_bfloat16* h;   // NHWC bf16 [pixels, h_ld]: the activation the previous layer produced (its sign = act')
  int h_ld;
  int act_fn;
  float* colsum;            // [n] fp32, accumulated with atomics: sum over pixels of da
};

struct CouplingEpi {
  float* z;
  int clamp_type;
  const float* cs;
  const float* csh;
  float* logdet;
  int reverse;
};

struct SplitKEpi {   // partial sums of a K slice: ws[pixel, col] += acc  (fp32, pixel-major, vector red)
  float* ws;
  int ld;
  // optional in-kernel fix-up: the CTA that adds the LAST slice of a tile reads the sums back, applies the affine +
  // activation, writes bf16 NHWC and clears workspace and counter (so no separate reduction launch is needed)
  unsigned int* counters;   // one per (n_tile, m_tile), zero before the launch; null = partial sums only
  long long slice_stride;   // != 0: slice z stores its partial tile at ws + z*slice_stride (no atomics, no zeroing)
  int k_split;
  int act_fn;
  __nv_bfloat16* out;
  int out_ld, out_off, vec_ok;
};

struct LstmEpi {
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
  int h_lo_off;          // split-precision mode: residual of h at channel + h_lo_off (0 = off)
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// debug cycle accounting: CNT_BEGIN/CNT_END accumulate clock64 deltas into a local counter when tracing is on
#define CNT_BEGIN() const long long cnt_t0_ = g.timeline ? clock64() : 0
#define CNT_END(var) do { if (g.timeline) var += clock64() - cnt_t0_; } while (0)
#define RFK_PUT(slot, val)                                                                                  \
  do {                                                                                                      \
    if (g.timeline) g.timeline[(((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (slot)] = (unsigned long long)(val); \
  } while (0)
#define RFK_STAMP(slot)                                                                               \
  do {                                                                                                \
    if (g.timeline) g.timeline[(((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (slot)] = gtime(); \
  } while (0)

// ------------------------------------------------------------------------------------------
// epilogues.  Thread (quadrant q, lane) owns accumulator row q*32+lane (= one pixel); the two warps
// of a quadrant (`half` 0/1) split the tile's columns.  `ss` = scale[BN] then shift[BN] in shared memory.
// ------------------------------------------------------------------------------------------
struct TileCtx {
  int b, y, x;      // this thread's pixel
  bool valid;
  int x0, y0, n0;   // tile origin
  int n_tile;
  int half, q;
  uint32_t stg;     // this half's staging buffer (shared-space address), 0 when not allocated
  uint32_t tmem_empty_bar;
  int remote_release;   // CTA pair, odd CTA: the accumulator-drained barrier lives in the leader CTA
  long long* dbg;   // 5 per-phase cycle counters of the TMA-store epilogue, or null
  int tile_id;      // n_tile * m_tiles + m_tile (split-K fix-up counter index)
  volatile unsigned int* flag;  // one shared-memory word for epilogue-wide broadcasts
};

__device__ __forceinline__ void release_accumulator(const TileCtx& t) {
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    if (t.remote_release) mbar_arrive_cluster(t.tmem_empty_bar, 0);
    else mbar_arrive(t.tmem_empty_bar);
  }
}

template <int ACT>
__device__ __forceinline__ void epilogue(const GemmArgs& g, const PlainEpi<ACT>& e, const CUtensorMap* tmO,
                                         uint32_t taddr, const float* ss, const TileCtx& t) {
  const int lane = threadIdx.x & 31;
  if (e.tma_store) {
    // ---- bf16 NHWC through swizzled shared memory + TMA store (the TMA unit clips out-of-range pixels/channels)
    const int nblk = (g.BN + 63) >> 6;
    const int r = t.q * 32 + lane;
    const bool issuer = lane == 0;            // every warp stores its own 32 rows (sub-box of the tile)
    const int r0 = t.q * 32;
    const int sub_x = t.x0 + (r0 & ((1 << g.tw_log2) - 1));
    const int sub_y = t.y0 + ((r0 >> g.tw_log2) & ((1 << g.th_log2) - 1));
    const int sub_n = t.n0 + (r0 >> (g.tw_log2 + g.th_log2));
    const int last_blk = ((nblk - 1 - t.half) & ~1) + t.half;  // last block this half owns (may be < half: none)
    bool released = false;
    for (int blk = t.half; blk < nblk; blk += 2) {
      const int cols = min(64, g.BN - blk * 64);
      long long t0 = t.dbg ? clock64() : 0;
#define EPI_PHASE(i) do { if (t.dbg) { long long t1 = clock64(); t.dbg[i] += t1 - t0; t0 = t1; } } while (0)
      uint32_t v[64];
#pragma unroll
      for (int c = 0; c < 64; c += 16)
        if (c < cols) tmem_ld16_nowait(taddr + blk * 64 + c, v + c);  // warp-uniform predicate
      tmem_wait_ld();
      EPI_PHASE(0);
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        if (c < cols) {
          const float4* sc = reinterpret_cast<const float4*>(ss + blk * 64 + c);
          const float4* sh = reinterpret_cast<const float4*>(ss + g.BN + blk * 64 + c);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 s4 = sc[k], h4 = sh[k];
            float a0 = fmaf(__uint_as_float(v[c + 4 * k + 0]), s4.x, h4.x);
            float a1 = fmaf(__uint_as_float(v[c + 4 * k + 1]), s4.y, h4.y);
            float a2 = fmaf(__uint_as_float(v[c + 4 * k + 2]), s4.z, h4.z);
            float a3 = fmaf(__uint_as_float(v[c + 4 * k + 3]), s4.w, h4.w);
            if (ACT == RFK_ACT_LEAKY) {
              a0 = apply_act(a0, RFK_ACT_LEAKY); a1 = apply_act(a1, RFK_ACT_LEAKY);
              a2 = apply_act(a2, RFK_ACT_LEAKY); a3 = apply_act(a3, RFK_ACT_LEAKY);
            }
            uint32_t p0 = pack_bf16(a0, a1), p1 = pack_bf16(a2, a3);
            if (ACT == RFK_ACT_RELU) {  // relu(bf16(x)) == bf16(relu(x)): do it on the packed pair (one HMNMX2)
              p0 = relu_bf16x2(p0);
              p1 = relu_bf16x2(p1);
            }
            pk[(c >> 1) + 2 * k] = p0;
            pk[(c >> 1) + 2 * k + 1] = p1;
          }
        }
      }
      if (blk == last_blk) { release_accumulator(t); released = true; }
      EPI_PHASE(1);
      if (issuer) bulk_wait_read0();   // this warp's previous TMA store has finished reading its staging rows
      __syncwarp();
      EPI_PHASE(2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (8 * j < cols) {
          const uint32_t dst = t.stg + r * 128 + ((j ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                       "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                       : "memory");
        }
      }
      fence_async_smem();
      __syncwarp();
      EPI_PHASE(3);
      if (issuer) {
        tma_store_4d(tmO, t.stg + r0 * 128, t.n_tile * g.BN + blk * 64, sub_x, sub_y, sub_n);
        bulk_commit();
      }
      EPI_PHASE(4);
    }
    if (!released) release_accumulator(t);
    return;
  }
  // ---- direct stores: f32 NCHW (coalesced over the pixels of a warp) or unaligned bf16 NHWC
  const long long pix = ((long long)t.b * g.H + t.y) * g.W + t.x;
  const long long plane = (long long)g.H * g.W;
  for (int c0 = 16 * t.half; c0 < g.BN; c0 += 32) {
    const int col0 = t.n_tile * g.BN + c0;
    if (col0 >= g.n) break;  // warp-uniform
    uint32_t r[16];
    tmem_ld16_nowait(taddr + c0, r);
    tmem_wait_ld();
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = apply_act(fmaf(__uint_as_float(r[j]), ss[c0 + j], ss[g.BN + c0 + j]), ACT);
    if (!t.valid) continue;
    if (e.out_kind == RFK_OUT_NHWC_BF16) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.out) + pix * e.out_ld + e.out_off + col0;
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        if (e.vec_ok && col0 + 8 * h8 + 8 <= g.n) {
          uint4 u = make_uint4(pack_bf16(v[8 * h8], v[8 * h8 + 1]), pack_bf16(v[8 * h8 + 2], v[8 * h8 + 3]),
                               pack_bf16(v[8 * h8 + 4], v[8 * h8 + 5]), pack_bf16(v[8 * h8 + 6], v[8 * h8 + 7]));
          *reinterpret_cast<uint4*>(dst + 8 * h8) = u;
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (col0 + 8 * h8 + k < g.n) dst[8 * h8 + k] = __float2bfloat16(v[8 * h8 + k]);
        }
      }
      if (e.lo_off) {   // split precision: second bf16 word holds what the first one rounded away
#pragma unroll
        for (int k = 0; k < 16; ++k)
          if (col0 + k < g.n) dst[e.lo_off + k] = __float2bfloat16(v[k] - __bfloat162float(__float2bfloat16(v[k])));
      }
    } else {
      float* dst = reinterpret_cast<float*>(e.out) + (((long long)t.b * g.n + col0) * g.H + t.y) * g.W + t.x;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (col0 + j < g.n) dst[j * plane] = v[j];
    }
  }
  release_accumulator(t);
}

// Butterfly reduce-scatter over the 32 lanes of a warp: every lane holds 32 values (one row of a 32-row x 32-column block);
// afterwards lane l holds the sum over all 32 rows of column colsum_col(l).  31 shuffles instead of 32 x 5.
__device__ __forceinline__ int colsum_col(int lane) {
  return ((lane >> 4) & 1) * 16 + ((lane >> 3) & 1) * 8 + ((lane >> 2) & 1) * 4 + ((lane >> 1) & 1) * 2 + (lane & 1);
}
__device__ __forceinline__ float warp_colsum32(float (&d)[32], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 16;
    const float send = up ? d[i] : d[i + 16];
    const float keep = up ? d[i + 16] : d[i];
    d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 8;
    const float send = up ? d[i] : d[i + 8];
    const float keep = up ? d[i + 8] : d[i];
    d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 4;
    const float send = up ? d[i] : d[i + 4];
    const float keep = up ? d[i + 4] : d[i];
    d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 2;
    const float send = up ? d[i] : d[i + 2];
    const float keep = up ? d[i + 2] : d[i];
    d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    const float send = up ? d[0] : d[1];
    const float keep = up ? d[1] : d[0];
    d[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return d[0];
}

// One 64-column block of the fused data-gradient + activation-backward epilogue.  The warp's 32 rows x 64 columns of the saved
// activation h were brought into `buf` (this warp's slice of a staging block, 128-byte swizzled) by a TMA load; every thread
// reads its own row from there, turns the accumulator into da = dh * act'(h) * scale, writes da back IN PLACE (same thread,
// same 128 bytes) and the caller TMA-stores the slice.  Returns the column sums through csum[2] (lane l: column colsum_col(l)
// of each 32-column half).
__device__ __forceinline__ void actbwd_block(const GemmArgs& g, const ActBwdEpi& e, uint32_t taddr, const float* ss, int c0,
                                             uint32_t row_smem, int rsw, bool valid, int lane, float (&cs)[2]) {
  uint32_t pk[32];
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t v[32];
    tmem_ld16_nowait(taddr + c0 + pass * 32, v);
    tmem_ld16_nowait(taddr + c0 + pass * 32 + 16, v + 16);
    uint32_t hw[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t src = row_smem + (((4 * pass + j) ^ rsw) << 4);
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hw[4 * j]), "=r"(hw[4 * j + 1]), "=r"(hw[4 * j + 2]),
                   "=r"(hw[4 * j + 3]) : "r"(src));
    }
    tmem_wait_ld();
    float d[32];
    const float4* sc4 = reinterpret_cast<const float4*>(ss + c0 + pass * 32);
    const bool has_scale = g.scale != nullptr;   // null: the ActNorm scale is folded into the weight rows (pack mode 5)
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      const float4 s4 = has_scale ? sc4[q4] : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
      const float2 ha = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[2 * q4]));
      const float2 hb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[2 * q4 + 1]));
      float m0, m1, m2, m3;
      if (e.act_fn == RFK_ACT_RELU) {
        m0 = ha.x > 0.0f ? s4.x : 0.0f; m1 = ha.y > 0.0f ? s4.y : 0.0f; m2 = hb.x > 0.0f ? s4.z : 0.0f; m3 = hb.y > 0.0f ? s4.w : 0.0f;
      } else if (e.act_fn == RFK_ACT_LEAKY) {
        m0 = ha.x > 0.0f ? s4.x : 0.2f * s4.x; m1 = ha.y > 0.0f ? s4.y : 0.2f * s4.y;
        m2 = hb.x > 0.0f ? s4.z : 0.2f * s4.z; m3 = hb.y > 0.0f ? s4.w : 0.2f * s4.w;
      } else {
        m0 = s4.x; m1 = s4.y; m2 = s4.z; m3 = s4.w;
      }
      if (!valid) m0 = m1 = m2 = m3 = 0.0f;   // a row outside the image can still see inside pixels through the 3x3 taps
      d[4 * q4] = __uint_as_float(v[4 * q4]) * m0;
      d[4 * q4 + 1] = __uint_as_float(v[4 * q4 + 1]) * m1;
      d[4 * q4 + 2] = __uint_as_float(v[4 * q4 + 2]) * m2;
      d[4 * q4 + 3] = __uint_as_float(v[4 * q4 + 3]) * m3;
      pk[pass * 16 + 2 * q4] = pack_bf16(d[4 * q4], d[4 * q4 + 1]);
      pk[pass * 16 + 2 * q4 + 1] = pack_bf16(d[4 * q4 + 2], d[4 * q4 + 3]);
    }
    cs[pass] = warp_colsum32(d, lane);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t dst = row_smem + ((j ^ rsw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]),
                 "r"(pk[4 * j + 3])
                 : "memory");
  }
}

__device__ __forceinline__ void epilogue(const GemmArgs& g, const ActBwdEpi& e, const CUtensorMap* tmO, uint32_t taddr,
                                         const float* ss, const TileCtx& t) {}   // (dispatched through epilogue_actbwd)

__device__ __forceinline__ void epilogue(const GemmArgs& g, const CouplingEpi& e, const CUtensorMap*, uint32_t taddr,
                                         const float* ss, const TileCtx& t) {
  const int half_c = g.n >> 1;
  const long long plane = (long long)g.H * g.W;
  float* zp = e.z + (((long long)t.b * g.n + half_c) * g.H + t.y) * g.W + t.x;
  float acc = 0.0f;
  for (int c0 = 16 * t.half; c0 < g.BN; c0 += 32) {
    if (c0 >= g.n) break;
    uint32_t r[16];
    tmem_ld16_nowait(taddr + c0, r);
    tmem_wait_ld();
    if (t.valid) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = (c0 >> 1) + k;
        if (j < half_c) {
          const int cs_ = c0 + 2 * k, cr_ = cs_ + 1;
          float sh = fmaf(__uint_as_float(r[2 * k]), ss[cs_], ss[g.BN + cs_]);
          float raw = fmaf(__uint_as_float(r[2 * k + 1]), ss[cr_], ss[g.BN + cr_]);
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
  release_accumulator(t);
  if (e.logdet) {
    // rows of one image are contiguous in the tile: reduce inside aligned lane groups of min(32, TW*TH)
    const int ppi_log2 = g.tw_log2 + g.th_log2;
    const int seg = ppi_log2 >= 5 ? 32 : (1 << ppi_log2);
    for (int o = seg >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const int lane = threadIdx.x & 31;
    if ((lane & (seg - 1)) == 0 && t.b < g.B) atomicAdd(e.logdet + t.b, e.reverse ? -acc : acc);
  }
}

__device__ __forceinline__ void epilogue(const GemmArgs& g, const SplitKEpi& e, const CUtensorMap*, uint32_t taddr,
                                         const float* ss, const TileCtx& t) {
  const long long pix = ((long long)t.b * g.H + t.y) * g.W + t.x;
  for (int c0 = 16 * t.half; c0 < g.BN; c0 += 32) {
    uint32_t r[16];
    tmem_ld16_nowait(taddr + c0, r);
    tmem_wait_ld();
    if (!t.valid) continue;
    if (e.slice_stride) {   // every K slice owns a workspace slab: plain stores, summed by the last-arriving slice
      // slab layout [tile][16-column group][128 rows][16 floats]: the 32 lanes of a warp (32 rows) write 2 KB contiguously
      float4* dst = reinterpret_cast<float4*>(e.ws + (long long)blockIdx.z * e.slice_stride + (long long)t.tile_id * g.BN * 128 +
                                              ((long long)(c0 >> 4) * 128 + (t.q * 32 + (threadIdx.x & 31))) * 16);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        __stcg(dst + k, make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]),
                                    __uint_as_float(r[4 * k + 3])));
      continue;
    }
    float* dst = e.ws + pix * e.ld + t.n_tile * g.BN + c0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * k), "f"(__uint_as_float(r[4 * k])),
                   "f"(__uint_as_float(r[4 * k + 1])), "f"(__uint_as_float(r[4 * k + 2])),
                   "f"(__uint_as_float(r[4 * k + 3]))
                   : "memory");
  }
  release_accumulator(t);
  if (!e.counters) return;
  if (e.slice_stride) {
    // ---- cooperative fix-up: all k_split CTAs of a tile are co-resident (the host only splits when pixel tiles x slices
    // fit the GPU), so each waits until every slice has stored its slab and then reduces ITS share of the column
    // groups -- the reduction of one tile is spread over k_split SMs instead of serialising 1 MB of loads on one.
    // release: the CTA barrier orders every epilogue thread's slab stores before the signalling thread's gpu-scope fence
    // (fences are cumulative), so ONE fence per CTA publishes the whole partial tile
    named_bar(3, kEpiWarps * 32);
    unsigned int* cnt = e.counters + t.tile_id;
    if (t.q == 0 && t.half == 0 && (threadIdx.x & 31) == 0) {
      __threadfence();
      atomicAdd(cnt, 1u);
      const long long t0 = clock64();
      while (atomicAdd(cnt, 0u) < (unsigned)e.k_split) {    // arrivals; departures count on from k_split
        __nanosleep(64);
        if (clock64() - t0 > 4000000000LL) __trap();
      }
      __threadfence();   // acquire side, again one fence per CTA; the slabs are read with ld.global.cg (L2) below
    }
    named_bar(3, kEpiWarps * 32);
    if (t.valid) {
      const int groups = g.BN >> 4;
      for (int gi = (int)blockIdx.z * 2 + t.half; gi < groups; gi += 2 * e.k_split) {
        const int c0 = gi << 4, col0 = t.n_tile * g.BN + c0;
        if (col0 >= g.n) continue;
        float v[16];
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          // all slices' loads of these 8 columns are issued before the first add (k_split <= 9: 18 x 128 bits in flight);
          // a rolled loop would serialise one L2 round trip per slice
          float4 part[9][2];
#pragma unroll
          for (int sl = 0; sl < 9; ++sl) {
            if (sl < e.k_split) {
              const float4* ss4 = reinterpret_cast<const float4*>(e.ws + (long long)sl * e.slice_stride +
                                                                  (long long)t.tile_id * g.BN * 128 +
                                                                  ((long long)gi * 128 + (t.q * 32 + (threadIdx.x & 31))) * 16) + 2 * hq;
              part[sl][0] = __ldcg(ss4);
              part[sl][1] = __ldcg(ss4 + 1);
            } else {
              part[sl][0] = part[sl][1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
          }
          float4 s0 = part[0][0], s1 = part[0][1];
#pragma unroll
          for (int sl = 1; sl < 9; ++sl) {
            s0.x += part[sl][0].x; s0.y += part[sl][0].y; s0.z += part[sl][0].z; s0.w += part[sl][0].w;
            s1.x += part[sl][1].x; s1.y += part[sl][1].y; s1.z += part[sl][1].z; s1.w += part[sl][1].w;
          }
          v[8 * hq + 0] = s0.x; v[8 * hq + 1] = s0.y; v[8 * hq + 2] = s0.z; v[8 * hq + 3] = s0.w;
          v[8 * hq + 4] = s1.x; v[8 * hq + 5] = s1.y; v[8 * hq + 6] = s1.z; v[8 * hq + 7] = s1.w;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = apply_act(fmaf(v[j], ss[c0 + j], ss[g.BN + c0 + j]), e.act_fn);
        __nv_bfloat16* dst = e.out + pix * e.out_ld + e.out_off + col0;
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          if (e.vec_ok && col0 + 8 * h8 + 8 <= g.n) {
            *reinterpret_cast<uint4*>(dst + 8 * h8) =
                make_uint4(pack_bf16(v[8 * h8], v[8 * h8 + 1]), pack_bf16(v[8 * h8 + 2], v[8 * h8 + 3]),
                           pack_bf16(v[8 * h8 + 4], v[8 * h8 + 5]), pack_bf16(v[8 * h8 + 6], v[8 * h8 + 7]));
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (col0 + 8 * h8 + k < g.n) dst[8 * h8 + k] = __float2bfloat16(v[8 * h8 + k]);
          }
        }
      }
    }
    named_bar(3, kEpiWarps * 32);   // all reads of the slabs are done before this CTA reports its departure
    if (t.q == 0 && t.half == 0 && (threadIdx.x & 31) == 0) {
      if (atomicAdd(cnt, 1u) == (unsigned)(2 * e.k_split - 1)) *cnt = 0u;   // last one out resets the counter
    }
    return;
  }
  // ---- fix-up by the last-arriving slice (all 256 epilogue threads take part in the hand-shake)
  __threadfence();
  named_bar(3, kEpiWarps * 32);
  volatile unsigned int* flag = t.flag;
  if (t.q == 0 && t.half == 0 && (threadIdx.x & 31) == 0) *flag = atomicAdd(e.counters + t.tile_id, 1u);
  named_bar(3, kEpiWarps * 32);
  const bool last = *flag == (unsigned)(e.k_split - 1);
  named_bar(3, kEpiWarps * 32);   // everyone has read the flag before the next tile may overwrite it
  if (!last) return;
  __threadfence();
  if (t.valid) {
    for (int c0 = 16 * t.half; c0 < g.BN; c0 += 32) {
      const int col0 = t.n_tile * g.BN + c0;
      if (col0 >= g.n) break;
      float4* src = reinterpret_cast<float4*>(e.ws + pix * e.ld + col0);
      float v[16];
      if (e.slice_stride) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = 0.0f;
        for (int sl = 0; sl < e.k_split; ++sl) {
          const float4* ss4 = reinterpret_cast<const float4*>(e.ws + (long long)sl * e.slice_stride + pix * e.ld + col0);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 a = __ldcg(ss4 + k);
            v[4 * k] += a.x; v[4 * k + 1] += a.y; v[4 * k + 2] += a.z; v[4 * k + 3] += a.w;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 a = __ldcg(src + k);
          src[k] = make_float4(0.f, 0.f, 0.f, 0.f);   // leave the workspace clean for the next launch
          v[4 * k] = a.x; v[4 * k + 1] = a.y; v[4 * k + 2] = a.z; v[4 * k + 3] = a.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = apply_act(fmaf(v[j], ss[c0 + j], ss[g.BN + c0 + j]), e.act_fn);
      __nv_bfloat16* dst = e.out + pix * e.out_ld + e.out_off + col0;
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        if (e.vec_ok && col0 + 8 * h8 + 8 <= g.n) {
          *reinterpret_cast<uint4*>(dst + 8 * h8) =
              make_uint4(pack_bf16(v[8 * h8], v[8 * h8 + 1]), pack_bf16(v[8 * h8 + 2], v[8 * h8 + 3]),
                         pack_bf16(v[8 * h8 + 4], v[8 * h8 + 5]), pack_bf16(v[8 * h8 + 6], v[8 * h8 + 7]));
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (col0 + 8 * h8 + k < g.n) dst[8 * h8 + k] = __float2bfloat16(v[8 * h8 + k]);
        }
      }
    }
  }
  if (t.q == 0 && t.half == 0 && (threadIdx.x & 31) == 0) e.counters[t.tile_id] = 0u;
}

__device__ __forceinline__ void epilogue(const GemmArgs& g, const LstmEpi& e, const CUtensorMap*, uint32_t taddr,
                                         const float* ss, const TileCtx& t) {
  const long long plane = (long long)g.H * g.W;
  const long long pofs = (long long)t.y * g.W + t.x;
  const long long pix = ((long long)t.b * g.H + t.y) * g.W + t.x;
  const float* bias = ss + g.BN;  // the staged `shift` vector is the (row-permuted) conv bias
  for (int j0 = 8 * t.half; j0 < e.ht_pad; j0 += 16) {
    if (j0 >= e.ht) break;  // warp-uniform
    uint32_t ri[8], rf[8], ro[8], rg[8];
    tmem_ld8_nowait(taddr + 0 * e.ht_pad + j0, ri);
    tmem_ld8_nowait(taddr + 1 * e.ht_pad + j0, rf);
    tmem_ld8_nowait(taddr + 2 * e.ht_pad + j0, ro);
    tmem_ld8_nowait(taddr + 3 * e.ht_pad + j0, rg);
    tmem_wait_ld();
    if (!t.valid) continue;
    float hv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      hv[k] = 0.0f;
      if (j0 + k < e.ht) {
        const int ch = t.n_tile * e.ht + j0 + k;
        const float bi = bias[j0 + k], bf = bias[e.ht_pad + j0 + k], bo = bias[2 * e.ht_pad + j0 + k],
                    bg = bias[3 * e.ht_pad + j0 + k];
        const long long co = ch * plane + pofs;
        float c = e.c_prev ? e.c_prev[t.b * e.c_prev_bs + co] : 0.0f;
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
        e.c_next[t.b * e.c_next_bs + co] = cn;
        e.h_out[t.b * e.h_bs + co] = h;
        hv[k] = h;
      }
    }
    if (e.h_nhwc) {
      __nv_bfloat16* dst = e.h_nhwc + pix * e.h_ld + e.h_off + t.n_tile * e.ht + j0;
      if (e.h_vec_ok && j0 + 8 <= e.ht) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(hv[0], hv[1]), pack_bf16(hv[2], hv[3]),
                                                    pack_bf16(hv[4], hv[5]), pack_bf16(hv[6], hv[7]));
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (j0 + k < e.ht) dst[k] = __float2bfloat16(hv[k]);
      }
      if (e.h_lo_off) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (j0 + k < e.ht) dst[e.h_lo_off + k] = __float2bfloat16(hv[k] - __bfloat162float(__float2bfloat16(hv[k])));
      }
    }
  }
  release_accumulator(t);
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
// kPair: CTA-pair build (cluster of two, tcgen05 cta_group::2).  A kernel that contains cta_group::2 instructions can only be
// launched with an even cluster width, so the pair mode is a separate instantiation, not a run-time switch.
template <class Epi, bool kPair>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmH, const GemmArgs g, const Epi ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

  // shared-memory map (all tile regions are multiples of 1024 B)
  const int k_iters = g.taps * g.kchunks;
  const uint32_t a_chunk_bytes = (uint32_t)BM * g.bk * 2;
  const int k_groups = g.kg_per_split;             // pipeline stages of K this CTA accumulates
  const int kg0 = blockIdx.z * g.kg_per_split;      // first one (split-K: gridDim.z slices)
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;   // CTA pair: 0 = leader (issues the MMAs, owns their barriers)
  const int bn_local = kPair ? g.BN >> 1 : g.BN;          // weight rows held by this CTA
  const uint32_t b_chunk_bytes = (uint32_t)bn_local * g.bk * 2;
  const uint32_t b_res_bytes = g.b_resident ? (uint32_t)(k_groups * g.kgroup) * b_chunk_bytes : 0u;
  const uint32_t stage_bytes = (uint32_t)g.kgroup * (a_chunk_bytes + (g.b_resident ? 0u : b_chunk_bytes));
  const uint32_t stage_base = base + b_res_bytes;
  const uint32_t stg_base = stage_base + g.stages * stage_bytes;
  const uint32_t ss_off = b_res_bytes + g.stages * stage_bytes + (uint32_t)g.use_stg * STG_BYTES;   // use_stg = staging blocks (0, 2, 4)
  float* ss = reinterpret_cast<float*>(smem + ss_off);
  const uint32_t bar_off = ss_off + 2u * g.BN * 4u;
  const uint32_t bar_base = base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (g.stages + s); };
  auto tmem_full_bar = [&](int b) { return bar_base + 8u * (2 * g.stages + b); };
  auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (2 * g.stages + 2 + b); };
  const uint32_t b_full_bar = bar_base + 8u * (2 * g.stages + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8u * (2 * g.stages + 5));
  auto h_bar = [&](int w, int b) { return bar_base + 8u * (2 * g.stages + 5) + 16u + 8u * (3 * w + b); };   // ActBwdEpi: per warp, per buffer

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y;
  if (threadIdx.x == 0) RFK_STAMP(0);  // CTA start

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmH);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), kEpiWarps * (kPair ? 2 : 1));
    }
    mbar_init(b_full_bar, 1);
    for (int w = 0; w < kEpiWarps; ++w)
      for (int b = 0; b < 3; ++b) mbar_init(h_bar(w, b), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (kPair) cluster_sync_all();   // the peer's barriers exist before anything (TMA, commits, remote arrives) targets them
  if (warp == 2) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                   "r"((uint32_t)g.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                   "r"((uint32_t)g.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  // Programmatic dependent launch: everything above touched only this CTA's own shared memory / TMEM.  The weights
  // are parameters (never written by the preceding kernel), so their TMA loads also go out before the dependency wait.
  if (warp == 0 && lane == 0 && g.b_resident) {
    if (kPair) {   // each CTA fetches its half of the rows; both halves count on the leader's barrier
      if (rank == 0) mbar_expect_tx(b_full_bar, 2u * b_res_bytes);
      for (int it = 0; it < k_groups * g.kgroup; ++it)
        tma_load_2d_2sm(base + it * b_chunk_bytes, &tmB, b_full_bar & kPeerBitMask, (kg0 * g.kgroup + it) * g.bk,
                        n_tile * g.BN + (int)rank * bn_local);
    } else {
      mbar_expect_tx(b_full_bar, b_res_bytes);
      for (int it = 0; it < k_groups * g.kgroup; ++it)
        tma_load_2d(base + it * b_chunk_bytes, &tmB, b_full_bar, (kg0 * g.kgroup + it) * g.bk, n_tile * g.BN);
    }
  }
  pdl_trigger();   // the next kernel in the stream may start its own prologue
  pdl_wait();      // from here on we read what the preceding kernel produced
  // per-channel scale / shift of this CTA's output-channel slice -> shared memory
  for (int i = threadIdx.x; i < g.BN; i += blockDim.x) {
    const int col = n_tile * g.BN + i;
    ss[i] = (g.scale && col < g.n_ss) ? g.scale[col] : 1.0f;
    ss[g.BN + i] = (g.shift && col < g.n_ss) ? g.shift[col] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) RFK_STAMP(1);  // setup done

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

  // tile schedule: CTA i takes tiles i, i + grid, ...; a CTA pair takes tile PAIRS (2*it + rank); with an odd tile count
  // the odd CTA's last tile does not exist (its loads are zero-filled, its stores clipped)
  const int it0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int it_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int it_end = kPair ? (g.m_tiles + 1) >> 1 : g.m_tiles;
  auto tile_of = [&](int it) { return kPair ? 2 * it + (int)rank : it; };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      long long c_wait_empty = 0;
      for (int it = it0; it < it_end; it += it_step) {
        const int mt = tile_of(it);
        int x0, y0, n0;
        tile_origin(mt, x0, y0, n0);
        int tap = (kg0 * g.kgroup) / g.kchunks, kc = (kg0 * g.kgroup) % g.kchunks;
        for (int grp = 0; grp < k_groups; ++grp) {
          { CNT_BEGIN(); mbar_wait(empty_bar(s), ph ^ 1u); CNT_END(c_wait_empty); }
          const uint32_t a_dst = stage_base + s * stage_bytes;
          if (!kPair) mbar_expect_tx(full_bar(s), stage_bytes);
          else if (rank == 0) mbar_expect_tx(full_bar(s), 2u * stage_bytes);   // both CTAs' tiles land on the leader's barrier
          for (int j = 0; j < g.kgroup; ++j) {
            const int dy = g.taps == 9 ? tap / 3 - 1 : 0;
            const int dx = g.taps == 9 ? tap % 3 - 1 : 0;
            const int part = kc / g.kch;
            const int ch0 = (kc - part * g.kch) * g.bk + (part == 1 ? g.lo_coord : 0);
            if (kPair)
              tma_load_4d_2sm(a_dst + j * a_chunk_bytes, &tmA, full_bar(s) & kPeerBitMask, ch0, x0 + dx, y0 + dy, n0);
            else
              tma_load_4d(a_dst + j * a_chunk_bytes, &tmA, full_bar(s), ch0, x0 + dx, y0 + dy, n0);
            if (!g.b_resident)
              tma_load_2d(a_dst + g.kgroup * a_chunk_bytes + j * b_chunk_bytes, &tmB, full_bar(s),
                          ((kg0 + grp) * g.kgroup + j) * g.bk, n_tile * g.BN);
            if (++kc == g.kchunks) { kc = 0; ++tap; }
          }
          if (++s == g.stages) { s = 0; ph ^= 1u; }
        }
      }
      RFK_STAMP(3);  // all loads issued
      RFK_PUT(8, c_wait_empty);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && rank == 0) {   // CTA pair: only the leader issues
      // kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24 (M = 256 for a pair)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.BN >> 3) << 17) |
                             ((uint32_t)((kPair ? 2 * BM : BM) >> 4) << 24);
      const uint64_t desc_hi = umma_desc_kmajor(0, g.bk);  // everything but the start address
      const int ksteps = g.bk / UMMA_K;
      if (g.b_resident) {
        mbar_wait(b_full_bar, 0);
        RFK_STAMP(2);  // weights resident
      }
      int s = 0;
      uint32_t ph = 0, tl = 0;
      long long c_wait_full = 0, c_wait_tempty = 0, c_issue = 0;
      for (int it = it0; it < it_end; it += it_step, ++tl) {
        const uint32_t buf = tl & 1u;
        { CNT_BEGIN(); mbar_wait(tmem_empty_bar(buf), ((tl >> 1) & 1u) ^ 1u); CNT_END(c_wait_tempty); }  // accumulator drained
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * g.BN;
        uint32_t b_res_addr = base;
        uint32_t accumulate = 0;
        for (int grp = 0; grp < k_groups; ++grp) {
          { CNT_BEGIN(); mbar_wait(full_bar(s), ph); CNT_END(c_wait_full); }
          tc_fence_after();
          CNT_BEGIN();
          uint32_t a_addr = stage_base + s * stage_bytes;
          uint32_t b_addr = g.b_resident ? b_res_addr : a_addr + g.kgroup * a_chunk_bytes;
          for (int j = 0; j < g.kgroup; ++j) {
            const uint64_t adesc = desc_hi | (uint64_t)((a_addr & 0x3FFFFu) >> 4);
            const uint64_t bdesc = desc_hi | (uint64_t)((b_addr & 0x3FFFFu) >> 4);
#pragma unroll 4
            for (int k = 0; k < ksteps; ++k) {
              // advance 32 B (16 bf16) along K inside the swizzle row: +2 in 16-byte units
              if (kPair) umma_bf16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accumulate);
              else umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accumulate);
              accumulate = 1;
            }
            a_addr += a_chunk_bytes;
            b_addr += b_chunk_bytes;
          }
          b_res_addr += g.kgroup * b_chunk_bytes;
          if (kPair) umma_commit_2sm(empty_bar(s));   // frees the stage in BOTH CTAs
          else umma_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
          CNT_END(c_issue);
          if (++s == g.stages) { s = 0; ph ^= 1u; }
        }
        if (kPair) umma_commit_2sm(tmem_full_bar(buf));
        else umma_commit(tmem_full_bar(buf));  // accumulator complete
      }
      RFK_STAMP(4);  // all MMAs issued
      RFK_PUT(9, c_wait_full);
      RFK_PUT(10, c_wait_tempty);
      RFK_PUT(13, c_issue);
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
    TileCtx t;
    t.q = warp & 3;
    t.half = (warp - kEpiWarp0) >> 2;
    t.n_tile = n_tile;
    t.stg = g.use_stg ? stg_base + t.half * STG_BYTES : 0u;
    t.flag = reinterpret_cast<volatile unsigned int*>(tmem_slot) + 2;
    const int row = t.q * 32 + lane;
    const int ppi_log2 = g.tw_log2 + g.th_log2;
    uint32_t tl = 0;
    long long c_wait_tfull = 0, c_epi = 0;
    long long phase_cnt[5] = {0, 0, 0, 0, 0};
    t.dbg = g.timeline ? phase_cnt : nullptr;
    t.remote_release = kPair && rank != 0;
    if constexpr (std::is_same<Epi, ActBwdEpi>::value) {
      // ---- fused data gradient + activation backward: the warp's 32 x 64 slices of h arrive by TMA, double-buffered ----
      const int ew = warp - kEpiWarp0;
      const int nblk = g.BN >> 6;
      const int nb = t.half < nblk ? (nblk - t.half + 1) >> 1 : 0;          // 64-column blocks per tile owned by this warp
      const int r0 = t.q * 32, r = r0 + lane, rsw = r & 7;
      float csum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      auto sub_origin = [&](int x0, int y0, int n0, int& sx, int& sy, int& sn) {
        sx = x0 + (r0 & ((1 << g.tw_log2) - 1));
        sy = y0 + ((r0 >> g.tw_log2) & ((1 << g.th_log2) - 1));
        sn = n0 + (r0 >> (g.tw_log2 + g.th_log2));
      };
      const uint32_t nbuf = (uint32_t)g.use_stg >> 1;         // staging buffers per warp: 3 (6 blocks) or 2 (4 blocks, when shared memory is short)
      auto slice = [&](uint32_t b) { return stg_base + ((uint32_t)t.half * nbuf + b) * STG_BYTES + (uint32_t)r0 * 128u; };
      auto issue_load = [&](int it2, int j, uint32_t b) {     // lane 0: block j of tile it2 -> buffer b
        int x0, y0, n0, sx, sy, sn;
        tile_origin(tile_of(it2), x0, y0, n0);
        sub_origin(x0, y0, n0, sx, sy, sn);
        // the TMA store that last read buffer b was issued nbuf - 1 blocks ago: with three buffers the newest store may
        // still be pending, with two it has to be waited for
        if (nbuf == 3) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else bulk_wait_read0();
        mbar_expect_tx(h_bar(ew, b), 32u * 128u);
        tma_load_4d(slice(b), &tmH, h_bar(ew, b), n_tile * g.BN + (t.half + 2 * j) * 64, sx, sy, sn);
      };
      auto prefetch_tile = [&](int it2) {                     // lane 0: this warp's slices of tile it2 -> L2, a whole tile ahead
        int x0, y0, n0, sx, sy, sn;
        tile_origin(tile_of(it2), x0, y0, n0);
        sub_origin(x0, y0, n0, sx, sy, sn);
        for (int j = 0; j < nb; ++j) tma_prefetch_4d(&tmH, n_tile * g.BN + (t.half + 2 * j) * 64, sx, sy, sn);
      };
      uint32_t gb = 0;                                        // blocks processed so far: buffer = gb % nbuf, phase = (gb / nbuf) & 1
      if (nb > 0 && it0 < it_end && lane == 0) issue_load(it0, 0, 0);
      for (int it = it0; it < it_end; it += it_step, ++tl) {
        const int mt = tile_of(it);
        const uint32_t buf = tl & 1u;
        tile_origin(mt, t.x0, t.y0, t.n0);
        t.b = t.n0 + (row >> ppi_log2);
        t.y = t.y0 + ((row >> g.tw_log2) & ((1 << g.th_log2) - 1));
        t.x = t.x0 + (row & ((1 << g.tw_log2) - 1));
        t.valid = mt < g.m_tiles && t.b < g.B && t.y < g.H && t.x < g.W;
        t.tmem_empty_bar = tmem_empty_bar(buf);
        if (lane == 0 && it + it_step < it_end) prefetch_tile(it + it_step);
        mbar_wait(tmem_full_bar(buf), (tl >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(t.q * 32) << 16) + buf * g.BN;
        int sx, sy, sn;
        sub_origin(t.x0, t.y0, t.n0, sx, sy, sn);
        for (int j = 0; j < nb; ++j, ++gb) {
          const uint32_t b = gb % nbuf, bn = (gb + 1u) % nbuf;
          if (lane == 0) {                                    // next block (of this tile or the first of the next) -> next buffer
            if (j + 1 < nb) issue_load(it, j + 1, bn);
            else if (it + it_step < it_end) issue_load(it + it_step, 0, bn);
          }
          mbar_wait(h_bar(ew, b), (gb / nbuf) & 1u);
          float cs[2];
          const int blk = t.half + 2 * j;
          actbwd_block(g, ep, taddr, ss, blk * 64, slice(b) + (uint32_t)lane * 128u, rsw, t.valid, lane, cs);
          csum[(j & 1) * 2] += cs[0];
          csum[(j & 1) * 2 + 1] += cs[1];
          if (j == nb - 1) release_accumulator(t);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmO, slice(b), n_tile * g.BN + blk * 64, sx, sy, sn);
            bulk_commit();
          }
        }
        if (nb == 0) release_accumulator(t);
      }
      for (int j = 0; j < nb; ++j) {
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const int col = n_tile * g.BN + (t.half + 2 * j) * 64 + pass * 32 + colsum_col(lane);
          if (col < g.n) atomicAdd(ep.colsum + col, csum[(j & 1) * 2 + pass]);
        }
      }
    } else {
      for (int it = it0; it < it_end; it += it_step, ++tl) {
        const int mt = tile_of(it);
        const uint32_t buf = tl & 1u;
        tile_origin(mt, t.x0, t.y0, t.n0);
        t.b = t.n0 + (row >> ppi_log2);
        t.y = t.y0 + ((row >> g.tw_log2) & ((1 << g.th_log2) - 1));
        t.x = t.x0 + (row & ((1 << g.tw_log2) - 1));
        t.valid = mt < g.m_tiles && t.b < g.B && t.y < g.H && t.x < g.W;
        t.tmem_empty_bar = tmem_empty_bar(buf);
        t.tile_id = n_tile * g.m_tiles + mt;
        { CNT_BEGIN(); mbar_wait(tmem_full_bar(buf), (tl >> 1) & 1u); CNT_END(c_wait_tfull); }
        if (tl == 0 && warp == kEpiWarp0 && lane == 0) RFK_STAMP(5);  // first accumulator ready
        tc_fence_after();
        { CNT_BEGIN(); epilogue(g, ep, &tmO, tmem_base + ((uint32_t)(t.q * 32) << 16) + buf * g.BN, ss, t); CNT_END(c_epi); }
        if (tl == 0 && warp == kEpiWarp0 && lane == 0) RFK_STAMP(6);  // first epilogue done
      }
    }
    if (g.use_stg && lane == 0) bulk_wait0();  // this warp's outstanding TMA stores still read shared memory
    if (warp == kEpiWarp0 && lane == 0) { RFK_STAMP(7); RFK_PUT(11, c_wait_tfull); RFK_PUT(12, c_epi);
      RFK_PUT(14, phase_cnt[0]); RFK_PUT(15, phase_cnt[1]); RFK_PUT(2, phase_cnt[2]); RFK_PUT(5, phase_cnt[3]); RFK_PUT(6, phase_cnt[4]); }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // the peer may still be arriving on this CTA's barriers / reading its weights
  if (warp == 2) {
    if (kPair)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols)
                   : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
EncodeTiledFn encode_fn() {
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

int ilog2_ceil(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

static unsigned long long* g_timeline = nullptr;
static long long g_timeline_cap = 0;
static int g_conv_split = 0;   // rfk_set_conv_split: bf16x3 split-precision operands (process-global, set once at start-up)
int conv_split_mode() { return g_conv_split; }
unsigned long long* debug_timeline(long long* capacity_ctas) { *capacity_ctas = g_timeline_cap; return g_timeline; }

struct Plan {
  GemmArgs g;
  CUtensorMap tmA, tmB, tmO, tmH;
  dim3 grid;
  size_t smem;
  int TW, TH, NIMG;
};

int encode_weight_map(CUtensorMap* map, const char* who, const void* ptr, long long ktot, int rows, int box_rows, int bk) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("%s: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)", who);
    return RFK_ECUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  cuuint32_t ones[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled(weights) failed with CUresult %d (rows=%d ktot=%lld box_rows=%d)", who, (int)r, rows,
              ktot, box_rows);
    return RFK_ECUDA;
  }
  return RFK_OK;
}

int encode_act_map(CUtensorMap* map, const char* who, const char* what, const void* ptr, int channels, int ld,
                   int B, int H, int W, int TW, int TH, int NIMG, int bk) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("%s: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)", who);
    return RFK_ECUDA;
  }
  // NHWC bf16 viewed as 4-D {C, W, H, B}; box {64, TW, TH, NIMG}; loads: OOB -> zeros (the conv padding); stores: clipped
  cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)NIMG};
  cuuint32_t ones[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled(%s) failed with CUresult %d (B=%d H=%d W=%d ld=%d channels=%d)", who, what, (int)r,
              B, H, W, ld, channels);
    return RFK_ECUDA;
  }
  return RFK_OK;
}

// stg_wanted: the epilogue can use TMA stores (needs 2 x 16 KB of staging shared memory)
static int make_plan(Plan& p, const char* who, const void* act, int B, int H, int W, int act_ld, int cin_pad,
                     const void* wgt, int n, int n_pad, int taps, int BN, bool stg_wanted, int k_split = 1,
                     bool allow_pair = false, int stg_blocks = 2) {
  RFK_REQUIRE(act && wgt && B > 0 && H > 0 && W > 0, "%s: null pointer or empty shape", who);
  const int parts_ld = g_conv_split ? 2 : 1, parts_k = g_conv_split ? 3 : 1;
  RFK_REQUIRE(cin_pad > 0 && (cin_pad % 64 == 0 || cin_pad == 32) && parts_ld * cin_pad <= act_ld,
              "%s: cin_pad=%d must be 32 or a multiple of 64, and %d x cin_pad <= act_ld=%d", who, cin_pad, parts_ld, act_ld);
  const int bk = cin_pad % 64 == 0 ? 64 : 32;
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
  g.B = B; g.H = H; g.W = W; g.n = n; g.BN = BN; g.taps = taps; g.bk = bk; g.kchunks = parts_k * cin_pad / bk;
  g.kch = cin_pad / bk;
  g.lo_coord = act_ld / 2;
  if (g_conv_split) RFK_REQUIRE(act_ld % 16 == 0 && cin_pad <= act_ld / 2, "%s: split precision needs rows of [hi | lo] halves (act_ld=%d, cin_pad=%d)", who, act_ld, cin_pad);
  int twl = ilog2_ceil(W);
  if (twl > 7) twl = 7;
  int thl = ilog2_ceil(H);
  if (thl > 7 - twl) thl = 7 - twl;
  g.tw_log2 = twl; g.th_log2 = thl;
  p.TW = 1 << twl; p.TH = 1 << thl; p.NIMG = BM / (p.TW * p.TH);
  g.tiles_x = ceil_div(W, p.TW);
  g.tiles_y = ceil_div(H, p.TH);
  g.m_tiles = g.tiles_x * g.tiles_y * ceil_div(B, p.NIMG);
  const int n_tiles = n_pad / BN;

  // shared-memory budget: [resident weights] [stages] [2 staging blocks] [scale/shift] [barriers]
  const int k_iters = taps * g.kchunks;
  // 64-byte-row chunks carry only two MMAs each: group three (a filter row) or two per pipeline stage
  g.kgroup = bk == 32 ? (k_iters % 3 == 0 ? 3 : (k_iters % 2 == 0 ? 2 : 1)) : 1;
  RFK_REQUIRE(k_split >= 1 && (k_iters / g.kgroup) % k_split == 0, "%s: %d K stages do not split %d ways", who,
              k_iters / g.kgroup, k_split);
  g.kg_per_split = k_iters / g.kgroup / k_split;
  const int b_chunk = g.kgroup * BN * bk * 2;
  const int a_stage = g.kgroup * BM * bk * 2;
  const int kHBars = 8 * 3 * kEpiWarps;   // ActBwdEpi's per-warp h-tile barriers (allocated for every epilogue: 192 B)
  const int fixed = 1024 /*alignment slack*/ + (stg_wanted ? stg_blocks * STG_BYTES : 0) + 2 * BN * 4 + 8 * (2 * 8 + 5) + 16 + kHBars;
  g.use_stg = stg_wanted ? stg_blocks : 0;
  int resident = 0, stages = 0;
  g.pair = 0;
  {
    // CTA pairs: when the resident weights leave room for only a few activation stages (the MMA warp then waits ~2k cycles
    // per tile for TMA data), two CTAs of a cluster share ONE copy of the weights -- half the rows each -- and issue
    // M = 256 MMAs (tcgen05 cta_group::2): twice the pipeline depth at the same shared-memory size.
    static const int pair_mode = [] { const char* e = getenv("RFK_GEMM_PAIR"); return e ? atoi(e) : 1; }();
    const long long full_res = (long long)k_iters * BN * bk * 2;
    static const int pair_min_stages = [] { const char* e = getenv("RFK_GEMM_PAIR_MIN_STAGES"); return e ? atoi(e) : 3; }();
    if (allow_pair && pair_mode && stg_wanted && k_split == 1 && n_tiles == 1 && BN % 32 == 0 &&
        full_res >= (pair_mode == 2 ? 0 : 96 * 1024) && g.m_tiles >= 2 * sm_count() &&
        (long long)SMEM_LIMIT - fixed - full_res / 2 >= (long long)pair_min_stages * a_stage)
      g.pair = 1;
  }
  {
    const long long res_bytes = (long long)(k_iters / k_split) * BN * bk * 2 / (g.pair ? 2 : 1);
    const long long room = (long long)SMEM_LIMIT - fixed - res_bytes;
    if (room >= 3LL * a_stage || g.pair) {
      resident = 1;
      stages = (int)(room / a_stage);
    } else {
      stages = (SMEM_LIMIT - fixed) / (a_stage + b_chunk);
    }
  }
  // a CTA that only ever sees one pixel tile gains nothing from resident weights: it would wait for ALL of them
  // before its first MMA, whereas streamed weight chunks arrive stage by stage (latency-bound small levels / sampling)
  if (resident && g.m_tiles * n_tiles * k_split <= sm_count()) {
    resident = 0;
    stages = (SMEM_LIMIT - fixed) / (a_stage + b_chunk);
  }
  if (const char* s = getenv("RFK_GEMM_RESIDENT")) {
    if (atoi(s) == 0 && resident) {
      resident = 0;
      stages = (SMEM_LIMIT - fixed) / (a_stage + b_chunk);
    }
  }
  if (const char* s = getenv("RFK_GEMM_STAGES")) stages = std::min(stages, std::max(1, atoi(s)));
  if (stages > 8) stages = 8;
  RFK_REQUIRE(stages >= 1, "%s: tile does not fit in shared memory (BN=%d)", who, BN);
  g.stages = stages;
  g.b_resident = resident;
  int cols = 32;
  while (cols < 2 * BN) cols <<= 1;
  g.tmem_cols = cols;  // two accumulators
  const int stage_bytes = a_stage + (resident ? 0 : b_chunk);
  p.smem = (size_t)1024 + (resident ? (size_t)(k_iters / k_split) * BN * bk * 2 / (g.pair ? 2 : 1) : 0) + (size_t)stages * stage_bytes +
           (stg_wanted ? stg_blocks * STG_BYTES : 0) + 2 * BN * 4 + 8 * (2 * stages + 5) + 16 + kHBars;
  RFK_REQUIRE(p.smem <= (size_t)SMEM_LIMIT, "%s: internal error: %zu B of shared memory planned", who, p.smem);
  int ctas_x = sm_count() / (n_tiles * k_split);
  if (ctas_x < 1) ctas_x = 1;
  if (ctas_x > g.m_tiles) ctas_x = g.m_tiles;
  if (g.pair) {
    RFK_REQUIRE(resident, "%s: internal error: CTA pairs need resident weights", who);
    ctas_x &= ~1;   // whole pairs
  }
  p.grid = dim3((unsigned)ctas_x, (unsigned)n_tiles, (unsigned)k_split);
  g.timeline = (g_timeline && (long long)ctas_x * n_tiles * k_split <= g_timeline_cap) ? g_timeline : nullptr;
  g.scale = nullptr; g.shift = nullptr; g.n_ss = 0;

  int rc = encode_act_map(&p.tmA, who, "A", act, g_conv_split ? act_ld : cin_pad, act_ld, B, H, W, p.TW, p.TH, p.NIMG, bk);
  if (rc) return rc;
  p.tmO = p.tmA;  // placeholder unless the epilogue stores through TMA
  p.tmH = p.tmA;  // placeholder unless the epilogue loads a saved activation (ActBwdEpi)
  // B: weights [n_pad, taps*cin_pad] viewed as 2-D {K, N}; box {64, BN}
  const cuuint64_t ktot = (cuuint64_t)taps * parts_k * cin_pad;
  cuuint64_t dimsB[2] = {ktot, (cuuint64_t)n_pad};
  cuuint64_t strB[1] = {ktot * 2};
  cuuint32_t boxB[2] = {(cuuint32_t)bk, (cuuint32_t)(g.pair ? BN / 2 : BN)};
  cuuint32_t ones[2] = {1, 1};
  CUresult r = enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wgt), dimsB, strB, boxB, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled(B) failed with CUresult %d (n_pad=%d ktot=%llu BN=%d)", who, (int)r, n_pad,
              (unsigned long long)ktot, BN);
    return RFK_ECUDA;
  }
  return RFK_OK;
}

template <class T> struct PairCapable { static constexpr bool value = false; };
template <int ACT> struct PairCapable<PlainEpi<ACT>> { static constexpr bool value = true; };
template <> struct PairCapable<ActBwdEpi> { static constexpr bool value = true; };

template <class Epi, bool kPair>
static int launch_impl(const Plan& p, const Epi& ep, cudaStream_t st, const char* who) {
  static size_t configured = 0;
  if (p.smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<Epi, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", who, p.smem, cudaGetErrorString(e));
      return RFK_ECUDA;
    }
    configured = p.smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = p.grid;
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (kPair) {   // two CTAs along x = the two SMs of a TPC
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaLaunchKernelEx(&cfg, conv_gemm_kernel<Epi, kPair>, p.tmA, p.tmB, p.tmO, p.tmH, p.g, ep);
  return check_launch(who);
}

template <class Epi>
static int launch(const Plan& p, const Epi& ep, cudaStream_t st, const char* who) {
  if constexpr (PairCapable<Epi>::value) {
    if (p.g.pair) return launch_impl<Epi, true>(p, ep, st, who);
  }
  return launch_impl<Epi, false>(p, ep, st, who);
}

// N tile: the whole (padded) channel count when it fits one accumulator; otherwise the largest divisor that is a
// multiple of `quantum`.  Small pixel counts get narrower tiles so that more SMs share the work.
static int pick_bn(int n_pad, int quantum, int m_tiles_hint) {
  int best = 0;
  int bn_max = 256;
  if (const char* s = getenv("RFK_GEMM_BN_MAX")) bn_max = std::max(quantum, atoi(s) / quantum * quantum);
  for (int bn = bn_max; bn >= quantum; bn -= quantum)
    if (n_pad % bn == 0) { best = bn; break; }
  if (!best) return 0;
  const int sms = sm_count();
  while (best % 2 == 0 && (best / 2) % quantum == 0 && best / 2 >= 64 && m_tiles_hint * (n_pad / best) * 2 <= sms) best /= 2;
  return best;
}

static int m_tiles_of(int B, int H, int W) {
  int twl = ilog2_ceil(W);
  if (twl > 7) twl = 7;
  int thl = ilog2_ceil(H);
  if (thl > 7 - twl) thl = 7 - twl;
  const int TW = 1 << twl, TH = 1 << thl;
  return ceil_div(W, TW) * ceil_div(H, TH) * ceil_div(B, BM / (TW * TH));
}

}  // namespace rfk

using namespace rfk;

extern "C" int rfk_set_conv_split(int on) {
  g_conv_split = on ? 1 : 0;
  return RFK_OK;
}

extern "C" int rfk_debug_set_timeline(unsigned long long* buf, long long capacity_ctas) {
  g_timeline = buf;
  g_timeline_cap = buf ? capacity_ctas : 0;
  return RFK_OK;
}

extern "C" int rfk_conv_gemm(const void* act, int B, int H, int W, int act_ld, int cin_pad, const void* wgt, int n,
                             int n_pad, int taps, const float* scale, const float* shift, int act_fn, int out_kind,
                             void* out, int out_ld, int out_off, void* stream) {
  RFK_REQUIRE(out, "rfk_conv_gemm: null output");
  RFK_REQUIRE(out_kind == RFK_OUT_NHWC_BF16 || out_kind == RFK_OUT_NCHW_F32, "rfk_conv_gemm: bad out_kind %d", out_kind);
  RFK_REQUIRE(act_fn >= 0 && act_fn <= 2, "rfk_conv_gemm: bad act_fn %d", act_fn);
  RFK_REQUIRE(n_pad > 0 && n_pad % 16 == 0, "rfk_conv_gemm: n_pad=%d must be a positive multiple of 16", n_pad);
  PlainEpi<0> e;
  e.act_fn = act_fn; e.out_kind = out_kind; e.out = out; e.out_ld = out_ld; e.out_off = out_off; e.vec_ok = 0;
  e.tma_store = 0; e.lo_off = 0;
  bool tma_ok = false;
  if (out_kind == RFK_OUT_NHWC_BF16) {
    const int span = g_conv_split ? out_ld / 2 : out_ld;   // split precision: [hi | lo] halves of the row
    RFK_REQUIRE(out_off >= 0 && out_off + n <= span, "rfk_conv_gemm: output window [%d,%d) exceeds %d channels", out_off,
                out_off + n, span);
    e.vec_ok = out_ld % 8 == 0 && out_off % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    tma_ok = e.vec_ok && getenv("RFK_GEMM_NO_TMA_STORE") == nullptr && !g_conv_split;
    if (g_conv_split) e.lo_off = span;
  }
  const int mth = m_tiles_of(B, H, W);
  int BN = 0;
  if (tma_ok) {
    // TMA stores move 64-channel blocks: a tile that is not the only one must be a whole number of blocks
    BN = (n_pad <= 256 && n_pad % 64 != 0) ? n_pad : pick_bn(n_pad, 64, mth);
    if (!BN) tma_ok = false;
  }
  if (!BN) BN = pick_bn(n_pad, 16, mth);
  Plan p;
  int rc = make_plan(p, "rfk_conv_gemm", act, B, H, W, act_ld, cin_pad, wgt, n, n_pad, taps, BN, tma_ok, 1, tma_ok);
  if (rc) return rc;
  p.g.scale = scale; p.g.shift = shift; p.g.n_ss = n;
  if (tma_ok) {
    e.tma_store = 1;
    // the store map's box is ONE EPILOGUE WARP's 32 pixel rows of the tile (x fastest, then y, then image): every warp
    // stores its own sub-box, so the epilogue needs no cross-warp barrier
    const int sx = std::min(p.TW, 32), sy = std::min(p.TH, 32 / sx), sn = 32 / (sx * sy);
    rc = encode_act_map(&p.tmO, "rfk_conv_gemm", "out", reinterpret_cast<const __nv_bfloat16*>(out) + out_off, n, out_ld, B, H,
                        W, sx, sy, sn);
    if (rc) return rc;
  }
  if (act_fn == RFK_ACT_RELU) {
    PlainEpi<RFK_ACT_RELU> e1;
    e1.act_fn = act_fn; e1.out_kind = e.out_kind; e1.out = e.out; e1.out_ld = e.out_ld; e1.out_off = e.out_off;
    e1.vec_ok = e.vec_ok; e1.tma_store = e.tma_store; e1.lo_off = e.lo_off;
    return launch(p, e1, (cudaStream_t)stream, "rfk_conv_gemm");
  }
  if (act_fn == RFK_ACT_LEAKY) {
    PlainEpi<RFK_ACT_LEAKY> e2;
    e2.act_fn = act_fn; e2.out_kind = e.out_kind; e2.out = e.out; e2.out_ld = e.out_ld; e2.out_off = e.out_off;
    e2.vec_ok = e.vec_ok; e2.tma_store = e.tma_store; e2.lo_off = e.lo_off;
    return launch(p, e2, (cudaStream_t)stream, "rfk_conv_gemm");
  }
  return launch(p, e, (cudaStream_t)stream, "rfk_conv_gemm");
}

extern "C" int rfk_conv_gemm_actbwd(const void* act, int B, int H, int W, int act_ld, int cin_pad, const void* wgt, int n,
                                    int n_pad, int taps, const float* scale, int act_fn, const void* h, int h_ld, void* out,
                                    int out_ld, float* colsum, void* stream) {
  RFK_REQUIRE(out && h && colsum, "rfk_conv_gemm_actbwd: null pointer");
  RFK_REQUIRE(act_fn >= 0 && act_fn <= 2, "rfk_conv_gemm_actbwd: bad act_fn %d", act_fn);
  RFK_REQUIRE(!g_conv_split, "rfk_conv_gemm_actbwd: training kernels do not run in split-precision mode");
  RFK_REQUIRE(n == n_pad && n % 64 == 0 && n <= 512, "rfk_conv_gemm_actbwd: n=%d must be a multiple of 64 (<= 512) without padding", n);
  RFK_REQUIRE(out_ld % 8 == 0 && n <= out_ld && h_ld % 8 == 0 && n <= h_ld &&
              ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(h)) & 15) == 0,
              "rfk_conv_gemm_actbwd: out / h must be 16-byte aligned NHWC bf16 with row strides that are multiples of 8");
  const int BN = n <= 256 ? n : pick_bn(n_pad, 64, m_tiles_of(B, H, W));
  RFK_REQUIRE(BN > 0 && BN % 64 == 0 && BN <= 256, "rfk_conv_gemm_actbwd: no N tile for n=%d", n);
  Plan p;
  // three staging buffers per epilogue warp (96 KB) unless that starves the main loop of pipeline stages or of the CTA-pair
  // mode: then two (64 KB)
  int rc = make_plan(p, "rfk_conv_gemm_actbwd", act, B, H, W, act_ld, cin_pad, wgt, n, n_pad, taps, BN, true, 1, true, 6);
  if (rc || (!p.g.pair && (p.g.stages < 3 || !p.g.b_resident))) {
    Plan p4;
    const int rc4 = make_plan(p4, "rfk_conv_gemm_actbwd", act, B, H, W, act_ld, cin_pad, wgt, n, n_pad, taps, BN, true, 1, true, 4);
    if (rc4 == 0 && (rc || p4.g.pair || p4.g.stages > p.g.stages || (p4.g.b_resident && !p.g.b_resident))) { p = p4; rc = 0; }
  }
  if (rc) return rc;
  p.g.scale = scale; p.g.shift = nullptr; p.g.n_ss = n;
  const int sx = std::min(p.TW, 32), sy = std::min(p.TH, 32 / sx), sn = 32 / (sx * sy);
  rc = encode_act_map(&p.tmO, "rfk_conv_gemm_actbwd", "out", out, n, out_ld, B, H, W, sx, sy, sn);
  if (rc) return rc;
  rc = encode_act_map(&p.tmH, "rfk_conv_gemm_actbwd", "h", h, n, h_ld, B, H, W, sx, sy, sn);   // same 32-row slices, loaded
  if (rc) return rc;
  ActBwdEpi e;
  e.h = (const __nv_bfloat16*)h; e.h_ld = h_ld; e.act_fn = act_fn; e.colsum = colsum;
  return launch(p, e, (cudaStream_t)stream, "rfk_conv_gemm_actbwd");
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
  int rc = make_plan(p, "rfk_conv_gemm_coupling", act, B, H, W, act_ld, cin_pad, wgt, n, n_pad, taps, n_pad, false);
  if (rc) return rc;
  p.g.scale = scale; p.g.shift = shift; p.g.n_ss = n;
  CouplingEpi e;
  e.z = z; e.clamp_type = clamp_type; e.cs = clamp_scale; e.csh = clamp_shift; e.logdet = logdet; e.reverse = reverse;
  return launch(p, e, (cudaStream_t)stream, "rfk_conv_gemm_coupling");
}

extern "C" int rfk_conv_gemm_splitk_fused(const void* act, int B, int H, int W, int act_ld, int cin_pad, const void* wgt,
                                          int n, int n_pad, int taps, int k_split, float* ws, int ws_ld,
                                          unsigned int* counters, const float* scale, const float* shift, int act_fn,
                                          void* out, int out_ld, int out_off, void* stream) {
  RFK_REQUIRE(ws && counters && out && ws_ld >= n_pad && ws_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0,
              "rfk_conv_gemm_splitk_fused: null pointer or bad workspace (16-byte aligned, ws_ld >= n_pad, ws_ld %% 4 == 0)");
  RFK_REQUIRE(n_pad > 0 && n_pad % 16 == 0 && act_fn >= 0 && act_fn <= 2, "rfk_conv_gemm_splitk_fused: bad n_pad / act_fn");
  RFK_REQUIRE(!g_conv_split, "rfk_conv_gemm_splitk_fused: not available in split-precision mode");
  RFK_REQUIRE(out_off >= 0 && out_off + n <= out_ld, "rfk_conv_gemm_splitk_fused: output window exceeds out_ld");
  Plan p;
  int rc = make_plan(p, "rfk_conv_gemm_splitk_fused", act, B, H, W, act_ld, cin_pad, wgt, n, n_pad, taps,
                     pick_bn(n_pad, 16, 1 << 20), false, k_split);
  if (rc) return rc;
  p.g.scale = scale; p.g.shift = shift; p.g.n_ss = n;
  RFK_REQUIRE(k_split >= 1 && k_split <= 9, "rfk_conv_gemm_splitk_fused: k_split=%d (1..9)", k_split);
  RFK_REQUIRE((long long)p.grid.x * p.grid.y * p.grid.z <= sm_count(),
              "rfk_conv_gemm_splitk_fused: %u x %u x %u CTAs exceed the %d SMs (the slices of a tile wait for each other, so "
              "all CTAs must be co-resident)", p.grid.x, p.grid.y, p.grid.z, sm_count());
  SplitKEpi e;
  e.ws = ws; e.ld = ws_ld; e.counters = counters; e.k_split = k_split; e.act_fn = act_fn;
  e.slice_stride = (long long)p.g.m_tiles * 128 * n_pad;   // one slab per K slice: [tile][16-col group][128 rows][16]
  RFK_REQUIRE(ws_ld == n_pad, "rfk_conv_gemm_splitk_fused: ws_ld must equal n_pad (%d)", n_pad);
  e.out = (__nv_bfloat16*)out; e.out_ld = out_ld; e.out_off = out_off;
  e.vec_ok = out_ld % 8 == 0 && out_off % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  return launch(p, e, (cudaStream_t)stream, "rfk_conv_gemm_splitk_fused");
}

extern "C" int rfk_conv_gemm_splitk(const void* act, int B, int H, int W, int act_ld, int cin_pad, const void* wgt, int n,
                                    int n_pad, int taps, int k_split, float* ws, int ws_ld, void* stream) {
  RFK_REQUIRE(ws && ws_ld >= n_pad && ws_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0,
              "rfk_conv_gemm_splitk: workspace must be 16-byte aligned with ws_ld >= n_pad, ws_ld %% 4 == 0");
  RFK_REQUIRE(n_pad > 0 && n_pad % 16 == 0, "rfk_conv_gemm_splitk: n_pad=%d must be a positive multiple of 16", n_pad);
  Plan p;
  int rc = make_plan(p, "rfk_conv_gemm_splitk", act, B, H, W, act_ld, cin_pad, wgt, n, n_pad, taps, pick_bn(n_pad, 16, 1 << 20),
                     false, k_split);
  if (rc) return rc;
  SplitKEpi e;
  e.ws = ws; e.ld = ws_ld; e.counters = nullptr; e.slice_stride = 0; e.k_split = k_split; e.act_fn = 0; e.out = nullptr;
  e.out_ld = 0; e.out_off = 0; e.vec_ok = 0;
  return launch(p, e, (cudaStream_t)stream, "rfk_conv_gemm_splitk");
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
  int rc = make_plan(p, "rfk_conv_gemm_lstm", act, B, H, W, act_ld, cin_pad, wgt, n_pad, n_pad, taps, BN, false);
  if (rc) return rc;
  p.g.scale = nullptr; p.g.shift = bias; p.g.n_ss = n_pad;
  LstmEpi e;
  e.hidden = hidden; e.ht = ht; e.ht_pad = ht_pad; e.c_prev = c_prev; e.c_prev_bs = c_prev_bstride;
  e.peep = peep; e.c_next = c_next; e.c_next_bs = c_next_bstride; e.h_out = h_out; e.h_bs = h_bstride;
  e.h_nhwc = (__nv_bfloat16*)h_nhwc; e.h_off = h_off; e.h_ld = h_ld;
  e.h_vec_ok = h_nhwc && h_ld % 8 == 0 && h_off % 8 == 0 && ht % 8 == 0 && (reinterpret_cast<uintptr_t>(h_nhwc) & 15) == 0;
  e.h_lo_off = (h_nhwc && g_conv_split) ? h_ld / 2 : 0;
  if (h_nhwc) RFK_REQUIRE(h_off >= 0 && h_off + hidden <= (g_conv_split ? h_ld / 2 : h_ld), "rfk_conv_gemm_lstm: h window exceeds h_ld");
  return launch(p, e, (cudaStream_t)stream, "rfk_conv_gemm_lstm");
}
