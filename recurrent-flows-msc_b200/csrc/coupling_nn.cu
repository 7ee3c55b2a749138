// The whole coupling network of a GlowStep in ONE kernel (sm_100a, CTA pairs):
//
//   h1   = act(ActNorm(conv3x3(nn_in)))     GEMM1: implicit GEMM over the filter taps, A by TMA, W1 in shared memory
//   h2   = act(ActNorm(conv1x1(h1)))        GEMM2: A = bf16 h1 in TENSOR MEMORY (TS form), W2 in shared memory
//   taps = tap-split form of conv3x3(h2)    GEMM3: A = bf16 h2 in tensor memory, W9 in shared memory
//
// (Flow/glow_modules.py:229-240: net.0 = Conv2dNorm 3x3, net.1 = act, net.2 = Conv2dNorm 1x1, net.3 = act,
// net.4 = Conv2dZeros 3x3; the tap-split form of net.4 is described at rfk_coupling_tail_taps in rfk.h.)
//
// Neither 256-channel hidden tensor touches HBM: 2 x 299 MB written + 2 x 299 MB read per GlowStep at level 1 of the
// 570-frame workload become 37 MB in + 84 MB out (ncu: 37.7 MB read, 29.1 MB written to DRAM -- most of the tap planes stay
// in L2).  (Training keeps h1 / h2 for the backward: they leave as side outputs by TMA store, written once and never read
// back by this pass.)
//
// W1 (up to 147 KB) + W2 (128 KB) + W9 do not fit one CTA's shared memory, so the kernel runs as CTA PAIRS
// (cluster of two, tcgen05 cta_group::2): every MMA is M = 256 (128 pixels per CTA), each CTA holds HALF of the rows of
// every weight matrix.  When W1 still does not fit (3x3 convs with more than 32 input channels) its chunks stream with
// the activations.
//
// The two ActNorms cost nothing at run time: their per-channel scale is folded into the weight rows when the weights are
// packed (rfk_pack_weight_folded) and their shift rides through the GEMM as 16 extra K columns (t_hi, t_lo, 0 ...) that
// meet a CONSTANT-ONE A operand -- a 4 KB shared-memory tile for GEMM1, eight tensor-memory columns for GEMM2.  The
// activation epilogues are tcgen05.ld -> activation -> bf16 -> tcgen05.st and nothing else (a per-channel affine read from
// shared memory made them LSU-bound: 16 warps x 8 LDS.128 per 16 columns, ~590 cycles per iteration, which paced GEMM2 / 3).
//
// Tensor memory (512 columns, two 256-column regions whose roles alternate from tile to tile):
//   R0: GEMM1 accumulator (fp32) -> overwritten IN PLACE by bf16 h1 (16 fp32 columns become 8 packed columns, each epilogue
//       warp rewrites only columns it has itself read; the constant-one columns sit in the first gap) -> after GEMM2: bf16 h2
//       (compact, columns [0, hid/2)) and the GEMM3 accumulator (columns [128, 128 + min(n3_pad, 128)); more than 128 tap
//       planes run as two passes over it)
//   R1: GEMM2 accumulator (fp32); it is the next tile's R0.
// GEMM1 of tile t+1 is issued right behind GEMM3 of tile t, so the tensor pipe idles only while the first chunk of an
// activation epilogue is being produced.  All hand-offs are mbarriers; barriers that gate MMAs live in the leader CTA (the
// peer arrives remotely), completion barriers are multicast to both CTAs by tcgen05.commit.
#include <algorithm>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rfk {

constexpr int kNNThreads = 640;   // warp 0: TMA, warp 1: MMA (leader CTA only), warp 2: TMEM alloc, warps 4-19: epilogue
constexpr int kNNEpiWarp0 = 4;
constexpr int kNNEpiWarps = 16;
constexpr int kNNSmemLimit = 232448;
constexpr int kNNStoreSlice = 32 * 32;   // staging of the side outputs per epilogue warp: four warps share a 32 pixels x 64 channels block

struct NNArgs {
  int B, H, W;
  int tw_log2, th_log2, tiles_x, tiles_y, m_tiles;
  int taps, kchunks, bk, kgroup, k_groups;   // GEMM1: taps * kchunks chunks of bk channels, kgroup chunks per pipeline stage
  int w1_resident;
  int hid;                                   // N of GEMM1 and GEMM2, K of GEMM2 and GEMM3 (multiple of 64, <= 256)
  int n3, n3_pad;                            // tap planes 9*C, padded to a multiple of 16 (<= 256)
  int stages;
  int act_fn;
  int store_h;                               // 1: h1 / h2 leave as bf16 NHWC side outputs
  float* taps_out;                           // fp32 NCHW [B, n3, H, W]
  unsigned long long* dbg;                   // debug: 16 cycle counters + 16 stamps per CTA (rfk_debug_set_timeline; tools/nn_fused_timeline.py), else null
};

#define NN_CNT_BEGIN() const unsigned cnt_t0_ = dbg ? (unsigned)clock() : 0u
#define NN_CNT_END(var) do { if (dbg) var += (unsigned)clock() - cnt_t0_; } while (0)

__device__ __forceinline__ void umma_bf16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 16 fp32 accumulator columns of one pixel -> activation -> 8 packed bf16x2 words
__device__ __forceinline__ void act_pack(const uint32_t (&v)[16], int act_fn, uint32_t (&pk)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float a0 = __uint_as_float(v[2 * k]), a1 = __uint_as_float(v[2 * k + 1]);
    if (act_fn == RFK_ACT_LEAKY) { a0 = apply_act(a0, RFK_ACT_LEAKY); a1 = apply_act(a1, RFK_ACT_LEAKY); }
    uint32_t p = pack_bf16(a0, a1);
    if (act_fn == RFK_ACT_RELU) p = relu_bf16x2(p);   // relu(bf16(x)) == bf16(relu(x))
    pk[k] = p;
  }
}

// K-major shared-memory matrix descriptor for ONE K = 16 slice stored as 32-byte rows (SWIZZLE_32B), 8-row groups 256 B apart
__device__ __forceinline__ uint64_t umma_desc_k16(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;   // SWIZZLE_32B
  return d;
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// kStore: h1 / h2 side outputs (keeps the packed words of a whole epilogue in registers); kDbg: cycle counters and stamps
template <bool kStore, bool kDbg>
__global__ void __launch_bounds__(kNNThreads, 1)
coupling_nn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                   const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW9,
                   const __grid_constant__ CUtensorMap tmS1, const __grid_constant__ CUtensorMap tmS2,
                   const __grid_constant__ CUtensorMap tmH1, const __grid_constant__ CUtensorMap tmH2, const NNArgs g) {
  unsigned long long* const dbg = kDbg ? g.dbg : nullptr;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t rank = cluster_ctarank();           // 0 = leader: issues the MMAs, owns the barriers that gate them
  const int hid_local = g.hid >> 1, n3_local = g.n3_pad >> 1;   // weight rows held by this CTA
  const int n3_passes = (g.n3_pad + 127) >> 7;                  // GEMM3 runs in passes of up to 128 accumulator columns
  const int k1_iters = g.taps * g.kchunks, k2chunks = g.hid >> 6;
  const uint32_t a_chunk = 128u * (uint32_t)g.bk * 2u, b1_chunk = (uint32_t)hid_local * (uint32_t)g.bk * 2u;
  const uint32_t w2_chunk = (uint32_t)hid_local * 128u, w9_chunk = (uint32_t)n3_local * 128u;
  // shared memory: [W1 half, if resident] [W2 half] [W9 half] [shift columns of W1 | of W2 | constant-one A tile: 4 KB each]
  //                [stages] [side-output staging] [barriers]
  const uint32_t w1_bytes = g.w1_resident ? (uint32_t)k1_iters * b1_chunk : 0u;
  const uint32_t w2_bytes = (uint32_t)k2chunks * w2_chunk, w9_bytes = (uint32_t)k2chunks * w9_chunk;
  const uint32_t w1_base = base, w2_base = base + w1_bytes, w9_base = w2_base + w2_bytes;
  const uint32_t bias_bytes = (uint32_t)hid_local * 32u;                 // one K = 16 slice of this CTA's weight rows
  const uint32_t bias1_base = w9_base + w9_bytes, bias2_base = bias1_base + 4096u, ones_base = bias2_base + 4096u;
  const uint32_t stage_bytes = (uint32_t)g.kgroup * (a_chunk + (g.w1_resident ? 0u : b1_chunk));
  const uint32_t stage_base = ones_base + 4096u;
  const uint32_t stg_base = stage_base + (uint32_t)g.stages * stage_bytes;
  const uint32_t bar_off = (stg_base - base) + (kStore ? (uint32_t)kNNEpiWarps * kNNStoreSlice : 0u);
  const uint32_t bar_base = base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (g.stages + s); };
  auto acc_full = [&](int i) { return bar_base + 8u * (2 * g.stages + i); };            // GEMM i+1 complete (both CTAs)
  const uint32_t w_full = bar_base + 8u * (2 * g.stages + 3);
  auto a2_ready = [&](int kc) { return bar_base + 8u * (2 * g.stages + 4 + kc); };       // bf16 h1 channels [64kc, 64kc+64) in TMEM
  auto a3_ready = [&](int kc) { return bar_base + 8u * (2 * g.stages + 8 + kc); };       // bf16 h2 ...
  const uint32_t d3_empty = bar_base + 8u * (2 * g.stages + 12);                         // tap accumulator drained
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8u * (2 * g.stages + 13));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmW9);
    tma_prefetch_desc(&tmS1);
    tma_prefetch_desc(&tmS2);
    if (kStore) {
      tma_prefetch_desc(&tmH1);
      tma_prefetch_desc(&tmH2);
    }
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(acc_full(i), 1);
    mbar_init(w_full, 1);
    for (int kc = 0; kc < 4; ++kc) {
      mbar_init(a2_ready(kc), 2 * kNNEpiWarps);    // every epilogue warp of both CTAs writes 16 channels of a 64-channel chunk
      mbar_init(a3_ready(kc), 2 * kNNEpiWarps);
    }
    mbar_init(d3_empty, 2 * kNNEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // constant-one A tile of the shift MMA of GEMM1: 128 rows x 16 K, every 16-byte half row = (1, 1, 0, 0, 0, 0, 0, 0), which
  // is invariant under the 32-byte swizzle
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(ones_base + 16u * (uint32_t)i), "r"(0x3F803F80u), "r"(0u) : "memory");
  fence_async_smem();
  cluster_sync_all();   // the peer's barriers and one-tile exist before anything (TMA, commits, remote arrives, MMAs) targets them
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // the weights are parameters (never written by the preceding kernel): their loads go out before the dependency wait.
  // Each CTA fetches its half of the rows; both halves count on the leader's barrier.
  if (warp == 0 && lane == 0) {
    if (rank == 0) mbar_expect_tx(w_full, 2u * (w1_bytes + w2_bytes + w9_bytes + 2u * bias_bytes));
    const uint32_t wbar = w_full & kPeerBitMask;
    if (g.w1_resident)
      for (int it = 0; it < k1_iters; ++it)
        tma_load_2d_2sm(w1_base + it * b1_chunk, &tmW1, wbar, it * g.bk, (int)rank * hid_local);
    for (int kc = 0; kc < k2chunks; ++kc) tma_load_2d_2sm(w2_base + kc * w2_chunk, &tmW2, wbar, kc * 64, (int)rank * hid_local);
    for (int kc = 0; kc < k2chunks; ++kc) tma_load_2d_2sm(w9_base + kc * w9_chunk, &tmW9, wbar, kc * 64, (int)rank * n3_local);
    tma_load_2d_2sm(bias1_base, &tmS1, wbar, k1_iters * g.bk, (int)rank * hid_local);   // the 16 shift columns behind the weights
    tma_load_2d_2sm(bias2_base, &tmS2, wbar, g.hid, (int)rank * hid_local);
  }
  pdl_trigger();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

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
  // a CTA pair takes tile PAIRS (2*it + rank); with an odd tile count the odd CTA's last tile does not exist (its loads are
  // zero-filled, its stores clipped / skipped)
  const int it0 = (int)(blockIdx.x >> 1), it_step = (int)(gridDim.x >> 1), it_end = (g.m_tiles + 1) >> 1;

  if (warp == 0) {
    // ===== TMA producer: the filter taps of the network input (+ the W1 chunks when they are not resident) =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = it0; it < it_end; it += it_step) {
        int x0, y0, n0;
        tile_origin(2 * it + (int)rank, x0, y0, n0);
        int tap = 0, kc = 0;
        for (int grp = 0; grp < g.k_groups; ++grp) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t a_dst = stage_base + s * stage_bytes;
          if (rank == 0) mbar_expect_tx(full_bar(s), 2u * stage_bytes);   // both CTAs' tiles land on the leader's barrier
          const uint32_t fbar = full_bar(s) & kPeerBitMask;
          for (int j = 0; j < g.kgroup; ++j) {
            const int dy = g.taps == 9 ? tap / 3 - 1 : 0;
            const int dx = g.taps == 9 ? tap % 3 - 1 : 0;
            tma_load_4d_2sm(a_dst + j * a_chunk, &tmA, fbar, kc * g.bk, x0 + dx, y0 + dy, n0);
            if (!g.w1_resident)
              tma_load_2d_2sm(a_dst + g.kgroup * a_chunk + j * b1_chunk, &tmW1, fbar, (grp * g.kgroup + j) * g.bk,
                              (int)rank * hid_local);
            if (++kc == g.kchunks) { kc = 0; ++tap; }
          }
          if (++s == g.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (lane == 0 && rank == 0) {
      // kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major, N>>3 at bit 17, M>>4 at bit 24 (M = 256 for the pair)
      const uint32_t idesc12 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.hid >> 3) << 17) | (16u << 24);
      const uint64_t desc_hi1 = umma_desc_kmajor(0, g.bk), desc_hi64 = umma_desc_kmajor(0, 64);
      const int ksteps1 = g.bk / 16;
      mbar_wait(w_full, 0);
      int s = 0;
      uint32_t ph = 0, tl = 0;
      unsigned c_g1 = 0, c_full = 0, c_d3 = 0, c_a2_0 = 0, c_a2_n = 0, c_a3_0 = 0, c_a3_n = 0;
      const unsigned c_start = dbg ? (unsigned)clock() : 0u;
      for (int it = it0; it < it_end; it += it_step, ++tl) {
        const uint32_t par = tl & 1u;
        const uint32_t r0 = tmem_base + (par ? 256u : 0u), r1 = tmem_base + (par ? 0u : 256u);
        const unsigned c_g1_0 = dbg ? (unsigned)clock() : 0u;
        unsigned long long* stp = (dbg && tl == 5u) ? dbg + (size_t)gridDim.x * 16 + (size_t)blockIdx.x * 16 : nullptr;
        if (stp) stp[0] = (unsigned)clock();                       // GEMM1 issue starts
        // ---- GEMM1 -> R0.  R0 was the previous tile's GEMM2 accumulator; this thread has waited for every a3_ready of that
        // tile, i.e. all epilogue warps have finished reading it.
        uint32_t b_res = w1_base;
        uint32_t accumulate = 0;
        for (int grp = 0; grp < g.k_groups; ++grp) {
          { NN_CNT_BEGIN(); mbar_wait(full_bar(s), ph); NN_CNT_END(c_full); }
          tc_fence_after();
          uint32_t a_addr = stage_base + s * stage_bytes;
          uint32_t b_addr = g.w1_resident ? b_res : a_addr + g.kgroup * a_chunk;
          for (int j = 0; j < g.kgroup; ++j) {
            const uint64_t adesc = desc_hi1 | (uint64_t)((a_addr & 0x3FFFFu) >> 4);
            const uint64_t bdesc = desc_hi1 | (uint64_t)((b_addr & 0x3FFFFu) >> 4);
#pragma unroll 4
            for (int k = 0; k < ksteps1; ++k) {
              // (copying each A slice to tensor memory with tcgen05.cp and running GEMM1 in TS form like GEMM2 / GEMM3 was
              // tried: correct, but 6.4k instead of 5.4k cycles per tile pair)
              umma_bf16_2sm(r0, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc12, accumulate);
              accumulate = 1;
            }
            a_addr += a_chunk;
            b_addr += b1_chunk;
          }
          b_res += g.kgroup * b1_chunk;
          umma_commit_2sm(empty_bar(s));   // frees the stage in BOTH CTAs
          if (++s == g.stages) { s = 0; ph ^= 1u; }
        }
        umma_bf16_2sm(r0, umma_desc_k16(ones_base), umma_desc_k16(bias1_base), idesc12, 1u);   // + ActNorm shift
        umma_commit_2sm(acc_full(0));
        if (dbg) c_g1 += (unsigned)clock() - c_g1_0;
        if (stp) stp[1] = (unsigned)clock();                       // GEMM1 issued + committed
        // ---- GEMM2 -> R1, A = bf16 h1 (in place in R0: channels [16u, 16u+16) sit in columns [16u, 16u+8)).  R1 held the
        // previous tile's h2 and tap accumulator: GEMM3 of that tile precedes us in the pipe, its accumulator must be drained.
        if (tl > 0) {
          { NN_CNT_BEGIN(); mbar_wait(d3_empty, (tl * (uint32_t)n3_passes - 1u) & 1u); NN_CNT_END(c_d3); }
          tc_fence_after();
        }
        accumulate = 0;
        for (int kc = 0; kc < k2chunks; ++kc) {
          { NN_CNT_BEGIN(); mbar_wait(a2_ready(kc), par); if (kc == 0) NN_CNT_END(c_a2_0); else NN_CNT_END(c_a2_n); }
          tc_fence_after();
          if (stp && kc == 0) stp[2] = (unsigned)clock();          // h1 chunk 0 seen
          if (stp && kc == 3) stp[3] = (unsigned)clock();          // h1 chunk 3 seen
          const uint64_t bdesc = desc_hi64 | (uint64_t)(((w2_base + kc * w2_chunk) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t a_col = (uint32_t)(64 * kc + 16 * k);
            umma_bf16_ts_2sm(r1, r0 + a_col, bdesc + (uint64_t)(2 * k), idesc12, accumulate);
            accumulate = 1;
          }
          if (kc == 0) umma_bf16_ts_2sm(r1, r0 + 8u, umma_desc_k16(bias2_base), idesc12, 1u);   // constant-one columns x shift
        }
        umma_commit_2sm(acc_full(1));
        if (stp) stp[4] = (unsigned)clock();                       // GEMM2 issued + committed
        // ---- GEMM3 -> R0 + 128, A = bf16 h2 (compact, R0 columns [0, hid/2)).  (Two accumulators fed by alternate K steps or by
        // alternate chunks were tried -- a dependent MMA starts ~160 cycles after its predecessor whatever N is -- and left the
        // kernel time unchanged: the chunks arrive at the pace of epilogue 2.)
        // More than 128 tap planes: passes of up to 128 accumulator columns, each drained by epilogue 3 before the next starts
        // (pass p takes weight rows [64p, 64p + N_p/2) of BOTH CTAs' halves, see the column -> plane map in epilogue 3).
        for (int ps = 0; ps < n3_passes; ++ps) {
          const int n_p = min(128, g.n3_pad - 128 * ps);
          const uint32_t idesc3 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_p >> 3) << 17) | (16u << 24);
          if (ps > 0) {
            mbar_wait(d3_empty, (tl * (uint32_t)n3_passes + (uint32_t)ps - 1u) & 1u);
            tc_fence_after();
          }
          accumulate = 0;
          for (int kc = 0; kc < k2chunks; ++kc) {
            if (ps == 0) {
              { NN_CNT_BEGIN(); mbar_wait(a3_ready(kc), par); if (kc == 0) NN_CNT_END(c_a3_0); else NN_CNT_END(c_a3_n); }
              tc_fence_after();
              if (stp && kc == 0) stp[5] = (unsigned)clock();          // h2 chunk 0 seen
              if (stp && kc == 3) stp[6] = (unsigned)clock();          // h2 chunk 3 seen
            }
            const uint64_t bdesc = desc_hi64 | (uint64_t)(((w9_base + kc * w9_chunk + (uint32_t)ps * 64u * 128u) & 0x3FFFFu) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_ts_2sm(r0 + 128u, r0 + (uint32_t)(kc * 32 + k * 8), bdesc + (uint64_t)(2 * k), idesc3, accumulate);
              accumulate = 1;
            }
          }
          umma_commit_2sm(acc_full(2));
        }
        if (stp) stp[7] = (unsigned)clock();                       // GEMM3 issued + committed
      }
      if (dbg) {
        unsigned long long* d = dbg + (size_t)blockIdx.x * 16;
        d[0] = c_g1; d[1] = c_full; d[2] = c_d3; d[3] = c_a2_0; d[4] = c_a2_n; d[5] = c_a3_0; d[6] = c_a3_n;
        d[7] = (unsigned)clock() - c_start;
      }
    }
    __syncwarp();
  } else if (warp >= kNNEpiWarp0) {
    // ===== epilogue warps: quadrant q owns TMEM lanes [32q, 32q+32) = tile rows; the four warps of a quadrant take the
    // 16-column slices `part` of every 64-channel chunk, so a chunk is complete -- and its MMAs can start -- after ONE short
    // iteration of all warps =====
    const int q = warp & 3, part = (warp - kNNEpiWarp0) >> 2;
    const int row = q * 32 + lane;
    const int ppi_log2 = g.tw_log2 + g.th_log2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const long long plane = (long long)g.H * g.W;
    const uint32_t qstg = stg_base + (uint32_t)q * 4096u;   // side outputs: this quadrant's staging block (32 pixels x 64 channels)
    const int r0row = q * 32;
    auto arrive_leader = [&](uint32_t bar) {
      if (rank == 0) mbar_arrive(bar);
      else mbar_arrive_cluster(bar, 0);
    };
    uint32_t tl = 0;
    unsigned c_w0 = 0, c_e1 = 0, c_w1 = 0, c_e2 = 0, c_w2 = 0, c_e3 = 0, c_e3ld = 0;
    const unsigned c_start = dbg ? (unsigned)clock() : 0u;
    for (int it = it0; it < it_end; it += it_step, ++tl) {
      const uint32_t par = tl & 1u;
      const uint32_t r0 = tmem_base + (par ? 256u : 0u) + lane_off, r1 = tmem_base + (par ? 0u : 256u) + lane_off;
      const int mt = 2 * it + (int)rank;
      int x0, y0, n0;
      tile_origin(mt, x0, y0, n0);
      const int b = n0 + (row >> ppi_log2);
      const int y = y0 + ((row >> g.tw_log2) & ((1 << g.th_log2) - 1));
      const int x = x0 + (row & ((1 << g.tw_log2) - 1));
      const bool valid = mt < g.m_tiles && b < g.B && y < g.H && x < g.W;
      const int sub_x = x0 + (r0row & ((1 << g.tw_log2) - 1));
      const int sub_y = y0 + ((r0row >> g.tw_log2) & ((1 << g.th_log2) - 1));
      const int sub_n = n0 + (r0row >> ppi_log2);
      // one activation epilogue: accumulator `src` -> activation -> bf16 at `dst` (8 columns per 16 channels); a 64-channel
      // chunk is announced as soon as this warp's 16 channels of it are in place.
      auto convert = [&](uint32_t src, uint32_t dst, bool in_place, int bar0, const CUtensorMap* map) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c0 = 16 * part + 64 * u;
          if (c0 < g.hid) {
            uint32_t v[16], pk[8];
            tmem_ld16_nowait(src + c0, v);
            tmem_wait_ld();
            act_pack(v, g.act_fn, pk);
            tmem_st8(dst + (uint32_t)(in_place ? c0 : c0 >> 1), pk);
            if (in_place && c0 == 0) {   // the constant-one columns of GEMM2's shift MMA: channels (1, 1, 0, ... 0) in the first gap
              const uint32_t one[8] = {0x3F803F80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
              tmem_st8(dst + 8u, one);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader(bar_base + 8u * (uint32_t)(bar0 + u));
          }
        }
        if (kStore) {
          // Side outputs, AFTER the loop (i.e. in the time this warp would otherwise wait for the next GEMM): the bf16 words are
          // read back from tensor memory; the four warps of a quadrant hold the four 16-channel pieces of the same 32 pixels x
          // 64 channels, assemble 128-byte rows in the quadrant's staging block (128-byte swizzle) and ONE TMA store moves it.
          // The quadrant barriers also make the read-back safe: tensor-memory lanes belong to a quadrant, and a warp can only
          // move on to overwrite them (next epilogue) after all four have passed the last barrier, i.e. finished reading.
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (64 * u < g.hid) {
              const int c0 = 16 * part + 64 * u;
              uint32_t w[8];
              tmem_ld8_nowait(dst + (uint32_t)(in_place ? c0 : c0 >> 1), w);
              tmem_wait_ld();
              if (part == 0 && lane == 0) bulk_wait_read0();   // the quadrant's previous store has finished reading the block
              named_bar(1 + q, 128);
              const uint32_t rowa = qstg + (uint32_t)lane * 128u, x7 = (uint32_t)lane & 7u;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + ((((uint32_t)(2 * part)) ^ x7) << 4)), "r"(w[0]),
                           "r"(w[1]), "r"(w[2]), "r"(w[3])
                           : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + ((((uint32_t)(2 * part + 1)) ^ x7) << 4)),
                           "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                           : "memory");
              fence_async_smem();
              named_bar(1 + q, 128);
              if (part == 0 && lane == 0 && mt < g.m_tiles) {
                tma_store_4d(map, qstg, 64 * u, sub_x, sub_y, sub_n);
                bulk_commit();
              }
            }
          }
        }
      };
      // ---- epilogue 1: GEMM1 accumulator -> bf16 h1, written back in place (each warp rewrites columns it has itself read)
      unsigned long long* stp = (dbg && tl == 5u && warp == kNNEpiWarp0 && lane == 0)
                                    ? dbg + (size_t)gridDim.x * 16 + (size_t)blockIdx.x * 16 : nullptr;
      { NN_CNT_BEGIN(); mbar_wait(acc_full(0), par); NN_CNT_END(c_w0); }
      tc_fence_after();
      unsigned c_t = dbg ? (unsigned)clock() : 0u;
      if (stp) stp[8] = c_t;                                       // GEMM1 accumulator seen
      convert(r0, r0, true, 2 * g.stages + 4, &tmH1);
      if (stp) stp[9] = (unsigned)clock();                         // epilogue 1 done
      // ---- epilogue 2: GEMM2 accumulator (R1) -> bf16 h2, compact in R0 (free: GEMM2 has consumed h1)
      if (dbg) c_e1 += (unsigned)clock() - c_t;
      { NN_CNT_BEGIN(); mbar_wait(acc_full(1), par); NN_CNT_END(c_w1); }
      tc_fence_after();
      c_t = dbg ? (unsigned)clock() : 0u;
      if (stp) stp[10] = c_t;                                      // GEMM2 accumulator seen
      convert(r1, r0, false, 2 * g.stages + 8, &tmH2);
      if (stp) stp[11] = (unsigned)clock();                        // epilogue 2 done
      // ---- epilogue 3: tap planes -> fp32 NCHW
      if (dbg) c_e2 += (unsigned)clock() - c_t;
      for (int ps = 0; ps < n3_passes; ++ps) {
        { NN_CNT_BEGIN(); mbar_wait(acc_full(2), (tl * (uint32_t)n3_passes + (uint32_t)ps) & 1u); NN_CNT_END(c_w2); }
        tc_fence_after();
        if (ps == 0) {
          c_t = dbg ? (unsigned)clock() : 0u;
          if (stp) stp[12] = c_t;                                    // GEMM3 accumulator seen
        }
        // accumulator column c of pass ps holds tap plane 64 ps + c (weight rows of the leader's half) for c < N_p / 2, else
        // n3_pad / 2 + 64 ps + (c - N_p / 2) (the peer's half); with a single pass that is plane c
        const int n_p = min(128, g.n3_pad - 128 * ps), half_p = n_p >> 1;
        // this warp's (at most two) 16-column slices; the accumulator is handed back as soon as it is in registers, before
        // the global stores go out
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c0 = 16 * part + 64 * i;
          uint32_t r[16];
          if (c0 < n_p) {   // warp-uniform
            NN_CNT_BEGIN();
            tmem_ld16_nowait(r0 + 128u + c0, r);
            tmem_wait_ld();
            NN_CNT_END(c_e3ld);
          }
          if (i == (n_p > 64 ? 1 : 0)) {   // nothing left to drain
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader(d3_empty);
          }
          if (c0 < n_p && valid) {
            float* dst = g.taps_out + ((long long)b * g.n3 * g.H + y) * g.W + x;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = c0 + j;
              const int pl = c < half_p ? 64 * ps + c : n3_local + 64 * ps + (c - half_p);
              if (c < n_p && pl < g.n3) dst[pl * plane] = __uint_as_float(r[j]);
            }
          }
          if (n_p <= 64) break;
        }
      }
      if (dbg) c_e3 += (unsigned)clock() - c_t;
      if (stp) stp[13] = (unsigned)clock();                        // epilogue 3 done
    }
    if (dbg && warp == kNNEpiWarp0 && lane == 0) {
      unsigned long long* d = dbg + (size_t)blockIdx.x * 16;
      d[8] = c_w0; d[9] = c_e1; d[10] = c_w1; d[11] = c_e2; d[12] = c_w2; d[13] = c_e3; d[14] = c_e3ld;
      d[15] = (unsigned)clock() - c_start;
    }
    if (kStore && part == 0 && lane == 0) bulk_wait0();   // outstanding TMA stores still read shared memory
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be arriving on this CTA's barriers / reading its weights
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace rfk

using namespace rfk;

extern "C" int rfk_coupling_nn_fused(const void* act, int B, int H, int W, int act_ld, int cin_pad, int taps, const void* w1,
                                     int hid, const void* w2, int act_fn, const void* w9, int n3, int w9_rows, float* taps_out,
                                     void* h1_out, void* h2_out, int h_ld, void* stream) {
  const char* who = "rfk_coupling_nn_fused";
  RFK_REQUIRE(act && w1 && w2 && w9 && taps_out && B > 0 && H > 0 && W > 0, "%s: null pointer or empty shape", who);
  RFK_REQUIRE(!conv_split_mode(), "%s: the hidden tiles are bf16 in tensor memory; not available in split-precision mode", who);
  RFK_REQUIRE(cin_pad > 0 && (cin_pad % 64 == 0 || cin_pad == 32) && cin_pad <= act_ld && act_ld % 8 == 0,
              "%s: cin_pad=%d must be 32 or a multiple of 64 and <= act_ld=%d (multiple of 8)", who, cin_pad, act_ld);
  RFK_REQUIRE(taps == 1 || taps == 9, "%s: taps=%d (only 1x1 and 3x3 kernels)", who, taps);
  RFK_REQUIRE(hid >= 64 && hid % 64 == 0 && hid <= 256, "%s: hidden=%d must be 64, 128, 192 or 256", who, hid);
  const int n3_pad = (n3 + 15) / 16 * 16;   // N of a cta_group::2 MMA: multiples of 16
  RFK_REQUIRE(n3 > 0 && n3_pad <= 256 && w9_rows >= n3, "%s: n3=%d (at most 256 tap planes), w9_rows=%d", who, n3, w9_rows);
  RFK_REQUIRE(act_fn >= 0 && act_fn <= 2, "%s: bad act_fn %d", who, act_fn);
  RFK_REQUIRE((h1_out == nullptr) == (h2_out == nullptr), "%s: h1_out and h2_out go together", who);
  const int store_h = h1_out != nullptr;
  if (store_h)
    RFK_REQUIRE(h_ld % 8 == 0 && hid <= h_ld &&
                    ((reinterpret_cast<uintptr_t>(h1_out) | reinterpret_cast<uintptr_t>(h2_out)) & 15) == 0,
                "%s: h1_out / h2_out must be 16-byte aligned NHWC bf16 with a row stride (h_ld=%d) that is a multiple of 8", who,
                h_ld);
  RFK_REQUIRE(((reinterpret_cast<uintptr_t>(act) | reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(w2) |
                reinterpret_cast<uintptr_t>(w9)) & 15) == 0, "%s: act / weights must be 16-byte aligned", who);
  NNArgs g;
  g.B = B; g.H = H; g.W = W; g.taps = taps; g.hid = hid; g.n3 = n3; g.n3_pad = n3_pad; g.act_fn = act_fn; g.store_h = store_h;
  {
    long long cap = 0;
    unsigned long long* tlb = debug_timeline(&cap);
    g.dbg = (tlb && cap >= 2 * sm_count()) ? tlb : nullptr;   // 16 counters + 16 stamps per CTA
  }
  g.taps_out = taps_out;
  g.bk = cin_pad % 64 == 0 ? 64 : 32;
  g.kchunks = cin_pad / g.bk;
  int twl = ilog2_ceil(W);
  if (twl > 7) twl = 7;
  int thl = ilog2_ceil(H);
  if (thl > 7 - twl) thl = 7 - twl;
  g.tw_log2 = twl; g.th_log2 = thl;
  const int TW = 1 << twl, TH = 1 << thl, NIMG = 128 / (TW * TH);
  g.tiles_x = ceil_div(W, TW);
  g.tiles_y = ceil_div(H, TH);
  g.m_tiles = g.tiles_x * g.tiles_y * ceil_div(B, NIMG);

  // shared-memory plan
  const int k1_iters = taps * g.kchunks;
  const long long a_chunk = 128LL * g.bk * 2, b1_chunk = (long long)(hid / 2) * g.bk * 2;
  const long long w2_bytes = (long long)(hid / 64) * (hid / 2) * 128, w9_bytes = (long long)(hid / 64) * (n3_pad / 2) * 128;
  const long long fixed = 1024 + w2_bytes + w9_bytes + 3 * 4096 + (store_h ? kNNEpiWarps * kNNStoreSlice : 0) +
                          8 * (2 * 12 + 13) + 16;
  const long long w1_bytes = (long long)k1_iters * b1_chunk;
  static const int force_stream = [] { const char* e = getenv("RFK_NN_W1_STREAM"); return e ? atoi(e) : 0; }();
  g.w1_resident = !force_stream && kNNSmemLimit - fixed - w1_bytes >= 4 * a_chunk;
  const long long room = kNNSmemLimit - fixed - (g.w1_resident ? w1_bytes : 0);
  const long long per_chunk = a_chunk + (g.w1_resident ? 0 : b1_chunk);
  // 64-byte-row chunks carry only two MMAs each: three per pipeline stage (one filter row; measured 7.9k -> 5.3k cycles per
  // tile pair in GEMM1 against one chunk per stage, even with only two stages)
  static const int kgroup_env = [] { const char* e = getenv("RFK_NN_KGROUP"); return e ? atoi(e) : 0; }();
  g.kgroup = 1;
  if (g.bk == 32 && k1_iters % 3 == 0 && room >= 2 * 3 * per_chunk) g.kgroup = 3;
  if (kgroup_env > 0 && k1_iters % kgroup_env == 0) g.kgroup = kgroup_env;
  g.k_groups = k1_iters / g.kgroup;
  int stages = (int)(room / (g.kgroup * per_chunk));
  if (stages > 12) stages = 12;
  RFK_REQUIRE(stages >= 2, "%s: weights leave no room for the activation pipeline (cin_pad=%d hid=%d n3=%d)", who, cin_pad, hid, n3);
  g.stages = stages;
  const size_t smem = (size_t)(1024 + (g.w1_resident ? w1_bytes : 0) + w2_bytes + w9_bytes + stages * g.kgroup * per_chunk +
                               3 * 4096 + (store_h ? kNNEpiWarps * kNNStoreSlice : 0) + 8 * (2 * stages + 13) + 16);
  RFK_REQUIRE(smem <= (size_t)kNNSmemLimit, "%s: internal error: %zu B of shared memory planned", who, smem);

  CUtensorMap tmA, tmW1, tmW2, tmW9, tmS1, tmS2, tmH1, tmH2;
  int rc = encode_act_map(&tmA, who, "A", act, cin_pad, act_ld, B, H, W, TW, TH, NIMG, g.bk);
  if (rc) return rc;
  const long long k1tot = (long long)taps * cin_pad + 16, k2tot = hid + 16;   // 16 shift columns behind each weight matrix
  rc = encode_weight_map(&tmW1, who, w1, k1tot, hid, hid / 2, g.bk);
  if (rc) return rc;
  rc = encode_weight_map(&tmW2, who, w2, k2tot, hid, hid / 2, 64);
  if (rc) return rc;
  rc = encode_weight_map(&tmW9, who, w9, hid, w9_rows, n3_pad / 2, 64);
  if (rc) return rc;
  {
    // the shift columns as one K = 16 slice of 32-byte rows (SWIZZLE_32B)
    EncodeTiledFn enc = encode_fn();
    RFK_REQUIRE(enc, "%s: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)", who);
    const void* ptrs[2] = {w1, w2};
    const long long ktots[2] = {k1tot, k2tot};
    CUtensorMap* maps[2] = {&tmS1, &tmS2};
    for (int i = 0; i < 2; ++i) {
      cuuint64_t dims[2] = {(cuuint64_t)ktots[i], (cuuint64_t)hid};
      cuuint64_t strides[1] = {(cuuint64_t)ktots[i] * 2};
      cuuint32_t box[2] = {16u, (cuuint32_t)(hid / 2)};
      cuuint32_t ones[2] = {1, 1};
      CUresult r = enc(maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptrs[i]), dims, strides, box, ones,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled(shift columns %d) failed with CUresult %d", who, i + 1, (int)r);
        return RFK_ECUDA;
      }
    }
  }
  tmH1 = tmA;
  tmH2 = tmA;
  if (store_h) {
    // the store maps' box is one QUADRANT's 32 pixel rows x 64 channels (128-byte rows, SWIZZLE_128B)
    const int sx = std::min(TW, 32), sy = std::min(TH, 32 / sx), sn = 32 / (sx * sy);
    rc = encode_act_map(&tmH1, who, "h1", h1_out, hid, h_ld, B, H, W, sx, sy, sn, 64);
    if (rc) return rc;
    rc = encode_act_map(&tmH2, who, "h2", h2_out, hid, h_ld, B, H, W, sx, sy, sn, 64);
    if (rc) return rc;
  }
  using KernelFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap,
                            NNArgs);
  const int variant = g.dbg ? 2 : store_h;
  KernelFn fn = variant == 2 ? (store_h ? (KernelFn)coupling_nn_kernel<true, true> : (KernelFn)coupling_nn_kernel<false, true>)
                : store_h    ? (KernelFn)coupling_nn_kernel<true, false>
                             : (KernelFn)coupling_nn_kernel<false, false>;
  static size_t configured[4] = {0, 0, 0, 0};
  const int slot = (g.dbg ? 2 : 0) + store_h;
  if (smem > configured[slot]) {
    cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", who, smem, cudaGetErrorString(e));
      return RFK_ECUDA;
    }
    configured[slot] = smem;
  }
  int pairs = sm_count() / 2;
  if (pairs > (g.m_tiles + 1) / 2) pairs = (g.m_tiles + 1) / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(kNNThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;   // two CTAs along x = the two SMs of a TPC
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaLaunchKernelEx(&cfg, fn, tmA, tmW1, tmW2, tmW9, tmS1, tmS2, tmH1, tmH2, g);
  return check_launch(who);
}
