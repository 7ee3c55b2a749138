/*
 * rfk.h -- C ABI of librfk.so: sm_100a CUDA kernels for the Glow flow step and the
 * ConvLSTM cell of cdglissov/recurrent-flows-msc.
 *
 * The reference has no FFI: the hot path sits behind plain torch.nn.Module classes
 * (SURVEY.md section 8b).  This header is the boundary a host binding (ctypes in this
 * repo; cffi / pybind / a C++ trainer elsewhere) links against.  Every entry point names
 * the reference code it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer unless it says "host"; no entry
 *     point allocates, frees, or synchronises; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - "NCHW f32" tensors are contiguous float32 [B,C,H,W]; "NHWC bf16" tensors are
 *     bfloat16 [B,H,W,ld] with `ld` (elements) >= the channels used;
 *   - return value: 0 on success, a negative RFK_E* code otherwise, with a
 *     human-readable message available from rfk_last_error() (thread-local);
 *   - nothing here falls back to the CPU: without an sm_100 device the launch fails and
 *     the error is returned.
 */
#ifndef RFK_H_
#define RFK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFK_VERSION 1

#define RFK_OK 0
#define RFK_EINVAL (-1)   /* bad argument (shape, alignment, enum)           */
#define RFK_ECUDA (-2)    /* CUDA runtime / driver error at launch           */
#define RFK_EUNSUPPORTED (-3)

/* activation after the per-channel affine of a conv epilogue (Utils/modules.py:8-19) */
#define RFK_ACT_NONE 0
#define RFK_ACT_RELU 1
#define RFK_ACT_LEAKY 2   /* LeakyReLU(0.2) */

/* log-scale clamp of the affine coupling (Flow/glow_modules.py:252-268) */
#define RFK_CLAMP_NONE 0
#define RFK_CLAMP_REALNVP 1   /* scale[j]*tanh(s)+scale_shift[j] */
#define RFK_CLAMP_GLOW 2      /* log(sigmoid(s+2))               */
#define RFK_CLAMP_SOFT 3      /* 2.5*0.636*atan(s/2.5)           */

/* how (mean, raw log-scale) are interleaved in a parameter tensor (Utils/utils.py:86-91) */
#define RFK_PAIR_CROSS 0   /* mean = ch 2j, raw = ch 2j+1      (coupling, Split2d) */
#define RFK_PAIR_SPLIT 1   /* mean = ch j,  raw = ch n+j       (ListGlow prior)    */

/* std from the raw log-scale (Flow/glow_modules.py:340-344, Flow/glow.py:139) */
#define RFK_STD_SOFTPLUS 0 /* softplus(raw)+1e-8 */
#define RFK_STD_EXP 1      /* exp(raw)           */

/* output kind of rfk_conv_gemm */
#define RFK_OUT_NHWC_BF16 0
#define RFK_OUT_NCHW_F32 1

int rfk_version(void);
const char* rfk_last_error(void);
/* host out-params; returns RFK_ECUDA when no device is visible */
int rfk_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- a4  Squeeze2d.forward (Flow/glow_modules.py:298-310) -------------------------------
 * undo=0: x [B,C,H,W] -> y [B,4C,H/2,W/2], y[b,4c+2dy+dx,i,j] = x[b,c,2i+dy,2j+dx]
 * undo=1: x [B,C,H,W] -> y [B,C/4,2H,2W]  (exact inverse).  B,C,H,W describe x. Bit exact. */
int rfk_squeeze2d(const float* x, float* y, int B, int C, int H, int W, int undo, void* stream);

/* ---- a1  ActNorm.forward (Flow/glow_modules.py:38-54) ------------------------------------
 * reverse=0: y = (x + bias[c]) * exp(logs[c]);  reverse=1: y = x*exp(-logs[c]) - bias[c].
 * The scalar log-det term H*W*sum(logs) is the caller's (it needs no kernel). */
int rfk_actnorm(const float* x, float* y, const float* bias, const float* logs,
                int B, int C, int HW, int reverse, void* stream);

/* ---- a1  ActNorm.initialize (Flow/glow_modules.py:26-31) ---------------------------------
 * Per-channel statistics of x [B,C,HW] over (B,HW): bias[c] = -mean,
 * logs[c] = log(1/(std_unbiased + 1e-6)).  mean_out/std_out (nullable) receive the raw stats. */
int rfk_actnorm_init(const float* x, float* bias, float* logs, float* mean_out, float* std_out,
                     int B, int C, int HW, void* stream);

/* ---- f3  BatchNormFlow (Flow/glow_modules.py:56-104), the flow_norm='batchnorm' alternative to ActNorm ---------
 * Per-POSITION parameters [1,C,H,W], statistics over the batch dimension only.  n = C*H*W.
 * rfk_batch_stats_pos: mean[i], var[i] = biased variance + eps over b of x[b,i]   (training-mode forward)
 * rfk_affine_pos     : y[b,i] = x[b,i]*a[i] + c[i].  Forward a = e^{log_gamma}/sqrt(var), c = beta - mean*a;
 *                      reverse a = sqrt(var)/e^{log_gamma}, c = mean - beta*a (the [C,H,W] algebra is the caller's). */
int rfk_batch_stats_pos(const float* x, float* mean, float* var, int B, long long n, float eps, void* stream);
int rfk_affine_pos(const float* x, float* y, const float* a, const float* c, int B, long long n, void* stream);

/* ---- a2 (+a1 folded)  InvConv.forward apply (Flow/glow_modules.py:213,218) ----------------
 * y[b,o,p] = sum_i Wm[o,i] * x[b,i,p] + bvec[o]   (Wm [C,C] row-major f32, bvec nullable).
 * With Wm = W*diag(exp(logs)), bvec = Wm*bias this is ActNorm followed by InvConv in one pass.
 * Optional side output: channels [0,side_n) of y are also written as bf16 into an NHWC buffer
 * at channel offset side_off with row stride side_ld (the coupling network's z1 input).
 * Optional log-det fold: logdet[b] += alpha * (*addend) for b < B (addend = device scalar H*W*(sum logs + log|det W|),
 * Flow/glow_modules.py:43,196), so a GlowStep needs no separate launch for its parameter-only log-det term. */
int rfk_mix1x1(const float* x, float* y, const float* Wm, const float* bvec, int B, int C, int HW,
               void* side_nhwc_bf16, int side_n, int side_off, int side_ld,
               float* logdet, const float* addend, float alpha, void* stream);

/* ---- layout helpers -----------------------------------------------------------------------
 * pack: channels [c_lo, c_lo+n) of src NCHW f32 [B,Csrc,HW] (batch stride src_bstride elements,
 * 0 = dense) -> dst NHWC bf16 at channel dst_off (row stride dst_ld).  Replaces torch.cat([z1, condition]) (Flow/glow_modules.py:273,355). */
int rfk_pack_nhwc_bf16(const float* src, long long src_bstride, int B, int Csrc, int HW, int c_lo, int n,
                       void* dst, int dst_off, int dst_ld, void* stream);
/* dst[b, dst_off+j, p] = src[b, src_off+j, p] for j<n  (the torch.cat halves, glow_modules.py:290,368) */
int rfk_copy_channels(const float* src, int src_C, int src_off, float* dst, int dst_C, int dst_off,
                      int n, int B, int HW, void* stream);

/* ---- a3/a5/a7/a9  convolution as implicit GEMM on tcgen05 (Flow/glow_modules.py:111,129,
 * Utils/modules.py:338-343) -------------------------------------------------------------------
 * act : NHWC bf16 [B,H,W,act_ld]; the first cin_pad channels (multiple of 64) are consumed.
 * wgt : bf16 [n_pad, taps*cin_pad] row-major, k = tap*cin_pad + c, tap = 3*ky+kx (taps = 1 or 9,
 *       'same' zero padding); rows >= n are zero; n_pad is a multiple of 16.
 * epilogue: v = acc*scale[o] + shift[o] (either nullable), then act_fn, then
 *   out_kind RFK_OUT_NHWC_BF16: out bf16 [B,H,W,out_ld] at channel offset out_off
 *   out_kind RFK_OUT_NCHW_F32 : out f32 [B,n,H,W]  (out_ld/out_off ignored)
 * ActNorm after the conv (Conv2dNorm) is scale=exp(logs), shift=bias*exp(logs);
 * Conv2dZeros is scale=exp(3*logs), shift=conv.bias*exp(3*logs). */
int rfk_conv_gemm(const void* act, int B, int H, int W, int act_ld, int cin_pad,
                  const void* wgt, int n, int n_pad, int taps,
                  const float* scale, const float* shift, int act_fn,
                  int out_kind, void* out, int out_ld, int out_off, void* stream);

/* Same GEMM with the affine-coupling tail fused into the epilogue
 * (Flow/glow_modules.py:275-290): n = C output channels as (shift_j, raw_j) pairs;
 * z [B,C,H,W] f32 is updated IN PLACE on channels [C/2, C):
 *   reverse=0: z2 = (z2 + shift)*exp(ls), logdet[b] += sum ls
 *   reverse=1: z2 = z2*exp(-ls) - shift,  logdet[b] -= sum ls
 * ls = clamp(raw) with clamp_scale/clamp_shift [C/2] for RFK_CLAMP_REALNVP. logdet nullable. */
int rfk_conv_gemm_coupling(const void* act, int B, int H, int W, int act_ld, int cin_pad,
                           const void* wgt, int n, int n_pad, int taps,
                           const float* scale, const float* shift,
                           float* z, int clamp_type, const float* clamp_scale, const float* clamp_shift,
                           float* logdet, int reverse, void* stream);

/* Same GEMM with the ConvLSTM cell update fused into the epilogue (Utils/modules.py:367-377).
 * wgt rows are tile-interleaved: row (t*4+g)*ht_pad + j holds gate g in (i,f,o,g) of hidden
 * channel t*ht + j (rows with j >= ht are zero); n_pad = n_tiles*4*ht_pad; bias likewise.
 * c_prev, c_next, h_out: f32 NCHW with batch strides (elements) so h_out can alias a slice of the
 * [B,T,Hc,H,W] output stack; peep (nullable) = Wci,Wcf,Wco as [3,Hc,H,W];
 * h_nhwc (nullable): bf16 copy of h' into the next step's NHWC input at channel h_off, stride h_ld. */
int rfk_conv_gemm_lstm(const void* act, int B, int H, int W, int act_ld, int cin_pad,
                       const void* wgt, int hidden, int ht, int ht_pad, int taps, const float* bias,
                       const float* c_prev, long long c_prev_bstride, const float* peep,
                       float* c_next, long long c_next_bstride, float* h_out, long long h_bstride,
                       void* h_nhwc, int h_off, int h_ld, void* stream);

/* Split-K form of rfk_conv_gemm for launches whose pixel tiles cannot fill the GPU (the RFN ConvLSTM runs on 2x2 maps:
 * 120 pixels, K = 9*712): the K loop is cut into k_split slices (gridDim.z), each CTA adds its partial tile into the
 * caller-zeroed fp32 workspace ws[pixel, ws_ld] (pixel = (b*H + y)*W + x, columns = output channels) with vector
 * red.global.add.  No epilogue math: pair it with rfk_convlstm_pointwise_ws (or any pixel-major consumer). */
int rfk_conv_gemm_splitk(const void* act, int B, int H, int W, int act_ld, int cin_pad,
                         const void* wgt, int n, int n_pad, int taps, int k_split,
                         float* ws, int ws_ld, void* stream);

/* Split-K with the reduction fused in: like rfk_conv_gemm (NHWC bf16 output, per-channel affine + activation), but K
 * is cut into k_split slices that run on different SMs.  Slice z stores its partial tile with plain stores into its own
 * slab of the fp32 workspace (ws must hold k_split * ceil(pixels/128)*128 * n_pad floats with ws_ld == n_pad; tile-local
 * layout, contents need not be initialised), and the
 * CTA that contributes the last slice of a tile (per-tile counter, zero before the launch, zero again after it) sums the
 * slabs, applies the epilogue and writes the bf16 tile.  For layers whose pixel tiles alone cannot fill the GPU (deep
 * levels of the flow; every level when sampling a few sequences).  (An earlier version added the partial tiles with
 * red.global.add into one slab: the L2 atomic rate made it slower than not splitting at all.) */
int rfk_conv_gemm_splitk_fused(const void* act, int B, int H, int W, int act_ld, int cin_pad,
                               const void* wgt, int n, int n_pad, int taps, int k_split,
                               float* ws, int ws_ld, unsigned int* counters,
                               const float* scale, const float* shift, int act_fn,
                               void* out, int out_ld, int out_off, void* stream);

/* ---- a3  coupling tail, standalone (Flow/glow_modules.py:275-290) -------------------------
 * nn_out [B,C,H,W] f32 = output of the coupling network; z as in rfk_conv_gemm_coupling. */
int rfk_coupling_tail(const float* nn_out, float* z, int B, int C, int HW,
                      int clamp_type, const float* clamp_scale, const float* clamp_shift,
                      float* logdet, int reverse, void* stream);

/* Coupling tail fed by a TAP-SPLIT convolution.  For few output channels (9*C <= 256) the coupling network's
 * last 3x3 conv (256 -> C) is cheaper as ONE 1x1 GEMM with N = 9*C (rfk_conv_gemm, taps=1, weight row
 * t*C + c = W[c, :, ky, kx], t = 3*ky+kx, f32 NCHW output `taps` [B,9C,H,W]): the activations are read once
 * instead of once per tap.  This kernel sums the nine planes shifted by their tap offset (zero outside the image),
 * applies Conv2dZeros' affine (scale/shift [C]) and then the coupling update exactly like rfk_coupling_tail. */
int rfk_coupling_tail_taps(const float* taps, float* z, int B, int C, int H, int W,
                           const float* scale, const float* shift,
                           int clamp_type, const float* clamp_scale, const float* clamp_shift,
                           float* logdet, int reverse, void* stream);

/* conv1x1 -> ActNorm -> activation -> tap-split conv3x3 in ONE kernel (Flow/glow_modules.py:235-237 = net.2, net.3,
 * net.4 of the coupling network): the hidden tensor h2 is produced in tensor memory as bf16 and consumed from there
 * by the second GEMM, so it never touches HBM.  act: NHWC bf16 h1 (first cin_pad channels, multiple of 64);
 * w2: bf16 [hid, cin_pad]; scale2/shift2: ActNorm affine [hid]; w9: bf16 [n3_pad, hid] in tap-split row order
 * (see rfk_coupling_tail_taps); taps: fp32 NCHW [B, n3, H, W].  hid in {64,128,192,256}, n3_pad <= 128. */
int rfk_conv1x1_taps_fused(const void* act, int B, int H, int W, int act_ld, int cin_pad,
                           const void* w2, int hid, const float* scale2, const float* shift2, int act_fn,
                           const void* w9, int n3, int n3_pad, float* taps, void* stream);

/* The WHOLE coupling network in one kernel (Flow/glow_modules.py:229-240 = net.0 .. net.4 of AffineCoupling):
 * conv3x3 (or 1x1) -> ActNorm -> activation -> conv1x1 -> ActNorm -> activation -> tap-split conv3x3, with BOTH hidden
 * tensors kept in tensor memory (CTA pairs, tcgen05 cta_group::2; csrc/coupling_nn.cu).  act: NHWC bf16 network input
 * [cond | z1] (first cin_pad channels: 32 or a multiple of 64); w1f: bf16 [hid, taps*cin_pad + 16] and w2f: bf16
 * [hid, hid + 16], both with their ActNorm folded in (rfk_pack_weight_folded: the per-channel scale lives in the weight rows,
 * the shift rides through the GEMM against a constant-one operand, so the epilogues are activation + rounding only);
 * w9: bf16 [w9_rows >= n3, hid] in tap-split row order (see rfk_coupling_tail_taps); taps_out: fp32 NCHW [B, n3, H, W],
 * n3 <= 256 (GEMM3 runs in passes of 128 accumulator columns).  h1_out / h2_out (both or neither; NHWC bf16, row stride h_ld): the hidden activations as side outputs for a
 * backward pass -- written once by TMA store, never read back here. */
int rfk_coupling_nn_fused(const void* act, int B, int H, int W, int act_ld, int cin_pad, int taps, const void* w1f,
                          int hid, const void* w2f, int act_fn, const void* w9, int n3, int w9_rows, float* taps_out,
                          void* h1_out, void* h2_out, int h_ld, void* stream);

/* rfk_coupling_tail_taps fused with the 1x1 mix that follows it (forward: the next GlowStep's ActNorm+InvConv;
 * reverse: the same GlowStep's InvConv^-1 + ActNorm^-1): z [B,C,H,W] holds the coupling's input (z1 | not yet updated z2),
 * the updated tensor is formed in shared memory only and y = Wm*(z1 | z2') + bvec is written (plus the optional bf16
 * side output and log-det scalar exactly as in rfk_mix1x1).  cpl_logdet (nullable) receives +/- sum(ls). */
int rfk_coupling_taps_mix(const float* taps, const float* z, float* y, int B, int C, int H, int W,
                          const float* scale, const float* shift,
                          int clamp_type, const float* clamp_scale, const float* clamp_shift,
                          float* cpl_logdet, int reverse,
                          const float* Wm, const float* bvec,
                          void* side_nhwc_bf16, int side_n, int side_off, int side_ld,
                          float* logdet, const float* addend, float alpha, void* stream);

/* ---- a5/a7  Gaussian log-density and sampling (Flow/glow_modules.py:362-368, glow.py:139,154)
 * params [B,2n,HW] f32 (nullable = zeros) holds (mean, raw) per `pairing`; std per `std_kind`.
 * logp : logdet[b] += sum_{j<n,p} log N(z[b,z_off+j,p]; mean, std)        (z has z_C channels)
 * sample: out[b,out_off+j,p] = mean + std*temperature*eps[b,j,p]           (out has out_C channels) */
int rfk_gauss_logp(const float* z, int z_C, int z_off, const float* params, int n, int B, int HW,
                   int pairing, int std_kind, float* logdet, void* stream);
int rfk_gauss_sample(const float* eps, const float* params, int n, int B, int HW, int pairing,
                     int std_kind, float temperature, float* out, int out_C, int out_off, void* stream);

/* ---- a9  ConvLSTM cell update, standalone (Utils/modules.py:369-377) ----------------------
 * cc [B,4Hc,HW] f32 in gate order i,f,o,g (bias already added); peep nullable [3,Hc,HW]. */
int rfk_convlstm_pointwise(const float* cc, const float* c_prev, const float* peep,
                           float* h_out, float* c_next, int B, int Hc, int HW, void* stream);

/* ---- backward kernels (driven by Flow/training.py and Utils/training.py behind ListGlow.log_prob / ConvLSTM) -------------
 * Data gradient of a conv: rfk_conv_gemm on the tap-flipped, transposed weights (Wd[ci, co, ky, kx] = W[co, ci, 2-ky, 2-kx]).
 *
 * rfk_act_affine_bwd: backward of h = act(a*scale + shift) (Conv2dNorm + ActFun, Flow/glow_modules.py:139-147) from the
 *   saved output h and the upstream gradient dh (both NHWC bf16, `rows` pixels, n channels, row stride ld):
 *   da = dh*act'(h)*scale (bf16, row stride da_ld);  r_dv[c] += sum dh*act'(h)  (d bias = scale*r_dv);
 *   r_dvv[c] += dvv_factor * sum dh*act'(h)*v = dvv_factor * sum dh*h  (d logs; factor 3 for Conv2dZeros' exp(3*logs)).
 *   dv_scaled != 0 multiplies the r_dv sums by scale[c], which makes them d bias directly.  r_dv / r_dvv: caller-zeroed fp32 [n].
 * rfk_conv_wgrad: dw[tap][n][c] += sum_p dy[p, n] * x[p + off(tap), c]  (x, dy NHWC bf16; zero outside the image;
 *   dw fp32, caller-zeroed, accumulated with atomics; tap = 3*ky + kx).  layout 0: dw[taps][cout][dw_ld]; layout 1: the conv
 *   weight's own layout dw[cout][dw_ld][taps] with input channel c stored at perm[c] (perm nullable; staging order ->
 *   weight order).  ws (nullable, 16-byte aligned, ws_bytes): scratch for the per-pixel-slice partial tiles; when it is
 *   large enough the slices are summed by a second kernel (dw += sum, no atomics), else they are added atomically.
 *   tcgen05 kernel with MN-major operands (csrc/wgrad_tc.cu); RFK_WGRAD_WMMA=1 selects the mma.sync one. */
int rfk_act_affine_bwd(const void* dh, const void* h, int ld, int n, const float* scale, int act_fn,
                       void* da, int da_ld, float* r_dv, float* r_dvv, float dvv_factor, int dv_scaled, long long rows,
                       void* stream);
int rfk_conv_wgrad(const void* x, int x_ld, int cin, const void* dy, int dy_ld, int cout, int B, int H, int W, int taps,
                   float* dw, int dw_ld, int layout, const int* perm, void* ws, long long ws_bytes, void* stream);

/* Affine-coupling tail backward, tap-split form (backward of rfk_coupling_tail_taps; Flow/glow_modules.py:237-273).
 *   dz [B,C,H,W] holds the gradient w.r.t. the coupling output; its z2 half (channels C/2..C) is overwritten with the
 *   gradient w.r.t. z2.  z_out = the coupling output, taps = the nine tap planes [B,9C,H,W], g_ld (nullable) [B] = gradient
 *   w.r.t. logdet.  dsum [B,C,H,W] receives the gradient w.r.t. the Conv2dZeros pre-affine sums (channel 2j: shift part,
 *   2j+1: log-scale part).  d_scale/d_shift [C] (Conv2dZeros exp(3*logs) and bias*exp(3*logs) affine) and, for the realnvp
 *   clamp, d_clamp_scale/d_clamp_shift [C/2] are caller-zeroed fp32 accumulators.  logs_factor != 0 (3 for Conv2dZeros)
 *   converts on the fly: d_scale receives d logs = f*(d scale*scale + d shift*shift), d_shift receives d bias = d shift*scale.
 * rfk_taps_scatter: dtaps[p, t*C + c] = dsum[c](p - off(t)) as NHWC bf16 (row stride ld >= 9C; pad columns untouched).
 * rfk_mix1x1_wgrad: dW[o,i] += sum dy[b,o,p]*x[b,i,p], db[o] += sum dy[b,o,p] (fp32 NCHW, C <= 64, caller-zeroed).
 * rfk_gauss_logp_bwd: backward of rfk_gauss_logp with upstream g[b]: dz[:, z_off:z_off+n] += ..., dparams (nullable with
 *   params) [B,2n,HW] = gradient w.r.t. the (mean, raw-scale) planes in the same pairing. */
int rfk_coupling_taps_bwd(const float* taps, const float* z_out, float* dz, float* dsum, int B, int C, int H, int W,
                          const float* scale, const float* shift, int clamp_type, const float* clamp_scale,
                          const float* clamp_shift, const float* g_ld, float* d_scale, float* d_shift,
                          float* d_clamp_scale, float* d_clamp_shift, float logs_factor, void* stream);
int rfk_taps_scatter(const float* dsum, void* dtaps, int ld, int B, int C, int H, int W, void* stream);
int rfk_mix1x1_wgrad(const float* x, const float* dy, int B, int C, int HW, float* dW, float* db, void* stream);
int rfk_gauss_logp_bwd(const float* z, int z_C, int z_off, const float* params, int n, int B, int HW, int pairing,
                       int std_kind, const float* g, float* dz, float* dparams, void* stream);

/* Weight repacking (one launch per conv weight per optimizer step): fp32 [N, Cin, taps] (tap = 3*ky + kx) -> bf16 K-major
 * GEMM operand dst[rows_pad, ktot], zero-padded.  mode 0: forward conv, row n, k = t*kp + j <- W[n, perm[j], t] (perm
 * nullable = identity over Cin, staging-buffer channel order); mode 1: data-gradient conv, row r (rows = len(perm) or Cin),
 * k = t*kp + co <- W[co, perm[r], taps-1-t]; mode 2: tap-split 1x1 form, row t*N + c, k = j <- W[c, j, t]; mode 3: tap-split
 * data gradient, row t*R + j (R = rows/taps), k = co <- W[co, perm[j], taps-1-t]; perm[j] < 0 (or j >= Cin without perm) = zero row. */
int rfk_pack_weight(const float* src, int N, int Cin, int taps, int mode, const int* perm, int rows, int kp,
                    void* dst, int rows_pad, int ktot, void* stream);

/* Conv weights with an ActNorm folded in (Flow/glow_modules.py:140-146: Conv2dNorm = conv, then ActNorm).
 * mode 4 -- forward weight + the ActNorm that FOLLOWS the conv: dst bf16 [rows_pad, ktot], ktot = taps*kp + 16, rows = N;
 *   row n, k = t*kp + j <- W[n, perm[j], t] * exp(logs[n]); columns taps*kp and taps*kp + 1 hold the shift
 *   bias[n] * exp(logs[n]) split into two bf16 words (hi, lo), the other 14 extra columns are zero.  Operand of
 *   rfk_coupling_nn_fused, which multiplies the extra columns by a constant one.
 * mode 5 -- data-gradient weight (rfk_pack_weight mode 1 without perm) + the ActNorm that PRODUCED the conv's input: row r
 *   (rows = Cin), k = t*kp + co <- W[co, r, taps-1-t] * exp(logs[r]); ktot = taps*kp, bias unused.  Operand of
 *   rfk_conv_gemm_actbwd called with scale = NULL. */
int rfk_pack_weight_folded(const float* src, int N, int Cin, int taps, int mode, const int* perm, int rows, int kp,
                           const float* logs, const float* bias, void* dst, int rows_pad, int ktot, void* stream);

/* Split-precision ("bf16x3") convolutions: the fp32-accurate mode behind the 1e-3 parity gate (BASELINE.json north_star:
 * "bf16/tf32 ... within rtol 1e-3"; tf32 keeps 11 significant bits, this mode 16).  Every conv operand is the sum of two
 * bf16 words, a = a_hi + a_lo: activation buffers hold [hi | lo] halves per row (row stride 2 x cin_pad; written by
 * rfk_pack_nhwc_bf16 + rfk_pack_nhwc_bf16_lo and by the conv epilogues), packed weights hold [w_hi | w_hi | w_lo] per tap,
 * and the SAME tcgen05 kernels run a K loop three times as long that visits the activation halves as hi, lo, hi:
 * a_hi w_hi + a_lo w_hi + a_hi w_lo, accumulated in fp32.  Process-global switch (set once at start-up, not thread-safe);
 * the bf16-only fusions (rfk_conv1x1_taps_fused, rfk_conv_gemm_splitk_fused, TMA-store epilogue) refuse / step aside. */
int rfk_set_conv_split(int on);
/* Programmatic dependent launch for the launches that follow (process-global; returns the previous setting, which is NOT an
 * error code).  Every kernel of the library waits (griddepcontrol.wait) before reading what a predecessor wrote and lets
 * its successor's prologue start early; kernels that write packed weights do not release their dependents early, because
 * the convolution kernels prefetch weights before their own wait.  The environment variable RFK_PDL=1/0 overrides. */
int rfk_set_pdl(int on);
int rfk_pack_nhwc_bf16_lo(const float* src, long long src_bstride, int B, int Csrc, int HW, int c_lo, int n, void* dst,
                          int dst_off, int dst_ld, void* stream);

/* Batched refresh of parameter-derived tensors: ONE launch per kind for a whole model, driven by a device table of
 * 24 x 64-bit words per entry (pointers and integers alike; csrc/prepare.cu documents the word layout of each kind).
 * They replace, per optimizer step, ~340 rfk_pack_weight launches and ~1500 ATen launches that rebuilt the folded
 * ActNorm . InvConv matrices (Flow/glow_modules.py:188-205) and the per-channel (scale, shift) of every ActNorm /
 * Conv2dZeros (Flow/glow_modules.py:41-50,120).
 *   rfk_pack_weights_batched    entries = the arguments of rfk_pack_weight; max_elements = largest rows_pad*ktot
 *   rfk_affine_prepare_batched  scale = exp(f*logs), shift = bias*scale
 *   rfk_fold_prepare_batched    Wf = P (L o mask + I)(U o mask^T + diag(sign_s e^log_s)) diag(e^logs), its transpose,
 *                               bf = Wf bias, {per-pixel log-det, HW * per-pixel log-det}
 *   rfk_fold_backward_batched   (dWf, dbf, g_sum = sum_b dloss/dlogdet[b]) -> ACCUMULATES d bias, d logs, d lower, d upper,
 *                               d log_s (autograd equivalent of Flow/glow_modules.py:33-54,188-205)
 * rfk_add_channels: dst[:, dst_off : dst_off+n] += src[:, src_off : src_off+n] (fp32 NCHW). */
int rfk_pack_weights_batched(const long long* table, int n_entries, long long max_elements, void* stream);
int rfk_affine_prepare_batched(const long long* table, int n_entries, void* stream);
int rfk_fold_prepare_batched(const long long* table, int n_entries, void* stream);
int rfk_fold_backward_batched(const long long* table, int n_entries, const float* g_sum, void* stream);
int rfk_add_channels(float* dst, int dst_C, int dst_off, const float* src, int src_C, int src_off, int n, int B, int HW,
                     void* stream);
/* Data gradient of a convolution FUSED with the backward of the h = act(ActNorm(.)) that produced its input (the coupling
 * network's hidden layers, Flow/glow_modules.py:229-238): out (bf16 NHWC, TMA stores) = (act (*) wgt)[p,c] * act'(h[p,c]) *
 * scale[c] with h the saved activation (its sign decides act'), and colsum[c] += sum_p out[p,c] (atomics).  Replaces a
 * rfk_conv_gemm + rfk_act_affine_bwd pair: the dh tensor never exists in HBM.  n must be a multiple of 64 without padding.
 * rfk_actnorm_param_bwd then gives the ActNorm gradients of that layer from the layer's own weight gradient:
 *   d bias[c] += colsum[c];  d logs[c] += sum_k W[c,k] dW[c,k] + bias[c] colsum[c];  grad_W += dW   (K = Cin*taps). */
int rfk_conv_gemm_actbwd(const void* act, int B, int H, int W, int act_ld, int cin_pad, const void* wgt, int n, int n_pad,
                         int taps, const float* scale, int act_fn, const void* h, int h_ld, void* out, int out_ld,
                         float* colsum, void* stream);
int rfk_actnorm_param_bwd(const float* W, const float* dW, long long K, const float* colsum, const float* bias,
                          float* grad_W, float* d_logs, float* d_bias, int n, void* stream);
/* out[b,c,p] (fp32 NCHW) = ws[(b*HW + p)*ld + c] + bias[c] (bias nullable): the pixel-major accumulator of rfk_conv_gemm_splitk
 * in the layout the cell-update / backward kernels read; zero != 0 clears ws for the next split-K launch. */
int rfk_ws_to_nchw(float* ws, int ld, const float* bias, float* out, int B, int C, int HW, int zero, void* stream);

/* Adam over all parameters in one launch (torch.optim.Adam without weight decay / amsgrad; the reference trains with
 * Adam, RFN/trainer.py).  p, g, m, v: flat fp32 buffers of n elements (n % 4 == 0, 16-byte aligned); g is multiplied by
 * grad_scale first (1/world after a sum-allreduce); *step = the 1-based step count, on the device (graph replay). */
int rfk_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float grad_scale, const float* step, void* stream);

/* ConvLSTM cell update, backward (Utils/modules.py:369-377): gates recomputed from the saved pre-activations cc
 * [B,4Hc,H,W] (bias included; order i,f,o,g) and c_prev (nullable = 0); dh [B,Hc,H,W] with batch stride dh_bstride,
 * dc_in (nullable) = gradient w.r.t. c_next.  Outputs dcc [B,4Hc,H,W], dc_prev [B,Hc,H,W]; dbias [4Hc] (nullable)
 * accumulates sum_{b,p} dcc (caller-zeroed).  Peepholes (nullable [3,Hc,HW]) are constants, as in the reference. */
int rfk_convlstm_pointwise_bwd(const float* cc, const float* c_prev, const float* peep, const float* dh,
                               long long dh_bstride, const float* dc_in, float* dcc, float* dc_prev, float* dbias,
                               int B, int Hc, int HW, void* stream);

/* Gather of nine tap planes stored NHWC bf16 (output of a tap-split 1x1 GEMM with N = 9*n_stride, row stride ld):
 * out[b, j, y, x] = sum_t T[b, y+ky-1, x+kx-1, t*n_stride + j] (zero outside the image), fp32 NCHW [B, n, H, W];
 * n <= n_stride <= 128, n_stride a multiple of 8 (tap segments are whole 16-byte chunks). */
int rfk_taps_gather_nhwc(const void* T, int ld, int n, int n_stride, int B, int H, int W, float* out, void* stream);
/* The same gather ACCUMULATED into two fp32 NCHW tensors instead of written to a fresh one: channels [0, n0) are added to
 * acc0[:, 0:n0] (acc0_C channels per sample), channels [n0, n) to acc1[:, 0:n-n0] (acc1_C channels).  In the coupling
 * backward (Flow/glow_modules.py:271-291 differentiated) these are the condition's gradient and dz[:, :C/2] += d z1: the
 * network-input gradient tensor and two rfk_add_channels launches per GlowStep disappear. */
int rfk_taps_gather_nhwc_acc(const void* T, int ld, int n, int n_stride, int B, int H, int W, float* acc0, int acc0_C,
                             int n0, float* acc1, int acc1_C, void* stream);

/* Debug aid: when buf != NULL, every conv-GEMM CTA of later launches (grids of at most capacity_ctas CTAs)
 * writes 16 words to buf[16*cta..]: %globaltimer stamps (ns) 0 start, 1 setup done, 2 weights resident, 3 last TMA
 * issued, 4 last MMA issued, 5 first accumulator ready, 6 first epilogue done, 7 all done; SM-cycle totals 8 producer
 * waiting for a free stage, 9 MMA waiting for data, 10 MMA waiting for a drained accumulator, 11 epilogue waiting for
 * an accumulator, 12 epilogue busy.  rfk_coupling_nn_fused uses the same buffer when capacity_ctas >= 2 x the SM count: 16
 * SM-cycle counters per CTA (MMA thread: GEMM1 section, waits for TMA data / the drained tap accumulator / the first and later
 * chunks of h1 and h2; epilogue warp 0: waits for the three accumulators, busy in each epilogue) followed, after all
 * counter blocks, by 16 clock stamps of one tile per CTA (tools/nn_fused_timeline.py prints both).  NULL switches it off
 * (the default). */
int rfk_debug_set_timeline(unsigned long long* buf, long long capacity_ctas);

/* ConvLSTM cell update (Utils/modules.py:369-377) from a pixel-major gate buffer cc[(b*HW+p)*cc_ld + g*Hc + ch]
 * (g in i,f,o,g; natural weight-row order) plus bias [4*Hc] (nullable).  c_prev (nullable = 0), h_out, c_next: f32 NCHW
 * with batch strides; peep nullable [3,Hc,HW]; h_nhwc (nullable): bf16 copy of h' at channel h_off of an NHWC buffer
 * with row stride h_ld; zero_cc != 0 clears the gate buffer behind the read (ready for the next split-K step). */
int rfk_convlstm_pointwise_ws(float* cc, int cc_ld, const float* bias, const float* c_prev, long long c_prev_bstride,
                              const float* peep, float* h_out, long long h_bstride, float* c_next,
                              long long c_next_bstride, void* h_nhwc, int h_off, int h_ld,
                              int B, int Hc, int HW, int zero_cc, void* stream);

/* logdet[b] += *addend  (device scalar; the parameter-only log-det terms of ActNorm / InvConv) */
int rfk_add_scalar(float* logdet, const float* addend, float alpha, int B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RFK_H_ */
