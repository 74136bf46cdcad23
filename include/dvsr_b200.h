/*
 * include/dvsr_b200.h -- C ABI of libdvsr_b200.so, the B200 (sm_100a) kernel library behind the
 * DynaVSR hot path (EDVR forward/backward + MAML inner step).
 *
 * Drop-in boundary.  The reference's native surface is the pybind11 module
 * `deform_conv_cuda` (codes/models/archs/dcn/src/deform_conv_cuda.cpp:681-695); everything else on
 * the path is reached through PyTorch ops (cuDNN convs, pooling, interpolate, optimiser steps).
 * This header replaces both with plain-pointer entry points:
 *
 *   dvsr_mdcn_forward_nchw   <- modulated_deform_conv_cuda_forward   (deform_conv_cuda.cpp:486-564)
 *   dvsr_mdcn_backward_nchw  <- modulated_deform_conv_cuda_backward  (deform_conv_cuda.cpp:566-679)
 *   dvsr_conv_fprop / dvsr_conv_wgrad / dvsr_mdcn_bwd_data / dvsr_pack_weights
 *                            <- nn.Conv2d / nn.Conv3d call sites EDVR_arch.py:68-90,141-159,224-249,
 *                               arch_util.py:42-43, LRimg_estimator.py:77-88 and the fused NHWC DCN
 *   dvsr_upsample_* / dvsr_pool_* / dvsr_tsa_* / dvsr_pad_*  <- EDVR_arch.py:107-120,166-202,311;
 *                               LRimg_estimator.py:75-76
 *   dvsr_loss_* / dvsr_update_*  <- Video_base_model.py:39-50,191-195; test_dynavsr.py:223-231,277
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless stated otherwise; the caller owns all memory,
 *     including workspaces (nothing is allocated or freed inside the library);
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it, never synchronised,
 *     so every entry point is CUDA-graph capturable;
 *   - return value 0 = success, <0 = error (DVSR_ERR_*); dvsr_last_error() gives the message of the
 *     last failing call on the calling thread;
 *   - activations inside the library are NHWC ("pixel rows of channels"); the *_nchw entry points
 *     accept the reference's NCHW-contiguous tensors (deform_conv_cuda.cpp:493-494).
 */
#ifndef DVSR_B200_H_
#define DVSR_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DVSR_OK 0
#define DVSR_ERR_INVALID (-1)     /* bad shape / argument                      */
#define DVSR_ERR_CUDA (-2)        /* a CUDA runtime / launch error             */
#define DVSR_ERR_UNSUPPORTED (-3) /* valid but not implemented (e.g. groups>1) */

#define DVSR_ACT_NONE 0
#define DVSR_ACT_RELU 1
#define DVSR_ACT_LRELU 2
#define DVSR_ACT_SIGMOID_SPLIT 3 /* sigmoid on output channels >= sig_split, identity below */

#define DVSR_MAX_SEG 5

/* Operand precision of the resident-weight tensor-core convolution (conv_tc2.cu), per launch (dvsr_policy.precision):
 *   DVSR_PREC_BF16X3 (0, default) activations and weights split into bf16 hi + lo parts, the three significant products
 *                    accumulated in fp32 (~5e-6 per layer; weights packed with mode 9 / 10);
 *   DVSR_PREC_TF32   single-pass TF32 with operands rounded to nearest (~3e-4 per layer; pack mode 5 / 6);
 *   DVSR_PREC_BF16   the BF16X3 layouts and packs, only the x_hi.w_hi product issued: plain bf16 operands, fp32
 *                    accumulation (~3e-3 per layer, a third of the tensor-core work) -- for the inner adaptation steps,
 *                    whose errors reach the frame attenuated (profiles/r1_precision_study.md). */
#define DVSR_PREC_BF16X3 0
#define DVSR_PREC_TF32 1
#define DVSR_PREC_BF16 2

/* Launch policy of ONE call.  Zero-initialised = library defaults; nothing about a launch lives in process-global state, so
 * several engines (adapt.AdaptationPool pipelines, two pools with different settings) can share a process.
 *   cta_budget  CTAs one launch of a persistent kernel (conv_tc2, conv_wgrad_tc, mdcn_tc) may use; 0 = every SM of the
 *               current device (queried, not assumed).  With P frames in flight on P streams ~SMs / P lets their launches run
 *               side by side.
 *   min_tiles   conv_tc2: at least this many 128-pixel tiles per persistent CTA (0 = 1: lowest single-launch latency; larger
 *               amortises the resident weights and leaves SMs to the other frames in flight).
 *   min_chunks  conv_wgrad_tc split-K: at least this many pixel chunks per CTA (0 = 4).
 *   precision   DVSR_PREC_* of conv_tc2.
 *   mdcn_staged DCN forward gather: 0 = staged shared-memory window when the launch has >= 2 tiles per SM, 1 = always the
 *               direct-gather kernel, 2 = always the staged kernel (A/B measurements, tests). */
typedef struct dvsr_policy {
    int cta_budget, min_tiles, min_chunks, precision, mdcn_staged;
} dvsr_policy;

/* One input "segment" of a convolution's K dimension: a group of C input channels read from one
 * NHWC tensor.  torch.cat([a, b], 1) feeding a conv is two segments; the 5-frame 1x1 fusion convs
 * of TSA are five; a Conv3d is one segment per temporal tap.  Output image n reads source image
 *   (n / T) * Tsrc + (t_fixed >= 0 ? t_fixed : n % T + dt)
 * (T = Tsrc = 1, dt = 0, t_fixed = -1 for an ordinary conv); a frame index n % T + dt outside [0, Tsrc)
 * contributes zeros (temporal taps of a Conv3d data gradient). */
typedef struct dvsr_conv_seg {
    const float* ptr;
    int C;               /* channels in this segment                                   */
    int pix_stride;      /* floats between consecutive pixels (>= C; allows channel slices) */
    long long img_stride;/* floats between consecutive source images                   */
    int T, Tsrc, dt, t_fixed;
} dvsr_conv_seg;

typedef struct dvsr_conv_desc {
    int N, H, W;         /* output images; SOURCE spatial size                         */
    int Ho, Wo;          /* produced spatial size                                      */
    int KH, KW, stride, pad, dil;
    int transposed;      /* 0: src = o*stride - pad + k*dil.  1 (data gradient): src = (o + pad - k*dil)/stride when divisible */
    int wshare;          /* 1: every segment uses the same packed weight rows (data gradient of a broadcast input) */
    int nseg;
    dvsr_conv_seg seg[DVSR_MAX_SEG];
    int Co;
    /* modulated deformable sampling of segment 0 (0 = plain convolution) */
    int deform, dg;
    const float* offset; int off_pix_stride;   /* [pix][(g*KH*KW + k)*2 + {dy,dx}]  */
    const float* mask;   int mask_pix_stride;  /* [pix][g*KH*KW + k]                 */
    /* epilogue: v = acc + bias; v = act(v); v += res; (PixelShuffle(2) store if shuffle == 2) */
    const float* bias;
    int act; float slope; int sig_split;
    const float* res; int res_pix_stride;
    int shuffle;
    int accumulate;      /* y += v instead of y = v                                    */
    float* y; int y_pix_stride;
    /* output placement (tensor-core path only; 0 = dense): produced pixel (oy, ox) is stored at
     * (oy*out_step + out_off_y, ox*out_step + out_off_x) of an out_H x out_W image and skipped when that falls
     * outside -- one parity class of the data gradient of a stride-2 convolution. */
    int out_step, out_off_y, out_off_x, out_H, out_W;
    dvsr_policy policy;  /* launch policy of this call (all zero = defaults) */
} dvsr_conv_desc;

/* Where element (co, seg s, ci, tap) of a PyTorch-layout weight lives:
 *   co * co_stride + seg_base[s] + ci * ci_stride + tap
 * Conv2d [Co, Cin, KH, KW] with cat-segments: co_stride = Cin*KH*KW, ci_stride = KH*KW,
 * seg_base[s] = (channel offset of s) * KH*KW.  Conv3d [Co, Ci, KT, KH, KW] with one segment per kt:
 * ci_stride = KT*KH*KW, seg_base[kt] = kt*KH*KW.
 *
 * Interleaved channels (ci_bits > 0): kernel channel ci = (hi << ci_bits) | lo addresses weight element
 *   co * co_stride + seg_base[s] + lo * ci_stride + hi * ci_hi_stride + tap      and exists only for lo < ci_lo_valid.
 * Used by the RGB Conv3d of MFDN (LRimg_estimator.py:77): its three temporal taps are folded into the channel axis of ONE
 * tensor-core conv over a [.., 12] tensor (channel 4*kt + c, c < 3), ci_stride = KT*KH*KW, ci_hi_stride = KH*KW. */
typedef struct dvsr_wlayout {
    long long co_stride, ci_stride;
    long long seg_base[DVSR_MAX_SEG];
    int seg_C[DVSR_MAX_SEG];
    int nseg, taps, Co;
    int ci_bits, ci_lo_valid;
    long long ci_hi_stride;
} dvsr_wlayout;

/* One weight-packing job (pack_table.cu).  mode 0/1: CUDA-core layouts; 2/3/4: streaming tcgen05 layouts (a0 = padded
 * rows; mode 4: parity-restricted taps, a1 = KWf, a2 = KWs, a3 = 2a+b); 5/6: resident-weight tcgen05 layouts (a0 =
 * weight blocks per 64-channel output group; segments [seg, seg_hi)); 7-10: BF16x3 variants of 5/6.  block_start is only used by dvsr_pack_table. */
typedef struct dvsr_pack_job {
    const float* w;
    float* wp;
    dvsr_wlayout wl;
    int mode, seg, seg_hi;
    int a0, a1, a2, a3;
    long long total;        /* floats in wp */
    long long block_start;  /* first 256-thread block of this job inside a table launch */
} dvsr_pack_job;

const char* dvsr_last_error(void);
int dvsr_version(void);
/* number of SMs of the current device (what cta_budget = 0 resolves to) */
int dvsr_sm_count(void);

/* ---- convolution family (conv_simt.cu; tensor-core fast path in conv_tc.cu) ---------------------- */
/* Packed layouts.  mode 0 (forward):  wp[k][co],  k = kofs(s) + tap*C_s + ci.
 *                  mode 1 (data gradient of segment `seg`): wp[tap*Co + co][ci]. */
int dvsr_pack_weights(const float* w, float* wp, const dvsr_wlayout* wl, int mode, int seg, void* stream);
int dvsr_pack_job_run(const dvsr_pack_job* job, void* stream);
/* Re-pack every job of a DEVICE-resident table in one launch (after a fused parameter update). */
int dvsr_pack_table(const dvsr_pack_job* table_dev, int n_jobs, long long total_blocks, void* stream);
/* Copy every packed buffer of the table to (to_packs = 0) / from (to_packs = 1) one arena of total_blocks * 256 floats
 * (job j lives at arena + block_start * 256): restoring the packs of the meta-weights for the next frame
 * (test_dynavsr.py:208) is then one copy launch instead of a re-pack. */
int dvsr_pack_table_copy(const dvsr_pack_job* table_dev, int n_jobs, long long total_blocks, float* arena, int to_packs,
                         void* stream);
/* y = epilogue(conv(x, wp)); wp packed with mode 0 (or mode 1 together with d->transposed). */
int dvsr_conv_fprop(const dvsr_conv_desc* d, const float* wp, void* stream);
/* gw[PyTorch layout] += sum_pix A[pix][k] * gy[pix][co]; A described by d (deformable or plain). */
int dvsr_conv_wgrad(const dvsr_conv_desc* d, const float* gy, int gy_pix_stride, float* gw,
                    const dvsr_wlayout* wl, void* stream);
/* tcgen05 / TMEM / TMA implicit GEMM (conv_tc.cu): stride-1 convolutions and their data gradients with
 * 16 <= Co <= 256; TF32 inputs, fp32 accumulation.  Weights packed by dvsr_pack_weights_tc:
 * mode 2 = forward, mode 3 = data gradient of segment `seg`; dvsr_conv_tc_packed_floats gives the size. */
int dvsr_conv_tc_supported(const dvsr_conv_desc* d);
long long dvsr_conv_tc_packed_floats(const dvsr_wlayout* wl, int mode, int seg);
int dvsr_pack_weights_tc(const float* w, float* wp, const dvsr_wlayout* wl, int mode, int seg, void* stream);
/* mode-3 packing restricted to the taps (a + 2t, b + 2u) of a KHf x KWf kernel (parity class (a, b) of a stride-2
 * data gradient); the packed tap index is t * KWs + u with KHs = ceil((KHf-a)/2), KWs = ceil((KWf-b)/2). */
int dvsr_pack_weights_tc_parity(const float* w, float* wp, const dvsr_wlayout* wl, int seg, int KHf, int KWf, int a,
                                int b, void* stream);
int dvsr_conv_tc_fprop(const dvsr_conv_desc* d, const float* wp, void* stream);
/* Persistent resident-weight variant (conv_tc2.cu): stride-1 convs whose packed weights for 64 output channels fit
 * in shared memory (<= 18 blocks of (tap, 32 input channels)); halo reuse across taps.  accum_in (optional) is added
 * before bias / activation, which lets the host split the K dimension of wider inputs over several launches.
 * Weights: dvsr_pack_weights_tc2 mode 5 / 9 (forward, segments [seg_lo, seg_hi)), mode 6 / 10 (data gradient of seg_lo);
 * 5 / 6 = TF32 rows (8 KiB blocks per (32-channel chunk, tap)); 9 / 10 = BF16x3 (16 KiB blocks per (64-channel pair, tap):
 * rows 0-63 hold the bf16 hi parts of 64 output channels, rows 64-127 the lo parts, so ONE N = 128 MMA gives x_hi.w_hi and
 * x_hi.w_lo); modes 7 / 8 are the [hi | lo]-per-row variant read by dvsr_mdcn_tc_fprop.  d->policy.precision selects the
 * operand precision and must match the pack mode. */
int dvsr_conv_tc2_supported(const dvsr_conv_desc* d);
long long dvsr_conv_tc2_packed_floats(const dvsr_wlayout* wl, int mode, int seg_lo, int seg_hi);
int dvsr_pack_weights_tc2(const float* w, float* wp, const dvsr_wlayout* wl, int mode, int seg_lo, int seg_hi, void* stream);
int dvsr_conv_tc2_fprop(const dvsr_conv_desc* d, const float* wp, const float* accum_in, int accum_pix_stride, void* stream);
/* debugging aid: 8 x 64 clock64 stamps of CTA (0,0) (producer / rounding / MMA / epilogue events); NULL = off */
int dvsr_conv_tc2_set_trace(long long* dev_buffer);
/* Tensor-core weight gradient of segment `seg` of a stride-1 convolution (conv_wgrad_tc.cu): both operands are
 * consumed MN-major straight from the NHWC tensors, x through one halo tile per pixel chunk. */
int dvsr_conv_wgrad_tc_supported(const dvsr_conv_desc* d, int seg);
int dvsr_conv_wgrad_tc(const dvsr_conv_desc* d, int seg, const float* gy, int gy_pix_stride, float* gw,
                       const dvsr_wlayout* wl, void* stream);
/* Small-Cout direct convolution (conv_last, 64 -> 3): one thread per output pixel. */
int dvsr_conv_small_co(const dvsr_conv_desc* d, const float* wp, void* stream);

/* ---- modulated deformable convolution ------------------------------------------------------------ */
/* Gradients w.r.t. the sampled input, the offsets and the (post-sigmoid) mask.  `d` describes the
 * forward op (deform = 1); wd is the mode-1 packed weight.  gx must be zero-filled (or hold a
 * gradient to accumulate into): contributions are added with red.global.add. */
int dvsr_mdcn_bwd_data(const dvsr_conv_desc* d, const float* gy, int gy_pix_stride, const float* wd,
                       float* gx, int gx_pix_stride, float* goff, int goff_pix_stride,
                       float* gmask, int gmask_pix_stride, void* stream);

/* Tensor-core backward (mdcn_bwd_tc.cu) -- replaces modulated_deform_conv_cuda_backward's GEMMs + col2im / col2im_coord kernels
 * (deform_conv_cuda.cpp:617-666, deform_conv_cuda_kernel.cu:634-766) with ONE kernel: grad_col = gy . W^T as tcgen05 MMAs into
 * TMEM, consumed in place by the gather warps (grad_offset / grad_mask stores, grad_input red.global.add.v4), which also build the
 * modulated samples transposed in shared memory for the weight-gradient MMAs (accumulators resident in TMEM, added to gw at the
 * end).  EDVR geometry: 64 -> 64 channels, 8 deformable groups, <= 9 taps.  wp10 = dvsr_pack_weights_tc2 mode 10 of the weight.
 * gx (may be NULL) must be zero-filled or hold a gradient to accumulate into; goff / gmask (may be NULL) are overwritten;
 * gw (may be NULL; PyTorch layout through wl) is accumulated into. */
int dvsr_mdcn_bwd_tc_supported(const dvsr_conv_desc* d);
int dvsr_mdcn_bwd_tc(const dvsr_conv_desc* d, const float* gy, int gy_pix_stride, const float* wp10, float* gx,
                     int gx_pix_stride, float* goff, int goff_pix_stride, float* gmask, int gmask_pix_stride, float* gw,
                     const dvsr_wlayout* wl, void* stream);

/* Tensor-core forward (mdcn_tc.cu): gather warps build the modulated bilinear samples as BF16x3 operand rows in shared
 * memory, resident weights (pack mode 7), persistent CTAs.  8 channels per deformable group, C*KH*KW <= 576, Co <= 64. */
int dvsr_mdcn_tc_supported(const dvsr_conv_desc* d);
int dvsr_mdcn_tc_fprop(const dvsr_conv_desc* d, const float* wp, void* stream);
/* 3x3 / stride-1 / pad-1 launches (every DCN of EDVR) stage a 26 x 18-pixel input window per 16 x 8 output tile and 32-channel
 * chunk in shared memory with one TMA box and gather the 36 corners per pixel from there (global fallback for corners a large
 * offset pushes outside the window) when the launch has at least two tiles per SM (d->policy.mdcn_staged overrides). */

/* Reference operator boundary, NCHW fp32: the tensors and the (h, w) geometry pairs of modulated_deform_conv_cuda_forward /
 * _backward (deform_conv_cuda.cpp:486-492, :566-573).  groups must be 1 and each pair isotropic (stride_h == stride_w, ...):
 * no YML / Python call site of the reference uses anything else (deform_conv.py:104-106 passes each value twice);
 * other requests return DVSR_ERR_UNSUPPORTED.  Workspace: dvsr_mdcn_workspace_bytes() bytes (negative = error code). */
long long dvsr_mdcn_workspace_bytes(int B, int C, int H, int W, int Co, int kh, int kw, int stride_h, int stride_w,
                                    int pad_h, int pad_w, int dil_h, int dil_w, int dg, int backward);
int dvsr_mdcn_forward_nchw(const float* x, const float* offset, const float* mask, const float* weight,
                           const float* bias, float* y, int B, int C, int H, int W, int Co, int kh,
                           int kw, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                           int dg, void* workspace, long long workspace_bytes, void* stream);
int dvsr_mdcn_backward_nchw(const float* x, const float* offset, const float* mask, const float* weight,
                            const float* gy, float* gx, float* goffset, float* gmask, float* gweight,
                            float* gbias, int B, int C, int H, int W, int Co, int kh, int kw, int stride_h,
                            int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int groups, int dg,
                            void* workspace, long long workspace_bytes, void* stream);

/* ---- layout, resampling, pooling, padding (elementwise.cu) --------------------------------------- */
int dvsr_nchw_to_nhwc(const float* x, float* y, int N, int C, int H, int W, void* stream);
int dvsr_nhwc_to_nchw(const float* x, float* y, int N, int C, int H, int W, void* stream);
/* bilinear, align_corners=False, integer scale; y = mul * up(x) (+ y if accumulate) */
int dvsr_upsample_bilinear(const float* x, float* y, int N, int H, int W, int C, int scale, float mul,
                           int accumulate, void* stream);
int dvsr_upsample_bilinear_bwd(const float* gy, float* gx, int N, int H, int W, int C, int scale,
                               float mul, void* stream);
/* y[..., 0:C] = maxpool3x3s2p1(x), y[..., C:2C] = avgpool3x3s2p1(x) (count_include_pad) */
int dvsr_pool_maxavg(const float* x, float* y, int N, int H, int W, int C, void* stream);
int dvsr_pool_maxavg_bwd(const float* x, const float* gy, float* gx, int N, int H, int W, int C, void* stream);
/* mode 0 = reflect, 1 = replicate; pads H and W by p (and, if padT, the T axis of [B,T,H,W,C] by 1, replicate) */
int dvsr_pad2d(const float* x, float* y, int N, int H, int W, int C, int p, int mode, void* stream);
int dvsr_pad2d_bwd(const float* gy, float* gx, int N, int H, int W, int C, int p, int mode, void* stream);
int dvsr_pad3d_replicate(const float* x, float* y, int B, int T, int H, int W, int C, void* stream);
int dvsr_pad3d_replicate_bwd(const float* gy, float* gx, int B, int T, int H, int W, int C, void* stream);
/* Replication-padded clip with the KT = 3 temporal taps folded into channels: x [B*T, H, W, C] (C <= 4) ->
 * y [B*T, H+2, W+2, 12], y[.., 4*kt + c] = x[b, clamp(t + kt - 1), clamp(h - 1), clamp(w - 1), c], zero for c >= C.
 * A Conv3d(C -> Co, 3^3) over the padded clip is then ONE 3x3 conv over y (LRimg_estimator.py:75-77,100-102). */
int dvsr_tcat_pad3(const float* x, float* y, int B, int T, int H, int W, int C, void* stream);
/* Blur-and-subsample degradation HR -> LR on NHWC frames x [T, H, W, C] (C <= 4), Degradation.apply of
 * random_kernel_generator.py:83-130: reflection pad L/2, depthwise L x L kernel k (already centre-shifted; [Tk, L, L]),
 * stride `scale`; kmode 0 = one kernel, 1 = kernel t per frame, 2 = kernel (t - 1) mod Tk (the T == Tk + 2 rule, :113-116);
 * quantize != 0 applies round(y * 255) / 255 (vsrbase.py:188).  y: [T, Ho, Wo, C], Ho = (H + 2*(L/2) - L) / scale + 1. */
int dvsr_degrade(const float* x, const float* k, float* y, int T, int H, int W, int C, int L, int scale, int Tk, int kmode,
                 int quantize, void* stream);
/* per-image per-channel spatial mean: m[n][c]; y = x - m (sign=-1) or x + m (sign=+1) */
int dvsr_spatial_mean(const float* x, float* m, int N, int HW, int C, void* stream);
int dvsr_add_channel_bias(const float* x, const float* m, float* y, int N, int HW, int C, float sign, void* stream);

/* Result frame -> 8-bit image, the device half of utils/util.py:112-142 (tensor2img) + :262-269 (calculate_psnr):
 * out[i] = uint8(rint(clamp(x, 0, 1) * 255)) for an [npix][C] channels-last frame (= the HWC image the reference writes),
 * channel order kept (reverse = 0, 'rgb') or reversed (reverse = 1, the 'bgr' default of tensor2img).  With ref (an image
 * of the same layout) and sse: *sse += sum (out - ref)^2, exact (PSNR = 20 log10(255 / sqrt(sse / (npix * C)))).
 * out must be 4-byte aligned; the caller zeroes *sse. */
int dvsr_frame_to_u8(const float* x, unsigned char* out, const unsigned char* ref, unsigned long long* sse,
                     long long npix, int C, int reverse, void* stream);

/* gpre = gy * act'(y - res) [+ un-PixelShuffle]; gbias[c] += sum_pix gpre[pix][c] (if gbias).  In-place allowed
 * when shuffle == 0.  y is the saved forward OUTPUT (relu / lrelu / sigmoid-split derivatives are functions of
 * it); res (optional) is the residual that the forward epilogue added AFTER the activation. */
int dvsr_act_bwd(const float* gy, const float* y, const float* res, float* gpre, float* gbias, long long npix,
                 int C, int act, float slope, int sig_split, int shuffle, int Ho, int Wo, void* stream);

/* ---- TSA fusion (tsa.cu) -------------------------------------------------------------------------- */
/* temporal attention (EDVR_arch.py:166-176): prob[n][f][pix] = sigmoid(sum_c emb[n][f][pix][c]*emb_ref[n][pix][c]);
 * out[n][pix][f*C + c] = aligned[n][f][pix][c] * prob  (pixel-major, F*C channels: the input of the 1x1 fusion convs) */
int dvsr_tsa_temporal(const float* aligned, const float* emb, const float* emb_ref, float* prob, float* out,
                      int B, int F, long long HW, int C, void* stream);
int dvsr_tsa_temporal_bwd(const float* aligned, const float* emb, const float* emb_ref, const float* prob,
                          const float* gout, float* galigned, float* gemb, float* gemb_ref,
                          int B, int F, long long HW, int C, void* stream);
/* out = fea * sigmoid(att) * 2 + att_add (EDVR_arch.py:200-202) */
int dvsr_tsa_combine(const float* fea, const float* att, const float* att_add, float* out, long long n, void* stream);
int dvsr_tsa_combine_bwd(const float* fea, const float* att, const float* gout, float* gfea, float* gatt,
                         long long n, void* stream);

/* ---- losses and the fused inner update (update.cu) ------------------------------------------------ */
#define DVSR_LOSS_L1 0
#define DVSR_LOSS_L2 1
#define DVSR_LOSS_CB 2
#define DVSR_LOSS_HUBER 3 /* delta passed in `eps` (loss.py:5-17, default 1e-2) */
/* loss[0] (+)= weight * mean(f(a - b)).  `loss` must be zeroed by the caller when accumulate == 0 is not
 * wanted; ga (optional) = weight * f'(a-b) / n  (caller multiplies by the upstream scalar). */
int dvsr_loss_fwd(const float* a, const float* b, float* loss, float* ga, long long n, int kind, float weight,
                  float eps, void* stream);
/* y = x * s[0] (device scalar) */
int dvsr_scale_by_device_scalar(const float* x, const float* s, float* y, long long n, void* stream);
/* p[i] -= lr(i) * (g[i] + wd * p[i]), lr(i) = lr0 for i < split else lr1 (two param groups: test_dynavsr.py:213-231,
 * train_dynavsr.py:335-344).  One launch for the whole flat EDVR+MFDN parameter buffer. */
int dvsr_update_sgd(float* p, const float* g, long long n, long long split, float lr0, float lr1, float wd, void* stream);
/* torch.optim.Adam semantics (L2 weight decay wd folded into the gradient, no amsgrad; Video_base_model.py:128-135);
 * bias corrections bc1 = 1-b1^t, bc2 = 1-b2^t. */
int dvsr_update_adam(float* p, const float* g, float* m, float* v, long long n, long long split, float lr0,
                     float lr1, float b1, float b2, float eps, float bc1, float bc2, float wd, void* stream);
/* Meta-training exchange fused with the outer update (train_dynavsr.py:438): grads_dev is a DEVICE array of n_peers pointers to
 * the ranks' flat meta-gradient buffers (peer-mapped symmetric memory, NVLink), grad_offset the element offset of the gradient
 * inside each buffer.  p -= update(scale * sum_r grads[r][i]) with Adam (adam != 0; torch.optim.Adam semantics as
 * dvsr_update_adam) or SGD.  The caller barriers all ranks before (gradients complete) and after (buffers reusable). */
int dvsr_update_peers(float* p, const float* const* grads_dev, int n_peers, long long grad_offset, float scale, float* m, float* v,
                      long long n, long long split, float lr0, float lr1, float b1, float b2, float eps, float bc1, float bc2,
                      float wd, int adam, void* stream);
/* The same exchange in reduce-scatter form (scales with the rank count): rank `rank` reduces slice `rank` of the flat buffer
 * ((n / 4 + n_peers - 1) / n_peers float4 elements) over all ranks' buffers bufs_dev[r] + buf_offset, updates that slice of the
 * meta-weights from p (read only) with its own slice of the moments, and writes the UPDATED WEIGHTS of the slice into the same
 * slice of every rank's buffer.  After the caller's closing barrier each buffer holds the complete new weights (copy them over p).
 * Optimiser state is sharded: m / v are only valid for the caller's slice. */
int dvsr_update_peers_sliced(const float* p, float* const* bufs_dev, int n_peers, int rank, long long buf_offset, float scale, float* m,
                             float* v, long long n, long long split, float lr0, float lr1, float b1, float b2, float eps, float bc1,
                             float bc2, float wd, int adam, void* stream);
/* sum |x| over channels [c0, c1) of an NHWC tensor -> out[0] (+=); the `offset_mean > 100` check of
 * deform_conv.py:285-287 without a host sync per call. */
int dvsr_abs_sum(const float* x, float* out, long long npix, int pix_stride, int c0, int c1, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVSR_B200_H_ */
