/*
 * btsbot_b200 -- C ABI of the B200 (sm_100a) alert-scoring hot path.
 *
 * The reference (nabeelre/BTSbot) has no FFI: its seam is Python (`btsbot/architectures.py`
 * classes calling `timm.create_model`, `btsbot/alert_utils.py` numpy helpers).  Each entry point
 * below names the reference code whose arithmetic it replaces (file:line under /root/reference).
 * The Python package `btsbot_b200` binds these with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - plain pointers + sizes, no torch types; every pointer is DEVICE memory owned by the caller
 *    unless a parameter is documented as host memory;
 *  - the library never allocates or frees device memory, never synchronises the device and keeps no
 *    pointer after the call returns; work is enqueued on `stream` (a cudaStream_t passed as void*);
 *  - activations between kernels are NHWC "pixel rows": a [B,H,W,C] map is the row-major matrix
 *    [B*H*W, C]; API-facing images stay NCHW float32 like the reference's tensors;
 *  - return value: BTSB_OK (0) or a negative BTSB_E* code; btsb_last_error_string() describes the
 *    last failure on the calling thread.  No exceptions or abort() cross this boundary;
 *  - dtype codes: BTSB_F32 / BTSB_BF16 / BTSB_F64 (arithmetic always accumulates in fp32, the
 *    preprocessing norms in fp64);
 *  - there is no CPU path: on a device that is not compute capability 10.x every compute entry
 *    returns BTSB_EARCH.
 */
#ifndef BTSBOT_B200_H
#define BTSBOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BTSB_OK 0
#define BTSB_EINVAL (-1)  /* bad shape / dtype / alignment / null pointer */
#define BTSB_EARCH (-2)   /* device is not sm_100 */
#define BTSB_ECUDA (-3)   /* CUDA runtime / driver error, see btsb_last_error_string */

#define BTSB_F32 0
#define BTSB_BF16 1
#define BTSB_F64 2
/* bf16 compute (MMA operands, LayerNorm outputs, hidden activations: bf16) with the call's RESIDUAL-STREAM tensor --
 * named per function below -- stored as IEEE fp16 (same bytes, 11 significand bits; stores saturate). */
#define BTSB_BF16_XF16 3

/* activation codes (metadata branch / heads) */
#define BTSB_ACT_NONE 0
#define BTSB_ACT_GELU 1 /* exact erf GELU (torch.nn.GELU default) */
#define BTSB_ACT_RELU 2

/* GEMM epilogues */
#define BTSB_EPI_BIAS 0      /* out = acc + bias[n] */
#define BTSB_EPI_BIAS_GELU 1 /* out = gelu(acc + bias[n]) */
#define BTSB_EPI_SCALE_RES 2 /* out = res[m,n] + gamma[n] * (acc + bias[n]) */
#define BTSB_EPI_BIAS_SILU 3 /* out = silu(acc + bias[n])   (MaxViT MBConv: 1x1 conv + folded BatchNorm + SiLU) */

int btsb_version(void);
const char* btsb_last_error_string(void);
/* 0 when the current device can run the library (compute capability 10.x), else BTSB_EARCH. */
int btsb_device_ok(void);
/* number of kernels this library has launched in this process (all threads); bench.py's gpu_launches. */
uint64_t btsb_launch_count(void);

/* ---- host-side marshalling for the end-to-end scoring call (no CUDA work; host pointers): float32 -> bf16, round to
 * nearest even (NaN stays NaN), on a pool of host threads (threads <= 0: one per hardware thread, at most 32).  The bf16
 * mode's first use of a pixel is its bf16 rounding, so packing on the host halves the PCIe bytes of a scoring step with
 * bit-identical logits; btsb_preprocess_crop_norm accepts the packed rows (in_dtype BTSB_BF16, cast + transpose path).
 * Replaces nothing in the reference (its triplets reach the GPU as fp32, inference_example.py:62-64). */
int btsb_host_pack_bf16(const float* src, uint16_t* dst, int64_t n, int threads);

/* ---- K1: array preparation -------------------------------------------------------------------
 * crop_norm: replaces alert_utils.crop_triplets / crop_norm_cutout (alert_utils.py:54-107) fused with the
 * cast + NHWC->NCHW transpose every caller performs next (inference_example.py:62-64, train.py:139-155,
 * val.py:92-94).  in: [n,63,63,3] HWC, in_dtype F32 or F64 (or BF16 rows from btsb_host_pack_bf16: full 63x63, no normalisation, NCHW only).  out: [n,3,s,s] NCHW float32.
 * margin = (63-s)/2 (floor).  normalize!=0: each cutout is divided by the L2 norm of its cropped window
 * (norm accumulated in fp64; for F64 input the quotient is formed in fp64 and rounded once to fp32, which
 * is what `.astype(np.float32)` does to the reference's float64 result).  normalize==0: crop/cast/transpose only.
 * out_hwc!=0 keeps the reference's [n,s,s,3] HWC layout (the drop-in return value of crop_triplets).
 */
int btsb_preprocess_crop_norm(const void* in, int in_dtype, int64_t n, int crop_to_size, int normalize,
                              int out_hwc, float* out, void* stream);

/* pad_norm: numeric tail of alert_utils.make_triplet (alert_utils.py:147-193) for n alerts.
 * stamps: [n,3,63*63] float32, stamp (a,c) stored densely row-major with its own width at the start of its
 * slot; hw: [n,3,2] int32 (rows, cols), each in [1,63].  Per cutout, in order science/template/difference:
 * nanmedian==+-inf -> drop; NaN->0, +-inf->+-FLT_MAX; if normalize and not yet dropped divide by the L2 norm
 * (float32 like numpy: an overflowing norm is +inf, a zero norm yields NaN as in the reference); all-zero
 * -> drop; pad bottom/right to 63x63 with float32(1e-9) AFTER normalisation.
 * out: [n,63,63,3] HWC, out_dtype F64 (reference's storage type) or F32.  drop: [n] uint8.
 */
int btsb_preprocess_pad_norm(const float* stamps, const int32_t* hw, int64_t n, int normalize,
                             void* out, int out_dtype, uint8_t* drop, void* stream);

/* ingest (HOST memory, no device work): batched gunzip + FITS parse of alert stamps -- replaces the per-stamp
 * gzip.open + astropy.io.fits.open loop of alert_utils.make_triplet (alert_utils.py:137-147).  blobs[i] / sizes[i] are
 * the gzipped FITS bytes of stamp i (three per alert: science, template, difference); stamps [n,63*63] float32 and
 * hw [n,2] int32 (rows, cols) are exactly what btsb_preprocess_pad_norm reads.  A pool of `threads` host threads
 * (<= 0: one per hardware thread, at most 32) inflates and parses; BITPIX -32/-64/16/32/8 with BSCALE/BZERO. */
int btsb_ingest_fits_gz(const unsigned char* const* blobs, const int64_t* sizes, int64_t n, float* stamps,
                        int32_t* hw, int threads);

/* training-time batch gather + augmentation (utils.py:44-48, train.py:178-199 RandomHorizontalFlip /
 * RandomVerticalFlip / RandomRightAngleRotation): out[b] = rot90^k(vflip(hflip(images[idx[b]]))), square [3,S,S] fp32
 * images; flags[b] bit0 = hflip, bit1 = vflip, bits2-3 = k counter-clockwise quarter turns (flags NULL = no aug). */
int btsb_augment_gather_f32(const float* images, const int64_t* idx, const uint8_t* flags, int64_t B, int S,
                            float* out, void* stream);

/* ---- K2a: ConvNeXt stem -- timm stem.0 Conv2d(3,C0,k4,s4)+bias and stem.1 LayerNorm2d(eps 1e-6)
 * (called at architectures.py:108,132).  x: [B,3,H,W] NCHW float32.  w: [48,C0] float32 with
 * k = (ci*4+ky)*4+kx (transposed conv weight), bias/ln_w/ln_b: [C0] float32.
 * out: [B*h*w, C0] (h=(H-4)/4+1, w likewise), out_dtype F32 or BF16.  C0 <= 128.
 */
int btsb_convnext_stem_fwd(const float* x, int64_t B, int H, int W, const float* w, const float* bias,
                           const float* ln_w, const float* ln_b, int C0, void* out, int out_dtype, void* stream);

/* ---- K3: depthwise 7x7 (pad 3) + bias + LayerNorm2d -- timm blocks.j.conv_dw + blocks.j.norm.
 * x, out: [B*H*W, C] dtype F32|BF16 (same for both); BF16_XF16: x is the fp16 residual stream, out (the fc1 operand)
 * bf16.  w: [49,C] float32 (k = ky*7+kx), bias/ln_w/ln_b [C].
 */
int btsb_convnext_dwln_fwd(const void* x, int dtype, int64_t B, int H, int W, int C, const float* w,
                           const float* bias, const float* ln_w, const float* ln_b, void* out, void* stream);

/* ---- K5a: downsample prologue -- timm stages.i.downsample.0 LayerNorm2d, emitted directly as the
 * 2x2/s2 patch matrix the conv (downsample.1) consumes as a GEMM.  x: [B*H*W, C]; out: [B*Ho*Wo, 4C] with
 * column (dy*2+dx)*C + c, Ho=(H-2)/2+1 (floor: last odd row/col dropped, as Conv2d does).  dtype F32|BF16 (both
 * tensors); BF16_XF16: x is the fp16 residual stream, out (the GEMM operand) bf16.
 */
int btsb_convnext_lnpatch_fwd(const void* x, int dtype, int64_t B, int H, int W, int C, const float* ln_w,
                              const float* ln_b, void* out, void* stream);

/* ---- head prologue: global average pool + LayerNorm2d + flatten (architectures.py:109-113,136-141,
 * 309-313).  x: [B*HW, C] dtype F32|BF16 -> out [B, C] float32.  ln_w==NULL: pool only.
 */
int btsb_convnext_poolln_fwd(const void* x, int dtype, int64_t B, int HW, int C, const float* ln_w,
                             const float* ln_b, float* out, void* stream);

/* ---- K4 / K5b: pointwise GEMM with fused epilogue -- timm mlp.fc1 (+GELU), mlp.fc2 (*gamma + shortcut),
 * downsample.1.  out[M,N] = epi(A[M,K] . Wt[N,K]^T + bias).  dtype F32: CUDA-core fp32 (the 1e-4 path);
 * dtype BF16: tcgen05/TMEM tensor cores with TMA operand staging, fp32 accumulate, A/Wt/res/out bf16.
 * bias/gamma: [N] float32.  res: [M,N] in `dtype` (EPI_SCALE_RES only).  F32: K % 4 == 0.  BF16: K % 16 == 0,
 * N % 16 == 0, all pointers 16-byte aligned.  BF16_XF16: as BF16 with res and out in the fp16 residual stream.
 */
int btsb_gemm_fwd(const void* A, const void* Wt, const float* bias, const float* gamma, const void* res,
                  void* out, int64_t M, int N, int K, int dtype, int epilogue, void* stream);

/* ---- K2a on tensor cores (bf16): the patch stem as im2col + tcgen05 GEMM whose epilogue applies bias and the
 * LayerNorm2d over the C0 output channels (thread = output row, three passes over the TMEM accumulator).
 * im2col: x [B,3,H,W] fp32 -> patches [B*h*w, 64] bf16 (k = (ci*4+ky)*4+kx, columns 48..63 zero).
 * gemm_ln: out[M,N] = LayerNorm_rows(A[M,K] . Wt[N,K]^T + bias) * ln_w + ln_b, bf16 operands/output, N <= 128. */
int btsb_stem_im2col_bf16(const float* x, void* patches, int64_t B, int H, int W, void* stream);
/* the same stem as ONE kernel: the im2col rows are built in shared memory by producer warps straight from the NCHW fp32
 * image (the [M,64] patch matrix never exists in HBM).  w_pad: [C0,64] bf16 (columns 48..63 zero); out [B*h*w, C0] bf16
 * (out_dtype BTSB_BF16) or fp16 (BTSB_BF16_XF16: the rows open the fp16 residual stream). */
int btsb_stem_fused_fwd(const float* x, int64_t B, int H, int W, const void* w_pad, const float* bias,
                        const float* ln_w, const float* ln_b, void* out, int C0, int out_dtype, void* stream);
int btsb_gemm_ln_fwd(const void* A, const void* Wt, const float* bias, const float* ln_w, const float* ln_b,
                     void* out, int64_t M, int N, int K, void* stream);

/* ---- K4 fused: ONE kernel for fc1 -> GELU -> fc2 -> *gamma -> +shortcut; the 4C hidden activation stays in
 * TMEM / shared memory (timm blocks.j.mlp + gamma + residual).  BF16 only; C a multiple of 16 in [64,160], 256 or 320
 * (ConvNeXt nano/pico stages 0-2, where the hidden tensor would be 4x the activation traffic).
 * y: dw+LN output [M,C]; res: block input [M,C]; W1 [4C,C], W2 [C,4C] bf16 row-major; b1 [4C], b2/gamma [C] f32.
 * dtype BTSB_BF16: res / out bf16; BTSB_BF16_XF16: res / out are the fp16 residual stream (y, W1, W2 stay bf16).
 * out == res (in place) is allowed; C = 256 / 320 then add the update to the rows with a bulk tensor reduction.
 */
int btsb_convnext_mlp_fused_fwd(const void* y, const void* res, const void* W1, const float* b1,
                                const void* W2, const float* b2, const float* gamma, void* out, int64_t M,
                                int C, int dtype, void* stream);

/* ---- K6: metadata branch + fusion head in one kernel (architectures.py:146-164,168-170; um_nn 282-290;
 * image-only heads 109-119; frozen_fusion 357-365).
 * feat: [B,F] (feat_dtype F32|BF16) or NULL (F=0); meta: [B,Mm] float32 or NULL (Mm=0).
 * BatchNorm1d is passed folded: bn_scale = w/sqrt(var+eps), bn_shift = b - mean*bn_scale.
 * Weights are float32 and TRANSPOSED ([in,out], out contiguous): m1t [Mm,m1], m2t [m1,m2],
 * h0t [F+m2 (or F, or m2), c1], h1t [c1,c2], h2 [c2].  meta_act after m1; meta_out_act after m2
 * (GELU for mm_*, NONE for frozen_fusion, RELU for um_nn); head_act after h0 and h1.
 * When Mm>0 and c1==0 the head is just `h2` applied to the meta embedding (um_nn: Linear(m2,1)).
 * h0_init (optional, NULL = off): [B,c1] float32 = h0b + feat . h0t[:F], the image-feature part of the head's first
 * layer pre-computed by the caller on the tensor cores (btsb_gemm_bf16_f32out); `feat` is then not read and the
 * kernel contracts only the m2 embedding rows h0t[F:].
 * logits: [B] float32.
 */
typedef struct {
  const void* feat; int feat_dtype; int F;
  const float* meta; int Mm;
  const float *bn_scale, *bn_shift;
  const float *m1t, *m1b; int m1;
  const float *m2t, *m2b; int m2;
  int meta_act, meta_out_act;
  const float *h0t, *h0b; int c1;
  const float *h1t, *h1b; int c2;
  const float *h2, *h2b;
  int head_act;
  const float* h0_init;
} btsb_head_params;
int btsb_meta_head_fwd(const btsb_head_params* p, int64_t B, float* logits, void* stream);

/* ---- scoring epilogue (train.py:530-538, val.py:153,168, inference_example.py:91):
 * score = sigmoid(logit), label = score > 0.5.  scores/labels may be NULL. */
int btsb_score_epilogue(const float* logits, int64_t B, float* scores, uint8_t* labels, void* stream);

/* ---- K7: training (fp32).  Replaces what autograd + ATen/cuDNN + torch.optim do inside train.py:496-547
 * (forward with saved intermediates, BCEWithLogitsLoss(pos_weight) :211-212,525, backward :526, AdamW.step :527).
 * All tensors float32, row-major, contiguous unless strides are given. */
/* C[M,N] (+)= sum_k A(m,k) B(k,n) with element strides (sam,sak) / (sbk,sbn): covers dgrad (NN) and wgrad (TN, split-K). */
int btsb_gemm_f32_strided(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                          float* C, int64_t M, int64_t N, int64_t K, int accumulate, void* stream);
/* out[n] (+)= sum_m X[m,n] * (Y ? Y[m,n] : 1)  -- bias gradients, d(gamma) */
int btsb_colsum_f32(const float* X, const float* Y, float* out, int64_t M, int N, int accumulate, void* stream);
/* dout == NULL: out = act(pre); else out = dout * act'(pre)   (act: BTSB_ACT_*) */
int btsb_act_f32(const float* pre, const float* dout, float* out, int64_t n, int act, void* stream);
/* out[m,n] = (res ? res[m,n] : 0) + g[n] * X[m,n]   -- layer scale + shortcut, and its backward */
int btsb_colscale_f32(const float* X, const float* g, const float* res, float* out, int64_t M, int N, void* stream);
/* X[m,n] += b[n] (in place) */
int btsb_bias_add_f32(float* X, const float* b, int64_t M, int N, void* stream);
int btsb_layernorm_fwd_f32(const float* u, const float* w, const float* b, float* y, int64_t M, int C, float eps, void* stream);
/* du may be NULL; dw/db are accumulated (+=). */
int btsb_layernorm_bwd_f32(const float* u, const float* w, const float* dy, float* du, float* dw, float* db,
                           int64_t M, int C, float eps, void* stream);
/* depthwise 7x7 pad 3 on NHWC rows; w49 [49,C]; bias may be NULL; flip=1 uses the 180-degree rotated kernel (dgrad). */
int btsb_dwconv7_f32(const float* x, const float* w49, const float* bias, float* out, int64_t B, int H, int W, int C,
                     int flip, void* stream);
/* dw49[k,c] += sum du * shifted x;  dbias[c] += sum du */
int btsb_dwconv7_wgrad_f32(const float* x, const float* du, float* dw49, float* dbias, int64_t B, int H, int W, int C,
                           void* stream);
/* x [B,3,H,W] -> patches [B*ho*wo, 48] (k = (ci*4+ky)*4+kx): the stem conv as a GEMM operand */
int btsb_stem_im2col_f32(const float* x, float* patches, int64_t B, int H, int W, void* stream);
/* reverse=0: rows [B*H*W,C] -> 2x2/s2 patches [B*Ho*Wo,4C]; reverse=1: patches -> rows (dropped pixels get 0) */
int btsb_patch2x2_f32(const float* src, float* dst, int64_t B, int H, int W, int C, int reverse, void* stream);
/* reverse=0: mean over HW, [B*HW,C] -> [B,C]; reverse=1: d[B,C]/HW broadcast to [B*HW,C] */
int btsb_pool_f32(const float* src, float* dst, int64_t B, int HW, int C, int reverse, void* stream);
/* BatchNorm1d in training mode (batch statistics, running stats updated with the unbiased variance like torch) */
int btsb_bn1d_train_fwd_f32(const float* x, const float* w, const float* b, float* run_mean, float* run_var,
                            float momentum, float eps, float* y, float* save_mean, float* save_rstd, int64_t B, int F,
                            void* stream);
int btsb_bn1d_bwd_f32(const float* x, const float* dy, const float* w, const float* save_mean, const float* save_rstd,
                      float* dx, float* dw, float* db, int64_t B, int F, void* stream);
/* inverted dropout; reuse_mask=0 draws mask[i] = hash(seed,i) >= p, reuse_mask=1 applies the stored mask (backward) */
int btsb_dropout_f32(const float* x, float* y, uint8_t* mask, int64_t n, float p, uint64_t seed, int reuse_mask,
                     void* stream);
/* BCEWithLogitsLoss(pos_weight), mean reduction: *loss = mean l_i; dlogits (may be NULL) = dscale * dl/dx */
int btsb_bce_logits_f32(const float* logits, const float* labels, float pos_weight, float* loss, float* dlogits,
                        int64_t B, float dscale, void* stream);
/* fused AdamW over a flat parameter buffer (torch.optim.AdamW semantics; g is multiplied by grad_scale first) */
int btsb_adamw_f32(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float wd, int64_t step, float grad_scale, void* stream);

/* the same update for up to BTSB_ADAMW_BATCH parameter tensors in ONE launch (torch's "foreach"/fused AdamW role):
 * the caller fills p/g/m/v/n[0..count); first_block is scratch written by the call.  All tensors share `step`. */
#define BTSB_ADAMW_BATCH 48
typedef struct {
  float* p[BTSB_ADAMW_BATCH];
  const float* g[BTSB_ADAMW_BATCH];
  float* m[BTSB_ADAMW_BATCH];
  float* v[BTSB_ADAMW_BATCH];
  int64_t n[BTSB_ADAMW_BATCH];
  int32_t first_block[BTSB_ADAMW_BATCH + 1];
  int32_t count;
} btsb_adamw_batch;
int btsb_adamw_multi_f32(btsb_adamw_batch* batch, float lr, float beta1, float beta2, float eps, float wd,
                         int64_t step, float grad_scale, void* stream);
/* CUDA-graph-capturable variants (a captured training step is replayed with the host-side scalars frozen): the AdamW step
 * count and the dropout seed offset are read from a device counter that btsb_counter_add_i64 bumps inside the graph
 * (torch.optim's `capturable=True` idea). */
int btsb_adamw_multi_ctr_f32(btsb_adamw_batch* batch, float lr, float beta1, float beta2, float eps, float wd,
                             const int64_t* step_dev, float grad_scale, void* stream);
int btsb_dropout_ctr_f32(const float* x, float* y, uint8_t* mask, int64_t n, float p, uint64_t seed,
                         const int64_t* counter, int reuse_mask, void* stream);
int btsb_counter_add_i64(int64_t* counter, int64_t v, void* stream);

/* ---- K7 on tensor cores (bf16 mode of the training step; what autocast(bf16) + cuBLAS do under train.py:496-547).
 * The three GEMMs of every Linear / 1x1 conv run on tcgen05 with bf16 operands and fp32 accumulation; results are
 * fp32 so that the residual stream, LayerNorm and the element-wise backward stay in fp32.
 * cast_dual_bf16: one pass over a [M,N] tensor (in / in2 both of in_dtype F32 | BF16) writes the row-major bf16 copy out_rm [M,N] (forward / dgrad
 *   operand) and/or the transposed copy out_t [N,ld] (wgrad operand; ld % 8 == 0, ld >= M), folding in
 *   op 0: v = in;  op 1: v = gelu(in);  op 2: v = in2 * gelu'(in);  then v *= colvec[n] (if colvec) and
 *   colsum[n] += sum_m v (if colsum; bias gradients); auxsum[n] += sum_m in[m,n] * aux[m,n] (if aux, fp32 [M,N]: the
 *   layer-scale gradient sum(dout * v) rides on the pass that scales dout by gamma).  N even.
 * gemm_bf16_f32out: out[M,N] fp32 = A[M,K] . Wt[N,K]^T (+ bias[N] if not NULL); N % 16 == 0, K % 8 == 0.
 * gemm_bf16_wgrad:  out[M,N] fp32 += At[M,K] . Bt[N,K]^T where K is the (huge) activation row count and At / Bt are
 *   transposed copies with row pitch ld; K is split over the SMs and partial tiles are reduced with red.global.add
 *   (the caller zeroes `out`; summation order, hence the last fp32 bits, varies run to run).  N % 16 == 0. */
int btsb_cast_dual_bf16(const void* in, const void* in2, const float* colvec, void* out_rm, void* out_t,
                        float* colsum, int64_t M, int N, int64_t ld, int op, int in_dtype, const float* aux,
                        float* auxsum, void* stream);
int btsb_gemm_bf16_f32out(const void* A, const void* Wt, const float* bias, float* out, int64_t M, int N, int K,
                          void* stream);
int btsb_gemm_bf16_wgrad(const void* At, const void* Bt, int64_t ld, float* out, int M, int N, int64_t K,
                         void* stream);
/* the same reduction straight from the ROW-MAJOR activations (no transposed copies): out[M,N] fp32 += A[K,M]^T . B[K,N],
 * A and B row-major bf16 with leading dimensions M and N -- both operands are fed to UMMA as MN-major (a 64-column x
 * 64-row TMA box with the 128-byte swizzle is the canonical MN-major atom).  M % 8 == 0, N % 16 == 0. */
int btsb_gemm_bf16_wgrad_mn(const void* A, const void* B, float* out, int M, int N, int64_t K, void* stream);

/* ---- MaxViT (timm maxvit_tiny_rw_224 behind btsbot/architectures.py:25-101) -------------
 * The 1x1 convolutions / Linear layers of the MBConv, attention and MLP blocks are btsb_gemm_fwd calls (tcgen05 in
 * bf16) with BatchNorm folded into the weights; the kernels below are everything between those GEMMs.  Activations
 * are NHWC pixel rows [B*H*W, C] in `dtype` (F32 | BF16), math in fp32.  At most 65535 images per call.
 *
 * stem1: F.interpolate(x, (S,S), bilinear, align_corners=False) (architectures.py:44-50,90-96) fused into stem.conv1
 *   (3x3, stride 2, pad 1, no bias) + stem.norm1 (BatchNorm, folded) + SiLU.  x [B,3,Hin,Win] fp32 NCHW ->
 *   out [B*(S/2)*(S/2), C1]; w [27][C1] fp32 (k = (ci*3+ky)*3+kx, BN scale folded in), shift [C1].  S == Hin skips
 *   nothing: the interpolation weights degenerate to the identity, as torch's do.
 * im2col3: 3x3 / stride 1 / pad 1 patch matrix [B*H*W, 9C], column (ky*3+kx)*C + c, for stem.conv2 as a GEMM.
 * avgpool2: MBConv shortcut AvgPool2d(2).
 * dw3: conv2_kxk depthwise 3x3 (stride 1|2, pad 1, no bias) + norm2 (folded) + SiLU; also writes the SE squeeze
 *   pooled[b,c] = mean_{h,w} out (fixed summation order).  w [9][C] fp32 (BN scale folded), shift [C].
 * se: gate[b,c] = sigmoid(W2 . silu(W1 . pooled[b] + b1) + b2);  w1 [R][C] and w2t = W2^T [R][C], fp32.
 * scale: x[b,p,c] *= gate[b,c] in place (SE excite, ahead of the conv3_1x1 GEMM).
 * layernorm_rows: LayerNorm(eps 1e-6) over C for every row (PartitionAttentionCl.norm1 / norm2); C % 64 == 0, <= 512.
 * attn: AttentionCl over 7x7 windows (grid_mode 0, 'block') or the 7x7 dilated grid (grid_mode 1): qkv rows
 *   [B*H*W, 3C] in image order, head h = columns [96h, 96h+96) = q|k|v (head_first); out rows [B*H*W, C];
 *   softmax(q k^T / sqrt(32) + table[rel_pos_index(i,j), h]) v;  table [169, C/32] fp32.  The window / grid partition
 *   and its reverse are index arithmetic inside the kernel.
 * lnpool: final LayerNorm2d(C) then global average pool -> out [B, C] fp32.
 */
int btsb_maxvit_stem1_fwd(const float* x, int64_t B, int Hin, int Win, int S, const float* w, const float* shift,
                          int C1, void* out, int dtype, void* stream);
int btsb_maxvit_im2col3_fwd(const void* x, void* out, int64_t B, int H, int W, int C, int dtype, void* stream);
int btsb_maxvit_avgpool2_fwd(const void* x, void* out, int64_t B, int H, int W, int C, int dtype, void* stream);
/* stem.conv2 as an implicit GEMM on tcgen05 (bf16): 3x3 / stride 1 / pad 1, 32 input channels, N = 64 output channels.
 * The A operand of every tap is one 4-D bulk tensor copy of the shifted 8 x 16 x 32 input box (TMA zero-fills the
 * padding); no patch matrix.  x [B,H,W,32] bf16 NHWC, w [N, 288] bf16 (k = (ky*3+kx)*32 + c), bias [N] fp32 or NULL,
 * out [B*H*W, N] bf16.  H % 8 == 0, W % 16 == 0; BTSB_EINVAL otherwise (use im2col3 + gemm). */
int btsb_conv3x3_c32_fwd(const void* x, const void* w, const float* bias, void* out, int64_t B, int H, int W, int N,
                         void* stream);
int btsb_maxvit_dw3_fwd(const void* x, int64_t B, int H, int W, int C, int stride, const float* w, const float* shift,
                        void* out, float* pooled, int dtype, void* stream);
int btsb_maxvit_se_fwd(const float* pooled, int64_t B, int C, int R, const float* w1, const float* b1, const float* w2t,
                       const float* b2, float* gate, void* stream);
int btsb_maxvit_scale_fwd(void* x, const float* gate, int64_t B, int HW, int C, int dtype, void* stream);
int btsb_layernorm_rows_fwd(const void* x, const float* ln_w, const float* ln_b, void* out, int64_t M, int C, int dtype,
                            void* stream);
int btsb_maxvit_attn_fwd(const void* qkv, void* out, int64_t B, int H, int W, int C, int grid_mode, const float* table,
                         int dtype, void* stream);
int btsb_maxvit_lnpool_fwd(const void* x, const float* ln_w, const float* ln_b, float* out, int64_t B, int HW, int C,
                           int dtype, void* stream);

/* debugging aid: device buffer (>= 17*64*8 int64) that block 0 of the fused-MLP kernel fills with clock64() stamps of its
 * pipeline hand-offs (scripts/mlp_trace.py); NULL switches it off (default). */
int btsb_debug_mlp_trace(void* buf);

/* dtype helpers used by the weight packer: float32 -> bf16 (round-to-nearest-even) and back. */
int btsb_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream);
int btsb_cast_bf16_to_f32(const void* in, float* out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BTSBOT_B200_H */
