/* nkbk.h -- C ABI of libnkbk.so: the B200 (sm_100a) kernels behind the
 * nkb-classification hot path.
 *
 * The reference (nkb-tech/nkb-classification) is 100 % Python and has no FFI
 * of its own; its boundary for this path is the config-driven Python API
 * (SURVEY.md section 8b).  Every entry point below therefore names the Python
 * call sites it replaces; the Python mirror in nkb_classification_b200/ binds
 * these symbols through ctypes and keeps the reference's signatures.
 *
 * Conventions
 *   - extern "C", plain pointers / sizes / scalars, no torch types.
 *   - Device pointers unless a parameter says "host".  The caller owns every
 *     buffer; the library owns only its NCCL communicator.
 *   - Every launch is asynchronous on `stream` (a cudaStream_t passed as
 *     void*; NULL = the legacy default stream) and re-entrant per stream.
 *   - Return value: 0 = NKBK_OK, negative = error; the message is available
 *     from nkbk_last_error() (thread local).  Nothing throws across the ABI.
 *   - There is no CPU fallback: without a CUDA device every compute entry
 *     point returns NKBK_E_CUDA.
 */
#ifndef NKBK_H_
#define NKBK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NKBK_ABI_VERSION 1

enum {
    NKBK_OK = 0,
    NKBK_E_ARG = -1,          /* NULL / out-of-range argument            */
    NKBK_E_SHAPE = -2,        /* unsupported size                         */
    NKBK_E_CUDA = -3,         /* CUDA runtime error (message has detail)  */
    NKBK_E_NCCL = -4,         /* NCCL error / communicator not initialised */
    NKBK_E_UNSUPPORTED = -5   /* valid request this build cannot serve    */
};

enum { NKBK_F32 = 0, NKBK_BF16 = 1 };                    /* element types */
enum { NKBK_MODE_STRETCH = 0, NKBK_MODE_LETTERBOX = 1 }; /* K1 geometry   */
enum { NKBK_LOSS_CE = 0, NKBK_LOSS_FOCAL = 1 };          /* K2 loss kind  */

int nkbk_abi_version(void);
const char* nkbk_last_error(void);
/* Number of kernels this library has launched in the calling process. */
int64_t nkbk_launch_count(void);

/* Launch option of the calling thread (default 0).  With 1, every following K1 launch of this thread
 * (nkbk_preprocess_crops / nkbk_preprocess_crops_aug) is enqueued as a PROGRAMMATIC dependent of the kernel in front
 * of it in the stream (cudaLaunchAttributeProgrammaticStreamSerialization): when that kernel is this library's fused
 * heads step -- which releases its dependents as soon as its CTAs are resident -- K1 of batch i+1 starts on the SMs the
 * heads step of batch i leaves free (it occupies ceil(B / rows_per_cta) SMs) instead of after it.  This is the
 * pipelined form of the reference's loop (engine.py:39-62: the DataLoader prepares batch i+1 while the model steps on
 * batch i) in ONE stream.  Only for callers that own their buffers: K1 must not write memory the kernel in front of it
 * still reads (the caller gives K1 its own output buffer), and K1 reads nothing that kernel writes.  Any other kernel
 * in front of K1 releases its dependents only when it ends, i.e. ordinary stream order.  Returns the previous value. */
int nkbk_k1_overlap_previous(int enable);

/* Launch option of the calling thread (default 1).  With 0 the nkbk_heads_* calls of this thread do not use the one-launch
 * persistent kernel (k2_fused_step) and run as their separate kernels (forward, dW, exchange / finalize) instead -- same
 * results contract.  The persistent kernel takes whole SMs (512 threads, ~221 KB of shared memory per CTA) and, at
 * world > 1, holds them while it waits for the other ranks; the separate kernels have small CTAs that slot in between
 * those of a K1 running on another stream, which is the faster arrangement for a loop that overlaps K1 of batch i+1
 * with the heads step of batch i at a full batch per GPU (8 x B200, 4096 crops per GPU: 0.532 vs 0.554 ms per step).
 * Returns the previous value. */
int nkbk_heads_one_launch(int enable);

/* ------------------------------------------------------------------------
 * K1  fused crop + cv2-INTER_LINEAR-exact uint8 resize + Normalize + HWC->NCHW
 *
 * Replaces, per sample: AnnotatedYOLODataset.__getitem__ slice
 * (nkb_classification/dataset.py:398-409), Transforms.__call__ (:96-102) with
 * the A.Resize | A.LongestMaxSize + A.PadIfNeeded(BORDER_CONSTANT), A.Normalize,
 * ToTensorV2 pipeline of configs/singletask_config.py:203-219, default_collate
 * + pin + fp32 H2D (dataset.py:608-629, engine.py:40,101) and the per-box loop
 * of Evaluator.classify_crops (metrics/det_cls_val.py:228-244).
 *
 *   frames_base  uint8 source pixels, interleaved 3-channel rows
 *   frame_desc   int64 [n_frames][4] = {byte offset from frames_base, height,
 *                width, row pitch in bytes}; width >= 2
 *   boxes        int32 [n][4] = x0, y0, x1, y1 (pixel, x1/y1 exclusive),
 *                as produced by bbox_xywhn2xyxy (dataset.py:414-421)
 *   frame_idx    int32 [n]
 *   mode         NKBK_MODE_STRETCH   -> A.Resize(out_h, out_w)
 *                NKBK_MODE_LETTERBOX -> A.LongestMaxSize(max_size) +
 *                                       centred constant pad to out_h x out_w
 *   pad_value    host uint8[3] (letterbox border, in OUTPUT channel order)
 *   mean255      host float[3]  = f32(mean) * f32(max_pixel_value)
 *   denom        host float[3]  = 1 / (f32(std) * f32(max_pixel_value))
 *   channel_swap non-zero: source is BGR, emit RGB (cv2.cvtColor BGR2RGB)
 *   out          [n][3][out_h][out_w] of out_dtype (NKBK_F32 | NKBK_BF16)
 *   out_u8       optional [n][out_h][out_w][3] resized uint8 pixels (may be NULL)
 *   bad_count    optional int32 counter, incremented once per crop whose box
 *                is empty or leaves its frame (such crops are filled with the
 *                normalised pad value)
 * ---------------------------------------------------------------------- */
int nkbk_preprocess_crops(const void* frames_base, const int64_t* frame_desc, int n_frames, const int32_t* boxes,
                          const int32_t* frame_idx, int n, int mode, int out_h, int out_w, int max_size,
                          const uint8_t* pad_value, const float* mean255, const float* denom, int channel_swap,
                          void* out, int out_dtype, uint8_t* out_u8, int32_t* bad_count, void* stream);

/* K1 with the deterministic part of the TRAIN-time pipeline fused in: per-crop parameters drawn on
 * the host (configs/singletask_config.py:172-194 -- A.HorizontalFlip, A.VerticalFlip, A.RandomBrightnessContrast
 * with brightness_by_max, A.HueSaturationValue, A.CoarseDropout) are applied, in that order, to the resized / padded
 * uint8 image before Normalize; same arithmetic as albumentations 1.x (cv2.flip, the clip(f32(v)*alpha +
 * beta*255).astype(uint8) look-up table, cv2 RGB2HSV -> hue/sat/val LUTs -> cv2 HSV2RGB, rectangle fill).
 * All other arguments as nkbk_preprocess_crops.
 *   aug_flags   int32 [n]   bit 0 horizontal flip, bit 1 vertical flip, bit 2 brightness/contrast on,
 *                           bit 3 hue/saturation/value shift on, bits 8.. number of holes (<= max_holes)
 *   aug_alpha   fp32 [n]    1 + contrast draw        (read when bit 2 is set)
 *   aug_beta    fp32 [n]    f32(brightness draw * max_pixel_value)
 *   aug_holes   int32 [n][max_holes][4] = x1, y1, x2, y2 in output pixels (ends exclusive); max_holes <= 16
 *   hole_fill   host uint8[3], CoarseDropout fill_value in OUTPUT channel order (NULL = 0)
 *   aug_hsv_lut uint8 [n][3][256]: per sample the hue, sat and val tables albumentations builds
 *               (mod(h + shift, 180), clip(s + shift, 0, 255), clip(v + shift, 0, 255), truncated); read only for
 *               samples with bit 3 set; may be NULL when no sample sets it
 *   hsv_trunc_cols  cv2's 8-bit HSV2RGB converts x * 255 to uint8 by TRUNCATION in its vectorised body (the first
 *               (out_w / lanes) * lanes pixels of every image row; lanes = 32 with AVX2) and by round-to-nearest-
 *               even in the scalar tail of the row: output columns [0, hsv_trunc_cols) take the former, the rest the
 *               latter (pass (out_w / 32) * 32 to reproduce an AVX2 host, 0 for the scalar rounding everywhere) */
int nkbk_preprocess_crops_aug(const void* frames_base, const int64_t* frame_desc, int n_frames, const int32_t* boxes,
                              const int32_t* frame_idx, int n, int mode, int out_h, int out_w, int max_size,
                              const uint8_t* pad_value, const float* mean255, const float* denom, int channel_swap,
                              const int32_t* aug_flags, const float* aug_alpha, const float* aug_beta,
                              const int32_t* aug_holes, int max_holes, const uint8_t* hole_fill,
                              const uint8_t* aug_hsv_lut, int hsv_trunc_cols, void* out, int out_dtype,
                              uint8_t* out_u8, int32_t* bad_count, void* stream);

/* The hue / sat / val tables of A.HueSaturationValue built ON THE DEVICE from the per-sample shifts, so that a batch
 * uploads 24 bytes per sample instead of 768 (albumentations' `_shift_hsv_uint8`: an int16 ramp plus the float64 shift,
 * `mod 180` for hue / `clip(0, 255)` for saturation and value, truncated to uint8 -- the same float64 operations here;
 * a zero shift, or a sample without bit 3 in aug_flags, keeps the identity table).
 *   hsv_shift   fp64 [n][3] hue, sat, val shift draws (device)
 *   aug_flags   int32 [n] (device; as for nkbk_preprocess_crops_aug)
 *   out_lut     uint8 [n][3][256] (device): what nkbk_preprocess_crops_aug takes as aug_hsv_lut */
/* Debug helper (host-synchronous): the per-CTA timeline of the calling process's last TMA-kernel K1 launch made with the
 * environment variable NKBK_K1_TIMING set: out_host uint64 [n_ctas][3] = {SM id, %globaltimer ns at CTA entry, at the
 * exit of its last warp}; returns the number of CTAs written (profiles/tools/k1_timeline.py). */
int64_t nkbk_debug_k1_timeline(uint64_t* out_host, int64_t max_ctas);

int nkbk_build_hsv_luts(const double* hsv_shift, const int32_t* aug_flags, int n, uint8_t* out_lut, void* stream);

/* Host-only helper (no CUDA): the per-axis coefficient table K1 uses, for
 * parity tests on machines without a GPU.  Writes dsize entries each. */
int nkbk_debug_axis_table(int dsize, int ssize, int horizontal, int32_t* src_index, int32_t* coef0, int32_t* coef1);
/* Host-only helper: letterbox geometry K1 uses. out4 = new_h, new_w, top, left. */
int nkbk_debug_letterbox(int h, int w, int max_size, int out_h, int out_w, int32_t* out4);
/* Host-only helper: the brightness/contrast look-up table K1 evaluates per pixel (256 entries). */
int nkbk_debug_brightness_contrast_lut(float alpha, float beta, uint8_t* lut256);
/* Host-only helper: K1's HueSaturationValue pixel function (RGB2HSV, the [3][256] LUTs, HSV2RGB) on n RGB pixels. */
int nkbk_debug_hsv_shift(const uint8_t* rgb_in, int64_t n, const uint8_t* lut768, int trunc, uint8_t* rgb_out);

/* ------------------------------------------------------------------------
 * K2  all task heads as one segmented GEMM + softmax + CE/focal loss + grads
 *
 * Replaces: the per-task nn.Linear of SingletaskClassifier / MultitaskClassifier
 * (nkb_classification/model.py:41-43, :114-116), FocalLoss.forward
 * (losses.py:59-94), nn.CrossEntropyLoss(weight) (losses.py:155-159),
 * MultitaskCriterion.__call__ (losses.py:110-147), their autograd backward
 * (engine.py:55-58) and the softmax of BaseLogger.log_iter (logging.py:270,279).
 *
 * Heads are concatenated: W_cat [NC][D], b_cat [NC], NC = seg_offsets[T];
 * task t owns rows seg_offsets[t] .. seg_offsets[t+1].
 *
 * The "reduce buffer" is one contiguous fp32 array (size from
 * nkbk_heads_reduce_buf_len):   [ dW_sum NC*D | db_sum NC | loss_sum T | denom T ]
 * holding UNNORMALISED sums over the local rows, so that N ranks can all-reduce
 * it once (K4) and then call nkbk_heads_finalize; 1 GPU calls finalize directly.
 *
 *   emb            [B][D] row-major, NKBK_F32 or NKBK_BF16
 *   W_cat, b_cat   fp32
 *   seg_offsets    host int32 [T+1]
 *   labels         int64 [B][T]; ignore_index rows contribute nothing
 *   class_weight   optional fp32 [NC] (CE `weight` / focal `alpha`), may be NULL
 *   out_logits     optional fp32 [B][NC]
 *   out_probs      optional fp32 [B][NC]   per-task softmax (fp32)
 *   dlogits        fp32 [B][NC]  d(loss_sum)/d(logit) (required when grads wanted;
 *                  may be NULL for a forward-only call -> reduce_buf then only
 *                  receives loss_sum / denom and its dW/db parts are zeroed)
 *   reduce_buf     fp32, see above
 *   workspace      scratch, size >= nkbk_heads_workspace_bytes(B, D, NC, T)
 * ---------------------------------------------------------------------- */
int64_t nkbk_heads_reduce_buf_len(int D, int NC, int T);
int64_t nkbk_heads_workspace_bytes(int B, int D, int NC, int T);

int nkbk_heads_fwd_loss_bwd(const void* emb, int emb_dtype, int B, int D, const float* W_cat, const float* b_cat,
                            const int32_t* seg_offsets, int T, const int64_t* labels, int loss_kind, float gamma,
                            const float* class_weight, int64_t ignore_index, float* out_logits, float* out_probs,
                            float* dlogits, float* reduce_buf, void* workspace, size_t workspace_bytes,
                            void* stream);

/* nkbk_heads_fwd_loss_bwd with K3 fused into the forward epilogue (the logits are still on chip there):
 *   out_pred  optional int32 [B][T]  per-task argmax (ties -> lowest index, NaN maximal: torch.argmax)
 *   cm_step   optional int64 confusion counts, layout as nkbk_argmax_confusion; ACCUMULATES; needs labels
 * One launch fewer per step than nkbk_heads_fwd_loss_bwd + nkbk_argmax_confusion; same results. */
int nkbk_heads_step(const void* emb, int emb_dtype, int B, int D, const float* W_cat, const float* b_cat,
                    const int32_t* seg_offsets, int T, const int64_t* labels, int loss_kind, float gamma,
                    const float* class_weight, int64_t ignore_index, float* out_logits, float* out_probs,
                    float* dlogits, float* reduce_buf, int32_t* out_pred, int64_t* cm_step, void* workspace,
                    size_t workspace_bytes, void* stream);

/* The whole training step of the heads in ONE launch: nkbk_heads_step, then -- by `exchange` --
 * nkbk_heads_finalize (NKBK_EXCHANGE_LOCAL: one GPU, or the caller all-reduces itself) or
 * nkbk_peer_allreduce_finalize (NKBK_EXCHANGE_PEER: the K4' exchange over NVLink peer memory), with identical results.
 * When the shape qualifies (gradients wanted, rows 16-byte aligned, NC * D small enough for register accumulators,
 * weights + a two-stage row ring within shared memory) this is one persistent cooperative kernel that reads every
 * embedding row from HBM once and whose epilogue IS the exchange; otherwise the same step runs as the separate launches.
 * Replaces, per training step: model.py:41-43 / :114-116, losses.py:59-94 / :110-147 / :155-159, the backward of
 * engine.py:55-58 through the heads, logging.py:270-281 (softmax, argmax) and metrics.py:31 (confusion counts).
 *   cm_step / cm_total / n_cm   as nkbk_heads_finalize (cm_step is scratch that is left zeroed); n_cm = 0 to skip
 *   out_loss   fp32 [T+1] per-task mean losses and their sum (over all ranks with NKBK_EXCHANGE_PEER)
 * The first nkbk_heads_workspace_bytes() of `workspace` need no initialisation. */
enum { NKBK_EXCHANGE_LOCAL = 0, NKBK_EXCHANGE_PEER = 1 };
int nkbk_heads_train_step(const void* emb, int emb_dtype, int B, int D, const float* W_cat, const float* b_cat,
                          const int32_t* seg_offsets, int T, const int64_t* labels, int loss_kind, float gamma,
                          const float* class_weight, int64_t ignore_index, float* out_logits, float* out_probs,
                          float* dlogits, float* reduce_buf, int32_t* out_pred, int64_t* cm_step, int64_t* cm_total,
                          int64_t n_cm, float* out_loss, int exchange, void* workspace, size_t workspace_bytes,
                          void* stream);

/* Optional hint for the NEXT nkbk_heads_* call of this thread: an identifier of the contents of W_cat that changes
 * whenever the weights change (e.g. torch's tensor version counter) and is never 0.  The tcgen05 forward keeps a bf16
 * copy of the weights in the caller's workspace; with the hint it re-packs that copy only when the version differs from
 * the one the workspace holds -- a validation or inference loop packs once.  Without the hint every call re-packs. */
void nkbk_heads_weights_version(int64_t version);

/* Which kernels served the calling thread's last nkbk_heads_step / nkbk_heads_train_step / nkbk_heads_fwd_loss_bwd:
 * NKBK_PATH_FUSED (k2_fused_step, one launch), NKBK_PATH_TC_FWD (tcgen05 forward + dW kernel) or NKBK_PATH_FFMA_FWD
 * (exact-fp32 FFMA forward + dW kernel).  Lets a caller see a fallback instead of guessing it from timings. */
enum { NKBK_PATH_FFMA_FWD = 1, NKBK_PATH_TC_FWD = 2, NKBK_PATH_FUSED = 4 };
int nkbk_heads_last_path(void);

/* Debug helper (host-synchronous): per-CTA phase time stamps of the last fused nkbk_heads_* launch made with the
 * environment variable NKBK_FUSED_TIMING=1 (profiles/tools/k2_phases.py).  out_host: uint64 [n_ctas][16]; returns the
 * number of CTAs written, 0 when there is nothing to report. */
int nkbk_debug_fused_timing(uint64_t* out_host, int max_ctas);

/* In place: dW, db <- sums / denom[task]; out_loss (fp32 [T+1]) <- per-task mean
 * losses and their unweighted sum (losses.py:140-147).  denom == 0 -> 0.
 * Optionally folds the (all-reduced) per-step confusion counts into the epoch
 * totals in the same launch: cm_total[i] += cm_step[i]; cm_step[i] = 0
 * (both NULL / n_cm = 0 to skip). */
int nkbk_heads_finalize(float* reduce_buf, int D, const int32_t* seg_offsets, int T, float* out_loss,
                        int64_t* cm_total, int64_t* cm_step, int64_t n_cm, void* stream);

/* d(sum_t task_scale[t] * loss_t)/d(emb) [B][D] (out_dtype) from dlogits and the (all-reduced, finalised or not
 * -- finalize leaves denom untouched) reduce_buf.  task_scale: optional device fp32 [T] (NULL = all ones, i.e. the
 * gradient of the summed loss that MultitaskCriterion returns). */
int nkbk_heads_demb(const float* dlogits, const float* reduce_buf, const float* W_cat, const int32_t* seg_offsets,
                    int T, int B, int D, const float* task_scale, void* out_demb, int out_dtype, void* stream);

/* Loss on already-computed logits: the reference's `criterion(pred, true)` call
 * (FocalLoss.forward losses.py:59-94, nn.CrossEntropyLoss losses.py:155-159,
 * MultitaskCriterion losses.py:110-147) when the heads were evaluated elsewhere.
 *   logits    [B][ld] NKBK_F32 | NKBK_BF16, task t in columns seg_offsets[t]..
 *   out_probs optional fp32 [B][NC]; dlogits optional fp32 [B][NC] = d(total loss)/d(logit)
 *   out_loss  fp32 [T+1] per-task mean losses and their sum
 *   workspace >= nkbk_loss_workspace_bytes(B, T) */
int64_t nkbk_loss_workspace_bytes(int B, int T);
int nkbk_loss_fwd_bwd(const void* logits, int dtype, int B, int ld, const int32_t* seg_offsets, int T,
                      const int64_t* labels, int loss_kind, float gamma, const float* class_weight,
                      int64_t ignore_index, float* out_probs, float* dlogits, float* out_loss, void* workspace,
                      size_t workspace_bytes, void* stream);

/* The UNREDUCED loss: FocalLoss(reduction="sum" | "none") of the reference (losses.py:87-94).
 *   out_row_loss  fp32 [B][T]: the loss term of every (row, task) -- -alpha (1 - pt)^gamma log pt, or the weighted
 *                 cross-entropy term -- 0 for rows whose label is ignore_index (or outside [0, C_t))
 *   dlogits       optional fp32 [B][NC] = d(row loss of the row)/d(logit), NOT divided by anything
 *   workspace     >= nkbk_loss_workspace_bytes(B, T)
 * "sum" is the sum of the row terms, "none" their vector over the kept rows; both are a few bytes of host glue. */
int nkbk_loss_rows(const void* logits, int dtype, int B, int ld, const int32_t* seg_offsets, int T,
                   const int64_t* labels, int loss_kind, float gamma, const float* class_weight, int64_t ignore_index,
                   float* out_row_loss, float* dlogits, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * K3  per-task argmax + confusion-matrix accumulation
 *
 * Replaces: argmax + D2H lists of BaseLogger.log_iter (logging.py:272-281),
 * the confusion matrix inside sklearn balanced_accuracy_score
 * (metrics.py:31) and inference()'s per-task argmax (inference.py:53,59).
 *
 *   logits   [B][ld] NKBK_F32 | NKBK_BF16, ld >= seg_offsets[T]
 *   labels   optional int64 [B][T] (NULL -> predictions only)
 *   out_pred optional int32 [B][T]
 *   cm       optional int64, task t at offset sum_{s<t} C_s^2, row = gt,
 *            col = pred; ACCUMULATES (caller zeroes at epoch start)
 * Ties -> lowest index; NaN is maximal (torch.argmax semantics).
 * ---------------------------------------------------------------------- */
int nkbk_argmax_confusion(const void* logits, int dtype, int B, int ld, const int32_t* seg_offsets, int T,
                          const int64_t* labels, int32_t* out_pred, int64_t* cm, void* stream);

/* ------------------------------------------------------------------------
 * K5  exact one-vs-rest ROC-AUC statistics on the device
 *
 * Replaces, for the quantity that needs every sample of the epoch: the per-step
 * confidence lists of BaseLogger.log_iter (logging.py:270,279) + label_binarize +
 * roc_auc_score of compute_targetwise_metrics (metrics.py:33-42).
 *
 *   probs       fp32 [N][ld], the epoch's K2 probabilities (device resident)
 *   labels      int64 [N][T]; a row whose label is outside [0, C_t) is a
 *               negative for every class of task t (what label_binarize makes of it)
 *   out_counts  int64 [NC][3] = { num2, P, Q } per class column:
 *               P / Q = positives / negatives, num2 = 2 * #{pos > neg} + #{pos == neg}
 *               so that  AUC = num2 / (2 * P * Q)  == sklearn.metrics.roc_auc_score
 *               (trapezoid with ties == Mann-Whitney U); P == 0 or Q == 0 -> undefined
 *   workspace   >= nkbk_auc_workspace_bytes(N, NC) (two N-float lists per column)
 * N < 2^31.  Integer arithmetic: exact and order independent.
 * ---------------------------------------------------------------------- */
int64_t nkbk_auc_workspace_bytes(int64_t N, int NC);
int nkbk_roc_auc_counts(const float* probs, int64_t N, int ld, const int32_t* seg_offsets, int T,
                        const int64_t* labels, int64_t* out_counts, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ------------------------------------------------------------------------
 * K4  batch sharding: the only exchange step of the path
 *
 * The reference is single-GPU (no torch.distributed anywhere);
 * this is the new data-parallel step: one all-reduce of the K2 reduce buffer
 * and one of the int64 confusion counts, NCCL over NVLink/NVSwitch.
 * ---------------------------------------------------------------------- */
#define NKBK_UNIQUE_ID_BYTES 128
int nkbk_comm_unique_id(void* out_id_host);               /* rank 0, host buffer of 128 bytes */
int nkbk_comm_init(int rank, int world, const void* id_host, int device);
int nkbk_comm_world(void);                                 /* 0 when not initialised */
/* Sum-all-reduce both payloads in one NCCL group; either may be NULL / 0. */
int nkbk_allreduce_heads(float* reduce_buf, int64_t n_f32, int64_t* cm, int64_t n_i64, void* stream);
int nkbk_comm_shutdown(void);

/* ------------------------------------------------------------------------
 * K4' the same exchange step fused with nkbk_heads_finalize over NVLink peer
 *     memory: ONE kernel pushes the local reduce buffer and
 *     step confusion counts into every peer's inbox (cudaIpc-mapped, 16-byte
 *     stores over NVLink/NVSwitch), waits on per-(rank, CTA) flags, sums the
 *     `world` copies in rank order (bit-identical on every rank) and applies
 *     the finalize step.  Replaces { nkbk_allreduce_heads, nkbk_heads_finalize }.
 *
 *   nkbk_peer_init     allocates this rank's inbox (capacity in fp32 / int64
 *                      elements of the largest payload) and returns its
 *                      64-byte IPC handle in a host buffer; world == 1 is allowed
 *   nkbk_peer_connect  maps the inboxes of all ranks: handles_host is
 *                      world x 64 bytes, rank-major (all-gathered by the caller);
 *                      the caller must barrier once afterwards, before first use
 *   nkbk_peer_allreduce_finalize   same arguments and results as
 *                      nkbk_heads_finalize, but the sums are over all ranks;
 *                      every rank must call it the same number of times with the
 *                      same shapes, on one stream per process
 *   nkbk_peer_status   host-synchronous: 0 = healthy, r + 1 = a wait on rank r
 *                      timed out (results of that step are invalid)
 * ---------------------------------------------------------------------- */
#define NKBK_IPC_HANDLE_BYTES 64
int nkbk_peer_init(int rank, int world, int device, int64_t max_f32, int64_t max_i64, void* out_handle_host);
int nkbk_peer_connect(const void* handles_host);
int nkbk_peer_world(void);                                 /* 0 when not connected */
int nkbk_peer_allreduce_finalize(float* reduce_buf, int D, const int32_t* seg_offsets, int T, float* out_loss,
                                 int64_t* cm_total, int64_t* cm_step, int64_t n_cm, void* stream);
int nkbk_peer_status(int32_t* out_status_host);
/* Tear-down is two-phase because an exported allocation must not be freed while a peer still maps it:
 * every rank calls nkbk_peer_disconnect (unmaps the peers), the caller barriers, then nkbk_peer_shutdown frees. */
int nkbk_peer_disconnect(void);
int nkbk_peer_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* NKBK_H_ */
