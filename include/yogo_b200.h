/*
 * yogo_b200 - C ABI of the B200-native YOGO hot path (libyogo_b200.so).
 *
 * The reference (czbiohub-sf/yogo) has no native boundary of its own: its hot path calls
 * torch / torchvision library kernels from Python.  This header is
 * the boundary a maintainer binds instead; each entry point cites the reference call
 * site it replaces.  Conventions:
 *   - extern "C", plain pointers and sizes, no torch types.  All data pointers are DEVICE
 *     pointers unless the name ends in _host.  The library never owns tensor memory.
 *   - every function returns 0 on success or a negative yg_status; yg_last_error()
 *     returns a thread-local message.  No C++ exception crosses the ABI.
 *   - every launch is asynchronous on the caller's stream (last argument, a cudaStream_t
 *     passed as void*).
 *   - activations are NHWC; `dtype` selects their storage: YG_F32 or YG_BF16 (accumulation
 *     is always fp32).  Parameters and parameter gradients are fp32 in the reference's
 *     OIHW layout (state_dict compatible, model.py:94-147).
 */
#ifndef YOGO_B200_H
#define YOGO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  YG_OK = 0,
  YG_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  YG_ERR_CUDA = -2,      /* CUDA runtime / driver error (message has the string) */
  YG_ERR_ARCH = -3,      /* device is not sm_100 */
  YG_ERR_WORKSPACE = -4  /* workspace too small */
} yg_status;

enum { YG_F32 = 0, YG_BF16 = 1, YG_U8 = 2 };
enum { YG_ACT_NONE = 0, YG_ACT_LRELU = 1, YG_ACT_SILU = 2 }; /* LeakyReLU(0.01) / SiLU */
enum { YG_IMPL_AUTO = 0, YG_IMPL_SIMT = 1, YG_IMPL_TCGEN05 = 2 };

int yg_version(void);
const char* yg_last_error(void);
/* 0 if the current device is compute capability 10.x, YG_ERR_ARCH otherwise. */
int yg_device_check(void);
/* Process-wide implementation switch for the 3x3 convolutions (AUTO = tcgen05 where the
 * shape qualifies, SIMT otherwise).  Used by the parity tests to cross-check both. */
int yg_set_conv_impl(int impl);
int yg_get_conv_impl(void);
/* debug/tuning switch of the tcgen05 engine (default 25 + 8192 + 16384): bit 0 = weights resident in shared memory when they fit,
 * bit 1 = cp.async producer for 16/32-channel operands, bit 2 = L2 prefetch warp, bit 3 = W-fold of small-channel
 * stride-1 layers, bit 4 = column-pair fold of small-channel stride-2 dgrad, bits 5-6 = cap the K chunk at 32 / 16,
 * bits 7-11 = profiling knobs of tools/bench_conv.py (skip stores / MMAs / epilogue / TMA, cycle counters),
 * bit 12 = swizzle-phase experiment, bit 13 = 2-D halo boxes, bit 14 = CTA pairs (cta_group::2) for 128->128 layers,
 * bit 15 = CTA pairs for stride-2 dgrad (slower), bit 16 = mma.sync wgrad also for 32->64 stride 2 (slower),
 * bit 17 = no mma.sync wgrad at all, bit 18 = no CTA pairs for N = 64 stride-1 dgrad,
 * bit 19 = no merged row taps (one N = 192 MMA) in the 64-input-channel wgrad, bit 20 = stride-2 wgrad fetches the dz tile once
 * per parity group instead of once per tile, bit 21 = fp32 tensors as split-bf16 "x3" convolutions on the tensor cores
 * (csrc/x3.cu; off by default: the exact SIMT kernels are the fp32 parity path), bit 22 = MMA issue by the per-tap table walk
 * instead of the flat host-built list (same MMAs in the same order; slower), bits 24-26 = A-operand pipeline depth of the
 * tcgen05 conv engine: 0 = at most 4 stages (default), 1..6 = at most 2..7, 7 = every stage that fits in shared memory. */
int yg_set_tc_options(int options);
int yg_get_tc_options(void);
/* profiling hook: copies n (<= 2048) cycle counters written by the engine's MMA warps (8 per CTA) to host memory. */
int yg_tc_debug_read(unsigned long long* out, int n);
/* number of CUDA kernels this library has launched in this process (bench.py gpu_launches). */
unsigned long long yg_launch_count(void);

/* ---- epilogue descriptor shared by the convolution entry points -------------------- */
typedef struct {
  const float* scale;      /* [Cout] or NULL (=1): folded BatchNorm scale in eval mode   */
  const float* shift;      /* [Cout] or NULL (=0): conv bias, or folded BN shift         */
  int act;                 /* YG_ACT_*                                                   */
  const float* dropscale;  /* [N*Cout] or NULL: Dropout2d keep/(1-p) per (n, c)          */
  double* stats;           /* [2*Cout] or NULL: += sum(v), sum(v*v) of v = acc*scale+shift
                              (BatchNorm batch statistics, taken before act)             */
  void* preact;            /* NHWC, same dtype as y, or NULL: v before the activation    */
  void* actmask;           /* [N*Ho*Wo*Cout/8] bytes or NULL: bit e (NHWC element index, bit e%8 of byte
                              e/8) = (v > 0).  One bit instead of 16 to carry LeakyReLU's slope to
                              the backward pass; honoured by yg_conv_fwd, needs Cout % 32 == 0 */
} yg_fwd_epilogue;

typedef struct {
  const void* saved;       /* NHWC tensor of the layer that produced this input:
                              post-activation output (lrelu, no BN), pre-activation
                              (silu, no BN) or raw conv output (BN); NULL = no act bwd  */
  int act;                 /* YG_ACT_* of that layer                                     */
  const float* dropscale;  /* [N*Cin] or NULL                                            */
  const float* bn_scale;   /* [Cin] gamma*invstd, NULL if that layer has no BN           */
  const float* bn_shift;   /* [Cin] beta - mean*gamma*invstd                             */
  const float* bn_mean;    /* [Cin]                                                      */
  const float* bn_invstd;  /* [Cin]                                                      */
  double* bn_sums;         /* [2*Cin]: += sum(g), sum(g*xhat)                            */
  const void* actmask;     /* the producer's yg_fwd_epilogue.actmask or NULL.  Used instead of `saved`
                              by the tcgen05 path when act is LeakyReLU and that layer has no BN
                              (16x less data to read); other paths read `saved`, so pass both */
} yg_bwd_epilogue;

/* ---- first layer: direct stencil on the NCHW image (Cin = 1 or 3) -------------------
 * replaces nn.Conv2d(input_channels, C1, 3, stride=2, padding=1) (+BatchNorm2d+act),
 * /root/reference/yogo/model_defns.py:33-37.  x is (N,Cin,H,W) in x_dtype (YG_U8 or
 * YG_F32; YOGO.forward's `x.float()` of model.py:272-273 is folded in), w is OIHW fp32.
 * If ep->stats is set and y is NULL only the statistics are accumulated (pass 1 of
 * train-mode BN); otherwise y (NHWC, dtype) is written. */
int yg_conv_first_fwd(const void* x, int x_dtype, const float* w, void* y, int dtype,
                      int N, int H, int W, int Cin, int Cout, int stride,
                      const yg_fwd_epilogue* ep, void* stream);
/* backward of the first layer from da = grad wrt its (post-activation) output.
 * pass 1 (bn_sums != NULL in `be`, dw == NULL): accumulate BN backward sums.
 * pass 2 (dw != NULL): dW (OIHW fp32, overwritten) and, if dshift, d(bias or beta);
 *   bn_dy_mean/bn_dyx_mean = sums/M from pass 1 (NULL when no BN). */
int yg_conv_first_bwd(const void* x, int x_dtype, const float* w, const void* da, int dtype,
                      int N, int H, int W, int Cin, int Cout, int stride,
                      const yg_bwd_epilogue* be, const float* fwd_shift,
                      const float* bn_dy_mean, const float* bn_dyx_mean,
                      float* dw, float* dshift, float clip, void* workspace, size_t workspace_bytes,
                      void* stream);
size_t yg_conv_first_bwd_workspace(int Cin, int Cout);
/* Single-channel images (Cin == 1): the 9-tap Gram statistics of the image, gram[54] = S[9] then the upper
 * triangle of G[9][9] (fp64, overwritten).  With them the BatchNorm batch statistics of the first layer
 * (yg_conv_first_stats_from_gram: stats[2*Cout] += sum y, sum y^2) and the BN part of its backward
 * (yg_conv_first_bwd_finalize, from P = sum g*x_t and Sg = sum g as produced by yg_conv_first_bwd with
 * bn_dy_mean == NULL) follow in closed form instead of extra passes over the 16-channel activations. */
int yg_conv_first_gram(const void* x, int x_dtype, int N, int H, int W, int stride, double* gram, void* stream);
int yg_conv_first_stats_from_gram(const double* gram, const float* w, const float* bias, double count, int Cout,
                                  double* stats, void* stream);
int yg_conv_first_bwd_finalize(const float* P, const float* Sg, const double* gram, const float* w, const float* bias,
                               const float* gamma, const float* mean, const float* invstd, double count,
                               int batch_stats, float clip, int Cout, float* dw, float* dbias, float* dgamma,
                               float* dbeta, void* stream);

/* Alignment: activation tensors handed to the convolution entry points should be 32-byte aligned (torch allocations are
 * 256-byte aligned); tensors that are only 16-byte aligned run on the generic SIMT path. */
/* ---- generic convolution (3x3 pad 1 or 1x1 pad 0, stride 1 or 2), NHWC -------------
 * replaces nn.Conv2d fprop / dgrad / wgrad (cuDNN) at model_defns.py:34-67 and the
 * elementwise BatchNorm/LeakyReLU/SiLU/Dropout2d kernels that follow each conv. */
int yg_conv_fwd(const void* x, const float* w_oihw, void* y, int dtype,
                int N, int H, int W, int Cin, int Cout, int ksize, int stride,
                const yg_fwd_epilogue* ep, void* stream);
/* dx = conv_transpose(dz, w); the epilogue turns it into the gradient wrt the previous
 * layer's conv output (activation / dropout / BN-sum fusion), see yg_bwd_epilogue. */
int yg_conv_dgrad(const void* dz, const float* w_oihw, void* dx, int dtype,
                  int N, int H, int W, int Cin, int Cout, int ksize, int stride,
                  const yg_bwd_epilogue* be, void* stream);
/* dw (OIHW fp32) = sum over pixels dz * x ; dbias[Cout] = sum dz (NULL to skip).
 * Both are clamped to [-clip, clip] (YOGO's per-parameter grad hook, model.py:76-77);
 * clip <= 0 disables.  Deterministic split-K through `workspace`. */
size_t yg_conv_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int ksize, int stride);
int yg_conv_wgrad(const void* x, const void* dz, float* dw, float* dbias, int dtype,
                  int N, int H, int W, int Cin, int Cout, int ksize, int stride,
                  float clip, void* workspace, size_t workspace_bytes, void* stream);

/* ---- BatchNorm2d pieces (nn.BatchNorm2d, model_defns.py:35,55,60) ------------------ */
/* from stats (sum, sumsq over M = N*H*W values per channel): mean, invstd, fused
 * scale/shift, and the running-stat update (momentum 0.1, unbiased variance). */
int yg_bn_finalize(const double* stats, double count, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float momentum, float eps,
                   float* mean, float* invstd, float* scale, float* shift, int C, void* stream);
/* eval mode: scale = gamma/sqrt(rv+eps), shift = beta - rm*scale (+ conv_bias*scale). */
int yg_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean,
                    const float* running_var, const float* conv_bias, float eps,
                    float* scale, float* shift, int C, void* stream);
/* a = act(y*scale[c] + shift[c]) * dropscale[n,c]   (elementwise, NHWC); actmask (or NULL): sign bits of the
 * activation input, same format as yg_fwd_epilogue.actmask (needs C % 8 == 0). */
int yg_bn_act_apply(const void* y, void* a, int dtype, int N, int HW, int C,
                    const float* scale, const float* shift, int act, const float* dropscale,
                    void* actmask, void* stream);
/* stats[0..C) += sum(y), stats[C..2C) += sum(y*y) over the N*HW pixels of an NHWC tensor (train-mode batch statistics
 * of a convolution output in one streaming pass; the alternative is yg_fwd_epilogue.stats).  C % 8 == 0. */
int yg_bn_stats(const void* y, int dtype, int N, int HW, int C, double* stats, void* stream);
/* sums[0..C) += sum(g), sums[C..2C) += sum(g * xhat), xhat = (y - mean) * invstd: the two reductions of BatchNorm
 * backward for a gradient g that already includes the activation backward (the alternative is yg_bwd_epilogue.bn_sums). */
int yg_bn_bwd_sums(const void* g, const void* y, int dtype, int N, int HW, int C, const float* mean,
                   const float* invstd, double* sums, void* stream);
/* dz = gamma*invstd*(g - sum_g/M - xhat*sum_gx/M) in place over g (batch_stats != 0), or
 * dz = gamma*invstd*g when the layer ran with running statistics (batch_stats == 0, BN in
 * eval mode inside a training graph: YOGO(tuning=True), model.py:69-70).  Also writes
 * dgamma = clamp(sum_gx), dbeta = clamp(sum_g). y = raw conv output (for xhat). */
int yg_bn_bwd_apply(void* g, const void* y, int dtype, int N, int HW, int C,
                    const double* sums, const float* gamma, const float* mean, const float* invstd,
                    float* dgamma, float* dbeta, float clip, int batch_stats, void* stream);

/* ---- head: 1x1 conv to 5+C channels fused with YOGO.forward's transform -------------
 * replaces model_defns.py:67 + model.py:277-313.  x NHWC (N,Sy,Sx,Cin) -> out (N,5+C,Sy,Sx)
 * fp32 NCHW; t_raw (N,Sy,Sx,5+C) fp32 keeps the raw logits for the backward (or NULL).
 * cxs/cys: the module's _Cxs/_Cys buffers (Sy*Sx fp32, model.py:48-59) or NULL to recompute. */
int yg_head_fwd(const void* x, int dtype, const float* w, const float* bias, float* out, float* t_raw,
                int N, int Sy, int Sx, int Cin, int num_classes,
                float anchor_w, float anchor_h, float width_mult, float height_mult,
                int inference, const float* cxs, const float* cys, void* stream);
/* dpred (N,5+C,Sy,Sx) fp32 + t_raw -> dx NHWC (with the previous block's bwd epilogue),
 * dw (OIHW [5+C,Cin,1,1]) and dbias, clamped. Training head only (inference == 0). */
size_t yg_head_bwd_workspace(int N, int Sy, int Sx, int Cin, int num_classes);
int yg_head_bwd(const float* dpred, const float* t_raw, const void* x, const float* w, void* dx, int dtype,
                float* dw, float* dbias, int N, int Sy, int Sx, int Cin, int num_classes,
                float anchor_w, float anchor_h, float width_mult, float height_mult,
                const yg_bwd_epilogue* be, float clip, void* workspace, size_t workspace_bytes,
                void* stream);

/* ---- YOGOLoss forward + backward in one pass -----------------------------------------
 * replaces yogo_loss.py:38-129 (+ torchvision box_convert / complete_box_iou_loss).
 * pred (N,5+C,Sy,Sx) fp32, label (N,6,Sy,Sx) fp32.  out4 (device, 4 floats) =
 * [loss, iou_loss, objectness_loss, classification_loss]; dpred may be NULL. */
size_t yg_yogo_loss_workspace(int N, int Sy, int Sx);
int yg_yogo_loss_fwd_bwd(const float* pred, const float* label, float* out4, float* dpred,
                         int N, int num_classes, int Sy, int Sx,
                         float no_obj_weight, float iou_weight, float classify_weight,
                         float label_smoothing, void* workspace, size_t workspace_bytes, void* stream);

/* ---- threshold + NMS + per-class counts, whole batch ----------------------------------
 * replaces format_preds (utils/prediction_formatting.py:23-93, torchvision.ops.nms) and
 * get_prediction_class_counts (infer.py:60-124).  preds (B,5+C,Sy,Sx) fp32.
 * Outputs: keep_count[B] (int32), rows (B, Sy*Sx, 5+C) fp32 - image b's kept rows packed
 * at rows[b][0..keep_count[b]) in the reference's output order - keep_index (B, Sy*Sx)
 * int32 grid-cell index of each kept row, class_counts[C] (int64, summed over the batch,
 * overwritten).  xyxy != 0 returns converted boxes (box_format="xyxy"). */
size_t yg_format_preds_workspace(int B, int num_classes, int Sy, int Sx);
int yg_format_preds_batch(const float* preds, int B, int num_classes, int Sy, int Sx,
                          float obj_thresh, double iou_thresh, int xyxy, float min_class_conf,
                          int* keep_count, float* rows, int* keep_index, long long* class_counts,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- evaluation matching ------------------------------------------------
 * cost[i][j] = 1 - IoU(label box i, prediction box j), the matrix format_preds_and_labels_v2 gives to the Hungarian
 * solver (utils/prediction_formatting.py:296-298, torchvision.ops.box_iou); bit-identical to the reference's.
 * labels / preds: rows of label_stride / pred_stride floats whose first four are xyxy (pass labels + 1 for
 * [mask, x1, y1, x2, y2, class] rows); cost (n_labels, n_preds) fp32, row-major. */
int yg_box_iou_cost(const float* labels, int label_stride, int n_labels, const float* preds, int pred_stride,
                    int n_preds, float* cost, void* stream);

/* ---- input side --------------------------------------------------------
 * RandomHorizontal/VerticalFlipWithBBs (data/data_transforms.py:51-98) of a whole batch, out of place:
 * images (N,C,H,W) uint8 (YG_U8) or fp32, labels (N,6,Sy,Sx) = [mask,x1,y1,x2,y2,class] with x1' = 1 - x2, x2' = 1 - x1
 * (hflip) / y1' = 1 - y2, y2' = 1 - y1 (vflip) on every cell and the cells mirrored; both flips may be combined. */
int yg_flip_images(const void* in, void* out, int dtype, int N, int C, int H, int W, int hflip, int vflip, void* stream);
int yg_flip_labels(const float* in, float* out, int N, int Sy, int Sx, int hflip, int vflip, void* stream);
/* format_labels_tensor (data/yogo_dataset.py:24-46) for a batch of ragged label lists: labels (total,5) =
 * [class,x1,y1,x2,y2] rows of all images back to back, offsets[B+1] (int32) the row range of each image, max_labels the
 * largest per-image count; out (B,6,Sy,Sx) fp32 overwritten; owner (B*Sy*Sx int32) is scratch; *err (device int) is set
 * to 1 when a label centre falls outside the grid (the reference raises IndexError). */
int yg_format_labels_batch(const float* labels, const int* offsets, int B, int max_labels, int Sy, int Sx,
                           int* owner, float* out, int* err, void* stream);

/* ---- optimizer: fused AdamW over one flat fp32 buffer ---------------
 * replaces torch.optim.AdamW.step (train.py:213-217, 324). grads are pre-scaled by
 * grad_scale (1/world_size after an all-reduce SUM). */
int yg_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                  float lr, float beta1, float beta2, float eps, float weight_decay,
                  long long step, float grad_scale, void* stream);
/* graph-replayable variant: hyper_dev (device, 8 floats) = [lr, beta1, beta2, eps, weight_decay,
 * 1-beta1^t, sqrt(1-beta2^t), grad_scale] */
int yg_adamw_flat_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                      const float* hyper_dev, void* stream);

/* ---- Dropout2d keep-scales of every block of a step + BatchNorm step counters, one launch -------------------------
 * replaces nn.Dropout2d's mask draw (model_defns.py:44-57) and `num_batches_tracked += 1`.  out: `total` floats, the
 * blocks' (N, C) scale matrices back to back; table (device): nblocks x int4 (offset, count, p as float bits, 0);
 * state (device, 3 x uint64): [seed, call counter, 0] - the kernel advances the counter, so CUDA-graph replays draw
 * fresh masks; counters (device): ncounters pointers to int64 `num_batches_tracked` buffers, each incremented by one. */
int yg_dropout_scales(float* out, const void* table, int nblocks, int total, void* state, const void* counters,
                      int ncounters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YOGO_B200_H */
