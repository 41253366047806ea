/*
 * yolo_b200.h -- C-ABI of libyolo_b200.so: the sm_100a kernels behind the
 * YOLOv3 detection hot path of GabeTsai/YOLO-For-Turbines.
 *
 * The reference has no native layer (it is 100 % Python on top of ATen/cuDNN),
 * so every entry point below names the reference *Python* call site it
 * replaces (paths are relative to the reference checkout).  The Python mirror
 * of the reference's module surface (yolo_for_turbines_b200/{model,utils}.py)
 * binds these with ctypes; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every function returns 0 (YB_OK) or a negative YB_ERR_* code and leaves a
 *     human-readable message retrievable with yolo_last_error() (thread-local);
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - every launch is asynchronous on the given cudaStream_t, performs no host
 *     synchronisation and no allocation: workspaces are sized by the
 *     *_workspace_bytes() queries and owned by the caller (torch tensors);
 *   - no global mutable state, so one thread per GPU may call concurrently;
 *   - plain C types only (no torch types), so any FFI can bind it.
 */
#ifndef YOLO_B200_H_
#define YOLO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef YOLO_B200_NO_CUDA_TYPES
typedef struct CUstream_st* yb_stream_t; /* == cudaStream_t */
#endif

#define YB_OK 0
#define YB_ERR_INVALID (-1)     /* bad argument                               */
#define YB_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed          */
#define YB_ERR_UNSUPPORTED (-3) /* shape/feature outside what the kernels do  */
#define YB_ERR_WORKSPACE (-4)   /* caller-provided workspace too small        */

/* bits of the device status word the kernels OR into (see yolo_conv_fwd)      */
#define YB_STATUS_NAN_INPUT 1u /* model.py:175  assert no NaN in the input     */
#define YB_STATUS_NAN_LAYER 2u /* model.py:183  ValueError("Nan in layer")     */

/* activation codes: model.py:62-70 (CNNBlock)                                 */
#define YB_ACT_NONE 0
#define YB_ACT_LEAKY 1 /* nn.LeakyReLU(0.1) */
#define YB_ACT_MISH 2  /* nn.Mish()         */

/* box formats of utils.py:38 calc_iou: "center" = cx,cy,w,h; anything else =
 * top-left x,y + w,h (the reference's "corners" is NOT x1y1x2y2).             */
#define YB_BOX_CENTER 0
#define YB_BOX_CORNERS 1

const char* yolo_last_error(void);
int yolo_version(void);
/* compute capability + SM count of `device`; fails unless it is sm_100.        */
int yolo_device_info(int device, int* cc_major, int* cc_minor, int* sm_count);

/* ------------------------------------------------------------------------- *
 * K1/K2  fused conv + folded BN + activation (+ residual, + 2x nearest
 * upsample on store, + zero-copy concat through channel pitches)
 * replaces CNNBlock.forward (model.py:80-86), ResidualBlock.forward's add
 * (model.py:115-121), nn.Upsample + torch.cat (model.py:189-191, :222).
 * Activations are NHWC bf16 (channel pitch may exceed C so that a tensor can
 * live inside a wider concat buffer); weights are [Cout_pad][kh][kw][Cin] bf16.
 * ------------------------------------------------------------------------- */
typedef struct yolo_conv_desc {
  int32_t batch, h_in, w_in, c_in;    /* logical input (c_in multiple of 32)  */
  int32_t in_pitch;                   /* elements between consecutive pixels  */
  int32_t c_out;                      /* logical output channels              */
  int32_t c_out_pad;                  /* rows of w_packed / scale / bias      */
  int32_t out_pitch;                  /* elements between output pixels       */
  int32_t ksize, stride, pad;         /* 1|3, 1|2, 0|1  (model.py:199-205)    */
  int32_t act;                        /* YB_ACT_*                             */
  int32_t has_residual, res_pitch;    /* y = act(..) + residual               */
  int32_t upsample2x;                 /* store every pixel to a 2x2 block     */
  int32_t out_fp32;                   /* 1: y is float (head), else bf16      */
  int32_t check_nan;                  /* OR YB_STATUS_NAN_LAYER on NaN output */
  int32_t a_mode;                     /* 0 auto, 1 force tiled-2D, 2 im2col   */
  int32_t block_n_hint;               /* tuning: 0 auto | 32 | 64 | 128 | 256 */
  int32_t stages_hint;                /* tuning: 0 auto | smem pipeline depth */
  int32_t impl_hint;                  /* 0 auto (persistent v2) | 1 one-tile-per-CTA v1 | 2 v2 */
  int32_t cta_pair_hint;              /* 0 auto | 1 single CTA | 2 tcgen05 cta_group::2 pair  */
  /* optional rectangular geometry along W (0 = same as the square fields): used by the engine's
   * pixel-pair folding of stride-2 layers, where a 3x3/s2/p1 conv becomes 3x2, stride (2,1), pad (1|0)  */
  int32_t ksize_w, stride_w;          /* filter width / stride along W                         */
  int32_t pad_w_hi_plus1;             /* 0: right pad = pad; else right pad = value - 1        */
  /* fused stem: > 0 marks the network's first conv (model.py:21) run straight from the NCHW fp32 image;
   * the desc then describes the pair-folded GEMM (ksize 1, c_in 64 = 2 x 32 taps, c_out 64, w_in = W/2),
   * x is not needed at plan time and the plan is launched with yolo_conv_fwd_stem.                      */
  int32_t stem_c;                     /* image channels (3), 0 = ordinary layer                */
  /* 1: the plan will be launched with yolo_conv_fwd_stats (reserves 8*c_out_pad bytes of shared memory for the
   * per-CTA channel sums)                                                                                        */
  int32_t want_stats;
  /* Data gradient of a 3x3 / stride-2 / pad-1 layer as two stride-1 sub-convolutions over dz, one per output ROW
   * parity r (0 | 1), instead of a 4x larger zero-stuffed conv:  dx[2a+r][2b+t][ci] = sum over (da, db) of
   * Wr[(t,ci)][da][db][co] * dz[a+da][b+db][co]  with da in {0} (r = 0) or {0, 1} (r = 1), db in {0, 1}.
   * s2_parity = r + 1 selects the mode: the desc is then ksize = 1 | 2, ksize_w = 2, pad = 0, pad_h_hi_plus1 =
   * ksize, pad_w_hi_plus1 = 2, c_out = c_out_pad = 2 * s2_cin (column n = t * s2_cin + ci), and every GEMM row
   * (img, a, b) is stored to pixel (2a + r, 2b + t) of the (batch, 2*h_out, 2*w_out) tensor y with out_pitch
   * elements per PIXEL (the residual operand is read the same way with res_pitch).
   * Weights: yolo_pack_weights_dgrad_s2.                                                                        */
  int32_t pad_h_hi_plus1;             /* 0: bottom pad = pad; else bottom pad = value - 1           */
  int32_t s2_parity, s2_cin;
  /* tuning switches of the persistent kernel, 0 = library default (on), 1 = off:
   * pdl_hint         launch with programmatic stream serialization (the prologue of layer i+1 overlaps the tail
   *                  of layer i; griddepcontrol.wait orders every global access after the previous launch);
   * tail_split_hint  cut the tiles of a last round that is at most half full into two half-width tiles.        */
  int32_t pdl_hint, tail_split_hint;
  /* row_hint: 0 auto | 1 off | 2 on with the smem descriptor's base_offset field set to the tap (measured WRONG on
   * B200: tcgen05 swizzles on absolute shared-memory address bits, the field must stay 0; kept as an A/B switch).  Row-window mode: 3x3 layers whose weights fit shared memory (the early,
   * L2-bandwidth-bound layers) load every filter row once per output-row segment and take the column taps as shifted
   * views of that tile, and keep the weights resident: ~2.5x less L2 -> SM traffic than nine im2col loads per tile.   */
  int32_t row_hint;
  /* Head conv with the anchor decode in its epilogue (ScalePredictionBlock's last conv, model.py:135-138, followed by
   * cells_to_boxes, utils.py:86-148): decode_mode = 1 makes y a CANDIDATE tensor -- fp32 rows [cx,cy,w,h,obj,cls] --
   * instead of the fp32 head: GEMM row (image, i, j) and anchor a go to row
   *   image * dec_rows_per_image + dec_row_offset + a * S * S + i * S + j        (S = h_out = w_out),
   * exactly what yolo_decode writes from the stored head (same arithmetic, bit-identical), so the (5+nc)*4 B per cell
   * head tensor never reaches HBM.  Needs 3 * (5 + dec_nc) <= c_out_pad = one tile (nc <= 80), out_fp32 = 1,
   * act = none.  dec_anchor_bits: the three (w, h) anchors already multiplied by S, as IEEE-754 bit patterns.       */
  int32_t decode_mode, dec_nc, dec_rows_per_image, dec_row_offset;
  int32_t dec_anchor_bits[6];
  /* mc_hint: 0 | 1 off (default), 2 on.  Weight-tile multicast: clusters of two CTA pairs on neighbouring M tiles of
   * one N tile; every CTA TMA-loads a quarter of the 256-row weight tile and multicasts it to the matching CTA of the
   * other pair (25 % less L2 -> SM traffic).  Correct (bit-identical) but MEASURED SLOWER on B200: only 33 clusters of
   * 4 CTAs are co-resident (132 of 148 SMs) and every big 3x3 layer loses 3-8 % (profiles/r2_multicast_ab.txt), so it
   * stays an opt-in A/B switch.                                                                                         */
  int32_t mc_hint;
} yolo_conv_desc;

/* Size of the opaque, caller-owned plan blob (64-byte aligned storage).       */
size_t yolo_conv_plan_bytes(void);
/* Encodes the TMA tensor maps and picks the tile configuration.  All device
 * pointers are baked into the plan (static buffers => CUDA-graph friendly).   */
int yolo_conv_plan_init(void* plan_host, size_t plan_bytes, const yolo_conv_desc* desc,
                        const void* x, const void* w_packed, const float* scale,
                        const float* bias, const void* residual, void* y);
int yolo_conv_fwd(const void* plan_host, uint32_t* status, yb_stream_t stream);
/* Same launch, plus per-channel statistics of the STORED bf16 output accumulated into sums2c (device doubles,
 * [2*c] += sum, [2*c+1] += sum of squares; caller zeroes): the BatchNorm batch statistics of the training
 * forward (model.py:61 under model.train()) without a second pass over the tensor.  bf16, non-upsampled plans.
 * fin != NULL: the CTA that finishes last also does what yolo_bn_finalize does (all pointers device pointers;
 * counter: a zero uint32 that is zero again on exit), so no finalize launch sits between the conv and
 * yolo_bn_act_fwd.                                                                                              */
typedef struct yolo_bn_finalize_desc {
  long long P;                              /* rows (batch * h_out * w_out)                                   */
  const float *gamma, *beta;
  float eps, momentum;
  float *running_mean, *running_var;        /* may be NULL                                                     */
  float *mean, *rstd, *scale, *bias;        /* outputs, as yolo_bn_finalize                                    */
  unsigned int* counter;
} yolo_bn_finalize_desc;
int yolo_conv_fwd_stats(const void* plan_host, uint32_t* status, double* sums2c, const yolo_bn_finalize_desc* fin,
                        yb_stream_t stream);
/* Fused stem launch: x_nchw = (B,3,H,W) fp32; also ORs YB_STATUS_NAN_INPUT (model.py:175).               */
int yolo_conv_fwd_stem(const void* plan_host, const float* x_nchw, uint32_t* status, yb_stream_t stream);
/* tile configuration chosen by plan_init: info8 = block_n, block_k, stages, tiles_n, tiles_m,
 * impl (1 | 2 | 3 = persistent kernel in row-window mode), CTAs per cluster, launched CTAs */
int yolo_conv_plan_info(const void* plan_host, int32_t* info8);
/* DEV TOOL: co-resident clusters of `cluster_size` CTAs of the 256-wide pair kernel (SM stranding per cluster size). */
int yolo_conv_max_clusters(int cluster_size, int* max_clusters);
/* DEV TOOL (scripts/conv_trace.py): yolo_conv_fwd with 32 x uint64 %globaltimer stamps / counters per launched CTA written to
 * trace_dev: [0] entry, [1] prologue done, [2] griddepcontrol.wait returned, [3] first TMA load issued,
 * [4] first operand stage landed, [5] last MMA committed, [6] first accumulator ready, [7] epilogue drained,
 * [8] exit, [9] tiles of this CTA (whole + half), [10..14], [19..22] epilogue box `box` of the first tile, [16..18] ns the MMA warp
 * waited for operands / for a free accumulator and the producer for a free ring stage.  Never used by the product path.                            */
int yolo_conv_fwd_trace(const void* plan_host, uint32_t* status, unsigned long long* trace_dev, int box, yb_stream_t stream);

/* TEST-ONLY reference: the same math on CUDA cores (direct convolution, one
 * thread per output element).  Never called by the product path.              */
int yolo_conv_fwd_simt(const yolo_conv_desc* desc, const void* x, const void* w_packed,
                       const float* scale, const float* bias, const void* residual,
                       void* y, uint32_t* status, yb_stream_t stream);

/* OIHW fp32 (nn.Conv2d.weight) -> [c_out_pad][k*k][c_in_pad] bf16, zero padded.
 * Replaces nothing in the reference; it is the repack the loader
 * (model.py:293-305) feeds.                                                   */
int yolo_pack_weights(const float* w_oihw, int c_out, int c_in, int ksize, int c_out_pad,
                      int c_in_pad, void* w_packed, yb_stream_t stream);
/* Stem weights: OIHW (c_out,3,3,3) -> [c_out_pad][32] bf16 with K index
 * (kh*3+kw)*3+c, matching yolo_input_patchify.                                */
int yolo_pack_stem_weights(const float* w_oihw, int c_out, int c_in, int c_out_pad,
                           void* w_packed, yb_stream_t stream);
/* BatchNorm2d (eval) folding: scale = g/sqrt(var+eps), bias = b - mean*scale;
 * gamma==NULL => scale=1, bias=conv bias (head conv).  model.py:61,84-86.     */
int yolo_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var,
                 const float* conv_bias, float eps, int c, int c_pad, float* scale,
                 float* bias, yb_stream_t stream);
/* NCHW fp32 -> NHWC bf16 (c padded with zeros to c_pad); ORs
 * YB_STATUS_NAN_INPUT into *status when x holds a NaN (model.py:175).          */
int yolo_nchw_to_nhwc_bf16(const float* x, int batch, int c, int h, int w, int c_pad,
                           int out_pitch, void* y, uint32_t* status, yb_stream_t stream);
/* NHWC bf16/fp32 -> NCHW fp32 (module-level drop-in outputs).                  */
int yolo_nhwc_to_nchw_f32(const void* x, int in_is_fp32, int batch, int c, int h, int w,
                          int in_pitch, float* y, yb_stream_t stream);
/* Stem im2col: NCHW fp32 (B,c,H,W), 9*c <= 32 -> [B*H*W][32] bf16 rows holding
 * the 9*c taps (kh,kw,c) of the pad-1 3x3 window, zero padded, so that the
 * Cin=3 stem conv (model.py:21) runs as a K=32 GEMM on the tensor cores.  Also
 * NaN-checks x (model.py:175).                                                */
int yolo_input_patchify(const float* x, int batch, int c, int h, int w, void* y,
                        uint32_t* status, yb_stream_t stream);

/* Inference pre-processing -- replaces the albumentations/OpenCV CPU pipeline of config.py:101-113
 * (`set_only_image_transforms`, used by demo.py:37-39): LongestMaxSize(S) with cv2.INTER_LINEAR's uint8 fixed-point
 * arithmetic, centred zero padding to S x S, /255, HWC -> CHW; fp32 (batch, channels, S, S) out.  descs_dev is a
 * DEVICE array of yolo_letterbox_desc_bytes()-sized records {const uint8_t* data; int32 h, w, nh, nw, top, left}:
 * source image (dense HWC uint8, device), resized size and padding offsets (host-computed, see preprocess.py).  */
size_t yolo_letterbox_desc_bytes(void);
int yolo_letterbox_u8(const void* descs_dev, int batch, int size, int channels, float* out, yb_stream_t stream);

/* ------------------------------------------------------------------------- *
 * K3  anchor decode -- replaces utils.py:86-148 cells_to_boxes.
 * head: (B,3,S,S,5+nc) with arbitrary element strides st[5]; fp32.
 * anchors6: 3x(w,h) already multiplied by S (utils.py:303).  Rows
 * [cx,cy,w,h,obj,cls] are written to out[(b*out_boxes_per_image + out_offset +
 * a*S*S + i*S + j)*6].  is_pred=0 follows utils.py:114-116.  writeback=1
 * reproduces the reference's in-place mutation of head[...,0:4] (:106-110).
 * ------------------------------------------------------------------------- */
int yolo_decode(const float* head, const int64_t* strides5_host, int batch, int S, int nc,
                const float* anchors6_host, int is_pred, int writeback, float* out,
                int out_boxes_per_image, int out_offset, yb_stream_t stream);
/* The scales of a detector (<= 4) in one launch, concatenated per image in the given order as utils.py:300-309 does.
 * Every head must be the dense [B*S*S pixels][pitch] fp32 layout the model produces (is_pred, no write-back);
 * YB_ERR_UNSUPPORTED otherwise (call yolo_decode per scale).  strides5_host: 5 per scale; anchors6_host: 6 per scale. */
int yolo_decode_multi(const float* const* heads_host, const int64_t* strides5_host, int batch, const int32_t* S_host,
                      int nc, const float* anchors6_host, int num_scales, float* out, int out_boxes_per_image,
                      yb_stream_t stream);

/* ------------------------------------------------------------------------- *
 * K4+K5+K6  threshold compaction, stable segmented sort, class-aware greedy
 * NMS -- replaces utils.py:150-191 non_max_suppression for a whole batch.
 * boxes: [total][6] fp32 rows [x,y,w,h,score,cls]; image b owns rows
 * [img_offsets[b], img_offsets[b+1]).  On return keep_idx[keep_off[b] ..
 * keep_off[b+1]) are the row indices the reference would return for image b,
 * in the reference's order (descending score, ties by original position).
 * ------------------------------------------------------------------------- */
size_t yolo_nms_workspace_bytes(int total, int batch);
/* class_bits: 0 = labels are arbitrary floats (grouped by float equality, NaN != NaN, as utils.py:178);
 * 8 | 16 = the caller guarantees integer labels in [0, 2^class_bits) (true for cells_to_boxes output), which
 * shortens the grouping sort from 4 radix passes to 1 | 2; violations are counted in workspace int32 #2.    */
int yolo_nms(const float* boxes, const int32_t* img_offsets, int batch, int total,
             float iou_thr, double obj_thr, int box_format, int class_bits, int32_t* keep_idx,
             int32_t* keep_off, void* workspace, size_t workspace_bytes, yb_stream_t stream);

/* Element-wise IoU -- replaces utils.py:38-84 calc_iou (n1 or n2 may be 1 to
 * broadcast) and utils.py:22-36 iou_aligned (aligned=1: rows are [w,h]).      */
int yolo_iou(const float* boxes1, int n1, int stride1, const float* boxes2, int n2, int stride2,
             int box_format, int aligned, float* out, yb_stream_t stream);

/* ------------------------------------------------------------------------- *
 * K7  mAP matching -- replaces the per-detection loop of utils.py:234-260.
 * dets [D][7] / gts [G][7] rows [img,cx,cy,w,h,score,cls].  gts must be
 * grouped by image (stable) and det_gt_lo/hi[d] give the gt range of det d's
 * image.  det_rank[d] = position of det d in the reference's per-class
 * descending-score order (any strictly order-preserving integer).  Outputs:
 * tp[d] in {0,1} (fp32), best_iou[d], best_gt[d] (-1 if none).
 * ------------------------------------------------------------------------- */
int yolo_map_match(const float* dets, int D, const float* gts, int G, const int32_t* det_gt_lo,
                   const int32_t* det_gt_hi, const int32_t* det_rank, float iou_thr,
                   int box_format, float* tp, float* best_iou, int32_t* best_gt,
                   int32_t* gt_claim /* [G] scratch */, yb_stream_t stream);

/* Per-class average precision -- replaces the per-class tail of utils.py:262-272 (cumsum, precision / recall with
 * (1, 0) prepended, torch.trapz).  tp_sorted [D] = TP flags (fp32 0/1) in evaluation order (class ascending, then
 * stable descending score); class c owns [cls_start[c], cls_end[c]); n_gt[c] = ground truths of the class.
 * ap[c] = AP of the class, 0 when it has no ground truth or no detections.                                       */
int yolo_map_ap(const float* tp_sorted, const int32_t* cls_start, const int32_t* cls_end, const int32_t* n_gt,
                int num_classes, float* ap, yb_stream_t stream);

/* Accuracy reductions -- replaces the per-scale body of utils.py:356-371 (check_model_accuracy).
 * head (B,3,S,S,5+nc) / target (B,3,S,S,6) fp32 with element strides; counts6 (device u64, accumulated):
 * correct_class, total_class, correct_obj, total_obj, correct_noobj, total_noobj.                          */
int yolo_accuracy_counts(const float* head, const int64_t* hstrides5_host, const float* target,
                         const int64_t* tstrides5_host, int batch, int S, int nc, float obj_thr,
                         unsigned long long* counts6, yb_stream_t stream);

/* K8 (forward only) -- the four terms of YOLOLoss.forward (loss.py:29-81) for one scale.  pred (B,3,S,S,5+nc)
 * and target (B,3,S,S,6) fp32 with element strides; anchors6 = the scale's 3 (w,h) anchors in grid units.
 * sums6 (device double[6], MUST be zero on entry): sum softplus over no-obj cells, #no-obj, sum (logit-iou)^2
 * over obj cells, sum of the 4 squared box terms, sum cross-entropy, #obj.  mutate=1 also applies loss.py:71-72's
 * in-place updates of pred[...,1:3] and target[...,2:4] (only when an object cell exists).                    */
int yolo_loss_fwd(float* pred, const int64_t* pstrides5_host, float* target, const int64_t* tstrides5_host,
                  int batch, int S, int nc, const float* anchors6_host, int mutate, double* sums6,
                  yb_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Training step (SURVEY config #4): replaces what autograd + cuDNN run for
 * `model.train(); out = model(x); loss.backward(); optimizer.step()`
 * (code/train.py:42-69).  Activations stay NHWC bf16, statistics fp32/fp64,
 * parameters, gradients and optimizer state fp32.
 *
 * Forward of one CNNBlock in train mode (model.py:80-86): yolo_conv_fwd with
 * scale=1/bias=0/act none writes the raw conv output z; yolo_bn_stats +
 * yolo_bn_finalize give the nn.BatchNorm2d batch statistics (and update the
 * running ones); yolo_bn_act_fwd applies BN + activation (+ residual, + the
 * 2x nearest upsample store of model.py:222).
 * Backward: yolo_bn_act_bwd turns the gradient of the block output into the
 * gradient dz of the raw conv output (+ dgamma, dbeta); the data gradient is
 * yolo_conv_fwd again on dz with transposed / flipped weights (stride-2
 * layers: on the zero-stuffed copy `stuffed`); yolo_wgrad is the weight
 * gradient GEMM on tcgen05 (MN-major operands straight from the NHWC tensors).
 * ------------------------------------------------------------------------- */
/* sums2c[2c] += sum_p z[p][c], sums2c[2c+1] += sum_p z[p][c]^2 (device doubles, caller zeroes them)          */
int yolo_bn_stats(const void* z, long long P, int C, int pitch, double* sums2c, yb_stream_t stream);
/* batch mean / rstd (biased variance), scale = gamma*rstd, bias = beta - mean*scale; running stats updated with
 * `momentum` and the unbiased variance (nn.BatchNorm2d, model.py:61); running_* may be NULL.                  */
int yolo_bn_finalize(const double* sums2c, long long P, int C, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, float* mean, float* rstd,
                     float* scale, float* bias, yb_stream_t stream);
/* yolo_bn_stats + yolo_bn_finalize in ONE launch: the block that finishes last (ticket in *counter, a device uint32
 * that must be 0 on entry and is 0 again on exit) finalises the layer.                                           */
int yolo_bn_stats_finalize(const void* z, long long P, int C, int pitch, double* sums2c, unsigned int* counter,
                           const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                           float* running_var, float* mean, float* rstd, float* scale, float* bias, yb_stream_t stream);
/* y = act(z*scale + bias) (+ residual); up2x: z is (B,h,w,C) and every pixel is stored to its 2x2 block of the
 * (B,2h,2w,*) tensor y (nn.Upsample + torch.cat, model.py:189-191, :222)                                       */
int yolo_bn_act_fwd(const void* z, long long P, int C, int z_pitch, const float* scale, const float* bias, int act,
                    const void* residual, int res_pitch, void* y, int y_pitch, int up2x, int h, int w,
                    yb_stream_t stream);
/* dA: gradient of the block output (up2x: taken as the 2x2 block sums of a (B,2h,2w,*) tensor = backward of
 * nn.Upsample).  Writes dz (bf16), dgamma/dbeta (fp32, assigned) and, when `stuffed` != NULL, dz zero-stuffed
 * to (B,2h,2w,*) for the stride-2 convs' data gradient.  sums2c: 2C zeroed doubles, counter: zero uint32 (left zero),
 * c1c0: 2C floats scratch.  Two launches: reduce (+ finalize by the last block) and apply.                       */
int yolo_bn_act_bwd(const void* dA, int dA_pitch, int up2x, const void* z, int z_pitch, long long P, int C, int h,
                    int w, const float* scale, const float* bias, const float* mean, const float* rstd, int act,
                    double* sums2c, unsigned int* counter, float* dgamma, float* dbeta, float* c1c0, void* dz,
                    int dz_pitch, void* stuffed, int stuffed_pitch, yb_stream_t stream);
/* bias gradient of the head conv (model.py:137, bias=True): dbias[c] = sum_p dz[p][c], c < C                   */
int yolo_bias_grad(const void* dz, long long P, int C_pad, int pitch, int C, double* sums2c, float* dbias,
                   yb_stream_t stream);
/* Weight gradient GEMM.  `desc` is the FORWARD geometry of the layer (yolo_conv_desc; in_pitch = pitch of x);
 * x = the layer input (NHWC bf16), dz = gradient of its raw output [P][dz_pitch] bf16; dw_packed =
 * [c_out_pad][k*k][c_in] fp32, ACCUMULATED (split-K partial tiles are added with red.global).                */
size_t yolo_wgrad_plan_bytes(void);
int yolo_wgrad_plan_init(void* plan_host, size_t plan_bytes, const yolo_conv_desc* desc, const void* x,
                         const void* dz, int dz_pitch, float* dw_packed, int splits_hint);
int yolo_wgrad(const void* plan_host, yb_stream_t stream);
/* info6 = n-tile, stages, splits, tiles_m, tiles_n, CTAs */
int yolo_wgrad_plan_info(const void* plan_host, int32_t* info6);
/* packed fp32 dW -> nn.Conv2d.weight.grad layout (OIHW, assigned); stem=1: packed rows are the 32-wide patch
 * order of yolo_input_patchify                                                                                */
int yolo_unpack_wgrad(const float* packed, int c_out, int c_in, int ksize, int c_in_pad, int stem,
                      float* grad_oihw, yb_stream_t stream);
/* OIHW fp32 -> the data-gradient weight pack [rows_pad >= c_in][k*k][cols_pad >= c_out] bf16 with
 * out[ci][tap][co] = w[co][ci][k*k-1-tap]: yolo_conv_fwd on dz with this pack computes d conv / d input.   */
int yolo_pack_weights_dgrad(const float* w_oihw, int c_out, int c_in, int ksize, int rows_pad, int cols_pad,
                            void* w_packed, yb_stream_t stream);
/* Both operand packs of one layer in one pass (training: the weights change every step).  The padding entries of
 * w_fwd [c_out_pad][k*k][c_in_pad] and w_dgrad [c_in_pad][k*k][c_out_pad] are NOT written: zero them once.       */
int yolo_pack_weights_train(const float* w_oihw, int c_out, int c_in, int ksize, int c_in_pad, int c_out_pad,
                            void* w_fwd, void* w_dgrad, yb_stream_t stream);
/* Weight pack of the stride-2 data-gradient sub-convolution of row parity r (see yolo_conv_desc.s2_parity):
 * out[(t*rows_half + ci)][da][db][co] bf16, zero padded to [2*rows_half][(r+1)*2][cols_pad]; rows_half >= c_in. */
int yolo_pack_weights_dgrad_s2(const float* w_oihw, int c_out, int c_in, int r, int rows_half, int cols_pad,
                               void* w_packed, yb_stream_t stream);
/* torch.optim.SGD(momentum, weight_decay) (train.py:171-172) on flat fp32 buffers: g' = g*grad_scale + wd*p;
 * buf = first_step ? g' : momentum*buf + g'; p -= lr*buf                                                      */
int yolo_sgd_step(float* param, const float* grad, float* momentum_buf, long long n, float lr, float momentum,
                  float weight_decay, float grad_scale, int first_step, yb_stream_t stream);
/* The same with the learning rate read from device memory at run time: a step captured in a CUDA graph follows a
 * per-iteration schedule (train.py:71-74 LinearLR warm-up, CosineAnnealingLR) without being captured again.        */
int yolo_sgd_step_dev(float* param, const float* grad, float* momentum_buf, long long n, const float* lr_dev,
                      float momentum, float weight_decay, float grad_scale, int first_step, yb_stream_t stream);
/* Training-target encoder -- replaces YOLODataset.__getitem__'s anchor assignment (dataset.py:119-167, iou_aligned
 * utils.py:22-36) + collate_fn's per-scale stacking (utils.py:694-700) for a batch.  boxes: device [total][5] doubles
 * x, y, w, h, class in the image's own order; image b owns rows [offsets[b], offsets[b+1]).  anchors18_host: the 9
 * (w, h) anchors scale-major as fractions of the image.  t0/t1/t2: (batch, 3, S, S, 6) fp32, fully written.        */
int yolo_encode_targets(const double* boxes, const int32_t* offsets, int batch, const float* anchors18_host, int S0,
                        int S1, int S2, float ignore_iou_threshold, float* t0, float* t1, float* t2,
                        yb_stream_t stream);
/* K8 backward: gradient of the summed, lambda-weighted YOLOLoss terms of one scale (loss.py:54-81) w.r.t. pred.
 * sums6 = the device sums yolo_loss_fwd produced for the same pred/target; dpred gets all 5+nc entries of every
 * cell (element strides dstrides5; out_bf16 selects bf16 or fp32), scaled by grad_scale and, per term, by
 * term_scales4_host = upstream gradients of [box, object, no-object, class] (NULL = all 1).                  */
int yolo_loss_bwd(const float* pred, const int64_t* pstrides5_host, const float* target,
                  const int64_t* tstrides5_host, int batch, int S, int nc, const float* anchors6_host,
                  const double* sums6, float grad_scale, const float* term_scales4_host, void* dpred,
                  const int64_t* dstrides5_host, int out_bf16, yb_stream_t stream);

/* Stable LSD radix sort of (u64 key, i32 value) pairs on bits [0,end_bit)
 * (end_bit multiple of 8); K5's building block, exported for tests and for
 * the mAP score ordering.  n_dev: device int32 holding the live count (<=max_n).
 * Result lands in keys/vals (copied back if the pass count is odd).           */
size_t yolo_sort_workspace_bytes(int max_n);
int yolo_sort_pairs(uint64_t* keys, int32_t* vals, const int32_t* n_dev, int max_n, int end_bit,
                    void* workspace, size_t workspace_bytes, yb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLO_B200_H_ */
