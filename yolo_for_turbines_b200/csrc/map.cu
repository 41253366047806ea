// K7: mAP matching step.  Replaces the O(D*G) Python loop of utils.py:234-260
// (one torch.tensor() + calc_iou call per (detection, ground-truth) pair).
//
// Parallel form used here (float-identical to the sequential loop):
//   best_gt[d]  = first gt of d's image and class with the strictly largest
//                 IoU > 0                                   (utils.py:240-249)
//   TP[d]       = best_iou[d] > thr  AND  d is the first detection, in the
//                 reference's descending-score order of d's class, whose
//                 best_gt is that gt and whose best_iou > thr (utils.py:252-257)
// "first" is resolved with an atomicMin over det_rank on a per-gt claim word.
#include "common.cuh"

namespace {

__global__ void k_map_best(const float* __restrict__ dets, int D, const float* __restrict__ gts,
                           const int32_t* __restrict__ lo, const int32_t* __restrict__ hi,
                           const int32_t* __restrict__ rank, float thr, int fmt,
                           float* __restrict__ best_iou, int32_t* __restrict__ best_gt,
                           int32_t* __restrict__ claim) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const float* r = dets + size_t(d) * 7;
  const float cls = r[6];
  const CBox db = yb_make_cbox(r[1], r[2], r[3], r[4], fmt);
  const float da = __fmul_rn(r[3], r[4]);
  float bi = 0.f;   // utils.py:240
  int bg = -1;
  for (int g = lo[d]; g < hi[d]; ++g) {
    const float* t = gts + size_t(g) * 7;
    if (!(t[6] == cls)) continue;  // same class list (utils.py:210-211), same image by range
    const CBox gb = yb_make_cbox(t[1], t[2], t[3], t[4], fmt);
    const float iou = yb_iou(db, da, gb, __fmul_rn(t[3], t[4]));
    if (iou > bi) { bi = iou; bg = g; }  // utils.py:247
  }
  best_iou[d] = bi;
  best_gt[d] = bg;
  if (bg >= 0 && bi > thr) atomicMin(&claim[bg], rank[d]);  // utils.py:252-255
}

__global__ void k_map_tp(int D, const int32_t* __restrict__ rank, float thr,
                         const float* __restrict__ best_iou, const int32_t* __restrict__ best_gt,
                         const int32_t* __restrict__ claim, float* __restrict__ tp) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const int bg = best_gt[d];
  tp[d] = (bg >= 0 && best_iou[d] > thr && claim[bg] == rank[d]) ? 1.f : 0.f;
}

// The tail of utils.py:262-272 for every class at once: one CTA per class walks that class's detections in evaluation
// order (tp_sorted[start[c] .. end[c])), cumTP by block scan with a carry, precision = cumTP / (cumTP + cumFP) and
// recall = cumTP / n_gt in fp32 as torch does, with (1, 0) prepended, and the trapezoid sum of torch.trapz:
// sum((p_i + p_{i-1}) * (r_i - r_{i-1})) / 2.  The products are fp32; their sum is accumulated in double (torch sums
// in fp32 pairwise order: the two agree to a few 1e-8).  Classes without ground truth get ap = 0 and are left out of
// the mean by the caller; a class with ground truth and no detections has AP 0 as well.
constexpr int AP_THREADS = 256;
__global__ void __launch_bounds__(AP_THREADS)
k_map_ap(const float* __restrict__ tp_sorted, const int32_t* __restrict__ start, const int32_t* __restrict__ end,
         const int32_t* __restrict__ n_gt, float* __restrict__ ap) {
  __shared__ double s_part[AP_THREADS / 32];
  const int c = blockIdx.x;
  const int s0 = start[c], s1 = end[c], ngt = n_gt[c];
  if (ngt <= 0 || s1 <= s0) {       // uniform
    if (threadIdx.x == 0) ap[c] = 0.f;
    return;
  }
  const float fngt = float(ngt);
  double acc = 0.0;
  int carry = 0;
  for (int base = s0; base < s1; base += AP_THREADS) {
    const int i = base + threadIdx.x;
    const bool in = i < s1;
    const int t = in ? int(tp_sorted[i]) : 0;
    int tot;
    const int excl = carry + block_exclusive_scan<AP_THREADS>(t, &tot);
    if (in) {
      const int k = i - s0;                                   // detections of the class before this one
      const float ctp = float(excl + t), ctp_prev = float(excl);
      const float prec = __fdiv_rn(ctp, float(k + 1)), rec = __fdiv_rn(ctp, fngt);
      const float prec_prev = k == 0 ? 1.f : __fdiv_rn(ctp_prev, float(k));
      const float rec_prev = k == 0 ? 0.f : __fdiv_rn(ctp_prev, fngt);
      acc += double(__fmul_rn(__fadd_rn(prec, prec_prev), __fsub_rn(rec, rec_prev)));
    }
    carry += tot;
    __syncthreads();   // block_exclusive_scan's shared words are reused by the next chunk
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tsum = 0.0;
    for (int w = 0; w < AP_THREADS / 32; ++w) tsum += s_part[w];
    ap[c] = float(tsum) * 0.5f;
  }
}

__global__ void k_fill_i32(int32_t* p, int n, int32_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

extern "C" int yolo_map_match(const float* dets, int D, const float* gts, int G,
                              const int32_t* det_gt_lo, const int32_t* det_gt_hi,
                              const int32_t* det_rank, float iou_thr, int box_format, float* tp,
                              float* best_iou, int32_t* best_gt, int32_t* gt_claim,
                              yb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  YB_REQUIRE(D >= 0 && G >= 0, "yolo_map_match: negative count");
  if (D == 0) return YB_OK;
  YB_REQUIRE(dets && det_gt_lo && det_gt_hi && det_rank && tp && best_iou && best_gt,
             "yolo_map_match: null pointer");
  YB_REQUIRE(G == 0 || (gts && gt_claim), "yolo_map_match: null gt pointer");
  if (G > 0) {
    k_fill_i32<<<yb_cdiv(G, 256), 256, 0, stream>>>(gt_claim, G, 0x7fffffff);
    YB_CHECK_LAUNCH();
  }
  k_map_best<<<yb_cdiv(D, 128), 128, 0, stream>>>(dets, D, gts, det_gt_lo, det_gt_hi, det_rank,
                                                  iou_thr, box_format, best_iou, best_gt, gt_claim);
  YB_CHECK_LAUNCH();
  k_map_tp<<<yb_cdiv(D, 256), 256, 0, stream>>>(D, det_rank, iou_thr, best_iou, best_gt, gt_claim, tp);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_map_ap(const float* tp_sorted, const int32_t* cls_start, const int32_t* cls_end, const int32_t* n_gt,
                           int num_classes, float* ap, yb_stream_t stream) {
  YB_REQUIRE(num_classes >= 0, "yolo_map_ap: negative class count");
  if (num_classes == 0) return YB_OK;
  YB_REQUIRE(cls_start && cls_end && n_gt && ap, "yolo_map_ap: null pointer");
  k_map_ap<<<num_classes, AP_THREADS, 0, (cudaStream_t)stream>>>(tp_sorted, cls_start, cls_end, n_gt, ap);
  YB_CHECK_LAUNCH();
  return YB_OK;
}
