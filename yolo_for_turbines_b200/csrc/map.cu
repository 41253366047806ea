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
