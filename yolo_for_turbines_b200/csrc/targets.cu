// Training-target encoder: what `YOLODataset.__getitem__` (code/dataset.py:119-167) and the target
// half of `collate_fn` (code/utils.py:694-700) build on the CPU, for a whole batch on the device.
//
// The reference loops over an image's boxes IN ORDER and over the 9 anchors by descending `iou_aligned`
// (utils.py:22-36); a box claims the best free anchor of every scale (obj = 1, [S*x - j, S*y - i, w*S, h*S], class)
// and marks other free anchors with IoU > 0.5 as ignore (-1).  "Free" reads element 0 (the x offset) of the cell,
// not its objectness flag (dataset.py:143) -- reproduced.  Later boxes see earlier boxes' cells, so the unit of
// parallelism is (image, scale): one thread walks the image's boxes and the three anchors of its scale; the scales
// are independent because `has_anchor` is per scale.  Cell indices and offsets are computed in double precision like
// the reference's Python floats, IoUs in fp32 with the reference's operation order (no FMA contraction).
#include "common.cuh"

namespace {

struct TargetParams {
  const double* boxes;     // [total][5] x, y, w, h, class (YOLO format, fractions of the image)
  const int32_t* offsets;  // [batch + 1]
  float anchors[18];       // 9 x (w, h), scale-major (anchors[0] + anchors[1] + anchors[2], dataset.py:39)
  float* t[3];             // (batch, 3, S, S, 6) fp32, zeroed by the launcher
  int S[3];
  int batch;
  float ignore_thr;
};

__global__ void k_encode_targets(const TargetParams p) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= p.batch * 3) return;
  const int b = gid / 3, scale = gid - 3 * b;
  const int S = p.S[scale];
  float* T = p.t[scale] + size_t(b) * 3 * S * S * 6;
  for (int n = p.offsets[b]; n < p.offsets[b + 1]; ++n) {
    const double x = p.boxes[5 * n], y = p.boxes[5 * n + 1], wd = p.boxes[5 * n + 2], hd = p.boxes[5 * n + 3];
    const float w = float(wd), h = float(hd);  // torch.tensor(box[2:4]) is fp32 (dataset.py:130)
    float iou[9];
    int order[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) {
      const float aw = p.anchors[2 * a], ah = p.anchors[2 * a + 1];
      const float inter = __fmul_rn(fminf(w, aw), fminf(h, ah));                                  // utils.py:34
      const float uni = __fsub_rn(__fadd_rn(__fmul_rn(w, h), __fmul_rn(aw, ah)), inter);          // utils.py:35
      iou[a] = __fdiv_rn(inter, uni);
      order[a] = a;
    }
    for (int u = 1; u < 9; ++u) {  // stable descending insertion sort (argsort, dataset.py:131)
      const int key = order[u];
      int v = u - 1;
      while (v >= 0 && iou[order[v]] < iou[key]) { order[v + 1] = order[v]; --v; }
      order[v + 1] = key;
    }
    const int i = int(double(S) * y), j = int(double(S) * x);                                     // dataset.py:142
    if (i < 0 || i >= S || j < 0 || j >= S) continue;  // the reference would raise IndexError; boxes are clipped upstream
    bool has_anchor = false;
    for (int u = 0; u < 9; ++u) {
      const int a = order[u];
      if (a / 3 != scale) continue;
      float* cell = T + ((size_t(a % 3) * S + i) * S + j) * 6;
      const bool taken = cell[0] != 0.f;                                                          // dataset.py:143
      if (!taken && !has_anchor) {
        cell[0] = float(double(S) * x - double(j));                                               // dataset.py:148-156
        cell[1] = float(double(S) * y - double(i));
        cell[2] = float(wd * double(S));
        cell[3] = float(hd * double(S));
        cell[4] = 1.f;
        cell[5] = float(int(p.boxes[5 * n + 4]));
        has_anchor = true;
      } else if (!taken && iou[a] > p.ignore_thr) {
        cell[4] = -1.f;                                                                           // dataset.py:160-161
      }
    }
  }
}

}  // namespace

extern "C" int yolo_encode_targets(const double* boxes, const int32_t* offsets, int batch, const float* anchors18_host,
                                   int S0, int S1, int S2, float ignore_iou_threshold, float* t0, float* t1, float* t2,
                                   yb_stream_t stream_) {
  YB_REQUIRE(offsets && anchors18_host && t0 && t1 && t2 && batch >= 0 && S0 >= 1 && S1 >= 1 && S2 >= 1,
             "yolo_encode_targets: bad argument");
  if (batch == 0) return YB_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  TargetParams p;
  p.boxes = boxes; p.offsets = offsets; p.batch = batch; p.ignore_thr = ignore_iou_threshold;
  for (int k = 0; k < 18; ++k) p.anchors[k] = anchors18_host[k];
  p.t[0] = t0; p.t[1] = t1; p.t[2] = t2;
  p.S[0] = S0; p.S[1] = S1; p.S[2] = S2;
  for (int s = 0; s < 3; ++s)
    YB_CHECK_CUDA(cudaMemsetAsync(p.t[s], 0, size_t(batch) * 3 * p.S[s] * p.S[s] * 6 * sizeof(float), stream));
  k_encode_targets<<<yb_cdiv(batch * 3, 64), 64, 0, stream>>>(p);
  YB_CHECK_LAUNCH();
  return YB_OK;
}
