// K3: anchor decode.  Replaces utils.py:86-148 (cells_to_boxes): ~15 ATen
// element-wise kernels + cat + arange become one pass that reads each head row
// once with coalesced loads (a warp per (b, anchor, i, j) cell, lanes across the
// 5+nc channels) and writes one 24-byte [cx,cy,w,h,obj,cls] row.
// HBM-bound: (5+nc)*4 B read + 24 B written per cell.
#include "common.cuh"

namespace {

struct DecodeParams {
  const float* head;
  long long st[5];  // element strides of (B,3,S,S,C)
  int batch, S, nc;
  float anchors[6];
  int is_pred, writeback;
  float* out;
  int out_boxes_per_image, out_offset;
  float inv_s;  // (float)(1.0 / S): utils.py:125 multiplies by a Python double cast to fp32
};

__device__ __forceinline__ float sigmoid_f32(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256) k_decode(const DecodeParams p) {
  const int warps_per_block = blockDim.x >> 5;
  const long long cell = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long cells_per_image = 3ll * p.S * p.S;
  if (cell >= cells_per_image * p.batch) return;  // warp-uniform
  const int b = int(cell / cells_per_image);
  int r = int(cell - (long long)b * cells_per_image);
  const int a = r / (p.S * p.S);
  r -= a * p.S * p.S;
  const int i = r / p.S, j = r - i * p.S;
  float* row = const_cast<float*>(p.head) + b * p.st[0] + a * p.st[1] + i * p.st[2] + j * p.st[3];
  const long long cs = p.st[4];
  float* o = p.out + (size_t(b) * p.out_boxes_per_image + p.out_offset + size_t(a) * p.S * p.S +
                      size_t(i) * p.S + j) * 6;

  float cls;
  if (p.is_pred) {
    // argmax over raw logits (utils.py:112): first maximal index, NaN counts as maximal
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    bool best_nan = false;
    for (int c = lane; c < p.nc; c += 32) {
      const float v = row[(5 + c) * cs];
      const bool vn = v != v;
      if (best_i == 0x7fffffff || (!best_nan && (vn || v > best))) {
        best = v; best_i = c; best_nan = vn;
      }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, d);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, d);
      const bool on = ov != ov;
      bool take;
      if (oi == 0x7fffffff) take = false;
      else if (best_i == 0x7fffffff) take = true;
      else if (best_nan || on) take = on && (!best_nan || oi < best_i);
      else take = (ov > best) || (ov == best && oi < best_i);
      if (take) { best = ov; best_i = oi; best_nan = on; }
    }
    cls = (best_i == 0x7fffffff) ? 0.f : float(best_i);
  } else {
    cls = row[5 * cs];  // utils.py:116
  }
  if (lane == 0) {
    float t0 = row[0], t1 = row[cs], t2 = row[2 * cs], t3 = row[3 * cs], t4 = row[4 * cs];
    float obj = t4;
    if (p.is_pred) {
      t0 = sigmoid_f32(t0);                         // utils.py:106
      t1 = sigmoid_f32(t1);
      t2 = __fmul_rn(expf(t2), p.anchors[2 * a]);   // utils.py:110
      t3 = __fmul_rn(expf(t3), p.anchors[2 * a + 1]);
      obj = sigmoid_f32(t4);                        // utils.py:111
      if (p.writeback) {                            // the reference mutates its input
        row[0] = t0; row[cs] = t1; row[2 * cs] = t2; row[3 * cs] = t3;
      }
    }
    o[0] = __fmul_rn(p.inv_s, __fadd_rn(t0, float(j)));  // utils.py:125
    o[1] = __fmul_rn(p.inv_s, __fadd_rn(t1, float(i)));  // utils.py:142
    o[2] = __fmul_rn(p.inv_s, t2);                       // utils.py:143
    o[3] = __fmul_rn(p.inv_s, t3);
    o[4] = obj;
    o[5] = cls;
  }
}

}  // namespace

extern "C" int yolo_decode(const float* head, const int64_t* strides5_host, int batch, int S, int nc,
                           const float* anchors6_host, int is_pred, int writeback, float* out,
                           int out_boxes_per_image, int out_offset, yb_stream_t stream) {
  YB_REQUIRE(head && strides5_host && anchors6_host && out, "yolo_decode: null pointer");
  YB_REQUIRE(batch >= 0 && S >= 1 && nc >= 1, "yolo_decode: bad shape (batch %d S %d nc %d)", batch, S, nc);
  YB_REQUIRE(out_boxes_per_image >= out_offset + 3 * S * S, "yolo_decode: output rows do not fit");
  if (batch == 0) return YB_OK;
  DecodeParams p;
  p.head = head;
  for (int k = 0; k < 5; ++k) p.st[k] = strides5_host[k];
  p.batch = batch; p.S = S; p.nc = nc;
  for (int k = 0; k < 6; ++k) p.anchors[k] = anchors6_host[k];
  p.is_pred = is_pred; p.writeback = writeback;
  p.out = out; p.out_boxes_per_image = out_boxes_per_image; p.out_offset = out_offset;
  p.inv_s = (float)(1.0 / (double)S);
  const long long cells = 3ll * S * S * batch;
  const int wpb = 8;
  k_decode<<<(unsigned)((cells + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(p);
  YB_CHECK_LAUNCH();
  return YB_OK;
}
