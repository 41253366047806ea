// Accuracy reductions of the evaluation loop.  Replaces the per-scale mask-index gathers of
// utils.py:356-371 (check_model_accuracy): obj/no-obj masks from the target's objectness, class accuracy
// = argmax(logits) == label on object cells, objectness accuracy = (sigmoid(logit) > thr) on object /
// no-object cells.  One warp per (b, anchor, i, j) cell, counts reduced per block, six global atomics.
#include "common.cuh"

namespace {

struct AccParams {
  const float* head;
  const float* target;
  long long hs[5], ts[5];
  int batch, S, nc;
  float thr;
  unsigned long long* counts;  // correct_class, total_class, correct_obj, total_obj, correct_noobj, total_noobj
};

__global__ void __launch_bounds__(256) k_accuracy(const AccParams p) {
  __shared__ unsigned int s_cnt[6];
  if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long cell = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long per_img = 3ll * p.S * p.S;
  if (cell < per_img * p.batch) {
    const int b = int(cell / per_img);
    int r = int(cell - (long long)b * per_img);
    const int a = r / (p.S * p.S);
    r -= a * p.S * p.S;
    const int i = r / p.S, j = r - i * p.S;
    const float* t = p.target + b * p.ts[0] + a * p.ts[1] + i * p.ts[2] + j * p.ts[3];
    const float* h = p.head + b * p.hs[0] + a * p.hs[1] + i * p.hs[2] + j * p.hs[3];
    const float tobj = t[4 * p.ts[4]];
    const bool is_obj = tobj == 1.0f, is_noobj = tobj == 0.0f;  // utils.py:358-359 (-1 = ignore)
    if (is_obj || is_noobj) {  // warp-uniform
      const float logit = h[4 * p.hs[4]];
      const bool pred = (1.0f / (1.0f + expf(-logit))) > p.thr;  // utils.py:366
      if (is_obj) {
        float best = -INFINITY;
        int best_i = 0x7fffffff;
        bool best_nan = false;
        for (int c = lane; c < p.nc; c += 32) {  // argmax, first maximal index, NaN maximal (utils.py:362)
          const float v = h[(5 + c) * p.hs[4]];
          const bool vn = v != v;
          if (best_i == 0x7fffffff || (!best_nan && (vn || v > best))) { best = v; best_i = c; best_nan = vn; }
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
        if (lane == 0) {
          atomicAdd(&s_cnt[1], 1u);
          atomicAdd(&s_cnt[3], 1u);
          if (float(best_i) == t[5 * p.ts[4]]) atomicAdd(&s_cnt[0], 1u);
          if (pred) atomicAdd(&s_cnt[2], 1u);  // (pred == target 1.0)
        }
      } else if (lane == 0) {
        atomicAdd(&s_cnt[5], 1u);
        if (!pred) atomicAdd(&s_cnt[4], 1u);   // (pred == target 0.0)
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 6 && s_cnt[threadIdx.x]) atomicAdd(&p.counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

}  // namespace

extern "C" int yolo_accuracy_counts(const float* head, const int64_t* hstrides5_host, const float* target,
                                    const int64_t* tstrides5_host, int batch, int S, int nc, float obj_thr,
                                    unsigned long long* counts6, yb_stream_t stream) {
  YB_REQUIRE(head && target && hstrides5_host && tstrides5_host && counts6, "yolo_accuracy_counts: null pointer");
  YB_REQUIRE(batch >= 0 && S >= 1 && nc >= 1, "yolo_accuracy_counts: bad shape");
  if (batch == 0) return YB_OK;
  AccParams p;
  p.head = head; p.target = target;
  for (int k = 0; k < 5; ++k) { p.hs[k] = hstrides5_host[k]; p.ts[k] = tstrides5_host[k]; }
  p.batch = batch; p.S = S; p.nc = nc; p.thr = obj_thr; p.counts = counts6;
  const long long cells = 3ll * S * S * batch;
  k_accuracy<<<(unsigned)((cells + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
  YB_CHECK_LAUNCH();
  return YB_OK;
}
