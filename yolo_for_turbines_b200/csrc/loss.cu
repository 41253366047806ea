// K8 (forward only): the four terms of YOLOLoss.forward (code/loss.py:29-81) for one scale in one pass.
// Replaces the mask-index gathers + BCEWithLogits / MSE / CrossEntropy launches of loss.py:42-76 with a
// thread per (b, anchor, i, j) cell accumulating six sums in double precision:
//   [0] sum softplus(logit_obj)      over no-object cells  (BCEWithLogits with target 0, loss.py:54)
//   [1] number of no-object cells
//   [2] sum (logit_obj - iou)^2      over object cells     (loss.py:60-67; iou = calc_iou "center", detached)
//   [3] sum of the 4 box terms       over object cells     (loss.py:71-73, with the reference's index quirk:
//                                     sigmoid is applied to entries 1:3 = ty and tw; tx and th stay raw)
//   [4] sum cross-entropy            over object cells     (loss.py:76)
//   [5] number of object cells
// The host turns them into [5*box, 1*object, 0.5*no_obj, 1*class] exactly as loss.py:78-81.  With mutate=1 a
// second launch reproduces the reference's in-place side effects (loss.py:71-72) when any object cell exists.
#include "common.cuh"

namespace {

struct LossParams {
  float* pred;
  float* target;
  long long ps[5], ts[5];
  int batch, S, nc;
  float anchors[6];
  int mutate;
  double* sums;
};

__global__ void __launch_bounds__(256) k_loss_fwd(const LossParams p) {
  __shared__ double s_sum[6];
  if (threadIdx.x < 6) s_sum[threadIdx.x] = 0.0;
  __syncthreads();
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_img = 3ll * p.S * p.S;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  if (cell < per_img * p.batch) {
    const int b = int(cell / per_img);
    int r = int(cell - (long long)b * per_img);
    const int a = r / (p.S * p.S);
    r -= a * p.S * p.S;
    const int i = r / p.S, j = r - i * p.S;
    float* t = p.target + b * p.ts[0] + a * p.ts[1] + i * p.ts[2] + j * p.ts[3];
    float* q = p.pred + b * p.ps[0] + a * p.ps[1] + i * p.ps[2] + j * p.ps[3];
    const long long tc = p.ts[4], pc = p.ps[4];
    const float tobj = t[4 * tc];
    if (tobj == 0.0f) {  // loss.py:43,54
      const float x = q[4 * pc];
      acc[0] = double(fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))));  // BCEWithLogits(x, 0) = softplus(x)
      acc[1] = 1.0;
    } else if (tobj == 1.0f) {  // loss.py:42
      const float tx = q[0], ty = q[pc], tw = q[2 * pc], th = q[3 * pc], to = q[4 * pc];
      const float gx = t[0], gy = t[tc], gw = t[2 * tc], gh = t[3 * tc];
      const float aw = p.anchors[2 * a], ah = p.anchors[2 * a + 1];
      const float sx = 1.f / (1.f + expf(-tx)), sy = 1.f / (1.f + expf(-ty)), sw = 1.f / (1.f + expf(-tw));
      // loss.py:60-64: IoU of [sigmoid(xy), exp(wh)*anchor] against the target box, "center" format
      const float pw = __fmul_rn(expf(tw), aw), ph = __fmul_rn(expf(th), ah);
      const CBox pb = yb_make_cbox(sx, sy, pw, ph, YB_BOX_CENTER);
      const CBox gb = yb_make_cbox(gx, gy, gw, gh, YB_BOX_CENTER);
      const float iou = yb_iou(pb, __fmul_rn(pw, ph), gb, __fmul_rn(gw, gh));
      const float dobj = to - iou * tobj;  // loss.py:67 (MSE on the raw logit)
      acc[2] = double(dobj) * dobj;
      // loss.py:71-73
      const float lw = logf(1e-16f + gw / aw), lh = logf(1e-16f + gh / ah);
      const float d0 = tx - gx, d1 = sy - gy, d2 = sw - lw, d3 = th - lh;
      acc[3] = double(d0) * d0 + double(d1) * d1 + double(d2) * d2 + double(d3) * d3;
      // loss.py:76 cross entropy = logsumexp(logits) - logit[label]
      float mx = -INFINITY;
      for (int c = 0; c < p.nc; ++c) mx = fmaxf(mx, q[(5 + c) * pc]);
      float se = 0.f;
      for (int c = 0; c < p.nc; ++c) se += expf(q[(5 + c) * pc] - mx);
      const int label = int(t[5 * tc]);
      const float picked = (label >= 0 && label < p.nc) ? q[(5 + label) * pc] : __int_as_float(0x7fc00000);
      acc[4] = double(mx + logf(se) - picked);
      acc[5] = 1.0;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double v = acc[k];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&s_sum[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 6 && s_sum[threadIdx.x] != 0.0) atomicAdd(&p.sums[threadIdx.x], s_sum[threadIdx.x]);
}

// loss.py:71-72 mutate predictions[..., 1:3] and targets[..., 2:4] of EVERY cell, but only inside
// `if obj_mask.any()`: applied afterwards, conditioned on the object count this launch produced.
__global__ void __launch_bounds__(256) k_loss_mutate(const LossParams p) {
  if (!(p.sums[5] > 0.0)) return;
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_img = 3ll * p.S * p.S;
  if (cell >= per_img * p.batch) return;
  const int b = int(cell / per_img);
  int r = int(cell - (long long)b * per_img);
  const int a = r / (p.S * p.S);
  r -= a * p.S * p.S;
  const int i = r / p.S, j = r - i * p.S;
  float* t = p.target + b * p.ts[0] + a * p.ts[1] + i * p.ts[2] + j * p.ts[3];
  float* q = p.pred + b * p.ps[0] + a * p.ps[1] + i * p.ps[2] + j * p.ps[3];
  q[p.ps[4]] = 1.f / (1.f + expf(-q[p.ps[4]]));
  q[2 * p.ps[4]] = 1.f / (1.f + expf(-q[2 * p.ps[4]]));
  t[2 * p.ts[4]] = logf(1e-16f + t[2 * p.ts[4]] / p.anchors[2 * a]);
  t[3 * p.ts[4]] = logf(1e-16f + t[3 * p.ts[4]] / p.anchors[2 * a + 1]);
}

}  // namespace

extern "C" int yolo_loss_fwd(float* pred, const int64_t* pstrides5_host, float* target, const int64_t* tstrides5_host,
                             int batch, int S, int nc, const float* anchors6_host, int mutate, double* sums6,
                             yb_stream_t stream) {
  YB_REQUIRE(pred && target && pstrides5_host && tstrides5_host && anchors6_host && sums6, "yolo_loss_fwd: null pointer");
  YB_REQUIRE(batch >= 0 && S >= 1 && nc >= 1, "yolo_loss_fwd: bad shape");
  if (batch == 0) return YB_OK;
  LossParams p;
  p.pred = pred; p.target = target;
  for (int k = 0; k < 5; ++k) { p.ps[k] = pstrides5_host[k]; p.ts[k] = tstrides5_host[k]; }
  p.batch = batch; p.S = S; p.nc = nc;
  for (int k = 0; k < 6; ++k) p.anchors[k] = anchors6_host[k];
  p.mutate = mutate; p.sums = sums6;
  const long long cells = 3ll * S * S * batch;
  k_loss_fwd<<<(unsigned)((cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  YB_CHECK_LAUNCH();
  if (mutate) {  // sums6 must have been zero before this call: the object count gates the side effects
    k_loss_mutate<<<(unsigned)((cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    YB_CHECK_LAUNCH();
  }
  return YB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// K8 backward: d(sum of the four weighted loss terms of one scale) / d(pred), loss.py:54-81 under autograd
// (train.py:56-67).  sums6 are the counts the forward launch produced (n_noobj = sums6[1], n_obj = sums6[5]):
//   no-object cell : d p4 = lambda_noobj * sigmoid(p4) / n_noobj
//   object cell    : d p4 = lambda_obj * 2 (p4 - iou) / n_obj            (iou is detached, loss.py:64)
//                    d p0..3 = lambda_box * 2 d_k chain_k / (4 n_obj)    (sigmoid on entries 1:3 only, loss.py:71)
//                    d p5+c = lambda_class * (softmax_c - [c == label]) / n_obj
//   ignore cell (-1): zero.
// Every one of the 5+nc entries of every cell is written (zeros included); out_bf16 selects the element type
// of dpred (the trainer writes the head conv's bf16 dz directly; the autograd wrapper asks for fp32).
namespace {

struct LossBwdParams {
  const float* pred;
  const float* target;
  long long ps[5], ts[5], ds[5];
  int batch, S, nc;
  float anchors[6];
  const double* sums;
  float gscale;
  float tw[4];  // upstream gradient of the [box, object, no-object, class] terms (1 when the terms are just summed)
  void* dpred;
  int out_bf16;
};

__device__ __forceinline__ void put_grad(const LossBwdParams& p, long long off, float v) {
  if (p.out_bf16) static_cast<__nv_bfloat16*>(p.dpred)[off] = __float2bfloat16_rn(v);
  else static_cast<float*>(p.dpred)[off] = v;
}

__global__ void __launch_bounds__(256) k_loss_bwd(const LossBwdParams p) {
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_img = 3ll * p.S * p.S;
  if (cell >= per_img * p.batch) return;
  const int b = int(cell / per_img);
  int r = int(cell - (long long)b * per_img);
  const int a = r / (p.S * p.S);
  r -= a * p.S * p.S;
  const int i = r / p.S, j = r - i * p.S;
  const float* t = p.target + b * p.ts[0] + a * p.ts[1] + i * p.ts[2] + j * p.ts[3];
  const float* q = p.pred + b * p.ps[0] + a * p.ps[1] + i * p.ps[2] + j * p.ps[3];
  const long long d0 = b * p.ds[0] + a * p.ds[1] + i * p.ds[2] + j * p.ds[3];
  const long long tc = p.ts[4], pc = p.ps[4], dc = p.ds[4];
  const float n_noobj = float(p.sums[1]), n_obj = float(p.sums[5]);
  const float tobj = t[4 * tc];
  if (tobj == 1.0f && n_obj > 0.f) {
    const float tx = q[0], ty = q[pc], tw = q[2 * pc], th = q[3 * pc], to = q[4 * pc];
    const float gx = t[0], gy = t[tc], gw = t[2 * tc], gh = t[3 * tc];
    const float aw = p.anchors[2 * a], ah = p.anchors[2 * a + 1];
    const float sx = 1.f / (1.f + expf(-tx)), sy = 1.f / (1.f + expf(-ty)), sw = 1.f / (1.f + expf(-tw));
    const float pw = __fmul_rn(expf(tw), aw), ph = __fmul_rn(expf(th), ah);
    const CBox pb = yb_make_cbox(sx, sy, pw, ph, YB_BOX_CENTER);
    const CBox gb = yb_make_cbox(gx, gy, gw, gh, YB_BOX_CENTER);
    const float iou = yb_iou(pb, __fmul_rn(pw, ph), gb, __fmul_rn(gw, gh));
    const float lw = logf(1e-16f + gw / aw), lh = logf(1e-16f + gh / ah);
    const float kb = p.gscale * p.tw[0] * 5.f * 2.f / (4.f * n_obj);
    put_grad(p, d0, kb * (tx - gx));
    put_grad(p, d0 + dc, kb * (sy - gy) * sy * (1.f - sy));
    put_grad(p, d0 + 2 * dc, kb * (sw - lw) * sw * (1.f - sw));
    put_grad(p, d0 + 3 * dc, kb * (th - lh));
    put_grad(p, d0 + 4 * dc, p.gscale * p.tw[1] * 2.f * (to - iou * tobj) / n_obj);
    float mx = -INFINITY;
    for (int c = 0; c < p.nc; ++c) mx = fmaxf(mx, q[(5 + c) * pc]);
    float se = 0.f;
    for (int c = 0; c < p.nc; ++c) se += expf(q[(5 + c) * pc] - mx);
    const int label = int(t[5 * tc]);
    const float gc = p.gscale * p.tw[3] / n_obj;
    const float inv = gc / se;
    for (int c = 0; c < p.nc; ++c)
      put_grad(p, d0 + (5 + c) * dc, expf(q[(5 + c) * pc] - mx) * inv - (c == label ? gc : 0.f));
  } else {
    float g4 = 0.f;
    if (tobj == 0.0f) g4 = p.gscale * p.tw[2] * 0.5f / (n_noobj * (1.f + expf(-q[4 * pc])));
    for (int c = 0; c < 5 + p.nc; ++c) put_grad(p, d0 + c * dc, c == 4 ? g4 : 0.f);
  }
}

}  // namespace

extern "C" int yolo_loss_bwd(const float* pred, const int64_t* pstrides5_host, const float* target,
                             const int64_t* tstrides5_host, int batch, int S, int nc, const float* anchors6_host,
                             const double* sums6, float grad_scale, const float* term_scales4_host, void* dpred,
                             const int64_t* dstrides5_host, int out_bf16, yb_stream_t stream) {
  YB_REQUIRE(pred && target && pstrides5_host && tstrides5_host && anchors6_host && sums6 && dpred && dstrides5_host,
             "yolo_loss_bwd: null pointer");
  YB_REQUIRE(batch >= 0 && S >= 1 && nc >= 1, "yolo_loss_bwd: bad shape");
  if (batch == 0) return YB_OK;
  LossBwdParams p;
  p.pred = pred; p.target = target;
  for (int k = 0; k < 5; ++k) { p.ps[k] = pstrides5_host[k]; p.ts[k] = tstrides5_host[k]; p.ds[k] = dstrides5_host[k]; }
  p.batch = batch; p.S = S; p.nc = nc;
  for (int k = 0; k < 6; ++k) p.anchors[k] = anchors6_host[k];
  p.sums = sums6; p.gscale = grad_scale; p.dpred = dpred; p.out_bf16 = out_bf16;
  for (int k = 0; k < 4; ++k) p.tw[k] = term_scales4_host ? term_scales4_host[k] : 1.f;
  const long long cells = 3ll * S * S * batch;
  k_loss_bwd<<<(unsigned)((cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  YB_CHECK_LAUNCH();
  return YB_OK;
}
