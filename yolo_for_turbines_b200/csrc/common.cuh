// Shared host/device helpers for libyolo_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "yolo_b200.h"

void yb_set_error(const char* fmt, ...);

#define YB_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      yb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__,     \
                   __LINE__);                                                            \
      return YB_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)
#define YB_CHECK_LAUNCH() YB_CHECK_CUDA(cudaGetLastError())
#define YB_REQUIRE(cond, ...)                                                            \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      yb_set_error(__VA_ARGS__);                                                         \
      return YB_ERR_INVALID;                                                             \
    }                                                                                    \
  } while (0)

static inline int yb_cdiv(int a, int b) { return (a + b - 1) / b; }

// Carves 256-byte aligned sub-buffers out of a caller-owned workspace.  Run
// once with base == nullptr to size it, once with the real pointer to place.
struct WsCarver {
  char* base;
  size_t off;
  explicit WsCarver(void* b) : base(static_cast<char*>(b)), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
  size_t bytes() const { return (off + 255) & ~size_t(255); }
};

// ---- device helpers ---------------------------------------------------------

// Exclusive scan of one int per thread over a block of NT threads (NT multiple
// of 32, <= 1024).  Returns the exclusive prefix; *total gets the block sum.
template <int NT>
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  __syncthreads();  // readers of the previous call are done with warp_sums
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int s = (lane < NT / 32) ? warp_sums[lane] : 0;
    int si = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, si, d);
      if (lane >= d) si += t;
    }
    warp_sums[lane] = si - s;           // exclusive prefix of the warp sums
    if (lane == 31) warp_sums[32] = si;  // block total
  }
  __syncthreads();
  if (total) *total = warp_sums[32];
  return warp_sums[warp] + incl - v;
}

// The reference compares IoUs produced by this exact fp32 op sequence
// (utils.py:57-83); every op is an explicit round-to-nearest intrinsic so that
// nvcc cannot contract a mul+add into an FMA and change a kept index.
struct CBox {  // corner form derived once per box
  float x1, y1, x2, y2;
};
__device__ __forceinline__ CBox yb_make_cbox(float x, float y, float w, float h, int fmt) {
  CBox c;
  if (fmt == YB_BOX_CENTER) {  // utils.py:60,63: xy - wh / 2
    c.x1 = __fsub_rn(x, __fmul_rn(w, 0.5f));
    c.y1 = __fsub_rn(y, __fmul_rn(h, 0.5f));
  } else {  // utils.py:66-67: boxes are taken as top-left x,y + w,h
    c.x1 = x;
    c.y1 = y;
  }
  c.x2 = __fadd_rn(c.x1, w);  // utils.py:72-73
  c.y2 = __fadd_rn(c.y1, h);
  return c;
}
// torch.max / torch.min propagate NaN (fmaxf/fminf do not).
__device__ __forceinline__ float yb_nanmax(float a, float b) { return (a > b || a != a) ? a : b; }
__device__ __forceinline__ float yb_nanmin(float a, float b) { return (a < b || a != a) ? a : b; }
// torch.clamp(min=0) also propagates NaN.
__device__ __forceinline__ float yb_clamp0(float d) { return d < 0.f ? 0.f : d; }

__device__ __forceinline__ float yb_iou(const CBox& a, float area_a, const CBox& b, float area_b) {
  const float xA = yb_nanmax(a.x1, b.x1), yA = yb_nanmax(a.y1, b.y1);  // utils.py:70-71
  const float xB = yb_nanmin(a.x2, b.x2), yB = yb_nanmin(a.y2, b.y2);  // utils.py:72-73
  const float iw = yb_clamp0(__fsub_rn(xB, xA));                       // utils.py:75
  const float ih = yb_clamp0(__fsub_rn(yB, yA));                       // utils.py:76
  const float inter = __fmul_rn(iw, ih);                               // utils.py:77
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);       // utils.py:81
  return __fdiv_rn(inter, __fadd_rn(uni, 1e-6f));                      // utils.py:83
}

// Order-preserving map float -> u32 (ascending), with -0.0 folded onto +0.0 so
// that equal floats get equal keys, as Python's float comparison sees them.
__device__ __forceinline__ uint32_t yb_float_key_asc(float f) {
  f = __fadd_rn(f, 0.0f);  // -0.0 -> +0.0
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// nn.BatchNorm2d training-mode finalize of one channel from its batch sums (shared by csrc/train.cu and by the conv
// kernel's last-CTA epilogue, csrc/conv2.cu): biased variance for the normalisation, momentum update of the running
// statistics with the UNBIASED variance.
struct BnFinalize {
  long long P;
  const float *gamma, *beta;
  float eps, momentum;
  float *running_mean, *running_var, *mean, *rstd, *scale, *bias;
};
__device__ __forceinline__ void bn_finalize_channel(const BnFinalize& f, int c, double s, double q) {
  const double n = double(f.P);
  const double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = float(1.0 / sqrt(var + double(f.eps)));
  const float sc = f.gamma[c] * rstd;
  f.mean[c] = float(mean);
  f.rstd[c] = rstd;
  f.scale[c] = sc;
  f.bias[c] = f.beta[c] - float(mean) * sc;
  if (f.running_mean) {
    const double unbiased = f.P > 1 ? var * n / (n - 1.0) : var;
    f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * float(mean);
    f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * float(unbiased);
  }
}
