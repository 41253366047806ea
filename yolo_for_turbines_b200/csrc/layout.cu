// Layout/packing kernels around K1: weight repack (feeds the Darknet loader,
// model.py:293-328), eval-mode BatchNorm folding (model.py:61,84), the NCHW fp32
// <-> NHWC bf16 boundary conversions (the reference's tensors are NCHW fp32) and
// the stem "patchify" that turns the Cin=3 first conv (model.py:21) into a K=32
// GEMM.  All are streaming, HBM-bound kernels.
#include "common.cuh"

namespace {

__global__ void k_pack_weights(const float* __restrict__ w, int c_out, int c_in, int ks,
                               int c_out_pad, int c_in_pad, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)c_out_pad * ks * ks * c_in_pad;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = int(i % c_in_pad);
  long long r = i / c_in_pad;
  const int tap = int(r % (ks * ks));
  const int o = int(r / (ks * ks));
  float v = 0.f;
  if (o < c_out && c < c_in) v = w[((size_t(o) * c_in + c) * ks + tap / ks) * ks + tap % ks];
  out[i] = __float2bfloat16_rn(v);
}

__global__ void k_pack_stem(const float* __restrict__ w, int c_out, int c_in, int c_out_pad,
                            __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c_out_pad * 32) return;
  const int k = i & 31, o = i >> 5;
  float v = 0.f;
  if (o < c_out && k < 9 * c_in) {
    const int tap = k / c_in, c = k - tap * c_in;
    v = w[((size_t(o) * c_in + c) * 3 + tap / 3) * 3 + tap % 3];
  }
  out[i] = __float2bfloat16_rn(v);
}

__global__ void k_fold_bn(const float* __restrict__ g, const float* __restrict__ b,
                          const float* __restrict__ mean, const float* __restrict__ var,
                          const float* __restrict__ conv_bias, float eps, int c, int c_pad,
                          float* __restrict__ scale, float* __restrict__ bias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c_pad) return;
  float s = 0.f, t = 0.f;
  if (i < c) {
    if (g) {
      s = g[i] / sqrtf(var[i] + eps);
      t = b[i] - mean[i] * s;
    } else {
      s = 1.f;
      t = conv_bias ? conv_bias[i] : 0.f;
    }
  }
  scale[i] = s;
  bias[i] = t;
}

__global__ void k_nchw_to_nhwc(const float* __restrict__ x, int batch, int c, int h, int w,
                               int c_pad, int pitch, __nv_bfloat16* __restrict__ y,
                               uint32_t* __restrict__ status) {
  const long long total = (long long)batch * h * w * c_pad;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = int(i % c_pad);
  const long long pix = i / c_pad;
  const int hw = h * w;
  const int b = int(pix / hw), r = int(pix - (long long)b * hw);
  float v = 0.f;
  if (ch < c) {
    v = x[(size_t(b) * c + ch) * hw + r];
    if (v != v && status) atomicOr(status, YB_STATUS_NAN_INPUT);
  }
  y[size_t(pix) * pitch + ch] = __float2bfloat16_rn(v);
}

__global__ void k_nhwc_to_nchw(const void* __restrict__ x, int is_f32, int batch, int c, int h,
                               int w, int pitch, float* __restrict__ y) {
  const long long total = (long long)batch * c * h * w;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int hw = h * w;
  const int r = int(i % hw);
  const long long t = i / hw;
  const int ch = int(t % c), b = int(t / c);
  const size_t src = (size_t(b) * hw + r) * pitch + ch;
  y[i] = is_f32 ? static_cast<const float*>(x)[src]
                : __bfloat162float(static_cast<const __nv_bfloat16*>(x)[src]);
}

// One thread per pixel: gathers the 3x3 window of every input channel (<= 3),
// k = (kh*3+kw)*C + c, and writes one 64-byte row.  Reads are coalesced along w.
template <int C>
__global__ void __launch_bounds__(256)
k_patchify(const float* __restrict__ x, int batch, int h, int w,
           __nv_bfloat16* __restrict__ y, uint32_t* __restrict__ status) {
  constexpr int c = C;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)batch * h * w;
  if (pix >= total) return;
  const int hw = h * w;
  const int b = int(pix / hw), r = int(pix - (long long)b * hw);
  const int i = r / w, j = r - i * w;
  float v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = 0.f;
  bool nan = false;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int ii = i + kh - 1;
    if (ii < 0 || ii >= h) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int jj = j + kw - 1;
      if (jj < 0 || jj >= w) continue;
#pragma unroll
      for (int ch = 0; ch < c; ++ch) {
        const float t = x[(size_t(b) * c + ch) * hw + size_t(ii) * w + jj];
        v[(kh * 3 + kw) * c + ch] = t;
        if (kh == 1 && kw == 1) nan |= (t != t);  // each input element checked exactly once
      }
    }
  }
  if (nan && status) atomicOr(status, YB_STATUS_NAN_INPUT);
  uint4* out = reinterpret_cast<uint4*>(y + size_t(pix) * 32);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
    __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0);
    u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2);
    u.w = *reinterpret_cast<uint32_t*>(&p3);
    out[q] = u;
  }
}

inline unsigned blocks_for(long long total, int threads) {
  return (unsigned)((total + threads - 1) / threads);
}

}  // namespace

extern "C" int yolo_pack_weights(const float* w_oihw, int c_out, int c_in, int ksize, int c_out_pad,
                                 int c_in_pad, void* w_packed, yb_stream_t stream) {
  YB_REQUIRE(w_oihw && w_packed, "yolo_pack_weights: null pointer");
  YB_REQUIRE(c_out >= 1 && c_in >= 1 && (ksize == 1 || ksize == 3) && c_out_pad >= c_out &&
                 c_in_pad >= c_in, "yolo_pack_weights: bad shape");
  const long long total = (long long)c_out_pad * ksize * ksize * c_in_pad;
  k_pack_weights<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, c_out, c_in, ksize, c_out_pad, c_in_pad, static_cast<__nv_bfloat16*>(w_packed));
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_pack_stem_weights(const float* w_oihw, int c_out, int c_in, int c_out_pad,
                                      void* w_packed, yb_stream_t stream) {
  YB_REQUIRE(w_oihw && w_packed, "yolo_pack_stem_weights: null pointer");
  YB_REQUIRE(c_in >= 1 && 9 * c_in <= 32 && c_out >= 1 && c_out_pad >= c_out,
             "yolo_pack_stem_weights: needs 9*c_in <= 32");
  k_pack_stem<<<blocks_for((long long)c_out_pad * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, c_out, c_in, c_out_pad, static_cast<__nv_bfloat16*>(w_packed));
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var,
                            const float* conv_bias, float eps, int c, int c_pad, float* scale,
                            float* bias, yb_stream_t stream) {
  YB_REQUIRE(scale && bias && c >= 1 && c_pad >= c, "yolo_fold_bn: bad arguments");
  YB_REQUIRE(!gamma || (beta && mean && var), "yolo_fold_bn: incomplete BatchNorm tensors");
  k_fold_bn<<<blocks_for(c_pad, 256), 256, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, conv_bias,
                                                                     eps, c, c_pad, scale, bias);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_nchw_to_nhwc_bf16(const float* x, int batch, int c, int h, int w, int c_pad,
                                      int out_pitch, void* y, uint32_t* status, yb_stream_t stream) {
  YB_REQUIRE(x && y && batch >= 1 && c >= 1 && h >= 1 && w >= 1 && c_pad >= c && out_pitch >= c_pad,
             "yolo_nchw_to_nhwc_bf16: bad arguments");
  const long long total = (long long)batch * h * w * c_pad;
  k_nchw_to_nhwc<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      x, batch, c, h, w, c_pad, out_pitch, static_cast<__nv_bfloat16*>(y), status);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_nhwc_to_nchw_f32(const void* x, int in_is_fp32, int batch, int c, int h, int w,
                                     int in_pitch, float* y, yb_stream_t stream) {
  YB_REQUIRE(x && y && batch >= 1 && c >= 1 && h >= 1 && w >= 1 && in_pitch >= c,
             "yolo_nhwc_to_nchw_f32: bad arguments");
  const long long total = (long long)batch * c * h * w;
  k_nhwc_to_nchw<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, in_is_fp32, batch, c, h, w,
                                                                          in_pitch, y);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_input_patchify(const float* x, int batch, int c, int h, int w, void* y,
                                   uint32_t* status, yb_stream_t stream) {
  YB_REQUIRE(x && y && batch >= 1 && h >= 1 && w >= 1, "yolo_input_patchify: bad arguments");
  YB_REQUIRE(c >= 1 && 9 * c <= 32, "yolo_input_patchify: needs 9*c <= 32 (got c=%d)", c);
  const long long total = (long long)batch * h * w;
  __nv_bfloat16* yo = static_cast<__nv_bfloat16*>(y);
  const unsigned g = blocks_for(total, 256);
  if (c == 3) k_patchify<3><<<g, 256, 0, (cudaStream_t)stream>>>(x, batch, h, w, yo, status);
  else if (c == 2) k_patchify<2><<<g, 256, 0, (cudaStream_t)stream>>>(x, batch, h, w, yo, status);
  else k_patchify<1><<<g, 256, 0, (cudaStream_t)stream>>>(x, batch, h, w, yo, status);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Inference pre-processing: what code/config.py:101-113 `set_only_image_transforms` does on
// the CPU through albumentations + OpenCV (demo.py:37-39), for a batch of uint8 HWC images of different sizes:
//   LongestMaxSize(S) [cv2.resize INTER_LINEAR, uint8 fixed point] -> PadIfNeeded(S, S, constant 0, centred)
//   -> Normalize(mean 0, std 1, max_pixel 255) -> ToTensorV2 (HWC -> CHW), written as fp32 (B, 3, S, S).
// One thread per output pixel recomputes OpenCV's coordinate / 11-bit weight arithmetic (resize.cpp, generic
// linear path; see oracle/preprocess_oracle.py for the restatement that is pinned bit-for-bit against cv2):
// horizontally a coordinate beyond the border snaps to the border pixel, vertically only the row index is
// clipped; vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
namespace {

struct LetterboxImage {
  const uint8_t* data;  // HWC uint8, dense
  int h, w;             // source size
  int nh, nw;           // resized size (LongestMaxSize), computed on the host with Python's round-half-even
  int top, left;        // PadIfNeeded offsets
};

__device__ __forceinline__ void lb_axis(int d, int src, int dst, bool vertical, int& i0, int& i1, int& w0, int& w1) {
  const double inv_scale = double(dst) / double(src);
  const double scale = 1.0 / inv_scale;
  float f = float((double(d) + 0.5) * scale - 0.5);
  int s = int(floorf(f));
  f -= float(s);
  if (!vertical) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  i0 = min(max(s, 0), src - 1);
  i1 = min(max(s + 1, 0), src - 1);
}

__global__ void __launch_bounds__(256) k_letterbox(const LetterboxImage* __restrict__ imgs, int size, int channels,
                                                   float* __restrict__ out) {
  const LetterboxImage im = imgs[blockIdx.z];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= size || y >= size) return;
  float* o = out + (size_t(blockIdx.z) * channels * size + y) * size + x;
  const size_t plane = size_t(size) * size;
  const int rx = x - im.left, ry = y - im.top;
  if (rx < 0 || rx >= im.nw || ry < 0 || ry >= im.nh) {
    for (int c = 0; c < channels; ++c) o[c * plane] = 0.f;
    return;
  }
  const float inv255 = 1.0f / 255.0f;
  if (im.nh == im.h && im.nw == im.w) {  // LongestMaxSize leaves the image alone when scale == 1
    const uint8_t* p = im.data + (size_t(ry) * im.w + rx) * channels;
    for (int c = 0; c < channels; ++c) o[c * plane] = __fmul_rn(float(p[c]), inv255);
    return;
  }
  int x0, x1, a0, a1, y0, y1, b0, b1;
  lb_axis(rx, im.w, im.nw, false, x0, x1, a0, a1);
  lb_axis(ry, im.h, im.nh, true, y0, y1, b0, b1);
  const uint8_t* r0 = im.data + size_t(y0) * im.w * channels;
  const uint8_t* r1 = im.data + size_t(y1) * im.w * channels;
  for (int c = 0; c < channels; ++c) {
    const int s0 = int(r0[x0 * channels + c]) * a0 + int(r0[x1 * channels + c]) * a1;  // horizontal pass, 2^11 scale
    const int s1 = int(r1[x0 * channels + c]) * a0 + int(r1[x1 * channels + c]) * a1;
    int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
    o[c * plane] = __fmul_rn(float(v), inv255);
  }
}

}  // namespace

extern "C" size_t yolo_letterbox_desc_bytes(void) { return sizeof(LetterboxImage); }

// descs_dev: device array of `batch` descriptors laid out as {const uint8_t* data; int32 h, w, nh, nw, top, left}
// (yolo_letterbox_desc_bytes() bytes each).
extern "C" int yolo_letterbox_u8(const void* descs_dev, int batch, int size, int channels, float* out, yb_stream_t stream) {
  YB_REQUIRE(descs_dev && out && batch >= 0 && size >= 1 && channels >= 1 && channels <= 4, "yolo_letterbox_u8: bad argument");
  if (batch == 0) return YB_OK;
  dim3 grid((size + 31) / 32, (size + 7) / 8, batch);
  k_letterbox<<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const LetterboxImage*>(descs_dev), size, channels, out);
  YB_CHECK_LAUNCH();
  return YB_OK;
}
