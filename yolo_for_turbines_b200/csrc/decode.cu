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

// Fast path for the model's own head layout: a dense [B*S*S pixels][pitch] fp32 buffer whose (B,3,S,S,C)
// view has strides (S*S*pitch, C, S*pitch, pitch, 1).  A CTA stages DEC_PIX whole pixel rows in shared
// memory with coalesced loads (row pitch + 1 words => the per-cell scans below are bank-conflict free),
// then one thread per (pixel, anchor) cell scans its logits from shared memory.  Same arithmetic as
// k_decode; HBM traffic is exactly one read of the head and one 24-byte row per cell.
constexpr int DEC_PIX = 32;
constexpr int DEC_THREADS = 128;

__device__ __forceinline__ void decode_dense_block(const DecodeParams& p, int pitch, long long total_pix, long long block) {
  extern __shared__ float s_head[];
  const int sp = pitch + 1;
  const long long pix0 = block * DEC_PIX;
  const int npix = (int)min((long long)DEC_PIX, total_pix - pix0);
  const float* src = p.head + pix0 * pitch;
  if ((pitch & 3) == 0) {  // 16-byte loads, 8 in flight per thread before the (scalar, padded-row) smem stores
    const int nvec = npix * pitch / 4;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    for (int base = threadIdx.x; base < nvec; base += DEC_THREADS * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * DEC_THREADS;
        v[u] = i < nvec ? __ldg(src4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * DEC_THREADS;
        if (i < nvec) {
          const int e = i * 4, r = e / pitch, c = e - r * pitch;
          float* dst = s_head + r * sp + c;
          dst[0] = v[u].x; dst[1] = v[u].y; dst[2] = v[u].z; dst[3] = v[u].w;
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < npix * pitch; i += DEC_THREADS) {
      const int r = i / pitch, c = i - r * pitch;
      s_head[r * sp + c] = src[i];
    }
  }
  __syncthreads();
  const int C = 5 + p.nc;
  const int ss = p.S * p.S;
  for (int cell = threadIdx.x; cell < 3 * npix; cell += DEC_THREADS) {
    const int a = cell / npix, r = cell - a * npix;  // consecutive threads -> consecutive output rows
    const float* t = s_head + r * sp + a * C;
    float best = t[5];
    int best_i = 0;
    bool best_nan = best != best;
    for (int c = 1; c < p.nc; ++c) {  // utils.py:112: first maximal logit, NaN counts as maximal
      const float v = t[5 + c];
      const bool vn = v != v;
      if (!best_nan && (vn || v > best)) { best = v; best_i = c; best_nan = vn; }
    }
    const long long pix = pix0 + r;
    const int b = (int)(pix / ss);
    const int rem = (int)(pix - (long long)b * ss);
    const int i = rem / p.S, j = rem - i * p.S;
    float* o = p.out + (size_t(b) * p.out_boxes_per_image + p.out_offset + size_t(a) * ss + rem) * 6;
    const float sx = sigmoid_f32(t[0]), sy = sigmoid_f32(t[1]);
    const float w = __fmul_rn(expf(t[2]), p.anchors[2 * a]), h = __fmul_rn(expf(t[3]), p.anchors[2 * a + 1]);
    float2* o2 = reinterpret_cast<float2*>(o);
    o2[0] = make_float2(__fmul_rn(p.inv_s, __fadd_rn(sx, float(j))), __fmul_rn(p.inv_s, __fadd_rn(sy, float(i))));
    o2[1] = make_float2(__fmul_rn(p.inv_s, w), __fmul_rn(p.inv_s, h));
    o2[2] = make_float2(sigmoid_f32(t[4]), float(best_i));
  }
}

__global__ void __launch_bounds__(DEC_THREADS) k_decode_dense(const DecodeParams p, int pitch, long long total_pix) {
  decode_dense_block(p, pitch, total_pix, blockIdx.x);
}

// All scales of a detector in ONE launch: block ranges [first[i], first[i+1]) belong to scale i.  The 13x13 and 26x26
// heads are launch-latency bound on their own (11.7 / 22.4 us for 1.4 / 5.5 MB of logits at batch 64); their blocks now
// run beside the 52x52 scale's.
constexpr int DEC_MAX_SCALES = 4;
struct DecodeMulti {
  DecodeParams p[DEC_MAX_SCALES];
  long long total_pix[DEC_MAX_SCALES];
  unsigned first[DEC_MAX_SCALES + 1];
  int pitch[DEC_MAX_SCALES];
  int n;
};
__global__ void __launch_bounds__(DEC_THREADS) k_decode_dense_multi(const __grid_constant__ DecodeMulti m) {
  int i = 0;
  while (i + 1 < m.n && blockIdx.x >= m.first[i + 1]) ++i;   // uniform per block
  decode_dense_block(m.p[i], m.pitch[i], m.total_pix[i], (long long)(blockIdx.x - m.first[i]));
}

bool dense_layout(const DecodeParams& p, int is_pred, int writeback) {
  const long long pitch = p.st[3];
  const int C = 5 + p.nc;
  return is_pred && !writeback && p.st[4] == 1 && p.st[1] == C && pitch >= 3 * C && pitch <= 384 &&
         p.st[2] == (long long)p.S * pitch && p.st[0] == (long long)p.S * p.S * pitch &&
         (reinterpret_cast<uintptr_t>(p.out) & 7) == 0;
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
  const long long pitch = p.st[3];
  if (dense_layout(p, is_pred, writeback)) {
    const long long total_pix = (long long)batch * S * S;
    const size_t smem = size_t(DEC_PIX) * (pitch + 1) * sizeof(float);
    k_decode_dense<<<(unsigned)((total_pix + DEC_PIX - 1) / DEC_PIX), DEC_THREADS, smem, (cudaStream_t)stream>>>(
        p, (int)pitch, total_pix);
    YB_CHECK_LAUNCH();
    return YB_OK;
  }
  const long long cells = 3ll * S * S * batch;
  const int wpb = 8;
  k_decode<<<(unsigned)((cells + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(p);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_decode_multi(const float* const* heads_host, const int64_t* strides5_host, int batch, const int32_t* S_host,
                                 int nc, const float* anchors6_host, int num_scales, float* out, int out_boxes_per_image,
                                 yb_stream_t stream) {
  YB_REQUIRE(heads_host && strides5_host && S_host && anchors6_host && out, "yolo_decode_multi: null pointer");
  YB_REQUIRE(num_scales >= 1 && num_scales <= DEC_MAX_SCALES && batch >= 0 && nc >= 1, "yolo_decode_multi: bad shape");
  if (batch == 0) return YB_OK;
  DecodeMulti m;
  m.n = num_scales;
  int off = 0;
  unsigned blocks = 0;
  int max_pitch = 0;
  for (int i = 0; i < num_scales; ++i) {
    DecodeParams& p = m.p[i];
    YB_REQUIRE(heads_host[i] && S_host[i] >= 1, "yolo_decode_multi: bad scale %d", i);
    p.head = heads_host[i];
    for (int k = 0; k < 5; ++k) p.st[k] = strides5_host[5 * i + k];
    p.batch = batch; p.S = S_host[i]; p.nc = nc;
    for (int k = 0; k < 6; ++k) p.anchors[k] = anchors6_host[6 * i + k];
    p.is_pred = 1; p.writeback = 0;
    p.out = out; p.out_boxes_per_image = out_boxes_per_image; p.out_offset = off;
    p.inv_s = (float)(1.0 / (double)p.S);
    if (!dense_layout(p, 1, 0)) {
      yb_set_error("yolo_decode_multi: scale %d is not a dense [pixels][pitch] head (use yolo_decode)", i);
      return YB_ERR_UNSUPPORTED;
    }
    m.pitch[i] = (int)p.st[3];
    m.total_pix[i] = (long long)batch * p.S * p.S;
    m.first[i] = blocks;
    blocks += (unsigned)((m.total_pix[i] + DEC_PIX - 1) / DEC_PIX);
    if (m.pitch[i] > max_pitch) max_pitch = m.pitch[i];
    off += 3 * p.S * p.S;
  }
  m.first[num_scales] = blocks;
  YB_REQUIRE(out_boxes_per_image >= off, "yolo_decode_multi: output rows do not fit");
  const size_t smem = size_t(DEC_PIX) * (max_pitch + 1) * sizeof(float);
  k_decode_dense_multi<<<blocks, DEC_THREADS, smem, (cudaStream_t)stream>>>(m);
  YB_CHECK_LAUNCH();
  return YB_OK;
}
