// Training-step kernels around the conv GEMMs (SURVEY config #4; reference: code/train.py:53-69 with
// model.train()).  Everything here is HBM-bound NHWC bf16 row work:
//
//   forward   CNNBlock in train mode (model.py:80-86 with nn.BatchNorm2d batch statistics):
//             z = conv(x) [tcgen05 kernel, bf16]  ->  k_bn_stats (sum z, sum z^2 per channel)
//             -> k_bn_finalize (mean / rstd, running-stat update, scale/bias)  -> k_bn_act_fwd
//             a = act(z*scale + bias) (+ residual) (optionally stored 2x2-replicated = nn.Upsample)
//   backward  k_bn_act_bwd_reduce (sum dy, sum dy*xhat)  -> k_bn_bwd_finalize (dgamma, dbeta, means)
//             -> k_bn_act_bwd_apply  dz = scale * (dy - mean(dy) - xhat * mean(dy*xhat))
//             (optionally also zero-stuffed to 2x resolution: the stride-2 convs' dgrad input)
//   k_unpack_wgrad   packed fp32 dW [Cout][tap][Cin] -> nn.Conv2d.weight.grad (OIHW)
//   k_sgd            torch.optim.SGD(momentum, weight_decay) on the flat parameter buffer (train.py:171-172)
//
// Rows are handled 8 channels (one 16-byte load) per thread; per-channel sums go through shared-memory
// double atomics and one global double atomic per channel per block.
#include <stdlib.h>

#include "common.cuh"
#include "conv_ptx.cuh"

namespace {

using convptx::bf16_hi;
using convptx::bf16_lo;
using convptx::pack_bf16;

constexpr int TR_THREADS = 256;

// (Programmatic dependent launch was tried on the row kernels between two convs and measured SLOWER: backward pass
// 8.67 -> 9.75 ms (LeakyReLU, batch 32, 416^2).  The early-started data-gradient conv CTAs hold their SMs' shared memory
// while they wait, which keeps the weight-gradient kernels of the side stream off those SMs.)
__device__ __forceinline__ void unpack8(const uint4 u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ uint4 ld8(const __nv_bfloat16* base, size_t row, int pitch, int c) {
  return __ldg(reinterpret_cast<const uint4*>(base + row * size_t(pitch) + c));
}

// Forward Mish in 7 operations with one exp and one reciprocal: with n = e^y and a = n (n + 2) + 2,
// y tanh(softplus(y)) = y (a - 2) / a = y - 2 y / a.  No clamp is needed: a overflows to +inf for y > 44, 1 / a = 0 and
// the result is y (exact to fp32 there); for very negative y, a -> 2 and the result cancels to 0 with an absolute
// error below |y| 2^-23 (the true value is y e^y).  The pass is instruction-issue bound (see EW_ITEMS below).
__device__ __forceinline__ float mish_fwd(float y) { return convptx::mish_fast(y); }   // 7 SASS operations, conv_ptx.cuh
__device__ __forceinline__ float act_fwd(float y, int act) {
  if (act == YB_ACT_LEAKY) return fmaxf(y, 0.1f * y);
  if (act == YB_ACT_MISH) return mish_fwd(y);
  return y;
}
// d act(y) / dy.  Mish in closed form with one exp and one reciprocal: with n = e^y and a = n (n + 2) + 2,
//   d/dy [y tanh(softplus(y))] = n w / a^2,   w = 4 (y + 1) + n (4 y + 6 + n (4 + n))
// (14 operations; the tanh / sigmoid form needed 19, and the kernel that calls this is instruction-issue bound).
// y is clamped at 20, where the derivative is 1 to fp32 precision and n^3, n w and 1 / a^2 are all still in range.
__device__ __forceinline__ float act_grad(float y, int act) {
  if (act == YB_ACT_LEAKY) return y > 0.f ? 1.f : 0.1f;
  if (act == YB_ACT_MISH) {
    y = fminf(y, 20.f);
    const float n = convptx::ex2_ftz(y * 1.4426950408889634f);   // bare MUFU forms: see mish_fast (conv_ptx.cuh)
    const float a = fmaf(n, n + 2.f, 2.f);
    const float w = fmaf(n, fmaf(n, n + 4.f, fmaf(4.f, y, 6.f)), fmaf(4.f, y, 4.f));
    const float r = convptx::rcp_ftz(a);
    return (n * w) * (r * r);
  }
  return 1.f;
}
__device__ __forceinline__ void ld8f(const float* __restrict__ p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Block-level finish of the per-channel reductions: every thread holds 8 + 8 float partial sums for the 8
// channels of its group; they are transposed through shared memory (row stride 257 floats: conflict free),
// summed over the block's row lanes, and each block issues ONE double atomic per output.
constexpr int RED_STRIDE = 257;
__device__ __forceinline__ void block_channel_reduce(const float (&s)[8], const float (&q)[8], int groups, int lanes, bool active,
                                                     int C, float* red /* [16][RED_STRIDE] */, double* __restrict__ sums) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[k * RED_STRIDE + threadIdx.x] = active ? s[k] : 0.f;
    red[(8 + k) * RED_STRIDE + threadIdx.x] = active ? q[k] : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 2 * C; o += TR_THREADS) {
    const int c = o >> 1, which = o & 1;
    const int grp = c >> 3, k = c & 7;
    const float* row = red + (which * 8 + k) * RED_STRIDE + grp;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += row[l * groups];
    atomicAdd(&sums[o], double(acc));
  }
}

struct RowGeom {
  long long P;       // rows (pixels) of the layer output
  int C, groups;     // channels, C / 8
  int h, w;          // output height / width (only for the 2x variants)
};

// ---------------------------------------------------------------------------------------------------------
// per-channel sums over rows: sums[2c] += sum z, sums[2c+1] += sum z^2
__global__ void __launch_bounds__(TR_THREADS) k_bn_stats(const __nv_bfloat16* __restrict__ z, int pitch, RowGeom g,
                                                         double* __restrict__ sums) {
  __shared__ float red[16 * RED_STRIDE];
  const int lanes = TR_THREADS / g.groups;
  const int grp = threadIdx.x % g.groups, lane = threadIdx.x / g.groups;
  const long long per_block = (g.P + gridDim.x - 1) / gridDim.x;
  const long long r0 = per_block * blockIdx.x;
  const long long r1 = r0 + per_block < g.P ? r0 + per_block : g.P;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool active = lane < lanes;
  if (active) {
    long long r = r0 + lane;
    for (; r + 3ll * lanes < r1; r += 4ll * lanes) {  // four independent 16-byte loads in flight per thread
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = ld8(z, size_t(r + (long long)j * lanes), pitch, grp * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
        unpack8(u[j], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] = fmaf(f[k], f[k], q[k]); }
      }
    }
    for (; r < r1; r += lanes) {
      float f[8];
      unpack8(ld8(z, size_t(r), pitch, grp * 8), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] = fmaf(f[k], f[k], q[k]); }
    }
  }
  block_channel_reduce(s, q, g.groups, lanes, active, g.C, red, sums);
}

// nn.BatchNorm2d training semantics: normalise with the biased batch variance, update the running
// statistics with momentum and the UNBIASED variance.
__global__ void k_bn_finalize(const double* __restrict__ sums, int C, const BnFinalize f) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) bn_finalize_channel(f, c, sums[2 * c], sums[2 * c + 1]);
}

// The block that takes the last ticket sees every other block's (L2-resident) atomics: it finishes the layer.
__device__ __forceinline__ bool last_block_done(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *counter = 0u;  // ready for the next launch
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// k_bn_stats + k_bn_finalize in one launch
__global__ void __launch_bounds__(TR_THREADS) k_bn_stats_fused(const __nv_bfloat16* __restrict__ z, int pitch, RowGeom g,
                                                               double* __restrict__ sums, unsigned int* counter,
                                                               const BnFinalize f) {
  __shared__ float red[16 * RED_STRIDE];
  const int lanes = TR_THREADS / g.groups;
  const int grp = threadIdx.x % g.groups, lane = threadIdx.x / g.groups;
  const long long per_block = (g.P + gridDim.x - 1) / gridDim.x;
  const long long r0 = per_block * blockIdx.x;
  const long long r1 = r0 + per_block < g.P ? r0 + per_block : g.P;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool active = lane < lanes;
  if (active) {
    long long r = r0 + lane;
    for (; r + 3ll * lanes < r1; r += 4ll * lanes) {
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = ld8(z, size_t(r + (long long)j * lanes), pitch, grp * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[8];
        unpack8(u[j], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
      }
    }
    for (; r < r1; r += lanes) {
      float v[8];
      unpack8(ld8(z, size_t(r), pitch, grp * 8), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
    }
  }
  block_channel_reduce(s, q, g.groups, lanes, active, g.C, red, sums);
  if (last_block_done(counter)) {
    for (int c = threadIdx.x; c < g.C; c += TR_THREADS) bn_finalize_channel(f, c, __ldcg(sums + 2 * c), __ldcg(sums + 2 * c + 1));
  }
}

// Streaming row kernels (k_bn_act_fwd, k_bn_act_bwd_apply): every thread owns EW_ITEMS (row, 8-channel group) items
// and issues ALL of their 16-byte loads before the first use.  With one item per thread both kernels ran at ~2.8 TB/s
// whether the tensors came from L2 or from DRAM (time proportional to the grid size: profiles/
// r2_train_launches_warm_before.csv).  Measured effect over the 72 layers of a step (same list, _after): the pure
// multiply-add pass 2 of the backward 1.54 -> 1.15 ms; the forward pass with Mish 1.69 -> 1.65 ms, i.e. that one is
// bound by instruction issue (exp + reciprocal + ~16 ALU ops per element), not by loads in flight.
// Item j of thread t is t + j * nq with nq a multiple of `groups`, so the items of a thread share their channel group
// (scale / bias are fetched once) and a warp's accesses stay contiguous.
constexpr int EW_ITEMS = 4;
__host__ __device__ inline unsigned ew_quarter(unsigned n_items, unsigned groups) {
  const unsigned q = (n_items + EW_ITEMS - 1) / EW_ITEMS;
  return (q + groups - 1) / groups * groups;
}
__device__ __forceinline__ void st8_stream(__nv_bfloat16* p, const uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

// a = act(z * scale + bias) (+ residual)
__global__ void __launch_bounds__(TR_THREADS) k_bn_act_fwd(const __nv_bfloat16* __restrict__ z, int z_pitch, RowGeom g,
                                                           const float* __restrict__ scale, const float* __restrict__ bias,
                                                           int act, const __nv_bfloat16* __restrict__ res, int res_pitch,
                                                           __nv_bfloat16* __restrict__ y, int y_pitch, unsigned nq) {
  // 32-bit index split (the host wrapper guarantees P * groups < 2^31): a 64-bit division per thread costs more
  // instructions than the whole activation
  const unsigned t = blockIdx.x * unsigned(TR_THREADS) + threadIdx.x;
  if (t >= nq) return;
  const unsigned groups = unsigned(g.groups), rq = nq / groups;
  const unsigned r0 = t / groups;
  const int c = int(t - r0 * groups) * 8;
  uint4 zu[EW_ITEMS], ru[EW_ITEMS];
  bool ok[EW_ITEMS];
#pragma unroll
  for (int j = 0; j < EW_ITEMS; ++j) {
    const unsigned r = r0 + unsigned(j) * rq;
    ok[j] = r < unsigned(g.P);
    if (ok[j]) zu[j] = ld8(z, size_t(r), z_pitch, c);
  }
  if (res) {
#pragma unroll
    for (int j = 0; j < EW_ITEMS; ++j)
      if (ok[j]) ru[j] = ld8(res, size_t(r0 + unsigned(j) * rq), res_pitch, c);
  }
  float sc[8], bi[8];
  ld8f(scale + c, sc); ld8f(bias + c, bi);
#pragma unroll
  for (int j = 0; j < EW_ITEMS; ++j) {
    if (!ok[j]) continue;
    float f[8];
    unpack8(zu[j], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = act_fwd(fmaf(f[k], sc[k], bi[k]), act);
    if (res) {
      float rr[8];
      unpack8(ru[j], rr);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += rr[k];
    }
    st8_stream(y + size_t(r0 + unsigned(j) * rq) * y_pitch + c, pack8(f));
  }
}

// the same with every pixel stored to its 2x2 block of a (2h, 2w) tensor (= nn.Upsample(2) of the activation); two layers
__global__ void __launch_bounds__(TR_THREADS) k_bn_act_fwd_up2x(const __nv_bfloat16* __restrict__ z, int z_pitch, RowGeom g,
                                                                const float* __restrict__ scale, const float* __restrict__ bias,
                                                                int act, const __nv_bfloat16* __restrict__ res, int res_pitch,
                                                                __nv_bfloat16* __restrict__ y, int y_pitch) {
  const unsigned idx = blockIdx.x * unsigned(TR_THREADS) + threadIdx.x;
  if (idx >= unsigned(g.P) * unsigned(g.groups)) return;
  const unsigned r = idx / unsigned(g.groups);
  const int c = int(idx - r * unsigned(g.groups)) * 8;
  float f[8];
  unpack8(ld8(z, size_t(r), z_pitch, c), f);
  float sc[8], bi[8];
  ld8f(scale + c, sc); ld8f(bias + c, bi);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = act_fwd(fmaf(f[k], sc[k], bi[k]), act);
  if (res) {
    float rr[8];
    unpack8(ld8(res, size_t(r), res_pitch, c), rr);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] += rr[k];
  }
  const uint4 o = pack8(f);
  const long long hw = (long long)g.h * g.w;
  const long long img = r / hw;
  const int rem = int(r - img * hw);
  const int i = rem / g.w, j = rem - i * g.w;
  const size_t W2 = size_t(2 * g.w);
  const size_t r00 = (size_t(img) * (2 * g.h) + 2 * i) * W2 + 2 * j;
  *reinterpret_cast<uint4*>(y + r00 * y_pitch + c) = o;
  *reinterpret_cast<uint4*>(y + (r00 + 1) * y_pitch + c) = o;
  *reinterpret_cast<uint4*>(y + (r00 + W2) * y_pitch + c) = o;
  *reinterpret_cast<uint4*>(y + (r00 + W2 + 1) * y_pitch + c) = o;
}

// gradient arriving at this layer's output, 8 channels of row r; up2x = backward of nn.Upsample(2): the sum of
// the 2x2 block of the (2h, 2w) gradient tensor
__device__ __forceinline__ void load_dA(const __nv_bfloat16* __restrict__ dA, int pitch, const RowGeom& g, long long r, int c,
                                        int up2x, float (&d)[8]) {
  if (!up2x) {
    unpack8(ld8(dA, size_t(r), pitch, c), d);
    return;
  }
  const long long hw = (long long)g.h * g.w;
  const long long img = r / hw;
  const int rem = int(r - img * hw);
  const int i = rem / g.w, j = rem - i * g.w;
  const size_t W2 = size_t(2 * g.w);
  const size_t r00 = (size_t(img) * (2 * g.h) + 2 * i) * W2 + 2 * j;
  float t[8];
  unpack8(ld8(dA, r00, pitch, c), d);
  unpack8(ld8(dA, r00 + 1, pitch, c), t);
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k] += t[k];
  unpack8(ld8(dA, r00 + W2, pitch, c), t);
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k] += t[k];
  unpack8(ld8(dA, r00 + W2 + 1, pitch, c), t);
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k] += t[k];
}

struct BnBwdParams {
  const __nv_bfloat16* dA; int dA_pitch, up2x;
  const __nv_bfloat16* z; int z_pitch;
  const float *scale, *bias, *mean, *rstd;
  int act;
  RowGeom g;
};

// Pass 1: sums[2c] += sum dy, sums[2c+1] += sum dy * z with dy = dA * act'(z*scale+bias); the last block turns them
// into dbeta = sum dy, dgamma = sum dy*xhat = rstd * (sum dy*z - mean * sum dy) and the two coefficient vectors of
// pass 2:  dz = scale*dy + c1*z + c0,  c1 = -scale*rstd*dgamma/P,  c0 = -scale*dbeta/P - c1*mean
// (= scale * (dy - mean(dy) - xhat * mean(dy*xhat)), the BatchNorm backward).
struct BnBwdFinalize {
  long long P;
  float *dgamma, *dbeta, *c1, *c0;
};
__device__ __forceinline__ void bn_bwd_finalize_channel(const BnBwdParams& p, const BnBwdFinalize& f, int c, double s, double q) {
  const double mu = double(p.mean[c]), rs = double(p.rstd[c]), sc = double(p.scale[c]);
  const double dgam = rs * (q - mu * s);
  f.dbeta[c] = float(s);
  f.dgamma[c] = float(dgam);
  const double c1 = -sc * rs * dgam / double(f.P);
  f.c1[c] = float(c1);
  f.c0[c] = float(-sc * s / double(f.P) - c1 * mu);
}

__global__ void __launch_bounds__(TR_THREADS, 3) k_bn_act_bwd_reduce(const BnBwdParams p, double* __restrict__ sums,
                                                                     unsigned int* counter, const BnBwdFinalize f,
                                                                     __nv_bfloat16* __restrict__ dy_out, int dy_pitch) {
  __shared__ float red[16 * RED_STRIDE];
  const RowGeom& g = p.g;
  const int lanes = TR_THREADS / g.groups;
  const int grp = threadIdx.x % g.groups, lane = threadIdx.x / g.groups;
  const long long per_block = (g.P + gridDim.x - 1) / gridDim.x;
  const long long r0 = per_block * blockIdx.x;
  const long long r1 = r0 + per_block < g.P ? r0 + per_block : g.P;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool active = lane < lanes;
  if (active) {
    const int c = grp * 8;
    float sc[8], bi[8];
    ld8f(p.scale + c, sc); ld8f(p.bias + c, bi);
    // dy = dA * act'(y) is also WRITTEN (bf16, into the dz buffer): pass 2 then needs neither dA nor the
    // activation derivative again -- these kernels are issue-bound (ncu), not DRAM-bound
    auto accum = [&](const float (&zf)[8], float (&d)[8], long long row) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float dy = d[k] * act_grad(fmaf(zf[k], sc[k], bi[k]), p.act);
        d[k] = dy;
        s[k] += dy;
        q[k] = fmaf(dy, zf[k], q[k]);
      }
      *reinterpret_cast<uint4*>(dy_out + size_t(row) * dy_pitch + c) = pack8(d);
    };
    long long r = r0 + lane;
    if (!p.up2x) {
      for (; r + lanes < r1; r += 2ll * lanes) {  // two rows (four 16-byte loads) in flight per thread; four rows measured
        // slower (2.82 vs 2.64 ms over the 72 layers: the Mish derivative makes this pass issue-bound, and the deeper
        // unroll costs a resident block per SM)
        const uint4 z0 = ld8(p.z, size_t(r), p.z_pitch, c), z1 = ld8(p.z, size_t(r + lanes), p.z_pitch, c);
        const uint4 d0 = ld8(p.dA, size_t(r), p.dA_pitch, c), d1 = ld8(p.dA, size_t(r + lanes), p.dA_pitch, c);
        float zf[8], d[8];
        unpack8(z0, zf); unpack8(d0, d); accum(zf, d, r);
        unpack8(z1, zf); unpack8(d1, d); accum(zf, d, r + lanes);
      }
    }
    for (; r < r1; r += lanes) {
      float zf[8], d[8];
      unpack8(ld8(p.z, size_t(r), p.z_pitch, c), zf);
      load_dA(p.dA, p.dA_pitch, g, r, c, p.up2x, d);
      accum(zf, d, r);
    }
  }
  block_channel_reduce(s, q, g.groups, lanes, active, g.C, red, sums);
  if (last_block_done(counter)) {
    for (int c = threadIdx.x; c < g.C; c += TR_THREADS)
      bn_bwd_finalize_channel(p, f, c, __ldcg(sums + 2 * c), __ldcg(sums + 2 * c + 1));
  }
}

// bias-only conv: sums[2c] is the bias gradient
__global__ void k_bias_finalize(const double* __restrict__ sums, int C, float* __restrict__ dbias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) dbias[c] = float(sums[2 * c]);
}

// Pass 2: dz = scale*dy + c1*z + c0, in place over the dy that pass 1 left in the dz buffer (EW_ITEMS items per thread,
// all loads first: see k_bn_act_fwd)
__global__ void __launch_bounds__(TR_THREADS) k_bn_act_bwd_apply(const BnBwdParams p, const float* __restrict__ c1,
                                                                 const float* __restrict__ c0, __nv_bfloat16* __restrict__ dz,
                                                                 int dz_pitch, unsigned nq) {
  const RowGeom& g = p.g;
  const unsigned t = blockIdx.x * unsigned(TR_THREADS) + threadIdx.x;
  if (t >= nq) return;
  const unsigned groups = unsigned(g.groups), rq = nq / groups;
  const unsigned r0 = t / groups;
  const int c = int(t - r0 * groups) * 8;
  uint4 zu[EW_ITEMS], du[EW_ITEMS];
  bool ok[EW_ITEMS];
#pragma unroll
  for (int j = 0; j < EW_ITEMS; ++j) {
    const unsigned r = r0 + unsigned(j) * rq;
    ok[j] = r < unsigned(g.P);
    if (ok[j]) {
      zu[j] = ld8(p.z, size_t(r), p.z_pitch, c);
      du[j] = *reinterpret_cast<const uint4*>(dz + size_t(r) * dz_pitch + c);   // dy, left here by pass 1
    }
  }
  float sc[8], a1[8], a0[8];
  ld8f(p.scale + c, sc); ld8f(c1 + c, a1); ld8f(c0 + c, a0);
#pragma unroll
  for (int j = 0; j < EW_ITEMS; ++j) {
    if (!ok[j]) continue;
    float zf[8], d[8], o[8];
    unpack8(zu[j], zf); unpack8(du[j], d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaf(sc[k], d[k], fmaf(a1[k], zf[k], a0[k]));
    *reinterpret_cast<uint4*>(dz + size_t(r0 + unsigned(j) * rq) * dz_pitch + c) = pack8(o);
  }
}

// the same plus a zero-stuffed copy at 2x resolution (value at (2i, 2j)): the stride-2 layers' data-gradient operand
__global__ void __launch_bounds__(TR_THREADS, 4) k_bn_act_bwd_apply_stuffed(const BnBwdParams p, const float* __restrict__ c1,
                                                                            const float* __restrict__ c0, __nv_bfloat16* __restrict__ dz,
                                                                            int dz_pitch, __nv_bfloat16* __restrict__ stuffed,
                                                                            int stuffed_pitch) {
  const RowGeom& g = p.g;
  const unsigned idx = blockIdx.x * unsigned(TR_THREADS) + threadIdx.x;   // 32-bit split, see k_bn_act_fwd
  if (idx >= unsigned(g.P) * unsigned(g.groups)) return;
  const unsigned r = idx / unsigned(g.groups);
  const int c = int(idx - r * unsigned(g.groups)) * 8;
  float zf[8], d[8], o[8];
  unpack8(ld8(p.z, size_t(r), p.z_pitch, c), zf);
  unpack8(*reinterpret_cast<const uint4*>(dz + size_t(r) * dz_pitch + c), d);   // dy, left here by pass 1
  {
    float sc[8], a1[8], a0[8];
    ld8f(p.scale + c, sc); ld8f(c1 + c, a1); ld8f(c0 + c, a0);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaf(sc[k], d[k], fmaf(a1[k], zf[k], a0[k]));
  }
  const uint4 u = pack8(o);
  *reinterpret_cast<uint4*>(dz + size_t(r) * dz_pitch + c) = u;
  const long long hw = (long long)g.h * g.w;
  const long long img = r / hw;
  const int rem = int(r - img * hw);
  const int i = rem / g.w, j = rem - i * g.w;
  const size_t W2 = size_t(2 * g.w);
  const size_t r00 = (size_t(img) * (2 * g.h) + 2 * i) * W2 + 2 * j;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  *reinterpret_cast<uint4*>(stuffed + r00 * stuffed_pitch + c) = u;
  *reinterpret_cast<uint4*>(stuffed + (r00 + 1) * stuffed_pitch + c) = zero;
  *reinterpret_cast<uint4*>(stuffed + (r00 + W2) * stuffed_pitch + c) = zero;
  *reinterpret_cast<uint4*>(stuffed + (r00 + W2 + 1) * stuffed_pitch + c) = zero;
}

// forward + data-gradient operand packs of one layer in one pass over the fp32 weights:
//   fwd[co][tap][ci] = w[co][ci][tap]            ([c_out_pad][taps][c_in_pad], padding pre-zeroed by the caller)
//   bwd[ci][taps-1-tap][co] = w[co][ci][tap]     ([c_in_pad][taps][c_out_pad])
// One block per 32 (co) x 32 (ci) tile staged through shared memory: OIHW rows are read as contiguous runs of
// 32*taps floats, both packs are written as 64-byte runs (ci-contiguous / co-contiguous).
constexpr int PK_T = 32;
template <int TAPS>
__global__ void __launch_bounds__(256) k_pack_weights_both(const float* __restrict__ w, int c_out, int c_in, int /*taps*/,
                                                          int c_in_pad, int c_out_pad, __nv_bfloat16* __restrict__ fwd,
                                                          __nv_bfloat16* __restrict__ bwd) {
  extern __shared__ float s_tile[];  // [PK_T][PK_T * taps + 1]
  constexpr int taps = TAPS;         // compile-time: every index split below is a multiply-shift, not a division
  constexpr int L = PK_T * taps, pitch = L + 1;
  const int ci0 = blockIdx.x * PK_T, co0 = blockIdx.y * PK_T;
  const int nci = min(PK_T, c_in - ci0);
  for (int i = threadIdx.x; i < PK_T * L; i += 256) {
    const int r = i / L, c = i - r * L;
    float v = 0.f;
    if (co0 + r < c_out && c < nci * taps) v = w[(size_t(co0 + r) * c_in + ci0) * taps + c];
    s_tile[r * pitch + c] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PK_T * L; i += 256) {   // forward pack: ci fastest
    const int cil = i % PK_T, t = (i / PK_T) % taps, r = i / (PK_T * taps);
    if (co0 + r < c_out && cil < nci)
      fwd[(size_t(co0 + r) * taps + t) * c_in_pad + ci0 + cil] = __float2bfloat16_rn(s_tile[r * pitch + cil * taps + t]);
  }
  for (int i = threadIdx.x; i < PK_T * L; i += 256) {   // data-gradient pack: co fastest, taps flipped
    const int r = i % PK_T, t = (i / PK_T) % taps, cil = i / (PK_T * taps);
    if (co0 + r < c_out && cil < nci)
      bwd[(size_t(ci0 + cil) * taps + (taps - 1 - t)) * c_out_pad + co0 + r] = __float2bfloat16_rn(s_tile[r * pitch + cil * taps + t]);
  }
}

// packed fp32 [c_out_pad][taps][c_in_pad] -> OIHW, one block per output channel through shared memory so that both
// the packed rows and the OIHW row are accessed contiguously
template <int TAPS>
__global__ void __launch_bounds__(256) k_unpack_wgrad_rows(const float* __restrict__ packed, int c_in, int /*taps*/, int c_in_pad,
                                                          float* __restrict__ grad) {
  extern __shared__ float s_tile[];  // [taps][c_in]
  constexpr int taps = TAPS;
  const int co = blockIdx.x;
  for (int i = threadIdx.x; i < taps * c_in; i += 256) {
    const int t = i / c_in, ci = i - t * c_in;
    s_tile[i] = packed[(size_t(co) * taps + t) * c_in_pad + ci];
  }
  __syncthreads();
  float* g = grad + size_t(co) * c_in * taps;
  for (int i = threadIdx.x; i < taps * c_in; i += 256) {
    const int ci = i / taps, t = i - ci * taps;
    g[i] = s_tile[t * c_in + ci];
  }
}

// packed fp32 [c_out_pad][taps][c_in_pad] -> OIHW (c_out, c_in, k, k); stem: packed [c_out_pad][32] with
// K index (kh*3+kw)*c_in + c (the patch-matrix order of yolo_input_patchify)
__global__ void k_unpack_wgrad(const float* __restrict__ packed, int c_out, int c_in, int taps, int c_in_pad, int stem,
                               float* __restrict__ grad) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)c_out * c_in * taps;
  if (idx >= total) return;
  const int tap = int(idx % taps);
  const int ci = int((idx / taps) % c_in);
  const int co = int(idx / ((long long)taps * c_in));
  grad[idx] = stem ? packed[size_t(co) * c_in_pad + tap * c_in + ci] : packed[(size_t(co) * taps + tap) * c_in_pad + ci];
}

// data-gradient weights: out[ci][tap'][co] = w[co][ci][taps-1-tap'] (transposed, spatially flipped), bf16,
// zero padded to [rows_pad][taps][cols_pad] -- the K-major B operand of the forward kernel run on dz
__global__ void k_pack_weights_dgrad(const float* __restrict__ w, int c_out, int c_in, int taps, int rows_pad, int cols_pad,
                                     __nv_bfloat16* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)rows_pad * taps * cols_pad;
  if (idx >= total) return;
  const int co = int(idx % cols_pad);
  const int tap = int((idx / cols_pad) % taps);
  const int ci = int(idx / ((long long)cols_pad * taps));
  float v = 0.f;
  if (co < c_out && ci < c_in) v = w[(size_t(co) * c_in + ci) * taps + (taps - 1 - tap)];
  out[idx] = __float2bfloat16_rn(v);
}

// stride-2 data gradient, row parity r: out[(t*rows_half + ci)][da][db][co]; kh = 1 (r = 0) or 2 - 2*da (r = 1);
// t = 0 uses kw = 1 at db = 0 only, t = 1 uses kw = 2 at db = 0 and kw = 0 at db = 1
__global__ void k_pack_weights_dgrad_s2(const float* __restrict__ w, int c_out, int c_in, int r, int rows_half, int cols_pad,
                                        __nv_bfloat16* __restrict__ out) {
  const int taps = (r + 1) * 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = 2ll * rows_half * taps * cols_pad;
  if (idx >= total) return;
  const int co = int(idx % cols_pad);
  const int tap = int((idx / cols_pad) % taps);
  const int row = int(idx / ((long long)cols_pad * taps));
  const int t = row / rows_half, ci = row - t * rows_half;
  const int da = tap >> 1, db = tap & 1;
  const int kh = r == 0 ? 1 : 2 - 2 * da;
  const int kw = t == 0 ? (db == 0 ? 1 : -1) : (db == 0 ? 2 : 0);
  float v = 0.f;
  if (kw >= 0 && co < c_out && ci < c_in) v = w[((size_t(co) * c_in + ci) * 3 + kh) * 3 + kw];
  out[idx] = __float2bfloat16_rn(v);
}

// torch.optim.SGD: g' = g * gscale + wd * p;  buf = first ? g' : mu * buf + g';  p -= lr * buf
__global__ void __launch_bounds__(256) k_sgd(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                            long long n, float lr, const float* __restrict__ lr_dev, float mu, float wd,
                                            float gscale, int first) {
  if (lr_dev != nullptr) lr = __ldg(lr_dev);   // a captured step reads its learning rate from device memory
  const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 pv = *reinterpret_cast<float4*>(p + i);
    const float4 gv = *reinterpret_cast<const float4*>(g + i);
    float4 bv = first ? make_float4(0, 0, 0, 0) : *reinterpret_cast<float4*>(buf + i);
    float* pp = &pv.x; const float* gp = &gv.x; float* bp = &bv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gg = fmaf(wd, pp[k], gp[k] * gscale);
      bp[k] = first ? gg : fmaf(mu, bp[k], gg);
      pp[k] -= lr * bp[k];
    }
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(buf + i) = bv;
  } else {
    for (long long j = i; j < n; ++j) {
      const float gg = fmaf(wd, p[j], g[j] * gscale);
      const float b = first ? gg : fmaf(mu, buf[j], gg);
      buf[j] = b;
      p[j] -= lr * b;
    }
  }
}

int check_rows(long long P, int C, int pitch, const char* what) {
  YB_REQUIRE(P >= 1 && C >= 8 && C % 8 == 0 && C <= 2048 && pitch >= C && pitch % 8 == 0, "%s: bad rows (P %lld, C %d, pitch %d)",
             what, P, C, pitch);
  YB_REQUIRE(P * (C / 8) < (1ll << 31), "%s: more than 2^31 (row, channel group) pairs", what);
  return YB_OK;
}
int reduce_grid(long long P, int groups, int per_sm = 3) {
  const int lanes = TR_THREADS / groups;
  long long blocks = (P + (long long)lanes * 8 - 1) / ((long long)lanes * 8);  // >= 8 rows per thread
  const int cap = 148 * per_sm;  // every block ends with one double atomic per channel statistic: the tail grows with the grid
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" int yolo_bn_stats(const void* z, long long P, int C, int pitch, double* sums2c, yb_stream_t stream) {
  YB_REQUIRE(z && sums2c, "yolo_bn_stats: null pointer");
  if (int rc = check_rows(P, C, pitch, "yolo_bn_stats")) return rc;
  RowGeom g{P, C, C / 8, 0, 0};
  k_bn_stats<<<reduce_grid(P, g.groups), TR_THREADS, 0, (cudaStream_t)stream>>>(
      static_cast<const __nv_bfloat16*>(z), pitch, g, sums2c);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_bn_finalize(const double* sums2c, long long P, int C, const float* gamma, const float* beta, float eps,
                                float momentum, float* running_mean, float* running_var, float* mean, float* rstd,
                                float* scale, float* bias, yb_stream_t stream) {
  YB_REQUIRE(sums2c && gamma && beta && mean && rstd && scale && bias && P >= 1 && C >= 1, "yolo_bn_finalize: bad argument");
  const BnFinalize f{P, gamma, beta, eps, momentum, running_mean, running_var, mean, rstd, scale, bias};
  k_bn_finalize<<<yb_cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(sums2c, C, f);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_bn_stats_finalize(const void* z, long long P, int C, int pitch, double* sums2c, unsigned int* counter,
                                      const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                                      float* running_var, float* mean, float* rstd, float* scale, float* bias,
                                      yb_stream_t stream) {
  YB_REQUIRE(z && sums2c && counter && gamma && beta && mean && rstd && scale && bias, "yolo_bn_stats_finalize: null pointer");
  if (int rc = check_rows(P, C, pitch, "yolo_bn_stats_finalize")) return rc;
  RowGeom g{P, C, C / 8, 0, 0};
  const BnFinalize f{P, gamma, beta, eps, momentum, running_mean, running_var, mean, rstd, scale, bias};
  k_bn_stats_fused<<<reduce_grid(P, g.groups), TR_THREADS, 0, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16*>(z), pitch, g,
                                                                                     sums2c, counter, f);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_bn_act_fwd(const void* z, long long P, int C, int z_pitch, const float* scale, const float* bias, int act,
                               const void* residual, int res_pitch, void* y, int y_pitch, int up2x, int h, int w,
                               yb_stream_t stream) {
  YB_REQUIRE(z && scale && bias && y, "yolo_bn_act_fwd: null pointer");
  if (int rc = check_rows(P, C, z_pitch, "yolo_bn_act_fwd")) return rc;
  YB_REQUIRE(y_pitch >= C && y_pitch % 8 == 0 && (!residual || (res_pitch >= C && res_pitch % 8 == 0)), "yolo_bn_act_fwd: bad pitch");
  YB_REQUIRE(!up2x || (h >= 1 && w >= 1 && P % ((long long)h * w) == 0), "yolo_bn_act_fwd: bad 2x geometry");
  RowGeom g{P, C, C / 8, h, w};
  const long long n = P * g.groups;
  if (up2x) {
    k_bn_act_fwd_up2x<<<(unsigned)((n + TR_THREADS - 1) / TR_THREADS), TR_THREADS, 0, (cudaStream_t)stream>>>(
        static_cast<const __nv_bfloat16*>(z), z_pitch, g, scale, bias, act, static_cast<const __nv_bfloat16*>(residual), res_pitch,
        static_cast<__nv_bfloat16*>(y), y_pitch);
  } else {
    const unsigned nq = ew_quarter((unsigned)n, (unsigned)g.groups);
    k_bn_act_fwd<<<(nq + TR_THREADS - 1) / TR_THREADS, TR_THREADS, 0, (cudaStream_t)stream>>>(
        static_cast<const __nv_bfloat16*>(z), z_pitch, g, scale, bias, act, static_cast<const __nv_bfloat16*>(residual), res_pitch,
        static_cast<__nv_bfloat16*>(y), y_pitch, nq);
  }
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_bn_act_bwd(const void* dA, int dA_pitch, int up2x, const void* z, int z_pitch, long long P, int C, int h, int w,
                               const float* scale, const float* bias, const float* mean, const float* rstd, int act,
                               double* sums2c /* zeroed */, unsigned int* counter /* zero */, float* dgamma, float* dbeta,
                               float* c1c0 /* [2C] scratch */, void* dz, int dz_pitch, void* stuffed, int stuffed_pitch,
                               yb_stream_t stream_) {
  YB_REQUIRE(dA && z && scale && bias && mean && rstd && sums2c && counter && dgamma && dbeta && c1c0 && dz,
             "yolo_bn_act_bwd: null pointer");
  if (int rc = check_rows(P, C, z_pitch, "yolo_bn_act_bwd")) return rc;
  YB_REQUIRE(dA_pitch >= C && dA_pitch % 8 == 0 && dz_pitch >= C && dz_pitch % 8 == 0, "yolo_bn_act_bwd: bad pitch");
  YB_REQUIRE((!up2x && !stuffed) || (h >= 1 && w >= 1 && P % ((long long)h * w) == 0), "yolo_bn_act_bwd: bad 2x geometry");
  YB_REQUIRE(!stuffed || (stuffed_pitch >= C && stuffed_pitch % 8 == 0), "yolo_bn_act_bwd: bad stuffed pitch");
  cudaStream_t stream = (cudaStream_t)stream_;
  BnBwdParams p;
  p.dA = static_cast<const __nv_bfloat16*>(dA); p.dA_pitch = dA_pitch; p.up2x = up2x;
  p.z = static_cast<const __nv_bfloat16*>(z); p.z_pitch = z_pitch;
  p.scale = scale; p.bias = bias; p.mean = mean; p.rstd = rstd; p.act = act;
  p.g = RowGeom{P, C, C / 8, h, w};
  const BnBwdFinalize f{P, dgamma, dbeta, c1c0, c1c0 + C};
  k_bn_act_bwd_reduce<<<reduce_grid(P, p.g.groups), TR_THREADS, 0, stream>>>(p, sums2c, counter, f,
                                                                             static_cast<__nv_bfloat16*>(dz), dz_pitch);
  YB_CHECK_LAUNCH();
  const long long n = P * p.g.groups;
  if (stuffed) {
    k_bn_act_bwd_apply_stuffed<<<(unsigned)((n + TR_THREADS - 1) / TR_THREADS), TR_THREADS, 0, stream>>>(
        p, c1c0, c1c0 + C, static_cast<__nv_bfloat16*>(dz), dz_pitch, static_cast<__nv_bfloat16*>(stuffed), stuffed_pitch);
  } else {
    const unsigned nq = ew_quarter((unsigned)n, (unsigned)p.g.groups);
    k_bn_act_bwd_apply<<<(nq + TR_THREADS - 1) / TR_THREADS, TR_THREADS, 0, stream>>>(p, c1c0, c1c0 + C,
                                                                                      static_cast<__nv_bfloat16*>(dz), dz_pitch, nq);
  }
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_bias_grad(const void* dz, long long P, int C_pad, int pitch, int C, double* sums2c /* zeroed */, float* dbias,
                              yb_stream_t stream_) {
  YB_REQUIRE(dz && sums2c && dbias && C >= 1 && C <= C_pad, "yolo_bias_grad: bad argument");
  if (int rc = check_rows(P, C_pad, pitch, "yolo_bias_grad")) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  RowGeom g{P, C_pad, C_pad / 8, 0, 0};
  k_bn_stats<<<reduce_grid(P, g.groups), TR_THREADS, 0, stream>>>(static_cast<const __nv_bfloat16*>(dz),
                                                                                          pitch, g, sums2c);
  YB_CHECK_LAUNCH();
  k_bias_finalize<<<yb_cdiv(C, 128), 128, 0, stream>>>(sums2c, C, dbias);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_unpack_wgrad(const float* packed, int c_out, int c_in, int ksize, int c_in_pad, int stem, float* grad_oihw,
                                 yb_stream_t stream) {
  YB_REQUIRE(packed && grad_oihw && c_out >= 1 && c_in >= 1 && (ksize == 1 || ksize == 3) && c_in_pad >= (stem ? 9 * c_in : c_in),
             "yolo_unpack_wgrad: bad argument");
  const long long total = (long long)c_out * c_in * ksize * ksize;
  const size_t row_bytes = size_t(ksize) * ksize * c_in * sizeof(float);
  if (!stem && row_bytes <= 48 * 1024) {
    if (ksize == 3)
      k_unpack_wgrad_rows<9><<<c_out, 256, row_bytes, (cudaStream_t)stream>>>(packed, c_in, 9, c_in_pad, grad_oihw);
    else
      k_unpack_wgrad_rows<1><<<c_out, 256, row_bytes, (cudaStream_t)stream>>>(packed, c_in, 1, c_in_pad, grad_oihw);
  } else {
    k_unpack_wgrad<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(packed, c_out, c_in, ksize * ksize, c_in_pad,
                                                                                    stem, grad_oihw);
  }
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_pack_weights_dgrad(const float* w_oihw, int c_out, int c_in, int ksize, int rows_pad, int cols_pad,
                                       void* w_packed, yb_stream_t stream) {
  YB_REQUIRE(w_oihw && w_packed && c_out >= 1 && c_in >= 1 && (ksize == 1 || ksize == 3) && rows_pad >= c_in && cols_pad >= c_out,
             "yolo_pack_weights_dgrad: bad argument");
  const long long total = (long long)rows_pad * ksize * ksize * cols_pad;
  k_pack_weights_dgrad<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, c_out, c_in, ksize * ksize, rows_pad, cols_pad, static_cast<__nv_bfloat16*>(w_packed));
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_pack_weights_train(const float* w_oihw, int c_out, int c_in, int ksize, int c_in_pad, int c_out_pad,
                                       void* w_fwd, void* w_dgrad, yb_stream_t stream) {
  YB_REQUIRE(w_oihw && w_fwd && w_dgrad && c_out >= 1 && c_in >= 1 && (ksize == 1 || ksize == 3) && c_in_pad >= c_in &&
                 c_out_pad >= c_out, "yolo_pack_weights_train: bad argument");
  const int taps = ksize * ksize;
  dim3 grid((c_in + PK_T - 1) / PK_T, (c_out + PK_T - 1) / PK_T);
  const size_t smem = PK_T * (PK_T * taps + 1) * sizeof(float);
  if (taps == 9)
    k_pack_weights_both<9><<<grid, 256, smem, (cudaStream_t)stream>>>(w_oihw, c_out, c_in, taps, c_in_pad, c_out_pad,
                                                                     static_cast<__nv_bfloat16*>(w_fwd), static_cast<__nv_bfloat16*>(w_dgrad));
  else
    k_pack_weights_both<1><<<grid, 256, smem, (cudaStream_t)stream>>>(w_oihw, c_out, c_in, taps, c_in_pad, c_out_pad,
                                                                     static_cast<__nv_bfloat16*>(w_fwd), static_cast<__nv_bfloat16*>(w_dgrad));
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_pack_weights_dgrad_s2(const float* w_oihw, int c_out, int c_in, int r, int rows_half, int cols_pad,
                                          void* w_packed, yb_stream_t stream) {
  YB_REQUIRE(w_oihw && w_packed && c_out >= 1 && c_in >= 1 && (r == 0 || r == 1) && rows_half >= c_in && cols_pad >= c_out,
             "yolo_pack_weights_dgrad_s2: bad argument");
  const long long total = 2ll * rows_half * (r + 1) * 2 * cols_pad;
  k_pack_weights_dgrad_s2<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, c_out, c_in, r, rows_half, cols_pad, static_cast<__nv_bfloat16*>(w_packed));
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_sgd_step(float* param, const float* grad, float* momentum_buf, long long n, float lr, float momentum,
                             float weight_decay, float grad_scale, int first_step, yb_stream_t stream) {
  YB_REQUIRE(param && grad && momentum_buf && n >= 0, "yolo_sgd_step: bad argument");
  YB_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(momentum_buf)) & 15) == 0,
             "yolo_sgd_step: buffers must be 16-byte aligned");
  if (n == 0) return YB_OK;
  const long long threads = (n + 3) / 4;
  k_sgd<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, n, lr, nullptr, momentum,
                                                                            weight_decay, grad_scale, first_step);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

extern "C" int yolo_sgd_step_dev(float* param, const float* grad, float* momentum_buf, long long n, const float* lr_dev,
                                 float momentum, float weight_decay, float grad_scale, int first_step, yb_stream_t stream) {
  YB_REQUIRE(param && grad && momentum_buf && lr_dev && n >= 0, "yolo_sgd_step_dev: bad argument");
  YB_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(momentum_buf)) & 15) == 0,
             "yolo_sgd_step_dev: buffers must be 16-byte aligned");
  if (n == 0) return YB_OK;
  const long long threads = (n + 3) / 4;
  k_sgd<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, n, 0.f, lr_dev, momentum,
                                                                            weight_decay, grad_scale, first_step);
  YB_CHECK_LAUNCH();
  return YB_OK;
}
