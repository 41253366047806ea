// K1/K2: fused convolution for the Darknet-53 + 3-scale head forward pass.
//
//   y = act( conv(x, W) * scale + bias ) (+ residual)       NHWC bf16, fp32 accumulate
//
// replaces, per CNNBlock (model.py:80-86), cudnn_convolution + cudnn_batch_norm
// + leaky_relu/mish (3 launches, 3 HBM round trips), the separate residual add
// of ResidualBlock.forward (model.py:118), nn.Upsample (model.py:222, fused as
// a 2x2 replicated store) and torch.cat (model.py:190, fused as channel-pitched
// stores/loads into one buffer).
//
// Implicit GEMM on the 5th-gen tensor cores:
//   D[M = B*Ho*Wo, N = Cout] = A[M, K = k*k*Cin] * W[N, K]^T
//   - A tiles (128 output pixels x KC channels of one filter tap) arrive by TMA:
//     im2col-mode tensor maps for 3x3 (padding = hardware zero fill, stride =
//     traversal stride), plain tiled maps for 1x1;  128B/64B swizzle.
//   - W tiles (BLOCK_N x KC, K-major) arrive by tiled TMA.
//   - one elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) with the
//     fp32 accumulator in TMEM; tcgen05.commit releases smem stages.
//   - 4 epilogue warps read TMEM with tcgen05.ld, apply folded BN + activation
//     (+ residual), convert to bf16 and store.
// Warp roles: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2..5 = epilogue.
#include <string.h>

#include <new>

#include "conv_plan.cuh"
#include "conv_ptx.cuh"

namespace {

constexpr int CONV_THREADS = 192;

using namespace convptx;

// ---------------------------------------------------------------- the kernel
template <int BLOCK_N, int KC>
__global__ void __launch_bounds__(CONV_THREADS)
k_conv_tcgen05(const __grid_constant__ ConvKParams p) {
  constexpr int ROW_BYTES = KC * 2;
  constexpr uint32_t A_BYTES = BLOCK_M * ROW_BYTES;
  constexpr uint32_t B_BYTES = BLOCK_N * ROW_BYTES;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BLOCK_N >> 3) << 17) |
                             (uint32_t(BLOCK_M >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stages = p.stages;
  const uint32_t bar_base = smem_base + stages * STAGE_BYTES;  // full[s], empty[s], tmem_full
  const uint32_t tmem_slot = bar_base + (2 * stages + 1) * 8;
  auto full_bar = [&](int s) { return bar_base + s * 8; };
  auto empty_bar = [&](int s) { return bar_base + (stages + s) * 8; };
  const uint32_t tmem_full_bar = bar_base + 2 * stages * 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1-D grid (gridDim.y is limited to 65535): N-tiles of one M-tile are adjacent so that the CTAs
  // sharing an A tile run together and hit it in L2.
  const int m0 = int(blockIdx.x / p.tiles_n) * BLOCK_M;
  const int n0 = int(blockIdx.x % p.tiles_n) * BLOCK_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int cw = 0, ch = 0, img = 0;
      if (p.a_im2col) {
        const int hw = p.h_out * p.w_out;
        img = m0 / hw;
        const int rem = m0 - img * hw;
        const int po = rem / p.w_out, qo = rem - po * p.w_out;
        cw = qo * p.stride_w - p.pad;  // top-left tap of the first output pixel, input coords
        ch = po * p.stride - p.pad;
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
        const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
        if (p.a_im2col) {
          const int r = tap / p.ksize_w, t = tap - r * p.ksize_w;
          tma_load_im2col_4d(&p.tmA, full_bar(s), sa, cc * KC, cw, ch, img, (uint16_t)t, (uint16_t)r);
        } else {
          tma_load_2d(&p.tmA, full_bar(s), sa, cc * KC, m0);
        }
        tma_load_2d(&p.tmB, full_bar(s), sb, kb * KC, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
        const uint64_t adesc = make_kmajor_desc<ROW_BYTES>(sa);
        const uint64_t bdesc = make_kmajor_desc<ROW_BYTES>(sb);
#pragma unroll
        for (int k = 0; k < KC / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in (addr >> 4) units
          umma_bf16(tmem_base, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), IDESC,
                    (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));  // implies tcgen05.fence::before_thread_sync
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int quad = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int row = quad * 32 + lane;
    const int m = m0 + row;
    const bool valid = m < p.M;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    size_t out_row[4];
    int n_out_rows = 1;
    if (p.upsample2x) {
      const int hw = p.h_out * p.w_out;
      const int img = m / hw;
      const int rem = m - img * hw;
      const int po = rem / p.w_out, qo = rem - po * p.w_out;
      const int W2 = 2 * p.w_out;
      const size_t r00 = (size_t(img) * (2 * p.h_out) + 2 * po) * W2 + 2 * qo;
      out_row[0] = r00; out_row[1] = r00 + 1; out_row[2] = r00 + W2; out_row[3] = r00 + W2 + 1;
      n_out_rows = 4;
    } else {
      out_row[0] = size_t(m);
    }
    bool saw_nan = false;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(c0), v);
      if (!valid) continue;
      const int n = n0 + c0;
      float o[32];
      const float4* sp = reinterpret_cast<const float4*>(p.scale + n);
      const float4* bp = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 s4 = __ldg(sp + j), b4 = __ldg(bp + j);
        o[4 * j + 0] = apply_act(fmaf(__uint_as_float(v[4 * j + 0]), s4.x, b4.x), p.act);
        o[4 * j + 1] = apply_act(fmaf(__uint_as_float(v[4 * j + 1]), s4.y, b4.y), p.act);
        o[4 * j + 2] = apply_act(fmaf(__uint_as_float(v[4 * j + 2]), s4.z, b4.z), p.act);
        o[4 * j + 3] = apply_act(fmaf(__uint_as_float(v[4 * j + 3]), s4.w, b4.w), p.act);
      }
      if (p.has_residual) {
        const uint4* rp = reinterpret_cast<const uint4*>(
            static_cast<const __nv_bfloat16*>(p.residual) + size_t(m) * p.res_pitch + n);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 r = __ldg(rp + j);
          o[8 * j + 0] += bf16_lo(r.x); o[8 * j + 1] += bf16_hi(r.x);
          o[8 * j + 2] += bf16_lo(r.y); o[8 * j + 3] += bf16_hi(r.y);
          o[8 * j + 4] += bf16_lo(r.z); o[8 * j + 5] += bf16_hi(r.z);
          o[8 * j + 6] += bf16_lo(r.w); o[8 * j + 7] += bf16_hi(r.w);
        }
      }
      if (p.check_nan) {
#pragma unroll
        for (int j = 0; j < 32; ++j) saw_nan |= (o[j] != o[j]);
      }
      if (p.out_fp32) {
        for (int rr = 0; rr < n_out_rows; ++rr) {
          float4* yp = reinterpret_cast<float4*>(static_cast<float*>(p.y) +
                                                 out_row[rr] * p.out_pitch + n);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            yp[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }
      } else {
        uint4 w4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          w4[j].x = pack_bf16(o[8 * j + 0], o[8 * j + 1]);
          w4[j].y = pack_bf16(o[8 * j + 2], o[8 * j + 3]);
          w4[j].z = pack_bf16(o[8 * j + 4], o[8 * j + 5]);
          w4[j].w = pack_bf16(o[8 * j + 6], o[8 * j + 7]);
        }
        for (int rr = 0; rr < n_out_rows; ++rr) {
          uint4* yp = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.y) +
                                               out_row[rr] * p.out_pitch + n);
#pragma unroll
          for (int j = 0; j < 4; ++j) yp[j] = w4[j];
        }
      }
    }
    if (saw_nan) atomicOr(p.status, YB_STATUS_NAN_LAYER);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------- test-only SIMT reference
__global__ void k_conv_simt(const yolo_conv_desc d, const __nv_bfloat16* __restrict__ x,
                            const __nv_bfloat16* __restrict__ w, const float* __restrict__ scale,
                            const float* __restrict__ bias, const __nv_bfloat16* __restrict__ res,
                            void* __restrict__ y, uint32_t* status, int h_out, int w_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long M = (long long)d.batch * h_out * w_out;
  if (idx >= M * d.c_out_pad) return;
  const int n = int(idx % d.c_out_pad);
  const long long m = idx / d.c_out_pad;
  const int img = int(m / (h_out * w_out));
  const int rem = int(m - (long long)img * h_out * w_out);
  const int po = rem / w_out, qo = rem - po * w_out;
  float acc = 0.f;
  const int kw = d.ksize_w > 0 ? d.ksize_w : d.ksize, sw = d.stride_w > 0 ? d.stride_w : d.stride;
  for (int r = 0; r < d.ksize; ++r) {
    const int hi = po * d.stride - d.pad + r;
    if (hi < 0 || hi >= d.h_in) continue;
    for (int t = 0; t < kw; ++t) {
      const int wi = qo * sw - d.pad + t;
      if (wi < 0 || wi >= d.w_in) continue;
      const __nv_bfloat16* xp = x + ((size_t(img) * d.h_in + hi) * d.w_in + wi) * d.in_pitch;
      const __nv_bfloat16* wp = w + (size_t(n) * d.ksize * kw + r * kw + t) * d.c_in;
      for (int c = 0; c < d.c_in; ++c) acc += __bfloat162float(xp[c]) * __bfloat162float(wp[c]);
    }
  }
  float o = apply_act(fmaf(acc, scale[n], bias[n]), d.act);
  if (d.has_residual) o += __bfloat162float(res[size_t(m) * d.res_pitch + n]);
  if (d.check_nan && o != o) atomicOr(status, YB_STATUS_NAN_LAYER);
  size_t rows[4];
  int nr = 1;
  if (d.upsample2x) {
    const int W2 = 2 * w_out;
    const size_t r00 = (size_t(img) * (2 * h_out) + 2 * po) * W2 + 2 * qo;
    rows[0] = r00; rows[1] = r00 + 1; rows[2] = r00 + W2; rows[3] = r00 + W2 + 1;
    nr = 4;
  } else {
    rows[0] = size_t(m);
  }
  for (int rr = 0; rr < nr; ++rr) {
    if (d.out_fp32) static_cast<float*>(y)[rows[rr] * d.out_pitch + n] = o;
    else static_cast<__nv_bfloat16*>(y)[rows[rr] * d.out_pitch + n] = __float2bfloat16_rn(o);
  }
}

// ---------------------------------------------------------------- host side
void* driver_fn(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  return fn;
}

int validate_desc(const yolo_conv_desc* d, int* h_out, int* w_out) {
  YB_REQUIRE(d, "conv: null desc");
  YB_REQUIRE(d->batch >= 1 && d->h_in >= 1 && d->w_in >= 1, "conv: bad input shape");
  YB_REQUIRE(d->c_in >= 32 && d->c_in % 32 == 0, "conv: c_in (%d) must be a multiple of 32", d->c_in);
  YB_REQUIRE(d->in_pitch >= d->c_in && d->in_pitch % 8 == 0, "conv: bad in_pitch %d", d->in_pitch);
  YB_REQUIRE(d->c_out >= 1 && d->c_out_pad >= d->c_out && d->c_out_pad % 32 == 0,
             "conv: c_out_pad (%d) must be a multiple of 32 covering c_out (%d)", d->c_out_pad, d->c_out);
  YB_REQUIRE(d->out_pitch % 8 == 0, "conv: bad out_pitch %d", d->out_pitch);
  if (d->s2_parity) {
    YB_REQUIRE((d->s2_parity == 1 || d->s2_parity == 2) && d->ksize == d->s2_parity && yb_kw(d) == 2 && d->pad == 0 &&
                   d->stride == 1 && yb_sw(d) == 1 && yb_pad_h_hi(d) == d->ksize - 1 && yb_pad_hi(d) == 1 &&
                   d->s2_cin >= 32 && d->s2_cin % 32 == 0 && d->c_out_pad == 2 * d->s2_cin && d->c_out == d->c_out_pad &&
                   !d->upsample2x && !d->out_fp32 && d->stem_c == 0 && d->out_pitch >= d->s2_cin,
               "conv: bad stride-2 data-gradient geometry");
  } else {
    YB_REQUIRE((d->ksize == 1 && d->pad == 0) || (d->ksize == 3 && d->pad == 1),
               "conv: only 1x1/pad0 and 3x3/pad1 (model.py:201)");
    YB_REQUIRE(d->out_pitch >= d->c_out_pad, "conv: bad out_pitch %d", d->out_pitch);
    YB_REQUIRE(!d->has_residual || d->res_pitch >= d->c_out_pad, "conv: bad res_pitch");
  }
  YB_REQUIRE(d->stride == 1 || d->stride == 2, "conv: stride must be 1 or 2");
  YB_REQUIRE(yb_kw(d) >= 1 && yb_kw(d) <= 3 && (yb_sw(d) == 1 || yb_sw(d) == 2) && yb_pad_hi(d) >= 0 && yb_pad_hi(d) <= 1,
             "conv: bad rectangular geometry (ksize_w %d stride_w %d)", d->ksize_w, d->stride_w);
  YB_REQUIRE(d->act >= YB_ACT_NONE && d->act <= YB_ACT_MISH, "conv: bad activation code %d", d->act);
  YB_REQUIRE(!d->has_residual || d->res_pitch % 8 == 0, "conv: bad res_pitch");
  YB_REQUIRE(!(d->has_residual && d->out_fp32), "conv: residual with fp32 output unsupported");
  *h_out = (d->h_in + d->pad + yb_pad_h_hi(d) - d->ksize) / d->stride + 1;
  *w_out = (d->w_in + d->pad + yb_pad_hi(d) - yb_kw(d)) / yb_sw(d) + 1;
  return YB_OK;
}

template <int BN, int KC>
int launch_conv(const ConvPlan* pl, const ConvKParams& kp, cudaStream_t stream) {
  YB_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tcgen05<BN, KC>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, pl->smem_bytes));
  k_conv_tcgen05<BN, KC><<<dim3((unsigned)pl->grid_x * (unsigned)pl->grid_y), CONV_THREADS, pl->smem_bytes, stream>>>(kp);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

}  // namespace

extern "C" size_t yolo_conv_plan_bytes(void) { return sizeof(ConvPlan) + 64; }

extern "C" int yolo_conv_plan_init(void* plan_host, size_t plan_bytes, const yolo_conv_desc* d,
                                   const void* x, const void* w_packed, const float* scale,
                                   const float* bias, const void* residual, void* y) {
  YB_REQUIRE(plan_host && plan_bytes >= sizeof(ConvPlan), "conv plan: buffer too small");
  YB_REQUIRE((reinterpret_cast<uintptr_t>(plan_host) & 63) == 0, "conv plan: buffer must be 64B aligned");
  int h_out, w_out;
  int rc = validate_desc(d, &h_out, &w_out);
  if (rc) return rc;
  const bool stem = d->stem_c > 0;
  YB_REQUIRE((x || stem) && w_packed && scale && bias && y, "conv plan: null tensor pointer");
  YB_REQUIRE(!d->has_residual || residual, "conv plan: residual pointer missing");
  YB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
             "conv plan: tensors must be 16-byte aligned");

  static PFN_encodeTiled encTiled = (PFN_encodeTiled)driver_fn("cuTensorMapEncodeTiled");
  static PFN_encodeIm2col encIm2col = (PFN_encodeIm2col)driver_fn("cuTensorMapEncodeIm2col");
  if (!encTiled || !encIm2col) {
    yb_set_error("conv plan: cuTensorMapEncode* driver entry points unavailable (no GPU driver?)");
    return YB_ERR_CUDA;
  }

  ConvPlan* pl = new (plan_host) ConvPlan();
  pl->d = *d;
  const int kc = (d->c_in % 64 == 0) ? 64 : 32;
  int bn = d->block_n_hint;  // a hint: ignored when it does not tile this layer
  if (bn != 32 && bn != 64 && bn != 128 && bn != 256) bn = 0;
  if (bn != 0 && d->c_out_pad % bn != 0) bn = 0;
  if (bn == 0) bn = d->c_out_pad % 128 == 0 ? 128 : (d->c_out_pad % 64 == 0 ? 64 : 32);
  YB_REQUIRE((bn == 32 || bn == 64 || bn == 128 || bn == 256) && d->c_out_pad % bn == 0,
             "conv plan: block_n %d does not tile c_out_pad %d", bn, d->c_out_pad);
  const int taps = d->ksize * yb_kw(d);
  const int cchunks = d->c_in / kc;
  const int num_kb = taps * cchunks;
  const bool plain_1x1 = d->ksize == 1 && yb_kw(d) == 1 && d->stride == 1 && yb_sw(d) == 1;
  const int im2col = (d->a_mode == 2) || (d->a_mode == 0 && !plain_1x1);
  YB_REQUIRE(im2col || plain_1x1, "conv plan: tiled A needs 1x1 stride 1");
  const long long M = (long long)d->batch * h_out * w_out;
  YB_REQUIRE(M < (1ll << 31), "conv plan: too many output pixels");

  const int stage_bytes = (BLOCK_M + bn) * kc * 2;
  int stages = d->stages_hint;
  if (stages == 0) {
    const int budget = (bn == 256) ? 200 * 1024 : 100 * 1024;  // <=128-wide tiles: 2 CTAs per SM
    stages = budget / stage_bytes;
    if (stages > 8) stages = 8;
  }
  if (stages > num_kb) stages = num_kb;
  if (stages < 1) stages = 1;
  pl->smem_bytes = stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
  if (d->impl_hint == 1)
    YB_REQUIRE(pl->smem_bytes <= 227 * 1024, "conv plan: %d stages of %d B exceed shared memory", stages, stage_bytes);

  const CUtensorMapSwizzle swz = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult cr;
  if (stem) {
    cr = CUDA_SUCCESS;  // no A tensor map: the stem kernel gathers A from the NCHW image itself
    memset(&pl->kp.tmA, 0, sizeof(pl->kp.tmA));
  } else if (im2col) {
    cuuint64_t dims[4] = {(cuuint64_t)d->c_in, (cuuint64_t)d->w_in, (cuuint64_t)d->h_in, (cuuint64_t)d->batch};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_pitch * 2, (cuuint64_t)d->w_in * d->in_pitch * 2,
                             (cuuint64_t)d->h_in * d->w_in * d->in_pitch * 2};
    int lower[2] = {-d->pad, -d->pad};  // {W, H}
    int upper[2] = {yb_pad_hi(d) - (yb_kw(d) - 1), yb_pad_h_hi(d) - (d->ksize - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)yb_sw(d), (cuuint32_t)d->stride, 1};
    cr = encIm2col(&pl->kp.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides,
                   lower, upper, (cuuint32_t)kc, (cuuint32_t)BLOCK_M, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    // Same workaround CUTLASS applies (copy_traits_sm90_im2col.hpp) for drivers <= 13.1 on
    // tensors smaller than 128 KiB.
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const unsigned long long bytes = (unsigned long long)d->batch * d->h_in * d->w_in * d->in_pitch * 2ull;
    if (cr == CUDA_SUCCESS && drv <= 13010 && bytes < 131072ull)
      reinterpret_cast<uint64_t*>(&pl->kp.tmA)[1] &= ~(1ull << 21);
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)d->c_in, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)d->in_pitch * 2};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)BLOCK_M};
    cuuint32_t estr[2] = {1, 1};
    cr = encTiled(&pl->kp.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (cr != CUDA_SUCCESS) {
    yb_set_error("conv plan: tensor map A encode failed (CUresult %d, im2col %d)", (int)cr, im2col);
    return YB_ERR_CUDA;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)taps * d->c_in, (cuuint64_t)d->c_out_pad};
    cuuint64_t strides[1] = {(cuuint64_t)taps * d->c_in * 2};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    cr = encTiled(&pl->kp.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      yb_set_error("conv plan: tensor map B encode failed (CUresult %d)", (int)cr);
      return YB_ERR_CUDA;
    }
  }
  ConvKParams& kp = pl->kp;
  kp.scale = scale; kp.bias = bias; kp.residual = residual; kp.y = y; kp.status = nullptr;
  kp.M = (int)M; kp.h_out = h_out; kp.w_out = w_out;
  kp.out_pitch = d->out_pitch; kp.res_pitch = d->res_pitch;
  kp.num_kb = num_kb; kp.cchunks = cchunks; kp.stages = stages; kp.tiles_n = d->c_out_pad / bn;
  kp.ksize = d->ksize; kp.stride = d->stride; kp.pad = d->pad; kp.ksize_w = yb_kw(d); kp.stride_w = yb_sw(d);
  kp.act = d->act; kp.has_residual = d->has_residual; kp.upsample2x = d->upsample2x;
  kp.out_fp32 = d->out_fp32; kp.check_nan = d->check_nan; kp.a_im2col = im2col;
  pl->block_n = bn; pl->kc = kc;
  pl->grid_x = d->c_out_pad / bn;
  pl->grid_y = (int)((M + BLOCK_M - 1) / BLOCK_M);
  pl->w = w_packed;
  pl->impl = (d->impl_hint == 1 && !stem && !d->s2_parity) ? 1 : 2;
  pl->stem_direct = 0;
  pl->ncta = 1;
  if (pl->impl == 2) {
    rc = conv2_plan_setup(pl, d, h_out, w_out, im2col, encTiled, x, residual, y);
    if (rc) return rc;
  }
  pl->magic = PLAN_MAGIC;
  return YB_OK;
}

extern "C" int yolo_conv_plan_info(const void* plan_host, int32_t* info5) {
  const ConvPlan* pl = static_cast<const ConvPlan*>(plan_host);
  YB_REQUIRE(pl && pl->magic == PLAN_MAGIC && info5, "conv plan info: bad plan");
  info5[0] = pl->block_n; info5[1] = pl->kc; info5[2] = pl->impl == 2 ? pl->kp2.stages : pl->kp.stages;
  info5[3] = pl->grid_x; info5[4] = pl->grid_y;
  info5[5] = (pl->impl == 2 && pl->kp2.row_mode) ? 3 : pl->impl; info5[6] = pl->ncta; info5[7] = pl->impl == 2 ? pl->grid2 : pl->grid_x * pl->grid_y;
  return YB_OK;
}

extern "C" int yolo_conv_fwd(const void* plan_host, uint32_t* status, yb_stream_t stream_) {
  const ConvPlan* pl = static_cast<const ConvPlan*>(plan_host);
  YB_REQUIRE(pl && pl->magic == PLAN_MAGIC, "conv fwd: plan not initialised");
  YB_REQUIRE(status || !pl->kp.check_nan, "conv fwd: status word required when check_nan is set");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (pl->impl == 2) return conv2_launch(pl, status, stream);
  ConvKParams kp = pl->kp;
  kp.status = status;
#define YB_CONV_CASE(BN, KC) \
  if (pl->block_n == BN && pl->kc == KC) return launch_conv<BN, KC>(pl, kp, stream);
  YB_CONV_CASE(32, 32) YB_CONV_CASE(64, 32) YB_CONV_CASE(128, 32) YB_CONV_CASE(256, 32)
  YB_CONV_CASE(32, 64) YB_CONV_CASE(64, 64) YB_CONV_CASE(128, 64) YB_CONV_CASE(256, 64)
#undef YB_CONV_CASE
  yb_set_error("conv fwd: no kernel for block_n %d kc %d", pl->block_n, pl->kc);
  return YB_ERR_UNSUPPORTED;
}

extern "C" int yolo_conv_max_clusters(int cluster_size, int* max_clusters) { return conv2_query_max_clusters(cluster_size, max_clusters); }

extern "C" int yolo_conv_fwd_trace(const void* plan_host, uint32_t* status, unsigned long long* trace_dev, int box,
                                   yb_stream_t stream_) {
  const ConvPlan* pl = static_cast<const ConvPlan*>(plan_host);
  YB_REQUIRE(pl && pl->magic == PLAN_MAGIC && pl->impl == 2 && trace_dev, "conv trace: needs a persistent-kernel plan");
  return conv2_launch(pl, status, (cudaStream_t)stream_, nullptr, nullptr, nullptr, trace_dev, box);
}

extern "C" int yolo_conv_fwd_stats(const void* plan_host, uint32_t* status, double* sums2c, const yolo_bn_finalize_desc* fin,
                                   yb_stream_t stream_) {
  const ConvPlan* pl = static_cast<const ConvPlan*>(plan_host);
  YB_REQUIRE(pl && pl->magic == PLAN_MAGIC, "conv fwd: plan not initialised");
  YB_REQUIRE(pl->impl == 2 && sums2c, "conv fwd stats: needs the persistent kernel and a sums buffer");
  YB_REQUIRE(status || !pl->kp.check_nan, "conv fwd: status word required when check_nan is set");
  if (!fin) return conv2_launch(pl, status, (cudaStream_t)stream_, sums2c);
  YB_REQUIRE(fin->P >= 1 && fin->gamma && fin->beta && fin->mean && fin->rstd && fin->scale && fin->bias && fin->counter,
             "conv fwd stats: incomplete finalize descriptor");
  const BnFinalize f{fin->P, fin->gamma, fin->beta, fin->eps, fin->momentum, fin->running_mean, fin->running_var,
                     fin->mean, fin->rstd, fin->scale, fin->bias};
  return conv2_launch(pl, status, (cudaStream_t)stream_, sums2c, &f, fin->counter);
}

extern "C" int yolo_conv_fwd_stem(const void* plan_host, const float* x_nchw, uint32_t* status, yb_stream_t stream) {
  const ConvPlan* pl = static_cast<const ConvPlan*>(plan_host);
  YB_REQUIRE(pl && pl->magic == PLAN_MAGIC, "conv stem: plan not initialised");
  return conv2_launch_stem(pl, x_nchw, status, (cudaStream_t)stream);
}

extern "C" int yolo_conv_fwd_simt(const yolo_conv_desc* d, const void* x, const void* w_packed,
                                  const float* scale, const float* bias, const void* residual,
                                  void* y, uint32_t* status, yb_stream_t stream) {
  int h_out, w_out;
  int rc = validate_desc(d, &h_out, &w_out);
  if (rc) return rc;
  const long long total = (long long)d->batch * h_out * w_out * d->c_out_pad;
  k_conv_simt<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      *d, static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w_packed), scale, bias,
      static_cast<const __nv_bfloat16*>(residual), y, status, h_out, w_out);
  YB_CHECK_LAUNCH();
  return YB_OK;
}
