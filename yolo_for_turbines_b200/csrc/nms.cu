// K4 + K5 + K6: confidence-threshold compaction (warp/block ballot scan, order
// preserving), stable segmented sort and per-(image, class) greedy NMS for a
// whole batch.  Replaces utils.py:150-191 (non_max_suppression) including its
// filter (:165), its stable descending sort (:166) and its greedy loop
// (:170-187).  Bit-exact contract: the kept rows and their order are the
// reference's, because (a) the IoU is the reference's fp32 op sequence
// (common.cuh yb_iou), (b) the sort is stable on (score desc, input position)
// and (c) "box k survives iff no earlier surviving box of the same class has
// !(iou < thr)" is the fixed point of the reference's loop.
#include "sort.cuh"

namespace {

#ifdef YB_NMS_PROFILE   // dev build only (scripts/nms_round_profile.py): per-phase clock sums of the register path, segments >= 1024 boxes
__device__ unsigned long long g_nms_prof[8];
#define NMS_T(i) do { if (prof) { const long long t_ = clock64(); pt[i] += t_ - t_last; t_last = t_; } } while (0)
#else
#define NMS_T(i) do { } while (0)
#endif


__device__ __forceinline__ int find_image(const int32_t* __restrict__ off, int batch, int i) {
  int lo = 0, hi = batch;  // largest b with off[b] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// utils.py:165  `box[4] > obj_threshold` is a Python double comparison.
__device__ __forceinline__ bool passes(float score, double thr) { return (double)score > thr; }

__global__ void __launch_bounds__(COMPACT_THREADS)
k_thr_count(const float* __restrict__ boxes, int total, double thr, int32_t* __restrict__ tile_cnt) {
  const int base = blockIdx.x * COMPACT_TILE + threadIdx.x * COMPACT_ITEMS;
  int c = 0;
#pragma unroll
  for (int k = 0; k < COMPACT_ITEMS; ++k) {
    const int i = base + k;
    if (i < total && passes(boxes[size_t(i) * 6 + 4], thr)) ++c;
  }
  int tot;
  block_exclusive_scan<COMPACT_THREADS>(c, &tot);
  if (threadIdx.x == 0) tile_cnt[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(COMPACT_THREADS)
k_thr_write(const float* __restrict__ boxes, const int32_t* __restrict__ img_off, int batch,
            int total, double thr, const int32_t* __restrict__ tile_off,
            uint64_t* __restrict__ key, int32_t* __restrict__ val) {
  const int base = blockIdx.x * COMPACT_TILE + threadIdx.x * COMPACT_ITEMS;
  float sc[COMPACT_ITEMS];
  unsigned flags = 0;
  int c = 0;
#pragma unroll
  for (int k = 0; k < COMPACT_ITEMS; ++k) {
    const int i = base + k;
    sc[k] = (i < total) ? boxes[size_t(i) * 6 + 4] : 0.f;
    if (i < total && passes(sc[k], thr)) { flags |= 1u << k; ++c; }
  }
  int pos = tile_off[blockIdx.x] + block_exclusive_scan<COMPACT_THREADS>(c, nullptr);
  if (!flags) return;
  int b = find_image(img_off, batch, base);
#pragma unroll
  for (int k = 0; k < COMPACT_ITEMS; ++k) {
    if (flags & (1u << k)) {
      const int i = base + k;
      while (b + 1 < batch && img_off[b + 1] <= i) ++b;  // skip empty images too
      // descending score: invert the ascending key
      key[pos] = (uint64_t(uint32_t(b)) << 32) | uint64_t(~yb_float_key_asc(sc[k]));
      val[pos] = i;
      ++pos;
    }
  }
}

// After sort #1 (image, score desc, position): p is the reference's output
// order.  Build the (image, class) grouping key for sort #2.
__global__ void k_class_keys(const float* __restrict__ boxes, const uint64_t* __restrict__ key1,
                             const int32_t* __restrict__ val1, const int32_t* __restrict__ n_dev,
                             uint64_t* __restrict__ key2, int32_t* __restrict__ val2,
                             int32_t* __restrict__ idx1, int class_bits, int32_t* __restrict__ violations) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= *n_dev) return;
  const int idx = val1[p];
  idx1[p] = idx;
  float cls = __fadd_rn(boxes[size_t(idx) * 6 + 5], 0.0f);  // -0.0 == 0.0 in utils.py:178
  uint32_t cb = (cls != cls) ? 0x7fc00000u : __float_as_uint(cls);
  if (class_bits > 0) {  // caller promises integer labels in [0, 2^class_bits): a 1-pass grouping key
    const float lim = float(1 << class_bits);
    if (cls >= 0.f && cls < lim && cls == floorf(cls)) cb = uint32_t(cls);
    else { cb = 0; atomicAdd(violations, 1); }
  }
  // integer labels: (image << class_bits) | class, so the grouping sort covers one contiguous bit range; float labels
  // keep the image in the upper word (k_nms_segments recognises the NaN class by the lower one)
  key2[p] = class_bits > 0 ? (((key1[p] >> 32) << class_bits) | cb) : ((key1[p] & 0xffffffff00000000ull) | cb);
  val2[p] = p;
}

__global__ void k_gather_segments(const float* __restrict__ boxes, const uint64_t* __restrict__ key2,
                                  const int32_t* __restrict__ val2, const int32_t* __restrict__ idx1,
                                  const int32_t* __restrict__ n_dev, int box_format,
                                  float4* __restrict__ cbox, float* __restrict__ area,
                                  uint8_t* __restrict__ suppressed, int32_t* __restrict__ seg_starts,
                                  int32_t* __restrict__ nseg) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= *n_dev) return;
  const float* r = boxes + size_t(idx1[val2[q]]) * 6;
  const float w = r[2], h = r[3];
  const CBox c = yb_make_cbox(r[0], r[1], w, h, box_format);
  float a = __fmul_rn(w, h);  // utils.py:79-80
  // torch.max/min propagate NaN (utils.py:70-73): a NaN corner makes every IoU with this box NaN, and
  // so does a NaN area.  Fold both into area = NaN so that the hot loop can use plain fmaxf/fminf.
  if (c.x1 != c.x1 || c.y1 != c.y1 || c.x2 != c.x2 || c.y2 != c.y2) {
    cbox[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    a = __int_as_float(0x7fc00000);
  } else {
    cbox[q] = make_float4(c.x1, c.y1, c.x2, c.y2);
  }
  area[q] = a;
  // bit 1 = "special": anything for which the 4-compare disjointness test below would not be the exact
  // answer (non-finite corner or area, empty or inverted extent, negative area).  Bit 0 = suppressed.
  const bool ordinary = fabsf(c.x1) < INFINITY && fabsf(c.y1) < INFINITY && fabsf(c.x2) < INFINITY &&
                        fabsf(c.y2) < INFINITY && c.x2 > c.x1 && c.y2 > c.y1 && a >= 0.f && a < INFINITY;
  suppressed[q] = ordinary ? 0 : 2;
  if (q == 0 || key2[q] != key2[q - 1]) seg_starts[atomicAdd(nseg, 1)] = q;
}

// `!(iou < thr)` of utils.py:175-179 for boxes prepared by k_gather_segments (no NaN corners; NaN
// folded into the area).  Bit-identical to yb_iou, but: plain fmaxf/fminf, and pairs that do not
// overlap skip the multiply/divide -- their intersection is exactly 0, so iou is +-0 (kept, as
// 0 < thr) unless the denominator is 0 or NaN (iou NaN => suppressed).
__device__ __forceinline__ bool nms_suppresses(const float4 e, const float ae, const float4 l,
                                               const float al, const float thr, const bool thr_pos) {
  const float xA = fmaxf(e.x, l.x), yA = fmaxf(e.y, l.y);
  const float xB = fminf(e.z, l.z), yB = fminf(e.w, l.w);
  const float dw = __fsub_rn(xB, xA), dh = __fsub_rn(yB, yA);
  const float s = __fadd_rn(ae, al);
  if (thr_pos && (dw <= 0.f || dh <= 0.f) && fabsf(dw) < INFINITY && fabsf(dh) < INFINITY) {
    const float d = __fadd_rn(s, 1e-6f);  // union == s because the intersection is exactly 0
    return !(d == d && d != 0.f);
  }
  const float inter = __fmul_rn(yb_clamp0(dw), yb_clamp0(dh));
  const float iou = __fdiv_rn(inter, __fadd_rn(__fsub_rn(s, inter), 1e-6f));
  return !(iou < thr);
}

// Hot-loop form: two ordinary boxes (see k_gather_segments) whose extents are disjoint have an exactly
// zero intersection and a positive finite denominator, so iou == 0 < thr: four compares settle it.
__device__ __forceinline__ bool nms_pair(const float4 e, const float ae, const float4 l, const float al,
                                         const bool general, const float thr, const bool thr_pos) {
  if (!general) {
    if (e.z <= l.x || l.z <= e.x || e.w <= l.y || l.w <= e.y) return false;
    const float iw = __fsub_rn(fminf(e.z, l.z), fmaxf(e.x, l.x)), ih = __fsub_rn(fminf(e.w, l.w), fmaxf(e.y, l.y));
    const float inter = __fmul_rn(iw, ih);
    const float iou = __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(ae, al), inter), 1e-6f));
    return !(iou < thr);
  }
  return nms_suppresses(e, ae, l, al, thr, thr_pos);
}

// One CTA per (image, class) segment, boxes in descending-score order, walked in chunks of 32:
//   (0) the chunk's boxes go to shared memory; the warp that owns them publishes their alive bits;
//   (a) 32x32 suppression bit matrix, one pair per thread, plus per-chunk SPATIAL BIN MASKS: for each of
//       32 x-bins and 32 y-bins of [0,1), which chunk boxes overlap the bin;
//   (b) warp 0 resolves the chunk serially (the reference's greedy loop restricted to 32 boxes);
//   (c) every thread applies the chunk's survivors to the later boxes it OWNS.  Owned boxes live in
//       registers for the whole segment (<= 16 per thread, 8192 per segment; longer segments fall back to
//       global memory).  A later box is only tested against survivors that share an x-bin AND a y-bin
//       with it -- disjoint extents cannot suppress (iou == 0 < thr) -- so the usual cost per (box,
//       chunk) is a few shared-memory words instead of up to 32 IoU tests.
constexpr int NMS_QPT = 48;  // owned boxes per thread: 48 x 512 = 24 576 boxes per segment take the fast path (NT = 512)

__device__ __forceinline__ int nms_bin(float v) {  // monotone and clamped => overlapping extents share a bin
  return (int)fminf(fmaxf(v * 32.f, 0.f), 31.f);
}
__device__ __forceinline__ uint32_t nms_pack_bins(const float4 b) {
  return uint32_t(nms_bin(b.x)) | (uint32_t(nms_bin(b.z)) << 8) | (uint32_t(nms_bin(b.y)) << 16) |
         (uint32_t(nms_bin(b.w)) << 24);
}
// packed bins -> the two span-table indices of the register path (x in the low half word, y in the high one):
// (span << 5) + first bin, y offset by 128; 256 (all ones) when the box spans more than four bins on that axis
__device__ __forceinline__ uint32_t nms_tab_index(uint32_t bins) {
  const int xl = bins & 31, sx = int((bins >> 8) & 31) - xl, yl = (bins >> 16) & 31, sy = int((bins >> 24) & 31) - yl;
  const uint32_t ix = sx > 3 ? 256u : uint32_t((sx << 5) + xl);
  const uint32_t iy = sy > 3 ? 256u : uint32_t(128 + (sy << 5) + yl);
  return ix | (iy << 16);
}
// survivors of the current chunk that could overlap a box with these bins (superset)
__device__ __forceinline__ uint32_t nms_candidates(uint32_t bins, const uint32_t* s_binx, const uint32_t* s_biny) {
  const int xl = bins & 31, xh = (bins >> 8) & 31, yl = (bins >> 16) & 31, yh = (bins >> 24) & 31;
  uint32_t mx = 0, my = 0;
  if (xh - xl > 3) mx = 0xffffffffu; else for (int b = xl; b <= xh; ++b) mx |= s_binx[b];
  if (yh - yl > 3) my = 0xffffffffu; else for (int b = yl; b <= yh; ++b) my |= s_biny[b];
  return mx & my;
}

// position of the (r+1)-th set bit of mask (r < popc(mask)): five popc steps instead of __fns' software loop
__device__ __forceinline__ int nms_nth_set_bit(uint32_t mask, int r) {
  int base = 0, c;
  c = __popc(mask & 0xffffu); if (r >= c) { r -= c; base += 16; mask >>= 16; }
  c = __popc(mask & 0xffu);   if (r >= c) { r -= c; base += 8;  mask >>= 8; }
  c = __popc(mask & 0xfu);    if (r >= c) { r -= c; base += 4;  mask >>= 4; }
  c = __popc(mask & 0x3u);    if (r >= c) { r -= c; base += 2;  mask >>= 2; }
  c = int(mask & 1u);         if (r >= c) { base += 1; }
  return base;
}

// Segments of at most 32 boxes -- every (image, class) group of a trained detector, and most groups of the standalone
// sweep -- are settled by ONE WARP each, in registers: lane l holds box l, row i of the suppression matrix is one
// ballot, and only rows of boxes that are still alive are evaluated.  Longer segments are appended to big_list for
// k_nms_segments (a 512-thread CTA per tiny segment cost ~6 us of barriers and dependent loads each).
__global__ void __launch_bounds__(256)
k_nms_small(const float4* __restrict__ cbox, const float* __restrict__ area, const uint64_t* __restrict__ key2,
            const int32_t* __restrict__ val2, const int32_t* __restrict__ seg_starts, const int32_t* __restrict__ nseg_dev,
            const int32_t* __restrict__ n_dev, float thr, const uint8_t* __restrict__ flags, uint8_t* __restrict__ keep,
            const bool int_classes, int32_t* __restrict__ big_list, int32_t* __restrict__ nbig) {
  const int n = *n_dev, nseg = *nseg_dev;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool thr_pos = thr > 0.f;
  for (int sg = warp; sg < nseg; sg += nwarps) {
    const int s0 = seg_starts[sg];
    const uint64_t k = key2[s0];
    const int q = s0 + lane;
    const bool mine = q < n && key2[q] == k;          // sorted keys: the members are lanes 0 .. m-1
    const uint32_t members = __ballot_sync(0xffffffffu, mine);
    const bool more = (s0 + 32 < n) && key2[s0 + 32] == k;
    if (members == 0xffffffffu && more) {              // longer than a warp
      if (lane == 0) big_list[atomicAdd(nbig, 1)] = sg;
      continue;
    }
    if (!int_classes && uint32_t(k) == 0x7fc00000u) {  // NaN class: != is always true (utils.py:178) => all kept
      if (mine) keep[val2[q]] = 1;
      continue;
    }
    const int m = __popc(members);
    const float4 b = mine ? cbox[q] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float a = mine ? area[q] : 0.f;
    const bool gen = !thr_pos || (mine && (flags[q] & 2));
    const int out = mine ? val2[q] : 0;
    uint32_t alive = members;
    for (int i = 0; i + 1 < m; ++i) {
      if (!((alive >> i) & 1u)) continue;              // uniform: a suppressed box suppresses nothing
      const float4 bi = make_float4(__shfl_sync(0xffffffffu, b.x, i), __shfl_sync(0xffffffffu, b.y, i),
                                    __shfl_sync(0xffffffffu, b.z, i), __shfl_sync(0xffffffffu, b.w, i));
      const float ai = __shfl_sync(0xffffffffu, a, i);
      const bool gi = __shfl_sync(0xffffffffu, (int)gen, i) != 0;
      const bool sup = mine && lane > i && nms_pair(bi, ai, b, a, gi || gen, thr, thr_pos);
      alive &= ~__ballot_sync(0xffffffffu, sup);
    }
    if (mine && ((alive >> lane) & 1u)) keep[out] = 1;  // keep[] is zero-initialised
  }
}

// NT threads per segment CTA: 512 suits few long segments, 128 many short ones (4x the resident CTAs per SM: 24 KB of
// bins and a quarter of the threads each) -- the host picks by the average segment length it can expect.
template <int NT>
__global__ void __launch_bounds__(NT)
k_nms_segments(const float4* __restrict__ cbox, const float* __restrict__ area,
               const uint64_t* __restrict__ key2, const int32_t* __restrict__ val2,
               const int32_t* __restrict__ seg_starts, const int32_t* __restrict__ nseg_dev,
               const int32_t* __restrict__ n_dev, float thr, uint8_t* __restrict__ suppressed,
               uint8_t* __restrict__ keep, const bool int_classes, const int32_t* __restrict__ big_list) {
  __shared__ float4 s_cbox[32];
  __shared__ float s_carea[32];
  __shared__ uint32_t s_row[32], s_binx[32], s_biny[32];
  // s_tab[axis][span][b] = OR of the bin masks b .. b + span: the survivors a box with bin range [b, b + span] can meet on
  // that axis, in ONE shared-memory read (the per-bin loop of nms_candidates made step (c) a chain of dependent reads)
  __shared__ uint32_t s_tab[257];   // [256] = all ones: the entry of a box that spans more than four bins on an axis
  __shared__ uint32_t s_alive, s_cgen, s_kept, s_win[4];
  extern __shared__ uint32_t s_bins[];  // [NMS_QPT][NMS_THREADS] span-table indices of the owned boxes (nms_tab_index), 96 KB (dynamic)
  __shared__ int s_cpos[32];
  __shared__ int s_end, s_fnew;
  constexpr int NMS_THREADS = NT, NMS_WARPS = NT / 32, NMS_REG_CAP = NT * NMS_QPT;
  const bool thr_pos = thr > 0.f;
  const int n = *n_dev, nseg = *nseg_dev;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int si = blockIdx.x; si < nseg; si += gridDim.x) {   // nseg_dev = number of entries of big_list (k_nms_small)
    const int s0 = seg_starts[big_list[si]];
    const uint64_t k = key2[s0];
    if (tid == 0) {  // upper bound of k in the sorted keys
      int lo = s0, hi = n;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (key2[mid] == k) lo = mid; else hi = mid;
      }
      s_end = hi;
    }
    __syncthreads();
    const int s1 = s_end;
    if (!int_classes && uint32_t(k) == 0x7fc00000u) {  // NaN class: != is always true (utils.py:178) => all kept
      for (int q = s0 + tid; q < s1; q += NMS_THREADS) keep[val2[q]] = 1;
      __syncthreads();
      continue;
    }
    const int m = s1 - s0;
    const bool regpath = m <= NMS_REG_CAP;
    // per owned box: packed spatial bins in shared memory (slot-major => conflict free), dead / general bits in
    // two registers.  Coordinates are re-read from global memory only for the rare box that shares bins with a
    // survivor, so the per-round loop can walk just the live later slots (a dynamic index, hence not registers).
    uint64_t supp = 0, gen = thr_pos ? 0ull : ~0ull;
    if (regpath) {
      for (int j = 0; j < NMS_QPT; ++j) {
        const int q = s0 + j * NMS_THREADS + tid;
        if (q < s1) {
          if (suppressed[q] & 2) gen |= 1ull << j;
          s_bins[j * NMS_THREADS + tid] = nms_tab_index(nms_pack_bins(cbox[q]));
        } else {
          supp |= 1ull << j;
        }
      }
    }
    if (regpath) {
      // ---- register path: chunks are the next 32 boxes that are still ALIVE (dead boxes are skipped, which
      // cuts the number of serial rounds from m/32 to ~(#survivors)/32), taken from a 128-position window.
      int f = 0;  // frontier: first position (relative to s0) not yet consumed
#ifdef YB_NMS_PROFILE
      const bool prof = tid == 0 && m >= 1024;
      long long pt[6] = {0, 0, 0, 0, 0, 0}, t_last = clock64();
      unsigned long long rounds = 0;
#endif
      while (f < m) {
        const int wb = f >> 5;
        // every warp publishes the blocks it owns that fall into the window
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int blk = wb + d;
          if ((blk % NMS_WARPS) == warp) {
            uint32_t al = 0;
            if (blk * 32 < m) al = __ballot_sync(0xffffffffu, ((supp >> (blk / NMS_WARPS)) & 1ull) == 0);
            if (d == 0) al &= ~((1u << (f & 31)) - 1u);
            if (lane == 0) s_win[d] = al;
          }
        }
        __syncthreads();
        NMS_T(0);
        // (0b) warp 0 selects the members -- lane l takes the (l+1)-th alive position -- and publishes the new frontier;
        // the other warps only need the member count (the selection cost ~80 instructions in each of the 16 warps, and
        // the kernel is bound by instruction issue: scripts/nms_round_profile.py)
        const uint32_t w0 = s_win[0], w1 = s_win[1], w2 = s_win[2], w3 = s_win[3];
        const int c0n = __popc(w0), c1n = __popc(w1), c2n = __popc(w2), c3n = __popc(w3);
        const int total = c0n + c1n + c2n + c3n;
        if (total == 0) {  // uniform
          f = (wb + 4) << 5;
          __syncthreads();  // s_win is rewritten by the next round
          continue;
        }
        const int nmem = total < 32 ? total : 32;
        int mypos = -1;
        if (warp == 0) {
          if (lane < nmem) {
            int r = lane, d = 0;
            uint32_t wsel = w0;
            if (r >= c0n) { r -= c0n; d = 1; wsel = w1;
              if (r >= c1n) { r -= c1n; d = 2; wsel = w2;
                if (r >= c2n) { r -= c2n; d = 3; wsel = w3; } } }
            mypos = ((wb + d) << 5) + nms_nth_set_bit(wsel, r);
          }
          const int lastpos = __shfl_sync(0xffffffffu, mypos, nmem - 1);
          if (lane == 0) s_fnew = (total <= 32) ? ((wb + 4) << 5) : (lastpos + 1);
        }
        // (0c) member boxes -> shared memory
        int my_out = 0;   // warp 0: this member's row in the reference's output order (for step (b): no dependent load there)
        if (warp == 0) {
          const bool valid = lane < nmem;
          const int qi = s0 + (valid ? mypos : 0);
          my_out = val2[qi];
          s_cbox[lane] = valid ? cbox[qi] : make_float4(0.f, 0.f, 0.f, 0.f);
          s_carea[lane] = valid ? area[qi] : 0.f;
          const int st = valid ? suppressed[qi] : 0;
          const uint32_t g = __ballot_sync(0xffffffffu, valid && (!thr_pos || (st & 2)));
          if (lane == 0) s_cgen = g;
          s_cpos[lane] = valid ? qi : -1;
        }
        __syncthreads();
        const int f_new = s_fnew;
        NMS_T(1);
        // (a) pair matrix (rows warp and warp+16) and spatial bin masks (bins warp and warp+16)
        {
          const float4 bj = s_cbox[lane];
          const float aj = s_carea[lane];
          const uint32_t cgen = s_cgen;
          const bool vj = lane < nmem;
          const uint32_t pb = nms_pack_bins(bj);
          const int xl = pb & 31, xh = (pb >> 8) & 31, yl = (pb >> 16) & 31, yh = (pb >> 24) & 31;
#pragma unroll
          for (int i = warp; i < 32; i += NMS_WARPS) {
            bool sup = false;
            if (lane > i && vj)
              sup = nms_pair(s_cbox[i], s_carea[i], bj, aj, ((cgen >> i) | (cgen >> lane)) & 1u, thr, thr_pos);
            const uint32_t row = __ballot_sync(0xffffffffu, sup);
            const uint32_t mxb = __ballot_sync(0xffffffffu, vj && xl <= i && i <= xh);
            const uint32_t myb = __ballot_sync(0xffffffffu, vj && yl <= i && i <= yh);
            if (lane == 0) { s_row[i] = row; s_binx[i] = mxb; s_biny[i] = myb; }
          }
        }
        __syncthreads();
        NMS_T(2);
        // (b) warp 0 resolves the chunk serially; rows are read up front so the dependent chain is pure ALU
        if (warp == 0) {
          uint32_t alive = nmem == 32 ? 0xffffffffu : ((1u << nmem) - 1u);
          uint32_t kept = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t ri = s_row[i];
            if ((alive >> i) & 1u) { kept |= 1u << i; alive &= ~ri; }
          }
          if (lane == 0) s_kept = kept;
          if ((kept >> lane) & 1u) keep[my_out] = 1;  // keep[] is zero-initialised
        } else {
          // meanwhile the other warps build the span tables from the bin masks of step (a)
          if (tid == 32) s_tab[256] = 0xffffffffu;
          for (int e = tid - 32; e < 256; e += NMS_THREADS - 32) {
            const int bsel = e & 31, span = (e >> 5) & 3;
            const uint32_t* src = (e >> 7) ? s_biny : s_binx;
            uint32_t o = src[bsel];
            if (span >= 1 && bsel + 1 < 32) o |= src[bsel + 1];
            if (span >= 2 && bsel + 2 < 32) o |= src[bsel + 2];
            if (span >= 3 && bsel + 3 < 32) o |= src[bsel + 3];
            s_tab[e] = o;
          }
        }
        __syncthreads();
        NMS_T(3);
        // (c) the chunk's survivors knock out the later boxes this thread owns: only live slots are visited
        const uint32_t kept = s_kept;
        const uint32_t kgen = kept & s_cgen;
        {
          // slots whose position j*NT + tid is >= f_new: all slots above jf, plus slot jf itself when tid >= rf
          const int jf = f_new / NMS_THREADS, rf = f_new - jf * NMS_THREADS;
          constexpr uint64_t ALL = (NMS_QPT == 64) ? ~0ull : ((1ull << NMS_QPT) - 1ull);
          uint64_t later = (jf >= NMS_QPT) ? 0ull : (ALL << jf) & ALL;
          if (jf < NMS_QPT && tid < rf) later &= ~(1ull << jf);
          // survivors of this chunk that could overlap owned box j (superset): general boxes meet every survivor
          auto cand_of = [&](int j) -> uint32_t {
            if ((gen >> j) & 1ull) return kept;
            const uint32_t w = s_bins[j * NMS_THREADS + tid];   // the box's two table indices (nms_tab_index)
            return (s_tab[w & 0xffffu] & s_tab[w >> 16] & kept) | kgen;
          };
          // pass 1: which live slots have a candidate at all (shared memory only).  The slot masks are walked as two
          // 32-bit words: this loop runs once per (live owned box, round) and 64-bit ffs / shift / clear doubled its
          // bookkeeping (segments of <= 32 * NT boxes never enter the second word)
          uint64_t hit = 0;
          {
            const uint64_t act = later & ~supp;
#pragma unroll
            for (int hw = 0; hw < 2; ++hw) {
              uint32_t a32 = uint32_t(act >> (32 * hw));
              const uint32_t g32 = uint32_t(gen >> (32 * hw));
              uint32_t h32 = 0;
              const uint32_t* bins = s_bins + (32 * hw) * NMS_THREADS + tid;
              while (a32) {
                const int jl = __ffs(a32) - 1;
                a32 &= a32 - 1;
                uint32_t cand;
                if ((g32 >> jl) & 1u) {
                  cand = kept;
                } else {
                  const uint32_t w = bins[jl * NMS_THREADS];
                  cand = (s_tab[w & 0xffffu] & s_tab[w >> 16] & kept) | kgen;
                }
                if (cand) h32 |= 1u << jl;
              }
              hit |= uint64_t(h32) << (32 * hw);
            }
          }
          // pass 2: the few that do fetch their coordinates -- four loads in flight instead of one L2 round trip per slot
          while (hit) {
            int js[4];
            float4 bb[4];
            float aa[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              js[u] = -1;
              if (hit) { js[u] = __ffsll((long long)hit) - 1; hit &= hit - 1; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (js[u] >= 0) {
                const int q = s0 + js[u] * NMS_THREADS + tid;
                bb[u] = cbox[q];
                aa[u] = area[q];
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (js[u] >= 0) {
                const bool qgen = (gen >> js[u]) & 1ull;
                uint32_t cand = cand_of(js[u]);
                while (cand) {
                  const int t = __ffs(cand) - 1;
                  cand &= cand - 1;
                  if (nms_pair(s_cbox[t], s_carea[t], bb[u], aa[u], qgen || ((kgen >> t) & 1u), thr, thr_pos)) {
                    supp |= 1ull << js[u];
                    break;
                  }
                }
              }
            }
          }
        }
        f = f_new;
#ifdef YB_NMS_PROFILE
        NMS_T(4);
        ++rounds;
#endif
      }
#ifdef YB_NMS_PROFILE
      if (prof) {
        for (int i = 0; i < 5; ++i) atomicAdd(&g_nms_prof[i], (unsigned long long)pt[i]);
        atomicAdd(&g_nms_prof[5], rounds);
        atomicAdd(&g_nms_prof[6], 1ull);
      }
#endif
      __syncthreads();
      continue;
    }
    // ---- long segments (> 8192 boxes): consecutive 32-box chunks, state in global memory
    for (int ci = 0; ci * 32 < m; ++ci) {
      const int c0 = s0 + ci * 32;
      if (warp == 0) {
        const int qi = c0 + lane;
        const bool valid = qi < s1;
        s_cbox[lane] = valid ? cbox[qi] : make_float4(0.f, 0.f, 0.f, 0.f);
        s_carea[lane] = valid ? area[qi] : 0.f;
        const int st = valid ? suppressed[qi] : 1;
        const uint32_t g = __ballot_sync(0xffffffffu, valid && (!thr_pos || (st & 2)));
        const uint32_t al = __ballot_sync(0xffffffffu, (st & 1) == 0);
        if (lane == 0) { s_cgen = g; s_alive = al; }
      }
      __syncthreads();
      {
        const float4 bj = s_cbox[lane];
        const float aj = s_carea[lane];
        const uint32_t cgen = s_cgen;
        const bool vj = c0 + lane < s1;
        const uint32_t pb = nms_pack_bins(bj);
        const int xl = pb & 31, xh = (pb >> 8) & 31, yl = (pb >> 16) & 31, yh = (pb >> 24) & 31;
#pragma unroll
        for (int i = warp; i < 32; i += NMS_WARPS) {
          bool sup = false;
          if (lane > i && vj)
            sup = nms_pair(s_cbox[i], s_carea[i], bj, aj, ((cgen >> i) | (cgen >> lane)) & 1u, thr, thr_pos);
          const uint32_t row = __ballot_sync(0xffffffffu, sup);
          const uint32_t mxb = __ballot_sync(0xffffffffu, vj && xl <= i && i <= xh);
          const uint32_t myb = __ballot_sync(0xffffffffu, vj && yl <= i && i <= yh);
          if (lane == 0) { s_row[i] = row; s_binx[i] = mxb; s_biny[i] = myb; }
        }
      }
      __syncthreads();
      if (warp == 0) {
        const int qi = c0 + lane;
        const bool valid = qi < s1;
        uint32_t alive = s_alive & __ballot_sync(0xffffffffu, valid);
        uint32_t kept = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const uint32_t ri = s_row[i];
          if ((alive >> i) & 1u) { kept |= 1u << i; alive &= ~ri; }
        }
        if (valid && ((kept >> lane) & 1u)) keep[val2[qi]] = 1;
        if (lane == 0) s_kept = kept;
      }
      __syncthreads();
      const uint32_t kept = s_kept;
      if (kept != 0) {
        const uint32_t kgen = kept & s_cgen;
        for (int q = c0 + 32 + tid; q < s1; q += NMS_THREADS) {
          const int state = suppressed[q];
          if (state & 1) continue;
          const float4 b = cbox[q];
          const bool qgen = !thr_pos || (state & 2);
          uint32_t cand = qgen ? kept : ((nms_candidates(nms_pack_bins(b), s_binx, s_biny) & kept) | kgen);
          if (!cand) continue;
          const float al = area[q];
          while (cand) {
            const int t = __ffs(cand) - 1;
            cand &= cand - 1;
            if (nms_pair(s_cbox[t], s_carea[t], b, al, qgen || ((kgen >> t) & 1u), thr, thr_pos)) {
              suppressed[q] = state | 1;
              break;
            }
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(COMPACT_THREADS)
k_keep_count(const uint8_t* __restrict__ keep, const int32_t* __restrict__ n_dev,
             int32_t* __restrict__ tile_cnt) {
  const int n = *n_dev;
  const int base = blockIdx.x * COMPACT_TILE + threadIdx.x * COMPACT_ITEMS;
  int c = 0;
#pragma unroll
  for (int k = 0; k < COMPACT_ITEMS; ++k) c += (base + k < n && keep[base + k]) ? 1 : 0;
  int tot;
  block_exclusive_scan<COMPACT_THREADS>(c, &tot);
  if (threadIdx.x == 0) {
    tile_cnt[blockIdx.x] = tot;
    if (blockIdx.x == 0) tile_cnt[gridDim.x] = 0;  // sentinel: after the scan it holds the grand total
  }
}

__global__ void __launch_bounds__(COMPACT_THREADS)
k_keep_write(const uint8_t* __restrict__ keep, const uint64_t* __restrict__ key1,
             const int32_t* __restrict__ idx1, const int32_t* __restrict__ n_dev,
             const int32_t* __restrict__ tile_off, int32_t* __restrict__ keep_idx) {
  const int n = *n_dev;
  const int base = blockIdx.x * COMPACT_TILE + threadIdx.x * COMPACT_ITEMS;
  unsigned flags = 0;
  int c = 0;
#pragma unroll
  for (int k = 0; k < COMPACT_ITEMS; ++k)
    if (base + k < n && keep[base + k]) { flags |= 1u << k; ++c; }
  int pos = tile_off[blockIdx.x] + block_exclusive_scan<COMPACT_THREADS>(c, nullptr);
#pragma unroll
  for (int k = 0; k < COMPACT_ITEMS; ++k) {
    if (flags & (1u << k)) {
      keep_idx[pos++] = idx1[base + k];
    }
  }
}

// keep_off[b] = number of survivors in images < b = survivors at positions p < (first p of image >= b):
// one warp per entry; binary search on the image-sorted keys, then tile prefix + partial-tile count.
__global__ void k_keep_offsets(const uint8_t* __restrict__ keep, const uint64_t* __restrict__ key1,
                               const int32_t* __restrict__ n_dev, const int32_t* __restrict__ tile_off, int batch,
                               int32_t* __restrict__ keep_off) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b > batch) return;
  const int n = *n_dev;
  int lo = 0, hi = n;  // first p with image(p) >= b
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (int(key1[mid] >> 32) < b) lo = mid + 1; else hi = mid;
  }
  const int tile = lo / COMPACT_TILE;
  int c = 0;
  for (int p = tile * COMPACT_TILE + lane; p < lo; p += 32) c += keep[p] ? 1 : 0;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if (lane == 0) keep_off[b] = tile_off[tile] + c;
}

struct NmsWs {
  int32_t* tile_cnt;
  int32_t* scalars;  // [0] n_valid, [1] nseg
  uint64_t* key1;
  int32_t* val1;
  uint64_t* key2;
  int32_t* val2;
  int32_t* idx1;
  float4* cbox;
  float* area;
  uint8_t* suppressed;
  uint8_t* keep;
  int32_t* seg_starts;
  int32_t* big_list;
  SortBuffers sb;
  int ctiles;
};

void nms_carve(WsCarver& ws, int total, NmsWs* w) {
  const int t = total > 0 ? total : 1;
  w->ctiles = yb_cdiv(t, COMPACT_TILE);
  w->tile_cnt = ws.take<int32_t>(w->ctiles + 1);
  w->scalars = ws.take<int32_t>(8);
  w->keep = ws.take<uint8_t>(t);     // directly behind the scalars: one memset clears both
  w->key1 = ws.take<uint64_t>(t);
  w->val1 = ws.take<int32_t>(t);
  w->key2 = ws.take<uint64_t>(t);
  w->val2 = ws.take<int32_t>(t);
  w->idx1 = ws.take<int32_t>(t);
  w->cbox = ws.take<float4>(t);
  w->area = ws.take<float>(t);
  w->suppressed = ws.take<uint8_t>(t);
  w->seg_starts = ws.take<int32_t>(t);
  w->big_list = ws.take<int32_t>(t);
  sort_carve(ws, t, &w->sb);
}

}  // namespace

extern "C" size_t yolo_nms_workspace_bytes(int total, int batch) {
  (void)batch;
  WsCarver ws(nullptr);
  NmsWs w;
  nms_carve(ws, total, &w);
  return ws.bytes();
}

extern "C" int yolo_nms(const float* boxes, const int32_t* img_offsets, int batch, int total,
                        float iou_thr, double obj_thr, int box_format, int class_bits, int32_t* keep_idx,
                        int32_t* keep_off, void* workspace, size_t workspace_bytes,
                        yb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  YB_REQUIRE(batch >= 1 && total >= 0, "yolo_nms: batch must be >= 1 and total >= 0");
  YB_REQUIRE(box_format == YB_BOX_CENTER || box_format == YB_BOX_CORNERS, "yolo_nms: bad box_format");
  YB_REQUIRE(keep_off != nullptr, "yolo_nms: keep_off is null");
  YB_REQUIRE(class_bits == 0 || class_bits == 8 || class_bits == 16, "yolo_nms: class_bits must be 0, 8 or 16");
  if (workspace_bytes < yolo_nms_workspace_bytes(total, batch)) {
    yb_set_error("yolo_nms: workspace too small (%zu < %zu)", workspace_bytes,
                 yolo_nms_workspace_bytes(total, batch));
    return YB_ERR_WORKSPACE;
  }
  if (total == 0) {   // otherwise k_keep_offsets writes every entry
    YB_CHECK_CUDA(cudaMemsetAsync(keep_off, 0, sizeof(int32_t) * (batch + 1), stream));
    return YB_OK;
  }
  YB_REQUIRE(boxes && img_offsets && keep_idx && workspace, "yolo_nms: null pointer");
  WsCarver ws(workspace);
  NmsWs w;
  nms_carve(ws, total, &w);
  int32_t* n_valid = w.scalars;
  int32_t* nseg = w.scalars + 1;
  // counters + keep flags (the NMS kernels only write the 1s) in one memset: keep sits right behind the scalars
  YB_CHECK_CUDA(cudaMemsetAsync(w.scalars, 0, size_t(reinterpret_cast<uint8_t*>(w.keep) - reinterpret_cast<uint8_t*>(w.scalars)) + size_t(total), stream));

  // K4: ordered threshold compaction -> (key, row index) pairs
  k_thr_count<<<w.ctiles, COMPACT_THREADS, 0, stream>>>(boxes, total, obj_thr, w.tile_cnt);
  YB_CHECK_LAUNCH();
  int rc = exclusive_scan_small(w.tile_cnt, w.ctiles, n_valid, stream);
  if (rc) return rc;
  k_thr_write<<<w.ctiles, COMPACT_THREADS, 0, stream>>>(boxes, img_offsets, batch, total, obj_thr,
                                                        w.tile_cnt, w.key1, w.val1);
  YB_CHECK_LAUNCH();

  // K5: (image, score desc, position) then regroup by (image, class)
  int img_bits = 0;
  while ((1 << img_bits) < batch) ++img_bits;
  const int img_hi = 32 + ((img_bits + 7) / 8) * 8;
  // one call per sort; the results are followed to whichever buffer the last pass wrote (no copy-back launches).  Sort
  // #2 ping-pongs between its own buffers and the buffer pair sort #1 did NOT end in (key1's order is needed again below).
  uint64_t* k1 = nullptr; int32_t* v1 = nullptr;
  rc = radix_sort_pairs(w.key1, w.val1, n_valid, total, 0, img_hi, w.sb, stream, &k1, &v1);
  if (rc) return rc;
  const int eb = 256, eg = yb_cdiv(total, eb);
  k_class_keys<<<eg, eb, 0, stream>>>(boxes, k1, v1, n_valid, w.key2, w.val2, w.idx1, class_bits, w.scalars + 2);
  YB_CHECK_LAUNCH();
  SortBuffers sb2 = w.sb;
  if (k1 != w.key1) { sb2.keys_alt = w.key1; sb2.vals_alt = w.val1; }
  uint64_t* k2 = nullptr; int32_t* v2 = nullptr;
  rc = radix_sort_pairs(w.key2, w.val2, n_valid, total, 0, class_bits > 0 ? class_bits + (img_hi - 32) : img_hi, sb2, stream, &k2, &v2);
  if (rc) return rc;
  k_gather_segments<<<eg, eb, 0, stream>>>(boxes, k2, v2, w.idx1, n_valid, box_format,
                                           w.cbox, w.area, w.suppressed, w.seg_starts, nseg);
  YB_CHECK_LAUNCH();

  // K6: greedy NMS per (image, class) segment: a warp each for segments of <= 32 boxes, a CTA each for the others
  int32_t* nbig = w.scalars + 3;
  int dev = 0, sms = 148;
  YB_CHECK_CUDA(cudaGetDevice(&dev));
  YB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // CTA size: segments are (image, class) groups; with integer class labels of <= 2^class_bits classes the average
  // segment is total / (batch * classes) boxes long at most -- short segments want many small CTAs
  // Measured on random-init YOLOv3 heads (class-skewed segments, profiles/r1_nms_cta_size.txt): 128 / 256 / 512 threads
  // give 0.99 / 0.71 / 0.61 ms at 416 (conf 0.5) and 4.6 / 3.1 / 1.9 ms at 608 (conf 0.01), 1024 threads 0.94 / 2.8 ms.
  {
    int sgrid = yb_cdiv(total, 8);   // one warp per segment, 8 warps per block; at most `total` segments
    if (sgrid > sms * 8) sgrid = sms * 8;
    k_nms_small<<<sgrid, 256, 0, stream>>>(w.cbox, w.area, k2, v2, w.seg_starts, nseg, n_valid, iou_thr, w.suppressed,
                                          w.keep, class_bits > 0, w.big_list, nbig);
    YB_CHECK_LAUNCH();
  }
  const int nt = 512;
  const size_t nms_smem = size_t(NMS_QPT) * nt * sizeof(uint32_t);
  int grid = sms * (nt == 1024 ? 1 : (nt == 512 ? 3 : (nt == 256 ? 6 : 12)));
  if (grid > total) grid = total;
#define YB_NMS_LAUNCH(NTV)                                                                                              \
  do {                                                                                                                  \
    YB_CHECK_CUDA(cudaFuncSetAttribute(k_nms_segments<NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem)); \
    k_nms_segments<NTV><<<grid, NTV, nms_smem, stream>>>(w.cbox, w.area, k2, v2, w.seg_starts, nbig, n_valid,              \
                                                       iou_thr, w.suppressed, w.keep, class_bits > 0, w.big_list);      \
  } while (0)
  if (nt == 128) YB_NMS_LAUNCH(128); else if (nt == 256) YB_NMS_LAUNCH(256); else if (nt == 1024) YB_NMS_LAUNCH(1024);
  else YB_NMS_LAUNCH(512);
#undef YB_NMS_LAUNCH
  YB_CHECK_LAUNCH();

  // survivors, in the reference's order, + per-image offsets
  k_keep_count<<<w.ctiles, COMPACT_THREADS, 0, stream>>>(w.keep, n_valid, w.tile_cnt);
  YB_CHECK_LAUNCH();
  rc = exclusive_scan_small(w.tile_cnt, w.ctiles + 1, nullptr, stream);
  if (rc) return rc;
  k_keep_write<<<w.ctiles, COMPACT_THREADS, 0, stream>>>(w.keep, k1, w.idx1, n_valid,
                                                         w.tile_cnt, keep_idx);
  YB_CHECK_LAUNCH();
  k_keep_offsets<<<yb_cdiv((batch + 1) * 32, 128), 128, 0, stream>>>(w.keep, k1, n_valid, w.tile_cnt, batch,
                                                                     keep_off);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

#ifdef YB_NMS_PROFILE
extern "C" int yolo_debug_nms_prof(unsigned long long* out8, int reset) {
  if (out8) cudaMemcpyFromSymbol(out8, g_nms_prof, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(g_nms_prof, z, sizeof(z)); }
  return 0;
}
#endif

// ---- element-wise IoU: utils.py:38-84 calc_iou / utils.py:22-36 iou_aligned --
namespace {
__global__ void k_iou(const float* __restrict__ b1, int n1, int st1, const float* __restrict__ b2,
                      int n2, int st2, int fmt, int aligned, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = b1 + size_t(n1 == 1 ? 0 : i) * st1;
  const float* q = b2 + size_t(n2 == 1 ? 0 : i) * st2;
  if (aligned) {  // utils.py:34-36
    const float inter = __fmul_rn(yb_nanmin(p[0], q[0]), yb_nanmin(p[1], q[1]));
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(p[0], p[1]), __fmul_rn(q[0], q[1])), inter);
    out[i] = __fdiv_rn(inter, uni);
    return;
  }
  const CBox a = yb_make_cbox(p[0], p[1], p[2], p[3], fmt);
  const CBox b = yb_make_cbox(q[0], q[1], q[2], q[3], fmt);
  out[i] = yb_iou(a, __fmul_rn(p[2], p[3]), b, __fmul_rn(q[2], q[3]));
}
}  // namespace

extern "C" int yolo_iou(const float* boxes1, int n1, int stride1, const float* boxes2, int n2,
                        int stride2, int box_format, int aligned, float* out, yb_stream_t stream) {
  YB_REQUIRE(n1 >= 0 && n2 >= 0, "yolo_iou: negative count");
  YB_REQUIRE(n1 == n2 || n1 == 1 || n2 == 1, "yolo_iou: shapes do not broadcast");
  const int n = (n1 == 0 || n2 == 0) ? 0 : (n1 > n2 ? n1 : n2);
  if (n == 0) return YB_OK;
  k_iou<<<yb_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(boxes1, n1, stride1, boxes2, n2, stride2,
                                                          box_format, aligned, out, n);
  YB_CHECK_LAUNCH();
  return YB_OK;
}
