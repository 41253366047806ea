// K5: stable LSD radix sort (8-bit digits) + small scan helper.  HBM/L2-bound
// integer work: keys are streamed with coalesced 8-byte loads, ranks come from
// warp match/ballot, no tensor cores involved.
#include "sort.cuh"

namespace {

__global__ void __launch_bounds__(SORT_THREADS)
k_radix_hist(const uint64_t* __restrict__ keys, const int32_t* __restrict__ n_dev, int shift,
             int32_t* __restrict__ hist, int num_tiles) {
  __shared__ int h[256];
  const int n = *n_dev;
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * SORT_TILE;
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    const int i = base + r * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[threadIdx.x * num_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(1024)
k_exclusive_scan(int32_t* __restrict__ data, int n, int32_t* __restrict__ total_out) {
  const int per = (n + 1023) / 1024;
  const int lo = min((int)threadIdx.x * per, n), hi = min(lo + per, n);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += data[i];
  int total;
  int off = block_exclusive_scan<1024>(s, &total);
  for (int i = lo; i < hi; ++i) {
    const int v = data[i];
    data[i] = off;
    off += v;
  }
  if (threadIdx.x == 0 && total_out) *total_out = total;
}

// One CTA per digit: exclusive scan of that digit's per-tile counts (contiguous, coalesced) and the
// digit's total.  Replaces a single-CTA scan over all 256 x tiles entries.
__global__ void __launch_bounds__(256)
k_radix_scan_digits(int32_t* __restrict__ hist, int num_tiles, int32_t* __restrict__ digit_total) {
  int32_t* row = hist + size_t(blockIdx.x) * num_tiles;
  int carry = 0;
  for (int base = 0; base < num_tiles; base += 256) {
    const int i = base + threadIdx.x;
    const int v = i < num_tiles ? row[i] : 0;
    int tot;
    const int ex = block_exclusive_scan<256>(v, &tot);
    if (i < num_tiles) row[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = carry;
}

// Stable scatter: items are taken in index order (round-major, then thread id);
// within a round the rank of an item among equal digits is
//   [items of lower warps] + [lower lanes of the same warp]  (match_any ballot).
__global__ void __launch_bounds__(SORT_THREADS)
k_radix_scatter(const uint64_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                uint64_t* __restrict__ keys_out, int32_t* __restrict__ vals_out,
                const int32_t* __restrict__ n_dev, int shift, const int32_t* __restrict__ hist,
                const int32_t* __restrict__ digit_total, int num_tiles) {
  __shared__ int digit_base[256];
  __shared__ int warp_cnt[SORT_THREADS / 32][256];
  const int n = *n_dev;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base = blockIdx.x * SORT_TILE;
  if (base >= n) return;  // uniform: the whole tile is past the live count
  {
    const int doff = block_exclusive_scan<SORT_THREADS>(digit_total[tid], nullptr);  // keys with a smaller digit
    digit_base[tid] = doff + hist[tid * num_tiles + blockIdx.x];                      // + same digit, earlier tiles
  }
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    const int round_base = base + r * SORT_THREADS;
    if (round_base >= n) break;  // uniform
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; ++w) warp_cnt[w][tid] = 0;
    __syncthreads();
    const int i = round_base + tid;
    const bool valid = i < n;
    const uint64_t key = valid ? keys_in[i] : 0ull;
    const unsigned d = valid ? (unsigned)((key >> shift) & 255u) : 256u;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) warp_cnt[warp][d] = __popc(peers);
    __syncthreads();
    {
      int off = digit_base[tid];
#pragma unroll
      for (int w = 0; w < SORT_THREADS / 32; ++w) {
        const int c = warp_cnt[w][tid];
        warp_cnt[w][tid] = off;
        off += c;
      }
      digit_base[tid] = off;
    }
    __syncthreads();
    if (valid) {
      const int pos = warp_cnt[warp][d] + rank;
      keys_out[pos] = key;
      vals_out[pos] = vals_in[i];
    }
    __syncthreads();
  }
}

}  // namespace

int exclusive_scan_small(int32_t* data, int n, int32_t* total_out, cudaStream_t stream) {
  k_exclusive_scan<<<1, 1024, 0, stream>>>(data, n, total_out);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

int radix_sort_pairs(uint64_t* keys, int32_t* vals, const int32_t* n_dev, int max_n, int bit_lo,
                     int bit_hi, const SortBuffers& sb, cudaStream_t stream, uint64_t** keys_res, int32_t** vals_res) {
  if (keys_res) *keys_res = keys;
  if (vals_res) *vals_res = vals;
  if (max_n <= 0 || bit_hi <= bit_lo) return YB_OK;
  YB_REQUIRE(bit_lo % 8 == 0 && bit_hi % 8 == 0 && bit_hi <= 64, "radix_sort_pairs: bad bit range");
  uint64_t* kin = keys;
  int32_t* vin = vals;
  uint64_t* kout = sb.keys_alt;
  int32_t* vout = sb.vals_alt;
  const int tiles = sb.num_tiles;
  for (int shift = bit_lo; shift < bit_hi; shift += 8) {
    k_radix_hist<<<tiles, SORT_THREADS, 0, stream>>>(kin, n_dev, shift, sb.hist, tiles);
    YB_CHECK_LAUNCH();
    k_radix_scan_digits<<<256, 256, 0, stream>>>(sb.hist, tiles, sb.digit_total);
    YB_CHECK_LAUNCH();
    k_radix_scatter<<<tiles, SORT_THREADS, 0, stream>>>(kin, vin, kout, vout, n_dev, shift, sb.hist,
                                                        sb.digit_total, tiles);
    YB_CHECK_LAUNCH();
    uint64_t* tk = kin; kin = kout; kout = tk;
    int32_t* tv = vin; vin = vout; vout = tv;
  }
  if (keys_res && vals_res) {   // the caller follows the result to whichever buffer the last pass wrote: no copy-back
    *keys_res = kin;
    *vals_res = vin;
  } else if (kin != keys) {
    YB_CHECK_CUDA(cudaMemcpyAsync(keys, kin, size_t(max_n) * sizeof(uint64_t),
                                  cudaMemcpyDeviceToDevice, stream));
    YB_CHECK_CUDA(cudaMemcpyAsync(vals, vin, size_t(max_n) * sizeof(int32_t),
                                  cudaMemcpyDeviceToDevice, stream));
  }
  return YB_OK;
}

extern "C" size_t yolo_sort_workspace_bytes(int max_n) {
  WsCarver ws(nullptr);
  SortBuffers sb;
  sort_carve(ws, max_n, &sb);
  return ws.bytes();
}

extern "C" int yolo_sort_pairs(uint64_t* keys, int32_t* vals, const int32_t* n_dev, int max_n,
                               int end_bit, void* workspace, size_t workspace_bytes,
                               yb_stream_t stream) {
  YB_REQUIRE(max_n >= 0 && end_bit > 0 && end_bit <= 64 && end_bit % 8 == 0,
             "yolo_sort_pairs: bad arguments");
  if (workspace_bytes < yolo_sort_workspace_bytes(max_n)) {
    yb_set_error("yolo_sort_pairs: workspace too small");
    return YB_ERR_WORKSPACE;
  }
  WsCarver ws(workspace);
  SortBuffers sb;
  sort_carve(ws, max_n, &sb);
  return radix_sort_pairs(keys, vals, n_dev, max_n, 0, end_bit, sb, (cudaStream_t)stream);
}
