// K5 building blocks: device-wide stable LSD radix sort of (u64 key, i32 value)
// pairs and an order-preserving stream compaction.  Both take their live
// element count from device memory so that a whole decode->NMS pipeline runs
// without a host round trip and can be captured in a CUDA graph.
#pragma once
#include "common.cuh"

constexpr int SORT_THREADS = 256;
constexpr int SORT_ROUNDS = 4;
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;  // 1024 keys per CTA

struct SortBuffers {  // carved from the caller's workspace
  uint64_t* keys_alt;
  int32_t* vals_alt;
  int32_t* hist;         // [256][num_tiles], digit-major
  int32_t* digit_total;  // [256]
  int num_tiles;
};
inline void sort_carve(WsCarver& ws, int max_n, SortBuffers* sb) {
  sb->num_tiles = yb_cdiv(max_n > 0 ? max_n : 1, SORT_TILE);
  sb->keys_alt = ws.take<uint64_t>(max_n);
  sb->vals_alt = ws.take<int32_t>(max_n);
  sb->hist = ws.take<int32_t>(size_t(256) * sb->num_tiles);
  sb->digit_total = ws.take<int32_t>(256);
}

// Sorts ascending on key bits [bit_lo, bit_hi) (multiples of 8).  With keys_res / vals_res null the result is in
// keys/vals on return (a copy pass is appended for odd pass counts); otherwise they receive the buffers that hold
// it (keys/vals or sb.keys_alt/vals_alt) and nothing is copied.
int radix_sort_pairs(uint64_t* keys, int32_t* vals, const int32_t* n_dev, int max_n, int bit_lo,
                     int bit_hi, const SortBuffers& sb, cudaStream_t stream, uint64_t** keys_res = nullptr,
                     int32_t** vals_res = nullptr);

// Order-preserving compaction support: per-tile counts -> exclusive offsets.
constexpr int COMPACT_THREADS = 256;
constexpr int COMPACT_ITEMS = 8;
constexpr int COMPACT_TILE = COMPACT_THREADS * COMPACT_ITEMS;  // 2048
// In-place exclusive scan of `n` ints by one CTA; total written to *total_out.
int exclusive_scan_small(int32_t* data, int n, int32_t* total_out, cudaStream_t stream);
