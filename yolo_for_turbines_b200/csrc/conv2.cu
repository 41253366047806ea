// K1/K2 v2: persistent fused convolution.  Same math and boundary as conv.cu (see its header for
// the reference call sites); what changes is how the SM is kept busy:
//
//   * persistent CTAs (one per SM, or one CTA PAIR per two SMs) loop over output tiles, so barrier
//     init, tensor-map prefetch and the TMEM allocation are paid once per launch, not per tile;
//   * the fp32 accumulator is double buffered in TMEM (2 x BLOCK_N columns): the epilogue of tile i
//     overlaps the TMA/MMA main loop of tile i+1;
//   * the epilogue stages bf16 output boxes (128 rows x 64 channels, 128B-swizzled) in shared memory
//     and writes them with TMA bulk stores -- full 128-byte lines instead of 16-byte fragments per
//     thread -- and the residual operand comes in through the same boxes with TMA loads;
//   * NCTA == 2: tcgen05.mma.cta_group::2 pairs two SMs on one 256 x BLOCK_N tile.  Each CTA loads its
//     own 128 rows of A and HALF of the weight tile, so per-CTA L2->smem traffic per FLOP halves.
//
// Warp roles (192 threads): 0 = TMA producer, 1 = TMEM alloc + MMA issuer (leader CTA only issues),
// 2..5 = epilogue (TMEM lane quadrant = warp % 4).
#include <new>
#include <string.h>
#include <type_traits>

#include "conv_plan.cuh"
#include "conv_ptx.cuh"

namespace {

using namespace convptx;

constexpr int CONV2_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 = two epilogue groups
constexpr int WSLOTS = 2;           // per-warp ring of 32-row output/residual boxes in shared memory
constexpr int EPI_WARPS = 8;
enum { EPI_BF16 = 0, EPI_BF16_RES = 1, EPI_F32 = 2, EPI_DIRECT = 3 };   // epilogue store / residual modes
constexpr int SCRATCH_BYTES = 256;  // per epilogue warp: scale[32] | bias[32] of the 32 columns being processed

template <int BLOCK_N, int KC, int NCTA>
struct Cfg {
  static constexpr int ROW_BYTES = KC * 2;
  static constexpr uint32_t A_BYTES = BLOCK_M * ROW_BYTES;
  static constexpr int B_ROWS = BLOCK_N / NCTA;
  static constexpr uint32_t B_BYTES = B_ROWS * ROW_BYTES;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // 64-channel boxes (128-byte rows).  Measured alternative: 32-channel boxes free 32 KB for a 6th stage of the
  // A/B ring but double the per-box epilogue overhead -- net slower on the N=256, short-K layers.
  static constexpr int BOXC = BLOCK_N < 64 ? BLOCK_N : 64;   // channels per epilogue box
  static constexpr int BOX_ROW_BYTES = BOXC * 2;             // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
  static constexpr uint32_t WBOX_BYTES = 32 * BOX_ROW_BYTES; // one warp's 32 rows of a box
  static constexpr uint32_t EPI_BYTES = EPI_WARPS * WSLOTS * WBOX_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BLOCK_N >> 3) << 17) |
                                    (uint32_t((BLOCK_M * NCTA) >> 4) << 24);
  static constexpr uint32_t IDESC_HALF = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BLOCK_N >> 4) << 17) |
                                         (uint32_t((BLOCK_M * NCTA) >> 4) << 24);   // tail-split tiles: N = BLOCK_N / 2
  __host__ __device__ static constexpr int NUM_BARS(int stages) { return 2 * stages + 4 + EPI_WARPS * WSLOTS; }
  // row-window mode: one ring stage = a window of up to 136 input pixels x 64 channels (128-byte rows, 17 swizzle
  // atoms); the layer's weights (num_kb k-blocks of B_BYTES) sit in front of the ring
  static constexpr uint32_t ROW_STAGE_BYTES = 136 * 128;
  static int smem_bytes_row(int stages, int num_kb) {
    return num_kb * B_BYTES + stages * ROW_STAGE_BYTES + EPI_BYTES + (NUM_BARS(stages) + 2) * 8 + 16 + EPI_WARPS * SCRATCH_BYTES;   // +2: keeps the scratch 16-byte aligned
  }
  // ring | epilogue slots | barriers | TMEM slot + last-CTA flag (16 B) | per-warp scale/bias scratch | channel sums.
  // The dynamic segment is 1024-byte aligned (no static shared memory in this kernel; checked at entry).
  static int smem_bytes(int stages, int extra) {
    return stages * STAGE_BYTES + EPI_BYTES + NUM_BARS(stages) * 8 + 16 + EPI_WARPS * SCRATCH_BYTES + extra;
  }
};

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// scale/bias + activation on 32 accumulator columns.  The per-column factors come from the warp's shared-memory
// scratch (scale[32] | bias[32], broadcast reads): with ~224 KB of the SM carved out as shared memory there is no L1
// left, and per-box global loads of scale / bias cost an L2 round trip each (1.5-3.5 us per box under load, measured
// with yolo_conv_fwd_trace) -- they were the epilogue's critical path.  LeakyReLU(0.1) is branch-free: max(v, 0.1 v).
__device__ __forceinline__ void bn_act32(const uint32_t (&v)[32], float (&o)[32], uint32_t scratch, int act) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 s4, b4;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(s4.x), "=f"(s4.y), "=f"(s4.z), "=f"(s4.w) : "r"(scratch + 16u * j));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(scratch + 128u + 16u * j));
    o[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), s4.x, b4.x);
    o[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), s4.y, b4.y);
    o[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), s4.z, b4.z);
    o[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), s4.w, b4.w);
  }
  if (act == YB_ACT_LEAKY) {
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.1f * o[j]);  // NaN stays NaN (both operands NaN)
  } else if (act == YB_ACT_MISH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = mish_fast(o[j]);
  }
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// virtual tile v -> (M tile, first column, width); see ConvKParams2::t_full
template <int BLOCK_N>
__device__ __forceinline__ void decode_tile(const ConvKParams2& p, int v, int& mt, int& n0, int& nw) {
  int t = v, half = 0;
  nw = BLOCK_N;
  if (v >= p.t_full) {
    const int u = v - p.t_full;
    t = p.t_full + (u >> 1);
    half = u & 1;
    nw = BLOCK_N >> 1;
  }
  int nt;
  if (p.tiles_n_sh >= 0) {   // power-of-two tile counts (all YOLOv3 layers): no integer division on the tile path
    nt = t & ((1 << p.tiles_n_sh) - 1);
    mt = t >> p.tiles_n_sh;
  } else {
    nt = t % p.tiles_n;
    mt = t / p.tiles_n;
  }
  n0 = nt * BLOCK_N + half * nw;
}

constexpr int STEM_GATHER_WARPS = 4;  // STEM mode: warps 10..13 build the A tile from the fp32 NCHW image
constexpr int STEM_IMG_STAGES = 6;    // STEM mode: ring of TMA-loaded image windows (3 channels x 3 rows x (2 wb + 2) pixels, fp32)

// STEM = true: the network's first conv (Cin = 3, 3x3/s1/p1) straight from the NCHW fp32 image -- no patch matrix in HBM.
// A tile is one segment of an output row (row_wb pixel PAIRS, as in row-window mode).  Warp 0 TMA-loads the segment's
// 3 channels x 3 rows x (2 row_wb + 2) pixels window (zero fill = padding) into a ring; four gather warps turn each
// pixel pair's 2 x 27 taps into the 128-byte K-major row of the pair-folded GEMM (bf16, 128B swizzle) in the A ring;
// fence.proxy.async hands it to the tensor core; the 64 x 64 weight tile is resident.  NaN inputs raise the status
// flag here (model.py:175).
//
// ROW = true: row-window mode (ConvKParams2::row_mode) for the early 3x3 layers, which are bound by L2 -> SM traffic:
// resident weights, one TMA window per filter row, column taps as shifted shared-memory views.
// TRACE = true: the yolo_conv_fwd_trace build (globaltimer stamps); the product instantiations carry none of it.
template <int BLOCK_N, int KC, int NCTA, bool STEM = false, bool ROW = false, bool TRACE = false>
__global__ void __launch_bounds__(CONV2_THREADS + (STEM ? STEM_GATHER_WARPS * 32 : 0), 1)
k_conv_v2(const __grid_constant__ ConvKParams2 p) {
  using C = Cfg<BLOCK_N, KC, NCTA>;
  static_assert(!ROW || (KC == 64 && NCTA == 2 && !STEM), "row-window mode: 128-byte rows, CTA pairs");
  static_assert(!STEM || (KC == 64 && NCTA == 1 && BLOCK_N == 64), "stem mode: pair-folded 3->32 GEMM, single CTA");
  constexpr bool ROWTILES = ROW || STEM;   // tiles are output-row segments (row_coords), stores use the 4-D map
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0u) __trap();  // the 128B-swizzled stages need 1024-byte aligned bases
  const int stages = p.stages;
  // ROW / STEM: the resident weights come first; STEM: the A ring is followed by the image-window ring
  const uint32_t ring_base = (ROW || STEM) ? smem_base + uint32_t(p.num_kb) * C::B_BYTES : smem_base;
  const uint32_t stage_bytes = ROW ? C::ROW_STAGE_BYTES : (STEM ? C::A_BYTES : C::STAGE_BYTES);
  const uint32_t img_base = ring_base + stages * stage_bytes;
  const uint32_t epi_base = img_base + (STEM ? STEM_IMG_STAGES * uint32_t(p.stem_img_bytes) : 0u);
  const uint32_t bar_base = epi_base + C::EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + s * 8; };
  auto empty_bar = [&](int s) { return bar_base + (stages + s) * 8; };
  auto tfull_bar = [&](int a) { return bar_base + (2 * stages + a) * 8; };
  auto tempty_bar = [&](int a) { return bar_base + (2 * stages + 2 + a) * 8; };
  auto res_bar = [&](int w, int sl) { return bar_base + (2 * stages + 4 + w * WSLOTS + sl) * 8; };
  const uint32_t wres_bar = bar_base + C::NUM_BARS(stages) * 8;                 // ROW / STEM: the resident weights have landed
  auto img_full_bar = [&](int i) { return wres_bar + 16 + i * 8; };              // STEM: image window i has landed
  auto img_empty_bar = [&](int i) { return wres_bar + 16 + (STEM_IMG_STAGES + i) * 8; };
  constexpr int EXTRA_BARS = ROW ? 2 : (STEM ? 2 + 2 * STEM_IMG_STAGES : 0);     // even: keeps the scratch 16-byte aligned
  const uint32_t tmem_slot = bar_base + (C::NUM_BARS(stages) + EXTRA_BARS) * 8;
  const uint32_t scratch_base = tmem_slot + 16;
  const uint32_t stats_base = scratch_base + EPI_WARPS * SCRATCH_BYTES;  // training forward: [2 * c_out_pad] fp32 channel sums of this CTA

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if constexpr (TRACE) {
    if (p.trace && threadIdx.x == 0) p.trace[32 * size_t(blockIdx.x)] = gtimer();
  }
  if (p.stats != nullptr) {
    for (int i = threadIdx.x; i < 2 * p.c_out_pad; i += blockDim.x)
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(stats_base + 4u * i), "r"(0u) : "memory");
  }
  // A cluster is one CTA pair, or (p.mc == 2) two pairs on neighbouring M tiles sharing the weight tile by multicast.
  const uint32_t crank = (NCTA == 2) ? cluster_ctarank() : 0u;
  const uint32_t rank = crank & 1u;                 // rank inside the MMA pair
  const int pq = int(crank >> 1);                   // which pair of the cluster (0 unless p.mc == 2)
  const int mc = (NCTA == 2) ? p.mc : 1;
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x / (NCTA * mc);
  const int num_clusters = gridDim.x / (NCTA * mc);
  const uint16_t pair_mask = uint16_t(0x3u << (2 * pq));          // tcgen05.commit multicast: this pair's two CTAs
  const uint16_t ring_mask = mc == 2 ? uint16_t(0xF) : pair_mask;  // ... and, for ring stages, every CTA that writes into them

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmY);
    if (p.has_residual) tma_prefetch_desc(&p.tmR);
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), STEM ? STEM_GATHER_WARPS : 1);  // STEM: one arrive per gather warp, no TMA bytes
      mbar_init(empty_bar(s), (NCTA == 2) ? p.mc : 1);   // one tcgen05.commit per pair whose loads land in this stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EPI_WARPS * NCTA);  // one arrive per epilogue warp, per CTA
    }
    for (int w = 0; w < EPI_WARPS; ++w)
      for (int sl = 0; sl < WSLOTS; ++sl) mbar_init(res_bar(w, sl), 1);
    if constexpr (ROW || STEM) mbar_init(wres_bar, 1);
    if constexpr (STEM) {
      for (int i = 0; i < STEM_IMG_STAGES; ++i) {
        mbar_init(img_full_bar(i), 1);
        mbar_init(img_empty_bar(i), STEM_GATHER_WARPS);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_n<NCTA>(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  if constexpr (NCTA == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  unsigned long long* const trace = (TRACE && p.trace) ? p.trace + 32 * size_t(blockIdx.x) : nullptr;
  if (trace && threadIdx.x == 0) trace[1] = gtimer();
  if constexpr (STEM) {
    if (warp == 0) {
      if (elect_one()) {
        mbar_expect_tx(wres_bar, C::B_BYTES);
        tma_load_2d(&p.tmB, wres_bar, smem_base, 0, 0);
      }
      __syncwarp();
    }
  }
  if constexpr (ROW) {
    // The layer's weights (this CTA's half of every k-block) are parameters, not the previous launch's output: they
    // are fetched before the dependency wait, i.e. while the previous layer is still draining.
    if (warp == 0) {
      if (elect_one()) {
        if (leader) mbar_expect_tx(wres_bar, uint32_t(p.num_kb) * C::B_BYTES * NCTA);
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma_load_2d_2sm(&p.tmB, wres_bar, smem_base + uint32_t(kb) * C::B_BYTES, kb * KC, (int)rank * C::B_ROWS);
      }
      __syncwarp();
    }
  }
  if (p.pdl) {
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped
    // the tail of the previous launch in the stream.  The next launch may start its own prologue now; every access
    // to global memory below is ordered after the COMPLETION of the previous launch (and, through its own wait, of
    // all earlier ones), so reads of its output and the reuse of older buffers are both safe.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  if (trace && threadIdx.x == 0) trace[2] = gtimer();
  // ROW: virtual tile v -> this CTA's row segment: (valid, w block, output row = image * h_out + ho)
  auto row_coords = [&](int v, bool& ok, int& wblk, int& orow) {
    int seg = v * NCTA + (int)rank;
    ok = seg < p.row_total;
    if (!ok) seg = 0;            // peer CTA of a ragged last pair: loads valid data, stores nothing
    if (p.row_nblk_sh >= 0) {
      wblk = seg & ((1 << p.row_nblk_sh) - 1);
      orow = seg >> p.row_nblk_sh;
    } else {
      wblk = seg % p.row_nblk;
      orow = seg / p.row_nblk;
    }
  };

  if (warp == 0) {
    // ===== TMA producer: the whole warp runs the (warp-uniform) loop, one elected lane issues =====
    int s = 0, ntiles = 0;
    uint32_t ph = 0;
    unsigned long long wait_ns = 0;
    if constexpr (STEM) {
      int si = 0;
      uint32_t iph = 0;
      for (int v = cluster_id; v < p.num_vtiles; v += num_clusters, ++ntiles) {
        bool ok;
        int wblk, orow;
        row_coords(v, ok, wblk, orow);
        const int img = orow / p.row_h_out, ho = orow - img * p.row_h_out;
        mbar_wait(img_empty_bar(si), iph ^ 1u);
        if (elect_one()) {
          // window start: pixel 2 * segment start - 4.  TMA needs a 16-byte aligned start in the innermost (fp32 pixel)
          // dimension -- an x0 of -1 raises an illegal-instruction fault -- so the left halo pixel sits in column 3
          mbar_expect_tx(img_full_bar(si), uint32_t(p.stem_boxw) * 9u * 4u);
          tma_load_4d(&p.tmA, img_full_bar(si), img_base + si * uint32_t(p.stem_img_bytes), 2 * wblk * p.row_wb - 4, ho - 1, 0, img);
        }
        __syncwarp();
        if (++si == STEM_IMG_STAGES) { si = 0; iph ^= 1u; }
      }
    } else if constexpr (ROW) {
      const int nwin = p.ksize * p.cchunks;   // windows per tile: one per (filter row, 64-channel chunk)
      for (int v = cluster_id; v < p.num_vtiles; v += num_clusters, ++ntiles) {
        bool ok;
        int wblk, orow;
        row_coords(v, ok, wblk, orow);
        const int img = orow / p.row_h_out, ho = orow - img * p.row_h_out;
        const int w0 = wblk * p.row_wb - p.pad, h0 = ho * p.stride - p.pad;
        int r = 0, cc = 0;
        for (int i = 0; i < nwin; ++i) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (elect_one()) {
            if (leader) mbar_expect_tx(full_bar(s), uint32_t(p.row_wb + p.ksize_w - 1) * 128u * NCTA);
            tma_load_4d_2sm(&p.tmA, full_bar(s), ring_base + s * C::ROW_STAGE_BYTES, cc * KC, w0, h0 + r, img);
            if (trace && ntiles == 0 && i == 0) trace[3] = gtimer();
          }
          __syncwarp();
          if (++cc == p.cchunks) { cc = 0; ++r; }
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
    } else
    for (int v = cluster_id; v < p.num_vtiles; v += num_clusters, ++ntiles) {
      int mt, n0, nw;
      decode_tile<BLOCK_N>(p, v, mt, n0, nw);
      int m0 = ((mt * mc + pq) * NCTA + (int)rank) * BLOCK_M;
      if (m0 >= p.M) m0 = 0;  // peer CTA of a ragged last pair: load valid rows, results are discarded
      const bool whole = nw == BLOCK_N;
      const int nb0 = n0 + (int)rank * (whole ? C::B_ROWS : C::B_ROWS / 2);   // this CTA's rows of the weight tile
      const uint32_t b_bytes = whole ? C::B_BYTES : C::B_BYTES / 2;
      const bool b_second = whole && p.b_half;  // half-height weight boxes: a whole tile takes two
      int cw = 0, ch = 0, img = 0;
      if (p.a_im2col) {
        const int hw = p.h_out * p.w_out;
        img = m0 / hw;
        const int rem = m0 - img * hw;
        const int po = rem / p.w_out, qo = rem - po * p.w_out;
        cw = qo * p.stride_w - p.pad;
        ch = po * p.stride - p.pad;
      }
      int tr = 0, tq = 0, cc = 0;  // filter tap (row, col) and channel chunk of the current k-block
      for (int kb = 0; kb < p.num_kb; ++kb) {
        if (trace) {
          const unsigned long long w0 = gtimer();
          mbar_wait(empty_bar(s), ph ^ 1u);
          wait_ns += gtimer() - w0;
        } else {
          mbar_wait(empty_bar(s), ph ^ 1u);
        }
        if (elect_one()) {
          if (leader) mbar_expect_tx(full_bar(s), (STEM ? C::B_BYTES : C::A_BYTES + b_bytes) * NCTA);
          const uint32_t sa = smem_base + s * C::STAGE_BYTES, sb = sa + C::A_BYTES;
          if constexpr (STEM) {
            tma_load_2d(&p.tmB, full_bar(s), sb, kb * KC, nb0);
          } else if constexpr (NCTA == 1) {
            if (p.a_im2col) tma_load_im2col_4d(&p.tmA, full_bar(s), sa, cc * KC, cw, ch, img, (uint16_t)tq, (uint16_t)tr);
            else tma_load_2d(&p.tmA, full_bar(s), sa, cc * KC, m0);
            tma_load_2d(&p.tmB, full_bar(s), sb, kb * KC, nb0);
            if (b_second) tma_load_2d(&p.tmB, full_bar(s), sb + C::B_BYTES / 2, kb * KC, nb0 + C::B_ROWS / 2);
          } else {
            if (p.a_im2col) tma_load_im2col_4d_2sm(&p.tmA, full_bar(s), sa, cc * KC, cw, ch, img, (uint16_t)tq, (uint16_t)tr);
            else tma_load_2d_2sm(&p.tmA, full_bar(s), sa, cc * KC, m0);
            if (mc == 2 && whole) {
              // this CTA fetches rows [pq * 64, +64) of its pair-half of the weight tile for BOTH pairs
              tma_load_2d_2sm_mc(&p.tmB, full_bar(s), sb + uint32_t(pq) * (C::B_BYTES / 2), kb * KC, nb0 + pq * (C::B_ROWS / 2),
                                 uint16_t((1u << rank) | (1u << (2 + rank))));
            } else {
              tma_load_2d_2sm(&p.tmB, full_bar(s), sb, kb * KC, nb0);
              if (b_second) tma_load_2d_2sm(&p.tmB, full_bar(s), sb + C::B_BYTES / 2, kb * KC, nb0 + C::B_ROWS / 2);
            }
          }
          if (trace && ntiles == 0 && kb == 0) trace[3] = gtimer();
        }
        __syncwarp();
        if (++cc == p.cchunks) { cc = 0; if (++tq == p.ksize_w) { tq = 0; ++tr; } }
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
    if (trace && lane == 0) { trace[9] = (unsigned long long)ntiles; trace[18] = wait_ns; }
  } else if (warp == 1) {
    if (leader) {
      // ===== MMA issuer: warp-uniform loop, tcgen05.mma / commit from one elected lane =====
      int s = 0;
      uint32_t ph = 0, tl = 0;
      unsigned long long wfull_ns = 0, wacc_ns = 0;
      const uint64_t adesc0 = make_kmajor_desc<C::ROW_BYTES>((ROW || STEM) ? ring_base : smem_base);
      const uint64_t bdesc0 = make_kmajor_desc<C::ROW_BYTES>((ROW || STEM) ? smem_base : smem_base + C::A_BYTES);
      if constexpr (STEM) {
        mbar_wait(wres_bar, 0);   // resident weights
        tc_fence_after();
        for (int v = cluster_id; v < p.num_vtiles; v += num_clusters, ++tl) {
          const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
          mbar_wait(tempty_bar(acc), aph ^ 1u);
          tc_fence_after();
          mbar_wait(full_bar(s), ph);   // the gather warps have written (and proxy-fenced) this A tile
          tc_fence_after();
          if (elect_one()) {
            const uint64_t soff = uint64_t((uint32_t(s) * C::A_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              umma_bf16_n<NCTA>(tmem_base + acc * BLOCK_N, adesc0 + soff + uint64_t(2 * k), bdesc0 + uint64_t(2 * k), C::IDESC,
                                k != 0 ? 1u : 0u);
            umma_commit_n<NCTA>(empty_bar(s));
            umma_commit_n<NCTA>(tfull_bar(acc));
          }
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      } else if constexpr (ROW) {
        mbar_wait(wres_bar, 0);   // resident weights
        tc_fence_after();
        const int nwin = p.ksize * p.cchunks;
        for (int v = cluster_id; v < p.num_vtiles; v += num_clusters, ++tl) {
          const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
          mbar_wait(tempty_bar(acc), aph ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          int r = 0, cc = 0;
          for (int i = 0; i < nwin; ++i) {
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            if (trace && tl == 0 && i == 0 && lane == 0) trace[4] = gtimer();
            if (elect_one()) {
              const uint64_t soff = uint64_t((uint32_t(s) * C::ROW_STAGE_BYTES) >> 4);
              for (int st = 0; st < p.ksize_w; ++st) {
                // column tap st = the same window, st pixels (128-byte rows) further: a shifted descriptor start
                const uint64_t a_tap = adesc0 + soff + uint64_t(8 * st) + (p.row_bo ? (uint64_t(st) << 49) : 0ull);
                const uint64_t b_tap = bdesc0 + uint64_t(((uint32_t)((r * p.ksize_w + st) * p.cchunks + cc) * C::B_BYTES) >> 4);
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  umma_bf16_n<NCTA>(d_tmem, a_tap + uint64_t(2 * k), b_tap + uint64_t(2 * k), C::IDESC, (i | st | k) != 0 ? 1u : 0u);
              }
              umma_commit_n<NCTA>(empty_bar(s));
              if (i == nwin - 1) umma_commit_n<NCTA>(tfull_bar(acc));
            }
            __syncwarp();
            if (++cc == p.cchunks) { cc = 0; ++r; }
            if (++s == stages) { s = 0; ph ^= 1u; }
          }
        }
      } else
      for (int v = cluster_id; v < p.num_vtiles; v += num_clusters, ++tl) {
        const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
        const uint32_t idesc = v < p.t_full ? C::IDESC : C::IDESC_HALF;
        if (trace) {
          const unsigned long long w0 = gtimer();
          mbar_wait(tempty_bar(acc), aph ^ 1u);
          wacc_ns += gtimer() - w0;
        } else {
          mbar_wait(tempty_bar(acc), aph ^ 1u);  // the epilogue has drained this accumulator
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          if (trace) {
            const unsigned long long w0 = gtimer();
            mbar_wait(full_bar(s), ph);
            wfull_ns += gtimer() - w0;
          } else {
            mbar_wait(full_bar(s), ph);
          }
          tc_fence_after();
          if (trace && tl == 0 && kb == 0 && lane == 0) trace[4] = gtimer();
          if (elect_one()) {
            const uint64_t soff = uint64_t((uint32_t(s) * C::STAGE_BYTES) >> 4);  // stage offset in descriptor units
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              umma_bf16_n<NCTA>(d_tmem, adesc0 + soff + uint64_t(2 * k), bdesc0 + soff + uint64_t(2 * k), idesc,
                                (kb | k) != 0 ? 1u : 0u);
            umma_commit_n<NCTA>(empty_bar(s), ring_mask);
            if (kb == p.num_kb - 1) umma_commit_n<NCTA>(tfull_bar(acc), pair_mask);
          }
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
      if (trace && lane == 0) { trace[5] = gtimer(); trace[16] = wfull_ns; trace[17] = wacc_ns; }
    }
  } else if (STEM && warp >= 2 + EPI_WARPS) {
    // ===== STEM gather warps: one thread per A row = one output pixel pair =====
    if constexpr (STEM) {
      const int gr = (warp - 2 - EPI_WARPS) * 32 + lane;   // A row = pixel pair gr of the tile's row segment
      int s = 0, si = 0;
      uint32_t ph = 0, iph = 0;
      bool saw_nan = false;
      const bool in_seg = gr < p.row_wb;
      for (int v = cluster_id; v < p.num_vtiles; v += num_clusters) {
        mbar_wait(img_full_bar(si), iph);
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t row_addr = ring_base + s * C::A_BYTES + gr * 128;
        // window layout: [channel][filter row][stem_boxw floats], column 0 = pixel 2 * segment start - 4; the pair's two
        // 3x3 windows span columns 2 gr + 3 .. 2 gr + 6
        const uint32_t wbase = img_base + si * uint32_t(p.stem_img_bytes) + uint32_t(2 * gr + 3) * 4u;
        float xv[3][3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t a = wbase + uint32_t((c * 3 + kh) * p.stem_boxw) * 4u;
            if (in_seg) {
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xv[c][kh][0]) : "r"(a));
              asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(xv[c][kh][1]), "=f"(xv[c][kh][2]) : "r"(a + 4u));
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xv[c][kh][3]) : "r"(a + 12u));
            } else {
              xv[c][kh][0] = xv[c][kh][1] = xv[c][kh][2] = xv[c][kh][3] = 0.f;
            }
          }
          saw_nan |= (xv[c][1][1] != xv[c][1][1]) | (xv[c][1][2] != xv[c][1][2]);  // each image element checked once
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_local(img_empty_bar(si));   // this warp has read the window
        if (++si == STEM_IMG_STAGES) { si = 0; iph ^= 1u; }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
#pragma unroll
          for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
              for (int c = 0; c < 3; ++c) v[(kh * 3 + kw) * 3 + c] = xv[c][kh][kw + half];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t chunk = uint32_t(half * 4 + q) ^ uint32_t(gr & 7);
            const uint32_t w0 = pack_bf16(v[8 * q + 0], v[8 * q + 1]), w1 = pack_bf16(v[8 * q + 2], v[8 * q + 3]);
            const uint32_t w2 = pack_bf16(v[8 * q + 4], v[8 * q + 5]), w3 = pack_bf16(v[8 * q + 6], v[8 * q + 7]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + (chunk << 4)), "r"(w0), "r"(w1),
                         "r"(w2), "r"(w3) : "memory");
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to tcgen05.mma's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive_local(full_bar(s));
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
      if (saw_nan) atomicOr(p.status, YB_STATUS_NAN_INPUT);
    }
  } else {
    auto run_epilogue = [&](auto em_tag) {
    constexpr int EM = decltype(em_tag)::value;
    // ===== epilogue: all eight warps work on EVERY tile.  Warp w reads TMEM lane quadrant w % 4 (its 32 rows) and
    // column half (w - 2) / 4 of the tile, so two warps per scheduler share the tile's epilogue and its latency is
    // half of what one group of four needs -- that latency is the tail of every launch and, on the short-K 1x1
    // layers, the critical path (accumulator double buffering only hides it when the main loop is longer).
    // Each warp works alone: own smem slots, own TMA loads/stores, no CTA barrier.
    const int ew = warp - 2;
    const int chalf = ew >> 2;          // which half of the tile's boxes this warp owns
    const int quad = warp & 3;
    // The store / residual path is a COMPILE-TIME mode of this block (EM), picked once per launch below: the hot loop
    // then carries no flag tests (ncu: the flag-test version spent ~55 % of its 1 070 instructions per box on them).
    //   EPI_BF16      bf16 boxes staged in shared memory, TMA stores
    //   EPI_BF16_RES  + residual boxes TMA-loaded into the same slots (prefetched one box ahead)
    //   EPI_F32       fp32 boxes (the scale heads): 32-column fp32 boxes through TMA stores
    //   EPI_DIRECT    per-thread global stores (2x2 upsample replication, stride-2 data-gradient scatter)
    constexpr bool f32_staged = EM == EPI_F32;
    constexpr bool direct = EM == EPI_DIRECT;
    const bool up_coalesced = direct && C::BOXC == 64 && p.upsample2x && !p.out_fp32;   // bf16 2x2 upsample store
    const uint32_t wslot_base = epi_base + uint32_t(ew) * WSLOTS * C::WBOX_BYTES;
    const uint32_t scratch = scratch_base + uint32_t(ew) * SCRATCH_BYTES;
    const uint32_t swz = (C::BOX_ROW_BYTES == 128) ? (lane & 7) : ((lane >> 1) & 3);
    uint32_t soff[8];   // byte offset of 16-byte chunk c of this lane's row inside a swizzled box
#pragma unroll
    for (int c = 0; c < 8; ++c) soff[c] = (uint32_t(c) ^ swz) << 4;
    const int act = p.act;
    uint32_t tl = 0, wbox = 0;
    bool saw_nan = false;
    // Residual prefetch cursor.  The residual box of output box k lands in slot k % WSLOTS and is overwritten in
    // place by the output box, so its load may be issued as soon as the TMA store of box k - WSLOTS has read the
    // slot.  The cursor runs up to one box ahead of the box being computed -- across tile boundaries, where the
    // loads of the next tile's first two boxes fly while this group waits for its accumulator -- instead of
    // starting every load right before its data is needed (one exposed L2 round trip per box).
    constexpr bool res_staged = EM == EPI_BF16_RES;
    int pv = cluster_id, pb = -1;   // (virtual tile, box) the cursor points at; pb < 0: not yet placed in the tile
    uint32_t pk = 0;                                         // boxes whose residual load has been issued
    // boxes [box_lo, box_hi) of tile number t (of this CTA) with width w belong to this warp: half of the boxes each;
    // one-box tiles alternate between the two column-half groups, so that consecutive tiles' epilogues overlap
    auto box_lo = [&](int w, uint32_t t) { const int nb = w / C::BOXC; return nb >= 2 ? chalf * (nb / 2) : 0; };
    auto box_hi = [&](int w, uint32_t t) {
      const int nb = w / C::BOXC;
      return nb >= 2 ? (chalf + 1) * (nb / 2) : (uint32_t(chalf) == (t & 1u) ? nb : 0);
    };
    uint32_t ptl = 0;                  // tile number of the cursor's tile
    int c_n0 = 0, c_bhi = 0, c_a = 0, c_b = 0;   // cached for the cursor's tile: first column, end box, row coords
    // Settles the cursor on the next tile in which this warp owns valid rows and boxes (others use no slots) and
    // caches its coordinates: (first GEMM row) or, in row-window mode, (w block, output row).
    auto res_cursor_settle = [&]() {
      while (pv < p.num_vtiles) {
        int cmt, cnw;
        decode_tile<BLOCK_N>(p, pv, cmt, c_n0, cnw);
        bool rows_ok;
        if constexpr (ROWTILES) {
          bool ok;
          row_coords(pv, ok, c_a, c_b);
          rows_ok = ok && quad * 32 < p.row_wb;
        } else {
          c_a = ((cmt * mc + pq) * NCTA + (int)rank) * BLOCK_M + quad * 32;
          rows_ok = c_a < p.M;
        }
        c_bhi = box_hi(cnw, ptl);
        if (rows_ok && box_lo(cnw, ptl) < c_bhi) {
          if (pb < 0) pb = box_lo(cnw, ptl);
          break;
        }
        pv += num_clusters;
        ++ptl;
        pb = -1;
      }
    };
    auto res_issue_next = [&]() {      // warp-uniform; the caller guarantees the slot is free
      const uint32_t cslot = pk % WSLOTS;
      if (lane == 0) {
        mbar_expect_tx(res_bar(ew, cslot), C::WBOX_BYTES);
        if constexpr (ROWTILES)
          tma_load_4d(&p.tmR, res_bar(ew, cslot), wslot_base + cslot * C::WBOX_BYTES, c_n0 + pb * C::BOXC, quad * 32, c_a, c_b);
        else
          tma_load_2d(&p.tmR, res_bar(ew, cslot), wslot_base + cslot * C::WBOX_BYTES, c_n0 + pb * C::BOXC, c_a);
      }
      ++pk;
      if (++pb == c_bhi) { pb = -1; pv += num_clusters; ++ptl; res_cursor_settle(); }
    };
    if (res_staged) res_cursor_settle();
    constexpr int NCH = (BLOCK_N / 64 > C::BOXC / 32) ? BLOCK_N / 64 : C::BOXC / 32;   // 32-column chunks one warp can own: half a tile, or a whole one-box tile
    float m_sc[NCH], m_bi[NCH];   // this warp's per-column scale / bias (lane l: columns sb_lo + 32 i + l)
    int sb_lo = -1, sb_hi = -1;
    for (int v = cluster_id; v < p.num_vtiles; v += num_clusters, ++tl) {
      const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
      int mt, n0, nw;
      decode_tile<BLOCK_N>(p, v, mt, n0, nw);
      const int b_lo = box_lo(nw, tl), b_hi = box_hi(nw, tl);
      const int m0w = ((mt * mc + pq) * NCTA + (int)rank) * BLOCK_M + quad * 32;
      const int m = m0w + lane;
      bool valid = m < p.M;
      bool wvalid = m0w < p.M;
      int row_wblk = 0, row_orow = 0;
      if constexpr (ROWTILES) {   // rows of this tile = pixels quad*32 + lane of one output-row segment
        bool ok;
        row_coords(v, ok, row_wblk, row_orow);
        wvalid = ok && quad * 32 < p.row_wb;
        valid = ok && quad * 32 + lane < p.row_wb;
      }
      const bool tr0 = trace != nullptr && ew == 0 && tl == 0 && lane == 0;
      // this tile's per-column scale / bias, fetched while the accumulator is still being produced: lane l holds
      // columns n0 + 32 i + l; chunk 0 is handed to the scratch per 32-column step and the registers rotate
      float r_sc[NCH], r_bi[NCH];
      const int c_lo = n0 + b_lo * C::BOXC, c_hi = n0 + b_hi * C::BOXC;
      // (re)fetched only when the column range changes: with one N tile per layer (all 1x1 layers up to 256 channels, the
      // 52x52 3x3 layers) that is once per launch -- on short-K tiles the accumulator of the NEXT tile is usually ready
      // when a tile's epilogue ends, so a per-tile fetch exposed its L2 round trip (1-4 us under load) on every tile
      if (c_lo != sb_lo || c_hi != sb_hi) {
        sb_lo = c_lo; sb_hi = c_hi;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const bool in = c_lo + 32 * i < c_hi;
          m_sc[i] = in ? __ldg(p.scale + c_lo + 32 * i + lane) : 0.f;
          m_bi[i] = in ? __ldg(p.bias + c_lo + 32 * i + lane) : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < NCH; ++i) { r_sc[i] = m_sc[i]; r_bi[i] = m_bi[i]; }
      if (res_staged && wvalid && b_lo < b_hi) {
        // tile boundary: every earlier store of this warp was committed long ago -- take both slots
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
        while (pk < wbox + WSLOTS && pv < p.num_vtiles) res_issue_next();
      }
      size_t out_row[4];
      int n_out_rows = 1;
      out_row[0] = size_t(m);
      if constexpr (!direct) {
      } else if (p.upsample2x) {
        const int hw = p.h_out * p.w_out;
        const int img = m / hw;
        const int rem = m - img * hw;
        const int po = rem / p.w_out, qo = rem - po * p.w_out;
        const int W2 = 2 * p.w_out;
        const size_t r00 = (size_t(img) * (2 * p.h_out) + 2 * po) * W2 + 2 * qo;
        out_row[0] = r00; out_row[1] = r00 + 1; out_row[2] = r00 + W2; out_row[3] = r00 + W2 + 1;
        n_out_rows = 4;
      } else if (p.s2_parity) {
        // GEMM row (img, a, b) -> pixel (2a + r, 2b) of the (2 h_out, 2 w_out) gradient tensor; column half t adds 1
        const int hw = p.h_out * p.w_out;
        const int img = m / hw;
        const int rem = m - img * hw;
        const int po = rem / p.w_out, qo = rem - po * p.w_out;
        out_row[0] = (size_t(img) * (2 * p.h_out) + 2 * po + (p.s2_parity - 1)) * size_t(2 * p.w_out) + 2 * qo;
      }
      const bool staged = !direct && wvalid;
      // every warp waits for the accumulator and hands it back, also when it owns no box of this tile
      mbar_wait(tfull_bar(acc), aph);
      tc_fence_after();
      if (trace && tl == 0 && ew == 0 && lane == 0) trace[6] = gtimer();
      if (b_lo >= b_hi) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (NCTA == 1) mbar_arrive_local(tempty_bar(acc)); else mbar_arrive_leader(tempty_bar(acc));
        }
      }
#pragma unroll 1
      for (int b = b_lo; b < b_hi; ++b) {
        const uint32_t slot = wbox % WSLOTS, sph = (wbox / WSLOTS) & 1u;
        const uint32_t slot_addr = wslot_base + slot * C::WBOX_BYTES;
        const int nb = n0 + b * C::BOXC;
        if (staged && !res_staged) {
          if (lane == 0) {
            if constexpr (f32_staged) bulk_wait_group_read<0>();   // an fp32 box pair uses both slots: all earlier stores have read
            else bulk_wait_group_read<WSLOTS - 1>();  // the TMA store that last used this slot has read it out
          }
          __syncwarp();
        }
        const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * BLOCK_N + b * C::BOXC;
        uint32_t v0[32], v1[32];
        tmem_ld32_nowait(taddr, v0);
        if constexpr (C::BOXC == 64) tmem_ld32_nowait(taddr + 32, v1);
        tmem_wait_ld();
        if (b == b_hi - 1) {
          // every tcgen05.ld of this accumulator has completed: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (NCTA == 1) mbar_arrive_local(tempty_bar(acc)); else mbar_arrive_leader(tempty_bar(acc));
          }
        }
        if (tr0 && b == b_lo + p.trace_box) trace[10] = gtimer();
        if (staged && res_staged) {
          // one box ahead: the store of the previous box (committed a TMEM round trip ago) has read its slot
          if (pk < wbox + WSLOTS && pv < p.num_vtiles) {
            if (lane == 0) bulk_wait_group_read<0>();
            __syncwarp();
            res_issue_next();
          }
          mbar_wait(res_bar(ew, slot), sph);
        }
        if (tr0 && b == b_lo + p.trace_box) trace[11] = gtimer();
        const uint32_t row_addr = slot_addr + lane * C::BOX_ROW_BYTES;
        auto process_half = [&](const uint32_t (&v)[32], const int h) {
          const int n = nb + h * 32;
          float o[32];
          __syncwarp();  // the previous step's broadcast reads are done
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(scratch + 4u * lane), "f"(r_sc[0]) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(scratch + 128u + 4u * lane), "f"(r_bi[0]) : "memory");
#pragma unroll
          for (int i = 0; i + 1 < NCH; ++i) { r_sc[i] = r_sc[i + 1]; r_bi[i] = r_bi[i + 1]; }
          __syncwarp();
          bn_act32(v, o, scratch, act);
          if (tr0 && b == b_lo + p.trace_box && h == 0) trace[19] = gtimer();
          size_t drow = size_t(m);   // direct-path addressing: (row, column) of the residual / output element
          int dcol = n;
          if constexpr (direct) {
            if (p.s2_parity) {
              const int th = n >= p.s2_cin ? 1 : 0;
              drow = out_row[0] + th;
              dcol = n - th * p.s2_cin;
            }
          }
          if (res_staged || (direct && p.has_residual)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 r;
              if constexpr (!direct) {
                const uint32_t a = row_addr + soff[h * 4 + j];
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
              } else {
                r = valid ? __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) +
                                                                 drow * p.res_pitch + dcol) + j)
                          : make_uint4(0, 0, 0, 0);
              }
              o[8 * j + 0] += bf16_lo(r.x); o[8 * j + 1] += bf16_hi(r.x);
              o[8 * j + 2] += bf16_lo(r.y); o[8 * j + 3] += bf16_hi(r.y);
              o[8 * j + 4] += bf16_lo(r.z); o[8 * j + 5] += bf16_hi(r.z);
              o[8 * j + 6] += bf16_lo(r.w); o[8 * j + 7] += bf16_hi(r.w);
            }
          }
          if (tr0 && b == b_lo + p.trace_box && h == 0) trace[20] = gtimer();
          if (p.check_nan && valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) saw_nan |= (o[j] != o[j]);
          }
          if constexpr (f32_staged) {
            // half h of the box -> its own 4 KB slot: 32 rows x 32 fp32 (128-byte rows, 128B swizzle)
            const uint32_t fa = wslot_base + uint32_t(h) * 4096u + uint32_t(lane) * 128u;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(fa + ((uint32_t(j) ^ uint32_t(lane & 7)) << 4)),
                           "r"(__float_as_uint(o[4 * j])), "r"(__float_as_uint(o[4 * j + 1])),
                           "r"(__float_as_uint(o[4 * j + 2])), "r"(__float_as_uint(o[4 * j + 3])) : "memory");
          } else if constexpr (!direct) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t a = row_addr + soff[h * 4 + j];
              const uint32_t w0 = pack_bf16(o[8 * j + 0], o[8 * j + 1]), w1 = pack_bf16(o[8 * j + 2], o[8 * j + 3]);
              const uint32_t w2 = pack_bf16(o[8 * j + 4], o[8 * j + 5]), w3 = pack_bf16(o[8 * j + 6], o[8 * j + 7]);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                           : "memory");
            }
          } else if (up_coalesced) {
            // 2x2 upsample store, step 1: stage the bf16 half box (swizzled like a TMA box); the warp writes it out
            // below with whole 128-byte lines per 8 lanes instead of one 16-byte fragment per lane and row
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t a = wslot_base + lane * C::BOX_ROW_BYTES + soff[h * 4 + j];
              const uint32_t w0 = pack_bf16(o[8 * j + 0], o[8 * j + 1]), w1 = pack_bf16(o[8 * j + 2], o[8 * j + 3]);
              const uint32_t w2 = pack_bf16(o[8 * j + 4], o[8 * j + 5]), w3 = pack_bf16(o[8 * j + 6], o[8 * j + 7]);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                           : "memory");
            }
          } else if (valid) {
            if (p.out_fp32) {
              for (int rr = 0; rr < n_out_rows; ++rr) {
                float4* yp = reinterpret_cast<float4*>(static_cast<float*>(p.y) + out_row[rr] * p.out_pitch + n);
#pragma unroll
                for (int j = 0; j < 8; ++j) yp[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
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
                uint4* yp = p.s2_parity
                                ? reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.y) + drow * p.out_pitch + dcol)
                                : reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.y) + out_row[rr] * p.out_pitch + n);
#pragma unroll
                for (int j = 0; j < 4; ++j) yp[j] = w4[j];
              }
            }
          }
        };
        process_half(v0, 0);
        if (tr0 && b == b_lo + p.trace_box) trace[21] = gtimer();
        if constexpr (C::BOXC == 64) process_half(v1, 1);
        if (tr0 && b == b_lo + p.trace_box) trace[22] = gtimer();
        if constexpr (direct) {
          if (up_coalesced) {
            // step 2: lanes 8 i .. 8 i + 7 own row 4 it + i of the box (16 bytes each = one 128-byte line) and store it
            // to the row's 2x2 block of output pixels (nn.Upsample(scale_factor=2), model.py:222): four full lines
            // per row instead of 4 x 8 scattered 16-byte stores -- the LSU-bound pattern that cost 15-18 us per launch
            __syncwarp();
            const unsigned long long my_r00 = valid ? (unsigned long long)out_row[0] : ~0ull;
            const size_t W2 = size_t(2 * p.w_out);
            __nv_bfloat16* const ybase = static_cast<__nv_bfloat16*>(p.y) + nb + (lane & 7) * 8;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + (lane >> 3);
              uint4 val;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                           : "r"(wslot_base + uint32_t(r) * C::BOX_ROW_BYTES + ((uint32_t(lane & 7) ^ uint32_t(r & 7)) << 4)));
              const unsigned long long r00 = __shfl_sync(0xffffffffu, my_r00, r);
              if (r00 != ~0ull) {
                *reinterpret_cast<uint4*>(ybase + size_t(r00) * p.out_pitch) = val;
                *reinterpret_cast<uint4*>(ybase + size_t(r00 + 1) * p.out_pitch) = val;
                *reinterpret_cast<uint4*>(ybase + size_t(r00 + W2) * p.out_pitch) = val;
                *reinterpret_cast<uint4*>(ybase + size_t(r00 + W2 + 1) * p.out_pitch) = val;
              }
            }
            __syncwarp();   // the slot is rewritten by the next box
          }
        }
        if (staged) {
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (tr0 && b == b_lo + p.trace_box) trace[12] = gtimer();
          if (lane == 0) {
            if constexpr (f32_staged) {
              tma_store_2d(&p.tmY, wslot_base, nb, m0w);
              if constexpr (C::BOXC == 64) tma_store_2d(&p.tmY, wslot_base + 4096u, nb + 32, m0w);
            } else {
              if constexpr (ROWTILES) tma_store_4d(&p.tmY, slot_addr, nb, quad * 32, row_wblk, row_orow);  // pixels >= row_wb clipped
              else tma_store_2d(&p.tmY, slot_addr, nb, m0w);  // rows >= M are clipped by the tensor map
            }
            bulk_commit_group();
          }
          if (EM == EPI_BF16 && p.stats != nullptr) {
            // Training forward (BatchNorm batch statistics, model.py:61 in train mode): column sums of the box just
            // staged, taken from the bf16 values that are actually stored.  Lane l owns the channel pair (2l, 2l+1):
            // one 4-byte word per row, 32 lanes read one 128-byte row -> conflict free.
            constexpr int WORDS = C::BOXC / 2;
            if (lane < WORDS) {
              float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
              const uint32_t chunk = uint32_t(lane) >> 2, word = uint32_t(lane) & 3u;
              const int rows = p.M - m0w < 32 ? p.M - m0w : 32;
#pragma unroll 4
              for (int r = 0; r < rows; ++r) {
                const uint32_t rs = (C::BOX_ROW_BYTES == 128) ? uint32_t(r & 7) : uint32_t((r >> 1) & 3);
                uint32_t wv;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv)
                             : "r"(slot_addr + uint32_t(r) * C::BOX_ROW_BYTES + ((chunk ^ rs) << 4) + (word << 2)));
                const float lo = bf16_lo(wv), hi = bf16_hi(wv);
                s0 += lo; s1 += hi;
                q0 = fmaf(lo, lo, q0); q1 = fmaf(hi, hi, q1);
              }
              const uint32_t dst = stats_base + 8u * uint32_t(nb + 2 * lane);   // [2c] sum, [2c+1] sum of squares
              asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(dst), "f"(s0) : "memory");
              asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(dst + 4), "f"(q0) : "memory");
              asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(dst + 8), "f"(s1) : "memory");
              asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(dst + 12), "f"(q1) : "memory");
            }
          }
          ++wbox;
          if (tr0) { if (b == b_lo + p.trace_box) trace[13] = gtimer(); if (b == b_hi - 1) trace[14] = gtimer(); }
        }
      }
    }
    if (saw_nan) atomicOr(p.status, YB_STATUS_NAN_LAYER);
    if (lane == 0) bulk_wait_group_all();
    if (trace && lane == 0) atomicMax(trace + 7, gtimer());
    };
    // ===== head conv + anchor decode (yolo_conv_desc::decode_mode) =====
    // The accumulator row of a pixel holds the 3 x (5 + nc) head logits of its three anchor cells.  One thread scans
    // its row in column order -- sigmoid / exp on the first five of each anchor, first-maximum argmax over the class
    // logits (NaN counts as maximal), the arithmetic of csrc/decode.cu op for op (utils.py:102-143) -- and writes the
    // three 24-byte candidate rows: the fp32 head ((5 + nc) * 4 B per cell, written and read back) never exists.
    // A whole tile belongs to ONE group of four warps (tile parity), the other group takes the next tile.
    auto run_decode_epilogue = [&]() {
      const int ew = warp - 2, chalf = ew >> 2, quad = warp & 3;
      const uint32_t scratch = scratch_base + uint32_t(ew) * SCRATCH_BYTES;
      constexpr int NHALF = BLOCK_N / 32;
      const int ch = 5 + p.dec_nc, S = p.dec_S, SS = S * S;
      float* const out = static_cast<float*>(p.y);
      uint32_t tl = 0;
      for (int v = cluster_id; v < p.num_vtiles; v += num_clusters, ++tl) {
        const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
        const bool mine = uint32_t(chalf) == (tl & 1u);
        const int m = (v * NCTA + (int)rank) * BLOCK_M + quad * 32 + lane;   // tiles_n == 1: tile v = M tile v
        const bool valid = m < p.M;
        float r_sc[NHALF], r_bi[NHALF];
#pragma unroll
        for (int i = 0; i < NHALF; ++i) {
          r_sc[i] = mine ? __ldg(p.scale + 32 * i + lane) : 0.f;
          r_bi[i] = mine ? __ldg(p.bias + 32 * i + lane) : 0.f;
        }
        mbar_wait(tfull_bar(acc), aph);
        tc_fence_after();
        if (!mine) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (NCTA == 1) mbar_arrive_local(tempty_bar(acc)); else mbar_arrive_leader(tempty_bar(acc));
          }
          continue;
        }
        const int img = valid ? m / SS : 0;
        const int rem = valid ? m - img * SS : 0;
        const int ci = rem / S, cj = rem - ci * S;
        float* const orow = out + (size_t(img) * p.dec_rpi + p.dec_off + size_t(ci) * S + cj) * 6;
        int a = 0, k = 0, bi = -1;
        bool bnan = false;
        float best = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 0.f;
        // finished anchors are parked (raw t0..t4 + class) in a small per-thread array and decoded in ONE loop after the
        // scan: with the sigmoid / exp / store code inlined into each of the 32 unrolled column steps the loop body was
        // ~150 KB of code and ran from instruction-cache misses (0.20 ms instead of ~0.02 ms on the 52x52 head)
        float fin[3][6];
#pragma unroll 1
        for (int hb = 0; hb < NHALF; ++hb) {
          uint32_t vv[32];
          tmem_ld32_nowait(tmem_base + (uint32_t(quad * 32) << 16) + acc * BLOCK_N + hb * 32, vv);
          tmem_wait_ld();
          if (hb == NHALF - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (NCTA == 1) mbar_arrive_local(tempty_bar(acc)); else mbar_arrive_leader(tempty_bar(acc));
            }
          }
          __syncwarp();
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(scratch + 4u * lane), "f"(r_sc[0]) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(scratch + 128u + 4u * lane), "f"(r_bi[0]) : "memory");
#pragma unroll
          for (int i = 0; i + 1 < NHALF; ++i) { r_sc[i] = r_sc[i + 1]; r_bi[i] = r_bi[i + 1]; }
          __syncwarp();
          float o[32];
          bn_act32(vv, o, scratch, p.act);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const float val = o[jj];
            if (a < 3) {
              if (k < 5) {
                t0 = k == 0 ? val : t0; t1 = k == 1 ? val : t1; t2 = k == 2 ? val : t2; t3 = k == 3 ? val : t3;
                t4 = k == 4 ? val : t4;
              } else {
                // argmax over raw logits: first maximal index, NaN counts as maximal.  Branch-free (selects): the test
                // differs per lane, and a divergent branch per column costs far more than three selects
                const bool vn = val != val;
                const bool take = (bi < 0) | (!bnan & (vn | (val > best)));
                best = take ? val : best;
                bi = take ? k - 5 : bi;
                bnan = take ? vn : bnan;
              }
              if (++k == ch) {
                fin[a][0] = t0; fin[a][1] = t1; fin[a][2] = t2; fin[a][3] = t3; fin[a][4] = t4;
                fin[a][5] = bi < 0 ? 0.f : float(bi);
                k = 0; ++a; bi = -1; bnan = false;
              }
            }
          }
        }
        if (valid) {
#pragma unroll 1
          for (int a2 = 0; a2 < 3; ++a2) {
            const float x = 1.0f / (1.0f + expf(-fin[a2][0])), y = 1.0f / (1.0f + expf(-fin[a2][1]));   // utils.py:106
            const float w = __fmul_rn(expf(fin[a2][2]), p.dec_anchors[2 * a2]);                           // utils.py:110
            const float h = __fmul_rn(expf(fin[a2][3]), p.dec_anchors[2 * a2 + 1]);
            const float obj = 1.0f / (1.0f + expf(-fin[a2][4]));                                          // utils.py:111
            float2* dst = reinterpret_cast<float2*>(orow + size_t(a2) * SS * 6);
            dst[0] = make_float2(__fmul_rn(p.dec_inv_s, __fadd_rn(x, float(cj))),                         // utils.py:125
                                 __fmul_rn(p.dec_inv_s, __fadd_rn(y, float(ci))));                        // utils.py:142
            dst[1] = make_float2(__fmul_rn(p.dec_inv_s, w), __fmul_rn(p.dec_inv_s, h));                   // utils.py:143
            dst[2] = make_float2(obj, fin[a2][5]);
          }
        }
      }
    };
    // The same scan with the class count known at compile time (COCO's 80 on the 256-column tile, the turbine model's 2
    // on the 32-column tile): after full unrolling every column's role (which anchor, which of its 5 + nc values) is a
    // constant, so a class column costs one compare-or-unordered, one NaN guard and two selects, and the decode of a
    // finished anchor sits at exactly three places in straight-line code.
    auto run_decode_epilogue_static = [&](auto nc_tag) {
      constexpr int NC = decltype(nc_tag)::value, CH = 5 + NC;
      const int ew = warp - 2, chalf = ew >> 2, quad = warp & 3;
      const uint32_t scratch = scratch_base + uint32_t(ew) * SCRATCH_BYTES;
      constexpr int NHALF = BLOCK_N / 32;
      const int S = p.dec_S, SS = S * S;
      float* const out = static_cast<float*>(p.y);
      uint32_t tl = 0;
      for (int v = cluster_id; v < p.num_vtiles; v += num_clusters, ++tl) {
        const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
        const bool mine = uint32_t(chalf) == (tl & 1u);
        const int m = (v * NCTA + (int)rank) * BLOCK_M + quad * 32 + lane;
        const bool valid = m < p.M;
        float r_sc[NHALF], r_bi[NHALF];
#pragma unroll
        for (int i = 0; i < NHALF; ++i) {
          r_sc[i] = mine ? __ldg(p.scale + 32 * i + lane) : 0.f;
          r_bi[i] = mine ? __ldg(p.bias + 32 * i + lane) : 0.f;
        }
        mbar_wait(tfull_bar(acc), aph);
        tc_fence_after();
        if (!mine) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (NCTA == 1) mbar_arrive_local(tempty_bar(acc)); else mbar_arrive_leader(tempty_bar(acc));
          }
          continue;
        }
        const int img = valid ? m / SS : 0;
        const int rem = valid ? m - img * SS : 0;
        const int ci = rem / S, cj = rem - ci * S;
        float* const orow = out + (size_t(img) * p.dec_rpi + p.dec_off + size_t(ci) * S + cj) * 6;
        float best = -INFINITY, t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 0.f;
        int bi = -1;
#pragma unroll
        for (int hb = 0; hb < NHALF; ++hb) {
          if (hb * 32 < 3 * CH) {   // (compile time) chunks past the last anchor are not read at all
            uint32_t vv[32];
            tmem_ld32_nowait(tmem_base + (uint32_t(quad * 32) << 16) + acc * BLOCK_N + hb * 32, vv);
            tmem_wait_ld();
          if ((hb + 1) * 32 >= 3 * CH) {   // last chunk that is read: hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (NCTA == 1) mbar_arrive_local(tempty_bar(acc)); else mbar_arrive_leader(tempty_bar(acc));
            }
          }
          __syncwarp();
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(scratch + 4u * lane), "f"(r_sc[hb]) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(scratch + 128u + 4u * lane), "f"(r_bi[hb]) : "memory");
          __syncwarp();
          float o[32];
          bn_act32(vv, o, scratch, YB_ACT_NONE);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int c = hb * 32 + jj;
            if (c < 3 * CH) {
              const int k = c % CH, a = c / CH;
              const float val = o[jj];
              if (k == 0) t0 = val;
              else if (k == 1) t1 = val;
              else if (k == 2) t2 = val;
              else if (k == 3) t3 = val;
              else if (k == 4) t4 = val;
              else {
                // first maximal index, NaN counts as maximal: take when val > best or val is NaN, unless best already is
                const bool take = !(val <= best) & (best == best);
                best = take ? val : best;
                bi = take ? k - 5 : bi;
              }
              if (k == CH - 1) {
                if (valid) {
                  const float x = 1.0f / (1.0f + expf(-t0)), y = 1.0f / (1.0f + expf(-t1));           // utils.py:106
                  const float w = __fmul_rn(expf(t2), p.dec_anchors[2 * a]);                           // utils.py:110
                  const float h = __fmul_rn(expf(t3), p.dec_anchors[2 * a + 1]);
                  const float obj = 1.0f / (1.0f + expf(-t4));                                         // utils.py:111
                  float2* dst = reinterpret_cast<float2*>(orow + size_t(a) * SS * 6);
                  dst[0] = make_float2(__fmul_rn(p.dec_inv_s, __fadd_rn(x, float(cj))),                // utils.py:125
                                       __fmul_rn(p.dec_inv_s, __fadd_rn(y, float(ci))));               // utils.py:142
                  dst[1] = make_float2(__fmul_rn(p.dec_inv_s, w), __fmul_rn(p.dec_inv_s, h));          // utils.py:143
                  dst[2] = make_float2(obj, bi < 0 ? 0.f : float(bi));
                }
                best = -INFINITY;
                bi = -1;
              }
            }
          }
          }
        }
      }
    };
    if (p.dec_mode) {
      if constexpr (BLOCK_N == 256 && !STEM && !ROW) {
        if (p.dec_nc == 80) run_decode_epilogue_static(std::integral_constant<int, 80>{});
        else run_decode_epilogue();
      } else if constexpr (BLOCK_N == 32 && !STEM && !ROW) {
        if (p.dec_nc == 2) run_decode_epilogue_static(std::integral_constant<int, 2>{});
        else run_decode_epilogue();
      }
    } else if (p.upsample2x || p.s2_parity != 0) run_epilogue(std::integral_constant<int, EPI_DIRECT>{});
    else if (p.out_fp32) run_epilogue(std::integral_constant<int, EPI_F32>{});
    else if (p.has_residual) run_epilogue(std::integral_constant<int, EPI_BF16_RES>{});
    else run_epilogue(std::integral_constant<int, EPI_BF16>{});
  }

  tc_fence_before();
  if constexpr (NCTA == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) tmem_dealloc_n<NCTA>(tmem_base, C::TMEM_COLS);
  if (trace && threadIdx.x == 0) trace[8] = gtimer();
  if (p.stats != nullptr) {  // one double atomic per channel statistic per CTA
    for (int i = threadIdx.x; i < 2 * p.c_out_pad; i += blockDim.x) {
      float v;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(stats_base + 4u * i));
      if (v != 0.f) atomicAdd(p.stats + i, double(v));
    }
    if (p.fin_counter != nullptr) {
      // the CTA that takes the last ticket sees every CTA's sums: it turns them into the layer's BatchNorm
      // mean / rstd / scale / bias (and running statistics) -- no separate finalize launch on the critical path
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(p.fin_counter, 1u);
        const uint32_t last = ticket == gridDim.x - 1 ? 1u : 0u;
        if (last) *p.fin_counter = 0u;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tmem_slot + 8), "r"(last) : "memory");
      }
      __syncthreads();
      uint32_t is_last;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(is_last) : "r"(tmem_slot + 8));
      if (is_last) {
        __threadfence();
        for (int c = threadIdx.x; c < p.c_out_pad; c += blockDim.x)
          bn_finalize_channel(p.fin, c, __ldcg(p.stats + 2 * c), __ldcg(p.stats + 2 * c + 1));
      }
    }
  }
}

template <int BN, int KC, int NCTA, bool STEM = false, bool ROW = false>
int launch2(const ConvPlan* pl, const ConvKParams2& kp, cudaStream_t stream) {
  auto kern = k_conv_v2<BN, KC, NCTA, STEM, ROW, false>;
  if (kp.trace != nullptr) {   // dev tool: instantiated for the CTA-pair kernels with 128-byte rows only
    if constexpr (KC == 64 && NCTA == 2 && !STEM && BN >= 64) {
      kern = k_conv_v2<BN, KC, NCTA, STEM, ROW, true>;
    } else {
      yb_set_error("conv trace: no trace build of this tile configuration (block_n %d, kc %d, ctas %d)", BN, KC, NCTA);
      return YB_ERR_UNSUPPORTED;
    }
  }
  YB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->smem_bytes));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)pl->grid2);
  cfg.blockDim = dim3(CONV2_THREADS + (STEM ? STEM_GATHER_WARPS * 32 : 0));
  cfg.dynamicSmemBytes = (size_t)pl->smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pl->csize > 0 ? pl->csize : NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kp.pdl ? 2 : 1;
  YB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, kp));
  return YB_OK;
}

template <int BN, int KC>
int smem_for(int ncta, int stages, int extra) {
  return ncta == 2 ? Cfg<BN, KC, 2>::smem_bytes(stages, extra) : Cfg<BN, KC, 1>::smem_bytes(stages, extra);
}

int smem_bytes_row_v2(int bn, int stages, int num_kb) {
  return bn == 64 ? Cfg<64, 64, 2>::smem_bytes_row(stages, num_kb) : Cfg<128, 64, 2>::smem_bytes_row(stages, num_kb);
}

int smem_bytes_v2(int bn, int kc, int ncta, int stages, int extra) {
#define YB_C2(BN, KC) if (bn == BN && kc == KC) return smem_for<BN, KC>(ncta, stages, extra);
  YB_C2(32, 32) YB_C2(64, 32) YB_C2(128, 32) YB_C2(256, 32) YB_C2(32, 64) YB_C2(64, 64) YB_C2(128, 64) YB_C2(256, 64)
#undef YB_C2
  return 1 << 30;
}

}  // namespace

// co-resident clusters of 4 CTAs of the 256-wide pair kernel (B200: 33 = 132 of 148 SMs), queried once
static int clusters_of_4() {
  static int n4 = -1;
  if (n4 < 0) {
    int v = 0;
    n4 = (conv2_query_max_clusters(4, &v) == YB_OK) ? v : 0;
  }
  return n4;
}

int conv2_plan_setup(ConvPlan* pl, const yolo_conv_desc* d, int h_out, int w_out, int im2col, PFN_encodeTiled encTiled,
                     const void* x, const void* residual, void* y) {
  const int kc = pl->kc;
  const bool stem = d->stem_c > 0;
  // default: a cta_group::2 pair (measured faster or equal on every YOLOv3 layer); 1 forces single CTAs
  int ncta = d->cta_pair_hint == 1 ? 1 : 2;
  int bn = pl->block_n;
  if (ncta == 2 && d->block_n_hint == 0 && d->c_out_pad % 256 == 0) bn = 256;  // a pair splits the weight tile
  if (ncta == 2 && bn < 64) ncta = 1;
  if (stem) ncta = 1;  // HBM-bound layer; keeps the gather warps' arrivals CTA-local
  const long long M = (long long)d->batch * h_out * w_out;
  const int tiles_m = (int)((M + BLOCK_M * ncta - 1) / (BLOCK_M * ncta));
  const int tiles_n = d->c_out_pad / bn;
  const int taps = d->ksize * yb_kw(d);
  const int num_kb = taps * (d->c_in / kc);
  const int extra = d->want_stats ? 8 * d->c_out_pad : 0;  // per-CTA channel sums of yolo_conv_fwd_stats
  int stages = d->stages_hint;
  if (stages <= 0) {
    stages = 8;
    while (stages > 1 && smem_bytes_v2(bn, kc, ncta, stages, extra) > 227 * 1024) --stages;
  }
  // the smem ring spans tiles in a persistent kernel: never clamp it to this layer's k-block count
  while (stages > 1 && smem_bytes_v2(bn, kc, ncta, stages, extra) > 227 * 1024) --stages;  // a hint is a ceiling
  const int smem = smem_bytes_v2(bn, kc, ncta, stages, extra);
  YB_REQUIRE(smem <= 227 * 1024, "conv v2: %d stages do not fit shared memory (block_n %d)", stages, bn);

  int dev = 0, sms = 148;
  YB_CHECK_CUDA(cudaGetDevice(&dev));
  YB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int direct = d->upsample2x || d->s2_parity;
  // ---- weight-tile multicast across two CTA pairs (ConvKParams2::mc) ------------------------------------------
  int mc = 1;
  if (d->mc_hint == 2 && ncta == 2 && bn == 256 && kc == 64 && !stem && !d->out_fp32 && !direct && !d->want_stats &&
      tiles_m >= 8 && d->block_n_hint == 0 && d->stages_hint == 0) {
    if (clusters_of_4() >= 16) mc = 2;
  }
  const int max_clusters = mc == 2 ? clusters_of_4() : sms / ncta;
  const int tiles_m_sched = (tiles_m + mc - 1) / mc;   // M tiles per cluster step
  // Tail splitting (ConvKParams2::t_full): the last round of a persistent launch holds num_tiles % clusters tiles.
  // When at most half of the CTA pairs would get one, every such tile is cut into two half-width tiles: the round
  // then costs about half a tile time (13x13 3x3 layers: 2.5 instead of 3 rounds).
  int t_full = tiles_m_sched * tiles_n, num_vtiles = t_full, b_half = mc == 2 ? 1 : 0;   // mc: 64-row weight boxes
  if (d->tail_split_hint != 1 && bn >= 128 && !d->out_fp32 && !direct && !d->want_stats && !stem) {
    const int rem = t_full % max_clusters;
    if (rem > 0 && 2 * rem <= max_clusters) {
      t_full -= rem;
      num_vtiles = t_full + 2 * rem;
      b_half = 1;
    }
  }

  // ---- row-window mode (ConvKParams2::row_mode): 3x3 layers whose whole weight tensor fits beside the ring ----------
  int row_mode = 0, row_wb = 0, row_nblk = 1, row_stages = 0, row_smem = 0;
  {
    const int kw = yb_kw(d);
    const bool shape_ok = !stem && ncta == 2 && kc == 64 && d->ksize == 3 && (kw == 3 || kw == 2) && yb_sw(d) == 1 &&
                          d->pad == 1 && (d->stride == 1 || d->stride == 2) && tiles_n == 1 && (bn == 64 || bn == 128) &&
                          !direct && !d->out_fp32 && !d->want_stats && d->a_mode == 0 && d->block_n_hint == 0 &&
                          d->stages_hint == 0 && x != nullptr;
    if (d->row_hint != 1 && shape_ok) {
      int nblk = (w_out + 127) / 128;
      while (nblk <= w_out && (w_out % nblk != 0 || w_out / nblk > 128)) ++nblk;
      const int wb = nblk <= w_out ? w_out / nblk : 0;
      int st = 6;
      while (st > 2 && smem_bytes_row_v2(bn, st, num_kb) > 227 * 1024) --st;
      // worth it when the window is reasonably full (the MMA always runs 128 rows) and at least 3 windows are in flight
      if (wb >= 64 && st >= 3 && smem_bytes_row_v2(bn, st, num_kb) <= 227 * 1024) {
        row_mode = 1; row_wb = wb; row_nblk = nblk; row_stages = st; row_smem = smem_bytes_row_v2(bn, st, num_kb);
      }
    }
  }
  if (row_mode) { t_full = num_vtiles = (int)(((long long)d->batch * h_out * row_nblk + 1) / 2); b_half = 0; }
  // ---- fused stem: tiles are row segments of pixel pairs too; the image window box holds 2 wb + 2 pixels (<= 256) -----
  int stem_boxw = 0, stem_img_bytes = 0, stem_smem = 0;
  if (stem) {
    int nblk = (w_out + 123) / 124;
    while (nblk <= w_out && (w_out % nblk != 0 || w_out / nblk > 124 || (w_out / nblk) % 2 != 0)) ++nblk;
    YB_REQUIRE(nblk <= w_out && w_out / nblk >= 32, "conv stem: image width %d (pairs) has no usable row segmentation", w_out);
    row_wb = w_out / nblk; row_nblk = nblk;
    stem_boxw = (2 * row_wb + 5 + 3) / 4 * 4;   // pixels -4 .. 2 wb + 1 of the segment, padded to 16 bytes
    stem_img_bytes = (stem_boxw * 9 * 4 + 1023) / 1024 * 1024;
    stages = 4;
    stem_smem = Cfg<64, 64, 1>::B_BYTES + stages * Cfg<64, 64, 1>::A_BYTES + STEM_IMG_STAGES * stem_img_bytes + Cfg<64, 64, 1>::EPI_BYTES +
                (Cfg<64, 64, 1>::NUM_BARS(stages) + 2 + 2 * STEM_IMG_STAGES) * 8 + 16 + EPI_WARPS * SCRATCH_BYTES;
    YB_REQUIRE(stem_smem <= 227 * 1024, "conv stem: shared memory");
    t_full = num_vtiles = d->batch * h_out * row_nblk;
    b_half = 0;
  }

  ConvKParams2& kp = pl->kp2;
  kp.tmA = pl->kp.tmA;  // same A geometry as v1 (128-row boxes of KC channels); unused by the stem
  if (row_mode) {   // input windows: (channel, w, h, image) tiles of 64 x (row_wb + kw - 1) x 1 x 1, zero fill = padding
    cuuint64_t dims[4] = {(cuuint64_t)d->c_in, (cuuint64_t)d->w_in, (cuuint64_t)d->h_in, (cuuint64_t)d->batch};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_pitch * 2, (cuuint64_t)d->w_in * d->in_pitch * 2,
                             (cuuint64_t)d->h_in * d->w_in * d->in_pitch * 2};
    cuuint32_t box[4] = {64u, (cuuint32_t)(row_wb + yb_kw(d) - 1), 1u, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encTiled(&kp.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    YB_REQUIRE(cr == CUDA_SUCCESS, "conv v2: row-window tensor map A encode failed (%d)", (int)cr);
  }
  // weight tile: BLOCK_N / NCTA rows per CTA (two half-height boxes when the tail is split)
  {
    const CUtensorMapSwizzle swz = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint64_t dims[2] = {(cuuint64_t)taps * d->c_in, (cuuint64_t)d->c_out_pad};
    cuuint64_t strides[1] = {(cuuint64_t)taps * d->c_in * 2};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)(bn / ncta / (b_half ? 2 : 1))};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = encTiled(&kp.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(pl->w), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    YB_REQUIRE(cr == CUDA_SUCCESS, "conv v2: tensor map B encode failed (%d)", (int)cr);
  }
  const int boxc = bn < 64 ? bn : 64;
  const CUtensorMapSwizzle bswz = boxc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (d->decode_mode) {
    YB_REQUIRE(d->out_fp32 && !direct && !d->has_residual && d->act == YB_ACT_NONE && tiles_n == 1 && (bn == 256 || bn == 32) &&
                   3 * (5 + d->dec_nc) <= d->c_out_pad && d->dec_nc >= 1 && h_out == w_out && !stem && !row_mode,
               "conv decode mode: needs a bias-only head conv whose 3 x (5 + nc) channels form one tile of 32 or 256");
    kp.tmY = kp.tmA;   // no TMA store in this mode: the epilogue writes 24-byte candidate rows
    kp.tmR = kp.tmA;
  } else if (d->out_fp32 && !direct) {   // scale heads: fp32 boxes of 32 columns x 32 rows
    cuuint64_t dims[2] = {(cuuint64_t)d->c_out_pad, (cuuint64_t)M};
    cuuint32_t box[2] = {32u, 32u};
    cuuint32_t estr[2] = {1, 1};
    cuuint64_t ystr[1] = {(cuuint64_t)d->out_pitch * 4};
    CUresult cr = encTiled(&kp.tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, y, dims, ystr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    YB_REQUIRE(cr == CUDA_SUCCESS, "conv v2: fp32 tensor map Y encode failed (%d)", (int)cr);
    kp.tmR = kp.tmY;
  } else if (row_mode || stem) {   // output / residual: (channel, w in segment, segment, image * h_out + ho), 32-pixel boxes per warp
    cuuint64_t dims[4] = {(cuuint64_t)d->c_out_pad, (cuuint64_t)row_wb, (cuuint64_t)row_nblk, (cuuint64_t)d->batch * h_out};
    cuuint32_t box[4] = {(cuuint32_t)boxc, 32u, 1u, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    cuuint64_t ystr[3] = {(cuuint64_t)d->out_pitch * 2, (cuuint64_t)row_wb * d->out_pitch * 2, (cuuint64_t)w_out * d->out_pitch * 2};
    CUresult cr = encTiled(&kp.tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, dims, ystr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           bswz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    YB_REQUIRE(cr == CUDA_SUCCESS, "conv v2: row-window tensor map Y encode failed (%d)", (int)cr);
    if (d->has_residual) {
      cuuint64_t rstr[3] = {(cuuint64_t)d->res_pitch * 2, (cuuint64_t)row_wb * d->res_pitch * 2, (cuuint64_t)w_out * d->res_pitch * 2};
      cr = encTiled(&kp.tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(residual), dims, rstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bswz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      YB_REQUIRE(cr == CUDA_SUCCESS, "conv v2: row-window tensor map R encode failed (%d)", (int)cr);
    } else {
      kp.tmR = kp.tmY;
    }
  } else if (!direct) {
    cuuint64_t dims[2] = {(cuuint64_t)d->c_out_pad, (cuuint64_t)M};
    cuuint32_t box[2] = {(cuuint32_t)boxc, 32u};  // one epilogue warp's 32 rows
    cuuint32_t estr[2] = {1, 1};
    cuuint64_t ystr[1] = {(cuuint64_t)d->out_pitch * 2};
    CUresult cr = encTiled(&kp.tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y, dims, ystr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, bswz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    YB_REQUIRE(cr == CUDA_SUCCESS, "conv v2: tensor map Y encode failed (%d)", (int)cr);
    if (d->has_residual) {
      cuuint64_t rstr[1] = {(cuuint64_t)d->res_pitch * 2};
      cr = encTiled(&kp.tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(residual), dims, rstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bswz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      YB_REQUIRE(cr == CUDA_SUCCESS, "conv v2: tensor map R encode failed (%d)", (int)cr);
    } else {
      kp.tmR = kp.tmY;
    }
  } else {
    kp.tmY = kp.tmA;
    kp.tmR = kp.tmA;
  }
  kp.scale = pl->kp.scale; kp.bias = pl->kp.bias; kp.residual = residual; kp.y = y; kp.status = nullptr;
  kp.M = (int)M; kp.h_out = h_out; kp.w_out = w_out;
  kp.out_pitch = d->out_pitch; kp.res_pitch = d->res_pitch;
  kp.num_kb = num_kb; kp.cchunks = d->c_in / kc; kp.stages = stages;
  kp.tiles_n = tiles_n; kp.num_tiles = tiles_m * tiles_n;
  kp.ksize = d->ksize; kp.stride = d->stride; kp.pad = d->pad; kp.ksize_w = yb_kw(d); kp.stride_w = yb_sw(d);
  kp.act = d->act; kp.has_residual = d->has_residual; kp.upsample2x = d->upsample2x;
  kp.out_fp32 = d->out_fp32; kp.check_nan = d->check_nan; kp.a_im2col = im2col;
  kp.stem_x = nullptr; kp.stem_h = d->h_in; kp.stem_w = d->w_in * 2;
  kp.stats = nullptr; kp.c_out_pad = d->c_out_pad; kp.fin_counter = nullptr;
  kp.s2_parity = d->s2_parity; kp.s2_cin = d->s2_cin;
  kp.t_full = t_full; kp.num_vtiles = num_vtiles; kp.b_half = b_half;
  auto log2_or_neg = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  kp.tiles_n_sh = log2_or_neg(tiles_n); kp.row_nblk_sh = log2_or_neg(row_nblk);
  kp.row_mode = row_mode; kp.row_wb = row_wb; kp.row_nblk = row_nblk; kp.row_h_out = h_out;
  kp.row_total = d->batch * h_out * row_nblk; kp.row_bo = d->row_hint == 2 ? 1 : 0;
  if (row_mode) kp.stages = row_stages;
  kp.pdl = d->pdl_hint == 1 ? 0 : 1;
  kp.mc = mc;
  kp.trace = nullptr; kp.trace_box = 0;
  kp.dec_mode = d->decode_mode ? 1 : 0; kp.dec_nc = d->dec_nc; kp.dec_S = h_out; kp.dec_rpi = d->dec_rows_per_image;
  kp.dec_off = d->dec_row_offset;
  kp.dec_inv_s = (float)(1.0 / (double)(h_out > 0 ? h_out : 1));   // utils.py:125: a Python double 1/S cast to fp32
  for (int i = 0; i < 6; ++i) memcpy(&kp.dec_anchors[i], &d->dec_anchor_bits[i], 4);
  pl->stem_direct = stem ? 1 : 0;
  if (stem) {
    YB_REQUIRE(bn == 64 && kc == 64 && d->stem_c == 3 && d->ksize == 1 && tiles_n == 1 && !d->has_residual && !direct,
               "conv stem: needs the pair-folded 3->32 stem geometry (c_in 64, c_out 64)");
    kp.stages = stages;
    kp.stem_boxw = stem_boxw; kp.stem_img_bytes = stem_img_bytes;
    pl->enc_tiled = (void*)encTiled;
  }

  int clusters = max_clusters;
  if (clusters > kp.num_vtiles) clusters = kp.num_vtiles;
  pl->grid2 = clusters * ncta * mc;
  pl->csize = ncta * mc;
  pl->ncta = ncta;
  pl->block_n = bn;
  pl->smem_bytes = stem ? stem_smem : (row_mode ? row_smem : smem);
  pl->grid_x = tiles_n;
  pl->grid_y = tiles_m;
  return YB_OK;
}

// DEV: how many clusters of `cluster` CTAs of the 256x64 pair kernel can be co-resident (SM stranding by cluster size)
int conv2_query_max_clusters(int cluster, int* out) {
  auto kern = k_conv_v2<256, 64, 2, false>;
  const int smem = Cfg<256, 64, 2>::smem_bytes(5, 0);
  YB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (cluster > 8) YB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(cluster * 148));
  cfg.blockDim = dim3(CONV2_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  YB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(out, kern, &cfg));
  return YB_OK;
}

int conv2_launch_stem(const ConvPlan* pl, const float* x_nchw, uint32_t* status, cudaStream_t stream) {
  YB_REQUIRE(pl->stem_direct && pl->block_n == 64 && pl->kc == 64 && pl->ncta == 1, "conv stem: plan is not a stem plan");
  YB_REQUIRE(x_nchw && status, "conv stem: null pointer");
  YB_REQUIRE((reinterpret_cast<uintptr_t>(x_nchw) & 15) == 0, "conv stem: the image must be 16-byte aligned");
  ConvKParams2 kp = pl->kp2;
  kp.status = status;
  kp.stem_x = x_nchw;
  {   // the caller's image moves from call to call: its (W, H, 3, B) fp32 tensor map is encoded per launch (host, ~1 us)
    const int W = kp.stem_w, H = kp.stem_h;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3ull, (cuuint64_t)pl->d.batch};
    cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)kp.stem_boxw, 3u, 3u, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = ((PFN_encodeTiled)pl->enc_tiled)(&kp.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x_nchw), dims, strides,
                                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    YB_REQUIRE(cr == CUDA_SUCCESS, "conv stem: image tensor map encode failed (%d)", (int)cr);
  }
  return launch2<64, 64, 1, true>(pl, kp, stream);
}

int conv2_launch(const ConvPlan* pl, uint32_t* status, cudaStream_t stream, double* stats, const BnFinalize* fin,
                 unsigned int* fin_counter, unsigned long long* trace, int trace_box) {
  YB_REQUIRE(!pl->stem_direct, "conv fwd: stem plans are launched with yolo_conv_fwd_stem");
  YB_REQUIRE(!stats || (pl->d.want_stats && !(pl->d.upsample2x || pl->d.out_fp32)),
             "conv fwd: statistics need a plan built with want_stats and the staged bf16 output path");
  ConvKParams2 kp = pl->kp2;
  kp.status = status;
  kp.stats = stats;
  kp.trace = trace;
  kp.trace_box = trace_box;
  if (stats && fin && fin_counter) { kp.fin = *fin; kp.fin_counter = fin_counter; }
  if (kp.row_mode) {
    if (pl->block_n == 64) return launch2<64, 64, 2, false, true>(pl, kp, stream);
    if (pl->block_n == 128) return launch2<128, 64, 2, false, true>(pl, kp, stream);
    yb_set_error("conv v2: no row-window kernel for block_n %d", pl->block_n);
    return YB_ERR_UNSUPPORTED;
  }
#define YB_L2(BN, KC)                                                        \
  if (pl->block_n == BN && pl->kc == KC)                                     \
    return pl->ncta == 2 ? launch2<BN, KC, 2>(pl, kp, stream) : launch2<BN, KC, 1>(pl, kp, stream);
  YB_L2(32, 32) YB_L2(64, 32) YB_L2(128, 32) YB_L2(256, 32) YB_L2(32, 64) YB_L2(64, 64) YB_L2(128, 64) YB_L2(256, 64)
#undef YB_L2
  yb_set_error("conv v2: no kernel for block_n %d kc %d", pl->block_n, pl->kc);
  return YB_ERR_UNSUPPORTED;
}
