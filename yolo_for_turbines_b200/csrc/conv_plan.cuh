// Plan blob shared by the conv kernels (opaque to the C-ABI caller).
#pragma once
#include <cuda.h>

#include "common.cuh"

constexpr int BLOCK_M = 128;
constexpr uint32_t PLAN_MAGIC = 0x59423230u;  // "YB20"

inline int yb_kw(const yolo_conv_desc* d) { return d->ksize_w > 0 ? d->ksize_w : d->ksize; }
inline int yb_sw(const yolo_conv_desc* d) { return d->stride_w > 0 ? d->stride_w : d->stride; }
inline int yb_pad_hi(const yolo_conv_desc* d) { return d->pad_w_hi_plus1 > 0 ? d->pad_w_hi_plus1 - 1 : d->pad; }
inline int yb_pad_h_hi(const yolo_conv_desc* d) { return d->pad_h_hi_plus1 > 0 ? d->pad_h_hi_plus1 - 1 : d->pad; }

// v1: one CTA per output tile (conv.cu)
struct ConvKParams {
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmB;
  const float* scale;
  const float* bias;
  const void* residual;
  void* y;
  uint32_t* status;
  int M, h_out, w_out;
  int out_pitch, res_pitch;
  int num_kb, cchunks, stages, tiles_n;
  int ksize, stride, pad;
  int ksize_w, stride_w;
  int act, has_residual, upsample2x, out_fp32, check_nan, a_im2col;
};

// v2: persistent, TMEM double-buffered, TMA-store epilogue, optional CTA pair (conv2.cu)
struct ConvKParams2 {
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmB;
  alignas(64) CUtensorMap tmY;  // output boxes   [128 rows x min(64, BLOCK_N) cols]
  alignas(64) CUtensorMap tmR;  // residual boxes, same geometry
  const float* scale;
  const float* bias;
  const void* residual;
  void* y;
  uint32_t* status;
  int M, h_out, w_out;
  int out_pitch, res_pitch;
  int num_kb, cchunks, stages, tiles_n, num_tiles;  // tiles of (128 * NCTA) x BLOCK_N
  int ksize, stride, pad;
  int ksize_w, stride_w;
  int act, has_residual, upsample2x, out_fp32, check_nan, a_im2col;
  const float* stem_x;  // STEM mode: fp32 NCHW image (set per launch)
  int stem_h, stem_w;
  int stem_boxw, stem_img_bytes;   // STEM: floats per line of an image window in shared memory, bytes per window stage
  int c_out_pad;
  int s2_parity, s2_cin;  // stride-2 data-gradient sub-convolution (yolo_conv_desc.s2_parity)
  BnFinalize fin;       // training forward: finalize by the CTA that finishes last (fin_counter != nullptr)
  unsigned int* fin_counter;
  double* stats;        // training forward: per-channel [sum, sum of squares] of the stored bf16 output (set per launch)
  // Tail splitting: virtual tiles [0, t_full) are whole (128*NCTA x BLOCK_N) tiles; every later pair of virtual
  // tiles is one tile of the last, partially filled round cut into two BLOCK_N/2-wide halves, so that the round
  // costs half a tile time when at most half of the CTA pairs would have had work (b_half: the weight tensor map
  // then holds half-height boxes and a whole tile issues two of them).
  int t_full, num_vtiles, b_half;
  // Row-window mode (conv2.cu, template flag ROW): a CTA tile is ONE segment of an output row (row_wb <= 128 pixels,
  // row_nblk segments per row), the weights of the whole layer stay resident in shared memory and every filter ROW is
  // loaded once as a window of row_wb + ksize_w - 1 input pixels whose ksize_w column taps are shifted views of the same
  // shared-memory tile (descriptor start + tap * 128 B).
  int row_mode, row_wb, row_nblk, row_total;   // row_total = batch * h_out * row_nblk segments
  int row_h_out, row_bo;
  int tiles_n_sh, row_nblk_sh;                 // log2 when tiles_n / row_nblk is a power of two (the usual case), else -1                       // row_bo: 1 = also set the descriptor's base_offset field to the tap
  // head conv + anchor decode (yolo_conv_desc::decode_mode): y = candidate rows
  int dec_mode, dec_nc, dec_S, dec_rpi, dec_off;
  float dec_inv_s, dec_anchors[6];
  // mc = 2: clusters of two CTA pairs that work on neighbouring M tiles of the same N tile and share the weight tile
  // through TMA multicast (each CTA loads a quarter of it for both pairs): 25 % less L2 -> SM traffic on the big layers
  int mc;
  int pdl;              // launched with programmatic stream serialization: griddepcontrol.* brackets the prologue
  unsigned long long* trace;  // dev tool (yolo_conv_fwd_trace): 32 globaltimer stamps / counters per CTA, nullptr otherwise
  int trace_box;              // which epilogue box of the first tile gets the fine-grained stamps
};

struct ConvPlan {
  ConvKParams kp;
  ConvKParams2 kp2;
  int impl, ncta, grid2, stem_direct, csize;
  void* enc_tiled;   // cuTensorMapEncodeTiled (the stem's image map is encoded per launch)
  const void* w;
  yolo_conv_desc d;
  int block_n, kc, grid_x, grid_y, smem_bytes;
  uint32_t magic;
};


typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// conv2.cu
int conv2_plan_setup(ConvPlan* pl, const yolo_conv_desc* d, int h_out, int w_out, int im2col, PFN_encodeTiled encTiled,
                     const void* x, const void* residual, void* y);
int conv2_launch(const ConvPlan* pl, uint32_t* status, cudaStream_t stream, double* stats = nullptr,
                 const BnFinalize* fin = nullptr, unsigned int* fin_counter = nullptr,
                 unsigned long long* trace = nullptr, int trace_box = 0);
int conv2_query_max_clusters(int cluster, int* out);
int conv2_launch_stem(const ConvPlan* pl, const float* x_nchw, uint32_t* status, cudaStream_t stream);
