// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
// GEMM view: M = output pixels, N = Cout, K = taps x Cin.
//   - persistent, warp-specialised CTA (one or two per SM) over 16x16-pixel output tiles = 2 accumulators of M=128
//     (16 image rows x 8 pixels each); TMEM double buffered (4 accumulators) so the epilogue of tile i overlaps the
//     MMAs of tile i+1
//   - A operand: for every (segment, Cin chunk) ONE TMA halo image [16+k-1][16+k-1][CK] lands in swizzled shared
//     memory; every (dy,dx) tap is the same image addressed through a UMMA descriptor whose start is shifted by whole
//     pixels and whose 8-row-group stride is the halo pitch. Out-of-image coordinates are zero filled by TMA, which IS
//     the "SAME" zero padding of tf.nn.conv2d (tools_wscale/GAN.py:691)
//   - B operand: packed weights [k-tile][Npad][CK]; streamed in stages of k vertical taps, or resident for the CTA's
//     lifetime when they fit; cta_group::2 CTA pairs hold half of every tile each and issue M=256 MMAs
//   - epilogue: tcgen05.ld -> +shift -> act -> pixel_norm -> 16-bit / fp32 store (per-warp TMA store for 64-channel
//     groups, direct 32-byte stores otherwise; optionally x2 nearest replicated), all fp32
#pragma once
#include "common.h"

namespace mpg {

constexpr int kIgTileW = 16;
constexpr int kIgTileH = 16;
constexpr int kIgThreads = 256;     // 4 role warps + 4 epilogue warps
constexpr int kIgMaxThreads = 384;  // ... or + 8 epilogue warps (one CTA per SM plans)
constexpr int kIgMaxStagesA = 4;
constexpr int kIgMaxStagesB = 12;

struct IgemmParams {
  int n, h, w;
  int tiles_x, tiles_y, num_tiles;
  int nseg;
  int seg_ks[2];
  int seg_nchunk[2];
  int npad;  // UMMA N
  int cout;
  int act, pixel_norm, upsample;
  int in_dtype;  // MPG_BF16 or MPG_F16 operands
  int out_dtype, out_cstride;
  int na, nb;  // pipeline depth of the A / B rings
  int a_stage_bytes, b_stage_bytes;
  int b_tile_bytes;  // one (seg,chunk,dx,dy) weight tile
  int bgroup;        // 1: a B stage holds all ks dy-taps of a (chunk,dx) (same cadence as A); 0: one tap
  // per-warp TMA-store epilogue (16-bit outputs made of 64-channel groups): 4 KB of swizzled staging per epilogue warp
  // at stage_off; stage_bytes = total staging
  int tma_store, stage_off, stage_bytes;
  int threads;  // 256 or 384 (launch block size)
  int pair;  // cta_group::2 CTA pairs: each CTA holds half of every weight tile (wide, weight-streaming layers)
  // resident weights: all `ktiles` weight tiles live in smem for the CTA's lifetime (thin layers)
  int bres, ktiles;
  int dbg;  // profiling only (env MPG_IGEMM_DBG): bit0 skip global stores, bit1 skip the TMEM loads too
  uint32_t tmem_cols;
  const float* shift;  // [npad] device
  void* out;
  // side output (mpg_conv_plan_set_side): side_out[pix][0..7] = sum_c y[pix][c] * side_w[c][0..7], the 1x1 shortcut of the NEXT
  // residual block computed from the fp32 epilogue values while they are in registers (TMA-store epilogue, 8 epilogue warps).
  // side_off: dynamic smem offset of 4 KB weights [128][8] + 4 KB partial-sum exchange
  int side, side_off;
  const float* side_w;  // device [128][8] fp32
  float* side_out;      // [n,h,w,8] fp32
};

// ck in {16, 32, 64}; returns cudaError_t as int
int igemm_launch(int ck, const CUtensorMap& tm_x0, const CUtensorMap& tm_x1, const CUtensorMap& tm_w,
                 const CUtensorMap& tm_y, const IgemmParams& p, int grid, size_t smem_bytes, cudaStream_t stream);
int igemm_set_smem_attr(int device, int ck, int pair, size_t smem_bytes, int side = 0);

}  // namespace mpg
