// Tap-folded implicit-GEMM convolution for NARROW output-channel counts (tcgen05 + TMEM + TMA).
//
// Why: an SS-mode tcgen05.mma spends >= 64 cycles streaming its 128 x 16 A block out of shared memory
// whatever N is, so the tap-by-tap kernel (conv_igemm.cu: N = Cout) runs a Cout = 32 layer at 25 % and a
// Cout = 8 layer at 6 % of the tensor pipe.  Here the k horizontal taps become extra GEMM columns:
//
//     P[y][x][dx*Cp + co] = sum_{dy,ci} X[y+dy-pad][x][ci] * W[dy][dx][ci][co]      (N = k*Cp per MMA)
//     out[y][x][co]       = act( sum_dx P[y][x+dx-pad][dx*Cp + co] + shift[co] )    (epilogue)
//
// so one A block feeds k taps at once (5x fewer MMAs for a 5x5 conv).  The shifted sum over dx runs in the
// epilogue: an accumulator holds 4 image rows x 32 pixels (lane = row*32 + x), each epilogue warp owns one
// image row, and the +-pad pixel shifts are warp shuffles.  A 32-pixel window therefore yields 32-(k-1)
// output pixels; neighbouring tiles overlap by k-1 columns.
//   - A operand: per (segment, Cin chunk) ONE TMA box [R+k-1 rows][32 px][CK]; vertical taps = UMMA
//     descriptor start advanced by whole image rows (swizzle-atom aligned).  Zero fill = SAME padding.
//   - B operand: packed weights [k-tile=(seg,chunk,dy)][Npad][CK]; resident in smem when small.
//   - optional 1x1 shortcut segment: extra K-slabs whose weight tile is non-zero only in the centre-dx
//     column block.
#pragma once
#include "common.h"

namespace mpg {

constexpr int kNfWin = 32;      // pixels per image row held by an accumulator
constexpr int kNfRowsAcc = 4;   // image rows per accumulator (4 * 32 = 128 MMA rows)
constexpr int kNfThreads = 256;     // 4 role warps + 4 epilogue warps (two CTAs per SM)
constexpr int kNfMaxThreads = 384;  // + 8 epilogue warps (one CTA per SM)
constexpr int kNfMaxStagesA = 8;
constexpr int kNfMaxStagesB = 8;
constexpr int kNfMaxBufs = 10;  // TMEM accumulator buffers (tiles in flight between the MMA and epilogue warps)

struct NfoldParams {
  int n, h, w;
  int tiles_x, tiles_y, num_tiles;
  int valid_w;  // 32 - (ks0 - 1)
  int rows;     // output rows per tile = naccs * 4
  int naccs, nbuf;
  int nseg;
  int seg_ks[2];
  int seg_nchunk[2];
  int npad;  // UMMA N = round_up(ks0 * cp, 16)
  int cp;    // cout rounded up to 8
  int cout;
  int act, pixel_norm;
  int in_dtype, out_dtype, out_cstride;
  int na, nb;
  int a_stage_bytes, b_tile_bytes;
  int bres, ktiles;
  int a_cpasync;  // window images staged by producer warps with cp.async instead of tiled TMA (conv_nfold.cu)
  int nprod;      // ... number of producer warps: warps 0 and 3 plus (nprod - 2) extra warps after the epilogue warps
  int seg_cin[2], seg_cstride[2];  // for the cp.async producer: real channels / channel stride of each input
  const void* x[2];                // ... and the input tensors themselves
  const float* resid;  // optional fp32 [n,h,w,8] tensor added before the activation (cout <= 8): the block's 1x1 shortcut
                       // computed by the producer of its input (mpg_conv_plan_set_side)
  int pair;  // cta_group::2 CTA pairs with resident half weight tiles, one accumulator per tile (conv_nfold.cu)
  int threads;  // launch block size: 256 or 384
  int dbg;  // profiling only (env MPG_NFOLD_DBG): bit0 skip stores, bit1 skip the whole epilogue body, bit2 skip MMAs
  uint32_t tmem_cols;
  const float* shift;  // [cp] device
  const void* wpacked;  // device: weight tiles in their swizzled smem image, b_tile_bytes apart
  void* out;
};

int nfold_launch(int ck, const CUtensorMap& tm_x0, const CUtensorMap& tm_x1, const CUtensorMap& tm_w,
                 const NfoldParams& p, int grid, size_t smem_bytes, cudaStream_t stream);
int nfold_set_smem_attr(int device, int ck, int ks, int pair, size_t smem_bytes);

}  // namespace mpg
