// Row-streaming tcgen05 convolution with the VERTICAL taps folded into the GEMM N dimension (narrow / medium Cout).
//
// Why: an SS-mode tcgen05.mma spends >= 64 cycles streaming its 128 x 16 A block out of shared memory whatever N
// is, so the tap-by-tap kernel (conv_igemm.cu, N = Cout) runs a Cout = 48 layer at <= 37 % of the tensor pipe. The
// horizontal fold of conv_nfold.cu fixes the MMA count but pays for it elsewhere: its epilogue needs k warp shuffles per
// output value (lanes = pixels along x, and the fold shifts along x), a 32-pixel window yields only 32-(k-1) outputs, and a
// 4-row tile re-loads 4 + k - 1 window rows (2x read amplification, tiled-TMA row-rate bound). Here the fold is vertical:
//
//     P[r][x][dy*Cp + co] = sum_{dx,ci} X[r][x+dx-pad][ci] * W[dy][dx][ci][co]          (one image row r, N = k*Cp)
//     out[y][x][co]       = act( sum_dy P[y+dy-pad][x][dy*Cp + co] + shift[co] )        (epilogue, SAME lane)
//
//   * a CTA owns a 128-pixel-wide column strip and STREAMS image rows through it: every input row is loaded exactly once
//     (one TMA box [1 row][128+k-1 px][CK] per Cin chunk, zero fill = SAME padding), the k horizontal taps are the same
//     staged row addressed through UMMA descriptors whose start is shifted by whole pixels;
//   * an accumulator = the 128 pixels of ONE image row (TMEM lane = x), so the sum over dy combines values of the SAME
//     lane from k consecutive accumulators: each epilogue thread keeps k-1 running partial rows in registers
//     (S[j] = S[j-1] + P[j]: one FADD per value, no shuffles, no shared memory);
//   * CTA pairs (cta_group::2): the two CTAs take horizontally adjacent strips and the same rows; each keeps HALF of every
//     weight tile resident in shared memory for its whole lifetime (loaded once with 1-D bulk copies), the leader issues
//     M = 256 MMAs. Accumulators rotate through TMEM so the epilogue of row i overlaps the MMAs of rows i+1...;
//   * optional 1x1 shortcut segment: extra K-slabs issued as a narrow MMA into the centre-dy column block;
//   * work = the flattened (image, strip pair, row) space cut into equal contiguous ranges, one per CTA pair; a range that
//     crosses an image / strip boundary restarts its running sums (k-1 extra rows).
#pragma once
#include "common.h"

namespace mpg {

constexpr int kVfStrip = 128;      // output pixels per CTA and image row (= MMA rows of one CTA)
constexpr int kVfMaxStagesA = 8;
constexpr int kVfMaxBufs = 4;      // TMEM accumulators in flight between the MMA and epilogue warps

struct VfoldParams {
  int n, h, w;
  int strips2;        // strip pairs per image row: ceil(ceil(w / 128) / 2)
  int total_rows;     // n * strips2 * h: flattened row units (one unit = one image row of a strip pair)
  int rows_per_pair;  // contiguous row units per CTA pair
  int ks;             // kernel size of the main segment (3 or 5)
  int nseg;
  int seg_nchunk[2];  // Cin chunks of CK channels
  int seg_klast[2];   // K=16 steps of the last chunk (the others run CK/16)
  int npad;           // UMMA N of the main MMAs = round_up(ks * cp, 16)
  int cp, cout;       // cp = cout rounded up to 8
  int n_sc, sc_col;   // shortcut MMA: N and first accumulator column
  int act, pixel_norm;
  int in_dtype, out_dtype, out_cstride;
  int na, a_stage_bytes;
  int b_tile_bytes;     // this CTA's half of a main weight tile (1024-byte multiple)
  int b_sc_tile_bytes;  // ... of a shortcut weight tile
  int b_bytes;          // resident weight bytes per CTA
  int nbuf;
  int epi_groups;       // epilogue warps per TMEM lane quarter (block = 128 + 128 * epi_groups threads)
  int dbg;              // profiling only (env MPG_VFOLD_DBG): bit0 skip stores, bit1 skip the epilogue body, bit2 skip MMAs
  uint32_t tmem_cols;
  const float* shift;   // [cp] device
  const void* wpacked;  // device: [rank][resident image of b_bytes]
  const float* resid;   // optional fp32 [n,h,w,8] added before the activation (cout <= 8), see mpg_conv_plan_set_side
  void* out;
};

int vfold_launch(int ck, int nchw, const CUtensorMap& tm_x0, const CUtensorMap& tm_x1, const VfoldParams& p, int grid,
                 size_t smem_bytes, cudaStream_t stream);
int vfold_set_smem_attr(int device, int ck, int ks, int nchw, int groups, size_t smem_bytes);

}  // namespace mpg
