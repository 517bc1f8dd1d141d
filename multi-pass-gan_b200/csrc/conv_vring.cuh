// Row-streaming tcgen05 convolution whose sum over the vertical taps is accumulated IN TENSOR MEMORY (a ring of output rows).
//
// conv_vfold.cu folds the k vertical taps into N and sums the k column blocks in the epilogue; that epilogue reads k times
// more TMEM than it stores, and TMEM reads run at ~40-64 B/clk/SM, so layers whose rows carry few K-steps (Cin <= 64) end up
// bound by the epilogue instead of the tensor pipe. Here the MMA itself does the sum:
//
//   * TMEM holds a RING of R = 512 / cs output-row accumulators ("slots": 128 pixels x cs columns, cs = round_up(Cout, 16));
//     output row o lives in slot (o + k - 1) mod R;
//   * image row r (N = k * cs columns, column block j <-> vertical tap dy = k-1-j) is accumulated straight onto the k
//     consecutive slots r .. r+k-1: ONE MMA per (dx, K-step) of N = k*cs columns, or two narrower ones when the k slots wrap
//     around the end of the ring (B descriptor advanced by whole column blocks);
//   * a slot is complete once the image row with its last tap has been issued; the epilogue reads just those cs columns
//     (k times less TMEM traffic, no adds), writes zeros back (tcgen05.st) and hands the slot to the MMA warp again. Every MMA
//     accumulates; only the first image row of a row range overwrites (accumulate = 0), which also discards the partial sums a
//     previous range left in the ring;
//   * single CTA per SM (cta_group::1): the narrow wrap MMAs need the whole B tile, which a CTA pair splits in halves. The
//     weights stay resident as [dx][K-step] tiles of N x 16 channels (32-byte rows, SWIZZLE_32B), <= ~150 KB.
// A operand, strips, flattened (image, strip, row) work ranges and SAME padding by TMA zero fill are as in conv_vfold.cuh.
#pragma once
#include "common.h"

namespace mpg {

constexpr int kVrStrip = 128;
constexpr int kVrMaxStagesA = 8;
constexpr int kVrMaxSlots = 32;

struct VringParams {
  int n, h, w;
  int strips;        // 128-pixel strips per image row
  int total_rows;    // n * strips * h
  int rows_per_cta;  // contiguous (image, strip, row) units per CTA
  int ks;
  int nseg;
  int seg_nchunk[2];  // Cin chunks of CK channels (one TMA box each)
  int seg_klast[2];   // K=16 steps of the last chunk
  int cs;             // slot width in columns = round_up(cout, 16)
  int nslots;         // R
  int cp, cout;
  int act, pixel_norm;
  int in_dtype, out_dtype, out_cstride;
  int na, a_stage_bytes;
  int b_tile_bytes;     // one (K-step, dx) tile: ks * cs rows of 32 bytes
  int b_sc_tile_bytes;  // one shortcut K-step tile: cs rows of 32 bytes
  int b_bytes;
  int epi_groups;
  int dbg;  // profiling only (env MPG_VRING_DBG): bit0 skip stores, bit1 skip the epilogue math, bit2 skip MMAs
  const float* shift;
  const void* wpacked;
  const float* resid;  // optional fp32 [n,h,w,8] added before the activation (cout <= 8)
  void* out;
};

int vring_launch(int ck, int nchw, const CUtensorMap& tm_x0, const CUtensorMap& tm_x1, const VringParams& p, int grid,
                 size_t smem_bytes, cudaStream_t stream);
int vring_set_smem_attr(int device, int ck, int ks, int nchw, int groups, size_t smem_bytes);

}  // namespace mpg
