// Row-streaming tcgen05 convolution whose sum over the vertical taps is accumulated IN TENSOR MEMORY (a ring of output rows).
//
// conv_vfold.cu folds the k vertical taps into N and sums the k column blocks in the epilogue; that epilogue reads k times
// more TMEM than it stores, and TMEM reads run at ~40-64 B/clk/SM, so layers whose rows carry few K-steps (Cin <= 64) end up
// bound by the epilogue instead of the tensor pipe. Here the MMA itself does the sum:
//
//   * TMEM holds a RING of R = 512 / cs output-row accumulators ("slots": 128 pixels x cs columns, cs = round_up(Cout, 16));
//     output row o lives in slot (o + k - 1) mod R;
//   * image row r (N = k * cs columns, column block j <-> vertical tap dy = k-1-j) is accumulated straight onto the k
//     consecutive slots r .. r+k-1 by ONE MMA per (dx, K-step). The ring is followed by k-1 OVERFLOW slots: when the k slots
//     would wrap around the end of the ring the MMA simply runs on into them (overflow slot R+i aliases ring slot i), so an
//     MMA never has to be split -- which is what lets CTA pairs, whose B halves fit exactly one N, use the scheme;
//   * a slot is complete once the image row with its last tap has been issued; the epilogue reads just those cs columns (plus
//     the alias for the first k-1 slots: k times less TMEM traffic than the fold), writes zeros back (tcgen05.st) and hands
//     the slot to the MMA warp again. Every MMA accumulates. The partial sums a row range leaves behind in the ring sit in
//     slots whose outputs the next range discards (the k-1 rows above its first output row) and are cleared with them;
//   * single CTA per SM for the thin layers (no cross-CTA hand-overs: a row is only a few hundred MMA cycles), CTA pairs
//     (cta_group::2, adjacent strips, HALF of every weight tile per CTA) for wide N: an SS-mode MMA fetches its operands at
//     ~64 B/clk per SM, so a single CTA needs (4 KB + N * 32 B) / 64 cycles per MMA -- 184 for N = 240 against 120 tensor
//     cycles -- while a pair fetches only N/2 weight rows per CTA. Weights stay resident as [K-step][dx] tiles of N x 16
//     channels (32-byte rows, SWIZZLE_32B).
// A operand, strips, flattened (image, strip, row) work ranges and SAME padding by TMA zero fill are as in conv_vfold.cuh.
#pragma once
#include "common.h"

namespace mpg {

constexpr int kVrStrip = 128;
constexpr int kVrMaxStagesA = 8;
constexpr int kVrMaxSlots = 32;

struct VringParams {
  int n, h, w;
  int pair;          // cta_group::2 CTA pairs on adjacent strips, half of every weight tile per CTA
  int units_x;       // 128-pixel strips (or strip pairs) per image row
  int total_rows;    // n * units_x * h
  int rows_per_cta;  // contiguous (image, strip [pair], row) units per CTA (pair)
  int ks;
  int nseg;
  int seg_nchunk[2];  // Cin chunks of CK channels (one TMA box each)
  int seg_klast[2];   // K=16 steps of the last chunk
  int cs;             // slot width in columns = round_up(cout, 16)
  int nslots;         // R ring slots; k-1 overflow slots follow them in tensor memory ((R + k - 1) * cs <= 512)
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
int vring_set_smem_attr(int device, int ck, int ks, int nchw, int groups, int pair, size_t smem_bytes);

}  // namespace mpg
