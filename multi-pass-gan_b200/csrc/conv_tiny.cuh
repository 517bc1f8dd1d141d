// CUDA-core convolution for the last, almost channel-less layers of a generator (Cout <= 2, Cin <= 8, 16-bit NHWC
// input at pixel stride 8): ru3 of gen_resnet (GAN/multipassGAN-4x.py:564: 8->2 and 2(+8 shortcut)->1).
//
// Why not the tensor cores: a 5x5 2->1 conv is 50 MAC per pixel. On the tap-folded tcgen05 kernel the layer is
// bound by the per-tile role hand-overs (0.13 ms for 8x512^2), on CUDA cores it is a bandwidth kernel
// (reads 16 B + 16 B per pixel, writes 2-16 B): HBM roofline, not tensor roofline.
//
// Layout: block = 256 threads = 64x16 output pixels, thread = 4 horizontally adjacent pixels. The (64+k-1) x (16+k-1)
// input window is converted to fp32 ONCE while it is staged into shared memory as channel planes [ci][y][x], so the
// inner loop is two conflict-free LDS.128 per (ci, dy) feeding 4*k*Cout FMAs whose weights are immediate
// constant-bank operands (weights live in the kernel parameter block, loops fully unrolled).
#pragma once
#include "common.h"

namespace mpg {

constexpr int kTinyTileW = 64;
constexpr int kTinyTileH = 16;
constexpr int kTinyMaxW0 = 5 * 5 * 8 * 2;
constexpr int kTinyMaxW1 = 8 * 2;

struct TinyParams {
  int n, h, w;
  int act;
  int in_dtype, out_dtype, out_cstride;
  const void* x0;  // [n,h,w,8] 16-bit, main segment (k x k)
  const void* x1;  // [n,h,w,8] 16-bit, 1x1 shortcut segment (or null)
  void* out;
  float shift[2];
  float w0[kTinyMaxW0];  // [dy][dx][ci][co], scale folded
  float w1[kTinyMaxW1];  // [ci][co]
};

bool tiny_eligible(const mpg_conv_desc& d);
// returns cudaError_t as int
int tiny_launch(const mpg_conv_desc& d, const TinyParams& p, cudaStream_t stream);

}  // namespace mpg
