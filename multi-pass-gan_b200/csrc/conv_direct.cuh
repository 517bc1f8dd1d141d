// CUDA-core direct convolution launcher (see conv_direct.cu).
#pragma once
#include "common.h"

namespace mpg {

struct DirectParams {
  int n;
  int src_h, src_w;  // spatial size of the input tensors in memory
  int in_upsample;   // the conv sees a nearest-upsampled view: h = src_h * in_upsample
  int h, w;          // virtual input size
  int oh, ow;        // conv output size = ceil(h / stride)
  int stride;
  int nseg;
  int seg_ks[2], seg_cin[2], seg_cstride[2], seg_pad[2];
  long long seg_woff[2];  // offset (floats) of each segment in w
  const void* x[2];
  const float* wts;    // [seg][tap][cin][coutp] fp32, scale folded
  const float* shift;  // [coutp]
  int cout, coutp;     // coutp = round_up(cout, 8)
  int act, pixel_norm, upsample;
  int in_dtype, out_dtype, out_cstride;
  void* out;
};

int direct_launch(const DirectParams& p, cudaStream_t stream);

}  // namespace mpg
