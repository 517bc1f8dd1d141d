// CUDA-core direct convolution (fp32 FMA accumulate).
//
// Used for (a) the thin-channel edge layers of the generators (Cin or Cout < 8: HBM-bound, SURVEY
// App. A.7 last rows), (b) the fp32 activation path that is held to <= 1e-4 against the oracle,
// (c) strided / even-kernel layers of the discriminator (tools_wscale/GAN.py:686-691 with
// stride 2, k=4).  SAME padding follows TensorFlow: pad_before = floor(pad_total/2).
#include "conv_direct.cuh"

namespace mpg {

namespace {

__device__ __forceinline__ float act_f(float v, int act) {
  switch (act) {
    case MPG_ACT_RELU:
      return fmaxf(v, 0.0f);
    case MPG_ACT_LRELU:
      return 0.6f * v + 0.4f * fabsf(v);
    case MPG_ACT_TANH:
      return tanhf(v);
    default:
      return v;
  }
}

template <typename T>
__device__ __forceinline__ float ld_scalar(const T* p);
template <>
__device__ __forceinline__ float ld_scalar<float>(const float* p) {
  return __ldg(p);
}
template <>
__device__ __forceinline__ float ld_scalar<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <>
__device__ __forceinline__ float ld_scalar<__half>(const __half* p) {
  return __half2float(*p);
}

// 8 output channels per thread
template <typename TIn>
__global__ void __launch_bounds__(128)
conv_direct_kernel(const DirectParams p) {
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long npix = static_cast<long long>(p.n) * p.oh * p.ow;
  if (pix >= npix) return;
  const int g = blockIdx.y;  // output channel group
  const int ox = static_cast<int>(pix % p.ow);
  const long long r = pix / p.ow;
  const int oy = static_cast<int>(r % p.oh);
  const int n = static_cast<int>(r / p.oh);

  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = p.shift[g * 8 + j];

  for (int s = 0; s < p.nseg; ++s) {
    const TIn* xs = reinterpret_cast<const TIn*>(p.x[s]);
    const int ks = p.seg_ks[s];
    const int cin = p.seg_cin[s];
    const int cstr = p.seg_cstride[s];
    const int pad = p.seg_pad[s];
    const float* ws = p.wts + p.seg_woff[s];
    for (int dy = 0; dy < ks; ++dy) {
      const int iy = oy * p.stride + dy - pad;
      if (iy < 0 || iy >= p.h) continue;
      for (int dx = 0; dx < ks; ++dx) {
        const int ix = ox * p.stride + dx - pad;
        if (ix < 0 || ix >= p.w) continue;
        const TIn* xp = xs + ((static_cast<long long>(n) * p.src_h + iy / p.in_upsample) * p.src_w +
                              ix / p.in_upsample) * cstr;
        const float* wp = ws + (static_cast<long long>(dy * ks + dx) * cin) * p.coutp + g * 8;
        for (int ci = 0; ci < cin; ++ci) {
          const float xv = ld_scalar<TIn>(xp + ci);
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
          acc[0] = fmaf(xv, w0.x, acc[0]);
          acc[1] = fmaf(xv, w0.y, acc[1]);
          acc[2] = fmaf(xv, w0.z, acc[2]);
          acc[3] = fmaf(xv, w0.w, acc[3]);
          acc[4] = fmaf(xv, w1.x, acc[4]);
          acc[5] = fmaf(xv, w1.y, acc[5]);
          acc[6] = fmaf(xv, w1.z, acc[6]);
          acc[7] = fmaf(xv, w1.w, acc[7]);
          wp += p.coutp;
        }
      }
    }
  }

  const int ups = p.upsample;
  const int fh = p.oh * ups, fw = p.ow * ups;
  for (int uy = 0; uy < ups; ++uy) {
    for (int ux = 0; ux < ups; ++ux) {
      const long long opix = (static_cast<long long>(n) * fh + (oy * ups + uy)) * fw + (ox * ups + ux);
      if (p.out_dtype != MPG_F32) {
        uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + opix * p.out_cstride + g * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (g * 8 + j < p.out_cstride)
            o[j] = float_to_h16((g * 8 + j < p.cout) ? act_f(acc[j], p.act) : 0.0f, p.out_dtype);
      } else {
        float* o = reinterpret_cast<float*>(p.out) + opix * p.out_cstride + g * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (g * 8 + j < p.out_cstride) o[j] = (g * 8 + j < p.cout) ? act_f(acc[j], p.act) : 0.0f;
      }
    }
  }
}

// in-place x * rsqrt(mean_c(x^2) + 1e-8)  (tools_wscale/GAN.py:472-474); one warp per pixel
__global__ void pixel_norm_kernel(void* x, int dtype, long long npix, int c, int cstride) {
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= npix) return;
  float* pf = reinterpret_cast<float*>(x) + warp * cstride;
  uint16_t* ph = reinterpret_cast<uint16_t*>(x) + warp * cstride;
  float ssq = 0.0f;
  for (int i = lane; i < c; i += 32) {
    const float v = dtype == MPG_F32 ? pf[i] : h16_to_float(ph[i], dtype);
    ssq = fmaf(v, v, ssq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
  const float rn = rsqrtf(ssq / static_cast<float>(c) + 1e-8f);
  for (int i = lane; i < c; i += 32) {
    if (dtype == MPG_F32)
      pf[i] = pf[i] * rn;
    else
      ph[i] = float_to_h16(h16_to_float(ph[i], dtype) * rn, dtype);
  }
}

}  // namespace

int direct_launch(const DirectParams& p, cudaStream_t stream) {
  const long long npix = static_cast<long long>(p.n) * p.oh * p.ow;
  dim3 grid(static_cast<unsigned>((npix + 127) / 128), static_cast<unsigned>(p.coutp / 8));
  if (p.in_dtype == MPG_F32)
    conv_direct_kernel<float><<<grid, 128, 0, stream>>>(p);
  else if (p.in_dtype == MPG_F16)
    conv_direct_kernel<__half><<<grid, 128, 0, stream>>>(p);
  else
    conv_direct_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  if (p.pixel_norm) {
    const long long fpix = npix * p.upsample * p.upsample;
    const unsigned blocks = static_cast<unsigned>((fpix * 32 + 255) / 256);
    pixel_norm_kernel<<<blocks, 256, 0, stream>>>(p.out, p.out_dtype, fpix, p.cout, p.out_cstride);
    e = cudaGetLastError();
  }
  return static_cast<int>(e);
}

}  // namespace mpg
