// CUDA-core convolution for Cout <= 2 / Cin <= 8 layers (design: conv_tiny.cuh).
#include "conv_tiny.cuh"

namespace mpg {

namespace {

template <int KS, int CIN0, int CIN1, int COUT>
__global__ void __launch_bounds__(256) conv_tiny_kernel(const __grid_constant__ TinyParams p) {
  constexpr int PAD = KS / 2;
  constexpr int WH = kTinyTileH + KS - 1;
  constexpr int WW = kTinyTileW + KS - 1;
  constexpr int PITCH = kTinyTileW + 8;  // multiple of 4 floats: every thread's window starts 16-byte aligned
  __shared__ __align__(16) float sm[CIN0][WH][PITCH];

  const int x0 = blockIdx.x * kTinyTileW, y0 = blockIdx.y * kTinyTileH, n = blockIdx.z;
  const int idt = p.in_dtype;

  // ---- stage the input window as fp32 channel planes (zero outside the image = SAME padding, GAN.py:691)
  {
    const uint8_t* base = static_cast<const uint8_t*>(p.x0) + static_cast<size_t>(n) * p.h * p.w * 16;
    for (int i = threadIdx.x; i < WH * WW; i += 256) {
      const int wy = i / WW, wx = i - wy * WW;
      const int gy = y0 - PAD + wy, gx = x0 - PAD + wx;
      uint32_t q[4] = {0u, 0u, 0u, 0u};
      if (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) {
        const uint8_t* src = base + (static_cast<size_t>(gy) * p.w + gx) * 16;
        if (CIN0 <= 2) {
          q[0] = __ldg(reinterpret_cast<const uint32_t*>(src));
        } else if (CIN0 <= 4) {
          const uint2 t = __ldg(reinterpret_cast<const uint2*>(src));
          q[0] = t.x;
          q[1] = t.y;
        } else {
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(src));
          q[0] = t.x;
          q[1] = t.y;
          q[2] = t.z;
          q[3] = t.w;
        }
      }
#pragma unroll
      for (int c = 0; c < CIN0; ++c)
        sm[c][wy][wx] = h16_to_float(static_cast<uint16_t>(q[c >> 1] >> ((c & 1) * 16)), idt);
    }
  }
  __syncthreads();

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][COUT];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[j][co] = p.shift[co];

#pragma unroll
  for (int c = 0; c < CIN0; ++c) {
#pragma unroll
    for (int dy = 0; dy < KS; ++dy) {
      const float4 a = *reinterpret_cast<const float4*>(&sm[c][ty + dy][4 * tx]);
      const float4 b = *reinterpret_cast<const float4*>(&sm[c][ty + dy][4 * tx + 4]);
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int dx = 0; dx < KS; ++dx)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int co = 0; co < COUT; ++co)
            acc[j][co] = fmaf(v[j + dx], p.w0[((dy * KS + dx) * CIN0 + c) * COUT + co], acc[j][co]);
    }
  }

  const int gy = y0 + ty;
  const int gx = x0 + 4 * tx;
  if (gy >= p.h || gx >= p.w) return;
  const size_t pix0 = (static_cast<size_t>(n) * p.h + gy) * p.w + gx;

  if (CIN1 > 0) {  // 1x1 shortcut of the resBlock (GAN/multipassGAN-4x.py:521), read straight from global memory
    const uint4* s = reinterpret_cast<const uint4*>(p.x1) + pix0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (gx + j < p.w) {
        const uint4 t = __ldg(s + j);
        const uint32_t q[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int c = 0; c < CIN1; ++c) {
          const float xv = h16_to_float(static_cast<uint16_t>(q[c >> 1] >> ((c & 1) * 16)), idt);
#pragma unroll
          for (int co = 0; co < COUT; ++co) acc[j][co] = fmaf(xv, p.w1[c * COUT + co], acc[j][co]);
        }
      }
    }
  }

  const float ca = p.act == MPG_ACT_RELU ? 0.5f : (p.act == MPG_ACT_LRELU ? 0.6f : 1.0f);
  const float cb = p.act == MPG_ACT_RELU ? 0.5f : (p.act == MPG_ACT_LRELU ? 0.4f : 0.0f);
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      const float x = acc[j][co];
      acc[j][co] = p.act == MPG_ACT_TANH ? tanhf(x) : fmaf(cb, fabsf(x), ca * x);
    }

  if (p.out_dtype == MPG_F32) {
    float* o = static_cast<float*>(p.out) + pix0 * p.out_cstride;
    if (p.out_cstride == 1 && gx + 3 < p.w && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
      *reinterpret_cast<float4*>(o) = make_float4(acc[0][0], acc[1][0], acc[2][0], acc[3][0]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gx + j < p.w)
          for (int c = 0; c < p.out_cstride; ++c)
            o[j * p.out_cstride + c] = (c == 0) ? acc[j][0] : ((COUT > 1 && c == 1) ? acc[j][COUT - 1] : 0.0f);
    }
  } else {
    uint16_t* o = static_cast<uint16_t*>(p.out) + pix0 * p.out_cstride;
    const int od = p.out_dtype;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (gx + j < p.w) {
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        q.x = pack_h16x2(acc[j][0], COUT > 1 ? acc[j][COUT - 1] : 0.0f, od);
        uint4* op = reinterpret_cast<uint4*>(o + j * p.out_cstride);
        op[0] = q;
        for (int c = 8; c < p.out_cstride; c += 8) op[c >> 3] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
}

template <int KS, int CIN0, int CIN1, int COUT>
int launch_t(const TinyParams& p, cudaStream_t st) {
  dim3 grid(ceil_div(p.w, kTinyTileW), ceil_div(p.h, kTinyTileH), p.n);
  conv_tiny_kernel<KS, CIN0, CIN1, COUT><<<grid, 256, 0, st>>>(p);
  return static_cast<int>(cudaGetLastError());
}

template <int KS, int CIN0>
int launch_k(const mpg_conv_desc& d, const TinyParams& p, cudaStream_t st) {
  const bool sc = d.nseg == 2;
  if (d.cout == 1) return sc ? launch_t<KS, CIN0, 8, 1>(p, st) : launch_t<KS, CIN0, 0, 1>(p, st);
  return sc ? launch_t<KS, CIN0, 8, 2>(p, st) : launch_t<KS, CIN0, 0, 2>(p, st);
}

int cin_bucket(int cin) { return cin <= 2 ? 2 : (cin <= 4 ? 4 : 8); }

}  // namespace

bool tiny_eligible(const mpg_conv_desc& d) {
  if (!is_h16(d.in_dtype) || d.stride != 1 || d.upsample != 1 || d.in_upsample != 1 || d.pixel_norm) return false;
  if (d.cout > 2 || d.out_cstride < d.cout) return false;
  if (d.out_dtype != MPG_F32 && (d.out_dtype != d.in_dtype || d.out_cstride % 8 != 0)) return false;
  if (d.seg_ksize[0] != 3 && d.seg_ksize[0] != 5) return false;
  if (d.seg_cin[0] > 8 || d.seg_cstride[0] != 8) return false;
  if (d.nseg == 2 && (d.seg_ksize[1] != 1 || d.seg_cin[1] > 8 || d.seg_cstride[1] != 8)) return false;
  return true;
}

int tiny_launch(const mpg_conv_desc& d, const TinyParams& p, cudaStream_t st) {
  const int cb = cin_bucket(d.seg_cin[0]);
  if (d.seg_ksize[0] == 5) return cb == 2 ? launch_k<5, 2>(d, p, st) : (cb == 4 ? launch_k<5, 4>(d, p, st) : launch_k<5, 8>(d, p, st));
  return cb == 2 ? launch_k<3, 2>(d, p, st) : (cb == 4 ? launch_k<3, 4>(d, p, st) : launch_k<3, 8>(d, p, st));
}

}  // namespace mpg
