// Bandwidth-bound tensor plumbing around the convolutions:
//   mpg_pack_channels   nearest-neighbour resize + channel slice/concat + dtype cast + channel padding
//                       (tf.image.resize_images(...,1) tools_wscale/GAN.py:517,541, GAN/multipassGAN-out.py:357;
//                        tf.concat / tf.slice GAN/multipassGAN-out.py:330-332,357)
//   mpg_dens_residual   out = dens + {channel | TF1 legacy bicubic resize of channel} of the network input
//                       (addBicubicUpsample, GAN/multipassGAN-out.py:327-332 -> tools_wscale/GAN.py:541 mode 2)
#include <string.h>

#include <vector>

#include "common.h"

namespace mpg {
namespace {

constexpr int kMaxSrc = 8;
constexpr int kPackRows = 8;

struct PackSrc {
  const void* ptr;
  int dtype, cstride, c0, nch, fh, fw;
  int sh, sw;    // source image size = out size / factor (host-computed: integer divisions bounded this kernel)
  int lfh, lfw;  // log2 of the factors when they are powers of two, else -1
};
__device__ __forceinline__ int pack_div(int v, int f, int lf) { return lf >= 0 ? (v >> lf) : v / f; }
struct PackParams {
  PackSrc src[kMaxSrc];
  int nsrc;
  int n, oh, ow;
  int out_dtype, out_cstride;
  void* out;
};

__global__ void __launch_bounds__(256) pack_channels_kernel(const PackParams p) {
  // grid = (ceil(ow/256), n*oh): no 64-bit divisions per pixel (they, not the memory system, bounded this kernel)
  const int x = static_cast<int>(blockIdx.x) * 256 + static_cast<int>(threadIdx.x);
  if (x >= p.ow) return;
  // kPackRows image rows per block: one-row blocks finish in well under a microsecond and the kernel was bound by the
  // block launch rate (8192 blocks for 8 x 512^2), not by memory
  const int row0 = static_cast<int>(blockIdx.y) * kPackRows;
  int n = row0 / p.oh;
  int y = row0 - n * p.oh - 1;
  for (int rr = 0; rr < kPackRows; ++rr) {
  const int row = row0 + rr;
  if (row >= p.n * p.oh) return;
  if (++y == p.oh) {
    y = 0;
    ++n;
  }
  const long long pix = static_cast<long long>(row) * p.ow + x;
  int oc = 0;
  if (p.out_dtype != MPG_F32 && (p.out_cstride & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
    // 16-bit output at a 16-byte pixel granularity: gather 8 channels in registers, one 128-bit store per chunk
    // (the scalar 2-byte stores of the generic path ran at a tenth of the HBM rate)
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + pix * p.out_cstride);
    unsigned long long lo = 0ull, hi = 0ull;  // dynamic shifts instead of a dynamically indexed array (stays in registers)
    for (int s = 0; s < p.nsrc; ++s) {
      const PackSrc& q = p.src[s];
      const long long spix = (static_cast<long long>(n) * q.sh + pack_div(y, q.fh, q.lfh)) * q.sw + pack_div(x, q.fw, q.lfw);
      const bool vec4 = q.dtype == MPG_F32 && q.nch == 4 && ((spix * q.cstride + q.c0) & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(q.ptr) & 15) == 0;
      float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vec4) f4 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(q.ptr) + spix * q.cstride + q.c0));
      for (int c = 0; c < q.nch; ++c, ++oc) {
        float v;
        if (vec4)
          v = c == 0 ? f4.x : (c == 1 ? f4.y : (c == 2 ? f4.z : f4.w));
        else if (q.dtype == MPG_F32)
          v = __ldg(reinterpret_cast<const float*>(q.ptr) + spix * q.cstride + q.c0 + c);
        else
          v = h16_to_float(reinterpret_cast<const uint16_t*>(q.ptr)[spix * q.cstride + q.c0 + c], q.dtype);
        const int k = oc & 7;
        const unsigned long long hv = static_cast<unsigned long long>(float_to_h16(v, p.out_dtype)) << ((k & 3) * 16);
        if (k < 4) lo |= hv; else hi |= hv;
        if (k == 7) {
          o[oc >> 3] = make_uint4(static_cast<uint32_t>(lo), static_cast<uint32_t>(lo >> 32), static_cast<uint32_t>(hi),
                                  static_cast<uint32_t>(hi >> 32));
          lo = hi = 0ull;
        }
      }
    }
    if (oc & 7) {
      o[oc >> 3] = make_uint4(static_cast<uint32_t>(lo), static_cast<uint32_t>(lo >> 32), static_cast<uint32_t>(hi),
                              static_cast<uint32_t>(hi >> 32));
      oc = (oc | 7) + 1;
    }
    for (; oc < p.out_cstride; oc += 8) o[oc >> 3] = make_uint4(0u, 0u, 0u, 0u);
    continue;
  }
  for (int s = 0; s < p.nsrc; ++s) {
    const PackSrc& q = p.src[s];
    const long long spix = (static_cast<long long>(n) * q.sh + pack_div(y, q.fh, q.lfh)) * q.sw + pack_div(x, q.fw, q.lfw);
    for (int c = 0; c < q.nch; ++c, ++oc) {
      float v;
      if (q.dtype == MPG_F32)
        v = __ldg(reinterpret_cast<const float*>(q.ptr) + spix * q.cstride + q.c0 + c);
      else
        v = h16_to_float(reinterpret_cast<const uint16_t*>(q.ptr)[spix * q.cstride + q.c0 + c], q.dtype);
      if (p.out_dtype == MPG_F32)
        reinterpret_cast<float*>(p.out)[pix * p.out_cstride + oc] = v;
      else
        reinterpret_cast<uint16_t*>(p.out)[pix * p.out_cstride + oc] = float_to_h16(v, p.out_dtype);
    }
  }
  for (; oc < p.out_cstride; ++oc) {
    if (p.out_dtype == MPG_F32)
      reinterpret_cast<float*>(p.out)[pix * p.out_cstride + oc] = 0.0f;
    else
      reinterpret_cast<uint16_t*>(p.out)[pix * p.out_cstride + oc] = 0;
  }
  }
}

// Dense fp32 -> 16-bit cast with channel padding (one source, no resize, all of its channels): the common case of the training
// step (every tensor-core conv gets a 16-bit copy of its fp32 input / output gradient). One thread per 8-channel chunk of the
// OUTPUT (one 16-byte store; consecutive threads = consecutive chunks = coalesced on both sides). The generic kernel below
// gives a thread a whole pixel: with 128 channels that is 512 contiguous bytes per lane, 512 bytes apart -- measured 83 us
// per call on [16,64,64,128] (36 % of the training loop body) against ~8 us for the bytes moved.
__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ in, uint4* __restrict__ out, long long npix,
                                                        int cin, int cs8, int dtype) {
  const long long total = npix * cs8;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const long long pix = e / cs8;
    const int c0 = static_cast<int>(e - pix * cs8) * 8;
    const float* p = in + pix * cin + c0;
    float v[8];
    if (c0 + 8 <= cin && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < cin) ? __ldg(p + j) : 0.0f;
    }
    uint4 q;
    q.x = pack_h16x2(v[0], v[1], dtype);
    q.y = pack_h16x2(v[2], v[3], dtype);
    q.z = pack_h16x2(v[4], v[5], dtype);
    q.w = pack_h16x2(v[6], v[7], dtype);
    out[e] = q;
  }
}

struct DensParams {
  const float* dens;  // [n, oh, ow]
  const void* src;    // [n, sh, sw, cstride]
  int src_dtype, src_cstride, src_c;
  int mode;  // 0: same-resolution channel add, 2: TF1 legacy bicubic
  int n, oh, ow, sh, sw;
  const int* iy;    // [oh][4]
  const float* wy;  // [oh][4]
  const int* ix;    // [ow][4]
  const float* wx;  // [ow][4]
  float* out;       // [n, oh, ow]
};

__device__ __forceinline__ float src_at(const DensParams& p, int n, int y, int x) {
  const long long o = ((static_cast<long long>(n) * p.sh + y) * p.sw + x) * p.src_cstride + p.src_c;
  if (p.src_dtype == MPG_F32) return __ldg(reinterpret_cast<const float*>(p.src) + o);
  return h16_to_float(reinterpret_cast<const uint16_t*>(p.src)[o], p.src_dtype);
}

__global__ void __launch_bounds__(256) dens_residual_kernel(const DensParams p) {
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long npix = static_cast<long long>(p.n) * p.oh * p.ow;
  if (pix >= npix) return;
  const int x = static_cast<int>(pix % p.ow);
  const long long r = pix / p.ow;
  const int y = static_cast<int>(r % p.oh);
  const int n = static_cast<int>(r / p.oh);
  float add;
  if (p.mode == 0) {
    add = src_at(p, n, y, x);
  } else {
    // TF1 ResizeBicubic: interpolate along x for each of the 4 rows, then along y (fp32)
    add = 0.0f;
#pragma unroll
    for (int ty = 0; ty < 4; ++ty) {
      const int sy = p.iy[y * 4 + ty];
      float row = 0.0f;
#pragma unroll
      for (int tx = 0; tx < 4; ++tx) row += src_at(p, n, sy, p.ix[x * 4 + tx]) * p.wx[x * 4 + tx];
      add += row * p.wy[y * 4 + ty];
    }
  }
  p.out[pix] = p.dens[pix] + add;
}

// Density output of the out.py generators in ONE pass: the 1x1 conv to a single channel (g_cdensOut, gain 1, no activation,
// GAN/multipassGAN-out.py:282) plus the additive residual (:327-332) -- the conv as its own launch re-read the whole
// feature tensor through the tensor-core path (N padded to 16) and the residual made another pass over the fp32 image.
struct DensOutParams {
  DensParams d;      // d.dens unused; d.mode: -1 no residual, 0 channel add, 2 TF1 bicubic
  const void* x;     // [n, oh, ow, x_cstride] features
  int x_dtype, x_cstride, cin;
  float bias;
  float w[64];
};

// One thread per output pixel, grid = (ceil(ow / 256), n * oh): 32-bit index arithmetic, the weights are uniform constant-bank
// operands, and the bicubic tables are read as ONE 16-byte vector each -- the kernel is bound by its load instructions
// (L1 hits), not by HBM: a lanes-per-pixel variant with shuffles issued 2-4x more of them and measured 1.5-2x slower.
__global__ void __launch_bounds__(256) dens_out_kernel(const __grid_constant__ DensOutParams q) {
  const DensParams& p = q.d;
  const int x = static_cast<int>(blockIdx.x) * 256 + static_cast<int>(threadIdx.x);
  if (x >= p.ow) return;
  const int rowi = static_cast<int>(blockIdx.y);
  const int n = rowi / p.oh;
  const int y = rowi - n * p.oh;
  const long long pix = static_cast<long long>(rowi) * p.ow + x;
  float acc = q.bias;
  if (q.x_dtype == MPG_F32) {
    const float* xp = reinterpret_cast<const float*>(q.x) + pix * q.x_cstride;
    for (int c = 0; c < q.cin; ++c) acc = fmaf(__ldg(xp + c), q.w[c], acc);
  } else {
    const uint4* xp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(q.x) + pix * q.x_cstride);
    const int nv = (q.cin + 7) >> 3;
#pragma unroll 4
    for (int v = 0; v < nv; ++v) {
      const uint4 u = __ldg(xp + v);
      const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = v * 8 + j * 2;  // (channels beyond cin are zero in the tensor and have zero weights here)
        acc = fmaf(h16_to_float(static_cast<uint16_t>(wd[j] & 0xFFFFu), q.x_dtype), q.w[c], acc);
        acc = fmaf(h16_to_float(static_cast<uint16_t>(wd[j] >> 16), q.x_dtype), q.w[c + 1], acc);
      }
    }
  }
  if (p.mode == 0) {
    acc += src_at(p, n, y, x);
  } else if (p.mode == 2) {
    // TF1 ResizeBicubic: interpolate along x for each of the 4 rows, then along y (fp32); tables are [size][4]
    const int4 ix = __ldg(reinterpret_cast<const int4*>(p.ix) + x);
    const float4 wx = __ldg(reinterpret_cast<const float4*>(p.wx) + x);
    const int4 iy = __ldg(reinterpret_cast<const int4*>(p.iy) + y);
    const float4 wy = __ldg(reinterpret_cast<const float4*>(p.wy) + y);
    const int sy[4] = {iy.x, iy.y, iy.z, iy.w};
    const float wyv[4] = {wy.x, wy.y, wy.z, wy.w};
    float add = 0.0f;
#pragma unroll
    for (int ty = 0; ty < 4; ++ty) {
      float row = src_at(p, n, sy[ty], ix.x) * wx.x;
      row += src_at(p, n, sy[ty], ix.y) * wx.y;
      row += src_at(p, n, sy[ty], ix.z) * wx.z;
      row += src_at(p, n, sy[ty], ix.w) * wx.w;
      add += row * wyv[ty];
    }
    acc += add;
  }
  p.out[pix] = acc;
}

// Standalone tf.image.resize_images for any channel count (tools_wscale/GAN.py:541 avg_depool modes 0 / 2 when the
// result feeds something other than the additive density residual). One thread per (pixel, channel).
struct ResizeParams {
  const void* src;  // [n, sh, sw, src_cstride]
  void* out;        // [n, oh, ow, out_cstride]
  int src_dtype, src_cstride, out_dtype, out_cstride, c;
  int mode;  // 0: TF1 legacy bilinear (align_corners=False, no half-pixel centres), 2: TF1 legacy bicubic
  int n, oh, ow, sh, sw;
  float scale_y, scale_x;  // in / out, computed in fp32 like CalculateResizeScale
  const int* iy;
  const float* wy;
  const int* ix;
  const float* wx;
};

__device__ __forceinline__ float rs_at(const ResizeParams& p, int n, int y, int x, int ch) {
  const long long o = ((static_cast<long long>(n) * p.sh + y) * p.sw + x) * p.src_cstride + ch;
  if (p.src_dtype == MPG_F32) return __ldg(reinterpret_cast<const float*>(p.src) + o);
  return h16_to_float(reinterpret_cast<const uint16_t*>(p.src)[o], p.src_dtype);
}

__global__ void __launch_bounds__(256) resize_images_kernel(const ResizeParams p) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(p.n) * p.oh * p.ow * p.out_cstride;
  if (e >= total) return;
  const int ch = static_cast<int>(e % p.out_cstride);
  long long r = e / p.out_cstride;
  const int x = static_cast<int>(r % p.ow);
  r /= p.ow;
  const int y = static_cast<int>(r % p.oh);
  const int n = static_cast<int>(r / p.oh);
  float v = 0.0f;
  if (ch < p.c) {
    if (p.mode == 0) {
      // ResizeBilinear: in = out * scale; lower = floor(in), upper = min(lower + 1, size - 1), lerp = in - lower;
      // top = tl + (tr - tl) * x_lerp; bottom = bl + (br - bl) * x_lerp; out = top + (bottom - top) * y_lerp
      const float fy = static_cast<float>(y) * p.scale_y, fx = static_cast<float>(x) * p.scale_x;
      const int y0 = static_cast<int>(floorf(fy)), x0 = static_cast<int>(floorf(fx));
      const int y1 = min(y0 + 1, p.sh - 1), x1 = min(x0 + 1, p.sw - 1);
      const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
      const float tl = rs_at(p, n, y0, x0, ch), tr = rs_at(p, n, y0, x1, ch);
      const float bl = rs_at(p, n, y1, x0, ch), br = rs_at(p, n, y1, x1, ch);
      const float top = tl + (tr - tl) * lx;
      const float bot = bl + (br - bl) * lx;
      v = top + (bot - top) * ly;
    } else {
#pragma unroll
      for (int ty = 0; ty < 4; ++ty) {
        const int sy = p.iy[y * 4 + ty];
        float row = 0.0f;
#pragma unroll
        for (int tx = 0; tx < 4; ++tx) row += rs_at(p, n, sy, p.ix[x * 4 + tx], ch) * p.wx[x * 4 + tx];
        v += row * p.wy[y * 4 + ty];
      }
    }
  }
  if (p.out_dtype == MPG_F32) reinterpret_cast<float*>(p.out)[e] = v;
  else reinterpret_cast<uint16_t*>(p.out)[e] = float_to_h16(v, p.out_dtype);
}

// Range check of a stored activation tensor (validation mode): every 16-bit store of the conv kernels is a saturating
// conversion (cvt.rn.satfinite, common.h), so a value beyond the type's range lands exactly on +-max-finite. Counts the
// elements that are +-max-finite, +-inf or NaN (fp32: non-finite only).
__global__ void __launch_bounds__(256) count_saturated_kernel(const void* __restrict__ t, long long count, int dtype,
                                                              unsigned long long* __restrict__ counter) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  unsigned int local = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    if (dtype == MPG_F32) {
      const uint32_t u = __float_as_uint(static_cast<const float*>(t)[i]) & 0x7fffffffu;
      local += u >= 0x7f800000u;
    } else {
      const uint32_t u = static_cast<const uint16_t*>(t)[i] & 0x7fffu;
      local += u >= (dtype == MPG_F16 ? 0x7bffu : 0x7f7fu);  // max finite (65504 / 3.39e38) or above (inf, NaN)
    }
  }
  local = __reduce_add_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(counter, static_cast<unsigned long long>(local));
}

// TF1 bicubic coefficient table (Keys, A = -0.75, 1024 entries), fp32 like resize_bicubic_op.cc
void bicubic_axis(int in_size, int out_size, std::vector<int>& idx, std::vector<float>& wts) {
  static float tab[1025][2];
  static bool init = false;
  if (!init) {
    const float a = -0.75f;
    for (int i = 0; i <= 1024; ++i) {
      float x = static_cast<float>(i) / 1024.0f;
      tab[i][0] = ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
      x += 1.0f;
      tab[i][1] = ((a * x - 5.0f * a) * x + 8.0f * a) * x - 4.0f * a;
    }
    init = true;
  }
  idx.resize(static_cast<size_t>(out_size) * 4);
  wts.resize(static_cast<size_t>(out_size) * 4);
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  for (int o = 0; o < out_size; ++o) {
    const float in_f = static_cast<float>(o) * scale;
    const int i = static_cast<int>(floorf(in_f));
    const float delta = in_f - static_cast<float>(i);
    const int off = static_cast<int>(lrintf(delta * 1024.0f));
    wts[o * 4 + 0] = tab[off][1];
    wts[o * 4 + 1] = tab[off][0];
    wts[o * 4 + 2] = tab[1024 - off][0];
    wts[o * 4 + 3] = tab[1024 - off][1];
    for (int t = 0; t < 4; ++t) {
      int s = i - 1 + t;
      s = s < 0 ? 0 : (s > in_size - 1 ? in_size - 1 : s);
      idx[o * 4 + t] = s;
    }
  }
}

}  // namespace
}  // namespace mpg

extern "C" {

int mpg_pack_channels(mpg_handle h, const mpg_chan_src* srcs, int nsrc, void* out, int out_dtype, int out_cstride,
                      int n, int oh, int ow, void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(h && srcs && out, "mpg_pack_channels: null argument");
  MPG_CHECK_ARG(nsrc >= 1 && nsrc <= kMaxSrc, "mpg_pack_channels: nsrc %d not in [1,%d]", nsrc, kMaxSrc);
  PackParams p;
  int total = 0;
  for (int s = 0; s < nsrc; ++s) {
    MPG_CHECK_ARG(srcs[s].ptr && srcs[s].factor_h >= 1 && srcs[s].factor_w >= 1 && is_dtype(srcs[s].dtype),
                  "pack: bad source %d", s);
    MPG_CHECK_ARG(oh % srcs[s].factor_h == 0 && ow % srcs[s].factor_w == 0, "pack: size not divisible by factor");
    MPG_CHECK_ARG(srcs[s].c0 >= 0 && srcs[s].nch > 0 && srcs[s].c0 + srcs[s].nch <= srcs[s].cstride,
                  "pack: channel range of source %d", s);
    auto lg2 = [](int f) {
      int l = 0;
      while ((1 << l) < f) ++l;
      return (1 << l) == f ? l : -1;
    };
    p.src[s] = {srcs[s].ptr, srcs[s].dtype, srcs[s].cstride, srcs[s].c0, srcs[s].nch, srcs[s].factor_h,
                srcs[s].factor_w, oh / srcs[s].factor_h, ow / srcs[s].factor_w, lg2(srcs[s].factor_h), lg2(srcs[s].factor_w)};
    total += srcs[s].nch;
  }
  MPG_CHECK_ARG(total <= out_cstride, "pack: %d channels do not fit out_cstride %d", total, out_cstride);
  p.nsrc = nsrc;
  p.n = n;
  p.oh = oh;
  p.ow = ow;
  p.out_dtype = out_dtype;
  p.out_cstride = out_cstride;
  p.out = out;
  if (nsrc == 1 && srcs[0].factor_h == 1 && srcs[0].factor_w == 1 && srcs[0].dtype == MPG_F32 && is_h16(out_dtype) &&
      srcs[0].c0 == 0 && srcs[0].nch == srcs[0].cstride && out_cstride % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(srcs[0].ptr) & 3) == 0) {
    const long long npix = static_cast<long long>(n) * oh * ow;
    const long long total = npix * (out_cstride / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > h->sm_count * 16LL) blocks = h->sm_count * 16LL;
    cast_pad_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(srcs[0].ptr), static_cast<uint4*>(out), npix, srcs[0].nch, out_cstride / 8, out_dtype);
    MPG_CUDA(cudaGetLastError());
    return MPG_OK;
  }
  MPG_CHECK_ARG((static_cast<long long>(n) * oh + kPackRows - 1) / kPackRows <= 65535, "pack: n*oh = %lld rows exceed the grid limit", static_cast<long long>(n) * oh);
  pack_channels_kernel<<<dim3(static_cast<unsigned>((ow + 255) / 256), static_cast<unsigned>((n * oh + kPackRows - 1) / kPackRows)), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

int mpg_bicubic_plan_create(mpg_handle h, int in_h, int in_w, int out_h, int out_w, void** plan_out) {
  using namespace mpg;
  MPG_CHECK_ARG(h && plan_out && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "bicubic plan: bad argument");
  std::vector<int> iy, ix;
  std::vector<float> wy, wx;
  bicubic_axis(in_h, out_h, iy, wy);
  bicubic_axis(in_w, out_w, ix, wx);
  // layout: [iy | ix] ints then [wy | wx] floats in one allocation
  const size_t ni = iy.size() + ix.size();
  const size_t bytes = ni * sizeof(int) + ni * sizeof(float);
  char* d = nullptr;
  MPG_CUDA(cudaMalloc(&d, bytes));
  MPG_CUDA(cudaMemcpy(d, iy.data(), iy.size() * 4, cudaMemcpyHostToDevice));
  MPG_CUDA(cudaMemcpy(d + iy.size() * 4, ix.data(), ix.size() * 4, cudaMemcpyHostToDevice));
  MPG_CUDA(cudaMemcpy(d + ni * 4, wy.data(), wy.size() * 4, cudaMemcpyHostToDevice));
  MPG_CUDA(cudaMemcpy(d + ni * 4 + wy.size() * 4, wx.data(), wx.size() * 4, cudaMemcpyHostToDevice));
  *plan_out = d;
  return MPG_OK;
}

int mpg_bicubic_plan_destroy(void* plan) {
  if (plan) cudaFree(plan);
  return MPG_OK;
}

int mpg_resize_images(mpg_handle h, const void* src, int src_dtype, int src_cstride, int c, int n, int src_h, int src_w,
                      void* out, int out_dtype, int out_cstride, int out_h, int out_w, int mode, void* bicubic_plan,
                      void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(h && src && out, "mpg_resize_images: null argument");
  MPG_CHECK_ARG(mode == 0 || mode == 2, "mpg_resize_images: mode must be 0 (bilinear) or 2 (TF1 bicubic); nearest is mpg_pack_channels");
  MPG_CHECK_ARG(is_dtype(src_dtype) && is_dtype(out_dtype), "mpg_resize_images: bad dtype");
  MPG_CHECK_ARG(c > 0 && c <= src_cstride && c <= out_cstride && n > 0 && src_h > 0 && src_w > 0 && out_h > 0 && out_w > 0,
                "mpg_resize_images: bad shape");
  MPG_CHECK_ARG(mode == 0 || bicubic_plan != nullptr, "mpg_resize_images: bicubic plan missing");
  ResizeParams p;
  p.src = src;
  p.out = out;
  p.src_dtype = src_dtype;
  p.src_cstride = src_cstride;
  p.out_dtype = out_dtype;
  p.out_cstride = out_cstride;
  p.c = c;
  p.mode = mode;
  p.n = n;
  p.oh = out_h;
  p.ow = out_w;
  p.sh = src_h;
  p.sw = src_w;
  p.scale_y = static_cast<float>(src_h) / static_cast<float>(out_h);
  p.scale_x = static_cast<float>(src_w) / static_cast<float>(out_w);
  p.iy = p.ix = nullptr;
  p.wy = p.wx = nullptr;
  if (mode == 2) {
    const char* d = static_cast<const char*>(bicubic_plan);
    const size_t ni = static_cast<size_t>(out_h + out_w) * 4;
    p.iy = reinterpret_cast<const int*>(d);
    p.ix = p.iy + static_cast<size_t>(out_h) * 4;
    p.wy = reinterpret_cast<const float*>(d + ni * 4);
    p.wx = p.wy + static_cast<size_t>(out_h) * 4;
  }
  const long long total = static_cast<long long>(n) * out_h * out_w * out_cstride;
  DeviceGuard guard(h->device);
  resize_images_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

int mpg_count_saturated(mpg_handle h, const void* t, long long count, int dtype, unsigned long long* counter_dev,
                        void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(h && t && counter_dev && count >= 0 && is_dtype(dtype), "mpg_count_saturated: bad argument");
  if (count == 0) return MPG_OK;
  long long blocks = (count + 255) / 256;
  const long long cap = static_cast<long long>(h->sm_count) * 16;
  if (blocks > cap) blocks = cap;
  DeviceGuard guard(h->device);
  count_saturated_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(t, count, dtype, counter_dev);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

int mpg_dens_residual(mpg_handle h, const float* dens, const void* src, int src_dtype, int src_cstride, int src_c,
                      int mode, void* bicubic_plan, int n, int out_h, int out_w, int src_h, int src_w, float* out,
                      void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(h && dens && src && out, "mpg_dens_residual: null argument");
  MPG_CHECK_ARG(mode == 0 || mode == 2, "mpg_dens_residual: mode must be 0 (channel add) or 2 (TF1 bicubic)");
  MPG_CHECK_ARG(mode == 2 || (src_h == out_h && src_w == out_w), "mpg_dens_residual: mode 0 needs equal sizes");
  MPG_CHECK_ARG(mode == 0 || bicubic_plan != nullptr, "mpg_dens_residual: bicubic plan missing");
  DensParams p;
  p.dens = dens;
  p.src = src;
  p.src_dtype = src_dtype;
  p.src_cstride = src_cstride;
  p.src_c = src_c;
  p.mode = mode;
  p.n = n;
  p.oh = out_h;
  p.ow = out_w;
  p.sh = src_h;
  p.sw = src_w;
  if (mode == 2) {
    const char* d = static_cast<const char*>(bicubic_plan);
    const size_t ni = static_cast<size_t>(out_h + out_w) * 4;
    p.iy = reinterpret_cast<const int*>(d);
    p.ix = p.iy + static_cast<size_t>(out_h) * 4;
    p.wy = reinterpret_cast<const float*>(d + ni * 4);
    p.wx = p.wy + static_cast<size_t>(out_h) * 4;
  } else {
    p.iy = p.ix = nullptr;
    p.wy = p.wx = nullptr;
  }
  p.out = out;
  const long long npix = static_cast<long long>(n) * out_h * out_w;
  dens_residual_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

int mpg_dens_out(mpg_handle h, const void* x, int x_dtype, int x_cstride, int cin, const float* w_host, float bias,
                 const void* src, int src_dtype, int src_cstride, int src_c, int mode, void* bicubic_plan, int n, int out_h,
                 int out_w, int src_h, int src_w, float* out, void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(h && x && w_host && out, "mpg_dens_out: null argument");
  MPG_CHECK_ARG(cin >= 1 && cin <= 64 && x_cstride >= cin, "mpg_dens_out: cin=%d (1..64) x_cstride=%d", cin, x_cstride);
  MPG_CHECK_ARG(is_dtype(x_dtype), "mpg_dens_out: bad x_dtype %d", x_dtype);
  MPG_CHECK_ARG(x_dtype == MPG_F32 || (x_cstride % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0),
                "mpg_dens_out: 16-bit features need a channel stride that is a multiple of 8 and a 16-byte aligned tensor");
  MPG_CHECK_ARG(mode == -1 || mode == 0 || mode == 2, "mpg_dens_out: mode must be -1 (no residual), 0 (channel add) or 2 (TF1 bicubic)");
  MPG_CHECK_ARG(mode == -1 || src != nullptr, "mpg_dens_out: residual source missing");
  MPG_CHECK_ARG(mode != 0 || (src_h == out_h && src_w == out_w), "mpg_dens_out: mode 0 needs equal sizes");
  MPG_CHECK_ARG(mode != 2 || bicubic_plan != nullptr, "mpg_dens_out: bicubic plan missing");
  DensOutParams q;
  memset(&q, 0, sizeof(q));
  DensParams& p = q.d;
  p.src = src;
  p.src_dtype = src_dtype;
  p.src_cstride = src_cstride;
  p.src_c = src_c;
  p.mode = mode;
  p.n = n;
  p.oh = out_h;
  p.ow = out_w;
  p.sh = src_h;
  p.sw = src_w;
  if (mode == 2) {
    const char* d = static_cast<const char*>(bicubic_plan);
    const size_t ni = static_cast<size_t>(out_h + out_w) * 4;
    p.iy = reinterpret_cast<const int*>(d);
    p.ix = p.iy + static_cast<size_t>(out_h) * 4;
    p.wy = reinterpret_cast<const float*>(d + ni * 4);
    p.wx = p.wy + static_cast<size_t>(out_h) * 4;
  }
  p.out = out;
  q.x = x;
  q.x_dtype = x_dtype;
  q.x_cstride = x_cstride;
  q.cin = cin;
  q.bias = bias;
  for (int c = 0; c < cin; ++c) q.w[c] = w_host[c];
  DeviceGuard guard(h->device);
  const long long npix = static_cast<long long>(n) * out_h * out_w;
  (void)npix;
  MPG_CHECK_ARG(static_cast<long long>(n) * out_h <= 65535, "mpg_dens_out: n*out_h = %lld rows exceed the grid limit", static_cast<long long>(n) * out_h);
  const dim3 blocks(static_cast<unsigned>((out_w + 255) / 256), static_cast<unsigned>(n * out_h));
  dens_out_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(q);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

}  // extern "C"
