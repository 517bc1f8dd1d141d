// Volume pipeline kernels (HBM-bandwidth bound): slice assembler, 3-D axis permutation, threshold.
// They replace the host-side numpy / scipy code of generate3DUniForNewNetwork
// (GAN/multipassGAN-out.py:390-618, GAN/multipassGAN-4x.py:1090-1169); see include/mpg.h.
#include "common.h"

namespace mpg {
namespace {

struct AsmParams {
  int dims[3];
  int vol_c;
  int axis_of[3];
  int zoom[3];
  int nchan;
  int chan_src[8];
  float chan_scale[8];
  int add_adj;
  int out_dtype, out_cstride;
  int odims[3];  // n_slices, H, W of the interpolated stack
  long long vstride[3];
  const float* vol;
  const float* dens;
  int slice0, count, dens_slice0;
  void* out;
};

struct Lerp {
  int lo, hi;
  float t;
};

// align-corners coordinate of scipy.ndimage.zoom(order=1): o * (n_in-1)/(n_out-1)
__device__ __forceinline__ Lerp lerp_coord(int o, int n_in, int zoom) {
  Lerp l;
  if (zoom == 1) {
    l.lo = l.hi = o;
    l.t = 0.0f;
    return l;
  }
  const int n_out = n_in * zoom;
  const double c = (n_out > 1) ? static_cast<double>(o) * static_cast<double>(n_in - 1) / static_cast<double>(n_out - 1) : 0.0;
  int lo = static_cast<int>(floor(c));
  lo = lo < 0 ? 0 : (lo > n_in - 1 ? n_in - 1 : lo);
  l.lo = lo;
  l.hi = lo + 1 > n_in - 1 ? n_in - 1 : lo + 1;
  l.t = static_cast<float>(c - static_cast<double>(lo));
  return l;
}

__device__ __forceinline__ float sample(const AsmParams& p, const Lerp (&l)[3], int ch) {
  // l[k] addresses source axis axis_of[k]; up to 8 corners, skipping zero-weight ones
  float acc = 0.0f;
#pragma unroll
  for (int c0 = 0; c0 < 2; ++c0) {
    const float w0 = c0 ? l[0].t : 1.0f - l[0].t;
    if (w0 == 0.0f) continue;
    const long long o0 = static_cast<long long>(c0 ? l[0].hi : l[0].lo) * p.vstride[0];
#pragma unroll
    for (int c1 = 0; c1 < 2; ++c1) {
      const float w1 = c1 ? l[1].t : 1.0f - l[1].t;
      if (w1 == 0.0f) continue;
      const long long o1 = o0 + static_cast<long long>(c1 ? l[1].hi : l[1].lo) * p.vstride[1];
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const float w2 = c2 ? l[2].t : 1.0f - l[2].t;
        if (w2 == 0.0f) continue;
        const long long o2 = o1 + static_cast<long long>(c2 ? l[2].hi : l[2].lo) * p.vstride[2];
        acc = fmaf(w0 * w1 * w2, __ldg(p.vol + o2 * p.vol_c + ch), acc);
      }
    }
  }
  return acc;
}

__global__ void __launch_bounds__(256) slice_assemble_kernel(const AsmParams p) {
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int H = p.odims[1], W = p.odims[2];
  const long long npix = static_cast<long long>(p.count) * H * W;
  if (pix >= npix) return;
  const int j = static_cast<int>(pix % W);
  const long long r = pix / W;
  const int i = static_cast<int>(r % H);
  const int s = p.slice0 + static_cast<int>(r / H);

  Lerp l[3];
  l[0] = lerp_coord(s, p.dims[p.axis_of[0]], p.zoom[0]);
  l[1] = lerp_coord(i, p.dims[p.axis_of[1]], p.zoom[1]);
  l[2] = lerp_coord(j, p.dims[p.axis_of[2]], p.zoom[2]);

  float v[16];
  int oc = 0;
  if (p.dens) v[oc++] = __ldg(p.dens + (static_cast<long long>(s - p.dens_slice0) * H + i) * W + j);
  for (int c = 0; c < p.nchan; ++c) v[oc++] = p.chan_scale[c] * sample(p, l, p.chan_src[c]);
  if (p.add_adj) {
    for (int d = -1; d <= 1; d += 2) {
      const int sn = s + d;
      float a = 0.0f;
      if (sn >= 0 && sn < p.odims[0]) {
        Lerp ln[3] = {lerp_coord(sn, p.dims[p.axis_of[0]], p.zoom[0]), l[1], l[2]};
        a = p.chan_scale[0] * sample(p, ln, p.chan_src[0]);
      }
      v[oc++] = a;
    }
  }
  if (p.out_dtype == MPG_F32) {
    float* o = reinterpret_cast<float*>(p.out) + pix * p.out_cstride;
    for (int c = 0; c < p.out_cstride; ++c) o[c] = c < oc ? v[c] : 0.0f;
  } else {
    uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + pix * p.out_cstride;
    for (int c = 0; c < p.out_cstride; ++c) o[c] = float_to_h16(c < oc ? v[c] : 0.0f, p.out_dtype);
  }
}

// Fast path of the assembler for the 4-channel volumes of the 4x recipe (vol_c == 4, fp32 rows of 4 channels, no adjacent
// slices): one thread owns one (i, j) of the slice plane and produces ALL slices of the batch. The in-plane bilinear
// combination of the four (l1, l2) corners is one 128-bit load per corner and is evaluated once per distinct source
// position along the slice axis (a batch of 8 output slices at zoom 4 touches 3-4 of them), each slice is then a lerp of two
// such values: ~16 128-bit loads per thread for 8 output pixels, where the generic kernel issues 24 scalar loads and
// three double-precision coordinate computations per output pixel. The value of a slice depends on (s, i, j) only, never on
// the batch it is computed in (sharded and tiled runs stay bit-identical to whole runs).
__device__ __forceinline__ float comp4(const float4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }

__global__ void __launch_bounds__(128) slice_assemble_c4_kernel(const AsmParams p) {
  const int H = p.odims[1], W = p.odims[2];
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const int j = pix % W, i = pix / W;
  const Lerp l1 = lerp_coord(i, p.dims[p.axis_of[1]], p.zoom[1]);
  const Lerp l2 = lerp_coord(j, p.dims[p.axis_of[2]], p.zoom[2]);
  const long long o00 = l1.lo * p.vstride[1] + l2.lo * p.vstride[2], o01 = l1.lo * p.vstride[1] + l2.hi * p.vstride[2];
  const long long o10 = l1.hi * p.vstride[1] + l2.lo * p.vstride[2], o11 = l1.hi * p.vstride[1] + l2.hi * p.vstride[2];
  const float w00 = (1.0f - l1.t) * (1.0f - l2.t), w01 = (1.0f - l1.t) * l2.t, w10 = l1.t * (1.0f - l2.t), w11 = l1.t * l2.t;
  const float4* vol4 = reinterpret_cast<const float4*>(p.vol);
  auto plane = [&](int x) {
    const long long b = static_cast<long long>(x) * p.vstride[0];
    const float4 a = __ldg(vol4 + b + o00), c = __ldg(vol4 + b + o01), d = __ldg(vol4 + b + o10), e = __ldg(vol4 + b + o11);
    float4 r;
    r.x = fmaf(w11, e.x, fmaf(w10, d.x, fmaf(w01, c.x, w00 * a.x)));
    r.y = fmaf(w11, e.y, fmaf(w10, d.y, fmaf(w01, c.y, w00 * a.y)));
    r.z = fmaf(w11, e.z, fmaf(w10, d.z, fmaf(w01, c.z, w00 * a.z)));
    r.w = fmaf(w11, e.w, fmaf(w10, d.w, fmaf(w01, c.w, w00 * a.w)));
    return r;
  };
  int xa = -1, xb = -1;
  float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
  float4* out = reinterpret_cast<float4*>(p.out);
  for (int k = 0; k < p.count; ++k) {
    const int s = p.slice0 + k;
    const Lerp l0 = lerp_coord(s, p.dims[p.axis_of[0]], p.zoom[0]);
    const float4 na = l0.lo == xa ? va : (l0.lo == xb ? vb : plane(l0.lo));
    const float4 nb = l0.hi == l0.lo ? na : (l0.hi == xb ? vb : (l0.hi == xa ? va : plane(l0.hi)));
    xa = l0.lo, xb = l0.hi, va = na, vb = nb;
    float4 v;
    const float t = l0.t, u = 1.0f - l0.t;
    v.x = fmaf(t, vb.x, u * va.x);
    v.y = fmaf(t, vb.y, u * va.y);
    v.z = fmaf(t, vb.z, u * va.z);
    v.w = fmaf(t, vb.w, u * va.w);
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    int oc = 0;
    if (p.dens) o[oc++] = __ldg(p.dens + (static_cast<long long>(s - p.dens_slice0) * H + i) * W + j);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < p.nchan) o[oc + c] = p.chan_scale[c] * comp4(v, p.chan_src[c]);
    out[(static_cast<long long>(k) * H + i) * W + j] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// 32x32 shared-memory tile transpose between input axis 2 (x) and input axis `a` (y); axis `b` is batch.
__global__ void __launch_bounds__(256)
transpose_tile_kernel(const float* __restrict__ in, float* __restrict__ out, int nx, int ny, long long in_sy,
                      long long in_sb, long long out_sx, long long out_sb, float threshold) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const long long bi = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int x = bx + tx, y = by + ty + k;
    if (x < nx && y < ny) tile[ty + k][tx] = in[bi * in_sb + static_cast<long long>(y) * in_sy + x];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int y = by + tx, x = bx + ty + k;
    if (x < nx && y < ny) {
      float v = tile[tx][ty + k];
      if (v < threshold) v = 0.0f;
      out[bi * out_sb + static_cast<long long>(x) * out_sx + y] = v;
    }
  }
}

// Same transpose with 128-bit global accesses: 64x64 tile, every thread moves float4s on both sides (a warp reads two
// 256-byte row segments and writes two 256-byte row segments). Needs nx, ny multiples of 4 and 16-byte aligned rows;
// the 32x32 scalar kernel above is the fallback. The scalar kernel reached 66 % of the measured HBM copy rate.
__global__ void __launch_bounds__(256)
transpose_tile64_kernel(const float* __restrict__ in, float* __restrict__ out, int nx, int ny, long long in_sy,
                        long long in_sb, long long out_sx, long long out_sb, float threshold) {
  __shared__ float tile[64][65];
  const int bx = blockIdx.x * 64, by = blockIdx.y * 64;
  const long long bi = blockIdx.z;
  const int c4 = threadIdx.x & 15, r = threadIdx.x >> 4;  // 16 float4 columns x 16 rows per pass
#pragma unroll
  for (int k = 0; k < 64; k += 16) {
    const int x = bx + 4 * c4, y = by + r + k;
    if (x < nx && y < ny) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(in + bi * in_sb + static_cast<long long>(y) * in_sy + x));
      tile[r + k][4 * c4 + 0] = v.x;
      tile[r + k][4 * c4 + 1] = v.y;
      tile[r + k][4 * c4 + 2] = v.z;
      tile[r + k][4 * c4 + 3] = v.w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 64; k += 16) {
    const int x = bx + r + k, y = by + 4 * c4;  // output row = input column x, 4 consecutive input rows y..y+3
    if (x < nx && y < ny) {
      float4 v = make_float4(tile[4 * c4 + 0][r + k], tile[4 * c4 + 1][r + k], tile[4 * c4 + 2][r + k], tile[4 * c4 + 3][r + k]);
      v.x = v.x < threshold ? 0.0f : v.x;
      v.y = v.y < threshold ? 0.0f : v.y;
      v.z = v.z < threshold ? 0.0f : v.z;
      v.w = v.w < threshold ? 0.0f : v.w;
      __stcs(reinterpret_cast<float4*>(out + bi * out_sb + static_cast<long long>(x) * out_sx + y), v);
    }
  }
}

// Axis change between passes on several GPUs as ONE kernel: the rank's slab [d0 = S/G, d1, d2] is transposed
// tile by tile exactly like transpose_tile64_kernel, but every float4 is stored straight into the OUTPUT slab of the rank
// that owns it in the next pass -- peer memory mapped over NVLink / NVSwitch (CUDA VMM handles from
// torch.distributed._symmetric_memory). This replaces pack-transpose -> ncclAllToAll -> unpack-permute (three passes over
// the slab and a staging copy) by one read of the slab and one remote write.
//   axis roles: x = input axis 2 (contiguous), y = input axis `ay` (becomes contiguous in the output), b = the third.
//   `split`: the input axis that is sliced across ranks (index / per = destination rank, index % per = local index);
//   axis 0 carries this rank's offset rank*per (the received slabs are ordered by source rank).
struct P2PArgs {
  float* peer[16];
  int d[3];            // input dims
  long long so[3];     // output stride of each INPUT axis (in the destination slab)
  int ay, ab;          // input axes playing y / batch
  int split, per, a0;  // a0: absolute index along axis 0 of the first row of `in` (rank*per for a whole slab)
  float threshold;
};
__global__ void __launch_bounds__(256) transpose_p2p_kernel(const float* __restrict__ in, const P2PArgs a) {
  __shared__ float tile[64][65];
  const int nx = a.d[2], ny = a.d[a.ay];
  const int bx = blockIdx.x * 64, by = blockIdx.y * 64;
  const int bi = blockIdx.z;
  const long long in_sy = a.ay == 1 ? a.d[2] : static_cast<long long>(a.d[1]) * a.d[2];
  const long long in_sb = a.ab == 1 ? a.d[2] : static_cast<long long>(a.d[1]) * a.d[2];
  const int c4 = threadIdx.x & 15, r = threadIdx.x >> 4;
#pragma unroll
  for (int k = 0; k < 64; k += 16) {
    const int x = bx + 4 * c4, y = by + r + k;
    if (x < nx && y < ny) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(in + bi * in_sb + static_cast<long long>(y) * in_sy + x));
      tile[r + k][4 * c4 + 0] = v.x;
      tile[r + k][4 * c4 + 1] = v.y;
      tile[r + k][4 * c4 + 2] = v.z;
      tile[r + k][4 * c4 + 3] = v.w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 64; k += 16) {
    const int x = bx + r + k, y = by + 4 * c4;
    if (x < nx && y < ny) {
      int idx[3];
      idx[2] = x;
      idx[a.ay] = y;
      idx[a.ab] = bi;
      const int dst = idx[a.split] / a.per;
      idx[a.split] -= dst * a.per;
      idx[0] += a.a0;
      float4 v = make_float4(tile[4 * c4 + 0][r + k], tile[4 * c4 + 1][r + k], tile[4 * c4 + 2][r + k], tile[4 * c4 + 3][r + k]);
      v.x = v.x < a.threshold ? 0.0f : v.x;
      v.y = v.y < a.threshold ? 0.0f : v.y;
      v.z = v.z < a.threshold ? 0.0f : v.z;
      v.w = v.w < a.threshold ? 0.0f : v.w;
      *reinterpret_cast<float4*>(a.peer[dst] + idx[0] * a.so[0] + idx[1] * a.so[1] + idx[2] * a.so[2]) = v;
    }
  }
}

// perm[2] == 2: rows stay contiguous, only the two outer axes move (or nothing moves)
__global__ void __launch_bounds__(256)
permute_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int d0, int d1, int d2, long long so0,
                    long long so1, float threshold) {
  const long long row = blockIdx.x;  // input row index = i0 * d1 + i1
  const int i0 = static_cast<int>(row / d1), i1 = static_cast<int>(row % d1);
  const float* src = in + row * d2;
  float* dst = out + i0 * so0 + i1 * so1;
  for (int x = threadIdx.x; x < d2; x += blockDim.x) {
    float v = src[x];
    if (v < threshold) v = 0.0f;
    dst[x] = v;
  }
}

__global__ void __launch_bounds__(256) threshold_kernel(float* v, long long n, float thr) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float x = v[i];
    if (x < thr) v[i] = 0.0f;
  }
}

}  // namespace
}  // namespace mpg

extern "C" {

int mpg_slice_assemble(mpg_handle h, const mpg_assemble_desc* d, const float* vol, const float* dens, int slice0,
                       int count, void* out, void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(h && d && vol && out, "mpg_slice_assemble: null argument");
  int seen[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    MPG_CHECK_ARG(d->axis_of[k] >= 0 && d->axis_of[k] < 3, "assemble: axis_of[%d]=%d", k, d->axis_of[k]);
    seen[d->axis_of[k]]++;
    MPG_CHECK_ARG(d->zoom[k] >= 1 && d->dims[k] >= 1, "assemble: zoom/dims must be >= 1");
  }
  MPG_CHECK_ARG(seen[0] == 1 && seen[1] == 1 && seen[2] == 1, "assemble: axis_of is not a permutation");
  MPG_CHECK_ARG(d->nchan >= 1 && d->nchan <= 8, "assemble: nchan %d not in [1,8]", d->nchan);
  const int total = d->nchan + (dens ? 1 : 0) + (d->add_adj ? 2 : 0);
  MPG_CHECK_ARG(total <= d->out_cstride && total <= 16, "assemble: %d channels > out_cstride %d", total, d->out_cstride);
  AsmParams p;
  for (int k = 0; k < 3; ++k) {
    p.dims[k] = d->dims[k];
    p.axis_of[k] = d->axis_of[k];
    p.zoom[k] = d->zoom[k];
  }
  const long long st[3] = {static_cast<long long>(d->dims[1]) * d->dims[2], d->dims[2], 1};
  for (int k = 0; k < 3; ++k) {
    p.odims[k] = d->dims[d->axis_of[k]] * d->zoom[k];
    p.vstride[k] = st[d->axis_of[k]];
  }
  p.vol_c = d->vol_c;
  p.nchan = d->nchan;
  for (int c = 0; c < 8; ++c) {
    p.chan_src[c] = d->chan_src[c];
    p.chan_scale[c] = d->chan_scale[c];
    if (c < d->nchan) MPG_CHECK_ARG(d->chan_src[c] >= 0 && d->chan_src[c] < d->vol_c, "assemble: chan_src[%d]", c);
  }
  p.add_adj = d->add_adj;
  p.out_dtype = d->out_dtype;
  p.out_cstride = d->out_cstride;
  p.vol = vol;
  p.dens = dens;
  p.slice0 = slice0;
  p.count = count;
  p.dens_slice0 = d->dens_slice0;
  MPG_CHECK_ARG(!dens || slice0 >= d->dens_slice0, "assemble: slice0 %d precedes dens_slice0 %d", slice0, d->dens_slice0);
  p.out = out;
  MPG_CHECK_ARG(slice0 >= 0 && count >= 1 && slice0 + count <= p.odims[0], "assemble: slices [%d,%d) outside [0,%d)",
                slice0, slice0 + count, p.odims[0]);
  const long long plane_px = static_cast<long long>(p.odims[1]) * p.odims[2];
  if (p.vol_c == 4 && p.out_dtype == MPG_F32 && p.out_cstride == 4 && !p.add_adj && total <= 4 && plane_px < (1ll << 31) &&
      ((reinterpret_cast<uintptr_t>(vol) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    slice_assemble_c4_kernel<<<static_cast<unsigned>((plane_px + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    MPG_CUDA(cudaGetLastError());
    return MPG_OK;
  }
  const long long npix = static_cast<long long>(count) * p.odims[1] * p.odims[2];
  slice_assemble_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

int mpg_transpose3d(mpg_handle h, const float* in, float* out, int d0, int d1, int d2, const int perm[3],
                    float threshold, void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(h && in && out && perm, "mpg_transpose3d: null argument");
  MPG_CHECK_ARG(in != out, "mpg_transpose3d: in-place permutation is not supported");
  const int d[3] = {d0, d1, d2};
  int seen[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    MPG_CHECK_ARG(perm[k] >= 0 && perm[k] < 3, "transpose3d: perm[%d]=%d", k, perm[k]);
    seen[perm[k]]++;
  }
  MPG_CHECK_ARG(seen[0] == 1 && seen[1] == 1 && seen[2] == 1, "transpose3d: perm is not a permutation");
  MPG_CHECK_ARG(d0 > 0 && d1 > 0 && d2 > 0, "transpose3d: empty volume");
  const long long od[3] = {d[perm[0]], d[perm[1]], d[perm[2]]};
  const long long st_out[3] = {od[1] * od[2], od[2], 1};
  long long so[3];  // output stride of each INPUT axis
  for (int k = 0; k < 3; ++k) so[perm[k]] = st_out[k];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float thr = threshold > 0.0f ? threshold : -INFINITY;
  if (perm[2] == 2) {
    permute_rows_kernel<<<static_cast<unsigned>(static_cast<long long>(d0) * d1), 256, 0, st>>>(in, out, d0, d1, d2,
                                                                                                so[0], so[1], thr);
  } else {
    const int a = perm[2];
    const int b = 3 - 2 - a;  // the remaining axis (0 or 1)
    const long long st_in[3] = {static_cast<long long>(d1) * d2, d2, 1};
    MPG_CHECK_ARG(d[b] <= 65535, "transpose3d: batch axis %d exceeds 65535", d[b]);
    const bool vec4 = d2 % 4 == 0 && d[a] % 4 == 0 && st_in[a] % 4 == 0 && st_in[b] % 4 == 0 && so[2] % 4 == 0 &&
                      so[b] % 4 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec4) {
      dim3 grid(static_cast<unsigned>(ceil_div(d2, 64)), static_cast<unsigned>(ceil_div(d[a], 64)), static_cast<unsigned>(d[b]));
      transpose_tile64_kernel<<<grid, 256, 0, st>>>(in, out, d2, d[a], st_in[a], st_in[b], so[2], so[b], thr);
    } else {
      dim3 grid(static_cast<unsigned>(ceil_div(d2, 32)), static_cast<unsigned>(ceil_div(d[a], 32)),
                static_cast<unsigned>(d[b]));
      transpose_tile_kernel<<<grid, 256, 0, st>>>(in, out, d2, d[a], st_in[a], st_in[b], so[2], so[b], thr);
    }
  }
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

/* Fused axis change across ranks (see transpose_p2p_kernel). slab: this rank's [S/G, S, S] rows of the finished pass;
 * peer_out[r]: device pointer (peer-mapped) of rank r's output slab; split_axis: 2 = the new slab axis is the old axis 2
 * (received block [A, b, c_loc]), 1 = the old axis 1 ([A, b_loc, c]); final_perm: permutation applied to the received
 * block (as in mpg_transpose3d), final_perm[2] must not be 2. Needs S/G % 4 == 0. The caller brackets the launch with
 * cross-rank barriers (all ranks done reading the previous contents / all stores landed). */
int mpg_reslab_p2p_part(mpg_handle h, const float* part, void* const* peer_out, int world, int S, int a0, int count,
                        int split_axis, const int final_perm[3], float threshold, void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(h && part && peer_out && final_perm, "mpg_reslab_p2p: null argument");
  MPG_CHECK_ARG(world >= 2 && world <= 16 && S % world == 0, "mpg_reslab_p2p: world=%d S=%d", world, S);
  MPG_CHECK_ARG(a0 >= 0 && count >= 1 && a0 + count <= S, "mpg_reslab_p2p: rows [%d, %d) outside [0, %d)", a0, a0 + count, S);
  MPG_CHECK_ARG(split_axis == 1 || split_axis == 2, "mpg_reslab_p2p: split_axis must be 1 or 2");
  const int per = S / world;
  MPG_CHECK_ARG(per % 4 == 0 && S % 4 == 0, "mpg_reslab_p2p: S/G = %d must be a multiple of 4", per);
  int seen[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    MPG_CHECK_ARG(final_perm[k] >= 0 && final_perm[k] < 3, "mpg_reslab_p2p: final_perm[%d]=%d", k, final_perm[k]);
    ++seen[final_perm[k]];
  }
  MPG_CHECK_ARG(seen[0] == 1 && seen[1] == 1 && seen[2] == 1 && final_perm[2] != 2, "mpg_reslab_p2p: unsupported permutation");
  // axis 0 of the part plays the y role when final_perm[2] == 0: its extent must then allow 128-bit stores
  MPG_CHECK_ARG(final_perm[2] != 0 || (count % 4 == 0 && a0 % 4 == 0), "mpg_reslab_p2p: rows [%d, +%d) must be 4-aligned for this permutation", a0, count);
  P2PArgs a;
  for (int r = 0; r < world; ++r) {
    MPG_CHECK_ARG(peer_out[r] && (reinterpret_cast<uintptr_t>(peer_out[r]) & 15) == 0, "mpg_reslab_p2p: peer pointer %d", r);
    a.peer[r] = static_cast<float*>(peer_out[r]);
  }
  a.d[0] = count;
  a.d[1] = S;
  a.d[2] = S;
  // received block dims: split 2 -> (S, S, per); split 1 -> (S, per, S); output = permute(received, final_perm)
  const long long rd[3] = {S, split_axis == 1 ? per : S, split_axis == 2 ? per : S};
  const long long od[3] = {rd[final_perm[0]], rd[final_perm[1]], rd[final_perm[2]]};
  const long long st_out[3] = {od[1] * od[2], od[2], 1};
  for (int k = 0; k < 3; ++k) a.so[final_perm[k]] = st_out[k];
  a.ay = final_perm[2];
  a.ab = 3 - 2 - a.ay;
  a.split = split_axis;
  a.per = per;
  a.a0 = a0;
  a.threshold = threshold > 0.0f ? threshold : -INFINITY;
  MPG_CHECK_ARG((reinterpret_cast<uintptr_t>(part) & 15) == 0, "mpg_reslab_p2p: slab not 16-byte aligned");
  dim3 grid(static_cast<unsigned>(ceil_div(a.d[2], 64)), static_cast<unsigned>(ceil_div(a.d[a.ay], 64)),
            static_cast<unsigned>(a.d[a.ab]));
  DeviceGuard guard(h->device);
  transpose_p2p_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(part, a);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

int mpg_reslab_p2p(mpg_handle h, const float* slab, void* const* peer_out, int world, int rank, int S, int split_axis,
                   const int final_perm[3], float threshold, void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(world >= 2 && rank >= 0 && rank < world && S % world == 0, "mpg_reslab_p2p: world=%d rank=%d S=%d", world, rank, S);
  return mpg_reslab_p2p_part(h, slab, peer_out, world, S, rank * (S / world), S / world, split_axis, final_perm, threshold, stream);
}

int mpg_threshold(mpg_handle h, float* vol, long long count, float threshold, void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(h && vol && count >= 0, "mpg_threshold: bad argument");
  if (count == 0) return MPG_OK;
  long long blocks = (count + 255) / 256;
  const long long cap = static_cast<long long>(h->sm_count) * 16;
  if (blocks > cap) blocks = cap;
  threshold_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(vol, count, threshold);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

}  // extern "C"
