// Training-step kernels (GAN/multipassGAN-4x.py:572-620,744-768,889-902): fp32 NHWC
// forward / dgrad / wgrad convolutions on HWIO weights that live in device memory (they change every Adam step),
// training-mode batch norm, activations, the GAN losses, TF1 Adam and the BN moving-average update.
// C-ABI: mpg_train_* (include/mpg.h).  The training tiles are tiny (16 x 64 x 64 pixels), so these are
// register-tiled CUDA-core kernels; the inference path keeps the tcgen05 kernels.
#include "common.h"

namespace mpg {
namespace {

__host__ __device__ inline int same_pad_before_tf(int in, int k, int s) {
  const int out = (in + s - 1) / s;
  int total = (out - 1) * s + k - in;
  if (total < 0) total = 0;
  return total / 2;
}

// ------------------------------------------------------------------------------------------------------------
// Tiled direct convolution: out[n,oy,ox,co] = sum_{ky,kx,ci} in[n, oy*s - pad + ky, ox*s - pad + kx, ci] * W(ky,kx,ci,co) (+ bias)
// `wmode` 0: W = w[ky][kx][ci][co] (forward, HWIO with wci = cin, wco = cout)
//         1: W = w[k-1-ky][k-1-kx][co][ci]  (stride-1 dgrad: the kernel is run with cin := Cout_fwd, cout := Cin_fwd)
//         2: strided dgrad (discriminator k4 s2 convs): `in` = dy [n,h,w_,cin := Cout_fwd], out = dx [n,oh,ow,cout := Cin_fwd];
//            tap (ky,kx) of output pixel (oy,ox) reads dy at ((oy+pad-ky)/stride, (ox+pad-kx)/stride) when both divide
//            exactly, W = w[ky][kx][co][ci]. Three quarters of the staged taps are zeros at stride 2 -- accepted: the layers
//            are tiny and this keeps them on the tiled/split-K kernel instead of a one-thread-per-element gather.
// Small grids (the 8x8 / 16x16 discriminator maps give 16-64 blocks) are split over K along blockIdx.z and combined
// with atomicAdd into a zeroed output; bias is added by split 0.
// `in_up`: the conv reads a nearest-upsampled view of `in` (in is [n, h/in_up, w/in_up, cin]).
// Block = 256 threads -> 8 x 16 output pixels x 64 output channels; thread tile 4 pixels x 8 channels.
struct ConvArgs {
  const float* in;
  const float* w;
  const float* bias;
  float* out;
  int n, h, w_, cin, cout, k, stride, pad, oh, ow, in_up, wmode, accumulate;
  int its_per_split;  // split-K: K-iterations ((ky,kx,cin chunk) triples) per blockIdx.z; > total = no split
};

constexpr int kTP = 128;  // pixels per block (8 rows x 16 cols)
constexpr int kTC = 64;   // output channels per block
constexpr int kKC = 16;   // input channels per smem chunk

__global__ void __launch_bounds__(256) conv_tiled_kernel(const ConvArgs a) {
  __shared__ float xs[kTP][kKC + 1];
  __shared__ __align__(16) float ws[kKC][kTC];
  const int tiles_x = (a.ow + 15) / 16, tiles_y = (a.oh + 7) / 8;
  int b = blockIdx.x;
  const int tx_ = b % tiles_x;
  b /= tiles_x;
  const int ty_ = b % tiles_y;
  const int n = b / tiles_y;
  const int co0 = blockIdx.y * kTC;
  const int tid = threadIdx.x;
  const int cg = tid & 7;   // channel group: 8 channels
  const int pg = tid >> 3;  // pixel group: 4 pixels (one row segment)
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  const int sh = a.h / a.in_up, sw = a.w_ / a.in_up;
  const int nchunk = (a.cin + kKC - 1) / kKC;
  const int total_its = a.k * a.k * nchunk;
  const int it0 = static_cast<int>(blockIdx.z) * a.its_per_split;
  const int it1 = (it0 + a.its_per_split) < total_its ? (it0 + a.its_per_split) : total_its;
  // each thread stages the same 8 input elements / 4 weight elements every iteration: precompute their coordinates
  for (int it = it0; it < it1; ++it) {
    const int tap = it / nchunk;
    const int c0 = (it - tap * nchunk) * kKC;
    const int ky = tap / a.k, kx = tap - ky * a.k;
    __syncthreads();
    // ---- stage inputs: 128 pixels x 16 channels
    for (int e = tid; e < kTP * kKC; e += 256) {
      const int p = e >> 4, c = e & 15;
      const int oy = ty_ * 8 + (p >> 4), ox = tx_ * 16 + (p & 15);
      float v = 0.0f;
      if (a.wmode == 2) {
        const int ty = oy + a.pad - ky, tx = ox + a.pad - kx;
        if (c0 + c < a.cin && ty >= 0 && tx >= 0 && ty % a.stride == 0 && tx % a.stride == 0) {
          const int iy = ty / a.stride, ix = tx / a.stride;
          if (iy < a.h && ix < a.w_) v = a.in[((static_cast<size_t>(n) * a.h + iy) * a.w_ + ix) * a.cin + c0 + c];
        }
      } else {
        const int iy = oy * a.stride - a.pad + ky, ix = ox * a.stride - a.pad + kx;
        if (c0 + c < a.cin && iy >= 0 && iy < a.h && ix >= 0 && ix < a.w_)
          v = a.in[((static_cast<size_t>(n) * sh + iy / a.in_up) * sw + ix / a.in_up) * a.cin + c0 + c];
      }
      xs[p][c] = v;
    }
    // ---- stage weights: 16 input channels x 64 output channels
    for (int e = tid; e < kKC * kTC; e += 256) {
      const int c = e >> 6, o = e & 63;
      float v = 0.0f;
      if (c0 + c < a.cin && co0 + o < a.cout) {
        if (a.wmode == 0)
          v = a.w[((static_cast<size_t>(ky) * a.k + kx) * a.cin + c0 + c) * a.cout + co0 + o];
        else if (a.wmode == 1)
          v = a.w[((static_cast<size_t>(a.k - 1 - ky) * a.k + (a.k - 1 - kx)) * a.cout + co0 + o) * a.cin + c0 + c];
        else
          v = a.w[((static_cast<size_t>(ky) * a.k + kx) * a.cout + co0 + o) * a.cin + c0 + c];
      }
      ws[c][o] = v;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kKC; ++c) {
      // a thread's 8 output channels are {4 cg .. 4 cg + 3} and {32 + 4 cg ..}: the 8 lanes of a quarter warp read 128
      // contiguous bytes (8 contiguous channels per thread made both 128-bit loads 2-way bank conflicts, as in wgrad_kernel)
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[c][cg * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&ws[c][32 + cg * 4]);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float xv = xs[pg * 4 + i][c];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv, wv[j], acc[i][j]);
      }
    }
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = pg * 4 + i;
    const int oy = ty_ * 8 + (p >> 4), ox = tx_ * 16 + (p & 15);
    if (oy >= a.oh || ox >= a.ow) continue;
    float* o = a.out + ((static_cast<size_t>(n) * a.oh + oy) * a.ow + ox) * a.cout;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = co0 + (j < 4 ? cg * 4 + j : 32 + cg * 4 + (j - 4));
      if (co < a.cout) {
        float v = acc[i][j] + ((a.bias && blockIdx.z == 0) ? a.bias[co] : 0.0f);
        if (split) {
          atomicAdd(&o[co], v);
        } else {
          if (a.accumulate) v += o[co];
          o[co] = v;
        }
      }
    }
  }
}

// wgrad: dW[ky][kx][ci][co] = sum_{n,oy,ox} in[n, oy*s-pad+ky, ox*s-pad+kx, ci] * dY[n,oy,ox,co]
// grid = (taps * ci_tiles * co_tiles, pixel splits); block tile 32 ci x 64 co, thread tile 1 ci x 8 co; atomicAdd.
struct WgradArgs {
  const float* in;
  const float* dy;
  float* dw;
  int n, h, w_, cin, cout, k, stride, pad, oh, ow, in_up, px_per_split;
};
// TCI = input channels per thread (1, 2, 4): block tile 32*TCI ci x 64 co, thread tile TCI ci x 8 co. The first version
// (1 x 8 with the 8 output channels contiguous) was shared-memory bound: 8 FMAs per 36 bytes of LDS and a 2-way bank
// conflict on both 128-bit loads (ncu: 16.8 M conflicts, 26 % issue utilisation, 177 us for 128->256 on 8x8 maps). Now a
// thread's channels are {4 cg .. 4 cg + 3} and {32 + 4 cg ..}: a quarter warp reads 128 contiguous bytes, and the wide layers
// do 32 FMAs per three 128-bit loads.
// PH = pixel phases (narrow inputs, cin <= 16): the 32 thread rows hold 32/PH channels x PH interleaved pixel subsets instead
// of 32 channels (the 2-channel first layer of the discriminator used 2 of 32 rows); the phases are summed through shared
// memory before the atomics.
template <int TCI, int PH>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs a) {
  static_assert(PH == 1 || TCI == 1, "pixel phases only with one channel per thread");
  constexpr int CIB = 32 * TCI / PH;  // input channels per block tile
  constexpr int XST = CIB + 4;        // row stride of xs: keeps 128-bit rows aligned
  __shared__ __align__(16) float xs[32 * XST];
  __shared__ __align__(16) float ys[32][64];
  __shared__ long long xbase[32];  // element offset of the tap's input pixel for each of the 32 staged output pixels, -1 = zero
  const int ci_tiles = (a.cin + CIB - 1) / CIB, co_tiles = (a.cout + 63) / 64;
  int b = blockIdx.x;
  const int cot = b % co_tiles;
  b /= co_tiles;
  const int cit = b % ci_tiles;
  const int tap = b / ci_tiles;
  const int ky = tap / a.k, kx = tap % a.k;
  const int tid = threadIdx.x;
  const int row = tid >> 3, cg = tid & 7;
  const int ci0 = PH == 1 ? row * TCI : row % CIB;  // first channel of this thread inside the block tile
  const int phase = PH == 1 ? 0 : row / CIB;
  const long long npix = static_cast<long long>(a.n) * a.oh * a.ow;
  const long long p_begin = static_cast<long long>(blockIdx.y) * a.px_per_split;
  long long p_end = p_begin + a.px_per_split;
  if (p_end > npix) p_end = npix;
  const int sh = a.h / a.in_up, sw = a.w_ / a.in_up;
  float acc[TCI][8];
#pragma unroll
  for (int t = 0; t < TCI; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.0f;
  for (long long p0 = p_begin; p0 < p_end; p0 += 32) {
    __syncthreads();
    // the pixel -> (n, oy, ox) decomposition once per staged pixel (32 threads), not once per element
    if (tid < 32) {
      const long long p = p0 + tid;
      long long base = -1;
      if (p < p_end) {
        const unsigned pu = static_cast<unsigned>(p);  // n * oh * ow < 2^31 (checked by the host wrapper)
        const int ox = static_cast<int>(pu % static_cast<unsigned>(a.ow));
        const unsigned r = pu / static_cast<unsigned>(a.ow);
        const int oy = static_cast<int>(r % static_cast<unsigned>(a.oh));
        const int n = static_cast<int>(r / static_cast<unsigned>(a.oh));
        const int iy = oy * a.stride - a.pad + ky, ix = ox * a.stride - a.pad + kx;
        if (iy >= 0 && iy < a.h && ix >= 0 && ix < a.w_)
          base = ((static_cast<long long>(n) * sh + iy / a.in_up) * sw + ix / a.in_up) * a.cin + cit * CIB;
      }
      xbase[tid] = base;
    }
    __syncthreads();
    for (int e = tid; e < 32 * CIB; e += 256) {
      const int pl = e / CIB, c = e % CIB;
      const long long base = xbase[pl];
      xs[pl * XST + c] = (base >= 0 && cit * CIB + c < a.cin) ? a.in[base + c] : 0.0f;
    }
    for (int e = tid; e < 32 * 64; e += 256) {
      const int pl = e >> 6, o = e & 63;
      const long long p = p0 + pl;
      float v = 0.0f;
      if (p < p_end && cot * 64 + o < a.cout) v = a.dy[p * a.cout + cot * 64 + o];
      ys[pl][o] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int pl = phase; pl < 32; pl += PH) {
      float xv[TCI];
      if (TCI == 4) {
        const float4 q = *reinterpret_cast<const float4*>(&xs[pl * XST + ci0]);
        xv[0] = q.x, xv[1 % TCI] = q.y, xv[2 % TCI] = q.z, xv[3 % TCI] = q.w;
      } else if (TCI == 2) {
        const float2 q = *reinterpret_cast<const float2*>(&xs[pl * XST + ci0]);
        xv[0] = q.x, xv[1 % TCI] = q.y;
      } else {
        xv[0] = xs[pl * XST + ci0];
      }
      const float4 y0 = *reinterpret_cast<const float4*>(&ys[pl][cg * 4]);
      const float4 y1 = *reinterpret_cast<const float4*>(&ys[pl][32 + cg * 4]);
#pragma unroll
      for (int t = 0; t < TCI; ++t) {
        acc[t][0] = fmaf(xv[t], y0.x, acc[t][0]);
        acc[t][1] = fmaf(xv[t], y0.y, acc[t][1]);
        acc[t][2] = fmaf(xv[t], y0.z, acc[t][2]);
        acc[t][3] = fmaf(xv[t], y0.w, acc[t][3]);
        acc[t][4] = fmaf(xv[t], y1.x, acc[t][4]);
        acc[t][5] = fmaf(xv[t], y1.y, acc[t][5]);
        acc[t][6] = fmaf(xv[t], y1.z, acc[t][6]);
        acc[t][7] = fmaf(xv[t], y1.w, acc[t][7]);
      }
    }
  }
  if (PH == 1) {
#pragma unroll
    for (int t = 0; t < TCI; ++t) {
      const int ci = cit * CIB + ci0 + t;
      if (ci >= a.cin) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int co = cot * 64 + (j < 4 ? cg * 4 + j : 32 + cg * 4 + (j - 4));
        if (co < a.cout) atomicAdd(&a.dw[(static_cast<size_t>(tap) * a.cin + ci) * a.cout + co], acc[t][j]);
      }
    }
  } else {
    // sum the pixel phases: row = phase * CIB + channel
    __syncthreads();
    float* red = &ys[0][0];
#pragma unroll
    for (int j = 0; j < 8; ++j) red[row * 64 + (j < 4 ? cg * 4 + j : 32 + cg * 4 + (j - 4))] = acc[0][j];
    __syncthreads();
    for (int e = tid; e < CIB * 64; e += 256) {
      const int c = e >> 6, o = e & 63;
      const int ci = cit * CIB + c, co = cot * 64 + o;
      if (ci >= a.cin || co >= a.cout) continue;
      float sum = 0.0f;
#pragma unroll
      for (int ph = 0; ph < PH; ++ph) sum += red[(ph * CIB + c) * 64 + o];
      atomicAdd(&a.dw[(static_cast<size_t>(tap) * a.cin + ci) * a.cout + co], sum);
    }
  }
}

template <int TCI, int PH>
static void launch_wgrad(mpg_handle h, WgradArgs& a, long long npix, cudaStream_t st) {
  constexpr int CIB = 32 * TCI / PH;
  const int tiles = a.k * a.k * ((a.cin + CIB - 1) / CIB) * ((a.cout + 63) / 64);
  int splits = (h->sm_count * 4 + tiles - 1) / tiles;
  if (splits < 1) splits = 1;
  long long per = (npix + splits - 1) / splits;
  per = (per + 31) / 32 * 32;
  if (per < 32) per = 32;
  splits = static_cast<int>((npix + per - 1) / per);
  a.px_per_split = static_cast<int>(per);
  wgrad_kernel<TCI, PH><<<dim3(static_cast<unsigned>(tiles), static_cast<unsigned>(splits)), 256, 0, st>>>(a);
}

// Filter gradient of the THIN stride-1 convolutions (cout <= 32, any cin; generator layers 4->8, 8->32, 32->8, 8->2,
// 2->1 and the narrow 1x1 shortcuts): thread = one (tap, ci) pair holding all cout partial sums in registers, block =
// 8 x 16 output pixels staged in shared memory (input window with halo for <= CI channels, dy tile), persistent over
// tiles so every block issues its atomics once per channel chunk. The generic wgrad_kernel above tiles 32 ci x 64 co
// per tap, which leaves >90 % of its threads idle on these shapes (1.1 ms per layer at 16 x 64 x 64).
constexpr int kWtTH = 8, kWtTW = 16;
constexpr int kWtXsFloats = 4352;
template <int CO>
__global__ void __launch_bounds__(256) wgrad_thin_kernel(const WgradArgs a, int ci_chunk) {
  __shared__ float xs[kWtXsFloats];
  __shared__ __align__(16) float dys[kWtTH * kWtTW][CO];
  const int k = a.k, pad = a.pad;
  const int hw = kWtTW + k - 1, hh = kWtTH + k - 1;
  const int tiles_x = (a.ow + kWtTW - 1) / kWtTW, tiles_y = (a.oh + kWtTH - 1) / kWtTH;
  const int ntiles = a.n * tiles_x * tiles_y;
  const int sh = a.h / a.in_up, sw = a.w_ / a.in_up;
  const int tid = threadIdx.x;
  const int npairs = k * k * ci_chunk;
  const int tap = tid / ci_chunk, cl = tid - tap * ci_chunk;
  const int ky = tap / k, kx = tap - ky * k;
  const bool active = tid < npairs;
  for (int c0 = 0; c0 < a.cin; c0 += ci_chunk) {
    float acc[CO];
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[j] = 0.0f;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int n = t / (tiles_x * tiles_y);
      const int r = t - n * tiles_x * tiles_y;
      const int ty = r / tiles_x, tx = r - ty * tiles_x;
      const int oy0 = ty * kWtTH, ox0 = tx * kWtTW;
      __syncthreads();
      for (int e = tid; e < hh * hw * ci_chunk; e += 256) {
        const int c = e % ci_chunk, q = e / ci_chunk;
        const int wy = q / hw, wx = q - wy * hw;
        const int iy = oy0 - pad + wy, ix = ox0 - pad + wx;
        float v = 0.0f;
        if (c0 + c < a.cin && iy >= 0 && iy < a.h && ix >= 0 && ix < a.w_)
          v = a.in[((static_cast<size_t>(n) * sh + iy / a.in_up) * sw + ix / a.in_up) * a.cin + c0 + c];
        xs[q * ci_chunk + c] = v;
      }
      for (int e = tid; e < kWtTH * kWtTW * CO; e += 256) {
        const int co = e % CO, q = e / CO;
        const int oy = oy0 + q / kWtTW, ox = ox0 + q % kWtTW;
        float v = 0.0f;
        if (co < a.cout && oy < a.oh && ox < a.ow) v = a.dy[((static_cast<size_t>(n) * a.oh + oy) * a.ow + ox) * a.cout + co];
        dys[q][co] = v;
      }
      __syncthreads();
      if (active) {
        for (int py = 0; py < kWtTH; ++py) {
          const float* xr = xs + ((py + ky) * hw + kx) * ci_chunk + cl;
#pragma unroll 4
          for (int px = 0; px < kWtTW; ++px) {
            const float xv = xr[px * ci_chunk];
            const float4* d4 = reinterpret_cast<const float4*>(&dys[py * kWtTW + px][0]);
#pragma unroll
            for (int j4 = 0; j4 < CO / 4; ++j4) {
              const float4 d = d4[j4];
              acc[j4 * 4 + 0] = fmaf(xv, d.x, acc[j4 * 4 + 0]);
              acc[j4 * 4 + 1] = fmaf(xv, d.y, acc[j4 * 4 + 1]);
              acc[j4 * 4 + 2] = fmaf(xv, d.z, acc[j4 * 4 + 2]);
              acc[j4 * 4 + 3] = fmaf(xv, d.w, acc[j4 * 4 + 3]);
            }
          }
        }
      }
    }
    if (active && c0 + cl < a.cin) {
      float* o = a.dw + (static_cast<size_t>(tap) * a.cin + c0 + cl) * a.cout;
#pragma unroll
      for (int j = 0; j < CO; ++j)
        if (j < a.cout) atomicAdd(&o[j], acc[j]);
    }
  }
}

// per-channel double-precision sums over `rows` rows of a [rows, c] matrix: out[0..c) += sum a*b?, out[c..2c) ...
__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == MPG_ACT_RELU) return fmaxf(z, 0.0f);
  if (act == MPG_ACT_LRELU) return 0.6f * z + 0.4f * fabsf(z);  // tools_wscale/GAN.py:733-737
  if (act == MPG_ACT_TANH) return tanhf(z);
  return z;
}
// derivative expressed through the OUTPUT y = act(z) (relu / lrelu keep the sign of z)
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
  if (act == MPG_ACT_RELU) return y > 0.0f ? 1.0f : 0.0f;
  if (act == MPG_ACT_LRELU) return y > 0.0f ? 1.0f : (y < 0.0f ? 0.2f : 0.6f);  // d/dz (0.6 z + 0.4 |z|), 0.6 at z = 0
  if (act == MPG_ACT_TANH) return 1.0f - y * y;
  return 1.0f;
}

// mode 0: s0 = sum x, s1 = sum x^2          (BN statistics)
// mode 1: s0 = sum dz, s1 = sum dz * xhat    (BN backward; xhat = (x - mean) * invstd); with yact != NULL the `dz`
//         argument is dy, the gradient w.r.t. the ACTIVATED output, and dz = dy * act'(yact) is formed on the fly
// mode 2: s0 = sum x                         (bias gradient)
__global__ void __launch_bounds__(256) colstats_kernel(const float* x, const float* dz, const float* mean,
                                                        const float* invstd, double* out, long long rows, int c,
                                                        int mode, const float* yact = nullptr, int act = 0) {
  // Vector path (c % 4 == 0, c <= 1024): a thread owns 4 adjacent channels (one 128-bit load per row) and walks the rows
  // with 4 independent loads in flight; the block combines its row groups in shared memory, so every block issues ONE
  // double atomic per channel and statistic. (The scalar version below kept 4 x 4 bytes per thread in flight and was
  // latency bound at ~30 % of the HBM rate: 34 us for a 33 MB layer, 68 launches per loop body.)
  if ((c & 3) == 0 && c <= 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (mode != 1 || ((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(yact)) & 15) == 0)) {
    __shared__ double red[2][256][4];
    const int groups = c >> 2;                       // float4 groups per row
    const int lanes = groups < 256 ? groups : 256;   // threads per row
    const int rows_per_iter = 256 / lanes;
    const int gl = threadIdx.x % lanes, r_l = threadIdx.x / lanes;
    const bool active = r_l < rows_per_iter;
    const long long rstep = static_cast<long long>(gridDim.x) * rows_per_iter;
    for (int g0 = 0; g0 < groups; g0 += lanes) {
      const int g = g0 + gl;
      double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
      if (active && g < groups) {
        const float4* xp = reinterpret_cast<const float4*>(x) + g;
        const float4* dp = reinterpret_cast<const float4*>(dz) + g;
        const float4* yp = reinterpret_cast<const float4*>(yact) + g;
        float4 mu = make_float4(0, 0, 0, 0), is = make_float4(1, 1, 1, 1);
        if (mode == 1) {
          mu = reinterpret_cast<const float4*>(mean)[g];
          is = reinterpret_cast<const float4*>(invstd)[g];
        }
        long long r = static_cast<long long>(blockIdx.x) * rows_per_iter + r_l;
        for (; r < rows; r += 4 * rstep) {
          float4 v[4], d[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const long long rr = r + j * rstep;
            v[j] = rr < rows ? __ldg(xp + rr * groups) : make_float4(0, 0, 0, 0);
            if (mode == 1) {
              d[j] = rr < rows ? __ldg(dp + rr * groups) : make_float4(0, 0, 0, 0);
              if (yact != nullptr && rr < rows) {
                const float4 ya = __ldg(yp + rr * groups);
                d[j].x *= act_grad_from_out(ya.x, act);
                d[j].y *= act_grad_from_out(ya.y, act);
                d[j].z *= act_grad_from_out(ya.z, act);
                d[j].w *= act_grad_from_out(ya.w, act);
              }
            }
          }
          // fp32 partial sums over the 4 rows in flight, accumulated in double across iterations
          float a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            if (mode == 1) {
              const float dv[4] = {d[j].x, d[j].y, d[j].z, d[j].w};
              const float m4[4] = {mu.x, mu.y, mu.z, mu.w}, i4[4] = {is.x, is.y, is.z, is.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                a0[q] += dv[q];
                a1[q] = fmaf(dv[q], (xv[q] - m4[q]) * i4[q], a1[q]);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                a0[q] += xv[q];
                if (mode == 0) a1[q] = fmaf(xv[q], xv[q], a1[q]);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            s0[q] += a0[q];
            s1[q] += a1[q];
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        red[0][threadIdx.x][q] = s0[q];
        red[1][threadIdx.x][q] = s1[q];
      }
      __syncthreads();
      // threads of row group 0 fold the other row groups of their channel quad
      if (r_l == 0 && g < groups) {
        for (int o = 1; o < rows_per_iter; ++o)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            s0[q] += red[0][o * lanes + gl][q];
            s1[q] += red[1][o * lanes + gl][q];
          }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          atomicAdd(&out[4 * g + q], s0[q]);
          if (mode != 2) atomicAdd(&out[c + 4 * g + q], s1[q]);
        }
      }
      __syncthreads();
    }
    return;
  }
  // thread -> channel (tid % c) when c <= 256, rows strided
  const int lanes_per_row = c < 256 ? c : 256;
  const int rows_per_iter = 256 / lanes_per_row;
  const int ch_l = threadIdx.x % lanes_per_row;
  const int r_l = threadIdx.x / lanes_per_row;
  if (r_l >= rows_per_iter) return;
  for (int ch = ch_l; ch < c; ch += lanes_per_row) {
    double s0 = 0.0, s1 = 0.0;
    const long long rstep = static_cast<long long>(gridDim.x) * rows_per_iter;
    long long r = static_cast<long long>(blockIdx.x) * rows_per_iter + r_l;
    for (; r < rows; r += rstep) {
      const float xv = x[r * c + ch];
      if (mode == 0) {
        s0 += xv;
        s1 += static_cast<double>(xv) * xv;
      } else if (mode == 1) {
        float d = dz[r * c + ch];
        if (yact != nullptr) d *= act_grad_from_out(yact[r * c + ch], act);
        s0 += d;
        s1 += static_cast<double>(d) * ((xv - mean[ch]) * invstd[ch]);
      } else {
        s0 += xv;
      }
    }
    atomicAdd(&out[ch], s0);
    if (mode != 2) atomicAdd(&out[c + ch], s1);
  }
}

// finalize BN statistics: mean, biased variance, invstd; EMA of the moving statistics (tf.contrib batch_norm)
__global__ void bn_finalize_kernel(const double* sums, float* mean, float* var, float* invstd, float* moving_mean,
                                   float* moving_var, long long rows, int c, float eps, float decay) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double m = sums[ch] / static_cast<double>(rows);
  double v = sums[c + ch] / static_cast<double>(rows) - m * m;
  if (v < 0.0) v = 0.0;
  mean[ch] = static_cast<float>(m);
  var[ch] = static_cast<float>(v);
  invstd[ch] = static_cast<float>(1.0 / sqrt(v + static_cast<double>(eps)));
  if (moving_mean) {  // moving = moving * decay + batch * (1 - decay)
    moving_mean[ch] = moving_mean[ch] * decay + static_cast<float>(m) * (1.0f - decay);
    moving_var[ch] = moving_var[ch] * decay + static_cast<float>(v) * (1.0f - decay);
  }
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* x, const float* gamma, const float* beta,
                                                        const float* mean, const float* invstd, float* y,
                                                        long long total, int c, int act) {
  // 128-bit path: c % 4 == 0 keeps a float4 inside one pixel, the channel index is 32-bit arithmetic (the scalar form did one
  // 64-bit modulo and one 4-byte load per element: these element-wise passes were a third of the training loop body)
  if ((c & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 && total < (1LL << 33)) {
    const unsigned c4 = static_cast<unsigned>(c) >> 2;
    const long long t4 = total >> 2;
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < t4; e += static_cast<long long>(gridDim.x) * 256) {
      const int ch = static_cast<int>(static_cast<unsigned>(e % c4)) * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + e);
      const float4 m = *reinterpret_cast<const float4*>(mean + ch), is = *reinterpret_cast<const float4*>(invstd + ch);
      const float4 g = *reinterpret_cast<const float4*>(gamma + ch), b = *reinterpret_cast<const float4*>(beta + ch);
      float4 o;
      o.x = act_fwd((v.x - m.x) * is.x * g.x + b.x, act);
      o.y = act_fwd((v.y - m.y) * is.y * g.y + b.y, act);
      o.z = act_fwd((v.z - m.z) * is.z * g.z + b.z, act);
      o.w = act_fwd((v.w - m.w) * is.w * g.w + b.w, act);
      reinterpret_cast<float4*>(y)[e] = o;
    }
    return;
  }
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const int ch = static_cast<int>(e % c);
    const float z = (x[e] - mean[ch]) * invstd[ch] * gamma[ch] + beta[ch];
    y[e] = act_fwd(z, act);
  }
}

// bn_finalize_kernel + bn_apply_kernel in one launch (c % 4 == 0, c <= 1024): every block derives mean / invstd of all channels
// from the double sums (same arithmetic as bn_finalize_kernel) into shared memory, block 0 also publishes them and updates the
// moving statistics; the element pass is bn_apply_kernel's 128-bit path.
__global__ void __launch_bounds__(256) bn_finalize_apply_kernel(const float* x, const float* gamma, const float* beta,
                                                                 const double* sums, float* mean, float* var, float* invstd,
                                                                 float* moving_mean, float* moving_var, float* y,
                                                                 long long rows, int c, float eps, float decay, int act) {
  __shared__ __align__(16) float s_mean[1024];
  __shared__ __align__(16) float s_inv[1024];
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    const double m = sums[ch] / static_cast<double>(rows);
    double v = sums[c + ch] / static_cast<double>(rows) - m * m;
    if (v < 0.0) v = 0.0;
    const float is = static_cast<float>(1.0 / sqrt(v + static_cast<double>(eps)));
    s_mean[ch] = static_cast<float>(m);
    s_inv[ch] = is;
    if (blockIdx.x == 0) {
      mean[ch] = static_cast<float>(m);
      var[ch] = static_cast<float>(v);
      invstd[ch] = is;
      if (moving_mean) {  // moving = moving * decay + batch * (1 - decay)
        moving_mean[ch] = moving_mean[ch] * decay + static_cast<float>(m) * (1.0f - decay);
        moving_var[ch] = moving_var[ch] * decay + static_cast<float>(v) * (1.0f - decay);
      }
    }
  }
  __syncthreads();
  const unsigned c4 = static_cast<unsigned>(c) >> 2;
  const long long t4 = (rows * c) >> 2;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < t4; e += static_cast<long long>(gridDim.x) * 256) {
    const int ch = static_cast<int>(static_cast<unsigned>(e % c4)) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + e);
    const float4 m = *reinterpret_cast<const float4*>(s_mean + ch), is = *reinterpret_cast<const float4*>(s_inv + ch);
    const float4 g = *reinterpret_cast<const float4*>(gamma + ch), b = *reinterpret_cast<const float4*>(beta + ch);
    float4 o;
    o.x = act_fwd((v.x - m.x) * is.x * g.x + b.x, act);
    o.y = act_fwd((v.y - m.y) * is.y * g.y + b.y, act);
    o.z = act_fwd((v.z - m.z) * is.z * g.z + b.z, act);
    o.w = act_fwd((v.w - m.w) * is.w * g.w + b.w, act);
    reinterpret_cast<float4*>(y)[e] = o;
  }
}

// dz = dy * act'(y)   (in place allowed)
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* y, const float* dy, float* dz, long long total, int act) {
  if ((total & 3) == 0 && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0) {
    const long long t4 = total >> 2;
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < t4; e += static_cast<long long>(gridDim.x) * 256) {
      const float4 a = reinterpret_cast<const float4*>(y)[e], d = reinterpret_cast<const float4*>(dy)[e];
      float4 o;
      o.x = d.x * act_grad_from_out(a.x, act);
      o.y = d.y * act_grad_from_out(a.y, act);
      o.z = d.z * act_grad_from_out(a.z, act);
      o.w = d.w * act_grad_from_out(a.w, act);
      reinterpret_cast<float4*>(dz)[e] = o;
    }
    return;
  }
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256)
    dz[e] = dy[e] * act_grad_from_out(y[e], act);
}

// dx = gamma * invstd / N * (N * dz - dbeta - xhat * dgamma);  dgamma/dbeta are read from `sums` (double)
// `dz` is dy (gradient w.r.t. the activated output) and dz = dy * act'(y) is formed here when y != NULL
__global__ void __launch_bounds__(256) bn_bwd_kernel(const float* x, const float* dz, const float* gamma,
                                                      const float* mean, const float* invstd, const double* sums,
                                                      float* dx, float* dgamma, float* dbeta, long long rows, int c,
                                                      const float* y, int act) {
  const long long total = rows * c;
  const float inv_n = 1.0f / static_cast<float>(rows);
  if ((c & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dx) |
                        reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    const unsigned c4 = static_cast<unsigned>(c) >> 2;
    const long long t4 = total >> 2;
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < t4; e += static_cast<long long>(gridDim.x) * 256) {
      const int ch = static_cast<int>(static_cast<unsigned>(e % c4)) * 4;
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + e);
      float4 dv = reinterpret_cast<const float4*>(dz)[e];
      if (y != nullptr) {
        const float4 ya = __ldg(reinterpret_cast<const float4*>(y) + e);
        dv.x *= act_grad_from_out(ya.x, act);
        dv.y *= act_grad_from_out(ya.y, act);
        dv.z *= act_grad_from_out(ya.z, act);
        dv.w *= act_grad_from_out(ya.w, act);
      }
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, da[4] = {dv.x, dv.y, dv.z, dv.w};
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float db = static_cast<float>(sums[ch + j]), dg = static_cast<float>(sums[c + ch + j]);
        const float xh = (xa[j] - mean[ch + j]) * invstd[ch + j];
        o[j] = gamma[ch + j] * invstd[ch + j] * (da[j] - inv_n * (db + xh * dg));
      }
      reinterpret_cast<float4*>(dx)[e] = make_float4(o[0], o[1], o[2], o[3]);
      if (e * 4 < c) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dgamma[e * 4 + j] += static_cast<float>(sums[c + e * 4 + j]);
          dbeta[e * 4 + j] += static_cast<float>(sums[e * 4 + j]);
        }
      }
    }
    return;
  }
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const int ch = static_cast<int>(e % c);
    const float db = static_cast<float>(sums[ch]), dg = static_cast<float>(sums[c + ch]);
    const float xh = (x[e] - mean[ch]) * invstd[ch];
    const float d = y != nullptr ? dz[e] * act_grad_from_out(y[e], act) : dz[e];
    dx[e] = gamma[ch] * invstd[ch] * (d - inv_n * (db + xh * dg));
    if (e < c) {
      dgamma[e] += static_cast<float>(sums[c + e]);
      dbeta[e] += static_cast<float>(sums[e]);
    }
  }
}

__global__ void __launch_bounds__(256) bias_act_kernel(const float* x, float* y, long long total, int act) {
  if ((total & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    const long long t4 = total >> 2;
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < t4; e += static_cast<long long>(gridDim.x) * 256) {
      const float4 p = reinterpret_cast<const float4*>(x)[e];
      reinterpret_cast<float4*>(y)[e] = make_float4(act_fwd(p.x, act), act_fwd(p.y, act), act_fwd(p.z, act), act_fwd(p.w, act));
    }
    return;
  }
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256)
    y[e] = act_fwd(x[e], act);
}
__global__ void __launch_bounds__(256) add_act_kernel(const float* a, const float* b, float* y, long long total, int act) {
  if ((total & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    const long long t4 = total >> 2;
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < t4; e += static_cast<long long>(gridDim.x) * 256) {
      const float4 p = reinterpret_cast<const float4*>(a)[e], q = reinterpret_cast<const float4*>(b)[e];
      reinterpret_cast<float4*>(y)[e] = make_float4(act_fwd(p.x + q.x, act), act_fwd(p.y + q.y, act), act_fwd(p.z + q.z, act),
                                                    act_fwd(p.w + q.w, act));
    }
    return;
  }
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256)
    y[e] = act_fwd(a[e] + b[e], act);
}
__global__ void __launch_bounds__(256) axpy_kernel(float* y, const float* x, float alpha, long long total) {
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256)
    y[e] += alpha * x[e];
}
__global__ void __launch_bounds__(256) mul_kernel(float* out, const float* a, const float* b, long long total) {
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256)
    out[e] = a[e] * b[e];
}
// ---- pieces of the 8x progressive-growing trainer (GAN/multipassGAN-8x.py:596-597 lerp, tools_wscale/GAN.py:162-169
//      avg_pool, :1101-1143 WGAN-GP terms)
__global__ void __launch_bounds__(256) avgpool2_fwd_kernel(const float* x, float* y, int n, int oh, int ow, int c) {
  const long long total = static_cast<long long>(n) * oh * ow * c;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const int ch = static_cast<int>(e % c);
    long long r = e / c;
    const int ox = static_cast<int>(r % ow);
    r /= ow;
    const int oy = static_cast<int>(r % oh);
    const long long img = r / oh;
    const long long row = static_cast<long long>(2 * ow) * c;
    const float* p = x + ((img * (2 * oh) + 2 * oy) * (2 * ow) + 2 * ox) * c + ch;
    y[e] = 0.25f * ((p[0] + p[c]) + (p[row] + p[row + c]));
  }
}
__global__ void __launch_bounds__(256) avgpool2_bwd_kernel(const float* dy, float* dx, int n, int oh, int ow, int c,
                                                            int accumulate) {
  const long long total = static_cast<long long>(n) * (2 * oh) * (2 * ow) * c;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const int ch = static_cast<int>(e % c);
    long long r = e / c;
    const int x = static_cast<int>(r % (2 * ow));
    r /= 2 * ow;
    const int y = static_cast<int>(r % (2 * oh));
    const long long img = r / (2 * oh);
    const float g = 0.25f * dy[((img * oh + (y >> 1)) * ow + (x >> 1)) * c + ch];
    dx[e] = accumulate ? dx[e] + g : g;
  }
}
__global__ void __launch_bounds__(256) lerp_kernel(float* out, const float* a, const float* b, float t, long long total) {
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256)
    out[e] = a[e] + (b[e] - a[e]) * t;
}
__global__ void __launch_bounds__(256) scale_kernel(float* y, const float* x, float alpha, long long total) {
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256)
    y[e] = alpha * x[e];
}
// strided pick / zero-stuffed 16-bit scatter: a stride-2 conv = the stride-1 conv sampled at even positions, its input gradient
// = the stride-1 dgrad of dy scattered onto those positions (mpg_conv_plan_update_ex)
__global__ void __launch_bounds__(256) pick_kernel(const float* in, float* out, int n, int oh, int ow, int c, int stride,
                                                    int in_cstride, int out_cstride, int out_c0) {
  const long long total = static_cast<long long>(n) * oh * ow * c;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const int ch = static_cast<int>(e % c);
    long long r = e / c;
    const int ox = static_cast<int>(r % ow);
    r /= ow;
    const int oy = static_cast<int>(r % oh);
    const long long img = r / oh;
    const long long src = ((img * (oh * stride) + static_cast<long long>(oy) * stride) * (ow * stride) + static_cast<long long>(ox) * stride) * in_cstride + ch;
    out[((img * oh + oy) * ow + ox) * out_cstride + out_c0 + ch] = in[src];
  }
}
__global__ void __launch_bounds__(256) stuff16_kernel(const float* dy, uint16_t* out, int dtype, int n, int oh, int ow, int c,
                                                       int stride, int cs) {
  const int h = oh * stride, w = ow * stride;
  const long long total = static_cast<long long>(n) * h * w * cs;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const int ch = static_cast<int>(e % cs);
    long long r = e / cs;
    const int x = static_cast<int>(r % w);
    r /= w;
    const int y = static_cast<int>(r % h);
    const long long img = r / h;
    float v = 0.0f;
    if (ch < c && y % stride == 0 && x % stride == 0) v = dy[((img * oh + y / stride) * ow + x / stride) * c + ch];
    out[e] = float_to_h16(v, dtype);
  }
}
// pixel_norm (tools_wscale/GAN.py:472-474): y = x * rsqrt(mean_c x^2 + 1e-8), one thread per pixel; backward:
// dx = r dy - x r^3 mean_c(dy x)
__global__ void __launch_bounds__(256) pixel_norm_fwd_kernel(const float* x, float* y, long long rows, int c) {
  for (long long p = blockIdx.x * 256LL + threadIdx.x; p < rows; p += static_cast<long long>(gridDim.x) * 256) {
    const float* xp = x + p * c;
    float s = 0.0f;
    for (int i = 0; i < c; ++i) s = fmaf(xp[i], xp[i], s);
    const float r = rsqrtf(s / static_cast<float>(c) + 1e-8f);
    for (int i = 0; i < c; ++i) y[p * c + i] = xp[i] * r;
  }
}
__global__ void __launch_bounds__(256) pixel_norm_bwd_kernel(const float* x, const float* dy, float* dx, long long rows, int c) {
  for (long long p = blockIdx.x * 256LL + threadIdx.x; p < rows; p += static_cast<long long>(gridDim.x) * 256) {
    const float* xp = x + p * c;
    const float* gp = dy + p * c;
    float s = 0.0f, d = 0.0f;
    for (int i = 0; i < c; ++i) {
      s = fmaf(xp[i], xp[i], s);
      d = fmaf(gp[i], xp[i], d);
    }
    const float r = rsqrtf(s / static_cast<float>(c) + 1e-8f);
    const float k = r * r * r * d / static_cast<float>(c);
    for (int i = 0; i < c; ++i) dx[p * c + i] = r * gp[i] - xp[i] * k;
  }
}
// one block per sample: norm = sqrt(sum (g + 1e-4)^2); loss += lambda * (norm - target)^2 / rows;
// v = d loss / d g = (2 * lambda / rows) * (norm - target) * (g + 1e-4) / norm      (GAN/multipassGAN-8x.py:1130-1133)
__global__ void __launch_bounds__(256) gp_penalty_kernel(const float* g, float* v, double* loss, float* norms, int rows,
                                                          long long n, float lambda, float target) {
  __shared__ double red[256];
  const int b = blockIdx.x;
  const float* gb = g + static_cast<long long>(b) * n;
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const double t = static_cast<double>(gb[i]) + 1e-4;
    s += t * t;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double norm = sqrt(red[0]);
  const double coef = 2.0 * lambda / rows * (norm - target) / norm;
  for (long long i = threadIdx.x; i < n; i += 256)
    v[static_cast<long long>(b) * n + i] = static_cast<float>(coef * (static_cast<double>(gb[i]) + 1e-4));
  if (threadIdx.x == 0) {
    atomicAdd(loss, static_cast<double>(lambda) * (norm - target) * (norm - target) / rows);
    if (norms) norms[b] = static_cast<float>(norm);
  }
}
// loss += scale * mean(x^power) (power 1 or 2), dx (+)= its gradient: the WGAN critic terms mean(-disc), mean(gen),
// wgan_epsilon * mean(disc^2) (:1111-1112, 1140-1141)
__global__ void __launch_bounds__(256) mean_pow_kernel(const float* x, float scale, int power, double* loss, float* dx,
                                                        long long total, int accumulate) {
  __shared__ double red[256];
  double s = 0.0;
  for (long long e = threadIdx.x; e < total; e += 256) {
    const float v = x[e];
    s += power == 2 ? static_cast<double>(v) * v : static_cast<double>(v);
    const float g = scale / static_cast<float>(total) * (power == 2 ? 2.0f * v : 1.0f);
    if (dx) dx[e] = accumulate ? dx[e] + g : g;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(loss, static_cast<double>(scale) * red[0] / static_cast<double>(total));
}

__global__ void __launch_bounds__(256) dsum_to_f32_kernel(const double* s, float* out, int c, int accumulate) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < c) out[i] = (accumulate ? out[i] : 0.0f) + static_cast<float>(s[i]);
}

// ---- losses: each writes dloss/dinput (scaled) and atomically adds the loss value (double) to *loss
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[8];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 8) r = sh[threadIdx.x];
  if (threadIdx.x < 32)
    for (int o = 4; o > 0; o >>= 1) r += __shfl_down_sync(0xFFFFFFFFu, r, o);
  __syncthreads();
  return r;
}
// tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log(1 + exp(-|x|)); mean over `count`, times scale
__global__ void __launch_bounds__(256) bce_kernel(const float* x, float z, float scale, double* loss, float* dx,
                                                   long long count, int accumulate) {
  double part = 0.0;
  const float inv = scale / static_cast<float>(count);
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < count; e += static_cast<long long>(gridDim.x) * 256) {
    const float v = x[e];
    part += static_cast<double>(fmaxf(v, 0.0f) - v * z + log1pf(expf(-fabsf(v))));
    const float g = (1.0f / (1.0f + expf(-v)) - z) * inv;
    if (dx) dx[e] = (accumulate ? dx[e] : 0.0f) + g;
  }
  part = block_sum(part);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, part * static_cast<double>(inv));
}
// mean |y - g| * scale ; dg = -sign(y - g) * scale / count
__global__ void __launch_bounds__(256) l1_kernel(const float* y, const float* g, float scale, double* loss, float* dg,
                                                  long long count, int accumulate) {
  double part = 0.0;
  const float inv = scale / static_cast<float>(count);
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < count; e += static_cast<long long>(gridDim.x) * 256) {
    const float d = y[e] - g[e];
    part += static_cast<double>(fabsf(d));
    const float s = d > 0.0f ? -1.0f : (d < 0.0f ? 1.0f : 0.0f);
    if (dg) dg[e] = (accumulate ? dg[e] : 0.0f) + s * inv;
  }
  part = block_sum(part);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, part * static_cast<double>(inv));
}
// tf.nn.l2_loss(a - b) * scale = 0.5 * sum (a-b)^2 * scale ; db = -(a - b) * scale
__global__ void __launch_bounds__(256) l2half_kernel(const float* a, const float* b, float scale, double* loss, float* db,
                                                      long long count, int accumulate) {
  double part = 0.0;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < count; e += static_cast<long long>(gridDim.x) * 256) {
    const float d = a[e] - b[e];
    part += 0.5 * static_cast<double>(d) * d;
    if (db) db[e] = (accumulate ? db[e] : 0.0f) - d * scale;
  }
  part = block_sum(part);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, part * static_cast<double>(scale));
}

// TF1 Adam ("epsilon hat"): m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr_t * m / (sqrt(v) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float* p, const float* g, float* m, float* v, long long count,
                                                    float lr_t, float b1, float b2, float eps) {
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < count; e += static_cast<long long>(gridDim.x) * 256) {
    const float gv = g[e];
    const float mv = b1 * m[e] + (1.0f - b1) * gv;
    const float vv = b2 * v[e] + (1.0f - b2) * gv * gv;
    m[e] = mv;
    v[e] = vv;
    p[e] -= lr_t * mv / (sqrtf(vv) + eps);
  }
}

// same update with the bias-corrected step size read from device memory (CUDA-graph replays: the value changes every
// step, a by-value kernel argument would be frozen at capture time)
__global__ void __launch_bounds__(256) adam_dev_kernel(float* p, const float* g, float* m, float* v, long long count,
                                                        const float* lr_t_dev, float b1, float b2, float eps) {
  const float lr_t = *lr_t_dev;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < count; e += static_cast<long long>(gridDim.x) * 256) {
    const float gv = g[e];
    const float mv = b1 * m[e] + (1.0f - b1) * gv;
    const float vv = b2 * v[e] + (1.0f - b2) * gv * gv;
    m[e] = mv;
    v[e] = vv;
    p[e] -= lr_t * mv / (sqrtf(vv) + eps);
  }
}

// fully connected head [rows, nin] x [nin] -> [rows]: forward, dgrad, wgrad
__global__ void __launch_bounds__(256) fc_fwd_kernel(const float* x, const float* w, const float* bias, float* y, int nin) {
  double part = 0.0;
  const float* xr = x + static_cast<size_t>(blockIdx.x) * nin;
  for (int i = threadIdx.x; i < nin; i += 256) part += static_cast<double>(xr[i]) * w[i];
  part = block_sum(part);
  if (threadIdx.x == 0) y[blockIdx.x] = static_cast<float>(part) + bias[0];
}
__global__ void __launch_bounds__(256) fc_bwd_kernel(const float* x, const float* w, const float* dy, float* dx, float* dw,
                                                      float* dbias, int rows, int nin) {
  // one thread per input feature: dx[r][i] = dy[r] * w[i]; dw[i] += sum_r dy[r] * x[r][i]
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < nin) {
    float acc = 0.0f;
    for (int r = 0; r < rows; ++r) {
      const float d = dy[r];
      if (dx) dx[static_cast<size_t>(r) * nin + i] = d * w[i];
      acc = fmaf(d, x[static_cast<size_t>(r) * nin + i], acc);
    }
    dw[i] += acc;
  }
  if (i == 0) {
    float s = 0.0f;
    for (int r = 0; r < rows; ++r) s += dy[r];
    dbias[0] += s;
  }
}

// out[e] = in[e * cstride + c]  /  in-place strided scatter-add (channel extraction of the D input gradient)
__global__ void __launch_bounds__(256) take_channel_kernel(const float* in, float* out, long long npix, int cstride, int c,
                                                            int accumulate) {
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < npix; e += static_cast<long long>(gridDim.x) * 256)
    out[e] = (accumulate ? out[e] : 0.0f) + in[e * cstride + c];
}

// tensorResample (GAN/multipassGAN-8x.py:545-594) for 2-D frames: out[b,y,x,:] = sum over the 4 cells around pos - 0.5 of
// (1 - |dy|)(1 - |dx|) * value[b, iy, ix, :]; pos[..., 0] runs along H, pos[..., 1] along W; indices are not clamped and
// out-of-range cells contribute nothing (tf.gather_nd on the GPU). BWD = 1: dvalue[b, iy, ix, :] += w * dout[b, y, x, :].
template <int BWD>
__global__ void __launch_bounds__(256) resample_kernel(const float* value, const float* pos, float* out, float* dvalue,
                                                        const float* dout, int n, int hh, int ww, int c) {
  const long long total = static_cast<long long>(n) * hh * ww;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const long long b = e / (static_cast<long long>(hh) * ww);
    const float q0 = pos[2 * e] - 0.5f, q1 = pos[2 * e + 1] - 0.5f;
    const float f0 = floorf(q0), f1 = floorf(q1);
    const int i0 = static_cast<int>(f0), i1 = static_cast<int>(f1);
    for (int ch = 0; ch < c; ++ch) {
      float acc = 0.0f;
      const float g = BWD ? dout[e * c + ch] : 0.0f;
#pragma unroll
      for (int c0 = 0; c0 < 2; ++c0)
#pragma unroll
        for (int c1 = 0; c1 < 2; ++c1) {
          const int y = i0 + c0, x = i1 + c1;
          if (y < 0 || y >= hh || x < 0 || x >= ww) continue;
          const float w = (1.0f - fabsf(q0 - static_cast<float>(y))) * (1.0f - fabsf(q1 - static_cast<float>(x)));
          const long long src = ((b * hh + y) * ww + x) * c + ch;
          if (BWD) atomicAdd(&dvalue[src], w * g);
          else acc = fmaf(w, value[src], acc);
        }
      if (!BWD) out[e * c + ch] = acc;
    }
  }
}

// getSemiLagrPosBatch (tools_wscale/tilecreator_t.py:1341-1378) for 2-D tiles, one thread per high-res cell: the MAC velocity
// (vx, vy at channel c0, c0 + 1 of the low-res tile rows) is interpolated to the S x S grid (order-1 map_coordinates at index
// (i + 0.5) * L / S, mode 'nearest'), centred (mean with the +1 neighbour along its own axis, last cell repeated), scaled by
// S / L, and pos = cell centre - centred velocity * dt with dt = dt0 * (n_t / 2 - frame), frame = row % n_t.
__global__ void __launch_bounds__(256) semilagr_pos_kernel(const float* x, float* pos, int n, int L, int S, int cstride, int c0,
                                                            float dt0, int n_t) {
  const long long total = static_cast<long long>(n) * S * S;
  const float f = static_cast<float>(L) / static_cast<float>(S), scale = static_cast<float>(S) / static_cast<float>(L);
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const int X = static_cast<int>(e % S), Y = static_cast<int>((e / S) % S);
    const long long b = e / (static_cast<long long>(S) * S);
    const float dt = dt0 * static_cast<float>(n_t / 2 - static_cast<int>(b % n_t));
    const float* tile = x + b * L * L * cstride;
    auto up = [&](int ch, int yy, int xx) {  // interpolated component at high-res cell (yy, xx)
      if (S == L) return tile[(yy * L + xx) * cstride + ch];
      float cy = (static_cast<float>(yy) + 0.5f) * f, cx = (static_cast<float>(xx) + 0.5f) * f;
      cy = fminf(fmaxf(cy, 0.0f), static_cast<float>(L - 1));
      cx = fminf(fmaxf(cx, 0.0f), static_cast<float>(L - 1));
      const int y0 = min(static_cast<int>(floorf(cy)), L - 1), x0 = min(static_cast<int>(floorf(cx)), L - 1);
      const int y1 = min(y0 + 1, L - 1), x1 = min(x0 + 1, L - 1);
      const float ty = cy - static_cast<float>(y0), tx = cx - static_cast<float>(x0);
      const float a = tile[(y0 * L + x0) * cstride + ch] * (1.0f - ty) + tile[(y1 * L + x0) * cstride + ch] * ty;
      const float c = tile[(y0 * L + x1) * cstride + ch] * (1.0f - ty) + tile[(y1 * L + x1) * cstride + ch] * ty;
      return a * (1.0f - tx) + c * tx;
    };
    const float vy = 0.5f * (up(c0 + 1, Y, X) + up(c0 + 1, min(Y + 1, S - 1), X)) * scale;
    const float vx = 0.5f * (up(c0, Y, X) + up(c0, Y, min(X + 1, S - 1))) * scale;
    pos[2 * e] = static_cast<float>(Y) + 0.5f - vy * dt;
    pos[2 * e + 1] = static_cast<float>(X) + 0.5f - vx * dt;
  }
}

inline int grid_for(long long total, int sm) {
  long long b = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm) * 8;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace
}  // namespace mpg

using namespace mpg;

// Launch conv_tiled_kernel; grids that leave most SMs idle are split over K (blockIdx.z) into a zeroed output.
static int launch_conv_tiled(mpg_handle h, ConvArgs a, cudaStream_t st) {
  const unsigned gx = static_cast<unsigned>(a.n * ((a.oh + 7) / 8) * ((a.ow + 15) / 16));
  const unsigned gy = static_cast<unsigned>((a.cout + kTC - 1) / kTC);
  const int total_its = a.k * a.k * ((a.cin + kKC - 1) / kKC);
  int nsplit = 1;
  const long long blocks = static_cast<long long>(gx) * gy;
  if (blocks < 2LL * h->sm_count && total_its >= 8) {
    nsplit = static_cast<int>((3LL * h->sm_count + blocks - 1) / blocks);
    if (nsplit > total_its / 4) nsplit = total_its / 4;
    if (nsplit < 1) nsplit = 1;
  }
  a.its_per_split = (total_its + nsplit - 1) / nsplit;
  nsplit = (total_its + a.its_per_split - 1) / a.its_per_split;
  if (nsplit > 1 && !a.accumulate)
    MPG_CUDA(cudaMemsetAsync(a.out, 0, sizeof(float) * static_cast<size_t>(a.n) * a.oh * a.ow * a.cout, st));
  conv_tiled_kernel<<<dim3(gx, gy, static_cast<unsigned>(nsplit)), 256, 0, st>>>(a);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

extern "C" {

int mpg_train_conv_fwd(mpg_handle h, const float* x, const float* w, const float* bias, float* y, int n, int hh, int ww,
                       int cin, int cout, int k, int stride, int in_up, void* stream) {
  MPG_CHECK_ARG(h && x && w && y && n > 0 && hh > 0 && ww > 0 && cin > 0 && cout > 0 && k > 0 && stride > 0 && in_up > 0,
                "mpg_train_conv_fwd: bad argument");
  ConvArgs a;
  a.in = x;
  a.w = w;
  a.bias = bias;
  a.out = y;
  a.n = n;
  a.h = hh;
  a.w_ = ww;
  a.cin = cin;
  a.cout = cout;
  a.k = k;
  a.stride = stride;
  a.pad = same_pad_before_tf(hh, k, stride);
  a.oh = (hh + stride - 1) / stride;
  a.ow = (ww + stride - 1) / stride;
  a.in_up = in_up;
  a.wmode = 0;
  a.accumulate = 0;
  return launch_conv_tiled(h, a, static_cast<cudaStream_t>(stream));
}

/* dx[n,hh,ww,cin] (+)= dgrad of y = conv2d_SAME(x, w, stride) given dy[n,oh,ow,cout] */
int mpg_train_conv_dgrad(mpg_handle h, const float* dy, const float* w, float* dx, int n, int hh, int ww, int cin,
                         int cout, int k, int stride, int accumulate, void* stream) {
  MPG_CHECK_ARG(h && dy && w && dx && n > 0 && cin > 0 && cout > 0 && k > 0 && stride > 0, "mpg_train_conv_dgrad: bad argument");
  const int pad = same_pad_before_tf(hh, k, stride);
  if (stride == 1) {
    ConvArgs a;
    a.in = dy;
    a.w = w;
    a.bias = nullptr;
    a.out = dx;
    a.n = n;
    a.h = hh;
    a.w_ = ww;
    a.cin = cout;  // roles swap
    a.cout = cin;
    a.k = k;
    a.stride = 1;
    a.pad = k - 1 - pad;
    a.oh = hh;
    a.ow = ww;
    a.in_up = 1;
    a.wmode = 1;
    a.accumulate = accumulate;
    return launch_conv_tiled(h, a, static_cast<cudaStream_t>(stream));
  }
  ConvArgs a;  // strided dgrad on the tiled kernel (wmode 2)
  a.in = dy;
  a.w = w;
  a.bias = nullptr;
  a.out = dx;
  a.n = n;
  a.h = (hh + stride - 1) / stride;  // dy spatial size
  a.w_ = (ww + stride - 1) / stride;
  a.cin = cout;  // roles swap
  a.cout = cin;
  a.k = k;
  a.stride = stride;
  a.pad = pad;
  a.oh = hh;
  a.ow = ww;
  a.in_up = 1;
  a.wmode = 2;
  a.accumulate = accumulate;
  return launch_conv_tiled(h, a, static_cast<cudaStream_t>(stream));
}

/* dw[k,k,cin,cout] += wgrad ; dbias[cout] += sum dy (dbias may be NULL); `scratch`: >= cout doubles */
int mpg_train_conv_wgrad(mpg_handle h, const float* x, const float* dy, float* dw, float* dbias, double* scratch, int n,
                         int hh, int ww, int cin, int cout, int k, int stride, int in_up, void* stream) {
  MPG_CHECK_ARG(h && x && dy && dw && n > 0 && cin > 0 && cout > 0 && k > 0 && stride > 0 && in_up > 0,
                "mpg_train_conv_wgrad: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WgradArgs a;
  a.in = x;
  a.dy = dy;
  a.dw = dw;
  a.n = n;
  a.h = hh;
  a.w_ = ww;
  a.cin = cin;
  a.cout = cout;
  a.k = k;
  a.stride = stride;
  a.pad = same_pad_before_tf(hh, k, stride);
  a.oh = (hh + stride - 1) / stride;
  a.ow = (ww + stride - 1) / stride;
  a.in_up = in_up;
  const long long npix = static_cast<long long>(n) * a.oh * a.ow;
  MPG_CHECK_ARG(npix < (1ll << 31), "mpg_train_conv_wgrad: %lld output pixels exceed 2^31", npix);
  if (stride == 1 && (k == 1 || k == 3 || k == 5) && cout <= 32) {
    // thin layers: (tap, ci)-per-thread kernel, persistent over 8x16-pixel tiles
    int ci_chunk = k == 5 ? 8 : (k == 3 ? 24 : 32);
    if (ci_chunk > cin) ci_chunk = cin;
    const int ntiles = n * ((a.oh + kWtTH - 1) / kWtTH) * ((a.ow + kWtTW - 1) / kWtTW);
    int grid = h->sm_count * 2;
    if (grid > ntiles) grid = ntiles;
    if (cout <= 4) wgrad_thin_kernel<4><<<grid, 256, 0, st>>>(a, ci_chunk);
    else if (cout <= 8) wgrad_thin_kernel<8><<<grid, 256, 0, st>>>(a, ci_chunk);
    else wgrad_thin_kernel<32><<<grid, 256, 0, st>>>(a, ci_chunk);
    MPG_CUDA(cudaGetLastError());
  } else {
    if (cin <= 2) launch_wgrad<1, 16>(h, a, npix, st);
    else if (cin <= 4) launch_wgrad<1, 8>(h, a, npix, st);
    else if (cin <= 8) launch_wgrad<1, 4>(h, a, npix, st);
    else if (cin <= 16) launch_wgrad<1, 2>(h, a, npix, st);
    else if (cin >= 128) launch_wgrad<4, 1>(h, a, npix, st);
    else if (cin >= 64) launch_wgrad<2, 1>(h, a, npix, st);
    else launch_wgrad<1, 1>(h, a, npix, st);
    MPG_CUDA(cudaGetLastError());
  }
  if (dbias) {
    MPG_CHECK_ARG(scratch != nullptr, "mpg_train_conv_wgrad: scratch missing");
    MPG_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * cout, st));
    colstats_kernel<<<h->sm_count * 2, 256, 0, st>>>(dy, nullptr, nullptr, nullptr, scratch, npix, cout, 2);
    dsum_to_f32_kernel<<<(cout + 255) / 256, 256, 0, st>>>(scratch, dbias, cout, 1);
    MPG_CUDA(cudaGetLastError());
  }
  return MPG_OK;
}

/* dbias[cout] += sum over rows of dy[rows, cout] (the bias half of mpg_train_conv_wgrad, for callers that compute the
 * filter gradient with mpg_train_conv_wgrad_tc); `scratch`: >= cout doubles */
int mpg_train_bias_grad(mpg_handle h, const float* dy, float* dbias, double* scratch, long long rows, int cout, void* stream) {
  MPG_CHECK_ARG(h && dy && dbias && scratch && rows > 0 && cout > 0, "mpg_train_bias_grad: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MPG_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * cout, st));
  colstats_kernel<<<h->sm_count * 2, 256, 0, st>>>(dy, nullptr, nullptr, nullptr, scratch, rows, cout, 2);
  dsum_to_f32_kernel<<<(cout + 255) / 256, 256, 0, st>>>(scratch, dbias, cout, 1);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

/* training-mode batch norm (tf.contrib.layers.batch_norm, is_training=True, tools_wscale/GAN.py:110):
 * batch mean / biased variance over all rows, y = act(gamma * (x - mean) * rsqrt(var + eps) + beta);
 * moving statistics updated in place with `decay` when moving_mean != NULL. scratch: 2*c doubles. */
int mpg_train_bn_fwd(mpg_handle h, const float* x, const float* gamma, const float* beta, float* y, float* mean,
                     float* var, float* invstd, float* moving_mean, float* moving_var, double* scratch, long long rows,
                     int c, float eps, float decay, int act, void* stream) {
  MPG_CHECK_ARG(h && x && gamma && beta && y && mean && var && invstd && scratch && rows > 0 && c > 0, "mpg_train_bn_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MPG_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * c, st));
  colstats_kernel<<<h->sm_count * 2, 256, 0, st>>>(x, nullptr, nullptr, nullptr, scratch, rows, c, 0);
  if ((c & 3) == 0 && c <= 1024 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
      rows * c < (1LL << 33)) {
    bn_finalize_apply_kernel<<<grid_for(rows * c / 4, h->sm_count), 256, 0, st>>>(x, gamma, beta, scratch, mean, var, invstd,
                                                                                   moving_mean, moving_var, y, rows, c, eps, decay,
                                                                                   act);
  } else {
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(scratch, mean, var, invstd, moving_mean, moving_var, rows, c, eps, decay);
    bn_apply_kernel<<<grid_for(rows * c, h->sm_count), 256, 0, st>>>(x, gamma, beta, mean, invstd, y, rows * c, c, act);
  }
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

/* backward of mpg_train_bn_fwd: dy is the gradient w.r.t. the activated output y; dgamma/dbeta accumulate.
 * dz: unused since round 2 (dy * act'(y) is formed inside the statistics and the dx pass instead of a pass of its own);
 * may be NULL. scratch: 2*c doubles. */
int mpg_train_bn_bwd(mpg_handle h, const float* x, const float* y, const float* dy, const float* gamma, const float* mean,
                     const float* invstd, float* dz, float* dx, float* dgamma, float* dbeta, double* scratch, long long rows,
                     int c, int act, void* stream) {
  MPG_CHECK_ARG(h && x && y && dy && gamma && mean && invstd && dx && dgamma && dbeta && scratch, "mpg_train_bn_bwd: bad argument");
  (void)dz;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MPG_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * c, st));
  colstats_kernel<<<h->sm_count * 2, 256, 0, st>>>(x, dy, mean, invstd, scratch, rows, c, 1, y, act);
  bn_bwd_kernel<<<grid_for(rows * c, h->sm_count), 256, 0, st>>>(x, dy, gamma, mean, invstd, scratch, dx, dgamma, dbeta, rows, c, y,
                                                                  act);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

int mpg_train_act_fwd(mpg_handle h, const float* x, float* y, long long count, int act, void* stream) {
  MPG_CHECK_ARG(h && x && y, "mpg_train_act_fwd: bad argument");
  bias_act_kernel<<<grid_for(count, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, count, act);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
int mpg_train_add_act_fwd(mpg_handle h, const float* a, const float* b, float* y, long long count, int act, void* stream) {
  MPG_CHECK_ARG(h && a && b && y, "mpg_train_add_act_fwd: bad argument");
  add_act_kernel<<<grid_for(count, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, y, count, act);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
/* dz = dy * act'(y) (dz may alias dy) */
int mpg_train_act_bwd(mpg_handle h, const float* y, const float* dy, float* dz, long long count, int act, void* stream) {
  MPG_CHECK_ARG(h && y && dy && dz, "mpg_train_act_bwd: bad argument");
  act_bwd_kernel<<<grid_for(count, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, dy, dz, count, act);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
int mpg_train_axpy(mpg_handle h, float* y, const float* x, float alpha, long long count, void* stream) {
  MPG_CHECK_ARG(h && x && y, "mpg_train_axpy: bad argument");
  axpy_kernel<<<grid_for(count, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, x, alpha, count);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

/* 2x2 average pooling, stride 2, VALID (tools_wscale/GAN.py:162-169 avg_pool defaults); hh, ww: INPUT size (even) */
int mpg_train_avgpool2_fwd(mpg_handle h, const float* x, float* y, int n, int hh, int ww, int c, void* stream) {
  MPG_CHECK_ARG(h && x && y && hh % 2 == 0 && ww % 2 == 0, "mpg_train_avgpool2_fwd: bad argument");
  const long long total = static_cast<long long>(n) * (hh / 2) * (ww / 2) * c;
  avgpool2_fwd_kernel<<<grid_for(total, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, hh / 2, ww / 2, c);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
int mpg_train_avgpool2_bwd(mpg_handle h, const float* dy, float* dx, int n, int hh, int ww, int c, int accumulate,
                           void* stream) {
  MPG_CHECK_ARG(h && dy && dx && hh % 2 == 0 && ww % 2 == 0, "mpg_train_avgpool2_bwd: bad argument");
  const long long total = static_cast<long long>(n) * hh * ww * c;
  avgpool2_bwd_kernel<<<grid_for(total, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, dx, n, hh / 2, ww / 2, c,
                                                                                                  accumulate);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
/* out = a + (b - a) * t with t already clipped to [0,1] (lerp, GAN/multipassGAN-8x.py:596-597) */
int mpg_train_lerp(mpg_handle h, float* out, const float* a, const float* b, float t, long long count, void* stream) {
  MPG_CHECK_ARG(h && out && a && b, "mpg_train_lerp: bad argument");
  lerp_kernel<<<grid_for(count, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, a, b, t, count);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
/* y = alpha * x */
int mpg_train_scale(mpg_handle h, float* y, const float* x, float alpha, long long count, void* stream) {
  MPG_CHECK_ARG(h && x && y, "mpg_train_scale: bad argument");
  scale_kernel<<<grid_for(count, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, x, alpha, count);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
int mpg_train_pick(mpg_handle h, const float* in, float* out, int n, int oh, int ow, int c, int stride, int in_cstride,
                   int out_cstride, int out_c0, void* stream) {
  MPG_CHECK_ARG(h && in && out && stride >= 1 && c <= in_cstride && out_c0 + c <= out_cstride, "mpg_train_pick: bad argument");
  const long long total = static_cast<long long>(n) * oh * ow * c;
  pick_kernel<<<grid_for(total, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, n, oh, ow, c, stride, in_cstride,
                                                                                          out_cstride, out_c0);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
int mpg_train_stuff16(mpg_handle h, const float* dy, void* out16, int out_dtype, int n, int oh, int ow, int c, int stride,
                      int out_cstride, void* stream) {
  MPG_CHECK_ARG(h && dy && out16 && stride >= 1 && c <= out_cstride && mpg::is_h16(out_dtype), "mpg_train_stuff16: bad argument");
  const long long total = static_cast<long long>(n) * oh * stride * ow * stride * out_cstride;
  stuff16_kernel<<<grid_for(total, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, static_cast<uint16_t*>(out16),
                                                                                             out_dtype, n, oh, ow, c, stride, out_cstride);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
/* pixel_norm of the growing generator in training mode (tools_wscale/GAN.py:472-474) and its backward; x [rows, c] */
int mpg_train_pixel_norm_fwd(mpg_handle h, const float* x, float* y, long long rows, int c, void* stream) {
  MPG_CHECK_ARG(h && x && y && rows > 0 && c > 0, "mpg_train_pixel_norm_fwd: bad argument");
  pixel_norm_fwd_kernel<<<grid_for(rows, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, rows, c);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
int mpg_train_pixel_norm_bwd(mpg_handle h, const float* x, const float* dy, float* dx, long long rows, int c, void* stream) {
  MPG_CHECK_ARG(h && x && dy && dx && rows > 0 && c > 0, "mpg_train_pixel_norm_bwd: bad argument");
  pixel_norm_bwd_kernel<<<grid_for(rows, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, dx, rows, c);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
/* WGAN-GP gradient penalty (GAN/multipassGAN-8x.py:1130-1133): g [rows, n] = gradient of mean(critic) w.r.t. the
 * interpolated samples; *loss += mean_b lambda (||g_b + 1e-4|| - target)^2, v [rows, n] = d penalty / d g, norms [rows]
 * (may be NULL) = the per-sample norms */
int mpg_train_gp_penalty(mpg_handle h, const float* g, float* v, double* loss, float* norms, int rows, long long n,
                         float lambda, float target, void* stream) {
  MPG_CHECK_ARG(h && g && v && loss && rows > 0 && n > 0, "mpg_train_gp_penalty: bad argument");
  gp_penalty_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, v, loss, norms, rows, n, lambda, target);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
/* *loss += scale * mean(x^power), power 1 or 2; dx (may be NULL) (+)= the gradient (WGAN critic terms :1111-1112, 1140) */
int mpg_train_mean_pow(mpg_handle h, const float* x, float scale, int power, double* loss, float* dx, long long count,
                       int accumulate, void* stream) {
  MPG_CHECK_ARG(h && x && loss && (power == 1 || power == 2) && count > 0, "mpg_train_mean_pow: bad argument");
  mean_pow_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, scale, power, loss, dx, count, accumulate);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

/* out = a * b elementwise (run-time weight scaling W_eff = v * wscale, tools_wscale/GAN.py:664-668, and its gradient) */
int mpg_train_mul(mpg_handle h, float* out, const float* a, const float* b, long long count, void* stream) {
  MPG_CHECK_ARG(h && out && a && b, "mpg_train_mul: bad argument");
  mul_kernel<<<grid_for(count, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, a, b, count);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

/* losses: *loss (device double) += value; d* (+)= gradient */
int mpg_train_bce_logits(mpg_handle h, const float* logits, float label, float scale, double* loss, float* dlogits,
                         long long count, int accumulate, void* stream) {
  MPG_CHECK_ARG(h && logits && count > 0, "mpg_train_bce_logits: bad argument");
  bce_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, label, scale, loss, dlogits, count, accumulate);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
int mpg_train_l1_mean(mpg_handle h, const float* y, const float* g, float scale, double* loss, float* dg, long long count,
                      int accumulate, void* stream) {
  MPG_CHECK_ARG(h && y && g && count > 0, "mpg_train_l1_mean: bad argument");
  l1_kernel<<<grid_for(count, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, g, scale, loss, dg, count, accumulate);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
int mpg_train_l2_half(mpg_handle h, const float* a, const float* b, float scale, double* loss, float* db, long long count,
                      int accumulate, void* stream) {
  MPG_CHECK_ARG(h && a && b && count > 0, "mpg_train_l2_half: bad argument");
  l2half_kernel<<<grid_for(count, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, scale, loss, db, count, accumulate);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

int mpg_train_adam(mpg_handle h, float* param, const float* grad, float* m, float* v, long long count, float lr_t,
                   float beta1, float beta2, float eps, void* stream) {
  MPG_CHECK_ARG(h && param && grad && m && v && count > 0, "mpg_train_adam: bad argument");
  adam_kernel<<<grid_for(count, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, m, v, count, lr_t, beta1, beta2, eps);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

/* mpg_train_adam with lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) read from a device scalar (graph replays) */
int mpg_train_adam_dev(mpg_handle h, float* param, const float* grad, float* m, float* v, long long count,
                       const float* lr_t_dev, float beta1, float beta2, float eps, void* stream) {
  MPG_CHECK_ARG(h && param && grad && m && v && lr_t_dev && count > 0, "mpg_train_adam_dev: bad argument");
  adam_dev_kernel<<<grid_for(count, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, m, v, count, lr_t_dev, beta1, beta2, eps);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

int mpg_train_fc_fwd(mpg_handle h, const float* x, const float* w, const float* bias, float* y, int rows, int nin, void* stream) {
  MPG_CHECK_ARG(h && x && w && bias && y && rows > 0 && nin > 0, "mpg_train_fc_fwd: bad argument");
  fc_fwd_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, y, nin);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
int mpg_train_fc_bwd(mpg_handle h, const float* x, const float* w, const float* dy, float* dx, float* dw, float* dbias,
                     int rows, int nin, void* stream) {
  MPG_CHECK_ARG(h && x && w && dy && dw && dbias && rows > 0 && nin > 0, "mpg_train_fc_bwd: bad argument");
  fc_bwd_kernel<<<(nin + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, dy, dx, dw, dbias, rows, nin);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

int mpg_train_take_channel(mpg_handle h, const float* in, float* out, long long npix, int cstride, int c, int accumulate,
                           void* stream) {
  MPG_CHECK_ARG(h && in && out && npix > 0 && c >= 0 && c < cstride, "mpg_train_take_channel: bad argument");
  take_channel_kernel<<<grid_for(npix, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, npix, cstride, c, accumulate);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

/* getSemiLagrPosBatch (tools_wscale/tilecreator_t.py:1341-1378, 2-D): x [n, L, L, cstride] low-res tile rows ordered
 * (sample, frame) with (vx, vy) at channels c0, c0 + 1; pos [n, S, S, 2] (y, x) = the positions tensorResample reads,
 * dt of row r = dt0 * (n_t / 2 - r % n_t) */
int mpg_train_semilagr_pos(mpg_handle h, const float* x, float* pos, int n, int L, int S, int cstride, int c0, float dt0,
                           int n_t, void* stream) {
  MPG_CHECK_ARG(h && x && pos && n > 0 && L > 0 && S >= L && cstride >= c0 + 2 && c0 >= 0 && n_t >= 1,
                "mpg_train_semilagr_pos: bad argument");
  const long long total = static_cast<long long>(n) * S * S;
  semilagr_pos_kernel<<<grid_for(total, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, pos, n, L, S, cstride, c0,
                                                                                                dt0, n_t);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
/* tensorResample (GAN/multipassGAN-8x.py:545-594, 2-D): out[n,hh,ww,c] = value re-sampled at pos[n,hh,ww,2] (bilinear around
 * pos - 0.5, no clamping, out-of-range cells contribute 0) */
int mpg_train_resample_fwd(mpg_handle h, const float* value, const float* pos, float* out, int n, int hh, int ww, int c,
                           void* stream) {
  MPG_CHECK_ARG(h && value && pos && out && n > 0 && hh > 0 && ww > 0 && c > 0, "mpg_train_resample_fwd: bad argument");
  const long long total = static_cast<long long>(n) * hh * ww;
  resample_kernel<0><<<grid_for(total, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(value, pos, out, nullptr, nullptr,
                                                                                              n, hh, ww, c);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}
/* dvalue[n,hh,ww,c] += gradient of mpg_train_resample_fwd w.r.t. value given dout (scatter-add; zero dvalue first) */
int mpg_train_resample_bwd(mpg_handle h, const float* dout, const float* pos, float* dvalue, int n, int hh, int ww, int c,
                           void* stream) {
  MPG_CHECK_ARG(h && dout && pos && dvalue && n > 0 && hh > 0 && ww > 0 && c > 0, "mpg_train_resample_bwd: bad argument");
  const long long total = static_cast<long long>(n) * hh * ww;
  resample_kernel<1><<<grid_for(total, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(nullptr, pos, nullptr, dvalue, dout,
                                                                                              n, hh, ww, c);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

}  // extern "C"
