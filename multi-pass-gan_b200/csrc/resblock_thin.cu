// Fused THIN residual block (the reference's resBlock, GAN/multipassGAN-4x.py:505-526, for the channel-poor ends of
// gen_resnet :560 ru1 = 4->8->32 and :564 ru4 = 8->2->1):
//
//     y = act( convB_kxk(act(convA_kxk(x))) + conv_1x1(x) )           (scale/shift of inference BN folded)
//
// in ONE launch: x is read once (optionally through the nearest xf view of max_depool :553-554 and straight from the
// fp32 rows of the slice assembler), the cmid-channel intermediate never leaves shared memory, y is written once.
// Unfused, these layers were 3-4 launches (pack + convA + convB/shortcut) that each round-tripped full-resolution
// activations through HBM and were bound by per-tile role hand-overs of the tcgen05 pipeline (0.25 ms for 3 % of
// the FLOPs of the 4x generator).
//
// Why register-level mma.sync (HMMA m16n8k16) and not tcgen05 here: with 4-8 input channels a tcgen05 SS-mode MMA
// streams a 128x16 A block from shared memory (>= 64 cycles) to do 8-32 columns of work; these blocks are 1 % of the
// generator's FLOPs, so the legacy tensor path (measured 1024 MAC/clk/SM on B200, tools/micro/hmma_bench.cu) is
// already far above what the HBM traffic needs, and it has no TMEM / mbarrier hand-overs between the two chained convs.
//
// Tiling: persistent CTAs (2 per SM), tile = 32x32 output pixels. Per tile
//   stage 0  input window 40x40 -> smem as 16-bit [y][x][CPP] (CPP = 4 or 8 channels per pixel), zero outside the image
//            (= SAME padding, tools_wscale/GAN.py:691)
//   stage 1  convA on the 36x36 window the second conv needs: implicit GEMM M = pixels, K = taps*CPP, N = 8;
//            + shift, act, ZERO outside the image (SAME padding of convB's input), 16-bit -> smem [y][x][8]
//   stage 2  convB (K = 25 taps * 8) + 1x1 shortcut (K = CPP from the centre of the input window), N = 8*NT2;
//            + shift, act, store. GEMM columns are permuted (col (nt, c) = channel 8*(c/2) + 2*nt + c%2) so a lane
//            owns 8 contiguous channels of its pixel: one 16-byte store per pixel row, 512 contiguous bytes per warp.
// A fragments come from ldmatrix (16-byte pixels) or 32-bit loads (8-byte pixels); B fragments (weights) are packed on
// the host in fragment order (one conflict-free LDS.64 per lane).
#include <string.h>

#include <vector>

#include "common.h"

struct mpg_resblock_plan_s {
  mpg_handle h;
  mpg_resblock_desc d;
  void* d_blob;
  size_t blob_bytes;
  int cpp, nt2;
  size_t smem_bytes;
  int grid;
  double flops;
};

namespace mpg {
namespace {

constexpr int kTW = 32, kTH = 32;         // output tile
constexpr int kKS = 5, kTaps = 25;        // filter size of both convs
constexpr int kIW = kTW + 2 * (kKS - 1);  // 40: input window
constexpr int kIH = kTH + 2 * (kKS - 1);
constexpr int kMW = kTW + (kKS - 1);      // 36: window of the intermediate
constexpr int kMH = kTH + (kKS - 1);
constexpr int kThreads = 256;
#ifndef MPG_RB_MT2
#define MPG_RB_MT2 4
#endif
#ifndef MPG_RB_OCC
#define MPG_RB_OCC 2
#endif
constexpr int kRbMT2 = MPG_RB_MT2;  // 4: two image rows per pass (64 accumulator registers), 2: one row (32)
constexpr int kRbOcc = MPG_RB_OCC;  // CTAs per SM the register budget is sized for

struct RbParams {
  const void* x;
  void* y;
  const uint8_t* blob;
  int n, h, w;
  int in_up, in_cstride, out_cstride;
  int act;
  int tiles_x, tiles_y, num_tiles;
  unsigned long long* sat;  // validation mode: counts intermediate values beyond the 16-bit range (null = off)
};

__host__ __device__ constexpr int tap_off(int tap, int pitch) {
  return (tap >= kTaps ? kTaps - 1 : tap) / kKS * pitch + (tap >= kTaps ? kTaps - 1 : tap) % kKS;
}

template <bool BF16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (BF16)
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ float act_apply(float x, float ca, float cb) { return fmaf(cb, fabsf(x), ca * x); }

// CPP: 16-bit channels per pixel of the staged input (4: <= 4 input channels, fp32 source; 8: <= 8, 16-bit source)
// NT2: 8-column MMA tiles of the block output (4: 32 channels, 16-bit output; 1: <= 8 channels)
template <int CPP, int NT2, bool BF16, bool IN_F32, bool OUT_F32>
__global__ void __launch_bounds__(kThreads, kRbOcc) resblock_thin_kernel(const RbParams p) {
  constexpr int MT2 = kRbMT2;  // m-tiles (16 pixels each) a warp accumulates at once in stage 2
  constexpr int KS1 = (kTaps * CPP + 15) / 16;  // K steps of convA: 7 (4 taps per step) or 13 (2 taps per step)
  constexpr int KS2 = (kTaps + 1) / 2;          // 13: two taps x 8 channels per step
  constexpr int IN_PX = CPP * 2;                // bytes per staged input pixel
  constexpr uint32_t WA_BYTES = KS1 * 256, WB_BYTES = KS2 * NT2 * 256, WS_BYTES = NT2 * 256;
  constexpr uint32_t BLOB_BYTES = WA_BYTES + WB_BYTES + WS_BYTES + 32 + NT2 * 32;

  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* s_in = smem;                              // [kIH][kIW][CPP] 16-bit
  uint8_t* s_mid = s_in + kIH * kIW * IN_PX;         // [kMH][kMW][8] 16-bit
  uint8_t* s_w = s_mid + kMH * kMW * 16;             // weight fragments + shifts
  const uint32_t a_in = smem_addr(s_in), a_mid = smem_addr(s_mid), a_w = smem_addr(s_w);
  const float* s_shiftA = reinterpret_cast<const float*>(s_w + WA_BYTES + WB_BYTES + WS_BYTES);
  const float* s_shiftB = s_shiftA + 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int dt = BF16 ? MPG_BF16 : MPG_F16;

  for (uint32_t i = tid; i < BLOB_BYTES / 16; i += kThreads)
    reinterpret_cast<uint4*>(s_w)[i] = __ldg(reinterpret_cast<const uint4*>(p.blob) + i);

  const float ca = p.act == MPG_ACT_RELU ? 0.5f : (p.act == MPG_ACT_LRELU ? 0.6f : 1.0f);
  const float cb = p.act == MPG_ACT_RELU ? 0.5f : (p.act == MPG_ACT_LRELU ? 0.4f : 0.0f);
  const int hs = p.h / p.in_up, ws = p.w / p.in_up;

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int per_img = p.tiles_x * p.tiles_y;
    const int n = tile / per_img;
    const int r = tile - n * per_img;
    const int ty0 = (r / p.tiles_x) * kTH, tx0 = (r % p.tiles_x) * kTW;

    // ---------------- stage 0: input window -> smem (16-bit), zero outside the image. All loads of a thread are issued
    //                  before the first conversion (a rolled loop exposed one global-memory latency per window row)
    {
      constexpr int NLD = (kIH * kIW + kThreads - 1) / kThreads;
      uint4 raw[NLD];
#pragma unroll
      for (int j = 0; j < NLD; ++j) {
        const int i = tid + j * kThreads;
        const int wy = i / kIW, wx = i - wy * kIW;
        const int gy = ty0 - (kKS - 1) + wy, gx = tx0 - (kKS - 1) + wx;
        const bool in = i < kIH * kIW && gy >= 0 && gy < p.h && gx >= 0 && gx < p.w;
        raw[j] = make_uint4(0u, 0u, 0u, 0u);
        if (in) {
          const size_t spix = static_cast<size_t>(n * hs + gy / p.in_up) * ws + gx / p.in_up;
          raw[j] = IN_F32 ? __ldg(reinterpret_cast<const uint4*>(static_cast<const float*>(p.x) + spix * p.in_cstride))
                          : __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.x) + spix * p.in_cstride));
        }
      }
#pragma unroll
      for (int j = 0; j < NLD; ++j) {
        const int i = tid + j * kThreads;
        if (i >= kIH * kIW) break;
        if (IN_F32) {
          uint2 q;
          q.x = pack_h16x2(__uint_as_float(raw[j].x), __uint_as_float(raw[j].y), dt);
          q.y = pack_h16x2(__uint_as_float(raw[j].z), __uint_as_float(raw[j].w), dt);
          *reinterpret_cast<uint2*>(s_in + i * IN_PX) = q;
        } else {
          *reinterpret_cast<uint4*>(s_in + i * IN_PX) = raw[j];
        }
      }
    }
    __syncthreads();

    // ---------------- stage 1: convA on the kMH x kMW window (linear pixel q = y*kMW + x), N = 8
    {
      uint2 wA[KS1];
#pragma unroll
      for (int ks = 0; ks < KS1; ++ks) wA[ks] = lds64(a_w + (ks * 32 + lane) * 8);
      const float shA0 = s_shiftA[2 * t], shA1 = s_shiftA[2 * t + 1];
      // two m-tiles (independent accumulator chains) per iteration: a single chain of dependent HMMAs is latency bound
      constexpr int NMT = (kMH * kMW) / 16;
      for (int mt0 = warp * 2; mt0 < NMT; mt0 += (kThreads / 32) * 2) {
        float acc[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f;
        const int mtu[2] = {mt0, mt0 + 1 < NMT ? mt0 + 1 : mt0};  // an odd tail recomputes the same tile (stored once)
        if (CPP == 8) {
          // ldmatrix: lane l supplies the row address of matrix l/8 = (pixel half l/8 & 1, tap half l/16)
          uint32_t base[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int q = mtu[u] * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            base[u] = a_in + ((q / kMW) * kIW + q % kMW) * IN_PX;
          }
          const bool hi = lane >= 16;
#pragma unroll
          for (int ks = 0; ks < KS1; ++ks) {
            const uint32_t o = (hi ? tap_off(2 * ks + 1, kIW) : tap_off(2 * ks, kIW)) * IN_PX;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              uint32_t a[4];
              ldmatrix_x4(a, base[u] + o);
              mma16816<BF16>(acc[u], a, wA[ks].x, wA[ks].y);
            }
          }
        } else {
          // 8-byte pixels: k = tap*4 + ch; a0/a1 = rows g / g+8 at tap 4ks + t/2, a2/a3 two taps further
          uint32_t b0[2], b1[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int q0 = mtu[u] * 16 + g, q1 = q0 + 8;
            b0[u] = a_in + ((q0 / kMW) * kIW + q0 % kMW) * IN_PX + (t & 1) * 4;
            b1[u] = a_in + ((q1 / kMW) * kIW + q1 % kMW) * IN_PX + (t & 1) * 4;
          }
          const bool odd = (t >> 1) != 0;
#pragma unroll
          for (int ks = 0; ks < KS1; ++ks) {
            const uint32_t o0 = (odd ? tap_off(4 * ks + 1, kIW) : tap_off(4 * ks, kIW)) * IN_PX;
            const uint32_t o1 = (odd ? tap_off(4 * ks + 3, kIW) : tap_off(4 * ks + 2, kIW)) * IN_PX;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              uint32_t a[4];
              a[0] = lds32(b0[u] + o0);
              a[1] = lds32(b1[u] + o0);
              a[2] = lds32(b0[u] + o1);
              a[3] = lds32(b1[u] + o1);
              mma16816<BF16>(acc[u], a, wA[ks].x, wA[ks].y);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u == 1 && mt0 + 1 >= NMT) break;
#pragma unroll
          for (int hrow = 0; hrow < 2; ++hrow) {
            const int q = mtu[u] * 16 + g + hrow * 8;
            const int my = q / kMW, mx = q - my * kMW;
            const int gy = ty0 - (kKS - 1) / 2 + my, gx = tx0 - (kKS - 1) / 2 + mx;
            const bool in = gy >= 0 && gy < p.h && gx >= 0 && gx < p.w;
            const float v0 = in ? act_apply(acc[u][2 * hrow] + shA0, ca, cb) : 0.0f;
            const float v1 = in ? act_apply(acc[u][2 * hrow + 1] + shA1, ca, cb) : 0.0f;
            if (p.sat != nullptr) {  // the intermediate never reaches global memory: count its saturating stores here
              const float lim = BF16 ? 3.3895314e38f : 65504.0f;
              const unsigned c = (fabsf(v0) >= lim || v0 != v0) + (fabsf(v1) >= lim || v1 != v1);
              if (c) atomicAdd(p.sat, static_cast<unsigned long long>(c));
            }
            *reinterpret_cast<uint32_t*>(s_mid + q * 16 + t * 4) = pack_h16x2(v0, v1, dt);
          }
        }
      }
    }
    __syncthreads();

    // ---------------- stage 2: convB + 1x1 shortcut; warp = 4 image rows, two passes of 2 rows x 2 halves (4 m-tiles)
#pragma unroll 1
    for (int pass = 0; pass < 8 / MT2; ++pass) {
      const int y0 = warp * 4 + pass * (MT2 / 2);  // first tile row of this pass (MT2/2 rows x 2 halves)
      float acc[MT2][NT2][4];
#pragma unroll
      for (int m = 0; m < MT2; ++m)
#pragma unroll
        for (int nt = 0; nt < NT2; ++nt) acc[m][nt][0] = acc[m][nt][1] = acc[m][nt][2] = acc[m][nt][3] = 0.f;
      // ldmatrix row address of this lane inside m-tile 0 of the pass: pixel x = (l & 7) + 8 * ((l >> 3) & 1)
      const uint32_t lbase = a_mid + (y0 * kMW + (lane & 7) + ((lane >> 3) & 1) * 8) * 16;
      const bool hi = lane >= 16;
#pragma unroll
      for (int ks = 0; ks < KS2; ++ks) {
        uint2 bf[NT2];
#pragma unroll
        for (int nt = 0; nt < NT2; ++nt) bf[nt] = lds64(a_w + WA_BYTES + ((ks * NT2 + nt) * 32 + lane) * 8);
        const uint32_t toff = (hi ? tap_off(2 * ks + 1, kMW) : tap_off(2 * ks, kMW)) * 16;
#pragma unroll
        for (int m = 0; m < MT2; ++m) {
          uint32_t a[4];
          ldmatrix_x4(a, lbase + toff + ((m >> 1) * kMW + (m & 1) * 16) * 16);
#pragma unroll
          for (int nt = 0; nt < NT2; ++nt) mma16816<BF16>(acc[m][nt], a, bf[nt].x, bf[nt].y);
        }
      }
      {  // shortcut: K = CPP channels of the centre pixel of the input window
        uint2 bf[NT2];
#pragma unroll
        for (int nt = 0; nt < NT2; ++nt) bf[nt] = lds64(a_w + WA_BYTES + WB_BYTES + (nt * 32 + lane) * 8);
#pragma unroll
        for (int m = 0; m < MT2; ++m) {
          const int yy = y0 + (m >> 1) + (kKS - 1), xx = (m & 1) * 16 + (kKS - 1);
          uint32_t a[4] = {0u, 0u, 0u, 0u};
          if (CPP == 8 || t < 2) {
            a[0] = lds32(a_in + (yy * kIW + xx + g) * IN_PX + t * 4);
            a[1] = lds32(a_in + (yy * kIW + xx + g + 8) * IN_PX + t * 4);
          }
#pragma unroll
          for (int nt = 0; nt < NT2; ++nt) mma16816<BF16>(acc[m][nt], a, bf[nt].x, bf[nt].y);
        }
      }
      // epilogue: lane (g, t) holds GEMM columns (nt, 2t + e) of rows g and g+8 = channels 8t + 2nt + e (NT2 == 4)
#pragma unroll
      for (int m = 0; m < MT2; ++m) {
        const int gy = ty0 + y0 + (m >> 1);
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int gx = tx0 + (m & 1) * 16 + g + hrow * 8;
          if (gy >= p.h || gx >= p.w) continue;
          const size_t pix = (static_cast<size_t>(n) * p.h + gy) * p.w + gx;
          if (NT2 == 4 && !OUT_F32) {
            uint32_t wv[4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const float v0 = act_apply(acc[m][nt][2 * hrow] + s_shiftB[nt * 8 + 2 * t], ca, cb);
              const float v1 = act_apply(acc[m][nt][2 * hrow + 1] + s_shiftB[nt * 8 + 2 * t + 1], ca, cb);
              wv[nt] = pack_h16x2(v0, v1, dt);
            }
            *reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.y) + pix * p.out_cstride + t * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          } else {
            // <= 8 output channels (identity column order): lane t owns channels 2t, 2t+1
            const float v0 = act_apply(acc[m][0][2 * hrow] + s_shiftB[2 * t], ca, cb);
            const float v1 = act_apply(acc[m][0][2 * hrow + 1] + s_shiftB[2 * t + 1], ca, cb);
            if (OUT_F32) {
              float* o = static_cast<float*>(p.y) + pix * p.out_cstride;
              if (2 * t < p.out_cstride) o[2 * t] = v0;
              if (2 * t + 1 < p.out_cstride) o[2 * t + 1] = v1;
            } else {
              *reinterpret_cast<uint32_t*>(static_cast<uint16_t*>(p.y) + pix * p.out_cstride + 2 * t) = pack_h16x2(v0, v1, dt);
            }
          }
        }
      }
    }
    __syncthreads();  // everyone is done with s_in / s_mid before the next tile overwrites them
  }
}

typedef void (*RbKernel)(const RbParams);

RbKernel rb_kernel(int cpp, int nt2, bool bf16, bool in_f32, bool out_f32) {
  if (cpp == 4 && nt2 == 4 && in_f32 && !out_f32)
    return bf16 ? resblock_thin_kernel<4, 4, true, true, false> : resblock_thin_kernel<4, 4, false, true, false>;
  if (cpp == 8 && nt2 == 4 && !in_f32 && !out_f32)
    return bf16 ? resblock_thin_kernel<8, 4, true, false, false> : resblock_thin_kernel<8, 4, false, false, false>;
  if (cpp == 8 && nt2 == 1 && !in_f32 && out_f32)
    return bf16 ? resblock_thin_kernel<8, 1, true, false, true> : resblock_thin_kernel<8, 1, false, false, true>;
  if (cpp == 8 && nt2 == 1 && !in_f32 && !out_f32)
    return bf16 ? resblock_thin_kernel<8, 1, true, false, false> : resblock_thin_kernel<8, 1, false, false, false>;
  return nullptr;
}

uint16_t to_h16(float f, bool bf16) {
  if (bf16) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40u);
    u += 0x7fffu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
  }
  __half hv = __float2half_rn(f);
  uint16_t r;
  memcpy(&r, &hv, 2);
  return r;
}

// B fragments of one K x (8*NT) GEMM operand given as a callback W(k, col): [kstep][ntile][lane] x {b0, b1},
// b0 = (k = 2t, 2t+1; n = g), b1 = (k = 2t+8, 2t+9; n = g), low half first.
template <typename F>
void pack_frags(std::vector<uint16_t>& out, int ksteps, int ntiles, bool bf16, F W) {
  for (int ks = 0; ks < ksteps; ++ks)
    for (int nt = 0; nt < ntiles; ++nt)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        const int ks16 = ks * 16, col = nt * 8 + g;
        out.push_back(to_h16(W(ks16 + 2 * t, col), bf16));
        out.push_back(to_h16(W(ks16 + 2 * t + 1, col), bf16));
        out.push_back(to_h16(W(ks16 + 2 * t + 8, col), bf16));
        out.push_back(to_h16(W(ks16 + 2 * t + 9, col), bf16));
      }
}

}  // namespace
}  // namespace mpg

extern "C" {

int mpg_resblock_plan_create(mpg_handle h, const mpg_resblock_desc* dsc, const float* w_a, const float* w_b,
                             const float* w_s, const float* scale_a, const float* scale_b, const float* scale_s,
                             const float* shift_a, const float* shift_bs, mpg_resblock_plan* out) {
  using namespace mpg;
  MPG_CHECK_ARG(h && dsc && w_a && w_b && w_s && out, "mpg_resblock_plan_create: null argument");
  mpg_resblock_desc d = *dsc;
  if (d.in_upsample <= 0) d.in_upsample = 1;
  MPG_CHECK_ARG(d.n > 0 && d.h > 0 && d.w > 0 && d.h % d.in_upsample == 0 && d.w % d.in_upsample == 0,
                "resblock: bad spatial size n=%d h=%d w=%d in_upsample=%d", d.n, d.h, d.w, d.in_upsample);
  MPG_CHECK_ARG(d.act == MPG_ACT_NONE || d.act == MPG_ACT_RELU || d.act == MPG_ACT_LRELU, "resblock: activation %d not supported", d.act);
  if (d.ksize != kKS || d.cmid < 1 || d.cmid > 8 || !is_h16(d.mma_dtype)) {
    set_error("resblock: the fused thin block needs ksize 5, cmid <= 8 and a 16-bit MMA type (got k=%d cmid=%d mma_dtype=%d)",
              d.ksize, d.cmid, d.mma_dtype);
    return MPG_ENOSUP;
  }
  int cpp, nt2;
  const bool in_f32 = d.in_dtype == MPG_F32;
  const bool out_f32 = d.out_dtype == MPG_F32;
  if (in_f32 && d.cin <= 4 && d.in_cstride % 4 == 0 && d.in_cstride >= 4) cpp = 4;
  else if (!in_f32 && d.in_dtype == d.mma_dtype && d.cin <= 8 && d.in_cstride % 8 == 0) cpp = 8;
  else {
    set_error("resblock: input must be fp32 with <= 4 channels (cstride %% 4 == 0) or the 16-bit MMA type with <= 8 channels (cstride %% 8 == 0)");
    return MPG_ENOSUP;
  }
  if (d.cout == 32 && !out_f32 && d.out_dtype == d.mma_dtype && d.out_cstride == 32) nt2 = 4;
  else if (d.cout <= 8 && d.out_cstride >= d.cout && (out_f32 || (d.out_dtype == d.mma_dtype && d.out_cstride == 8))) nt2 = 1;
  else {
    set_error("resblock: output must be 32 channels (16-bit, cstride 32) or <= 8 channels (fp32 any cstride, or 16-bit cstride 8)");
    return MPG_ENOSUP;
  }
  const bool bf16 = d.mma_dtype == MPG_BF16;
  if (rb_kernel(cpp, nt2, bf16, in_f32, out_f32) == nullptr) {
    set_error("resblock: no kernel instance for cpp=%d nt2=%d in_f32=%d out_f32=%d", cpp, nt2, in_f32, out_f32);
    return MPG_ENOSUP;
  }
  MPG_CHECK_ARG(!out_f32 || d.out_cstride <= 8, "resblock: fp32 output cstride %d > 8", d.out_cstride);

  const int ks1 = (kTaps * cpp + 15) / 16, ks2 = (kTaps + 1) / 2;
  std::vector<uint16_t> frags;
  auto sc = [](const float* s, int c) { return s ? s[c] : 1.0f; };
  // convA: k = tap*cpp + ch, col = mid channel
  pack_frags(frags, ks1, 1, bf16, [&](int k, int col) -> float {
    const int tap = k / cpp, ch = k % cpp;
    if (tap >= kTaps || ch >= d.cin || col >= d.cmid) return 0.0f;
    return w_a[(static_cast<size_t>(tap) * d.cin + ch) * d.cmid + col] * sc(scale_a, col);
  });
  // convB: k = tap*8 + ch; column (nt, c) -> channel 8*(c/2) + 2*nt + c%2 when the block has 32 outputs
  auto chan_of = [&](int col) { return nt2 == 4 ? 8 * ((col & 7) >> 1) + 2 * (col >> 3) + (col & 1) : col; };
  pack_frags(frags, ks2, nt2, bf16, [&](int k, int col) -> float {
    const int tap = k / 8, ch = k % 8, co = chan_of(col);
    if (tap >= kTaps || ch >= d.cmid || co >= d.cout) return 0.0f;
    return w_b[(static_cast<size_t>(tap) * d.cmid + ch) * d.cout + co] * sc(scale_b, co);
  });
  // shortcut: k = input channel
  pack_frags(frags, 1, nt2, bf16, [&](int k, int col) -> float {
    const int co = chan_of(col);
    if (k >= d.cin || co >= d.cout) return 0.0f;
    return w_s[static_cast<size_t>(k) * d.cout + co] * sc(scale_s, co);
  });
  std::vector<uint8_t> blob(frags.size() * 2 + 32 + nt2 * 32, 0);
  memcpy(blob.data(), frags.data(), frags.size() * 2);
  float* sh = reinterpret_cast<float*>(blob.data() + frags.size() * 2);
  for (int c = 0; c < d.cmid; ++c) sh[c] = shift_a ? shift_a[c] : 0.0f;
  for (int col = 0; col < nt2 * 8; ++col) {
    const int co = chan_of(col);
    sh[8 + col] = (shift_bs && co < d.cout) ? shift_bs[co] : 0.0f;
  }

  mpg_resblock_plan p = new mpg_resblock_plan_s();
  memset(p, 0, sizeof(*p));
  p->h = h;
  p->d = d;
  p->cpp = cpp;
  p->nt2 = nt2;
  p->blob_bytes = blob.size();
  p->flops = 2.0 * d.n * d.h * d.w * (static_cast<double>(kTaps) * d.cin * d.cmid + static_cast<double>(kTaps) * d.cmid * d.cout +
                                      static_cast<double>(d.cin) * d.cout);
  DeviceGuard guard(h->device);
  cudaError_t e = cudaMalloc(&p->d_blob, blob.size());
  if (e == cudaSuccess) e = cudaMemcpy(p->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("resblock: weight upload failed: %s", cudaGetErrorString(e));
    if (p->d_blob) cudaFree(p->d_blob);
    delete p;
    return static_cast<int>(e);
  }
  p->smem_bytes = static_cast<size_t>(kIH) * kIW * cpp * 2 + static_cast<size_t>(kMH) * kMW * 16 + blob.size();
  const int tiles = ceil_div(d.w, kTW) * ceil_div(d.h, kTH) * d.n;
  p->grid = tiles < kRbOcc * h->sm_count ? tiles : kRbOcc * h->sm_count;
  e = cudaFuncSetAttribute(rb_kernel(cpp, nt2, bf16, in_f32, out_f32), cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(p->smem_bytes));  // per device and per instance; plan creation is not a hot path
  if (e != cudaSuccess) {
    set_error("resblock: cudaFuncSetAttribute(max dynamic smem %zu) failed: %s", p->smem_bytes, cudaGetErrorString(e));
    cudaFree(p->d_blob);
    delete p;
    return static_cast<int>(e);
  }
  *out = p;
  return MPG_OK;
}

int mpg_resblock_plan_run(mpg_resblock_plan p, const void* x, void* y, void* stream) {
  return mpg_resblock_plan_run_checked(p, x, y, nullptr, stream);
}

int mpg_resblock_plan_run_checked(mpg_resblock_plan p, const void* x, void* y, unsigned long long* sat_counter_dev,
                                  void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(p && x && y, "mpg_resblock_plan_run: null argument");
  MPG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "resblock: input not 16-byte aligned");
  const mpg_resblock_desc& d = p->d;
  MPG_CHECK_ARG(d.out_dtype == MPG_F32 || (reinterpret_cast<uintptr_t>(y) & 15) == 0, "resblock: output not 16-byte aligned");
  RbParams q;
  q.x = x;
  q.y = y;
  q.blob = static_cast<const uint8_t*>(p->d_blob);
  q.n = d.n;
  q.h = d.h;
  q.w = d.w;
  q.in_up = d.in_upsample;
  q.in_cstride = d.in_cstride;
  q.out_cstride = d.out_cstride;
  q.act = d.act;
  q.tiles_x = ceil_div(d.w, kTW);
  q.tiles_y = ceil_div(d.h, kTH);
  q.num_tiles = q.tiles_x * q.tiles_y * d.n;
  q.sat = sat_counter_dev;
  DeviceGuard guard(p->h->device);
  rb_kernel(p->cpp, p->nt2, d.mma_dtype == MPG_BF16, d.in_dtype == MPG_F32, d.out_dtype == MPG_F32)
      <<<p->grid, kThreads, p->smem_bytes, static_cast<cudaStream_t>(stream)>>>(q);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("resblock launch failed: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return MPG_OK;
}

int mpg_resblock_plan_destroy(mpg_resblock_plan p) {
  if (!p) return MPG_OK;
  if (p->d_blob) cudaFree(p->d_blob);
  delete p;
  return MPG_OK;
}

double mpg_resblock_plan_flops(mpg_resblock_plan p) { return p ? p->flops : 0.0; }

}  // extern "C"
