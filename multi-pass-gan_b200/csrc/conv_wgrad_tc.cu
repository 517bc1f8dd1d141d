// tcgen05 weight-gradient kernel of the training step (GAN/multipassGAN-4x.py:889-898 runs
// tf.train.AdamOptimizer.minimize, whose Conv2DBackpropFilter this replaces for the wide stride-1 generator convs).
//
//   dW[dy][dx][ci][co] += sum_{n,y,x} X[n][y+dy-p][x+dx-p][ci] * dY[n][y][x][co]
//
// GEMM view per tap: D[M][N] += A[M][K] * B[K][N] with K = pixels. Both operands are NHWC, i.e. the CHANNEL index is
// contiguous and the pixel (K) index strides: "MN-major" operands for tcgen05 (instruction-descriptor transpose bits).
// A TMA box [64 ch][pixels] with 128-byte swizzle lands in shared memory as 128-byte pixel rows, which is exactly the
// MN-major SWIZZLE_128B canonical layout (8 consecutive K = 8 pixel rows = one 1 KB atom, SBO = 1024; the second
// 64-channel block sits LBO bytes away). So no transposition happens anywhere:
//   * M operand = the 128-channel tensor (X when Cin = 128, otherwise dY), two 64-channel blocks
//   * N operand = the other tensor (32 / 64 / 128 channels; 32 channels use the 64-byte swizzle)
//   * the k horizontal taps re-use ONE staged image row of X (width W+k-1, zero filled by TMA = SAME padding): tap dx
//     is the same row addressed from a start shifted by dx pixels (the swizzle is a function of the absolute address)
//   * one fp32 accumulator [128 x N] per tap in TMEM, up to 512/N taps per CTA; a CTA owns (dy, a dx range) and a
//     slice of the N*H image rows; partial sums are added to dW with red.global (split-K over pixels)
// Roofline: tensor pipe in principle (2*pixels*k^2*Cin*Cout FLOP), in practice the tiled-TMA row rate (~5 cycles per
// box row per SM, measured) because every staged pixel row only feeds k MMAs.
#include <string.h>

#include "common.h"
#include "ptx.cuh"

namespace mpg {
namespace {

constexpr int kWgMaxStages = 8;

struct WgTcParams {
  int n, h, w, cin, cout, k, pad;
  int m_is_x;           // 1: M operand = X (Cin = 128), N operand = dY; 0: M = dY (Cout = 128), N = X
  int nN;               // UMMA N = channels of the N operand
  int nblkN, rbN;       // 64-/32-channel blocks of the N operand, bytes per pixel row of a block (128 or 64)
  int groups_per_dy, taps_per_group;
  int rows_per_split;   // image rows (of the n*h row space) per CTA along gridDim.y
  int nstages, stage_bytes;
  int offM1, offN0, offN1;  // block offsets inside a stage (M block 0 at 0)
  int x_row_bytes, dy_row_bytes;  // TMA bytes of one staged X / dY block row
  uint32_t tmem_cols;
  float* dw;
};

__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(256, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy, const WgTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[kWgMaxStages], empty[kWgMaxStages];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x;
  const int dy = g / p.groups_per_dy;
  const int dx0 = (g - dy * p.groups_per_dy) * p.taps_per_group;
  const int nt = (p.k - dx0) < p.taps_per_group ? (p.k - dx0) : p.taps_per_group;
  const int total_rows = p.n * p.h;
  const int r0 = blockIdx.y * p.rows_per_split;
  const int r1 = (r0 + p.rows_per_split) < total_rows ? (r0 + p.rows_per_split) : total_rows;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_dy);
    for (int i = 0; i < p.nstages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== producer: one X row (with halo) and one dY row per image row =====================
    if (lane == 0) {
      const int nblkX = p.m_is_x ? 2 : p.nblkN, nblkD = p.m_is_x ? p.nblkN : 2;
      const int cwX = p.m_is_x ? 64 : (p.rbN >> 1), cwD = p.m_is_x ? (p.rbN >> 1) : 64;
      const int offX[2] = {p.m_is_x ? 0 : p.offN0, p.m_is_x ? p.offM1 : p.offN1};
      const int offD[2] = {p.m_is_x ? p.offN0 : 0, p.m_is_x ? p.offN1 : p.offM1};
      const uint32_t bytes = static_cast<uint32_t>(nblkX * p.x_row_bytes + nblkD * p.dy_row_bytes);
      int st = 0;
      uint32_t ph = 0;
      for (int r = r0; r < r1; ++r) {
        const int img = r / p.h, y = r - img * p.h;
        const int iy = y + dy - p.pad;
        if (iy < 0 || iy >= p.h) continue;  // the whole row of this vertical tap lies in the zero padding
        mbar_wait(&empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&full[st], bytes);
        uint8_t* base = smem + static_cast<size_t>(st) * p.stage_bytes;
        for (int b = 0; b < nblkX; ++b) tma_load_4d(base + offX[b], &tm_x, &full[st], b * cwX, -p.pad, iy, img);
        for (int b = 0; b < nblkD; ++b) tma_load_4d(base + offD[b], &tm_dy, &full[st], b * cwD, 0, y, img);
        if (++st == p.nstages) {
          st = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: bf16 x bf16 -> fp32, BOTH operands MN-major (transpose bits 15 / 16), M = 128
    const uint32_t idesc = umma_idesc_f16kind(128, p.nN, 1u) | (1u << 15) | (1u << 16);
    const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
    const uint32_t stage16 = static_cast<uint32_t>(p.stage_bytes) >> 4;
    // M operand: two SWIZZLE_128B blocks offM1 apart (LBO), 8-pixel groups 1024 B apart (SBO)
    const uint32_t m_lo_extra = (static_cast<uint32_t>(p.offM1) >> 4) << 16;
    const uint32_t m_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    // N operand: blocks of rbN-byte pixel rows
    const uint32_t n_lo_extra = (p.nblkN > 1 ? (static_cast<uint32_t>(p.offN1 - p.offN0) >> 4) : 1u) << 16;
    const uint32_t n_hi = ((8u * static_cast<uint32_t>(p.rbN)) >> 4) | (1u << 14) | ((p.rbN == 128 ? 2u : 4u) << 29);
    const uint32_t rbM16 = 128u >> 4, rbN16 = static_cast<uint32_t>(p.rbN) >> 4;
    const uint32_t offN0_16 = static_cast<uint32_t>(p.offN0) >> 4;
    const int ksteps = p.w >> 4;
    const bool leader = elect_one() != 0;
    int st = 0;
    uint32_t ph = 0;
    uint32_t accumulate = 0;
    for (int r = r0; r < r1; ++r) {
      const int img = r / p.h, y = r - img * p.h;
      const int iy = y + dy - p.pad;
      if (iy < 0 || iy >= p.h) continue;
      mbar_wait(&full[st], ph);
      tc_fence_after();
      const uint32_t base16 = smem_lo + static_cast<uint32_t>(st) * stage16;
      if (leader) {
        for (int t = 0; t < nt; ++t) {
          // the X operand starts dx pixels into the staged row; dY is never shifted
          const uint32_t shift = static_cast<uint32_t>(dx0 + t);
          uint32_t m16 = base16 + (p.m_is_x ? shift * rbM16 : 0u);
          uint32_t n16 = base16 + offN0_16 + (p.m_is_x ? 0u : shift * rbN16);
          const uint32_t d = tmem_base + static_cast<uint32_t>(t * p.nN);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t ad = (static_cast<uint64_t>(m_hi) << 32) | ((m16 & 0x3FFFu) | m_lo_extra);
            const uint64_t bd = (static_cast<uint64_t>(n_hi) << 32) | ((n16 & 0x3FFFu) | n_lo_extra);
            umma_bf16_ss(d, ad, bd, idesc, (ks > 0) ? 1u : accumulate);
            m16 += 16u * rbM16;
            n16 += 16u * rbN16;
          }
        }
        umma_commit(&empty[st]);
      }
      __syncwarp();
      accumulate = 1;
      if (++st == p.nstages) {
        st = 0;
        ph ^= 1u;
      }
    }
    if (leader) umma_commit(&done_bar);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> red.global into dW[k,k,cin,cout] =====================
    int nvalid = 0;
    for (int r = r0; r < r1; ++r) {
      const int y = r % p.h;
      const int iy = y + dy - p.pad;
      nvalid += (iy >= 0 && iy < p.h) ? 1 : 0;
    }
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    if (nvalid > 0) {
      const int ew = warp & 3;
      const int m = ew * 32 + lane;  // channel of the M operand
      for (int t = 0; t < nt; ++t) {
        const int tap = dy * p.k + dx0 + t;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(t * p.nN);
        for (int c0 = 0; c0 < p.nN; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (p.m_is_x) {  // lane = ci, columns = co: 16 consecutive floats per lane
            float* o = p.dw + (static_cast<size_t>(tap) * p.cin + m) * p.cout + c0;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              red_add_v4(o + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {  // lane = co, columns = ci: consecutive lanes hit consecutive floats
#pragma unroll
            for (int j = 0; j < 16; ++j)
              red_add_f32(p.dw + (static_cast<size_t>(tap) * p.cin + (c0 + j)) * p.cout + m, __uint_as_float(v[j]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace
}  // namespace mpg

extern "C" {

/* dw[k,k,cin,cout] (fp32, HWIO) += Conv2DBackpropFilter(x, dy) for a stride-1 SAME conv, on the tensor cores.
 * x: bf16 NHWC [n,h,w,cin], dy: bf16 NHWC [n,h,w,cout] (channel stride == channel count). Shapes: k in {1,3,5},
 * w % 16 == 0, w + k - 1 <= 256, one of (cin, cout) == 128 and the other in {32, 64, 128}.
 * Returns MPG_ENOSUP for other shapes (callers fall back to mpg_train_conv_wgrad). */
int mpg_train_conv_wgrad_tc(mpg_handle h, const void* x, const void* dy, float* dw, int n, int hh, int ww, int cin,
                            int cout, int k, void* stream) {
  using namespace mpg;
  MPG_CHECK_ARG(h && x && dy && dw && n > 0 && hh > 0 && ww > 0, "mpg_train_conv_wgrad_tc: bad argument");
  const bool shape_ok = (k == 1 || k == 3 || k == 5) && ww % 16 == 0 && ww + k - 1 <= 256 &&
                        ((cin == 128 && (cout == 32 || cout == 64 || cout == 128)) ||
                         (cout == 128 && (cin == 32 || cin == 64)));
  if (!shape_ok) {
    set_error("mpg_train_conv_wgrad_tc: unsupported shape cin=%d cout=%d k=%d w=%d", cin, cout, k, ww);
    return MPG_ENOSUP;
  }
  MPG_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0, "mpg_train_conv_wgrad_tc: tensors not 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WgTcParams p;
  memset(&p, 0, sizeof(p));
  p.n = n;
  p.h = hh;
  p.w = ww;
  p.cin = cin;
  p.cout = cout;
  p.k = k;
  p.pad = k / 2;
  p.m_is_x = cin == 128 ? 1 : 0;
  p.nN = p.m_is_x ? cout : cin;
  p.rbN = p.nN >= 64 ? 128 : 64;
  p.nblkN = p.nN >= 64 ? p.nN / 64 : 1;
  p.taps_per_group = 512 / p.nN < k ? 512 / p.nN : k;
  p.groups_per_dy = ceil_div(k, p.taps_per_group);
  const int groups = k * p.groups_per_dy;
  const int xw = ww + k - 1;
  // stage layout: [M block 0][M block 1][N block 0][N block 1], every block 1024-byte aligned
  const int m_row_px = p.m_is_x ? xw : ww, n_row_px = p.m_is_x ? ww : xw;
  const int m_blk = round_up(m_row_px * 128, 1024), n_blk = round_up(n_row_px * p.rbN, 1024);
  p.offM1 = m_blk;
  p.offN0 = 2 * m_blk;
  p.offN1 = p.offN0 + n_blk;
  p.stage_bytes = 2 * m_blk + p.nblkN * n_blk;
  p.x_row_bytes = xw * (p.m_is_x ? 128 : p.rbN);
  p.dy_row_bytes = ww * (p.m_is_x ? p.rbN : 128);
  int nst = (200 * 1024) / p.stage_bytes;
  p.nstages = nst > kWgMaxStages ? kWgMaxStages : nst;
  MPG_CHECK_ARG(p.nstages >= 2, "mpg_train_conv_wgrad_tc: row of %d pixels does not fit two pipeline stages", ww);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(p.taps_per_group * p.nN)) cols <<= 1;
  p.tmem_cols = cols;
  const int total_rows = n * hh;
  int splits = h->sm_count / groups;
  if (splits < 1) splits = 1;
  if (splits > total_rows) splits = total_rows;
  p.rows_per_split = ceil_div(total_rows, splits);
  splits = ceil_div(total_rows, p.rows_per_split);
  p.dw = dw;

  CUtensorMap tm_x, tm_dy;
  {
    const int cw = p.m_is_x ? 64 : (p.rbN >> 1);
    const uint64_t cs = static_cast<uint64_t>(cin) * 2;
    const uint64_t dims[4] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(ww), static_cast<uint64_t>(hh), static_cast<uint64_t>(n)};
    const uint64_t strides[3] = {cs, cs * ww, cs * ww * hh};
    const uint32_t box[4] = {static_cast<uint32_t>(cw), static_cast<uint32_t>(xw), 1u, 1u};
    int r = encode_tmap(h, &tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box,
                        cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (r) return r;
  }
  {
    const int cw = p.m_is_x ? (p.rbN >> 1) : 64;
    const uint64_t cs = static_cast<uint64_t>(cout) * 2;
    const uint64_t dims[4] = {static_cast<uint64_t>(cout), static_cast<uint64_t>(ww), static_cast<uint64_t>(hh), static_cast<uint64_t>(n)};
    const uint64_t strides[3] = {cs, cs * ww, cs * ww * hh};
    const uint32_t box[4] = {static_cast<uint32_t>(cw), static_cast<uint32_t>(ww), 1u, 1u};
    int r = encode_tmap(h, &tm_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dy, dims, strides, box,
                        cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (r) return r;
  }
  const size_t smem_bytes = static_cast<size_t>(p.nstages) * p.stage_bytes + 1024;
  DeviceGuard guard(h->device);
  static size_t attr_set[kMaxDevices] = {};  // per device: cudaFuncSetAttribute applies to the current device only
  const bool cached = h->device >= 0 && h->device < kMaxDevices;
  if (!cached || smem_bytes > attr_set[h->device]) {
    MPG_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
    if (cached) attr_set[h->device] = smem_bytes;
  }
  conv_wgrad_tc_kernel<<<dim3(static_cast<unsigned>(groups), static_cast<unsigned>(splits)), 256, smem_bytes, st>>>(tm_x, tm_dy, p);
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

}  // extern "C"
