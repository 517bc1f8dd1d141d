// Row-streaming tcgen05 convolution, vertical taps folded into N (design: conv_vfold.cuh).
#include <string.h>

#include "conv_vfold.cuh"
#include "ptx.cuh"

namespace mpg {

namespace {

// A CTA pair walks its contiguous range of flattened (image, strip pair, row) units as a sequence of segments: `len`
// output rows y0.. of image n, this CTA's strip starting at pixel x0.
struct VfIter {
  int cur, end;
  int n, x0, y0, len;
  __device__ __forceinline__ bool next(const VfoldParams& p, uint32_t rank) {
    if (cur >= end) return false;
    const int u = cur / p.h;
    y0 = cur - u * p.h;
    len = min(p.h - y0, end - cur);
    n = u / p.strips2;
    x0 = (2 * (u - n * p.strips2) + static_cast<int>(rank)) * kVfStrip;
    cur += len;
    return true;
  }
};

// NCHW: 8-channel chunks of the output an epilogue warp owns (registers: (KS-1) * NCHW * 8 running sums per thread);
// G: epilogue warps per TMEM lane quarter (warp group g owns chunks [g*NCHW, (g+1)*NCHW)). The epilogue reads k times
// more TMEM than it stores, so it needs many warps in flight: the block is 4 role warps + 4*G epilogue warps.
template <int CK, int KS, int NCHW, int G>
__global__ void __launch_bounds__(128 + 128 * G, 1)
conv_vfold_kernel(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1,
                  const VfoldParams p) {
  constexpr int RB = CK * 2;  // bytes per pixel of one K-chunk == swizzle span
  constexpr uint32_t LAYOUT = (RB == 128) ? 2u : (RB == 64 ? 4u : 6u);
  constexpr uint32_t SBO = 8u * RB;
  constexpr int KSTEPS = CK / 16;
  constexpr int PAD = KS >> 1;
  constexpr int WIN = kVfStrip + KS - 1;  // staged pixels per image row
  constexpr uint32_t PX16 = RB >> 4;      // one pixel in 16-byte descriptor units

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_a[kVfMaxStagesA], empty_a[kVfMaxStagesA];
  __shared__ __align__(8) uint64_t tmem_full[kVfMaxBufs], tmem_empty[kVfMaxBufs];
  __shared__ __align__(8) uint64_t full_b, b_ready;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_shift[64];
  __shared__ float s_pn[2][G][128];  // pixel_norm partial sums: [slot][warp group][accumulator row]

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint8_t* smB = smem;                   // resident weights first (b_bytes is a 1024-byte multiple)
  uint8_t* smA = smem + p.b_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = static_cast<int>(blockIdx.x >> 1);
  const int row_begin = pair * p.rows_per_pair;
  const int row_end = min(row_begin + p.rows_per_pair, p.total_rows);
  constexpr int epi_active = 4 * G;

  if (threadIdx.x < 64) s_shift[threadIdx.x] = threadIdx.x < p.cp ? p.shift[threadIdx.x] : 0.0f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x0);
    if (p.nseg > 1) tma_prefetch_desc(&tm_x1);
    for (int i = 0; i < p.na; ++i) {
      mbar_init(&full_a[i], 1);
      mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < kVfMaxBufs; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * epi_active);  // one arrive per active epilogue warp of both CTAs
    }
    mbar_init(&full_b, 1);
    mbar_init(&b_ready, 2);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc2(&tmem_base_slot, p.tmem_cols);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== A producer: one staged image row per (segment, Cin chunk) =====================
    if (lane == 0) {
      constexpr uint32_t bytes = static_cast<uint32_t>(WIN) * RB;
      int st = 0;
      uint32_t ph = 0;
      VfIter it{row_begin, row_end, 0, 0, 0, 0};
      while (it.next(p, rank)) {
        const int nrows = it.len + KS - 1;
        for (int wr = 0; wr < nrows; ++wr) {
          const int gy = it.y0 - PAD + wr;
          for (int s = 0; s < p.nseg; ++s) {
            const CUtensorMap* tm = (s == 0) ? &tm_x0 : &tm_x1;
            for (int ch = 0; ch < p.seg_nchunk[s]; ++ch) {
              mbar_wait(&empty_a[st], ph ^ 1u);
              // both CTAs' rows complete on the LEADER's barrier, which expects the bytes of both
              if (rank == 0) mbar_arrive_expect_tx(&full_a[st], 2u * bytes);
              tma_load_4d_2cta(smA + static_cast<size_t>(st) * p.a_stage_bytes, tm, mapa_u32(smem_u32(&full_a[st]), 0),
                               ch * CK, it.x0 - PAD, gy, it.n);
              if (++st == p.na) {
                st = 0;
                ph ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== resident weights: this CTA's half of every tile, loaded once =====================
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpacked) + static_cast<size_t>(rank) * p.b_bytes;
      mbar_arrive_expect_tx(&full_b, static_cast<uint32_t>(p.b_bytes));
      for (int off = 0; off < p.b_bytes; off += 32768) {
        const int nb = min(32768, p.b_bytes - off);
        bulk_load_1d(smB + off, wsrc + off, static_cast<uint32_t>(nb), &full_b);
      }
      mbar_wait(&full_b, 0);
      mbar_arrive_cluster(mapa_u32(smem_u32(&b_ready), 0));
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) =====================
    const uint32_t fmt = p.in_dtype == MPG_F16 ? 0u : 1u;
    const uint32_t idesc = umma_idesc_f16kind(256, p.npad, fmt);
    const uint32_t idesc_sc = umma_idesc_f16kind(256, p.n_sc, fmt);
    constexpr uint32_t DESC_HI = (SBO >> 4) | (1u << 14) | (LAYOUT << 29);
    const uint32_t smA_lo = ((smem_u32(smA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t smB_lo = ((smem_u32(smB) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a_stage16 = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
    const uint32_t b_tile16 = static_cast<uint32_t>(p.b_tile_bytes) >> 4;
    const uint32_t b_sc_tile16 = static_cast<uint32_t>(p.b_sc_tile_bytes) >> 4;
    const uint32_t npad = static_cast<uint32_t>(p.npad);
    const uint32_t sc_col = static_cast<uint32_t>(p.sc_col);
    const int na = p.na, nbuf = p.nbuf, nseg = p.nseg;
    const int nch0 = p.seg_nchunk[0], kl0 = p.seg_klast[0];
    const int nch1 = p.seg_nchunk[1], kl1 = p.seg_klast[1];
    const bool do_mma = !(p.dbg & 4);
    const bool leader = elect_one() != 0;
    int sa = 0;
    uint32_t pa = 0;
    int buf = 0;
    uint32_t use = 0;
    mbar_wait_cluster(&b_ready, 0);
    tc_fence_after();
    VfIter it{row_begin, row_end, 0, 0, 0, 0};
    while (it.next(p, rank)) {
      const int nrows = it.len + KS - 1;
      for (int wr = 0; wr < nrows; ++wr) {
        mbar_wait(&tmem_empty[buf], (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t dbase = tmem_base + static_cast<uint32_t>(buf) * npad;
        uint32_t b_cur = smB_lo;
        uint32_t accumulate = 0;
        for (int ch = 0; ch < nch0; ++ch) {
          mbar_wait_cluster(&full_a[sa], pa);
          tc_fence_after();
          const uint32_t a_row = smA_lo + static_cast<uint32_t>(sa) * a_stage16;
          const int nk = (ch == nch0 - 1) ? kl0 : KSTEPS;
          if (leader && do_mma) {
#pragma unroll
            for (int dx = 0; dx < KS; ++dx) {
              const uint32_t a_tap = a_row + static_cast<uint32_t>(dx) * PX16;
              const uint32_t b_tap = b_cur + static_cast<uint32_t>(dx) * b_tile16;
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                if (k < nk) {
                  const uint64_t ad = (static_cast<uint64_t>(DESC_HI) << 32) | (a_tap + k * 2);
                  const uint64_t bd = (static_cast<uint64_t>(DESC_HI) << 32) | (b_tap + k * 2);
                  umma_bf16_ss_2cta(dbase, ad, bd, idesc, (dx > 0 || k > 0) ? 1u : accumulate);
                }
              }
            }
          }
          b_cur += static_cast<uint32_t>(KS) * b_tile16;
          accumulate = 1;
          if (leader) umma_commit_2cta(&empty_a[sa], 3);
          __syncwarp();
          if (++sa == na) {
            sa = 0;
            pa ^= 1u;
          }
        }
        if (nseg > 1) {
          // 1x1 shortcut: the staged row of the second input, centre pixel shift, into the centre-dy column block
          for (int ch = 0; ch < nch1; ++ch) {
            mbar_wait_cluster(&full_a[sa], pa);
            tc_fence_after();
            const uint32_t a_tap = smA_lo + static_cast<uint32_t>(sa) * a_stage16 + static_cast<uint32_t>(PAD) * PX16;
            const int nk = (ch == nch1 - 1) ? kl1 : KSTEPS;
            if (leader && do_mma) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                if (k < nk) {
                  const uint64_t ad = (static_cast<uint64_t>(DESC_HI) << 32) | (a_tap + k * 2);
                  const uint64_t bd = (static_cast<uint64_t>(DESC_HI) << 32) | (b_cur + k * 2);
                  umma_bf16_ss_2cta(dbase + sc_col, ad, bd, idesc_sc, 1u);
                }
              }
            }
            b_cur += b_sc_tile16;
            if (leader) umma_commit_2cta(&empty_a[sa], 3);
            __syncwarp();
            if (++sa == na) {
              sa = 0;
              pa ^= 1u;
            }
          }
        }
        if (leader) umma_commit_2cta(&tmem_full[buf], 3);
        __syncwarp();
        if (++buf == nbuf) {
          buf = 0;
          ++use;
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + epi_active) {
    // ===================== epilogue: running sums over dy in registers + shift + act + store =====================
    const int ew = warp & 3;          // TMEM lane quarter: pixels [32*ew, 32*ew + 32) of the strip
    const int grp = (warp - 4) >> 2;  // warp group: owns chunks [grp*NCHW, grp*NCHW + NCHW)
    const int nchunks = p.cp >> 3;
    const float act_a = p.act == MPG_ACT_LRELU ? 0.6f : 1.0f;
    const float act_b = p.act == MPG_ACT_LRELU ? 0.4f : 0.0f;
    const bool is_relu = p.act == MPG_ACT_RELU;
    const bool is_tanh = p.act == MPG_ACT_TANH;
    const float inv_c = 1.0f / static_cast<float>(p.cout);
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const int cp = p.cp;
    int pn_slot = 0;
    int buf = 0;
    uint32_t use = 0;
    VfIter it{row_begin, row_end, 0, 0, 0, 0};
    while (it.next(p, rank)) {
      const int gx = it.x0 + ew * 32 + lane;
      const bool col_ok = gx < p.w && !(p.dbg & 1);
      float S[KS - 1][NCHW * 8];
#pragma unroll
      for (int j = 0; j < KS - 1; ++j)
#pragma unroll
        for (int i = 0; i < NCHW * 8; ++i) S[j][i] = 0.0f;
      const int nrows = it.len + KS - 1;
#pragma unroll 1
      for (int wr = 0; wr < nrows; ++wr) {
        mbar_wait(&tmem_full[buf], use & 1u);
        tc_fence_after();
        const uint32_t taddr = lane_addr + static_cast<uint32_t>(buf * p.npad);
        float o[NCHW * 8];
        if (!(p.dbg & 2)) {
#pragma unroll
          for (int c = 0; c < NCHW; ++c) {
            const int cc = grp * NCHW + c;
            if (cc < nchunks) {
              uint32_t r[KS][8];
#pragma unroll
              for (int dy = 0; dy < KS; ++dy) tmem_ld8(taddr + static_cast<uint32_t>(dy * cp + cc * 8), r[dy]);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                o[c * 8 + j] = S[KS - 2][c * 8 + j] + __uint_as_float(r[KS - 1][j]);
#pragma unroll
                for (int dy = KS - 2; dy >= 1; --dy) S[dy][c * 8 + j] = S[dy - 1][c * 8 + j] + __uint_as_float(r[dy][j]);
                S[0][c * 8 + j] = __uint_as_float(r[0][j]);
              }
            }
          }
        }
        // the accumulator is free as soon as its columns are in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[buf]), 0));
          else mbar_arrive(&tmem_empty[buf]);
        }
        if (++buf == p.nbuf) {
          buf = 0;
          ++use;
        }
        const int orow = wr - (KS - 1);  // output row (relative to y0) completed by this image row
        if (orow < 0 || (p.dbg & 2)) continue;
        const int y = it.y0 + orow;
        const size_t pix = (static_cast<size_t>(it.n) * p.h + y) * p.w + gx;
        float ssq = 0.0f;
#pragma unroll
        for (int c = 0; c < NCHW; ++c) {
          const int cc = grp * NCHW + c;
          if (cc < nchunks) {
            float rs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (p.resid != nullptr && cc == 0 && col_ok) {
              const float4* rp = reinterpret_cast<const float4*>(p.resid + pix * 8);
              const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
              rs[0] = r0.x; rs[1] = r0.y; rs[2] = r0.z; rs[3] = r0.w;
              rs[4] = r1.x; rs[5] = r1.y; rs[6] = r1.z; rs[7] = r1.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = o[c * 8 + j] + s_shift[cc * 8 + j] + rs[j];
              float v = is_relu ? fmaxf(x, 0.0f) : fmaf(act_b, fabsf(x), act_a * x);
              if (is_tanh) v = tanhf(x);
              o[c * 8 + j] = v;
              ssq = fmaf(v, v, ssq);
            }
          }
        }
        if (G > 1 && p.pixel_norm) {  // add the other warp groups' chunks (same pixels, other channels)
          s_pn[pn_slot][grp][ew * 32 + lane] = ssq;
          named_bar_sync(1 + ew, 32 * G);
          ssq = 0.0f;
#pragma unroll
          for (int g2 = 0; g2 < G; ++g2) ssq += s_pn[pn_slot][g2][ew * 32 + lane];
          pn_slot ^= 1;
        }
        const float rn = p.pixel_norm ? rsqrtf(ssq * inv_c + 1e-8f) : 1.0f;  // tools_wscale/GAN.py:472-474
        if (col_ok) {
          if (p.out_dtype != MPG_F32) {
            const int od = p.out_dtype;
            uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + pix * p.out_cstride;
#pragma unroll
            for (int c = 0; c < NCHW; ++c) {
              const int cc = grp * NCHW + c;
              if (cc < nchunks) {
                uint4 q;
                q.x = pack_h16x2(o[c * 8 + 0] * rn, o[c * 8 + 1] * rn, od);
                q.y = pack_h16x2(o[c * 8 + 2] * rn, o[c * 8 + 3] * rn, od);
                q.z = pack_h16x2(o[c * 8 + 4] * rn, o[c * 8 + 5] * rn, od);
                q.w = pack_h16x2(o[c * 8 + 6] * rn, o[c * 8 + 7] * rn, od);
                *reinterpret_cast<uint4*>(op + cc * 8) = q;
              }
            }
          } else {
            float* op = reinterpret_cast<float*>(p.out) + pix * p.out_cstride;
#pragma unroll
            for (int c = 0; c < NCHW; ++c) {
              const int cc = grp * NCHW + c;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int co = cc * 8 + j;
                if (cc < nchunks && co < p.out_cstride) op[co] = (co < p.cout) ? o[c * 8 + j] * rn : 0.0f;
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, p.tmem_cols);
  }
}

typedef void (*VfKernel)(const CUtensorMap, const CUtensorMap, const VfoldParams);

// (chunks per warp, warp groups) configurations: nchw * g >= cp / 8
template <int CK, int KS>
VfKernel vf_kernel_cfg(int nchw, int g) {
  if (nchw == 1) {
    switch (g) {
      case 1: return conv_vfold_kernel<CK, KS, 1, 1>;
      case 2: return conv_vfold_kernel<CK, KS, 1, 2>;
      case 3: return conv_vfold_kernel<CK, KS, 1, 3>;
      default: return conv_vfold_kernel<CK, KS, 1, 4>;
    }
  }
  if (nchw == 2) {
    switch (g) {
      case 1: return conv_vfold_kernel<CK, KS, 2, 1>;
      case 2: return conv_vfold_kernel<CK, KS, 2, 2>;
      case 3: return conv_vfold_kernel<CK, KS, 2, 3>;
      default: return conv_vfold_kernel<CK, KS, 2, 4>;
    }
  }
  return conv_vfold_kernel<CK, KS, 3, 2>;
}

VfKernel vf_kernel(int ck, int ks, int nchw, int g) {
  if (ck == 64) return ks == 5 ? vf_kernel_cfg<64, 5>(nchw, g) : vf_kernel_cfg<64, 3>(nchw, g);
  return ks == 5 ? vf_kernel_cfg<32, 5>(nchw, g) : vf_kernel_cfg<32, 3>(nchw, g);
}

}  // namespace

static size_t g_vf_smem_attr[kMaxDevices][64] = {};  // per device: cudaFuncSetAttribute applies to the current device only

int vfold_set_smem_attr(int device, int ck, int ks, int nchw, int g, size_t smem_bytes) {
  const int slot = (ck == 64 ? 0 : 32) + (ks == 5 ? 0 : 16) + (nchw - 1) * 4 + (g - 1);
  const bool cached = device >= 0 && device < kMaxDevices;
  if (cached && smem_bytes <= g_vf_smem_attr[device][slot]) return 0;
  DeviceGuard guard(device);
  cudaError_t e = cudaFuncSetAttribute(vf_kernel(ck, ks, nchw, g), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem_bytes));
  if (e == cudaSuccess && cached) g_vf_smem_attr[device][slot] = smem_bytes;
  return static_cast<int>(e);
}

int vfold_launch(int ck, int nchw, const CUtensorMap& tm_x0, const CUtensorMap& tm_x1, const VfoldParams& p, int grid,
                 size_t smem_bytes, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(static_cast<unsigned>(128 + 128 * p.epi_groups), 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, vf_kernel(ck, p.ks, nchw, p.epi_groups), tm_x0, tm_x1, p));
}

}  // namespace mpg
