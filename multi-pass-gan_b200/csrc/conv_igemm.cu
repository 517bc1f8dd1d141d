// tcgen05 implicit-GEMM convolution kernel (see conv_igemm.cuh for the design).
#include <string.h>

#include "conv_igemm.cuh"
#include "ptx.cuh"

namespace mpg {

namespace {

// Branch-free activation: act(v) = ca*v + cb*|v| covers none (1,0), relu (0.5,0.5: exact) and the
// reference's lrelu 0.6*x + 0.4*|x| (tools_wscale/GAN.py:733-737); tanh is a rare separate pass.
__device__ __forceinline__ void epi_chunk16(uint32_t taddr, const float* __restrict__ shift16, float ca, float cb,
                                            bool is_tanh, float (&v)[16]) {
  uint32_t r[16];
  tmem_ld16(taddr, r);
  const float4* s4 = reinterpret_cast<const float4*>(shift16);
  const float4 s0 = s4[0], s1 = s4[1], s2 = s4[2], s3 = s4[3];
  const float sh[16] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w, s3.x, s3.y, s3.z, s3.w};
  tmem_ld_wait();
  if (ca == 0.5f && cb == 0.5f) {  // relu: one FMNMX instead of FMUL + FFMA (same bits: 0.5x + 0.5|x| is exact)
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(__uint_as_float(r[j]) + sh[j], 0.0f);
    return;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float x = __uint_as_float(r[j]) + sh[j];
    v[j] = fmaf(cb, fabsf(x), ca * x);
  }
  if (is_tanh) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = tanhf(__uint_as_float(r[j]) + sh[j]);
  }
}

struct TileCoord {
  int n, y0, x0;
};
__device__ __forceinline__ TileCoord decode_tile(int t, const IgemmParams& p) {
  const int per_img = p.tiles_x * p.tiles_y;
  TileCoord c;
  c.n = t / per_img;
  const int r = t - c.n * per_img;
  const int ty = r / p.tiles_x;
  c.y0 = ty * kIgTileH;
  c.x0 = (r - ty * p.tiles_x) * kIgTileW;
  return c;
}

// PAIR: the two CTAs of a cluster form a cta_group::2 pair. Each CTA owns its own 16x16-pixel tile (A operand,
// accumulators, epilogue) but holds only HALF of every weight tile; the leader issues M=256 MMAs that read both
// halves, so the L2->SM weight traffic per SM halves (measured: an SM ingests ~27 B/clk from L2, which bounds
// the single-CTA kernel at 970 KB per tile vs 26 K MMA cycles).
template <int CK, bool PAIR, bool SIDE>
__global__ void __launch_bounds__(kIgMaxThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1,
                  const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_y,
                  const IgemmParams p) {
  constexpr int RB = CK * 2;  // bytes per pixel row of one K-chunk == swizzle span
  constexpr uint32_t LAYOUT = (RB == 128) ? 2u : (RB == 64 ? 4u : 6u);
  constexpr uint32_t SBO = 8u * RB;  // 8 pixels per swizzle group
  constexpr int KSTEPS = CK / 16;    // UMMA K = 16 bf16

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_a[kIgMaxStagesA], empty_a[kIgMaxStagesA];
  __shared__ __align__(8) uint64_t full_b[kIgMaxStagesB], empty_b[kIgMaxStagesB];
  __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_shift[256];
  __shared__ float s_pn[2][2][128];  // pixel_norm partial sums: [accumulator][column half][accumulator row]

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + static_cast<size_t>(p.na) * p.a_stage_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // tile schedule: CTA (or CTA pair) q takes tiles q, q+G, ...; in PAIR mode the pair takes tiles (2q, 2q+1) and
  // both CTAs run the same number of iterations (an odd last tile is recomputed by the peer, not stored)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int tstep = PAIR ? static_cast<int>(gridDim.x >> 1) * 2 : static_cast<int>(gridDim.x);
  const int tfirst = PAIR ? static_cast<int>(blockIdx.x >> 1) * 2 + static_cast<int>(rank) : static_cast<int>(blockIdx.x);
#define MPG_TILE_LOOP(t) for (int t = tfirst; t - static_cast<int>(rank) < p.num_tiles; t += tstep)
#define MPG_TILE_CLAMP(t) ((t) < p.num_tiles ? (t) : p.num_tiles - 1)

  for (int i = threadIdx.x; i < p.npad; i += blockDim.x) s_shift[i] = p.shift[i];
  // epilogue warps: 4 (256-thread launch) or 8 (384-thread launch: two warps per TMEM lane quarter split the columns)
  const int n_epi_warps = (static_cast<int>(blockDim.x) >> 5) - 4;
  // (pixel_norm needs the whole pixel's sum of squares: the two warps of a lane quarter each reduce their column half
  //  and exchange the partial sums through shared memory -- s_pn below -- so all 8 warps stay usable)
  const int epi_active = (n_epi_warps == 8) ? 8 : 4;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x0);
    if (p.nseg > 1) tma_prefetch_desc(&tm_x1);
    tma_prefetch_desc(&tm_w);
    if (p.tma_store) tma_prefetch_desc(&tm_y);
    for (int i = 0; i < p.na; ++i) {
      mbar_init(&full_a[i], 1);
      mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < p.nb; ++i) {
      mbar_init(&full_b[i], 1);
      mbar_init(&empty_b[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], (PAIR ? 2 : 1) * epi_active);  // one arrive per active epilogue warp (both CTAs in PAIR mode)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc2(&tmem_base_slot, p.tmem_cols);
      tmem_relinquish2();
    } else {
      tmem_alloc(&tmem_base_slot, p.tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const float act_a = p.act == MPG_ACT_RELU ? 0.5f : (p.act == MPG_ACT_LRELU ? 0.6f : 1.0f);
  const float act_b = p.act == MPG_ACT_RELU ? 0.5f : (p.act == MPG_ACT_LRELU ? 0.4f : 0.0f);
  const bool act_tanh = p.act == MPG_ACT_TANH;

  if (warp == 0) {
    // ===================== A producer: one halo image per (segment, chunk) ===============
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      MPG_TILE_LOOP(t) {
        const TileCoord tc = decode_tile(MPG_TILE_CLAMP(t), p);
        for (int s = 0; s < p.nseg; ++s) {
          const int ks = p.seg_ks[s];
          const int pad = ks >> 1;
          const CUtensorMap* tm = (s == 0) ? &tm_x0 : &tm_x1;
          const uint32_t hbytes = static_cast<uint32_t>((kIgTileH + ks - 1) * (kIgTileW + ks - 1) * RB);
          for (int ch = 0; ch < p.seg_nchunk[s]; ++ch) {
            mbar_wait(&empty_a[st], ph ^ 1u);
            if (PAIR) {
              // both CTAs' images complete on the LEADER's barrier, which expects the bytes of both
              if (rank == 0) mbar_arrive_expect_tx(&full_a[st], 2u * hbytes);
              tma_load_4d_2cta(smA + static_cast<size_t>(st) * p.a_stage_bytes, tm, mapa_u32(smem_u32(&full_a[st]), 0),
                               ch * CK, tc.x0 - pad, tc.y0 - pad, tc.n);
            } else {
              mbar_arrive_expect_tx(&full_a[st], hbytes);
              tma_load_4d(smA + static_cast<size_t>(st) * p.a_stage_bytes, tm, &full_a[st], ch * CK, tc.x0 - pad,
                          tc.y0 - pad, tc.n);
            }
            if (++st == p.na) {
              st = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== B producer: weight tiles, `bgroup` consecutive dy taps per stage =====
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      const int nrows = PAIR ? p.npad / 2 : p.npad;  // PAIR: this CTA holds weight rows [rank*N/2, (rank+1)*N/2)
      const int row0 = PAIR ? static_cast<int>(rank) * nrows : 0;
      const uint32_t tile_bytes = static_cast<uint32_t>(nrows * RB);
      if (p.bres) {
        // resident weights: every k-tile is loaded once per CTA and stays in shared memory
        // (PAIR: each CTA keeps its half of every tile; both halves complete on the leader's barrier)
        if (PAIR) {
          if (rank == 0) mbar_arrive_expect_tx(&full_b[0], 2u * tile_bytes * static_cast<uint32_t>(p.ktiles));
          const uint32_t bar = mapa_u32(smem_u32(&full_b[0]), 0);
          for (int kt = 0; kt < p.ktiles; ++kt)
            tma_load_2d_2cta(smB + static_cast<size_t>(kt) * p.b_tile_bytes, &tm_w, bar, 0, kt * p.npad + row0);
        } else {
          mbar_arrive_expect_tx(&full_b[0], tile_bytes * static_cast<uint32_t>(p.ktiles));
          for (int kt = 0; kt < p.ktiles; ++kt)
            tma_load_2d(smB + static_cast<size_t>(kt) * p.b_tile_bytes, &tm_w, &full_b[0], 0, kt * p.npad);
        }
      }
      MPG_TILE_LOOP(t) {
        if (p.bres) break;
        int kt = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int ks = p.seg_ks[s];
          const int gb = p.bgroup ? ks : 1;
          const int ngroups = p.seg_nchunk[s] * ks * (ks / gb);
          for (int i = 0; i < ngroups; ++i) {
            mbar_wait(&empty_b[st], ph ^ 1u);
            uint8_t* dst = smB + static_cast<size_t>(st) * p.b_stage_bytes;
            if (PAIR) {
              if (rank == 0) mbar_arrive_expect_tx(&full_b[st], 2u * tile_bytes * gb);
              const uint32_t bar = mapa_u32(smem_u32(&full_b[st]), 0);
              for (int g = 0; g < gb; ++g, ++kt)
                tma_load_2d_2cta(dst + static_cast<size_t>(g) * p.b_tile_bytes, &tm_w, bar, 0, kt * p.npad + row0);
            } else {
              mbar_arrive_expect_tx(&full_b[st], tile_bytes * gb);
              for (int g = 0; g < gb; ++g, ++kt)
                tma_load_2d(dst + static_cast<size_t>(g) * p.b_tile_bytes, &tm_w, &full_b[st], 0, kt * p.npad);
            }
            if (++st == p.nb) {
              st = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1 && (!PAIR || rank == 0)) {
    // ===================== MMA issuer (PAIR: leader CTA only) ==========================================================
    // The whole warp walks the loop (warp-uniform control flow keeps the descriptors in uniform registers); one lane,
    // elected once, issues the tcgen05.mma / tcgen05.commit instructions. The loop body is kept to a few 32-bit adds
    // per MMA: with CK = 32 a (dx, dy) tap is only 4 MMAs (256 tensor-pipe cycles), and the earlier loop (64-bit stage
    // address products, per-tap re-derivation of both descriptor words) took ~290 cycles per tap -- the 32->128 layer
    // was bound by this thread, not by the tensor pipe (knock-out runs, tools/thin_probe.py "skeleton").
    const uint32_t idesc = umma_idesc_f16kind(PAIR ? 256 : 128, p.npad, p.in_dtype == MPG_F16 ? 0u : 1u);
    // smem-descriptor words: hi = SBO | version 1 | layout type; lo = (addr >> 4) | LBO(1)
    constexpr uint32_t B_HI = (SBO >> 4) | (1u << 14) | (LAYOUT << 29);
    constexpr uint32_t PX16 = RB >> 4;  // one pixel of the halo image in 16-byte descriptor units
    const uint32_t smA_lo = ((smem_u32(smA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t smB_lo = ((smem_u32(smB) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a_stage16 = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
    const uint32_t b_stage16 = static_cast<uint32_t>(p.b_stage_bytes) >> 4;
    const uint32_t b_tile16 = static_cast<uint32_t>(p.b_tile_bytes) >> 4;
    const uint32_t npad = static_cast<uint32_t>(p.npad);
    const bool bres = p.bres != 0;
    const bool do_mma = !(p.dbg & 4);
    const int na = p.na, nb = p.nb, nseg = p.nseg;
    const bool leader = elect_one() != 0;
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int it = 0;
    if (bres) {
      mbar_wait(&full_b[0], 0);
      tc_fence_after();
    }
    for (int t = tfirst; t - static_cast<int>(rank) < p.num_tiles; t += tstep, ++it) {
      const int buf = it & 1;
      mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + static_cast<uint32_t>(buf * 2) * npad;
      const uint32_t d1 = d0 + npad;
      uint32_t accumulate = 0;
      uint32_t b_res = smB_lo;  // resident-weights cursor (k-tile order == issue order)
      for (int s = 0; s < nseg; ++s) {
        const int ks = p.seg_ks[s];
        const int gb = p.bgroup ? ks : 1;
        // 16 image rows x 8 px per accumulator: 8-row-group stride = one halo row; the start is shifted by whole
        // pixels, i.e. NOT aligned to the swizzle atom (legal: the swizzle is a function of the absolute address)
        const uint32_t row16 = static_cast<uint32_t>(kIgTileW + ks - 1) * PX16;
        const uint32_t a_hi = row16 | (1u << 14) | (LAYOUT << 29);
        const int nchunk = p.seg_nchunk[s];
        for (int ch = 0; ch < nchunk; ++ch) {
          mbar_wait(&full_a[sa], pa);
          tc_fence_after();
          const uint32_t a_img = smA_lo + static_cast<uint32_t>(sa) * a_stage16;
          for (int dx = 0; dx < ks; ++dx) {
            uint32_t a_tap = a_img + static_cast<uint32_t>(dx) * PX16;
            for (int dy0 = 0; dy0 < ks; dy0 += gb) {
              uint32_t b_lo;
              if (bres) {
                b_lo = b_res;
                b_res += static_cast<uint32_t>(gb) * b_tile16;
              } else {
                mbar_wait(&full_b[sb], pb);
                tc_fence_after();
                b_lo = smB_lo + static_cast<uint32_t>(sb) * b_stage16;
              }
              if (leader) {
                uint32_t a_g = a_tap, b_g = b_lo, acc = accumulate;
                if (do_mma) {
                  for (int g = 0; g < gb; ++g) {
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k) {
                      const uint64_t bd = (static_cast<uint64_t>(B_HI) << 32) | (b_g + k * 2);
                      const uint64_t ad0 = (static_cast<uint64_t>(a_hi) << 32) | (a_g + k * 2);
                      const uint64_t ad1 = (static_cast<uint64_t>(a_hi) << 32) | (a_g + 8u * PX16 + k * 2);
                      if (PAIR) {
                        umma_bf16_ss_2cta(d0, ad0, bd, idesc, (k > 0) ? 1u : acc);
                        umma_bf16_ss_2cta(d1, ad1, bd, idesc, (k > 0) ? 1u : acc);
                      } else {
                        umma_bf16_ss(d0, ad0, bd, idesc, (k > 0) ? 1u : acc);
                        umma_bf16_ss(d1, ad1, bd, idesc, (k > 0) ? 1u : acc);
                      }
                    }
                    acc = 1;
                    a_g += row16;
                    b_g += b_tile16;
                  }
                }
                if (!bres) {
                  if (PAIR) umma_commit_2cta(&empty_b[sb], 3);
                  else umma_commit(&empty_b[sb]);
                }
              }
              __syncwarp();
              accumulate = 1;
              a_tap += static_cast<uint32_t>(gb) * row16;
              if (!bres && ++sb == nb) {
                sb = 0;
                pb ^= 1u;
              }
            }
          }
          if (leader) {
            if (PAIR) umma_commit_2cta(&empty_a[sa], 3);
            else umma_commit(&empty_a[sa]);
          }
          __syncwarp();
          if (++sa == na) {
            sa = 0;
            pa ^= 1u;
          }
        }
      }
      if (leader) {
        if (PAIR) umma_commit_2cta(&tmem_full[buf], 3);
        else umma_commit(&tmem_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 4 + epi_active) {
    // ===================== epilogue: TMEM -> registers -> (staged TMA store | direct global stores) ====================
    const int ew = warp & 3;  // TMEM lane quarter this warp may access
    const int m = ew * 32 + lane;
    // accumulator row m -> pixel: 16 image rows x 8 px
    const int prow = m >> 3;
    const int pcol = m & 7;
    const int ups = p.upsample;
    const int oh = p.h * ups, ow = p.w * ups;
    const float inv_c = 1.0f / static_cast<float>(p.cout);
    const int nchunk16 = p.npad >> 4;
    const int csplit = (epi_active == 8) ? ((nchunk16 + 1) >> 1) * 16 : p.npad;
    const int cbeg = (warp >= 8) ? csplit : 0;
    const int cend = (warp >= 8) ? p.npad : csplit;
    const bool st32 = (p.out_cstride % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0) && !(p.dbg & 8);
    uint8_t* stg = smem + p.stage_off + static_cast<size_t>(warp - 4) * 4096;  // TMA-store staging of this warp
    // side output: weights [128][8] and the partial-sum exchange between the two warps of a lane quarter
    const float4* s_sw = reinterpret_cast<const float4*>(smem + p.side_off);
    float4* s_sx = reinterpret_cast<float4*>(smem + p.side_off + 4096);
    if (SIDE && p.side) {
      for (int i = threadIdx.x - 128; i < 256; i += static_cast<int>(blockDim.x) - 128)
        reinterpret_cast<float4*>(smem + p.side_off)[i] = reinterpret_cast<const float4*>(p.side_w)[i];
      asm volatile("bar.sync 5, 256;" ::: "memory");  // the 8 epilogue warps
    }
    int it = 0;
    for (int t = tfirst; t - static_cast<int>(rank) < p.num_tiles; t += tstep, ++it) {
      const int buf = it & 1;
      const TileCoord tc = decode_tile(MPG_TILE_CLAMP(t), p);
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int acc = 0; acc < 2; ++acc) {
        const int y = tc.y0 + prow;
        const int x = tc.x0 + acc * 8 + pcol;
        const bool valid = (y < p.h) && (x < p.w) && (t < p.num_tiles) && !(p.dbg & 1);
        if (p.dbg & 2) continue;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) +
                               static_cast<uint32_t>((buf * 2 + acc) * p.npad);
        float rn = 1.0f;
        if (p.pixel_norm) {
          float ssq = 0.0f;
          for (int c0 = cbeg; c0 < cend; c0 += 16) {
            float v[16];
            epi_chunk16(taddr + c0, &s_shift[c0], act_a, act_b, act_tanh, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) ssq = fmaf(v[j], v[j], ssq);
          }
          if (epi_active == 8) {  // add the partner warp's half (same lane quarter, other column half)
            s_pn[acc][warp >= 8 ? 1 : 0][m] = ssq;
            quarter_pair_sync(ew);
            ssq += s_pn[acc][warp >= 8 ? 0 : 1][m];
          }
          rn = rsqrtf(ssq * inv_c + 1e-8f);  // tools_wscale/GAN.py:472-474
        }
        if (p.tma_store) {
          // 16-bit outputs with whole 64-channel groups: the warp stages its 32 pixels x 64 channels (4 image rows x
          // 8 px, 128 B per pixel) in its own 4 KB of shared memory in the 128B-swizzled box layout and ONE lane
          // issues a bulk tensor store. Full 128-byte lines reach L2 instead of 32 scattered 32-byte sectors per
          // store instruction (the 32->128 layer was bound by those: epilogue alone 0.28 ms vs 0.24 ms of MMAs).
          // No block barrier is involved (the earlier staged variant needed two per tile and lost).
          const int od = p.out_dtype;
          float sacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int cg = cbeg; cg < cend; cg += 64) {
            if (lane == 0) tma_store_wait_read();  // the previous store of this warp has finished reading the staging
            __syncwarp();
            // (one TMEM load in flight at a time: holding all four raised the kernel to 141 registers, which costs the
            //  two-CTAs-per-SM layers their occupancy and measured slower here as well)
#pragma unroll 1
            for (int q4 = 0; q4 < 4; ++q4) {
              float v[16];
              epi_chunk16(taddr + cg + q4 * 16, &s_shift[cg + q4 * 16], act_a, act_b, act_tanh, v);
              if (p.pixel_norm) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] *= rn;
              }
              if (SIDE && p.side) {  // 1x1 shortcut of the next block: 16 channels x 8 outputs from registers, weights broadcast
                const float4* wq = s_sw + (cg + q4 * 16) * 2;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float4 w0 = wq[2 * j], w1 = wq[2 * j + 1];
                  sacc[0] = fmaf(v[j], w0.x, sacc[0]);
                  sacc[1] = fmaf(v[j], w0.y, sacc[1]);
                  sacc[2] = fmaf(v[j], w0.z, sacc[2]);
                  sacc[3] = fmaf(v[j], w0.w, sacc[3]);
                  sacc[4] = fmaf(v[j], w1.x, sacc[4]);
                  sacc[5] = fmaf(v[j], w1.y, sacc[5]);
                  sacc[6] = fmaf(v[j], w1.z, sacc[6]);
                  sacc[7] = fmaf(v[j], w1.w, sacc[7]);
                }
              }
              uint4 lo, hi;
              lo.x = pack_h16x2(v[0], v[1], od);
              lo.y = pack_h16x2(v[2], v[3], od);
              lo.z = pack_h16x2(v[4], v[5], od);
              lo.w = pack_h16x2(v[6], v[7], od);
              hi.x = pack_h16x2(v[8], v[9], od);
              hi.y = pack_h16x2(v[10], v[11], od);
              hi.z = pack_h16x2(v[12], v[13], od);
              hi.w = pack_h16x2(v[14], v[15], od);
              const uint32_t sw = static_cast<uint32_t>(lane & 7);
              *reinterpret_cast<uint4*>(stg + lane * 128 + (((2u * q4) ^ sw) << 4)) = lo;
              *reinterpret_cast<uint4*>(stg + lane * 128 + (((2u * q4 + 1u) ^ sw) << 4)) = hi;
            }
            fence_proxy_async();
            __syncwarp();
            if (SIDE && p.side && cg + 64 >= cend) {
              // fold the two column halves (warps w and w+4 of this lane quarter) and store 8 fp32 per pixel
              if (epi_active == 8) {
                if (warp >= 8) {
                  s_sx[2 * m] = make_float4(sacc[0], sacc[1], sacc[2], sacc[3]);
                  s_sx[2 * m + 1] = make_float4(sacc[4], sacc[5], sacc[6], sacc[7]);
                }
                quarter_pair_sync(ew);
                if (warp < 8) {
                  const float4 a = s_sx[2 * m], b = s_sx[2 * m + 1];
                  sacc[0] += a.x; sacc[1] += a.y; sacc[2] += a.z; sacc[3] += a.w;
                  sacc[4] += b.x; sacc[5] += b.y; sacc[6] += b.z; sacc[7] += b.w;
                }
                quarter_pair_sync(ew);  // the exchange buffer is free for the next accumulator
              }
              if (warp < 8 && valid) {
                float4* so = reinterpret_cast<float4*>(p.side_out + ((static_cast<size_t>(tc.n) * p.h + y) * p.w + x) * 8);
                so[0] = make_float4(sacc[0], sacc[1], sacc[2], sacc[3]);
                so[1] = make_float4(sacc[4], sacc[5], sacc[6], sacc[7]);
              }
            }
            if (lane == 0 && t < p.num_tiles && !(p.dbg & 1)) {
              if (ups == 1) {
                tma_store_4d(&tm_y, stg, cg, tc.x0 + acc * 8, tc.y0 + ew * 4, tc.n);
              } else {  // nearest x2: the four replicas of the block, map dims (c, ux, x, uy, n*h + y)
                const int row = tc.n * p.h + tc.y0 + ew * 4;
                tma_store_5d(&tm_y, stg, cg, 0, tc.x0 + acc * 8, 0, row);
                tma_store_5d(&tm_y, stg, cg, 1, tc.x0 + acc * 8, 0, row);
                tma_store_5d(&tm_y, stg, cg, 0, tc.x0 + acc * 8, 1, row);
                tma_store_5d(&tm_y, stg, cg, 1, tc.x0 + acc * 8, 1, row);
              }
              tma_store_commit();
            }
          }
          continue;
        }
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
          float v[16];
          epi_chunk16(taddr + c0, &s_shift[c0], act_a, act_b, act_tanh, v);
          if (p.pixel_norm) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= rn;
          }
          if (valid) {
            for (int uy = 0; uy < ups; ++uy) {
              for (int ux = 0; ux < ups; ++ux) {
                const size_t pix = (static_cast<size_t>(tc.n) * oh + (y * ups + uy)) * ow + (x * ups + ux);
                if (p.out_dtype != MPG_F32) {
                  const int od = p.out_dtype;
                  uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + pix * p.out_cstride + c0;
                  if (st32 && c0 + 16 <= p.out_cstride) {
                    // one 32-byte store = one full L2 sector per lane (two 16-byte halves cost two partial writes)
                    st_global_v8(o, pack_h16x2(v[0], v[1], od), pack_h16x2(v[2], v[3], od), pack_h16x2(v[4], v[5], od),
                                 pack_h16x2(v[6], v[7], od), pack_h16x2(v[8], v[9], od), pack_h16x2(v[10], v[11], od),
                                 pack_h16x2(v[12], v[13], od), pack_h16x2(v[14], v[15], od));
                  } else {
                    if (c0 + 8 <= p.out_cstride) {
                      uint4 q;
                      q.x = pack_h16x2(v[0], v[1], od);
                      q.y = pack_h16x2(v[2], v[3], od);
                      q.z = pack_h16x2(v[4], v[5], od);
                      q.w = pack_h16x2(v[6], v[7], od);
                      *reinterpret_cast<uint4*>(o) = q;
                    }
                    if (c0 + 16 <= p.out_cstride) {
                      uint4 q;
                      q.x = pack_h16x2(v[8], v[9], od);
                      q.y = pack_h16x2(v[10], v[11], od);
                      q.z = pack_h16x2(v[12], v[13], od);
                      q.w = pack_h16x2(v[14], v[15], od);
                      *reinterpret_cast<uint4*>(o + 8) = q;
                    }
                  }
                } else {
                  float* o = reinterpret_cast<float*>(p.out) + pix * p.out_cstride + c0;
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (c0 + j < p.out_cstride) o[j] = (c0 + j < p.cout) ? v[j] : 0.0f;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[buf]), 0));
        else mbar_arrive(&tmem_empty[buf]);
      }
    }
    if (p.tma_store && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves early
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
#undef MPG_TILE_LOOP
#undef MPG_TILE_CLAMP
}

}  // namespace

typedef void (*IgKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const IgemmParams);

static IgKernel ig_kernel(int ck, int pair, int side = 0) {
  // SIDE (epilogue side output, mpg_conv_plan_set_side) is its own instantiation: its 8 extra accumulators and weight
  // registers must not cost the two-CTAs-per-SM layers their occupancy
  if (side) return conv_igemm_kernel<64, true, true>;
  if (pair) return ck == 64 ? conv_igemm_kernel<64, true, false> : (ck == 32 ? conv_igemm_kernel<32, true, false> : conv_igemm_kernel<16, true, false>);
  return ck == 64 ? conv_igemm_kernel<64, false, false> : (ck == 32 ? conv_igemm_kernel<32, false, false> : conv_igemm_kernel<16, false, false>);
}

// The attribute is per kernel instantiation AND per device (cudaFuncSetAttribute applies to the current device only):
// only ever raise it, and remember what each device has.
static size_t g_smem_attr[kMaxDevices][7] = {};

int igemm_set_smem_attr(int device, int ck, int pair, size_t smem_bytes, int side) {
  const int slot = side ? 6 : (ck == 64 ? 0 : (ck == 32 ? 1 : 2)) + (pair ? 3 : 0);
  const bool cached = device >= 0 && device < kMaxDevices;
  if (cached && smem_bytes <= g_smem_attr[device][slot]) return 0;
  DeviceGuard guard(device);
  cudaError_t e = cudaFuncSetAttribute(ig_kernel(ck, pair, side), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem_bytes));
  if (e == cudaSuccess && cached) g_smem_attr[device][slot] = smem_bytes;
  return static_cast<int>(e);
}

int igemm_launch(int ck, const CUtensorMap& tm_x0, const CUtensorMap& tm_x1, const CUtensorMap& tm_w,
                 const CUtensorMap& tm_y, const IgemmParams& p, int grid, size_t smem_bytes, cudaStream_t stream) {
  if (!p.pair) {
    ig_kernel(ck, 0)<<<grid, p.threads, smem_bytes, stream>>>(tm_x0, tm_x1, tm_w, tm_y, p);
    return static_cast<int>(cudaGetLastError());
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(static_cast<unsigned>(p.threads), 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, ig_kernel(ck, 1, p.side), tm_x0, tm_x1, tm_w, tm_y, p));
}

}  // namespace mpg
