// Tap-folded tcgen05 convolution for narrow Cout (design: conv_nfold.cuh).
#include <string.h>

#include "conv_nfold.cuh"
#include "ptx.cuh"

namespace mpg {

namespace {

struct NfTile {
  int n, y0, gx0, ox0;  // gx0: global x of lane 0 of the window; ox0: first output x of the tile
};
__device__ __forceinline__ NfTile nf_decode(int t, const NfoldParams& p) {
  const int per_img = p.tiles_x * p.tiles_y;
  NfTile c;
  c.n = t / per_img;
  const int r = t - c.n * per_img;
  const int ty = r / p.tiles_x;
  c.y0 = ty * p.rows;
  c.ox0 = (r - ty * p.tiles_x) * p.valid_w;
  c.gx0 = c.ox0 - (p.seg_ks[0] >> 1);
  return c;
}

// PAIR (cta_group::2, cluster of 2; deep-K layers such as 5x5 128->32): each CTA keeps HALF of every weight tile
// resident in shared memory for its whole lifetime (the full set does not fit next to the window images), the leader
// issues M=256 MMAs over both CTAs' windows. With no weight traffic left a tile is ONE accumulator, so three of
// them rotate through TMEM (3 x 160 columns) and the shuffle-sum epilogue of tile i overlaps the MMAs of i+1, i+2 --
// the single-CTA configuration had to hold 3 accumulators per weight pass and could not overlap its epilogue.
template <int CK, int KS, bool PAIR>
__global__ void __maxnreg__(128)
conv_nfold_kernel(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1,
                  const __grid_constant__ CUtensorMap tm_w, const NfoldParams p) {
  // CK == 8 ("NS8"): thin inputs (<= 8 channels, pixel stride 16 B). The window image is staged UN-swizzled with
  // whole image rows as the TMA inner dimension (one TMA row per image row instead of one per pixel: the tiled
  // TMA path costs ~7 cycles per box row whatever its size), which is exactly the no-swizzle K-major core-matrix
  // layout: 8 consecutive pixels x 8 channels = one 8x16-byte core matrix. The second K half of a K=16 MMA is the
  // SAME pixels one image row below (LBO = row pitch), so one MMA covers two vertical taps.
  constexpr bool NS8 = (CK == 8);
  constexpr int RB = CK * 2;  // bytes per pixel of one K-chunk (== swizzle span when swizzled)
  constexpr uint32_t LAYOUT = NS8 ? 0u : ((RB == 128) ? 2u : (RB == 64 ? 4u : 6u));
  constexpr uint32_t SBO = 8u * RB;
  constexpr int KSTEPS = NS8 ? 1 : CK / 16;
  constexpr uint32_t ROW_BYTES = kNfWin * RB;  // one image row of the window in smem

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_a[kNfMaxStagesA], empty_a[kNfMaxStagesA];
  __shared__ __align__(8) uint64_t full_b[kNfMaxStagesB], empty_b[kNfMaxStagesB];
  __shared__ __align__(8) uint64_t tmem_full[kNfMaxBufs], tmem_empty[kNfMaxBufs];
  __shared__ __align__(8) uint64_t b_ready;  // PAIR: both CTAs' resident weights have landed (leader's copy is used)
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_shift[32];
  __shared__ float s_pn[2][2][128];  // pixel_norm partial sums: [slot][chunk parity][accumulator row]

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + static_cast<size_t>(p.na) * p.a_stage_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // tile schedule: CTA q takes tiles q, q+G, ...; a PAIR takes tiles (2q, 2q+1) and both CTAs run the same number
  // of iterations (an odd last tile is recomputed by the peer, not stored)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int tstep = PAIR ? static_cast<int>(gridDim.x >> 1) * 2 : static_cast<int>(gridDim.x);
  const int tfirst = PAIR ? static_cast<int>(blockIdx.x >> 1) * 2 + static_cast<int>(rank) : static_cast<int>(blockIdx.x);
#define NF_TILE_LOOP(t) for (int t = tfirst; t - static_cast<int>(rank) < p.num_tiles; t += tstep)
#define NF_TILE_CLAMP(t) ((t) < p.num_tiles ? (t) : p.num_tiles - 1)

  if (threadIdx.x < 32) s_shift[threadIdx.x] = threadIdx.x < p.cp ? p.shift[threadIdx.x] : 0.0f;
  // 4 epilogue warps (256-thread launch, two CTAs per SM) or 8 (384 threads: two warps per TMEM lane quarter take
  // the even / odd 8-channel chunks; for pixel_norm they exchange their partial sums of squares through s_pn)
  const int n_extra = p.a_cpasync ? p.nprod - 2 : 0;  // extra cp.async producer warps sit after the epilogue warps
  const int n_epi_launched = static_cast<int>(blockDim.x >> 5) - 4 - n_extra;
  const int epi_active = (n_epi_launched == 8 && p.cp > 8) ? 8 : 4;
  const int first_extra = 4 + n_epi_launched;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x0);
    if (p.nseg > 1) tma_prefetch_desc(&tm_x1);
    for (int i = 0; i < p.na; ++i) {
      // cp.async staging: one arrival per producer lane (2 warps), plus the peer CTA's relay in PAIR mode (leader only)
      mbar_init(&full_a[i], p.a_cpasync ? 32u * static_cast<uint32_t>(p.nprod) + ((PAIR && rank == 0) ? 1u : 0u) : 1u);
      mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < kNfMaxStagesB; ++i) {
      mbar_init(&full_b[i], 1);
      mbar_init(&empty_b[i], 1);
    }
    for (int i = 0; i < kNfMaxBufs; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], (PAIR ? 2 : 1) * epi_active);  // one arrive per active epilogue warp (both CTAs in PAIR mode)
    }
    mbar_init(&b_ready, 2);
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

  // ===================== A producer, cp.async form (p.a_cpasync; swizzled layouts, resident weights) ================
  // Tiled TMA costs ~5.5 cycles per box row (one pixel of one K-chunk, <= 128 B) per SM whatever its size (measured:
  // the load-only skeleton of the 5x5 128->32 layer took 0.198 ms with 128-byte rows and 0.358 ms with 64-byte rows),
  // which bounded that layer below its MMA time. Here warps 0 and 3 copy the window with 16-byte cp.async (LDGSTS,
  // 512 B per warp instruction) straight into the 128B/64B/32B-swizzled image the UMMA descriptors expect; out-of-image
  // pixels and channels beyond the tensor are zero-filled (= SAME padding / TMA out-of-bounds fill). Every lane's copies
  // arrive on the stage's mbarrier asynchronously (cp.async.mbarrier.arrive.noinc); in PAIR mode the peer CTA collects
  // its own lanes on its local barrier and warp 2 relays one cluster-scope arrival to the leader.
  if (p.a_cpasync && !NS8 && (warp == 0 || warp == 3 || warp >= first_extra)) {
    if (warp == 3 && lane == 0) {  // resident weights first (same as the TMA form below)
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpacked);
      const uint32_t tile_bytes = static_cast<uint32_t>(p.b_tile_bytes);
      if (PAIR) {
        mbar_arrive_expect_tx(&full_b[0], tile_bytes * static_cast<uint32_t>(p.ktiles));
        for (int kt = 0; kt < p.ktiles; ++kt)
          bulk_load_1d(smB + static_cast<size_t>(kt) * tile_bytes, wsrc + (static_cast<size_t>(kt) * 2 + rank) * tile_bytes,
                       tile_bytes, &full_b[0]);
        mbar_wait(&full_b[0], 0);
        mbar_arrive_cluster(mapa_u32(smem_u32(&b_ready), 0));
      } else {
        mbar_arrive_expect_tx(&full_b[0], tile_bytes * static_cast<uint32_t>(p.ktiles));
        bulk_load_1d(smB, wsrc, tile_bytes * static_cast<uint32_t>(p.ktiles), &full_b[0]);
      }
    }
    __syncwarp();
    constexpr int CPR = RB / 16;    // 16-byte chunks per pixel row of a K-chunk
    constexpr int PPI = 32 / CPR;   // pixels per warp instruction
    // producer warp index -> (column group cgp in {0,1}, row group rg): a warp copies its column group of every
    // (nprod/2)-th window row
    const int pidx = warp == 0 ? 0 : (warp == 3 ? 1 : 2 + (warp - first_extra));
    const int pw = pidx & 1, rg = pidx >> 1, nrg = p.nprod >> 1;
    const int j = lane % CPR, q = lane / CPR;
    const uint32_t smA_u32 = smem_u32(smA);
    int st = 0;
    uint32_t ph = 0;
    NF_TILE_LOOP(t) {
      const NfTile tc = nf_decode(NF_TILE_CLAMP(t), p);
      for (int s = 0; s < p.nseg; ++s) {
        const int ks = p.seg_ks[s];
        const int wy0 = tc.y0 - (ks >> 1);
        const int cin = p.seg_cin[s];
        const size_t cs2 = static_cast<size_t>(p.seg_cstride[s]) * 2;
        const uint8_t* img = static_cast<const uint8_t*>(p.x[s]) + static_cast<size_t>(tc.n) * p.h * p.w * cs2;
        for (int ch = 0; ch < p.seg_nchunk[s]; ++ch) {
          if (lane == 0) mbar_wait(&empty_a[st], ph ^ 1u);
          __syncwarp();
          const uint32_t stage = smA_u32 + static_cast<uint32_t>(st) * static_cast<uint32_t>(p.a_stage_bytes);
          const int c0 = ch * CK + j * 8;  // first channel of this lane's 16 bytes
          const uint32_t cbytes = c0 >= cin ? 0u : ((cin - c0) >= 8 ? 16u : static_cast<uint32_t>((cin - c0) * 2));
          // a lane owns NCOL fixed window columns (wx = pw*PPI + q + 2*PPI*k): per image row ONE 64-bit address, then NCOL
          // copies at constant offsets (a per-pixel recomputation made every copy a ~150-cycle dependent chain)
          constexpr int NCOL = kNfWin / (2 * PPI) > 0 ? kNfWin / (2 * PPI) : 1;  // (>= 1 only to let the unused NS8 instance compile)
          const int wx0 = pw * PPI + q;
          uint32_t dst_col[NCOL];
          uint32_t ok_col = 0;
#pragma unroll
          for (int k = 0; k < NCOL; ++k) {
            const int wx = wx0 + 2 * PPI * k;
            const uint32_t sw = RB == 128 ? (wx & 7) : (RB == 64 ? ((wx >> 1) & 3) : ((wx >> 2) & 1));
            dst_col[k] = stage + static_cast<uint32_t>(wx) * RB + ((static_cast<uint32_t>(j) ^ sw) << 4);
            const int gx = tc.gx0 + wx;
            ok_col |= (gx >= 0 && gx < p.w && cbytes != 0u) ? (1u << k) : 0u;
          }
          const uint8_t* col0 = img + (static_cast<long long>(tc.gx0) + wx0) * static_cast<long long>(cs2) +
                                static_cast<size_t>(c0 < cin ? c0 : 0) * 2;
          const int nrows = p.rows + ks - 1;
#pragma unroll 2
          for (int wy = rg; wy < nrows; wy += nrg) {
            const int gy = wy0 + wy;
            const bool row_ok = gy >= 0 && gy < p.h;
            const uint8_t* rowp = col0 + static_cast<long long>(row_ok ? gy : 0) * p.w * static_cast<long long>(cs2);
            const uint32_t drow = static_cast<uint32_t>(wy) * (kNfWin * RB);
#pragma unroll
            for (int k = 0; k < NCOL; ++k) {
              const bool ok = row_ok && ((ok_col >> k) & 1u);
              cp_async_16_zfill(dst_col[k] + drow, ok ? rowp + static_cast<size_t>(2 * PPI * k) * cs2 : img, ok ? cbytes : 0u);
            }
          }
          cp_async_mbar_arrive_noinc(&full_a[st]);
          if (++st == p.na) {
            st = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (p.a_cpasync && !NS8 && PAIR && rank != 0 && warp == 2) {
    // peer CTA: relay "this CTA's window has landed" to the leader's stage barrier
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      NF_TILE_LOOP(t) {
        for (int s = 0; s < p.nseg; ++s)
          for (int ch = 0; ch < p.seg_nchunk[s]; ++ch) {
            mbar_wait(&full_a[st], ph);
            fence_proxy_async_all();
            mbar_arrive_cluster(mapa_u32(smem_u32(&full_a[st]), 0));
            if (++st == p.na) {
              st = 0;
              ph ^= 1u;
            }
          }
      }
    }
  } else if (warp == 0) {
    // ===================== A producer: one window image per (segment, chunk) =====================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      NF_TILE_LOOP(t) {
        const NfTile tc = nf_decode(NF_TILE_CLAMP(t), p);
        for (int s = 0; s < p.nseg; ++s) {
          const int ks = p.seg_ks[s];
          // NS8 loads one more row: the unused second half of the last tap pair must read finite data
          const uint32_t bytes = static_cast<uint32_t>(p.rows + ks - (NS8 ? 0 : 1)) * ROW_BYTES;
          const CUtensorMap* tm = (s == 0) ? &tm_x0 : &tm_x1;
          for (int ch = 0; ch < p.seg_nchunk[s]; ++ch) {
            mbar_wait(&empty_a[st], ph ^ 1u);
            if (PAIR) {  // both CTAs' windows complete on the LEADER's barrier, which expects the bytes of both
              if (rank == 0) mbar_arrive_expect_tx(&full_a[st], 2u * bytes);
              tma_load_4d_2cta(smA + static_cast<size_t>(st) * p.a_stage_bytes, tm, mapa_u32(smem_u32(&full_a[st]), 0),
                               ch * CK, tc.gx0, tc.y0 - (ks >> 1), tc.n);
            } else {
            mbar_arrive_expect_tx(&full_a[st], bytes);
            if (NS8)  // tensor viewed as [N][H][W*8]: inner coordinate in elements
              tma_load_3d(smA + static_cast<size_t>(st) * p.a_stage_bytes, tm, &full_a[st], tc.gx0 * 8, tc.y0 - (ks >> 1), tc.n);
            else
              tma_load_4d(smA + static_cast<size_t>(st) * p.a_stage_bytes, tm, &full_a[st], ch * CK, tc.gx0,
                          tc.y0 - (ks >> 1), tc.n);
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
    // ===================== B producer: weight tiles [Npad][CK] per (segment, chunk, dy) ============
    if (lane == 0) {
      // weight tiles live in global memory in their swizzled smem image: plain bulk copies (see ptx.cuh)
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpacked);
      const uint32_t tile_bytes = static_cast<uint32_t>(p.b_tile_bytes);
      if (PAIR) {
        // resident halves: k-tile kt of this CTA = rows [rank*N/2, (rank+1)*N/2) of the full tile (whole swizzle atoms)
        mbar_arrive_expect_tx(&full_b[0], tile_bytes * static_cast<uint32_t>(p.ktiles));
        for (int kt = 0; kt < p.ktiles; ++kt)
          bulk_load_1d(smB + static_cast<size_t>(kt) * tile_bytes,
                       wsrc + (static_cast<size_t>(kt) * 2 + rank) * tile_bytes, tile_bytes, &full_b[0]);
        mbar_wait(&full_b[0], 0);
        mbar_arrive_cluster(mapa_u32(smem_u32(&b_ready), 0));
      } else if (p.bres) {
        mbar_arrive_expect_tx(&full_b[0], tile_bytes * static_cast<uint32_t>(p.ktiles));
        bulk_load_1d(smB, wsrc, tile_bytes * static_cast<uint32_t>(p.ktiles), &full_b[0]);
      } else {
        int st = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
          for (int kt = 0; kt < p.ktiles; ++kt) {
            mbar_wait(&empty_b[st], ph ^ 1u);
            mbar_arrive_expect_tx(&full_b[st], tile_bytes);
            bulk_load_1d(smB + static_cast<size_t>(st) * p.b_tile_bytes, wsrc + static_cast<size_t>(kt) * p.b_tile_bytes,
                         tile_bytes, &full_b[st]);
            if (++st == p.nb) {
              st = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1 && (!PAIR || rank == 0)) {
    // ===================== MMA issuer (warp-uniform loop, one lane elected once issues; PAIR: leader CTA only) =====
    // lean 32-bit descriptor arithmetic only (see conv_igemm.cu: the issue loop, not the tensor pipe, bounded CK<=32 layers)
    const uint32_t idesc = umma_idesc_f16kind(PAIR ? 256 : 128, p.npad, p.in_dtype == MPG_F16 ? 0u : 1u);
    constexpr uint32_t DESC_HI = (SBO >> 4) | (1u << 14) | (LAYOUT << 29);
    // leading-dimension byte offset (distance between the two K halves): ignored when swizzled; NS8: A = one image
    // row, B = the npad/8 core matrices of the first K half
    constexpr uint32_t DESC_LO = NS8 ? ((ROW_BYTES >> 4) << 16) : (1u << 16);
    constexpr uint32_t ROW16 = ROW_BYTES >> 4;
    const uint32_t b_desc_lo = NS8 ? ((static_cast<uint32_t>(p.npad) * 16u) >> 4) << 16 : (1u << 16);
    const uint32_t smA_lo = ((smem_u32(smA) & 0x3FFFFu) >> 4) | DESC_LO;
    const uint32_t smB_lo = ((smem_u32(smB) & 0x3FFFFu) >> 4) | b_desc_lo;
    const uint32_t a_stage16 = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
    const uint32_t b_tile16 = static_cast<uint32_t>(p.b_tile_bytes) >> 4;
    const uint32_t npad = static_cast<uint32_t>(p.npad);
    const int naccs = p.naccs, nbuf = p.nbuf, na = p.na, nb = p.nb, nseg = p.nseg;
    const bool bres = p.bres != 0;
    const bool do_mma = !(p.dbg & 4);
    const bool leader = elect_one() != 0;
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int buf = 0;
    uint32_t use = 0;
    if (PAIR) {
      mbar_wait_cluster(&b_ready, 0);
      tc_fence_after();
    } else if (bres) {
      mbar_wait(&full_b[0], 0);
      tc_fence_after();
    }
    NF_TILE_LOOP(t) {
      mbar_wait(&tmem_empty[buf], (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t dbase = tmem_base + static_cast<uint32_t>(buf * naccs) * npad;
      uint32_t b_res = smB_lo;
      uint32_t accumulate = 0;
      for (int s = 0; s < nseg; ++s) {
        const int ks = p.seg_ks[s];
        const int nchunk = p.seg_nchunk[s];
        for (int ch = 0; ch < nchunk; ++ch) {
          if (PAIR) mbar_wait_cluster(&full_a[sa], pa);  // (the peer's relay arrives with release.cluster)
          else mbar_wait(&full_a[sa], pa);
          if (p.a_cpasync) fence_proxy_async_all();  // cp.async wrote through the generic proxy, the MMA reads through the async one
          tc_fence_after();
          uint32_t a_row = smA_lo + static_cast<uint32_t>(sa) * a_stage16;  // image row dy of the window
          for (int dy = 0; dy < ks; dy += (NS8 ? 2 : 1)) {
            uint32_t b_lo;
            if (bres) {
              b_lo = b_res;
              b_res += b_tile16;
            } else {
              mbar_wait(&full_b[sb], pb);
              tc_fence_after();
              b_lo = smB_lo + static_cast<uint32_t>(sb) * b_tile16;
            }
            if (leader) {
              if (do_mma) {
                uint32_t ag = a_row, d = dbase;
                for (int acc = 0; acc < naccs; ++acc) {
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k) {
                    const uint64_t bd = (static_cast<uint64_t>(DESC_HI) << 32) | (b_lo + k * 2);
                    const uint64_t ad = (static_cast<uint64_t>(DESC_HI) << 32) | (ag + k * 2);
                    if (PAIR) umma_bf16_ss_2cta(d, ad, bd, idesc, (k > 0) ? 1u : accumulate);
                    else umma_bf16_ss(d, ad, bd, idesc, (k > 0) ? 1u : accumulate);
                  }
                  ag += static_cast<uint32_t>(kNfRowsAcc) * ROW16;
                  d += npad;
                }
              }
              if (!bres) umma_commit(&empty_b[sb]);
            }
            __syncwarp();
            accumulate = 1;
            a_row += (NS8 ? 2u : 1u) * ROW16;
            if (!bres && ++sb == nb) {
              sb = 0;
              pb ^= 1u;
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
      if (++buf == nbuf) {
        buf = 0;
        ++use;
      }
    }
  } else if (warp >= 4 && warp < 4 + epi_active) {
    // ===================== epilogue: shifted sum over dx (warp shuffles) + shift + act + store ====
    const int ew = warp & 3;  // TMEM lane quarter == image row within the accumulator
    constexpr int pad0 = KS >> 1;
    const bool lane_valid = (lane >= pad0) && (lane < kNfWin - pad0);
    const float act_a = p.act == MPG_ACT_RELU ? 0.5f : (p.act == MPG_ACT_LRELU ? 0.6f : 1.0f);
    const float act_b = p.act == MPG_ACT_RELU ? 0.5f : (p.act == MPG_ACT_LRELU ? 0.4f : 0.0f);
    const bool is_relu = p.act == MPG_ACT_RELU;
    const float inv_c = 1.0f / static_cast<float>(p.cout);
    const int nchunks = p.cp >> 3;
    const int cgrp = (warp - 4) >> 2;           // chunk parity this warp handles when 8 warps are active
    const bool all_chunks = epi_active == 4;
    int pn_slot = 0;
    int it = 0;
    for (int t = tfirst; t - static_cast<int>(rank) < p.num_tiles; t += tstep, ++it) {
      const int buf = it % p.nbuf;
      const uint32_t use = static_cast<uint32_t>(it / p.nbuf);
      const NfTile tc = nf_decode(NF_TILE_CLAMP(t), p);
      mbar_wait(&tmem_full[buf], use & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int acc = 0; acc < p.naccs; ++acc) {
        if (p.dbg & 2) continue;
        const int y = tc.y0 + acc * kNfRowsAcc + ew;
        const int gx = tc.gx0 + lane;
        const bool valid = lane_valid && (y < p.h) && (gx < p.w) && (t < p.num_tiles) && !(p.dbg & 1);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) +
                               static_cast<uint32_t>((buf * p.naccs + acc) * p.npad);
        float o[32];
        float ssq = 0.0f;
        float rs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (p.resid != nullptr && valid) {
          const float4* rp = reinterpret_cast<const float4*>(p.resid + ((static_cast<size_t>(tc.n) * p.h + y) * p.w + gx) * 8);
          const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
          rs[0] = r0.x; rs[1] = r0.y; rs[2] = r0.z; rs[3] = r0.w;
          rs[4] = r1.x; rs[5] = r1.y; rs[6] = r1.z; rs[7] = r1.w;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nchunks && (all_chunks || (c & 1) == cgrp)) {
            uint32_t r[KS][8];
#pragma unroll
            for (int dx = 0; dx < KS; ++dx) tmem_ld8(taddr + static_cast<uint32_t>(dx * p.cp + c * 8), r[dx]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float t[KS];
#pragma unroll
              for (int dx = 0; dx < KS; ++dx)  // independent shuffles: lane x takes column block dx of lane x+dx-pad
                t[dx] = __shfl_sync(0xFFFFFFFFu, __uint_as_float(r[dx][j]), (lane + dx - pad0) & 31);
              float a = t[0];
#pragma unroll
              for (int dx = 1; dx < KS; ++dx) a += t[dx];
              const float x = a + s_shift[c * 8 + j] + (c == 0 ? rs[j] : 0.0f);
              const float v = is_relu ? fmaxf(x, 0.0f) : fmaf(act_b, fabsf(x), act_a * x);
              o[c * 8 + j] = v;
              ssq = fmaf(v, v, ssq);
            }
          }
        }
        if (p.pixel_norm && epi_active == 8) {  // add the partner warp's chunks (same image row, other chunk parity)
          s_pn[pn_slot][cgrp][ew * 32 + lane] = ssq;
          quarter_pair_sync(ew);
          ssq += s_pn[pn_slot][cgrp ^ 1][ew * 32 + lane];
          pn_slot ^= 1;
        }
        const float rn = p.pixel_norm ? rsqrtf(ssq * inv_c + 1e-8f) : 1.0f;  // tools_wscale/GAN.py:472-474
        if (valid) {
          const size_t pix = (static_cast<size_t>(tc.n) * p.h + y) * p.w + gx;
          if (p.out_dtype != MPG_F32) {
            const int od = p.out_dtype;
            uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + pix * p.out_cstride;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c < nchunks && (all_chunks || (c & 1) == cgrp)) {
                uint4 q;
                q.x = pack_h16x2(o[c * 8 + 0] * rn, o[c * 8 + 1] * rn, od);
                q.y = pack_h16x2(o[c * 8 + 2] * rn, o[c * 8 + 3] * rn, od);
                q.z = pack_h16x2(o[c * 8 + 4] * rn, o[c * 8 + 5] * rn, od);
                q.w = pack_h16x2(o[c * 8 + 6] * rn, o[c * 8 + 7] * rn, od);
                *reinterpret_cast<uint4*>(op + c * 8) = q;
              }
            }
          } else {
            float* op = reinterpret_cast<float*>(p.out) + pix * p.out_cstride;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < p.out_cstride && (j >> 3) < nchunks && (all_chunks || ((j >> 3) & 1) == cgrp))
                op[j] = (j < p.cout) ? o[j] * rn : 0.0f;
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
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves early
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
#undef NF_TILE_LOOP
#undef NF_TILE_CLAMP
}

}  // namespace

typedef void (*NfKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const NfoldParams);

static NfKernel nf_kernel(int ck, int ks, int pair) {
  if (pair) {
    if (ks == 5) return ck == 64 ? conv_nfold_kernel<64, 5, true> : conv_nfold_kernel<32, 5, true>;
    return ck == 64 ? conv_nfold_kernel<64, 3, true> : conv_nfold_kernel<32, 3, true>;
  }
  if (ck == 8) return ks == 5 ? conv_nfold_kernel<8, 5, false> : conv_nfold_kernel<8, 3, false>;
  if (ks == 5)
    return ck == 64 ? conv_nfold_kernel<64, 5, false> : (ck == 32 ? conv_nfold_kernel<32, 5, false> : conv_nfold_kernel<16, 5, false>);
  return ck == 64 ? conv_nfold_kernel<64, 3, false> : (ck == 32 ? conv_nfold_kernel<32, 3, false> : conv_nfold_kernel<16, 3, false>);
}

static size_t g_nf_smem_attr[kMaxDevices][12] = {};  // per device: cudaFuncSetAttribute applies to the current device only

int nfold_set_smem_attr(int device, int ck, int ks, int pair, size_t smem_bytes) {
  const int slot = pair ? 8 + (ck == 64 ? 0 : 1) + (ks == 5 ? 0 : 2)
                        : ((ck == 8) ? (ks == 5 ? 6 : 7) : (ck == 64 ? 0 : (ck == 32 ? 1 : 2)) + (ks == 5 ? 0 : 3));
  const bool cached = device >= 0 && device < kMaxDevices;
  if (cached && smem_bytes <= g_nf_smem_attr[device][slot]) return 0;
  DeviceGuard guard(device);
  cudaError_t e = cudaFuncSetAttribute(nf_kernel(ck, ks, pair), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem_bytes));
  if (e == cudaSuccess && cached) g_nf_smem_attr[device][slot] = smem_bytes;
  return static_cast<int>(e);
}

int nfold_launch(int ck, const CUtensorMap& tm_x0, const CUtensorMap& tm_x1, const CUtensorMap& tm_w,
                 const NfoldParams& p, int grid, size_t smem_bytes, cudaStream_t stream) {
  if (!p.pair) {
    nf_kernel(ck, p.seg_ks[0], 0)<<<grid, p.threads, smem_bytes, stream>>>(tm_x0, tm_x1, tm_w, p);
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
  return static_cast<int>(cudaLaunchKernelEx(&cfg, nf_kernel(ck, p.seg_ks[0], 1), tm_x0, tm_x1, tm_w, p));
}

}  // namespace mpg
