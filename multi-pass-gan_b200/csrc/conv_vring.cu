// Row-streaming tcgen05 convolution, vertical-tap sum accumulated in a TMEM ring (design: conv_vring.cuh).
#include <string.h>

#include "conv_vring.cuh"
#include "ptx.cuh"

namespace mpg {

namespace {

// A CTA walks its contiguous range of flattened (image, strip, row) units as a sequence of segments: `len` output rows
// y0.. of image n, strip starting at pixel x0.
struct VrIter {
  int cur, end;
  int n, x0, y0, len;
  // units_x: strips (single CTA) or strip pairs (CTA pair: the CTA of cluster rank r takes strip 2*unit + r) per image row
  __device__ __forceinline__ bool next(const VringParams& p, int per_unit, int rank) {
    if (cur >= end) return false;
    const int u = cur / p.h;
    y0 = cur - u * p.h;
    len = min(p.h - y0, end - cur);
    n = u / p.units_x;
    x0 = ((u - n * p.units_x) * per_unit + rank) * kVrStrip;
    cur += len;
    return true;
  }
};

// NCHW: 8-column chunks of a slot an epilogue warp owns; G: epilogue warps per TMEM lane quarter; PAIR: cta_group::2 pairs
// (adjacent strips, HALF of every weight tile resident per CTA, M = 256 MMAs issued by the leader).
template <int CK, int KS, int NCHW, int G, bool PAIR>
__global__ void __launch_bounds__(128 + 128 * G, 1)
conv_vring_kernel(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1,
                  const VringParams p) {
  constexpr int RB = CK * 2;  // bytes per pixel of one staged K-chunk == swizzle span of the A operand
  constexpr uint32_t A_LAYOUT = (RB == 128) ? 2u : 4u;
  constexpr uint32_t A_HI = ((8u * RB) >> 4) | (1u << 14) | (A_LAYOUT << 29);
  constexpr uint32_t B_HI = (256u >> 4) | (1u << 14) | (6u << 29);  // weights: 32-byte rows, SWIZZLE_32B, 8-row groups 256 B apart
  constexpr int KSTEPS = CK / 16;
  constexpr int PAD = KS >> 1;
  constexpr int WIN = kVrStrip + KS - 1;
  constexpr uint32_t PX16 = RB >> 4;
  constexpr uint32_t TMEM_COLS = 512;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_a[kVrMaxStagesA], empty_a[kVrMaxStagesA];
  __shared__ __align__(8) uint64_t slot_full[kVrMaxSlots], slot_free[kVrMaxSlots];
  __shared__ __align__(8) uint64_t full_b, b_ready, seg_flushed;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_shift[64];
  __shared__ float s_pn[2][G][128];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint8_t* smB = smem;  // resident weights first (b_bytes is a 1024-byte multiple)
  uint8_t* smA = smem + p.b_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  constexpr int PER_UNIT = PAIR ? 2 : 1;
  const int irank = static_cast<int>(rank);
  const int row_begin = static_cast<int>(PAIR ? blockIdx.x >> 1 : blockIdx.x) * p.rows_per_cta;
  const int row_end = min(row_begin + p.rows_per_cta, p.total_rows);
  const int R = p.nslots;
  const int cs = p.cs;

  if (threadIdx.x < 64) s_shift[threadIdx.x] = threadIdx.x < p.cp ? p.shift[threadIdx.x] : 0.0f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x0);
    if (p.nseg > 1) tma_prefetch_desc(&tm_x1);
    for (int i = 0; i < p.na; ++i) {
      mbar_init(&full_a[i], 1);
      mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < R; ++i) {
      mbar_init(&slot_full[i], 1);
      mbar_init(&slot_free[i], (PAIR ? 2 : 1) * 4 * G);  // one arrive per epilogue warp (of both CTAs)
    }
    mbar_init(&full_b, 1);
    mbar_init(&b_ready, 2);
    mbar_init(&seg_flushed, (PAIR ? 2 : 1) * 4 * G);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc2(&tmem_base_slot, TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(&tmem_base_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // tensor memory is undefined after allocation and every MMA accumulates: clear it once
  if (warp >= 4 && warp < 8) {
    const uint32_t la = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (uint32_t c = 0; c < TMEM_COLS; c += 8) tmem_st8_fill(la + c, 0u);
    tmem_st_wait();
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // (also: the peer's barriers must be initialised before anything signals them)
  else __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ===================== A producer: one staged image row per (segment, Cin chunk) =====================
    if (lane == 0) {
      constexpr uint32_t bytes = static_cast<uint32_t>(WIN) * RB;
      int st = 0;
      uint32_t ph = 0;
      VrIter it{row_begin, row_end, 0, 0, 0, 0};
      while (it.next(p, PER_UNIT, irank)) {
        const int nrows = it.len + KS - 1;
        for (int wr = 0; wr < nrows; ++wr) {
          const int gy = it.y0 - PAD + wr;
          for (int s = 0; s < p.nseg; ++s) {
            const CUtensorMap* tm = (s == 0) ? &tm_x0 : &tm_x1;
            for (int ch = 0; ch < p.seg_nchunk[s]; ++ch) {
              mbar_wait(&empty_a[st], ph ^ 1u);
              if (PAIR) {  // both CTAs' rows complete on the LEADER's barrier, which expects the bytes of both
                if (rank == 0) mbar_arrive_expect_tx(&full_a[st], 2u * bytes);
                tma_load_4d_2cta(smA + static_cast<size_t>(st) * p.a_stage_bytes, tm, mapa_u32(smem_u32(&full_a[st]), 0),
                                 ch * CK, it.x0 - PAD, gy, it.n);
              } else {
                mbar_arrive_expect_tx(&full_a[st], bytes);
                tma_load_4d(smA + static_cast<size_t>(st) * p.a_stage_bytes, tm, &full_a[st], ch * CK, it.x0 - PAD, gy, it.n);
              }
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
    // ===================== resident weights, loaded once =====================
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpacked) + static_cast<size_t>(rank) * p.b_bytes;
      mbar_arrive_expect_tx(&full_b, static_cast<uint32_t>(p.b_bytes));
      for (int off = 0; off < p.b_bytes; off += 32768) {
        const int nb = min(32768, p.b_bytes - off);
        bulk_load_1d(smB + off, wsrc + off, static_cast<uint32_t>(nb), &full_b);
      }
      if (PAIR) {
        mbar_wait(&full_b, 0);
        mbar_arrive_cluster(mapa_u32(smem_u32(&b_ready), 0));
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues; PAIR: leader CTA only) ===========
    const uint32_t fmt = p.in_dtype == MPG_F16 ? 0u : 1u;
    const uint32_t idesc = umma_idesc_f16kind(PAIR ? 256 : 128, KS * cs, fmt);
    const uint32_t idesc_sc = umma_idesc_f16kind(PAIR ? 256 : 128, cs, fmt);
    const uint32_t smA_lo = ((smem_u32(smA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t smB_lo = ((smem_u32(smB) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a_stage16 = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
    const uint32_t b_tile16 = static_cast<uint32_t>(p.b_tile_bytes) >> 4;
    const uint32_t b_sc_tile16 = static_cast<uint32_t>(p.b_sc_tile_bytes) >> 4;
    const int na = p.na, nseg = p.nseg;
    const int nch0 = p.seg_nchunk[0], kl0 = p.seg_klast[0];
    const int nch1 = p.seg_nchunk[1], kl1 = p.seg_klast[1];
    const bool do_mma = !(p.dbg & 4);
    const bool leader = elect_one() != 0;
    int sa = 0;
    uint32_t pa = 0;
    // Ring positions are a function of the IMAGE ROW (slot of a row's first column block = row mod R), not of a running
    // counter: which contributions of an output row land in an overflow slot -- i.e. the order its fp32 partial sums are
    // added in -- then depends on the row alone, and the result is bit-identical however the rows are cut into ranges
    // (batch size, number of GPUs). The price: a range leaves k-1 partially written slots at positions unrelated to the
    // next range, so the epilogue clears them and the next range waits for that (seg_flushed).
    int b0 = 0;
    uint32_t need_wait = 0;  // bit s: slot s was handed to the MMAs and its output row is (or will be) read by the epilogue
    uint32_t free_par = 0;   // bit s: parity of the next completion of slot_free[s] to wait for
    uint32_t garbage = 0;    // slots the previous range left partially written (cleared by the epilogue's flush)
    int seg = 0;
    if (PAIR) mbar_wait_cluster(&b_ready, 0);
    else mbar_wait(&full_b, 0);
    tc_fence_after();
    VrIter it{row_begin, row_end, 0, 0, 0, 0};
    while (it.next(p, PER_UNIT, irank)) {
      const int nrows = it.len + KS - 1;
      for (int wr = 0; wr < nrows; ++wr) {
        int first_new = KS - 1;  // column blocks [first_new, KS) touch slots this row is the first to use
        if (wr == 0) {
          if (seg > 0) {
            if (PAIR) mbar_wait_cluster(&seg_flushed, static_cast<uint32_t>(seg - 1) & 1u);
            else mbar_wait(&seg_flushed, static_cast<uint32_t>(seg - 1) & 1u);
            tc_fence_after();
            need_wait &= ~garbage;
          }
          ++seg;
          b0 = it.y0 % R;
          first_new = 0;
        }
        for (int j = first_new; j < KS; ++j) {
          int sn = b0 + j;
          if (sn >= R) sn -= R;
          const uint32_t bit = 1u << sn;
          if (need_wait & bit) {  // the epilogue has read and cleared the slot's previous output row
            if (PAIR) mbar_wait_cluster(&slot_free[sn], (free_par >> sn) & 1u);
            else mbar_wait(&slot_free[sn], (free_par >> sn) & 1u);
            free_par ^= bit;
          }
          need_wait |= bit;
        }
        tc_fence_after();
        // the k column blocks go to the PHYSICAL slots b0 .. b0+k-1: past the end of the ring they land in the k-1
        // overflow slots, which alias ring slots 0 .. k-2 (the epilogue adds the two halves), so an MMA never splits
        const uint32_t d = tmem_base + static_cast<uint32_t>(b0 * cs);
        uint32_t b_cur = smB_lo;
        for (int ch = 0; ch < nch0; ++ch) {
          if (PAIR) mbar_wait_cluster(&full_a[sa], pa);
          else mbar_wait(&full_a[sa], pa);
          tc_fence_after();
          const uint32_t a_row = smA_lo + static_cast<uint32_t>(sa) * a_stage16;
          const int nk = (ch == nch0 - 1) ? kl0 : KSTEPS;
          if (leader && do_mma) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              if (k < nk) {
#pragma unroll
                for (int dx = 0; dx < KS; ++dx) {
                  const uint64_t ad = (static_cast<uint64_t>(A_HI) << 32) | (a_row + static_cast<uint32_t>(dx) * PX16 + k * 2);
                  const uint64_t bd = (static_cast<uint64_t>(B_HI) << 32) | (b_cur + static_cast<uint32_t>(k * KS + dx) * b_tile16);
                  if (PAIR) umma_bf16_ss_2cta(d, ad, bd, idesc, 1u);
                  else umma_bf16_ss(d, ad, bd, idesc, 1u);
                }
              }
            }
          }
          b_cur += static_cast<uint32_t>(nk * KS) * b_tile16;
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
        if (nseg > 1) {
          // 1x1 shortcut: centre pixel shift, onto the slot of the output row at this image row (column block PAD)
          const uint32_t ds = d + static_cast<uint32_t>(PAD * cs);
          for (int ch = 0; ch < nch1; ++ch) {
            if (PAIR) mbar_wait_cluster(&full_a[sa], pa);
            else mbar_wait(&full_a[sa], pa);
            tc_fence_after();
            const uint32_t a_tap = smA_lo + static_cast<uint32_t>(sa) * a_stage16 + static_cast<uint32_t>(PAD) * PX16;
            const int nk = (ch == nch1 - 1) ? kl1 : KSTEPS;
            if (leader && do_mma) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                if (k < nk) {
                  const uint64_t ad = (static_cast<uint64_t>(A_HI) << 32) | (a_tap + k * 2);
                  const uint64_t bd = (static_cast<uint64_t>(B_HI) << 32) | (b_cur + static_cast<uint32_t>(k) * b_sc_tile16);
                  if (PAIR) umma_bf16_ss_2cta(ds, ad, bd, idesc_sc, 1u);
                  else umma_bf16_ss(ds, ad, bd, idesc_sc, 1u);
                }
              }
            }
            b_cur += static_cast<uint32_t>(nk) * b_sc_tile16;
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
        if (leader) {  // slot b0 has received its last tap
          if (PAIR) umma_commit_2cta(&slot_full[b0], 3);
          else umma_commit(&slot_full[b0]);
        }
        __syncwarp();
        if (++b0 == R) b0 = 0;
      }
      garbage = 0;  // slots b0 .. b0+k-2 (b0 already advanced) hold partial sums of rows below the range
      for (int j = 0; j < KS - 1; ++j) {
        int sg = b0 + j;
        if (sg >= R) sg -= R;
        garbage |= 1u << sg;
      }
    }
  } else if (warp >= 4 && warp < 4 + 4 * G) {
    // ===================== epilogue: read one finished slot, clear it, shift + act + store =====================
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int nchunks = p.cp >> 3;
    const float act_a = p.act == MPG_ACT_LRELU ? 0.6f : 1.0f;
    const float act_b = p.act == MPG_ACT_LRELU ? 0.4f : 0.0f;
    const bool is_relu = p.act == MPG_ACT_RELU;
    const bool is_tanh = p.act == MPG_ACT_TANH;
    const float inv_c = 1.0f / static_cast<float>(p.cout);
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    int pn_slot = 0;
    int slot = 0;
    uint32_t full_par = 0;  // bit s: parity of the next completion of slot_full[s]
    VrIter it{row_begin, row_end, 0, 0, 0, 0};
    while (it.next(p, PER_UNIT, irank)) {
      const int gx = it.x0 + ew * 32 + lane;
      const bool col_ok = gx < p.w && !(p.dbg & 1);
      const int nrows = it.len + KS - 1;
      slot = it.y0 % R;  // ring position = image row mod R (see the MMA warp)
#pragma unroll 1
      for (int wr = 0; wr < nrows; ++wr) {
        mbar_wait(&slot_full[slot], (full_par >> slot) & 1u);
        full_par ^= 1u << slot;
        tc_fence_after();
        const uint32_t taddr = lane_addr + static_cast<uint32_t>(slot * cs + grp * NCHW * 8);
        const uint32_t talias = taddr + static_cast<uint32_t>(R * cs);  // overflow slot R + slot (exists for slot < KS-1)
        const bool alias = slot < KS - 1;
        const int orow = wr - (KS - 1);  // output row (relative to y0) held by this slot; < 0: rows above the range
        float o[NCHW * 8];
        if (orow >= 0) {
          // all loads of the row in flight at once, ONE wait (a wait per chunk serialised NCHW TMEM round trips per row)
          uint32_t r[NCHW][8], r2[NCHW][8];
#pragma unroll
          for (int c = 0; c < NCHW; ++c) {
            if (grp * NCHW + c < nchunks) {
              tmem_ld8(taddr + static_cast<uint32_t>(c * 8), r[c]);
              if (alias) tmem_ld8(talias + static_cast<uint32_t>(c * 8), r2[c]);
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < NCHW; ++c) {
            if (grp * NCHW + c < nchunks) {
#pragma unroll
              for (int j = 0; j < 8; ++j) o[c * 8 + j] = __uint_as_float(r[c][j]) + (alias ? __uint_as_float(r2[c][j]) : 0.0f);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < NCHW; ++c)
          if (grp * NCHW + c < nchunks) {
            tmem_st8_fill(taddr + static_cast<uint32_t>(c * 8), 0u);
            if (alias) tmem_st8_fill(talias + static_cast<uint32_t>(c * 8), 0u);
          }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&slot_free[slot]), 0));
          else mbar_arrive(&slot_free[slot]);
        }
        const bool last_row = wr == nrows - 1;
        if (++slot == R) slot = 0;
        if (last_row) {
          // the k-1 slots after the last output row hold partial sums of rows below the range: clear them (the MMAs that
          // wrote them completed before slot_full of this row) and tell the MMA warp the ring is clean
          for (int j = 0; j < KS - 1; ++j) {
            int sg = slot + j;
            if (sg >= R) sg -= R;
            const uint32_t tg = lane_addr + static_cast<uint32_t>(sg * cs + grp * NCHW * 8);
#pragma unroll
            for (int c = 0; c < NCHW; ++c)
              if (grp * NCHW + c < nchunks) {
                tmem_st8_fill(tg + static_cast<uint32_t>(c * 8), 0u);
                if (sg < KS - 1) tmem_st8_fill(tg + static_cast<uint32_t>(R * cs + c * 8), 0u);
              }
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&seg_flushed), 0));
            else mbar_arrive(&seg_flushed);
          }
        }
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
  if (PAIR) cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves early
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef void (*VrKernel)(const CUtensorMap, const CUtensorMap, const VringParams);

template <int CK, int KS, bool PAIR>
VrKernel vr_kernel_cfg(int nchw, int g) {
  if (g == 1) {
    switch (nchw) {
      case 1: return conv_vring_kernel<CK, KS, 1, 1, PAIR>;
      default: return conv_vring_kernel<CK, KS, 2, 1, PAIR>;
    }
  }
  if (g == 3) return conv_vring_kernel<CK, KS, 2, 3, PAIR>;
  if (g == 4) {
    switch (nchw) {
      case 1: return conv_vring_kernel<CK, KS, 1, 4, PAIR>;
      default: return conv_vring_kernel<CK, KS, 2, 4, PAIR>;
    }
  }
  switch (nchw) {
    case 1: return conv_vring_kernel<CK, KS, 1, 2, PAIR>;
    case 2: return conv_vring_kernel<CK, KS, 2, 2, PAIR>;
    case 3: return conv_vring_kernel<CK, KS, 3, 2, PAIR>;
    default: return conv_vring_kernel<CK, KS, 4, 2, PAIR>;
  }
}

VrKernel vr_kernel(int ck, int ks, int nchw, int g, int pair) {
  if (pair) {
    if (ck == 64) return ks == 5 ? vr_kernel_cfg<64, 5, true>(nchw, g) : vr_kernel_cfg<64, 3, true>(nchw, g);
    return ks == 5 ? vr_kernel_cfg<32, 5, true>(nchw, g) : vr_kernel_cfg<32, 3, true>(nchw, g);
  }
  if (ck == 64) return ks == 5 ? vr_kernel_cfg<64, 5, false>(nchw, g) : vr_kernel_cfg<64, 3, false>(nchw, g);
  return ks == 5 ? vr_kernel_cfg<32, 5, false>(nchw, g) : vr_kernel_cfg<32, 3, false>(nchw, g);
}

}  // namespace

static size_t g_vr_smem_attr[kMaxDevices][128] = {};  // per device: cudaFuncSetAttribute applies to the current device only

int vring_set_smem_attr(int device, int ck, int ks, int nchw, int g, int pair, size_t smem_bytes) {
  const int slot = (pair ? 64 : 0) + (ck == 64 ? 0 : 32) + (ks == 5 ? 0 : 16) + (g - 1) * 4 + (nchw - 1);
  const bool cached = device >= 0 && device < kMaxDevices;
  if (cached && smem_bytes <= g_vr_smem_attr[device][slot]) return 0;
  DeviceGuard guard(device);
  cudaError_t e = cudaFuncSetAttribute(vr_kernel(ck, ks, nchw, g, pair), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem_bytes));
  if (e == cudaSuccess && cached) g_vr_smem_attr[device][slot] = smem_bytes;
  return static_cast<int>(e);
}

int vring_launch(int ck, int nchw, const CUtensorMap& tm_x0, const CUtensorMap& tm_x1, const VringParams& p, int grid,
                 size_t smem_bytes, cudaStream_t stream) {
  const unsigned threads = 128u + 128u * static_cast<unsigned>(p.epi_groups);
  if (!p.pair) {
    vr_kernel(ck, p.ks, nchw, p.epi_groups, 0)<<<grid, threads, smem_bytes, stream>>>(tm_x0, tm_x1, p);
    return static_cast<int>(cudaGetLastError());
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, vr_kernel(ck, p.ks, nchw, p.epi_groups, 1), tm_x0, tm_x1, p));
}

}  // namespace mpg
