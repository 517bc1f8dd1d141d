// Convolution plans: argument validation, weight packing, kernel selection and launch.
// C-ABI: mpg_conv_plan_* (include/mpg.h).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.h"
#include "conv_direct.cuh"
#include "conv_igemm.cuh"
#include "conv_nfold.cuh"
#include "conv_tiny.cuh"
#include "conv_vfold.cuh"
#include "conv_vring.cuh"

struct mpg_conv_plan_s {
  mpg_handle h;
  mpg_conv_desc d;
  int kind;  // 1 igemm, 2 direct, 3 tap-folded igemm (narrow Cout), 4 CUDA-core tiny (Cout <= 2, Cin <= 8),
             // 5 row-streaming igemm with the vertical taps folded into N (conv_vfold.cu)
  mpg::NfoldParams np;
  mpg::VfoldParams vp;
  int vf_nchw;
  mpg::VringParams rp;
  double flops;
  int oh, ow;
  // ---- igemm
  int ck, npad;
  int seg_nchunk[2];
  void* d_wpacked;
  float* d_shift;
  CUtensorMap tm_w;
  CUtensorMap tm_x[2];
  const void* tm_x_ptr[2];
  CUtensorMap tm_y;
  const void* tm_y_ptr;
  mpg::IgemmParams ip;
  size_t smem_bytes;
  int grid;
  float* d_side_w;  // igemm side output (mpg_conv_plan_set_side): [128][8] fp32
  // ---- direct
  float* d_wdirect;
  mpg::DirectParams dp;
  // ---- tiny
  mpg::TinyParams tp;
};

namespace {

using namespace mpg;

uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40u);  // NaN
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return static_cast<uint16_t>(u >> 16);
}

uint16_t f32_to_f16_rn(float f) {
  __half hv = __float2half_rn(f);
  uint16_t r;
  memcpy(&r, &hv, 2);
  return r;
}

CUtensorMapSwizzle swizzle_for(int ck) {
  return ck == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                  : (ck == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// Element offset of (row n, K-element c) inside one weight tile stored as the SWIZZLED shared-memory image
// the UMMA descriptor expects (rows of rb = 2*ck bytes; the TMA/UMMA swizzle XORs the 16-byte chunk index
// with address bits [7,10) / [7,9) / [7,8) for 128 / 64 / 32-byte rows; tiles are 1024-byte aligned).
size_t swz_elem(int n, int c, int ck) {
  const int rb = ck * 2;
  const int chunk = (c * 2) >> 4;
  const int x = rb == 128 ? (n & 7) : (rb == 64 ? ((n >> 1) & 3) : ((n >> 2) & 1));
  return (static_cast<size_t>(n) * rb + static_cast<size_t>((chunk ^ x) << 4) + ((c * 2) & 15)) / 2;
}

// Device-side re-pack of fp32 HWIO weights into the igemm tile layout [ktile][npad][ck] (ktile order
// (segment, chunk, dx, dy)), used by the training step where the weights change every optimizer step.
// mode 0: W(dy,dx,ci,n) = w[dy][dx][ci][n]                      (forward; w is [k,k,cin,cout])
// mode 1: W(dy,dx,ci,n) = w[k-1-dy][k-1-dx][n][ci]              (dgrad; w is the FORWARD tensor [k,k,cout_d,cin_d])
struct RepackArgs {
  const float* w[2];
  int mode[2];
  int ks[2], cin[2], nchunk[2], kt0[2];
  int src_k;     // segment 0: kernel size of the SOURCE tensor (== ks[0], or smaller: embedded at offset ks[0] - src_k, zero taps before)
  int src_cout;  // segment 0, mode 0: cout of the source tensor (this plan covers its channels [cout_off, cout_off + cout))
  int cout_off;
  int nseg, ck, npad, cout, f16;
  uint16_t* out;
  long long total;
};
__global__ void __launch_bounds__(256) repack_igemm_kernel(const RepackArgs a) {
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < a.total; e += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(e % a.ck);
    long long r = e / a.ck;
    const int n = static_cast<int>(r % a.npad);
    const int kt = static_cast<int>(r / a.npad);
    const int s = (a.nseg > 1 && kt >= a.kt0[1]) ? 1 : 0;
    const int ks = a.ks[s];
    int q = kt - a.kt0[s];
    const int dy = q % ks;
    q /= ks;
    const int dx = q % ks;
    const int ch = q / ks;
    const int ci = ch * a.ck + c;
    float v = 0.0f;
    if (ci < a.cin[s] && n < a.cout) {
      // source taps: a k' x k' tensor embedded in the k x k kernel at offset o = k - k' (the TF SAME window of an even kernel,
      // e.g. 4x4 -> taps -1..+2 of a 5x5); segment 1 is never embedded
      const int kf = s == 0 ? a.src_k : ks, o = ks - kf;
      const int sc = (s == 0 && a.mode[s] == 0) ? a.src_cout : a.cout, co = (s == 0 && a.mode[s] == 0) ? a.cout_off : 0;
      if (a.mode[s] == 0) {
        if (dy >= o && dx >= o) v = a.w[s][((static_cast<size_t>(dy - o) * kf + (dx - o)) * a.cin[s] + ci) * sc + co + n];
      } else {
        const int fy = ks - 1 - dy - o, fx = ks - 1 - dx - o;
        if (fy >= 0 && fx >= 0) v = a.w[s][((static_cast<size_t>(fy) * kf + fx) * a.cout + n) * a.cin[s] + ci];
      }
    }
    a.out[e] = a.f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

int same_pad_before(int in, int k, int s) {
  const int out = (in + s - 1) / s;
  int total = (out - 1) * s + k - in;
  if (total < 0) total = 0;
  return total / 2;  // TF: pad_before = floor(pad_total / 2)
}

bool igemm_eligible(const mpg_conv_desc& d) {
  if (!is_h16(d.in_dtype) || d.stride != 1 || d.in_upsample != 1) return false;
  if (d.out_dtype != MPG_F32 && d.out_dtype != d.in_dtype) return false;
  if (d.cout > 128) return false;
  if (d.upsample != 1 && d.upsample != 2) return false;
  for (int s = 0; s < d.nseg; ++s) {
    const int k = d.seg_ksize[s];
    if (!(k == 1 || k == 3 || k == 5)) return false;
    if (d.seg_cstride[s] % 8 != 0) return false;
    if (d.seg_cin[s] > d.seg_cstride[s]) return false;
  }
  if (d.out_dtype != MPG_F32 && d.out_cstride % 8 != 0) return false;
  return true;
}

int build_igemm(mpg_conv_plan p, const float* w[2], const float* scale[2], const float* shift) {
  const mpg_conv_desc& d = p->d;
  int maxcin = 0;
  for (int s = 0; s < d.nseg; ++s) maxcin = d.seg_cin[s] > maxcin ? d.seg_cin[s] : maxcin;
  // K-chunk (= swizzle span): zero-padded channels cost real MMAs, so take the chunk with the fewest
  // K=16 steps summed over the segments, preferring the larger one unless a smaller saves >= 20 %
  auto ksteps_for = [&](int c) {
    int t = 0;
    for (int s = 0; s < d.nseg; ++s) t += d.seg_ksize[s] * d.seg_ksize[s] * ceil_div(d.seg_cin[s], c) * (c / 16);
    return t;
  };
  int ck = maxcin > 32 ? 64 : (maxcin > 16 ? 32 : 16);
  for (int c = ck / 2; c >= 16; c /= 2)
    if (ksteps_for(c) * 5 <= ksteps_for(ck) * 4) ck = c;
  if (const char* e = getenv("MPG_IGEMM_CK")) {
    const int c = atoi(e);
    if (c == 16 || c == 32 || c == 64) ck = c;
  }
  const int rb = ck * 2;
  const int npad = round_up(d.cout, 16);
  p->ck = ck;
  p->npad = npad;
  int ktiles = 0;
  int maxks = 1;
  for (int s = 0; s < d.nseg; ++s) {
    p->seg_nchunk[s] = ceil_div(d.seg_cin[s], ck);
    ktiles += p->seg_nchunk[s] * d.seg_ksize[s] * d.seg_ksize[s];
    maxks = d.seg_ksize[s] > maxks ? d.seg_ksize[s] : maxks;
  }
  // ---- pack weights: [ktile][npad][ck] 16-bit, ktile order = (seg, chunk, dx, dy); loaded by tiled TMA (measured:
  //      for the streamed 16 KB tiles of the wide layers the tiled path beats 1-D bulk copies, 1.02 vs 1.39 ms skeleton)
  std::vector<uint16_t> wp(static_cast<size_t>(ktiles) * npad * ck, 0);
  size_t kt = 0;
  for (int s = 0; s < d.nseg; ++s) {
    const int ks = d.seg_ksize[s], cin = d.seg_cin[s];
    for (int ch = 0; ch < p->seg_nchunk[s]; ++ch)
      for (int dx = 0; dx < ks; ++dx)
        for (int dy = 0; dy < ks; ++dy, ++kt)
          for (int n = 0; n < d.cout; ++n) {
            const float sc = scale[s] ? scale[s][n] : 1.0f;
            for (int c = 0; c < ck; ++c) {
              const int ci = ch * ck + c;
              if (ci >= cin) break;
              const float v = w[s][((static_cast<size_t>(dy) * ks + dx) * cin + ci) * d.cout + n] * sc;
              wp[(kt * npad + n) * ck + c] = d.in_dtype == MPG_F16 ? f32_to_f16_rn(v) : f32_to_bf16_rn(v);
            }
          }
  }
  MPG_CUDA(cudaMalloc(&p->d_wpacked, wp.size() * 2));
  MPG_CUDA(cudaMemcpy(p->d_wpacked, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
  std::vector<float> sh(npad, 0.0f);
  if (shift)
    for (int n = 0; n < d.cout; ++n) sh[n] = shift[n];
  MPG_CUDA(cudaMalloc(&p->d_shift, npad * sizeof(float)));
  MPG_CUDA(cudaMemcpy(p->d_shift, sh.data(), npad * sizeof(float), cudaMemcpyHostToDevice));


  IgemmParams& ip = p->ip;
  memset(&ip, 0, sizeof(ip));
  ip.n = d.n;
  ip.h = d.h;
  ip.w = d.w;
  ip.tiles_x = ceil_div(d.w, kIgTileW);
  ip.tiles_y = ceil_div(d.h, kIgTileH);
  ip.num_tiles = ip.tiles_x * ip.tiles_y * d.n;
  ip.nseg = d.nseg;
  for (int s = 0; s < d.nseg; ++s) {
    ip.seg_ks[s] = d.seg_ksize[s];
    ip.seg_nchunk[s] = p->seg_nchunk[s];
  }
  ip.npad = npad;
  ip.cout = d.cout;
  ip.act = d.act;
  ip.pixel_norm = d.pixel_norm;
  ip.upsample = d.upsample;
  ip.in_dtype = d.in_dtype;
  ip.out_dtype = d.out_dtype;
  ip.out_cstride = d.out_cstride;
  // one TMA halo image per (segment, chunk); every (dy,dx) tap is a shifted UMMA descriptor into it. (The earlier
  // per-dx-image staging and the block-barrier TMA-store epilogue both measured slower on B200 and were removed.)
  ip.tma_store = 0;
  ip.a_stage_bytes = (kIgTileH + maxks - 1) * (kIgTileW + maxks - 1) * rb;
  ip.a_stage_bytes = round_up(ip.a_stage_bytes, 1024);
  ip.ktiles = ktiles;
  // CTA pairs (cta_group::2) for the layers that stream their weights: each CTA keeps half of every weight tile
  ip.pair = (npad % 32 == 0 && ip.tiles_x * ip.tiles_y * d.n >= 2 &&
             static_cast<size_t>(ktiles) * round_up(npad * rb, 1024) > 64 * 1024) ? 1 : 0;
  if (const char* e = getenv("MPG_IGEMM_PAIR")) ip.pair = (atoi(e) && npad % 32 == 0) ? 1 : 0;
  const int brows = ip.pair ? npad / 2 : npad;
  ip.b_tile_bytes = round_up(brows * rb, 1024);
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ck), static_cast<uint64_t>(ktiles) * npad};
    const uint64_t strides[1] = {static_cast<uint64_t>(rb)};
    const uint32_t box[2] = {static_cast<uint32_t>(ck), static_cast<uint32_t>(brows)};
    int r = encode_tmap(p->h, &p->tm_w, d.in_dtype == MPG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        p->d_wpacked, dims, strides, box, swizzle_for(ck));
    if (r) return r;
  }
  // resident weights: thin layers keep every weight tile in shared memory for the CTA's lifetime, which
  // removes the per-tile weight TMA round trips that bound them (measured: 0.18 of 0.22 ms was load skeleton)
  ip.bres = (!ip.pair && static_cast<size_t>(ktiles) * ip.b_tile_bytes <= 64 * 1024) ? 1 : 0;
  // CTA pairs: a CTA only holds half of every tile, so e.g. the 5x5 32->128 layer (25 x 4 KB halves) stays
  // resident next to three halo images and the MMA thread never waits on a weight stage
  if (ip.pair && static_cast<size_t>(ktiles) * ip.b_tile_bytes + 3 * static_cast<size_t>(ip.a_stage_bytes) <= 200 * 1024) ip.bres = 1;
  if (const char* e = getenv("MPG_IGEMM_BRES"))
    ip.bres = (atoi(e) && static_cast<size_t>(ktiles) * ip.b_tile_bytes + 2 * static_cast<size_t>(ip.a_stage_bytes) <= 200 * 1024) ? 1 : 0;
  // group the ks vertical taps of one (chunk,dx) into a single B stage when that stays small: every stage
  // hand-over (mbarrier wait + tcgen05.commit round trip) costs ~450 cycles on top of bytes/27 B/clk (measured,
  // tools/thin_probe.py), so the MMA thread should wait/commit once per 2*ks*CK/16 MMAs, not once per 2*CK/16
  ip.bgroup = (maxks * ip.b_tile_bytes <= 40 * 1024) ? 1 : 0;
  if (const char* e = getenv("MPG_IGEMM_BGROUP")) ip.bgroup = atoi(e) ? 1 : 0;
  ip.b_stage_bytes = ip.b_tile_bytes * (ip.bgroup ? maxks : 1);
  ip.stage_bytes = 0;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(4 * npad)) cols <<= 1;
  // CTAs per SM: thin layers are latency bound per tile (TMA round trips), not throughput bound, so two
  // small CTAs per SM hide it (registers cap it at 2); wide layers keep one CTA with deep rings
  const int b_res_bytes = ip.bres ? ktiles * ip.b_tile_bytes : 0;
  int occ = 1;
  if (cols <= 256 && (ip.bres ? b_res_bytes : 3 * ip.b_stage_bytes) + 2 * ip.a_stage_bytes <= 100 * 1024) occ = 2;
  if (const char* e = getenv("MPG_IGEMM_OCC")) occ = atoi(e) > 0 ? atoi(e) : 1;
  if (ip.pair) occ = 1;
  // 8 epilogue warps when the SM holds one CTA (two 256-thread CTAs need the registers)
  ip.threads = (occ == 1) ? kIgMaxThreads : kIgThreads;
  if (const char* e = getenv("MPG_IGEMM_THREADS")) ip.threads = atoi(e) == 384 ? 384 : 256;
  const int budget = (210 * 1024) / occ - (occ > 1 ? 2048 : 0);
  int nb, na;
  if (ip.bres) {
    nb = 1;
    ip.b_stage_bytes = b_res_bytes;
  } else {
    // Ring depths: one B stage only carries 2*CK/16 MMAs (~0.1-0.35 us of tensor work) but a TMA
    // round trip is ~1 us, so small weight tiles need a deep ring; A stages carry ks times more work.
    nb = (96 * 1024 / occ) / ip.b_stage_bytes;
    nb = nb > kIgMaxStagesB ? kIgMaxStagesB : (nb < 4 ? 4 : nb);
    if (ip.bgroup && nb > 4) nb = 4;
    while (nb > 2 && nb * ip.b_stage_bytes + 2 * ip.a_stage_bytes > budget) --nb;
    if (const char* e = getenv("MPG_IGEMM_NB")) nb = atoi(e);
  }
  na = (budget - nb * ip.b_stage_bytes) / ip.a_stage_bytes;
  na = na > kIgMaxStagesA ? kIgMaxStagesA : na;
  if (const char* e = getenv("MPG_IGEMM_NA")) na = atoi(e);
  if (nb < 1 || nb > kIgMaxStagesB || na < 2 || na > kIgMaxStagesA ||
      static_cast<size_t>(na) * ip.a_stage_bytes + static_cast<size_t>(nb) * ip.b_stage_bytes > 224 * 1024 ||
      cols * static_cast<uint32_t>(occ) > 512u) {
    set_error("conv: bad pipeline depth na=%d nb=%d occ=%d (a_stage %d B, b_stage %d B)", na, nb, occ, ip.a_stage_bytes,
              ip.b_stage_bytes);
    return MPG_EINVAL;
  }
  // per-warp staged TMA-store epilogue (conv_igemm.cu): 16-bit outputs made of whole 64-channel groups; needs 4 KB of
  // staging per epilogue warp next to the rings (one A stage is given up for it when that keeps >= 2)
  {
    const int epi_warps = (ip.threads == kIgMaxThreads) ? 8 : 4;
    // nearest x2 replicated store (upsample 2): the same staged block goes out as FOUR bulk stores through a 5-D map
    // (c, ux, x, uy, n*h + y); rows of different images are adjacent in that last dimension, so the image height must be
    // a whole number of tiles. (The per-lane replicated 32-byte stores it replaces ran at 0.24 TB/s: 2.2 ms for the
    // 128->128 layer in front of the first x2 of the 8x generator, 14x its siblings.)
    bool ts = is_h16(d.out_dtype) && (d.upsample == 1 || (d.upsample == 2 && d.h % kIgTileH == 0)) && d.cout % 64 == 0 &&
              d.out_cstride == d.cout && occ == 1 && (epi_warps == 4 || npad % 128 == 0);
    if (const char* e = getenv("MPG_IGEMM_TMASTORE")) ts = ts && atoi(e) != 0;
    const size_t staging = static_cast<size_t>(epi_warps) * 4096;
    if (ts) {
      while (na > 2 && static_cast<size_t>(na) * ip.a_stage_bytes + static_cast<size_t>(nb) * ip.b_stage_bytes + staging > 223 * 1024) --na;
      ts = static_cast<size_t>(na) * ip.a_stage_bytes + static_cast<size_t>(nb) * ip.b_stage_bytes + staging <= 223 * 1024;
    }
    ip.tma_store = ts ? 1 : 0;
    ip.stage_bytes = ts ? static_cast<int>(staging) : 0;
  }
  ip.nb = nb;
  ip.na = na;
  if (const char* e = getenv("MPG_IGEMM_DBG")) ip.dbg = atoi(e);
  ip.tmem_cols = cols;
  ip.shift = p->d_shift;
  ip.stage_off = ip.na * ip.a_stage_bytes + ip.nb * ip.b_stage_bytes;
  p->smem_bytes = static_cast<size_t>(ip.stage_off) + static_cast<size_t>(ip.stage_bytes) + 1024;
  p->grid = ip.num_tiles < p->h->sm_count * occ ? ip.num_tiles : p->h->sm_count * occ;
  if (const char* e = getenv("MPG_IGEMM_GRID")) p->grid = atoi(e) > 0 ? atoi(e) : p->grid;
  if (ip.pair) {  // whole pairs only; a pair covers two tiles
    const int pairs_needed = (ip.num_tiles + 1) / 2;
    int pairs = p->h->sm_count / 2;
    if (pairs > pairs_needed) pairs = pairs_needed;
    p->grid = pairs * 2;
  }
  int r = igemm_set_smem_attr(p->h->device, ck, ip.pair, p->smem_bytes);
  if (r) {
    set_error("cudaFuncSetAttribute(max dynamic smem %zu) failed: %s", p->smem_bytes,
              cudaGetErrorString(static_cast<cudaError_t>(r)));
    return r;
  }
  p->tm_x_ptr[0] = p->tm_x_ptr[1] = nullptr;
  p->tm_y_ptr = nullptr;
  return 0;
}

bool nfold_eligible(const mpg_conv_desc& d) {
  if (!igemm_eligible(d) || d.upsample != 1 || d.cout > 32 || d.act == MPG_ACT_TANH) return false;
  if (d.seg_ksize[0] != 3 && d.seg_ksize[0] != 5) return false;
  if (d.nseg == 2 && d.seg_ksize[1] != 1) return false;
  const int cp = round_up(d.cout, 8);
  if (d.out_dtype == MPG_F32 ? d.out_cstride > 32 : d.out_cstride != cp) return false;
  return true;
}

// Tap-folded kernel (conv_nfold.cu): weights packed as [k-tile=(seg,chunk,dy)][Npad][CK] with row
// n = dx*cp + co; the 1x1 shortcut segment only fills the centre-dx column block.
int build_nfold(mpg_conv_plan p, const float* w[2], const float* scale[2], const float* shift) {
  const mpg_conv_desc& d = p->d;
  int maxcin = 0;
  for (int s = 0; s < d.nseg; ++s) maxcin = d.seg_cin[s] > maxcin ? d.seg_cin[s] : maxcin;
  auto ksteps_for = [&](int c) {
    int t = 0;
    for (int s = 0; s < d.nseg; ++s) t += d.seg_ksize[s] * ceil_div(d.seg_cin[s], c) * (c / 16);
    return t;
  };
  int ck = maxcin > 32 ? 64 : (maxcin > 16 ? 32 : 16);
  for (int c = ck / 2; c >= 16; c /= 2)
    if (ksteps_for(c) * 5 <= ksteps_for(ck) * 4) ck = c;
  if (const char* e = getenv("MPG_NFOLD_CK")) {
    const int c = atoi(e);
    if (c == 16 || c == 32 || c == 64) ck = c;
  }
  // thin inputs (every segment <= 8 channels at pixel stride 8): un-swizzled whole-row staging, two vertical taps
  // per K=16 MMA (conv_nfold.cu "NS8")
  bool ns8 = true;
  for (int s = 0; s < d.nseg; ++s) ns8 = ns8 && d.seg_cin[s] <= 8 && d.seg_cstride[s] == 8;
  if (const char* e = getenv("MPG_NFOLD_NS8")) ns8 = ns8 && atoi(e) != 0;
  if (ns8) ck = 8;
  const int rb = ck * 2;
  const int ks0 = d.seg_ksize[0], pad0 = ks0 / 2;
  const int cp = round_up(d.cout, 8);
  const int npad = round_up(ks0 * cp, 16);
  p->ck = ck;
  p->npad = npad;
  int ktiles = 0;
  for (int s = 0; s < d.nseg; ++s) {
    p->seg_nchunk[s] = ceil_div(d.seg_cin[s], ck);
    ktiles += ns8 ? (d.seg_ksize[s] + 1) / 2 : p->seg_nchunk[s] * d.seg_ksize[s];
  }
  const size_t tile_elems = static_cast<size_t>(round_up(npad * (ns8 ? 32 : rb), 1024)) / 2;
  std::vector<uint16_t> wp(static_cast<size_t>(ktiles) * tile_elems, 0);
  size_t kt = 0;
  if (ns8) {
    // tile = (segment, tap pair j): K index = t*8 + ci with dy = 2j + t; no-swizzle K-major core matrices:
    // element (n, k) at ((k/8) * (npad/8) + n/8) * 64 + (n%8) * 8 + (k%8)   [in 16-bit elements]
    for (int s = 0; s < d.nseg; ++s) {
      const int ks = d.seg_ksize[s], cin = d.seg_cin[s];
      for (int j = 0; j < (ks + 1) / 2; ++j, ++kt)
        for (int t = 0; t < 2; ++t) {
          const int dy = 2 * j + t;
          if (dy >= ks) continue;
          for (int dx = 0; dx < ks; ++dx) {
            const int col_dx = (ks == 1) ? pad0 : dx;
            for (int n = 0; n < d.cout; ++n) {
              const float sc = scale[s] ? scale[s][n] : 1.0f;
              const int row = col_dx * cp + n;
              for (int ci = 0; ci < cin; ++ci) {
                const float v = w[s][((static_cast<size_t>(dy) * ks + dx) * cin + ci) * d.cout + n] * sc;
                const size_t off = (static_cast<size_t>(t) * (npad / 8) + row / 8) * 64 + (row % 8) * 8 + ci;
                wp[kt * tile_elems + off] = d.in_dtype == MPG_F16 ? f32_to_f16_rn(v) : f32_to_bf16_rn(v);
              }
            }
          }
        }
    }
  }
  for (int s = 0; s < d.nseg && !ns8; ++s) {
    const int ks = d.seg_ksize[s], cin = d.seg_cin[s];
    for (int ch = 0; ch < p->seg_nchunk[s]; ++ch)
      for (int dy = 0; dy < ks; ++dy, ++kt)
        for (int dx = 0; dx < ks; ++dx) {
          const int col_dx = (ks == 1) ? pad0 : dx;  // shortcut: centre column block of the main segment
          for (int n = 0; n < d.cout; ++n) {
            const float sc = scale[s] ? scale[s][n] : 1.0f;
            for (int c = 0; c < ck; ++c) {
              const int ci = ch * ck + c;
              if (ci >= cin) break;
              const float v = w[s][((static_cast<size_t>(dy) * ks + dx) * cin + ci) * d.cout + n] * sc;
              wp[kt * tile_elems + swz_elem(col_dx * cp + n, c, ck)] = d.in_dtype == MPG_F16 ? f32_to_f16_rn(v) : f32_to_bf16_rn(v);
            }
          }
        }
  }
  MPG_CUDA(cudaMalloc(&p->d_wpacked, wp.size() * 2));
  MPG_CUDA(cudaMemcpy(p->d_wpacked, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
  std::vector<float> sh(32, 0.0f);
  if (shift)
    for (int n = 0; n < d.cout; ++n) sh[n] = shift[n];
  MPG_CUDA(cudaMalloc(&p->d_shift, 32 * sizeof(float)));
  MPG_CUDA(cudaMemcpy(p->d_shift, sh.data(), 32 * sizeof(float), cudaMemcpyHostToDevice));
  NfoldParams& q = p->np;
  memset(&q, 0, sizeof(q));
  q.n = d.n;
  q.h = d.h;
  q.w = d.w;
  q.valid_w = kNfWin - (ks0 - 1);
  if (4 * npad <= 256) {
    // thin layers: the serial per-CTA roles (one producer thread, one MMA thread, 4 epilogue warps) bound them, so
    // two small CTAs per SM (256 TMEM columns each) beat one CTA with deeper buffering (measured, thin_probe.py)
    q.naccs = 2;
    q.nbuf = 2;
  } else if (8 * npad <= 512) {
    q.naccs = 4;
    q.nbuf = 2;
  } else if (4 * npad <= 512) {
    q.naccs = 2;
    q.nbuf = 2;
  } else {
    q.naccs = 512 / npad;
    q.nbuf = 1;
  }
  // CTA pairs with resident half weight tiles (conv_nfold.cu PAIR): deep-K layers whose weights neither fit one CTA's
  // shared memory nor leave TMEM room to double-buffer a 3-accumulator weight pass (5x5 128->32: N = 160)
  {
    int cin_total = 0;
    for (int s = 0; s < d.nseg; ++s) cin_total += d.seg_cin[s];
    const size_t half = static_cast<size_t>(npad / 2) * rb;
    const size_t a_stage1 = round_up((kNfRowsAcc + ks0 - 1) * kNfWin * rb, 1024);
    bool pair = !ns8 && ck >= 32 && npad % 32 == 0 && half % 1024 == 0 && cin_total >= 64 && 2 * npad <= 512 &&
                static_cast<size_t>(ktiles) * round_up(npad * rb, 1024) > 64 * 1024 &&
                static_cast<size_t>(ktiles) * half + 3 * a_stage1 <= 200 * 1024 &&
                d.n * ceil_div(d.h, kNfRowsAcc) * ceil_div(d.w, kNfWin - (ks0 - 1)) >= 2;
    if (const char* e = getenv("MPG_NFOLD_PAIR")) pair = pair && atoi(e) != 0;
    q.pair = pair ? 1 : 0;
    if (pair) {
      q.naccs = 1;
      q.nbuf = 512 / npad > 4 ? 4 : 512 / npad;
    }
  }
  if (const char* e = getenv("MPG_NFOLD_NACCS")) {
    const int a = atoi(e);
    if (a >= 1 && a * npad <= 512) q.naccs = a;
  }
  if (const char* e = getenv("MPG_NFOLD_NBUF")) {
    const int b = atoi(e);
    if (b >= 1 && b <= kNfMaxBufs) q.nbuf = b;
  }
  while (q.nbuf > 1 && q.nbuf * q.naccs * npad > 512) --q.nbuf;
  q.rows = q.naccs * kNfRowsAcc;
  q.tiles_x = ceil_div(d.w, q.valid_w);
  q.tiles_y = ceil_div(d.h, q.rows);
  q.num_tiles = q.tiles_x * q.tiles_y * d.n;
  q.nseg = d.nseg;
  for (int s = 0; s < d.nseg; ++s) {
    q.seg_ks[s] = d.seg_ksize[s];
    q.seg_nchunk[s] = p->seg_nchunk[s];
  }
  q.npad = npad;
  q.cp = cp;
  q.cout = d.cout;
  q.act = d.act;
  q.pixel_norm = d.pixel_norm;
  q.in_dtype = d.in_dtype;
  q.out_dtype = d.out_dtype;
  q.out_cstride = d.out_cstride;
  q.a_stage_bytes = round_up((q.rows + ks0 - (ns8 ? 0 : 1)) * kNfWin * rb, 1024);
  q.b_tile_bytes = round_up(npad * (ns8 ? 32 : rb), 1024);
  q.ktiles = ktiles;
  q.bres = (static_cast<size_t>(ktiles) * q.b_tile_bytes <= 64 * 1024) ? 1 : 0;
  if (const char* e = getenv("MPG_NFOLD_BRES")) q.bres = (atoi(e) && static_cast<size_t>(ktiles) * q.b_tile_bytes <= 128 * 1024) ? 1 : 0;
  if (q.pair) {
    q.b_tile_bytes = (npad / 2) * rb;  // this CTA's half of a k-tile
    q.bres = 1;
  }
  // Window images by cp.async producer warps instead of tiled TMA (conv_nfold.cu). OFF by default: measured on B200 the
  // two-warp LDGSTS producer is 3.4x SLOWER than tiled TMA (load-only skeleton of the 5x5 128->32 layer 0.684 ms vs
  // 0.198 ms; ~157 cycles per 512-byte warp copy whatever the address arithmetic costs), so the TMA row rate
  // (~5.5 cycles per <=128-byte box row per SM) stays the bound of that layer. Kept behind MPG_NFOLD_CPASYNC=1 (parity
  // tested) for swizzled layouts with resident weights, where warp 3 is free to be the second producer warp.
  q.a_cpasync = 0;
  q.nprod = 2;
  if (const char* e = getenv("MPG_NFOLD_CPASYNC")) {
    q.a_cpasync = (!ns8 && q.bres && atoi(e) != 0) ? 1 : 0;
    if (atoi(e) >= 4 && atoi(e) <= 8 && atoi(e) % 2 == 0) q.nprod = atoi(e);  // 4 / 6 / 8 producer warps
  }
  for (int s = 0; s < d.nseg; ++s) {
    q.seg_cin[s] = d.seg_cin[s];
    q.seg_cstride[s] = d.seg_cstride[s];
  }
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(q.nbuf * q.naccs * npad)) cols <<= 1;
  q.tmem_cols = cols;
  const int b_bytes_min = q.bres ? ktiles * q.b_tile_bytes : 3 * q.b_tile_bytes;
  int occ = (cols <= 256 && b_bytes_min + 2 * q.a_stage_bytes <= 100 * 1024) ? 2 : 1;
  if (const char* e = getenv("MPG_NFOLD_OCC")) occ = atoi(e) > 0 ? atoi(e) : 1;
  if (q.pair) occ = 1;
  q.threads = (occ == 1) ? kNfMaxThreads : kNfThreads;
  if (const char* e = getenv("MPG_NFOLD_THREADS")) q.threads = atoi(e) == 384 ? 384 : 256;
  if (q.a_cpasync && q.nprod > 2) {
    if (occ != 1) q.nprod = 2;  // the extra producer warps need the registers of a one-CTA-per-SM launch
    else q.threads += 32 * (q.nprod - 2);
  }
  const int budget = (210 * 1024) / occ - (occ > 1 ? 2048 : 0);
  int nb;
  if (q.bres) {
    nb = ktiles;
  } else {
    nb = (budget - 2 * q.a_stage_bytes) / q.b_tile_bytes;
    nb = nb > kNfMaxStagesB ? kNfMaxStagesB : nb;
    if (nb > 4 && (budget - nb * q.b_tile_bytes) / q.a_stage_bytes < 3) nb = 4;
  }
  int na = (budget - nb * q.b_tile_bytes) / q.a_stage_bytes;
  na = na > kNfMaxStagesA ? kNfMaxStagesA : na;
  if (nb < 1 || na < 2 || cols * static_cast<uint32_t>(occ) > 512u) {
    set_error("conv(nfold): bad pipeline depth na=%d nb=%d occ=%d (a_stage %d B, b_tile %d B)", na, nb, occ,
              q.a_stage_bytes, q.b_tile_bytes);
    return MPG_EINVAL;
  }
  q.na = na;
  q.nb = q.bres ? 1 : nb;
  q.shift = p->d_shift;
  q.wpacked = p->d_wpacked;
  if (const char* e = getenv("MPG_NFOLD_DBG")) q.dbg = atoi(e);
  p->smem_bytes = static_cast<size_t>(na) * q.a_stage_bytes + static_cast<size_t>(nb) * q.b_tile_bytes + 1024;
  p->grid = q.num_tiles < p->h->sm_count * occ ? q.num_tiles : p->h->sm_count * occ;
  if (q.pair) {  // whole pairs only; a pair covers two tiles
    const int pairs_needed = (q.num_tiles + 1) / 2;
    int pairs = p->h->sm_count / 2;
    if (pairs > pairs_needed) pairs = pairs_needed;
    p->grid = pairs * 2;
  }
  int r = nfold_set_smem_attr(p->h->device, ck, ks0, q.pair, p->smem_bytes);
  if (r) {
    set_error("cudaFuncSetAttribute(nfold, max dynamic smem %zu) failed: %s", p->smem_bytes,
              cudaGetErrorString(static_cast<cudaError_t>(r)));
    return r;
  }
  p->tm_x_ptr[0] = p->tm_x_ptr[1] = nullptr;
  return 0;
}

// ---- row-streaming kernel with the vertical taps folded into N (conv_vfold.cu) ----
struct VfGeom {
  int ck, cp, npad, nchunks, groups, nchw, n_sc, sc_col;
  int seg_nchunk[2], seg_klast[2];
  int b_tile_bytes, b_sc_tile_bytes, b_bytes, a_stage_bytes, na, nbuf;
};

bool vfold_geometry(const mpg_conv_desc& d, VfGeom* g) {
  if (!igemm_eligible(d) || d.upsample != 1) return false;
  const int ks = d.seg_ksize[0];
  if (ks != 3 && ks != 5) return false;
  if (d.nseg == 2 && d.seg_ksize[1] != 1) return false;
  const int cp = round_up(d.cout, 8);
  if (ks * cp > 256 || cp / 8 > (ks == 5 ? 6 : 8)) return false;
  if (d.out_dtype == MPG_F32 ? d.out_cstride > cp : d.out_cstride != cp) return false;
  g->cp = cp;
  g->npad = round_up(ks * cp, 16);
  g->nchunks = cp / 8;
  // epilogue warp groups x chunks per warp (the epilogue is TMEM-read bound: as many warps in flight as registers allow)
  static const int cfg[9][2] = {{0, 0}, {1, 1}, {2, 1}, {3, 1}, {2, 2}, {3, 2}, {2, 3}, {4, 2}, {4, 2}};
  g->groups = cfg[g->nchunks][0];
  g->nchw = cfg[g->nchunks][1];
  if (const char* e = getenv("MPG_VFOLD_GROUPS")) {
    const int gg = atoi(e);
    if (gg >= 1 && gg <= 4 && ceil_div(g->nchunks, gg) <= (gg == 2 ? 3 : 2)) {
      g->groups = gg;
      g->nchw = ceil_div(g->nchunks, gg);
    }
  }
  const int pad = ks / 2;
  if ((pad * cp) % 16 == 0) {
    g->n_sc = round_up(cp, 16);
    g->sc_col = pad * cp;
  } else {  // a narrow MMA would start at a column that is no multiple of 16: use the whole accumulator width
    g->n_sc = g->npad;
    g->sc_col = 0;
  }
  int maxcin = 0;
  for (int s = 0; s < d.nseg; ++s) maxcin = d.seg_cin[s] > maxcin ? d.seg_cin[s] : maxcin;
  const int budget = 212 * 1024;
  const int first_ck = maxcin > 32 ? 64 : 32;
  for (int ck = first_ck; ck >= 32; ck /= 2) {
    const int rb = ck * 2;
    g->ck = ck;
    g->b_tile_bytes = round_up(g->npad / 2 * rb, 1024);
    g->b_sc_tile_bytes = round_up(g->n_sc / 2 * rb, 1024);
    g->a_stage_bytes = round_up((kVfStrip + ks - 1) * rb, 1024);
    int stages_per_row = 0;
    g->b_bytes = 0;
    for (int s = 0; s < d.nseg; ++s) {
      g->seg_nchunk[s] = ceil_div(d.seg_cin[s], ck);
      g->seg_klast[s] = ceil_div(d.seg_cin[s] - (g->seg_nchunk[s] - 1) * ck, 16);
      stages_per_row += g->seg_nchunk[s];
      g->b_bytes += s == 0 ? g->seg_nchunk[s] * ks * g->b_tile_bytes : g->seg_nchunk[s] * g->b_sc_tile_bytes;
    }
    if (d.nseg == 1) g->seg_nchunk[1] = g->seg_klast[1] = 0;
    int na = (budget - g->b_bytes) / g->a_stage_bytes;
    na = na > kVfMaxStagesA ? kVfMaxStagesA : na;
    g->na = na;
    if (na >= stages_per_row + 1 && na >= 2) break;
    if (ck == 32) return false;
  }
  g->nbuf = 512 / g->npad > kVfMaxBufs ? kVfMaxBufs : 512 / g->npad;
  return g->nbuf >= 2;
}

int build_vfold(mpg_conv_plan p, const float* w[2], const float* scale[2], const float* shift) {
  const mpg_conv_desc& d = p->d;
  VfGeom g;
  if (!vfold_geometry(d, &g)) {
    set_error("conv(vfold): layer does not fit the row-streaming kernel");
    return MPG_ENOSUP;
  }
  const int ks = d.seg_ksize[0], pad = ks / 2, ck = g.ck, cp = g.cp;
  p->ck = ck;
  p->npad = g.npad;
  p->vf_nchw = g.nchw;
  for (int s = 0; s < 2; ++s) p->seg_nchunk[s] = g.seg_nchunk[s];
  // resident image per CTA rank: main tiles (chunk, dx) then shortcut tiles (chunk); a tile = this rank's half of the
  // N rows (row n = dy*cp + co) in the swizzled K-major layout
  std::vector<uint16_t> wp(static_cast<size_t>(g.b_bytes), 0);  // 2 ranks x b_bytes / 2 elements
  auto cvt = [&](float v) { return d.in_dtype == MPG_F16 ? f32_to_f16_rn(v) : f32_to_bf16_rn(v); };
  for (int r = 0; r < 2; ++r) {
    size_t off = static_cast<size_t>(r) * g.b_bytes / 2;  // in elements
    for (int ch = 0; ch < g.seg_nchunk[0]; ++ch)
      for (int dx = 0; dx < ks; ++dx, off += g.b_tile_bytes / 2)
        for (int nr = 0; nr < g.npad / 2; ++nr) {
          const int col = r * (g.npad / 2) + nr;
          const int dy = col / cp, co = col % cp;
          if (dy >= ks || co >= d.cout) continue;
          const float sc = scale[0] ? scale[0][co] : 1.0f;
          for (int c = 0; c < ck; ++c) {
            const int ci = ch * ck + c;
            if (ci >= d.seg_cin[0]) break;
            wp[off + swz_elem(nr, c, ck)] = cvt(w[0][((static_cast<size_t>(dy) * ks + dx) * d.seg_cin[0] + ci) * d.cout + co] * sc);
          }
        }
    for (int ch = 0; d.nseg > 1 && ch < g.seg_nchunk[1]; ++ch, off += g.b_sc_tile_bytes / 2)
      for (int nr = 0; nr < g.n_sc / 2; ++nr) {
        const int col = g.sc_col + r * (g.n_sc / 2) + nr;
        const int dy = col / cp, co = col % cp;
        if (dy != pad || co >= d.cout) continue;
        const float sc = scale[1] ? scale[1][co] : 1.0f;
        for (int c = 0; c < ck; ++c) {
          const int ci = ch * ck + c;
          if (ci >= d.seg_cin[1]) break;
          wp[off + swz_elem(nr, c, ck)] = cvt(w[1][static_cast<size_t>(ci) * d.cout + co] * sc);
        }
      }
  }
  MPG_CUDA(cudaMalloc(&p->d_wpacked, wp.size() * 2));
  MPG_CUDA(cudaMemcpy(p->d_wpacked, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
  std::vector<float> sh(64, 0.0f);
  if (shift)
    for (int n = 0; n < d.cout; ++n) sh[n] = shift[n];
  MPG_CUDA(cudaMalloc(&p->d_shift, 64 * sizeof(float)));
  MPG_CUDA(cudaMemcpy(p->d_shift, sh.data(), 64 * sizeof(float), cudaMemcpyHostToDevice));
  VfoldParams& q = p->vp;
  memset(&q, 0, sizeof(q));
  q.n = d.n;
  q.h = d.h;
  q.w = d.w;
  q.strips2 = ceil_div(ceil_div(d.w, kVfStrip), 2);
  q.total_rows = d.n * q.strips2 * d.h;
  int npairs = p->h->sm_count / 2;
  if (const char* e = getenv("MPG_VFOLD_PAIRS")) npairs = atoi(e) > 0 ? atoi(e) : npairs;
  q.rows_per_pair = ceil_div(q.total_rows, npairs);
  if (q.rows_per_pair < 1) q.rows_per_pair = 1;
  p->grid = 2 * ceil_div(q.total_rows, q.rows_per_pair);
  q.ks = ks;
  q.nseg = d.nseg;
  for (int s = 0; s < 2; ++s) {
    q.seg_nchunk[s] = g.seg_nchunk[s];
    q.seg_klast[s] = g.seg_klast[s];
  }
  q.npad = g.npad;
  q.cp = cp;
  q.cout = d.cout;
  q.n_sc = g.n_sc;
  q.sc_col = g.sc_col;
  q.act = d.act;
  q.pixel_norm = d.pixel_norm;
  q.in_dtype = d.in_dtype;
  q.out_dtype = d.out_dtype;
  q.out_cstride = d.out_cstride;
  q.na = g.na;
  if (const char* e = getenv("MPG_VFOLD_NA")) q.na = (atoi(e) >= 2 && atoi(e) <= g.na) ? atoi(e) : g.na;
  q.a_stage_bytes = g.a_stage_bytes;
  q.b_tile_bytes = g.b_tile_bytes;
  q.b_sc_tile_bytes = g.b_sc_tile_bytes;
  q.b_bytes = g.b_bytes;
  q.nbuf = g.nbuf;
  if (const char* e = getenv("MPG_VFOLD_NBUF")) q.nbuf = (atoi(e) >= 1 && atoi(e) <= g.nbuf) ? atoi(e) : g.nbuf;
  q.epi_groups = g.groups;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(q.nbuf * q.npad)) cols <<= 1;
  q.tmem_cols = cols;
  q.shift = p->d_shift;
  q.wpacked = p->d_wpacked;
  if (const char* e = getenv("MPG_VFOLD_DBG")) q.dbg = atoi(e);
  p->smem_bytes = static_cast<size_t>(q.b_bytes) + static_cast<size_t>(q.na) * q.a_stage_bytes + 1024;
  int r = vfold_set_smem_attr(p->h->device, ck, ks, g.nchw, g.groups, p->smem_bytes);
  if (r) {
    set_error("cudaFuncSetAttribute(vfold, max dynamic smem %zu) failed: %s", p->smem_bytes,
              cudaGetErrorString(static_cast<cudaError_t>(r)));
    return r;
  }
  p->tm_x_ptr[0] = p->tm_x_ptr[1] = nullptr;
  return 0;
}

// The row-streaming kernel pays off when a strip pair is mostly real pixels and every CTA pair gets a long row range.
bool vfold_preferred(const mpg_conv_desc& d, int sm_count) {
  VfGeom g;
  if (!vfold_geometry(d, &g)) return false;
  // The epilogue reads k*cp accumulator columns per output pixel and TMEM reads run at ~40-64 B/clk/SM (measured: the
  // epilogue alone costs ~12.7 cycles per column and image row), so the fold only pays when the MMAs of a row take longer:
  // k * (Cin/16) K-steps of max(64, N/2) cycles, i.e. 5x5 layers with >= 48 input channels.
  int min_cin = 48;
  if (const char* e = getenv("MPG_VFOLD_MINCIN")) min_cin = atoi(e);
  if (d.seg_ksize[0] != 5 || d.seg_cin[0] < min_cin) return false;
  const int strips = ceil_div(d.w, kVfStrip);
  if (d.w < 2 * kVfStrip || (strips % 2 != 0 && strips < 5)) return false;
  // (independent of the batch size on purpose: the same layer must pick the same kernel in a 1-GPU and an N-GPU run, which
  //  process different numbers of slices per launch, for the sharded volume to stay bit-identical)
  (void)sm_count;
  return d.h >= 64;
}

// ---- row-streaming kernel with the vertical-tap sum accumulated in a TMEM ring (conv_vring.cu) ----
struct VrGeom {
  int ck, cp, cs, nslots, nchunks, groups, nchw, pair;
  int seg_nchunk[2], seg_klast[2], seg_ksteps[2];
  int b_tile_bytes, b_sc_tile_bytes, b_bytes, a_stage_bytes, na;
};

// pair: -1 = choose, 0 = single CTA, 1 = CTA pairs
bool vring_geometry(const mpg_conv_desc& d, VrGeom* g, int pair = -1) {
  if (!igemm_eligible(d) || d.upsample != 1) return false;
  const int ks = d.seg_ksize[0];
  if (ks != 3 && ks != 5) return false;
  if (d.nseg == 2 && d.seg_ksize[1] != 1) return false;
  const int cp = round_up(d.cout, 8), cs = round_up(d.cout, 16);
  if (ks * cs > 256 || cp > 64) return false;
  if (d.out_dtype == MPG_F32 ? d.out_cstride > cp : d.out_cstride != cp) return false;
  g->cp = cp;
  g->cs = cs;
  g->nslots = 512 / cs - (ks - 1);  // ring slots; the k-1 overflow slots follow
  if (g->nslots > kVrMaxSlots) g->nslots = kVrMaxSlots;
  if (g->nslots < ks + 1) return false;
  g->nchunks = cp / 8;
  // epilogue warp groups per TMEM lane quarter x 8-column chunks per warp (measured: 48 columns 3 x 2, 64 columns 4 x 2
  // are 4-6 % faster than 2 x 3 / 2 x 4)
  g->groups = g->nchunks > 6 ? 4 : (g->nchunks > 4 ? 3 : (g->nchunks > 1 ? 2 : 1));
  g->nchw = ceil_div(g->nchunks, g->groups);
  if (const char* e = getenv("MPG_VRING_GROUPS")) {
    const int gg = atoi(e);
    if (gg == 2 && g->nchunks > 1) g->groups = 2, g->nchw = ceil_div(g->nchunks, 2);
    if (gg == 4 && g->nchunks >= 3) g->groups = 4, g->nchw = ceil_div(g->nchunks, 4);
  }
  int maxcin = 0;
  for (int s = 0; s < d.nseg; ++s) maxcin = d.seg_cin[s] > maxcin ? d.seg_cin[s] : maxcin;
  const int ck = maxcin > 32 ? 64 : 32, rb = ck * 2;
  g->ck = ck;
  g->a_stage_bytes = round_up((kVrStrip + ks - 1) * rb, 1024);
  int stages_per_row = 0, tiles0 = 0, tiles1 = 0;
  g->seg_nchunk[1] = g->seg_klast[1] = g->seg_ksteps[1] = 0;
  for (int s = 0; s < d.nseg; ++s) {
    g->seg_nchunk[s] = ceil_div(d.seg_cin[s], ck);
    g->seg_klast[s] = ceil_div(d.seg_cin[s] - (g->seg_nchunk[s] - 1) * ck, 16);
    g->seg_ksteps[s] = (g->seg_nchunk[s] - 1) * (ck / 16) + g->seg_klast[s];
    stages_per_row += g->seg_nchunk[s];
    (s == 0 ? tiles0 : tiles1) = s == 0 ? g->seg_ksteps[s] * ks : g->seg_ksteps[s];
  }
  const int budget = 212 * 1024;
  const int strips = ceil_div(d.w, kVrStrip);
  for (int pr = (pair < 0 ? 0 : pair); pr <= (pair < 0 ? 1 : pair); ++pr) {
    g->pair = pr;
    g->b_tile_bytes = ks * cs * 32 / (pr ? 2 : 1);
    g->b_sc_tile_bytes = cs * 32 / (pr ? 2 : 1);
    g->b_bytes = round_up(tiles0 * g->b_tile_bytes + tiles1 * g->b_sc_tile_bytes, 1024);
    int na = (budget - g->b_bytes) / g->a_stage_bytes;
    na = na > kVrMaxStagesA ? kVrMaxStagesA : na;
    g->na = na;
    const bool fits = na >= stages_per_row + 1 && na >= 2;
    if (pair >= 0) return fits && (!pr || strips >= 2);
    // choose: a single CTA whenever the weights fit it. Pairs halve the per-CTA weight fetch of every MMA, but their ring is
    // handed over across two SMs and (for wide slots) only k+1 slots deep: measured slower on every layer that fits one CTA
    // (5x5 48->48: 0.179 ms single, 0.266 ms pairs), faster than the fold where only the halves fit (5x5 96->48: 0.314 vs 0.408)
    if (pr == 0 && fits) return true;
    if (pr == 1 && fits && strips >= 2) return true;
  }
  return false;
}

int build_vring(mpg_conv_plan p, const float* w[2], const float* scale[2], const float* shift) {
  const mpg_conv_desc& d = p->d;
  VrGeom g;
  int want = -1;
  if (const char* e = getenv("MPG_VRING_PAIR")) want = atoi(e) != 0 ? 1 : 0;
  if (!vring_geometry(d, &g, want) && !(want >= 0 && vring_geometry(d, &g, -1))) {
    set_error("conv(vring): layer does not fit the TMEM-ring row-streaming kernel");
    return MPG_ENOSUP;
  }
  const int ks = d.seg_ksize[0], cs = g.cs, nr = g.pair ? 2 : 1;
  p->ck = g.ck;
  p->npad = ks * cs;
  p->vf_nchw = g.nchw;
  for (int s = 0; s < 2; ++s) p->seg_nchunk[s] = g.seg_nchunk[s];
  // resident image (per CTA rank): main tiles [K-step][dx] of ks*cs rows (row = (ks-1-dy)*cs + co) x 16 channels, then
  // shortcut tiles [K-step] of cs rows; 32-byte rows in the SWIZZLE_32B K-major layout. A pair splits every tile's rows.
  std::vector<uint16_t> wp(static_cast<size_t>(g.b_bytes) / 2 * nr, 0);
  auto cvt = [&](float v) { return d.in_dtype == MPG_F16 ? f32_to_f16_rn(v) : f32_to_bf16_rn(v); };
  const int half0 = ks * cs / nr, half1 = cs / nr;
  for (int r = 0; r < nr; ++r) {
    size_t off = static_cast<size_t>(r) * g.b_bytes / 2;
    for (int kk = 0; kk < g.seg_ksteps[0]; ++kk)
      for (int dx = 0; dx < ks; ++dx, off += g.b_tile_bytes / 2)
        for (int dy = 0; dy < ks; ++dy)
          for (int co = 0; co < d.cout; ++co) {
            const int row = (ks - 1 - dy) * cs + co - r * half0;
            if (row < 0 || row >= half0) continue;
            const float sc = scale[0] ? scale[0][co] : 1.0f;
            for (int c = 0; c < 16; ++c) {
              const int ci = kk * 16 + c;
              if (ci >= d.seg_cin[0]) break;
              wp[off + swz_elem(row, c, 16)] = cvt(w[0][((static_cast<size_t>(dy) * ks + dx) * d.seg_cin[0] + ci) * d.cout + co] * sc);
            }
          }
    for (int kk = 0; d.nseg > 1 && kk < g.seg_ksteps[1]; ++kk, off += g.b_sc_tile_bytes / 2)
      for (int co = 0; co < d.cout; ++co) {
        const int row = co - r * half1;
        if (row < 0 || row >= half1) continue;
        const float sc = scale[1] ? scale[1][co] : 1.0f;
        for (int c = 0; c < 16; ++c) {
          const int ci = kk * 16 + c;
          if (ci >= d.seg_cin[1]) break;
          wp[off + swz_elem(row, c, 16)] = cvt(w[1][static_cast<size_t>(ci) * d.cout + co] * sc);
        }
      }
  }
  MPG_CUDA(cudaMalloc(&p->d_wpacked, wp.size() * 2));
  MPG_CUDA(cudaMemcpy(p->d_wpacked, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
  std::vector<float> sh(64, 0.0f);
  if (shift)
    for (int n = 0; n < d.cout; ++n) sh[n] = shift[n];
  MPG_CUDA(cudaMalloc(&p->d_shift, 64 * sizeof(float)));
  MPG_CUDA(cudaMemcpy(p->d_shift, sh.data(), 64 * sizeof(float), cudaMemcpyHostToDevice));
  VringParams& q = p->rp;
  memset(&q, 0, sizeof(q));
  q.n = d.n;
  q.h = d.h;
  q.w = d.w;
  q.pair = g.pair;
  q.units_x = ceil_div(ceil_div(d.w, kVrStrip), nr);
  q.total_rows = d.n * q.units_x * d.h;
  int nctas = p->h->sm_count / nr;
  if (const char* e = getenv("MPG_VRING_CTAS")) nctas = atoi(e) > 0 ? atoi(e) : nctas;
  q.rows_per_cta = ceil_div(q.total_rows, nctas);
  if (q.rows_per_cta < 1) q.rows_per_cta = 1;
  p->grid = nr * ceil_div(q.total_rows, q.rows_per_cta);
  q.ks = ks;
  q.nseg = d.nseg;
  for (int s = 0; s < 2; ++s) {
    q.seg_nchunk[s] = g.seg_nchunk[s];
    q.seg_klast[s] = g.seg_klast[s];
  }
  q.cs = cs;
  q.nslots = g.nslots;
  if (const char* e = getenv("MPG_VRING_SLOTS")) q.nslots = (atoi(e) >= ks + 1 && atoi(e) <= g.nslots) ? atoi(e) : g.nslots;
  q.cp = g.cp;
  q.cout = d.cout;
  q.act = d.act;
  q.pixel_norm = d.pixel_norm;
  q.in_dtype = d.in_dtype;
  q.out_dtype = d.out_dtype;
  q.out_cstride = d.out_cstride;
  q.na = g.na;
  q.a_stage_bytes = g.a_stage_bytes;
  q.b_tile_bytes = g.b_tile_bytes;
  q.b_sc_tile_bytes = g.b_sc_tile_bytes;
  q.b_bytes = g.b_bytes;
  q.epi_groups = g.groups;
  q.shift = p->d_shift;
  q.wpacked = p->d_wpacked;
  if (const char* e = getenv("MPG_VRING_DBG")) q.dbg = atoi(e);
  p->smem_bytes = static_cast<size_t>(q.b_bytes) + static_cast<size_t>(q.na) * q.a_stage_bytes + 1024;
  int r = vring_set_smem_attr(p->h->device, g.ck, ks, g.nchw, g.groups, g.pair, p->smem_bytes);
  if (r) {
    set_error("cudaFuncSetAttribute(vring, max dynamic smem %zu) failed: %s", p->smem_bytes,
              cudaGetErrorString(static_cast<cudaError_t>(r)));
    return r;
  }
  p->tm_x_ptr[0] = p->tm_x_ptr[1] = nullptr;
  return 0;
}

// Measured per layer (tools/thin_probe.py, tools/step_times*.py): the ring beats the tap-by-tap kernel, both folds and the
// CUDA-core paths on every eligible layer except <= 8 output channels (the horizontal fold is as fast there).
bool vring_preferred(const mpg_conv_desc& d, int sm_count) {
  VrGeom g;
  if (!vring_geometry(d, &g)) return false;
  if (const char* e = getenv("MPG_CONV_VRING")) {
    if (atoi(e) == 0) return false;
    if (atoi(e) == 2) return true;
  }
  if (g.cp <= 8) return false;
  if (g.pair) {
    // pairs pay off on deep-K 5x5 layers only (5x5 96->48: 0.59 ms vs 0.88 with the fold; 3x3 128/128->64 at 256^2: 0.43 vs
    // 0.35 on the tap-by-tap kernel)
    if (d.seg_ksize[0] != 5 || d.seg_cin[0] < 64) return false;
    if (const char* e = getenv("MPG_CONV_VRING_PAIRS"))
      if (atoi(e) == 0) return false;
  }
  const int strips = ceil_div(d.w, kVrStrip);
  if (strips * kVrStrip * 4 > d.w * 5) return false;  // > 25 % of a strip row would be padding
  if (g.pair && strips % 2 != 0 && strips < 5) return false;
  // (independent of the batch size on purpose, see vfold_preferred)
  (void)sm_count;
  return d.h >= 64;
}

int build_direct(mpg_conv_plan p, const float* w[2], const float* scale[2], const float* shift) {
  const mpg_conv_desc& d = p->d;
  DirectParams& dp = p->dp;
  memset(&dp, 0, sizeof(dp));
  const int coutp = round_up(d.cout, 8);
  size_t total = 0;
  for (int s = 0; s < d.nseg; ++s) {
    dp.seg_woff[s] = static_cast<long long>(total);
    total += static_cast<size_t>(d.seg_ksize[s]) * d.seg_ksize[s] * d.seg_cin[s] * coutp;
  }
  std::vector<float> wd(total + coutp, 0.0f);
  for (int s = 0; s < d.nseg; ++s) {
    const int ks = d.seg_ksize[s], cin = d.seg_cin[s];
    for (int tap = 0; tap < ks * ks; ++tap)
      for (int ci = 0; ci < cin; ++ci)
        for (int n = 0; n < d.cout; ++n)
          wd[dp.seg_woff[s] + (static_cast<size_t>(tap) * cin + ci) * coutp + n] =
              w[s][(static_cast<size_t>(tap) * cin + ci) * d.cout + n] * (scale[s] ? scale[s][n] : 1.0f);
  }
  if (shift)
    for (int n = 0; n < d.cout; ++n) wd[total + n] = shift[n];
  MPG_CUDA(cudaMalloc(&p->d_wdirect, wd.size() * sizeof(float)));
  MPG_CUDA(cudaMemcpy(p->d_wdirect, wd.data(), wd.size() * sizeof(float), cudaMemcpyHostToDevice));
  dp.n = d.n;
  dp.in_upsample = d.in_upsample;
  dp.h = d.h;
  dp.w = d.w;
  dp.src_h = d.h / d.in_upsample;
  dp.src_w = d.w / d.in_upsample;
  dp.stride = d.stride;
  dp.oh = p->oh;
  dp.ow = p->ow;
  dp.nseg = d.nseg;
  for (int s = 0; s < d.nseg; ++s) {
    dp.seg_ks[s] = d.seg_ksize[s];
    dp.seg_cin[s] = d.seg_cin[s];
    dp.seg_cstride[s] = d.seg_cstride[s];
    dp.seg_pad[s] = same_pad_before(d.h, d.seg_ksize[s], d.stride);
  }
  dp.shift = p->d_wdirect + total;
  dp.cout = d.cout;
  dp.coutp = coutp;
  dp.act = d.act;
  dp.pixel_norm = d.pixel_norm;
  dp.upsample = d.upsample;
  dp.in_dtype = d.in_dtype;
  dp.out_dtype = d.out_dtype;
  dp.out_cstride = d.out_cstride;
  return 0;
}

// CUDA-core kernel for the channel-less tail layers (conv_tiny.cu): weights go into the kernel parameter block as
// [dy][dx][ci bucket][cout], scale folded.
int build_tiny(mpg_conv_plan p, const float* w[2], const float* scale[2], const float* shift) {
  const mpg_conv_desc& d = p->d;
  TinyParams& t = p->tp;
  memset(&t, 0, sizeof(t));
  t.n = d.n;
  t.h = d.h;
  t.w = d.w;
  t.act = d.act;
  t.in_dtype = d.in_dtype;
  t.out_dtype = d.out_dtype;
  t.out_cstride = d.out_cstride;
  const int ks = d.seg_ksize[0], cin = d.seg_cin[0];
  const int cb = cin <= 2 ? 2 : (cin <= 4 ? 4 : 8);
  for (int tap = 0; tap < ks * ks; ++tap)
    for (int ci = 0; ci < cin; ++ci)
      for (int n = 0; n < d.cout; ++n)
        t.w0[(tap * cb + ci) * d.cout + n] = w[0][(static_cast<size_t>(tap) * cin + ci) * d.cout + n] * (scale[0] ? scale[0][n] : 1.0f);
  if (d.nseg == 2)
    for (int ci = 0; ci < d.seg_cin[1]; ++ci)
      for (int n = 0; n < d.cout; ++n)
        t.w1[ci * d.cout + n] = w[1][static_cast<size_t>(ci) * d.cout + n] * (scale[1] ? scale[1][n] : 1.0f);
  for (int n = 0; n < d.cout; ++n) t.shift[n] = shift ? shift[n] : 0.0f;
  return 0;
}

}  // namespace

extern "C" {

int mpg_conv_plan_create(mpg_handle h, const mpg_conv_desc* dsc, const float* w_seg0, const float* w_seg1,
                         const float* scale_seg0, const float* scale_seg1, const float* shift,
                         mpg_conv_plan* out) {
  MPG_CHECK_ARG(h && dsc && out && w_seg0, "mpg_conv_plan_create: null argument");
  mpg_conv_desc d = *dsc;
  if (d.upsample <= 0) d.upsample = 1;
  if (d.in_upsample <= 0) d.in_upsample = 1;
  if (d.stride <= 0) d.stride = 1;
  MPG_CHECK_ARG(d.n > 0 && d.h > 0 && d.w > 0, "conv: bad spatial size n=%d h=%d w=%d", d.n, d.h, d.w);
  MPG_CHECK_ARG(d.nseg == 1 || d.nseg == 2, "conv: nseg must be 1 or 2 (got %d)", d.nseg);
  MPG_CHECK_ARG(d.nseg == 1 || w_seg1 != nullptr, "conv: segment 1 weights missing");
  MPG_CHECK_ARG(d.cout > 0 && d.out_cstride >= d.cout, "conv: cout=%d out_cstride=%d", d.cout, d.out_cstride);
  MPG_CHECK_ARG(d.act >= MPG_ACT_NONE && d.act <= MPG_ACT_TANH, "conv: unknown activation %d", d.act);
  MPG_CHECK_ARG(mpg::is_dtype(d.in_dtype), "conv: bad in_dtype %d", d.in_dtype);
  MPG_CHECK_ARG(mpg::is_dtype(d.out_dtype), "conv: bad out_dtype %d", d.out_dtype);
  MPG_CHECK_ARG(d.h % d.in_upsample == 0 && d.w % d.in_upsample == 0, "conv: h,w not divisible by in_upsample");
  for (int s = 0; s < d.nseg; ++s) {
    MPG_CHECK_ARG(d.seg_cin[s] > 0 && d.seg_cstride[s] >= d.seg_cin[s] && d.seg_ksize[s] > 0,
                  "conv: segment %d cin=%d cstride=%d k=%d", s, d.seg_cin[s], d.seg_cstride[s], d.seg_ksize[s]);
  }
  int kind = d.force_kind;
  const bool elig = igemm_eligible(d);
  if (kind == 0) {
    // Every 16-bit stride-1 conv goes to the tensor cores, thin layers included: a 25-tap conv with
    // Cin <= 16 is 25 K=16 MMAs per 128 pixels, far cheaper than the CUDA-core loop (TMA zero-fills
    // the missing channels, the weights of padded output channels are zero).
    kind = elig ? 1 : 2;
    // narrow Cout: fold the horizontal taps into the GEMM N dimension (conv_nfold.cu)
    // (pays off when the folded N stays small, or when K is deep enough that the 5x fewer MMAs outweigh the
    //  un-overlapped epilogue of the N=160 single-buffered configuration; measured: tools/thin_probe.py)
    int cin_total = 0;
    for (int s = 0; s < d.nseg; ++s) cin_total += d.seg_cin[s];
    bool thin_in = true;  // every segment <= 8 channels at pixel stride 8: the un-swizzled two-taps-per-MMA staging applies
    for (int s = 0; s < d.nseg; ++s) thin_in = thin_in && d.seg_cin[s] <= 8 && d.seg_cstride[s] == 8;
    (void)thin_in;
    // (re-measured with the lean issue loops on the 8x generators, tools/step_times_8x.py: widening the rule to 5x5 layers
    //  with >= 24 input channels, or padding N = 48 to 64 for CTA pairs, both left the frame time unchanged or worse)
    bool nf = nfold_eligible(d) && (round_up(d.cout, 8) * d.seg_ksize[0] <= 64 || cin_total >= 64);
    if (const char* e = getenv("MPG_CONV_NFOLD")) nf = (atoi(e) == 2) ? nfold_eligible(d) : (nf && atoi(e) != 0);
    if (kind == 1 && nf) kind = 3;
    // medium / narrow Cout on wide images: stream image rows, fold the VERTICAL taps into N (conv_vfold.cu)
    if (kind == 1 || kind == 3) {
      bool vf = vfold_preferred(d, h->sm_count);
      if (const char* e = getenv("MPG_CONV_VFOLD")) {
        VfGeom g;
        vf = (atoi(e) == 2) ? vfold_geometry(d, &g) : (vf && atoi(e) != 0);
      }
      // the TMEM-ring variant first: it beats the fold on every layer whose weights fit one CTA (measured, thin_probe.py)
      if (vring_preferred(d, h->sm_count)) kind = 6;
      else if (vf) kind = 5;
    }
    // Cout <= 2 from <= 8 channels: a bandwidth kernel on CUDA cores beats the per-tile hand-overs of the tensor path
    bool tiny = tiny_eligible(d);
    if (const char* e = getenv("MPG_CONV_TINY")) tiny = tiny && atoi(e) != 0;
    if (tiny) kind = 4;
  }
  if (kind == 4 && !tiny_eligible(d)) {
    mpg::set_error("conv: tiny CUDA-core path needs 16-bit input at channel stride 8, cin <= 8, cout <= 2, k0 in {3,5}, 1x1 shortcut, stride 1");
    return MPG_ENOSUP;
  }
  if (kind == 3 && !nfold_eligible(d)) {
    mpg::set_error("conv: tap-folded path needs the tcgen05 constraints plus cout <= 32, k0 in {3,5}, 1x1 shortcut, no upsample");
    return MPG_ENOSUP;
  }
  if (kind == 1 && !elig) {
    mpg::set_error("conv: tcgen05 path needs bf16/f16 input (same 16-bit output type or f32), stride 1, k in {1,3,5}, cstride %% 8 == 0, cout <= 128");
    return MPG_ENOSUP;
  }
  if (kind == 5) {
    VfGeom g;
    if (!vfold_geometry(d, &g)) {
      mpg::set_error("conv: row-streaming path needs the tcgen05 constraints plus k0 in {3,5}, k0 * round_up(cout, 8) <= 256, 1x1 shortcut, no upsample");
      return MPG_ENOSUP;
    }
  }
  if (kind == 6) {
    VrGeom g;
    if (!vring_geometry(d, &g)) {
      mpg::set_error("conv: TMEM-ring path needs the tcgen05 constraints plus k0 in {3,5}, k0 * round_up(cout, 16) <= 256, cout <= 64, 1x1 shortcut, no upsample, weights <= ~150 KB");
      return MPG_ENOSUP;
    }
  }
  MPG_CHECK_ARG(kind >= 1 && kind <= 6, "conv: bad force_kind %d", d.force_kind);

  mpg_conv_plan p = new mpg_conv_plan_s();
  memset(p, 0, sizeof(*p));
  p->h = h;
  p->d = d;
  p->kind = kind;
  p->oh = (d.h + d.stride - 1) / d.stride;
  p->ow = (d.w + d.stride - 1) / d.stride;
  p->flops = 0.0;
  for (int s = 0; s < d.nseg; ++s)
    p->flops += 2.0 * d.n * p->oh * p->ow * static_cast<double>(d.seg_ksize[s]) * d.seg_ksize[s] * d.seg_cin[s] * d.cout;
  const float* w[2] = {w_seg0, w_seg1};
  const float* sc[2] = {scale_seg0, scale_seg1};
  mpg::DeviceGuard guard(h->device);
  int r = (kind == 1) ? build_igemm(p, w, sc, shift)
          : (kind == 3 ? build_nfold(p, w, sc, shift)
                       : (kind == 4 ? build_tiny(p, w, sc, shift) : (kind == 5 ? build_vfold(p, w, sc, shift) : (kind == 6 ? build_vring(p, w, sc, shift) : build_direct(p, w, sc, shift)))));
  if (r) {
    mpg_conv_plan_destroy(p);
    return r;
  }
  *out = p;
  return MPG_OK;
}

int mpg_conv_plan_set_side(mpg_conv_plan p, const float* w_side, int side_cout) {
  MPG_CHECK_ARG(p && w_side && side_cout >= 1, "mpg_conv_plan_set_side: bad argument");
  const mpg::IgemmParams& ip = p->ip;
  if (p->kind != 1 || !ip.tma_store || !ip.pair || p->ck != 64 || ip.threads != mpg::kIgMaxThreads || ip.npad != 128 || p->d.cout != 128 || ip.pixel_norm ||
      ip.upsample != 1 || side_cout > 8) {
    mpg::set_error("conv: a side output needs a 128-channel tcgen05 plan with the per-warp TMA-store epilogue (8 epilogue warps, no "
                   "pixel_norm / upsample) and <= 8 side channels");
    return MPG_ENOSUP;
  }
  const size_t need = static_cast<size_t>(ip.stage_off) + static_cast<size_t>(ip.stage_bytes) + 8192 + 1024;
  if (need > 227 * 1024) {
    mpg::set_error("conv: no shared memory left for the side output (%zu B)", need);
    return MPG_ENOSUP;
  }
  mpg::DeviceGuard guard(p->h->device);
  std::vector<float> w(128 * 8, 0.0f);
  for (int c = 0; c < 128; ++c)
    for (int k = 0; k < side_cout; ++k) w[c * 8 + k] = w_side[static_cast<size_t>(c) * side_cout + k];
  if (!p->d_side_w) MPG_CUDA(cudaMalloc(&p->d_side_w, w.size() * sizeof(float)));
  MPG_CUDA(cudaMemcpy(p->d_side_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  p->ip.side = 1;
  p->ip.side_off = ip.stage_off + ip.stage_bytes;
  p->ip.side_w = p->d_side_w;
  p->smem_bytes = need;
  int r = mpg::igemm_set_smem_attr(p->h->device, p->ck, ip.pair, p->smem_bytes, 1);
  if (r) {
    mpg::set_error("cudaFuncSetAttribute(max dynamic smem %zu) failed: %s", p->smem_bytes, cudaGetErrorString(static_cast<cudaError_t>(r)));
    return r;
  }
  return MPG_OK;
}

int mpg_conv_plan_run(mpg_conv_plan p, const void* x0, const void* x1, void* y, void* stream) {
  return mpg_conv_plan_run_ex(p, x0, x1, y, nullptr, nullptr, stream);
}

int mpg_conv_plan_run_ex(mpg_conv_plan p, const void* x0, const void* x1, void* y, float* y_side, const float* residual,
                         void* stream) {
  MPG_CHECK_ARG(p && x0 && y, "mpg_conv_plan_run: null argument");
  MPG_CHECK_ARG((y_side != nullptr) == (p->kind == 1 && p->ip.side != 0), "conv: y_side must be given exactly when the plan has a side output");
  MPG_CHECK_ARG(y_side == nullptr || (reinterpret_cast<uintptr_t>(y_side) & 15) == 0, "conv: y_side not 16-byte aligned");
  if (residual != nullptr) {
    MPG_CHECK_ARG((p->kind == 3 || p->kind == 5 || p->kind == 6) && p->d.cout <= 8 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
                  "conv: an fp32 residual input needs a tap-folded plan with <= 8 output channels and a 16-byte aligned tensor");
  }
  MPG_CHECK_ARG(p->d.nseg == 1 || x1 != nullptr, "mpg_conv_plan_run: segment 1 input missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mpg::DeviceGuard guard(p->h->device);  // the plan's device, whatever the caller's current device is
  const mpg_conv_desc& d = p->d;
  if (p->kind == 1) {
    const void* xs[2] = {x0, x1};
    for (int s = 0; s < d.nseg; ++s) {
      if (p->tm_x_ptr[s] == xs[s]) continue;
      MPG_CHECK_ARG((reinterpret_cast<uintptr_t>(xs[s]) & 15) == 0, "conv: input %d not 16-byte aligned", s);
      const uint64_t cs = static_cast<uint64_t>(d.seg_cstride[s]) * 2;
      const uint64_t dims[4] = {static_cast<uint64_t>(d.seg_cin[s]), static_cast<uint64_t>(d.w),
                                static_cast<uint64_t>(d.h), static_cast<uint64_t>(d.n)};
      const uint64_t strides[3] = {cs, cs * d.w, cs * d.w * d.h};
      const uint32_t box[4] = {static_cast<uint32_t>(p->ck),
                               static_cast<uint32_t>(mpg::kIgTileW + d.seg_ksize[s] - 1),
                               static_cast<uint32_t>(mpg::kIgTileH + d.seg_ksize[s] - 1), 1u};
      int r = mpg::encode_tmap(p->h, &p->tm_x[s], d.in_dtype == MPG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, xs[s], dims, strides,
                               box, swizzle_for(p->ck));
      if (r) return r;
      p->tm_x_ptr[s] = xs[s];
    }
    if (d.out_dtype != MPG_F32)
      MPG_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "conv: output not 16-byte aligned");
    mpg::IgemmParams ip = p->ip;
    ip.out = y;
    ip.side_out = y_side;
    if (ip.tma_store && p->tm_y_ptr != y) {
      // per-warp store box: 64 channels x 8 px x 4 image rows, 128B-swizzled staging
      const CUtensorMapDataType dt = d.out_dtype == MPG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
      const uint64_t cs = static_cast<uint64_t>(d.out_cstride) * 2;
      int r;
      if (d.upsample == 2) {
        // y[n, 2h+uy, 2w+ux, c] as (c, ux, w, uy, n*H + h)
        const uint64_t dims[5] = {static_cast<uint64_t>(d.out_cstride), 2u, static_cast<uint64_t>(d.w), 2u,
                                  static_cast<uint64_t>(d.n) * d.h};
        const uint64_t strides[4] = {cs, 2 * cs, 2 * cs * d.w, 4 * cs * d.w};
        const uint32_t box[5] = {64u, 1u, 8u, 1u, 4u};
        r = mpg::encode_tmap(p->h, &p->tm_y, dt, 5, y, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      } else {
        const uint64_t dims[4] = {static_cast<uint64_t>(d.out_cstride), static_cast<uint64_t>(d.w),
                                  static_cast<uint64_t>(d.h), static_cast<uint64_t>(d.n)};
        const uint64_t strides[3] = {cs, cs * d.w, cs * d.w * d.h};
        const uint32_t box[4] = {64u, 8u, 4u, 1u};
        r = mpg::encode_tmap(p->h, &p->tm_y, dt, 4, y, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      }
      if (r) return r;
      p->tm_y_ptr = y;
    }
    int r = mpg::igemm_launch(p->ck, p->tm_x[0], d.nseg > 1 ? p->tm_x[1] : p->tm_x[0], p->tm_w,
                              ip.tma_store ? p->tm_y : p->tm_w, ip, p->grid, p->smem_bytes, st);
    if (r) {
      mpg::set_error("conv igemm launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(r)));
      return r;
    }
    return MPG_OK;
  }
  if (p->kind == 3) {
    const void* xs[2] = {x0, x1};
    for (int s = 0; s < d.nseg; ++s) {
      if (p->tm_x_ptr[s] == xs[s]) continue;
      MPG_CHECK_ARG((reinterpret_cast<uintptr_t>(xs[s]) & 15) == 0, "conv: input %d not 16-byte aligned", s);
      const uint64_t cs = static_cast<uint64_t>(d.seg_cstride[s]) * 2;
      int r;
      if (p->ck == 8) {  // NS8: [N][H][W*8] view, whole 32-pixel rows as the inner box dimension, no swizzle
        const uint64_t dims[3] = {static_cast<uint64_t>(d.w) * 8, static_cast<uint64_t>(d.h), static_cast<uint64_t>(d.n)};
        const uint64_t strides[2] = {cs * d.w, cs * d.w * d.h};
        const uint32_t box[3] = {static_cast<uint32_t>(mpg::kNfWin * 8), static_cast<uint32_t>(p->np.rows + d.seg_ksize[s]), 1u};
        r = mpg::encode_tmap(p->h, &p->tm_x[s], d.in_dtype == MPG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, xs[s], dims,
                             strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
      } else {
        const uint64_t dims[4] = {static_cast<uint64_t>(d.seg_cin[s]), static_cast<uint64_t>(d.w),
                                  static_cast<uint64_t>(d.h), static_cast<uint64_t>(d.n)};
        const uint64_t strides[3] = {cs, cs * d.w, cs * d.w * d.h};
        const uint32_t box[4] = {static_cast<uint32_t>(p->ck), static_cast<uint32_t>(mpg::kNfWin),
                                 static_cast<uint32_t>(p->np.rows + d.seg_ksize[s] - 1), 1u};
        r = mpg::encode_tmap(p->h, &p->tm_x[s], d.in_dtype == MPG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, xs[s], dims, strides,
                             box, swizzle_for(p->ck));
      }
      if (r) return r;
      p->tm_x_ptr[s] = xs[s];
    }
    if (d.out_dtype != MPG_F32)
      MPG_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "conv: output not 16-byte aligned");
    mpg::NfoldParams q = p->np;
    q.out = y;
    q.x[0] = x0;
    q.x[1] = x1;
    q.resid = residual;
    int r = mpg::nfold_launch(p->ck, p->tm_x[0], d.nseg > 1 ? p->tm_x[1] : p->tm_x[0], p->tm_w, q, p->grid,
                              p->smem_bytes, st);
    if (r) {
      mpg::set_error("conv nfold launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(r)));
      return r;
    }
    return MPG_OK;
  }
  if (p->kind == 5 || p->kind == 6) {
    const void* xs[2] = {x0, x1};
    for (int s = 0; s < d.nseg; ++s) {
      if (p->tm_x_ptr[s] == xs[s]) continue;
      MPG_CHECK_ARG((reinterpret_cast<uintptr_t>(xs[s]) & 15) == 0, "conv: input %d not 16-byte aligned", s);
      const uint64_t cs = static_cast<uint64_t>(d.seg_cstride[s]) * 2;
      const uint64_t dims[4] = {static_cast<uint64_t>(d.seg_cin[s]), static_cast<uint64_t>(d.w),
                                static_cast<uint64_t>(d.h), static_cast<uint64_t>(d.n)};
      const uint64_t strides[3] = {cs, cs * d.w, cs * d.w * d.h};
      // one staged image row of the strip (the shortcut input uses the same window, centre pixel shift)
      const uint32_t box[4] = {static_cast<uint32_t>(p->ck), static_cast<uint32_t>(mpg::kVfStrip + d.seg_ksize[0] - 1), 1u, 1u};
      int r = mpg::encode_tmap(p->h, &p->tm_x[s], d.in_dtype == MPG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, xs[s], dims, strides,
                               box, swizzle_for(p->ck));
      if (r) return r;
      p->tm_x_ptr[s] = xs[s];
    }
    if (d.out_dtype != MPG_F32)
      MPG_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "conv: output not 16-byte aligned");
    if (p->kind == 6) {
      mpg::VringParams q = p->rp;
      q.out = y;
      q.resid = residual;
      int r = mpg::vring_launch(p->ck, p->vf_nchw, p->tm_x[0], d.nseg > 1 ? p->tm_x[1] : p->tm_x[0], q, p->grid, p->smem_bytes, st);
      if (r) {
        mpg::set_error("conv vring launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(r)));
        return r;
      }
      return MPG_OK;
    }
    mpg::VfoldParams q = p->vp;
    q.out = y;
    q.resid = residual;
    int r = mpg::vfold_launch(p->ck, p->vf_nchw, p->tm_x[0], d.nseg > 1 ? p->tm_x[1] : p->tm_x[0], q, p->grid, p->smem_bytes, st);
    if (r) {
      mpg::set_error("conv vfold launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(r)));
      return r;
    }
    return MPG_OK;
  }
  if (p->kind == 4) {
    MPG_CHECK_ARG((reinterpret_cast<uintptr_t>(x0) & 15) == 0 && (reinterpret_cast<uintptr_t>(x1) & 15) == 0 &&
                      (d.out_dtype == MPG_F32 || (reinterpret_cast<uintptr_t>(y) & 15) == 0),
                  "conv: tensors not 16-byte aligned");
    mpg::TinyParams t = p->tp;
    t.x0 = x0;
    t.x1 = x1;
    t.out = y;
    int r = mpg::tiny_launch(d, t, st);
    if (r) {
      mpg::set_error("conv tiny launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(r)));
      return r;
    }
    return MPG_OK;
  }
  mpg::DirectParams dp = p->dp;
  dp.x[0] = x0;
  dp.x[1] = x1;
  dp.wts = p->d_wdirect;
  dp.out = y;
  int r = mpg::direct_launch(dp, st);
  if (r) {
    mpg::set_error("conv direct launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(r)));
    return r;
  }
  return MPG_OK;
}

/* Training: refresh the packed weights / shift of a tcgen05 (kind 1) plan from DEVICE fp32 tensors, stream ordered.
 * mode 0: w_seg is this conv's HWIO weight; mode 1: w_seg is the HWIO weight of the FORWARD conv whose input
 * gradient this plan computes (flipped taps, swapped channels). shift_dev may be NULL (= keep). */
int mpg_conv_plan_update(mpg_conv_plan p, const float* w_seg0_dev, const float* w_seg1_dev, int mode0, int mode1,
                         const float* shift_dev, void* stream) {
  return mpg_conv_plan_update_ex(p, w_seg0_dev, w_seg1_dev, mode0, mode1, shift_dev, 0, 0, 0, stream);
}

/* ... with segment 0 read from a SMALLER and / or WIDER source tensor: src_k x src_k taps embedded in the plan's k x k kernel
 * at offset k - src_k (a 4x4 TF-SAME conv is taps -1..+2 of a 5x5 one: the discriminator's k = 4 convs on the tensor cores),
 * and, in mode 0, a plan that covers output channels [cout_off, cout_off + cout) of a source with src_cout channels (cout >
 * 128 split over several plans). 0 = same as the plan. */
int mpg_conv_plan_update_ex(mpg_conv_plan p, const float* w_seg0_dev, const float* w_seg1_dev, int mode0, int mode1,
                            const float* shift_dev, int src_k, int src_cout, int cout_off, void* stream) {
  MPG_CHECK_ARG(p && w_seg0_dev, "mpg_conv_plan_update: null argument");
  MPG_CHECK_ARG(src_k >= 0 && src_k <= p->d.seg_ksize[0] && cout_off >= 0 && (src_cout == 0 || src_cout >= cout_off + p->d.cout),
                "mpg_conv_plan_update_ex: src_k=%d src_cout=%d cout_off=%d do not fit the plan", src_k, src_cout, cout_off);
  MPG_CHECK_ARG(p->kind == 1, "mpg_conv_plan_update: only tcgen05 igemm plans (force_kind 1) can be refreshed on the device");
  MPG_CHECK_ARG(p->d.nseg == 1 || w_seg1_dev, "mpg_conv_plan_update: segment 1 weights missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RepackArgs a;
  memset(&a, 0, sizeof(a));
  a.w[0] = w_seg0_dev;
  a.w[1] = w_seg1_dev;
  a.mode[0] = mode0;
  a.mode[1] = mode1;
  int kt = 0;
  for (int s = 0; s < p->d.nseg; ++s) {
    a.ks[s] = p->d.seg_ksize[s];
    a.cin[s] = p->d.seg_cin[s];
    a.nchunk[s] = p->seg_nchunk[s];
    a.kt0[s] = kt;
    kt += p->seg_nchunk[s] * a.ks[s] * a.ks[s];
  }
  a.src_k = src_k > 0 ? src_k : p->d.seg_ksize[0];
  a.src_cout = src_cout > 0 ? src_cout : p->d.cout;
  a.cout_off = cout_off;
  a.nseg = p->d.nseg;
  a.ck = p->ck;
  a.npad = p->npad;
  a.cout = p->d.cout;
  a.f16 = p->d.in_dtype == MPG_F16 ? 1 : 0;
  a.out = static_cast<uint16_t*>(p->d_wpacked);
  a.total = static_cast<long long>(kt) * p->npad * p->ck;
  long long blocks = (a.total + 255) / 256;
  if (blocks > p->h->sm_count * 8) blocks = p->h->sm_count * 8;
  repack_igemm_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(a);
  MPG_CUDA(cudaGetLastError());
  if (shift_dev) MPG_CUDA(cudaMemcpyAsync(p->d_shift, shift_dev, sizeof(float) * p->d.cout, cudaMemcpyDeviceToDevice, st));
  return MPG_OK;
}

int mpg_conv_plan_destroy(mpg_conv_plan p) {
  if (!p) return MPG_OK;
  if (p->d_wpacked) cudaFree(p->d_wpacked);
  if (p->d_shift) cudaFree(p->d_shift);
  if (p->d_wdirect) cudaFree(p->d_wdirect);
  if (p->d_side_w) cudaFree(p->d_side_w);
  delete p;
  return MPG_OK;
}

int mpg_conv_plan_kind(mpg_conv_plan p) { return p ? p->kind : MPG_EINVAL; }
double mpg_conv_plan_flops(mpg_conv_plan p) { return p ? p->flops : 0.0; }

}  // extern "C"
