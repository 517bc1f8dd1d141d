// Tile cut / overlap-crop stitch: TileCreator.createTiles / cutTile / concatTiles of
// tools_wscale/tilecreator_t.py:403-450,886-918 for 2-D slices (the z extent of a slice is 1).
// Pure data movement, HBM-bandwidth bound: one 4-byte (or 2-byte) word per thread, coalesced along (x, c).
#include "common.h"

namespace mpg {
namespace {

struct TileArgs {
  const void* in;
  void* out;
  int n, h, w;        // frames
  int th, tw;         // tile size (without padding)
  int sy, sx;         // strides between tile origins (< tile size => overlapping tiles)
  int ty, tx;         // tiles per frame
  int pad;            // cut: np.pad(tile, pad, 'edge'); stitch: tileBorder cropped from every side
  int row_words;      // words per pixel (channels * elem_bytes / word_bytes)
};

template <typename WordT>
__global__ void __launch_bounds__(256) tiles_cut_kernel(const TileArgs a) {
  const int oh = a.th + 2 * a.pad, ow = a.tw + 2 * a.pad;
  const long long total = static_cast<long long>(a.n) * a.ty * a.tx * oh * ow * a.row_words;
  const WordT* in = reinterpret_cast<const WordT*>(a.in);
  WordT* out = reinterpret_cast<WordT*>(a.out);
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(e % a.row_words);
    long long r = e / a.row_words;
    const int x = static_cast<int>(r % ow);
    r /= ow;
    const int y = static_cast<int>(r % oh);
    r /= oh;
    const int ix = static_cast<int>(r % a.tx);
    r /= a.tx;
    const int iy = static_cast<int>(r % a.ty);
    const int n = static_cast<int>(r / a.ty);
    // 'edge' padding replicates the TILE's own border pixels (np.pad on the cut tile, tilecreator_t.py:430)
    int yy = y - a.pad, xx = x - a.pad;
    yy = yy < 0 ? 0 : (yy >= a.th ? a.th - 1 : yy);
    xx = xx < 0 ? 0 : (xx >= a.tw ? a.tw - 1 : xx);
    const int gy = iy * a.sy + yy, gx = ix * a.sx + xx;
    out[e] = in[((static_cast<long long>(n) * a.h + gy) * a.w + gx) * a.row_words + c];
  }
}

template <typename WordT>
__global__ void __launch_bounds__(256) tiles_stitch_kernel(const TileArgs a) {
  const int ch = a.th - 2 * a.pad, cw = a.tw - 2 * a.pad;  // kept centre of every tile
  const int fh = a.ty * ch, fw = a.tx * cw;
  const long long total = static_cast<long long>(a.n) * fh * fw * a.row_words;
  const WordT* in = reinterpret_cast<const WordT*>(a.in);
  WordT* out = reinterpret_cast<WordT*>(a.out);
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(e % a.row_words);
    long long r = e / a.row_words;
    const int x = static_cast<int>(r % fw);
    r /= fw;
    const int y = static_cast<int>(r % fh);
    const int n = static_cast<int>(r / fh);
    const int iy = y / ch, ix = x / cw;
    const int yy = y - iy * ch + a.pad, xx = x - ix * cw + a.pad;
    const long long tile = (static_cast<long long>(n) * a.ty + iy) * a.tx + ix;
    out[e] = in[((tile * a.th + yy) * a.tw + xx) * a.row_words + c];
  }
}

// Exact inverse of an overlapped cut (stride = tile - 2*border, no padding) of a frame of size
// t*(tile - 2*border) + 2*border: every tile contributes its centre, tiles on a frame edge also keep their outer
// border (there the tile edge IS the frame edge, so a network applied per tile saw the same zero padding as one
// applied to the whole frame). This is what makes a tiled apply bit-identical to the untiled one when
// border >= receptive-field radius; concatTiles (tilecreator_t.py:886-918) drops that band.
template <typename WordT>
__global__ void __launch_bounds__(256) tiles_stitch_overlap_kernel(const TileArgs a) {
  const int ch = a.th - 2 * a.pad, cw = a.tw - 2 * a.pad;
  const int fh = a.ty * ch + 2 * a.pad, fw = a.tx * cw + 2 * a.pad;
  const long long total = static_cast<long long>(a.n) * fh * fw * a.row_words;
  const WordT* in = reinterpret_cast<const WordT*>(a.in);
  WordT* out = reinterpret_cast<WordT*>(a.out);
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(e % a.row_words);
    long long r = e / a.row_words;
    const int x = static_cast<int>(r % fw);
    r /= fw;
    const int y = static_cast<int>(r % fh);
    const int n = static_cast<int>(r / fh);
    int iy = (y - a.pad) / ch, ix = (x - a.pad) / cw;
    iy = (y < a.pad) ? 0 : (iy >= a.ty ? a.ty - 1 : iy);
    ix = (x < a.pad) ? 0 : (ix >= a.tx ? a.tx - 1 : ix);
    const int yy = y - iy * ch, xx = x - ix * cw;
    const long long tile = (static_cast<long long>(n) * a.ty + iy) * a.tx + ix;
    out[e] = in[((tile * a.th + yy) * a.tw + xx) * a.row_words + c];
  }
}

inline int grid_of(long long total, int sm) {
  long long b = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm) * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace
}  // namespace mpg

using namespace mpg;

extern "C" {

int mpg_tiles_count(int extent, int tile, int stride) { return (extent - tile) / (stride > 0 ? stride : tile) + 1; }

int mpg_tiles_cut(mpg_handle h, const void* in, void* out, int n, int hh, int ww, int c, int elem_bytes, int th, int tw,
                  int stride_y, int stride_x, int pad, void* stream) {
  MPG_CHECK_ARG(h && in && out && n > 0 && hh > 0 && ww > 0 && c > 0 && th > 0 && tw > 0 && pad >= 0, "mpg_tiles_cut: bad argument");
  MPG_CHECK_ARG(elem_bytes == 2 || elem_bytes == 4, "mpg_tiles_cut: elem_bytes must be 2 or 4");
  MPG_CHECK_ARG(th <= hh && tw <= ww, "mpg_tiles_cut: tile %dx%d larger than frame %dx%d", th, tw, hh, ww);
  TileArgs a;
  a.in = in;
  a.out = out;
  a.n = n;
  a.h = hh;
  a.w = ww;
  a.th = th;
  a.tw = tw;
  a.sy = stride_y > 0 ? stride_y : th;  // strides <= 0: regular non-overlapping grid (tilecreator_t.py:412-414)
  a.sx = stride_x > 0 ? stride_x : tw;
  a.ty = (hh - th) / a.sy + 1;
  a.tx = (ww - tw) / a.sx + 1;
  a.pad = pad;
  const int row_bytes = c * elem_bytes;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (row_bytes % 4 == 0) {
    a.row_words = row_bytes / 4;
    const long long total = static_cast<long long>(n) * a.ty * a.tx * (th + 2 * pad) * (tw + 2 * pad) * a.row_words;
    tiles_cut_kernel<uint32_t><<<grid_of(total, h->sm_count), 256, 0, st>>>(a);
  } else {
    a.row_words = row_bytes / 2;
    const long long total = static_cast<long long>(n) * a.ty * a.tx * (th + 2 * pad) * (tw + 2 * pad) * a.row_words;
    tiles_cut_kernel<uint16_t><<<grid_of(total, h->sm_count), 256, 0, st>>>(a);
  }
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

int mpg_tiles_stitch(mpg_handle h, const void* tiles, void* out, int n, int ty, int tx, int th, int tw, int c,
                     int elem_bytes, int border, void* stream) {
  MPG_CHECK_ARG(h && tiles && out && n > 0 && ty > 0 && tx > 0 && th > 0 && tw > 0 && c > 0 && border >= 0,
                "mpg_tiles_stitch: bad argument");
  MPG_CHECK_ARG(elem_bytes == 2 || elem_bytes == 4, "mpg_tiles_stitch: elem_bytes must be 2 or 4");
  MPG_CHECK_ARG(th > 2 * border && tw > 2 * border, "mpg_tiles_stitch: border %d leaves nothing of a %dx%d tile", border, th, tw);
  TileArgs a;
  a.in = tiles;
  a.out = out;
  a.n = n;
  a.h = a.w = 0;
  a.th = th;
  a.tw = tw;
  a.sy = a.sx = 0;
  a.ty = ty;
  a.tx = tx;
  a.pad = border;
  const int row_bytes = c * elem_bytes;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (row_bytes % 4 == 0) {
    a.row_words = row_bytes / 4;
    const long long total = static_cast<long long>(n) * ty * (th - 2 * border) * tx * (tw - 2 * border) * a.row_words;
    tiles_stitch_kernel<uint32_t><<<grid_of(total, h->sm_count), 256, 0, st>>>(a);
  } else {
    a.row_words = row_bytes / 2;
    const long long total = static_cast<long long>(n) * ty * (th - 2 * border) * tx * (tw - 2 * border) * a.row_words;
    tiles_stitch_kernel<uint16_t><<<grid_of(total, h->sm_count), 256, 0, st>>>(a);
  }
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

/* out[n, ty*(th-2b)+2b, tx*(tw-2b)+2b, c]: overlap-crop stitch that keeps the outer border of frame-edge tiles
 * (inverse of mpg_tiles_cut with stride = tile - 2*border, pad = 0). */
int mpg_tiles_stitch_overlap(mpg_handle h, const void* tiles, void* out, int n, int ty, int tx, int th, int tw, int c,
                             int elem_bytes, int border, void* stream) {
  MPG_CHECK_ARG(h && tiles && out && n > 0 && ty > 0 && tx > 0 && th > 0 && tw > 0 && c > 0 && border >= 0,
                "mpg_tiles_stitch_overlap: bad argument");
  MPG_CHECK_ARG(elem_bytes == 2 || elem_bytes == 4, "mpg_tiles_stitch_overlap: elem_bytes must be 2 or 4");
  MPG_CHECK_ARG(th > 2 * border && tw > 2 * border, "mpg_tiles_stitch_overlap: border %d leaves nothing of a %dx%d tile", border, th, tw);
  TileArgs a;
  a.in = tiles;
  a.out = out;
  a.n = n;
  a.h = a.w = 0;
  a.th = th;
  a.tw = tw;
  a.sy = a.sx = 0;
  a.ty = ty;
  a.tx = tx;
  a.pad = border;
  const int row_bytes = c * elem_bytes;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long px = static_cast<long long>(n) * (ty * (th - 2 * border) + 2 * border) * (tx * (tw - 2 * border) + 2 * border);
  if (row_bytes % 4 == 0) {
    a.row_words = row_bytes / 4;
    tiles_stitch_overlap_kernel<uint32_t><<<grid_of(px * a.row_words, h->sm_count), 256, 0, st>>>(a);
  } else {
    a.row_words = row_bytes / 2;
    tiles_stitch_overlap_kernel<uint16_t><<<grid_of(px * a.row_words, h->sm_count), 256, 0, st>>>(a);
  }
  MPG_CUDA(cudaGetLastError());
  return MPG_OK;
}

}  // extern "C"
