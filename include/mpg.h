/*
 * mpg.h - C ABI of the B200-native (sm_100a) Multi-pass-GAN generator hot path.
 *
 * The reference (maxwerhahn/Multi-pass-GAN) is pure Python/TensorFlow-1.x and has no FFI of
 * its own; every entry point below replaces one TensorFlow / numpy / scipy call site of the
 * hot path (file:line are relative to the reference checkout).  The Python layer API that sits
 * on top of this library (multi-pass-gan_b200/GAN.py) mirrors tools_wscale/GAN.py.
 *
 * Conventions
 *   - plain C symbols, device pointers + cudaStream_t (passed as void*), caller owns every
 *     buffer; the library allocates nothing persistent except opaque handles/plans
 *   - activations are NHWC; the channel stride of bf16 tensors is a multiple of 8 elements
 *   - every function returns 0 on success, a negative MPG_E* code on invalid arguments, or a
 *     positive cudaError_t / CUresult passthrough; mpg_last_error() gives the message of the
 *     last failure on the calling thread; nothing throws or exits across the ABI
 *   - weights/bias/scale arrays handed to *_plan_create are HOST pointers (they are packed
 *     once into the device layout the tensor-core kernel wants)
 */
#ifndef MPG_H_
#define MPG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPG_OK 0
#define MPG_EINVAL (-1)   /* bad shape / dtype / alignment / null pointer */
#define MPG_ENOSUP (-2)   /* combination not supported by any kernel       */
#define MPG_EDRIVER (-3)  /* CUDA driver entry point unavailable            */
#define MPG_ENOMEM (-4)

#define MPG_ACT_NONE 0
#define MPG_ACT_RELU 1  /* tf.nn.relu                     GAN/multipassGAN-out.py:227,233 */
#define MPG_ACT_LRELU 2 /* 0.6*x + 0.4*|x|                tools_wscale/GAN.py:733-737    */
#define MPG_ACT_TANH 3  /* GAN.convolutional_layer default tools_wscale/GAN.py:80        */

#define MPG_BF16 0
#define MPG_F32 1
#define MPG_F16 2 /* IEEE half: same tcgen05 kind::f16 rate as bf16, 3 more mantissa bits */

typedef struct mpg_handle_s* mpg_handle;
typedef struct mpg_conv_plan_s* mpg_conv_plan;

int mpg_version(void);
const char* mpg_last_error(void);

/* One handle per (process, device). Holds the device id, SM count and driver entry points. */
int mpg_create(mpg_handle* out, int device);
int mpg_destroy(mpg_handle h);
int mpg_sm_count(mpg_handle h);

/* ------------------------------------------------------------------------------------------
 * Fused convolution  (replaces tf.nn.conv2d(...,"SAME") tools_wscale/GAN.py:686-691, the bias
 * add :104-105, inference batch-norm :108-110, the activation :112-113, the residual
 * tf.add(B, s)+relu GAN/multipassGAN-out.py:233 / GAN/multipassGAN-4x.py:523, pixel_norm
 * tools_wscale/GAN.py:472-474 and the nearest x2 depool :541 that may follow).
 *
 *   y = pixel_norm?( act( sum_seg conv2d_SAME(x_seg, w_seg * scale_seg[cout]) + shift[cout] ) )
 *
 * n,h,w describe the (virtual, after in_upsample) conv input; output is ceil(h/stride).
 * Segment 1 (optional) is the 1x1 shortcut of a resBlock
 * evaluated as extra K-slabs of the same implicit GEMM. `scale` folds wscale-independent
 * per-output-channel factors (inference BN gamma/sqrt(var+eps)); `shift` is the folded
 * bias/BN offset summed over segments.
 * -----------------------------------------------------------------------------------------*/
typedef struct mpg_conv_desc {
  int n, h, w;        /* input == output spatial size (before `upsample`)                 */
  int nseg;           /* 1 or 2                                                            */
  int seg_cin[2];     /* real input channels of each segment                               */
  int seg_cstride[2]; /* channel stride (elements) of each NHWC input tensor               */
  int seg_ksize[2];   /* 1, 3 or 5 (tensor-core path); any k incl. 4 on the CUDA-core path */
  int cout;           /* real output channels                                              */
  int act;            /* MPG_ACT_*                                                         */
  int pixel_norm;     /* 0/1 ; eps 1e-8 as tools_wscale/GAN.py:472                         */
  int upsample;       /* 1, or 2 = nearest x2 replicated store                             */
  int in_upsample;    /* >=1: the conv reads a nearest-upsampled view of x (CUDA-core path) */
  int stride;         /* 1 (tensor-core path) or 2 (CUDA-core path, discriminator)         */
  int force_kind;     /* 0 auto, 1 tcgen05 implicit GEMM, 2 CUDA-core direct, 3 tap-folded, 4 tiny, 5 / 6 row-streaming */
  int in_dtype;       /* MPG_BF16 / MPG_F16 (tensor-core path) or MPG_F32 (CUDA-core path) */
  int out_dtype;      /* MPG_BF16 / MPG_F16 (same 16-bit type as in_dtype) or MPG_F32      */
  int out_cstride;    /* channel stride (elements) of y; channels >= cout are written as 0 */
} mpg_conv_desc;

/* w_seg: HOST fp32 HWIO [k,k,cin,cout] already multiplied by the wscale constant
 * (tools_wscale/GAN.py:664-668); scale_seg: HOST fp32 [cout] or NULL (=1); shift: HOST fp32
 * [cout] or NULL (=0). */
int mpg_conv_plan_create(mpg_handle h, const mpg_conv_desc* d, const float* w_seg0,
                         const float* w_seg1, const float* scale_seg0, const float* scale_seg1,
                         const float* shift, mpg_conv_plan* out);
int mpg_conv_plan_run(mpg_conv_plan p, const void* x_seg0, const void* x_seg1, void* y,
                      void* stream);
/* Cross-launch shortcut fusion (the 1x1 shortcut conv of the NEXT residual block reads this conv's whole output: for ru3 of
 * gen_resnet, GAN/multipassGAN-4x.py:521,563, that is a 537 MB re-read of the 128-channel tensor per slice batch):
 * mpg_conv_plan_set_side gives a 128-channel tcgen05 plan a SIDE output y_side[n,h,w,8] (fp32) = y x w_side, computed in the
 * epilogue from the fp32 values while they are in registers (w_side: HOST fp32 [128][side_cout], the shortcut's weight with
 * wscale / BN scale folded; side_cout <= 8); the consumer (a tap-folded plan with <= 8 output channels, built WITHOUT the
 * shortcut segment) adds it through `residual` before its activation. MPG_ENOSUP when the plan cannot carry a side output
 * (the caller keeps the two-segment form). */
int mpg_conv_plan_set_side(mpg_conv_plan p, const float* w_side, int side_cout);
int mpg_conv_plan_run_ex(mpg_conv_plan p, const void* x_seg0, const void* x_seg1, void* y, float* y_side,
                         const float* residual, void* stream);
int mpg_conv_plan_destroy(mpg_conv_plan p);
/* Training: refresh the packed weights / shift of a tcgen05 plan (force_kind 1) from DEVICE fp32 HWIO tensors,
 * stream ordered. mode 0: w_seg is this conv's weight; mode 1: w_seg is the weight of the FORWARD conv whose
 * input gradient this plan computes (taps flipped, channels swapped = Conv2DBackpropInput). */
int mpg_conv_plan_update(mpg_conv_plan p, const float* w_seg0_dev, const float* w_seg1_dev, int mode0, int mode1,
                         const float* shift_dev, void* stream);
/* ... with segment 0 read from a smaller and / or wider source tensor: src_k x src_k taps embedded in the plan's k x k kernel at
 * offset k - src_k (TF's SAME window of a 4x4 conv = taps -1..+2 of a 5x5 one: the k = 4 convs of disc_binclass,
 * GAN/multipassGAN-4x.py:593-614, on the tensor cores) and, in mode 0, output channels [cout_off, cout_off + cout) of a source
 * with src_cout channels (cout > 128 split over several plans). 0 = same as the plan. */
int mpg_conv_plan_update_ex(mpg_conv_plan p, const float* w_seg0_dev, const float* w_seg1_dev, int mode0, int mode1,
                            const float* shift_dev, int src_k, int src_cout, int cout_off, void* stream);
/* which kernel the plan dispatches to: 1 = tcgen05 implicit GEMM, 2 = CUDA-core direct, 3 = tcgen05 with the
 * horizontal taps folded into N (narrow Cout), 4 = CUDA-core kernel for cout <= 2 from <= 8 channels,
 * 5 = row-streaming tcgen05 kernel with the vertical taps folded into N (k * round_up(cout, 8) <= 256, wide images),
 * 6 = row-streaming tcgen05 kernel that accumulates the vertical-tap sum in a TMEM ring (cout <= 64) */
int mpg_conv_plan_kind(mpg_conv_plan p);
/* algorithmic FLOPs of one run (2*MAC, un-padded channels) */
double mpg_conv_plan_flops(mpg_conv_plan p);


/* ------------------------------------------------------------------------------------------
 * Fused THIN residual block = the reference's resBlock (GAN/multipassGAN-4x.py:505-526) in one launch, for the
 * channel-poor ends of gen_resnet (:560 ru1 4->8->32, :564 ru4 8->2->1):
 *
 *   y = act( conv_kxk(act(conv_kxk(x, w_a*scale_a) + shift_a), w_b*scale_b) + conv_1x1(x, w_s*scale_s) + shift_bs )
 *
 * x may be the fp32 rows of the slice assembler (<= 4 channels) read through a nearest x`in_upsample` view
 * (max_depool, GAN/multipassGAN-4x.py:553-554) or a 16-bit NHWC tensor (<= 8 channels, cstride % 8 == 0); the
 * intermediate never leaves shared memory. Runs on register-level tensor-core MMAs (mma.sync m16n8k16, fp32
 * accumulate, 16-bit operands of type mma_dtype). Supported: ksize 5, cmid <= 8, and either cout == 32 (16-bit output,
 * cstride 32) or cout <= 8 (fp32 output with cstride <= 8, or 16-bit output with cstride 8); anything else
 * returns MPG_ENOSUP and the caller uses separate mpg_conv plans. Weights are HOST fp32 HWIO with the wscale
 * constant folded; scale_* (inference BN gamma/sqrt(var+eps)) may be NULL (= 1), shift_* NULL (= 0);
 * shift_bs is the sum of the folded offsets of conv B and of the shortcut.
 * -----------------------------------------------------------------------------------------*/
typedef struct mpg_resblock_desc {
  int n, h, w;        /* spatial size of the block (after in_upsample)                     */
  int cin, cmid, cout;
  int ksize;          /* 5                                                                 */
  int in_upsample;    /* >= 1: x is [n, h/f, w/f, in_cstride], read through a nearest view */
  int in_dtype;       /* MPG_F32, or the 16-bit mma_dtype                                  */
  int in_cstride;
  int mma_dtype;      /* MPG_F16 / MPG_BF16                                                */
  int out_dtype;      /* mma_dtype or MPG_F32                                              */
  int out_cstride;
  int act;            /* MPG_ACT_NONE / RELU / LRELU, after conv A and after the add       */
} mpg_resblock_desc;
typedef struct mpg_resblock_plan_s* mpg_resblock_plan;
int mpg_resblock_plan_create(mpg_handle h, const mpg_resblock_desc* d, const float* w_a, const float* w_b,
                             const float* w_s, const float* scale_a, const float* scale_b, const float* scale_s,
                             const float* shift_a, const float* shift_bs, mpg_resblock_plan* out);
int mpg_resblock_plan_run(mpg_resblock_plan p, const void* x, void* y, void* stream);
/* same, and *sat_counter_dev (may be NULL) += intermediate values that left the 16-bit range (the intermediate never
 * reaches global memory, so mpg_count_saturated cannot see it) */
int mpg_resblock_plan_run_checked(mpg_resblock_plan p, const void* x, void* y, unsigned long long* sat_counter_dev,
                                  void* stream);
int mpg_resblock_plan_destroy(mpg_resblock_plan p);
double mpg_resblock_plan_flops(mpg_resblock_plan p);

/* ------------------------------------------------------------------------------------------
 * Tensor plumbing around the convolutions (bandwidth bound)
 * -----------------------------------------------------------------------------------------*/
typedef struct mpg_chan_src {
  const void* ptr; /* NHWC device tensor [n, oh/factor_h, ow/factor_w, cstride]           */
  int dtype;       /* MPG_BF16 / MPG_F16 / MPG_F32                                       */
  int cstride;
  int c0, nch;     /* channel range [c0, c0+nch) taken from this source                  */
  int factor_h, factor_w; /* nearest-neighbour replication factors (src = floor(dst/f)) */
} mpg_chan_src;

/* out[n,y,x,:] = concat_s( src_s[n, y/fh_s, x/fw_s, c0_s : c0_s+nch_s] ), zero padded to out_cstride.
 * Replaces tf.image.resize_images(.., method=1) / keras resize_images (tools_wscale/GAN.py:517,541;
 * GAN/multipassGAN-out.py:357,363), tf.concat / tf.slice (GAN/multipassGAN-out.py:330-332,357)
 * and the fp32 -> bf16 cast + channel padding the tensor-core kernels need. */
int mpg_pack_channels(mpg_handle h, const mpg_chan_src* srcs, int nsrc, void* out, int out_dtype,
                      int out_cstride, int n, int oh, int ow, void* stream);

/* TF 1.x legacy bicubic resize (tf.image.resize_images(.., 2), align_corners=False, Keys A=-0.75,
 * 1024-entry table; tools_wscale/GAN.py:541 with mode=2 from GAN/multipassGAN-out.py:330):
 * per-axis tap indices/weights are precomputed once per (in,out) size. */
int mpg_bicubic_plan_create(mpg_handle h, int in_h, int in_w, int out_h, int out_w, void** plan_out);
int mpg_bicubic_plan_destroy(void* plan);

/* Standalone tf.image.resize_images(x, [out_h, out_w], mode) of an NHWC tensor with c channels (tools_wscale/GAN.py:541,
 * GAN.avg_depool): mode 0 = TF1 legacy bilinear (align_corners=False, no half-pixel centres: in = out * in/out,
 * lerp between floor and min(floor+1, size-1)), mode 2 = TF1 legacy bicubic (needs a plan of mpg_bicubic_plan_create).
 * Mode 1 (nearest) is mpg_pack_channels. Channels >= c of `out` are written as 0. */
int mpg_resize_images(mpg_handle h, const void* src, int src_dtype, int src_cstride, int c, int n, int src_h, int src_w,
                      void* out, int out_dtype, int out_cstride, int out_h, int out_w, int mode, void* bicubic_plan,
                      void* stream);

/* Range check (validation mode of the engine): *counter_dev += number of elements of `t` that sit on the 16-bit type's
 * largest finite value or beyond (+-65504 for MPG_F16: every 16-bit store of the kernels is a SATURATING conversion, so
 * an out-of-range activation lands exactly there), are infinite or NaN (MPG_F32: non-finite only). The reference
 * computes in fp32 (tools_wscale/GAN.py:19-35) and cannot overflow this way: a non-zero count means the 16-bit path
 * has left the reference's result and the caller must fall back to precision fp32. */
int mpg_count_saturated(mpg_handle h, const void* t, long long count, int dtype, unsigned long long* counter_dev,
                        void* stream);

/* out[n,y,x] = dens[n,y,x] + R(src[..., src_c])  with R = identity (mode 0, GAN/multipassGAN-out.py:332)
 * or the TF1 bicubic resize (mode 2, GAN/multipassGAN-out.py:330). dens/out fp32 [n,out_h,out_w]. */
int mpg_dens_residual(mpg_handle h, const float* dens, const void* src, int src_dtype, int src_cstride,
                      int src_c, int mode, void* bicubic_plan, int n, int out_h, int out_w, int src_h,
                      int src_w, float* out, void* stream);

/* Density output of the out.py generators in one pass: out[n,y,x] = sum_c x[n,y,x,c] * w[c] + bias + R(src[..., src_c]),
 * i.e. the 1x1 conv g_cdensOut (gain 1, no activation, GAN/multipassGAN-out.py:282) fused with the additive residual of
 * mpg_dens_residual (:327-332). x: NHWC features (MPG_F32, or 16-bit at a channel stride that is a multiple of 8);
 * w_host: HOST fp32 [cin] (wscale folded), cin <= 64; mode -1 = no residual, 0 / 2 as mpg_dens_residual. */
int mpg_dens_out(mpg_handle h, const void* x, int x_dtype, int x_cstride, int cin, const float* w_host, float bias,
                 const void* src, int src_dtype, int src_cstride, int src_c, int mode, void* bicubic_plan, int n,
                 int out_h, int out_w, int src_h, int src_w, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Volume pipeline (replaces the host numpy/scipy code of generate3DUniForNewNetwork,
 * GAN/multipassGAN-out.py:390-618 and GAN/multipassGAN-4x.py:1090-1169)
 * -----------------------------------------------------------------------------------------*/

/* Slice assembler. Builds `count` consecutive network-input slices [count, H, W, out_cstride]
 * starting at slice index `slice0` from the low-res field volume `vol` = fp32 [L0, L1, L2, vol_c].
 *
 * Output index (s, i, j) addresses source axis axis_of[0], axis_of[1], axis_of[2] respectively
 * (a permutation of 0,1,2 = the numpy transposes of GAN/multipassGAN-out.py:402,408,414,466,472,478).
 * Along each output index k the source coordinate is either the index itself (zoom[k] == 1) or the
 * align-corners linear interpolation of scipy.ndimage.zoom(order=1) with factor zoom[k]:
 * coord = o * (n_in-1)/(n_out-1), n_out = n_in*zoom[k] (GAN/multipassGAN-out.py:401-421,
 * GAN/multipassGAN-4x.py:1095-1103).
 * Channel c of the output is  chan_scale[c] * lerp(vol[..., chan_src[c]])  for c < nchan
 * (the velocity-channel swaps of GAN/multipassGAN-out.py:403-405,409-411,473-475 and the
 * velScale / *upRes factors of GAN/multipassGAN-4x.py:277-283).
 * If `dens` != NULL, output channel 0 is instead read from the fp32 volume dens[S_s, H, W]
 * (already laid out in slice order; first-pass density, GAN/multipassGAN-4x.py:1113) and the
 * field channels follow from channel 1; dens row 0 is slice `dens_slice0` (a rank's slab).
 * If add_adj != 0 two more channels are appended: channel chan_src[0] of slice s-1 and s+1 of the
 * interpolated stack, zero outside [0, n_slices) (GAN/multipassGAN-out.py:423-436).
 * Remaining channels up to out_cstride are zero. */
typedef struct mpg_assemble_desc {
  int dims[3];      /* L0, L1, L2 of vol                                          */
  int vol_c;        /* channels of vol                                            */
  int axis_of[3];   /* source axis addressed by output index (slice, row, col)   */
  int zoom[3];      /* integer zoom per OUTPUT index (1 = none)                   */
  int nchan;        /* field channels to emit                                     */
  int chan_src[8];
  float chan_scale[8];
  int add_adj;
  int out_dtype;    /* MPG_BF16 / MPG_F16 / MPG_F32                               */
  int out_cstride;
  int dens_slice0;  /* absolute slice index of dens[0] (0 unless the volume is sharded)  */
} mpg_assemble_desc;

int mpg_slice_assemble(mpg_handle h, const mpg_assemble_desc* d, const float* vol, const float* dens,
                       int slice0, int count, void* out, void* stream);

/* out = permute(in, perm) for a dense fp32 3-D volume in[d0,d1,d2]; out axis k is in axis perm[k]
 * (numpy .transpose(perm); GAN/multipassGAN-out.py:459,521,587-590, GAN/multipassGAN-4x.py:1142-1144).
 * If threshold > 0, values < threshold are written as 0 (GAN/multipassGAN-out.py:612-615). */
int mpg_transpose3d(mpg_handle h, const float* in, float* out, int d0, int d1, int d2, const int perm[3],
                    float threshold, void* stream);

/* in-place v < threshold -> 0 (GAN/multipassGAN-out.py:614-615, GAN/multipassGAN-4x.py:1156-1157) */
/* Multi-GPU axis change between passes as ONE kernel (replaces the np.array(rows).reshape(S,S,S).transpose(..)
 * of GAN/multipassGAN-out.py:459,521 / -4x.py:1142 when the volume is sharded by slice): this rank's slab [S/G,S,S] is
 * transposed and every element is stored directly into the output slab of the rank that owns it in the next pass
 * (peer_out[r] = peer-mapped device pointer of rank r's slab, NVLink / NVSwitch). split_axis 2: received block
 * [A, b, c_loc]; 1: [A, b_loc, c]; final_perm: permutation of the received block (final_perm[2] != 2). The caller
 * brackets the launch with cross-rank barriers. */
int mpg_reslab_p2p(mpg_handle h, const float* slab, void* const* peer_out, int world, int rank, int S, int split_axis,
                   const int final_perm[3], float threshold, void* stream);
/* Same exchange for `count` consecutive rows [a0, a0+count) of the old slice axis only (`part` points at row a0 of the
 * rank's slab): lets the caller push every finished slice batch to its owners while the pass is still running, so the
 * axis change leaves the critical path and ONE cross-rank barrier per pass boundary remains (GAN/multipassGAN-out.py:
 * 443-459: the rows of a slice batch are final when its sess.run returns). */
int mpg_reslab_p2p_part(mpg_handle h, const float* part, void* const* peer_out, int world, int S, int a0, int count,
                        int split_axis, const int final_perm[3], float threshold, void* stream);
int mpg_threshold(mpg_handle h, float* vol, long long count, float threshold, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tile cut / overlap-crop stitch of 2-D slices (tools_wscale/tilecreator_t.py: createTiles :403-434,
 * cutTile :436-450, concatTiles :886-918).  Tiles are ordered (frame, tile row, tile column); NHWC, any
 * 2- or 4-byte element type.
 * -----------------------------------------------------------------------------------------*/
/* number of tiles along one axis: (extent - tile) // stride + 1  (stride <= 0 means stride = tile) */
int mpg_tiles_count(int extent, int tile, int stride);
/* out[n*ty*tx, th+2*pad, tw+2*pad, c]: tile (iy,ix) starts at (iy*stride_y, ix*stride_x); stride < tile gives
 * overlapping tiles; pad > 0 replicates the tile's own edge pixels (np.pad(tile, 'edge')). */
int mpg_tiles_cut(mpg_handle h, const void* in, void* out, int n, int hh, int ww, int c, int elem_bytes, int th,
                  int tw, int stride_y, int stride_x, int pad, void* stream);
/* out[n, ty*(th-2*border), tx*(tw-2*border), c]: crop `border` from every side of every tile and concatenate
 * x, then y (concatTiles with tileBorder). */
int mpg_tiles_stitch(mpg_handle h, const void* tiles, void* out, int n, int ty, int tx, int th, int tw, int c,
                     int elem_bytes, int border, void* stream);
/* Overlap-crop stitch that also keeps the outer border of frame-edge tiles: exact inverse of mpg_tiles_cut with
 * stride = tile - 2*border, pad = 0 on a frame of t*(tile - 2*border) + 2*border pixels per axis. With
 * border >= the generator's receptive-field radius a per-tile apply is bit-identical to the
 * whole-slice apply; TileCreator.concatTiles (tools_wscale/tilecreator_t.py:886-918) drops that band instead. */
int mpg_tiles_stitch_overlap(mpg_handle h, const void* tiles, void* out, int n, int ty, int tx, int th, int tw, int c,
                             int elem_bytes, int border, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training step of the 4x model (generator + spatial discriminator, GAN/multipassGAN-4x.py:528-620,
 * 744-768, 889-902, loop :1316-1397).  All tensors fp32 NHWC, weights HWIO fp32 in DEVICE memory
 * (Adam rewrites them every step).  `scratch` is caller-owned device memory of the stated size.
 * -----------------------------------------------------------------------------------------*/

/* y = conv2d_SAME(nearest_up(x, in_up), w, stride) + bias   (tools_wscale/GAN.py:94,105; max_depool :517) */
int mpg_train_conv_fwd(mpg_handle h, const float* x, const float* w, const float* bias, float* y, int n, int hh,
                       int ww, int cin, int cout, int k, int stride, int in_up, void* stream);
/* dx (+)= d loss / d x of that convolution (tf.gradients -> Conv2DBackpropInput) */
int mpg_train_conv_dgrad(mpg_handle h, const float* dy, const float* w, float* dx, int n, int hh, int ww, int cin,
                         int cout, int k, int stride, int accumulate, void* stream);
/* dw += Conv2DBackpropFilter, dbias += sum dy (dbias may be NULL); scratch: cout doubles */
int mpg_train_conv_wgrad(mpg_handle h, const float* x, const float* dy, float* dw, float* dbias, double* scratch,
                         int n, int hh, int ww, int cin, int cout, int k, int stride, int in_up, void* stream);
/* Tensor-core filter gradient (csrc/conv_wgrad_tc.cu): dw[k,k,cin,cout] (fp32 HWIO) += Conv2DBackpropFilter(x, dy)
 * of a stride-1 SAME conv, x / dy bf16 NHWC with channel stride == channel count. Supported: k in {1,3,5},
 * w % 16 == 0, one of (cin, cout) == 128 and the other in {32, 64, 128}; other shapes return MPG_ENOSUP
 * (use mpg_train_conv_wgrad). Replaces the backward of tf.nn.conv2d (tools_wscale/GAN.py:691) inside
 * tf.train.AdamOptimizer.minimize (GAN/multipassGAN-4x.py:889-898). */
int mpg_train_conv_wgrad_tc(mpg_handle h, const void* x, const void* dy, float* dw, int n, int hh, int ww, int cin,
                            int cout, int k, void* stream);
/* dbias[cout] += column sums of dy[rows, cout]; scratch: >= cout doubles */
int mpg_train_bias_grad(mpg_handle h, const float* dy, float* dbias, double* scratch, long long rows, int cout,
                        void* stream);
/* tf.contrib.layers.batch_norm(is_training=True) (+ activation) and its moving-average update
 * (tools_wscale/GAN.py:110; UPDATE_OPS GAN/multipassGAN-4x.py:776-779). scratch: 2*c doubles */
int mpg_train_bn_fwd(mpg_handle h, const float* x, const float* gamma, const float* beta, float* y, float* mean,
                     float* var, float* invstd, float* moving_mean, float* moving_var, double* scratch,
                     long long rows, int c, float eps, float decay, int act, void* stream);
/* backward: dy = gradient w.r.t. the activated output y; dx = gradient w.r.t. x; dgamma / dbeta accumulate. `dz` is
 * unused (kept for ABI stability, may be NULL): dy * act'(y) is formed inside the statistics and the dx pass */
int mpg_train_bn_bwd(mpg_handle h, const float* x, const float* y, const float* dy, const float* gamma,
                     const float* mean, const float* invstd, float* dz, float* dx, float* dgamma, float* dbeta,
                     double* scratch, long long rows, int c, int act, void* stream);
int mpg_train_act_fwd(mpg_handle h, const float* x, float* y, long long count, int act, void* stream);
int mpg_train_add_act_fwd(mpg_handle h, const float* a, const float* b, float* y, long long count, int act,
                          void* stream); /* relu(tf.add(B, s)) GAN/multipassGAN-4x.py:523 */
int mpg_train_act_bwd(mpg_handle h, const float* y, const float* dy, float* dz, long long count, int act,
                      void* stream);
int mpg_train_axpy(mpg_handle h, float* y, const float* x, float alpha, long long count, void* stream);
/* out = a * b: run-time weight scaling W_eff = v * wscale (tools_wscale/GAN.py:664-668) and its gradient */
int mpg_train_mul(mpg_handle h, float* out, const float* a, const float* b, long long count, void* stream);
/* losses (GAN/multipassGAN-4x.py:751-768): *loss (device double) += scale * value, gradient (+)= into d* */
int mpg_train_bce_logits(mpg_handle h, const float* logits, float label, float scale, double* loss,
                         float* dlogits, long long count, int accumulate, void* stream);
int mpg_train_l1_mean(mpg_handle h, const float* y, const float* g, float scale, double* loss, float* dg,
                      long long count, int accumulate, void* stream);
int mpg_train_l2_half(mpg_handle h, const float* a, const float* b, float scale, double* loss, float* db,
                      long long count, int accumulate, void* stream);
/* tf.train.AdamOptimizer update, TF1 "epsilon hat" form (GAN/multipassGAN-4x.py:889-898), one flat launch */
int mpg_train_adam(mpg_handle h, float* param, const float* grad, float* m, float* v, long long count, float lr_t,
                   float beta1, float beta2, float eps, void* stream);
/* GAN.fully_connected_layer with one output (tools_wscale/GAN.py:438-456; d_l5) */
/* same, lr_t read from a device scalar (so a captured CUDA graph of the step can be replayed) */
int mpg_train_adam_dev(mpg_handle h, float* param, const float* grad, float* m, float* v, long long count,
                       const float* lr_t_dev, float beta1, float beta2, float eps, void* stream);
int mpg_train_fc_fwd(mpg_handle h, const float* x, const float* w, const float* bias, float* y, int rows, int nin,
                     void* stream);
int mpg_train_fc_bwd(mpg_handle h, const float* x, const float* w, const float* dy, float* dx, float* dw,
                     float* dbias, int rows, int nin, void* stream);
/* tensorResample of the 8x trainer (GAN/multipassGAN-8x.py:545-594; used on the frame triplets in front of the temporal
 * critic, :1195-1197, 1241-1242): out[n,hh,ww,c] = value re-sampled bilinearly at pos[n,hh,ww,2] - 0.5 (pos[...,0] along hh),
 * no index clamping, out-of-range cells contribute 0; _bwd scatter-adds dout into dvalue (zero it first) */
int mpg_train_resample_fwd(mpg_handle h, const float* value, const float* pos, float* out, int n, int hh, int ww, int c,
                           void* stream);
/* the positions it reads: getSemiLagrPosBatch of getTempoinput (tools_wscale/tilecreator_t.py:1341-1378, called from
 * selectRandomTempoTiles :1382-1413) for 2-D tiles. x [n, L, L, cstride] low-res tile rows ordered (sample, frame), (vx, vy) at
 * channels c0, c0 + 1; pos [n, S, S, 2] (y, x) = cell centre - (velocity interpolated to S x S, MAC-centred, * S / L) * dt,
 * dt of row r = dt0 * (n_t / 2 - r % n_t) */
int mpg_train_semilagr_pos(mpg_handle h, const float* x, float* pos, int n, int L, int S, int cstride, int c0, float dt0,
                           int n_t, void* stream);
int mpg_train_resample_bwd(mpg_handle h, const float* dout, const float* pos, float* dvalue, int n, int hh, int ww, int c,
                           void* stream);
int mpg_train_take_channel(mpg_handle h, const float* in, float* out, long long npix, int cstride, int c,
                           int accumulate, void* stream);

/* Pieces of the 8x progressive-growing trainer (GAN/multipassGAN-8x.py): 2x2 average pooling of growBlockDisc
 * (:771-772 -> tools_wscale/GAN.py:162-169; hh, ww = INPUT size), lerp blending of the growing stages (:596-597, t already
 * clipped), y = alpha x, the WGAN-GP gradient penalty (:1130-1133: g [rows, n] -> *loss += mean_b lambda (|g_b + 1e-4| -
 * target)^2, v = d penalty / d g, optional per-sample norms) and the critic terms scale * mean(x^power) (:1111-1112, 1140). */
int mpg_train_avgpool2_fwd(mpg_handle h, const float* x, float* y, int n, int hh, int ww, int c, void* stream);
int mpg_train_avgpool2_bwd(mpg_handle h, const float* dy, float* dx, int n, int hh, int ww, int c, int accumulate,
                           void* stream);
int mpg_train_lerp(mpg_handle h, float* out, const float* a, const float* b, float t, long long count, void* stream);
int mpg_train_scale(mpg_handle h, float* y, const float* x, float alpha, long long count, void* stream);
int mpg_train_gp_penalty(mpg_handle h, const float* g, float* v, double* loss, float* norms, int rows, long long n,
                         float lambda, float target, void* stream);
int mpg_train_mean_pow(mpg_handle h, const float* x, float scale, int power, double* loss, float* dx, long long count,
                       int accumulate, void* stream);
/* Strided convs through a stride-1 tensor-core plan: out[n,oy,ox,out_c0 + c] = in[n, oy*stride, ox*stride, c] (pick), and its
 * adjoint for the input gradient: a 16-bit full-resolution tensor that holds dy at the sampled positions and zeros elsewhere
 * (channels >= c of the cstride are zero too) */
int mpg_train_pick(mpg_handle h, const float* in, float* out, int n, int oh, int ow, int c, int stride, int in_cstride,
                   int out_cstride, int out_c0, void* stream);
int mpg_train_stuff16(mpg_handle h, const float* dy, void* out16, int out_dtype, int n, int oh, int ow, int c, int stride,
                      int out_cstride, void* stream);
/* pixel_norm (tools_wscale/GAN.py:472-474) of the growing generator in training mode, x [rows, c], and its backward */
int mpg_train_pixel_norm_fwd(mpg_handle h, const float* x, float* y, long long rows, int c, void* stream);
int mpg_train_pixel_norm_bwd(mpg_handle h, const float* x, const float* dy, float* dx, long long rows, int c, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPG_H_ */
