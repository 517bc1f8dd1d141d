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
  int force_kind;     /* 0 auto, 1 tcgen05 implicit GEMM, 2 CUDA-core direct               */
  int in_dtype;       /* MPG_BF16 (tensor-core path) or MPG_F32 (fp32 CUDA-core path)      */
  int out_dtype;      /* MPG_BF16 or MPG_F32                                               */
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
int mpg_conv_plan_destroy(mpg_conv_plan p);
/* which kernel the plan dispatches to: 1 = tcgen05 implicit GEMM, 2 = CUDA-core direct */
int mpg_conv_plan_kind(mpg_conv_plan p);
/* algorithmic FLOPs of one run (2*MAC, un-padded channels) */
double mpg_conv_plan_flops(mpg_conv_plan p);

#ifdef __cplusplus
}
#endif
#endif /* MPG_H_ */
