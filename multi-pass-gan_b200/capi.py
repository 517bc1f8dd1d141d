"""ctypes binding of libmpg_b200.so (the C ABI declared in include/mpg.h).

There is no CPU fallback: importing this module without the built library, or creating a handle
without a B200, raises.  PyTorch is used only to own device memory and streams.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpg_b200.so")

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
BF16, F32, F16 = 0, 1, 2
KIND_VRING = 6  # tcgen05, row streaming, vertical-tap sum accumulated in a TMEM ring of output rows
KIND_VFOLD = 5  # tcgen05, image rows streamed through 128-pixel strips, vertical taps folded into N
KIND_TCGEN05, KIND_DIRECT, KIND_NFOLD, KIND_TINY = 1, 2, 3, 4  # NFOLD: tcgen05 with the horizontal taps folded into N; TINY: CUDA-core kernel for cout <= 2 / cin <= 8

_ACT_BY_NAME = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "lrelu": ACT_LRELU, "tanh": ACT_TANH}


class MpgError(RuntimeError):
    """Raised for any non-zero status of the C ABI (mirrors TF's InvalidArgumentError role)."""


class MpgRangeError(MpgError):
    """A 16-bit activation saturated (range_check mode): the fp16 path would silently clip where the fp32 reference
    does not. `.counts` = {layer label: saturated elements}."""

    def __init__(self, counts):
        self.counts = dict(counts)
        worst = sorted(self.counts.items(), key=lambda kv: -kv[1])[:4]
        super().__init__("16-bit activations saturated (|x| >= max finite) in %d layer(s): %s -- rerun with precision fp32"
                         % (len(self.counts), "; ".join("%s: %d" % kv for kv in worst)))


class ConvDesc(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int), ("h", ctypes.c_int), ("w", ctypes.c_int),
        ("nseg", ctypes.c_int),
        ("seg_cin", ctypes.c_int * 2),
        ("seg_cstride", ctypes.c_int * 2),
        ("seg_ksize", ctypes.c_int * 2),
        ("cout", ctypes.c_int),
        ("act", ctypes.c_int),
        ("pixel_norm", ctypes.c_int),
        ("upsample", ctypes.c_int),
        ("in_upsample", ctypes.c_int),
        ("stride", ctypes.c_int),
        ("force_kind", ctypes.c_int),
        ("in_dtype", ctypes.c_int),
        ("out_dtype", ctypes.c_int),
        ("out_cstride", ctypes.c_int),
    ]


class ResblockDesc(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("h", ctypes.c_int), ("w", ctypes.c_int),
                ("cin", ctypes.c_int), ("cmid", ctypes.c_int), ("cout", ctypes.c_int), ("ksize", ctypes.c_int),
                ("in_upsample", ctypes.c_int), ("in_dtype", ctypes.c_int), ("in_cstride", ctypes.c_int),
                ("mma_dtype", ctypes.c_int), ("out_dtype", ctypes.c_int), ("out_cstride", ctypes.c_int),
                ("act", ctypes.c_int)]


class ChanSrc(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("dtype", ctypes.c_int), ("cstride", ctypes.c_int),
                ("c0", ctypes.c_int), ("nch", ctypes.c_int), ("factor_h", ctypes.c_int),
                ("factor_w", ctypes.c_int)]


class AssembleDesc(ctypes.Structure):
    _fields_ = [("dims", ctypes.c_int * 3), ("vol_c", ctypes.c_int), ("axis_of", ctypes.c_int * 3),
                ("zoom", ctypes.c_int * 3), ("nchan", ctypes.c_int), ("chan_src", ctypes.c_int * 8),
                ("chan_scale", ctypes.c_float * 8), ("add_adj", ctypes.c_int), ("out_dtype", ctypes.c_int),
                ("out_cstride", ctypes.c_int), ("dens_slice0", ctypes.c_int)]


_lib = None


def lib():
    """Load the shared library once; fail loudly if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpgError(
            "libmpg_b200.so is missing (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback for the CUDA path." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, ip, dp = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    fp = ctypes.POINTER(ctypes.c_float)
    L.mpg_version.restype = ip
    L.mpg_last_error.restype = ctypes.c_char_p
    L.mpg_create.argtypes = [ctypes.POINTER(vp), ip]
    L.mpg_destroy.argtypes = [vp]
    L.mpg_sm_count.argtypes = [vp]
    L.mpg_conv_plan_create.argtypes = [vp, ctypes.POINTER(ConvDesc), fp, fp, fp, fp, fp, ctypes.POINTER(vp)]
    L.mpg_conv_plan_run.argtypes = [vp, vp, vp, vp, vp]
    L.mpg_conv_plan_set_side.argtypes = [vp, fp, ip]
    L.mpg_conv_plan_run_ex.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.mpg_conv_plan_destroy.argtypes = [vp]
    L.mpg_conv_plan_update.argtypes = [vp, vp, vp, ip, ip, vp, vp]
    L.mpg_conv_plan_update_ex.argtypes = [vp, vp, vp, ip, ip, vp, ip, ip, ip, vp]
    L.mpg_conv_plan_kind.argtypes = [vp]
    L.mpg_conv_plan_flops.argtypes = [vp]
    L.mpg_conv_plan_flops.restype = dp
    L.mpg_resblock_plan_create.argtypes = [vp, ctypes.POINTER(ResblockDesc), fp, fp, fp, fp, fp, fp, fp, fp, ctypes.POINTER(vp)]
    L.mpg_resblock_plan_run.argtypes = [vp, vp, vp, vp]
    L.mpg_resblock_plan_run_checked.argtypes = [vp, vp, vp, vp, vp]
    L.mpg_resblock_plan_destroy.argtypes = [vp]
    L.mpg_resblock_plan_flops.argtypes = [vp]
    L.mpg_resblock_plan_flops.restype = dp
    L.mpg_pack_channels.argtypes = [vp, ctypes.POINTER(ChanSrc), ip, vp, ip, ip, ip, ip, ip, vp]
    L.mpg_bicubic_plan_create.argtypes = [vp, ip, ip, ip, ip, ctypes.POINTER(vp)]
    L.mpg_bicubic_plan_destroy.argtypes = [vp]
    L.mpg_dens_residual.argtypes = [vp, vp, vp, ip, ip, ip, ip, vp, ip, ip, ip, ip, ip, vp, vp]
    L.mpg_dens_out.argtypes = [vp, vp, ip, ip, ip, fp, ctypes.c_float, vp, ip, ip, ip, ip, vp, ip, ip, ip, ip, ip, vp, vp]
    L.mpg_count_saturated.argtypes = [vp, vp, ctypes.c_longlong, ip, vp, vp]
    L.mpg_resize_images.argtypes = [vp, vp, ip, ip, ip, ip, ip, ip, vp, ip, ip, ip, ip, ip, vp, vp]
    L.mpg_slice_assemble.argtypes = [vp, ctypes.POINTER(AssembleDesc), vp, vp, ip, ip, vp, vp]
    L.mpg_transpose3d.argtypes = [vp, vp, vp, ip, ip, ip, ctypes.POINTER(ctypes.c_int), ctypes.c_float, vp]
    L.mpg_threshold.argtypes = [vp, vp, ctypes.c_longlong, ctypes.c_float, vp]
    L.mpg_reslab_p2p.argtypes = [vp, vp, ctypes.POINTER(vp), ip, ip, ip, ip, ctypes.POINTER(ip), ctypes.c_float, vp]
    L.mpg_reslab_p2p_part.argtypes = [vp, vp, ctypes.POINTER(vp), ip, ip, ip, ip, ip, ctypes.POINTER(ip), ctypes.c_float, vp]
    L.mpg_tiles_count.argtypes = [ip, ip, ip]
    L.mpg_tiles_cut.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip, ip, ip, ip, ip, ip, vp]
    L.mpg_tiles_stitch.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip, ip, ip, ip, vp]
    L.mpg_tiles_stitch_overlap.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip, ip, ip, ip, vp]
    ll, fl = ctypes.c_longlong, ctypes.c_float
    L.mpg_train_conv_fwd.argtypes = [vp, vp, vp, vp, vp, ip, ip, ip, ip, ip, ip, ip, ip, vp]
    L.mpg_train_conv_dgrad.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip, ip, ip, ip, ip, vp]
    L.mpg_train_conv_wgrad.argtypes = [vp, vp, vp, vp, vp, vp, ip, ip, ip, ip, ip, ip, ip, ip, vp]
    L.mpg_train_conv_wgrad_tc.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip, ip, ip, vp]
    L.mpg_train_bias_grad.argtypes = [vp, vp, vp, vp, ll, ip, vp]
    L.mpg_train_bn_fwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ll, ip, fl, fl, ip, vp]
    L.mpg_train_bn_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ll, ip, ip, vp]
    L.mpg_train_act_fwd.argtypes = [vp, vp, vp, ll, ip, vp]
    L.mpg_train_add_act_fwd.argtypes = [vp, vp, vp, vp, ll, ip, vp]
    L.mpg_train_act_bwd.argtypes = [vp, vp, vp, vp, ll, ip, vp]
    L.mpg_train_axpy.argtypes = [vp, vp, vp, fl, ll, vp]
    L.mpg_train_mul.argtypes = [vp, vp, vp, vp, ll, vp]
    L.mpg_train_bce_logits.argtypes = [vp, vp, fl, fl, vp, vp, ll, ip, vp]
    L.mpg_train_l1_mean.argtypes = [vp, vp, vp, fl, vp, vp, ll, ip, vp]
    L.mpg_train_l2_half.argtypes = [vp, vp, vp, fl, vp, vp, ll, ip, vp]
    L.mpg_train_adam.argtypes = [vp, vp, vp, vp, vp, ll, fl, fl, fl, fl, vp]
    L.mpg_train_adam_dev.argtypes = [vp, vp, vp, vp, vp, ll, vp, fl, fl, fl, vp]
    L.mpg_train_fc_fwd.argtypes = [vp, vp, vp, vp, vp, ip, ip, vp]
    L.mpg_train_fc_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, ip, ip, vp]
    L.mpg_train_take_channel.argtypes = [vp, vp, vp, ll, ip, ip, ip, vp]
    L.mpg_train_resample_fwd.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip, vp]
    L.mpg_train_resample_bwd.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip, vp]
    L.mpg_train_semilagr_pos.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip, fl, ip, vp]
    L.mpg_train_avgpool2_fwd.argtypes = [vp, vp, vp, ip, ip, ip, ip, vp]
    L.mpg_train_avgpool2_bwd.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip, vp]
    L.mpg_train_lerp.argtypes = [vp, vp, vp, vp, fl, ll, vp]
    L.mpg_train_scale.argtypes = [vp, vp, vp, fl, ll, vp]
    L.mpg_train_gp_penalty.argtypes = [vp, vp, vp, vp, vp, ip, ll, fl, fl, vp]
    L.mpg_train_mean_pow.argtypes = [vp, vp, fl, ip, vp, vp, ll, ip, vp]
    L.mpg_train_pixel_norm_fwd.argtypes = [vp, vp, vp, ll, ip, vp]
    L.mpg_train_pick.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip, ip, ip, ip, vp]
    L.mpg_train_stuff16.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip, ip, ip, vp]
    L.mpg_train_pixel_norm_bwd.argtypes = [vp, vp, vp, vp, ll, ip, vp]
    _lib = L
    return L


def check(status, what):
    if status != 0:
        msg = lib().mpg_last_error().decode("utf-8", "replace")
        raise MpgError("%s failed (status %d): %s" % (what, status, msg))


class Handle:
    """mpg_handle: one per (process, device)."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        check(lib().mpg_create(ctypes.byref(self._h), int(device)), "mpg_create")
        self.device = int(device)

    @property
    def ptr(self):
        return self._h

    @property
    def sm_count(self):
        return lib().mpg_sm_count(self._h)

    def close(self):
        if self._h:
            lib().mpg_destroy(self._h)
            self._h = ctypes.c_void_p()


_handles = {}


def default_handle(device=0):
    h = _handles.get(device)
    if h is None:
        h = _handles[device] = Handle(device)
    return h


def _fptr(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _f32c(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class ConvPlan:
    """mpg_conv_plan: fused conv (+shortcut segment) + shift + act + pixel_norm + x2 store.

    weights: list of HWIO float32 arrays (already wscale-multiplied), one per segment.
    """

    def __init__(self, handle, n, h, w, weights, cstrides, cout, out_cstride, act=None, scales=None,
                 shift=None, pixel_norm=False, upsample=1, in_upsample=1, stride=1,
                 in_dtype=BF16, out_dtype=BF16, force_kind=0):
        self.handle = handle
        d = ConvDesc()
        d.n, d.h, d.w = int(n), int(h), int(w)
        d.nseg = len(weights)
        ws = [_f32c(wt) for wt in weights]
        for s, wt in enumerate(ws):
            assert wt.ndim == 4 and wt.shape[0] == wt.shape[1] and wt.shape[3] == cout, wt.shape
            d.seg_ksize[s] = wt.shape[0]
            d.seg_cin[s] = wt.shape[2]
            d.seg_cstride[s] = int(cstrides[s])
        d.cout = int(cout)
        d.act = _ACT_BY_NAME[act] if not isinstance(act, int) else act
        d.pixel_norm = int(bool(pixel_norm))
        d.upsample = int(upsample)
        d.in_upsample = int(in_upsample)
        d.stride = int(stride)
        d.force_kind = int(force_kind)
        d.in_dtype = int(in_dtype)
        d.out_dtype = int(out_dtype)
        d.out_cstride = int(out_cstride)
        self.desc = d
        scs = [None, None]
        if scales is not None:
            for s, sc in enumerate(scales):
                scs[s] = _f32c(sc)
        sh = _f32c(shift)
        self._p = ctypes.c_void_p()
        check(lib().mpg_conv_plan_create(handle.ptr, ctypes.byref(d), _fptr(ws[0]),
                                         _fptr(ws[1]) if len(ws) > 1 else None, _fptr(scs[0]), _fptr(scs[1]),
                                         _fptr(sh), ctypes.byref(self._p)), "mpg_conv_plan_create")
        self.kind = lib().mpg_conv_plan_kind(self._p)
        self.flops = lib().mpg_conv_plan_flops(self._p)
        self.out_h = (d.h + d.stride - 1) // d.stride * d.upsample
        self.out_w = (d.w + d.stride - 1) // d.stride * d.upsample

    def run(self, x0, x1, y, stream=0):
        """x0/x1/y: torch CUDA tensors (or raw device pointers); stream: cudaStream_t as int."""
        p0 = x0.data_ptr() if hasattr(x0, "data_ptr") else int(x0)
        p1 = None if x1 is None else (x1.data_ptr() if hasattr(x1, "data_ptr") else int(x1))
        py = y.data_ptr() if hasattr(y, "data_ptr") else int(y)
        check(lib().mpg_conv_plan_run(self._p, p0, p1, py, stream), "mpg_conv_plan_run")

    def set_side(self, w_side):
        """Give the plan a side output y_side = y x w_side (w_side: float32 [cout, side_cout], scales folded)."""
        w = _f32c(w_side)
        check(lib().mpg_conv_plan_set_side(self._p, _fptr(w), int(w.shape[1])), "mpg_conv_plan_set_side")
        self.has_side = True

    def run_ex(self, x0, x1, y, y_side=None, residual=None, stream=0):
        check(lib().mpg_conv_plan_run_ex(self._p, _ptr(x0), _ptr(x1), _ptr(y), _ptr(y_side), _ptr(residual), stream),
              "mpg_conv_plan_run_ex")

    def update(self, w0, w1=None, mode0=0, mode1=0, shift=None, stream=0, src_k=0, src_cout=0, cout_off=0):
        """Refresh the packed weights from device fp32 HWIO tensors (training; tcgen05 plans only). src_k / src_cout /
        cout_off: segment 0 comes from a smaller (embedded) and / or wider source tensor (mpg_conv_plan_update_ex)."""
        check(lib().mpg_conv_plan_update_ex(self._p, _ptr(w0), _ptr(w1), int(mode0), int(mode1), _ptr(shift), int(src_k),
                                            int(src_cout), int(cout_off), stream), "mpg_conv_plan_update_ex")

    def close(self):
        if self._p:
            lib().mpg_conv_plan_destroy(self._p)
            self._p = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ResblockPlan:
    """mpg_resblock_plan: the reference's thin resBlock (GAN/multipassGAN-4x.py:505-526) as one fused launch.
    w_a [k,k,cin,cmid], w_b [k,k,cmid,cout], w_s [1,1,cin,cout]: HWIO float32, wscale folded; scales = inference-BN
    factors per conv or None; shift_a / shift_bs = folded offsets (shift_bs = conv B + shortcut)."""

    def __init__(self, handle, n, h, w, w_a, w_b, w_s, in_dtype, in_cstride, mma_dtype, out_dtype, out_cstride, act="relu",
                 scale_a=None, scale_b=None, scale_s=None, shift_a=None, shift_bs=None, in_upsample=1):
        self.handle = handle
        wa, wb, ws = _f32c(w_a), _f32c(w_b), _f32c(w_s)
        assert wa.ndim == 4 and wb.ndim == 4 and ws.ndim == 4 and ws.shape[0] == 1 and wa.shape[0] == wb.shape[0]
        assert wa.shape[3] == wb.shape[2] and ws.shape[2] == wa.shape[2] and ws.shape[3] == wb.shape[3]
        d = ResblockDesc()
        d.n, d.h, d.w = int(n), int(h), int(w)
        d.cin, d.cmid, d.cout, d.ksize = wa.shape[2], wa.shape[3], wb.shape[3], wa.shape[0]
        d.in_upsample, d.in_dtype, d.in_cstride = int(in_upsample), int(in_dtype), int(in_cstride)
        d.mma_dtype, d.out_dtype, d.out_cstride = int(mma_dtype), int(out_dtype), int(out_cstride)
        d.act = _ACT_BY_NAME[act] if not isinstance(act, int) else act
        self.desc = d
        keep = [_f32c(a) for a in (scale_a, scale_b, scale_s, shift_a, shift_bs)]
        self._p = ctypes.c_void_p()
        check(lib().mpg_resblock_plan_create(handle.ptr, ctypes.byref(d), _fptr(wa), _fptr(wb), _fptr(ws), _fptr(keep[0]),
                                             _fptr(keep[1]), _fptr(keep[2]), _fptr(keep[3]), _fptr(keep[4]),
                                             ctypes.byref(self._p)), "mpg_resblock_plan_create")
        self.flops = lib().mpg_resblock_plan_flops(self._p)

    def run(self, x, y, stream=0, sat_counter=None):
        check(lib().mpg_resblock_plan_run_checked(self._p, _ptr(x), _ptr(y), _ptr(sat_counter), stream), "mpg_resblock_plan_run")

    def close(self):
        if self._p:
            lib().mpg_resblock_plan_destroy(self._p)
            self._p = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(t):
    if t is None:
        return None
    return t.data_ptr() if hasattr(t, "data_ptr") else int(t)


def pack_channels(handle, sources, out, out_dtype, out_cstride, n, oh, ow, stream=0):
    """sources: list of (tensor_or_ptr, dtype, cstride, c0, nch, factor_h, factor_w)."""
    arr = (ChanSrc * len(sources))()
    for i, (t, dt, cs, c0, nch, fh, fw) in enumerate(sources):
        arr[i].ptr = _ptr(t)
        arr[i].dtype, arr[i].cstride, arr[i].c0, arr[i].nch = int(dt), int(cs), int(c0), int(nch)
        arr[i].factor_h, arr[i].factor_w = int(fh), int(fw)
    check(lib().mpg_pack_channels(handle.ptr, arr, len(sources), _ptr(out), int(out_dtype), int(out_cstride),
                                  int(n), int(oh), int(ow), stream), "mpg_pack_channels")


class BicubicPlan:
    def __init__(self, handle, in_h, in_w, out_h, out_w):
        self._p = ctypes.c_void_p()
        check(lib().mpg_bicubic_plan_create(handle.ptr, in_h, in_w, out_h, out_w, ctypes.byref(self._p)),
              "mpg_bicubic_plan_create")

    @property
    def ptr(self):
        return self._p

    def close(self):
        if self._p:
            lib().mpg_bicubic_plan_destroy(self._p)
            self._p = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def dens_residual(handle, dens, src, src_dtype, src_cstride, src_c, mode, bicubic_plan, n, out_h, out_w, src_h,
                  src_w, out, stream=0):
    check(lib().mpg_dens_residual(handle.ptr, _ptr(dens), _ptr(src), int(src_dtype), int(src_cstride), int(src_c),
                                  int(mode), bicubic_plan.ptr if bicubic_plan is not None else None, int(n),
                                  int(out_h), int(out_w), int(src_h), int(src_w), _ptr(out), stream),
          "mpg_dens_residual")


def dens_out(handle, x, x_dtype, x_cstride, cin, w, bias, src, src_dtype, src_cstride, src_c, mode, bicubic_plan, n, out_h,
             out_w, src_h, src_w, out, stream=0):
    """g_cdensOut (1x1 conv to one channel) + the additive density residual in one launch (mpg_dens_out)."""
    wv = np.ascontiguousarray(w, dtype=np.float32).reshape(-1)
    check(lib().mpg_dens_out(handle.ptr, _ptr(x), int(x_dtype), int(x_cstride), int(cin),
                             wv.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.c_float(float(bias)),
                             _ptr(src) if src is not None else None, int(src_dtype), int(src_cstride), int(src_c), int(mode),
                             bicubic_plan.ptr if bicubic_plan is not None else None, int(n), int(out_h), int(out_w),
                             int(src_h), int(src_w), _ptr(out), stream), "mpg_dens_out")


def resize_images(handle, src, src_dtype, src_cstride, c, n, src_h, src_w, out, out_dtype, out_cstride, out_h, out_w, mode,
                  bicubic_plan=None, stream=0):
    """tf.image.resize_images(x, [out_h, out_w], mode) with mode 0 (TF1 bilinear) or 2 (TF1 bicubic)."""
    check(lib().mpg_resize_images(handle.ptr, _ptr(src), int(src_dtype), int(src_cstride), int(c), int(n), int(src_h),
                                  int(src_w), _ptr(out), int(out_dtype), int(out_cstride), int(out_h), int(out_w),
                                  int(mode), bicubic_plan.ptr if bicubic_plan is not None else None, stream),
          "mpg_resize_images")


def count_saturated(handle, t, count, dtype, counter, stream=0):
    """counter (device int64/uint64 scalar tensor or pointer) += saturated / non-finite elements of t."""
    check(lib().mpg_count_saturated(handle.ptr, _ptr(t), int(count), int(dtype), _ptr(counter), stream), "mpg_count_saturated")


def make_assemble_desc(dims, vol_c, axis_of, zoom, chan_src, chan_scale=None, add_adj=False, out_dtype=BF16,
                       out_cstride=8, dens_slice0=0):
    d = AssembleDesc()
    for k in range(3):
        d.dims[k], d.axis_of[k], d.zoom[k] = int(dims[k]), int(axis_of[k]), int(zoom[k])
    d.vol_c = int(vol_c)
    d.nchan = len(chan_src)
    for c, src in enumerate(chan_src):
        d.chan_src[c] = int(src)
        d.chan_scale[c] = float(1.0 if chan_scale is None else chan_scale[c])
    d.add_adj = int(bool(add_adj))
    d.out_dtype = int(out_dtype)
    d.out_cstride = int(out_cstride)
    d.dens_slice0 = int(dens_slice0)
    return d


def slice_assemble(handle, desc, vol, dens, slice0, count, out, stream=0):
    check(lib().mpg_slice_assemble(handle.ptr, ctypes.byref(desc), _ptr(vol), _ptr(dens), int(slice0), int(count),
                                   _ptr(out), stream), "mpg_slice_assemble")


def transpose3d(handle, src, dst, dims, perm, threshold=0.0, stream=0):
    pa = (ctypes.c_int * 3)(*[int(x) for x in perm])
    check(lib().mpg_transpose3d(handle.ptr, _ptr(src), _ptr(dst), int(dims[0]), int(dims[1]), int(dims[2]), pa,
                                float(threshold), stream), "mpg_transpose3d")


def reslab_p2p(handle, slab, peer_ptrs, rank, S, split_axis, final_perm, threshold=0.0, stream=0):
    """Fused transpose + direct stores into every rank's output slab (peer_ptrs: device pointers, one per rank)."""
    world = len(peer_ptrs)
    pp = (ctypes.c_void_p * world)(*[int(x) for x in peer_ptrs])
    pa = (ctypes.c_int * 3)(*[int(x) for x in final_perm])
    check(lib().mpg_reslab_p2p(handle.ptr, _ptr(slab), pp, world, int(rank), int(S), int(split_axis), pa, float(threshold),
                               stream), "mpg_reslab_p2p")


def reslab_p2p_part(handle, part, peer_ptrs, S, a0, count, split_axis, final_perm, threshold=0.0, stream=0):
    """The same exchange for rows [a0, a0+count) of the old slice axis only (`part` = those rows)."""
    world = len(peer_ptrs)
    pp = (ctypes.c_void_p * world)(*[int(x) for x in peer_ptrs])
    pa = (ctypes.c_int * 3)(*[int(x) for x in final_perm])
    check(lib().mpg_reslab_p2p_part(handle.ptr, _ptr(part), pp, world, int(S), int(a0), int(count), int(split_axis), pa,
                                    float(threshold), stream), "mpg_reslab_p2p_part")


def threshold(handle, vol, count, thr, stream=0):
    check(lib().mpg_threshold(handle.ptr, _ptr(vol), int(count), float(thr), stream), "mpg_threshold")


def train_call(name, handle, *args):
    """Generic checked call of an mpg_train_* entry point; tensors are passed as torch tensors / None / ints."""
    conv = [(_ptr(a) if (a is None or hasattr(a, "data_ptr")) else a) for a in args]
    check(getattr(lib(), "mpg_train_" + name)(handle.ptr, *conv), "mpg_train_" + name)


def tiles_count(extent, tile, stride):
    return lib().mpg_tiles_count(int(extent), int(tile), int(stride))


def tiles_cut(handle, src, dst, n, h, w, c, elem_bytes, th, tw, stride_y=-1, stride_x=-1, pad=0, stream=0):
    """TileCreator.createTiles on device NHWC frames -> [n*ty*tx, th+2*pad, tw+2*pad, c]."""
    check(lib().mpg_tiles_cut(handle.ptr, _ptr(src), _ptr(dst), int(n), int(h), int(w), int(c), int(elem_bytes), int(th),
                              int(tw), int(stride_y), int(stride_x), int(pad), stream), "mpg_tiles_cut")


def tiles_stitch(handle, tiles, dst, n, ty, tx, th, tw, c, elem_bytes, border=0, stream=0):
    """TileCreator.concatTiles (tileBorder crop + concatenation) on device tiles."""
    check(lib().mpg_tiles_stitch(handle.ptr, _ptr(tiles), _ptr(dst), int(n), int(ty), int(tx), int(th), int(tw), int(c),
                                 int(elem_bytes), int(border), stream), "mpg_tiles_stitch")


def tiles_stitch_overlap(handle, tiles, dst, n, ty, tx, th, tw, c, elem_bytes, border, stream=0):
    """Inverse of tiles_cut(stride = tile - 2*border, pad=0): centres of all tiles + the outer border of frame-edge tiles."""
    check(lib().mpg_tiles_stitch_overlap(handle.ptr, _ptr(tiles), _ptr(dst), int(n), int(ty), int(tx), int(th), int(tw), int(c),
                                         int(elem_bytes), int(border), stream), "mpg_tiles_stitch_overlap")
