"""Lowers a recorded graph (graph.py) to fused sm_100a kernel launches through the C ABI (capi.py).

Plays the role of `sess.run(sampler, feed_dict)` (GAN/multipassGAN-out.py:446,508,570;
GAN/multipassGAN-4x.py:1128) for one fixed slice-batch size.  Fusion rules (SURVEY table 2.2):
  conv + bias + inference BN + activation                      -> one conv plan (K1-K4)
  relu(convB + conv1x1_shortcut) of a resBlock                 -> 2-segment implicit GEMM (K5)
  pixel_norm after the activation                              -> conv epilogue (K6)
  nearest x2 resize of a block output                          -> replicated store in the epilogue (K7)
  nearest resize / slice / concat / cast feeding a conv        -> one pack_channels launch (K7, K9)
  dens + bicubic(input density) | dens + input density         -> dens_residual (K8)
Inputs stay on the device; activations are fp16 (default) or bf16 with fp32 accumulation, or fp32 end to end.
"""
import os

import numpy as np
import torch

from . import capi

_BN_EPS = 1e-3  # tf.contrib.layers.batch_norm default epsilon (tools_wscale/GAN.py:110)


_TORCH_DTYPE = {capi.BF16: torch.bfloat16, capi.F16: torch.float16, capi.F32: torch.float32}


def _round_up(a, b):
    return -(-a // b) * b


class Buf:
    """A device tensor [n, h, w, cstride]; `ptr` may be rebound per run (placeholders, final output)."""

    def __init__(self, n, h, w, c, cstride, dtype, tensor=None, name=None, root=None):
        self.n, self.h, self.w, self.c, self.cstride, self.dtype = n, h, w, c, cstride, dtype
        self.tensor = tensor
        self.ptr = tensor.data_ptr() if tensor is not None else None
        self.name = name
        self.root = root  # reshape aliases resolve their pointer through the root buffer

    @property
    def nbytes(self):
        return self.n * self.h * self.w * self.cstride * (4 if self.dtype == capi.F32 else 2)


class View:
    """Lazy NHWC tensor: concatenation of channel ranges of buffers, each nearest-replicated."""

    def __init__(self, h, w, sources, bicubic_of=None, resize_mode=2):
        self.h, self.w = h, w
        self.sources = sources  # list of (Buf, c0, nch, fh, fw)
        self.c = sum(s[2] for s in sources)
        # source View of a TF1 bilinear (resize_mode 0) / bicubic (2) resize: not expressible as a lazy gather, it is
        # either fused into the density residual (bicubic) or materialised by mpg_resize_images
        self.bicubic_of = bicubic_of
        self.resize_mode = resize_mode
        self._mat = {}

    def whole_buf(self):
        """The single buffer this view aliases 1:1, or None."""
        if self.bicubic_of is None and len(self.sources) == 1:
            b, c0, nch, fh, fw = self.sources[0]
            if c0 == 0 and nch == b.c and fh == 1 and fw == 1 and b.h == self.h and b.w == self.w:
                return b
        return None


class Group:
    def __init__(self, conv):
        self.convs = [conv]
        self.act = None
        self.pn = False
        self.ups = 1
        self.out_node = conv


class _PendingDens:
    """Placeholder view of a 1x1 density-output conv whose launch is fused with the residual add that consumes it."""

    def __init__(self, grp):
        self.grp = grp


class CompiledNet:
    def __init__(self, output, weights, batch, precision="fp16", handle=None, device=0, verbose=False, dry=False,
                 range_check=False):
        """dry=True records the fused launch list without touching the GPU (host-logic tests).
        range_check=True (validation mode): every layer output is scanned for saturated 16-bit stores / non-finite values
        after it is produced (mpg_count_saturated); `run` raises capi.MpgRangeError instead of returning clipped data."""
        assert precision in ("bf16", "fp16", "fp32")
        self.dry = dry
        self.h = None if dry else (handle or capi.default_handle(device))
        self.device = None if dry else torch.device("cuda", self.h.device)
        self.batch = int(batch)
        self.precision = precision
        self.weights = weights
        self.act_dtype = {"bf16": capi.BF16, "fp16": capi.F16, "fp32": capi.F32}[precision]
        self.steps = []  # (label, callable(stream))
        self.step_flops = {}
        self.step_bufs = {}  # step index -> output Buf (debugging / per-layer parity)
        self.launches = 0
        self.flops = 0.0
        self.plans = []
        self.placeholders = {}  # name -> Buf
        self.activation_bytes = 0
        self.verbose = verbose
        self.timed_step = None  # index into self.steps bracketed by CUDA events when set (bench roofline)
        self.timed_events = []
        self.range_check = bool(range_check) and not dry
        self._lower(output)
        self.sat_counts = (torch.zeros(len(self.steps) + 1, dtype=torch.int64, device=self.device)
                           if self.range_check else None)

    # ------------------------------------------------------------------ helpers
    def _alloc(self, h, w, c, dtype, cstride=None, name=None):
        if cstride is None:
            cstride = c if dtype == capi.F32 else _round_up(c, 8)
        t = None
        if not self.dry:
            t = torch.empty((self.batch, h, w, cstride), dtype=_TORCH_DTYPE[dtype], device=self.device)
        b = Buf(self.batch, h, w, c, cstride, dtype, t, name)
        self.activation_bytes += b.nbytes
        return b

    def _materialize(self, view, dtype):
        """Return a Buf holding `view` in `dtype` with a conv-friendly channel stride."""
        if view.bicubic_of is not None:
            return self._materialize_resize(view, dtype)
        b = view.whole_buf()
        if b is not None and b.dtype == dtype and (dtype == capi.F32 or b.cstride % 8 == 0):
            return b
        if dtype in view._mat:
            return view._mat[dtype]
        out = self._alloc(view.h, view.w, view.c, dtype)
        srcs = [(sb, sb.dtype, sb.cstride, c0, nch, fh, fw) for (sb, c0, nch, fh, fw) in view.sources]
        handle = self.h

        def step(stream, srcs=srcs, out=out):
            capi.pack_channels(handle, [(self._p(s[0]),) + s[1:] for s in srcs], self._p(out), out.dtype,
                               out.cstride, out.n, out.h, out.w, stream)

        self.step_bufs[len(self.steps)] = out
        self.steps.append(("pack %dx%dx%d" % (view.h, view.w, view.c), step))
        view._mat[dtype] = out
        return out

    def _plain(self, v):
        """A lazily gatherable View of `v` (a bilinear / bicubic resize is materialised first)."""
        if v.bicubic_of is None:
            return v
        return View(v.h, v.w, [(self._materialize(v, self.act_dtype), 0, v.c, 1, 1)])

    def _materialize_resize(self, view, dtype):
        """tf.image.resize_images(.., 0 | 2) (tools_wscale/GAN.py:541) as a standalone launch."""
        if dtype in view._mat:
            return view._mat[dtype]
        sv = view.bicubic_of
        src = self._materialize(sv, capi.F32 if self.precision == "fp32" else dtype)
        out = self._alloc(view.h, view.w, view.c, dtype)
        mode = view.resize_mode
        plan = None
        if mode == 2 and not self.dry:
            plan = capi.BicubicPlan(self.h, sv.h, sv.w, view.h, view.w)
            self.plans.append(plan)
        handle = self.h

        def step(stream):
            capi.resize_images(handle, self._p(src), src.dtype, src.cstride, view.c, out.n, sv.h, sv.w, self._p(out),
                               out.dtype, out.cstride, view.h, view.w, mode, plan, stream)

        self.step_bufs[len(self.steps)] = out
        self.steps.append(("resize mode %d %dx%d->%dx%dx%d" % (mode, sv.h, sv.w, view.h, view.w, view.c), step))
        view._mat[dtype] = out
        return out

    def _w(self, var):
        try:
            return self.weights[var.name]
        except KeyError:
            raise KeyError("no value for variable '%s' in the injected weights" % var.name)

    def _conv_affine(self, node):
        """(W_eff HWIO fp32, per-cout scale or None, per-cout shift) of one conv node."""
        wv = node.attrs["weight"]
        w_eff = self._w(wv["var"]).astype(np.float32) * np.float32(wv["wscale"])  # tools_wscale/GAN.py:668
        bias = self._w(node.attrs["bias"]).astype(np.float32)
        bn = node.attrs["bn"]
        if bn is None:
            return w_eff, None, bias
        if bn["train"] not in (False, None):
            raise NotImplementedError("training-mode batch norm is handled by the training step, not CompiledNet")
        gamma, beta = self._w(bn["gamma"]).astype(np.float64), self._w(bn["beta"]).astype(np.float64)
        mean, var = self._w(bn["moving_mean"]).astype(np.float64), self._w(bn["moving_variance"]).astype(np.float64)
        scale = gamma / np.sqrt(var + _BN_EPS)
        shift = (bias.astype(np.float64) - mean) * scale + beta
        return w_eff, scale.astype(np.float32), shift.astype(np.float32)

    # ------------------------------------------------------------------ lowering
    def _lower(self, output):
        g = output.graph
        # reachable sub-DAG
        need = set()
        stack = [output.node]
        while stack:
            n = stack.pop()
            if n.id in need:
                continue
            need.add(n.id)
            stack.extend(t.node for t in n.inputs)
        nodes = [n for n in g.nodes if n.id in need]
        uses = {n.id: [] for n in nodes}
        for n in nodes:
            for t in n.inputs:
                uses[t.node.id].append(n)

        # ---- group formation
        absorbed = {}
        groups_by_out = {}
        for n in nodes:
            if n.op != "conv" or n.id in absorbed:
                continue
            grp = Group(n)
            t = n
            while True:
                us = uses[t.id]
                if len(us) != 1 or t is output.node:
                    break
                u = us[0]
                if u.op == "act" and grp.act is None and not grp.pn:
                    grp.act = u.attrs["kind"]
                elif u.op == "add" and grp.act is None and not grp.pn and len(grp.convs) == 1:
                    other = u.inputs[1].node if u.inputs[0].node is t else u.inputs[0].node
                    ok = (other.op == "conv" and other is not t and other.id > n.id and other.id not in absorbed
                          and len(uses[other.id]) == 1 and other.attrs["stride"] == 1 and n.attrs["stride"] == 1
                          and other.out.shape == n.out.shape)
                    if not ok:
                        break
                    grp.convs.append(other)
                    absorbed[other.id] = grp
                elif u.op == "pixel_norm" and not grp.pn:
                    grp.pn = True
                elif (u.op == "resize" and u.attrs["method"] == 1 and n.attrs["stride"] == 1
                      and u.out.shape[1] == 2 * t.out.shape[1] and u.out.shape[2] == 2 * t.out.shape[2]):
                    grp.ups = 2
                    absorbed[u.id] = grp
                    t = u
                    break
                else:
                    break
                absorbed[u.id] = grp
                t = u
            grp.out_node = t
            absorbed[n.id] = grp
            groups_by_out[t.id] = grp

        # ---- thin residual blocks (resBlock of GAN/multipassGAN-4x.py:505-526 with <= 8 input / middle channels):
        #      conv A + conv B + 1x1 shortcut become ONE launch (csrc/resblock_thin.cu)
        rb_of_a, rb_of_b = {}, {}
        fuse = int(os.environ.get("MPG_FUSE_RESBLOCK", "1"))
        if self.precision != "fp32" and fuse:
            for gb in groups_by_out.values():
                ga = self._match_thin_resblock(gb, groups_by_out, uses, output)
                # <= 8 output channels (ru4: 8->2->1): with one 8-column MMA tile per A fragment the fused kernel is bound
                # by its ldmatrix traffic and measured slower (0.137 ms) than the two CUDA-core conv_tiny launches
                # (0.096 ms for 8 x 512^2): fused only on request (MPG_FUSE_RESBLOCK=2)
                if ga is not None and (fuse >= 2 or gb.convs[0].out.shape[3] == 32):
                    rb_of_a[id(ga)] = gb
                    rb_of_b[id(gb)] = ga

        # ---- the 1x1 shortcut of a block whose input is a 128-channel tcgen05 conv output is computed in THAT conv's epilogue
        #      (mpg_conv_plan_set_side) and added by the consumer as an fp32 residual: removes the re-read of the 128-channel
        #      tensor (537 MB per slice batch for ru3 of gen_resnet, GAN/multipassGAN-4x.py:521,563)
        self._side_of = {}   # id(producer group) -> (consumer group, shortcut conv)
        self._side_buf = {}  # id(consumer group) -> fp32 [n,h,w,8] side tensor, once the producer was emitted with it
        if self.precision != "fp32" and int(os.environ.get("MPG_FUSE_SHORTCUT", "1")):
            for gx in groups_by_out.values():
                if len(gx.convs) != 2 or gx.pn or gx.ups != 1 or id(gx) in rb_of_b:
                    continue
                c5, cs = sorted(gx.convs, key=lambda c: -c.attrs["ksize"])
                if c5.attrs["ksize"] not in (3, 5) or cs.attrs["ksize"] != 1 or cs.out.shape[3] > 8:
                    continue
                gp = groups_by_out.get(cs.inputs[0].node.id)
                if (gp is None or gp.pn or gp.ups != 1 or gp.convs[0].out.shape[3] != 128 or id(gp) in self._side_of
                        or id(gp) in rb_of_b or id(gp) in rb_of_a):
                    continue
                self._side_of[id(gp)] = (gx, cs)

        # ---- density output (g_cdensOut: 1x1 conv to ONE channel, no activation, GAN/multipassGAN-out.py:282) whose only
        #      consumer is the additive residual (:327-332): one bandwidth-bound launch (mpg_dens_out) instead of a padded
        #      tensor-core conv plus a second pass over the fp32 image
        dens_fused = set()
        if int(os.environ.get("MPG_FUSE_DENSOUT", "1")):
            for gd in groups_by_out.values():
                c = gd.convs[0]
                if (len(gd.convs) == 1 and c.attrs["ksize"] == 1 and c.attrs["stride"] == 1 and c.out.shape[3] == 1
                        and gd.act is None and not gd.pn and gd.ups == 1 and gd.out_node is c and c is not output.node
                        and c.attrs["bn"] is None and c.inputs[0].shape[3] <= 64
                        and len(uses[c.id]) == 1 and uses[c.id][0].op == "add" and uses[c.id][0].out.shape[3] == 1):
                    dens_fused.add(id(gd))

        views = {}
        for n in nodes:
            if n.id in groups_by_out:
                grp = groups_by_out[n.id]
                if id(grp) in dens_fused:
                    views[n.id] = _PendingDens(grp)
                    continue
                if id(grp) in rb_of_a:
                    xin = views[grp.convs[0].inputs[0].node.id]
                    if self._thin_resblock_input(xin, grp.convs[0]) is not None:
                        views[n.id] = None  # produced inside the fused launch of the block, never materialised
                        continue
                    del rb_of_b[id(rb_of_a.pop(id(grp)))]  # input layout not supported: separate launches
                if id(grp) in rb_of_b:
                    views[n.id] = self._emit_thin_resblock(rb_of_b[id(grp)], grp, views)
                    continue
                views[n.id] = self._emit_group(grp, views)
                continue
            if n.id in absorbed:
                continue
            op = n.op
            if op == "placeholder":
                per = int(np.prod(n.out.shape[1:]))
                b = Buf(self.batch, 1, 1, per, per, capi.F32, None, n.attrs["name"])
                self.placeholders[n.attrs["name"]] = b
                views[n.id] = View(1, 1, [(b, 0, per, 1, 1)])
            elif op == "reshape":
                views[n.id] = self._lower_reshape(n, views[n.inputs[0].node.id])
            elif op == "slice":
                views[n.id] = self._lower_slice(n, views[n.inputs[0].node.id])
            elif op == "concat":
                vs = [views[t.node.id] for t in n.inputs]
                srcs = []
                for v in vs:
                    srcs.extend(self._plain(v).sources)
                views[n.id] = View(n.out.shape[1], n.out.shape[2], srcs)
            elif op == "resize":
                views[n.id] = self._lower_resize(n, views[n.inputs[0].node.id])
            elif op == "add":
                views[n.id] = self._lower_add(n, views)
            else:
                raise NotImplementedError("op '%s' is not fusable into a conv on this path and has no standalone "
                                          "kernel (node %d)" % (op, n.id))
        outv = views[output.node.id]
        ob = outv.whole_buf()
        if ob is None or ob.dtype != capi.F32:
            ob = self._materialize(outv, capi.F32)
        self.out_buf = ob
        self.out_shape = tuple(output.shape[1:])
        self.launches = len(self.steps)

    def _lower_reshape(self, n, v):
        shp = n.out.shape
        if len(shp) == 4 and (shp[1], shp[2], shp[3]) == (v.h, v.w, v.c):
            return v  # NHWC -> same NHWC: no-op (GAN/multipassGAN-out.py:297 on the concatenated input)
        b = v.whole_buf()
        if b is None or b.cstride != b.c:
            b = self._materialize(v, capi.F32 if (b is None or b.dtype == capi.F32) else b.dtype)
            if b.cstride != b.c:
                raise NotImplementedError("reshape of a channel-padded tensor")
        if len(shp) == 2:
            h, w, c = 1, 1, shp[1]
        else:
            h, w, c = shp[1], shp[2], shp[3]
        assert h * w * c == b.h * b.w * b.c
        alias = Buf(b.n, h, w, c, c, b.dtype, None, b.name, root=b.root or b)
        return View(h, w, [(alias, 0, c, 1, 1)])

    def _lower_slice(self, n, v):
        v = self._plain(v)
        c0, c1 = n.attrs["c0"], n.attrs["c1"]
        srcs, pos = [], 0
        for (b, s0, nch, fh, fw) in v.sources:
            lo, hi = max(c0, pos), min(c1, pos + nch)
            if lo < hi:
                srcs.append((b, s0 + lo - pos, hi - lo, fh, fw))
            pos += nch
        return View(v.h, v.w, srcs)

    def _lower_resize(self, n, v):
        oh, ow = n.out.shape[1], n.out.shape[2]
        m = n.attrs["method"]
        if m == 1:
            if oh % v.h or ow % v.w:
                raise NotImplementedError("nearest resize with a non-integer factor")
            fh, fw = oh // v.h, ow // v.w
            v = self._plain(v)
            return View(oh, ow, [(b, c0, nch, f0 * fh, f1 * fw) for (b, c0, nch, f0, f1) in v.sources])
        if m in (0, 2):  # 0: TF1 bilinear (GAN.avg_depool default mode), 2: TF1 bicubic
            return View(oh, ow, list(v.sources), bicubic_of=self._plain(v), resize_mode=m)
        raise NotImplementedError("resize method %d is not a tf.image.ResizeMethod used by tools_wscale/GAN.py" % m)

    def _lower_add(self, n, views):
        va, vb = views[n.inputs[0].node.id], views[n.inputs[1].node.id]
        if n.out.shape[3] != 1:
            raise NotImplementedError("standalone add is only implemented for the 1-channel density residual")
        pend = va if isinstance(va, _PendingDens) else (vb if isinstance(vb, _PendingDens) else None)
        feat = None
        if pend is not None:
            conv = pend.grp.convs[0]
            other = vb if pend is va else va
            feat = self._materialize(views[conv.inputs[0].node.id], self.act_dtype)
            oh, ow = feat.h, feat.w
        else:
            dens = va.whole_buf()
            other = vb
            if dens is None or dens.dtype != capi.F32:
                dens, other = vb.whole_buf(), va
            if dens is None or dens.dtype != capi.F32:
                raise NotImplementedError("density residual: neither operand is a dense fp32 tensor")
            oh, ow = dens.h, dens.w
        out = self._alloc(oh, ow, 1, capi.F32)
        handle = self.h
        if other.bicubic_of is not None and other.resize_mode != 2:
            other = View(other.h, other.w, [(self._materialize(other, capi.F32), 0, other.c, 1, 1)])
        if other.bicubic_of is not None:
            sv = other.bicubic_of
            (sb, c0, nch, fh, fw), = sv.sources
            assert nch == 1 and fh == 1 and fw == 1
            plan = None
            if not self.dry:
                plan = capi.BicubicPlan(handle, sv.h, sv.w, oh, ow)
                self.plans.append(plan)
            mode, sh, sw = 2, sv.h, sv.w
        else:
            (sb, c0, nch, fh, fw), = other.sources
            if fh != 1 or fw != 1:
                sb = self._materialize(other, capi.F32)
                c0 = 0
            plan, mode, sh, sw = None, 0, oh, ow

        if feat is not None:
            w_eff, sc, shift = self._conv_affine(conv)
            cin = w_eff.shape[2]
            wvec = (w_eff[0, 0, :, 0] * (sc[0] if sc is not None else 1.0)).astype(np.float32)
            bias = float(shift[0])
            flops = 2.0 * self.batch * oh * ow * cin
            self.flops += flops

            def step(stream):
                capi.dens_out(handle, self._p(feat), feat.dtype, feat.cstride, cin, wvec, bias, self._p(sb), sb.dtype, sb.cstride,
                              c0, mode, plan, out.n, oh, ow, sh, sw, self._p(out), stream)

            label = "dens_out %s k1 %d->1 + residual mode %d %dx%d" % (
                conv.attrs["weight"]["var"].name.rsplit("/", 2)[-2], cin, mode, oh, ow)
            if self.verbose:
                print(label)
            self.step_flops[len(self.steps)] = flops
            self.steps.append((label, step))
            return View(out.h, out.w, [(out, 0, 1, 1, 1)])

        def step(stream):
            capi.dens_residual(handle, self._p(dens), self._p(sb), sb.dtype, sb.cstride, c0, mode, plan, out.n, dens.h,
                               dens.w, sh, sw, self._p(out), stream)

        self.steps.append(("dens_residual mode %d" % mode, step))
        return View(out.h, out.w, [(out, 0, 1, 1, 1)])

    @staticmethod
    def _p(buf):
        return buf.root.ptr if buf.root is not None else buf.ptr

    def _emit_group(self, grp, views):
        convs = sorted(grp.convs, key=lambda c: -c.attrs["ksize"])
        first = convs[0]
        resid = self._side_buf.get(id(grp))  # the shortcut of this block was computed by the producer of its input
        ws, scs, shift_total, ins = [], [], None, []
        any_scale = False
        for c in convs:
            w_eff, sc, sh = self._conv_affine(c)
            shift_total = sh if shift_total is None else shift_total + sh
            if resid is not None and c is not first:
                continue  # its weights live in the producer's epilogue; only the folded offset stays here
            ws.append(w_eff)
            scs.append(sc)
            any_scale = any_scale or sc is not None
            ins.append(self._materialize(views[c.inputs[0].node.id], self.act_dtype))
        cout = first.out.shape[3]
        ih, iw = first.inputs[0].shape[1], first.inputs[0].shape[2]
        stride = first.attrs["stride"]
        out_dtype = capi.F32 if (self.precision == "fp32" or cout == 1) else self.act_dtype
        oh, ow = -(-ih // stride) * grp.ups, -(-iw // stride) * grp.ups
        out = self._alloc(oh, ow, cout, out_dtype)
        flops = sum(2.0 * self.batch * (oh // grp.ups) * (ow // grp.ups) * c.attrs["ksize"] ** 2
                    * c.inputs[0].shape[3] * cout for c in convs)
        self.flops += flops
        x0 = ins[0]
        x1 = ins[1] if len(ins) > 1 else None
        side = None
        if self.dry:
            plan, kind = None, self._predict_kind(convs if resid is None else convs[:1], ins, cout, out_dtype, stride, grp.ups,
                                                  out.cstride, grp.act)
        else:
            def make(ws_, scs_, ins_):
                return capi.ConvPlan(self.h, self.batch, ih, iw, ws_, [b.cstride for b in ins_], cout, out.cstride,
                                     act=grp.act, scales=scs_ if any(sc_ is not None for sc_ in scs_) else None,
                                     shift=shift_total, pixel_norm=grp.pn, upsample=grp.ups, stride=stride,
                                     in_dtype=self.act_dtype, out_dtype=out_dtype)

            plan = make(ws, scs, ins)
            if resid is not None and not (plan.kind in (capi.KIND_NFOLD, capi.KIND_VFOLD, capi.KIND_VRING) and cout <= 8):
                # the consumer kernel cannot add a residual: back to the two-segment form (the side tensor stays unused)
                plan.close()
                resid = None
                ws, scs, ins = [], [], []
                for c in convs:
                    w_eff, sc, _ = self._conv_affine(c)
                    ws.append(w_eff)
                    scs.append(sc)
                    ins.append(self._materialize(views[c.inputs[0].node.id], self.act_dtype))
                x0, x1 = ins[0], (ins[1] if len(ins) > 1 else None)
                plan = make(ws, scs, ins)
            self.plans.append(plan)
            kind = plan.kind
            if id(grp) in self._side_of:
                gx, cs = self._side_of[id(grp)]
                w_s, sc_s, _ = self._conv_affine(cs)  # [1,1,128,k]; its offset is added by the consumer
                w_side = w_s[0, 0] * (sc_s[None, :] if sc_s is not None else 1.0)
                try:
                    plan.set_side(w_side.astype(np.float32))
                    side = self._alloc(oh, ow, 8, capi.F32, cstride=8)
                    self._side_buf[id(gx)] = side
                except capi.MpgError:
                    side = None  # this plan cannot carry a side output: the consumer keeps its shortcut segment

        def step(stream):
            if side is None and resid is None:
                plan.run(self._p(x0), self._p(x1) if x1 is not None else None, self._p(out), stream)
            else:
                plan.run_ex(self._p(x0), self._p(x1) if x1 is not None else None, self._p(out),
                            y_side=self._p(side) if side is not None else None,
                            residual=self._p(resid) if resid is not None else None, stream=stream)

        if x1 is not None and resid is not None:
            x1 = None
        label = "conv[%s] %s k%s %s->%d %dx%d%s%s%s%s%s" % (
            {capi.KIND_TCGEN05: "tc", capi.KIND_NFOLD: "nf", capi.KIND_TINY: "ct", capi.KIND_VFOLD: "vf", capi.KIND_VRING: "vr"}.get(kind, "cc"),
            "+".join(c.attrs["weight"]["var"].name.rsplit("/", 2)[-2] for c in convs),
            "/".join(str(c.attrs["ksize"]) for c in convs),
            "/".join(str(c.inputs[0].shape[3]) for c in convs), cout, ih, iw,
            " " + grp.act if grp.act else "", " pn" if grp.pn else "", " up2" if grp.ups == 2 else "",
            " +side8" if side is not None else "", " (shortcut via residual)" if resid is not None else "")
        if self.verbose:
            print(label)
        self.step_flops[len(self.steps)] = flops
        self.step_bufs[len(self.steps)] = out
        self.steps.append((label, step))
        return View(oh, ow, [(out, 0, cout, 1, 1)])

    # ------------------------------------------------------------------ fused thin residual block
    @staticmethod
    def _match_thin_resblock(gb, groups_by_out, uses, output):
        """Group `gb` = act(conv5x5(A_out) + conv1x1(X)); returns the group of A = act(conv5x5(X)) when the three convs
        form a thin resBlock the fused kernel covers, else None."""
        if len(gb.convs) != 2 or gb.pn or gb.ups != 1 or gb.act not in (None, "relu", "lrelu"):
            return None
        cb, cs = sorted(gb.convs, key=lambda c: -c.attrs["ksize"])
        if cb.attrs["ksize"] != 5 or cs.attrs["ksize"] != 1 or cb.attrs["stride"] != 1 or cs.attrs["stride"] != 1:
            return None
        ta = cb.inputs[0].node
        ga = groups_by_out.get(ta.id)
        if ga is None or len(ga.convs) != 1 or ga.pn or ga.ups != 1 or ga.act != gb.act or ta is output.node:
            return None
        ca = ga.convs[0]
        if ca.attrs["ksize"] != 5 or ca.attrs["stride"] != 1 or len(uses[ta.id]) != 1:
            return None
        if cs.inputs[0].node is not ca.inputs[0].node:
            return None
        cin, cmid, cout = ca.inputs[0].shape[3], ca.out.shape[3], cb.out.shape[3]
        if cin > 8 or cmid > 8 or not (cout == 32 or cout <= 8):
            return None
        return ga

    def _thin_resblock_input(self, v, ca):
        """(buffer, nearest factor) when the block input view can be read directly by the fused kernel."""
        if v is None or v.bicubic_of is not None or len(v.sources) != 1:
            return None
        b, c0, nch, fh, fw = v.sources[0]
        cin = ca.inputs[0].shape[3]
        if c0 != 0 or nch != cin or fh != fw:
            return None
        if b.dtype == capi.F32 and cin <= 4 and b.cstride % 4 == 0:
            return b, fh
        if b.dtype == self.act_dtype and b.dtype != capi.F32 and b.cstride % 8 == 0:
            return b, fh
        return None

    def _emit_thin_resblock(self, ga, gb, views):
        ca = ga.convs[0]
        cb, cs = sorted(gb.convs, key=lambda c: -c.attrs["ksize"])
        xin = views[ca.inputs[0].node.id]
        src, up = self._thin_resblock_input(xin, ca)
        (wa, sca, sha), (wb, scb, shb), (ws, scs, shs) = (self._conv_affine(c) for c in (ca, cb, cs))
        cin, cmid, cout = wa.shape[2], wa.shape[3], wb.shape[3]
        out_dtype = capi.F32 if cout == 1 else self.act_dtype
        out = self._alloc(xin.h, xin.w, cout, out_dtype)
        flops = 2.0 * self.batch * xin.h * xin.w * (25.0 * cin * cmid + 25.0 * cmid * cout + cin * cout)
        self.flops += flops
        plan = None
        if not self.dry:
            plan = capi.ResblockPlan(self.h, self.batch, xin.h, xin.w, wa, wb, ws, src.dtype, src.cstride, self.act_dtype,
                                     out_dtype, out.cstride, act=gb.act, scale_a=sca, scale_b=scb, scale_s=scs,
                                     shift_a=sha, shift_bs=shb + shs, in_upsample=up)
            self.plans.append(plan)
            assert abs(plan.flops - flops) < 1e-6 * flops

        idx = len(self.steps)

        def step(stream):
            plan.run(self._p(src), self._p(out), stream,
                     sat_counter=self.sat_counts[idx:idx + 1] if self.sat_counts is not None else None)

        nm = lambda c: c.attrs["weight"]["var"].name.rsplit("/", 2)[-2]
        label = "resblock[hm] %s>%s+%s k5/5/1 %d->%d->%d %dx%d%s%s" % (
            nm(ca), nm(cb), nm(cs), cin, cmid, cout, xin.h, xin.w, " " + gb.act if gb.act else "",
            " in_up%d" % up if up > 1 else "")
        if self.verbose:
            print(label)
        self.step_flops[len(self.steps)] = flops
        self.step_bufs[len(self.steps)] = out
        self.steps.append((label, step))
        return View(xin.h, xin.w, [(out, 0, cout, 1, 1)])

    def _predict_kind(self, convs, ins, cout, out_dtype, stride, ups=1, out_cstride=None, act=None):
        """Mirror of the auto rule in csrc/conv_plan.cu (dry runs only)."""
        if self.act_dtype == capi.F32 or stride != 1 or cout > 128:
            return capi.KIND_DIRECT
        if any(c.attrs["ksize"] not in (1, 3, 5) for c in convs) or any(b.cstride % 8 for b in ins):
            return capi.KIND_DIRECT
        cp = _round_up(cout, 8)
        if (ups == 1 and cout <= 2 and convs[0].attrs["ksize"] in (3, 5)
                and (len(convs) == 1 or convs[1].attrs["ksize"] == 1)
                and all(c.inputs[0].shape[3] <= 8 for c in convs) and all(b.cstride == 8 for b in ins)
                and (out_dtype == capi.F32 or out_cstride % 8 == 0)):
            return capi.KIND_TINY
        if (ups == 1 and cout <= 32 and act != "tanh" and convs[0].attrs["ksize"] in (3, 5)
                and (len(convs) == 1 or convs[1].attrs["ksize"] == 1)
                and (out_cstride <= 32 if out_dtype == capi.F32 else out_cstride == cp)
                and (cp * convs[0].attrs["ksize"] <= 64 or sum(c.inputs[0].shape[3] for c in convs) >= 64)):
            return capi.KIND_NFOLD
        return capi.KIND_TCGEN05

    # ------------------------------------------------------------------ execution
    def run(self, feeds, out=None, stream=None):
        """feeds: placeholder name -> device tensor / pointer holding [batch, n_flat] fp32 rows.
        out: optional device tensor / pointer receiving the [batch, n_output] fp32 rows; returns the
        output tensor when `out` is None."""
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        if self.dry:
            raise RuntimeError("dry CompiledNet cannot run")
        for name, b in self.placeholders.items():
            if name not in feeds:
                raise KeyError("placeholder '%s' was not fed" % name)
            f = feeds[name]
            b.ptr = f.data_ptr() if hasattr(f, "data_ptr") else int(f)
        ob = self.out_buf
        root = ob.root or ob
        saved = root.ptr
        if out is not None:
            root.ptr = out.data_ptr() if hasattr(out, "data_ptr") else int(out)
        try:
            if self.range_check:
                for i, (_, step) in enumerate(self.steps):
                    step(st)
                    ob_i = self.step_bufs.get(i)
                    if ob_i is not None:
                        capi.count_saturated(self.h, self._p(ob_i), ob_i.n * ob_i.h * ob_i.w * ob_i.cstride, ob_i.dtype,
                                             self.sat_counts[i:i + 1], st)
            elif self.timed_step is None:
                for _, step in self.steps:
                    step(st)
            else:
                cur = torch.cuda.current_stream(self.device)
                ts = cur if cur.cuda_stream == st else torch.cuda.ExternalStream(st, device=self.device)
                for i, (_, step) in enumerate(self.steps):
                    if i == self.timed_step:
                        e0 = torch.cuda.Event(enable_timing=True)
                        e1 = torch.cuda.Event(enable_timing=True)
                        e0.record(ts)
                        step(st)
                        e1.record(ts)
                        self.timed_events.append((e0, e1))
                    else:
                        step(st)
        finally:
            root.ptr = saved
        if self.range_check:
            self.check_range()
        if out is None:
            return root.tensor.view(self.batch, -1)
        return out

    def saturated(self, reset=True):
        """{layer label: saturated / non-finite elements} accumulated since the last reset (synchronises)."""
        if self.sat_counts is None:
            return {}
        host = self.sat_counts.cpu().tolist()
        if reset:
            self.sat_counts.zero_()
        return {self.steps[i][0]: int(c) for i, c in enumerate(host[:len(self.steps)]) if c}

    def check_range(self):
        bad = self.saturated()
        if bad:
            raise capi.MpgRangeError(bad)

    def dominant_step(self):
        """Index, label and FLOPs of the conv step with the most algorithmic FLOPs."""
        best = max(range(len(self.steps)), key=lambda i: self.step_flops.get(i, 0.0))
        return best, self.steps[best][0], self.step_flops[best]

    def close(self):
        for p in self.plans:
            if p is not None:
                p.close()
        self.plans = []
