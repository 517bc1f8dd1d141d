"""Training step of the 4x model on the GPU: generator `gen_resnet` + spatial discriminator
`disc_binclass`, the GAN losses, TF1 Adam and the BN moving-average updates of GAN/multipassGAN-4x.py
(:528-620 networks, :726-730 wiring, :744-768 losses, :783-787 variable split, :889-898 optimizers,
:1316-1397 loop body), data parallel over `torch.distributed` (one NCCL all-reduce of the flat gradient per
optimizer step).  Everything that computes runs in hand-written kernels behind the C ABI (mpg_train_*);
PyTorch only owns the buffers.  fp32 throughout (the reference trains in fp32).

Variables live in two flat fp32 buffers (generator / discriminator, the reference's `"g_" in name` /
`"d_" in name` split) so that the weight scaling, the gradient all-reduce and Adam are one launch each.
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import capi
from . import parallel as par
from . import weights as W

BN_EPS = 1e-3  # tf.contrib.layers.batch_norm default


class ParamSet:
    """Flat parameter storage of one optimizer's variables."""

    def __init__(self, device):
        self.device = device
        self.specs = []  # (name, shape, offset, numel, scale)
        self.total = 0
        self.index = {}

    def add(self, name, shape, scale=1.0):
        n = int(np.prod(shape))
        self.index[name] = len(self.specs)
        self.specs.append((name, tuple(int(s) for s in shape), self.total, n, float(scale)))
        self.total += (n + 3) // 4 * 4  # keep every variable 16-byte aligned
        return name

    def finalize(self, values):
        f32 = dict(dtype=torch.float32, device=self.device)
        host_v = np.zeros(self.total, np.float32)
        host_s = np.ones(self.total, np.float32)
        for name, shape, off, n, sc in self.specs:
            host_v[off:off + n] = np.asarray(values[name], np.float32).reshape(-1)
            host_s[off:off + n] = np.float32(sc)
        self.v = torch.from_numpy(host_v).to(self.device)        # the TF variables (unscaled N(0,1) weights)
        self.scale = torch.from_numpy(host_s).to(self.device)    # wscale constant per element (1 for bias / BN)
        self.w = torch.empty(self.total, **f32)                  # effective values v * scale
        self.gw = torch.zeros(self.total, **f32)                 # gradient w.r.t. the effective values
        self.g = torch.zeros(self.total, **f32)                  # gradient w.r.t. the variables
        self.m = torch.zeros(self.total, **f32)
        self.vv = torch.zeros(self.total, **f32)
        self.t = 0

    def view(self, buf, name):
        _, shape, off, n, _ = self.specs[self.index[name]]
        return buf[off:off + n]

    def export(self):
        host = self.v.cpu().numpy()
        return {name: host[off:off + n].reshape(shape).copy() for name, shape, off, n, _ in self.specs}


class _Conv:
    """One GAN.convolutional_layer (tools_wscale/GAN.py:80-119) in training mode."""

    def __init__(self, tr, ps, scope, k, stride, cin, cout, bn, act, in_up=1, gain=math.sqrt(2.0)):
        self.tr, self.ps, self.scope = tr, ps, scope
        self.k, self.stride, self.cin, self.cout, self.bn, self.in_up = k, stride, cin, cout, bn, in_up
        self.act = capi._ACT_BY_NAME[act]
        std = np.float32(gain / np.sqrt(k * k * cin))  # tools_wscale/GAN.py:664-667
        self.wn = ps.add(scope + "/weight", (k, k, cin, cout), std)
        self.bn_ = ps.add(scope + "/bias", (cout,))
        if bn:
            self.beta = ps.add(scope + "/beta", (cout,))
            self.gamma = ps.add(scope + "/gamma", (cout,))
            tr.moving[scope + "/moving_mean"] = None
            tr.moving[scope + "/moving_variance"] = None
        # tensor-core path (precision "fp16"): every stride-1 conv runs forward on the tcgen05 implicit-GEMM kernel with fp16
        # operands and dgrad on the same kernel with bf16 gradients (range) and flipped/transposed weights; channel counts
        # are padded to 8 in the 16-bit copies (zero channels), nearest-upsampled inputs are materialised by the pack
        self.fast = tr.precision == "fp16" and k in (1, 3, 5) and stride == 1
        self.cpi, self.cpo = -(-cin // 8) * 8, -(-cout // 8) * 8
        # filter gradient on the tensor cores (csrc/conv_wgrad_tc.cu): bf16 copies of the input and of the output gradient
        self.fast_w = (tr.precision == "fp16" and k in (1, 3, 5) and stride == 1 and in_up == 1
                       and ((cin == 128 and cout in (32, 64, 128)) or (cout == 128 and cin in (32, 64))))
        self.plan_f = self.plan_d = None
        # discriminator convs (k = 4, stride 2 / 1; GAN/multipassGAN-4x.py:593-614) on the tensor cores too: TF's SAME window
        # of a 4x4 kernel is taps -1..+2 of a 5x5 one, so the conv runs as a stride-1 5x5 plan at the INPUT resolution with the
        # weights embedded (mpg_conv_plan_update_ex) and a stride-2 layer keeps every other output pixel (4x the MACs, on a
        # pipe that is ~50x faster than the fp32 CUDA-core kernel these layers used: 42 % of the serialized loop body);
        # the input gradient is the same plan's dgrad of dy scattered onto the sampled positions. cout > 128 = several plans.
        self.emb = tr.precision == "fp16" and k == 4 and in_up == 1 and stride in (1, 2)
        self.plans_f = []

    def build_plans(self, n, h, w):
        if self.emb and not self.plans_f:
            tr = self.tr
            for c0 in range(0, self.cout, 128):
                cc = min(128, self.cout - c0)
                self.plans_f.append((c0, cc, capi.ConvPlan(tr.h, n, h, w, [np.zeros((5, 5, self.cin, cc), np.float32)], [self.cpi], cc, cc,
                                                           act=None, shift=np.zeros(cc, np.float32), in_dtype=capi.F16,
                                                           out_dtype=capi.F32, force_kind=1)))
            self.plan_d = capi.ConvPlan(tr.h, n, h, w, [np.zeros((5, 5, self.cout, self.cin), np.float32)], [self.cpo], self.cin,
                                        self.cin, act=None, shift=np.zeros(self.cin, np.float32), in_dtype=capi.BF16,
                                        out_dtype=capi.F32, force_kind=1)
            return
        if not self.fast or self.plan_f is not None:
            return
        tr = self.tr
        zf = np.zeros((self.k, self.k, self.cin, self.cout), np.float32)
        zd = np.zeros((self.k, self.k, self.cout, self.cin), np.float32)
        self.plan_f = capi.ConvPlan(tr.h, n, h, w, [zf], [self.cpi], self.cout, self.cout, act=None,
                                    shift=np.zeros(self.cout, np.float32), in_dtype=capi.F16, out_dtype=capi.F32, force_kind=1)
        if self.in_up == 1:  # layers fed by the (upsampled) data never need an input gradient
            self.plan_d = capi.ConvPlan(tr.h, n, h, w, [zd], [self.cpo], self.cin, self.cin, act=None,
                                        shift=np.zeros(self.cin, np.float32), in_dtype=capi.BF16, out_dtype=capi.F32, force_kind=1)

    def refresh(self):
        if self.emb and self.plans_f:
            ps, tr = self.ps, self.tr
            bias = ps.view(ps.w, self.bn_)
            for c0, cc, pl in self.plans_f:
                pl.update(ps.view(ps.w, self.wn), mode0=0, shift=bias[c0:c0 + cc], stream=tr.st, src_k=4, src_cout=self.cout,
                          cout_off=c0)
            self.plan_d.update(ps.view(ps.w, self.wn), mode0=1, stream=tr.st, src_k=4)
            tr.launches += len(self.plans_f) + 1
            return
        if self.plan_f is None:
            return
        ps, tr = self.ps, self.tr
        self.plan_f.update(ps.view(ps.w, self.wn), mode0=0, shift=ps.view(ps.w, self.bn_), stream=tr.st)
        tr.launches += 1
        if self.plan_d is not None:
            self.plan_d.update(ps.view(ps.w, self.wn), mode0=1, stream=tr.st)
            tr.launches += 1

    def forward(self, x, n, h, w):
        """x: [n, h/in_up, w/in_up, cin]; h, w: conv input size. Returns (y, saved)."""
        tr, ps = self.tr, self.ps
        oh, ow = -(-h // self.stride), -(-w // self.stride)
        lin = tr.buf((n, oh, ow, self.cout))
        if self.emb and self.plans_f:
            x16 = tr.buf16((n, h, w, self.cpi), torch.float16)
            capi.pack_channels(tr.h, [(x, capi.F32, self.cin, 0, self.cin, 1, 1)], x16, capi.F16, self.cpi, n, h, w, tr.st)
            tr.launches += 1
            for c0, cc, pl in self.plans_f:
                full = tr.buf((n, h, w, cc))
                pl.run(x16, None, full, tr.st)
                tr.call("pick", full, lin, n, oh, ow, cc, self.stride, cc, self.cout, c0, tr.st)
                tr.launches += 1
        elif self.fast:
            x16 = tr.buf16((n, h, w, self.cpi), torch.float16)
            capi.pack_channels(tr.h, [(x, capi.F32, self.cin, 0, self.cin, self.in_up, self.in_up)], x16, capi.F16, self.cpi, n, h,
                               w, tr.st)
            self.plan_f.run(x16, None, lin, tr.st)
            tr.launches += 2
        else:
            tr.call("conv_fwd", x, ps.view(ps.w, self.wn), ps.view(ps.w, self.bn_), lin, n, h, w, self.cin, self.cout, self.k,
                    self.stride, self.in_up, tr.st)
        rows = n * oh * ow
        sv = dict(x=x, lin=lin, n=n, h=h, w=w, rows=rows)
        if self.bn:
            y = tr.buf((n, oh, ow, self.cout))
            sv["mean"], sv["var"], sv["invstd"] = tr.buf((self.cout,)), tr.buf((self.cout,)), tr.buf((self.cout,))
            tr.call("bn_fwd", lin, ps.view(ps.w, self.gamma), ps.view(ps.w, self.beta), y, sv["mean"], sv["var"],
                    sv["invstd"], tr.moving[self.scope + "/moving_mean"], tr.moving[self.scope + "/moving_variance"],
                    tr.scratch, rows, self.cout, BN_EPS, tr.bn_decay_for(self), self.act, tr.st)
        elif self.act != capi.ACT_NONE:
            y = tr.buf((n, oh, ow, self.cout))
            tr.call("act_fwd", lin, y, rows * self.cout, self.act, tr.st)
        else:
            y = lin
        sv["y"] = y
        return y, sv

    def backward(self, sv, dy, dx=None, accumulate=False, param_grads=True):
        """dy: gradient w.r.t. the activated output. Writes / accumulates dx when given."""
        tr, ps = self.tr, self.ps
        rows = sv["rows"]
        if self.bn:
            dlin = tr.buf(sv["lin"].shape)
            tr.call("bn_bwd", sv["lin"], sv["y"], dy, ps.view(ps.w, self.gamma), sv["mean"], sv["invstd"], None, dlin,
                    ps.view(ps.gw, self.gamma), ps.view(ps.gw, self.beta), tr.scratch, rows, self.cout, self.act, tr.st)
        elif self.act != capi.ACT_NONE:
            dlin = tr.buf(sv["lin"].shape)
            tr.call("act_bwd", sv["y"], dy, dlin, rows * self.cout, self.act, tr.st)
        else:
            dlin = dy
        if self.emb and self.plans_f:
            if param_grads:
                tr.call("conv_wgrad", sv["x"], dlin, ps.view(ps.gw, self.wn), ps.view(ps.gw, self.bn_), tr.scratch, sv["n"],
                        sv["h"], sv["w"], self.cin, self.cout, self.k, self.stride, self.in_up, tr.st)
            if dx is not None:
                oh, ow = dlin.shape[1], dlin.shape[2]
                g16 = tr.buf16((sv["n"], sv["h"], sv["w"], self.cpo), torch.bfloat16)
                tr.call("stuff16", dlin, g16, capi.BF16, sv["n"], oh, ow, self.cout, self.stride, self.cpo, tr.st)
                tgt = tr.buf(tuple(dx.shape)) if accumulate else dx
                self.plan_d.run(g16, None, tgt, tr.st)
                tr.launches += 1
                if accumulate:
                    tr.call("axpy", dx, tgt, 1.0, dx.numel(), tr.st)
            return
        g16 = None
        use_wtc = param_grads and self.fast_w and sv["w"] % 16 == 0
        if use_wtc or (dx is not None and self.fast):
            g16 = tr.buf16(tuple(dlin.shape[:3]) + (self.cpo,), torch.bfloat16)
            capi.pack_channels(tr.h, [(dlin, capi.F32, self.cout, 0, self.cout, 1, 1)], g16, capi.BF16, self.cpo, sv["n"],
                               sv["h"], sv["w"], tr.st)
            tr.launches += 1
        if use_wtc:
            xb = tr.buf16(tuple(sv["x"].shape), torch.bfloat16)
            capi.pack_channels(tr.h, [(sv["x"], capi.F32, self.cin, 0, self.cin, 1, 1)], xb, capi.BF16, self.cin, sv["n"],
                               sv["h"], sv["w"], tr.st)
            tr.launches += 1
            tr.call("conv_wgrad_tc", xb, g16, ps.view(ps.gw, self.wn), sv["n"], sv["h"], sv["w"], self.cin, self.cout, self.k,
                    tr.st)
            tr.call("bias_grad", dlin, ps.view(ps.gw, self.bn_), tr.scratch, rows, self.cout, tr.st)
        elif param_grads:
            tr.call("conv_wgrad", sv["x"], dlin, ps.view(ps.gw, self.wn), ps.view(ps.gw, self.bn_), tr.scratch, sv["n"],
                    sv["h"], sv["w"], self.cin, self.cout, self.k, self.stride, self.in_up, tr.st)
        if dx is not None:
            assert self.in_up == 1
            if self.fast:
                tgt = tr.buf(tuple(dx.shape)) if accumulate else dx
                self.plan_d.run(g16, None, tgt, tr.st)
                tr.launches += 1
                if accumulate:
                    tr.call("axpy", dx, tgt, 1.0, dx.numel(), tr.st)
            else:
                tr.call("conv_dgrad", dlin, ps.view(ps.w, self.wn), dx, sv["n"], sv["h"], sv["w"], self.cin, self.cout, self.k,
                        self.stride, 1 if accumulate else 0, tr.st)


class Trainer4x:
    """multipassGAN-4x.py training loop body for `upsampling_mode 2` tiles (tileSizeLow^2 x 4 -> (4 tileSizeLow)^2)."""

    def __init__(self, tileSizeLow=16, upRes=4, batch=16, values=None, seed=1, batch_norm=True, bn_decay=0.999,
                 learning_rate=2e-4, adam_beta1=0.5, weight_dld=1.0, k2_l=(1.0, 1.0, 1.0, 1.0), device=0, group=None,
                 precision="fp32", graphs=False):
        """graphs: replay each optimizer step as ONE captured CUDA graph (the first call of a step runs eagerly, the
        second is captured, later ones are replays; with world > 1 the graph ends before the gradient all-reduce). The step is ~2300 launches of a few microseconds each,
        so without the graph the host launch path, not the GPU, bounds it.
        precision "fp32": every kernel fp32 (the parity mode, what the reference computes); "fp16": the wide stride-1
        convolutions run forward / dgrad on the tcgen05 kernel (fp16 activations, bf16 gradients, fp32 accumulation and
        fp32 master weights / optimizer), everything else stays fp32."""
        assert precision in ("fp32", "fp16")
        self.precision = precision
        self._graphs = {}
        self.h = capi.default_handle(device)
        self.device = torch.device("cuda", device)
        self.L, self.u, self.S, self.B, self.C = int(tileSizeLow), int(upRes), int(tileSizeLow) * int(upRes), int(batch), 4
        self.bn, self.bn_decay = bool(batch_norm), float(bn_decay)
        self.lr, self.beta1, self.beta2, self.eps = float(learning_rate), float(adam_beta1), 0.999, 1e-8
        self.weight_dld, self.k2_l = float(weight_dld), tuple(float(k) for k in k2_l)
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.use_graphs = bool(graphs)  # world > 1: the gradient all-reduce and the Adam launch stay outside the graph
        self.moving = {}
        self.pg, self.pd = ParamSet(self.device), ParamSet(self.device)
        C = self.C
        # ---- generator (GAN/multipassGAN-4x.py:528-569): 4 resBlocks (A: k5 relu, B: k5, s: 1x1), BN on ru1-ru3
        self.rbs = []
        chans = [(C, 2 * C, 8 * C, self.bn), (8 * C, 128, 128, self.bn), (128, 32, 8, self.bn), (8, 2, 1, False)]
        for i, (cin, s1, s2, bn) in enumerate(chans):
            up = self.u if i == 0 else 1
            a = _Conv(self, self.pg, "generator/g_cA%d" % i, 5, 1, cin, s1, bn, "relu", in_up=up)
            b = _Conv(self, self.pg, "generator/g_cB%d" % i, 5, 1, s1, s2, bn, None)
            s = _Conv(self, self.pg, "generator/g_s%d" % i, 1, 1, cin, s2, bn, None, in_up=up)
            self.rbs.append((a, b, s))
        # ---- discriminator (:572-620): k4 convs s2,s2,s2,s1 with lrelu, BN on d_c2..d_c4, FC head
        self.dcs = [_Conv(self, self.pd, "discriminator/d_c1", 4, 2, 2, 32, False, "lrelu"),
                    _Conv(self, self.pd, "discriminator/d_c2", 4, 2, 32, 64, self.bn, "lrelu"),
                    _Conv(self, self.pd, "discriminator/d_c3", 4, 2, 64, 128, self.bn, "lrelu"),
                    _Conv(self, self.pd, "discriminator/d_c4", 4, 1, 128, 256, self.bn, "lrelu")]
        self.fc_in = (self.S // 8) ** 2 * 256
        self.fc_w = self.pd.add("discriminator/d_l5/weight", (self.fc_in, 1), np.float32(math.sqrt(2.0) / np.sqrt(self.fc_in)))
        self.fc_b = self.pd.add("discriminator/d_l5/bias", (1,))
        # ---- values
        vals = dict(values) if values else {}
        for ps in (self.pg, self.pd):
            for name, shape, _, _, _ in ps.specs:
                if name not in vals:
                    leaf = name.rsplit("/", 1)[-1]
                    kind = "normal" if leaf == "weight" else ("const", {"bias": 0.1, "gamma": 1.0, "beta": 0.0}[leaf])
                    vals[name] = W.init_variable(seed, name, shape, kind)
            ps.finalize(vals)
        for name in list(self.moving):
            c = None
            for ps in (self.pg, self.pd):
                key = name.rsplit("/", 1)[0] + "/bias"
                if key in ps.index:
                    c = ps.specs[ps.index[key]][1][0]
            init = vals.get(name)
            if init is None:
                init = np.full((c,), 0.0 if name.endswith("moving_mean") else 1.0, np.float32)
            self.moving[name] = torch.from_numpy(np.ascontiguousarray(init, np.float32)).to(self.device)
        for a, b, sc in self.rbs:
            for c in (a, b, sc):
                c.build_plans(self.B, self.S, self.S)
        hh = self.S
        for c in self.dcs:
            c.build_plans(self.B, hh, hh)
            hh = -(-hh // c.stride)
        self.dirty = {"g": True, "d": True}  # variable sets whose effective / packed weights are stale
        self.scratch = torch.zeros(4096, dtype=torch.float64, device=self.device)
        self.losses = torch.zeros(8, dtype=torch.float64, device=self.device)
        self.st = 0
        self.launches = 0
        self._bufs = []

    # ------------------------------------------------------------------ plumbing
    def bn_decay_for(self, conv):
        """gen_resnet builds GAN(_in) with the class default 0.999 (GAN/multipassGAN-4x.py:546, tools_wscale/GAN.py:19);
        only disc_binclass passes the `bnDecay` flag (:606)."""
        return self.bn_decay if conv.scope.rsplit("/", 1)[-1].startswith("d_") else 0.999

    def buf(self, shape):
        t = torch.empty(shape, dtype=torch.float32, device=self.device)
        self._bufs.append(t)
        return t

    def buf16(self, shape, dtype):
        t = torch.empty(shape, dtype=dtype, device=self.device)
        self._bufs.append(t)
        return t

    def call(self, name, *args):
        self.launches += 1
        capi.train_call(name, self.h, *args)

    def _dev(self, a):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=torch.float32).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)

    def _refresh_weights(self):
        """Effective weights and packed tensor-core copies -- only of the variable set an optimizer has changed since the last
        refresh (a generator step leaves the discriminator's copies valid and vice versa: half of the re-packing launches)."""
        if self.dirty["g"]:
            self.call("mul", self.pg.w, self.pg.v, self.pg.scale, self.pg.total, self.st)  # W_eff = v * wscale (GAN.py:668)
            for a, b, sc in self.rbs:
                for c in (a, b, sc):
                    c.refresh()
        if self.dirty["d"]:
            self.call("mul", self.pd.w, self.pd.v, self.pd.scale, self.pd.total, self.st)
            for c in self.dcs:
                c.refresh()
        self.dirty["g"] = self.dirty["d"] = False

    # ------------------------------------------------------------------ networks
    def gen_forward(self, x):
        """x: [B, L*L*4] rows -> gen_part [B, S, S, 1]."""
        B, L, S = self.B, self.L, self.S
        inp = x.view(B, L, L, self.C)
        saved = []
        for i, (a, b, s) in enumerate(self.rbs):
            ya, sa = a.forward(inp, B, S, S)
            yb, sb = b.forward(ya, B, S, S)
            ys, ss = s.forward(inp, B, S, S)
            out = self.buf(yb.shape)
            self.call("add_act_fwd", yb, ys, out, out.numel(), capi.ACT_RELU, self.st)  # relu(tf.add(gc2, gs1)) :523
            saved.append((sa, sb, ss, out))
            inp = out
        return inp, saved

    def gen_backward(self, saved, dout):
        for i in range(len(self.rbs) - 1, -1, -1):
            a, b, s = self.rbs[i]
            sa, sb, ss, out = saved[i]
            dsum = self.buf(out.shape)
            self.call("act_bwd", out, dout, dsum, out.numel(), capi.ACT_RELU, self.st)
            dya = self.buf(sa["y"].shape)
            b.backward(sb, dsum, dx=dya)
            if i > 0:
                dinp = self.buf(sa["x"].shape)
                a.backward(sa, dya, dx=dinp)
                s.backward(ss, dsum, dx=dinp, accumulate=True)
                dout = dinp
            else:  # the block input is the (upsampled) data: no gradient needed
                a.backward(sa, dya)
                s.backward(ss, dsum)

    def disc_forward(self, in_low, in_high):
        """in_low: [B, L*L] (the first n_input/C floats of the flat x rows, App. D.5), in_high: [B, S*S]."""
        B, L, S = self.B, self.L, self.S
        xin = self.buf((B, S, S, 2))
        capi.pack_channels(self.h, [(in_low, capi.F32, 1, 0, 1, self.u, self.u), (in_high, capi.F32, 1, 0, 1, 1, 1)], xin,
                           capi.F32, 2, B, S, S, self.st)
        self.launches += 1
        feats, saved = [], []
        cur, hh = xin, S
        for c in self.dcs:
            cur, sv = c.forward(cur, B, hh, hh)
            hh = -(-hh // c.stride)
            feats.append(cur)
            saved.append(sv)
        logits = self.buf((B, 1))
        self.call("fc_fwd", cur, self.pd.view(self.pd.w, self.fc_w), self.pd.view(self.pd.w, self.fc_b), logits, B,
                  self.fc_in, self.st)
        return logits, feats, saved

    def disc_backward(self, saved, feats, dlogits, dfeats=None, need_input_grad=False, param_grads=True):
        pd = self.pd
        d4 = self.buf(feats[3].shape)
        if param_grads:
            self.call("fc_bwd", feats[3], pd.view(pd.w, self.fc_w), dlogits, d4, pd.view(pd.gw, self.fc_w),
                      pd.view(pd.gw, self.fc_b), self.B, self.fc_in, self.st)
        else:
            dummy_w, dummy_b = self.buf((self.fc_in,)), self.buf((1,))
            self.call("fc_bwd", feats[3], pd.view(pd.w, self.fc_w), dlogits, d4, dummy_w, dummy_b, self.B, self.fc_in, self.st)
        dcur = d4
        dxin = None
        for i in range(3, -1, -1):
            if dfeats is not None and dfeats[i] is not None:
                self.call("axpy", dcur, dfeats[i], 1.0, dcur.numel(), self.st)
            c = self.dcs[i]
            if i > 0:
                dprev = self.buf(saved[i]["x"].shape)
                c.backward(saved[i], dcur, dx=dprev, param_grads=param_grads)
                dcur = dprev
            elif need_input_grad:
                dxin = self.buf(saved[0]["x"].shape)
                c.backward(saved[0], dcur, dx=dxin, param_grads=param_grads)
            else:
                c.backward(saved[0], dcur, param_grads=param_grads)
        return dxin

    # ------------------------------------------------------------------ optimizer
    def _prep_adam(self, ps):
        """Host half of tf.train.AdamOptimizer: step count and bias-corrected step size (into a device scalar)."""
        ps.t += 1
        ps.lr_t = self.lr * math.sqrt(1.0 - self.beta2 ** ps.t) / (1.0 - self.beta1 ** ps.t)
        if self.use_graphs:
            if not hasattr(ps, "lr_dev"):
                ps.lr_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
            ps.lr_dev.fill_(ps.lr_t)

    def _adam(self, ps):
        self.call("mul", ps.g, ps.gw, ps.scale, ps.total, self.st)  # d/dv = d/dW_eff * wscale
        if self.use_graphs and self.world > 1:
            return  # captured part ends here; _run_step all-reduces and applies Adam eagerly
        self._adam_apply(ps)

    def _adam_apply(self, ps):
        par.allreduce_mean(ps.g, self.group)
        if self.use_graphs:
            self.call("adam_dev", ps.v, ps.g, ps.m, ps.vv, ps.total, ps.lr_dev, self.beta1, self.beta2, self.eps, self.st)
        else:
            self.call("adam", ps.v, ps.g, ps.m, ps.vv, ps.total, ps.lr_t, self.beta1, self.beta2, self.eps, self.st)

    def _forward_all(self, x, y):
        self._refresh_weights()
        gen_part, gsaved = self.gen_forward(x)
        in_low = x[:, : self.L * self.L].contiguous()  # App. D.5: first n_input/C floats of the interleaved row
        disc, dfeat, dsv = self.disc_forward(in_low, y)
        gen, gfeat, gsv = self.disc_forward(in_low, gen_part.view(self.B, -1))
        return x, y, gen_part, gsaved, (disc, dfeat, dsv), (gen, gfeat, gsv)

    def _run_step(self, key, ps, body, x_rows, y_rows):
        """Run one optimizer step eagerly, or (graphs=True) capture it on its second call and replay it afterwards."""
        x, y = self._dev(x_rows), self._dev(y_rows)
        self._prep_adam(ps)
        which = "g" if ps is self.pg else "d"
        if not self.use_graphs:
            self._bufs = []
            self.st = torch.cuda.current_stream(self.device).cuda_stream
            body(x, y)
            self.dirty[which] = True
            return self.losses
        key = key + (self.dirty["g"], self.dirty["d"])  # the captured body refreshes exactly the stale sets
        ent = self._graphs.get(key)
        if ent is None:  # warm-up: plans, shared-memory attributes and allocator pools settle outside any capture
            self._bufs = []
            self.st = torch.cuda.current_stream(self.device).cuda_stream
            body(x, y)
            self._after_graph(ps)
            self._graphs[key] = "warm"
            self.dirty[which] = True
            return self.losses
        if ent == "warm":
            sx, sy = x.clone(), y.clone()
            launches = self.launches
            torch.cuda.synchronize(self.device)
            stale = dict(self.dirty)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._bufs = []
                self.st = torch.cuda.current_stream(self.device).cuda_stream
                self.dirty.update(stale)
                body(sx, sy)
            ent = dict(graph=g, bufs=self._bufs, x=sx, y=sy, launches=self.launches - launches)
            self._graphs[key] = ent
            self._bufs = []
        else:
            ent["x"].copy_(x)
            ent["y"].copy_(y)
            self.launches += ent["launches"]
        ent["graph"].replay()
        self._after_graph(ps)
        self.dirty["g"] = self.dirty["d"] = False  # the replayed body refreshed what was stale ...
        self.dirty[which] = True                   # ... and its optimizer changed this set
        return self.losses

    def _after_graph(self, ps):
        if self.use_graphs and self.world > 1:
            self.st = torch.cuda.current_stream(self.device).cuda_stream
            self._adam_apply(ps)

    def disc_step(self, x_rows, y_rows):
        """sess.run(disc_optimizer, ...) (:1321): Adam on the d_ variables with disc_loss (:751-755)."""
        return self._run_step(("d",), self.pd, self._disc_body, x_rows, y_rows)

    def _disc_body(self, x, y):
        x, y, gen_part, gsaved, (disc, dfeat, dsv), (gen, gfeat, gsv) = self._forward_all(x, y)
        self.pd.gw.zero_()
        self.losses.zero_()
        dl_r, dl_f = self.buf(disc.shape), self.buf(gen.shape)
        self.call("bce_logits", disc, 1.0, self.weight_dld, self.losses[0:1], dl_r, self.B, 0, self.st)
        self.call("bce_logits", gen, 0.0, 1.0, self.losses[0:1], dl_f, self.B, 0, self.st)
        self.disc_backward(dsv, dfeat, dl_r)
        self.disc_backward(gsv, gfeat, dl_f)
        self._adam(self.pd)

    def gen_step(self, x_rows, y_rows, kk, kk2):
        """sess.run(gen_optimizer, ...) (:1352): Adam on the g_ variables with gen_loss_complete (:757-768)."""
        return self._run_step(("g", float(kk), float(kk2)), self.pg, lambda x, y: self._gen_body(x, y, kk, kk2), x_rows, y_rows)

    def _gen_body(self, x, y, kk, kk2):
        x, y, gen_part, gsaved, (disc, dfeat, dsv), (gen, gfeat, gsv) = self._forward_all(x, y)
        self.pg.gw.zero_()
        self.losses.zero_()
        dl_f = self.buf(gen.shape)
        self.call("bce_logits", gen, 1.0, 1.0, self.losses[1:2], dl_f, self.B, 0, self.st)          # gen_loss
        dgen = self.buf(gen_part.shape)
        self.call("l1_mean", y, gen_part, float(kk), self.losses[2:3], dgen, gen_part.numel(), 0, self.st)  # kk * gen_l1_loss
        dfe = []
        for i in range(4):  # kk2 * k2_li * tf.nn.l2_loss(dy_i - gy_i)
            d = self.buf(gfeat[i].shape)
            self.call("l2_half", dfeat[i], gfeat[i], float(kk2) * self.k2_l[i], self.losses[3:4], d, d.numel(), 0, self.st)
            dfe.append(d)
        dxin = self.disc_backward(gsv, gfeat, dl_f, dfeats=dfe, need_input_grad=True, param_grads=False)
        self.call("take_channel", dxin, dgen, self.B * self.S * self.S, 2, 1, 1, self.st)
        self.gen_backward(gsaved, dgen)
        self._adam(self.pg)

    def iteration(self, batches_d, batches_g, kk=5.0, kk2=1e-5):
        """One loop body (:1316-1397). Returns python floats (one D2H read of the loss vector per step)."""
        out = {}
        for xb, yb in batches_d:
            out["disc_loss"] = float(self.disc_step(xb, yb)[0].item())
        for xb, yb in batches_g:
            l = self.gen_step(xb, yb, kk, kk2).cpu().numpy()
            out.update(gen_loss=float(l[1]), gen_l1_loss_scaled=float(l[2]), disc_loss_layer_scaled=float(l[3]),
                       gen_loss_complete=float(l[1] + l[2] + l[3]))
        return out

    def values(self):
        out = {}
        out.update(self.pg.export())
        out.update(self.pd.export())
        for n, t in self.moving.items():
            out[n] = t.cpu().numpy().copy()
        return out

    def save_checkpoint(self, prefix):
        """saver.save(sess, test_path + 'model_%04d.ckpt') (GAN/multipassGAN-4x.py:1173): all generator /
        discriminator variables and BN moving statistics as a TF checkpoint-V2 bundle (tfckpt.py)."""
        from . import tfckpt
        tfckpt.write_checkpoint(prefix, self.values())

    def grads(self, which):
        ps = self.pg if which == "g" else self.pd
        host = ps.g.cpu().numpy()
        return {name: host[off:off + n].reshape(shape).copy() for name, shape, off, n, _ in ps.specs}
