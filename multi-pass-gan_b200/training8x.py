"""Pieces of the 8x progressive-growing trainer on the GPU (GAN/multipassGAN-8x.py), for the configuration of
the shipped first-network training command (GAN/example_run_training.py:4: firstNNArch 1, upsamplingMode 2, use_wgan_gp 1,
no batch norm / gDrop / minibatch stddev in the discriminator):

* `GrowingDisc` -- growing_disc / growBlockDisc (:752-866): the spatial critic grown stage by stage, each stage blended in
  with lerp(old, new, percentage - (j-1)) (:596-597); forward, backward (parameter and input gradients) and the
  WGAN-GP gradient penalty (:1120-1138).  The penalty differentiates a gradient (tf.gradients inside the loss).  The critic
  is piecewise linear (convs, lrelu = 0.6x + 0.4|x|, average pooling, lerp, one dense layer), so with g = d mean(D)/d y and
  v = d penalty/d g the parameter gradient is  d/d theta [ u^T J(theta) v ]  (u = 1/B): ONE tangent pass of v through the
  linearised critic (same weights, no biases, the lrelu slopes of the primal pass) gives the tangent activations t_l, and
  the backward signals of that linear net are exactly the primal ones, so  dW_l += wgrad(t_{l-1}, delta_l)  with the
  delta_l already computed for g.  No second-order kernels are needed; biases get no gradient from the penalty.
* `critic_step` -- the WGAN-GP discriminator loss (:1111-1143) and its gradients in one call.
* `stage_mask` / `StagedAdam` -- the per-stage optimizers (:1304-1362: optimizer z only updates the variables whose name
  contains "%i" % 2**i, i <= z+1 -- a substring rule, reproduced), TF1 Adam over the flat buffer with a 0/1 mask.
* `WeightEMA` -- tf.contrib.opt.MovingAverageOptimizer(…, 0.999) (:1356-1361): shadow -= (1 - decay) (shadow - value).
Everything that computes runs in the fp32 training kernels behind the C ABI (mpg_train_*), PyTorch owns the buffers.
Checked against the fp64 autograd oracle (oracle/training8x.py), which is pinned by executing the reference's own functions.
Not built: the growing generator's training graph, the temporal discriminator, loss scaling, the training loop / CLI.
"""
import math

import numpy as np
import torch

from . import capi
from . import weights as W
from .training import ParamSet


class _Ctx:
    def __init__(self, device):
        self.h = capi.default_handle(device)
        self.device = torch.device("cuda", device)
        self.st = 0
        self.launches = 0
        self.scratch = torch.zeros(4096, dtype=torch.float64, device=self.device)

    def buf(self, shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def zeros(self, shape):
        return torch.zeros(shape, dtype=torch.float32, device=self.device)

    def call(self, name, *args):
        self.launches += 1
        capi.train_call(name, self.h, *args)


class _Conv8:
    """GAN.convolutional_layer (tools_wscale/GAN.py:80-119) without batch norm, stride 1, fp32: forward, backward, and the
    tangent pass / filter gradient of the gradient penalty."""

    def __init__(self, cx, ps, scope, k, cin, cout, act, gain=math.sqrt(2.0)):
        self.cx, self.ps, self.scope, self.k, self.cin, self.cout = cx, ps, scope, k, cin, cout
        self.act = capi._ACT_BY_NAME[act]
        self.wn = ps.add(scope + "/weight", (k, k, cin, cout), np.float32(gain / np.sqrt(k * k * cin)))
        self.bn = ps.add(scope + "/bias", (cout,))
        self.zero_bias = None

    def forward(self, x, n, h, w):
        cx, ps = self.cx, self.ps
        lin = cx.buf((n, h, w, self.cout))
        cx.call("conv_fwd", x, ps.view(ps.w, self.wn), ps.view(ps.w, self.bn), lin, n, h, w, self.cin, self.cout, self.k, 1, 1, cx.st)
        y = lin
        if self.act != capi.ACT_NONE:
            y = cx.buf(lin.shape)
            cx.call("act_fwd", lin, y, lin.numel(), self.act, cx.st)
        return y, dict(x=x, y=y, n=n, h=h, w=w)

    def backward(self, sv, dy, dx=None, accumulate=False, param_grads=True):
        cx, ps = self.cx, self.ps
        dlin = dy
        if self.act != capi.ACT_NONE:
            dlin = cx.buf(dy.shape)
            cx.call("act_bwd", sv["y"], dy, dlin, dy.numel(), self.act, cx.st)
        sv["dlin"] = dlin
        if param_grads:
            cx.call("conv_wgrad", sv["x"], dlin, ps.view(ps.gw, self.wn), ps.view(ps.gw, self.bn), cx.scratch, sv["n"], sv["h"],
                    sv["w"], self.cin, self.cout, self.k, 1, 1, cx.st)
        if dx is not None:
            cx.call("conv_dgrad", dlin, ps.view(ps.w, self.wn), dx, sv["n"], sv["h"], sv["w"], self.cin, self.cout, self.k, 1,
                    1 if accumulate else 0, cx.st)

    def tangent(self, t_in, sv):
        """Linearised layer at the primal point: act'(lin) * conv(t_in, W); also dW += wgrad(t_in, delta) (see module doc)."""
        cx, ps = self.cx, self.ps
        if self.zero_bias is None:
            self.zero_bias = cx.zeros((self.cout,))
        t_lin = cx.buf(sv["y"].shape)
        cx.call("conv_fwd", t_in, ps.view(ps.w, self.wn), self.zero_bias, t_lin, sv["n"], sv["h"], sv["w"], self.cin, self.cout,
                self.k, 1, 1, cx.st)
        cx.call("conv_wgrad", t_in, sv["dlin"], ps.view(ps.gw, self.wn), None, cx.scratch, sv["n"], sv["h"], sv["w"], self.cin,
                self.cout, self.k, 1, 1, cx.st)
        if self.act == capi.ACT_NONE:
            return t_lin
        t = cx.buf(t_lin.shape)
        cx.call("act_bwd", sv["y"], t_lin, t, t.numel(), self.act, cx.st)  # dz = dy * act'(y): the slope of the primal pass
        return t


class GrowingDisc:
    """growing_disc (GAN/multipassGAN-8x.py:782-866), upsampling_mode 2."""

    def __init__(self, tileSizeLow=16, upRes=8, n_inputChannels=6, start_fms=256, max_fms=256, filterSize=3,
                 first_nn_arch=True, batch=16, values=None, seed=1, device=0):
        self.cx = _Ctx(device)
        self.L, self.u, self.S, self.C, self.B = int(tileSizeLow), int(upRes), int(tileSizeLow) * int(upRes), int(n_inputChannels), int(batch)
        self.stages = int(round(math.log(self.u, 2)))
        self.first = bool(first_nn_arch)
        self.ps = ps = ParamSet(self.cx.device)
        cx = self.cx
        sc = "spatial-disc/"
        k = 4 if self.first else int(filterSize)
        self.c_from = {self.u: _Conv8(cx, ps, sc + "d_cfromDensity%d" % self.u, 1, 2, int(start_fms / self.u), None)}
        self.blocks = {}
        for j in range(self.stages, 0, -1):
            fms = int(min(start_fms / (2 ** j), max_fms))
            out2 = min(min(fms * 2, max_fms), start_fms // 2)
            up = 2 ** j
            c1 = (fms * 3 if up == 2 else fms * 2) if self.first else fms
            # (firstNNArch 0 declares in_channels = fms for cB although its input has c1 = fms channels: same thing)
            a = _Conv8(cx, ps, sc + "dBlock%d/d_cA%d" % (up, up), k, fms, c1, "lrelu")
            b = _Conv8(cx, ps, sc + "dBlock%d/d_cB%d" % (up, up), k, c1, out2, "lrelu")
            self.blocks[j] = (a, b, out2)
            self.c_from[2 ** (j - 1)] = _Conv8(cx, ps, sc + "d_cfromDensity%d" % (2 ** (j - 1)), 1, 2, out2, None)
        last = self.blocks[1][2]
        if not self.first:
            self.tail = (_Conv8(cx, ps, sc + "d_cA1", int(filterSize), last, 32, "lrelu"),
                         _Conv8(cx, ps, sc + "d_cB1", int(filterSize), 32, 4, None))
            last = 4
        self.fc_in = self.L * self.L * last
        self.fc_w = ps.add(sc + "d_l61/weight", (self.fc_in, 1), np.float32(1.0 / np.sqrt(self.fc_in)))  # gain 1 (:862)
        self.fc_b = ps.add(sc + "d_l61/bias", (1,))
        vals = dict(values) if values else {}
        for name, shape, _, _, _ in ps.specs:
            if name not in vals:
                vals[name] = W.init_variable(seed, name, shape, "normal" if name.endswith("/weight") else ("const", 0.1))
        ps.finalize(vals)
        self.losses = torch.zeros(4, dtype=torch.float64, device=cx.device)
        self.zero1 = cx.zeros((1,))

    # ------------------------------------------------------------------ helpers
    def refresh(self):
        ps = self.ps
        self.cx.call("mul", ps.w, ps.v, ps.scale, ps.total, self.cx.st)  # W_eff = v * wscale (tools_wscale/GAN.py:668)

    def _t(self, percentage, j):
        return float(min(max(percentage - (j - 1), 0.0), 1.0))

    def _pool(self, x, n, h, w, c):
        y = self.cx.buf((n, h // 2, w // 2, c))
        self.cx.call("avgpool2_fwd", x, y, n, h, w, c, self.cx.st)
        return y

    def _input(self, in_low, in_high):
        """concat(nearest x upRes of channel 0 of the low-res rows, in_high) (:796-813)."""
        B, S = in_high.shape[0], self.S
        xin = self.cx.buf((B, S, S, 2))
        capi.pack_channels(self.cx.h, [(in_low, capi.F32, self.C, 0, 1, self.u, self.u), (in_high, capi.F32, 1, 0, 1, 1, 1)], xin,
                           capi.F32, 2, B, S, S, self.cx.st)
        return xin

    # ------------------------------------------------------------------ forward / backward
    def forward_from_input(self, xin, percentage):
        cx, B, S = self.cx, xin.shape[0], self.S
        sv = dict(xin=xin, B=B, pct=float(percentage), lvl={})
        x_, sv["from_u"] = self.c_from[self.u].forward(xin, B, S, S)
        inH, res = xin, S
        for j in range(self.stages, 0, -1):
            a, b, out2 = self.blocks[j]
            inH = self._pool(inH, B, res, res, 2)
            x1, sa = a.forward(x_, B, res, res)
            x2, sb = b.forward(x1, B, res, res)
            pooled = self._pool(x2, B, res, res, out2)
            old, so = self.c_from[2 ** (j - 1)].forward(inH, B, res // 2, res // 2)
            t = self._t(percentage, j)
            blend = cx.buf(pooled.shape)
            cx.call("lerp", blend, old, pooled, t, blend.numel(), cx.st)
            sv["lvl"][j] = dict(sa=sa, sb=sb, so=so, res=res, t=t, inH=inH, pooled=pooled)
            x_, res = blend, res // 2
        if self.first:
            flat = sv["lvl"][1]["pooled"]  # cursor quirk: flatten() sees the pooled block output, not the last blend
        else:
            y1, sv["t1"] = self.tail[0].forward(x_, B, res, res)
            flat, sv["t2"] = self.tail[1].forward(y1, B, res, res)
        sv["flat"] = flat
        logits = cx.buf((B, 1))
        ps = self.ps
        cx.call("fc_fwd", flat, ps.view(ps.w, self.fc_w), ps.view(ps.w, self.fc_b), logits, B, self.fc_in, cx.st)
        return logits, sv

    def forward(self, in_low, in_high, percentage):
        """in_low [B, L*L*C], in_high [B, S*S] device fp32 rows -> (logits [B,1], saved state)."""
        return self.forward_from_input(self._input(in_low, in_high), percentage)

    def backward(self, sv, dlogits, need_input_grad=False, param_grads=True):
        """Returns d loss / d xin [B,S,S,2] when asked (channel 1 = the high-res sample)."""
        cx, ps, B = self.cx, self.ps, sv["B"]
        sv["dlogits"] = dlogits
        dflat = cx.buf(sv["flat"].shape)
        dw = ps.view(ps.gw, self.fc_w) if param_grads else cx.buf((self.fc_in,))
        db = ps.view(ps.gw, self.fc_b) if param_grads else cx.buf((1,))
        cx.call("fc_bwd", sv["flat"], ps.view(ps.w, self.fc_w), dlogits, dflat, dw, db, B, self.fc_in, cx.st)
        dblend = None
        if not self.first:
            d1 = cx.buf(sv["t1"]["y"].shape)
            self.tail[1].backward(sv["t2"], dflat, dx=d1, param_grads=param_grads)
            dblend = cx.buf(sv["t1"]["x"].shape)
            self.tail[0].backward(sv["t1"], d1, dx=dblend, param_grads=param_grads)
        dH = None  # gradient w.r.t. the pooled input image of the level being processed
        for j in range(1, self.stages + 1):
            a, b, out2 = self.blocks[j]
            lv = sv["lvl"][j]
            res, t = lv["res"], lv["t"]
            dpool = cx.zeros(lv["pooled"].shape)
            if dblend is not None:
                cx.call("scale", dpool, dblend, t, dpool.numel(), cx.st)
                dold = cx.buf(dblend.shape)
                cx.call("scale", dold, dblend, 1.0 - t, dold.numel(), cx.st)
                dHj = cx.buf(lv["inH"].shape) if need_input_grad else None
                self.c_from[2 ** (j - 1)].backward(lv["so"], dold, dx=dHj, param_grads=param_grads)
            else:
                dHj = cx.zeros(lv["inH"].shape) if need_input_grad else None
            if j == 1 and self.first:
                cx.call("axpy", dpool, dflat, 1.0, dpool.numel(), cx.st)
            if need_input_grad:
                if dH is not None:  # the deeper level's image gradient comes up through this level's pooling
                    cx.call("avgpool2_bwd", dH, dHj, B, res // 2, res // 2, 2, 1, cx.st)
                dH = dHj
            dx2 = cx.buf(lv["sb"]["y"].shape)
            cx.call("avgpool2_bwd", dpool, dx2, B, res, res, out2, 0, cx.st)
            dx1 = cx.buf(lv["sa"]["y"].shape)
            b.backward(lv["sb"], dx2, dx=dx1, param_grads=param_grads)
            dblend = cx.buf(lv["sa"]["x"].shape)
            a.backward(lv["sa"], dx1, dx=dblend, param_grads=param_grads)
        dxin = None
        if need_input_grad:
            dxin = cx.buf(sv["xin"].shape)
            cx.call("avgpool2_bwd", dH, dxin, B, self.S, self.S, 2, 0, cx.st)
            self.c_from[self.u].backward(sv["from_u"], dblend, dx=dxin, accumulate=True, param_grads=param_grads)
        else:
            self.c_from[self.u].backward(sv["from_u"], dblend, param_grads=param_grads)
        return dxin

    # ------------------------------------------------------------------ gradient penalty
    def gradient_penalty(self, in_low, y_gp, percentage, lam=10.0, target=1.0, loss=None):
        """WGAN-GP term of :1120-1138 for the interpolated samples y_gp [B, S*S]: adds the penalty to `loss` (device double)
        and its parameter gradient to ps.gw. Returns the per-sample gradient norms."""
        cx, ps, B, S = self.cx, self.ps, y_gp.shape[0], self.S
        loss = self.losses[1:2] if loss is None else loss
        logits, sv = self.forward(in_low, y_gp, percentage)
        dl = cx.buf((B, 1))
        dl.fill_(1.0 / B)  # d_out_loss = reduce_mean(d_out)
        dxin = self.backward(sv, dl, need_input_grad=True, param_grads=False)
        g = cx.buf((B, S * S))
        cx.call("take_channel", dxin, g, B * S * S, 2, 1, 0, cx.st)  # tf.gradients(..., [y_gp_d, x_disc])[0]
        v = cx.buf((B, S * S))
        norms = cx.buf((B,))
        cx.call("gp_penalty", g, v, loss, norms, B, S * S, float(lam), float(target), cx.st)
        # tangent pass of (0, v) through the linearised critic; every layer adds wgrad(tangent input, primal delta)
        txin = cx.zeros((B, S, S, 2))
        capi.pack_channels(cx.h, [(cx.zeros((B, S * S)), capi.F32, 1, 0, 1, 1, 1), (v, capi.F32, 1, 0, 1, 1, 1)], txin, capi.F32,
                           2, B, S, S, cx.st)
        t_x = self.c_from[self.u].tangent(txin, sv["from_u"])
        tH, res = txin, S
        for j in range(self.stages, 0, -1):
            a, b, out2 = self.blocks[j]
            lv = sv["lvl"][j]
            tH = self._pool(tH, B, res, res, 2)
            t1 = a.tangent(t_x, lv["sa"])
            t2 = b.tangent(t1, lv["sb"])
            tp = self._pool(t2, B, res, res, out2)
            if "dlin" in lv["so"]:
                told = self.c_from[2 ** (j - 1)].tangent(tH, lv["so"])
                t_x = cx.buf(tp.shape)
                cx.call("lerp", t_x, told, tp, lv["t"], t_x.numel(), cx.st)
            else:  # the blend of this level never reached the logits (firstNNArch cursor quirk, j = 1)
                t_x = tp
            if j == 1 and self.first:
                t_flat = tp
            res //= 2
        if not self.first:
            y1 = self.tail[0].tangent(t_x, sv["t1"])
            t_flat = self.tail[1].tangent(y1, sv["t2"])
        dummy_x, dummy_b = cx.buf(t_flat.shape), cx.buf((1,))
        cx.call("fc_bwd", t_flat, ps.view(ps.w, self.fc_w), sv["dlogits"], dummy_x, ps.view(ps.gw, self.fc_w), dummy_b, B,
                self.fc_in, cx.st)
        return norms

    # ------------------------------------------------------------------ the critic's optimizer step input
    def critic_step(self, in_low, y_real, y_fake, percentage, lerp_factor, weight_dld=1.0, lam=10.0, target=1.0, eps=0.001):
        """disc_loss of :1111-1143 (use_wgan_gp, not LSGAN): mean(-D(y)) * weight_dld + mean(D(G)) + eps * mean(D(y)^2) +
        gradient penalty at lerp_factor * y + (1 - lerp_factor) * G. Leaves d loss / d variables in ps.g; returns the loss
        tensor [total, penalty] (device doubles)."""
        cx, ps = self.cx, self.ps
        cx.st = torch.cuda.current_stream(cx.device).cuda_stream
        self.refresh()
        ps.gw.zero_()
        self.losses.zero_()
        disc, sv_r = self.forward(in_low, y_real, percentage)
        gen, sv_f = self.forward(in_low, y_fake, percentage)
        dl_r, dl_f = cx.buf(disc.shape), cx.buf(gen.shape)
        cx.call("mean_pow", disc, -float(weight_dld), 1, self.losses[0:1], dl_r, disc.numel(), 0, cx.st)
        cx.call("mean_pow", disc, float(eps), 2, self.losses[0:1], dl_r, disc.numel(), 1, cx.st)
        cx.call("mean_pow", gen, 1.0, 1, self.losses[0:1], dl_f, gen.numel(), 0, cx.st)
        self.backward(sv_r, dl_r)
        self.backward(sv_f, dl_f)
        lf = lerp_factor.to(device=cx.device, dtype=torch.float32).view(-1, 1)
        y_gp = (lf * y_real + (1.0 - lf) * y_fake).contiguous()
        self.gradient_penalty(in_low, y_gp, percentage, lam, target, loss=self.losses[1:2])
        cx.call("mul", ps.g, ps.gw, ps.scale, ps.total, cx.st)  # d/dv = d/dW_eff * wscale
        total = self.losses[0:1] + self.losses[1:2]
        return torch.cat([total, self.losses[1:2]])

    def grads(self):
        ps = self.ps
        host = ps.g.cpu().numpy()
        return {name: host[off:off + n].reshape(shape).copy() for name, shape, off, n, _ in ps.specs}


def stage_mask(ps, z, n_stages=3):
    """0/1 mask over the flat parameter buffer for optimizer z (:1332-1338, 1349-1355): every variable at the last stage, else
    those whose NAME contains "%i" % 2**i for some i <= z+1 (substring rule, reproduced)."""
    m = np.zeros(ps.total, np.float32)
    for name, shape, off, n, _ in ps.specs:
        if z == n_stages - 1 or any(("%i" % (2 ** i)) in name for i in range(0, z + 2)):
            m[off:off + n] = 1.0
    return torch.from_numpy(m).to(ps.device)


class StagedAdam:
    """One tf.train.AdamOptimizer per growing stage (:1304-1362), each with its own moments and step count, updating only
    its stage's variables (the others receive no gradient in `minimize(..., var_list=...)`)."""

    def __init__(self, cx, ps, learning_rates, beta1=0.0, beta2=0.99, eps=1e-8, n_stages=3):
        self.cx, self.ps = cx, ps
        self.lrs = [float(l) for l in learning_rates]
        self.b1, self.b2, self.eps = float(beta1), float(beta2), float(eps)
        f32 = dict(dtype=torch.float32, device=ps.device)
        self.state = [dict(mask=stage_mask(ps, z, n_stages), m=torch.zeros(ps.total, **f32), v=torch.zeros(ps.total, **f32), t=0)
                      for z in range(n_stages)]
        self.tmp = torch.zeros(ps.total, **f32)

    def step(self, z):
        st, ps, cx = self.state[z], self.ps, self.cx
        st["t"] += 1
        lr_t = self.lrs[z] * math.sqrt(1.0 - self.b2 ** st["t"]) / (1.0 - self.b1 ** st["t"])
        # variables outside the stage are not in var_list: zero gradient, zero moments -> the Adam update is exactly 0 for them
        cx.call("mul", self.tmp, ps.g, st["mask"], ps.total, cx.st)
        cx.call("adam", ps.v, self.tmp, st["m"], st["v"], ps.total, lr_t, self.b1, self.b2, self.eps, cx.st)


class WeightEMA:
    """Shadow copies of tf.contrib.opt.MovingAverageOptimizer(opt, 0.999) (:1356-1361), kept on the device."""

    def __init__(self, ps, decay=0.999):
        self.ps, self.decay = ps, float(decay)
        self.shadow = ps.v.clone()

    def update(self, mask=None):
        d = (1.0 - self.decay) * (self.shadow - self.ps.v)
        self.shadow -= d if mask is None else d * mask

    def export(self):
        host = self.shadow.cpu().numpy()
        return {name: host[off:off + n].reshape(shape).copy() for name, shape, off, n, _ in self.ps.specs}
