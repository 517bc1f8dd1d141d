"""The 8x progressive-growing trainer on the GPU (GAN/multipassGAN-8x.py), for the configurations of the two
shipped training commands (GAN/example_run_training.py:4,7: use_wgan_gp 1, lambda_t 1.0, no batch norm / gDrop / minibatch
stddev; first network = firstNNArch 1 + upsamplingMode 2, refinement network = upsamplingMode 1):

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
* `GrowingGen` -- growing_gen in TRAINING mode (:626-750, output=False): per-stage density outputs blended with lerp, nearest
  x2 between stages, resBlocks with pixel_norm (forward + backward kernels), the TF1-bicubic residual of the input density.
* `Trainer8x` -- the loop body of :1898-2075 without the temporal terms: critic step, generator step (g_loss_d + lambda l1)
  through the critic's input gradient, the optimizers of growing stage z, the generator EMA.
Everything that computes runs in the fp32 training kernels behind the C ABI (mpg_train_*), PyTorch owns the buffers.
Checked against the fp64 autograd oracle (oracle/training8x.py), which is pinned by executing the reference's own functions.
* `Trainer8x.train` -- the training loop around it (:1898-2089): the growing / blending / learning-rate schedules
  (schedule8x.GrowthSchedule, pinned by traces of the reference's own loop statements), discRuns / genRuns, the nearest resize
  of the stage's target tiles to the full tile size (:1057-1058), the 1-in-20 empty-density batches of getinput (:1527-1533),
  growing events, the save rule (:2076-2084) and `save` / `load` of `model_%04d.ckpt` + `model_ema_%04d.ckpt` (:1804-1807)
  as TF checkpoint-V2 bundles that multipassGAN-out.py restores.
* the temporal critic (lambda_t, both shipped commands): `GrowingDisc(kind="tempo")` = growing_disc_tempo (:868-923) on three
  aligned frames per pixel, its WGAN-GP loss with one gradient norm per (sample, frame) (:1262-1289), the t_adam_* staged
  optimizers and the generator term kkt * mean(-T(G frames)) (`Trainer8x(lambda_t=...)`, `t_disc_step`, `gen_step(..., x_t, y_t)`).
  Frame alignment of the shipped commands (adv_flag 1, adv_mode 0): `tensorResample` (:545-594) of the generated and the target
  frames at given positions (`y_pos`; kernels mpg_train_resample_fwd/_bwd), at the full tile size.
  `TempoBatches` = getTempoinput / TileCreator.selectRandomTempoTiles: three-frame tiles from device-resident sequences, rows
  (sample, frame), and the semi-Lagrangian positions of every frame (getSemiLagrPosBatch, kernel mpg_train_semilagr_pos).
Not built: the in-graph advection of adv_mode 1 / 2, loss scaling (numerically the identity), the feature-layer loss (lambda2,
0 in the shipped commands), the .uni loading of the sequences and the command line.
"""
import math

import numpy as np
import torch

from . import capi
from . import parallel as par
from . import schedule8x
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
    """growing_disc (GAN/multipassGAN-8x.py:782-866). upsampling_mode 2 (first network): every growBlockDisc ends in a 2x2
    average pooling and the image is pooled along (:771-772, 828-829); upsampling_mode 1 / 3 (refinement networks): no
    pooling anywhere, every stage works at the full tile size (:773-774)."""

    def __init__(self, tileSizeLow=16, upRes=8, n_inputChannels=6, start_fms=256, max_fms=256, filterSize=3,
                 first_nn_arch=True, batch=16, values=None, seed=1, device=0, cx=None, upsampling_mode=2, kind="spatial"):
        """kind "spatial": growing_disc, the conditional critic of (nearest-upsampled low-res density, high-res sample);
        kind "tempo": growing_disc_tempo (:868-923), the unconditional critic of three aligned frames [B, S*S, 3] -- the same
        growing structure with the name prefix "t" in scope "tempo-disc" and a 3-channel image."""
        self.cx = cx if cx is not None else _Ctx(device)
        if kind not in ("spatial", "tempo"):
            raise ValueError("kind must be 'spatial' or 'tempo'")
        self.kind = kind
        self.cimg = 2 if kind == "spatial" else 3   # channels of the image the critic looks at
        if int(upsampling_mode) not in (1, 2, 3):
            raise ValueError("upsampling_mode %r: built are 2 (first network) and 1 / 3 (refinement networks)" % (upsampling_mode,))
        self.pool = int(upsampling_mode) == 2
        if not self.pool and first_nn_arch:
            raise ValueError("firstNNArch is the first network's architecture (upsampling_mode 2)")
        self.L, self.u, self.S, self.C, self.B = int(tileSizeLow), int(upRes), int(tileSizeLow) * int(upRes), int(n_inputChannels), int(batch)
        self.stages = int(round(math.log(self.u, 2)))
        self.first = bool(first_nn_arch)
        self.ps = ps = ParamSet(self.cx.device)
        cx = self.cx
        sc, p = ("spatial-disc/", "d") if kind == "spatial" else ("tempo-disc/", "t")
        k = 4 if self.first else int(filterSize)
        self.c_from = {self.u: _Conv8(cx, ps, sc + "%s_cfromDensity%d" % (p, self.u), 1, self.cimg, int(start_fms / self.u), None)}
        self.blocks = {}
        for j in range(self.stages, 0, -1):
            fms = int(min(start_fms / (2 ** j), max_fms))
            out2 = min(min(fms * 2, max_fms), start_fms // 2)
            up = 2 ** j
            c1 = (fms * 3 if up == 2 else fms * 2) if self.first else fms
            # (firstNNArch 0 declares in_channels = fms for cB although its input has c1 = fms channels: same thing)
            a = _Conv8(cx, ps, sc + "%sBlock%d/%s_cA%d" % (p, up, p, up), k, fms, c1, "lrelu")
            b = _Conv8(cx, ps, sc + "%sBlock%d/%s_cB%d" % (p, up, p, up), k, c1, out2, "lrelu")
            self.blocks[j] = (a, b, out2)
            self.c_from[2 ** (j - 1)] = _Conv8(cx, ps, sc + "%s_cfromDensity%d" % (p, 2 ** (j - 1)), 1, self.cimg, out2, None)
        last = self.blocks[1][2]
        if not self.first:
            self.tail = (_Conv8(cx, ps, sc + "%s_cA1" % p, int(filterSize), last, 32, "lrelu"),
                         _Conv8(cx, ps, sc + "%s_cB1" % p, int(filterSize), 32, 4, None))
            last = 4
        self.fc_in = (self.L * self.L if self.pool else self.S * self.S) * last
        self.fc_w = ps.add(sc + "%s_l61/weight" % p, (self.fc_in, 1), np.float32(1.0 / np.sqrt(self.fc_in)))  # gain 1 (:862)
        self.fc_b = ps.add(sc + "%s_l61/bias" % p, (1,))
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
            ro = res // 2 if self.pool else res
            if self.pool:
                inH = self._pool(inH, B, res, res, self.cimg)
            x1, sa = a.forward(x_, B, res, res)
            x2, sb = b.forward(x1, B, res, res)
            pooled = self._pool(x2, B, res, res, out2) if self.pool else x2
            old, so = self.c_from[2 ** (j - 1)].forward(inH, B, ro, ro)
            t = self._t(percentage, j)
            blend = cx.buf(pooled.shape)
            cx.call("lerp", blend, old, pooled, t, blend.numel(), cx.st)
            sv["lvl"][j] = dict(sa=sa, sb=sb, so=so, res=res, t=t, inH=inH, pooled=pooled)
            x_, res = blend, ro
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
        # without pooling every d_cfromDensity conv reads the SAME image: their input gradients add up in place
        dxin_np = cx.zeros(sv["xin"].shape) if (need_input_grad and not self.pool) else None
        for j in range(1, self.stages + 1):
            a, b, out2 = self.blocks[j]
            lv = sv["lvl"][j]
            res, t = lv["res"], lv["t"]
            dpool = cx.zeros(lv["pooled"].shape)
            if dblend is not None:
                cx.call("scale", dpool, dblend, t, dpool.numel(), cx.st)
                dold = cx.buf(dblend.shape)
                cx.call("scale", dold, dblend, 1.0 - t, dold.numel(), cx.st)
                if self.pool:
                    dHj = cx.buf(lv["inH"].shape) if need_input_grad else None
                    self.c_from[2 ** (j - 1)].backward(lv["so"], dold, dx=dHj, param_grads=param_grads)
                else:
                    self.c_from[2 ** (j - 1)].backward(lv["so"], dold, dx=dxin_np, accumulate=True, param_grads=param_grads)
            else:
                dHj = cx.zeros(lv["inH"].shape) if need_input_grad else None
            if j == 1 and self.first:
                cx.call("axpy", dpool, dflat, 1.0, dpool.numel(), cx.st)
            if need_input_grad and self.pool:
                if dH is not None:  # the deeper level's image gradient comes up through this level's pooling
                    cx.call("avgpool2_bwd", dH, dHj, B, res // 2, res // 2, self.cimg, 1, cx.st)
                dH = dHj
            if self.pool:
                dx2 = cx.buf(lv["sb"]["y"].shape)
                cx.call("avgpool2_bwd", dpool, dx2, B, res, res, out2, 0, cx.st)
            else:
                dx2 = dpool
            dx1 = cx.buf(lv["sa"]["y"].shape)
            b.backward(lv["sb"], dx2, dx=dx1, param_grads=param_grads)
            dblend = cx.buf(lv["sa"]["x"].shape)
            a.backward(lv["sa"], dx1, dx=dblend, param_grads=param_grads)
        dxin = None
        if need_input_grad:
            if self.pool:
                dxin = cx.buf(sv["xin"].shape)
                cx.call("avgpool2_bwd", dH, dxin, B, self.S, self.S, self.cimg, 0, cx.st)
            else:
                dxin = dxin_np
            self.c_from[self.u].backward(sv["from_u"], dblend, dx=dxin, accumulate=True, param_grads=param_grads)
        else:
            self.c_from[self.u].backward(sv["from_u"], dblend, param_grads=param_grads)
        return dxin

    # ------------------------------------------------------------------ gradient penalty
    def gradient_penalty(self, in_low, y_gp, percentage, lam=10.0, target=1.0, loss=None):
        """WGAN-GP term of :1120-1138 for the interpolated samples y_gp [B, S*S] of the spatial critic: adds the penalty to
        `loss` (device double) and its parameter gradient to ps.gw. Returns the gradient norms."""
        return self._gp(self._input(in_low, y_gp), percentage, lam, target, loss)

    def _gp(self, xin, percentage, lam=10.0, target=1.0, loss=None):
        """Gradient penalty at the critic input image xin [B,S,S,cimg]. Spatial critic: the gradient w.r.t. channel 1 (the
        sample; `tf.gradients(d_out_loss, [y_gp_d, x_disc])[0]`). Temporal critic (:1279-1285): the samples are
        [B, S*S, 3], so reduce_sum(axis=1) gives one norm per (sample, frame)."""
        cx, ps, B, S = self.cx, self.ps, xin.shape[0], self.S
        loss = self.losses[1:2] if loss is None else loss
        logits, sv = self.forward_from_input(xin, percentage)
        dl = cx.buf((B, 1))
        dl.fill_(1.0 / B)  # d_out_loss = reduce_mean(d_out)
        dxin = self.backward(sv, dl, need_input_grad=True, param_grads=False)
        if self.kind == "tempo":
            gT, vT, norms = cx.buf((B, 3, S * S)), cx.buf((B, 3, S * S)), cx.buf((B * 3,))
            capi.transpose3d(cx.h, dxin, gT, (B, S * S, 3), (0, 2, 1), 0.0, cx.st)
            cx.call("gp_penalty", gT, vT, loss, norms, B * 3, S * S, float(lam), float(target), cx.st)
            txin = cx.buf((B, S, S, 3))
            capi.transpose3d(cx.h, vT, txin, (B, 3, S * S), (0, 2, 1), 0.0, cx.st)
        else:
            g = cx.buf((B, S * S))
            cx.call("take_channel", dxin, g, B * S * S, 2, 1, 0, cx.st)  # tf.gradients(..., [y_gp_d, x_disc])[0]
            v = cx.buf((B, S * S))
            if self.pool:
                norms = cx.buf((B,))
                cx.call("gp_penalty", g, v, loss, norms, B, S * S, float(lam), float(target), cx.st)
            else:
                # upsampling_mode 1 / 3 keep the samples as [B, S, S, 1] images, so the reference's reduce_sum(axis=1) (:1130)
                # sums over the image rows only: one norm per (sample, column). Columns become rows for the kernel and back.
                gT, vT, norms = cx.buf((B, S, S)), cx.buf((B, S, S)), cx.buf((B * S,))
                capi.transpose3d(cx.h, g, gT, (B, S, S), (0, 2, 1), 0.0, cx.st)
                cx.call("gp_penalty", gT, vT, loss, norms, B * S, S, float(lam), float(target), cx.st)
                capi.transpose3d(cx.h, vT, v, (B, S, S), (0, 2, 1), 0.0, cx.st)
            # tangent pass of (0, v) through the linearised critic; every layer adds wgrad(tangent input, primal delta)
            txin = cx.zeros((B, S, S, 2))
            capi.pack_channels(cx.h, [(cx.zeros((B, S * S)), capi.F32, 1, 0, 1, 1, 1), (v, capi.F32, 1, 0, 1, 1, 1)], txin,
                               capi.F32, 2, B, S, S, cx.st)
        t_x = self.c_from[self.u].tangent(txin, sv["from_u"])
        tH, res = txin, S
        for j in range(self.stages, 0, -1):
            a, b, out2 = self.blocks[j]
            lv = sv["lvl"][j]
            if self.pool:
                tH = self._pool(tH, B, res, res, self.cimg)
            t1 = a.tangent(t_x, lv["sa"])
            t2 = b.tangent(t1, lv["sb"])
            tp = self._pool(t2, B, res, res, out2) if self.pool else t2
            if "dlin" in lv["so"]:
                told = self.c_from[2 ** (j - 1)].tangent(tH, lv["so"])
                t_x = cx.buf(tp.shape)
                cx.call("lerp", t_x, told, tp, lv["t"], t_x.numel(), cx.st)
            else:  # the blend of this level never reached the logits (firstNNArch cursor quirk, j = 1)
                t_x = tp
            if j == 1 and self.first:
                t_flat = tp
            if self.pool:
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
        cx = self.cx
        cx.st = torch.cuda.current_stream(cx.device).cuda_stream
        lf = lerp_factor.to(device=cx.device, dtype=torch.float32).view(-1, 1)
        y_gp = (lf * y_real + (1.0 - lf) * y_fake).contiguous()
        return self._critic(self._input(in_low, y_real), self._input(in_low, y_fake), self._input(in_low, y_gp), percentage,
                            weight_dld, lam, target, eps)

    def critic_step_frames(self, real, fake, percentage, lerp_factor, weight_dld=1.0, lam=10.0, target=1.0, eps=0.001):
        """t_disc_loss of :1262-1289 for the temporal critic: real / fake [B, S*S*3] rows of three aligned frames per pixel
        (y_resampled / g_resampled of :1213-1214, 1232-1233), lerp_factor [B, 1] (tf.random_uniform([B, 1, 1]))."""
        cx, S = self.cx, self.S
        if self.kind != "tempo":
            raise ValueError("critic_step_frames is the temporal critic's step")
        cx.st = torch.cuda.current_stream(cx.device).cuda_stream
        B = real.shape[0]
        lf = lerp_factor.to(device=cx.device, dtype=torch.float32).view(-1, 1)
        gp = (lf * real + (1.0 - lf) * fake).contiguous()
        return self._critic(real.view(B, S, S, 3), fake.view(B, S, S, 3), gp.view(B, S, S, 3), percentage, weight_dld, lam, target,
                            eps)

    def _critic(self, x_real, x_fake, x_gp, percentage, weight_dld, lam, target, eps):
        cx, ps = self.cx, self.ps
        self.refresh()
        ps.gw.zero_()
        self.losses.zero_()
        disc, sv_r = self.forward_from_input(x_real, percentage)
        gen, sv_f = self.forward_from_input(x_fake, percentage)
        dl_r, dl_f = cx.buf(disc.shape), cx.buf(gen.shape)
        cx.call("mean_pow", disc, -float(weight_dld), 1, self.losses[0:1], dl_r, disc.numel(), 0, cx.st)
        cx.call("mean_pow", disc, float(eps), 2, self.losses[0:1], dl_r, disc.numel(), 1, cx.st)
        cx.call("mean_pow", gen, 1.0, 1, self.losses[0:1], dl_f, gen.numel(), 0, cx.st)
        self.backward(sv_r, dl_r)
        self.backward(sv_f, dl_f)
        self._gp(x_gp, percentage, lam, target, loss=self.losses[1:2])
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


class GrowingGen:
    """growing_gen in TRAINING mode (GAN/multipassGAN-8x.py:626-750, output=False; pixel_norm, no batch norm): resBlocks
    A(k, relu) -> pixel_norm -> B(k) ; s(1x1) ; pixel_norm(relu(B + s)), a density output per stage (1x1, gain 1) plus the
    residual input density, blended with the density of the previous stage by lerp(old, new, percentage - (j-1)).
    Forward and backward.
    * upsampling_mode 2 / firstNNArch (first network): input rows [B, L*L*C], nearest x2 per stage, 5 / 3 / 2 resBlocks,
      TF1-bicubic residual of input channel 0, the previous stage's density nearest-upsampled before the blend.
    * upsampling_mode 1 / 3 (refinement networks, use_res_net): input rows [B, S*S*(C+1)] (concat(first-pass density, resized
      low-res fields), :1042-1044), two head resBlocks (:715-716), two resBlocks per stage, no resampling anywhere, the
      residual is input channel 0 itself (:738-739)."""

    def __init__(self, tileSizeLow=16, upRes=8, n_inputChannels=6, start_fms=256, max_fms=256, filterSize=3, batch=16,
                 addBicubicUpsample=True, values=None, seed=1, device=0, cx=None, upsampling_mode=2):
        self.cx = cx if cx is not None else _Ctx(device)
        if int(upsampling_mode) not in (1, 2, 3):
            raise ValueError("upsampling_mode %r: built are 2 (first network) and 1 / 3 (refinement networks)" % (upsampling_mode,))
        self.up = int(upsampling_mode) == 2
        self.L, self.u, self.S, self.C = int(tileSizeLow), int(upRes), int(tileSizeLow) * int(upRes), int(n_inputChannels)
        self.stages = int(round(math.log(self.u, 2)))
        self.bicubic = bool(addBicubicUpsample)
        self.ps = ps = ParamSet(self.cx.device)
        cx, k = self.cx, int(filterSize)
        sc = "generator/"

        def rb(pre, name, cin, s1, s2):
            return (_Conv8(cx, ps, pre + "g_cA_" + name, k, cin, s1, "relu"), _Conv8(cx, ps, pre + "g_cB_" + name, k, s1, s2, None),
                    _Conv8(cx, ps, pre + "g_s_" + name, 1, cin, s2, None))

        self.cin0 = cin = self.C if self.up else self.C + 1
        self.head = []
        if not self.up:
            m = min(int(max_fms), int(start_fms) // 2)
            for s1, s2, name in ((16, m // 8, "1"), (m // 4, m // 2, "2")):
                self.head.append(rb(sc, name, cin, s1, s2))
                cin = s2
        self.c_dens = {1: _Conv8(cx, ps, sc + "g_cdensOut1", 1, cin, 1, None, gain=1.0)}
        self.blocks = {}
        for j in range(1, self.stages + 1):
            fms = min(int(start_fms / (2 ** j)), max_fms)
            up = 2 ** j
            if not self.up:
                plan = [(fms, fms, "first"), (fms // 2, fms // 2, "second")]
            elif up == 2:
                plan = [(fms, fms, n) for n in ("first", "second", "third", "fourth", "fifth")]
            elif up == 4:
                plan = [(fms * 2, fms, "first"), (fms, fms, "second"), (fms, fms, "third")]
            else:
                plan = [(fms * 2, fms, "first"), (fms, fms, "second")]
            rbs = []
            for s1, s2, name in plan:
                rbs.append(rb(sc + "genBlock%d/" % up, name, cin, s1, s2))
                cin = s2
            self.blocks[j] = rbs
            self.c_dens[up] = _Conv8(cx, ps, sc + "genBlock%d/g_cdensOut%d" % (up, up), 1, cin, 1, None, gain=1.0)
        vals = dict(values) if values else {}
        for name, shape, _, _, _ in ps.specs:
            if name not in vals:
                vals[name] = W.init_variable(seed, name, shape, "normal" if name.endswith("/weight") else ("const", 0.1))
        ps.finalize(vals)
        self.bic_plans = {}

    def refresh(self):
        ps = self.ps
        self.cx.call("mul", ps.w, ps.v, ps.scale, ps.total, self.cx.st)

    def _up2(self, x, B, h, w, c):
        y = self.cx.buf((B, 2 * h, 2 * w, c))
        capi.pack_channels(self.cx.h, [(x, capi.F32, c, 0, c, 2, 2)], y, capi.F32, c, B, 2 * h, 2 * w, self.cx.st)
        return y

    def _up2_bwd(self, dy, B, h, w, c):
        """backward of the nearest x2: sum over the 2x2 replicas."""
        cx = self.cx
        t = cx.buf((B, h, w, c))
        cx.call("avgpool2_fwd", dy, t, B, 2 * h, 2 * w, c, cx.st)
        cx.call("scale", t, t, 4.0, t.numel(), cx.st)
        return t

    def _pn(self, x):
        y = self.cx.buf(x.shape)
        self.cx.call("pixel_norm_fwd", x, y, x.numel() // x.shape[-1], x.shape[-1], self.cx.st)
        return y

    def _rb_fwd(self, rb, inp, B, res):
        cx = self.cx
        a, b, s = rb
        ya, sa = a.forward(inp, B, res, res)
        yap = self._pn(ya)
        yb, sb = b.forward(yap, B, res, res)
        ys, ss = s.forward(inp, B, res, res)
        r = cx.buf(yb.shape)
        cx.call("add_act_fwd", yb, ys, r, r.numel(), capi.ACT_RELU, cx.st)
        return self._pn(r), dict(sa=sa, sb=sb, ss=ss, ya=ya, r=r)

    def _rb_bwd(self, rb, q, d, need_input_grad):
        """d: gradient w.r.t. the block's (pixel-normed) output; returns the gradient w.r.t. its input (None for data)."""
        cx = self.cx
        a, b, s = rb
        d_r = cx.buf(q["r"].shape)
        cx.call("pixel_norm_bwd", q["r"], d, d_r, q["r"].numel() // q["r"].shape[-1], q["r"].shape[-1], cx.st)
        d_sum = cx.buf(q["r"].shape)
        cx.call("act_bwd", q["r"], d_r, d_sum, d_sum.numel(), capi.ACT_RELU, cx.st)
        d_yap = cx.buf(q["ya"].shape)
        b.backward(q["sb"], d_sum, dx=d_yap)
        d_inp = cx.buf(q["sa"]["x"].shape) if need_input_grad else None
        s.backward(q["ss"], d_sum, dx=d_inp)
        d_ya = cx.buf(q["ya"].shape)
        cx.call("pixel_norm_bwd", q["ya"], d_yap, d_ya, q["ya"].numel() // q["ya"].shape[-1], q["ya"].shape[-1], cx.st)
        a.backward(q["sa"], d_ya, dx=d_inp, accumulate=True)
        return d_inp

    def forward(self, x_rows, percentage):
        """x_rows: device fp32 [B, L*L*C] (upsampling_mode 2) or [B, S*S*(C+1)] (1 / 3) -> (gen rows [B, S*S], saved)."""
        cx, B, L, C = self.cx, x_rows.shape[0], self.L, self.C
        res = L if self.up else self.S
        x0 = x_rows.view(B, res, res, self.cin0)
        sv = dict(B=B, lvl={}, pct=float(percentage), head=[])
        cur, ch = x0, self.cin0
        for rb in self.head:
            cur, q = self._rb_fwd(rb, cur, B, res)
            sv["head"].append(q)
            ch = cur.shape[-1]
        old, sv["d1"] = self.c_dens[1].forward(cur, B, res, res)
        for j in range(1, self.stages + 1):
            if self.up:
                inp = self._up2(cur, B, res, res, ch)
                res *= 2
            else:
                inp = cur
            rbs_sv = []
            for rb in self.blocks[j]:
                inp, q = self._rb_fwd(rb, inp, B, res)
                rbs_sv.append(q)
            cur, ch = inp, inp.shape[-1]
            dens, sd = self.c_dens[2 ** j].forward(cur, B, res, res)
            if self.bicubic:
                out = cx.buf(dens.shape)
                if self.up:
                    key = (L, res)
                    if key not in self.bic_plans:
                        self.bic_plans[key] = capi.BicubicPlan(cx.h, L, L, res, res)
                    capi.dens_residual(cx.h, dens, x0, capi.F32, C, 0, 2, self.bic_plans[key], B, res, res, L, L, out, cx.st)
                else:
                    capi.dens_residual(cx.h, dens, x0, capi.F32, self.cin0, 0, 0, None, B, res, res, res, res, out, cx.st)
                dens = out
            oldu = self._up2(old, B, res // 2, res // 2, 1) if self.up else old
            t = float(min(max(percentage - (j - 1), 0.0), 1.0))
            blend = cx.buf(dens.shape)
            cx.call("lerp", blend, oldu, dens, t, blend.numel(), cx.st)
            sv["lvl"][j] = dict(rbs=rbs_sv, sd=sd, t=t, res=res)
            old = blend
        return old.view(B, self.S * self.S), sv

    def backward(self, sv, dout):
        """dout [B, S*S]: gradient w.r.t. the generated rows. Accumulates the parameter gradients into ps.gw."""
        cx, B = self.cx, sv["B"]
        d_old = dout.reshape(B, self.S, self.S, 1)
        d_up_next = None  # gradient w.r.t. the (upsampled) input of the next finer stage
        for j in range(self.stages, 0, -1):
            lv = sv["lvl"][j]
            res, t = lv["res"], lv["t"]
            d_dens = cx.buf(d_old.shape)
            cx.call("scale", d_dens, d_old, t, d_dens.numel(), cx.st)
            d_oldu = cx.buf(d_old.shape)
            cx.call("scale", d_oldu, d_old, 1.0 - t, d_oldu.numel(), cx.st)
            ch = lv["sd"]["x"].shape[-1]
            d_cur = cx.buf(lv["sd"]["x"].shape)
            self.c_dens[2 ** j].backward(lv["sd"], d_dens, dx=d_cur)
            if d_up_next is not None:
                cx.call("axpy", d_cur, self._up2_bwd(d_up_next, B, res, res, ch) if self.up else d_up_next, 1.0, d_cur.numel(),
                        cx.st)
            d = d_cur
            rbs = self.blocks[j]
            for i in range(len(rbs) - 1, -1, -1):
                # the first block of the first network reads the upsampled DATA: no input gradient needed
                d = self._rb_bwd(rbs[i], lv["rbs"][i], d, not (self.up and j == 1 and i == 0))
            d_up_next = d
            d_old = self._up2_bwd(d_oldu, B, res // 2, res // 2, 1) if self.up else d_oldu
        if not self.head:
            self.c_dens[1].backward(sv["d1"], d_old)
            return
        d = cx.buf(sv["d1"]["x"].shape)
        self.c_dens[1].backward(sv["d1"], d_old, dx=d)
        cx.call("axpy", d, d_up_next, 1.0, d.numel(), cx.st)
        for i in range(len(self.head) - 1, -1, -1):
            d = self._rb_bwd(self.head[i], sv["head"][i], d, i > 0)

    def grads(self):
        ps = self.ps
        host = ps.g.cpu().numpy()
        return {name: host[off:off + n].reshape(shape).copy() for name, shape, off, n, _ in ps.specs}


class StageBatches:
    """`batches(currentUpres)` for Trainer8x.train out of one tilesampler.TileSampler per data resolution. The reference
    builds a new TileCreator with `upres=currentUpres` and re-loads `density_low_<currentUpres>_%04d.uni` as the targets at
    every growing event (GAN/multipassGAN-8x.py:1916-1960); here the frames of every stage stay resident on the device and
    the growing event only switches the sampler. getinput (:1497-1536) = selectRandomTiles + the reshape to rows."""

    def __init__(self, samplers, batch_size, augment=False, tile_t=None):
        """tile_t: 1 when the samplers hold three-frame sequences (TileSampler(dim_t=3), the temporal training data): getinput
        then draws single frames of them, as the reference's selectRandomTiles default does."""
        self.samplers, self.batch_size, self.augment, self.tile_t = dict(samplers), int(batch_size), bool(augment), tile_t

    def __call__(self, currentUpres):
        if currentUpres not in self.samplers:
            raise KeyError("no training data at %dx (have %s)" % (currentUpres, sorted(self.samplers)))
        return self.samplers[currentUpres].batch_rows(self.batch_size, augment=self.augment, tile_t=self.tile_t)


class TempoBatches:
    """`tempo_batches(currentUpres)` for Trainer8x.train: getTempoinput (GAN/multipassGAN-8x.py:1475-1495) =
    TileCreator.selectRandomTempoTiles (tools_wscale/tilecreator_t.py:1382-1413). The samplers hold
    THREE-frame data (TileCreator(dim_t=3): the frames of a sequence stored as channel groups, low [N,1,L,L,C*3], high
    [N,1,S,S,3]); a batch is `batch_size // 3` random three-frame tiles (same random decisions as selectRandomTiles), re-ordered
    to rows (sample, frame), plus the semi-Lagrangian re-sampling positions of every frame (getSemiLagrPosBatch :1341-1378,
    kernel mpg_train_semilagr_pos) with dt * (+1, 0, -1): the neighbouring frames are pulled onto the middle one."""

    def __init__(self, samplers, batch_size, n_t=3, dt=0.5, vel_channel=1, device=0, augment=False):
        """augment: selectRandomTiles(..., augment=True) -> generateTile on the three-frame tiles (samplers built with
        dim_t=3 and init_data_augmentation: scaling / rot90 / flip with every frame's velocity vectors fixed up)."""
        self.samplers, self.batch_size, self.n_t, self.dt = dict(samplers), int(batch_size), int(n_t), float(dt)
        self.c0, self.h, self.augment = int(vel_channel), capi.default_handle(device), bool(augment)

    def __call__(self, currentUpres):
        if currentUpres not in self.samplers:
            raise KeyError("no three-frame training data at %dx (have %s)" % (currentUpres, sorted(self.samplers)))
        s, n_t = self.samplers[currentUpres], self.n_t
        B = max(1, self.batch_size // n_t)
        low, high = s.select_random_tiles(B, True, self.augment)       # [B,1,T,T,C*n_t], [B,1,Su,Su,n_t]
        T, Su = low.shape[2], high.shape[2]
        C = low.shape[-1] // n_t
        if low.shape[-1] != C * n_t or high.shape[-1] != n_t:
            raise ValueError("TempoBatches needs %d-frame data: low channels C*%d, high channels %d" % (n_t, n_t, n_t))
        x = low.reshape(B, T, T, n_t, C).permute(0, 3, 1, 2, 4).contiguous().view(B * n_t, T * T * C)
        y = high.reshape(B, Su, Su, n_t).permute(0, 3, 1, 2).contiguous().view(B * n_t, Su * Su)
        pos = torch.empty((B * n_t, Su * Su * 2), dtype=torch.float32, device=x.device)
        st = torch.cuda.current_stream(x.device).cuda_stream
        capi.train_call("semilagr_pos", self.h, x, pos, B * n_t, T, Su, C, self.c0, self.dt, n_t, st)
        return x, y, pos


class Trainer8x:
    """Loop body of the 8x progressive-growing training (GAN/multipassGAN-8x.py:1898-2075, spatial part): one critic step
    (WGAN-GP, :1111-1143) and one generator step (g_loss_d + lambda * l1, :1117,1145) with the optimizers of growing stage z
    (:1304-1362) and the generator weight EMA. Temporal discriminator terms (lambda_t) and loss scaling are not built.
    upsampling_mode 2 + firstNNArch = the first network (GAN/example_run_training.py:4): x rows [B, L*L*C], y rows
    [B, S*S].  upsampling_mode 1 / 3 = the refinement networks (:7): y rows [B, S*S*2] carry (target density, first-pass
    density) per pixel; the generator input is concat(first-pass density, nearest-resized low-res fields) (:1042-1044), the
    target is channel 0 (:1061-1062) and the critic's gradient penalty reduces over image rows (GrowingDisc)."""

    def __init__(self, tileSizeLow=16, upRes=8, n_inputChannels=6, start_fms=256, max_fms=256, filterSize=3, batch=16,
                 learning_rate=1e-4, adam_beta1=0.0, adam_beta2=0.99, lambda_l1=1.0, values=None, seed=1, device=0,
                 upsampling_mode=2, first_nn_arch=None, lambda_t=0.0, group=None):
        """group: torch.distributed process group for data-parallel training (one process per GPU, every rank draws its own
        batches): the flat gradient of each optimizer step is averaged over the ranks with ONE all-reduce
        (parallel.allreduce_mean, as in Trainer4x); None with an initialised default group = all ranks."""
        self.cx = cx = _Ctx(device)
        self.group = group
        self.refine = int(upsampling_mode) != 2
        first = (not self.refine) if first_nn_arch is None else bool(first_nn_arch)
        if first == self.refine:
            raise ValueError("built: firstNNArch 1 with upsamplingMode 2, firstNNArch 0 with upsamplingMode 1 / 3")
        self.gen = GrowingGen(tileSizeLow, upRes, n_inputChannels, start_fms, max_fms, filterSize, batch, True, values, seed, device,
                              cx=cx, upsampling_mode=upsampling_mode)
        self.disc = GrowingDisc(tileSizeLow, upRes, n_inputChannels, start_fms, max_fms, filterSize, first, batch, values, seed,
                                device, cx=cx, upsampling_mode=upsampling_mode)
        n = self.gen.stages
        self.opt_g = StagedAdam(cx, self.gen.ps, [learning_rate] * n, adam_beta1, adam_beta2, n_stages=n)
        self.opt_d = StagedAdam(cx, self.disc.ps, [learning_rate] * n, adam_beta1, adam_beta2, n_stages=n)
        self.ema = WeightEMA(self.gen.ps, 0.999)
        # temporal critic (lambda_t > 1e-6 = useTempoD, :196-201): growing_disc_tempo on three aligned frames, its own staged
        # optimizers (t_adam_*, :1318-1330) and the generator term kkt * mean(-T(G frames)) (:1296-1302)
        self.k_t = float(lambda_t)
        self.tdisc = self.opt_t = None
        if self.k_t > 1e-6:
            self.tdisc = GrowingDisc(tileSizeLow, upRes, n_inputChannels, start_fms, max_fms, filterSize, first, batch, values, seed,
                                     device, cx=cx, upsampling_mode=upsampling_mode, kind="tempo")
            self.opt_t = StagedAdam(cx, self.tdisc.ps, [learning_rate] * n, adam_beta1, adam_beta2, n_stages=n)
        self.k_l1 = float(lambda_l1)
        self.learning_rate = float(learning_rate)
        self.losses = torch.zeros(4, dtype=torch.float64, device=cx.device)
        self.save_no = 0

    def _inputs(self, x_rows, y_rows):
        """(generator input rows, target rows) of the training graph (:1040-1062)."""
        if not self.refine:
            return x_rows, y_rows
        cx, g = self.cx, self.gen
        B, S, L, C = x_rows.shape[0], g.S, g.L, g.C
        if y_rows.shape[1] != S * S * 2:
            raise ValueError("refinement networks train on target rows of S*S*2 values (target, first-pass density)")
        x_in = cx.buf((B, S * S * (C + 1)))
        capi.pack_channels(cx.h, [(y_rows, capi.F32, 2, 1, 1, 1, 1), (x_rows, capi.F32, C, 0, C, g.u, g.u)], x_in, capi.F32, C + 1,
                           B, S, S, cx.st)
        y_in = cx.buf((B, S * S))
        cx.call("take_channel", y_rows, y_in, B * S * S, 2, 0, 0, cx.st)
        return x_in, y_in

    def disc_step(self, x_rows, y_rows, percentage, z, lerp_factor):
        cx = self.cx
        cx.st = torch.cuda.current_stream(cx.device).cuda_stream
        self.gen.refresh()
        x_in, y_in = self._inputs(x_rows, y_rows)
        gen_y, _ = self.gen.forward(x_in, percentage)
        out = self.disc.critic_step(x_rows, y_in, gen_y, percentage, lerp_factor)
        par.allreduce_mean(self.disc.ps.g, self.group)
        self.opt_d.step(z)
        return out

    def _frames(self, rows):
        """[B*3, S*S] rows ordered (sample, frame) -> [B, S*S*3]: `transpose(reshape(., [-1, 3, n_output]), [0, 2, 1])` of
        :1213-1214 / :1232-1233 (three aligned frames per pixel, the temporal critic's channel axis)."""
        cx, n = self.cx, self.gen.S * self.gen.S
        if rows.shape[0] % 3 or rows.shape[1] != n:
            raise ValueError("temporal batches are 3 consecutive frames per sample: [B*3, S*S] rows")
        B = rows.shape[0] // 3
        out = cx.buf((B, n * 3))
        capi.transpose3d(cx.h, rows, out, (B, 3, n), (0, 2, 1), 0.0, cx.st)
        return out

    def _align(self, rows, y_pos):
        """tensorResample (:545-594) of the frames [n, S*S] at the advected positions y_pos [n, cur*cur*2] (adv_flag 1,
        adv_mode 0: :1195-1197 for the generated frames, :1241-1242 for the targets). y_pos None = frames already aligned
        (adv_flag 0). Positions live on the grid of the stage's tiles (cur = tileSize * currentUpres): as in :1192-1204 the
        frames are nearest-resized to that grid (a strided pick), re-sampled, and nearest-resized back to the full tile."""
        if y_pos is None:
            return rows
        cx, S, n = self.cx, self.gen.S, rows.shape[0]
        cur = int(round(math.sqrt(y_pos.shape[1] // 2)))
        if y_pos.shape[0] != n or cur * cur * 2 != y_pos.shape[1] or S % cur:
            raise ValueError("y_pos must be [B*3, cur*cur*2] positions on a grid dividing the tile (got %s)" % (tuple(y_pos.shape),))
        f = S // cur
        src = rows
        if f > 1:
            src = cx.buf((n, cur * cur))
            cx.call("pick", rows, src, n, cur, cur, 1, f, 1, 1, 0, cx.st)
        out = cx.buf((n, cur * cur))
        cx.call("resample_fwd", src, y_pos, out, n, cur, cur, 1, cx.st)
        if f > 1:
            up = cx.buf((n, S * S))
            capi.pack_channels(cx.h, [(out, capi.F32, 1, 0, 1, f, f)], up, capi.F32, 1, n, S, S, cx.st)
            out = up
        return out

    def _align_bwd(self, d_rows, y_pos):
        """Adjoint of _align: gradient w.r.t. the un-aligned frames [n, S*S]."""
        cx, S, n = self.cx, self.gen.S, d_rows.shape[0]
        cur = int(round(math.sqrt(y_pos.shape[1] // 2)))
        f = S // cur
        d = d_rows
        res = S
        while res > cur:                     # adjoint of the nearest resize up: sum over the replicas, a factor 2 at a time
            t = cx.buf((n, (res // 2) ** 2))
            cx.call("avgpool2_fwd", d, t, n, res, res, 1, cx.st)
            cx.call("scale", t, t, 4.0, t.numel(), cx.st)
            d, res = t, res // 2
        d_src = cx.zeros((n, cur * cur))
        cx.call("resample_bwd", d, y_pos, d_src, n, cur, cur, 1, cx.st)
        if f == 1:
            return d_src
        out = cx.zeros((n, S, S))            # adjoint of the strided pick: the gradient lands on the sampled positions
        out[:, ::f, ::f] = d_src.view(n, cur, cur)
        return out.view(n, S * S)

    def t_disc_step(self, x_t_rows, y_t_rows, percentage, z, lerp_factor, y_pos=None):
        """One step of t_disc_optimizer[z] (:2001-2013) on frame triplets: x_t_rows [B*3, L*L*C], y_t_rows [B*3, S*S(*2)]
        ordered (sample, frame); y_pos: see _align (None = the triplets are already aligned, adv_flag 0 :1229-1231)."""
        if self.tdisc is None:
            raise ValueError("lambda_t is 0: no temporal critic")
        cx = self.cx
        cx.st = torch.cuda.current_stream(cx.device).cuda_stream
        self.gen.refresh()
        x_in, y_in = self._inputs(x_t_rows, y_t_rows)
        gen_ts, _ = self.gen.forward(x_in, percentage)
        out = self.tdisc.critic_step_frames(self._frames(self._align(y_in.contiguous(), y_pos)),
                                            self._frames(self._align(gen_ts, y_pos)), percentage, lerp_factor)
        par.allreduce_mean(self.tdisc.ps.g, self.group)
        self.opt_t.step(z)
        return out

    def gen_step(self, x_rows, y_rows, percentage, z, x_t_rows=None, y_t_rows=None, y_pos=None):
        """gen_optimizer[z] (:2015-2043) on gen_loss_complete = g_loss_d + kk * l1 [+ kkt * g_loss_t on the temporal batch].
        Returns the device doubles [g_loss_d, kk * l1, kkt * g_loss_t, 0]."""
        cx, g, d = self.cx, self.gen, self.disc
        cx.st = torch.cuda.current_stream(cx.device).cuda_stream
        g.refresh()
        d.refresh()
        g.ps.gw.zero_()
        self.losses.zero_()
        x_in, y_in = self._inputs(x_rows, y_rows)
        gen_y, gsv = g.forward(x_in, percentage)
        logits, dsv = d.forward(x_rows, gen_y, percentage)
        dl = cx.buf(logits.shape)
        cx.call("mean_pow", logits, -1.0, 1, self.losses[0:1], dl, logits.numel(), 0, cx.st)      # g_loss_d = mean(-gen) :1117
        dxin = d.backward(dsv, dl, need_input_grad=True, param_grads=False)
        dgen = cx.buf(gen_y.shape)
        cx.call("l1_mean", y_in, gen_y, self.k_l1, self.losses[1:2], dgen, gen_y.numel(), 0, cx.st)  # lambda * mean|y - G| :1099,1145
        cx.call("take_channel", dxin, dgen, gen_y.numel(), 2, 1, 1, cx.st)
        g.backward(gsv, dgen)
        if self.tdisc is not None and x_t_rows is not None:
            # + kkt * mean(-T(G(x_t))) (:1296-1302): a second pass of the same generator on the frame triplets; its parameter
            # gradients accumulate onto those of the spatial terms
            t = self.tdisc
            t.refresh()
            x_in_t, _ = self._inputs(x_t_rows, y_t_rows)
            gen_ts, gsv_t = g.forward(x_in_t, percentage)
            B, S = gen_ts.shape[0] // 3, g.S
            logits_t, tsv = t.forward_from_input(self._frames(self._align(gen_ts, y_pos)).view(B, S, S, 3), percentage)
            dl_t = cx.buf(logits_t.shape)
            cx.call("mean_pow", logits_t, -self.k_t, 1, self.losses[2:3], dl_t, logits_t.numel(), 0, cx.st)
            dx_t = t.backward(tsv, dl_t, need_input_grad=True, param_grads=False)
            dgen_t = cx.buf(gen_ts.shape)
            capi.transpose3d(cx.h, dx_t, dgen_t, (B, S * S, 3), (0, 2, 1), 0.0, cx.st)
            if y_pos is not None:  # back through tensorResample: scatter-add onto the un-aligned frames
                dgen_t = self._align_bwd(dgen_t, y_pos)
            g.backward(gsv_t, dgen_t)
        cx.call("mul", g.ps.g, g.ps.gw, g.ps.scale, g.ps.total, cx.st)
        par.allreduce_mean(g.ps.g, self.group)
        self.opt_g.step(z)
        self.ema.update(self.opt_g.state[z]["mask"])
        return self.losses

    # ------------------------------------------------------------------ the loop around the two steps
    def target_rows(self, y_rows):
        """y_in of :1057-1058: the stage's target tiles [B, (L * currentUpres)^2] nearest-resized to the full tile size S.
        (The reference takes the stage's tile size from 2 ** ceil(percentage); its schedule runs one stage ahead of its data
        (schedule8x doc), so the size is taken from the rows themselves here.)"""
        cx, S, B = self.cx, self.gen.S, y_rows.shape[0]
        if self.refine:
            return y_rows                 # upsampling_mode 1 / 3: the data is at the full resolution from the start (:1061)
        cur = int(round(math.sqrt(y_rows.shape[1])))
        if cur * cur != y_rows.shape[1] or S % cur:
            raise ValueError("target rows of %d values are not square tiles dividing %d" % (y_rows.shape[1], S))
        if cur == S:
            return y_rows
        out = cx.buf((B, S * S))
        capi.pack_channels(cx.h, [(y_rows, capi.F32, 1, 0, 1, S // cur, S // cur)], out, capi.F32, 1, B, S, S, cx.st)
        return out

    def train(self, batches, schedule, discRuns=1, genRuns=1, lambda_f=1.0, add_adj_idcs=True, zero_density=True, save_dir=None,
              saveInterval=200, alwaysSave=True, on_grow=None, log=None, log_interval=0, lerp_seed=0, max_iters=None,
              tempo_batches=None):
        """The training loop of GAN/multipassGAN-8x.py:1898-2089 (spatial part). `batches(currentUpres)` returns one batch
        (x_rows [B, L*L*C], y_rows [B, (L*currentUpres)^2]) of device fp32 rows (getinput, :1497); `schedule` is a
        schedule8x.GrowthSchedule. Per iteration: discRuns critic steps, genRuns generator steps (each on a fresh batch) with
        the optimizers of the stage's index, blend value and decayed learning rate; at a growing event the model is saved and
        `on_grow(new_upres)` is called (the reference re-loads its data there, :1916-1963); the model is saved when
        `(disc_cost + gen_cost < lastCost or alwaysSave) and lastSave >= saveInterval` (:2076-2084).  The critic's
        interpolation factors are torch's uniform numbers (TF's random stream is not reproducible).  With lambda_t > 0 and
        `tempo_batches(currentUpres)` -> (x_t rows [B*3, ...], y_t rows [B*3, ...][, y_pos [B*3, S*S*2]]) frame triplets
        (getTempoinput; without y_pos they are taken as aligned, adv_flag 0), every iteration also runs discRuns temporal-critic steps (:2001-2013) and the generator step carries
        the temporal term.  Returns a list of
        (it, disc_loss, g_loss_d, l1) at the logged iterations (log_interval 0: only the last iteration is read back)."""
        cx = self.cx
        gen_rng = torch.Generator(device=cx.device)
        gen_rng.manual_seed(int(lerp_seed))
        last_save, last_cost = 1, 1e10
        kkin = self.k_l1
        history, done = [], 0
        d_loss = g_loss = None
        n_total = len(schedule) if max_iters is None else min(len(schedule), int(max_iters))

        def batch(upres):
            xs, ys = batches(upres)
            if zero_density:                                                      # :1527-1533 on (copies of) the device rows
                xz, yz = xs.clone(), ys.clone()
                if schedule8x.zero_density_batch(xz.view(xz.shape[0], -1, self.gen.C), yz, add_adj_idcs and self.gen.C >= 6):
                    xs, ys = xz, yz
            return xs, self.target_rows(ys)

        for st in schedule:
            if done >= n_total:
                break
            if st.grew:
                if save_dir is not None:
                    self.save(save_dir)                                             # saveModel(0.0) :1913
                if on_grow is not None:
                    on_grow(st.currentUpres)
            lrs_g, lrs_d = schedule8x.learning_rates(self.learning_rate, st.lrgs, schedule.decayIter, schedule.decayLR,
                                                     self.gen.stages)
            self.opt_g.lrs, self.opt_d.lrs = lrs_g, lrs_d
            if self.opt_t is not None:
                self.opt_t.lrs = list(lrs_d)                                        # learning_rates_t = the same decay (:1005-1006)
            for _ in range(discRuns):
                xs, ys = batch(st.currentUpres)
                lf = torch.rand((xs.shape[0], 1), generator=gen_rng, device=cx.device)
                d_loss = self.disc_step(xs, ys, st.percentage, st.index, lf)
            tempo = self.tdisc is not None and tempo_batches is not None
            if tempo:
                for _ in range(discRuns):
                    xt, yt, pos = (tuple(tempo_batches(st.currentUpres)) + (None,))[:3]
                    lf = torch.rand((xt.shape[0] // 3, 1), generator=gen_rng, device=cx.device)
                    self.t_disc_step(xt, self._tempo_targets(yt), st.percentage, st.index, lf, pos)
            for _ in range(genRuns):
                xs, ys = batch(st.currentUpres)
                kkin = lambda_f * kkin                                              # :2019
                self.k_l1 = kkin
                xt = yt = pos = None
                if tempo:
                    xt, yt, pos = (tuple(tempo_batches(st.currentUpres)) + (None,))[:3]
                    yt = self._tempo_targets(yt)
                g_loss = self.gen_step(xs, ys, st.percentage, st.index, xt, yt, pos).clone()
            done += 1
            read = (log_interval and (st.it + 1) % log_interval == 0) or done == n_total or not alwaysSave
            if read:
                dl = float(d_loss[0]) if d_loss is not None else 0.0
                gl = g_loss.cpu().numpy() if g_loss is not None else np.zeros(4)
                history.append((st.it, dl, float(gl[0]), float(gl[1])))
                if log is not None:
                    log("it %d upres %d stage %d blend %.4f lr %.3e: disc %.5f gen %.5f l1 %.5f" %
                        (st.it, st.currentUpres, st.index, st.percentage, lrs_g[0], dl, gl[0], gl[1]))
            cost = (history[-1][1] + history[-1][2]) if (read and history) else 0.0
            if ((not alwaysSave and cost < last_cost) or alwaysSave) and last_save >= saveInterval:
                last_save = 1
                last_cost = cost
                if save_dir is not None:
                    self.save(save_dir)
            else:
                last_save += 1
        return history

    def _tempo_targets(self, y_t_rows):
        return self.target_rows(y_t_rows)

    def _optimizers(self):
        out = [("g", self.opt_g, self.gen.ps), ("d", self.opt_d, self.disc.ps)]
        if self.tdisc is not None:
            out.append(("t", self.opt_t, self.tdisc.ps))
        return out

    # ------------------------------------------------------------------ checkpoints
    def values(self):
        out = self.gen.ps.export()
        out.update(self.disc.ps.export())
        if self.tdisc is not None:
            out.update(self.tdisc.ps.export())
        return out

    def save(self, test_path):
        """saveModel (:1804-1807): `model_%04d.ckpt` = the variables, `model_ema_%04d.ckpt` = the same names with the generator's
        moving averages swapped in (MovingAverageOptimizer.swapping_saver), both as TF checkpoint-V2 bundles (tfckpt.py) under
        the reference's variable names, which is what multipassGAN-out.py restores (:367-386).  The optimizer moments travel
        in the first file under TF's slot names (`<var>/<g|d>_adam_<2**(z+1)>[_1]`) for resumed runs."""
        import os
        from . import tfckpt
        vals = self.values()
        ema = dict(vals)
        ema.update(self.ema.export())
        full = dict(vals)
        for tag, opt, ps in self._optimizers():
            for z, st in enumerate(opt.state):
                if st["t"] == 0:
                    continue
                m, v, mask = st["m"].cpu().numpy(), st["v"].cpu().numpy(), st["mask"].cpu().numpy()
                for name, shape, off, n, _ in ps.specs:
                    if mask[off] != 0:
                        full["%s/%s_adam_%d" % (name, tag, 2 ** (z + 1))] = m[off:off + n].reshape(shape).copy()
                        full["%s/%s_adam_%d_1" % (name, tag, 2 ** (z + 1))] = v[off:off + n].reshape(shape).copy()
                full["mpg_b200/%s_adam_%d/steps" % (tag, 2 ** (z + 1))] = np.array([st["t"]], np.float32)
        os.makedirs(test_path, exist_ok=True)
        no = self.save_no
        tfckpt.write_checkpoint(os.path.join(test_path, "model_%04d.ckpt" % no), full)
        tfckpt.write_checkpoint(os.path.join(test_path, "model_ema_%04d.ckpt" % no), ema)
        self.save_no += 1
        return no

    def load(self, test_path, no):
        """saver.restore + saver_2.restore of :1373-1377: variables, optimizer moments and the generator's moving averages."""
        import os
        from . import tfckpt
        full = tfckpt.read_checkpoint(os.path.join(test_path, "model_%04d.ckpt" % no), verify_data=True)
        ema = tfckpt.read_checkpoint(os.path.join(test_path, "model_ema_%04d.ckpt" % no), verify_data=True)
        for tag, opt, ps in self._optimizers():
            host = ps.v.cpu().numpy()
            for name, shape, off, n, _ in ps.specs:
                host[off:off + n] = np.asarray(full[name], np.float32).reshape(-1)
            ps.v.copy_(torch.from_numpy(host))
            for z, st in enumerate(opt.state):
                key = "mpg_b200/%s_adam_%d/steps" % (tag, 2 ** (z + 1))
                if key not in full:
                    continue
                st["t"] = int(full[key][0])
                m, v = st["m"].cpu().numpy(), st["v"].cpu().numpy()
                for name, shape, off, n, _ in ps.specs:
                    k = "%s/%s_adam_%d" % (name, tag, 2 ** (z + 1))
                    if k in full:
                        m[off:off + n] = np.asarray(full[k], np.float32).reshape(-1)
                        v[off:off + n] = np.asarray(full[k + "_1"], np.float32).reshape(-1)
                st["m"].copy_(torch.from_numpy(m))
                st["v"].copy_(torch.from_numpy(v))
        ps = self.gen.ps
        host = self.ema.shadow.cpu().numpy()
        for name, shape, off, n, _ in ps.specs:
            host[off:off + n] = np.asarray(ema[name], np.float32).reshape(-1)
        self.ema.shadow.copy_(torch.from_numpy(host))
        self.save_no = int(no) + 1
