"""ORACLE (test infrastructure only): one training iteration of the 4x model on CPU with torch autograd.

Restates GAN/multipassGAN-4x.py: graph wiring :726-730, losses :744-768, variable split :783-787, Adam
optimizers :889-898 (TF1 "epsilon hat" update), BN UPDATE_OPS :776-779 and the loop body
:1316-1397 (discRuns discriminator steps, then genRuns generator steps; every sess.run re-evaluates the whole
graph with train=True).  Temporal terms are out of scope (lambda_t 0).
"""
import math

import numpy as np
import torch

from . import gan as og
from . import networks as on


class TrainContext(og.Context):
    """Context whose variables are persistent autograd leaves."""

    def __init__(self, store, dtype):
        super().__init__(store, dtype)
        self.leaves = {}

    def var(self, leaf, shape, kind):
        name = "/".join(self.scopes + [leaf])
        if name not in self.leaves:
            t = torch.as_tensor(self.store.get(name, shape, kind)).to(self.dtype).clone()
            t.requires_grad_(leaf not in ("moving_mean", "moving_variance"))
            self.leaves[name] = t
        return self.leaves[name], name


def bce_logits(x, z):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log(1 + exp(-|x|))."""
    return torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-x.abs()))


class Adam:
    """tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)."""

    def __init__(self, lr, beta1, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m, self.v = {}, {}

    def step(self, values, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for name, g in grads.items():
            g = np.asarray(g, dtype=np.float64)
            m = self.m.get(name, np.zeros_like(g))
            v = self.v.get(name, np.zeros_like(g))
            m = self.b1 * m + (1 - self.b1) * g
            v = self.b2 * v + (1 - self.b2) * g * g
            self.m[name], self.v[name] = m, v
            values[name] = (np.asarray(values[name], np.float64) - lr_t * m / (np.sqrt(v) + self.eps)).astype(values[name].dtype)


def forward_losses(values, x_rows, y_rows, cfg, hp, dtype=torch.float64):
    """Whole training graph (GAN/multipassGAN-4x.py:726-768) with train=True. Returns (ctx, dict of losses)."""
    store = og.VarStore(seed=hp.get("seed", 1))
    store.values = values  # shared: variables created on first use persist in the caller's dict
    ctx = TrainContext(store, dtype)
    x = torch.as_tensor(np.asarray(x_rows)).to(dtype)
    y = torch.as_tensor(np.asarray(y_rows)).to(dtype)
    bn = cfg.batch_norm
    gen_part, _ = on.gen_resnet(x, ctx, cfg, train=True)
    disc, dy1, dy2, dy3, dy4 = on.disc_binclass(x, y, ctx, cfg, train=True, use_batch_norm=bn)
    gen, gy1, gy2, gy3, gy4 = on.disc_binclass(x, gen_part, ctx, cfg, train=True, use_batch_norm=bn)
    k2 = hp.get("k2_l", (1.0, 1.0, 1.0, 1.0))
    L = {}
    L["disc_loss_disc"] = bce_logits(disc, 1.0).mean()
    L["disc_loss_gen"] = bce_logits(gen, 0.0).mean()
    L["disc_loss"] = L["disc_loss_disc"] * hp.get("weight_dld", 1.0) + L["disc_loss_gen"]
    L["disc_loss_layer"] = sum(k * 0.5 * ((a - b) ** 2).sum() for k, a, b in
                               zip(k2, (dy1, dy2, dy3, dy4), (gy1, gy2, gy3, gy4)))  # tf.nn.l2_loss = sum(t^2)/2
    L["gen_loss"] = bce_logits(gen, 1.0).mean()
    L["gen_l1_loss"] = (y - gen_part).abs().mean()
    L["gen_loss_complete"] = L["gen_loss"] + L["gen_l1_loss"] * hp["kk"] + L["disc_loss_layer"] * hp["kk2"]
    L["gen_part"] = gen_part
    return ctx, L


def _apply(values, ctx, loss, select, opt):
    names = [n for n, t in ctx.leaves.items() if t.requires_grad and select(n)]
    grads = torch.autograd.grad(loss, [ctx.leaves[n] for n in names], allow_unused=True)
    g = {n: (gr.detach().numpy() if gr is not None else np.zeros(ctx.leaves[n].shape)) for n, gr in zip(names, grads)}
    opt.step(values, g)
    for n, t in ctx.bn_updates.items():  # every UPDATE_OP collected so far runs with both optimizers (:776-787)
        values[n] = t.detach().numpy().astype(values[n].dtype)
    return g


def is_g_var(name):
    return "g_" in name  # GAN/multipassGAN-4x.py:784


def is_d_var(name):
    return "d_" in name  # :787


def train_iteration(values, batches_d, batches_g, cfg, hp, opt_d, opt_g, dtype=torch.float64):
    """One loop body: len(batches_d) discriminator steps then len(batches_g) generator steps.
    `values` (name -> array) is updated in place. Returns the scalar losses and the gradients of the last steps."""
    out = {}
    for (xb, yb) in batches_d:
        ctx, L = forward_losses(values, xb, yb, cfg, hp, dtype)
        out["grads_d"] = _apply(values, ctx, L["disc_loss"], is_d_var, opt_d)
        out["disc_loss"] = float(L["disc_loss"])
    for (xb, yb) in batches_g:
        ctx, L = forward_losses(values, xb, yb, cfg, hp, dtype)
        out["grads_g"] = _apply(values, ctx, L["gen_loss_complete"], is_g_var, opt_g)
        for k in ("gen_loss", "gen_l1_loss", "disc_loss_layer", "gen_loss_complete"):
            out[k] = float(L[k])
    return out
