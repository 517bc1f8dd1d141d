"""Oracle networks wrapped as `sess.run`-style callables on flat rows (test helper)."""
import numpy as np
import torch

from oracle import gan as og
from oracle import networks as on


def oracle_gen_resnet(weights, L, mode=2, dtype=torch.float64, upRes=4, batch_norm=True):
    cfg = on.make_cfg_4x(L, upRes=upRes, upsampling_mode=mode, batch_norm=batch_norm)

    def run(rows):
        ctx = og.Context(og.VarStore(values=weights), dtype)
        y, _ = on.gen_resnet(torch.as_tensor(np.asarray(rows)).to(dtype), ctx, cfg)
        return y.numpy()

    return run


def oracle_growing_gen(weights, idx, spec, L, upRes=8, dtype=torch.float64, **cfg_kw):
    cfg = on.make_cfg_out(L, upRes=upRes, **cfg_kw)
    cu = on.log2i(upRes)

    def run(x_rows, y_rows=None):
        ctx = og.Context(og.VarStore(values=weights), dtype)
        x = torch.as_tensor(np.asarray(x_rows)).to(dtype)
        with ctx.variable_scope("gen_%d" % idx):
            if idx == 1:
                y, _ = on.growing_gen(x, ctx, cfg, currentUpres=cu, output=True, firstGen=True,
                                      filterSize=spec.filterSize, startFms=spec.startFms, maxFms=spec.maxFms,
                                      add_adj_idcs=spec.add_adj_idcs, first_nn_arch=spec.first_nn_arch,
                                      use_res_net=spec.use_res_net)
            else:
                xin = on.sampler_input_2(x, torch.as_tensor(np.asarray(y_rows)).to(dtype), cfg)
                y, _ = on.growing_gen(xin, ctx, cfg, currentUpres=cu, output=True, firstGen=False,
                                      filterSize=spec.filterSize, startFms=spec.startFms, maxFms=spec.maxFms,
                                      add_adj_idcs=False, first_nn_arch=False, use_res_net=spec.use_res_net)
        return y.numpy()

    return run


def err_stats(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    d = got - ref
    return dict(max_abs=float(np.abs(d).max()), ref_max=float(np.abs(ref).max()),
                rel_l2=float(np.linalg.norm(d) / (np.linalg.norm(ref) + 1e-300)))
