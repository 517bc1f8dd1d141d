"""Per-layer comparison of the 16-bit tensor-core path against the fp32 CUDA-core path on a golden net case.
python tools/debug_layers.py out_net1_u8 [fp16]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import engine, graph as G, networks as N, pipeline as P, weights as W

tag = sys.argv[1] if len(sys.argv) > 1 else "out_net1_u8"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
nets = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "nets.npz"))
cfg = json.loads(str(nets[tag + "_cfg"]))
s = cfg["spec"]
spec = P.NetSpec(use_res_net=s["use_res_net"], add_adj_idcs=s["add_adj_idcs"], startFms=s["startFms"],
                 maxFms=s["maxFms"], filterSize=s["filterSize"], first_nn_arch=s["first_nn_arch"])
G.reset_default_graph()
out = P.build_out_graph(s["idx"], spec, N.config_out(cfg["L"], upRes=cfg["u"]))
w = W.randomize_bn_stats(W.init_graph_variables(G.get_default_graph(), cfg["seed"]), cfg["seed"])
x = nets[tag + "_x"]
feeds = {"x": torch.from_numpy(x).cuda()}
if tag + "_y" in nets:
    feeds["y"] = torch.from_numpy(nets[tag + "_y"]).cuda()
a = engine.CompiledNet(out, w, x.shape[0], precision="fp32")
b = engine.CompiledNet(out, w, x.shape[0], precision=prec)
ya = a.run(feeds).float().cpu().numpy()
yb = b.run(feeds).float().cpu().numpy()
torch.cuda.synchronize()
print("final rel", np.linalg.norm(ya - yb) / np.linalg.norm(ya))
for i in sorted(a.step_bufs):
    ba, bb = a.step_bufs[i], b.step_bufs[i]
    ta = ba.tensor.float().cpu().numpy()[..., :ba.c]
    tb = bb.tensor.float().cpu().numpy()[..., :bb.c]
    rel = np.linalg.norm(ta - tb) / (np.linalg.norm(ta) + 1e-30)
    print("%2d rel=%.3e max=%.3e  %s | %s" % (i, rel, np.abs(ta - tb).max(), a.steps[i][0], b.steps[i][0]))
