"""One LML+gradient sweep of B candidate kernels x n points (the C3 workload at reduced B) for ncu.  usage: c3_eval.py [B] [n]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gaussianprocessfundamentals_b200 import engine as eng  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
trees, hps = bench.candidate_trees(256)
trees, hps = trees[:B], hps[:B]
x, _ = bench.make_xy(n, 2)
progs = [eng.DeviceProgram.get(t, 1, False, 1) for t in trees]
plan = eng.Plan(progs, [n] * B, want_grad=True)
for b in range(B):
    plan.set_data(b, torch.tensor(x), torch.tensor(bench.make_xy(n, 1000 + b)[1]))
    plan.set_hp(b, hps[b], 1e-2)
for _ in range(2):
    plan.eval(eng.STAGES_LML_GRAD)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
ev[0].record()
for i, st in enumerate([eng.STAGE_ASSEMBLE, eng.STAGE_POTRF, eng.STAGE_NLL, eng.STAGE_INVERSE, eng.STAGE_GRAD]):
    plan.eval(st); ev[i + 1].record()
torch.cuda.synchronize()
print("B", B, "n", n, "ops/tree", np.mean([p.compiled.n_ops for p in progs]), "hp/tree", np.mean([p.n_hp for p in progs]),
      "ms:", ["%.2f" % ev[i].elapsed_time(ev[i + 1]) for i in range(5)])
nll, grads, info = plan.results()
print("nll0", nll[0], "info", int(info.max()))
