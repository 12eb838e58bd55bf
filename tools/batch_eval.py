"""One LML+gradient evaluation of a batched bench workload (for ncu launch lists).  usage: batch_eval.py c3|c4 [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gaussianprocessfundamentals_b200 import engine as eng  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "c4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
trees, hps, ns, xs, ys = bench.build_workload(key, 0, 1)
progs = eng.DeviceProgram.get_many(trees, 1, False, 1)
plan = eng.Plan(progs, ns, want_grad=True)
for b in range(len(ns)):
    plan.set_data(b, torch.tensor(xs[b]), torch.tensor(ys[b]))
    plan.set_hp(b, hps[b], 1e-2)
for _ in range(reps):
    plan.eval(eng.STAGES_LML_GRAD)
torch.cuda.synchronize()
nll, grads, info = plan.results()
print(key, "GPs", len(ns), "nll0", nll[0], "info", int(info.max()), "launches/eval", eng.launch_count() // reps)
