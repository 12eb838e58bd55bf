"""One LML+gradient evaluation of the n-point composite GP (for ncu launch lists).  usage: one_eval.py N [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessfundamentals_b200 import engine as eng  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tree = ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)])
rng = np.random.default_rng(1)
x = np.linspace(0, 1, n)[:, None]
y = x * np.sin(40 * x) + 0.1 * rng.standard_normal((n, 1))
prog = eng.DeviceProgram.get(tree, 1, False, 1)
plan = eng.Plan([prog], [n], want_grad=True)
plan.set_data(0, torch.tensor(x), torch.tensor(y))
plan.set_hp(0, np.array([0.1, 0.1, 0.1, 0.01]), 1e-2)
for _ in range(reps):
    plan.eval(eng.STAGES_LML_GRAD)
torch.cuda.synchronize()
nll, grads, info = plan.results()
print("n", n, "nll", nll[0], "grad", grads[0], "info", info[0], "launches/eval", eng.launch_count() // reps)
