"""Phase clocks of the diagonal-block kernel (developer build: GPB_NVCC_EXTRA=-DGPB_DIAG_CLOCKS=1 python csrc/build.py --force)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessfundamentals_b200 import _lib, engine as eng  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
x = np.linspace(0, 1, n)[:, None]
y = np.sin(9 * x)
prog = eng.DeviceProgram.get(("SE",), 1, False, 1)
plan = eng.Plan([prog], [n], want_grad=True)
plan.set_data(0, torch.tensor(x), torch.tensor(y))
plan.set_hp(0, np.array([0.1]), 1e-2)
for _ in range(3):
    plan.eval(eng.STAGES_LML)
torch.cuda.synchronize()
out = (ctypes.c_longlong * 8)()
rc = _lib.load().gpb_debug_diag_clocks(out)
names = ["load", "panel phases", "update phases", "logdet/identity rows", "inverse level 0", "inverse levels 8..64", "store W"]
print("rc", rc, "total", sum(out[:7]))
for nm, v in zip(names, out):
    print("%-24s %8d cycles" % (nm, v))
