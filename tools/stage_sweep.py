"""Developer tool: per-stage CUDA-event times of an SE plan under the current GPB_* environment.
usage: python tools/stage_sweep.py B:n [B:n ...]   e.g.  GPB_POTRF_KB=4 python tools/stage_sweep.py 1:8192 256:2048"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessfundamentals_b200 import engine as eng  # noqa: E402

STAGES = [("assemble", eng.STAGE_ASSEMBLE), ("potrf", eng.STAGE_POTRF), ("nll", eng.STAGE_NLL),
          ("trtri", eng.STAGE_TRTRI), ("lauum", eng.STAGE_LAUUM), ("grad", eng.STAGE_GRAD)]


def run(B, n, reps):
    rng = np.random.default_rng(0)
    prog = eng.DeviceProgram.get(("SE",), 1, False, 1)
    plan = eng.Plan([prog] * B, [n] * B, want_grad=True)
    for b in range(B):
        x = np.sort(rng.uniform(0, 1, n))[:, None]
        y = np.sin(20 * x[:, 0]) + 0.1 * rng.standard_normal(n)
        plan.set_data(b, torch.from_numpy(x), torch.from_numpy(y))
        plan.set_hp(b, [0.3], 1e-2)
    acc = {k: [] for k, _ in STAGES}
    whole = []
    for it in range(reps + 1):
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(len(STAGES) + 1)]
        marks[0].record()
        for i, (_, bit) in enumerate(STAGES):
            plan.eval(bit)
            marks[i + 1].record()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.eval(eng.STAGES_LML_GRAD); e1.record(); torch.cuda.synchronize()
        if it == 0:
            continue
        for i, (k, _) in enumerate(STAGES):
            acc[k].append(marks[i].elapsed_time(marks[i + 1]))
        whole.append(e0.elapsed_time(e1))
    nll, _, info = plan.results()
    rec = {"B": B, "n": n, "env": {k: v for k, v in os.environ.items() if k.startswith("GPB_")},
           "stages_ms": {k: round(float(np.median(v)), 4) for k, v in acc.items()},
           "whole_ms": round(float(np.median(whole)), 4), "nll0": float(nll[0]), "info_max": int(max(info))}
    fl = B * n ** 3 / 3.0
    rec["potrf_tflops"] = round(fl / rec["stages_ms"]["potrf"] / 1e9, 2)
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    for a in sys.argv[1:]:
        B, n = a.split(":")
        run(int(B), int(n), 3 if int(n) >= 16384 else 5)
