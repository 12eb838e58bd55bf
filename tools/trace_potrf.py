"""Developer tool: launch timeline of one plan evaluation (engine.trace) - where the look-ahead schedule of the
factorisation spends its time.  usage: python tools/trace_potrf.py [n] [stages] [out.txt]"""
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessfundamentals_b200 import engine as eng  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    stages = int(sys.argv[2]) if len(sys.argv) > 2 else (eng.STAGE_ASSEMBLE | eng.STAGE_POTRF | eng.STAGE_NLL)
    out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/trace_potrf_%d.txt" % n
    rng = np.random.default_rng(0)
    x = np.sort(rng.uniform(0, 1, n))[:, None]
    y = np.sin(20 * x[:, 0]) + 0.1 * rng.standard_normal(n)
    prog = eng.DeviceProgram.get(("SE",), 1, False, 1)
    plan = eng.Plan([prog], [n], want_grad=True)
    plan.set_data(0, torch.from_numpy(x), torch.from_numpy(y))
    plan.set_hp(0, [0.3], 1e-2)
    for _ in range(2):
        plan.eval(stages)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.eval(stages); e1.record(); torch.cuda.synchronize()
    print("untraced eval: %.3f ms" % e0.elapsed_time(e1))
    with eng.trace() as t:
        plan.eval(stages)
    spans = t.spans
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        for sp in spans:
            f.write("%s %d %d %d %.1f %.1f\n" % sp)
    t_end = max(s[5] for s in spans)
    print("traced eval: %.3f ms, %d launches -> %s" % (t_end / 1e3, len(spans), out))
    agg = collections.OrderedDict()
    for tag, a, b, st, t0, t1 in spans:
        agg.setdefault((tag, st), []).append(t1 - t0)
    print("%-10s %3s %5s %10s %9s %9s %9s" % ("tag", "st", "n", "sum_ms", "mean_us", "min_us", "max_us"))
    for (tag, st), v in agg.items():
        print("%-10s %3d %5d %10.3f %9.1f %9.1f %9.1f" % (tag, st, len(v), sum(v) / 1e3, np.mean(v), min(v), max(v)))
    # chain: start of consecutive diagonal blocks
    diag = [(a, t0, t1) for tag, a, b, st, t0, t1 in spans if tag == "diag"]
    print("diag k: start_us span_us gap_to_next_us")
    for i, (k, t0, t1) in enumerate(diag):
        nxt = diag[i + 1][1] if i + 1 < len(diag) else t_end
        if i % 4 == 0 or i + 1 == len(diag):
            print("  %3d %9.1f %7.1f %8.1f" % (k, t0, t1 - t0, nxt - t0))


if __name__ == "__main__":
    main()
