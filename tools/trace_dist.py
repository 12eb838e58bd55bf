"""Developer tool: per-rank launch timeline (engine.trace) of the distributed LML + gradient evaluation.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
           tools/trace_dist.py [n] [out_prefix]
Writes <out_prefix>_rank<r>.txt (one line per launch) and prints a per-tag summary for ranks 0 and 1."""
import collections
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gaussianprocessfundamentals_b200 import engine as eng  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    prefix = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/trace_dist_%d" % n
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    w = bench.WORKLOADS["m32kd"]
    d = w["d"]
    x, y, ell = bench.make_c5(n, d)
    prog = eng.DeviceProgram.get(w["tree"], d, False, 1)
    P, Q = eng.ProcessGrid.default_shape(world)
    plan = eng.Plan([prog], [n], want_grad=True, grid=eng.ProcessGrid(P, Q))
    plan.set_data(0, torch.tensor(x), torch.tensor(y))
    plan.set_hp(0, ell, 1e-2)
    bits = [eng.STAGE_ASSEMBLE, eng.STAGE_POTRF, eng.STAGE_NLL, eng.STAGE_TRTRI, eng.STAGE_LAUUM, eng.STAGE_GRAD]
    for _ in range(2):
        for b in bits:
            plan.eval(b)
    torch.cuda.synchronize()
    dist.barrier()
    marks = {}
    with eng.trace() as t:
        for b in bits:
            plan.eval(b)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks[b] = ev
    spans = t.spans
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    with open("%s_rank%d.txt" % (prefix, rank), "w") as f:
        for sp in spans:
            f.write("%s %d %d %d %.1f %.1f\n" % sp)
    dist.barrier()
    if rank < 2:
        agg = collections.OrderedDict()
        for tag, a, b, st, t0, t1 in spans:
            agg.setdefault((tag, st), []).append(t1 - t0)
        lines = ["rank %d of %d, n = %d: %d launches, end %.1f ms" % (rank, world, n, len(spans), max(s[5] for s in spans) / 1e3),
                 "%-10s %3s %5s %10s %9s %9s %9s" % ("tag", "st", "n", "sum_ms", "mean_us", "min_us", "max_us")]
        for (tag, st), v in agg.items():
            lines.append("%-10s %3d %5d %10.3f %9.1f %9.1f %9.1f" % (tag, st, len(v), sum(v) / 1e3, np.mean(v), min(v), max(v)))
        print("\n".join(lines), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
