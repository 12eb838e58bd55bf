import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from gaussianprocessfundamentals_b200 import engine as eng
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
trees, hps, ns, xs, ys = bench.build_workload("c4", 0, 1)
trees, hps, ns, xs, ys = trees[:nb], hps[:nb], ns[:nb], xs[:nb], ys[:nb]
progs = eng.DeviceProgram.get_many(trees, 1, False, 1)
plan = eng.Plan(progs, ns, want_grad=True)
for b in range(len(ns)):
    plan.set_data(b, torch.tensor(xs[b]), torch.tensor(ys[b])); plan.set_hp(b, hps[b], 1e-2)
res = []
for rep in range(3):
    plan.eval(eng.STAGES_LML_GRAD); torch.cuda.synchronize()
    nll, grads, info = plan.results()
    res.append((nll.copy(), info.copy()))
    bad = np.nonzero(info)[0]
    print("rep", rep, "bad GPs", len(bad), bad[:10], info[bad[:10]], "nll sum", np.nansum(nll))
print("KB", os.environ.get("GPB_POTRF_KB"), "DIAG", os.environ.get("GPB_DIAG"), "max |nll diff| between reps", np.nanmax(np.abs(res[0][0]-res[1][0])))
