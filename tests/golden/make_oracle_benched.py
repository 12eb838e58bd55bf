"""Generates tests/golden/oracle_benched_sizes.json: the CPU oracle (oracle/gp_oracle.py) evaluated on the exact inputs of
the parity tests at the BENCHED sizes (tests/test_gpu_parity_sizes.py), in the build container.

Why a fixture and not only the live oracle: at n = 8192 the covariance matrix of C2 has a condition number of ~1e6, so
the CPU Cholesky of the oracle is itself only good to cond * eps ~ 1e-10 in the likelihood - on one GPU box of the pool
the live oracle (another CPU / BLAS code path) came out 1.9e-10 away from the value computed here, while the device
result is bitwise the same on every box and agrees with this container's oracle to 1e-13 for both distance formulas.
The tests therefore compare the device with these committed values at the north-star tolerance (1e-10 / 1e-8) and
require the live oracle of the box to agree with them to 1e-8.

usage: python tests/golden/make_oracle_benched.py   (about 5 minutes on 8 cores)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as orc  # noqa: E402
from tests import benched_cases as bc  # noqa: E402


def flat(gref, gnoise):
    return [float(v) for v in np.concatenate([np.asarray(g).reshape(-1) for g in gref] + [[gnoise]])]


def both(tree, hp, x, y):
    out = {}
    for key, rd in (("reference_distance", True), ("direct_distance", False)):
        nll, gref, gnoise = orc.nll_and_grad(tree, hp, 1e-2, x, y, reference_distance=rd)
        out[key] = {"nll": float(nll), "grad": flat(gref, gnoise)}
    return out


def main():
    gold = {"c2": {}, "c3": {}, "c4": {}, "m16k": {}}
    for n in bc.C2_SIZES:
        x, y = bc.c2_inputs(n)
        gold["c2"][str(n)] = both(bc.COMPOSITE, bc.C2_HP, x, y)
        print("c2", n, gold["c2"][str(n)]["reference_distance"]["nll"], gold["c2"][str(n)]["direct_distance"]["nll"], flush=True)
    order, trees, hps, x, ys = bc.c3_inputs()
    for k, b in enumerate(order):
        gold["c3"][str(b)] = both(trees[b], bc.c3_hp_struct(trees[b], hps[b]), x, ys[k])
        print("c3", b, gold["c3"][str(b)]["reference_distance"]["nll"], flush=True)
    x, y, ls, nb, n = bc.c4_inputs()
    blocks = []
    for j in range(nb):
        nll, gref, gn = orc.nll_and_grad(("SE",), [ls[j]], 1e-2, x[j * n:(j + 1) * n], y[j * n:(j + 1) * n],
                                         reference_distance=True)
        blocks.append({"nll": float(nll), "grad_ls": float(np.asarray(gref[0]).reshape(-1)[0]), "grad_noise": float(gn)})
    gold["c4"]["blocks"] = blocks
    ref, alpha = bc.m16k_lapack()
    gold["m16k"] = {"nll": float(ref), "alpha_head": [float(v) for v in alpha.reshape(-1)[:64]],
                    "alpha_absmax": float(np.max(np.abs(alpha))), "alpha_sum": float(np.sum(alpha))}
    gold["made_with"] = {"torch": torch.__version__, "numpy": np.__version__, "threads": torch.get_num_threads()}
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_benched_sizes.json")
    with open(out, "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
