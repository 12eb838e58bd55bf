"""Inputs of the parity tests at the benched sizes, shared by tests/test_gpu_parity_sizes.py and the script that freezes
the oracle's values for them (tests/golden/make_oracle_benched.py) - one construction, so the two cannot drift."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

COMPOSITE = ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)])
C2_HP = [0.1, 0.1, 0.1, [0.01]]
C2_FLAT = np.array([0.1, 0.1, 0.1, 0.01])
C2_SIZES = (8192, 6145, 8191)


def c2_inputs(n):
    import bench
    return bench.make_xy(n, 1)


def c3_inputs(n=2048, B=16):
    """the 16 largest programs of bench.py's C3 draw (seed 2): (order, trees, hps, x, ys)"""
    import bench
    trees, hps = bench.candidate_trees(256)
    order = sorted(range(256), key=lambda b: -len(hps[b]))[:B]
    x, _ = bench.make_xy(n, 2)
    ys = [bench.make_xy(n, 1000 + b)[1] for b in order]
    return order, trees, hps, x, ys


def c3_hp_struct(tree, flat):
    from gaussianprocessfundamentals_b200.program import compile_spec
    return [flat[o] if s == 1 else flat[o:o + s] for o, s in compile_spec(tree, 1, False).entries]


def c4_inputs(nb=32, n=1024):
    N = nb * n
    x = (np.arange(N) / N)[:, None]                     # exact in binary: the strict-< rule cuts exactly n points each
    rng = np.random.default_rng(3)
    y = np.sin((50 + (np.arange(N) // n) % 7)[:, None] * 40 * x) + 0.1 * rng.standard_normal((N, 1))
    ls = rng.uniform(0.2, 1.0, nb) / nb
    return x, y, ls, nb, n


def m16k_lapack(n=16384):
    """NLL and alpha of C2's kernel at n = 16384 by LAPACK (scipy) on the oracle's covariance matrix"""
    import torch
    from scipy.linalg import cho_factor, cho_solve
    from oracle import gp_oracle as orc
    x, y = c2_inputs(n)
    hp_t = [torch.tensor(0.1, dtype=torch.float64), torch.tensor(0.1, dtype=torch.float64),
            torch.tensor(0.1, dtype=torch.float64), torch.tensor([0.01], dtype=torch.float64)]
    with torch.no_grad():
        K = orc.kernel_matrix(COMPOSITE, hp_t, torch.tensor(x), torch.tensor(x), reference_distance=True).numpy()
    K[np.diag_indices(n)] += 1e-2
    c, low = cho_factor(K, lower=True, overwrite_a=True, check_finite=False)
    alpha = cho_solve((c, low), y, check_finite=False)
    ref = 0.5 * float(y.reshape(-1) @ alpha.reshape(-1)) + float(np.sum(np.log(np.diag(c)))) + 0.5 * n * np.log(2 * np.pi)
    return ref, alpha
