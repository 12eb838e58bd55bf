"""CPU tests of the kernel-program compiler and interpreter (csrc/program.cuh compiled for the host by the test
harness) against the oracle: values, reverse-mode derivatives and the trace-gradient formula."""
import numpy as np
import pytest
import torch

from gaussianprocessfundamentals_b200.program import compile_spec
from oracle import gp_oracle as orc
from tests.harness.build_harness import load

COMPOSITE = ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)])
CASES = [
    (("SE",), [0.1], 1, False, 1),
    (("PER",), [0.5, 0.3], 1, False, 1),
    (("LIN",), [[0.01]], 1, False, 1),
    (COMPOSITE, [0.1, 0.1, 0.1, [0.01]], 1, False, 1),
    (COMPOSITE, [0.1, 0.7, 0.1, 0.2, 1.5, [0.01], 0.3], 1, True, 1),
    (("ADD", [("MAT32",), ("MAT52",), ("WN",)]), [0.2, -0.3], 1, False, 1),
    (("ADD", [("MUL", [("SE",), ("PER",), ("LIN",)]), ("MUL", [("SE",), ("ADD", [("LIN",), ("PER",)])])]),
     [0.2, 0.4, 0.25, [0.3], 0.15, [-0.2], 0.6, 0.35], 1, False, 1),
    (("CP", [("SE",), ("PER",), ("ADD", [("SE",), ("LIN",)])]), [0.31, 0.67, 0.1, 0.2, 0.15, 0.08, [0.5]], 1, False, 1),
    (("CP", [("SE",), ("PER",), ("ADD", [("SE",), ("LIN",)])]), [0.31, 0.67, 0.1, 0.2, 0.15, 0.08, [0.5]], 1, False, 2),
    (("CP", [("SE",), ("PER",)]), [0.5, 0.1, 0.2, 0.15], 1, False, 0),
    (("ADD", [("SE_ARD",), ("LIN",), ("SE",)]), [[0.3, 0.5, 0.8], [0.1, -0.2, 0.3], 0.4], 3, False, 1),
    (("MUL", [("ADD", [("CP", [("SE",), ("LIN",)]), ("PER",)]), ("SE",)]), [0.4, 0.2, [0.1], 0.3, 0.2, 0.5], 1, False, 1),
]


def _flat(hp):
    out = []
    for h in hp:
        out.extend(np.asarray(h, dtype=np.float64).reshape(-1).tolist())
    return np.asarray(out, dtype=np.float64)


def _host_matrix(lib, prog, cp_mode, x, x2, hpf, W=None):
    n, m = x.shape[0], x2.shape[0]
    K = np.zeros((n, m))
    g = np.zeros(max(prog.n_hp, 1))
    code = np.ascontiguousarray(prog.code)
    x = np.ascontiguousarray(x); x2 = np.ascontiguousarray(x2)
    Wc = None if W is None else np.ascontiguousarray(W)
    rc = lib.h_matrix(code.ctypes.data, prog.n_ops, prog.dim, cp_mode, x.ctypes.data, n, x2.ctypes.data, m,
                      hpf.ctypes.data, K.ctypes.data, None if W is None else Wc.ctypes.data,
                      None if W is None else g.ctypes.data)
    assert rc == 0
    return K, g[:prog.n_hp]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_values_match_oracle(case):
    lib = load()
    tree, hp, d, scaled, cp_mode = CASES[case]
    rng = np.random.default_rng(case)
    x = np.sort(rng.uniform(0, 1, (37, d)), axis=0)
    x2 = np.sort(rng.uniform(-0.2, 1.2, (23, d)), axis=0)
    prog = compile_spec(tree, d, scaled)
    assert prog.n_hp == _flat(hp).size
    K, _ = _host_matrix(lib, prog, cp_mode, x, x2, _flat(hp))
    hp_t = [torch.tensor(np.asarray(h, dtype=np.float64)) for h in hp]
    Kref = orc.kernel_matrix(tree, hp_t, torch.tensor(x), torch.tensor(x2), scaled, cp_mode,
                             reference_distance=False).numpy()
    assert np.max(np.abs(K - Kref)) <= 4e-15 * max(1.0, np.max(np.abs(Kref)))
    if d == 1:  # the reference's sqrt(a^2 - 2ab + b^2) formula agrees where it is finite
        Kref2 = orc.kernel_matrix(tree, hp_t, torch.tensor(x), torch.tensor(x2), scaled, cp_mode,
                                  reference_distance=True).numpy()
        ok = np.isfinite(Kref2)
        assert ok.mean() > 0.9
        assert np.max(np.abs(K[ok] - Kref2[ok])) <= 1e-11 * max(1.0, np.max(np.abs(Kref)))


@pytest.mark.parametrize("case", range(len(CASES)))
def test_trace_gradient_matches_autodiff(case):
    """dNLL/dtheta = sum_ij 1/2 (Kinv - alpha alpha^T)_ij dK_ij/dtheta with the interpreter's derivatives equals
    autodiff through the oracle's Cholesky (what the reference's GradientTape computes)."""
    lib = load()
    tree, hp, d, scaled, cp_mode = CASES[case]
    rng = np.random.default_rng(100 + case)
    n = 60
    x = np.sort(rng.uniform(0, 1, (n, d)), axis=0)
    y = np.sin(6 * x[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    noise = 0.05
    prog = compile_spec(tree, d, scaled)
    hpf = _flat(hp)
    K, _ = _host_matrix(lib, prog, cp_mode, x, x, hpf)
    Kn = K + noise * np.eye(n)
    Kinv = np.linalg.inv(Kn)
    alpha = Kinv @ y
    G = 0.5 * (Kinv - alpha @ alpha.T)
    _, g = _host_matrix(lib, prog, cp_mode, x, x, hpf, W=G)
    ref, gref, gnoise = orc.nll_and_grad(tree, hp, noise, x, y, scaled, cp_mode, reference_distance=False)
    gflat = np.concatenate([np.asarray(v).reshape(-1) for v in gref]) if gref else np.zeros(0)
    assert g.shape == gflat.shape
    assert np.max(np.abs(g - gflat)) <= 1e-8 * max(1e-6, np.max(np.abs(gflat))), (g, gflat)
    assert abs(np.trace(G) - gnoise) <= 1e-8 * abs(gnoise)


def test_compile_layout_matches_reference_order():
    # SURVEY App. C KAT: MUL[ADD[SE,PER],LIN] consumes [l_se, l_per, p_per, c]
    prog = compile_spec(COMPOSITE, 1, False)
    assert prog.entries == [(0, 1), (1, 1), (2, 1), (3, 1)]
    assert prog.code[:, 0].tolist() == [1, 2, 16, 3, 17]
    prog = compile_spec(("MUL", [("LIN",), ("ADD", [("PER",), ("SE",)])]), 2, True)
    assert prog.entries == [(0, 2), (2, 1), (3, 1), (4, 1), (5, 1), (6, 1), (7, 1)]
    # change points first (Operators.py:451-453)
    prog = compile_spec(("CP", [("SE",), ("PER",), ("LIN",)]), 1, False)
    assert prog.entries[:2] == [(0, 1), (1, 1)] and prog.n_hp == 6
    assert prog.code[:, 0].tolist() == [1, 18, 2, 18, 16, 3, 18, 16]
    with pytest.raises(ValueError):
        compile_spec(("CP", [("SE",), ("PER",)]), 2, False)
