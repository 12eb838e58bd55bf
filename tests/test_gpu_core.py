"""GPU parity tests of the CUDA path (through the C-ABI) against the CPU oracle.  Tolerances (north_star):
relative <= 1e-10 on the log-likelihood, <= 1e-8 on gradients; matrices element-wise to a few ulp."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LL_RTOL = 1e-10
GRAD_RTOL = 1e-8


def _eng():
    from gaussianprocessfundamentals_b200 import engine
    return engine


def _data(n, d=1, seed=0, kind="grid"):
    rng = np.random.default_rng(seed)
    if d == 1 and kind == "grid":
        x = np.linspace(0.0, 1.0, n)[:, None]
    else:
        x = np.sort(rng.uniform(0, 1, size=(n, d)), axis=0)
    y = np.sin(12 * x[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    return x, y


COMPOSITE = ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)])
TREES = {
    "se": (("SE",), [0.1]),
    "per": (("PER",), [0.5, 0.3]),
    "lin": (("LIN",), [[0.01]]),
    "composite": (COMPOSITE, [0.1, 0.1, 0.1, [0.01]]),
    "mat": (("ADD", [("MAT32",), ("MAT52",), ("WN",)]), [0.2, 0.3]),
    "deep": (("ADD", [("MUL", [("SE",), ("PER",), ("LIN",)]), ("MUL", [("SE",), ("ADD", [("LIN",), ("PER",)])])]),
             [0.2, 0.4, 0.25, [0.3], 0.15, [-0.2], 0.6, 0.35]),
    "cp": (("CP", [("SE",), ("PER",), ("ADD", [("SE",), ("LIN",)])]), [0.31, 0.67, 0.1, 0.2, 0.15, 0.08, [0.5]]),
}


def _flat(entries_hp):
    out = []
    for h in entries_hp:
        out.extend(np.asarray(h, dtype=np.float64).reshape(-1).tolist())
    return np.asarray(out, dtype=np.float64)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [2, 4])   # 128x128 tiles / 64x128 tiles (csrc/gemm.cuh)
@pytest.mark.parametrize("akm,bkm", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("shape", [(128, 128, 128), (300, 200, 150), (257, 129, 151), (64, 1, 16), (1, 5, 7)])
def test_gemm_dmma(akm, bkm, shape, cfg):
    eng = _eng()
    M, N, K = shape
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K))
    Bm = rng.standard_normal((N, K))
    C0 = rng.standard_normal((M, N))
    ev = lambda v: v + (v & 1)
    # column-major storage: an MN-major operand stores (mn, k) at [mn + k*ld]; a K-major one at [k + mn*ld]
    lda = ev(K + 2) if akm else ev(M + 4)
    ldb = ev(K + 2) if bkm else ev(N + 6)
    ldc = M + 3
    Ah = np.zeros((lda * (M if akm else K),))
    Bh = np.zeros((ldb * (N if bkm else K),))
    Ch = np.zeros((ldc * N,))
    for i in range(M):
        for k in range(K):
            Ah[(k + i * lda) if akm else (i + k * lda)] = A[i, k]
    for j in range(N):
        for k in range(K):
            Bh[(k + j * ldb) if bkm else (j + k * ldb)] = Bm[j, k]
    for i in range(M):
        Ch[i + np.arange(N) * ldc] = C0[i, :]
    dev = torch.device("cuda")
    At, Bt, Ct = (torch.tensor(v, dtype=torch.float64, device=dev) for v in (Ah, Bh, Ch))
    eng.gemm(akm | cfg, bkm, At, lda, Bt, ldb, Ct, ldc, M, N, K, -0.75, 0.5)
    got = Ct.cpu().numpy()
    want = -0.75 * A @ Bm.T + 0.5 * C0
    res = np.array([[got[i + j * ldc] for j in range(N)] for i in range(M)])
    assert np.max(np.abs(res - want)) <= 1e-12 * max(1.0, np.max(np.abs(want))) * K
    # beta == 0 must ignore NaNs in C
    Ct2 = torch.full_like(Ct, float("nan"))
    eng.gemm(akm | cfg, bkm, At, lda, Bt, ldb, Ct2, ldc, M, N, K, 1.0, 0.0)
    got2 = Ct2.cpu().numpy()
    res2 = np.array([[got2[i + j * ldc] for j in range(N)] for i in range(M)])
    assert np.max(np.abs(res2 - A @ Bm.T)) <= 1e-12 * K * max(1.0, np.max(np.abs(A @ Bm.T)))


@pytest.mark.parametrize("name", list(TREES))
@pytest.mark.parametrize("n,m", [(200, 200), (130, 75)])
def test_assemble_matches_oracle(name, n, m):
    eng = _eng()
    tree, hp = TREES[name]
    x, _ = _data(n)
    x2 = np.linspace(-0.2, 1.3, m)[:, None] if m != n else x
    prog = eng.DeviceProgram.get(tree, 1, False, 1)
    dev = torch.device("cuda")
    X = torch.tensor(x, device=dev)
    X2 = None if m == n else torch.tensor(x2, device=dev)
    hpf = torch.tensor(_flat(hp), device=dev)
    noise = torch.tensor([0.01], dtype=torch.float64, device=dev)
    K = eng.assemble(prog, X, X2, hpf, noise if m == n else None).cpu().numpy()
    hp_t = [torch.tensor(np.asarray(h, dtype=np.float64)) for h in hp]
    Kref = orc.kernel_matrix(tree, hp_t, torch.tensor(x), torch.tensor(x2), reference_distance=False).numpy()
    if m == n:
        Kref = Kref + 0.01 * np.eye(n)
    scale = max(1.0, np.max(np.abs(Kref)))
    assert K.shape == Kref.shape
    err = np.abs(K - Kref)
    if np.max(err) > 5e-15 * scale and name == "se":
        # arbitrate with a 40-digit evaluation of the worst entry: the oracle's vectorised CPU exp() is itself only
        # accurate to about an ulp and differs between host CPUs
        import mpmath as mp
        mp.mp.dps = 40
        i, j = np.unravel_index(np.argmax(err), err.shape)
        exact = mp.e ** (-(mp.mpf(float(x[i, 0])) - mp.mpf(float(x2[j, 0]))) ** 2 / (2 * mp.mpf(hp[0]) ** 2))
        exact = float(exact) + (0.01 if (m == n and i == j) else 0.0)
        assert abs(K[i, j] - exact) <= 5e-15 * scale, (i, j, K[i, j], Kref[i, j], exact)
        return
    assert np.max(err) <= 5e-15 * scale, (np.unravel_index(np.argmax(err), err.shape), np.max(err))


def _run_plan(eng, tree, hp, n, noise=1e-2, d=1, scaled=False, cp_mode=1, seed=0, want_grad=True, host=False, x=None,
              y=None):
    if x is None:
        x, y = _data(n, d, seed)
    prog = eng.DeviceProgram.get(tree, d, scaled, cp_mode)
    plan = eng.Plan([prog], [n], want_grad=want_grad)
    hpf = _flat(hp)
    if host:
        nll, grads, info = plan.eval_host([hpf], [noise], [x], [y.reshape(-1)],
                                          stages=eng.STAGES_LML_GRAD if want_grad else eng.STAGES_LML)
    else:
        plan.set_data(0, torch.tensor(x), torch.tensor(y))
        plan.set_hp(0, hpf, noise)
        plan.eval(eng.STAGES_LML_GRAD if want_grad else eng.STAGES_LML)
        nll, grads, info = plan.results()
    return plan, x, y, nll, grads, info


@pytest.mark.parametrize("name", ["se", "composite", "deep", "cp", "mat"])
@pytest.mark.parametrize("n", [64, 128, 129, 300, 1000, 1025])
def test_nll_and_grad_match_oracle(name, n):
    eng = _eng()
    tree, hp = TREES[name]
    plan, x, y, nll, grads, info = _run_plan(eng, tree, hp, n, host=(n % 2 == 0))
    assert info[0] == 0
    ref, gref, gnoise = orc.nll_and_grad(tree, hp, 1e-2, x, y, reference_distance=False)
    assert abs(nll[0] - ref) <= LL_RTOL * abs(ref), (nll[0], ref)
    gflat = np.concatenate([np.asarray(g).reshape(-1) for g in gref] + [[gnoise]])
    got = grads[0]
    for k in range(len(gflat)):
        assert abs(got[k] - gflat[k]) <= GRAD_RTOL * max(abs(gflat[k]), 1e-3 * np.max(np.abs(gflat))), (k, got, gflat)


def test_factor_buffers_match_lapack():
    eng = _eng()
    tree, hp = TREES["composite"]
    n = 700
    plan, x, y, nll, grads, info = _run_plan(eng, tree, hp, n, want_grad=False)
    hp_t = [torch.tensor(np.asarray(h, dtype=np.float64)) for h in hp]
    _, K, L, alpha = orc.nll(tree, hp_t, torch.tensor(1e-2, dtype=torch.float64), torch.tensor(x), torch.tensor(y),
                             reference_distance=False, return_parts=True)
    Lg = torch.tril(plan.lower_matrix(0, eng.BUF_A)).cpu().numpy()
    assert np.max(np.abs(Lg - L.numpy())) <= 1e-10 * np.max(np.abs(L.numpy()))
    # standalone back substitution -> alpha
    plan.eval(eng.STAGE_BACKSOLVE)
    a = plan.buffer(0, eng.BUF_ALPHA).cpu().numpy()
    assert np.max(np.abs(a - alpha.numpy().reshape(-1))) <= 1e-8 * np.max(np.abs(alpha.numpy()))


def test_inverse_buffers():
    eng = _eng()
    tree, hp = TREES["se"]
    n = 900
    plan, x, y, nll, grads, info = _run_plan(eng, tree, hp, n)
    hp_t = [torch.tensor(np.asarray(h, dtype=np.float64)) for h in hp]
    _, K, L, alpha = orc.nll(tree, hp_t, torch.tensor(1e-2, dtype=torch.float64), torch.tensor(x), torch.tensor(y),
                             reference_distance=False, return_parts=True)
    Kn = K.numpy() + 1e-2 * np.eye(n)
    Kinv = np.linalg.inv(Kn)
    got = torch.tril(plan.lower_matrix(0, eng.BUF_KINV)).cpu().numpy()
    assert np.max(np.abs(got - np.tril(Kinv))) <= 1e-9 * np.max(np.abs(Kinv))
    a = plan.buffer(0, eng.BUF_ALPHA).cpu().numpy()
    assert np.max(np.abs(a - alpha.numpy().reshape(-1))) <= 1e-9 * np.max(np.abs(alpha.numpy()))


def test_scaled_and_smooth_changepoint_gradients():
    eng = _eng()
    n = 400
    x, y = _data(n)
    # scaled base kernels (p_scaled_base_kernel): every leaf gains sg
    tree = ("ADD", [("SE",), ("MUL", [("PER",), ("LIN",)])])
    hp = [0.2, 0.7, 0.3, 0.25, 1.3, [0.1], 0.4]
    plan, _, _, nll, grads, info = _run_plan(eng, tree, hp, n, scaled=True, x=x, y=y)
    ref, gref, gnoise = orc.nll_and_grad(tree, hp, 1e-2, x, y, scaled=True, reference_distance=False)
    assert abs(nll[0] - ref) <= LL_RTOL * abs(ref)
    gflat = np.concatenate([np.asarray(g).reshape(-1) for g in gref] + [[gnoise]])
    assert np.max(np.abs(grads[0] - gflat)) <= GRAD_RTOL * np.max(np.abs(gflat))
    # approx-indicator change points carry gradients (Operators.py:379-385)
    tree, hp = TREES["cp"]
    plan, _, _, nll, grads, info = _run_plan(eng, tree, hp, n, cp_mode=2, x=x, y=y)
    ref, gref, gnoise = orc.nll_and_grad(tree, hp, 1e-2, x, y, cp_mode=2, reference_distance=False)
    assert abs(nll[0] - ref) <= LL_RTOL * abs(ref)
    gflat = np.concatenate([np.asarray(g).reshape(-1) for g in gref] + [[gnoise]])
    assert np.max(np.abs(grads[0] - gflat)) <= GRAD_RTOL * np.max(np.abs(gflat))


def test_multidim_and_ard():
    eng = _eng()
    n, d = 500, 3
    rng = np.random.default_rng(4)
    x = rng.uniform(0, 1, size=(n, d))
    y = np.sum(np.sin(3 * x), axis=1, keepdims=True) + 0.1 * rng.standard_normal((n, 1))
    tree = ("ADD", [("SE_ARD",), ("LIN",), ("SE",)])
    hp = [[0.3, 0.5, 0.8], [0.1, -0.2, 0.3], 0.4]
    plan, _, _, nll, grads, info = _run_plan(eng, tree, hp, n, d=d, x=x, y=y)
    ref, gref, gnoise = orc.nll_and_grad(tree, hp, 1e-2, x, y, reference_distance=False)
    assert abs(nll[0] - ref) <= LL_RTOL * abs(ref)
    gflat = np.concatenate([np.asarray(g).reshape(-1) for g in gref] + [[gnoise]])
    assert np.max(np.abs(grads[0] - gflat)) <= GRAD_RTOL * np.max(np.abs(gflat))


def test_ragged_batch_matches_singletons():
    eng = _eng()
    names = ["se", "composite", "deep", "per", "cp"]
    ns = [257, 128, 640, 33, 300]
    progs, xs, ys, hps = [], [], [], []
    for k, (nm, n) in enumerate(zip(names, ns)):
        tree, hp = TREES[nm]
        x, y = _data(n, seed=k)
        progs.append(eng.DeviceProgram.get(tree, 1, False, 1))
        xs.append(x); ys.append(y.reshape(-1)); hps.append(_flat(hp))
    plan = eng.Plan(progs, ns, want_grad=True)
    nll, grads, info = plan.eval_host(hps, [1e-2] * len(ns), xs, ys)
    assert np.all(info == 0)
    for k, (nm, n) in enumerate(zip(names, ns)):
        tree, hp = TREES[nm]
        ref, gref, gnoise = orc.nll_and_grad(tree, hp, 1e-2, xs[k], ys[k].reshape(-1, 1), reference_distance=False)
        assert abs(nll[k] - ref) <= LL_RTOL * abs(ref), (nm, nll[k], ref)
        gflat = np.concatenate([np.asarray(g).reshape(-1) for g in gref] + [[gnoise]])
        assert np.max(np.abs(grads[k] - gflat)) <= GRAD_RTOL * np.max(np.abs(gflat)), nm


def test_not_positive_definite_reports_info():
    eng = _eng()
    n = 256
    x = np.zeros((n, 1))  # identical inputs, zero noise -> singular
    y = np.ones((n, 1))
    plan, _, _, nll, grads, info = _run_plan(eng, ("SE",), [0.1], n, noise=0.0, x=x, y=y)
    assert info[0] > 0 and math.isnan(nll[0])


def test_known_answer_identity_kernel():
    """WN kernel: K = (1 + s2) I  ->  closed-form NLL"""
    eng = _eng()
    n = 513
    x, y = _data(n)
    plan, _, _, nll, grads, info = _run_plan(eng, ("WN",), [], n, noise=0.5, x=x, y=y)
    want = 0.5 * float(y.T @ y) / 1.5 + 0.5 * n * math.log(1.5) + 0.5 * n * math.log(2 * math.pi)
    assert abs(nll[0] - want) <= 1e-12 * abs(want)


def test_large_roundtrip_properties():
    """n = 4096 (beyond the look-ahead threshold): L L^T == K + s2 I and Kinv (K + s2 I) == I on random probes."""
    eng = _eng()
    tree, hp = TREES["composite"]
    n = 4096
    x, y = _data(n)
    prog = eng.DeviceProgram.get(tree, 1, False, 1)
    plan = eng.Plan([prog], [n], want_grad=True)
    plan.set_data(0, torch.tensor(x), torch.tensor(y))
    plan.set_hp(0, _flat(hp), 1e-2)
    dev = torch.device("cuda")
    Kfull = eng.assemble(prog, torch.tensor(x, device=dev), None, torch.tensor(_flat(hp), device=dev),
                         torch.tensor([1e-2], dtype=torch.float64, device=dev))
    plan.eval(eng.STAGES_LML)
    L = torch.tril(plan.lower_matrix(0, eng.BUF_A))
    v = torch.randn(n, 8, dtype=torch.float64, device=dev)
    r1 = L @ (L.t() @ v) - Kfull @ v
    assert float(r1.abs().max()) <= 1e-9 * float((Kfull @ v).abs().max())
    plan.eval(eng.STAGE_INVERSE)
    Kl = torch.tril(plan.lower_matrix(0, eng.BUF_KINV))
    Kinv = Kl + torch.tril(Kl, -1).t()
    r2 = Kinv @ (Kfull @ v) - v
    assert float(r2.abs().max()) <= 1e-6 * float(v.abs().max())


@pytest.mark.parametrize("env_extra", [{"GPB_POTRF_KB": "2"}, {"GPB_POTRF_KB": "3"}, {"GPB_POTRF_KB": "4"},
                                       {"GPB_POTRF_KB": "8"},
                                       {"GPB_POTRF_T2": "384", "GPB_POTRF_T4": "1024", "GPB_POTRF_T8": "1664"}])
def test_wide_outer_panels_on_ragged_batches_subprocess(env_extra):
    """The outer panels of the Cholesky driver are 128 W wide with W chosen per outer step from the remaining rows
    (8 / 4 / 2 / 1 from 20480 / 10240 / 5120 rows; batches: 4 from n = 1024).  GPB_POTRF_KB (read once per process) forces
    one width, GPB_POTRF_T2/_T4/_T8 move the thresholds so that a 2300-point GP walks through all four widths; a
    subprocess evaluates a ragged batch - sizes on both sides of block boundaries, odd and even numbers of blocks - and
    single GPs on the look-ahead path against the oracle."""
    import subprocess
    import sys
    script = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from gaussianprocessfundamentals_b200 import engine as eng
from oracle import gp_oracle as orc
tree = ("ADD", [("SE",), ("PER",)])
hp = np.array([0.2, 0.5, 0.3])
ns = [700, 640, 513, 384, 1000, 129, 257, 256]
prog = eng.DeviceProgram.get(tree, 1, False, 1)
plan = eng.Plan([prog] * len(ns), ns, want_grad=True)
data = []
for b, n in enumerate(ns):
    rng = np.random.default_rng(50 + b)
    x = np.sort(rng.uniform(0, 1, (n, 1)), axis=0); y = np.sin(7 * x) + 0.1 * rng.standard_normal((n, 1))
    data.append((x, y)); plan.set_data(b, torch.tensor(x), torch.tensor(y)); plan.set_hp(b, hp, 1e-2)
plan.eval(eng.STAGES_LML_GRAD); torch.cuda.synchronize()
nll, grads, info = plan.results()
assert int(np.max(info)) == 0
for b, n in enumerate(ns):
    want, g, gn = orc.nll_and_grad(tree, [np.asarray(v) for v in hp], 1e-2, data[b][0], data[b][1], reference_distance=False)
    assert abs(nll[b] - want) <= 1e-10 * abs(want), (n, nll[b], want)
    gw = np.concatenate([np.asarray(t).reshape(-1) for t in g] + [[gn]])
    assert np.max(np.abs(grads[b] - gw)) <= 1e-8 * np.max(np.abs(gw)), (n, grads[b], gw)
# one GP per plan: the look-ahead driver (depth 2, three priority levels) with the same widths
for n in (1153, 2300, 1536):
    rng = np.random.default_rng(90 + n)
    x = np.sort(rng.uniform(0, 1, (n, 1)), axis=0); y = np.sin(7 * x) + 0.1 * rng.standard_normal((n, 1))
    single = eng.Plan([prog], [n], want_grad=True)
    single.set_data(0, torch.tensor(x), torch.tensor(y)); single.set_hp(0, hp, 1e-2)
    single.eval(eng.STAGES_LML_GRAD); torch.cuda.synchronize()
    nll1, g1, info1 = single.results()
    want, g, gn = orc.nll_and_grad(tree, [np.asarray(v) for v in hp], 1e-2, x, y, reference_distance=False)
    gw = np.concatenate([np.asarray(t).reshape(-1) for t in g] + [[gn]])
    assert int(info1[0]) == 0 and abs(nll1[0] - want) <= 1e-10 * abs(want), (n, nll1[0], want)
    assert np.max(np.abs(g1[0] - gw)) <= 1e-8 * np.max(np.abs(gw)), (n, g1[0], gw)
print("ok")
""" % ROOT
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


# ---- kernels specialised per program (csrc/jit.cu) vs the interpreter kernels ---------------------------------------------
def test_programs_run_on_specialised_kernels():
    """every program of this file is compiled to its own assembly / gradient kernels (NVRTC); the interpreter is the
    fallback only"""
    eng = _eng()
    from gaussianprocessfundamentals_b200 import _lib
    assert _lib.load().gpb_jit_available() == 1
    for name, (tree, hp) in TREES.items():
        prog = eng.DeviceProgram.get(tree, 1, False, 1)
        assert prog.specialised, (name, prog.jit_note)


def test_c3_grammar_worst_case_tree_evaluates():
    """depth-3, 3-ary MUL of PER leaves (27 leaves, 54 hyper-parameters, interpreter tape 106 > 64): accepted and correct
    on the specialised kernels"""
    eng = _eng()
    tree = ("MUL", [("MUL", [("MUL", [("PER",)] * 3)] * 3)] * 3)
    rng = np.random.default_rng(11)
    hp = []
    for _ in range(27):
        hp += [float(rng.uniform(2.0, 4.0)), float(rng.uniform(0.3, 0.9))]     # long length scales: the product stays O(1)
    n = 300
    plan, x, y, nll, grads, info = _run_plan(eng, tree, hp, n)
    assert info[0] == 0
    ref, gref, gnoise = orc.nll_and_grad(tree, hp, 1e-2, x, y, reference_distance=False)
    assert abs(nll[0] - ref) <= LL_RTOL * abs(ref)
    gflat = np.concatenate([np.asarray(g).reshape(-1) for g in gref] + [[gnoise]])
    assert np.max(np.abs(grads[0] - gflat)) <= GRAD_RTOL * np.max(np.abs(gflat))


def test_interpreter_kernels_subprocess():
    """GPB_JIT=0 (read once per process) keeps every program on the interpreter kernels - the path of a host without
    NVRTC: the oracle parity tests of this file must hold there too, and the two paths must agree to rounding"""
    import subprocess
    import sys
    script = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from gaussianprocessfundamentals_b200 import engine as eng, _lib
from oracle import gp_oracle as orc
from tests.test_gpu_core import TREES, _run_plan, _flat
assert _lib.load().gpb_jit_available() == 0
out = {}
for name, (tree, hp) in TREES.items():
    for n in (129, 700):
        plan, x, y, nll, grads, info = _run_plan(eng, tree, hp, n, host=(n == 700))
        assert not eng.DeviceProgram.get(tree, 1, False, 1).specialised
        ref, gref, gnoise = orc.nll_and_grad(tree, hp, 1e-2, x, y, reference_distance=False)
        assert info[0] == 0 and abs(nll[0] - ref) <= 1e-10 * abs(ref), (name, n, nll[0], ref)
        gflat = np.concatenate([np.asarray(g).reshape(-1) for g in gref] + [[gnoise]])
        assert np.max(np.abs(grads[0] - gflat)) <= 1e-8 * np.max(np.abs(gflat)), (name, n)
        print("R", name, n, repr(float(nll[0])), " ".join(repr(float(v)) for v in grads[0]))
# a tree beyond the interpreter's tape is refused with a clear message
try:
    eng.DeviceProgram.get(("MUL", [("MUL", [("MUL", [("PER",)] * 3)] * 3)] * 3), 1, False, 1)
    print("NOT REFUSED")
except _lib.GpbError as exc:
    assert "tape" in str(exc)
print("ok")
""" % ROOT
    env = dict(os.environ, GPB_JIT="0")
    out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "ok" in out.stdout and "NOT REFUSED" not in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    eng = _eng()
    for line in out.stdout.splitlines():
        if not line.startswith("R "):
            continue
        _, name, n, nll_i, *g_i = line.split()
        tree, hp = TREES[name]
        plan, x, y, nll, grads, info = _run_plan(eng, tree, hp, int(n), host=(int(n) == 700))
        assert abs(nll[0] - float(nll_i)) <= 1e-12 * abs(nll[0]), (name, n)
        gi = np.array([float(v) for v in g_i])
        assert np.max(np.abs(grads[0] - gi)) <= 1e-10 * np.max(np.abs(gi)), (name, n)


def test_diagonal_block_factor_is_accurate_to_a_few_ulp():
    """The pivot chain of the diagonal-block kernel uses its own 1/sqrt (hardware seed + one third-order step, csrc/linalg.cu
    d2_rsqrt).  A sloppy seed would still pass the 1e-10 likelihood tolerance, so the factor itself is checked: the residual
    K - L L^T of a 128 x 128 block, accumulated in extended precision, must stay within a few ulp of |K|."""
    eng = _eng()
    n = 128
    x, y = _data(n)
    tree, hp = TREES["composite"]
    prog = eng.DeviceProgram.get(tree, 1, False, 1)
    plan = eng.Plan([prog], [n], want_grad=True)
    plan.set_data(0, torch.tensor(x), torch.tensor(y)); plan.set_hp(0, _flat(hp), 1e-2)
    dev = torch.device("cuda")
    K = eng.assemble(prog, torch.tensor(x, device=dev), None, torch.tensor(_flat(hp), device=dev),
                     torch.tensor([1e-2], dtype=torch.float64, device=dev)).cpu().numpy()
    plan.eval(eng.STAGES_LML)
    L = torch.tril(plan.lower_matrix(0, eng.BUF_A)).cpu().numpy()
    Lx = L.astype(np.longdouble)
    resid = np.abs(K.astype(np.longdouble) - Lx @ Lx.T)
    eps = np.finfo(np.float64).eps
    assert float(resid.max()) <= 24 * eps * float(np.abs(K).max()), float(resid.max() / (eps * np.abs(K).max()))
    # and the inverse of the block: W L = I to rounding (W = inv(L) from the recursive doubling inside the kernel)
    plan.eval(eng.STAGE_INVERSE)
    W = torch.tril(plan.lower_matrix(0, eng.BUF_A)).cpu().numpy().astype(np.longdouble)
    ident = np.abs(W @ Lx - np.eye(n))
    cond = float(np.abs(W).max() * np.abs(L).max())
    assert float(ident.max()) <= 64 * eps * cond, (float(ident.max()), cond)


def test_large_batch_is_reproducible_and_matches_single_gp_plans():
    """Regression (round 2): the diagonal-block kernel wrote a micro-block back in place while slower warps were still
    reading it - invisible in small batches, non-reproducible results once ~1000 blocks kept every SM busy.  600 GPs of
    three block rows each: three evaluations must agree bit for bit, every GP must be finite, a sample must agree with
    one-GP plans bit for bit (same kernels, same order of operations) and with the oracle."""
    eng = _eng()
    tree, hp = TREES["se"]
    B, n = 600, 384
    rng = np.random.default_rng(77)
    xs = [np.sort(rng.uniform(0, 1, (n, 1)), axis=0) for _ in range(B)]
    ys = [np.sin((5 + b % 9) * xs[b]) + 0.1 * rng.standard_normal((n, 1)) for b in range(B)]
    prog = eng.DeviceProgram.get(tree, 1, False, 1)
    plan = eng.Plan([prog] * B, [n] * B, want_grad=True)
    for b in range(B):
        plan.set_data(b, torch.tensor(xs[b]), torch.tensor(ys[b])); plan.set_hp(b, _flat(hp), 1e-2)
    runs = []
    for _ in range(3):
        plan.eval(eng.STAGES_LML_GRAD)
        torch.cuda.synchronize()
        nll, grads, info = plan.results()
        assert int(np.max(info)) == 0 and np.all(np.isfinite(nll))
        runs.append((nll.copy(), np.stack(grads)))
    for nll, g in runs[1:]:
        assert np.array_equal(nll, runs[0][0]) and np.array_equal(g, runs[0][1])
    single = eng.Plan([prog], [n], want_grad=True)
    for b in list(range(0, B, 37)):
        single.set_data(0, torch.tensor(xs[b]), torch.tensor(ys[b])); single.set_hp(0, _flat(hp), 1e-2)
        single.eval(eng.STAGES_LML_GRAD)
        torch.cuda.synchronize()
        nll1, g1, info1 = single.results()
        assert nll1[0] == runs[0][0][b] and np.array_equal(g1[0], runs[0][1][b]), b
    for b in (3, 299, 598):
        ref, gref, gnoise = orc.nll_and_grad(tree, hp, 1e-2, xs[b], ys[b], reference_distance=False)
        assert abs(runs[0][0][b] - ref) <= LL_RTOL * abs(ref)
        gflat = np.concatenate([np.asarray(t).reshape(-1) for t in gref] + [[gnoise]])
        assert np.max(np.abs(runs[0][1][b] - gflat)) <= GRAD_RTOL * np.max(np.abs(gflat))
