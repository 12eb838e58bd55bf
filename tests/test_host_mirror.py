"""CPU: host-side mirror of the reference's operator surface (no device work): tree utilities, hyper-parameter order,
defaults, bounds, segment / partition bookkeeping - against the golden vectors of the unmodified reference - and the
C-ABI library's exported symbols."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import gaussianprocessfundamentals_b200.global_parameters as global_param
from gaussianprocessfundamentals_b200 import _lib
from gaussianprocessfundamentals_b200.DataHandling import DataInput as di
from gaussianprocessfundamentals_b200.KernelBasics import BaseKernels as bk
from gaussianprocessfundamentals_b200.KernelBasics import Operators as op
from gaussianprocessfundamentals_b200.KernelBasics import PartitioningModel as pm
from gaussianprocessfundamentals_b200.KernelBasics import PartitionOperator as po
from gaussianprocessfundamentals_b200.MeanFunctionBasics import BaseMeanFunctions as bmf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
LEAF = {"SE": bk.SquaredExponentialKernel, "PER": bk.PeriodicKernel, "LIN": bk.LinearKernel,
        "MAT32": bk.MaternKernel3_2, "MAT52": bk.MaternKernel5_2, "WN": bk.WhiteNoiseKernel}


def build(spec, d=1):
    kind = spec[0]
    if kind in LEAF:
        return LEAF[kind](d)
    children = [build(c, d) for c in spec[1]]
    if kind == "ADD":
        return op.AdditionOperator(d, children)
    if kind == "MUL":
        return op.MultiplicationOperator(d, children)
    if kind == "CP":
        return op.ChangePointOperator(d, children, [torch.tensor(c, dtype=torch.float64) for c in spec[2]])
    raise ValueError(kind)


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    return z, json.loads(bytes(z["__meta__"]).decode("utf-8"))


def test_strings_names_and_counts_match_reference(gold):
    z, meta = gold
    for name, m in meta.items():
        if not isinstance(m, dict) or m.get("kind") != "holistic":
            continue
        global_param.p_scaled_base_kernel = m["scaled"]
        try:
            kern = build(json.loads(m["spec"]))
            assert kern.get_string_representation() == m["string"], name
            assert kern.get_hyper_parameter_names(0) == m["names"], name
            assert kern.get_number_of_hyper_parameter() == int(z[name + "/n_hp"]), name
        finally:
            global_param.p_scaled_base_kernel = False


def test_sort_changes_hp_order_like_reference(gold):
    _, meta = gold
    spec = ["MUL", [["ADD", [["SE"], ["PER"]]], ["LIN"]]]
    k1, k2 = build(spec), build(spec)
    assert k1.get_string_representation() == meta["sort_kat"]["before"]
    assert k1.type_compare_to(k2)
    assert k1.get_string_representation() == meta["sort_kat"]["after"]
    assert k1.get_hyper_parameter_names(0) == meta["sort_kat"]["names_after"]
    assert k1.to_spec() == ("MUL", [("LIN",), ("ADD", [("PER",), ("SE",)])])


def test_defaults_and_bounds_match_reference(gold):
    z, _ = gold
    kern = build(["MUL", [["ADD", [["SE"], ["PER"]]], ["LIN"]]])
    xr = [[0.25, 2.25]]
    d = np.concatenate([h.numpy().reshape(-1) for h in kern.get_default_hyper_parameter(xr, 500)])
    assert np.array_equal(d, z["defaults/composite"])
    b = kern.get_hyper_parameter_bounds(xr, 500)
    lo = np.concatenate([np.asarray(v[0]).reshape(-1) for v in b])
    hi = np.concatenate([np.asarray(v[1]).reshape(-1) for v in b])
    assert np.allclose(lo, z["bounds/composite_lo"], rtol=1e-15, atol=0, equal_nan=True)
    assert np.allclose(hi, z["bounds/composite_hi"], rtol=1e-15, atol=0, equal_nan=True)


def test_blockwise_segments_bit_exact(gold):
    z, _ = gold
    x, y, cps = z["blockwise/x"], z["blockwise/y"], z["blockwise/cps"]
    b = di.BlockwiseDataInput(x, y, x, y, [torch.tensor(c, dtype=torch.float64) for c in cps])
    assert len(b.data_inputs) == len(cps) + 1
    for i, blk in enumerate(b.data_inputs):
        assert blk.n_train == int(z["blockwise/seg%d_n" % i])
        assert np.array_equal(blk.data_x_train.numpy(), z["blockwise/seg%d_x" % i])


def test_partition_indices_and_reordering_bit_exact(gold):
    z, _ = gold
    x, y, edges = z["partition/x"], z["partition/y"], z["partition/edges"]
    model = pm.PartitioningModel(pm.PartitioningClass.SELF_SUFFICIENT, [])
    model.init_partitioning([pm.IntervalCriterion(edges[i], edges[i + 1]) for i in range(4)])
    idx = model.get_data_record_indices_per_partition(x)
    for i, ix in enumerate(idx):
        assert np.array_equal(np.asarray(ix, dtype=np.int64), z["partition/idx%d" % i])
    pdi = model.partition_data_input(di.DataInput(x, y, x, y))
    assert np.array_equal(pdi.data_x_train.numpy(), z["partition/x_reordered"])
    assert [blk.n_train for blk in pdi.data_inputs] == [len(z["partition/idx%d" % i]) for i in range(4)]
    copy = model.deepcopy()
    assert copy.get_number_of_partitions() == 4 and copy.partition_class == model.partition_class


def test_child_slices_advance_over_change_points_and_empty_blocks():
    kern = build(["CP", [["SE"], ["PER"], ["ADD", [["SE"], ["LIN"]]]], [0.3, 0.6]])
    assert kern.get_number_of_hyper_parameter() == 2 + 1 + 2 + 2
    assert [(s.start, s.stop) for s in kern.child_slices()] == [(2, 3), (3, 5), (5, 7)]
    hp = kern.get_default_hyper_parameter([[0.0, 1.0]], 100)
    assert [float(h.reshape(-1)[0]) for h in hp[:2]] == [0.3, 0.6]
    kern.set_last_hyper_parameter(hp)
    assert len(kern.get_last_hyper_parameter()) == 7


def test_simplification_and_pruning():
    se, per, lin = bk.SquaredExponentialKernel(1), bk.PeriodicKernel(1), bk.LinearKernel(1)
    prod = op.MultiplicationOperator(1, [op.AdditionOperator(1, [se, per]), lin])
    assert prod.get_simplified_version().get_string_representation() == "((LIN x SE) + (LIN x PER))"
    cp = build(["CP", [["SE"], ["PER"], ["LIN"]], [0.4, 1.7]])
    pruned, changed = cp.get_simplified_kernel([0.0, 1.0])
    assert changed and pruned.get_string_representation() == "(SE ][ PER)"
    same, changed = build(["CP", [["SE"], ["PER"]], [0.4]]).get_simplified_kernel([0.0, 1.0])
    assert not changed


def test_unscaling_of_fitted_hyper_parameters():
    se, per, lin = bk.SquaredExponentialKernel(1), bk.PeriodicKernel(1), bk.LinearKernel(1)
    k = op.AdditionOperator(1, [se, per, lin])
    k.set_last_hyper_parameter([torch.tensor(-0.2, dtype=torch.float64), torch.tensor(0.3, dtype=torch.float64),
                                torch.tensor(-0.4, dtype=torch.float64), torch.tensor([0.5], dtype=torch.float64)])
    raw = [float(h.reshape(-1)[0]) for h in k.get_last_hyper_parameter()]
    assert raw == [0.2, 0.3, 0.4, 0.5]            # |l|, |p| are stored (BaseKernels.py:429-432, :629-634)
    sc = [float(h.reshape(-1)[0]) for h in k.get_last_hyper_parameter((10.0, 2.0))]
    assert sc == [0.4, 0.6, 0.8, 0.5 * 2.0 + 10.0]  # l*s1, l*s1, p*s1, c*s1+s0 (BaseKernels.py:259-269,417-427,617-627)


def test_library_exports_every_declared_symbol():
    """the C-ABI shared library loads on a machine without a GPU and exports what include/gpb.h declares"""
    header = open(os.path.join(ROOT, "include", "gpb.h")).read()
    declared = set(re.findall(r"\b(gpb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.gpb_version.restype = ctypes.c_int
    assert lib.gpb_version() >= 100


def test_no_cpu_fallback():
    from gaussianprocessfundamentals_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.GpbError):
        engine.require_cuda()
    kern = bk.SquaredExponentialKernel(1)
    x = np.linspace(0, 1, 8)[:, None]
    with pytest.raises(_lib.GpbError):
        kern.get_tf_tensor([torch.tensor(0.1, dtype=torch.float64)], x, x)


def test_distributed_layout_host_arithmetic():
    """gpb_dist_owner / gpb_dist_panel_segments (include/gpb.h): the block -> rank map and the staging order of a panel's
    tiles, against a plain restatement.  No GPU, no NCCL involved."""
    import ctypes
    lib = _lib.load()
    for P, Q in ((1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (4, 2), (1, 8), (8, 1), (3, 2)):
        for I in range(0, 20):
            for J in range(0, 20):
                assert lib.gpb_dist_owner(I, J, P, Q) == (I % P) * Q + (J % Q)
        seen = set()
        for I in range(P * 3):
            for J in range(Q * 3):
                seen.add(lib.gpb_dist_owner(I, J, P, Q))
        assert seen == set(range(P * Q))                       # every rank owns blocks
        for k in (0, 1, 5, 17):
            for nt in (0, 1, 2, 7, 33):
                base, cnt, first = ((ctypes.c_int * 8)(), (ctypes.c_int * 8)(), (ctypes.c_int * 8)())
                assert lib.gpb_dist_panel_segments(k, nt, P, base, cnt, first) == 0
                slots = {}
                for o in range(P):
                    tiles = [t for t in range(nt) if (k + t) % P == o]     # tiles of process row o, ascending
                    assert cnt[o] == len(tiles)
                    if tiles:
                        assert first[o] == tiles[0]
                    for idx, t in enumerate(tiles):
                        slots[t] = base[o] + idx
                # the staging order is a permutation of the panel's tiles, contiguous per process row
                assert sorted(slots.values()) == list(range(nt))
    assert lib.gpb_dist_owner(-1, 0, 1, 1) == -1
    # ownership in groups of W block columns (1 x Q grids: the owner factorises a whole outer panel)
    assert lib.gpb_dist_col_width(1) in (1, 2, 4) and lib.gpb_dist_col_width(2) == 1
    for Q in (1, 2, 3, 4, 8):
        for W in (1, 2, 4):
            for I in range(6):
                for J in range(40):
                    assert lib.gpb_dist_owner_w(I, J, 1, Q, W) == (J // W) % Q
                    assert lib.gpb_dist_owner_w(I, J, 2, Q, W) == (I % 2) * Q + (J // W) % Q
            for q in range(Q):
                for lo in range(0, 23):
                    for hi in (lo, lo + 1, lo + 2, lo + 7, 40):
                        want = [J for J in range(lo, hi) if (J // W) % Q == q]
                        buf = (ctypes.c_int * 64)()
                        assert lib.gpb_dist_owned_cols(lo, hi, Q, q, W, buf, 64) == len(want)
                        assert list(buf[:len(want)]) == want
    assert lib.gpb_dist_panel_segments(0, 4, 9, base, cnt, first) != 0      # P > 8 is refused


def test_lbfgs_fitter_host_logic_without_device(monkeypatch):
    """LbfgsFitter's host side - flatten / unflatten of the hyper-parameter list, bounds from the kernel, a non-PD trial
    point reported as +inf, best point stored on the kernel - with the device objective replaced by a quadratic."""
    from gaussianprocessfundamentals_b200.Optimizer import Fitter as fitter
    from gaussianprocessfundamentals_b200 import global_parameters as gparam

    kern = op.AdditionOperator(1, [bk.SquaredExponentialKernel(1), bk.LinearKernel(1)])
    f = fitter.LbfgsFitter.__new__(fitter.LbfgsFitter)
    f.max_evaluations, f.tolerance, f.history, f.result = 50, 1e-12, [], None

    class _Cov:
        kernel = kern
    class _Gp:
        covariance_matrix = _Cov()
    _Gp.kernel = kern
    f._gp = _Gp()
    target = [torch.tensor(0.3, dtype=torch.float64), torch.tensor([0.7], dtype=torch.float64)]
    calls = {"n": 0}

    def objective(variables):
        calls["n"] += 1
        if float(variables[0]) > 5.0:
            raise ArithmeticError("not positive definite")
        val = sum(float(((v - t) ** 2).sum()) for v, t in zip(variables, target))
        return val, [2 * (v - t) for v, t in zip(variables, target)]

    f._objective = objective
    f._metric_value = lambda variables: torch.tensor([[objective(variables)[0]]], dtype=torch.float64)
    f._initial_variables = lambda: ([torch.tensor(1.0, dtype=torch.float64), torch.tensor([0.01], dtype=torch.float64)],
                                    [[0.0, 1.0]], 100)
    monkeypatch.setattr(gparam, "p_check_hyper_parameters", False)
    monkeypatch.setattr(gparam, "p_optimize_noise", False)
    pre, post, hps, noise, _ = f.fit()
    assert float(post) < 1e-12 < float(pre)
    assert abs(float(torch.as_tensor(hps[0])) - 0.3) < 1e-6 and abs(float(torch.as_tensor(hps[1]).reshape(-1)[0]) - 0.7) < 1e-6
    assert [tuple(torch.as_tensor(h).shape) for h in hps] == [(), (1,)]
    # bounds: lengthscale in [5 R / n, R / 3] = [0.05, 1/3]; the linear offset is unbounded
    monkeypatch.setattr(gparam, "p_check_hyper_parameters", True)
    b = f._flat_bounds(f._initial_variables()[0], [[0.0, 1.0]], 100)
    assert b[0][0] == pytest.approx(0.05) and b[0][1] == pytest.approx(1.0 / 3.0) and b[1] == (None, None)
    target[0] = torch.tensor(2.0, dtype=torch.float64)           # optimum outside the box: the fit stops at the bound
    pre, post, hps, noise, _ = f.fit()
    assert abs(float(torch.as_tensor(hps[0])) - 1.0 / 3.0) < 1e-9


def test_overlapped_inverse_task_graph_respects_its_dependencies():
    """gpb_trtri_schedule (include/gpb.h) runs the scheduler that gpb_plan_eval uses to overlap the recursive-doubling
    inverse with the factorisation - host arithmetic only.  For several sizes and progress sequences: every task is
    launched exactly once, never before the columns of L it reads are final, T before W of a sub-problem, and the inverses
    of the two diagonal halves (all lower-level tasks and block copies inside them) before the task that multiplies
    with them."""
    import ctypes
    import random
    lib = _lib.load()
    NB = 128

    def schedule(n, fcs):
        arr = (ctypes.c_longlong * len(fcs))(*fcs)
        cap = 4 * (n // NB + 2) * 4
        out = (ctypes.c_int * (4 * cap))()
        cnt = lib.gpb_trtri_schedule(n, arr, len(fcs), out, cap)
        assert 0 <= cnt <= cap, (n, cnt)
        return [tuple(out[4 * i:4 * i + 4]) for i in range(cnt)]

    def check(n, fcs):
        tasks = schedule(n, fcs)
        pos = {}
        for i, (kind, s, p, fc) in enumerate(tasks):
            assert (kind, s, p) not in pos
            pos[(kind, s, p)] = i
        nblk = (n + NB - 1) // NB
        levels = []
        s = NB
        while s < n:
            levels.append((s, (n - s + 2 * s - 1) // (2 * s)))
            s *= 2
        assert all((0, NB, j) in pos for j in range(nblk))
        assert all((1, s, p) in pos and (2, s, p) in pos for s, cnt in levels for p in range(cnt))
        assert len(tasks) == nblk + 2 * sum(cnt for _, cnt in levels)

        def inverse_complete_before(lo, hi, i):
            hi = min(hi, n)
            if any(pos[(0, NB, j)] >= i for j in range(lo // NB, (hi + NB - 1) // NB)):
                return False
            for s, cnt in levels:
                for p in range(cnt):
                    r0, rA = 2 * s * p, 2 * s * p + s
                    if r0 >= lo and rA < hi and pos[(2, s, p)] >= i:
                        return False
            return True

        for (kind, s, p), i in pos.items():
            fc = tasks[i][3]
            if kind == 0:
                assert min((p + 1) * NB, n) <= fc
            else:
                r0, rA = 2 * s * p, 2 * s * p + s
                if kind == 1:
                    assert rA <= fc and inverse_complete_before(r0, rA, i)
                else:
                    assert min(rA + s, n) <= fc and pos[(1, s, p)] < i and inverse_complete_before(rA, rA + s, i)

    rng = random.Random(3)
    for n in (129, 255, 256, 257, 1000, 1024, 1025, 2321, 4096, 5000, 8192):
        nblk = (n + NB - 1) // NB
        two = []
        k = 0
        while k < nblk:                                   # the factorisation's progress with 256-wide outer panels
            kb = 2 if k + 2 <= nblk - 1 else 1
            two.append(min((k + kb) * NB, n))
            k += kb
        check(n, two + [n])
        check(n, [min((k + 1) * NB, n) for k in range(nblk)] + [n])
        check(n, sorted(rng.sample(range(0, n + 1), min(9, n))) + [n])
        check(n, [n])
    # a progress sequence that never reaches n leaves the graph undrained: reported as a negative count
    arr = (ctypes.c_longlong * 1)(512)
    assert lib.gpb_trtri_schedule(2048, arr, 1, None, 0) < 0
