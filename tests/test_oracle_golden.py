"""CPU: pins the oracle (oracle/gp_oracle.py) against the golden vectors produced by the UNMODIFIED reference sources
(tests/golden/make_golden.py), and checks it independently against mpmath and closed forms."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import gp_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    meta = json.loads(bytes(z["__meta__"]).decode("utf-8"))
    return z, meta


def _tree(spec):
    """json spec (lists) -> oracle tuple tree; a CP node carries its change points as third element"""
    kind = spec[0]
    if len(spec) == 1:
        return (kind,)
    return (kind, [_tree(c) for c in spec[1]])


def holistic_names(meta):
    return [k for k, v in meta.items() if isinstance(v, dict) and v.get("kind") == "holistic"]


def _case(z, meta, name):
    m = meta[name]
    tree = _tree(json.loads(m["spec"]))
    n_hp = int(z[name + "/n_hp"])
    hp = [z[name + "/hp%d" % i] for i in range(n_hp)]
    grads = [z[name + "/grad%d" % i] for i in range(n_hp)]
    return tree, hp, grads, m


def test_golden_file_present(gold):
    z, meta = gold
    assert len(holistic_names(meta)) >= 12


@pytest.mark.parametrize("faithful", [True, False])
def test_oracle_matches_reference_nll_and_gradients(gold, faithful):
    z, meta = gold
    for name in holistic_names(meta):
        tree, hp, grads, m = _case(z, meta, name)
        x, y, noise = z[name + "/x"], z[name + "/y"], float(z[name + "/noise"])
        val, g, gn = orc.nll_and_grad(tree, hp, noise, x, y, scaled=m["scaled"], cp_mode=m["cp_mode"],
                                      reference_distance=faithful, optimize_noise=m["optimize_noise"])
        ref = float(z[name + "/nll"][0])
        # default jitter 1e-8 is numerically singular: the restatement is op-for-op identical, the
        # directly-summed distance is not expected to reproduce those digits
        loose = name == "se_default_jitter_n100" and not faithful
        tol = 1e-6 if loose else (1e-13 if faithful else 1e-10)
        assert abs(val - ref) <= tol * max(1.0, abs(ref)), (name, val, ref)
        gtol = 1e-4 if loose else (1e-11 if faithful else 1e-8)
        scale = max(np.max(np.abs(np.concatenate([np.reshape(v, -1) for v in grads] + [[1e-300]]))), 1e-12)
        for gi, gr in zip(g, grads):
            assert np.max(np.abs(np.reshape(gi, -1) - np.reshape(gr, -1))) <= gtol * scale, name
        assert abs(gn - float(z[name + "/grad_noise"])) <= gtol * max(abs(float(z[name + "/grad_noise"])), scale), name


def test_indicator_change_points_get_no_gradient(gold):
    z, meta = gold
    name = "cp_indicator_n240"
    assert bool(z[name + "/grad0_none"]) and bool(z[name + "/grad1_none"])       # TF: None for the comparison op
    assert not bool(z["cp_approx_n240/grad0_none"])


def test_oracle_matrices_match_reference(gold):
    z, meta = gold
    for name in ("se_n64_mats", "composite_n96_mats"):
        tree, hp, _, m = _case(z, meta, name)
        hp_t = [torch.tensor(h) for h in hp]
        out, K, L, alpha = orc.nll(tree, hp_t, torch.tensor(float(z[name + "/noise"]), dtype=torch.float64), torch.tensor(z[name + "/x"]),
                                   torch.tensor(z[name + "/y"]), return_parts=True)
        assert np.array_equal(K.numpy(), z[name + "/K"])
        # the factor is compared through its backward error and loosely entry-wise: LAPACK's blocked reduction order
        # depends on the thread count and these small matrices are ill-conditioned
        Kn = z[name + "/K"] + float(z[name + "/noise"]) * np.eye(K.shape[0])
        for Lm in (L.numpy(), z[name + "/L"]):
            assert np.max(np.abs(Lm @ Lm.T - Kn)) <= 1e-13 * np.max(np.abs(Kn))
        assert np.max(np.abs(L.numpy() - z[name + "/L"])) <= 1e-7 * np.max(np.abs(z[name + "/L"]))
        assert np.max(np.abs(alpha.numpy() - z[name + "/alpha"])) <= 1e-6 * np.max(np.abs(z[name + "/alpha"]))


def test_blockwise_segments_and_block_nll(gold):
    z, meta = gold
    x, y, cps = z["blockwise/x"], z["blockwise/y"], z["blockwise/cps"]
    segs = orc.blockwise_segments(x, cps)
    specs = json.loads(meta["blockwise"]["specs"])
    sizes = meta["blockwise"]["hp_sizes"]
    flat = z["blockwise/hp_children"]
    counts = [orc.n_hp_entries(_tree(s)) for s in specs]
    slices = orc.hp_slices(counts)
    total = 0.0
    for i, seg in enumerate(segs):
        assert int(z["blockwise/seg%d_n" % i]) == len(seg)
        assert np.array_equal(x[seg], z["blockwise/seg%d_x" % i])
        ref = z["blockwise/block_nll"][i]
        if len(seg) == 0:
            assert np.isnan(ref)
            continue
        hp = [torch.tensor(np.atleast_1d(v) if _tree(specs[i])[0] == "LIN" else v) for v in flat[slices[i]]]
        if specs[i][0] == "ADD":   # [SE l, LIN c]: c is a [1] vector
            hp = [torch.tensor(flat[slices[i]][0]), torch.tensor([flat[slices[i]][1]])]
        v = float(orc.nll(_tree(specs[i]), hp, torch.tensor(1e-2, dtype=torch.float64), torch.tensor(x[seg]),
                          torch.tensor(y[seg])))
        assert abs(v - ref) <= 1e-12 * abs(ref), i
        total += v
    # holistic change-point kernel == sum over (non-empty) blocks  
    assert abs(total - float(z["blockwise/holistic_nll"][0])) <= 1e-10 * abs(total)


def test_partition_indices_bit_exact(gold):
    z, meta = gold
    x, edges = z["partition/x"], z["partition/edges"]
    cols = [np.logical_and(x[:, 0] >= edges[i], x[:, 0] < edges[i + 1]).astype(np.float64) for i in range(4)]
    idx = orc.partition_indices(cols)
    order = []
    for i, ix in enumerate(idx):
        assert np.array_equal(ix.astype(np.int64), z["partition/idx%d" % i])
        order.extend(ix.tolist())
    assert np.array_equal(x[np.asarray(order, dtype=np.int64)], z["partition/x_reordered"])


def test_mpmath_spot_check():
    """independent 50-digit evaluation of the SE-kernel NLL at n = 12"""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    rng = np.random.default_rng(7)
    n = 12
    x = np.linspace(0, 1, n)
    y = rng.standard_normal(n)
    l, s2 = 0.3, 0.05
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = mp.e ** (-(mp.mpf(float(x[i])) - mp.mpf(float(x[j]))) ** 2 / (2 * mp.mpf(l) ** 2))
        K[i, i] += mp.mpf(s2)
    yv = mp.matrix([mp.mpf(float(v)) for v in y])
    alpha = mp.lu_solve(K, yv)
    nll = (yv.T * alpha)[0] / 2 + mp.log(mp.det(K)) / 2 + mp.mpf(n) / 2 * mp.log(2 * mp.pi)
    val, _, _ = orc.nll_and_grad(("SE",), [l], s2, x[:, None], y[:, None], reference_distance=False)
    assert abs(val - float(nll)) <= 1e-12 * abs(float(nll))


def test_closed_form_white_noise():
    n = 40
    rng = np.random.default_rng(3)
    x = np.linspace(0, 1, n)[:, None]
    y = rng.standard_normal((n, 1))
    val, g, gn = orc.nll_and_grad(("WN",), [], 0.5, x, y)
    want = 0.5 * float((y.T @ y)[0, 0]) / 1.5 + 0.5 * n * math.log(1.5) + 0.5 * n * math.log(2 * math.pi)
    assert abs(val - want) <= 1e-13 * abs(want)
    assert abs(gn - (0.5 * n / 1.5 - 0.5 * float((y.T @ y)[0, 0]) / 1.5 ** 2)) <= 1e-12


def test_reference_outputs_against_lapack_independent_of_torch(gold):
    """The golden vectors were produced by the reference's own code on a TensorFlow stand-in whose leaf ops are torch's.
    Guard against a shared torch artefact: recompute NLL, L and alpha of the stored cases with NumPy / SciPy LAPACK only
    (no torch, no autodiff), from the kernel matrix K the reference itself produced."""
    import scipy.linalg as sla
    z, meta = gold
    checked = 0
    for name, m in meta.items():
        if not (isinstance(m, dict) and m.get("kind") == "holistic" and (name + "/K") in z.files):
            continue
        K, y, s2 = z[name + "/K"], z[name + "/y"], float(z[name + "/noise"])
        n = K.shape[0]
        c, low = sla.cho_factor(K + s2 * np.eye(n), lower=True)
        L = np.tril(c)
        alpha = sla.cho_solve((c, low), y)
        nll = 0.5 * float(y.T @ alpha) + float(np.sum(np.log(np.diag(L)))) + 0.5 * n * np.log(2 * np.pi)
        assert abs(nll - float(z[name + "/nll"][0])) <= 1e-12 * abs(nll), name
        assert np.max(np.abs(L - z[name + "/L"])) <= 1e-12 * np.max(np.abs(L)), name
        assert np.max(np.abs(alpha - z[name + "/alpha"])) <= 1e-9 * np.max(np.abs(alpha)), name
        # and the kernel matrix itself from the formulas of SURVEY App. A, in plain NumPy, for the SE case
        if json.loads(m["spec"]) == ["SE"]:
            x, l = z[name + "/x"], float(z[name + "/hp0"])
            assert np.max(np.abs(K - np.exp(-0.5 * (x - x.T) ** 2 / (l * l)))) <= 1e-14, name
        checked += 1
    assert checked >= 2


# ---- round-2 goldens: rank-3 batch likelihood, partitioned K_s, blockwise metrics -------------------------------------
@pytest.fixture(scope="module")
def gold2():
    z = np.load(os.path.join(os.path.dirname(GOLD), "reference_golden_r2.npz"))
    meta = json.loads(bytes(z["__meta__"]).decode("utf-8"))
    return z, meta


@pytest.mark.parametrize("name", ["batch3_se", "batch3_composite"])
def test_oracle_batch_aggregate_matches_reference(gold2, name):
    """Metrics/LogLikelihood.py:49,62-63 + Metrics/Metrics.py:152-154: value and gradient"""
    z, meta = gold2
    tree = _tree(json.loads(meta[name]["spec"]))
    flat, sizes = z[name + "/hp"], meta[name]["hp_sizes"]
    hp, pos = [], 0
    for sz in sizes:        # one-element vectors broadcast like the reference's rank-0 entries
        hp.append(torch.tensor(flat[pos:pos + sz], dtype=torch.float64, requires_grad=True))
        pos += sz
    raw = torch.tensor(float(z[name + "/noise"]), dtype=torch.float64, requires_grad=True)
    x, y = torch.tensor(z[name + "/x"]), torch.tensor(z[name + "/y"])
    val = orc.batch_nll(tree, hp, raw, x, y)
    grads = torch.autograd.grad(val, hp + [raw])
    want = float(z[name + "/nll"][0])
    assert abs(float(val) - want) <= 1e-12 * abs(want)
    g = np.concatenate([t.numpy().reshape(-1) for t in grads[:-1]])
    assert np.max(np.abs(g - z[name + "/grad"])) <= 1e-10 * np.max(np.abs(z[name + "/grad"]))
    assert abs(float(grads[-1]) - float(z[name + "/grad_noise"][0])) <= 1e-10 * abs(float(z[name + "/grad_noise"][0]))
    # the aggregate is NOT the mean of the per-entry likelihoods (the quirk is material)
    per = z[name + "/per_entry_nll"]
    assert abs(np.mean(per) - want) > 1.0
    mean_val = orc.batch_nll(tree, [h.detach() for h in hp], raw.detach(), x, y, reference_aggregate=False)
    assert abs(float(mean_val) - np.mean(per)) <= 1e-11 * abs(np.mean(per))


def test_oracle_partitioned_K_s_matches_reference(gold2):
    """block-rectangular K_s with dead rows / columns (Auxiliary/NonSquareBlockMatrices.py:8-103)"""
    z, meta = gold2
    for name in ("ppred_full", "ppred_dead"):
        specs = json.loads(meta[name]["specs"])
        sizes = meta[name]["hp_sizes"]
        flat = z[name + "/hp"]
        x, xt = z[name + "/x"], z[name + "/xt"]
        edges = z[name + "/edges"]
        Ks = z[name + "/K_s"]
        assert Ks.shape == (x.shape[0], xt.shape[0])
        r = c = pos = 0
        counts = [orc.n_hp_entries(_tree(sp)) for sp in specs]
        it = iter(sizes)
        for i, sp in enumerate(specs):
            lo, hi = edges[i], edges[i + 1]
            a = np.where(np.logical_and(x[:, 0] >= lo, x[:, 0] < hi))[0]
            b = np.where(np.logical_and(xt[:, 0] >= lo, xt[:, 0] < hi))[0]
            assert np.array_equal(a, z[name + "/idx%d" % i]) and np.array_equal(b, z[name + "/idxt%d" % i])
            hp = []
            for _ in range(counts[i]):
                sz = next(it)
                hp.append(torch.tensor(flat[pos:pos + sz]))
                pos += sz
            if len(a) and len(b):
                blk = orc.kernel_matrix(_tree(sp), hp, torch.tensor(x[a]), torch.tensor(xt[b])).numpy()
                got = Ks[r:r + len(a), c:c + len(b)]
                assert np.max(np.abs(got - blk)) <= 1e-13 * max(1.0, np.max(np.abs(blk)))
                # everything outside the diagonal blocks is exactly zero
                assert np.all(Ks[r:r + len(a), :c] == 0) and np.all(Ks[r:r + len(a), c + len(b):] == 0)
            r += len(a)
            c += len(b)
        assert (r, c) == Ks.shape
