"""GPU parity AT THE BENCHED SIZES (VERDICT r1 "parity is pinned only at small n"): the exact configurations bench.py
times are compared with the CPU oracle (oracle/gp_oracle.py, itself pinned against the unmodified reference) -
  C2    (SE+PER)xLIN, n = 8192: the look-ahead driver with 256-wide outer panels, the triangular inverse overlapped with
        the factorisation and CUDA-graph replay through gpb_plan_eval_host, plus n = 6145 / 8191 (ragged last block on
        both sides of the kb = 2 boundary)
  C3    16 heterogeneous candidate kernels of bench.py's grammar at n = 2048, one batched plan
  C4    32 partition blocks of n = 1024 through PartitionedGaussianProcess + blockwise_LL, index bookkeeping bit-exact
  M16k  n = 16384 likelihood against LAPACK (scipy cho_factor) on the oracle's covariance matrix
Tolerances (north_star): relative <= 1e-10 on the likelihood, <= 1e-8 on gradients."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from oracle import gp_oracle as orc
from tests import benched_cases as bc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
LL_RTOL, GRAD_RTOL = 1e-10, 1e-8
# The live oracle of the test box against the values frozen in the build container: the oracle's own CPU Cholesky is
# good to cond(K) * eps only (C2 at n = 8192: cond ~ 1e6; one pool box came out 1.9e-10 away), so the DEVICE is held to
# the north-star tolerance against the committed values and the live oracle to this looser one.
LIVE_ORACLE_RTOL = 1e-8
COMPOSITE, C2_HP = bc.COMPOSITE, bc.C2_HP
with open(os.path.join(ROOT, "tests", "golden", "oracle_benched_sizes.json")) as _f:
    GOLD = json.load(_f)


def _eng():
    from gaussianprocessfundamentals_b200 import engine
    return engine


def _flatten(gref, gnoise):
    return np.concatenate([np.asarray(g).reshape(-1) for g in gref] + [[gnoise]])


def _check(nll, grad, gold, what):
    """device vs the frozen oracle values `gold` = {"reference_distance": {nll, grad}, "direct_distance": {nll, grad}}.

    Arbitration: the reference computes r^2 as a^2 - 2ab + b^2, whose cancellation noise an
    ill-conditioned matrix amplifies beyond 1e-10 all by itself; the device sums (x - x')^2 directly.  The device must
    agree with the reference-formula value to the tolerance, or else agree with the cancellation-free CPU evaluation to
    the tolerance and be no farther from the reference-formula value than twice the distance between the two."""
    ref, exact = gold["reference_distance"], gold["direct_distance"]
    g = np.asarray(ref["grad"])
    if abs(nll - ref["nll"]) > LL_RTOL * abs(ref["nll"]):
        assert abs(nll - exact["nll"]) <= LL_RTOL * abs(exact["nll"]), (what, nll, exact["nll"], ref["nll"])
        assert abs(nll - ref["nll"]) <= 2.0 * abs(exact["nll"] - ref["nll"]), (what, nll, exact["nll"], ref["nll"])
        g = np.asarray(exact["grad"])
    assert np.max(np.abs(grad - g)) <= GRAD_RTOL * np.max(np.abs(g)), (what, grad, g)


def _check_live_oracle(live_nll, live_grad, gold, what):
    ref = gold["reference_distance"]
    assert abs(live_nll - ref["nll"]) <= LIVE_ORACLE_RTOL * abs(ref["nll"]), ("live oracle vs frozen", what, live_nll, ref["nll"])
    g = np.asarray(ref["grad"])
    assert np.max(np.abs(live_grad - g)) <= 1e-6 * np.max(np.abs(g)), ("live oracle vs frozen", what)


@pytest.mark.parametrize("n", list(bc.C2_SIZES))
def test_c2_benched_path_matches_oracle(n):
    """the exact data, kernel, hyper-parameters and launch path of bench.py's C2 workload (n = 8192)"""
    eng = _eng()
    x, y = bc.c2_inputs(n)
    gold = GOLD["c2"][str(n)]
    live, gref, gnoise = orc.nll_and_grad(COMPOSITE, C2_HP, 1e-2, x, y, reference_distance=True)
    _check_live_oracle(live, _flatten(gref, gnoise), gold, n)
    prog = eng.DeviceProgram.get(COMPOSITE, 1, False, 1)
    plan = eng.Plan([prog], [n], want_grad=True)
    flat = bc.C2_FLAT
    # (1) resident inputs, direct launches on the plan's streams (depth-2 look-ahead, kb = 2 for n >= 6144, fused inverse)
    plan.set_data(0, torch.tensor(x), torch.tensor(y))
    plan.set_hp(0, flat, 1e-2)
    plan.eval(eng.STAGES_LML_GRAD)
    nll, grads, info = plan.results()
    assert info[0] == 0
    _check(nll[0], grads[0], gold, "plan.eval")
    # (2) host buffers: first call captures the CUDA graph, the following ones replay it
    for rep in range(3):
        nll_h, grads_h, info_h = plan.eval_host([flat], [1e-2], [x], [y.reshape(-1)])
        assert info_h[0] == 0
        _check(nll_h[0], grads_h[0], gold, "eval_host #%d" % rep)
    # (3) stage by stage (how bench.py times the stages): same numbers
    for bit in (eng.STAGE_ASSEMBLE, eng.STAGE_POTRF, eng.STAGE_NLL, eng.STAGE_TRTRI, eng.STAGE_LAUUM, eng.STAGE_GRAD):
        plan.eval(bit)
    nll_s, grads_s, info_s = plan.results()
    _check(nll_s[0], grads_s[0], gold, "stage by stage")
    # K^-1 against the oracle's alpha: K^-1 y == alpha (lower triangle mirrored), a full-size property of the inverse
    alpha = plan.buffer(0, eng.BUF_ALPHA).clone()
    Kl = torch.tril(plan.lower_matrix(0, eng.BUF_KINV))
    yv = torch.tensor(y.reshape(-1), device=alpha.device)
    a2 = Kl @ yv + torch.tril(Kl, -1).t() @ yv
    assert float((a2 - alpha).abs().max()) <= 1e-8 * float(alpha.abs().max())


def test_c3_heterogeneous_candidates_match_oracle():
    """16 candidates of bench.py's C3 grammar (seed 2) at the benched n = 2048 in one batched plan: the 16 largest
    programs of the benched draw (the heaviest assembly / gradient work)"""
    eng = _eng()
    n, B = 2048, 16
    order, trees, hps, x, ys = bc.c3_inputs(n, B)
    progs = eng.DeviceProgram.get_many([trees[b] for b in order], 1, False, 1)
    assert all(p.specialised for p in progs), [p.jit_note for p in progs if not p.specialised]
    plan = eng.Plan(progs, [n] * B, want_grad=True)
    nll, grads, info = plan.eval_host([hps[b] for b in order], [1e-2] * B, [x] * B, [y.reshape(-1) for y in ys])
    assert int(np.max(info)) == 0
    for k, b in enumerate(order):
        gold = GOLD["c3"][str(b)]
        _check(nll[k], grads[k], gold, ("candidate", b, trees[b]))
        if k < 4:       # the live oracle of this box on a sample of the candidates
            live, gref, gnoise = orc.nll_and_grad(trees[b], bc.c3_hp_struct(trees[b], hps[b]), 1e-2, x, ys[k],
                                                  reference_distance=True)
            _check_live_oracle(live, _flatten(gref, gnoise), gold, ("candidate", b))


def test_c4_partition_blocks_match_oracle(request):
    """32 blocks x n = 1024 of bench.py's C4 construction (x = arange(N) / N, cut at multiples of 1 / blocks) through
    PartitionedGaussianProcess + blockwise_LL; the partition index lists are compared bit-exactly"""
    from gaussianprocessfundamentals_b200 import compat
    compat.install_as_gpbasics()
    import gpbasics.global_parameters as global_param
    global_param.init(1)
    import gpbasics.KernelBasics.BaseKernels as bk
    import gpbasics.KernelBasics.PartitionOperator as po
    import gpbasics.KernelBasics.PartitioningModel as pm
    import gpbasics.DataHandling.DataInput as di
    import gpbasics.MeanFunctionBasics.BaseMeanFunctions as bmf
    import gpbasics.Statistics.GaussianProcess as gproc
    import gpbasics.Metrics.Auxiliary as met_aux
    import gpbasics.Metrics.Metrics as met
    x, y, ls, nb, n = bc.c4_inputs()
    model = pm.PartitioningModel(pm.PartitioningClass.SELF_SUFFICIENT, [])
    model.init_partitioning([pm.IntervalCriterion(j / nb, (j + 1) / nb) for j in range(nb)])
    idx = model.get_data_record_indices_per_partition(x)
    want_idx = orc.partition_indices([np.logical_and(x[:, 0] >= j / nb, x[:, 0] < (j + 1) / nb).astype(np.float64)
                                      for j in range(nb)])
    for j in range(nb):
        assert np.array_equal(np.asarray(idx[j]), want_idx[j])
        assert np.array_equal(want_idx[j], np.arange(j * n, (j + 1) * n))
    kern = po.PartitionOperator(1, [bk.SquaredExponentialKernel(1) for _ in range(nb)], model)
    hp = [torch.tensor(v, dtype=torch.float64) for v in ls]
    pdi = model.partition_data_input(di.DataInput(x, y, x, y))
    pdi.set_mean_function(bmf.ZeroMeanFunction(1))
    pgp = gproc.PartitionedGaussianProcess(kern, bmf.ZeroMeanFunction(1))
    pgp.set_data_input(pdi)
    metric = met_aux.get_metric_by_type(met.MetricType.blockwise_LL, pgp)
    noise = torch.tensor(1e-2, dtype=torch.float64)
    val = float(metric.get_metric(hp, noise, None))
    grads, gnoise = metric.get_gradients(hp, noise, with_noise=True)
    total, gtot, gn_tot = 0.0, [], 0.0
    for j in range(nb):
        gold = GOLD["c4"]["blocks"][j]
        ref = gold["nll"]
        assert abs(metric.last_block_values[j] - ref) <= LL_RTOL * abs(ref), (j, metric.last_block_values[j], ref)
        total += ref
        gtot.append(gold["grad_ls"])
        gn_tot += gold["grad_noise"]
        if j % 8 == 0:   # the live oracle of this box on a sample of the blocks
            live, _, _ = orc.nll_and_grad(("SE",), [ls[j]], 1e-2, x[j * n:(j + 1) * n], y[j * n:(j + 1) * n],
                                          reference_distance=True)
            assert abs(live - ref) <= LIVE_ORACLE_RTOL * abs(ref), ("live oracle vs frozen", j, live, ref)
    assert abs(val - total) <= LL_RTOL * abs(total)
    got = np.array([float(np.asarray(v).reshape(-1)[0]) for v in grads])
    assert np.max(np.abs(got - np.array(gtot))) <= GRAD_RTOL * np.max(np.abs(gtot))
    assert abs(float(gnoise) - gn_tot) <= GRAD_RTOL * abs(gn_tot)


def test_m16k_likelihood_matches_lapack():
    """n = 16384 (between the two sizes the metric names): the NLL of the device factorisation against LAPACK's
    Cholesky (scipy) of the oracle's covariance matrix - an independent linear-algebra stack at a size where the
    256-wide outer panels and 128 panel steps accumulate"""
    eng = _eng()
    n = 16384
    x, y = bc.c2_inputs(n)
    gold = GOLD["m16k"]
    ref_live, alpha = bc.m16k_lapack(n)       # LAPACK on this box ...
    assert abs(ref_live - gold["nll"]) <= LIVE_ORACLE_RTOL * abs(gold["nll"]), (ref_live, gold["nll"])
    ref = gold["nll"]                         # ... the device is held to the value frozen in the build container
    prog = eng.DeviceProgram.get(COMPOSITE, 1, False, 1)
    plan = eng.Plan([prog], [n], want_grad=False)
    nll, _, info = plan.eval_host([np.array([0.1, 0.1, 0.1, 0.01])], [1e-2], [x], [y.reshape(-1)], stages=eng.STAGES_LML)
    assert info[0] == 0
    assert abs(nll[0] - ref) <= LL_RTOL * abs(ref), (nll[0], ref)
    plan.eval(eng.STAGE_BACKSOLVE)
    a = plan.buffer(0, eng.BUF_ALPHA).cpu().numpy()
    assert np.max(np.abs(a - alpha.reshape(-1))) <= 1e-7 * np.max(np.abs(alpha))
