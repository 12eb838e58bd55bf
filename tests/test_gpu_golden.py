"""GPU: the drop-in operator surface (kernel objects -> DataInput -> GaussianProcess -> Metric -> Fitter) evaluated by the
CUDA path, compared with the golden vectors of the UNMODIFIED reference (tests/golden/make_golden.py).
Tolerances (north_star): relative <= 1e-10 on the likelihood, <= 1e-8 on gradients."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
LL_RTOL, GRAD_RTOL = 1e-10, 1e-8


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    return z, json.loads(bytes(z["__meta__"]).decode("utf-8"))


@pytest.fixture(scope="module")
def gpb():
    """the reference's import names, resolved to the B200 implementation"""
    from gaussianprocessfundamentals_b200 import compat
    compat.install_as_gpbasics()
    import gpbasics.global_parameters as global_param
    global_param.init(1)
    import gpbasics.KernelBasics.BaseKernels as bk
    import gpbasics.KernelBasics.Operators as op
    import gpbasics.KernelBasics.PartitionOperator as po
    import gpbasics.KernelBasics.PartitioningModel as pm
    import gpbasics.DataHandling.DataInput as di
    import gpbasics.MeanFunctionBasics.BaseMeanFunctions as bmf
    import gpbasics.Statistics.GaussianProcess as gproc
    import gpbasics.Metrics.Auxiliary as met_aux
    import gpbasics.Metrics.Metrics as met
    import gpbasics.Metrics.MatrixHandlingTypes as mht
    import gpbasics.Optimizer.Fitter as fitter

    class NS:
        pass
    ns = NS()
    ns.__dict__.update(locals())
    return ns


def build(g, spec, d=1):
    leaf = {"SE": g.bk.SquaredExponentialKernel, "PER": g.bk.PeriodicKernel, "LIN": g.bk.LinearKernel,
            "MAT32": g.bk.MaternKernel3_2, "MAT52": g.bk.MaternKernel5_2, "WN": g.bk.WhiteNoiseKernel}
    kind = spec[0]
    if kind in leaf:
        return leaf[kind](d)
    children = [build(g, c, d) for c in spec[1]]
    if kind == "ADD":
        return g.op.AdditionOperator(d, children)
    if kind == "MUL":
        return g.op.MultiplicationOperator(d, children)
    if kind == "CP":
        return g.op.ChangePointOperator(d, children, [torch.tensor(c, dtype=torch.float64) for c in spec[2]])
    raise ValueError(kind)


def _holistic(meta):
    return [k for k, v in meta.items() if isinstance(v, dict) and v.get("kind") == "holistic"]


def test_holistic_cases_match_reference(gold, gpb):
    z, meta = gold
    g = gpb
    for name in _holistic(meta):
        m = meta[name]
        g.global_param.p_scaled_base_kernel = m["scaled"]
        g.global_param.p_cp_operator_type = g.global_param.ChangePointOperatorType(m["cp_mode"])
        try:
            kern = build(g, json.loads(m["spec"]))
            n_hp = int(z[name + "/n_hp"])
            hp = [torch.tensor(z[name + "/hp%d" % i]) for i in range(n_hp)]
            x, y = z[name + "/x"], z[name + "/y"]
            din = g.di.DataInput(x, y, x, y)
            din.set_mean_function(g.bmf.ZeroMeanFunction(1))
            gp = g.gproc.GaussianProcess(kern, g.bmf.ZeroMeanFunction(1))
            gp.set_data_input(din)
            metric = g.met_aux.get_metric_by_type(g.met.MetricType.LL, gp, g.mht.MatrixApproximations.NONE,
                                                  g.mht.NumericalMatrixHandlingType.CHOLESKY_BASED)
            raw = float(z[name + "/noise"])
            noise = torch.tensor(abs(raw) if m["optimize_noise"] else raw, dtype=torch.float64)
            val = metric.get_metric(hp, noise, None)
            assert tuple(val.shape) == (1, 1)
            ref = float(z[name + "/nll"][0])
            ill = name == "se_default_jitter_n100"   # jitter 1e-8: numerically singular, tolerance waived
            assert abs(float(val) - ref) <= (1e-5 if ill else LL_RTOL) * abs(ref), (name, float(val), ref)
            grads, gnoise = metric.get_gradients(hp, noise, with_noise=True)
            gref = [z[name + "/grad%d" % i] for i in range(n_hp)]
            scale = max(np.max(np.abs(np.concatenate([np.reshape(v, -1) for v in gref]))), 1e-300)
            for i, (a, b) in enumerate(zip(grads, gref)):
                assert tuple(a.shape) == tuple(np.shape(b)), (name, i)
                assert np.max(np.abs(a.numpy() - b)) <= (1e-3 if ill else GRAD_RTOL) * scale, (name, i, a, b)
            gn_ref = float(z[name + "/grad_noise"])
            gn = float(gnoise) * (np.sign(raw) if m["optimize_noise"] else 1.0)
            assert abs(gn - gn_ref) <= (1e-3 if ill else GRAD_RTOL) * max(abs(gn_ref), scale), name
        finally:
            g.global_param.p_scaled_base_kernel = False
            g.global_param.p_cp_operator_type = g.global_param.ChangePointOperatorType.INDICATOR


def test_covariance_matrix_getters_match_reference(gold, gpb):
    z, meta = gold
    g = gpb
    for name in ("se_n64_mats", "composite_n96_mats"):
        kern = build(g, json.loads(meta[name]["spec"]))
        n_hp = int(z[name + "/n_hp"])
        hp = [torch.tensor(z[name + "/hp%d" % i]) for i in range(n_hp)]
        x, y = z[name + "/x"], z[name + "/y"]
        din = g.di.DataInput(x, y, x, y)
        din.set_mean_function(g.bmf.ZeroMeanFunction(1))
        gp = g.gproc.GaussianProcess(kern, g.bmf.ZeroMeanFunction(1))
        gp.set_data_input(din)
        cov = gp.covariance_matrix
        noise = torch.tensor(float(z[name + "/noise"]), dtype=torch.float64)
        K = cov.get_K(hp).cpu().numpy()
        assert np.max(np.abs(K - z[name + "/K"])) <= 1e-13 * np.max(np.abs(z[name + "/K"]))
        Kn = cov.get_K_noised(hp, noise).cpu().numpy()
        assert np.max(np.abs(Kn - (z[name + "/K"] + float(noise) * np.eye(len(x))))) <= 1e-13
        L = cov.get_L_K(hp, noise).cpu().numpy()
        assert np.allclose(np.triu(L, 1), 0.0)
        assert np.max(np.abs(L @ L.T - Kn)) <= 1e-13 * np.max(np.abs(Kn))
        assert np.max(np.abs(L - z[name + "/L"])) <= 1e-6 * np.max(np.abs(z[name + "/L"]))
        a = cov.get_L_alpha(hp, noise).cpu().numpy()
        assert a.shape == z[name + "/alpha"].shape
        assert np.max(np.abs(a - z[name + "/alpha"])) <= 1e-6 * np.max(np.abs(z[name + "/alpha"]))
        cov.reset()
        Kinv = cov.get_K_inv(hp, noise).cpu().numpy()
        assert np.max(np.abs(Kinv @ Kn - np.eye(len(x)))) <= 1e-6
        Linv = cov.get_L_inv_K(hp, noise).cpu().numpy()
        assert np.max(np.abs(Linv @ z[name + "/L"] - np.eye(len(x)))) <= 1e-6


def test_blockwise_gp_matches_reference_blocks_and_holistic(gold, gpb):
    z, meta = gold
    g = gpb
    x, y, cps = z["blockwise/x"], z["blockwise/y"], z["blockwise/cps"]
    specs = json.loads(meta["blockwise"]["specs"])
    children = [build(g, s) for s in specs]
    cpk = g.op.ChangePointOperator(1, children, [torch.tensor(c, dtype=torch.float64) for c in cps])
    flat = z["blockwise/hp_children"]
    hp_children = []
    pos = 0
    for d in cpk.get_hyper_parameter_dimensionalities()[len(cps):]:
        size = 1 if len(d) == 0 else d[0]
        hp_children.append(torch.tensor(flat[pos:pos + size]).reshape(d))
        pos += size
    full = [torch.tensor(c, dtype=torch.float64) for c in cps] + hp_children
    bdi = g.di.BlockwiseDataInput(x, y, x, y, [torch.tensor(c, dtype=torch.float64) for c in cps])
    bdi.set_mean_function(g.bmf.ZeroMeanFunction(1))
    bgp = g.gproc.BlockwiseGaussianProcess(cpk, g.bmf.ZeroMeanFunction(1))
    bgp.set_data_input(bdi)
    metric = g.met_aux.get_metric_by_type(g.met.MetricType.blockwise_LL, bgp)
    noise = torch.tensor(1e-2, dtype=torch.float64)
    total = float(metric.get_metric(full, noise, None))
    ref_blocks = z["blockwise/block_nll"]
    for v, r in zip(metric.last_block_values, ref_blocks):
        if v is None:
            assert np.isnan(r)
        else:
            assert abs(v - r) <= LL_RTOL * abs(r)
    assert abs(total - np.nansum(ref_blocks)) <= LL_RTOL * abs(np.nansum(ref_blocks))
    # holistic evaluation of the same change-point kernel (one dense n x n Cholesky) agrees with the block sum
    din = g.di.DataInput(x, y, x, y)
    din.set_mean_function(g.bmf.ZeroMeanFunction(1))
    gp_h = g.gproc.GaussianProcess(cpk.deepcopy(), g.bmf.ZeroMeanFunction(1))
    gp_h.set_data_input(din)
    mh = g.met_aux.get_metric_by_type(g.met.MetricType.LL, gp_h)
    hol = float(mh.get_metric(full, noise, None))
    assert abs(hol - float(z["blockwise/holistic_nll"][0])) <= LL_RTOL * abs(hol)
    assert abs(hol - total) <= 1e-9 * abs(hol)


def test_partitioned_gp_matches_reference(gold, gpb):
    z, meta = gold
    g = gpb
    assert meta["partition"]["ok"]
    x, y, edges = z["partition/x"], z["partition/y"], z["partition/edges"]
    model = g.pm.PartitioningModel(g.pm.PartitioningClass.SELF_SUFFICIENT, [])
    model.init_partitioning([g.pm.IntervalCriterion(edges[i], edges[i + 1]) for i in range(4)])
    specs = json.loads(meta["partition"]["specs"])
    kern = g.po.PartitionOperator(1, [build(g, s) for s in specs], model)
    flat = z["partition/hp"]
    hp, pos = [], 0
    for d in kern.get_hyper_parameter_dimensionalities():
        size = 1 if len(d) == 0 else d[0]
        hp.append(torch.tensor(flat[pos:pos + size]).reshape(d))
        pos += size
    pdi = model.partition_data_input(g.di.DataInput(x, y, x, y))
    pdi.set_mean_function(g.bmf.ZeroMeanFunction(1))
    pgp = g.gproc.PartitionedGaussianProcess(kern, g.bmf.ZeroMeanFunction(1))
    pgp.set_data_input(pdi)
    metric = g.met_aux.get_metric_by_type(g.met.MetricType.blockwise_LL, pgp)
    noise = torch.tensor(1e-2, dtype=torch.float64)
    val = float(metric.get_metric(hp, noise, None))
    ref = float(z["partition/blockwise_nll"][0])
    assert abs(val - ref) <= LL_RTOL * abs(ref)
    grads = metric.get_gradients(hp, noise)
    gflat = np.concatenate([np.asarray(v).reshape(-1) for v in grads])
    assert np.max(np.abs(gflat - z["partition/grad"])) <= GRAD_RTOL * np.max(np.abs(z["partition/grad"]))
    Kd = kern.get_tf_tensor(hp, x, x).cpu().numpy()
    assert np.max(np.abs(Kd - z["partition/K_dense"])) <= 1e-13 * np.max(np.abs(z["partition/K_dense"]))


def test_fitter_one_step_matches_reference(gold, gpb):
    z, meta = gold
    g = gpb
    x, y = z["fit/x"], z["fit/y"]
    kern = build(g, ["MUL", [["ADD", [["SE"], ["PER"]]], ["LIN"]]])
    din = g.di.DataInput(x, y, x, y)
    din.set_mean_function(g.bmf.ZeroMeanFunction(1))
    gp = g.gproc.GaussianProcess(kern, g.bmf.ZeroMeanFunction(1))
    g.global_param.p_cov_matrix_jitter = torch.tensor(1e-2, dtype=torch.float64)
    try:
        f = g.fitter.VariationalSgdFitter(din, gp, g.met.MetricType.LL, False, g.mht.MatrixApproximations.NONE,
                                          g.mht.NumericalMatrixHandlingType.CHOLESKY_BASED)
        pre, post, hps, nz, idx = f.fit()
    finally:
        g.global_param.p_cov_matrix_jitter = torch.tensor(1e-8, dtype=torch.float64)
    assert abs(float(pre) - float(z["fit/pre"][0])) <= LL_RTOL * abs(float(z["fit/pre"][0]))
    got = np.concatenate([np.asarray(v).reshape(-1) for v in f.last_gradients])
    assert np.max(np.abs(got - z["fit/grads"])) <= GRAD_RTOL * np.max(np.abs(z["fit/grads"]))
    assert abs(float(nz) - float(z["fit/noise"])) == 0.0 and idx is None
    # the update itself is the optimiser's business; the step is small, so post ~ pre
    assert abs(float(post) - float(z["fit/post"][0])) <= 1e-4 * abs(float(z["fit/post"][0]))
    assert len(hps) == 4


def test_autograd_bridge_and_adam_fit_loop(gpb):
    g = gpb
    n = 400
    rng = np.random.default_rng(5)
    x = np.linspace(0, 1, n)[:, None]
    y = np.sin(10 * x) + 0.1 * rng.standard_normal((n, 1))
    kern = build(g, ["ADD", [["SE"], ["LIN"]]])
    din = g.di.DataInput(x, y, x, y)
    din.set_mean_function(g.bmf.ZeroMeanFunction(1))
    gp = g.gproc.GaussianProcess(kern, g.bmf.ZeroMeanFunction(1))
    gp.set_data_input(din)
    metric = g.met_aux.get_metric_by_type(g.met.MetricType.LL, gp)
    hp = [torch.tensor(0.3, dtype=torch.float64, requires_grad=True),
          torch.tensor([0.1], dtype=torch.float64, requires_grad=True)]
    noise = torch.tensor(0.05, dtype=torch.float64, requires_grad=True)
    val = metric.get_metric(hp, noise, None)
    val.sum().backward()
    ref = metric.get_gradients([h.detach() for h in hp], noise.detach(), with_noise=True)
    assert torch.allclose(hp[0].grad, ref[0][0]) and torch.allclose(hp[1].grad, ref[0][1])
    assert torch.allclose(noise.grad, ref[1])
    g.global_param.p_cov_matrix_jitter = torch.tensor(0.05, dtype=torch.float64)
    try:
        f = g.fitter.AdamFitter(din, g.gproc.GaussianProcess(kern, g.bmf.ZeroMeanFunction(1)), g.met.MetricType.LL, False,
                                g.mht.MatrixApproximations.NONE, g.mht.NumericalMatrixHandlingType.CHOLESKY_BASED,
                                steps=25, learning_rate=0.02)
        pre, post, hps, nz, _ = f.fit()
    finally:
        g.global_param.p_cov_matrix_jitter = torch.tensor(1e-8, dtype=torch.float64)
    assert float(post) < float(pre)
    # quasi-Newton loop on the same problem: must reach a lower NLL than 25 Adam steps, in few evaluations
    g.global_param.p_cov_matrix_jitter = torch.tensor(0.05, dtype=torch.float64)
    try:
        f2 = g.fitter.LbfgsFitter(din, g.gproc.GaussianProcess(kern, g.bmf.ZeroMeanFunction(1)), g.met.MetricType.LL, False,
                                  g.mht.MatrixApproximations.NONE, g.mht.NumericalMatrixHandlingType.CHOLESKY_BASED,
                                  max_evaluations=40)
        pre2, post2, hps2, nz2, _ = f2.fit()
    finally:
        g.global_param.p_cov_matrix_jitter = torch.tensor(1e-8, dtype=torch.float64)
    assert float(post2) <= float(post) + 1e-9 and len(f2.history) <= 41
    gfin = metric.get_gradients([torch.as_tensor(h) for h in hps2], torch.tensor(0.05, dtype=torch.float64))
    assert max(float(torch.as_tensor(t).abs().max()) for t in gfin) <= 1e-2 * max(1.0, abs(float(post2)))


def test_distances_and_prediction(gpb):
    g = gpb
    import gpbasics.Auxiliary.Distances as dist
    rng = np.random.default_rng(2)
    a, b = rng.uniform(0, 1, (70, 3)), rng.uniform(0, 1, (45, 3))
    d2 = dist.euclidian_distance(a, b).cpu().numpy()
    d1 = dist.manhattan_distance(a, b).cpu().numpy()
    assert np.max(np.abs(d2 - np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)))) <= 1e-15
    assert np.max(np.abs(d1 - np.abs(a[:, None, :] - b[None, :, :]).sum(-1))) <= 1e-15
    n = 300
    x = np.linspace(0, 1, n)[:, None]
    y = np.sin(8 * x) + 0.05 * rng.standard_normal((n, 1))
    xt = np.linspace(0.05, 0.95, 77)[:, None]
    yt = np.sin(8 * xt)
    kern = build(g, ["SE"])
    din = g.di.DataInput(x, y, xt, yt)
    din.set_mean_function(g.bmf.ZeroMeanFunction(1))
    gp = g.gproc.GaussianProcess(kern, g.bmf.ZeroMeanFunction(1))
    gp.set_data_input(din)
    hp = [torch.tensor(0.15, dtype=torch.float64)]
    full, mean, post = gp.predict(hp, None, torch.tensor(1e-2, dtype=torch.float64))
    K = np.exp(-0.5 * (x - x.T) ** 2 / 0.15 ** 2) + 1e-2 * np.eye(n)
    Ks = np.exp(-0.5 * (x - xt.T) ** 2 / 0.15 ** 2)
    want = Ks.T @ np.linalg.solve(K, y)
    assert np.max(np.abs(full.cpu().numpy().reshape(-1) - want.reshape(-1))) <= 1e-8
    var = gp.aux.get_posterior_var(hp, torch.tensor(1e-2, dtype=torch.float64)).cpu().numpy()
    Kss = np.exp(-0.5 * (xt - xt.T) ** 2 / 0.15 ** 2)
    assert np.max(np.abs(var - (Kss - Ks.T @ np.linalg.solve(K, Ks)))) <= 1e-7


# ---- the callers right after the likelihood path: prediction, MSE, BIC (tests/golden/make_golden_predict.py) -------
def test_prediction_mse_bic_match_reference(gpb):
    g = gpb
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_predict.npz"))
    meta = json.loads(bytes(z["__meta__"]).decode("utf-8"))
    A = (g.mht.MatrixApproximations.NONE, g.mht.NumericalMatrixHandlingType.CHOLESKY_BASED)
    for name, m in meta.items():
        kern = build(g, json.loads(m["spec"]))
        flat, hp, pos = z[name + "/hp"], [], 0
        for d in kern.get_hyper_parameter_dimensionalities():
            size = 1 if len(d) == 0 else d[0]
            hp.append(torch.tensor(flat[pos:pos + size]).reshape(d))
            pos += size
        noise = torch.tensor(float(z[name + "/noise"]), dtype=torch.float64)
        din = g.di.DataInput(z[name + "/x"], z[name + "/y"], z[name + "/xt"], z[name + "/yt"])
        din.set_mean_function(g.bmf.ZeroMeanFunction(1))
        gp = g.gproc.GaussianProcess(kern, g.bmf.ZeroMeanFunction(1))
        gp.set_data_input(din)
        mse = g.met_aux.get_metric_by_type(g.met.MetricType.MSE, gp, *A)
        bic = g.met_aux.get_metric_by_type(g.met.MetricType.BIC, gp, *A)
        v_mse, v_bic = float(mse.get_metric(hp, noise, None)), float(bic.get_metric(hp, noise, None))
        assert abs(v_mse - float(z[name + "/mse"][0])) <= 1e-9 * abs(float(z[name + "/mse"][0])), name
        assert abs(v_bic - float(z[name + "/bic"][0])) <= LL_RTOL * abs(float(z[name + "/bic"][0])), name
        gp.aux.reset(); gp.covariance_matrix.reset()
        mu = gp.aux.get_posterior_mu(hp, noise).cpu().numpy().reshape(-1)
        scale = np.max(np.abs(z[name + "/post_mu"]))
        assert np.max(np.abs(mu - z[name + "/post_mu"])) <= 1e-9 * scale, name
        Ks = gp.covariance_matrix.get_K_s(hp).cpu().numpy()
        assert np.max(np.abs(Ks - z[name + "/K_s"])) <= 1e-13 * np.max(np.abs(z[name + "/K_s"])), name
        var = gp.aux.get_posterior_var(hp, noise).cpu().numpy()
        # the posterior covariance is a difference of O(1) terms: absolute tolerance relative to the prior scale
        assert np.max(np.abs(var - z[name + "/post_var"])) <= 1e-9 * max(1.0, np.max(np.abs(Ks))), name
        total, mean_mu, post = gp.predict(hp, None, noise)
        assert np.max(np.abs(total.cpu().numpy().reshape(-1) - z[name + "/predict_total"])) <= 1e-9 * scale, name


def test_candidate_batch_matches_one_at_a_time(gpb):
    """search.CandidateBatch (BASELINE config 3): all candidates in one batched plan = the reference's one metric per
    candidate"""
    g = gpb
    from gaussianprocessfundamentals_b200 import search
    rng = np.random.default_rng(33)
    n = 257
    x = np.sort(rng.uniform(0, 1, (n, 1)), axis=0)
    y = np.sin(6 * x) + 0.5 * x + 0.1 * rng.standard_normal((n, 1))
    specs = [["SE"], ["PER"], ["ADD", [["SE"], ["LIN"]]], ["MUL", [["SE"], ["PER"]]],
             ["ADD", [["MUL", [["SE"], ["LIN"]]], ["PER"]]], ["MAT32"], ["MUL", [["MAT52"], ["LIN"]]]]
    kernels = [build(g, s) for s in specs]
    din = g.di.DataInput(x, y, x, y)
    din.set_mean_function(g.bmf.ZeroMeanFunction(1))
    hps = []
    for k in kernels:
        hp = []
        for d in k.get_hyper_parameter_dimensionalities():
            size = 1 if len(d) == 0 else d[0]
            hp.append(torch.tensor(rng.uniform(0.3, 1.2, size=size)).reshape(d))
        hps.append(hp)
    noise = torch.tensor(1e-2, dtype=torch.float64)
    batch = search.CandidateBatch(kernels, din)
    nll, grads, gnoise = batch.evaluate(hps, noise)
    for i, k in enumerate(kernels):
        gp = g.gproc.GaussianProcess(k, g.bmf.ZeroMeanFunction(1))
        gp.set_data_input(din)
        metric = g.met_aux.get_metric_by_type(g.met.MetricType.LL, gp)
        want = float(metric.get_metric(hps[i], noise, None))
        wg, wn = metric.get_gradients(hps[i], noise, with_noise=True)
        assert abs(nll[i] - want) <= LL_RTOL * abs(want), i
        a = np.concatenate([np.asarray(t).reshape(-1) for t in grads[i]])
        b = np.concatenate([np.asarray(t).reshape(-1) for t in wg])
        assert np.max(np.abs(a - b)) <= GRAD_RTOL * np.max(np.abs(b)), i
        assert abs(gnoise[i] - float(wn)) <= GRAD_RTOL * abs(float(wn)), i
    assert batch.best(hps, noise) == int(np.argmin(nll))
