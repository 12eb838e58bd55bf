"""GPU: the parts of the drop-in surface that had no golden vector in round 1, compared with outputs of the UNMODIFIED
reference (tests/golden/make_golden_r2.py): the rank-3 BatchDataInput likelihood with its gradient, partitioned
prediction (block-rectangular K_s with empty partitions), blockwise BIC / MSE, and a fitter driven by BIC.
Tolerances (north_star): relative <= 1e-10 on the likelihood, <= 1e-8 on gradients."""
import json
import os

import numpy as np
import pytest
import torch

from tests.test_gpu_golden import build, gpb  # noqa: F401  (fixture + kernel builder)

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LL_RTOL, GRAD_RTOL = 1e-10, 1e-8
COMPOSITE = ["MUL", [["ADD", [["SE"], ["PER"]]], ["LIN"]]]


@pytest.fixture(scope="module")
def gold2():
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_r2.npz"))
    return z, json.loads(bytes(z["__meta__"]).decode("utf-8"))


def _hp_list(kern, flat):
    hp, pos = [], 0
    for d in kern.get_hyper_parameter_dimensionalities():
        size = 1 if len(d) == 0 else d[0]
        hp.append(torch.tensor(flat[pos:pos + size]).reshape(d))
        pos += size
    assert pos == len(flat)
    return hp


@pytest.mark.parametrize("name", ["batch3_se", "batch3_composite"])
def test_rank3_batch_likelihood_and_gradient_match_reference(gold2, gpb, name):
    """Metrics/LogLikelihood.py:49,62-63 with Metrics/Metrics.py:152-154"""
    import gpbasics.DataHandling.BatchDataInput as bdi
    z, meta = gold2
    g = gpb
    kern = build(g, json.loads(meta[name]["spec"]))
    hp = _hp_list(kern, z[name + "/hp"])
    noise = torch.tensor(float(z[name + "/noise"]), dtype=torch.float64)
    x, y = z[name + "/x"], z[name + "/y"]
    din = bdi.BatchDataInput(x, y, x, y)
    din.set_mean_function(g.bmf.ZeroMeanFunction(1))
    gp = g.gproc.GaussianProcess(kern, g.bmf.ZeroMeanFunction(1))
    gp.set_data_input(din)
    metric = g.met_aux.get_metric_by_type(g.met.MetricType.LL, gp, g.mht.MatrixApproximations.NONE,
                                          g.mht.NumericalMatrixHandlingType.CHOLESKY_BASED)
    ref = float(z[name + "/nll"][0])
    val = metric.get_metric(hp, noise, None)
    assert tuple(val.shape) == (1, 1)
    assert abs(float(val) - ref) <= LL_RTOL * abs(ref)
    grads, gnoise = metric.get_gradients(hp, noise, with_noise=True)
    got = np.concatenate([np.asarray(v).reshape(-1) for v in grads])
    assert np.max(np.abs(got - z[name + "/grad"])) <= GRAD_RTOL * np.max(np.abs(z[name + "/grad"]))
    assert abs(float(gnoise) - float(z[name + "/grad_noise"][0])) <= GRAD_RTOL * abs(float(z[name + "/grad_noise"][0]))
    # the autograd bridge differentiates the same aggregate
    hp_v = [h.clone().requires_grad_(True) for h in hp]
    out = metric.get_metric(hp_v, noise, None)
    out.backward()
    got2 = np.concatenate([h.grad.numpy().reshape(-1) for h in hp_v])
    assert np.max(np.abs(got2 - z[name + "/grad"])) <= GRAD_RTOL * np.max(np.abs(z[name + "/grad"]))
    # the per-entry mode returns the mean of the true likelihoods (and its gradient is consistent with it)
    metric.reference_batch_aggregate = False
    per = z[name + "/per_entry_nll"]
    v2 = float(metric.get_metric(hp, noise, None))
    assert abs(v2 - np.mean(per)) <= LL_RTOL * abs(np.mean(per))
    g_mean = metric.get_gradients(hp, noise)
    eps = 1e-6
    hp_p = [h.clone() for h in hp]
    hp_p[0] = hp_p[0] + eps
    hp_m = [h.clone() for h in hp]
    hp_m[0] = hp_m[0] - eps
    fd = (float(metric.get_metric(hp_p, noise, None)) - float(metric.get_metric(hp_m, noise, None))) / (2 * eps)
    assert abs(float(np.asarray(g_mean[0]).reshape(-1)[0]) - fd) <= 1e-5 * max(1.0, abs(fd))
    # switching back restores the reference aggregate (the device-side weights follow the flag)
    metric.reference_batch_aggregate = True
    assert abs(float(metric.get_metric(hp, noise, None)) - ref) <= LL_RTOL * abs(ref)
    got3 = np.concatenate([np.asarray(v).reshape(-1) for v in metric.get_gradients(hp, noise)])
    assert np.max(np.abs(got3 - z[name + "/grad"])) <= GRAD_RTOL * np.max(np.abs(z[name + "/grad"]))


def _partition_setup(g, z, meta, name):
    edges = z[name + "/edges"]
    model = g.pm.PartitioningModel(g.pm.PartitioningClass.SELF_SUFFICIENT, [])
    model.init_partitioning([g.pm.IntervalCriterion(edges[i], edges[i + 1]) for i in range(len(edges) - 1)])
    specs = json.loads(meta[name]["specs"])
    kern = g.po.PartitionOperator(1, [build(g, s) for s in specs], model)
    return model, kern, _hp_list(kern, z[name + "/hp"])


@pytest.mark.parametrize("name", ["ppred_full", "ppred_dead"])
def test_partitioned_rectangular_K_s_matches_reference(gold2, gpb, name):
    """Auxiliary/NonSquareBlockMatrices.py:8-103 through PartitionOperator.get_tf_tensor (PartitionOperator.py:24-83):
    block-rectangular train x test covariance, incl. partitions without test points (dead rows) or training points
    (dead columns); index bookkeeping bit-exact"""
    z, meta = gold2
    g = gpb
    model, kern, hp = _partition_setup(g, z, meta, name)
    x, xt = z[name + "/x"], z[name + "/xt"]
    idx = model.get_data_record_indices_per_partition(x)
    idx_t = model.get_data_record_indices_per_partition(xt)
    for i, (a, b) in enumerate(zip(idx, idx_t)):
        assert np.array_equal(np.asarray(a), z[name + "/idx%d" % i])
        assert np.array_equal(np.asarray(b), z[name + "/idxt%d" % i])
    Ks = kern.get_tf_tensor(hp, x, xt).cpu().numpy()
    ref = z[name + "/K_s"]
    assert Ks.shape == ref.shape
    assert np.max(np.abs(Ks - ref)) <= 1e-13 * np.max(np.abs(ref))
    assert np.array_equal(Ks == 0.0, ref == 0.0)        # the zero pattern (dead rows / columns, off-diagonal blocks)


def test_partitioned_prediction_and_blockwise_metrics_match_reference(gold2, gpb):
    """Statistics/GaussianProcess.py:42-85 (partitioned predict), Metrics/BayesianInformationCriterion.py:43-63,
    Metrics/MeanSquaredError.py:45-81"""
    z, meta = gold2
    g = gpb
    name = "ppred_full"
    m = meta[name]
    assert m["predict_ok"] and m["bic_ok"] and m["mse_ok"] and m["ll_ok"]
    model, kern, hp = _partition_setup(g, z, meta, name)
    base = g.di.DataInput(z[name + "/x"], z[name + "/y"], z[name + "/xt"], z[name + "/yt"])
    pdi = model.partition_data_input(base)
    pdi.set_mean_function(g.bmf.ZeroMeanFunction(1))
    assert np.array_equal(pdi.data_x_test.cpu().numpy(), z[name + "/xt_reordered"])
    pgp = g.gproc.PartitionedGaussianProcess(kern, g.bmf.ZeroMeanFunction(1))
    pgp.set_data_input(pdi)
    noise = torch.tensor(1e-2, dtype=torch.float64)
    A = (g.mht.MatrixApproximations.NONE, g.mht.NumericalMatrixHandlingType.CHOLESKY_BASED)
    for key, mt, tol in [("ll", g.met.MetricType.blockwise_LL, LL_RTOL), ("bic", g.met.MetricType.blockwise_BIC, LL_RTOL),
                         ("mse", g.met.MetricType.blockwise_MSE, 1e-9)]:
        metric = g.met_aux.get_metric_by_type(mt, pgp, *A)
        val = float(metric.get_metric(hp, noise, None))
        ref = float(z[name + "/blockwise_" + key][0])
        assert abs(val - ref) <= tol * abs(ref), key
    total, mean_mu, post = pgp.predict(hp, None, noise)
    scale = np.max(np.abs(z[name + "/predict_post_mu"]))
    assert np.max(np.abs(post.cpu().numpy().reshape(-1) - z[name + "/predict_post_mu"])) <= 1e-9 * scale
    assert np.max(np.abs(total.cpu().numpy().reshape(-1) - z[name + "/predict_total"])) <= 1e-9 * scale
    # gradient of the blockwise BIC = 2 x gradient of the blockwise likelihood (the penalty is constant)
    bic = g.met_aux.get_metric_by_type(g.met.MetricType.blockwise_BIC, pgp, *A)
    bll = g.met_aux.get_metric_by_type(g.met.MetricType.blockwise_LL, pgp, *A)
    gb = np.concatenate([np.asarray(v).reshape(-1) for v in bic.get_gradients(hp, noise)])
    gl = np.concatenate([np.asarray(v).reshape(-1) for v in bll.get_gradients(hp, noise)])
    assert np.max(np.abs(gb - 2 * gl)) <= 1e-12 * np.max(np.abs(gl))


def test_fitters_accept_bic_and_refuse_mse(gpb):
    """every fitter built with MetricType.BIC used to die in _objective; MSE has no device gradient"""
    g = gpb
    rng = np.random.default_rng(5)
    x = np.linspace(0.25, 2.25, 160)[:, None]
    y = np.sin(5 * x) + 0.1 * rng.standard_normal(x.shape)
    din = g.di.DataInput(x, y, x, y)
    din.set_mean_function(g.bmf.ZeroMeanFunction(1))
    A = (g.mht.MatrixApproximations.NONE, g.mht.NumericalMatrixHandlingType.CHOLESKY_BASED)
    g.global_param.p_cov_matrix_jitter = torch.tensor(1e-2, dtype=torch.float64)
    try:
        gp = g.gproc.GaussianProcess(build(g, COMPOSITE), g.bmf.ZeroMeanFunction(1))
        f_bic = g.fitter.VariationalSgdFitter(din, gp, g.met.MetricType.BIC, False, *A)
        pre, post, hps, nz, _ = f_bic.fit()
        gp2 = g.gproc.GaussianProcess(build(g, COMPOSITE), g.bmf.ZeroMeanFunction(1))
        f_ll = g.fitter.VariationalSgdFitter(din, gp2, g.met.MetricType.LL, False, *A)
        pre_ll, _, _, _, _ = f_ll.fit()
        penalty = 4 * np.log(160)
        assert abs(float(pre) - (2 * float(pre_ll) + penalty)) <= 1e-10 * abs(float(pre))
        gb = np.concatenate([np.asarray(v).reshape(-1) for v in f_bic.last_gradients])
        gl = np.concatenate([np.asarray(v).reshape(-1) for v in f_ll.last_gradients])
        assert np.max(np.abs(gb - 2 * gl)) <= 1e-10 * np.max(np.abs(gl))
        gp3 = g.gproc.GaussianProcess(build(g, COMPOSITE), g.bmf.ZeroMeanFunction(1))
        adam = g.fitter.AdamFitter(din, gp3, g.met.MetricType.BIC, False, *A, steps=5, learning_rate=1e-3)
        pre_a, post_a, _, _, _ = adam.fit()
        assert float(post_a) < float(pre_a)
        with pytest.raises(NotImplementedError):
            g.fitter.VariationalSgdFitter(din, g.gproc.GaussianProcess(build(g, COMPOSITE), g.bmf.ZeroMeanFunction(1)),
                                          g.met.MetricType.MSE, False, *A)
    finally:
        g.global_param.p_cov_matrix_jitter = torch.tensor(1e-8, dtype=torch.float64)


def test_new_mean_function_reaches_the_device(gpb):
    """the plan kept the targets it was built with; the reference re-reads get_detrended_y_train() on every
    get_metric (Metrics/LogLikelihood.py:35)"""
    g = gpb
    rng = np.random.default_rng(6)
    x = np.linspace(0.0, 1.0, 150)[:, None]
    y = 3.0 + np.sin(7 * x) + 0.1 * rng.standard_normal(x.shape)
    kern = build(g, ["SE"])
    hp = [torch.tensor(0.2, dtype=torch.float64)]
    noise = torch.tensor(1e-2, dtype=torch.float64)
    din = g.di.DataInput(x, y, x, y)
    din.set_mean_function(g.bmf.ZeroMeanFunction(1))
    gp = g.gproc.GaussianProcess(kern, g.bmf.ZeroMeanFunction(1))
    gp.set_data_input(din)
    metric = g.met_aux.get_metric_by_type(g.met.MetricType.LL, gp)
    v0 = float(metric.get_metric(hp, noise, None))
    # same object, new mean function: y_detrended changes by the constant
    const = g.bmf.ConstantMeanFunction(1)
    const.set_last_hyper_parameter([torch.tensor(3.0, dtype=torch.float64)])
    din.set_mean_function(const)
    v1 = float(metric.get_metric(hp, noise, None))
    din2 = g.di.DataInput(x, y - 3.0, x, y - 3.0)
    din2.set_mean_function(g.bmf.ZeroMeanFunction(1))
    gp2 = g.gproc.GaussianProcess(build(g, ["SE"]), g.bmf.ZeroMeanFunction(1))
    gp2.set_data_input(din2)
    v2 = float(g.met_aux.get_metric_by_type(g.met.MetricType.LL, gp2).get_metric(hp, noise, None))
    assert abs(v1 - v2) <= 1e-10 * abs(v2)
    assert abs(v0 - v1) > 1e-3 * abs(v1)
