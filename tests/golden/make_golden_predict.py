"""Generates tests/golden/reference_golden_predict.npz: posterior mean / covariance, MSE and BIC of the UNMODIFIED
reference (/root/reference/main/gpbasics on top of oracle/tf_shim; see make_golden.py for what that means) on small
train / test splits - the callers right after the likelihood path.  Run from the repository
root:    python tests/golden/make_golden_predict.py
The file it writes is committed; tests/test_gpu_golden.py compares the CUDA path with it."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, "/root/reference/main")

import tensorflow as tf  # noqa: E402  (the shim)
import gpbasics.global_parameters as global_param  # noqa: E402

global_param.init(4)

import gpbasics.KernelBasics.BaseKernels as bk  # noqa: E402
import gpbasics.KernelBasics.Operators as op  # noqa: E402
import gpbasics.DataHandling.DataInput as di  # noqa: E402
import gpbasics.MeanFunctionBasics.BaseMeanFunctions as bmf  # noqa: E402
import gpbasics.Statistics.GaussianProcess as gproc  # noqa: E402
import gpbasics.Metrics.Auxiliary as met_aux  # noqa: E402
import gpbasics.Metrics.Metrics as met  # noqa: E402
import gpbasics.Metrics.MatrixHandlingTypes as mht  # noqa: E402

OUT, META = {}, {}
LEAF = {"SE": bk.SquaredExponentialKernel, "PER": bk.PeriodicKernel, "LIN": bk.LinearKernel,
        "MAT32": bk.MaternKernel3_2, "MAT52": bk.MaternKernel5_2}


def build(spec):
    if spec[0] in LEAF:
        return LEAF[spec[0]](1)
    cls = op.AdditionOperator if spec[0] == "ADD" else op.MultiplicationOperator
    return cls(1, [build(c) for c in spec[1]])


def tf_hp(values):
    return [tf.Variable(np.asarray(v, dtype=np.float64), dtype=tf.float64) for v in values]


def case(name, spec, hp_values, n, n_test, noise, seed, blockwise_cps=None):
    rng = np.random.default_rng(seed)
    x_all = np.sort(rng.uniform(0.0, 1.0, size=(n + n_test, 1)), axis=0)
    y_all = np.sin(9 * x_all) + 0.3 * x_all + 0.1 * rng.standard_normal(x_all.shape)
    test_idx = np.sort(rng.choice(n + n_test, size=n_test, replace=False))
    mask = np.ones(n + n_test, dtype=bool)
    mask[test_idx] = False
    x, y, xt, yt = x_all[mask], y_all[mask], x_all[~mask], y_all[~mask]
    kernel = build(spec)
    hp = tf_hp(hp_values)
    nz = tf.constant(noise, dtype=tf.float64)
    din = di.DataInput(x, y, xt, yt)
    din.set_mean_function(bmf.ZeroMeanFunction(1))
    gp = gproc.GaussianProcess(kernel, bmf.ZeroMeanFunction(1))
    gp.set_data_input(din)
    A = (mht.MatrixApproximations.NONE, mht.NumericalMatrixHandlingType.CHOLESKY_BASED)
    mse = met_aux.get_metric_by_type(met.MetricType.MSE, gp, *A)
    bic = met_aux.get_metric_by_type(met.MetricType.BIC, gp, *A)
    OUT[name + "/x"], OUT[name + "/y"], OUT[name + "/xt"], OUT[name + "/yt"] = x, y, xt, yt
    OUT[name + "/noise"] = np.float64(noise)
    OUT[name + "/hp"] = np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in hp_values])
    OUT[name + "/mse"] = np.asarray(mse.get_metric(hp, nz, None).numpy()).reshape(-1)
    OUT[name + "/bic"] = np.asarray(bic.get_metric(hp, nz, None).numpy()).reshape(-1)
    gp.aux.reset(); gp.covariance_matrix.reset()
    OUT[name + "/post_mu"] = gp.aux.get_posterior_mu(hp, nz).numpy().reshape(-1)
    OUT[name + "/post_var"] = gp.aux.get_posterior_var(hp, nz).numpy()
    OUT[name + "/K_s"] = gp.covariance_matrix.get_K_s(hp).numpy()
    total, mean_mu, post_mu = gp.predict(hp, None, nz)
    OUT[name + "/predict_total"] = np.asarray(total.numpy()).reshape(-1)
    META[name] = {"spec": json.dumps(spec), "n": int(x.shape[0]), "n_test": int(xt.shape[0])}


case("se", ("SE",), [0.15], 160, 40, 1e-2, 21)
case("composite", ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)]), [0.2, 0.4, 0.3, [0.1]], 220, 55, 1e-2, 22)
case("matern", ("ADD", [("MAT32",), ("MAT52",)]), [0.3, 0.2], 130, 31, 5e-2, 23)

OUT["__meta__"] = np.frombuffer(json.dumps(META).encode("utf-8"), dtype=np.uint8)
np.savez_compressed(os.path.join(HERE, "reference_golden_predict.npz"), **OUT)
print("wrote reference_golden_predict.npz with", len(OUT), "arrays")
for k_, v_ in META.items():
    print(k_, v_, "mse", OUT[k_ + "/mse"], "bic", OUT[k_ + "/bic"])
