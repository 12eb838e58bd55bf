"""Generates tests/golden/reference_golden_r2.npz by running the UNMODIFIED reference (/root/reference/main/gpbasics on
top of oracle/tf_shim; see make_golden.py for what that means) on the parts of the path that had no golden vector yet:

  batch3/*      rank-3 BatchDataInput likelihood (Metrics/LogLikelihood.py:49,62-63; Metrics/Metrics.py:152-154;
                SURVEY App. B-3: log-determinant summed over the whole batch, data fit per entry, reduce_mean) with its
                gradient w.r.t. the hyper-parameters and the noise
  ppred/*       partitioned prediction: block-rectangular K_s of a PartitionOperator incl. empty train / test partitions
                (Auxiliary/NonSquareBlockMatrices.py:8-103, KernelBasics/PartitionOperator.py:24-83), posterior mean of a
                PartitionedGaussianProcess (Statistics/GaussianProcess.py:42-85)
  bmetric/*     blockwise BIC / MSE of a PartitionedGaussianProcess (Metrics/BayesianInformationCriterion.py:43-63,
                Metrics/MeanSquaredError.py:45-81)

Run from the repository root:    python tests/golden/make_golden_r2.py
The file it writes is committed; tests/test_oracle_golden.py (CPU) and tests/test_gpu_golden.py (GPU) read it.
"""
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
import gpbasics.KernelBasics.PartitionOperator as po  # noqa: E402
import gpbasics.KernelBasics.PartitioningModel as pm  # noqa: E402
import gpbasics.DataHandling.DataInput as di  # noqa: E402
import gpbasics.DataHandling.BatchDataInput as bdi  # noqa: E402
import gpbasics.MeanFunctionBasics.BaseMeanFunctions as bmf  # noqa: E402
import gpbasics.Statistics.GaussianProcess as gproc  # noqa: E402
import gpbasics.Metrics.Auxiliary as met_aux  # noqa: E402
import gpbasics.Metrics.Metrics as met  # noqa: E402
import gpbasics.Metrics.MatrixHandlingTypes as mht  # noqa: E402

OUT, META = {}, {}
LEAF = {"SE": bk.SquaredExponentialKernel, "PER": bk.PeriodicKernel, "LIN": bk.LinearKernel}
A = (mht.MatrixApproximations.NONE, mht.NumericalMatrixHandlingType.CHOLESKY_BASED)


def build(spec, d=1):
    if spec[0] in LEAF:
        return LEAF[spec[0]](d)
    cls = op.AdditionOperator if spec[0] == "ADD" else op.MultiplicationOperator
    return cls(d, [build(c, d) for c in spec[1]])


def tf_hp(values):
    return [tf.Variable(np.asarray(v, dtype=np.float64), dtype=tf.float64) for v in values]


COMPOSITE = ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)])

# ---- rank-3 batch likelihood ----------------------------------------------------------------------------------------
for name, spec, hpv, B, n, noise in [("batch3_se", ("SE",), [0.2], 3, 50, 1e-2),
                                     ("batch3_composite", COMPOSITE, [0.2, 0.4, 0.3, [0.1]], 4, 40, 5e-2)]:
    rng = np.random.default_rng(50 + B)
    xs = np.stack([np.linspace(0.0, 1.0, n)[:, None] + 0.01 * b for b in range(B)])
    ys = np.stack([np.sin((8 + b) * xs[b]) + 0.1 * rng.standard_normal((n, 1)) for b in range(B)])
    kernel = build(spec)
    hp = tf_hp(hpv)
    din = bdi.BatchDataInput(tf.constant(xs), tf.constant(ys), tf.constant(xs), tf.constant(ys))
    din.set_mean_function(bmf.ZeroMeanFunction(1))
    gp = gproc.GaussianProcess(kernel, bmf.ZeroMeanFunction(1))
    gp.set_data_input(din)
    metric = met_aux.get_metric_by_type(met.MetricType.LL, gp, *A)
    raw = tf.Variable(noise, dtype=tf.float64)
    with tf.GradientTape() as g:
        val = metric.get_metric(hp, raw, None)
        grads = g.gradient(val, hp + [raw])
    OUT[name + "/x"], OUT[name + "/y"] = xs, ys
    OUT[name + "/noise"] = np.float64(noise)
    OUT[name + "/hp"] = np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in hpv])
    OUT[name + "/nll"] = np.asarray(val.numpy()).reshape(-1)
    OUT[name + "/grad"] = np.concatenate([gr.numpy().reshape(-1) for gr in grads[:-1]])
    OUT[name + "/grad_noise"] = np.asarray(grads[-1].numpy()).reshape(-1)
    # the true per-entry values, for the record (what reference_batch_aggregate=False averages)
    per = []
    for b in range(B):
        d1 = di.DataInput(xs[b], ys[b], xs[b], ys[b])
        d1.set_mean_function(bmf.ZeroMeanFunction(1))
        g1 = gproc.GaussianProcess(build(spec), bmf.ZeroMeanFunction(1))
        g1.set_data_input(d1)
        per.append(float(met_aux.get_metric_by_type(met.MetricType.LL, g1, *A).get_metric(hp, raw, None).numpy().reshape(-1)[0]))
    OUT[name + "/per_entry_nll"] = np.asarray(per)
    META[name] = {"spec": json.dumps(spec), "B": B, "n": n, "hp_sizes": [int(np.size(v)) for v in hpv]}


# ---- partitioned prediction and blockwise metrics ----------------------------------------------------------------------
class Interval(pm.PartitionCriterion):
    def __init__(self, lo, hi):
        super().__init__(pm.PartitioningClass.SELF_SUFFICIENT)
        self.lo, self.hi = lo, hi

    def get_score(self, x_vector):
        c = np.asarray(x_vector)[:, 0]
        return np.logical_and(c >= self.lo, c < self.hi).astype(np.float64)

    def deepcopy(self):
        return Interval(self.lo, self.hi)

    def get_json(self):
        return {"lo": self.lo, "hi": self.hi}


def partition_case(name, edges, specs, hpv, n, n_test, seed, test_lo=0.0, test_hi=1.0, metrics=True):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, 1))
    y = np.sin(9 * x) + 0.3 * x + 0.1 * rng.standard_normal((n, 1))
    xt = rng.uniform(test_lo, test_hi, (n_test, 1))
    yt = np.sin(9 * xt) + 0.3 * xt + 0.1 * rng.standard_normal((n_test, 1))
    model = pm.PartitioningModel(pm.PartitioningClass.SELF_SUFFICIENT, [])
    model.init_partitioning([Interval(edges[i], edges[i + 1]) for i in range(len(edges) - 1)])
    kernel = po.PartitionOperator(1, [build(s) for s in specs], model)
    hp = tf_hp(hpv)
    nz = tf.constant(1e-2, dtype=tf.float64)
    OUT[name + "/x"], OUT[name + "/y"], OUT[name + "/xt"], OUT[name + "/yt"] = x, y, xt, yt
    OUT[name + "/edges"] = np.asarray(edges)
    OUT[name + "/hp"] = np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in hpv])
    meta = {"specs": json.dumps(specs), "hp_sizes": [int(np.size(v)) for v in hpv]}
    # block-rectangular K_s straight from the operator (x rows / x_test columns, each side partition-major)
    idx = model.get_data_record_indices_per_partition(x)
    idx_t = model.get_data_record_indices_per_partition(xt)
    for i, (a, b) in enumerate(zip(idx, idx_t)):
        OUT[name + "/idx%d" % i] = np.asarray(a, dtype=np.int64)
        OUT[name + "/idxt%d" % i] = np.asarray(b, dtype=np.int64)
    try:
        Ks = kernel.get_tf_tensor(hp, x, xt)
        OUT[name + "/K_s"] = Ks.numpy()
        meta["K_s_ok"] = True
    except Exception as e:
        meta["K_s_ok"] = False
        meta["K_s_error"] = repr(e)
    if metrics:
        base = di.DataInput(x, y, xt, yt)
        pdi = model.partition_data_input(base)
        pdi.set_mean_function(bmf.ZeroMeanFunction(1))
        pgp = gproc.PartitionedGaussianProcess(kernel, bmf.ZeroMeanFunction(1))
        pgp.set_data_input(pdi)
        OUT[name + "/xt_reordered"] = pdi.data_x_test.numpy()
        for key, mt in [("bic", met.MetricType.blockwise_BIC), ("mse", met.MetricType.blockwise_MSE),
                        ("ll", met.MetricType.blockwise_LL)]:
            try:
                m = met_aux.get_metric_by_type(mt, pgp, *A)
                OUT[name + "/blockwise_" + key] = np.asarray(m.get_metric(hp, nz, None).numpy()).reshape(-1)
                meta[key + "_ok"] = True
            except Exception as e:
                meta[key + "_ok"] = False
                meta[key + "_error"] = repr(e)
        try:
            total, mean_mu, post_mu = pgp.predict(hp, None, nz)
            OUT[name + "/predict_total"] = np.asarray(total.numpy()).reshape(-1)
            OUT[name + "/predict_post_mu"] = np.asarray(post_mu.numpy()).reshape(-1)
            meta["predict_ok"] = True
        except Exception as e:
            meta["predict_ok"] = False
            meta["predict_error"] = repr(e)
    META[name] = meta


# every partition holds train and test points: the whole chain of the reference runs
partition_case("ppred_full", [0.0, 0.35, 0.7, 1.0 + 1e-9], [("SE",), ("PER",), COMPOSITE],
               [0.1, 0.3, 0.2, 0.12, 0.2, 0.15, [0.3]], 210, 60, 61)
# a partition without test points (dead rows) and one without training points (dead columns): K_s only
partition_case("ppred_dead", [0.0, 0.3, 0.5, 0.5, 0.8, 1.0 + 1e-9], [("SE",), ("PER",), ("SE",), ("LIN",), ("SE",)],
               [0.1, 0.3, 0.2, 0.15, [0.3], 0.2], 150, 40, 62, test_lo=0.31, test_hi=0.99, metrics=False)

OUT["__meta__"] = np.frombuffer(json.dumps(META).encode("utf-8"), dtype=np.uint8)
np.savez_compressed(os.path.join(HERE, "reference_golden_r2.npz"), **OUT)
print("wrote reference_golden_r2.npz with", len(OUT), "arrays")
for k_, v_ in META.items():
    print(k_, v_)
for k_ in sorted(OUT):
    if k_.endswith(("nll", "blockwise_bic", "blockwise_mse", "blockwise_ll", "per_entry_nll")):
        print(k_, OUT[k_])
