"""Generates tests/golden/reference_golden.npz by running the UNMODIFIED reference sources (/root/reference/main/gpbasics)
in this container on top of oracle/tf_shim (TensorFlow itself is not installable here).

What executes is the reference's own Python: its kernel classes and operator trees, its hyper-parameter slicing, its
CovarianceMatrix / LogLikelihood / BlockwiseLogLikelihood code, its BlockwiseDataInput and PartitioningModel index
bookkeeping, and its VariationalSgdFitter.fit(); the shim supplies only the TensorFlow leaf ops (float64 torch CPU) and
tf.GradientTape (torch.autograd).  Run from the repository root:

    python tests/golden/make_golden.py

The file it writes is committed; tests/test_oracle_golden.py (CPU) pins the oracle against it and
tests/test_gpu_golden.py (GPU) compares the CUDA path with it.  /root/reference is NOT needed by any test.
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
import gpbasics.MeanFunctionBasics.BaseMeanFunctions as bmf  # noqa: E402
import gpbasics.Statistics.GaussianProcess as gproc  # noqa: E402
import gpbasics.Metrics.Auxiliary as met_aux  # noqa: E402
import gpbasics.Metrics.Metrics as met  # noqa: E402
import gpbasics.Metrics.MatrixHandlingTypes as mht  # noqa: E402
import gpbasics.Optimizer.Fitter as fitter  # noqa: E402

OUT = {}
META = {}


def build(spec, d):
    kind = spec[0]
    leaf = {"SE": bk.SquaredExponentialKernel, "PER": bk.PeriodicKernel, "LIN": bk.LinearKernel,
            "MAT32": bk.MaternKernel3_2, "MAT52": bk.MaternKernel5_2, "WN": bk.WhiteNoiseKernel}
    if kind in leaf:
        return leaf[kind](d)
    children = [build(c, d) for c in spec[1]]
    if kind == "ADD":
        return op.AdditionOperator(d, children)
    if kind == "MUL":
        return op.MultiplicationOperator(d, children)
    if kind == "CP":
        return op.ChangePointOperator(d, children, [tf.Variable(c, dtype=tf.float64) for c in spec[2]])
    raise ValueError(kind)


def tf_hp(values):
    return [tf.Variable(np.asarray(v, dtype=np.float64), dtype=tf.float64) for v in values]


def data(n, seed, d=1):
    rng = np.random.default_rng(seed)
    if d == 1:
        x = np.linspace(0.0, 1.0, n)[:, None]
    else:
        x = np.sort(rng.uniform(0, 1, (n, d)), axis=0)
    y = np.sin(12 * x[:, :1]) * (1 + x[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    return x, y


def holistic_case(name, spec, hp_values, n, noise, seed, scaled=False, cp_mode=None, store_matrices=False,
                  optimize_noise=False):
    global_param.p_scaled_base_kernel = scaled
    global_param.p_cp_operator_type = cp_mode or global_param.ChangePointOperatorType.INDICATOR
    x, y = data(n, seed)
    kernel = build(spec, 1)
    hp = tf_hp(hp_values)
    cps = [tf.Variable(c, dtype=tf.float64) for c in spec[2]] if spec[0] == "CP" else []
    full_hp = cps + hp
    din = di.DataInput(x, y, x, y)
    din.set_mean_function(bmf.ZeroMeanFunction(1))
    gp = gproc.GaussianProcess(kernel, bmf.ZeroMeanFunction(1))
    gp.set_data_input(din)
    metric = met_aux.get_metric_by_type(met.MetricType.LL, gp, mht.MatrixApproximations.NONE,
                                        mht.NumericalMatrixHandlingType.CHOLESKY_BASED)
    raw = tf.Variable(noise, dtype=tf.float64)
    with tf.GradientTape() as g:
        nz = tf.abs(raw) if optimize_noise else raw
        val = metric.get_metric(full_hp, nz, None)
        grads = g.gradient(val, full_hp + [raw])
    OUT[name + "/x"] = x
    OUT[name + "/y"] = y
    OUT[name + "/noise"] = np.float64(noise)
    OUT[name + "/nll"] = val.numpy().reshape(-1)
    OUT[name + "/n_hp"] = np.int64(len(full_hp))
    for i, (h, gr) in enumerate(zip(full_hp, grads[:-1])):
        OUT[name + "/hp%d" % i] = h.numpy()
        OUT[name + "/grad%d" % i] = np.zeros_like(h.numpy()) if gr is None else gr.numpy()
        OUT[name + "/grad%d_none" % i] = np.bool_(gr is None)
    OUT[name + "/grad_noise"] = grads[-1].numpy()
    if store_matrices:
        cov = gp.covariance_matrix
        cov.reset()
        OUT[name + "/K"] = cov.get_K(full_hp).numpy()
        OUT[name + "/L"] = cov.get_L_K(full_hp, raw).numpy()
        OUT[name + "/alpha"] = cov.get_L_alpha(full_hp, raw).numpy()
    META[name] = {"spec": json.dumps(spec), "scaled": bool(scaled), "cp_mode": global_param.p_cp_operator_type.value,
                  "kind": "holistic", "string": kernel.get_string_representation(),
                  "names": kernel.get_hyper_parameter_names(0), "optimize_noise": bool(optimize_noise)}
    global_param.p_scaled_base_kernel = False
    global_param.p_cp_operator_type = global_param.ChangePointOperatorType.INDICATOR


COMPOSITE = ("MUL", [("ADD", [("SE",), ("PER",)]), ("LIN",)])

holistic_case("se_n200", ("SE",), [0.1], 200, 1e-2, 0, store_matrices=False)
holistic_case("se_n64_mats", ("SE",), [0.15], 64, 1e-2, 1, store_matrices=True)
holistic_case("per_n150", ("PER",), [0.5, 0.3], 150, 1e-2, 2)
holistic_case("lin_n150", ("LIN",), [[0.01]], 150, 5e-2, 3)
holistic_case("composite_n300", COMPOSITE, [0.1, 0.1, 0.1, [0.01]], 300, 1e-2, 4)
holistic_case("composite_n96_mats", COMPOSITE, [0.12, 0.2, 0.15, [0.3]], 96, 1e-2, 5, store_matrices=True)
holistic_case("composite_scaled_n200", COMPOSITE, [0.1, 0.7, 0.1, 0.2, 1.5, [0.01], 0.3], 200, 1e-2, 6, scaled=True)
holistic_case("matern_wn_n120", ("ADD", [("MAT32",), ("MAT52",), ("WN",)]), [0.2, 0.3], 120, 1e-2, 7)
holistic_case("deep_n260", ("ADD", [("MUL", [("SE",), ("PER",), ("LIN",)]),
                                    ("MUL", [("SE",), ("ADD", [("LIN",), ("PER",)])])]),
              [0.2, 0.4, 0.25, [0.3], 0.15, [-0.2], 0.6, 0.35], 260, 1e-2, 8)
holistic_case("cp_indicator_n240", ("CP", [("SE",), ("PER",), ("ADD", [("SE",), ("LIN",)])], [0.31, 0.67]),
              [0.1, 0.2, 0.15, 0.08, [0.5]], 240, 1e-2, 9)
holistic_case("cp_approx_n240", ("CP", [("SE",), ("PER",), ("ADD", [("SE",), ("LIN",)])], [0.31, 0.67]),
              [0.1, 0.2, 0.15, 0.08, [0.5]], 240, 1e-2, 10,
              cp_mode=global_param.ChangePointOperatorType.APPROX_INDICATOR)
holistic_case("cp_sigmoid_n200", ("CP", [("SE",), ("PER",)], [0.5]), [0.1, 0.2, 0.15], 200, 1e-2, 11,
              cp_mode=global_param.ChangePointOperatorType.SIGMOID)
holistic_case("se_optimize_noise_n180", ("SE",), [0.2], 180, -0.05, 12, optimize_noise=True)
holistic_case("se_default_jitter_n100", ("SE",), [0.1], 100, 1e-8, 13)

# ---- hp order after the implicit child sort of type_compare_to (SURVEY App. C KAT) --------------------------------
k1 = build(COMPOSITE, 1)
k2 = build(COMPOSITE, 1)
before = k1.get_string_representation()
k1.type_compare_to(k2)
META["sort_kat"] = {"before": before, "after": k1.get_string_representation(),
                    "names_after": k1.get_hyper_parameter_names(0)}

# ---- defaults and bounds ----------------------------------------------------------------------------------------------
kd = build(COMPOSITE, 1)
xr = [[0.25, 2.25]]
OUT["defaults/composite"] = np.concatenate([h.numpy().reshape(-1) for h in kd.get_default_hyper_parameter(xr, 500)])
bl = kd.get_hyper_parameter_bounds(xr, 500)
OUT["bounds/composite_lo"] = np.concatenate([np.asarray(b[0].numpy()).reshape(-1) for b in bl])
OUT["bounds/composite_hi"] = np.concatenate([np.asarray(b[1].numpy()).reshape(-1) for b in bl])

# ---- block-wise (change points as segment boundaries) ------------------------------------------------------------------
n = 330
x, y = data(n, 20)
cps = [0.2, 0.2 + 1e-12, 0.55, 0.9]   # includes an (almost) empty segment
bdi = di.BlockwiseDataInput(x, y, x, y, [tf.constant(c, dtype=tf.float64) for c in cps])
bdi.set_mean_function(bmf.ZeroMeanFunction(1))
OUT["blockwise/x"] = x
OUT["blockwise/y"] = y
OUT["blockwise/cps"] = np.asarray(cps)
for i, blk in enumerate(bdi.data_inputs):
    OUT["blockwise/seg%d_x" % i] = blk.data_x_train.numpy()
    OUT["blockwise/seg%d_n" % i] = np.int64(blk.n_train)
specs = [("SE",), ("PER",), ("ADD", [("SE",), ("LIN",)]), ("SE",), ("MUL", [("SE",), ("PER",)])]
children = [build(s, 1) for s in specs]
cpk = op.ChangePointOperator(1, children, [tf.Variable(c, dtype=tf.float64) for c in cps])
hpv = [0.1, 0.2, 0.15, 0.08, [0.5], 0.12, 0.3, 0.25, 0.2]
hp_children = tf_hp(hpv)
full = [tf.Variable(c, dtype=tf.float64) for c in cps] + hp_children
bgp = gproc.BlockwiseGaussianProcess(cpk, bmf.ZeroMeanFunction(1))
bgp.set_data_input(bdi)
noise = tf.constant(1e-2, dtype=tf.float64)
# per-block values straight from SegmentedCovarianceMatrix (hp start index = #cps, CovarianceMatrix.py:319-320)
Ls = bgp.covariance_matrix.get_L_K_blocks(full, noise)
als = bgp.covariance_matrix.get_L_alpha_blocks(full, noise)
block_nll = []
for i, blk in enumerate(bdi.data_inputs):
    if Ls[i] is None:
        block_nll.append(np.nan)
        continue
    yb = blk.get_detrended_y_train().numpy()
    L = Ls[i].numpy()
    a = als[i].numpy()
    block_nll.append(float(0.5 * (yb.T @ a) + np.sum(np.log(np.diag(L))) + 0.5 * blk.n_train * np.log(np.pi * 2)))
OUT["blockwise/block_nll"] = np.asarray(block_nll)
OUT["blockwise/hp_children"] = np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in hpv])
META["blockwise"] = {"specs": json.dumps(specs), "hp_sizes": [len(np.atleast_1d(v)) for v in hpv]}
# the same data through the holistic change-point kernel: equal to the block sum for non-empty blocks
gp_h = gproc.GaussianProcess(cpk.deepcopy(), bmf.ZeroMeanFunction(1))
din = di.DataInput(x, y, x, y)
din.set_mean_function(bmf.ZeroMeanFunction(1))
gp_h.set_data_input(din)
mh = met_aux.get_metric_by_type(met.MetricType.LL, gp_h)
OUT["blockwise/holistic_nll"] = mh.get_metric(full, noise, None).numpy().reshape(-1)


# ---- partition operator + partitioned GP + blockwise_LL ------------------------------------------------------------------
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


rng = np.random.default_rng(30)
n = 280
xp = rng.uniform(0, 1, (n, 1))          # unsorted on purpose: partitioning re-orders partition-major
yp = np.sin(9 * xp) + 0.1 * rng.standard_normal((n, 1))
edges = [0.0, 0.3, 0.3, 0.62, 1.0 + 1e-9]   # second interval empty
model = pm.PartitioningModel(pm.PartitioningClass.SELF_SUFFICIENT, [])
model.init_partitioning([Interval(edges[i], edges[i + 1]) for i in range(4)])
idx = model.get_data_record_indices_per_partition(xp)
for i, ix in enumerate(idx):
    OUT["partition/idx%d" % i] = np.asarray(ix, dtype=np.int64)
pspecs = [("SE",), ("PER",), COMPOSITE, ("SE",)]
pkernel = po.PartitionOperator(1, [build(s, 1) for s in pspecs], model)
php_v = [0.1, 0.3, 0.2, 0.12, 0.2, 0.15, [0.3], 0.2]
php = tf_hp(php_v)
base = di.DataInput(xp, yp, xp, yp)
pdi = model.partition_data_input(base)
pdi.set_mean_function(bmf.ZeroMeanFunction(1))
pgp = gproc.PartitionedGaussianProcess(pkernel, bmf.ZeroMeanFunction(1))
pgp.set_data_input(pdi)
bll = met_aux.get_metric_by_type(met.MetricType.blockwise_LL, pgp)
OUT["partition/x"] = xp
OUT["partition/y"] = yp
OUT["partition/edges"] = np.asarray(edges)
OUT["partition/x_reordered"] = pdi.data_x_train.numpy()
OUT["partition/hp"] = np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in php_v])
try:
    with tf.GradientTape() as g:
        v = bll.get_metric(php, noise, None)
        gr = g.gradient(v, php)
    OUT["partition/blockwise_nll"] = v.numpy().reshape(-1)
    OUT["partition/grad"] = np.concatenate([np.zeros_like(h.numpy()).reshape(-1) if gi is None else gi.numpy().reshape(-1)
                                            for h, gi in zip(php, gr)])
    META["partition"] = {"specs": json.dumps(pspecs), "ok": True}
except Exception as e:  # the reference cannot build a DataInput for an empty partition
    META["partition"] = {"specs": json.dumps(pspecs), "ok": False, "error": repr(e)}
Kd = pkernel.get_tf_tensor(php, xp, xp)
OUT["partition/K_dense"] = Kd.numpy()

# ---- VariationalSgdFitter.fit(): one step (Optimizer/Fitter.py:61-170) --------------------------------------------------
x, y = data(150, 40)
x = x * 2.0 + 0.25
fk = build(COMPOSITE, 1)
fdi = di.DataInput(x, y, x, y)
fdi.set_mean_function(bmf.ZeroMeanFunction(1))
fgp = gproc.GaussianProcess(fk, bmf.ZeroMeanFunction(1))
global_param.p_cov_matrix_jitter = tf.constant(1e-2, dtype=tf.float64)   # well conditioned
f = fitter.VariationalSgdFitter(fdi, fgp, met.MetricType.LL, False, mht.MatrixApproximations.NONE,
                                mht.NumericalMatrixHandlingType.CHOLESKY_BASED)
# capture what sgd_opt.minimize differentiated
import tensorflow_probability as tfp  # noqa: E402
captured = {}
_orig = tfp.optimizer.VariationalSGD.minimize


def _spy(self, loss, var_list, tape=None):
    _orig(self, loss, var_list, tape)
    captured["grads"] = [g.numpy() for g in self.last_grads]
    captured["loss"] = self.last_loss.numpy()


tfp.optimizer.VariationalSGD.minimize = _spy
pre, post, hps, nz, _ = f.fit()
tfp.optimizer.VariationalSGD.minimize = _orig
global_param.p_cov_matrix_jitter = tf.constant(1e-8, dtype=tf.float64)
OUT["fit/x"] = x
OUT["fit/y"] = y
OUT["fit/pre"] = pre.numpy().reshape(-1)
OUT["fit/post"] = post.numpy().reshape(-1)
OUT["fit/grads"] = np.concatenate([g.reshape(-1) for g in captured["grads"]])
OUT["fit/hp_after"] = np.concatenate([h.numpy().reshape(-1) for h in hps])
OUT["fit/noise"] = np.float64(nz.numpy())

OUT["__meta__"] = np.frombuffer(json.dumps(META).encode("utf-8"), dtype=np.uint8)
np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **OUT)
print("wrote", os.path.join(HERE, "reference_golden.npz"), "with", len(OUT), "arrays")
for k_, v_ in META.items():
    print(k_, v_)
