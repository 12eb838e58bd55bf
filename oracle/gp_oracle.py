"""CPU ORACLE - TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.

A float64 restatement, op for op and unfused, of the exact-GP likelihood path of gpbasics 2.0.0 (paths relative to
/root/reference/main/gpbasics).  The arithmetic of the reference lives in TensorFlow (setup.py:3-9: tensorflow>=2.4,
tensorflow-probability>=0.12, un-pinned, not installable here); it is restated with torch CPU float64 ops of the same
meaning, so that reverse-mode autodiff through the Cholesky is available exactly as TF's GradientTape provides it
(Optimizer/Fitter.py:124-132).

Pinning: the reference ships no tests / golden vectors.  This oracle is pinned instead against
outputs of the UNMODIFIED reference sources executed in the build container on top of `oracle/tf_shim` (a stand-in for
the handful of TensorFlow ops the path uses); the generator is tests/golden/make_golden.py and the vectors are
tests/golden/*.npz (tests/test_oracle_golden.py).  It is additionally checked against mpmath and closed forms.

Kernel trees are plain tuples:
    ("SE",) ("PER",) ("LIN",) ("MAT32",) ("MAT52",) ("WN",) ("SE_ARD",)
    ("ADD", [children]) ("MUL", [children]) ("CP", [children])            # CP consumes len(children)-1 change points first
`hp` is the reference's hyper-parameter list (one tensor per entry, scalars or [d] vectors) in depth-first order.
"""
import math
from typing import List, Sequence

import numpy as np
import torch

DT = torch.float64

CP_SIGMOID, CP_INDICATOR, CP_APPROX = 0, 1, 2


# ---- Auxiliary/Distances.py:4-12 -------------------------------------------------------------------------------
def euclidian_distance(a: torch.Tensor, b: torch.Tensor, reference_formula: bool = True) -> torch.Tensor:
    if reference_formula:  # Distances.py:6-7 (sqrt of a^2 - 2ab + b^2; NaN-prone)
        return torch.sqrt(torch.sum(a * a, -1, keepdim=True) - 2 * torch.matmul(a, b.transpose(-1, -2))
                          + torch.sum(b * b, -1, keepdim=True).transpose(-1, -2))
    d = a.unsqueeze(-2) - b.unsqueeze(-3)
    return torch.sqrt(torch.sum(d * d, -1))


def manhattan_distance(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:  # Distances.py:12
    return torch.sum(torch.abs(a.unsqueeze(-2) - b.unsqueeze(-3)), -1)


# ---- number of hp list entries per node (BaseKernels.py:153-159,308-314,475-481; Operators.py:28-32,507-511) ----
def n_hp_entries(tree, scaled: bool = False) -> int:
    kind = tree[0]
    if kind in ("SE", "LIN", "MAT32", "MAT52", "SE_ARD"):
        return 1 + (1 if scaled else 0)
    if kind == "PER":
        return 2 + (1 if scaled else 0)
    if kind in ("WN", "L2", "L1"):
        return 0
    if kind in ("ADD", "MUL"):
        return sum(n_hp_entries(c, scaled) for c in tree[1])
    if kind == "CP":
        return sum(n_hp_entries(c, scaled) for c in tree[1]) + len(tree[1]) - 1
    raise ValueError(kind)


def _cp_s(x, cp, mode):
    if mode == CP_INDICATOR:  # Operators.py:396-400
        return (x.reshape(-1) < cp.reshape(())).to(DT).reshape(-1, 1)
    if mode == CP_SIGMOID:  # Operators.py:387-394
        return 0.5 * (1 + torch.tanh((cp - x) / 0.0025))
    return 1.0 / (1.0 + torch.exp(-1.0 * 100.0 * (x - cp)))  # Operators.py:379-385


def kernel_matrix(tree, hp: Sequence[torch.Tensor], x: torch.Tensor, x2: torch.Tensor, scaled: bool = False,
                  cp_mode: int = CP_INDICATOR, reference_distance: bool = True) -> torch.Tensor:
    """Kernel.get_tf_tensor (KernelBasics/Kernel.py:51) for the tuple tree."""
    kind = tree[0]
    assert len(hp) == n_hp_entries(tree, scaled), (tree, len(hp))
    if kind == "SE":  # BaseKernels.py:282-290
        dist = euclidian_distance(x, x2, reference_distance)
        sq = dist * dist
        k = torch.exp(-0.5 * (sq / (hp[0] * hp[0])))
        return hp[1] * k if scaled else k
    if kind == "PER":  # BaseKernels.py:446-453
        dist = manhattan_distance(x, x2)
        u = math.pi * (dist / hp[1])
        sine = torch.sin(u) ** 2
        k = torch.exp(-2 * sine / (hp[0] * hp[0]))
        return hp[2] * k if scaled else k
    if kind == "LIN":  # BaseKernels.py:119-130
        k = torch.matmul(x - hp[0], (x2 - hp[0]).transpose(-1, -2))
        return hp[1] * k if scaled else k
    if kind == "MAT32":  # BaseKernels.py:707-716
        dist = manhattan_distance(x, x2)
        l = torch.abs(hp[0])
        f = (math.sqrt(3.0) * dist) / l
        k = (1.0 + f) * torch.exp(-f)
        return hp[1] * k if scaled else k
    if kind == "MAT52":  # BaseKernels.py:864-876
        dist = manhattan_distance(x, x2)
        l = torch.abs(hp[0])
        f = (math.sqrt(5.0) * dist) / l
        third = (5.0 * dist * dist) / (3.0 * l * l)
        k = (1.0 + f + third) * torch.exp(-f)
        return hp[1] * k if scaled else k
    if kind == "WN":  # BaseKernels.py:646-662
        n, m = x.shape[0], x2.shape[0]
        k = torch.zeros(n, m, dtype=DT)
        q = min(n, m)
        k[:q, :q] = k[:q, :q] + torch.eye(q, dtype=DT)
        return k
    if kind == "L2":
        return euclidian_distance(x, x2, reference_distance)
    if kind == "L1":
        return manhattan_distance(x, x2)
    if kind == "SE_ARD":  # extension (config C5); equals SE(l=1) on x / l  
        d = (x.unsqueeze(-2) - x2.unsqueeze(-3)) / hp[0]
        k = torch.exp(-0.5 * torch.sum(d * d, -1))
        return hp[1] * k if scaled else k
    children = tree[1]
    if kind in ("ADD", "MUL"):  # Operators.py:306-326, 207-225 (left fold, consecutive hp slices)
        idx = 0
        res = None
        for c in children:
            nc = n_hp_entries(c, scaled)
            kc = kernel_matrix(c, hp[idx:idx + nc], x, x2, scaled, cp_mode, reference_distance)
            idx += nc
            res = kc if res is None else (res + kc if kind == "ADD" else res * kc)
        return res
    if kind == "CP":  # Operators.py:410-476
        if len(children) == 1:
            return kernel_matrix(children[0], hp, x, x2, scaled, cp_mode, reference_distance)
        ncp = len(children) - 1
        cps = hp[:ncp]
        idx = ncp
        prev = torch.tensor(1.0, dtype=DT)
        parts = []
        for i, c in enumerate(children):
            nc = n_hp_entries(c, scaled)
            kc = kernel_matrix(c, hp[idx:idx + nc], x, x2, scaled, cp_mode, reference_distance)
            idx += nc
            kc = kc * prev
            if i < ncp:
                s1 = _cp_s(x, cps[i], cp_mode)
                s2 = _cp_s(x2, cps[i], cp_mode)
                ind = torch.matmul(s1, s2.transpose(-1, -2))
                prev = torch.matmul(1.0 - s1, (1.0 - s2).transpose(-1, -2))
                kc = kc * ind
            parts.append(kc)
        res = parts[0]
        for pmat in parts[1:]:
            res = res + pmat
        return res
    raise ValueError(kind)


# ---- Statistics/CovarianceMatrix.py:197-265, Metrics/Metrics.py:138-154, Metrics/LogLikelihood.py:30-65 ----------
def nll(tree, hp: Sequence[torch.Tensor], noise: torch.Tensor, x: torch.Tensor, y: torch.Tensor, scaled: bool = False,
        cp_mode: int = CP_INDICATOR, reference_distance: bool = True, return_parts: bool = False):
    n = x.shape[0]
    K = kernel_matrix(tree, hp, x, x, scaled, cp_mode, reference_distance)
    Kn = K + noise * torch.eye(n, dtype=DT)                                   # CovarianceMatrix.py:201-202
    L = torch.linalg.cholesky(Kn)                                            # :250
    z = torch.linalg.solve_triangular(L, y, upper=False)                     # :260-262
    alpha = torch.linalg.solve_triangular(L.transpose(-1, -2), z, upper=True)
    data_fit = -0.5 * torch.matmul(y.transpose(-1, -2), alpha)               # LogLikelihood.py:39
    logdet = 2 * torch.sum(torch.log(torch.diagonal(L)))                     # Metrics.py:153-154
    penalty = -0.5 * logdet                                                  # LogLikelihood.py:41-42
    log_2_pi = math.log(math.pi * 2)                                         # :44
    norm = -0.5 * (n * log_2_pi)                                             # :45-46
    ll = (data_fit + penalty) + norm                                         # :49
    out = -ll                                                                # :65  shape [1,1]
    if return_parts:
        return out, K, L, alpha
    return out


def nll_and_grad(tree, hp_values: Sequence[np.ndarray], noise: float, x: np.ndarray, y: np.ndarray,
                 scaled: bool = False, cp_mode: int = CP_INDICATOR, reference_distance: bool = True,
                 optimize_noise: bool = False):
    """NLL and d NLL / d hp (autodiff, Optimizer/Fitter.py:124-132).  With optimize_noise the noise passed to the
    metric is |raw| and raw heads the gradient list (Fitter.py:94-95,107-108)."""
    hp = [torch.tensor(np.asarray(h, dtype=np.float64), dtype=DT, requires_grad=True) for h in hp_values]
    raw = torch.tensor(float(noise), dtype=DT, requires_grad=True)
    nz = torch.abs(raw) if optimize_noise else raw
    xt = torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=DT)
    yt = torch.as_tensor(np.asarray(y, dtype=np.float64), dtype=DT).reshape(-1, 1)
    # Up to n = 3000 the oracle runs on ONE thread: the multi-threaded LAPACK of some pool boxes returned a likelihood
    # 1.6e-10 off for the n = 1000 composite test problem (an 80-bit evaluation, tools/accuracy_probe.py, sides with the
    # device to 4e-13 and with the single-threaded oracle to 5e-13), which is more than the 1e-10 the tests hold the
    # device to.  Larger sizes keep all threads (tests compare them with frozen values instead).
    threads = torch.get_num_threads()
    single = xt.shape[-2] <= 3000 and threads > 1
    if single:
        torch.set_num_threads(1)
    try:
        val = nll(tree, hp, nz, xt, yt, scaled, cp_mode, reference_distance)
        grads = torch.autograd.grad(val.sum(), hp + [raw], allow_unused=True)
    finally:
        if single:
            torch.set_num_threads(threads)
    g = [np.zeros_like(np.asarray(h, dtype=np.float64)) if gi is None else gi.detach().numpy().copy()
         for h, gi in zip(list(hp_values) + [noise], grads)]
    return float(val.detach().reshape(-1)[0]), g[:-1], float(np.asarray(g[-1]))


def batch_nll(tree, hp: Sequence[torch.Tensor], noise: torch.Tensor, x: torch.Tensor, y: torch.Tensor, scaled: bool = False,
              cp_mode: int = CP_INDICATOR, reference_distance: bool = True, reference_aggregate: bool = True):
    """LogLikelihood.get_metric on a rank-3 BatchDataInput x [B, n, d], y [B, n, 1] (Metrics/LogLikelihood.py:30-65).
    The data fit is per entry [B, 1, 1] (:39) but get_log_determinant_cholesky reduces over EVERY axis
    (Metrics/Metrics.py:153-154: tf.reduce_sum without axis), so the same batch-wide log-determinant is added to every
    entry before p_batch_metric_aggregator = tf.reduce_mean (:62-63, global_parameters.py:64) - SURVEY App. B-3.
    reference_aggregate=False returns the mean of the true per-entry NLLs instead."""
    n = x.shape[-2]
    K = kernel_matrix(tree, hp, x, x, scaled, cp_mode, reference_distance)
    Kn = K + noise * torch.eye(n, dtype=DT)
    L = torch.linalg.cholesky(Kn)
    z = torch.linalg.solve_triangular(L, y, upper=False)
    alpha = torch.linalg.solve_triangular(L.transpose(-1, -2), z, upper=True)
    data_fit = -0.5 * torch.matmul(y.transpose(-1, -2), alpha)                       # [B, 1, 1]
    diag = torch.diagonal(L, dim1=-2, dim2=-1)                                       # tf.linalg.diag_part: [B, n]
    if reference_aggregate:
        logdet = 2 * torch.sum(torch.log(diag))                                      # scalar over the whole batch
    else:
        logdet = 2 * torch.sum(torch.log(diag), -1).reshape(-1, 1, 1)
    ll = (data_fit + (-0.5) * logdet) + (-0.5 * (n * math.log(math.pi * 2)))
    return -torch.mean(ll)


def blockwise_nll(trees: Sequence, hps: Sequence[Sequence[torch.Tensor]], noise, xs, ys, **kw):
    """BlockwiseLogLikelihood.get_metric (Metrics/LogLikelihood.py:77-104): sum of per-block NLLs."""
    total = None
    for t, h, xb, yb in zip(trees, hps, xs, ys):
        v = nll(t, h, noise, xb, yb, **kw)
        total = v if total is None else total + v
    return total


# ---- index bookkeeping (integer, bit-exact) ------------------------------------------------------------------------
def blockwise_segments(x: np.ndarray, change_points: Sequence[float]) -> List[np.ndarray]:
    """BlockwiseDataInput.__init__ (DataHandling/DataInput.py:231-244): segment i = {x < cp_0}, {cp_{i-1} <= x < cp_i},
    {x >= cp_last}; tf.where(...)[:, 0] on an [n,1] input = ascending row indices."""
    x = np.asarray(x, dtype=np.float64)
    out = []
    ncp = len(change_points)
    for i in range(ncp + 1):
        if i == 0:
            mask = x < float(change_points[0])
        elif i == ncp:
            mask = x >= float(change_points[i - 1])
        else:
            mask = np.logical_and(x < float(change_points[i]), x >= float(change_points[i - 1]))
        out.append(np.where(mask)[0].astype(np.int64))
    return out


def partition_indices(score_columns: Sequence[np.ndarray], smallest_distance: bool = False,
                      tie_noise: np.ndarray = None) -> List[np.ndarray]:
    """PartitioningModel.get_data_record_indices_per_partition (KernelBasics/PartitioningModel.py:109-131)."""
    score = np.transpose(np.array(score_columns))
    if smallest_distance:
        if tie_noise is not None:
            score = score + tie_noise
        col_min = np.amin(score, axis=1)
        score = score == col_min.reshape(-1, 1)
    out = []
    for i in range(len(score_columns)):
        out.append(np.where(score[:, i] == 1)[0])
    if len(out) == 0:
        out = [np.linspace(0, score.shape[0] - 1, score.shape[0], dtype=int)]
    return out


def hp_slices(child_counts: Sequence[int], start: int = 0) -> List[slice]:
    """Consecutive hp slices per child; the index advances even for empty blocks
    (KernelBasics/PartitionOperator.py:63-82, Statistics/CovarianceMatrix.py:316-339)."""
    out = []
    idx = start
    for c in child_counts:
        out.append(slice(idx, idx + c))
        idx += c
    return out
