"""Negative log marginal likelihood (mirror of gpbasics/Metrics/LogLikelihood.py:16-104).

    get_metric = 1/2 y^T alpha + sum(log diag L) + 1/2 n log(2 pi)      (LogLikelihood.py:39-49, :65; minimum = optimum)

One host call evaluates it - and, on request, its gradient w.r.t. every hyper-parameter and the noise - through the
fused plan: assembly, blocked Cholesky with y carried as an extra row, (inverse, trace gradient).  When a
hyper-parameter tensor has requires_grad the returned value is connected to torch.autograd, so optimisers that
differentiate `get_metric` (the reference's `sgd_opt.minimize(opt, hp)`, Optimizer/Fitter.py:154-158) keep working."""
from typing import List

import numpy as np
import torch

from .. import global_parameters as global_param
from ..KernelBasics import Operators as op
from . import MatrixHandlingTypes as mht
from .Metrics import AbstractMetric, Metric, MetricType


class _NllFunction(torch.autograd.Function):
    """bridges the device gradient into torch.autograd: forward evaluates NLL and gradient together"""

    @staticmethod
    def forward(ctx, evaluator, noise, *hp):
        value, grads, gnoise = evaluator(list(hp), noise, True)
        ctx.grads = [torch.as_tensor(g, dtype=torch.float64) for g in grads]
        ctx.gnoise = gnoise
        ctx.shapes = [h.shape for h in hp]
        return torch.tensor([[value]], dtype=torch.float64)

    @staticmethod
    def backward(ctx, gout):
        s = gout.reshape(-1)[0]
        hp_grads = [(s * g).reshape(shape) for g, shape in zip(ctx.grads, ctx.shapes)]
        return (None, s * torch.tensor(ctx.gnoise, dtype=torch.float64)) + tuple(hp_grads)


def _wants_autograd(hyper_parameter, noise) -> bool:
    if not torch.is_grad_enabled():
        return False
    ts = [h for h in hyper_parameter if isinstance(h, torch.Tensor)]
    if isinstance(noise, torch.Tensor):
        ts.append(noise)
    return any(t.requires_grad for t in ts)


def _evaluate(evaluator, hyper_parameter, noise):
    if _wants_autograd(hyper_parameter, noise):
        hp = [h if isinstance(h, torch.Tensor) else torch.as_tensor(h, dtype=torch.float64) for h in hyper_parameter]
        nz = noise if isinstance(noise, torch.Tensor) else torch.as_tensor(noise, dtype=torch.float64)
        return _NllFunction.apply(evaluator, nz, *hp)
    value, _, _ = evaluator(hyper_parameter, noise, False)
    return torch.tensor([[value]], dtype=torch.float64)


class AbstractLogLikelihood(Metric):
    pass


class LogLikelihood(AbstractLogLikelihood):
    def __init__(self, data_input, covariance_matrix, local_approx, numerical_matrix_handling, subset_size: int = None,
                 reference_batch_aggregate: bool = True):
        super().__init__(data_input, covariance_matrix, MetricType.LL, local_approx, numerical_matrix_handling,
                         subset_size)
        self.reference_batch_aggregate = reference_batch_aggregate
        self._batch_blocks = None
        self._batch_weights = None

    # ---- evaluation ------------------------------------------------------------------------------------------------
    def _eval(self, hyper_parameter, noise, want_grad):
        if self.data_input.data_x_train.dim() == 3:
            return self._eval_batch(hyper_parameter, noise, want_grad)
        return self.covariance_matrix.nll_and_grad(hyper_parameter, noise, want_grad)

    def _eval_batch(self, hyper_parameter, noise, want_grad):
        """rank-3 input [B, n, d]: B GPs sharing kernel and hyper-parameters, one batched plan.  The reference sums
        the log-determinant over the whole batch before averaging (Metrics.py:153-154, LogLikelihood.py:49,62-63;
        SURVEY App. B-3):   value = mean_b(1/2 y_b^T alpha_b) + sum_b sum(log diag L_b) + 1/2 n log(2 pi).
        `reference_batch_aggregate` reproduces that (value and gradient: the aggregate is linear in the per-GP terms, so
        the device reports every GP's gradient with the weights (1/B, 1) on its two terms and the host adds them up);
        otherwise the mean of the true per-entry NLLs is returned."""
        from ..Statistics._device import DeviceBlocks
        x, y = self.data_input.data_x_train, self.data_input.get_detrended_y_train()
        B, n = x.shape[0], x.shape[1]
        kern = self.covariance_matrix.kernel
        if self._batch_blocks is not None and not self._batch_blocks.matches([kern] * B, y):
            self._batch_blocks = None
        if self._batch_blocks is None:
            self._batch_blocks = DeviceBlocks([kern] * B, [x[b] for b in range(B)], [y[b] for b in range(B)], True,
                                              y_source=y)
            self._batch_weights = None
        blocks = self._batch_blocks
        weights = (1.0 / B, 1.0) if self.reference_batch_aggregate else (1.0, 1.0)
        if self._batch_weights != weights:
            for b in range(B):
                blocks.plan.set_grad_weights(b, *weights)
            self._batch_weights = weights
        s2 = float(torch.as_tensor(noise, dtype=torch.float64))
        nll, grads = blocks.evaluate([hyper_parameter] * B, [s2] * B, want_grad)
        kern._remember(hyper_parameter)
        const = 0.5 * n * np.log(np.pi * 2)
        if self.reference_batch_aggregate:
            quad, half_logdet = blocks.plan.last_terms()          # one D2H for the whole batch (part of eval_host)
            value = float(np.mean(0.5 * quad) + np.sum(half_logdet) + const)
        else:
            value = float(np.mean(nll))
        if not want_grad:
            return value, None, None
        glists, gnoise = blocks.grads_as_lists(grads, [hyper_parameter] * B)
        scale = 1.0 if self.reference_batch_aggregate else 1.0 / B
        total = [sum(torch.as_tensor(g[i]) for g in glists) * scale for i in range(len(hyper_parameter))]
        return value, total, float(np.sum(gnoise) * scale)

    def get_metric(self, hyper_parameter: List[torch.Tensor], noise, indices=None, reset: bool = True) -> torch.Tensor:
        if reset:
            self.covariance_matrix.reset()
            self.last_covariance_matrix = None
        return _evaluate(self._eval, hyper_parameter, noise)

    def get_gradients(self, hyper_parameter, noise, reset: bool = True, with_noise: bool = False):
        if reset:
            self.covariance_matrix.reset()
            self.last_covariance_matrix = None
        _, grads, gnoise = self._eval(hyper_parameter, noise, True)
        grads = [torch.as_tensor(g, dtype=torch.float64) for g in grads]
        return (grads, torch.tensor(gnoise, dtype=torch.float64)) if with_noise else grads

    def get_metric_and_gradients(self, hyper_parameter, noise):
        """(NLL [1,1], [d NLL / d hp], d NLL / d noise) from one fused evaluation"""
        self.covariance_matrix.reset()
        value, grads, gnoise = self._eval(hyper_parameter, noise, True)
        return torch.tensor([[value]], dtype=torch.float64), \
            [torch.as_tensor(g, dtype=torch.float64) for g in grads], torch.tensor(gnoise, dtype=torch.float64)


class BlockwiseLogLikelihood(AbstractMetric):
    """sum of the per-block NLLs of a Blockwise / PartitionedGaussianProcess (LogLikelihood.py:77-104), evaluated as one
    batched plan over all non-empty blocks.  The reference starts slicing the hyper-parameter list at 0 even for
    change-point kernels whose list starts with the change points (dead `hasattr` test, :80-83); the
    evident intent - skip them, as SegmentedCovarianceMatrix does (CovarianceMatrix.py:319-320) - is implemented."""

    def __init__(self, _gp, local_approx, numerical_matrix_handling, subset_size: int = None):
        if local_approx is not mht.MatrixApproximations.NONE or \
                numerical_matrix_handling is not mht.NumericalMatrixHandlingType.CHOLESKY_BASED:
            raise NotImplementedError("only the exact Cholesky-based path is implemented on the B200 path")
        self.local_approx = local_approx
        self.numerical_matrix_handling = numerical_matrix_handling
        self.subset_size = subset_size
        self._gp = _gp
        self.last_block_values = None

    def _eval(self, hyper_parameter, noise, want_grad):
        cov = self._gp.covariance_matrix
        per_block, grads, gnoise = cov.block_nll_and_grad(hyper_parameter, noise, want_grad)
        self.last_block_values = per_block
        return float(sum(v for v in per_block if v is not None)), grads, gnoise

    def get_metric(self, hyper_parameter, noise, indices=None, reset: bool = True) -> torch.Tensor:
        return _evaluate(self._eval, hyper_parameter, noise)

    def get_gradients(self, hyper_parameter, noise, reset: bool = True, with_noise: bool = False):
        _, grads, gnoise = self._eval(hyper_parameter, noise, True)
        return (grads, torch.tensor(gnoise, dtype=torch.float64)) if with_noise else grads
