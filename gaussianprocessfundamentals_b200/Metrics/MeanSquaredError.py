"""Mean squared error of the posterior mean at the test inputs (mirror of gpbasics/Metrics/MeanSquaredError.py:18-81).

The caller right after the likelihood path: alpha comes from the blocked Cholesky with carried y and
the blocked back substitution, K_s from the fused assembly kernel, K_s^T alpha from the FP64 tensor-core GEMM."""
from typing import List

import torch

from .. import engine
from ..KernelBasics import Operators as op
from . import MatrixHandlingTypes as mht
from .Metrics import AbstractMetric, Metric, MetricType


class AbstractMSE(Metric):
    pass


class MeanSquaredError(AbstractMSE):
    def __init__(self, data_input, covariance_matrix, aux_gp, local_approx, numerical_matrix_handling,
                 subset_size: int = None):
        super().__init__(data_input, covariance_matrix, MetricType.MSE, local_approx, numerical_matrix_handling,
                         subset_size)
        self.aux_gp = aux_gp

    def get_metric(self, hyper_parameter: List[torch.Tensor], noise, indices=None) -> torch.Tensor:
        self.aux_gp.reset()
        self.covariance_matrix.reset()
        post_mu = self.get_posterior_mu(hyper_parameter, noise, indices).reshape(-1, 1)
        y_test = self.data_input.get_detrended_y_test().to(post_mu.device)
        return torch.mean((post_mu - y_test) ** 2)                          # MeanSquaredError.py:30

    def get_posterior_mu(self, hyper_parameter, noise, indices=None):
        alpha = self.get_alpha(hyper_parameter, noise, None, indices)       # Metrics.py:138-139
        K_s = self.covariance_matrix.get_K_s(hyper_parameter)
        return engine.matmul(K_s, alpha, trans_a=True).reshape(int(self.data_input.n_test))


class BlockwiseMeanSquaredError(AbstractMetric):
    """MSE over the concatenated per-block posterior means (MeanSquaredError.py:45-81).  The reference tests
    `hasattr(self._gp, 'change_point_positions')`, which is never true, and so starts slicing at 0 even for change-point
    kernels; the evident intent - skip the leading change points - is implemented.  Blocks without
    training or test points contribute nothing (the reference would fail on them)."""

    def __init__(self, _gp, local_approx, numerical_matrix_handling, subset_size: int = None):
        if local_approx is not mht.MatrixApproximations.NONE or \
                numerical_matrix_handling is not mht.NumericalMatrixHandlingType.CHOLESKY_BASED:
            raise NotImplementedError("only the exact Cholesky-based path is implemented on the B200 path")
        self.local_approx = local_approx
        self.numerical_matrix_handling = numerical_matrix_handling
        self.subset_size = subset_size
        self.aux_gp = _gp.aux
        self._gp = _gp
        self.data_input = _gp.data_input

    def get_metric(self, hyper_parameter, noise, indices=None) -> torch.Tensor:
        kernel = self._gp.covariance_matrix.kernel
        index = len(kernel.change_point_positions) if isinstance(kernel, op.ChangePointOperator) else 0
        mus, ys = [], []
        for sub_gp in self._gp.constituent_gps:
            c = sub_gp.covariance_matrix.kernel.get_number_of_hyper_parameter()
            if sub_gp.data_input.n_train > 0 and sub_gp.data_input.n_test > 0:
                sub = MeanSquaredError(sub_gp.data_input, sub_gp.covariance_matrix, sub_gp.aux, self.local_approx,
                                       self.numerical_matrix_handling, self.subset_size)
                sub.aux_gp.reset()
                sub.covariance_matrix.reset()
                mus.append(sub.get_posterior_mu(list(hyper_parameter[index:index + c]), noise, indices))
                ys.append(sub.data_input.get_detrended_y_test().reshape(-1))
            index += c
        post_mu = torch.cat(mus, dim=0).reshape(-1, 1)
        y_test = torch.cat(ys, dim=0).reshape(-1, 1).to(post_mu.device)
        return torch.mean((post_mu - y_test) ** 2)
