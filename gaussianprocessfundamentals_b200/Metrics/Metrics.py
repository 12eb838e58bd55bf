"""Metric base classes (mirror of gpbasics/Metrics/Metrics.py:17-154).

Only the exact path is implemented on the device: MatrixApproximations.NONE with
NumericalMatrixHandlingType.CHOLESKY_BASED, where `get_alpha` is CovarianceMatrix.get_L_alpha and `get_log_determinant`
is 2 * sum(log(diag L)) (Metrics.py:138-139, :152-154).  The approximations (Nystroem, SKC bounds, SKI, subset of data)
and the inverse / pinv / CG handlers are outside the north-star path and raise NotImplementedError."""
from enum import Enum
from typing import List

import torch

from .. import global_parameters as global_param
from . import MatrixHandlingTypes as mht

global_param.ensure_init()


class MetricType(Enum):
    LL = 1
    MSE = 5
    BIC = 6
    blockwise_LL = 10
    blockwise_MSE = 50
    blockwise_BIC = 60


class AbstractMetric:
    # a metric is given in a form having: optimum = minimum
    def get_metric(self, hyper_parameter: List[torch.Tensor], noise, indices=None) -> torch.Tensor:
        raise NotImplementedError

    def get_gradients(self, hyper_parameter: List[torch.Tensor], noise, reset: bool = True):
        """d metric / d hyper_parameter (list shaped like `hyper_parameter`).  Declared but never implemented in the
        reference (Metrics.py:31-32); here it is the fused trace-gradient of the CUDA path."""
        raise NotImplementedError


class Metric(AbstractMetric):
    def __init__(self, data_input, covariance_matrix, metric_type: MetricType, local_approx, numerical_matrix_handling,
                 subset_size: int = None):
        if local_approx is not mht.MatrixApproximations.NONE:
            raise NotImplementedError("only the exact GP (MatrixApproximations.NONE) is implemented on the B200 path")
        if numerical_matrix_handling is not mht.NumericalMatrixHandlingType.CHOLESKY_BASED:
            raise NotImplementedError("only NumericalMatrixHandlingType.CHOLESKY_BASED is implemented on the B200 path")
        self.covariance_matrix = covariance_matrix
        self.local_approx = local_approx
        self.numerical_matrix_handling = numerical_matrix_handling
        self.subset_size = subset_size
        self.data_input = data_input
        self.covariance_matrix.set_data_input(self.data_input)
        self.type = metric_type
        self.last_covariance_matrix = None

    def get_covariance_matrix(self, hyper_parameter, noise, indices=None):
        if self.last_covariance_matrix is None:
            self.last_covariance_matrix = self.covariance_matrix.get_K_noised(hyper_parameter, noise)
        return self.last_covariance_matrix

    # The reference binds get_alpha / get_log_determinant to one of several handlers in its constructor
    # (Metrics.py:77-107); with CHOLESKY_BASED - the only handler of the B200 path - these are the two below.
    def get_alpha_cholesky(self, hyper_parameter, noise, y=None, indices=None):
        """alpha = L^T \\ (L \\ y) (Metrics.py:138-139): blocked Cholesky with carried y + blocked back substitution"""
        return self.covariance_matrix.get_L_alpha(hyper_parameter, noise)

    def get_log_determinant_cholesky(self, hyper_parameter, noise, indices=None):
        """log det (K + s2 I) = 2 sum log diag L (Metrics.py:152-154)"""
        L = self.covariance_matrix.get_L_K(hyper_parameter, noise)
        return 2 * torch.sum(torch.log(torch.diagonal(L)))

    def get_default_covariance_matrix(self, hyper_parameter, noise, indices=None):
        return self.get_covariance_matrix(hyper_parameter, noise, indices)

    def get_alpha(self, hyper_parameter, noise, y=None, indices=None):
        return self.get_alpha_cholesky(hyper_parameter, noise, y, indices)

    def get_log_determinant(self, hyper_parameter, noise, indices=None):
        return self.get_log_determinant_cholesky(hyper_parameter, noise, indices)
