"""BIC = 2 NLL + |theta| log n on top of the device likelihood (mirror of
gpbasics/Metrics/BayesianInformationCriterion.py:18-63).  Host-side add-on."""
import math

import torch

from .LogLikelihood import AbstractLogLikelihood, BlockwiseLogLikelihood
from .Metrics import AbstractMetric, Metric, MetricType


class AbstractBIC(Metric):
    pass


class BIC(AbstractBIC):
    def __init__(self, data_input, covariance_matrix, log_likelihood: AbstractLogLikelihood):
        super().__init__(data_input, covariance_matrix, MetricType.BIC, log_likelihood.local_approx,
                         log_likelihood.numerical_matrix_handling, log_likelihood.subset_size)
        self.log_likelihood = log_likelihood

    def get_metric(self, hyper_parameter, noise, indices=None, reset: bool = True) -> torch.Tensor:
        nll = self.log_likelihood.get_metric(hyper_parameter, noise, indices, reset)
        penalty = self.covariance_matrix.kernel.get_number_of_hyper_parameter() * math.log(self.data_input.n_train)
        return 2 * nll + penalty

    def _eval(self, hyper_parameter, noise, want_grad):
        """(BIC, [d BIC / d hp], d BIC / d noise): the penalty is constant, so the gradient is twice the likelihood's
        (what the reference's GradientTape yields when a fitter is built with MetricType.BIC, Optimizer/Fitter.py:154-158)"""
        value, grads, gnoise = self.log_likelihood._eval(hyper_parameter, noise, want_grad)
        penalty = self.covariance_matrix.kernel.get_number_of_hyper_parameter() * math.log(self.data_input.n_train)
        if not want_grad:
            return 2 * value + penalty, None, None
        return 2 * value + penalty, [2 * torch.as_tensor(g, dtype=torch.float64) for g in grads], 2 * gnoise

    def get_gradients(self, hyper_parameter, noise, reset: bool = True, with_noise: bool = False):
        if reset:
            self.covariance_matrix.reset()
        _, grads, gnoise = self._eval(hyper_parameter, noise, True)
        return (grads, torch.tensor(gnoise, dtype=torch.float64)) if with_noise else grads


class BlockwiseBIC(AbstractMetric):
    def __init__(self, _gp, local_approx, numerical_matrix_handling, subset_size: int = None):
        self.local_approx = local_approx
        self.numerical_matrix_handling = numerical_matrix_handling
        self.subset_size = subset_size
        self._gp = _gp

    def get_metric(self, hyper_parameter, noise, indices=None) -> torch.Tensor:
        ll = BlockwiseLogLikelihood(self._gp, self.local_approx, self.numerical_matrix_handling, self.subset_size)
        nll = ll.get_metric(hyper_parameter, noise, indices)
        penalty = self._gp.covariance_matrix.kernel.get_number_of_hyper_parameter() * math.log(self._gp.data_input.n_train)
        return 2 * nll + penalty

    def _eval(self, hyper_parameter, noise, want_grad):
        ll = BlockwiseLogLikelihood(self._gp, self.local_approx, self.numerical_matrix_handling, self.subset_size)
        value, grads, gnoise = ll._eval(hyper_parameter, noise, want_grad)
        penalty = self._gp.covariance_matrix.kernel.get_number_of_hyper_parameter() * math.log(self._gp.data_input.n_train)
        if not want_grad:
            return 2 * value + penalty, None, None
        return 2 * value + penalty, [2 * torch.as_tensor(g, dtype=torch.float64) for g in grads], 2 * gnoise

    def get_gradients(self, hyper_parameter, noise, reset: bool = True, with_noise: bool = False):
        _, grads, gnoise = self._eval(hyper_parameter, noise, True)
        return (grads, torch.tensor(gnoise, dtype=torch.float64)) if with_noise else grads
