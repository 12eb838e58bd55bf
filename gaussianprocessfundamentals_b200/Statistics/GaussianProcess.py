"""Gaussian-process objects (mirror of gpbasics/Statistics/GaussianProcess.py:19-201): wiring of kernel, mean
function, covariance matrix and auxiliary properties; holistic, block-wise (change points) and partitioned variants
with their `constituent_gps`."""
import logging
from typing import List, Tuple

import numpy as np
import torch

from .. import global_parameters as global_param
from ..KernelBasics import Operators as op
from . import Auxiliary as ax
from . import CovarianceMatrix as cm

global_param.ensure_init()


class AbstractGaussianProcess:
    def __init__(self, kernel, mean_function):
        self.mean_function = mean_function
        self.kernel = kernel
        self.covariance_matrix: cm.CovarianceMatrix = None
        self.aux: ax.AuxiliaryGpProperties = None
        self.data_input = None
        self.inducing_points = None

    def set_inducing_points(self, inducing_points):
        self.inducing_points = inducing_points

    def set_data_input(self, data_input):
        self.data_input = data_input
        self.covariance_matrix.set_data_input(data_input)
        self.aux.set_data_input(data_input)

    def predict(self, kernel_hyper_param=None, mean_function_hyper_param=None, noise=None) \
            -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(mean + posterior mean, mean, posterior mean) at the test inputs (GaussianProcess.py:42-85).  The reference's
        block-wise branch reads a non-existent attribute (:72); the evident intent - skip the leading
        change points of a CP kernel - is implemented."""
        self.aux.reset()
        self.covariance_matrix.reset()
        if noise is None:
            noise = global_param.p_cov_matrix_jitter
        if mean_function_hyper_param is None:
            mean_function_hyper_param = self.mean_function.get_last_hyper_parameter() or \
                self.mean_function.get_default_hyper_parameter()
        if kernel_hyper_param is None:
            kernel_hyper_param = self.kernel.get_last_hyper_parameter()
            if kernel_hyper_param is None:
                kernel_hyper_param = self.kernel.get_default_hyper_parameter(self.data_input.get_x_range(),
                                                                             self.data_input.n_train)
        mean_mu = self.mean_function.get_tf_tensor(mean_function_hyper_param, self.data_input.data_x_test).cuda()
        if isinstance(self, (PartitionedGaussianProcess, BlockwiseGaussianProcess)):
            index = len(self.kernel.change_point_positions) if isinstance(self.kernel, op.ChangePointOperator) else 0
            mus = []
            for sub_gp in self.constituent_gps:
                c = sub_gp.kernel.get_number_of_hyper_parameter()
                n_test = int(sub_gp.data_input.n_test)
                if sub_gp.data_input.n_train > 0 and n_test > 0:
                    sub_gp.aux.reset(); sub_gp.covariance_matrix.reset()
                    mus.append(sub_gp.aux.get_posterior_mu(list(kernel_hyper_param[index:index + c]), noise))
                elif n_test > 0:
                    # test points in a block without training points: the posterior is the prior, whose mean is 0 (the
                    # reference cannot build such a block at all); keeps the concatenation aligned with data_x_test
                    mus.append(torch.zeros(n_test, dtype=torch.float64, device="cuda"))
                index += c
            posterior_mu = torch.cat(mus, dim=0)
        else:
            posterior_mu = self.aux.get_posterior_mu(kernel_hyper_param, noise)
        return mean_mu + posterior_mu, mean_mu, posterior_mu

    def copy(self):
        raise NotImplementedError


class GaussianProcess(AbstractGaussianProcess):
    def __init__(self, kernel, mean_function):
        super().__init__(kernel, mean_function)
        self.covariance_matrix = cm.HolisticCovarianceMatrix(self.kernel)
        self.aux = ax.HolisticAuxiliaryGpProperties(self.covariance_matrix, self.mean_function)

    def copy(self):
        gp = GaussianProcess(self.kernel, self.mean_function)
        gp.set_inducing_points(self.inducing_points)
        return gp


class PredefinedGaussianProcess(AbstractGaussianProcess):
    def __init__(self, covariance_matrix, mean_function):
        super().__init__(covariance_matrix.kernel, mean_function)
        self.covariance_matrix = covariance_matrix
        self.aux = ax.HolisticAuxiliaryGpProperties(self.covariance_matrix, self.mean_function)

    def copy(self):
        gp = PredefinedGaussianProcess(self.covariance_matrix, self.mean_function)
        gp.set_inducing_points(self.inducing_points)
        return gp


class _SegmentedGaussianProcess(AbstractGaussianProcess):
    def __init__(self, kernel, mean_function):
        super().__init__(kernel, mean_function)
        self.constituent_gps: List[GaussianProcess] = [GaussianProcess(cn, self.mean_function)
                                                       for cn in kernel.child_nodes]
        self.covariance_matrix = cm.SegmentedCovarianceMatrix(kernel)
        self.aux = ax.BlockwiseAuxiliaryGpProperties(self.covariance_matrix, self.mean_function)

    def set_data_input(self, data_input):
        assert len(data_input.data_inputs) == len(self.constituent_gps), \
            "Data Input does not fit constituent GPs of the segmented GP"
        self.data_input = data_input
        self.covariance_matrix.set_data_input(data_input)
        self.aux.set_data_input(data_input)
        for sub_gp, blk in zip(self.constituent_gps, data_input.data_inputs):
            sub_gp.set_data_input(blk)

    def copy(self):
        gp = type(self)(self.kernel, self.mean_function)
        gp.set_inducing_points(self.inducing_points)
        return gp


class BlockwiseGaussianProcess(_SegmentedGaussianProcess):
    """top-level ChangePointOperator kernel over a BlockwiseDataInput (GaussianProcess.py:140-169)"""


class PartitionedGaussianProcess(_SegmentedGaussianProcess):
    """top-level PartitionOperator kernel over a PartitionedDataInput (GaussianProcess.py:172-201)"""
