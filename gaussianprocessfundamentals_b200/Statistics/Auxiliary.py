"""Posterior mean / variance helpers (mirror of gpbasics/Statistics/Auxiliary.py:14-107).  Downstream of the likelihood
path; products run on the FP64 tensor-core GEMM, and the variance uses W = L^-1 from the
recursive-doubling inverse instead of the reference's explicit tf.linalg.inv(L) (Auxiliary.py:57-66)."""
from typing import List

import torch

from .. import engine
from .. import global_parameters as global_param

global_param.ensure_init()


class AuxiliaryGpProperties:
    def __init__(self, covariance_matrix, mean_function):
        self.covariance_matrix = covariance_matrix
        self.data_input = None
        self.mean_function = mean_function
        self.reset()
        self.detrended_y_test = None

    def reset(self):
        self.detrended_y_train = None
        self.inv_L_K_dot_K_s = None
        self.posterior_mu = None
        self.posterior_var = None
        self.posterior_sd = None

    def set_data_input(self, data_input):
        self.data_input = data_input
        self.reset()


class HolisticAuxiliaryGpProperties(AuxiliaryGpProperties):
    def get_inverse_cholesky_k_times_k_s(self, hyper_parameter, noise):
        if self.data_input is None:
            return None
        if self.inv_L_K_dot_K_s is None:
            W = self.covariance_matrix.get_L_inv_K(hyper_parameter, noise)
            self.inv_L_K_dot_K_s = engine.matmul(W, self.covariance_matrix.get_K_s(hyper_parameter))
        return self.inv_L_K_dot_K_s

    def get_posterior_mu(self, hyper_parameter, noise):
        if self.data_input is None:
            return None
        if self.posterior_mu is None:
            alpha = self.covariance_matrix.get_L_alpha(hyper_parameter, noise)
            K_s = self.covariance_matrix.get_K_s(hyper_parameter)
            self.posterior_mu = engine.matmul(K_s, alpha, trans_a=True).reshape(-1)
        return self.posterior_mu

    def get_posterior_var(self, hyper_parameter, noise):
        if self.data_input is None:
            return None
        if self.posterior_var is None:
            v = self.get_inverse_cholesky_k_times_k_s(hyper_parameter, noise)
            vtv = engine.matmul(v, v, trans_a=True)
            self.posterior_var = self.covariance_matrix.get_K_ss(hyper_parameter) - vtv
        return self.posterior_var

    def get_posterior_sd(self, hyper_parameter, noise):
        if self.data_input is None:
            return None
        if self.posterior_sd is None:
            self.posterior_sd = torch.sqrt(self.get_posterior_var(hyper_parameter, noise))
        return self.posterior_sd


class BlockwiseAuxiliaryGpProperties(HolisticAuxiliaryGpProperties):
    pass
