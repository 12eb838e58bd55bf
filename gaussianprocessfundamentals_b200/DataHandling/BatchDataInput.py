"""Rank-3 same-kernel batches (mirror of gpbasics/DataHandling/BatchDataInput.py:24-112): X [B, n, d], y [B, n, 1]."""
from typing import List

import torch

from .. import global_parameters as global_param
from ..MeanFunctionBasics import BaseMeanFunctions as bmf
from .AbstractDataInput import AbstractDataInput, _t

global_param.ensure_init()


def is_equidistant(input_data: torch.Tensor):
    length = input_data.shape[1]
    diff = input_data[:, :length - 1, :] - input_data[:, 1:, :]
    mean = diff.mean(dim=0)
    allowed = 1 / (100 * length)
    return torch.logical_and(torch.abs(diff.max(dim=1).values - mean) < allowed,
                             torch.abs(diff.min(dim=1).values - mean) < allowed)


class BatchDataInput(AbstractDataInput):
    def get_x_range(self) -> List[List[float]]:
        out = []
        for d in range(self.get_input_dimensionality()):
            both = torch.cat([self.data_x_train[:, :, d], self.data_x_test[:, :, d]], dim=0)
            out.append([float(both.min()), float(both.max())])
        return out

    def _detrend(self, x, y):
        if isinstance(self.mean_function, bmf.ZeroMeanFunction):
            return y
        hp = self.mean_function.get_last_hyper_parameter()
        mean = torch.stack([self.mean_function.get_tf_tensor(hp, xb).reshape(-1, 1) for xb in x])
        return y - mean

    def get_detrended_y_train(self):
        if self.mean_function is None:
            return None
        if self.detrended_y_train is None:
            self.detrended_y_train = self._detrend(self.data_x_train, self.data_y_train)
        return self.detrended_y_train

    def get_detrended_y_test(self):
        if self.mean_function is None:
            return None
        if self.detrended_y_test is None:
            self.detrended_y_test = self._detrend(self.data_x_test, self.data_y_test)
        return self.detrended_y_test

    def is_equidistant_input_x(self):
        return is_equidistant(self.data_x_train)

    def get_random_subset(self, subset_size: int):
        raise Exception("get_random_subset -- Not implemented for BatchDataInput.")

    def get_grid_subset(self, subset_size: int):
        raise Exception("get_grid_subset -- Not implemented for BatchDataInput.")

    def get_subset(self, subset_size: int, subset_of_data_approach):
        raise Exception("get_subset -- Not implemented for BatchDataInput.")

    @staticmethod
    def get_k_fold_data_inputs(data_x_train, data_y_train, k: int, seed: int = 3061941):
        x, y = _t(data_x_train), _t(data_y_train)
        assert x.dim() == 3 and y.dim() == 3, "Only Batched Data is valid Input."
        folds = AbstractDataInput.get_k_fold_data_inputs(x, y, k, seed)
        return [BatchDataInput(f.data_x_train, f.data_y_train, f.data_x_test, f.data_y_test, seed=f.seed) for f in folds]
