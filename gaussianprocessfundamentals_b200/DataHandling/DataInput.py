"""DataInput and its partitioned / block-wise containers (mirror of gpbasics/DataHandling/DataInput.py:26-253).

The segment rule of BlockwiseDataInput is integer bookkeeping that must match the reference bit for bit
(DataInput.py:231-244): segment 0 = {x < cp_0}, segment i = {cp_{i-1} <= x < cp_i}, last = {x >= cp_last}, row indices
ascending, 1-d inputs only."""
import logging
from typing import List

import numpy as np
import torch

from .. import global_parameters as global_param
from ..MeanFunctionBasics import BaseMeanFunctions as bmf
from ..Metrics import MatrixHandlingTypes as mht
from .AbstractDataInput import AbstractDataInput, _t

global_param.ensure_init()


def is_equidistant(input_vector) -> bool:
    v = np.asarray(input_vector, dtype=np.float64)
    diff = v[:len(v) - 1] - v[1:]
    mean = np.mean(diff)
    allowed_error = 1 / (100 * len(v))
    return bool(np.abs(np.max(diff) - mean) < allowed_error and np.abs(np.min(diff) - mean) < allowed_error)


class DataInput(AbstractDataInput):
    def __init__(self, data_x_train, data_y_train, data_x_test=None, data_y_test=None, test_ratio: float = -1,
                 seed: int = 3061941):
        if data_x_test is None or data_y_test is None:
            data_x_test, data_y_test = (None, None) if data_x_test is None else (data_x_test, data_y_test)
        super().__init__(data_x_train, data_y_train, data_x_test, data_y_test, test_ratio, seed)

    def get_inducting_x_train(self, indices) -> torch.Tensor:
        self.inducting_x_train = self.data_x_train[torch.as_tensor(indices, dtype=torch.long)]
        return self.inducting_x_train

    def get_inducting_x_test(self, indices) -> torch.Tensor:
        self.inducting_x_test = self.data_x_test[torch.as_tensor(indices, dtype=torch.long)]
        return self.inducting_x_test

    def get_x_range(self) -> List[List[float]]:
        out = []
        for d in range(self.get_input_dimensionality()):
            lo = float(min(self.data_x_train[:, d].min(), self.data_x_test[:, d].min()))
            hi = float(max(self.data_x_train[:, d].max(), self.data_x_test[:, d].max()))
            out.append([lo, hi])
        return out

    def _detrend(self, x, y):
        if isinstance(self.mean_function, bmf.ZeroMeanFunction):
            return y.clone()
        mean = self.mean_function.get_tf_tensor(self.mean_function.get_last_hyper_parameter(), x).reshape(-1, 1)
        return y - mean

    def get_detrended_y_train(self) -> torch.Tensor:
        if self.mean_function is None:
            logging.error("Mean Function is None.")
            return None
        if self.detrended_y_train is None:
            self.detrended_y_train = self._detrend(self.data_x_train, self.data_y_train)
        return self.detrended_y_train

    def get_detrended_y_test(self) -> torch.Tensor:
        if self.mean_function is None:
            logging.error("Mean Function is None.")
            return None
        if self.detrended_y_test is None:
            self.detrended_y_test = self._detrend(self.data_x_test, self.data_y_test)
        return self.detrended_y_test

    def get_detrended_y_test_individual(self, data_x_test, data_y_test) -> torch.Tensor:
        return self._detrend(_t(data_x_test), _t(data_y_test))

    def _subset(self, idx: torch.Tensor):
        separate = not torch.equal(self.data_x_train, self.data_x_test)
        x_te, y_te = (self.data_x_test, self.data_y_test) if separate else (self.data_x_train, self.data_y_train)
        sub = DataInput(self.data_x_train[idx], self.data_y_train[idx], x_te, y_te)
        sub.set_mean_function(self.mean_function)
        return sub

    def get_random_subset(self, subset_size: int):
        g = torch.Generator().manual_seed(int(self.seed))
        idx = torch.sort(torch.randint(0, self.n_train, (subset_size,), generator=g)).values
        return self._subset(idx)

    def get_grid_subset(self, subset_size: int):
        idx = torch.as_tensor(np.linspace(start=0, stop=self.n_train, num=subset_size, endpoint=False, dtype=int))
        return self._subset(idx)

    def is_equidistant_input_x(self) -> bool:
        return is_equidistant(self.data_x_train.numpy())

    def get_subset(self, subset_size: int, subset_of_data_approach):
        if subset_of_data_approach is mht.SubsetOfDataApproaches.SOD_GRID:
            return self.get_grid_subset(subset_size)
        if subset_of_data_approach is mht.SubsetOfDataApproaches.SOD_RANDOM:
            return self.get_random_subset(subset_size)
        raise Exception("Invalid subset-of-data approach: %s" % str(subset_of_data_approach))

    @staticmethod
    def get_k_fold_data_inputs(data_x_train, data_y_train, k: int, seed: int = 3061941):
        x, y = _t(data_x_train), _t(data_y_train)
        assert x.dim() == 2 and y.dim() == 2, "Only non-batched Data is valid Input."
        folds = AbstractDataInput.get_k_fold_data_inputs(x, y, k, seed)
        return [DataInput(f.data_x_train, f.data_y_train, f.data_x_test, f.data_y_test, seed=f.seed) for f in folds]


class PartitionedDataInput(DataInput):
    def __init__(self, data_x_train, data_y_train, data_x_test, data_y_test, data_inputs: List[DataInput]):
        super().__init__(data_x_train, data_y_train, data_x_test, data_y_test)
        self.data_inputs: List[DataInput] = data_inputs

    def set_mean_function(self, mean_function):
        super().set_mean_function(mean_function)
        for data_input in self.data_inputs:
            data_input.set_mean_function(mean_function)


def blockwise_segment_indices(x: torch.Tensor, change_points) -> List[torch.Tensor]:
    """ascending row indices of each change-point segment of a 1-d input column (DataInput.py:231-244)"""
    col = x.reshape(x.shape[0], -1)[:, 0]
    cps = [float(torch.as_tensor(c, dtype=torch.float64).reshape(-1)[0]) for c in change_points]
    out = []
    for i in range(len(cps) + 1):
        if i == 0:
            mask = col < cps[0]
        elif i == len(cps):
            mask = col >= cps[i - 1]
        else:
            mask = torch.logical_and(col < cps[i], col >= cps[i - 1])
        out.append(torch.nonzero(mask)[:, 0])
    return out


class BlockwiseDataInput(PartitionedDataInput):
    def __init__(self, data_x_train, data_y_train, data_x_test, data_y_test, change_points):
        x_tr, y_tr, x_te, y_te = _t(data_x_train), _t(data_y_train), _t(data_x_test), _t(data_y_test)
        tr_idx = blockwise_segment_indices(x_tr, change_points)
        te_idx = blockwise_segment_indices(x_te, change_points)
        blocks = [DataInput(x_tr[a], y_tr[a], x_te[b], y_te[b]) for a, b in zip(tr_idx, te_idx)]
        self.train_indices, self.test_indices = tr_idx, te_idx
        super().__init__(x_tr, y_tr, x_te, y_te, blocks)
