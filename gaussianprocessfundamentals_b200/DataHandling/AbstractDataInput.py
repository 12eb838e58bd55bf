"""Input contract of the likelihood path (mirror of gpbasics/DataHandling/AbstractDataInput.py:14-168).

Holds X / y (train and test) as float64 host tensors with the reference's shape rules ([n, d] / [n, 1], or rank 3 for
batches); the device copies the kernels read are created lazily and cached (`device_x_train`, `device_y_train`)."""
import logging
from typing import List

import numpy as np
import torch

from .. import global_parameters as global_param

global_param.ensure_init()


def _t(v) -> torch.Tensor:
    if isinstance(v, torch.Tensor):
        return v.detach().to("cpu", torch.float64)
    return torch.as_tensor(np.asarray(v, dtype=np.float64))


class AbstractDataInput:
    def __init__(self, data_x_train, data_y_train, data_x_test=None, data_y_test=None, test_ratio: float = -1,
                 seed: int = 3061941):
        data_x_train, data_y_train = _t(data_x_train), _t(data_y_train)
        if data_x_test is not None:
            data_x_test = _t(data_x_test)
        if data_y_test is not None:
            data_y_test = _t(data_y_test)
        assert (data_x_test is None or (data_x_train.dim() == data_x_test.dim()
                                        and (data_y_test is None or data_y_train.dim() == data_y_test.dim())
                                        and data_x_train.dim() == data_y_train.dim())) \
            and data_x_train.dim() in (2, 3), \
            "Shape of input and target data needs to be [instance#, length, dimensionality] or [length, dimensionality]."
        assert data_y_train.shape[-1] == 1, "Target training data (data_y_train) has to be unidimensional"
        assert data_y_test is None or data_y_test.shape[-1] == 1, "Target test data (data_y_test) has to be unidimensional"
        assert data_x_test is None or data_x_train.shape[-1] == data_x_test.shape[-1], \
            "Dimensionality of training and test input data need to match"
        assert test_ratio <= 1, "test_ratio has to be in the range [0; 1]"
        self.seed = seed
        if test_ratio > 0 and data_x_test is not None:
            logging.warning("test_ratio is ignored if test_data is explicitly given.")
        if data_x_test is None and test_ratio != 0:
            if test_ratio < 0:
                logging.warning("test_ratio is not given although explicit test data was not provided. "
                                "default value '0.2' is assumed for test_ratio.")
                test_ratio = 0.2
            length = data_x_train.shape[0]
            test_size = min(length - 1, int(length * test_ratio))
            train_size = length - test_size
            g = torch.Generator().manual_seed(int(seed))
            perm = torch.randperm(length, generator=g)
            idx_train = torch.sort(perm[:train_size]).values
            idx_test = torch.sort(perm[train_size:]).values
            self.data_x_train, self.data_y_train = data_x_train[idx_train], data_y_train[idx_train]
            self.data_x_test, self.data_y_test = data_x_train[idx_test], data_y_train[idx_test]
        elif data_x_test is None:
            self.data_x_train, self.data_y_train = data_x_train, data_y_train
            self.data_x_test, self.data_y_test = data_x_train, data_y_train
        else:
            self.data_x_train, self.data_y_train = data_x_train, data_y_train
            self.data_x_test, self.data_y_test = data_x_test, data_y_test
        self.detrended_y_test = None
        self.detrended_y_train = None
        self.mean_function = None
        self.n_train: int = self.data_x_train.shape[-2]
        self.n_test: int = self.data_x_test.shape[-2]
        self.inducting_x_train = None
        self.inducting_x_test = None
        inducting_min = 20
        self.n_inducting_train = max(inducting_min, int(self.n_train * global_param.p_nystroem_ratio))
        self.n_inducting_test = max(inducting_min, int(self.n_test * global_param.p_nystroem_ratio))
        self._dev = {}

    # ---- device residency (new: the reference keeps tf constants wherever TF places them) ------------------------
    def _device(self, name: str, tensor: torch.Tensor) -> torch.Tensor:
        from .. import engine
        engine.require_cuda()
        key = (name, tensor.data_ptr(), tuple(tensor.shape))
        cached = self._dev.get(name)
        if cached is None or cached[0] != key:
            cached = (key, tensor.contiguous().cuda())
            self._dev[name] = cached
        return cached[1]

    @property
    def device_x_train(self) -> torch.Tensor:
        return self._device("x_train", self.data_x_train)

    @property
    def device_x_test(self) -> torch.Tensor:
        return self._device("x_test", self.data_x_test)

    def get_input_dimensionality(self) -> int:
        return int(self.data_x_train.shape[-1])

    def set_seed(self, seed: int):
        self.seed = seed
        self.inducting_x_test = None
        self.inducting_x_train = None

    def set_mean_function(self, mean_function):
        self.detrended_y_train = None
        self.detrended_y_test = None
        self.mean_function = mean_function
        if self.mean_function.get_last_hyper_parameter() is None:
            self.mean_function.last_hyper_parameter = self.mean_function.get_default_hyper_parameter()

    def get_x_range(self) -> List[List[float]]:
        raise NotImplementedError

    def get_detrended_y_train(self) -> torch.Tensor:
        raise NotImplementedError

    def get_detrended_y_test(self) -> torch.Tensor:
        raise NotImplementedError

    @staticmethod
    def get_k_fold_data_inputs(x_train, y_train, k: int, seed: int = 3061941):
        x_train, y_train = _t(x_train), _t(y_train)
        length = x_train.shape[0]
        g = torch.Generator().manual_seed(int(seed))
        perm = torch.randperm(length, generator=g)
        sizes = [length // k] * (k - 1) + [length - (length // k) * (k - 1)]
        parts = torch.split(perm, sizes)
        out = []
        for i in range(k):
            idx_test = torch.sort(parts[i]).values
            idx_train = torch.sort(torch.cat([parts[j] for j in range(k) if j != i])).values
            out.append(AbstractDataInput(x_train[idx_train], y_train[idx_train], x_train[idx_test], y_train[idx_test],
                                         seed=seed))
        return out
