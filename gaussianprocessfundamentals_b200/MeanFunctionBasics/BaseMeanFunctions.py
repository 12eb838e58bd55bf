"""Zero / constant / linear mean functions (mirror of gpbasics/MeanFunctionBasics/BaseMeanFunctions.py:37-120).
Only ZeroMeanFunction is on the likelihood path's fast path (DataInput.py:86-87: y passes through unchanged); the other
two are provided so that detrending call sites keep working.  Host-side O(n) arithmetic."""
from typing import List

import torch

from . import MeanFunction as mf


class BaseMeanFunction(mf.MeanFunction):
    def __init__(self, manifestation, input_dimensionality: int):
        super().__init__(mf.MeanFunctionType.BASE_MEAN_FUNCTION, manifestation, input_dimensionality)

    def get_number_base_mean_function(self) -> int:
        return 1

    def set_last_hyper_parameter(self, last_hyper_parameter: List[torch.Tensor]):
        assert len(last_hyper_parameter) == self.get_number_of_hyper_parameter(), \
            "Wrong size/shape of given 'last_hyper_param'"
        self.last_hyper_parameter = last_hyper_parameter

    def get_number_of_hyper_parameter(self) -> int:
        return len(self.get_default_hyper_parameter())

    def get_string_representation(self) -> str:
        return self.manifestation.name

    def get_string_representation_weight(self) -> int:
        return self.manifestation.value - 100


class ConstantMeanFunction(BaseMeanFunction):
    def __init__(self, input_dimensionality: int):
        super().__init__(mf.MeanFunctionManifestation.C, input_dimensionality)

    def get_tf_tensor(self, hyper_parameter, x_vector) -> torch.Tensor:
        assert x_vector is not None, "Input vector x uninitialized: " + str(self)
        assert len(hyper_parameter) == self.get_number_of_hyper_parameter(), "Invalid hyper_param size: " + str(self)
        x = torch.as_tensor(x_vector, dtype=torch.float64)
        self.last_hyper_parameter = hyper_parameter
        return torch.zeros(x.shape[0], dtype=torch.float64) + torch.as_tensor(hyper_parameter[0], dtype=torch.float64)

    def get_default_hyper_parameter(self):
        return [torch.tensor(0.01, dtype=torch.float64)]

    def get_hyper_parameter_dimensionalities(self):
        return [[]]

    def deepcopy(self):
        other = type(self)(self.input_dimensionality)
        if self.last_hyper_parameter is not None:
            other.set_last_hyper_parameter(self.last_hyper_parameter)
        return other


class ZeroMeanFunction(ConstantMeanFunction):
    def get_string_representation(self) -> str:
        return "ZERO_MEAN"

    def get_default_hyper_parameter(self):
        return [torch.tensor(0.0, dtype=torch.float64)]

    def deepcopy(self):
        return ZeroMeanFunction(self.input_dimensionality)


class LinearMeanFunction(BaseMeanFunction):
    """m(x) = sum_d a_d x_d + b"""

    def __init__(self, input_dimensionality: int):
        super().__init__(mf.MeanFunctionManifestation.LIN, input_dimensionality)

    def get_tf_tensor(self, hyper_parameter, x_vector) -> torch.Tensor:
        assert len(hyper_parameter) == self.get_number_of_hyper_parameter(), "Invalid hyper_param size: " + str(self)
        x = torch.as_tensor(x_vector, dtype=torch.float64)
        a = torch.as_tensor(hyper_parameter[0], dtype=torch.float64).reshape(-1)
        b = torch.as_tensor(hyper_parameter[1], dtype=torch.float64)
        self.last_hyper_parameter = hyper_parameter
        return (x * a).sum(-1) + b

    def get_default_hyper_parameter(self):
        return [torch.full((self.input_dimensionality,), 0.01, dtype=torch.float64), torch.tensor(0.01, dtype=torch.float64)]

    def get_hyper_parameter_dimensionalities(self):
        return [[self.input_dimensionality], []]

    def deepcopy(self):
        other = LinearMeanFunction(self.input_dimensionality)
        if self.last_hyper_parameter is not None:
            other.set_last_hyper_parameter(self.last_hyper_parameter)
        return other
