"""Mean-function base class (mirror of gpbasics/MeanFunctionBasics/MeanFunction.py:9-77).  Mean functions detrend y on
the host before the likelihood path starts (DataHandling/DataInput.py:77-97); they are O(n) and outside the hot path."""
from enum import Enum
from typing import List

import torch

from ..Auxiliary import BasicGPComponent as bgpc


class ConstantHyperParamType(Enum):
    NONE_CONSTANT = 0
    ALL_CONSTANT = 3


class MeanFunctionType(Enum):
    BASE_MEAN_FUNCTION = 1
    OPERATOR = 2


class MeanFunctionManifestation(Enum):
    C = 101
    LIN = 102
    EXP = 103
    LOGIT = 104
    ADD = 201
    MUL = 202
    CP = 203


class MeanFunction(bgpc.Component):
    def __init__(self, mean_function_type, manifestation, input_dimensionality: int):
        assert input_dimensionality >= 1, "input_dimensionality for a mean function ought to be 1 or larger"
        self.type = mean_function_type
        self.manifestation = manifestation
        self.last_hyper_parameter: List[torch.Tensor] = None
        self.input_dimensionality = input_dimensionality

    def get_tf_tensor(self, hyper_parameter, x_vector) -> torch.Tensor:
        raise NotImplementedError

    def get_mean_function_type(self):
        return self.type

    def get_mean_function_manifestation(self):
        return self.manifestation

    def get_number_of_hyper_parameter(self) -> int:
        raise NotImplementedError

    def get_string_representation(self) -> str:
        raise NotImplementedError

    def set_last_hyper_parameter(self, last_hyper_parameter):
        self.last_hyper_parameter = last_hyper_parameter

    def get_last_hyper_parameter(self):
        return self.last_hyper_parameter

    def deepcopy(self):
        raise NotImplementedError

    def get_default_hyper_parameter(self):
        raise NotImplementedError
