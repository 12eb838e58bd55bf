"""Abstract kernel, enums and shared host-side state (mirror of gpbasics/KernelBasics/Kernel.py:9-140).

A kernel object is a host-side description only.  Its arithmetic lives in the CUDA interpreter: `to_spec()` turns the
tree into the tuple form that gaussianprocessfundamentals_b200.program compiles, and `get_tf_tensor` (name kept for
drop-in compatibility, Kernel.py:51) launches the fused assembly kernel.
"""
from enum import Enum
from typing import List

import numpy as np
import torch

from ..Auxiliary import BasicGPComponent as bgpc
from .. import global_parameters as global_param


class ConstantHyperParamType(Enum):
    NONE_CONSTANT = 0
    JUST_CONSTANT_BASE_KERNELS = 1
    JUST_CONSTANT_CP = 2
    ALL_CONSTANT = 3


class KernelType(Enum):
    BASE_KERNEL = 1
    OPERATOR = 2


class KernelManifestation(Enum):
    C = 101
    LIN = 102
    RQ = 103
    PER = 104
    SE = 105
    WN = 106
    MAT32 = 107
    MAT52 = 108
    SE_ARD = 109  # extension (not in the reference)

    ADD = 201
    MUL = 202
    CP = 203
    PART = 204


def as_scalar_tensor(v) -> torch.Tensor:
    return torch.as_tensor(v, dtype=torch.float64).detach().clone()


class Kernel(bgpc.Component):
    def __init__(self, kernel_type: KernelType, manifestation: KernelManifestation, input_dimensionality: int):
        assert input_dimensionality >= 1, "input_dimensionality for a kernel ought to be one or larger"
        self.kernel_type = kernel_type
        self.manifestation = manifestation
        self.last_hyper_parameter: List[torch.Tensor] = None
        self.input_dimensionality = int(input_dimensionality)
        self.noise: torch.Tensor = None

    # ---- device evaluation ------------------------------------------------------------------------------------
    def to_spec(self):
        """tuple tree consumed by program.compile_spec"""
        raise NotImplementedError

    def get_tf_tensor(self, hyper_parameter: List[torch.Tensor], x_vector, x_vector_) -> torch.Tensor:
        """K[n, m] = k(x_vector, x_vector_) on the device; same contract as Kernel.get_tf_tensor (Kernel.py:51)."""
        from .. import engine
        from ..program import flatten_hp
        assert x_vector is not None and x_vector_ is not None, "Input vectors x and x_ uninitialized: " + str(self)
        assert len(hyper_parameter) == self.get_number_of_hyper_parameter(), "Invalid hyper_param size: " + str(self)
        engine.require_cuda()
        prog = engine.DeviceProgram.get(self.to_spec(), self.input_dimensionality, global_param.p_scaled_base_kernel,
                                        global_param.cp_mode_code())
        same = x_vector is x_vector_
        X = torch.as_tensor(x_vector, dtype=torch.float64).cuda().contiguous()
        X2 = None if same else torch.as_tensor(x_vector_, dtype=torch.float64).cuda().contiguous()
        hp = torch.as_tensor(flatten_hp(prog.compiled.entries, hyper_parameter, prog.n_hp)).cuda()
        K = engine.assemble(prog, X, X2, hp, None)
        self._remember(hyper_parameter)
        return K

    def _remember(self, hyper_parameter):
        pass

    # ---- reference API ----------------------------------------------------------------------------------------
    def get_kernel_type(self) -> KernelType:
        return self.kernel_type

    def get_kernel_manifestation(self) -> KernelManifestation:
        return self.manifestation

    def get_number_of_hyper_parameter(self) -> int:
        raise NotImplementedError

    def get_string_representation(self) -> str:
        raise NotImplementedError

    def get_number_base_kernels(self) -> int:
        raise NotImplementedError

    def get_default_hyper_parameter(self, xrange: List[List[float]], n: int, from_distribution: bool = False):
        raise NotImplementedError

    def set_last_hyper_parameter(self, last_hyper_parameter: List[torch.Tensor]):
        raise NotImplementedError

    def get_last_hyper_parameter(self, scaling_x_param=None):
        raise NotImplementedError

    def set_noise(self, noise):
        noise = torch.as_tensor(noise, dtype=torch.float64)
        if noise.dim() == 0:
            self.noise = noise
        else:
            raise Exception("Invalid Noise set for Kernel")

    def get_noise(self):
        return self.noise

    def deepcopy(self):
        raise NotImplementedError

    def get_string_representation_weight(self) -> float:
        return 0

    def sort_child_nodes(self):
        pass

    def get_json(self) -> dict:
        raise NotImplementedError

    def get_number_of_child_nodes(self) -> int:
        raise NotImplementedError

    def get_derivative_matrices(self, hyper_parameter, x_vector, x_vector_):
        """The reference's analytic derivative matrices are never called and two of them are wrong.
        Gradients are produced by the fused trace kernel instead (Metric.get_gradients)."""
        raise NotImplementedError("derivative matrices are not materialised; use Metric.get_gradients")

    def get_hyper_parameter_names(self, kernel_id: int = -1) -> List[str]:
        raise NotImplementedError

    def get_dimensionality(self):
        return self.input_dimensionality

    def set_dimensionality(self, input_dimensionality: int):
        self.input_dimensionality = int(input_dimensionality)

    def get_simplified_version(self):
        return self

    def type_compare_to(self, other):
        return self == other

    def get_hash_tuple(self):
        plain = []
        if isinstance(self.last_hyper_parameter, list):
            for hyp in self.last_hyper_parameter:
                v = np.asarray(torch.as_tensor(hyp).detach().cpu().numpy()).tolist()
                if isinstance(v, float):
                    plain.append(v)
                elif isinstance(v, list):
                    plain.extend(v)
        if self.noise is None:
            return self.manifestation.value, None, tuple(plain)
        return self.manifestation.value, float(self.noise), tuple(plain)

    def __hash__(self):
        return hash(self.get_hash_tuple())
