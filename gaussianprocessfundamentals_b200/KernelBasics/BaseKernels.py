"""Base kernels (mirror of gpbasics/KernelBasics/BaseKernels.py): LIN, SE, PER on the north-star path plus WN, MAT32,
MAT52 (and an SE-ARD extension) as further opcodes of the same fused assembly kernel.

The reference classes repeat the same bookkeeping per kernel; here one data-driven base class carries it and each
kernel declares only what differs: its opcode, its parameters, their defaults / priors / bounds and how a fitted value
is mapped back to data units.  Formulas (evaluated on the device, csrc/program.cuh gpb_leaf):
    LIN   (x-c).(x'-c)                          BaseKernels.py:119-130     hp [c (d,)]
    SE    exp(-1/2 r^2 / l^2)                   :282-290                   hp [l]
    PER   exp(-2 sin^2(pi L1 / p) / l^2)        :446-453                   hp [l, p]
    MAT32 (1+f) e^-f, f = sqrt3 L1/|l|          :707-716                   hp [l]
    MAT52 (1+f+5 L1^2/(3 l^2)) e^-f             :864-876                   hp [l]
    WN    1[i == j]                             :646-662                   no hp
every kernel gains a trailing scale `sg` under global_param.p_scaled_base_kernel (:127-128, :288-289, :452-453).
"""
import logging
import math
from typing import List, Tuple

import torch

from .. import global_parameters as global_param
from . import Kernel as k

global_param.ensure_init()

_INF = float("inf")


def _f64(v, shape=None):
    t = torch.as_tensor(v, dtype=torch.float64)
    if shape is not None:
        t = t.expand(shape).clone() if t.dim() == 0 and len(shape) > 0 else t.reshape(shape)
    return t


class BaseKernel(k.Kernel):
    SPEC = None          # opcode name understood by program.compile_spec
    MANIFESTATION = None
    PARAMS: Tuple[str, ...] = ()   # name suffixes of the un-scaled hyper-parameters, list order
    ABS_ON_SET: Tuple[int, ...] = ()  # entries stored as |.| after a fit (BaseKernels.py:429-432, :629-634)

    def __init__(self, input_dimensionality: int):
        super().__init__(k.KernelType.BASE_KERNEL, self.MANIFESTATION, input_dimensionality)
        if self.manifestation.value > 199:
            logging.critical("Invalid manifestation for BaseKernel: %s", self.manifestation)
        self.latest_cov_mat = None

    # ---- structure ----------------------------------------------------------------------------------------------
    def to_spec(self):
        return (self.SPEC,)

    def _scaled(self) -> bool:
        return bool(global_param.p_scaled_base_kernel) and len(self.PARAMS) > 0

    def _param_shape(self, idx: int) -> list:
        return []

    def get_number_base_kernels(self) -> int:
        return 1

    def get_number_of_child_nodes(self) -> int:
        return 1

    def get_number_of_hyper_parameter(self) -> int:
        return len(self.PARAMS) + (1 if self._scaled() else 0)

    def get_hyper_parameter_dimensionalities(self) -> List[list]:
        dims = [self._param_shape(i) for i in range(len(self.PARAMS))]
        if self._scaled():
            dims.append([])
        return dims

    def get_string_representation(self) -> str:
        return self.manifestation.name

    def get_string_representation_weight(self) -> int:
        return self.manifestation.value - 100

    def get_hyper_parameter_names(self, kernel_id: int = -1) -> List[str]:
        rep = self.get_string_representation()
        if kernel_id >= 0:
            rep += "_%i" % kernel_id
        names = [rep + "_" + p for p in self.PARAMS]
        if self._scaled():
            names.append(rep + "_sg")
        return names

    def type_compare_to(self, other):
        return isinstance(other, type(self))

    def get_json(self) -> dict:
        hp = self.get_last_hyper_parameter()
        return {"type": self.get_string_representation(),
                "hyper_param": [torch.as_tensor(h).detach().cpu().numpy().tolist() for h in hp]}

    def deepcopy(self):
        other = type(self)(input_dimensionality=self.input_dimensionality)
        if self.last_hyper_parameter is not None and self.get_number_of_hyper_parameter() > 0:
            other.set_last_hyper_parameter(list(self.last_hyper_parameter))
        if self.noise is not None:
            other.set_noise(self.noise)
        return other

    # ---- hyper-parameters ---------------------------------------------------------------------------------------
    def _remember(self, hyper_parameter):
        self.last_hyper_parameter = hyper_parameter

    def set_last_hyper_parameter(self, last_hyper_parameter: List[torch.Tensor]):
        if not isinstance(last_hyper_parameter, list):
            raise Exception("Wrong type for last_hyper_parameter to be set!")
        assert len(last_hyper_parameter) == self.get_number_of_hyper_parameter(), "Invalid hyper_param size: %s" % str(self)
        stored = list(last_hyper_parameter)
        for idx in self.ABS_ON_SET:
            stored[idx] = torch.abs(torch.as_tensor(stored[idx], dtype=torch.float64))
        self.last_hyper_parameter = stored

    def _to_data_units(self, hp, scaling_x_param):
        """fitted hp -> data units for inputs that were scaled as x' = (x - s0) / s1"""
        return list(hp[:len(self.PARAMS)])

    def get_last_hyper_parameter(self, scaling_x_param=None):
        result = self.last_hyper_parameter
        if scaling_x_param is None or result is None:
            return result
        out = self._to_data_units(result, scaling_x_param)
        if self._scaled():
            out.append(result[len(self.PARAMS)])
        return out

    def _default_fixed(self, xrange, n) -> list:
        raise NotImplementedError

    def _default_random(self, xrange, n) -> list:
        raise NotImplementedError

    def _prior(self, xrange, n) -> List[dict]:
        raise NotImplementedError

    def _bounds(self, xrange, n) -> list:
        raise NotImplementedError

    def get_default_hyper_parameter_fixed(self, xrange, n):
        hp = self._default_fixed(xrange, n)
        if self._scaled():
            hp.append(_f64(0.1))
        return hp

    def get_default_hyper_parameter_distribution(self, xrange, n):
        hp = self._default_random(xrange, n)
        if self._scaled():
            hp.append(torch.abs(0.1 + 0.2 * torch.randn((), dtype=torch.float64)))
        return hp

    def get_default_hyper_parameter(self, xrange, n, from_distribution: bool = False):
        if from_distribution:
            return self.get_default_hyper_parameter_distribution(xrange, n)
        return self.get_default_hyper_parameter_fixed(xrange, n)

    def get_hyper_parameter_distribution_definition(self, xrange, n) -> List[dict]:
        prior = self._prior(xrange, n)
        if self._scaled():
            prior.append({"shape": [], "mean": 0.1, "stddev": 0.2, "type": "random_normal"})
        return prior

    def get_hyper_parameter_bounds(self, xrange, n):
        bounds = self._bounds(xrange, n)
        if self._scaled():
            bounds.append((_f64(global_param.p_cov_matrix_jitter * 100), _f64(_INF)))
        return bounds


# ---- shared length-scale behaviour (SE, MAT32, MAT52; PER adds a period) ---------------------------------------------
class _LengthScaleKernel(BaseKernel):
    PARAMS = ("l",)
    ABS_ON_SET = (0,)

    @staticmethod
    def _width(xrange):
        return xrange[0][1] - xrange[0][0]

    def _default_fixed(self, xrange, n):
        return [_f64(self._width(xrange) / 10)]                       # BaseKernels.py:323-332

    def _default_random(self, xrange, n):
        return [torch.abs(self._width(xrange) / 10 + 0.2 * torch.randn((), dtype=torch.float64))]   # :334-350

    def _prior(self, xrange, n):
        return [{"shape": [], "mean": self._width(xrange) / 10, "stddev": 0.2, "type": "random_normal"}]

    def _bounds(self, xrange, n):
        w = self._width(xrange)
        return [(_f64(5 * w / n), _f64(w / 3))]                        # :296-306

    def _to_data_units(self, hp, s):
        return [hp[0] * s[1]]                                          # :417-427


class SquaredExponentialKernel(_LengthScaleKernel):
    SPEC, MANIFESTATION = "SE", k.KernelManifestation.SE


class MaternKernel3_2(_LengthScaleKernel):
    SPEC, MANIFESTATION = "MAT32", k.KernelManifestation.MAT32


class MaternKernel5_2(_LengthScaleKernel):
    SPEC, MANIFESTATION = "MAT52", k.KernelManifestation.MAT52


class PeriodicKernel(_LengthScaleKernel):
    SPEC, MANIFESTATION = "PER", k.KernelManifestation.PER
    PARAMS = ("l", "p")
    ABS_ON_SET = (0, 1)

    def _default_fixed(self, xrange, n):
        v = self._width(xrange) / 10
        return [_f64(v), _f64(v)]                                      # :490-501

    @staticmethod
    def _avg_dist(xrange, n):
        return min(r[1] - r[0] for r in xrange) / n

    def _default_random(self, xrange, n):
        l = torch.abs(self._width(xrange) / 10 + 0.2 * torch.randn((), dtype=torch.float64))
        a = self._avg_dist(xrange, n)
        lo, hi = a * 5, a * (n / 2)
        p = lo + (hi - lo) * torch.rand((), dtype=torch.float64)       # :503-527
        return [l, p]

    def _prior(self, xrange, n):
        a = self._avg_dist(xrange, n)
        return [{"shape": [], "mean": self._width(xrange) / 10, "stddev": 0.2, "type": "random_normal"},
                {"shape": [], "minval": a * 5, "maxval": a * (n / 2), "type": "random_uniform"}]

    def _bounds(self, xrange, n):
        w = self._width(xrange)
        # the period bounds are logarithms in the reference (sic, :466-467); kept for drop-in behaviour
        return [(_f64(5 * w / n), _f64(w / 3)), (_f64(math.log(10 * (w / n))), _f64(math.log(w / 5)))]

    def _to_data_units(self, hp, s):
        return [hp[0] * s[1], hp[1] * s[1]]                            # :617-627


class LinearKernel(BaseKernel):
    SPEC, MANIFESTATION = "LIN", k.KernelManifestation.LIN
    PARAMS = ("c",)

    def _param_shape(self, idx):
        return [self.input_dimensionality]

    def _default_fixed(self, xrange, n):
        return [torch.full((self.input_dimensionality,), 0.01, dtype=torch.float64)]   # :168-176

    def _span(self, xrange):
        lo = min(r[0] - (r[1] - r[0]) for r in xrange)
        hi = max(r[1] + (r[1] - r[0]) for r in xrange)
        return lo, hi

    def _default_random(self, xrange, n):
        lo, hi = self._span(xrange)
        return [lo + (hi - lo) * torch.rand((self.input_dimensionality,), dtype=torch.float64)]   # :178-196

    def _prior(self, xrange, n):
        x_min = torch.tensor([r[0] for r in xrange], dtype=torch.float64)
        x_max = torch.tensor([r[1] for r in xrange], dtype=torch.float64)
        return [{"shape": [self.input_dimensionality], "minval": x_min - (x_max - x_min),
                 "maxval": x_max + (x_max - x_min), "type": "random_uniform"}]

    def _bounds(self, xrange, n):
        d = self.input_dimensionality
        return [(torch.full((d,), -_INF, dtype=torch.float64), torch.full((d,), _INF, dtype=torch.float64))]

    def _to_data_units(self, hp, s):
        return [hp[0] * s[1] + s[0]]                                   # :259-269


class WhiteNoiseKernel(BaseKernel):
    SPEC, MANIFESTATION = "WN", k.KernelManifestation.WN
    PARAMS = ()

    def _default_fixed(self, xrange, n):
        return []

    def _default_random(self, xrange, n):
        return []

    def _prior(self, xrange, n):
        return []

    def _bounds(self, xrange, n):
        return []

    def set_last_hyper_parameter(self, last_hyper_parameter):
        pass

    def get_last_hyper_parameter(self, scaling_x_param=None):
        return []


class SquaredExponentialArdKernel(BaseKernel):
    """exp(-1/2 sum_d (x_d - x'_d)^2 / l_d^2): per-dimension length scales.  Extension for config C5 of BASELINE.json
    (the reference's SE is isotropic, BaseKernels.py:365-371); equal to SE(l=1) on x / l."""
    SPEC, MANIFESTATION = "SE_ARD", k.KernelManifestation.SE_ARD
    PARAMS = ("l",)
    ABS_ON_SET = (0,)

    def _param_shape(self, idx):
        return [self.input_dimensionality]

    def _default_fixed(self, xrange, n):
        return [torch.tensor([(r[1] - r[0]) / 10 for r in xrange], dtype=torch.float64)]

    def _default_random(self, xrange, n):
        base = torch.tensor([(r[1] - r[0]) / 10 for r in xrange], dtype=torch.float64)
        return [torch.abs(base + 0.2 * torch.randn(self.input_dimensionality, dtype=torch.float64))]

    def _prior(self, xrange, n):
        return [{"shape": [self.input_dimensionality], "mean": [(r[1] - r[0]) / 10 for r in xrange], "stddev": 0.2,
                 "type": "random_normal"}]

    def _bounds(self, xrange, n):
        lo = torch.tensor([5 * (r[1] - r[0]) / n for r in xrange], dtype=torch.float64)
        hi = torch.tensor([(r[1] - r[0]) / 3 for r in xrange], dtype=torch.float64)
        return [(lo, hi)]

    def _to_data_units(self, hp, s):
        return [hp[0] * s[1]]


class ConstantKernel(BaseKernel):
    """Dead in the reference as well: its constructor raises (BaseKernels.py:54-57)."""
    SPEC, MANIFESTATION = None, k.KernelManifestation.C

    def __init__(self, input_dimensionality: int):
        raise Exception("Not up to date. Implementation for ConstantKernel is not available.")
