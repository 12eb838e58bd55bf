"""Enums selecting approximations / numerical handling.  Names and values are the reference's
(gpbasics/Metrics/MatrixHandlingTypes.py) - a constant table that is part of the drop-in surface, because callers pass
these members positionally into get_metric_by_type and the fitters.  The B200 path implements the exact GP only:
MatrixApproximations.NONE with NumericalMatrixHandlingType.CHOLESKY_BASED (everything else raises NotImplementedError
where it is consumed)."""
from enum import Enum


class GlobalApproximationsType(Enum):
    """common base so that `isinstance(x, GlobalApproximationsType)` covers both families below"""


class MatrixApproximations(GlobalApproximationsType):
    NONE = 0                # exact GP: the only member implemented on the device
    SKC_LOWER_BOUND = 1
    SKC_UPPER_BOUND = 2
    BASIC_NYSTROEM = 3
    SKI = 4


class SubsetOfDataApproaches(GlobalApproximationsType):
    SOD_RANDOM = 5
    SOD_GRID = 6
    SOD_SMOOTHED_GRID = 7


class NumericalMatrixHandlingType(Enum):
    STRICT_INVERSE = 0
    PSEUDO_INVERSE = 1
    CHOLESKY_BASED = 2      # the only member implemented on the device
    LINEAR_CONJUGATE_GRADIENT = 3


IMPLEMENTED = (MatrixApproximations.NONE, NumericalMatrixHandlingType.CHOLESKY_BASED)
