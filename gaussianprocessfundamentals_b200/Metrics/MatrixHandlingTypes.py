"""Enums selecting approximations / numerical handling (mirror of gpbasics/Metrics/MatrixHandlingTypes.py).  The
B200 path implements MatrixApproximations.NONE with NumericalMatrixHandlingType.CHOLESKY_BASED - the exact GP."""
from enum import Enum


class GlobalApproximationsType(Enum):
    pass


class MatrixApproximations(GlobalApproximationsType):
    NONE = 0
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
    CHOLESKY_BASED = 2
    LINEAR_CONJUGATE_GRADIENT = 3
