"""(mirror of gpbasics/Optimizer/FitterType.py)"""
from enum import Enum


class FitterType(Enum):
    GRADIENT = 0
    NON_GRADIENT = 1
