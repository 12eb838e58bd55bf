"""Enums selecting approximations / numerical handling.  Names and values are the reference's
(gpbasics/Metrics/MatrixHandlingTypes.py), because callers pass them positionally into get_metric_by_type and the
fitters; the B200 path implements the exact GP only: MatrixApproximations.NONE with
NumericalMatrixHandlingType.CHOLESKY_BASED (everything else raises NotImplementedError where it is consumed)."""
from enum import Enum


class GlobalApproximationsType(Enum):
    """common base so that `isinstance(x, GlobalApproximationsType)` covers both families below"""


def _family(name: str, first_value: int, *members: str):
    return GlobalApproximationsType(name, {m: first_value + i for i, m in enumerate(members)}, module=__name__)


# covariance-matrix approximations (values 0..4) and subset-of-data approaches (values 5..7)
MatrixApproximations = _family("MatrixApproximations", 0, "NONE", "SKC_LOWER_BOUND", "SKC_UPPER_BOUND", "BASIC_NYSTROEM",
                               "SKI")
SubsetOfDataApproaches = _family("SubsetOfDataApproaches", 5, "SOD_RANDOM", "SOD_GRID", "SOD_SMOOTHED_GRID")

# how alpha and the log-determinant are obtained; only CHOLESKY_BASED is implemented on the device
NumericalMatrixHandlingType = Enum("NumericalMatrixHandlingType",
                                   {m: i for i, m in enumerate(("STRICT_INVERSE", "PSEUDO_INVERSE", "CHOLESKY_BASED",
                                                                "LINEAR_CONJUGATE_GRADIENT"))}, module=__name__)

IMPLEMENTED = (MatrixApproximations.NONE, NumericalMatrixHandlingType.CHOLESKY_BASED)
