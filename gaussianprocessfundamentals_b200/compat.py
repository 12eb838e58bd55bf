"""Drop-in switch: `install_as_gpbasics()` registers this package under the reference's import name, so that
`import gpbasics.KernelBasics.BaseKernels as bk` etc. resolve to the B200 implementation without touching call sites."""
import importlib
import sys

_SUBMODULES = [
    "global_parameters",
    "Auxiliary", "Auxiliary.BasicGPComponent", "Auxiliary.Distances",
    "KernelBasics", "KernelBasics.Kernel", "KernelBasics.BaseKernels", "KernelBasics.Operators",
    "KernelBasics.PartitioningModel", "KernelBasics.PartitionOperator",
    "MeanFunctionBasics", "MeanFunctionBasics.MeanFunction", "MeanFunctionBasics.BaseMeanFunctions",
    "DataHandling", "DataHandling.AbstractDataInput", "DataHandling.DataInput", "DataHandling.BatchDataInput",
    "Statistics", "Statistics.CovarianceMatrix", "Statistics.GaussianProcess", "Statistics.Auxiliary",
    "Metrics", "Metrics.MatrixHandlingTypes", "Metrics.Metrics", "Metrics.LogLikelihood",
    "Metrics.BayesianInformationCriterion", "Metrics.Auxiliary",
    "Optimizer", "Optimizer.FitterType", "Optimizer.Fitter",
]


def install_as_gpbasics(name: str = "gpbasics"):
    pkg = importlib.import_module(__package__)
    sys.modules[name] = pkg
    for sub in _SUBMODULES:
        sys.modules[name + "." + sub] = importlib.import_module(__package__ + "." + sub)
    return pkg
