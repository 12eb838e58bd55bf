"""gaussianprocessfundamentals_b200 - B200-native exact-GP likelihood path behind the gpbasics 2.0.0 operator surface.

Layout: `csrc/` holds the hand-written sm_100a kernels and the C-ABI (include/gpb.h); `engine.py` / `program.py` drive
it; the sub-packages mirror the reference's module paths (KernelBasics, Statistics, Metrics, Optimizer, DataHandling,
MeanFunctionBasics, Auxiliary, global_parameters) so that reference call sites read unchanged.
"""
__version__ = "0.1.0"
