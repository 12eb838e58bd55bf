"""Process-wide configuration of the likelihood path (mirror of gpbasics/global_parameters.py:16-98).

Same flag names and `init(tf_parallel, worker)` signature as the reference, so callers that mutate
`global_param.p_*` keep working.  Differences, all deliberate:
  * a defined default state exists before `init()` (the reference raises NameError there);
  * `p_dtype` is torch.float64 and `p_cov_matrix_jitter` a 0-d torch tensor instead of TF objects;
  * `tf_parallel` only sizes the optional worker pool - the arithmetic runs on the GPU, not on TF thread pools.
The kernel-program compiler snapshots `p_scaled_base_kernel` / `p_cp_operator_type` when a kernel tree is compiled.
"""
import logging
import multiprocessing
import os
from enum import Enum

import torch


class ChangePointOperatorType(Enum):
    SIGMOID = 0
    INDICATOR = 1
    APPROX_INDICATOR = 2


def _mean_aggregator(t):
    return torch.mean(t)


initiated = False
p_dtype = torch.float64
p_cp_operator_type = ChangePointOperatorType.INDICATOR
p_cov_matrix_jitter = torch.tensor(1e-8, dtype=torch.float64)
p_optimize_noise = False
p_nystroem_ratio = 0.1
p_check_hyper_parameters = False
p_used_base_kernel = []
p_used_base_mean_functions = []
p_split_kernel = None
p_gradient_fitter = None
p_non_gradient_fitter = None
p_max_threads = 1
p_logging_level = logging.INFO
p_scaled_base_kernel = False
p_batch_metric_aggregator = _mean_aggregator
p_scale_data_y = True
pool = None


def ensure_init():
    """The reference aborts the process when `init` was skipped; here the defaults above are simply kept."""
    if not initiated:
        logging.debug("global parameters used with their defaults (init() was not called)")


def init(tf_parallel: int = 1, worker: bool = False):
    global initiated, p_dtype, p_cp_operator_type, p_cov_matrix_jitter, p_optimize_noise, p_nystroem_ratio, \
        p_check_hyper_parameters, p_used_base_kernel, p_used_base_mean_functions, p_split_kernel, p_gradient_fitter, \
        p_non_gradient_fitter, p_max_threads, p_logging_level, p_scaled_base_kernel, p_batch_metric_aggregator, pool, \
        p_scale_data_y
    initiated = True
    p_dtype = torch.float64
    p_cp_operator_type = ChangePointOperatorType.INDICATOR
    p_cov_matrix_jitter = torch.tensor(1e-8, dtype=torch.float64)
    p_optimize_noise = False
    p_nystroem_ratio = 0.1
    p_check_hyper_parameters = False
    p_used_base_kernel = []
    p_used_base_mean_functions = []
    p_split_kernel = None
    p_gradient_fitter = None
    p_non_gradient_fitter = None
    p_max_threads = max(1, (os.cpu_count() or 1) - int(tf_parallel))
    p_logging_level = logging.INFO
    p_scaled_base_kernel = False
    p_batch_metric_aggregator = _mean_aggregator
    pool = None
    p_scale_data_y = True
    logging.basicConfig(format="%(levelname)s: %(message)s", level=p_logging_level)
    logging.info("Process-%s:Initialization of global parameters finished." % os.getpid())


def set_up_pool(maxtasksperchild: int = -1):
    global pool
    if p_max_threads > 1:
        if maxtasksperchild is None or maxtasksperchild < 1:
            maxtasksperchild = None
        pool = multiprocessing.get_context("spawn").Pool(processes=p_max_threads, maxtasksperchild=maxtasksperchild)
    else:
        pool = None
        logging.warning("No Multiprocessing Pool set up, due to non-parallel execution")


def shutdown_pool():
    global pool
    if pool is not None:
        pool.close()
        pool.join()
        pool = None


def cp_mode_code() -> int:
    return int(p_cp_operator_type.value)
