"""Fitters (mirror of gpbasics/Optimizer/Fitter.py:20-181).

`VariationalSgdFitter.fit()` keeps the reference's contract: evaluate the metric, take exactly ONE gradient step, store
the hyper-parameters on the kernel, evaluate again, return (pre_fit_metric, post_fit_metric, hyper_parameters, noise,
indices) (Fitter.py:61-170).  The reference delegates the step to tfp.optimizer.VariationalSGD, whose update inside
the burn-in phase is a plain gradient step at `burnin_max_learning_rate` (1e-6, TFP default); that is what is applied
here.  Parity is defined on (NLL, gradient) - the quantities the device path produces - not on the optimiser's
update.

`AdamFitter` is the real fit loop the reference leaves to its callers: many fused
LML+gradient evaluations with the data resident on the GPU and only the hyper-parameter vector crossing the bus."""
import logging
from typing import List, Tuple, Union

import torch

from .. import global_parameters as global_param
from ..Metrics import Auxiliary as met_aux
from ..Metrics import MatrixHandlingTypes as mht
from ..Metrics import Metrics as met
from . import FitterType as ft

global_param.ensure_init()


class Fitter:
    def __init__(self, data_input, gaussian_process, metric_type: met.MetricType, fitter_type: ft.FitterType,
                 from_distribution: bool, local_approx, numerical_matrix_handling, subset_size: int = None):
        self._gp = gaussian_process
        if metric_type in (met.MetricType.MSE, met.MetricType.blockwise_MSE):
            # the fitters drive the fused likelihood gradient of the device path; the adjoint of the posterior mean is
            # not part of it (the reference would differentiate MSE through its GradientTape)
            raise NotImplementedError("fitting on %s is not implemented on the B200 path: use MetricType.LL / BIC or "
                                      "their blockwise variants" % metric_type)
        if isinstance(data_input, list):
            self.metric = []
            for instance in data_input:
                copied = self._gp.copy()
                copied.set_data_input(instance)
                self.metric.append(met_aux.get_metric_by_type(metric_type, copied, local_approx,
                                                              numerical_matrix_handling, subset_size))
        else:
            self._gp.set_data_input(data_input)
            self.metric = met_aux.get_metric_by_type(metric_type, self._gp, local_approx, numerical_matrix_handling,
                                                     subset_size)
        self.data_input = data_input
        self.type = fitter_type
        self.from_distribution = from_distribution

    # ---- shared machinery ------------------------------------------------------------------------------------------
    def _first_input(self):
        return self.data_input[0] if isinstance(self.data_input, list) else self.data_input

    def _objective(self, variables: List[torch.Tensor]):
        """(value, [gradients aligned with `variables`]); with p_optimize_noise variables[0] is the raw noise and the
        metric sees |raw| (Fitter.py:94-95)"""
        metrics = self.metric if isinstance(self.metric, list) else [self.metric]
        if global_param.p_optimize_noise:
            raw, hp = variables[0], variables[1:]
            noise = torch.abs(raw.detach())
        else:
            raw, hp = None, variables
            noise = global_param.p_cov_matrix_jitter
        hp = [h.detach() for h in hp]
        total, grads, gnoise = 0.0, None, 0.0
        for m in metrics:
            m.covariance_matrix.reset() if hasattr(m, "covariance_matrix") else None
            v, g, gn = m._eval(hp, noise, True)
            total += v
            gnoise += gn
            g = [torch.as_tensor(t, dtype=torch.float64) for t in g]
            grads = g if grads is None else [a + b for a, b in zip(grads, g)]
        k = float(len(metrics))
        grads = [g / k for g in grads]
        if raw is not None:
            sign = torch.sign(raw.detach())
            grads = [torch.as_tensor(gnoise / k, dtype=torch.float64) * sign] + grads
        return total / k, grads

    def _metric_value(self, variables) -> torch.Tensor:
        metrics = self.metric if isinstance(self.metric, list) else [self.metric]
        if global_param.p_optimize_noise:
            noise, hp = torch.abs(variables[0].detach()), variables[1:]
        else:
            noise, hp = global_param.p_cov_matrix_jitter, variables
        vals = [m.get_metric([h.detach() for h in hp], noise, None) for m in metrics]
        return sum(vals) / len(vals)

    def _initial_variables(self):
        first = self._first_input()
        xrange, n = first.get_x_range(), first.n_train
        kernel = self._gp.covariance_matrix.kernel
        hp = [torch.as_tensor(h, dtype=torch.float64).clone() for h in
              kernel.get_default_hyper_parameter(xrange, n, self.from_distribution)]
        if global_param.p_optimize_noise:
            hp = [global_param.p_cov_matrix_jitter.clone()] + hp
        return hp, xrange, n

    def _bounded(self, variables, grads, xrange, n):
        """gradient replacement outside the kernel's bounds (Fitter.py:122-152)"""
        bounds = self._gp.kernel.get_hyper_parameter_bounds(xrange, n)
        out = list(grads)
        for idx, (lo, hi) in enumerate(bounds):
            v = variables[idx]
            if bool(torch.all(v < lo)):
                out[idx] = -torch.abs(lo / v)
            elif bool(torch.all(v > hi)):
                out[idx] = torch.abs(v / hi)
        return out

    def _store(self, variables):
        kernel = self._gp.covariance_matrix.kernel
        if global_param.p_optimize_noise:
            kernel.set_last_hyper_parameter(list(variables[1:]))
            kernel.set_noise(torch.abs(variables[0]))
        else:
            kernel.set_last_hyper_parameter(list(variables))
            kernel.set_noise(global_param.p_cov_matrix_jitter)

    def fit(self) -> Tuple[torch.Tensor, torch.Tensor, List[torch.Tensor], torch.Tensor, torch.Tensor]:
        raise NotImplementedError


class GradientFitter(Fitter):
    pass


class VariationalSgdFitter(Fitter):
    BURNIN_LEARNING_RATE = 1e-6

    def __init__(self, data_input, gaussian_process, metric_type: met.MetricType, from_distribution: bool, local_approx,
                 numerical_matrix_handling, subset_size: int = None):
        super().__init__(data_input, gaussian_process, metric_type, ft.FitterType.NON_GRADIENT, from_distribution,
                         local_approx, numerical_matrix_handling, subset_size)
        self.last_gradients = None

    def fit(self):
        variables, xrange, n = self._initial_variables()
        pre_fit_metric = self._metric_value(variables)
        _, grads = self._objective(variables)
        if global_param.p_check_hyper_parameters:
            grads = self._bounded(variables, grads, xrange, n)
        self.last_gradients = grads
        variables = [v - self.BURNIN_LEARNING_RATE * g.reshape(v.shape) for v, g in zip(variables, grads)]
        self._store(variables)
        post_fit_metric = self._metric_value(variables)
        kernel = self._gp.covariance_matrix.kernel
        return pre_fit_metric, post_fit_metric, kernel.get_last_hyper_parameter(), kernel.get_noise(), None


class AdamFitter(Fitter):
    """A real fit loop: `steps` fused LML+gradient evaluations driven by Adam on the hyper-parameter vector."""

    def __init__(self, data_input, gaussian_process, metric_type: met.MetricType, from_distribution: bool, local_approx,
                 numerical_matrix_handling, subset_size: int = None, steps: int = 50, learning_rate: float = 1e-2):
        super().__init__(data_input, gaussian_process, metric_type, ft.FitterType.GRADIENT, from_distribution,
                         local_approx, numerical_matrix_handling, subset_size)
        self.steps, self.learning_rate = int(steps), float(learning_rate)
        self.history: List[float] = []

    def fit(self):
        variables, xrange, n = self._initial_variables()
        pre_fit_metric = self._metric_value(variables)
        m = [torch.zeros_like(v) for v in variables]
        s = [torch.zeros_like(v) for v in variables]
        b1, b2, eps = 0.9, 0.999, 1e-8
        self.history = []
        for t in range(1, self.steps + 1):
            value, grads = self._objective(variables)
            self.history.append(value)
            if global_param.p_check_hyper_parameters:
                grads = self._bounded(variables, grads, xrange, n)
            for i, g in enumerate(grads):
                g = g.reshape(variables[i].shape)
                m[i] = b1 * m[i] + (1 - b1) * g
                s[i] = b2 * s[i] + (1 - b2) * g * g
                variables[i] = variables[i] - self.learning_rate * (m[i] / (1 - b1 ** t)) / \
                    (torch.sqrt(s[i] / (1 - b2 ** t)) + eps)
        self._store(variables)
        post_fit_metric = self._metric_value(variables)
        kernel = self._gp.covariance_matrix.kernel
        return pre_fit_metric, post_fit_metric, kernel.get_last_hyper_parameter(), kernel.get_noise(), None


class LbfgsFitter(Fitter):
    """Quasi-Newton fit loop: L-BFGS-B over the flat hyper-parameter vector, every function / gradient
    evaluation one fused device evaluation (assembly, Cholesky with carried y, inverse, trace gradient).  Box constraints
    are the kernel's own `get_hyper_parameter_bounds` when `p_check_hyper_parameters` is set (the reference only uses
    them to replace gradients, Fitter.py:122-152); a non positive-definite trial point is reported to the line search
    as +inf instead of aborting the fit."""

    def __init__(self, data_input, gaussian_process, metric_type: met.MetricType, from_distribution: bool, local_approx,
                 numerical_matrix_handling, subset_size: int = None, max_evaluations: int = 60, tolerance: float = 1e-9):
        super().__init__(data_input, gaussian_process, metric_type, ft.FitterType.GRADIENT, from_distribution,
                         local_approx, numerical_matrix_handling, subset_size)
        self.max_evaluations, self.tolerance = int(max_evaluations), float(tolerance)
        self.history: List[float] = []
        self.result = None

    @staticmethod
    def _flatten(variables):
        import numpy as np
        return np.concatenate([v.detach().cpu().numpy().reshape(-1) for v in variables]).astype("float64")

    @staticmethod
    def _unflatten(flat, like):
        out, pos = [], 0
        for v in like:
            k = v.numel()
            out.append(torch.as_tensor(flat[pos:pos + k], dtype=torch.float64).reshape(v.shape).clone())
            pos += k
        return out

    def _flat_bounds(self, variables, xrange, n):
        """(lo, hi) per scalar, or None; the raw noise (p_optimize_noise) is unbounded"""
        if not global_param.p_check_hyper_parameters:
            return None
        import numpy as np
        bounds = []
        if global_param.p_optimize_noise:
            bounds.append((None, None))
        for (lo, hi), v in zip(self._gp.kernel.get_hyper_parameter_bounds(xrange, n),
                               variables[1:] if global_param.p_optimize_noise else variables):
            lo = np.broadcast_to(np.asarray(lo, dtype="float64"), tuple(v.shape)).reshape(-1)
            hi = np.broadcast_to(np.asarray(hi, dtype="float64"), tuple(v.shape)).reshape(-1)
            for a, b in zip(lo, hi):
                a, b = (None if not np.isfinite(a) else float(a)), (None if not np.isfinite(b) else float(b))
                if a is not None and b is not None and a > b:
                    a, b = None, None          # the reference's log-scaled period bounds can be inverted
                bounds.append((a, b))
        return bounds

    def fit(self):
        import numpy as np
        from scipy.optimize import minimize
        variables, xrange, n = self._initial_variables()
        pre_fit_metric = self._metric_value(variables)
        self.history = []

        def fun(flat):
            trial = self._unflatten(flat, variables)
            try:
                value, grads = self._objective(trial)
            except ArithmeticError:            # engine.NotPositiveDefinite: let the line search back off
                return np.inf, np.zeros_like(flat)
            self.history.append(float(value))
            g = np.concatenate([np.asarray(t, dtype="float64").reshape(-1) for t in grads])
            if not np.isfinite(value) or not np.all(np.isfinite(g)):
                return np.inf, np.zeros_like(flat)
            return float(value), g

        self.result = minimize(fun, self._flatten(variables), jac=True, method="L-BFGS-B",
                               bounds=self._flat_bounds(variables, xrange, n),
                               options={"maxfun": self.max_evaluations, "ftol": self.tolerance, "gtol": 1e-10})
        best = self._unflatten(self.result.x, variables)
        self._store(best)
        post_fit_metric = self._metric_value(best)
        kernel = self._gp.covariance_matrix.kernel
        return pre_fit_metric, post_fit_metric, kernel.get_last_hyper_parameter(), kernel.get_noise(), None


if global_param.p_gradient_fitter is None:
    global_param.p_gradient_fitter = VariationalSgdFitter
