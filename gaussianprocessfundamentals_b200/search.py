"""Batched evaluation of many candidate kernels on the same data (BASELINE config 3: "batched kernel search").

The reference evaluates the candidates of a structure search one at a time: one GaussianProcess + LogLikelihood metric
per candidate, each a chain of TensorFlow ops (`get_metric_by_type(MetricType.LL, gp)` then `metric.get_metric(hp, noise)`,
Metrics/Auxiliary.py:13-51, Metrics/LogLikelihood.py:30-65).  Here all candidates of a round form ONE batched device plan
(assembly, blocked Cholesky with carried y, inverse and trace gradient of every candidate in the same launches), and with
more than one process the candidates are sharded across the ranks by estimated cost with no data-path collective -
only the per-candidate scalars and flat gradients are all-gathered at the end (sharding.Sharding).

    batch = CandidateBatch(kernels, data_input)                # every rank passes the same list
    nll, grads, gnoise = batch.evaluate(hp_lists, noise)       # per candidate: NLL, [d NLL / d hp], d NLL / d noise

Every rank returns the values of ALL candidates, identical to a single-process evaluation.
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import global_parameters as global_param
from .program import compile_spec, unflatten_grad
from .sharding import Sharding, estimated_cost


class CandidateBatch:
    def __init__(self, kernels: Sequence, data_input, group=None, rank: Optional[int] = None,
                 world: Optional[int] = None, want_grad: bool = True):
        self.kernels = list(kernels)
        self.data_input = data_input
        self.want_grad = bool(want_grad)
        n = int(data_input.n_train)
        # equal n: the cost differs only through the kernel program (assembly / gradient), so weigh by its length
        scaled = bool(global_param.p_scaled_base_kernel)
        self._compiled = [compile_spec(k.to_spec(), k.get_dimensionality(), scaled) for k in self.kernels]
        costs = [estimated_cost(n, want_grad) + 40.0 * n * n * c.n_ops for c in self._compiled]
        self.sharding = Sharding(costs, group=group, rank=rank, world=world)
        self._blocks = None

    # the local share as one batched plan; a hook so that the host logic can be tested without a device
    def _make_blocks(self, kernels, xs, ys):
        from .Statistics._device import DeviceBlocks
        return DeviceBlocks(kernels, xs, ys, want_grad=self.want_grad)

    def _local_blocks(self):
        if self._blocks is None and self.sharding.mine:
            x = self.data_input.data_x_train
            y = self.data_input.get_detrended_y_train()
            mine = self.sharding.mine
            self._blocks = self._make_blocks([self.kernels[i] for i in mine], [x] * len(mine), [y] * len(mine))
        return self._blocks

    def evaluate(self, hp_lists: Sequence[list], noise, want_grad: Optional[bool] = None):
        """hp_lists[i]: the reference-style hyper-parameter list of candidate i.  Returns (nll [C], grads, gnoise):
        grads[i] is shaped like hp_lists[i] and gnoise[i] = d NLL_i / d noise (both None without gradients).
        Raises engine.NotPositiveDefinite on every rank if any candidate's matrix is not positive definite."""
        from . import engine
        want_grad = self.want_grad if want_grad is None else bool(want_grad)
        if want_grad and not self.want_grad:
            raise ValueError("this batch was created without gradient workspace")
        assert len(hp_lists) == len(self.kernels)
        s2 = float(torch.as_tensor(noise, dtype=torch.float64))
        mine = self.sharding.mine
        blocks = self._local_blocks()
        if mine:
            nll_l, grads_l = blocks.evaluate([hp_lists[i] for i in mine], [s2] * len(mine), want_grad, check=False)
            info_l = blocks.last[2]
        else:
            nll_l, grads_l, info_l = [], [], []
        sizes = [c.n_hp + 1 for c in self._compiled]
        nll, grads, info = self.sharding.combine(nll_l, grads_l if want_grad else None, info_l, sizes)
        bad = np.nonzero(info)[0]
        if bad.size:
            raise engine.NotPositiveDefinite(int(info[bad[0]]))
        for k, hp in zip(self.kernels, hp_lists):
            k._remember(hp)
        if not want_grad:
            return nll, None, None
        glists = [unflatten_grad(c.entries, g[:-1], like=hp) for c, g, hp in zip(self._compiled, grads, hp_lists)]
        return nll, glists, np.array([float(g[-1]) for g in grads])

    def best(self, hp_lists: Sequence[list], noise) -> int:
        """index of the candidate with the smallest NLL (the reference's metrics are 'minimum = optimum',
        Metrics/Metrics.py:27-29)"""
        nll, _, _ = self.evaluate(hp_lists, noise, want_grad=False)
        return int(np.argmin(nll))
