"""Device-side state shared by the covariance-matrix classes: one engine.Plan per (kernel programs, block sizes),
the data resident in its workspace, and bookkeeping of what the factorisation buffer currently holds."""
from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import engine
from .. import global_parameters as global_param
from ..program import flatten_hp, unflatten_grad


class DeviceBlocks:
    """B independent (kernel, X, y) blocks evaluated as one batched plan."""

    def __init__(self, kernels: Sequence, xs: Sequence[torch.Tensor], ys: Sequence[torch.Tensor], want_grad: bool = True,
                 grid=None, y_source=None):
        """y_source: the tensor(s) the detrended targets were taken from (default: `ys`); `matches` compares their
        identity and in-place version so that a new mean function or edited targets are uploaded again."""
        engine.require_cuda()
        self.kernels = list(kernels)
        self.ns = [int(x.shape[0]) for x in xs]
        scaled = bool(global_param.p_scaled_base_kernel)
        cp_mode = global_param.cp_mode_code()
        dims = {kern.get_dimensionality() for kern in self.kernels}
        if len(dims) == 1:      # programs that do not exist yet are compiled in parallel
            self.programs = engine.DeviceProgram.get_many([kern.to_spec() for kern in self.kernels], dims.pop(), scaled,
                                                          cp_mode)
        else:
            self.programs = [engine.DeviceProgram.get(kern.to_spec(), kern.get_dimensionality(), scaled, cp_mode)
                             for kern in self.kernels]
        self.key = (tuple(p.compiled.signature() for p in self.programs), tuple(self.ns), scaled, cp_mode, want_grad)
        self.plan = engine.Plan(self.programs, self.ns, want_grad=want_grad, grid=grid)
        self.want_grad = want_grad
        for b, (x, y) in enumerate(zip(xs, ys)):
            self.plan.set_data(b, x, y)
        self.holds = None          # "L" after the factorisation, "W" after the inverse
        self.last = None           # (nll[B], grads[B], info[B]) of the latest evaluation
        self._y_source = ys if y_source is None else y_source     # strong reference: ids are compared, never reused
        self._y_stamp = self._stamp(self._y_source)

    @staticmethod
    def _stamp(ys):
        ys = ys if isinstance(ys, (list, tuple)) else [ys]
        return tuple((id(t), getattr(t, "_version", 0)) for t in ys)

    def matches(self, kernels: Sequence, y_source) -> bool:
        """same kernel programs, same global switches and the same (unmodified) target tensors as at construction"""
        scaled = bool(global_param.p_scaled_base_kernel)
        cp_mode = global_param.cp_mode_code()
        if len(kernels) != len(self.programs) or (scaled, cp_mode) != self.key[2:4]:
            return False
        from ..program import compile_spec
        sigs = tuple(compile_spec(k.to_spec(), k.get_dimensionality(), scaled).signature() for k in kernels)
        return sigs == self.key[0] and self._stamp(y_source) == self._y_stamp

    def flat_hp(self, hp_lists: Sequence[list]) -> List[np.ndarray]:
        return [flatten_hp(p.compiled.entries, hp, p.n_hp) for p, hp in zip(self.programs, hp_lists)]

    def evaluate(self, hp_lists: Sequence[list], noises: Sequence[float], grad: bool, check: bool = True):
        """assembly -> Cholesky -> NLL (-> inverse -> gradient); host buffers in and out (the end-to-end call).
        check=False leaves the per-GP `info` in self.last for the caller (sharded evaluations raise on all ranks)."""
        stages = engine.STAGES_LML_GRAD if grad else engine.STAGES_LML
        nll, grads, info = self.plan.eval_host(self.flat_hp(hp_lists), [float(v) for v in noises], stages=stages)
        self.holds = "W" if grad else "L"
        self.last = (nll, grads, info)
        bad = np.nonzero(info)[0]
        if check and bad.size:
            raise engine.NotPositiveDefinite(int(info[bad[0]]))
        return nll, grads

    def grads_as_lists(self, grads, hp_lists):
        return [unflatten_grad(p.compiled.entries, g[:-1], like=hp) for p, g, hp in zip(self.programs, grads, hp_lists)], \
            [float(g[-1]) for g in grads]

    # ---- matrices (clones: the workspace is overwritten by the next evaluation) -------------------------------
    def lower(self, b: int, which: int) -> torch.Tensor:
        n = self.ns[b]
        buf = self.plan.buffer(b, which)
        ld = buf.shape[1]
        from .. import _lib
        _lib.check(self.plan.lib.gpb_zero_upper(buf.data_ptr(), n, ld, engine._stream_ptr()), "gpb_zero_upper")
        return buf[:n, :n].t().clone()

    def symmetric(self, b: int, which: int) -> torch.Tensor:
        n = self.ns[b]
        buf = self.plan.buffer(b, which)
        ld = buf.shape[1]
        from .. import _lib
        _lib.check(self.plan.lib.gpb_symmetrize(buf.data_ptr(), n, ld, engine._stream_ptr()), "gpb_symmetrize")
        return buf[:n, :n].t().clone()

    def vector(self, b: int, which: int) -> torch.Tensor:
        return self.plan.buffer(b, which).clone().reshape(-1, 1)
