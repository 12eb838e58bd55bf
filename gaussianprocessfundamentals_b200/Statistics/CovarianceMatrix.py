"""Covariance-matrix objects (mirror of gpbasics/Statistics/CovarianceMatrix.py:15-565): memoise-until-reset access to
K, K+s2 I, chol(K+s2 I), alpha, inverses - holistic or per block.

Every getter runs hand-written CUDA through libgpb (no TensorFlow, no CPU fallback):
  get_K / get_K_noised / get_K_s / get_K_ss   fused assembly kernel            (CovarianceMatrix.py:187-206, :213-221, :277-286)
  get_L_K                                     blocked FP64 tensor-core Cholesky (:247-254, per block :469-479)
  get_L_alpha                                 carried-RHS forward solve + blocked back substitution (:256-265, :492-506)
  get_K_inv / get_L_inv_K                     recursive-doubling triangular inverse, W^T W  (replaces the explicit
                                              tf.linalg.inv calls of :208-211, :267-275)
The likelihood itself does not go through these getters one by one: Metrics.LogLikelihood calls `nll_and_grad`, the
fused plan (assembly -> Cholesky with carried y -> NLL -> inverse -> trace gradient) in one host call.
Results are torch CUDA float64 tensors with the reference's shapes."""
from enum import Enum
from typing import List, Optional

import numpy as np
import torch

from .. import engine
from .. import global_parameters as global_param
from ..DataHandling import DataInput as di
from ..KernelBasics import Kernel as k
from ..KernelBasics import Operators as op
from ..program import flatten_hp
from ._device import DeviceBlocks

global_param.ensure_init()


class CovarianceMatrixType(Enum):
    HOLISTIC = 0
    SEGMENTED = 1
    GLOBALIZED_SEGMENTED = 2


def _noise_value(noise) -> float:
    if noise is None:
        raise Exception("No Data Input given or Noise unspecified")
    t = torch.as_tensor(noise, dtype=torch.float64)
    if t.dim() != 0:
        raise Exception("No Data Input given or Noise unspecified")
    return float(t)


class CovarianceMatrix:
    _FIELDS = ("K", "noised_K", "K_ss", "noised_K_ss", "L_K_ss", "L_K", "L_inv_K", "K_inv", "K_s", "L_alpha")

    def __init__(self, matrix_type: CovarianceMatrixType, kernel: k.Kernel):
        self.kernel = kernel
        self.data_input = None
        self.type = matrix_type
        self._blocks: Optional[DeviceBlocks] = None
        self.reset()

    def reset(self):
        """forget every memoised matrix (hyper-parameters changed); the device plan and the resident data stay"""
        for f in self._FIELDS:
            setattr(self, f, None)

    def set_data_input(self, data_input):
        self.data_input = data_input
        self._blocks = None
        self.reset()

    def is_segmented(self) -> bool:
        return self.type == CovarianceMatrixType.SEGMENTED

    def _need_data(self):
        if self.data_input is None:
            raise Exception("No Data Input given")


class HolisticCovarianceMatrix(CovarianceMatrix):
    def __init__(self, kernel: k.Kernel):
        super().__init__(CovarianceMatrixType.HOLISTIC, kernel)

    # ---- the fused path -------------------------------------------------------------------------------------------
    def _device_blocks(self) -> DeviceBlocks:
        self._need_data()
        # the reference re-reads get_detrended_y_train() on every get_metric: a new mean function or edited targets must
        # reach the device too, so the plan is keyed on the kernel program AND on the target tensor's identity / version
        y = self.data_input.get_detrended_y_train()
        if self._blocks is not None and not self._blocks.matches([self.kernel], y):
            self._blocks = None
        if self._blocks is None:
            self._blocks = DeviceBlocks([self.kernel], [self.data_input.data_x_train], [y], want_grad=True,
                                        grid=getattr(self, "_grid", None))
        return self._blocks

    def distribute(self, grid):
        """Factorise this ONE matrix over the processes of `grid` (engine.ProcessGrid; one process per GPU): distributed
        Cholesky with NCCL panel broadcasts, inverse and trace gradient split by block column (csrc/dist.cu).  Every
        rank must hold the same data and make the same calls in the same order; every rank gets the same NLL, gradient,
        L and alpha.  BASELINE config 5 / SURVEY 8(e)."""
        self._grid = grid
        self._blocks = None
        return self

    def nll_and_grad(self, hyper_parameter: List[torch.Tensor], noise, want_grad: bool = True):
        """(nll, [d nll / d hp], d nll / d noise) of the GP on the training data; grads are None without want_grad"""
        blocks = self._device_blocks()
        nll, grads = blocks.evaluate([hyper_parameter], [_noise_value(noise)], want_grad)
        self.kernel._remember(hyper_parameter)
        if not want_grad:
            return float(nll[0]), None, None
        glists, gnoise = blocks.grads_as_lists(grads, [hyper_parameter])
        return float(nll[0]), glists[0], gnoise[0]

    def _factorise(self, hyper_parameter, noise, inverse: bool):
        blocks = self._device_blocks()
        blocks.evaluate([hyper_parameter], [_noise_value(noise)], inverse)
        self.kernel._remember(hyper_parameter)
        return blocks

    # ---- reference getters ----------------------------------------------------------------------------------------
    def get_K(self, hyper_parameter):
        self._need_data()
        if self.K is None:
            x = self.data_input.data_x_train
            self.K = self.kernel.get_tf_tensor(hyper_parameter, x, x)
        return self.K

    def get_K_noised(self, hyper_parameter, noise):
        s2 = _noise_value(noise)
        self._need_data()
        if self.noised_K is None:
            K = self.get_K(hyper_parameter).clone()
            K.diagonal().add_(s2)
            self.noised_K = K
        return self.noised_K

    def get_K_ss(self, hyper_parameter):
        self._need_data()
        if self.K_ss is None:
            x = self.data_input.data_x_test
            self.K_ss = self.kernel.get_tf_tensor(hyper_parameter, x, x)
        return self.K_ss

    def get_K_ss_noised(self, hyper_parameter, noise):
        s2 = _noise_value(noise)
        self._need_data()
        if self.noised_K_ss is None:
            K = self.get_K_ss(hyper_parameter).clone()
            K.diagonal().add_(s2)
            self.noised_K_ss = K
        return self.noised_K_ss

    def get_K_s(self, hyper_parameter):
        self._need_data()
        if self.K_s is None:
            self.K_s = self.kernel.get_tf_tensor(hyper_parameter, self.data_input.data_x_train,
                                                 self.data_input.data_x_test)
        return self.K_s

    def get_L_K(self, hyper_parameter, noise):
        self._need_data()
        if self.L_K is None:
            blocks = self._factorise(hyper_parameter, noise, inverse=False)
            self.L_K = blocks.lower(0, engine.BUF_A)
            blocks.plan.eval(engine.STAGE_BACKSOLVE)
            self.L_alpha = blocks.vector(0, engine.BUF_ALPHA)
        return self.L_K

    def get_L_alpha(self, hyper_parameter, noise):
        self._need_data()
        if self.L_alpha is None:
            self.get_L_K(hyper_parameter, noise)
        return self.L_alpha

    def get_L_K_ss(self, hyper_parameter, noise):
        self._need_data()
        if self.L_K_ss is None:
            x = self.data_input.data_x_test
            tmp = DeviceBlocks([self.kernel], [x], [torch.zeros(x.shape[0], 1, dtype=torch.float64)], want_grad=False)
            tmp.evaluate([hyper_parameter], [_noise_value(noise)], False)
            self.L_K_ss = tmp.lower(0, engine.BUF_A)
        return self.L_K_ss

    def _inverse(self, hyper_parameter, noise):
        blocks = self._factorise(hyper_parameter, noise, inverse=True)
        self.L_inv_K = blocks.lower(0, engine.BUF_A)
        self.K_inv = blocks.symmetric(0, engine.BUF_KINV)
        if self.L_alpha is None:
            self.L_alpha = blocks.vector(0, engine.BUF_ALPHA)

    def get_K_inv(self, hyper_parameter, noise):
        self._need_data()
        if self.K_inv is None:
            self._inverse(hyper_parameter, noise)
        return self.K_inv

    def get_L_inv_K(self, hyper_parameter, noise):
        self._need_data()
        if self.L_inv_K is None:
            self._inverse(hyper_parameter, noise)
        return self.L_inv_K


class SegmentedCovarianceMatrix(CovarianceMatrix):
    """Block-diagonal covariance of a PartitionOperator / ChangePointOperator kernel over a PartitionedDataInput: the
    non-empty blocks form one batched plan (a single launch sequence for all blocks) instead of the reference's Python
    loop of small TF ops (CovarianceMatrix.py:316-339, :469-506).  Child hyper-parameters are consecutive slices that
    start after the change points of a CP kernel and advance over empty blocks too (:319-320, :334-337)."""

    def __init__(self, kernel):
        super().__init__(CovarianceMatrixType.SEGMENTED, kernel)

    def set_data_input(self, data_input: di.PartitionedDataInput):
        assert len(data_input.data_inputs) == len(self.kernel.child_nodes), \
            "Invalid data input. Data input does not match segments prescribed by given kernel"
        super().set_data_input(data_input)

    # ---- bookkeeping ----------------------------------------------------------------------------------------------
    def _slices(self) -> List[slice]:
        if isinstance(self.kernel, op.ChangePointOperator):
            return self.kernel.child_slices()
        out, idx = [], 0
        for cn in self.kernel.child_nodes:
            c = cn.get_number_of_hyper_parameter()
            out.append(slice(idx, idx + c))
            idx += c
        return out

    def _active(self) -> List[int]:
        return [i for i, blk in enumerate(self.data_input.data_inputs) if blk.n_train > 0]

    def shard(self, group=None, rank=None, world=None):
        """Distribute the non-empty blocks over the ranks of `group` (default: the world group) by estimated cost
        (sharding.assign).  Afterwards every rank evaluates only its own blocks in `block_nll_and_grad` and the
        per-block scalars and gradients are all-gathered, so the metric is identical on every rank and identical to
        the single-process value (independent units, no data-path collective).  The per-block matrix
        getters (get_L_K_blocks, ...) remain single-process."""
        from ..sharding import Sharding, estimated_cost
        self._need_data()
        act = self._active()
        costs = [estimated_cost(self.data_input.data_inputs[i].n_train) for i in act]
        self._sharding = Sharding(costs, group=group, rank=rank, world=world)
        self._blocks = None
        return self._sharding

    def _local_positions(self, act) -> List[int]:
        """positions (into the active list) of the blocks this rank evaluates"""
        sh = getattr(self, "_sharding", None)
        return list(range(len(act))) if sh is None else list(sh.mine)

    def _make_blocks(self, kernels, xs, ys):
        return DeviceBlocks(kernels, xs, ys, want_grad=True)

    def _device_blocks(self) -> DeviceBlocks:
        self._need_data()
        act = self._active()
        act = [act[p] for p in self._local_positions(act)]
        kernels = [self.kernel.child_nodes[i] for i in act]
        ys = [self.data_input.data_inputs[i].get_detrended_y_train() for i in act]
        if self._blocks is not None and not self._blocks.matches(kernels, ys):
            self._blocks = None       # a child kernel, a global switch or the detrended targets changed
        if self._blocks is None:
            xs = [self.data_input.data_inputs[i].data_x_train for i in act]
            self._blocks = self._make_blocks(kernels, xs, ys) if act else None
        return self._blocks

    def block_nll_and_grad(self, hyper_parameter, noise, want_grad: bool = True):
        """per-block NLL (list aligned with kernel.child_nodes, None for empty blocks) and, optionally, the gradient
        w.r.t. the full hyper-parameter list (zeros for change points) and the noise"""
        blocks = self._device_blocks()
        act, sl = self._active(), self._slices()
        sh = getattr(self, "_sharding", None)
        hp_lists = [list(hyper_parameter[sl[i]]) for i in act]
        glists = gnoise = None
        if sh is None:
            nll, grads = blocks.evaluate(hp_lists, [_noise_value(noise)] * len(act), want_grad)
            if want_grad:
                glists, gnoise = blocks.grads_as_lists(grads, hp_lists)
        else:
            # this rank's share, then one all-gather of (nll, info, flat gradient) per block
            mine = sh.mine
            if mine:
                nll_l, grads_l = blocks.evaluate([hp_lists[p] for p in mine], [_noise_value(noise)] * len(mine),
                                                 want_grad, check=False)
                info_l = blocks.last[2]
            else:
                nll_l, grads_l, info_l = [], [], []
            sizes = [sum(int(torch.as_tensor(h).numel()) for h in hp) + 1 for hp in hp_lists]
            nll, grads, info = sh.combine(nll_l, grads_l if want_grad else None, info_l, sizes)
            bad = np.nonzero(info)[0]
            if bad.size:
                raise engine.NotPositiveDefinite(int(info[bad[0]]))
            if want_grad:
                from ..program import unflatten_grad, compile_spec
                scaled = bool(global_param.p_scaled_base_kernel)
                glists, gnoise = [], []
                for pos, i in enumerate(act):
                    cn = self.kernel.child_nodes[i]
                    entries = compile_spec(cn.to_spec(), cn.get_dimensionality(), scaled).entries
                    glists.append(unflatten_grad(entries, grads[pos][:-1], like=hp_lists[pos]))
                    gnoise.append(float(grads[pos][-1]))
        per_block = [None] * len(self.kernel.child_nodes)
        for pos, i in enumerate(act):
            per_block[i] = float(nll[pos])
            self.kernel.child_nodes[i]._remember(hp_lists[pos])
        if not want_grad:
            return per_block, None, None
        full = [torch.zeros_like(torch.as_tensor(h, dtype=torch.float64)) for h in hyper_parameter]
        for pos, i in enumerate(act):
            for off, g in enumerate(glists[pos]):
                full[sl[i].start + off] = torch.as_tensor(g)
        return per_block, full, float(sum(gnoise))

    # ---- per-block getters (lists aligned with kernel.child_nodes; None for empty blocks) ---------------------------
    def _per_block(self, maker, which_n="n_train"):
        out = []
        for i, (cn, sl) in enumerate(zip(self.kernel.child_nodes, self._slices())):
            blk = self.data_input.data_inputs[i]
            out.append(maker(i, cn, sl, blk) if getattr(blk, which_n) > 0 else None)
        return out

    def get_K_blocks(self, hyper_parameter):
        return self._per_block(lambda i, cn, sl, blk: cn.get_tf_tensor(list(hyper_parameter[sl]), blk.data_x_train,
                                                                       blk.data_x_train))

    def get_K_noised_blocks(self, hyper_parameter, noise):
        s2 = _noise_value(noise)
        out = []
        for Kb in self.get_K_blocks(hyper_parameter):
            if Kb is None:
                out.append(None)
            else:
                Kb = Kb.clone()
                Kb.diagonal().add_(s2)
                out.append(Kb)
        return out

    def get_K_ss_blocks(self, hyper_parameter):
        return self._per_block(lambda i, cn, sl, blk: cn.get_tf_tensor(list(hyper_parameter[sl]), blk.data_x_test,
                                                                       blk.data_x_test), "n_test")

    def get_K_ss_noised_blocks(self, hyper_parameter, noise):
        s2 = _noise_value(noise)
        out = []
        for Kb in self.get_K_ss_blocks(hyper_parameter):
            if Kb is None:
                out.append(None)
            else:
                Kb = Kb.clone()
                Kb.diagonal().add_(s2)
                out.append(Kb)
        return out

    def _factor_blocks(self, hyper_parameter, noise, inverse: bool):
        if getattr(self, "_sharding", None) is not None:
            raise RuntimeError("the per-block matrix getters are single-process: this covariance matrix was shard()ed "
                               "and holds only this rank's blocks (use block_nll_and_grad, or an unsharded copy)")
        blocks = self._device_blocks()
        act, sl = self._active(), self._slices()
        blocks.evaluate([list(hyper_parameter[sl[i]]) for i in act], [_noise_value(noise)] * len(act), inverse)
        return blocks, act

    def _scatter(self, act, values):
        out = [None] * len(self.kernel.child_nodes)
        for pos, i in enumerate(act):
            out[i] = values[pos]
        return out

    def get_L_K_blocks(self, hyper_parameter, noise):
        blocks, act = self._factor_blocks(hyper_parameter, noise, False)
        Ls = [blocks.lower(pos, engine.BUF_A) for pos in range(len(act))]
        blocks.plan.eval(engine.STAGE_BACKSOLVE)
        self._alpha_blocks = self._scatter(act, [blocks.vector(pos, engine.BUF_ALPHA) for pos in range(len(act))])
        return self._scatter(act, Ls)

    def get_L_alpha_blocks(self, hyper_parameter, noise):
        self.get_L_K_blocks(hyper_parameter, noise)
        return self._alpha_blocks

    def get_L_inv_K_blocks(self, hyper_parameter, noise):
        blocks, act = self._factor_blocks(hyper_parameter, noise, True)
        return self._scatter(act, [blocks.lower(pos, engine.BUF_A) for pos in range(len(act))])

    def get_K_inv_blocks(self, hyper_parameter, noise):
        blocks, act = self._factor_blocks(hyper_parameter, noise, True)
        return self._scatter(act, [blocks.symmetric(pos, engine.BUF_KINV) for pos in range(len(act))])

    def get_L_K_ss_blocks(self, hyper_parameter, noise):
        out = []
        for i, (cn, sl) in enumerate(zip(self.kernel.child_nodes, self._slices())):
            blk = self.data_input.data_inputs[i]
            if blk.n_test > 0:
                tmp = DeviceBlocks([cn], [blk.data_x_test], [torch.zeros(blk.n_test, 1, dtype=torch.float64)], False)
                tmp.evaluate([list(hyper_parameter[sl])], [_noise_value(noise)], False)
                out.append(tmp.lower(0, engine.BUF_A))
            else:
                out.append(None)
        return out

    # ---- dense views (the reference's LinearOperatorBlockDiag(...).to_dense(), CovarianceMatrix.py:310-312) ---------
    @staticmethod
    def _dense(blocks):
        present = [b for b in blocks if b is not None]
        return torch.block_diag(*present) if present else torch.zeros(0, 0, dtype=torch.float64, device="cuda")

    def get_K(self, hyper_parameter):
        self._need_data()
        if self.K is None:
            self.K = self._dense(self.get_K_blocks(hyper_parameter))
        return self.K

    def get_K_noised(self, hyper_parameter, noise):
        self._need_data()
        if self.noised_K is None:
            self.noised_K = self._dense(self.get_K_noised_blocks(hyper_parameter, noise))
        return self.noised_K

    def get_K_ss(self, hyper_parameter):
        self._need_data()
        if self.K_ss is None:
            self.K_ss = self._dense(self.get_K_ss_blocks(hyper_parameter))
        return self.K_ss

    def get_K_ss_noised(self, hyper_parameter, noise):
        self._need_data()
        if self.noised_K_ss is None:
            self.noised_K_ss = self._dense(self.get_K_ss_noised_blocks(hyper_parameter, noise))
        return self.noised_K_ss

    def get_L_K_ss(self, hyper_parameter, noise):
        self._need_data()
        if self.L_K_ss is None:
            self.L_K_ss = self._dense(self.get_L_K_ss_blocks(hyper_parameter, noise))
        return self.L_K_ss

    def get_L_K(self, hyper_parameter, noise):
        self._need_data()
        if self.L_K is None:
            self.L_K = self._dense(self.get_L_K_blocks(hyper_parameter, noise))
        return self.L_K

    def get_L_alpha(self, hyper_parameter, noise):
        self._need_data()
        if self.L_alpha is None:
            parts = [a for a in self.get_L_alpha_blocks(hyper_parameter, noise) if a is not None]
            self.L_alpha = torch.cat(parts, dim=0)
        return self.L_alpha

    def get_L_inv_K(self, hyper_parameter, noise):
        self._need_data()
        if self.L_inv_K is None:
            self.L_inv_K = self._dense(self.get_L_inv_K_blocks(hyper_parameter, noise))
        return self.L_inv_K

    def get_K_inv(self, hyper_parameter, noise):
        self._need_data()
        if self.K_inv is None:
            self.K_inv = self._dense(self.get_K_inv_blocks(hyper_parameter, noise))
        return self.K_inv

    def get_K_s(self, hyper_parameter):
        self._need_data()
        if self.K_s is None:
            self.K_s = self.kernel.get_tf_tensor(hyper_parameter, self.data_input.data_x_train,
                                                 self.data_input.data_x_test)
        return self.K_s
