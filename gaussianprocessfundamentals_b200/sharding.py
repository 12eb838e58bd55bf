"""Sharding of independent GPs across the ranks of one 8 x B200 box (first row).

The blocks of a Blockwise / PartitionedGaussianProcess (BlockwiseLogLikelihood sums independent per-block terms,
Metrics/LogLikelihood.py:85-104) and the candidates of a kernel search are independent evaluations: they are
distributed over the ranks by estimated cost, every rank evaluates its own share with the batched plan, and there is
no collective on the data path.  The only exchange is one all-gather of the per-GP scalars (NLL, info) and flat
gradients at the end of an evaluation, so that every rank returns the same totals a single-process call would.

Host logic only: it runs on `gloo` (CPU tensors) exactly as on `nccl` (CUDA tensors), which is how the tests cover the
world_size > 1 path without a GPU.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def estimated_cost(n: int, want_grad: bool = True) -> float:
    """leading-order FLOPs of one evaluation: n^3 / 3 (Cholesky) + 2 n^3 / 3 (inverse for the gradient)"""
    n = float(n)
    return n * n * n * (1.0 if want_grad else 1.0 / 3.0) + 64.0 * n * n


def assign(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of units to ranks.  Deterministic (ties: lower unit index first,
    lower rank first) so that every rank computes the same map without communication; each rank's list is ascending.
    Equal costs degenerate to round-robin."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (load[q], len(out[q]), q))
        out[r].append(i)
        load[r] += float(costs[i])
    return [sorted(v) for v in out]


class Sharding:
    """The unit -> rank map of one workload plus the final all-gather."""

    def __init__(self, costs: Sequence[float], group: Optional["dist.ProcessGroup"] = None,
                 rank: Optional[int] = None, world: Optional[int] = None):
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
        self.rank, self.world = int(rank), int(world)
        self.n_units = len(costs)
        self.map = assign(costs, self.world)
        self.mine = self.map[self.rank]
        self.owner = np.empty(self.n_units, dtype=np.int64)
        for r, units in enumerate(self.map):
            self.owner[units] = r

    def _device(self) -> torch.device:
        if self.world > 1 and dist.get_backend(self.group) == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def gather(self, local_rows: Sequence[np.ndarray], width: int) -> List[np.ndarray]:
        """local_rows[p] = float64 row (<= width entries) of this rank's p-th unit; returns the rows of all units in
        unit order on every rank.  One all-gather of a [max units per rank, width] tensor."""
        assert len(local_rows) == len(self.mine)
        per = max(len(u) for u in self.map) if self.n_units else 0
        buf = np.zeros((per, width), dtype=np.float64)
        for p, row in enumerate(local_rows):
            row = np.asarray(row, dtype=np.float64).reshape(-1)
            buf[p, :row.size] = row
        if self.world == 1:
            parts = [buf]
        else:
            dev = self._device()
            mine = torch.from_numpy(buf).to(dev)
            outs = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(outs, mine, group=self.group)
            parts = [o.cpu().numpy() for o in outs]
        rows: List[Optional[np.ndarray]] = [None] * self.n_units
        for r, units in enumerate(self.map):
            for p, u in enumerate(units):
                rows[u] = parts[r][p]
        return rows  # type: ignore[return-value]

    def combine(self, nll: Sequence[float], grads: Optional[Sequence[np.ndarray]], info: Sequence[int],
                grad_sizes: Sequence[int]) -> Tuple[np.ndarray, Optional[List[np.ndarray]], np.ndarray]:
        """per-unit (nll, flat gradient, info) of this rank's units -> the same for all units, on every rank.
        grad_sizes[u] = length of unit u's flat gradient (known everywhere: it follows from the kernel tree)."""
        width = 2 + (max(grad_sizes) if (grads is not None and len(grad_sizes)) else 0)
        local = []
        for p in range(len(self.mine)):
            row = [float(nll[p]), float(info[p])]
            if grads is not None:
                row += list(np.asarray(grads[p], dtype=np.float64).reshape(-1))
            local.append(np.asarray(row))
        rows = self.gather(local, width)
        all_nll = np.array([r[0] for r in rows], dtype=np.float64)
        all_info = np.array([int(r[1]) for r in rows], dtype=np.int64)
        all_grads = None
        if grads is not None:
            all_grads = [rows[u][2:2 + grad_sizes[u]].copy() for u in range(self.n_units)]
        return all_nll, all_grads, all_info
