"""Pairwise distances (mirror of gpbasics/Auxiliary/Distances.py:4-12).

On the likelihood path the distances are never materialised: the assembly kernel computes them per pair in registers
(csrc/program.cuh gpb_sqdist / gpb_l1dist).  These two functions exist for callers that want the distance matrix
itself; they run the same device code through the one-opcode programs L2 / L1.  The reference forms
sqrt(|a|^2 - 2ab + |b|^2), which goes NaN under round-off; the device code sums (a_d - b_d)^2
directly, which is identical wherever the reference is finite."""
import torch

from .. import engine


def _distance(a, b, kind: str) -> torch.Tensor:
    engine.require_cuda()
    a = torch.as_tensor(a, dtype=torch.float64).cuda().contiguous()
    b = torch.as_tensor(b, dtype=torch.float64).cuda().contiguous()
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1], "expects a[n,d], b[m,d]"
    prog = engine.DeviceProgram.get((kind,), int(a.shape[1]), False, 1)
    empty = torch.empty(0, dtype=torch.float64, device=a.device)
    return engine.assemble(prog, a, b, empty, None)


def euclidian_distance(a, b) -> torch.Tensor:
    return _distance(a, b, "L2")


def manhattan_distance(a, b) -> torch.Tensor:
    return _distance(a, b, "L1")
