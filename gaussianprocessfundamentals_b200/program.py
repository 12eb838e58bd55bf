"""Kernel tree -> flat postfix program for the CUDA interpreter (csrc/program.cuh).

A *spec* is a nested tuple mirroring the reference's kernel objects:
    ("SE",) ("PER",) ("LIN",) ("MAT32",) ("MAT52",) ("WN",) ("SE_ARD",)
    ("ADD", [specs]) ("MUL", [specs]) ("CP", [specs])
The flat hyper-parameter vector is the concatenation of the reference's hyper-parameter *list* in its own order:
children are visited depth first in their current order (Operators.py:207-225, :306-326), a change-point node
contributes its change points first (Operators.py:451-453, :507-511), a leaf contributes [params..., sg] with sg only
under p_scaled_base_kernel (BaseKernels.py:153-159, :308-314, :475-481).  N-ary ADD / MUL are emitted as left folds,
exactly the order in which the reference accumulates them.
"""
from typing import List, Tuple

import numpy as np

OP = {"SE": 1, "PER": 2, "LIN": 3, "MAT32": 4, "MAT52": 5, "WN": 6, "SE_ARD": 7, "L2": 8, "L1": 9, "ADD2": 16, "MUL2": 17, "CPW": 18}
MAX_OPS, MAX_STACK, MAX_TAPE, MAX_DIM, MAX_HP = 256, 8, 64, 16, 128


def leaf_entries(kind: str, dim: int, scaled: bool) -> List[int]:
    """sizes of the hp list entries of a base kernel (BaseKernels.py get_hyper_parameter_dimensionalities)."""
    if kind in ("WN", "L2", "L1"):
        return []
    if kind == "PER":
        sizes = [1, 1]
    elif kind in ("LIN", "SE_ARD"):
        sizes = [dim]
    else:
        sizes = [1]
    if scaled:
        sizes.append(1)
    return sizes


class CompiledProgram:
    def __init__(self, code: np.ndarray, entries: List[Tuple[int, int]], n_hp: int, dim: int, scaled: bool, tape: int = 0):
        self.tape = tape            # entries the interpreter's gradient tape would need (MAX_TAPE: interpreter only)
        self.code = code            # int32 [n_ops, 4]
        self.entries = entries      # (offset, size) of each hp list entry in the flat vector
        self.n_hp = n_hp
        self.dim = dim
        self.scaled = scaled

    @property
    def n_ops(self) -> int:
        return int(self.code.shape[0])

    def signature(self) -> bytes:
        return self.code.tobytes() + bytes([self.dim, int(self.scaled)])


def compile_spec(spec, dim: int, scaled: bool = False) -> CompiledProgram:
    ops: List[List[int]] = []
    entries: List[Tuple[int, int]] = []
    state = {"off": 0, "sp": 0, "max_sp": 0, "tape": 0}

    def push():
        state["sp"] += 1
        state["max_sp"] = max(state["max_sp"], state["sp"])

    def emit(node):
        kind = node[0]
        if kind in ("SE", "PER", "LIN", "MAT32", "MAT52", "WN", "SE_ARD", "L2", "L1"):
            sizes = leaf_entries(kind, dim, scaled)
            ops.append([OP[kind], state["off"], 1 if (scaled and kind not in ("WN", "L2", "L1")) else 0, 0])
            for s in sizes:
                entries.append((state["off"], s))
                state["off"] += s
            state["tape"] += sum(sizes)
            push()
            return
        children = node[1]
        if len(children) == 0:
            raise ValueError("operator without children")
        if kind in ("ADD", "MUL"):
            emit(children[0])
            for c in children[1:]:
                emit(c)
                ops.append([OP["ADD2" if kind == "ADD" else "MUL2"], 0, 0, 0])
                state["sp"] -= 1
                if kind == "MUL":
                    state["tape"] += 2
            return
        if kind == "CP":
            k = len(children)
            if k == 1:
                emit(children[0])
                return
            if dim != 1:
                raise ValueError("change-point kernels are defined for 1-d inputs only (Operators.py:398)")
            cp_off = state["off"]
            for _ in range(k - 1):
                entries.append((state["off"], 1))
                state["off"] += 1
            for i, c in enumerate(children):
                emit(c)
                ops.append([OP["CPW"], cp_off, i, k])
                state["tape"] += 4
                if i > 0:
                    ops.append([OP["ADD2"], 0, 0, 0])
                    state["sp"] -= 1
            return
        raise ValueError("unknown kernel node %r" % (kind,))

    emit(spec)
    if len(ops) > MAX_OPS:
        raise ValueError("kernel tree too large: %d ops > %d" % (len(ops), MAX_OPS))
    if state["max_sp"] > MAX_STACK:
        raise ValueError("kernel tree too deep: stack %d > %d" % (state["max_sp"], MAX_STACK))
    # MAX_TAPE limits the INTERPRETER's gradient sweep only: kernels specialised per program (csrc/jit.cu) have no tape,
    # so the limit is enforced by gpb_program_create, which knows whether the run-time compiler is available
    if state["off"] > MAX_HP:
        raise ValueError("too many hyper-parameters: %d > %d" % (state["off"], MAX_HP))
    if dim > MAX_DIM:
        raise ValueError("input dimensionality %d > %d" % (dim, MAX_DIM))
    return CompiledProgram(np.asarray(ops, dtype=np.int32).reshape(-1, 4), entries, state["off"], dim, scaled,
                           tape=state["tape"])


def flatten_hp(entries: List[Tuple[int, int]], hp_list, n_hp: int) -> np.ndarray:
    """reference hp list (scalars / [d] vectors; torch, numpy or python numbers) -> flat float64 vector"""
    if len(hp_list) != len(entries):
        raise AssertionError("Invalid hyper_param size: got %d entries, kernel consumes %d" % (len(hp_list), len(entries)))
    flat = np.zeros(n_hp, dtype=np.float64)
    for (off, size), h in zip(entries, hp_list):
        if hasattr(h, "detach"):
            h = h.detach().cpu().numpy()
        v = np.asarray(h, dtype=np.float64).reshape(-1)
        if v.size != size:
            raise AssertionError("hyper-parameter entry has %d values, expected %d" % (v.size, size))
        flat[off:off + size] = v
    return flat


def unflatten_grad(entries: List[Tuple[int, int]], flat: np.ndarray, like=None):
    """flat gradient -> list shaped like the reference hp list (scalars for size-1 entries unless `like` says [1])"""
    out = []
    for idx, (off, size) in enumerate(entries):
        g = np.array(flat[off:off + size], dtype=np.float64)
        shape = None
        if like is not None:
            h = like[idx]
            shape = tuple(h.shape) if hasattr(h, "shape") else ()
        if shape is None:
            shape = () if size == 1 else (size,)
        out.append(g.reshape(shape))
    return out
