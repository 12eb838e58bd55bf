"""Composite kernels (mirror of gpbasics/KernelBasics/Operators.py): ADD, MUL and the change-point operator CP, plus
the host-side tree utilities the reference's callers rely on (sorting, structural comparison, normal forms, hashing,
JSON export, change-point pruning).

No matrix arithmetic happens here: `to_spec()` describes the tree, the CUDA interpreter evaluates it in one pass
(csrc/program.cuh).  Hyper-parameters are consumed depth first in the *current* child order - sorting children
(`sort_child_nodes`, called implicitly by `type_compare_to`, Operators.py:188-189) therefore changes the order of the
hyper-parameter list exactly as it does in the reference.
"""
import logging
from typing import List, Tuple

import torch

from .. import global_parameters as global_param
from . import BaseKernels as bk
from . import Kernel as k

global_param.ensure_init()


def _as_float(t) -> float:
    return float(torch.as_tensor(t, dtype=torch.float64).reshape(-1)[0])


class Operator(k.Kernel):
    def __init__(self, manifestation: k.KernelManifestation, input_dimensionality: int, child_nodes: List[k.Kernel]):
        super().__init__(k.KernelType.OPERATOR, manifestation, input_dimensionality)
        self.child_nodes: List[k.Kernel] = child_nodes
        self.operator_sign = "UNKNOWN"
        if self.manifestation.value < 200:
            logging.critical("Invalid manifestation for Operator: %s", manifestation)
        self.sortable = False

    # ---- hyper-parameter bookkeeping: children own consecutive slices (Operators.py:94-105, :214-223) ----------
    def _leading(self) -> int:
        """entries that precede the children's slices (the change points of a CP node)"""
        return 0

    def child_slices(self) -> List[slice]:
        out, idx = [], self._leading()
        for cn in self.child_nodes:
            c = cn.get_number_of_hyper_parameter()
            out.append(slice(idx, idx + c))
            idx += c
        return out

    def get_number_of_hyper_parameter(self) -> int:
        return self._leading() + sum(cn.get_number_of_hyper_parameter() for cn in self.child_nodes)

    def get_default_hyper_parameter(self, xrange, n, from_distribution: bool = False):
        out = []
        for cn in self.child_nodes:
            out = out + cn.get_default_hyper_parameter(xrange, n, from_distribution)
        return out

    def set_last_hyper_parameter(self, last_hyper_parameter: List[torch.Tensor]):
        assert len(last_hyper_parameter) == self.get_number_of_hyper_parameter(), "Invalid hyper_param size: " + str(self)
        for cn, sl in zip(self.child_nodes, self.child_slices()):
            cn.set_last_hyper_parameter(list(last_hyper_parameter[sl]))

    def get_last_hyper_parameter(self, scaling_x_param=None):
        out = []
        for cn in self.child_nodes:
            out.extend(cn.get_last_hyper_parameter(scaling_x_param))
        return out

    def get_hyper_parameter_bounds(self, xrange, n):
        out = []
        for cn in self.child_nodes:
            out.extend(cn.get_hyper_parameter_bounds(xrange, n))
        return out

    def get_hyper_parameter_dimensionalities(self) -> List[list]:
        out = []
        for cn in self.child_nodes:
            out.extend(cn.get_hyper_parameter_dimensionalities())
        return out

    def get_hyper_parameter_distribution_definition(self, xrange, n) -> List[dict]:
        out = []
        for cn in self.child_nodes:
            out.extend(cn.get_hyper_parameter_distribution_definition(xrange, n))
        return out

    def get_hyper_parameter_names(self, kernel_id: int = -1) -> List[str]:
        names: List[str] = []
        for cn in self.child_nodes:
            names += cn.get_hyper_parameter_names(kernel_id)
            if isinstance(cn, bk.BaseKernel) and kernel_id >= 0:
                kernel_id += 1
        return names

    def _remember(self, hyper_parameter):
        # the reference's leaves record the slice they were evaluated with (BaseKernels.py:130, :292, :455)
        for cn, sl in zip(self.child_nodes, self.child_slices()):
            cn._remember(list(hyper_parameter[sl]))

    # ---- tree utilities ------------------------------------------------------------------------------------------
    def add_kernel(self, kernel):
        self.child_nodes = self.child_nodes + [kernel]

    def replace_child_node(self, index: int, new_child_node: k.Kernel):
        assert index < len(self.child_nodes), "cannot replace child node at index %d: only %d child nodes" % (
            index, len(self.child_nodes))
        self.child_nodes[index] = new_child_node

    def get_number_base_kernels(self) -> int:
        return sum(cn.get_number_base_kernels() for cn in self.child_nodes)

    def get_number_of_child_nodes(self) -> int:
        return len(self.child_nodes)

    def get_string_representation(self) -> str:
        if len(self.child_nodes) == 1:
            return self.child_nodes[0].get_string_representation()
        inner = (" " + self.operator_sign + " ").join(cn.get_string_representation() for cn in self.child_nodes)
        return "(" + inner + ")"

    def get_string_representation_weight(self) -> int:
        return sum(cn.get_string_representation_weight() for cn in self.child_nodes)

    def get_json(self) -> dict:
        return {"type": self.manifestation.name, "child_nodes": [cn.get_json() for cn in self.child_nodes]}

    def set_noise(self, noise):
        super().set_noise(noise)
        for cn in self.child_nodes:
            cn.set_noise(noise)

    def sort_child_nodes(self):
        if self.sortable:
            self.child_nodes = sorted(self.child_nodes, key=lambda node: node.get_string_representation_weight())
        for cn in self.child_nodes:
            if isinstance(cn, Operator):
                cn.sort_child_nodes()

    def set_dimensionality(self, input_dimensionality: int):
        super().set_dimensionality(input_dimensionality)
        for cn in self.child_nodes:
            cn.set_dimensionality(input_dimensionality)

    def get_simplified_version(self):
        raise NotImplementedError

    def type_compare_to(self, other):
        while isinstance(other, Operator) and len(other.child_nodes) == 1:
            other = other.child_nodes[0]
        if len(self.child_nodes) == 1:
            return self.child_nodes[0].type_compare_to(other)
        if not isinstance(other, Operator) or len(other.child_nodes) != len(self.child_nodes):
            return False
        self.sort_child_nodes()
        other.sort_child_nodes()
        return all(cn.type_compare_to(on) for cn, on in zip(self.child_nodes, other.child_nodes))

    def get_hash_tuple(self):
        return super().get_hash_tuple() + (sum(hash(cn) for cn in self.child_nodes),)

    def _copy_children(self):
        return [cn.deepcopy() for cn in self.child_nodes]


class _FoldOperator(Operator):
    """n-ary ADD / MUL: a left fold over the children (Operators.py:207-225, :306-326)"""
    SPEC = None

    def to_spec(self):
        if len(self.child_nodes) == 1:
            return self.child_nodes[0].to_spec()
        return (self.SPEC, [cn.to_spec() for cn in self.child_nodes])

    def deepcopy(self):
        other = type(self)(self.input_dimensionality, self._copy_children())
        if self.noise is not None:
            other.set_noise(self.noise)
        return other


class MultiplicationOperator(_FoldOperator):
    SPEC = "MUL"

    def __init__(self, input_dimensionality: int, child_nodes: List[k.Kernel]):
        super().__init__(k.KernelManifestation.MUL, input_dimensionality, child_nodes)
        self.operator_sign = "x"
        self.sortable = True

    def get_simplified_version(self):
        """flatten nested products and distribute over the first sum (Operators.py:271-297)"""
        flat, first_sum = [], None
        for cn in (c.get_simplified_version() for c in self.child_nodes):
            if isinstance(cn, MultiplicationOperator):
                flat.extend(cn.child_nodes)
            else:
                if isinstance(cn, AdditionOperator) and first_sum is None:
                    first_sum = len(flat)
                flat.append(cn)
        if first_sum is None:
            return MultiplicationOperator(self.input_dimensionality, flat)
        others = [c for i, c in enumerate(flat) if i != first_sum]
        terms = [MultiplicationOperator(self.input_dimensionality, others + [t]) for t in flat[first_sum].child_nodes]
        return AdditionOperator(self.input_dimensionality, terms).get_simplified_version()


class AdditionOperator(_FoldOperator):
    SPEC = "ADD"

    def __init__(self, input_dimensionality: int, child_nodes: List[k.Kernel]):
        super().__init__(k.KernelManifestation.ADD, input_dimensionality, child_nodes)
        self.operator_sign = "+"
        self.sortable = True

    def get_simplified_version(self):
        flat = []
        for cn in (c.get_simplified_version() for c in self.child_nodes):
            if isinstance(cn, AdditionOperator):
                flat.extend(cn.child_nodes)
            else:
                flat.append(cn)
        return AdditionOperator(self.input_dimensionality, flat)


class ChangePointOperator(Operator):
    """K = sum_i K_i * prev_i * (s_i s_i'^T), prev_{i+1} = (1-s_i)(1-s_i')^T, s_i(x) = 1[x < cp_i] (or its sigmoid /
    logistic surrogates, selected by global_param.p_cp_operator_type); hp = [cp_0..cp_{m-1}] ++ children
    (Operators.py:410-476, :451-453, :507-511).  1-d inputs only, as in the reference (:398)."""

    def __init__(self, input_dimensionality: int, child_nodes: List[k.Kernel], change_point_positions: List[torch.Tensor]):
        super().__init__(k.KernelManifestation.CP, input_dimensionality, child_nodes)
        assert (len(child_nodes) - 1) == len(change_point_positions), \
            "Error. Change Point positions and/or their positions wrongly initialized."
        self.operator_sign = "]["
        self.change_point_positions: List[torch.Tensor] = [k.as_scalar_tensor(c) for c in change_point_positions]

    def to_spec(self):
        if len(self.child_nodes) == 1:
            return self.child_nodes[0].to_spec()
        return ("CP", [cn.to_spec() for cn in self.child_nodes])

    def _leading(self) -> int:
        return len(self.change_point_positions)

    def get_hyper_parameter_dimensionalities(self) -> List[list]:
        # the reference reports one [m] entry here although the list carries m scalars;
        # the list form - what get_tf_tensor consumes - is followed
        return [[] for _ in self.change_point_positions] + super().get_hyper_parameter_dimensionalities()

    def set_last_hyper_parameter(self, last_hyper_parameter: List[torch.Tensor]):
        assert len(last_hyper_parameter) == self.get_number_of_hyper_parameter(), \
            "Invalid hyper_param size: %s" % str(last_hyper_parameter)
        if len(self.child_nodes) > 1:
            self.change_point_positions = list(last_hyper_parameter[0:len(self.change_point_positions)])
            self.last_hyper_parameter = self.change_point_positions
        for cn, sl in zip(self.child_nodes, self.child_slices()):
            cn.set_last_hyper_parameter(list(last_hyper_parameter[sl]))

    def get_default_hyper_parameter(self, xrange, n, from_distribution: bool = False):
        return list(self.change_point_positions) + super().get_default_hyper_parameter(xrange, n, from_distribution)

    def get_last_hyper_parameter(self, scaling_x_param=None):
        return list(self.change_point_positions) + super().get_last_hyper_parameter(scaling_x_param)

    def get_hyper_parameter_bounds(self, xrange, n):
        w = xrange[0][1] - xrange[0][0]
        lo = torch.tensor(xrange[0][0] - 1.5 * w, dtype=torch.float64)
        hi = torch.tensor(xrange[0][1] + 1.5 * w, dtype=torch.float64)
        return [(lo, hi)] * len(self.change_point_positions) + super().get_hyper_parameter_bounds(xrange, n)

    def add_kernel(self, kernel: k.Kernel, new_cp_position):
        assert kernel is not None, "Adding None as kernel to ChangePoint is not allowed."
        assert new_cp_position is not None
        assert len(self.change_point_positions) == 0 or \
            _as_float(new_cp_position) - _as_float(self.change_point_positions[-1]) > 0, \
            "New Changepoints _must_ be larger in value than the former largest change point."
        self.child_nodes.append(kernel)
        self.change_point_positions.append(k.as_scalar_tensor(new_cp_position))

    def add_preceding_kernel(self, kernel, new_cp_position):
        assert kernel is not None, "Adding None as kernel to ChangePoint is not allowed."
        assert new_cp_position is not None
        # the reference compares against the *last* change point here (Operators.py:530-534); kept
        assert len(self.change_point_positions) == 0 or \
            _as_float(self.change_point_positions[-1]) - _as_float(new_cp_position) > 0, \
            "A preceding change point must be smaller than the existing ones."
        self.child_nodes = [kernel] + self.child_nodes
        self.change_point_positions = [k.as_scalar_tensor(new_cp_position)] + self.change_point_positions

    def get_simplified_kernel(self, data_range: List[float]) -> Tuple[k.Kernel, bool]:
        """drop change points (and the segment they delimit) that left the data range or overtook their successor
        (Operators.py:538-587)"""
        cps = [_as_float(c) for c in self.change_point_positions]
        smooth = global_param.p_cp_operator_type == global_param.ChangePointOperatorType.SIGMOID
        blur = 4 if smooth else 0
        drop_cp, drop_child = set(), set()
        for i, c in enumerate(cps):
            if c >= data_range[1] + blur:
                drop_cp.add(i); drop_child.add(i + 1)
            if c <= data_range[0] - blur:
                drop_cp.add(i); drop_child.add(i)
            if i < len(cps) - 1 and c >= cps[i + 1]:
                drop_cp.add(i); drop_child.add(i + 1)
        if not drop_cp:
            return self, False
        new_cps = [self.change_point_positions[i].clone() for i in range(len(cps)) if i not in drop_cp]
        new_children = [cn for i, cn in enumerate(self.child_nodes) if i not in drop_child]
        assert len(new_cps) + 1 == len(new_children), "Error in get_simplified_kernel"
        return ChangePointOperator(self.input_dimensionality, new_children, new_cps), True

    def deepcopy(self):
        other = ChangePointOperator(self.input_dimensionality, self._copy_children(),
                                    [c.clone() for c in self.change_point_positions])
        if self.noise is not None:
            other.set_noise(self.noise)
        return other

    def get_json(self) -> dict:
        if len(self.child_nodes) == 1:
            return {"type": self.manifestation.name, "child_nodes": [self.child_nodes[0].get_json()]}
        cps = [_as_float(c) for c in self.change_point_positions]
        nodes = []
        for i, cn in enumerate(self.child_nodes):
            j = cn.get_json()
            j["start_index"] = 0 if i == 0 else cps[i - 1]
            j["stop_index"] = 1.0 if i == len(cps) else cps[i]
            nodes.append(j)
        return {"type": self.manifestation.name, "child_nodes": nodes}

    def get_simplified_version(self):
        return ChangePointOperator(self.input_dimensionality, [cn.get_simplified_version() for cn in self.child_nodes],
                                   self.change_point_positions)

    def get_hash_tuple(self):
        return k.Kernel.get_hash_tuple(self) + (sum(hash(cn) for cn in self.child_nodes),) + \
            tuple(hash(cn) for cn in self.child_nodes)
