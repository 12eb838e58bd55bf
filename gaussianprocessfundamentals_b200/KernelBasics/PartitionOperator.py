"""Block-diagonal kernel over a data partition (mirror of gpbasics/KernelBasics/PartitionOperator.py:15-123).

The reference densifies the blocks with LinearOperatorBlockDiag(...).to_dense(); here every non-empty block is one
launch of the fused assembly kernel into its diagonal sub-block of a zeroed matrix, and the likelihood path never
densifies at all: PartitionedGaussianProcess evaluates the blocks as one batched plan.  Hyper-parameters are sliced
consecutively per child and the index advances even for empty blocks (PartitionOperator.py:63-82)."""
from typing import List

import numpy as np
import torch

from .. import global_parameters as global_param
from . import Kernel as k
from . import Operators as op
from . import PartitioningModel as pm

global_param.ensure_init()


class PartitionOperator(op.Operator):
    def __init__(self, input_dimensionality: int, child_nodes: List[k.Kernel], partitioning_model: pm.PartitioningModel):
        assert len(child_nodes) == partitioning_model.get_number_of_partitions(), \
            "One partitioning criterion for each kernel has to be supplied"
        super().__init__(k.KernelManifestation.PART, input_dimensionality, child_nodes)
        self.operator_sign = "|"
        self.partitioning_model = partitioning_model

    def to_spec(self):
        raise NotImplementedError("a partition operator is evaluated block by block (see get_list_of_block_matrices)")

    def get_list_of_block_matrices(self, hyper_parameter, x_vector, x_vector_):
        indices = self.partitioning_model.get_data_record_indices_per_partition(x_vector)
        indices_ = indices if x_vector is x_vector_ else \
            self.partitioning_model.get_data_record_indices_per_partition(x_vector_)
        assert len(indices) == len(indices_) == len(self.child_nodes)
        x = torch.as_tensor(x_vector, dtype=torch.float64)
        x_ = torch.as_tensor(x_vector_, dtype=torch.float64)
        square, blocks = True, []
        for cn, sl, a, b in zip(self.child_nodes, self.child_slices(), indices, indices_):
            if len(a) == 0 or len(b) == 0:
                if len(a) != len(b):
                    blocks.append([len(a), len(b)])
            else:
                xa = x[torch.as_tensor(a, dtype=torch.long)]
                xb = xa if (x_vector is x_vector_) else x_[torch.as_tensor(b, dtype=torch.long)]
                blocks.append(cn.get_tf_tensor(list(hyper_parameter[sl]), xa, xb))
            if len(a) != len(b):
                square = False
        return blocks, square, indices, indices_

    def get_tf_tensor(self, hyper_parameter, x_vector, x_vector_) -> torch.Tensor:
        assert len(hyper_parameter) == self.get_number_of_hyper_parameter(), "Invalid hyper_param size: " + str(self)
        blocks, _, _, _ = self.get_list_of_block_matrices(hyper_parameter, x_vector, x_vector_)
        rows = sum(b.shape[0] if isinstance(b, torch.Tensor) else b[0] for b in blocks)
        cols = sum(b.shape[1] if isinstance(b, torch.Tensor) else b[1] for b in blocks)
        n, m = len(x_vector), len(x_vector_)
        assert rows == n and cols == m, "x_vector.shape=%s, result_shape=%s" % (str((n, m)), str((rows, cols)))
        out = torch.zeros((n, m), dtype=torch.float64, device="cuda")
        r = c = 0
        for b in blocks:
            if isinstance(b, torch.Tensor):
                out[r:r + b.shape[0], c:c + b.shape[1]] = b
                r, c = r + b.shape[0], c + b.shape[1]
            else:
                r, c = r + b[0], c + b[1]
        return out

    def add_kernel(self, kernel: k.Kernel, criterion: pm.PartitionCriterion):
        assert kernel is not None, "Adding None as kernel is not allowed."
        self.child_nodes.append(kernel)
        self.partitioning_model.add_partitioning_criterion(criterion)

    def deepcopy(self):
        other = PartitionOperator(self.input_dimensionality, self._copy_children(), self.partitioning_model.deepcopy())
        if self.noise is not None:
            other.set_noise(self.noise)
        return other

    def get_json(self) -> dict:
        if len(self.child_nodes) == 1:
            return {"type": self.manifestation.name, "child_nodes": [self.child_nodes[0].get_json()]}
        nodes = []
        for cn, crit in zip(self.child_nodes, self.partitioning_model.partitioning):
            j = cn.get_json()
            j["partitioning_criterion"] = crit.get_json()
            nodes.append(j)
        return {"type": self.manifestation.name, "child_nodes": nodes}

    def get_simplified_version(self):
        return PartitionOperator(self.input_dimensionality, [cn.get_simplified_version() for cn in self.child_nodes],
                                 self.partitioning_model)

    def get_hash_tuple(self):
        return k.Kernel.get_hash_tuple(self) + (sum(hash(cn) for cn in self.child_nodes),) + \
            tuple(hash(cn) for cn in self.child_nodes) + (hash(self.partitioning_model),)
