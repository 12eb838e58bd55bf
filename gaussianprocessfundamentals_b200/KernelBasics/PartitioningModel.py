"""Data partitioning (mirror of gpbasics/KernelBasics/PartitioningModel.py:12-151): per-partition row indices from a
score matrix and the partition-major re-ordering of a DataInput.  Integer bookkeeping - bit-exact with the reference
for SELF_SUFFICIENT criteria and for tie-free SMALLEST_DISTANCE scores (the reference breaks ties with unseeded noise,
PartitioningModel.py:118; pass `rng` to make that reproducible).  Host-side NumPy, negligible cost."""
import logging
from enum import Enum
from typing import List, Optional

import numpy as np
import torch

from .. import global_parameters as global_param
from ..DataHandling import DataInput as di

global_param.ensure_init()


class PartitioningClass(Enum):
    SELF_SUFFICIENT = 0,     # (sic) a 1-tuple in the reference as well (PartitioningModel.py:15)
    SMALLEST_DISTANCE = 1


class PartitionCriterion:
    def __init__(self, partitioning_type: PartitioningClass):
        self.partitioning_type = partitioning_type

    def get_score(self, x_vector: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def deepcopy(self):
        raise NotImplementedError

    def get_json(self) -> dict:
        raise NotImplementedError


class IntervalCriterion(PartitionCriterion):
    """score 1 inside [lower, upper) of one input dimension, else 0.  The reference's PartitionCriterion is abstract
    (PartitioningModel.py:21-32); this concrete criterion is what the block-partitioned benchmark (config C4) uses."""

    def __init__(self, lower: float, upper: float, dimension: int = 0):
        super().__init__(PartitioningClass.SELF_SUFFICIENT)
        self.lower, self.upper, self.dimension = float(lower), float(upper), int(dimension)

    def get_score(self, x_vector: np.ndarray) -> np.ndarray:
        col = np.asarray(x_vector, dtype=np.float64)[:, self.dimension]
        return np.logical_and(col >= self.lower, col < self.upper).astype(np.float64)

    def deepcopy(self):
        return IntervalCriterion(self.lower, self.upper, self.dimension)

    def get_json(self) -> dict:
        return {"type": "interval", "lower": self.lower, "upper": self.upper, "dimension": self.dimension}

    def __hash__(self):
        return hash((self.lower, self.upper, self.dimension))


class PartitioningModel:
    def __init__(self, partition_class: PartitioningClass, ignored_dimensions: List[int],
                 rng: Optional[np.random.Generator] = None):
        self.partitioning: List[PartitionCriterion] = []
        self.partition_class = partition_class
        self.ignored_dimensions = ignored_dimensions
        self.rng = rng

    def automatic_init_criteria(self, data_input, optimize_metric, model_selection_metric,
                                number_of_partitions: int = None, predecessor_criterion=None):
        pass

    def init_partitioning(self, partitioning: List[PartitionCriterion]):
        if len(self.partitioning) > 0:
            logging.warning("%s: Overwriting old partitioning." % str(self))
        self.partitioning = partitioning

    def get_number_of_partitions(self) -> int:
        return len(self.partitioning)

    def add_partitioning_criterion(self, criterion: PartitionCriterion):
        assert criterion is not None, "Criterion cannot be None"
        assert criterion.partitioning_type == self.partition_class, \
            "Partitioning Criterion does not match Partitioning Model"
        self.partitioning.append(criterion)

    def filter_data_by_ignored_dimensions(self, vector: np.ndarray):
        if len(self.ignored_dimensions) == 0:
            return vector
        assert vector.shape[1] > max(self.ignored_dimensions)
        keep = [i not in self.ignored_dimensions for i in range(vector.shape[1])]
        return vector[:, keep]

    def get_data_record_indices_per_partition(self, x_vector) -> List[np.ndarray]:
        x = np.asarray(x_vector.detach().cpu().numpy() if isinstance(x_vector, torch.Tensor) else x_vector)
        columns = [c.get_score(self.filter_data_by_ignored_dimensions(x)) for c in self.partitioning]
        score = np.transpose(np.array(columns))
        if self.partition_class == PartitioningClass.SMALLEST_DISTANCE:
            noise = (self.rng.normal(0, 1e-10, score.shape) if self.rng is not None
                     else np.random.normal(0, 1e-10, score.shape))
            score = score + noise
            score = score == np.amin(score, axis=1).reshape(-1, 1)
        per_partition = [np.where(score[:, i] == 1)[0] for i in range(self.get_number_of_partitions())]
        if len(per_partition) == 0:
            per_partition = [np.linspace(0, len(x) - 1, len(x), dtype=int)]
        return per_partition

    def partition_data_input(self, data_input: di.DataInput) -> di.PartitionedDataInput:
        if len(self.partitioning) <= 1:
            logging.warning("Dataset cannot be partitioned as only one / none partition criterion is available.")
            return di.PartitionedDataInput(data_input.data_x_train, data_input.data_y_train, data_input.data_x_test,
                                           data_input.data_y_test, [data_input])
        separate = not torch.equal(data_input.data_x_train, data_input.data_x_test)
        train_idx = self.get_data_record_indices_per_partition(data_input.data_x_train)
        test_idx = self.get_data_record_indices_per_partition(data_input.data_x_test) if separate else train_idx
        assert len(train_idx) == len(test_idx)
        blocks, xs, ys, xts, yts = [], [], [], [], []
        for tr, te in zip(train_idx, test_idx):
            tr_t, te_t = torch.as_tensor(tr, dtype=torch.long), torch.as_tensor(te, dtype=torch.long)
            bx, by = data_input.data_x_train[tr_t], data_input.data_y_train[tr_t]
            bxt = data_input.data_x_test[te_t]
            byt = data_input.data_y_test[te_t] if data_input.data_y_test is not None else None
            blocks.append(di.DataInput(data_x_train=bx, data_y_train=by, data_x_test=bxt, data_y_test=byt))
            xs.append(bx); ys.append(by); xts.append(bxt); yts.append(byt)
        y_test = torch.cat(yts, dim=0) if all(v is not None for v in yts) else None
        out = di.PartitionedDataInput(torch.cat(xs, dim=0), torch.cat(ys, dim=0), torch.cat(xts, dim=0), y_test, blocks)
        out.train_indices, out.test_indices = train_idx, test_idx
        return out

    def deepcopy(self):
        # the reference swaps the constructor arguments here (PartitioningModel.py:143-145)
        other = PartitioningModel(self.partition_class, list(self.ignored_dimensions), self.rng)
        other.init_partitioning([pc.deepcopy() for pc in self.partitioning])
        return other

    def get_hash_tuple(self):
        return tuple(self.ignored_dimensions) + (sum(hash(c) for c in self.partitioning),)

    def __hash__(self):
        return hash(self.get_hash_tuple())
