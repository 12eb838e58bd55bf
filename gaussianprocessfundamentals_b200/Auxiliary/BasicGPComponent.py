"""Base class of kernels and mean functions (mirror of gpbasics/Auxiliary/BasicGPComponent.py:6-42)."""
from typing import List, Tuple

import torch


class Component:
    def get_hyper_parameter_bounds(self, xrange: List[List[float]], n: int) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        raise NotImplementedError

    def get_hyper_parameter_dimensionalities(self) -> List[list]:
        raise NotImplementedError

    def get_hyper_parameter_distribution_definition(self, xrange: List[List[float]], n: int) -> List[dict]:
        raise NotImplementedError

    @staticmethod
    def serialize_hyper_parameter(hyper_parameter: List[torch.Tensor]) -> torch.Tensor:
        return torch.cat([torch.as_tensor(h, dtype=torch.float64).reshape(-1) for h in hyper_parameter], dim=0)

    @staticmethod
    def deserialize_hyper_parameter(hyper_parameter: torch.Tensor, dimensionalities: List[list]) -> List[torch.Tensor]:
        """Inverse of serialize.  The reference always slices from offset 0 (BasicGPComponent.py:37);
        the evident intent - consecutive slices - is implemented."""
        out, index = [], 0
        flat = torch.as_tensor(hyper_parameter, dtype=torch.float64).reshape(-1)
        for dim in dimensionalities:
            size = 1 if len(dim) == 0 else int(dim[0])
            out.append(flat[index:index + size].reshape(list(dim)))
            index += size
        return out
