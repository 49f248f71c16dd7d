"""Grid plugin classes with the reference's surface (grids/*.py): raymarch() / interpolate()."""
from .base import BLASGrid, HashGridBase
from .occtree import Occtree
from .permuto_grid import PermutoGrid, PermutoEncoding
from .hash_grid_tinycudann import HashGridTinyCudaNN
from .hash_grid_torch import HashGridTorch

__all__ = ["BLASGrid", "HashGridBase", "Occtree", "PermutoGrid", "PermutoEncoding", "HashGridTinyCudaNN", "HashGridTorch"]
