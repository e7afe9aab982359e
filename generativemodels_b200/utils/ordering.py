"""Sequence orderings for the VQ-VAE + transformer pipeline — interface of ``generative/utils/ordering.py``.

An ``Ordering`` maps the flattened voxel index of a 2-D / 3-D latent grid to its position in the 1-D token sequence
(raster scan, boustrophedon "s_curve", or a random permutation), after optional transpositions / 90-degree rotations
/ reflections of the index grid applied in ``transformation_order``.  Host-side index bookkeeping: numpy, computed
once per ordering; the permutations are applied to token tensors with plain indexing.
"""
from __future__ import annotations

import numpy as np
import torch

from .enums import OrderingTransformations, OrderingType


class Ordering:
    def __init__(self, ordering_type: str, spatial_dims: int, dimensions: tuple, reflected_spatial_dims: tuple = (),
                 transpositions_axes: tuple = (), rot90_axes: tuple = (),
                 transformation_order: tuple = (OrderingTransformations.TRANSPOSE.value,
                                                OrderingTransformations.ROTATE_90.value,
                                                OrderingTransformations.REFLECT.value)) -> None:
        super().__init__()
        self.ordering_type = ordering_type
        if self.ordering_type not in list(OrderingType):
            raise ValueError(f"ordering_type must be one of the following {list(OrderingType)}, but got "
                             f"{self.ordering_type}.")
        self.spatial_dims = spatial_dims
        self.dimensions = dimensions
        if len(dimensions) != self.spatial_dims + 1:
            raise ValueError(f"dimensions must be of length {self.spatial_dims + 1}, but got {len(dimensions)}.")
        self.reflected_spatial_dims = reflected_spatial_dims
        self.transpositions_axes = transpositions_axes
        self.rot90_axes = rot90_axes
        if len(set(transformation_order)) != len(transformation_order):
            raise ValueError(f"No duplicates are allowed. Received {transformation_order}.")
        for transformation in transformation_order:
            if transformation not in list(OrderingTransformations):
                raise ValueError(f"Valid transformations are {list(OrderingTransformations)} but received "
                                 f"{transformation}.")
        self.transformation_order = transformation_order
        self.template = self._transformed_template()
        self._sequence_ordering = self._scan(self.template)
        self._revert_sequence_ordering = np.argsort(self._sequence_ordering)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return x[self._sequence_ordering]

    def get_sequence_ordering(self) -> np.ndarray:
        return self._sequence_ordering

    def get_revert_sequence_ordering(self) -> np.ndarray:
        return self._revert_sequence_ordering

    # ------------------------------------------------------------------------------------------
    def _transformed_template(self) -> np.ndarray:
        grid = np.arange(int(np.prod(self.dimensions[1:]))).reshape(*self.dimensions[1:])
        for step in self.transformation_order:
            if step == OrderingTransformations.TRANSPOSE.value:
                for axes in self.transpositions_axes:
                    grid = np.transpose(grid, axes=axes)
            elif step == OrderingTransformations.ROTATE_90.value:
                for axes in self.rot90_axes:
                    grid = np.rot90(grid, axes=axes)
            elif step == OrderingTransformations.REFLECT.value:
                for axis, flip in enumerate(self.reflected_spatial_dims):
                    if flip:
                        grid = np.flip(grid, axis=axis)
        return grid

    def _scan(self, grid: np.ndarray) -> np.ndarray:
        coords = {OrderingType.RASTER_SCAN.value: self.raster_scan_idx, OrderingType.S_CURVE.value: self.s_curve_idx,
                  OrderingType.RANDOM.value: self.random_idx}[str(self.ordering_type)](
            grid.shape[0], grid.shape[1], grid.shape[2] if self.spatial_dims == 3 else None)
        return np.array([grid[tuple(c)] for c in coords])

    @staticmethod
    def raster_scan_idx(rows: int, cols: int, depths: int | None = None) -> np.ndarray:
        if depths:
            return np.stack(np.meshgrid(np.arange(rows), np.arange(cols), np.arange(depths), indexing="ij"),
                            -1).reshape(-1, 3)
        return np.stack(np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij"), -1).reshape(-1, 2)

    @staticmethod
    def s_curve_idx(rows: int, cols: int, depths: int | None = None) -> np.ndarray:
        out = []
        for r in range(rows):
            for c in (range(cols) if r % 2 == 0 else reversed(range(cols))):
                if depths:
                    out.extend((r, c, d) for d in (range(depths) if c % 2 == 0 else reversed(range(depths))))
                else:
                    out.append((r, c))
        return np.array(out)

    @staticmethod
    def random_idx(rows: int, cols: int, depths: int | None = None) -> np.ndarray:
        idx = Ordering.raster_scan_idx(rows, cols, depths)
        np.random.shuffle(idx)          # numpy's global generator, as in the reference
        return idx
