"""Raster <-> tile-major token order — same interface as the reference's ``vorta/attention/tile.py:7-78``.

The routed attention kernel never needs these as separate passes (the tile-major order is produced by the
gather that stages Q/K/V and undone by the attention epilogue's row map); they exist for callers that use the
reference's helpers directly.  ``sp_size`` is accepted for signature parity and must describe an already
assembled sequence: the reference's frame interleave for ``sp_size > 1`` (tile.py:20-24) is a by-product of its
frame sharding and is deliberately not reproduced (N-GPU result == 1-GPU result).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from .. import _lib as L
from .. import ops
from ._plans import get_plan

_MAPS = {}


def _maps(tile_size: Sequence[int], latent_shape: Sequence[int], device: torch.device):
    key = (tuple(tile_size), tuple(latent_shape), str(device))
    if key not in _MAPS:
        plan = get_plan(latent_shape, tile_size, (1, 1, 1), (1, 1, 1), n_unpooled=0)
        fwd = torch.from_numpy(plan.export(L.EXPORT_TILE_MAP)[:plan.seq_len].copy())
        inv = torch.empty_like(fwd)
        inv[fwd.long()] = torch.arange(fwd.numel(), dtype=torch.int32)
        _MAPS[key] = (fwd.to(device), inv.to(device))
    return _MAPS[key]


def _apply(sequence: torch.Tensor, row_map: torch.Tensor, head_dim: int) -> torch.Tensor:
    x = sequence if head_dim == 1 else sequence.transpose(1, 2)       # -> (B, H, S, D) view
    y = ops.gather_rows(x, row_map)
    return y if head_dim == 1 else y.transpose(1, 2)


def tile_layout(sequence: torch.Tensor, sp_size: int, tile_size: Tuple[int, int, int],
                latent_shape: Tuple[int, int, int], head_dim: int = 2) -> torch.Tensor:
    fwd, _ = _maps(tile_size, latent_shape, sequence.device)
    return _apply(sequence, fwd, head_dim)


def untile_layout(sequence: torch.Tensor, sp_size: int, tile_size: Tuple[int, int, int],
                  latent_shape: Tuple[int, int, int], head_dim: int = 2) -> torch.Tensor:
    _, inv = _maps(tile_size, latent_shape, sequence.device)
    return _apply(sequence, inv, head_dim)
