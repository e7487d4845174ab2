"""Plan cache: one ``ops.Plan`` per geometry, shared by every layer and step."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import ops

_CACHE: Dict[tuple, ops.Plan] = {}


def _t3(x) -> Tuple[int, int, int]:
    return tuple(int(v) for v in x)


def infer_lowres_window(group_info, latent_shape: Sequence[int]) -> Tuple[int, int, int]:
    """Recover the (f, h, w) group window from a reference-style LowresGroupInfo: group 0 starts at the origin,
    so the largest coordinate among its members + 1 is the window (coreset_select.py:43-49)."""
    win = getattr(group_info, "compress_window_size", None)
    if win is not None:
        return _t3(win)
    T, H, W = _t3(latent_shape)
    toks = np.concatenate([group_info.center_indices[0].detach().cpu().numpy().reshape(-1),
                           group_info.margin_indices[0].detach().cpu().numpy().reshape(-1)])
    return (int((toks // (H * W)).max()) + 1, int(((toks % (H * W)) // W).max()) + 1, int((toks % W).max()) + 1)


def get_plan(latent_shape, tile_size, window_size, lowres_window, n_unpooled: int, text_len: int = 0,
             text_valid: int = 0) -> ops.Plan:
    # the plan's device tables live on the device that is current when it is created
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    # text_valid is part of the key: prompts of different length (true CFG alternates two) each keep their own
    # schedule tables instead of rebuilding one plan with a device synchronize on every switch
    key = (_t3(latent_shape), _t3(tile_size), _t3(window_size), _t3(lowres_window), int(n_unpooled), int(text_len),
           int(text_valid), dev)
    plan = _CACHE.get(key)
    if plan is None:
        if len(_CACHE) >= 64:                      # bounded: drop the oldest geometry
            _CACHE.pop(next(iter(_CACHE)))
        plan = ops.Plan(key[0], key[1], key[2], key[3], n_unpooled=int(n_unpooled), text_len=int(text_len),
                        text_valid=int(text_valid))
        _CACHE[key] = plan
    return plan


def plan_for_group_info(group_info, latent_shape: Optional[Sequence[int]] = None) -> ops.Plan:
    """Plan that only serves the coreset helpers (trivial tiling)."""
    lat = getattr(group_info, "latent_video_shape", None) or latent_shape
    if lat is None:
        raise ValueError("latent_shape is required when the LowresGroupInfo does not carry it")
    win = infer_lowres_window(group_info, lat)
    return get_plan(lat, (1, 1, 1), (1, 1, 1), win, group_info.num_unpooled_tokens_per_group)


def clear() -> None:
    _CACHE.clear()
