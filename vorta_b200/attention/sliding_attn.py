"""Sliding-tile attention — interface of the reference's ``vorta/attention/sliding_attn_flex.py``.

``create_sliding_tile_attn_mask_func`` (:72-134) returns a ``SlidingTileSchedule`` instead of a torch
``BlockMask``: the 3-D window of every query tile is closed-form, so the schedule is a table of contiguous key
runs in tile-major order (built inside ``vb_plan_create``) and no (S/128)^2 mask is ever evaluated.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple, Union

import numpy as np
import torch

from .. import _lib as L
from .. import ops
from ._plans import get_plan


@dataclass
class SlidingTileSchedule:
    """What the reference passes around as ``flex_attn_mask_func`` (a BlockMask): geometry + text lengths."""
    latent_shape: Tuple[int, int, int]
    window_size: Tuple[int, int, int]
    tile_size: Tuple[int, int, int]
    text_seq_length: int = 0
    text_seq_length_no_pad: int = 0

    def plan(self) -> ops.Plan:
        return get_plan(self.latent_shape, self.tile_size, self.window_size, (1, 1, 1), 0, self.text_seq_length,
                        self.text_seq_length_no_pad)

    def tile_windows(self) -> np.ndarray:
        """(num_tiles, 6) lo/hi tile coordinates of every query tile's key window."""
        return self.plan().export(L.EXPORT_TILE_WINDOW).reshape(-1, 6)

    def key_runs(self) -> np.ndarray:
        return self.plan().export(L.EXPORT_SLIDING_RUNS).reshape(-1, 2)


def create_sliding_tile_attn_mask_func(latent_shape, window_size, tile_size, text_seq_length: int,
                                       text_seq_length_no_pad: int, device=None) -> SlidingTileSchedule:
    sched = SlidingTileSchedule(tuple(int(x) for x in latent_shape), tuple(int(x) for x in window_size),
                                tuple(int(x) for x in tile_size), int(text_seq_length), int(text_seq_length_no_pad))
    sched.plan()        # validates (tile divides latent) like the reference's first-block _check_input
    return sched


def sliding_tile_flex_attn(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                           flex_attn_func: Optional[Callable] = None,
                           encoder_query: Optional[torch.Tensor] = None, encoder_key: Optional[torch.Tensor] = None,
                           encoder_value: Optional[torch.Tensor] = None,
                           tile_size: Tuple[int, int, int] = (6, 8, 8), latent_shape: Tuple[int, int, int] = (30, 45, 80),
                           head_dim: int = 2, window_size: Tuple[int, int, int] = (3, 3, 3),
                           text_seq_length_no_pad: Optional[int] = None
                           ) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """Signature of sliding_attn_flex.py:137-211.  ``flex_attn_func`` is accepted for parity: when it is a
    ``functools.partial(..., block_mask=SlidingTileSchedule)`` (what the processors pass) the schedule is taken
    from it; otherwise ``window_size`` / ``text_seq_length_no_pad`` are used.  Inputs / outputs in raster order,
    (B, H, S, D) for head_dim=1 or (B, S, H, D) for head_dim=2."""
    sched = None
    if flex_attn_func is not None:
        sched = getattr(flex_attn_func, "keywords", {}).get("block_mask")
    is_mmdit = encoder_query is not None
    text_len = encoder_query.shape[2 if head_dim == 1 else 1] if is_mmdit else 0
    if isinstance(sched, SlidingTileSchedule):
        window_size, text_valid = sched.window_size, sched.text_seq_length_no_pad
    else:
        text_valid = text_len if text_seq_length_no_pad is None else int(text_seq_length_no_pad)
    if head_dim == 2:
        query, key, value = (t.transpose(1, 2) for t in (query, key, value))
        if is_mmdit:
            encoder_query, encoder_key, encoder_value = (t.transpose(1, 2) for t in
                                                         (encoder_query, encoder_key, encoder_value))
    if is_mmdit:
        query = torch.cat([query, encoder_query], dim=2)
        key = torch.cat([key, encoder_key], dim=2)
        value = torch.cat([value, encoder_value], dim=2)
    plan = get_plan(latent_shape, tile_size, window_size, (1, 1, 1), 0, text_len, text_valid)
    H = query.shape[1]
    out = ops.routed_attention(plan, query, key, value, branch=[L.BRANCH_SLIDING] * H)
    if is_mmdit:
        video, text = out[:, :, :-text_len], out[:, :, -text_len:]
        if head_dim == 2:
            return video.transpose(1, 2), text.transpose(1, 2)
        return video, text
    return out.transpose(1, 2) if head_dim == 2 else out
