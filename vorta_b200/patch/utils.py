"""Processor kwargs preparation — the entry points of the reference's ``vorta/patch/utils.py``
(prepare_wan_self_attn_kwargs :8-36, prepare_hunyuan_self_attn_kwargs :39-56, wan_/hunyuan_pixel2token :59-95).

The config dict the reference scripts pass (``latent_shape, window_size, tile_size, lowres_window_size,
lowres_reduction_rate``) is rewritten IN PLACE into the kwargs every self-attention processor call receives:
the two ``lowres_*`` keys are replaced by ``lowres_group_info``; Wan additionally gets ``flex_attn_mask_func`` (text
lengths 0); ``tau_sparse`` is added when given.  Same keys in, same keys out as the reference.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple

import torch

from ..attention import create_sliding_tile_attn_mask_func, get_group_info


def _swap_in_group_info(kw: Dict[str, Any], device: torch.device, tau_sparse: Optional[float]) -> Dict[str, Any]:
    window, rate = kw.pop("lowres_window_size"), kw.pop("lowres_reduction_rate")
    kw["lowres_group_info"] = get_group_info(kw["latent_shape"], window, reduction_rate=rate, device=device)
    if tau_sparse is not None:
        kw["tau_sparse"] = tau_sparse
    return kw


def prepare_wan_self_attn_kwargs(self_attention_kwargs: Dict[str, Any], device: torch.device,
                                 tau_sparse: Optional[float] = None) -> Dict[str, Any]:
    kw = _swap_in_group_info(self_attention_kwargs, device, tau_sparse)
    # a schedule handle here (a BlockMask in the reference); Wan has no text tokens in its self-attention
    kw["flex_attn_mask_func"] = create_sliding_tile_attn_mask_func(
        latent_shape=kw["latent_shape"], window_size=kw["window_size"], tile_size=kw["tile_size"],
        text_seq_length=0, text_seq_length_no_pad=0, device=device)
    return kw


def prepare_hunyuan_self_attn_kwargs(self_attention_kwargs: Dict[str, Any], device: torch.device,
                                     tau_sparse: Optional[float] = None) -> Dict[str, Any]:
    # the HunyuanVideo mask depends on the prompt's text length and is built per call (pipeline_hunyuan.py:380-392)
    return _swap_in_group_info(self_attention_kwargs, device, tau_sparse)


def _pixels_to_tokens(video_shape: Sequence[int], strides: Tuple[int, int, int]) -> Tuple[int, int, int]:
    """(frames, height, width) in pixels -> latent token grid.  Each axis shrinks by VAE stride x patch size; a
    remainder of exactly one is the causal VAE's leading frame (77 f -> 20, 81 f -> 21); anything else is an error."""
    grid = []
    for pixels, stride in zip(video_shape, strides):
        extra = pixels % stride
        if extra > 1:
            raise ValueError(f"Number of pixel {pixels} is not a multiple of pixel2token {stride}.")
        grid.append(pixels // stride + extra)
    return tuple(grid)


def wan_pixel2token(video_shape: Sequence[int]) -> Tuple[int, int, int]:
    """Wan 2.1: VAE 4x temporal / 8x spatial, patch (1, 2, 2)."""
    return _pixels_to_tokens(video_shape, (4, 16, 16))


def hunyuan_pixel2token(video_shape: Sequence[int]) -> Tuple[int, int, int]:
    """HunyuanVideo: the same strides as Wan."""
    return _pixels_to_tokens(video_shape, (4, 16, 16))
