"""Coreset ("lowres") token selection — same names and argument meaning as the reference's
``vorta/attention/coreset_select.py`` (LowresGroupInfo :8-12, get_group_info :15-60, MatchingResults :62-65,
pool_sequence_by_similarity :68-124, unpool_sequence_by_similarity :127-185), computed by the sm_100a kernels
behind the C ABI (vb_plan_create / vb_coreset_select / vb_coreset_tables / vb_gather_rows).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib as L
from .. import ops
from ._plans import plan_for_group_info


@dataclass
class LowresGroupInfo:
    center_indices: torch.Tensor          # (G, 1) int64
    margin_indices: torch.Tensor          # (G, g-1) int64
    num_unpooled_tokens_per_group: int
    # extras the reference dataclass does not carry; they let the processors find the cached plan without
    # re-deriving the geometry from the index tables
    latent_video_shape: Optional[Tuple[int, int, int]] = None
    compress_window_size: Optional[Tuple[int, int, int]] = None


def get_group_info(latent_video_shape: Sequence[int], compress_window_size: Sequence[int],
                   reduction_rate: float = 0.5, device: torch.device = torch.device("cpu")) -> LowresGroupInfo:
    """Group tables from the plan's closed-form builder (host integer code in the C library)."""
    lat = tuple(int(x) for x in latent_video_shape)
    win = tuple(int(x) for x in compress_window_size)
    g = int(np.prod(win))
    n_unpooled = int(g * (1 - reduction_rate)) - 1            # coreset_select.py:54
    # the group tables do not depend on the sliding-tile parameters; use a trivial tiling for this plan
    plan = ops.Plan(lat, (1, 1, 1), (1, 1, 1), win, n_unpooled=n_unpooled)
    center = torch.from_numpy(plan.export(L.EXPORT_CENTER_INDICES)).reshape(-1, 1)
    margin = torch.from_numpy(plan.export(L.EXPORT_MARGIN_INDICES)).reshape(center.shape[0], g - 1)
    return LowresGroupInfo(center.to(device), margin.to(device), n_unpooled, lat, win)


@dataclass
class MatchingResults:
    unpooled_argsort_sim: torch.Tensor    # (B, h, G, n_unpooled) int64: kept margins, ascending similarity
    pooled_argsort_sim: torch.Tensor      # (B, h, G, g-1-n_unpooled) int64: dropped margins


def pool_sequence_by_similarity(hidden_states: torch.Tensor, lowres_group_info: LowresGroupInfo,
                                matching_results: Optional[MatchingResults] = None,
                                latent_shape: Optional[Sequence[int]] = None
                                ) -> Tuple[torch.Tensor, MatchingResults]:
    """(B, h, S, 128) -> (B, h, S_c, 128) = [centres | kept margins], plus the matching."""
    plan = plan_for_group_info(lowres_group_info, latent_shape)
    if matching_results is None:
        un, po, kept, _ = ops.coreset_select(plan, hidden_states, want_tokens=True)
        matching_results = MatchingResults(un, po)
    else:
        kept, _, _ = ops.coreset_tables(plan, matching_results.unpooled_argsort_sim,
                                        matching_results.pooled_argsort_sim)
    pooled = ops.gather_rows(hidden_states, kept, n_rows=plan.coreset_len)
    return pooled, matching_results


def unpool_sequence_by_similarity(pooled_hidden_states: torch.Tensor, lowres_group_info: LowresGroupInfo,
                                  matching_results: MatchingResults,
                                  latent_shape: Optional[Sequence[int]] = None) -> torch.Tensor:
    """(B, h, S_c, 128) -> (B, h, S, 128): kept tokens get their own row, dropped margins their centre's."""
    plan = plan_for_group_info(lowres_group_info, latent_shape)
    _, _, src = ops.coreset_tables(plan, matching_results.unpooled_argsort_sim, matching_results.pooled_argsort_sim)
    return ops.gather_rows(pooled_hidden_states, src)
