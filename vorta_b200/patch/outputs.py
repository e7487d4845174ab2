"""Output containers of the routed transformers and pipelines — ``vorta/patch/outputs.py:8-30``.

The reference derives them from diffusers' ``BaseOutput``; this restatement keeps the part callers rely on:
attribute access, ``out["sample"]``, ``out[0]`` and ``to_tuple()`` over the fields that are not ``None``."""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any, List, Optional, Tuple

import torch


class _Output:
    def to_tuple(self) -> Tuple[Any, ...]:
        return tuple(getattr(self, f.name) for f in fields(self) if getattr(self, f.name) is not None)

    def __getitem__(self, key):
        if isinstance(key, str):
            return getattr(self, key)
        return self.to_tuple()[key]


@dataclass
class RoutedTransformerModelOutput(_Output):
    sample: torch.Tensor
    reg_loss: Optional[torch.Tensor] = None
    last_layer_distill_loss: Optional[torch.Tensor] = None
    hidden_layer_distill_loss: Optional[torch.Tensor] = None
    routing_scores: Optional[List[torch.Tensor]] = None


@dataclass
class VideoPipelineOutput(_Output):
    """Denoised latents (``output_type="latent"``) and the routing scores of every step."""
    frames: torch.Tensor
    routing_scores: Optional[List[List[torch.Tensor]]] = None
