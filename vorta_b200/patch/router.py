"""Router head — the reference's ``vorta/patch/router.py:17-43`` (state-dict compatible: ``linear.weight`` /
``linear.bias``), with the forward pass on the sm_100a router kernel (fp32 arithmetic)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
from torch import nn

from .. import ops


class Router(nn.Module):
    def __init__(self, embedding_dim: int, heads: int, num_experts: int = 3):
        super().__init__()
        if num_experts != 3:
            raise ValueError("the routed attention path has exactly 3 experts (full, coreset, sliding tile)")
        self.heads = heads
        self.num_experts = num_experts
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, heads * num_experts, bias=True)
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, temb: torch.Tensor) -> torch.Tensor:
        """temb (B, E) -> routing score (B, H, 3) = softmax(linear(silu(temb)))."""
        scores, _ = ops.router_forward(temb, self.linear.weight, self.linear.bias, self.heads)
        return scores[0].to(temb.dtype)


def route_step(routers: Sequence[Router], temb: torch.Tensor, tau_sparse: Optional[float]
               ) -> Tuple[torch.Tensor, list]:
    """All routers of a denoise step in ONE launch (they only depend on ``temb``; modeling_wan.py:215).
    Returns scores (L, B, H, 3) on the device and the per-layer branch lists on the host (one D2H copy)."""
    weight = torch.stack([r.linear.weight for r in routers])
    bias = torch.stack([r.linear.bias for r in routers])
    scores, branch = ops.router_forward(temb, weight, bias, routers[0].heads, tau_sparse)
    return scores, branch.cpu().tolist()
