"""Cost-balanced head placement for the Ulysses exchange.

The reference gives rank r the contiguous heads [r*H/P, (r+1)*H/P) (vorta/ulysses/utils.py:60-66).  With routed
attention the heads of one layer differ in cost by ~6x (full : coreset : sliding), every rank waits for the slowest
one at the "out" exchange, and the routing of the whole step is known before the first block runs — so the host
picks, per layer, which H/P heads each rank receives.  The choice only changes slot numbers inside the exchange
buffers (``head_at`` of ``vb_ulysses_*``; ``out_heads`` of ``vb_attn_fwd``): no extra bytes move, and results are
bit-identical to the contiguous placement because heads are independent.

``balance_heads`` is deterministic and depends only on (branch ids, costs, P): every rank derives the same table
from its own copy of the routing decisions.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

# measured kernel efficiency relative to the full branch plus the branch's gather passes
# (profiles/r1_sweep_attn_8k_128k.csv, total_tflops column)
_BRANCH_OVERHEAD = (1.0, 1.12, 1.25)


def enabled() -> bool:
    return os.environ.get("VB_ULYSSES_BALANCE", "1") != "0"


def branch_costs(plan) -> List[float]:
    """Relative cost of one head of each branch for this geometry (algorithmic FLOPs x measured overhead)."""
    return [plan.flops_per_head(e) * _BRANCH_OVERHEAD[e] for e in range(3)]


def balance_heads(branch: Sequence[int], costs: Sequence[float], world: int) -> Optional[List[int]]:
    """Longest-processing-time placement of H heads on ``world`` ranks, H/world heads each.

    Returns ``head_at`` with ``head_at[r * H/P + i]`` = head held in slot i of rank r (ascending head order inside
    a rank), or None when the contiguous placement is already as good (identity tables skip the host work)."""
    H = len(branch)
    if world <= 1 or H % world != 0:
        return None
    hp = H // world
    cost = [float(costs[int(e)]) if 0 <= int(e) < len(costs) else 0.0 for e in branch]
    contiguous = max(sum(cost[r * hp:(r + 1) * hp]) for r in range(world))
    order = sorted(range(H), key=lambda h: (-cost[h], h))
    load = [0.0] * world
    held: List[List[int]] = [[] for _ in range(world)]
    for h in order:
        r = min((r for r in range(world) if len(held[r]) < hp), key=lambda r: (load[r], r))
        held[r].append(h)
        load[r] += cost[h]
    if max(load) >= contiguous * (1.0 - 1e-9):
        return None
    return [h for r in range(world) for h in sorted(held[r])]
