"""Ulysses exchange over NVLink / NVSwitch peer memory, fused into the kernels on either side of the attention.

The reference (vorta/ulysses/utils.py:15-93) and the NCCL path of this package move Q, K, V and O with all-to-all
collectives plus layout copies.  Here every rank maps every peer's receive buffers (torch symmetric memory, one
allocation per geometry) and

  * "in"  : ``vb_ulysses_scatter_qkv`` stores each (token, head) row of Q, K, V straight into the buffer of the rank that
            owns the head — one pass, no staging buffer, no collective;
  * "out" : the attention kernel's epilogue stores each output row straight into the buffer of the rank that owns the
            token (``vb_attn_args.out_peer_ptrs``), already in the (S_loc, H, 128) layout the output projection reads.

Two device-side barriers per layer order the remote stores against their consumers (see ``PeerExchange``).  Used for
the top-1 (Eval) processors without a text segment; everything else takes the NCCL path in ``utils.py``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .. import _lib as L
from .parallel_states import SP_STATE

_EXCHANGES: Dict[tuple, "PeerExchange"] = {}
_DISABLED_REASON: Optional[str] = None


class PeerExchange:
    """Symmetric receive buffers of one geometry (heads H, local tokens S_loc, world P), all 128-channel bf16:
    ``qkv`` (3, S, H/P, 128): this rank's head chunk over the full sequence, written by every peer;
    ``out`` (S_loc, H, 128): this rank's token shard over all heads, written by every peer's attention epilogue.

    Ordering per layer (all on the caller's stream):
        scatter_qkv -> barrier A -> attention (reads qkv, stores into peers' out) -> barrier B -> consumer reads out.
    A rank leaves barrier B only after every rank finished its attention, so the next layer's scatter cannot overwrite
    a qkv buffer that is still being read; it reaches the next barrier A only after its own consumer of ``out`` was
    issued, so nobody stores into ``out`` while it is still needed."""

    def __init__(self, heads: int, s_loc: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm
        P, rank = SP_STATE.sp_size, SP_STATE.group_local_rank
        self.P, self.rank, self.heads, self.s_loc, self.hp = P, rank, heads, s_loc, heads // P
        self.S = s_loc * P
        group = SP_STATE.group if SP_STATE.group is not None else dist.group.WORLD
        self.qkv = symm.empty((3, self.S, self.hp, 128), dtype=torch.bfloat16, device=device)
        self.out = symm.empty((s_loc, heads, 128), dtype=torch.bfloat16, device=device)
        self.h_qkv = symm.rendezvous(self.qkv, group)
        self.h_out = symm.rendezvous(self.out, group)
        self.qkv_ptrs = (C.c_void_p * P)(*[int(p) for p in self.h_qkv.buffer_ptrs])
        self.out_ptrs = [int(p) for p in self.h_out.buffer_ptrs]

    def scatter_qkv(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, head_at: Optional[Sequence[int]] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """q, k, v: (1, H, S_loc, 128) views of this rank's token shard.  Returns (1, H/P, S, 128) views of the local
        receive buffer, valid after barrier A (issued here).  ``head_at``: slot -> head table of a balanced placement
        (``balance.balance_heads``), None = contiguous head chunks."""
        i64x3 = C.c_int64 * 3
        table = (C.c_int32 * self.heads)(*[int(h) for h in head_at]) if head_at is not None else None
        with torch.cuda.device(q.device):
            L.check(L.lib().vb_ulysses_scatter_qkv(
                q.data_ptr(), k.data_ptr(), v.data_ptr(), i64x3(q.stride(2), k.stride(2), v.stride(2)),
                i64x3(q.stride(1), k.stride(1), v.stride(1)), self.qkv_ptrs, self.S, self.s_loc, self.heads, self.P,
                self.rank, table, torch.cuda.current_stream(q.device).cuda_stream))
        self.h_qkv.barrier(channel=0)
        return tuple(self.qkv[i].unsqueeze(0).transpose(1, 2) for i in range(3))

    def finish_out(self) -> torch.Tensor:
        """Barrier B, then this rank's (1, H, S_loc, 128) view of the gathered outputs."""
        self.h_out.barrier(channel=1)
        return self.out.unsqueeze(0).transpose(1, 2)


def get_exchange(heads: int, s_loc: int, device: torch.device) -> Optional[PeerExchange]:
    """The cached exchange for this geometry, or None when peer memory is unavailable / disabled
    (VB_ULYSSES=nccl), in which case the caller takes the NCCL path."""
    global _DISABLED_REASON
    if not SP_STATE.enabled or os.environ.get("VB_ULYSSES", "peer") == "nccl" or _DISABLED_REASON is not None:
        return None
    if SP_STATE.sp_size > 8 or heads % SP_STATE.sp_size != 0:
        return None
    key = (heads, s_loc, str(device), SP_STATE.sp_size)
    ex = _EXCHANGES.get(key)
    if ex is None:
        try:
            ex = PeerExchange(heads, s_loc, device)
        except Exception as e:                      # no P2P mapping on this box: say so once, use NCCL
            _DISABLED_REASON = f"{type(e).__name__}: {e}"
            if SP_STATE.rank == 0:
                print(f"vorta_b200.ulysses: peer-memory exchange unavailable ({_DISABLED_REASON}); using NCCL all-to-all",
                      flush=True)
            return None
        _EXCHANGES[key] = ex
    return ex


def disabled_reason() -> Optional[str]:
    return _DISABLED_REASON
