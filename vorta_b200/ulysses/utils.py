"""Ulysses head <-> sequence exchange — interface of the reference's ``vorta/ulysses/utils.py``
(_all_to_all_4D :15-93, all_to_all_4D :123-124, all_gather :127-162, shrink_dim :218-223).

The exchange is one equal-split NCCL all-to-all over NVLink / NVSwitch issued through ``torch.distributed``;
the layout passes on either side are the library's pack / unpack kernels (one pass each instead of the
reference's two transposed copies, and no device-wide synchronize after the collective).

``exchange_qkv`` / ``exchange_out`` are the forms the processors use: Q, K and V travel in ONE collective, the
receive buffer is consumed by the attention kernel through strides, and the attention output is written
directly in send order, so a layer costs two collectives and two layout passes in total.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .. import _lib as L
from .parallel_states import SP_STATE


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _a2a(recv: torch.Tensor, send: torch.Tensor) -> None:
    """Equal-split all-to-all over the SP group (works on NCCL and, for host-logic tests, gloo)."""
    dist.all_to_all_single(recv, send, group=SP_STATE.group)


def _head_table(head_at: Optional[Sequence[int]], heads: int):
    """ctypes int32[H] of a slot -> head table (None = the reference's contiguous chunks)."""
    if head_at is None:
        return None
    if len(head_at) != heads:
        raise ValueError(f"head_at has {len(head_at)} entries for {heads} heads")
    return (C.c_int32 * heads)(*[int(h) for h in head_at])


def pack_heads(x: torch.Tensor, world: int, head_at: Optional[Sequence[int]] = None) -> torch.Tensor:
    """x: (n_tensors, S_loc, H, 128) bf16 contiguous -> (n_tensors, P, S_loc, H/P, 128)."""
    n, s_loc, H, D = x.shape
    send = torch.empty((n, world, s_loc, H // world, D), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.lib().vb_ulysses_pack_heads(x.data_ptr(), send.data_ptr(), s_loc, H, world, n, x.stride(0),
                                              send.stride(0), _head_table(head_at, H), _stream(x.device)))
    return send


def unpack_heads(recv: torch.Tensor, head_at: Optional[Sequence[int]] = None) -> torch.Tensor:
    """recv: (P, S_loc, H/P, 128) -> (S_loc, H, 128)."""
    world, s_loc, hp, D = recv.shape
    y = torch.empty((s_loc, hp * world, D), dtype=recv.dtype, device=recv.device)
    with torch.cuda.device(recv.device):
        L.check(L.lib().vb_ulysses_unpack_heads(recv.data_ptr(), y.data_ptr(), s_loc, hp * world, world,
                                                _head_table(head_at, hp * world), _stream(recv.device)))
    return y


def exchange_qkv(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, extra_rows: int = 0,
                 head_at: Optional[Sequence[int]] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """q, k, v: (B, H, S_loc, 128) views: this rank's token shard, all heads.
    Each tensor may have its own (batch, token, head) strides.
    Returns (B, H/P, S + extra_rows, 128) views of token-major memory: this rank's head chunk over the full
    sequence; ``extra_rows`` uninitialised rows are left at the end for the caller (HunyuanVideo text tokens).
    ``head_at``: slot -> head table of a balanced placement (``balance.balance_heads``); rank r then holds heads
    ``head_at[r*H/P:(r+1)*H/P]`` instead of the contiguous chunk.
    Any batch size, like the reference's ``_all_to_all_4D`` (utils.py:15-93): samples travel one after the other."""
    P = SP_STATE.sp_size
    B, H, s_loc, D = q.shape
    if B != 1:
        parts = [exchange_qkv(q[b:b + 1], k[b:b + 1], v[b:b + 1], extra_rows, head_at) for b in range(B)]
        # (B, S + extra, hp, D) memory, handed back as the same (B, hp, S + extra, D) view a single sample gets
        return tuple(torch.cat([p[i].transpose(1, 2) for p in parts], dim=0).transpose(1, 2) for i in range(3))
    if H % P != 0:
        raise ValueError(f"heads {H} must be divisible by the sequence-parallel size {P}")
    for t in (q, k, v):
        if t.stride(3) != 1:
            raise ValueError("q, k, v must have contiguous channels")
    hp = H // P
    send = torch.empty((3, P, s_loc, hp, D), dtype=q.dtype, device=q.device)
    with torch.cuda.device(q.device):
        i64x3 = C.c_int64 * 3
        L.check(L.lib().vb_ulysses_pack_qkv(q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                            i64x3(q.stride(2), k.stride(2), v.stride(2)),
                                            i64x3(q.stride(1), k.stride(1), v.stride(1)),
                                            send.data_ptr(), s_loc, H, P, _head_table(head_at, H),
                                            _stream(q.device)))
    out = []
    for i in range(3):
        # chunk p of the receive buffer = tokens of rank p for my head chunk: (S, hp, D) token-major, in place
        recv = torch.empty((P * s_loc + extra_rows, hp, D), dtype=q.dtype, device=q.device)
        _a2a(recv[:P * s_loc], send[i])
        out.append(recv.unsqueeze(0).transpose(1, 2))                 # (1, hp, S + extra, D) view
    return tuple(out)


def exchange_out(o: torch.Tensor, head_at: Optional[Sequence[int]] = None) -> torch.Tensor:
    """o: (B, H/P, S, 128) view of (B, S, H/P, 128) memory -> (B, H, S_loc, 128) view of (B, S_loc, H, 128).
    ``head_at`` must be the table the matching ``exchange_qkv`` used; heads come back in their original order."""
    P = SP_STATE.sp_size
    B, hp, S, D = o.shape
    if B != 1:
        parts = [exchange_out(o[b:b + 1], head_at).transpose(1, 2) for b in range(B)]
        return torch.cat(parts, dim=0).transpose(1, 2)
    s_loc = S // P
    send = o.transpose(1, 2).reshape(P, s_loc, hp, D)
    if not send.is_contiguous():
        send = send.contiguous()
    recv = torch.empty_like(send)
    _a2a(recv, send)
    y = unpack_heads(recv, head_at)                                   # (S_loc, H, D)
    return y.unsqueeze(0).transpose(1, 2)


def all_to_all_4D(input_: torch.Tensor, scatter_idx: int, gather_idx: int) -> torch.Tensor:
    """Reference signature (utils.py:123): (B, H, S/P, D) <-> (B, H/P, S, D) for one tensor."""
    if not SP_STATE.enabled:
        return input_
    P = SP_STATE.sp_size
    if scatter_idx == 1 and gather_idx == 2:
        B, H, s_loc, D = input_.shape
        if B != 1:       # utils.py:15-93 handles any batch: one sample after the other
            parts = [all_to_all_4D(input_[b:b + 1], 1, 2).transpose(1, 2) for b in range(B)]
            return torch.cat(parts, dim=0).transpose(1, 2)
        x = input_.transpose(1, 2).reshape(1, s_loc, H, D)
        send = pack_heads(x.contiguous(), P)[0]                       # (P, S_loc, hp, D)
        recv = torch.empty_like(send)
        _a2a(recv, send)
        return recv.reshape(P * s_loc, H // P, D).unsqueeze(0).transpose(1, 2)
    if scatter_idx == 2 and gather_idx == 1:
        return exchange_out(input_)
    raise RuntimeError("scatter_idx must be 1 or 2 and gather_idx must be 1 or 2")       # utils.py:93


def all_gather(input_: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """utils.py:127-162 (forward): concatenate every rank's tensor along ``dim`` in rank order."""
    if not SP_STATE.enabled:
        return input_
    parts = [torch.empty_like(input_) for _ in range(SP_STATE.sp_size)]
    dist.all_gather(parts, input_.contiguous(), group=SP_STATE.group)
    return torch.cat(parts, dim=dim)


def shrink_dim(tensor: torch.Tensor, dim: int) -> torch.Tensor:
    """utils.py:218-223: this rank's contiguous slice along ``dim`` (no communication)."""
    if SP_STATE.enabled:
        local = tensor.size(dim) // SP_STATE.sp_size
        return tensor.narrow(dim, local * SP_STATE.group_local_rank, local)
    return tensor


def local_heads(values, heads: int, head_at: Optional[Sequence[int]] = None):
    """Slice a per-head list (branch ids, weights rows) down to the heads this rank holds."""
    if not SP_STATE.enabled:
        return values
    hp = heads // SP_STATE.sp_size
    r = SP_STATE.group_local_rank
    if head_at is not None:
        return [values[h] for h in head_at[r * hp:(r + 1) * hp]]
    return values[r * hp:(r + 1) * hp]
