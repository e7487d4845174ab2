"""Ulysses exchange over NVLink / NVSwitch peer memory, fused into the kernels on either side of the attention.

The reference (vorta/ulysses/utils.py:15-93) and the NCCL path of this package move Q, K, V and O with all-to-all
collectives plus layout copies.  Here every rank maps every peer's receive buffers (torch symmetric memory, one
allocation per geometry) and

  * "in"  : ``vb_ulysses_scatter_qkv`` stores each (token, head) row of Q, K, V straight into the buffer of the rank that
            owns the head — one pass, no staging buffer, no collective;
  * "out" : the attention kernel's epilogue stores each output row straight into the buffer of the rank that owns the
            token (``vb_attn_args.out_peer_ptrs``), already in the (S_loc, H, 128) layout the output projection reads.

Two device-side barriers per layer order the remote stores against their consumers (see ``PeerExchange``).  Used for
the top-1 (Eval) processors at batch size 1, Wan and HunyuanVideo (whose replicated text rows are filled locally on the
way in and stored to every rank on the way out); the blended (Train) processors and larger batches take the NCCL path
in ``utils.py``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .. import _lib as L
from . import balance
from .parallel_states import SP_STATE

_SIDE_CTAS = int(os.environ.get("VB_ULYSSES_SIDE_CTAS", "32"))      # grid cap of the side-stream store kernels
_EXCHANGES: Dict[tuple, "PeerExchange"] = {}
_DISABLED_REASON: Optional[str] = None


class PeerExchange:
    """Symmetric receive buffers of one geometry (heads H, local tokens S_loc, world P), all 128-channel bf16:
    ``qkv`` (3, S + T, H/P, 128): this rank's head chunk over the full sequence, video rows written by every peer,
                                  the T replicated text rows (HunyuanVideo; T = 0 for Wan) filled locally;
    ``out`` (S_loc + T, H, 128): this rank's token shard over all heads, written by every peer's attention epilogue;
                                  text rows of every head are stored to every rank (the head all-gather of
                                  vorta/attention/hunyuan.py:186-187).

    Ordering per layer (all on the caller's stream):
        scatter_qkv -> barrier A -> attention (reads qkv, stores into peers' out) -> barrier B -> consumer reads out.
    A rank leaves barrier B only after every rank finished its attention, so the next layer's scatter cannot overwrite
    a qkv buffer that is still being read; it reaches the next barrier A only after its own consumer of ``out`` was
    issued, so nobody stores into ``out`` while it is still needed."""

    def __init__(self, heads: int, s_loc: int, device: torch.device, text_len: int = 0):
        import torch.distributed._symmetric_memory as symm
        P, rank = SP_STATE.sp_size, SP_STATE.group_local_rank
        self.P, self.rank, self.heads, self.s_loc, self.hp = P, rank, heads, s_loc, heads // P
        self.S = s_loc * P
        self.text_len = int(text_len)
        # head slots of the receive buffer: more than H / P, so that a rank can hold an uneven number of heads and
        # query halves of full heads (balance.place_units)
        self.slots = balance.max_slots(heads, P)
        group = SP_STATE.group if SP_STATE.group is not None else dist.group.WORLD
        self.qkv = symm.empty((3, self.S + self.text_len, self.slots, 128), dtype=torch.bfloat16, device=device)
        self.out = symm.empty((s_loc + self.text_len, heads, 128), dtype=torch.bfloat16, device=device)
        self.h_qkv = symm.rendezvous(self.qkv, group)
        self.h_out = symm.rendezvous(self.out, group)
        self.qkv_ptrs = (C.c_void_p * P)(*[int(p) for p in self.h_qkv.buffer_ptrs])
        self.out_ptrs = [int(p) for p in self.h_out.buffer_ptrs]
        # bytes this rank stores into OTHER ranks' buffers over NVLink, counted from the placement tables (the NVLink data
        # counters of nvidia-smi read N/A on this pool): "in" = Q/K/V rows of scatter_qkv, "out" = attention output rows
        self.nvlink_tx_bytes = 0
        self._side = None

    def _tables(self, placement):
        peers, slots, heads = [], [], []
        for p, units in enumerate(placement):
            if len(units) > self.slots:
                raise ValueError(f"rank {p} was given {len(units)} units for {self.slots} slots")
            for slot, (h, _part) in enumerate(units):
                peers.append(p); slots.append(slot); heads.append(int(h))
        n = len(peers)
        remote_in = sum(1 for p in peers if p != self.rank)
        frac_out = sum(1.0 if part == balance.WHOLE else 1.0 / balance.part_kn(part)[1] for _, part in placement[self.rank])
        self.nvlink_tx_bytes += 3 * self.s_loc * remote_in * 256 + int(frac_out * (self.S - self.s_loc) * 256)
        arr = C.c_int32 * n
        return arr(*peers), arr(*slots), arr(*heads), n

    def _scatter(self, tensors, tables, mask: int, max_ctas: int = 0) -> None:
        """tensors: [q, k, v] (1, H, S_loc, 128) views or None for the ones ``mask`` does not select."""
        i64x3 = C.c_int64 * 3
        ref = next(t for t in tensors if t is not None)
        ptr = [t.data_ptr() if t is not None else None for t in tensors]
        st_s = i64x3(*[t.stride(2) if t is not None else 8 for t in tensors])
        st_h = i64x3(*[t.stride(1) if t is not None else 8 for t in tensors])
        peers, slots, heads, n = tables
        with torch.cuda.device(ref.device):
            L.check(L.lib().vb_ulysses_scatter_slots_partial(
                ptr[0], ptr[1], ptr[2], st_s, st_h, self.qkv_ptrs, self.S + self.text_len, self.s_loc, self.slots, self.P,
                self.rank, peers, slots, heads, n, mask, max_ctas, torch.cuda.current_stream(ref.device).cuda_stream))

    def scatter_qkv(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                    placement: Sequence[Sequence[Tuple[int, int]]], text: Optional[Sequence[torch.Tensor]] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """q, k, v: (1, H, S_loc, 128) views of this rank's token shard.  ``placement[p]`` = the (head, part) units of
        rank p in slot order (``balance.place_units`` or ``contiguous_placement``); every head of a unit is stored into
        that slot of rank p's buffer — a head split into query parts goes to several ranks.  Returns (1, slots, S + T, 128)
        views of the local receive buffer, valid after barrier A (issued here); only the first len(placement[rank])
        slots hold data.  ``text``: the replicated (1, H, T, 128) text rows of q, k, v; the rows of this rank's heads are
        copied behind the video rows locally (no traffic)."""
        self._scatter([q, k, v], self._tables(placement), 7)
        return self._finish_scatter(placement, text, q.device)

    def _finish_scatter(self, placement, text, device):
        if self.text_len:
            mine = [h for h, _ in placement[self.rank]]
            sel = torch.tensor(mine, device=device)
            for i, t in enumerate(text):          # (1, H, T, 128) -> rows [S, S + T) of my first len(mine) slots
                self.qkv[i, self.S:, :len(mine)].copy_(t[0].index_select(0, sel).transpose(0, 1))
        self.h_qkv.barrier(channel=0)
        return tuple(self.qkv[i].unsqueeze(0).transpose(1, 2) for i in range(3))

    # ---- overlapped form: K and V leave on a side stream while the next projection GEMM runs -----------------------
    def overlap_begin(self, placement):
        """Start an exchange whose tensors are handed over one by one (``overlap_send``) as the projections produce
        them; ``overlap_end`` issues barrier A.  The side stream has high priority and its store kernels use a capped
        grid, so they share the GPU with the GEMM of the next projection instead of running exposed after all three."""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.qkv.device, priority=-1)
        self._tables_now = self._tables(placement)
        self._placement_now = placement
        self._pending = False

    def overlap_send(self, which: int, x: torch.Tensor, side: bool) -> None:
        """x: (1, H, S_loc, 128) view of tensor ``which`` (0 = q, 1 = k, 2 = v), complete on the current stream."""
        tensors = [None, None, None]
        tensors[which] = x
        if not side:
            self._scatter(tensors, self._tables_now, 1 << which)
            return
        main = torch.cuda.current_stream(x.device)
        ready = torch.cuda.Event()
        ready.record(main)
        self._side.wait_event(ready)
        x.record_stream(self._side)
        with torch.cuda.stream(self._side):
            self._scatter(tensors, self._tables_now, 1 << which, max_ctas=_SIDE_CTAS)
        self._pending = True

    def overlap_end(self, device, text=None):
        if self._pending:
            torch.cuda.current_stream(device).wait_stream(self._side)
        return self._finish_scatter(self._placement_now, text, device)

    def zero_padded_text(self, text_valid: int) -> None:
        """Rows of padded text queries are written by nobody (hunyuan.py:176 pads with zeros): clear them locally."""
        if self.text_len > text_valid:
            self.out[self.s_loc + text_valid:].zero_()

    def finish_out(self) -> torch.Tensor:
        """Barrier B, then this rank's (1, H, S_loc + T, 128) view of the gathered outputs."""
        self.h_out.barrier(channel=1)
        return self.out.unsqueeze(0).transpose(1, 2)


def get_exchange(heads: int, s_loc: int, device: torch.device, text_len: int = 0) -> Optional[PeerExchange]:
    """The cached exchange for this geometry, or None when peer memory is unavailable / disabled
    (VB_ULYSSES=nccl), in which case the caller takes the NCCL path."""
    global _DISABLED_REASON
    if not SP_STATE.enabled or os.environ.get("VB_ULYSSES", "peer") == "nccl" or _DISABLED_REASON is not None:
        return None
    if SP_STATE.sp_size > 8 or heads % SP_STATE.sp_size != 0:
        return None
    key = (heads, s_loc, str(device), SP_STATE.sp_size, int(text_len))
    ex = _EXCHANGES.get(key)
    if ex is None:
        try:
            ex = PeerExchange(heads, s_loc, device, text_len)
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


def contiguous_placement(heads: int, world: int):
    """The reference's chunking (vorta/ulysses/utils.py:60-66) as a placement table: rank r holds whole heads
    [r * H / P, (r + 1) * H / P)."""
    hp = heads // world
    return [[(h, balance.WHOLE) for h in range(r * hp, (r + 1) * hp)] for r in range(world)]


def layer_placement(branch: Sequence[int], plan, heads: int, world: int, slots: int):
    """Placement of one layer's heads for the peer exchange: cost-balanced units (whole heads and, at P >= 4, query
    halves of full heads) when balancing is on, the contiguous chunks otherwise."""
    if branch is None or not balance.enabled():
        return contiguous_placement(heads, world)
    return balance.place_units(list(branch), balance.branch_costs(plan), world, slots,
                               allow_split=balance.split_enabled(world))


def local_units(placement, rank: int, branch: Sequence[int], slots: int):
    """(branch id per slot, output head per slot) of this rank for ``vb_attn_fwd``: unused slots are skipped, the parts
    of a split full head become VB_BRANCH_FULL_PART(k, n)."""
    ids, out_heads = [], []
    for h, part in placement[rank]:
        e = int(branch[h])
        if part != balance.WHOLE:
            if e != L.BRANCH_FULL:
                raise ValueError("only full-attention heads are split into query parts")
            e = 16 + part                       # VB_BRANCH_FULL_PART(k, n)
        ids.append(e)
        out_heads.append(int(h))
    pad = slots - len(ids)
    return ids + [L.BRANCH_SKIP] * pad, out_heads + [0] * pad


def nvlink_tx_bytes(reset: bool = False) -> int:
    """Bytes this rank has stored into other ranks' exchange buffers since the last reset (all geometries)."""
    total = sum(ex.nvlink_tx_bytes for ex in _EXCHANGES.values())
    if reset:
        for ex in _EXCHANGES.values():
            ex.nvlink_tx_bytes = 0
    return total


def overlap_enabled() -> bool:
    """VB_ULYSSES_OVERLAP=1 (opt-in) sends K and V on a side stream while the next projection GEMM runs.  Measured
    slower than one pass after all three projections — Wan-14B on 2 GPUs 1688 vs 1620 ms / step, alternating runs on one
    box (profiles/r2p_*): the cuBLAS GEMMs are persistent over all SMs, so the 32 store CTAs delay whole GEMM waves
    instead of filling idle SMs, and a 32-CTA grid does not saturate NVLink.  Results are bit-identical either way."""
    return os.environ.get("VB_ULYSSES_OVERLAP", "0") == "1"
