"""CPU oracle for VORTA's routed sparse attention path — TEST INFRASTRUCTURE ONLY.

A plain torch/numpy restatement of the reference algorithm (wenhao728/VORTA), one function per reference
function, each citing the reference file:line it follows.  Nothing under ``vorta_b200/`` imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may use it, and only as the checker / CPU baseline.

Pinning: the reference ships no tests or golden vectors.  This restatement is pinned by
``oracle/make_golden.py``, which imports the *reference's own modules* from /root/reference (with stub
``diffusers`` modules), runs them on seeded inputs and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against those files.

Numerics: floating-point work runs in whatever dtype the caller passes (fp32 for parity runs, fp64 for the
"true ranking" of near-tied similarities), exactly like the reference's dtype-generic torch code.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BRANCH_FULL, BRANCH_CORESET, BRANCH_SLIDING = 0, 1, 2   # expert order, vorta/attention/wan.py:352-354


# ----------------------------------------------------------------------------------------------------------
# coreset groups — vorta/attention/coreset_select.py:15-60
# ----------------------------------------------------------------------------------------------------------
@dataclass
class GroupInfo:
    center_indices: torch.Tensor      # (G, 1) int64
    margin_indices: torch.Tensor      # (G, g-1) int64
    num_unpooled_tokens_per_group: int


def get_group_info(latent_shape: Sequence[int], window: Sequence[int], reduction_rate: float = 0.5) -> GroupInfo:
    """Non-overlapping (f, h, w) windows in raster group order, raster member order; the centre is member
    (f//2, h//2, w//2) (coreset_select.py:51); n_unpooled = int(g * (1 - r)) - 1 (:54)."""
    T, H, W = (int(x) for x in latent_shape)
    fw, hw, ww = (int(x) for x in window)
    gf, gh, gw = T // fw, H // hw, W // ww
    # token id of member (a, b, c) of group (i, j, k): explicit index arithmetic with broadcasting
    i = np.arange(gf)[:, None, None, None, None, None]
    j = np.arange(gh)[None, :, None, None, None, None]
    k = np.arange(gw)[None, None, :, None, None, None]
    a = np.arange(fw)[None, None, None, :, None, None]
    b = np.arange(hw)[None, None, None, None, :, None]
    c = np.arange(ww)[None, None, None, None, None, :]
    tok = ((i * fw + a) * H + (j * hw + b)) * W + (k * ww + c)
    members = tok.reshape(gf * gh * gw, fw * hw * ww).astype(np.int64)
    slot = (fw // 2) * hw * ww + (hw // 2) * ww + ww // 2
    center = members[:, slot:slot + 1]
    margin = np.delete(members, slot, axis=1)
    n_unpooled = int(fw * hw * ww * (1 - reduction_rate)) - 1
    return GroupInfo(torch.from_numpy(center.copy()), torch.from_numpy(margin.copy()), n_unpooled)


# ----------------------------------------------------------------------------------------------------------
# similarity selection / pool / unpool — coreset_select.py:68-185
# ----------------------------------------------------------------------------------------------------------
def similarity(x: torch.Tensor, info: GroupInfo) -> torch.Tensor:
    """cos(centre, margin) per group (coreset_select.py:98-103): F.normalize (eps 1e-12) then a dot product."""
    c = F.normalize(x[:, :, info.center_indices[:, 0], :], p=2, dim=-1)          # (B, h, G, D)
    m = F.normalize(x[:, :, info.margin_indices, :], p=2, dim=-1)                # (B, h, G, g-1, D)
    return torch.einsum("bhgd,bhgmd->bhgm", c, m)


def match(x: torch.Tensor, info: GroupInfo) -> Tuple[torch.Tensor, torch.Tensor]:
    """Ascending argsort of the similarities; the first n_unpooled (least similar) margins are kept, the rest
    are dropped (coreset_select.py:105-113).  Returns (unpooled_argsort_sim, pooled_argsort_sim)."""
    order = similarity(x, info).argsort(dim=-1, descending=False)
    n_u = info.num_unpooled_tokens_per_group
    return order[..., :n_u], order[..., n_u:]


def pool(x: torch.Tensor, info: GroupInfo, matching: Tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
    """[all centres (group order) | kept margins (group-major, ascending similarity)] (coreset_select.py:118-123)."""
    un, _ = matching
    B, h, _, D = x.shape
    kept_tok = kept_token_ids(info, un)                                           # (B, h, G * n_u)
    centres = x[:, :, info.center_indices[:, 0], :]
    kept = torch.gather(x, 2, kept_tok[..., None].expand(-1, -1, -1, D))
    return torch.cat([centres, kept], dim=2)


def kept_token_ids(info: GroupInfo, unpooled_argsort: torch.Tensor) -> torch.Tensor:
    B, h = unpooled_argsort.shape[:2]
    mi = info.margin_indices[None, None].expand(B, h, -1, -1)
    return torch.gather(mi, 3, unpooled_argsort).flatten(2, 3)


def dropped_token_ids(info: GroupInfo, pooled_argsort: torch.Tensor) -> torch.Tensor:
    B, h = pooled_argsort.shape[:2]
    mi = info.margin_indices[None, None].expand(B, h, -1, -1)
    return torch.gather(mi, 3, pooled_argsort)                                    # (B, h, G, n_p)


def unpool(y: torch.Tensor, info: GroupInfo, matching: Tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
    """Kept tokens receive their own output, every dropped margin receives its centre's output
    (coreset_select.py:154-184)."""
    un, po = matching
    B, h, _, D = y.shape
    G = info.center_indices.shape[0]
    S = G * (1 + info.margin_indices.shape[1])
    out = torch.zeros(B, h, S, D, dtype=y.dtype)
    centres, kept = y[:, :, :G], y[:, :, G:]
    out[:, :, info.center_indices[:, 0]] = centres
    out.scatter_(2, kept_token_ids(info, un)[..., None].expand(-1, -1, -1, D), kept)
    drop_tok = dropped_token_ids(info, po)                                        # (B, h, G, n_p)
    n_p = drop_tok.shape[-1]
    if n_p > 0:
        src = centres[:, :, :, None, :].expand(-1, -1, -1, n_p, -1).reshape(B, h, G * n_p, D)
        out.scatter_(2, drop_tok.flatten(2, 3)[..., None].expand(-1, -1, -1, D), src)
    return out


# ----------------------------------------------------------------------------------------------------------
# tile-major layout — vorta/attention/tile.py:7-78 (sp_size == 1)
# ----------------------------------------------------------------------------------------------------------
def tile_permutation(latent_shape: Sequence[int], tile: Sequence[int]) -> torch.Tensor:
    """perm[p] = raster token at tile-major position p; tile_layout(x) == x[..., perm, :] (tile.py:26-29)."""
    T, H, W = (int(x) for x in latent_shape)
    tt, th, tw = (int(x) for x in tile)
    ids = torch.arange(T * H * W).reshape(T // tt, tt, H // th, th, W // tw, tw)
    return ids.permute(0, 2, 4, 1, 3, 5).reshape(-1)


# ----------------------------------------------------------------------------------------------------------
# sliding-tile mask — vorta/attention/sliding_attn_flex.py:72-134
# ----------------------------------------------------------------------------------------------------------
def _clamp_ref(x: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    return x.clamp(lo, hi)      # torch semantics: lo > hi -> hi everywhere (the reference relies on it)


def sliding_tile_mask(latent_shape, window, tile, text_len: int = 0, text_valid: int = 0) -> torch.Tensor:
    """Dense boolean (S+text, S+text) mask in TILE-MAJOR index space, as mask_mod defines it
    (sliding_attn_flex.py:101-129).  Small grids only."""
    T, H, W = (int(x) for x in latent_shape)
    nt = [T // tile[0], H // tile[1], W // tile[2]]
    tau = tile[0] * tile[1] * tile[2]
    S = T * H * W
    N, Nv = S + text_len, S + text_valid
    idx = torch.arange(N)
    tid = idx // tau
    coords = [tid // (nt[1] * nt[2]), (tid % (nt[1] * nt[2])) // nt[2], tid % nt[2]]
    q, kv = idx[:, None], idx[None, :]
    vid = torch.ones(N, N, dtype=torch.bool)
    for d in range(3):
        half = window[d] // 2
        centre = _clamp_ref(coords[d], half, nt[d] - 1 - half)
        vid &= (centre[:, None] - coords[d][None, :]).abs() <= half
    vid &= (q < S) & (kv < S)
    text_to_all = (q >= S) & (q < Nv) & (kv < Nv)
    video_to_text = (q < S) & (kv >= S) & (kv < Nv)
    return text_to_all | video_to_text | vid


def tile_windows(latent_shape, window, tile) -> np.ndarray:
    """(num_tiles, 6) lo/hi tile coordinates of each query tile's key window — the closed form of the mask."""
    nt = [latent_shape[d] // tile[d] for d in range(3)]
    out = np.zeros((nt[0] * nt[1] * nt[2], 6), dtype=np.int32)
    for a in range(nt[0]):
        for b in range(nt[1]):
            for c in range(nt[2]):
                t = (a * nt[1] + b) * nt[2] + c
                for d, q in enumerate((a, b, c)):
                    half = window[d] // 2
                    centre = int(_clamp_ref(torch.tensor(q), half, nt[d] - 1 - half))
                    ks = [k for k in range(nt[d]) if abs(centre - k) <= half]
                    out[t, d], out[t, 3 + d] = ks[0], ks[-1]
    return out


# ----------------------------------------------------------------------------------------------------------
# attention branches
# ----------------------------------------------------------------------------------------------------------
def sdpa(q, k, v, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, is_causal=False)


def full_attention(q, k, v, text_len: int = 0, text_valid: int = 0) -> torch.Tensor:
    """wan.py:142-144; hunyuan.py:169-176: attend over the first S + text_valid tokens, zero the padded tail."""
    N = q.shape[2]
    n_eff = N - (text_len - text_valid)
    out = sdpa(q[:, :, :n_eff], k[:, :, :n_eff], v[:, :, :n_eff])
    return F.pad(out, (0, 0, 0, N - n_eff), value=0.0)


def coreset_attention(q, k, v, info: GroupInfo, text_len: int = 0, text_valid: int = 0,
                      kv_from_k: bool = False, q_matching=None, k_matching=None) -> torch.Tensor:
    """wan.py:243-270 (one matching from Q reused for K and V) / hunyuan.py:410-457 (K and V use K's own matching;
    the pooled video tokens are followed by the text tokens; unpool with Q's matching)."""
    S = q.shape[2] - text_len
    qv, kv_, vv = q[:, :, :S], k[:, :, :S], v[:, :, :S]
    mq = q_matching if q_matching is not None else match(qv, info)
    mk = (k_matching if k_matching is not None else match(kv_, info)) if kv_from_k else mq
    pq, pk, pv = pool(qv, info, mq), pool(kv_, info, mk), pool(vv, info, mk)
    if text_len:
        pq = torch.cat([pq, q[:, :, S:]], 2)
        pk = torch.cat([pk, k[:, :, S:]], 2)
        pv = torch.cat([pv, v[:, :, S:]], 2)
    o = full_attention(pq, pk, pv, text_len, text_valid)
    S_c = pq.shape[2] - text_len
    video = unpool(o[:, :, :S_c], info, mq)
    return torch.cat([video, o[:, :, S_c:]], 2) if text_len else video


def sliding_tile_attention(q, k, v, latent_shape, window, tile, text_len: int = 0, text_valid: int = 0,
                           dense: Optional[bool] = None) -> torch.Tensor:
    """wan.py:272-294 + sliding_attn_flex.py:137-211.  Inputs / outputs in raster order.  ``dense`` evaluates the
    mask densely (small grids); otherwise every query tile attends to the gathered keys of its window, which is
    the same function."""
    S = q.shape[2] - text_len
    perm = tile_permutation(latent_shape, tile)
    if dense is None:
        dense = (S + text_len) <= 4096
    if dense:
        order = torch.cat([perm, torch.arange(S, S + text_len)])
        qt, kt, vt = q[:, :, order], k[:, :, order], v[:, :, order]
        mask = sliding_tile_mask(latent_shape, window, tile, text_len, text_valid)
        dead = ~mask.any(dim=1)                       # padded text queries: fully masked rows -> 0
        mask = mask.clone()
        mask[dead, 0] = True
        ot = sdpa(qt, kt, vt, mask)
        ot[:, :, dead] = 0
        out = torch.empty_like(ot)
        out[:, :, order] = ot
        return out
    win = tile_windows(latent_shape, window, tile)
    nt = [latent_shape[d] // tile[d] for d in range(3)]
    tau = tile[0] * tile[1] * tile[2]
    tiles = perm.reshape(-1, tau)
    text_keys = torch.arange(S, S + text_valid)
    out = torch.zeros_like(q)
    for t in range(tiles.shape[0]):
        lo, hi = win[t, :3], win[t, 3:]
        ids = [(a * nt[1] + b) * nt[2] + c for a in range(lo[0], hi[0] + 1) for b in range(lo[1], hi[1] + 1)
               for c in range(lo[2], hi[2] + 1)]
        keys = torch.cat([tiles[ids].reshape(-1), text_keys])
        out[:, :, tiles[t]] = sdpa(q[:, :, tiles[t]], k[:, :, keys], v[:, :, keys])
    if text_valid:
        nv = S + text_valid
        out[:, :, S:nv] = sdpa(q[:, :, S:nv], k[:, :, :nv], v[:, :, :nv])
    return out


# ----------------------------------------------------------------------------------------------------------
# router and routing — vorta/patch/router.py:33-43; wan.py:396-400
# ----------------------------------------------------------------------------------------------------------
def router_forward(temb: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, heads: int,
                   module_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """router.py:41-43.  ``module_dtype=torch.bfloat16`` restates what torch does when the router runs in bf16 (the
    reference's inference default): every module computes in fp32 and rounds its OUTPUT to bf16 — SiLU, Linear (one
    rounding after the fp32 accumulate + bias), Softmax."""
    def rnd(t):
        return t if module_dtype is None else t.to(module_dtype).to(torch.float32)
    logits = rnd(F.linear(rnd(F.silu(temb.float())), weight.float(), bias.float()))
    return rnd(torch.softmax(logits.unflatten(1, (heads, 3)), dim=-1))


def route_top1(routing_score: torch.Tensor, tau: Optional[float]) -> torch.Tensor:
    """Per-head branch id from the FIRST sample of the batch; below-threshold heads fall back to full attention."""
    score, idx = routing_score[0].topk(1, dim=-1)
    idx = idx.clone()
    if tau is not None:
        idx[score < tau] = 0
    return idx.squeeze(-1)


def routed_attention(q, k, v, info: GroupInfo, latent_shape, window, tile, branch: Optional[torch.Tensor] = None,
                     weights: Optional[torch.Tensor] = None, text_len: int = 0, text_valid: int = 0,
                     kv_from_k: bool = False) -> torch.Tensor:
    """Eval semantics (``branch``: wan.py:351-383, each head runs its branch) or Train semantics (``weights``
    (B, H, 3): wan.py:227-239 + :296-300, out = sum_e w_e * O_e)."""
    def run(e: int, hs):
        qq, kk, vv = q[:, hs], k[:, hs], v[:, hs]
        if e == BRANCH_FULL:
            return full_attention(qq, kk, vv, text_len, text_valid)
        if e == BRANCH_CORESET:
            return coreset_attention(qq, kk, vv, info, text_len, text_valid, kv_from_k)
        return sliding_tile_attention(qq, kk, vv, latent_shape, window, tile, text_len, text_valid)

    if weights is not None:
        H = q.shape[1]
        outs = [run(e, list(range(H))) for e in range(3)]
        return (weights[:, :, :, None, None].to(q.dtype) * torch.stack(outs, dim=2)).sum(dim=2)
    out = torch.zeros_like(q)
    for e in range(3):
        hs = [h for h in range(q.shape[1]) if int(branch[h]) == e]
        if hs:
            out[:, hs] = run(e, hs)
    return out


# ----------------------------------------------------------------------------------------------------------
# Ulysses all-to-all as a pure permutation — vorta/ulysses/utils.py:15-93 (verified on gloo ranks)
# ----------------------------------------------------------------------------------------------------------
def ulysses_scatter_heads(shards: List[torch.Tensor]) -> List[torch.Tensor]:
    """scatter_idx=1, gather_idx=2: rank r holds (B, H, S/P, D) -> rank r gets head chunk r of the full
    sequence, (B, H/P, S, D), sequence concatenated in rank order."""
    P = len(shards)
    full = torch.cat(shards, dim=2)
    hp = full.shape[1] // P
    return [full[:, r * hp:(r + 1) * hp].contiguous() for r in range(P)]


def ulysses_gather_heads(shards: List[torch.Tensor]) -> List[torch.Tensor]:
    """scatter_idx=2, gather_idx=1: the exact inverse."""
    P = len(shards)
    full = torch.cat(shards, dim=1)
    sp = full.shape[2] // P
    return [full[:, :, r * sp:(r + 1) * sp].contiguous() for r in range(P)]


# ----------------------------------------------------------------------------------------------------------
# algorithmic work — BASELINE.md section 3
# ----------------------------------------------------------------------------------------------------------
def branch_flops(branch: int, latent_shape, window, tile, lowres_window, reduction_rate=0.5, text_valid=0,
                 head_dim=128) -> float:
    S = int(np.prod(latent_shape))
    if branch == BRANCH_FULL:
        return 4.0 * (S + text_valid) ** 2 * head_dim
    if branch == BRANCH_CORESET:
        g = int(np.prod(lowres_window))
        s_c = (S // g) * (1 + int(g * (1 - reduction_rate)) - 1)
        return 4.0 * (s_c + text_valid) ** 2 * head_dim
    nt = [latent_shape[d] // tile[d] for d in range(3)]
    kw = int(np.prod(tile))
    for d in range(3):
        kw *= min(nt[d], 2 * (window[d] // 2) + 1)
    return 4.0 * head_dim * (S * (kw + text_valid) + text_valid * (S + text_valid))
