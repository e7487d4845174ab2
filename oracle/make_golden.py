"""Generate tests/golden/*.pt by running the REFERENCE's own code (imported from /root/reference) on seeded inputs.

Run in the build container only:   CXX=/usr/bin/g++ python oracle/make_golden.py
The reference has no tests or golden vectors of its own; these files are the pin for
``oracle/vorta_oracle.py`` and the fixtures the GPU parity tests compare against.  Everything is small
(a few hundred tokens) so the whole set stays well under a few MB.
"""
from __future__ import annotations

import os
import sys

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import fixtures as FX  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
D = 128


def seeded(shape, seed, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(dtype)


def save(name, obj):
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


# ----------------------------------------------------------------------------------------------------------
def golden_group_info(ref):
    cases = [((4, 6, 4), (2, 3, 2), 0.5), ((6, 6, 8), (3, 3, 2), 0.5), ((4, 6, 8), (1, 3, 4), 0.5),
             ((20, 30, 52), (2, 3, 2), 0.5), ((21, 30, 52), (3, 3, 2), 0.5), ((4, 8, 12), (2, 2, 2), 0.5),
             ((4, 6, 8), (2, 3, 2), 0.75)]
    out = []
    for lat, win, r in cases:
        info = ref.cs.get_group_info(lat, win, reduction_rate=r)
        big = info.center_indices.numel() > 2000
        out.append(dict(latent=lat, window=win, rate=r, n_unpooled=info.num_unpooled_tokens_per_group,
                        center=None if big else info.center_indices.clone(),
                        margin=None if big else info.margin_indices.clone(),
                        center_sum=int(info.center_indices.sum()), margin_sum=int(info.margin_indices.sum()),
                        center_head=info.center_indices[:16].clone(), margin_tail=info.margin_indices[-4:].clone(),
                        shape=(tuple(info.center_indices.shape), tuple(info.margin_indices.shape))))
    save("group_info.pt", out)


def golden_coreset(ref):
    """pool / unpool / matching in fp32 and fp64 on bf16-valued inputs (the contract dtype)."""
    out = []
    for lat, win, r, h, seed in [((4, 6, 8), (2, 3, 2), 0.5, 2, 11), ((6, 6, 8), (3, 3, 2), 0.5, 2, 12),
                                 ((4, 6, 8), (2, 3, 2), 0.75, 1, 14)]:
        S = lat[0] * lat[1] * lat[2]
        info = ref.cs.get_group_info(lat, win, reduction_rate=r)
        x = seeded((1, h, S, D), seed)
        rec = dict(latent=lat, window=win, rate=r, x=x)
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            pooled, m = ref.cs.pool_sequence_by_similarity(x.to(dt), info)
            rec[f"unpooled_{tag}"] = m.unpooled_argsort_sim.clone()
            rec[f"pooled_{tag}"] = m.pooled_argsort_sim.clone()
            if tag == "f32":
                rec["pooled_seq"] = pooled.to(torch.bfloat16)
                y = seeded(tuple(pooled.shape), seed + 100)
                rec["y"] = y
                rec["unpooled_seq"] = ref.cs.unpool_sequence_by_similarity(y.float(), info, m).to(torch.bfloat16)
                # reuse Q's matching for another tensor (wan.py:252-255)
                k = seeded((1, h, S, D), seed + 200)
                rec["k"] = k
                rec["pooled_k_with_q_matching"] = ref.cs.pool_sequence_by_similarity(k.float(), info, m)[0].to(
                    torch.bfloat16)
        out.append(rec)
    save("coreset.pt", out)


FULLSIZE_CORESET_CASES = [
    # (name, latent, coreset window, rate, heads, seed, heads per reference call)
    ("wan13_480p_81f", (21, 30, 52), (3, 3, 2), 0.5, 12, 101, 4),       # BASELINE configs[0]/[1]
    ("wan14_720p_81f", (21, 45, 80), (3, 3, 2), 0.5, 16, 102, 4),       # BASELINE configs[2] geometry, 16 of 40 heads
    ("wan14_720p_77f_native", (20, 45, 80), (2, 3, 2), 0.5, 4, 103, 4),  # the reference's own training geometry
    # spatially smooth activations (neighbouring tokens nearly parallel, cosines 0.99+): the regime where fp32 and
    # fp64 rankings of the reference can differ; `smooth` = weight of the per-token noise on a shared direction
    ("wan14_720p_81f_smooth", (21, 45, 80), (3, 3, 2), 0.5, 8, 104, 4, 0.02),
]


def fullsize_coreset_input(lat, heads, seed, smooth=None):
    """The input of a full-size coreset case, regenerated from its seed by the tests ((1, H, S, 128) bf16)."""
    S = lat[0] * lat[1] * lat[2]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((1, heads, S, D), generator=g)
    if smooth is not None:
        x = torch.randn((1, heads, 1, D), generator=g) + smooth * x
    return x.to(torch.bfloat16)


def golden_coreset_fullsize(ref):
    """BASELINE-size index tables of the REFERENCE's pool_sequence_by_similarity (coreset_select.py:68-124), run in
    fp32 (north_star's contract dtype) AND in fp64, on bf16-valued random inputs.  Stored: the fp64 tables (uint8)
    and every (head, group) row where the fp32 run differs from the fp64 run, with the fp64 cosine gap between the
    margins whose order flipped — the evidence for DESIGN.md section 3.3 (the kernel returns the fp64 ranking; the
    fp32 reference may legitimately differ from it only where two similarities are closer than fp32 can resolve)."""
    import torch.nn.functional as F
    out = []
    for name, lat, win, r, heads, seed, chunk, *rest in FULLSIZE_CORESET_CASES:
        smooth = rest[0] if rest else None
        info = ref.cs.get_group_info(lat, win, reduction_rate=r)
        x = fullsize_coreset_input(lat, heads, seed, smooth)
        tabs = {}
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            un, po = [], []
            for h0 in range(0, heads, chunk):
                _, m = ref.cs.pool_sequence_by_similarity(x[:, h0:h0 + chunk].to(dt), info)
                un.append(m.unpooled_argsort_sim.to(torch.uint8))
                po.append(m.pooled_argsort_sim.to(torch.uint8))
            tabs[tag] = (torch.cat(un, dim=1), torch.cat(po, dim=1))
        full32 = torch.cat(tabs["f32"], dim=-1)[0]          # (H, G, g-1): the complete ascending-similarity order
        full64 = torch.cat(tabs["f64"], dim=-1)[0]
        diff = (full32 != full64).any(dim=-1).nonzero()     # (n, 2): head, group
        rows = []
        for h, grp in diff.tolist():
            c = x[0, h, info.center_indices[grp, 0]].double()
            m = x[0, h, info.margin_indices[grp]].double()
            sim = F.normalize(m, dim=-1) @ F.normalize(c, dim=-1)
            a, b = full32[h, grp].long(), full64[h, grp].long()
            pos = (a != b).nonzero().flatten()
            # margins whose rank differs between the two runs; the largest fp64 gap among them bounds what fp32 lost
            gap = (sim[a[pos]] - sim[b[pos]]).abs().max().item()
            rows.append(dict(order_f32=full32[h, grp].clone(), gap_f64=gap))
        n_groups = heads * full64.shape[1]
        print(f"{name}: {len(rows)} of {n_groups} (head, group) rows differ between the reference in fp32 and fp64; "
              f"largest fp64 gap {max([r_['gap_f64'] for r_ in rows], default=0.0):.2e}")
        g1 = full64.shape[-1]
        out.append(dict(name=name, latent=lat, window=win, rate=r, heads=heads, seed=seed, smooth=smooth,
                        n_unpooled=info.num_unpooled_tokens_per_group,
                        unpooled_f64=tabs["f64"][0], pooled_f64=tabs["f64"][1], n_groups=n_groups,
                        # rows where the reference's fp32 run differs from its fp64 run: (head, group), the fp32
                        # order of the g-1 margins, and the fp64 cosine gap between the margins that swapped
                        f32_differs_at=diff.to(torch.int32),
                        f32_order=torch.stack([r_["order_f32"] for r_ in rows]) if rows else torch.zeros((0, g1), dtype=torch.uint8),
                        f32_gap=torch.tensor([r_["gap_f64"] for r_ in rows], dtype=torch.float64)))
    save("coreset_fullsize.pt", out)


def golden_tile_and_mask(ref):
    import torch.nn.attention.flex_attention as fa
    orig = ref.saf.create_block_mask
    ref.saf.create_block_mask = lambda *a, **k: orig(*a, **{**k, "_compile": False})   # dense evaluation, no Inductor
    out = []
    try:
        for lat, win, tile, tl, tv in [((6, 12, 12), (1, 3, 3), (2, 4, 4), 0, 0), ((8, 8, 12), (3, 3, 3), (2, 2, 4), 0, 0),
                                       ((4, 8, 8), (3, 3, 3), (2, 4, 4), 5, 3), ((10, 9, 8), (3, 3, 3), (5, 9, 8), 0, 0),
                                       ((4, 8, 12), (3, 3, 3), (2, 4, 4), 0, 0), ((4, 8, 12), (3, 5, 3), (1, 2, 3), 16, 11),
                                       ((4, 6, 8), (3, 3, 3), (2, 3, 4), 0, 0), ((4, 6, 8), (2, 3, 1), (1, 3, 2), 8, 8)]:
            S = lat[0] * lat[1] * lat[2]
            bm = ref.saf.create_sliding_tile_attn_mask_func(lat, win, tile, tl, tv, torch.device("cpu"))
            idx = torch.arange(S + tl)
            z = torch.zeros((), dtype=torch.int64)
            dense = bm.mask_mod(z, z, idx[:, None], idx[None, :])
            x = torch.arange(S, dtype=torch.float32).reshape(1, 1, S, 1)
            tiled = ref.tile.tile_layout(x, 1, tile, lat, head_dim=1).reshape(-1).long()
            back = ref.tile.untile_layout(tiled.reshape(1, 1, S, 1).float(), 1, tile, lat, head_dim=1).reshape(-1)
            assert torch.equal(back.long(), torch.arange(S))
            out.append(dict(latent=lat, window=win, tile=tile, text_len=tl, text_valid=tv,
                            pairs=int(dense.sum()), keys_q0=int(dense[0].sum()),
                            mask_bits=torch.from_numpy(__import__("numpy").packbits(dense.numpy())),
                            tile_perm=tiled.to(torch.int32)))
    finally:
        ref.saf.create_block_mask = orig
    save("tile_mask.pt", out)


def golden_router(ref):
    out = []
    for E, H, B, seed in [(1536, 12, 2, 21), (1024, 40, 1, 22), (768, 24, 1, 23)]:
        torch.manual_seed(seed)
        r = ref.router.Router(E, H, 3)
        temb = seeded((B, E), seed + 1, torch.float32)
        with torch.no_grad():
            score = r(temb)
        proc = ref.att.WanAttnProcessorTripleEval()
        dec = {}
        for tau in (0.0, 0.3, 0.36, 0.4, 0.5):
            s, idx = score[0].topk(1, dim=-1)           # wan.py:398-400
            idx = idx.clone()
            idx[s < tau] = 0
            dec[tau] = idx.squeeze(-1).to(torch.int32)
        out.append(dict(E=E, H=H, weight=r.linear.weight.detach().clone(), bias=r.linear.bias.detach().clone(),
                        temb=temb, score=score.clone(), decisions=dec))
    save("router.pt", out)


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def golden_wan_processor(ref):
    """Branch outputs and processor outputs of the reference Wan processors (fp32 arithmetic on CPU; inputs,
    weights and the post-projection q, k, v are bf16-representable so the GPU path sees identical values)."""
    c = FX.WAN_CASE
    lat, tile, win, lw, r, H = c["latent"], c["tile"], c["window"], c["lowres_window"], c["rate"], c["heads"]
    S = lat[0] * lat[1] * lat[2]
    attn = FX.FakeWanAttn(H, seed=31)
    hs = FX.det_tensor((1, S, H * D), 32, 1.0)
    rot = FX.wan_rotary(S, 33)
    info = ref.cs.get_group_info(lat, lw, reduction_rate=r)
    bm = ref.saf.create_sliding_tile_attn_mask_func(lat, win, tile, 0, 0, torch.device("cpu"))
    ev, tr = ref.att.WanAttnProcessorTripleEval(check_input=True), ref.att.WanAttnProcessorTripleTrain(check_input=True)
    kw = dict(lowres_group_info=info, flex_attn_mask_func=bm, window_size=win, tile_size=tile, latent_shape=lat)
    rec = {}
    with torch.no_grad():
        q, k, v, _ = ev._input_proj(attn, hs, None, rot)
        q, k, v = _bf16_round(q), _bf16_round(k), _bf16_round(v)
        rec.update(q=q.to(torch.bfloat16), k=k.to(torch.bfloat16), v=v.to(torch.bfloat16))
        rec["o_full"] = ev._attn(attn, q, k, v, None, None, False)[0]
        rec["o_coreset"] = ev._lowres_attn(attn, q, k, v, info)
        rec["o_sliding"] = ev._sliding_attn(q, k, v, bm, win, tile, lat)
        _, m = ref.cs.pool_sequence_by_similarity(q, info)
        rec["unpooled_argsort"], rec["pooled_argsort"] = m.unpooled_argsort_sim, m.pooled_argsort_sim
        mix = torch.tensor(FX.MIX)
        rec["eval_mix"] = ev(attn, hs, None, None, rot, tau_sparse=0.3, routing_score=mix, **kw).to(torch.bfloat16)
        rec["eval_mix_tau075"] = ev(attn, hs, None, None, rot, tau_sparse=0.75, routing_score=mix, **kw).to(torch.bfloat16)
        rec["train_mix"] = tr(attn, hs, None, None, rot, routing_score=mix, **kw).to(torch.bfloat16)
        rec["original"] = ev(attn, hs, None, None, rot, tau_sparse=0.3, routing_score=mix, use_original_attn=True,
                             **kw).to(torch.bfloat16)
        # SURVEY section 4 invariant 1: Train with one-hot scores == Eval
        onehot = torch.tensor([[[1., 0., 0.], [0., 1., 0.], [0., 0., 1.]]])
        a = tr(attn, hs, None, None, rot, routing_score=onehot, **kw)
        b = ev(attn, hs, None, None, rot, tau_sparse=0.3, routing_score=onehot, **kw)
        rec["train_onehot_equals_eval"] = bool(torch.equal(a, b))
    save("wan_processor.pt", rec)


def golden_hunyuan_processor(ref):
    c = FX.HUNYUAN_CASE
    lat, tile, win, lw, r, H = c["latent"], c["tile"], c["window"], c["lowres_window"], c["rate"], c["heads"]
    TL, TV = c["text_len"], c["text_valid"]
    S = lat[0] * lat[1] * lat[2]
    info = ref.cs.get_group_info(lat, lw, reduction_rate=r)
    bm = ref.saf.create_sliding_tile_attn_mask_func(lat, win, tile, TL, TV, torch.device("cpu"))
    mask = torch.zeros(1, 1, 1, S + TL, dtype=torch.bool)
    mask[..., :S + TV] = True                            # modeling_hunyuan.py:213-229
    rope = FX.hunyuan_rotary(S, 41)
    out = {}
    kw = dict(lowres_group_info=info, flex_attn_mask_func=bm, window_size=win, tile_size=tile, latent_shape=lat)
    for kind in ("dual", "single"):
        attn = FX.FakeHunyuanAttn(H, dual=kind == "dual", seed=42 if kind == "dual" else 43)
        hs = FX.det_tensor((1, S, H * D), 44, 1.0)
        ehs = FX.det_tensor((1, TL, H * D), 45, 1.0)
        ev = ref.att.HunyuanVideoFlashAttnProcessorTripleEval(check_input=True)
        tr = ref.att.HunyuanVideoFlashAttnProcessorTripleTrain(check_input=True)
        rec = {}
        with torch.no_grad():
            q, k, v = ev._step_to_qkv_and_unflatten(attn, hs, ehs)
            q, k = ev._step_qk_norm(attn, q, k)
            q, k = ev._step_rotary_emb(attn, q, k, TL, rope)
            q, k, v = ev._step_encoder_to_qkv_and_concat(attn, q, k, v, ehs)
            q, k, v = _bf16_round(q), _bf16_round(k), _bf16_round(v)
            rec.update(q=q.to(torch.bfloat16), k=k.to(torch.bfloat16), v=v.to(torch.bfloat16))
            rec["o_full"] = torch.cat(ev._step_attention(q, k, v, mask, TL), dim=2)
            rec["o_coreset"] = torch.cat(ev._step_lowres_attention(q, k, v, mask, TL, info), dim=2)
            rec["o_sliding"] = torch.cat(ev._step_sliding_attention(q, k, v, TL, bm, tile, lat), dim=2)
            mix = torch.tensor(FX.MIX)
            a, b = ev(attn, hs, ehs, mask, rope, routing_score=mix, tau_sparse=0.3, **kw)
            rec["eval_mix_video"], rec["eval_mix_text"] = a.to(torch.bfloat16), b.to(torch.bfloat16)
            a, b = tr(attn, hs, ehs, mask, rope, routing_score=mix, **kw)
            rec["train_mix_video"], rec["train_mix_text"] = a.to(torch.bfloat16), b.to(torch.bfloat16)
        out[kind] = rec
    save("hunyuan_processor.pt", out)


def _ulysses_worker(rank, world, port, shapes, ret):
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.cuda.synchronize = lambda *a, **k: None       # reference calls it unconditionally (ulysses/utils.py:49,81)
    ref = ref_loader.load()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref.uly.SP_STATE.setup_sp_group(world)
    B, H, S, d = shapes
    full = torch.arange(B * H * S * d, dtype=torch.float32).reshape(B, H, S, d)
    mine = full[:, :, rank * (S // world):(rank + 1) * (S // world)].contiguous()
    gathered = ref.uly.all_to_all_4D(mine, scatter_idx=1, gather_idx=2)
    back = ref.uly.all_to_all_4D(gathered, scatter_idx=2, gather_idx=1)
    ret[rank] = (gathered.clone(), bool(torch.equal(back, mine)))
    dist.destroy_process_group()


def golden_ulysses():
    import torch.multiprocessing as mp
    world, shapes = 4, (1, 8, 16, 4)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ulysses_worker, args=(world, 29731, shapes, ret), nprocs=world, join=True)
    save("ulysses.pt", dict(world=world, shape=shapes, gathered=[ret[r][0] for r in range(world)],
                            roundtrip_ok=[ret[r][1] for r in range(world)]))


def main():
    os.makedirs(OUT, exist_ok=True)
    which = set(sys.argv[1:])
    ref = ref_loader.load()
    steps = [("group", golden_group_info), ("coreset", golden_coreset), ("coreset_full", golden_coreset_fullsize),
             ("mask", golden_tile_and_mask),
             ("router", golden_router), ("wan", golden_wan_processor), ("hunyuan", golden_hunyuan_processor)]
    for name, fn in steps:
        if not which or name in which:
            fn(ref)
    if not which or "ulysses" in which:
        golden_ulysses()


if __name__ == "__main__":
    main()
