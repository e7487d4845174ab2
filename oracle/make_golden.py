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
    for lat, win, r, h, seed in [((4, 6, 8), (2, 3, 2), 0.5, 3, 11), ((6, 6, 8), (3, 3, 2), 0.5, 2, 12),
                                 ((4, 8, 12), (2, 2, 2), 0.5, 3, 13), ((4, 6, 8), (2, 3, 2), 0.75, 2, 14)]:
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
    for E, H, B, seed in [(1536, 12, 2, 21), (5120, 40, 1, 22), (3072, 24, 1, 23)]:
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


class FakeWanAttn(nn.Module):
    """Stand-in for diffusers' Attention with the members the Wan processors touch (wan.py:72-94,158-159)."""

    def __init__(self, heads):
        super().__init__()
        hd = heads * D
        self.heads = heads
        self.to_q, self.to_k, self.to_v = nn.Linear(hd, hd), nn.Linear(hd, hd), nn.Linear(hd, hd)
        self.norm_q, self.norm_k = nn.RMSNorm(hd, eps=1e-6), nn.RMSNorm(hd, eps=1e-6)
        self.add_k_proj = None
        self.to_out = nn.ModuleList([nn.Linear(hd, hd), nn.Dropout(0.0)])


def wan_rotary(S, seed):
    g = torch.Generator().manual_seed(seed)
    ang = torch.rand(1, 1, S, D // 2, generator=g, dtype=torch.float64) * 6.283185307179586
    return torch.polar(torch.ones_like(ang), ang)      # complex128, (1, 1, S, D/2)  (wan.py:34-37)


def golden_wan_processor(ref):
    """Branch outputs and processor outputs of the reference Wan processors (fp32 on CPU)."""
    lat, tile, win, lw, r = (4, 6, 8), (2, 3, 4), (3, 3, 3), (2, 3, 2), 0.5
    S, H = lat[0] * lat[1] * lat[2], 3
    torch.manual_seed(31)
    attn = FakeWanAttn(H)
    hs = seeded((1, S, H * D), 32, torch.float32) * 1.0
    rot = wan_rotary(S, 33)
    info = ref.cs.get_group_info(lat, lw, reduction_rate=r)
    bm = ref.saf.create_sliding_tile_attn_mask_func(lat, win, tile, 0, 0, torch.device("cpu"))
    ev, tr = ref.att.WanAttnProcessorTripleEval(check_input=True), ref.att.WanAttnProcessorTripleTrain(check_input=True)
    kw = dict(lowres_group_info=info, flex_attn_mask_func=bm, window_size=win, tile_size=tile, latent_shape=lat)
    rec = dict(latent=lat, tile=tile, window=win, lowres_window=lw, rate=r, heads=H, hidden_states=hs, rotary=rot,
               state_dict={k: v.clone() for k, v in attn.state_dict().items()})
    with torch.no_grad():
        q, k, v, _ = ev._input_proj(attn, hs, None, rot)
        rec.update(q=q.clone(), k=k.clone(), v=v.clone())
        rec["o_full"] = ev._attn(attn, q, k, v, None, None, False)[0]
        rec["o_coreset"] = ev._lowres_attn(attn, q, k, v, info)
        rec["o_sliding"] = ev._sliding_attn(q, k, v, bm, win, tile, lat)
        outs = {}
        for name, score in (("full", [1., 0., 0.]), ("coreset", [0., 1., 0.]), ("sliding", [0., 0., 1.])):
            sc = torch.tensor(score).repeat(1, H, 1)
            outs[f"eval_{name}"] = ev(attn, hs, None, None, rot, tau_sparse=0.3, routing_score=sc, **kw)
        mix = torch.tensor([[[0.7, 0.2, 0.1], [0.1, 0.8, 0.1], [0.2, 0.2, 0.6]]])
        outs["eval_mix"] = ev(attn, hs, None, None, rot, tau_sparse=0.3, routing_score=mix, **kw)
        outs["eval_mix_tau075"] = ev(attn, hs, None, None, rot, tau_sparse=0.75, routing_score=mix, **kw)
        outs["train_mix"] = tr(attn, hs, None, None, rot, routing_score=mix, **kw)
        outs["original"] = ev(attn, hs, None, None, rot, tau_sparse=0.3, routing_score=mix, use_original_attn=True, **kw)
        rec["mix"] = mix
        rec.update(outs)
    save("wan_processor.pt", rec)


class FakeHunyuanAttn(nn.Module):
    """Stand-in with the members the HunyuanVideo processors touch (hunyuan.py:49-54,115-128,202-207)."""

    def __init__(self, heads, dual):
        super().__init__()
        hd = heads * D
        self.heads = heads
        self.to_q, self.to_k, self.to_v = nn.Linear(hd, hd), nn.Linear(hd, hd), nn.Linear(hd, hd)
        self.norm_q, self.norm_k = nn.RMSNorm(D, eps=1e-6), nn.RMSNorm(D, eps=1e-6)
        if dual:
            self.add_q_proj, self.add_k_proj, self.add_v_proj = nn.Linear(hd, hd), nn.Linear(hd, hd), nn.Linear(hd, hd)
            self.norm_added_q, self.norm_added_k = nn.RMSNorm(D, eps=1e-6), nn.RMSNorm(D, eps=1e-6)
            self.to_out = nn.ModuleList([nn.Linear(hd, hd), nn.Dropout(0.0)])
            self.to_add_out = nn.Linear(hd, hd)
        else:
            self.add_q_proj = self.add_k_proj = self.add_v_proj = None
            self.norm_added_q = self.norm_added_k = None
            self.to_out = None
            self.to_add_out = None


def golden_hunyuan_processor(ref):
    lat, tile, win, lw, r = (4, 8, 8), (2, 4, 4), (3, 3, 3), (2, 2, 2), 0.5
    S, H, TL, TV = lat[0] * lat[1] * lat[2], 3, 16, 11
    info = ref.cs.get_group_info(lat, lw, reduction_rate=r)
    bm = ref.saf.create_sliding_tile_attn_mask_func(lat, win, tile, TL, TV, torch.device("cpu"))
    mask = torch.zeros(1, 1, 1, S + TL, dtype=torch.bool)
    mask[..., :S + TV] = True                            # modeling_hunyuan.py:213-229
    g = torch.Generator().manual_seed(41)
    ang = torch.rand(S, D // 2, generator=g) * 6.283185307179586
    rope = (ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1))
    out = dict(latent=lat, tile=tile, window=win, lowres_window=lw, rate=r, heads=H, text_len=TL, text_valid=TV,
               attention_mask=mask, rope_cos=rope[0], rope_sin=rope[1])
    kw = dict(lowres_group_info=info, flex_attn_mask_func=bm, window_size=win, tile_size=tile, latent_shape=lat)
    for kind in ("dual", "single"):
        torch.manual_seed(42 if kind == "dual" else 43)
        attn = FakeHunyuanAttn(H, dual=kind == "dual")
        hs = seeded((1, S, H * D), 44, torch.float32)
        ehs = seeded((1, TL, H * D), 45, torch.float32)
        ev = ref.att.HunyuanVideoFlashAttnProcessorTripleEval(check_input=True)
        tr = ref.att.HunyuanVideoFlashAttnProcessorTripleTrain(check_input=True)
        rec = dict(hidden_states=hs, encoder_hidden_states=ehs,
                   state_dict={k: v.clone() for k, v in attn.state_dict().items()})
        with torch.no_grad():
            q, k, v = ev._step_to_qkv_and_unflatten(attn, hs, ehs)
            q, k = ev._step_qk_norm(attn, q, k)
            q, k = ev._step_rotary_emb(attn, q, k, TL, rope)
            q, k, v = ev._step_encoder_to_qkv_and_concat(attn, q, k, v, ehs)
            rec.update(q=q.clone(), k=k.clone(), v=v.clone())
            of = ev._step_attention(q, k, v, mask, TL)
            oc = ev._step_lowres_attention(q, k, v, mask, TL, info)
            os_ = ev._step_sliding_attention(q, k, v, TL, bm, tile, lat)
            rec["o_full"] = torch.cat(of, dim=2)
            rec["o_coreset"] = torch.cat(oc, dim=2)
            rec["o_sliding"] = torch.cat(os_, dim=2)
            mix = torch.tensor([[[0.7, 0.2, 0.1], [0.1, 0.8, 0.1], [0.2, 0.2, 0.6]]])
            rec["mix"] = mix
            a, b = ev(attn, hs, ehs, mask, rope, routing_score=mix, tau_sparse=0.3, **kw)
            rec["eval_mix_video"], rec["eval_mix_text"] = a, b
            a, b = tr(attn, hs, ehs, mask, rope, routing_score=mix, **kw)
            rec["train_mix_video"], rec["train_mix_text"] = a, b
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
    steps = [("group", golden_group_info), ("coreset", golden_coreset), ("mask", golden_tile_and_mask),
             ("router", golden_router), ("wan", golden_wan_processor), ("hunyuan", golden_hunyuan_processor)]
    for name, fn in steps:
        if not which or name in which:
            fn(ref)
    if not which or "ulysses" in which:
        golden_ulysses()


if __name__ == "__main__":
    main()
