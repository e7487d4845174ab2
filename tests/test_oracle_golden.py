"""The CPU oracle (oracle/vorta_oracle.py) against the golden vectors produced by the REFERENCE's own code
(oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import vorta_oracle as O


def test_group_info_matches_reference(golden):
    for rec in golden("group_info.pt"):
        info = O.get_group_info(rec["latent"], rec["window"], rec["rate"])
        assert info.num_unpooled_tokens_per_group == rec["n_unpooled"]
        assert (tuple(info.center_indices.shape), tuple(info.margin_indices.shape)) == rec["shape"]
        assert int(info.center_indices.sum()) == rec["center_sum"]
        assert int(info.margin_indices.sum()) == rec["margin_sum"]
        assert torch.equal(info.center_indices[:16], rec["center_head"])
        assert torch.equal(info.margin_indices[-4:], rec["margin_tail"])
        if rec["center"] is not None:
            assert torch.equal(info.center_indices, rec["center"])
            assert torch.equal(info.margin_indices, rec["margin"])


def test_reference_main_known_answer():
    # the reference's own __main__ demo (coreset_select.py:191-203)
    info = O.get_group_info((4, 6, 4), (2, 3, 2), 0.5)
    assert info.center_indices[:, 0].tolist() == [29, 31, 41, 43, 77, 79, 89, 91]
    assert info.num_unpooled_tokens_per_group == 5
    assert tuple(info.margin_indices.shape) == (8, 11)


@pytest.mark.parametrize("tag,dtype", [("f32", torch.float32), ("f64", torch.float64)])
def test_matching_pool_unpool_match_reference(golden, tag, dtype):
    for rec in golden("coreset.pt"):
        info = O.get_group_info(rec["latent"], rec["window"], rec["rate"])
        x = rec["x"].to(dtype)
        un, po = O.match(x, info)
        assert torch.equal(un, rec[f"unpooled_{tag}"])
        assert torch.equal(po, rec[f"pooled_{tag}"])
        if tag == "f32":
            pooled = O.pool(x, info, (un, po))
            assert torch.equal(pooled.to(torch.bfloat16), rec["pooled_seq"])
            assert torch.equal(O.pool(rec["k"].float(), info, (un, po)).to(torch.bfloat16),
                               rec["pooled_k_with_q_matching"])
            assert torch.equal(O.unpool(rec["y"].float(), info, (un, po)).to(torch.bfloat16), rec["unpooled_seq"])


def test_tile_permutation_and_mask_match_reference(golden):
    for rec in golden("tile_mask.pt"):
        lat, win, tile, tl, tv = rec["latent"], rec["window"], rec["tile"], rec["text_len"], rec["text_valid"]
        assert torch.equal(O.tile_permutation(lat, tile).to(torch.int32), rec["tile_perm"])
        mask = O.sliding_tile_mask(lat, win, tile, tl, tv)
        n = mask.shape[0]
        ref = torch.from_numpy(np.unpackbits(rec["mask_bits"].numpy())[:n * n].reshape(n, n).astype(bool))
        assert torch.equal(mask, ref)
        assert int(mask.sum()) == rec["pairs"]
        assert int(mask[0].sum()) == rec["keys_q0"]


def test_mask_known_answers():
    # allowed-pair counts probed from the reference
    for lat, win, tile, tl, tv, pairs, keys in [((6, 12, 12), (1, 3, 3), (2, 4, 4), 0, 0, 248832, 288),
                                                ((8, 8, 12), (3, 3, 3), (2, 2, 4), 0, 0, 331776, 432),
                                                ((4, 8, 8), (3, 3, 3), (2, 4, 4), 5, 3, 67081, 259),
                                                ((10, 9, 8), (3, 3, 3), (5, 9, 8), 0, 0, 518400, 720)]:
        m = O.sliding_tile_mask(lat, win, tile, tl, tv)
        assert int(m.sum()) == pairs
        assert int(m[0].sum()) == keys


def test_tile_windows_equal_dense_mask():
    for lat, win, tile in [((4, 8, 12), (3, 3, 3), (2, 4, 4)), ((4, 6, 8), (2, 3, 1), (1, 3, 2)),
                           ((6, 4, 4), (5, 1, 3), (1, 2, 2))]:
        mask = O.sliding_tile_mask(lat, win, tile)
        tau = tile[0] * tile[1] * tile[2]
        nt = [lat[d] // tile[d] for d in range(3)]
        wins = O.tile_windows(lat, win, tile)
        tile_mask = mask[::tau, ::tau]
        for t in range(wins.shape[0]):
            lo, hi = wins[t, :3], wins[t, 3:]
            allowed = {(a * nt[1] + b) * nt[2] + c for a in range(lo[0], hi[0] + 1) for b in range(lo[1], hi[1] + 1)
                       for c in range(lo[2], hi[2] + 1)}
            assert set(torch.nonzero(tile_mask[t]).flatten().tolist()) == allowed


def test_router_matches_reference(golden):
    for rec in golden("router.pt"):
        score = O.router_forward(rec["temb"], rec["weight"], rec["bias"], rec["H"])
        assert torch.allclose(score, rec["score"], atol=1e-6, rtol=1e-5)
        for tau, dec in rec["decisions"].items():
            assert torch.equal(O.route_top1(score, tau).to(torch.int32), dec)


def _close(a, b, cos_min=0.99999, max_abs=1e-4):
    a, b = a.float(), b.float()
    cos = torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()
    assert cos >= cos_min, cos
    assert (a - b).abs().max().item() <= max_abs, (a - b).abs().max().item()


def test_wan_branches_match_reference(golden):
    rec, c = golden("wan_processor.pt"), FX.WAN_CASE
    info = O.get_group_info(c["latent"], c["lowres_window"], c["rate"])
    q, k, v = rec["q"].float(), rec["k"].float(), rec["v"].float()
    assert rec["train_onehot_equals_eval"]
    _close(O.full_attention(q, k, v), rec["o_full"])
    un, po = O.match(q, info)
    assert torch.equal(un, rec["unpooled_argsort"]) and torch.equal(po, rec["pooled_argsort"])
    _close(O.coreset_attention(q, k, v, info), rec["o_coreset"])
    _close(O.sliding_tile_attention(q, k, v, c["latent"], c["window"], c["tile"], dense=True), rec["o_sliding"])
    _close(O.sliding_tile_attention(q, k, v, c["latent"], c["window"], c["tile"], dense=False), rec["o_sliding"])
    # routed / blended combination from the branch outputs
    branch = torch.tensor([0, 1, 2])
    routed = O.routed_attention(q, k, v, info, c["latent"], c["window"], c["tile"], branch=branch)
    ref = torch.stack([rec["o_full"][:, 0], rec["o_coreset"][:, 1], rec["o_sliding"][:, 2]], dim=1)
    _close(routed, ref)
    w = torch.tensor(FX.MIX)
    blended = O.routed_attention(q, k, v, info, c["latent"], c["window"], c["tile"], weights=w)
    ref = (w[:, :, :, None, None] * torch.stack([rec["o_full"], rec["o_coreset"], rec["o_sliding"]], dim=2)).sum(2)
    _close(blended, ref)


@pytest.mark.parametrize("kind", ["dual", "single"])
def test_hunyuan_branches_match_reference(golden, kind):
    rec, c = golden("hunyuan_processor.pt")[kind], FX.HUNYUAN_CASE
    info = O.get_group_info(c["latent"], c["lowres_window"], c["rate"])
    tl, tv = c["text_len"], c["text_valid"]
    q, k, v = rec["q"].float(), rec["k"].float(), rec["v"].float()
    _close(O.full_attention(q, k, v, tl, tv), rec["o_full"])
    _close(O.coreset_attention(q, k, v, info, tl, tv, kv_from_k=True), rec["o_coreset"])
    _close(O.sliding_tile_attention(q, k, v, c["latent"], c["window"], c["tile"], tl, tv, dense=True), rec["o_sliding"])
    _close(O.sliding_tile_attention(q, k, v, c["latent"], c["window"], c["tile"], tl, tv, dense=False),
           rec["o_sliding"])
    # SURVEY.md section 4 invariant 4: padded text query rows are exactly zero in all three branches
    S = q.shape[2] - tl
    for name in ("o_full", "o_coreset", "o_sliding"):
        assert rec[name][:, :, S + tv:].abs().max().item() == 0.0


def test_ulysses_permutation_matches_reference(golden):
    rec = golden("ulysses.pt")
    B, H, S, d = rec["shape"]
    P = rec["world"]
    assert all(rec["roundtrip_ok"])
    full = torch.arange(B * H * S * d, dtype=torch.float32).reshape(B, H, S, d)
    shards = [full[:, :, r * (S // P):(r + 1) * (S // P)].contiguous() for r in range(P)]
    gathered = O.ulysses_scatter_heads(shards)
    for r in range(P):
        assert torch.equal(gathered[r], rec["gathered"][r])
    back = O.ulysses_gather_heads(gathered)
    for r in range(P):
        assert torch.equal(back[r], shards[r])


def test_oracle_matches_reference_fullsize_coreset_tables(golden):
    """BASELINE-size index tables: the oracle run in fp64 equals the reference run in fp64 (first heads of every case
    of tests/golden/coreset_fullsize.pt), and where the reference's own fp32 run differs from its fp64 run the fp64
    cosine gap of the swapped margins is far below what fp32 accumulation can resolve."""
    from oracle.make_golden import fullsize_coreset_input
    for rec in golden("coreset_fullsize.pt"):
        info = O.get_group_info(rec["latent"], rec["window"], rec["rate"])
        assert info.num_unpooled_tokens_per_group == rec["n_unpooled"]
        x = fullsize_coreset_input(rec["latent"], rec["heads"], rec["seed"], rec["smooth"])[:, :2]
        un, po = O.match(x.double(), info)
        assert torch.equal(un.to(torch.uint8), rec["unpooled_f64"][:, :2])
        assert torch.equal(po.to(torch.uint8), rec["pooled_f64"][:, :2])
        assert rec["f32_differs_at"].shape[0] == rec["f32_gap"].numel() == rec["f32_order"].shape[0]
        if rec["f32_gap"].numel():
            assert rec["f32_gap"].max().item() <= 4e-5          # kSimGap of vb_coreset_select_kernel
