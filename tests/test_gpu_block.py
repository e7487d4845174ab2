"""Fused elementwise kernels of the DiT block against a plain PyTorch fp32 reference of
the same op (floating-point kernels: tolerance = one bf16 rounding of the fp32 result)."""
import pytest
import torch

from vorta_b200 import ops

pytestmark = pytest.mark.gpu


def _close(out, ref, atol=None):
    ref16 = ref.to(torch.bfloat16).float()
    err = (out.float() - ref).abs()
    # within one bf16 ulp of the fp32 reference
    tol = ref.abs() * 2.0 ** -7 + 1e-6
    assert bool((err <= tol).all()), float((err - tol).max())
    assert float((out.float() - ref16).abs().max()) <= float(tol.max())


@pytest.mark.parametrize("B,S,dim", [(1, 333, 1536), (2, 130, 5120), (1, 64, 384), (1, 71, 3072), (2, 9, 2048)])
def test_ln_modulate(B, S, dim):
    g = torch.Generator().manual_seed(dim)
    x = (torch.randn((B, S, dim), generator=g) * 2 + 0.3).to(torch.bfloat16).cuda()
    scale = (torch.randn((B, dim), generator=g) * 0.2).cuda()
    shift = (torch.randn((B, dim), generator=g) * 0.2).cuda()
    w = (1 + 0.1 * torch.randn((dim,), generator=g)).cuda()
    b = (0.1 * torch.randn((dim,), generator=g)).cuda()
    xf = x.float()
    ln = torch.nn.functional.layer_norm(xf, (dim,), None, None, 1e-6)
    _close(ops.ln_modulate(x, None, None, scale, shift, 1e-6), ln * (1 + scale[:, None]) + shift[:, None])
    _close(ops.ln_modulate(x, w, b, None, None, 1e-6), torch.nn.functional.layer_norm(xf, (dim,), w, b, 1e-6))
    _close(ops.ln_modulate(x, None, None, None, None, 1e-6), ln)


@pytest.mark.parametrize("B,S,dim", [(1, 333, 1536), (2, 130, 5120)])
def test_gate_residual(B, S, dim):
    g = torch.Generator().manual_seed(dim + 1)
    x = torch.randn((B, S, dim), generator=g).to(torch.bfloat16).cuda()
    y = torch.randn((B, S, dim), generator=g).to(torch.bfloat16).cuda()
    gate = torch.randn((B, dim), generator=g).cuda()
    _close(ops.gate_residual(x, y, gate), x.float() + y.float() * gate[:, None])
    _close(ops.gate_residual(x, y, None), x.float() + y.float())


@pytest.mark.parametrize("B,S,heads", [(1, 200, 12), (2, 77, 40), (1, 96, 3), (1, 53, 24), (1, 5, 16)])
def test_rmsnorm_rope(B, S, heads):
    dim = heads * 128
    g = torch.Generator().manual_seed(heads)
    x = torch.randn((B, S, dim), generator=g).to(torch.bfloat16).cuda()
    w = (1 + 0.1 * torch.randn((dim,), generator=g)).to(torch.bfloat16).cuda()
    ang = torch.rand((S, 64), generator=g, dtype=torch.float64) * 6.283185307179586
    cos, sin = ang.cos().float().cuda(), ang.sin().float().cuda()
    xf = x.float()
    normed = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()
    _close(ops.rmsnorm_rope(x, w, 1e-6, None, None), normed)
    # reference rotation: complex multiply of channel pairs per head (wan.py:34-37), in fp64
    z = torch.view_as_complex(normed.double().reshape(B, S, heads, 64, 2))
    rot = torch.view_as_real(z * torch.polar(torch.ones_like(ang), ang).cuda()[None, :, None]).reshape(B, S, dim).float()
    _close(ops.rmsnorm_rope(x, w, 1e-6, cos, sin), rot)


@pytest.mark.parametrize("B,S,T,heads", [(1, 200, 16, 24), (2, 77, 9, 3), (1, 1030, 256, 24), (1, 5, 0, 16)])
def test_headnorm_rope_joint_sequence(B, S, T, heads):
    """HunyuanVideo Q / K prologue (hunyuan.py:62-134): per-head RMSNorm, RoPE on the video rows only, rows placed in the
    joint [video | text] tensor — against fp32 torch (norm per head, the real-valued pair rotation of diffusers'
    apply_rotary_emb(use_real=True, unbind_dim=-1), torch.cat)."""
    dim = heads * 128
    g = torch.Generator().manual_seed(heads + S)
    xv = torch.randn((B, S, dim), generator=g).to(torch.bfloat16).cuda()
    xt = torch.randn((B, T, dim), generator=g).to(torch.bfloat16).cuda()
    wv = (1 + 0.1 * torch.randn((128,), generator=g)).to(torch.bfloat16).cuda()
    wt = (1 + 0.1 * torch.randn((128,), generator=g)).to(torch.bfloat16).cuda()
    ang = torch.rand((S, 64), generator=g) * 6.283185307179586
    cos, sin = ang.cos().cuda(), ang.sin().cuda()

    def ref(x, w, rope):
        xf = x.float().unflatten(2, (heads, 128))
        y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()
        if rope:
            c = cos.repeat_interleave(2, dim=1)[None, :, None]          # diffusers layout: (S, 128) pair-repeated
            s = sin.repeat_interleave(2, dim=1)[None, :, None]
            re, im = y.reshape(*y.shape[:-1], 64, 2).unbind(-1)
            y = y * c + torch.stack([-im, re], dim=-1).flatten(3) * s
        return y.flatten(2)

    joint = torch.full((B, S + T, dim), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.headnorm_rope(xv, wv, 1e-6, heads, cos, sin, out=joint)
    if T:
        ops.headnorm_rope(xt, wt, 1e-6, heads, out=joint, dst_row0=S)
    _close(joint, torch.cat([ref(xv, wv, True), ref(xt, wt, False)], dim=1))
    # single-stream form: one tensor over the joint sequence, in place, RoPE on the first S rows; no-norm copy
    both = torch.cat([xv, xt], dim=1)
    want = torch.cat([ref(xv, wv, True), ref(xt, wv, False)], dim=1)
    got = ops.headnorm_rope(both.clone(), wv, 1e-6, heads, cos, sin, rope_rows=S)
    _close(got, want)
    plain = torch.empty_like(joint)
    ops.headnorm_rope(xv, None, 0.0, heads, out=plain)
    assert torch.equal(plain[:, :S], xv)
    with pytest.raises(ValueError):
        ops.headnorm_rope(xv, wv, 1e-6, heads, out=joint, dst_row0=T + 1)        # does not fit


def test_hunyuan_fused_prologue_equals_eager_steps():
    """The processors' fused q / k / v path against the reference-shaped eager steps (same module, same inputs)."""
    from oracle import fixtures as FX
    from vorta_b200.attention import HunyuanVideoFlashAttnProcessor
    S, T, H = 128, 16, 3
    proc = HunyuanVideoFlashAttnProcessor()
    c, s = FX.hunyuan_rotary(S, 14)
    rope = (c.cuda(), s.cuda())
    for dual in (True, False):
        attn = FX.FakeHunyuanAttn(H, dual=dual, seed=11).to("cuda", torch.bfloat16)
        assert proc._qkv_fused(attn, FX.det_tensor((1, S, H * 128), 12).to("cuda", torch.bfloat16),
                               FX.det_tensor((1, T, H * 128), 13).to("cuda", torch.bfloat16), rope) is None   # grad mode
        hs = FX.det_tensor((1, S, H * 128), 12).to("cuda", torch.bfloat16)
        ehs = FX.det_tensor((1, T, H * 128), 13).to("cuda", torch.bfloat16)
        with torch.no_grad():
            fused = proc._qkv_fused(attn, hs, ehs, rope)
            q, k, v = proc._step_to_qkv_and_unflatten(attn, hs, ehs)
            q, k = proc._step_qk_norm(attn, q, k)
            q, k = proc._step_rotary_emb(attn, q, k, T, rope)
            eager = proc._step_encoder_to_qkv_and_concat(attn, q, k, v, ehs)
        assert fused is not None
        for a, b in zip(fused, eager):
            assert a.shape == b.shape
            scale = b.float().abs().max().item()
            assert (a.float() - b.float()).abs().max().item() <= 2.0 ** -6 * scale      # two bf16 roundings apart
        assert torch.equal(fused[2], eager[2])                                           # V: same GEMM, no arithmetic after
