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
