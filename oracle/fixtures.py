"""Deterministic, storage-free test inputs shared by oracle/make_golden.py and tests/ — TEST INFRASTRUCTURE ONLY.

``det_tensor`` fills a tensor from an integer hash of the element index, mapped to k/128 * scale with k an 8-bit
integer and scale a power of two, so every value is exactly representable in bf16 and identical on every machine
(no dependence on a library RNG).  The golden files therefore only need to store the reference's OUTPUTS.
"""
from __future__ import annotations

import math

import torch
from torch import nn

D = 128
_M32 = (1 << 32) - 1


def det_tensor(shape, seed: int, scale: float = 1.0, dtype=torch.float32) -> torch.Tensor:
    n = int(math.prod(shape))
    x = (torch.arange(n, dtype=torch.int64) * 2654435761 + (seed + 1) * 40503) & _M32
    x = (x ^ (x >> 15)) & _M32
    x = (x * 2246822519) & _M32
    x = (x ^ (x >> 13)) & _M32
    x = (x * 3266489917) & _M32
    x = (x ^ (x >> 16)) & _M32
    k = (x & 255) - 128                                    # 8-bit signed: exact in bf16
    return (k.to(torch.float64) / 128.0 * scale).to(dtype).reshape(shape)


def fill_module_(module: nn.Module, seed: int) -> nn.Module:
    """Overwrite every parameter with det_tensor values (weights ~ U(-1,1)/sqrt(fan_in) rounded to a power of two,
    biases small, norm weights near one)."""
    with torch.no_grad():
        for i, (name, p) in enumerate(sorted(module.named_parameters())):
            if p.dim() == 2:
                scale = 2.0 ** round(math.log2(1.5 / math.sqrt(p.shape[1])))
                p.copy_(det_tensor(tuple(p.shape), seed * 1000 + i, scale))
            elif "norm" in name:
                p.copy_(1.0 + det_tensor(tuple(p.shape), seed * 1000 + i, 2.0 ** -3))
            else:
                p.copy_(det_tensor(tuple(p.shape), seed * 1000 + i, 2.0 ** -4))
    return module


class FakeWanAttn(nn.Module):
    """Stand-in for diffusers' Attention with the members the Wan processors touch
    (reference: vorta/attention/wan.py:72-94, 122-127, 158-159)."""

    def __init__(self, heads: int, seed: int = 0):
        super().__init__()
        hd = heads * D
        self.heads = heads
        self.to_q, self.to_k, self.to_v = nn.Linear(hd, hd), nn.Linear(hd, hd), nn.Linear(hd, hd)
        self.norm_q, self.norm_k = nn.RMSNorm(hd, eps=1e-6), nn.RMSNorm(hd, eps=1e-6)
        self.add_k_proj = None
        self.add_v_proj = None
        self.norm_added_k = None
        self.to_out = nn.ModuleList([nn.Linear(hd, hd), nn.Dropout(0.0)])
        fill_module_(self, seed)


class FakeHunyuanAttn(nn.Module):
    """Stand-in with the members the HunyuanVideo processors touch (hunyuan.py:49-54, 115-128, 202-207).
    dual=True: MM-DiT dual-stream block (separate text projections); False: single-stream block."""

    def __init__(self, heads: int, dual: bool, seed: int = 0):
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
        fill_module_(self, seed)


def wan_rotary(S: int, seed: int) -> torch.Tensor:
    """complex128 phases of shape (1, 1, S, D/2), the layout wan.py:34-37 multiplies by."""
    ang = (det_tensor((1, 1, S, D // 2), seed, 1.0, torch.float64) + 1.0) * math.pi
    return torch.polar(torch.ones_like(ang), ang)


def hunyuan_rotary(S: int, seed: int):
    """(cos, sin), each (S, D), in the pair-repeated layout diffusers' apply_rotary_emb expects."""
    ang = (det_tensor((S, D // 2), seed, 1.0, torch.float32) + 1.0) * math.pi
    return ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1)


# geometry of the processor-level golden cases
WAN_CASE = dict(latent=(4, 6, 8), tile=(2, 3, 4), window=(3, 3, 3), lowres_window=(2, 3, 2), rate=0.5, heads=3)
HUNYUAN_CASE = dict(latent=(2, 8, 8), tile=(1, 4, 4), window=(3, 3, 3), lowres_window=(2, 2, 2), rate=0.5, heads=3,
                    text_len=16, text_valid=11)
MIX = [[[0.7, 0.2, 0.1], [0.1, 0.8, 0.1], [0.2, 0.2, 0.6]]]     # heads -> full, coreset, sliding
