"""Random-init Wan 2.1 DiT shell that hosts the attention processors, for the "DiT denoise-step ms" metric.

The reference patches diffusers' ``WanTransformer3DModel`` (vorta/patch/modeling_wan.py); diffusers is not
installed here and pretrained weights are not reachable, so this module provides the same module tree
(``blocks[i].attn1 / attn2 / ffn / norm*``, ``condition_embedder.time_proj``, ``rope``, ``patch_embedding``,
``proj_out``, ``scale_shift_table``) with the public Wan 2.1 dimensions.  Everything except
self / cross attention is a library call (cuBLAS GEMMs, torch elementwise): those are the "next" rows of SURVEY.md
section 8f, not the hot path.  ``vorta_b200.patch.apply_vorta_transformer`` installs the routers, the routed
forward functions and the processors on it exactly as the reference does on the diffusers model.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import nn


@dataclass
class WanConfig:
    dim: int
    heads: int
    ffn_dim: int
    num_layers: int
    in_channels: int = 16
    out_channels: int = 16
    patch_size: Tuple[int, int, int] = (1, 2, 2)
    text_dim: int = 4096
    freq_dim: int = 256
    eps: float = 1e-6
    rope_max_seq_len: int = 1024

    @property
    def head_dim(self) -> int:
        return self.dim // self.heads


WAN_CONFIGS = {
    "wan2.1-t2v-1.3b": WanConfig(dim=1536, heads=12, ffn_dim=8960, num_layers=30),
    "wan2.1-t2v-14b": WanConfig(dim=5120, heads=40, ffn_dim=13824, num_layers=40),
    # not a real model: a 2-block, 4-head shell for the contract test of bench.py (tests/test_bench_contract.py)
    "wan-contract-test": WanConfig(dim=512, heads=4, ffn_dim=1024, num_layers=2, text_dim=64),
}


class FP32LayerNorm(nn.LayerNorm):
    """LayerNorm evaluated in fp32 whatever the parameter dtype (what diffusers' Wan blocks use); returns fp32
    when fed fp32, like the reference's ``self.norm1(hidden_states.float())`` call sites expect."""

    def forward(self, x):
        w = self.weight.float() if self.weight is not None else None
        b = self.bias.float() if self.bias is not None else None
        return torch.nn.functional.layer_norm(x.float(), self.normalized_shape, w, b, self.eps).to(x.dtype)


class Attention(nn.Module):
    """The container the processors are written against (diffusers ``Attention``): projections, qk norms and a
    pluggable processor; ``forward`` forwards every keyword to the processor (SURVEY.md section 8b "Caller")."""

    def __init__(self, dim: int, heads: int, eps: float = 1e-6, cross: bool = False):
        super().__init__()
        self.heads = heads
        self.is_cross = cross
        self.to_q = nn.Linear(dim, dim)
        self.to_k = nn.Linear(dim, dim)
        self.to_v = nn.Linear(dim, dim)
        self.norm_q = nn.RMSNorm(dim, eps=eps)        # rms_norm_across_heads (wan.py:85-89)
        self.norm_k = nn.RMSNorm(dim, eps=eps)
        self.add_k_proj = None
        self.add_v_proj = None
        self.norm_added_k = None
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        self.processor = None

    def set_processor(self, processor) -> None:
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kwargs)


class FeedForward(nn.Module):
    def __init__(self, dim: int, ffn_dim: int):
        super().__init__()
        self.proj_in = nn.Linear(dim, ffn_dim)
        self.act = nn.GELU(approximate="tanh")
        self.proj_out = nn.Linear(ffn_dim, dim)

    def forward(self, x):
        if x.is_cuda and x.dim() == 3:
            # bias + tanh-GELU in the cuBLASLt epilogue of the first GEMM: no separate pass over (S, ffn_dim)
            h = torch._addmm_activation(self.proj_in.bias, x.flatten(0, 1), self.proj_in.weight.t(), use_gelu=True)
            return self.proj_out(h).unflatten(0, x.shape[:2])
        return self.proj_out(self.act(self.proj_in(x)))


class WanTransformerBlock(nn.Module):
    """Module tree of a Wan block; ``forward`` is assigned by apply_vorta_transformer (modeling_wan.py:305)."""

    def __init__(self, cfg: WanConfig):
        super().__init__()
        self.norm1 = FP32LayerNorm(cfg.dim, eps=cfg.eps, elementwise_affine=False)
        self.attn1 = Attention(cfg.dim, cfg.heads, cfg.eps)
        self.norm2 = FP32LayerNorm(cfg.dim, eps=cfg.eps, elementwise_affine=True)
        self.attn2 = Attention(cfg.dim, cfg.heads, cfg.eps, cross=True)
        self.norm3 = FP32LayerNorm(cfg.dim, eps=cfg.eps, elementwise_affine=False)
        self.ffn = FeedForward(cfg.dim, cfg.ffn_dim)
        self.scale_shift_table = nn.Parameter(torch.randn(1, 6, cfg.dim) / cfg.dim ** 0.5)


class WanRotaryPosEmbed(nn.Module):
    def __init__(self, cfg: WanConfig):
        super().__init__()
        self.attention_head_dim = cfg.head_dim
        self.patch_size = cfg.patch_size
        d = cfg.head_dim
        dims = [d - 2 * (d // 3), d // 3, d // 3]            # 44 / 42 / 42 for head_dim 128 (modeling_wan.py:249-256)
        freqs = []
        for dd in dims:
            inv = 1.0 / (10000.0 ** (torch.arange(0, dd, 2, dtype=torch.float64) / dd))
            ang = torch.outer(torch.arange(cfg.rope_max_seq_len, dtype=torch.float64), inv)
            freqs.append(torch.polar(torch.ones_like(ang), ang))
        # plain attribute, not a buffer: Module.to(dtype) must not cast the complex table (diffusers keeps it the
        # same way and moves it in the rope forward, modeling_wan.py:248)
        self.freqs = torch.cat(freqs, dim=1)


class WanConditionEmbedder(nn.Module):
    def __init__(self, cfg: WanConfig):
        super().__init__()
        self.freq_dim = cfg.freq_dim
        self.time_embedder = nn.Sequential(nn.Linear(cfg.freq_dim, cfg.dim), nn.SiLU(), nn.Linear(cfg.dim, cfg.dim))
        self.act_fn = nn.SiLU()
        self.time_proj = nn.Linear(cfg.dim, cfg.dim * 6)
        self.text_embedder = nn.Sequential(nn.Linear(cfg.text_dim, cfg.dim), nn.GELU(approximate="tanh"),
                                           nn.Linear(cfg.dim, cfg.dim))

    def forward(self, timestep, encoder_hidden_states, encoder_hidden_states_image=None):
        half = self.freq_dim // 2
        exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timestep.device) / half
        ang = timestep.float()[:, None] * exponent.exp()[None]
        emb = torch.cat([ang.cos(), ang.sin()], dim=-1)
        dtype = self.time_embedder[0].weight.dtype
        temb = self.time_embedder(emb.to(dtype))
        timestep_proj = self.time_proj(self.act_fn(temb))
        encoder_hidden_states = self.text_embedder(encoder_hidden_states)
        return temb, timestep_proj, encoder_hidden_states, None


class WanDiT(nn.Module):
    """Module tree of WanTransformer3DModel; ``forward`` is assigned by apply_vorta_transformer."""

    def __init__(self, cfg: WanConfig):
        super().__init__()
        self.config = cfg
        p = cfg.patch_size
        self.rope = WanRotaryPosEmbed(cfg)
        self.patch_embedding = nn.Conv3d(cfg.in_channels, cfg.dim, kernel_size=p, stride=p)
        self.condition_embedder = WanConditionEmbedder(cfg)
        self.blocks = nn.ModuleList([WanTransformerBlock(cfg) for _ in range(cfg.num_layers)])
        self.norm_out = FP32LayerNorm(cfg.dim, eps=cfg.eps, elementwise_affine=False)
        self.proj_out = nn.Linear(cfg.dim, cfg.out_channels * p[0] * p[1] * p[2])
        self.scale_shift_table = nn.Parameter(torch.randn(1, 2, cfg.dim) / cfg.dim ** 0.5)
        self.gradient_checkpointing = False

    @staticmethod
    def build(name: str, device, dtype=torch.bfloat16, seed: int = 0) -> "WanDiT":
        """Random-init weights of the named architecture, created directly on the device."""
        cfg = WAN_CONFIGS[name]
        torch.manual_seed(seed)
        with torch.device(device):
            model = WanDiT(cfg)
        model = model.to(dtype)
        return model.eval().requires_grad_(False)
