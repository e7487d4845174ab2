"""Random-init HunyuanVideo DiT shell hosting the attention processors (SURVEY.md appendix E dimensions).

Same role as ``vorta_b200/dit/wan.py``: the module tree the reference patches on diffusers'
``HunyuanVideoTransformer3DModel`` (vorta/patch/modeling_hunyuan.py) — ``transformer_blocks[i].attn / norm1 / ...``,
``single_transformer_blocks[i].attn / norm / proj_mlp / proj_out``, ``time_text_embed``, ``rope``, ``x_embedder``,
``context_embedder`` — with everything outside the attention path as library calls.  The text token refiner is a
two-layer MLP stand-in (256 tokens, not on the path).  ``forward`` functions are assigned by
``vorta_b200.patch.modeling_hunyuan.apply_vorta_transformer``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Tuple

import torch
from torch import nn

from .wan import FP32LayerNorm


@dataclass
class HunyuanConfig:
    heads: int = 24
    head_dim: int = 128
    num_layers: int = 20                 # dual-stream blocks
    num_single_layers: int = 40
    mlp_ratio: float = 4.0
    in_channels: int = 16
    out_channels: int = 16
    patch_size: int = 2
    patch_size_t: int = 1
    text_embed_dim: int = 4096
    pooled_projection_dim: int = 768
    rope_theta: float = 256.0
    rope_axes_dim: Tuple[int, int, int] = (16, 56, 56)
    guidance_embeds: bool = True

    @property
    def dim(self) -> int:
        return self.heads * self.head_dim


HUNYUAN_CONFIGS = {"hunyuanvideo": HunyuanConfig()}


class HunyuanAttention(nn.Module):
    """Members the HunyuanVideo processors touch (hunyuan.py:49-54, 115-128, 202-207)."""

    def __init__(self, dim: int, heads: int, dual: bool):
        super().__init__()
        hd = dim // heads
        self.heads = heads
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.norm_q, self.norm_k = nn.RMSNorm(hd, eps=1e-6), nn.RMSNorm(hd, eps=1e-6)
        if dual:
            self.add_q_proj, self.add_k_proj, self.add_v_proj = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
            self.norm_added_q, self.norm_added_k = nn.RMSNorm(hd, eps=1e-6), nn.RMSNorm(hd, eps=1e-6)
            self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
            self.to_add_out = nn.Linear(dim, dim)
        else:                                   # single-stream block: pre-only attention
            self.add_q_proj = self.add_k_proj = self.add_v_proj = None
            self.norm_added_q = self.norm_added_k = None
            self.to_out = None
            self.to_add_out = None
        self.processor = None

    def set_processor(self, processor) -> None:
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kwargs)


class AdaLayerNormZero(nn.Module):
    """emb -> 6 modulation vectors; the normalisation itself is applied by the (fused) block forward."""

    def __init__(self, dim: int, n_out: int = 6):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(dim, n_out * dim)
        self.norm = FP32LayerNorm(dim, eps=1e-6, elementwise_affine=False)
        self.n_out = n_out

    def forward(self, x: torch.Tensor, emb: torch.Tensor):
        """diffusers-style call (AdaLayerNormZero: norm(x) * (1 + scale) + shift and the remaining vectors); the routed
        block forwards do not use it — they read ``silu`` / ``linear`` and fuse the normalisation."""
        v = self.linear(self.silu(emb)).float().chunk(self.n_out, dim=1)
        if self.n_out == 2:                       # AdaLayerNormContinuous: (scale, shift)
            return (self.norm(x).float() * (1 + v[0][:, None]) + v[1][:, None]).to(x.dtype)
        y = (self.norm(x).float() * (1 + v[1][:, None]) + v[0][:, None]).to(x.dtype)
        return (y, *v[2:])


class MLP(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.proj_in = nn.Linear(dim, hidden)
        self.act = nn.GELU(approximate="tanh")
        self.proj_out = nn.Linear(hidden, dim)

    def forward(self, x):
        if x.is_cuda and x.dim() == 3:
            h = torch._addmm_activation(self.proj_in.bias, x.flatten(0, 1), self.proj_in.weight.t(), use_gelu=True)
            return self.proj_out(h).unflatten(0, x.shape[:2])
        return self.proj_out(self.act(self.proj_in(x)))


class HunyuanVideoTransformerBlock(nn.Module):
    def __init__(self, cfg: HunyuanConfig):
        super().__init__()
        d, hidden = cfg.dim, int(cfg.dim * cfg.mlp_ratio)
        self.norm1, self.norm1_context = AdaLayerNormZero(d), AdaLayerNormZero(d)
        self.attn = HunyuanAttention(d, cfg.heads, dual=True)
        self.norm2 = FP32LayerNorm(d, eps=1e-6, elementwise_affine=False)
        self.norm2_context = FP32LayerNorm(d, eps=1e-6, elementwise_affine=False)
        self.ff, self.ff_context = MLP(d, hidden), MLP(d, hidden)


class HunyuanVideoSingleTransformerBlock(nn.Module):
    def __init__(self, cfg: HunyuanConfig):
        super().__init__()
        d, hidden = cfg.dim, int(cfg.dim * cfg.mlp_ratio)
        self.norm = AdaLayerNormZero(d, n_out=3)
        self.proj_mlp = nn.Linear(d, hidden)
        self.act_mlp = nn.GELU(approximate="tanh")
        self.attn = HunyuanAttention(d, cfg.heads, dual=False)
        self.proj_out = nn.Linear(d + hidden, d)


class HunyuanVideoRotaryPosEmbed(nn.Module):
    def __init__(self, cfg: HunyuanConfig):
        super().__init__()
        self.patch_size, self.patch_size_t = cfg.patch_size, cfg.patch_size_t
        self.rope_dim, self.theta = cfg.rope_axes_dim, cfg.rope_theta


class TimestepMLP(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1, self.act, self.linear_2 = nn.Linear(in_dim, dim), nn.SiLU(), nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class TokenRefinerStandIn(TimestepMLP):
    """Stand-in for diffusers' HunyuanVideoTokenRefiner with its call signature ``(hidden_states, timestep,
    attention_mask)`` (reference call site: modeling_hunyuan.py:210).  The real refiner is a 2-block transformer over the
    256 text tokens conditioned on the timestep — outside the routed-attention path; here it is a per-token MLP."""

    def forward(self, hidden_states, timestep=None, attention_mask=None):
        return super().forward(hidden_states)


class HunyuanVideoConditionEmbedding(nn.Module):
    def __init__(self, cfg: HunyuanConfig):
        super().__init__()
        self.freq_dim = 256
        self.timestep_embedder = TimestepMLP(self.freq_dim, cfg.dim)
        self.text_embedder = TimestepMLP(cfg.pooled_projection_dim, cfg.dim)
        self.guidance_embedder = TimestepMLP(self.freq_dim, cfg.dim) if cfg.guidance_embeds else None
        self.image_condition_type = None

    def time_proj(self, t: torch.Tensor) -> torch.Tensor:
        half = self.freq_dim // 2
        exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
        ang = t.float()[:, None] * exponent.exp()[None]
        return torch.cat([ang.cos(), ang.sin()], dim=-1)


class HunyuanDiT(nn.Module):
    def __init__(self, cfg: HunyuanConfig):
        super().__init__()
        self.config = cfg
        p, pt = cfg.patch_size, cfg.patch_size_t
        self.x_embedder = nn.Conv3d(cfg.in_channels, cfg.dim, kernel_size=(pt, p, p), stride=(pt, p, p))
        self.context_embedder = TokenRefinerStandIn(cfg.text_embed_dim, cfg.dim)
        self.time_text_embed = HunyuanVideoConditionEmbedding(cfg)
        self.rope = HunyuanVideoRotaryPosEmbed(cfg)
        self.transformer_blocks = nn.ModuleList([HunyuanVideoTransformerBlock(cfg) for _ in range(cfg.num_layers)])
        self.single_transformer_blocks = nn.ModuleList(
            [HunyuanVideoSingleTransformerBlock(cfg) for _ in range(cfg.num_single_layers)])
        self.norm_out = AdaLayerNormZero(cfg.dim, n_out=2)
        self.proj_out = nn.Linear(cfg.dim, pt * p * p * cfg.out_channels)
        self.gradient_checkpointing = False

    @staticmethod
    def build(name_or_cfg, device, dtype=torch.bfloat16, seed: int = 0) -> "HunyuanDiT":
        cfg = HUNYUAN_CONFIGS[name_or_cfg] if isinstance(name_or_cfg, str) else name_or_cfg
        torch.manual_seed(seed)
        with torch.device(device):
            model = HunyuanDiT(cfg)
        return model.to(dtype).eval().requires_grad_(False)
