"""Install boundary for Wan — the reference's ``vorta/patch/modeling_wan.py`` (apply_vorta_transformer :265-310,
wan_block_routed_forward :195-239, wan_transformer_3d_routed_forward :38-192, wan_rope_forward :242-262).

Works on any model with the Wan module tree (diffusers' WanTransformer3DModel or ``vorta_b200.dit.WanDiT``).
Differences from the reference, by design:
* the routers of ALL blocks are evaluated in one launch at the start of the forward (they only read ``temb``,
  modeling_wan.py:215) and their top-1 decisions reach the host in one copy, instead of one sync per block;
* under Ulysses the sequence is sharded by TOKENS inside the forward (the reference shards frames in the
  pipeline, pipeline_wan.py:120-122), and the RoPE table is narrowed per rank in the processor (wan.py:97);
* the router-training losses of the reference forward (reg / distillation) are evaluated in the forward direction
  only; training through them needs the attention backward kernels, which do not exist yet.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import torch
import torch.nn.functional as F

from .. import ops
from ..attention import WanAttnProcessor2_0, WanAttnProcessorTripleEval, WanAttnProcessorTripleTrain
from ..ulysses import SP_STATE, all_gather
from .outputs import RoutedTransformerModelOutput
from .router import Router, route_step


def wan_rope_forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
    """(1, 1, T*H*W, D/2) complex phases for the FULL token grid (modeling_wan.py:242-262); every rank builds the
    whole table and the processor narrows it to its token shard."""
    _, _, num_frames, height, width = hidden_states.shape
    p_t, p_h, p_w = self.patch_size
    ppf, pph, ppw = num_frames // p_t, height // p_h, width // p_w
    d = self.attention_head_dim
    freqs = self.freqs.to(hidden_states.device).split_with_sizes([d // 2 - 2 * (d // 6), d // 6, d // 6], dim=1)
    f = freqs[0][:ppf].view(ppf, 1, 1, -1).expand(ppf, pph, ppw, -1)
    h = freqs[1][:pph].view(1, pph, 1, -1).expand(ppf, pph, ppw, -1)
    w = freqs[2][:ppw].view(1, 1, ppw, -1).expand(ppf, pph, ppw, -1)
    return torch.cat([f, h, w], dim=-1).reshape(1, 1, ppf * pph * ppw, -1)


def wan_block_routed_forward(self, hidden_states, encoder_hidden_states, temb, rotary_emb,
                             temb_before_proj: Optional[torch.Tensor], use_original_attn: bool = False,
                             self_attention_kwargs: Optional[Dict[str, Any]] = None,
                             routing_score: Optional[torch.Tensor] = None, branch=None):
    """Dataflow of modeling_wan.py:195-239.  ``routing_score`` / ``branch`` may be supplied by the caller when the
    step's routing was precomputed; otherwise the block's router runs here like in the reference."""
    # adaLN parameters in fp32, (B, dim) each; every elementwise stage below is ONE fused pass with fp32 math
    # (the reference runs them as chains of fp32 torch ops, modeling_wan.py:201-238)
    shift_msa, scale_msa, gate_msa, c_shift_msa, c_scale_msa, c_gate_msa = (
        t.contiguous() for t in (self.scale_shift_table + temb.float()).unbind(dim=1))

    norm_hidden_states = ops.ln_modulate(hidden_states, None, None, scale_msa, shift_msa, self.norm1.eps)
    if not use_original_attn and routing_score is None:
        routing_score = self.router(temb_before_proj)
    kwargs = dict(self_attention_kwargs or {})
    if branch is not None:
        kwargs["branch"] = branch
    if type(self.attn1.processor) is WanAttnProcessor2_0:     # dense baseline: the base processor takes no routing kwargs
        attn_output = self.attn1(hidden_states=norm_hidden_states, rotary_emb=rotary_emb)
    else:
        attn_output = self.attn1(hidden_states=norm_hidden_states, rotary_emb=rotary_emb, routing_score=routing_score,
                                 use_original_attn=use_original_attn, **kwargs)
    hidden_states = ops.gate_residual(hidden_states, attn_output, gate_msa)

    norm_hidden_states = ops.ln_modulate(hidden_states, self.norm2.weight, self.norm2.bias, None, None, self.norm2.eps)
    attn_output = self.attn2(hidden_states=norm_hidden_states, encoder_hidden_states=encoder_hidden_states)
    hidden_states = ops.gate_residual(hidden_states, attn_output, None)

    norm_hidden_states = ops.ln_modulate(hidden_states, None, None, c_scale_msa, c_shift_msa, self.norm3.eps)
    ff_output = self.ffn(norm_hidden_states)
    hidden_states = ops.gate_residual(hidden_states, ff_output, c_gate_msa)
    return hidden_states, routing_score


def wan_transformer_3d_routed_forward(self, hidden_states: torch.Tensor, timestep: torch.Tensor,
                                      encoder_hidden_states: torch.Tensor,
                                      encoder_hidden_states_image: Optional[torch.Tensor] = None,
                                      return_dict: bool = True, attention_kwargs=None,
                                      self_attention_kwargs: Optional[Dict[str, Any]] = None,
                                      return_losses: bool = False, reture_hidden_layer_distill_loss: bool = False,
                                      return_routing_scores: bool = False):
    """One DiT forward = one denoise step: signature, dataflow and return structure of modeling_wan.py:38-192
    (``reture_…`` is the reference's spelling).  ``return_dict=False`` gives the reference's 5-tuple
    ``(sample, reg_loss, last_layer_distill_loss, hidden_layer_distill_loss, routing_scores)``.
    ``return_losses`` evaluates the router-training losses of :109-171 in the forward direction (L2 on the
    full-attention score, distillation against the un-routed blocks run without gradient); their backward through the
    attention kernels does not exist yet."""
    batch_size, _, num_frames, height, width = hidden_states.shape
    p_t, p_h, p_w = self.config.patch_size
    ppf, pph, ppw = num_frames // p_t, height // p_h, width // p_w

    rotary_emb = self.rope(hidden_states)
    hidden_states = self.patch_embedding(hidden_states).flatten(2).transpose(1, 2)
    temb, timestep_proj, encoder_hidden_states, encoder_hidden_states_image = self.condition_embedder(
        timestep, encoder_hidden_states, encoder_hidden_states_image)
    timestep_proj = timestep_proj.unflatten(1, (6, -1))
    if encoder_hidden_states_image is not None:           # I2V: [257 image tokens | text tokens] (:83-84)
        encoder_hidden_states = torch.concat([encoder_hidden_states_image, encoder_hidden_states], dim=1)

    if SP_STATE.enabled:        # token sharding: S / P contiguous tokens per rank
        s_loc = hidden_states.shape[1] // SP_STATE.sp_size
        hidden_states = hidden_states.narrow(1, SP_STATE.group_local_rank * s_loc, s_loc).contiguous()

    kwargs = dict(self_attention_kwargs or {})
    tau = kwargs.get("tau_sparse")
    eval_mode = isinstance(self.blocks[0].attn1.processor, WanAttnProcessorTripleEval)
    # dense baseline (apply_sp_flashattn_transformer): plain processors, no routers; same token-sharded forward
    dense_baseline = type(self.blocks[0].attn1.processor) is WanAttnProcessor2_0
    # routing of the whole step in one launch: it depends on temb only
    if not dense_baseline and not eval_mode and torch.is_grad_enabled() and any(q.requires_grad for q in self.blocks[0].router.parameters()):
        # the Train processors exist to fit the routers; that needs the backward of the attention kernels
        raise NotImplementedError("vorta_b200: router training is not supported yet (no attention backward); run the "
                                  "Train processors under torch.no_grad()")
    reg_loss = hidden_layer_distill_loss = last_layer_distill_loss = None
    if dense_baseline:
        if return_losses or return_routing_scores:
            raise ValueError("the dense baseline (apply_sp_flashattn_transformer) has no routers: no losses / scores")
        scores = branches = None
    else:
        scores, branches = route_step([b.router for b in self.blocks], temb, tau if eval_mode else None)
    self._vb_last_branches = branches if eval_mode else None      # fp32 decisions of this step, per layer (bench / tests)
    ref_hidden_states = hidden_states.detach().clone() if return_losses else None
    for i, block in enumerate(self.blocks):
        if dense_baseline:
            hidden_states, _ = block(hidden_states, encoder_hidden_states, timestep_proj, rotary_emb,
                                     temb_before_proj=temb, use_original_attn=True)
            continue
        score_i = scores[i].to(temb.dtype)
        hidden_states, _ = block(hidden_states, encoder_hidden_states, timestep_proj, rotary_emb,
                                 temb_before_proj=temb, use_original_attn=False, self_attention_kwargs=kwargs,
                                 routing_score=score_i, branch=branches[i] if eval_mode else None)
        if return_losses:
            with torch.no_grad():                          # reference branch: the un-routed block (:122-128)
                ref_hidden_states, _ = block(ref_hidden_states, encoder_hidden_states, timestep_proj, rotary_emb,
                                             temb_before_proj=temb, use_original_attn=True)
            reg_loss = accumulate_loss(reg_loss, torch.square(score_i[:, :, 0]).mean().float())      # :131-132
            if reture_hidden_layer_distill_loss:
                hidden_layer_distill_loss = accumulate_loss(
                    hidden_layer_distill_loss, F.mse_loss(ref_hidden_states.float(), hidden_states.float()))
    # one device -> host copy for the whole step (the reference copies per block, :119-120)
    routing_scores = list(scores.detach().to(temb.dtype).cpu().unbind(0)) if return_routing_scores else []     # noqa: E501

    shift, scale = ((self.scale_shift_table.float() + temb.float().unsqueeze(1))).unbind(dim=1)
    shift, scale = shift.contiguous(), scale.contiguous()
    hidden_states = self.proj_out(ops.ln_modulate(hidden_states, None, None, scale, shift, self.norm_out.eps))
    if return_losses:
        with torch.no_grad():
            ref_hidden_states = self.proj_out(ops.ln_modulate(ref_hidden_states, None, None, scale, shift,
                                                              self.norm_out.eps))
        last_layer_distill_loss = F.mse_loss(ref_hidden_states.float(), hidden_states.float())      # :165-171
    if SP_STATE.enabled:
        hidden_states = all_gather(hidden_states, dim=1)
    hidden_states = hidden_states.reshape(batch_size, ppf, pph, ppw, p_t, p_h, p_w, -1)
    hidden_states = hidden_states.permute(0, 7, 1, 4, 2, 5, 3, 6)
    output = hidden_states.flatten(6, 7).flatten(4, 5).flatten(2, 3)
    if not return_dict:
        return (output, reg_loss, last_layer_distill_loss, hidden_layer_distill_loss, routing_scores)
    return RoutedTransformerModelOutput(sample=output, reg_loss=reg_loss, last_layer_distill_loss=last_layer_distill_loss,
                                        hidden_layer_distill_loss=hidden_layer_distill_loss,
                                        routing_scores=routing_scores)


def accumulate_loss(current_loss, new_loss):
    """vorta/utils/misc.py:91-92."""
    return new_loss if current_loss is None else current_loss + new_loss


def load_router_checkpoint(checkpoint_file: os.PathLike, model) -> None:
    """Router-only checkpoint ``router.pt`` (reference: vorta/train/checkpoint.py:63-74).  Keys are the transformer's
    own state-dict names, ``blocks.{i}.router.linear.{weight,bias}`` (``transformer_blocks`` / ``single_transformer_blocks``
    for HunyuanVideo); like the reference, entries whose key the model does not have are dropped silently and every
    other parameter keeps its value."""
    state = torch.load(checkpoint_file, map_location="cpu", weights_only=True)
    own = model.state_dict()
    own.update({k: v for k, v in state.items() if k in own})
    model.load_state_dict(own)


def apply_vorta_transformer(model, train_router: bool = False, checkpoint_file: Optional[os.PathLike] = None,
                            attn_processor_kwargs: Optional[Dict[str, Any]] = None,
                            router_dtype: Optional[torch.dtype] = None):
    """Same signature and effect as modeling_wan.py:265-310."""
    processor_cls = WanAttnProcessorTripleTrain if train_router else WanAttnProcessorTripleEval
    model.__class__.forward = wan_transformer_3d_routed_forward
    model.rope.__class__.forward = wan_rope_forward
    param = next(model.parameters())
    dtype = router_dtype or param.dtype
    embedding_dim = model.condition_embedder.time_proj.in_features
    attn_processor_kwargs = dict(attn_processor_kwargs or {})
    attn_processor_kwargs.update(check_input=True)
    for block in model.blocks:
        if not hasattr(block, "router"):
            block.router = Router(embedding_dim=embedding_dim, heads=block.attn1.heads, num_experts=3).to(
                device=param.device, dtype=dtype)
        if train_router:
            block.router.requires_grad_(True)
        block.attn1.set_processor(processor_cls(**attn_processor_kwargs))
        block.attn2.set_processor(WanAttnProcessor2_0())
        attn_processor_kwargs.update(check_input=False)       # only the first block checks (modeling_wan.py:303)
        block.__class__.forward = wan_block_routed_forward
    if checkpoint_file is not None:
        load_router_checkpoint(checkpoint_file, model)
    return model


def apply_sp_flashattn_transformer(model):
    """Baseline-only variant (modeling_wan.py:313-323): dense attention processors, no routing.  The reference leaves
    the model forward alone because its PIPELINE shards frames (pipeline_wan.py:120-122); here sharding lives in the
    model forward (tokens, not frames), so the baseline installs the same forward — every rank takes its S / P tokens,
    the processors exchange heads for the full sequence, and the output is all-gathered."""
    model.__class__.forward = wan_transformer_3d_routed_forward
    model.rope.__class__.forward = wan_rope_forward
    for block in model.blocks:
        block.attn1.set_processor(WanAttnProcessor2_0())
        block.attn2.set_processor(WanAttnProcessor2_0())
        block.__class__.forward = wan_block_routed_forward
    return model
