from .router import Router, route_step
from .utils import (hunyuan_pixel2token, prepare_hunyuan_self_attn_kwargs,
                    prepare_wan_self_attn_kwargs, wan_pixel2token)
from .modeling_wan import (apply_sp_flashattn_transformer, apply_vorta_transformer, load_router_checkpoint,
                           wan_block_routed_forward, wan_rope_forward, wan_transformer_3d_routed_forward)
from . import modeling_hunyuan, modeling_wan
from .outputs import RoutedTransformerModelOutput, VideoPipelineOutput
from .pipeline import FlowMatchEulerScheduler, hunyuan_denoise, wan_denoise
