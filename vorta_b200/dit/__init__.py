from .wan import WAN_CONFIGS, Attention, WanConfig, WanDiT, WanTransformerBlock
from .hunyuan import HUNYUAN_CONFIGS, HunyuanConfig, HunyuanDiT
