from .wan import WAN_CONFIGS, Attention, WanConfig, WanDiT, WanTransformerBlock
