"""Same export list as the reference's ``vorta/attention/__init__.py:1-16``."""
from .coreset_select import (LowresGroupInfo, MatchingResults, get_group_info, pool_sequence_by_similarity,
                             unpool_sequence_by_similarity)
from .hunyuan import (HunyuanVideoFlashAttnProcessor, HunyuanVideoFlashAttnProcessorTripleEval,
                      HunyuanVideoFlashAttnProcessorTripleTrain)
from .sliding_attn import SlidingTileSchedule, create_sliding_tile_attn_mask_func, sliding_tile_flex_attn
from .tile import tile_layout, untile_layout
from .wan import WanAttnProcessor2_0, WanAttnProcessorTripleEval, WanAttnProcessorTripleTrain
