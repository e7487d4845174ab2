from .router import Router, route_step
from .utils import (Pixel2TokenFactory, hunyuan_pixel2token, prepare_hunyuan_self_attn_kwargs,
                    prepare_wan_self_attn_kwargs, wan_pixel2token)
