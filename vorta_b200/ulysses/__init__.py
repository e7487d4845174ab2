from .parallel_states import SP_STATE, SequenceParallelState
from .utils import (all_gather, all_to_all_4D, exchange_out, exchange_qkv, local_heads, pack_heads, shrink_dim,
                    unpack_heads)
from . import balance
from .balance import balance_heads, branch_costs
