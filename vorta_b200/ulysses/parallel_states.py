"""Sequence-parallel process-group state — interface of the reference's ``vorta/ulysses/parallel_states.py:7-75``.

One process per GPU; ranks of an SP group are contiguous blocks of ``sp_size`` ranks (parallel_states.py:55-72).
Differences from the reference, by design: the sequence is sharded by TOKENS (S / P
contiguous tokens per rank), not frames, so grids whose frame count is not divisible by P work, and the N-GPU
result is defined to equal the single-GPU result.
"""
from __future__ import annotations

import os

import torch.distributed as dist


class SequenceParallelState:
    def __init__(self):
        self._enabled = False
        self._sp_size = 1
        self._group_id = 0
        self._group_local_rank = 0
        self._group = None

    @property
    def rank(self) -> int:
        return int(os.getenv("RANK", "0"))

    @property
    def local_rank(self) -> int:
        return int(os.getenv("LOCAL_RANK", "0"))

    @property
    def world_size(self) -> int:
        return int(os.getenv("WORLD_SIZE", "1"))

    @property
    def enabled(self) -> bool:
        return self._enabled

    @property
    def sp_size(self) -> int:
        return self._sp_size

    @property
    def group_id(self) -> int:
        return self._group_id

    @property
    def group_local_rank(self) -> int:
        return self._group_local_rank

    @property
    def group(self):
        return self._group

    @property
    def num_sp_groups(self) -> int:
        return self.world_size // self.sp_size

    def cleanup(self) -> None:
        if dist.is_initialized():
            dist.destroy_process_group()
        self.__init__()

    def setup_sp_group(self, sequence_parallel_size: int) -> None:
        if self.world_size % sequence_parallel_size != 0:
            raise ValueError(f"{self.world_size=} must be divisible by {sequence_parallel_size=}!")
        if sequence_parallel_size > 1:
            self._enabled = True
            self._sp_size = sequence_parallel_size
            self._group_id = self.rank // sequence_parallel_size
            self._group_local_rank = self.rank % sequence_parallel_size
            # every rank must create every group (torch.distributed contract); keep our own
            for gid in range(self.world_size // sequence_parallel_size):
                ranks = list(range(gid * sequence_parallel_size, (gid + 1) * sequence_parallel_size))
                grp = dist.new_group(ranks)
                if gid == self._group_id:
                    self._group = grp
        else:
            self._enabled = False
            self._sp_size = 1
            self._group_id = self.rank
            self._group_local_rank = 0
            self._group = None


SP_STATE = SequenceParallelState()
