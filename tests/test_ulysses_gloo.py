"""Host-side logic of the Ulysses layer on CPU: world_size-2 gloo ranks (no CUDA kernels are called — the pack /
unpack kernels are covered by the GPU suite; here the collective plumbing, group state and head / token bookkeeping
are checked against the oracle permutation pinned to the reference's own all_to_all_4D)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vorta_oracle as O


def _pack_ref(x, P):
    """test-side restatement of the pack kernel: (S_loc, H, D) -> (P, S_loc, H/P, D)"""
    s_loc, H, D = x.shape
    return x.reshape(s_loc, P, H // P, D).transpose(0, 1).contiguous()


def _unpack_ref(recv):
    P, s_loc, hp, D = recv.shape
    return recv.transpose(0, 1).reshape(s_loc, P * hp, D).contiguous()


def _worker(rank, world, port, ret):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from vorta_b200.ulysses import SP_STATE, all_gather, local_heads, shrink_dim
    from vorta_b200.ulysses import utils as U
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with pytest.raises(ValueError):
            SP_STATE.setup_sp_group(3)                           # parallel_states.py:56-57
        SP_STATE.setup_sp_group(world)
        assert SP_STATE.enabled and SP_STATE.sp_size == world and SP_STATE.group_local_rank == rank
        B, H, S, D = 1, 4, 12, 8
        full = torch.arange(B * H * S * D, dtype=torch.float32).reshape(B, H, S, D)
        s_loc = S // world
        mine = full[:, :, rank * s_loc:(rank + 1) * s_loc]
        # "in" direction: pack (test side) -> product collective -> token-major view
        send = _pack_ref(mine[0].transpose(0, 1).contiguous(), world)
        recv = torch.empty_like(send)
        U._a2a(recv, send)
        got = recv.reshape(world * s_loc, H // world, D).unsqueeze(0).transpose(1, 2)
        # "out" direction: inverse
        send2 = got.transpose(1, 2).reshape(world, s_loc, H // world, D).contiguous()
        recv2 = torch.empty_like(send2)
        U._a2a(recv2, send2)
        back = _unpack_ref(recv2).unsqueeze(0).transpose(1, 2)
        # bookkeeping helpers
        assert torch.equal(shrink_dim(full, 2), mine)
        assert local_heads(list(range(H)), H) == list(range(rank * (H // world), (rank + 1) * (H // world)))
        gathered = all_gather(mine.contiguous(), dim=2)
        ret[rank] = (got.contiguous(), bool(torch.equal(back, mine)), bool(torch.equal(gathered, full)))
    finally:
        SP_STATE.cleanup()


def test_ulysses_exchange_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29541, ret), nprocs=world, join=True)
    B, H, S, D = 1, 4, 12, 8
    full = torch.arange(B * H * S * D, dtype=torch.float32).reshape(B, H, S, D)
    shards = [full[:, :, r * (S // world):(r + 1) * (S // world)].contiguous() for r in range(world)]
    want = O.ulysses_scatter_heads(shards)
    for r in range(world):
        got, roundtrip, gathered = ret[r]
        assert torch.equal(got, want[r])          # N-rank layout == the reference's all_to_all_4D semantics
        assert roundtrip and gathered


def test_sp_state_single_process_defaults():
    from vorta_b200.ulysses import SP_STATE, shrink_dim
    assert not SP_STATE.enabled and SP_STATE.sp_size == 1
    x = torch.arange(10)
    assert shrink_dim(x, 0) is x
