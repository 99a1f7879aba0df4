"""The sharded render of SURVEY.md 8e on the GPU: pb200_set_main_input (device bus into the main mixer), the pipelined
piece-by-piece render on one rank against the oracle's single-graph render, and -- when the box has two GPUs -- two NCCL
ranks against the same oracle render."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ORACLE_LIB, ROOT
from phonic_b200.player import DelayEffect, FilterEffect, Player, ReverbEffect
from test_distributed_gloo import FRAMES, N_SUBTREES, build

pytestmark = pytest.mark.gpu


def chain(q):
    q.add_effect(FilterEffect(0, 3000.0, 0.707))
    q.add_effect(DelayEffect())
    q.add_effect(ReverbEffect(0.6, 0.35))


def oracle_full(oracle_api, subtrees):
    p = Player(oracle_api, 48000)
    build(p, subtrees)
    chain(p)
    full = p.render(FRAMES)
    p.close()
    return full


def test_main_input_alone_is_a_copy(cuda_api):
    p = Player(cuda_api, 48000)
    bus = (torch.randn(4096, 2, device="cuda") * 0.1).contiguous()
    out = torch.zeros_like(bus)
    p.set_main_input(bus.data_ptr(), 4096)
    p.render_device(out.data_ptr(), 4096)
    assert torch.equal(out, bus)
    p.render_device(out.data_ptr(), 1024)   # detached: the empty main mixer is finished, silence
    assert not out[:1024].any()
    p.close()


def test_main_input_adds_to_the_main_mixers_own_children(cuda_api, oracle_api):
    """A rank that holds subtrees AND the main chain: external bus + its own sub-mixers, then the main effects."""
    outs = []
    rng = np.random.default_rng(5)
    ext = (rng.standard_normal((FRAMES, 2)) * 0.05).astype(np.float32)
    for api in (cuda_api, oracle_api):
        p = Player(api, 48000)
        build(p, [0, 1])
        chain(p)
        if api is cuda_api:
            d = torch.from_numpy(ext).cuda()
            o = torch.zeros(FRAMES, 2, device="cuda")
            p.set_main_input(d.data_ptr(), FRAMES)
            p.render_device(o.data_ptr(), FRAMES)
            outs.append(o.cpu().numpy())
        else:
            p.set_main_input(ext.ctypes.data, FRAMES)
            outs.append(p.render(FRAMES))
        p.close()
    assert float(np.abs(outs[1]).max()) > 0.05
    assert float(np.abs(outs[0] - outs[1]).max()) <= 1e-5


def test_pipelined_render_one_rank_matches_oracle(cuda_api, oracle_api):
    from phonic_b200.distributed import MainBusStage, render_sharded
    p = Player(cuda_api, 48000)
    build(p, list(range(N_SUBTREES)))
    bus = torch.zeros(FRAMES, 2, device="cuda")
    out = torch.zeros(FRAMES, 2, device="cuda")
    stage = MainBusStage(cuda_api, 48000, chain)
    stats = {}
    render_sharded(p, bus, 5 * 1024, stage, out, stats=stats)
    p.close()
    stage.close()
    assert stats["pieces"] == 5 and stats["main_bus_launches"] > 0 and stats["shard_launches"] > 0
    full = oracle_full(oracle_api, list(range(N_SUBTREES)))
    assert float(np.abs(full).max()) > 0.05
    assert float(np.abs(out.cpu().numpy() - full).max()) <= 1e-5


def nccl_worker(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import phonic_b200
    from phonic_b200.distributed import MainBusStage, assign_subtrees, render_sharded, subtree_weight
    api = phonic_b200.load_api()
    mine = assign_subtrees([subtree_weight(6, [2])] * N_SUBTREES, world)[rank]
    p = Player(api, 48000, device_ordinal=rank)
    build(p, mine)
    bus = torch.zeros(FRAMES, 2, device="cuda")
    out = torch.zeros(FRAMES, 2, device="cuda")
    stage = MainBusStage(api, 48000, chain, device_ordinal=rank) if rank == 0 else None
    render_sharded(p, bus, 6 * 1024, stage, out)
    if rank == 0:
        np.save(result_path, out.cpu().numpy())
    p.close()
    # the same render with the pieces pushed into rank 0's IPC staging buses instead of reduced (twice: the flags' generation)
    from phonic_b200.distributed import PeerBus
    peer = PeerBus(api, FRAMES, 4, rank)
    for it in range(2):
        p = Player(api, 48000, device_ordinal=rank)
        build(p, mine)
        stage = MainBusStage(api, 48000, chain, device_ordinal=rank) if rank == 0 else None
        out.zero_()
        render_sharded(p, bus, 6 * 1024, stage, out, peer=peer)
        if rank == 0:
            np.save(result_path.replace(".npy", f"_peer{it}.npy"), out.cpu().numpy())
            stage.close()
        p.close()
    peer.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_nccl_ranks_match_oracle(tmp_path, oracle_api):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    result = str(tmp_path / "nccl.npy")
    mp.spawn(nccl_worker, args=(2, port, result), nprocs=2, join=True)
    full = oracle_full(oracle_api, list(range(N_SUBTREES)))
    got = np.load(result)
    assert float(np.abs(full).max()) > 0.05
    assert float(np.abs(got - full).max()) <= 1e-5
    for it in range(2):
        got = np.load(result.replace(".npy", f"_peer{it}.npy"))
        assert float(np.abs(got - full).max()) <= 1e-5
