"""N>1 path on CPU: world_size-2 `gloo` processes render disjoint sub-mixer subtrees (the oracle stands in
for the GPU renderer behind the same C-ABI), reduce the stereo partial buses onto rank 0 and compare with
the single-process render of the whole graph."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ORACLE_LIB, ROOT

N_SUBTREES = 5
VOICES = 6
FRAMES = 24 * 1024


def build(player, subtree_ids):
    from phonic_b200 import workloads as W
    from phonic_b200.player import FilterEffect
    buf = W.synth_buffer(20000, 44100, seed=1)
    bid = player.upload_buffer(buf, 44100)
    for s in subtree_ids:
        mh = player.add_mixer(None)
        W.add_voice_bank(player, W.VoiceBankSpec(voices=VOICES, voices_per_sampler=3), bid, mh.id,
                         seed_offset=1000 * (s + 1), time_scale=0.05)
        player.add_effect(FilterEffect(0, 2500.0, 0.707), mh.id)


def worker(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from phonic_b200._capi import CApi
    from phonic_b200.distributed import assign_subtrees, reduce_partial_bus, subtree_weight
    from phonic_b200.player import Player
    api = CApi(ORACLE_LIB, "po_")
    weights = [subtree_weight(VOICES, [2])] * N_SUBTREES
    mine = assign_subtrees(weights, world)[rank]
    p = Player(api, 48000)
    build(p, mine)
    part = torch.from_numpy(p.render(FRAMES))
    reduce_partial_bus(part, dst=0)
    if rank == 0:
        np.save(result_path, part.numpy())
        # cfg5's last stage: the main-bus effects on the reduced sum, rank 0 only
        from phonic_b200.distributed import finish_on_main_bus
        from phonic_b200.player import DelayEffect
        np.save(result_path.replace(".npy", "_main.npy"), finish_on_main_bus(api, part.numpy(), 48000, lambda q: q.add_effect(DelayEffect())))
    # the pipelined form: pieces rendered, reduced and run through rank 0's main-bus stage while the next piece renders
    from phonic_b200.distributed import MainBusStage, render_sharded
    from phonic_b200.player import DelayEffect as Delay2
    p = Player(api, 48000)
    build(p, mine)
    bus, out = torch.zeros(FRAMES, 2), torch.zeros(FRAMES, 2)
    stage = MainBusStage(api, 48000, lambda q: q.add_effect(Delay2())) if rank == 0 else None
    stats = {}
    render_sharded(p, bus, 5 * 1024, stage, out, stats=stats)
    if rank == 0:
        assert stats["pieces"] == 5 and stats["main_bus_ms"] > 0.0
        np.save(result_path.replace(".npy", "_pipelined.npy"), out.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_assign_subtrees_is_the_reference_heuristic():
    from phonic_b200.distributed import assign_subtrees
    # heaviest first, each to the lightest bin, first minimum wins
    assert assign_subtrees([5, 3, 3, 2, 2, 1], 2) == [[0, 3, 5], [1, 2, 4]]
    assert assign_subtrees([1, 1, 1, 1], 4) == [[0], [1], [2], [3]]
    bins = assign_subtrees(list(range(1, 65)), 8)
    assert sorted(i for b in bins for i in b) == list(range(64))
    loads = [sum(i + 1 for i in b) for b in bins]
    assert max(loads) - min(loads) <= 8


def test_two_rank_reduce_matches_single_process_render(tmp_path, oracle_api):
    from phonic_b200.player import Player
    port = 29500 + (os.getpid() % 2000)
    result = str(tmp_path / "reduced.npy")
    mp.spawn(worker, args=(2, port, result), nprocs=2, join=True)
    reduced = np.load(result)
    p = Player(oracle_api, 48000)
    # single process: all subtrees in the order the two ranks hold them
    build(p, list(range(N_SUBTREES)))
    full = p.render(FRAMES)
    assert float(np.abs(full).max()) > 0.05
    # the reduce changes the f32 summation order across subtrees: <= 1e-6, not bit-exact
    assert float(np.abs(reduced - full).max()) <= 1e-6
    # the same graph with a Delay on the main bus, rendered in one process, against rank 0's two-stage result
    from phonic_b200.player import DelayEffect
    p2 = Player(oracle_api, 48000)
    build(p2, list(range(N_SUBTREES)))
    p2.add_effect(DelayEffect())
    full_main = p2.render(FRAMES)
    two_stage = np.load(result.replace(".npy", "_main.npy"))
    assert float(np.abs(full_main).max()) > 0.05
    assert float(np.abs(two_stage - full_main).max()) <= 2e-6
    # ... and the pipelined render (per-piece reduce, main-bus stage fed through pb200_set_main_input) gives the same bytes
    # as the two-stage one: same partial sums, same reduce, same chain on the same 1024-frame chunks
    pipelined = np.load(result.replace(".npy", "_pipelined.npy"))
    assert np.array_equal(pipelined, two_stage)


def test_main_bus_stage_alone_is_bit_exact(oracle_api):
    """Without a reduce in between (one rank), rendering the sub-mixers first and the main-bus chain second gives the bytes
    of the single-graph render: the bus re-enters as an equal-rate, unity-gain source (a copy) and the main chain sees the
    same 1024-frame chunks."""
    from phonic_b200.distributed import finish_on_main_bus
    from phonic_b200.player import DelayEffect, FilterEffect, Player
    p = Player(oracle_api, 48000)
    build(p, [0, 1, 2])
    bus = p.render(FRAMES)
    p.close()

    def chain(q):
        q.add_effect(FilterEffect(0, 3000.0, 0.707))
        q.add_effect(DelayEffect())
    two_stage = finish_on_main_bus(oracle_api, bus, 48000, chain)
    p = Player(oracle_api, 48000)
    build(p, [0, 1, 2])
    chain(p)
    full = p.render(FRAMES)
    p.close()
    assert float(np.abs(full).max()) > 0.05
    assert np.array_equal(two_stage, full)


def test_main_input_stage_is_bit_exact(oracle_api):
    """pb200_set_main_input: the main mixer fed with the sub-mixers' summed bus (rendered by another renderer, in pieces)
    gives the bytes of the single-graph render, and so does rendering the shard in pieces."""
    import torch
    from phonic_b200.distributed import MainBusStage, render_sharded
    from phonic_b200.player import DelayEffect, FilterEffect, Player

    def chain(q):
        q.add_effect(FilterEffect(0, 3000.0, 0.707))
        q.add_effect(DelayEffect())
    p = Player(oracle_api, 48000)
    build(p, [0, 1, 2])
    bus, out = torch.zeros(FRAMES, 2), torch.zeros(FRAMES, 2)
    stage = MainBusStage(oracle_api, 48000, chain)
    stats = {}
    render_sharded(p, bus, 7 * 1024, stage, out, stats=stats)
    p.close()
    stage.close()
    assert stats["pieces"] == 4
    p = Player(oracle_api, 48000)
    build(p, [0, 1, 2])
    chain(p)
    full = p.render(FRAMES)
    p.close()
    assert float(np.abs(full).max()) > 0.05
    assert np.array_equal(out.numpy(), full)


def test_main_input_serves_one_render_call(oracle_api):
    from phonic_b200.player import Player
    p = Player(oracle_api, 48000)
    bus = (np.random.default_rng(3).standard_normal((2048, 2)) * 0.1).astype(np.float32)
    p.set_main_input(bus.ctypes.data, 2048)
    assert np.array_equal(p.render(2048), bus)
    assert not p.render(1024).any()          # detached again: the empty main mixer writes nothing
    with pytest.raises(Exception):
        p.set_main_input(bus.ctypes.data, 1000)  # not a multiple of the block
    p.close()


def test_load_aware_partition_of_cfg5():
    """Rank 0 carries the main-bus chain as a preload: with enough ranks it ends up holding the main bus only."""
    import bench
    from phonic_b200.distributed import assign_subtrees
    assert bench.cfg5_submixers(1) == [64]
    for world in (2, 4, 8):
        per = bench.cfg5_submixers(world)
        assert sum(per) == 64 * world and len(per) == world
        assert per[0] <= min(per[1:]), per                  # rank 0 never holds more sub-mixers than anybody else
        assert max(per[1:]) - min(per[1:]) <= 1, per        # the others are balanced
    assert bench.cfg5_submixers(8)[0] == 0
    # the preload is only a starting load: the packing itself stays the reference's heuristic
    assert assign_subtrees([4, 4, 4, 4], 2, preload=[8, 0]) == [[2], [0, 1, 3]]   # (a tie goes to the first bin, like min_by_key)
    assert assign_subtrees([4, 4, 4, 4], 2, preload=[0, 0]) == assign_subtrees([4, 4, 4, 4], 2)


def test_piece_bounds():
    from phonic_b200.distributed import piece_bounds
    assert piece_bounds(5 * 1024, 2 * 1024) == [(0, 2048), (2048, 2048), (4096, 1024)]
    assert piece_bounds(1024, 65536) == [(0, 1024)]
    with pytest.raises(AssertionError):
        piece_bounds(1000, 1024)


def test_consecutive_sharded_renders_continue_the_stream(oracle_api):
    """Two render_sharded calls on the same shard renderer and main stage == one render of twice the length."""
    import torch
    from phonic_b200.distributed import MainBusStage, render_sharded
    from phonic_b200.player import DelayEffect, Player

    def chain(q):
        q.add_effect(DelayEffect())
    half = FRAMES // 2
    p = Player(oracle_api, 48000)
    build(p, [0, 1])
    stage = MainBusStage(oracle_api, 48000, chain)
    got = []
    for _ in range(2):
        bus, out = torch.zeros(half, 2), torch.zeros(half, 2)
        render_sharded(p, bus, 4 * 1024, stage, out)
        got.append(out.numpy().copy())
    p.close()
    stage.close()
    p = Player(oracle_api, 48000)
    build(p, [0, 1])
    chain(p)
    full = p.render(FRAMES)
    p.close()
    assert np.array_equal(np.concatenate(got), full)
