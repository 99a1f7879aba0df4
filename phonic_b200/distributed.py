"""Multi-GPU plumbing of the offline render: the graph is partitioned by whole direct
sub-mixers of the main mixer, one process per GPU renders its subtrees into a stereo partial bus, and
the partials are combined with ONE reduce per render (NCCL over NVLink on GPUs; any torch.distributed
backend works, `gloo` is used by the CPU tests). Main-bus effects are nonlinear in the sum and run on
rank 0 after the reduce.
"""
from __future__ import annotations

from typing import List, Sequence


def assign_subtrees(weights: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy weight bin-packing of sub-mixer subtrees onto ranks -- the heuristic the reference uses
    to spread sub-mixers over its worker threads (WorkerTaskBatcher::update,
    src/source/mixed/submixer/thread_pool.rs:92-121): heaviest first (stable), each to the currently
    lightest bin (first minimum wins, like Iterator::min_by_key)."""
    order = sorted(range(len(weights)), key=lambda i: -weights[i])  # Python's sort is stable, like sort_by
    bins: List[List[int]] = [[] for _ in range(world_size)]
    totals = [0] * world_size
    for i in order:
        b = min(range(world_size), key=lambda k: totals[k])
        totals[b] += weights[i]
        bins[b].append(i)
    return bins


def subtree_weight(n_voices: int, effect_weights: Sequence[int] = ()) -> int:
    """MixedSource::weight (src/source/mixed.rs:734-748): active sampler voices + effect weights.
    Effect::weight(): Filter 2, Eq5 3, Chorus 3, Delay 3, Compressor 4, Reverb 5."""
    return max(n_voices, 1) + sum(effect_weights)


def reduce_partial_bus(bus, dst: int = 0, group=None):
    """Sum the per-rank stereo partial buses onto `dst` (one collective per render). `bus` is a torch
    tensor [frames, 2] f32 on the rank's device (CUDA for NCCL, CPU for gloo); reduced in place."""
    import torch.distributed as dist
    dist.reduce(bus, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return bus


def finish_on_main_bus(api, bus, sample_rate: int, add_effects, device_ordinal: int = -1, stats=None):
    """Rank 0's last stage: the reduced stereo bus runs through the MAIN mixer's effect chain
    (nonlinear in the sum, so it cannot be sharded). The bus enters a fresh renderer as a file source at its own rate --
    the resampler's equal-rate bypass, unity gain, centre pan: a bit-exact copy (scene `file_bypass`) -- and
    `add_effects(player)` attaches the main-bus effects (cfg5: Delay + Reverb). `bus` is a host array [frames, 2] f32
    (a multiple of the 1024-frame block); returns the final audio [frames, 2]. `stats`: optional dict that receives the
    stage's device time."""
    import numpy as np

    from .player import FilePlaybackOptions, Player
    bus = np.ascontiguousarray(bus, dtype=np.float32)
    frames = bus.shape[0]
    p = Player(api, sample_rate, device_ordinal=device_ordinal)
    try:
        bid = p.upload_buffer(bus, sample_rate)
        p.play_file_source(bid, FilePlaybackOptions())
        add_effects(p)
        out = p.render(frames)
        if stats is not None:
            st = p.last_render_stats()
            stats["device_ms"] = st.device_ms
            stats["kernel_launches"] = int(st.kernel_launches)
            stats["skeleton_kernel_ms"] = st.skeleton_kernel_ms
            stats["voice_kernel_ms"] = st.voice_kernel_ms
            stats["effect_kernel_ms"] = st.effect_kernel_ms
    finally:
        p.close()
    return out
