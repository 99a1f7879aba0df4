"""Multi-GPU plumbing of the offline render: the graph is partitioned by whole direct
sub-mixers of the main mixer, one process per GPU renders its subtrees into a stereo partial bus, and
the partials are combined with ONE reduce per render (NCCL over NVLink on GPUs; any torch.distributed
backend works, `gloo` is used by the CPU tests). Main-bus effects are nonlinear in the sum and run on
rank 0 after the reduce.
"""
from __future__ import annotations

import os
import sys
from typing import List, Sequence


def assign_subtrees(weights: Sequence[int], world_size: int, preload: Sequence[int] | None = None) -> List[List[int]]:
    """Greedy weight bin-packing of sub-mixer subtrees onto ranks -- the heuristic the reference uses
    to spread sub-mixers over its worker threads (WorkerTaskBatcher::update,
    src/source/mixed/submixer/thread_pool.rs:92-121): heaviest first (stable), each to the currently
    lightest bin (first minimum wins, like Iterator::min_by_key). `preload`: load a rank carries before any subtree
    (rank 0's main-bus effect chain, which nobody else can run), in the same units as `weights`."""
    order = sorted(range(len(weights)), key=lambda i: -weights[i])  # Python's sort is stable, like sort_by
    bins: List[List[int]] = [[] for _ in range(world_size)]
    totals = list(preload) if preload is not None else [0] * world_size
    assert len(totals) == world_size
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


class MainBusStage:
    """Rank 0's main mixer in a sharded render: a renderer that holds nothing but the main mixer's own effect chain and
    receives the reduced sub-mixer bus piece by piece through `pb200_set_main_input` (device memory stays on the device:
    no host round trip, no second sample upload). `add_effects(player)` attaches the chain (cfg5: Delay + Reverb)."""

    def __init__(self, api, sample_rate: int, add_effects, device_ordinal: int = -1):
        from .player import Player
        self.player = Player(api, sample_rate, device_ordinal=device_ordinal)
        add_effects(self.player)
        self.device_ms = 0.0
        self.kernel_launches = 0

    def process(self, bus, out):
        """bus, out: torch tensors [n, 2] f32, contiguous, both on the renderer's device (CPU tensors for the oracle)."""
        n = bus.shape[0]
        self.player.set_main_input(bus.data_ptr(), n)
        if out.is_cuda:
            self.player.render_device(out.data_ptr(), n)
        else:
            self.player.render_into(out.numpy())
        st = self.player.last_render_stats()
        self.device_ms += st.device_ms
        self.kernel_launches += int(st.kernel_launches)

    def process_inputs(self, bus_ptrs, out, n: int):
        """Same, the main mixer's input being the sum of several device buses (raw pointers, one per rank, added in order)."""
        self.player.set_main_inputs(bus_ptrs, n)
        self.player.render_device(out.data_ptr(), n)
        st = self.player.last_render_stats()
        self.device_ms += st.device_ms
        self.kernel_launches += int(st.kernel_launches)

    def close(self):
        self.player.close()


class PeerBus:
    """The transport of a sharded render between GPUs of one node without a collective kernel: rank 0 owns one staging
    bus per other rank (plus one 'landed' flag per rank and piece), allocated with pb200_device_alloc and shared through
    CUDA IPC; a rank pushes a finished piece with a stream-ordered DMA copy followed by the flag (pb200_push_async). Nothing
    waits for SMs on the sending device (an NCCL reduce kernel queues 5-10 ms behind a busy shard's launches) and rank 0 adds
    the buses in rank order inside its main mixer (pb200_set_main_inputs): a fixed summation order."""

    def __init__(self, api, frames: int, max_pieces: int, device_ordinal: int, group=None):
        import ctypes as C

        import torch.distributed as dist
        self.api, self.frames, self.max_pieces = api, frames, max_pieces
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.group = group
        self.gen = 0
        n_peers = self.world - 1
        self.owned = self.rank == 0
        handles = [None, None]
        if self.owned:
            self.staging, self.flags = C.c_void_p(), C.c_void_p()
            self._check(api.device_alloc(device_ordinal, n_peers * frames * 8, C.byref(self.staging)))
            self._check(api.device_alloc(device_ordinal, max(n_peers * max_pieces, 1) * 4, C.byref(self.flags)))
            hs = [(C.c_ubyte * 64)(), (C.c_ubyte * 64)()]
            self._check(api.ipc_export(self.staging, hs[0]))
            self._check(api.ipc_export(self.flags, hs[1]))
            handles = [bytes(hs[0]), bytes(hs[1])]
        dist.broadcast_object_list(handles, src=0, group=group)
        if not self.owned:
            self.staging, self.flags = C.c_void_p(), C.c_void_p()
            for h, dst in ((handles[0], self.staging), (handles[1], self.flags)):
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                self._check(api.ipc_open(buf, device_ordinal, C.byref(dst)))
        dist.barrier(group)

    @staticmethod
    def _check(code):
        if code != 0:
            raise RuntimeError(f"peer bus: C-ABI error {code}")

    def slot(self, rank: int, off: int) -> int:
        return self.staging.value + ((rank - 1) * self.frames + off) * 8

    def flag(self, rank: int, piece: int) -> int:
        return self.flags.value + ((rank - 1) * self.max_pieces + piece) * 4

    def close(self):
        import torch.distributed as dist
        dist.barrier(self.group)   # nobody still pushes into / reads from the staging memory
        if self.owned:
            self.api.device_free(self.staging)
            self.api.device_free(self.flags)
        else:
            self.api.ipc_close(self.staging)
            self.api.ipc_close(self.flags)


def piece_bounds(frames: int, piece_frames: int):
    """[(offset, length)] of the pieces a sharded render is cut into (whole multiples of the 1024-frame block)."""
    assert frames % 1024 == 0 and piece_frames % 1024 == 0 and piece_frames > 0
    return [(o, min(piece_frames, frames - o)) for o in range(0, frames, piece_frames)]


def render_sharded(player, bus, piece_frames: int, main_stage: MainBusStage | None = None, out=None, group=None, stats=None,
                   peer: PeerBus | None = None):
    """One render of a graph partitioned over the ranks of `group`, pipelined piece by piece:

      every rank   renders its sub-mixer subtrees into `bus` (its partial stereo bus) in ONE render call -- the renderer's
                   own pipelining across time blocks stays intact -- while a second host thread follows
                   `pb200_render_progress` and starts the reduce of every finished piece onto rank 0 (asynchronously:
                   NCCL's stream / gloo's thread);
      rank 0 only  a third thread waits for the reduce of piece p and runs the main mixer's own effect chain on it
                   (`main_stage`, nonlinear in the sum, so not shardable) into `out[p]`, while the shards render on.

    With `peer` (GPUs of one node, rank 0 holding a main stage) the pieces travel as DMA pushes into rank 0's staging buses
    instead of a reduce, and rank 0's main mixer adds them in rank order.
    `bus`, `out`: torch [frames, 2] f32 on the rank's device. Without a process group the reduce is skipped (one rank).
    Returns `out` on rank 0 when there is a main stage, else `bus` (the reduced sum on rank 0)."""
    import queue
    import threading
    import time

    import torch
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if distributed else 0
    frames = bus.shape[0]
    pieces = piece_bounds(frames, piece_frames)
    ready: "queue.Queue" = queue.Queue()
    err = []
    wait_s = [0.0]
    works = []
    if peer is not None:
        assert bus.is_cuda and distributed and len(pieces) <= peer.max_pieces and frames == peer.frames
        peer.gen += 1   # (every rank counts the renders: a flag of an earlier render never matches)
    base = player.render_progress()
    render_over = threading.Event()
    trace = [] if os.environ.get("PB200_TRACE_REDUCE") else None
    tl = [] if os.environ.get("PB200_TRACE_PIPE") else None   # (debug) timeline: (what, piece, ms since the call started)
    t_call = time.perf_counter()
    mark = (lambda what, k: tl.append((what, k, round((time.perf_counter() - t_call) * 1e3, 2)))) if tl is not None else (lambda what, k: None)
    has_stage = main_stage is not None and rank == 0
    if has_stage:
        assert out is not None

    def follow():  # every rank: the collectives, in piece order, as the pieces become final
        try:
            if bus.is_cuda:
                torch.cuda.set_device(bus.device)
            for (off, n) in pieces:
                while player.render_progress() - base < off + n:
                    if render_over.is_set() and player.render_progress() - base < off + n:
                        return   # the render call failed; the caller raises
                    time.sleep(0.00005)
                mark("ready", len(works))
                if peer is not None:
                    if rank != 0:
                        player.push_async(peer.slot(rank, off), bus.data_ptr() + off * 8, n * 8, peer.flag(rank, len(works)), peer.gen)
                    work = True
                else:
                    work = dist.reduce(bus[off:off + n], dst=0, op=dist.ReduceOp.SUM, group=group, async_op=True) if distributed else True
                if trace is not None and work is not True:   # (debug: how long each collective takes from its issue on this rank)
                    t0 = time.perf_counter()
                    work.wait()
                    if bus.is_cuda:
                        torch.cuda.current_stream().synchronize()
                    trace.append((time.perf_counter() - t0) * 1e3)
                works.append(work)
                if has_stage:
                    ready.put(work)
        except BaseException as e:
            err.append(e)
        finally:
            if has_stage:
                ready.put(None)

    def consume():  # rank 0: the main mixer's own chain on every reduced piece
        try:
            if bus.is_cuda:
                torch.cuda.set_device(bus.device)
                side = torch.cuda.Stream(device=bus.device)
            for (off, n) in pieces:
                work = ready.get()
                if work is None:
                    return
                t0 = time.perf_counter()
                if work is not True:
                    if bus.is_cuda:
                        with torch.cuda.stream(side):
                            work.wait()          # the side stream waits for the collective; then the host for the stream
                        side.synchronize()
                    else:
                        work.wait()
                if peer is not None:   # every other rank's piece has landed in its staging bus?
                    k = pieces.index((off, n))
                    while True:
                        fl = main_stage.player.peek_u32(peer.flags.value, (peer.world - 1) * peer.max_pieces)
                        if all(fl[(r - 1) * peer.max_pieces + k] == peer.gen for r in range(1, peer.world)):
                            break
                        time.sleep(0.0001)
                wait_s[0] += time.perf_counter() - t0
                mark("landed", pieces.index((off, n)))
                if peer is not None:
                    main_stage.process_inputs([bus.data_ptr() + off * 8] + [peer.slot(r, off) for r in range(1, peer.world)], out[off:off + n], n)
                else:
                    main_stage.process(bus[off:off + n], out[off:off + n])
                mark("main done", pieces.index((off, n)))
        except BaseException as e:  # surfaced by the caller
            err.append(e)

    threads = []
    if distributed or has_stage:
        threads.append(threading.Thread(target=follow, name="piece-reduce"))
    if has_stage:
        threads.append(threading.Thread(target=consume, name="main-bus-stage"))
    for t in threads:
        t.start()
    try:
        if bus.is_cuda:
            player.render_device(bus.data_ptr(), frames)
        else:
            player.render_into(bus.numpy())
        mark("render call returned", -1)
    finally:
        render_over.set()
        for t in threads:
            t.join()
    mark("threads joined", -1)
    for w in works:
        if w is not True:
            w.wait()
    if peer is not None:
        if rank != 0:
            player.push_sync()
        dist.barrier(group)   # rank 0 has consumed every staging piece before anybody pushes the next render's
    if bus.is_cuda:
        torch.cuda.synchronize(bus.device)
    mark("end", -1)
    if tl is not None:
        print(f"[rank {rank}] " + " | ".join(f"{w} {k}: {t}" for (w, k, t) in sorted(tl, key=lambda x: x[2])), file=sys.stderr)
    if err:
        raise err[0]
    if stats is not None:
        st = player.last_render_stats()
        stats["shard_ms"] = st.device_ms
        stats["shard_launches"] = int(st.kernel_launches) + (len(pieces) if distributed and peer is None else 0)
        stats["pieces"] = len(pieces)
        if trace is not None:
            stats["reduce_latency_ms"] = [round(x, 2) for x in trace]
        for k in ("voice_kernel_ms", "skeleton_kernel_ms", "effect_kernel_ms", "voice_frames"):
            stats[k] = getattr(st, k)
        if has_stage:
            stats["main_bus_ms"] = main_stage.device_ms
            stats["main_bus_launches"] = main_stage.kernel_launches
            stats["reduce_wait_ms"] = wait_s[0] * 1e3
    return out if has_stage else bus
