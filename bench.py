#!/usr/bin/env python
"""bench.py -- voice-samples/s of the offline resample+FX+mix path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|sinc|cfg5shard|cfg5] [--impl reference]

A *step* is one complete offline render of the workload (cfg2: 256 Sampler voices with AHDSR + glide,
cubic 44.1->48 kHz resampling, FilterEffect LP on the bus, "10 s" = 469 WavStream blocks = 480 256
frames). `value` counts voices x output frames per second of DEVICE time with the scene already
resident in HBM (CUDA events inside the renderer, on the streams the kernels are launched on);
`e2e` is the same metric through the public Player API from HOST buffers: sample upload, event
scheduling, graph upload, render, read-back of the WAV data into pinned host memory.

N > 1 (torchrun, one rank per GPU): every rank renders its own voice bank as one sub-mixer subtree
(weak scaling, the path shards by independent subtrees) and the stereo bus partials are summed on
rank 0 with one NCCL reduce per render.

`--impl reference`: the reference's own CPU path cannot be compiled here (no cargo); the arm times
the oracle port (oracle/, scalar C++) on the host cores instead -- cfg2 has a single mixer, so the
reference's SubMixerThreadPool has nothing to distribute and one audio thread is what phonic uses.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SR = 48000
FLOP_PER_CHANNEL_SAMPLE = 25.0   # Hermite 19 + phase 2 + gain/pan/envelope 4
METRIC = "voice-samples/sec (resample+FX+mix, 48 kHz)"


def oracle_api(fast=True):
    from phonic_b200._capi import CApi
    lib = os.path.join(ROOT, "oracle", "_build", "libphonic_oracle_fast.so" if fast else "libphonic_oracle.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return CApi(lib, "po_")


def workload_spec(name):
    from phonic_b200 import workloads as W
    if name == "cfg2":
        return dict(voices=256, n_mixers=0, seconds=10, desc="cfg2: 256 Sampler voices (AHDSR + glide), cubic 44.1->48k, "
                    "FilterEffect LP 2 kHz on the bus, 10 s (480256 frames)")
    if name == "cfg5shard":
        return dict(voices=8192, n_mixers=64, seconds=10, desc="cfg5 per-GPU shard: 8192 voices in 64 sub-mixer subtrees, "
                    "10 s (480256 frames)")
    if name == "cfg3":
        return dict(voices=4096, n_mixers=64, seconds=60, effects="cfg3", time_scale=6.0,
                    desc="cfg3: 4096 voices over 64 sub-mixers (8 Samplers x 8 voices each), Eq5 + Compressor + Chorus per "
                    "sub-mixer, 60 s (2880512 frames); cfg2 voice events stretched x6 so notes span the render")
    if name == "cfg5":
        return dict(voices=8192, n_mixers=64, seconds=10, main_bus=True,
                    desc="cfg5, one GPU's share per rank: 8192 voices in 64 sub-mixer subtrees, NCCL reduce of the stereo bus, "
                    "Delay + Reverb on the main bus (rank 0, after the reduce), 10 s (480256 frames)")
    if name == "cfg4":
        return dict(voices=160, n_mixers=0, seconds=10, desc="cfg4: granular, 160 voices x 100 grains/s x 100 ms Hann grains "
                    "(16k grains/s) from a 362835-frame mono buffer, AHDSR, 10 s (480256 frames)")
    if name == "sinc":
        return dict(voices=1024, n_mixers=0, seconds=10, desc="sinc micro-benchmark: 1024 HighQuality (rubato sinc, 256 taps x 4 "
                    "sub-phases) mono file sources 44.1->48k, 10 s (480256 frames)")
    raise SystemExit(f"unknown workload {name}")


_BUFFER = None


def sample_buffer():
    """The synthetic sample data (host memory), synthesised once: it is the *input* of the path."""
    global _BUFFER
    if _BUFFER is None:
        from phonic_b200 import workloads as W
        _BUFFER = W.synth_buffer(int(4.0 * 44100), 44100, seed=1)
    return _BUFFER


# cfg5 at N > 1: rank 0 also runs the main mixer's Delay + Reverb, a serial chain nobody else can take (49 ms for a 10 s
# render, against 37 ms for a 64-sub-mixer shard: DESIGN.md 5). The sub-mixers are spread with the reference's greedy weight
# heuristic, rank 0 starting with the chain's cost in voice units -- at 8 ranks it ends up holding the main bus alone.
MAIN_CHAIN_VOICE_UNITS = int(8192 * 49.0 / 37.0)


def cfg5_submixers(world):
    """Sub-mixers per rank of the `world`-rank cfg5 run (world x 64 sub-mixers of 128 voices in total)."""
    from phonic_b200.distributed import assign_subtrees
    if world == 1:
        return [64]
    bins = assign_subtrees([128] * (64 * world), world, preload=[MAIN_CHAIN_VOICE_UNITS] + [0] * (world - 1))
    return [len(b) for b in bins]


_CFG4_BUFFER = None


def cfg4_buffer():
    """cfg4's synthetic pad-shaped sample (host memory), synthesised once like `sample_buffer`: input data, not the path."""
    global _CFG4_BUFFER
    if _CFG4_BUFFER is None:
        from phonic_b200 import workloads as W
        _CFG4_BUFFER = W.synth_buffer(362835, 48000, seed=4)
    return _CFG4_BUFFER


def build_scene(player, name, rank=0, as_subtree=False, world=1):
    """Builds the workload on `player` from the host sample buffer (upload + graph + events)."""
    from phonic_b200 import workloads as W
    from phonic_b200.player import FilterEffect
    spec = workload_spec(name)
    buf = sample_buffer()
    if name == "cfg5" and world > 1:
        n_mine = cfg5_submixers(world)[rank]
        if n_mine:
            W.build_subtrees(player, n_mine, spec["voices"] // spec["n_mixers"], W.VoiceBankSpec(), seed_base=100000 * rank, buffer=buf, fast=True)
        return n_mine * (spec["voices"] // spec["n_mixers"])
    if name == "cfg2":
        if not as_subtree:
            W.build_cfg2(player, W.VoiceBankSpec(voices=spec["voices"]), buffer=buf, fast=True)
        else:  # one GPU's shard of an N-GPU graph: the bank lives on a sub-mixer of the main mixer
            bid = player.upload_buffer(buf, 44100)
            mh = player.add_mixer(None)
            W.add_voice_bank_fast(player, W.VoiceBankSpec(voices=spec["voices"]), bid, mh.id, seed_offset=7919 * rank)
            player.add_effect(FilterEffect(0, 2000.0, 0.707), mh.id)
    elif name == "cfg4":
        W.build_cfg4(player, spec["voices"], buffer=cfg4_buffer())
    elif name == "sinc":
        W.build_sinc_bank(player, spec["voices"], buffer=buf)
    else:
        W.build_subtrees(player, spec["n_mixers"], spec["voices"] // spec["n_mixers"], W.VoiceBankSpec(),
                         effects=spec.get("effects", "none"), time_scale=spec.get("time_scale", 1.0),
                         seed_base=100000 * rank, buffer=buf, fast=True)
    return spec["voices"]


class ClockSampler:
    """SM clock + clock-event (throttle) reasons during the timed region (B200_PROFILING.md's clocks line). Sampled through
    NVML in this process every 50 ms -- a looping `nvidia-smi --query-gpu` holds driver locks for milliseconds per sample,
    which a step made of many short launches and host synchronisations feels (cfg5: 88 -> 110 ms); nvidia-smi is the
    fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []      # nvidia-smi fallback
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc = None
        self.nvml = None
        self.stop_flag = threading.Event()
        self.t = None
        self.period = float(os.environ.get("PB200_CLOCKS_PERIOD", "0.05"))

    def _nvml_loop(self):
        n = self.nvml
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            h = n.nvmlDeviceGetHandleByIndex(phys)
            mx = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
            names = {n.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     n.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", n.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag.is_set():
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                self.mx.append(mx)
                bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
                for b, name in names.items():
                    if bits & b:
                        self.reasons.add(name)
                self.stop_flag.wait(self.period)
        except Exception as e:  # keep what was sampled
            self.reasons.add(f"nvml sampling stopped: {e!r}"[:80])

    def start(self):
        if os.environ.get("PB200_NO_CLOCKS"):   # (debug: how much the sampling itself costs)
            return
        try:
            if os.environ.get("PB200_CLOCKS") == "smi":
                raise RuntimeError("nvidia-smi requested")
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.t.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def fp32_fma_peak_tflops(torch, device):
    """FP32 FMA peak measured live: a dependent-chain-free torch.addcmul loop would be HBM bound, so use
    the arithmetic peak 148 SM x 128 lanes x 2 flop x SM clock(max) and say so (derived, not measured)."""
    props = torch.cuda.get_device_properties(device)
    clock_hz = 1.965e9
    try:
        out = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits", "-i",
                                       str(device.index or 0)], text=True)
        clock_hz = float(out.strip().splitlines()[0]) * 1e6
    except Exception:
        pass
    return props.multi_processor_count * 128 * 2 * clock_hz / 1e12


def build_cpu_sample(p, name, rank=0, as_subtree=False):
    """The bounded sample of `name` the host-side arms render (same per-voice events, fewer voices where the full
    workload would take minutes on one core). Returns (voices in the sample, description)."""
    from phonic_b200 import workloads as W
    spec = workload_spec(name)
    if name == "cfg2":
        build_scene(p, name, rank=rank, as_subtree=as_subtree)
        return spec["voices"], f"full {name} render"
    if name == "cfg4":
        W.build_cfg4(p, 32, buffer=cfg4_buffer())
        return 32, "cfg4 with 32 of 160 voices (same per-voice events)"
    if name == "sinc":
        W.build_sinc_bank(p, 16, buffer=sample_buffer())
        return 16, "sinc bank with 16 of 1024 voices"
    if name == "cfg3":
        W.build_subtrees(p, 2, 64, W.VoiceBankSpec(), effects="cfg3", time_scale=spec["time_scale"], seed_base=100000 * rank,
                         buffer=sample_buffer(), fast=True)
        return 128, "cfg3 with 128 of 4096 voices (2 of 64 sub-mixers, same per-voice events and effect chains)"
    W.build_subtrees(p, 4, 128, W.VoiceBankSpec(), effects="none", seed_base=100000 * rank, buffer=sample_buffer(), fast=True)
    if spec.get("main_bus") and not as_subtree:
        W.add_main_bus_sends(p)
    return 512, f"{name} with 512 of {spec['voices']} voices (4 sub-mixers x 128, same per-voice events)"


def cpu_baseline(name, seconds_budget=25.0):
    """Oracle port on the host cores, rank 0 only, bounded sample of the same workload."""
    from phonic_b200 import workloads as W
    from phonic_b200.player import Player
    api = oracle_api(fast=True)
    spec = workload_spec(name)
    frames = W.frames_for(spec["seconds"], SR)
    p = Player(api, SR)
    v, sample = build_cpu_sample(p, name)
    t0 = time.perf_counter()
    out = np.zeros((frames, 2), np.float32)
    p.render_into(out)
    dt = time.perf_counter() - t0
    p.close()
    return {"value": v * frames / dt, "unit": "voice-samples/s", "cores": 1, "kind": "port", "sample": sample + ", once",
            "seconds": dt}, out


# parity bars of tests/ (DESIGN.md §2): max |gpu - oracle| per workload; feedback effects are judged on the rms floor
PARITY_BARS = {"cfg2": 2.5e-7, "cfg3": 1e-5, "cfg4": 0.0, "sinc": 1e-5, "cfg5shard": 0.0, "cfg5": 1e-4}


def check_parity(api, name, device_ordinal, ref):
    """Renders the scene the oracle just rendered (the cpu_baseline sample: the FULL workload for cfg2) on the GPU through
    the same C-ABI and asserts the parity bar. Outside every timed region; the oracle is the checker, not the product."""
    from phonic_b200.player import Player
    p = Player(api, SR, device_ordinal=device_ordinal)
    build_cpu_sample(p, name)
    gpu = np.zeros_like(ref)
    p.render_into(gpu)
    p.close()
    d = gpu.astype(np.float64) - ref
    err = float(np.abs(d).max())
    rms = float(np.sqrt(np.mean(d ** 2)))
    rms_db = 20.0 * np.log10(max(rms, 1e-30))
    bar = PARITY_BARS[name]
    ok = err <= bar and (name != "cfg5" or rms_db < -90.0) and float(np.abs(ref).max()) > 1e-3
    if not ok:
        raise SystemExit(f"bench.py: PARITY FAILED on {name}: max|gpu-oracle| = {err:.3e} (bar {bar:g}), rms {rms_db:.1f} dBFS")
    return {"parity_max_abs_err": err, "parity_rms_dbfs": rms_db, "parity_bar": bar,
            "parity_scene": "the cpu_baseline sample rendered on both sides (cfg2: the full 256-voice, 480256-frame workload)"}


def run_reference(args):
    """Reference arm: the reference's CPU path (oracle port, Rust toolchain absent) on the host cores.
    N=1: cfg2 has one mixer -> phonic renders it on one audio thread. N>1: the arm's config is N voice banks
    on N direct sub-mixers of the main mixer, which phonic's SubMixerThreadPool renders on N worker threads
    (one sub-mixer per worker, main thread sums: src/source/mixed.rs:505-537) -- emulated with N host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import threading
    from phonic_b200 import workloads as W
    from phonic_b200.player import Player
    api = oracle_api(fast=True)
    spec = workload_spec(args.workload)
    frames = W.frames_for(spec["seconds"], SR)
    voices = spec["voices"]
    world = max(1, args.gpus)
    sample_buffer()
    if args.workload == "cfg4":
        cfg4_buffer()
    times = []
    v_per, sample = voices, ""
    for i in range(args.warmup + args.steps):
        players, outs = [], []
        for r in range(world):
            p = Player(api, SR)
            v_per, sample = build_cpu_sample(p, args.workload, rank=r, as_subtree=world > 1)
            players.append(p)
            outs.append(np.zeros((frames, 2), np.float32))
        t0 = time.perf_counter()
        if world == 1:
            players[0].render_into(outs[0])
        else:
            ths = [threading.Thread(target=players[r].render_into, args=(outs[r],)) for r in range(world)]
            [t.start() for t in ths]
            [t.join() for t in ths]
            total = outs[0]
            for r in range(1, world):
                total += outs[r]
            if spec.get("main_bus"):  # the main mixer's effects run on the sum, on the audio thread
                from phonic_b200.distributed import finish_on_main_bus
                total = finish_on_main_bus(api, total, SR, W.add_main_bus_sends)
        dt = time.perf_counter() - t0
        for p in players:
            p.close()
        if i >= args.warmup:
            times.append(dt)
    total_t = sum(times)
    value = world * v_per * frames * len(times) / total_t
    sample = sample + " per step"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "voice-samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total_t / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": spec["desc"] + (f"; x{world} banks on {world} sub-mixers" if world > 1 else ""),
                       "note": "oracle port of the reference's CPU path (Rust toolchain absent); one host thread per direct "
                               "sub-mixer like SubMixerThreadPool, a single audio thread when the graph has no sub-mixers"},
            "cpu_baseline": {"value": value, "unit": "voice-samples/s", "cores": world, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "voice-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="cfg5 at N > 1: pieces travel as DMA pushes into rank 0's IPC staging buses (peer) or as NCCL reduces (nccl)")
    ap.add_argument("--piece-frames", type=int, default=65536,
                    help="cfg5: frames per piece of the pipelined sharded render (shards render piece p+1 while rank 0's main-bus stage runs p)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import phonic_b200
    from phonic_b200 import workloads as W
    from phonic_b200.distributed import MainBusStage, PeerBus, piece_bounds, reduce_partial_bus, render_sharded
    from phonic_b200.player import Player

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: phonic_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    api = phonic_b200.load_api()
    spec = workload_spec(args.workload)
    frames = W.frames_for(spec["seconds"], SR)
    voices = spec["voices"]
    n_total = args.warmup + args.steps
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)  # > 126 MB L2

    sample_buffer()  # synthesise the input data before anything is timed
    if args.workload == "cfg4":
        cfg4_buffer()

    # ---- device-resident arm: scenes built and uploaded before the timed region -------------------------------
    main_bus = bool(spec.get("main_bus"))
    players, stages = [], []
    for i in range(n_total):
        p = Player(api, SR, device_ordinal=local_rank)
        build_scene(p, args.workload, rank=rank, as_subtree=world > 1 or main_bus, world=world)
        players.append(p)
        # cfg5: rank 0 also owns the main mixer's Delay + Reverb (not shardable); its renderer takes the reduced bus on the device
        stages.append(MainBusStage(api, SR, W.add_main_bus_sends, device_ordinal=local_rank) if main_bus and rank == 0 else None)
    out_dev = torch.zeros(frames, 2, dtype=torch.float32, device=device)
    bus_dev = torch.zeros(frames, 2, dtype=torch.float32, device=device) if main_bus else out_dev
    peer = None
    if main_bus and world > 1 and args.transport == "peer":
        peer = PeerBus(api, frames, len(piece_bounds(frames, args.piece_frames)), local_rank)
    clocks = ClockSampler(local_rank)
    dev_ms, wall_ms, voice_ms, skel_ms, fx_ms, launches, vframes = [], [], [], [], [], 0, 0
    sinc_ms, grain_ms, sinc_frames, grain_samples = [], [], 0, 0
    red_ms_all, main_ms_all, shard_ms_all, wait_ms_all = [], [], [], []
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for i, p in enumerate(players):
        if i == args.warmup:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            clocks.start()
        flush.fill_(float(i))  # evict L2 between steps
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        red_ms = main_ms = 0.0
        extra_launches = 0
        if main_bus:
            # pipelined: piece p+1 renders on every rank while piece p is reduced (NCCL) and runs through rank 0's main-bus
            # chain. The step's device time is the span of the whole pipeline: one pair of events around it.
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            e0.record()
            pst = {}
            render_sharded(p, bus_dev, args.piece_frames, stages[i], out_dev, stats=pst, peer=peer)
            e1.record()
            torch.cuda.synchronize()
            step_ms = e0.elapsed_time(e1)
            if "reduce_latency_ms" in pst:
                print(f"[rank {rank}] step {i} reduce latencies {pst['reduce_latency_ms']} shard {pst['shard_ms']:.1f}", file=sys.stderr)
            extra_launches = pst["shard_launches"] + pst.get("main_bus_launches", 0)
            if i >= args.warmup:
                shard_ms_all.append(pst["shard_ms"]); main_ms_all.append(pst.get("main_bus_ms", 0.0)); wait_ms_all.append(pst.get("reduce_wait_ms", 0.0))
        else:
            p.render_device(out_dev.data_ptr(), frames)
            if world > 1:  # one NCCL reduce of the stereo bus partial per render (rank 0 owns the main bus)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier()  # absorbs the host-side skew between the ranks' step loops: the span below is the collective itself
                e0.record()
                reduce_partial_bus(out_dev, dst=0)
                e1.record()
                torch.cuda.synchronize()
                red_ms = e0.elapsed_time(e1)
                extra_launches = 1
            step_ms = None
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        st = p.last_render_stats()
        if i >= args.warmup:
            dev_ms.append(step_ms if step_ms is not None else st.device_ms + red_ms)
            red_ms_all.append(red_ms)
            wall_ms.append((t1 - t0) * 1e3)
            voice_ms.append(pst["voice_kernel_ms"] if main_bus else st.voice_kernel_ms)
            skel_ms.append(pst["skeleton_kernel_ms"] if main_bus else st.skeleton_kernel_ms)
            fx_ms.append(pst["effect_kernel_ms"] if main_bus else st.effect_kernel_ms)
            launches += (0 if main_bus else int(st.kernel_launches)) + extra_launches
            vframes += int(pst["voice_frames"]) if main_bus else int(st.voice_frames)
            sinc_ms.append(st.sinc_kernel_ms); grain_ms.append(st.grain_kernel_ms)
            sinc_frames += int(st.sinc_frames); grain_samples += int(st.grain_samples)
        if i + 1 == n_total:
            checksum = float(out_dev.abs().sum().item())
        # the renderer's work areas (snapshots, buses) go back to the process-wide pool for the next step's renderer: every
        # step runs in the same, warm device memory instead of a fresh multi-GB range (steps 20: 10.4 -> 7.3 ms per step)
        p.close()
        if stages[i] is not None:
            stages[i].close()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk = clocks.stop()
    # the collective alone (not overlapped), for the record: every piece's reduce back to back behind a barrier
    reduce_alone_ms = None
    if world > 1:
        pieces = piece_bounds(frames, args.piece_frames) if main_bus else [(0, frames)]
        spans = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            e0.record()
            for (off, n) in pieces:
                reduce_partial_bus(bus_dev[off:off + n], dst=0)
            e1.record()
            torch.cuda.synchronize()
            spans.append(e0.elapsed_time(e1))
        reduce_alone_ms = min(spans)

    # ---- end-to-end arm: host buffers in, host WAV data out, every step -------------------------------------------
    out_host = torch.zeros(frames, 2, dtype=torch.float32).pin_memory()
    out_np = out_host.numpy()
    e2e_ms = []
    h2d = d2h = 0
    for i in range(n_total):
        flush.fill_(float(i))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        p = Player(api, SR, device_ordinal=local_rank)
        build_scene(p, args.workload, rank=rank, as_subtree=world > 1 or main_bus, world=world)   # uploads the sample buffer (H2D) + schedules events
        if main_bus:
            stage = MainBusStage(api, SR, W.add_main_bus_sends, device_ordinal=local_rank) if rank == 0 else None
            render_sharded(p, bus_dev, args.piece_frames, stage, out_dev, peer=peer)
            if rank == 0:
                out_host.copy_(out_dev, non_blocking=False)   # the WAV data, D2H
                stage.close()
            torch.cuda.synchronize()
        elif world > 1:
            p.render_device(out_dev.data_ptr(), frames)
            reduce_partial_bus(out_dev, dst=0)
            if rank == 0:
                out_host.copy_(out_dev, non_blocking=False)
            torch.cuda.synchronize()
        else:
            p.render_into(out_np)                                         # graph upload (H2D), kernels, WAV data D2H
        t1 = time.perf_counter()
        p.close()
        if i >= args.warmup:
            e2e_ms.append((t1 - t0) * 1e3)
    buf_bytes = (362835 if args.workload == "cfg4" else int(4.0 * 44100)) * 4
    h2d = buf_bytes + voices * 176 + voices * 3 * 64       # sample buffer + voice state + event records
    d2h = frames * 2 * 4

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    shard_ms_max = allmax(sum(shard_ms_all)) if main_bus else 0.0   # the slowest rank's shard render (device span)
    total_dev_ms = allmax(sum(dev_ms))
    total_e2e_ms = allmax(sum(e2e_ms))
    total_voice_ms = allmax(sum(voice_ms))
    total_vframes = allsum(float(vframes))
    total_launches = allsum(float(launches))
    K = args.steps
    value = world * voices * frames * K / (total_dev_ms / 1e3)
    e2e_value = world * voices * frames * K / (total_e2e_ms / 1e3)

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        fma_peak = fp32_fma_peak_tflops(torch, device)
        passes = {"voice_kernel_ms_per_step": total_voice_ms / K, "skeleton_kernel_ms_per_step": sum(skel_ms) / K,
                  "effect_kernel_ms_per_step": sum(fx_ms) / K,
                  "note": "per-pass spans are event-to-event on their own stream; the three passes overlap across time blocks"}
        fma_src = ("derived: SMs x 128 lanes x 2 x clocks.max.sm (MEASURED_PEAKS.json holds only HBM and bf16 tensor peaks)")
        if args.workload == "sinc":
            # dominant kernel: sinc_kernel. Algorithmic work = output frames x source channels x 4 x 256 FMA (2048 flop)
            flops = sinc_frames * 1 * 2048.0
            t = sum(sinc_ms) / 1e3
            ach = flops / t / 1e12
            smem_bytes = sinc_frames * 258 * (4 + 1) * 4.0   # what any one-coefficient-read-per-output mapping must move
            roofline = {"bound": "fp32_fma", "kernel": "sinc_kernel", "achieved": ach, "peak": fma_peak, "unit": "TFLOP/s",
                        "frac": ach / fma_peak, "traffic": None, "peak_source": fma_src,
                        "sinc_kernel_ms_per_step": sum(sinc_ms) / K, "output_frames_per_step": sinc_frames / K,
                        "shared_memory_gbs": smem_bytes / t / 1e9,
                        "shared_memory_peak_gbs": torch.cuda.get_device_properties(device).multi_processor_count * 128 * 1.965,
                        "shared_memory_note": "the kernel is bound by the shared-memory pipe (128 B/clk/SM): 258 x (4 + channels) "
                                              "words per output against 258 x 4 x channels FFMA caps a mono source at 20 % of the FP32 pipe",
                        **passes}
        elif args.workload == "cfg4":
            # dominant kernel: grain_kernel. Algorithmic traffic = 8 B contribution written + 8 B read back per grain sample
            byts = grain_samples * 16.0
            t = sum(grain_ms) / 1e3
            ach = byts / t / 1e9
            roofline = {"bound": "hbm", "kernel": "grain_kernel", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                        "grain_kernel_ms_per_step": sum(grain_ms) / K, "grain_samples_per_step": grain_samples / K,
                        "note2": "one thread per grain runs the grain's serial f64 recurrences: latency-bound today", **passes}
        else:
            # The voice path is three passes that overlap across time blocks; the step is as fast as the pass with the longest
            # span, so THAT is the kernel the roofline names. Algorithmic work of the path = active voice-frames x 2 channels
            # x 25 flop, FMA-pipe bound because the voices share one L2-resident sample buffer.
            flops = (total_vframes / world) * 2 * FLOP_PER_CHANNEL_SAMPLE
            spans = {"skeleton_kernel": sum(skel_ms), "replay_kernel": total_voice_ms, "mix_fx_kernel": sum(fx_ms)}
            if main_bus:
                spans["mix_fx_kernel (main-bus Delay + Reverb, rank 0)"] = sum(main_ms_all)
            dominant = max(spans, key=lambda k: spans[k])
            ach = flops / (spans[dominant] / 1e3) / 1e12
            step_ach = flops / (total_dev_ms / 1e3) / 1e12
            replay_ach = flops / (total_voice_ms / 1e3) / 1e12
            roofline = {"bound": "fp32_fma", "kernel": dominant, "achieved": ach, "peak": fma_peak, "unit": "TFLOP/s",
                        "frac": ach / fma_peak, "traffic": None,
                        "peak_source": fma_src + "; voices share one L2-resident buffer so the path is not HBM bound",
                        "achieved_note": "the path's algorithmic flops (active voice-frames x 2 ch x 25) over the span of its slowest "
                                         "pass, the one named in `kernel`; step_frac is the same work over ms_per_step",
                        "step_achieved": step_ach, "step_frac": step_ach / fma_peak,
                        "replay_kernel_achieved": replay_ach, "replay_kernel_frac": replay_ach / fma_peak,
                        "fp32_peak_nominal_tflops": 74.45, "fp32_peak_used_tflops": fma_peak,
                        "hbm_achieved_gbs": (total_vframes / world) * 8.0 / (spans[dominant] / 1e3) / 1e9,
                        "hbm_peak_gbs": peaks["hbm_gbs"], "hbm_peak_kind": peak_kind, **passes}
        # measured DRAM traffic of the kernel the roofline names (one ncu --set full capture, committed under profiles/)
        try:
            tr = (json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload) or {}).get("kernels", {})
            ent = tr.get(roofline["kernel"].split(" ")[0])
            if ent and ent.get("bytes_per_launch"):
                roofline["traffic"] = ent["bytes_per_launch"]
                roofline["traffic_unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)" + (
                    f", one time block of {ent['block_frames']} frames" if ent.get("block_frames") else "")
                roofline["traffic_source"] = ent["source"]
                if ent.get("note"):
                    roofline["traffic_note"] = ent["note"]
        except (OSError, ValueError):
            pass
        line = {"metric": METRIC, "value": value, "unit": "voice-samples/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
                "ms_per_step": total_dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": spec["desc"] + (f"; x{world} ranks, one subtree per GPU, NCCL reduce of the stereo bus" if world > 1 else ""),
                           "l2": "flushed between steps (256 MiB fill)", "timing": "CUDA events around the whole pipeline" if main_bus else "CUDA events on the renderer's own streams",
                           "wall_ms_per_step": sum(wall_ms) / K, "checksum": checksum},
                "e2e": {"value": e2e_value, "unit": "voice-samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": total_e2e_ms / K},
                "gpu_launches": int(total_launches), "clocks": clk, "roofline": roofline}
        if world > 1 or main_bus:
            line["multi_gpu"] = {"reduce_ms": reduce_alone_ms, "reduce_note": "every piece's NCCL reduce back to back behind a barrier, "
                                 "measured after the timed steps (inside them the reduces overlap the next piece's render)",
                                 "reduce_in_step_ms": sum(red_ms_all) / K if not main_bus else None,
                                 "pieces": len(piece_bounds(frames, args.piece_frames)) if main_bus else 1,
                                 "transport": ("peer: DMA pushes into rank 0's IPC staging buses, summed in rank order by its main mixer"
                                               if peer is not None else "nccl reduce")}
            if main_bus:
                line["multi_gpu"].update({"shard_ms": sum(shard_ms_all) / K, "shard_ms_slowest_rank": shard_ms_max / K,
                                          "submixers_per_rank": cfg5_submixers(world), "main_bus_ms": sum(main_ms_all) / K,
                                          "reduce_wait_ms": sum(wait_ms_all) / K,
                                          "note": "ms_per_step is the span of the whole pipeline (events around it): shard pieces, "
                                                  "per-piece reduce, rank 0's main-bus chain on piece p while p+1 renders"})
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"], ref_out = cpu_baseline(args.workload)
            line.update(check_parity(api, args.workload, local_rank, ref_out))
        elif world > 1:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if peer is not None:
        peer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
