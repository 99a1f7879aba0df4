"""cfg5 on one GPU, taken apart: the shard alone, the shard with the progress follower, the whole pipeline (debug aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.distributed import MainBusStage, render_sharded
from phonic_b200.player import Player
api = phonic_b200.load_api()
frames = W.frames_for(10, 48000)
pf = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
bus = torch.zeros(frames, 2, device="cuda"); out = torch.zeros(frames, 2, device="cuda")
def shard():
    p = Player(api, 48000); bench.build_scene(p, "cfg5", as_subtree=True); return p
extra = []
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
MODES = sys.argv[2].split(",") if len(sys.argv) > 2 else ["alone", "pipeline"]
for mode in MODES:
    if mode == "idle-players":   # what bench.py keeps alive during its device arm
        for _ in range(5):
            extra.append(shard()); extra.append(MainBusStage(api, 48000, W.add_main_bus_sends))
        continue
    if mode == "flush":
        flush.fill_(1.0); torch.cuda.synchronize(); continue
    if mode == "clocks":
        cs = bench.ClockSampler(0); cs.start(); time.sleep(0.3); continue
    p = shard()
    stage = MainBusStage(api, 48000, W.add_main_bus_sends) if mode in ("pipeline", "stage-after") else None
    torch.cuda.synchronize(); t0 = time.perf_counter(); st = {}
    if mode == "alone":
        p.render_device(bus.data_ptr(), frames); st["shard_ms"] = p.last_render_stats().device_ms
    elif mode == "stage-after":
        p.render_device(bus.data_ptr(), frames); st["shard_ms"] = p.last_render_stats().device_ms
        stage.process(bus, out); st["main_bus_ms"] = stage.device_ms
    else:
        render_sharded(p, bus, pf, stage, out, stats=st)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(mode, "wall %.1f" % dt, {k: round(v, 1) for k, v in st.items() if k in ("shard_ms", "main_bus_ms")})
    p.close()
    if stage: stage.close()
if "benchlike" in sys.argv:
    pairs = [(shard(), MainBusStage(api, 48000, W.add_main_bus_sends, device_ordinal=0)) for _ in range(4)]
    for i, (p, stage) in enumerate(pairs):
        flush.fill_(float(i)); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if "noevents" not in sys.argv: e0.record()
        st = {}
        t0 = time.perf_counter()
        render_sharded(p, bus, pf, stage, out, stats=st)
        if "noevents" not in sys.argv: e1.record()
        torch.cuda.synchronize()
        print("benchlike wall %.1f" % ((time.perf_counter() - t0) * 1e3), {k: round(v, 1) for k, v in st.items() if k in ("shard_ms", "main_bus_ms")})
    for p, stage in pairs:
        p.close(); stage.close()
