"""PB200_HOST_PROF: host-side phases of one render call of a bench workload (debug aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import Player
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5shard"
api = phonic_b200.load_api()
spec = bench.workload_spec(wl)
frames = W.frames_for(spec["seconds"], 48000)
out = torch.zeros(frames, 2, device="cuda")
for it in range(3):
    t0 = time.perf_counter(); p = Player(api, 48000); bench.build_scene(p, wl); t1 = time.perf_counter()
    if it == 2: os.environ["PB200_HOST_PROF"] = "1"
    p.render_device(out.data_ptr(), frames); t2 = time.perf_counter()
    print("build %.1f ms render call %.1f ms device %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, p.last_render_stats().device_ms))
    p.close()
