"""PB200_HOST_PROF: host phases + device timeline of one render of a bench workload, stand-alone and as a rank's subtree (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PB200_HOST_PROF"] = "1"
import torch, bench, phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import Player
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
api = phonic_b200.load_api()
spec = bench.workload_spec(wl)
frames = W.frames_for(spec["seconds"], 48000)
out = torch.zeros(frames, 2, device="cuda")
for sub in (False, False, True, True):
    p = Player(api, 48000); bench.build_scene(p, wl, as_subtree=sub)
    print("== as_subtree", sub, file=sys.stderr)
    p.render_device(out.data_ptr(), frames)
    p.close()
