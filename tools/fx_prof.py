"""PB200_FX_PROF cycle breakdown of mix_fx on a bench workload (debug aid).
usage: fx_prof.py [workload] [level: 1 main mixer only | 2 per mixer and stage]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import Player
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
spec = bench.workload_spec(wl)
api = phonic_b200.load_api()
frames = W.frames_for(spec["seconds"], 48000)
for it in range(2):
    if it == 1: os.environ["PB200_FX_PROF"] = sys.argv[2] if len(sys.argv) > 2 else "1"
    p = Player(api, 48000); bench.build_scene(p, wl); p.render(frames)
    st = p.last_render_stats(); print("device ms", st.device_ms, "fx ms", st.effect_kernel_ms, "skel", st.skeleton_kernel_ms, "voice", st.voice_kernel_ms)
    p.close()
