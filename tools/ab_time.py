"""A/B timing helper: device time of one workload with the library named by PB200_LIB (default: the in-tree build)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import Player
api = phonic_b200.load_api()
which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
for i in range(3):
    p = Player(api, 48000)
    if which == "cfg4": W.build_cfg4(p, 160)
    else: W.build_cfg2(p)
    p.render(W.frames_for(10, 48000)); st = p.last_render_stats()
    print(os.environ.get("PB200_LIB", "in-tree")[-60:], which, "device", round(st.device_ms, 2), "skel", round(st.skeleton_kernel_ms, 2), "voice", round(st.voice_kernel_ms, 2), "fx", round(st.effect_kernel_ms, 2))
    p.close()
