import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import Player
out = os.path.join(ROOT, 'gpurun_out', 'skel_prof_cfg4.csv')
api = phonic_b200.load_api()
frames = W.frames_for(10, 48000)
for it in range(2):
    if it == 1: os.environ["PB200_SKEL_PROF"] = out
    p = Player(api, 48000); W.build_cfg4(p, 160); p.render(frames); st = p.last_render_stats(); p.close()
d = np.loadtxt(out, delimiter=",", skiprows=1)
print(os.environ.get("PB200_LIB", "in-tree")[-40:], "skel ms", round(st.skeleton_kernel_ms, 2), "mean Mcycles per voice:", np.round(d[:, 1:].mean(axis=0) / 1e6, 2))
