"""Per-voice cycle breakdown of the skeleton pass on cfg2 (PB200_SKEL_PROF debug counters)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import Player

out = os.path.join(ROOT, "gpurun_out", "skel_prof.csv")
api = phonic_b200.load_api()
frames = W.frames_for(10, 48000)
for it in range(2):
    if it == 1:
        os.environ["PB200_SKEL_PROF"] = out
    p = Player(api, 48000)
    W.build_cfg2(p)
    p.render(frames)
    st = p.last_render_stats()
    p.close()
print("skeleton ms", st.skeleton_kernel_ms, "device ms", st.device_ms)
d = np.loadtxt(out, delimiter=",", skiprows=1)
tot, simple, general, fr = d[:, 1], d[:, 2], d[:, 3], d[:, 4]
print("voices", len(d), "sum frames", fr.sum())
print("per-voice cycles inside phase_run of non-fused simple pieces: max %.3g mean %.3g" % (tot.max(), tot.mean()))
print("simple: max %.3g mean %.3g   general: max %.3g mean %.3g" % (simple.max(), simple.mean(), general.max(), general.mean()))
work = simple + general
order = np.argsort(-work)[:12]
for i in order:
    print("voice %3d work %.3g cyc (simple %.3g [phase %.3g] general %.3g) frames %d -> %.1f cyc/frame" % (i, work[i], simple[i], tot[i], general[i], fr[i], work[i] / max(fr[i], 1)))
# per CTA (8 voices): sum over blocks of the slowest voice is not available; report max work per group
g = work.reshape(-1, 8)
print("per-group max work: max %.3g mean %.3g" % (g.max(axis=1).max(), g.max(axis=1).mean()))
