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
workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
per_group = {"cfg2": 8, "cfg4": 8, "sinc": 1}[workload]
for it in range(2):
    if it == 1:
        os.environ["PB200_SKEL_PROF"] = out
    p = Player(api, 48000)
    if workload == "cfg2":
        W.build_cfg2(p)
    elif workload == "cfg4":
        W.build_cfg4(p)
    else:
        W.build_sinc_bank(p, 256)
    p.render(frames)
    st = p.last_render_stats()
    p.close()
print("skeleton ms", st.skeleton_kernel_ms, "device ms", st.device_ms)
d = np.loadtxt(out, delimiter=",", skiprows=1)
sync, simple, general, res = d[:, 1], d[:, 2], d[:, 3], d[:, 4]
work = simple + general
other = res - work - sync
print("voices", len(d))
print("resident cycles (sum over blocks): max %.3g mean %.3g" % (res.max(), res.mean()))
print("free-run voice section: max %.3g mean %.3g | waiting at barriers: max %.3g mean %.3g | other code: max %.3g mean %.3g"
      % (work.max(), work.mean(), sync.max(), sync.mean(), other.max(), other.mean()))
g = lambda x: x.reshape(-1, per_group)
gi = int(np.argmax(g(res).max(axis=1)))
print("slowest group %d: per voice work / sync / other (Mcycles)" % gi)
for v in range(per_group):
    i = gi * per_group + v
    print("  voice %3d free-run %.2f sync %.2f other %.2f resident %.2f (event-chunk path %.2f)" % (i, simple[i] / 1e6, sync[i] / 1e6, (res[i] - simple[i] - sync[i]) / 1e6, res[i] / 1e6, general[i] / 1e6))
