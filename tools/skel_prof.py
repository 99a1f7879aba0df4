"""Per-voice cycle counters of the skeleton pass on cfg2 (PB200_SKEL_PROF debug aid): where the voices of the
slowest group wait for each other."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import Player
out = os.path.join(ROOT, 'gpurun_out', 'skel_prof.csv')
api = phonic_b200.load_api()
frames = W.frames_for(10, 48000)
for it in range(2):
    if it == 1: os.environ["PB200_SKEL_PROF"] = out
    p = Player(api, 48000); W.build_cfg2(p); p.render(frames); p.close()
d = np.loadtxt(out, delimiter=",", skiprows=1)
g = d[:, 1:5].reshape(-1, 8, 4)
x = d[:, 5:].reshape(-1, 8, 8)
gi = int(np.argmax(g[:, :, 3].max(axis=1)))
print("group", gi, "columns: sync1(before run) sync2(after voices) sync3(after thread0 bookkeeping) free-run-work  [Mcycles]")
for v in range(8): print("  voice", gi * 8 + v, np.round(g[gi, v] / 1e6, 2))
print("mean over all voices:", np.round(d[:, 1:5].mean(axis=0) / 1e6, 2))
print("columns: simple calls, their Mcycles, jumped tiles, literal frames, general-path frames, their Mcycles, Mcycles in the fast tile loop, Mcycles in soft sc_advance")
for v in range(8):
    r = x[gi, v]
    print("  voice", gi * 8 + v, int(r[0]), round(r[1] / 1e6, 2), int(r[2]), int(r[3]), int(r[4]), round(r[5] / 1e6, 2), round(r[6] / 1e6, 2), round(r[7] / 1e6, 2))
m = d[:, 5:].mean(axis=0)
print("mean:", int(m[0]), round(m[1] / 1e6, 2), int(m[2]), int(m[3]), int(m[4]), round(m[5] / 1e6, 2), round(m[6] / 1e6, 2), round(m[7] / 1e6, 2))
