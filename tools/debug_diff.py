"""Debug helper: render scenes on the CUDA renderer and the oracle, print where they differ."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import phonic_b200
from phonic_b200._capi import CApi
from phonic_b200.player import Player
from scenes import SCENES, SR

names = sys.argv[1:] or sorted(SCENES)
cu = phonic_b200.load_api()
orc = CApi(os.path.join(ROOT, "oracle", "_build", "libphonic_oracle.so"), "po_")
for name in names:
    outs = []
    for api in (cu, orc):
        p = Player(api, SR)
        info = SCENES[name](p)
        outs.append(p.render(info["frames"]))
    g, r = outs
    d = np.abs(g - r)
    bad = np.flatnonzero(d.max(axis=1) > 0)
    print(f"{name}: frames={len(r)} peak={np.abs(r).max():.4f} maxerr={d.max():.3e} nbad={bad.size}", end="")
    if bad.size:
        print(f" first={bad[0]} last={bad[-1]}")
        # run-lengths of bad regions
        runs = np.split(bad, np.flatnonzero(np.diff(bad) > 1) + 1)
        print("   runs:", [(int(x[0]), int(x[-1])) for x in runs[:12]], "..." if len(runs) > 12 else "")
        i = bad[0]
        print("   gpu", g[i:i + 4].tolist(), "\n   ref", r[i:i + 4].tolist())
    else:
        print()
