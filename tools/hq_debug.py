"""Renders scenes on the device and the oracle and prints where they differ (debug aid)."""
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

names = sys.argv[1:] or ["hq_mono_to_eof"]
oracle = CApi(os.path.join(ROOT, "oracle", "_build", "libphonic_oracle.so"), "po_")
for name in names:
    outs = []
    for api in (phonic_b200.load_api(), oracle):
        p = Player(api, SR)
        info = SCENES[name](p)
        outs.append(p.render(info["frames"]))
        if "h" in info:
            st = info["h"].status()
            print(name, "status", st.is_playing, st.exhausted, st.playback_pos, st.end_frame)
        if "g" in info:
            print(name, "voices", info["g"].voice_states())
    g, r = outs
    d = np.abs(g - r).max(axis=1)
    bad = np.flatnonzero(d > 0)
    print(name, "max err %.3e" % d.max(), "differing frames", bad.size, bad[:10], "peak", np.abs(r).max())
    if bad.size:
        i = bad[0]
        print(" gpu", g[i:i + 4].ravel(), "\n ref", r[i:i + 4].ravel())
