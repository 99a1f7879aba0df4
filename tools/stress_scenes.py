"""Renders every parity scene repeatedly in ONE process (pool re-use, stale buffers) and reports the first API error."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import phonic_b200
from phonic_b200.player import Player
from scenes import SCENES, SR
api = phonic_b200.load_api()
ref = {}
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    for name in sorted(n for n in SCENES if len(sys.argv) < 3 or n in sys.argv[2:]):
        try:
            p = Player(api, SR)
            info = SCENES[name](p)
            out = p.render(info["frames"])
            p.close()
        except Exception as e:
            print("ERROR", rep, name, repr(e)); sys.exit(1)
        if name in ref and not np.array_equal(ref[name], out):
            print("NONDETERMINISTIC", rep, name, float(np.abs(ref[name] - out).max())); sys.exit(2)
        ref.setdefault(name, out)
print("ok")
