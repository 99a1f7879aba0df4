"""Generates tests/golden/scenes.json from the ORACLE (the Rust reference cannot be built here, so
these vectors pin the oracle against regressions and give the bit-exact GPU scenes a committed target).
Run: python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

from phonic_b200._capi import CApi  # noqa: E402
from phonic_b200.player import Player  # noqa: E402
from scenes import SCENES, SR  # noqa: E402

api = CApi(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_build", "libphonic_oracle.so"), "po_")
out = {}
for name in sorted(SCENES):
    p = Player(api, SR)
    info = SCENES[name](p)
    a = p.render(info["frames"])
    nz = np.flatnonzero(np.abs(a).max(axis=1) > 0)
    first = int(nz[0]) if nz.size else -1
    out[name] = {
        "frames": int(info["frames"]),
        "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
        "peak": float(np.abs(a).max()),
        "sum": float(a.astype(np.float64).sum()),
        "first_nonzero_frame": first,
        "last_nonzero_frame": int(nz[-1]) if nz.size else -1,
        "probe": [float(x) for x in a[first:first + 8].reshape(-1)] if first >= 0 else [],
    }
    print(name, out[name]["peak"], out[name]["first_nonzero_frame"], out[name]["last_nonzero_frame"])
json.dump(out, open(os.path.join(HERE, "scenes.json"), "w"), indent=1)
