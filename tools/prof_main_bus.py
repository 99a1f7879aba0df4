"""Times cfg5's main-bus stage (Delay + Reverb on a 10 s stereo bus) alone; PB200_FX_PROF=2 adds the per-stage cycle breakdown."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.distributed import finish_on_main_bus
api = phonic_b200.load_api()
frames = W.frames_for(10, 48000)
bus = (np.random.default_rng(1).standard_normal((frames, 2)) * 0.1).astype(np.float32)
for it in range(3):
    st = {}
    finish_on_main_bus(api, bus, 48000, W.add_main_bus_sends, stats=st)
    print("main bus", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
