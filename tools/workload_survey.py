"""Times the CUDA renderer on the other BASELINE configs (not bench lines): device ms, wall ms."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200._capi import CApi
from phonic_b200.player import Player

SR = 48000
which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg1", "cfg3s", "cfg5shard"]
cu = phonic_b200.load_api()
orc = CApi(os.path.join(ROOT, "oracle", "_build", "libphonic_oracle_fast.so"), "po_")


def scene(p, name):
    if name == "cfg1":
        buf = W.synth_buffer(44100 * 6, 44100, seed=5, channels=2)
        W.build_cfg1(p, buf, 44100)
        return 1, W.frames_for(10, SR)
    if name == "cfg3s":  # cfg3 at 10 s instead of 60 s
        W.build_subtrees(p, 64, 64, W.VoiceBankSpec(), effects="cfg3")
        return 4096, W.frames_for(10, SR)
    if name == "cfg5shard":
        W.build_subtrees(p, 64, 128, W.VoiceBankSpec(), effects="none")
        W.add_main_bus_sends(p)
        return 8192, W.frames_for(10, SR)
    if name == "cfg5shard_nofx":
        W.build_subtrees(p, 64, 128, W.VoiceBankSpec(), effects="none")
        return 8192, W.frames_for(10, SR)
    raise SystemExit(name)


for name in which:
    for rep in range(2):
        p = Player(cu, SR)
        t0 = time.perf_counter(); voices, frames = scene(p, name); t1 = time.perf_counter()
        out = np.zeros((frames, 2), np.float32)
        p.render_into(out); t2 = time.perf_counter()
        st = p.last_render_stats()
        p.close()
        print(f"{name} gpu rep{rep}: build={1e3*(t1-t0):.0f}ms render_wall={1e3*(t2-t1):.1f}ms device={st.device_ms:.1f}ms "
              f"skel={st.skeleton_kernel_ms:.1f} replay={st.voice_kernel_ms:.1f} fx={st.effect_kernel_ms:.1f} "
              f"-> {voices*frames/(st.device_ms/1e3)/1e9:.2f} Gvs/s  peak={np.abs(out).max():.3f}", flush=True)
    if "--cpu" in sys.argv or name in ("cfg1",):
        p = Player(orc, SR)
        voices, frames = scene(p, name)
        out2 = np.zeros((frames, 2), np.float32)
        t0 = time.perf_counter(); p.render_into(out2); dt = time.perf_counter() - t0
        print(f"{name} cpu oracle: {1e3*dt:.0f}ms -> {voices*frames/dt/1e9:.3f} Gvs/s  maxdiff={np.abs(out-out2).max():.2e}", flush=True)
