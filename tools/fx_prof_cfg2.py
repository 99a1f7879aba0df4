"""PB200_FX_PROF cycle breakdown of the main mixer's mix_fx CTA on cfg2 (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import Player
api = phonic_b200.load_api()
frames = W.frames_for(10, 48000)
for it in range(2):
    if it == 1: os.environ["PB200_FX_PROF"] = "1"
    p = Player(api, 48000); W.build_cfg2(p); p.render(frames)
    st = p.last_render_stats(); print("device ms", st.device_ms, "fx ms", st.effect_kernel_ms, "skel", st.skeleton_kernel_ms, "voice", st.voice_kernel_ms)
    p.close()
