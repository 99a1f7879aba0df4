"""Skeleton-pass time of cfg2 variants (debug aid): which part of the score costs what."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import Player
api = phonic_b200.load_api()
frames = W.frames_for(10, 48000)
buf = W.synth_buffer(int(4.0 * 44100), 44100, seed=1)
def run(tag, **kw):
    best = None
    for it in range(3):
        p = Player(api, 48000)
        W.build_cfg2(p, W.VoiceBankSpec(**kw), buffer=buf)
        p.render(frames)
        st = p.last_render_stats()
        p.close()
        if best is None or st.device_ms < best[0]: best = (st.device_ms, st.skeleton_kernel_ms, st.voice_kernel_ms, st.effect_kernel_ms, st.voice_frames)
    print(f"{tag:28s} device {best[0]:6.2f} ms  skeleton {best[1]:6.2f}  replay {best[2]:6.2f}  fx {best[3]:6.2f}  active voice-frames {best[4] / 1e6:.1f}M")
run("cfg2")
run("no glide", glide=False)
run("no note-off", note_off=False)
run("no glide, no note-off", glide=False, note_off=False)
run("8 voices (1 group)", voices=8)
run("64 voices", voices=64)
run("256 voices, 4 per sampler", voices_per_sampler=4)
run("256 voices, 16 per sampler", voices_per_sampler=16)
run("256 voices, 1 per sampler", voices_per_sampler=1)
