"""Per-effect cost of the mixer pass (debug aid): one looping stereo file through ONE effect on the main bus, 10 s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import (ChorusEffect, CompressorEffect, DelayEffect, DistortionEffect, Eq5Effect, FilePlaybackOptions,
                                FilterEffect, GainEffect, GateEffect, PanningEffect, Player, ReverbEffect)
api = phonic_b200.load_api()
frames = W.frames_for(10, 48000)
buf = W.synth_buffer(44100, 44100, seed=1, channels=2)
cases = [("none", None), ("Filter", FilterEffect(0, 2000.0, 0.707)), ("Eq5", Eq5Effect()), ("Compressor", CompressorEffect()),
         ("Limiter", CompressorEffect.new_limiter()), ("Chorus", ChorusEffect()), ("Delay", DelayEffect()), ("Reverb", ReverbEffect(0.6, 0.35)),
         ("Gain+DC", GainEffect(-3.0, 2)), ("Gate", GateEffect(-30.0, 0.005, 0.1, 0.2, -60.0)), ("Distortion", DistortionEffect(2, 2.0, 1.0)),
         ("Panning", PanningEffect())]
for name, fx in cases:
    best = None
    for it in range(3):
        p = Player(api, 48000)
        b = p.upload_buffer(buf, 44100)
        o = FilePlaybackOptions(volume=0.5)
        o.repeat_forever()
        p.play_file_source(b, o)
        if fx is not None:
            h = p.add_effect(fx)
            if name == "Panning": h.set_parameter("pan ", 0.3, 0)
        p.render(frames)
        st = p.last_render_stats()
        p.close()
        if best is None or st.effect_kernel_ms < best[0]: best = (st.effect_kernel_ms, st.device_ms)
    print(f"{name:12s} mixer pass {best[0]:7.2f} ms   ({best[0] * 1e3 / 469:6.1f} us per 1024-frame chunk)   render {best[1]:7.2f} ms")
