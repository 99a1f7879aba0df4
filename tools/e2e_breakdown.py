import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import Player, FilterEffect
api = phonic_b200.load_api()
frames = W.frames_for(10, 48000)
import torch, bench
out = torch.zeros(frames, 2, dtype=torch.float32).pin_memory().numpy()   # what bench.py's e2e arm renders into
os.environ["PB200_HOST_PROF"] = "1"
for it in range(4):
    t = [time.perf_counter()]
    buf = W.synth_buffer(int(4.0 * 44100), 44100, seed=1); t.append(time.perf_counter())
    p = Player(api, 48000); t.append(time.perf_counter())
    bid = p.upload_buffer(buf, 44100); t.append(time.perf_counter())
    hs = W.add_voice_bank_fast(p, W.VoiceBankSpec(voices=256), bid); t.append(time.perf_counter())
    p.add_effect(FilterEffect(0, 2000.0, 0.707)); t.append(time.perf_counter())
    p.render_into(out); t.append(time.perf_counter())
    st = p.last_render_stats()
    p.close(); t.append(time.perf_counter())
    names = ["synth_buffer", "Player()", "upload_buffer", "add_voice_bank", "add_effect", "render_into", "close"]
    print(it, " ".join(f"{n}={1e3*(b-a):.2f}ms" for n, a, b in zip(names, t, t[1:])), f"| device_ms={st.device_ms:.2f}")
