"""fuzz seed 28 rebuilt by hand, with variants (GPU vs oracle) (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import phonic_b200
from conftest import ORACLE_LIB
from phonic_b200._capi import CApi
from phonic_b200 import player as P
from scenes import tone
gpu, orc = phonic_b200.load_api(), CApi(ORACLE_LIB, "po_")
def run(label, comps1=2, comp2=True, dist2=True, f1=True, f3=True, vol_ev=True, seek_ev=True, calls=(32768, 32768), m3=True):
    outs = []
    for api in (gpu, orc):
        p = P.Player(api, 48000)
        b0 = p.upload_buffer(tone(15070, 44100, channels=2, seed=196), 48000)
        b = p.upload_buffer(tone(34658, 44100, channels=2, seed=197), 32000)
        m1 = p.add_mixer(None); m2 = p.add_mixer(None); mm3 = p.add_mixer(None) if m3 else m1
        for _ in range(comps1): p.add_effect(P.CompressorEffect(), m1.id)
        if comp2: p.add_effect(P.CompressorEffect(), m2.id)
        if dist2: p.add_effect(P.DistortionEffect(), m2.id)
        h1 = None
        if f1:
            o = P.FilePlaybackOptions(volume=0.26466819197719127, panning=0.636166600471562, speed=2.0, repeat=2, target_mixer=mm3.id); o.fade_in = 0.19288975494263616
            h1 = p.play_file_source(b, o, start_time=22807)
        o = P.FilePlaybackOptions(volume=0.18874803636233076, panning=0.03950292982505221, speed=0.93, repeat=2, target_mixer=m1.id); o.fade_in = 0.03824785638303426
        h2 = p.play_file_source(b, o, start_time=11274)
        if f3: p.play_file_source(b, P.FilePlaybackOptions(volume=0.40154412217587465, panning=-0.8288326513097459, speed=0.5, repeat=1, target_mixer=m2.id), start_time=28992)
        if seek_ev and h1 is not None: h1.seek(0.2414765625, sample_time=37619)
        if vol_ev: h2.set_volume(0.23486032953448177, sample_time=21420)
        outs.append(np.concatenate([p.render(c) for c in calls])); p.close()
    d = np.abs(outs[0] - outs[1]).max(axis=1); big = np.flatnonzero(d > 1e-5)
    if big.size and os.environ.get("SHOW"):
        i = int(big[0])
        print("gpu   ", outs[0][i - 2:i + 3].tolist()); print("oracle", outs[1][i - 2:i + 3].tolist())
    runs = []
    if big.size:
        st = big[0]; prev = big[0]
        for x in big[1:]:
            if x - prev > 64: runs.append((int(st), int(prev))); st = x
            prev = x
        runs.append((int(st), int(prev)))
    print(f"{label:34s} max {d.max():.2e} first {int(big[0]) if big.size else None} runs {runs[:6]}", flush=True)
run("minimal", f1=False, f3=False, comp2=False, dist2=False)
