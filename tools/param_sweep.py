"""Every effect alone on the main bus with one event per parameter, GPU vs oracle (debug aid: parameter-update paths)."""
import os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import phonic_b200
from conftest import ORACLE_LIB
from phonic_b200._capi import CApi
from phonic_b200 import player as P
from phonic_b200 import workloads as W
from scenes import tone
SRX = 48000
apis = {"gpu": phonic_b200.load_api(), "oracle": CApi(ORACLE_LIB, "po_")}
CASES = {
    "filter": (lambda: P.FilterEffect(0, 3000.0, 0.707), [("cuto", 800.0), ("fltq", 2.0)]),
    "eq5": (lambda: P.Eq5Effect(), [("gan1", 4.0), ("frq3", 2000.0), ("bw_2", 2.5)]),
    "compressor": (lambda: P.CompressorEffect(), [("thrs", -24.0), ("rato", 4.0), ("knee", 6.0), ("attk", 0.05), ("rels", 0.5), ("gain", 3.0), ("look", 0.02)]),
    "limiter": (lambda: P.CompressorEffect.new_limiter(), [("thrs", -6.0), ("rels", 0.3), ("look", 0.01)]),
    "chorus": (lambda: P.ChorusEffect(), [("rate", 2.0), ("phas", 1.0), ("dpth", 0.5), ("fdbk", -0.4), ("dlay", 20.0), ("wet_", 0.7), ("fltf", 4000.0), ("fltq", 0.5)]),
    "delay": (lambda: P.DelayEffect(), [("dlay", 200.0), ("fdbk", 0.3), ("cuto", 3000.0), ("driv", 0.5), ("wet_", 0.8), ("wdth", 0.9), ("lfor", 2.0), ("lfdt", 0.3), ("ldfb", 0.2), ("lfdf", -0.3)]),
    "reverb": (lambda: P.ReverbEffect(0.6, 0.35), [("room", 0.8), ("wet ", 0.6)]),
    "gate": (lambda: P.GateEffect(), [("thrs", -20.0), ("attk", 0.01), ("hold", 0.05), ("rels", 0.5), ("rnge", -30.0)]),
    "gain": (lambda: P.GainEffect(), [("gain", 0.5)]),
    "pan": (lambda: P.PanningEffect(), [("pan ", -0.4), ("wdth", 1.5)]),
    "distortion": (lambda: P.DistortionEffect(), [("driv", 2.0), ("mix ", 0.5)]),
}
for name, (make, params) in CASES.items():
    for pid, val in params:
        outs = []
        err = None
        for key in ("gpu", "oracle"):
            try:
                p = P.Player(apis[key], SRX)
                b = p.upload_buffer(tone(30000, 44100, seed=51), 44100)
                fx = p.add_effect(make())
                fx.set_parameter(pid, val, 20011)
                p.play_file_source(b, P.FilePlaybackOptions(volume=0.6, repeat=3))
                outs.append(p.render(W.frames_for(2, SRX))); p.close()
            except Exception as e:
                err = repr(e)[:80]; break
        if err: print(name, pid, "ERROR", err); continue
        d = np.abs(outs[0] - outs[1])
        rms = 20 * np.log10(np.sqrt(np.mean((outs[0] - outs[1]) ** 2)) + 1e-30)
        print(f"{name:10s} {pid} peak {np.abs(outs[1]).max():.3f} max {d.max():.2e} rms {rms:7.1f} dBFS" + ("   <-- above 1e-5" if d.max() > 1e-5 else ""))
