"""Re-runs one fuzz seed with one kind of call switched off at a time (GPU vs oracle): which call the mismatch needs."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import phonic_b200
import fuzz_scenes as F
from conftest import ORACLE_LIB
from phonic_b200._capi import CApi
from phonic_b200 import player as P
seed = int(sys.argv[1])
gpu, orc = phonic_b200.load_api(), CApi(ORACLE_LIB, "po_")
def run(label):
    x, y = F.build_and_render(gpu, seed), F.build_and_render(orc, seed)
    d = np.abs(x - y).max(axis=1)
    big = np.flatnonzero(d > 1e-5)
    print(f"{label:40s} max {d.max():.2e} first {int(big[0]) if big.size else None}", flush=True)
run("as is")
orig_sp = P.EffectHandle.set_parameter
for pid in ("rels", "rato", "driv", "thrs", "cuto", "gain"):
    def sp(self, p_, v, t=None, _pid=pid):
        if p_ == _pid: return
        return orig_sp(self, p_, v, t)
    P.EffectHandle.set_parameter = sp
    run("without effect param " + pid)
P.EffectHandle.set_parameter = orig_sp
for cls, n in ((P.FilePlaybackHandle, "seek"), (P.FilePlaybackHandle, "set_volume"), (P.FilePlaybackHandle, "set_speed"), (P.FilePlaybackHandle, "stop"),
               (P.GeneratorPlaybackHandle, "set_note_speed"), (P.GeneratorPlaybackHandle, "note_off"), (P.Player, "remove_mixer"), (P.Player, "remove_effect")):
    o = getattr(cls, n)
    setattr(cls, n, lambda self, *a, **k: None)
    run("without " + cls.__name__ + "." + n)
    setattr(cls, n, o)
# structure variants
class Dummy:
    id = 0
    def set_parameter(self, *a, **k): pass
o = P.Player.add_effect
P.Player.add_effect = lambda self, e, *a, **k: Dummy()
P.Player.remove_effect = lambda self, i: None
run("without any effect")
for kind in ("CompressorEffect", "DistortionEffect", "GateEffect", "ChorusEffect", "Eq5Effect", "FilterEffect", "GainEffect", "PanningEffect"):
    P.Player.add_effect = lambda self, e, *a, _k=kind, **k: Dummy() if type(e).__name__ == _k else o(self, e, *a, **k)
    run("without " + kind)
P.Player.add_effect = o
