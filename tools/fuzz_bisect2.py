"""Variants of one fuzz_scenes2 seed, each in its own process (a variant may kill the CUDA context): which feature a mismatch needs."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and sys.argv[2] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
    import numpy as np
    import phonic_b200
    import fuzz_scenes2 as F
    from conftest import ORACLE_LIB
    from phonic_b200._capi import CApi
    from phonic_b200 import player as P
    seed, label = int(sys.argv[1]), sys.argv[3]
    skip = os.environ.get("FZ_SKIP", "")
    class Dummy:
        id = 0
        def set_parameter(self, *a, **k): pass
    if skip.startswith("fx:"):
        kind = skip[3:]
        o = P.Player.add_effect
        P.Player.add_effect = lambda self, e, *a, **k: Dummy() if (kind == "all" or type(e).__name__ == kind) else o(self, e, *a, **k)
        P.Player.move_effect = lambda self, *a, **k: None
    if skip == "loop":
        o2 = P.Player.play_file_source
        def pfs(self, b, opt, start_time=None):
            opt.loop_range = None; return o2(self, b, opt, start_time=start_time)
        P.Player.play_file_source = pfs
    if skip == "onecall":
        o3 = P.Player.render
        state = {"buf": []}
    if skip == "sparams": P.GeneratorPlaybackHandle.set_parameter = lambda self, *a, **k: None
    if skip == "pan": P.FilePlaybackHandle.set_panning = lambda self, *a, **k: None; P.GeneratorPlaybackHandle.set_note_panning = lambda self, *a, **k: None
    if skip == "rmmixer": P.Player.remove_mixer = lambda self, i: None
    if skip == "gens": P.Player.add_generator = lambda self, *a, **k: type("G", (), {"note_on": lambda s, *a, **k: 0, "note_off": lambda s, *a, **k: None, "set_note_speed": lambda s, *a, **k: None, "set_note_volume": lambda s, *a, **k: None, "set_note_panning": lambda s, *a, **k: None, "set_parameter": lambda s, *a, **k: None})()
    x, fb = F.build_and_render(phonic_b200.load_api(), seed)
    y, _ = F.build_and_render(CApi(ORACLE_LIB, "po_"), seed)
    d = np.abs(x - y).max(axis=1); big = np.flatnonzero(~(d <= 1e-5))
    print(f"{label:28s} max {np.nanmax(d):.2e} nan {int(np.isnan(x).sum())} first {int(big[0]) if big.size else None} {'(feedback)' if fb else ''}", flush=True)
    sys.exit(0)
seed = sys.argv[1]
variants = [("as is", {}), ("no pipeline", {"PB200_NO_FX_PIPELINE": "1"}), ("no effects", {"FZ_SKIP": "fx:all"}), ("no loop ranges", {"FZ_SKIP": "loop"}),
            ("no sampler params", {"FZ_SKIP": "sparams"}), ("no panning events", {"FZ_SKIP": "pan"}), ("no remove_mixer", {"FZ_SKIP": "rmmixer"}), ("no generators", {"FZ_SKIP": "gens"}),
            ("no persistent", {"PB200_NO_PERSISTENT": "1"}), ("no jumps", {"PB200_SKEL_DEBUG": "4"}), ("no simple calls", {"PB200_SKEL_DEBUG": "2"}), ("no autonomy", {"PB200_NO_AUTONOMOUS": "1"}), ("no direct child", {"PB200_NO_DIRECT_CHILD": "1"})]
for kind in ("CompressorEffect", "DistortionEffect", "GateEffect", "ChorusEffect", "Eq5Effect", "FilterEffect", "GainEffect", "PanningEffect"):
    variants.append(("no " + kind, {"FZ_SKIP": "fx:" + kind}))
for label, env in variants:
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, __file__, seed, "child", label], env=e, capture_output=True, text=True, timeout=120)
    out = (r.stdout.strip().splitlines() or [""])[-1]
    print(out if out else f"{label:28s} FAILED: " + (r.stderr.strip().splitlines() or ["?"])[-1][:160], flush=True)
