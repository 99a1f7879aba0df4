"""Random mixer graphs (nested sub-mixers, effect chains, file sources and samplers, events of every kind, structural
changes between two render calls) rendered by the GPU library and the oracle: prints every seed whose outputs differ by
more than the bar (debug aid; the seeds that ever failed become tests).  usage: fuzz_scenes.py [first seed] [count]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import phonic_b200
from conftest import ORACLE_LIB
from phonic_b200._capi import CApi
from phonic_b200 import player as P
from scenes import tone

SRX = 48000
FX = {
    "filter": (lambda r: P.FilterEffect(int(r.integers(0, 3)), float(r.uniform(300, 6000)), float(r.uniform(0.3, 2.0))), [("cuto", 200.0, 8000.0), ("fltq", 0.2, 3.0)]),
    "eq5": (lambda r: P.Eq5Effect(), [("gan1", -6.0, 6.0), ("gan3", -6.0, 6.0), ("frq2", 200.0, 2000.0)]),
    "compressor": (lambda r: P.CompressorEffect(), [("thrs", -40.0, -6.0), ("rato", 1.5, 12.0), ("rels", 0.1, 1.0)]),
    "chorus": (lambda r: P.ChorusEffect(), [("rate", 0.1, 4.0), ("dpth", 0.0, 0.8), ("wet_", 0.1, 0.9)]),
    "gate": (lambda r: P.GateEffect(), [("thrs", -50.0, -10.0), ("rels", 0.02, 0.5)]),
    "gain": (lambda r: P.GainEffect(), [("gain", 0.1, 2.0)]),
    "pan": (lambda r: P.PanningEffect(), [("pan ", -1.0, 1.0), ("wdth", 0.0, 2.0)]),
    "distortion": (lambda r: P.DistortionEffect(), [("driv", 0.0, 3.0), ("mix ", 0.0, 1.0)]),
}


def build_and_render(api, seed):
    r = np.random.default_rng(seed)
    p = P.Player(api, SRX)
    rates = [int(r.choice([44100, 48000, 32000])) for _ in range(2)]
    bufs = [p.upload_buffer(tone(int(r.integers(8000, 40000)), 44100, channels=int(r.integers(1, 3)), seed=seed * 7 + i), rates[i]) for i in range(2)]
    mixers = [None]
    for _ in range(int(r.integers(0, 4))):
        parent = mixers[int(r.integers(0, len(mixers)))]
        mixers.append(p.add_mixer(parent.id if parent is not None else None))
    effects = []
    for m in mixers:
        for _ in range(int(r.integers(0, 3))):
            name = str(r.choice(list(FX)))
            make, params = FX[name]
            fx = p.add_effect(make(r), m.id if m is not None else None) if m is not None else p.add_effect(make(r))
            effects.append((fx, params))
    frames_total = 64 * 1024
    files, gens = [], []
    for _ in range(int(r.integers(1, 5))):
        m = mixers[int(r.integers(0, len(mixers)))]
        o = P.FilePlaybackOptions(volume=float(r.uniform(0.1, 0.6)), panning=float(r.uniform(-1, 1)), speed=float(r.choice([1.0, 0.5, 1.7, 2.0, 0.93])),
                                  repeat=int(r.integers(0, 4)), target_mixer=m.id if m is not None else P.A.MAIN_MIXER)
        if r.random() < 0.3: o.fade_in = float(r.uniform(0.01, 0.2))
        bi = int(r.integers(0, 2))
        files.append((p.play_file_source(bufs[bi], o, start_time=int(r.integers(0, frames_total // 2))), rates[bi]))
    for _ in range(int(r.integers(0, 3))):
        m = mixers[int(r.integers(0, len(mixers)))]
        env = P.AhdsrParameters(attack=float(r.uniform(0.001, 0.05)), hold=float(r.uniform(0, 0.05)), decay=float(r.uniform(0.02, 0.3)), sustain=float(r.uniform(0.2, 0.9)), release=float(r.uniform(0.02, 0.4))) if r.random() < 0.8 else None
        g = p.add_generator(bufs[int(r.integers(0, 2))], P.GeneratorPlaybackOptions(voices=int(r.integers(1, 6)), volume=float(r.uniform(0.3, 0.9)), target_mixer=m.id if m is not None else P.A.MAIN_MIXER), env)
        notes = []
        for _ in range(int(r.integers(1, 9))):
            t = int(r.integers(0, frames_total - 2000))
            nid = g.note_on(int(r.integers(40, 90)), volume=float(r.uniform(0.2, 0.8)), panning=float(r.uniform(-1, 1)), sample_time=t)
            notes.append((nid, t))
        for nid, t in notes:
            k = r.random()
            t2 = t + int(r.integers(100, 30000))
            if k < 0.4: g.note_off(nid, sample_time=t2)
            elif k < 0.6: g.set_note_speed(nid, float(r.uniform(0.5, 2.0)), glide=float(r.uniform(5, 60)) if r.random() < 0.7 else None, sample_time=t2)
            elif k < 0.7: g.set_note_volume(nid, float(r.uniform(0.1, 1.0)), sample_time=t2)
        gens.append(g)
    for fx, params in effects:
        for _ in range(int(r.integers(0, 3))):
            pid, lo, hi = params[int(r.integers(0, len(params)))]
            fx.set_parameter(pid, float(r.uniform(lo, hi)), int(r.integers(0, frames_total)))
    for f, rate in files:
        k = r.random()
        t = int(r.integers(1000, frames_total))
        # whole-frame seek targets only: a seek that lands on an odd sample of a stereo buffer makes the REFERENCE spin forever at
        # the end of the file (preloaded.rs:138-146 truncates position x rate x channels to a sample index; cubic.rs:43-48 never
        # consumes the last, partial frame), and the oracle reproduces that
        if k < 0.25: f.seek((int(r.integers(0, 8000)) + 0.25) / rate, sample_time=t)
        elif k < 0.5: f.set_speed(float(r.uniform(0.5, 2.0)), glide=float(r.uniform(5, 50)) if r.random() < 0.5 else None, sample_time=t)
        elif k < 0.65: f.stop(stop_time=t)
        elif k < 0.8: f.set_volume(float(r.uniform(0.1, 0.8)), sample_time=t)
    a = p.render(frames_total // 2)
    # structural changes between the calls
    if len(mixers) > 1 and r.random() < 0.4:
        p.remove_mixer(mixers[-1].id)
    if effects and r.random() < 0.4:
        try: p.remove_effect(effects[0][0].id)
        except Exception: pass
    b = p.render(frames_total // 2)
    p.close()
    return np.concatenate([a, b])


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    gpu, orc = phonic_b200.load_api(), CApi(ORACLE_LIB, "po_")
    bad = 0
    for seed in range(first, first + count):
        try:
            x, y = build_and_render(gpu, seed), build_and_render(orc, seed)
        except Exception as e:
            print("seed", seed, "ERROR", repr(e)[:200]); bad += 1; continue
        d = np.abs(x - y).max(axis=1)
        if d.max() > 1e-5:
            print("seed", seed, "MISMATCH max %.2e first frame %d peak %.3f" % (d.max(), int(np.flatnonzero(d > 1e-5)[0]), np.abs(y).max())); bad += 1
    print("done:", count, "seeds,", bad, "bad")
