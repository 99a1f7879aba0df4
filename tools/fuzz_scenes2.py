"""Second fuzzer: wider than fuzz_scenes.py -- other WavStream block sizes, three render calls, HighQuality file sources,
granular samplers, sampler parameter automation, loop ranges, Delay / Reverb (judged on the error floor), move_effect.
usage: fuzz_scenes2.py [first seed] [count]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import phonic_b200
from conftest import ORACLE_LIB
from phonic_b200._capi import CApi
from phonic_b200 import player as P
from scenes import tone
from fuzz_scenes import FX

SRX = 48000
FX2 = dict(FX)
FX2["delay"] = (lambda r: P.DelayEffect(), [("fdbk", 0.1, 0.7), ("wet_", 0.2, 0.8)])
FX2["reverb"] = (lambda r: P.ReverbEffect(float(r.uniform(0.3, 0.8)), float(r.uniform(0.2, 0.5))), [("wet ", 0.2, 0.6)])


def build_and_render(api, seed):
    r = np.random.default_rng(1000003 * 7 + seed)
    bf = int(r.choice([1024, 1024, 512, 999, 256]))
    if os.environ.get("FZ_BF"): bf = int(os.environ["FZ_BF"])
    p = P.Player(api, SRX, block_frames=bf)
    rates = [int(r.choice([44100, 48000, 32000])) for _ in range(2)]
    lens = [int(r.integers(9000, 40000)) for _ in range(2)]
    bufs = [p.upload_buffer(tone(lens[i], 44100, channels=int(r.integers(1, 3)), seed=seed * 11 + i), rates[i]) for i in range(2)]
    mixers = [None]
    for _ in range(int(r.integers(0, 4))):
        parent = mixers[int(r.integers(0, len(mixers)))]
        mixers.append(p.add_mixer(parent.id if parent is not None else None))
    effects, feedback = [], False
    for m in mixers:
        for _ in range(int(r.integers(0, 4))):
            name = str(r.choice(list(FX2)))
            if os.environ.get("FZ_NOFB") and name in ("delay", "reverb"): name = "filter"
            if os.environ.get("FZ_ONLY") and name != os.environ["FZ_ONLY"]: name = "gain"
            feedback |= name in ("delay", "reverb")
            make, params = FX2[name]
            fx = p.add_effect(make(r), m.id) if m is not None else p.add_effect(make(r))
            effects.append((fx, params, m))
    blocks_total = (80 * 1024) // bf
    frames_total = blocks_total * bf
    tm = lambda m: m.id if m is not None else P.A.MAIN_MIXER
    for _ in range(int(r.integers(1, 5))):
        m = mixers[int(r.integers(0, len(mixers)))]
        bi = int(r.integers(0, 2))
        hq = r.random() < 0.25 and not os.environ.get("FZ_NOHQ")
        o = P.FilePlaybackOptions(volume=float(r.uniform(0.1, 0.5)), panning=float(r.uniform(-1, 1)), speed=1.0 if hq else float(r.choice([1.0, 0.5, 1.7, 2.0, 0.93])),
                                  repeat=int(r.integers(0, 4)), target_mixer=tm(m), resampling_quality=1 if hq else 0)
        if r.random() < 0.3:
            a0 = int(r.integers(0, lens[bi] // 2)); o.loop_range = (a0, a0 + int(r.integers(2000, lens[bi] // 2)))
        f = p.play_file_source(bufs[bi], o, start_time=int(r.integers(0, frames_total // 2)))
        k = r.random(); t = int(r.integers(1000, frames_total))
        if k < 0.2: f.seek((int(r.integers(0, 8000)) + 0.25) / rates[bi], sample_time=t)
        elif k < 0.4 and not hq: f.set_speed(float(r.uniform(0.5, 2.0)), glide=float(r.uniform(5, 50)) if r.random() < 0.5 else None, sample_time=t)
        elif k < 0.55: f.stop(stop_time=t)
        elif k < 0.7: f.set_volume(float(r.uniform(0.1, 0.8)), sample_time=t)
        elif k < 0.8: f.set_panning(float(r.uniform(-1, 1)), sample_time=t)
    for _ in range(int(r.integers(0, 3))):
        m = mixers[int(r.integers(0, len(mixers)))]
        env = P.AhdsrParameters(attack=float(r.uniform(0.001, 0.05)), hold=float(r.uniform(0, 0.05)), decay=float(r.uniform(0.02, 0.3)), sustain=float(r.uniform(0.2, 0.9)), release=float(r.uniform(0.02, 0.4))) if r.random() < 0.8 else None
        gran = None
        if r.random() < 0.3 and not os.environ.get("FZ_NOGRAN"):
            gran = P.GranularParameters(overlap_mode=int(r.integers(0, 2)), window=int(r.integers(0, 6)), size=float(r.uniform(20, 150)), density=float(r.uniform(5, 60)),
                                        position=float(r.uniform(0, 0.8)), step=float(r.choice([0.0, 1.0, 0.5])))
        g = p.add_generator(bufs[int(r.integers(0, 2))], P.GeneratorPlaybackOptions(voices=int(r.integers(1, 6)), volume=float(r.uniform(0.3, 0.9)), target_mixer=tm(m)), env, granular=gran)
        notes = []
        for _ in range(int(r.integers(1, 9))):
            t = int(r.integers(0, frames_total - 2000))
            notes.append((g.note_on(int(r.integers(40, 90)), volume=float(r.uniform(0.2, 0.8)), panning=float(r.uniform(-1, 1)), sample_time=t), t))
        for nid, t in notes:
            k = r.random(); t2 = t + int(r.integers(100, 30000))
            if k < 0.4: g.note_off(nid, sample_time=t2)
            elif k < 0.55 and gran is None: g.set_note_speed(nid, float(r.uniform(0.5, 2.0)), glide=float(r.uniform(5, 60)) if r.random() < 0.7 else None, sample_time=t2)
            elif k < 0.65: g.set_note_volume(nid, float(r.uniform(0.1, 1.0)), sample_time=t2)
            elif k < 0.75: g.set_note_panning(nid, float(r.uniform(-1, 1)), sample_time=t2)
        if gran is None:
            for _ in range(int(r.integers(0, 3))):
                pid, lo, hi = [("STRN", -12, 12), ("SFTN", -50, 50), ("SVOL", 0.2, 1.0), ("SPAN", -0.8, 0.8), ("AATK", 0.001, 0.05), ("ADCY", 0.02, 0.3), ("ASTN", 0.2, 0.9), ("AREL", 0.02, 0.4)][int(r.integers(0, 8))]
                try: g.set_parameter(pid, float(np.round(r.uniform(lo, hi)) if pid in ("STRN", "SFTN") else r.uniform(lo, hi)), int(r.integers(0, frames_total)))
                except P.PhonicError: pass
    for fx, params, m in effects:
        for _ in range(int(r.integers(0, 3))):
            pid, lo, hi = params[int(r.integers(0, len(params)))]
            fx.set_parameter(pid, float(r.uniform(lo, hi)), int(r.integers(0, frames_total)))
    cuts = sorted(int(x) for x in r.integers(1, blocks_total, size=2))
    parts = [p.render(cuts[0] * bf)]
    if effects and r.random() < 0.4 and not os.environ.get("FZ_NOMOVE"):
        fx, _, m = effects[int(r.integers(0, len(effects)))]
        try: p.move_effect(str(r.choice(["start", "end"])), fx.id, tm(m))
        except P.PhonicError: pass
    if cuts[1] > cuts[0]: parts.append(p.render((cuts[1] - cuts[0]) * bf))
    if len(mixers) > 1 and r.random() < 0.3: p.remove_mixer(mixers[-1].id)
    parts.append(p.render((blocks_total - cuts[1]) * bf))
    p.close()
    return np.concatenate(parts), feedback


def verdict(x, y, feedback):
    d = x - y
    mx = float(np.abs(d).max())
    if not feedback:
        return mx <= 1e-5, mx
    rms = float(np.sqrt(np.mean(d ** 2)))
    return (20 * np.log10(rms + 1e-30) < -90.0 and mx < 1e-3), mx


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    only = sys.argv[3] if len(sys.argv) > 3 else "both"
    orc = CApi(ORACLE_LIB, "po_")
    gpu = phonic_b200.load_api() if only == "both" else None
    bad = 0
    import faulthandler
    for seed in range(first, first + count):
        faulthandler.dump_traceback_later(60, exit=True)
        try:
            y, fb = build_and_render(orc, seed)
            if gpu is None:
                print("seed", seed, "oracle peak %.3f" % np.abs(y).max(), "feedback" if fb else "", flush=True); continue
            x, _ = build_and_render(gpu, seed)
        except Exception as e:
            print("seed", seed, "ERROR", repr(e)[:200], flush=True); bad += 1
            if "illegal memory" in repr(e) or "CUDA error" in repr(e): print("stopping: the CUDA context is gone"); break
            continue
        finally:
            faulthandler.cancel_dump_traceback_later()
        ok, mx = verdict(x, y, fb)
        if not ok:
            dd = np.abs(x - y).max(axis=1)
            print("seed", seed, "MISMATCH max %.2e first frame %d peak %.3f %s" % (mx, int(np.flatnonzero(dd > 1e-5)[0]), np.abs(y).max(), "(feedback fx)" if fb else ""), flush=True); bad += 1
    print("done:", count, "seeds,", bad, "bad")
