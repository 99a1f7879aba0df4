"""Scenes shared with UPSTREAM phonic (rust/tests/dump_reference.rs): described as plain data so that the Rust harness
and this module build exactly the same graph -- the Rust side through phonic's public API on the CPU, this side through
the Player mirror on any implementation of the C-ABI. Only deterministic features appear (no OS-seeded state:
no Reverb, no random LFO shapes, no grain randomisation).

  python tests/reference_scenes.py --write     # writes tests/golden/reference/{scenes.json, inputs/*.f32}
  cargo test --manifest-path rust/Cargo.toml --features dump-reference --test dump_reference   # needs cargo + phonic
      -> tests/golden/reference/out/<scene>.wav, which tests/test_reference_fixtures.py compares with the oracle
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF_DIR = os.path.join(HERE, "golden", "reference")
SR = 48000
BLOCK = 1024


def _tone(frames, rate, channels, seed):
    from phonic_b200 import workloads as W
    return W.synth_buffer(frames, rate, seed=seed + 11, channels=channels)


BUFFERS = {
    "mono_44k": dict(frames=30000, rate=44100, channels=1, seed=0, loop=None),
    "stereo_48k": dict(frames=20000, rate=48000, channels=2, seed=3, loop=None),
    "mono_44k_loop": dict(frames=60000, rate=44100, channels=1, seed=5, loop=[10000, 50000]),
    "mono_48k_long": dict(frames=90000, rate=48000, channels=1, seed=7, loop=[20000, 80000]),
}

FILE_DEFAULT = dict(volume=1.0, panning=0.0, speed=1.0, repeat=None, loop_range=None, fade_in=None, fade_out=0.05, hq=False)
AHDSR = dict(attack=0.01, hold=0.0, decay=0.2, sustain=0.7, release=0.3)


def _file(buffer, start=None, events=(), **opts):
    return dict(type="file", buffer=buffer, start=start, options={**FILE_DEFAULT, **opts}, events=list(events))


SCENES = [
    dict(name="ref_file_mono_default", frames=40 * BLOCK, effects=[], sources=[_file("mono_44k")]),
    dict(name="ref_file_stereo_fast_loop", frames=48 * BLOCK, effects=[],
         sources=[_file("stereo_48k", start=1500, volume=0.7, panning=-0.3, speed=1.37, repeat=2, loop_range=[4000, 9000], fade_in=0.02)]),
    dict(name="ref_file_events", frames=70 * BLOCK, effects=[],
         sources=[_file("mono_44k_loop", volume=0.9, events=[
             dict(t=3000, kind="set_volume", value=0.4), dict(t=3000, kind="set_panning", value=0.6),
             dict(t=5000, kind="set_speed", speed=1.5, glide=24.0), dict(t=20000, kind="seek", seconds=0.25),
             dict(t=30000, kind="set_speed", speed=0.8, glide=None), dict(t=41000, kind="set_panning", value=-1.0),
             dict(t=60000, kind="stop")])]),
    dict(name="ref_hq_mono", frames=40 * BLOCK, effects=[], sources=[_file("mono_44k", hq=True)]),
    dict(name="ref_hq_stereo_down", frames=40 * BLOCK, effects=[],
         sources=[_file("stereo_48k", speed=1.37, repeat=2, loop_range=[4000, 9000], hq=True)]),
    dict(name="ref_sampler_notes", frames=96 * BLOCK, effects=[],
         sources=[dict(type="sampler", buffer="mono_44k_loop", voices=3, volume=0.8, panning=0.1, ahdsr=AHDSR, granular=None, events=[
             dict(t=100, kind="note_on", id=0, note=60, volume=0.6, panning=-0.3),
             dict(t=4000, kind="note_on", id=1, note=67, volume=0.5, panning=0.4),
             dict(t=9000, kind="note_on", id=2, note=55, volume=0.5, panning=0.0),
             dict(t=15000, kind="note_on", id=3, note=72, volume=0.4, panning=0.2),     # steals a voice
             dict(t=20000, kind="set_note_speed", ref=1, speed=1.8, glide=18.0),
             dict(t=26000, kind="set_note_volume", ref=2, value=0.2),
             dict(t=27000, kind="set_note_panning", ref=2, value=0.9),
             dict(t=40000, kind="note_off", ref=1), dict(t=52000, kind="note_off", ref=2), dict(t=60000, kind="all_notes_off")])]),
    dict(name="ref_fx_filter_eq5", frames=48 * BLOCK,
         effects=[dict(kind="filter", type=0, cutoff=1200.0, q=0.9), dict(kind="eq5", gains=[3.0, -4.0, 2.0, -6.0, 5.0])],
         sources=[_file("mono_44k_loop", volume=0.8, repeat="forever")]),
    dict(name="ref_fx_compressor", frames=48 * BLOCK,
         effects=[dict(kind="compressor", threshold=-24.0, ratio=6.0, knee=6.0, attack=0.005, release=0.1, makeup=6.0, lookahead=0.005)],
         sources=[_file("mono_44k_loop", volume=0.9, repeat="forever")]),
    dict(name="ref_fx_chorus", frames=48 * BLOCK,
         effects=[dict(kind="chorus", rate=1.2, phase=1.0, depth=0.4, feedback=0.4, delay=12.0, wet=0.5, filter_type=0, filter_freq=8000.0, filter_resonance=0.5)],
         sources=[_file("stereo_48k", volume=0.8, repeat="forever", loop_range=[1000, 19000])]),
    dict(name="ref_fx_delay", frames=96 * BLOCK, effects=[dict(kind="delay")],
         sources=[_file("mono_44k", volume=0.8)]),
    dict(name="ref_gran_cloud", frames=64 * BLOCK, effects=[],
         sources=[dict(type="sampler", buffer="mono_48k_long", voices=2, volume=1.0, panning=0.0, ahdsr=AHDSR,
                       granular=dict(overlap_mode=0, window=0, size=80.0, density=30.0, position=0.2, step=0.7, playback_direction=0), events=[
             dict(t=200, kind="note_on", id=0, note=60, volume=0.6, panning=-0.4),
             dict(t=9000, kind="note_on", id=1, note=64, volume=0.5, panning=0.3),
             dict(t=40000, kind="note_off", ref=0), dict(t=45000, kind="note_off", ref=1)])]),
    dict(name="ref_submixers", frames=48 * BLOCK, effects=[dict(kind="filter", type=3, cutoff=200.0, q=0.707)],
         mixers=[dict(effects=[dict(kind="filter", type=0, cutoff=3000.0, q=0.8)], sources=[_file("mono_44k", volume=0.5, speed=1.2)]),
                 dict(effects=[dict(kind="eq5", gains=[0.0, 4.0, 0.0, -3.0, 0.0])], sources=[_file("stereo_48k", volume=0.5, start=4000)])],
         sources=[_file("mono_44k_loop", volume=0.4, repeat="forever")]),
]


def buffer_data(name):
    b = BUFFERS[name]
    return _tone(b["frames"], b["rate"], b["channels"], b["seed"])


def write_manifest():
    os.makedirs(os.path.join(REF_DIR, "inputs"), exist_ok=True)
    for name in BUFFERS:
        np.ascontiguousarray(buffer_data(name), dtype="<f4").tofile(os.path.join(REF_DIR, "inputs", name + ".f32"))
    json.dump(dict(sample_rate=SR, buffers=BUFFERS, scenes=SCENES), open(os.path.join(REF_DIR, "scenes.json"), "w"), indent=1)


# ---- the same description on the Player mirror --------------------------------------------------------------------------
def _add_effect(p, e, mixer_id):
    from phonic_b200.player import ChorusEffect, CompressorEffect, DelayEffect, Eq5Effect, FilterEffect
    k = e["kind"]
    if k == "filter":
        return p.add_effect(FilterEffect(e["type"], e["cutoff"], e["q"]), mixer_id)
    if k == "eq5":
        h = p.add_effect(Eq5Effect(), mixer_id)
        for i, g in enumerate(e["gains"]):
            h.set_parameter(f"gan{i + 1}", g, sample_time=0)
        return h
    if k == "compressor":
        return p.add_effect(CompressorEffect(e["threshold"], e["ratio"], e["knee"], e["attack"], e["release"], e["makeup"], e["lookahead"]), mixer_id)
    if k == "chorus":
        return p.add_effect(ChorusEffect(e["rate"], e["phase"], e["depth"], e["feedback"], e["delay"], e["wet"], e["filter_type"],
                                         e["filter_freq"], e["filter_resonance"]), mixer_id)
    if k == "delay":
        return p.add_effect(DelayEffect(), mixer_id)
    raise ValueError(k)


def _add_source(p, s, bufs, mixer_id):
    from phonic_b200 import _capi as A
    from phonic_b200.player import AhdsrParameters, FilePlaybackOptions, GeneratorPlaybackOptions, GranularParameters
    bid = bufs[s["buffer"]]
    if s["type"] == "file":
        o = s["options"]
        fo = FilePlaybackOptions(volume=o["volume"], panning=o["panning"], speed=o["speed"], fade_in=o["fade_in"], fade_out=o["fade_out"],
                                 resampling_quality=1 if o["hq"] else 0, target_mixer=mixer_id or A.MAIN_MIXER)
        if o["repeat"] == "forever":
            fo.repeat_forever()
        elif o["repeat"] is not None:
            fo.repeat = o["repeat"]
        if o["loop_range"]:
            fo.loop_range = tuple(o["loop_range"])
        h = p.play_file_source(bid, fo, start_time=s["start"])
        for e in s["events"]:
            k, t = e["kind"], e["t"]
            if k == "set_volume": h.set_volume(e["value"], t)
            elif k == "set_panning": h.set_panning(e["value"], t)
            elif k == "set_speed": h.set_speed(e["speed"], e["glide"], t)
            elif k == "seek": h.seek(e["seconds"], t)
            elif k == "stop": h.stop(t)
        return h
    a = s["ahdsr"]
    env = AhdsrParameters(attack=a["attack"], hold=a["hold"], decay=a["decay"], sustain=a["sustain"], release=a["release"]) if a else None
    g = s["granular"]
    gran = GranularParameters(overlap_mode=g["overlap_mode"], window=g["window"], size=g["size"], density=g["density"], position=g["position"],
                              step=g["step"], playback_direction=g["playback_direction"]) if g else None
    h = p.add_generator(bid, GeneratorPlaybackOptions(volume=s["volume"], panning=s["panning"], voices=s["voices"]), env, mixer_id=mixer_id, granular=gran)
    ids = {}
    for e in s["events"]:
        k, t = e["kind"], e["t"]
        if k == "note_on": ids[e["id"]] = h.note_on(e["note"], volume=e["volume"], panning=e["panning"], sample_time=t)
        elif k == "note_off": h.note_off(ids[e["ref"]], sample_time=t)
        elif k == "all_notes_off": h.all_notes_off(sample_time=t)
        elif k == "set_note_speed": h.set_note_speed(ids[e["ref"]], e["speed"], glide=e["glide"], sample_time=t)
        elif k == "set_note_volume": h.set_note_volume(ids[e["ref"]], e["value"], sample_time=t)
        elif k == "set_note_panning": h.set_note_panning(ids[e["ref"]], e["value"], sample_time=t)
    return h


def build(p, scene):
    bufs = {}
    used = {s["buffer"] for s in scene["sources"]} | {s["buffer"] for m in scene.get("mixers", []) for s in m["sources"]}
    for name in sorted(used):
        b = BUFFERS[name]
        bufs[name] = p.upload_buffer(buffer_data(name), b["rate"], loop_range=tuple(b["loop"]) if b["loop"] else None)
    for m in scene.get("mixers", []):
        mh = p.add_mixer(None)
        for s in m["sources"]:
            _add_source(p, s, bufs, mh.id)
        for e in m["effects"]:
            _add_effect(p, e, mh.id)
    for s in scene["sources"]:
        _add_source(p, s, bufs, None)
    for e in scene["effects"]:
        _add_effect(p, e, None)
    return scene["frames"]


if __name__ == "__main__":
    if "--write" in sys.argv:
        write_manifest()
        print("wrote", REF_DIR)
