"""Seeded test scenes, described once through the Player API and rendered by whatever
implementation of the C-ABI the Player was given."""
import math

import numpy as np

from phonic_b200 import workloads as W
from phonic_b200.player import (AhdsrParameters, ChorusEffect, CompressorEffect, DelayEffect, Eq5Effect,
                                FilePlaybackOptions, DistortionEffect, FilterEffect, GainEffect, GateEffect, GeneratorPlaybackOptions, GranularParameters, PanningEffect,
                                Player, ReverbEffect)

SR = 48000
BLOCK = 1024


def tone(frames, rate, channels=1, seed=0):
    return W.synth_buffer(frames, rate, seed=seed + 11, channels=channels)


def file_mono_default(p: Player):
    """one mono 44.1 kHz file, default options, plays to EOF (ratio < 1 branch)"""
    b = p.upload_buffer(tone(30000, 44100), 44100)
    return {"h": p.play_file_source(b, FilePlaybackOptions()), "frames": 40 * BLOCK}


def file_stereo_fast_loop(p: Player):
    """stereo file at speed 1.37 (ratio >= 1 branch), loop range x2, fade-in, volume/pan options, late start"""
    b = p.upload_buffer(tone(20000, 48000, channels=2, seed=3), 48000)
    o = FilePlaybackOptions(volume=0.7, panning=-0.3, speed=1.37, repeat=2, loop_range=(4000, 9000), fade_in=0.02)
    return {"h": p.play_file_source(b, o, start_time=1500), "frames": 48 * BLOCK}


def file_events(p: Player):
    """seek, speed glide, volume/pan ramps, scheduled stop with fade-out, embedded loop"""
    b = p.upload_buffer(tone(60000, 44100, seed=5), 44100, loop_range=(10000, 50000))
    h = p.play_file_source(b, FilePlaybackOptions(volume=0.9))
    h.set_volume(0.4, 3000)
    h.set_panning(0.6, 3000)
    h.set_speed(1.5, 24.0, 5000)
    h.seek(0.25, 20000)
    h.set_speed(0.8, None, 30000)
    h.set_panning(-1.0, 41000)
    h.stop(60000)
    return {"h": h, "frames": 70 * BLOCK}


def file_bypass_and_immediate(p: Player):
    """equal rates (bypass copy), immediate (None) events between render calls are exercised by the test"""
    b = p.upload_buffer(tone(50000, 48000, seed=6), 48000)
    return {"h": p.play_file_source(b, FilePlaybackOptions(fade_out=0.01)), "frames": 60 * BLOCK}


def hq_mono_to_eof(p: Player):
    """HighQuality (rubato sinc) 44.1 -> 48 kHz, mono, plays through EOF: zero-padded tail chunks"""
    b = p.upload_buffer(tone(30000, 44100, seed=21), 44100)
    return {"h": p.play_file_source(b, FilePlaybackOptions(resampling_quality=1)), "frames": 40 * BLOCK}


def hq_stereo_down_loop(p: Player):
    """HighQuality stereo at speed 1.37 (down-sampling: cutoff scaled by the ratio), loop range x2 (zero padding
    at every loop end), fade-in, volume/pan options, late start"""
    b = p.upload_buffer(tone(20000, 48000, channels=2, seed=22), 48000)
    o = FilePlaybackOptions(volume=0.7, panning=-0.3, speed=1.37, repeat=2, loop_range=(4000, 9000), fade_in=0.02,
                            resampling_quality=1)
    return {"h": p.play_file_source(b, o, start_time=1500), "frames": 40 * BLOCK}


def hq_events(p: Player):
    """HighQuality with seek (drops pending output only), volume/pan ramps, same-speed set_speed, scheduled stop"""
    b = p.upload_buffer(tone(60000, 44100, seed=23), 44100, loop_range=(10000, 50000))
    h = p.play_file_source(b, FilePlaybackOptions(volume=0.9, resampling_quality=1))
    h.set_volume(0.4, 3000)
    h.set_panning(0.6, 3000)
    h.seek(0.25, 20000)
    h.set_speed(1.0, None, 25000)
    h.seek(0.9, 33333)
    h.set_panning(-1.0, 41000)
    h.stop(60000)
    return {"h": h, "frames": 70 * BLOCK}


def hq_equal_rates(p: Player):
    """HighQuality with equal rates: rubato bypass copy, still in 256-frame padded chunks at the loop end"""
    b = p.upload_buffer(tone(9000, 48000, channels=2, seed=24), 48000)
    o = FilePlaybackOptions(repeat=3, loop_range=(1000, 7268), resampling_quality=1)
    return {"h": p.play_file_source(b, o), "frames": 36 * BLOCK}


def hq_up_2x(p: Player):
    """HighQuality at speed 0.5 (2x up-sampling: > 512 frames per chunk) with another HighQuality voice beside it"""
    b = p.upload_buffer(tone(12000, 48000, seed=25), 48000)
    b2 = p.upload_buffer(tone(15000, 32000, channels=2, seed=26), 32000)
    h = p.play_file_source(b, FilePlaybackOptions(speed=0.5, volume=0.6, resampling_quality=1))
    p.play_file_source(b2, FilePlaybackOptions(volume=0.5, resampling_quality=1), start_time=777)
    return {"h": h, "frames": 36 * BLOCK}


def gran_cloud(p: Player):
    """granular Cloud mode: Hann 100 ms grains at 50 Hz from a moving playhead, AHDSR, note-offs, note volume / pan /
    speed changes, more notes than voices (stealing resets the grain pool)"""
    b = p.upload_buffer(tone(48000, 48000, seed=31), 48000)
    g = p.add_generator(b, GeneratorPlaybackOptions(volume=0.9, voices=3), AhdsrParameters(0.01, 0.0, 0.1, 0.7, 0.3),
                        granular=GranularParameters(window=0, size=100.0, density=50.0, position=0.2, step=1.0))
    n1 = g.note_on(60, 0.8, -0.3, sample_time=100)
    n2 = g.note_on(67, 0.5, 0.4, sample_time=5000)
    g.set_note_volume(n1, 0.3, sample_time=9000)
    g.set_note_panning(n2, -0.8, sample_time=12000)
    g.set_note_speed(n1, 1.5, glide=None, sample_time=15000)
    n3 = g.note_on(55, 0.6, 0.0, sample_time=20000)
    g.note_on(72, 0.4, 0.2, sample_time=26000)      # steals
    g.note_off(n2, sample_time=30000)
    g.note_off(n3, sample_time=40000)
    g.all_notes_off(50000)
    return {"g": g, "frames": 72 * BLOCK}


def gran_resampled_fixed(p: Player):
    """granular from a stereo 44.1 kHz file (mono buffer resampled on the device through the cubic path), fixed
    position, backward Triangle grains, no envelope: note_off stops triggering and the pool drains"""
    b = p.upload_buffer(tone(30000, 44100, channels=2, seed=32), 44100)
    g = p.play_generator(b, GeneratorPlaybackOptions(voices=2), None, start_time=300,
                         granular=GranularParameters(window=2, size=60.0, density=25.0, position=0.4, step=0.0,
                                                     playback_direction=1))
    n1 = g.note_on(62, 0.7, 0.5, sample_time=1000)
    n2 = g.note_on(50, 0.6, -0.5, sample_time=4000)
    g.note_off(n1, sample_time=20000)
    g.note_off(n2, sample_time=28000)
    g.stop(40000)
    return {"g": g, "frames": 48 * BLOCK}


def gran_sequential_loop(p: Player):
    """granular Sequential mode (crossfade-gated triggering) with an embedded loop the playhead runs into"""
    b = p.upload_buffer(tone(40000, 48000, seed=33), 48000, loop_range=(8000, 30000))
    g = p.add_generator(b, GeneratorPlaybackOptions(voices=2), AhdsrParameters(0.005, 0.0, 0.05, 0.8, 0.2),
                        granular=GranularParameters(overlap_mode=1, window=4, size=35.0, density=10.0, position=0.05, step=3.0))
    n1 = g.note_on(60, 0.8, 0.0, sample_time=0)
    n2 = g.note_on(64, 0.5, 0.6, sample_time=7777)
    g.note_off(n1, sample_time=45000)
    g.note_off(n2, sample_time=52000)
    return {"g": g, "frames": 72 * BLOCK}


def gran_dense(p: Player):
    """100 Hz x 400 ms grains: ~40 overlapping grains per voice, Blackman window"""
    b = p.upload_buffer(tone(36000, 48000, seed=34), 48000)
    g = p.add_generator(b, GeneratorPlaybackOptions(voices=2), AhdsrParameters(0.01, 0.0, 0.2, 0.6, 0.5),
                        granular=GranularParameters(window=1, size=400.0, density=100.0, position=0.3, step=0.5))
    n1 = g.note_on(57, 0.2, -0.2, sample_time=500)
    g.note_on(69, 0.15, 0.3, sample_time=9000)
    g.note_off(n1, sample_time=40000)
    g.all_notes_off(60000)
    return {"g": g, "frames": 96 * BLOCK}


def sampler_notes(p: Player):
    """8-voice sampler with AHDSR: more notes than voices (stealing), note-offs, glides, per-note vol/pan"""
    b = p.upload_buffer(tone(44100 * 2, 44100, seed=7), 44100, loop_range=(2000, 80000))
    ah = AhdsrParameters(attack=0.01, hold=0.05, decay=0.1, sustain=0.6, release=0.2)
    g = p.add_generator(b, GeneratorPlaybackOptions(volume=0.8, panning=0.1, voices=8), ah)
    rng = np.random.default_rng(42)
    ids = []
    t = 0
    for i in range(24):
        t += int(rng.integers(200, 4000))
        nid = g.note_on(int(rng.integers(40, 80)), volume=float(rng.uniform(0.1, 0.4)),
                        panning=float(rng.uniform(-1, 1)), sample_time=t)
        ids.append((nid, t))
        if i % 3 == 0:
            g.set_note_speed(nid, float(rng.uniform(0.5, 2.0)), glide=float(rng.uniform(6, 48)), sample_time=t + 3000)
        if i % 4 == 1:
            g.set_note_volume(nid, 0.2, sample_time=t + 2500)
            g.set_note_panning(nid, -0.5, sample_time=t + 2600)
        if i % 2 == 0:
            g.note_off(nid, sample_time=t + int(rng.integers(5000, 20000)))
    g.set_volume(0.5, 30000)
    g.set_panning(-0.4, 32000)
    g.all_notes_off(90000)
    return {"g": g, "frames": 110 * BLOCK}


def sampler_no_envelope(p: Player):
    """sampler without AHDSR: note_off -> 50 ms fade-out (chunk-dependent finish), transient + stop"""
    b = p.upload_buffer(tone(30000, 44100, channels=2, seed=8), 44100)
    g = p.play_generator(b, GeneratorPlaybackOptions(voices=4), None, start_time=700)
    n1 = g.note_on(60, 0.5, None, sample_time=1000)
    n2 = g.note_on(67, 0.4, 0.3, sample_time=1000)
    g.note_on(72, 0.3, -0.3, sample_time=5000)
    g.note_off(n1, sample_time=9000)
    g.note_off(n2, sample_time=9500)
    g.stop(30000)
    return {"g": g, "frames": 48 * BLOCK}


def cfg2_small(p: Player):
    hs, fx = W.build_cfg2(p, W.VoiceBankSpec(voices=48), time_scale=0.15)
    return {"hs": hs, "frames": 72 * BLOCK}


def filter_automation(p: Player):
    b = p.upload_buffer(tone(48000 * 2, 48000, channels=2, seed=9), 48000)
    p.play_file_source(b, FilePlaybackOptions(volume=0.8))
    fx = p.add_effect(FilterEffect(0, 800.0, 0.9))
    fx.set_parameter("cuto", 5000.0, 10000)
    fx.set_parameter("fltq", 2.0, 12000)
    fx.set_parameter_normalized("cuto", 0.3, 40000)
    fx.set_parameter("type", 3, 50000)
    return {"frames": 100 * BLOCK}


def fx_scene(effect, seconds=3.0, param_events=(), seed=10, buffer_seconds=1.0):
    def build(p: Player):
        b = p.upload_buffer(tone(int(44100 * buffer_seconds), 44100, channels=2, seed=seed), 44100)
        p.play_file_source(b, FilePlaybackOptions(volume=0.8))
        fx = p.add_effect(effect)
        for (pid, val, t) in param_events:
            fx.set_parameter(pid, val, t)
        return {"frames": W.frames_for(seconds, SR)}
    return build


def submixers_cfg3_small(p: Player):
    W.build_subtrees(p, 3, 12, W.VoiceBankSpec(), effects="cfg3", time_scale=0.1)
    return {"frames": 96 * BLOCK}


def submixers_cfg5_small(p: Player):
    W.build_subtrees(p, 4, 16, W.VoiceBankSpec(), effects="none", time_scale=0.1)
    W.add_main_bus_sends(p)
    return {"frames": 96 * BLOCK}


def many_groups(p: Player):
    """160 samplers of 8 voices (> one resident wave of warp-per-voice CTAs on 148 SMs): the skeleton takes its lane-per-voice mapping
    with the uniform phase loop, as it does for cfg3 / cfg5; no effects, so the whole path is bit-exact"""
    W.build_subtrees(p, 20, 64, W.VoiceBankSpec(), effects="none", time_scale=0.1)
    return {"frames": 96 * BLOCK}


def nested_and_gated(p: Player):
    """nested sub-mixers; one goes silent for > 2 s (silence gate + effect auto-bypass), then wakes up"""
    b = p.upload_buffer(tone(12000, 44100, seed=12), 44100)
    m1 = p.add_mixer(None)
    m2 = p.add_mixer(m1.id)
    p.add_effect(FilterEffect(0, 3000.0, 0.707), m2.id)
    p.add_effect(ChorusEffect(), m1.id)
    p.play_file_source(b, FilePlaybackOptions(volume=0.5, target_mixer=m2.id))
    p.play_file_source(b, FilePlaybackOptions(volume=0.5, target_mixer=m2.id), start_time=int(3.2 * SR))
    p.play_file_source(b, FilePlaybackOptions(volume=0.3, target_mixer=m1.id), start_time=5000)
    p.add_effect(CompressorEffect())
    return {"frames": W.frames_for(5, SR)}


def single_submixer_gated(p: Player):
    """the main mixer only passes ONE sub-mixer through (a rank's share of a sharded graph): the sub-mixer's kernel writes the
    output itself (renderer.cu direct_child). No effects anywhere (bit-exact); the sub-mixer -- and a nested one -- go silent
    for > 2 s, so the silence gate closes and re-opens."""
    b = p.upload_buffer(tone(12000, 44100, seed=21), 44100)
    m1 = p.add_mixer(None)
    m2 = p.add_mixer(m1.id)
    p.play_file_source(b, FilePlaybackOptions(volume=0.5, target_mixer=m1.id))
    p.play_file_source(b, FilePlaybackOptions(volume=0.4, panning=-0.3, target_mixer=m2.id), start_time=3000)
    p.play_file_source(b, FilePlaybackOptions(volume=0.5, target_mixer=m1.id), start_time=int(3.4 * SR) + 77)
    p.play_file_source(b, FilePlaybackOptions(volume=0.3, panning=0.5, target_mixer=m2.id), start_time=int(4.1 * SR))
    return {"frames": W.frames_for(5, SR)}


def single_submixer_filter(p: Player):
    """same shape with an effect chain on the single sub-mixer (staged chunks, bypass transitions) and a master volume != 1"""
    b = p.upload_buffer(tone(12000, 44100, seed=22), 44100)
    m1 = p.add_mixer(None)
    fx = p.add_effect(FilterEffect(0, 2500.0, 0.707), m1.id)
    fx.set_parameter("cuto", 1200.0, int(3.6 * SR))
    p.play_file_source(b, FilePlaybackOptions(volume=0.5, target_mixer=m1.id))
    p.play_file_source(b, FilePlaybackOptions(volume=0.5, target_mixer=m1.id), start_time=int(3.3 * SR) + 5)
    return {"frames": W.frames_for(5, SR)}


SCENES = {
    "single_submixer_gated": single_submixer_gated,
    "single_submixer_filter": single_submixer_filter,
    "file_mono_default": file_mono_default,
    "file_stereo_fast_loop": file_stereo_fast_loop,
    "file_events": file_events,
    "file_bypass": file_bypass_and_immediate,
    "hq_mono_to_eof": hq_mono_to_eof,
    "hq_stereo_down_loop": hq_stereo_down_loop,
    "hq_events": hq_events,
    "hq_equal_rates": hq_equal_rates,
    "hq_up_2x": hq_up_2x,
    "gran_cloud": gran_cloud,
    "gran_resampled_fixed": gran_resampled_fixed,
    "gran_sequential_loop": gran_sequential_loop,
    "gran_dense": gran_dense,
    "sampler_notes": sampler_notes,
    "sampler_no_envelope": sampler_no_envelope,
    "cfg2_small": cfg2_small,
    "filter_automation": filter_automation,
    "fx_filter": fx_scene(FilterEffect(0, 1000.0, 0.707)),
    "fx_eq5": fx_scene(Eq5Effect(), param_events=[("gan1", 4.0, 0), ("gan3", -5.0, 0), ("frq2", 1500.0, 20000), ("bw_3", 2.0, 30000)]),
    "fx_compressor": fx_scene(CompressorEffect()),
    "fx_limiter": fx_scene(CompressorEffect.new_limiter()),
    "fx_chorus": fx_scene(ChorusEffect(), param_events=[("rate", 2.0, 30000), ("dlay", 20.0, 40000)]),
    "fx_delay": fx_scene(DelayEffect(), seconds=4.0, param_events=[("dlay", 120.0, 0), ("fdbk", 0.6, 0), ("driv", 0.3, 50000), ("mode", 1, 90000)]),
    "fx_gain_dc": fx_scene(GainEffect(-4.5, 2), param_events=[("gain", 0.3, 20000), ("dcfm", 3, 40000), ("dcfm", 0, 70000), ("gain", 2.0, 90000)], buffer_seconds=2.8),
    "fx_panning": fx_scene(PanningEffect(), param_events=[("wdth", 0.4, 0), ("pan ", -0.6, 15000), ("invr", 1, 50000), ("wdth", 1.7, 80000),
                                                           ("pan ", 0.0, 100000), ("invr", 0, 110000), ("wdth", 1.0, 120000)], buffer_seconds=2.8),
    "fx_gate": fx_scene(GateEffect(-22.0, 0.004, 0.03, 0.15, -40.0), param_events=[("thrs", -15.0, 50000), ("rnge", -60.0, 80000), ("hold", 0.2, 100000)],
                        buffer_seconds=2.8, seed=14),
    "fx_dist_softclip": fx_scene(DistortionEffect(0, 2.5, 1.0), param_events=[("driv", 0.5, 30000), ("driv", 4.0, 60000)], buffer_seconds=2.8),
    "fx_dist_hardclip": fx_scene(DistortionEffect(1, 1.5, 0.7), param_events=[("mix ", 0.2, 20000), ("driv", 3.0, 50000), ("mix ", 1.0, 90000)], buffer_seconds=2.8),
    "fx_dist_fold": fx_scene(DistortionEffect(4, 3.0, 1.0), param_events=[("mix ", 0.0, 40000), ("mix ", 0.5, 70000), ("type", 1, 100000), ("type", 4, 110000)],
                             buffer_seconds=2.8),
    "fx_dist_diode": fx_scene(DistortionEffect(default=True), param_events=[("driv", 2.0, 10000), ("mix ", 0.6, 50000), ("type", 3, 80000), ("driv", 3.5, 100000)],
                              buffer_seconds=2.8),
    "fx_reverb": fx_scene(ReverbEffect(0.6, 0.35), seconds=4.0, param_events=[("room", 0.8, 60000)]),
    "submixers_cfg3_small": submixers_cfg3_small,
    "submixers_cfg5_small": submixers_cfg5_small,
    "nested_and_gated": nested_and_gated,
    "many_groups": many_groups,
}

# scenes whose whole path is +,-,*,/,sqrt in the reference's order: must be bit-exact on device
BIT_EXACT = {"single_submixer_gated", "many_groups", "fx_gain_dc", "fx_panning", "fx_dist_softclip", "fx_dist_hardclip", "fx_dist_fold", "file_mono_default", "file_stereo_fast_loop", "file_events", "file_bypass", "sampler_notes",
             "sampler_no_envelope", "hq_equal_rates", "gran_cloud", "gran_resampled_fixed", "gran_sequential_loop", "gran_dense"}
# bit-exact voice path + time-invariant biquads evaluated by the f64 block scan (exact up to O(1e-16)
# relative reassociation error before the f32 cast): at most a rare last-bit flip
NEAR_EXACT = {"cfg2_small", "fx_filter"}
