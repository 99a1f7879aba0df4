"""Structural mixer messages, effect messages and sampler parameter automation (every MixerMessage /
GeneratorPlaybackEvent variant):
  MixerMessage::{RemoveSource, RemoveMixer, MoveEffect, RemoveEffect, RemoveAllPendingEvents, ProcessEffectMessage}
  (src/source/mixed.rs:113-194, 294-499), Sampler::process_parameter_update (src/generator/sampler.rs:1069-1192),
  SamplerMessage::SetLoopRange (sampler.rs:1246-1271).
The oracle half runs on the CPU (sanity properties); the GPU half compares the CUDA renderer with the oracle."""
import numpy as np
import pytest

from phonic_b200 import PhonicError
from phonic_b200 import _capi as A
from phonic_b200 import workloads as W
from phonic_b200.player import (AhdsrParameters, DistortionEffect, FilePlaybackOptions, FilterEffect, GainEffect, GeneratorPlaybackOptions,
                                Player, ReverbEffect)

SR = 48000
BLOCK = 1024


def buf(frames=40000, rate=44100, seed=9, channels=1):
    return W.synth_buffer(frames, rate, seed=seed, channels=channels)


# ---- scenes: each returns the rendered audio (several render calls with control messages in between) ------------------
def scene_structural(api):
    """two sub-mixers with effects + sources; effects are moved / removed, a generator and a sub-mixer are removed
    between render calls"""
    p = Player(api, SR)
    b = p.upload_buffer(buf(), 44100, loop_range=(1000, 39000))
    m1, m2 = p.add_mixer(None), p.add_mixer(None)
    m3 = p.add_mixer(m2.id)
    o = FilePlaybackOptions(volume=0.5)
    o.repeat_forever()
    for mid, speed in ((None, 1.0), (m1.id, 1.25), (m2.id, 0.8), (m3.id, 1.5)):
        oo = FilePlaybackOptions(volume=0.4, speed=speed, target_mixer=mid or A.MAIN_MIXER)
        oo.repeat_forever()
        p.play_file_source(b, oo)
    g = p.add_generator(b, GeneratorPlaybackOptions(voices=4, target_mixer=m1.id), AhdsrParameters(attack=0.01, hold=0.0, decay=0.1, sustain=0.7, release=0.2), mixer_id=m1.id)
    g.note_on(60, volume=0.5, sample_time=500)
    g.note_on(67, volume=0.4, sample_time=9000)
    f1 = p.add_effect(FilterEffect(0, 1500.0, 0.9), m1.id)
    d1 = p.add_effect(DistortionEffect(0, 0.6, 0.8), m1.id)
    g1 = p.add_effect(GainEffect(0.5), m1.id)
    fm = p.add_effect(FilterEffect(2, 800.0, 1.2))
    dm = p.add_effect(DistortionEffect(1, 0.4, 1.0))
    parts = [p.render(8 * BLOCK)]
    p.move_effect("end", f1.id, m1.id)            # [dist, gain, filter]
    p.move_effect(-1, g1.id, m1.id)               # [gain, dist, filter]
    p.move_effect("start", dm.id)                 # main: [dist, filter]
    parts.append(p.render(8 * BLOCK))
    p.remove_effect(d1.id)
    p.remove_generator(g.id)
    parts.append(p.render(8 * BLOCK))
    p.remove_mixer(m2.id)                         # takes m3 and their sources along
    p.remove_effect(fm.id)
    parts.append(p.render(8 * BLOCK))
    p.remove_effect(dm.id)
    p.move_effect(5, f1.id, m1.id)
    parts.append(p.render(6 * BLOCK))
    with pytest.raises(PhonicError) as e:
        p.remove_mixer(m3.id)
    assert e.value.code == A.ERR_MIXER_NOT_FOUND
    with pytest.raises(PhonicError) as e:
        p.remove_effect(d1.id)
    assert e.value.code == A.ERR_EFFECT_NOT_FOUND
    with pytest.raises(PhonicError) as e:
        p.move_effect("end", f1.id, A.MAIN_MIXER)
    assert e.value.code == A.ERR_PARAMETER
    with pytest.raises(PhonicError) as e:
        g.note_on(60)
    assert e.value.code == A.ERR_SOURCE_NOT_PLAYING
    return np.concatenate(parts)


def scene_stop_all(api):
    """Player::stop_all_sources: playing transient sources fade out, scheduled ones and pending events disappear,
    a non-transient generator stays"""
    p = Player(api, SR)
    b = p.upload_buffer(buf(seed=12), 44100, loop_range=(1000, 39000))
    o = FilePlaybackOptions(volume=0.5)
    o.repeat_forever()
    h1 = p.play_file_source(b, o)
    h1.set_volume(0.1, 30 * BLOCK)                         # pending event: dropped
    o2 = FilePlaybackOptions(volume=0.5, speed=1.3)
    o2.repeat_forever()
    p.play_file_source(b, o2, start_time=20 * BLOCK)       # scheduled source: dropped
    gen = p.add_generator(b, GeneratorPlaybackOptions(voices=2), AhdsrParameters(attack=0.01, hold=0.0, decay=0.1, sustain=0.7, release=0.2))
    gen.note_on(64, volume=0.5, sample_time=100)
    gen.note_off_time = None
    tg = p.play_generator(b, GeneratorPlaybackOptions(voices=2), AhdsrParameters(attack=0.01, hold=0.0, decay=0.1, sustain=0.7, release=0.05))
    tg.note_on(72, volume=0.4, sample_time=200)
    parts = [p.render(6 * BLOCK)]
    p.stop_all_sources()
    parts.append(p.render(30 * BLOCK))
    with pytest.raises(PhonicError):
        h1.set_volume(0.3)
    gen.note_on(50, volume=0.3)                            # the added generator is still there
    parts.append(p.render(4 * BLOCK))
    return np.concatenate(parts)


def scene_reverb_reset(api, reset=True):
    p = Player(api, SR)
    b = p.upload_buffer(buf(20000, seed=14), 44100)
    p.play_file_source(b, FilePlaybackOptions(volume=0.8, repeat=0))
    rv = p.add_effect(ReverbEffect(0.8, 0.6))
    if reset:
        rv.send_message(A.MSG_REVERB_RESET, 30000)
    with pytest.raises(PhonicError):
        rv.send_message(77, 100)
    return p.render(48 * BLOCK)


def scene_sampler_params(api):
    """transpose / finetune / volume / panning / envelope automation + loop range messages on a cubic sampler"""
    p = Player(api, SR)
    b = p.upload_buffer(buf(60000, seed=15), 44100)
    env = AhdsrParameters(attack=0.02, hold=0.01, decay=0.2, sustain=0.6, release=0.3)
    g = p.add_generator(b, GeneratorPlaybackOptions(voices=4, volume=0.9), env)
    n1 = g.note_on(60, volume=0.5, panning=-0.2, sample_time=300)
    n2 = g.note_on(64, volume=0.4, panning=0.3, sample_time=5000)
    g.set_parameter("STRN", 3, 12000)
    g.set_parameter("SVOL", 0.5, 15000)
    g.set_parameter_normalized("SPAN", 0.9, 18000)
    g.set_parameter("SFTN", -37, 20000)
    g.set_parameter("ADCY", 0.05, 21000)
    g.set_parameter("ASTN", 0.3, 21000)
    g.set_parameter_normalized("AREL", 0.2, 22000)
    g.set_note_speed(n1, 1.1, glide=30.0, sample_time=23000)
    n3 = g.note_on(55, volume=0.5, sample_time=26000)      # starts with the automated base values
    g.set_loop_range((2000, 9000), 28000)
    g.note_off(n2, 30000)
    g.set_parameter("AATK", 0.001, 33000)
    g.set_parameter("AHLD", 0.0, 33000)
    n4 = g.note_on(72, volume=0.4, sample_time=36000)
    g.set_loop_range(None, 60000)
    g.note_off(n1, 70000)
    g.note_off(n3, 72000)
    g.note_off(n4, 74000)
    for bad in (("XXXX", 1.0), ("SVOL", float("nan"))):
        with pytest.raises(PhonicError) as e:
            g.set_parameter(bad[0], bad[1], 100)
        assert e.value.code == A.ERR_PARAMETER
    with pytest.raises(PhonicError):
        g.set_parameter_normalized("SVOL", 1.5, 100)
    with pytest.raises(PhonicError):
        g.set_loop_range((5000, 900000), 100)
    a = p.render(40 * BLOCK)
    g.set_parameter("STRN", -5)                            # immediate, between render calls
    bb = p.render(60 * BLOCK)
    return np.concatenate([a, bb]), g.voice_states()


SCENES = {"structural": scene_structural, "stop_all": scene_stop_all, "reverb_reset": scene_reverb_reset, "sampler_params": scene_sampler_params}


def audio(x):
    return x[0] if isinstance(x, tuple) else x


def test_oracle_structural_messages_change_the_mix(oracle_api):
    out = scene_structural(oracle_api)
    seg = [out[i * 8 * BLOCK:(i + 1) * 8 * BLOCK] for i in range(4)]
    assert all(np.abs(s).max() > 0.01 for s in seg)
    # removing a sub-mixer with two looping files lowers the level
    assert np.sqrt(np.mean(seg[3] ** 2)) < np.sqrt(np.mean(seg[2] ** 2))


def test_oracle_stop_all_sources(oracle_api):
    out = scene_stop_all(oracle_api)
    late = out[26 * BLOCK:36 * BLOCK]
    assert np.abs(out[:6 * BLOCK]).max() > 0.05
    # only the added generator's sustained note is left; the source scheduled for block 20 never started
    assert 0.0 < np.abs(late).max() < 0.5 * np.abs(out[:6 * BLOCK]).max()


def test_oracle_reverb_reset_cuts_the_tail(oracle_api):
    out, keep = scene_reverb_reset(oracle_api), scene_reverb_reset(oracle_api, reset=False)
    assert np.array_equal(out[:30000], keep[:30000])
    rms = lambda x: float(np.sqrt(np.mean(x.astype(np.float64) ** 2)))
    assert rms(keep[31000:40000]) > 1e-3 and rms(out[31000:40000]) < 0.3 * rms(keep[31000:40000])


def test_oracle_sampler_parameters(oracle_api):
    out, states = scene_sampler_params(oracle_api)
    assert np.isfinite(out).all() and np.abs(out).max() > 0.05
    assert all(s[3] == 0 for s in states)  # every note released and done
    # the same score without automation renders something else
    p = Player(oracle_api, SR)
    b = p.upload_buffer(buf(60000, seed=15), 44100)
    g = p.add_generator(b, GeneratorPlaybackOptions(voices=4, volume=0.9), AhdsrParameters(attack=0.02, hold=0.01, decay=0.2, sustain=0.6, release=0.3))
    g.note_on(60, volume=0.5, panning=-0.2, sample_time=300)
    plain = p.render(40 * BLOCK)
    assert np.array_equal(plain[:5000], out[:5000]) and not np.array_equal(plain[12000:16000], out[12000:16000])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENES))
def test_control_scene_matches_oracle(cuda_api, oracle_api, name):
    gpu, ref = SCENES[name](cuda_api), SCENES[name](oracle_api)
    if isinstance(ref, tuple):
        assert gpu[1] == ref[1]
    g, r = audio(gpu), audio(ref)
    assert g.shape == r.shape and np.abs(r).max() > 0.01
    err = float(np.abs(g - r).max())
    if name == "sampler_params":   # voice path only: bit-exact
        bad = np.flatnonzero((g != r).any(axis=1))
        assert bad.size == 0, f"first differing frame {bad[0]}, max err {err:.3e}"
    elif name == "reverb_reset":
        rms = float(np.sqrt(np.mean((g.astype(np.float64) - r) ** 2)))
        assert 20 * np.log10(max(rms, 1e-30)) < -90.0 and err < 1e-4
    else:
        assert err <= 1e-5, f"max abs err {err:.3e}"


# ---- observability: PlaybackStatusEvent stream + the main mixer's level meter ----------------------------
def scene_observability(api):
    p = Player(api, SR)
    p.set_metering_interval(0.25)
    b = p.upload_buffer(buf(150000, seed=21), 44100)
    b2 = p.upload_buffer(buf(30000, rate=48000, seed=22, channels=2), 48000, loop_range=(1000, 29000))
    h1 = p.play_file_source(b, FilePlaybackOptions(volume=0.6, speed=1.1))            # plays to its end (~3.1 s)
    o = FilePlaybackOptions(volume=0.5, fade_out=0.02)
    o.repeat_forever()
    h2 = p.play_file_source(b2, o, start_time=5000)
    h2.stop(100000)                                                                    # stopped, with fade-out
    o3 = FilePlaybackOptions(volume=0.4, fade_out=None, resampling_quality=1)
    h3 = p.play_file_source(b, o3, start_time=20000)
    h3.stop(70000)                                                                     # stopped abruptly (no fade)
    levels, events = [], []
    parts = []
    for _ in range(5):
        parts.append(p.render(40 * BLOCK))
        levels.append(p.audio_level())
        events += p.poll_status()
    assert p.poll_status() == []
    return np.concatenate(parts), events, levels


def test_oracle_status_stream_and_meter(oracle_api):
    out, events, levels = scene_observability(oracle_api)
    kinds = [(e[1], e[2]) for e in events]
    assert kinds.count(("stopped", 1)) == 1 and kinds.count(("stopped", 2)) == 1 and kinds.count(("stopped", 3)) == 1
    stopped = {e[2]: e for e in events if e[1] == "stopped"}
    assert stopped[1][3] is True and stopped[2][3] is False and stopped[3][3] is False and stopped[3][0] == 70000
    pos1 = [e for e in events if e[1] == "position" and e[2] == 1]
    assert len(pos1) == 3 and all(b[0] - a[0] >= SR for a, b in zip(pos1, pos1[1:]))
    assert all(b[3] > a[3] for a, b in zip(pos1, pos1[1:]))
    (peak, rms) = levels[0]
    assert 0.05 < max(peak) < 1.5 and 0.0 < max(rms) < max(peak)


@pytest.mark.gpu
def test_status_stream_and_meter_match_oracle(cuda_api, oracle_api):
    g_out, g_ev, g_lv = scene_observability(cuda_api)
    o_out, o_ev, o_lv = scene_observability(oracle_api)
    assert g_ev == o_ev                     # frames, ids, positions (ns) and exhausted flags: integers, exact
    assert float(np.abs(g_out - o_out).max()) <= 1e-5
    for (gp, gr), (op, orr) in zip(g_lv, o_lv):
        assert np.allclose(gp, op, rtol=0, atol=1e-5) and np.allclose(gr, orr, rtol=1e-5, atol=1e-7)
