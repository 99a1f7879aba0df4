"""Synthetic workloads of BASELINE.json's configs, built through the public
Player API so the same scene can be handed to any implementation of the C-ABI.

All randomness is seeded; all sample values are f32 in [-0.5, 0.5]; gains are <= 1/sqrt(voices)
so buses stay below 1.0.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .player import (AhdsrParameters, ChorusEffect, CompressorEffect, DelayEffect, Eq5Effect, FilePlaybackOptions,
                     FilterEffect, GeneratorPlaybackOptions, GranularParameters, Player, ReverbEffect)


def speed_from_note(note: int) -> float:
    """src/utils.rs:67-78"""
    return (440.0 * 2.0 ** ((note - 69.0) / 12.0)) / (440.0 * 2.0 ** ((60 - 69.0) / 12.0))


def synth_buffer(frames: int, sample_rate: int, seed: int = 1, channels: int = 1) -> np.ndarray:
    """Band-limited noise + sines in [-0.5, 0.5] (cfg2 'one 44.1 k mono buffer of 4 s', seed 1)."""
    rng = np.random.default_rng(seed)
    t = np.arange(frames, dtype=np.float64) / sample_rate
    out = np.zeros((frames, channels), dtype=np.float64)
    for c in range(channels):
        sig = np.zeros(frames)
        for k in range(6):
            f = 110.0 * (k + 1) * (1.0 + 0.01 * c)
            sig += rng.uniform(0.2, 1.0) / (k + 1) * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
        noise = rng.standard_normal(frames)
        # crude band limit: moving average over 8 samples
        kernel = np.ones(8) / 8.0
        noise = np.convolve(noise, kernel, mode="same")
        sig += 0.5 * noise
        sig *= 0.5 / np.max(np.abs(sig))
        out[:, c] = sig
    out32 = out.astype(np.float32)
    return out32[:, 0].copy() if channels == 1 else out32


@dataclass
class VoiceBankSpec:
    """cfg2-style sampler voice bank."""
    voices: int = 256
    voices_per_sampler: int = 8
    seconds: float = 10.0
    sample_rate: int = 48000
    buffer_rate: int = 44100
    buffer_seconds: float = 4.0
    seed: int = 2
    glide: bool = True
    note_off: bool = True


def frames_for(seconds: float, sample_rate: int, block: int = 1024) -> int:
    """WavStream renders whole 1024-frame blocks until whole-seconds(pos/sr) >= duration
    (src/output/wav.rs:222): a '10 s' render at 48 kHz is 469 blocks = 480 256 frames."""
    blocks = 0
    while (blocks * block) // sample_rate < seconds:
        blocks += 1
    return blocks * block


_SCORES: dict = {}


def voice_bank_score(spec: VoiceBankSpec, seed_offset: int = 0, time_scale: float = 1.0):
    """The seeded random score of a voice bank: per sampler a list of (note, t_on, pan, glide | None, t_off | None)
    with glide = (t_gl, target_note, rate). Cached: drawing it is workload synthesis, not part of the render path."""
    key = (tuple(sorted(spec.__dict__.items())), seed_offset, time_scale)
    if key in _SCORES:
        return _SCORES[key]
    rng = np.random.default_rng(spec.seed + seed_offset)
    sr = spec.sample_rate
    n_samplers = (spec.voices + spec.voices_per_sampler - 1) // spec.voices_per_sampler
    remaining = spec.voices
    score = []
    for _ in range(n_samplers):
        nv = min(spec.voices_per_sampler, remaining)
        remaining -= nv
        voices = []
        for _v in range(nv):
            note = int(rng.integers(36, 85))
            t_on = int(rng.uniform(0.0, 2.0 * time_scale) * sr)
            pan = float(rng.uniform(-0.8, 0.8))
            glide = None
            if spec.glide:
                t_gl = int(rng.uniform(2.0, 6.0) * time_scale * sr)
                tgt = note + (7 if rng.random() < 0.5 else -7)
                rate = float(rng.uniform(12.0, 60.0))
                glide = (t_gl, tgt, rate)
            t_off = int(rng.uniform(6.0, 8.0) * time_scale * sr) if spec.note_off else None
            voices.append((note, t_on, pan, glide, t_off))
        score.append(voices)
    _SCORES[key] = score
    return score


def add_voice_bank(player: Player, spec: VoiceBankSpec, buffer_id: int, mixer_id=None, seed_offset: int = 0,
                   time_scale: float = 1.0):
    """Adds `spec.voices` sampler voices (AHDSR + glide) to `mixer_id`; returns the sampler handles.

    Per voice: note uniform in 36..84, note-on uniform in [0, 2 s), glide SetSpeed to note+-7 at
    [2, 6 s) with 12..60 st/s, note-off at [6, 8 s). Times scale with `time_scale` for short tests.
    """
    gain = 1.0 / math.sqrt(max(spec.voices, 1))
    ahdsr = AhdsrParameters(attack=0.01, hold=0.0, decay=0.5, sustain=0.75, release=1.0)
    handles = []
    for voices in voice_bank_score(spec, seed_offset, time_scale):
        opts = GeneratorPlaybackOptions(volume=1.0, panning=0.0, voices=len(voices))
        h = player.add_generator(buffer_id, opts, ahdsr, mixer_id=mixer_id)
        handles.append(h)
        for note, t_on, pan, glide, t_off in voices:
            nid = h.note_on(note, volume=gain, panning=pan, sample_time=t_on)
            if glide is not None:
                h.set_note_speed(nid, speed_from_note(glide[1]), glide=glide[2], sample_time=glide[0])
            if t_off is not None:
                h.note_off(nid, sample_time=t_off)
    return handles


_SCORE_ARRAYS: dict = {}


def voice_bank_events(spec: VoiceBankSpec, seed_offset: int = 0, time_scale: float = 1.0):
    """The same score as `voice_bank_score`, as one array of pb200_event records in the order add_voice_bank queues
    them (note-on, glide, note-off per voice; note-addressed events refer to their NOTE_ON by batch index). `target`
    holds the sampler's index in the bank; `add_voice_bank_fast` replaces it with the generator id. Cached."""
    from . import _capi as A
    key = (tuple(sorted(spec.__dict__.items())), seed_offset, time_scale)
    if key in _SCORE_ARRAYS:
        return _SCORE_ARRAYS[key]
    gain = 1.0 / math.sqrt(max(spec.voices, 1))
    rows = []
    for si, voices in enumerate(voice_bank_score(spec, seed_offset, time_scale)):
        for note, t_on, pan, glide, t_off in voices:
            on = len(rows)
            rows.append((t_on, A.EV_NOTE_ON, si, 0, note, 0, gain, pan, 0.0, A.EVF_HAS_VOLUME | A.EVF_HAS_PANNING, 0.0, 0))
            if glide is not None:
                rows.append((glide[0], A.EV_SET_NOTE_SPEED, si, on, 0, 0, 0.0, 0.0, glide[2], A.EVF_NOTE_FROM_BATCH, speed_from_note(glide[1]), 0))
            if t_off is not None:
                rows.append((t_off, A.EV_NOTE_OFF, si, on, 0, 0, 0.0, 0.0, 0.0, A.EVF_NOTE_FROM_BATCH, 0.0, 0))
    arr = np.array(rows, dtype=A.event_dtype())
    counts = [len(v) for v in voice_bank_score(spec, seed_offset, time_scale)]
    _SCORE_ARRAYS[key] = (arr, counts)
    return _SCORE_ARRAYS[key]


def add_voice_bank_fast(player: Player, spec: VoiceBankSpec, buffer_id: int, mixer_id=None, seed_offset: int = 0,
                        time_scale: float = 1.0):
    """add_voice_bank with the events queued by one pb200_schedule_many call from the cached event array (what a host
    program does with a score it has loaded): same graph, same events in the same order, same audio."""
    template, counts = voice_bank_events(spec, seed_offset, time_scale)
    ahdsr = AhdsrParameters(attack=0.01, hold=0.0, decay=0.5, sustain=0.75, release=1.0)
    handles = [player.add_generator(buffer_id, GeneratorPlaybackOptions(volume=1.0, panning=0.0, voices=nv), ahdsr, mixer_id=mixer_id)
               for nv in counts]
    events = template.copy()
    events["target"] = np.asarray([h.id for h in handles], dtype=np.uint32)[template["target"]]
    player.schedule_array(events)
    return handles


def build_cfg2(player: Player, spec: VoiceBankSpec | None = None, time_scale: float = 1.0, buffer=None, fast: bool = False):
    """cfg2: 256 Sampler voices (AHDSR + glide), cubic resampling, FilterEffect LP 2 kHz on the bus.
    `buffer`: the (already synthesised) host sample data; generated here when None."""
    spec = spec or VoiceBankSpec()
    buf = buffer if buffer is not None else synth_buffer(int(spec.buffer_seconds * spec.buffer_rate), spec.buffer_rate, seed=1)
    bid = player.upload_buffer(buf, spec.buffer_rate)
    handles = (add_voice_bank_fast if fast else add_voice_bank)(player, spec, bid, None, 0, time_scale)
    fx = player.add_effect(FilterEffect(0, 2000.0, 0.707))
    return handles, fx


def build_cfg1(player: Player, buffer: np.ndarray, buffer_rate: int):
    """cfg1: one stereo file through FilterEffect LP 1 kHz + ReverbEffect(0.6, 0.35)."""
    bid = player.upload_buffer(buffer, buffer_rate)
    h = player.play_file_source(bid, FilePlaybackOptions())
    player.add_effect(FilterEffect(0, 1000.0, 0.707))
    player.add_effect(ReverbEffect(0.6, 0.35))
    return h


def build_cfg1_asset(player: Player, wav_path: str, repeat_forever: bool = False):
    """cfg1 on a real asset (BASELINE configs[0]): the WAV file decoded by the implementation's own reader
    (AudioFileBuffer::from_file, src/source/file/buffer.rs:64-119), default FilePlaybackOptions (cubic resampling to the
    player's output rate), FilterEffect LP 1 kHz / Q 0.707 + ReverbEffect(0.6, 0.35) on the main mixer."""
    bid, info = player.upload_wav(wav_path)
    o = FilePlaybackOptions()
    if repeat_forever:
        o.repeat_forever()
    h = player.play_file_source(bid, o)
    player.add_effect(FilterEffect(0, 1000.0, 0.707))
    player.add_effect(ReverbEffect(0.6, 0.35))
    return h, info


def build_subtrees(player: Player, n_mixers: int, voices_per_mixer: int, spec: VoiceBankSpec, effects: str = "none",
                   time_scale: float = 1.0, seed_base: int = 0, buffer=None, fast: bool = False):
    """cfg3 / cfg5 shape: `n_mixers` sub-mixers of the main mixer, each with a cfg2-style bank.
    effects: 'none' | 'cfg3' (Eq5 + Compressor + Chorus per sub-mixer)."""
    buf = buffer if buffer is not None else synth_buffer(int(spec.buffer_seconds * spec.buffer_rate), spec.buffer_rate, seed=1)
    bid = player.upload_buffer(buf, spec.buffer_rate)
    rng = np.random.default_rng(3 + seed_base)
    out = []
    for m in range(n_mixers):
        mh = player.add_mixer(None)
        sub = VoiceBankSpec(**{**spec.__dict__, "voices": voices_per_mixer})
        hs = (add_voice_bank_fast if fast else add_voice_bank)(player, sub, bid, mh.id, seed_offset=1000 * (m + 1) + seed_base, time_scale=time_scale)
        if effects == "cfg3":
            eq = player.add_effect(Eq5Effect(), mh.id)
            for b in range(5):
                eq.set_parameter(f"gan{b + 1}", float(rng.uniform(-6.0, 6.0)), sample_time=0)
            player.add_effect(CompressorEffect(), mh.id)
            player.add_effect(ChorusEffect(), mh.id)
        out.append((mh, hs))
    return out


def add_main_bus_sends(player: Player):
    """cfg5: Delay (375 ms, fb 0.5 = DelayEffect::new() defaults) + Reverb on the main bus."""
    player.add_effect(DelayEffect())
    player.add_effect(ReverbEffect(0.6, 0.35))


def build_cfg4(player: Player, voices: int = 160, voices_per_sampler: int = 8, time_scale: float = 1.0, seed: int = 4,
               wav_path: str | None = None, buffer=None):
    """cfg4: granular synthesis, 16k grains/s: a pad-ambient.wav-shaped mono buffer (362 835 frames
    @ 48 kHz, loop 286 619..362 834), `voices` voices x density 100 Hz x size 100 ms, Hann, Forward, step 1.0, no
    randomisation. Notes as in cfg2 (on in [0, 2 s), off in [6, 8 s)), AHDSR (10 ms, 0, 500 ms, 0.75, 1 s)."""
    if wav_path is not None:  # the real asset (stereo float32 + smpl loop): mixed down on device like sampler.rs:908-952
        bid, _ = player.upload_wav(wav_path)
    else:
        frames = 362835
        buf = buffer if buffer is not None else synth_buffer(frames, 48000, seed=seed)   # (`buffer`: the same data, synthesised once)
        bid = player.upload_buffer(buf, 48000, loop_range=(286619, 362834))
    rng = np.random.default_rng(seed + 1)
    ahdsr = AhdsrParameters(attack=0.01, hold=0.0, decay=0.5, sustain=0.75, release=1.0)
    gran = GranularParameters(window=0, size=100.0, density=100.0, position=0.1, step=1.0)
    gain = 1.0 / math.sqrt(max(voices, 1)) / 3.0
    handles = []
    remaining = voices
    while remaining > 0:
        nv = min(voices_per_sampler, remaining)
        remaining -= nv
        h = player.add_generator(bid, GeneratorPlaybackOptions(voices=nv), ahdsr, granular=gran)
        handles.append(h)
        for _ in range(nv):
            note = int(rng.integers(48, 73))
            t_on = int(rng.uniform(0.0, 2.0 * time_scale) * 48000)
            nid = h.note_on(note, volume=gain, panning=float(rng.uniform(-0.8, 0.8)), sample_time=t_on)
            h.note_off(nid, sample_time=int(rng.uniform(6.0, 8.0) * time_scale * 48000))
    return handles


def build_sinc_bank(player: Player, voices: int = 1024, buffer=None, seed: int = 5):
    """cfg4's sinc micro-benchmark: `voices` HighQuality file sources, 44.1 -> 48 kHz, looping over
    the whole 4 s buffer so that every voice resamples for the full render; start times spread over 100 ms."""
    buf = buffer if buffer is not None else synth_buffer(int(4.0 * 44100), 44100, seed=1)
    frames = buf.shape[0]
    bid = player.upload_buffer(buf, 44100)
    rng = np.random.default_rng(seed)
    gain = 1.0 / math.sqrt(max(voices, 1))
    hs = []
    for _ in range(voices):
        o = FilePlaybackOptions(volume=gain, panning=float(rng.uniform(-0.8, 0.8)), loop_range=(0, frames), resampling_quality=1)
        o.repeat_forever()
        hs.append(player.play_file_source(bid, o, start_time=int(rng.integers(0, 4800))))
    return hs
