"""Host-side mirror of phonic's `Player` + handle API for the offline (WAV-output) path.

Same names, argument meaning and error behaviour as the reference (src/player.rs:274-1046,
src/player/handles/*.rs); every call forwards to the C-ABI of include/phonic_b200.h. Durations
are `float` seconds here and cross the boundary as integer nanoseconds (std::time::Duration).
"""
from __future__ import annotations

import os

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _capi as A


class PhonicError(RuntimeError):
    """phonic::Error (src/error.rs:8-22) carried across the C-ABI as a code + message."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


def _nanos(seconds: Optional[float]) -> int:
    """Duration::from_secs_f64-like conversion to Duration::as_nanos()."""
    if seconds is None:
        return A.DURATION_NONE
    return int(round(float(seconds) * 1e9))


@dataclass
class FilePlaybackOptions:
    """FilePlaybackOptions (src/source/file.rs:34-218)."""
    volume: float = 1.0
    panning: float = 0.0
    speed: float = 1.0
    repeat: Optional[int] = None          # None | count | "forever" via repeat_forever()
    loop_range: Optional[tuple] = None
    fade_in: Optional[float] = None
    fade_out: Optional[float] = 0.05
    resampling_quality: int = 0           # ResamplingQuality::Default
    target_mixer: int = A.MAIN_MIXER

    def repeat_forever(self):
        self.repeat = A.REPEAT_FOREVER
        return self


@dataclass
class AhdsrParameters:
    """AhdsrParameters::new / new_with_scaling (src/utils/ahdsr.rs:50-98), times in seconds."""
    attack: float = 0.01
    hold: float = 1.0
    decay: float = 0.5
    sustain: float = 0.75
    release: float = 1.0
    attack_scaling: float = 0.0
    decay_scaling: float = 0.0
    release_scaling: float = 0.0


@dataclass
class GranularParameters:
    """GranularParameters (src/generator/sampler/granular.rs:239-284); Sampler::with_granular_playback."""
    overlap_mode: int = 0         # GrainOverlapMode::Cloud
    window: int = 2               # GrainWindowMode::Triangle
    size: float = 100.0
    density: float = 10.0
    variation: float = 0.0
    spray: float = 0.0
    pan_spread: float = 0.0
    playback_direction: int = 0   # GrainPlaybackDirection::Forward
    position: float = 0.5
    step: float = 0.0


@dataclass
class GeneratorPlaybackOptions:
    """GeneratorPlaybackOptions (src/generator.rs:41-141)."""
    volume: float = 1.0
    panning: float = 0.0
    voices: int = 8
    target_mixer: int = A.MAIN_MIXER


@dataclass
class FilterEffect:
    """FilterEffect::with_parameters (src/effect/filter.rs:104-116); None => FilterEffect::new()."""
    filter_type: int = 0
    cutoff: float = 20000.0
    q: float = 0.707
    default: bool = False


@dataclass
class Eq5Effect:
    """Eq5Effect::new() (src/effect/eq5.rs:153-170)."""


@dataclass
class CompressorEffect:
    """CompressorEffect::with_compressor_parameters (src/effect/compressor.rs:122-140)."""
    threshold: float = -12.0
    ratio: float = 8.0
    knee: float = 3.0
    attack_time: float = 0.02
    release_time: float = 2.0
    makeup_gain: float = 6.0
    lookahead_time: float = 0.04

    @staticmethod
    def new_limiter():  # compressor.rs:114-157
        return CompressorEffect(-0.01, 20.0, 0.0, 0.02, 2.0, 0.0, 0.02)


@dataclass
class ChorusEffect:
    """ChorusEffect::with_parameters (src/effect/chorus.rs:178-200)."""
    rate: float = 1.0
    phase: float = math.pi / 2.0
    depth: float = 0.25
    feedback: float = 0.5
    delay: float = 12.0
    wet: float = 0.5
    filter_type: int = 0
    filter_freq: float = 20000.0
    filter_resonance: float = 0.0


@dataclass
class DelayEffect:
    """DelayEffect::new() (src/effect/delay.rs:180-212)."""


@dataclass
class ReverbEffect:
    """ReverbEffect::with_parameters (src/effect/reverb.rs:153-159) with the RNG-drawn start
    state (fpd_l/r, 16 vibrato phases, reverb.rs:95-103,535-538) made explicit."""
    room_size: float = 0.6
    wet: float = 0.35
    fpd: Sequence[int] = (0x1234567, 0x89ABCDE)
    vib_phase: Sequence[float] = field(default_factory=lambda: [0.1 + 0.37 * i for i in range(16)])


@dataclass
class GainEffect:
    """GainEffect::with_parameters(gain_db, dc_mode) (src/effect/gain.rs:97-104); default => GainEffect::new()."""
    gain_db: float = 0.0
    dc_filter_mode: int = 0       # GainEffectDcFilterMode: 0 Off 1 Slow 2 Default 3 Fast
    default: bool = False


@dataclass
class GateEffect:
    """GateEffect::with_parameters (src/effect/gate.rs:67-81): dB, s, s, s, dB; default => GateEffect::new()."""
    threshold: float = -30.0
    attack_time: float = 0.005
    hold_time: float = 0.1
    release_time: float = 0.2
    range: float = -60.0
    default: bool = False


@dataclass
class DistortionEffect:
    """DistortionEffect::with_parameters(type, drive, mix) (src/effect/distortion.rs:247-253); type: 0 SoftClip 1 HardClip
    2 Diode 3 Fuzz 4 Fold; default => DistortionEffect::new() (Diode, drive 0, mix 1)."""
    distortion_type: int = 2
    drive: float = 0.0
    mix: float = 1.0
    default: bool = False


@dataclass
class PanningEffect:
    """PanningEffect::new() (src/effect/pan.rs:52-60); parameters 'pan ', 'wdth', 'invl', 'invr' via set_parameter."""


class BatchNote:
    """Placeholder for the NotePlaybackId of a note_on queued inside `Player.batch()`: `index` = position of the NOTE_ON in
    the batch; `id` is the real id once the batch has been flushed."""
    __slots__ = ("index", "id")

    def __init__(self, index: int):
        self.index = index
        self.id = None

    def __int__(self):
        if self.id is None:
            raise RuntimeError("note id is only known after the batch has been flushed")
        return self.id


class _Handle:
    def __init__(self, player: "Player", ident: int):
        self._p = player
        self.id = ident

    def _ev(self, kind, sample_time, **kw):
        ev = A.Event()
        ev.sample_time = A.TIME_NOW if sample_time is None else int(sample_time)
        ev.kind = kind
        ev.target = self.id
        for k, v in kw.items():
            if k == "note_id" and isinstance(v, BatchNote):
                if v.id is None:  # refers to a NOTE_ON of the open batch
                    ev.note_id = v.index
                    ev.flags |= A.EVF_NOTE_FROM_BATCH
                else:
                    ev.note_id = v.id
            else:
                setattr(ev, k, v)
        batch = self._p._batch
        if batch is not None:
            batch.append(ev)
            return ev
        self._p._check(self._p.api.schedule(self._p._r, C.byref(ev)))
        return ev


class FilePlaybackHandle(_Handle):
    """FilePlaybackHandle (src/player/handles/file.rs:31-268)."""

    def stop(self, stop_time=None):
        self._ev(A.EV_STOP_SOURCE, stop_time)

    def seek(self, position_seconds: float, sample_time=None):
        self._ev(A.EV_SEEK_SOURCE, sample_time, position_nanos=_nanos(position_seconds))

    def set_speed(self, speed: float, glide: Optional[float] = None, sample_time=None):
        self._ev(A.EV_SET_SOURCE_SPEED, sample_time, speed=speed, glide=glide or 0.0)

    def set_volume(self, volume: float, sample_time=None):
        self._ev(A.EV_SET_SOURCE_VOLUME, sample_time, value=volume)

    def set_panning(self, panning: float, sample_time=None):
        self._ev(A.EV_SET_SOURCE_PANNING, sample_time, value=panning)

    def status(self) -> A.SourceStatus:
        st = A.SourceStatus()
        self._p._check(self._p.api.source_status_get(self._p._r, self.id, C.byref(st)))
        return st

    def is_playing(self) -> bool:
        return bool(self.status().is_playing)


class GeneratorPlaybackHandle(_Handle):
    """GeneratorPlaybackHandle (src/player/handles/generator.rs:62-437)."""

    def stop(self, stop_time=None):
        self._ev(A.EV_STOP_SOURCE, stop_time)

    def set_volume(self, volume: float, sample_time=None):
        self._ev(A.EV_SET_SOURCE_VOLUME, sample_time, value=volume)

    def set_panning(self, panning: float, sample_time=None):
        self._ev(A.EV_SET_SOURCE_PANNING, sample_time, value=panning)

    def note_on(self, note: int, volume: Optional[float] = None, panning: Optional[float] = None,
                sample_time=None) -> int:
        flags = (A.EVF_HAS_VOLUME if volume is not None else 0) | (A.EVF_HAS_PANNING if panning is not None else 0)
        ev = self._ev(A.EV_NOTE_ON, sample_time, note=int(note), value=volume or 0.0,
                      value2=panning or 0.0, flags=flags)
        if self._p._batch is not None:
            note = BatchNote(len(self._p._batch) - 1)
            self._p._batch_notes.append(note)
            return note
        return int(ev.note_id)

    def note_off(self, note_id: int, sample_time=None):
        self._ev(A.EV_NOTE_OFF, sample_time, note_id=note_id)

    def all_notes_off(self, sample_time=None):
        self._ev(A.EV_ALL_NOTES_OFF, sample_time)

    def set_note_speed(self, note_id: int, speed: float, glide: Optional[float] = None, sample_time=None):
        self._ev(A.EV_SET_NOTE_SPEED, sample_time, note_id=note_id, speed=speed, glide=glide or 0.0)

    def set_note_volume(self, note_id: int, volume: float, sample_time=None):
        self._ev(A.EV_SET_NOTE_VOLUME, sample_time, note_id=note_id, value=volume)

    def set_note_panning(self, note_id: int, panning: float, sample_time=None):
        self._ev(A.EV_SET_NOTE_PANNING, sample_time, note_id=note_id, value=panning)

    def set_parameter(self, param_id: str, value: float, sample_time=None):
        """GeneratorPlaybackHandle::set_parameter((id, value), t): 'STRN' 'SFTN' 'SVOL' 'SPAN' 'AATK' 'AHLD' 'ADCY' 'ASTN' 'AREL'"""
        self._ev(A.EV_SET_GENERATOR_PARAMETER, sample_time, param_id=A.fourcc(param_id), value=value)

    def set_parameter_normalized(self, param_id: str, value: float, sample_time=None):
        self._ev(A.EV_SET_GENERATOR_PARAMETER, sample_time, param_id=A.fourcc(param_id), value=value, flags=A.EVF_NORMALIZED)

    def set_loop_range(self, loop_range, sample_time=None):
        """send_message(SamplerMessage::SetLoopRange(range), t); `loop_range` = (start frame, end frame) or None"""
        if loop_range is None:
            self._ev(A.EV_SET_GENERATOR_LOOP_RANGE, sample_time, flags=A.EVF_NO_RANGE)
        else:
            self._ev(A.EV_SET_GENERATOR_LOOP_RANGE, sample_time, position_nanos=int(loop_range[0]), note_id=int(loop_range[1]))

    def voice_states(self, capacity: int = 1024):
        arr = (A.VoiceState * capacity)()
        n = A.U32(0)
        self._p._check(self._p.api.sampler_voice_states(self._p._r, self.id, arr, capacity, C.byref(n)))
        return [(int(v.note_id), int(v.playback_pos), int(v.envelope_stage), int(v.active))
                for v in arr[:min(n.value, capacity)]]


class EffectHandle(_Handle):
    """EffectHandle (src/player/handles/effect.rs:47-126)."""

    def set_parameter(self, param_id: str, value: float, sample_time=None):
        self._ev(A.EV_SET_EFFECT_PARAMETER, sample_time, param_id=A.fourcc(param_id), value=value)

    def set_parameter_normalized(self, param_id: str, value: float, sample_time=None):
        self._ev(A.EV_SET_EFFECT_PARAMETER, sample_time, param_id=A.fourcc(param_id), value=value,
                 flags=A.EVF_NORMALIZED)

    def send_message(self, message: int, sample_time=None):
        """EffectHandle::send_message (handles/effect.rs:127-163); message = _capi.MSG_REVERB_RESET"""
        self._ev(A.EV_EFFECT_MESSAGE, sample_time, param_id=message)


class MixerHandle:
    """MixerHandle (src/player/handles/mixer.rs:37-72)."""

    def __init__(self, ident: int):
        self.id = ident


class Player:
    """Player::new(WavOutput::open_with_specs(path, sample_rate, 2, duration), None)."""

    def __init__(self, api: A.CApi, sample_rate: int = 48000, channel_count: int = 2,
                 device_ordinal: int = -1, master_volume: float = 1.0, block_frames: int = 1024):
        self.api = api
        self.sample_rate = sample_rate
        self.channel_count = channel_count
        self.block_frames = block_frames
        cfg = A.Config(sample_rate, channel_count, block_frames, device_ordinal, master_volume)
        r = C.c_void_p()
        code = api.create(C.byref(cfg), C.byref(r))
        if code != A.OK:
            raise PhonicError(code, "pb200_create failed")
        self._r = r
        self._batch = None       # events collected by batch()
        self._batch_notes = []

    def close(self):
        if self._r:
            self.api.destroy(self._r)
            self._r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, code: int):
        if code != A.OK:
            msg = self.api.last_error(self._r)
            raise PhonicError(code, msg.decode() if msg else "")

    # -- sample data -------------------------------------------------------------------------
    def upload_buffer(self, samples: np.ndarray, sample_rate: int, loop_range=None, add_pad_frame=True) -> int:
        """AudioFileBuffer::new; `samples` is [frames] (mono) or [frames, channels] float32."""
        a = np.ascontiguousarray(samples, dtype=np.float32)
        ch = 1 if a.ndim == 1 else a.shape[1]
        frames = a.shape[0]
        ls, le = loop_range if loop_range else (A.NO_LOOP, A.NO_LOOP)
        bid = A.U32()
        self._check(self.api.upload_buffer(self._r, a.ctypes.data_as(C.POINTER(A.F32)), frames, ch, sample_rate,
                                           ls, le, 1 if add_pad_frame else 0, C.byref(bid)))
        return bid.value

    def batch(self):
        """Context manager: handle calls made inside are collected and queued with ONE pb200_schedule_many call on exit
        (same order, same result as the individual calls). note_on returns a BatchNote placeholder usable for the note's
        later events inside or after the batch."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            if self._batch is not None:
                raise RuntimeError("batches do not nest")
            self._batch, self._batch_notes = [], []
            try:
                yield self
            except BaseException:
                self._batch, self._batch_notes = None, []   # the body failed: nothing of the partial batch is queued
                raise
            else:
                evs, notes = self._batch, self._batch_notes
                self._batch, self._batch_notes = None, []
                if evs:
                    arr = (A.Event * len(evs))(*evs)
                    done = A.U32()
                    self._check(self.api.schedule_many(self._r, arr, len(evs), C.byref(done)))
                    for n in notes:
                        n.id = int(arr[n.index].note_id)
        return cm()

    def schedule_array(self, events: np.ndarray) -> int:
        """Queue a whole score held in a numpy array of `_capi.event_dtype()` records with ONE pb200_schedule_many call
        (the array is updated in place: NOTE_ON ids are written back, batch references resolved). Returns the count."""
        assert events.dtype == A.event_dtype() and events.flags.c_contiguous
        done = A.U32()
        self._check(self.api.schedule_many(self._r, events.ctypes.data_as(C.POINTER(A.Event)), len(events), C.byref(done)))
        return done.value

    def upload_wav(self, path: str):
        """AudioFileBuffer::from_file for a RIFF/WAVE file (src/source/file/buffer.rs:64-119). Returns (buffer id, WavInfo)."""
        bid, info = A.U32(), A.WavInfo()
        self._check(self.api.upload_wav(self._r, os.fsencode(path), C.byref(bid), C.byref(info)))
        return bid.value, info

    def render_to_wav(self, path: str, seconds: float) -> int:
        """Player::new(WavOutput::open_with_specs(path, sr, 2, Duration::from_secs_f64(seconds))) (src/output/wav.rs:50-120):
        renders the whole duration and writes a 32-bit float WAV; returns the frames written."""
        written = A.U64()
        self._check(self.api.render_to_wav(self._r, os.fsencode(path), int(round(seconds * 1e9)), C.byref(written)))
        return written.value

    # -- graph -------------------------------------------------------------------------------
    def add_mixer(self, parent_mixer_id: Optional[int] = None) -> MixerHandle:
        mid = A.U32()
        self._check(self.api.add_mixer(self._r, parent_mixer_id or A.MAIN_MIXER, C.byref(mid)))
        return MixerHandle(mid.value)

    def remove_mixer(self, mixer_id: int):
        """Player::remove_mixer (src/player.rs:825-868)"""
        self._check(self.api.remove_mixer(self._r, mixer_id))

    def remove_generator(self, playback_id: int):
        """Player::remove_generator (src/player.rs:747-770)"""
        self._check(self.api.remove_source(self._r, playback_id))

    def remove_effect(self, effect_id: int):
        """Player::remove_effect (src/player.rs:977-991)"""
        self._check(self.api.remove_effect(self._r, effect_id))

    def move_effect(self, movement, effect_id: int, mixer_id: Optional[int] = None):
        """Player::move_effect (src/player.rs:942-974); movement = 'start' | 'end' | int offset (EffectMovement)"""
        kind, off = (A.MOVE_START, 0) if movement == "start" else (A.MOVE_END, 0) if movement == "end" else (A.MOVE_DIRECTION, int(movement))
        self._check(self.api.move_effect(self._r, effect_id, mixer_id or A.MAIN_MIXER, kind, off))

    def poll_status(self, capacity: int = 4096):
        """PlaybackStatusEvent stream (src/source/status.rs): [(frame, 'position', id, nanos) | (frame, 'stopped', id, exhausted)]"""
        buf = (A.StatusEvent * capacity)()
        n = A.U32()
        self._check(self.api.poll_status(self._r, buf, capacity, C.byref(n)))
        return [(e.frame, "position", e.playback_id, e.position_nanos) if e.kind == 0 else (e.frame, "stopped", e.playback_id, bool(e.exhausted))
                for e in buf[:n.value]]

    def set_metering_interval(self, seconds: Optional[float]):
        """PlayerConfig::metering_interval (src/player.rs:166,216)"""
        self._check(self.api.set_metering_interval(self._r, A.DURATION_NONE if seconds is None else int(round(seconds * 1e9))))

    def audio_level(self):
        """Player::audio_level (src/player.rs:464-471): ((peak L, peak R), (rms L, rms R)) of the main mixer's output"""
        lv = A.AudioLevel()
        self._check(self.api.get_audio_level(self._r, C.byref(lv)))
        return tuple(lv.peak), tuple(lv.rms)

    def stop_all_sources(self):
        """Player::stop_all_sources (src/player.rs:1012-1045)"""
        self._check(self.api.stop_all_sources(self._r))

    def add_effect(self, effect, mixer_id: Optional[int] = None) -> EffectHandle:
        eid = A.U32()
        mixer = mixer_id or A.MAIN_MIXER
        if isinstance(effect, FilterEffect):
            if effect.default:
                kind, p = A.FX_FILTER, None
            else:
                kind, p = A.FX_FILTER, A.FilterParams(effect.filter_type, effect.cutoff, effect.q)
        elif isinstance(effect, Eq5Effect):
            kind, p = A.FX_EQ5, None
        elif isinstance(effect, CompressorEffect):
            kind, p = A.FX_COMPRESSOR, A.CompressorParams(effect.threshold, effect.ratio, effect.knee,
                                                         effect.attack_time, effect.release_time,
                                                         effect.makeup_gain, effect.lookahead_time)
        elif isinstance(effect, ChorusEffect):
            kind, p = A.FX_CHORUS, A.ChorusParams(effect.rate, effect.phase, effect.depth, effect.feedback,
                                                 effect.delay, effect.wet, effect.filter_type,
                                                 effect.filter_freq, effect.filter_resonance)
        elif isinstance(effect, DelayEffect):
            kind, p = A.FX_DELAY, None
        elif isinstance(effect, ReverbEffect):
            kind = A.FX_REVERB
            p = A.ReverbParams(effect.room_size, effect.wet, (A.U32 * 2)(*effect.fpd), (A.F64 * 16)(*effect.vib_phase))
        elif isinstance(effect, GainEffect):
            kind, p = A.FX_GAIN, (None if effect.default else A.GainParams(effect.gain_db, effect.dc_filter_mode))
        elif isinstance(effect, PanningEffect):
            kind, p = A.FX_PANNING, None
        elif isinstance(effect, GateEffect):
            kind = A.FX_GATE
            p = None if effect.default else A.GateParams(effect.threshold, effect.attack_time, effect.hold_time,
                                                         effect.release_time, effect.range)
        elif isinstance(effect, DistortionEffect):
            kind = A.FX_DISTORTION
            p = None if effect.default else A.DistortionParams(effect.distortion_type, effect.drive, effect.mix)
        else:
            raise TypeError(effect)
        if p is None:
            self._check(self.api.add_effect(self._r, mixer, kind, None, 0, C.byref(eid)))
        else:
            self._check(self.api.add_effect(self._r, mixer, kind, C.cast(C.byref(p), C.c_void_p), C.sizeof(p), C.byref(eid)))
        return EffectHandle(self, eid.value)

    # -- sources -----------------------------------------------------------------------------
    def play_file_source(self, buffer_id: int, options: Optional[FilePlaybackOptions] = None,
                         start_time: Optional[int] = None) -> FilePlaybackHandle:
        o = options or FilePlaybackOptions()
        fo = A.FileOptions()
        self.api.file_options_default(C.byref(fo))
        fo.volume, fo.panning, fo.speed = o.volume, o.panning, o.speed
        fo.repeat = A.REPEAT_DEFAULT if o.repeat is None else int(o.repeat)
        if o.loop_range:
            fo.loop_start, fo.loop_end = o.loop_range
        fo.fade_in_nanos = _nanos(o.fade_in)
        fo.fade_out_nanos = _nanos(o.fade_out)
        fo.resampling_quality = o.resampling_quality
        fo.target_mixer = o.target_mixer
        pid = A.U32()
        self._check(self.api.play_file(self._r, buffer_id, C.byref(fo),
                                       A.TIME_NOW if start_time is None else int(start_time), C.byref(pid)))
        return FilePlaybackHandle(self, pid.value)

    def _sampler(self, buffer_id, options, ahdsr, transient, start_time, granular=None):
        o = options or GeneratorPlaybackOptions()
        so = A.SamplerOptions()
        self.api.sampler_options_default(C.byref(so))
        so.volume, so.panning, so.voices, so.target_mixer = o.volume, o.panning, o.voices, o.target_mixer
        so.transient = 1 if transient else 0
        if ahdsr is not None:
            so.has_ahdsr = 1
            so.ahdsr = A.Ahdsr(_nanos(ahdsr.attack), _nanos(ahdsr.hold), _nanos(ahdsr.decay), _nanos(ahdsr.release),
                               ahdsr.attack_scaling, ahdsr.decay_scaling, ahdsr.release_scaling, ahdsr.sustain)
        if granular is not None:
            g = granular
            so.has_granular = 1
            so.granular = A.Granular(g.overlap_mode, g.window, g.size, g.density, g.variation, g.spray, g.pan_spread,
                                     g.playback_direction, g.position, g.step)
        gid = A.U32()
        self._check(self.api.add_sampler(self._r, buffer_id, C.byref(so),
                                         A.TIME_NOW if start_time is None else int(start_time), C.byref(gid)))
        return GeneratorPlaybackHandle(self, gid.value)

    def play_generator(self, buffer_id: int, options=None, ahdsr: Optional[AhdsrParameters] = None,
                       start_time: Optional[int] = None, granular: Optional[GranularParameters] = None) -> GeneratorPlaybackHandle:
        """Player::play_generator(Sampler::from_file_source(..).with_ahdsr(..)[.with_granular_playback(..)], start_time)."""
        return self._sampler(buffer_id, options, ahdsr, True, start_time, granular)

    def add_generator(self, buffer_id: int, options=None, ahdsr: Optional[AhdsrParameters] = None,
                      mixer_id: Optional[int] = None, granular: Optional[GranularParameters] = None) -> GeneratorPlaybackHandle:
        """Player::add_generator(Sampler::from_file_source(..).with_ahdsr(..), mixer_id)."""
        o = options or GeneratorPlaybackOptions()
        if mixer_id is not None:
            o = GeneratorPlaybackOptions(o.volume, o.panning, o.voices, mixer_id)
        return self._sampler(buffer_id, o, ahdsr, False, None, granular)

    # -- render ------------------------------------------------------------------------------
    def render(self, frames: int) -> np.ndarray:
        """`frames` output frames as WavStream::process would write them (src/output/wav.rs:210-250)."""
        out = np.zeros((frames, self.channel_count), dtype=np.float32)
        self.render_into(out)
        return out

    def render_into(self, out: np.ndarray) -> int:
        written = A.U64()
        frames = out.shape[0]
        self._check(self.api.render(self._r, out.ctypes.data_as(C.POINTER(A.F32)), frames, C.byref(written)))
        return written.value

    def render_device(self, device_ptr: int, frames: int) -> int:
        written = A.U64()
        self._check(self.api.render_device(self._r, C.c_void_p(device_ptr), frames, C.byref(written)))
        return written.value

    def render_progress(self) -> int:
        """Output frames finalized so far over all render calls (callable from another thread while `render*` runs)."""
        return int(self.api.render_progress(self._r))

    def set_main_input(self, device_ptr: int | None, frames: int = 0) -> None:
        """Multi-GPU renders: the next render adds the stereo f32 bus at `device_ptr` (device memory; host memory for the
        oracle) to the main mixer's input -- the reduced output of the sub-mixers rendered on other ranks. None detaches."""
        self._check(self.api.set_main_input(self._r, C.c_void_p(device_ptr) if device_ptr else None, frames))

    def set_main_inputs(self, device_ptrs, frames: int) -> None:
        """Up to 8 external buses, added to the main mixer's input in the order given (one per rank of a sharded render)."""
        arr = (C.c_void_p * len(device_ptrs))(*[C.c_void_p(p) for p in device_ptrs])
        self._check(self.api.set_main_inputs(self._r, arr, len(device_ptrs), frames))

    def push_async(self, dst_peer: int, src_device: int, nbytes: int, flag_peer: int | None, flag_value: int) -> None:
        self._check(self.api.push_async(self._r, C.c_void_p(dst_peer), C.c_void_p(src_device), nbytes,
                                        C.c_void_p(flag_peer) if flag_peer else None, flag_value))

    def push_sync(self) -> None:
        self._check(self.api.push_sync(self._r))

    def peek_u32(self, src_device: int, count: int):
        out = (A.U32 * count)()
        self._check(self.api.peek_u32(self._r, C.c_void_p(src_device), count, out))
        return list(out)

    def output_sample_frame_position(self) -> int:
        return int(self.api.position(self._r))

    def last_render_stats(self) -> A.RenderStats:
        st = A.RenderStats()
        self._check(self.api.last_render_stats(self._r, C.byref(st)))
        return st
