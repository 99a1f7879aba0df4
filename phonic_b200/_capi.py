"""ctypes binding of the C-ABI declared in include/phonic_b200.h.

The binding is parametrised by (shared library path, symbol prefix) so that the very same
Python-side scene description can drive any implementation of that ABI. The product uses it
with ``libphonic_b200.so`` / ``pb200_``; nothing in this package refers to the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

U32, U64, I64, I32, F32, F64 = C.c_uint32, C.c_uint64, C.c_int64, C.c_int32, C.c_float, C.c_double

TIME_NOW = 0xFFFFFFFFFFFFFFFF
MAIN_MIXER = 0
REPEAT_DEFAULT = 0xFFFFFFFFFFFFFFFE
REPEAT_FOREVER = 0xFFFFFFFFFFFFFFFF
NO_LOOP = -1
DURATION_NONE = 0xFFFFFFFFFFFFFFFF

# pb200_error
OK = 0
ERR_SOURCE_NOT_PLAYING = 1
ERR_RESAMPLING = 7
ERR_GENERATOR_NOT_FOUND = 8
ERR_EFFECT_NOT_FOUND = 9
ERR_MIXER_NOT_FOUND = 10
ERR_PARAMETER = 11
ERR_SEND = 12
ERR_MEDIA_FILE_NOT_FOUND, ERR_MEDIA_FILE_PROBE, ERR_AUDIO_DECODING, ERR_IO = 2, 3, 5, 13
ERR_CUDA = 100
ERR_UNSUPPORTED = 101

# pb200_effect_kind
FX_FILTER, FX_EQ5, FX_COMPRESSOR, FX_CHORUS, FX_DELAY, FX_REVERB, FX_GAIN, FX_PANNING, FX_GATE, FX_DISTORTION = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10

# pb200_event_kind
EV_STOP_SOURCE = 1
EV_SET_SOURCE_VOLUME = 2
EV_SET_SOURCE_PANNING = 3
EV_SET_SOURCE_SPEED = 4
EV_SEEK_SOURCE = 5
EV_NOTE_ON = 10
EV_NOTE_OFF = 11
EV_ALL_NOTES_OFF = 12
EV_SET_NOTE_SPEED = 13
EV_SET_NOTE_VOLUME = 14
EV_SET_NOTE_PANNING = 15
EV_SET_GENERATOR_PARAMETER = 16
EV_SET_GENERATOR_LOOP_RANGE = 17
EV_SET_EFFECT_PARAMETER = 20
EV_EFFECT_MESSAGE = 21
MSG_REVERB_RESET = 1
MOVE_DIRECTION, MOVE_START, MOVE_END = 0, 1, 2

EVF_NORMALIZED, EVF_HAS_VOLUME, EVF_HAS_PANNING, EVF_NOTE_FROM_BATCH, EVF_NO_RANGE = 1, 2, 4, 8, 16


class Config(C.Structure):
    _fields_ = [("sample_rate", U32), ("channel_count", U32), ("block_frames", U32),
                ("device_ordinal", I32), ("master_volume", F32), ("reserved", U32 * 3)]


class FilterParams(C.Structure):
    _fields_ = [("filter_type", U32), ("cutoff", F32), ("q", F32)]


class CompressorParams(C.Structure):
    _fields_ = [("threshold", F32), ("ratio", F32), ("knee", F32), ("attack_time", F32),
                ("release_time", F32), ("makeup_gain", F32), ("lookahead_time", F32)]


class ChorusParams(C.Structure):
    _fields_ = [("rate", F32), ("phase", F32), ("depth", F32), ("feedback", F32), ("delay", F32),
                ("wet", F32), ("filter_type", U32), ("filter_freq", F32), ("filter_resonance", F32)]


class GainParams(C.Structure):
    _fields_ = [("gain_db", F32), ("dc_filter_mode", U32)]


class DistortionParams(C.Structure):
    _fields_ = [("distortion_type", U32), ("drive", F32), ("mix", F32)]


class GateParams(C.Structure):
    _fields_ = [("threshold", F32), ("attack_time", F32), ("hold_time", F32), ("release_time", F32), ("range", F32)]


class ReverbParams(C.Structure):
    _fields_ = [("room_size", F32), ("wet", F32), ("fpd", U32 * 2), ("vib_phase", F64 * 16)]


class FileOptions(C.Structure):
    _fields_ = [("volume", F32), ("panning", F32), ("speed", F64), ("repeat", U64),
                ("loop_start", I64), ("loop_end", I64), ("fade_in_nanos", U64),
                ("fade_out_nanos", U64), ("resampling_quality", U32), ("target_mixer", U32)]


class Ahdsr(C.Structure):
    _fields_ = [("attack_nanos", U64), ("hold_nanos", U64), ("decay_nanos", U64),
                ("release_nanos", U64), ("attack_scaling", F32), ("decay_scaling", F32),
                ("release_scaling", F32), ("sustain_level", F32)]


class Granular(C.Structure):
    _fields_ = [("overlap_mode", U32), ("window", U32), ("size", F32), ("density", F32), ("variation", F32),
                ("spray", F32), ("pan_spread", F32), ("playback_direction", U32), ("position", F32), ("step", F32)]


class SamplerOptions(C.Structure):
    _fields_ = [("volume", F32), ("panning", F32), ("voices", U32), ("target_mixer", U32),
                ("transient", U32), ("has_ahdsr", U32), ("ahdsr", Ahdsr),
                ("has_granular", U32), ("granular", Granular), ("reserved", U32)]


class Event(C.Structure):
    _fields_ = [("sample_time", U64), ("kind", U32), ("target", U32), ("note_id", U64),
                ("note", U32), ("param_id", U32), ("value", F32), ("value2", F32), ("glide", F32),
                ("flags", U32), ("speed", F64), ("position_nanos", U64)]


def event_dtype():
    """numpy structured dtype with the memory layout of pb200_event (for pb200_schedule_many from an array)."""
    import numpy as np
    dt = np.dtype([("sample_time", "<u8"), ("kind", "<u4"), ("target", "<u4"), ("note_id", "<u8"), ("note", "<u4"),
                   ("param_id", "<u4"), ("value", "<f4"), ("value2", "<f4"), ("glide", "<f4"), ("flags", "<u4"),
                   ("speed", "<f8"), ("position_nanos", "<u8")], align=True)
    assert dt.itemsize == C.sizeof(Event)
    return dt


class SourceStatus(C.Structure):
    _fields_ = [("is_playing", U32), ("exhausted", U32), ("end_frame", U64), ("playback_pos", U64)]


class VoiceState(C.Structure):
    _fields_ = [("note_id", U64), ("playback_pos", U64), ("envelope_stage", U32), ("active", U32)]


class RenderStats(C.Structure):
    _fields_ = [("device_ms", F64), ("voice_kernel_ms", F64), ("skeleton_kernel_ms", F64), ("effect_kernel_ms", F64),
                ("kernel_launches", U64), ("voice_frames", U64), ("sinc_kernel_ms", F64), ("grain_kernel_ms", F64),
                ("sinc_frames", U64), ("grain_samples", U64)]


# every symbol include/phonic_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
_R = C.c_void_p
class StatusEvent(C.Structure):
    _fields_ = [("frame", U64), ("kind", U32), ("playback_id", U32), ("position_nanos", U64), ("exhausted", U32), ("reserved", U32)]


class AudioLevel(C.Structure):
    _fields_ = [("peak", F32 * 2), ("rms", F32 * 2)]


class WavInfo(C.Structure):
    _fields_ = [("frames", U64), ("channels", U32), ("sample_rate", U32), ("loop_start", I64), ("loop_end", I64),
                ("bits_per_sample", U32), ("is_float", U32)]


SYMBOLS = {
    "create": (C.c_int, [_P(Config), _P(_R)]),
    "destroy": (None, [_R]),
    "last_error": (C.c_char_p, [_R]),
    "backend": (C.c_char_p, []),
    "upload_buffer": (C.c_int, [_R, _P(F32), U64, U32, U32, I64, I64, C.c_int, _P(U32)]),
    "add_mixer": (C.c_int, [_R, U32, _P(U32)]),
    "add_effect": (C.c_int, [_R, U32, U32, C.c_void_p, C.c_size_t, _P(U32)]),
    "file_options_default": (None, [_P(FileOptions)]),
    "play_file": (C.c_int, [_R, U32, _P(FileOptions), U64, _P(U32)]),
    "sampler_options_default": (None, [_P(SamplerOptions)]),
    "add_sampler": (C.c_int, [_R, U32, _P(SamplerOptions), U64, _P(U32)]),
    "schedule": (C.c_int, [_R, _P(Event)]),
    "render": (C.c_int, [_R, _P(F32), U64, _P(U64)]),
    "render_device": (C.c_int, [_R, C.c_void_p, U64, _P(U64)]),
    "position": (U64, [_R]),
    "source_status_get": (C.c_int, [_R, U32, _P(SourceStatus)]),
    "sampler_voice_states": (C.c_int, [_R, U32, _P(VoiceState), U32, _P(U32)]),
    "last_render_stats": (C.c_int, [_R, _P(RenderStats)]),
    "schedule_many": (C.c_int, [_R, _P(Event), U32, _P(U32)]),
    "decode_wav": (C.c_int, [C.c_char_p, _P(_P(F32)), _P(WavInfo)]),
    "free": (None, [C.c_void_p]),
    "upload_wav": (C.c_int, [_R, C.c_char_p, _P(U32), _P(WavInfo)]),
    "render_to_wav": (C.c_int, [_R, C.c_char_p, U64, _P(U64)]),
    "remove_source": (C.c_int, [_R, U32]),
    "remove_mixer": (C.c_int, [_R, U32]),
    "remove_effect": (C.c_int, [_R, U32]),
    "move_effect": (C.c_int, [_R, U32, U32, U32, I32]),
    "stop_all_sources": (C.c_int, [_R]),
    "poll_status": (C.c_int, [_R, _P(StatusEvent), U32, _P(U32)]),
    "set_metering_interval": (C.c_int, [_R, U64]),
    "get_audio_level": (C.c_int, [_R, _P(AudioLevel)]),
    "set_main_input": (C.c_int, [_R, C.c_void_p, U64]),
    "render_progress": (U64, [_R]),
    "set_main_inputs": (C.c_int, [_R, _P(C.c_void_p), U32, U64]),
    "trim_pool": (U64, [C.c_int]),
    "device_alloc": (C.c_int, [C.c_int, C.c_size_t, _P(C.c_void_p)]),
    "device_free": (C.c_int, [C.c_void_p]),
    "ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ipc_open": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_void_p)]),
    "ipc_close": (C.c_int, [C.c_void_p]),
    "push_async": (C.c_int, [_R, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, U32]),
    "push_sync": (C.c_int, [_R]),
    "peek_u32": (C.c_int, [_R, C.c_void_p, U32, _P(U32)]),
}


class CApi:
    """Loaded shared library with typed entry points ``api.<name>`` for every ABI symbol."""

    def __init__(self, path: str, prefix: str = "pb200_"):
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            if os.environ.get("PB200_LIB_LENIENT") and not hasattr(self.lib, prefix + name):
                continue  # A/B timing of an older build of the library (tools/ab_time.py)
            fn = getattr(self.lib, prefix + name)  # AttributeError => missing export
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)


def fourcc(s: str) -> int:
    b = s.encode("ascii")
    assert len(b) == 4, s
    return (b[0] << 24) | (b[1] << 16) | (b[2] << 8) | b[3]
