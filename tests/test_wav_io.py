"""WAV in / out (the steps either side of the path): the product's decoder and the oracle's
independently written one against a third, pure-numpy reading of the same files; `smpl` loops; the 32-bit float output
file. Decoding needs no GPU; the upload -> render -> file round trip is a GPU test."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from conftest import ROOT
from phonic_b200 import _capi as A
from phonic_b200.player import FilePlaybackOptions, Player

SR = 48000


def write_wav(path, data, rate, fmt, smpl=None, extensible=False, odd_chunk=True):
    """data: float64 [frames, ch] in [-1, 1). fmt: 'u8' | 's16' | 's24' | 's32' | 'f32' | 'f64'. Returns what a decoder must yield (f32)."""
    frames, ch = data.shape
    if fmt == "u8":
        q = np.clip(np.round(data * 128.0) + 128, 0, 255).astype(np.uint8); raw = q.tobytes(); expect = (q.astype(np.float32) - 128) / 128
        tag, bits = 1, 8
    elif fmt == "s16":
        q = np.clip(np.round(data * 32768.0), -32768, 32767).astype("<i2"); raw = q.tobytes(); expect = q.astype(np.float32) / 32768
        tag, bits = 1, 16
    elif fmt == "s24":
        q = np.clip(np.round(data * 8388608.0), -8388608, 8388607).astype(np.int32)
        raw = b"".join(struct.pack("<i", int(x))[:3] for x in q.reshape(-1)); expect = q.astype(np.float32) / 8388608
        tag, bits = 1, 24
    elif fmt == "s32":
        q = np.clip(np.round(data * 2147483648.0), -2147483648, 2147483647).astype("<i4"); raw = q.tobytes()
        expect = (q.astype(np.float64) / 2147483648.0).astype(np.float32)
        tag, bits = 1, 32
    elif fmt == "f32":
        q = data.astype("<f4"); raw = q.tobytes(); expect = q.astype(np.float32); tag, bits = 3, 32
    else:
        q = data.astype("<f8"); raw = q.tobytes(); expect = q.astype(np.float32); tag, bits = 3, 64
    align = ch * bits // 8
    if extensible:
        guid = struct.pack("<H", tag) + bytes([0, 0, 0, 0, 0x10, 0, 0x80, 0, 0, 0xAA, 0, 0x38, 0x9B, 0x71])
        fmt_body = struct.pack("<HHIIHHHHI", 0xFFFE, ch, rate, rate * align, align, bits, 22, bits, 3) + guid
    else:
        fmt_body = struct.pack("<HHIIHH", tag, ch, rate, rate * align, align, bits)
    chunks = b"fmt " + struct.pack("<I", len(fmt_body)) + fmt_body
    if odd_chunk:  # an unknown chunk of odd size: its pad byte must be skipped
        chunks += b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\0"
    chunks += b"data" + struct.pack("<I", len(raw)) + raw + (b"\0" if len(raw) & 1 else b"")
    if smpl:
        body = struct.pack("<9I", 0, 0, 0, 60, 0, 0, 0, 1, 0) + struct.pack("<6I", 0, 0, smpl[0], smpl[1], 0, 0)
        chunks += b"smpl" + struct.pack("<I", len(body)) + body
    open(path, "wb").write(b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks)
    return expect


def signal(frames, ch, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(frames)[:, None] / 44100.0
    return 0.4 * np.sin(2 * np.pi * (220.0 + 30 * np.arange(ch)[None, :]) * t) + 0.2 * rng.uniform(-1, 1, (frames, ch))


def decode(lib, prefix, path):
    fn = getattr(lib, prefix + "decode_wav")
    fn.restype = C.c_int
    fn.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(A.WavInfo)]
    free = getattr(lib, prefix + "free")
    free.argtypes = [C.c_void_p]
    free.restype = None
    p, info = C.POINTER(C.c_float)(), A.WavInfo()
    rc = fn(os.fsencode(path), C.byref(p), C.byref(info))
    if rc != 0:
        return rc, None, None
    a = np.ctypeslib.as_array(p, shape=(info.frames * info.channels,)).copy().reshape(info.frames, info.channels)
    free(p)
    return rc, a, info


@pytest.fixture(scope="module")
def libs(oracle_lib):
    product = C.CDLL(os.path.join(ROOT, "phonic_b200", "csrc", "libphonic_b200.so"))
    return [(product, "pb200_"), (oracle_lib, "po_")]


@pytest.mark.parametrize("fmt", ["u8", "s16", "s24", "s32", "f32", "f64"])
@pytest.mark.parametrize("ch", [1, 2])
def test_decoders_match_numpy_reading(libs, tmp_path, fmt, ch):
    path = str(tmp_path / f"t_{fmt}_{ch}.wav")
    expect = write_wav(path, signal(1237, ch, 3), 44100, fmt, extensible=(fmt in ("s24", "f32") and ch == 2))
    for lib, prefix in libs:
        rc, a, info = decode(lib, prefix, path)
        assert rc == 0
        assert (info.frames, info.channels, info.sample_rate) == (1237, ch, 44100)
        assert (info.loop_start, info.loop_end) == (A.NO_LOOP, A.NO_LOOP)
        assert np.array_equal(a, expect), f"{prefix}: {fmt} x{ch} differs"  # bit-exact: each value is one exact division


def test_smpl_loop_and_clamping(libs, tmp_path):
    data = signal(1000, 2, 4)
    for smpl, want in [((100, 900), (100, 900)), ((100, 5000), (100, 1001)), ((900, 100), (A.NO_LOOP, A.NO_LOOP)),
                       ((2000, 3000), (A.NO_LOOP, A.NO_LOOP))]:
        path = str(tmp_path / f"loop_{smpl[0]}_{smpl[1]}.wav")
        write_wav(path, data, 48000, "f32", smpl=smpl)
        for lib, prefix in libs:
            rc, a, info = decode(lib, prefix, path)
            # buffer.rs:105-115: both ends clamped to the frame count including the pad frame; kept only if end > start
            assert rc == 0 and (info.loop_start, info.loop_end) == want, (prefix, smpl)


def test_decode_errors_mirror_reference(libs, tmp_path):
    bad = str(tmp_path / "bad.wav")
    open(bad, "wb").write(b"RIFF\x04\0\0\0AVI ")
    empty = str(tmp_path / "empty.wav")
    write_wav(empty, np.zeros((0, 2)), 48000, "s16")
    adpcm = str(tmp_path / "adpcm.wav")
    write_wav(adpcm, signal(64, 1, 1), 48000, "s16")
    b = bytearray(open(adpcm, "rb").read()); b[20] = 2; open(adpcm, "wb").write(bytes(b))  # format tag 2 (ADPCM): unsupported
    for lib, prefix in libs:
        assert decode(lib, prefix, str(tmp_path / "missing.wav"))[0] == A.ERR_MEDIA_FILE_NOT_FOUND
        assert decode(lib, prefix, bad)[0] == A.ERR_MEDIA_FILE_PROBE
        assert decode(lib, prefix, adpcm)[0] == A.ERR_MEDIA_FILE_PROBE
        assert decode(lib, prefix, empty)[0] == A.ERR_AUDIO_DECODING


def read_f32_wav(path):
    d = open(path, "rb").read()
    assert d[:4] == b"RIFF" and d[8:16] == b"WAVEfmt " and struct.unpack("<I", d[4:8])[0] == len(d) - 8
    fmt_len = struct.unpack("<I", d[16:20])[0]
    tag, ch, rate, _, align, bits = struct.unpack("<HHIIHH", d[20:36])
    assert (tag, bits, fmt_len) == (0xFFFE, 32, 40) and d[44:46] == b"\x03\x00"  # extensible, IEEE float sub-format
    p = 20 + fmt_len
    assert d[p:p + 4] == b"data"
    n = struct.unpack("<I", d[p + 4:p + 8])[0]
    return np.frombuffer(d[p + 8:p + 8 + n], "<f4").reshape(-1, ch), rate


def test_oracle_wav_round_trip(oracle_api, tmp_path):
    """24-bit looping file in -> 2 s WAV out through the oracle; WavStream writes whole 1024-frame blocks (wav.rs:222)."""
    src = str(tmp_path / "in.wav")
    write_wav(src, signal(30000, 2, 5), 44100, "s24", smpl=(2000, 29000))
    p = Player(oracle_api, SR)
    bid, info = p.upload_wav(src)
    o = FilePlaybackOptions(volume=0.7)
    o.repeat_forever()
    p.play_file_source(bid, o)
    out = str(tmp_path / "out.wav")
    frames = p.render_to_wav(out, 2.0)
    p.close()
    assert frames == 94 * 1024  # first block count with whole-seconds(pos / 48000) >= 2
    audio, rate = read_f32_wav(out)
    assert rate == SR and audio.shape == (frames, 2) and np.abs(audio[-1024:]).max() > 0.01  # still looping at the end


@pytest.mark.gpu
def test_product_wav_round_trip_matches_oracle(cuda_api, oracle_api, tmp_path):
    src = str(tmp_path / "in.wav")
    write_wav(src, signal(30000, 2, 5), 44100, "s24", smpl=(2000, 29000))
    outs = []
    for name, api in (("gpu", cuda_api), ("ref", oracle_api)):
        p = Player(api, SR)
        bid, info = p.upload_wav(src)
        assert (info.frames, info.channels, info.loop_start, info.loop_end) == (30000, 2, 2000, 29000)
        o = FilePlaybackOptions(volume=0.7)
        o.repeat_forever()
        p.play_file_source(bid, o)
        out = str(tmp_path / f"out_{name}.wav")
        assert p.render_to_wav(out, 2.0) == 94 * 1024
        p.close()
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1]  # header and every sample byte: the cubic file path is bit-exact
