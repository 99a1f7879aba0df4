"""Pins the oracle (CPU restatement) against every assertion the reference's own tests hold for
this path and the derived known-answer vectors of SURVEY.md Appendix D.
No GPU needed."""
import ctypes as C

import numpy as np
import pytest

F32P = C.POINTER(C.c_float)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def test_cubic_known_answer(oracle_lib):
    # Appendix D: mono [0.2,1.0,0.5,0.0] 44100->48000, ratio f32 = 0.918749988079071
    lib = oracle_lib
    lib.po_test_cubic.restype = C.c_uint64
    inp = f32([0.2, 1.0, 0.5, 0.0])
    out = np.zeros(1024, np.float32)
    consumed = C.c_uint64()
    ratio = C.c_float()
    n = lib.po_test_cubic(inp.ctypes.data_as(F32P), C.c_uint64(4), 1, 44100, 48000, out.ctypes.data_as(F32P),
                          C.c_uint64(1024), C.byref(consumed), C.byref(ratio))
    assert n == 3 and consumed.value == 4
    assert ratio.value == np.float32(0.918749988079071)
    assert out[:3].tolist() == [np.float32(0.2), np.float32(0.9777594804763794), np.float32(0.5956249237060547)]


def test_reference_unit_test_resampling(oracle_lib):
    # src/source/file/preloaded.rs:486-533 `resampling`
    lib = oracle_lib
    lib.po_test_preloaded_write.restype = C.c_uint64
    buf = f32([0.2, 1.0, 0.5, 0.0])
    out = np.zeros(1024, np.float32)
    written = lib.po_test_preloaded_write(buf.ctypes.data_as(F32P), C.c_uint64(4), 1, 44100, 48000,
                                          out.ctypes.data_as(F32P), C.c_uint64(1024))
    assert written >= 4 * 44100 // 48000
    assert abs(float(out.sum()) - float(buf.sum())) < 0.1
    # "HighQuality" half: equal rates => bypass copy
    buf = f32([0.2, 1.0, 0.5])
    out = np.zeros(1024, np.float32)
    written = lib.po_test_preloaded_write(buf.ctypes.data_as(F32P), C.c_uint64(3), 1, 48000, 48000,
                                          out.ctypes.data_as(F32P), C.c_uint64(1024))
    assert written >= 3 * 44100 // 48000
    assert abs(float(out.sum()) - float(buf.sum())) < 0.2
    assert float(out[3:].sum()) < 0.1


def test_biquad_lowpass_known_answer(oracle_lib):
    # Appendix D: LP fc=1000 Q=f32(0.707) at 44100 Hz (what FilterEffect::with_parameters uses)
    lib = oracle_lib
    co = (C.c_double * 6)()
    imp = np.zeros(4, np.float32)
    lib.po_test_biquad(0, 44100, C.c_float(1000.0), C.c_float(0.707), C.c_float(0.0), co, imp.ctypes.data_as(F32P), 4)
    assert co[0] == pytest.approx(0.90413974532964481, rel=1e-14)
    assert co[1] == pytest.approx(0.064518219546102942, rel=1e-14)
    assert co[2] == pytest.approx(0.0046039350386941303, rel=1e-13)
    np.testing.assert_allclose(imp, f32([0.0046039349, 0.0174906794, 0.0323072597, 0.0438246652]), rtol=2e-7)
    lib.po_test_biquad(0, 48000, C.c_float(1000.0), C.c_float(0.707), C.c_float(0.0), co, imp.ctypes.data_as(F32P), 4)
    assert co[0] == pytest.approx(0.91157503639126503, rel=1e-14)
    assert co[2] == pytest.approx(0.0039160766917348353, rel=1e-13)


def _ahdsr(lib, n, off_at, a=0.01, h=1.0, d=0.5, s=0.75, r=1.0, sc=(0.0, 0.0, 0.0), sr=48000):
    out = np.zeros(n, np.float32)
    st = np.zeros(n, np.uint32)
    rates = np.zeros(3, np.float32)
    ns = lambda x: C.c_uint64(int(round(x * 1e9)))
    lib.po_test_ahdsr(ns(a), ns(h), ns(d), C.c_float(s), ns(r), C.c_float(sc[0]), C.c_float(sc[1]), C.c_float(sc[2]),
                      sr, n, off_at, out.ctypes.data_as(F32P), st.ctypes.data_as(C.POINTER(C.c_uint32)),
                      rates.ctypes.data_as(F32P))
    return out, st, rates


def test_ahdsr_reference_stage_transitions(oracle_lib):
    # src/utils/ahdsr.rs:588-664: note_on -> Attack, note_off -> Release, end -> Idle; Appendix D numbers
    out, st, rates = _ahdsr(oracle_lib, 200000, 100000)
    assert rates[0] == np.float32(0.0020833334419876337)
    assert st[0] == 1                      # Attack right after note_on
    assert int((st == 1).sum()) == 479     # Attack lasts 480 run() calls; the 480th switches to Hold
    assert st[479] == 2 and out[479] == 1.0
    assert st[100000] == 5                 # Release at note_off
    assert st[-1] == 0 and out[-1] == 0.0  # Idle at the end
    assert np.all(np.diff(out[:480]) > 0)  # attack is monotone
    rel = out[100000:100000 + 48000]
    assert np.all(np.diff(rel[: int((st[100000:] == 5).sum())]) < 0)


def test_ahdsr_scaling_monotone(oracle_lib):
    out, st, _ = _ahdsr(oracle_lib, 2000, 10**9, a=0.02, sc=(0.5, -0.5, 0.5))
    att = out[st == 1]
    assert np.all(np.diff(att) >= 0) and att.max() <= 1.0 + 1e-6


def test_exp_smoother_known_answer(oracle_lib):
    last = C.c_float()
    n = oracle_lib.po_test_exp_smoother(C.c_float(0.0), C.c_float(1.0), 48000, C.byref(last))
    assert n == 1588
    assert last.value == pytest.approx(0.99668497, abs=1e-7)
    # reference test: at 44100 the ramp is monotone without overshoot (smoothing.rs:556-600)
    n2 = oracle_lib.po_test_exp_smoother(C.c_float(0.0), C.c_float(1.0), 44100, C.byref(last))
    assert 0 < n2 < n and last.value <= 1.0 + 1e-4


def test_fader_known_answer(oracle_lib):
    inertia = C.c_float()
    n = oracle_lib.po_test_fader(C.c_uint64(50_000_000), 48000, C.byref(inertia))
    # NB: SURVEY Appendix D lists 0.00191694498 (numpy's SIMD f32 exp, 1 ulp off at 0.99808306);
    # the correctly rounded expf glibc/Rust use gives 0.9980831146 => inertia 0.00191688538.
    assert inertia.value == np.float32(0.0019168853759765625)
    assert n in (4801, 4802)


def test_panning_and_db(oracle_lib):
    lib = oracle_lib
    l, r = C.c_float(), C.c_float()
    lib.po_test_panning(C.c_float(0.0), C.byref(l), C.byref(r))
    assert (l.value, r.value) == (1.0, 1.0)
    lib.po_test_panning(C.c_float(-0.3), C.byref(l), C.byref(r))
    assert l.value == pytest.approx(1.1401754618, abs=1e-7) and r.value == pytest.approx(0.8366600275, abs=1e-7)
    lib.po_test_panning(C.c_float(1.0), C.byref(l), C.byref(r))
    assert l.value == 0.0 and r.value == pytest.approx(1.4142135382, abs=1e-7)
    # src/utils.rs:94-104
    lib.po_test_db_to_linear.restype = C.c_float
    lib.po_test_linear_to_db.restype = C.c_float
    assert lib.po_test_linear_to_db(C.c_float(1.0)) == 0.0
    assert lib.po_test_linear_to_db(C.c_float(0.0)) == -200.0
    assert lib.po_test_db_to_linear(C.c_float(-200.0)) == 0.0
    assert lib.po_test_db_to_linear(C.c_float(0.0)) == 1.0
    assert lib.po_test_linear_to_db(C.c_float(lib.po_test_db_to_linear(C.c_float(20.0)))) == pytest.approx(20.0, abs=1e-4)
    assert np.isnan(lib.po_test_db_to_linear(C.c_float(float("nan"))))
    assert np.isnan(lib.po_test_linear_to_db(C.c_float(-1.0)))


@pytest.mark.parametrize("note,speed,rate,ratio", [
    (58, 0.8908987181403394, 53878, 0.81851590),
    (60, 1.0, 48000, 0.91874999),
    (61, None, 45305, 0.97340250),
    (72, 2.0, 24000, 1.83749998),
])
def test_note_to_ratio(oracle_lib, note, speed, rate, ratio):
    s, r, q = C.c_double(), C.c_uint32(), C.c_float()
    oracle_lib.po_test_note_ratio(note, 44100, 48000, C.byref(s), C.byref(r), C.byref(q))
    if speed is not None:
        assert s.value == pytest.approx(speed, rel=1e-15)
    assert r.value == rate
    assert q.value == pytest.approx(ratio, abs=1e-7)


# ---- HighQuality (rubato sinc) path: PARITY UNPINNED against rubato itself (crate not available); these pin the
# ---- restatement against the reference's own unit test at this boundary and against signal-level properties.
def _player(oracle_api, sr=48000):
    from phonic_b200.player import Player
    return Player(oracle_api, sr)


def test_reference_unit_test_resampling_high_quality_half(oracle_api):
    # src/source/file/preloaded.rs:513-532: [0.2, 1.0, 0.5] @48000 -> 48000, ResamplingQuality::HighQuality
    from phonic_b200.player import FilePlaybackOptions
    p = _player(oracle_api)
    buf = f32([0.2, 1.0, 0.5])
    b = p.upload_buffer(buf, 48000, add_pad_frame=False)
    p.play_file_source(b, FilePlaybackOptions(resampling_quality=1))
    out = p.render(1024)[:, 0]
    assert np.count_nonzero(out) >= 3 * 44100 // 48000
    assert abs(float(out.sum()) - float(buf.sum())) < 0.2
    assert float(out[3:].sum()) < 0.1


def test_high_quality_resampler_reconstructs_a_sine(oracle_api):
    """44.1 -> 48 kHz: a 1 kHz sine comes out as a 1 kHz sine of the same amplitude; its sub-sample delay is the one
    the restated index arithmetic implies (first output at idx = -sinc_len/2 + t_ratio, filter centre 1/128 late)."""
    from phonic_b200.player import FilePlaybackOptions
    sr_in, sr_out, f = 44100, 48000, 1000.0
    x = (0.5 * np.sin(2 * np.pi * f * np.arange(20000) / sr_in)).astype(np.float32)
    p = _player(oracle_api, sr_out)
    b = p.upload_buffer(x, sr_in)
    p.play_file_source(b, FilePlaybackOptions(resampling_quality=1, fade_out=None))
    y = p.render(24 * 1024)[:, 0]
    seg = slice(2000, 16000)
    n = np.arange(len(y))[seg]
    w = 2 * np.pi * f / sr_out
    A = np.stack([np.sin(w * n), np.cos(w * n)], 1)
    c, *_ = np.linalg.lstsq(A, y[seg].astype(np.float64), rcond=None)
    assert abs(np.hypot(*c) - 0.5) < 2e-5
    assert np.abs(y[seg] - A @ c).max() < 2e-6
    delay_in_samples = -np.arctan2(c[1], c[0]) / (2 * np.pi * f) * sr_in
    t_ratio = sr_in / sr_out
    assert delay_in_samples == pytest.approx(1.0 - t_ratio - 1.0 / 128.0, abs=2e-3)


def test_high_quality_chunking(oracle_api):
    """256-frame input chunks. 1000 input frames = 3 chunks + a 232-frame tail; the tail comes up while output of the
    3rd chunk is still pending, and the zero-pad path counts it as consumed without processing it
    (preloaded.rs:296-304): only 768 input frames ever reach the resampler."""
    from phonic_b200.player import FilePlaybackOptions
    p = _player(oracle_api)
    b = p.upload_buffer(np.full(1000, 0.25, np.float32), 24000, add_pad_frame=False)
    h = p.play_file_source(b, FilePlaybackOptions(resampling_quality=1, fade_out=None))
    out = p.render(4 * 1024)[:, 0]
    # 2x up-sampling of a DC signal: DC gain 1 in the steady state (table normalised to F / sum)
    assert np.abs(out[400:1200] - 0.25).max() < 2e-6
    # 768 input frames -> 1536 output frames + half the filter length (256 output frames) of ring-down
    last = int(np.flatnonzero(np.abs(out) > 1e-7)[-1])
    assert 1536 < last < 1536 + 256
    # EOF is reached inside the 2nd 1024-frame call; the source is gone after it
    assert not h.is_playing() and not out[2048:].any()


# ---- granular playback (src/generator/sampler/granular.rs), derived from the cited formulas -------------------------
def _grain_window_lut(mode, N=2048):
    ph = (np.arange(N, dtype=np.float32) / np.float32(N)).astype(np.float32)
    pi = np.float32(np.pi)
    if mode == 2:   # Triangle (granular.rs:131-135)
        return np.where(ph < np.float32(0.5), np.float32(2) * ph, np.float32(2) * (np.float32(1) - ph)).astype(np.float32)
    if mode == 4:   # Trapezoid (granular.rs:152-160)
        rw = np.float32(0.1)
        return np.where(ph < rw, ph / rw, np.where(ph > np.float32(1) - rw, (np.float32(1) - ph) / rw, np.float32(1))).astype(np.float32)
    raise ValueError(mode)


def test_single_grain_is_the_window_times_the_sample(oracle_api):
    """One Triangle grain over a constant buffer: out = ((0.25 * (window(k / size) * volume)) * gain), f32, with the
    2048-point LUT lerp of GrainWindow::sample (granular.rs:201-215) and the 0.001 envelope threshold."""
    from phonic_b200.player import GeneratorPlaybackOptions, GranularParameters
    sr, size_ms, vol, pan = 48000, 100.0, np.float32(0.8), np.float32(-0.5)
    p = _player(oracle_api, sr)
    b = p.upload_buffer(np.full(4800, 0.25, np.float32), sr, add_pad_frame=False)
    g = p.add_generator(b, GeneratorPlaybackOptions(voices=1), None,
                        granular=GranularParameters(window=2, size=size_ms, density=1.0, position=0.5, step=0.0))
    g.note_on(60, float(vol), float(pan), sample_time=0)
    out = p.render(8 * 1024)
    S = int(np.float32(size_ms) * np.float32(1.0) * np.float32(sr) / np.float32(1000.0))
    assert S == 4800
    lut = _grain_window_lut(2)
    k = np.arange(S)
    phase = np.cumsum(np.concatenate([[0.0], np.full(S - 1, 1.0 / S)]))   # window_phase += 1/size (f64)
    idxf = phase * 2047.0
    idx = idxf.astype(np.int64) & 2047
    frac = (idxf - np.trunc(idxf)).astype(np.float32)
    env = (lut[idx] * (np.float32(1) - frac) + lut[(idx + 1) & 2047] * frac).astype(np.float32) * vol
    windowed = (np.float32(0.25) * env).astype(np.float32)
    lg = (np.float32(1) - pan) * np.float32(0.5)
    rg = (np.float32(1) + pan) * np.float32(0.5)
    exp_l = np.where(env > np.float32(0.001), windowed * lg, np.float32(0)).astype(np.float32)
    exp_r = np.where(env > np.float32(0.001), windowed * rg, np.float32(0)).astype(np.float32)
    assert np.array_equal(out[:S, 0], exp_l) and np.array_equal(out[:S, 1], exp_r)
    assert not out[S:].any()   # density 1 Hz: the next grain is 48000 frames away


def test_grain_trigger_cadence(oracle_api):
    """Cloud mode: trigger_phase starts at 1.0 and accumulates density / sr in f32 (granular.rs:788-809)."""
    from phonic_b200.player import GeneratorPlaybackOptions, GranularParameters
    sr, density = 48000, 37.0
    p = _player(oracle_api, sr)
    b = p.upload_buffer(np.full(4800, 0.25, np.float32), sr, add_pad_frame=False)
    g = p.add_generator(b, GeneratorPlaybackOptions(voices=1), None,
                        granular=GranularParameters(window=4, size=1.0, density=density, position=0.5, step=0.0))
    g.note_on(60, 1.0, 0.0, sample_time=0)
    frames = 16 * 1024
    out = p.render(frames)[:, 0]
    inc = np.float32(density) / np.float32(sr)
    ph, trig = np.float32(1.0), []
    for t in range(frames):
        ph = np.float32(ph + inc)
        if ph >= np.float32(1.0):
            ph = np.float32(ph - np.float32(1.0))
            trig.append(t)
    onsets = np.flatnonzero((out != 0) & (np.concatenate([[0], out[:-1]]) == 0))
    # 1 ms Trapezoid grain = 48 samples; sample 0 has envelope 0 (skipped), the blip starts one frame after the trigger
    assert onsets.tolist() == [t + 1 for t in trig if t + 1 < frames]


def test_distortion_shapers_known_answers(oracle_lib):
    """src/effect/distortion.rs:124-190, values derived by hand from the cited formulas (f32)."""
    lib = oracle_lib
    lib.po_test_dist_shape.restype = C.c_float
    lib.po_test_dist_compensation.restype = C.c_float
    shape = lambda t, x, d: lib.po_test_dist_shape(C.c_uint32(t), C.c_float(x), C.c_float(d))
    # drive 0: gain = 1 for the polynomial / clip / fold shapers -> identity inside [-1, 1]
    for t in (0, 1, 4):
        for x in (-0.75, -0.1, 0.0, 0.3, 0.9):
            assert shape(t, x, 0.0) == np.float32(x)
    # SoftClip, full drive: gain 15 -> saturates at +-1 beyond |x| >= 1/15, cubic below
    assert shape(0, 0.5, 4.0) == 1.0 and shape(0, -0.5, 4.0) == -1.0
    x = np.float32(0.02) * np.float32(15.0)
    assert shape(0, 0.02, 4.0) == np.float32(1.5) * (x - (x * x * x) / np.float32(3.0))
    # HardClip, full drive: gain 25, threshold 1/25 -> clamp(x) * gain
    assert shape(1, 0.5, 4.0) == np.float32(np.float32(1.0) / np.float32(25.0)) * np.float32(25.0)
    assert shape(1, 0.01, 4.0) == np.float32(0.01) * np.float32(25.0)
    # Fuzz removes the negative half wave: 1.5 * (s + |s|) == 0 for s < 0
    assert shape(3, -0.4, 2.0) == 0.0 and shape(3, 0.4, 2.0) > 0.0
    # Diode at 0 is exp(0) - 1 = 0 -> atan(0) = 0
    assert shape(2, 0.0, 3.0) == 0.0
    # Fold, full drive: gain 4, threshold 0.25; x = 0.5 * 4 = 2 -> |(|2 - .25| % 1) - .5| - .25 = 0.0
    assert shape(4, 0.5, 4.0) == 0.0
    assert shape(4, 0.05, 4.0) == np.float32(0.05) * np.float32(4.0)
    # identity shapers need no loudness compensation at drive 0; every table entry is finite and positive
    for t in (0, 1, 4):
        assert lib.po_test_dist_compensation(C.c_uint32(t), C.c_float(0.0)) == 1.0
    for t in range(5):
        for d in (0.0, 1.0, 2.5, 4.0):
            c = lib.po_test_dist_compensation(C.c_uint32(t), C.c_float(d))
            assert np.isfinite(c) and c > 0.0
        # more drive -> louder output -> smaller compensation (the wavefolder folds the peaks back instead)
        if t != 4:
            assert lib.po_test_dist_compensation(C.c_uint32(t), C.c_float(4.0)) < lib.po_test_dist_compensation(C.c_uint32(t), C.c_float(0.5))


def test_schedule_many_equals_individual_calls(oracle_api):
    """pb200_schedule_many: one call for a whole score, same audio and same note ids as the individual handle calls
    (note-addressed events refer to NOTE_ONs of the same batch by index)."""
    from phonic_b200 import workloads as W
    from phonic_b200.player import BatchNote, Player
    outs, ids = [], []
    for batched in (False, True):
        p = Player(oracle_api, 48000)
        buf = W.synth_buffer(20000, 44100, seed=1)
        bid = p.upload_buffer(buf, 44100)
        if batched:
            with p.batch():
                hs = W.add_voice_bank(p, W.VoiceBankSpec(voices=12, voices_per_sampler=4), bid, None, 0, 0.05)
                late = hs[0].note_on(50, volume=0.2, sample_time=1000)
                assert isinstance(late, BatchNote)
            hs[0].note_off(late, sample_time=9000)   # a placeholder keeps working after the flush
            ids.append(int(late))
        else:
            hs = W.add_voice_bank(p, W.VoiceBankSpec(voices=12, voices_per_sampler=4), bid, None, 0, 0.05)
            late = hs[0].note_on(50, volume=0.2, sample_time=1000)
            hs[0].note_off(late, sample_time=9000)
            ids.append(int(late))
        outs.append(p.render(24 * 1024))
        p.close()
    assert ids[0] == ids[1]
    assert float(np.abs(outs[0]).max()) > 0.01
    assert np.array_equal(outs[0], outs[1])


def test_schedule_many_rejects_forward_references(oracle_api):
    import ctypes as CT
    from phonic_b200 import _capi as A
    from phonic_b200.player import Player
    p = Player(oracle_api, 48000)
    evs = (A.Event * 1)()
    evs[0].kind = A.EV_NOTE_OFF
    evs[0].target = 1
    evs[0].note_id = 0
    evs[0].flags = A.EVF_NOTE_FROM_BATCH
    done = A.U32(7)
    assert oracle_api.schedule_many(p._r, evs, 1, CT.byref(done)) == A.ERR_PARAMETER and done.value == 0
    p.close()


def test_score_array_equals_handle_calls(oracle_api):
    """workloads.add_voice_bank_fast (one pb200_schedule_many call from a numpy score) == add_voice_bank (handle calls)."""
    from phonic_b200 import workloads as W
    from phonic_b200.player import Player
    outs = []
    for fast in (False, True):
        p = Player(oracle_api, 48000)
        bid = p.upload_buffer(W.synth_buffer(20000, 44100, seed=1), 44100)
        spec = W.VoiceBankSpec(voices=20, voices_per_sampler=8)
        (W.add_voice_bank_fast if fast else W.add_voice_bank)(p, spec, bid, None, 3, 0.05)
        outs.append(p.render(24 * 1024))
        p.close()
    assert float(np.abs(outs[0]).max()) > 0.01
    assert np.array_equal(outs[0], outs[1])
