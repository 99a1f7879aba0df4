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
