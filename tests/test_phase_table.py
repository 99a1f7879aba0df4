"""The exact 64-frame phase jumps (phonic_b200/csrc/phase_table.cuh) on the CPU: the header is host/device code, so the
same source the skeleton kernel runs is compiled here with g++ and checked against a literal restatement of
CubicInterpolator::process's f32 recurrence (src/utils/resampler/cubic.rs:72-110).

 1. the integer model == the float loop, frame by frame (state and pushes), for every ratio class;
 2. every jump the table accepts == 64 literal frames (bit-identical sub_pos, same push count), from long orbits that
    start at 0 (what a note-on leaves), from random grid states, and from states planted right at / next to every break
    point and margin edge;
 3. the table declines (literal fallback) only on a small fraction of tiles."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "phase_table_harness.cpp")
LIB = os.path.join(ROOT, "tests", "_build", "libphase_table_harness.so")
HDR = os.path.join(ROOT, "phonic_b200", "csrc", "phase_table.cuh")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    l = C.CDLL(LIB)
    l.pt_words.restype = C.c_uint32
    l.pt_words.argtypes = [C.c_float, C.POINTER(C.c_uint32)]
    l.pt_build.argtypes = [C.c_float, C.POINTER(C.c_uint32)]
    l.pt_check_model.restype = C.c_uint64
    l.pt_check_model.argtypes = [C.c_float, C.POINTER(C.c_float), C.c_uint32, C.c_uint32]
    l.pt_check_jumps.restype = C.c_uint64
    l.pt_check_jumps.argtypes = [C.c_float, C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
    l.pt_table_stats.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    l.pt_check_accum.restype = C.c_uint64
    l.pt_check_accum.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint64)]
    return l


def file_ratio(speed, in_rate=44100, out_rate=48000):
    """FileSourceImpl::update_speed (src/source/file/common.rs:165) + CubicResampler::new: the f32 ratio of a voice"""
    new_rate = int(out_rate / speed)
    return np.float32(in_rate / new_rate)


def cfg2_ratios():
    return sorted({float(file_ratio(2.0 ** ((n - 60) / 12.0))) for n in range(29, 92)})


def geom(lib, ratio):
    g = (C.c_uint32 * 7)()
    words = lib.pt_words(C.c_float(ratio), g)
    return dict(mode=g[0], sh=g[1], R=g[2], L=g[3], n_bp=g[4], margin=g[5], nth=g[6], words=words)


def build(lib, ratio):
    g = geom(lib, ratio)
    tab = np.zeros(g["words"], np.uint32)
    lib.pt_build(C.c_float(ratio), tab.ctypes.data_as(C.POINTER(C.c_uint32)))
    return g, tab


def grid_states(g, rng, n, ratio):
    """random sub_pos values on the ratio's grid, inside the reachable range"""
    one = 1 << g["sh"]
    top = one + g["R"] if ratio < 1.0 else one
    S = rng.integers(0, top, n, dtype=np.int64)
    # reachable states carry at most 24 significant bits
    msb = np.floor(np.log2(np.maximum(S, 1))).astype(np.int64)
    drop = np.maximum(msb - 23, 0)
    S = (S >> drop) << drop
    return (S.astype(np.float64) / one).astype(np.float32)


RATIOS = sorted(set(cfg2_ratios() + [0.91875, 0.5, 0.75, 0.7500001, 0.24301, 0.1300001, 0.0700001, 0.999, 1.0000021, 1.001, 1.25, 1.37, 1.5,
                                     1.9999, 2.0001, 2.5, 2.99, 3.0000002, 3.3, 3.999, 5.3, 7.5, 7.99, 11.7, 13.99]))


def test_integer_model_equals_the_f32_recurrence(lib):
    rng = np.random.default_rng(1)
    for ratio in RATIOS:
        g = geom(lib, ratio)
        if g["mode"] == 0:
            continue
        starts = np.concatenate([np.zeros(1, np.float32), grid_states(g, rng, 4000, ratio)])
        bad = lib.pt_check_model(C.c_float(ratio), starts.ctypes.data_as(C.POINTER(C.c_float)), len(starts), 300)
        assert bad == 0, (ratio, g)


def test_modes_cover_the_expected_classes(lib):
    assert geom(lib, 1.0)["mode"] == 0 and geom(lib, 3.0)["mode"] == 0 and geom(lib, 20.0)["mode"] == 0 and geom(lib, 0.03)["mode"] == 0
    assert geom(lib, 0.5)["mode"] == 1 and geom(lib, 0.75)["mode"] == 1      # no sum ever rounds
    assert geom(lib, float(np.float32(0.91875)))["mode"] in (1, 2)
    assert geom(lib, 2.5)["mode"] == 3 and geom(lib, 5.3)["mode"] == 3       # ratio + 1 stays inside the ratio's binade
    assert geom(lib, 1.37)["mode"] == 4 and geom(lib, 3.3)["mode"] == 4 and geom(lib, 7.5)["mode"] == 4


def test_jumps_equal_64_literal_frames(lib):
    rng = np.random.default_rng(2)
    declined_frac = []
    for ratio in RATIOS:
        g, tab = build(lib, ratio)
        if g["mode"] == 0:
            continue
        tp = tab.ctypes.data_as(C.POINTER(C.c_uint32))
        out = (C.c_uint64 * 4)()
        # (a) the orbit a note-on starts: sub_pos = 0, 20000 tiles = 27 s at 48 kHz
        z = np.zeros(1, np.float32)
        bad = lib.pt_check_jumps(C.c_float(ratio), tp, z.ctypes.data_as(C.POINTER(C.c_float)), 1, 20000, out)
        assert bad == 0, (ratio, g, list(out))
        applied, declined = out[0], out[1]
        declined_frac.append((declined / 20000.0, ratio))
        # (b) random grid states, 8 tiles each
        starts = grid_states(g, rng, 20000, ratio)
        bad = lib.pt_check_jumps(C.c_float(ratio), tp, starts.ctypes.data_as(C.POINTER(C.c_float)), len(starts), 8, out)
        assert bad == 0, (ratio, g, list(out))
        # (c) states at and around every break point and margin edge
        if g["mode"] in (2, 4):
            n_bp = g["n_bp"]
            bp = tab[16 + 258:16 + 258 + n_bp].astype(np.int64)
            one = 1 << g["sh"]
            offs = np.array([-g["margin"] - 2, -g["margin"] - 1, -g["margin"], -g["margin"] + 1, -3, -2, -1, 0, 1, 2, 3,
                             g["margin"] - 1, g["margin"], g["margin"] + 1, g["margin"] + 2], np.int64)
            S = (bp[:, None] + offs[None, :]).reshape(-1)
            for res in range(0, g["L"]):
                S2 = np.concatenate([S, (S & ~(g["L"] - 1)) | res])
                S2 = S2[(S2 >= 0) & (S2 < one)]
                msb = np.floor(np.log2(np.maximum(S2, 1))).astype(np.int64)
                drop = np.maximum(msb - 23, 0)
                S2 = (S2 >> drop) << drop
                st = (S2.astype(np.float64) / one).astype(np.float32)
                bad = lib.pt_check_jumps(C.c_float(ratio), tp, st.ctypes.data_as(C.POINTER(C.c_float)), len(st), 2, out)
                assert bad == 0, (ratio, g, res, list(out))
    worst = max(declined_frac)
    assert worst[0] < 0.05, f"table declines {worst[0]:.1%} of the tiles at ratio {worst[1]}"
    print("declined fraction per ratio (orbit from 0):", [(round(r, 5), round(f, 4)) for f, r in declined_frac])


def test_off_grid_states_are_declined_then_join_the_grid(lib):
    """After a glide sub_pos is whatever the previous ratios left: one literal tile puts it on the new ratio's grid."""
    for ratio in (0.7071068, 0.24301, 1.37, 3.3):
        ratio = float(np.float32(ratio))
        g, tab = build(lib, ratio)
        tp = tab.ctypes.data_as(C.POINTER(C.c_uint32))
        st = np.array([0.123456789, 0.9999999, 1e-7, 0.33333334], np.float32)
        out = (C.c_uint64 * 4)()
        bad = lib.pt_check_jumps(C.c_float(ratio), tp, st.ctypes.data_as(C.POINTER(C.c_float)), len(st), 50, out)
        assert bad == 0
        assert out[0] >= 4 * 45  # at most a few literal tiles per start


def test_table_coverage(lib):
    out = (C.c_uint64 * 4)()
    for ratio in cfg2_ratios():
        g, tab = build(lib, ratio)
        if g["mode"] not in (2, 4):
            continue
        lib.pt_table_stats(tab.ctypes.data_as(C.POINTER(C.c_uint32)), out)
        one = 1 << g["sh"]
        assert out[1] / one > 0.98, (ratio, out[1] / one)          # states inside a valid interval
        assert out[3] / max(out[2], 1) < 0.02, (ratio, out[3], out[2])  # entries that failed verification


def test_accumulate_jump_equals_the_literal_chain(lib):
    """AHDSR stage chains (`env += rate` per frame, ahdsr.rs:448-516): n steps at once == n literal f32 adds."""
    rng = np.random.default_rng(3)
    n = 400000
    o = np.exp(rng.uniform(np.log(1e-4), np.log(4.0), n)).astype(np.float32)
    # rates of real envelopes (1 / (t * sr)), ties, huge and tiny steps, both signs
    d = (np.exp(rng.uniform(np.log(1e-9), np.log(0.3), n)) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    d[: n // 10] = (np.ldexp(1.0, rng.integers(-30, -3, n // 10)) * rng.choice([-1.5, -1.0, -0.5, 0.5, 1.0, 1.5], n // 10)).astype(np.float32)
    # states next to binade edges
    o[n // 2: n // 2 + n // 10] = np.nextafter(np.ldexp(1.0, rng.integers(-8, 2, n // 10)).astype(np.float32),
                                               np.float32(0) if True else np.float32(9), dtype=np.float32)
    o[n // 2 + n // 10: n // 2 + n // 5] = np.ldexp(1.0, rng.integers(-8, 2, n // 10)).astype(np.float32)
    steps = rng.integers(1, 65, n).astype(np.uint32)
    steps[::3] = 64
    out = (C.c_uint64 * 4)()
    bad = lib.pt_check_accum(o.ctypes.data_as(C.POINTER(C.c_float)), d.ctypes.data_as(C.POINTER(C.c_float)),
                             steps.ctypes.data_as(C.POINTER(C.c_uint32)), n, out)
    assert bad == 0, (out[2], o[out[2]], d[out[2]], steps[out[2]])
    assert out[0] > 0.5 * n
    # the hold countdown: 48000.0, 47999.0, ...
    o2 = np.arange(100.0, 60000.0, 7.0, dtype=np.float32)
    d2 = np.full_like(o2, -1.0)
    s2 = np.full(len(o2), 64, np.uint32)
    out = (C.c_uint64 * 4)()
    assert lib.pt_check_accum(o2.ctypes.data_as(C.POINTER(C.c_float)), d2.ctypes.data_as(C.POINTER(C.c_float)),
                              s2.ctypes.data_as(C.POINTER(C.c_uint32)), len(o2), out) == 0
    assert out[0] > 0.9 * len(o2)
