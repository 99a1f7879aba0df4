"""The skeleton pass replaces the reference's `while sub_pos < ratio { sub_pos += 1.0 }` loop
(src/utils/resampler/cubic.rs:94-103) by a closed form of the repeated f32 `+= 1.0`
(phonic_b200/csrc/voice.cuh, voice_advance). This pins the closed form bit for bit against the sequential
adds in numpy float32, including binade edges. CPU only."""
import numpy as np

f = np.float32


def seq(s, n):
    t = f(s)
    for _ in range(n):
        t = f(t + f(1.0))
    return t


def closed(s, n):
    a1, a2, a3, a4 = min(n, 1), min(max(n - 1, 0), 2), min(max(n - 3, 0), 4), min(max(n - 7, 0), 8)
    t = f(s)
    for a in (a1, a2, a3, a4):
        t = f(t + f(a))
    return t


def test_closed_form_equals_sequential_adds():
    rng = np.random.default_rng(0)
    cases = []
    for i in range(60000):
        kind = i % 4
        if kind == 0:
            s = f(rng.random())
        elif kind == 1:
            s = f(1.0) - f(rng.random() * 1e-6)
        elif kind == 2:
            s = f(rng.random() * 2.0 ** -int(rng.integers(1, 20)))
        else:
            s = np.nextafter(f(rng.integers(0, 2 ** 24) / 2 ** 24), f(1), dtype=f)
        if s >= 1:
            s = np.nextafter(f(1), f(0), dtype=f)
        cases.append(s)
    cases += [f(0.0), np.nextafter(f(1), f(0), dtype=f), f(0.5), f(0.25), f(2.0 ** -24), f(1 - 2.0 ** -24)]
    for s in cases:
        n = int(rng.integers(0, 16))
        assert seq(s, n) == closed(s, n), (repr(s), n)
    for s in cases[-6:]:
        for n in range(16):
            assert seq(s, n) == closed(s, n), (repr(s), n)


def test_push_count_is_floor_or_floor_plus_one():
    rng = np.random.default_rng(1)
    for _ in range(5000):
        ratio = f(rng.uniform(1.0, 13.9))
        s = f(rng.random())
        t, k = s, 0
        while t < ratio:
            t = f(t + f(1.0))
            k += 1
        assert k in (int(ratio), int(ratio) + 1)
        assert f(t - ratio) < 1.0 and f(t - ratio) >= 0.0
