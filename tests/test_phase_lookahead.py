"""The skeleton pass runs the ratio < 1 phase recurrence of CubicInterpolator::process (src/utils/resampler/cubic.rs:
73-89) with the push flag computed one frame ahead from thresholds, and with comparisons evaluated as one saturating
FMA (phonic_b200/csrc/voice.cuh, phase_run). This restates both tricks in numpy float32 / exact arithmetic and pins
them bit for bit against the literal recurrence. CPU only."""
from fractions import Fraction

import numpy as np

f = np.float32
K = 2.0 ** 60


def pred(x):
    return np.nextafter(f(x), f(-np.inf), dtype=f)


def succ(x):
    return np.nextafter(f(x), f(np.inf), dtype=f)


def fma_sat(a, b, c):
    """fma.rn.sat.f32: round(a * b + c) exactly once, then clamp to [0, 1]."""
    exact = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
    r = f(float(exact)) if abs(exact) < Fraction(2) ** 120 else f(np.inf if exact > 0 else -np.inf)
    # float(Fraction) rounds to f64 first; the results used here are either >= 1, <= 0 or exactly representable
    return f(min(max(float(r), 0.0), 1.0))


def step_weight(x):
    return f(-(float(pred(x)) * K))


def step(s, w):
    return fma_sat(s, f(K), w)


def literal(s, ratio, n):
    pushes = 0
    out = []
    for _ in range(n):
        if s >= f(1.0):
            s = f(s - f(1.0))
            pushes += 1
        s = f(s + ratio)
        out.append(s)
    return out, pushes


def lookahead(s, ratio, n):
    w1 = step_weight(f(1.0))
    thrA = f(f(1.0) - ratio)
    while f(pred(thrA) + ratio) >= f(1.0):
        thrA = pred(thrA)
    while f(thrA + ratio) < f(1.0):
        thrA = succ(thrA)
    wA = step_weight(thrA)
    out, pushes = [], f(0)
    p = step(s, w1)
    s = f(f(s - p) + ratio)
    pushes = f(pushes + p)
    out.append(s)
    p = step(s, w1)
    if ratio < f(0.499):
        for _ in range(1, n):
            pn = f(step(s, wA) - step(s, w1))
            s = f(f(s - p) + ratio)
            pushes = f(pushes + p)
            p = pn
            out.append(s)
    else:
        thrB = f(f(2.0) - ratio)
        while f(f(pred(thrB) - f(1.0)) + ratio) >= f(1.0):
            thrB = pred(thrB)
        while f(f(thrB - f(1.0)) + ratio) < f(1.0):
            thrB = succ(thrB)
        wB = step_weight(thrB)
        for _ in range(1, n):
            pn = f(f(step(s, wA) - step(s, w1)) + step(s, wB))
            s = f(f(s - p) + ratio)
            pushes = f(pushes + p)
            p = pn
            out.append(s)
    return out, int(pushes)


def test_saturating_fma_is_an_exact_comparison():
    rng = np.random.default_rng(3)
    xs = [f(1.0), f(0.5), f(2.0 ** -20), f(1.5), f(0.08125), f(1.0 - 2.0 ** -24)] + [f(rng.uniform(1e-6, 2.0)) for _ in range(200)]
    for x in xs:
        w = step_weight(x)
        for s in (pred(pred(x)), pred(x), x, succ(x), succ(succ(x)), f(0.0), f(1.9999999), f(rng.uniform(0, 2))):
            assert step(s, w) == f(1.0 if s >= x else 0.0), (repr(x), repr(s))


def test_lookahead_equals_literal_recurrence():
    rng = np.random.default_rng(4)
    ratios = [f(0.91875), f(0.22968750), f(0.4989), f(0.499), f(0.4991), f(0.5), f(0.50000006), f(0.75), f(0.999998), f(0.001),
              f(0.3333333), f(0.6666667), pred(f(0.5)), pred(f(1.0 - 1e-6))]
    ratios += [f(rng.uniform(0.01, 0.999)) for _ in range(120)]
    for ratio in ratios:
        starts = [f(0.0), f(rng.random()), f(1.0), pred(f(1.0)), f(1.0 + float(ratio)) if ratio < 0.98 else f(1.5), f(1.99), f(rng.uniform(1.0, 1.99))]
        for s0 in starts:
            a, pa = literal(f(s0), ratio, 160)
            b, pb = lookahead(f(s0), ratio, 160)
            assert a == b and pa == pb, (repr(ratio), repr(s0))


def test_push_count_is_the_rounded_phase_balance():
    """Pushes per piece are not counted per frame on the device: s_out = s_in - pushes + n * ratio (ratio < 1) or
    s_in + pushes - n * ratio (ratio >= 1) up to the recurrence's own rounding, so pushes = rint() of the balance."""
    rng = np.random.default_rng(5)
    for _ in range(300):
        ratio = f(rng.uniform(0.01, 0.999))
        s0 = f(rng.uniform(0.0, 1.99))
        n = int(rng.integers(1, 1025))
        out, pushes = literal(s0, ratio, n)
        assert int(np.rint((float(s0) - float(out[-1])) + n * float(ratio))) == pushes
    for _ in range(300):
        ratio = f(rng.uniform(1.0, 13.99))
        s = s0 = f(rng.random())
        n = int(rng.integers(1, 1025))
        pushes = 0
        for _ in range(n):
            while s < ratio:
                s = f(s + f(1.0))
                pushes += 1
            s = f(s - ratio)
        assert int(np.rint((float(s) - float(s0)) + n * float(ratio))) == pushes
