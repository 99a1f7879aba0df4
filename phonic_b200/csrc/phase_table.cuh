// Exact 64-frame jumps of CubicInterpolator's f32 phase recurrence (src/utils/resampler/cubic.rs:72-110).
//
// The reference advances `sub_pos` by one rounded f32 operation chain per output frame; whether an input sample is
// pushed depends on the rounded value, so a voice's position at frame n is the result of n dependent f32 adds
// (SURVEY.md H1). This file turns that chain into integer arithmetic and jumps over a whole replay tile at once:
//
//  * Every reachable sub_pos is a multiple of u = ulp(ratio) (a sum never lands in a finer binade than the ratio's),
//    so the state is an integer S = sub_pos / u, the ratio an integer R in [2^23, 2^24), 1.0 the power of two ONE.
//    One frame is then `S += R` (ratio < 1; minus ONE when S >= ONE) or `S += n ONE - R` (ratio >= 1) with the sum
//    rounded to 24 significant bits, ties to even: phase_step_down / phase_step_up below restate exactly what the
//    f32 unit does (tests/test_phase_table.py checks the model against the literal float loop bit for bit).
//  * Without rounding the recurrence is a rotation S -> S + K R (mod ONE): closed form. Rounding only happens when a
//    sum enters a binade above the ratio's, and then perturbs S by a few units that depend only on (a) which binade
//    each of the K sums falls in and whether it wraps -- a pattern that is constant between K x (#binades) break
//    points of S -- and (b) the low bits of S. For one ratio the K-frame map is therefore a table
//    (interval of S) x (S mod L) -> (pushes W, perturbation P), S_K = S + K R - W ONE + P.
//  * The table is not trusted to the analysis: the build kernel places the break points analytically, keeps a margin
//    around each, and then VERIFIES every (interval, residue) entry by running the integer model for the interval's
//    lowest and highest state of that residue: the decision sequences must be identical (fl() is monotone, so every
//    state in between takes the same decisions) and so must W and P. Entries that fail are marked invalid and the
//    skeleton takes the literal per-frame loop for such a tile (as it does for states inside a margin, states that
//    are not on the grid yet -- the first tile after a glide -- integer ratios, ratio >= 14 and ratio < 1/16).
//  * Ratios whose sums can never round (R a multiple of ONE's coarsest granule; ratio >= 1 with ratio + 1 inside the
//    ratio's own binade) need no table: the jump is three integer operations.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

namespace pb {

constexpr uint32_t PT_K = 64;          // frames per jump (== TILE)
constexpr uint32_t PT_MAX_BP = 256;    // K x binades above the ratio's (ratio >= 1/16)
constexpr uint32_t PT_COARSE = 256;    // buckets of the first-level index over S
constexpr uint32_t PT_HEADER = 16;

enum PhaseMode : uint32_t { PT_LITERAL = 0, PT_DOWN_EXACT = 1, PT_DOWN_TABLE = 2, PT_UP_EXACT = 3, PT_UP_TABLE = 4 };

// header words of one ratio's table
enum { PT_H_RATIO = 0, PT_H_MODE = 1, PT_H_SH = 2, PT_H_R = 3, PT_H_LMASK = 4, PT_H_NBP = 5, PT_H_CSHIFT = 6, PT_H_MARGIN = 7, PT_H_WORDS = 8 };

struct PhaseGeom {
  uint32_t mode;
  uint32_t sh;      // ONE = 1 << sh
  uint32_t R;       // ratio / u
  uint32_t L;       // residue modulus (power of two)
  uint32_t n_bp;    // analytic break points
  uint32_t margin;  // states closer than this to a break point take the literal loop
  uint32_t nth;     // binade thresholds per frame (ratio < 1)
};

PB_HD uint32_t pt_f32_bits(float f) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__float_as_int(f);
#else
  union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
PB_HD float pt_bits_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __int_as_float((int)u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
PB_HD int pt_msb(uint32_t x) {  // index of the highest set bit, x > 0
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)x);
#else
  return 31 - __builtin_clz(x);
#endif
}

// round `t` to 24 significant bits, ties to even (what fl() does to an exact integer sum)
PB_HD uint32_t pt_rne24(uint32_t t) {
  const int b = pt_msb(t);
  if (b <= 23) return t;
  const int sh = b - 23;
  const uint32_t g = 1u << sh, rem = t & (g - 1u), half = g >> 1;
  uint32_t t0 = t - rem;
  if (rem > half || (rem == half && ((t0 >> sh) & 1u))) t0 += g;
  return t0;
}

// Geometry of a ratio: which jump applies and the table's dimensions. Mirrors simple_call_ok's domain.
PB_HD PhaseGeom phase_geom(float ratio) {
  PhaseGeom g;
  g.mode = PT_LITERAL; g.sh = 0; g.R = 0; g.L = 1; g.n_bp = 0; g.margin = 0; g.nth = 0;
  const uint32_t bits = pt_f32_bits(ratio);
  if (!(ratio > 0.0f) || !(ratio < 14.0f)) return g;
  const float d1 = ratio - 1.0f;
  if ((d1 < 0.0f ? -d1 : d1) < 0.000001f) return g;           // bypass copy (cubic.rs:53-58)
  const int e = (int)((bits >> 23) & 0xFFu) - 127;             // ratio in [2^e, 2^(e+1))
  if (e < -4) return g;                                        // ratio < 1/16: literal
  g.sh = (uint32_t)(23 - e);
  g.R = (bits & 0x7FFFFFu) | 0x800000u;
  if (ratio < 1.0f) {
    // sums reach binades e+1 .. 0 (values up to 1 + ratio): granules 2 .. 2^-e units
    const uint32_t coarsest = 1u << (uint32_t)(-e);
    if ((g.R & (coarsest - 1u)) == 0u) { g.mode = PT_DOWN_EXACT; g.L = coarsest; return g; }  // exact while S is a multiple of it too
    g.mode = PT_DOWN_TABLE;
    g.L = coarsest << 1;                                       // the tie rule looks one bit above the granule
    g.nth = (uint32_t)(-e);                                    // thresholds 2^24, 2^25, ..., ONE
    g.n_bp = PT_K * g.nth;
    g.margin = PT_K * (coarsest >> 1) + 2u;                    // largest possible drift of the true orbit from the exact one
  } else {
    if (ratio == (float)(int)ratio) return g;                  // integer ratios: literal (the loop's exit test can tie)
    // t = sub_pos + n is < ratio + 1: it only rounds when that crosses into binade e+1
    const uint32_t ONE = 1u << g.sh;
    if (g.R + ONE <= 0x1000000u) { g.mode = PT_UP_EXACT; return g; }
    g.mode = PT_UP_TABLE;
    g.L = 4;
    g.n_bp = PT_K;
    g.margin = PT_K + 2u;
  }
  return g;
}

// words of one ratio's table: header | coarse[PT_COARSE + 1] | bpx[n_bp + 2] | lo[n_bp + 1] | hi[n_bp + 1] | entry[(n_bp + 1) * L]
// bpx[0] = 0 and bpx[n_bp + 1] = ONE are sentinels around the sorted break points bpx[1 .. n_bp]
PB_HD uint32_t phase_table_words(const PhaseGeom& g) {
  if (g.mode != PT_DOWN_TABLE && g.mode != PT_UP_TABLE) return PT_HEADER;
  const uint32_t w = PT_HEADER + (PT_COARSE + 1u) + (g.n_bp + 2u) + 2u * (g.n_bp + 1u) + (g.n_bp + 1u) * g.L;
  return (w + 3u) & ~3u;  // tables start 16-byte aligned (bulk copies into shared memory)
}

// ---- the integer model of one frame ---------------------------------------------------------------------------------
// ratio < 1 (cubic.rs:73-89): `if sub_pos >= 1 { push; sub_pos -= 1 }  emit  sub_pos += ratio`
// `sig` folds in every decision the frame took (wrap, binade of the sum) for the build's verification.
PB_HD void phase_step_down(uint32_t& S, uint32_t& W, uint64_t& sig, const uint32_t ONE, const uint32_t R) {
  uint32_t w = 0;
  if (S >= ONE) { S -= ONE; w = 1; }
  W += w;
  const uint32_t t = S + R;
  sig = sig * 6364136223846793005ull + (uint64_t)((uint32_t)pt_msb(t) * 2u + w) + 1442695040888963407ull;
  S = pt_rne24(t);
}
// ratio >= 1 (cubic.rs:94-105): `while sub_pos < ratio { push; sub_pos += 1 }  sub_pos -= ratio  emit`
PB_HD void phase_step_up(uint32_t& S, uint32_t& W, uint64_t& sig, const uint32_t ONE, const uint32_t R) {
  uint32_t t = S, n = 0;
  while (t < R) { t = pt_rne24(t + ONE); ++n; }   // only a sum >= 2^24 rounds, and that one ends the loop
  W += n;
  sig = sig * 6364136223846793005ull + (uint64_t)((uint32_t)pt_msb(t) * 64u + n) + 1442695040888963407ull;
  S = t - R;                                      // exact (Sterbenz)
}

struct PhaseSim { uint32_t S, W; uint64_t sig; };
PB_HD PhaseSim phase_simulate(uint32_t S, const PhaseGeom& g) {
  PhaseSim r;
  r.S = S; r.W = 0; r.sig = 0;
  const uint32_t ONE = 1u << g.sh;
  if (g.mode == PT_DOWN_EXACT || g.mode == PT_DOWN_TABLE)
    for (uint32_t k = 0; k < PT_K; ++k) phase_step_down(r.S, r.W, r.sig, ONE, g.R);
  else
    for (uint32_t k = 0; k < PT_K; ++k) phase_step_up(r.S, r.W, r.sig, ONE, g.R);
  return r;
}

// ---- table build -----------------------------------------------------------------------------------------------------
// Break point `j` of the exact (unrounded) orbit, in [0, ONE): the state at which frame k's sum reaches a threshold.
PB_HD uint32_t phase_break_point(const PhaseGeom& g, uint32_t j) {
  const uint32_t ONE = 1u << g.sh, mask = ONE - 1u;
  if (g.mode == PT_DOWN_TABLE) {
    const uint32_t k = j / g.nth, i = j % g.nth;
    const uint32_t theta = 0x1000000u << i;                    // 2^24 .. ONE
    return (theta - (k + 1u) * g.R) & mask;                    // S + (k + 1) R == theta (mod ONE)
  }
  const uint32_t rho = g.R & mask;                             // frame k pushes one more sample while sub_pos < frac(ratio)
  return (rho + j * g.R) & mask;                               // S - k R == rho (mod ONE)
}

// One (interval, residue) entry: verified against the model at both ends of the interval. 0 = invalid.
// entry = 1 << 31 | W << 16 | (P + 32768)
PB_HD uint32_t phase_build_entry(const PhaseGeom& g, uint32_t lo, uint32_t hi, uint32_t residue) {
  if (lo > hi) return 0u;
  const uint32_t Lm = g.L - 1u;
  uint32_t a = (lo & ~Lm) | residue;
  if (a < lo) a += g.L;
  uint32_t b = (hi & ~Lm) | residue;
  if (b > hi) { if (b < g.L) return 0u; b -= g.L; }
  if (a > b) return 0u;
  const PhaseSim ra = phase_simulate(a, g), rb = phase_simulate(b, g);
  if (ra.sig != rb.sig || ra.W != rb.W) return 0u;
  const int64_t ONE = (int64_t)1 << g.sh, KR = (int64_t)PT_K * g.R;
  const bool down = g.mode == PT_DOWN_TABLE;
  const int64_t pa = down ? (int64_t)ra.S - ((int64_t)a + KR - (int64_t)ra.W * ONE) : (int64_t)ra.S - ((int64_t)a + (int64_t)ra.W * ONE - KR);
  const int64_t pb_ = down ? (int64_t)rb.S - ((int64_t)b + KR - (int64_t)rb.W * ONE) : (int64_t)rb.S - ((int64_t)b + (int64_t)rb.W * ONE - KR);
  if (pa != pb_ || pa < -32768 || pa > 32767 || ra.W > 0xFFFu) return 0u;
  return 0x80000000u | (ra.W << 16) | (uint32_t)(pa + 32768);
}

// Cooperative build of one ratio's table by `nthreads` workers; worker `tid` calls each phase in turn with a barrier
// between phases (a CTA on the device, a plain loop over tid on the host). `scratch` holds PT_MAX_BP words.
PB_HD void phase_build_header(uint32_t* tab, float ratio, const PhaseGeom& g) {
  for (uint32_t i = 0; i < PT_HEADER; ++i) tab[i] = 0;
  tab[PT_H_RATIO] = pt_f32_bits(ratio); tab[PT_H_MODE] = g.mode; tab[PT_H_SH] = g.sh; tab[PT_H_R] = g.R;
  tab[PT_H_LMASK] = g.L - 1u; tab[PT_H_NBP] = g.n_bp; tab[PT_H_MARGIN] = g.margin;
  tab[PT_H_CSHIFT] = g.sh > 8u ? g.sh - 8u : 0u;             // PT_COARSE = 256 buckets over [0, ONE)
  tab[PT_H_WORDS] = phase_table_words(g);
}
PB_HD uint32_t* phase_tab_coarse(uint32_t* tab) { return tab + PT_HEADER; }
PB_HD uint32_t* phase_tab_bp(uint32_t* tab) { return tab + PT_HEADER + PT_COARSE + 2u; }  // bpx + 1: the real break points
PB_HD uint32_t* phase_tab_lo(uint32_t* tab, uint32_t n_bp) { return phase_tab_bp(tab) + n_bp + 1u; }
PB_HD uint32_t* phase_tab_hi(uint32_t* tab, uint32_t n_bp) { return phase_tab_lo(tab, n_bp) + n_bp + 1u; }
PB_HD uint32_t* phase_tab_entry(uint32_t* tab, uint32_t n_bp) { return phase_tab_hi(tab, n_bp) + n_bp + 1u; }

// phase 1: raw break points -> scratch
PB_HD void phase_build_p1(const PhaseGeom& g, uint32_t* scratch, uint32_t tid, uint32_t nthreads) {
  for (uint32_t j = tid; j < g.n_bp; j += nthreads) scratch[j] = phase_break_point(g, j);
}
// phase 2: rank sort scratch -> bp[]
PB_HD void phase_build_p2(uint32_t* tab, const PhaseGeom& g, const uint32_t* scratch, uint32_t tid, uint32_t nthreads) {
  uint32_t* bp = phase_tab_bp(tab);
  for (uint32_t j = tid; j < g.n_bp; j += nthreads) {
    const uint32_t v = scratch[j];
    uint32_t rank = 0;
    for (uint32_t i = 0; i < g.n_bp; ++i) rank += (scratch[i] < v || (scratch[i] == v && i < j)) ? 1u : 0u;
    bp[rank] = v;
  }
}
// phase 3: coarse index, interval bounds with the margin applied
PB_HD void phase_build_p3(uint32_t* tab, const PhaseGeom& g, uint32_t tid, uint32_t nthreads) {
  const uint32_t* bp = phase_tab_bp(tab);
  uint32_t* coarse = phase_tab_coarse(tab);
  uint32_t* lo = phase_tab_lo(tab, g.n_bp);
  uint32_t* hi = phase_tab_hi(tab, g.n_bp);
  const uint32_t cshift = tab[PT_H_CSHIFT], ONE = 1u << g.sh;
  if (tid == 0) { phase_tab_bp(tab)[-1] = 0u; phase_tab_bp(tab)[g.n_bp] = ONE; }
  for (uint32_t c = tid; c <= PT_COARSE; c += nthreads) {      // coarse[c] = number of break points below the bucket's first state
    const uint32_t first = c << cshift;
    uint32_t n = 0;
    for (uint32_t i = 0; i < g.n_bp; ++i) n += bp[i] < first ? 1u : 0u;
    coarse[c] = n;
  }
  for (uint32_t i = tid; i <= g.n_bp; i += nthreads) {         // interval i = [bp[i-1], bp[i]) shrunk by the margin
    const uint64_t l = i == 0 ? 0u : (uint64_t)bp[i - 1] + g.margin;
    const int64_t h = i == g.n_bp ? (int64_t)ONE - 1 : (int64_t)bp[i] - 1 - (int64_t)g.margin;
    if (h < 0 || l > (uint64_t)h) { lo[i] = 1u; hi[i] = 0u; }
    else { lo[i] = (uint32_t)l; hi[i] = (uint32_t)h; }
  }
}
// phase 4: the verified entries
PB_HD void phase_build_p4(uint32_t* tab, const PhaseGeom& g, uint32_t tid, uint32_t nthreads) {
  const uint32_t* lo = phase_tab_lo(tab, g.n_bp);
  const uint32_t* hi = phase_tab_hi(tab, g.n_bp);
  uint32_t* entry = phase_tab_entry(tab, g.n_bp);
  const uint32_t total = (g.n_bp + 1u) * g.L;
  for (uint32_t x = tid; x < total; x += nthreads) {
    const uint32_t i = x / g.L, res = x % g.L;
    entry[x] = phase_build_entry(g, lo[i], hi[i], res);
  }
}

// ---- the jump --------------------------------------------------------------------------------------------------------
// A table's header in registers + where its body lives (global memory, or the copy a warp staged in shared memory).
struct PhaseRef {
  const uint32_t* body;   // coarse[] | bp[] | lo[] | hi[] | entry[]
  const uint32_t* entry;
  uint32_t mode, sh, R, KR, lmask, lshift, n_bp, cshift, margin;
  float scale, inv_scale; // 2^sh (S = sub_pos * scale) and 2^-sh
  uint32_t smem_body;     // device: shared-memory address of a staged body, 0 = read through `body`
};
PB_HD PhaseRef phase_ref(const uint32_t* tab, const uint32_t* body) {
  PhaseRef r;
  r.body = nullptr; r.entry = nullptr; r.mode = PT_LITERAL; r.sh = r.R = r.KR = r.lmask = r.lshift = r.n_bp = r.cshift = r.margin = 0;
  r.scale = r.inv_scale = 1.0f; r.smem_body = 0;
  if (tab == nullptr) return r;
  r.mode = tab[PT_H_MODE]; r.sh = tab[PT_H_SH]; r.R = tab[PT_H_R]; r.lmask = tab[PT_H_LMASK]; r.n_bp = tab[PT_H_NBP];
  r.cshift = tab[PT_H_CSHIFT]; r.margin = tab[PT_H_MARGIN];
  r.KR = PT_K * r.R;
  r.scale = pt_bits_f32((127u + r.sh) << 23); r.inv_scale = pt_bits_f32((127u - r.sh) << 23);
  r.body = body ? body : tab + PT_HEADER;
  r.entry = r.body + PT_COARSE + 1u + (r.n_bp + 2u) + 2u * (r.n_bp + 1u);
  r.lshift = 0; while ((1u << r.lshift) <= r.lmask) ++r.lshift;
  return r;
}

// Advance `s` (sub_pos at a frame boundary) by PT_K frames; `np` += samples pushed. Returns false (nothing changed)
// when the jump does not apply and the caller has to run the literal loop for this tile. All integer arithmetic is
// modulo 2^32: every intermediate true value is below 2^31.
PB_HD bool phase_jump(const PhaseRef& t, float& s, uint32_t& np) {
  const uint32_t mode = t.mode;
  if (mode == PT_LITERAL) return false;
  const uint32_t sh = t.sh;
  const uint32_t ONE = 1u << sh;
  const float x = s * t.scale;                                  // S = s / u exactly (power-of-two scaling)
  if (!(x >= 0.0f) || !(x < 1073741824.0f)) return false;
  uint32_t S = (uint32_t)x;
  if ((float)S != x) return false;                              // not on the ratio's grid (yet)
  uint32_t W0 = 0;
  const bool down = mode <= PT_DOWN_TABLE;
  if (down && S >= ONE) { S -= ONE; W0 = 1; }                   // the first frame's wrap (`sub_pos -= 1.0` is exact)
  if (S >= ONE) return false;
  uint32_t SK, W;
  if (mode == PT_DOWN_EXACT) {
    if (S & t.lmask) return false;                               // a state left behind by another ratio: its sums still round
    W = (S + t.KR - t.R) >> sh;                                  // wraps taken by frames 0 .. K-1
    SK = S + t.KR - (W << sh);
  } else if (mode == PT_UP_EXACT || (mode == PT_UP_TABLE && ((S | t.R) & 1u) == 0u)) {
    // no sum can round: either t < 2^24 always, or every t is even (a rational ratio with a small denominator, whose
    // short orbit would otherwise sit on the table's break points forever)
    SK = (S - t.KR) & (ONE - 1u);
    W = (t.KR + SK - S) >> sh;
  } else {
    const uint32_t* coarse = t.body;
    const uint32_t* bpx = coarse + PT_COARSE + 1u;
    uint32_t i = coarse[S >> t.cshift];
    uint32_t next = bpx[i + 1u];
    while (next <= S) { ++i; next = bpx[i + 1u]; }
    const uint32_t prev = bpx[i];
    // interval i = [prev, next) minus the margin at both ends (never looser than phase_build_p3's bounds)
    if (S - prev < t.margin || next - S <= t.margin) return false;
    const uint32_t e = t.entry[(i << t.lshift) + (S & t.lmask)];
    if (!(e & 0x80000000u)) return false;
    W = (e >> 16) & 0xFFFu;
    const uint32_t P = (e & 0xFFFFu) - 32768u;
    SK = down ? S + t.KR - (W << sh) + P : S + (W << sh) - t.KR + P;
  }
  s = (float)SK * t.inv_scale;                                  // SK has at most 24 significant bits: exact
  np += W0 + W;
  return true;
}
PB_HD bool phase_jump(const uint32_t* tab, float& s, uint32_t& np) { return phase_jump(phase_ref(tab, nullptr), s, np); }

// ---- bare f32 accumulate chains (AHDSR stages, ahdsr.rs:448-516) ----------------------------------------------------
// `n` steps of o = fl(o + d) at once. While every sum stays inside o's binade, one step adds the same integer number
// of ulps: d = (c + f) ulp with |f| < 1/2 rounds to c ulps whatever o's mantissa is (a tie, |f| == 1/2, depends on
// its parity and is declined), so n steps move the mantissa by n c. Returns false (o untouched) when a sum could
// leave the binade -- the caller then runs the literal chain for this tile.
PB_HD bool accum_jump(float& o, const float d, const uint32_t n) {
  const uint32_t bits = pt_f32_bits(o);
  const uint32_t E = (bits >> 23) & 0xFFu;
  if ((bits >> 31) || E < 24u || E > 253u) return false;       // o > 0, normal, ulp(o) a normal number too
  const float q = d * pt_bits_f32((277u - E) << 23);            // d / ulp(o): exact scaling by a power of two
  if (!(q > -4194304.0f && q < 4194304.0f)) return false;
#if defined(__CUDA_ARCH__)
  const float c = rintf(q);
#else
  const float c = __builtin_rintf(q);
#endif
  const float f = q - c;
  if (f == 0.5f || f == -0.5f) return false;
  const int64_t M = (int64_t)((bits & 0x7FFFFFu) | 0x800000u);
  const int64_t M1 = M + (int64_t)n * (int64_t)c;
  if (M1 > 0xFFFFFF) return false;                               // would cross into the next binade
  if (q < 0.0f && (M < 0x800001 || M1 < 0x800001)) return false; // a sum just below 2^E rounds on the finer grid
  o = pt_bits_f32((bits & 0xFF800000u) | ((uint32_t)M1 & 0x7FFFFFu));
  return true;
}

}  // namespace pb
