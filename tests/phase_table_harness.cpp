// Test harness (CPU): the integer model + jump tables of phonic_b200/csrc/phase_table.cuh against the LITERAL f32
// recurrence of CubicInterpolator::process (src/utils/resampler/cubic.rs:72-110), written here independently.
// Built by tests/test_phase_table.py with g++ -O2 -ffp-contract=off.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../phonic_b200/csrc/phase_table.cuh"

using namespace pb;

// one output frame of cubic.rs: returns samples pushed
static inline uint32_t literal_frame(float& sub_pos, float ratio) {
  uint32_t pushed = 0;
  if (ratio < 1.0f) {
    if (sub_pos >= 1.0f) { ++pushed; sub_pos -= 1.0f; }
    sub_pos += ratio;
  } else {
    while (sub_pos < ratio) { ++pushed; sub_pos += 1.0f; }
    sub_pos -= ratio;
  }
  return pushed;
}

extern "C" {

// words of the table for `ratio`; fills geometry
uint32_t pt_words(float ratio, uint32_t* geom7) {
  PhaseGeom g = phase_geom(ratio);
  if (geom7) { geom7[0] = g.mode; geom7[1] = g.sh; geom7[2] = g.R; geom7[3] = g.L; geom7[4] = g.n_bp; geom7[5] = g.margin; geom7[6] = g.nth; }
  return phase_table_words(g);
}

void pt_build(float ratio, uint32_t* tab) {
  PhaseGeom g = phase_geom(ratio);
  phase_build_header(tab, ratio, g);
  if (g.mode != PT_DOWN_TABLE && g.mode != PT_UP_TABLE) return;
  std::vector<uint32_t> scratch(PT_MAX_BP);
  const uint32_t nt = 8;
  for (uint32_t t = 0; t < nt; ++t) phase_build_p1(g, scratch.data(), t, nt);
  for (uint32_t t = 0; t < nt; ++t) phase_build_p2(tab, g, scratch.data(), t, nt);
  for (uint32_t t = 0; t < nt; ++t) phase_build_p3(tab, g, t, nt);
  for (uint32_t t = 0; t < nt; ++t) phase_build_p4(tab, g, t, nt);
}

// The integer model against the literal float loop, one frame at a time, from `n` start states (as float sub_pos
// values on the ratio's grid). Returns the number of mismatching frames.
uint64_t pt_check_model(float ratio, const float* starts, uint32_t n, uint32_t frames) {
  PhaseGeom g = phase_geom(ratio);
  if (g.mode == PT_LITERAL) return 0;
  const uint32_t ONE = 1u << g.sh;
  const float scale = pt_bits_f32((127u + g.sh) << 23), inv = pt_bits_f32((127u - g.sh) << 23);
  const bool down = ratio < 1.0f;
  uint64_t bad = 0;
  for (uint32_t i = 0; i < n; ++i) {
    float s = starts[i];
    uint32_t S = (uint32_t)(s * scale);
    if ((float)S * inv != s) { ++bad; continue; }
    for (uint32_t f = 0; f < frames; ++f) {
      uint32_t W = 0; uint64_t sig = 0;
      const uint32_t pushed = literal_frame(s, ratio);
      if (down) phase_step_down(S, W, sig, ONE, g.R); else phase_step_up(S, W, sig, ONE, g.R);
      if (pushed != W || (float)S * inv != s || (uint32_t)(s * scale) != S) { ++bad; break; }
    }
  }
  return bad;
}

// Jumps against the literal loop: from each start state walk `tiles` tiles; on every tile compare phase_jump with 64
// literal frames. out[0] += tiles where the jump applied, out[1] += tiles where it declined, returns mismatches.
uint64_t pt_check_jumps(float ratio, const uint32_t* tab, const float* starts, uint32_t n, uint32_t tiles, uint64_t* out) {
  uint64_t bad = 0;
  for (uint32_t i = 0; i < n; ++i) {
    float s = starts[i];
    for (uint32_t t = 0; t < tiles; ++t) {
      float sl = s;
      uint32_t wl = 0;
      for (uint32_t f = 0; f < PT_K; ++f) wl += literal_frame(sl, ratio);
      float sj = s;
      uint32_t wj = 0;
      if (phase_jump(tab, sj, wj)) {
        out[0]++;
        if (wj != wl || std::memcmp(&sj, &sl, 4) != 0) { ++bad; if (bad < 4) { out[2] = i; out[3] = t; } }
      } else {
        out[1]++;
      }
      s = sl;
    }
  }
  return bad;
}

// invalid entries / intervals of a built table (diagnostics)
void pt_table_stats(const uint32_t* tab, uint64_t* out) {
  const uint32_t mode = tab[PT_H_MODE];
  out[0] = out[1] = out[2] = out[3] = 0;
  if (mode != PT_DOWN_TABLE && mode != PT_UP_TABLE) return;
  const uint32_t n_bp = tab[PT_H_NBP], L = tab[PT_H_LMASK] + 1;
  const uint32_t* lo = tab + PT_HEADER + PT_COARSE + 1 + n_bp + 2;
  const uint32_t* hi = lo + n_bp + 1;
  const uint32_t* entry = hi + n_bp + 1;
  for (uint32_t i = 0; i <= n_bp; ++i) {
    if (lo[i] > hi[i]) { out[0]++; continue; }
    out[1] += (uint64_t)hi[i] - lo[i] + 1;     // states covered by a valid interval
    for (uint32_t r = 0; r < L; ++r) { out[2]++; if (!(entry[i * L + r] & 0x80000000u)) out[3]++; }
  }
}

// accum_jump against the literal chain; returns mismatches, out[0] += applied, out[1] += declined
uint64_t pt_check_accum(const float* o, const float* d, const uint32_t* n, uint32_t count, uint64_t* out) {
  uint64_t bad = 0;
  for (uint32_t i = 0; i < count; ++i) {
    volatile float lit = o[i];
    for (uint32_t k = 0; k < n[i]; ++k) lit = lit + d[i];
    float j = o[i];
    if (accum_jump(j, d[i], n[i])) {
      out[0]++;
      float l2 = lit;
      if (std::memcmp(&j, &l2, 4) != 0) { ++bad; out[2] = i; }
    } else {
      out[1]++;
    }
  }
  return bad;
}

}  // extern "C"
