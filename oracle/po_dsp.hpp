// ORACLE -- TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement (scalar C++, f32/f64 exactly as the reference, operation order identical,
// compile with -ffp-contract=off) of emuell/phonic v0.16.0's DSP primitives on the offline
// render path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may build, load or call anything under oracle/.
//
// Parity pinning: the reference ships no golden audio for this path; this file
// is pinned against the reference's own unit-test assertions and the derived known-answer
// vectors of SURVEY.md Appendix D (tests/test_oracle_kat.py). The Rust reference cannot be
// built in this environment (no cargo), so it was never run here.
//
// Every struct cites the reference file:line it follows (paths relative to the reference root).
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <optional>
#include <vector>

namespace po {

constexpr float F32_EPSILON = 1.1920929e-07f;  // f32::EPSILON
constexpr size_t USIZE_MAX = std::numeric_limits<size_t>::max();

// std::time::Duration: (secs: u64, nanos: u32). as_secs_f32 = secs as f32 + nanos as f32 / 1e9.
struct Duration {
  uint64_t secs = 0;
  uint32_t nanos = 0;
  static Duration from_nanos(uint64_t n) { return {n / 1000000000ull, (uint32_t)(n % 1000000000ull)}; }
  static Duration from_millis(uint64_t ms) { return from_nanos(ms * 1000000ull); }
  static Duration from_secs(uint64_t s) { return {s, 0}; }
  // Duration::from_secs_f32: the exact value rounded to the nearest nanosecond (core::time, try_from_secs_f32)
  static Duration from_secs_f32(float s) { return from_nanos((uint64_t)std::nearbyint((double)s * 1.0e9)); }
  bool is_zero() const { return secs == 0 && nanos == 0; }
  float as_secs_f32() const { return (float)secs + (float)nanos / 1.0e9f; }
  double as_secs_f64() const { return (double)secs + (double)nanos / 1.0e9; }
};

// src/utils.rs:41-51
inline float db_to_linear(float value) {
  const float DB_TO_LIN_FACTOR = 2.302585092994046f / 20.0f;  // LN_10 / 20
  if (std::isnan(value)) return NAN;
  if (value == 0.0f) return 1.0f;
  if (value > -200.0f) return std::exp(value * DB_TO_LIN_FACTOR);
  return 0.0f;
}
// src/utils.rs:26-38
inline float linear_to_db(float value) {
  const float LIN_TO_DB_FACTOR = 20.0f / 2.302585092994046f;
  if (value < 0.0f || std::isnan(value)) return NAN;
  if (value == 1.0f) return 0.0f;
  if (value > 1e-12f) return std::log(value) * LIN_TO_DB_FACTOR;
  return -200.0f;
}
// src/utils.rs:56-62
inline void panning_factors(float pan, float& l, float& r) {
  const float POWER = 0.70710678118654752440f;  // FRAC_1_SQRT_2
  float normalized = (std::min(std::max(pan, -1.0f), 1.0f) + 1.0f) / 2.0f;
  l = std::sqrt(1.0f - normalized) / POWER;
  r = std::sqrt(normalized) / POWER;
}
// src/utils.rs:67-78
inline double pitch_from_note(uint8_t note) { return 440.0 * std::pow(2.0, ((double)note - 69.0) / 12.0); }
inline double speed_from_note(uint8_t note) { return pitch_from_note(note) / pitch_from_note(60); }

// ---- buffer ops: src/utils/buffer.rs:86-173 (lane-wise, order independent) ------------------
inline void clear_buffer(float* b, size_t n) { std::fill(b, b + n, 0.0f); }
inline void scale_buffer(float* b, size_t n, float g) { for (size_t i = 0; i < n; ++i) b[i] *= g; }
inline void add_buffers(float* d, const float* s, size_t n) { for (size_t i = 0; i < n; ++i) d[i] += s[i]; }
inline float max_abs_sample(const float* b, size_t n) {
  float m = 0.0f;
  for (size_t i = 0; i < n; ++i) m = std::max(m, std::fabs(b[i]));
  return m;
}

// ---- src/utils/smoothing.rs:131-228 -----------------------------------------------------------
struct ExpSmoothed {
  float current = 0, target = 0, inertia = 1.0f / 256.0f, sample_rate_comp = 44100.0f / 66666.0f;
  ExpSmoothed() {}
  ExpSmoothed(float v, uint32_t sr) : current(v), target(v), sample_rate_comp(44100.0f / (float)sr) {}
  bool need_ramp() const {
    const float EPS = F32_EPSILON * 100.0f;
    float add = (target - current) * inertia * sample_rate_comp;
    return std::fabs(add) > EPS;
  }
  void ramp() { current += (target - current) * inertia * sample_rate_comp; }
  float next() {  // smoothing.rs:21-28
    if (need_ramp()) { ramp(); return current; }
    return target;
  }
  void init(float v) { target = v; current = v; }
  void set_target(float t) { target = t; if (!need_ramp()) current = target; }
  void set_sample_rate(uint32_t sr) { sample_rate_comp = 44100.0f / (float)sr; }
};

// ---- src/utils/smoothing.rs:247-402 -----------------------------------------------------------
struct LinearSmoothed {
  float current = 0, target = 0, step = 0.01f, current_step = 0;
  uint32_t num_pending_steps = 0;
  float sample_rate_comp = 44100.0f / 66666.0f;
  LinearSmoothed() {}
  LinearSmoothed(float v, float step_) : current(v), target(v), step(step_) {}
  bool need_ramp() const { return num_pending_steps > 0; }
  void ramp() {
    if (num_pending_steps > 0) {
      current += current_step;
      num_pending_steps -= 1;
      if (num_pending_steps == 0) current = target;
    }
  }
  float next() { if (need_ramp()) { ramp(); return current; } return target; }
  void init(float v) { target = v; current = v; num_pending_steps = 0; }
  static uint32_t round_to_u32(float x) {  // f32::round().max(0.0) as u32 (saturating)
    float r = std::max(std::round(x), 0.0f);
    if (std::isnan(r)) return 0;
    if (r >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)r;
  }
  void set_target(float t) {  // set_target_with_duration(target, None)
    target = t;
    if (current == target) { num_pending_steps = 0; return; }
    current_step = (current > target) ? -step * sample_rate_comp : step * sample_rate_comp;
    float pending = (target - current) / current_step;
    num_pending_steps = round_to_u32(pending);
    if (num_pending_steps == 0) current = target;
  }
  void set_sample_rate(uint32_t sr) {
    sample_rate_comp = 44100.0f / (float)sr;
    current_step = (current > target) ? -step * sample_rate_comp : step * sample_rate_comp;
  }
};

// ---- src/utils/smoothing.rs:424-534 -----------------------------------------------------------
struct SpringSmoothed {
  float current = 0, velocity = 0, target = 0, omega = 5.5f / 4410.0f;
  float sample_rate_comp = 44100.0f / 66666.0f;
  SpringSmoothed() {}
  SpringSmoothed(float v, size_t duration) : current(v), target(v), omega(5.5f / (float)duration) {}
  bool need_ramp() const {
    const float EPS = F32_EPSILON * 100.0f;
    return std::fabs(velocity) > EPS || std::fabs(target - current) > EPS;
  }
  void ramp() {
    float om = omega * sample_rate_comp;
    float k = om * om;
    float d = 2.0f * om;
    velocity += (target - current) * k - velocity * d;
    current += velocity;
  }
  float next() { if (need_ramp()) { ramp(); return current; } return target; }
  void init(float v) { current = v; velocity = 0; target = v; }
  void set_target(float v) { target = v; }
  void set_sample_rate(uint32_t sr) { sample_rate_comp = 44100.0f / (float)sr; }
};

// src/utils/smoothing.rs:60-71 -- NB: advances per *sample*, not per frame.
template <class S>
inline void apply_smoothed_gain(float* b, size_t n, S& sm) {
  if (sm.need_ramp()) {
    for (size_t i = 0; i < n; ++i) b[i] *= sm.next();
  } else {
    float g = sm.target;
    if (std::fabs(1.0f - g) > 0.000001f) scale_buffer(b, n, g);
  }
}
// src/utils/smoothing.rs:74-122
template <class S>
inline void apply_smoothed_panning(float* b, size_t n, size_t ch, S& sm) {
  if (ch < 2) return;
  if (sm.need_ramp()) {
    for (size_t i = 0; i + ch <= n; i += ch) {
      float l, r;
      panning_factors(sm.next(), l, r);
      b[i] *= l;
      b[i + 1] *= r;
    }
  } else {
    float pan = sm.target;
    if (std::fabs(pan) > 0.000001f) {
      float l, r;
      panning_factors(pan, l, r);
      for (size_t i = 0; i + ch <= n; i += ch) { b[i] *= l; b[i + 1] *= r; }
    }
  }
}

// ---- src/utils/fader.rs:27-123 ------------------------------------------------------------------
struct VolumeFader {
  enum State { Stopped, IsRunning, Finished };
  State state = Stopped;
  float current_volume = 1, target_volume = 1, inertia = 1;
  size_t channel_count = 1;
  uint32_t sample_rate = 44100;
  VolumeFader() {}
  VolumeFader(size_t ch, uint32_t sr) : channel_count(ch), sample_rate(sr) {}
  void start(float from, float to, Duration d) {
    if (d.is_zero()) {
      current_volume = to; target_volume = to; state = Finished;
    } else {
      state = IsRunning; current_volume = from; target_volume = to;
      const float LN100 = 4.605f;
      float samples_duration = (float)sample_rate * d.as_secs_f32() / LN100;
      inertia = 1.0f - std::exp(-1.0f / samples_duration);
    }
  }
  void start_fade_in(Duration d) { if (state == IsRunning) start(current_volume, 1.0f, d); else start(0.0f, 1.0f, d); }
  void start_fade_out(Duration d) { if (state == IsRunning) start(current_volume, 0.0f, d); else start(1.0f, 0.0f, d); }
  void reset() { state = Stopped; current_volume = 1; target_volume = 1; }
  void process(float* out, size_t n) {
    if (state != IsRunning) {
      if (target_volume != 1.0f) scale_buffer(out, n, target_volume);
    } else {
      for (size_t i = 0; i + channel_count <= n; i += channel_count) {
        current_volume += (target_volume - current_volume) * inertia;
        for (size_t c = 0; c < channel_count; ++c) out[i + c] *= current_volume;
      }
      if (std::fabs(current_volume - target_volume) < 0.0001f) state = Finished;
    }
  }
};

// ---- src/utils/ahdsr.rs:26-346 -------------------------------------------------------------------
struct AhdsrParameters {
  uint32_t sample_rate = 0;
  Duration attack_time, hold_time, decay_time, release_time;
  float attack_scaling = 0, attack_rate = 0, decay_scaling = 0, decay_rate = 0, sustain_level = 0;
  float release_scaling = 0, release_rate = 0;
  static constexpr uint32_t UNINITIALIZED_SAMPLE_RATE = 66666;
  static constexpr float F32_MAX = std::numeric_limits<float>::max();

  void set_attack_time(Duration t) {
    attack_time = t;
    float s = t.as_secs_f32();
    attack_rate = (s == 0.0f) ? F32_MAX : 1.0f / (s * (float)sample_rate);
  }
  void set_hold_time(Duration t) { hold_time = t; }
  void set_decay_time(Duration t) {
    decay_time = t;
    decay_rate = t.is_zero() ? F32_MAX : (1.0f - sustain_level) / (t.as_secs_f32() * (float)sample_rate);
  }
  bool set_sustain_level(float l) { if (!(l >= 0.0f && l <= 1.0f)) return false; sustain_level = l; return true; }
  void set_release_time(Duration t) {
    release_time = t;
    float s = t.as_secs_f32();
    release_rate = (s == 0.0f) ? F32_MAX : 1.0f / (s * (float)sample_rate);
  }
  // setup_with_scaling, ahdsr.rs:293-313 (order matters: decay rate uses the *previous* sustain)
  bool setup_with_scaling(Duration a, float as, Duration h, Duration d, float ds, float s, Duration r, float rs) {
    set_attack_time(a);
    if (!(as >= -1.0f && as <= 1.0f)) return false;
    attack_scaling = as;
    set_hold_time(h);
    set_decay_time(d);
    if (!(ds >= -1.0f && ds <= 1.0f)) return false;
    decay_scaling = ds;
    if (!set_sustain_level(s)) return false;
    set_release_time(r);
    if (!(rs >= -1.0f && rs <= 1.0f)) return false;
    release_scaling = rs;
    return true;
  }
  bool setup(Duration a, Duration h, Duration d, float s, Duration r) {  // ahdsr.rs:273-287
    set_attack_time(a); set_hold_time(h); set_decay_time(d);
    if (!set_sustain_level(s)) return false;
    set_release_time(r);
    return true;
  }
  // new_with_scaling, ahdsr.rs:75-98
  static bool create(AhdsrParameters& p, Duration a, float as, Duration h, Duration d, float ds, float s, Duration r, float rs) {
    p = AhdsrParameters();
    p.sample_rate = UNINITIALIZED_SAMPLE_RATE;
    return p.setup_with_scaling(a, as, h, d, ds, s, r, rs);
  }
  bool set_sample_rate(uint32_t sr) {  // ahdsr.rs:123-136
    if (sample_rate != sr) {
      sample_rate = sr;
      return setup(attack_time, hold_time, decay_time, sustain_level, release_time);
    }
    return true;
  }
  static float apply_scaling(float value, float scaling) {  // ahdsr.rs:324-345
    const float EULER_DIV_2 = 2.718281828459045f / 2.0f;
    if (scaling == 0.0f || value == 0.0f) return value;
    float s = -scaling;
    if (s > 0.0f) return std::pow(value, 1.0f + std::pow(s, EULER_DIV_2) * 16.0f);
    return 1.0f - std::pow(1.0f - value, 1.0f + std::pow(-s, EULER_DIV_2) * 16.0f);
  }
};

// ---- src/utils/ahdsr.rs:367-580 -----------------------------------------------------------------
struct AhdsrEnvelope {
  enum Stage { Idle = 0, Attack = 1, Hold = 2, Decay = 3, Sustain = 4, Release = 5 };
  Stage stage = Idle;
  float target_volume = 0, hold_samples_remaining = 0, release_output = 0, output = 0;
  static constexpr float SILENCE = 0.001f;

  void note_on(const AhdsrParameters& p, float volume) {
    target_volume = volume;
    if (p.attack_rate == AhdsrParameters::F32_MAX) {
      output = volume;
      if (!p.hold_time.is_zero()) {
        stage = Hold;
        hold_samples_remaining = p.hold_time.as_secs_f32() * (float)p.sample_rate;
      } else {
        stage = Decay;
      }
    } else {
      output = 0.0f;
      stage = Attack;
    }
  }
  void note_off(const AhdsrParameters& p) {
    if (!p.release_time.is_zero()) {
      target_volume = 0.0f;
      release_output = output;
      stage = (release_output > F32_EPSILON) ? Release : Idle;
    } else {
      output = 0.0f; release_output = 0.0f; stage = Idle;
    }
  }
  void reset() { output = 0.0f; stage = Idle; }
  float run(const AhdsrParameters& p) {
    switch (stage) {
      case Attack:
        output += p.attack_rate;
        if (output >= target_volume) {
          output = target_volume;
          target_volume = p.sustain_level;
          if (!p.hold_time.is_zero()) {
            stage = Hold;
            hold_samples_remaining = p.hold_time.as_secs_f32() * (float)p.sample_rate;
          } else {
            stage = Decay;
          }
        }
        break;
      case Hold:
        hold_samples_remaining -= 1.0f;
        if (hold_samples_remaining <= 0.0f) stage = p.decay_time.is_zero() ? Sustain : Decay;
        break;
      case Decay:
        if (output > p.sustain_level) {
          output -= p.decay_rate;
          if (output <= p.sustain_level) { output = p.sustain_level; stage = Sustain; }
        } else {
          output += p.decay_rate;
          if (output >= p.sustain_level) { output = p.sustain_level; stage = Sustain; }
        }
        break;
      case Sustain: break;
      case Release:
        output -= release_output * p.release_rate;
        if (output <= SILENCE) { output = 0.0f; stage = Idle; }
        break;
      case Idle: break;
    }
    // scaling (ahdsr.rs:518-551)
    if (stage == Attack && p.attack_scaling != 0.0f) {
      float progress = output / std::max(target_volume, F32_EPSILON);
      return AhdsrParameters::apply_scaling(progress, p.attack_scaling) * target_volume;
    }
    if (stage == Decay && p.decay_scaling != 0.0f) {
      float range = std::max(std::fabs(target_volume - p.sustain_level), F32_EPSILON);
      float progress = (target_volume > p.sustain_level) ? (target_volume - output) / range
                                                         : (output - target_volume) / range;
      float sp = AhdsrParameters::apply_scaling(progress, p.decay_scaling);
      return (target_volume > p.sustain_level) ? target_volume - (sp * range) : target_volume + (sp * range);
    }
    if (stage == Release && p.release_scaling != 0.0f) {
      float initial = std::max(output, F32_EPSILON);
      float progress = 1.0f - (output / initial);
      float sp = AhdsrParameters::apply_scaling(progress, p.release_scaling);
      return initial * (1.0f - sp);
    }
    return output;
  }
};

// ---- src/utils/resampler/cubic.rs:10-143 ---------------------------------------------------------
struct CubicInterpolator {
  float input[4] = {0, 0, 0, 0};
  float sub_pos = 0, ratio = 1;
  bool is_initialized = false;
  void reset() { input[0] = input[1] = input[2] = input[3] = 0; sub_pos = 0; is_initialized = false; }
  void push_sample(float v) { input[3] = input[2]; input[2] = input[1]; input[1] = input[0]; input[0] = v; }
  float interpolate(float fraction) const {
    float ym1 = input[3], y0 = input[2], y1 = input[1], y2 = input[0];
    float c0 = y0;
    float c1 = (y1 - ym1) * 0.5f;
    float c2 = ym1 - y0 * 2.5f + y1 * 2.0f - y2 * 0.5f;
    float c3 = (y2 - ym1) * 0.5f + (y0 - y1) * 1.5f;
    return ((c3 * fraction + c2) * fraction + c1) * fraction + c0;
  }
  // returns (consumed samples, produced samples), cubic.rs:36-114
  std::pair<size_t, size_t> process(const float* in, size_t in_len, float* out, size_t out_len,
                                    size_t ci, size_t cc) {
    size_t num_in = in_len / cc, num_out = out_len / cc;
    size_t consumed = 0, produced = 0;
    if (std::fabs(ratio - 1.0f) < 0.000001f) {
      size_t m = std::min(in_len, out_len);
      std::memcpy(out, in, m * sizeof(float));  // copies all channels (each interpolator repeats it)
      return {m, m};
    }
    if (!is_initialized && num_in >= 3) {
      is_initialized = true;
      for (size_t f = 0; f < 3; ++f) { push_sample(in[f * cc + ci]); consumed += 1; }
    }
    if (ratio < 1.0f) {
      while (produced < num_out) {
        if (sub_pos >= 1.0f) {
          if (consumed >= num_in) break;
          push_sample(in[consumed * cc + ci]);
          consumed += 1;
          sub_pos -= 1.0f;
        }
        out[produced * cc + ci] = interpolate(sub_pos);
        produced += 1;
        sub_pos += ratio;
      }
    } else {
      bool stop = false;
      while (!stop && produced < num_out) {
        while (sub_pos < ratio) {
          if (consumed >= num_in) { stop = true; break; }
          push_sample(in[consumed * cc + ci]);
          consumed += 1;
          sub_pos += 1.0f;
        }
        if (stop) break;
        sub_pos -= ratio;
        out[produced * cc + ci] = interpolate(1.0f - sub_pos);
        produced += 1;
      }
    }
    return {consumed * cc, produced * cc};
  }
};

// ---- src/utils/resampler/cubic.rs:150-207, src/utils/resampler.rs:12-35 -------------------------
struct CubicResampler {
  uint32_t input_rate, output_rate;
  size_t channel_count;
  std::vector<CubicInterpolator> interpolators;
  CubicResampler(uint32_t in_rate, uint32_t out_rate, size_t cc)
      : input_rate(in_rate), output_rate(out_rate), channel_count(cc), interpolators(cc) {
    float r = (float)((double)in_rate / (double)out_rate);
    for (auto& i : interpolators) i.ratio = r;
  }
  std::pair<size_t, size_t> process(const float* in, size_t in_len, float* out, size_t out_len) {
    std::pair<size_t, size_t> res{0, 0};
    for (size_t c = 0; c < channel_count; ++c) res = interpolators[c].process(in, in_len, out, out_len, c, channel_count);
    return res;
  }
  void update(uint32_t in_rate, uint32_t out_rate) {
    input_rate = in_rate; output_rate = out_rate;
    float r = (float)((double)in_rate / (double)out_rate);
    for (auto& i : interpolators) i.ratio = r;
  }
  void reset() { for (auto& i : interpolators) i.reset(); }
};

// ---- rubato ^0.16 `SincFixedIn<f32>` (THIRD-PARTY crate, not vendored under /root/reference) ---------
// PARITY UNPINNED: restated from rubato 0.16's published algorithm (src/sinc.rs make_sincs, src/windows.rs
// blackman_harris / make_window, src/sinc_interpolator/mod.rs ScalarInterpolator::get_sinc_interpolated,
// src/asynchro_sinc.rs SincFixedIn::{new, process_into_buffer}, get_nearest_times_4, interp_cubic). Neither
// the crate nor any golden vector of it is available in this environment; the reference's only test at this
// boundary (preloaded.rs:513-532) takes the equal-rate bypass and never enters rubato. rubato picks an
// AVX/SSE/NEON interpolator at run time whose summation order differs from the scalar one restated here, so
// even reference-vs-reference is only ~1e-6 exact. Call sites: src/utils/resampler/rubato.rs:22-56,100-103.
struct RubatoSincTable {
  size_t sinc_len = 0, factor = 0;
  std::vector<float> sincs;  // [factor][sinc_len]
  const float* row(size_t sub) const { return sincs.data() + sub * sinc_len; }
  // windows.rs: blackman_harris (periodic, f32) squared for BlackmanHarris2
  static std::vector<float> blackman_harris2(size_t npoints) {
    std::vector<float> w(npoints);
    const float PI = 3.14159265358979323846264338327950288f;
    const float pi2 = 2.0f * PI, pi4 = 4.0f * PI, pi6 = 6.0f * PI;
    const float np_f = (float)npoints;
    const float a = 0.35875f, b = 0.48829f, c = 0.14128f, d = 0.01168f;
    for (size_t x = 0; x < npoints; ++x) {
      const float xf = (float)x;
      float v = a - b * std::cos(pi2 * xf / np_f) + c * std::cos(pi4 * xf / np_f) - d * std::cos(pi6 * xf / np_f);
      w[x] = v * v;
    }
    return w;
  }
  static float sinc(float value) {
    const float PI = 3.14159265358979323846264338327950288f;
    if (value == 0.0f) return 1.0f;
    return std::sin(value * PI) / (value * PI);
  }
  // sinc.rs make_sincs(npoints, factor, f_cutoff, window)
  void make(size_t npoints, size_t fac, float f_cutoff) {
    sinc_len = npoints; factor = fac;
    const size_t totpoints = npoints * fac;
    std::vector<float> y(totpoints);
    std::vector<float> window = blackman_harris2(totpoints);
    float sum = 0.0f;
    for (size_t x = 0; x < totpoints; ++x) {
      float val = window[x] * sinc(((float)x - (float)(totpoints / 2)) * f_cutoff / (float)fac);
      sum += val;
      y[x] = val;
    }
    sum /= (float)fac;
    sincs.assign(fac * npoints, 0.0f);
    for (size_t p = 0; p < npoints; ++p)
      for (size_t n = 0; n < fac; ++n) sincs[(fac - n - 1) * npoints + p] = y[fac * p + n] / sum;
  }
  // one table per cutoff (make_interpolator: f_cutoff scaled by the ratio when downsampling)
  static std::shared_ptr<RubatoSincTable> get(size_t sinc_len, double resample_ratio, float f_cutoff, size_t factor) {
    static std::vector<std::pair<uint32_t, std::shared_ptr<RubatoSincTable>>> cache;
    sinc_len = 8 * (size_t)std::ceil((float)sinc_len / 8.0f);
    float fc = resample_ratio >= 1.0 ? f_cutoff : f_cutoff * (float)resample_ratio;
    uint32_t key;
    std::memcpy(&key, &fc, 4);
    for (auto& e : cache) if (e.first == key && e.second->sinc_len == sinc_len && e.second->factor == factor) return e.second;
    auto t = std::make_shared<RubatoSincTable>();
    t->make(sinc_len, factor, fc);
    cache.emplace_back(key, t);
    return t;
  }
  // ScalarInterpolator::get_sinc_interpolated: 8 interleaved accumulators, summed acc0+...+acc7
  float dot(const float* wave, size_t index, size_t subindex) const {
    const float* w = wave + index;
    const float* s = row(subindex);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < sinc_len; i += 8)
      for (size_t k = 0; k < 8; ++k) acc[k] += w[i + k] * s[i + k];
    return acc[0] + acc[1] + acc[2] + acc[3] + acc[4] + acc[5] + acc[6] + acc[7];
  }
};

struct RubatoSincFixedIn {
  size_t nbr_channels, chunk_size;
  double last_index, resample_ratio, resample_ratio_original, target_ratio, max_relative_ratio;
  std::shared_ptr<RubatoSincTable> interpolator;
  std::vector<std::vector<float>> buffer;  // [ch][chunk_size + 2 * sinc_len]
  RubatoSincFixedIn(double ratio, double max_rel, size_t sinc_len, float f_cutoff, size_t oversampling, size_t chunk, size_t ch)
      : nbr_channels(ch), chunk_size(chunk), resample_ratio(ratio), resample_ratio_original(ratio), target_ratio(ratio),
        max_relative_ratio(max_rel), interpolator(RubatoSincTable::get(sinc_len, ratio, f_cutoff, oversampling)) {
    buffer.assign(ch, std::vector<float>(chunk + 2 * interpolator->sinc_len, 0.0f));
    last_index = -(double)(interpolator->sinc_len / 2);
  }
  size_t output_frames_max() const { return (size_t)((double)chunk_size * resample_ratio_original * max_relative_ratio + 10.0); }
  bool set_resample_ratio(double new_ratio, bool ramp) {
    if (new_ratio / resample_ratio_original >= 1.0 / max_relative_ratio && new_ratio / resample_ratio_original <= max_relative_ratio) {
      if (!ramp) resample_ratio = new_ratio;
      target_ratio = new_ratio;
      return true;
    }
    return false;  // ResampleError::RatioOutOfBounds
  }
  static void get_nearest_times_4(double t, long factor, long (*points)[2]) {
    long index = (long)std::floor(t);
    long frac = (long)std::floor((t - std::floor(t)) * (double)factor);
    if (frac == 0) {
      points[0][0] = index - 1; points[0][1] = factor - 1;
      points[1][0] = index; points[1][1] = 0;
      points[2][0] = index; points[2][1] = 1;
      points[3][0] = index; points[3][1] = 2;
    } else if (frac == factor - 2) {
      points[0][0] = index; points[0][1] = frac - 1;
      points[1][0] = index; points[1][1] = frac;
      points[2][0] = index; points[2][1] = frac + 1;
      points[3][0] = index + 1; points[3][1] = 0;
    } else if (frac == factor - 1) {
      points[0][0] = index; points[0][1] = frac - 1;
      points[1][0] = index; points[1][1] = frac;
      points[2][0] = index + 1; points[2][1] = 0;
      points[3][0] = index + 1; points[3][1] = 1;
    } else {
      for (long k = 0; k < 4; ++k) { points[k][0] = index; points[k][1] = frac - 1 + k; }
    }
  }
  static float interp_cubic(float x, const float* y) {
    float a0 = y[1];
    float a1 = -(1.0f / 3.0f) * y[0] - 0.5f * y[1] + y[2] - (1.0f / 6.0f) * y[3];
    float a2 = 0.5f * (y[0] + y[2]) - y[1];
    float a3 = 0.5f * (y[1] - y[2]) + (1.0f / 6.0f) * (y[3] - y[0]);
    float x2 = x * x;
    float x3 = x2 * x;
    return a0 + a1 * x + a2 * x2 + a3 * x3;
  }
  // process_into_buffer (cubic interpolation arm); wave_in[ch][chunk_size], returns frames written per channel
  size_t process_into_buffer(const std::vector<std::vector<float>>& wave_in, std::vector<std::vector<float>>& wave_out) {
    const size_t sinc_len = interpolator->sinc_len;
    const long oversampling = (long)interpolator->factor;
    double t_ratio = 1.0 / resample_ratio;
    const double t_ratio_end = 1.0 / target_ratio;
    const double approximate_nbr_frames = (double)chunk_size * (0.5 * resample_ratio + 0.5 * target_ratio);
    const double t_ratio_increment = (t_ratio_end - t_ratio) / approximate_nbr_frames;
    const long end_idx = (long)chunk_size - ((long)sinc_len + 1) - (long)std::ceil(t_ratio_end);
    for (auto& buf : buffer) std::memmove(buf.data(), buf.data() + chunk_size, 2 * sinc_len * sizeof(float));
    for (size_t c = 0; c < nbr_channels; ++c) std::memcpy(buffer[c].data() + 2 * sinc_len, wave_in[c].data(), chunk_size * sizeof(float));
    double idx = last_index;
    size_t n = 0;
    float points[4];
    long nearest[4][2];
    while (idx < (double)end_idx) {
      t_ratio += t_ratio_increment;
      idx += t_ratio;
      get_nearest_times_4(idx, oversampling, nearest);
      double frac = idx * (double)oversampling - std::floor(idx * (double)oversampling);
      float frac_offset = (float)frac;
      for (size_t c = 0; c < nbr_channels; ++c) {
        for (int k = 0; k < 4; ++k)
          points[k] = interpolator->dot(buffer[c].data(), (size_t)(nearest[k][0] + 2 * (long)sinc_len), (size_t)nearest[k][1]);
        wave_out[c][n] = interp_cubic(frac_offset, points);
      }
      n += 1;
    }
    last_index = idx - (double)chunk_size;
    resample_ratio = target_ratio;
    return n;
  }
};

// ---- trait AudioResampler (src/utils/resampler.rs:43-64) ----------------------------------------------
struct AudioResampler {
  virtual ~AudioResampler() {}
  virtual size_t required_input_buffer_size() const = 0;  // 0 = None
  virtual size_t max_input_buffer_size() const = 0;       // 0 = None
  virtual std::pair<size_t, size_t> process(const float* in, size_t in_len, float* out, size_t out_len) = 0;
  virtual bool update(uint32_t in_rate, uint32_t out_rate) = 0;  // false = Err (the caller `expect`s)
  virtual void reset() = 0;
};
struct CubicAudioResampler : AudioResampler {
  CubicResampler r;
  CubicAudioResampler(uint32_t in_rate, uint32_t out_rate, size_t cc) : r(in_rate, out_rate, cc) {}
  size_t required_input_buffer_size() const override { return 0; }
  size_t max_input_buffer_size() const override { return 0; }
  std::pair<size_t, size_t> process(const float* in, size_t in_len, float* out, size_t out_len) override { return r.process(in, in_len, out, out_len); }
  bool update(uint32_t in_rate, uint32_t out_rate) override { r.update(in_rate, out_rate); return true; }
  void reset() override { r.reset(); }
};

// ---- RubatoResampler (src/utils/resampler/rubato.rs:12-154) -----------------------------------------
struct RubatoResampler : AudioResampler {
  uint32_t input_rate, output_rate;
  size_t channel_count;
  RubatoSincFixedIn resampler;
  std::vector<std::vector<float>> input, output;
  std::vector<float> pending;          // TempBuffer (utils/buffer.rs:499-610)
  size_t pending_start = 0, pending_end = 0;
  static constexpr size_t CHUNK_SIZE = 256;
  RubatoResampler(uint32_t in_rate, uint32_t out_rate, size_t cc)
      : input_rate(in_rate), output_rate(out_rate), channel_count(cc),
        resampler((double)out_rate / (double)in_rate, 1.0, 256, 0.95f, 128, CHUNK_SIZE, cc) {
    input.assign(cc, std::vector<float>(CHUNK_SIZE, 0.0f));                   // input_buffer_allocate(true)
    output.assign(cc, std::vector<float>(resampler.output_frames_max(), 0.0f)); // output_buffer_allocate(true)
    pending.assign(cc * resampler.output_frames_max(), 0.0f);
  }
  size_t required_input_buffer_size() const override { return CHUNK_SIZE * channel_count; }  // input_frames_next()
  size_t max_input_buffer_size() const override { return CHUNK_SIZE * channel_count; }       // input_frames_max()
  size_t pending_copy_to(float* out, size_t out_len) {
    size_t n = std::min(out_len, pending_end - pending_start);
    std::memcpy(out, pending.data() + pending_start, n * sizeof(float));
    return n;
  }
  std::pair<size_t, size_t> process(const float* in, size_t in_len, float* out, size_t out_len) override {
    if (input_rate == output_rate) {  // bypass (rubato.rs:73-78)
      size_t m = std::min(in_len, out_len);
      std::memcpy(out, in, m * sizeof(float));
      return {m, m};
    }
    if (pending_start < pending_end) {  // flush pending outs (rubato.rs:81-86)
      size_t w = pending_copy_to(out, out_len);
      pending_start += w;
      return {0, w};
    }
    if (in_len == 0) return {0, 0};
    for (size_t c = 0; c < channel_count; ++c)
      for (size_t f = 0; f < CHUNK_SIZE; ++f) input[c][f] = in[f * channel_count + c];
    size_t frames = resampler.process_into_buffer(input, output);
    size_t total = channel_count * frames;
    if (total > out_len) {
      pending_start = 0; pending_end = total;
      for (size_t f = 0; f < frames; ++f)
        for (size_t c = 0; c < channel_count; ++c) pending[f * channel_count + c] = output[c][f];
      size_t w = pending_copy_to(out, out_len);
      pending_start += w;
      return {channel_count * CHUNK_SIZE, w};
    }
    for (size_t f = 0; f < frames; ++f)
      for (size_t c = 0; c < channel_count; ++c) out[f * channel_count + c] = output[c][f];
    return {channel_count * CHUNK_SIZE, total};
  }
  bool update(uint32_t in_rate, uint32_t out_rate) override {
    input_rate = in_rate; output_rate = out_rate;
    return resampler.set_resample_ratio((double)out_rate / (double)in_rate, false);
  }
  void reset() override { pending_start = 0; pending_end = 0; }  // rubato.rs:150-153: only the pending buffer
};

// f64 -> u32 `as` cast (saturating, NaN -> 0)
inline uint32_t f64_as_u32(double v) {
  if (std::isnan(v)) return 0;
  if (v <= 0.0) return 0;
  if (v >= 4294967295.0) return 0xFFFFFFFFu;
  return (uint32_t)v;
}
inline size_t f64_as_usize(double v) {
  if (std::isnan(v) || v <= 0.0) return 0;
  if (v >= 18446744073709551615.0) return USIZE_MAX;
  return (size_t)v;
}

// ---- src/utils/dsp/filters/biquad.rs:30-331 ------------------------------------------------------
enum class BiquadType { Lowpass, Highpass, Bandpass, Notch, Peak, Allpass, Bell, Lowshelf, Highshelf };
struct BiquadCoefficients {
  BiquadType filter_type = BiquadType::Lowpass;
  uint32_t sample_rate = 0;
  float cutoff = 0, q = 0, gain = 0;
  double a1 = 0, a2 = 0, a3 = 0, m0 = 0, m1 = 0, m2 = 0;
  bool set(BiquadType t, uint32_t sr, float c, float q_, float g) {
    if (filter_type != t || sample_rate != sr || cutoff != c || q != q_ || gain != g) {
      filter_type = t; sample_rate = sr; cutoff = c; q = q_; gain = g;
      return apply();
    }
    return true;
  }
  bool set_filter_type(BiquadType t) { if (filter_type != t) { filter_type = t; return apply(); } return true; }
  bool set_cutoff(float c) { if (cutoff != c) { cutoff = c; return apply(); } return true; }
  bool apply() {
    if (sample_rate == 0) return false;
    if (q <= 0.0f) return false;
    if (cutoff > (float)sample_rate / 2.0f) return false;
    const double PI = 3.14159265358979323846;
    double g = std::tan(PI * (double)cutoff / (double)sample_rate);
    double k = 1.0 / (double)q;
    double a = 0;
    switch (filter_type) {
      case BiquadType::Bell:
        a = std::pow(10.0, (double)gain / 40.0);
        k = 1.0 / ((double)q * a);
        break;
      case BiquadType::Lowshelf:
        a = std::pow(10.0, (double)gain / 40.0);
        g = g / std::sqrt(a);
        break;
      case BiquadType::Highshelf:
        a = std::pow(10.0, (double)gain / 40.0);
        g = g * std::sqrt(a);
        break;
      default: break;
    }
    a1 = 1.0 / (1.0 + g * (g + k));
    a2 = g * a1;
    a3 = g * a2;
    switch (filter_type) {
      case BiquadType::Lowpass: m0 = 0.0; m1 = 0.0; m2 = 1.0; break;
      case BiquadType::Highpass: m0 = 1.0; m1 = -k; m2 = -1.0; break;
      case BiquadType::Bandpass: m0 = 0.0; m1 = 1.0; m2 = 0.0; break;
      case BiquadType::Notch: m0 = 1.0; m1 = -k; m2 = 0.0; break;
      case BiquadType::Peak: m0 = 1.0; m1 = -k; m2 = -2.0; break;
      case BiquadType::Allpass: m0 = 1.0; m1 = -2.0 * k; m2 = 0.0; break;
      case BiquadType::Bell: m0 = 1.0; m1 = k * (a * a - 1.0); m2 = 0.0; break;
      case BiquadType::Lowshelf: m0 = 1.0; m1 = k * (a - 1.0); m2 = a * a - 1.0; break;
      case BiquadType::Highshelf: m0 = a * a; m1 = k * (1.0 - a) * a; m2 = 1.0 - a * a; break;
    }
    return true;
  }
};
struct BiquadFilter {
  double ic1eq = 0, ic2eq = 0;
  double process_sample(const BiquadCoefficients& c, double v0) {
    double v3 = v0 - ic2eq;
    double v1 = c.a1 * ic1eq + c.a2 * v3;
    double v2 = ic2eq + c.a2 * ic1eq + c.a3 * v3;
    ic1eq = 2.0 * v1 - ic1eq;
    ic2eq = 2.0 * v2 - ic2eq;
    return c.m0 * v0 + c.m1 * v1 + c.m2 * v2;
  }
  void reset() { ic1eq = 0; ic2eq = 0; }
};

// ---- src/parameter/scaling.rs:45-75, src/parameter/float.rs (denormalize_value, clamp_value) -----
struct ParamScaling {
  enum Kind { Linear, Exponential, Decibel, Sigmoid } kind = Linear;
  float a = 0, b = 0;
  float scale(float v) const {
    switch (kind) {
      case Linear: return v;
      case Exponential: return std::pow(v, a);
      case Sigmoid: {
        auto sg = [&](float x) { return 1.0f / (1.0f + std::exp(-a * (x - 0.5f))); };
        float y = sg(v), y0 = sg(0.0f), y1 = sg(1.0f);
        return (y - y0) / (y1 - y0);
      }
      case Decibel: {
        float db = a + v * (b - a);
        float lin = db_to_linear(db);
        float lo = db_to_linear(a), hi = db_to_linear(b);
        return (lin - lo) / (hi - lo);
      }
    }
    return v;
  }
};
struct FloatParam {
  uint32_t id = 0;
  float min = 0, max = 1, def = 0;
  ParamScaling scaling;
  float clamp_value(float v) const { return std::min(std::max(v, min), max); }
  float denormalize(float n) const { return min + scaling.scale(n) * (max - min); }  // float.rs
};
constexpr uint32_t fourcc(const char (&s)[5]) {
  return ((uint32_t)(uint8_t)s[0] << 24) | ((uint32_t)(uint8_t)s[1] << 16) | ((uint32_t)(uint8_t)s[2] << 8) | (uint32_t)(uint8_t)s[3];
}

// ParameterValueUpdate::{Raw(f32), Normalized(f32)} (src/parameter.rs:106-111)
struct ParamUpdate { float value; bool normalized; };

// SmoothedParameterValue<S> (src/parameter/smoothed.rs:17-158)
template <class S>
struct SmoothedParam {
  FloatParam desc;
  S value;
  void from_description(const FloatParam& d) { desc = d; value.init(d.def); }
  void set_sample_rate(uint32_t sr) { value.set_sample_rate(sr); }
  bool need_ramp() const { return value.need_ramp(); }
  float next_value() { return value.next(); }
  float target_value() const { return value.target; }
  float current_value() const { return value.current; }
  void init_value(float v) { value.init(v); }
  void apply_update(const ParamUpdate& u) {
    if (u.normalized) value.set_target(desc.denormalize(std::min(std::max(u.value, 0.0f), 1.0f)));
    else value.set_target(desc.clamp_value(u.value));
  }
};

}  // namespace po
