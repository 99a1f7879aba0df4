// ORACLE -- TEST INFRASTRUCTURE ONLY (see po_dsp.hpp header). CPU restatement of the six effects
// BASELINE.json names: FilterEffect, Eq5Effect, CompressorEffect, ChorusEffect, DelayEffect,
// ReverbEffect, plus the DSP primitives they use (SVF, DC blocker, delay lines, LFO, follower).
// Non-deterministic reference state (OS-seeded RNG: reverb fpd + vibrato phases, LFO random
// shapes) is an explicit input here.
#pragma once
#include "po_graph.hpp"

namespace po {

// ---- src/utils/dsp/filters/svf.rs:28-222 -------------------------------------------------------------
enum class SvfType { Lowpass = 0, Highpass = 1, Bandpass = 2 };
struct SvfCoefficients {
  SvfType filter_type = SvfType::Lowpass;
  uint32_t sample_rate = 0;
  float cutoff = 0, resonance = 0;
  double g = 0, k = 0, a1 = 0, a2 = 0, a3 = 0;
  bool set(SvfType t, uint32_t sr, float c, float r) {
    if (filter_type != t || sample_rate != sr || cutoff != c || resonance != r) {
      filter_type = t; sample_rate = sr; cutoff = c; resonance = r;
      return apply();
    }
    return true;
  }
  bool set_filter_type(SvfType t) { if (filter_type != t) { filter_type = t; return apply(); } return true; }
  bool apply() {
    if (sample_rate == 0) return false;
    if (resonance < 0.0f || resonance > 1.0f) return false;
    if (cutoff > (float)sample_rate / 2.0f) return false;
    const double PI = 3.14159265358979323846;
    g = std::tan(PI * (double)cutoff / (double)sample_rate);
    k = std::max(2.0 * (1.0 - (double)resonance * 0.97), 0.03);
    a1 = 1.0 / (1.0 + g * (g + k));
    a2 = g * a1;
    a3 = g * a2;
    return true;
  }
};
struct SvfFilter {
  double ic1eq = 0, ic2eq = 0;
  double process_sample(const SvfCoefficients& c, double in) {
    double v3 = in - ic2eq;
    double v1 = c.a1 * ic1eq + c.a2 * v3;
    double v2 = ic2eq + c.a2 * ic1eq + c.a3 * v3;
    ic1eq = 2.0 * v1 - ic1eq;
    ic2eq = 2.0 * v2 - ic2eq;
    switch (c.filter_type) {
      case SvfType::Lowpass: return v2;
      case SvfType::Bandpass: return v1;
      case SvfType::Highpass: return in - c.k * v1 - v2;
    }
    return v2;
  }
  void reset() { ic1eq = 0; ic2eq = 0; }
};

// ---- src/utils/dsp/filters/dc.rs:35-88 ----------------------------------------------------------------
struct DcFilter {
  double y1 = 0, x1 = 0, r = 0.999;
  DcFilter() {}
  DcFilter(uint32_t sr, double hz) : r(1.0 - (6.28318530717958647692 * hz / (double)sr)) {}
  void reset() { x1 = 0; y1 = 0; }
  double process_sample(double s) { y1 = s - x1 + r * y1; x1 = s; return y1; }
};

// ---- src/utils/dsp/envelope.rs:5-75 ---------------------------------------------------------------------
struct EnvelopeFollower {
  float current_value = 0, attack_coeff = 0, release_coeff = 0;
  uint32_t sample_rate = 44100;
  EnvelopeFollower() { set_attack_time(0.01f); set_release_time(0.1f); }
  EnvelopeFollower(uint32_t sr, float a, float r) : sample_rate(sr) { set_attack_time(a); set_release_time(r); }
  void set_attack_time(float t) { attack_coeff = t > 0.0f ? std::exp(-1.0f / (t * (float)sample_rate)) : 0.0f; }
  void set_release_time(float t) { release_coeff = t > 0.0f ? std::exp(-1.0f / (t * (float)sample_rate)) : 0.0f; }
  float run(float in) {
    if (in > current_value) current_value = in + attack_coeff * (current_value - in);
    else current_value = in + release_coeff * (current_value - in);
    return current_value;
  }
  void reset(float v) { current_value = v; }
};

// ---- src/utils/dsp/lfo.rs:9-253 (deterministic waveforms only) ---------------------------------------------
inline float sine_approx(float x) {
  const float PI = 3.14159265358979323846f;
  const float B = 4.0f / PI;
  const float C = -4.0f / (PI * PI);
  const float P = 0.225f;
  float y = B * x + C * x * std::fabs(x);
  return P * (y * std::fabs(y) - y) + y;
}
inline float rem_euclid_f32(float a, float b) {
  float r = std::fmod(a, b);
  return r < 0.0f ? r + std::fabs(b) : r;
}
struct Lfo {
  enum Waveform { Sine = 0, Triangle, RampUp, RampDown, Square, Random, SmoothRandom };
  float phase = 0, phase_inc = 0;
  Waveform waveform = Sine;
  Lfo() : Lfo(44100, 1.0, Sine) {}
  Lfo(uint32_t sr, double rate, Waveform w) : phase(0), phase_inc((float)(rate / (double)sr)), waveform(w) {}
  void reset() { phase = 0; }
  void set_rate(uint32_t sr, double rate) { phase_inc = (float)(rate / (double)sr); }
  void set_phase(float p) { phase = rem_euclid_f32(p, 1.0f); }
  void set_phase_degrees(float p) { set_phase(p / 6.28318530717958647692f); }
  float run() {
    const float TAU = 6.28318530717958647692f;
    float v = 0;
    switch (waveform) {
      case Sine: { float p = phase < 0.5f ? phase * TAU : (phase - 1.0f) * TAU; v = sine_approx(p); break; }
      case Triangle: v = phase < 0.25f ? phase * 4.0f : (phase < 0.75f ? 2.0f - phase * 4.0f : phase * 4.0f - 4.0f); break;
      case RampUp: v = phase * 2.0f - 1.0f; break;
      case RampDown: v = 1.0f - phase * 2.0f; break;
      case Square: v = phase < 0.5f ? 1.0f : -1.0f; break;
      default: assert(false && "oracle: OS-seeded random LFO shapes are not reproducible (SURVEY H4)"); break;
    }
    phase += phase_inc;
    if (phase >= 1.0f) phase -= 1.0f;
    return v;
  }
};

inline size_t next_pow2(size_t v) { size_t p = 1; while (p < v) p <<= 1; return p; }
inline size_t f32_ceil_usize(float v) { float c = std::ceil(v); return c <= 0.0f || std::isnan(c) ? 0 : (size_t)c; }

// ---- src/utils/dsp/delay.rs:19-66 (DelayLine<2>) ------------------------------------------------------------
struct DelayLine2 {
  std::vector<double> buffer;  // [frames][2]
  size_t mask = 0, write_pos = 0;
  explicit DelayLine2(size_t max_size = 1) { size_t n = next_pow2(max_size); buffer.assign(n * 2, 0.0); mask = n - 1; }
  void flush() { std::fill(buffer.begin(), buffer.end(), 0.0); write_pos = 0; }
  void process(size_t delay, double& l, double& r) {
    write_pos &= mask;
    buffer[write_pos * 2] = l; buffer[write_pos * 2 + 1] = r;
    write_pos = (write_pos + 1) & mask;
    if (write_pos > delay) write_pos = 0;
    l = buffer[write_pos * 2]; r = buffer[write_pos * 2 + 1];
  }
};

// ---- src/utils/dsp/delay.rs:79-155 (InterpolatedDelayLine<1>) ------------------------------------------------
struct InterpolatedDelayLine1 {
  std::vector<double> buffer;
  size_t mask = 0, write_pos = 0;
  InterpolatedDelayLine1() {}
  explicit InterpolatedDelayLine1(size_t max_size) { size_t n = next_pow2(max_size); buffer.assign(n, 0.0); mask = n - 1; }
  void flush() { std::fill(buffer.begin(), buffer.end(), 0.0); write_pos = 0; }
  float process(float input, float feedback, float delay) {
    double read_pos = (double)write_pos - (double)delay;
    double fl = std::floor(read_pos);
    double fraction = read_pos - fl;
    int64_t index1 = (int64_t)fl;
    int64_t index2 = index1 + 1;
    size_t i1 = (size_t)index1 & mask, i2 = (size_t)index2 & mask;
    double v1 = buffer[i1], v2 = buffer[i2];
    float out = (float)(v1 + (v2 - v1) * fraction);
    buffer[write_pos & mask] = (double)input + (double)out * (double)feedback;
    write_pos = (write_pos + 1) & mask;
    return out;
  }
};

// ---- src/utils/dsp/delay.rs:172-270 (LookupDelayLine<2>) ------------------------------------------------------
struct LookupDelayLine2 {
  std::vector<double> buffer;  // [frames][2]
  size_t write_pos = 0, mask = 0, delay_frames = 0;
  double peak_value = 0;
  size_t peak_pos = 0;
  LookupDelayLine2() {}
  LookupDelayLine2(uint32_t sr, float delay_time) {
    delay_frames = f32_ceil_usize(delay_time * (float)sr);
    if (delay_frames > 0) { size_t n = next_pow2(delay_frames); buffer.assign(n * 2, 0.0); mask = n - 1; }
  }
  void process(const float in[2], float out[2]) {
    if (delay_frames == 0) { out[0] = in[0]; out[1] = in[1]; return; }
    size_t frames = buffer.size() / 2;
    size_t read_index = (write_pos + frames - delay_frames) & mask;
    out[0] = (float)buffer[read_index * 2]; out[1] = (float)buffer[read_index * 2 + 1];
    size_t wi = write_pos & mask;
    buffer[wi * 2] = (double)in[0]; buffer[wi * 2 + 1] = (double)in[1];
    bool peak_expired = peak_pos == read_index;
    double new_peak = std::max(std::max(0.0, (double)std::fabs(in[0])), (double)std::fabs(in[1]));
    if (new_peak >= peak_value) {
      peak_value = new_peak; peak_pos = write_pos;
    } else if (peak_expired) {
      peak_value = 0.0;
      for (size_t i = 0; i < delay_frames; ++i) {
        size_t fi = (write_pos + frames - i) & mask;
        double fp = std::max(std::max(0.0, std::fabs(buffer[fi * 2])), std::fabs(buffer[fi * 2 + 1]));
        if (fp >= peak_value) { peak_value = fp; peak_pos = fi; }
      }
    }
    write_pos = (write_pos + 1) & mask;
  }
};

// ---- src/utils/dsp/delay.rs:283-350 (AllpassDelayLine<2>) -------------------------------------------------------
struct AllpassDelayLine2 {
  std::vector<double> buffer;
  size_t delay = 0, write_pos = 0;
  explicit AllpassDelayLine2(size_t max_size) : buffer(max_size * 2, 0.0) {}
  void flush() { std::fill(buffer.begin(), buffer.end(), 0.0); write_pos = 0; }
  void set_delay(size_t d) { delay = std::min(d, buffer.size() / 2 - 1); }
  void process(double& l, double& r) {
    size_t read_pos = write_pos + 1;
    if (read_pos > delay) read_pos = 0;
    double dl = buffer[read_pos * 2], dr = buffer[read_pos * 2 + 1];
    double bl = l - (dl * 0.5), br = r - (dr * 0.5);
    double ol = bl * 0.5, orr = br * 0.5;
    buffer[write_pos * 2] = bl; buffer[write_pos * 2 + 1] = br;
    write_pos += 1;
    if (write_pos > delay) write_pos = 0;
    ol += buffer[write_pos * 2]; orr += buffer[write_pos * 2 + 1];
    l = ol; r = orr;
  }
};

inline uint32_t enum_from_update(const ParamUpdate& u, uint32_t count) {
  if (u.normalized) {
    float n = std::min(std::max(u.value, 0.0f), 1.0f);
    return (uint32_t)std::round(n * (float)(count - 1));
  }
  float v = std::round(u.value);
  if (v < 0) v = 0;
  if (v > (float)(count - 1)) v = (float)(count - 1);
  return (uint32_t)v;
}
// FloatParameterValue (src/parameter/float.rs): plain, unsmoothed
struct PlainParam {
  FloatParam desc; float value = 0;
  void from_description(const FloatParam& d) { desc = d; value = d.def; }
  void apply_update(const ParamUpdate& u) {
    if (u.normalized) value = desc.denormalize(std::min(std::max(u.value, 0.0f), 1.0f));
    else value = desc.clamp_value(u.value);
  }
};

static const ParamScaling SC_LIN{ParamScaling::Linear, 0, 0};
static const ParamScaling SC_EXP25{ParamScaling::Exponential, 2.5f, 0};
static const ParamScaling SC_EXP2{ParamScaling::Exponential, 2.0f, 0};

// ---- src/effect/filter.rs:48-238 ------------------------------------------------------------------------------
struct FilterEffect : Effect {
  size_t channel_count = 0; uint32_t sample_rate = 0;
  std::vector<BiquadFilter> filters;
  BiquadCoefficients coeffs;
  uint32_t filter_type = 0;  // FilterEffectType
  SmoothedParam<ExpSmoothed> cutoff;
  SmoothedParam<LinearSmoothed> q;
  static BiquadType to_biquad(uint32_t t) {
    switch (t) { case 0: return BiquadType::Lowpass; case 1: return BiquadType::Bandpass; case 2: return BiquadType::Notch; default: return BiquadType::Highpass; }
  }
  FilterEffect() {
    coeffs.set(BiquadType::Lowpass, 44100, 22050.0f, 0.707f, 0.0f);
    cutoff.from_description({fourcc("cuto"), 20.0f, 20000.0f, 20000.0f, SC_EXP25});
    q.from_description({fourcc("fltq"), 0.001f, 4.0f, 0.707f, SC_LIN});
  }
  FilterEffect(uint32_t type, float cut, float q_) : FilterEffect() {  // with_parameters, filter.rs:104-116
    filter_type = type;
    cutoff.init_value(cut);
    q.init_value(q_);
    float c = std::min(std::max(cut, 20.0f), 44100.0f / 2.0f);
    coeffs.set(to_biquad(type), 44100, c, q_, 0.0f);
  }
  const char* name() const override { return "Filter"; }
  size_t weight() const override { return 2; }
  bool initialize(uint32_t sr, size_t ch, size_t) override {
    sample_rate = sr; channel_count = ch;
    float c = std::min(std::max(coeffs.cutoff, 20.0f), (float)sr / 2.0f);
    coeffs.set_cutoff(c);
    filters.assign(ch, BiquadFilter());
    cutoff.set_sample_rate(sr); q.set_sample_rate(sr);
    return true;
  }
  void process(float* buf, size_t len, uint64_t) override {
    if (cutoff.need_ramp() || q.need_ramp()) {
      for (size_t i = 0; i + channel_count <= len; i += channel_count) {
        float c = std::min(std::max(cutoff.next_value(), 20.0f), (float)sample_rate / 2.0f);
        float qq = q.next_value();
        coeffs.set(to_biquad(filter_type), sample_rate, c, qq, 0.0f);
        for (size_t ch = 0; ch < channel_count; ++ch) buf[i + ch] = (float)filters[ch].process_sample(coeffs, (double)buf[i + ch]);
      }
    } else {
      for (size_t ch = 0; ch < channel_count; ++ch)
        for (size_t i = ch; i < len; i += channel_count) buf[i] = (float)filters[ch].process_sample(coeffs, (double)buf[i]);
    }
  }
  bool process_tail(size_t& f) const override { f = (size_t)sample_rate / 10; return true; }
  bool process_parameter_update(uint32_t id, const ParamUpdate& u) override {
    if (id == fourcc("type")) { filter_type = enum_from_update(u, 4); coeffs.set_filter_type(to_biquad(filter_type)); }
    else if (id == fourcc("cuto")) cutoff.apply_update(u);
    else if (id == fourcc("fltq")) q.apply_update(u);
    else return false;
    return true;
  }
};

// ---- src/effect/gain.rs:51-206 ---------------------------------------------------------------------------------
struct GainEffect : Effect {
  SmoothedParam<ExpSmoothed> gain;
  uint32_t dc_filter_mode = 0;  // GainEffectDcFilterMode: Off, Slow, Default, Fast
  std::vector<DcFilter> dc_filters;
  uint32_t sample_rate = 0; size_t channel_count = 0;
  static double mode_hz(uint32_t m) { return m == 1 ? 1.0 : (m == 3 ? 20.0 : 5.0); }  // DcFilterMode::hz; Off -> Default
  GainEffect() { gain.from_description({fourcc("gain"), 0.000001f, 15.848932f, 1.0f, ParamScaling{ParamScaling::Decibel, -60.0f, 24.0f}}); }
  GainEffect(float gain_db, uint32_t dc_mode) : GainEffect() {  // with_parameters (gain.rs:97-104)
    gain.init_value(db_to_linear(std::min(std::max(gain_db, -60.0f), 24.0f)));
    dc_filter_mode = dc_mode;
  }
  const char* name() const override { return "Gain"; }
  size_t weight() const override { return 1; }
  bool initialize(uint32_t sr, size_t ch, size_t) override {
    sample_rate = sr; channel_count = ch;
    gain.set_sample_rate(sr);
    dc_filters.assign(ch, DcFilter(sr, mode_hz(dc_filter_mode)));
    return true;
  }
  void process(float* buf, size_t len, uint64_t) override {
    if (dc_filter_mode != 0)
      for (size_t ch = 0; ch < channel_count; ++ch)
        for (size_t i = ch; i < len; i += channel_count) buf[i] = (float)dc_filters[ch].process_sample((double)buf[i]);
    if (gain.need_ramp()) {
      for (size_t i = 0; i + channel_count <= len; i += channel_count) {
        float g = gain.next_value();
        for (size_t ch = 0; ch < channel_count; ++ch) buf[i + ch] *= g;
      }
    } else {
      scale_buffer(buf, len, gain.target_value());
    }
  }
  bool process_tail(size_t& f) const override {
    f = dc_filter_mode != 0 ? (size_t)sample_rate / (size_t)mode_hz(dc_filter_mode) : 0;
    return true;
  }
  bool process_parameter_update(uint32_t id, const ParamUpdate& u) override {
    if (id == fourcc("gain")) gain.apply_update(u);
    else if (id == fourcc("dcfm")) {
      dc_filter_mode = enum_from_update(u, 4);
      if (dc_filter_mode != 0) for (auto& f : dc_filters) f.r = 1.0 - (6.28318530717958647692 * mode_hz(dc_filter_mode) / (double)sample_rate);
      else for (auto& f : dc_filters) f.reset();
    } else return false;
    return true;
  }
};

// ---- src/effect/pan.rs:17-192 ----------------------------------------------------------------------------------
struct PanningEffect : Effect {
  size_t channel_count = 0;
  SmoothedParam<ExpSmoothed> pan, width;
  bool invert_l = false, invert_r = false;
  PanningEffect() {
    pan.from_description({fourcc("pan "), -1.0f, 1.0f, 0.0f, SC_LIN});
    width.from_description({fourcc("wdth"), 0.0f, 2.0f, 1.0f, SC_LIN});
  }
  const char* name() const override { return "Panning"; }
  size_t weight() const override { return 1; }
  bool initialize(uint32_t sr, size_t ch, size_t) override {
    if (ch != 2) return false;
    channel_count = ch;
    pan.set_sample_rate(sr); width.set_sample_rate(sr);
    return true;
  }
  void process(float* buf, size_t len, uint64_t) override {
    const float il = invert_l ? -1.0f : 1.0f, ir = invert_r ? -1.0f : 1.0f;
    const bool has_invert = il < 0.0f || ir < 0.0f;
    const bool pan_ramping = pan.need_ramp(), width_ramping = width.need_ramp();
    if (!has_invert && !pan_ramping && !width_ramping && std::fabs(pan.target_value()) < 1e-6f &&
        std::fabs(width.target_value() - 1.0f) < 1e-6f)
      return;
    for (size_t i = 0; i + 2 <= len; i += 2) {
      float l = buf[i] * il, r = buf[i + 1] * ir;
      float w = width_ramping ? width.next_value() : width.target_value();
      if (std::fabs(w - 1.0f) > 1e-6f) {
        float mid = (l + r) * 0.5f, side = (l - r) * 0.5f;
        l = mid + side * w;
        r = mid - side * w;
      }
      float p = pan_ramping ? pan.next_value() : pan.target_value();
      if (std::fabs(p) > 1e-6f) {
        float pl, pr;
        panning_factors(p, pl, pr);
        l *= pl; r *= pr;
      }
      buf[i] = l; buf[i + 1] = r;
    }
  }
  bool process_tail(size_t& f) const override { f = 0; return true; }
  bool process_parameter_update(uint32_t id, const ParamUpdate& u) override {
    auto as_bool = [](const ParamUpdate& x) { return x.normalized ? std::min(std::max(x.value, 0.0f), 1.0f) >= 0.5f : x.value != 0.0f; };
    if (id == fourcc("pan ")) pan.apply_update(u);
    else if (id == fourcc("wdth")) width.apply_update(u);
    else if (id == fourcc("invl")) invert_l = as_bool(u);
    else if (id == fourcc("invr")) invert_r = as_bool(u);
    else return false;
    return true;
  }
};

// ---- src/effect/gate.rs:12-224 ---------------------------------------------------------------------------------
struct GateEffect : Effect {
  PlainParam threshold, attack_time, hold_time, release_time, range;
  EnvelopeFollower envelope_follower;
  uint32_t hold_counter = 0;
  float gate_gain_db = -60.0f, attack_coeff = 0.0f, release_coeff = 0.0f;
  uint32_t sample_rate = 0; size_t channel_count = 0;
  GateEffect() {
    threshold.from_description({fourcc("thrs"), -60.0f, 0.0f, -30.0f, SC_LIN});
    attack_time.from_description({fourcc("attk"), 0.001f, 0.5f, 0.005f, SC_LIN});
    hold_time.from_description({fourcc("hold"), 0.0f, 2.0f, 0.1f, SC_LIN});
    release_time.from_description({fourcc("rels"), 0.01f, 2.0f, 0.2f, SC_LIN});
    range.from_description({fourcc("rnge"), -60.0f, 0.0f, -60.0f, SC_LIN});
  }
  GateEffect(float thr, float atk, float hold, float rel, float rng) : GateEffect() {
    threshold.value = thr; attack_time.value = atk; hold_time.value = hold; release_time.value = rel; range.value = rng;
  }
  void update_coefficients() {  // gate.rs:83-95
    if (sample_rate > 0) {
      envelope_follower.set_attack_time(attack_time.value);
      envelope_follower.set_release_time(release_time.value);
      float sr = (float)sample_rate;
      attack_coeff = std::exp(-1.0f / (attack_time.value * sr));
      release_coeff = std::exp(-1.0f / (release_time.value * sr));
    }
  }
  const char* name() const override { return "Gate"; }
  size_t weight() const override { return 2; }
  bool initialize(uint32_t sr, size_t ch, size_t) override {
    if (ch != 2) return false;
    sample_rate = sr; channel_count = ch;
    envelope_follower = EnvelopeFollower(sr, attack_time.value, release_time.value);
    envelope_follower.reset(-120.0f);
    hold_counter = 0;
    gate_gain_db = range.value;
    update_coefficients();
    return true;
  }
  void process(float* buf, size_t len, uint64_t) override {
    const float thr = threshold.value, range_db = range.value;
    const uint32_t hold_samples = f64_as_u32((double)(hold_time.value * (float)sample_rate));
    for (size_t i = 0; i + 2 <= len; i += 2) {
      float frame_peak = std::max(std::fabs(buf[i]), std::fabs(buf[i + 1]));
      float input_db = frame_peak > 1e-6f ? 20.0f * std::log10(frame_peak) : -120.0f;
      float envelope = envelope_follower.run(input_db);
      float target_gain_db;
      if (envelope >= thr) { hold_counter = hold_samples; target_gain_db = 0.0f; }
      else if (hold_counter > 0) { hold_counter -= 1; target_gain_db = 0.0f; }
      else target_gain_db = range_db;
      if (target_gain_db > gate_gain_db) gate_gain_db = attack_coeff * gate_gain_db + (1.0f - attack_coeff) * target_gain_db;
      else gate_gain_db = release_coeff * gate_gain_db + (1.0f - release_coeff) * target_gain_db;
      float gain = gate_gain_db <= -60.0f ? 0.0f : db_to_linear(gate_gain_db);
      buf[i] *= gain; buf[i + 1] *= gain;
    }
  }
  bool process_tail(size_t& f) const override {
    f = (size_t)std::ceil(hold_time.value * (float)sample_rate) + (size_t)std::ceil(release_time.value * (float)sample_rate);
    return true;
  }
  bool process_parameter_update(uint32_t id, const ParamUpdate& u) override {
    if (id == fourcc("thrs")) threshold.apply_update(u);
    else if (id == fourcc("attk")) attack_time.apply_update(u);
    else if (id == fourcc("hold")) hold_time.apply_update(u);
    else if (id == fourcc("rels")) release_time.apply_update(u);
    else if (id == fourcc("rnge")) range.apply_update(u);
    else return false;
    update_coefficients();
    return true;
  }
};

// ---- src/effect/distortion.rs:19-386 ---------------------------------------------------------------------------
// Waveshapers (distortion.rs:124-190). `powi(2)` / `powi(3)` are multiplication chains.
inline float dist_shape(uint32_t type, float sample, float drive) {
  const float MAX_DRIVE = 4.0f;
  const float t = drive / MAX_DRIVE;
  switch (type) {
    case 0: {  // SoftClip
      const float gain = 1.0f + (t * t) * (15.0f - 1.0f);
      const float x = sample * gain;
      if (x >= 1.0f) return 1.0f;
      if (x > -1.0f) return gain <= 1.0f ? sample : (3.0f / 2.0f) * (x - ((x * x) * x) / 3.0f);
      return -1.0f;
    }
    case 1: {  // HardClip: sample.clamp(-threshold, threshold) * gain
      const float gain = 1.0f + (t * t) * (25.0f - 1.0f);
      const float threshold = 1.0f / gain;
      float c = sample;
      if (c < -threshold) c = -threshold;
      if (c > threshold) c = threshold;
      return c * gain;
    }
    case 2: {  // Diode
      const float curve = 0.6f * (t * t) + 0.4f * t;
      const float gain = 1.0f + curve * (20.0f - 1.0f);
      const float diode_clipping = std::exp((0.1f * sample) / (0.0253f * 1.68f)) - 1.0f;
      return 2.0f / 3.14159265358979323846f * std::atan(diode_clipping * gain);
    }
    case 3: {  // Fuzz
      const float gain = 1.0f + (1.0f - std::exp(-3.0f * t)) * (30.0f - 1.0f);
      const float amplified = sample * gain;
      const float saturated = amplified < 0.0f ? -1.0f * (1.0f - std::exp(-std::fabs(amplified))) : 1.0f * (1.0f - std::exp(-std::fabs(amplified)));
      return 1.5f * (saturated + std::fabs(saturated));
    }
    default: {  // Fold
      const float gain = 1.0f + (t * t) * (4.0f - 1.0f);
      const float x = sample * gain;
      const float threshold = 1.0f / gain;
      if (x > threshold || x < -threshold)
        return std::fabs(std::fmod(std::fabs(x - threshold), threshold * 4.0f) - threshold * 2.0f) - threshold;
      return x;
    }
  }
}
// DistortionType::rms_compensation (distortion.rs:84-122)
inline float dist_rms_compensation(uint32_t type, float drive) {
  const int N = 256;
  const float PARTIALS[5][2] = {{1.0f, 0.60f}, {2.7f, 0.25f}, {5.3f, 0.10f}, {9.1f, 0.03f}, {14.6f, 0.02f}};
  float partials_peak = 0.0f;
  for (auto& p : PARTIALS) partials_peak += p[1];
  float input_sum_sq = 0.0f, output_sum_sq = 0.0f;
  for (int i = 0; i < N; ++i) {
    const float t = 6.28318530717958647692f * ((float)i + 0.5f) / (float)N;
    float sum = 0.0f;
    for (auto& p : PARTIALS) sum += p[1] * std::sin(p[0] * t);
    const float sample = sum / partials_peak;
    input_sum_sq += sample * sample;
    const float o = dist_shape(type, sample, drive);
    output_sum_sq += o * o;
  }
  const float input_rms = std::sqrt(input_sum_sq / (float)N), output_rms = std::sqrt(output_sum_sq / (float)N);
  return output_rms > 1e-10f ? input_rms / output_rms : 1.0f;
}
struct DistortionEffect : Effect {
  uint32_t distortion_type = 2;  // DistortionType::Diode (TYPE default)
  SmoothedParam<LinearSmoothed> drive;
  SmoothedParam<ExpSmoothed> mix;
  float luts[5][256];
  size_t channel_count = 0;
  DistortionEffect() {
    drive.from_description({fourcc("driv"), 0.0f, 4.0f, 0.0f, SC_LIN});
    drive.value.step = 0.01f;
    mix.from_description({fourcc("mix "), 0.0f, 1.0f, 1.0f, SC_LIN});
    mix.value.inertia = 0.1f;
    for (uint32_t ty = 0; ty < 5; ++ty)
      for (int i = 0; i < 256; ++i) luts[ty][i] = dist_rms_compensation(ty, (float)i / 255.0f * 4.0f);
  }
  DistortionEffect(uint32_t type, float drive_, float mix_) : DistortionEffect() {  // with_parameters (distortion.rs:247-253)
    distortion_type = type;
    drive.init_value(drive_);
    mix.init_value(mix_);
  }
  float lookup(float d) const {  // lookup_gain_compensation (distortion.rs:271-279)
    const float* lut = luts[distortion_type];
    float pos = std::min(std::max(d / 4.0f, 0.0f), 1.0f) * 255.0f;
    size_t lo = (size_t)pos;
    size_t hi = std::min(lo + 1, (size_t)255);
    float frac = pos - (float)lo;
    return lut[lo] + (lut[hi] - lut[lo]) * frac;
  }
  const char* name() const override { return "Distortion"; }
  size_t weight() const override { return 1; }
  bool initialize(uint32_t sr, size_t ch, size_t) override {
    channel_count = ch;
    mix.set_sample_rate(sr); drive.set_sample_rate(sr);
    return true;
  }
  void process(float* buf, size_t len, uint64_t) override {
    if (!mix.need_ramp() && mix.target_value() == 0.0f) {
    } else if (!mix.need_ramp() && mix.target_value() >= 1.0f) {
      if (!drive.need_ramp()) {
        const float d = drive.target_value();
        const float comp = lookup(d);
        for (size_t i = 0; i < len; ++i) buf[i] = dist_shape(distortion_type, buf[i], d) * comp;
      } else {
        for (size_t i = 0; i + channel_count <= len; i += channel_count) {
          const float d = drive.next_value();
          const float comp = lookup(d);
          for (size_t ch = 0; ch < channel_count; ++ch) buf[i + ch] = dist_shape(distortion_type, buf[i + ch], d) * comp;
        }
      }
    } else {
      for (size_t i = 0; i + channel_count <= len; i += channel_count) {
        const float d = drive.next_value();
        const float comp = lookup(d);
        const float m = mix.next_value();
        for (size_t ch = 0; ch < channel_count; ++ch) {
          const float dry = buf[i + ch];
          const float wet = dist_shape(distortion_type, dry, d) * comp;
          buf[i + ch] = (1.0f - m) * dry + m * wet;
        }
      }
    }
  }
  bool process_tail(size_t& f) const override { f = 0; return true; }
  bool process_parameter_update(uint32_t id, const ParamUpdate& u) override {
    if (id == fourcc("type")) distortion_type = enum_from_update(u, 5);
    else if (id == fourcc("driv")) drive.apply_update(u);
    else if (id == fourcc("mix ")) mix.apply_update(u);
    else return false;
    return true;
  }
};

// ---- src/effect/eq5.rs:19-364 ------------------------------------------------------------------------------------
struct Eq5Effect : Effect {
  uint32_t sample_rate = 0; size_t channel_count = 0;
  SmoothedParam<ExpSmoothed> gains[5], frequencies[5];
  SmoothedParam<LinearSmoothed> bandwidths[5];
  BiquadCoefficients coeffs[5];
  std::vector<std::array<BiquadFilter, 5>> filters;
  Eq5Effect() {
    static const char* gid[5] = {"gan1", "gan2", "gan3", "gan4", "gan5"};
    static const char* fid[5] = {"frq1", "frq2", "frq3", "frq4", "frq5"};
    static const char* bid[5] = {"bw_1", "bw_2", "bw_3", "bw_4", "bw_5"};
    static const float fdef[5] = {100.0f, 1000.0f, 4000.0f, 8000.0f, 12000.0f};
    static const float bmax[5] = {1.0f, 4.0f, 4.0f, 4.0f, 1.0f};
    auto cc = [](const char* s) { return ((uint32_t)(uint8_t)s[0] << 24) | ((uint32_t)(uint8_t)s[1] << 16) | ((uint32_t)(uint8_t)s[2] << 8) | (uint32_t)(uint8_t)s[3]; };
    for (int i = 0; i < 5; ++i) {
      gains[i].from_description({cc(gid[i]), -20.0f, 20.0f, 0.0f, SC_LIN});
      frequencies[i].from_description({cc(fid[i]), 20.0f, 20000.0f, fdef[i], SC_EXP25});
      bandwidths[i].from_description({cc(bid[i]), 0.0001f, bmax[i], bmax[i], SC_LIN});
    }
  }
  static BiquadType band_type(int i) { return i == 0 ? BiquadType::Lowshelf : (i == 4 ? BiquadType::Highshelf : BiquadType::Bell); }
  bool update_filter_coefficients() {  // eq5.rs:173-188
    for (int i = 0; i < 5; ++i) {
      float c = std::min(std::max(frequencies[i].current_value(), 20.0f), (float)sample_rate / 2.0f);
      if (!coeffs[i].set(band_type(i), sample_rate, c, bandwidths[i].current_value(), gains[i].current_value())) return false;
    }
    return true;
  }
  void ramp_filter_coefficients() {  // eq5.rs:191-209
    for (int i = 0; i < 5; ++i) {
      float qv = (i == 0 || i == 4) ? bandwidths[i].next_value() : 1.0f / std::max(bandwidths[i].next_value(), 0.001f);
      float c = std::min(std::max(frequencies[i].next_value(), 20.0f), (float)sample_rate / 2.0f);
      float g = gains[i].next_value();
      coeffs[i].set(band_type(i), sample_rate, c, qv, g);
    }
  }
  const char* name() const override { return "Eq5"; }
  size_t weight() const override { return 3; }
  bool initialize(uint32_t sr, size_t ch, size_t) override {
    sample_rate = sr; channel_count = ch;
    for (int i = 0; i < 5; ++i) { gains[i].set_sample_rate(sr); frequencies[i].set_sample_rate(sr); bandwidths[i].set_sample_rate(sr); }
    if (!update_filter_coefficients()) return false;
    filters.assign(ch, {});
    for (int i = 0; i < 5; ++i) {
      gains[i].init_value(gains[i].target_value());
      frequencies[i].init_value(frequencies[i].target_value());
      bandwidths[i].init_value(bandwidths[i].target_value());
    }
    return true;
  }
  void process(float* out, size_t len, uint64_t) override {
    bool need_ramp = false;
    for (int i = 0; i < 5; ++i) need_ramp |= frequencies[i].need_ramp();
    for (int i = 0; i < 5; ++i) need_ramp |= bandwidths[i].need_ramp();
    for (int i = 0; i < 5; ++i) need_ramp |= gains[i].need_ramp();
    size_t frames = len / channel_count;
    for (size_t f = 0; f < frames; ++f) {
      if (need_ramp) ramp_filter_coefficients();
      for (size_t ch = 0; ch < channel_count; ++ch) {
        float s = out[f * channel_count + ch];
        for (int i = 0; i < 5; ++i) s = (float)filters[ch][i].process_sample(coeffs[i], (double)s);
        out[f * channel_count + ch] = s;
      }
    }
  }
  bool process_tail(size_t& f) const override { f = (size_t)sample_rate / 5; return true; }
  bool process_parameter_update(uint32_t id, const ParamUpdate& u) override {
    bool found = false;
    for (int i = 0; i < 5 && !found; ++i) {
      if (id == gains[i].desc.id) { gains[i].apply_update(u); found = true; }
      else if (id == frequencies[i].desc.id) { frequencies[i].apply_update(u); found = true; }
      else if (id == bandwidths[i].desc.id) { bandwidths[i].apply_update(u); found = true; }
    }
    if (!found) return false;
    return update_filter_coefficients();
  }
};

// ---- src/effect/compressor.rs:24-331 --------------------------------------------------------------------------------
struct CompressorEffect : Effect {
  uint32_t sample_rate = 0; size_t channel_count = 0;
  PlainParam threshold, ratio, knee_width, attack_time, release_time, lookahead_time;
  SmoothedParam<ExpSmoothed> makeup_gain;
  EnvelopeFollower follower;
  LookupDelayLine2 delay_line;
  CompressorEffect() {
    threshold.from_description({fourcc("thrs"), -60.0f, 0.0f, -12.0f, SC_LIN});
    ratio.from_description({fourcc("rato"), 1.0f, 20.0f, 8.0f, SC_LIN});
    knee_width.from_description({fourcc("knee"), 0.0f, 12.0f, 3.0f, SC_LIN});
    attack_time.from_description({fourcc("attk"), 0.001f, 0.5f, 0.02f, SC_LIN});
    release_time.from_description({fourcc("rels"), 0.1f, 2.0f, 2.0f, SC_LIN});
    makeup_gain.from_description({fourcc("gain"), -24.0f, 24.0f, 6.0f, SC_LIN});
    lookahead_time.from_description({fourcc("look"), 0.001f, 0.2f, 0.04f, SC_LIN});
  }
  // with_compressor_parameters, compressor.rs:122-140 (set_value asserts range; values are used as is)
  CompressorEffect(float thr, float rat, float knee, float atk, float rel, float makeup, float look) : CompressorEffect() {
    threshold.value = thr; ratio.value = rat; knee_width.value = knee; attack_time.value = atk;
    release_time.value = rel; makeup_gain.init_value(makeup); lookahead_time.value = look;
  }
  const char* name() const override { return "Compressor"; }
  size_t weight() const override { return 4; }
  bool initialize(uint32_t sr, size_t ch, size_t) override {
    sample_rate = sr; channel_count = ch;
    if (ch != 2) return false;
    makeup_gain.set_sample_rate(sr);
    delay_line = LookupDelayLine2(sr, lookahead_time.value);
    follower = EnvelopeFollower(sr, attack_time.value, release_time.value);
    follower.reset(ratio.value >= 20.0f ? -120.0f : 0.0f);
    return true;
  }
  void process(float* out, size_t len, uint64_t) override {
    for (size_t i = 0; i + 2 <= len; i += 2) {
      float in[2] = {out[i], out[i + 1]};
      float delayed[2];
      delay_line.process(in, delayed);
      float input_db;
      if (ratio.value >= 20.0f) {
        float peak = (float)delay_line.peak_value;
        input_db = peak > 1e-6f ? 20.0f * std::log10(peak) : -120.0f;
      } else {
        float peak = std::max(std::fabs(in[0]), std::fabs(in[1]));
        input_db = peak > 1e-6f ? 20.0f * std::log10(peak) : -120.0f;
      }
      float envelope = follower.run(input_db);
      float t = threshold.value, w = knee_width.value;
      float slope = ratio.value >= 20.0f ? 1.0f : 1.0f - 1.0f / ratio.value;
      float gr_db;
      if (w > 0.0f && envelope > (t - w / 2.0f) && envelope < (t + w / 2.0f)) {
        float knee_lower = t - w / 2.0f;
        float x = (envelope - knee_lower) / w;
        gr_db = x * x * slope * w / 2.0f;
      } else if (envelope > (t + w / 2.0f)) {
        gr_db = (envelope - t) * slope;
      } else {
        gr_db = 0.0f;
      }
      float makeup = makeup_gain.next_value();
      float total_gain = db_to_linear(makeup - gr_db);
      out[i] = delayed[0] * total_gain;
      out[i + 1] = delayed[1] * total_gain;
    }
  }
  bool process_tail(size_t& f) const override {
    f = f32_ceil_usize(lookahead_time.value * (float)sample_rate) + f32_ceil_usize(release_time.value * (float)sample_rate);
    return true;
  }
  bool process_parameter_update(uint32_t id, const ParamUpdate& u) override {
    float old_look = lookahead_time.value;
    if (id == fourcc("thrs")) threshold.apply_update(u);
    else if (id == fourcc("rato")) ratio.apply_update(u);
    else if (id == fourcc("knee")) knee_width.apply_update(u);
    else if (id == fourcc("attk")) attack_time.apply_update(u);
    else if (id == fourcc("rels")) release_time.apply_update(u);
    else if (id == fourcc("gain")) makeup_gain.apply_update(u);
    else if (id == fourcc("look")) lookahead_time.apply_update(u);
    else return false;
    if (sample_rate > 0) { follower.set_attack_time(attack_time.value); follower.set_release_time(release_time.value); }
    if (lookahead_time.value != old_look && sample_rate > 0) delay_line = LookupDelayLine2(sample_rate, lookahead_time.value);
    return true;
  }
};

// ---- src/effect/chorus.rs:48-460 --------------------------------------------------------------------------------------
struct ChorusEffect : Effect {
  uint32_t sample_rate = 0; size_t channel_count = 0;
  SmoothedParam<LinearSmoothed> rate, phase;
  SmoothedParam<ExpSmoothed> depth, feedback, wet_mix, filter_freq, filter_resonance;
  SmoothedParam<SpringSmoothed> delay;
  uint32_t filter_type = 0;  // ChorusEffectFilterType: Lowpass, Highpass, Bandpass == SvfType order
  float lfo_range = 0; double current_phase = 0;
  Lfo left_osc, right_osc;
  InterpolatedDelayLine1 delay_left, delay_right;
  SvfCoefficients filter_coeffs;
  SvfFilter filter_left, filter_right;
  static constexpr float PI_F = 3.14159265358979323846f;
  ChorusEffect() {
    rate.value = LinearSmoothed(0, 0.005f);
    rate.from_description({fourcc("rate"), 0.01f, 10.0f, 1.0f, SC_EXP2});
    phase.value = LinearSmoothed(0, 0.001f);
    phase.from_description({fourcc("phas"), 0.0f, PI_F, PI_F / 2.0f, SC_LIN});
    depth.from_description({fourcc("dpth"), 0.0f, 1.0f, 0.25f, SC_LIN});
    feedback.from_description({fourcc("fdbk"), -1.0f, 1.0f, 0.5f, SC_LIN});
    delay.value = SpringSmoothed(0, 1000);
    delay.from_description({fourcc("dlay"), 0.0f, 100.0f, 12.0f, SC_LIN});
    wet_mix.from_description({fourcc("wet_"), 0.0f, 1.0f, 0.5f, SC_LIN});
    filter_freq.from_description({fourcc("fltf"), 20.0f, 20000.0f, 20000.0f, SC_EXP25});
    filter_resonance.from_description({fourcc("fltq"), 0.0f, 1.0f, 0.0f, SC_LIN});
  }
  ChorusEffect(float r, float p, float d, float fb, float dl, float wet, uint32_t ft, float ff, float fq) : ChorusEffect() {
    rate.init_value(r); phase.init_value(p); depth.init_value(d); feedback.init_value(fb); delay.init_value(dl);
    wet_mix.init_value(wet); filter_type = ft; filter_freq.init_value(ff); filter_resonance.init_value(fq);
  }
  void reset_lfos() {
    double r = (double)rate.current_value();
    left_osc = Lfo(sample_rate, r, Lfo::Sine);
    right_osc = Lfo(sample_rate, r, Lfo::Sine);
    double off = (double)phase.current_value();
    left_osc.set_phase_degrees((float)current_phase);
    right_osc.set_phase_degrees((float)(current_phase + off));
  }
  void update_lfos() {
    double r = (double)rate.next_value();
    left_osc.set_rate(sample_rate, r);
    right_osc.set_rate(sample_rate, r);
    double off = (double)phase.next_value();
    left_osc.set_phase_degrees((float)current_phase);
    right_osc.set_phase_degrees((float)(current_phase + off));
  }
  void reset() {
    delay_left.flush(); delay_right.flush(); filter_left.reset(); filter_right.reset();
    rate.init_value(rate.target_value()); phase.init_value(phase.target_value());
    current_phase = 0.0;
    reset_lfos();
  }
  const char* name() const override { return "Chorus"; }
  size_t weight() const override { return 3; }
  bool initialize(uint32_t sr, size_t ch, size_t) override {
    sample_rate = sr; channel_count = ch;
    if (ch != 2) return false;
    rate.set_sample_rate(sr); phase.set_sample_rate(sr); depth.set_sample_rate(sr); feedback.set_sample_rate(sr);
    delay.set_sample_rate(sr); wet_mix.set_sample_rate(sr); filter_freq.set_sample_rate(sr); filter_resonance.set_sample_rate(sr);
    lfo_range = 256.0f * ((float)sr / 44100.0f);
    size_t max_depth = f32_ceil_usize(lfo_range);
    size_t max_delay = f32_ceil_usize(100.0f * (float)sr / 1000.0f);
    size_t max_buffer = 2 + max_delay + 2 * max_depth + 1;
    delay_left = InterpolatedDelayLine1(max_buffer);
    delay_right = InterpolatedDelayLine1(max_buffer);
    float c = std::min(std::max(filter_freq.target_value(), 20.0f), (float)sr / 2.0f);
    filter_coeffs = SvfCoefficients();
    if (!filter_coeffs.set((SvfType)filter_type, sr, c, filter_resonance.target_value())) return false;
    reset();
    return true;
  }
  void process(float* out, size_t len, uint64_t) override {
    for (size_t i = 0; i + 2 <= len; i += 2) {
      float li = out[i], ri = out[i + 1];
      float delay_ms = delay.next_value();
      float dp = depth.next_value();
      float fb = std::min(std::max(feedback.next_value(), -0.999f), 0.999f);
      float wet = wet_mix.next_value();
      float dry = 1.0f - wet;
      if (rate.need_ramp() || phase.need_ramp()) update_lfos();
      if (filter_freq.need_ramp() || filter_resonance.need_ramp()) {
        float c = std::min(std::max(filter_freq.next_value(), 20.0f), (float)sample_rate / 2.0f);
        float r = filter_resonance.next_value();
        filter_coeffs.set((SvfType)filter_type, sample_rate, c, r);
      }
      double fl = filter_left.process_sample(filter_coeffs, (double)li);
      double fr = filter_right.process_sample(filter_coeffs, (double)ri);
      float delay_in_samples = delay_ms * (float)sample_rate * 0.001f;
      float depth_in_samples = lfo_range * dp;
      float llfo = left_osc.run(), rlfo = right_osc.run();
      float lpos = 2.0f + delay_in_samples + (1.0f + llfo) * depth_in_samples;
      float rpos = 2.0f + delay_in_samples + (1.0f + rlfo) * depth_in_samples;
      float lo = delay_left.process((float)fl, fb, lpos);
      float ro = delay_right.process((float)fr, fb, rpos);
      out[i] = li * dry + lo * wet;
      out[i + 1] = ri * dry + ro * wet;
    }
    const double PI = 3.14159265358979323846;
    double phase_inc = 2.0 * PI * (double)rate.current_value() / (double)sample_rate;
    current_phase += (double)len / (double)channel_count * phase_inc;
    while (current_phase >= 2.0 * PI) current_phase -= 2.0 * PI;
  }
  bool process_tail(size_t& f) const override {
    float delay_ms = delay.target_value();
    float depth_ms = 256.0f * 1000.0f / (float)sample_rate;
    float total_ms = delay_ms + depth_ms;
    float fb = std::fabs(feedback.target_value());
    if (fb >= 1.0f) { f = USIZE_MAX; return true; }
    if (fb < 0.001f) { f = f32_ceil_usize(total_ms * (float)sample_rate / 1000.0f); return true; }
    float total_samples = total_ms * (float)sample_rate / 1000.0f;
    float decay = total_samples + (float)((double)total_samples * std::log10(0.001) / std::log10((double)fb));
    f = f32_ceil_usize(decay);
    return true;
  }
  bool process_parameter_update(uint32_t id, const ParamUpdate& u) override {
    if (id == fourcc("rate")) rate.apply_update(u);
    else if (id == fourcc("phas")) phase.apply_update(u);
    else if (id == fourcc("dpth")) depth.apply_update(u);
    else if (id == fourcc("fdbk")) feedback.apply_update(u);
    else if (id == fourcc("dlay")) delay.apply_update(u);
    else if (id == fourcc("wet_")) wet_mix.apply_update(u);
    else if (id == fourcc("fltt")) { filter_type = enum_from_update(u, 3); filter_coeffs.set_filter_type((SvfType)filter_type); }
    else if (id == fourcc("fltf")) filter_freq.apply_update(u);
    else if (id == fourcc("fltq")) filter_resonance.apply_update(u);
    else return false;
    return true;
  }
};

// ---- src/effect/delay.rs:70-521 ----------------------------------------------------------------------------------------
struct DelayEffect : Effect {
  uint32_t sample_rate = 0;
  uint32_t mode = 0;  // 0 Stereo, 1 PingPong
  SmoothedParam<SpringSmoothed> delay_time;
  SmoothedParam<ExpSmoothed> feedback, filter_cutoff, drive, wet_mix, stereo_width, lfo_rate, lfo_depth_time, lfo_depth_feedback, lfo_depth_filter;
  uint32_t filter_type = 0, lfo_shape = 0;
  InterpolatedDelayLine1 delay_left, delay_right;
  Lfo lfo;
  SvfCoefficients filter_coeffs;
  SvfFilter filter_left, filter_right;
  DcFilter dc_left, dc_right;
  float feedback_left = 0, feedback_right = 0;
  static constexpr float MAX_DELAY_MS = 4000.0f, MAX_LFO_TIME_MOD_MS = 50.0f, FILTER_RESONANCE = 0.302f;
  DelayEffect() {
    delay_time.value = SpringSmoothed(0, 20000);
    delay_time.from_description({fourcc("dlay"), 1.0f, MAX_DELAY_MS, 375.0f, SC_LIN});
    feedback.from_description({fourcc("fdbk"), 0.0f, 1.0f, 0.5f, SC_LIN});
    filter_cutoff.from_description({fourcc("cuto"), 20.0f, 20000.0f, 6000.0f, SC_EXP25});
    drive.from_description({fourcc("driv"), 0.0f, 1.0f, 0.0f, SC_LIN});
    wet_mix.from_description({fourcc("wet_"), 0.0f, 1.0f, 0.5f, SC_LIN});
    stereo_width.from_description({fourcc("wdth"), 0.0f, 1.0f, 0.5f, SC_LIN});
    lfo_rate.from_description({fourcc("lfor"), 0.01f, 10.0f, 1.0f, SC_EXP2});
    lfo_depth_time.from_description({fourcc("lfdt"), -1.0f, 1.0f, 0.0f, SC_LIN});
    lfo_depth_feedback.from_description({fourcc("ldfb"), -1.0f, 1.0f, 0.0f, SC_LIN});
    lfo_depth_filter.from_description({fourcc("lfdf"), -1.0f, 1.0f, 0.0f, SC_LIN});
  }
  static double saturate(double input, float drv) {
    if (drv < 0.001f) return input;
    double gain = 1.0 + (double)drv * 4.0;
    double x = input * gain;
    double x2 = x * x;
    double o = x * (27.0 + x2) / (27.0 + 9.0 * x2);
    return o / std::sqrt(gain);
  }
  static float process_feedback(SvfFilter& f, const SvfCoefficients& c, DcFilter& dc, float delayed, float drv) {
    double filtered = f.process_sample(c, (double)delayed);
    double sat = saturate(filtered, drv);
    float clean = (float)dc.process_sample(sat);
    return std::min(std::max(clean, -4.0f), 4.0f);
  }
  const char* name() const override { return "Delay"; }
  size_t weight() const override { return 3; }
  bool initialize(uint32_t sr, size_t ch, size_t) override {
    sample_rate = sr;
    if (ch != 2) return false;
    delay_time.set_sample_rate(sr); feedback.set_sample_rate(sr); filter_cutoff.set_sample_rate(sr); drive.set_sample_rate(sr);
    wet_mix.set_sample_rate(sr); stereo_width.set_sample_rate(sr); lfo_rate.set_sample_rate(sr);
    lfo_depth_time.set_sample_rate(sr); lfo_depth_feedback.set_sample_rate(sr); lfo_depth_filter.set_sample_rate(sr);
    size_t max_delay = f32_ceil_usize((MAX_DELAY_MS + MAX_LFO_TIME_MOD_MS) * (float)sr / 1000.0f);
    delay_left = InterpolatedDelayLine1(max_delay + 4);
    delay_right = InterpolatedDelayLine1(max_delay + 4);
    float c = std::min(std::max(filter_cutoff.target_value(), 20.0f), (float)sr / 2.0f);
    filter_coeffs = SvfCoefficients();
    if (!filter_coeffs.set((SvfType)filter_type, sr, c, FILTER_RESONANCE)) return false;
    lfo = Lfo(sr, (double)lfo_rate.target_value(), (Lfo::Waveform)lfo_shape);
    dc_left = DcFilter(sr, 5.0); dc_right = DcFilter(sr, 5.0);
    feedback_left = 0; feedback_right = 0;
    return true;
  }
  void process(float* out, size_t len, uint64_t) override {
    float srf = (float)sample_rate;
    for (size_t i = 0; i + 2 <= len; i += 2) {
      float li = out[i], ri = out[i + 1];
      float lfo_val = lfo.run();
      if (lfo_rate.need_ramp()) { float r = lfo_rate.next_value(); lfo.set_rate(sample_rate, (double)r); }
      float base_delay_ms = delay_time.next_value();
      float time_mod_ms = lfo_val * lfo_depth_time.next_value() * MAX_LFO_TIME_MOD_MS;
      float delay_ms = std::max(base_delay_ms + time_mod_ms, 1.0f);
      float delay_samples = delay_ms * 0.001f * srf;
      float filter_depth = lfo_depth_filter.next_value();
      float filter_mod = std::pow(2.0f, lfo_val * filter_depth * 2.0f);
      float c = std::min(std::max(filter_cutoff.next_value() * filter_mod, 20.0f), srf / 2.0f);
      filter_coeffs.set((SvfType)filter_type, sample_rate, c, FILTER_RESONANCE);
      float base_fb = feedback.next_value();
      float fb_depth = lfo_depth_feedback.next_value();
      float fb = std::min(std::max(base_fb + lfo_val * fb_depth * (1.0f - std::fabs(base_fb)), 0.0f), 0.999f);
      float drv = drive.next_value();
      float wet = wet_mix.next_value();
      float width = stereo_width.next_value();
      float wet_l, wet_r;
      if (mode == 0) {
        float l_in = li + feedback_left * fb;
        float dl = delay_left.process(l_in, 0.0f, delay_samples);
        float cl = process_feedback(filter_left, filter_coeffs, dc_left, dl, drv);
        feedback_left = cl;
        float r_in = ri + feedback_right * fb;
        float dr = delay_right.process(r_in, 0.0f, delay_samples);
        float cr = process_feedback(filter_right, filter_coeffs, dc_right, dr, drv);
        feedback_right = cr;
        wet_l = cl; wet_r = cr;
      } else {
        float mono = (li + ri) * 0.5f;
        float l_in = mono + feedback_right * fb;
        float dl = delay_left.process(l_in, 0.0f, delay_samples);
        float cl = process_feedback(filter_left, filter_coeffs, dc_left, dl, drv);
        float r_in = feedback_left * fb;
        float dr = delay_right.process(r_in, 0.0f, delay_samples);
        float cr = process_feedback(filter_right, filter_coeffs, dc_right, dr, drv);
        feedback_left = cl; feedback_right = cr;
        wet_l = cl; wet_r = cr;
      }
      float dry_gain = std::min((1.0f - wet) * 2.0f, 1.0f);
      float wet_gain = std::min(wet * 2.0f, 1.0f);
      float ol = li * dry_gain + wet_l * wet_gain;
      float orr = ri * dry_gain + wet_r * wet_gain;
      float mid = (ol + orr) * 0.5f;
      float side = (ol - orr) * 0.5f;
      out[i] = mid + side * width;
      out[i + 1] = mid - side * width;
    }
  }
  bool process_tail(size_t& f) const override {
    if (drive.target_value() > 0.0f) return false;
    double delay_ms = (double)(delay_time.target_value() + MAX_LFO_TIME_MOD_MS);
    double fb = (double)std::fabs(feedback.target_value());
    if (fb >= 0.9999) { f = USIZE_MAX; return true; }
    if (fb < 0.001) { f = f64_as_usize(std::ceil(delay_ms * (double)sample_rate / 1000.0)); return true; }
    double ds = delay_ms * (double)sample_rate / 1000.0;
    double decay = ds + ds * std::log10(0.001) / std::log10(fb);
    f = std::max<size_t>(f64_as_usize(std::ceil(decay)), 1);
    return true;
  }
  bool process_parameter_update(uint32_t id, const ParamUpdate& u) override {
    if (id == fourcc("mode")) mode = enum_from_update(u, 2);
    else if (id == fourcc("dlay")) delay_time.apply_update(u);
    else if (id == fourcc("fdbk")) feedback.apply_update(u);
    else if (id == fourcc("ftyp")) filter_type = enum_from_update(u, 3);
    else if (id == fourcc("cuto")) filter_cutoff.apply_update(u);
    else if (id == fourcc("driv")) drive.apply_update(u);
    else if (id == fourcc("wet_")) wet_mix.apply_update(u);
    else if (id == fourcc("wdth")) stereo_width.apply_update(u);
    else if (id == fourcc("lfor")) lfo_rate.apply_update(u);
    else if (id == fourcc("lfos")) { lfo_shape = enum_from_update(u, 7); lfo.waveform = (Lfo::Waveform)lfo_shape; }
    else if (id == fourcc("lfdt")) lfo_depth_time.apply_update(u);
    else if (id == fourcc("ldfb")) lfo_depth_feedback.apply_update(u);
    else if (id == fourcc("lfdf")) lfo_depth_filter.apply_update(u);
    else return false;
    return true;
  }
};

// ---- src/effect/reverb.rs:38-615 -----------------------------------------------------------------------------------------
struct ReverbDelayLine2 {  // reverb.rs:518-604
  std::vector<double> buffer;  // [size+1][2]
  size_t count = 1, delay = 1;
  double feedback[2] = {0, 0};
  double depth;
  double vib_phase[2];
  ReverbDelayLine2(size_t size, double depth_, double p0, double p1) : buffer((size + 1) * 2, 0.0), depth(depth_) { vib_phase[0] = p0; vib_phase[1] = p1; }
  void flush() { std::fill(buffer.begin(), buffer.end(), 0.0); }
  void get(double vib_depth, double blend, double out[2]) const {
    for (int ch = 0; ch < 2; ++ch) {
      double offset = (std::sin(vib_phase[ch]) + 1.0) * vib_depth;
      double working = (double)count + offset;
      double wf = std::floor(working);
      double frac = working - wf;
      size_t wi = (size_t)wf;
      size_t r1 = wi; if (r1 > delay) r1 -= delay + 1;
      size_t r2 = wi + 1; if (r2 > delay) r2 -= delay + 1;
      double v1 = buffer[r1 * 2 + ch], v2 = buffer[r2 * 2 + ch];
      double ip = v1 * (1.0 - frac) + v2 * frac;
      ip = (1.0 - blend) * ip + (v1 * blend);
      out[ch] = ip;
    }
  }
  void set(double l, double r) { buffer[count * 2] = l + feedback[0]; buffer[count * 2 + 1] = r + feedback[1]; }
  void step(double speed) {
    count += 1;
    if (count > delay) count = 0;
    vib_phase[0] += depth * speed;
    vib_phase[1] += depth * speed;
  }
  void set_delay(size_t d) { delay = std::min(d, buffer.size() / 2 - 1); }
};

struct ReverbEffect : Effect {
  uint32_t sample_rate = 0; size_t channel_count = 0;
  SmoothedParam<LinearSmoothed> room_size;
  SmoothedParam<ExpSmoothed> wet;
  BiquadCoefficients ca, cb, cc;
  BiquadFilter a_l, a_r, b_l, b_r, c_l, c_r;
  uint32_t fpd_l, fpd_r;
  std::vector<ReverbDelayLine2> lines;  // a..h
  AllpassDelayLine2 ai, aj, ak, al;
  DelayLine2 m;
  // injected state: fpd[2], vib_phase[8][2]
  ReverbEffect(float room, float wet_, const uint32_t fpd[2], const double vib[16])
      : fpd_l(fpd[0]), fpd_r(fpd[1]), ai(4511), aj(4311), ak(3911), al(3311), m(3111) {
    room_size.value = LinearSmoothed(0, 0.01f);
    room_size.from_description({fourcc("room"), 0.0f, 1.0f, 0.6f, SC_LIN});
    wet.from_description({fourcc("wet "), 0.0f, 1.0f, 0.35f, SC_LIN});
    room_size.init_value(room);
    wet.init_value(wet_);
    static const size_t sizes[8] = {8111, 7511, 7311, 6911, 6311, 6111, 5511, 4911};
    static const double depths[8] = {0.003251, 0.002999, 0.002917, 0.002749, 0.002503, 0.002423, 0.002146, 0.002088};
    for (int i = 0; i < 8; ++i) lines.emplace_back(sizes[i], depths[i], vib[i * 2], vib[i * 2 + 1]);
  }
  const char* name() const override { return "Reverb"; }
  size_t weight() const override { return 5; }
  bool process_message(uint32_t message) override {  // ReverbEffectMessage::Reset (reverb.rs:469-487)
    if (message != 1u) return false;
    for (auto& l : lines) l.flush();
    ai.flush(); aj.flush(); ak.flush(); al.flush(); m.flush();
    return true;
  }
  bool initialize(uint32_t sr, size_t ch, size_t) override {
    sample_rate = sr; channel_count = ch;
    if (ch != 2) return false;
    room_size.set_sample_rate(sr); wet.set_sample_rate(sr);
    return true;
  }
  void update_filter_coefs(float cutoff) {
    float c = std::min(std::max(cutoff, 20.0f), (float)sample_rate / 2.0f);
    if (!ca.set(BiquadType::Lowpass, sample_rate, c, 1.618034f, 0.0f)) return;
    if (!cb.set(BiquadType::Lowpass, sample_rate, c, 0.618034f, 0.0f)) return;
    cc.set(BiquadType::Lowpass, sample_rate, c, 0.5f, 0.0f);
  }
  size_t update_delay_sizes(double size) {
    static const double mult[8] = {79.0, 73.0, 71.0, 67.0, 61.0, 59.0, 53.0, 47.0};
    for (int i = 0; i < 8; ++i) lines[i].set_delay(f64_as_usize(mult[i] * size));
    ai.set_delay(f64_as_usize(43.0 * size));
    aj.set_delay(f64_as_usize(41.0 * size));
    ak.set_delay(f64_as_usize(37.0 * size));
    al.set_delay(f64_as_usize(31.0 * size));
    return f64_as_usize(29.0 * size);
  }
  void process_frame(float* frame, double blend, double regen, size_t predelay, double w) {
    const double vib_speed = 0.1, vib_depth = 7.0;
    double il = (double)frame[0], ir = (double)frame[1];
    if (std::fabs(il) < 1.18e-23) il = (double)fpd_l * 1.18e-17;
    if (std::fabs(ir) < 1.18e-23) ir = (double)fpd_r * 1.18e-17;
    double dry_l = il, dry_r = ir;
    m.process(predelay, il, ir);
    il = a_l.process_sample(ca, il);
    ir = a_r.process_sample(ca, ir);
    il *= w; ir *= w;
    il = std::sin(il); ir = std::sin(ir);
    double i_l = il, i_r = ir; ai.process(i_l, i_r);
    double j_l = i_l, j_r = i_r; aj.process(j_l, j_r);
    double k_l = j_l, k_r = j_r; ak.process(k_l, k_r);
    double l_l = k_l, l_r = k_r; al.process(l_l, l_r);
    lines[0].set(l_l, l_r); lines[1].set(k_l, k_r); lines[2].set(j_l, j_r); lines[3].set(i_l, i_r);
    lines[4].set(i_l, i_r); lines[5].set(j_l, j_r); lines[6].set(k_l, k_r); lines[7].set(l_l, l_r);
    for (int i = 0; i < 8; ++i) lines[i].step(vib_speed);
    double o[8][2];
    for (int i = 0; i < 8; ++i) lines[i].get(vib_depth, blend, o[i]);
    for (int ch = 0; ch < 2; ++ch) {
      double A = o[0][ch], B = o[1][ch], C = o[2][ch], D = o[3][ch], E = o[4][ch], F = o[5][ch], G = o[6][ch], H = o[7][ch];
      lines[0].feedback[ch] = (A - (B + C + D)) * regen;
      lines[1].feedback[ch] = (B - (A + C + D)) * regen;
      lines[2].feedback[ch] = (C - (A + B + D)) * regen;
      lines[3].feedback[ch] = (D - (A + B + C)) * regen;
      lines[4].feedback[ch] = (E - (F + G + H)) * regen;
      lines[5].feedback[ch] = (F - (E + G + H)) * regen;
      lines[6].feedback[ch] = (G - (E + F + H)) * regen;
      lines[7].feedback[ch] = (H - (E + F + G)) * regen;
    }
    il = (o[0][0] + o[1][0] + o[2][0] + o[3][0] + o[4][0] + o[5][0] + o[6][0] + o[7][0]) / 8.0;
    ir = (o[0][1] + o[1][1] + o[2][1] + o[3][1] + o[4][1] + o[5][1] + o[6][1] + o[7][1]) / 8.0;
    il = b_l.process_sample(cb, il);
    ir = b_r.process_sample(cb, ir);
    il = std::min(std::max(il, -1.0), 1.0);
    ir = std::min(std::max(ir, -1.0), 1.0);
    il = std::asin(il); ir = std::asin(ir);
    il = c_l.process_sample(cc, il);
    ir = c_r.process_sample(cc, ir);
    if (w != 1.0) { il += dry_l * (1.0 - w); ir += dry_r * (1.0 - w); }
    frame[0] = (float)il; frame[1] = (float)ir;
  }
  struct Derived { float cutoff; double size, blend, regen; };
  static Derived derive(double room, double w) {
    Derived d;
    d.cutoff = (float)(10000.0 - (room * w * 3000.0));
    d.size = (room * room * 75.0) + 25.0;
    double t = 1.0 - (0.82 - (((1.0 - room) * 0.7) + (d.size * 0.002)));
    double depth_factor = 1.0 - (t * t) * (t * t);  // powi(4) == ((t*t)*(t*t)) in LLVM's expansion
    d.blend = 0.955 - (d.size * 0.007);
    d.regen = depth_factor * 0.5;
    return d;
  }
  void process(float* out, size_t len, uint64_t) override {
    if (room_size.need_ramp() || wet.need_ramp()) {
      for (size_t i = 0; i + 2 <= len; i += 2) {
        double room = (double)room_size.next_value();
        double w = (double)wet.next_value();
        Derived d = derive(room, w);
        size_t predelay = update_delay_sizes(d.size);
        update_filter_coefs(d.cutoff);
        process_frame(out + i, d.blend, d.regen, predelay, w);
      }
    } else {
      double room = (double)room_size.target_value();
      double w = (double)wet.target_value();
      Derived d = derive(room, w);
      size_t predelay = update_delay_sizes(d.size);
      update_filter_coefs(d.cutoff);
      for (size_t i = 0; i + 2 <= len; i += 2) process_frame(out + i, d.blend, d.regen, predelay, w);
    }
  }
  bool process_tail(size_t& f) const override {
    double room = (double)room_size.target_value();
    double size = (room * room * 75.0) + 25.0;
    size_t max_delay = f64_as_usize(79.0 * size);
    double t = 1.0 - (0.82 - (((1.0 - room) * 0.7) + (size * 0.002)));
    double fb = 1.0 - (t * t) * (t * t);
    if (fb >= 1.0) { f = USIZE_MAX; return true; }
    if (fb == 0.0) { f = max_delay; return true; }
    f = max_delay + f64_as_usize((double)max_delay * std::log10(0.001) / std::log10(fb));
    return true;
  }
  bool process_parameter_update(uint32_t id, const ParamUpdate& u) override {
    if (id == fourcc("room")) room_size.apply_update(u);
    else if (id == fourcc("wet ")) wet.apply_update(u);
    else return false;
    return true;
  }
};

}  // namespace po
