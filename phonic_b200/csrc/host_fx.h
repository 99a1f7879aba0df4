// Host side of the effect chain: builds the initial device state of each effect exactly as the
// reference's constructors + Effect::initialize do (host libm == the libm the Rust reference links),
// and resolves ParameterValueUpdate::{Raw, Normalized} to plain values (src/parameter/*.rs).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/phonic_b200.h"
#include "effects_state.h"

namespace pbh {
using namespace pb;

constexpr float F32_EPS_H = 1.1920929e-07f;

inline uint32_t cc4(const char* s) {
  return ((uint32_t)(uint8_t)s[0] << 24) | ((uint32_t)(uint8_t)s[1] << 16) | ((uint32_t)(uint8_t)s[2] << 8) | (uint32_t)(uint8_t)s[3];
}

// src/utils.rs:41-51
inline float db_to_linear(float v) {
  const float K = 2.302585092994046f / 20.0f;
  if (std::isnan(v)) return NAN;
  if (v == 0.0f) return 1.0f;
  if (v > -200.0f) return std::exp(v * K);
  return 0.0f;
}

// FloatParameter descriptor + ParameterScaling (src/parameter/float.rs:126-141, scaling.rs:45-75)
struct ParamDesc {
  uint32_t id;
  float min, max;
  int scaling;  // 0 linear, 1 exponential(a), 2 enum(count = max + 1), 3 decibel(a = min dB, b = max dB), 4 boolean
  float a;
  float b = 0.0f;
};
inline float denormalize(const ParamDesc& d, float n) {
  n = std::min(std::max(n, 0.0f), 1.0f);
  if (d.scaling == 2) return std::round(n * d.max);
  if (d.scaling == 4) return n >= 0.5f ? 1.0f : 0.0f;  // BooleanParameter::denormalize_value (boolean.rs:83-86)
  float s = n;
  if (d.scaling == 1) s = std::pow(n, d.a);
  else if (d.scaling == 3) {  // ParameterScaling::Decibel (scaling.rs:64-72)
    const float db = d.a + n * (d.b - d.a);
    const float lo = db_to_linear(d.a), hi = db_to_linear(d.b);
    s = (db_to_linear(db) - lo) / (hi - lo);
  }
  return d.min + s * (d.max - d.min);
}
inline float resolve_plain(const ParamDesc& d, float v, bool normalized) {
  if (normalized) return denormalize(d, v);
  if (d.scaling == 2) return std::min(std::max(std::round(v), 0.0f), d.max);
  if (d.scaling == 4) return v != 0.0f ? 1.0f : 0.0f;
  return std::min(std::max(v, d.min), d.max);
}

inline const std::vector<ParamDesc>& param_table(uint32_t kind) {
  static const float PI_F = 3.14159265358979323846f;
  static const std::vector<ParamDesc> filter = {
      {cc4("type"), 0, 3, 2, 0}, {cc4("cuto"), 20.0f, 20000.0f, 1, 2.5f}, {cc4("fltq"), 0.001f, 4.0f, 0, 0}};
  static const std::vector<ParamDesc> eq5 = [] {
    std::vector<ParamDesc> v;
    const float bmax[5] = {1.0f, 4.0f, 4.0f, 4.0f, 1.0f};
    for (int i = 0; i < 5; ++i) {
      char g[5] = {'g', 'a', 'n', (char)('1' + i), 0}, f[5] = {'f', 'r', 'q', (char)('1' + i), 0}, b[5] = {'b', 'w', '_', (char)('1' + i), 0};
      v.push_back({cc4(g), -20.0f, 20.0f, 0, 0});
      v.push_back({cc4(f), 20.0f, 20000.0f, 1, 2.5f});
      v.push_back({cc4(b), 0.0001f, bmax[i], 0, 0});
    }
    return v;
  }();
  static const std::vector<ParamDesc> comp = {
      {cc4("thrs"), -60.0f, 0.0f, 0, 0}, {cc4("rato"), 1.0f, 20.0f, 0, 0}, {cc4("knee"), 0.0f, 12.0f, 0, 0},
      {cc4("attk"), 0.001f, 0.5f, 0, 0}, {cc4("rels"), 0.1f, 2.0f, 0, 0}, {cc4("gain"), -24.0f, 24.0f, 0, 0},
      {cc4("look"), 0.001f, 0.2f, 0, 0}};
  static const std::vector<ParamDesc> chorus = {
      {cc4("rate"), 0.01f, 10.0f, 1, 2.0f}, {cc4("phas"), 0.0f, PI_F, 0, 0}, {cc4("dpth"), 0.0f, 1.0f, 0, 0},
      {cc4("fdbk"), -1.0f, 1.0f, 0, 0}, {cc4("dlay"), 0.0f, 100.0f, 0, 0}, {cc4("wet_"), 0.0f, 1.0f, 0, 0},
      {cc4("fltt"), 0, 2, 2, 0}, {cc4("fltf"), 20.0f, 20000.0f, 1, 2.5f}, {cc4("fltq"), 0.0f, 1.0f, 0, 0}};
  static const std::vector<ParamDesc> delay = {
      {cc4("mode"), 0, 1, 2, 0}, {cc4("dlay"), 1.0f, 4000.0f, 0, 0}, {cc4("fdbk"), 0.0f, 1.0f, 0, 0},
      {cc4("ftyp"), 0, 2, 2, 0}, {cc4("cuto"), 20.0f, 20000.0f, 1, 2.5f}, {cc4("driv"), 0.0f, 1.0f, 0, 0},
      {cc4("wet_"), 0.0f, 1.0f, 0, 0}, {cc4("wdth"), 0.0f, 1.0f, 0, 0}, {cc4("lfor"), 0.01f, 10.0f, 1, 2.0f},
      {cc4("lfos"), 0, 6, 2, 0}, {cc4("lfdt"), -1.0f, 1.0f, 0, 0}, {cc4("ldfb"), -1.0f, 1.0f, 0, 0},
      {cc4("lfdf"), -1.0f, 1.0f, 0, 0}};
  static const std::vector<ParamDesc> reverb = {{cc4("room"), 0.0f, 1.0f, 0, 0}, {cc4("wet "), 0.0f, 1.0f, 0, 0}};
  static const std::vector<ParamDesc> gain = {{cc4("gain"), 0.000001f, 15.848932f, 3, -60.0f, 24.0f}, {cc4("dcfm"), 0, 3, 2, 0}};
  static const std::vector<ParamDesc> panning = {{cc4("pan "), -1.0f, 1.0f, 0, 0}, {cc4("wdth"), 0.0f, 2.0f, 0, 0},
                                                 {cc4("invl"), 0, 1, 4, 0}, {cc4("invr"), 0, 1, 4, 0}};
  static const std::vector<ParamDesc> gate = {{cc4("thrs"), -60.0f, 0.0f, 0, 0}, {cc4("attk"), 0.001f, 0.5f, 0, 0}, {cc4("hold"), 0.0f, 2.0f, 0, 0},
                                              {cc4("rels"), 0.01f, 2.0f, 0, 0}, {cc4("rnge"), -60.0f, 0.0f, 0, 0}};
  static const std::vector<ParamDesc> distortion = {{cc4("type"), 0, 4, 2, 0}, {cc4("driv"), 0.0f, 4.0f, 0, 0}, {cc4("mix "), 0.0f, 1.0f, 0, 0}};
  static const std::vector<ParamDesc> none;
  switch (kind) {
    case FX_GATE: return gate;
    case FX_DISTORTION: return distortion;
    case FX_GAIN: return gain;
    case FX_PANNING: return panning;
    case FX_FILTER: return filter;
    case FX_EQ5: return eq5;
    case FX_COMPRESSOR: return comp;
    case FX_CHORUS: return chorus;
    case FX_DELAY: return delay;
    case FX_REVERB: return reverb;
  }
  return none;
}

// ---- host mirrors of the coefficient math (same formulas as effects.cuh, host libm) ----------------------
inline void biquad_apply_h(BiquadCoef& c) {
  const double PI = 3.14159265358979323846;
  double g = std::tan(PI * (double)c.cutoff / (double)c.sample_rate);
  double k = 1.0 / (double)c.q;
  double a = 0.0;
  if (c.type == 6) { a = std::pow(10.0, (double)c.gain / 40.0); k = 1.0 / ((double)c.q * a); }
  else if (c.type == 7) { a = std::pow(10.0, (double)c.gain / 40.0); g = g / std::sqrt(a); }
  else if (c.type == 8) { a = std::pow(10.0, (double)c.gain / 40.0); g = g * std::sqrt(a); }
  c.a1 = 1.0 / (1.0 + g * (g + k));
  c.a2 = g * c.a1;
  c.a3 = g * c.a2;
  switch (c.type) {
    case 0: c.m0 = 0.0; c.m1 = 0.0; c.m2 = 1.0; break;
    case 1: c.m0 = 1.0; c.m1 = -k; c.m2 = -1.0; break;
    case 2: c.m0 = 0.0; c.m1 = 1.0; c.m2 = 0.0; break;
    case 3: c.m0 = 1.0; c.m1 = -k; c.m2 = 0.0; break;
    case 4: c.m0 = 1.0; c.m1 = -k; c.m2 = -2.0; break;
    case 5: c.m0 = 1.0; c.m1 = -2.0 * k; c.m2 = 0.0; break;
    case 6: c.m0 = 1.0; c.m1 = k * (a * a - 1.0); c.m2 = 0.0; break;
    case 7: c.m0 = 1.0; c.m1 = k * (a - 1.0); c.m2 = a * a - 1.0; break;
    default: c.m0 = a * a; c.m1 = k * (1.0 - a) * a; c.m2 = 1.0 - a * a; break;
  }
}
inline void biquad_set_h(BiquadCoef& c, uint32_t type, uint32_t sr, float cutoff, float q, float gain) {
  if (c.type != type || c.sample_rate != sr || c.cutoff != cutoff || c.q != q || c.gain != gain) {
    c.type = type; c.sample_rate = sr; c.cutoff = cutoff; c.q = q; c.gain = gain;
    biquad_apply_h(c);
  }
}
inline void svf_set_h(SvfCoef& c, uint32_t type, uint32_t sr, float cutoff, float res) {
  c.type = type; c.sample_rate = sr; c.cutoff = cutoff; c.resonance = res;
  const double PI = 3.14159265358979323846;
  c.g = std::tan(PI * (double)cutoff / (double)sr);
  c.k = std::max(2.0 * (1.0 - (double)res * 0.97), 0.03);
  c.a1 = 1.0 / (1.0 + c.g * (c.g + c.k));
  c.a2 = c.g * c.a1;
  c.a3 = c.g * c.a2;
}

inline ExpSm exp_init(float v) { return ExpSm{v, v}; }
inline LinSm lin_init(float v, float step, float comp) { return LinSm{v, v, step, step * comp, 0}; }
inline SpringSm spring_init(float v, size_t duration) { return SpringSm{v, 0.0f, v, 5.5f / (float)duration}; }
inline uint32_t next_pow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }
inline uint32_t ceil_u32(float v) { float c = std::ceil(v); return c > 0.0f ? (uint32_t)c : 0u; }

struct FxBuild {
  std::vector<uint8_t> state;  // kind-specific state blob
  size_t aux_doubles = 0;      // delay-line storage the effect needs (zero-initialised)
  std::string error;
  int code = 0;
};

// aux offsets inside the blob are relative; the renderer rebases them by adding the arena offset.
template <class T>
inline T& blob(FxBuild& b) { b.state.assign(sizeof(T), 0); return *reinterpret_cast<T*>(b.state.data()); }

inline FxBuild build_filter(const pb200_filter_params* p, uint32_t sr) {  // filter.rs:84-164
  FxBuild b;
  FilterState& s = blob<FilterState>(b);
  const float comp = 44100.0f / (float)sr;
  std::memset(&s.coef, 0, sizeof(s.coef));
  s.coef.type = 0xFFFFFFFFu;  // force first set
  biquad_set_h(s.coef, 0 /*Lowpass*/, 44100, 22050.0f, 0.707f, 0.0f);
  float cutoff = 20000.0f, q = 0.707f;
  uint32_t ftype = 0;
  if (p) {
    if (p->filter_type > 3 || !(p->cutoff >= 20.0f && p->cutoff <= 20000.0f) || !(p->q >= 0.001f && p->q <= 4.0f)) {
      b.code = PB200_ERR_PARAMETER; b.error = "Value out of bounds"; return b;
    }
    ftype = p->filter_type; cutoff = p->cutoff; q = p->q;
    const uint32_t map[4] = {0, 2, 3, 1};  // FilterEffectType -> BiquadFilterType (filter.rs:33-42)
    float c = std::min(std::max(cutoff, 20.0f), 44100.0f / 2.0f);
    biquad_set_h(s.coef, map[ftype], 44100, c, q, 0.0f);  // NB: stays at 44.1 kHz
  }
  const uint32_t map[4] = {0, 2, 3, 1};
  s.filter_type = map[ftype];
  // initialize(): clamp cutoff to the real nyquist, set_cutoff re-applies only when it changed
  float c = std::min(std::max(s.coef.cutoff, 20.0f), (float)sr / 2.0f);
  if (s.coef.cutoff != c) { s.coef.cutoff = c; biquad_apply_h(s.coef); }
  s.cutoff = exp_init(cutoff);
  s.q = lin_init(q, 0.01f, comp);
  return b;
}

inline FxBuild build_eq5(uint32_t sr) {  // eq5.rs:153-170, 266-295
  FxBuild b;
  Eq5State& s = blob<Eq5State>(b);
  const float comp = 44100.0f / (float)sr;
  const float fdef[5] = {100.0f, 1000.0f, 4000.0f, 8000.0f, 12000.0f};
  const float bdef[5] = {1.0f, 4.0f, 4.0f, 4.0f, 1.0f};
  for (int i = 0; i < 5; ++i) {
    s.gains[i] = exp_init(0.0f);
    s.freqs[i] = exp_init(fdef[i]);
    s.bws[i] = lin_init(bdef[i], 0.01f, comp);
    uint32_t type = i == 0 ? 7 : (i == 4 ? 8 : 6);
    float c = std::min(std::max(fdef[i], 20.0f), (float)sr / 2.0f);
    std::memset(&s.coef[i], 0, sizeof(BiquadCoef));
    s.coef[i].type = 0xFFFFFFFFu;
    biquad_set_h(s.coef[i], type, sr, c, bdef[i], 0.0f);
  }
  return b;
}

inline FxBuild build_compressor(const pb200_compressor_params* p, uint32_t sr) {  // compressor.rs:96-228
  FxBuild b;
  CompState& s = blob<CompState>(b);
  s.threshold = p ? p->threshold : -12.0f;
  s.ratio = p ? p->ratio : 8.0f;
  s.knee = p ? p->knee : 3.0f;
  s.attack_time = p ? p->attack_time : 0.02f;
  s.release_time = p ? p->release_time : 2.0f;
  s.makeup = exp_init(p ? p->makeup_gain : 6.0f);
  s.lookahead_time = p ? p->lookahead_time : 0.04f;
  s.atk_coeff = s.attack_time > 0.0f ? std::exp(-1.0f / (s.attack_time * (float)sr)) : 0.0f;
  s.rel_coeff = s.release_time > 0.0f ? std::exp(-1.0f / (s.release_time * (float)sr)) : 0.0f;
  s.env_cur = s.ratio >= 20.0f ? -120.0f : 0.0f;
  s.delay_frames = ceil_u32(s.lookahead_time * (float)sr);
  s.buf_frames = s.delay_frames ? next_pow2(s.delay_frames) : 0;
  s.mask = s.buf_frames ? s.buf_frames - 1 : 0;
  // capacity for the largest lookahead a parameter update may ask for (0.2 s)
  s.aux_capacity_frames = next_pow2(ceil_u32(0.2f * (float)sr) + 1);
  s.aux = 0;
  b.aux_doubles = (size_t)s.aux_capacity_frames * 2;
  return b;
}

inline double dc_mode_hz(uint32_t m) { return m == 1 ? 1.0 : (m == 3 ? 20.0 : 5.0); }  // DcFilterMode::hz; Off -> Default

inline FxBuild build_gain(const pb200_gain_params* p, uint32_t sr) {  // gain.rs:85-146
  FxBuild b;
  GainState& s = blob<GainState>(b);
  if (p && p->dc_filter_mode > 3) { b.code = PB200_ERR_PARAMETER; b.error = "bad DC filter mode"; return b; }
  s.gain = exp_init(p ? db_to_linear(std::min(std::max(p->gain_db, -60.0f), 24.0f)) : 1.0f);
  s.dc_mode = p ? p->dc_filter_mode : 0;
  s.dc_r = 1.0 - (6.28318530717958647692 * dc_mode_hz(s.dc_mode) / (double)sr);
  s.dc_x1[0] = s.dc_x1[1] = s.dc_y1[0] = s.dc_y1[1] = 0.0;
  return b;
}

// coefficients of GateEffect::update_coefficients (gate.rs:83-95) + EnvelopeFollower::set_*_time (envelope.rs:36-49)
inline void gate_update_coefficients_h(GateState& s, uint32_t sr) {
  const float srf = (float)sr;
  s.env_atk = s.attack_time > 0.0f ? std::exp(-1.0f / (s.attack_time * srf)) : 0.0f;
  s.env_rel = s.release_time > 0.0f ? std::exp(-1.0f / (s.release_time * srf)) : 0.0f;
  s.attack_coeff = std::exp(-1.0f / (s.attack_time * srf));
  s.release_coeff = std::exp(-1.0f / (s.release_time * srf));
}

inline FxBuild build_gate(const pb200_gate_params* p, uint32_t sr) {  // gate.rs:51-81, 124-150
  FxBuild b;
  GateState& s = blob<GateState>(b);
  s.threshold = p ? p->threshold : -30.0f;
  s.attack_time = p ? p->attack_time : 0.005f;
  s.hold_time = p ? p->hold_time : 0.1f;
  s.release_time = p ? p->release_time : 0.2f;
  s.range = p ? p->range : -60.0f;
  if (!(s.threshold >= -60.0f && s.threshold <= 0.0f) || !(s.attack_time >= 0.001f && s.attack_time <= 0.5f) ||
      !(s.hold_time >= 0.0f && s.hold_time <= 2.0f) || !(s.release_time >= 0.01f && s.release_time <= 2.0f) ||
      !(s.range >= -60.0f && s.range <= 0.0f)) {
    b.code = PB200_ERR_PARAMETER; b.error = "Value out of bounds"; return b;
  }
  s.env_cur = -120.0f;
  s.hold_counter = 0;
  s.gate_gain_db = s.range;
  gate_update_coefficients_h(s, sr);
  return b;
}

// DistortionType::shape_function (distortion.rs:124-190) on the host: only used to build the RMS compensation tables at
// construction time, as DistortionEffect::new() does (distortion.rs:256-268); the audio path evaluates it on device.
inline float dist_shape_h(uint32_t type, float sample, float drive) {
  const float t = drive / 4.0f;
  if (type == 0) {
    const float gain = 1.0f + (t * t) * (15.0f - 1.0f), x = sample * gain;
    if (x >= 1.0f) return 1.0f;
    if (x > -1.0f) return gain <= 1.0f ? sample : (3.0f / 2.0f) * (x - ((x * x) * x) / 3.0f);
    return -1.0f;
  }
  if (type == 1) {
    const float gain = 1.0f + (t * t) * (25.0f - 1.0f), threshold = 1.0f / gain;
    return std::min(std::max(sample, -threshold), threshold) * gain;
  }
  if (type == 2) {
    const float gain = 1.0f + (0.6f * (t * t) + 0.4f * t) * (20.0f - 1.0f);
    const float dc = std::exp((0.1f * sample) / (0.0253f * 1.68f)) - 1.0f;
    return 2.0f / 3.14159265358979323846f * std::atan(dc * gain);
  }
  if (type == 3) {
    const float gain = 1.0f + (1.0f - std::exp(-3.0f * t)) * (30.0f - 1.0f), a = sample * gain;
    const float sat = a < 0.0f ? -1.0f * (1.0f - std::exp(-std::fabs(a))) : 1.0f * (1.0f - std::exp(-std::fabs(a)));
    return 1.5f * (sat + std::fabs(sat));
  }
  const float gain = 1.0f + (t * t) * (4.0f - 1.0f), x = sample * gain, threshold = 1.0f / gain;
  if (x > threshold || x < -threshold) return std::fabs(std::fmod(std::fabs(x - threshold), threshold * 4.0f) - threshold * 2.0f) - threshold;
  return x;
}
inline float dist_rms_compensation_h(uint32_t type, float drive) {  // distortion.rs:84-122
  static const float PART[5][2] = {{1.0f, 0.60f}, {2.7f, 0.25f}, {5.3f, 0.10f}, {9.1f, 0.03f}, {14.6f, 0.02f}};
  float peak = 0.0f;
  for (int k = 0; k < 5; ++k) peak += PART[k][1];
  float in_sq = 0.0f, out_sq = 0.0f;
  for (int i = 0; i < 256; ++i) {
    const float t = 6.28318530717958647692f * ((float)i + 0.5f) / 256.0f;
    float sum = 0.0f;
    for (int k = 0; k < 5; ++k) sum += PART[k][1] * std::sin(PART[k][0] * t);
    const float x = sum / peak, y = dist_shape_h(type, x, drive);
    in_sq += x * x;
    out_sq += y * y;
  }
  const float in_rms = std::sqrt(in_sq / 256.0f), out_rms = std::sqrt(out_sq / 256.0f);
  return out_rms > 1e-10f ? in_rms / out_rms : 1.0f;
}
inline FxBuild build_distortion(const pb200_distortion_params* p, uint32_t sr) {  // distortion.rs:226-253
  FxBuild b;
  DistState& s = blob<DistState>(b);
  if (p && (p->distortion_type > 4 || !(p->drive >= 0.0f && p->drive <= 4.0f) || !(p->mix >= 0.0f && p->mix <= 1.0f))) {
    b.code = PB200_ERR_PARAMETER; b.error = "Value out of bounds"; return b;
  }
  const float comp = 44100.0f / (float)sr;
  s.type = p ? p->distortion_type : 2u;
  s.drive = lin_init(p ? p->drive : 0.0f, 0.01f, comp);
  s.mix = exp_init(p ? p->mix : 1.0f);
  struct Luts { float v[5][256]; };
  static const Luts luts = [] {  // the tables depend on nothing but the type: built once per process (thread-safe static)
    Luts l;
    for (uint32_t ty = 0; ty < 5; ++ty)
      for (int i = 0; i < 256; ++i) l.v[ty][i] = dist_rms_compensation_h(ty, (float)i / 255.0f * 4.0f);
    return l;
  }();
  std::memcpy(s.lut, luts.v, sizeof(s.lut));
  return b;
}

inline FxBuild build_panning() {  // pan.rs:52-60
  FxBuild b;
  PanState& s = blob<PanState>(b);
  s.pan = exp_init(0.0f);
  s.width = exp_init(1.0f);
  s.invert_l = s.invert_r = 0;
  return b;
}

inline float rem_euclid1(float a) { float r = std::fmod(a, 1.0f); return r < 0.0f ? r + 1.0f : r; }

inline FxBuild build_chorus(const pb200_chorus_params* p, uint32_t sr) {  // chorus.rs:143-309
  FxBuild b;
  ChorusState& s = blob<ChorusState>(b);
  const float comp = 44100.0f / (float)sr;
  const float PI_F = 3.14159265358979323846f;
  if (p && p->filter_type > 2) { b.code = PB200_ERR_PARAMETER; b.error = "bad chorus filter type"; return b; }
  s.rate = lin_init(p ? p->rate : 1.0f, 0.005f, comp);
  s.phase = lin_init(p ? p->phase : PI_F / 2.0f, 0.001f, comp);
  s.depth = exp_init(p ? p->depth : 0.25f);
  s.feedback = exp_init(p ? p->feedback : 0.5f);
  s.delay = spring_init(p ? p->delay : 12.0f, 1000);
  s.wet = exp_init(p ? p->wet : 0.5f);
  s.filter_type = p ? p->filter_type : 0;
  s.filter_freq = exp_init(p ? p->filter_freq : 20000.0f);
  s.filter_res = exp_init(p ? p->filter_resonance : 0.0f);
  s.lfo_range = 256.0f * ((float)sr / 44100.0f);
  uint32_t max_depth = ceil_u32(s.lfo_range);
  uint32_t max_delay = ceil_u32(100.0f * (float)sr / 1000.0f);
  uint32_t n = next_pow2(2 + max_delay + 2 * max_depth + 1);
  s.dl = IDelay{n - 1, 0, 0, 0};
  s.dr = IDelay{n - 1, 0, n, 0};
  b.aux_doubles = (size_t)n * 2;
  float c = std::min(std::max(s.filter_freq.target, 20.0f), (float)sr / 2.0f);
  if (!(s.filter_res.target >= 0.0f && s.filter_res.target <= 1.0f)) { b.code = PB200_ERR_PARAMETER; b.error = "Invalid filter resonance"; return b; }
  svf_set_h(s.coef, s.filter_type, sr, c, s.filter_res.target);
  // reset(): current_phase = 0, Lfo::new(sr, rate, Sine), phases from current_phase (+ offset)
  s.current_phase = 0.0;
  float inc = (float)((double)s.rate.current / (double)sr);
  s.left_osc = LfoSt{rem_euclid1((float)s.current_phase / 6.28318530717958647692f), inc, 0};
  s.right_osc = LfoSt{rem_euclid1((float)(s.current_phase + (double)s.phase.current) / 6.28318530717958647692f), inc, 0};
  return b;
}

inline FxBuild build_delay(uint32_t sr) {  // delay.rs:180-332
  FxBuild b;
  DelayState& s = blob<DelayState>(b);
  s.delay_time = spring_init(375.0f, 20000);
  s.feedback = exp_init(0.5f);
  s.cutoff = exp_init(6000.0f);
  s.drive = exp_init(0.0f);
  s.wet = exp_init(0.5f);
  s.width = exp_init(0.5f);
  s.lfo_rate = exp_init(1.0f);
  s.lfo_dt = exp_init(0.0f); s.lfo_dfb = exp_init(0.0f); s.lfo_dflt = exp_init(0.0f);
  s.mode = 0; s.filter_type = 0; s.lfo_shape = 0;
  uint32_t max_delay = ceil_u32((4000.0f + 50.0f) * (float)sr / 1000.0f);
  uint32_t n = next_pow2(max_delay + 4);
  s.dl = IDelay{n - 1, 0, 0, 0};
  s.dr = IDelay{n - 1, 0, n, 0};
  b.aux_doubles = (size_t)n * 2;
  float c = std::min(std::max(6000.0f, 20.0f), (float)sr / 2.0f);
  svf_set_h(s.coef, 0, sr, c, 0.302f);
  s.lfo = LfoSt{0.0f, (float)(1.0 / (double)sr), 0};
  s.dc_r = 1.0 - (6.28318530717958647692 * 5.0 / (double)sr);
  return b;
}

inline FxBuild build_reverb(const pb200_reverb_params* p, uint32_t sr) {  // reverb.rs:94-159, 388-407
  FxBuild b;
  ReverbState& s = blob<ReverbState>(b);
  const float comp = 44100.0f / (float)sr;
  if (!(p->room_size >= 0.0f && p->room_size <= 1.0f) || !(p->wet >= 0.0f && p->wet <= 1.0f)) {
    b.code = PB200_ERR_PARAMETER; b.error = "Value out of bounds"; return b;
  }
  s.room = lin_init(p->room_size, 0.01f, comp);
  s.wet = exp_init(p->wet);
  s.fpd_l = p->fpd[0]; s.fpd_r = p->fpd[1];
  std::memset(&s.ca, 0, sizeof(BiquadCoef)); std::memset(&s.cb, 0, sizeof(BiquadCoef)); std::memset(&s.cc, 0, sizeof(BiquadCoef));
  // BiquadFilterCoefficients::default(): Lowpass, sr 0, all zeros -> first set() always applies
  const uint32_t sizes[8] = {8111, 7511, 7311, 6911, 6311, 6111, 5511, 4911};
  const double depths[8] = {0.003251, 0.002999, 0.002917, 0.002749, 0.002503, 0.002423, 0.002146, 0.002088};
  size_t off = 0;
  for (int i = 0; i < 8; ++i) {
    RvLine& L = s.lines[i];
    L.aux = (uint32_t)off; L.size = sizes[i]; L.count = 1; L.delay = 1; L.depth = depths[i];
    L.feedback[0] = L.feedback[1] = 0.0;
    L.vib_phase[0] = p->vib_phase[i * 2]; L.vib_phase[1] = p->vib_phase[i * 2 + 1];
    off += (size_t)(sizes[i] + 1) * 2;
  }
  const uint32_t asz[4] = {4511, 4311, 3911, 3311};
  for (int i = 0; i < 4; ++i) { s.ap[i] = RvAllpass{(uint32_t)off, asz[i], 0, 0}; off += (size_t)asz[i] * 2; }
  uint32_t mn = next_pow2(3111);
  s.m_aux = (uint32_t)off; s.m_mask = mn - 1; s.m_write_pos = 0;
  off += (size_t)mn * 2;
  b.aux_doubles = off;
  return b;
}

// rebase aux offsets stored in a state blob by `base` doubles
inline void rebase_aux(uint32_t kind, std::vector<uint8_t>& state, uint32_t base) {
  switch (kind) {
    case FX_COMPRESSOR: reinterpret_cast<CompState*>(state.data())->aux += base; break;
    case FX_CHORUS: { auto* s = reinterpret_cast<ChorusState*>(state.data()); s->dl.aux += base; s->dr.aux += base; break; }
    case FX_DELAY: { auto* s = reinterpret_cast<DelayState*>(state.data()); s->dl.aux += base; s->dr.aux += base; break; }
    case FX_REVERB: {
      auto* s = reinterpret_cast<ReverbState*>(state.data());
      for (int i = 0; i < 8; ++i) s->lines[i].aux += base;
      for (int i = 0; i < 4; ++i) s->ap[i].aux += base;
      s->m_aux += base;
      break;
    }
    default: break;
  }
}

}  // namespace pbh
