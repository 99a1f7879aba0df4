// ORACLE -- TEST INFRASTRUCTURE ONLY (see po_dsp.hpp header).
// Primitive-level entry points so tests can pin the restatement against the reference's own
// unit-test assertions and the derived known-answer vectors (Appendix D).
#include "po_effects.hpp"

using namespace po;

extern "C" {

// CubicResampler::process on one interleaved buffer; returns produced samples, *consumed in samples.
uint64_t po_test_cubic(const float* in, uint64_t in_len, uint32_t channels, uint32_t in_rate, uint32_t out_rate,
                       float* out, uint64_t out_len, uint64_t* consumed, float* ratio_out) {
  CubicResampler rs(in_rate, out_rate, channels);
  auto res = rs.process(in, in_len, out, out_len);
  if (consumed) *consumed = res.first;
  if (ratio_out) *ratio_out = rs.interpolators[0].ratio;
  return res.second;
}

// PreloadedFileSource::write (test `resampling`, src/source/file/preloaded.rs:486-533)
uint64_t po_test_preloaded_write(const float* data, uint64_t samples, uint32_t channels, uint32_t rate, uint32_t out_rate,
                                 float* out, uint64_t out_len) {
  auto fb = std::make_shared<AudioFileBuffer>();
  fb->buffer.assign(data, data + samples); fb->sample_rate = rate; fb->channel_count = channels;
  FilePlaybackOptions o;
  PreloadedFileSource src(fb, o, out_rate);
  return src.write(out, out_len, SourceTime{0});
}

// BiquadFilterCoefficients::set + impulse response through BiquadFilter
void po_test_biquad(uint32_t type, uint32_t sr, float cutoff, float q, float gain, double* coeffs6, float* impulse, uint32_t n) {
  BiquadCoefficients c;
  c.set((BiquadType)type, sr, cutoff, q, gain);
  coeffs6[0] = c.a1; coeffs6[1] = c.a2; coeffs6[2] = c.a3; coeffs6[3] = c.m0; coeffs6[4] = c.m1; coeffs6[5] = c.m2;
  BiquadFilter f;
  for (uint32_t i = 0; i < n; ++i) impulse[i] = (float)f.process_sample(c, i == 0 ? 1.0 : 0.0);
}

// AHDSR: run the envelope for n frames after note_on, note_off at frame `off_at` (or never if >= n);
// writes the per-frame output and stage.
void po_test_ahdsr(uint64_t a_ns, uint64_t h_ns, uint64_t d_ns, float sustain, uint64_t r_ns, float as, float ds, float rs,
                   uint32_t sr, uint32_t n, uint32_t off_at, float* out, uint32_t* stages, float* rates3) {
  AhdsrParameters p;
  AhdsrParameters::create(p, Duration::from_nanos(a_ns), as, Duration::from_nanos(h_ns), Duration::from_nanos(d_ns), ds, sustain,
                          Duration::from_nanos(r_ns), rs);
  p.set_sample_rate(sr);
  if (rates3) { rates3[0] = p.attack_rate; rates3[1] = p.decay_rate; rates3[2] = p.release_rate; }
  AhdsrEnvelope e;
  e.note_on(p, 1.0f);
  for (uint32_t i = 0; i < n; ++i) {
    if (i == off_at) e.note_off(p);
    out[i] = e.run(p);
    stages[i] = (uint32_t)e.stage;
  }
}

// ExponentialSmoothedValue: number of next() calls that still ramp from `from` to `to`
uint32_t po_test_exp_smoother(float from, float to, uint32_t sr, float* last_ramped) {
  ExpSmoothed s(from, sr);
  s.set_target(to);
  uint32_t n = 0;
  while (s.need_ramp() && n < 10000000u) { float v = s.next(); if (last_ramped) *last_ramped = v; ++n; }
  return n;
}

// VolumeFader fade-out: frames until |cur - target| < 1e-4
uint32_t po_test_fader(uint64_t dur_ns, uint32_t sr, float* inertia) {
  VolumeFader f(1, sr);
  f.start_fade_out(Duration::from_nanos(dur_ns));
  if (inertia) *inertia = f.inertia;
  uint32_t n = 0;
  float x = 1.0f;
  while (f.state == VolumeFader::IsRunning && n < 10000000u) { x = 1.0f; f.process(&x, 1); ++n; }
  return n;
}

void po_test_panning(float pan, float* l, float* r) { panning_factors(pan, *l, *r); }
float po_test_db_to_linear(float v) { return db_to_linear(v); }
float po_test_linear_to_db(float v) { return linear_to_db(v); }

// note -> (speed, resampler output rate, f32 ratio) for a file at in_rate played at out_rate
void po_test_note_ratio(uint32_t note, uint32_t in_rate, uint32_t out_rate, double* speed, uint32_t* rate, float* ratio) {
  double s = speed_from_note((uint8_t)note);
  uint32_t r = f64_as_u32((double)out_rate / s);
  *speed = s; *rate = r; *ratio = (float)((double)in_rate / (double)r);
}

// DistortionType::shape_function / rms_compensation (src/effect/distortion.rs:84-190)
float po_test_dist_shape(uint32_t type, float sample, float drive) { return dist_shape(type, sample, drive); }
float po_test_dist_compensation(uint32_t type, float drive) { return dist_rms_compensation(type, drive); }

}
