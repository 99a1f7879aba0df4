// ORACLE -- TEST INFRASTRUCTURE ONLY (see po_dsp.hpp header). CPU restatement of the reference's granular
// playback: src/generator/sampler/granular.rs (GrainWindow LUTs :104-216, GranularParameters :239-331,
// GrainPool :340-934, Grain :961-1120).
//
// Deterministic subset: the reference seeds every GrainPool's SmallRng from the OS (granular.rs:413), so any
// setting that lets a random draw reach the audio (variation, spray, pan_spread > 0, Random direction) is not
// reproducible in the reference itself and is rejected at the C-ABI. With those at zero every draw is
// multiplied by 0 (granular.rs:563-571, 826-858), so the draws themselves are skipped here. The modulation
// matrix has no routings by default (src/modulation/matrix.rs), so all modulation buffers are zero.
// Parity pinning: no reference test asserts audio for this path; pinned by the derived vectors of
// tests/test_oracle_kat.py (window LUT points, Catmull-Rom identities, trigger cadence).
#pragma once
#include "po_dsp.hpp"

namespace po {

enum class GrainOverlapMode : uint32_t { Cloud = 0, Sequential = 1 };
enum class GrainPlaybackDirection : uint32_t { Forward = 0, Backward = 1, Random = 2 };
enum GrainWindowMode : uint32_t { GW_Hann = 0, GW_Blackman, GW_Triangle, GW_Tukey, GW_Trapezoid, GW_Exponential, GW_RampUp, GW_RampDown, GW_COUNT };

inline float sequential_crossfade_point(uint32_t mode) {  // granular.rs:76-93
  switch (mode) {
    case GW_Hann: case GW_Blackman: case GW_Triangle: case GW_Tukey: return 0.5f;
    case GW_Trapezoid: return 0.9f;
    default: return 0.8f;
  }
}

// granular.rs:104-216
struct GrainWindow {
  static constexpr size_t N = 2048;
  static constexpr size_t MASK = N - 1;
  std::vector<float> luts;  // [GW_COUNT][N]
  GrainWindow() : luts(GW_COUNT * N) {
    const float PI = 3.14159265358979323846264338327950288f;
    for (size_t i = 0; i < N; ++i) {
      float phase = (float)i / (float)N;
      lut(GW_Hann)[i] = 0.5f * (1.0f - std::cos(2.0f * PI * phase));
      float pi_phase = PI * phase;
      lut(GW_Blackman)[i] = 0.42f - 0.5f * std::cos(2.0f * pi_phase) + 0.08f * std::cos(4.0f * pi_phase);
      lut(GW_Triangle)[i] = phase < 0.5f ? 2.0f * phase : 2.0f * (1.0f - phase);
      const float alpha = 0.5f;
      const float width = alpha / 2.0f;
      if (phase < width) { float u = phase / width; lut(GW_Tukey)[i] = 0.5f * (1.0f - std::cos(PI * u)); }
      else if (phase > 1.0f - width) { float u = (1.0f - phase) / width; lut(GW_Tukey)[i] = 0.5f * (1.0f - std::cos(PI * u)); }
      else lut(GW_Tukey)[i] = 1.0f;
      const float ramp_width = 0.1f;
      if (phase < ramp_width) lut(GW_Trapezoid)[i] = phase / ramp_width;
      else if (phase > 1.0f - ramp_width) lut(GW_Trapezoid)[i] = (1.0f - phase) / ramp_width;
      else lut(GW_Trapezoid)[i] = 1.0f;
      const float decay_rate = 6.0f;
      float center_dist = std::fabs(phase - 0.5f);
      lut(GW_Exponential)[i] = std::exp(-decay_rate * center_dist);
      if (phase < 0.9f) lut(GW_RampUp)[i] = phase / 0.9f;
      else { float u = (phase - 0.9f) / 0.1f; lut(GW_RampUp)[i] = 0.5f * (1.0f + std::cos(PI * u)); }
      if (phase < 0.1f) { float u = phase / 0.1f; lut(GW_RampDown)[i] = 0.5f * (1.0f - std::cos(PI * u)); }
      else lut(GW_RampDown)[i] = 1.0f - ((phase - 0.1f) / 0.9f);
    }
  }
  float* lut(uint32_t mode) { return luts.data() + (size_t)mode * N; }
  const float* lut(uint32_t mode) const { return luts.data() + (size_t)mode * N; }
  float sample(uint32_t mode, double phase) const {  // granular.rs:201-215
    double index_float = phase * (double)(N - 1);
    size_t index = f64_as_usize(index_float) & MASK;
    float fraction = (float)(index_float - std::trunc(index_float));
    size_t next_index = (index + 1) & MASK;
    const float* l = lut(mode);
    if (index < N - 1) return l[index] * (1.0f - fraction) + l[next_index] * fraction;
    return l[N - 1];
  }
  static const GrainWindow& shared() { static GrainWindow w; return w; }  // GRAIN_WINDOW_LUT (granular.rs:221)
};

struct GranularParameters {  // granular.rs:239-331
  GrainOverlapMode overlap_mode = GrainOverlapMode::Cloud;
  uint32_t window = GW_Triangle;
  float size = 100.0f, density = 10.0f, variation = 0.0f, spray = 0.0f, pan_spread = 0.0f;
  GrainPlaybackDirection playback_direction = GrainPlaybackDirection::Forward;
  float position = 0.5f, step = 0.0f;
  bool validate() const {
    if (size < 1.0f || size > 1000.0f) return false;
    if (density < 1.0f || density > 100.0f) return false;
    if (spray < 0.0f || spray > 1.0f) return false;
    if (variation < 0.0f || variation > 1.0f) return false;
    if (pan_spread < 0.0f || pan_spread > 1.0f) return false;
    if (position < 0.0f || position > 1.0f) return false;
    if (step < -4.0f || step > 4.0f) return false;
    return true;
  }
};

inline double rem_euclid_f64(double a, double b) { double r = std::fmod(a, b); return r < 0.0 ? r + std::fabs(b) : r; }
inline float gran_rem_euclid_f32(float a, float b) { float r = std::fmod(a, b); return r < 0.0f ? r + std::fabs(b) : r; }

struct GrainOutput { float envelope, panning, position; };

struct Grain {  // granular.rs:961-1120
  bool active = false;
  float volume = 1.0f, panning = 0.0f;
  double position = 0.0, increment = 0.0;
  size_t samples_remaining = 0;
  double window_phase = 0.0, window_increment = 0.0;
  uint32_t window_mode = GW_Triangle;
  bool has_loop = false; double loop_start = 0, loop_end = 0;
  void activate(uint32_t mode, double pos, double speed, float vol, float pan, size_t grain_size_samples,
                size_t file_length_frames, bool reverse, bool has_loop_, double ls, double le) {
    active = true;
    window_mode = mode;
    position = std::min(std::max(pos, 0.0), 1.0);
    volume = std::min(std::max(vol, 0.0f), 100.0f);
    panning = std::min(std::max(pan, -1.0f), 1.0f);
    samples_remaining = grain_size_samples;
    has_loop = has_loop_; loop_start = ls; loop_end = le;
    double base_increment = file_length_frames > 0 ? speed / (double)file_length_frames : 0.0;
    increment = base_increment * (reverse ? -1.0 : 1.0);
    window_phase = 0.0;
    window_increment = grain_size_samples > 0 ? 1.0 / (double)grain_size_samples : 0.0;
  }
  void deactivate() { active = false; samples_remaining = 0; }
  GrainOutput process(const GrainWindow& w) {
    float envelope_value = w.sample(window_mode, window_phase);
    float pos = (float)position;
    position += increment;
    window_phase += window_increment;
    samples_remaining = samples_remaining > 0 ? samples_remaining - 1 : 0;
    if (has_loop) {
      double loop_len = loop_end - loop_start;
      if (loop_len > 0.0) position = loop_start + rem_euclid_f64(position - loop_start, loop_len);
    } else if (position < 0.0) {
      position += 1.0;
    } else if (position > 1.0) {
      position -= 1.0;
    }
    if (samples_remaining == 0) active = false;
    return {envelope_value * volume, panning, pos};
  }
};

struct GrainPool {  // granular.rs:340-934, POOL_SIZE = GRAIN_POOL_SIZE = 100 (voice.rs:33)
  static constexpr size_t POOL_SIZE = 100;
  static constexpr float ENVELOPE_THRESHOLD = 0.001f;
  GrainOverlapMode overlap_mode = GrainOverlapMode::Cloud;
  std::array<Grain, POOL_SIZE> grain_pool;
  std::vector<size_t> active_grain_indices;
  bool has_primary = false; size_t primary_grain_index = 0;
  std::shared_ptr<std::vector<float>> sample_buffer;
  bool has_loop_range = false; float loop_start = 0, loop_end = 0;  // sample_loop_range (normalised)
  bool playing_loop_range = false, trigger_new_grains = true;
  float trigger_phase = 0.0f;
  double speed = 1.0;
  float volume = 1.0f, panning = 0.0f, playhead = 0.0f;
  uint32_t sample_rate;

  GrainPool(uint32_t sr, std::shared_ptr<std::vector<float>> buf, bool has_loop, float ls, float le)
      : sample_buffer(buf), has_loop_range(has_loop), loop_start(ls), loop_end(le), sample_rate(sr) {
    active_grain_indices.reserve(POOL_SIZE);
  }
  static double fold_into_loop_range(double position, double ls, double le) {
    double loop_len = le - ls;
    return loop_len > 0.0 ? ls + rem_euclid_f64(position - ls, loop_len) : ls;
  }
  bool is_exhausted() const { return !trigger_new_grains && active_grain_indices.empty(); }
  float playback_position(const GranularParameters& p, float position_mod) const {
    float base = p.step == 0.0f ? p.position : playhead;
    if (position_mod != 0.0f) base += position_mod;
    if (playing_loop_range && has_loop_range) base = (float)fold_into_loop_range((double)base, (double)loop_start, (double)loop_end);
    return gran_rem_euclid_f32(base, 1.0f);
  }
  void start(const GranularParameters& p, double speed_, float volume_, float panning_) {
    trigger_new_grains = true;
    trigger_phase = 1.0f;
    speed = speed_; volume = volume_; panning = panning_;
    playhead = p.position;
    playing_loop_range = false;
  }
  void stop() { trigger_new_grains = false; }
  void reset() {
    active_grain_indices.clear();
    for (auto& g : grain_pool) g.deactivate();
    trigger_new_grains = true;
    has_primary = false;
  }
  bool update_trigger_phase(const GranularParameters& p, float density_mod) {
    if (overlap_mode == GrainOverlapMode::Sequential) return true;
    float density_mult = 1.0f + density_mod;
    float density = std::min(std::max(p.density * density_mult, 1.0f), 100.0f);
    float trigger_increment = density / (float)sample_rate;
    trigger_phase += trigger_increment;
    if (trigger_phase >= 1.0f) { trigger_phase -= 1.0f; return true; }
    return false;
  }
  bool activate_new_grain(const GranularParameters& p, double position) {
    size_t index = POOL_SIZE;
    for (size_t i = 0; i < POOL_SIZE; ++i) if (!grain_pool[i].active) { index = i; break; }
    if (index == POOL_SIZE) return false;
    // zero variation / pan spread: every random draw is multiplied by 0 (see header)
    float vol = volume * 1.0f;
    const float size_scale = 1.0f;
    float size_mult = 1.0f + 0.0f;
    float grain_size_ms = std::min(std::max(p.size * size_mult, 1.0f), 1000.0f);
    size_t grain_size = std::max<size_t>(f64_as_usize((double)(grain_size_ms * size_scale * (float)sample_rate / 1000.0f)), 2);
    float pan = std::min(std::max(panning + 0.0f, -1.0f), 1.0f);
    double varied_speed = speed * 1.0;
    bool reverse = p.playback_direction == GrainPlaybackDirection::Backward;
    bool use_loop = playing_loop_range && has_loop_range;
    grain_pool[index].activate(p.window, position, varied_speed, vol, pan, grain_size, sample_buffer->size(), reverse,
                               use_loop, (double)loop_start, (double)loop_end);
    auto it = std::find(active_grain_indices.begin(), active_grain_indices.end(), index);
    if (it != active_grain_indices.end()) active_grain_indices.erase(it);
    active_grain_indices.push_back(index);
    if (overlap_mode == GrainOverlapMode::Sequential) { has_primary = true; primary_grain_index = index; }
    return true;
  }
  bool try_trigger_grain(const GranularParameters& p) {  // granular.rs:524-603, all modulation inputs zero
    if (overlap_mode != p.overlap_mode) { overlap_mode = p.overlap_mode; has_primary = false; }
    if (overlap_mode == GrainOverlapMode::Sequential && has_primary) {
      const Grain& g = grain_pool[primary_grain_index];
      if (g.active && g.window_phase < (double)sequential_crossfade_point(p.window)) return false;
    }
    if (!trigger_new_grains || !update_trigger_phase(p, 0.0f)) return false;
    double grain_position = (double)playback_position(p, 0.0f) + 0.0;
    if (playing_loop_range && has_loop_range) grain_position = fold_into_loop_range(grain_position, (double)loop_start, (double)loop_end);
    grain_position = rem_euclid_f64(grain_position, 1.0);
    return activate_new_grain(p, grain_position);
  }
  void advance_playhead(size_t buffer_frame_count, float step, float speed_mod) {  // granular.rs:607-640
    float speed_mult = 1.0f + speed_mod;
    float modulated_step = step * speed_mult;
    float position_increment = modulated_step / (float)buffer_frame_count;
    playhead += position_increment;
    if (has_loop_range) {
      if (playing_loop_range) playhead = (float)fold_into_loop_range((double)playhead, (double)loop_start, (double)loop_end);
      else if (playhead >= loop_start && playhead < loop_end) playing_loop_range = true;
      else { if (playhead >= 1.0f) playhead -= 1.0f; else if (playhead < 0.0f) playhead += 1.0f; }
    } else if (playhead >= 1.0f) playhead -= 1.0f;
    else if (playhead < 0.0f) playhead += 1.0f;
  }
  float sample_at_position(float normalized_pos) const {  // granular.rs:901-933
    const std::vector<float>& b = *sample_buffer;
    size_t len = b.size();
    size_t max_index = len - 1;
    float float_index = normalized_pos * (float)max_index;
    size_t index = std::min(f64_as_usize((double)float_index), max_index);
    float fraction = float_index - (float)index;
    size_t i1 = index;
    size_t i2 = i1 < max_index ? i1 + 1 : 0;
    size_t i0 = i1 > 0 ? i1 - 1 : max_index;
    size_t i3 = i2 < max_index ? i2 + 1 : 0;
    float y0 = b[i0], y1 = b[i1], y2 = b[i2], y3 = b[i3];
    float a = -0.5f * y0 + 1.5f * y1 - 1.5f * y2 + 0.5f * y3;
    float bb = y0 - 2.5f * y1 + 2.0f * y2 - 0.5f * y3;
    float c = -0.5f * y0 + 0.5f * y2;
    float d = y1;
    return a * fraction * fraction * fraction + bb * fraction * fraction + c * fraction + d;
  }
  size_t process(float* out, size_t len, size_t channel_count, const GranularParameters& p) {  // granular.rs:642-784
    const GrainWindow& w = GrainWindow::shared();
    size_t sample_frame_count = sample_buffer->size();
    bool move_playhead = p.step != 0.0f && sample_frame_count > 0;
    assert(channel_count == 2 && "oracle: granular playback is restated for stereo output only");
    for (size_t f = 0; f + 2 <= len; f += 2) {
      try_trigger_grain(p);
      if (move_playhead) advance_playhead(sample_frame_count, p.step, 0.0f);
      for (size_t gi : active_grain_indices) {
        Grain& g = grain_pool[gi];
        if (g.active) {
          GrainOutput o = g.process(w);
          if (o.envelope > ENVELOPE_THRESHOLD) {
            float sample = sample_at_position(o.position);
            float windowed = sample * o.envelope;
            float left_gain = (1.0f - o.panning) * 0.5f;
            float right_gain = (1.0f + o.panning) * 0.5f;
            out[f] += windowed * left_gain;
            out[f + 1] += windowed * right_gain;
          }
        }
      }
    }
    active_grain_indices.erase(std::remove_if(active_grain_indices.begin(), active_grain_indices.end(),
                                              [&](size_t i) { return !grain_pool[i].active; }),
                               active_grain_indices.end());
    return len;
  }
};

}  // namespace po
