// Voice path on device: PreloadedFileSource + CubicResampler + VolumeFader + Amplified/Panned
// smoothing + AHDSR, one thread per voice, one CTA per group (Sampler or file playback).
//
// Every f32/f64 operation keeps the reference's operation order and is separately rounded
// (the translation unit is compiled with -fmad=false): the voice path is bit-exact against the
// reference's scalar code, which is what makes loop/seek/event indices exact and keeps the f32
// phase recurrence of cubic.rs:73-110 in lock-step.
#pragma once
#include "dev_structs.h"

namespace pb {

#define PB_DEV __device__ __forceinline__

constexpr float F32_EPS = 1.1920929e-07f;
constexpr float SMOOTH_INERTIA = 1.0f / 256.0f;  // ExponentialSmoothedValue::DEFAULT_INERTIA

struct RenderConsts {
  uint32_t sample_rate;
  float rate_comp;  // 44100 / sample_rate as f32 (ExponentialSmoothedValue::sample_rate_comp)
};

// ---- ExponentialSmoothedValue (src/utils/smoothing.rs:195-223) --------------------------------------
PB_DEV bool exp_need_ramp(const ExpSm& s, float comp) {
  float add = (s.target - s.current) * SMOOTH_INERTIA * comp;
  return fabsf(add) > F32_EPS * 100.0f;
}
PB_DEV float exp_next(ExpSm& s, float comp) {
  float add = (s.target - s.current) * SMOOTH_INERTIA * comp;
  if (fabsf(add) > F32_EPS * 100.0f) {
    s.current += add;
    return s.current;
  }
  return s.target;
}
PB_DEV void exp_set_target(ExpSm& s, float t, float comp) {
  s.target = t;
  if (!exp_need_ramp(s, comp)) s.current = s.target;
}

// src/utils.rs:56-62
PB_DEV void panning_factors(float pan, float& l, float& r) {
  const float POWER = 0.70710678118654752440f;
  float normalized = (fminf(fmaxf(pan, -1.0f), 1.0f) + 1.0f) / 2.0f;
  l = sqrtf(1.0f - normalized) / POWER;
  r = sqrtf(normalized) / POWER;
}

// ---- CubicInterpolator (src/utils/resampler/cubic.rs:117-142) -----------------------------------------
PB_DEV void hist_push(float* h, float v) { h[3] = h[2]; h[2] = h[1]; h[1] = h[0]; h[0] = v; }
PB_DEV float hermite(const float* h, float fraction) {
  float ym1 = h[3], y0 = h[2], y1 = h[1], y2 = h[0];
  float c0 = y0;
  float c1 = (y1 - ym1) * 0.5f;
  float c2 = ym1 - y0 * 2.5f + y1 * 2.0f - y2 * 0.5f;
  float c3 = (y2 - ym1) * 0.5f + (y0 - y1) * 1.5f;
  return ((c3 * fraction + c2) * fraction + c1) * fraction + c0;
}

PB_DEV uint32_t f64_as_u32(double v) {  // Rust `as u32`: saturating, NaN -> 0
  if (!(v > 0.0)) return 0u;
  if (v >= 4294967295.0) return 0xFFFFFFFFu;
  return (uint32_t)v;
}

// FileSourceImpl::update_speed (src/source/file/common.rs:141-169) + CubicResampler::update
PB_DEV void update_speed(VoiceState& v, uint32_t in_rate, uint32_t out_rate) {
  double speed_diff = v.target_speed - v.current_speed;
  if (v.glide_rate > 0.0f && fabs(speed_diff) > 0.0001) {
    double semitone_diff = fabs(12.0 * log2(v.target_speed / v.current_speed));
    float duration_secs = (float)semitone_diff / v.glide_rate;
    if (duration_secs > 0.0f) {
      float duration_frames = duration_secs * (float)out_rate;
      double step = (v.target_speed - v.current_speed) / (double)duration_frames;
      double change = step * 64.0;
      if (fabs(v.target_speed - v.current_speed) < fabs(change)) v.current_speed = v.target_speed;
      else v.current_speed += change;
    } else {
      v.current_speed = v.target_speed;
    }
  } else {
    v.current_speed = v.target_speed;
  }
  uint32_t new_rate = f64_as_u32((double)out_rate / v.current_speed);
  v.ratio = (float)((double)in_rate / (double)new_rate);
}

PB_DEV void resampler_reset(VoiceState& v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) v.hidx[i] = -1;
  v.sub_pos = 0.0f;
  v.initialized = 0;
}

// PreloadedFileSource::set_speed (preloaded.rs:182-193)
PB_DEV void file_set_speed(VoiceState& v, double speed, float glide, uint32_t in_rate, uint32_t out_rate) {
  if (!v.finished) {
    v.to_next_speed_update = 0;
    v.target_speed = speed;
    v.glide_rate = glide > 0.0f ? glide : 0.0f;
    if (v.glide_rate == 0.0f) {
      v.current_speed = speed;
      update_speed(v, in_rate, out_rate);
    }
  }
}
// PreloadedFileSource::stop (preloaded.rs:196-209); VolumeFader::start_fade_out (fader.rs:68-92)
PB_DEV void file_stop(VoiceState& v, const GroupParams& gp) {
  if (!v.finished) {
    if (gp.has_fade_out) {
      float from = (v.fader_state == FADER_RUNNING) ? v.fader_cur : 1.0f;
      v.fader_state = FADER_RUNNING;
      v.fader_cur = from;
      v.fader_tgt = 0.0f;
      v.fader_inertia = gp.fade_out_inertia;
    } else {
      v.stopped_exhausted = v.pos_eof;
      v.finished = 1;
    }
  }
}
// PreloadedFileSource::seek (preloaded.rs:138-146); index resolved on the host
PB_DEV void file_seek(VoiceState& v, uint32_t pos) {
  if (!v.finished) {
    v.playback_pos = pos;
    resampler_reset(v);
  }
}
// PreloadedFileSource::reset (preloaded.rs:212-230)
PB_DEV void file_reset(VoiceState& v) {
  if (!v.finished) { v.stopped_exhausted = v.pos_eof; v.finished = 1; }
  v.playback_pos = 0;
  v.repeat_count = v.repeat;
  v.pos_eof = 0;
  v.finished = 0;
  resampler_reset(v);
  v.fader_state = FADER_STOPPED;
  v.fader_cur = 1.0f;
  v.fader_tgt = 1.0f;
}

// ---- AhdsrEnvelope (src/utils/ahdsr.rs:402-552) --------------------------------------------------------
PB_DEV void env_note_on(VoiceState& v, const GroupParams& gp, float volume) {
  v.env_target = volume;
  if (gp.attack_rate == 3.402823466e+38f) {
    v.env_out = volume;
    if (!gp.hold_is_zero) { v.env_stage = ENV_HOLD; v.env_hold = gp.hold_samples; }
    else v.env_stage = ENV_DECAY;
  } else {
    v.env_out = 0.0f;
    v.env_stage = ENV_ATTACK;
  }
}
PB_DEV void env_note_off(VoiceState& v, const GroupParams& gp) {
  if (!gp.release_is_zero) {
    v.env_target = 0.0f;
    v.env_release_out = v.env_out;
    v.env_stage = (v.env_release_out > F32_EPS) ? ENV_RELEASE : ENV_IDLE;
  } else {
    v.env_out = 0.0f; v.env_release_out = 0.0f; v.env_stage = ENV_IDLE;
  }
}
PB_DEV float env_apply_scaling(float value, float scaling) {
  const float EULER_DIV_2 = 2.718281828459045f / 2.0f;
  if (scaling == 0.0f || value == 0.0f) return value;
  float s = -scaling;
  if (s > 0.0f) return powf(value, 1.0f + powf(s, EULER_DIV_2) * 16.0f);
  return 1.0f - powf(1.0f - value, 1.0f + powf(-s, EULER_DIV_2) * 16.0f);
}
PB_DEV float env_run(VoiceState& v, const GroupParams& gp) {
  switch (v.env_stage) {
    case ENV_ATTACK:
      v.env_out += gp.attack_rate;
      if (v.env_out >= v.env_target) {
        v.env_out = v.env_target;
        v.env_target = gp.sustain_level;
        if (!gp.hold_is_zero) { v.env_stage = ENV_HOLD; v.env_hold = gp.hold_samples; }
        else v.env_stage = ENV_DECAY;
      }
      break;
    case ENV_HOLD:
      v.env_hold -= 1.0f;
      if (v.env_hold <= 0.0f) v.env_stage = gp.decay_is_zero ? ENV_SUSTAIN : ENV_DECAY;
      break;
    case ENV_DECAY:
      if (v.env_out > gp.sustain_level) {
        v.env_out -= gp.decay_rate;
        if (v.env_out <= gp.sustain_level) { v.env_out = gp.sustain_level; v.env_stage = ENV_SUSTAIN; }
      } else {
        v.env_out += gp.decay_rate;
        if (v.env_out >= gp.sustain_level) { v.env_out = gp.sustain_level; v.env_stage = ENV_SUSTAIN; }
      }
      break;
    case ENV_RELEASE:
      v.env_out -= v.env_release_out * gp.release_rate;
      if (v.env_out <= 0.001f) { v.env_out = 0.0f; v.env_stage = ENV_IDLE; }
      break;
    default: break;
  }
  if (v.env_stage == ENV_ATTACK && gp.attack_scaling != 0.0f) {
    float progress = v.env_out / fmaxf(v.env_target, F32_EPS);
    return env_apply_scaling(progress, gp.attack_scaling) * v.env_target;
  }
  if (v.env_stage == ENV_DECAY && gp.decay_scaling != 0.0f) {
    float range = fmaxf(fabsf(v.env_target - gp.sustain_level), F32_EPS);
    bool down = v.env_target > gp.sustain_level;
    float progress = down ? (v.env_target - v.env_out) / range : (v.env_out - v.env_target) / range;
    float sp = env_apply_scaling(progress, gp.decay_scaling);
    return down ? v.env_target - (sp * range) : v.env_target + (sp * range);
  }
  if (v.env_stage == ENV_RELEASE && gp.release_scaling != 0.0f) {
    float initial = fmaxf(v.env_out, F32_EPS);
    float progress = 1.0f - (v.env_out / initial);
    float sp = env_apply_scaling(progress, gp.release_scaling);
    return initial * (1.0f - sp);
  }
  return v.env_out;
}

// SamplerVoice::reset (voice.rs:222-236)
PB_DEV void voice_reset(VoiceState& v) {
  if (v.has_note) { file_reset(v); v.has_note = 0; }
  v.has_release = 0;
}

// ---- per write-call context (decisions the reference takes once per `write` call) -----------------------
struct CallCtx {
  uint32_t chunk_left;       // frames still requested in this Source::write call
  uint32_t call_left;        // frames left in the current write_buffer call (glide sub-chunk)
  uint32_t produced_in_call; // output of the current resampler.process call
  uint32_t ls, le;           // active loop range / buffer range in samples
  uint32_t hq_off;           // block-relative frame of the next output (HighQuality voices read the stream scratch there)
  bool new_call;             // a new resampler.process call starts at the next frame
  bool gliding;              // write() took the pitch-slide arm (preloaded.rs:419)
  bool ended;                // write_buffer broke out (EOF) -> no more frames in this call
  bool fader_running, fader_scale;
  bool vol_ramp, vol_scale;
  bool pan_ramp, pan_apply;
  bool env_per_frame;
  float pan_l, pan_r, env_const;
};

PB_DEV void loop_range_samples(const VoiceState& v, const DevBuffer& b, uint32_t& ls, uint32_t& le) {
  ls = 0; le = b.n_samples;
  if (v.repeat > 0) {
    if (v.loop_ovr_start >= 0) { ls = (uint32_t)v.loop_ovr_start * b.channels; le = (uint32_t)v.loop_ovr_end * b.channels; }
    else if (b.loop_start >= 0) { ls = (uint32_t)b.loop_start * b.channels; le = (uint32_t)b.loop_end * b.channels; }
  }
}

// Start of PreloadedFileSource::write + the wrappers' per-call decisions. Returns false when the
// source is finished (write returns 0).
PB_DEV bool voice_begin_call(VoiceState& v, CallCtx& c, const GroupParams& gp, const DevBuffer& b, uint32_t n_frames,
                             float comp, bool with_env, uint32_t call_off) {
  c.chunk_left = n_frames;
  c.hq_off = call_off;
  c.call_left = 0;
  c.produced_in_call = 0;
  c.new_call = true;
  c.ended = false;
  if (v.finished) { c.chunk_left = 0; c.ended = true; return false; }
  c.gliding = v.current_speed != v.target_speed;
  if (!c.gliding) v.to_next_speed_update = 0;
  loop_range_samples(v, b, c.ls, c.le);
  // VolumeFader::process mode (fader.rs:103-108)
  c.fader_running = v.fader_state == FADER_RUNNING;
  c.fader_scale = !c.fader_running && v.fader_tgt != 1.0f;
  // apply_smoothed_gain mode (smoothing.rs:60-71)
  c.vol_ramp = exp_need_ramp(v.vol, comp);
  c.vol_scale = !c.vol_ramp && fabsf(1.0f - v.vol.target) > 0.000001f;
  // apply_smoothed_panning mode (smoothing.rs:74-122)
  c.pan_ramp = exp_need_ramp(v.pan, comp);
  c.pan_apply = !c.pan_ramp && fabsf(v.pan.target) > 0.000001f;
  c.pan_l = 1.0f; c.pan_r = 1.0f;
  if (c.pan_apply) panning_factors(v.pan.target, c.pan_l, c.pan_r);
  // envelope mode (voice.rs:470-486)
  c.env_per_frame = with_env && !(v.env_stage == ENV_SUSTAIN || v.env_stage == ENV_IDLE);
  c.env_const = v.env_out;
  return true;
}

// write_buffer post-call bookkeeping (preloaded.rs:311-329)
PB_DEV void after_process_call(VoiceState& v, CallCtx& c) {
  if (v.playback_pos >= c.le) {
    if (v.repeat_count > 0) {
      if (v.repeat_count != REPEAT_FOREVER) v.repeat_count -= 1;
      v.playback_pos = c.ls;
    } else {
      v.pos_eof = 1;
    }
  }
}

// The interpolator history as values (replay) next to the indices every pass keeps in VoiceState.
struct HistVals { float h[2][4]; };

template <int CC>
PB_DEV void hist_load(HistVals& hv, const VoiceState& v, const float* __restrict__ buf) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hv.h[0][i] = v.hidx[i] >= 0 ? __ldg(buf + v.hidx[i]) : 0.0f;
    hv.h[1][i] = (CC == 2 && v.hidx[i] >= 0) ? __ldg(buf + v.hidx[i] + 1) : 0.0f;
  }
}

// CubicInterpolator::push_sample for the frame at v.playback_pos (cubic.rs:117-122)
template <int CC, bool AUDIO>
PB_DEV void push_frame(VoiceState& v, HistVals& hv, const float* __restrict__ buf) {
  v.hidx[3] = v.hidx[2]; v.hidx[2] = v.hidx[1]; v.hidx[1] = v.hidx[0]; v.hidx[0] = (int32_t)v.playback_pos;
  if (AUDIO) {
    hist_push(hv.h[0], __ldg(buf + v.playback_pos));
    if (CC == 2) hist_push(hv.h[1], __ldg(buf + v.playback_pos + 1));
  }
  v.playback_pos += CC;
}

// One output frame of the resampler inside write_buffer. Returns false on the EOF break.
// AUDIO=false advances exactly the same state (positions, phase, history indices) without touching samples.
template <int CC, bool AUDIO>
PB_DEV bool resample_frame(VoiceState& v, CallCtx& c, HistVals& hv, const float* __restrict__ buf, float& x0, float& x1) {
  for (;;) {
    if (c.new_call) {  // CubicInterpolator::process prologue (cubic.rs:47-69)
      c.new_call = false;
      c.produced_in_call = 0;
      const bool bypass = fabsf(v.ratio - 1.0f) < 0.000001f;
      uint32_t avail = c.le > v.playback_pos ? (c.le - v.playback_pos) / CC : 0u;
      if (!bypass && !v.initialized && avail >= 3) {
        v.initialized = 1;
#pragma unroll
        for (int f = 0; f < 3; ++f) push_frame<CC, AUDIO>(v, hv, buf);
      }
    }
    const bool has_input = v.playback_pos < c.le;
    if (fabsf(v.ratio - 1.0f) < 0.000001f) {  // bypass copy (cubic.rs:53-58)
      if (has_input) {
        if (AUDIO) {
          x0 = __ldg(buf + v.playback_pos);
          x1 = CC == 2 ? __ldg(buf + v.playback_pos + 1) : x0;
        }
        v.playback_pos += CC;
        c.produced_in_call++;
        return true;
      }
    } else if (v.ratio < 1.0f) {  // cubic.rs:72-90
      bool ok = true;
      if (v.sub_pos >= 1.0f) {
        if (has_input) {
          push_frame<CC, AUDIO>(v, hv, buf);
          v.sub_pos -= 1.0f;
        } else {
          ok = false;
        }
      }
      if (ok) {
        if (AUDIO) {
          x0 = hermite(hv.h[0], v.sub_pos);
          x1 = CC == 2 ? hermite(hv.h[1], v.sub_pos) : x0;
        }
        v.sub_pos += v.ratio;
        c.produced_in_call++;
        return true;
      }
    } else {  // cubic.rs:91-110
      bool ok = true;
      while (v.sub_pos < v.ratio) {
        if (v.playback_pos >= c.le) { ok = false; break; }
        push_frame<CC, AUDIO>(v, hv, buf);
        v.sub_pos += 1.0f;
      }
      if (ok) {
        v.sub_pos -= v.ratio;
        if (AUDIO) {
          float fr = 1.0f - v.sub_pos;
          x0 = hermite(hv.h[0], fr);
          x1 = CC == 2 ? hermite(hv.h[1], fr) : x0;
        }
        c.produced_in_call++;
        return true;
      }
    }
    // input exhausted: process() returned early; write_buffer loops (preloaded.rs:311-329)
    after_process_call(v, c);
    if (v.pos_eof && c.produced_in_call == 0) return false;
    c.new_call = true;
  }
}

// Advance the current write call by up to `n` frames. AUDIO=true renders them into `out` (interleaved
// stereo); AUDIO=false (the skeleton pass) only advances the state: positions, f32 phase and ramp
// recurrences, envelope stage machine. Returns the number of frames written (< n only when the
// source ran dry).
template <int CC, bool AUDIO>
PB_DEV uint32_t voice_frames(VoiceState& v, CallCtx& c, HistVals& hv, const GroupParams& gp, const DevBuffer& b,
                             uint32_t out_rate, float comp, uint32_t n, float* __restrict__ out, const bool acc = false) {
  const float* __restrict__ buf = b.data;
  uint32_t f = 0;
  for (; f < n && !c.ended; ++f) {
    if (c.call_left == 0) {  // next write_buffer call (preloaded.rs:419-447)
      if (c.gliding) {
        if (v.to_next_speed_update == 0) {
          if (v.current_speed != v.target_speed) update_speed(v, b.sample_rate, out_rate);
          v.to_next_speed_update = 64;
        }
        c.call_left = min(c.chunk_left, v.to_next_speed_update);
      } else {
        c.call_left = c.chunk_left;
      }
      loop_range_samples(v, b, c.ls, c.le);
      c.new_call = true;
    }
    float x0 = 0.0f, x1 = 0.0f;
    if (!resample_frame<CC, AUDIO>(v, c, hv, buf, x0, x1)) { c.ended = true; break; }
    c.call_left--;
    c.chunk_left--;
    if (c.gliding) v.to_next_speed_update--;
    if (c.call_left == 0) after_process_call(v, c);  // output slice full: post-call check still runs
    // VolumeFader::process (fader.rs:103-116)
    if (c.fader_running) {
      v.fader_cur += (v.fader_tgt - v.fader_cur) * v.fader_inertia;
      x0 *= v.fader_cur; x1 *= v.fader_cur;
    } else if (c.fader_scale) {
      x0 *= v.fader_tgt; x1 *= v.fader_tgt;
    }
    // ChannelMappedSource: mono -> stereo duplicates (buffer.rs:199-206); stereo passes through
    float l = x0, r = x1;
    // AmplifiedSource: per *sample* smoothing (smoothing.rs:61-64)
    if (c.vol_ramp) { l *= exp_next(v.vol, comp); r *= exp_next(v.vol, comp); }
    else if (c.vol_scale) { l *= v.vol.target; r *= v.vol.target; }
    // PannedSource
    if (c.pan_ramp) {
      float pl, pr;
      float p = exp_next(v.pan, comp);
      if (AUDIO) { panning_factors(p, pl, pr); l *= pl; r *= pr; }
    } else if (c.pan_apply) {
      l *= c.pan_l; r *= c.pan_r;
    }
    // AHDSR (voice.rs:470-486)
    if (gp.has_env) {
      float e = c.env_per_frame ? env_run(v, gp) : c.env_const;
      l *= e; r *= e;
    }
    if (AUDIO) {  // acc: Sampler::write adds the voice to the output (sampler.rs:989-1006)
      out[2 * f] = acc ? out[2 * f] + l : l;
      out[2 * f + 1] = acc ? out[2 * f + 1] + r : r;
    }
  }
  return f;
}

// End of PreloadedFileSource::write (preloaded.rs:449-472); `call_end_frame` = absolute frame after the call
PB_DEV void voice_end_call(VoiceState& v, CallCtx& c, uint64_t call_end_frame) {
  if (c.fader_running) {
    if (fabsf(v.fader_cur - v.fader_tgt) < 0.0001f) v.fader_state = FADER_FINISHED;
  }
  bool fade_out_completed = v.fader_state == FADER_FINISHED && v.fader_tgt == 0.0f;
  if (v.pos_eof || fade_out_completed) {
    v.stopped_exhausted = v.pos_eof;
    v.finished = 1;
    v.end_frame = call_end_frame;
  }
}

// Comparison results as 1.0f / 0.0f computed on the FMA pipe: [s >= x] = sat((s - pred(x)) * 2^60) as ONE
// FFMA.SAT (the product and the difference are exact before the single rounding, so the sign is that of
// s - pred(x); a positive difference is at least one ulp >= 2^-60 for |x| >= 2^-36). The skeleton's phase chains
// are single-warp dependent chains: FSET / predicates sit on another pipe and cost ~2x the forwarding latency.
constexpr float STEP_K = 1152921504606846976.0f;  // 2^60
PB_DEV float fma_sat(float a, float b, float c) { float d; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
PB_DEV float f32_pred(float x) { return __int_as_float(__float_as_int(x) - 1); }  // x > 0
PB_DEV float f32_succ(float x) { return __int_as_float(__float_as_int(x) + 1); }  // x > 0
// w such that step(s, w) = [s >= x] for x > 0
PB_DEV float step_weight(float x) { return -(f32_pred(x) * STEP_K); }
PB_DEV float step(float s, float w) { return fma_sat(s, STEP_K, w); }
PB_DEV float fset_ge(float a, float b) { return step(a, step_weight(b)); }                               // b > 0
PB_DEV float fset_lt(float a, float b) { return fma_sat(a, -STEP_K, b * STEP_K); }                       // a, b >= 1

// ---- the resampler's f32 phase recurrence on the fast path -----------------------------------------------------------
// (no input exhaustion possible, 0 < ratio < 64, not the bypass ratio): sub_pos advances exactly as
// CubicInterpolator::process does (cubic.rs:72-110). Everything that only depends on the ratio is resolved once
// (PhaseK); a run of pieces with the same ratio (one simple call) carries (s, p) from piece to piece.
//
// ratio < 1 (cubic.rs:73-89): per frame p = [s >= 1]; s = (s - p) + ratio (`- 1.0` and `- 0.0` are exact). The push
// flag of the NEXT frame is a function of this frame's s alone, because fl() is monotone:
//   p' = [fl(s + ratio) >= 1] = [s >= thrA]            if s < 1   (p = 0)
//   p' = [fl((s - 1) + ratio) >= 1] = [s >= thrB]      if s >= 1  (p = 1)
// with thrA < 1 < thrB the smallest such floats, i.e. p' = [thrA <= s < 1] + [s >= thrB]; for ratio < 0.499 the second
// term never fires (s < 1 + ratio < thrB). The flag is computed one frame ahead, off the s -> (s - p) -> (+ ratio)
// chain (tests/test_phase_lookahead.py). The first frame of a run is taken literally: right after a ratio change s
// may still exceed 1 + ratio.
//
// ratio >= 1 (cubic.rs:94-105): `while sub_pos < ratio { push; sub_pos += 1.0 }; sub_pos -= ratio`. With sub_pos in
// [0,1) at frame start the trip count is n0 = floor(ratio) or n0 + 1. Repeated f32 `+= 1.0` only rounds when the sum
// enters a new binade ([1,2), [2,4), [4,8), ...) and is exact inside one, so t_n = sub_pos after n pushes has the
// closed form t_n = fl(fl(fl(fl(s + a1) + a2) + a3) + a4), a = (min(n,1), min(n-1,2), min(n-3,4), min(n-7,8))
// (bit-identical to the sequential adds; tests/test_phase_closed_form.py). Integer ratios (the one case where
// t_{n0-1} can reach `ratio`) and ratio >= 14 take the literal loop.
//
// Pushes are not counted per frame: s_out = s_in -/+ (pushes - span * ratio) up to the recurrence's own rounding
// errors (< 1e-3 per piece), so the count is the nearest integer of that difference in f64.
struct PhaseK {
  float ratio, w1, wA, wB, wR;
  float a2, a3, a4;
  int mode;  // 0: ratio < 0.499   1: ratio < 1   2: ratio >= 1, closed form   3: literal loop
  int nm;    // floor(ratio) - 1
};

PB_DEV PhaseK phase_consts(const float ratio) {
  PhaseK k;
  k.ratio = ratio;
  k.w1 = step_weight(1.0f);
  k.wA = k.wB = k.wR = 0.0f; k.a2 = k.a3 = k.a4 = 0.0f; k.nm = 0;
  if (ratio < 1.0f) {
    float thrA = 1.0f - ratio;
    while (f32_pred(thrA) + ratio >= 1.0f) thrA = f32_pred(thrA);
    while (thrA + ratio < 1.0f) thrA = f32_succ(thrA);
    k.wA = step_weight(thrA);
    k.mode = 0;
    if (!(ratio < 0.499f)) {
      float thrB = 2.0f - ratio;
      while ((f32_pred(thrB) - 1.0f) + ratio >= 1.0f) thrB = f32_pred(thrB);
      while ((thrB - 1.0f) + ratio < 1.0f) thrB = f32_succ(thrB);
      k.wB = step_weight(thrB);
      k.mode = 1;
    }
  } else {
    const int n0 = (int)ratio;
    k.nm = n0 - 1;
    k.a2 = (float)min(max(k.nm - 1, 0), 2); k.a3 = (float)min(max(k.nm - 3, 0), 4); k.a4 = (float)min(max(k.nm - 7, 0), 8);
    k.wR = ratio * STEP_K;
    k.mode = (ratio < 14.0f && ratio != (float)n0) ? 2 : 3;
  }
  return k;
}

// `span` frames. `first`: this is the first piece of a run with this ratio (p is not valid yet).
// ACC additionally runs an independent `o += d` chain (an envelope stage's bare accumulate) in the latency shadow.
template <bool ACC>
PB_DEV uint32_t phase_piece(float& s, float& p, const PhaseK& k, const uint32_t span, const bool first, float& o, const float d) {
  if (span == 0) return 0u;
  const float ratio = k.ratio;
  const float s_in = s;
  uint32_t f = 0;
  if (k.mode <= 1) {
    if (first) {  // frame 0 and the flag of frame 1, literally
      p = step(s, k.w1);
      s = (s - p) + ratio;
      if (ACC) o += d;
      p = step(s, k.w1);
      f = 1;
    }
    const float w1 = k.w1, wA = k.wA;
    if (k.mode == 0) {
#pragma unroll 4
      for (; f < span; ++f) {
        const float pn = step(s, wA) - step(s, w1);
        s = (s - p) + ratio;
        p = pn;
        if (ACC) o += d;
      }
    } else {
      const float wB = k.wB;
#pragma unroll 4
      for (; f < span; ++f) {
        const float pn = (step(s, wA) - step(s, w1)) + step(s, wB);
        s = (s - p) + ratio;
        p = pn;
        if (ACC) o += d;
      }
    }
    return (uint32_t)__double2int_rn(((double)s_in - (double)s) + (double)span * (double)ratio);
  }
  if (k.mode == 2) {
    if (first && !(s < 1.0f)) {  // sub_pos may be >= 1 right after a ratio change: one literal frame
      while (s < ratio) s += 1.0f;
      s -= ratio;
      if (ACC) o += d;
      f = 1;
    }
    // from here on sub_pos = t - ratio with t in [ratio, ratio + 1): always in [0, 1)
    const float wR = k.wR;
#define PB_PHASE_LOOP(TM_EXPR)                                       \
    for (; f < span; ++f) {                                      \
      const float tm = (TM_EXPR);                                \
      const float t0 = tm + 1.0f;                                \
      const float more = fma_sat(t0, -STEP_K, wR); /* [t0 < ratio] */ \
      const float t = t0 + more; /* + 0.0 is exact */            \
      s = t - ratio;                                             \
      if (ACC) o += d;                                           \
    }
    const int nm = k.nm;
    if (nm <= 0) { PB_PHASE_LOOP(s) }
    else if (nm == 1) { PB_PHASE_LOOP(s + 1.0f) }
    else if (nm <= 3) { PB_PHASE_LOOP((s + 1.0f) + k.a2) }
    else if (nm <= 7) { PB_PHASE_LOOP(((s + 1.0f) + 2.0f) + k.a3) }
    else { PB_PHASE_LOOP((((s + 1.0f) + 2.0f) + 4.0f) + k.a4) }
#undef PB_PHASE_LOOP
    return (uint32_t)__double2int_rn(((double)s - (double)s_in) + (double)span * (double)ratio);
  }
  uint32_t np = 0;
  for (; f < span; ++f) {
    while (s < ratio) { s += 1.0f; ++np; }
    s -= ratio;
    if (ACC) o += d;
  }
  return np;
}

// The same piece with ONE instruction stream for modes 0-2, for the lane-per-voice skeleton (a warp = the voices of a
// group, each with its own ratio): every lane evaluates both the ratio < 1 step and the closed-form ratio >= 1 step and
// keeps its own, so the lanes of a warp stay converged whatever their ratios are. Constants of the arm a lane does not
// use are neutral (`x + 0.0f` is exact for x >= 0; a step weight of -inf never fires). The envelope accumulate always
// runs (d = 0 when not fused; the caller then discards o). Bit-identical to phase_piece<ACC> per lane.
PB_DEV uint32_t phase_piece_uniform(float& s, float& p, const PhaseK& k, const uint32_t span, const bool first, float& o, const float d) {
  if (span == 0) return 0u;
  if (k.mode == 3) return phase_piece<true>(s, p, k, span, first, o, d);
  const float ratio = k.ratio;
  const float s_in = s;
  const bool down = k.mode <= 1;
  uint32_t f = 0;
  if (first) {
    if (down) {
      p = step(s, k.w1);
      s = (s - p) + ratio;
      o += d;
      p = step(s, k.w1);
      f = 1;
    } else if (!(s < 1.0f)) {
      while (s < ratio) s += 1.0f;
      s -= ratio;
      o += d;
      f = 1;
    }
  }
  const float w1 = k.w1, wA = k.wA, wB = k.mode == 1 ? k.wB : -__int_as_float(0x7f800000), wR = k.wR;
  const float nmf = (float)max(k.nm, 0);
  const float a1 = down ? 0.0f : fminf(nmf, 1.0f), a2 = down ? 0.0f : k.a2, a3 = down ? 0.0f : k.a3, a4 = down ? 0.0f : k.a4;
#pragma unroll 2
  for (; f < span; ++f) {
    const float pn = (step(s, wA) - step(s, w1)) + step(s, wB);
    const float sd = (s - p) + ratio;
    const float tm = (((s + a1) + a2) + a3) + a4;
    const float t0 = tm + 1.0f;
    const float more = fma_sat(t0, -STEP_K, wR);  // [t0 < ratio]
    const float su = (t0 + more) - ratio;
    s = down ? sd : su;
    p = pn;
    o += d;
  }
  const double bal = down ? ((double)s_in - (double)s) : ((double)s - (double)s_in);
  return (uint32_t)__double2int_rn(bal + (double)span * (double)ratio);
}

// One piece on its own (the ratio may change from piece to piece: pitch glides)
PB_DEV uint32_t phase_run(float& s, const float ratio, const uint32_t span) {
  float o = 0.0f, p = 0.0f;
  const PhaseK k = phase_consts(ratio);
  return phase_piece<false>(s, p, k, span, true, o, 0.0f);
}

// How many frames the current envelope stage can run as a bare chain of the reference's own f32 accumulate
// before a threshold crossing comes within reach (10 % + 2 steps of slack), the per-frame increment `d`
// (`x -= step` == `x += -step` exactly) and whether the chain runs on env_hold instead of env_out.
PB_DEV uint32_t env_bare_steps(const VoiceState& v, const GroupParams& gp, float& d, bool& on_hold) {
  const uint32_t stage = v.env_stage;
  float room = 0.0f, step = 1.0f;
  on_hold = false;
  d = 0.0f;
  if (stage == ENV_ATTACK) { room = v.env_target - v.env_out; step = gp.attack_rate; d = step; }
  else if (stage == ENV_HOLD) { room = v.env_hold; step = 1.0f; d = -1.0f; on_hold = true; }
  else if (stage == ENV_DECAY && v.env_out > gp.sustain_level) { room = v.env_out - gp.sustain_level; step = gp.decay_rate; d = -step; }
  else if (stage == ENV_RELEASE) { room = v.env_out - 0.001f; step = v.env_release_out * gp.release_rate; d = -step; }
  if (room > 0.0f && step > 0.0f) {
    const float q = fminf(room / step * 0.9f, 1.0e6f);
    return q > 3.0f ? (uint32_t)q - 2u : 0u;
  }
  return 0u;
}

// The AHDSR stage machine advanced by `w` frames, state only (ahdsr.rs:448-516)
PB_DEV void env_chain(VoiceState& v, const GroupParams& gp, const uint32_t w) {
  uint32_t i = 0;
  while (i < w) {
    const uint32_t stage = v.env_stage;
    if (stage == ENV_SUSTAIN || stage == ENV_IDLE) break;
    float d;
    bool on_hold;
    const uint32_t m = min(env_bare_steps(v, gp, d, on_hold), w - i);
    if (m) {
      float o = on_hold ? v.env_hold : v.env_out;
#pragma unroll 8
      for (uint32_t j = 0; j < m; ++j) o += d;
      if (on_hold) v.env_hold = o; else v.env_out = o;
      i += m;
    } else {
      (void)env_run(v, gp);
      ++i;
    }
  }
}

// The fader / gain / pan recurrences of `w` frames, state only (fader.rs:109-116, smoothing.rs:61-64,74-122)
PB_DEV void advance_ramps(VoiceState& v, const CallCtx& c, const uint32_t w, const float comp) {
  // (2) VolumeFader ramp (fader.rs:109-116)
  if (c.fader_running) {
    float cur = v.fader_cur;
    const float tgt = v.fader_tgt, inertia = v.fader_inertia;
    for (uint32_t i = 0; i < w; ++i) cur += (tgt - cur) * inertia;
    v.fader_cur = cur;
  }
  // (3) gain ramp, two steps per frame (smoothing.rs:61-64); once it stops needing a ramp it stays put
  if (c.vol_ramp) {
    float cur = v.vol.current;
    const float tgt = v.vol.target;
    for (uint32_t i = 0; i < 2 * w; ++i) {
      const float add = (tgt - cur) * SMOOTH_INERTIA * comp;
      if (!(fabsf(add) > F32_EPS * 100.0f)) break;
      cur += add;
    }
    v.vol.current = cur;
  }
  // (4) pan ramp, one step per frame
  if (c.pan_ramp) {
    float cur = v.pan.current;
    const float tgt = v.pan.target;
    for (uint32_t i = 0; i < w; ++i) {
      const float add = (tgt - cur) * SMOOTH_INERTIA * comp;
      if (!(fabsf(add) > F32_EPS * 100.0f)) break;
      cur += add;
    }
    v.pan.current = cur;
  }
}

// ---- skeleton fast path -------------------------------------------------------------------------------------
// State-only advance of one write call by up to `n` frames, bit-identical to voice_frames<CC,false> but with
// the independent recurrences separated into tight loops: (1) the resampler's f32 phase/position
// recurrence (cubic.rs:72-110), then -- only while they are actually moving -- (2) the fader, (3) the
// per-sample gain ramp, (4) the pan ramp, (5) the AHDSR stage machine. Falls back to the general per-frame
// code near loop ends / EOF, where input exhaustion has to be handled frame by frame.
template <int CC>
PB_DEV uint32_t voice_advance(VoiceState& v, CallCtx& c, const GroupParams& gp, const DevBuffer& b, uint32_t out_rate,
                              float comp, uint32_t n) {
  HistVals hv_unused;
  uint32_t done = 0;
  while (done < n && !c.ended) {
    if (c.call_left == 0) {  // next write_buffer call (preloaded.rs:419-447)
      if (c.gliding) {
        if (v.to_next_speed_update == 0) {
          if (v.current_speed != v.target_speed) update_speed(v, b.sample_rate, out_rate);
          v.to_next_speed_update = 64;
        }
        c.call_left = min(c.chunk_left, v.to_next_speed_update);
      } else {
        c.call_left = c.chunk_left;
      }
      loop_range_samples(v, b, c.ls, c.le);
      c.new_call = true;
    }
    const uint32_t span = min(n - done, c.call_left);
    const float ratio = v.ratio;
    const bool bypass = fabsf(ratio - 1.0f) < 0.000001f;
    const uint32_t avail = c.le > v.playback_pos ? (c.le - v.playback_pos) / CC : 0u;
    // worst-case pushes of `span` frames (+3 for a pending preload); exhaustion cannot happen below that
    const uint32_t per_frame = bypass ? 1u : (ratio < 1.0f ? 1u : (uint32_t)ratio + 2u);
    const bool fast = !bypass && ratio > 0.0f && ratio < 64.0f && (uint64_t)span * per_frame + 4u < (uint64_t)avail;
    uint32_t w;
    if (fast) {
      if (c.new_call) {  // CubicInterpolator::process prologue (cubic.rs:47-69)
        c.new_call = false;
        c.produced_in_call = 0;
        if (!v.initialized) {
          v.initialized = 1;
          v.hidx[3] = v.hidx[0];
          v.hidx[2] = (int32_t)v.playback_pos; v.hidx[1] = (int32_t)(v.playback_pos + CC); v.hidx[0] = (int32_t)(v.playback_pos + 2 * CC);
          v.playback_pos += 3 * CC;
        }
      }
      float s = v.sub_pos;
      const uint32_t np = phase_run(s, ratio, span);
      v.sub_pos = s;
      if (np >= 4) {
        v.playback_pos += np * CC;
        v.hidx[0] = (int32_t)(v.playback_pos - CC); v.hidx[1] = (int32_t)(v.playback_pos - 2 * CC);
        v.hidx[2] = (int32_t)(v.playback_pos - 3 * CC); v.hidx[3] = (int32_t)(v.playback_pos - 4 * CC);
      } else {
        for (uint32_t i = 0; i < np; ++i) {
          v.hidx[3] = v.hidx[2]; v.hidx[2] = v.hidx[1]; v.hidx[1] = v.hidx[0]; v.hidx[0] = (int32_t)v.playback_pos;
          v.playback_pos += CC;
        }
      }
      c.produced_in_call += span;
      c.call_left -= span;
      c.chunk_left -= span;
      if (c.gliding) v.to_next_speed_update -= span;
      if (c.call_left == 0) after_process_call(v, c);
      w = span;
    } else if (bypass && span > 0u && span <= avail) {
      // equal rates (cubic.rs:53-58): a plain copy, one input frame per output frame and no history; with `span` frames
      // of input left nothing else can happen
      if (c.new_call) { c.new_call = false; c.produced_in_call = 0; }
      v.playback_pos += span * CC;
      c.produced_in_call += span;
      c.call_left -= span;
      c.chunk_left -= span;
      if (c.gliding) v.to_next_speed_update -= span;
      if (c.call_left == 0) after_process_call(v, c);
      w = span;
    } else {
      // general path for the resampler only: switch the other recurrences off, they are advanced below
      CallCtx t = c;
      t.fader_running = false; t.fader_scale = false; t.vol_ramp = false; t.vol_scale = false;
      t.pan_ramp = false; t.pan_apply = false; t.env_per_frame = false;
      GroupParams gq = gp;
      gq.has_env = 0;
      w = voice_frames<CC, false>(v, t, hv_unused, gq, b, out_rate, comp, span, nullptr);
      c.chunk_left = t.chunk_left; c.call_left = t.call_left; c.produced_in_call = t.produced_in_call;
      c.ls = t.ls; c.le = t.le; c.new_call = t.new_call; c.ended = t.ended;
    }
    advance_ramps(v, c, w, comp);
    // (5) AHDSR (ahdsr.rs:448-516); Sustain and Idle do not move
    if (gp.has_env && c.env_per_frame) env_chain(v, gp, w);
    done += w;
    if (w < span) break;
  }
  return done;
}

// Snapshot the skeleton pass emits at the start of every replay segment: the complete voice state and
// write-call context, from which a replay thread reproduces the next `n` frames bit for bit.
struct Segment {
  VoiceState v;
  CallCtx c;
  uint32_t out_off;  // first frame, relative to the time block
  uint32_t n;        // frames requested (the replay may produce fewer when the source runs dry)
  uint32_t gp_idx;   // the group's parameter version in force (GroupParams array index)
  uint32_t _pad;
};

// Generator-level gain/pan checkpoint (AmplifiedSource/PannedSource around a Sampler, player.rs:1075-1081)
struct GroupSeg {
  ExpSm vol, pan;
  uint32_t out_off, n;
  uint32_t flags;  // bit0 vol_ramp, bit1 vol_scale, bit2 pan_ramp, bit3 pan_apply
  uint32_t _pad;
};


// ---- simple calls -----------------------------------------------------------------------------------------
// A write call is "simple" when nothing but the phase recurrence and the envelope can move during it: no
// glide (one write_buffer call, constant ratio), no fader / gain / pan ramp, and enough input left that
// neither a loop wrap nor EOF can be reached. The skeleton then emits one full Segment at the call's first
// frame and, at every later 64-frame tile boundary inside the call, only this 32-byte record; the replay
// rebuilds the complete state from the call's Segment + the record (apply_tile_rec).
struct __align__(16) TileRec {
  uint32_t pos;        // playback_pos at the tile boundary
  float sub_pos;
  float env_out, env_hold, env_target;
  uint32_t stage_n;    // env_stage << 16 | frames of this piece
  uint32_t base;       // index of the call's Segment
  uint32_t gen;        // time-block generation tag (stale records of earlier blocks are ignored)
};

template <int CC>
PB_DEV bool simple_call_ok(const VoiceState& v, const CallCtx& c, const DevBuffer& b, uint32_t n) {
  if (c.gliding || c.fader_running || c.vol_ramp || c.pan_ramp) return false;
  const float ratio = v.ratio;
  if (fabsf(ratio - 1.0f) < 0.000001f || !(ratio > 0.0f) || !(ratio < 64.0f)) return false;
  if (ratio >= 1.0f && !(ratio < 14.0f && ratio != (float)(int)ratio && v.sub_pos < 1.0f)) return false;
  uint32_t ls, le;
  loop_range_samples(v, b, ls, le);
  const uint32_t avail = le > v.playback_pos ? (le - v.playback_pos) / CC : 0u;
  const uint32_t per_frame = ratio < 1.0f ? 1u : (uint32_t)ratio + 2u;
  return (uint64_t)n * per_frame + 4u < (uint64_t)avail;
}

// Rebuild the voice state `frames_in` frames into a simple call from the state at the call's first frame.
template <int CC>
PB_DEV void apply_tile_rec(VoiceState& v, CallCtx& c, const DevBuffer& b, const TileRec& r, uint32_t frames_in) {
  // start of the (single) write_buffer call + CubicInterpolator::process prologue (preloaded.rs:419-447, cubic.rs:47-69)
  c.call_left = c.chunk_left;
  loop_range_samples(v, b, c.ls, c.le);
  c.new_call = false;
  // all pushes since the call started (3 preload frames included) took consecutive input frames
  const uint32_t k = (r.pos - v.playback_pos) / CC;
  int32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // v.hidx[i - k] as a select chain: a dynamic index would force the whole VoiceState into local memory
    const uint32_t j = (uint32_t)i - min(k, (uint32_t)i);
    const int32_t old = j == 0 ? v.hidx[0] : j == 1 ? v.hidx[1] : j == 2 ? v.hidx[2] : v.hidx[3];
    h[i] = (uint32_t)i < k ? (int32_t)(r.pos - (uint32_t)(i + 1) * CC) : old;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) v.hidx[i] = h[i];
  v.initialized = 1;
  v.playback_pos = r.pos;
  v.sub_pos = r.sub_pos;
  c.produced_in_call = frames_in;
  c.call_left -= frames_in;
  c.chunk_left -= frames_in;
  v.env_out = r.env_out; v.env_hold = r.env_hold; v.env_target = r.env_target;
  v.env_stage = r.stage_n >> 16;
}

// ---- replay fast path ------------------------------------------------------------------------------------
// Is the Segment at hand the first frame of a call the skeleton treated as simple? (Same predicate on the same
// state, so the answer is the skeleton's; a call it had to treat as general for bookkeeping reasons only -- segment
// capacity -- is still rendered correctly by either path.)
template <int CC>
PB_DEV bool simple_call_start(const VoiceState& v, const CallCtx& c, const DevBuffer& b) {
  return c.new_call && c.call_left == 0 && c.produced_in_call == 0 && !c.ended && !v.hq && simple_call_ok<CC>(v, c, b, c.chunk_left);
}

// CubicInterpolator::process prologue of a simple call (cubic.rs:60-69), as simple_call() does it in the skeleton
template <int CC>
PB_DEV void simple_call_prologue(VoiceState& v, CallCtx& c, HistVals& hv, const DevBuffer& b) {
  c.call_left = c.chunk_left;
  loop_range_samples(v, b, c.ls, c.le);
  c.new_call = false;
  if (!v.initialized) {
    v.initialized = 1;
#pragma unroll
    for (int f = 0; f < 3; ++f) push_frame<CC, true>(v, hv, b.data);
  }
}

// `n` frames inside a simple call, bit-identical to voice_frames<CC, true> under the simple-call preconditions (constant
// ratio != 1, one write_buffer call, no fader / gain / pan ramp, the input cannot run out): only the phase recurrence,
// the history pushes, the Hermite evaluation, the constant scale factors (`x * 1.0f` is exact, so an absent stage is a
// multiplication by one) and the envelope remain. Call bookkeeping (call_left, produced_in_call, history indices) is
// not maintained: the replay reloads it with the next Segment / TileRec.
constexpr uint32_t SIMPLE_WIN = 16;  // input samples staged per refill

template <int CC>
PB_DEV void simple_frames(VoiceState& v, const CallCtx& c, HistVals& hv, const GroupParams& gp, const DevBuffer& b,
                          const uint32_t n, float* __restrict__ out, const bool acc, float* __restrict__ win, const uint32_t wstride) {
  const float* __restrict__ buf = b.data;
  const float ratio = v.ratio;
  float s = v.sub_pos;
  uint32_t pos = v.playback_pos;
  const float fs = c.fader_scale ? v.fader_tgt : 1.0f;
  const float vs = c.vol_scale ? v.vol.target : 1.0f;
  const float pl = c.pan_apply ? c.pan_l : 1.0f, pr = c.pan_apply ? c.pan_r : 1.0f;
  const bool env_pf = gp.has_env && c.env_per_frame;
  const float ec = gp.has_env ? c.env_const : 1.0f;
  const bool down = ratio < 1.0f;
  // The pushes consume consecutive input samples: they are staged SIMPLE_WIN at a time in the thread's window of shared
  // memory (element i of thread t at win[i * stride], one bank per lane), so the L2 / HBM latency of the sample
  // buffer is paid once per window by independent loads in flight together instead of once per push by a load the
  // Hermite evaluation waits for. Reads past the last sample a push can reach are clamped into the buffer and never used.
  const uint32_t last = b.n_samples - 1u;
  uint32_t wbase = pos;
  auto refill = [&]() {
#pragma unroll
    for (uint32_t i = 0; i < SIMPLE_WIN; ++i) win[i * wstride] = __ldg(buf + min(wbase + i, last));
  };
  refill();
  auto push = [&]() {
    if (pos - wbase >= SIMPLE_WIN) { wbase = pos; refill(); }
    hist_push(hv.h[0], win[(pos - wbase) * wstride]);
    if (CC == 2) hist_push(hv.h[1], win[(pos - wbase + 1u) * wstride]);
    pos += CC;
  };
  for (uint32_t f = 0; f < n; ++f) {
    float fr;
    if (down) {  // cubic.rs:72-90
      if (s >= 1.0f) { push(); s -= 1.0f; }
      fr = s;
      s += ratio;
    } else {     // cubic.rs:91-110
      while (s < ratio) { push(); s += 1.0f; }
      s -= ratio;
      fr = 1.0f - s;
    }
    float x0 = hermite(hv.h[0], fr);
    float x1 = CC == 2 ? hermite(hv.h[1], fr) : x0;
    x0 *= fs; x1 *= fs;                       // VolumeFader (fader.rs:103-116), not running
    float l = x0 * vs, r = x1 * vs;           // AmplifiedSource (smoothing.rs:60-71), not ramping
    l *= pl; r *= pr;                         // PannedSource, not ramping
    const float e = env_pf ? env_run(v, gp) : ec;  // AHDSR (voice.rs:470-486)
    l *= e; r *= e;
    out[2 * f] = acc ? out[2 * f] + l : l;
    out[2 * f + 1] = acc ? out[2 * f + 1] + r : r;
  }
  v.sub_pos = s;
  v.playback_pos = pos;
}


}  // namespace pb
