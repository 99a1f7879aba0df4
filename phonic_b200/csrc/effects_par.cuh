// Chunk-parallel forms of the stereo-coupled effects. A bus effect has two channels, so running it
// frame-serially leaves the GPU idle and pays a full memory round trip per delay-line access
// (cfg1's reverb took 4 s on one thread vs 0.1 s on one CPU core). Every feedback path of
// these effects goes through a delay line whose read lags its write by at least `lag` frames; inside a
// sub-block shorter than the lag every delay-line read sees data written before the sub-block, so all
// frames of the sub-block can be evaluated in parallel stage by stage. What stays serial is only what the
// reference makes serial by construction: f32/f64 phase accumulators and the compressor's branchy
// one-pole follower -- bare dependent chains of 1-4 ops per frame, run by single threads between the
// parallel stages. Per-frame arithmetic is the reference's own, in its order, so results equal the serial
// path bit for bit except where a time-invariant SVF/biquad is evaluated by the f64 block scan.
//
// Each function is called by all threads of the mixer CTA and returns false (uniformly) when the fast
// form does not apply (ramping parameters, lag shorter than a sub-block, limiter mode, ...); the caller
// then runs the frame-serial reference form.
#pragma once
#include "effects.cuh"

namespace pb {

constexpr uint32_t FX_THREADS_C = 256;  // threads of the mixer CTA (== FX_THREADS)

struct ParWork {
  uint8_t* base;    // shared-memory work area
  uint32_t bytes;
};

template <class T>
PB_DEV T* carve(uint8_t*& p, uint32_t n) {
  T* r = reinterpret_cast<T*>(p);
  p += ((n * sizeof(T) + 15u) & ~15u);
  return r;
}

// ---- CompressorEffect (compressor.rs:230-294), compressor mode (ratio < 20) ------------------------------
PB_DEV bool comp_parallel(CompState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t n, uint32_t tid, uint32_t nt, ParWork w) {
  const bool limiter = s.ratio >= 20.0f;  // detector = running peak of the look-ahead window (LookupDelayLine, delay.rs:206-265)
  if (limiter && (s.delay_frames == 0 || (size_t)5 * n * sizeof(float) + (size_t)s.buf_frames * 2 * sizeof(float) > w.bytes)) return false;
  uint8_t* p = w.base;
  float* in_db = carve<float>(p, n);
  float* env = carve<float>(p, n);
  float* dl = carve<float>(p, n);
  float* dr = carve<float>(p, n);
  float* mk = carve<float>(p, n);
  double* line = cx.aux_arena + s.aux;
  const uint32_t D = s.delay_frames, wp = s.write_pos, mask = s.mask, bufn = s.buf_frames;
  // 1. delayed frames (LookupDelayLine::process reads, delay.rs:214-222) and detector level of the undelayed input
  for (uint32_t f = tid; f < n; f += nt) {
    const float in0 = CB_L(f), in1 = CB_R(f);
    float d0 = in0, d1 = in1;
    if (D != 0) {
      if (f >= D) { d0 = CB_L(f - D); d1 = CB_R(f - D); }
      else { const uint32_t ri = (wp + f + bufn - D) & mask; d0 = (float)line[ri * 2]; d1 = (float)line[ri * 2 + 1]; }
    }
    dl[f] = d0; dr[f] = d1;
    const float peak = fmaxf(fabsf(in0), fabsf(in1));
    in_db[f] = peak > 1e-6f ? 20.0f * log10f(peak) : -120.0f;
  }
  if (limiter) {
    // The window peak is a serial state machine over the delay line (new maximum / expiry of the old one -> rescan).
    // The line holds f32 inputs widened to f64, so a float copy in shared memory is lossless; warp 0 walks the chunk
    // on that copy, every lane holding the same scalar state, and the lanes share the rare O(delay) rescans.
    float* ring = carve<float>(p, (size_t)bufn * 2);
    for (uint32_t i = tid; i < bufn * 2; i += nt) ring[i] = (float)line[i];
    __syncthreads();
    if (tid < 32) {
      const uint32_t lane = tid;
      float peak = (float)s.peak_value;
      uint32_t pos = s.peak_pos;
      auto rescan = [&](const uint32_t newest, const uint32_t first_i, const uint32_t last_i) {
        // `for i in first..last { fi = (newest - i) & mask; if fp >= peak { peak = fp; pos = fi } }`: the LAST i holding the maximum wins
        float bv = 0.0f;
        uint32_t bi = 0xFFFFFFFFu;
        for (uint32_t i = first_i + lane; i < last_i; i += 32) {
          const uint32_t fi = (newest + bufn - i) & mask;
          const float fp = fmaxf(fmaxf(0.0f, fabsf(ring[fi * 2])), fabsf(ring[fi * 2 + 1]));
          if (fp >= bv) { bv = fp; bi = i; }
        }
        for (uint32_t o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
          const uint32_t oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
          if (oi != 0xFFFFFFFFu && (bi == 0xFFFFFFFFu || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; }
        }
        peak = 0.0f;
        if (bi != 0xFFFFFFFFu) { peak = bv; pos = (newest + bufn - bi) & mask; }
      };
      if (s.peak_dirty) { rescan(wp, 1, D + 1); }  // compressor chunks do not track it (effects.cuh comp_process does the same rescan)
      for (uint32_t f = 0; f < n; ++f) {
        const float in0 = CB_L(f), in1 = CB_R(f);
        const uint32_t wi = (wp + f) & mask, read_index = (wp + f + bufn - D) & mask;
        if (lane == 0) { ring[wi * 2] = in0; ring[wi * 2 + 1] = in1; }
        const bool expired = pos == read_index;
        const float new_peak = fmaxf(fmaxf(0.0f, fabsf(in0)), fabsf(in1));
        if (new_peak >= peak) { peak = new_peak; pos = wi; }
        else if (expired) { __syncwarp(); rescan(wi, 0, D); }
        if (lane == 0) in_db[f] = peak;  // the window peak; converted to dB below
      }
      if (lane == 0) { s.peak_value = (double)peak; s.peak_pos = pos; }
    }
    __syncthreads();
    for (uint32_t f = tid; f < n; f += nt) { const float peak = in_db[f]; in_db[f] = peak > 1e-6f ? 20.0f * log10f(peak) : -120.0f; }
  }
  __syncthreads();
  // 2. serial parts: EnvelopeFollower::run (envelope.rs:51-60) and the makeup-gain smoother
  if (tid == 0) {
    float cur = s.env_cur;
    const float atk = s.atk_coeff, rel = s.rel_coeff;
    uint32_t f = 0;
    for (; f + 8 <= n; f += 8) {  // loads and stores of 8 frames around the dependent chain
      float x[8], e[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = in_db[f + j];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float c = x[j] > cur ? atk : rel; cur = x[j] + c * (cur - x[j]); e[j] = cur; }
#pragma unroll
      for (int j = 0; j < 8; ++j) env[f + j] = e[j];
    }
    for (; f < n; ++f) {
      const float x = in_db[f];
      const float c = x > cur ? atk : rel;
      cur = x + c * (cur - x);
      env[f] = cur;
    }
    s.env_cur = cur;
  } else if (tid == 32) {
    ExpSm m = s.makeup;
    if (exp_need_ramp(m, cx.comp)) { for (uint32_t f = 0; f < n; ++f) mk[f] = exp_next(m, cx.comp); }
    else { for (uint32_t f = 0; f < n; ++f) mk[f] = m.target; }
    s.makeup = m;
  }
  // delay line writes (after every read of this chunk has been taken)
  if (D != 0)
    for (uint32_t f = tid; f < n; f += nt) {
      const uint32_t wi = (wp + f) & mask;
      line[wi * 2] = (double)CB_L(f);
      line[wi * 2 + 1] = (double)CB_R(f);
    }
  __syncthreads();
  // 3. gain computer + output (compressor.rs:262-291)
  const float t = s.threshold, wk = s.knee;
  const float slope = limiter ? 1.0f : 1.0f - 1.0f / s.ratio;
  for (uint32_t f = tid; f < n; f += nt) {
    const float envelope = env[f];
    float gr_db;
    if (wk > 0.0f && envelope > (t - wk / 2.0f) && envelope < (t + wk / 2.0f)) {
      const float knee_lower = t - wk / 2.0f;
      const float x = (envelope - knee_lower) / wk;
      gr_db = x * x * slope * wk / 2.0f;
    } else if (envelope > (t + wk / 2.0f)) {
      gr_db = (envelope - t) * slope;
    } else {
      gr_db = 0.0f;
    }
    const float total_gain = db_to_linear_dev(mk[f] - gr_db);
    CB_L(f) = dl[f] * total_gain;
    CB_R(f) = dr[f] * total_gain;
  }
  if (tid == 0 && D != 0) { s.write_pos = (wp + n) & mask; s.peak_dirty = limiter ? 0u : 1u; }
  __syncthreads();
  return true;
}

// SvfFilter::process_sample (svf.rs:211-222) as a (a1,a2,a3,m0,m1,m2) tick: LP = v2, BP = v1, HP = in - k v1 - v2
PB_DEV BiquadCoef svf_as_biquad(const SvfCoef& c) {
  BiquadCoef b;
  b.a1 = c.a1; b.a2 = c.a2; b.a3 = c.a3;
  if (c.type == 0) { b.m0 = 0.0; b.m1 = 0.0; b.m2 = 1.0; }
  else if (c.type == 2) { b.m0 = 0.0; b.m1 = 1.0; b.m2 = 0.0; }
  else { b.m0 = 1.0; b.m1 = -c.k; b.m2 = -1.0; }
  b.cutoff = c.cutoff; b.q = 0.0f; b.gain = 0.0f; b.type = 0; b.sample_rate = c.sample_rate; b._pad = 0;
  return b;
}

// DcFilter::process_sample (dc.rs:84-88) over a block by one warp. y[n] = (v[n] - v[n-1]) + r y[n-1] is linear in y:
// every lane runs its sub-block once from a zero state, the sub-block start states are chained through r^len (32
// steps, every lane redundantly), then every lane re-runs the reference's own recurrence from its true start state.
// Exact in real arithmetic; in f64 the start states differ from the serial chain by O(1e-16) relative -- used only where
// the parity bar is a tolerance (the Delay's feedback path); GainEffect's bit-exact DC filter stays serial.
PB_DEV void dc_scan_channel(const double r, double& x1, double& y1, const double* __restrict__ v, float* __restrict__ out,
                            const uint32_t len, const uint32_t lane) {
  const uint32_t B = (len + 31u) / 32u;
  const uint32_t lo = min(lane * B, len), hi = min(lo + B, len);
  const double before = lo == 0 ? x1 : v[pidx(lo - 1)];
  double z = 0.0, rp = 1.0, prev = before;
  for (uint32_t n = lo; n < hi; ++n) {
    const double cur = v[pidx(n)];
    z = cur - prev + r * z;
    prev = cur;
    rp *= r;
  }
  double start = y1, mine = y1;
  for (uint32_t l = 0; l < 32; ++l) {
    const double zl = __shfl_sync(0xFFFFFFFFu, z, l), rl = __shfl_sync(0xFFFFFFFFu, rp, l);
    if (lane == l) mine = start;
    start = rl * start + zl;
  }
  double y = mine;
  prev = before;
  for (uint32_t n = lo; n < hi; ++n) {
    const double cur = v[pidx(n)];
    y = cur - prev + r * y;
    prev = cur;
    out[pidx(n)] = fminf(fmaxf((float)y, -4.0f), 4.0f);
  }
  const double last = len ? v[pidx(len - 1)] : x1;
  __syncwarp();
  if (len) { x1 = last; y1 = start; }
}

// Lfo (lfo.rs:122-169, 234-239) for one warp: only the phase accumulate is a serial f32 chain (lane 0, three dependent
// operations per frame); the waveform is a pure function of the phase and is evaluated by the 32 lanes in parallel.
PB_DEV void lfo_fill_warp(LfoSt& l, float* out, uint32_t n, uint32_t lane) {
  const uint32_t waveform = l.waveform;
  if (lane == 0) {
    float ph = l.phase;
    const float inc = l.phase_inc;
    for (uint32_t f = 0; f < n; ++f) {
      out[f] = ph;
      ph += inc;
      if (ph >= 1.0f) ph -= 1.0f;
    }
    l.phase = ph;
  }
  __syncwarp();
  for (uint32_t f = lane; f < n; f += 32) out[f] = lfo_wave(waveform, out[f]);
  __syncwarp();
}

// ---- ChorusEffect::process (chorus.rs:311-394) while no parameter is ramping ----------------------------------
PB_DEV bool chorus_parallel(ChorusState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t n, uint32_t tid, uint32_t nt, ParWork w) {
  if (lin_need_ramp(s.rate) || lin_need_ramp(s.phase) || exp_need_ramp(s.depth, cx.comp) || exp_need_ramp(s.feedback, cx.comp) ||
      exp_need_ramp(s.wet, cx.comp) || exp_need_ramp(s.filter_freq, cx.comp) || exp_need_ramp(s.filter_res, cx.comp) ||
      spring_need_ramp(s.delay))
    return false;
  const float srf = (float)cx.sample_rate;
  const float delay_ms = s.delay.target, depth = s.depth.target;
  const float fb = fminf(fmaxf(s.feedback.target, -0.999f), 0.999f);
  const float wet = s.wet.target, dry = 1.0f - wet;
  const float delay_in_samples = delay_ms * srf * 0.001f;
  const float depth_in_samples = s.lfo_range * depth;
  // every read position is >= 2 + delay_in_samples (1 + lfo >= 0): sub-blocks shorter than that never read their own writes
  const float pos_min = 2.0f + delay_in_samples;
  uint32_t L = 512;
  while (L >= 64 && (float)L + 2.0f >= pos_min) L >>= 1;
  if (L < 64) return false;
  uint8_t* p = w.base;
  float* llfo = carve<float>(p, n);
  float* rlfo = carve<float>(p, n);
  float* fl = carve<float>(p, PLANE);
  float* fr = carve<float>(p, PLANE);
  double* bl = cx.aux_arena + s.dl.aux;
  double* br = cx.aux_arena + s.dr.aux;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  // LFOs (two serial phase chains) next to the input SVF (two warps, block scan)
  for (uint32_t f = tid; f < n; f += nt) { fl[pidx(f)] = CB_L(f); fr[pidx(f)] = CB_R(f); }
  __syncthreads();
  if (warp < 2) {
    const BiquadCoef c = svf_as_biquad(s.coef);
    double ic1 = warp == 0 ? s.fl_ic1 : s.fr_ic1, ic2 = warp == 0 ? s.fl_ic2 : s.fr_ic2;
    biquad_scan_channel(c, ic1, ic2, warp == 0 ? fl : fr, cb.scratch + (size_t)warp * 1060, cb.lane_state + (size_t)warp * 64, n, lane);
    if (lane == 0) { if (warp == 0) { s.fl_ic1 = ic1; s.fl_ic2 = ic2; } else { s.fr_ic1 = ic1; s.fr_ic2 = ic2; } }
  } else if (warp == 2) {
    lfo_fill_warp(s.left_osc, llfo, n, lane);
  } else if (warp == 3) {
    lfo_fill_warp(s.right_osc, rlfo, n, lane);
  }
  __syncthreads();
  const uint32_t wpl = s.dl.write_pos, wpr = s.dr.write_pos, mkl = s.dl.mask, mkr = s.dr.mask;
  for (uint32_t f0 = 0; f0 < n; f0 += L) {
    const uint32_t f1 = min(f0 + L, n);
    for (uint32_t i = f0 * 2 + tid; i < f1 * 2; i += nt) {
      const uint32_t f = i >> 1, ch = i & 1;
      double* buf = ch ? br : bl;
      const uint32_t wp = ((ch ? wpr : wpl) + f) & (ch ? mkr : mkl), mk = ch ? mkr : mkl;
      const float lfo = ch ? rlfo[f] : llfo[f];
      const float pos = 2.0f + delay_in_samples + (1.0f + lfo) * depth_in_samples;
      // InterpolatedDelayLine::process (delay.rs:107-155)
      const double read_pos = (double)wp - (double)pos;
      const double flr = floor(read_pos);
      const double fraction = read_pos - flr;
      const long long index1 = (long long)flr;
      const uint32_t i1 = (uint32_t)((unsigned long long)index1 & mk), i2 = (uint32_t)((unsigned long long)(index1 + 1) & mk);
      const double v1 = buf[i1], v2 = buf[i2];
      const float out = (float)(v1 + (v2 - v1) * fraction);
      const float filtered = ch ? fr[pidx(f)] : fl[pidx(f)];
      buf[wp] = (double)filtered + (double)out * (double)fb;
      const float in = ch ? CB_R(f) : CB_L(f);
      const float o = in * dry + out * wet;
      if (ch) CB_R(f) = o; else CB_L(f) = o;
    }
    __syncthreads();
  }
  if (tid == 0) {
    s.dl.write_pos = (wpl + n) & mkl;
    s.dr.write_pos = (wpr + n) & mkr;
    const double PI = 3.14159265358979323846;
    double phase_inc = 2.0 * PI * (double)s.rate.current / (double)cx.sample_rate;
    s.current_phase += (double)(n * 2) / 2.0 * phase_inc;
    while (s.current_phase >= 2.0 * PI) s.current_phase -= 2.0 * PI;
  }
  __syncthreads();
  return true;
}

// ---- DelayEffect::process (delay.rs:334-454) while no parameter is ramping and the LFO does not move the filter ----
PB_DEV bool delay_parallel(DelayState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t n, uint32_t tid, uint32_t nt, ParWork w) {
  if (spring_need_ramp(s.delay_time) || exp_need_ramp(s.feedback, cx.comp) || exp_need_ramp(s.cutoff, cx.comp) ||
      exp_need_ramp(s.drive, cx.comp) || exp_need_ramp(s.wet, cx.comp) || exp_need_ramp(s.width, cx.comp) ||
      exp_need_ramp(s.lfo_rate, cx.comp) || exp_need_ramp(s.lfo_dt, cx.comp) || exp_need_ramp(s.lfo_dfb, cx.comp) ||
      exp_need_ramp(s.lfo_dflt, cx.comp))
    return false;
  if (s.lfo_dflt.target != 0.0f) return false;  // LFO -> filter: coefficients change every frame (serial form)
  const float srf = (float)cx.sample_rate;
  const float base_delay_ms = s.delay_time.target, dt_depth = s.lfo_dt.target;
  // shortest delay this chunk can see: |lfo| <= 1
  const float min_delay_samples = fmaxf(base_delay_ms - fabsf(dt_depth) * 50.0f, 1.0f) * 0.001f * srf;
  uint32_t L = 1024;
  while (L >= 32 && (float)L + 2.0f >= min_delay_samples) L >>= 1;
  if (L < 32) return false;
  {  // the (constant) filter coefficients the per-frame `set` call would leave in place (delay.rs:363-370)
    const float c = fminf(fmaxf(s.cutoff.target * powf(2.0f, 0.0f), 20.0f), srf / 2.0f);
    if (s.coef.type != s.filter_type || s.coef.sample_rate != cx.sample_rate || s.coef.cutoff != c || s.coef.resonance != 0.302f) {
      __syncthreads();
      if (tid == 0) svf_set(s.coef, s.filter_type, cx.sample_rate, c, 0.302f);
      __syncthreads();
    }
  }
  uint8_t* p = w.base;
  float* lfo = carve<float>(p, n);
  float* dly[2] = {carve<float>(p, PLANE), carve<float>(p, PLANE)};      // delayed, then clean (f32)
  double* flt[2] = {carve<double>(p, PLANE), carve<double>(p, PLANE)};   // filtered / saturated (f64)
  double* bufs[2] = {cx.aux_arena + s.dl.aux, cx.aux_arena + s.dr.aux};
  const uint32_t lane = tid & 31, warp = tid >> 5;
  if (warp == 0) lfo_fill_warp(s.lfo, lfo, n, lane);
  __syncthreads();
  const float base_fb = s.feedback.target, fb_depth = s.lfo_dfb.target, drive = s.drive.target;
  const float wet = s.wet.target, width = s.width.target;
  const uint32_t mode = s.mode;
  const uint32_t wp0[2] = {s.dl.write_pos, s.dr.write_pos}, mk[2] = {s.dl.mask, s.dr.mask};
  const BiquadCoef c = svf_as_biquad(s.coef);
  float fb_prev[2] = {s.fb_l, s.fb_r};  // feedback_left / feedback_right entering the sub-block
  for (uint32_t f0 = 0; f0 < n; f0 += L) {
    const uint32_t f1 = min(f0 + L, n), len = f1 - f0;
    // (1) delayed reads of both lines (written before this sub-block)
    for (uint32_t i = f0 * 2 + tid; i < f1 * 2; i += nt) {
      const uint32_t f = i >> 1, ch = i & 1;
      const float lv = lfo[f];
      const float delay_ms = fmaxf(base_delay_ms + lv * dt_depth * 50.0f, 1.0f);
      const float delay_samples = delay_ms * 0.001f * srf;
      const uint32_t wp = (wp0[ch] + f) & mk[ch];
      const double read_pos = (double)wp - (double)delay_samples;
      const double flr = floor(read_pos);
      const double fraction = read_pos - flr;
      const long long index1 = (long long)flr;
      const uint32_t i1 = (uint32_t)((unsigned long long)index1 & mk[ch]), i2 = (uint32_t)((unsigned long long)(index1 + 1) & mk[ch]);
      const double v1 = bufs[ch][i1], v2 = bufs[ch][i2];
      dly[ch][pidx(f - f0)] = (float)(v1 + (v2 - v1) * fraction);
    }
    __syncthreads();
    // (2) feedback path: SVF (block scan, f64 out) -> saturate -> DC blocker (serial chain) -> clamp
    if (warp < 2) {
      double ic1 = warp == 0 ? s.fl_ic1 : s.fr_ic1, ic2 = warp == 0 ? s.fl_ic2 : s.fr_ic2;
      biquad_scan_channel(c, ic1, ic2, dly[warp], cb.scratch + (size_t)warp * 1060, cb.lane_state + (size_t)warp * 64, len, lane, flt[warp]);
      if (lane == 0) { if (warp == 0) { s.fl_ic1 = ic1; s.fl_ic2 = ic2; } else { s.fr_ic1 = ic1; s.fr_ic2 = ic2; } }
    }
    __syncthreads();
    for (uint32_t i = tid; i < len * 2; i += nt) { const uint32_t f = i >> 1, ch = i & 1; flt[ch][pidx(f)] = delay_saturate(flt[ch][pidx(f)], drive); }
    __syncthreads();
    if (warp < 2) {  // DC blocker + clamp -> clean (dly), one warp per channel
      const uint32_t ch = warp;
      double x1 = ch ? s.dcr_x1 : s.dcl_x1, y1 = ch ? s.dcr_y1 : s.dcl_y1;
      dc_scan_channel(s.dc_r, x1, y1, flt[ch], dly[ch], len, lane);
      if (lane == 0) { if (ch) { s.dcr_x1 = x1; s.dcr_y1 = y1; } else { s.dcl_x1 = x1; s.dcl_y1 = y1; } }
    }
    __syncthreads();
    // (3) line writes (input + previous frame's clean feedback) and the output mix
    for (uint32_t f = f0 + tid; f < f1; f += nt) {
      const float lv = lfo[f];
      const float fb = fminf(fmaxf(base_fb + lv * fb_depth * (1.0f - fabsf(base_fb)), 0.0f), 0.999f);
      const float li = CB_L(f), ri = CB_R(f);
      const float pl = f == f0 ? fb_prev[0] : dly[0][pidx(f - f0 - 1)];
      const float pr = f == f0 ? fb_prev[1] : dly[1][pidx(f - f0 - 1)];
      float l_in, r_in;
      if (mode == 0) { l_in = li + pl * fb; r_in = ri + pr * fb; }
      else { const float mono = (li + ri) * 0.5f; l_in = mono + pr * fb; r_in = pl * fb; }
      const float cl = dly[0][pidx(f - f0)], cr = dly[1][pidx(f - f0)];
      // InterpolatedDelayLine::process writes input + output * 0.0 (delay.rs:145-150)
      const double zl = (double)(float)0.0f;
      (void)zl;
      bufs[0][(wp0[0] + f) & mk[0]] = (double)l_in + (double)0.0;  // out * 0.0 == +-0.0
      bufs[1][(wp0[1] + f) & mk[1]] = (double)r_in + (double)0.0;
      const float dry_gain = fminf((1.0f - wet) * 2.0f, 1.0f);
      const float wet_gain = fminf(wet * 2.0f, 1.0f);
      const float ol = li * dry_gain + cl * wet_gain;
      const float orr = ri * dry_gain + cr * wet_gain;
      const float mid = (ol + orr) * 0.5f;
      const float side = (ol - orr) * 0.5f;
      CB_L(f) = mid + side * width;
      CB_R(f) = mid - side * width;
    }
    __syncthreads();
    fb_prev[0] = dly[0][pidx(len - 1)];
    fb_prev[1] = dly[1][pidx(len - 1)];
    __syncthreads();
  }
  if (tid == 0) {
    s.dl.write_pos = (wp0[0] + n) & mk[0];
    s.dr.write_pos = (wp0[1] + n) & mk[1];
    s.fb_l = fb_prev[0]; s.fb_r = fb_prev[1];
  }
  __syncthreads();
  return true;
}

// ---- ReverbEffect (reverb.rs:217-369, 409-447) while room size and wet are not ramping -----------------------------
// position of a "count += 1; if count > delay { count = 0 }" counter k steps after it was w0
PB_DEV uint32_t cyc_pos(uint32_t w0, uint32_t k, uint32_t delay) {
  uint32_t v;
  if (w0 > delay) { if (k == 0) return w0; v = k - 1; }
  else v = w0 + k;
  const uint32_t period = delay + 1;
  if (v >= 4 * period) return v % period;  // k <= 1025 and period >= 257 here: at most a few wraps
  while (v >= period) v -= period;
  return v;
}

constexpr uint32_t RV_BATCH = FX_THREADS_C / 2;  // frames per delay-line read/write batch (one thread per frame x channel)
__host__ __device__ constexpr size_t rv_work_bytes(uint32_t L) { return ((size_t)10 * (L + L / 32 + 4) + (size_t)2 * 16 * RV_BATCH + 32 + 16) * sizeof(double); }

// RVL = frames per sub-block: every feedback path of the reverb goes through a delay of at least RVL frames (checked by
// the caller), so inside a sub-block each stage is a pass over all frames. 1024 (a whole chunk) when the room is large
// enough, else 256 (predelay >= 29 * 25 = 725 always).
template <uint32_t RVL>
PB_DEV void reverb_parallel_impl(ReverbState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t n, uint32_t tid, uint32_t nt, ParWork w,
                                 const RvDerived& d, const uint32_t predelay) {
  constexpr uint32_t PL = RVL + RVL / 32 + 4;
  constexpr uint32_t E = 2 * RVL / FX_THREADS_C;  // (frame, channel) elements per thread and pass
  const double wet = (double)s.wet.target;
  const double blend = d.blend, regen = d.regen;
  const double vib_speed = 0.1, vib_depth = 7.0;
  uint8_t* p = w.base;
  double* A[2] = {carve<double>(p, PL), carve<double>(p, PL)};
  double* AP[4][2];
  for (int i = 0; i < 4; ++i) { AP[i][0] = carve<double>(p, PL); AP[i][1] = carve<double>(p, PL); }
  double* VT = carve<double>(p, 16 * RV_BATCH);  // (cos, sin)(k * inc_line), k = 1..RV_BATCH: [line][k - 1][2]
  double* VB = carve<double>(p, 32);             // (sin, cos) of every (line, channel)'s phase at the batch start
  double* FB = carve<double>(p, 16 * RV_BATCH);  // feedback of every frame of the batch, same layout
  double* carry = carve<double>(p, 16);          // feedback of the last frame before the batch, [line*2+ch]
  const uint32_t lane = tid & 31, warp = tid >> 5;
  double* aux = cx.aux_arena;
  // counters at chunk start
  const uint32_t m_w0 = s.m_write_pos & s.m_mask;
  uint32_t ap_w0[4], ap_delay[4], ln_c0[8], ln_delay[8];
  uint32_t ap_aux[4], ln_aux[8];
  for (int i = 0; i < 4; ++i) { ap_w0[i] = s.ap[i].write_pos; ap_delay[i] = s.ap[i].delay; ap_aux[i] = s.ap[i].aux; }
  for (int i = 0; i < 8; ++i) { ln_c0[i] = s.lines[i].count; ln_delay[i] = s.lines[i].delay; ln_aux[i] = s.lines[i].aux; }
  const uint32_t m_aux = s.m_aux;
  const double fpd_l = (double)s.fpd_l * 1.18e-17, fpd_r = (double)s.fpd_r * 1.18e-17;
  const int src[8] = {3, 2, 1, 0, 0, 1, 2, 3};  // a<-l, b<-k, c<-j, d<-i, e<-i, f<-j, g<-k, h<-l
  if (tid < 16) carry[tid] = s.lines[tid >> 1].feedback[tid & 1];
  // The vibrato phase of a line advances by a constant per frame (reverb.rs:596-604), so sin(phase) of frame k of a batch
  // is a rotation of the batch's start phase by k * inc: the 2048 f64 sines of a batch become 16 sincos + a table
  // lookup and two multiplies per read. The phase itself is still accumulated serially (the state must stay the
  // reference's); the rotated argument differs from the accumulated one by rounding noise (~1e-11 rad).
  for (uint32_t i = tid; i < 8 * RV_BATCH; i += nt) {
    const uint32_t line = i / RV_BATCH, k = i % RV_BATCH + 1;
    double sk, ck;
    sincos((double)k * (s.lines[line].depth * vib_speed), &sk, &ck);
    VT[i * 2] = ck; VT[i * 2 + 1] = sk;
  }

  for (uint32_t f0 = 0; f0 < n; f0 += RVL) {
    const uint32_t len = min(RVL, n - f0);
    // (a) input + denormal dither, predelay DelayLine<2>::process (delay.rs:47-66): reads then writes
    {
      double keep[E];  // q is a compile-time index: the array stays in registers
#pragma unroll
      for (uint32_t q = 0; q < E; ++q) {
        const uint32_t i = tid + q * FX_THREADS_C;
        if (i < len * 2) {
          const uint32_t f = i >> 1, ch = i & 1;
          double x = (double)(ch ? CB_R(f0 + f) : CB_L(f0 + f));
          if (fabs(x) < 1.18e-23) x = ch ? fpd_r : fpd_l;
          keep[q] = x;
          A[ch][pidx(f)] = aux[m_aux + (size_t)cyc_pos(m_w0, f0 + f + 1, predelay) * 2 + ch];
        }
      }
      __syncthreads();
#pragma unroll
      for (uint32_t q = 0; q < E; ++q) {
        const uint32_t i = tid + q * FX_THREADS_C;
        if (i < len * 2) {
          const uint32_t f = i >> 1, ch = i & 1;
          aux[m_aux + (size_t)cyc_pos(m_w0, f0 + f, predelay) * 2 + ch] = keep[q];
        }
      }
    }
    // (b) biquad A (block scan, in place), * wet, sin
    {
      const BiquadCoef cf = s.ca;
      biquad_scan_planes64(cf, s.a_ic, A, cb.lane_state, len, tid);   // (all 256 threads; ends with a barrier)
    }
    // (c) sin(x * wet), then the four Schroeder allpasses in series (delay.rs:314-350). An allpass reads the slot the NEXT
    // frame overwrites, but never a slot written inside this sub-block (delay >= RVL): the chain is pointwise per
    // (frame, channel) -- all reads of the four stages first, then, behind one barrier, all writes.
    {
      double keep[4][E];
#pragma unroll
      for (uint32_t q = 0; q < E; ++q) {
        const uint32_t i = tid + q * FX_THREADS_C;
        if (i < len * 2) {
          const uint32_t f = i >> 1, ch = i & 1;
          double delayed[4];
#pragma unroll
          for (int st = 0; st < 4; ++st) delayed[st] = aux[ap_aux[st] + (size_t)cyc_pos(ap_w0[st], f0 + f + 1, ap_delay[st]) * 2 + ch];
          double in = sin(A[ch][pidx(f)] * wet);
#pragma unroll
          for (int st = 0; st < 4; ++st) {
            const double buf = in - (delayed[st] * 0.5);
            keep[st][q] = buf;
            in = buf * 0.5 + delayed[st];
            AP[st][ch][pidx(f)] = in;
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (uint32_t q = 0; q < E; ++q) {
        const uint32_t i = tid + q * FX_THREADS_C;
        if (i < len * 2) {
          const uint32_t f = i >> 1, ch = i & 1;
#pragma unroll
          for (int st = 0; st < 4; ++st) aux[ap_aux[st] + (size_t)cyc_pos(ap_w0[st], f0 + f, ap_delay[st]) * 2 + ch] = keep[st][q];
        }
      }
    }
    // (d) eight modulated delay lines + Householder feedback, in batches of RV_BATCH frames (the vibrato reads reach up
    // to 15 frames ahead of the write position, never into frames of the same batch that are not written yet)
    for (uint32_t b0 = 0; b0 < len; b0 += RV_BATCH) {
      const uint32_t bl = min(RV_BATCH, len - b0);
      if (tid < 16) {  // vibrato phase chains (reverb.rs:596-604): serial f64 accumulate per (line, channel)
        RvLine& L = s.lines[tid >> 1];
        double ph = L.vib_phase[tid & 1];
        double sb, cb_;
        sincos(ph, &sb, &cb_);
        VB[tid * 2] = sb; VB[tid * 2 + 1] = cb_;
        const double inc = L.depth * vib_speed;
        for (uint32_t f = 0; f < bl; ++f) ph += inc;
        L.vib_phase[tid & 1] = ph;
      }
      __syncthreads();
      // gets (reverb.rs:554-586) for frame (b0 + fb) of the sub-block, channel ch
      double o[8];
      const uint32_t fb_i = tid >> 1, ch = tid & 1;
      const bool act = fb_i < bl;
      const uint32_t fabs_ = f0 + b0 + fb_i;  // frame index within the chunk
      if (act) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double* lb = aux + ln_aux[i];
          const uint32_t dl = ln_delay[i];
          const uint32_t cnt = cyc_pos(ln_c0[i], fabs_ + 1, dl);
          const double vsin = VB[(i * 2 + ch) * 2] * VT[(i * RV_BATCH + fb_i) * 2] + VB[(i * 2 + ch) * 2 + 1] * VT[(i * RV_BATCH + fb_i) * 2 + 1];
          const double offset = (vsin + 1.0) * vib_depth;
          const double working = (double)cnt + offset;
          const double wf = floor(working);
          const double frac = working - wf;
          const uint32_t wi = (uint32_t)wf;
          uint32_t r1 = wi; if (r1 > dl) r1 -= dl + 1;
          uint32_t r2 = wi + 1; if (r2 > dl) r2 -= dl + 1;
          const double v1 = lb[(size_t)r1 * 2 + ch], v2 = lb[(size_t)r2 * 2 + ch];
          double ip = v1 * (1.0 - frac) + v2 * frac;
          ip = (1.0 - blend) * ip + (v1 * blend);
          o[i] = ip;
        }
        FB[(0 * 2 + ch) * RV_BATCH + fb_i] = (o[0] - (o[1] + o[2] + o[3])) * regen;
        FB[(1 * 2 + ch) * RV_BATCH + fb_i] = (o[1] - (o[0] + o[2] + o[3])) * regen;
        FB[(2 * 2 + ch) * RV_BATCH + fb_i] = (o[2] - (o[0] + o[1] + o[3])) * regen;
        FB[(3 * 2 + ch) * RV_BATCH + fb_i] = (o[3] - (o[0] + o[1] + o[2])) * regen;
        FB[(4 * 2 + ch) * RV_BATCH + fb_i] = (o[4] - (o[5] + o[6] + o[7])) * regen;
        FB[(5 * 2 + ch) * RV_BATCH + fb_i] = (o[5] - (o[4] + o[6] + o[7])) * regen;
        FB[(6 * 2 + ch) * RV_BATCH + fb_i] = (o[6] - (o[4] + o[5] + o[7])) * regen;
        FB[(7 * 2 + ch) * RV_BATCH + fb_i] = (o[7] - (o[4] + o[5] + o[6])) * regen;
      }
      __syncthreads();
      // sets: set(frame) = allpass output of the frame + the feedback of frame - 1 (the neighbouring thread's, or the
      // carry of the previous batch for the first frame)
      if (act) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          double* lb = aux + ln_aux[i];
          const double prev_fb = fb_i == 0 ? carry[i * 2 + ch] : FB[(i * 2 + ch) * RV_BATCH + fb_i - 1];
          const uint32_t cnt = cyc_pos(ln_c0[i], fabs_, ln_delay[i]);
          lb[(size_t)cnt * 2 + ch] = AP[src[i]][ch][pidx(b0 + fb_i)] + prev_fb;
        }
        // mean of the eight lines -> input of biquad B (reuse A)
        A[ch][pidx(b0 + fb_i)] = (o[0] + o[1] + o[2] + o[3] + o[4] + o[5] + o[6] + o[7]) / 8.0;
      }
      __syncthreads();
      if (tid < 16) carry[tid] = FB[tid * RV_BATCH + bl - 1];  // read again only behind the next batch's two barriers
    }
    // (e) biquad B, clamp, asin, biquad C, dry mix
    {
      const BiquadCoef cf = s.cb;
      biquad_scan_planes64(cf, s.b_ic, A, cb.lane_state, len, tid);   // (all 256 threads; ends with a barrier)
    }
    for (uint32_t i = tid; i < len * 2; i += nt) {
      const uint32_t f = i >> 1, ch = i & 1;
      A[ch][pidx(f)] = asin(fmin(fmax(A[ch][pidx(f)], -1.0), 1.0));
    }
    __syncthreads();
    {
      const BiquadCoef cf = s.cc;
      biquad_scan_planes64(cf, s.c_ic, A, cb.lane_state, len, tid);   // (all 256 threads; ends with a barrier)
    }
    for (uint32_t i = tid; i < len * 2; i += nt) {
      const uint32_t f = i >> 1, ch = i & 1;
      double x = (double)(ch ? CB_R(f0 + f) : CB_L(f0 + f));
      if (fabs(x) < 1.18e-23) x = ch ? fpd_r : fpd_l;
      double y = A[ch][pidx(f)];
      if (wet != 1.0) y += x * (1.0 - wet);
      if (ch) CB_R(f0 + f) = (float)y; else CB_L(f0 + f) = (float)y;
    }
    __syncthreads();
  }
  if (tid < 16) s.lines[tid >> 1].feedback[tid & 1] = carry[tid];
  if (tid == 0) {
    s.m_write_pos = cyc_pos(m_w0, n, predelay);
    for (int i = 0; i < 4; ++i) s.ap[i].write_pos = cyc_pos(ap_w0[i], n, ap_delay[i]);
    for (int i = 0; i < 8; ++i) s.lines[i].count = cyc_pos(ln_c0[i], n, ln_delay[i]);
  }
  __syncthreads();
}

PB_DEV bool reverb_parallel(ReverbState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t n, uint32_t tid, uint32_t nt, ParWork w) {
  if (lin_need_ramp(s.room) || exp_need_ramp(s.wet, cx.comp)) return false;
  const double room = (double)s.room.target, wet = (double)s.wet.target;
  const RvDerived d = reverb_derive(room, wet);
  __syncthreads();
  if (tid == 0) {  // per-call updates of ReverbEffect::process (reverb.rs:428-441)
    (void)reverb_update_sizes(s, d.size);
    reverb_update_filters(s, cx, d.cutoff);
  }
  __syncthreads();
  const uint32_t predelay = f64_as_usize32(29.0 * d.size);
  // shortest feedback lag: the predelay, the shortest allpass (ap[3]) and the shortest line (lines[7]) less its look-ahead
  const uint32_t lag = min(min(predelay, s.ap[3].delay), s.lines[7].delay >= 16 ? s.lines[7].delay - 16 : 0u);
  if (lag >= 1024 && w.bytes >= rv_work_bytes(1024)) reverb_parallel_impl<1024>(s, cx, cb, n, tid, nt, w, d, predelay);
  else if (lag >= 256 && w.bytes >= rv_work_bytes(256)) reverb_parallel_impl<256>(s, cx, cb, n, tid, nt, w, d, predelay);
  else return false;
  return true;
}

// Effect::process for the mixer CTA: the chunk-parallel form when it applies, else the frame-serial reference form
PB_DEV void fx_process_chunk(const FxHeader& h, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames, uint32_t tid, uint32_t nt, ParWork w) {
  uint8_t* st = cx.state_arena + h.state_offset;
  bool done = false;
  switch (h.kind) {
    case FX_COMPRESSOR: done = comp_parallel(*(CompState*)st, cx, cb, frames, tid, nt, w); break;
    case FX_CHORUS: done = chorus_parallel(*(ChorusState*)st, cx, cb, frames, tid, nt, w); break;
    case FX_DELAY: done = delay_parallel(*(DelayState*)st, cx, cb, frames, tid, nt, w); break;
    case FX_REVERB: done = reverb_parallel(*(ReverbState*)st, cx, cb, frames, tid, nt, w); break;
    case FX_GATE: gate_process(*(GateState*)st, cx, cb, frames, tid, nt, reinterpret_cast<float*>(w.base)); done = true; break;
    default: break;
  }
  if (!done) fx_process(h, cx, cb, frames, tid);
}

}  // namespace pb
