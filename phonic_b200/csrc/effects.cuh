// Effect chain on device: the six effects BASELINE.json names, restated for one mixer bus.
// Filters and delay storage stay f64 as in the reference; f32 round-trips between
// stages are kept. Per-channel effects (Filter, Eq5) run one lane per channel; stereo-coupled effects
// (Compressor, Chorus, Delay, Reverb) run on lane 0, the reverb spreads its 8x2 delay-line reads over
// 16 lanes. The unit is compiled with -fmad=false so every op is rounded as the scalar reference does;
// device libm (tan/pow/sin/asin/log10/exp) may differ from the host libm in the last ulp.
#pragma once
#include "voice.cuh"
#include "effects_state.h"

namespace pb {

// ---- smoothers (src/utils/smoothing.rs) -------------------------------------------------------------
PB_DEV bool lin_need_ramp(const LinSm& s) { return s.pending > 0; }
PB_DEV float lin_next(LinSm& s) {
  if (s.pending > 0) {
    s.current += s.current_step;
    s.pending -= 1;
    if (s.pending == 0) s.current = s.target;
    return s.current;
  }
  return s.target;
}
PB_DEV void lin_set_target(LinSm& s, float t, float comp) {
  s.target = t;
  if (s.current == s.target) { s.pending = 0; return; }
  s.current_step = (s.current > s.target) ? -s.step * comp : s.step * comp;
  float pending = (s.target - s.current) / s.current_step;
  float r = fmaxf(roundf(pending), 0.0f);
  s.pending = (r >= 4294967296.0f) ? 0xFFFFFFFFu : (uint32_t)r;
  if (s.pending == 0) s.current = s.target;
}
PB_DEV bool spring_need_ramp(const SpringSm& s) {
  const float EPS = F32_EPS * 100.0f;
  return fabsf(s.velocity) > EPS || fabsf(s.target - s.current) > EPS;
}
PB_DEV float spring_next(SpringSm& s, float comp) {
  if (spring_need_ramp(s)) {
    float om = s.omega * comp;
    float k = om * om;
    float d = 2.0f * om;
    s.velocity += (s.target - s.current) * k - s.velocity * d;
    s.current += s.velocity;
    return s.current;
  }
  return s.target;
}

// ---- BiquadFilterCoefficients::set/apply (src/utils/dsp/filters/biquad.rs:127-283) ---------------------
enum BqType : uint32_t { BQ_LOWPASS, BQ_HIGHPASS, BQ_BANDPASS, BQ_NOTCH, BQ_PEAK, BQ_ALLPASS, BQ_BELL, BQ_LOWSHELF, BQ_HIGHSHELF };
PB_DEV void biquad_apply(BiquadCoef& c) {
  const double PI = 3.14159265358979323846;
  double g = tan(PI * (double)c.cutoff / (double)c.sample_rate);
  double k = 1.0 / (double)c.q;
  double a = 0.0;
  if (c.type == BQ_BELL) { a = pow(10.0, (double)c.gain / 40.0); k = 1.0 / ((double)c.q * a); }
  else if (c.type == BQ_LOWSHELF) { a = pow(10.0, (double)c.gain / 40.0); g = g / sqrt(a); }
  else if (c.type == BQ_HIGHSHELF) { a = pow(10.0, (double)c.gain / 40.0); g = g * sqrt(a); }
  c.a1 = 1.0 / (1.0 + g * (g + k));
  c.a2 = g * c.a1;
  c.a3 = g * c.a2;
  switch (c.type) {
    case BQ_LOWPASS: c.m0 = 0.0; c.m1 = 0.0; c.m2 = 1.0; break;
    case BQ_HIGHPASS: c.m0 = 1.0; c.m1 = -k; c.m2 = -1.0; break;
    case BQ_BANDPASS: c.m0 = 0.0; c.m1 = 1.0; c.m2 = 0.0; break;
    case BQ_NOTCH: c.m0 = 1.0; c.m1 = -k; c.m2 = 0.0; break;
    case BQ_PEAK: c.m0 = 1.0; c.m1 = -k; c.m2 = -2.0; break;
    case BQ_ALLPASS: c.m0 = 1.0; c.m1 = -2.0 * k; c.m2 = 0.0; break;
    case BQ_BELL: c.m0 = 1.0; c.m1 = k * (a * a - 1.0); c.m2 = 0.0; break;
    case BQ_LOWSHELF: c.m0 = 1.0; c.m1 = k * (a - 1.0); c.m2 = a * a - 1.0; break;
    default: c.m0 = a * a; c.m1 = k * (1.0 - a) * a; c.m2 = 1.0 - a * a; break;
  }
}
PB_DEV void biquad_set(BiquadCoef& c, uint32_t type, uint32_t sr, float cutoff, float q, float gain) {
  if (c.type != type || c.sample_rate != sr || c.cutoff != cutoff || c.q != q || c.gain != gain) {
    c.type = type; c.sample_rate = sr; c.cutoff = cutoff; c.q = q; c.gain = gain;
    // parameter validation failures of the reference (`expect`) cannot occur for in-range parameters
    if (q > 0.0f && cutoff <= (float)sr / 2.0f) biquad_apply(c);
  }
}
PB_DEV double biquad_tick(const BiquadCoef& c, double& ic1, double& ic2, double v0) {
  double v3 = v0 - ic2;
  double v1 = c.a1 * ic1 + c.a2 * v3;
  double v2 = ic2 + c.a2 * ic1 + c.a3 * v3;
  ic1 = 2.0 * v1 - ic1;
  ic2 = 2.0 * v2 - ic2;
  return c.m0 * v0 + c.m1 * v1 + c.m2 * v2;
}

// ---- SvfFilterCoefficients (src/utils/dsp/filters/svf.rs:112-169) ------------------------------------------
PB_DEV void svf_set(SvfCoef& c, uint32_t type, uint32_t sr, float cutoff, float res) {
  if (c.type != type || c.sample_rate != sr || c.cutoff != cutoff || c.resonance != res) {
    c.type = type; c.sample_rate = sr; c.cutoff = cutoff; c.resonance = res;
    if (res >= 0.0f && res <= 1.0f && cutoff <= (float)sr / 2.0f) {
      const double PI = 3.14159265358979323846;
      c.g = tan(PI * (double)cutoff / (double)sr);
      c.k = fmax(2.0 * (1.0 - (double)res * 0.97), 0.03);
      c.a1 = 1.0 / (1.0 + c.g * (c.g + c.k));
      c.a2 = c.g * c.a1;
      c.a3 = c.g * c.a2;
    }
  }
}
PB_DEV void svf_set_type(SvfCoef& c, uint32_t type) {
  if (c.type != type) svf_set(c, type, c.sample_rate, c.cutoff, c.resonance);
}
PB_DEV double svf_tick(const SvfCoef& c, double& ic1, double& ic2, double in) {
  double v3 = in - ic2;
  double v1 = c.a1 * ic1 + c.a2 * v3;
  double v2 = ic2 + c.a2 * ic1 + c.a3 * v3;
  ic1 = 2.0 * v1 - ic1;
  ic2 = 2.0 * v2 - ic2;
  if (c.type == 0) return v2;
  if (c.type == 2) return v1;
  return in - c.k * v1 - v2;
}

// ---- Lfo (src/utils/dsp/lfo.rs), deterministic waveforms ----------------------------------------------------
PB_DEV float sine_approx(float x) {
  const float PI = 3.14159265358979323846f;
  const float B = 4.0f / PI;
  const float C = -4.0f / (PI * PI);
  const float P = 0.225f;
  float y = B * x + C * x * fabsf(x);
  return P * (y * fabsf(y) - y) + y;
}
PB_DEV void lfo_set_phase_degrees(LfoSt& l, float p) {
  float q = p / 6.28318530717958647692f;
  float r = fmodf(q, 1.0f);
  l.phase = r < 0.0f ? r + 1.0f : r;
}
PB_DEV void lfo_set_rate(LfoSt& l, uint32_t sr, double rate) { l.phase_inc = (float)(rate / (double)sr); }
// the waveform at a given phase (lfo.rs:122-169): a pure function of the phase
PB_DEV float lfo_wave(uint32_t waveform, float phase) {
  const float TAU = 6.28318530717958647692f;
  switch (waveform) {
    case 0: { float p = phase < 0.5f ? phase * TAU : (phase - 1.0f) * TAU; return sine_approx(p); }
    case 1: return phase < 0.25f ? phase * 4.0f : (phase < 0.75f ? 2.0f - phase * 4.0f : phase * 4.0f - 4.0f);
    case 2: return phase * 2.0f - 1.0f;
    case 3: return 1.0f - phase * 2.0f;
    default: return phase < 0.5f ? 1.0f : -1.0f;
  }
}
PB_DEV float lfo_run(LfoSt& l) {
  const float v = lfo_wave(l.waveform, l.phase);
  l.phase += l.phase_inc;
  if (l.phase >= 1.0f) l.phase -= 1.0f;
  return v;
}

// exp of an f32 as the host libm gives it (glibc's expf is correctly rounded to within 0.502 ulp): the f64 exp rounded to f32.
// Used where a coefficient is re-derived on the device after a parameter event -- a one-pole release coefficient
// exp(-1 / (2 s x 48 kHz)) = 1 - 1.04e-5 has only ~7 significant bits in its distance from 1, so a last-bit difference
// of the device's expf shows up as 5e-4 in a compressor's output (the same coefficients come from the host at construction).
PB_DEV float expf_host_rounding(const float x) { return (float)exp((double)x); }
PB_DEV float db_to_linear_dev(float value) {  // src/utils.rs:41-51
  const float DB_TO_LIN_FACTOR = 2.302585092994046f / 20.0f;
  if (isnan(value)) return value;
  if (value == 0.0f) return 1.0f;
  if (value > -200.0f) return expf(value * DB_TO_LIN_FACTOR);
  return 0.0f;
}

// InterpolatedDelayLine<1>::process (src/utils/dsp/delay.rs:107-155)
PB_DEV float idelay_process(IDelay& d, double* __restrict__ buf, float input, float feedback, float delay) {
  double read_pos = (double)d.write_pos - (double)delay;
  double fl = floor(read_pos);
  double fraction = read_pos - fl;
  long long index1 = (long long)fl;
  uint32_t i1 = (uint32_t)((unsigned long long)index1 & d.mask);
  uint32_t i2 = (uint32_t)((unsigned long long)(index1 + 1) & d.mask);
  double v1 = buf[i1], v2 = buf[i2];
  float out = (float)(v1 + (v2 - v1) * fraction);
  buf[d.write_pos & d.mask] = (double)input + (double)out * (double)feedback;
  d.write_pos = (d.write_pos + 1) & d.mask;
  return out;
}

struct FxCtx {
  uint32_t sample_rate;
  float comp;
  uint8_t* state_arena;
  double* aux_arena;
};

// A chunk staged in shared memory, planar and padded (one pad word per 32 samples) so that both the
// frame-serial effects and the lane-per-sub-block scans access it without bank conflicts.
struct ChunkBuf {
  float* ch[2];
  double* scratch;   // [2][pidx(1024) + 1] zero-state responses of the scan
  double* lane_state;// [2][32][2] sub-block end states / start states
};
PB_DEV uint32_t pidx(uint32_t f) { return f + (f >> 5); }
constexpr uint32_t CHUNK_MAX = 1024;  // frames per chunk (<= WavStream block)
constexpr uint32_t PLANE = 1060;      // pidx(CHUNK_MAX - 1) + a few words
#define CB_L(f) cb.ch[0][pidx(f)]
#define CB_R(f) cb.ch[1][pidx(f)]


// ---- block-parallel evaluation of one time-invariant biquad over a staged chunk ---------------------------
// The Cytomic SVF tick (biquad.rs:314-323) is linear in (ic1eq, ic2eq, x):
//   s' = A s + B x,  y = C s + D x,   A = [[2a1-1, -2a2], [2a2, 1-2a3]]
// One warp handles one channel: lane j runs the *reference's own tick* over sub-block j from a zero state
// (zero-state response y0 and end state e_j, f64), lane 0 then chains the true sub-block start states
// s_{j+1} = A^B s_j + e_j, and every lane adds the homogeneous part y[n] = y0[n] + (C A^k) s_j.
// Exact in real arithmetic; in f64 it differs from the frame-serial evaluation by O(1e-16) relative, far
// below the f32 output quantum. Only used while no parameter is ramping.
struct Mat2 { double a, b, c, d; };
PB_DEV Mat2 mat_mul(const Mat2& x, const Mat2& y) {
  return Mat2{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}
PB_DEV Mat2 mat_pow(Mat2 m, uint32_t k) {
  Mat2 r{1.0, 0.0, 0.0, 1.0};
  while (k) {
    if (k & 1u) r = mat_mul(r, m);
    m = mat_mul(m, m);
    k >>= 1;
  }
  return r;
}
// x: f32 input (planar padded); result written back to x as f32, or -- when out64 != nullptr -- left as f64
// in out64 (planar padded) for callers whose next stage works in f64.
PB_DEV void biquad_scan_channel(const BiquadCoef& c, double& ic1_io, double& ic2_io, float* x, double* y0, double* lane_state,
                                uint32_t len, uint32_t lane, double* out64 = nullptr, const double* in64 = nullptr) {
  const uint32_t B = (len + 31) / 32;             // samples per lane
  const uint32_t lo = min(lane * B, len), hi = min(lo + B, len);
  const Mat2 A{2.0 * c.a1 - 1.0, -2.0 * c.a2, 2.0 * c.a2, 1.0 - 2.0 * c.a3};
  // phase 1: zero-state response of my sub-block with the reference tick
  double ic1 = 0.0, ic2 = 0.0;
  for (uint32_t n = lo; n < hi; ++n) y0[pidx(n)] = biquad_tick(c, ic1, ic2, in64 ? in64[pidx(n)] : (double)x[pidx(n)]);
  lane_state[lane * 2] = ic1;
  lane_state[lane * 2 + 1] = ic2;
  __syncwarp();
  // phase 2: chain the sub-block start states (lane 0), leave them in lane_state
  if (lane == 0) {
    const Mat2 AB = mat_pow(A, B);
    double s1 = ic1_io, s2 = ic2_io;
    for (uint32_t j = 0; j < 32; ++j) {
      const uint32_t jl = min(j * B, len), jh = min(jl + B, len);
      const double e1 = lane_state[j * 2], e2 = lane_state[j * 2 + 1];
      lane_state[j * 2] = s1;
      lane_state[j * 2 + 1] = s2;
      if (jh > jl) {
        const Mat2 P = (jh - jl == B) ? AB : mat_pow(A, jh - jl);
        const double n1 = P.a * s1 + P.b * s2 + e1;
        const double n2 = P.c * s1 + P.d * s2 + e2;
        s1 = n1; s2 = n2;
      }
    }
    ic1_io = s1; ic2_io = s2;
  }
  __syncwarp();
  // phase 3: add the homogeneous response of my sub-block's true start state
  const double s1 = lane_state[lane * 2], s2 = lane_state[lane * 2 + 1];
  // y = C s + D x with C = (m1 a1 + m2 a2, -m1 a2 + m2 (1 - a3))
  const double C1 = c.m1 * c.a1 + c.m2 * c.a2, C2 = -c.m1 * c.a2 + c.m2 * (1.0 - c.a3);
  double h1 = s1, h2 = s2;  // A^k s
  for (uint32_t n = lo; n < hi; ++n) {
    const double y = y0[pidx(n)] + (C1 * h1 + C2 * h2);
    if (out64) out64[pidx(n)] = y; else x[pidx(n)] = (float)y;
    const double t1 = A.a * h1 + A.b * h2, t2 = A.c * h1 + A.d * h2;
    h1 = t1; h2 = t2;
  }
  __syncwarp();
}

// Both channels of one time-invariant biquad with ALL threads of the mixer CTA (FX_THREADS = 256): threads
// [0,128) / [128,256) = channel 0 / 1, one sub-block of ceil(len/128) samples per thread. Same three phases as
// biquad_scan_channel, but the sub-block start states come from a parallel prefix over affine maps instead of one
// lane's serial chain: every sub-block before the last non-empty one has the same transition P = A^B, so
//   S_j = P^j W + sum_{i<j} P^(j-1-i) e_i   (W = state at the warp's first sub-block)
// is a Hillis-Steele scan over the end states e_i with the powers P^(2^k) (5 shuffle steps per warp), the four
// warps of a channel are chained through shared memory, and P^j W is built from the same powers.
// The chunk-parallel f64 dependent-op latency on this part is ~40 cycles, so depth is what matters:
// 8 + 8 ticks and ~25 scan levels instead of 32 + 32 ticks and ~110 serial levels.
// Ends with a CTA barrier. `ic1`, `ic2`: per-channel state at [ch * stride] (read at entry, written by the owner of the
// last sample).
PB_DEV void biquad_scan_stereo(const BiquadCoef& c, double* ic1, double* ic2, uint32_t stride, const ChunkBuf& cb, uint32_t len, uint32_t tid) {
  const uint32_t ch = tid >> 7, j = tid & 127u, lane = tid & 31u, w = j >> 5;
  const double in1 = ic1[ch * stride], in2 = ic2[ch * stride];
  __syncthreads();  // every thread has read the incoming state before the owner of the last sample replaces it
  if (len == 0) return;
  const uint32_t B = (len + 127u) / 128u;
  const uint32_t lo = min(j * B, len), hi = min(lo + B, len);
  float* x = cb.ch[ch];
  double* y0 = cb.scratch + (size_t)ch * 1060;
  double* ls = cb.lane_state + (size_t)ch * 64;
  const Mat2 A{2.0 * c.a1 - 1.0, -2.0 * c.a2, 2.0 * c.a2, 1.0 - 2.0 * c.a3};
  // phase 1: zero-state response of my sub-block with the reference tick
  double e1 = 0.0, e2 = 0.0;
  for (uint32_t n = lo; n < hi; ++n) y0[pidx(n)] = biquad_tick(c, e1, e2, (double)x[pidx(n)]);
  // phase 2a: inclusive scan E_j = sum_{i<=j} P^(j-i) e_i inside the warp
  const Mat2 P = mat_pow(A, B);
  Mat2 Pd = P;
  double E1 = e1, E2 = e2;
#pragma unroll
  for (uint32_t d = 1; d < 32; d <<= 1) {
    const double u1 = __shfl_up_sync(0xFFFFFFFFu, E1, d), u2 = __shfl_up_sync(0xFFFFFFFFu, E2, d);
    if (lane >= d) { E1 += Pd.a * u1 + Pd.b * u2; E2 += Pd.c * u1 + Pd.d * u2; }
    Pd = mat_mul(Pd, Pd);
  }
  // Pd = P^32: the transition of a whole warp
  if (lane == 31) { ls[w * 2] = E1; ls[w * 2 + 1] = E2; }
  __syncthreads();
  // phase 2b: state at my warp's first sub-block
  double W1 = in1, W2 = in2;
  for (uint32_t q = 0; q < w; ++q) {
    const double t1 = Pd.a * W1 + Pd.b * W2 + ls[q * 2], t2 = Pd.c * W1 + Pd.d * W2 + ls[q * 2 + 1];
    W1 = t1; W2 = t2;
  }
  // phase 2c: S_j = P^lane W + E_(lane-1)
  double p1 = __shfl_up_sync(0xFFFFFFFFu, E1, 1), p2 = __shfl_up_sync(0xFFFFFFFFu, E2, 1);
  if (lane == 0) { p1 = 0.0; p2 = 0.0; }
  Mat2 Q = P;
#pragma unroll
  for (uint32_t k = 0; k < 5; ++k) {
    if ((lane >> k) & 1u) { const double t1 = Q.a * W1 + Q.b * W2, t2 = Q.c * W1 + Q.d * W2; W1 = t1; W2 = t2; }
    Q = mat_mul(Q, Q);
  }
  // phase 3: add the homogeneous response of my sub-block's true start state
  const double C1 = c.m1 * c.a1 + c.m2 * c.a2, C2 = -c.m1 * c.a2 + c.m2 * (1.0 - c.a3);
  double h1 = W1 + p1, h2 = W2 + p2;
  for (uint32_t n = lo; n < hi; ++n) {
    x[pidx(n)] = (float)(y0[pidx(n)] + (C1 * h1 + C2 * h2));
    const double t1 = A.a * h1 + A.b * h2, t2 = A.c * h1 + A.d * h2;
    h1 = t1; h2 = t2;
  }
  // the owner of the last sample holds the chunk's end state: A^(its length) S_j + e_j
  if (hi == len && lo < hi) { ic1[ch * stride] = h1 + e1; ic2[ch * stride] = h2 + e2; }
  __syncthreads();
}

// The same scan over two f64 planes, in place (the reverb's three filters work on its f64 signal path): threads [0,128) /
// [128,256) = plane 0 / 1. `ic`: the filter's state as [plane][2]. Ends with a CTA barrier.
PB_DEV void biquad_scan_planes64(const BiquadCoef& c, double (*ic)[2], double* const* planes, double* lane_state, uint32_t len, uint32_t tid) {
  const uint32_t ch = tid >> 7, j = tid & 127u, lane = tid & 31u, w = j >> 5;
  const double in1 = ic[ch][0], in2 = ic[ch][1];
  __syncthreads();  // every thread has read the incoming state before the owner of the last sample replaces it
  if (len == 0) return;
  const uint32_t B = (len + 127u) / 128u;
  const uint32_t lo = min(j * B, len), hi = min(lo + B, len);
  double* x = planes[ch];
  double* ls = lane_state + (size_t)ch * 64;
  const Mat2 A{2.0 * c.a1 - 1.0, -2.0 * c.a2, 2.0 * c.a2, 1.0 - 2.0 * c.a3};
  double e1 = 0.0, e2 = 0.0;
  for (uint32_t n = lo; n < hi; ++n) x[pidx(n)] = biquad_tick(c, e1, e2, x[pidx(n)]);   // zero-state response, in place
  const Mat2 P = mat_pow(A, B);
  Mat2 Pd = P;
  double E1 = e1, E2 = e2;
#pragma unroll
  for (uint32_t d = 1; d < 32; d <<= 1) {
    const double u1 = __shfl_up_sync(0xFFFFFFFFu, E1, d), u2 = __shfl_up_sync(0xFFFFFFFFu, E2, d);
    if (lane >= d) { E1 += Pd.a * u1 + Pd.b * u2; E2 += Pd.c * u1 + Pd.d * u2; }
    Pd = mat_mul(Pd, Pd);
  }
  if (lane == 31) { ls[w * 2] = E1; ls[w * 2 + 1] = E2; }
  __syncthreads();
  double W1 = in1, W2 = in2;
  for (uint32_t q = 0; q < w; ++q) {
    const double t1 = Pd.a * W1 + Pd.b * W2 + ls[q * 2], t2 = Pd.c * W1 + Pd.d * W2 + ls[q * 2 + 1];
    W1 = t1; W2 = t2;
  }
  double p1 = __shfl_up_sync(0xFFFFFFFFu, E1, 1), p2 = __shfl_up_sync(0xFFFFFFFFu, E2, 1);
  if (lane == 0) { p1 = 0.0; p2 = 0.0; }
  Mat2 Q = P;
#pragma unroll
  for (uint32_t k = 0; k < 5; ++k) {
    if ((lane >> k) & 1u) { const double t1 = Q.a * W1 + Q.b * W2, t2 = Q.c * W1 + Q.d * W2; W1 = t1; W2 = t2; }
    Q = mat_mul(Q, Q);
  }
  const double C1 = c.m1 * c.a1 + c.m2 * c.a2, C2 = -c.m1 * c.a2 + c.m2 * (1.0 - c.a3);
  double h1 = W1 + p1, h2 = W2 + p2;
  for (uint32_t n = lo; n < hi; ++n) {
    x[pidx(n)] += C1 * h1 + C2 * h2;
    const double t1 = A.a * h1 + A.b * h2, t2 = A.c * h1 + A.d * h2;
    h1 = t1; h2 = t2;
  }
  if (hi == len && lo < hi) { ic[ch][0] = h1 + e1; ic[ch][1] = h2 + e2; }
  __syncthreads();
}

// ---- FilterEffect::process (filter.rs:166-201): warps 0/1 = channels ---------------------------------------
PB_DEV void filter_process(FilterState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames, uint32_t lane, uint32_t warp) {
  const bool ramp = exp_need_ramp(s.cutoff, cx.comp) || lin_need_ramp(s.q);
  if (ramp) {
    // coefficients are shared by both channels and re-derived per frame: frame-serial on one thread
    if (lane == 0 && warp == 0) {
      for (uint32_t f = 0; f < frames; ++f) {
        float c = fminf(fmaxf(exp_next(s.cutoff, cx.comp), 20.0f), (float)cx.sample_rate / 2.0f);
        float q = lin_next(s.q);
        biquad_set(s.coef, s.filter_type, cx.sample_rate, c, q, 0.0f);
        CB_L(f) = (float)biquad_tick(s.coef, s.ic1[0], s.ic2[0], (double)CB_L(f));
        CB_R(f) = (float)biquad_tick(s.coef, s.ic1[1], s.ic2[1], (double)CB_R(f));
      }
    }
  } else {
    const BiquadCoef c = s.coef;
    biquad_scan_stereo(c, s.ic1, s.ic2, 1, cb, frames, warp * 32 + lane);
  }
}

// ---- Eq5Effect (eq5.rs:173-209, 297-326) ---------------------------------------------------------------
PB_DEV uint32_t eq5_band_type(int i) { return i == 0 ? BQ_LOWSHELF : (i == 4 ? BQ_HIGHSHELF : BQ_BELL); }
PB_DEV void eq5_update_coefficients(Eq5State& s, const FxCtx& cx) {
  for (int i = 0; i < 5; ++i) {
    float c = fminf(fmaxf(s.freqs[i].current, 20.0f), (float)cx.sample_rate / 2.0f);
    biquad_set(s.coef[i], eq5_band_type(i), cx.sample_rate, c, s.bws[i].current, s.gains[i].current);
  }
}
PB_DEV void eq5_process(Eq5State& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames, uint32_t lane, uint32_t warp) {
  bool ramp = false;
  for (int i = 0; i < 5; ++i) ramp |= exp_need_ramp(s.freqs[i], cx.comp) || lin_need_ramp(s.bws[i]) || exp_need_ramp(s.gains[i], cx.comp);
  if (ramp) {
    if (lane == 0 && warp == 0) {
      for (uint32_t f = 0; f < frames; ++f) {
        for (int i = 0; i < 5; ++i) {  // ramp_filter_coefficients
          float bw = lin_next(s.bws[i]);
          float q = (i == 0 || i == 4) ? bw : 1.0f / fmaxf(bw, 0.001f);
          float c = fminf(fmaxf(exp_next(s.freqs[i], cx.comp), 20.0f), (float)cx.sample_rate / 2.0f);
          float g = exp_next(s.gains[i], cx.comp);
          biquad_set(s.coef[i], eq5_band_type(i), cx.sample_rate, c, q, g);
        }
        for (int ch = 0; ch < 2; ++ch) {
          float x = cb.ch[ch][pidx(f)];
          for (int i = 0; i < 5; ++i) x = (float)biquad_tick(s.coef[i], s.ic1[ch][i], s.ic2[ch][i], (double)x);
          cb.ch[ch][pidx(f)] = x;
        }
      }
    }
  } else {
    // five cascaded stages, each cast back to f32 before the next one (eq5.rs:317-321)
    for (int i = 0; i < 5; ++i) {
      const BiquadCoef c = s.coef[i];
      biquad_scan_stereo(c, &s.ic1[0][i], &s.ic2[0][i], 5, cb, frames, warp * 32 + lane);
    }
  }
}

// ---- CompressorEffect::process (compressor.rs:230-294) + LookupDelayLine (delay.rs:206-265) -------------
PB_DEV void comp_process(CompState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames) {
  double* line = cx.aux_arena + s.aux;
  const bool limiter = s.ratio >= 20.0f;
  if (s.peak_dirty && s.delay_frames != 0) {  // window peak = max over the last delay_frames frames (delay.rs:245-260)
    s.peak_value = 0.0;
    for (uint32_t i = 1; i <= s.delay_frames; ++i) {
      uint32_t fi = (s.write_pos + s.buf_frames - i) & s.mask;
      double fp = fmax(fmax(0.0, fabs(line[fi * 2])), fabs(line[fi * 2 + 1]));
      if (fp >= s.peak_value) { s.peak_value = fp; s.peak_pos = fi; }
    }
  }
  s.peak_dirty = 0;
  for (uint32_t f = 0; f < frames; ++f) {
    const float in0 = CB_L(f), in1 = CB_R(f);
    float d0 = in0, d1 = in1;
    if (s.delay_frames != 0) {
      uint32_t read_index = (s.write_pos + s.buf_frames - s.delay_frames) & s.mask;
      d0 = (float)line[read_index * 2]; d1 = (float)line[read_index * 2 + 1];
      uint32_t wi = s.write_pos & s.mask;
      line[wi * 2] = (double)in0; line[wi * 2 + 1] = (double)in1;
      bool peak_expired = s.peak_pos == read_index;
      double new_peak = fmax(fmax(0.0, (double)fabsf(in0)), (double)fabsf(in1));
      if (new_peak >= s.peak_value) {
        s.peak_value = new_peak; s.peak_pos = s.write_pos;
      } else if (peak_expired) {
        s.peak_value = 0.0;
        for (uint32_t i = 0; i < s.delay_frames; ++i) {
          uint32_t fi = (s.write_pos + s.buf_frames - i) & s.mask;
          double fp = fmax(fmax(0.0, fabs(line[fi * 2])), fabs(line[fi * 2 + 1]));
          if (fp >= s.peak_value) { s.peak_value = fp; s.peak_pos = fi; }
        }
      }
      s.write_pos = (s.write_pos + 1) & s.mask;
    }
    float input_db;
    if (limiter) {
      float peak = (float)s.peak_value;
      input_db = peak > 1e-6f ? 20.0f * log10f(peak) : -120.0f;
    } else {
      float peak = fmaxf(fabsf(in0), fabsf(in1));
      input_db = peak > 1e-6f ? 20.0f * log10f(peak) : -120.0f;
    }
    if (input_db > s.env_cur) s.env_cur = input_db + s.atk_coeff * (s.env_cur - input_db);
    else s.env_cur = input_db + s.rel_coeff * (s.env_cur - input_db);
    const float envelope = s.env_cur;
    const float t = s.threshold, w = s.knee;
    const float slope = limiter ? 1.0f : 1.0f - 1.0f / s.ratio;
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
    float makeup = exp_next(s.makeup, cx.comp);
    float total_gain = db_to_linear_dev(makeup - gr_db);
    CB_L(f) = d0 * total_gain;
    CB_R(f) = d1 * total_gain;
  }
}

// ---- ChorusEffect::process (chorus.rs:311-394) -----------------------------------------------------------
PB_DEV void chorus_process(ChorusState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames) {
  double* bl = cx.aux_arena + s.dl.aux;
  double* br = cx.aux_arena + s.dr.aux;
  const float srf = (float)cx.sample_rate;
  for (uint32_t f = 0; f < frames; ++f) {
    const float li = CB_L(f), ri = CB_R(f);
    float delay_ms = spring_next(s.delay, cx.comp);
    float depth = exp_next(s.depth, cx.comp);
    float fb = fminf(fmaxf(exp_next(s.feedback, cx.comp), -0.999f), 0.999f);
    float wet = exp_next(s.wet, cx.comp);
    float dry = 1.0f - wet;
    if (lin_need_ramp(s.rate) || lin_need_ramp(s.phase)) {  // update_lfos
      double r = (double)lin_next(s.rate);
      lfo_set_rate(s.left_osc, cx.sample_rate, r);
      lfo_set_rate(s.right_osc, cx.sample_rate, r);
      double off = (double)lin_next(s.phase);
      lfo_set_phase_degrees(s.left_osc, (float)s.current_phase);
      lfo_set_phase_degrees(s.right_osc, (float)(s.current_phase + off));
    }
    if (exp_need_ramp(s.filter_freq, cx.comp) || exp_need_ramp(s.filter_res, cx.comp)) {
      float c = fminf(fmaxf(exp_next(s.filter_freq, cx.comp), 20.0f), srf / 2.0f);
      float r = exp_next(s.filter_res, cx.comp);
      svf_set(s.coef, s.filter_type, cx.sample_rate, c, r);
    }
    double fl = svf_tick(s.coef, s.fl_ic1, s.fl_ic2, (double)li);
    double fr = svf_tick(s.coef, s.fr_ic1, s.fr_ic2, (double)ri);
    float delay_in_samples = delay_ms * srf * 0.001f;
    float depth_in_samples = s.lfo_range * depth;
    float llfo = lfo_run(s.left_osc), rlfo = lfo_run(s.right_osc);
    float lpos = 2.0f + delay_in_samples + (1.0f + llfo) * depth_in_samples;
    float rpos = 2.0f + delay_in_samples + (1.0f + rlfo) * depth_in_samples;
    float lo = idelay_process(s.dl, bl, (float)fl, fb, lpos);
    float ro = idelay_process(s.dr, br, (float)fr, fb, rpos);
    CB_L(f) = li * dry + lo * wet;
    CB_R(f) = ri * dry + ro * wet;
  }
  const double PI = 3.14159265358979323846;
  double phase_inc = 2.0 * PI * (double)s.rate.current / (double)cx.sample_rate;
  s.current_phase += (double)(frames * 2) / 2.0 * phase_inc;
  while (s.current_phase >= 2.0 * PI) s.current_phase -= 2.0 * PI;
}

// ---- DelayEffect::process (delay.rs:334-454) ------------------------------------------------------------
PB_DEV double delay_saturate(double input, float drive) {
  if (drive < 0.001f) return input;
  double gain = 1.0 + (double)drive * 4.0;
  double x = input * gain;
  double x2 = x * x;
  double o = x * (27.0 + x2) / (27.0 + 9.0 * x2);
  return o / sqrt(gain);
}
PB_DEV float delay_feedback_path(const SvfCoef& c, double& ic1, double& ic2, double& x1, double& y1, double r, float delayed, float drive) {
  double filtered = svf_tick(c, ic1, ic2, (double)delayed);
  double sat = delay_saturate(filtered, drive);
  y1 = sat - x1 + r * y1;  // DcFilter::process_sample (dc.rs:84-88)
  x1 = sat;
  float clean = (float)y1;
  return fminf(fmaxf(clean, -4.0f), 4.0f);
}
PB_DEV void delay_process(DelayState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames) {
  double* bl = cx.aux_arena + s.dl.aux;
  double* br = cx.aux_arena + s.dr.aux;
  const float srf = (float)cx.sample_rate;
  for (uint32_t f = 0; f < frames; ++f) {
    const float li = CB_L(f), ri = CB_R(f);
    float lfo_val = lfo_run(s.lfo);
    if (exp_need_ramp(s.lfo_rate, cx.comp)) { float r = exp_next(s.lfo_rate, cx.comp); lfo_set_rate(s.lfo, cx.sample_rate, (double)r); }
    float base_delay_ms = spring_next(s.delay_time, cx.comp);
    float time_mod_ms = lfo_val * exp_next(s.lfo_dt, cx.comp) * 50.0f;
    float delay_ms = fmaxf(base_delay_ms + time_mod_ms, 1.0f);
    float delay_samples = delay_ms * 0.001f * srf;
    float filter_depth = exp_next(s.lfo_dflt, cx.comp);
    float filter_mod = powf(2.0f, lfo_val * filter_depth * 2.0f);
    float c = fminf(fmaxf(exp_next(s.cutoff, cx.comp) * filter_mod, 20.0f), srf / 2.0f);
    svf_set(s.coef, s.filter_type, cx.sample_rate, c, 0.302f);
    float base_fb = exp_next(s.feedback, cx.comp);
    float fb_depth = exp_next(s.lfo_dfb, cx.comp);
    float fb = fminf(fmaxf(base_fb + lfo_val * fb_depth * (1.0f - fabsf(base_fb)), 0.0f), 0.999f);
    float drive = exp_next(s.drive, cx.comp);
    float wet = exp_next(s.wet, cx.comp);
    float width = exp_next(s.width, cx.comp);
    float wet_l, wet_r;
    if (s.mode == 0) {
      float l_in = li + s.fb_l * fb;
      float dl = idelay_process(s.dl, bl, l_in, 0.0f, delay_samples);
      float cl = delay_feedback_path(s.coef, s.fl_ic1, s.fl_ic2, s.dcl_x1, s.dcl_y1, s.dc_r, dl, drive);
      s.fb_l = cl;
      float r_in = ri + s.fb_r * fb;
      float dr = idelay_process(s.dr, br, r_in, 0.0f, delay_samples);
      float cr = delay_feedback_path(s.coef, s.fr_ic1, s.fr_ic2, s.dcr_x1, s.dcr_y1, s.dc_r, dr, drive);
      s.fb_r = cr;
      wet_l = cl; wet_r = cr;
    } else {
      float mono = (li + ri) * 0.5f;
      float l_in = mono + s.fb_r * fb;
      float dl = idelay_process(s.dl, bl, l_in, 0.0f, delay_samples);
      float cl = delay_feedback_path(s.coef, s.fl_ic1, s.fl_ic2, s.dcl_x1, s.dcl_y1, s.dc_r, dl, drive);
      float r_in = s.fb_l * fb;
      float dr = idelay_process(s.dr, br, r_in, 0.0f, delay_samples);
      float cr = delay_feedback_path(s.coef, s.fr_ic1, s.fr_ic2, s.dcr_x1, s.dcr_y1, s.dc_r, dr, drive);
      s.fb_l = cl; s.fb_r = cr;
      wet_l = cl; wet_r = cr;
    }
    float dry_gain = fminf((1.0f - wet) * 2.0f, 1.0f);
    float wet_gain = fminf(wet * 2.0f, 1.0f);
    float ol = li * dry_gain + wet_l * wet_gain;
    float orr = ri * dry_gain + wet_r * wet_gain;
    float mid = (ol + orr) * 0.5f;
    float side = (ol - orr) * 0.5f;
    CB_L(f) = mid + side * width;
    CB_R(f) = mid - side * width;
  }
}

// ---- ReverbEffect (reverb.rs:196-369, 409-447, 554-604) ---------------------------------------------------
struct RvDerived { float cutoff; double size, blend, regen; };
PB_DEV RvDerived reverb_derive(double room, double w) {
  RvDerived d;
  d.cutoff = (float)(10000.0 - (room * w * 3000.0));
  d.size = (room * room * 75.0) + 25.0;
  double t = 1.0 - (0.82 - (((1.0 - room) * 0.7) + (d.size * 0.002)));
  double depth_factor = 1.0 - (t * t) * (t * t);
  d.blend = 0.955 - (d.size * 0.007);
  d.regen = depth_factor * 0.5;
  return d;
}
PB_DEV uint32_t f64_as_usize32(double v) { return v > 0.0 ? (uint32_t)v : 0u; }
PB_DEV uint32_t reverb_update_sizes(ReverbState& s, double size) {
  const double mult[8] = {79.0, 73.0, 71.0, 67.0, 61.0, 59.0, 53.0, 47.0};
  for (int i = 0; i < 8; ++i) s.lines[i].delay = min(f64_as_usize32(mult[i] * size), s.lines[i].size);  // buffer.len()-1 == size
  const double am[4] = {43.0, 41.0, 37.0, 31.0};
  for (int i = 0; i < 4; ++i) s.ap[i].delay = min(f64_as_usize32(am[i] * size), s.ap[i].size - 1);
  return f64_as_usize32(29.0 * size);
}
PB_DEV void reverb_update_filters(ReverbState& s, const FxCtx& cx, float cutoff) {
  float c = fminf(fmaxf(cutoff, 20.0f), (float)cx.sample_rate / 2.0f);
  biquad_set(s.ca, BQ_LOWPASS, cx.sample_rate, c, 1.618034f, 0.0f);
  biquad_set(s.cb, BQ_LOWPASS, cx.sample_rate, c, 0.618034f, 0.0f);
  biquad_set(s.cc, BQ_LOWPASS, cx.sample_rate, c, 0.5f, 0.0f);
}
PB_DEV void rv_allpass(RvAllpass& a, double* __restrict__ b, double& l, double& r) {  // delay.rs:314-350
  uint32_t read_pos = a.write_pos + 1;
  if (read_pos > a.delay) read_pos = 0;
  double dl = b[read_pos * 2], dr = b[read_pos * 2 + 1];
  double bl = l - (dl * 0.5), br = r - (dr * 0.5);
  double ol = bl * 0.5, orr = br * 0.5;
  b[a.write_pos * 2] = bl; b[a.write_pos * 2 + 1] = br;
  a.write_pos += 1;
  if (a.write_pos > a.delay) a.write_pos = 0;
  ol += b[a.write_pos * 2]; orr += b[a.write_pos * 2 + 1];
  l = ol; r = orr;
}
PB_DEV void reverb_frame(ReverbState& s, const FxCtx& cx, float& frame0, float& frame1, double blend, double regen, uint32_t predelay, double w) {
  const double vib_speed = 0.1, vib_depth = 7.0;
  double il = (double)frame0, ir = (double)frame1;
  if (fabs(il) < 1.18e-23) il = (double)s.fpd_l * 1.18e-17;
  if (fabs(ir) < 1.18e-23) ir = (double)s.fpd_r * 1.18e-17;
  const double dry_l = il, dry_r = ir;
  {  // DelayLine<2>::process (delay.rs:47-66)
    double* m = cx.aux_arena + s.m_aux;
    s.m_write_pos &= s.m_mask;
    m[s.m_write_pos * 2] = il; m[s.m_write_pos * 2 + 1] = ir;
    s.m_write_pos = (s.m_write_pos + 1) & s.m_mask;
    if (s.m_write_pos > predelay) s.m_write_pos = 0;
    il = m[s.m_write_pos * 2]; ir = m[s.m_write_pos * 2 + 1];
  }
  il = biquad_tick(s.ca, s.a_ic[0][0], s.a_ic[0][1], il);
  ir = biquad_tick(s.ca, s.a_ic[1][0], s.a_ic[1][1], ir);
  il *= w; ir *= w;
  il = sin(il); ir = sin(ir);
  double ap[4][2];
  double xl = il, xr = ir;
  for (int i = 0; i < 4; ++i) {
    rv_allpass(s.ap[i], cx.aux_arena + s.ap[i].aux, xl, xr);
    ap[i][0] = xl; ap[i][1] = xr;
  }
  // a<-l, b<-k, c<-j, d<-i, e<-i, f<-j, g<-k, h<-l (reverb.rs:276-283)
  const int src[8] = {3, 2, 1, 0, 0, 1, 2, 3};
  double o[8][2];
  for (int i = 0; i < 8; ++i) {
    RvLine& L = s.lines[i];
    double* b = cx.aux_arena + L.aux;
    b[L.count * 2] = ap[src[i]][0] + L.feedback[0];
    b[L.count * 2 + 1] = ap[src[i]][1] + L.feedback[1];
    L.count += 1;
    if (L.count > L.delay) L.count = 0;
    L.vib_phase[0] += L.depth * vib_speed;
    L.vib_phase[1] += L.depth * vib_speed;
    for (int ch = 0; ch < 2; ++ch) {
      double offset = (sin(L.vib_phase[ch]) + 1.0) * vib_depth;
      double working = (double)L.count + offset;
      double wf = floor(working);
      double frac = working - wf;
      uint32_t wi = (uint32_t)wf;
      uint32_t r1 = wi; if (r1 > L.delay) r1 -= L.delay + 1;
      uint32_t r2 = wi + 1; if (r2 > L.delay) r2 -= L.delay + 1;
      double v1 = b[r1 * 2 + ch], v2 = b[r2 * 2 + ch];
      double ip = v1 * (1.0 - frac) + v2 * frac;
      ip = (1.0 - blend) * ip + (v1 * blend);
      o[i][ch] = ip;
    }
  }
  for (int ch = 0; ch < 2; ++ch) {
    double A = o[0][ch], B = o[1][ch], C = o[2][ch], D = o[3][ch], E = o[4][ch], F = o[5][ch], G = o[6][ch], H = o[7][ch];
    s.lines[0].feedback[ch] = (A - (B + C + D)) * regen;
    s.lines[1].feedback[ch] = (B - (A + C + D)) * regen;
    s.lines[2].feedback[ch] = (C - (A + B + D)) * regen;
    s.lines[3].feedback[ch] = (D - (A + B + C)) * regen;
    s.lines[4].feedback[ch] = (E - (F + G + H)) * regen;
    s.lines[5].feedback[ch] = (F - (E + G + H)) * regen;
    s.lines[6].feedback[ch] = (G - (E + F + H)) * regen;
    s.lines[7].feedback[ch] = (H - (E + F + G)) * regen;
  }
  il = (o[0][0] + o[1][0] + o[2][0] + o[3][0] + o[4][0] + o[5][0] + o[6][0] + o[7][0]) / 8.0;
  ir = (o[0][1] + o[1][1] + o[2][1] + o[3][1] + o[4][1] + o[5][1] + o[6][1] + o[7][1]) / 8.0;
  il = biquad_tick(s.cb, s.b_ic[0][0], s.b_ic[0][1], il);
  ir = biquad_tick(s.cb, s.b_ic[1][0], s.b_ic[1][1], ir);
  il = fmin(fmax(il, -1.0), 1.0);
  ir = fmin(fmax(ir, -1.0), 1.0);
  il = asin(il); ir = asin(ir);
  il = biquad_tick(s.cc, s.c_ic[0][0], s.c_ic[0][1], il);
  ir = biquad_tick(s.cc, s.c_ic[1][0], s.c_ic[1][1], ir);
  if (w != 1.0) { il += dry_l * (1.0 - w); ir += dry_r * (1.0 - w); }
  frame0 = (float)il; frame1 = (float)ir;
}
PB_DEV void reverb_process(ReverbState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames) {
  if (lin_need_ramp(s.room) || exp_need_ramp(s.wet, cx.comp)) {
    for (uint32_t f = 0; f < frames; ++f) {
      double room = (double)lin_next(s.room);
      double w = (double)exp_next(s.wet, cx.comp);
      RvDerived d = reverb_derive(room, w);
      uint32_t predelay = reverb_update_sizes(s, d.size);
      reverb_update_filters(s, cx, d.cutoff);
      reverb_frame(s, cx, CB_L(f), CB_R(f), d.blend, d.regen, predelay, w);
    }
  } else {
    double room = (double)s.room.target;
    double w = (double)s.wet.target;
    RvDerived d = reverb_derive(room, w);
    uint32_t predelay = reverb_update_sizes(s, d.size);
    reverb_update_filters(s, cx, d.cutoff);
    for (uint32_t f = 0; f < frames; ++f) reverb_frame(s, cx, CB_L(f), CB_R(f), d.blend, d.regen, predelay, w);
  }
}

// ---- GainEffect::process (gain.rs:148-170): DC filter per channel (serial one-pole, warps 0/1), then the gain -----
PB_DEV void gain_process(GainState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames, uint32_t tid, uint32_t nt) {
  if (s.dc_mode != 0) {
    if (tid == 0 || tid == 32) {  // DcFilter::process_sample (dc.rs:84-88), f64 state, one thread per channel
      const uint32_t ch = tid >> 5;
      double x1 = s.dc_x1[ch], y1 = s.dc_y1[ch];
      const double r = s.dc_r;
      // the chain is two dependent f64 operations per frame: loads / conversions / stores of 8 frames are issued around it
      uint32_t f = 0;
      for (; f + 8 <= frames; f += 8) {
        double x[8];
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (double)cb.ch[ch][pidx(f + j)];
#pragma unroll
        for (int j = 0; j < 8; ++j) { y1 = x[j] - x1 + r * y1; x1 = x[j]; y[j] = (float)y1; }
#pragma unroll
        for (int j = 0; j < 8; ++j) cb.ch[ch][pidx(f + j)] = y[j];
      }
      for (; f < frames; ++f) {
        const double x = (double)cb.ch[ch][pidx(f)];
        y1 = x - x1 + r * y1;
        x1 = x;
        cb.ch[ch][pidx(f)] = (float)y1;
      }
      s.dc_x1[ch] = x1; s.dc_y1[ch] = y1;
    }
    __syncthreads();
  }
  if (exp_need_ramp(s.gain, cx.comp)) {
    if (tid == 0) {
      for (uint32_t f = 0; f < frames; ++f) {
        const float g = exp_next(s.gain, cx.comp);
        CB_L(f) *= g; CB_R(f) *= g;
      }
    }
  } else {
    const float g = s.gain.target;  // scale_buffer
    for (uint32_t f = tid; f < frames; f += nt) { CB_L(f) *= g; CB_R(f) *= g; }
  }
}

// ---- GateEffect::process (gate.rs:152-198) -----------------------------------------------------------------------
// Detector level (log10) and the final gain (exp) are per-frame parallel stages around the serial part: the
// envelope follower, the open / hold / closed state machine and the one-pole gain smoothing, one thread.
PB_DEV void gate_process(GateState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames, uint32_t tid, uint32_t nt, float* work) {
  float* level = work;            // input_db, then gate_gain_db per frame
  for (uint32_t f = tid; f < frames; f += nt) {
    const float frame_peak = fmaxf(fabsf(CB_L(f)), fabsf(CB_R(f)));
    level[f] = frame_peak > 1e-6f ? 20.0f * log10f(frame_peak) : -120.0f;
  }
  __syncthreads();
  if (tid == 0) {
    const float thr = s.threshold, range_db = s.range;
    const float hs = s.hold_time * (float)cx.sample_rate;
    const uint32_t hold_samples = hs >= 4294967295.0f ? 0xFFFFFFFFu : (hs > 0.0f ? (uint32_t)hs : 0u);
    float cur = s.env_cur, g = s.gate_gain_db;
    uint32_t hold = s.hold_counter;
    const float ea = s.env_atk, er = s.env_rel, ac = s.attack_coeff, rc = s.release_coeff;
    auto step = [&](const float x) {
      cur = x + (x > cur ? ea : er) * (cur - x);   // EnvelopeFollower::run (envelope.rs:51-60)
      float target;
      if (cur >= thr) { hold = hold_samples; target = 0.0f; }
      else if (hold > 0) { hold -= 1; target = 0.0f; }
      else target = range_db;
      if (target > g) g = ac * g + (1.0f - ac) * target;
      else g = rc * g + (1.0f - rc) * target;
      return g;
    };
    uint32_t f = 0;
    for (; f + 8 <= frames; f += 8) {  // loads and stores of 8 frames around the two dependent chains
      float x[8], y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = level[f + j];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = step(x[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) level[f + j] = y[j];
    }
    for (; f < frames; ++f) level[f] = step(level[f]);
    s.env_cur = cur; s.gate_gain_db = g; s.hold_counter = hold;
  }
  __syncthreads();
  for (uint32_t f = tid; f < frames; f += nt) {
    const float g = level[f];
    const float gain = g <= -60.0f ? 0.0f : db_to_linear_dev(g);
    CB_L(f) *= gain; CB_R(f) *= gain;
  }
}

// ---- PanningEffect::process (pan.rs:105-160) ----------------------------------------------------------------------
PB_DEV void pan_process(PanState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames, uint32_t tid, uint32_t nt) {
  const float il = s.invert_l ? -1.0f : 1.0f, ir = s.invert_r ? -1.0f : 1.0f;
  const bool has_invert = il < 0.0f || ir < 0.0f;
  const bool pan_ramping = exp_need_ramp(s.pan, cx.comp), width_ramping = exp_need_ramp(s.width, cx.comp);
  if (!has_invert && !pan_ramping && !width_ramping && fabsf(s.pan.target) < 1e-6f && fabsf(s.width.target - 1.0f) < 1e-6f) return;
  auto frame = [&](uint32_t f, float w, float p) {
    float l = CB_L(f) * il, r = CB_R(f) * ir;
    if (fabsf(w - 1.0f) > 1e-6f) {
      const float mid = (l + r) * 0.5f, side = (l - r) * 0.5f;
      l = mid + side * w;
      r = mid - side * w;
    }
    if (fabsf(p) > 1e-6f) {
      float pl, pr;
      panning_factors(p, pl, pr);
      l *= pl; r *= pr;
    }
    CB_L(f) = l; CB_R(f) = r;
  };
  if (pan_ramping || width_ramping) {
    if (tid == 0)
      for (uint32_t f = 0; f < frames; ++f) {
        const float w = width_ramping ? exp_next(s.width, cx.comp) : s.width.target;
        const float p = pan_ramping ? exp_next(s.pan, cx.comp) : s.pan.target;
        frame(f, w, p);
      }
  } else {
    const float w = s.width.target, p = s.pan.target;
    for (uint32_t f = tid; f < frames; f += nt) frame(f, w, p);
  }
}

// ---- DistortionEffect (distortion.rs:69-190, 326-360): memoryless waveshaping, parallel over frames unless a smoother ramps
constexpr float DIST_MIX_INERTIA = 0.1f;  // ExponentialSmoothedValue::with_inertia(0.1) (distortion.rs:238)
PB_DEV bool dist_mix_need_ramp(const ExpSm& s, float comp) {
  return fabsf((s.target - s.current) * DIST_MIX_INERTIA * comp) > F32_EPS * 100.0f;
}
PB_DEV float dist_mix_next(ExpSm& s, float comp) {
  const float add = (s.target - s.current) * DIST_MIX_INERTIA * comp;
  if (fabsf(add) > F32_EPS * 100.0f) { s.current += add; return s.current; }
  return s.target;
}
PB_DEV float dist_shape(uint32_t type, float sample, float drive) {
  const float t = drive / 4.0f;
  switch (type) {
    case 0: {  // SoftClip
      const float gain = 1.0f + (t * t) * (15.0f - 1.0f);
      const float x = sample * gain;
      if (x >= 1.0f) return 1.0f;
      if (x > -1.0f) return gain <= 1.0f ? sample : (3.0f / 2.0f) * (x - ((x * x) * x) / 3.0f);
      return -1.0f;
    }
    case 1: {  // HardClip
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
      const float diode_clipping = expf((0.1f * sample) / (0.0253f * 1.68f)) - 1.0f;
      return 2.0f / 3.14159265358979323846f * atanf(diode_clipping * gain);
    }
    case 3: {  // Fuzz
      const float gain = 1.0f + (1.0f - expf(-3.0f * t)) * (30.0f - 1.0f);
      const float amplified = sample * gain;
      const float sat = amplified < 0.0f ? -1.0f * (1.0f - expf(-fabsf(amplified))) : 1.0f * (1.0f - expf(-fabsf(amplified)));
      return 1.5f * (sat + fabsf(sat));
    }
    default: {  // Fold
      const float gain = 1.0f + (t * t) * (4.0f - 1.0f);
      const float x = sample * gain;
      const float threshold = 1.0f / gain;
      if (x > threshold || x < -threshold) return fabsf(fmodf(fabsf(x - threshold), threshold * 4.0f) - threshold * 2.0f) - threshold;
      return x;
    }
  }
}
PB_DEV float dist_lookup(const DistState& s, float drive) {  // lookup_gain_compensation (distortion.rs:271-279)
  const float* lut = s.lut[s.type];
  const float pos = fminf(fmaxf(drive / 4.0f, 0.0f), 1.0f) * 255.0f;
  const uint32_t lo = (uint32_t)pos;
  const uint32_t hi = min(lo + 1u, 255u);
  const float frac = pos - (float)lo;
  return lut[lo] + (lut[hi] - lut[lo]) * frac;
}
PB_DEV void dist_process(DistState& s, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames, uint32_t tid, uint32_t nt) {
  const bool mix_ramp = dist_mix_need_ramp(s.mix, cx.comp), drive_ramp = lin_need_ramp(s.drive);
  const float mix_target = s.mix.target;
  const uint32_t type = s.type;
  __syncthreads();  // every thread has taken its decisions before thread 0 advances the smoothers
  if (!mix_ramp && mix_target == 0.0f) return;
  if (!mix_ramp && mix_target >= 1.0f) {
    if (!drive_ramp) {
      const float d = s.drive.target, comp = dist_lookup(s, d);
      for (uint32_t f = tid; f < frames; f += nt) { CB_L(f) = dist_shape(type, CB_L(f), d) * comp; CB_R(f) = dist_shape(type, CB_R(f), d) * comp; }
    } else if (tid == 0) {
      for (uint32_t f = 0; f < frames; ++f) {
        const float d = lin_next(s.drive), comp = dist_lookup(s, d);
        CB_L(f) = dist_shape(type, CB_L(f), d) * comp; CB_R(f) = dist_shape(type, CB_R(f), d) * comp;
      }
    }
  } else if (tid == 0) {
    for (uint32_t f = 0; f < frames; ++f) {
      const float d = lin_next(s.drive), comp = dist_lookup(s, d);
      const float m = dist_mix_next(s.mix, cx.comp);
      const float dl = CB_L(f), dr = CB_R(f);
      CB_L(f) = (1.0f - m) * dl + m * (dist_shape(type, dl, d) * comp);
      CB_R(f) = (1.0f - m) * dr + m * (dist_shape(type, dr, d) * comp);
    }
  }
}

// ---- Effect::process_tail (Option<usize>): returns false for None -------------------------------------------
PB_DEV uint64_t f32_ceil_u64(float v) { float c = ceilf(v); return c > 0.0f ? (uint64_t)c : 0ull; }
PB_DEV bool fx_process_tail(const FxHeader& h, const FxCtx& cx, uint64_t& frames) {
  const uint8_t* st = cx.state_arena + h.state_offset;
  const float srf = (float)cx.sample_rate;
  switch (h.kind) {
    case FX_FILTER: frames = cx.sample_rate / 10; return true;
    case FX_EQ5: frames = cx.sample_rate / 5; return true;
    case FX_GAIN: {
      const GainState& s = *(const GainState*)st;
      frames = s.dc_mode != 0 ? (uint64_t)cx.sample_rate / (s.dc_mode == 1 ? 1u : (s.dc_mode == 3 ? 20u : 5u)) : 0ull;
      return true;
    }
    case FX_PANNING: frames = 0; return true;
    case FX_DISTORTION: frames = 0; return true;
    case FX_GATE: {
      const GateState& s = *(const GateState*)st;
      frames = f32_ceil_u64(s.hold_time * srf) + f32_ceil_u64(s.release_time * srf);
      return true;
    }
    case FX_COMPRESSOR: {
      const CompState& s = *(const CompState*)st;
      frames = f32_ceil_u64(s.lookahead_time * srf) + f32_ceil_u64(s.release_time * srf);
      return true;
    }
    case FX_CHORUS: {
      const ChorusState& s = *(const ChorusState*)st;
      float total_ms = s.delay.target + 256.0f * 1000.0f / srf;
      float fb = fabsf(s.feedback.target);
      if (fb >= 1.0f) { frames = UINT64_MAX; return true; }
      if (fb < 0.001f) { frames = f32_ceil_u64(total_ms * srf / 1000.0f); return true; }
      float total_samples = total_ms * srf / 1000.0f;
      float decay = total_samples + (float)((double)total_samples * log10(0.001) / log10((double)fb));
      frames = f32_ceil_u64(decay);
      return true;
    }
    case FX_DELAY: {
      const DelayState& s = *(const DelayState*)st;
      if (s.drive.target > 0.0f) return false;
      double delay_ms = (double)(s.delay_time.target + 50.0f);
      double fb = (double)fabsf(s.feedback.target);
      if (fb >= 0.9999) { frames = UINT64_MAX; return true; }
      if (fb < 0.001) { frames = (uint64_t)ceil(delay_ms * (double)cx.sample_rate / 1000.0); return true; }
      double ds = delay_ms * (double)cx.sample_rate / 1000.0;
      double decay = ds + ds * log10(0.001) / log10(fb);
      uint64_t v = (uint64_t)ceil(decay);
      frames = v > 1 ? v : 1;
      return true;
    }
    case FX_REVERB: {
      const ReverbState& s = *(const ReverbState*)st;
      double room = (double)s.room.target;
      double size = (room * room * 75.0) + 25.0;
      uint64_t max_delay = (uint64_t)(79.0 * size);
      double t = 1.0 - (0.82 - (((1.0 - room) * 0.7) + (size * 0.002)));
      double fb = 1.0 - (t * t) * (t * t);
      if (fb >= 1.0) { frames = UINT64_MAX; return true; }
      if (fb == 0.0) { frames = max_delay; return true; }
      double extra = (double)max_delay * log10(0.001) / log10(fb);
      frames = max_delay + (extra > 0.0 ? (uint64_t)extra : 0ull);
      return true;
    }
  }
  return false;
}

// Effect::process_parameter_update with the value already resolved to a plain value on the host
// (normalized -> denormalized, clamped; enums -> index).
// Effect::process_message, cooperatively by the mixer's CTA. ReverbEffectMessage::Reset (src/effect/reverb.rs:469-487):
// flush() of the eight ReverbDelayLines (buffer only), the four allpasses and the predelay (buffer + write position).
PB_DEV void fx_process_message(FxHeader& h, const FxCtx& cx, const uint32_t msg, const uint32_t tid, const uint32_t nt) {
  if (h.kind != FX_REVERB || msg != 1u) return;
  ReverbState& s = *(ReverbState*)(cx.state_arena + h.state_offset);
  for (int l = 0; l < 8; ++l) {
    double* b = cx.aux_arena + s.lines[l].aux;
    for (uint32_t i = tid; i < (s.lines[l].size + 1u) * 2u; i += nt) b[i] = 0.0;
  }
  for (int l = 0; l < 4; ++l) {
    double* b = cx.aux_arena + s.ap[l].aux;
    for (uint32_t i = tid; i < s.ap[l].size * 2u; i += nt) b[i] = 0.0;
  }
  {
    double* b = cx.aux_arena + s.m_aux;
    for (uint32_t i = tid; i < (s.m_mask + 1u) * 2u; i += nt) b[i] = 0.0;
  }
  __syncthreads();
  if (tid == 0) { for (int l = 0; l < 4; ++l) s.ap[l].write_pos = 0; s.m_write_pos = 0; }
}

PB_DEV void fx_apply_param(FxHeader& h, const FxCtx& cx, const FxParamEvent& e) {
  uint8_t* st = cx.state_arena + h.state_offset;
  const uint32_t id = e.param_id;
  const float v = e.value;
#define CC4(a, b, c, d) (((uint32_t)(a) << 24) | ((uint32_t)(b) << 16) | ((uint32_t)(c) << 8) | (uint32_t)(d))
  switch (h.kind) {
    case FX_GAIN: {
      GainState& s = *(GainState*)st;
      if (id == CC4('g', 'a', 'i', 'n')) exp_set_target(s.gain, v, cx.comp);
      else if (id == CC4('d', 'c', 'f', 'm')) {
        s.dc_mode = (uint32_t)v & 3u;
        if (s.dc_mode != 0) s.dc_r = 1.0 - (6.28318530717958647692 * (s.dc_mode == 1 ? 1.0 : (s.dc_mode == 3 ? 20.0 : 5.0)) / (double)cx.sample_rate);
        else { s.dc_x1[0] = s.dc_x1[1] = 0.0; s.dc_y1[0] = s.dc_y1[1] = 0.0; }
      }
      break;
    }
    case FX_GATE: {
      GateState& s = *(GateState*)st;
      if (id == CC4('t', 'h', 'r', 's')) s.threshold = v;
      else if (id == CC4('a', 't', 't', 'k')) s.attack_time = v;
      else if (id == CC4('h', 'o', 'l', 'd')) s.hold_time = v;
      else if (id == CC4('r', 'e', 'l', 's')) s.release_time = v;
      else if (id == CC4('r', 'n', 'g', 'e')) s.range = v;
      const float srf = (float)cx.sample_rate;  // update_coefficients (gate.rs:83-95)
      s.env_atk = s.attack_time > 0.0f ? expf_host_rounding(-1.0f / (s.attack_time * srf)) : 0.0f;
      s.env_rel = s.release_time > 0.0f ? expf_host_rounding(-1.0f / (s.release_time * srf)) : 0.0f;
      s.attack_coeff = expf_host_rounding(-1.0f / (s.attack_time * srf));
      s.release_coeff = expf_host_rounding(-1.0f / (s.release_time * srf));
      break;
    }
    case FX_DISTORTION: {
      DistState& s = *(DistState*)st;
      if (id == CC4('t', 'y', 'p', 'e')) s.type = min((uint32_t)v, 4u);
      else if (id == CC4('d', 'r', 'i', 'v')) lin_set_target(s.drive, v, cx.comp);
      else if (id == CC4('m', 'i', 'x', ' ')) { s.mix.target = v; if (!dist_mix_need_ramp(s.mix, cx.comp)) s.mix.current = v; }
      break;
    }
    case FX_PANNING: {
      PanState& s = *(PanState*)st;
      if (id == CC4('p', 'a', 'n', ' ')) exp_set_target(s.pan, v, cx.comp);
      else if (id == CC4('w', 'd', 't', 'h')) exp_set_target(s.width, v, cx.comp);
      else if (id == CC4('i', 'n', 'v', 'l')) s.invert_l = v != 0.0f;
      else if (id == CC4('i', 'n', 'v', 'r')) s.invert_r = v != 0.0f;
      break;
    }
    case FX_FILTER: {
      FilterState& s = *(FilterState*)st;
      if (id == CC4('t', 'y', 'p', 'e')) {
        const uint32_t map[4] = {BQ_LOWPASS, BQ_BANDPASS, BQ_NOTCH, BQ_HIGHPASS};
        s.filter_type = map[(uint32_t)v & 3];
        if (s.coef.type != s.filter_type) { s.coef.type = s.filter_type; biquad_apply(s.coef); }
      } else if (id == CC4('c', 'u', 't', 'o')) exp_set_target(s.cutoff, v, cx.comp);
      else if (id == CC4('f', 'l', 't', 'q')) lin_set_target(s.q, v, cx.comp);
      break;
    }
    case FX_EQ5: {
      Eq5State& s = *(Eq5State*)st;
      const uint32_t band = (id & 0xFF) - '1';
      const uint32_t pre = id & 0xFFFFFF00u;
      if (band < 5) {
        if (pre == (CC4('g', 'a', 'n', 0))) exp_set_target(s.gains[band], v, cx.comp);
        else if (pre == (CC4('f', 'r', 'q', 0))) exp_set_target(s.freqs[band], v, cx.comp);
        else if (pre == (CC4('b', 'w', '_', 0))) lin_set_target(s.bws[band], v, cx.comp);
      }
      eq5_update_coefficients(s, cx);
      break;
    }
    case FX_COMPRESSOR: {
      CompState& s = *(CompState*)st;
      const float old_look = s.lookahead_time;
      if (id == CC4('t', 'h', 'r', 's')) s.threshold = v;
      else if (id == CC4('r', 'a', 't', 'o')) s.ratio = v;
      else if (id == CC4('k', 'n', 'e', 'e')) s.knee = v;
      else if (id == CC4('a', 't', 't', 'k')) s.attack_time = v;
      else if (id == CC4('r', 'e', 'l', 's')) s.release_time = v;
      else if (id == CC4('g', 'a', 'i', 'n')) exp_set_target(s.makeup, v, cx.comp);
      else if (id == CC4('l', 'o', 'o', 'k')) s.lookahead_time = v;
      s.atk_coeff = s.attack_time > 0.0f ? expf_host_rounding(-1.0f / (s.attack_time * (float)cx.sample_rate)) : 0.0f;
      s.rel_coeff = s.release_time > 0.0f ? expf_host_rounding(-1.0f / (s.release_time * (float)cx.sample_rate)) : 0.0f;
      if (s.lookahead_time != old_look) {  // LookupDelayLine::new (delay.rs:182-203)
        uint32_t df = (uint32_t)f32_ceil_u64(s.lookahead_time * (float)cx.sample_rate);
        uint32_t n = 1; while (n < df) n <<= 1;
        s.delay_frames = df; s.buf_frames = df ? n : 0; s.mask = df ? n - 1 : 0;
        s.write_pos = 0; s.peak_value = 0.0; s.peak_pos = 0;
        double* line = cx.aux_arena + s.aux;
        for (uint32_t i = 0; i < s.buf_frames * 2; ++i) line[i] = 0.0;
      }
      break;
    }
    case FX_CHORUS: {
      ChorusState& s = *(ChorusState*)st;
      if (id == CC4('r', 'a', 't', 'e')) lin_set_target(s.rate, v, cx.comp);
      else if (id == CC4('p', 'h', 'a', 's')) lin_set_target(s.phase, v, cx.comp);
      else if (id == CC4('d', 'p', 't', 'h')) exp_set_target(s.depth, v, cx.comp);
      else if (id == CC4('f', 'd', 'b', 'k')) exp_set_target(s.feedback, v, cx.comp);
      else if (id == CC4('d', 'l', 'a', 'y')) s.delay.target = v;
      else if (id == CC4('w', 'e', 't', '_')) exp_set_target(s.wet, v, cx.comp);
      else if (id == CC4('f', 'l', 't', 't')) { s.filter_type = (uint32_t)v; svf_set_type(s.coef, s.filter_type); }
      else if (id == CC4('f', 'l', 't', 'f')) exp_set_target(s.filter_freq, v, cx.comp);
      else if (id == CC4('f', 'l', 't', 'q')) exp_set_target(s.filter_res, v, cx.comp);
      break;
    }
    case FX_DELAY: {
      DelayState& s = *(DelayState*)st;
      if (id == CC4('m', 'o', 'd', 'e')) s.mode = (uint32_t)v;
      else if (id == CC4('d', 'l', 'a', 'y')) s.delay_time.target = v;
      else if (id == CC4('f', 'd', 'b', 'k')) exp_set_target(s.feedback, v, cx.comp);
      else if (id == CC4('f', 't', 'y', 'p')) s.filter_type = (uint32_t)v;
      else if (id == CC4('c', 'u', 't', 'o')) exp_set_target(s.cutoff, v, cx.comp);
      else if (id == CC4('d', 'r', 'i', 'v')) exp_set_target(s.drive, v, cx.comp);
      else if (id == CC4('w', 'e', 't', '_')) exp_set_target(s.wet, v, cx.comp);
      else if (id == CC4('w', 'd', 't', 'h')) exp_set_target(s.width, v, cx.comp);
      else if (id == CC4('l', 'f', 'o', 'r')) exp_set_target(s.lfo_rate, v, cx.comp);
      else if (id == CC4('l', 'f', 'o', 's')) { s.lfo_shape = (uint32_t)v; s.lfo.waveform = s.lfo_shape; }
      else if (id == CC4('l', 'f', 'd', 't')) exp_set_target(s.lfo_dt, v, cx.comp);
      else if (id == CC4('l', 'd', 'f', 'b')) exp_set_target(s.lfo_dfb, v, cx.comp);
      else if (id == CC4('l', 'f', 'd', 'f')) exp_set_target(s.lfo_dflt, v, cx.comp);
      break;
    }
    case FX_REVERB: {
      ReverbState& s = *(ReverbState*)st;
      if (id == CC4('r', 'o', 'o', 'm')) lin_set_target(s.room, v, cx.comp);
      else if (id == CC4('w', 'e', 't', ' ')) exp_set_target(s.wet, v, cx.comp);
      break;
    }
  }
#undef CC4
}

constexpr uint32_t FX_THREADS_FULL = 256;  // threads of the mixer CTA (== FX_THREADS in mixer_kernel.cuh)
// Effect::process dispatch, called by every thread of the mixer CTA
PB_DEV void fx_process(const FxHeader& h, const FxCtx& cx, const ChunkBuf& cb, uint32_t frames, uint32_t tid) {
  const uint32_t lane = tid & 31, warp = tid >> 5;
  uint8_t* st = cx.state_arena + h.state_offset;
  switch (h.kind) {
    case FX_FILTER: filter_process(*(FilterState*)st, cx, cb, frames, lane, warp); break;
    case FX_EQ5: eq5_process(*(Eq5State*)st, cx, cb, frames, lane, warp); break;
    case FX_GAIN: gain_process(*(GainState*)st, cx, cb, frames, tid, FX_THREADS_FULL); break;
    case FX_PANNING: pan_process(*(PanState*)st, cx, cb, frames, tid, FX_THREADS_FULL); break;
    case FX_DISTORTION: dist_process(*(DistState*)st, cx, cb, frames, tid, FX_THREADS_FULL); break;
    case FX_COMPRESSOR: if (tid == 0) comp_process(*(CompState*)st, cx, cb, frames); break;
    case FX_CHORUS: if (tid == 0) chorus_process(*(ChorusState*)st, cx, cb, frames); break;
    case FX_DELAY: if (tid == 0) delay_process(*(DelayState*)st, cx, cb, frames); break;
    case FX_REVERB: if (tid == 0) reverb_process(*(ReverbState*)st, cx, cb, frames); break;
  }
}

}  // namespace pb
