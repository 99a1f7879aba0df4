// Per-effect device state records (plain structs shared by the host graph compiler and the kernels).
#pragma once
#include "dev_structs.h"

namespace pb {

struct LfoSt { float phase, phase_inc; uint32_t waveform; };
struct IDelay { uint32_t mask, write_pos, aux; uint32_t _pad; };

// ==== per-kind state =================================================================================

struct FilterState {  // src/effect/filter.rs:48-57
  BiquadCoef coef;
  double ic1[2], ic2[2];
  ExpSm cutoff;
  LinSm q;
  uint32_t filter_type;  // already mapped to BqType
  uint32_t _pad;
};

struct Eq5State {  // src/effect/eq5.rs:19-33
  BiquadCoef coef[5];
  double ic1[2][5], ic2[2][5];
  ExpSm gains[5], freqs[5];
  LinSm bws[5];
};

struct CompState {  // src/effect/compressor.rs:24-38
  float threshold, ratio, knee, attack_time, release_time, lookahead_time;
  ExpSm makeup;
  float env_cur, atk_coeff, rel_coeff;
  uint32_t write_pos, mask, delay_frames, peak_pos, buf_frames;
  double peak_value;
  uint32_t aux;  // [buf_frames][2] doubles
  uint32_t aux_capacity_frames;
  uint32_t peak_dirty;  // the chunk-parallel path does not track the window peak; rescan before limiter use
  uint32_t _pad;
};

struct ChorusState {  // src/effect/chorus.rs:48-74
  LinSm rate, phase;
  ExpSm depth, feedback, wet, filter_freq, filter_res;
  SpringSm delay;
  uint32_t filter_type;
  float lfo_range;
  double current_phase;
  LfoSt left_osc, right_osc;
  IDelay dl, dr;
  SvfCoef coef;
  double fl_ic1, fl_ic2, fr_ic1, fr_ic2;
};

struct DelayState {  // src/effect/delay.rs:83-109
  SpringSm delay_time;
  ExpSm feedback, cutoff, drive, wet, width, lfo_rate, lfo_dt, lfo_dfb, lfo_dflt;
  uint32_t mode, filter_type, lfo_shape, _pad;
  IDelay dl, dr;
  LfoSt lfo;
  SvfCoef coef;
  double fl_ic1, fl_ic2, fr_ic1, fr_ic2;
  double dcl_x1, dcl_y1, dcr_x1, dcr_y1, dc_r;
  float fb_l, fb_r;
};

struct GainState {  // src/effect/gain.rs:51-60
  ExpSm gain;
  uint32_t dc_mode;          // GainEffectDcFilterMode: 0 Off 1 Slow 2 Default 3 Fast
  uint32_t _pad;
  double dc_r;               // DcFilter::r (shared by the channels)
  double dc_x1[2], dc_y1[2];
};

struct DistState {  // src/effect/distortion.rs:195-204
  LinSm drive;               // LinearSmoothedValue, step 0.01
  ExpSm mix;                 // ExponentialSmoothedValue, inertia 0.1
  uint32_t type;             // DistortionType: 0 SoftClip 1 HardClip 2 Diode 3 Fuzz 4 Fold
  float lut[5][256];         // compensation_luts, built on the host at construction like the reference
};

struct GateState {  // src/effect/gate.rs:12-28
  float threshold, attack_time, hold_time, release_time, range;
  float env_cur, env_atk, env_rel;   // EnvelopeFollower (envelope.rs:5-75)
  float gate_gain_db, attack_coeff, release_coeff;
  uint32_t hold_counter;
};

struct PanState {  // src/effect/pan.rs:17-25
  ExpSm pan, width;
  uint32_t invert_l, invert_r;
};

struct RvLine { uint32_t aux, size, count, delay; double depth; double feedback[2]; double vib_phase[2]; };
struct RvAllpass { uint32_t aux, size, delay, write_pos; };
struct ReverbState {  // src/effect/reverb.rs:38-72
  LinSm room;
  ExpSm wet;
  BiquadCoef ca, cb, cc;
  double a_ic[2][2], b_ic[2][2], c_ic[2][2];  // [channel][ic1, ic2]
  uint32_t fpd_l, fpd_r;
  RvLine lines[8];
  RvAllpass ap[4];
  uint32_t m_aux, m_mask, m_write_pos, _pad;
};

}  // namespace pb
