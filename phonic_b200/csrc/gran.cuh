// Granular playback of Sampler voices (src/generator/sampler/granular.rs; SamplerVoice::process granular arm,
// src/generator/sampler/voice.rs:412-427) in three passes:
//   skeleton : gran_advance() runs the per-frame control of GrainPool::process -- trigger oscillator (f32
//              accumulate), playhead (f32 accumulate), pool-slot allocation, Sequential-mode crossfade gate (exact
//              f64 window-phase chain of the primary grain) -- and emits one GrainRec per grain and time block.
//   grains   : grain_kernel, one thread per GrainRec, runs the grain's own f64 position / window-phase
//              recurrences (Grain::process, granular.rs:1081-1120), window LUT lerp, Catmull-Rom read of the mono
//              sample buffer at the f32 position (granular.rs:901-933) and stores the stereo contribution of every
//              sample of the grain in the block's grain storage.
//   replay   : gran_replay_frames() adds the contributions of a voice's live grains in activation order -- the
//              order of GrainPool::active_grain_indices, so the f32 sum is the reference's -- and applies the AHDSR.
// Deterministic subset only (no OS-seeded randomisation, no modulation routings): DESIGN.md §2.
#pragma once
#include "voice.cuh"

namespace pb {

struct GranEmit {
  GrainRec* recs;        // this block's record list (all granular voices)
  uint32_t* counters;    // [0] records, [1] storage frames
  uint32_t rec_cap;
  uint32_t* vrec;        // [n_rows][vrec_cap]: this block's records of each voice, in activation order
  uint32_t vrec_cap;
  uint32_t* tile_range;  // [n_rows][n_tiles][2]: candidate records (first, last) of every 64-frame tile
  uint32_t n_tiles;
  uint32_t gen;          // generation tag of this block
  uint32_t block_frames;
};

// f64::rem_euclid. fmod(a, b) == a exactly for 0 <= a < b: the common case skips the (slow) fmod
PB_DEV double gran_rem_euclid(double a, double b) {
  if (a >= 0.0 && a < b) return a;
  double r = fmod(a, b);
  return r < 0.0 ? r + fabs(b) : r;
}
PB_DEV float gran_rem_euclidf(float a, float b) {
  if (a >= 0.0f && a < b) return a;
  float r = fmodf(a, b);
  return r < 0.0f ? r + fabsf(b) : r;
}
PB_DEV double gran_fold(double position, double ls, double le) {  // GrainPool::fold_into_loop_range
  const double loop_len = le - ls;
  return loop_len > 0.0 ? ls + gran_rem_euclid(position - ls, loop_len) : ls;
}

// GrainPool::start (granular.rs:477-491)
PB_DEV void gran_start(GranState& s, const GranGroup& gg, double speed, float volume, float panning) {
  s.trigger_new = 1;
  s.trigger_phase = 1.0f;
  s.speed = speed; s.volume = volume; s.panning = panning;
  s.playhead = gg.position;
  s.playing_loop = 0;
}

// Emits the record of the grain in `slot` for this block. `cur_off` = block-relative frame of its next sample.
PB_DEV void gran_emit(GranState& s, const GranEmit& em, const GranGroup& gg, uint32_t row, uint32_t slot, uint32_t cur_off,
                      uint32_t done, const GrainRec& proto) {
  const uint32_t total = s.slot_total[slot];
  const uint32_t len = min(total - done, em.block_frames - cur_off);
  const uint32_t i = atomicAdd(&em.counters[0], 1u);
  const uint32_t st = atomicAdd(&em.counters[1], len);
  s.slot_rec[slot] = i;
  if (i < em.rec_cap) {
    GrainRec r = proto;
    r.row = row; r.start_off = cur_off; r.len = len; r.done = done; r.total = total; r.slot = slot; r.storage = st;
    r.buffer = gg.buffer; r.buf_len = gg.buf_len;
    em.recs[i] = r;
  }
  if (s.n_recs < em.vrec_cap) em.vrec[(size_t)row * em.vrec_cap + s.n_recs] = i;
  s.n_recs++;
}

// First use of a voice in a time block: grains that are still playing get a continuation record each, in
// activation order; the per-block bookkeeping restarts.
PB_DEV void gran_block_begin(GranState& s, const GranEmit& em, const GranGroup& gg, uint32_t row, uint64_t now, uint32_t cur_off) {
  s.gen = em.gen;
  s.n_recs = 0;
  s.first_active = 0;
  uint32_t kept = 0;
  for (uint32_t i = 0; i < s.n_order; ++i) {
    const uint32_t slot = s.order[i];
    if (s.slot_end[slot] > now) {
      s.order[kept++] = (uint8_t)slot;
      GrainRec proto;
      proto.position = 0.0; proto.increment = 0.0; proto.win_inc = 0.0; proto.loop_start = 0.0; proto.loop_end = 0.0;
      proto.volume = 0.0f; proto.panning = 0.0f; proto.window_mode = 0; proto.has_loop = 0; proto._pad = 0;
      const uint32_t done = s.slot_total[slot] - (uint32_t)(s.slot_end[slot] - now);
      gran_emit(s, em, gg, row, slot, cur_off, done, proto);
    }
  }
  s.n_order = kept;
}

// GrainPool::reset (granular.rs:497-504) when the voice resets at block-relative frame `cur_off` / absolute `now`:
// every live grain stops contributing from here on.
PB_DEV void gran_reset(GranState& s, const GranEmit& em, uint64_t now, uint32_t cur_off) {
  if (s.gen == em.gen) {
    for (uint32_t i = 0; i < s.n_order; ++i) {
      const uint32_t slot = s.order[i];
      if (s.slot_end[slot] > now && s.slot_rec[slot] < em.rec_cap) {
        GrainRec& r = em.recs[s.slot_rec[slot]];
        r.len = min(r.len, cur_off - r.start_off);
        r.total = r.done + r.len;  // cut short: it must not leave a carry behind (its slot may be re-used in this block)
      }
    }
  }
  for (uint32_t i = 0; i < GRAIN_POOL; ++i) s.slot_end[i] = 0;
  s.n_order = 0;
  s.max_end = 0;
  s.trigger_new = 1;
  s.has_primary = 0;
}

// GrainPool::try_trigger_grain -> activate_new_grain for the frame at absolute time `t` (granular.rs:524-603,
// 813-897) with all modulation inputs and random variations zero.
PB_DEV void gran_activate(GranState& s, const GranEmit& em, const GranGroup& gg, uint32_t row, uint64_t t, uint32_t cur_off) {
  uint32_t slot = GRAIN_POOL;
  for (uint32_t i = 0; i < GRAIN_POOL; ++i) if (s.slot_end[i] <= t) { slot = i; break; }
  if (slot == GRAIN_POOL) return;  // pool exhausted: the trigger is lost
  // GrainPool::playback_position (granular.rs:446-475) + loop fold + rem_euclid (granular.rs:575-585)
  float base = gg.step == 0.0f ? gg.position : s.playhead;
  const bool in_loop = s.playing_loop && gg.has_loop;
  if (in_loop) base = (float)gran_fold((double)base, (double)gg.loop_start, (double)gg.loop_end);
  base = gran_rem_euclidf(base, 1.0f);
  double grain_position = (double)base + 0.0;
  if (in_loop) grain_position = gran_fold(grain_position, (double)gg.loop_start, (double)gg.loop_end);
  grain_position = gran_rem_euclid(grain_position, 1.0);
  GrainRec proto;
  proto.window_mode = (uint16_t)gg.window;
  proto.position = fmin(fmax(grain_position, 0.0), 1.0);
  proto.volume = fminf(fmaxf(s.volume * 1.0f, 0.0f), 100.0f);
  proto.panning = fminf(fmaxf(fminf(fmaxf(s.panning + 0.0f, -1.0f), 1.0f), -1.0f), 1.0f);
  const double base_increment = gg.buf_len > 0 ? (s.speed * 1.0) / (double)gg.buf_len : 0.0;
  proto.increment = base_increment * (gg.backward ? -1.0 : 1.0);
  proto.win_inc = gg.grain_size > 0 ? 1.0 / (double)gg.grain_size : 0.0;
  proto.has_loop = in_loop ? 1 : 0; proto._pad = 0;
  proto.loop_start = (double)gg.loop_start; proto.loop_end = (double)gg.loop_end;
  // active_grain_indices: drop a stale entry of this slot, push (granular.rs:888-893)
  uint32_t kept = 0;
  for (uint32_t i = 0; i < s.n_order; ++i) if (s.order[i] != slot) s.order[kept++] = s.order[i];
  s.order[kept++] = (uint8_t)slot;
  s.n_order = kept;
  s.slot_total[slot] = gg.grain_size;
  s.slot_end[slot] = t + gg.grain_size;
  s.max_end = max(s.max_end, s.slot_end[slot]);
  gran_emit(s, em, gg, row, slot, cur_off, 0u, proto);
  if (s.overlap_mode == 1) {
    s.has_primary = 1; s.primary_slot = slot; s.primary_phase = 0.0; s.primary_inc = proto.win_inc;
    s.primary_end = s.slot_end[slot];
  }
}

// State-only advance of `n` frames of GrainPool::process starting at absolute frame `t0` (the skeleton pass).
__device__ __noinline__ void gran_advance(GranState* __restrict__ sp, const GranGroup* __restrict__ ggp, const GranEmit em,
                                          uint32_t row, uint64_t t0, uint32_t cur_off, uint32_t n) {
  GranState& s = *sp;
  const GranGroup gg = *ggp;
  if (s.gen != em.gen) gran_block_begin(s, em, gg, row, t0, cur_off);
  const uint32_t tile = cur_off / 64u;
  uint32_t* tr = em.tile_range + ((size_t)row * em.n_tiles + tile) * 2;
  {  // oldest record that can still sound in this tile
    uint32_t fa = s.first_active;
    while (fa < s.n_recs && fa < em.vrec_cap) {
      const uint32_t ri = em.vrec[(size_t)row * em.vrec_cap + fa];
      if (ri >= em.rec_cap) break;
      const GrainRec& r = em.recs[ri];
      if (r.start_off + r.len > cur_off) break;
      ++fa;
    }
    s.first_active = fa;
    if (tr[1] == 0xFFFFFFFFu) tr[0] = fa;  // the tile's first segment sets the lower bound (the host clears the table to ~0)
  }
  float trigger_phase = s.trigger_phase, playhead = s.playhead;
  const bool move_playhead = gg.step != 0.0f && gg.buf_len > 0;
  const float position_increment = (gg.step * (1.0f + 0.0f)) / (float)gg.buf_len;
  if (s.overlap_mode != gg.overlap_mode) { s.overlap_mode = (uint8_t)gg.overlap_mode; s.has_primary = 0; }  // granular.rs:536-539
  const bool sequential = gg.overlap_mode == 1;
  const bool trigger_new = s.trigger_new != 0;
  bool playing_loop = s.playing_loop != 0, has_primary = s.has_primary != 0;
  uint64_t primary_end = s.primary_end;
  double primary_phase = s.primary_phase, primary_inc = s.primary_inc;
  uint32_t i = 0;
  while (i < n) {
    if (!sequential) {
      // Cloud mode fast span: frames in which no grain triggers and the playhead update is a plain `+= increment`
      // (no wrap, no loop entry, and inside the loop range the fold is the identity: (double)x - ls and ls + that
      // are exact for f32 operands) are one FADD each; anything else falls through to the reference's own code.
      float lo = 0.0f, hi = 1.0f;
      if (move_playhead && gg.has_loop) {
        if (playing_loop) { lo = gg.loop_start; hi = gg.loop_end; }
        else if (playhead < gg.loop_start) { hi = fminf(gg.loop_start, 1.0f); }
        else if (playhead >= gg.loop_end) { lo = fmaxf(gg.loop_end, 0.0f); }
        else { lo = 1.0f; hi = 0.0f; }  // about to enter the loop: slow path
      }
      const float tinc = trigger_new ? gg.trigger_inc : 0.0f;
      const float pinc = move_playhead ? position_increment : 0.0f;
      if (!move_playhead) { lo = -3.0e38f; hi = 3.0e38f; }
      for (; i < n; ++i) {
        const float tpn = trigger_phase + tinc;
        const float phn = playhead + pinc;
        if ((tpn >= 1.0f) | !(phn >= lo) | !(phn < hi)) break;
        trigger_phase = tpn;
        playhead = phn;
      }
      if (i >= n) break;
    }
    const uint64_t t = t0 + i;
    // try_trigger_grain (granular.rs:524-560)
    bool trig = false;
    if (sequential) {
      const bool blocked = has_primary && t < primary_end && primary_phase < (double)gg.crossfade;
      trig = !blocked && trigger_new;
    } else if (trigger_new) {
      trigger_phase += gg.trigger_inc;
      if (trigger_phase >= 1.0f) { trigger_phase -= 1.0f; trig = true; }
    }
    if (trig) {
      s.playhead = playhead; s.playing_loop = playing_loop ? 1 : 0;
      gran_activate(s, em, gg, row, t, cur_off + i);
      has_primary = s.has_primary != 0; primary_end = s.primary_end; primary_phase = s.primary_phase; primary_inc = s.primary_inc;
    }
    // advance_playhead (granular.rs:607-640)
    if (move_playhead) {
      playhead += position_increment;
      if (gg.has_loop) {
        if (playing_loop) playhead = (float)gran_fold((double)playhead, (double)gg.loop_start, (double)gg.loop_end);
        else if (playhead >= gg.loop_start && playhead < gg.loop_end) playing_loop = true;
        else { if (playhead >= 1.0f) playhead -= 1.0f; else if (playhead < 0.0f) playhead += 1.0f; }
      } else if (playhead >= 1.0f) playhead -= 1.0f;
      else if (playhead < 0.0f) playhead += 1.0f;
    }
    // the primary grain's Grain::process of this frame (granular.rs:1094)
    if (has_primary && t < primary_end) primary_phase += primary_inc;
    ++i;
  }
  s.trigger_phase = trigger_phase; s.playhead = playhead; s.playing_loop = playing_loop ? 1 : 0;
  s.primary_phase = primary_phase;
  tr[1] = s.n_recs;
}

// ---- grain kernel ----------------------------------------------------------------------------------------------
struct GrainArgs {
  const GrainRec* recs;
  const uint32_t* counters;
  uint32_t rec_cap;
  const DevBuffer* buffers;
  const float* window_luts;      // [8][2048] (GrainWindow::new, granular.rs:110-196; built on the host)
  float2* storage;               // this block's grain storage
  uint32_t storage_cap;
  const GrainCarry* carry_in;    // [n_rows][GRAIN_POOL] written by the previous block's launch
  GrainCarry* carry_out;
};

constexpr float GRAIN_ENVELOPE_THRESHOLD = 0.001f;
// marks a sample the reference skips (`envelope <= threshold`): nothing is added to the frame
PB_DEV float2 grain_skip() { return make_float2(__int_as_float(0x7fc00000), 0.0f); }

__global__ void __launch_bounds__(128) grain_kernel(GrainArgs a) {
  const uint32_t n = min(a.counters[0], a.rec_cap);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const GrainRec r = a.recs[i];
  if (r.len == 0 || (size_t)r.storage + r.len > a.storage_cap) return;
  // a continuation record carries only (row, slot): the grain's state is what the last block's launch left behind
  GrainCarry st;
  if (r.done > 0) {
    st = a.carry_in[(size_t)r.row * GRAIN_POOL + r.slot];
  } else {
    st.position = r.position; st.window_phase = 0.0; st.increment = r.increment; st.win_inc = r.win_inc;
    st.loop_start = r.loop_start; st.loop_end = r.loop_end; st.volume = r.volume; st.panning = r.panning;
    st.window_mode = r.window_mode; st.has_loop = r.has_loop;
  }
  double position = st.position, phase = st.window_phase;
  const double increment = st.increment, win_inc = st.win_inc;
  const float volume = st.volume, panning = st.panning;
  const float* __restrict__ buf = a.buffers[r.buffer].data;
  const uint32_t max_index = r.buf_len - 1;
  const float* __restrict__ lut = a.window_luts + (size_t)st.window_mode * GRAIN_LUT_N;
  const float left_gain = (1.0f - panning) * 0.5f, right_gain = (1.0f + panning) * 0.5f;
  const double loop_len = st.loop_end - st.loop_start;
  float2* out = a.storage + r.storage;
  for (uint32_t k = 0; k < r.len; ++k) {
    // GrainWindow::sample (granular.rs:201-215)
    const double index_float = phase * (double)(GRAIN_LUT_N - 1);
    const uint32_t index = (uint32_t)index_float & (GRAIN_LUT_N - 1);
    const float fraction = (float)(index_float - trunc(index_float));
    float envelope_value;
    if (index < GRAIN_LUT_N - 1) envelope_value = __ldg(lut + index) * (1.0f - fraction) + __ldg(lut + ((index + 1) & (GRAIN_LUT_N - 1))) * fraction;
    else envelope_value = __ldg(lut + GRAIN_LUT_N - 1);
    const float pos = (float)position;
    // Grain::process (granular.rs:1091-1108)
    position += increment;
    phase += win_inc;
    if (st.has_loop) {
      if (loop_len > 0.0) position = st.loop_start + gran_rem_euclid(position - st.loop_start, loop_len);
    } else if (position < 0.0) position += 1.0;
    else if (position > 1.0) position -= 1.0;
    const float envelope = envelope_value * volume;
    float2 o = grain_skip();
    if (envelope > GRAIN_ENVELOPE_THRESHOLD) {
      // GrainPool::sample_at_position (granular.rs:901-933)
      const float float_index = pos * (float)max_index;
      uint32_t idx = float_index >= 4294967295.0f ? 0xFFFFFFFFu : (float_index > 0.0f ? (uint32_t)float_index : 0u);
      idx = min(idx, max_index);
      const float fr = float_index - (float)idx;
      const uint32_t i1 = idx;
      const uint32_t i2 = i1 < max_index ? i1 + 1 : 0;
      const uint32_t i0 = i1 > 0 ? i1 - 1 : max_index;
      const uint32_t i3 = i2 < max_index ? i2 + 1 : 0;
      const float y0 = __ldg(buf + i0), y1 = __ldg(buf + i1), y2 = __ldg(buf + i2), y3 = __ldg(buf + i3);
      const float ca = -0.5f * y0 + 1.5f * y1 - 1.5f * y2 + 0.5f * y3;
      const float cb = y0 - 2.5f * y1 + 2.0f * y2 - 0.5f * y3;
      const float cc = -0.5f * y0 + 0.5f * y2;
      const float cd = y1;
      const float sample = ca * fr * fr * fr + cb * fr * fr + cc * fr + cd;
      const float windowed = sample * envelope;
      o = make_float2(windowed * left_gain, windowed * right_gain);
    }
    out[k] = o;
  }
  if (r.done + r.len < r.total) {
    st.position = position; st.window_phase = phase;
    a.carry_out[(size_t)r.row * GRAIN_POOL + r.slot] = st;
  }
}

// ---- replay ------------------------------------------------------------------------------------------------------
struct GranReplay {
  const GrainRec* recs;
  const uint32_t* vrec;
  const uint32_t* tile_range;
  const float2* storage;
  uint32_t rec_cap, vrec_cap, n_tiles, storage_cap;
};

// Up to GRAN_CHUNK frames of a granular voice starting at block-relative frame c.hq_off: the ordered sum of the live
// grains' contributions (grains outer, frames inner -- per frame the adds keep the activation order), then the AHDSR
// exactly as voice_frames() applies it (voice.rs:470-486); the result is stored to / added to `out`. The frame loops are
// fully unrolled over registers: the contribution loads of one grain are independent and in flight together, and the
// next grain's record is fetched while this one is summed (the storage / record reads come from L2 or HBM: their latency,
// not the adds, bounds this function).
constexpr uint32_t GRAN_CHUNK = 16;
struct GranPiece { uint32_t s0, len, storage; };
PB_DEV GranPiece gran_piece(const GranReplay& g, uint32_t row, uint32_t i, uint32_t last) {
  GranPiece p{0u, 0u, 0u};
  if (i < last) {
    const uint32_t ri = g.vrec[(size_t)row * g.vrec_cap + i];
    if (ri < g.rec_cap) {
      const GrainRec& r = g.recs[ri];
      p.s0 = r.start_off; p.len = r.len; p.storage = r.storage;
      if ((size_t)p.storage + p.len > g.storage_cap) p.len = 0;
    }
  }
  return p;
}
PB_DEV uint32_t gran_replay_frames(VoiceState& v, CallCtx& c, const GroupParams& gp, const GranReplay& g, uint32_t row,
                                   uint32_t n, float* __restrict__ out, const bool acc) {
  const uint32_t lo = c.hq_off, hi = lo + n;
  float sl[GRAN_CHUNK], sr[GRAN_CHUNK];
#pragma unroll
  for (uint32_t j = 0; j < GRAN_CHUNK; ++j) { sl[j] = 0.0f; sr[j] = 0.0f; }
  const uint32_t tile = lo / 64u;
  const uint32_t* tr = g.tile_range + ((size_t)row * g.n_tiles + tile) * 2;
  const uint32_t first = tr[0], last = min(tr[1], g.vrec_cap);
  GranPiece nxt = gran_piece(g, row, first, last);
  for (uint32_t i = first; i < last; ++i) {
    const GranPiece p = nxt;
    nxt = gran_piece(g, row, i + 1, last);
    const uint32_t a0 = max(lo, p.s0), a1 = min(hi, p.s0 + p.len);
    if (a0 >= a1) continue;
    const float2* __restrict__ src = g.storage + p.storage;
    float2 x[GRAN_CHUNK];
#pragma unroll
    for (uint32_t j = 0; j < GRAN_CHUNK; ++j) {
      const uint32_t f = lo + j;
      x[j] = (f >= a0 && f < a1) ? __ldg(src + (f - p.s0)) : make_float2(__int_as_float(0x7fc00000), 0.0f);
    }
#pragma unroll
    for (uint32_t j = 0; j < GRAN_CHUNK; ++j)
      if (x[j].x == x[j].x) { sl[j] += x[j].x; sr[j] += x[j].y; }  // NaN marks samples the reference skips (and the gaps here)
  }
#pragma unroll
  for (uint32_t j = 0; j < GRAN_CHUNK; ++j) {
    if (j < n) {
      if (gp.has_env) {
        const float e = c.env_per_frame ? env_run(v, gp) : c.env_const;
        sl[j] *= e; sr[j] *= e;
      }
      out[2 * j] = acc ? out[2 * j] + sl[j] : sl[j];
      out[2 * j + 1] = acc ? out[2 * j + 1] + sr[j] : sr[j];
    }
  }
  c.hq_off += n;
  c.chunk_left -= n;
  return n;
}

}  // namespace pb
