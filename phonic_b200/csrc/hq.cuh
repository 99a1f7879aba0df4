// HighQuality (ResamplingQuality::HighQuality) file voices: phonic's RubatoResampler wrapper around
// rubato::SincFixedIn<f32> (src/utils/resampler/rubato.rs:12-154), driven by PreloadedFileSource::write_buffer
// (src/source/file/preloaded.rs:270-332).
//
// Three passes, like the cubic path, but the resampling arithmetic gets a kernel of its own:
//   skeleton  : hq_advance() walks write_buffer's process() iterations -- pending flushes, 256-frame input
//               chunks, zero padding at loop ends / EOF (and its "padded input counts as consumed" quirk),
//               the exact f64 `idx += t_ratio` recurrence that decides how many frames a chunk yields -- and
//               emits one HqRec per piece of resampler output the block consumes. No sample is touched.
//   sinc      : sinc_kernel (sinc_kernel.cuh) materialises those records into the per-block stream scratch
//               (interleaved stereo, one row per HighQuality voice).
//   replay    : hq_replay_frames() reads the stream at the voice's block offset and applies fader / gain / pan.
//
// rubato is a third-party crate that is not vendored with the reference: PARITY UNPINNED.
#pragma once
#include "voice.cuh"

namespace pb {

struct HqEmit {
  HqRec* recs;          // this block's record list (shared by all HighQuality voices)
  uint32_t* n_recs;     // its fill counter
  uint32_t cap;
  uint32_t buffer;      // DevBuffer index of the voice's sample data
};

// RubatoResampler::reset (rubato.rs:150-153): only the pending output is dropped; SincFixedIn keeps its
// history and last_index. `cur_off` = block-relative frame of the voice's next output.
PB_DEV void hq_reset_pending(HqState& h, const HqEmit& em, uint32_t cur_off) {
  if (h.rec != HQ_NONE && h.rec < em.cap) em.recs[h.rec].count = cur_off - em.recs[h.rec].out_off;
  h.pending = 0;
  h.rec = HQ_NONE;
}

// State-only advance of one write call of a HighQuality voice by up to `n` frames (the skeleton pass): walks
// write_buffer's loop iterations (preloaded.rs:283-329); frames that only drain pending output are taken in one step.
template <int CC>
__device__ __noinline__ uint32_t hq_advance(VoiceState& v, CallCtx& c, HqState* __restrict__ hp, const HqEmit em,
                                            const DevBuffer& b, float comp, uint32_t n) {
  HqState h = *hp;  // the resampler state lives in HBM between segments: the skeleton's hot cubic path keeps its registers
  uint32_t done = 0;
  uint32_t call_left = c.call_left, chunk_left = c.chunk_left, hq_off = c.hq_off;
  bool new_call = c.new_call;
  while (done < n && !c.ended) {
    if (call_left == 0) {  // write() without pitch slide: one write_buffer call (preloaded.rs:441-447)
      call_left = chunk_left;
      loop_range_samples(v, b, c.ls, c.le);
      new_call = true;
    }
    if (new_call) {  // first iteration of this write_buffer call
      new_call = false;
      if (h.pending > 0) {
        // process() only flushes pending output (rubato.rs:81-86): nothing is consumed -- except on the zero-pad
        // path, which counts the whole remaining input as consumed whatever process() did (preloaded.rs:296-304)
        const uint32_t remaining = c.le > v.playback_pos ? c.le - v.playback_pos : 0u;
        if (remaining < HQ_CHUNK * CC) v.playback_pos += remaining;
        after_process_call(v, c);
      }
    }
    if (h.pending > 0) {
      if (h.rec == HQ_NONE) {  // first frame this block takes from the open chunk: have it materialised
        const uint32_t i = atomicAdd(em.n_recs, 1u);
        h.rec = i;
        if (i < em.cap) {
          HqRec r;
          r.idx0 = h.idx0; r.t_ratio = h.t_ratio;
#pragma unroll
          for (int k = 0; k < 3; ++k) { r.src[k] = h.src[k]; r.valid[k] = h.valid[k]; }
          r.kind = v.hq == 2 ? 1 : 0;
          r.slot = h.slot; r.out_off = hq_off; r.skip = h.n_out - h.pending; r.count = h.pending;
          r.buffer = em.buffer; r.table = h.table;
          em.recs[i] = r;
        }
      }
      const uint32_t take = min(h.pending, n - done);  // (call_left >= n - done: a segment never outlives its call)
      h.pending -= take;
      call_left -= take; chunk_left -= take; hq_off += take; done += take;
      continue;
    }
    // pending is empty: the next iteration feeds the resampler (rubato.rs:93-133)
    const uint32_t remaining = c.le > v.playback_pos ? c.le - v.playback_pos : 0u;
    const bool pad = remaining < HQ_CHUNK * CC;
    uint32_t produced, consumed;
    h.src[0] = h.src[1]; h.src[1] = h.src[2]; h.valid[0] = h.valid[1]; h.valid[1] = h.valid[2];
    h.src[2] = v.playback_pos;
    h.rec = HQ_NONE;
    if (v.hq == 2) {
      // equal rates: process() copies min(input, output) samples (rubato.rs:73-78); the padded input is 256 frames
      const uint32_t room = call_left;
      if (pad) { produced = min(HQ_CHUNK, room); consumed = remaining; h.valid[2] = (uint16_t)min(remaining / CC, produced); }
      else { produced = min(remaining / CC, room); consumed = produced * CC; h.valid[2] = (uint16_t)produced; }
      h.idx0 = 0.0;
    } else {
      h.valid[2] = (uint16_t)min(remaining / CC, HQ_CHUNK);
      consumed = pad ? remaining : HQ_CHUNK * CC;
      // SincFixedIn::process_into_buffer: `while idx < end_idx { idx += t_ratio; n += 1 }` (fixed ratio), exactly
      double idx = h.last_index;
      const double t = h.t_ratio, end = (double)h.end_idx;
      uint32_t cnt = 0;
      for (;;) {  // eight adds per exit test: the chain of DADDs is the work, the compare-and-branch was most of the time
        const double a1 = idx + t, a2 = a1 + t, a3 = a2 + t, a4 = a3 + t, a5 = a4 + t, a6 = a5 + t, a7 = a6 + t;
        if (!(a7 < end)) break;
        idx = a7 + t; cnt += 8;
      }
      while (idx < end) { idx += t; ++cnt; }
      h.idx0 = h.last_index;
      h.last_index = idx - (double)HQ_CHUNK;
      produced = cnt;
    }
    h.n_out = produced;
    h.pending = produced;
    v.playback_pos += consumed;
    after_process_call(v, c);
    if (produced == 0 && v.pos_eof) { c.ended = true; break; }  // `playback_pos_eof && output_written == 0` (preloaded.rs:326-329)
  }
  c.call_left = call_left; c.chunk_left = chunk_left; c.hq_off = hq_off; c.new_call = new_call;
  *hp = h;
  advance_ramps(v, c, done, comp);
  return done;
}

// Replay of `n` frames of a HighQuality voice: the resampler output comes from the stream scratch row
// (interleaved stereo; mono sources are already duplicated), then VolumeFader / AmplifiedSource / PannedSource
// exactly as voice_frames() applies them. A HighQuality source never runs dry inside a call (EOF keeps feeding
// zero chunks until the call is full, preloaded.rs:283-329), so all `n` frames are produced.
PB_DEV uint32_t hq_replay_frames(VoiceState& v, CallCtx& c, const float* __restrict__ row, float comp, uint32_t n,
                                 float* __restrict__ out, const bool acc) {
  for (uint32_t f = 0; f < n; ++f) {
    const float2 x = *reinterpret_cast<const float2*>(row + (size_t)c.hq_off * 2);
    c.hq_off++;
    float x0 = x.x, x1 = x.y;
    if (c.fader_running) {
      v.fader_cur += (v.fader_tgt - v.fader_cur) * v.fader_inertia;
      x0 *= v.fader_cur; x1 *= v.fader_cur;
    } else if (c.fader_scale) {
      x0 *= v.fader_tgt; x1 *= v.fader_tgt;
    }
    float l = x0, r = x1;
    if (c.vol_ramp) { l *= exp_next(v.vol, comp); r *= exp_next(v.vol, comp); }
    else if (c.vol_scale) { l *= v.vol.target; r *= v.vol.target; }
    if (c.pan_ramp) {
      float pl, pr;
      panning_factors(exp_next(v.pan, comp), pl, pr);
      l *= pl; r *= pr;
    } else if (c.pan_apply) {
      l *= c.pan_l; r *= c.pan_r;
    }
    out[2 * f] = acc ? out[2 * f] + l : l;
    out[2 * f + 1] = acc ? out[2 * f + 1] + r : r;
  }
  return n;
}

}  // namespace pb
