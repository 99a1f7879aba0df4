// Skeleton kernel (pass 1 of the voice path): one CTA per group (a Sampler with N voices or one file
// playback), one thread per voice. It walks the mixer's exact chunk schedule for one time block
// (MixedSource::write chunking, src/source/mixed.rs:679-693), resolves note/speed/seek/stop events on
// device at the frame they are due (voice allocation and stealing included), and advances every voice's
// *control* state -- positions, the f32 phase recurrence, ramps, the envelope stage machine -- without
// touching a single audio sample. At every 64-frame tile boundary and every write-call boundary it
// emits a Segment snapshot; pass 2 (replay_kernel.cuh) renders all segments of the block in parallel.
//
// This is the serial part the reference's f32 recurrences force; everything per-sample
// that can be replayed from a checkpoint is left to pass 2.
#pragma once
#include "hq.cuh"
#include "gran.cuh"

namespace pb {

constexpr uint32_t TILE = 64;  // frames per replay tile (== FileSourceImpl::SPEED_UPDATE_CHUNK_SIZE)

struct SkeletonArgs {
  const GroupParams* groups;
  const uint32_t* group_list;         // groups of this launch (one size class)
  GroupState* gstate;
  VoiceState* voices;
  const DevBuffer* buffers;
  const DevEvent* events;
  const uint64_t* chunk_bounds;       // concatenated chunk boundaries of this time block
  const uint32_t* mixer_chunk_begin;  // [n_mixers + 1]
  uint8_t* group_flags;               // [n_groups][max_chunks]: source produced output in chunk k
  uint32_t max_chunks;
  uint32_t block_frames;
  uint64_t block_start;
  RenderConsts rc;
  // segment output
  Segment* segs;                      // [n_voices][seg_cap]
  uint16_t* seg_first;                // [n_voices][n_tiles]
  uint16_t* seg_count;                // [n_voices][n_tiles]
  GroupSeg* gsegs;                    // [n_groups][seg_cap]
  uint16_t* gseg_first;               // [n_groups][n_tiles]
  uint16_t* gseg_count;               // [n_groups][n_tiles]
  TileRec* recs;                      // [n_voices][n_tiles]: continuation records of simple calls
  uint32_t gen;                       // generation tag of this launch's records
  uint32_t seg_cap;
  uint32_t n_tiles;
  // HighQuality (rubato sinc) file voices: resampler state per voice + this block's record list (hq.cuh)
  HqState* hq_states;                 // [n_voices] or nullptr when the graph has no HighQuality source
  HqRec* hq_recs;
  uint32_t* hq_n_recs;
  uint32_t hq_cap;
  // granular samplers (gran.cuh): per-group parameters, per-voice GrainPool control state, this block's emit context
  const GranGroup* gran_groups;       // [n_groups] or nullptr when the graph has no granular sampler
  GranState* gran_states;             // [n_gran_rows]
  GranEmit gran;
  unsigned long long* prof;           // PB200_SKEL_PROF: [n_voices][4] cycles waiting at the free run's three barriers + its own work
  uint32_t debug_flags;  // timing experiments only (PB200_SKEL_DEBUG): 1 = no snapshot stores, 2 = no simple calls
};

constexpr int VK_MAX_VOICES = 1024;
#ifdef PB200_CYC
// build-time cycle counters of ONE voice (debug builds only, -DPB200_CYC=<global voice index>)
__device__ unsigned long long g_cyc[8];
#define CYC_ON(gv) ((gv) == PB200_CYC)
#define CYC_T() clock64()
#else
#define CYC_ON(gv) false
#define CYC_T() 0ll
#endif
constexpr uint32_t SB_MAX = 256;   // chunk boundaries cached in shared memory
constexpr uint32_t MAX_RUN = 64;   // chunks per free run

struct VoiceHeader {  // what Sampler::next_free_voice_index needs (sampler.rs:826-860)
  uint64_t note_id;
  uint64_t release_start;
  uint8_t active, in_release, has_release;
};

// Sampler::next_free_voice_index; evaluated redundantly by every thread on the shared headers
PB_DEV uint32_t next_free_voice_index(const VoiceHeader* h, uint32_t n, bool has_env) {
  for (uint32_t i = 0; i < n; ++i)
    if (!h[i].active) return i;
  uint32_t candidate = 0;
  bool has_earliest = false, has_oldest = false;
  uint64_t earliest = 0, oldest = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (has_env && h[i].in_release) {
      if (h[i].has_release) {
        if (!has_earliest || h[i].release_start < earliest) {
          has_earliest = true; earliest = h[i].release_start; has_oldest = false; candidate = i;
        }
      }
    } else if (!has_earliest) {
      if (h[i].active) {
        if (!has_oldest || h[i].note_id < oldest) { has_oldest = true; oldest = h[i].note_id; candidate = i; }
      }
    }
  }
  return candidate;
}

PB_DEV void publish_header(VoiceHeader* h, uint32_t i, const VoiceState& v) {
  h[i].note_id = v.note_id;
  h[i].release_start = v.release_start;
  h[i].active = v.has_note;
  h[i].in_release = v.env_stage == ENV_RELEASE;
  h[i].has_release = v.has_release;
}

// SamplerVoice::stop (voice.rs:196-219)
PB_DEV void sampler_voice_stop(VoiceState& v, const GroupParams& gp, uint64_t frame) {
  if (v.has_note) {
    v.has_release = 1;
    v.release_start = frame;
    if (gp.has_env) env_note_off(v, gp);
    else file_stop(v, gp);
  }
}


// One simple call (voice.cuh "simple calls"): the whole call's phase / envelope recurrences in one go, with a
// 32-byte TileRec stored at every tile boundary crossed instead of a full Segment.
// UNI: lane-per-voice skeleton -- one instruction stream for every ratio / envelope state (phase_piece_uniform).
template <int CC, bool UNI>
PB_DEV void simple_call(VoiceState& v, CallCtx& cc, const GroupParams& gp, const DevBuffer& buf, uint32_t n,
                        uint32_t call_off, TileRec* __restrict__ my_recs, uint32_t base, uint32_t gen, const bool cyc = false) {
  const long long cy0 = cyc ? CYC_T() : 0ll;
  long long cy_loop = 0;
  cc.call_left = cc.chunk_left;
  loop_range_samples(v, buf, cc.ls, cc.le);
  cc.new_call = false;
  if (!v.initialized) {  // CubicInterpolator::process prologue (cubic.rs:60-69)
    v.initialized = 1;
    v.hidx[3] = v.hidx[0];
    v.hidx[2] = (int32_t)v.playback_pos; v.hidx[1] = (int32_t)(v.playback_pos + CC); v.hidx[0] = (int32_t)(v.playback_pos + 2 * CC);
    v.playback_pos += 3 * CC;
  }
  const float ratio = v.ratio;
  const bool env = gp.has_env && cc.env_per_frame;
  float s = v.sub_pos, p = 0.0f;
  const PhaseK pk = phase_consts(ratio);  // the ratio is constant for the whole call
  bool first = true;
  uint32_t np = 0, off = call_off, remaining = n;
  uint32_t piece = min(remaining, TILE - (off % TILE));
  for (;;) {
    bool fused = false;
    if (UNI) {
      float d = 0.0f, o = 0.0f;
      bool on_hold = false;
      if (env && v.env_stage != ENV_SUSTAIN && v.env_stage != ENV_IDLE && env_bare_steps(v, gp, d, on_hold) >= piece) {
        fused = true;
        o = on_hold ? v.env_hold : v.env_out;
      } else {
        d = 0.0f;
      }
      np += phase_piece_uniform(s, p, pk, piece, first, o, d);
      if (fused) { if (on_hold) v.env_hold = o; else v.env_out = o; }
      else if (env) env_chain(v, gp, piece);
      fused = true;  // this piece is done
    } else if (env && v.env_stage != ENV_SUSTAIN && v.env_stage != ENV_IDLE) {
      float d;
      bool on_hold;
      if (env_bare_steps(v, gp, d, on_hold) >= piece) {  // the stage cannot end inside this piece: ride along
        float o = on_hold ? v.env_hold : v.env_out;
        const long long c0 = cyc ? CYC_T() : 0ll;
        np += phase_piece<true>(s, p, pk, piece, first, o, d);
        if (cyc) cy_loop += CYC_T() - c0;
        if (on_hold) v.env_hold = o; else v.env_out = o;
        fused = true;
      }
    }
    if (!fused) {
      float o_unused = 0.0f;
      const long long c0 = cyc ? CYC_T() : 0ll;
      np += phase_piece<false>(s, p, pk, piece, first, o_unused, 0.0f);
      if (cyc) cy_loop += CYC_T() - c0;
      if (env) env_chain(v, gp, piece);
    }
    first = false;
    off += piece; remaining -= piece;
    if (remaining == 0) break;
    piece = min(remaining, TILE);
    uint4 lo, hi;
    lo.x = v.playback_pos + np * CC; lo.y = __float_as_uint(s); lo.z = __float_as_uint(v.env_out); lo.w = __float_as_uint(v.env_hold);
    hi.x = __float_as_uint(v.env_target); hi.y = ((uint32_t)v.env_stage << 16) | piece; hi.z = base; hi.w = gen;
    uint4* dst = reinterpret_cast<uint4*>(my_recs + off / TILE);
    dst[0] = lo; dst[1] = hi;
  }
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
  cc.produced_in_call = n;
  cc.call_left = 0;
  cc.chunk_left = 0;
  after_process_call(v, cc);
#ifdef PB200_CYC
  if (cyc) { g_cyc[0] += CYC_T() - cy0; g_cyc[1] += cy_loop; g_cyc[2] += 1; g_cyc[3] += n; }
#endif
}

// WPV (warp per voice): lane 0 of warp i owns voice i, so voices never serialise each other's divergent
// control flow (the skeleton is a latency-bound chain of dependent f32 ops per voice, not a SIMT workload).
template <int MAXT, bool WPV>
PB_DEV void skeleton_block(const SkeletonArgs& a) {
  __shared__ VoiceHeader s_head[WPV ? 32 : MAXT];
  __shared__ GroupState s_gs;
  __shared__ uint32_t s_count;
  __shared__ uint64_t s_bounds[SB_MAX];   // this mixer's chunk boundaries of the block
  __shared__ uint32_t s_cnt[MAX_RUN];      // voices still holding a note after each chunk of a free run
  unsigned long long prof_free = 0;
  unsigned long long prof_q[4] = {0, 0, 0, 0};

  const long long cyc_block0 = CYC_T();
  const uint32_t g = a.group_list[blockIdx.x];
  const uint32_t tid = WPV ? ((threadIdx.x & 31u) == 0 ? (threadIdx.x >> 5) : 0xFFFFu) : threadIdx.x;
  const GroupParams gp = a.groups[g];
  const uint32_t nv = gp.n_voices;
  const DevBuffer buf = a.buffers[gp.buffer];
  const bool is_sampler = gp.kind == GROUP_SAMPLER;
  const bool mine = tid < nv;
  const float comp = a.rc.rate_comp;
  const uint32_t out_rate = a.rc.sample_rate;

  VoiceState v;
  if (mine) {
    v = a.voices[gp.first_voice + tid];
    publish_header(s_head, tid, v);
  }
  if (tid == 0) s_gs = a.gstate[g];

  // per-voice / per-group segment tables of this block
  const uint32_t vidx = gp.first_voice + tid;
  const bool is_hq = mine && v.hq != 0;
  HqState* const hqp = is_hq ? a.hq_states + vidx : nullptr;
  HqEmit hq_em;
  hq_em.recs = a.hq_recs; hq_em.n_recs = a.hq_n_recs; hq_em.cap = a.hq_cap; hq_em.buffer = gp.buffer;
  if (is_hq) hqp->rec = HQ_NONE;  // records are per time block
  const bool is_gran = a.gran_groups != nullptr && a.gran_groups[g].enabled != 0;
  const uint32_t gran_row = is_gran ? a.gran_groups[g].first_row + tid : 0u;
  GranState* const gsp = (is_gran && mine) ? a.gran_states + gran_row : nullptr;
  // SamplerVoice::stop (voice.rs:196-219) incl. the grain pool of a voice without envelope
  auto stop_voice = [&](const uint64_t frame) {
    if (gsp && v.has_note && !gp.has_env) gsp->trigger_new = 0;
    sampler_voice_stop(v, gp, frame);
  };
  // SamplerVoice::process epilogue (voice.rs:488-502): `frame` / `off` = absolute / block-relative frame after the call
  auto voice_epilogue = [&](const uint64_t frame, const uint32_t off) {
    bool reset = v.finished || (gp.has_env && v.env_stage == ENV_IDLE);
    if (gsp && !reset) reset = !gsp->trigger_new && gsp->max_end <= frame;  // GrainPool::is_exhausted
    if (reset) {
      if (gsp && v.has_note) gran_reset(*gsp, a.gran, frame, off);
      voice_reset(v);
    }
  };
  Segment* my_segs = a.segs + (size_t)vidx * a.seg_cap;
  uint16_t* my_first = a.seg_first + (size_t)vidx * a.n_tiles;
  uint16_t* my_count = a.seg_count + (size_t)vidx * a.n_tiles;
  TileRec* my_recs = a.recs + (size_t)vidx * a.n_tiles;
  uint32_t n_segs = 0, cur_tile = 0xFFFFFFFFu, cur_first = 0, cur_cnt = 0;
  if (WPV) {  // the whole warp clears its voice's per-tile segment counts
    const uint32_t wv = threadIdx.x >> 5;
    if (wv < nv) {
      uint16_t* cnt = a.seg_count + (size_t)(gp.first_voice + wv) * a.n_tiles;
      for (uint32_t i = threadIdx.x & 31u; i < a.n_tiles; i += 32) cnt[i] = 0;
    }
  } else if (mine) {
    for (uint32_t i = 0; i < a.n_tiles; ++i) my_count[i] = 0;
  }
  GroupSeg* g_segs = a.gsegs + (size_t)g * a.seg_cap;
  uint16_t* g_first = a.gseg_first + (size_t)g * a.n_tiles;
  uint16_t* g_count = a.gseg_count + (size_t)g * a.n_tiles;
  uint32_t n_gsegs = 0, gcur_tile = 0xFFFFFFFFu, gcur_first = 0, gcur_cnt = 0;
  for (uint32_t i = threadIdx.x; i < a.n_tiles; i += blockDim.x) g_count[i] = 0;
  __syncthreads();

  const uint32_t cb = a.mixer_chunk_begin[gp.mixer], ce = a.mixer_chunk_begin[gp.mixer + 1];
  for (uint32_t i = threadIdx.x; i < min(ce - cb, SB_MAX); i += blockDim.x) s_bounds[i] = a.chunk_bounds[cb + i];
  __syncthreads();
  auto bound = [&](const uint32_t i) -> uint64_t { return i - cb < SB_MAX ? s_bounds[i - cb] : a.chunk_bounds[i]; };
  // time of the next pending event (re-read only when the cursor moved)
  // next event that needs all voices at the same frame (see the free run below): its index and time
  uint32_t hard_idx = 0xFFFFFFFFu;
  uint64_t hard_time = UINT64_MAX;
  uint8_t* gflags = a.group_flags + (size_t)g * a.max_chunks;
  uint64_t my_frames = 0;

  // One Source::write call of this thread's voice (already opened by voice_begin_call): emits the call's
  // Segment / TileRec checkpoints and advances the control state. Returns the frames written.
  auto run_call = [&](CallCtx& cc, const uint32_t n, const uint32_t call_off, const uint64_t t) __attribute__((always_inline)) -> uint32_t {
    uint32_t written_frames = 0;
    bool simple = false;
    const bool cyc_v = CYC_ON(gp.first_voice + tid);
    const long long cyr0 = cyc_v ? CYC_T() : 0ll;
    if (n_segs < a.seg_cap && !(a.debug_flags & 2u) && !is_hq && !is_gran)
      simple = buf.channels == 2 ? simple_call_ok<2>(v, cc, buf, n) : simple_call_ok<1>(v, cc, buf, n);
    if (simple) {
      const uint32_t tile = call_off / TILE;
      if (!(a.debug_flags & 1u)) {
        Segment& s = my_segs[n_segs];
        s.v = v; s.c = cc; s.out_off = call_off; s.n = min(n, (tile + 1) * TILE - call_off);
      }
      if (tile != cur_tile) {
        if (cur_tile != 0xFFFFFFFFu) { my_first[cur_tile] = (uint16_t)cur_first; my_count[cur_tile] = (uint16_t)cur_cnt; }
        cur_tile = tile; cur_first = n_segs; cur_cnt = 0;
      }
      cur_cnt++;
      if (buf.channels == 2) simple_call<2, !WPV>(v, cc, gp, buf, n, call_off, my_recs, n_segs, a.gen, CYC_ON(gp.first_voice + tid));
      else simple_call<1, !WPV>(v, cc, gp, buf, n, call_off, my_recs, n_segs, a.gen, CYC_ON(gp.first_voice + tid));
      n_segs++;
      written_frames = n;
    } else {
      uint32_t off = call_off, remaining = n;
      while (remaining > 0 && !cc.ended) {
        const uint32_t tile = off / TILE;
        const uint32_t seg_len = min(remaining, (tile + 1) * TILE - off);
        if (n_segs < a.seg_cap && !(a.debug_flags & 1u)) {
          Segment& s = my_segs[n_segs];
          s.v = v; s.c = cc; s.out_off = off; s.n = seg_len;
          // per-tile (first, count) live in registers and are stored when the tile changes: no global
          // load sits on this latency-critical path
          if (tile != cur_tile) {
            if (cur_tile != 0xFFFFFFFFu) { my_first[cur_tile] = (uint16_t)cur_first; my_count[cur_tile] = (uint16_t)cur_cnt; }
            cur_tile = tile; cur_first = n_segs; cur_cnt = 0;
          }
          cur_cnt++;
          n_segs++;
        }
        uint32_t w;
        if (is_gran) {  // granular arm of SamplerVoice::process (voice.rs:412-427): the grain pool always fills the call
          gran_advance(gsp, a.gran_groups + g, a.gran, gran_row, t + (off - call_off), off, seg_len);
          w = seg_len;
          cc.chunk_left -= w; cc.hq_off += w;
          if (gp.has_env && cc.env_per_frame) env_chain(v, gp, w);
        } else if (is_hq) {
          // the out-of-line HighQuality state machine works on copies: taking the address of `v` / `cc` themselves
          // would move the hot cubic path's voice state from registers to local memory
          VoiceState vt = v;
          CallCtx ct = cc;
          w = buf.channels == 2 ? hq_advance<2>(vt, ct, hqp, hq_em, buf, comp, seg_len) : hq_advance<1>(vt, ct, hqp, hq_em, buf, comp, seg_len);
          v = vt; cc = ct;
        }
        else if (buf.channels == 2) w = voice_advance<2>(v, cc, gp, buf, out_rate, comp, seg_len);
        else w = voice_advance<1>(v, cc, gp, buf, out_rate, comp, seg_len);
        written_frames += w;
        off += w; remaining -= w;
        if (w < seg_len) break;
      }
    }
    my_frames += written_frames;
    voice_end_call(v, cc, t + n);
#ifdef PB200_CYC
    if (cyc_v) { g_cyc[4] += CYC_T() - cyr0; g_cyc[5] += simple ? 0 : n; g_cyc[6] += simple ? 0 : (unsigned long long)(CYC_T() - cyr0); }
#endif
    return written_frames;
  };
  // generator-level gain / pan (player.rs:1075-1081) of one call: checkpoint per (call x tile), advance ramps (thread 0)
  auto group_call = [&](const uint32_t n, const uint32_t call_off) {
    const bool vol_ramp = exp_need_ramp(s_gs.vol, comp);
    const bool vol_scale = !vol_ramp && fabsf(1.0f - s_gs.vol.target) > 0.000001f;
    const bool pan_ramp = exp_need_ramp(s_gs.pan, comp);
    const bool pan_apply = !pan_ramp && fabsf(s_gs.pan.target) > 0.000001f;
    const uint32_t flags = (vol_ramp ? 1u : 0u) | (vol_scale ? 2u : 0u) | (pan_ramp ? 4u : 0u) | (pan_apply ? 8u : 0u);
    if (flags == 0) return;  // unity gain, centre pan: the replay's default
    uint32_t off = call_off, remaining = n;
    while (remaining > 0) {
      const uint32_t tile = off / TILE;
      const uint32_t seg_len = min(remaining, (tile + 1) * TILE - off);
      if (n_gsegs < a.seg_cap) {
        GroupSeg& s = g_segs[n_gsegs];
        s.vol = s_gs.vol; s.pan = s_gs.pan; s.out_off = off; s.n = seg_len; s.flags = flags;
        if (tile != gcur_tile) {
          if (gcur_tile != 0xFFFFFFFFu) { g_first[gcur_tile] = (uint16_t)gcur_first; g_count[gcur_tile] = (uint16_t)gcur_cnt; }
          gcur_tile = tile; gcur_first = n_gsegs; gcur_cnt = 0;
        }
        gcur_cnt++;
        n_gsegs++;
      }
      if (vol_ramp) for (uint32_t i = 0; i < seg_len * 2; ++i) (void)exp_next(s_gs.vol, comp);
      if (pan_ramp) for (uint32_t i = 0; i < seg_len; ++i) (void)exp_next(s_gs.pan, comp);
      off += seg_len; remaining -= seg_len;
    }
  };

  for (uint32_t k = cb; k + 1 < ce; ++k) {
    const uint64_t c0 = bound(k), c1 = bound(k + 1);
    // ---- free run -------------------------------------------------------------------------------------
    // Between two note-ons of a Sampler nothing couples its voices: each one runs its own write calls through
    // consecutive chunks without a CTA barrier and only reports whether it still holds a note after each
    // chunk; thread 0 then replays the generator-level bookkeeping (active-voice count, stopped / dead,
    // generator gain/pan checkpoints, chunk flags) for those chunks in order. A voice holding a note implies
    // the generator writes (sampler.rs:978-981), so the voices need no group state while they run.
    // Only NOTE_ON (voice allocation looks at every voice, sampler.rs:826-860) and STOP (sets `stopping`, which
    // makes later events ignored, sampler.rs:664) need all voices at the same frame. Every other event is applied
    // inside the run: note-addressed events by the voice that holds the note (ids are unique), all-notes-off by
    // every voice, generator volume / panning by thread 0 during the bookkeeping.
    if (is_sampler) {
      if (hard_idx == 0xFFFFFFFFu || hard_idx < s_gs.ev_cursor) {
        hard_idx = s_gs.ev_cursor;
        while (hard_idx < gp.ev_end && a.events[hard_idx].kind != EVK_NOTE_ON && a.events[hard_idx].kind != EVK_STOP) ++hard_idx;
        hard_time = hard_idx < gp.ev_end ? a.events[hard_idx].time : UINT64_MAX;
      }
      uint32_t run = 0;
      if (!s_gs.dead && gp.start_time <= c0) {
        while (k + run + 1 < ce && run < MAX_RUN) {
          const uint64_t r0 = bound(k + run), r1 = bound(k + run + 1);
          if (hard_time <= r0) break;                               // a note-on / stop is due at this chunk
          if (s_gs.has_stop_time && s_gs.stop_time < r1) break;     // scheduled stop inside this chunk
          ++run;
        }
      }
      if (run) {
        const long long pq0 = a.prof ? clock64() : 0;
        for (uint32_t j = threadIdx.x; j < run; j += blockDim.x) s_cnt[j] = 0;
        __syncthreads();
        const long long prof_fr0 = a.prof ? clock64() : 0;
        if (a.prof) prof_q[0] += prof_fr0 - pq0;
        if (mine) {
          uint32_t ec = s_gs.ev_cursor;
          const bool ignore = s_gs.stopping != 0;
          for (uint32_t j = 0; j < run; ++j) {
            {  // events due at this chunk's start (MixedSource::process_events -> the first write call of the chunk)
              const uint64_t e0 = bound(k + j);
              while (ec < gp.ev_end && a.events[ec].time <= e0) {
                const DevEvent ev = a.events[ec];
                ++ec;
                if (ignore || ev.kind == EVK_SET_VOLUME || ev.kind == EVK_SET_PANNING) continue;
                if (ev.kind == EVK_ALL_NOTES_OFF) { stop_voice(e0); continue; }
                if (!(v.has_note && v.note_id == ev.note_id)) continue;
                if (ev.kind == EVK_NOTE_OFF) stop_voice(e0);
                else if (ev.kind == EVK_NOTE_SPEED) { file_set_speed(v, ev.speed, ev.glide, buf.sample_rate, out_rate); if (gsp) gsp->speed = ev.speed; }
                else if (ev.kind == EVK_NOTE_VOLUME) { v.note_volume = ev.value; exp_set_target(v.vol, gp.base_volume * ev.value, comp); if (gsp) gsp->volume = gp.base_volume * ev.value; }
                else if (ev.kind == EVK_NOTE_PANNING) {
                  v.note_panning = ev.value;
                  const float eff = fminf(fmaxf(gp.base_panning + ev.value, -1.0f), 1.0f);
                  exp_set_target(v.pan, eff, comp);
                  if (gsp) gsp->panning = eff;
                }
              }
            }
            if (v.has_note) {
              const uint64_t r0 = bound(k + j);
              const uint32_t rlen = (uint32_t)(bound(k + j + 1) - r0);
              CallCtx cc;
              cc.ended = true; cc.chunk_left = 0; cc.fader_running = false;
              if (voice_begin_call(v, cc, gp, buf, rlen, comp, gp.has_env != 0, (uint32_t)(r0 - a.block_start))) run_call(cc, rlen, (uint32_t)(r0 - a.block_start), r0);
              voice_epilogue(r0 + rlen, (uint32_t)(r0 - a.block_start) + rlen);
              if (v.has_note) atomicAdd(&s_cnt[j], 1u);
            }
          }
          publish_header(s_head, tid, v);
        }
        const long long pq1 = a.prof ? clock64() : 0;
        if (a.prof) prof_free += pq1 - prof_fr0;
        __syncthreads();
        const long long pq2 = a.prof ? clock64() : 0;
        if (a.prof) prof_q[1] += pq2 - pq1;
        if (tid == 0) {
          uint32_t ec = s_gs.ev_cursor;
          for (uint32_t j = 0; j < run; ++j) {
            {  // generator-level AmplifiedSource / PannedSource messages due at this chunk's start
              const uint64_t e0 = bound(k + j);
              while (ec < gp.ev_end && a.events[ec].time <= e0) {
                const DevEvent ev = a.events[ec];
                ++ec;
                if (ev.kind == EVK_SET_VOLUME) exp_set_target(s_gs.vol, ev.value, comp);
                else if (ev.kind == EVK_SET_PANNING) exp_set_target(s_gs.pan, ev.value, comp);
              }
            }
            bool writes = false;
            if (!s_gs.dead) {
              writes = !(s_gs.stopped || (s_gs.active_voices == 0 && !s_gs.stopping));
              if (writes) {
                const uint64_t r0 = bound(k + j);
                group_call((uint32_t)(bound(k + j + 1) - r0), (uint32_t)(r0 - a.block_start));
                s_gs.active_voices = s_cnt[j];
                if (s_gs.stopping && s_cnt[j] == 0) s_gs.stopped = 1;
              }
              if (gp.transient && s_gs.stopped && !s_gs.dead) { s_gs.dead = 1; s_gs.dead_time = bound(k + j + 1); }
            }
            gflags[k + j - cb] = writes ? 1 : 0;
          }
          s_gs.ev_cursor = ec;
        }
        const long long pq3 = a.prof ? clock64() : 0;
        if (a.prof) prof_q[2] += pq3 - pq2;
        __syncthreads();
        if (a.prof) prof_q[3] += clock64() - pq3;
        k += run - 1;
        continue;
      }
    }
    const uint32_t len = (uint32_t)(c1 - c0);
    const uint32_t boff = (uint32_t)(c0 - a.block_start);
    bool produced = false;
    uint32_t total = 0;
    bool skip = s_gs.dead != 0;
    // MixedSource::process_sources (mixed.rs:558-624) for this one source
    if (!skip && gp.start_time > c0) {
      uint64_t until = gp.start_time - c0;
      if (until >= len) skip = true;
      else total = (uint32_t)until;
    }
    if (skip) {
      if (tid == 0) gflags[k - cb] = 0;
      continue;
    }

    while (total < len) {
      const uint64_t t = c0 + total;
      uint64_t until_stop = UINT64_MAX;
      __syncthreads();
      if (s_gs.has_stop_time) until_stop = s_gs.stop_time > t ? s_gs.stop_time - t : 0;
      bool send_stop = false;
      if (until_stop == 0) { send_stop = true; until_stop = UINT64_MAX; }
      __syncthreads();
      if (send_stop && tid == 0) s_gs.has_stop_time = 0;
      const uint32_t n = (uint32_t)min((uint64_t)(len - total), until_stop);

      // ---- Source::write(n frames at time t) -------------------------------------------------------
      // 1. messages in queue order: the events process_events pushed at this chunk's start (consumed by
      //    the first write call of the chunk), then a Stop the mixer force-pushed at stop_time.
      uint32_t ev_end_now = s_gs.ev_cursor;
      while (ev_end_now < gp.ev_end && a.events[ev_end_now].time <= c0) ++ev_end_now;
      for (uint32_t e = s_gs.ev_cursor; e < ev_end_now; ++e) {
        const DevEvent ev = a.events[e];
        if (is_sampler) {
          const bool ignore = s_gs.stopping != 0;  // sampler.rs:664
          if (ev.kind == EVK_STOP) {
            // GeneratorPlaybackMessage::Stop (sampler.rs:733-738)
            __syncthreads();
            if (tid == 0) s_gs.stopping = gp.transient;
            if (mine) { stop_voice(t); publish_header(s_head, tid, v); }
            __syncthreads();
          } else if (ev.kind == EVK_SET_VOLUME) {  // generator-level AmplifiedSource message
            if (tid == 0) exp_set_target(s_gs.vol, ev.value, comp);
          } else if (ev.kind == EVK_SET_PANNING) {
            if (tid == 0) exp_set_target(s_gs.pan, ev.value, comp);
          } else if (!ignore) {
            if (ev.kind == EVK_NOTE_ON) {
              const uint32_t idx = next_free_voice_index(s_head, nv, gp.has_env != 0);
              __syncthreads();
              if (tid == idx) {  // SamplerVoice::start (voice.rs:122-193)
                if (gsp && v.has_note) gran_reset(*gsp, a.gran, t, boff + total);
                voice_reset(v);
                v.note = (uint8_t)ev.note;
                v.note_volume = ev.value;
                v.note_panning = ev.value2;
                float eff_vol = gp.base_volume * ev.value;
                float eff_pan = fminf(fmaxf(gp.base_panning + ev.value2, -1.0f), 1.0f);
                file_set_speed(v, ev.speed, 0.0f, buf.sample_rate, out_rate);
                exp_set_target(v.vol, eff_vol, comp);
                exp_set_target(v.pan, eff_pan, comp);
                if (gsp) gran_start(*gsp, a.gran_groups[g], ev.speed, eff_vol, eff_pan);
                if (gp.has_env) env_note_on(v, gp, 1.0f);
                v.has_note = 1;
                v.note_id = ev.note_id;
                publish_header(s_head, tid, v);
              }
              if (tid == 0) s_gs.active_voices += 1;
              __syncthreads();
            } else if (ev.kind == EVK_ALL_NOTES_OFF) {
              if (mine) { stop_voice(t); publish_header(s_head, tid, v); }
              __syncthreads();
            } else {
              // note-addressed events: first voice whose note_id matches (sampler.rs:776-822)
              uint32_t idx = 0xFFFFFFFFu;
              for (uint32_t i = 0; i < nv; ++i)
                if (s_head[i].active && s_head[i].note_id == ev.note_id) { idx = i; break; }
              __syncthreads();
              if (tid == idx) {
                if (ev.kind == EVK_NOTE_OFF) stop_voice(t);
                else if (ev.kind == EVK_NOTE_SPEED) { file_set_speed(v, ev.speed, ev.glide, buf.sample_rate, out_rate); if (gsp) gsp->speed = ev.speed; }
                else if (ev.kind == EVK_NOTE_VOLUME) { v.note_volume = ev.value; exp_set_target(v.vol, gp.base_volume * ev.value, comp); if (gsp) gsp->volume = gp.base_volume * ev.value; }
                else if (ev.kind == EVK_NOTE_PANNING) {
                  v.note_panning = ev.value;
                  const float eff = fminf(fmaxf(gp.base_panning + ev.value, -1.0f), 1.0f);
                  exp_set_target(v.pan, eff, comp);
                  if (gsp) gsp->panning = eff;
                }
                publish_header(s_head, tid, v);
              }
              __syncthreads();
            }
          }
        } else if (mine) {  // file playback: FilePlaybackMessage / Amplified / Panned messages
          if (ev.kind == EVK_STOP) file_stop(v, gp);
          else if (ev.kind == EVK_SET_SPEED) file_set_speed(v, ev.speed, ev.glide, buf.sample_rate, out_rate);
          else if (ev.kind == EVK_SEEK) {
            if (is_hq && !v.finished) hq_reset_pending(*hqp, hq_em, boff + total);
            file_seek(v, ev.seek_pos);
          }
          else if (ev.kind == EVK_SET_VOLUME) exp_set_target(v.vol, ev.value, comp);
          else if (ev.kind == EVK_SET_PANNING) exp_set_target(v.pan, ev.value, comp);
        }
      }
      __syncthreads();
      if (tid == 0) s_gs.ev_cursor = ev_end_now;
      if (send_stop) {  // PlaybackMessageQueue::send_stop (mixed.rs:591-598)
        if (is_sampler) {
          if (tid == 0) s_gs.stopping = gp.transient;
          if (mine) { stop_voice(t); publish_header(s_head, tid, v); }
        } else if (mine) {
          file_stop(v, gp);
        }
      }
      __syncthreads();

      // 2. does the source write at all? (sampler.rs:978-981 / preloaded.rs:400-403)
      bool group_writes;
      if (is_sampler) group_writes = !(s_gs.stopped || (s_gs.active_voices == 0 && !s_gs.stopping));
      else group_writes = true;
      const bool was_active = mine && (is_sampler ? v.has_note != 0 : true);
      CallCtx cc;
      cc.ended = true; cc.chunk_left = 0; cc.fader_running = false;
      bool call_open = false;
      const uint32_t call_off = boff + total;  // first frame of the call, relative to the block
      if (group_writes && was_active) call_open = voice_begin_call(v, cc, gp, buf, n, comp, gp.has_env != 0, call_off);
      __syncthreads();

      // 3. advance the voice through the call, one segment per (call x 64-frame tile)
      uint32_t written_frames = 0;
      if (call_open) written_frames = run_call(cc, n, call_off, t);
      if (is_sampler && group_writes && tid == 0) group_call(n, call_off);
      uint32_t written;
      if (is_sampler) {
        written = group_writes ? n : 0;
        if (was_active && group_writes) {
          voice_epilogue(t + n, call_off + n);
          publish_header(s_head, tid, v);
        }
        if (tid == 0) s_count = 0;
        __syncthreads();
        if (group_writes) {
          if (mine && v.has_note) atomicAdd(&s_count, 1u);
          __syncthreads();
          if (tid == 0) {
            s_gs.active_voices = s_count;
            if (s_gs.stopping && s_count == 0) s_gs.stopped = 1;
          }
        }
        __syncthreads();
      } else {
        // a file source's written count = frames its single voice produced (short on EOF)
        if (tid == 0) s_count = written_frames;
        __syncthreads();
        written = s_count;
        __syncthreads();
      }
      total += written;
      produced |= written > 0;
      // mixed.rs:612-619
      bool exhausted;
      if (is_sampler) exhausted = s_gs.stopped != 0;
      else { if (tid == 0) s_count = v.finished; __syncthreads(); exhausted = s_count != 0; __syncthreads(); }
      if (gp.transient && exhausted) {
        if (tid == 0) { s_gs.dead = 1; s_gs.dead_time = c1; }
        __syncthreads();
        break;
      } else if (written == 0) {
        break;
      }
    }
    if (tid == 0) gflags[k - cb] = produced ? 1 : 0;
    __syncthreads();
  }

  if (mine && cur_tile != 0xFFFFFFFFu) { my_first[cur_tile] = (uint16_t)cur_first; my_count[cur_tile] = (uint16_t)cur_cnt; }
  if (tid == 0 && gcur_tile != 0xFFFFFFFFu) { g_first[gcur_tile] = (uint16_t)gcur_first; g_count[gcur_tile] = (uint16_t)gcur_cnt; }
  if (mine) a.voices[gp.first_voice + tid] = v;
#ifdef PB200_CYC
  if (CYC_ON(gp.first_voice + tid) && mine) g_cyc[7] += CYC_T() - cyc_block0;
#endif
  if (a.prof && mine) {
    unsigned long long* pr = a.prof + (size_t)(gp.first_voice + tid) * 4;
    pr[0] += prof_q[0]; pr[1] += prof_q[1]; pr[2] += prof_q[3]; pr[3] += prof_free;
  }
  if (my_frames) atomicAdd((unsigned long long*)&s_gs.voice_frames, (unsigned long long)my_frames);
  __syncthreads();
  if (tid == 0) a.gstate[g] = s_gs;
}

// How one launch walks several consecutive time blocks (the persistent mode of small graphs: every group keeps
// its own pace through the whole render, a block is handed to the replay pass as soon as ALL groups have finished
// it; renderer.cu waits for `block_done[b]` with a stream memory operation). n_blocks = 1 and block_done = nullptr:
// the plain one-launch-per-block mode.
struct SkeletonLoop {
  uint32_t n_blocks;
  uint32_t chunk_begin_stride;   // uint32 per block in mixer_chunk_begin
  size_t group_flags_stride;     // per-block strides of the snapshot tables (elements)
  size_t segs_stride, seg_tab_stride, gsegs_stride, gseg_tab_stride, recs_stride;
  uint32_t* block_done;          // [n_blocks] CTAs that finished the block
};

template <int MAXT, bool WPV>
__global__ void __launch_bounds__(MAXT) skeleton_kernel(SkeletonArgs a, SkeletonLoop L) {
  for (uint32_t b = 0; b < L.n_blocks; ++b) {
    skeleton_block<MAXT, WPV>(a);
    if (L.block_done) {
      __syncthreads();
      if (threadIdx.x == 0) { __threadfence(); atomicAdd(L.block_done + b, 1u); }
    }
    a.mixer_chunk_begin += L.chunk_begin_stride;
    a.group_flags += L.group_flags_stride;
    a.block_start += a.block_frames;
    a.segs += L.segs_stride; a.seg_first += L.seg_tab_stride; a.seg_count += L.seg_tab_stride;
    a.gsegs += L.gsegs_stride; a.gseg_first += L.gseg_tab_stride; a.gseg_count += L.gseg_tab_stride;
    a.recs += L.recs_stride;
    a.gen += 1;
    __syncthreads();
  }
}

}  // namespace pb
