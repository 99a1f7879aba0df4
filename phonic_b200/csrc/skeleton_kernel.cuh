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
#include "phase_table.cuh"

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
  // PlaybackStatusEvent stream of file playbacks (nullptr: nobody listens)
  StatusRec* status;
  uint32_t* status_count;
  uint32_t* overflow;        // set when a voice's / group's snapshot list filled up (records may have been dropped)
  uint32_t status_cap;
  uint32_t pos_emit_rate;            // frames between Position events (1 s)
  // exact 64-frame phase jumps (phase_table.cuh): one table per steady ratio of the graph, directory sorted by ratio bits
  const uint32_t* phase_tabs;
  const uint2* phase_dir;             // (f32 bits of the ratio, word offset of its table)
  uint32_t n_phase;
  unsigned long long* prof;           // PB200_SKEL_PROF: [n_voices][4] cycles waiting at the free run's three barriers + its own work
  uint32_t debug_flags;  // timing experiments only (PB200_SKEL_DEBUG): 1 = no snapshot stores, 2 = no simple calls, 4 = no phase jumps
};

constexpr int VK_MAX_VOICES = 1024;
// speed_from_note (src/utils.rs:67-78) of the 128 MIDI notes, computed once on the host with the libm the reference's
// `powf` calls go to: a sampler pitch parameter change re-derives the speed of every sounding voice from its note
__constant__ double c_note_speed[128];
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
constexpr uint32_t MAX_RUN = 256;  // chunks per free run (fewer joins: the voices of a group take turns being the slowest)

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


// The jump table of `ratio`, or nullptr when the host did not foresee this ratio (a glide in progress): literal loop.
PB_DEV const uint32_t* find_phase_tab(const SkeletonArgs& a, const float ratio) {
  const uint32_t bits = (uint32_t)__float_as_int(ratio);
  uint32_t lo = 0, hi = a.n_phase;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint32_t k = a.phase_dir[mid].x;
    if (k == bits) return a.phase_tabs + a.phase_dir[mid].y;
    if (k < bits) lo = mid + 1; else hi = mid;
  }
  return nullptr;
}

// ---- a voice's jump table staged in shared memory (warp-per-voice skeleton) ------------------------------------------
// The jump is a chain of three dependent table reads; from L2 that is ~1000 cycles per tile on a path where one thread
// per voice works alone. The voice's warp therefore keeps the body of its current table in its own 16 KB slot of shared
// memory, fetched with ONE bulk async copy (cp.async.bulk, completion on an mbarrier) whenever the ratio changes.
constexpr uint32_t TAB_SLOT_WORDS = 4096;   // fits every table of ratio >= 1/8; larger ones are read from global memory
struct TabSlot { uint32_t* words; uint64_t* mbar; uint32_t* parity; };

PB_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
PB_DEV void tab_slot_init(const TabSlot& ts) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(ts.mbar)));
  *ts.parity = 0u;
}
PB_DEV void tab_slot_fetch(const TabSlot& ts, const uint32_t* __restrict__ src, const uint32_t bytes) {
  const uint32_t mbar = smem_u32(ts.mbar), dst = smem_u32(ts.words);
  // earlier generic-proxy reads of the slot are ordered before the async-proxy write
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
  const uint32_t parity = *ts.parity;
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  }
  *ts.parity = parity ^ 1u;
}

// Builds the tables of `dir[blockIdx.x]` (phase_table.cuh: analytic break points, every entry verified against the model).
__global__ void __launch_bounds__(256) phase_table_kernel(uint32_t* __restrict__ tabs, const uint2* __restrict__ dir) {
  __shared__ uint32_t scratch[PT_MAX_BP];
  const uint2 d = dir[blockIdx.x];
  const float ratio = __int_as_float((int)d.x);
  uint32_t* tab = tabs + d.y;
  const PhaseGeom g = phase_geom(ratio);
  if (threadIdx.x == 0) phase_build_header(tab, ratio, g);
  if (g.mode != PT_DOWN_TABLE && g.mode != PT_UP_TABLE) return;
  phase_build_p1(g, scratch, threadIdx.x, blockDim.x);
  __syncthreads();
  phase_build_p2(tab, g, scratch, threadIdx.x, blockDim.x);
  __syncthreads();
  phase_build_p3(tab, g, threadIdx.x, blockDim.x);
  __syncthreads();
  phase_build_p4(tab, g, threadIdx.x, blockDim.x);
}

// ---- simple calls and super-calls --------------------------------------------------------------------------------
// A simple call (voice.cuh "simple calls") advances only the phase recurrence and the envelope; the skeleton stores one
// full Segment at its first frame and a 32-byte TileRec at every tile boundary it crosses. A voice that stays steady --
// no event addressed to it, envelope in Sustain or far from its next threshold, input far from its loop end -- does not
// notice the mixer's chunk boundaries at all (they come from OTHER sources' events, mixed.rs:686-693): the following
// chunks are accepted into the open call ("super-call") without a new Segment, and whole tiles are advanced with one
// exact jump of the phase recurrence (phase_table.cuh) + a closed-form envelope step, however the chunk boundaries
// fall inside them. Frames of an unfinished tile stay pending until later chunks complete it or something ends the
// steady span (then they are advanced literally and the call is closed).
struct SuperCall {
  uint32_t open;
  uint32_t open_off;   // block-relative frame of the call's Segment
  uint32_t adv_off;    // phase / envelope state is advanced up to here (tile aligned once past the first piece)
  uint32_t end_off;    // chunks accepted up to here; end_off - adv_off frames are pending
  uint32_t np;         // samples pushed since the call opened (folded into playback_pos / hidx when it closes)
  uint32_t base;       // index of the call's Segment
  uint32_t budget;     // frames from adv_off for which the voice provably stays steady
  uint32_t le;         // end of the input range in samples (loop end / buffer end)
  float s, p;          // sub_pos at adv_off, look-ahead push flag of the literal loop
  bool first, env;
  PhaseRef tab;
};

// per-frame increment of the current envelope stage's bare accumulate chain (see env_bare_steps)
PB_DEV void env_increment(const VoiceState& v, const GroupParams& gp, float& d, bool& on_hold) {
  const uint32_t stage = v.env_stage;
  on_hold = false; d = 0.0f;
  if (stage == ENV_ATTACK) d = gp.attack_rate;
  else if (stage == ENV_HOLD) { d = -1.0f; on_hold = true; }
  else if (stage == ENV_DECAY && v.env_out > gp.sustain_level) d = -gp.decay_rate;
  else if (stage == ENV_RELEASE) d = -(v.env_release_out * gp.release_rate);
}

// The literal pieces and the envelope's stage machine are off the steady path: they live out of line (scalars in,
// scalars out -- nothing of the voice state has its address taken) so that the tile loop stays small.
struct PieceOut { float s, p, o; uint32_t np; };
template <bool UNI>
__device__ __noinline__ PieceOut phase_piece_nl(float s, float p, float o, const float d, const float ratio, const uint32_t span,
                                                const bool first, const bool acc) {
  const PhaseK pk = phase_consts(ratio);
  PieceOut r;
  if (UNI) r.np = phase_piece_uniform(s, p, pk, span, first, o, acc ? d : 0.0f);
  else if (acc) r.np = phase_piece<true>(s, p, pk, span, first, o, d);
  else r.np = phase_piece<false>(s, p, pk, span, first, o, 0.0f);
  r.s = s; r.p = p; r.o = o;
  return r;
}
struct EnvOut { float out, hold, target; uint32_t stage; };
__device__ __noinline__ EnvOut env_chain_nl(const float out, const float hold, const float target, const float release_out,
                                            const uint32_t stage, const GroupParams* __restrict__ gp, const uint32_t w) {
  VoiceState t;
  t.env_out = out; t.env_hold = hold; t.env_target = target; t.env_release_out = release_out; t.env_stage = (uint8_t)stage;
  const GroupParams g = *gp;
  env_chain(t, g, w);
  EnvOut r;
  r.out = t.env_out; r.hold = t.env_hold; r.target = t.env_target; r.stage = t.env_stage;
  return r;
}
PB_DEV void env_chain_call(VoiceState& v, const GroupParams* gp, const uint32_t w) {
  const EnvOut r = env_chain_nl(v.env_out, v.env_hold, v.env_target, v.env_release_out, v.env_stage, gp, w);
  v.env_out = r.out; v.env_hold = r.hold; v.env_target = r.target; v.env_stage = (uint8_t)r.stage;
}
// the general per-frame path of a call (glides, ramps, loop ends): works on copies like the HighQuality state machine
template <int CC>
__device__ __noinline__ uint32_t voice_advance_nl(VoiceState* v, CallCtx* c, const GroupParams* __restrict__ gp, const DevBuffer* __restrict__ b,
                                                  const uint32_t out_rate, const float comp, const uint32_t n) {
  VoiceState vt = *v;
  CallCtx ct = *c;
  const GroupParams g = *gp;
  const DevBuffer bb = *b;
  const uint32_t w = voice_advance<CC>(vt, ct, g, bb, out_rate, comp, n);
  *v = vt; *c = ct;
  return w;
}

PB_DEV void sc_open(SuperCall& sc, VoiceState& v, CallCtx& cc, const GroupParams& gp, const DevBuffer& buf, const uint32_t CC,
                    const uint32_t call_off, const uint32_t base, const PhaseRef& tab) {
  cc.call_left = cc.chunk_left;
  loop_range_samples(v, buf, cc.ls, cc.le);
  cc.new_call = false;
  if (!v.initialized) {  // CubicInterpolator::process prologue (cubic.rs:60-69)
    v.initialized = 1;
    v.hidx[3] = v.hidx[0];
    v.hidx[2] = (int32_t)v.playback_pos; v.hidx[1] = (int32_t)(v.playback_pos + CC); v.hidx[0] = (int32_t)(v.playback_pos + 2 * CC);
    v.playback_pos += 3 * CC;
  }
  sc.open = 1; sc.open_off = sc.adv_off = sc.end_off = call_off; sc.np = 0; sc.base = base; sc.budget = 0; sc.le = cc.le;
  sc.s = v.sub_pos; sc.p = 0.0f; sc.first = true; sc.env = gp.has_env && cc.env_per_frame;
  sc.tab = tab;
}

// Frames from adv_off for which nothing but the phase recurrence and a bare envelope chain can happen.
PB_DEV uint32_t sc_steady_budget(const SuperCall& sc, const VoiceState& v, const GroupParams& gp, const uint32_t CC) {
  uint32_t b = 0xFFFFFFFFu;
  if (sc.env && v.env_stage != ENV_SUSTAIN) {  // Sustain: env_run returns the constant level until a note-off
    if (v.env_stage == ENV_IDLE) return 0u;
    float d;
    bool on_hold;
    b = env_bare_steps(v, gp, d, on_hold);
  }
  const uint32_t pos = v.playback_pos + sc.np * CC;
  const uint32_t avail = sc.le > pos ? (sc.le - pos) / CC : 0u;
  const uint32_t per_frame = v.ratio < 1.0f ? 1u : (uint32_t)v.ratio + 2u;
  const uint32_t in_b = avail > 5u ? (avail - 5u) / per_frame : 0u;  // simple_call_ok: n * per_frame + 4 < avail
  return min(b, in_b);
}

PB_DEV void sc_store_rec(const SuperCall& sc, const VoiceState& v, TileRec* __restrict__ my_recs, const uint32_t gen, const uint32_t CC,
                         const uint32_t piece) {
  uint4 lo, hi;
  lo.x = v.playback_pos + sc.np * CC; lo.y = __float_as_uint(sc.s); lo.z = __float_as_uint(v.env_out); lo.w = __float_as_uint(v.env_hold);
  hi.x = __float_as_uint(v.env_target); hi.y = ((uint32_t)v.env_stage << 16) | piece; hi.z = sc.base; hi.w = gen;
  uint4* dst = reinterpret_cast<uint4*>(my_recs + sc.adv_off / TILE);
  dst[0] = lo; dst[1] = hi;
}

// Word `idx` of a table body: from the warp's staged copy (ld.shared, 32-bit address) or through the global pointer.
template <bool SMEM>
PB_DEV uint32_t pt_word(const PhaseRef& t, const uint32_t idx) {
  if (SMEM) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(t.smem_body + idx * 4u));
    return v;
  }
  return t.body[idx];
}

// The steady tile loop of a super-call: whole tiles past the first piece, the phase state kept as the integer S of
// phase_table.cuh across tiles (converted once on entry; the conversion doubles as the on-grid test). Stops at the
// first tile the jump declines or whose envelope is not provably a bare chain; sc_advance's general loop takes over.
template <bool SMEM>
PB_DEV void sc_fast_tiles(SuperCall& sc, VoiceState& v, const GroupParams& gp, TileRec* __restrict__ my_recs, const uint32_t gen,
                          const uint32_t CC, const uint32_t to_off, unsigned long long* __restrict__ stat) {
  const PhaseRef& t = sc.tab;
  const uint32_t mode = t.mode;
  if (mode == PT_LITERAL) return;
  const float x = sc.s * t.scale;
  if (!(x >= 0.0f) || !(x < 1073741824.0f)) return;
  uint32_t S = (uint32_t)x;
  if ((float)S != x) return;                                   // not on the ratio's grid (yet)
  if (mode == PT_DOWN_EXACT && (S & t.lmask)) return;
  const bool down = mode <= PT_DOWN_TABLE;
  // a ratio >= 1 whose states and R are all even never rounds: the exact rotation applies (see phase_jump)
  const bool table = (mode == PT_DOWN_TABLE) || (mode == PT_UP_TABLE && ((S | t.R) & 1u));
  const uint32_t sh = t.sh, ONE = 1u << sh, KR = t.KR, wrap = down ? ONE : 0xFFFFFFFFu;
  const uint32_t bpx = PT_COARSE + 1u, ent = (uint32_t)(t.entry - t.body);
  uint32_t np = sc.np, adv = sc.adv_off, budget = sc.budget, done = 0;
  const uint32_t pos0 = v.playback_pos;
  TileRec* __restrict__ rec = my_recs + adv / TILE;
  const uint32_t n_max = (to_off - adv) / TILE;
  // The loop body is branch-free up to its single exit test so that the record store and the envelope step fill the
  // issue slots of the S -> lookup -> S' dependency chain (one thread per voice: nothing else hides its latency).
  for (uint32_t it = 0; it < n_max; ++it) {
    const bool moving = sc.env && v.env_stage != ENV_SUSTAIN && v.env_stage != ENV_IDLE;
    if (moving && budget < TILE) break;
    const uint32_t w0 = S >= wrap ? 1u : 0u;                     // ratio < 1: the first frame's wrap
    const uint32_t Sr = S - (w0 << sh);
    bool ok = Sr < ONE;
    uint32_t SK, W;
    if (table) {
      const uint32_t c = pt_word<SMEM>(t, min(Sr, ONE - 1u) >> t.cshift);
      const uint32_t p0 = pt_word<SMEM>(t, bpx + c), n1 = pt_word<SMEM>(t, bpx + c + 1u), n2 = pt_word<SMEM>(t, bpx + min(c + 2u, t.n_bp + 1u));
      const bool step = n1 <= Sr;                                // at most one break point of the bucket lies below S ...
      const uint32_t i = c + (step ? 1u : 0u), next = step ? n2 : n1, prev = step ? n1 : p0;
      ok = ok && next > Sr;                                      // ... else the general loop takes this tile
      const uint32_t e = pt_word<SMEM>(t, ent + (min(i, t.n_bp) << t.lshift) + (Sr & t.lmask));
      ok = ok && (Sr - prev >= t.margin) && (next - Sr > t.margin) && (e & 0x80000000u);
      W = (e >> 16) & 0xFFFu;
      const uint32_t P = (e & 0xFFFFu) - 32768u;
      SK = down ? Sr + KR - (W << sh) + P : Sr + (W << sh) - KR + P;
    } else if (down) {
      W = (Sr + KR - t.R) >> sh;
      SK = Sr + KR - (W << sh);
    } else {
      SK = (Sr - KR) & (ONE - 1u);
      W = (KR + SK - Sr) >> sh;
    }
    {  // the tile's continuation record: the state at its first frame (rewritten by the general loop if we stop here)
      uint4 lo, hi;
      lo.x = pos0 + np * CC; lo.y = __float_as_uint((float)S * t.inv_scale); lo.z = __float_as_uint(v.env_out); lo.w = __float_as_uint(v.env_hold);
      hi.x = __float_as_uint(v.env_target); hi.y = ((uint32_t)v.env_stage << 16) | TILE; hi.z = sc.base; hi.w = gen;
      uint4* dst = reinterpret_cast<uint4*>(rec);
      dst[0] = lo; dst[1] = hi;
    }
    if (!ok) break;
    np += w0 + W; S = SK;
    if (moving) {
      float d;
      bool on_hold;
      env_increment(v, gp, d, on_hold);
      float o = on_hold ? v.env_hold : v.env_out;
      if (!accum_jump(o, d, TILE)) {
#pragma unroll 8
        for (uint32_t j = 0; j < TILE; ++j) o += d;
      }
      if (on_hold) v.env_hold = o; else v.env_out = o;
    }
    ++rec; ++done;
    budget = budget > TILE ? budget - TILE : 0u;
  }
  if (done) {
    sc.s = (float)S * t.inv_scale;  // at most 24 significant bits: exact
    sc.np = np; sc.adv_off = adv + done * TILE; sc.budget = budget;
    sc.first = true;                // the literal loop's look-ahead push flag is stale after a jump
    if (stat) stat[0] += done;
  }
}

// Advance the open call to `to_off`. !final: only whole tiles (the rest stays pending). UNI: lane-per-voice skeleton,
// one instruction stream for every ratio class in the literal pieces (phase_piece_uniform).
template <bool UNI>
PB_DEV void sc_advance(SuperCall& sc, VoiceState& v, const GroupParams& gp, const GroupParams* __restrict__ gp_mem, TileRec* __restrict__ my_recs,
                       const uint32_t gen, const uint32_t CC, const uint32_t to_off, const bool final, unsigned long long* __restrict__ stat) {
  // steady tiles: one exact jump of the phase recurrence + a closed-form step of a bare envelope chain
  if (sc.adv_off != sc.open_off && sc.adv_off + TILE <= to_off) {
    const long long f0 = stat ? clock64() : 0;
    if (sc.tab.smem_body) sc_fast_tiles<true>(sc, v, gp, my_recs, gen, CC, to_off, stat);
    else sc_fast_tiles<false>(sc, v, gp, my_recs, gen, CC, to_off, stat);
    if (stat) stat[4] += (unsigned long long)(clock64() - f0);
  }
  while (sc.adv_off < to_off) {
    const uint32_t tile_end = (sc.adv_off / TILE + 1u) * TILE;
    const uint32_t piece_end = min(tile_end, to_off);
    // the first piece belongs to the Segment (its length is stored there): it is never left pending
    if (piece_end < tile_end && !final && sc.adv_off != sc.open_off) break;
    const uint32_t piece = piece_end - sc.adv_off;
    if (sc.adv_off != sc.open_off) sc_store_rec(sc, v, my_recs, gen, CC, piece);  // a tile that opens inside the call
    const bool env_moving = sc.env && v.env_stage != ENV_SUSTAIN && v.env_stage != ENV_IDLE;
    bool bare = false, on_hold = false;
    float d = 0.0f;
    if (env_moving) {
      if (sc.budget >= piece) { bare = true; env_increment(v, gp, d, on_hold); }
      else bare = env_bare_steps(v, gp, d, on_hold) >= piece;
    }
    bool jumped = false;
    if (piece == TILE) {
      uint32_t w = 0;
      jumped = phase_jump(sc.tab, sc.s, w);
      sc.np += w;
    }
    if (jumped) {
      if (bare) {
        float o = on_hold ? v.env_hold : v.env_out;
        if (!accum_jump(o, d, TILE)) {
#pragma unroll 8
          for (uint32_t j = 0; j < TILE; ++j) o += d;
        }
        if (on_hold) v.env_hold = o; else v.env_out = o;
      } else if (env_moving) {
        env_chain_call(v, gp_mem, piece);
      }
    } else {
      const float o_in = bare ? (on_hold ? v.env_hold : v.env_out) : 0.0f;
      const PieceOut r = phase_piece_nl<UNI>(sc.s, sc.p, o_in, d, v.ratio, piece, sc.first, bare);
      sc.s = r.s; sc.p = r.p; sc.np += r.np;
      if (bare) { if (on_hold) v.env_hold = r.o; else v.env_out = r.o; }
      else if (env_moving) env_chain_call(v, gp_mem, piece);
    }
    if (stat) { if (jumped) stat[0] += 1; else stat[1] += piece; }
    sc.first = jumped;  // after a jump the look-ahead push flag of the literal loop is stale
    sc.adv_off = piece_end;
    sc.budget = sc.budget > piece ? sc.budget - piece : 0u;
  }
}

// Fold the call's pushes into the voice state (everything the call consumed is consecutive input).
PB_DEV void sc_close(SuperCall& sc, VoiceState& v, const uint32_t CC) {
  v.sub_pos = sc.s;
  const uint32_t np = sc.np;
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
  sc.open = 0;
}

// The granular arm of one Source::write call (SamplerVoice::process, voice.rs:412-427), out of line and on copies of the
// voice state: the per-tile loop calls gran_advance (itself out of line), and a call made from the skeleton's main body
// would save and restore that body's ~250 live registers around every tile.
struct SegCursor { uint32_t n_segs, cur_tile, cur_first, cur_cnt; };
__device__ __noinline__ void gran_call_nl(VoiceState* vp, CallCtx* cp, const GroupParams* __restrict__ gpp, GranState* gsp,
                                          const GranGroup* __restrict__ ggp, const GranEmit* __restrict__ emp, const uint32_t gran_row,
                                          const uint64_t t, const uint32_t call_off, const uint32_t n, Segment* __restrict__ my_segs,
                                          uint16_t* __restrict__ my_first, uint16_t* __restrict__ my_count, const uint32_t seg_cap,
                                          const uint32_t gp_idx, const bool store, SegCursor* cur) {
  VoiceState v = *vp;
  CallCtx cc = *cp;
  SegCursor c = *cur;
  const GroupParams gp = *gpp;
  const GranEmit em = *emp;
  uint32_t off = call_off, remaining = n;
  while (remaining > 0) {
    const uint32_t tile = off / TILE;
    const uint32_t seg_len = min(remaining, (tile + 1) * TILE - off);
    if (c.n_segs < seg_cap && store) {
      Segment& s = my_segs[c.n_segs];
      s.v = v; s.c = cc; s.out_off = off; s.n = seg_len; s.gp_idx = gp_idx;
      if (tile != c.cur_tile) {
        if (c.cur_tile != 0xFFFFFFFFu) { my_first[c.cur_tile] = (uint16_t)c.cur_first; my_count[c.cur_tile] = (uint16_t)c.cur_cnt; }
        c.cur_tile = tile; c.cur_first = c.n_segs; c.cur_cnt = 0;
      }
      c.cur_cnt++;
      c.n_segs++;
    }
    gran_advance(gsp, ggp, em, gran_row, t + (off - call_off), off, seg_len);
    cc.chunk_left -= seg_len; cc.hq_off += seg_len;
    if (gp.has_env && cc.env_per_frame) env_chain(v, gp, seg_len);
    off += seg_len; remaining -= seg_len;
  }
  *vp = v; *cp = cc; *cur = c;
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
  uint32_t tab_slots;            // warp-per-voice: shared-memory table slots (one per warp), 0 = tables are read from global memory
  // Autonomous voices (persistent mode): from block quiet_block[cta] on, the group has no event left that couples its
  // voices (note-on / stop / parameter / loop message, scheduled stop). Each voice's warp then walks the remaining blocks
  // at its own pace; a voice that finishes block b bumps auton_done[cta][b], and the LAST voice to do so carries out the
  // group-level bookkeeping of the block and hands it to the replay. Without this every block costs the group as much as
  // its slowest voice (one of eight is usually in a pitch glide), i.e. the sum of all voices' slow paths.
  const uint32_t* quiet_block;   // [ctas] first block of autonomy, >= n_blocks: never
  uint32_t* auton_done;          // [ctas][n_blocks]
  uint32_t* auton_cnt;           // [ctas][n_blocks][max_chunks] voices still holding a note after each chunk
};


// WPV (warp per voice): lane 0 of warp i owns voice i, so voices never serialise each other's divergent
// control flow (the skeleton is a latency-bound chain of dependent f32 ops per voice, not a SIMT workload).
// SIMPLE = false: the launch holds only granular / HighQuality voices, which never take a simple call -- the super-call
// machinery (and the registers it keeps live across the whole chunk loop) is compiled out.
template <int MAXT, bool WPV, bool AUTON, bool SIMPLE>
PB_DEV void skeleton_block(SkeletonArgs a, const TabSlot& tslot, const SkeletonLoop& L, const uint32_t b_first) {
  __shared__ VoiceHeader s_head[WPV ? 32 : MAXT];
  __shared__ GroupState s_gs;
  __shared__ uint32_t s_count;
  __shared__ uint64_t s_bounds[SB_MAX];   // this mixer's chunk boundaries of the block
  __shared__ uint32_t s_cnt[MAX_RUN];      // voices still holding a note after each chunk of a free run
  unsigned long long prof_free = 0;
  unsigned long long prof_q[4] = {0, 0, 0, 0};
  unsigned long long prof_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // simple calls, their cycles, jumped tiles, literal frames, general frames, their cycles, -, block cycles
  const long long prof_b0 = a.prof ? clock64() : 0;

  const long long cyc_block0 = CYC_T();
  const uint32_t g = a.group_list[blockIdx.x];
  const uint32_t tid = WPV ? ((threadIdx.x & 31u) == 0 ? (threadIdx.x >> 5) : 0xFFFFu) : threadIdx.x;
  uint32_t gp_idx = a.gstate[g].gp_idx;   // the parameter version in force (== g until a parameter event arrives)
  GroupParams gp = a.groups[gp_idx];
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

  uint32_t cb = a.mixer_chunk_begin[gp.mixer], ce = a.mixer_chunk_begin[gp.mixer + 1];
  for (uint32_t i = threadIdx.x; i < min(ce - cb, SB_MAX); i += blockDim.x) s_bounds[i] = a.chunk_bounds[cb + i];
  __syncthreads();
  // (autonomous voices are in different blocks at the same time: they read the schedule from global memory)
  auto bound = [&](const uint32_t i) -> uint64_t { return (!AUTON && i - cb < SB_MAX) ? s_bounds[i - cb] : a.chunk_bounds[i]; };
  // time of the next pending event (re-read only when the cursor moved)
  // next event that needs all voices at the same frame (see the free run below): its index and time
  uint32_t hard_idx = 0xFFFFFFFFu;
  uint64_t hard_time = UINT64_MAX;
  uint8_t* gflags = a.group_flags + (size_t)g * a.max_chunks;
  uint64_t my_frames = 0;

  // jump table of the voice's current ratio, looked up when the ratio changes
  uint32_t tab_bits = 0;
  PhaseRef tab_ref = phase_ref(nullptr, nullptr);
  auto tab_for = [&](const float ratio) -> const PhaseRef& {
    const uint32_t bits = (uint32_t)__float_as_int(ratio);
    if (bits != tab_bits) {
      tab_bits = bits;
      const uint32_t* t = (a.debug_flags & 4u) ? nullptr : find_phase_tab(a, ratio);
      const uint32_t* body = nullptr;
      if (WPV && t != nullptr && tslot.words != nullptr && (t[PT_H_MODE] == PT_DOWN_TABLE || t[PT_H_MODE] == PT_UP_TABLE)) {
        const uint32_t words = t[PT_H_WORDS] - PT_HEADER;
        if (words <= TAB_SLOT_WORDS) { tab_slot_fetch(tslot, t + PT_HEADER, words * 4u); body = tslot.words; }
      }
      tab_ref = phase_ref(t, body);
      if (body) tab_ref.smem_body = smem_u32(body);
    }
    return tab_ref;
  };
  SuperCall sc;   // the voice's open simple call, if any (see "simple calls and super-calls")
  sc.open = 0;
  // advance the pending frames of the open call literally and close it: the voice state is exact at sc.end_off again
  const uint32_t CCr = buf.channels;
  auto sc_finish = [&]() {
    if (!SIMPLE || !sc.open) return;
    sc_advance<!WPV>(sc, v, gp, a.groups + gp_idx, my_recs, a.gen, CCr, sc.end_off, true, a.prof ? prof_s + 2 : nullptr);
    sc_close(sc, v, CCr);
  };
  // One Source::write call of this thread's voice (already opened by voice_begin_call): emits the call's
  // Segment / TileRec checkpoints and advances the control state. Returns the frames written.
  auto run_call = [&](CallCtx& cc, const uint32_t n, const uint32_t call_off, const uint64_t t, const bool allow_lazy) __attribute__((always_inline)) -> uint32_t {
    uint32_t written_frames = 0;
    bool simple = false;
    const bool cyc_v = CYC_ON(gp.first_voice + tid);
    const long long cyr0 = cyc_v ? CYC_T() : 0ll;
    if (SIMPLE && n_segs < a.seg_cap && !(a.debug_flags & 2u) && !is_hq && !is_gran)
      simple = buf.channels == 2 ? simple_call_ok<2>(v, cc, buf, n) : simple_call_ok<1>(v, cc, buf, n);
    const long long prc0 = a.prof ? clock64() : 0;
    if (simple) {
      const uint32_t tile = call_off / TILE;
      if (!(a.debug_flags & 1u)) {
        Segment& s = my_segs[n_segs];
        s.v = v; s.c = cc; s.out_off = call_off; s.n = min(n, (tile + 1) * TILE - call_off); s.gp_idx = gp_idx;
      }
      if (tile != cur_tile) {
        if (cur_tile != 0xFFFFFFFFu) { my_first[cur_tile] = (uint16_t)cur_first; my_count[cur_tile] = (uint16_t)cur_cnt; }
        cur_tile = tile; cur_first = n_segs; cur_cnt = 0;
      }
      cur_cnt++;
      // open the call; inside a free run it may stay open as a super-call (whole tiles advanced, the rest pending)
      const PhaseRef& tr = tab_for(v.ratio);
      unsigned long long* st = a.prof ? prof_s + 2 : nullptr;
      const uint32_t end_off = call_off + n;
      sc_open(sc, v, cc, gp, buf, CCr, call_off, n_segs, tr);
      sc.end_off = end_off;
      bool lazy = false;
      if (allow_lazy) { sc.budget = sc_steady_budget(sc, v, gp, CCr); lazy = sc.budget >= (end_off + TILE - 1u) / TILE * TILE - call_off; }
      sc_advance<!WPV>(sc, v, gp, a.groups + gp_idx, my_recs, a.gen, CCr, end_off, !lazy, st);
      if (!lazy || (sc.adv_off % TILE) != 0u) sc_close(sc, v, CCr);
      cc.produced_in_call = n; cc.call_left = 0; cc.chunk_left = 0;
      if (!sc.open) after_process_call(v, cc);  // (an open super-call cannot have reached its loop end: sc_steady_budget)
      n_segs++;
      written_frames = n;
    } else if (is_gran) {
      VoiceState vt = v;
      CallCtx ct = cc;
      SegCursor cur{n_segs, cur_tile, cur_first, cur_cnt};
      gran_call_nl(&vt, &ct, a.groups + gp_idx, gsp, a.gran_groups + g, &a.gran, gran_row, t, call_off, n, my_segs, my_first, my_count,
                   a.seg_cap, gp_idx, !(a.debug_flags & 1u), &cur);
      v = vt; cc = ct;
      n_segs = cur.n_segs; cur_tile = cur.cur_tile; cur_first = cur.cur_first; cur_cnt = cur.cur_cnt;
      written_frames = n;
    } else {
      uint32_t off = call_off, remaining = n;
      while (remaining > 0 && !cc.ended) {
        const uint32_t tile = off / TILE;
        const uint32_t seg_len = min(remaining, (tile + 1) * TILE - off);
        if (n_segs < a.seg_cap && !(a.debug_flags & 1u)) {
          Segment& s = my_segs[n_segs];
          s.v = v; s.c = cc; s.out_off = off; s.n = seg_len; s.gp_idx = gp_idx;
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
        else {  // out of line, on copies (taking the address of `v` / `cc` themselves would move them to local memory)
          VoiceState vt = v;
          CallCtx ct = cc;
          w = buf.channels == 2 ? voice_advance_nl<2>(&vt, &ct, a.groups + gp_idx, a.buffers + gp.buffer, out_rate, comp, seg_len)
                                : voice_advance_nl<1>(&vt, &ct, a.groups + gp_idx, a.buffers + gp.buffer, out_rate, comp, seg_len);
          v = vt; cc = ct;
        }
        written_frames += w;
        off += w; remaining -= w;
        if (w < seg_len) break;
      }
    }
    if (a.prof) {
      const unsigned long long dt = (unsigned long long)(clock64() - prc0);
      if (simple) { prof_s[0] += 1; prof_s[1] += dt; } else { prof_s[4] += n; prof_s[5] += dt; }
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

  // One mixer chunk [r0, r0 + rlen) of this thread's voice inside a span that needs no group-wide agreement: the events
  // addressed to the voice, then its Source::write call (or the chunk's acceptance into the open super-call). `cnt`
  // counts the voices that still hold a note after the chunk.
  auto voice_chunk = [&](const uint64_t r0, const uint32_t rlen, uint32_t* cnt, uint32_t& ec, uint64_t& ev_t, const bool ignore) __attribute__((always_inline)) {
    const uint32_t r0_off = (uint32_t)(r0 - a.block_start);
    // events due at this chunk's start (MixedSource::process_events -> the first write call of the chunk)
    while (ev_t <= r0) {
      const DevEvent ev = a.events[ec];
      ++ec;
      ev_t = ec < gp.ev_end ? a.events[ec].time : UINT64_MAX;
      if (ignore || ev.kind == EVK_SET_VOLUME || ev.kind == EVK_SET_PANNING) continue;
      if (ev.kind != EVK_ALL_NOTES_OFF && !(v.has_note && v.note_id == ev.note_id)) continue;
      sc_finish();  // the event changes this voice: its state has to be exact at r0 first
      if (ev.kind == EVK_ALL_NOTES_OFF || ev.kind == EVK_NOTE_OFF) stop_voice(r0);
      else if (ev.kind == EVK_NOTE_SPEED) { file_set_speed(v, ev.speed, ev.glide, buf.sample_rate, out_rate); if (gsp) gsp->speed = ev.speed; }
      else if (ev.kind == EVK_NOTE_VOLUME) { v.note_volume = ev.value; exp_set_target(v.vol, gp.base_volume * ev.value, comp); if (gsp) gsp->volume = gp.base_volume * ev.value; }
      else if (ev.kind == EVK_NOTE_PANNING) {
        v.note_panning = ev.value;
        const float eff = fminf(fmaxf(gp.base_panning + ev.value, -1.0f), 1.0f);
        exp_set_target(v.pan, eff, comp);
        if (gsp) gsp->panning = eff;
      }
    }
    if (!v.has_note) return;
    if (SIMPLE && sc.open) {  // steady voice: accept the chunk into the open call when it provably stays steady to the tile's end
      const uint32_t r1_off = r0_off + rlen;
      const uint32_t need = (r1_off + TILE - 1u) / TILE * TILE - sc.adv_off;
      if (sc.budget < need) sc.budget = sc_steady_budget(sc, v, gp, CCr);
      if (sc.budget >= need) {
        sc.end_off = r1_off;
        unsigned long long* st = a.prof ? prof_s + 2 : nullptr;
        const long long q0 = a.prof ? clock64() : 0;
        sc_advance<!WPV>(sc, v, gp, a.groups + gp_idx, my_recs, a.gen, CCr, r1_off, false, st);
        if (a.prof) prof_s[7] += (unsigned long long)(clock64() - q0);
        my_frames += rlen;
        atomicAdd(cnt, 1u);
        return;
      }
      sc_finish();
    }
    CallCtx cc;
    cc.ended = true; cc.chunk_left = 0; cc.fader_running = false;
    if (voice_begin_call(v, cc, gp, buf, rlen, comp, gp.has_env != 0, r0_off)) run_call(cc, rlen, r0_off, r0, true);
    voice_epilogue(r0 + rlen, r0_off + rlen);
    if (v.has_note) atomicAdd(cnt, 1u);
  };

  if (AUTON) {
    // ---- autonomous voices (see SkeletonLoop): this warp's voice walks blocks b_first .. n_blocks-1 on its own --------
    const uint32_t cta = blockIdx.x;
    if (mine) {
      uint32_t ec = s_gs.ev_cursor;
      uint64_t ev_t = ec < gp.ev_end ? a.events[ec].time : UINT64_MAX;
      const uint32_t chunk_stride = L.chunk_begin_stride;
      for (uint32_t bb = b_first; bb < L.n_blocks; ++bb) {
        if (bb > b_first) {  // the tables of the next block (persistent mode: one snapshot slot per block)
          a.mixer_chunk_begin += chunk_stride;
          a.block_start += a.block_frames;
          a.segs += L.segs_stride; a.seg_first += L.seg_tab_stride; a.seg_count += L.seg_tab_stride;
          a.recs += L.recs_stride;
          a.gen += 1;
          my_segs = a.segs + (size_t)vidx * a.seg_cap;
          my_first = a.seg_first + (size_t)vidx * a.n_tiles;
          my_count = a.seg_count + (size_t)vidx * a.n_tiles;
          my_recs = a.recs + (size_t)vidx * a.n_tiles;
          n_segs = 0; cur_tile = 0xFFFFFFFFu; cur_first = 0; cur_cnt = 0;
          uint4* z = reinterpret_cast<uint4*>(my_count);   // n_tiles is a multiple of 16: 16-byte aligned rows
          for (uint32_t i = 0; i < a.n_tiles / 8u; ++i) z[i] = make_uint4(0u, 0u, 0u, 0u);
          cb = a.mixer_chunk_begin[gp.mixer]; ce = a.mixer_chunk_begin[gp.mixer + 1];
        }
        uint32_t* cnt = L.auton_cnt + ((size_t)cta * L.n_blocks + bb) * a.max_chunks;
        my_frames = 0;
        for (uint32_t k = cb; k + 1 < ce; ++k) {
          const uint64_t r0 = bound(k);
          voice_chunk(r0, (uint32_t)(bound(k + 1) - r0), cnt + (k - cb), ec, ev_t, false);
        }
        sc_finish();
        if (cur_tile != 0xFFFFFFFFu) { my_first[cur_tile] = (uint16_t)cur_first; my_count[cur_tile] = (uint16_t)cur_cnt; }
        if (n_segs >= a.seg_cap) atomicOr(a.overflow, 1u);
        if (my_frames) atomicAdd((unsigned long long*)&s_gs.voice_frames, (unsigned long long)my_frames);
        __threadfence();
        const uint32_t arrived = atomicAdd(L.auton_done + (size_t)cta * L.n_blocks + bb, 1u) + 1u;
        if (arrived == nv) {
          // ---- last voice out: the generator-level bookkeeping of block bb (what thread 0 does after a free run) ----
          __threadfence();
          const uint32_t* bbegin = a.mixer_chunk_begin;
          const uint32_t kb = bbegin[gp.mixer], ke = bbegin[gp.mixer + 1];
          const size_t boff_g = (size_t)(bb - b_first);
          g_segs = a.gsegs + boff_g * L.gsegs_stride + (size_t)g * a.seg_cap;
          g_first = a.gseg_first + boff_g * L.gseg_tab_stride + (size_t)g * a.n_tiles;
          g_count = a.gseg_count + boff_g * L.gseg_tab_stride + (size_t)g * a.n_tiles;
          gflags = a.group_flags + boff_g * L.group_flags_stride + (size_t)g * a.max_chunks;
          n_gsegs = 0; gcur_tile = 0xFFFFFFFFu; gcur_first = 0; gcur_cnt = 0;
          if (bb > b_first) for (uint32_t i = 0; i < a.n_tiles; ++i) g_count[i] = 0;
          uint32_t gec = s_gs.ev_cursor;
          for (uint32_t k = kb; k + 1 < ke; ++k) {
            const uint64_t e0 = a.chunk_bounds[k];
            while (gec < gp.ev_end && a.events[gec].time <= e0) {  // generator-level AmplifiedSource / PannedSource messages
              const DevEvent ev = a.events[gec];
              ++gec;
              if (ev.kind == EVK_SET_VOLUME) exp_set_target(s_gs.vol, ev.value, comp);
              else if (ev.kind == EVK_SET_PANNING) exp_set_target(s_gs.pan, ev.value, comp);
            }
            bool writes = false;
            if (!s_gs.dead) {
              writes = !(s_gs.stopped || (s_gs.active_voices == 0 && !s_gs.stopping));
              if (writes) {
                group_call((uint32_t)(a.chunk_bounds[k + 1] - e0), (uint32_t)(e0 - a.block_start));
                s_gs.active_voices = cnt[k - kb];
                if (s_gs.stopping && cnt[k - kb] == 0) s_gs.stopped = 1;
              }
              if (gp.transient && s_gs.stopped && !s_gs.dead) { s_gs.dead = 1; s_gs.dead_time = a.chunk_bounds[k + 1]; }
            }
            gflags[k - kb] = writes ? 1 : 0;
          }
          s_gs.ev_cursor = gec;
          if (gcur_tile != 0xFFFFFFFFu) { g_first[gcur_tile] = (uint16_t)gcur_first; g_count[gcur_tile] = (uint16_t)gcur_cnt; }
          if (bb + 1 == L.n_blocks) a.gstate[g] = s_gs;
          __threadfence();
          atomicAdd(L.block_done + bb, 1u);
        }
      }
      a.voices[gp.first_voice + tid] = v;
    }
    return;
  }
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
        while (hard_idx < gp.ev_end && a.events[hard_idx].kind != EVK_NOTE_ON && a.events[hard_idx].kind != EVK_STOP &&
               a.events[hard_idx].kind != EVK_SET_PARAM && a.events[hard_idx].kind != EVK_SET_LOOP) ++hard_idx;
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
          uint64_t ev_t = ec < gp.ev_end ? a.events[ec].time : UINT64_MAX;   // time of the group's next pending event
          const bool ignore = s_gs.stopping != 0;
          for (uint32_t j = 0; j < run; ++j) {
            const uint64_t r0 = bound(k + j);
            voice_chunk(r0, (uint32_t)(bound(k + j + 1) - r0), &s_cnt[j], ec, ev_t, ignore);
          }
          sc_finish();
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
      const bool file_was_finished = !is_sampler && mine && v.finished;   // (before this call's messages)
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
            } else if (ev.kind == EVK_SET_PARAM) {
              // Sampler::process_parameter_update: the version the host resolved becomes current; active voices re-derive
              // what depends on it (voice.rs:258-310)
              gp_idx = ev.seek_pos;
              gp = a.groups[gp_idx];
              if (tid == 0) s_gs.gp_idx = gp_idx;
              if (mine && v.has_note) {
                if (ev.note == PARAM_PITCH) {
                  const double eff = c_note_speed[v.note & 127u] * ev.speed;
                  file_set_speed(v, eff, 0.0f, buf.sample_rate, out_rate);
                  if (gsp) gsp->speed = eff;
                } else if (ev.note == PARAM_VOLUME) {
                  exp_set_target(v.vol, gp.base_volume * v.note_volume, comp);
                  if (gsp) gsp->volume = gp.base_volume * v.note_volume;
                } else if (ev.note == PARAM_PANNING) {
                  const float eff = fminf(fmaxf(gp.base_panning + v.note_panning, -1.0f), 1.0f);
                  exp_set_target(v.pan, eff, comp);
                  if (gsp) gsp->panning = eff;
                }
              }
              __syncthreads();
            } else if (ev.kind == EVK_SET_LOOP) {
              if (mine) {  // SamplerVoice::set_loop_range on every voice (voice.rs:313-339)
                const bool has = !(ev.flags & 2u);
                v.loop_ovr_start = has ? (int32_t)ev.seek_pos : -1; v.loop_ovr_end = has ? (int32_t)ev.note : -1;
                v.repeat = has ? REPEAT_FOREVER : 0u; v.repeat_count = v.repeat;
              }
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
      if (call_open) written_frames = run_call(cc, n, call_off, t, false);
      if (!is_sampler && mine && a.status) {  // PlaybackStatusEvent (file/common.rs:171-221, preloaded.rs:196-209,454-472)
        auto emit = [&](const uint32_t kind, const uint64_t frame) {
          const uint32_t i = atomicAdd(a.status_count, 1u);
          if (i < a.status_cap) { StatusRec sr; sr.frame = frame; sr.pos = v.playback_pos; sr.group = g; sr.kind = kind; a.status[i] = sr; }
        };
        if (call_open && t - min(t, v.pos_clock) >= a.pos_emit_rate) { v.pos_clock = t; emit(0u, t); }
        if (v.finished && !file_was_finished) emit(v.stopped_exhausted ? 1u : 2u, call_open ? t + n : t);
      }
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
  if ((mine && n_segs >= a.seg_cap) || (tid == 0 && n_gsegs >= a.seg_cap)) atomicOr(a.overflow, 1u);   // a full list may have dropped records
  if (mine) a.voices[gp.first_voice + tid] = v;
#ifdef PB200_CYC
  if (CYC_ON(gp.first_voice + tid) && mine) g_cyc[7] += CYC_T() - cyc_block0;
#endif
  if (a.prof && mine) {
    unsigned long long* pr = a.prof + (size_t)(gp.first_voice + tid) * 12;
    pr[0] += prof_q[0]; pr[1] += prof_q[1]; pr[2] += prof_q[3]; pr[3] += prof_free;
    for (int i = 0; i < 8; ++i) pr[4 + i] += prof_s[i];
  }
  if (my_frames) atomicAdd((unsigned long long*)&s_gs.voice_frames, (unsigned long long)my_frames);
  __syncthreads();
  if (tid == 0) a.gstate[g] = s_gs;
}

#ifndef SKEL_LPV_MINB
#define SKEL_LPV_MINB 1
#endif
template <int MAXT, bool WPV, bool SIMPLE = true>
__global__ void __launch_bounds__(MAXT, (WPV || MAXT > 256) ? 1 : SKEL_LPV_MINB) skeleton_kernel(SkeletonArgs a, SkeletonLoop L) {
  extern __shared__ __align__(128) uint32_t tab_smem[];   // WPV: [tab_slots][TAB_SLOT_WORDS] | mbarriers | parities
  TabSlot tslot;
  tslot.words = nullptr; tslot.mbar = nullptr; tslot.parity = nullptr;
  if (WPV && L.tab_slots) {
    const uint32_t w = threadIdx.x >> 5;
    if (w < L.tab_slots) {
      tslot.words = tab_smem + (size_t)w * TAB_SLOT_WORDS;
      tslot.mbar = reinterpret_cast<uint64_t*>(tab_smem + (size_t)L.tab_slots * TAB_SLOT_WORDS) + w;
      tslot.parity = tab_smem + (size_t)L.tab_slots * (TAB_SLOT_WORDS + 2u) + w;
      if ((threadIdx.x & 31u) == 0) tab_slot_init(tslot);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
  }
  for (uint32_t b = 0; b < L.n_blocks; ++b) {
    if constexpr (WPV && SIMPLE) {
      if (L.quiet_block != nullptr && b >= L.quiet_block[blockIdx.x]) {  // the group's voices go their own way from here on
        skeleton_block<MAXT, WPV, true, SIMPLE>(a, tslot, L, b);
        return;
      }
    }
    skeleton_block<MAXT, WPV, false, SIMPLE>(a, tslot, L, b);
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
