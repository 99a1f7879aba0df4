// Mixer kernels, launched per tree level (deepest first) for every time block.
//
// mix_sum_kernel   (M1): bus[f] = sum of gated sub-mixer buses then source (group) buses, in the
//                  reference's order (MixedSource::process_sub_mixers / process_sources,
//                  src/source/mixed.rs:505-624). Fully parallel over frames.
// mix_fx_kernel    (M2): one CTA per mixer walks the mixer's own chunk schedule: effect parameter
//                  events, the effect chain with the auto-bypass state machine
//                  (EffectProcessor::process, mixed/effect.rs:56-145) on a chunk staged in shared
//                  memory, the sub-mixer silence gate at every *parent* chunk end
//                  (SubMixerProcessor::process, mixed/submixer.rs:47-77) with an exact max|x| reduction,
//                  and for the main mixer the WavStream master volume per 1024-frame block
//                  (src/output/wav.rs:237) plus the final output store.
#pragma once
#include "effects_par.cuh"

namespace pb {

constexpr uint32_t PB_MAX_EXT = 8;   // external buses of a main mixer (PB200_MAX_MAIN_INPUTS)

struct MixerKernelArgs {
  const MixerParams* mixers;
  MixerState* mstate;
  const uint32_t* level_mixers;   // dense mixer indices of this level
  const uint32_t* child_index;    // children per mixer (reference order)
  const uint32_t* source_index;   // groups per mixer (playing_sources order)
  FxHeader* fx;
  const FxParamEvent* fx_events;
  FxCtx fxc;
  const uint64_t* chunk_bounds;
  const uint32_t* mixer_chunk_begin;
  const float* group_bus;         // [n_groups][block_frames][2]
  const uint8_t* group_flags;     // [n_groups][max_chunks]
  float* mixer_bus;               // [n_mixers][block_frames][2]
  uint8_t* mixer_flags;           // [n_mixers][max_chunks]: audible per *parent* chunk
  uint32_t max_chunks;
  uint32_t block_frames;
  uint32_t block_len;             // frames actually rendered in this block
  uint64_t block_start;
  // main mixer output
  float* out;                     // device output for this block (interleaved stereo) or nullptr
  ExpSm* master;                  // WavStream::smoothed_volume
  uint32_t wav_block_frames;      // 1024
  uint32_t work_bytes;            // dynamic shared memory of this launch (FX_WORK_SMALL unless a mixer of the level holds a reverb)
  // Effect-chain pipelining (n_stages > 1): CTA (mixer, stage) runs the effects [stage_begin[stage], stage_begin[stage+1]) of
  // its mixer; chunk q moves from stage to stage through the mixer bus in global memory, announced by fx_progress.
  uint32_t n_stages;              // gridDim.y of the launch; 1 = one CTA per mixer runs the whole chain
  const uint32_t* stage_begin;    // [n_mixers][MAX_FX_STAGES + 1] effect indices relative to fx_begin (by dense mixer index)
  uint32_t* fx_progress;          // [n_mixers][n_stages] chunks of this block finished by the stage (zeroed per block)
  uint8_t* fx_pflags;             // [n_mixers][n_stages][max_chunks] bit0 input still bypassed, bit1 some effect ran
  uint32_t* fx_ticket;            // CTA start counter of the launch (zeroed per launch)
  const float* ext_in[PB_MAX_EXT]; // main mixer: external stereo buses of this block added to its input, in order (pb200_set_main_inputs)
  uint32_t n_ext;
  uint32_t ext_len;               // frames of them inside this block
  double* meter;                  // main mixer: [wav block of the render][peak L, peak R, sum of squares L, R] or nullptr (MeteredSource)
  uint64_t render_start;          // first frame of the render call (meter rows count from it)
  uint32_t direct_out;            // this level's (single) mixer writes the render output itself: its parent, the main mixer, has nothing else to do
  unsigned long long* progress;   // mapped host word: the main mixer's CTA stores progress_value when the block's output is final
  unsigned long long progress_value;
  unsigned long long* prof;
  uint32_t prof_all;       // PB200_FX_PROF: [8] cycle counters of the main mixer's CTA (debug aid)
};

PB_DEV uint64_t sat_sub_u64(uint64_t a, uint64_t b) { return a > b ? a - b : 0; }
PB_DEV uint64_t sat_add_u64(uint64_t a, uint64_t b) { return (a > UINT64_MAX - b) ? UINT64_MAX : a + b; }

// ---- M1 -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mix_sum_kernel(MixerKernelArgs a) {
  const uint32_t m = a.level_mixers[blockIdx.y];
  const MixerParams mp = a.mixers[m];
  const uint32_t cb = a.mixer_chunk_begin[m], ce = a.mixer_chunk_begin[m + 1];
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;  // frame within the block
  if (f >= a.block_len) return;
  // chunk index of this frame: last boundary <= frame
  const uint64_t t = a.block_start + f;
  uint32_t lo = cb, hi = ce - 1;  // bounds[lo] <= t < bounds[hi]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a.chunk_bounds[mid] <= t) lo = mid; else hi = mid;
  }
  const uint32_t k = lo - cb;
  float2 s = make_float2(0.0f, 0.0f);
  if (a.n_ext && mp.parent == 0xFFFFFFFFu && f < a.ext_len)   // the sub-mixers rendered elsewhere, summed in the order given
    for (uint32_t e = 0; e < a.n_ext; ++e) { const float2 v = __ldcg(reinterpret_cast<const float2*>(a.ext_in[e]) + f); s.x += v.x; s.y += v.y; }
  for (uint32_t ci = mp.child_begin; ci < mp.child_end; ++ci) {
    const uint32_t c = a.child_index[ci];
    if (a.mixer_flags[(size_t)c * a.max_chunks + k]) {
      const float2 v = *reinterpret_cast<const float2*>(a.mixer_bus + ((size_t)c * a.block_frames + f) * 2);
      s.x += v.x; s.y += v.y;
    }
  }
  for (uint32_t si = mp.src_begin; si < mp.src_end; ++si) {
    const uint32_t g = a.source_index[si];
    if (a.group_flags[(size_t)g * a.max_chunks + k]) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(a.group_bus + ((size_t)g * a.block_frames + f) * 2));
      s.x += v.x; s.y += v.y;
    }
  }
  *reinterpret_cast<float2*>(a.mixer_bus + ((size_t)m * a.block_frames + f) * 2) = s;
}

// ---- M2 -------------------------------------------------------------------------------------------------
constexpr uint32_t FX_THREADS = FX_THREADS_C;
// frames after a parameter event during which a smoothed effect parameter may still be ramping (the default exponential
// smoother needs 278 x ln(delta / 3.3e-3) frames at 48 kHz: 4400 for a 20 kHz cutoff jump; 65536 covers every built-in smoother)
constexpr uint64_t FX_RAMP_WINDOW = 65536;
constexpr uint32_t MAX_FX_STAGES = 4;    // pipeline stages of a mixer's effect chain (one CTA each)
constexpr uint32_t FX_WORK_SMALL = 48 * 1024;   // every effect but the whole-chunk reverb fits: keeps the L1 carve-out and 3 CTAs per SM
constexpr uint32_t FX_WORK_BYTES = 120 * 1024;  // shared-memory work area of the chunk-parallel effects (the reverb's ten f64 planes of a whole chunk)

// MINB = 2: the 128-register build, two CTAs per SM, for levels whose pipeline needs more co-resident CTAs than SMs
// PLAIN: the build for mixers WITHOUT effects (gate / master volume / output only): a few dozen registers, so its CTAs find
// room next to the resident skeleton and replay CTAs instead of waiting for a whole free SM like the effect build does.
template <int MINB, bool PLAIN = false>
__global__ void __launch_bounds__(FX_THREADS, MINB) mix_fx_kernel(MixerKernelArgs a) {
  __shared__ float s_ch[2][PLANE];
  __shared__ double s_scratch[2][PLANE];
  __shared__ double s_lane_state[2][64];
  __shared__ float s_red[FX_THREADS / 32];
  __shared__ uint32_t s_run;      // effect e runs this chunk
  extern __shared__ __align__(16) uint8_t s_work[];
  const ParWork pw{s_work, a.work_bytes};

  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  const uint32_t n_stages = a.n_stages;
  // Pipelined launch: a CTA's place (mixer, stage) comes from a ticket taken when it starts running, stage-major, so every
  // CTA it will wait for holds a lower ticket and is therefore already running (or done) -- no co-residency requirement,
  // whatever order the hardware dispatches the grid in.
  uint32_t slot_x = blockIdx.x, stage = 0;
  if (n_stages > 1) {
    if (tid == 0) s_run = atomicAdd(a.fx_ticket, 1u);
    __syncthreads();
    const uint32_t ticket = s_run;
    __syncthreads();
    slot_x = ticket % gridDim.x; stage = ticket / gridDim.x;
  }
  const uint32_t m = a.level_mixers[slot_x];
  const uint32_t lane = tid & 31, warp = tid >> 5;
  const MixerParams mp = a.mixers[m];
  const bool is_main = mp.parent == 0xFFFFFFFFu;
  const bool has_fx = !PLAIN && mp.fx_end > mp.fx_begin;
  // the effects this CTA runs, and its place in the mixer's pipeline
  uint32_t e_lo = mp.fx_begin, e_hi = mp.fx_end;
  if (n_stages > 1) {
    const uint32_t* sb = a.stage_begin + (size_t)m * (MAX_FX_STAGES + 1);
    e_lo = mp.fx_begin + sb[stage]; e_hi = mp.fx_begin + sb[stage + 1];
    if (stage > 0 && e_lo >= e_hi) return;   // (a mixer with fewer effects than stages; stage 0 always runs: gate / master)
  }
  const bool first_stage = e_lo == mp.fx_begin, last_stage = e_hi == mp.fx_end;
  uint32_t prev_stage = stage;               // the stage whose output this one consumes
  if (!first_stage) { const uint32_t* sb = a.stage_begin + (size_t)m * (MAX_FX_STAGES + 1); do { --prev_stage; } while (sb[prev_stage] == sb[prev_stage + 1]); }
  volatile uint32_t* prog_in = a.fx_progress ? a.fx_progress + (size_t)m * n_stages + prev_stage : nullptr;
  uint32_t* prog_out = a.fx_progress ? a.fx_progress + (size_t)m * n_stages + stage : nullptr;
  const uint8_t* pf_in = a.fx_pflags ? a.fx_pflags + ((size_t)m * n_stages + prev_stage) * a.max_chunks : nullptr;
  uint8_t* pf_out = a.fx_pflags ? a.fx_pflags + ((size_t)m * n_stages + stage) * a.max_chunks : nullptr;
  uint32_t q = 0;                            // sequence number of the (merged) chunk: the same in every stage
  float* bus = a.mixer_bus + (size_t)m * a.block_frames * 2;
  const uint32_t cb = a.mixer_chunk_begin[m], ce = a.mixer_chunk_begin[m + 1];
  uint32_t pcb = 0, pk = 0;
  if (!is_main) { pcb = a.mixer_chunk_begin[mp.parent]; pk = pcb; }
  const uint32_t sr = a.fxc.sample_rate;
  ChunkBuf cbuf;
  cbuf.ch[0] = s_ch[0]; cbuf.ch[1] = s_ch[1];
  cbuf.scratch = &s_scratch[0][0];
  cbuf.lane_state = &s_lane_state[0][0];

  // chunk <-> shared memory: interleaved global (L2: another SM may have written it) <-> planar padded shared; two frames per
  // 16-byte access when the chunk starts on a 16-byte boundary, all of a thread's loads issued before its first store
  auto stage_in = [&](const float* __restrict__ g, const uint32_t len) {
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
      const float4* g4 = reinterpret_cast<const float4*>(g);
      const uint32_t n4 = len >> 1;
      for (uint32_t i0 = tid; i0 < n4; i0 += 4 * nt) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const uint32_t i = i0 + j * nt; if (i < n4) v[j] = __ldcg(g4 + i); }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t i = i0 + j * nt;
          if (i < n4) { const uint32_t f = 2 * i; s_ch[0][pidx(f)] = v[j].x; s_ch[1][pidx(f)] = v[j].y; s_ch[0][pidx(f + 1)] = v[j].z; s_ch[1][pidx(f + 1)] = v[j].w; }
        }
      }
      if ((len & 1u) && tid < 2) s_ch[tid][pidx(len - 1)] = __ldcg(g + 2 * (len - 1) + tid);
    } else {
      for (uint32_t i = tid; i < len * 2; i += nt) s_ch[i & 1][pidx(i >> 1)] = __ldcg(g + i);
    }
  };
  // returns this thread's max|x| of the chunk's samples (before `gain`)
  auto stage_out = [&](float* __restrict__ g, const uint32_t len, const float gain, const bool scale) -> float {
    float mx = 0.0f;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
      float4* g4 = reinterpret_cast<float4*>(g);
      const uint32_t n4 = len >> 1;
      for (uint32_t i = tid; i < n4; i += nt) {
        const uint32_t f = 2 * i;
        float4 v = make_float4(s_ch[0][pidx(f)], s_ch[1][pidx(f)], s_ch[0][pidx(f + 1)], s_ch[1][pidx(f + 1)]);
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        if (scale) { v.x *= gain; v.y *= gain; v.z *= gain; v.w *= gain; }
        g4[i] = v;
      }
      if ((len & 1u) && tid < 2) { float x = s_ch[tid][pidx(len - 1)]; mx = fmaxf(mx, fabsf(x)); if (scale) x *= gain; g[2 * (len - 1) + tid] = x; }
    } else {
      for (uint32_t i = tid; i < len * 2; i += nt) { float x = s_ch[i & 1][pidx(i >> 1)]; mx = fmaxf(mx, fabsf(x)); if (scale) x *= gain; g[i] = x; }
    }
    return mx;
  };
  float gate_mx = 0.0f;          // this thread's share of max|x| over the written-back chunks of the open parent chunk
  uint32_t gate_covered = 0;     // frames of the open parent chunk those chunks cover (uniform)
  long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto tick = [&](int i, long long& t0) { if (a.prof) { const long long t1 = clock64(); pt[i] += t1 - t0; t0 = t1; } };
  // pb200_render_progress: the main mixer's (last-stage) CTA writes the block's output last; once every thread's stores are
  // out, one system-scope store tells the host that the block is final
  auto publish_progress = [&]() {
    if (!a.progress) return;
    __threadfence_system();
    __syncthreads();
    if (tid == 0) { __threadfence_system(); *reinterpret_cast<volatile unsigned long long*>(a.progress) = a.progress_value; }
  };
  // audible input per chunk (mixed.rs:701-708): one flag per child / source and chunk, gathered for the whole block at once
  // (the children's and sources' kernels of this block have finished)
  constexpr uint32_t AUD_MAX = 1024;
  __shared__ uint8_t s_aud[AUD_MAX];
  auto audible_of = [&](const uint32_t kk) -> bool {
    bool aud = is_main && a.n_ext != 0;   // (an external bus counts as an audible sub-mixer)
    // (eight flags per round trip: the loads of a batch are independent, the early exit is taken between batches)
    for (uint32_t ci = mp.child_begin; ci < mp.child_end && !aud; ci += 8) {
      uint32_t any = 0;
#pragma unroll
      for (uint32_t j = 0; j < 8; ++j) if (ci + j < mp.child_end) any |= a.mixer_flags[(size_t)a.child_index[ci + j] * a.max_chunks + (kk - cb)];
      aud = any != 0;
    }
    for (uint32_t si = mp.src_begin; si < mp.src_end && !aud; si += 8) {
      uint32_t any = 0;
#pragma unroll
      for (uint32_t j = 0; j < 8; ++j) if (si + j < mp.src_end) any |= a.group_flags[(size_t)a.source_index[si + j] * a.max_chunks + (kk - cb)];
      aud = any != 0;
    }
    return aud;
  };
  // this mixer's chunk boundaries of the block in shared memory (a dependent L2 load per use otherwise: ~700 cycles each,
  // several per chunk)
  __shared__ uint64_t s_bounds[AUD_MAX + 1];
  const bool bounds_cached = ce - cb <= AUD_MAX + 1;
  if (bounds_cached) {
    for (uint32_t kk = cb + tid; kk < ce; kk += nt) s_bounds[kk - cb] = a.chunk_bounds[kk];
    __syncthreads();
  }
  auto bound_at = [&](const uint32_t kk) -> uint64_t { return bounds_cached ? s_bounds[kk - cb] : a.chunk_bounds[kk]; };
  uint64_t parent_next = is_main ? 0ull : a.chunk_bounds[pk + 1];   // end of the open parent chunk (kept in a register)
  uint32_t eff_byp = a.mstate[m].effects_bypassed;                  // MixedSource::effects_bypassed: this CTA is its only writer
  // WavStream's master-volume smoother: a shared copy (one L2 round trip per chunk otherwise); thread 0 is its only writer
  __shared__ ExpSm s_master;
  const bool direct = !is_main && a.direct_out != 0;   // (host: the main mixer has this one child, no effects, no sources, static master volume)
  if ((is_main || direct) && tid == 0) s_master = *a.master;
  __syncthreads();
  const uint32_t wbf = a.wav_block_frames;   // block_start is a multiple of it: boundaries are tested on 32-bit offsets
  bool any_fx_events = false;   // (static: event lists do not change during a launch)
  if (has_fx) {
    for (uint32_t e = mp.fx_begin; e < mp.fx_end; ++e) any_fx_events |= a.fx[e].ev_end > a.fx[e].ev_begin;
    for (uint32_t kk = cb + tid; kk + 1 < ce && kk - cb < AUD_MAX; kk += nt) s_aud[kk - cb] = audible_of(kk) ? 1 : 0;
    __syncthreads();
  } else if (is_main && a.out && !a.meter) {
    // a main mixer without effects (the top of a tree whose work sits in the sub-mixers): while the master volume is not
    // ramping the whole time block is one scaled copy, whatever the chunk boundaries are
    const ExpSm ms0 = s_master;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(bus) | reinterpret_cast<uintptr_t>(a.out)) & 15u) == 0 && (a.block_len & 1u) == 0;
    if (vec_ok && !exp_need_ramp(ms0, a.fxc.comp)) {   // (odd block_frames: the chunk loop below copies frame by frame)
      const float g = ms0.target;
      const bool scale = fabsf(1.0f - g) > 0.000001f;
      const float4* src = reinterpret_cast<const float4*>(bus);
      float4* dst = reinterpret_cast<float4*>(a.out);
      const uint32_t n4 = a.block_len / 2;
      for (uint32_t i = tid; i < n4; i += nt) {
        float4 v = src[i];
        if (scale) { v.x *= g; v.y *= g; v.z *= g; v.w *= g; }
        dst[i] = v;
      }
      publish_progress();
      return;
    }
  }
  for (uint32_t k = cb; k + 1 < ce; ++k) {
    long long t0 = a.prof ? clock64() : 0;
    const uint64_t c0 = bound_at(k);
    uint64_t c1 = bound_at(k + 1);
    uint32_t len = (uint32_t)(c1 - c0);
    const uint32_t boff = (uint32_t)(c0 - a.block_start);
    float* gchunk = bus + (size_t)boff * 2;
    // WavStream's master volume (wav.rs:237) is decided per 1024-frame block; while it is not ramping its state does not
    // move, so the decision holds for every chunk of the block
    bool master_direct = false, master_scale = false, chunk_done = false;
    float master_gain = 1.0f;
    if ((is_main || direct) && a.out && !a.meter) {
      const ExpSm ms0 = s_master;
      if (!exp_need_ramp(ms0, a.fxc.comp)) { master_direct = true; master_gain = ms0.target; master_scale = fabsf(1.0f - master_gain) > 0.000001f; }
    }

    if (has_fx) {
      // effect parameter events due at this chunk start (MixedSource::process_events, mixed.rs:683)
      for (uint32_t e = e_lo; e < e_hi && any_fx_events; ++e) {
        FxHeader& h = a.fx[e];
        // most chunks have no event due. The vote is also a barrier: every thread has read the cursor before thread 0 moves it
        if (!__syncthreads_or(h.ev_cursor < h.ev_end && a.fx_events[h.ev_cursor].time <= c0)) continue;
        for (;;) {  // events in order; a message (Effect::process_message) is carried out by the whole CTA
          if (tid == 0) {
            s_run = 0u;
            while (h.ev_cursor < h.ev_end && a.fx_events[h.ev_cursor].time <= c0) {
              const FxParamEvent pe = a.fx_events[h.ev_cursor];
              h.ev_cursor++;
              if (pe.normalized == 2u) { s_run = pe.param_id; break; }
              fx_apply_param(h, a.fxc, pe);
            }
          }
          __syncthreads();
          const uint32_t msg = s_run;
          __syncthreads();  // (everybody has read the verdict before thread 0 reuses the flag)
          if (msg == 0u) break;
          fx_process_message(h, a.fxc, msg, tid, nt);
          __syncthreads();
        }
      }
      const bool audible = (k - cb < AUD_MAX) ? s_aud[k - cb] != 0 : audible_of(k);
      // While the input is audible every effect runs and its bypass counters are reset chunk after chunk
      // (mixed/effect.rs:71-79): processing consecutive audible chunks in one go gives the same samples. Merge them up
      // to the next 1024-frame block boundary / parent chunk boundary / effect event -- most chunk boundaries of a busy
      // mixer come from its SOURCES' events (mixed.rs:686-693) and mean nothing to the effects.
      if (audible) {
        // An event of ANY effect of the mixer in (c0, nb] ends the merge, and so does a parameter event in the recent past:
        // a smoothed parameter decides per process() call -- per chunk -- whether it still ramps (chorus.rs / filter.rs /
        // eq5.rs ...: `if need_ramp { per-sample } else { constant }`), so while a ramp set off by an event may still be
        // running the chunks are the reference's. Both are judged on the event times alone (the cursors and the smoothers
        // belong to the effects' own stages), so that every stage of a pipeline cuts the same chunks.
        uint64_t next_ev = UINT64_MAX;
        bool can_merge = true;
        for (uint32_t e = mp.fx_begin; e < mp.fx_end && any_fx_events; ++e) {
          const FxHeader& h = a.fx[e];
          uint32_t lo2 = h.ev_begin, hi2 = h.ev_end;   // first event with time > c0
          while (lo2 < hi2) { const uint32_t mid = (lo2 + hi2) >> 1; if (a.fx_events[mid].time <= c0) lo2 = mid + 1; else hi2 = mid; }
          if (lo2 < h.ev_end) next_ev = min(next_ev, (uint64_t)a.fx_events[lo2].time);
          if (lo2 > h.ev_begin && c0 - a.fx_events[lo2 - 1].time < FX_RAMP_WINDOW) can_merge = false;
        }
        while (can_merge && k + 2 < ce) {
          const uint64_t nb = bound_at(k + 1);   // start of the next chunk
          if ((uint32_t)(nb - a.block_start) % wbf == 0) break;
          if (!is_main && nb == parent_next) break;
          if (!((k + 1 - cb < AUD_MAX) ? s_aud[k + 1 - cb] != 0 : audible_of(k + 1))) break;
          if (next_ev <= nb) break;
          ++k;
          c1 = bound_at(k + 1);
        }
        len = (uint32_t)(c1 - c0);
      }
      bool input_bypassed = !audible;
      bool all_bypassed = true;
      if (!first_stage) {  // chunk q leaves the previous stage: its samples are in the bus, its flags in pf_in[q]
        tick(0, t0);
        if (tid == 0) { while (*prog_in < q + 1u) __nanosleep(64); }
        __syncthreads();
        tick(7, t0);
        __threadfence();
        const uint8_t f = __ldcg(pf_in + q);
        input_bypassed = (f & 1u) != 0; all_bypassed = (f & 2u) == 0;
      }
      // (a pipeline stage cannot see the chain's verdict of the previous chunk; skipping is only a shortcut: with every
      // effect bypassed and the input silent each EffectProcessor stays bypassed by itself, mixed/effect.rs:88-91)
      bool skip_all = n_stages == 1 && eff_byp && input_bypassed;  // mixed.rs:629
      if (n_stages > 1 && input_bypassed) {  // nothing to stage when every effect of this stage stays bypassed
        bool any = false;
        for (uint32_t e = e_lo; e < e_hi; ++e) any |= !(a.fx[e].tail_counter == 0 && a.fx[e].silence_counter == UINT64_MAX);
        skip_all = !any;
      }
      tick(0, t0);
      if (!skip_all) {
        // stage the chunk: interleaved global -> planar padded shared (L2 loads: another SM may have written it)
        stage_in(gchunk, len);
        __syncthreads();
        tick(1, t0);
        for (uint32_t e = e_lo; e < e_hi; ++e) {
          FxHeader& h = a.fx[e];
          // EffectProcessor::process bypass transitions (mixed/effect.rs:64-109), decided by thread 0
          if (tid == 0) {
            bool bypassed = h.bypassed != 0;
            const bool should_bypass = input_bypassed && h.tail_counter == 0 && h.silence_counter == UINT64_MAX;
            if (should_bypass && !bypassed) bypassed = true;
            else if (!should_bypass && bypassed) { bypassed = false; h.tail_counter = UINT64_MAX; h.silence_counter = 0; }
            h.bypassed = bypassed;
            s_run = bypassed ? 0u : 1u;
          }
          __syncthreads();
          tick(2, t0);
          const bool run = s_run != 0;
          if (run) {
            fx_process_chunk(h, a.fxc, cbuf, len, tid, nt, pw);
            __syncthreads();
            tick(3, t0);
            if (input_bypassed) {
              uint64_t tail_frames;
              const bool has_tail = fx_process_tail(h, a.fxc, tail_frames);
              float mx = 0.0f;
              if (!has_tail) {  // unknown tail: exact max|x| of the processed chunk
                for (uint32_t i = tid; i < len; i += nt) mx = fmaxf(mx, fmaxf(fabsf(s_ch[0][pidx(i)]), fabsf(s_ch[1][pidx(i)])));
                for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
                if (lane == 0) s_red[warp] = mx;
                __syncthreads();
              }
              if (tid == 0) {
                if (has_tail) {
                  if (tail_frames == UINT64_MAX) h.tail_counter = tail_frames;
                  else if (h.tail_counter == UINT64_MAX) h.tail_counter = tail_frames;
                  else h.tail_counter = sat_sub_u64(h.tail_counter, len);
                  h.silence_counter = UINT64_MAX;
                } else {
                  for (uint32_t w = 0; w < nt / 32; ++w) mx = fmaxf(mx, s_red[w]);
                  if (mx < 0.001f) {
                    h.silence_counter = sat_add_u64(h.silence_counter, len);
                    if (h.silence_counter >= 2ull * sr) { h.tail_counter = 0; h.silence_counter = UINT64_MAX; }
                  } else {
                    h.silence_counter = 0;
                  }
                }
              }
            } else if (tid == 0) {
              h.tail_counter = UINT64_MAX; h.silence_counter = 0;
            }
            input_bypassed = false;
            all_bypassed = false;
          }
          __syncthreads();
        }
        tick(4, t0);
        if (last_stage) { eff_byp = all_bypassed ? 1u : 0u; if (tid == 0) a.mstate[m].effects_bypassed = eff_byp; }
        if (master_direct && last_stage) {
          // main mixer, master volume not ramping, no meter: the chunk goes from shared memory straight to the output
          // (nobody reads the main bus again), scaled as WavStream::process does per block (wav.rs:237)
          const float mxo = stage_out(a.out + (size_t)boff * 2, len, master_scale ? master_gain : 1.0f, master_scale);
          if (direct) { gate_mx = fmaxf(gate_mx, mxo); gate_covered += len; }   // (still a sub-mixer: its gate decides below)
          chunk_done = true;
        } else {
          // write the processed chunk back; a sub-mixer's silence gate wants max|x| of the parent chunk: collected here,
          // from shared memory, instead of re-reading the bus
          gate_mx = fmaxf(gate_mx, stage_out(gchunk, len, 1.0f, false));
          gate_covered += len;
        }
      } else if (n_stages > 1 && tid == 0) {
        for (uint32_t e = e_lo; e < e_hi; ++e) a.fx[e].bypassed = 1u;   // (the transition the skipped loop would have made)
        if (last_stage) a.mstate[m].effects_bypassed = all_bypassed ? 1u : 0u;
      }
      if (!last_stage) {  // hand chunk q to the next stage: samples first, then the flags, then the count
        if (!is_main && c1 == parent_next) { pk++; parent_next = a.chunk_bounds[pk + 1]; }   // (the gate itself is the last stage's business)
        __threadfence();
        __syncthreads();
        if (tid == 0) {
          pf_out[q] = (uint8_t)((input_bypassed ? 1u : 0u) | (all_bypassed ? 0u : 2u));
          __threadfence();
          atomicExch(prog_out, q + 1u);
        }
        ++q;
        continue;   // gate / master volume / output belong to the last stage
      }
      __syncthreads();
      ++q;
      tick(5, t0);
    }

    if (is_main && master_direct) {
      if (!chunk_done) for (uint32_t i = tid; i < len * 2; i += nt) a.out[(size_t)boff * 2 + i] = master_scale ? gchunk[i] * master_gain : gchunk[i];
    } else if (is_main) {
      // WavStream::process: apply_smoothed_gain once per 1024-frame block (wav.rs:228-237)
      if (((uint32_t)(c1 - a.block_start) % wbf) == 0 || k + 2 == ce) {
        const uint64_t wb0 = a.block_start + ((uint32_t)(c1 - 1 - a.block_start) / wbf) * wbf;
        const uint32_t o0 = (uint32_t)(wb0 - a.block_start);
        const uint32_t wl = (uint32_t)(c1 - wb0);
        float* wbuf = bus + (size_t)o0 * 2;
        if (a.meter) {  // MeteredSource::record of this block (metered.rs:107-128): per-channel peak and sum of squares
          __syncthreads();
          float pk = 0.0f;
          double sq = 0.0;
          // thread parity == channel: a thread only ever sees samples of one channel
          for (uint32_t i = tid; i < wl * 2; i += nt) { const float x = wbuf[i]; pk = fmaxf(pk, fabsf(x)); sq += (double)x * (double)x; }
          for (int o = 16; o >= 2; o >>= 1) { pk = fmaxf(pk, __shfl_xor_sync(0xFFFFFFFFu, pk, o)); sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o); }
          if (lane < 2) { s_scratch[0][warp * 4 + lane] = (double)pk; s_scratch[0][warp * 4 + 2 + lane] = sq; }
          __syncthreads();
          if (tid < 2) {
            double p2 = 0.0, s2 = 0.0;
            for (uint32_t w = 0; w < nt / 32; ++w) { p2 = fmax(p2, s_scratch[0][w * 4 + tid]); s2 += s_scratch[0][w * 4 + 2 + tid]; }
            double* row = a.meter + ((wb0 - a.render_start) / a.wav_block_frames) * 4;
            row[tid] = p2; row[2 + tid] = s2;
          }
          __syncthreads();
        }
        ExpSm ms = s_master;
        const bool ramp = exp_need_ramp(ms, a.fxc.comp);
        __syncthreads();
        if (ramp) {
          if (tid == 0) {
            for (uint32_t i = 0; i < wl * 2; ++i) wbuf[i] *= exp_next(ms, a.fxc.comp);
            *a.master = ms;
            s_master = ms;
          }
          __syncthreads();
          if (a.out) for (uint32_t i = tid; i < wl * 2; i += nt) a.out[(size_t)o0 * 2 + i] = wbuf[i];
        } else {
          const float g = ms.target;
          const bool scale = fabsf(1.0f - g) > 0.000001f;
          if (a.out) for (uint32_t i = tid; i < wl * 2; i += nt) a.out[(size_t)o0 * 2 + i] = scale ? wbuf[i] * g : wbuf[i];
        }
        __syncthreads();
      }
    } else {
      if (direct && master_direct && !chunk_done)   // (a chunk the effects did not touch, or a mixer without effects)
        for (uint32_t i = tid; i < len * 2; i += nt) a.out[(size_t)boff * 2 + i] = master_scale ? gchunk[i] * master_gain : gchunk[i];
      // parent chunk ends here? -> SubMixerProcessor::process gate over the parent chunk span
      if (c1 == parent_next) {
        const uint64_t p0 = a.chunk_bounds[pk];
        const uint32_t o0 = (uint32_t)(p0 - a.block_start);
        const uint32_t pl = (uint32_t)(c1 - p0);
        // frames of the span whose maximum the write-back above has not seen (chunks the effects skipped, mixers without effects)
        float mx = gate_mx;
        if (gate_covered != pl) { mx = 0.0f; for (uint32_t i = tid; i < pl * 2; i += nt) mx = fmaxf(mx, fabsf(bus[(size_t)o0 * 2 + i])); }
        gate_mx = 0.0f; gate_covered = 0;
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        if (lane == 0) s_red[warp] = mx;
        __syncthreads();
        if (tid == 0) {
          for (uint32_t w = 1; w < nt / 32; ++w) mx = fmaxf(mx, s_red[w]);
          MixerState& ms = a.mstate[m];
          uint32_t flag;
          if (mx < 0.001f) {
            ms.silence_counter += pl;
            flag = ms.silence_counter < 2ull * sr ? 1u : 0u;
          } else {
            ms.silence_counter = 0;
            flag = 1u;
          }
          a.mixer_flags[(size_t)m * a.max_chunks + (pk - pcb)] = (uint8_t)flag;
          s_run = flag;
        }
        __syncthreads();
        if (direct && master_direct && s_run == 0u)   // gated off: the main mixer does not add this span (submixer.rs:47-77)
          for (uint32_t i = tid; i < pl * 2; i += nt) a.out[(size_t)o0 * 2 + i] = 0.0f;
        __syncthreads();
        pk++;
        parent_next = a.chunk_bounds[pk + 1];
      }
    }
    tick(6, t0);
  }
  if ((is_main || direct) && last_stage) publish_progress();
  if (a.prof && tid == 0 && (is_main || a.prof_all))
    for (int i = 0; i < 8; ++i) atomicAdd(a.prof + (a.prof_all ? ((size_t)m * MAX_FX_STAGES + stage) * 8 : 0) + i, (unsigned long long)pt[i]);
}

}  // namespace pb
