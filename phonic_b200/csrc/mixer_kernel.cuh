// Mixer kernel: one CTA per mixer of one tree level. Walks the mixer's own chunk schedule for the time
// block; per chunk it (1) sums gated sub-mixer buses then source (group) buses in the reference's
// order (MixedSource::process_sub_mixers / process_sources, src/source/mixed.rs:505-624), (2) runs the
// effect chain with the auto-bypass state machine (EffectProcessor::process, mixed/effect.rs:56-145) on
// warp 0, (3) at every *parent* chunk end evaluates the sub-mixer silence gate
// (SubMixerProcessor::process, mixed/submixer.rs:47-77) with an exact max|x| reduction. The main mixer
// additionally applies the WavStream master volume per output block (src/output/wav.rs:237).
#pragma once
#include "effects.cuh"

namespace pb {

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
  uint64_t block_start;
  // main mixer output
  float* out;                     // device output for this block (interleaved stereo) or nullptr
  ExpSm* master;                  // WavStream::smoothed_volume
  uint32_t wav_block_frames;      // 1024
};

PB_DEV uint64_t sat_sub_u64(uint64_t a, uint64_t b) { return a > b ? a - b : 0; }
PB_DEV uint64_t sat_add_u64(uint64_t a, uint64_t b) { return (a > UINT64_MAX - b) ? UINT64_MAX : a + b; }

__global__ void __launch_bounds__(256) mixer_kernel(MixerKernelArgs a) {
  __shared__ float s_red[8];
  __shared__ uint32_t s_flag;
  const uint32_t m = a.level_mixers[blockIdx.x];
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  const MixerParams mp = a.mixers[m];
  const bool is_main = mp.parent == 0xFFFFFFFFu;
  float* bus = a.mixer_bus + (size_t)m * a.block_frames * 2;
  const uint32_t cb = a.mixer_chunk_begin[m], ce = a.mixer_chunk_begin[m + 1];
  uint32_t pcb = 0, pce = 0, pk = 0;
  if (!is_main) { pcb = a.mixer_chunk_begin[mp.parent]; pce = a.mixer_chunk_begin[mp.parent + 1]; pk = pcb; }
  const uint32_t sr = a.fxc.sample_rate;

  for (uint32_t k = cb; k + 1 < ce; ++k) {
    const uint64_t c0 = a.chunk_bounds[k], c1 = a.chunk_bounds[k + 1];
    const uint32_t len = (uint32_t)(c1 - c0);
    const uint32_t boff = (uint32_t)(c0 - a.block_start);
    float* cbuf = bus + (size_t)boff * 2;

    // effect parameter events due at this chunk start (MixedSource::process_events, mixed.rs:683)
    if (tid == 0) {
      for (uint32_t e = mp.fx_begin; e < mp.fx_end; ++e) {
        FxHeader& h = a.fx[e];
        while (h.ev_cursor < h.ev_end && a.fx_events[h.ev_cursor].time <= c0) {
          fx_apply_param(h, a.fxc, a.fx_events[h.ev_cursor]);
          h.ev_cursor++;
        }
      }
    }

    // (1) sum children then sources, in order
    bool audible = false;
    for (uint32_t ci = mp.child_begin; ci < mp.child_end; ++ci)
      audible |= a.mixer_flags[(size_t)a.child_index[ci] * a.max_chunks + (k - cb)] != 0;
    for (uint32_t si = mp.src_begin; si < mp.src_end; ++si)
      audible |= a.group_flags[(size_t)a.source_index[si] * a.max_chunks + (k - cb)] != 0;
    for (uint32_t i = tid; i < len * 2; i += nt) {
      float s = 0.0f;
      for (uint32_t ci = mp.child_begin; ci < mp.child_end; ++ci) {
        const uint32_t c = a.child_index[ci];
        if (a.mixer_flags[(size_t)c * a.max_chunks + (k - cb)]) s += a.mixer_bus[((size_t)c * a.block_frames + boff) * 2 + i];
      }
      for (uint32_t si = mp.src_begin; si < mp.src_end; ++si) {
        const uint32_t g = a.source_index[si];
        if (a.group_flags[(size_t)g * a.max_chunks + (k - cb)]) s += a.group_bus[((size_t)g * a.block_frames + boff) * 2 + i];
      }
      cbuf[i] = s;
    }
    __syncthreads();

    // (2) effects with auto-bypass (mixed.rs:627-655, mixed/effect.rs:56-145), warp 0
    if (warp == 0 && mp.fx_end > mp.fx_begin) {
      MixerState& ms = a.mstate[m];
      bool input_bypassed = !audible;
      if (!(ms.effects_bypassed && input_bypassed)) {
        bool all_bypassed = true;
        for (uint32_t e = mp.fx_begin; e < mp.fx_end; ++e) {
          FxHeader& h = a.fx[e];
          bool bypassed = h.bypassed != 0;
          uint64_t tail = h.tail_counter, silence = h.silence_counter;
          const bool should_bypass = input_bypassed && tail == 0 && silence == UINT64_MAX;
          if (should_bypass && !bypassed) bypassed = true;
          else if (!should_bypass && bypassed) { bypassed = false; tail = UINT64_MAX; silence = 0; }
          if (!bypassed) {
            fx_process(h, a.fxc, cbuf, len, lane);
            __syncwarp();
            if (input_bypassed) {
              uint64_t tail_frames;
              if (fx_process_tail(h, a.fxc, tail_frames)) {
                if (tail_frames == UINT64_MAX) tail = tail_frames;
                else if (tail == UINT64_MAX) tail = tail_frames;
                else tail = sat_sub_u64(tail, len);
                silence = UINT64_MAX;
              } else {
                float mx = 0.0f;
                for (uint32_t i = lane; i < len * 2; i += 32) mx = fmaxf(mx, fabsf(cbuf[i]));
                for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
                if (mx < 0.001f) {
                  silence = sat_add_u64(silence, len);
                  if (silence >= 2ull * sr) { tail = 0; silence = UINT64_MAX; }
                } else {
                  silence = 0;
                }
              }
            } else {
              tail = UINT64_MAX; silence = 0;
            }
            input_bypassed = false;
            all_bypassed = false;
          }
          __syncwarp();
          if (lane == 0) { h.bypassed = bypassed; h.tail_counter = tail; h.silence_counter = silence; }
          __syncwarp();
        }
        if (lane == 0) ms.effects_bypassed = all_bypassed;
      }
    }
    __syncthreads();

    if (is_main) {
      // WavStream::process: apply_smoothed_gain per 1024-frame block (wav.rs:228-237)
      if (((c1 - 0) % a.wav_block_frames) == 0 || k + 2 == ce) {
        // chunk ends a wav block: find the block start inside this time block
        const uint64_t wb0 = ((c1 - 1) / a.wav_block_frames) * a.wav_block_frames;
        const uint32_t o0 = (uint32_t)(wb0 - a.block_start);
        const uint32_t wl = (uint32_t)(c1 - wb0);
        float* wbuf = bus + (size_t)o0 * 2;
        ExpSm ms = *a.master;
        const bool ramp = exp_need_ramp(ms, a.fxc.comp);
        __syncthreads();
        if (ramp) {
          if (tid == 0) {
            for (uint32_t i = 0; i < wl * 2; ++i) wbuf[i] *= exp_next(ms, a.fxc.comp);
            *a.master = ms;
          }
        } else if (fabsf(1.0f - ms.target) > 0.000001f) {
          for (uint32_t i = tid; i < wl * 2; i += nt) wbuf[i] *= ms.target;
        }
        __syncthreads();
        if (a.out) for (uint32_t i = tid; i < wl * 2; i += nt) a.out[(size_t)o0 * 2 + i] = wbuf[i];
      }
    } else {
      // (3) parent chunk ends here? -> SubMixerProcessor::process gate over the parent chunk span
      if (c1 == a.chunk_bounds[pk + 1]) {
        const uint64_t p0 = a.chunk_bounds[pk];
        const uint32_t o0 = (uint32_t)(p0 - a.block_start);
        const uint32_t pl = (uint32_t)(c1 - p0);
        float mx = 0.0f;
        for (uint32_t i = tid; i < pl * 2; i += nt) mx = fmaxf(mx, fabsf(bus[(size_t)o0 * 2 + i]));
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        if (lane == 0) s_red[warp] = mx;
        __syncthreads();
        if (tid == 0) {
          for (uint32_t w = 1; w < (nt + 31) / 32; ++w) mx = fmaxf(mx, s_red[w]);
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
        }
        __syncthreads();
        pk++;
      }
    }
  }
  (void)s_flag; (void)pce;
}

}  // namespace pb
