// Replay kernel (pass 2 of the voice path): renders every Segment / TileRec the skeleton pass emitted, in parallel over
// (group x 64-frame tile). Thread (group g, tile t) walks the voices of g in *voice order* and adds each voice's frames
// of tile t to its own row of shared memory -- exactly Sampler::write's out = ((0 + v0) + v1) + ... (sampler.rs:989-1006)
// -- with the full audio math: cubic Hermite resampling from the shared sample buffer (read-only path, L1/L2 resident),
// fader, per-sample gain ramp, constant-power pan, AHDSR. Then the generator-level gain/pan is applied to the row and
// the warp stores its 32 rows coalesced to the group bus. No barrier, no cross-thread reduction: a warp holds 32
// consecutive tiles of ONE group and visits the same voice at the same time, so its lanes share the ratio, the call
// kind and the envelope stage (warp-uniform control flow in the resampler). Per-voice audio never touches HBM.
#pragma once
#include "skeleton_kernel.cuh"

namespace pb {

constexpr uint32_t ROWP = 2 * TILE + 1;   // row pitch (floats): lanes of a warp hit 32 different banks
constexpr uint32_t REPLAY_THREADS = 64;   // tiles per CTA
constexpr uint32_t REPLAY_SMEM = REPLAY_THREADS * (ROWP + SIMPLE_WIN) * sizeof(float);

struct ReplayArgs {
  const GroupParams* groups;
  const uint32_t* group_list;
  const DevBuffer* buffers;
  const Segment* segs;
  const uint16_t* seg_first;
  const uint16_t* seg_count;
  const TileRec* recs;    // [n_voices][n_tiles]
  uint32_t gen;
  const GroupSeg* gsegs;
  const uint16_t* gseg_first;
  const uint16_t* gseg_count;
  float* group_bus;       // [n_groups][block_frames][2]
  uint32_t seg_cap, n_tiles, block_frames;
  RenderConsts rc;
  const float* hq_scratch;      // [n_hq][block_frames][2]: resampler output stream of HighQuality voices (sinc_kernel.cuh)
  const HqState* hq_states;     // [n_voices] (only `slot` is read here)
  const GranGroup* gran_groups; // [n_groups] or nullptr
  GranReplay gran;              // this block's grain records + contribution storage (gran.cuh)
};

// All frames of one voice inside one tile: its TileRec / Segments replayed into `row` (tile-relative, interleaved stereo).
template <int CC>
PB_DEV void replay_voice_tile(const ReplayArgs& a, const GroupParams& gp0, const uint32_t g, const DevBuffer& buf, const size_t vidx, const uint32_t tile,
                              float* __restrict__ row, float* __restrict__ win, const bool acc, const float* __restrict__ hq_row,
                              const GranReplay* __restrict__ gran, const uint32_t gran_row) {
  const uint32_t cnt = a.seg_count[vidx * a.n_tiles + tile];
  const uint32_t first = cnt ? a.seg_first[vidx * a.n_tiles + tile] : 0u;
  const Segment* __restrict__ segs = a.segs + vidx * a.seg_cap;
  const TileRec rec = a.recs[vidx * a.n_tiles + tile];
  VoiceState v;
  CallCtx cc;
  HistVals hv;
  // the parameter version of the segment at hand (sampler parameter automation): the group's own record unless an event
  // switched versions
  GroupParams gp_alt;
  const GroupParams* gpp = &gp0;
  auto use_version = [&](const uint32_t idx) { if (idx == g) gpp = &gp0; else { gp_alt = a.groups[idx]; gpp = &gp_alt; } };
  uint32_t seg_i, seg_end_i = first + cnt, seg_pos, seg_stop;
  bool simple;
  if (rec.gen == a.gen && (rec.stage_n & 0xFFFFu)) {
    // the tile opens inside a simple call: state = the call's Segment advanced by the record
    const Segment& s = segs[rec.base];
    v = s.v; cc = s.c;
    use_version(s.gp_idx);
    seg_pos = tile * TILE; seg_stop = seg_pos + (rec.stage_n & 0xFFFFu);
    apply_tile_rec<CC>(v, cc, buf, rec, seg_pos - s.out_off);
    hist_load<CC>(hv, v, buf.data);
    seg_i = first - 1u;  // the tile's own segments (if any) follow
    simple = true;
  } else if (cnt) {
    seg_i = first;
    const Segment& s = segs[seg_i];
    v = s.v; cc = s.c; seg_pos = s.out_off; seg_stop = s.out_off + s.n;
    use_version(s.gp_idx);
    hist_load<CC>(hv, v, buf.data);
    simple = !gran && simple_call_start<CC>(v, cc, buf);
    if (simple) simple_call_prologue<CC>(v, cc, hv, buf);
  } else {
    return;  // the voice is silent in this tile
  }
  const uint32_t tile_lo = tile * TILE;
  for (;;) {
    float* out = row + (seg_pos - tile_lo) * 2;
    const uint32_t n = seg_stop - seg_pos;
    if (gran) {
      for (uint32_t o = 0; o < n; o += GRAN_CHUNK) gran_replay_frames(v, cc, *gpp, *gran, gran_row, min(GRAN_CHUNK, n - o), out + 2 * o, acc);
    } else if (v.hq) {
      hq_replay_frames(v, cc, hq_row, a.rc.rate_comp, n, out, acc);
    } else if (simple) {
      simple_frames<CC>(v, cc, hv, *gpp, buf, n, out, acc, win, REPLAY_THREADS);
    } else {
      voice_frames<CC, true>(v, cc, hv, *gpp, buf, a.rc.sample_rate, a.rc.rate_comp, n, out, acc);  // may run dry early
    }
    if (++seg_i >= seg_end_i) return;
    const Segment& s = segs[seg_i];
    v = s.v; cc = s.c; seg_pos = s.out_off; seg_stop = s.out_off + s.n;
    use_version(s.gp_idx);
    hist_load<CC>(hv, v, buf.data);
    simple = !gran && simple_call_start<CC>(v, cc, buf);
    if (simple) simple_call_prologue<CC>(v, cc, hv, buf);
  }
}

__global__ void __launch_bounds__(REPLAY_THREADS, 6) replay_kernel(ReplayArgs a) {
  extern __shared__ float smem[];
  const uint32_t g = a.group_list[blockIdx.y];
  const GroupParams gp = a.groups[g];
  const DevBuffer buf = a.buffers[gp.buffer];
  const uint32_t tid = threadIdx.x;
  const uint32_t tile = blockIdx.x * REPLAY_THREADS + tid;
  const bool is_sampler = gp.kind == GROUP_SAMPLER;
  float* row = smem + (size_t)tid * ROWP;
  float* win = smem + (size_t)REPLAY_THREADS * ROWP + tid;   // [SIMPLE_WIN][REPLAY_THREADS] input staging (simple_frames)
#pragma unroll 16
  for (uint32_t i = 0; i < 2 * TILE; ++i) row[i] = 0.0f;

  if (tile < a.n_tiles) {
    const bool is_gran = a.gran_groups != nullptr && a.gran_groups[g].enabled != 0;
    const GranReplay* gran = is_gran ? &a.gran : nullptr;
    for (uint32_t vi = 0; vi < gp.n_voices; ++vi) {
      const size_t vidx = gp.first_voice + vi;
      const float* hq_row = a.hq_states ? a.hq_scratch + (size_t)a.hq_states[vidx].slot * a.block_frames * 2 : nullptr;
      const uint32_t gran_row = is_gran ? a.gran_groups[g].first_row + vi : 0u;
      if (buf.channels == 2) replay_voice_tile<2>(a, gp, g, buf, vidx, tile, row, win, is_sampler, hq_row, gran, gran_row);
      else replay_voice_tile<1>(a, gp, g, buf, vidx, tile, row, win, is_sampler, hq_row, gran, gran_row);
    }
    // generator-level per-sample gain / per-frame pan (AmplifiedSource / PannedSource around a Sampler, player.rs:1075-1081)
    if (is_sampler) {
      const uint32_t cnt = a.gseg_count[(size_t)g * a.n_tiles + tile];
      const uint32_t first = cnt ? a.gseg_first[(size_t)g * a.n_tiles + tile] : 0u;
      for (uint32_t s = 0; s < cnt; ++s) {
        GroupSeg gs = a.gsegs[(size_t)g * a.seg_cap + first + s];
        float* o = row + (gs.out_off - tile * TILE) * 2;
        if (gs.flags & 1u) for (uint32_t i = 0; i < gs.n * 2; ++i) o[i] *= exp_next(gs.vol, a.rc.rate_comp);
        else if (gs.flags & 2u) for (uint32_t i = 0; i < gs.n * 2; ++i) o[i] *= gs.vol.target;
        if (gs.flags & 4u) {
          for (uint32_t i = 0; i < gs.n; ++i) {
            float l, r;
            panning_factors(exp_next(gs.pan, a.rc.rate_comp), l, r);
            o[2 * i] *= l; o[2 * i + 1] *= r;
          }
        } else if (gs.flags & 8u) {
          float l, r;
          panning_factors(gs.pan.target, l, r);
          for (uint32_t i = 0; i < gs.n; ++i) { o[2 * i] *= l; o[2 * i + 1] *= r; }
        }
      }
    }
  }
  __syncwarp();
  // the warp's 32 rows -> group bus, 128-byte coalesced stores
  const uint32_t lane = tid & 31u, w0 = tid & ~31u;
  float* gbus = a.group_bus + (size_t)g * a.block_frames * 2;
  for (uint32_t r = 0; r < 32; ++r) {
    const uint32_t rtile = blockIdx.x * REPLAY_THREADS + w0 + r;
    if (rtile >= a.n_tiles || rtile * TILE >= a.block_frames) break;
    const float* src = smem + (size_t)(w0 + r) * ROWP;
    float* dst = gbus + (size_t)rtile * TILE * 2;
    const uint32_t live = min(2u * TILE, (a.block_frames - rtile * TILE) * 2u);   // (a block size that is no multiple of 64 ends in a partial tile)
#pragma unroll
    for (uint32_t k = 0; k < 2 * TILE / 32; ++k) if (k * 32 + lane < live) dst[k * 32 + lane] = src[k * 32 + lane];
  }
}

}  // namespace pb
