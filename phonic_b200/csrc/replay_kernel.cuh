// Replay kernel (pass 2 of the voice path): renders every Segment the skeleton pass emitted, in parallel
// over (voice x 64-frame tile). One CTA covers TPC consecutive tiles of one group; thread (tile, voice)
// replays its segments with the full audio math -- cubic Hermite resampling from the shared sample
// buffer (read-only path, L1/L2 resident), fader, per-sample gain ramp, constant-power pan, AHDSR -- in
// 16-frame sub-tiles staged in shared memory, where the voices of a tile are summed in *voice order*
// (Sampler::write, sampler.rs:989-1006: out = ((0 + v0) + v1) + ...), the generator-level gain/pan is
// applied and the group bus is stored coalesced. Per-voice audio never touches HBM.
#pragma once
#include "skeleton_kernel.cuh"

namespace pb {

constexpr uint32_t SUB = 16;          // frames per shared-memory sub-tile
constexpr uint32_t ROW = 2 * SUB + 1; // padded row (floats): conflict-free row writes and column reads

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
  uint32_t vpad;          // voices per tile rounded up to a power of two
  uint32_t tpc;           // tiles per CTA
  RenderConsts rc;
  const float* hq_scratch;      // [n_hq][block_frames][2]: resampler output stream of HighQuality voices (sinc_kernel.cuh)
  const HqState* hq_states;     // [n_voices] (only `slot` is read here)
  const GranGroup* gran_groups; // [n_groups] or nullptr
  GranReplay gran;              // this block's grain records + contribution storage (gran.cuh)
};

template <int CC>
PB_DEV void replay_voice_subtile(VoiceState& v, CallCtx& cc, HistVals& hv, const Segment* __restrict__ segs, uint32_t& seg_i,
                                 uint32_t seg_end_i, uint32_t& seg_pos, uint32_t& seg_stop, bool& have, const GroupParams& gp,
                                 const DevBuffer& buf, const RenderConsts& rc, uint32_t sub_lo, uint32_t sub_hi, float* row,
                                 const float* __restrict__ hq_row, const GranReplay* __restrict__ gran, uint32_t gran_row) {
  // sub_lo/sub_hi: frame range of this sub-tile relative to the block
  while (have && seg_pos < sub_hi) {
    const uint32_t lo = max(seg_pos, sub_lo);
    const uint32_t hi = min(seg_stop, sub_hi);
    uint32_t wrote = 0;
    if (hi > lo) {
      if (gran) wrote = gran_replay_frames(v, cc, gp, *gran, gran_row, hi - lo, row + (lo - sub_lo) * 2);
      else if (v.hq) wrote = hq_replay_frames(v, cc, hq_row, rc.rate_comp, hi - lo, row + (lo - sub_lo) * 2);
      else wrote = voice_frames<CC, true>(v, cc, hv, gp, buf, rc.sample_rate, rc.rate_comp, hi - lo, row + (lo - sub_lo) * 2);
      seg_pos = lo + wrote;
    }
    if (seg_pos >= seg_stop || wrote < hi - lo) {  // segment done (or the source ran dry): next snapshot
      seg_i++;
      if (seg_i < seg_end_i) {
        const Segment& s = segs[seg_i];
        v = s.v; cc = s.c; seg_pos = s.out_off; seg_stop = s.out_off + s.n;
        hist_load<CC>(hv, v, buf.data);
      } else {
        have = false;
      }
    } else {
      break;  // sub-tile full
    }
  }
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) replay_kernel(ReplayArgs a) {
  extern __shared__ float smem[];
  const uint32_t g = a.group_list[blockIdx.y];
  const GroupParams gp = a.groups[g];
  const DevBuffer buf = a.buffers[gp.buffer];
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  const uint32_t vpad = a.vpad, tpc = a.tpc;
  const uint32_t tl = tid / vpad, vi = tid % vpad;        // tile within CTA, voice within group
  const uint32_t tile = blockIdx.x * tpc + tl;
  const bool is_sampler = gp.kind == GROUP_SAMPLER;
  const bool active_thread = vi < gp.n_voices && tile < a.n_tiles;

  float* rows = smem;                                      // [nt][ROW]
  float* g_gain = rows + (size_t)nt * ROW;                 // [tpc][2*TILE]
  float* g_pl = g_gain + (size_t)tpc * 2 * TILE;           // [tpc][TILE]
  float* g_pr = g_pl + (size_t)tpc * TILE;                 // [tpc][TILE]
  float* my_row = rows + (size_t)tid * ROW;

  // generator-level per-sample gain / per-frame pan for the whole tile (serial only while ramping)
  if (is_sampler && vi == 0 && tile < a.n_tiles) {
    float* gg = g_gain + (size_t)tl * 2 * TILE;
    float* pl = g_pl + (size_t)tl * TILE;
    float* pr = g_pr + (size_t)tl * TILE;
    for (uint32_t i = 0; i < TILE; ++i) { gg[2 * i] = 1.0f; gg[2 * i + 1] = 1.0f; pl[i] = 1.0f; pr[i] = 1.0f; }
    const uint32_t cnt = a.gseg_count[(size_t)g * a.n_tiles + tile];
    const uint32_t first = a.gseg_first[(size_t)g * a.n_tiles + tile];
    for (uint32_t s = 0; s < cnt; ++s) {
      GroupSeg gs = a.gsegs[(size_t)g * a.seg_cap + first + s];
      const uint32_t o = gs.out_off - tile * TILE;
      if (gs.flags & 1u) for (uint32_t i = 0; i < gs.n * 2; ++i) gg[o * 2 + i] = exp_next(gs.vol, a.rc.rate_comp);
      else if (gs.flags & 2u) for (uint32_t i = 0; i < gs.n * 2; ++i) gg[o * 2 + i] = gs.vol.target;
      if (gs.flags & 4u) for (uint32_t i = 0; i < gs.n; ++i) panning_factors(exp_next(gs.pan, a.rc.rate_comp), pl[o + i], pr[o + i]);
      else if (gs.flags & 8u) {
        float l, r;
        panning_factors(gs.pan.target, l, r);
        for (uint32_t i = 0; i < gs.n; ++i) { pl[o + i] = l; pr[o + i] = r; }
      }
    }
  }

  // first segment of this (voice, tile)
  VoiceState v;
  CallCtx cc;
  HistVals hv;
  uint32_t seg_i = 0, seg_end_i = 0, seg_pos = 0, seg_stop = 0;
  bool have = false;
  const Segment* segs = nullptr;
  const float* hq_row = nullptr;
  const bool is_gran = a.gran_groups != nullptr && a.gran_groups[g].enabled != 0;
  const GranReplay* gran = is_gran ? &a.gran : nullptr;
  const uint32_t gran_row = is_gran ? a.gran_groups[g].first_row + vi : 0u;
  if (active_thread) {
    const size_t vidx = gp.first_voice + vi;
    if (a.hq_states) hq_row = a.hq_scratch + (size_t)a.hq_states[vidx].slot * a.block_frames * 2;
    const uint32_t cnt = a.seg_count[vidx * a.n_tiles + tile];
    const uint32_t first = cnt ? a.seg_first[vidx * a.n_tiles + tile] : 0u;
    segs = a.segs + vidx * a.seg_cap;
    const TileRec rec = a.recs[vidx * a.n_tiles + tile];
    if (rec.gen == a.gen && (rec.stage_n & 0xFFFFu)) {
      // the tile opens inside a simple call: state = the call's Segment advanced by the record
      const Segment& s = segs[rec.base];
      v = s.v; cc = s.c;
      seg_pos = tile * TILE; seg_stop = seg_pos + (rec.stage_n & 0xFFFFu);
      if (buf.channels == 2) { apply_tile_rec<2>(v, cc, buf, rec, seg_pos - s.out_off); hist_load<2>(hv, v, buf.data); }
      else { apply_tile_rec<1>(v, cc, buf, rec, seg_pos - s.out_off); hist_load<1>(hv, v, buf.data); }
      seg_i = first - 1u;  // the tile's own segments (if any) follow
      seg_end_i = first + cnt;
      have = true;
    } else if (cnt) {
      seg_i = first;
      seg_end_i = seg_i + cnt;
      const Segment& s = segs[seg_i];
      v = s.v; cc = s.c; seg_pos = s.out_off; seg_stop = s.out_off + s.n;
      if (buf.channels == 2) hist_load<2>(hv, v, buf.data); else hist_load<1>(hv, v, buf.data);
      have = true;
    }
  }

  float* gbus = a.group_bus + (size_t)g * a.block_frames * 2;
  for (uint32_t st = 0; st < TILE / SUB; ++st) {
#pragma unroll
    for (uint32_t i = 0; i < 2 * SUB; ++i) my_row[i] = 0.0f;
    if (have) {
      const uint32_t sub_lo = tile * TILE + st * SUB, sub_hi = sub_lo + SUB;
      if (buf.channels == 2)
        replay_voice_subtile<2>(v, cc, hv, segs, seg_i, seg_end_i, seg_pos, seg_stop, have, gp, buf, a.rc, sub_lo, sub_hi, my_row, hq_row, gran, gran_row);
      else
        replay_voice_subtile<1>(v, cc, hv, segs, seg_i, seg_end_i, seg_pos, seg_stop, have, gp, buf, a.rc, sub_lo, sub_hi, my_row, hq_row, gran, gran_row);
    }
    __syncthreads();
    // ordered reduction over the voices of each tile, generator-level gain/pan, coalesced store
    for (uint32_t col = tid; col < tpc * 2 * SUB; col += nt) {
      const uint32_t rtl = col / (2 * SUB), c = col % (2 * SUB);
      const uint32_t rtile = blockIdx.x * tpc + rtl;
      if (rtile >= a.n_tiles) continue;
      const float* base = rows + (size_t)(rtl * vpad) * ROW + c;
      float s = 0.0f;
      if (is_sampler) {
        for (uint32_t i = 0; i < gp.n_voices; ++i) s += base[(size_t)i * ROW];
        const uint32_t f = st * SUB + (c >> 1);
        s *= g_gain[(size_t)rtl * 2 * TILE + f * 2 + (c & 1)];
        s *= (c & 1) ? g_pr[(size_t)rtl * TILE + f] : g_pl[(size_t)rtl * TILE + f];
      } else {
        s = base[0];
      }
      gbus[((size_t)rtile * TILE + st * SUB) * 2 + c] = s;
    }
    __syncthreads();
  }
}

}  // namespace pb
