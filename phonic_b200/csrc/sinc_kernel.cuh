// Sinc kernel: the arithmetic of rubato::SincFixedIn<f32>::process_into_buffer (cubic interpolation between the
// 4 nearest of 128 sub-phase filters, 256 taps each; call site src/utils/resampler/rubato.rs:100-103) for every
// HqRec the skeleton pass emitted for this time block. PARITY UNPINNED: rubato is not vendored.
//
// Mapping. A persistent CTA (one per SM) keeps the whole filter table in shared memory -- 128 rows x 256 taps,
// every row padded to 288 words with two leading zeros, so the row stride is a multiple of the 32 banks and a
// tap shifted by +-1 never leaves the row -- and takes batches of SINC_NB records. For every record it stages
// SincFixedIn's 768-frame buffer (the last two chunks + the open one, cut from the sample buffer with zero
// fill) planar in shared memory. One thread = one output frame (all channels):
//   idx = idx0 + (j + 1) * t_ratio            (closed form of rubato's `idx += t_ratio`; differs by < 1e-11 input
//                                              samples, the frame COUNT per chunk was fixed exactly by pass 1)
//   out = interp_cubic(frac, [dot(x[i_k ..], sinc[s_k]) for the 4 nearest (i_k, s_k)])
// The four sub-phases may straddle an index step (get_nearest_times_4); with the unified tap range p' in [-1, 256]
// all four read the same input x[index + p'] and only shift their coefficient by d_k in {-1, 0, +1}.
// Bank conflicts: output j (thread t; consecutive in a warp) walks the taps rotated by -j, i.e. tap (k - j) mod 258 in step k. A coefficient read is
// then at bank (k - t - d_k + 2) mod 32 whatever the sub-phase row, distinct across the warp; the input read is at
// index_t + k - t, which moves by t_ratio - 1 per lane: neighbouring words or a broadcast.
// Bound: shared-memory bandwidth. Per output 258 x (4 + CC) words against 258 x 4 x CC FFMA: 20 % (mono) /
// 33 % (stereo) of the FP32 pipe is the ceiling of any mapping that reads each coefficient once per output
// (DESIGN.md §4.5).
#pragma once
#include "hq.cuh"

namespace pb {

constexpr int SINC_NB = 4;            // records per batch
constexpr int SINC_THREADS = 768;
constexpr uint32_t SINC_ROW = 288;    // padded table row (words)
constexpr uint32_t SINC_TAPS = 258;   // unified tap range p' = -1 .. 256
constexpr uint32_t SINC_WIN = 3 * HQ_CHUNK;  // SincFixedIn buffer: chunk_size + 2 * sinc_len frames
constexpr size_t SINC_SMEM = ((size_t)HQ_FACTOR * SINC_ROW + (size_t)SINC_NB * 2 * SINC_WIN) * sizeof(float);

struct SincArgs {
  const HqRec* recs;
  const uint32_t* n_recs;
  uint32_t cap;
  const DevBuffer* buffers;
  const float* tables;     // [n_tables][128][256]
  float* scratch;          // [n_hq][block_frames][2]
  uint32_t block_frames;
  uint32_t table;          // the table this launch serves
  uint32_t do_copies;      // also serve the bypass (copy) records
  unsigned long long* frames_out;  // statistics: output frames materialised
};

PB_DEV float sinc_interp_cubic(float x, float y0, float y1, float y2, float y3) {
  const float a0 = y1;
  const float a1 = -(1.0f / 3.0f) * y0 - 0.5f * y1 + y2 - (1.0f / 6.0f) * y3;
  const float a2 = 0.5f * (y0 + y2) - y1;
  const float a3 = 0.5f * (y1 - y2) + (1.0f / 6.0f) * (y3 - y0);
  const float x2 = x * x;
  const float x3 = x2 * x;
  return a0 + a1 * x + a2 * x2 + a3 * x3;
}

__global__ void __launch_bounds__(SINC_THREADS, 1) sinc_kernel(SincArgs a) {
  extern __shared__ float sinc_smem[];
  float* tab = sinc_smem;                                   // [128][SINC_ROW]
  float* win = sinc_smem + (size_t)HQ_FACTOR * SINC_ROW;    // [SINC_NB][2][SINC_WIN] planar
  __shared__ HqRec s_rec[SINC_NB];
  __shared__ uint32_t s_pref[SINC_NB + 1];
  const uint32_t tid = threadIdx.x;
  const uint32_t n = min(*a.n_recs, a.cap);
  if (blockIdx.x * SINC_NB >= n) return;

  if (a.tables) {  // the filter table of this launch, rows padded: word q of a row holds tap q - 2
    const float* src = a.tables + (size_t)a.table * HQ_FACTOR * HQ_CHUNK;
    for (uint32_t i = tid; i < HQ_FACTOR * SINC_ROW; i += SINC_THREADS) {
      const uint32_t row = i / SINC_ROW, q = i % SINC_ROW;
      tab[i] = (q >= 2 && q < 2 + HQ_CHUNK) ? __ldg(src + row * HQ_CHUNK + (q - 2)) : 0.0f;
    }
  }

  for (uint32_t base = blockIdx.x * SINC_NB; base < n; base += gridDim.x * SINC_NB) {
    __syncthreads();
    if (tid < SINC_NB) {
      HqRec r;
      r.count = 0;
      if (base + tid < n) {
        r = a.recs[base + tid];
        const bool mine = r.kind == 0 ? r.table == a.table : a.do_copies != 0;
        if (!mine) r.count = 0;
        // clip to the block
        if (r.out_off >= a.block_frames) r.count = 0;
        else r.count = min(r.count, a.block_frames - r.out_off);
      }
      s_rec[tid] = r;
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0;
      for (int i = 0; i < SINC_NB; ++i) { s_pref[i] = acc; acc += s_rec[i].count; }
      s_pref[SINC_NB] = acc;
      if (acc) atomicAdd(a.frames_out, (unsigned long long)acc);
    }
    // stage SincFixedIn's buffer of every sinc record: [chunk k-2 | chunk k-1 | chunk k], planar, zero filled
    for (uint32_t i = tid; i < SINC_NB * SINC_WIN; i += SINC_THREADS) {
      const uint32_t rr = i / SINC_WIN, f = i % SINC_WIN;
      const HqRec& r = s_rec[rr];
      if (r.count == 0 || r.kind != 0) continue;
      const uint32_t ck = f / HQ_CHUNK, o = f % HQ_CHUNK;
      const DevBuffer b = a.buffers[r.buffer];
      float x0 = 0.0f, x1 = 0.0f;
      if (o < r.valid[ck]) {
        if (b.channels == 2) {
          const float2 x = __ldg(reinterpret_cast<const float2*>(b.data + r.src[ck]) + o);
          x0 = x.x; x1 = x.y;
        } else {
          x0 = __ldg(b.data + r.src[ck] + o);
          x1 = x0;
        }
      }
      win[(size_t)(rr * 2) * SINC_WIN + f] = x0;
      win[(size_t)(rr * 2 + 1) * SINC_WIN + f] = x1;
    }
    __syncthreads();
    const uint32_t total = s_pref[SINC_NB];
    for (uint32_t o = tid; o < total; o += SINC_THREADS) {
      uint32_t rr = 0;
#pragma unroll
      for (int i = 1; i < SINC_NB; ++i) rr += o >= s_pref[i] ? 1u : 0u;
      const HqRec& r = s_rec[rr];
      const uint32_t jj = o - s_pref[rr];
      const uint32_t j = r.skip + jj;
      float* dst = a.scratch + ((size_t)r.slot * a.block_frames + r.out_off + jj) * 2;
      const DevBuffer b = a.buffers[r.buffer];
      if (r.kind != 0) {  // equal-rate bypass: a copy of `valid` frames followed by the zero padding
        float x0 = 0.0f, x1 = 0.0f;
        if (j < r.valid[2]) {
          x0 = __ldg(b.data + r.src[2] + (size_t)j * b.channels);
          x1 = b.channels == 2 ? __ldg(b.data + r.src[2] + (size_t)j * 2 + 1) : x0;
        }
        *reinterpret_cast<float2*>(dst) = make_float2(x0, x1);
        continue;
      }
      const double idx = fma((double)(j + 1), r.t_ratio, r.idx0);
      const double fl = floor(idx);
      const int index = (int)fl;
      const int sub = (int)floor((idx - fl) * (double)HQ_FACTOR);
      const double t128 = idx * (double)HQ_FACTOR;
      const float frac = (float)(t128 - floor(t128));
      // get_nearest_times_4: sub-phases sub-1 .. sub+2 with carries into index -1 / +1
      const float* crow[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int s = sub - 1 + k, d = 0;
        if (s < 0) { s += (int)HQ_FACTOR; d = -1; }
        else if (s >= (int)HQ_FACTOR) { s -= (int)HQ_FACTOR; d = 1; }
        crow[k] = tab + (size_t)s * SINC_ROW + 2 - d;   // coefficient of unified tap p' is crow[k][p']
      }
      // unified taps p' = -1 .. 256 read x[index + 2 * sinc_len + p']; m = p' + 1 is rotated by -j
      const float* xl = win + (size_t)(rr * 2) * SINC_WIN + (index + 2 * (int)HQ_CHUNK - 1);
      const float* xr = xl + SINC_WIN;
      const bool stereo = b.channels == 2;
      uint32_t m = (SINC_TAPS - (j % SINC_TAPS)) % SINC_TAPS;  // by output index, not thread: the sum order of an output is fixed
      float al0 = 0.0f, al1 = 0.0f, al2 = 0.0f, al3 = 0.0f;
      float ar0 = 0.0f, ar1 = 0.0f, ar2 = 0.0f, ar3 = 0.0f;
      if (stereo) {
#pragma unroll 6
        for (uint32_t k = 0; k < SINC_TAPS; ++k) {
          const int pp = (int)m - 1;
          const float c0 = crow[0][pp], c1 = crow[1][pp], c2 = crow[2][pp], c3 = crow[3][pp];
          const float vl = xl[m], vr = xr[m];
          al0 = __fmaf_rn(vl, c0, al0); al1 = __fmaf_rn(vl, c1, al1); al2 = __fmaf_rn(vl, c2, al2); al3 = __fmaf_rn(vl, c3, al3);
          ar0 = __fmaf_rn(vr, c0, ar0); ar1 = __fmaf_rn(vr, c1, ar1); ar2 = __fmaf_rn(vr, c2, ar2); ar3 = __fmaf_rn(vr, c3, ar3);
          m = m + 1 == SINC_TAPS ? 0u : m + 1;
        }
      } else {
#pragma unroll 6
        for (uint32_t k = 0; k < SINC_TAPS; ++k) {
          const int pp = (int)m - 1;
          const float c0 = crow[0][pp], c1 = crow[1][pp], c2 = crow[2][pp], c3 = crow[3][pp];
          const float vl = xl[m];
          al0 = __fmaf_rn(vl, c0, al0); al1 = __fmaf_rn(vl, c1, al1); al2 = __fmaf_rn(vl, c2, al2); al3 = __fmaf_rn(vl, c3, al3);
          m = m + 1 == SINC_TAPS ? 0u : m + 1;
        }
      }
      const float yl = sinc_interp_cubic(frac, al0, al1, al2, al3);
      const float yr = stereo ? sinc_interp_cubic(frac, ar0, ar1, ar2, ar3) : yl;
      *reinterpret_cast<float2*>(dst) = make_float2(yl, yr);
    }
  }
}

}  // namespace pb
