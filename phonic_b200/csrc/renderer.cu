// phonic_b200 renderer: host graph builder + schedule compiler + CUDA launches behind the C-ABI of
// include/phonic_b200.h. No CPU fallback: every render path launches the kernels of this unit.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <atomic>
#include <vector>

#include "../../include/phonic_b200.h"
#include "replay_kernel.cuh"
#include "wav_io.h"
#include "sinc_kernel.cuh"
#include "mixer_kernel.cuh"
#include "host_fx.h"

using namespace pb;

namespace {

#define CUDA_TRY(expr)                                                                           \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      r->last_error = std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr;      \
      return PB200_ERR_CUDA;                                                                     \
    }                                                                                            \
  } while (0)

// Process-wide device memory pool (per device, power-of-two size classes): offline batch rendering
// creates and drops many short-lived renderers, and cudaMalloc/cudaFree (which synchronise the device)
// would otherwise dominate the end-to-end time of a render. Blocks are only handed back to the pool when
// the owning renderer is idle (every render call ends with a stream synchronise).
struct DevicePool {
  std::mutex mu;
  std::map<std::pair<int, size_t>, std::vector<void*>> free_blocks;
  std::vector<cudaEvent_t> free_events;
  static DevicePool& get() { static DevicePool p; return p; }
  static size_t size_class(size_t bytes) { size_t c = 256; while (c < bytes) c <<= 1; return c; }
  cudaError_t alloc(void** out, size_t bytes, size_t* granted) {
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t c = size_class(bytes);
    {
      std::lock_guard<std::mutex> g(mu);
      auto& v = free_blocks[{dev, c}];
      if (!v.empty()) { *out = v.back(); v.pop_back(); *granted = c; return cudaSuccess; }
    }
    cudaError_t e = cudaMalloc(out, c);
    if (e == cudaErrorMemoryAllocation) {  // the pool itself may hold the memory: give this device's idle blocks back and retry
      cudaGetLastError();
      trim(dev);
      e = cudaMalloc(out, c);
    }
    *granted = c;
    return e;
  }
  // cudaFree every pooled (idle) block of `dev` (-1: all devices); returns the bytes released (pb200_trim_pool)
  size_t trim(int dev) {
    std::vector<std::pair<int, void*>> victims;
    size_t bytes = 0;
    {
      std::lock_guard<std::mutex> g(mu);
      for (auto& kv : free_blocks) {
        if (dev >= 0 && kv.first.first != dev) continue;
        for (void* p : kv.second) { victims.push_back({kv.first.first, p}); bytes += kv.first.second; }
        kv.second.clear();
      }
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& v : victims) { cudaSetDevice(v.first); cudaFree(v.second); }
    cudaSetDevice(cur);
    return bytes;
  }
  void release(void* p, size_t cls) {
    if (!p) return;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    free_blocks[{dev, cls}].push_back(p);
  }
  cudaError_t event(cudaEvent_t* ev) {
    {
      std::lock_guard<std::mutex> g(mu);
      if (!free_events.empty()) { *ev = free_events.back(); free_events.pop_back(); return cudaSuccess; }
    }
    return cudaEventCreate(ev);
  }
  void release_event(cudaEvent_t ev) { std::lock_guard<std::mutex> g(mu); free_events.push_back(ev); }
  // idle non-blocking streams per (device, priority): creating four streams costs a short-lived renderer ~0.2 ms
  std::map<std::pair<int, int>, std::vector<cudaStream_t>> free_streams;
  cudaError_t stream(cudaStream_t* s, int dev, int prio) {
    {
      std::lock_guard<std::mutex> g(mu);
      auto& v = free_streams[{dev, prio}];
      if (!v.empty()) { *s = v.back(); v.pop_back(); return cudaSuccess; }
    }
    return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, prio);
  }
  void release_stream(cudaStream_t s, int dev, int prio) { std::lock_guard<std::mutex> g(mu); free_streams[{dev, prio}].push_back(s); }   // (caller has drained it)
};

template <class T>
struct DevVec {  // grow-only device array backed by the pool
  T* p = nullptr;
  size_t cap = 0;       // elements
  size_t cls = 0;       // pool size class in bytes
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    void* np = nullptr;
    size_t granted = 0;
    cudaError_t e = DevicePool::get().alloc(&np, n * sizeof(T), &granted);
    if (e != cudaSuccess) return e;
    if (p) DevicePool::get().release(p, cls);
    p = (T*)np; cls = granted; cap = granted / sizeof(T);
    return cudaSuccess;
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = reserve(std::max<size_t>(v.size(), 1));
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
  cudaError_t download(std::vector<T>& v, cudaStream_t s) const {
    if (v.empty()) return cudaSuccess;
    return cudaMemcpyAsync(v.data(), p, v.size() * sizeof(T), cudaMemcpyDeviceToHost, s);
  }
  // grow keeping the first `keep` elements (device-to-device copy) and zeroing everything behind them
  cudaError_t grow_preserve(size_t n, size_t keep, cudaStream_t s) {
    if (n <= cap) return cudaSuccess;
    void* np = nullptr;
    size_t granted = 0;
    cudaError_t e = DevicePool::get().alloc(&np, n * sizeof(T), &granted);
    if (e != cudaSuccess) return e;
    keep = std::min(keep, cap);
    if ((e = cudaMemsetAsync((char*)np + keep * sizeof(T), 0, granted - keep * sizeof(T), s)) != cudaSuccess) return e;
    if (p && keep && (e = cudaMemcpyAsync(np, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    if (p) DevicePool::get().release(p, cls);
    p = (T*)np; cls = granted; cap = granted / sizeof(T);
    return cudaSuccess;
  }
  void free() { if (p) DevicePool::get().release(p, cls); p = nullptr; cap = 0; cls = 0; }
};

// generation tags of TileRec records: unique per skeleton launch in this process, never 0
uint32_t next_generation();
// `n` consecutive generation tags (none of them 0)
uint32_t reserve_generations(uint32_t n) {
  for (;;) {
    uint32_t first = next_generation();
    bool ok = true;
    for (uint32_t i = 1; i < n; ++i) { uint32_t g = next_generation(); if (g != first + i) { ok = false; break; } }
    if (ok && first + n > first) return first;
  }
}
uint32_t next_generation() {
  static std::atomic<uint32_t> g{0};
  uint32_t v = g.fetch_add(1) + 1;
  if (v == 0) v = g.fetch_add(1) + 1;
  return v;
}

// cuStreamWaitValue32 through the runtime's driver entry point query (no link-time dependency on libcuda)
typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWaitValue32Fn stream_wait_value32() {
  static StreamWaitValue32Fn fn = []() -> StreamWaitValue32Fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return (StreamWaitValue32Fn)p;
  }();
  return fn;
}

struct HostBuffer { DevBuffer dev; size_t cls; };

struct HostEvent { DevEvent ev; uint64_t seq; double raw_speed = 0.0; uint32_t param_id = 0; };  // raw_speed: before the sampler's pitch factor

struct HostGroup {
  uint32_t public_id;
  GroupParams gp;
  std::vector<HostEvent> events;  // pending (not yet consumed) events, sorted at compile time
  // automatable Sampler state (sampler.rs:79-86): base pitch + the AHDSR times the rates in `gp` were derived from
  int32_t base_transpose = 0, base_finetune = 0;
  pb200_ahdsr env{};
  bool removed = false;           // MixerMessage::RemoveSource / its mixer was removed
};

struct HostFx {
  uint32_t public_id, kind, mixer;  // dense mixer index
  std::vector<FxParamEvent> events;
  std::vector<uint64_t> seqs;
};

struct HostMixer {
  uint32_t public_id;
  uint32_t parent;  // dense index or 0xFFFFFFFF
  uint32_t depth;
  bool removed = false;  // MixerMessage::RemoveMixer: detached from its parent, never rendered again
  std::vector<uint32_t> children, sources, effects;
};

constexpr uint32_t RING = 3;

}  // namespace

// The events of a render call. While the call runs they guard its exits: however it ends (each CUDA_TRY may return early)
// the renderer's streams are drained first -- no kernel of a failed call is left in flight -- and the events go back to the
// pool. After a successful call the set is kept by the renderer until somebody asks for the per-pass spans
// (pb200_last_render_stats): ~50 cudaEventElapsedTime calls that a caller who only wants the audio does not pay for.
struct SpanEvents {
  std::vector<cudaStream_t> streams;
  std::vector<cudaEvent_t> all;
  std::vector<cudaEvent_t> v0, v1, r0, r1, m0, m1, x;
  cudaEvent_t start = nullptr, end = nullptr, skel_end = nullptr;
  bool persistent = false, complete = false;
  cudaError_t get(cudaEvent_t* e) { const cudaError_t c = DevicePool::get().event(e); if (c == cudaSuccess) all.push_back(*e); return c; }
  ~SpanEvents() {
    if (!complete) for (cudaStream_t st : streams) if (st) cudaStreamSynchronize(st);
    for (cudaEvent_t e : all) DevicePool::get().release_event(e);
  }
};

// Pinned, device-mapped progress words (one per renderer) out of one process-wide slab: cudaHostAlloc / cudaFreeHost per
// renderer would synchronise the device every time a short-lived renderer comes or goes.
class ProgressWords {
 public:
  static ProgressWords& get() { static ProgressWords p; return p; }
  unsigned long long* take() {
    std::lock_guard<std::mutex> g(mu_);
    if (free_.empty()) {
      unsigned long long* slab = nullptr;
      if (cudaHostAlloc((void**)&slab, SLAB * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
      for (size_t i = 0; i < SLAB; ++i) free_.push_back(slab + (SLAB - 1 - i));
    }
    unsigned long long* w = free_.back();
    free_.pop_back();
    return w;
  }
  void give(unsigned long long* w) { if (w) { std::lock_guard<std::mutex> g(mu_); free_.push_back(w); } }
 private:
  static constexpr size_t SLAB = 256;
  std::mutex mu_;
  std::vector<unsigned long long*> free_;
};

struct pb200_renderer {
  pb200_config cfg;
  std::string last_error;
  int device = 0;
  cudaStream_t sv = nullptr, sm = nullptr;
  int prio[3] = {0, 0, 0};   // priorities of sv, sr_/sr2, sm (the pool hands streams out per device and priority)
  RenderConsts rc;

  std::vector<HostBuffer> buffers;
  std::map<uint32_t, uint32_t> gran_buffer_of;   // file buffer -> its mono, output-rate copy for granular playback
  std::vector<HostMixer> mixers;            // dense; [0] = main
  std::map<uint32_t, uint32_t> mixer_by_id; // public -> dense
  std::vector<HostGroup> groups;
  std::map<uint32_t, uint32_t> group_by_id;
  std::vector<HostFx> fxs;
  std::map<uint32_t, uint32_t> fx_by_id;
  uint32_t next_source_id = 1, next_mixer_id = 1, next_effect_id = 1;
  uint64_t next_note_id = 1, next_seq = 1;

  // host mirrors of device state
  std::vector<VoiceState> h_voices;
  std::vector<HqState> h_hq;                // parallel to h_voices; only HighQuality voices use their entry
  uint32_t n_hq = 0;                        // HighQuality voices (= rows of the stream scratch)
  std::vector<uint32_t> sinc_table_keys;    // f32 bits of each table's cutoff
  std::vector<float> sinc_tables;           // [n_tables][128][256]
  bool sinc_tables_dirty = false;
  std::vector<GranGroup> gran_groups;       // parallel to groups (enabled = 0 for everything but granular samplers)
  std::vector<GranState> h_gran;            // one per voice of a granular sampler ("row")
  uint32_t n_gran_rows = 0;
  std::vector<float> grain_luts;            // [8][2048] GRAIN_WINDOW_LUT, built once
  bool grain_luts_dirty = false;
  uint32_t gran_parity = 0;                 // carry double buffer: flips every time block
  std::vector<GroupState> h_gstate;
  std::vector<MixerState> h_mstate;
  std::vector<FxHeader> h_fx;
  std::vector<uint8_t> h_fx_state;
  size_t aux_doubles = 0;
  bool host_state_valid = true;  // host mirrors are newer or equal to device
  bool graph_dirty = true;
  size_t dev_n_gran_rows = 0;
  size_t dev_n_voices = 0, dev_n_groups = 0, dev_n_mixers = 0, dev_n_fx = 0, dev_fx_state_bytes = 0;

  // device arrays
  DevVec<DevBuffer> d_buffers;
  DevVec<VoiceState> d_voices;
  DevVec<HqState> d_hq;
  DevVec<float> d_sinc_tables, d_hq_scratch;
  DevVec<HqRec> d_hq_recs;
  DevVec<uint32_t> d_hq_nrecs;
  DevVec<unsigned long long> d_hq_frames;
  DevVec<GranGroup> d_gran_groups;
  DevVec<GranState> d_gran_states;
  DevVec<GrainRec> d_grain_recs;
  DevVec<uint32_t> d_gran_counters, d_gran_vrec, d_gran_tiles;
  DevVec<float2> d_grain_storage;
  DevVec<GrainCarry> d_grain_carry[2];      // double buffer, each [n_rows][GRAIN_POOL]: rows keep their offset when samplers are added
  size_t carry_rows = 0;                    // rows of the carry buffers that hold live state
  DevVec<float> d_grain_luts;
  DevVec<GroupParams> d_groups;
  DevVec<GroupState> d_gstate;
  DevVec<DevEvent> d_events;
  DevVec<MixerParams> d_mixers;
  DevVec<MixerState> d_mstate;
  DevVec<uint32_t> d_child_index, d_source_index, d_level_mixers, d_class_groups;
  DevVec<FxHeader> d_fx;
  DevVec<FxParamEvent> d_fx_events;
  DevVec<uint8_t> d_fx_state;
  DevVec<double> d_aux;
  size_t d_aux_used = 0;
  DevVec<uint64_t> d_bounds;
  DevVec<uint32_t> d_chunk_begin;
  DevVec<float> d_group_bus, d_mixer_bus, d_out;
  DevVec<Segment> d_segs;
  DevVec<GroupSeg> d_gsegs;
  DevVec<uint16_t> d_seg_first, d_seg_count, d_gseg_first, d_gseg_count;
  DevVec<TileRec> d_recs;
  DevVec<uint32_t> d_block_done, d_quiet_block, d_auton, d_stage_begin, d_fx_progress;
  DevVec<uint8_t> d_fx_pflags;
  // exact phase-jump tables (phase_table.cuh): one per steady resampling ratio seen so far, built on the device
  std::map<uint32_t, uint32_t> phase_off;   // f32 bits of the ratio -> word offset of its table
  DevVec<uint32_t> d_phase_tabs;
  DevVec<uint2> d_phase_dir, d_phase_new;
  size_t phase_words = 0;
  cudaStream_t sr_ = nullptr, sr2 = nullptr;  // replay streams (sr2: every other block of a thin graph)
  DevVec<uint8_t> d_group_flags, d_mixer_flags;
  DevVec<ExpSm> d_master;
  ExpSm h_master;
  bool master_uploaded = false;

  // observability: PlaybackStatusEvent stream + MeteredSource state of the main mixer
  DevVec<StatusRec> d_status;
  DevVec<uint32_t> d_status_count;
  std::vector<pb200_status_event> status_events;
  DevVec<double> d_meter;
  uint64_t meter_interval = UINT64_MAX;     // frames; UINT64_MAX = metering off
  uint64_t meter_clock = 0, meter_frames = 0;
  float meter_peak_hold[2] = {0.0f, 0.0f};
  double meter_sum_square[2] = {0.0, 0.0};
  pb200_audio_level audio_level{};
  uint64_t position = 0;  // frames
  bool finished = false;
  // pb200_render_progress: output frames that are final (read from other host threads while a render runs)
  std::unique_ptr<SpanEvents> spans;        // events of the last render call whose spans have not been read yet
  unsigned long long* progress = nullptr;   // pinned, mapped host word: the main mixer's last CTA of a time block stores into it
  uint64_t progress_base = 0;               // frames of all earlier render calls
  const float* ext_input[PB_MAX_EXT] = {};   // pb200_set_main_inputs: device stereo buses the next render adds to the main mixer
  uint32_t n_ext = 0;
  uint64_t ext_frames = 0;
  cudaStream_t sc = nullptr;                 // pb200_push_async: copies to a peer (DMA, no SMs)
  uint32_t* push_flags = nullptr;            // pinned ring of flag values in flight
  uint32_t push_seq = 0;
  uint32_t time_block = 32768;
  uint64_t voice_frames_total = 0;
  pb200_render_stats stats{};
};

namespace {

int fail(pb200_renderer* r, int code, const std::string& msg) {
  if (r) r->last_error = msg;
  return code;
}

// std::time::Duration::as_secs_f32 / as_secs_f64 from as_nanos()
float nanos_as_secs_f32(uint64_t n) { return (float)(n / 1000000000ull) + (float)(uint32_t)(n % 1000000000ull) / 1.0e9f; }
double nanos_as_secs_f64(uint64_t n) { return (double)(n / 1000000000ull) + (double)(uint32_t)(n % 1000000000ull) / 1.0e9; }

uint32_t f64_as_u32_h(double v) {
  if (!(v > 0.0)) return 0u;
  if (v >= 4294967295.0) return 0xFFFFFFFFu;
  return (uint32_t)v;
}

// VolumeFader::start inertia (src/utils/fader.rs:85-90), f32 math as in the reference
float fader_inertia(uint32_t sr, uint64_t nanos) {
  const float LN100 = 4.605f;
  float samples_duration = (float)sr * nanos_as_secs_f32(nanos) / LN100;
  return 1.0f - std::exp(-1.0f / samples_duration);
}

double speed_from_note_h(uint32_t note) {  // src/utils.rs:67-78
  double p = 440.0 * std::pow(2.0, ((double)(uint8_t)note - 69.0) / 12.0);
  double p60 = 440.0 * std::pow(2.0, (60.0 - 69.0) / 12.0);
  return p / p60;
}

VoiceState default_voice(const DevBuffer& b, uint32_t out_rate, double speed) {
  VoiceState v;
  std::memset(&v, 0, sizeof(v));
  v.current_speed = speed; v.target_speed = speed;
  v.end_frame = UINT64_MAX;
  v.repeat = b.loop_start >= 0 ? REPEAT_FOREVER : 0;
  v.repeat_count = v.repeat;
  v.loop_ovr_start = -1; v.loop_ovr_end = -1;
  for (int i = 0; i < 4; ++i) v.hidx[i] = -1;  // CubicInterpolator::new: input = [0.0; 4]
  uint32_t rate = f64_as_u32_h((double)out_rate / speed);  // file/common.rs:78-82
  v.ratio = (float)((double)b.sample_rate / (double)rate);
  v.fader_state = FADER_STOPPED; v.fader_cur = 1.0f; v.fader_tgt = 1.0f; v.fader_inertia = 1.0f;
  v.vol = ExpSm{1.0f, 1.0f};
  v.pan = ExpSm{0.0f, 0.0f};
  v.env_stage = ENV_IDLE;
  v.note = 60; v.note_volume = 1.0f;
  return v;
}

// AhdsrParameters::new_with_scaling + set_sample_rate (src/utils/ahdsr.rs:75-136, 160-313)
bool resolve_ahdsr(const pb200_ahdsr& a, uint32_t sr, GroupParams& gp) {
  const float F32_MAX = 3.402823466e+38f;
  auto in_unit = [](float s) { return s >= -1.0f && s <= 1.0f; };
  if (!in_unit(a.attack_scaling) || !in_unit(a.decay_scaling) || !in_unit(a.release_scaling)) return false;
  if (!(a.sustain_level >= 0.0f && a.sustain_level <= 1.0f)) return false;
  // after set_sample_rate(sr) -> setup(): all rates use `sr` and the final sustain level
  float at = nanos_as_secs_f32(a.attack_nanos);
  gp.attack_rate = at == 0.0f ? F32_MAX : 1.0f / (at * (float)sr);
  gp.hold_is_zero = a.hold_nanos == 0;
  gp.hold_samples = nanos_as_secs_f32(a.hold_nanos) * (float)sr;
  gp.decay_is_zero = a.decay_nanos == 0;
  gp.decay_rate = a.decay_nanos == 0 ? F32_MAX : (1.0f - a.sustain_level) / (nanos_as_secs_f32(a.decay_nanos) * (float)sr);
  gp.sustain_level = a.sustain_level;
  float rt = nanos_as_secs_f32(a.release_nanos);
  gp.release_is_zero = a.release_nanos == 0;
  gp.release_rate = rt == 0.0f ? F32_MAX : 1.0f / (rt * (float)sr);
  gp.attack_scaling = a.attack_scaling; gp.decay_scaling = a.decay_scaling; gp.release_scaling = a.release_scaling;
  gp.has_env = 1;
  return true;
}

// Duration::from_secs_f32(seconds.max(0.0)): the exact value rounded to the nearest nanosecond
uint64_t secs_f32_to_nanos(float s) { return (uint64_t)std::nearbyint((double)std::max(s, 0.0f) * 1.0e9); }

// Sampler parameter descriptors (src/generator/sampler.rs:94-183): plain value of a raw / normalized update
// (Sampler::parameter_update_value / parameter_update_value_integer, sampler.rs:862-906). Returns false for unknown ids.
bool sampler_param_plain(uint32_t id, float v, bool normalized, bool has_env, float& out) {
  const float n = std::min(std::max(v, 0.0f), 1.0f);
  auto cl = [](float x, float lo, float hi) { return std::min(std::max(x, lo), hi); };
  if (id == pbh::cc4("STRN")) { out = normalized ? std::round(-48.0f + n * 96.0f) : (float)std::min(std::max((int32_t)v, -48), 48); return true; }
  if (id == pbh::cc4("SFTN")) { out = normalized ? std::round(-100.0f + n * 200.0f) : (float)std::min(std::max((int32_t)v, -100), 100); return true; }
  if (id == pbh::cc4("SVOL")) {
    const pbh::ParamDesc d{id, 0.000001f, 15.848932f, 3, -60.0f, 24.0f};
    out = normalized ? pbh::denormalize(d, n) : cl(v, d.min, d.max);
    return true;
  }
  if (id == pbh::cc4("SPAN")) { out = normalized ? -1.0f + n * 2.0f : cl(v, -1.0f, 1.0f); return true; }
  if (!has_env) return false;
  if (id == pbh::cc4("AATK") || id == pbh::cc4("AHLD") || id == pbh::cc4("ADCY") || id == pbh::cc4("AREL")) {
    out = normalized ? 0.0f + std::pow(n, 2.0f) * 10.0f : cl(v, 0.0f, 10.0f);
    return true;
  }
  if (id == pbh::cc4("ASTN")) { out = normalized ? n : cl(v, 0.0f, 1.0f); return true; }
  return false;
}

// Sampler::process_parameter_update on the host mirror of a sampler (sampler.rs:1069-1135, AhdsrParameters setters
// ahdsr.rs:160-249). Returns which state of the sounding voices the device has to re-derive.
uint32_t apply_sampler_param(HostGroup& g, uint32_t sr, uint32_t id, float plain) {
  const float F32_MAX = 3.402823466e+38f;
  GroupParams& gp = g.gp;
  if (id == pbh::cc4("STRN")) { g.base_transpose = (int32_t)plain; return PARAM_PITCH; }
  if (id == pbh::cc4("SFTN")) { g.base_finetune = (int32_t)plain; return PARAM_PITCH; }
  if (id == pbh::cc4("SVOL")) { gp.base_volume = plain; return PARAM_VOLUME; }
  if (id == pbh::cc4("SPAN")) { gp.base_panning = plain; return PARAM_PANNING; }
  const uint64_t nanos = secs_f32_to_nanos(plain);
  if (id == pbh::cc4("AATK")) {
    g.env.attack_nanos = nanos;
    const float t = nanos_as_secs_f32(nanos);
    gp.attack_rate = t == 0.0f ? F32_MAX : 1.0f / (t * (float)sr);
  } else if (id == pbh::cc4("AHLD")) {
    g.env.hold_nanos = nanos;
    gp.hold_is_zero = nanos == 0; gp.hold_samples = nanos_as_secs_f32(nanos) * (float)sr;
  } else if (id == pbh::cc4("ADCY")) {
    g.env.decay_nanos = nanos;
    gp.decay_is_zero = nanos == 0;
    gp.decay_rate = nanos == 0 ? F32_MAX : (1.0f - gp.sustain_level) / (nanos_as_secs_f32(nanos) * (float)sr);
  } else if (id == pbh::cc4("ASTN")) {
    g.env.sustain_level = plain; gp.sustain_level = plain;  // (the decay rate keeps the sustain level it was set with)
  } else if (id == pbh::cc4("AREL")) {
    g.env.release_nanos = nanos;
    const float t = nanos_as_secs_f32(nanos);
    gp.release_is_zero = nanos == 0; gp.release_rate = t == 0.0f ? F32_MAX : 1.0f / (t * (float)sr);
  }
  return PARAM_ENVELOPE;
}
double sampler_pitch_factor(const HostGroup& g) {  // voice.rs:144-148
  return std::pow(2.0, (double)g.base_transpose / 12.0 + (double)g.base_finetune / 1200.0);
}

int sync_state_to_host(pb200_renderer* r) {
  if (r->host_state_valid) return PB200_OK;
  r->h_voices.resize(r->dev_n_voices);
  r->h_gstate.resize(r->dev_n_groups);
  r->h_mstate.resize(r->dev_n_mixers);
  r->h_fx.resize(r->dev_n_fx);
  r->h_fx_state.resize(r->dev_fx_state_bytes);
  CUDA_TRY(r->d_voices.download(r->h_voices, r->sm));
  if (r->n_hq) { r->h_hq.resize(r->dev_n_voices); CUDA_TRY(r->d_hq.download(r->h_hq, r->sm)); }
  if (r->dev_n_gran_rows) { r->h_gran.resize(r->dev_n_gran_rows); CUDA_TRY(r->d_gran_states.download(r->h_gran, r->sm)); }
  CUDA_TRY(r->d_gstate.download(r->h_gstate, r->sm));
  CUDA_TRY(r->d_mstate.download(r->h_mstate, r->sm));
  CUDA_TRY(r->d_fx.download(r->h_fx, r->sm));
  CUDA_TRY(r->d_fx_state.download(r->h_fx_state, r->sm));
  if (r->master_uploaded) CUDA_TRY(cudaMemcpyAsync(&r->h_master, r->d_master.p, sizeof(ExpSm), cudaMemcpyDeviceToHost, r->sm));
  CUDA_TRY(cudaStreamSynchronize(r->sm));
  r->host_state_valid = true;
  return PB200_OK;
}


// rubato ^0.16 make_sincs(256, 128, f_cutoff, BlackmanHarris2) in f32 (rubato src/sinc.rs, src/windows.rs; not
// vendored with the reference -> PARITY UNPINNED): the filter table SincFixedIn::new builds for
// phonic's RubatoResampler parameters (src/utils/resampler/rubato.rs:28-34). Returns the table index.
uint32_t sinc_table_for(pb200_renderer* r, double resample_ratio) {
  const float base_cutoff = 0.95f;
  const float fc = resample_ratio >= 1.0 ? base_cutoff : base_cutoff * (float)resample_ratio;  // make_interpolator
  uint32_t key;
  std::memcpy(&key, &fc, 4);
  for (size_t i = 0; i < r->sinc_table_keys.size(); ++i) if (r->sinc_table_keys[i] == key) return (uint32_t)i;
  const size_t npoints = HQ_CHUNK, factor = HQ_FACTOR, tot = npoints * factor;
  const float PI = 3.14159265358979323846264338327950288f;
  const float pi2 = 2.0f * PI, pi4 = 4.0f * PI, pi6 = 6.0f * PI, np_f = (float)tot;
  std::vector<float> y(tot);
  float sum = 0.0f;
  for (size_t x = 0; x < tot; ++x) {
    const float xf = (float)x;
    float w = 0.35875f - 0.48829f * std::cos(pi2 * xf / np_f) + 0.14128f * std::cos(pi4 * xf / np_f) - 0.01168f * std::cos(pi6 * xf / np_f);
    w = w * w;  // BlackmanHarris2
    const float v = (xf - (float)(tot / 2)) * fc / (float)factor;
    const float sc = v == 0.0f ? 1.0f : std::sin(v * PI) / (v * PI);
    const float val = w * sc;
    sum += val;
    y[x] = val;
  }
  sum /= (float)factor;
  const size_t off = r->sinc_tables.size();
  r->sinc_tables.resize(off + tot);
  for (size_t p = 0; p < npoints; ++p)
    for (size_t n = 0; n < factor; ++n) r->sinc_tables[off + (factor - n - 1) * npoints + p] = y[factor * p + n] / sum;
  r->sinc_table_keys.push_back(key);
  r->sinc_tables_dirty = true;
  return (uint32_t)(r->sinc_table_keys.size() - 1);
}

// GrainWindow::new (src/generator/sampler/granular.rs:110-196): the eight 2048-point window LUTs, f32 as the reference
void build_grain_luts(std::vector<float>& luts) {
  const size_t N = GRAIN_LUT_N;
  luts.assign(8 * N, 0.0f);
  const float PI = 3.14159265358979323846264338327950288f;
  for (size_t i = 0; i < N; ++i) {
    const float phase = (float)i / (float)N;
    luts[0 * N + i] = 0.5f * (1.0f - std::cos(2.0f * PI * phase));
    const float pi_phase = PI * phase;
    luts[1 * N + i] = 0.42f - 0.5f * std::cos(2.0f * pi_phase) + 0.08f * std::cos(4.0f * pi_phase);
    luts[2 * N + i] = phase < 0.5f ? 2.0f * phase : 2.0f * (1.0f - phase);
    const float alpha = 0.5f, width = alpha / 2.0f;
    if (phase < width) { const float u = phase / width; luts[3 * N + i] = 0.5f * (1.0f - std::cos(PI * u)); }
    else if (phase > 1.0f - width) { const float u = (1.0f - phase) / width; luts[3 * N + i] = 0.5f * (1.0f - std::cos(PI * u)); }
    else luts[3 * N + i] = 1.0f;
    const float ramp_width = 0.1f;
    if (phase < ramp_width) luts[4 * N + i] = phase / ramp_width;
    else if (phase > 1.0f - ramp_width) luts[4 * N + i] = (1.0f - phase) / ramp_width;
    else luts[4 * N + i] = 1.0f;
    const float decay_rate = 6.0f;
    luts[5 * N + i] = std::exp(-decay_rate * std::fabs(phase - 0.5f));
    if (phase < 0.9f) luts[6 * N + i] = phase / 0.9f;
    else { const float u = (phase - 0.9f) / 0.1f; luts[6 * N + i] = 0.5f * (1.0f + std::cos(PI * u)); }
    if (phase < 0.1f) { const float u = phase / 0.1f; luts[7 * N + i] = 0.5f * (1.0f - std::cos(PI * u)); }
    else luts[7 * N + i] = 1.0f - ((phase - 0.1f) / 0.9f);
  }
}

__global__ void downmix_kernel(const float* __restrict__ stereo, float* __restrict__ mono, uint32_t frames, uint32_t src_channels) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= frames) return;
  // `frame.iter().sum::<f32>() / channel_count` (sampler.rs:939-941); a mono source was duplicated by the mapping
  const float l = stereo[2 * i], rr = stereo[2 * i + 1];
  mono[i] = src_channels == 1 ? (0.0f + l) / 1.0f : ((0.0f + l) + rr) / 2.0f;
}

// Sampler::create_granular_sample_buffer (src/generator/sampler.rs:908-952): the file as a mono buffer at the output
// rate. Like the reference this plays the file through a temporary PreloadedFileSource (Default quality, repeat 0)
// in 1024-frame calls -- here a temporary renderer on the same device that borrows the sample data -- and mixes
// the result down. Returns the index of the (internal) buffer holding it.
int make_granular_buffer(pb200_renderer* r, uint32_t buffer_id, uint32_t* out_index) {
  const DevBuffer src = r->buffers[buffer_id].dev;
  const uint32_t sr = r->cfg.sample_rate;
  if (src.channels == 1 && src.sample_rate == sr) { *out_index = buffer_id; return PB200_OK; }  // "just copy": shared read-only
  if (cudaSetDevice(r->device) != cudaSuccess) return fail(r, PB200_ERR_CUDA, "granular sample buffer: cudaSetDevice failed");
  pb200_config cfg = r->cfg;
  cfg.master_volume = 1.0f;
  cfg.block_frames = 1024;          // (the resampling render below is sized in 1024-frame blocks whatever the caller's block size is)
  cfg.device_ordinal = r->device;   // (the borrowed sample pointer and the pool blocks below belong to THIS renderer's device)
  pb200_renderer* t = nullptr;
  if (int e = pb200_create(&cfg, &t)) return fail(r, e, "granular sample buffer: could not create the resampling renderer");
  HostBuffer borrowed;
  borrowed.dev = src; borrowed.dev.loop_start = -1; borrowed.dev.loop_end = -1; borrowed.cls = 0;
  t->buffers.push_back(borrowed);
  pb200_file_options fo;
  pb200_file_options_default(&fo);
  fo.repeat = 0;
  uint32_t pid = 0;
  int e = pb200_play_file(t, 0, &fo, PB200_TIME_NOW, &pid);
  const uint64_t in_frames = src.n_samples / src.channels;
  const uint64_t want = in_frames * sr / src.sample_rate + 100 + 2 * 1024;
  const uint64_t frames = (want + 1023) / 1024 * 1024;
  float* tmp = nullptr;
  size_t tmp_cls = 0;
  if (!e && DevicePool::get().alloc((void**)&tmp, frames * 2 * sizeof(float), &tmp_cls) != cudaSuccess) e = PB200_ERR_CUDA;
  uint64_t written = 0;
  if (!e) e = pb200_render_device(t, tmp, frames, &written);
  uint64_t produced = 0;
  if (!e) produced = t->stats.voice_frames;  // sum of what the source's write calls returned
  std::string msg = e ? t->last_error : std::string();
  pb200_destroy(t);
  if (e) { if (tmp) DevicePool::get().release(tmp, tmp_cls); return fail(r, e, "granular sample buffer: " + msg); }
  const uint64_t n = std::max<uint64_t>(produced, 1);
  HostBuffer hb;
  std::memset(&hb, 0, sizeof(hb));
  float* mono = nullptr;
  if (DevicePool::get().alloc((void**)&mono, n * sizeof(float), &hb.cls) != cudaSuccess) {
    DevicePool::get().release(tmp, tmp_cls);
    return fail(r, PB200_ERR_CUDA, "granular sample buffer: out of device memory");
  }
  cudaError_t ce;
  if (produced == 0) ce = cudaMemsetAsync(mono, 0, sizeof(float), r->sm);  // "ensure sample buffer is not empty"
  else { downmix_kernel<<<(uint32_t)((produced + 255) / 256), 256, 0, r->sm>>>(tmp, mono, (uint32_t)produced, src.channels); ce = cudaGetLastError(); }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(r->sm);
  DevicePool::get().release(tmp, tmp_cls);
  if (ce != cudaSuccess) {
    DevicePool::get().release(mono, hb.cls);
    return fail(r, PB200_ERR_CUDA, std::string("granular sample buffer: ") + cudaGetErrorString(ce));
  }
  hb.dev.data = mono; hb.dev.n_samples = (uint32_t)n; hb.dev.channels = 1; hb.dev.sample_rate = sr;
  hb.dev.loop_start = -1; hb.dev.loop_end = -1;
  r->buffers.push_back(hb);
  *out_index = (uint32_t)r->buffers.size() - 1;
  return PB200_OK;
}

HqState default_hq() {
  HqState h;
  std::memset(&h, 0, sizeof(h));
  h.rec = HQ_NONE;
  return h;
}
}  // namespace

extern "C" {

const char* pb200_backend(void) { return "cuda sm_100a: hand-written voice/mixer/effect kernels (no CPU fallback)"; }

int pb200_create(const pb200_config* config, pb200_renderer** out) {
  if (!config || !out) return PB200_ERR_PARAMETER;
  if (config->channel_count != 2 || config->sample_rate == 0) return PB200_ERR_PARAMETER;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    fprintf(stderr, "phonic_b200: no CUDA device available -- this renderer has no CPU fallback\n");
    return PB200_ERR_CUDA;
  }
  auto* r = new pb200_renderer();
  r->cfg = *config;
  if (r->cfg.block_frames == 0) r->cfg.block_frames = 1024;
  if (r->cfg.block_frames > CHUNK_MAX) { delete r; return PB200_ERR_UNSUPPORTED; }  // chunks are staged in shared memory
  if (config->device_ordinal >= 0) {
    if (config->device_ordinal >= n || cudaSetDevice(config->device_ordinal) != cudaSuccess) { delete r; return PB200_ERR_CUDA; }
    r->device = config->device_ordinal;
  } else {
    cudaGetDevice(&r->device);
  }
  // The mixer pass is the end of every time block's dependency chain and its CTAs are few and long-running: they go first
  // when SMs free up (also against the kernels of another renderer on the same device, e.g. the shard of a sharded render
  // next to rank 0's main-bus stage).
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);   // (numerically lower = higher priority)
  static const bool flat_prio = getenv("PB200_FLAT_PRIORITIES") != nullptr;
  // The serial skeleton pass comes next, the wide replay launches last (best whole-render time). PB200_PRIO_ORDER=replay
  // swaps the two: a time block that has been walked is finished before the walk of later blocks takes the SMs -- the first
  // pieces of a sharded render leave earlier (cfg5 rank with 106 sub-mixers: 30 -> 22 ms) but the render ends later (62 -> 72).
  static const bool skel_first = !(getenv("PB200_PRIO_ORDER") && !strcmp(getenv("PB200_PRIO_ORDER"), "replay"));
  const int p_mid = std::min(prio_lo, prio_hi + 1);
  const int p_mix = flat_prio ? prio_lo : prio_hi, p_skel = flat_prio || !skel_first ? prio_lo : p_mid, p_rep = flat_prio || skel_first ? prio_lo : p_mid;
  if (DevicePool::get().stream(&r->sv, r->device, p_skel) != cudaSuccess ||
      DevicePool::get().stream(&r->sr_, r->device, p_rep) != cudaSuccess ||
      DevicePool::get().stream(&r->sr2, r->device, p_rep) != cudaSuccess ||
      DevicePool::get().stream(&r->sm, r->device, p_mix) != cudaSuccess) { delete r; return PB200_ERR_CUDA; }
  r->prio[0] = p_skel; r->prio[1] = p_rep; r->prio[2] = p_mix;
  r->progress = ProgressWords::get().take();
  if (!r->progress) { delete r; return PB200_ERR_CUDA; }
  *r->progress = 0;
  {  // the note -> speed table is the same for every renderer: uploaded once per device
    static std::mutex mu;
    static std::vector<int> done;
    std::lock_guard<std::mutex> g(mu);
    if (std::find(done.begin(), done.end(), r->device) == done.end()) {
      double note_speed[128];
      for (uint32_t n = 0; n < 128; ++n) note_speed[n] = speed_from_note_h(n);
      if (cudaMemcpyToSymbol(c_note_speed, note_speed, sizeof(note_speed)) != cudaSuccess) { delete r; return PB200_ERR_CUDA; }
      done.push_back(r->device);
    }
  }
  r->rc.sample_rate = config->sample_rate;
  r->rc.rate_comp = 44100.0f / (float)config->sample_rate;
  HostMixer main;
  main.public_id = 0; main.parent = 0xFFFFFFFFu; main.depth = 0;
  r->mixers.push_back(main);
  r->mixer_by_id[0] = 0;
  r->h_mstate.push_back(MixerState{0, 1, 0});
  r->h_master = ExpSm{config->master_volume, config->master_volume};
  if (const char* tb = getenv("PB200_TIME_BLOCK")) {
    long v = atol(tb);
    if (v >= (long)r->cfg.block_frames) r->time_block = (uint32_t)(v / r->cfg.block_frames) * r->cfg.block_frames;
  }
  r->time_block = std::max(r->time_block / r->cfg.block_frames, 1u) * r->cfg.block_frames;
  *out = r;
  return PB200_OK;
}

void pb200_destroy(pb200_renderer* r) {
  if (!r) return;
  cudaSetDevice(r->device);
  r->spans.reset();
  if (r->sv) cudaStreamSynchronize(r->sv);
  if (r->sr_) cudaStreamSynchronize(r->sr_);
  if (r->sr2) cudaStreamSynchronize(r->sr2);
  if (r->sm) cudaStreamSynchronize(r->sm);
  for (auto& b : r->buffers) if (b.cls) DevicePool::get().release((void*)b.dev.data, b.cls);  // cls 0: borrowed
  r->d_quiet_block.free(); r->d_auton.free(); r->d_stage_begin.free(); r->d_fx_progress.free(); r->d_fx_pflags.free(); r->d_status.free(); r->d_status_count.free(); r->d_meter.free(); r->d_block_done.free(); r->d_phase_tabs.free(); r->d_phase_dir.free(); r->d_phase_new.free(); r->d_hq_frames.free(); r->d_hq.free(); r->d_sinc_tables.free(); r->d_hq_scratch.free(); r->d_hq_recs.free(); r->d_hq_nrecs.free();
  r->d_gran_groups.free(); r->d_gran_states.free(); r->d_grain_recs.free(); r->d_gran_counters.free(); r->d_gran_vrec.free();
  r->d_gran_tiles.free(); r->d_grain_storage.free(); r->d_grain_carry[0].free(); r->d_grain_carry[1].free(); r->d_grain_luts.free();
  r->d_buffers.free(); r->d_voices.free(); r->d_groups.free(); r->d_gstate.free(); r->d_events.free();
  r->d_mixers.free(); r->d_mstate.free(); r->d_child_index.free(); r->d_source_index.free(); r->d_level_mixers.free();
  r->d_class_groups.free(); r->d_fx.free(); r->d_fx_events.free(); r->d_fx_state.free(); r->d_aux.free();
  r->d_bounds.free(); r->d_chunk_begin.free(); r->d_group_bus.free(); r->d_mixer_bus.free(); r->d_out.free();
  r->d_group_flags.free(); r->d_mixer_flags.free(); r->d_master.free();
  r->d_segs.free(); r->d_gsegs.free(); r->d_seg_first.free(); r->d_seg_count.free(); r->d_gseg_first.free(); r->d_gseg_count.free(); r->d_recs.free();
  if (r->sr_) DevicePool::get().release_stream(r->sr_, r->device, r->prio[1]);
  if (r->sr2) DevicePool::get().release_stream(r->sr2, r->device, r->prio[1]);
  if (r->sv) DevicePool::get().release_stream(r->sv, r->device, r->prio[0]);
  if (r->sm) DevicePool::get().release_stream(r->sm, r->device, r->prio[2]);
  if (r->sc) { cudaStreamSynchronize(r->sc); cudaStreamDestroy(r->sc); }
  if (r->push_flags) cudaFreeHost(r->push_flags);
  ProgressWords::get().give(r->progress);
  delete r;
}

const char* pb200_last_error(const pb200_renderer* r) { return r ? r->last_error.c_str() : ""; }

int pb200_upload_buffer(pb200_renderer* r, const float* data, uint64_t frames, uint32_t ch, uint32_t rate,
                        int64_t loop_start, int64_t loop_end, int add_pad_frame, uint32_t* buffer_id) {
  if (!r || !data || !buffer_id) return PB200_ERR_PARAMETER;
  if (rate == 0) return fail(r, PB200_ERR_PARAMETER, "file buffer sample rate must be > 0");
  if (ch == 0) return fail(r, PB200_ERR_PARAMETER, "file buffer channel count must be > 0");
  if (frames == 0) return fail(r, PB200_ERR_PARAMETER, "file buffer must not be empty");
  if (ch > 2) return fail(r, PB200_ERR_UNSUPPORTED, "only mono and stereo buffers are supported");
  uint64_t total_frames = frames + (add_pad_frame ? 1 : 0);
  if (total_frames * ch >= 0xFFFFFFF0ull) return fail(r, PB200_ERR_UNSUPPORTED, "buffer too large");
  HostBuffer hb;
  std::memset(&hb, 0, sizeof(hb));
  hb.dev.loop_start = -1; hb.dev.loop_end = -1;
  if (loop_start >= 0 && loop_end >= 0) {
    if (loop_start >= loop_end || (uint64_t)loop_end > total_frames) return fail(r, PB200_ERR_PARAMETER, "file buffer loop range is out of bounds");
    hb.dev.loop_start = (int32_t)loop_start; hb.dev.loop_end = (int32_t)loop_end;
  }
  cudaSetDevice(r->device);
  float* d = nullptr;
  CUDA_TRY(DevicePool::get().alloc((void**)&d, total_frames * ch * sizeof(float), &hb.cls));
  CUDA_TRY(cudaMemsetAsync(d, 0, total_frames * ch * sizeof(float), r->sm));
  CUDA_TRY(cudaMemcpyAsync(d, data, frames * ch * sizeof(float), cudaMemcpyHostToDevice, r->sm));
  CUDA_TRY(cudaStreamSynchronize(r->sm));
  hb.dev.data = d;
  hb.dev.n_samples = (uint32_t)(total_frames * ch);
  hb.dev.channels = ch;
  hb.dev.sample_rate = rate;
  r->buffers.push_back(hb);
  r->graph_dirty = true;
  *buffer_id = (uint32_t)r->buffers.size() - 1;
  return PB200_OK;
}

int pb200_add_mixer(pb200_renderer* r, uint32_t parent, uint32_t* mixer_id) {
  if (!r || !mixer_id) return PB200_ERR_PARAMETER;
  auto it = r->mixer_by_id.find(parent);
  if (it == r->mixer_by_id.end()) return fail(r, PB200_ERR_MIXER_NOT_FOUND, "Mixer not found");
  if (int e = sync_state_to_host(r)) return e;
  HostMixer m;
  m.public_id = r->next_mixer_id++;
  m.parent = it->second;
  m.depth = r->mixers[it->second].depth + 1;
  uint32_t dense = (uint32_t)r->mixers.size();
  r->mixers.push_back(m);
  r->mixers[it->second].children.push_back(dense);
  r->mixer_by_id[m.public_id] = dense;
  r->h_mstate.push_back(MixerState{0, 1, 0});
  r->graph_dirty = true;
  *mixer_id = m.public_id;
  return PB200_OK;
}

int pb200_add_effect(pb200_renderer* r, uint32_t mixer, uint32_t kind, const void* params, size_t size, uint32_t* effect_id) {
  if (!r || !effect_id) return PB200_ERR_PARAMETER;
  auto it = r->mixer_by_id.find(mixer);
  if (it == r->mixer_by_id.end()) return fail(r, PB200_ERR_MIXER_NOT_FOUND, "Mixer not found");
  const uint32_t sr = r->cfg.sample_rate;
  pbh::FxBuild b;
  switch (kind) {
    case PB200_FX_FILTER:
      if (params && size != sizeof(pb200_filter_params)) return fail(r, PB200_ERR_PARAMETER, "bad filter params size");
      b = pbh::build_filter((const pb200_filter_params*)params, sr);
      break;
    case PB200_FX_EQ5:
      if (params) return fail(r, PB200_ERR_PARAMETER, "Eq5Effect has no parameter constructor");
      b = pbh::build_eq5(sr);
      break;
    case PB200_FX_COMPRESSOR:
      if (params && size != sizeof(pb200_compressor_params)) return fail(r, PB200_ERR_PARAMETER, "bad compressor params size");
      b = pbh::build_compressor((const pb200_compressor_params*)params, sr);
      break;
    case PB200_FX_CHORUS:
      if (params && size != sizeof(pb200_chorus_params)) return fail(r, PB200_ERR_PARAMETER, "bad chorus params size");
      b = pbh::build_chorus((const pb200_chorus_params*)params, sr);
      break;
    case PB200_FX_DELAY:
      if (params) return fail(r, PB200_ERR_PARAMETER, "DelayEffect has no parameter constructor");
      b = pbh::build_delay(sr);
      break;
    case PB200_FX_REVERB:
      if (!params || size != sizeof(pb200_reverb_params)) return fail(r, PB200_ERR_PARAMETER, "reverb needs explicit fpd/vib_phase state");
      b = pbh::build_reverb((const pb200_reverb_params*)params, sr);
      break;
    case PB200_FX_GAIN:
      if (params && size != sizeof(pb200_gain_params)) return fail(r, PB200_ERR_PARAMETER, "bad gain params size");
      b = pbh::build_gain((const pb200_gain_params*)params, sr);
      break;
    case PB200_FX_PANNING:
      if (params) return fail(r, PB200_ERR_PARAMETER, "PanningEffect has no parameter constructor");
      b = pbh::build_panning();
      break;
    case PB200_FX_GATE:
      if (params && size != sizeof(pb200_gate_params)) return fail(r, PB200_ERR_PARAMETER, "bad gate params size");
      b = pbh::build_gate((const pb200_gate_params*)params, sr);
      break;
    case PB200_FX_DISTORTION:
      if (params && size != sizeof(pb200_distortion_params)) return fail(r, PB200_ERR_PARAMETER, "bad distortion params size");
      b = pbh::build_distortion((const pb200_distortion_params*)params, sr);
      break;
    default: return fail(r, PB200_ERR_PARAMETER, "unknown effect kind");
  }
  if (b.code) return fail(r, b.code, b.error);
  if (r->mixers[it->second].effects.size() >= (size_t)MAX_EFFECTS_PER_MIXER) return fail(r, PB200_ERR_PARAMETER, "too many effects on one mixer");
  if (int e = sync_state_to_host(r)) return e;
  // device aux arena is append-only: rebase this effect's delay lines behind everything allocated so far
  size_t aux_base = (r->aux_doubles + 1) & ~size_t(1);
  if (aux_base + b.aux_doubles >= 0xFFFFFFF0ull) return fail(r, PB200_ERR_UNSUPPORTED, "effect delay storage exceeds 32 GiB");
  pbh::rebase_aux(kind, b.state, (uint32_t)aux_base);
  r->aux_doubles = aux_base + b.aux_doubles;
  FxHeader h;
  std::memset(&h, 0, sizeof(h));
  h.kind = kind;
  h.bypassed = 1;                     // EffectProcessor::new (mixed/effect.rs:23-30)
  h.tail_counter = 0;
  h.silence_counter = UINT64_MAX;
  size_t off = (r->h_fx_state.size() + 15) & ~size_t(15);
  r->h_fx_state.resize(off + b.state.size());
  std::memcpy(r->h_fx_state.data() + off, b.state.data(), b.state.size());
  h.state_offset = (uint32_t)off;
  h.aux_offset = (uint32_t)aux_base;
  HostFx fx;
  fx.public_id = r->next_effect_id++;
  fx.kind = kind;
  fx.mixer = it->second;
  uint32_t dense = (uint32_t)r->fxs.size();
  r->fxs.push_back(fx);
  r->h_fx.push_back(h);
  r->mixers[it->second].effects.push_back(dense);
  r->h_mstate[it->second].effects_bypassed = 0;  // AddEffect (mixed.rs:430-431)
  r->fx_by_id[fx.public_id] = dense;
  r->graph_dirty = true;
  *effect_id = fx.public_id;
  return PB200_OK;
}

void pb200_file_options_default(pb200_file_options* o) {
  if (!o) return;
  o->volume = 1.0f; o->panning = 0.0f; o->speed = 1.0; o->repeat = PB200_REPEAT_DEFAULT;
  o->loop_start = PB200_NO_LOOP; o->loop_end = PB200_NO_LOOP;
  o->fade_in_nanos = PB200_DURATION_NONE; o->fade_out_nanos = 50000000ull;
  o->resampling_quality = 0; o->target_mixer = PB200_MAIN_MIXER;
}

static int validate_vol_pan(pb200_renderer* r, float volume, float panning) {
  if (volume < 0.0f || std::isnan(volume)) return fail(r, PB200_ERR_PARAMETER, "playback options 'volume' value is invalid");
  if (!(panning >= -1.0f && panning <= 1.0f)) return fail(r, PB200_ERR_PARAMETER, "playback options 'panning' value is invalid");
  return PB200_OK;
}

// insert a source into the mixer's playing_sources keeping the reference's order
// (partition_point(start_time < t): new sources go *before* equal start times, mixed.rs:324-340)
static void insert_source(pb200_renderer* r, uint32_t mixer, uint32_t group, uint64_t start) {
  auto& v = r->mixers[mixer].sources;
  size_t pos = 0;
  while (pos < v.size() && r->groups[v[pos]].gp.start_time < start) ++pos;
  v.insert(v.begin() + pos, group);
}

int pb200_play_file(pb200_renderer* r, uint32_t buffer_id, const pb200_file_options* o, uint64_t start_time, uint32_t* playback_id) {
  if (!r || !o || !playback_id) return PB200_ERR_PARAMETER;
  if (buffer_id >= r->buffers.size()) return fail(r, PB200_ERR_PARAMETER, "unknown buffer");
  if (int e = validate_vol_pan(r, o->volume, o->panning)) return e;
  if (o->speed < 0.0 || std::isnan(o->speed) || std::isinf(o->speed)) return fail(r, PB200_ERR_PARAMETER, "playback options 'speed' value is invalid");
  if (o->resampling_quality > 1) return fail(r, PB200_ERR_PARAMETER, "unknown resampling quality");
  auto mit = r->mixer_by_id.find(o->target_mixer);
  if (mit == r->mixer_by_id.end()) return fail(r, PB200_ERR_MIXER_NOT_FOUND, "Mixer not found");
  if (int e = sync_state_to_host(r)) return e;
  const DevBuffer& b = r->buffers[buffer_id].dev;
  const uint32_t sr = r->cfg.sample_rate;
  HostGroup g;
  std::memset(&g.gp, 0, sizeof(g.gp));
  g.public_id = r->next_source_id++;
  g.gp.kind = GROUP_FILE;
  g.gp.first_voice = (uint32_t)r->h_voices.size();
  g.gp.n_voices = 1;
  g.gp.buffer = buffer_id;
  g.gp.mixer = mit->second;
  g.gp.transient = 1;
  g.gp.start_time = start_time == PB200_TIME_NOW ? 0 : start_time;
  g.gp.has_fade_out = (o->fade_out_nanos != PB200_DURATION_NONE && o->fade_out_nanos != 0) ? 1 : 0;
  if (g.gp.has_fade_out) g.gp.fade_out_inertia = fader_inertia(sr, o->fade_out_nanos);
  g.gp.base_volume = 1.0f;
  VoiceState v = default_voice(b, sr, o->speed);
  // PreloadedFileSource::from_shared_buffer (preloaded.rs:72-115)
  if (o->repeat != PB200_REPEAT_DEFAULT) v.repeat = o->repeat == PB200_REPEAT_FOREVER ? REPEAT_FOREVER : (uint32_t)std::min<uint64_t>(o->repeat, 0xFFFFFFFEull);
  v.repeat_count = v.repeat;
  if (o->loop_start >= 0 && o->loop_end >= 0) {
    uint64_t fc = b.n_samples / b.channels;
    v.loop_ovr_start = (int32_t)std::min<uint64_t>((uint64_t)o->loop_start, fc > 0 ? fc - 1 : 0);
    v.loop_ovr_end = (int32_t)std::min<uint64_t>((uint64_t)o->loop_end, fc);
  }
  if (o->fade_in_nanos != PB200_DURATION_NONE && o->fade_in_nanos != 0) {  // start_fade_in (fader.rs:60-66)
    v.fader_state = FADER_RUNNING; v.fader_cur = 0.0f; v.fader_tgt = 1.0f;
    v.fader_inertia = fader_inertia(sr, o->fade_in_nanos);
  }
  v.vol = ExpSm{o->volume, o->volume};
  v.pan = ExpSm{o->panning, o->panning};
  v.has_note = 1;  // a file playback is one always-active voice
  HqState hq = default_hq();
  if (o->resampling_quality == 1) {
    // RubatoResampler::new (rubato.rs:22-56) on the specs of FileSourceImpl::new (file/common.rs:78-87)
    const uint32_t rate = f64_as_u32_h((double)sr / o->speed);
    if (rate == 0) return fail(r, PB200_ERR_RESAMPLING, "Invalid resampling ratio");
    const double ratio = (double)rate / (double)b.sample_rate;
    if (!(ratio >= 1.0 / 16.0 && ratio <= 64.0)) return fail(r, PB200_ERR_UNSUPPORTED, "HighQuality resampling ratio outside [1/16, 64]");
    v.hq = rate == b.sample_rate ? 2 : 1;
    hq.t_ratio = 1.0 / ratio;
    hq.last_index = -(double)(HQ_CHUNK / 2);
    hq.idx0 = hq.last_index;
    hq.end_idx = (int32_t)HQ_CHUNK - ((int32_t)HQ_CHUNK + 1) - (int32_t)std::ceil(hq.t_ratio);
    hq.slot = r->n_hq++;
    hq.table = v.hq == 1 ? sinc_table_for(r, ratio) : 0u;
  }
  r->h_hq.resize(r->h_voices.size(), default_hq());
  r->h_hq.push_back(hq);
  r->h_voices.push_back(v);
  GroupState gs;
  std::memset(&gs, 0, sizeof(gs));
  gs.vol = ExpSm{1.0f, 1.0f};
  r->h_gstate.push_back(gs);
  uint32_t dense = (uint32_t)r->groups.size();
  r->groups.push_back(g);
  { GranGroup gg; std::memset(&gg, 0, sizeof(gg)); r->gran_groups.push_back(gg); }
  insert_source(r, mit->second, dense, g.gp.start_time);
  r->group_by_id[g.public_id] = dense;
  r->graph_dirty = true;
  *playback_id = g.public_id;
  return PB200_OK;
}

void pb200_sampler_options_default(pb200_sampler_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->volume = 1.0f; o->panning = 0.0f; o->voices = 8; o->target_mixer = PB200_MAIN_MIXER;
  o->ahdsr.attack_nanos = 10000000ull; o->ahdsr.hold_nanos = 1000000000ull; o->ahdsr.decay_nanos = 500000000ull;
  o->ahdsr.release_nanos = 1000000000ull; o->ahdsr.sustain_level = 0.75f;
}

int pb200_add_sampler(pb200_renderer* r, uint32_t buffer_id, const pb200_sampler_options* o, uint64_t start_time, uint32_t* generator_id) {
  if (!r || !o || !generator_id) return PB200_ERR_PARAMETER;
  if (buffer_id >= r->buffers.size()) return fail(r, PB200_ERR_PARAMETER, "unknown buffer");
  if (int e = validate_vol_pan(r, o->volume, o->panning)) return e;
  if (o->voices == 0) return fail(r, PB200_ERR_PARAMETER, "playback options voice count is '0'");
  if (o->voices > (uint32_t)VK_MAX_VOICES) return fail(r, PB200_ERR_UNSUPPORTED, "more than 1024 voices per sampler");
  cudaSetDevice(r->device);
  auto mit = r->mixer_by_id.find(o->target_mixer);
  if (mit == r->mixer_by_id.end()) return fail(r, PB200_ERR_MIXER_NOT_FOUND, "Mixer not found");
  if (int e = sync_state_to_host(r)) return e;
  const DevBuffer& b = r->buffers[buffer_id].dev;
  const uint32_t sr = r->cfg.sample_rate;
  HostGroup g;
  std::memset(&g.gp, 0, sizeof(g.gp));
  g.public_id = r->next_source_id++;
  g.gp.kind = GROUP_SAMPLER;
  g.gp.first_voice = (uint32_t)r->h_voices.size();
  g.gp.n_voices = o->voices;
  g.gp.buffer = buffer_id;
  g.gp.mixer = mit->second;
  g.gp.transient = o->transient ? 1 : 0;
  g.gp.start_time = start_time == PB200_TIME_NOW ? 0 : start_time;
  g.gp.has_fade_out = 1;  // sampler voices: fade_out_duration = 50 ms (sampler.rs:513-514)
  g.gp.fade_out_inertia = fader_inertia(sr, 50000000ull);
  g.gp.base_volume = 1.0f; g.gp.base_panning = 0.0f;
  if (o->has_ahdsr && !resolve_ahdsr(o->ahdsr, sr, g.gp)) return fail(r, PB200_ERR_PARAMETER, "Invalid AHDSR parameters");
  if (o->has_ahdsr) g.env = o->ahdsr;
  GranGroup gg;
  std::memset(&gg, 0, sizeof(gg));
  if (o->has_granular) {  // Sampler::with_granular_playback (sampler.rs:599-637)
    const pb200_granular_params& p = o->granular;
    if (p.overlap_mode > 1 || p.window > 7 || p.playback_direction > 2) return fail(r, PB200_ERR_PARAMETER, "Invalid granular parameters");
    // GranularParameters::validate (granular.rs:287-331)
    if (!(p.size >= 1.0f && p.size <= 1000.0f) || !(p.density >= 1.0f && p.density <= 100.0f) || !(p.spray >= 0.0f && p.spray <= 1.0f) ||
        !(p.variation >= 0.0f && p.variation <= 1.0f) || !(p.pan_spread >= 0.0f && p.pan_spread <= 1.0f) ||
        !(p.position >= 0.0f && p.position <= 1.0f) || !(p.step >= -4.0f && p.step <= 4.0f))
      return fail(r, PB200_ERR_PARAMETER, "Invalid granular parameters");
    if (p.variation != 0.0f || p.spray != 0.0f || p.pan_spread != 0.0f || p.playback_direction == 2)
      return fail(r, PB200_ERR_UNSUPPORTED, "OS-seeded grain randomisation is not reproducible");
    // create_granular_sample_buffer (sampler.rs:908-952) depends on the file buffer and the output rate only: samplers
    // that play the same buffer share one device copy of it instead of resampling it once each
    uint32_t gbuf = 0;
    auto cached = r->gran_buffer_of.find(buffer_id);
    if (cached != r->gran_buffer_of.end()) gbuf = cached->second;
    else {
      if (int e = make_granular_buffer(r, buffer_id, &gbuf)) return e;
      r->gran_buffer_of[buffer_id] = gbuf;
    }
    const DevBuffer& fb = r->buffers[buffer_id].dev;  // (re-read: the buffer vector may have grown)
    gg.enabled = 1;
    gg.overlap_mode = p.overlap_mode; gg.window = p.window; gg.backward = p.playback_direction == 1;
    gg.position = p.position; gg.step = p.step;
    gg.trigger_inc = std::min(std::max(p.density * (1.0f + 0.0f), 1.0f), 100.0f) / (float)sr;  // granular.rs:798-801
    gg.crossfade = p.window <= 3 ? 0.5f : (p.window == 4 ? 0.9f : 0.8f);                       // granular.rs:76-93
    const float grain_size_ms = std::min(std::max(p.size * (1.0f + 0.0f), 1.0f), 1000.0f);
    const float size_f = grain_size_ms * 1.0f * (float)sr / 1000.0f;                            // granular.rs:843-845
    gg.grain_size = std::max<uint32_t>(size_f >= 4294967295.0f ? 0xFFFFFFFFu : (size_f > 0.0f ? (uint32_t)size_f : 0u), 2u);
    gg.buffer = gbuf; gg.buf_len = r->buffers[gbuf].dev.n_samples;
    if (fb.loop_start >= 0) {  // voice.rs:355-361
      const float total = (float)(fb.n_samples / fb.channels);
      gg.has_loop = 1; gg.loop_start = (float)fb.loop_start / total; gg.loop_end = (float)fb.loop_end / total;
    }
    gg.first_row = r->n_gran_rows;
    r->n_gran_rows += o->voices;
    GranState gs0;
    std::memset(&gs0, 0, sizeof(gs0));
    gs0.trigger_new = 1; gs0.speed = 1.0; gs0.volume = 1.0f;
    r->h_gran.resize(r->n_gran_rows, gs0);
    if (r->grain_luts.empty()) { build_grain_luts(r->grain_luts); r->grain_luts_dirty = true; }
  }
  const DevBuffer& b2 = r->buffers[buffer_id].dev;
  for (uint32_t i = 0; i < o->voices; ++i) r->h_voices.push_back(default_voice(b2, sr, 1.0));
  r->h_hq.resize(r->h_voices.size(), default_hq());
  GroupState gs;
  std::memset(&gs, 0, sizeof(gs));
  gs.vol = ExpSm{o->volume, o->volume};
  gs.pan = ExpSm{o->panning, o->panning};
  r->h_gstate.push_back(gs);
  uint32_t dense = (uint32_t)r->groups.size();
  r->groups.push_back(g);
  r->gran_groups.push_back(gg);
  insert_source(r, mit->second, dense, g.gp.start_time);
  r->group_by_id[g.public_id] = dense;
  r->graph_dirty = true;
  *generator_id = g.public_id;
  return PB200_OK;
}

int pb200_schedule(pb200_renderer* r, pb200_event* ev) {
  if (!r || !ev) return PB200_ERR_PARAMETER;
  const bool now = ev->sample_time == PB200_TIME_NOW;
  if (ev->kind == PB200_EV_EFFECT_MESSAGE) {  // EffectHandle::send_message (handles/effect.rs:127-163)
    auto it = r->fx_by_id.find(ev->target);
    if (it == r->fx_by_id.end()) return fail(r, PB200_ERR_EFFECT_NOT_FOUND, "Effect not found");
    HostFx& fx = r->fxs[it->second];
    if (ev->param_id != PB200_MSG_REVERB_RESET || fx.kind != FX_REVERB) return fail(r, PB200_ERR_PARAMETER, "Invalid message for this effect");
    FxParamEvent pe;
    std::memset(&pe, 0, sizeof(pe));
    pe.time = now ? 0 : ev->sample_time;
    pe.param_id = ev->param_id;
    pe.normalized = 2;  // a message, not a parameter update
    fx.events.push_back(pe);
    fx.seqs.push_back(r->next_seq++);
    r->graph_dirty = true;
    return PB200_OK;
  }
  if (ev->kind == PB200_EV_SET_EFFECT_PARAMETER) {
    auto it = r->fx_by_id.find(ev->target);
    if (it == r->fx_by_id.end()) return fail(r, PB200_ERR_EFFECT_NOT_FOUND, "Effect not found");
    // EffectHandle::set_parameter (handles/effect.rs:72-77): a normalized update outside [0, 1] is a ParameterError
    if ((ev->flags & PB200_EVF_NORMALIZED) && !(ev->value >= 0.0f && ev->value <= 1.0f))
      return fail(r, PB200_ERR_PARAMETER, "Invalid parameter update: value should be a normalized value");
    HostFx& fx = r->fxs[it->second];
    const pbh::ParamDesc* desc = nullptr;
    for (auto& d : pbh::param_table(fx.kind)) if (d.id == ev->param_id) desc = &d;
    if (!desc) return fail(r, PB200_ERR_PARAMETER, "Unknown parameter for effect");
    FxParamEvent pe;
    std::memset(&pe, 0, sizeof(pe));
    pe.time = now ? 0 : ev->sample_time;  // handles/effect.rs:80: None => 0
    pe.param_id = ev->param_id;
    pe.value = pbh::resolve_plain(*desc, ev->value, (ev->flags & PB200_EVF_NORMALIZED) != 0);
    if (fx.kind == FX_DELAY && ev->param_id == pbh::cc4("lfos") && pe.value >= 5.0f)
      return fail(r, PB200_ERR_UNSUPPORTED, "OS-seeded random LFO shapes are not reproducible");
    fx.events.push_back(pe);
    fx.seqs.push_back(r->next_seq++);
    r->graph_dirty = true;
    return PB200_OK;
  }
  auto git = r->group_by_id.find(ev->target);
  if (git == r->group_by_id.end()) return fail(r, PB200_ERR_SOURCE_NOT_PLAYING, "Source is no longer playing");
  HostGroup& g = r->groups[git->second];
  const bool is_sampler = g.gp.kind == GROUP_SAMPLER;
  const DevBuffer& b = r->buffers[g.gp.buffer].dev;
  DevEvent de;
  std::memset(&de, 0, sizeof(de));
  de.time = now ? r->position : ev->sample_time;
  de.glide = ev->glide;
  double raw_speed = 0.0;
  uint32_t param_id = 0;
  switch (ev->kind) {
    case PB200_EV_STOP_SOURCE:
      if (!now) {  // MixerMessage::StopSource -> PlayingSource::stop_time (mixed.rs:388-399)
        if (int e = sync_state_to_host(r)) return e;
        r->h_gstate[git->second].has_stop_time = 1;
        r->h_gstate[git->second].stop_time = ev->sample_time;
        r->graph_dirty = true;
        return PB200_OK;
      }
      de.kind = EVK_STOP;
      break;
    case PB200_EV_SET_SOURCE_VOLUME: de.kind = EVK_SET_VOLUME; de.value = ev->value; break;
    case PB200_EV_SET_SOURCE_PANNING: de.kind = EVK_SET_PANNING; de.value = ev->value; break;
    case PB200_EV_SET_SOURCE_SPEED:
      if (is_sampler) return fail(r, PB200_ERR_PARAMETER, "set_speed needs a file source");
      {
        // HighQuality: rubato runs with max_resample_ratio_relative = 1.0 (rubato.rs:37); any other output rate fails
        // set_resample_ratio and FileSourceImpl::update_speed `expect`-panics (file/common.rs:166-168)
        if (int e = sync_state_to_host(r)) return e;
        const VoiceState& fv = r->h_voices[g.gp.first_voice];
        if (fv.hq) {
          const uint32_t new_rate = f64_as_u32_h((double)r->cfg.sample_rate / ev->speed);
          const double t_new = 1.0 / ((double)new_rate / (double)b.sample_rate);
          if (ev->glide > 0.0f || new_rate == 0 || t_new != r->h_hq[g.gp.first_voice].t_ratio)
            return fail(r, PB200_ERR_RESAMPLING, "HighQuality file sources cannot change speed");
        }
      }
      de.kind = EVK_SET_SPEED; de.speed = ev->speed;
      break;
    case PB200_EV_SEEK_SOURCE: {
      if (is_sampler) return fail(r, PB200_ERR_PARAMETER, "seek needs a file source");
      de.kind = EVK_SEEK;
      double pos = nanos_as_secs_f64(ev->position_nanos) * (double)b.sample_rate * (double)b.channels;  // preloaded.rs:140-143
      uint64_t p = !(pos > 0.0) ? 0 : (pos >= 1.8446744073709552e19 ? UINT64_MAX : (uint64_t)pos);
      de.seek_pos = (uint32_t)std::min<uint64_t>(p, b.n_samples);
      break;
    }
    case PB200_EV_NOTE_ON:
      if (!is_sampler) return fail(r, PB200_ERR_GENERATOR_NOT_FOUND, "Generator not found");
      de.kind = EVK_NOTE_ON;
      ev->note_id = r->next_note_id++;
      de.note_id = ev->note_id;
      de.note = ev->note & 0xFF;
      de.value = (ev->flags & PB200_EVF_HAS_VOLUME) ? ev->value : 1.0f;
      de.value2 = (ev->flags & PB200_EVF_HAS_PANNING) ? ev->value2 : 0.0f;
      // voice.rs:144-148: note speed * 2^(transpose/12 + finetune/1200); the factor in force at the event's time is
      // applied when the score is compiled (upload_graph)
      raw_speed = speed_from_note_h(ev->note);
      de.speed = raw_speed;
      break;
    case PB200_EV_NOTE_OFF: de.kind = EVK_NOTE_OFF; de.note_id = ev->note_id; break;
    case PB200_EV_ALL_NOTES_OFF: de.kind = EVK_ALL_NOTES_OFF; break;
    case PB200_EV_SET_NOTE_SPEED:
      de.kind = EVK_NOTE_SPEED; de.note_id = ev->note_id;
      raw_speed = ev->speed;
      de.speed = raw_speed;
      break;
    case PB200_EV_SET_NOTE_VOLUME: de.kind = EVK_NOTE_VOLUME; de.note_id = ev->note_id; de.value = ev->value; break;
    case PB200_EV_SET_NOTE_PANNING: de.kind = EVK_NOTE_PANNING; de.note_id = ev->note_id; de.value = ev->value; break;
    case PB200_EV_SET_GENERATOR_PARAMETER: {
      if (!is_sampler) return fail(r, PB200_ERR_GENERATOR_NOT_FOUND, "Generator not found");
      if ((ev->flags & PB200_EVF_NORMALIZED) && !(ev->value >= 0.0f && ev->value <= 1.0f))
        return fail(r, PB200_ERR_PARAMETER, "Invalid parameter update: value should be a normalized value");
      if (std::isnan(ev->value)) return fail(r, PB200_ERR_PARAMETER, "Invalid parameter value");
      float plain = 0.0f;
      if (!sampler_param_plain(ev->param_id, ev->value, (ev->flags & PB200_EVF_NORMALIZED) != 0, g.gp.has_env != 0, plain))
        return fail(r, PB200_ERR_PARAMETER, "Invalid or unknown sampler parameter");
      de.kind = EVK_SET_PARAM; de.value = plain; param_id = ev->param_id;
      break;
    }
    case PB200_EV_SET_GENERATOR_LOOP_RANGE: {
      if (!is_sampler) return fail(r, PB200_ERR_GENERATOR_NOT_FOUND, "Generator not found");
      if (r->gran_groups[git->second].enabled) return fail(r, PB200_ERR_UNSUPPORTED, "loop range messages to granular samplers");
      de.kind = EVK_SET_LOOP;
      if (ev->flags & PB200_EVF_NO_RANGE) de.flags2 = 2u;
      else {
        const uint64_t fc = b.n_samples / b.channels;  // frame_count() includes the pad frame
        if (ev->position_nanos >= ev->note_id || ev->position_nanos >= fc || ev->note_id > fc) return fail(r, PB200_ERR_PARAMETER, "Invalid loop range");
        de.seek_pos = (uint32_t)std::min<uint64_t>(ev->position_nanos, 0xFFFFFFFFull);
        de.note = (uint32_t)std::min<uint64_t>(ev->note_id, 0xFFFFFFFFull);
      }
      break;
    }
    default: return fail(r, PB200_ERR_PARAMETER, "unknown event kind");
  }
  if (de.kind >= EVK_NOTE_ON && !is_sampler) return fail(r, PB200_ERR_GENERATOR_NOT_FOUND, "Generator not found");
  HostEvent he;
  he.ev = de;
  he.seq = r->next_seq++;
  he.raw_speed = raw_speed;
  he.param_id = param_id;
  // immediate messages do not split mixer chunks; mark them so the schedule compiler skips them
  he.ev.flags = (now ? 1u : 0u) | de.flags2;
  he.ev.flags2 = 0;
  g.events.push_back(he);
  r->graph_dirty = true;
  return PB200_OK;
}

uint64_t pb200_position(const pb200_renderer* r) { return r ? r->position : 0; }

// ---- structural messages: MixedSource::process_messages runs them at the next block start = the next render call ----
static void drop_group(pb200_renderer* r, uint32_t dense) {
  HostGroup& g = r->groups[dense];
  auto& v = r->mixers[g.gp.mixer].sources;
  v.erase(std::remove(v.begin(), v.end(), dense), v.end());
  g.removed = true;
  g.events.clear();
  r->h_gstate[dense].dead = 1;
  r->h_gstate[dense].dead_time = r->position;
  r->h_gstate[dense].ev_cursor = g.gp.ev_begin;
  r->group_by_id.erase(g.public_id);
}

int pb200_remove_source(pb200_renderer* r, uint32_t playback_id) {
  if (!r) return PB200_ERR_PARAMETER;
  auto it = r->group_by_id.find(playback_id);
  if (it == r->group_by_id.end()) return fail(r, PB200_ERR_GENERATOR_NOT_FOUND, "Generator not found");
  if (int e = sync_state_to_host(r)) return e;
  drop_group(r, it->second);
  r->graph_dirty = true;
  return PB200_OK;
}

int pb200_remove_mixer(pb200_renderer* r, uint32_t mixer_id) {
  if (!r) return PB200_ERR_PARAMETER;
  if (mixer_id == PB200_MAIN_MIXER) return fail(r, PB200_ERR_PARAMETER, "Cannot remove the main mixer");
  auto it = r->mixer_by_id.find(mixer_id);
  if (it == r->mixer_by_id.end()) return fail(r, PB200_ERR_MIXER_NOT_FOUND, "Mixer not found");
  if (int e = sync_state_to_host(r)) return e;
  const uint32_t top = it->second;
  auto& sib = r->mixers[r->mixers[top].parent].children;
  sib.erase(std::remove(sib.begin(), sib.end(), top), sib.end());
  std::vector<uint32_t> gone{top};
  for (size_t i = 0; i < gone.size(); ++i)
    for (uint32_t ch : r->mixers[gone[i]].children) gone.push_back(ch);
  for (uint32_t mi : gone) {
    HostMixer& m = r->mixers[mi];
    m.removed = true;
    const std::vector<uint32_t> srcs = m.sources;
    for (uint32_t gdense : srcs) drop_group(r, gdense);
    for (uint32_t f : m.effects) r->fx_by_id.erase(r->fxs[f].public_id);
    m.effects.clear();
    r->mixer_by_id.erase(m.public_id);
  }
  r->graph_dirty = true;
  return PB200_OK;
}

int pb200_remove_effect(pb200_renderer* r, uint32_t effect_id) {
  if (!r) return PB200_ERR_PARAMETER;
  auto it = r->fx_by_id.find(effect_id);
  if (it == r->fx_by_id.end()) return fail(r, PB200_ERR_EFFECT_NOT_FOUND, "Effect not found");
  if (int e = sync_state_to_host(r)) return e;
  const uint32_t dense = it->second, mi = r->fxs[dense].mixer;
  auto& v = r->mixers[mi].effects;
  v.erase(std::remove(v.begin(), v.end(), dense), v.end());
  if (v.empty()) r->h_mstate[mi].effects_bypassed = 1;  // mixed.rs:436-438
  r->fx_by_id.erase(it);
  r->graph_dirty = true;
  return PB200_OK;
}

int pb200_move_effect(pb200_renderer* r, uint32_t effect_id, uint32_t mixer_id, uint32_t movement, int32_t offset) {
  if (!r) return PB200_ERR_PARAMETER;
  auto it = r->fx_by_id.find(effect_id);
  if (it == r->fx_by_id.end()) return fail(r, PB200_ERR_EFFECT_NOT_FOUND, "Effect not found");
  auto mit = r->mixer_by_id.find(mixer_id);
  if (mit == r->mixer_by_id.end() || r->fxs[it->second].mixer != mit->second) return fail(r, PB200_ERR_PARAMETER, "Effect does not belong to this mixer");
  if (movement > PB200_MOVE_END) return fail(r, PB200_ERR_PARAMETER, "unknown effect movement");
  if (int e = sync_state_to_host(r)) return e;
  auto& v = r->mixers[mit->second].effects;
  const size_t cur = std::find(v.begin(), v.end(), it->second) - v.begin();
  const uint32_t dense = v[cur];
  v.erase(v.begin() + cur);
  size_t pos;
  if (movement == PB200_MOVE_DIRECTION) pos = (size_t)std::min<int64_t>(std::max<int64_t>((int64_t)cur + offset, 0), (int64_t)v.size());
  else if (movement == PB200_MOVE_START) pos = 0;
  else pos = v.size();
  v.insert(v.begin() + pos, dense);
  r->graph_dirty = true;
  return PB200_OK;
}

int pb200_stop_all_sources(pb200_renderer* r) {
  if (!r) return PB200_ERR_PARAMETER;
  if (int e = sync_state_to_host(r)) return e;
  const uint64_t now = r->position;
  for (size_t gi = 0; gi < r->groups.size(); ++gi) {
    HostGroup& g = r->groups[gi];
    if (g.removed || r->h_gstate[gi].dead) continue;
    // MixerMessage::RemoveAllPendingEvents (mixed.rs:297-305): pending events of every source ...
    g.events.erase(std::remove_if(g.events.begin(), g.events.end(), [&](const HostEvent& e) { return !(e.ev.flags & 1u) && e.ev.time > now; }), g.events.end());
    if (!g.gp.transient) continue;
    if (g.gp.start_time > now) { drop_group(r, (uint32_t)gi); continue; }  // ... and sources that have not started
    // PlaybackMessageQueue::send_stop to every transient source; the player forgets it
    HostEvent he;
    std::memset(&he.ev, 0, sizeof(he.ev));
    he.ev.time = now; he.ev.kind = EVK_STOP; he.ev.flags = 1u;
    he.seq = r->next_seq++;
    g.events.push_back(he);
    r->group_by_id.erase(g.public_id);
  }
  for (auto& fx : r->fxs) {
    for (size_t i = 0; i < fx.events.size();) {
      if (fx.events[i].time > now) { fx.events.erase(fx.events.begin() + i); fx.seqs.erase(fx.seqs.begin() + i); }
      else ++i;
    }
  }
  r->graph_dirty = true;
  return PB200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// compile + render
// ---------------------------------------------------------------------------------------------------------
namespace {

struct SizeClass { uint32_t vpad; uint32_t threads; std::vector<uint32_t> groups; };

struct Compiled {
  std::vector<std::vector<uint32_t>> levels;  // mixers per depth
  std::vector<SizeClass> classes;
  uint32_t max_chunks = 1;
  std::vector<uint32_t> level_offsets, class_offsets, level_fx_count;
};

// Jump tables for every steady ratio the graph can reach: the speeds of playing voices and of every pending note-on /
// speed event (FileSourceImpl::update_speed's rate arithmetic, file/common.rs:165). Ratios in the middle of a glide are
// not foreseen and take the literal loop. New tables are appended and built by one kernel launch on `s`.
int update_phase_tables(pb200_renderer* r, cudaStream_t s) {
  static const bool off = getenv("PB200_NO_PHASE_TABLES") != nullptr;
  if (off) return PB200_OK;
  const uint32_t sr = r->cfg.sample_rate;
  std::vector<uint2> fresh;
  size_t words = r->phase_words;
  uint64_t last_key = ~0ull;
  auto add = [&](double speed, uint32_t in_rate) {
    uint64_t key;
    std::memcpy(&key, &speed, 8);
    key ^= (uint64_t)in_rate << 1;
    if (key == last_key) return;  // runs of equal speeds are common
    last_key = key;
    if (!(speed > 0.0) || !std::isfinite(speed)) return;
    const uint32_t rate = f64_as_u32_h((double)sr / speed);
    if (rate == 0) return;
    const float ratio = (float)((double)in_rate / (double)rate);
    uint32_t bits;
    std::memcpy(&bits, &ratio, 4);
    if (r->phase_off.count(bits) || r->phase_off.size() >= 8192) return;
    const PhaseGeom g = phase_geom(ratio);
    if (g.mode == PT_LITERAL) return;
    r->phase_off[bits] = (uint32_t)words;
    fresh.push_back(make_uint2(bits, (uint32_t)words));
    words += phase_table_words(g);
  };
  for (size_t gi = 0; gi < r->groups.size(); ++gi) {
    const HostGroup& g = r->groups[gi];
    if (r->h_gstate[gi].dead) continue;
    const uint32_t in_rate = r->buffers[g.gp.buffer].dev.sample_rate;
    for (uint32_t vi = g.gp.first_voice; vi < g.gp.first_voice + g.gp.n_voices; ++vi) {
      if (r->h_voices[vi].hq) continue;
      add(r->h_voices[vi].current_speed, in_rate);
      add(r->h_voices[vi].target_speed, in_rate);
    }
    for (const HostEvent& e : g.events)
      if (e.ev.kind == EVK_NOTE_ON || e.ev.kind == EVK_NOTE_SPEED || e.ev.kind == EVK_SET_SPEED) add(e.ev.speed, in_rate);
  }
  if (!fresh.empty()) {
    if (words >= 0xFFFFFFF0ull) return fail(r, PB200_ERR_UNSUPPORTED, "phase tables exceed 16 GiB");
    CUDA_TRY(r->d_phase_tabs.grow_preserve(words, r->phase_words, s));
    r->phase_words = words;
    CUDA_TRY(r->d_phase_new.upload(fresh, s));
    phase_table_kernel<<<(uint32_t)fresh.size(), 256, 0, s>>>(r->d_phase_tabs.p, r->d_phase_new.p);
    std::vector<uint2> dir;
    dir.reserve(r->phase_off.size());
    for (auto& kv : r->phase_off) dir.push_back(make_uint2(kv.first, kv.second));  // std::map: sorted by the ratio bits
    CUDA_TRY(r->d_phase_dir.upload(dir, s));
  }
  return PB200_OK;
}

int upload_graph(pb200_renderer* r, Compiled& c) {
  // drop consumed events, then flatten per-group / per-effect event lists (stable by time)
  std::vector<DevEvent> events;
  const size_t ng = r->groups.size();
  std::vector<GroupParams> gparams(ng);   // [0, ng): every group's current parameters; behind them the versions that
                                          // pending sampler parameter events switch to (EVK_SET_PARAM.seek_pos)
  const uint32_t sr_out = r->cfg.sample_rate;
  for (size_t gi = 0; gi < ng; ++gi) {
    HostGroup& g = r->groups[gi];
    uint32_t consumed = r->h_gstate[gi].ev_cursor - g.gp.ev_begin;
    if (consumed) {
      const size_t nc = std::min<size_t>(consumed, g.events.size());
      // parameter events the device has applied become the sampler's state (Sampler::process_parameter_update)
      for (size_t i = 0; i < nc; ++i)
        if (g.events[i].ev.kind == EVK_SET_PARAM && !r->h_gstate[gi].stopping) apply_sampler_param(g, sr_out, g.events[i].param_id, g.events[i].ev.value);
      g.events.erase(g.events.begin(), g.events.begin() + nc);
      r->h_gstate[gi].ev_cursor = 0;
    }
    std::stable_sort(g.events.begin(), g.events.end(), [](const HostEvent& a, const HostEvent& b) { return a.ev.time < b.ev.time; });
    g.gp.ev_begin = (uint32_t)events.size();
    g.gp.ev_end = (uint32_t)(events.size() + g.events.size());
    gparams[gi] = g.gp;
    // walk the score: note speeds take the pitch factor in force at their time, parameter events get their version
    HostGroup w;      // working copy of the automatable state
    w.public_id = g.public_id; w.gp = g.gp; w.base_transpose = g.base_transpose; w.base_finetune = g.base_finetune; w.env = g.env;
    double factor = g.gp.kind == GROUP_SAMPLER ? sampler_pitch_factor(w) : 1.0;
    for (auto& e : g.events) {
      if (e.ev.kind == EVK_NOTE_ON || e.ev.kind == EVK_NOTE_SPEED) e.ev.speed = e.raw_speed * factor;
      else if (e.ev.kind == EVK_SET_PARAM) {
        e.ev.note = apply_sampler_param(w, sr_out, e.param_id, e.ev.value);
        factor = sampler_pitch_factor(w);
        e.ev.speed = factor;
        e.ev.seek_pos = (uint32_t)gparams.size();
        gparams.push_back(w.gp);
      }
      events.push_back(e.ev);
    }
    r->h_gstate[gi].gp_idx = (uint32_t)gi;
  }
  // GroupState::ev_cursor is absolute into the flattened array
  for (size_t gi = 0; gi < ng; ++gi) r->h_gstate[gi].ev_cursor = r->groups[gi].gp.ev_begin;
  std::vector<FxParamEvent> fx_events;
  for (size_t fi = 0; fi < r->fxs.size(); ++fi) {
    HostFx& fx = r->fxs[fi];
    FxHeader& h = r->h_fx[fi];
    uint32_t consumed = h.ev_cursor - h.ev_begin;
    if (consumed) { fx.events.erase(fx.events.begin(), fx.events.begin() + std::min<size_t>(consumed, fx.events.size())); }
    std::stable_sort(fx.events.begin(), fx.events.end(), [](const FxParamEvent& a, const FxParamEvent& b) { return a.time < b.time; });
    h.ev_begin = (uint32_t)fx_events.size();
    for (auto& e : fx.events) fx_events.push_back(e);
    h.ev_end = (uint32_t)fx_events.size();
    h.ev_cursor = h.ev_begin;
  }
  // mixers
  std::vector<MixerParams> mparams(r->mixers.size());
  std::vector<uint32_t> child_index, source_index;
  std::vector<FxHeader> fx_sorted;  // FxHeader array must be contiguous per mixer in chain order
  std::vector<uint32_t> fx_perm;    // new position -> old dense index
  uint32_t max_depth = 0;
  for (size_t mi = 0; mi < r->mixers.size(); ++mi) {
    const HostMixer& m = r->mixers[mi];
    MixerParams& p = mparams[mi];
    std::memset(&p, 0, sizeof(p));
    p.parent = m.parent; p.depth = m.depth;
    p.child_begin = (uint32_t)child_index.size();
    for (uint32_t ch : m.children) child_index.push_back(ch);
    p.child_end = (uint32_t)child_index.size();
    p.src_begin = (uint32_t)source_index.size();
    for (uint32_t s : m.sources) source_index.push_back(s);
    p.src_end = (uint32_t)source_index.size();
    p.fx_begin = (uint32_t)fx_perm.size();
    for (uint32_t f : m.effects) fx_perm.push_back(f);
    p.fx_end = (uint32_t)fx_perm.size();
    if (!m.removed) max_depth = std::max(max_depth, m.depth);
  }
  // permute fx headers into chain order (and remember the permutation for later syncs)
  {
    std::vector<FxHeader> nh(fx_perm.size());
    std::vector<HostFx> nf(fx_perm.size());
    std::vector<uint32_t> inv(r->fxs.size(), 0xFFFFFFFFu);  // (removed effects are not listed by any mixer: they drop out here)
    for (size_t i = 0; i < fx_perm.size(); ++i) { nh[i] = r->h_fx[fx_perm[i]]; nf[i] = r->fxs[fx_perm[i]]; inv[fx_perm[i]] = (uint32_t)i; }
    r->h_fx.swap(nh); r->fxs.swap(nf);
    for (auto& kv : r->fx_by_id) kv.second = inv[kv.second];
    for (auto& m : r->mixers) for (auto& f : m.effects) f = inv[f];
  }
  c.levels.assign(max_depth + 1, {});
  for (size_t mi = 0; mi < r->mixers.size(); ++mi) if (!r->mixers[mi].removed) c.levels[r->mixers[mi].depth].push_back((uint32_t)mi);
  std::vector<uint32_t> level_mixers;
  c.level_offsets.clear();
  c.level_fx_count.clear();
  for (auto& l : c.levels) {  // mixers with effects first: the two groups go to different builds of the mixer kernel
    std::stable_partition(l.begin(), l.end(), [&](uint32_t m) { return !r->mixers[m].effects.empty(); });
    c.level_fx_count.push_back((uint32_t)std::count_if(l.begin(), l.end(), [&](uint32_t m) { return !r->mixers[m].effects.empty(); }));
    c.level_offsets.push_back((uint32_t)level_mixers.size());
    for (uint32_t m : l) level_mixers.push_back(m);
  }
  // size classes: groups bucketed by voices-per-group rounded up to a power of two
  c.classes.clear();
  for (uint32_t vpad = 1; vpad <= 1024; vpad <<= 1) {
    SizeClass sc;
    sc.vpad = vpad;
    sc.threads = std::max(32u, vpad);                        // skeleton: one thread per voice
    for (size_t gi = 0; gi < r->groups.size(); ++gi) {
      uint32_t nv = r->groups[gi].gp.n_voices;
      if (nv <= vpad && (vpad == 1 || nv > vpad / 2)) sc.groups.push_back((uint32_t)gi);
    }
    if (sc.groups.empty()) continue;
    c.classes.push_back(sc);
  }
  std::vector<uint32_t> class_groups;
  c.class_offsets.clear();
  for (auto& sc : c.classes) { c.class_offsets.push_back((uint32_t)class_groups.size()); for (uint32_t g : sc.groups) class_groups.push_back(g); }

  std::vector<DevBuffer> bufs;
  for (auto& b : r->buffers) bufs.push_back(b.dev);
  cudaStream_t s = r->sm;
  CUDA_TRY(r->d_buffers.upload(bufs, s));
  CUDA_TRY(r->d_groups.upload(gparams, s));
  CUDA_TRY(r->d_events.upload(events, s));
  CUDA_TRY(r->d_mixers.upload(mparams, s));
  CUDA_TRY(r->d_child_index.upload(child_index, s));
  CUDA_TRY(r->d_source_index.upload(source_index, s));
  CUDA_TRY(r->d_level_mixers.upload(level_mixers, s));
  CUDA_TRY(r->d_class_groups.upload(class_groups, s));
  CUDA_TRY(r->d_fx_events.upload(fx_events, s));
  // state
  CUDA_TRY(r->d_voices.upload(r->h_voices, s));
  if (r->n_hq) {
    r->h_hq.resize(r->h_voices.size(), default_hq());
    CUDA_TRY(r->d_hq.upload(r->h_hq, s));
    if (r->sinc_tables_dirty) { CUDA_TRY(r->d_sinc_tables.upload(r->sinc_tables, s)); r->sinc_tables_dirty = false; }
  }
  if (r->n_gran_rows) {
    CUDA_TRY(r->d_gran_groups.upload(r->gran_groups, s));
    CUDA_TRY(r->d_gran_states.upload(r->h_gran, s));
    if (r->grain_luts_dirty) { CUDA_TRY(r->d_grain_luts.upload(r->grain_luts, s)); r->grain_luts_dirty = false; }
  }
  CUDA_TRY(r->d_gstate.upload(r->h_gstate, s));
  CUDA_TRY(r->d_mstate.upload(r->h_mstate, s));
  CUDA_TRY(r->d_fx.upload(r->h_fx, s));
  CUDA_TRY(r->d_fx_state.upload(r->h_fx_state, s));
  std::vector<ExpSm> master{r->h_master};
  CUDA_TRY(r->d_master.upload(master, s));
  r->master_uploaded = true;
  // aux arena: grow preserving contents, zero the new tail
  if (r->aux_doubles > r->d_aux.cap) {
    double* np = nullptr;
    size_t granted = 0;
    CUDA_TRY(DevicePool::get().alloc((void**)&np, (r->aux_doubles + 1024) * sizeof(double), &granted));
    CUDA_TRY(cudaMemsetAsync(np, 0, granted, s));
    if (r->d_aux.p) {
      CUDA_TRY(cudaMemcpyAsync(np, r->d_aux.p, r->d_aux_used * sizeof(double), cudaMemcpyDeviceToDevice, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      DevicePool::get().release(r->d_aux.p, r->d_aux.cls);
    }
    r->d_aux.p = np; r->d_aux.cap = granted / sizeof(double); r->d_aux.cls = granted;
  }
  r->d_aux_used = r->aux_doubles;
  if (int e = update_phase_tables(r, s)) return e;
  CUDA_TRY(cudaStreamSynchronize(s));
  r->dev_n_gran_rows = r->n_gran_rows;
  r->dev_n_voices = r->h_voices.size(); r->dev_n_groups = r->h_gstate.size(); r->dev_n_mixers = r->h_mstate.size();
  r->dev_n_fx = r->h_fx.size(); r->dev_fx_state_bytes = r->h_fx_state.size();
  r->graph_dirty = false;
  return PB200_OK;
}

// the reference's exact chunk boundaries per mixer for [p0, p1): 1024-frame WavStream blocks, split at
// every pending event time of the mixer itself and of all its ancestors (mixed.rs:679-693)
void compile_schedule(pb200_renderer* r, uint64_t p0, uint64_t p1, uint32_t tb, std::vector<uint64_t>& bounds,
                      std::vector<uint32_t>& begin, uint32_t& n_blocks, uint32_t& max_chunks) {
  const size_t nm = r->mixers.size();
  std::vector<std::vector<uint64_t>> per_mixer(nm);
  std::vector<std::vector<uint64_t>> own(nm);
  for (auto& g : r->groups)
    for (auto& e : g.events)
      if (!(e.ev.flags & 1u) && e.ev.time > p0 && e.ev.time < p1) own[g.gp.mixer].push_back(e.ev.time);
  for (auto& fx : r->fxs)
    for (auto& e : fx.events)
      if (e.time > p0 && e.time < p1) own[fx.mixer].push_back(e.time);
  std::vector<uint32_t> order(nm);
  for (size_t i = 0; i < nm; ++i) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return r->mixers[a].depth < r->mixers[b].depth; });
  for (uint32_t mi : order) {
    std::vector<uint64_t>& v = per_mixer[mi];
    if (r->mixers[mi].parent == 0xFFFFFFFFu) {
      for (uint64_t t = p0; t <= p1; t += r->cfg.block_frames) v.push_back(t);
    } else {
      v = per_mixer[r->mixers[mi].parent];
    }
    v.insert(v.end(), own[mi].begin(), own[mi].end());
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  }
  n_blocks = (uint32_t)((p1 - p0 + tb - 1) / tb);
  bounds.clear();
  begin.assign((size_t)n_blocks * (nm + 1), 0);
  max_chunks = 1;
  std::vector<size_t> cursor(nm, 0);
  for (uint32_t b = 0; b < n_blocks; ++b) {
    const uint64_t b0 = p0 + (uint64_t)b * tb, b1 = std::min<uint64_t>(b0 + tb, p1);
    for (size_t mi = 0; mi < nm; ++mi) {
      begin[(size_t)b * (nm + 1) + mi] = (uint32_t)bounds.size();
      const auto& v = per_mixer[mi];
      size_t& cu = cursor[mi];
      while (cu < v.size() && v[cu] < b0) ++cu;
      size_t first = bounds.size();
      size_t i = cu;
      while (i < v.size() && v[i] <= b1) bounds.push_back(v[i++]);
      max_chunks = std::max<uint32_t>(max_chunks, (uint32_t)(bounds.size() - first));
    }
    begin[(size_t)b * (nm + 1) + nm] = (uint32_t)bounds.size();
  }
}

int render_impl(pb200_renderer* r, float* out_dev, float* out_host, uint64_t frames, uint64_t* frames_written) {
  const uint32_t bf = r->cfg.block_frames;
  if (frames % bf != 0) return fail(r, PB200_ERR_PARAMETER, "frames must be a multiple of block_frames");
  if (frames == 0) { if (frames_written) *frames_written = 0; return PB200_OK; }
  cudaSetDevice(r->device);
  const uint64_t progress_base = r->progress_base;
  struct ProgressDone {   // whatever way the call ends, pollers see it end
    pb200_renderer* r; uint64_t total;
    ~ProgressDone() { r->progress_base = total; __atomic_store_n(r->progress, (unsigned long long)total, __ATOMIC_RELEASE); }
  } progress_done{r, progress_base + frames};
  // WavStream finishes when the main mixer has nothing at all to do (wav.rs:231-234, mixed.rs:664-670)
  size_t live_mixers = 0, live_fx = 0;
  for (auto& m : r->mixers) if (!m.removed) { ++live_mixers; live_fx += m.effects.size(); }
  if (r->groups.empty() && live_fx == 0 && live_mixers == 1 && !r->n_ext) r->finished = true;
  if (r->finished) {
    if (out_host) std::memset(out_host, 0, frames * 2 * sizeof(float));
    if (out_dev) CUDA_TRY(cudaMemsetAsync(out_dev, 0, frames * 2 * sizeof(float), r->sm));
    if (out_dev) CUDA_TRY(cudaStreamSynchronize(r->sm));
    if (frames_written) *frames_written = 0;
    return PB200_OK;
  }
  // PB200_HOST_PROF: wall time of the host-side phases of a render call (debug aid)
  static const bool host_prof = getenv("PB200_HOST_PROF") != nullptr;
  auto hp_now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double hp_t = host_prof ? hp_now() : 0.0;
  auto hp_mark = [&](const char* what) { if (host_prof) { const double t = hp_now(); fprintf(stderr, "[host] %-28s %8.3f ms\n", what, t - hp_t); hp_t = t; } };
  Compiled c;
  if (int e = sync_state_to_host(r)) return e;
  hp_mark("sync_state_to_host");
  // (events and cursors live on the host between render calls: always rebuild the flattened lists)
  if (int e = upload_graph(r, c)) return e;
  hp_mark("upload_graph");

  const uint64_t p0 = r->position, p1 = p0 + frames;
  // Large graphs (every launch already fills the GPU) take longer time blocks: fewer launches and per-block joins
  // (cfg5shard 53.0 -> 50.9 ms); small graphs keep 32768 frames, where the un-overlapped tail of the last block matters.
  uint32_t tb = r->time_block;
  if (!getenv("PB200_TIME_BLOCK") && (r->h_voices.size() >= 2048 || r->n_gran_rows > 0)) tb *= 2;  // granular: cfg4 48 -> 36 ms
  std::vector<uint64_t> bounds;
  std::vector<uint32_t> begin;
  uint32_t n_blocks = 0, max_chunks = 1;
  compile_schedule(r, p0, p1, tb, bounds, begin, n_blocks, max_chunks);
  CUDA_TRY(r->d_bounds.upload(bounds, r->sm));
  CUDA_TRY(r->d_chunk_begin.upload(begin, r->sm));
  const size_t ng = r->groups.size(), nm = r->mixers.size();
  CUDA_TRY(r->d_group_bus.reserve(std::max<size_t>(1, (size_t)RING * ng * tb * 2)));
  CUDA_TRY(r->d_mixer_bus.reserve((size_t)RING * nm * tb * 2));
  CUDA_TRY(r->d_group_flags.reserve(std::max<size_t>(1, (size_t)RING * ng * max_chunks)));
  CUDA_TRY(r->d_mixer_flags.reserve((size_t)RING * nm * max_chunks));
  float* dout = out_dev;
  if (!dout) { CUDA_TRY(r->d_out.reserve(frames * 2)); dout = r->d_out.p; }
  CUDA_TRY(cudaStreamSynchronize(r->sm));

  CUDA_TRY(cudaFuncSetAttribute(mix_fx_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FX_WORK_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(mix_fx_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FX_WORK_SMALL));
  CUDA_TRY(cudaFuncSetAttribute(replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)REPLAY_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(skeleton_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (TAB_SLOT_WORDS * 4 + 12) + 16));
  const uint32_t n_tiles = ((tb + TILE - 1) / TILE + 15u) / 16u * 16u;   // (a block size that is no multiple of 64 ends in a partial tile; rows stay 16-byte aligned)
  const uint32_t seg_cap = n_tiles + max_chunks + 8;
  if (seg_cap >= 65535) return fail(r, PB200_ERR_UNSUPPORTED, "too many chunk boundaries in one time block");
  const size_t nvoices = std::max<size_t>(1, r->h_voices.size());
  // Persistent skeleton (small graphs): ONE launch walks all time blocks, every group at its own pace, and hands
  // block b to the replay pass through a counter the replay stream waits on (cuStreamWaitValue32). Needs every
  // skeleton CTA resident at once, one size class, a snapshot slot per block (no ring re-use, so the kernel never
  // has to wait for a consumer) and no HighQuality / granular voices (their per-block lists are host-managed).
  int sm_count_all = 0;
  cudaDeviceGetAttribute(&sm_count_all, cudaDevAttrMultiProcessorCount, r->device);
  bool persistent = n_blocks > 1 && r->n_hq == 0 && r->n_gran_rows == 0 && c.classes.size() == 1 &&
                    c.classes[0].groups.size() <= (size_t)sm_count_all && !getenv("PB200_NO_PERSISTENT") &&
                    stream_wait_value32() != nullptr;
  if (persistent) {
    const size_t bytes = (size_t)n_blocks * (nvoices * ((size_t)seg_cap * sizeof(Segment) + (size_t)n_tiles * (sizeof(TileRec) + 4)) +
                                             ng * ((size_t)seg_cap * sizeof(GroupSeg) + (size_t)n_tiles * 4 + max_chunks));
    if (bytes > (size_t)4 << 30) persistent = false;
  }
  const uint32_t nslots = persistent ? n_blocks : RING;
  CUDA_TRY(r->d_group_flags.reserve(std::max<size_t>(1, (size_t)nslots * ng * max_chunks)));
  CUDA_TRY(r->d_segs.reserve((size_t)nslots * nvoices * seg_cap));
  CUDA_TRY(r->d_gsegs.reserve(std::max<size_t>(1, (size_t)nslots * ng * seg_cap)));
  CUDA_TRY(r->d_seg_first.reserve((size_t)nslots * nvoices * n_tiles));
  CUDA_TRY(r->d_seg_count.reserve((size_t)nslots * nvoices * n_tiles));
  {  // tile records carry a process-wide unique generation tag; a freshly acquired buffer is cleared once
    const TileRec* before = r->d_recs.p;
    CUDA_TRY(r->d_recs.reserve((size_t)nslots * nvoices * n_tiles));
    if (r->d_recs.p != before) CUDA_TRY(cudaMemsetAsync(r->d_recs.p, 0, r->d_recs.cap * sizeof(TileRec), r->sv));
  }
  CUDA_TRY(r->d_gseg_first.reserve(std::max<size_t>(1, (size_t)nslots * ng * n_tiles)));
  CUDA_TRY(r->d_gseg_count.reserve(std::max<size_t>(1, (size_t)nslots * ng * n_tiles)));
  bool autonomous = false;
  if (persistent) {
    CUDA_TRY(r->d_block_done.reserve(n_blocks));
    CUDA_TRY(cudaMemsetAsync(r->d_block_done.p, 0, n_blocks * sizeof(uint32_t), r->sv));
    // Autonomous voices (SkeletonLoop): per group the first block from which no event couples its voices any more
    static const bool no_auton = getenv("PB200_NO_AUTONOMOUS") != nullptr;
    const SizeClass& sc0 = c.classes[0];
    if (!no_auton && sc0.vpad <= 8) {
      std::vector<uint32_t> quiet(sc0.groups.size(), n_blocks);
      for (size_t ci = 0; ci < sc0.groups.size(); ++ci) {
        const uint32_t gi = sc0.groups[ci];
        const HostGroup& g = r->groups[gi];
        const GroupState& gs = r->h_gstate[gi];
        if (g.gp.kind != GROUP_SAMPLER || gs.dead || gs.stopping || gs.stopped || gs.has_stop_time || g.gp.start_time > p0) continue;
        uint64_t last_hard = 0;
        bool any = false;
        for (auto& e : g.events)
          if (e.ev.kind == EVK_NOTE_ON || e.ev.kind == EVK_STOP || e.ev.kind == EVK_SET_PARAM || e.ev.kind == EVK_SET_LOOP) { any = true; last_hard = std::max(last_hard, e.ev.time); }
        // the first block that starts after the last coupling event (an event AT a block's first frame still belongs to it)
        const uint32_t qb = !any || last_hard < p0 ? 0u : (uint32_t)std::min<uint64_t>((last_hard - p0) / tb + 1, n_blocks);
        quiet[ci] = qb;
        autonomous |= qb < n_blocks;
      }
      if (autonomous) {
        CUDA_TRY(r->d_quiet_block.upload(quiet, r->sv));
        const size_t words = sc0.groups.size() * (size_t)n_blocks * (1 + (size_t)max_chunks);
        CUDA_TRY(r->d_auton.reserve(words));
        CUDA_TRY(cudaMemsetAsync(r->d_auton.p, 0, words * sizeof(uint32_t), r->sv));
      }
    }
  }

  // HighQuality voices: per-block record list + resampler output stream scratch (hq.cuh, sinc_kernel.cuh)
  const uint32_t n_hq = r->n_hq;
  uint32_t hq_cap = 0;
  int sm_count = 148;
  if (n_hq) {
    size_t cap = 0;
    for (size_t vi = 0; vi < r->h_voices.size(); ++vi) {
      if (!r->h_voices[vi].hq) continue;
      // a chunk yields ~256 * ratio frames (>= 1 per process() call on the bypass path), plus one record per
      // write call that re-opens a chunk and per block boundary
      const double per = r->h_voices[vi].hq == 1 ? std::max(1.0, std::floor(256.0 / r->h_hq[vi].t_ratio) - 2.0) : 1.0;
      cap += (size_t)((double)tb / per) + 2 * (size_t)max_chunks + 8;
    }
    if (cap >= 0x7FFFFFFFull) return fail(r, PB200_ERR_UNSUPPORTED, "too many HighQuality chunks in one time block");
    hq_cap = (uint32_t)cap;
    CUDA_TRY(r->d_hq_recs.reserve((size_t)RING * hq_cap));
    CUDA_TRY(r->d_hq_nrecs.reserve(std::max<size_t>(RING, (size_t)n_blocks)));
    CUDA_TRY(r->d_hq_scratch.reserve((size_t)RING * n_hq * tb * 2));
    CUDA_TRY(r->d_hq_frames.reserve(1));
    CUDA_TRY(cudaMemsetAsync(r->d_hq_frames.p, 0, sizeof(unsigned long long), r->sv));
    CUDA_TRY(cudaMemsetAsync(r->d_hq_nrecs.p, 0, std::max<size_t>(RING, (size_t)n_blocks) * sizeof(uint32_t), r->sv));
    CUDA_TRY(cudaFuncSetAttribute(sinc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SINC_SMEM));
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, r->device);
  }

  // granular samplers: per-block grain records, per-voice record lists, tile ranges, contribution storage, carry
  const uint32_t n_rows = r->n_gran_rows;
  uint32_t gran_rec_cap = 0, gran_vrec_cap = 0;
  size_t gran_storage_cap = 0;
  if (n_rows) {
    size_t recs = 0;
    for (size_t gi = 0; gi < r->groups.size(); ++gi) {
      const GranGroup& gg = r->gran_groups[gi];
      if (!gg.enabled) continue;
      const double per_frame = gg.overlap_mode == 1 ? 1.0 / std::max(1.0, std::floor(gg.grain_size * 0.5)) : (double)gg.trigger_inc;
      const size_t per_voice = (size_t)GRAIN_POOL + (size_t)std::ceil((double)tb * per_frame) + 8;
      const size_t overlap = gg.overlap_mode == 1 ? 3 : std::min<size_t>(GRAIN_POOL, (size_t)std::ceil((double)gg.grain_size * gg.trigger_inc) + 2);
      recs += per_voice * r->groups[gi].gp.n_voices;
      gran_vrec_cap = std::max<uint32_t>(gran_vrec_cap, (uint32_t)per_voice);
      gran_storage_cap += (size_t)tb * overlap * r->groups[gi].gp.n_voices;
    }
    if (recs >= 0x7FFFFFFFull || gran_storage_cap >= 0xFFFFFFFFull)
      return fail(r, PB200_ERR_UNSUPPORTED, "granular workload exceeds one time block's grain storage (lower PB200_TIME_BLOCK)");
    gran_rec_cap = (uint32_t)recs;
    CUDA_TRY(r->d_grain_recs.reserve((size_t)RING * gran_rec_cap));
    CUDA_TRY(r->d_gran_counters.reserve(2 * (size_t)std::max<uint32_t>(RING, n_blocks)));
    CUDA_TRY(r->d_gran_vrec.reserve((size_t)RING * n_rows * gran_vrec_cap));
    CUDA_TRY(r->d_gran_tiles.reserve((size_t)RING * n_rows * n_tiles * 2));
    CUDA_TRY(r->d_grain_storage.reserve((size_t)RING * gran_storage_cap));
    // carry of grains in flight: rows of samplers that were already playing keep their state when rows are added
    for (int h = 0; h < 2; ++h) CUDA_TRY(r->d_grain_carry[h].grow_preserve((size_t)n_rows * GRAIN_POOL, r->carry_rows * GRAIN_POOL, r->sv));
    r->carry_rows = n_rows;
    CUDA_TRY(cudaMemsetAsync(r->d_gran_counters.p, 0, 2 * (size_t)std::max<uint32_t>(RING, n_blocks) * sizeof(uint32_t), r->sv));
  }

  // Effect-chain pipelining (mixer_kernel.cuh): per tree level the number of pipeline stages, per mixer the effects of
  // each stage (contiguous ranges of the chain, balanced by the measured cost per chunk of each effect kind).
  std::vector<uint32_t> level_stages(c.levels.size(), 1), stage_offsets(c.levels.size(), 0), level_small(c.levels.size(), 0);
  {
    static const bool no_pipe = getenv("PB200_NO_FX_PIPELINE") != nullptr, no_small = getenv("PB200_NO_FX_SMALL") != nullptr;
    auto fx_cost = [](uint32_t kind) -> double {  // microseconds per 1024-frame chunk (tools/fx_cost.py)
      switch (kind) {
        case FX_FILTER: return 7; case FX_EQ5: return 15; case FX_COMPRESSOR: return 20; case FX_CHORUS: return 22; case FX_DELAY: return 33;
        case FX_REVERB: return 98; case FX_GAIN: return 27; case FX_GATE: return 23; default: return 6;
      }
    };
    std::vector<uint32_t> stage_begin;
    int dev_sms = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, r->device);
    for (size_t lvl = 0; lvl < c.levels.size() && !no_pipe; ++lvl) {
      size_t max_fx = 0;
      bool reverb = false;
      for (uint32_t mi : c.levels[lvl]) {
        max_fx = std::max(max_fx, r->mixers[mi].effects.size());
        for (uint32_t fi : r->mixers[mi].effects) reverb |= r->fxs[fi].kind == FX_REVERB;
      }
      if (max_fx < 2) continue;
      int per_sm = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mix_fx_kernel<1>, FX_THREADS, reverb ? FX_WORK_BYTES : FX_WORK_SMALL);
      const size_t want = std::min<size_t>(max_fx, MAX_FX_STAGES), nlm = std::max<size_t>(1, c.level_fx_count[lvl]);
      uint32_t S = (uint32_t)std::min<size_t>(want, (size_t)std::max(per_sm, 0) * dev_sms / nlm);
      if (S < want && !reverb && !no_small) {   // the 128-register build fits twice as many CTAs
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mix_fx_kernel<2>, FX_THREADS, FX_WORK_SMALL);
        const uint32_t S2 = (uint32_t)std::min<size_t>(want, (size_t)std::max(per_sm, 0) * dev_sms / nlm);
        if (S2 > S) { S = S2; level_small[lvl] = 1; }
      }
      if (S < 2) continue;
      level_stages[lvl] = S;
      stage_offsets[lvl] = 0;  // (indexed by dense mixer index: one table for the whole graph)
    }
    bool any = false;
    for (uint32_t S : level_stages) any |= S > 1;
    if (any) {
      stage_begin.assign(nm * (size_t)(MAX_FX_STAGES + 1), 0);
      for (size_t lvl = 0; lvl < c.levels.size(); ++lvl) {
        const uint32_t S = level_stages[lvl];
        if (S < 2) continue;
        for (uint32_t mi : c.levels[lvl]) {
          const auto& fx = r->mixers[mi].effects;
          const size_t E = fx.size();
          // contiguous partition of the chain into <= S parts with the smallest largest part (E <= 16: brute force DP)
          std::vector<double> pre(E + 1, 0.0);
          for (size_t i = 0; i < E; ++i) pre[i + 1] = pre[i] + fx_cost(r->fxs[fx[i]].kind);
          const uint32_t P = (uint32_t)std::min<size_t>(S, E);   // non-empty parts; the stages after them stay empty
          std::vector<std::vector<double>> best(P + 1, std::vector<double>(E + 1, 1e300));
          std::vector<std::vector<size_t>> cut(P + 1, std::vector<size_t>(E + 1, 0));
          best[0][0] = 0.0;
          for (uint32_t k = 1; k <= P; ++k)
            for (size_t j = k; j <= E; ++j)
              for (size_t i = k - 1; i < j; ++i) {
                const double v = std::max(best[k - 1][i], pre[j] - pre[i]);
                if (v < best[k][j]) { best[k][j] = v; cut[k][j] = i; }
              }
          uint32_t* sb = stage_begin.data() + (size_t)mi * (MAX_FX_STAGES + 1);   // (one fixed-stride row per mixer: levels differ in S)
          for (uint32_t k = P; k <= S; ++k) sb[k] = (uint32_t)E;
          size_t j = E;
          for (uint32_t k = P; k >= 1; --k) { sb[k] = (uint32_t)j; j = cut[k][j]; }
          sb[0] = 0;
        }
      }
      CUDA_TRY(r->d_stage_begin.upload(stage_begin, r->sm));
      CUDA_TRY(r->d_fx_progress.reserve(nm * (size_t)MAX_FX_STAGES + 1));
      CUDA_TRY(r->d_fx_pflags.reserve(nm * (size_t)MAX_FX_STAGES * max_chunks));
      CUDA_TRY(cudaStreamSynchronize(r->sm));
    }
  }
  hp_mark("schedule + work areas");
  r->spans.reset();
  std::unique_ptr<SpanEvents> sp(new SpanEvents);
  SpanEvents& events = *sp;
  events.streams = {r->sv, r->sr_, r->sr2, r->sm, r->sc};
  auto &ev_v0 = sp->v0, &ev_v1 = sp->v1, &ev_r0 = sp->r0, &ev_r1 = sp->r1, &ev_m0 = sp->m0, &ev_m1 = sp->m1, &ev_x = sp->x;
  for (auto* v : {&ev_v0, &ev_v1, &ev_r0, &ev_r1, &ev_m0, &ev_m1}) v->resize(n_blocks);
  ev_x.resize((n_hq || n_rows) ? 3 * (size_t)n_blocks : 0);  // grain begin / grain end = sinc begin / sinc end
  for (auto& e : ev_x) CUDA_TRY(events.get(&e));
  for (uint32_t b = 0; b < n_blocks; ++b) {
    CUDA_TRY(events.get(&ev_v0[b])); CUDA_TRY(events.get(&ev_v1[b]));
    CUDA_TRY(events.get(&ev_r1[b])); CUDA_TRY(events.get(&ev_m1[b]));
    CUDA_TRY(events.get(&ev_r0[b])); CUDA_TRY(events.get(&ev_m0[b]));
  }
  sp->persistent = persistent;
  cudaEvent_t &ev_start = sp->start, &ev_end = sp->end;
  CUDA_TRY(events.get(&ev_start)); CUDA_TRY(events.get(&ev_end));
  CUDA_TRY(cudaEventRecord(ev_start, r->sv));
  CUDA_TRY(cudaStreamWaitEvent(r->sm, ev_start, 0));
  CUDA_TRY(cudaStreamWaitEvent(r->sr_, ev_start, 0));
  CUDA_TRY(cudaStreamWaitEvent(r->sr2, ev_start, 0));
  // host output in page-locked memory: copy block by block on the copy stream (an async copy to pageable memory would
  // block this thread in the middle of the launch loop, so that case keeps the single copy at the end)
  bool host_copy_per_block = false;
  if (out_host) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, out_host) == cudaSuccess && pa.type == cudaMemoryTypeHost) {
      if (!r->sc) CUDA_TRY(cudaStreamCreateWithFlags(&r->sc, cudaStreamNonBlocking));
      host_copy_per_block = true;
    } else {
      cudaGetLastError();
    }
  }
  // A main mixer that only passes its single sub-mixer through (no effects, no sources of its own, no external input, no
  // meter, master volume at rest -- one rank's share of a sharded graph looks like this): the sub-mixer's kernel applies its
  // own silence gate and the master gain and writes the output; the main level's two launches per block are skipped.
  bool direct_child = false;
  {
    static const bool off = getenv("PB200_NO_DIRECT_CHILD") != nullptr;
    const bool master_rest = r->h_master.current == r->h_master.target;
    bool main_sources = false;
    for (auto& g : r->groups) if (!g.removed && g.gp.mixer == 0) main_sources = true;
    direct_child = !off && c.levels.size() >= 2 && c.levels[0].size() == 1 && c.levels[0][0] == 0 && c.levels[1].size() == 1 &&
                   r->mixers[0].effects.empty() && !main_sources && r->n_ext == 0 && r->meter_interval == UINT64_MAX && master_rest;
  }
  static const bool no_alt = getenv("PB200_NO_REPLAY_ALT") != nullptr;
  const bool replay_alt = persistent && !no_alt && r->n_hq == 0 && r->n_gran_rows == 0;
  uint64_t launches = 0;
  uint32_t gen0 = 0;
  cudaEvent_t& ev_skel_end = sp->skel_end;
  CUDA_TRY(events.get(&ev_skel_end));
  // PB200_SKEL_PROF=<file>: per-voice cycle counters of the skeleton pass (debug aid)
  unsigned long long* prof_buf = nullptr;
  if (getenv("PB200_SKEL_PROF")) {
    CUDA_TRY(cudaMalloc((void**)&prof_buf, nvoices * 12 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(prof_buf, 0, nvoices * 12 * sizeof(unsigned long long)));
  }

  // PlaybackStatusEvent records of the file playbacks: a Position per second of output + a Stopped each
  uint32_t n_file_groups = 0;
  for (auto& g : r->groups) if (g.gp.kind == GROUP_FILE && !g.removed) ++n_file_groups;
  const uint32_t pos_rate = r->cfg.sample_rate;  // FilePlaybackOptions::default().playback_pos_emit_rate = 1 s (file.rs:110)
  const uint32_t status_cap = n_file_groups ? n_file_groups * (uint32_t)(frames / pos_rate + 4) : 0u;
  if (status_cap) CUDA_TRY(r->d_status.reserve(status_cap));
  CUDA_TRY(r->d_status_count.reserve(2));   // [0] status records appended, [1] snapshot-list overflow flag
  CUDA_TRY(cudaMemsetAsync(r->d_status_count.p, 0, 2 * sizeof(uint32_t), r->sv));
  const bool metering = r->meter_interval != UINT64_MAX;
  const size_t meter_rows = (size_t)(frames / bf);
  if (metering) {
    CUDA_TRY(r->d_meter.reserve(meter_rows * 4));
    CUDA_TRY(cudaMemsetAsync(r->d_meter.p, 0, meter_rows * 4 * sizeof(double), r->sm));
  }
  unsigned long long* fx_prof = nullptr;
  const bool fx_prof_all = getenv("PB200_FX_PROF") && atoi(getenv("PB200_FX_PROF")) >= 2;   // per mixer and stage
  const size_t fx_prof_n = fx_prof_all ? r->mixers.size() * MAX_FX_STAGES * 8 : 8;
  if (getenv("PB200_FX_PROF")) { CUDA_TRY(cudaMalloc((void**)&fx_prof, fx_prof_n * 8)); CUDA_TRY(cudaMemset(fx_prof, 0, fx_prof_n * 8)); }
  for (uint32_t b = 0; b < n_blocks; ++b) {
    const uint64_t b0 = p0 + (uint64_t)b * tb;
    const uint32_t blen = (uint32_t)std::min<uint64_t>(tb, p1 - b0);
    const uint32_t slot = b % RING;
    // thin graphs: the replay launches of consecutive blocks are independent (ring slots apart) and far from filling the GPU --
    // alternate two streams so that a block's replay does not queue behind the previous one's (grain / sinc passes carry state: one stream)
    cudaStream_t srb = (replay_alt && (b & 1u)) ? r->sr2 : r->sr_;
    const uint32_t sslot = persistent ? b : slot;  // slot of the tables the skeleton pass owns
    // pass 1 (skeleton) may not overwrite the segment slot the replay of block b-RING still reads
    // ... nor the group-flag slot the mixer of block b-RING still reads
    if (!persistent && b >= RING) { CUDA_TRY(cudaStreamWaitEvent(r->sv, ev_r1[b - RING], 0)); CUDA_TRY(cudaStreamWaitEvent(r->sv, ev_m1[b - RING], 0)); }
    if (!persistent || b == 0) CUDA_TRY(cudaEventRecord(ev_v0[b], r->sv));
    SkeletonArgs va;
    va.groups = r->d_groups.p; va.gstate = r->d_gstate.p; va.voices = r->d_voices.p; va.buffers = r->d_buffers.p;
    va.events = r->d_events.p;
    va.chunk_bounds = r->d_bounds.p;
    va.mixer_chunk_begin = r->d_chunk_begin.p + (size_t)b * (nm + 1);
    va.group_flags = r->d_group_flags.p + (size_t)sslot * ng * max_chunks;
    va.max_chunks = max_chunks; va.block_frames = tb; va.block_start = b0; va.rc = r->rc;
    va.segs = r->d_segs.p + (size_t)sslot * nvoices * seg_cap;
    va.seg_first = r->d_seg_first.p + (size_t)sslot * nvoices * n_tiles;
    va.seg_count = r->d_seg_count.p + (size_t)sslot * nvoices * n_tiles;
    va.recs = r->d_recs.p + (size_t)sslot * nvoices * n_tiles;
    if (!persistent) va.gen = next_generation();
    else { if (b == 0) gen0 = reserve_generations(n_blocks); va.gen = gen0 + b; }
    va.gsegs = r->d_gsegs.p + (size_t)sslot * ng * seg_cap;
    va.gseg_first = r->d_gseg_first.p + (size_t)sslot * ng * n_tiles;
    va.gseg_count = r->d_gseg_count.p + (size_t)sslot * ng * n_tiles;
    va.seg_cap = seg_cap; va.n_tiles = n_tiles;
    va.hq_states = n_hq ? r->d_hq.p : nullptr;
    va.hq_recs = n_hq ? r->d_hq_recs.p + (size_t)slot * hq_cap : nullptr;
    va.hq_n_recs = n_hq ? r->d_hq_nrecs.p + b : nullptr;   // one counter per block: read back after the render
    va.hq_cap = hq_cap;
    va.gran_groups = n_rows ? r->d_gran_groups.p : nullptr;
    va.gran_states = n_rows ? r->d_gran_states.p : nullptr;
    std::memset(&va.gran, 0, sizeof(va.gran));
    if (n_rows) {
      va.gran.recs = r->d_grain_recs.p + (size_t)slot * gran_rec_cap;
      va.gran.counters = r->d_gran_counters.p + 2 * (size_t)b;
      va.gran.rec_cap = gran_rec_cap;
      va.gran.vrec = r->d_gran_vrec.p + (size_t)slot * n_rows * gran_vrec_cap;
      va.gran.vrec_cap = gran_vrec_cap;
      va.gran.tile_range = r->d_gran_tiles.p + (size_t)slot * n_rows * n_tiles * 2;
      va.gran.n_tiles = n_tiles;
      va.gran.gen = va.gen;
      va.gran.block_frames = blen;  // grains stop at the end of what this block really renders: their carry is taken there
      CUDA_TRY(cudaMemsetAsync(va.gran.tile_range, 0xFF, (size_t)n_rows * n_tiles * 2 * sizeof(uint32_t), r->sv));
    }
    va.phase_tabs = r->d_phase_tabs.p; va.phase_dir = r->d_phase_dir.p; va.n_phase = (uint32_t)r->phase_off.size();
    va.status = status_cap ? r->d_status.p : nullptr; va.status_count = r->d_status_count.p; va.overflow = r->d_status_count.p + 1; va.status_cap = status_cap; va.pos_emit_rate = pos_rate;
    va.prof = prof_buf;
    va.debug_flags = getenv("PB200_SKEL_DEBUG") ? (uint32_t)atoi(getenv("PB200_SKEL_DEBUG")) : 0u;
    SkeletonLoop sl;
    std::memset(&sl, 0, sizeof(sl));
    sl.n_blocks = 1;
    if (persistent) {
      sl.n_blocks = n_blocks;
      sl.chunk_begin_stride = (uint32_t)(nm + 1);
      sl.group_flags_stride = ng * max_chunks;
      sl.segs_stride = nvoices * seg_cap; sl.seg_tab_stride = nvoices * n_tiles;
      sl.gsegs_stride = ng * seg_cap; sl.gseg_tab_stride = ng * n_tiles; sl.recs_stride = nvoices * n_tiles;
      sl.block_done = r->d_block_done.p;
      if (autonomous) {
        sl.quiet_block = r->d_quiet_block.p;
        sl.auton_done = r->d_auton.p;
        sl.auton_cnt = r->d_auton.p + c.classes[0].groups.size() * (size_t)n_blocks;
      }
    }
    for (size_t ci = 0; ci < c.classes.size() && (!persistent || b == 0); ++ci) {
      const SizeClass& sc = c.classes[ci];
      va.group_list = r->d_class_groups.p + c.class_offsets[ci];
      // Warp per voice (latency-optimal: voices never serialise each other's control flow) while every group of the class
      // can be resident at once; beyond that one lane per voice (a warp per group of up to 32 voices) fits 16x more voices
      // per SM and the branch-free phase loops keep the lanes converged.
      static const int force_wpv = getenv("PB200_SKEL_WPV") ? atoi(getenv("PB200_SKEL_WPV")) : -1;
      const size_t wpv_resident = (size_t)sm_count_all * std::max<size_t>(1, 65536 / (224 * 32 * (size_t)sc.vpad));
      const bool wpv = force_wpv >= 0 ? force_wpv != 0 : (persistent || sc.groups.size() <= wpv_resident);
      sl.tab_slots = 0;
      // launches whose voices are all granular / HighQuality take the variant without the simple-call machinery
      bool plain_only = true;
      for (uint32_t gi : sc.groups) {
        const bool gran = r->gran_groups[gi].enabled != 0;
        const bool hq = r->groups[gi].gp.kind == GROUP_FILE && r->h_voices[r->groups[gi].gp.first_voice].hq != 0;
        plain_only &= gran || hq;
      }
      if (sc.vpad <= 8 && wpv && plain_only) {
        skeleton_kernel<256, true, false><<<(uint32_t)sc.groups.size(), sc.vpad * 32, 0, r->sv>>>(va, sl);
      } else if (sc.vpad <= 8 && wpv) {  // the voice's jump table in its warp's shared-memory slot
        sl.tab_slots = sc.vpad;
        const size_t smem = (size_t)sl.tab_slots * (TAB_SLOT_WORDS * 4 + 8 + 4) + 16;
        skeleton_kernel<256, true><<<(uint32_t)sc.groups.size(), sc.vpad * 32, smem, r->sv>>>(va, sl);
      }
      else if (sc.vpad <= 32 && wpv) skeleton_kernel<1024, true><<<(uint32_t)sc.groups.size(), sc.vpad * 32, 0, r->sv>>>(va, sl);
      else if (sc.threads <= 256 && plain_only) skeleton_kernel<256, false, false><<<(uint32_t)sc.groups.size(), sc.threads, 0, r->sv>>>(va, sl);
      else if (sc.threads <= 256) skeleton_kernel<256, false><<<(uint32_t)sc.groups.size(), sc.threads, 0, r->sv>>>(va, sl);
      else skeleton_kernel<1024, false><<<(uint32_t)sc.groups.size(), sc.threads, 0, r->sv>>>(va, sl);
      ++launches;
    }
    // pass 2 (replay): needs the segments of this block; may not overwrite a group-bus slot the mixer still reads
    if (!persistent) {
      CUDA_TRY(cudaEventRecord(ev_v1[b], r->sv));
      CUDA_TRY(cudaStreamWaitEvent(srb, ev_v1[b], 0));
    } else {
      if (b == 0) CUDA_TRY(cudaEventRecord(ev_skel_end, r->sv));
      if (stream_wait_value32()((CUstream)srb, (CUdeviceptr)(r->d_block_done.p + b), (cuuint32_t)c.classes[0].groups.size(),
                                CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
        return fail(r, PB200_ERR_CUDA, "cuStreamWaitValue32 failed");
      if (b > 0) CUDA_TRY(cudaEventRecord(ev_v0[b], srb));
      CUDA_TRY(cudaEventRecord(ev_v1[b], srb));
    }
    if (b >= RING) CUDA_TRY(cudaStreamWaitEvent(srb, ev_m1[b - RING], 0));
    ReplayArgs ra;
    ra.groups = r->d_groups.p; ra.buffers = r->d_buffers.p;
    ra.segs = va.segs; ra.seg_first = va.seg_first; ra.seg_count = va.seg_count;
    ra.recs = va.recs; ra.gen = va.gen;
    ra.gsegs = va.gsegs; ra.gseg_first = va.gseg_first; ra.gseg_count = va.gseg_count;
    ra.group_bus = r->d_group_bus.p + (size_t)slot * ng * tb * 2;
    ra.seg_cap = seg_cap; ra.n_tiles = n_tiles; ra.block_frames = tb; ra.rc = r->rc;
    ra.hq_states = va.hq_states;
    ra.hq_scratch = n_hq ? r->d_hq_scratch.p + (size_t)slot * n_hq * tb * 2 : nullptr;
    ra.gran_groups = va.gran_groups;
    std::memset(&ra.gran, 0, sizeof(ra.gran));
    if (!ev_x.empty()) CUDA_TRY(cudaEventRecord(ev_x[3 * (size_t)b], srb));
    if (n_rows) {  // every grain's contribution to this block
      GrainArgs ga;
      ga.recs = va.gran.recs; ga.counters = va.gran.counters; ga.rec_cap = gran_rec_cap; ga.buffers = r->d_buffers.p;
      ga.window_luts = r->d_grain_luts.p;
      ga.storage = r->d_grain_storage.p + (size_t)slot * gran_storage_cap;
      ga.storage_cap = (uint32_t)gran_storage_cap;
      ga.carry_in = r->d_grain_carry[r->gran_parity ^ 1u].p;
      ga.carry_out = r->d_grain_carry[r->gran_parity].p;
      r->gran_parity ^= 1u;
      grain_kernel<<<(gran_rec_cap + 127) / 128, 128, 0, srb>>>(ga);
      ++launches;
      ra.gran.recs = ga.recs; ra.gran.vrec = va.gran.vrec; ra.gran.tile_range = va.gran.tile_range; ra.gran.storage = ga.storage;
      ra.gran.rec_cap = gran_rec_cap; ra.gran.vrec_cap = gran_vrec_cap; ra.gran.n_tiles = n_tiles; ra.gran.storage_cap = ga.storage_cap;
    }
    if (!ev_x.empty()) CUDA_TRY(cudaEventRecord(ev_x[3 * (size_t)b + 1], srb));
    if (n_hq) {  // materialise the resampler output this block consumes: one launch per filter table
      SincArgs sa;
      sa.recs = va.hq_recs; sa.n_recs = va.hq_n_recs; sa.cap = hq_cap; sa.buffers = r->d_buffers.p;
      sa.frames_out = r->d_hq_frames.p;
      sa.tables = r->d_sinc_tables.p; sa.scratch = r->d_hq_scratch.p + (size_t)slot * n_hq * tb * 2; sa.block_frames = tb;
      const uint32_t n_tables = std::max<uint32_t>(1, (uint32_t)r->sinc_table_keys.size());
      for (uint32_t t = 0; t < n_tables; ++t) {
        sa.table = t; sa.do_copies = t == 0;
        sinc_kernel<<<sm_count, SINC_THREADS, SINC_SMEM, srb>>>(sa);
        ++launches;
      }
    }
    if (!ev_x.empty()) CUDA_TRY(cudaEventRecord(ev_x[3 * (size_t)b + 2], srb));
    const uint32_t live_tiles = (blen + TILE - 1) / TILE;
    CUDA_TRY(cudaEventRecord(ev_r0[b], srb));  // stream order: every wait of this block's replay is behind it
    static const bool skel_only = getenv("PB200_SKEL_ONLY") != nullptr;  // timing experiments: the skeleton pass alone (output invalid)
    if (!skel_only && ng > 0) {  // one launch over every group (the class lists are contiguous in d_class_groups)
      ra.group_list = r->d_class_groups.p;
      dim3 grid((live_tiles + REPLAY_THREADS - 1) / REPLAY_THREADS, ng);
      replay_kernel<<<grid, REPLAY_THREADS, REPLAY_SMEM, srb>>>(ra);
      ++launches;
    }
    CUDA_TRY(cudaEventRecord(ev_r1[b], srb));
    CUDA_TRY(cudaStreamWaitEvent(r->sm, ev_r1[b], 0));
    CUDA_TRY(cudaEventRecord(ev_m0[b], r->sm));
    MixerKernelArgs ma;
    ma.mixers = r->d_mixers.p; ma.mstate = r->d_mstate.p; ma.child_index = r->d_child_index.p; ma.source_index = r->d_source_index.p;
    ma.fx = r->d_fx.p; ma.fx_events = r->d_fx_events.p;
    ma.fxc.sample_rate = r->cfg.sample_rate; ma.fxc.comp = r->rc.rate_comp; ma.fxc.state_arena = r->d_fx_state.p; ma.fxc.aux_arena = r->d_aux.p;
    ma.chunk_bounds = r->d_bounds.p; ma.mixer_chunk_begin = va.mixer_chunk_begin;
    ma.group_bus = ra.group_bus; ma.group_flags = va.group_flags;
    ma.mixer_bus = r->d_mixer_bus.p + (size_t)slot * nm * tb * 2;
    ma.mixer_flags = r->d_mixer_flags.p + (size_t)slot * nm * max_chunks;
    ma.max_chunks = max_chunks; ma.block_frames = tb; ma.block_start = b0;
    ma.out = dout + (size_t)(b0 - p0) * 2; ma.master = r->d_master.p; ma.wav_block_frames = bf;
    ma.block_len = blen;
    ma.meter = metering ? r->d_meter.p : nullptr; ma.render_start = p0;
    ma.prof = fx_prof;
    ma.progress = out_dev ? r->progress : nullptr;
    ma.progress_value = progress_base + (b0 - p0) + blen;
    ma.n_ext = (b0 - p0 < r->ext_frames) ? r->n_ext : 0u;
    for (uint32_t e = 0; e < PB_MAX_EXT; ++e) ma.ext_in[e] = e < ma.n_ext ? r->ext_input[e] + (b0 - p0) * 2 : nullptr;
    ma.ext_len = ma.n_ext ? (uint32_t)std::min<uint64_t>(blen, r->ext_frames - (b0 - p0)) : 0u;
    ma.prof_all = fx_prof_all ? 1u : 0u;
    for (int lvl = (int)c.levels.size() - 1; lvl >= 0 && !skel_only; --lvl) {
      if (lvl == 0 && direct_child) break;          // the main mixer's only child has written the output itself
      ma.direct_out = (lvl == 1 && direct_child) ? 1u : 0u;
      ma.level_mixers = r->d_level_mixers.p + c.level_offsets[lvl];
      const uint32_t nlm = (uint32_t)c.levels[lvl].size();
      mix_sum_kernel<<<dim3((blen + 255) / 256, nlm), 256, 0, r->sm>>>(ma);
      bool has_reverb = false;
      for (uint32_t mi : c.levels[lvl]) for (uint32_t fi : r->mixers[mi].effects) has_reverb |= r->fxs[fi].kind == FX_REVERB;
      ma.work_bytes = has_reverb ? FX_WORK_BYTES : FX_WORK_SMALL;
      const uint32_t S = level_stages[lvl];
      ma.n_stages = S;
      ma.stage_begin = S > 1 ? r->d_stage_begin.p + stage_offsets[lvl] : nullptr;
      ma.fx_progress = S > 1 ? r->d_fx_progress.p : nullptr;
      ma.fx_pflags = S > 1 ? r->d_fx_pflags.p : nullptr;
      ma.fx_ticket = S > 1 ? r->d_fx_progress.p + nm * (size_t)MAX_FX_STAGES : nullptr;
      const uint32_t n_fx = c.level_fx_count[lvl], n_plain = nlm - n_fx;
      if (n_fx && S > 1) {
        // (an ordinary launch: the ticket order makes the hand-over deadlock-free; a cooperative launch would serialise the
        // kernel against everything else on the device, the replay of the next block included)
        CUDA_TRY(cudaMemsetAsync(r->d_fx_progress.p, 0, (nm * (size_t)MAX_FX_STAGES + 1) * sizeof(uint32_t), r->sm));
        if (level_small[lvl]) mix_fx_kernel<2><<<dim3(n_fx, S), FX_THREADS, ma.work_bytes, r->sm>>>(ma);
        else mix_fx_kernel<1><<<dim3(n_fx, S), FX_THREADS, ma.work_bytes, r->sm>>>(ma);
      } else if (n_fx) {
        mix_fx_kernel<1><<<n_fx, FX_THREADS, ma.work_bytes, r->sm>>>(ma);
      }
      if (n_plain) {  // the mixers without effects: the light build
        ma.level_mixers += n_fx;
        ma.n_stages = 1; ma.stage_begin = nullptr; ma.fx_progress = nullptr; ma.fx_pflags = nullptr; ma.fx_ticket = nullptr;
        ma.work_bytes = 0;
        mix_fx_kernel<4, true><<<n_plain, FX_THREADS, 0, r->sm>>>(ma);
      }
      launches += 1 + (n_fx ? 1 : 0) + (n_plain ? 1 : 0);
    }
    CUDA_TRY(cudaEventRecord(ev_m1[b], r->sm));
    if (host_copy_per_block) {  // the block's WAV data leaves for the (pinned) host buffer while later blocks still render
      CUDA_TRY(cudaStreamWaitEvent(r->sc, ev_m1[b], 0));
      CUDA_TRY(cudaMemcpyAsync(out_host + (b0 - p0) * 2, dout + (b0 - p0) * 2, (size_t)blen * 2 * sizeof(float), cudaMemcpyDeviceToHost, r->sc));
    }
  }
  CUDA_TRY(cudaEventRecord(ev_end, r->sm));
  hp_mark("launches enqueued");
  if (out_host && !host_copy_per_block) CUDA_TRY(cudaMemcpyAsync(out_host, dout, frames * 2 * sizeof(float), cudaMemcpyDeviceToHost, r->sm));
  if (host_copy_per_block) CUDA_TRY(cudaStreamSynchronize(r->sc));
  CUDA_TRY(cudaStreamSynchronize(r->sm));
  CUDA_TRY(cudaStreamSynchronize(r->sr_));
  CUDA_TRY(cudaStreamSynchronize(r->sr2));
  CUDA_TRY(cudaStreamSynchronize(r->sv));
  CUDA_TRY(cudaGetLastError());
  hp_mark("device done");

  if (fx_prof) {
    std::vector<unsigned long long> hp(fx_prof_n);
    cudaMemcpy(hp.data(), fx_prof, fx_prof_n * 8, cudaMemcpyDeviceToHost);
    cudaFree(fx_prof);
    unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 8; ++i) h[i] = hp[i];
    if (fx_prof_all) {
      for (size_t mi = 0; mi < r->mixers.size(); mi += std::max<size_t>(1, r->mixers.size() / 4))
        for (uint32_t st = 0; st < MAX_FX_STAGES; ++st) {
          const unsigned long long* q = hp.data() + (mi * MAX_FX_STAGES + st) * 8;
          if (q[0] + q[3] + q[6] == 0) continue;
          fprintf(stderr, "fx prof mixer %zu stage %u (Mcycles): ev %.2f stage-in %.2f decide %.2f process %.2f tail %.2f writeback %.2f gate %.2f wait %.2f\n", mi, st,
                  q[0] / 1e6, q[1] / 1e6, q[2] / 1e6, q[3] / 1e6, q[4] / 1e6, q[5] / 1e6, q[6] / 1e6, q[7] / 1e6);
        }
    }
    fprintf(stderr, "fx prof (Mcycles): events+audible %.2f stage %.2f bypass-decision %.2f process %.2f tail %.2f writeback %.2f master/gate %.2f\n",
            h[0] / 1e6, h[1] / 1e6, h[2] / 1e6, h[3] / 1e6, h[4] / 1e6, h[5] / 1e6, h[6] / 1e6);
  }
  if (prof_buf) {
    std::vector<unsigned long long> h(nvoices * 12);
    cudaMemcpy(h.data(), prof_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(prof_buf);
    if (FILE* f = fopen(getenv("PB200_SKEL_PROF"), "w")) {
      fprintf(f, "voice,wait_before_run,wait_after_voices,wait_after_bookkeeping,free_run_work,simple_calls,simple_cycles,jumped_tiles,literal_frames,general_frames,general_cycles,unused,block_cycles\n");
      for (size_t i = 0; i < nvoices; ++i) {
        fprintf(f, "%zu", i);
        for (int k = 0; k < 12; ++k) fprintf(f, ",%llu", h[12 * i + k]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
#ifdef PB200_CYC
  {
    unsigned long long h[8];
    cudaMemcpyFromSymbol(h, g_cyc, sizeof(h));
    fprintf(stderr, "cyc voice %d: simple_call total %.2fM (loops %.2fM) calls %llu frames %llu | run_call total %.2fM, non-simple frames %llu in %.2fM | block fn total %.2fM\n",
            (int)PB200_CYC, h[0] / 1e6, h[1] / 1e6, h[2], h[3], h[4] / 1e6, h[5], h[6] / 1e6, h[7] / 1e6);
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_cyc, z, sizeof(z));
  }
#endif
  r->stats = pb200_render_stats{};
  if (host_prof) {  // device timeline of the call, ms since its first event
    auto at = [&](cudaEvent_t e) { float t = 0; cudaEventElapsedTime(&t, ev_start, e); return t; };
    fprintf(stderr, "[dev] end %.3f", at(ev_end));
    if (persistent) fprintf(stderr, " | skeleton %.3f..%.3f", at(ev_v0[0]), at(ev_skel_end));
    fprintf(stderr, "\n[dev] per block (replay start..end, mixer start..end):");
    for (uint32_t b = 0; b < n_blocks; ++b) fprintf(stderr, " [%u] %.2f..%.2f %.2f..%.2f", b, at(ev_r0[b]), at(ev_r1[b]), at(ev_m0[b]), at(ev_m1[b]));
    fprintf(stderr, "\n");
  }
  r->stats.kernel_launches = launches;
  hp_mark("  event spans");
  if (n_hq) {
    std::vector<uint32_t> counts(n_blocks);
    CUDA_TRY(cudaMemcpy(counts.data(), r->d_hq_nrecs.p, n_blocks * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (uint32_t c2 : counts)
      if (c2 > hq_cap) return fail(r, PB200_ERR_CUDA, "HighQuality chunk record list overflowed");
    unsigned long long fo = 0;
    CUDA_TRY(cudaMemcpy(&fo, r->d_hq_frames.p, sizeof(fo), cudaMemcpyDeviceToHost));
    r->stats.sinc_frames = fo;
  }
  if (n_rows) {
    std::vector<uint32_t> counts(2 * (size_t)n_blocks);
    CUDA_TRY(cudaMemcpy(counts.data(), r->d_gran_counters.p, counts.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (uint32_t b = 0; b < n_blocks; ++b)
      if (counts[2 * b] > gran_rec_cap || counts[2 * b + 1] > gran_storage_cap)
        return fail(r, PB200_ERR_CUDA, "granular record list / grain storage overflowed");
    for (uint32_t b = 0; b < n_blocks; ++b) r->stats.grain_samples += counts[2 * b + 1];
  }
  uint32_t counts2[2] = {0, 0};
  CUDA_TRY(cudaMemcpy(counts2, r->d_status_count.p, sizeof(counts2), cudaMemcpyDeviceToHost));
  if (counts2[1]) return fail(r, PB200_ERR_CUDA, "a voice's snapshot list filled up (more write calls per time block than seg_cap allows): output invalid");
  if (status_cap) {  // PlaybackStatusEvent stream, in emission order (frame, then the mixer's source order ~ id)
    uint32_t n = counts2[0];
    if (n > status_cap) return fail(r, PB200_ERR_CUDA, "status event list overflowed");
    std::vector<StatusRec> recs(n);
    if (n) CUDA_TRY(cudaMemcpy(recs.data(), r->d_status.p, n * sizeof(StatusRec), cudaMemcpyDeviceToHost));
    std::sort(recs.begin(), recs.end(), [&](const StatusRec& a, const StatusRec& b) {
      if (a.frame != b.frame) return a.frame < b.frame;
      if (a.group != b.group) return r->groups[a.group].public_id < r->groups[b.group].public_id;
      return a.kind < b.kind;
    });
    for (const StatusRec& sr : recs) {
      pb200_status_event e;
      std::memset(&e, 0, sizeof(e));
      const DevBuffer& b = r->buffers[r->groups[sr.group].gp.buffer].dev;
      e.frame = sr.frame; e.playback_id = r->groups[sr.group].public_id;
      if (sr.kind == 0) {
        e.kind = PB200_STATUS_POSITION;
        const double second_pos = (double)(sr.pos / b.channels) / (double)b.sample_rate;   // file/common.rs:195-196
        e.position_nanos = (uint64_t)std::nearbyint(second_pos * 1.0e9);
      } else {
        e.kind = PB200_STATUS_STOPPED; e.exhausted = sr.kind == 1;
      }
      r->status_events.push_back(e);
    }
  }
  hp_mark("  list checks + status");
  r->host_state_valid = false;
  r->position = p1;
  // statistics + event cursors come back with the (small) group state
  {
    std::vector<GroupState> gs(r->dev_n_groups);
    CUDA_TRY(r->d_gstate.download(gs, r->sm));
    CUDA_TRY(cudaStreamSynchronize(r->sm));
    uint64_t vf = 0;
    for (auto& g : gs) vf += g.voice_frames;
    r->stats.voice_frames = vf - r->voice_frames_total;
    r->voice_frames_total = vf;
    // WavStream stops at the first block whose MixedSource::write returns 0 (wav.rs:231-234): the main mixer has no
    // playing source, effect, sub-mixer or pending event left (mixed.rs:664-670). Sources are dropped at the end of
    // the block they finished in; an event is popped by the chunk that starts at its time.
    uint64_t written = frames;
    if (live_fx == 0 && live_mixers == 1 && !r->n_ext) {
      bool all_dead = true;
      uint64_t fin = p0;
      for (size_t gi = 0; gi < gs.size(); ++gi) {
        if (!gs[gi].dead) { all_dead = false; break; }
        fin = std::max(fin, (gs[gi].dead_time + bf - 1) / bf * bf);
      }
      if (all_dead) {
        for (auto& g : r->groups)
          for (auto& e : g.events)
            if (!(e.ev.flags & 1u)) fin = std::max(fin, (e.ev.time / bf + 1) * bf);
        if (fin <= p1) {
          r->finished = true;
          written = fin - p0;
          if (written < frames) {
            if (out_host) std::memset(out_host + written * 2, 0, (frames - written) * 2 * sizeof(float));
            if (out_dev) { CUDA_TRY(cudaMemsetAsync(out_dev + written * 2, 0, (frames - written) * 2 * sizeof(float), r->sm)); CUDA_TRY(cudaStreamSynchronize(r->sm)); }
          }
        }
      }
    }
    if (metering) {  // AudioLevelState::record per WavStream block (metered.rs:107-148) from the device's per-block reductions
      std::vector<double> rows(meter_rows * 4);
      CUDA_TRY(cudaMemcpy(rows.data(), r->d_meter.p, rows.size() * sizeof(double), cudaMemcpyDeviceToHost));
      for (size_t bi = 0; bi < std::min<size_t>(meter_rows, (size_t)(written / bf)); ++bi) {  // (blocks the stream really wrote)
        const uint64_t t = p0 + bi * bf;
        for (int c2 = 0; c2 < 2; ++c2) {
          r->meter_peak_hold[c2] = std::max(r->meter_peak_hold[c2], (float)rows[bi * 4 + c2]);
          r->meter_sum_square[c2] += rows[bi * 4 + 2 + c2];
        }
        r->meter_frames += bf;
        if (t - std::min(t, r->meter_clock) >= r->meter_interval) {
          for (int c2 = 0; c2 < 2; ++c2) {
            r->audio_level.peak[c2] = r->meter_peak_hold[c2];
            r->audio_level.rms[c2] = r->meter_frames ? (float)std::sqrt(r->meter_sum_square[c2] / (double)r->meter_frames) : 0.0f;
            r->meter_peak_hold[c2] = 0.0f; r->meter_sum_square[c2] = 0.0;
          }
          r->meter_clock = t; r->meter_frames = 0;
        }
      }
    }
    if (r->finished) r->position = p0 + written;  // WavStream::playback_pos stops with the stream
    if (frames_written) *frames_written = written;
  }
  hp_mark("state download + bookkeeping");
  r->n_ext = 0; r->ext_frames = 0;   // (external main-mixer inputs serve one render call)
  sp->complete = true;               // every stream has been drained above: the spans can be read whenever they are wanted
  r->spans = std::move(sp);
  return PB200_OK;
}

}  // namespace

extern "C" {

int pb200_render(pb200_renderer* r, float* out, uint64_t frames, uint64_t* frames_written) {
  if (!r || !out) return PB200_ERR_PARAMETER;
  return render_impl(r, nullptr, out, frames, frames_written);
}

int pb200_render_device(pb200_renderer* r, float* out_device, uint64_t frames, uint64_t* frames_written) {
  if (!r || !out_device) return PB200_ERR_PARAMETER;
  return render_impl(r, out_device, nullptr, frames, frames_written);
}

int pb200_set_main_inputs(pb200_renderer* r, const float* const* buses_device, uint32_t count, uint64_t frames) {
  if (!r || (count && !buses_device)) return PB200_ERR_PARAMETER;
  if (count > PB_MAX_EXT) return fail(r, PB200_ERR_PARAMETER, "too many main-mixer inputs");
  if (count && frames % r->cfg.block_frames != 0) return fail(r, PB200_ERR_PARAMETER, "frames must be a multiple of block_frames");
  for (uint32_t i = 0; i < count; ++i) if (!buses_device[i]) return fail(r, PB200_ERR_PARAMETER, "null main-mixer input");
  for (uint32_t i = 0; i < count; ++i) r->ext_input[i] = buses_device[i];
  r->n_ext = count;
  r->ext_frames = count ? frames : 0;
  return PB200_OK;
}

int pb200_set_main_input(pb200_renderer* r, const float* bus_device, uint64_t frames) {
  return pb200_set_main_inputs(r, &bus_device, bus_device ? 1u : 0u, frames);
}

// ---- peer memory (sharded renders): plain cudaMalloc allocations shared between the ranks' processes through CUDA IPC; pushes
// are stream-ordered DMA copies plus a stream memory write as the "piece has landed" flag -- no kernel, so nothing waits for
// SMs on a device that is busy rendering (an NCCL reduce kernel does: 5-10 ms per piece behind a shard's launches) ----------
int pb200_device_alloc(int device_ordinal, size_t bytes, void** ptr) {
  if (!ptr || !bytes) return PB200_ERR_PARAMETER;
  if (device_ordinal >= 0 && cudaSetDevice(device_ordinal) != cudaSuccess) return PB200_ERR_CUDA;
  if (cudaMalloc(ptr, bytes) != cudaSuccess) return PB200_ERR_CUDA;
  if (cudaMemset(*ptr, 0, bytes) != cudaSuccess) return PB200_ERR_CUDA;
  return PB200_OK;
}
int pb200_device_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? PB200_OK : PB200_ERR_CUDA; }
int pb200_ipc_export(const void* ptr, void* handle64) {
  if (!ptr || !handle64) return PB200_ERR_PARAMETER;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)) != cudaSuccess) return PB200_ERR_CUDA;
  std::memcpy(handle64, &h, 64);
  return PB200_OK;
}
int pb200_ipc_open(const void* handle64, int device_ordinal, void** ptr) {
  if (!handle64 || !ptr) return PB200_ERR_PARAMETER;
  if (device_ordinal >= 0 && cudaSetDevice(device_ordinal) != cudaSuccess) return PB200_ERR_CUDA;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  return cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? PB200_OK : PB200_ERR_CUDA;
}
int pb200_ipc_close(void* ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? PB200_OK : PB200_ERR_CUDA; }

int pb200_push_async(pb200_renderer* r, void* dst_peer, const void* src_device, size_t bytes, uint32_t* flag_peer, uint32_t flag_value) {
  if (!r || !dst_peer || !src_device) return PB200_ERR_PARAMETER;
  cudaSetDevice(r->device);
  if (!r->sc) CUDA_TRY(cudaStreamCreateWithFlags(&r->sc, cudaStreamNonBlocking));
  CUDA_TRY(cudaMemcpyAsync(dst_peer, src_device, bytes, cudaMemcpyDefault, r->sc));
  if (flag_peer) {  // the flag follows the data on the same stream: a 4-byte copy out of a pinned ring (64 pushes may be in flight)
    if (!r->push_flags) CUDA_TRY(cudaHostAlloc((void**)&r->push_flags, 64 * sizeof(uint32_t), cudaHostAllocPortable));
    uint32_t* slot = r->push_flags + (r->push_seq++ & 63u);
    *slot = flag_value;
    CUDA_TRY(cudaMemcpyAsync(flag_peer, slot, sizeof(uint32_t), cudaMemcpyDefault, r->sc));
  }
  return PB200_OK;
}
int pb200_push_sync(pb200_renderer* r) {
  if (!r) return PB200_ERR_PARAMETER;
  if (r->sc) CUDA_TRY(cudaStreamSynchronize(r->sc));
  return PB200_OK;
}
int pb200_peek_u32(pb200_renderer* r, const uint32_t* src_device, uint32_t count, uint32_t* out_host) {
  if (!r || !src_device || !out_host) return PB200_ERR_PARAMETER;
  cudaSetDevice(r->device);
  if (!r->sc) CUDA_TRY(cudaStreamCreateWithFlags(&r->sc, cudaStreamNonBlocking));
  CUDA_TRY(cudaMemcpyAsync(out_host, src_device, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, r->sc));
  CUDA_TRY(cudaStreamSynchronize(r->sc));
  return PB200_OK;
}

uint64_t pb200_trim_pool(int device_ordinal) { return (uint64_t)DevicePool::get().trim(device_ordinal); }

uint64_t pb200_render_progress(const pb200_renderer* r) { return r && r->progress ? __atomic_load_n(r->progress, __ATOMIC_ACQUIRE) : 0; }

int pb200_schedule_many(pb200_renderer* r, pb200_event* events, uint32_t count, uint32_t* scheduled) {
  if (!r || (!events && count)) return PB200_ERR_PARAMETER;
  uint32_t i = 0;
  int rc = PB200_OK;
  for (; i < count; ++i) {
    pb200_event& ev = events[i];
    if (ev.flags & PB200_EVF_NOTE_FROM_BATCH) {
      const uint64_t idx = ev.note_id;
      if (idx >= i || events[idx].kind != PB200_EV_NOTE_ON) { rc = fail(r, PB200_ERR_PARAMETER, "batch note reference must point at an earlier NOTE_ON"); break; }
      ev.note_id = events[idx].note_id;
      ev.flags &= ~PB200_EVF_NOTE_FROM_BATCH;
    }
    if ((rc = pb200_schedule(r, &ev)) != PB200_OK) break;
  }
  if (scheduled) *scheduled = i;
  return rc;
}

int pb200_decode_wav(const char* path, float** interleaved, pb200_wav_info* info) {
  if (!path || !interleaved || !info) return PB200_ERR_PARAMETER;
  pbh::WavData w;
  if (int e = pbh::decode_wav_file(path, w)) return e;
  float* p = (float*)std::malloc(w.samples.size() * sizeof(float));
  if (!p) return PB200_ERR_IO;
  std::memcpy(p, w.samples.data(), w.samples.size() * sizeof(float));
  *interleaved = p;
  *info = w.info;
  return PB200_OK;
}

void pb200_free(void* p) { std::free(p); }

int pb200_upload_wav(pb200_renderer* r, const char* path, uint32_t* buffer_id, pb200_wav_info* info) {
  if (!r || !path || !buffer_id) return PB200_ERR_PARAMETER;
  pbh::WavData w;
  if (int e = pbh::decode_wav_file(path, w)) return fail(r, e, e == PB200_ERR_MEDIA_FILE_NOT_FOUND ? "Audio file not found" : "Audio file failed to probe / decode");
  if (info) *info = w.info;
  return pb200_upload_buffer(r, w.samples.data(), w.info.frames, w.info.channels, w.info.sample_rate, w.info.loop_start, w.info.loop_end, 1, buffer_id);
}

int pb200_render_to_wav(pb200_renderer* r, const char* path, uint64_t duration_nanos, uint64_t* frames_written) {
  if (!r || !path) return PB200_ERR_PARAMETER;
  const uint64_t frames = pbh::wav_stream_frames(duration_nanos, r->cfg.sample_rate, r->cfg.block_frames ? r->cfg.block_frames : 1024u);
  std::vector<float> out((size_t)frames * 2);
  uint64_t written = 0;
  if (frames) { if (int e = render_impl(r, nullptr, out.data(), frames, &written)) return e; }
  if (frames_written) *frames_written = written;
  if (int e = pbh::write_wav_f32(path, out.data(), written, 2, r->cfg.sample_rate)) return fail(r, e, "failed to write the WAV file");
  return PB200_OK;
}

int pb200_source_status_get(pb200_renderer* r, uint32_t id, pb200_source_status* st) {
  if (!r || !st) return PB200_ERR_PARAMETER;
  auto it = r->group_by_id.find(id);
  if (it == r->group_by_id.end()) return fail(r, PB200_ERR_SOURCE_NOT_PLAYING, "Source is no longer playing");
  if (int e = sync_state_to_host(r)) return e;
  std::memset(st, 0, sizeof(*st));
  st->end_frame = UINT64_MAX;
  const HostGroup& g = r->groups[it->second];
  const GroupState& gs = r->h_gstate[it->second];
  st->is_playing = gs.dead ? 0 : 1;
  if (g.gp.kind == GROUP_FILE && !gs.dead) {
    const VoiceState& v = r->h_voices[g.gp.first_voice];
    st->playback_pos = v.playback_pos;
    st->exhausted = v.stopped_exhausted;
    st->end_frame = v.end_frame;
  }
  return PB200_OK;
}

int pb200_sampler_voice_states(pb200_renderer* r, uint32_t id, pb200_voice_state* out, uint32_t capacity, uint32_t* count) {
  if (!r || !out || !count) return PB200_ERR_PARAMETER;
  auto it = r->group_by_id.find(id);
  if (it == r->group_by_id.end() || r->groups[it->second].gp.kind != GROUP_SAMPLER) return fail(r, PB200_ERR_GENERATOR_NOT_FOUND, "Generator not found");
  if (int e = sync_state_to_host(r)) return e;
  const HostGroup& g = r->groups[it->second];
  if (r->h_gstate[it->second].dead) { *count = 0; return PB200_OK; }  // dropped by its mixer
  uint32_t n = std::min(capacity, g.gp.n_voices);
  for (uint32_t i = 0; i < n; ++i) {
    const VoiceState& v = r->h_voices[g.gp.first_voice + i];
    out[i].note_id = v.has_note ? v.note_id : UINT64_MAX;
    out[i].playback_pos = v.playback_pos;
    out[i].envelope_stage = v.env_stage;
    out[i].active = v.has_note;
  }
  *count = g.gp.n_voices;
  return PB200_OK;
}

int pb200_poll_status(pb200_renderer* r, pb200_status_event* out, uint32_t capacity, uint32_t* count) {
  if (!r || !count || (!out && capacity)) return PB200_ERR_PARAMETER;
  const uint32_t n = (uint32_t)std::min<size_t>(capacity, r->status_events.size());
  for (uint32_t i = 0; i < n; ++i) out[i] = r->status_events[i];
  r->status_events.erase(r->status_events.begin(), r->status_events.begin() + n);
  *count = n;
  return PB200_OK;
}

int pb200_set_metering_interval(pb200_renderer* r, uint64_t interval_nanos) {
  if (!r) return PB200_ERR_PARAMETER;
  // SampleTimeClock::duration_to_sample_time (utils/time.rs:28-35)
  r->meter_interval = interval_nanos == PB200_DURATION_NONE ? UINT64_MAX : (uint64_t)(nanos_as_secs_f64(interval_nanos) * (double)r->cfg.sample_rate);
  return PB200_OK;
}

int pb200_get_audio_level(pb200_renderer* r, pb200_audio_level* out) {
  if (!r || !out) return PB200_ERR_PARAMETER;
  if (r->meter_interval == UINT64_MAX) return fail(r, PB200_ERR_PARAMETER, "metering is off (PlayerConfig::metering_interval is None)");
  *out = r->audio_level;
  return PB200_OK;
}

int pb200_last_render_stats(pb200_renderer* r, pb200_render_stats* st) {
  if (!r || !st) return PB200_ERR_PARAMETER;
  if (r->spans) {  // the per-pass spans of the last call, computed on first request
    const SpanEvents& e = *r->spans;
    float ms = 0;
    cudaEventElapsedTime(&ms, e.start, e.end);
    r->stats.device_ms = ms;
    for (size_t b = 0; b < e.r0.size(); ++b) {
      if (!e.persistent) { cudaEventElapsedTime(&ms, e.v0[b], e.v1[b]); r->stats.skeleton_kernel_ms += ms; }
      else if (b == 0) { cudaEventElapsedTime(&ms, e.v0[0], e.skel_end); r->stats.skeleton_kernel_ms += ms; }
      cudaEventElapsedTime(&ms, e.r0[b], e.r1[b]); r->stats.voice_kernel_ms += ms;   // the replay launches alone
      cudaEventElapsedTime(&ms, e.m0[b], e.m1[b]); r->stats.effect_kernel_ms += ms;
    }
    for (size_t b = 0; 3 * b + 2 < e.x.size(); ++b) {
      cudaEventElapsedTime(&ms, e.x[3 * b], e.x[3 * b + 1]); r->stats.grain_kernel_ms += ms;
      cudaEventElapsedTime(&ms, e.x[3 * b + 1], e.x[3 * b + 2]); r->stats.sinc_kernel_ms += ms;
    }
    r->spans.reset();
  }
  *st = r->stats;
  return PB200_OK;
}

}  // extern "C"
