// ORACLE -- TEST INFRASTRUCTURE ONLY (see po_dsp.hpp header).
//
// Exposes the CPU restatement through the SAME C signatures as include/phonic_b200.h, with the
// `pb200_` prefix replaced by `po_`, so that tests drive the oracle and the CUDA product with one
// scene description. Mirrors Player (src/player.rs:289-1046) + WavStream (src/output/wav.rs:196-250).
#define pb200_create po_create
#define pb200_destroy po_destroy
#define pb200_last_error po_last_error
#define pb200_backend po_backend
#define pb200_upload_buffer po_upload_buffer
#define pb200_add_mixer po_add_mixer
#define pb200_add_effect po_add_effect
#define pb200_file_options_default po_file_options_default
#define pb200_play_file po_play_file
#define pb200_sampler_options_default po_sampler_options_default
#define pb200_add_sampler po_add_sampler
#define pb200_schedule po_schedule
#define pb200_render po_render
#define pb200_render_device po_render_device
#define pb200_position po_position
#define pb200_source_status_get po_source_status_get
#define pb200_sampler_voice_states po_sampler_voice_states
#define pb200_last_render_stats po_last_render_stats
#define pb200_decode_wav po_decode_wav
#define pb200_free po_free
#define pb200_upload_wav po_upload_wav
#define pb200_render_to_wav po_render_to_wav
#define pb200_schedule_many po_schedule_many
#define pb200_remove_source po_remove_source
#define pb200_remove_mixer po_remove_mixer
#define pb200_remove_effect po_remove_effect
#define pb200_move_effect po_move_effect
#define pb200_stop_all_sources po_stop_all_sources
#define pb200_poll_status po_poll_status
#define pb200_set_metering_interval po_set_metering_interval
#define pb200_get_audio_level po_get_audio_level
#define pb200_set_main_input po_set_main_input
#define pb200_render_progress po_render_progress
#define pb200_set_main_inputs po_set_main_inputs
#define pb200_device_alloc po_device_alloc
#define pb200_device_free po_device_free
#define pb200_ipc_export po_ipc_export
#define pb200_ipc_open po_ipc_open
#define pb200_ipc_close po_ipc_close
#define pb200_push_async po_push_async
#define pb200_push_sync po_push_sync
#define pb200_peek_u32 po_peek_u32
#define pb200_trim_pool po_trim_pool
#include "../include/phonic_b200.h"

#include <atomic>
#include <chrono>
#include <map>
#include <string>

#include "po_effects.hpp"

using namespace po;

struct pb200_renderer {
  pb200_config cfg;
  std::string last_error;
  std::unique_ptr<MixedSource> main;
  std::map<uint32_t, MixedSource*> mixers;  // id -> mixer (0 = main)
  std::map<uint32_t, uint32_t> mixer_parents; // sub-mixer id -> parent id
  std::vector<std::shared_ptr<AudioFileBuffer>> buffers;
  struct SourceRef { uint32_t mixer; PlaybackQueues queues; PreloadedFileSource* file = nullptr; Sampler* sampler = nullptr; bool transient = true; };
  std::map<uint32_t, SourceRef> sources;
  struct EffectRef { uint32_t mixer; uint32_t kind; };
  std::map<uint32_t, EffectRef> effects;
  uint32_t next_source_id = 1, next_mixer_id = 1, next_effect_id = 1;
  uint64_t next_note_id = 1;
  // WavStream
  ExpSmoothed smoothed_volume;
  uint64_t playback_pos = 0;  // samples
  bool finished = false;
  std::atomic<uint64_t> progress{0};   // pb200_render_progress
  std::vector<float> block;
  double last_ms = 0;
  uint64_t last_voice_frames = 0;
  // observability: PlaybackStatusEvent stream (file sources push into `raw_status`) + MeteredSource of the main mixer
  std::vector<PreloadedFileSource::StatusEv> raw_status;
  std::vector<pb200_status_event> status_events;
  std::map<uint32_t, std::pair<uint32_t, uint32_t>> status_file_format;  // playback id -> (channels, rate) of its buffer
  uint64_t meter_interval = UINT64_MAX, meter_clock = 0, meter_frames = 0;
  float meter_peak_hold[2] = {0, 0};
  double meter_sum_square[2] = {0, 0};
  pb200_audio_level audio_level{};
};

static int fail(pb200_renderer* r, int code, const std::string& msg) {
  if (r) r->last_error = msg;
  return code;
}

extern "C" {

const char* pb200_backend(void) { return "oracle: scalar C++ restatement of emuell/phonic v0.16.0 (CPU)"; }

int pb200_create(const pb200_config* config, pb200_renderer** out) {
  if (!config || !out) return PB200_ERR_PARAMETER;
  if (config->channel_count != 2 || config->sample_rate == 0) return PB200_ERR_PARAMETER;
  auto* r = new pb200_renderer();
  r->cfg = *config;
  if (r->cfg.block_frames == 0) r->cfg.block_frames = 1024;
  r->main = std::make_unique<MixedSource>(config->channel_count, config->sample_rate);
  r->mixers[0] = r->main.get();
  // WavStream: smoothed_volume = ExponentialSmoothedValue::new(1.0, sr) then set_volume (wav.rs)
  r->smoothed_volume = ExpSmoothed(1.0f, config->sample_rate);
  r->smoothed_volume.init(config->master_volume);
  r->block.resize((size_t)r->cfg.block_frames * config->channel_count);
  *out = r;
  return PB200_OK;
}

void pb200_destroy(pb200_renderer* r) { delete r; }
const char* pb200_last_error(const pb200_renderer* r) { return r ? r->last_error.c_str() : ""; }

int pb200_upload_buffer(pb200_renderer* r, const float* data, uint64_t frames, uint32_t ch, uint32_t rate,
                        int64_t loop_start, int64_t loop_end, int add_pad_frame, uint32_t* buffer_id) {
  if (!r || !data || !buffer_id) return PB200_ERR_PARAMETER;
  if (rate == 0) return fail(r, PB200_ERR_PARAMETER, "file buffer sample rate must be > 0");
  if (ch == 0) return fail(r, PB200_ERR_PARAMETER, "file buffer channel count must be > 0");
  if (frames == 0) return fail(r, PB200_ERR_PARAMETER, "file buffer must not be empty");
  if (ch > 2) return fail(r, PB200_ERR_UNSUPPORTED, "only mono and stereo buffers are supported");
  auto fb = std::make_shared<AudioFileBuffer>();
  fb->buffer.assign(data, data + frames * ch);
  if (add_pad_frame) fb->buffer.insert(fb->buffer.end(), ch, 0.0f);
  fb->sample_rate = rate;
  fb->channel_count = ch;
  if (loop_start >= 0 && loop_end >= 0) {
    size_t fc = fb->frame_count();
    if (loop_start >= loop_end || (size_t)loop_end > fc) return fail(r, PB200_ERR_PARAMETER, "file buffer loop range is out of bounds");
    fb->has_loop = true; fb->loop_start = (size_t)loop_start; fb->loop_end = (size_t)loop_end;
  }
  r->buffers.push_back(fb);
  *buffer_id = (uint32_t)r->buffers.size() - 1;
  return PB200_OK;
}

int pb200_add_mixer(pb200_renderer* r, uint32_t parent, uint32_t* mixer_id) {
  if (!r || !mixer_id) return PB200_ERR_PARAMETER;
  auto it = r->mixers.find(parent);
  if (it == r->mixers.end()) return fail(r, PB200_ERR_MIXER_NOT_FOUND, "Mixer not found");
  auto proc = std::make_shared<SubMixerProcessor>();
  proc->mixer = std::make_unique<MixedSource>(r->cfg.channel_count, r->cfg.sample_rate);
  uint32_t id = r->next_mixer_id++;
  r->mixers[id] = proc->mixer.get();
  r->mixer_parents[id] = parent;
  MixedSource::Message m; m.kind = MixedSource::Message::AddMixer; m.id = id; m.mixer = proc;
  it->second->message_queue.push_back(std::move(m));
  *mixer_id = id;
  return PB200_OK;
}

int pb200_add_effect(pb200_renderer* r, uint32_t mixer, uint32_t kind, const void* params, size_t size, uint32_t* effect_id) {
  if (!r || !effect_id) return PB200_ERR_PARAMETER;
  auto it = r->mixers.find(mixer);
  if (it == r->mixers.end()) return fail(r, PB200_ERR_MIXER_NOT_FOUND, "Mixer not found");
  std::unique_ptr<Effect> fx;
  switch (kind) {
    case PB200_FX_FILTER:
      if (params) {
        if (size != sizeof(pb200_filter_params)) return fail(r, PB200_ERR_PARAMETER, "bad filter params size");
        auto* p = (const pb200_filter_params*)params;
        if (p->filter_type > 3 || !(p->cutoff >= 20.0f && p->cutoff <= 20000.0f) || !(p->q >= 0.001f && p->q <= 4.0f))
          return fail(r, PB200_ERR_PARAMETER, "Value out of bounds");
        fx = std::make_unique<FilterEffect>(p->filter_type, p->cutoff, p->q);
      } else fx = std::make_unique<FilterEffect>();
      break;
    case PB200_FX_EQ5:
      if (params) return fail(r, PB200_ERR_PARAMETER, "Eq5Effect has no parameter constructor");
      fx = std::make_unique<Eq5Effect>();
      break;
    case PB200_FX_COMPRESSOR:
      if (params) {
        if (size != sizeof(pb200_compressor_params)) return fail(r, PB200_ERR_PARAMETER, "bad compressor params size");
        auto* p = (const pb200_compressor_params*)params;
        fx = std::make_unique<CompressorEffect>(p->threshold, p->ratio, p->knee, p->attack_time, p->release_time, p->makeup_gain, p->lookahead_time);
      } else fx = std::make_unique<CompressorEffect>();
      break;
    case PB200_FX_CHORUS:
      if (params) {
        if (size != sizeof(pb200_chorus_params)) return fail(r, PB200_ERR_PARAMETER, "bad chorus params size");
        auto* p = (const pb200_chorus_params*)params;
        if (p->filter_type > 2) return fail(r, PB200_ERR_PARAMETER, "bad chorus filter type");
        fx = std::make_unique<ChorusEffect>(p->rate, p->phase, p->depth, p->feedback, p->delay, p->wet, p->filter_type, p->filter_freq, p->filter_resonance);
      } else fx = std::make_unique<ChorusEffect>();
      break;
    case PB200_FX_DELAY:
      if (params) return fail(r, PB200_ERR_PARAMETER, "DelayEffect has no parameter constructor");
      fx = std::make_unique<DelayEffect>();
      break;
    case PB200_FX_REVERB: {
      if (!params || size != sizeof(pb200_reverb_params)) return fail(r, PB200_ERR_PARAMETER, "reverb needs explicit fpd/vib_phase state");
      auto* p = (const pb200_reverb_params*)params;
      fx = std::make_unique<ReverbEffect>(p->room_size, p->wet, p->fpd, p->vib_phase);
      break;
    }
    case PB200_FX_GAIN:
      if (params) {
        if (size != sizeof(pb200_gain_params)) return fail(r, PB200_ERR_PARAMETER, "bad gain params size");
        auto* p = (const pb200_gain_params*)params;
        if (p->dc_filter_mode > 3) return fail(r, PB200_ERR_PARAMETER, "bad DC filter mode");
        fx = std::make_unique<GainEffect>(p->gain_db, p->dc_filter_mode);
      } else fx = std::make_unique<GainEffect>();
      break;
    case PB200_FX_PANNING:
      if (params) return fail(r, PB200_ERR_PARAMETER, "PanningEffect has no parameter constructor");
      fx = std::make_unique<PanningEffect>();
      break;
    case PB200_FX_GATE:
      if (params) {
        if (size != sizeof(pb200_gate_params)) return fail(r, PB200_ERR_PARAMETER, "bad gate params size");
        auto* p = (const pb200_gate_params*)params;
        if (!(p->threshold >= -60.0f && p->threshold <= 0.0f) || !(p->attack_time >= 0.001f && p->attack_time <= 0.5f) ||
            !(p->hold_time >= 0.0f && p->hold_time <= 2.0f) || !(p->release_time >= 0.01f && p->release_time <= 2.0f) ||
            !(p->range >= -60.0f && p->range <= 0.0f))
          return fail(r, PB200_ERR_PARAMETER, "Value out of bounds");
        fx = std::make_unique<GateEffect>(p->threshold, p->attack_time, p->hold_time, p->release_time, p->range);
      } else fx = std::make_unique<GateEffect>();
      break;
    case PB200_FX_DISTORTION:
      if (params) {
        if (size != sizeof(pb200_distortion_params)) return fail(r, PB200_ERR_PARAMETER, "bad distortion params size");
        auto* p = (const pb200_distortion_params*)params;
        if (p->distortion_type > 4 || !(p->drive >= 0.0f && p->drive <= 4.0f) || !(p->mix >= 0.0f && p->mix <= 1.0f))
          return fail(r, PB200_ERR_PARAMETER, "Value out of bounds");
        fx = std::make_unique<DistortionEffect>(p->distortion_type, p->drive, p->mix);
      } else fx = std::make_unique<DistortionEffect>();
      break;
    default: return fail(r, PB200_ERR_PARAMETER, "unknown effect kind");
  }
  // Player::add_effect: effect.initialize(sr, ch, MAX_MIX_BUFFER_SAMPLES / ch) (player.rs:905-909)
  if (!fx->initialize(r->cfg.sample_rate, r->cfg.channel_count, MixedSource::MAX_MIX_BUFFER_SAMPLES / r->cfg.channel_count))
    return fail(r, PB200_ERR_PARAMETER, "effect initialize failed");
  uint32_t id = r->next_effect_id++;
  MixedSource::Message m; m.kind = MixedSource::Message::AddEffect; m.id = id;
  m.effect = std::make_shared<EffectProcessor>(std::move(fx));
  it->second->message_queue.push_back(std::move(m));
  r->effects[id] = {mixer, kind};
  *effect_id = id;
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

int pb200_play_file(pb200_renderer* r, uint32_t buffer_id, const pb200_file_options* o, uint64_t start_time, uint32_t* playback_id) {
  if (!r || !o || !playback_id) return PB200_ERR_PARAMETER;
  if (buffer_id >= r->buffers.size()) return fail(r, PB200_ERR_PARAMETER, "unknown buffer");
  if (int e = validate_vol_pan(r, o->volume, o->panning)) return e;
  if (o->speed < 0.0 || std::isnan(o->speed) || std::isinf(o->speed)) return fail(r, PB200_ERR_PARAMETER, "playback options 'speed' value is invalid");
  if (o->resampling_quality > 1) return fail(r, PB200_ERR_PARAMETER, "unknown resampling quality");
  auto mit = r->mixers.find(o->target_mixer);
  if (mit == r->mixers.end()) return fail(r, PB200_ERR_MIXER_NOT_FOUND, "Mixer not found");
  FilePlaybackOptions fo;
  fo.volume = o->volume; fo.panning = o->panning; fo.speed = o->speed;
  if (o->repeat != PB200_REPEAT_DEFAULT) { fo.has_repeat = true; fo.repeat = o->repeat == PB200_REPEAT_FOREVER ? USIZE_MAX : (size_t)o->repeat; }
  if (o->loop_start >= 0 && o->loop_end >= 0) { fo.has_loop_range = true; fo.loop_start = (uint64_t)o->loop_start; fo.loop_end = (uint64_t)o->loop_end; }
  fo.has_fade_in = o->fade_in_nanos != PB200_DURATION_NONE; if (fo.has_fade_in) fo.fade_in = Duration::from_nanos(o->fade_in_nanos);
  fo.has_fade_out = o->fade_out_nanos != PB200_DURATION_NONE; if (fo.has_fade_out) fo.fade_out = Duration::from_nanos(o->fade_out_nanos);
  fo.resampling_quality = o->resampling_quality;
  auto fs = std::make_unique<PreloadedFileSource>(r->buffers[buffer_id], fo, r->cfg.sample_rate);
  pb200_renderer::SourceRef ref;
  ref.mixer = o->target_mixer;
  ref.file = fs.get();
  fs->status_sink = &r->raw_status; fs->pos_emit_rate = r->cfg.sample_rate;  // FilePlaybackOptions::default(): 1 s (file.rs:110)
  ref.queues.file = fs->queue;
  // ConvertedSource (converted.rs:15-46): file already runs at the output rate -> channel mapping only
  std::unique_ptr<Source> src = std::move(fs);
  if (src->channel_count() != r->cfg.channel_count) src = std::make_unique<ChannelMappedSource>(std::move(src), r->cfg.channel_count);
  auto amp = std::make_unique<AmplifiedSource>(std::move(src), o->volume);
  ref.queues.volume = amp->queue;
  auto pan = std::make_unique<PannedSource>(std::move(amp), o->panning);
  ref.queues.panning = pan->queue;
  uint32_t id = r->next_source_id++;
  ref.file->status_id = id;
  r->status_file_format[id] = {(uint32_t)r->buffers[buffer_id]->channel_count, r->buffers[buffer_id]->sample_rate};
  auto ps = std::make_shared<MixedSource::PlayingSource>();
  ps->is_transient = true; ps->playback_id = id; ps->queues = ref.queues; ps->source = std::move(pan);
  ps->start_time = start_time == PB200_TIME_NOW ? 0 : start_time;
  MixedSource::Message m; m.kind = MixedSource::Message::AddSource; m.source = ps; m.sample_time = ps->start_time;
  mit->second->message_queue.push_back(std::move(m));
  r->sources[id] = ref;
  *playback_id = id;
  return PB200_OK;
}

void pb200_sampler_options_default(pb200_sampler_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->volume = 1.0f; o->panning = 0.0f; o->voices = 8; o->target_mixer = PB200_MAIN_MIXER; o->transient = 0; o->has_ahdsr = 0;
  // AhdsrParameters::default() (ahdsr.rs:348-359)
  o->ahdsr.attack_nanos = 10000000ull; o->ahdsr.hold_nanos = 1000000000ull; o->ahdsr.decay_nanos = 500000000ull;
  o->ahdsr.release_nanos = 1000000000ull; o->ahdsr.sustain_level = 0.75f;
}

int pb200_add_sampler(pb200_renderer* r, uint32_t buffer_id, const pb200_sampler_options* o, uint64_t start_time, uint32_t* generator_id) {
  if (!r || !o || !generator_id) return PB200_ERR_PARAMETER;
  if (buffer_id >= r->buffers.size()) return fail(r, PB200_ERR_PARAMETER, "unknown buffer");
  if (int e = validate_vol_pan(r, o->volume, o->panning)) return e;
  if (o->voices == 0) return fail(r, PB200_ERR_PARAMETER, "playback options voice count is '0'");
  auto mit = r->mixers.find(o->target_mixer);
  if (mit == r->mixers.end()) return fail(r, PB200_ERR_MIXER_NOT_FOUND, "Mixer not found");
  auto sampler = std::make_unique<Sampler>(r->buffers[buffer_id], o->voices, r->cfg.channel_count, r->cfg.sample_rate);
  if (o->has_ahdsr) {
    AhdsrParameters p;
    const auto& a = o->ahdsr;
    if (!AhdsrParameters::create(p, Duration::from_nanos(a.attack_nanos), a.attack_scaling, Duration::from_nanos(a.hold_nanos),
                                 Duration::from_nanos(a.decay_nanos), a.decay_scaling, a.sustain_level,
                                 Duration::from_nanos(a.release_nanos), a.release_scaling))
      return fail(r, PB200_ERR_PARAMETER, "Invalid AHDSR parameters");
    if (!sampler->with_ahdsr(p)) return fail(r, PB200_ERR_PARAMETER, "Failed to initialize AHDSR parameters");
  }
  if (o->has_granular) {
    const auto& g = o->granular;
    if (g.overlap_mode > 1 || g.window >= GW_COUNT || g.playback_direction > 2) return fail(r, PB200_ERR_PARAMETER, "Invalid granular parameters");
    if (g.variation != 0.0f || g.spray != 0.0f || g.pan_spread != 0.0f || g.playback_direction == 2)
      return fail(r, PB200_ERR_UNSUPPORTED, "OS-seeded grain randomisation is not reproducible");
    GranularParameters gp;
    gp.overlap_mode = (GrainOverlapMode)g.overlap_mode; gp.window = g.window; gp.size = g.size; gp.density = g.density;
    gp.playback_direction = (GrainPlaybackDirection)g.playback_direction; gp.position = g.position; gp.step = g.step;
    if (!sampler->with_granular_playback(gp)) return fail(r, PB200_ERR_PARAMETER, "Invalid granular parameters");
  }
  sampler->transient = o->transient != 0;  // set_is_transient (player.rs:1062)
  pb200_renderer::SourceRef ref;
  ref.mixer = o->target_mixer; ref.sampler = sampler.get(); ref.transient = o->transient != 0;
  ref.queues.generator = sampler->queue;
  auto amp = std::make_unique<AmplifiedSource>(std::move(sampler), o->volume);
  ref.queues.volume = amp->queue;
  auto pan = std::make_unique<PannedSource>(std::move(amp), o->panning);
  ref.queues.panning = pan->queue;
  uint32_t id = r->next_source_id++;
  auto ps = std::make_shared<MixedSource::PlayingSource>();
  ps->is_transient = o->transient != 0; ps->playback_id = id; ps->queues = ref.queues; ps->source = std::move(pan);
  ps->start_time = start_time == PB200_TIME_NOW ? 0 : start_time;
  MixedSource::Message m; m.kind = MixedSource::Message::AddSource; m.source = ps; m.sample_time = ps->start_time;
  mit->second->message_queue.push_back(std::move(m));
  r->sources[id] = ref;
  *generator_id = id;
  return PB200_OK;
}

int pb200_schedule(pb200_renderer* r, pb200_event* ev) {
  if (!r || !ev) return PB200_ERR_PARAMETER;
  const bool now = ev->sample_time == PB200_TIME_NOW;
  if (ev->kind == PB200_EV_EFFECT_MESSAGE) {  // EffectHandle::send_message (handles/effect.rs:127-163)
    auto it = r->effects.find(ev->target);
    if (it == r->effects.end()) return fail(r, PB200_ERR_EFFECT_NOT_FOUND, "Effect not found");
    if (ev->param_id != PB200_MSG_REVERB_RESET || it->second.kind != PB200_FX_REVERB)
      return fail(r, PB200_ERR_PARAMETER, "Invalid message for this effect");
    MixedSource::Message m; m.kind = MixedSource::Message::Event;
    m.event.kind = MixerEvent::EffectMessage; m.event.target = ev->target;
    m.event.sample_time = now ? 0 : ev->sample_time;
    m.event.param_id = ev->param_id;
    r->mixers[it->second.mixer]->message_queue.push_back(std::move(m));
    return PB200_OK;
  }
  if (ev->kind == PB200_EV_SET_EFFECT_PARAMETER) {
    auto it = r->effects.find(ev->target);
    if (it == r->effects.end()) return fail(r, PB200_ERR_EFFECT_NOT_FOUND, "Effect not found");
    MixedSource::Message m; m.kind = MixedSource::Message::Event;
    m.event.kind = MixerEvent::EffectParameter; m.event.target = ev->target;
    m.event.sample_time = now ? 0 : ev->sample_time;  // handles/effect.rs:80: None => 0
    if ((ev->flags & PB200_EVF_NORMALIZED) && !(ev->value >= 0.0f && ev->value <= 1.0f))  // handles/effect.rs:72-77
      return fail(r, PB200_ERR_PARAMETER, "Invalid parameter update: value should be a normalized value");
    m.event.param_id = ev->param_id;
    m.event.param = {ev->value, (ev->flags & PB200_EVF_NORMALIZED) != 0};
    r->mixers[it->second.mixer]->message_queue.push_back(std::move(m));
    return PB200_OK;
  }
  auto sit = r->sources.find(ev->target);
  if (sit == r->sources.end()) return fail(r, PB200_ERR_SOURCE_NOT_PLAYING, "Source is no longer playing");
  auto& ref = sit->second;
  MixedSource* mixer = r->mixers[ref.mixer];
  auto push_event = [&](MixerEvent e) {
    e.target = ev->target; e.sample_time = ev->sample_time;
    MixedSource::Message m; m.kind = MixedSource::Message::Event; m.event = e;
    mixer->message_queue.push_back(std::move(m));
  };
  const bool has_glide = ev->glide > 0.0f;
  switch (ev->kind) {
    case PB200_EV_STOP_SOURCE:
      if (now) {
        if (ref.queues.file) { FileMsg m{FileMsg::Stop}; ref.queues.file->force_push(m); }
        else { GenMsg m; m.is_stop = true; ref.queues.generator->force_push(m); }
      } else {
        MixedSource::Message m; m.kind = MixedSource::Message::StopSource; m.id = ev->target; m.sample_time = ev->sample_time;
        mixer->message_queue.push_back(std::move(m));
      }
      return PB200_OK;
    case PB200_EV_SET_SOURCE_VOLUME:
      if (now) ref.queues.volume->force_push(ev->value);
      else { MixerEvent e; e.kind = MixerEvent::SetSourceVolume; e.value = ev->value; push_event(e); }
      return PB200_OK;
    case PB200_EV_SET_SOURCE_PANNING:
      if (now) ref.queues.panning->force_push(ev->value);
      else { MixerEvent e; e.kind = MixerEvent::SetSourcePanning; e.value = ev->value; push_event(e); }
      return PB200_OK;
    case PB200_EV_SET_SOURCE_SPEED:
      if (!ref.queues.file) return fail(r, PB200_ERR_PARAMETER, "set_speed needs a file source");
      // HighQuality: rubato is built with max_resample_ratio_relative = 1.0 (rubato.rs:37), so any other output rate
      // makes set_resample_ratio fail and FileSourceImpl::update_speed `expect`-panics (common.rs:166-168)
      if (ref.file && ref.file->hq) {
        auto* rr = static_cast<RubatoResampler*>(ref.file->resampler_box.get());
        uint32_t new_rate = f64_as_u32((double)r->cfg.sample_rate / ev->speed);
        if (has_glide || (double)new_rate / (double)rr->input_rate != rr->resampler.resample_ratio_original)
          return fail(r, PB200_ERR_RESAMPLING, "HighQuality file sources cannot change speed");
      }
      if (now) { FileMsg m{FileMsg::SetSpeed}; m.speed = ev->speed; m.has_glide = has_glide; m.glide = ev->glide; if (!ref.queues.file->push(m)) return fail(r, PB200_ERR_SEND, "File playback queue is full"); }
      else { MixerEvent e; e.kind = MixerEvent::SetSourceSpeed; e.speed = ev->speed; e.has_glide = has_glide; e.glide = ev->glide; push_event(e); }
      return PB200_OK;
    case PB200_EV_SEEK_SOURCE:
      if (!ref.queues.file) return fail(r, PB200_ERR_PARAMETER, "seek needs a file source");
      if (now) { FileMsg m{FileMsg::Seek}; m.position = Duration::from_nanos(ev->position_nanos); if (!ref.queues.file->push(m)) return fail(r, PB200_ERR_SEND, "File playback queue is full"); }
      else { MixerEvent e; e.kind = MixerEvent::SeekSource; e.position = Duration::from_nanos(ev->position_nanos); push_event(e); }
      return PB200_OK;
    default: break;
  }
  if (!ref.queues.generator) return fail(r, PB200_ERR_GENERATOR_NOT_FOUND, "Generator not found");
  GenEvent g;
  switch (ev->kind) {
    case PB200_EV_NOTE_ON:
      g.kind = GenEvent::NoteOn; ev->note_id = r->next_note_id++; g.note_id = ev->note_id; g.note = (uint8_t)ev->note;
      g.has_volume = (ev->flags & PB200_EVF_HAS_VOLUME) != 0; g.volume = ev->value;
      g.has_panning = (ev->flags & PB200_EVF_HAS_PANNING) != 0; g.panning = ev->value2;
      break;
    case PB200_EV_NOTE_OFF: g.kind = GenEvent::NoteOff; g.note_id = ev->note_id; break;
    case PB200_EV_ALL_NOTES_OFF: g.kind = GenEvent::AllNotesOff; break;
    case PB200_EV_SET_NOTE_SPEED: g.kind = GenEvent::SetSpeed; g.note_id = ev->note_id; g.speed = ev->speed; g.has_glide = has_glide; g.glide = ev->glide; break;
    case PB200_EV_SET_NOTE_VOLUME: g.kind = GenEvent::SetVolume; g.note_id = ev->note_id; g.volume = ev->value; break;
    case PB200_EV_SET_NOTE_PANNING: g.kind = GenEvent::SetPanning; g.note_id = ev->note_id; g.panning = ev->value; break;
    case PB200_EV_SET_GENERATOR_PARAMETER:  // GeneratorPlaybackHandle::set_parameter: the id must be one of the generator's parameters
      if (!Sampler::is_parameter(ev->param_id, ref.sampler->envelope_parameters.has_value()))
        return fail(r, PB200_ERR_PARAMETER, "Invalid or unknown sampler parameter");
      if (std::isnan(ev->value)) return fail(r, PB200_ERR_PARAMETER, "Invalid parameter value");
      if ((ev->flags & PB200_EVF_NORMALIZED) && !(ev->value >= 0.0f && ev->value <= 1.0f))
        return fail(r, PB200_ERR_PARAMETER, "Invalid parameter update: value should be a normalized value");
      g.kind = GenEvent::SetParameter; g.param_id = ev->param_id; g.param_value = ev->value; g.param_normalized = (ev->flags & PB200_EVF_NORMALIZED) != 0;
      break;
    case PB200_EV_SET_GENERATOR_LOOP_RANGE:
      g.kind = GenEvent::SetLoopRange; g.has_range = !(ev->flags & PB200_EVF_NO_RANGE); g.range_start = ev->position_nanos; g.range_end = ev->note_id;
      if (ref.sampler->granular_parameters) return fail(r, PB200_ERR_UNSUPPORTED, "loop range messages to granular samplers");
      if (g.has_range) {
        const uint64_t fc = ref.sampler->file_buffer->frame_count();
        if (g.range_start >= g.range_end || g.range_start >= fc || g.range_end > fc) return fail(r, PB200_ERR_PARAMETER, "Invalid loop range");
      }
      break;
    default: return fail(r, PB200_ERR_PARAMETER, "unknown event kind");
  }
  if (now) {
    GenMsg m; m.event = g;
    if (!ref.queues.generator->push(m)) return fail(r, PB200_ERR_SEND, "Generator playback queue is full");
  } else {
    MixerEvent e; e.kind = MixerEvent::TriggerGenerator; e.gen = g; push_event(e);
  }
  return PB200_OK;
}

// ---- structural messages (processed by MixedSource::process_messages at the next block start) -------------------------
int pb200_remove_source(pb200_renderer* r, uint32_t playback_id) {  // Player::remove_generator (player.rs:747-770)
  if (!r) return PB200_ERR_PARAMETER;
  auto it = r->sources.find(playback_id);
  if (it == r->sources.end()) return fail(r, PB200_ERR_GENERATOR_NOT_FOUND, "Generator not found");
  MixedSource::Message m; m.kind = MixedSource::Message::RemoveSource; m.id = playback_id;
  r->mixers[it->second.mixer]->message_queue.push_back(std::move(m));
  r->sources.erase(it);
  return PB200_OK;
}

int pb200_remove_mixer(pb200_renderer* r, uint32_t mixer_id) {  // Player::remove_mixer (player.rs:825-868)
  if (!r) return PB200_ERR_PARAMETER;
  if (mixer_id == PB200_MAIN_MIXER) return fail(r, PB200_ERR_PARAMETER, "Cannot remove the main mixer");
  auto pit = r->mixer_parents.find(mixer_id);
  if (pit == r->mixer_parents.end() || !r->mixers.count(mixer_id)) return fail(r, PB200_ERR_MIXER_NOT_FOUND, "Mixer not found");
  MixedSource::Message m; m.kind = MixedSource::Message::RemoveMixer; m.id = mixer_id;
  r->mixers[pit->second]->message_queue.push_back(std::move(m));
  // the removed subtree is gone for the player too (the reference drops the mixer's effects from its tracking maps;
  // its sources and sub-mixers die with the mixer object)
  std::vector<uint32_t> gone{mixer_id};
  for (size_t i = 0; i < gone.size(); ++i)
    for (auto& kv : r->mixer_parents) if (kv.second == gone[i]) gone.push_back(kv.first);
  for (uint32_t g : gone) {
    for (auto it = r->effects.begin(); it != r->effects.end();) { if (it->second.mixer == g) it = r->effects.erase(it); else ++it; }
    for (auto it = r->sources.begin(); it != r->sources.end();) { if (it->second.mixer == g) it = r->sources.erase(it); else ++it; }
    r->mixers.erase(g);
    r->mixer_parents.erase(g);
  }
  return PB200_OK;
}

int pb200_remove_effect(pb200_renderer* r, uint32_t effect_id) {  // Player::remove_effect (player.rs:977-991)
  if (!r) return PB200_ERR_PARAMETER;
  auto it = r->effects.find(effect_id);
  if (it == r->effects.end()) return fail(r, PB200_ERR_EFFECT_NOT_FOUND, "Effect not found");
  MixedSource::Message m; m.kind = MixedSource::Message::RemoveEffect; m.id = effect_id;
  r->mixers[it->second.mixer]->message_queue.push_back(std::move(m));
  r->effects.erase(it);
  return PB200_OK;
}

int pb200_move_effect(pb200_renderer* r, uint32_t effect_id, uint32_t mixer_id, uint32_t movement, int32_t offset) {  // player.rs:942-974
  if (!r) return PB200_ERR_PARAMETER;
  auto it = r->effects.find(effect_id);
  if (it == r->effects.end()) return fail(r, PB200_ERR_EFFECT_NOT_FOUND, "Effect not found");
  if (it->second.mixer != mixer_id) return fail(r, PB200_ERR_PARAMETER, "Effect does not belong to this mixer");
  if (movement > PB200_MOVE_END) return fail(r, PB200_ERR_PARAMETER, "unknown effect movement");
  MixedSource::Message m; m.kind = MixedSource::Message::MoveEffect; m.id = effect_id; m.movement = movement; m.offset = offset;
  r->mixers[mixer_id]->message_queue.push_back(std::move(m));
  return PB200_OK;
}

int pb200_stop_all_sources(pb200_renderer* r) {  // Player::stop_all_sources (player.rs:1012-1045)
  if (!r) return PB200_ERR_PARAMETER;
  for (auto it = r->sources.begin(); it != r->sources.end();) {
    if (it->second.transient) {
      if (it->second.queues.file) { FileMsg m{FileMsg::Stop}; it->second.queues.file->force_push(m); }
      else { GenMsg m; m.is_stop = true; it->second.queues.generator->force_push(m); }
      it = r->sources.erase(it);
    } else ++it;
  }
  for (auto& kv : r->mixers) {
    MixedSource::Message m; m.kind = MixedSource::Message::RemoveAllPendingEvents;
    kv.second->message_queue.push_back(std::move(m));
  }
  return PB200_OK;
}

int pb200_render(pb200_renderer* r, float* out, uint64_t frames, uint64_t* frames_written) {
  if (!r || !out) return PB200_ERR_PARAMETER;
  const uint32_t bf = r->cfg.block_frames, ch = r->cfg.channel_count;
  if (frames % bf != 0) return fail(r, PB200_ERR_PARAMETER, "frames must be a multiple of block_frames");
  auto t0 = std::chrono::steady_clock::now();
  uint64_t done = 0;
  const uint64_t progress_base = r->progress.load(std::memory_order_acquire);
  struct ProgressDone { pb200_renderer* r; uint64_t total; ~ProgressDone() { r->progress.store(total, std::memory_order_release); } } progress_done{r, progress_base + frames};
  while (done < frames && !r->finished) {  // WavStream::process, wav.rs:210-250
    SourceTime time{r->playback_pos / ch};
    size_t written = r->main->write(r->block.data(), r->block.size(), time);
    if (written == 0) { r->finished = true; break; }
    if (r->meter_interval != UINT64_MAX) {  // MeteredSource::write -> AudioLevelState::record (metered.rs:107-148)
      for (size_t i = 0; i + 1 < written; i += 2)
        for (int c = 0; c < 2; ++c) {
          const float x = r->block[i + c];
          if (std::fabs(x) > r->meter_peak_hold[c]) r->meter_peak_hold[c] = std::fabs(x);
          r->meter_sum_square[c] += (double)x * (double)x;
        }
      r->meter_frames += written / 2;
      if (time.pos_in_frames - std::min(time.pos_in_frames, r->meter_clock) >= r->meter_interval) {
        for (int c = 0; c < 2; ++c) {
          r->audio_level.peak[c] = r->meter_peak_hold[c];
          r->audio_level.rms[c] = r->meter_frames ? (float)std::sqrt(r->meter_sum_square[c] / (double)r->meter_frames) : 0.0f;
          r->meter_peak_hold[c] = 0; r->meter_sum_square[c] = 0;
        }
        r->meter_clock = time.pos_in_frames; r->meter_frames = 0;
      }
    }
    apply_smoothed_gain(r->block.data(), written, r->smoothed_volume);
    std::memcpy(out + done * ch, r->block.data(), written * sizeof(float));
    r->playback_pos += r->block.size();
    done += bf;
    r->progress.store(progress_base + done, std::memory_order_release);
  }
  r->main->ext_bus.clear(); r->main->ext_frames = 0;   // (external main-mixer inputs serve one render call)
  if (done < frames) std::memset(out + done * ch, 0, (frames - done) * ch * sizeof(float));
  if (frames_written) *frames_written = done;
  {  // PlaybackStatusEvent stream in a canonical order (frame, playback id, kind)
    auto& raw = r->raw_status;
    std::stable_sort(raw.begin(), raw.end(), [](const PreloadedFileSource::StatusEv& a, const PreloadedFileSource::StatusEv& b) {
      if (a.frame != b.frame) return a.frame < b.frame;
      if (a.id != b.id) return a.id < b.id;
      return a.kind < b.kind;
    });
    for (auto& e : raw) {
      pb200_status_event o;
      std::memset(&o, 0, sizeof(o));
      o.frame = e.frame; o.playback_id = e.id;
      if (e.kind == 0) {
        auto fmt = r->status_file_format[e.id];
        o.kind = PB200_STATUS_POSITION;
        o.position_nanos = (uint64_t)std::nearbyint(((double)(e.pos / fmt.first) / (double)fmt.second) * 1.0e9);
      } else { o.kind = PB200_STATUS_STOPPED; o.exhausted = e.kind == 1; }
      r->status_events.push_back(o);
    }
    raw.clear();
  }
  r->last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
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
  r->meter_interval = interval_nanos == PB200_DURATION_NONE ? UINT64_MAX : (uint64_t)(Duration::from_nanos(interval_nanos).as_secs_f64() * (double)r->cfg.sample_rate);
  return PB200_OK;
}

int pb200_get_audio_level(pb200_renderer* r, pb200_audio_level* out) {
  if (!r || !out) return PB200_ERR_PARAMETER;
  if (r->meter_interval == UINT64_MAX) return fail(r, PB200_ERR_PARAMETER, "metering is off (PlayerConfig::metering_interval is None)");
  *out = r->audio_level;
  return PB200_OK;
}

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

// ---- WAV in / out, written independently of the product's wav_io.h (chunk walk over the FILE, sample conversion per
// format branch): AudioFileBuffer::from_file for RIFF/WAVE (src/source/file/buffer.rs:64-119, decoder.rs:294-330) and
// WavOutput's 32-bit float file (src/output/wav.rs:61-70, 210-250). symphonia / hound are third-party: integer PCM is
// assumed to scale by 1 / 2^(bits-1).
static bool po_read_exact(FILE* f, void* dst, size_t n) { return std::fread(dst, 1, n, f) == n; }
static uint32_t po_le32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

int pb200_decode_wav(const char* path, float** interleaved, pb200_wav_info* info) {
  if (!path || !interleaved || !info) return PB200_ERR_PARAMETER;
  FILE* f = std::fopen(path, "rb");
  if (!f) return PB200_ERR_MEDIA_FILE_NOT_FOUND;
  unsigned char hd[12];
  if (!po_read_exact(f, hd, 12) || std::memcmp(hd, "RIFF", 4) || std::memcmp(hd + 8, "WAVE", 4)) { std::fclose(f); return PB200_ERR_MEDIA_FILE_PROBE; }
  unsigned fmt_tag = 0, channels = 0, bits = 0, align = 0;
  uint32_t rate = 0;
  std::vector<unsigned char> pcm;
  bool have_data = false, have_loop = false;
  uint32_t ls = 0, le = 0;
  unsigned char ck[8];
  while (po_read_exact(f, ck, 8)) {
    const uint32_t sz = po_le32(ck + 4);
    std::vector<unsigned char> body(sz);
    const size_t got = std::fread(body.data(), 1, sz, f);
    body.resize(got);
    if (sz & 1) std::fgetc(f);
    if (!std::memcmp(ck, "fmt ", 4) && got >= 16) {
      fmt_tag = body[0] | (body[1] << 8); channels = body[2] | (body[3] << 8); rate = po_le32(&body[4]);
      align = body[12] | (body[13] << 8); bits = body[14] | (body[15] << 8);
      if (fmt_tag == 0xFFFE && got >= 26) fmt_tag = body[24] | (body[25] << 8);
    } else if (!std::memcmp(ck, "data", 4) && !have_data) {
      pcm.swap(body); have_data = true;
    } else if (!std::memcmp(ck, "smpl", 4) && !have_loop && got >= 60 && po_le32(&body[28]) > 0) {
      ls = po_le32(&body[36 + 8]); le = po_le32(&body[36 + 12]); have_loop = true;
    }
    if (got < sz) break;
  }
  std::fclose(f);
  const bool int_ok = fmt_tag == 1 && (bits == 8 || bits == 16 || bits == 24 || bits == 32);
  const bool flt_ok = fmt_tag == 3 && (bits == 32 || bits == 64);
  if (!have_data || !channels || !rate || !(int_ok || flt_ok) || align != channels * (bits / 8)) return PB200_ERR_MEDIA_FILE_PROBE;
  const size_t frames = pcm.size() / align;
  if (!frames) return PB200_ERR_AUDIO_DECODING;
  float* out = (float*)std::malloc(frames * channels * sizeof(float));
  if (!out) return PB200_ERR_IO;
  const unsigned char* p = pcm.data();
  for (size_t i = 0; i < frames * channels; ++i, p += bits / 8) {
    switch (fmt_tag * 100 + bits) {
      case 108: out[i] = (float)((int)p[0] - 128) / 128.0f; break;
      case 116: out[i] = (float)(int16_t)(p[0] | (p[1] << 8)) / 32768.0f; break;
      case 124: { int32_t x = (int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24) >> 8; out[i] = (float)x / 8388608.0f; break; }
      case 132: out[i] = (float)((double)(int32_t)po_le32(p) / 2147483648.0); break;
      case 332: std::memcpy(&out[i], p, 4); break;
      default: { double d; std::memcpy(&d, p, 8); out[i] = (float)d; break; }
    }
  }
  *interleaved = out;
  info->frames = frames; info->channels = channels; info->sample_rate = rate;
  info->bits_per_sample = bits; info->is_float = fmt_tag == 3;
  info->loop_start = info->loop_end = PB200_NO_LOOP;
  if (have_loop) {
    const uint64_t fc = frames + 1;  // frame count after the pad frame was appended (buffer.rs:103-110)
    const uint64_t a = std::min<uint64_t>(ls, fc), b = std::min<uint64_t>(le, fc);
    if (b > a) { info->loop_start = (int64_t)a; info->loop_end = (int64_t)b; }
  }
  return PB200_OK;
}

void pb200_free(void* p) { std::free(p); }

int pb200_upload_wav(pb200_renderer* r, const char* path, uint32_t* buffer_id, pb200_wav_info* info) {
  if (!r || !path || !buffer_id) return PB200_ERR_PARAMETER;
  float* data = nullptr;
  pb200_wav_info wi;
  if (int e = pb200_decode_wav(path, &data, &wi)) return fail(r, e, "Audio file failed to open / probe / decode");
  if (info) *info = wi;
  const int e = pb200_upload_buffer(r, data, wi.frames, wi.channels, wi.sample_rate, wi.loop_start, wi.loop_end, 1, buffer_id);
  std::free(data);
  return e;
}

int pb200_render(pb200_renderer* r, float* out, uint64_t frames, uint64_t* frames_written);
int pb200_render_to_wav(pb200_renderer* r, const char* path, uint64_t duration_nanos, uint64_t* frames_written) {
  if (!r || !path) return PB200_ERR_PARAMETER;
  const uint32_t bf = r->cfg.block_frames, ch = r->cfg.channel_count, sr = r->cfg.sample_rate;
  uint64_t blocks = 0;  // wav.rs:222: stop once whole-seconds(pos / sr) >= duration
  while ((blocks * bf / sr) * 1000000000ull < duration_nanos) ++blocks;
  std::vector<float> audio((size_t)blocks * bf * ch);
  uint64_t written = 0;
  if (blocks) { if (int e = pb200_render(r, audio.data(), blocks * bf, &written)) return e; }
  if (frames_written) *frames_written = written;
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(r, PB200_ERR_IO, "cannot create the WAV file");
  const uint32_t data_bytes = (uint32_t)(written * ch * 4);
  auto w32 = [&](uint32_t v) { std::fwrite(&v, 4, 1, f); };
  auto w16 = [&](uint16_t v) { std::fwrite(&v, 2, 1, f); };
  std::fwrite("RIFF", 1, 4, f); w32(60 + data_bytes); std::fwrite("WAVEfmt ", 1, 8, f); w32(40);
  w16(0xFFFE); w16((uint16_t)ch); w32(sr); w32(sr * ch * 4); w16((uint16_t)(ch * 4)); w16(32); w16(22); w16(32); w32(ch == 2 ? 3u : 4u);
  const unsigned char guid[16] = {3, 0, 0, 0, 0, 0, 0x10, 0, 0x80, 0, 0, 0xAA, 0, 0x38, 0x9B, 0x71};
  std::fwrite(guid, 1, 16, f);
  std::fwrite("data", 1, 4, f); w32(data_bytes);
  std::fwrite(audio.data(), 4, (size_t)written * ch, f);
  std::fclose(f);
  return PB200_OK;
}

uint64_t pb200_render_progress(const pb200_renderer* r) { return r ? r->progress.load(std::memory_order_acquire) : 0; }

int pb200_set_main_inputs(pb200_renderer* r, const float* const* buses, uint32_t count, uint64_t frames) {
  if (!r || (count && !buses)) return PB200_ERR_PARAMETER;
  if (count > PB200_MAX_MAIN_INPUTS) return fail(r, PB200_ERR_PARAMETER, "too many main-mixer inputs");
  if (count && frames % r->cfg.block_frames != 0) return fail(r, PB200_ERR_PARAMETER, "frames must be a multiple of block_frames");
  r->main->ext_bus.assign(buses, buses + count);
  r->main->ext_start = r->playback_pos / r->cfg.channel_count;
  r->main->ext_frames = count ? frames : 0;
  return PB200_OK;
}
int pb200_set_main_input(pb200_renderer* r, const float* bus, uint64_t frames) { return pb200_set_main_inputs(r, &bus, bus ? 1u : 0u, frames); }
uint64_t pb200_trim_pool(int) { return 0; }
// peer memory: device-only
int pb200_device_alloc(int, size_t, void**) { return PB200_ERR_UNSUPPORTED; }
int pb200_device_free(void*) { return PB200_ERR_UNSUPPORTED; }
int pb200_ipc_export(const void*, void*) { return PB200_ERR_UNSUPPORTED; }
int pb200_ipc_open(const void*, int, void**) { return PB200_ERR_UNSUPPORTED; }
int pb200_ipc_close(void*) { return PB200_ERR_UNSUPPORTED; }
int pb200_push_async(pb200_renderer* r, void*, const void*, size_t, uint32_t*, uint32_t) { return fail(r, PB200_ERR_UNSUPPORTED, "oracle has no device memory"); }
int pb200_push_sync(pb200_renderer* r) { return fail(r, PB200_ERR_UNSUPPORTED, "oracle has no device memory"); }
int pb200_peek_u32(pb200_renderer* r, const uint32_t*, uint32_t, uint32_t*) { return fail(r, PB200_ERR_UNSUPPORTED, "oracle has no device memory"); }

int pb200_render_device(pb200_renderer* r, float*, uint64_t, uint64_t*) { return fail(r, PB200_ERR_UNSUPPORTED, "oracle has no device memory"); }

uint64_t pb200_position(const pb200_renderer* r) { return r ? r->playback_pos / r->cfg.channel_count : 0; }

int pb200_source_status_get(pb200_renderer* r, uint32_t id, pb200_source_status* st) {
  if (!r || !st) return PB200_ERR_PARAMETER;
  auto it = r->sources.find(id);
  if (it == r->sources.end()) return fail(r, PB200_ERR_SOURCE_NOT_PLAYING, "Source is no longer playing");
  std::memset(st, 0, sizeof(*st));
  st->end_frame = UINT64_MAX;
  // NB: the oracle keeps raw pointers into the graph; finished transient sources are dropped from
  // the mixer, so status is tracked lazily while they are alive only.
  MixedSource* mixer = r->mixers[it->second.mixer];
  bool alive = mixer->find_source(id) != nullptr;
  bool queued = false;
  for (auto& m : mixer->message_queue) if (m.kind == MixedSource::Message::AddSource && m.source && m.source->playback_id == id) queued = true;
  st->is_playing = (alive || queued) ? 1 : 0;
  if ((alive || queued) && it->second.file) {
    st->playback_pos = it->second.file->playback_pos;
    st->exhausted = it->second.file->stopped_exhausted;
    st->end_frame = it->second.file->end_frame;
  }
  return PB200_OK;
}

int pb200_sampler_voice_states(pb200_renderer* r, uint32_t id, pb200_voice_state* out, uint32_t capacity, uint32_t* count) {
  if (!r || !out || !count) return PB200_ERR_PARAMETER;
  auto it = r->sources.find(id);
  if (it == r->sources.end() || !it->second.sampler) return fail(r, PB200_ERR_GENERATOR_NOT_FOUND, "Generator not found");
  // a transient generator that got exhausted was dropped (and freed) by its mixer: no voices left
  MixedSource* mixer = r->mixers[it->second.mixer];
  bool alive = mixer->find_source(id) != nullptr;
  for (auto& m : mixer->message_queue) if (m.kind == MixedSource::Message::AddSource && m.source && m.source->playback_id == id) alive = true;
  if (!alive) { *count = 0; return PB200_OK; }
  Sampler* s = it->second.sampler;
  uint32_t n = (uint32_t)std::min<size_t>(capacity, s->voices.size());
  for (uint32_t i = 0; i < n; ++i) {
    const auto& v = s->voices[i];
    out[i].note_id = v.has_note ? v.note_id : UINT64_MAX;
    out[i].playback_pos = v.file->playback_pos;
    out[i].envelope_stage = (uint32_t)v.envelope.stage;
    out[i].active = v.has_note ? 1 : 0;
  }
  *count = (uint32_t)s->voices.size();
  return PB200_OK;
}

int pb200_last_render_stats(pb200_renderer* r, pb200_render_stats* st) {
  if (!r || !st) return PB200_ERR_PARAMETER;
  std::memset(st, 0, sizeof(*st));
  st->device_ms = r->last_ms;
  return PB200_OK;
}

}  // extern "C"
