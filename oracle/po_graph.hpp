// ORACLE -- TEST INFRASTRUCTURE ONLY (see po_dsp.hpp header). CPU restatement of the reference's
// source chain, sampler, mixer graph and WAV block driver. Parity pinning: see po_dsp.hpp.
#pragma once
#include "po_dsp.hpp"
#include "po_granular.hpp"

namespace po {

// SourceTime (src/source.rs:28-33); pos_instant is wall clock and unused offline.
struct SourceTime { uint64_t pos_in_frames = 0; };

// trait Source (src/source.rs:80-110)
struct Source {
  virtual ~Source() {}
  virtual uint32_t sample_rate() const = 0;
  virtual size_t channel_count() const = 0;
  virtual bool is_exhausted() const = 0;
  virtual size_t weight() const = 0;
  virtual size_t write(float* out, size_t len, const SourceTime& time) = 0;
};

// crossbeam_queue::ArrayQueue<T> (bounded; push fails when full, force_push drops the oldest)
template <class T>
struct ArrayQueue {
  size_t cap;
  std::deque<T> q;
  explicit ArrayQueue(size_t c) : cap(c) {}
  bool push(const T& v) { if (q.size() >= cap) return false; q.push_back(v); return true; }
  bool force_push(const T& v) {  // returns true when an element got displaced
    bool displaced = false;
    if (q.size() >= cap) { q.pop_front(); displaced = true; }
    q.push_back(v);
    return displaced;
  }
  bool pop(T& v) { if (q.empty()) return false; v = q.front(); q.pop_front(); return true; }
};

// ---- src/source/file/buffer.rs:14-150 -------------------------------------------------------------
struct AudioFileBuffer {
  std::vector<float> buffer;
  uint32_t sample_rate = 0;
  size_t channel_count = 0;
  bool has_loop = false;
  size_t loop_start = 0, loop_end = 0;  // frames
  size_t frame_count() const { return buffer.size() / channel_count; }
};

// ---- src/source/file.rs:34-84 ---------------------------------------------------------------------
struct FilePlaybackOptions {
  float volume = 1.0f, panning = 0.0f;
  double speed = 1.0;
  bool has_repeat = false; size_t repeat = 0;
  bool has_loop_range = false; uint64_t loop_start = 0, loop_end = 0;
  bool has_fade_in = false; Duration fade_in;
  bool has_fade_out = true; Duration fade_out = Duration::from_millis(50);
  uint32_t resampling_quality = 0;
  uint32_t target_mixer = 0;
};

struct FileMsg {
  enum Kind { Seek, SetSpeed, Stop, Kill } kind;
  Duration position{};
  double speed = 1; bool has_glide = false; float glide = 0;
  FileMsg(Kind k = Stop) : kind(k) {}
};

// ---- src/source/file/preloaded.rs:29-476 + src/source/file/common.rs:31-169 -----------------------
struct PreloadedFileSource : Source {
  std::shared_ptr<AudioFileBuffer> file_buffer;
  // FileSourceImpl
  VolumeFader volume_fader;
  std::unique_ptr<AudioResampler> resampler_box;  // Box<dyn AudioResampler> (file/common.rs:84-87)
  AudioResampler& resampler;
  std::vector<float> resampler_input_buffer;       // TempBuffer of max_input_buffer_size or 256 (common.rs:88-92)
  bool hq;
  bool has_fade_out; Duration fade_out_duration;
  uint32_t output_sample_rate;
  size_t output_channel_count;
  std::shared_ptr<ArrayQueue<FileMsg>> queue;
  bool playback_finished = false;
  size_t samples_to_next_speed_update = 0;
  float speed_glide_rate = 0;
  double current_speed, target_speed;
  // PreloadedFileSource
  size_t playback_repeat, playback_repeat_count, playback_pos = 0;
  bool playback_pos_eof = false;
  bool has_loop_override = false; uint64_t loop_override_start = 0, loop_override_end = 0;
  uint64_t end_frame = UINT64_MAX;  // oracle-side bookkeeping for status queries
  // PlaybackStatusEvent sink (file/common.rs:171-221): null for sampler voices; events are (frame, kind, pos)
  struct StatusEv { uint64_t frame; uint32_t kind; uint64_t pos; uint32_t id; };
  std::vector<StatusEv>* status_sink = nullptr;
  uint32_t status_id = 0; uint64_t pos_emit_rate = 0, pos_clock = 0, msg_frame = 0;
  void send_stopped(uint64_t frame) { if (status_sink) status_sink->push_back({frame, playback_pos_eof ? 1u : 2u, playback_pos, status_id}); }
  bool stopped_exhausted = false;

  static constexpr size_t SPEED_UPDATE_CHUNK_SIZE = 64;

  PreloadedFileSource(std::shared_ptr<AudioFileBuffer> fb, const FilePlaybackOptions& o, uint32_t out_rate)
      : file_buffer(fb),
        volume_fader(fb->channel_count, out_rate),
        resampler_box(make_resampler(o.resampling_quality, fb->sample_rate, f64_as_u32((double)out_rate / o.speed), fb->channel_count)),
        resampler(*resampler_box),
        resampler_input_buffer(resampler.max_input_buffer_size() ? resampler.max_input_buffer_size() : 256, 0.0f),
        hq(o.resampling_quality != 0),
        has_fade_out(o.has_fade_out), fade_out_duration(o.fade_out),
        output_sample_rate(out_rate), output_channel_count(fb->channel_count),
        queue(std::make_shared<ArrayQueue<FileMsg>>(128)),
        current_speed(o.speed), target_speed(o.speed) {
    if (o.has_fade_in && !o.fade_in.is_zero()) volume_fader.start_fade_in(o.fade_in);
    playback_repeat = o.has_repeat ? o.repeat : (fb->has_loop ? USIZE_MAX : 0);
    playback_repeat_count = playback_repeat;
    if (o.has_loop_range) {
      uint64_t fc = fb->frame_count();
      has_loop_override = true;
      loop_override_start = std::min<uint64_t>(o.loop_start, fc > 0 ? fc - 1 : 0);
      loop_override_end = std::min<uint64_t>(o.loop_end, fc);
    }
  }
  static std::unique_ptr<AudioResampler> make_resampler(uint32_t quality, uint32_t in_rate, uint32_t out_rate, size_t cc) {
    if (quality != 0) return std::make_unique<RubatoResampler>(in_rate, out_rate, cc);  // ResamplingQuality::HighQuality
    return std::make_unique<CubicAudioResampler>(in_rate, out_rate, cc);
  }
  uint32_t sample_rate() const override { return output_sample_rate; }
  size_t channel_count() const override { return output_channel_count; }
  bool is_exhausted() const override { return playback_finished; }
  size_t weight() const override { return 1; }

  void update_speed(uint32_t input_sample_rate) {  // common.rs:141-169
    double speed_diff = target_speed - current_speed;
    if (speed_glide_rate > 0.0f && std::fabs(speed_diff) > 0.0001) {
      double semitone_diff = std::fabs(12.0 * std::log2(target_speed / current_speed));
      float duration_secs = (float)semitone_diff / speed_glide_rate;
      if (duration_secs > 0.0f) {
        float duration_frames = duration_secs * (float)output_sample_rate;
        double step_per_frame = (target_speed - current_speed) / (double)duration_frames;
        double change = step_per_frame * (double)SPEED_UPDATE_CHUNK_SIZE;
        if (std::fabs(target_speed - current_speed) < std::fabs(change)) current_speed = target_speed;
        else current_speed += change;
      } else {
        current_speed = target_speed;
      }
    } else {
      current_speed = target_speed;
    }
    uint32_t new_rate = f64_as_u32((double)output_sample_rate / current_speed);
    bool ok = resampler.update(input_sample_rate, new_rate);
    assert(ok && "failed to update resampler specs");  // .expect() in the reference (common.rs:166-168)
    (void)ok;
  }
  void seek(Duration position) {  // preloaded.rs:138-146
    if (!is_exhausted()) {
      double buffer_pos = position.as_secs_f64() * (double)file_buffer->sample_rate * (double)file_buffer->channel_count;
      playback_pos = std::min(f64_as_usize(buffer_pos), file_buffer->buffer.size());
      resampler.reset();
    }
  }
  void set_speed(double speed, bool has_glide, float glide) {  // preloaded.rs:182-193
    if (!is_exhausted()) {
      samples_to_next_speed_update = 0;
      target_speed = speed;
      speed_glide_rate = has_glide ? glide : 0.0f;
      if (speed_glide_rate == 0.0f) {
        current_speed = speed;
        update_speed(file_buffer->sample_rate);
      }
    }
  }
  void stop() {  // preloaded.rs:196-209
    if (!is_exhausted()) {
      if (has_fade_out && !fade_out_duration.is_zero()) volume_fader.start_fade_out(fade_out_duration);
      else { stopped_exhausted = playback_pos_eof; playback_finished = true; send_stopped(msg_frame); }
    }
  }
  void kill() { if (!is_exhausted()) { stopped_exhausted = playback_pos_eof; playback_finished = true; send_stopped(msg_frame); } }
  void reset() {  // preloaded.rs:212-230
    if (!is_exhausted()) kill();
    playback_pos = 0;
    playback_repeat_count = playback_repeat;
    playback_pos_eof = false;
    playback_finished = false;
    resampler.reset();
    volume_fader.reset();
  }
  void set_loop_range(bool has, uint64_t s, uint64_t e) { has_loop_override = has; loop_override_start = s; loop_override_end = e; }
  void set_repeat(size_t r) { playback_repeat = r; playback_repeat_count = r; }
  bool loop_range(size_t& s, size_t& e) const {  // frames, preloaded.rs:150-156
    if (has_loop_override) { s = (size_t)loop_override_start; e = (size_t)loop_override_end; return true; }
    if (file_buffer->has_loop) { s = file_buffer->loop_start; e = file_buffer->loop_end; return true; }
    return false;
  }
  void process_messages() {
    FileMsg m;
    while (queue->pop(m)) {
      switch (m.kind) {
        case FileMsg::Seek: seek(m.position); break;
        case FileMsg::SetSpeed: set_speed(m.speed, m.has_glide, m.glide); break;
        case FileMsg::Stop: stop(); break;
        case FileMsg::Kill: kill(); break;
      }
    }
  }
  size_t write_buffer(float* out, size_t len) {  // preloaded.rs:270-332
    size_t written = 0;
    size_t cc = file_buffer->channel_count;
    size_t ls = 0, le = file_buffer->buffer.size();
    if (playback_repeat > 0) {
      size_t s, e;
      if (loop_range(s, e)) { ls = s * cc; le = e * cc; }
    }
    const size_t required_input_len = resampler.required_input_buffer_size();
    while (written < len) {
      size_t remaining_in = le > playback_pos ? le - playback_pos : 0;
      const float* in = file_buffer->buffer.data() + playback_pos;
      std::pair<size_t, size_t> res;
      if (remaining_in < required_input_len) {
        // pad the input with zeros for fixed-size resamplers; the input counts as consumed whatever
        // process() did with it (preloaded.rs:296-304)
        std::copy(in, in + remaining_in, resampler_input_buffer.begin());
        std::fill(resampler_input_buffer.begin() + remaining_in, resampler_input_buffer.end(), 0.0f);
        auto r2 = resampler.process(resampler_input_buffer.data(), resampler_input_buffer.size(), out + written, len - written);
        res = {remaining_in, r2.second};
      } else {
        res = resampler.process(in, remaining_in, out + written, len - written);
      }
      playback_pos += res.first;
      written += res.second;
      if (playback_pos >= le) {
        if (playback_repeat_count > 0) {
          if (playback_repeat_count != USIZE_MAX) playback_repeat_count -= 1;
          playback_pos = ls;
        } else {
          playback_pos_eof = true;
        }
      }
      if (playback_pos_eof && res.second == 0) break;
    }
    return written;
  }
  size_t write(float* out, size_t len, const SourceTime& time) override {  // preloaded.rs:396-475
    msg_frame = time.pos_in_frames;
    process_messages();
    if (playback_finished) return 0;
    size_t total = 0;
    if (current_speed != target_speed) {
      while (total < len) {
        if (samples_to_next_speed_update == 0) {
          if (current_speed != target_speed) update_speed(file_buffer->sample_rate);
          samples_to_next_speed_update = SPEED_UPDATE_CHUNK_SIZE * output_channel_count;
        }
        size_t chunk = std::min(len - total, samples_to_next_speed_update);
        size_t w = write_buffer(out + total, chunk);
        samples_to_next_speed_update -= w;
        total += w;
        if (w < chunk) break;
      }
    } else {
      samples_to_next_speed_update = 0;
      total = write_buffer(out, len);
    }
    volume_fader.process(out, total);
    // send_playback_position_status (file/common.rs:183-208): once per emit interval, at the end of a write call
    if (status_sink && pos_emit_rate && time.pos_in_frames - std::min(time.pos_in_frames, pos_clock) >= pos_emit_rate) {
      pos_clock = time.pos_in_frames;
      status_sink->push_back({time.pos_in_frames, 0u, playback_pos, status_id});
    }
    bool fade_out_completed = volume_fader.state == VolumeFader::Finished && volume_fader.target_volume == 0.0f;
    if (playback_pos_eof || fade_out_completed) {
      stopped_exhausted = playback_pos_eof;
      playback_finished = true;
      end_frame = time.pos_in_frames + len / output_channel_count;
      send_stopped(end_frame);
    }
    return total;
  }
};

// ---- src/source/mapped.rs:8-100 + src/utils/buffer.rs:183-268 (1<->2 channels only) ----------------
struct ChannelMappedSource : Source {
  std::unique_ptr<Source> source;
  size_t in_ch, out_ch;
  std::vector<float> input_buffer;
  static constexpr size_t MAX_MIX_BUFFER_SAMPLES = 8 * 1024;
  ChannelMappedSource(std::unique_ptr<Source> s, size_t out_channels)
      : source(std::move(s)), in_ch(source->channel_count()), out_ch(out_channels),
        input_buffer(MAX_MIX_BUFFER_SAMPLES / out_channels * in_ch) {}
  uint32_t sample_rate() const override { return source->sample_rate(); }
  size_t channel_count() const override { return out_ch; }
  bool is_exhausted() const override { return source->is_exhausted(); }
  size_t weight() const override { return source->weight(); }
  size_t write(float* out, size_t len, const SourceTime& time) override {
    if (len == 0 || in_ch == out_ch) return source->write(out, len, time);
    size_t total = 0;
    while (total < len) {
      size_t input_max = ((len - total) / out_ch) * in_ch;
      size_t buffer_max = std::min(input_max, input_buffer.size());
      SourceTime st{time.pos_in_frames + total / out_ch};
      size_t written = source->write(input_buffer.data(), buffer_max, st);
      if (written == 0) break;
      size_t frames = written / in_ch;
      float* o = out + total;
      if (in_ch == 1 && out_ch == 2) {
        for (size_t f = 0; f < frames; ++f) { o[2 * f] = input_buffer[f]; o[2 * f + 1] = input_buffer[f]; }
      } else if (in_ch == 2 && out_ch == 1) {
        for (size_t f = 0; f < frames; ++f) o[f] = (input_buffer[2 * f] + input_buffer[2 * f + 1]) / 2.0f;
      } else {
        assert(false && "oracle: only mono<->stereo channel mapping is restated");
      }
      total += frames * out_ch;
    }
    return total;
  }
};

// ---- src/source/amplified.rs:19-104 -----------------------------------------------------------------
struct AmplifiedSource : Source {
  std::unique_ptr<Source> source;
  ExpSmoothed volume;
  std::shared_ptr<ArrayQueue<float>> queue;
  AmplifiedSource(std::unique_ptr<Source> s, float v)
      : source(std::move(s)), volume(v, source->sample_rate()), queue(std::make_shared<ArrayQueue<float>>(1)) {}
  uint32_t sample_rate() const override { return source->sample_rate(); }
  size_t channel_count() const override { return source->channel_count(); }
  bool is_exhausted() const override { return source->is_exhausted(); }
  size_t weight() const override { return source->weight(); }
  void set_volume(float v) { volume.set_target(v); }
  size_t write(float* out, size_t len, const SourceTime& time) override {
    float v;
    while (queue->pop(v)) volume.set_target(v);
    size_t written = source->write(out, len, time);
    apply_smoothed_gain(out, written, volume);
    return written;
  }
};

// ---- src/source/panned.rs:19-104 --------------------------------------------------------------------
struct PannedSource : Source {
  std::unique_ptr<Source> source;
  ExpSmoothed panning;
  std::shared_ptr<ArrayQueue<float>> queue;
  PannedSource(std::unique_ptr<Source> s, float p)
      : source(std::move(s)), panning(p, source->sample_rate()), queue(std::make_shared<ArrayQueue<float>>(1)) {}
  uint32_t sample_rate() const override { return source->sample_rate(); }
  size_t channel_count() const override { return source->channel_count(); }
  bool is_exhausted() const override { return source->is_exhausted(); }
  size_t weight() const override { return source->weight(); }
  void set_panning(float p) { panning.set_target(p); }
  size_t write(float* out, size_t len, const SourceTime& time) override {
    float p;
    while (queue->pop(p)) panning.set_target(p);
    size_t written = source->write(out, len, time);
    apply_smoothed_panning(out, written, source->channel_count(), panning);
    return written;
  }
};

// ---- GeneratorPlaybackEvent / GeneratorPlaybackMessage (src/generator.rs:172-239) -------------------
struct GenEvent {
  enum Kind { NoteOn, NoteOff, AllNotesOff, SetSpeed, SetVolume, SetPanning, SetParameter, SetLoopRange } kind;
  uint32_t param_id = 0; float param_value = 0; bool param_normalized = false;  // SetParameter (generator.rs:172-226)
  bool has_range = false; uint64_t range_start = 0, range_end = 0;                 // ProcessMessage(SamplerMessage::SetLoopRange)
  uint64_t note_id = 0;
  uint8_t note = 60;
  bool has_volume = false; float volume = 1;
  bool has_panning = false; float panning = 0;
  double speed = 1; bool has_glide = false; float glide = 0;
};
struct GenMsg { bool is_stop = false; GenEvent event; };

// ---- src/generator/sampler/voice.rs:37-528 ------------------------------------------------------------
struct SamplerVoice {
  bool has_note = false; uint64_t note_id = 0;
  uint8_t note = 60;
  float note_volume = 1, note_panning = 0;
  // Panned<Amplified<ChannelMapped<PreloadedFileSource>>>
  std::unique_ptr<PannedSource> source;
  AmplifiedSource* amplified = nullptr;
  PreloadedFileSource* file = nullptr;
  AhdsrEnvelope envelope;
  bool has_release_start = false; uint64_t release_start_frame = 0;
  std::unique_ptr<GrainPool> grain_pool;  // voice.rs:46 (enable_granular_playback :342-380)
  static constexpr size_t MODULATION_PROCESSOR_BLOCK_SIZE = 64;  // src/modulation/processor.rs

  void enable_granular_playback(uint32_t sample_rate, std::shared_ptr<std::vector<float>> sample_buffer) {
    const AudioFileBuffer& fb = *file->file_buffer;
    bool has_loop = fb.has_loop;
    float ls = 0, le = 0;
    if (has_loop) { float total = (float)fb.frame_count(); ls = (float)fb.loop_start / total; le = (float)fb.loop_end / total; }
    grain_pool = std::make_unique<GrainPool>(sample_rate, sample_buffer, has_loop, ls, le);
  }

  SamplerVoice(std::unique_ptr<PreloadedFileSource> fs, size_t channel_count) {
    file = fs.get();
    auto mapped = std::make_unique<ChannelMappedSource>(std::move(fs), channel_count);
    auto amp = std::make_unique<AmplifiedSource>(std::move(mapped), 1.0f);
    amplified = amp.get();
    source = std::make_unique<PannedSource>(std::move(amp), 0.0f);
  }
  bool is_active() const { return has_note; }
  bool in_release_stage() const { return envelope.stage == AhdsrEnvelope::Release; }
  void reset() {  // voice.rs:222-236
    if (is_active()) { file->reset(); has_note = false; if (grain_pool) grain_pool->reset(); }
    has_release_start = false;
  }
  void start(uint64_t id, uint8_t n, float volume, float panning, int32_t base_transpose, int32_t base_finetune,
             float base_volume, float base_panning, const std::optional<AhdsrParameters>& env,
             const std::optional<GranularParameters>& gran) {  // voice.rs:122-193
    reset();
    note = n; note_volume = volume; note_panning = panning;
    double note_speed = speed_from_note(n);
    double pitch_factor = std::pow(2.0, (double)base_transpose / 12.0 + (double)base_finetune / 1200.0);
    double effective_speed = note_speed * pitch_factor;
    float effective_volume = base_volume * volume;
    float effective_panning = std::min(std::max(base_panning + panning, -1.0f), 1.0f);
    file->set_speed(effective_speed, false, 0.0f);
    amplified->set_volume(effective_volume);
    source->set_panning(effective_panning);
    if (grain_pool && gran) grain_pool->start(*gran, effective_speed, effective_volume, effective_panning);
    if (env) envelope.note_on(*env, 1.0f);
    has_note = true; note_id = id;
  }
  void stop(const std::optional<AhdsrParameters>& env, uint64_t current_frame) {  // voice.rs:196-219
    if (is_active()) {
      has_release_start = true; release_start_frame = current_frame;
      if (env) envelope.note_off(*env);
      else { file->stop(); if (grain_pool) grain_pool->stop(); }
    }
  }
  void set_speed(double speed, bool has_glide, float glide, int32_t base_transpose, int32_t base_finetune) {
    double pitch_factor = std::pow(2.0, (double)base_transpose / 12.0 + (double)base_finetune / 1200.0);
    file->set_speed(speed * pitch_factor, has_glide, glide);
    if (grain_pool) grain_pool->speed = speed * pitch_factor;
  }
  void set_base_pitch(int32_t base_transpose, int32_t base_finetune) {  // voice.rs:258-268
    double note_speed = speed_from_note(note);
    double pitch_factor = std::pow(2.0, (double)base_transpose / 12.0 + (double)base_finetune / 1200.0);
    double effective_speed = note_speed * pitch_factor;
    file->set_speed(effective_speed, false, 0.0f);
    if (grain_pool) grain_pool->speed = effective_speed;
  }
  void set_base_volume(float base_volume) {  // voice.rs:283-289
    amplified->set_volume(base_volume * note_volume);
    if (grain_pool) grain_pool->volume = base_volume * note_volume;
  }
  void set_base_panning(float base_panning) {  // voice.rs:304-310
    float eff = std::min(std::max(base_panning + note_panning, -1.0f), 1.0f);
    source->set_panning(eff);
    if (grain_pool) grain_pool->panning = eff;
  }
  void set_loop_range(bool has, uint64_t s, uint64_t e) {  // voice.rs:313-339
    file->set_loop_range(has, s, e);
    file->set_repeat(has ? USIZE_MAX : 0);
    if (grain_pool) {
      float total = (float)file->file_buffer->frame_count();
      grain_pool->has_loop_range = has;  // GrainPool::set_loop_range (granular.rs:516-518)
      grain_pool->loop_start = has ? (float)s / total : 0.0f; grain_pool->loop_end = has ? (float)e / total : 0.0f;
    }
  }
  void set_volume(float v, float base_volume) {
    note_volume = v; amplified->set_volume(base_volume * v);
    if (grain_pool) grain_pool->volume = base_volume * v;
  }
  void set_panning(float p, float base_panning) {
    note_panning = p;
    float eff = std::min(std::max(base_panning + p, -1.0f), 1.0f);
    source->set_panning(eff);
    if (grain_pool) grain_pool->panning = eff;
  }
  size_t process(float* out, size_t len, size_t channel_count, const std::optional<AhdsrParameters>& env,
                 const std::optional<GranularParameters>& gran, const SourceTime& time) {  // voice.rs:390-505
    size_t written;
    if (grain_pool && gran) {  // grain playback: chunks of MODULATION_PROCESSOR_BLOCK_SIZE frames (voice.rs:412-427)
      const size_t step = MODULATION_PROCESSOR_BLOCK_SIZE * channel_count;
      for (size_t o = 0; o < len; o += step) grain_pool->process(out + o, std::min(step, len - o), channel_count, *gran);
      written = len;
    } else {
      written = source->write(out, len, time);
    }
    if (env) {
      if (envelope.stage == AhdsrEnvelope::Sustain || envelope.stage == AhdsrEnvelope::Idle) {
        scale_buffer(out, written, envelope.output);
      } else {
        for (size_t i = 0; i + channel_count <= written; i += channel_count) {
          float e = envelope.run(*env);
          for (size_t c = 0; c < channel_count; ++c) out[i + c] *= e;
        }
      }
    }
    if (source->is_exhausted() || (grain_pool && grain_pool->is_exhausted()) || (env && envelope.stage == AhdsrEnvelope::Idle)) reset();
    return written;
  }
};

// ---- src/generator/sampler.rs:72-1028 -----------------------------------------------------------------
struct Sampler : Source {
  std::shared_ptr<ArrayQueue<GenMsg>> queue;
  size_t active_voices = 0;
  std::vector<SamplerVoice> voices;
  int32_t base_transpose = 0, base_finetune = 0;
  float base_volume = 1, base_panning = 0;
  std::optional<AhdsrParameters> envelope_parameters;
  std::optional<GranularParameters> granular_parameters;
  std::shared_ptr<AudioFileBuffer> file_buffer;
  bool transient = false, stopping = false, stopped = false;
  uint32_t output_sample_rate;
  size_t output_channel_count;
  std::vector<float> temp_buffer;

  Sampler(std::shared_ptr<AudioFileBuffer> fb, size_t voice_count, size_t out_ch, uint32_t out_rate)
      : queue(std::make_shared<ArrayQueue<GenMsg>>((4 + 5 + 10) * 2 + 16)),
        file_buffer(fb), output_sample_rate(out_rate), output_channel_count(out_ch), temp_buffer(8 * 1024) {
    FilePlaybackOptions vo;  // sampler.rs:509-514
    vo.has_fade_out = true; vo.fade_out = Duration::from_millis(50);
    voices.reserve(voice_count);
    for (size_t i = 0; i < voice_count; ++i)
      voices.emplace_back(std::make_unique<PreloadedFileSource>(fb, vo, out_rate), out_ch);
  }
  bool with_ahdsr(AhdsrParameters p) {
    if (!p.set_sample_rate(output_sample_rate)) return false;
    envelope_parameters = p;
    return true;
  }
  // Sampler::create_granular_sample_buffer (sampler.rs:908-952): the file as a mono buffer at the output rate
  static std::shared_ptr<std::vector<float>> create_granular_sample_buffer(std::shared_ptr<AudioFileBuffer> fb, uint32_t out_rate) {
    auto dest = std::make_shared<std::vector<float>>();
    if (fb->channel_count == 1 && fb->sample_rate == out_rate) { *dest = fb->buffer; return dest; }
    FilePlaybackOptions o;  // default().resampling_quality(Default).repeat(0)
    o.has_repeat = true; o.repeat = 0;
    PreloadedFileSource source(fb, o, out_rate);
    size_t cc = source.channel_count();
    std::vector<float> temp(1024 * cc, 0.0f);
    SourceTime time;
    for (;;) {
      size_t read = source.write(temp.data(), temp.size(), time);
      if (read == 0) break;
      for (size_t f = 0; f + cc <= read; f += cc) {
        float sum = 0.0f;
        for (size_t c = 0; c < cc; ++c) sum += temp[f + c];
        dest->push_back(sum / (float)cc);
      }
      time.pos_in_frames += read / cc;
    }
    if (dest->empty()) dest->push_back(0.0f);
    return dest;
  }
  bool with_granular_playback(const GranularParameters& p) {  // sampler.rs:599-637
    if (!p.validate()) return false;
    auto buf = create_granular_sample_buffer(file_buffer, output_sample_rate);
    for (auto& v : voices) v.enable_granular_playback(output_sample_rate, buf);
    granular_parameters = p;
    return true;
  }
  uint32_t sample_rate() const override { return output_sample_rate; }
  size_t channel_count() const override { return output_channel_count; }
  bool is_exhausted() const override { return stopped; }
  size_t weight() const override { return std::max<size_t>(active_voices, 1); }

  size_t next_free_voice_index() const {  // sampler.rs:826-860
    for (size_t i = 0; i < voices.size(); ++i) if (!voices[i].is_active()) return i;
    size_t candidate = 0;
    bool has_earliest = false; uint64_t earliest = 0;
    bool has_oldest = false; uint64_t oldest = 0;
    for (size_t i = 0; i < voices.size(); ++i) {
      const auto& v = voices[i];
      if (envelope_parameters && v.in_release_stage()) {
        if (v.has_release_start) {
          if (!has_earliest || v.release_start_frame < earliest) {
            has_earliest = true; earliest = v.release_start_frame;
            has_oldest = false;
            candidate = i;
          }
        }
      } else if (!has_earliest) {
        if (v.has_note) {
          if (!has_oldest || v.note_id < oldest) { has_oldest = true; oldest = v.note_id; candidate = i; }
        }
      }
    }
    return candidate;
  }
  void all_notes_off(uint64_t frame) { for (auto& v : voices) v.stop(envelope_parameters, frame); }
  SamplerVoice* find_voice(uint64_t id) {
    for (auto& v : voices) if (v.has_note && v.note_id == id) return &v;
    return nullptr;
  }
  // FloatParameter::denormalize_value with its ParameterScaling (parameter/float.rs, scaling.rs:45-75)
  static float denorm_linear(float n, float lo, float hi) { return lo + n * (hi - lo); }
  static float denorm_exp2(float n, float lo, float hi) { return lo + std::pow(n, 2.0f) * (hi - lo); }
  static float denorm_db(float n, float lo, float hi, float db_lo, float db_hi) {
    float db = db_lo + n * (db_hi - db_lo);
    float l = db_to_linear(db_lo), h = db_to_linear(db_hi);
    return lo + ((db_to_linear(db) - l) / (h - l)) * (hi - lo);
  }
  static float clampf(float v, float lo, float hi) { return std::min(std::max(v, lo), hi); }
  // Sampler::process_parameter_update (sampler.rs:1069-1192): base + envelope parameters. Returns false for ids the
  // reference answers with Error::ParameterError (logged and ignored on the audio thread, sampler.rs:694-697).
  bool process_parameter_update(uint32_t id, float value, bool normalized) {
    const float n = clampf(value, 0.0f, 1.0f);
    auto time_value = [&]() { return normalized ? denorm_exp2(n, 0.0f, 10.0f) : clampf(value, 0.0f, 10.0f); };
    if (id == fourcc("STRN")) {
      base_transpose = normalized ? (int32_t)std::round(-48.0f + n * 96.0f) : std::min(std::max((int32_t)value, -48), 48);
      for (auto& v : voices) if (v.is_active()) v.set_base_pitch(base_transpose, base_finetune);
      return true;
    }
    if (id == fourcc("SFTN")) {
      base_finetune = normalized ? (int32_t)std::round(-100.0f + n * 200.0f) : std::min(std::max((int32_t)value, -100), 100);
      for (auto& v : voices) if (v.is_active()) v.set_base_pitch(base_transpose, base_finetune);
      return true;
    }
    if (id == fourcc("SVOL")) {
      base_volume = normalized ? denorm_db(n, 0.000001f, 15.848932f, -60.0f, 24.0f) : clampf(value, 0.000001f, 15.848932f);
      for (auto& v : voices) if (v.is_active()) v.set_base_volume(base_volume);
      return true;
    }
    if (id == fourcc("SPAN")) {
      base_panning = normalized ? denorm_linear(n, -1.0f, 1.0f) : clampf(value, -1.0f, 1.0f);
      for (auto& v : voices) if (v.is_active()) v.set_base_panning(base_panning);
      return true;
    }
    if (envelope_parameters) {  // Sampler::set_envelope_parameter (sampler.rs:184-218)
      AhdsrParameters& p = *envelope_parameters;
      if (id == fourcc("AATK")) { p.set_attack_time(Duration::from_secs_f32(std::max(time_value(), 0.0f))); return true; }
      if (id == fourcc("AHLD")) { p.set_hold_time(Duration::from_secs_f32(std::max(time_value(), 0.0f))); return true; }
      if (id == fourcc("ADCY")) { p.set_decay_time(Duration::from_secs_f32(std::max(time_value(), 0.0f))); return true; }
      if (id == fourcc("ASTN")) { return p.set_sustain_level(normalized ? denorm_linear(n, 0.0f, 1.0f) : clampf(value, 0.0f, 1.0f)); }
      if (id == fourcc("AREL")) { p.set_release_time(Duration::from_secs_f32(std::max(time_value(), 0.0f))); return true; }
    }
    return false;
  }
  static bool is_parameter(uint32_t id, bool has_env) {
    if (id == fourcc("STRN") || id == fourcc("SFTN") || id == fourcc("SVOL") || id == fourcc("SPAN")) return true;
    return has_env && (id == fourcc("AATK") || id == fourcc("AHLD") || id == fourcc("ADCY") || id == fourcc("ASTN") || id == fourcc("AREL"));
  }
  void process_playback_messages(uint64_t frame) {  // sampler.rs:656-731
    GenMsg m;
    while (queue->pop(m)) {
      if (m.is_stop) {
        stopping = transient;
        all_notes_off(frame);
      } else if (!stopping) {
        const GenEvent& e = m.event;
        switch (e.kind) {
          case GenEvent::AllNotesOff: all_notes_off(frame); break;
          case GenEvent::NoteOn: {
            float vol = e.has_volume ? e.volume : 1.0f;
            float pan = e.has_panning ? e.panning : 0.0f;
            size_t idx = next_free_voice_index();
            voices[idx].start(e.note_id, e.note, vol, pan, base_transpose, base_finetune, base_volume, base_panning, envelope_parameters, granular_parameters);
            active_voices += 1;
            break;
          }
          case GenEvent::NoteOff: if (auto* v = find_voice(e.note_id)) v->stop(envelope_parameters, frame); break;
          case GenEvent::SetSpeed: if (auto* v = find_voice(e.note_id)) v->set_speed(e.speed, e.has_glide, e.glide, base_transpose, base_finetune); break;
          case GenEvent::SetVolume: if (auto* v = find_voice(e.note_id)) v->set_volume(e.volume, base_volume); break;
          case GenEvent::SetPanning: if (auto* v = find_voice(e.note_id)) v->set_panning(e.panning, base_panning); break;
          case GenEvent::SetParameter: (void)process_parameter_update(e.param_id, e.param_value, e.param_normalized); break;
          case GenEvent::SetLoopRange: {  // Sampler::process_message (sampler.rs:1246-1271)
            uint64_t fc = voices.empty() ? 0 : voices[0].file->file_buffer->frame_count();
            if (!e.has_range || (e.range_start < fc && e.range_end <= fc))
              for (auto& v : voices) v.set_loop_range(e.has_range, e.range_start, e.range_end);
            break;
          }
        }
      }
    }
  }
  size_t write(float* out, size_t len, const SourceTime& time) override {  // sampler.rs:974-1028
    process_playback_messages(time.pos_in_frames);
    if (stopped || (active_voices == 0 && !stopping)) return 0;
    clear_buffer(out, len);
    size_t active = 0;
    for (auto& v : voices) {
      if (v.is_active()) {
        float* mix = temp_buffer.data();
        clear_buffer(mix, len);
        size_t written = v.process(mix, len, output_channel_count, envelope_parameters, granular_parameters, time);
        add_buffers(out, mix, written);
        if (v.is_active()) active += 1;
      }
    }
    active_voices = active;
    if (stopping && active == 0) stopped = true;
    return len;
  }
};

// ---- trait Effect (src/effect.rs:86-215) ---------------------------------------------------------------
struct Effect {
  virtual ~Effect() {}
  virtual const char* name() const = 0;
  virtual size_t weight() const = 0;
  virtual bool initialize(uint32_t sample_rate, size_t channel_count, size_t max_frames) = 0;
  virtual void process_started() {}
  virtual void process_stopped() {}
  virtual void process(float* buf, size_t len, uint64_t time_frames) = 0;
  virtual bool process_tail(size_t& frames) const { (void)frames; return false; }  // Option<usize>
  virtual bool process_parameter_update(uint32_t id, const ParamUpdate& value) = 0;
  virtual bool process_message(uint32_t message) { (void)message; return false; }  // Effect::process_message (effect.rs)
};

// ---- src/source/mixed/effect.rs:10-153 ------------------------------------------------------------------
struct EffectProcessor {
  std::unique_ptr<Effect> effect;
  bool bypassed = true;
  size_t tail_counter = 0;
  size_t silence_counter = USIZE_MAX;
  static constexpr float SILENCE_THRESHOLD = 0.001f;
  static constexpr size_t SILENCE_SECONDS = 2;
  explicit EffectProcessor(std::unique_ptr<Effect> e) : effect(std::move(e)) {}
  size_t weight() const { return bypassed ? 1 : effect->weight(); }
  static size_t sat_add(size_t a, size_t b) { return (a > USIZE_MAX - b) ? USIZE_MAX : a + b; }
  void reset_tail_counters() { tail_counter = USIZE_MAX; silence_counter = 0; }
  bool process(float* out, size_t len, size_t channel_count, uint32_t sample_rate, bool input_bypassed, uint64_t time) {
    bool should_bypass = input_bypassed && tail_counter == 0 && silence_counter == USIZE_MAX;
    if (should_bypass && !bypassed) { effect->process_stopped(); bypassed = true; }
    else if (!should_bypass && bypassed) { effect->process_started(); bypassed = false; reset_tail_counters(); }
    if (bypassed) return false;
    effect->process(out, len, time);
    if (input_bypassed) {
      size_t tail_frames;
      if (effect->process_tail(tail_frames)) {
        if (tail_frames == USIZE_MAX) tail_counter = tail_frames;
        else if (tail_counter == USIZE_MAX) tail_counter = tail_frames;
        else { size_t fp = len / channel_count; tail_counter = tail_counter > fp ? tail_counter - fp : 0; }
        silence_counter = USIZE_MAX;
      } else {
        float m = max_abs_sample(out, len);
        if (m < SILENCE_THRESHOLD) {
          silence_counter = sat_add(silence_counter, len / channel_count);
          if (silence_counter >= SILENCE_SECONDS * (size_t)sample_rate) { tail_counter = 0; silence_counter = USIZE_MAX; }
        } else {
          silence_counter = 0;
        }
      }
    } else {
      reset_tail_counters();
    }
    return true;
  }
};

// ---- MixerMessage / MixerEvent (src/source/mixed.rs:47-194) --------------------------------------------
struct MixerEvent {
  enum Kind { SeekSource, SetSourceSpeed, SetSourceVolume, SetSourcePanning, TriggerGenerator, EffectParameter, EffectMessage } kind;
  uint32_t target = 0;  // playback id / effect id
  uint64_t sample_time = 0;
  Duration position; double speed = 1; bool has_glide = false; float glide = 0;
  float value = 0;
  GenEvent gen;
  uint32_t param_id = 0; ParamUpdate param{0, false};
};

// PlaybackMessageQueue (src/source/playback.rs)
struct PlaybackQueues {
  std::shared_ptr<ArrayQueue<FileMsg>> file;
  std::shared_ptr<ArrayQueue<GenMsg>> generator;
  std::shared_ptr<ArrayQueue<float>> volume, panning;
};

struct MixedSource;

// ---- src/source/mixed/submixer.rs:23-77 ------------------------------------------------------------------
struct SubMixerProcessor {
  std::unique_ptr<MixedSource> mixer;
  size_t silence_counter = 0;
  bool process(float* out, float* mix, size_t len, size_t channel_count, uint32_t sample_rate, const SourceTime& time);
};

// ---- src/source/mixed.rs:199-924 --------------------------------------------------------------------------
struct MixedSource : Source {
  struct PlayingSource {
    bool is_active = true, is_transient = true;
    uint32_t playback_id = 0;
    PlaybackQueues queues;
    std::unique_ptr<Source> source;
    uint64_t start_time = 0;
    bool has_stop_time = false; uint64_t stop_time = 0;
  };
  struct Message {  // the subset of MixerMessage the offline path uses
    enum Kind { AddSource, StopSource, AddMixer, AddEffect, Event, RemoveSource, RemoveMixer, RemoveEffect, MoveEffect,
                RemoveAllPendingEvents } kind;
    uint32_t movement = 0; int32_t offset = 0;  // MoveEffect: EffectMovement (0 Direction(offset), 1 Start, 2 End)
    std::shared_ptr<PlayingSource> source;  // AddSource (moved out on processing)
    uint32_t id = 0; uint64_t sample_time = 0;
    std::shared_ptr<SubMixerProcessor> mixer;
    std::shared_ptr<EffectProcessor> effect;
    MixerEvent event;
  };
  std::deque<std::unique_ptr<PlayingSource>> playing_sources;
  std::vector<std::pair<uint32_t, std::shared_ptr<SubMixerProcessor>>> mixers;
  std::vector<std::pair<uint32_t, std::shared_ptr<EffectProcessor>>> effects;
  bool effects_bypassed = true;
  std::deque<Message> message_queue;  // reference: ArrayQueue(4096); unbounded here (offline feed)
  std::deque<MixerEvent> events;
  size_t channels;
  uint32_t rate;
  std::vector<float> mix_buffer;
  static constexpr size_t MAX_MIX_BUFFER_SAMPLES = 8 * 1024;

  MixedSource(size_t ch, uint32_t sr) : channels(ch), rate(sr), mix_buffer(MAX_MIX_BUFFER_SAMPLES) {}
  uint32_t sample_rate() const override { return rate; }
  size_t channel_count() const override { return channels; }
  bool is_exhausted() const override { return false; }
  size_t weight() const override { return 1; }

  void insert_event(const MixerEvent& e) {  // utils/event.rs:31-38
    size_t pos = 0;
    while (pos < events.size() && events[pos].sample_time <= e.sample_time) ++pos;
    events.insert(events.begin() + pos, e);
  }
  PlayingSource* find_source(uint32_t id) {
    for (auto& s : playing_sources) if (s->playback_id == id) return s.get();
    return nullptr;
  }
  void process_messages(const SourceTime& time) {  // mixed.rs:294-499
    while (!message_queue.empty()) {
      Message m = std::move(message_queue.front());
      message_queue.pop_front();
      switch (m.kind) {
        case Message::AddSource: {
          size_t pos = 0;
          while (pos < playing_sources.size() && playing_sources[pos]->start_time < m.sample_time) ++pos;
          auto ps = std::make_unique<PlayingSource>(std::move(*m.source));
          playing_sources.insert(playing_sources.begin() + pos, std::move(ps));
          break;
        }
        case Message::StopSource:
          if (auto* s = find_source(m.id)) { s->has_stop_time = true; s->stop_time = m.sample_time; }
          break;
        case Message::AddMixer: mixers.emplace_back(m.id, m.mixer); break;
        case Message::AddEffect: effects.emplace_back(m.id, m.effect); effects_bypassed = false; break;
        case Message::Event: insert_event(m.event); break;
        case Message::RemoveAllPendingEvents:  // mixed.rs:297-305
          for (size_t i = 0; i < playing_sources.size();) {
            if (playing_sources[i]->is_transient && playing_sources[i]->start_time > time.pos_in_frames) playing_sources.erase(playing_sources.begin() + i);
            else ++i;
          }
          for (size_t i = 0; i < events.size();) {
            if (events[i].sample_time > time.pos_in_frames) events.erase(events.begin() + i);
            else ++i;
          }
          break;
        case Message::RemoveSource:  // mixed.rs:391-393
          for (size_t i = 0; i < playing_sources.size();) {
            if (playing_sources[i]->playback_id == m.id) playing_sources.erase(playing_sources.begin() + i);
            else ++i;
          }
          break;
        case Message::RemoveMixer:  // mixed.rs:417-419
          for (size_t i = 0; i < mixers.size();) { if (mixers[i].first == m.id) mixers.erase(mixers.begin() + i); else ++i; }
          break;
        case Message::RemoveEffect:  // mixed.rs:432-439
          for (size_t i = 0; i < effects.size(); ++i)
            if (effects[i].first == m.id) { effects.erase(effects.begin() + i); if (effects.empty()) effects_bypassed = true; break; }
          break;
        case Message::MoveEffect:  // mixed.rs:440-459
          for (size_t i = 0; i < effects.size(); ++i)
            if (effects[i].first == m.id) {
              auto fx = effects[i];
              effects.erase(effects.begin() + i);
              size_t pos;
              if (m.movement == 0) { int32_t t = (int32_t)i + m.offset; pos = (size_t)std::min(std::max(t, 0), (int32_t)effects.size()); }
              else if (m.movement == 1) pos = 0;
              else pos = effects.size();
              effects.insert(effects.begin() + pos, fx);
              break;
            }
          break;
      }
    }
  }
  void process_event(const MixerEvent& e) {  // mixed.rs:761-924
    switch (e.kind) {
      case MixerEvent::SeekSource:
        if (auto* s = find_source(e.target)) if (s->queues.file) { FileMsg m{FileMsg::Seek}; m.position = e.position; s->queues.file->push(m); }
        break;
      case MixerEvent::SetSourceSpeed:
        if (auto* s = find_source(e.target)) if (s->queues.file) { FileMsg m{FileMsg::SetSpeed}; m.speed = e.speed; m.has_glide = e.has_glide; m.glide = e.glide; s->queues.file->push(m); }
        break;
      case MixerEvent::SetSourceVolume:
        if (auto* s = find_source(e.target)) s->queues.volume->force_push(e.value);
        break;
      case MixerEvent::SetSourcePanning:
        if (auto* s = find_source(e.target)) s->queues.panning->force_push(e.value);
        break;
      case MixerEvent::TriggerGenerator:
        if (auto* s = find_source(e.target)) if (s->queues.generator) {
          GenMsg m; m.event = e.gen;
          if (!s->queues.generator->push(m)) s->queues.generator->force_push(m);
        }
        break;
      case MixerEvent::EffectParameter:
        for (auto& fx : effects) if (fx.first == e.target) { fx.second->effect->process_parameter_update(e.param_id, e.param); break; }
        break;
      case MixerEvent::EffectMessage:  // mixed.rs: ProcessEffectMessage -> Effect::process_message
        for (auto& fx : effects) if (fx.first == e.target) { fx.second->effect->process_message(e.param_id); break; }
        break;
    }
  }
  bool process_sub_mixers(float* out, size_t len, const SourceTime& time) {  // sequential arm, mixed.rs:539-553
    bool produced = false;
    for (auto& m : mixers) produced |= m.second->process(out, mix_buffer.data(), len, channels, rate, time);
    return produced;
  }
  bool process_sources(float* out, size_t len, const SourceTime& time) {  // mixed.rs:558-624
    bool produced = false;
    size_t out_frames = len / channels;
    for (auto& ps : playing_sources) {
      size_t total = 0;
      if (ps->start_time > time.pos_in_frames) {
        size_t until = (size_t)(ps->start_time - time.pos_in_frames);
        if (until > 0) {
          if (until >= out_frames) break;
          total += until * channels;
        }
      }
      while (total < len) {
        SourceTime st{time.pos_in_frames + total / channels};
        uint64_t until_stop = UINT64_MAX;
        if (ps->has_stop_time) {
          uint64_t d = ps->stop_time > st.pos_in_frames ? ps->stop_time - st.pos_in_frames : 0;
          until_stop = d * channels;
        }
        if (until_stop == 0) {
          if (ps->queues.file) { FileMsg m{FileMsg::Stop}; ps->queues.file->force_push(m); }        // send_stop
          else if (ps->queues.generator) { GenMsg m; m.is_stop = true; ps->queues.generator->force_push(m); }
          ps->has_stop_time = false;
          until_stop = UINT64_MAX;
        }
        size_t remaining = (size_t)std::min<uint64_t>(len - total, until_stop);
        size_t to_write = std::min(remaining, mix_buffer.size());
        size_t written = ps->source->write(mix_buffer.data(), to_write, st);
        add_buffers(out + total, mix_buffer.data(), written);
        total += written;
        produced |= written > 0;
        if (ps->is_transient && ps->source->is_exhausted()) { ps->is_active = false; break; }
        else if (written == 0) break;
      }
    }
    return produced;
  }
  std::vector<const float*> ext_bus;   // pb200_set_main_inputs (the oracle's "device" memory is host memory)
  uint64_t ext_start = 0, ext_frames = 0;
  void process_effects(float* out, size_t len, const SourceTime& time, bool input_bypassed) {  // mixed.rs:627-655
    if (effects_bypassed && input_bypassed) return;
    bool all_bypassed = true;
    for (auto& fx : effects) {
      bool active = fx.second->process(out, len, channels, rate, input_bypassed, time.pos_in_frames);
      if (active) { input_bypassed = false; all_bypassed = false; }
    }
    effects_bypassed = all_bypassed;
  }
  size_t write(float* out, size_t len, const SourceTime& time) override {  // mixed.rs:659-719
    process_messages(time);
    if (playing_sources.empty() && effects.empty() && mixers.empty() && events.empty() && ext_bus.empty()) return 0;
    clear_buffer(out, len);
    size_t out_frames = len / channels;
    size_t done = 0;
    while (done < out_frames) {
      uint64_t now = time.pos_in_frames + done;
      while (!events.empty() && events.front().sample_time <= now) {
        MixerEvent e = events.front();
        events.pop_front();
        process_event(e);
      }
      size_t remaining = out_frames - done;
      size_t in_temp = mix_buffer.size() / channels;
      size_t until_event = events.empty() ? USIZE_MAX : (size_t)(events.front().sample_time - now);
      size_t n = std::min(std::min(remaining, in_temp), until_event);
      if (n > 0) {
        SourceTime ct{time.pos_in_frames + done};
        float* chunk = out + done * channels;
        bool audible = false;
        if (!ext_bus.empty() && ct.pos_in_frames >= ext_start && ct.pos_in_frames < ext_start + ext_frames) {
          // pb200_set_main_inputs: the outputs of sub-mixers rendered elsewhere, added like SubMixerProcessors', in order
          const size_t m = (size_t)std::min<uint64_t>(n, ext_start + ext_frames - ct.pos_in_frames);
          for (const float* bus : ext_bus) {
            const float* e = bus + (ct.pos_in_frames - ext_start) * channels;
            for (size_t i = 0; i < m * channels; ++i) chunk[i] += e[i];
          }
          audible = true;
        }
        audible |= process_sub_mixers(chunk, n * channels, ct);
        audible |= process_sources(chunk, n * channels, ct);
        process_effects(chunk, n * channels, ct, !audible);
        done += n;
      }
    }
    for (size_t i = 0; i < playing_sources.size();) {
      if (playing_sources[i]->is_transient && !playing_sources[i]->is_active) playing_sources.erase(playing_sources.begin() + i);
      else ++i;
    }
    return len;
  }
};

inline bool SubMixerProcessor::process(float* out, float* mix, size_t len, size_t channel_count, uint32_t sample_rate, const SourceTime& time) {
  size_t written = mixer->write(mix, len, time);
  float m = max_abs_sample(mix, written);
  if (m < EffectProcessor::SILENCE_THRESHOLD) {
    silence_counter += len / channel_count;
    if (silence_counter < EffectProcessor::SILENCE_SECONDS * (size_t)sample_rate) { add_buffers(out, mix, written); return true; }
    return false;
  }
  silence_counter = 0;
  add_buffers(out, mix, written);
  return true;
}

}  // namespace po
