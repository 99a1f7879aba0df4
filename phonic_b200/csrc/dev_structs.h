// Device-resident data model of the renderer (shared by host graph compiler and kernels).
//
// Layout in HBM:
//   sample buffers   : one contiguous f32 array per uploaded AudioFileBuffer (interleaved)
//   voices           : VoiceState[n_voices]   (one record per sampler voice / file playback)
//   groups           : GroupParams[n_groups] (immutable) + GroupState[n_groups] (mutable)
//                      a group = one phonic `Source` attached to a mixer: a Sampler (N voices)
//                      or one file playback (1 voice)
//   events           : DevEvent[] sorted per group by (sample_time, insertion order)
//   chunk schedule   : per mixer, the reference's exact `MixedSource::write` chunk boundaries
//   group bus ring   : f32 [ring][group][block_frames][2]   (per time block scratch)
//   mixer buses      : f32 [ring][mixer][block_frames][2]
#pragma once
#include <stdint.h>

namespace pb {

constexpr uint32_t REPEAT_FOREVER = 0xFFFFFFFFu;
constexpr int MAX_EFFECTS_PER_MIXER = 16;

struct DevBuffer {
  const float* data;    // interleaved samples (incl. the +1 zero pad frame)
  uint32_t n_samples;   // buffer().len()
  uint32_t channels;    // 1 or 2
  uint32_t sample_rate;
  int32_t loop_start;   // frames, -1 = no embedded loop
  int32_t loop_end;
  uint32_t _pad;
};

// ExponentialSmoothedValue (src/utils/smoothing.rs:131-228); inertia and rate comp are constants
struct ExpSm {
  float current, target;
};

enum FaderState : uint8_t { FADER_STOPPED = 0, FADER_RUNNING = 1, FADER_FINISHED = 2 };
enum EnvStage : uint8_t { ENV_IDLE = 0, ENV_ATTACK = 1, ENV_HOLD = 2, ENV_DECAY = 3, ENV_SUSTAIN = 4, ENV_RELEASE = 5 };

// One voice = Panned<Amplified<ChannelMapped<PreloadedFileSource>>> (+ AhdsrEnvelope for samplers).
struct VoiceState {
  // PreloadedFileSource (src/source/file/preloaded.rs:29-37) + FileSourceImpl (file/common.rs:31-52)
  double current_speed, target_speed;
  uint64_t note_id;            // SamplerVoice::note_id (valid when has_note)
  uint64_t release_start;      // SamplerVoice::release_start_frame
  uint64_t end_frame;          // output frame at which playback finished (status only)
  uint64_t pos_clock;          // FileSourceImpl::playback_pos_sample_time_clock: output frame of the last Position event
  uint32_t playback_pos;       // sample index into the buffer
  uint32_t repeat, repeat_count;
  int32_t loop_ovr_start, loop_ovr_end;  // loop_range_override in frames, -1 = None
  uint32_t to_next_speed_update;         // frames until the next glide update
  float glide_rate;
  // CubicInterpolator x channels (src/utils/resampler/cubic.rs:10-15); sub_pos/ratio shared
  float ratio, sub_pos;
  int32_t hidx[4];             // sample index (channel 0) of input[0..3]; -1 = the zero a reset leaves
  // VolumeFader (src/utils/fader.rs:27-34)
  float fader_cur, fader_tgt, fader_inertia;
  // AmplifiedSource / PannedSource smoothers
  ExpSm vol, pan;
  // AhdsrEnvelope (src/utils/ahdsr.rs:367-373)
  float env_target, env_hold, env_release_out, env_out;
  float note_volume, note_panning;
  uint8_t env_stage;
  uint8_t fader_state;
  uint8_t initialized;         // CubicInterpolator::is_initialized
  uint8_t pos_eof, finished;   // playback_pos_eof, playback_finished
  uint8_t has_note, has_release, note;
  uint8_t stopped_exhausted;
  uint8_t hq;                  // ResamplingQuality::HighQuality: 0 cubic, 1 rubato sinc, 2 rubato with equal rates (bypass)
  uint8_t _pad[6];
};
static_assert(sizeof(VoiceState) % 8 == 0, "VoiceState must stay 8-byte aligned");

// RubatoResampler + rubato::SincFixedIn state of one HighQuality file voice (src/utils/resampler/rubato.rs:12-56).
// The resampler's input is a stream of 256-frame chunks cut from the sample buffer (zero padded at loop ends /
// EOF, preloaded.rs:296-304); SincFixedIn keeps the last two chunks as history, so the open chunk's output is a
// function of three (position, valid frames) pairs -- no sample has to be copied.
constexpr uint32_t HQ_CHUNK = 256;     // RubatoResampler CHUNK_SIZE (rubato.rs:27) == sinc_len
constexpr uint32_t HQ_FACTOR = 128;    // oversampling_factor (rubato.rs:31)
constexpr uint32_t HQ_NONE = 0xFFFFFFFFu;
struct HqState {
  double idx0;          // SincFixedIn::last_index before the open chunk was processed
  double last_index;    // ... after it
  double t_ratio;       // 1 / resample_ratio
  uint32_t src[3];      // first sample index of chunks k-2, k-1, k (k = open chunk)
  uint16_t valid[3];    // valid frames of each; the rest of the 256 is zero padding
  uint16_t _pad0;
  uint32_t n_out;       // output frames of the open chunk
  uint32_t pending;     // of those, not yet consumed (RubatoResampler::pending)
  uint32_t slot;        // row of this voice in the per-block stream scratch
  uint32_t table;       // sinc table index (one per cutoff)
  uint32_t rec;         // this block's HqRec of the open chunk, HQ_NONE = not emitted yet
  int32_t end_idx;      // chunk_size - (sinc_len + 1) - ceil(t_ratio)
};

// One piece of resampler output the sinc kernel has to materialise into the stream scratch of a time block.
struct HqRec {
  double idx0, t_ratio;
  uint32_t src[3];
  uint16_t valid[3];
  uint16_t kind;        // 0 sinc, 1 copy (equal-rate bypass, zero padded)
  uint32_t slot;
  uint32_t out_off;     // block-relative frame the first materialised output lands on
  uint32_t skip;        // outputs of the chunk consumed before `out_off` (earlier block)
  uint32_t count;       // outputs to materialise
  uint32_t buffer, table;
};

// ---- granular playback (src/generator/sampler/granular.rs) ---------------------------------------------------
constexpr uint32_t GRAIN_POOL = 100;   // GRAIN_POOL_SIZE (src/generator/sampler/voice.rs:33)
constexpr uint32_t GRAIN_LUT_N = 2048; // GRAIN_WINDOW_LUT (granular.rs:221)

// GranularParameters of a Sampler, resolved on the host (static: no parameter automation yet)
struct GranGroup {
  uint32_t enabled;
  uint32_t overlap_mode;      // 0 Cloud 1 Sequential
  uint32_t window;            // GrainWindowMode
  uint32_t backward;          // GrainPlaybackDirection::Backward
  float position, step;
  float trigger_inc;          // clamp(density, 1, 100) / sample_rate (granular.rs:798-801)
  float crossfade;            // GrainWindowMode::sequential_crossfade_point
  uint32_t grain_size;        // max((size_ms * 1.0 * sr / 1000) as usize, 2) (granular.rs:843-845)
  uint32_t buffer;            // DevBuffer of the mono sample data at the output rate (sampler.rs:908-952)
  uint32_t buf_len;           // its length in frames
  uint32_t has_loop;          // sample_loop_range (voice.rs:355-361), normalised
  float loop_start, loop_end;
  uint32_t first_row;         // stream-scratch row of the sampler's first voice
  uint32_t _pad;
};

// GrainPool state of one voice that the control pass needs (the grains' own f64 recurrences live in the grain kernel)
struct GranState {
  double speed;                       // GrainPool::speed
  double primary_phase, primary_inc;  // Sequential: window phase of the primary grain (exact f64 accumulate)
  uint64_t primary_end;
  uint64_t max_end;                   // latest end frame of any grain: active_grain_indices.is_empty() <=> max_end <= now
  uint64_t slot_end[GRAIN_POOL];      // absolute frame after a slot's grain played its last sample; 0 = free
  uint32_t slot_rec[GRAIN_POOL];      // this block's GrainRec of the slot's grain
  uint32_t slot_total[GRAIN_POOL];    // grain_size_samples of the slot's grain
  float trigger_phase, playhead, volume, panning;
  uint32_t gen;                       // time-block generation the slot_rec entries belong to
  uint32_t n_order;
  uint8_t order[GRAIN_POOL];          // active_grain_indices: slots in activation order (stale entries allowed)
  uint8_t playing_loop, trigger_new, overlap_mode, has_primary;
  uint32_t primary_slot;
  uint32_t first_active;              // index into this block's per-voice record list of the oldest live record
  uint32_t n_recs;                    // records of this voice in this block
};

// One grain's contribution inside one time block (grain kernel work item)
struct GrainRec {
  double position, increment, win_inc, loop_start, loop_end;  // Grain::activate (granular.rs:1025-1067)
  float volume, panning;
  uint32_t row;        // voice row
  uint32_t start_off;  // block-relative frame of its first sample in this block
  uint32_t len;        // samples it plays in this block (shortened when the voice resets)
  uint32_t done;       // samples played before this block (> 0: state comes from the carry buffer)
  uint32_t total;      // grain_size_samples
  uint32_t slot;
  uint32_t storage;    // first frame of its contribution in the block's grain storage
  uint32_t buffer, buf_len;
  uint16_t window_mode;
  uint8_t has_loop, _pad;
};

// A grain that plays on into the next time block: its complete state after the block's last sample
struct GrainCarry {
  double position, window_phase, increment, win_inc, loop_start, loop_end;
  float volume, panning;
  uint32_t window_mode, has_loop;
};

enum GroupKind : uint32_t { GROUP_SAMPLER = 0, GROUP_FILE = 1 };

// PlaybackStatusEvent of a file playback as the skeleton pass records it (src/source/status.rs:15-36)
struct StatusRec { uint64_t frame; uint64_t pos; uint32_t group; uint32_t kind; };  // kind: 0 Position, 1 Stopped (exhausted), 2 Stopped

struct GroupParams {
  uint32_t kind;
  uint32_t first_voice, n_voices;
  uint32_t buffer;
  uint32_t mixer;        // dense mixer index
  uint32_t ev_begin, ev_end;
  uint32_t transient;
  uint64_t start_time;   // PlayingSource::start_time
  // AhdsrParameters (src/utils/ahdsr.rs:26-39), sample rate applied on the host
  uint32_t has_env;
  float attack_rate, decay_rate, release_rate, sustain_level, hold_samples;
  float attack_scaling, decay_scaling, release_scaling;
  uint32_t hold_is_zero, decay_is_zero, release_is_zero;
  // fade-out of voices (50 ms default): inertia precomputed on the host with the reference's f32 math
  uint32_t has_fade_out;
  float fade_out_inertia;
  float base_volume, base_panning;  // Sampler::base_volume/base_panning
  uint32_t _pad;
};

struct GroupState {
  uint64_t stop_time;
  uint32_t has_stop_time;
  uint32_t ev_cursor;
  uint32_t active_voices;
  uint32_t stopping, stopped;
  uint32_t dead;          // removed from the mixer (transient && exhausted)
  ExpSm vol, pan;         // generator-level AmplifiedSource / PannedSource (player.rs:1075-1081)
  uint64_t voice_frames;  // statistics: active voice-frames rendered
  uint64_t dead_time;     // end of the mixer chunk in which the source was dropped (valid when dead)
  uint32_t gp_idx;        // index of the group's CURRENT parameter version in the GroupParams array (sampler parameter automation)
  uint32_t _pad;
};

enum EventKind : uint32_t {
  EVK_STOP = 1, EVK_SET_VOLUME = 2, EVK_SET_PANNING = 3, EVK_SET_SPEED = 4, EVK_SEEK = 5,
  EVK_NOTE_ON = 10, EVK_NOTE_OFF = 11, EVK_ALL_NOTES_OFF = 12, EVK_NOTE_SPEED = 13, EVK_NOTE_VOLUME = 14,
  EVK_NOTE_PANNING = 15,
  // Sampler::process_parameter_update (sampler.rs:1069-1192): seek_pos = index of the parameter version that becomes current,
  // note = what the active voices have to re-derive (PARAM_*), speed = the new pitch factor
  EVK_SET_PARAM = 16,
  // SamplerMessage::SetLoopRange (sampler.rs:1246-1271): seek_pos / note = first / end frame, flags bit1: None
  EVK_SET_LOOP = 17
};
enum ParamClass : uint32_t { PARAM_PITCH = 0, PARAM_VOLUME = 1, PARAM_PANNING = 2, PARAM_ENVELOPE = 3 };

struct DevEvent {
  uint64_t time;
  uint64_t note_id;
  double speed;        // effective speed (note speed * pitch factor resolved on the host)
  uint32_t kind;
  float value, value2; // volume / panning
  float glide;         // <= 0: None
  uint32_t note;
  uint32_t seek_pos;   // SEEK: clamped sample index (preloaded.rs:140-143)
  uint32_t flags;      // bit0: immediate message (does not split mixer chunks), bit1: SET_LOOP without a range
  uint32_t flags2;     // host scratch
};

// ---- mixers / effects -------------------------------------------------------------------------------
enum FxKind : uint32_t { FX_FILTER = 1, FX_EQ5 = 2, FX_COMPRESSOR = 3, FX_CHORUS = 4, FX_DELAY = 5, FX_REVERB = 6, FX_GAIN = 7, FX_PANNING = 8, FX_GATE = 9, FX_DISTORTION = 10 };

struct LinSm { float current, target, step, current_step; uint32_t pending; };
struct SpringSm { float current, velocity, target, omega; };

struct BiquadCoef { double a1, a2, a3, m0, m1, m2; float cutoff, q, gain; uint32_t type, sample_rate; uint32_t _pad; };
struct SvfCoef { double g, k, a1, a2, a3; float cutoff, resonance; uint32_t type, sample_rate; };

struct FxParamEvent { uint64_t time; uint32_t param_id; float value; uint32_t normalized; uint32_t _pad; };

// EffectProcessor (src/source/mixed/effect.rs:10-15) + per-kind state blob (see effects.cuh)
struct FxHeader {
  uint32_t kind;
  uint32_t bypassed;
  uint64_t tail_counter, silence_counter;  // usize::MAX sentinels as in the reference
  uint32_t ev_begin, ev_end, ev_cursor;
  uint32_t state_offset;   // byte offset of the kind-specific state in the fx state arena
  uint32_t aux_offset;     // byte offset of delay-line storage in the fx aux arena (f64 units)
  uint32_t _pad;
};

struct MixerParams {
  uint32_t parent;         // dense index, 0xFFFFFFFF for main
  uint32_t depth;
  uint32_t child_begin, child_end;   // into child index array (reference mixer order)
  uint32_t src_begin, src_end;       // into source (group) index array (playing_sources order)
  uint32_t fx_begin, fx_end;         // into FxHeader array (effect chain order)
  uint32_t _pad[2];
};

struct MixerState {
  uint64_t silence_counter;   // SubMixerProcessor::silence_counter
  uint32_t effects_bypassed;  // MixedSource::effects_bypassed
  uint32_t _pad;
};

}  // namespace pb
