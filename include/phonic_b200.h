/*
 * phonic_b200.h -- C-ABI of the B200-native offline renderer for phonic's mixer graph.
 *
 * This is the drop-in boundary: plain C types only, no torch/CUDA types.
 * Every entry point cites the reference (emuell/phonic v0.16.0) interface it stands in for;
 * paths are relative to the reference's repository root.
 *
 * Threading: one render call at a time per renderer (`&mut self` semantics of
 * `Source::write`, src/source.rs:80-110). Graph-building and scheduling calls must not
 * overlap a render call (offline use: no real-time constraint).
 *
 * Time: every `sample_time` is an absolute output frame position, exactly the `u64`
 * sample times phonic's handles take (src/player/handles/ *.rs). `PB200_TIME_NOW`
 * stands for `None` ("apply immediately" = at the next rendered block start).
 *
 * Errors: functions return 0 on success or a PB200_ERR_* code which maps 1:1 onto
 * `phonic::Error` (src/error.rs:8-22); `pb200_last_error` gives the message.
 * Nothing unwinds across this boundary.
 */
#ifndef PHONIC_B200_H
#define PHONIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define PB200_API
#else
#define PB200_API __attribute__((visibility("default")))
#endif

typedef struct pb200_renderer pb200_renderer;

/* ---- error codes: phonic::Error (src/error.rs:8-22) + device errors ---------------------- */
enum pb200_error {
  PB200_OK = 0,
  PB200_ERR_SOURCE_NOT_PLAYING = 1,   /* Error::SourceNotPlaying        */
  PB200_ERR_MEDIA_FILE_NOT_FOUND = 2, /* Error::MediaFileNotFound       */
  PB200_ERR_MEDIA_FILE_PROBE = 3,     /* Error::MediaFileProbeError     */
  PB200_ERR_MEDIA_FILE_SEEK = 4,      /* Error::MediaFileSeekError      */
  PB200_ERR_AUDIO_DECODING = 5,       /* Error::AudioDecodingError      */
  PB200_ERR_OUTPUT_DEVICE = 6,        /* Error::OutputDeviceError       */
  PB200_ERR_RESAMPLING = 7,           /* Error::ResamplingError         */
  PB200_ERR_GENERATOR_NOT_FOUND = 8,  /* Error::GeneratorNotFoundError  */
  PB200_ERR_EFFECT_NOT_FOUND = 9,     /* Error::EffectNotFoundError     */
  PB200_ERR_MIXER_NOT_FOUND = 10,     /* Error::MixerNotFoundError      */
  PB200_ERR_PARAMETER = 11,           /* Error::ParameterError          */
  PB200_ERR_SEND = 12,                /* Error::SendError               */
  PB200_ERR_IO = 13,                  /* Error::IoError                 */
  PB200_ERR_CUDA = 100,               /* CUDA runtime / no device (maps to OutputDeviceError) */
  PB200_ERR_UNSUPPORTED = 101         /* feature of the reference not (yet) rendered on device */
};

#define PB200_TIME_NOW UINT64_MAX /* `None` sample time of the handle methods */
#define PB200_MAIN_MIXER 0u       /* Player::MAIN_MIXER_ID (src/player.rs) */
#define PB200_REPEAT_DEFAULT (UINT64_MAX - 1) /* FilePlaybackOptions::repeat = None */
#define PB200_REPEAT_FOREVER UINT64_MAX     /* FilePlaybackOptions::repeat_forever() */
#define PB200_NO_LOOP (-1)
/* All durations cross the boundary as std::time::Duration::as_nanos() so that the
 * reference's own Duration -> f32/f64 conversions (as_secs_f32) are reproduced bit-exactly. */
#define PB200_DURATION_NONE UINT64_MAX

/* ---- renderer: Player + WavOutput (src/player.rs:289-403, src/output/wav.rs:41-120) ------- */
typedef struct pb200_config {
  uint32_t sample_rate;      /* WavOutput::open_with_specs sample rate (wav.rs:50-56) */
  uint32_t channel_count;    /* must be 2 (PlayerConfig::enforce_stereo_playback, player.rs:134) */
  uint32_t block_frames;     /* WavStream block: 1024 frames (wav.rs:25); 0 = default */
  int32_t device_ordinal;    /* CUDA device this renderer instance owns; -1 = current */
  float master_volume;       /* OutputDevice::set_volume (output.rs:41-47); 1.0 default */
  uint32_t reserved[3];
} pb200_config;

/* Player::new(WavOutput::open_with_specs(..)) */
PB200_API int pb200_create(const pb200_config *config, pb200_renderer **out);
/* drop(Player) */
PB200_API void pb200_destroy(pb200_renderer *r);
/* Display for Error */
PB200_API const char *pb200_last_error(const pb200_renderer *r);
/* static description of the build ("cuda sm_100a ..." / oracle) */
PB200_API const char *pb200_backend(void);

/* ---- sample data: AudioFileBuffer::new (src/source/file/buffer.rs:23-59) -------------------
 * `interleaved` is copied (to HBM). `add_pad_frame` != 0 appends the one zero frame that
 * AudioFileBuffer::from_audio_decoder adds for the cubic resampler (buffer.rs:103-104).
 * loop_start/loop_end in frames or PB200_NO_LOOP (RIFF smpl loops, decoder.rs:293-330). */
PB200_API int pb200_upload_buffer(pb200_renderer *r, const float *interleaved, uint64_t frames,
                                  uint32_t channel_count, uint32_t sample_rate,
                                  int64_t loop_start, int64_t loop_end, int add_pad_frame,
                                  uint32_t *buffer_id);

/* ---- mixers: Player::add_mixer (src/player.rs:772-836) ------------------------------------ */
PB200_API int pb200_add_mixer(pb200_renderer *r, uint32_t parent_mixer_id, uint32_t *mixer_id);

/* ---- effects: Player::add_effect (src/player.rs:893-939), Effect (src/effect.rs:86-215) --- */
enum pb200_effect_kind {
  PB200_FX_FILTER = 1,     /* FilterEffect     src/effect/filter.rs     */
  PB200_FX_EQ5 = 2,        /* Eq5Effect        src/effect/eq5.rs        */
  PB200_FX_COMPRESSOR = 3, /* CompressorEffect src/effect/compressor.rs */
  PB200_FX_CHORUS = 4,     /* ChorusEffect     src/effect/chorus.rs     */
  PB200_FX_DELAY = 5,      /* DelayEffect      src/effect/delay.rs      */
  PB200_FX_REVERB = 6,     /* ReverbEffect     src/effect/reverb.rs     */
  PB200_FX_GAIN = 7,       /* GainEffect       src/effect/gain.rs       */
  PB200_FX_PANNING = 8,    /* PanningEffect    src/effect/pan.rs        */
  PB200_FX_GATE = 9,       /* GateEffect       src/effect/gate.rs       */
  PB200_FX_DISTORTION = 10 /* DistortionEffect src/effect/distortion.rs */
};

/* FilterEffect::with_parameters(filter_type, cutoff, q) (filter.rs:104-116) */
typedef struct pb200_filter_params {
  uint32_t filter_type; /* FilterEffectType: 0 Lowpass 1 Bandpass 2 Bandstop 3 Highpass */
  float cutoff;
  float q;
} pb200_filter_params;

/* Eq5Effect has only Eq5Effect::new() (eq5.rs:153-170): pass params = NULL and set bands with
 * PB200_EV_SET_EFFECT_PARAMETER events ('gan1'..'gan5', 'frq1'..'frq5', 'bw_1'..'bw_5'). */

/* CompressorEffect::with_compressor_parameters / new_limiter (compressor.rs:114-157) */
typedef struct pb200_compressor_params {
  float threshold, ratio, knee, attack_time, release_time, makeup_gain, lookahead_time;
} pb200_compressor_params;

/* ChorusEffect::with_parameters (chorus.rs:178-200) */
typedef struct pb200_chorus_params {
  float rate, phase, depth, feedback, delay, wet;
  uint32_t filter_type; /* ChorusEffectFilterType: 0 Lowpass 1 Highpass 2 Bandpass */
  float filter_freq, filter_resonance;
} pb200_chorus_params;

/* DelayEffect has only DelayEffect::new() (delay.rs:180-212): pass params = NULL and use
 * PB200_EV_SET_EFFECT_PARAMETER events ('mode','dlay','fdbk','ftyp','cuto','driv','wet_','wdth',
 * 'lfor','lfos','lfdt','ldfb','lfdf'). The OS-seeded Random/SmoothRandom LFO shapes are rejected
 * with PB200_ERR_UNSUPPORTED (not reproducible in the reference itself). */

/* ReverbEffect::with_parameters(room_size, wet) (reverb.rs:153-159). The reference seeds
 * `fpd_l/fpd_r` and the 16 vibrato phases from rand::rng() (reverb.rs:95-103,137-144);
 * parity needs them injected, so they are explicit inputs here. */
typedef struct pb200_reverb_params {
  float room_size, wet;
  uint32_t fpd[2];
  double vib_phase[16]; /* [line 0..7][L,R] start phases in [0, 2pi) */
} pb200_reverb_params;

/* GainEffect::with_parameters(gain_db, dc_mode) (gain.rs:97-104). Parameters: 'gain' (linear gain; normalized
 * updates use ParameterScaling::Decibel(-60, 24)), 'dcfm'. */
typedef struct pb200_gain_params {
  float gain_db;           /* clamped to [-60, 24] */
  uint32_t dc_filter_mode; /* GainEffectDcFilterMode: 0 Off 1 Slow (1 Hz) 2 Default (5 Hz) 3 Fast (20 Hz) */
} pb200_gain_params;

/* GateEffect::with_parameters(threshold, attack, hold, release, range) (gate.rs:67-81): dB, seconds, seconds,
 * seconds, dB; out-of-range values are PB200_ERR_PARAMETER (the reference asserts). Parameters 'thrs', 'attk',
 * 'hold', 'rels', 'rnge'. */
typedef struct pb200_gate_params {
  float threshold, attack_time, hold_time, release_time, range;
} pb200_gate_params;

/* PanningEffect has only PanningEffect::new() (pan.rs:52-60): pass params = NULL and use
 * PB200_EV_SET_EFFECT_PARAMETER events ('pan ', 'wdth', 'invl', 'invr'; booleans: value != 0). */

/* DistortionEffect::with_parameters(distortion_type, drive, mix) (distortion.rs:247-253). Parameters: 'type', 'driv'
 * (0..4, linear ramp), 'mix ' (0..1). */
typedef struct pb200_distortion_params {
  uint32_t distortion_type; /* DistortionType: 0 SoftClip 1 HardClip 2 Diode (default) 3 Fuzz 4 Fold */
  float drive;
  float mix;
} pb200_distortion_params;

/* `params` may be NULL => Effect::new()/default(). */
PB200_API int pb200_add_effect(pb200_renderer *r, uint32_t mixer_id, uint32_t kind,
                               const void *params, size_t params_size, uint32_t *effect_id);

/* ---- file playback: Player::play_file_source (src/player.rs:511-600) ---------------------- */
typedef struct pb200_file_options { /* FilePlaybackOptions (src/source/file.rs:34-84) */
  float volume;               /* 1.0 */
  float panning;              /* 0.0 */
  double speed;               /* 1.0 */
  uint64_t repeat;            /* PB200_REPEAT_DEFAULT | count | PB200_REPEAT_FOREVER */
  int64_t loop_start;         /* loop_range override or PB200_NO_LOOP */
  int64_t loop_end;
  uint64_t fade_in_nanos;     /* Duration::as_nanos(); PB200_DURATION_NONE => None */
  uint64_t fade_out_nanos;    /* default 50 ms (file.rs:106) */
  uint32_t resampling_quality;/* 0 Default (cubic) 1 HighQuality (sinc) */
  uint32_t target_mixer;      /* PB200_MAIN_MIXER */
} pb200_file_options;

PB200_API void pb200_file_options_default(pb200_file_options *o);
PB200_API int pb200_play_file(pb200_renderer *r, uint32_t buffer_id, const pb200_file_options *o,
                              uint64_t start_time, uint32_t *playback_id);

/* ---- sampler: Sampler::from_file_source + with_ahdsr (src/generator/sampler.rs:487-596),
 *      Player::play_generator / add_generator (src/player.rs:700-770) ------------------------ */
typedef struct pb200_ahdsr { /* AhdsrParameters::new_with_scaling (src/utils/ahdsr.rs:75-98) */
  uint64_t attack_nanos; /* std::time::Duration::as_nanos() of each stage time */
  uint64_t hold_nanos;
  uint64_t decay_nanos;
  uint64_t release_nanos;
  float attack_scaling, decay_scaling, release_scaling;
  float sustain_level;
} pb200_ahdsr;

/* GranularParameters (src/generator/sampler/granular.rs:239-331); Sampler::with_granular_playback
 * (src/generator/sampler.rs:599-637). Every GrainPool seeds its SmallRng from the OS (granular.rs:413), so
 * settings that let a random draw reach the audio -- variation, spray, pan_spread > 0, Random direction --
 * are not reproducible in the reference itself and are rejected with PB200_ERR_UNSUPPORTED. */
typedef struct pb200_granular_params {
  uint32_t overlap_mode;       /* GrainOverlapMode: 0 Cloud 1 Sequential */
  uint32_t window;             /* GrainWindowMode: 0 Hann 1 Blackman 2 Triangle 3 Tukey 4 Trapezoid 5 Exponential 6 RampUp 7 RampDown */
  float size;                  /* ms, 1..1000 (default 100) */
  float density;               /* Hz, 1..100 (default 10) */
  float variation;             /* must be 0 */
  float spray;                 /* must be 0 */
  float pan_spread;            /* must be 0 */
  uint32_t playback_direction; /* GrainPlaybackDirection: 0 Forward 1 Backward (2 Random: unsupported) */
  float position;              /* 0..1 (default 0.5) */
  float step;                  /* -4..4 (default 0) */
} pb200_granular_params;

typedef struct pb200_sampler_options { /* GeneratorPlaybackOptions (src/generator.rs:41-72) */
  float volume;          /* 1.0 */
  float panning;         /* 0.0 */
  uint32_t voices;       /* 8 */
  uint32_t target_mixer; /* PB200_MAIN_MIXER */
  uint32_t transient;    /* 1: play_generator, 0: add_generator */
  uint32_t has_ahdsr;    /* with_ahdsr(..) */
  pb200_ahdsr ahdsr;
  uint32_t has_granular; /* with_granular_playback(..) */
  pb200_granular_params granular;
  uint32_t reserved;
} pb200_sampler_options;

PB200_API void pb200_sampler_options_default(pb200_sampler_options *o);
PB200_API int pb200_add_sampler(pb200_renderer *r, uint32_t buffer_id,
                                const pb200_sampler_options *o, uint64_t start_time,
                                uint32_t *generator_id);

/* ---- events: handle methods (src/player/handles/{file,generator,effect}.rs) ---------------- */
enum pb200_event_kind {
  /* FilePlaybackHandle / GeneratorPlaybackHandle: target = playback id */
  PB200_EV_STOP_SOURCE = 1,      /* stop(stop_time)                 handles/file.rs:69-98     */
  PB200_EV_SET_SOURCE_VOLUME = 2,/* set_volume(volume, t)           handles/file.rs:188-222   */
  PB200_EV_SET_SOURCE_PANNING = 3,/* set_panning(panning, t)        handles/file.rs:223-257   */
  PB200_EV_SET_SOURCE_SPEED = 4, /* set_speed(speed, glide, t)      handles/file.rs:135-187   */
  PB200_EV_SEEK_SOURCE = 5,      /* seek(position, t)               handles/file.rs:99-134    */
  /* GeneratorPlaybackHandle: target = generator id */
  PB200_EV_NOTE_ON = 10,         /* note_on(note, volume, panning, t) handles/generator.rs:200-238 */
  PB200_EV_NOTE_OFF = 11,        /* note_off(note_id, t)            */
  PB200_EV_ALL_NOTES_OFF = 12,   /* all_notes_off(t)                */
  PB200_EV_SET_NOTE_SPEED = 13,  /* set_note_speed(note_id, speed, glide, t) */
  PB200_EV_SET_NOTE_VOLUME = 14, /* set_note_volume(note_id, volume, t) */
  PB200_EV_SET_NOTE_PANNING = 15,/* set_note_panning(note_id, panning, t) */
  /* set_parameter((id, value), t) -> GeneratorPlaybackEvent::SetParameter (handles/generator.rs, src/generator.rs:172-226;
   * Sampler::process_parameter_update, src/generator/sampler.rs:1069-1192): param_id = 'STRN' transpose, 'SFTN' finetune
   * (integer parameters: raw values are truncated to i32), 'SVOL' volume, 'SPAN' panning, and with an AHDSR envelope
   * 'AATK' 'AHLD' 'ADCY' 'ASTN' 'AREL'; `value` raw or, with PB200_EVF_NORMALIZED, normalized. */
  PB200_EV_SET_GENERATOR_PARAMETER = 16,
  /* send_message(SamplerMessage::SetLoopRange(range), t) (src/generator/sampler.rs:1246-1271): position_nanos = first
   * frame, note_id = end frame of the loop; PB200_EVF_NO_RANGE: None (looping off). */
  PB200_EV_SET_GENERATOR_LOOP_RANGE = 17,
  /* EffectHandle: target = effect id */
  PB200_EV_SET_EFFECT_PARAMETER = 20, /* set_parameter((id, value), t) handles/effect.rs:67-98 */
  PB200_EV_EFFECT_MESSAGE = 21        /* send_message(message, t)      handles/effect.rs:127-163; param_id = PB200_MSG_* */
};
#define PB200_MSG_REVERB_RESET 1u /* ReverbEffectMessage::Reset (src/effect/reverb.rs:469-487) */

typedef struct pb200_event {
  uint64_t sample_time; /* absolute output frame, or PB200_TIME_NOW */
  uint32_t kind;        /* pb200_event_kind */
  uint32_t target;      /* playback / generator / effect id */
  uint64_t note_id;     /* NotePlaybackId for note events; out for NOTE_ON via pb200_schedule */
  uint32_t note;        /* MIDI note (NOTE_ON) */
  uint32_t param_id;    /* FourCC as big-endian u32, e.g. 'cuto' (SET_EFFECT_PARAMETER) */
  float value;          /* volume / panning / parameter value */
  float value2;         /* NOTE_ON: panning */
  float glide;          /* semitones per second; <= 0 => None */
  uint32_t flags;       /* PB200_EVF_* */
  double speed;         /* SET_*_SPEED */
  uint64_t position_nanos; /* SEEK position: Duration::as_nanos() */
} pb200_event;

#define PB200_EVF_NORMALIZED 1u /* ParameterValueUpdate::Normalized instead of ::Raw */
#define PB200_EVF_HAS_VOLUME 2u /* NOTE_ON: volume is Some(value) */
#define PB200_EVF_HAS_PANNING 4u/* NOTE_ON: panning is Some(value2) */
#define PB200_EVF_NO_RANGE 16u  /* SET_GENERATOR_LOOP_RANGE: the range is None */

/* Queue one event. For PB200_EV_NOTE_ON a fresh NotePlaybackId is allocated
 * (unique_note_id, src/generator.rs:30-33) and written back to `ev->note_id`. */
PB200_API int pb200_schedule(pb200_renderer *r, pb200_event *ev);

/* A whole score in one call: the events are queued in array order exactly as `count` pb200_schedule calls would (ids of
 * NOTE_ON events are written back). An event flagged PB200_EVF_NOTE_FROM_BATCH carries in `note_id` the INDEX (in this
 * array, lower than its own) of the NOTE_ON event whose freshly allocated id it refers to. Stops at the first error;
 * `*scheduled` = events queued. (The reference's handles push one message per call into lock-free queues,
 * src/player/handles/generator.rs:62-437; an offline score has no reason to cross the FFI once per event.) */
#define PB200_EVF_NOTE_FROM_BATCH 8u
PB200_API int pb200_schedule_many(pb200_renderer *r, pb200_event *events, uint32_t count, uint32_t *scheduled);

/* ---- structural messages: take effect at the next block start, i.e. with the next render call ------------------------
 * (MixedSource::process_messages runs at the top of every Source::write, src/source/mixed.rs:294-499) */
/* Player::remove_generator (src/player.rs:747-770) -> MixerMessage::RemoveSource (mixed.rs:391-393): the source is
 * dropped from its mixer, whatever it is playing. Also valid for file playbacks. */
PB200_API int pb200_remove_source(pb200_renderer *r, uint32_t playback_id);
/* Player::remove_mixer (src/player.rs:825-868) -> MixerMessage::RemoveMixer: the sub-mixer, its effects, sources and
 * sub-mixers disappear from the parent's sum. PB200_ERR_PARAMETER for the main mixer. */
PB200_API int pb200_remove_mixer(pb200_renderer *r, uint32_t mixer_id);
/* Player::remove_effect (src/player.rs:977-991) -> MixerMessage::RemoveEffect (mixed.rs:432-439) */
PB200_API int pb200_remove_effect(pb200_renderer *r, uint32_t effect_id);
/* Player::move_effect (src/player.rs:942-974) -> MixerMessage::MoveEffect (mixed.rs:440-459); EffectMovement:
 * Direction(offset) | Start | End. `mixer_id` must be the effect's mixer (PB200_ERR_PARAMETER otherwise). */
enum pb200_effect_movement { PB200_MOVE_DIRECTION = 0, PB200_MOVE_START = 1, PB200_MOVE_END = 2 };
PB200_API int pb200_move_effect(pb200_renderer *r, uint32_t effect_id, uint32_t mixer_id, uint32_t movement, int32_t offset);
/* Player::stop_all_sources (src/player.rs:1012-1045): Stop to every transient source + MixerMessage::
 * RemoveAllPendingEvents on every mixer (scheduled sources that have not started and all pending events are dropped,
 * mixed.rs:297-305). */
PB200_API int pb200_stop_all_sources(pb200_renderer *r);

/* ---- render: repeated WavStream::process (src/output/wav.rs:210-250) -----------------------
 * Renders `frames` output frames (interleaved f32, channel_count channels) as the reference
 * would in consecutive block_frames-sized `Source::write` calls on the main mixer, including
 * the master-volume smoothing of wav.rs:237. `out` is HOST memory. */
PB200_API int pb200_render(pb200_renderer *r, float *out_interleaved, uint64_t frames,
                           uint64_t *frames_written);
/* Multi-GPU renders: the main mixer of this renderer additionally receives `bus_device`, an interleaved
 * stereo f32 bus in DEVICE memory holding the summed output of the sub-mixers that were rendered elsewhere (the reduced
 * partial buses of the other ranks). It takes the place of those SubMixerProcessor outputs in MixedSource::write
 * (src/source/mixed.rs:701-716): added to the main mixer's input ahead of its own children, always audible. Frame 0 of
 * the bus is the first frame of the NEXT pb200_render* call, which consumes the attachment (the pointer is borrowed for
 * that one call); `frames` is a multiple of block_frames; nullptr detaches. */
PB200_API int pb200_set_main_input(pb200_renderer *r, const float *bus_device, uint64_t frames);
/* Same with up to PB200_MAX_MAIN_INPUTS buses, added in the order given (one per rank of a sharded render: the sum order is
 * fixed, unlike a tree reduce's). */
#define PB200_MAX_MAIN_INPUTS 8
PB200_API int pb200_set_main_inputs(pb200_renderer *r, const float *const *buses_device, uint32_t count, uint64_t frames);

/* ---- peer memory for sharded renders (CUDA build only; no reference counterpart: the reference's sub-mixer workers share
 * one address space, src/source/mixed/submixer/thread_pool.rs) --------------------------------------------------------------
 * A rank's finished piece travels to rank 0 as a stream-ordered DMA copy into memory rank 0 allocated and exported through
 * CUDA IPC, followed by a 4-byte flag: no kernel is involved, so the transfer does not queue behind the kernels of a device
 * that is busy rendering. pb200_device_alloc: zero-filled cudaMalloc memory (IPC-exportable, unlike a caching allocator's
 * sub-blocks); pb200_ipc_export / _open / _close: cudaIpcMemHandle_t as 64 opaque bytes; pb200_push_async: copy
 * `bytes` from this renderer's device to `dst_peer`, then store `flag_value` to `flag_peer` (may be NULL), in order, on the
 * renderer's copy stream; pb200_push_sync waits for all pushes; pb200_peek_u32 reads `count` words of device memory. */
/* The library keeps device memory of destroyed renderers in a process-wide pool (cudaMalloc / cudaFree synchronise the
 * device). pb200_trim_pool returns the idle blocks of one device (-1: all) to the driver and reports the bytes freed; the
 * pool also trims itself when an allocation fails. No reference counterpart. */
PB200_API uint64_t pb200_trim_pool(int device_ordinal);
PB200_API int pb200_device_alloc(int device_ordinal, size_t bytes, void **ptr);
PB200_API int pb200_device_free(void *ptr);
PB200_API int pb200_ipc_export(const void *ptr, void *handle64);
PB200_API int pb200_ipc_open(const void *handle64, int device_ordinal, void **ptr);
PB200_API int pb200_ipc_close(void *ptr);
PB200_API int pb200_push_async(pb200_renderer *r, void *dst_peer, const void *src_device, size_t bytes, uint32_t *flag_peer,
                               uint32_t flag_value);
PB200_API int pb200_push_sync(pb200_renderer *r);
PB200_API int pb200_peek_u32(pb200_renderer *r, const uint32_t *src_device, uint32_t count, uint32_t *out_host);
/* Output frames finalized so far, counted over ALL render calls of the renderer: after a call that rendered `frames` has
 * returned it has grown by `frames`; while a pb200_render_device call is still rendering it advances time block by time
 * block, and the frames it has passed are final in `out_device`. The one entry point that may be called from another host
 * thread while a render runs: a sharded render starts the reduce of a finished piece (and rank 0 its main-bus stage)
 * without waiting for the whole render. */
PB200_API uint64_t pb200_render_progress(const pb200_renderer *r);
/* Same as pb200_render, but `out` is DEVICE memory of the renderer's device (no host copy). CUDA build only. */
PB200_API int pb200_render_device(pb200_renderer *r, float *out_device, uint64_t frames,
                                  uint64_t *frames_written);
/* Player::output_sample_frame_position (src/player.rs) */
PB200_API uint64_t pb200_position(const pb200_renderer *r);

/* ---- WAV in / out: the steps either side of the path ------------------------------------------------
 * in : AudioFileBuffer::from_file for RIFF/WAVE (src/source/file/buffer.rs:64-119; `smpl` loops decoder.rs:294-330):
 *      PCM 8/16/24/32 and IEEE float 32/64, WAVE_FORMAT_EXTENSIBLE included; the first `smpl` loop becomes the
 *      buffer's loop range (clamped as the reference does, kept only if end > start).
 * out: WavOutput (src/output/wav.rs:50-120, 210-250): whole block_frames blocks while whole-seconds(pos / rate) <
 *      duration, written as 32-bit float WAV. */
typedef struct pb200_wav_info {
  uint64_t frames;          /* decoded frames (the +1 zero pad frame is added on upload) */
  uint32_t channels, sample_rate;
  int64_t loop_start, loop_end; /* frames, PB200_NO_LOOP when the file has no usable `smpl` loop */
  uint32_t bits_per_sample, is_float;
} pb200_wav_info;
/* Renderer-free decode into a malloc'd interleaved f32 buffer (release it with pb200_free). */
PB200_API int pb200_decode_wav(const char *path, float **interleaved, pb200_wav_info *info);
PB200_API void pb200_free(void *p);
/* decode + pb200_upload_buffer(add_pad_frame = 1, the file's loop). `info` may be NULL. */
PB200_API int pb200_upload_wav(pb200_renderer *r, const char *path, uint32_t *buffer_id, pb200_wav_info *info);
/* Player + WavOutput::open_with_specs(path, sample_rate, 2, duration): render and write the file.
 * duration_nanos = Duration::as_nanos(). */
PB200_API int pb200_render_to_wav(pb200_renderer *r, const char *path, uint64_t duration_nanos, uint64_t *frames_written);

/* ---- status: PlaybackStatusEvent::Stopped mirror (src/source/status.rs:15-36) --------------- */
typedef struct pb200_source_status {
  uint32_t is_playing;    /* FilePlaybackHandle::is_playing */
  uint32_t exhausted;     /* Stopped{exhausted} */
  uint64_t end_frame;     /* output frame at which the source finished (chunk end), or UINT64_MAX */
  uint64_t playback_pos;  /* PreloadedFileSource::playback_pos (sample index), files only */
} pb200_source_status;
PB200_API int pb200_source_status_get(pb200_renderer *r, uint32_t playback_id,
                                      pb200_source_status *st);

/* ---- observability: the PlaybackStatusEvent stream and the main mixer's level meter --------------------
 * PlaybackStatusEvent (src/source/status.rs:15-36) of file playbacks, in emission order: Position every second of output
 * time, reported at the end of the source's write call (FileSourceImpl::send_playback_position_status,
 * src/source/file/common.rs:171-208; FilePlaybackOptions' default playback_pos_emit_rate of 1 s), and Stopped when the
 * source finishes (preloaded.rs:196-209, 464-472). pb200_poll_status hands out and clears what the render calls so far
 * produced. */
enum pb200_status_kind { PB200_STATUS_POSITION = 0, PB200_STATUS_STOPPED = 1 };
typedef struct pb200_status_event {
  uint64_t frame;          /* output frame of the write call that emitted the event (its start; Stopped: its end) */
  uint32_t kind;           /* pb200_status_kind */
  uint32_t playback_id;
  uint64_t position_nanos; /* Position: Duration::from_secs_f64(playback frame / file rate).as_nanos() */
  uint32_t exhausted;      /* Stopped: played to its end (true) or was stopped (false) */
  uint32_t reserved;
} pb200_status_event;
PB200_API int pb200_poll_status(pb200_renderer *r, pb200_status_event *out, uint32_t capacity, uint32_t *count);

/* PlayerConfig::metering_interval + Player::audio_level (src/player.rs:166,216,464-471): MeteredSource around the main
 * mixer (src/source/metered.rs:107-148) -- per-channel peak and RMS of the main mixer's output, published whenever a
 * 1024-frame block starts `interval` or more after the last publication. PB200_DURATION_NONE switches metering off
 * (the default). On the device the per-block peak / sum of squares is a reduction fused into the main mixer's last
 * kernel. */
typedef struct pb200_audio_level {
  float peak[2];
  float rms[2];
} pb200_audio_level;
PB200_API int pb200_set_metering_interval(pb200_renderer *r, uint64_t interval_nanos);
PB200_API int pb200_get_audio_level(pb200_renderer *r, pb200_audio_level *out);

/* Voice-level introspection used by the parity tests: bit-exact integer state of every sampler
 * voice after the last render call (note id or UINT64_MAX, playback_pos sample index). */
typedef struct pb200_voice_state {
  uint64_t note_id;
  uint64_t playback_pos;
  uint32_t envelope_stage; /* AhdsrStage as 0 Idle 1 Attack 2 Hold 3 Decay 4 Sustain 5 Release */
  uint32_t active;
} pb200_voice_state;
PB200_API int pb200_sampler_voice_states(pb200_renderer *r, uint32_t generator_id,
                                         pb200_voice_state *out, uint32_t capacity,
                                         uint32_t *count);

/* ---- timing of the last render call (device side, CUDA events) ------------------------------ */
typedef struct pb200_render_stats {
  double device_ms;        /* all kernels of the last pb200_render* call */
  double voice_kernel_ms;  /* replay kernels (resample + gain/pan/envelope + ordered sum), summed over blocks;
                              NB: event-to-event spans on their own stream, they overlap the other passes */
  double skeleton_kernel_ms; /* control-only pass (events, phase/ramp recurrences, segment snapshots) */
  double effect_kernel_ms; /* mixer/effect kernels */
  uint64_t kernel_launches;
  uint64_t voice_frames;   /* active voice-frames rendered */
  double sinc_kernel_ms;   /* HighQuality resampler kernel (part of the voice_kernel_ms span) */
  double grain_kernel_ms;  /* granular grain kernel (part of the voice_kernel_ms span) */
  uint64_t sinc_frames;    /* resampler output frames the sinc kernel materialised */
  uint64_t grain_samples;  /* grain samples the grain kernel rendered */
} pb200_render_stats;
PB200_API int pb200_last_render_stats(pb200_renderer *r, pb200_render_stats *st);

#ifdef __cplusplus
}
#endif
#endif /* PHONIC_B200_H */
