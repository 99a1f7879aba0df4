//! `extern "C"` mirror of include/phonic_b200.h (every declared entry point; checked by
//! tests/test_rust_shim_matches_header.py).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct pb200_renderer {
    _private: [u8; 0],
}

pub const PB200_OK: c_int = 0;
pub const PB200_ERR_SOURCE_NOT_PLAYING: c_int = 1;
pub const PB200_ERR_MEDIA_FILE_NOT_FOUND: c_int = 2;
pub const PB200_ERR_MEDIA_FILE_PROBE: c_int = 3;
pub const PB200_ERR_MEDIA_FILE_SEEK: c_int = 4;
pub const PB200_ERR_AUDIO_DECODING: c_int = 5;
pub const PB200_ERR_OUTPUT_DEVICE: c_int = 6;
pub const PB200_ERR_RESAMPLING: c_int = 7;
pub const PB200_ERR_GENERATOR_NOT_FOUND: c_int = 8;
pub const PB200_ERR_EFFECT_NOT_FOUND: c_int = 9;
pub const PB200_ERR_MIXER_NOT_FOUND: c_int = 10;
pub const PB200_ERR_PARAMETER: c_int = 11;
pub const PB200_ERR_SEND: c_int = 12;
pub const PB200_ERR_IO: c_int = 13;
pub const PB200_ERR_CUDA: c_int = 100;
pub const PB200_ERR_UNSUPPORTED: c_int = 101;

pub const PB200_TIME_NOW: u64 = u64::MAX;
pub const PB200_MAIN_MIXER: u32 = 0;
pub const PB200_REPEAT_DEFAULT: u64 = u64::MAX - 1;
pub const PB200_REPEAT_FOREVER: u64 = u64::MAX;
pub const PB200_NO_LOOP: i64 = -1;
pub const PB200_DURATION_NONE: u64 = u64::MAX;

pub const PB200_FX_FILTER: u32 = 1;
pub const PB200_FX_EQ5: u32 = 2;
pub const PB200_FX_COMPRESSOR: u32 = 3;
pub const PB200_FX_CHORUS: u32 = 4;
pub const PB200_FX_DELAY: u32 = 5;
pub const PB200_FX_REVERB: u32 = 6;
pub const PB200_FX_GAIN: u32 = 7;
pub const PB200_FX_PANNING: u32 = 8;
pub const PB200_FX_GATE: u32 = 9;
pub const PB200_FX_DISTORTION: u32 = 10;

pub const PB200_EV_STOP_SOURCE: u32 = 1;
pub const PB200_EV_SET_SOURCE_VOLUME: u32 = 2;
pub const PB200_EV_SET_SOURCE_PANNING: u32 = 3;
pub const PB200_EV_SET_SOURCE_SPEED: u32 = 4;
pub const PB200_EV_SEEK_SOURCE: u32 = 5;
pub const PB200_EV_NOTE_ON: u32 = 10;
pub const PB200_EV_NOTE_OFF: u32 = 11;
pub const PB200_EV_ALL_NOTES_OFF: u32 = 12;
pub const PB200_EV_SET_NOTE_SPEED: u32 = 13;
pub const PB200_EV_SET_NOTE_VOLUME: u32 = 14;
pub const PB200_EV_SET_NOTE_PANNING: u32 = 15;
pub const PB200_EV_SET_GENERATOR_PARAMETER: u32 = 16;
pub const PB200_EV_SET_GENERATOR_LOOP_RANGE: u32 = 17;
pub const PB200_EV_SET_EFFECT_PARAMETER: u32 = 20;
pub const PB200_EV_EFFECT_MESSAGE: u32 = 21;
pub const PB200_MSG_REVERB_RESET: u32 = 1;

pub const PB200_EVF_NORMALIZED: u32 = 1;
pub const PB200_EVF_HAS_VOLUME: u32 = 2;
pub const PB200_EVF_HAS_PANNING: u32 = 4;
pub const PB200_EVF_NOTE_FROM_BATCH: u32 = 8;
pub const PB200_EVF_NO_RANGE: u32 = 16;

pub const PB200_MOVE_DIRECTION: u32 = 0;
pub const PB200_MOVE_START: u32 = 1;
pub const PB200_MOVE_END: u32 = 2;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_config {
    pub sample_rate: u32,
    pub channel_count: u32,
    pub block_frames: u32,
    pub device_ordinal: i32,
    pub master_volume: f32,
    pub reserved: [u32; 3],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_filter_params {
    pub filter_type: u32,
    pub cutoff: f32,
    pub q: f32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_compressor_params {
    pub threshold: f32,
    pub ratio: f32,
    pub knee: f32,
    pub attack_time: f32,
    pub release_time: f32,
    pub makeup_gain: f32,
    pub lookahead_time: f32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_chorus_params {
    pub rate: f32,
    pub phase: f32,
    pub depth: f32,
    pub feedback: f32,
    pub delay: f32,
    pub wet: f32,
    pub filter_type: u32,
    pub filter_freq: f32,
    pub filter_resonance: f32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_reverb_params {
    pub room_size: f32,
    pub wet: f32,
    pub fpd: [u32; 2],
    pub vib_phase: [f64; 16],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_file_options {
    pub volume: f32,
    pub panning: f32,
    pub speed: f64,
    pub repeat: u64,
    pub loop_start: i64,
    pub loop_end: i64,
    pub fade_in_nanos: u64,
    pub fade_out_nanos: u64,
    pub resampling_quality: u32,
    pub target_mixer: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_ahdsr {
    pub attack_nanos: u64,
    pub hold_nanos: u64,
    pub decay_nanos: u64,
    pub release_nanos: u64,
    pub attack_scaling: f32,
    pub decay_scaling: f32,
    pub release_scaling: f32,
    pub sustain_level: f32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_granular_params {
    pub overlap_mode: u32,
    pub window: u32,
    pub size: f32,
    pub density: f32,
    pub variation: f32,
    pub spray: f32,
    pub pan_spread: f32,
    pub playback_direction: u32,
    pub position: f32,
    pub step: f32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_sampler_options {
    pub volume: f32,
    pub panning: f32,
    pub voices: u32,
    pub target_mixer: u32,
    pub transient: u32,
    pub has_ahdsr: u32,
    pub ahdsr: pb200_ahdsr,
    pub has_granular: u32,
    pub granular: pb200_granular_params,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_event {
    pub sample_time: u64,
    pub kind: u32,
    pub target: u32,
    pub note_id: u64,
    pub note: u32,
    pub param_id: u32,
    pub value: f32,
    pub value2: f32,
    pub glide: f32,
    pub flags: u32,
    pub speed: f64,
    pub position_nanos: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_wav_info {
    pub frames: u64,
    pub channels: u32,
    pub sample_rate: u32,
    pub loop_start: i64,
    pub loop_end: i64,
    pub bits_per_sample: u32,
    pub is_float: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_source_status {
    pub is_playing: u32,
    pub exhausted: u32,
    pub end_frame: u64,
    pub playback_pos: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_voice_state {
    pub note_id: u64,
    pub playback_pos: u64,
    pub envelope_stage: u32,
    pub active: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_render_stats {
    pub device_ms: f64,
    pub voice_kernel_ms: f64,
    pub skeleton_kernel_ms: f64,
    pub effect_kernel_ms: f64,
    pub kernel_launches: u64,
    pub voice_frames: u64,
    pub sinc_kernel_ms: f64,
    pub grain_kernel_ms: f64,
    pub sinc_frames: u64,
    pub grain_samples: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_status_event {
    pub frame: u64,
    pub kind: u32,
    pub playback_id: u32,
    pub position_nanos: u64,
    pub exhausted: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_audio_level {
    pub peak: [f32; 2],
    pub rms: [f32; 2],
}

extern "C" {
    pub fn pb200_backend() -> *const c_char;
    pub fn pb200_create(config: *const pb200_config, out: *mut *mut pb200_renderer) -> c_int;
    pub fn pb200_destroy(r: *mut pb200_renderer);
    pub fn pb200_last_error(r: *const pb200_renderer) -> *const c_char;
    pub fn pb200_upload_buffer(r: *mut pb200_renderer, interleaved: *const f32, frames: u64, channels: u32, sample_rate: u32,
                               loop_start: i64, loop_end: i64, add_pad_frame: c_int, buffer_id: *mut u32) -> c_int;
    pub fn pb200_add_mixer(r: *mut pb200_renderer, parent: u32, mixer_id: *mut u32) -> c_int;
    pub fn pb200_add_effect(r: *mut pb200_renderer, mixer: u32, kind: u32, params: *const c_void, size: usize, effect_id: *mut u32) -> c_int;
    pub fn pb200_file_options_default(o: *mut pb200_file_options);
    pub fn pb200_play_file(r: *mut pb200_renderer, buffer_id: u32, o: *const pb200_file_options, start_time: u64, playback_id: *mut u32) -> c_int;
    pub fn pb200_sampler_options_default(o: *mut pb200_sampler_options);
    pub fn pb200_add_sampler(r: *mut pb200_renderer, buffer_id: u32, o: *const pb200_sampler_options, start_time: u64, generator_id: *mut u32) -> c_int;
    pub fn pb200_schedule(r: *mut pb200_renderer, ev: *mut pb200_event) -> c_int;
    pub fn pb200_schedule_many(r: *mut pb200_renderer, events: *mut pb200_event, count: u32, scheduled: *mut u32) -> c_int;
    pub fn pb200_remove_source(r: *mut pb200_renderer, playback_id: u32) -> c_int;
    pub fn pb200_remove_mixer(r: *mut pb200_renderer, mixer_id: u32) -> c_int;
    pub fn pb200_remove_effect(r: *mut pb200_renderer, effect_id: u32) -> c_int;
    pub fn pb200_move_effect(r: *mut pb200_renderer, effect_id: u32, mixer_id: u32, movement: u32, offset: i32) -> c_int;
    pub fn pb200_stop_all_sources(r: *mut pb200_renderer) -> c_int;
    pub fn pb200_render(r: *mut pb200_renderer, out_interleaved: *mut f32, frames: u64, frames_written: *mut u64) -> c_int;
    pub fn pb200_render_progress(r: *const pb200_renderer) -> u64;
    pub fn pb200_set_main_inputs(r: *mut pb200_renderer, buses_device: *const *const f32, count: u32, frames: u64) -> c_int;
    pub fn pb200_trim_pool(device_ordinal: c_int) -> u64;
    pub fn pb200_device_alloc(device_ordinal: c_int, bytes: usize, ptr: *mut *mut c_void) -> c_int;
    pub fn pb200_device_free(ptr: *mut c_void) -> c_int;
    pub fn pb200_ipc_export(ptr: *const c_void, handle64: *mut c_void) -> c_int;
    pub fn pb200_ipc_open(handle64: *const c_void, device_ordinal: c_int, ptr: *mut *mut c_void) -> c_int;
    pub fn pb200_ipc_close(ptr: *mut c_void) -> c_int;
    pub fn pb200_push_async(r: *mut pb200_renderer, dst_peer: *mut c_void, src_device: *const c_void, bytes: usize, flag_peer: *mut u32, flag_value: u32) -> c_int;
    pub fn pb200_push_sync(r: *mut pb200_renderer) -> c_int;
    pub fn pb200_peek_u32(r: *mut pb200_renderer, src_device: *const u32, count: u32, out_host: *mut u32) -> c_int;
    pub fn pb200_set_main_input(r: *mut pb200_renderer, bus_device: *const f32, frames: u64) -> c_int;
    pub fn pb200_render_device(r: *mut pb200_renderer, out_device: *mut f32, frames: u64, frames_written: *mut u64) -> c_int;
    pub fn pb200_position(r: *const pb200_renderer) -> u64;
    pub fn pb200_decode_wav(path: *const c_char, interleaved: *mut *mut f32, info: *mut pb200_wav_info) -> c_int;
    pub fn pb200_free(p: *mut c_void);
    pub fn pb200_upload_wav(r: *mut pb200_renderer, path: *const c_char, buffer_id: *mut u32, info: *mut pb200_wav_info) -> c_int;
    pub fn pb200_render_to_wav(r: *mut pb200_renderer, path: *const c_char, duration_nanos: u64, frames_written: *mut u64) -> c_int;
    pub fn pb200_source_status_get(r: *mut pb200_renderer, playback_id: u32, st: *mut pb200_source_status) -> c_int;
    pub fn pb200_sampler_voice_states(r: *mut pb200_renderer, generator_id: u32, out: *mut pb200_voice_state, capacity: u32, count: *mut u32) -> c_int;
    pub fn pb200_poll_status(r: *mut pb200_renderer, out: *mut pb200_status_event, capacity: u32, count: *mut u32) -> c_int;
    pub fn pb200_set_metering_interval(r: *mut pb200_renderer, interval_nanos: u64) -> c_int;
    pub fn pb200_get_audio_level(r: *mut pb200_renderer, out: *mut pb200_audio_level) -> c_int;
    pub fn pb200_last_render_stats(r: *mut pb200_renderer, st: *mut pb200_render_stats) -> c_int;
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_gain_params {
    pub gain_db: f32,
    pub dc_filter_mode: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_gate_params {
    pub threshold: f32,
    pub attack_time: f32,
    pub hold_time: f32,
    pub release_time: f32,
    pub range: f32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pb200_distortion_params {
    pub distortion_type: u32,
    pub drive: f32,
    pub mix: f32,
}
