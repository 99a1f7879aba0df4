//! phonic-b200: phonic's offline (WAV output) render path on an NVIDIA B200.
//!
//! `B200Player` mirrors the subset of `phonic::Player` (src/player.rs:274-1046) that builds and drives the mixer
//! graph -- same method names, same argument meaning, same `phonic::Error` variants -- and forwards to the C-ABI
//! of include/phonic_b200.h. `B200Output` is the matching `phonic::OutputDevice` (src/output.rs:33-65) for code
//! that wants to keep a `phonic::Player` for control and only swap the device: it renders nothing itself (the
//! graph lives on the GPU) but reports the WavOutput-compatible format and position.
//!
//! No CPU fallback: `B200Player::new` fails with `Error::OutputDeviceError` when there is no CUDA device.
pub mod sys;

use std::{
    collections::HashMap,
    ffi::{CStr, CString},
    path::Path,
    ptr,
    sync::{Arc, Mutex},
    time::Duration,
};

use phonic::{
    AudioFileBuffer, EffectId, EffectMovement, Error, FilePlaybackOptions, GeneratorPlaybackOptions, MixerId, NotePlaybackId,
    OutputDevice, PlaybackId, ResamplingQuality, Source,
};

/// The renderer handle, shared by the player and its playback handles (`&mut self` semantics of the C-ABI are kept
/// by the mutex; the reference's handles are `Send + Sync` too).
struct Renderer(*mut sys::pb200_renderer);
unsafe impl Send for Renderer {}
impl Drop for Renderer {
    fn drop(&mut self) {
        unsafe { sys::pb200_destroy(self.0) }
    }
}
type Shared = Arc<Mutex<Renderer>>;

fn check(code: i32, r: *const sys::pb200_renderer) -> Result<(), Error> {
    if code == sys::PB200_OK {
        return Ok(());
    }
    let msg = unsafe {
        let p = sys::pb200_last_error(r);
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    };
    // return codes are 1:1 with phonic::Error (src/error.rs:8-22)
    Err(match code {
        sys::PB200_ERR_SOURCE_NOT_PLAYING => Error::SourceNotPlaying,
        sys::PB200_ERR_MEDIA_FILE_NOT_FOUND => Error::MediaFileNotFound,
        sys::PB200_ERR_MEDIA_FILE_PROBE => Error::MediaFileProbeError,
        sys::PB200_ERR_MEDIA_FILE_SEEK => Error::MediaFileSeekError,
        sys::PB200_ERR_AUDIO_DECODING => Error::AudioDecodingError(msg.into()),
        sys::PB200_ERR_OUTPUT_DEVICE => Error::OutputDeviceError(msg.into()),
        sys::PB200_ERR_IO => Error::IoError(std::io::Error::new(std::io::ErrorKind::Other, msg)),
        sys::PB200_ERR_RESAMPLING => Error::ResamplingError(msg.into()),
        sys::PB200_ERR_GENERATOR_NOT_FOUND => Error::GeneratorNotFoundError(0),
        sys::PB200_ERR_EFFECT_NOT_FOUND => Error::EffectNotFoundError(0),
        sys::PB200_ERR_MIXER_NOT_FOUND => Error::MixerNotFoundError(0),
        sys::PB200_ERR_SEND => Error::SendError(msg),
        sys::PB200_ERR_CUDA => Error::OutputDeviceError(msg.into()),
        _ => Error::ParameterError(msg),
    })
}

fn nanos(d: Option<Duration>) -> u64 {
    d.map(|d| d.as_nanos() as u64).unwrap_or(sys::PB200_DURATION_NONE)
}

/// Effects that can be added to a mixer: the constructors of src/effect/*.rs with their reference arguments.
pub enum B200Effect {
    Filter { filter_type: u32, cutoff: f32, q: f32 },
    Eq5,
    Compressor { threshold: f32, ratio: f32, knee_width: f32, attack_time: f32, release_time: f32, makeup_gain: f32, lookahead_time: f32 },
    Chorus { rate: f32, phase: f32, depth: f32, feedback: f32, delay: f32, wet_mix: f32, filter_type: u32, filter_freq: f32, filter_resonance: f32 },
    Delay,
    /// `ReverbEffect::with_parameters(room_size, wet)` + the state the reference draws from `rand::rng()`
    Reverb { room_size: f32, wet: f32, fpd: [u32; 2], vib_phase: [f64; 16] },
    Gain { gain_db: f32, dc_filter_mode: u32 },
    Panning,
    Gate { threshold: f32, attack_time: f32, hold_time: f32, release_time: f32, range: f32 },
    Distortion { distortion_type: u32, drive: f32, mix: f32 },
}

/// `phonic::Player` for the offline path, rendering on the GPU.
pub struct B200Player {
    r: Shared,
    sample_rate: u32,
    buffers: HashMap<usize, u32>, // Arc<AudioFileBuffer> address -> uploaded buffer id
}

impl B200Player {
    pub const MAIN_MIXER_ID: MixerId = 0;

    /// `Player::new(WavOutput::open_with_specs(_, sample_rate, 2, _), None)`
    pub fn new(sample_rate: u32) -> Result<Self, Error> {
        Self::new_on_device(sample_rate, -1)
    }
    pub fn new_on_device(sample_rate: u32, device_ordinal: i32) -> Result<Self, Error> {
        let cfg = sys::pb200_config { sample_rate, channel_count: 2, block_frames: 1024, device_ordinal, master_volume: 1.0, reserved: [0; 3] };
        let mut r = ptr::null_mut();
        check(unsafe { sys::pb200_create(&cfg, &mut r) }, r)?;
        Ok(Self { r: Arc::new(Mutex::new(Renderer(r))), sample_rate, buffers: HashMap::new() })
    }

    fn with<T>(&self, f: impl FnOnce(*mut sys::pb200_renderer) -> T) -> T {
        let g = self.r.lock().unwrap();
        f(g.0)
    }

    pub fn output_sample_rate(&self) -> u32 { self.sample_rate }
    pub fn output_channel_count(&self) -> usize { 2 }
    pub fn output_sample_frame_position(&self) -> u64 { self.with(|r| unsafe { sys::pb200_position(r) }) }

    /// Uploads a decoded file once per `Arc<AudioFileBuffer>` (the reference shares it between clones,
    /// preloaded.rs:118-135). Decoded buffers already carry the +1 zero pad frame (buffer.rs:103-104).
    pub fn upload(&mut self, buffer: &Arc<AudioFileBuffer>) -> Result<u32, Error> {
        let key = Arc::as_ptr(buffer) as usize;
        if let Some(id) = self.buffers.get(&key) {
            return Ok(*id);
        }
        let (ls, le) = buffer.loop_range().map(|r| (r.start as i64, r.end as i64)).unwrap_or((sys::PB200_NO_LOOP, sys::PB200_NO_LOOP));
        let mut id = 0u32;
        self.with(|r| check(unsafe {
            sys::pb200_upload_buffer(r, buffer.buffer().as_ptr(), buffer.frame_count() as u64, buffer.channel_count() as u32,
                                     buffer.sample_rate(), ls, le, 0, &mut id)
        }, r))?;
        self.buffers.insert(key, id);
        Ok(id)
    }

    /// `Player::play_file(path, options)` for RIFF/WAVE files decoded by the renderer's own reader
    pub fn play_file<P: AsRef<Path>>(&mut self, path: P, options: FilePlaybackOptions) -> Result<B200FileHandle, Error> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes()).map_err(|e| Error::ParameterError(e.to_string()))?;
        let mut id = 0u32;
        self.with(|r| check(unsafe { sys::pb200_upload_wav(r, c.as_ptr(), &mut id, ptr::null_mut()) }, r))?;
        self.play_buffer_id(id, options, None)
    }

    /// `Player::play_file_source(PreloadedFileSource::from_shared_buffer(buffer, ..), start_time)`
    pub fn play_file_buffer<T: Into<Option<u64>>>(&mut self, buffer: &Arc<AudioFileBuffer>, options: FilePlaybackOptions, start_time: T)
        -> Result<B200FileHandle, Error> {
        let id = self.upload(buffer)?;
        self.play_buffer_id(id, options, start_time.into())
    }

    fn play_buffer_id(&mut self, buffer_id: u32, o: FilePlaybackOptions, start_time: Option<u64>) -> Result<B200FileHandle, Error> {
        o.validate()?;
        let fo = sys::pb200_file_options {
            volume: o.volume,
            panning: o.panning,
            speed: o.speed,
            repeat: match o.repeat { None => sys::PB200_REPEAT_DEFAULT, Some(usize::MAX) => sys::PB200_REPEAT_FOREVER, Some(n) => n as u64 },
            loop_start: o.loop_range.map(|(s, _)| s as i64).unwrap_or(sys::PB200_NO_LOOP),
            loop_end: o.loop_range.map(|(_, e)| e as i64).unwrap_or(sys::PB200_NO_LOOP),
            fade_in_nanos: nanos(o.fade_in_duration),
            fade_out_nanos: nanos(o.fade_out_duration),
            resampling_quality: match o.resampling_quality { ResamplingQuality::HighQuality => 1, _ => 0 },
            target_mixer: o.target_mixer.unwrap_or(Self::MAIN_MIXER_ID) as u32,
        };
        let start = start_time.or(o.start_time).unwrap_or(sys::PB200_TIME_NOW);
        let mut id = 0u32;
        self.with(|r| check(unsafe { sys::pb200_play_file(r, buffer_id, &fo, start, &mut id) }, r))?;
        Ok(B200FileHandle { r: self.r.clone(), id })
    }

    /// `Player::add_generator(Sampler::from_file_source(..).with_ahdsr(..), mixer)` / `play_generator` (transient)
    #[allow(clippy::too_many_arguments)]
    pub fn add_sampler(&mut self, buffer: &Arc<AudioFileBuffer>, options: GeneratorPlaybackOptions, ahdsr: Option<sys::pb200_ahdsr>,
                       granular: Option<sys::pb200_granular_params>, transient: bool, start_time: Option<u64>)
        -> Result<B200GeneratorHandle, Error> {
        let buffer_id = self.upload(buffer)?;
        let mut so: sys::pb200_sampler_options = unsafe { std::mem::zeroed() };
        unsafe { sys::pb200_sampler_options_default(&mut so) };
        so.volume = options.volume;
        so.panning = options.panning;
        so.voices = options.voices as u32;
        so.target_mixer = options.target_mixer.unwrap_or(Self::MAIN_MIXER_ID) as u32;
        so.transient = transient as u32;
        if let Some(a) = ahdsr { so.has_ahdsr = 1; so.ahdsr = a; }
        if let Some(g) = granular { so.has_granular = 1; so.granular = g; }
        let mut id = 0u32;
        self.with(|r| check(unsafe { sys::pb200_add_sampler(r, buffer_id, &so, start_time.unwrap_or(sys::PB200_TIME_NOW), &mut id) }, r))?;
        Ok(B200GeneratorHandle { r: self.r.clone(), id })
    }
    pub fn remove_generator(&mut self, playback_id: PlaybackId) -> Result<(), Error> {
        self.with(|r| check(unsafe { sys::pb200_remove_source(r, playback_id as u32) }, r))
    }

    pub fn add_mixer<M: Into<Option<MixerId>>>(&mut self, parent: M) -> Result<MixerId, Error> {
        let mut id = 0u32;
        let parent = parent.into().unwrap_or(Self::MAIN_MIXER_ID) as u32;
        self.with(|r| check(unsafe { sys::pb200_add_mixer(r, parent, &mut id) }, r))?;
        Ok(id as MixerId)
    }
    pub fn remove_mixer(&mut self, mixer_id: MixerId) -> Result<(), Error> {
        self.with(|r| check(unsafe { sys::pb200_remove_mixer(r, mixer_id as u32) }, r))
    }

    pub fn add_effect<M: Into<Option<MixerId>>>(&mut self, effect: B200Effect, mixer_id: M) -> Result<B200EffectHandle, Error> {
        use std::ffi::c_void;
        let mixer = mixer_id.into().unwrap_or(Self::MAIN_MIXER_ID) as u32;
        let mut id = 0u32;
        let add = |r: *mut sys::pb200_renderer, kind: u32, p: *const c_void, n: usize, id: &mut u32| check(unsafe { sys::pb200_add_effect(r, mixer, kind, p, n, id) }, r);
        self.with(|r| match effect {
            B200Effect::Filter { filter_type, cutoff, q } => {
                let p = sys::pb200_filter_params { filter_type, cutoff, q };
                add(r, sys::PB200_FX_FILTER, &p as *const _ as *const c_void, std::mem::size_of_val(&p), &mut id)
            }
            B200Effect::Eq5 => add(r, sys::PB200_FX_EQ5, ptr::null(), 0, &mut id),
            B200Effect::Compressor { threshold, ratio, knee_width, attack_time, release_time, makeup_gain, lookahead_time } => {
                let p = sys::pb200_compressor_params { threshold, ratio, knee: knee_width, attack_time, release_time, makeup_gain, lookahead_time };
                add(r, sys::PB200_FX_COMPRESSOR, &p as *const _ as *const c_void, std::mem::size_of_val(&p), &mut id)
            }
            B200Effect::Chorus { rate, phase, depth, feedback, delay, wet_mix, filter_type, filter_freq, filter_resonance } => {
                let p = sys::pb200_chorus_params { rate, phase, depth, feedback, delay, wet: wet_mix, filter_type, filter_freq, filter_resonance };
                add(r, sys::PB200_FX_CHORUS, &p as *const _ as *const c_void, std::mem::size_of_val(&p), &mut id)
            }
            B200Effect::Delay => add(r, sys::PB200_FX_DELAY, ptr::null(), 0, &mut id),
            B200Effect::Reverb { room_size, wet, fpd, vib_phase } => {
                let p = sys::pb200_reverb_params { room_size, wet, fpd, vib_phase };
                add(r, sys::PB200_FX_REVERB, &p as *const _ as *const c_void, std::mem::size_of_val(&p), &mut id)
            }
            B200Effect::Gain { gain_db, dc_filter_mode } => {
                let p = sys::pb200_gain_params { gain_db, dc_filter_mode };
                add(r, sys::PB200_FX_GAIN, &p as *const _ as *const c_void, std::mem::size_of_val(&p), &mut id)
            }
            B200Effect::Panning => add(r, sys::PB200_FX_PANNING, ptr::null(), 0, &mut id),
            B200Effect::Gate { threshold, attack_time, hold_time, release_time, range } => {
                let p = sys::pb200_gate_params { threshold, attack_time, hold_time, release_time, range };
                add(r, sys::PB200_FX_GATE, &p as *const _ as *const c_void, std::mem::size_of_val(&p), &mut id)
            }
            B200Effect::Distortion { distortion_type, drive, mix } => {
                let p = sys::pb200_distortion_params { distortion_type, drive, mix };
                add(r, sys::PB200_FX_DISTORTION, &p as *const _ as *const c_void, std::mem::size_of_val(&p), &mut id)
            }
        })?;
        Ok(B200EffectHandle { r: self.r.clone(), id })
    }
    pub fn move_effect<M: Into<Option<MixerId>>>(&mut self, movement: EffectMovement, effect_id: EffectId, mixer_id: M) -> Result<(), Error> {
        let (kind, off) = match movement {
            EffectMovement::Direction(o) => (sys::PB200_MOVE_DIRECTION, o),
            EffectMovement::Start => (sys::PB200_MOVE_START, 0),
            EffectMovement::End => (sys::PB200_MOVE_END, 0),
        };
        let mixer = mixer_id.into().unwrap_or(Self::MAIN_MIXER_ID) as u32;
        self.with(|r| check(unsafe { sys::pb200_move_effect(r, effect_id as u32, mixer, kind, off) }, r))
    }
    pub fn remove_effect(&mut self, effect_id: EffectId) -> Result<(), Error> {
        self.with(|r| check(unsafe { sys::pb200_remove_effect(r, effect_id as u32) }, r))
    }
    pub fn stop_all_sources(&mut self) -> Result<(), Error> {
        self.with(|r| check(unsafe { sys::pb200_stop_all_sources(r) }, r))
    }

    /// The next `frames` output frames (a multiple of 1024), interleaved stereo f32: `frames / 1024` WavStream blocks.
    pub fn render(&mut self, out: &mut [f32]) -> Result<u64, Error> {
        let frames = (out.len() / 2) as u64;
        let mut written = 0u64;
        self.with(|r| check(unsafe { sys::pb200_render(r, out.as_mut_ptr(), frames, &mut written) }, r))?;
        Ok(written)
    }

    /// `Player::new(WavOutput::open_with_specs(path, sr, 2, duration), ..)` run to the end (src/output/wav.rs:50-120)
    pub fn render_to_wav<P: AsRef<Path>>(&mut self, path: P, duration: Duration) -> Result<u64, Error> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes()).map_err(|e| Error::ParameterError(e.to_string()))?;
        let mut written = 0u64;
        self.with(|r| check(unsafe { sys::pb200_render_to_wav(r, c.as_ptr(), duration.as_nanos() as u64, &mut written) }, r))?;
        Ok(written)
    }
}

fn event(kind: u32, target: u32, sample_time: Option<u64>) -> sys::pb200_event {
    let mut ev: sys::pb200_event = unsafe { std::mem::zeroed() };
    ev.kind = kind;
    ev.target = target;
    ev.sample_time = sample_time.unwrap_or(sys::PB200_TIME_NOW);
    ev
}
fn schedule(r: &Shared, ev: &mut sys::pb200_event) -> Result<(), Error> {
    let g = r.lock().unwrap();
    check(unsafe { sys::pb200_schedule(g.0, ev) }, g.0)
}

/// `FilePlaybackHandle` (src/player/handles/file.rs:31-268)
pub struct B200FileHandle { r: Shared, id: u32 }
impl B200FileHandle {
    pub fn id(&self) -> PlaybackId { self.id as PlaybackId }
    pub fn is_playing(&self) -> bool {
        let g = self.r.lock().unwrap();
        let mut st: sys::pb200_source_status = unsafe { std::mem::zeroed() };
        unsafe { sys::pb200_source_status_get(g.0, self.id, &mut st) == sys::PB200_OK && st.is_playing != 0 }
    }
    pub fn stop<T: Into<Option<u64>>>(&self, stop_time: T) -> Result<(), Error> {
        schedule(&self.r, &mut event(sys::PB200_EV_STOP_SOURCE, self.id, stop_time.into()))
    }
    pub fn seek<T: Into<Option<u64>>>(&self, position: Duration, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_SEEK_SOURCE, self.id, sample_time.into());
        ev.position_nanos = position.as_nanos() as u64;
        schedule(&self.r, &mut ev)
    }
    pub fn set_speed<T: Into<Option<u64>>>(&self, speed: f64, glide: Option<f32>, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_SET_SOURCE_SPEED, self.id, sample_time.into());
        ev.speed = speed;
        ev.glide = glide.unwrap_or(0.0);
        schedule(&self.r, &mut ev)
    }
    pub fn set_volume<T: Into<Option<u64>>>(&self, volume: f32, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_SET_SOURCE_VOLUME, self.id, sample_time.into());
        ev.value = volume;
        schedule(&self.r, &mut ev)
    }
    pub fn set_panning<T: Into<Option<u64>>>(&self, panning: f32, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_SET_SOURCE_PANNING, self.id, sample_time.into());
        ev.value = panning;
        schedule(&self.r, &mut ev)
    }
}

/// `GeneratorPlaybackHandle` (src/player/handles/generator.rs:62-437)
pub struct B200GeneratorHandle { r: Shared, id: u32 }
impl B200GeneratorHandle {
    pub fn id(&self) -> PlaybackId { self.id as PlaybackId }
    pub fn stop<T: Into<Option<u64>>>(&self, stop_time: T) -> Result<(), Error> {
        schedule(&self.r, &mut event(sys::PB200_EV_STOP_SOURCE, self.id, stop_time.into()))
    }
    pub fn note_on<T: Into<Option<u64>>>(&self, note: u8, volume: Option<f32>, panning: Option<f32>, sample_time: T) -> Result<NotePlaybackId, Error> {
        let mut ev = event(sys::PB200_EV_NOTE_ON, self.id, sample_time.into());
        ev.note = note as u32;
        if let Some(v) = volume { ev.value = v; ev.flags |= sys::PB200_EVF_HAS_VOLUME; }
        if let Some(p) = panning { ev.value2 = p; ev.flags |= sys::PB200_EVF_HAS_PANNING; }
        schedule(&self.r, &mut ev)?;
        Ok(ev.note_id as NotePlaybackId)
    }
    pub fn note_off<T: Into<Option<u64>>>(&self, note_id: NotePlaybackId, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_NOTE_OFF, self.id, sample_time.into());
        ev.note_id = note_id as u64;
        schedule(&self.r, &mut ev)
    }
    pub fn all_notes_off<T: Into<Option<u64>>>(&self, sample_time: T) -> Result<(), Error> {
        schedule(&self.r, &mut event(sys::PB200_EV_ALL_NOTES_OFF, self.id, sample_time.into()))
    }
    pub fn set_note_speed<T: Into<Option<u64>>>(&self, note_id: NotePlaybackId, speed: f64, glide: Option<f32>, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_SET_NOTE_SPEED, self.id, sample_time.into());
        ev.note_id = note_id as u64;
        ev.speed = speed;
        ev.glide = glide.unwrap_or(0.0);
        schedule(&self.r, &mut ev)
    }
    pub fn set_note_volume<T: Into<Option<u64>>>(&self, note_id: NotePlaybackId, volume: f32, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_SET_NOTE_VOLUME, self.id, sample_time.into());
        ev.note_id = note_id as u64;
        ev.value = volume;
        schedule(&self.r, &mut ev)
    }
    pub fn set_note_panning<T: Into<Option<u64>>>(&self, note_id: NotePlaybackId, panning: f32, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_SET_NOTE_PANNING, self.id, sample_time.into());
        ev.note_id = note_id as u64;
        ev.value = panning;
        schedule(&self.r, &mut ev)
    }
    /// `set_parameter((id, ParameterValueUpdate::Raw | Normalized), t)`; `id` as four ASCII bytes, e.g. *b"STRN"
    pub fn set_parameter<T: Into<Option<u64>>>(&self, id: [u8; 4], value: f32, normalized: bool, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_SET_GENERATOR_PARAMETER, self.id, sample_time.into());
        ev.param_id = u32::from_be_bytes(id);
        ev.value = value;
        if normalized { ev.flags |= sys::PB200_EVF_NORMALIZED; }
        schedule(&self.r, &mut ev)
    }
    /// `send_message(SamplerMessage::SetLoopRange(range), t)`
    pub fn set_loop_range<T: Into<Option<u64>>>(&self, range: Option<std::ops::Range<u64>>, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_SET_GENERATOR_LOOP_RANGE, self.id, sample_time.into());
        match range {
            Some(r) => { ev.position_nanos = r.start; ev.note_id = r.end; }
            None => ev.flags |= sys::PB200_EVF_NO_RANGE,
        }
        schedule(&self.r, &mut ev)
    }
}

/// `EffectHandle` (src/player/handles/effect.rs:47-163)
pub struct B200EffectHandle { r: Shared, id: u32 }
impl B200EffectHandle {
    pub fn id(&self) -> EffectId { self.id as EffectId }
    pub fn set_parameter<T: Into<Option<u64>>>(&self, id: [u8; 4], value: f32, normalized: bool, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_SET_EFFECT_PARAMETER, self.id, sample_time.into());
        ev.param_id = u32::from_be_bytes(id);
        ev.value = value;
        if normalized { ev.flags |= sys::PB200_EVF_NORMALIZED; }
        schedule(&self.r, &mut ev)
    }
    /// `send_message(ReverbEffectMessage::Reset, t)`
    pub fn reset_reverb<T: Into<Option<u64>>>(&self, sample_time: T) -> Result<(), Error> {
        let mut ev = event(sys::PB200_EV_EFFECT_MESSAGE, self.id, sample_time.into());
        ev.param_id = sys::PB200_MSG_REVERB_RESET;
        schedule(&self.r, &mut ev)
    }
}

/// `OutputDevice` for a stock `phonic::Player`: format and clock of a WavOutput (stereo, the renderer's rate), no
/// audio thread. The graph a `phonic::Player` hands to `play()` is CPU code and is dropped: build the graph through
/// `B200Player` instead. Exists so that code written against `OutputDevice` (volume, position, pause / resume)
/// keeps working when the device is swapped.
pub struct B200Output { sample_rate: u32, volume: f32, running: bool, player: Shared }
impl B200Output {
    pub fn open(player: &B200Player) -> Self {
        Self { sample_rate: player.sample_rate, volume: 1.0, running: false, player: player.r.clone() }
    }
}
impl OutputDevice for B200Output {
    fn channel_count(&self) -> usize { 2 }
    fn sample_rate(&self) -> u32 { self.sample_rate }
    fn sample_position(&self) -> u64 {
        let g = self.player.lock().unwrap();
        unsafe { sys::pb200_position(g.0) * 2 }
    }
    fn volume(&self) -> f32 { self.volume }
    fn set_volume(&mut self, volume: f32) { self.volume = volume; }
    fn is_suspended(&self) -> bool { false }
    fn is_running(&self) -> bool { self.running }
    fn pause(&mut self) { self.running = false; }
    fn resume(&mut self) { self.running = true; }
    fn play(&mut self, _source: Box<dyn Source>) { self.running = true; }
    fn stop(&mut self) { self.running = false; }
    fn close(&mut self) { self.running = false; }
}
