//! Renders the shared parity scenes (tests/golden/reference/scenes.json, written by `python tests/reference_scenes.py
//! --write`) with UPSTREAM phonic on the CPU and writes tests/golden/reference/out/<scene>.wav. Those files are the
//! reference-run fixtures tests/test_reference_fixtures.py pins the oracle against (SURVEY.md §8c deliverable 4).
//!
//!   cargo test --manifest-path rust/Cargo.toml --features dump-reference --test dump_reference -- --nocapture
//!
//! Needs only phonic itself (the GPU library is not touched). Not runnable in the image this repo is developed in
//! (no cargo); anyone with a Rust toolchain can close the loop and commit the `out/` directory.
#![cfg(feature = "dump-reference")]

use std::{collections::HashMap, fs, path::PathBuf, sync::Arc, time::Duration};

use phonic::{
    effects::{ChorusEffect, ChorusEffectFilterType, CompressorEffect, DelayEffect, Eq5Effect, FilterEffect, FilterEffectType},
    generators::{AhdsrParameters, GrainOverlapMode, GrainPlaybackDirection, GrainWindowMode, GranularParameters, Sampler},
    outputs::WavOutput,
    parameters::FloatParameter,
    sources::PreloadedFileSource,
    AudioFileBuffer, Error, FilePlaybackOptions, GeneratorPlaybackOptions, MixerId, Player, ResamplingQuality,
};
use serde_json::Value;

fn repo() -> PathBuf {
    PathBuf::from(env!("CARGO_MANIFEST_DIR")).join("..")
}

fn load_buffer(name: &str, spec: &Value) -> Arc<AudioFileBuffer> {
    let raw = fs::read(repo().join("tests/golden/reference/inputs").join(format!("{name}.f32"))).expect("input buffer");
    let mut samples: Vec<f32> = raw.chunks_exact(4).map(|b| f32::from_le_bytes([b[0], b[1], b[2], b[3]])).collect();
    let channels = spec["channels"].as_u64().unwrap() as usize;
    // AudioFileBuffer::from_audio_decoder appends one zero frame for the cubic resampler (buffer.rs:103-104); the
    // Python side uploads with add_pad_frame = 1
    samples.extend(std::iter::repeat(0.0).take(channels));
    let loop_range = spec["loop"].as_array().map(|l| l[0].as_u64().unwrap() as usize..l[1].as_u64().unwrap() as usize);
    Arc::new(AudioFileBuffer::new(samples, channels, spec["rate"].as_u64().unwrap() as u32, loop_range).expect("buffer"))
}

fn secs(v: &Value) -> Option<Duration> {
    v.as_f64().map(Duration::from_secs_f64)
}

fn add_effects(player: &mut Player, effects: &Value, mixer: Option<MixerId>) -> Result<(), Error> {
    for e in effects.as_array().unwrap() {
        let f = |k: &str| e[k].as_f64().unwrap() as f32;
        match e["kind"].as_str().unwrap() {
            "filter" => {
                let t = match e["type"].as_u64().unwrap() {
                    0 => FilterEffectType::Lowpass,
                    1 => FilterEffectType::Bandpass,
                    2 => FilterEffectType::Bandstop,
                    _ => FilterEffectType::Highpass,
                };
                player.add_effect(FilterEffect::with_parameters(t, f("cutoff"), f("q")), mixer)?;
            }
            "eq5" => {
                let h = player.add_effect(Eq5Effect::new(), mixer)?;
                for (i, g) in e["gains"].as_array().unwrap().iter().enumerate() {
                    let id = [b'g', b'a', b'n', b'1' + i as u8];
                    h.set_parameter((phonic::four_cc::FourCC(id), phonic::ParameterValueUpdate::Raw(Arc::new(g.as_f64().unwrap() as f32))), 0u64)?;
                }
            }
            "compressor" => {
                player.add_effect(
                    CompressorEffect::with_compressor_parameters(f("threshold"), f("ratio"), f("knee"), f("attack"), f("release"), f("makeup"), f("lookahead")),
                    mixer,
                )?;
            }
            "chorus" => {
                let ft = match e["filter_type"].as_u64().unwrap() {
                    0 => ChorusEffectFilterType::Lowpass,
                    1 => ChorusEffectFilterType::Highpass,
                    _ => ChorusEffectFilterType::Bandpass,
                };
                player.add_effect(
                    ChorusEffect::with_parameters(f("rate"), f("phase"), f("depth"), f("feedback"), f("delay"), f("wet"), ft, f("filter_freq"), f("filter_resonance")),
                    mixer,
                )?;
            }
            "delay" => {
                player.add_effect(DelayEffect::new(), mixer)?;
            }
            other => panic!("unknown effect kind {other}"),
        }
    }
    let _ = FloatParameter::new; // (keeps the parameters import used on every phonic version)
    Ok(())
}

fn add_source(player: &mut Player, s: &Value, buffers: &HashMap<String, Arc<AudioFileBuffer>>, mixer: Option<MixerId>) -> Result<(), Error> {
    let sr = player.output_sample_rate();
    let buffer = buffers[s["buffer"].as_str().unwrap()].clone();
    if s["type"] == "file" {
        let o = &s["options"];
        let mut fo = FilePlaybackOptions::default()
            .volume(o["volume"].as_f64().unwrap() as f32)
            .panning(o["panning"].as_f64().unwrap() as f32)
            .speed(o["speed"].as_f64().unwrap());
        if o["repeat"] == "forever" {
            fo = fo.repeat_forever();
        } else if let Some(n) = o["repeat"].as_u64() {
            fo = fo.repeat(n as usize);
        }
        if let Some(l) = o["loop_range"].as_array() {
            fo = fo.loop_range(l[0].as_u64().unwrap()..l[1].as_u64().unwrap());
        }
        if let Some(d) = secs(&o["fade_in"]) {
            fo = fo.fade_in(d);
        }
        fo.fade_out_duration = secs(&o["fade_out"]);
        if o["hq"].as_bool().unwrap() {
            fo = fo.resampling_quality(ResamplingQuality::HighQuality);
        }
        if let Some(m) = mixer {
            fo = fo.target_mixer(m);
        }
        let source = PreloadedFileSource::from_shared_buffer(buffer, "fixture", fo, sr)?;
        let h = player.play_file_source(source, s["start"].as_u64())?;
        for e in s["events"].as_array().unwrap() {
            let t = e["t"].as_u64().unwrap();
            match e["kind"].as_str().unwrap() {
                "set_volume" => h.set_volume(e["value"].as_f64().unwrap() as f32, t)?,
                "set_panning" => h.set_panning(e["value"].as_f64().unwrap() as f32, t)?,
                "set_speed" => h.set_speed(e["speed"].as_f64().unwrap(), e["glide"].as_f64().map(|g| g as f32), t)?,
                "seek" => h.seek(Duration::from_secs_f64(e["seconds"].as_f64().unwrap()), t)?,
                "stop" => h.stop(t)?,
                other => panic!("unknown file event {other}"),
            }
        }
        return Ok(());
    }
    let mut go = GeneratorPlaybackOptions::default();
    go.volume = s["volume"].as_f64().unwrap() as f32;
    go.panning = s["panning"].as_f64().unwrap() as f32;
    go.voices = s["voices"].as_u64().unwrap() as usize;
    go.target_mixer = mixer;
    let file = PreloadedFileSource::from_shared_buffer(buffer, "fixture", FilePlaybackOptions::default(), sr)?;
    let mut sampler = Sampler::from_file_source(file, go, player.output_channel_count(), sr)?;
    if let Some(a) = s["ahdsr"].as_object() {
        let d = |k: &str| Duration::from_secs_f64(a[k].as_f64().unwrap());
        sampler = sampler.with_ahdsr(AhdsrParameters::new(d("attack"), d("hold"), d("decay"), a["sustain"].as_f64().unwrap() as f32, d("release"))?)?;
    }
    if let Some(g) = s["granular"].as_object() {
        let mut gp = GranularParameters::default();
        gp.overlap_mode = if g["overlap_mode"].as_u64().unwrap() == 0 { GrainOverlapMode::Cloud } else { GrainOverlapMode::Sequential };
        gp.window = match g["window"].as_u64().unwrap() {
            0 => GrainWindowMode::Hann, 1 => GrainWindowMode::Blackman, 2 => GrainWindowMode::Triangle, 3 => GrainWindowMode::Tukey,
            4 => GrainWindowMode::Trapezoid, 5 => GrainWindowMode::Exponential, 6 => GrainWindowMode::RampUp, _ => GrainWindowMode::RampDown,
        };
        gp.size = g["size"].as_f64().unwrap() as f32;
        gp.density = g["density"].as_f64().unwrap() as f32;
        gp.position = g["position"].as_f64().unwrap() as f32;
        gp.step = g["step"].as_f64().unwrap() as f32;
        gp.variation = 0.0; gp.spray = 0.0; gp.pan_spread = 0.0;   // every RNG draw is multiplied by zero
        gp.playback_direction = if g["playback_direction"].as_u64().unwrap() == 1 { GrainPlaybackDirection::Backward } else { GrainPlaybackDirection::Forward };
        sampler = sampler.with_granular_playback(gp)?;
    }
    let h = player.add_generator(sampler, mixer)?;
    let mut ids = HashMap::new();
    for e in s["events"].as_array().unwrap() {
        let t = e["t"].as_u64().unwrap();
        let note = |e: &Value| ids[&e["ref"].as_u64().unwrap()];
        match e["kind"].as_str().unwrap() {
            "note_on" => {
                let id = h.note_on(e["note"].as_u64().unwrap() as u8, e["volume"].as_f64().map(|v| v as f32), e["panning"].as_f64().map(|v| v as f32), t)?;
                ids.insert(e["id"].as_u64().unwrap(), id);
            }
            "note_off" => h.note_off(note(e), t)?,
            "all_notes_off" => h.all_notes_off(t)?,
            "set_note_speed" => h.set_note_speed(note(e), e["speed"].as_f64().unwrap(), e["glide"].as_f64().map(|g| g as f32), t)?,
            "set_note_volume" => h.set_note_volume(note(e), e["value"].as_f64().unwrap() as f32, t)?,
            "set_note_panning" => h.set_note_panning(note(e), e["value"].as_f64().unwrap() as f32, t)?,
            other => panic!("unknown generator event {other}"),
        }
    }
    Ok(())
}

#[test]
fn dump_reference_scenes() -> Result<(), Error> {
    let dir = repo().join("tests/golden/reference");
    let manifest: Value = serde_json::from_str(&fs::read_to_string(dir.join("scenes.json")).expect("run `python tests/reference_scenes.py --write` first")).unwrap();
    let sr = manifest["sample_rate"].as_u64().unwrap() as u32;
    let buffers: HashMap<String, Arc<AudioFileBuffer>> =
        manifest["buffers"].as_object().unwrap().iter().map(|(k, v)| (k.clone(), load_buffer(k, v))).collect();
    fs::create_dir_all(dir.join("out")).unwrap();
    for scene in manifest["scenes"].as_array().unwrap() {
        let name = scene["name"].as_str().unwrap();
        let frames = scene["frames"].as_u64().unwrap();
        // WavStream renders whole 1024-frame blocks while whole-seconds(pos / sr) < duration (wav.rs:222): ask for enough
        // whole seconds, the comparison uses the first `frames` frames
        let duration = Duration::from_secs(frames / sr as u64 + 1);
        let path = dir.join("out").join(format!("{name}.wav"));
        {
            let mut player = Player::new(WavOutput::open_with_specs(&path, sr, 2, duration)?, None);
            player.stop(); // hold the output until the whole graph and score are queued (sample times count from 0)
            if let Some(mixers) = scene.get("mixers").and_then(|m| m.as_array()) {
                for m in mixers {
                    let id = player.add_mixer(None)?.id();
                    for s in m["sources"].as_array().unwrap() {
                        add_source(&mut player, s, &buffers, Some(id))?;
                    }
                    add_effects(&mut player, &m["effects"], Some(id))?;
                }
            }
            for s in scene["sources"].as_array().unwrap() {
                add_source(&mut player, s, &buffers, None)?;
            }
            add_effects(&mut player, &scene["effects"], None)?;
            player.start();
            // the WAV writer thread runs ahead of real time (1 ms sleep per block); wait for it to pass the duration
            let total = duration.as_secs() * sr as u64;
            let mut last = u64::MAX;
            loop {
                std::thread::sleep(Duration::from_millis(50));
                let pos = player.output_sample_frame_position();
                if pos >= total || pos == last {
                    break;
                }
                last = pos;
            }
        } // dropping the player finalises the WAV file
        println!("wrote {}", path.display());
    }
    Ok(())
}
