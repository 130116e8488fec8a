//! dump_golden — writes REFERENCE vectors for libgooey_b200's parity tests.
//!
//! This file is meant to be dropped into a libgooey checkout as `examples/dump_golden.rs` and run there:
//!
//!     cargo run --release --no-default-features --features bounce --example dump_golden -- \
//!         <libgooey_b200>/tests/golden/ref/scripts <libgooey_b200>/tests/golden/ref
//!
//! It uses nothing but libgooey's public API (`gooey::instruments`, `gooey::engine`, `gooey::bounce`, `gooey::ffi`,
//! `gooey::utils::oversampler`) plus the two third-party pieces the bounce path depends on (`halfband::iir`, std
//! `DefaultHasher`).  Every output is raw little-endian (`.f32le` / `.u64le`); the tests in libgooey_b200
//! (`tests/test_ref_golden_cpu.py`, `tests/test_ref_golden_gpu.py`) pick the files up when present and otherwise
//! report the oracle as unpinned.
//!
//! Outputs
//!   hasher.u64le            DefaultHasher::new(); (n as u64).hash(); finish()   for n = 0..32   (gen/oscillator.rs:187-196)
//!   halfband_up8.f32le      Upsampler8::default().process(x) for x = impulse of 64 samples  -> 128 values
//!   halfband_down8.f32le    Downsampler8::default().process(a, b) for ((1,0),(0,0)...) then a fresh one for ((0,1),(0,0)...): 64 + 64
//!   oversampler.f32le       Oversampler2x / Oversampler4x::process(x, tanh(3x)) over 256 samples of a 1 kHz sine: 256 + 256
//!   preset_kit.f32le        every preset voice of tests/golden_cases.py::kit_patches, 4096 frames each, trigger at t = 0
//!   c1_kick.f32le           BASELINE config C1: Engine + KickDrum::new + pattern [1,0,...], bounce_to_buffer(Samples(44100))
//!   engine_<name>.f32le     for every <name>.calls in the scripts directory: the FFI calls replayed on gooey_engine_new(44100),
//!                           then gooey_engine_bounce_to_buffer(bars) (bars = the script's `bounce` line); sample_playback.calls
//!                           covers the loop mixer (direct / Resample / WSOLA, a queued take) and a sampler rack with its pattern
//!   sweep64.f32le           for sweep64.voices: 64 voices built with `<Voice>Config::new_full` / Tom2::set_config, 8192 frames
use std::collections::hash_map::DefaultHasher;
use std::fs;
use std::hash::{Hash, Hasher};
use std::io::Write;
use std::path::{Path, PathBuf};

use gooey::bounce::{bounce_to_buffer, BounceLength};
use gooey::engine::{Engine, Instrument, Sequencer};
use gooey::ffi;
use gooey::instruments::{
    FilterSlope, HiHat2, HiHat2Config, KickConfig, KickDrum, NoiseColor, SnareConfig, SnareDrum, Tom2, Tom2Config,
};
use gooey::utils::{Oversampler2x, Oversampler4x};
use halfband::iir::{Downsampler8, Upsampler8};

const SR: f32 = 44100.0;

fn write_f32(path: &Path, data: &[f32]) {
    let mut f = fs::File::create(path).expect("create output");
    for x in data {
        f.write_all(&x.to_le_bytes()).unwrap();
    }
    println!("wrote {} ({} values)", path.display(), data.len());
}

/// `for n in 0..frames { out.push(v.tick(t)); t += 1.0 / sr }` — the bounce loop of src/bounce.rs:48-53 on one voice.
fn render_voice(v: &mut dyn Instrument, velocity: f32, frames: usize) -> Vec<f32> {
    let mut out = Vec::with_capacity(frames);
    let mut t = 0.0f64;
    let dt = 1.0 / SR as f64;
    v.trigger_with_velocity(0.0, velocity);
    for _ in 0..frames {
        out.push(v.tick(t));
        t += dt;
    }
    out
}

fn tom_with(cfg: Tom2Config) -> Tom2 {
    let mut t = Tom2::new(SR);
    t.set_config(cfg);
    t
}

fn preset_kit() -> Vec<f32> {
    // order and velocities of tests/golden_cases.py::kit_patches: linspace(0.4, 1.0, 17)
    let mut voices: Vec<Box<dyn Instrument>> = vec![
        Box::new(KickDrum::with_config(SR, KickConfig::tight())),
        Box::new(KickDrum::with_config(SR, KickConfig::punch())),
        Box::new(KickDrum::with_config(SR, KickConfig::loose())),
        Box::new(KickDrum::with_config(SR, KickConfig::dirt())),
        Box::new(SnareDrum::with_config(SR, SnareConfig::tight())),
        Box::new(SnareDrum::with_config(SR, SnareConfig::loose())),
        Box::new(SnareDrum::with_config(SR, SnareConfig::hiss())),
        Box::new(SnareDrum::with_config(SR, SnareConfig::smack())),
        Box::new(HiHat2::with_config(SR, HiHat2Config::short())),
        Box::new(HiHat2::with_config(SR, HiHat2Config::loose())),
        Box::new(HiHat2::with_config(SR, HiHat2Config::dark())),
        Box::new(HiHat2::with_config(SR, HiHat2Config::soft())),
        Box::new(tom_with(Tom2Config::derp())),
        Box::new(tom_with(Tom2Config::ring())),
        Box::new(tom_with(Tom2Config::brush())),
        Box::new(tom_with(Tom2Config::void_preset())),
        Box::new(Tom2::new(SR)),
    ];
    let n = voices.len();
    let mut out = Vec::new();
    for (i, v) in voices.iter_mut().enumerate() {
        // numpy linspace(0.4, 1.0, n) evaluated in f64, then cast to f32
        let vel = (0.4f64 + (1.0f64 - 0.4f64) * (i as f64) / ((n - 1) as f64)) as f32;
        out.extend(render_voice(v.as_mut(), vel, 4096));
    }
    out
}

/// sweep64.voices: one line per voice: `<instrument> <aux> <velocity> <p0> ... <p23>` (tests/workloads.py drum sweep).
fn sweep(path: &Path, frames: usize) -> Vec<f32> {
    let text = fs::read_to_string(path).expect("sweep64.voices");
    let mut out = Vec::new();
    for line in text.lines() {
        let w: Vec<&str> = line.split_whitespace().collect();
        if w.is_empty() || w[0].starts_with('#') {
            continue;
        }
        let inst: u32 = w[0].parse().unwrap();
        let vel: f32 = w[2].parse().unwrap();
        let p: Vec<f32> = w[3..].iter().map(|s| s.parse().unwrap()).collect();
        let mut v: Box<dyn Instrument> = match inst {
            0 => Box::new(KickDrum::with_config(
                SR,
                KickConfig::new_full(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12], p[13], p[14], p[15], p[16], p[17]),
            )),
            1 => Box::new(SnareDrum::with_config(
                SR,
                SnareConfig::new_full(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12], p[13] as u8, p[14], p[15], p[16], p[17], p[18]),
            )),
            2 => {
                let mut c = HiHat2Config::new(p[0], p[1], p[2], NoiseColor::White, FilterSlope::Db12, p[3]);
                c.volume = p[4];
                Box::new(HiHat2::with_config(SR, c))
            }
            _ => Box::new(tom_with(Tom2Config { tune: p[0], bend: p[1], tone: p[2], color: p[3], decay: p[4], membrane: p[5], membrane_q: p[6], volume: p[7] })),
        };
        out.extend(render_voice(v.as_mut(), vel, frames));
    }
    out
}

/// Replays one `<name>.calls` script (written by tests/golden/make_ref_scripts.py) through the C FFI.
/// Interleaved synthetic PCM by formula — `Engine.synth_pcm` of libgooey_b200/engine.py: exact in f32 (an integer in -1000..=1000 over 1024).
fn synth_pcm(frames: usize, channels: usize, seed: u64) -> Vec<f32> {
    let mut out = Vec::with_capacity(frames * channels);
    for k in 0..frames as u64 {
        for ch in 0..channels as u64 {
            let v = ((k * 37 + ch * 101 + seed * 977) * (k % 89 + 3)) % 2001;
            out.push((v as i64 - 1000) as f32 / 1024.0);
        }
    }
    out
}

fn replay(path: &Path) -> Vec<f32> {
    let text = fs::read_to_string(path).expect("script");
    let e = ffi::gooey_engine_new(SR);
    let mut bars = 1u32;
    unsafe {
        for line in text.lines() {
            let w: Vec<&str> = line.split_whitespace().collect();
            if w.is_empty() || w[0].starts_with('#') {
                continue;
            }
            let u = |i: usize| -> u32 { w[i].parse().unwrap() };
            let f = |i: usize| -> f32 { w[i].parse().unwrap() };
            let b = |i: usize| -> bool { w[i] == "1" };
            match w[0] {
                "bounce" => bars = u(1),
                "set_bpm" => ffi::gooey_engine_set_bpm(e, f(1)),
                "set_swing" => ffi::gooey_engine_set_swing(e, f(1)),
                "set_master_gain" => ffi::gooey_engine_set_master_gain(e, f(1)),
                "set_kick_param" => ffi::gooey_engine_set_kick_param(e, u(1), f(2)),
                "set_snare_param" => ffi::gooey_engine_set_snare_param(e, u(1), f(2)),
                "set_hihat_param" => ffi::gooey_engine_set_hihat_param(e, u(1), f(2)),
                "set_tom_param" => ffi::gooey_engine_set_tom_param(e, u(1), f(2)),
                "set_bass_param" => ffi::gooey_engine_set_bass_param(e, u(1), f(2)),
                "sequencer_set_instrument_step" => ffi::gooey_engine_sequencer_set_instrument_step(e, u(1), u(2), b(3)),
                "sequencer_set_instrument_step_settings" => ffi::gooey_engine_sequencer_set_instrument_step_settings(
                    e, u(1), u(2), b(3), b(4), f(5), b(6), f(7), f(8), b(9), u(10) as u8,
                ),
                "mixer_add_track" => {
                    let name = std::ffi::CString::new(w[1]).unwrap();
                    ffi::gooey_engine_mixer_add_track(e, name.as_ptr());
                }
                "mixer_route_source" => {
                    ffi::gooey_engine_mixer_route_source(e, u(1), u(2));
                }
                "mixer_set_track_gain" => ffi::gooey_engine_mixer_set_track_gain(e, u(1), f(2)),
                "mixer_set_track_pan" => ffi::gooey_engine_mixer_set_track_pan(e, u(1), f(2)),
                "set_global_effect_param" => ffi::gooey_engine_set_global_effect_param(e, u(1), u(2), f(3)),
                "set_global_effect_enabled" => ffi::gooey_engine_set_global_effect_enabled(e, u(1), b(2)),
                "sequencer_start" => ffi::gooey_engine_sequencer_start(e),
                // sample-playback sources; the `_synth` forms load synth_pcm buffers (below) so the scripts carry no audio
                "loop_load_synth" => {
                    let pcm = synth_pcm(u(2) as usize, u(3) as usize, u(5) as u64);
                    assert!(ffi::gooey_engine_loop_load(e, u(1), pcm.as_ptr(), u(2), u(3), f(4)));
                }
                "loop_queue_swap_synth" => {
                    let pcm = synth_pcm(u(2) as usize, u(3) as usize, u(5) as u64);
                    assert!(ffi::gooey_engine_loop_queue_swap(e, u(1), pcm.as_ptr(), u(2), u(3), f(4), f(6), u(7)));
                }
                "loop_set_playing" => ffi::gooey_engine_loop_set_playing(e, u(1), b(2)),
                "loop_set_gain" => ffi::gooey_engine_loop_set_gain(e, u(1), f(2)),
                "loop_set_mute" => ffi::gooey_engine_loop_set_mute(e, u(1), b(2)),
                "loop_set_solo" => ffi::gooey_engine_loop_set_solo(e, u(1), b(2)),
                "loop_set_start" => ffi::gooey_engine_loop_set_start(e, u(1), f(2)),
                "loop_set_end" => ffi::gooey_engine_loop_set_end(e, u(1), f(2)),
                "loop_set_speed" => ffi::gooey_engine_loop_set_speed(e, u(1), f(2)),
                "loop_set_source_bpm" => ffi::gooey_engine_loop_set_source_bpm(e, u(1), f(2)),
                "loop_set_pitch_mode" => ffi::gooey_engine_loop_set_pitch_mode(e, u(1), u(2)),
                "loop_restart" => ffi::gooey_engine_loop_restart(e, u(1)),
                "loop_set_position" => ffi::gooey_engine_loop_set_position(e, u(1), f(2)),
                "sampler_register" => {
                    assert!(ffi::gooey_engine_sampler_register(e) >= 0);
                }
                "sampler_set_slot_synth" => {
                    let pcm = synth_pcm(u(3) as usize, u(4) as usize, u(6) as u64);
                    assert!(ffi::gooey_engine_sampler_set_slot_buffer(e, u(1), u(2), pcm.as_ptr(), u(3), u(4), f(5)));
                }
                "sampler_set_step" => {
                    ffi::gooey_engine_sampler_set_step(e, u(1), u(2), b(3), u(4), f(5));
                }
                "sampler_trigger" => {
                    ffi::gooey_engine_sampler_trigger(e, u(1), u(2), f(3));
                }
                "sampler_start_pattern" => {
                    ffi::gooey_engine_sampler_start_pattern(e, u(1), u(2));
                }
                other => panic!("dump_golden: unknown call `{other}` in {}", path.display()),
            }
        }
        let mut len = 0u32;
        let buf = ffi::gooey_engine_bounce_to_buffer(e, bars, &mut len);
        assert!(!buf.is_null());
        let out = std::slice::from_raw_parts(buf, len as usize).to_vec();
        ffi::gooey_engine_free_buffer(buf, len);
        ffi::gooey_engine_free(e);
        out
    }
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    if args.len() != 3 {
        eprintln!("usage: dump_golden <scripts dir> <output dir>");
        std::process::exit(2);
    }
    let scripts = PathBuf::from(&args[1]);
    let outdir = PathBuf::from(&args[2]);
    fs::create_dir_all(&outdir).unwrap();

    // 1. std DefaultHasher (SipHash-1-3, zero keys) of u64 values
    {
        let mut f = fs::File::create(outdir.join("hasher.u64le")).unwrap();
        for n in 0u64..32 {
            let mut h = DefaultHasher::new();
            n.hash(&mut h);
            f.write_all(&h.finish().to_le_bytes()).unwrap();
        }
    }
    // 2. halfband 0.2 impulse responses
    {
        let mut up = Upsampler8::default();
        let mut v = Vec::new();
        for n in 0..64 {
            let y = up.process(if n == 0 { 1.0 } else { 0.0 });
            v.push(y[0]);
            v.push(y[1]);
        }
        write_f32(&outdir.join("halfband_up8.f32le"), &v);
        let mut v = Vec::new();
        let mut d = Downsampler8::default();
        for n in 0..64 {
            v.push(d.process(if n == 0 { 1.0 } else { 0.0 }, 0.0));
        }
        let mut d = Downsampler8::default();
        for n in 0..64 {
            v.push(d.process(0.0, if n == 0 { 1.0 } else { 0.0 }));
        }
        write_f32(&outdir.join("halfband_down8.f32le"), &v);
    }
    // 3. Oversampler2x / 4x around tanh(3x)
    {
        let mut o2 = Oversampler2x::new();
        let mut o4 = Oversampler4x::new();
        let mut v = Vec::new();
        let sig: Vec<f32> = (0..256).map(|n| 0.8 * (2.0 * std::f32::consts::PI * 1000.0 * n as f32 / SR).sin()).collect();
        for x in &sig {
            v.push(o2.process(*x, |s| (s * 3.0).tanh()));
        }
        for x in &sig {
            v.push(o4.process(*x, |s| (s * 3.0).tanh()));
        }
        write_f32(&outdir.join("oversampler.f32le"), &v);
    }
    // 4. preset kit
    write_f32(&outdir.join("preset_kit.f32le"), &preset_kit());
    // 5. C1
    {
        let mut engine = Engine::new(SR);
        engine.set_bpm(120.0);
        engine.add_instrument("kick", Box::new(KickDrum::new(SR)));
        let pattern: Vec<bool> = (0..16).map(|i| i == 0).collect();
        engine.add_sequencer(Sequencer::with_pattern(120.0, SR, pattern, "kick"));
        write_f32(&outdir.join("c1_kick.f32le"), &bounce_to_buffer(&mut engine, BounceLength::Samples(44100)));
    }
    // 6. FFI engine scripts, 7. the voice sweep
    let mut entries: Vec<PathBuf> = fs::read_dir(&scripts).unwrap().map(|e| e.unwrap().path()).collect();
    entries.sort();
    for p in entries {
        let stem = p.file_stem().unwrap().to_string_lossy().to_string();
        match p.extension().and_then(|s| s.to_str()) {
            Some("calls") => write_f32(&outdir.join(format!("engine_{stem}.f32le")), &replay(&p)),
            Some("voices") => write_f32(&outdir.join(format!("{stem}.f32le")), &sweep(&p, 8192)),
            _ => {}
        }
    }
}
