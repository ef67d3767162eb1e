//! Dumps reference renders as raw little-endian buffers for tools/rust_dump/compare.py.
//! Every case mirrors a reference entry point verbatim (file:line in the comments); nothing here is product code.
use std::fs::File;
use std::io::Write;

use openwurli_dsp::dk_preamp::DkPreamp;
use openwurli_dsp::engine::WurliEngine;
use openwurli_dsp::oversampler::Oversampler;
use openwurli_dsp::power_amp::PowerAmp;
use openwurli_dsp::preamp::PreampModel;
use openwurli_dsp::speaker::Speaker;
use openwurli_dsp::tables;
use openwurli_dsp::tremolo::Tremolo;
use openwurli_dsp::voice::Voice;

fn write_f64(path: &str, x: &[f64]) {
    let mut f = File::create(path).unwrap();
    for v in x {
        f.write_all(&v.to_le_bytes()).unwrap();
    }
}

fn write_f32(path: &str, x: &[f32]) {
    let mut f = File::create(path).unwrap();
    for v in x {
        f.write_all(&v.to_le_bytes()).unwrap();
    }
}

/// `preamp-bench render` with default flags (tools/preamp-bench/src/main.rs:371-496).
fn chain_b(note: u8, velocity: u8, duration: f64, sample_rate: f64, r_ldr: f64, tremolo_depth: f64, volume: f64, speaker_char: f64) -> Vec<f64> {
    let do_oversample = sample_rate < 88200.0;
    let preamp_sr = if do_oversample { sample_rate * 2.0 } else { sample_rate };
    let vel_norm = velocity as f64 / 127.0;
    let noise_seed = (note as u32).wrapping_mul(2654435761);
    let mut voice = Voice::note_on(note, vel_norm, sample_rate, noise_seed, true);
    let n = (duration * sample_rate) as usize;
    let mut reed = vec![0.0f64; n];
    let mut off = 0;
    while off < n {
        let end = (off + 1024).min(n);
        voice.render(&mut reed[off..end]);
        off = end;
    }
    let mut preamp = DkPreamp::new(preamp_sr);
    let mut tremolo = if tremolo_depth > 0.0 {
        Some(Tremolo::new(tremolo_depth, preamp_sr))
    } else {
        preamp.reset();
        preamp.set_ldr_resistance(r_ldr);
        None
    };
    let mut pre = vec![0.0f64; n];
    let mut os = Oversampler::new();
    for i in 0..n {
        if do_oversample {
            let mut up = [0.0f64; 2];
            os.upsample_2x(&[reed[i]], &mut up);
            let mut processed = [0.0f64; 2];
            for j in 0..2 {
                if let Some(ref mut t) = tremolo {
                    preamp.set_ldr_resistance(t.process());
                }
                processed[j] = preamp.process_sample(up[j]);
            }
            let mut down = [0.0f64; 1];
            os.downsample_2x(&processed, &mut down);
            pre[i] = down[0];
        } else {
            if let Some(ref mut t) = tremolo {
                preamp.set_ldr_resistance(t.process());
            }
            pre[i] = preamp.process_sample(reed[i]);
        }
    }
    let mut power_amp = PowerAmp::new();
    let mut speaker = Speaker::new(sample_rate);
    speaker.set_character(speaker_char);
    (0..n).map(|i| speaker.process(power_amp.process(pre[i] * volume * volume)) * tables::POST_SPEAKER_GAIN).collect()
}

fn main() {
    let dir = std::env::args().nth(1).unwrap_or_else(|| ".".into());
    std::fs::create_dir_all(&dir).unwrap();
    // chain V: reed-renderer -n 60 -v 100 -d 2.0 (tools/reed-renderer/src/main.rs:83-96)
    write_f64(&format!("{dir}/voice_60_100.f64"), &Voice::render_note(60, 100.0 / 127.0, 2.0, 44100.0));
    write_f64(&format!("{dir}/voice_33_127.f64"), &Voice::render_note(33, 1.0, 1.0, 44100.0));
    write_f64(&format!("{dir}/voice_96_1.f64"), &Voice::render_note(96, 1.0 / 127.0, 1.0, 44100.0));
    // chain B: preamp-bench render, static LDR and tremolo, 2x oversampled and native rate
    write_f64(&format!("{dir}/bench_60_100_static.f64"), &chain_b(60, 100, 2.0, 44100.0, 1e6, 0.0, 0.60, 1.0));
    write_f64(&format!("{dir}/bench_60_100_trem05.f64"), &chain_b(60, 100, 2.0, 44100.0, 1e6, 0.5, 0.60, 1.0));
    write_f64(&format!("{dir}/bench_40_127_trem10_48k.f64"), &chain_b(40, 127, 1.0, 48000.0, 1e6, 1.0, 0.80, 0.4));
    write_f64(&format!("{dir}/bench_84_64_ldr19k_96k.f64"), &chain_b(84, 64, 1.0, 96000.0, 19000.0, 0.0, 0.60, 0.0));
    // chain E: WurliEngine driven like the plugin (set_sample_rate warms up), 512-sample blocks, a re-strike, a pedal and a release
    let sr = 44100.0;
    let mut eng = WurliEngine::new(sr);
    eng.set_sample_rate(sr);
    eng.set_volume(0.5);
    eng.set_tremolo_depth(0.5);
    eng.set_speaker_character(0.5);
    let events: [(usize, u8, u8, f32); 6] = [(0, 0, 60, 0.8), (4096, 0, 64, 0.6), (8192, 2, 1, 0.0), (12288, 1, 60, 0.0), (16384, 0, 60, 0.9), (24576, 2, 0, 0.0)];
    let total = 44100usize;
    let mut out = vec![0.0f32; total];
    let mut e = 0;
    let mut pos = 0;
    while pos < total {
        let len = 512.min(total - pos);
        while e < events.len() && events[e].0 < pos + len {
            let (_, kind, note, vel) = events[e];
            match kind {
                0 => eng.note_on(note, vel),
                1 => eng.note_off(note),
                _ => eng.set_sustain(note != 0),
            }
            e += 1;
        }
        eng.render(&mut out[pos..pos + len]);
        pos += len;
    }
    write_f32(&format!("{dir}/engine_stream.f32"), &out);
    println!("wrote reference dumps to {dir}");
}
