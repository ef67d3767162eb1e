// TEST INFRASTRUCTURE ONLY -- CPU oracle (parity checker), never on the product path.
//
// Restatement of WurliEngine ("chain E", crates/openwurli-dsp/src/engine.rs):
// 64 voice slots, stealing with 5 ms linear crossfade, sustain pedal, 5 ms
// parameter smoothers, NaN guards, f32 output.
#pragma once
#include "ow_chain.hpp"
#include <memory>

namespace ow {

struct LinearSmoother {  // engine.rs:67-130
    double current, target, step;
    uint32_t samples_remaining, ramp_samples;
    LinearSmoother(double initial, uint32_t ramp) : current(initial), target(initial), step(0.0), samples_remaining(0), ramp_samples(ramp) {}
    void set_target(double t) {
        if (std::fabs(t - target) < 1e-9) return;
        target = t;
        const double delta = t - current;
        if (ramp_samples == 0) { current = t; samples_remaining = 0; return; }
        step = delta / (double)ramp_samples;
        samples_remaining = ramp_samples;
    }
    void snap_to(double v) { current = v; target = v; step = 0.0; samples_remaining = 0; }
    void set_ramp_samples(uint32_t r) {
        ramp_samples = r;
        if (samples_remaining > 0) {
            step = (target - current) / (double)std::max(r, 1u);
            samples_remaining = r;
        }
    }
    inline double next() {
        if (samples_remaining > 0) {
            current += step;
            samples_remaining -= 1;
            if (samples_remaining == 0) current = target;
        }
        return current;
    }
};

static inline uint32_t ramp_samples_for_rate(double sr) { return std::max(f64_as_u32(sr * 0.005), 1u); }  // engine.rs:672-674

enum class VState : uint8_t { Free, Held, Sustained, Releasing };

struct VoiceSlot {  // engine.rs:36-61
    std::unique_ptr<Voice> voice;
    VState state = VState::Free;
    uint8_t midi_note = 0;
    uint64_t age = 0;
    std::unique_ptr<Voice> steal_voice;
    uint32_t steal_fade = 0, steal_fade_len = 0;
};

struct WurliEngine {
    static constexpr int MAX_VOICES = 64;
    std::vector<VoiceSlot> voices;
    uint64_t age_counter = 0;
    std::unique_ptr<AnyPreamp> preamp;
    int preamp_model = 0;  // 0 melange 12-node, 1 legacy 8-node (the default build)
    std::unique_ptr<Tremolo> tremolo;
    Oversampler oversampler;
    PowerAmp power_amp;
    std::unique_ptr<Speaker> speaker;
    std::vector<double> voice_buf, sum_buf, up_buf, out_buf;
    double sample_rate, os_sample_rate;
    bool oversample;
    bool sustain_held = false, mlp_enabled = true;
    LinearSmoother volume, tremolo_depth, speaker_character;
    uint64_t nan_guard_fires = 0;

    explicit WurliEngine(double sr, int model = 0)  // engine.rs:194-229 (no warm-up)
        : voices(MAX_VOICES), preamp_model(model), sample_rate(sr), volume(0.5, ramp_samples_for_rate(sr)),
          tremolo_depth(0.5, ramp_samples_for_rate(sr)), speaker_character(0.0, ramp_samples_for_rate(sr)) {
        oversample = sr < 88200.0;
        os_sample_rate = oversample ? sr * 2.0 : sr;
        preamp.reset(new AnyPreamp(preamp_model, os_sample_rate));
        tremolo.reset(new Tremolo(0.5, os_sample_rate));
        speaker.reset(new Speaker(sr));
        voice_buf.assign(8192, 0.0); sum_buf.assign(8192, 0.0); up_buf.assign(16384, 0.0); out_buf.assign(8192, 0.0);
    }

    void warm_up() {  // engine.rs:261-270
        float scratch[512];
        const size_t total = (size_t)f64_as_u64(sample_rate * 0.6);
        size_t done = 0;
        while (done < total) {
            const size_t len = std::min<size_t>(512, total - done);
            render(scratch, len);
            done += len;
        }
    }

    void set_sample_rate(double sr) {  // engine.rs:272-286
        sample_rate = sr;
        oversample = sr < 88200.0;
        os_sample_rate = oversample ? sr * 2.0 : sr;
        preamp.reset(new AnyPreamp(preamp_model, os_sample_rate));
        tremolo.reset(new Tremolo(tremolo_depth.target, os_sample_rate));
        oversampler = Oversampler();
        power_amp = PowerAmp();
        speaker.reset(new Speaker(sr));
        const uint32_t ramp = ramp_samples_for_rate(sr);
        volume.set_ramp_samples(ramp);
        tremolo_depth.set_ramp_samples(ramp);
        speaker_character.set_ramp_samples(ramp);
        warm_up();
    }

    void reset() {  // engine.rs:232-251
        for (auto& s : voices) { s.state = VState::Free; s.voice.reset(); s.steal_voice.reset(); s.steal_fade = 0; }
        preamp->reset();
        tremolo->reset();
        oversampler.reset();
        speaker->reset();
        age_counter = 0;
        sustain_held = false;
        volume.snap_to(volume.target);
        tremolo_depth.snap_to(tremolo_depth.target);
        speaker_character.snap_to(speaker_character.target);
        warm_up();
    }

    void ensure_buffer_capacity(size_t n) {
        if (sum_buf.size() < n) { voice_buf.resize(n, 0.0); sum_buf.resize(n, 0.0); up_buf.resize(2 * n, 0.0); out_buf.resize(n, 0.0); }
    }

    int allocate_voice() const {  // engine.rs:569-590
        int best_idx = 0;
        uint64_t best_priority = UINT64_MAX;
        for (int i = 0; i < MAX_VOICES; i++) {
            const VoiceSlot& s = voices[i];
            uint64_t pr;
            switch (s.state) {
                case VState::Free: return i;
                case VState::Releasing: pr = s.age; break;
                case VState::Sustained: pr = s.age + UINT64_MAX / 4; break;
                default: pr = s.age + UINT64_MAX / 2; break;
            }
            if (pr < best_priority) { best_priority = pr; best_idx = i; }
        }
        return best_idx;
    }

    void note_on(uint8_t note, float velocity) {  // engine.rs:299-338
        note = std::min<uint8_t>(std::max<uint8_t>(note, 33), 96);
        for (auto& s : voices) {
            if (s.state == VState::Sustained && s.midi_note == note) {
                s.state = VState::Releasing;
                if (s.voice) s.voice->note_off();
            }
        }
        const int idx = allocate_voice();
        VoiceSlot& slot = voices[idx];
        if (slot.state != VState::Free) {
            const uint32_t fade = f64_as_u32(sample_rate * 0.005);
            slot.steal_voice = std::move(slot.voice);
            slot.steal_fade = fade;
            slot.steal_fade_len = fade;
        }
        age_counter += 1;
        const uint32_t seed = (uint32_t)note * 2654435761u + (uint32_t)age_counter;
        slot.voice.reset(new Voice());
        slot.voice->note_on(note, (double)velocity, sample_rate, seed, mlp_enabled);
        slot.state = VState::Held;
        slot.midi_note = note;
        slot.age = age_counter;
    }

    void note_off(uint8_t note) {  // engine.rs:340-359
        note = std::min<uint8_t>(std::max<uint8_t>(note, 33), 96);
        int oldest = -1;
        for (int i = 0; i < MAX_VOICES; i++) {
            const VoiceSlot& s = voices[i];
            if (s.state == VState::Held && s.midi_note == note) {
                if (oldest < 0 || s.age < voices[oldest].age) oldest = i;  // min_by_key: first minimum
            }
        }
        if (oldest >= 0) {
            if (sustain_held) voices[oldest].state = VState::Sustained;
            else {
                voices[oldest].state = VState::Releasing;
                if (voices[oldest].voice) voices[oldest].voice->note_off();
            }
        }
    }

    void set_sustain(bool held) {  // engine.rs:361-374
        if (sustain_held && !held) {
            for (auto& s : voices) {
                if (s.state == VState::Sustained) {
                    s.state = VState::Releasing;
                    if (s.voice) s.voice->note_off();
                }
            }
        }
        sustain_held = held;
    }

    void render_voices_to_preamp_out(size_t len) {  // engine.rs:466-567 (offset always 0)
        for (size_t i = 0; i < len; i++) sum_buf[i] = 0.0;
        for (auto& slot : voices) {
            if (slot.state == VState::Free && !slot.steal_voice) continue;
            if (slot.voice) {
                slot.voice->render(voice_buf.data(), len);
                for (size_t i = 0; i < len; i++) sum_buf[i] += voice_buf[i];
            }
            if (slot.steal_voice) {
                slot.steal_voice->render(voice_buf.data(), len);
                const double fade_len = (double)slot.steal_fade_len;
                for (size_t i = 0; i < len; i++) {
                    const uint32_t i32 = (uint32_t)i;
                    const uint32_t remaining = slot.steal_fade > i32 ? slot.steal_fade - i32 : 0;
                    const double gain = (double)remaining / fade_len;
                    sum_buf[i] += voice_buf[i] * gain;
                }
                const uint32_t l32 = (uint32_t)len;
                slot.steal_fade = slot.steal_fade > l32 ? slot.steal_fade - l32 : 0;
                if (slot.steal_fade == 0) slot.steal_voice.reset();
            }
        }
        bool bad = false;
        for (size_t i = 0; i < len; i++) if (!std::isfinite(sum_buf[i])) { bad = true; break; }
        if (bad) {  // engine.rs:499-521
            nan_guard_fires++;
            for (size_t i = 0; i < len; i++) sum_buf[i] = 0.0;
            for (auto& slot : voices) {
                if (slot.state == VState::Free && !slot.steal_voice) continue;
                if (slot.voice) {
                    slot.voice->render(voice_buf.data(), len);
                    bool vb = false;
                    for (size_t i = 0; i < len; i++) if (!std::isfinite(voice_buf[i])) vb = true;
                    if (vb) { slot.state = VState::Free; slot.voice.reset(); }
                }
                if (slot.steal_voice) {
                    slot.steal_voice->render(voice_buf.data(), len);
                    bool vb = false;
                    for (size_t i = 0; i < len; i++) if (!std::isfinite(voice_buf[i])) vb = true;
                    if (vb) { slot.steal_voice.reset(); slot.steal_fade = 0; }
                }
            }
        }
        if (oversample) {
            oversampler.upsample_2x(sum_buf.data(), len, up_buf.data());
            for (size_t i = 0; i < len; i++) {
                const double depth = tremolo_depth.next();
                tremolo->set_depth(depth);
                for (int j = 0; j < 2; j++) {
                    const size_t idx = i * 2 + j;
                    const double r = tremolo->process();
                    preamp->set_ldr_resistance(r);
                    const double po = preamp->process_sample(up_buf[idx]);
                    up_buf[idx] = power_amp.process(po * FIXED_CIRCUIT_DRIVE);
                }
            }
            oversampler.downsample_2x(up_buf.data(), out_buf.data(), len);
        } else {
            for (size_t i = 0; i < len; i++) {
                const double depth = tremolo_depth.next();
                tremolo->set_depth(depth);
                const double r = tremolo->process();
                preamp->set_ldr_resistance(r);
                const double po = preamp->process_sample(sum_buf[i]);
                out_buf[i] = power_amp.process(po * FIXED_CIRCUIT_DRIVE);
            }
        }
    }

    void cleanup_voices() {  // engine.rs:592-602
        for (auto& s : voices) {
            if (s.state != VState::Free && s.voice && s.voice->is_silent()) { s.state = VState::Free; s.voice.reset(); }
        }
    }

    void render(float* out, size_t len) {  // engine.rs:425-462
        if (len == 0) return;
        ensure_buffer_capacity(len);
        render_voices_to_preamp_out(len);
        for (size_t i = 0; i < len; i++) {
            const double ch = speaker_character.next();
            speaker->set_character(ch);
            const double shaped = speaker->process(out_buf[i]);
            const double user_vol = volume.next();
            const double post_gain = shaped * POST_SPEAKER_GAIN * user_vol;
            const float sample = (float)post_gain;
            if (std::isfinite(sample)) out[i] = sample;
            else {
                preamp->reset();
                oversampler.reset();
                speaker->reset();
                out[i] = 0.0f;
            }
        }
        cleanup_voices();
    }

    int active_voice_count() const { int c = 0; for (auto& s : voices) if (s.state != VState::Free) c++; return c; }
    int count_state(VState st) const { int c = 0; for (auto& s : voices) if (s.state == st) c++; return c; }
};

}  // namespace ow
