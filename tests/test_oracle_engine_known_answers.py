"""The reference's WurliEngine unit tests (crates/openwurli-dsp/src/engine.rs:682-1178) restated against the CPU oracle's engine
restatement (oracle/ow_engine.hpp), one test per reference test with its line number.  The GPU engine path is compared with the oracle
sample by sample (tests/test_gpu_parity.py); these known answers pin the oracle's control logic -- allocation priority, stealing,
sustain pedal, re-strike, clamping, smoothers -- to the reference's own assertions, so a transliteration slip in the state machine
cannot hide behind a GPU == oracle comparison."""
import numpy as np

import oracle_lib as O
from oracle_lib import (NOTE_ON, NOTE_OFF, SUSTAIN, RENDER, QUERY, SET_VOLUME, SET_DEPTH, SET_CHARACTER, RESET, SET_SAMPLE_RATE, SET_MLP)

MAX_VOICES = 64


def run(ops, **kw):
    return O.engine_script(ops, **kw)


def test_note_on_allocates_and_note_off_releases():  # engine.rs:697-709
    _, q, _ = run([(NOTE_ON, 60, 0.8), (QUERY, 60, 0), (NOTE_OFF, 60, 0), (QUERY, 60, 0)])
    assert q[0]["held"] == 1 and q[1]["held"] == 0 and q[1]["releasing"] == 1


def test_polyphony_up_to_max_voices_and_stealing_when_full():  # engine.rs:711-730
    ops = [(NOTE_ON, 36 + n, 0.8) for n in range(MAX_VOICES)] + [(QUERY, 0, 0), (NOTE_ON, 96, 0.8), (QUERY, 96, 0)]
    _, q, _ = run(ops)
    assert q[0]["held"] == MAX_VOICES
    assert q[1]["held"] == MAX_VOICES and q[1]["has_steal"] == 1


def test_render_produces_output_and_idle_is_near_silent():  # engine.rs:732-753
    y, _, _ = run([(NOTE_ON, 60, 0.8), (RENDER, 256, 1)])
    assert float((y.astype(np.float64) ** 2).sum()) > 0.0
    z, _, _ = run([(RENDER, 512, 1)])
    assert float(np.abs(z).max()) < 0.05


def test_reset_clears_voices_and_sustain_state():  # engine.rs:755-762, 1013-1021
    _, q, _ = run([(NOTE_ON, 60, 0.8), (NOTE_ON, 72, 0.8), (RESET, 0, 0), (QUERY, 0, 0)])
    assert q[0]["active"] == 0
    _, q, _ = run([(SUSTAIN, 1, 0), (NOTE_ON, 60, 0.8), (NOTE_OFF, 60, 0), (RESET, 0, 0), (QUERY, 0, 0)])
    assert q[0]["sustain_flag"] == 0 and q[0]["active"] == 0


def test_sustain_defers_note_off_and_pedal_release_triggers_damping():  # engine.rs:764-776, 923-936
    _, q, _ = run([(SUSTAIN, 1, 0), (NOTE_ON, 60, 0.8), (NOTE_OFF, 60, 0), (QUERY, 60, 0), (SUSTAIN, 0, 0), (QUERY, 60, 0)])
    assert q[0]["sustained"] == 1 and q[0]["held"] == 0
    assert q[1]["sustained"] == 0 and q[1]["releasing"] == 1


def test_volume_smoother_ramps():  # engine.rs:778-787: default 0.5 ramping to 1.0 over ~220 samples
    _, _, sm = run([(SET_VOLUME, 1.0, 0), (RENDER, 1, 0)])
    assert 0.5 < sm[0] < 1.0
    _, _, sm = run([(SET_VOLUME, 1.0, 0), (RENDER, 220, 0)])
    assert sm[0] == 1.0  # snaps exactly on the last ramp step (engine.rs:112-121)


def _chord_peak(volume, depth, notes, vel, seconds):
    ops = [(SET_VOLUME, volume, 0), (SET_DEPTH, depth, 0), (SET_CHARACTER, 0.0, 0), (SET_MLP, 1, 0)] + [(RENDER, 1024, 0)] * 6
    ops += [(NOTE_ON, n, vel) for n in notes]
    total = int(44100.0 * seconds)
    ops += [(RENDER, 1024, 1)] * (total // 1024) + ([(RENDER, total % 1024, 1)] if total % 1024 else [])
    y, _, _ = run(ops)
    return float(np.abs(y).max())


def test_engine_peak_below_unity_at_vol_1():  # engine.rs:788-836: chord-ff, vol 1.0, tremolo bright -> peak <= 1.02
    assert _chord_peak(1.0, 1.0, [48, 55, 60, 63, 67, 70], 0.95, 1.0) <= 1.02


def test_user_volume_scales_output_linearly():  # engine.rs:838-882: ratio 2.0 +- 2 %
    p05 = _chord_peak(0.5, 0.0, [60], 0.95, 0.5)
    p10 = _chord_peak(1.0, 0.0, [60], 0.95, 0.5)
    assert 1.96 <= p10 / p05 <= 2.04, (p05, p10)


def test_higher_velocity_louder():  # engine.rs:884-906
    soft, _, _ = run([(SET_VOLUME, 0.5, 0), (NOTE_ON, 60, 0.2), (RENDER, 4096, 1)])
    loud, _, _ = run([(SET_VOLUME, 0.5, 0), (NOTE_ON, 60, 0.2), (RENDER, 4096, 0), (RESET, 0, 0), (NOTE_ON, 60, 1.0), (RENDER, 4096, 1)])
    rms = lambda x: float(np.sqrt((x.astype(np.float64) ** 2).mean()))
    assert rms(loud) > rms(soft)


def test_note_clamps_to_valid_range():  # engine.rs:910-919
    _, q, _ = run([(NOTE_ON, 0, 0.8), (NOTE_ON, 127, 0.8), (QUERY, 0, 0)])
    assert q[0]["held"] == 2


def test_sustain_held_voices_still_render():  # engine.rs:938-953
    y, _, _ = run([(SUSTAIN, 1, 0), (NOTE_ON, 60, 0.8), (RENDER, 1024, 0), (NOTE_OFF, 60, 0), (RENDER, 1024, 0), (SUSTAIN, 0, 0), (RENDER, 1024, 1)])
    assert float((y.astype(np.float64) ** 2).sum()) > 0.0


def test_voice_stealing_prefers_sustained_over_held():  # engine.rs:966-987
    ops = [(SUSTAIN, 1, 0)]
    for n in range(MAX_VOICES // 2):
        ops += [(NOTE_ON, 36 + n, 0.8), (NOTE_OFF, 36 + n, 0)]
    for n in range(MAX_VOICES // 2, MAX_VOICES):
        ops += [(NOTE_ON, 36 + n, 0.8)]
    ops += [(QUERY, 0, 0), (NOTE_ON, 127, 0.8), (QUERY, 0, 0)]
    _, q, _ = run(ops)
    assert q[0]["sustained"] + q[0]["held"] == MAX_VOICES
    assert q[1]["held"] == q[0]["held"] + 1 and q[1]["sustained"] == q[0]["sustained"] - 1


def test_reattack_releases_sustained_same_note():  # engine.rs:989-1007
    _, q, _ = run([(SUSTAIN, 1, 0), (NOTE_ON, 60, 0.8), (NOTE_OFF, 60, 0), (NOTE_ON, 60, 0.8), (QUERY, 60, 0)])
    assert q[0]["note_sustained"] == 0 and q[0]["note_held"] == 1


def test_pedal_up_only_releases_sustained_not_held():  # engine.rs:1009-1024
    _, q, _ = run([(SUSTAIN, 1, 0), (NOTE_ON, 60, 0.8), (NOTE_OFF, 60, 0), (NOTE_ON, 64, 0.8), (QUERY, 0, 0), (SUSTAIN, 0, 0), (QUERY, 0, 0)])
    assert q[0]["sustained"] == 1 and q[0]["held"] == 1
    assert q[1]["sustained"] == 0 and q[1]["held"] == 1


def test_note_off_for_nonexistent_note_is_noop():  # engine.rs:1037-1047
    _, q, _ = run([(NOTE_ON, 60, 0.8), (NOTE_OFF, 72, 0), (QUERY, 0, 0)])
    assert q[0]["held"] == 1


def test_volume_zero_and_back_no_nan():  # engine.rs:1051-1068
    ops = [(NOTE_ON, 60, 0.8)]
    for _ in range(4):
        ops += [(SET_VOLUME, 0.0, 0), (RENDER, 512, 1), (SET_VOLUME, 0.5, 0), (RENDER, 512, 1)]
    y, _, _ = run(ops)
    assert np.all(np.isfinite(y))


def test_no_catastrophic_output_spikes_under_continuous_play():  # engine.rs:1070-1110: peak < +14 dBFS
    chords = [[60, 64, 67], [62, 65, 69], [64, 67, 71], [65, 69, 72]]
    ops = []
    for i in range(8):
        ch = chords[i % 4]
        ops += [(NOTE_ON, n, 1.0) for n in ch] + [(RENDER, 256, 1)] * 86 + [(NOTE_OFF, n, 0) for n in ch]
        if i % 2 == 1:
            ops += [(RENDER, 256, 0)] * 5
    y, _, _ = run(ops)
    assert 20.0 * np.log10(max(float(np.abs(y).max()), 1e-12)) < 14.0


def test_sound_after_sample_rate_change_and_big_buffers():  # engine.rs:1114-1133
    y, _, _ = run([(SET_SAMPLE_RATE, 48000.0, 0), (NOTE_ON, 60, 0.8), (RENDER, 1024, 1)])
    assert float((y.astype(np.float64) ** 2).sum()) > 0.0
    y, _, _ = run([(NOTE_ON, 60, 0.8), (RENDER, 16384, 1)])
    assert np.all(np.isfinite(y))


def test_tremolo_smoother_does_not_pin_depth_to_zero():  # engine.rs:1138-1178: > 3 dB RMS swing at the default depth 0.5
    sr = 44100
    y, _, _ = run([(NOTE_ON, 60, 0.9)] + [(RENDER, 256, 1)] * (sr * 4 // 256))
    win, skip = sr // 50, 25
    env = []
    for i in range(skip, len(y) // win):
        seg = y[i * win:(i + 1) * win].astype(np.float64)
        env.append(20.0 * np.log10(np.sqrt((seg ** 2).mean()) + 1e-12))
    assert max(env) - min(env) > 3.0
