"""WAV output and CLI mirrors: quantisation rules of the two reference tools, RIFF round trip, flag parsing."""
import os
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwurli_b200 import wav
from openwurli_b200.cli import preamp_bench, reed_renderer


def test_pcm24_truncate_matches_rust_cast():
    # reed-renderer main.rs:119-123: (clamp(s,-1,1) * 8388607.0) as i32 -- truncation toward zero
    x = np.array([0.0, 1.0, -1.0, 2.0, -3.0, 0.5, -0.5, 1e-9, -1e-9, 0.9999999, np.nan])
    q = wav.pcm24_truncate(x)
    exp = [0, 8388607, -8388607, 8388607, -8388607, 4194303, -4194303, 0, 0, int(0.9999999 * 8388607.0), 0]
    assert q.tolist() == exp


def test_pcm24_round_half_away_from_zero_and_clamp():
    # preamp-bench main.rs:950-954: (s*scale*max).round() as i32, clamp(+-max)
    m = 8388607.0
    x = np.array([0.5 / m, -0.5 / m, 1.5 / m, -1.5 / m, 2.5 / m, 0.49999 / m, 1.0, -1.0, 1.5, -1.5, 0.0])
    q = wav.pcm24_round(x)
    assert q.tolist() == [1, -1, 2, -2, 3, 0, 8388607, -8388607, 8388607, -8388607, 0]
    # scale is applied before rounding
    assert wav.pcm24_round(np.array([1.0]), 0.5).tolist() == [4194304]   # 4194303.5 rounds away from zero
    # random cross-check against Python's decimal-free definition of round-half-away
    rng = np.random.default_rng(5)
    y = rng.uniform(-1.2, 1.2, 20000)
    s = y * m
    ref = np.clip(np.where(s >= 0, np.floor(s + 0.5), -np.floor(-s + 0.5)), -m, m).astype(np.int64)
    assert np.array_equal(wav.pcm24_round(y).astype(np.int64), ref)


def test_normalize_scale_rule():
    assert wav.normalize_scale(np.array([0.5, -0.6]), True) == 1.0          # peak <= 0.7: untouched
    assert wav.normalize_scale(np.array([0.5, -1.4]), True) == 0.7 / 1.4
    assert wav.normalize_scale(np.array([0.5, -1.4]), False) == 1.0


def test_wav_roundtrip_and_header(tmp_path):
    rng = np.random.default_rng(1)
    q = rng.integers(-8388607, 8388608, 1001).astype(np.int32)   # odd byte count -> pad byte
    p = tmp_path / "a.wav"
    wav.write_wav_pcm24(str(p), q, 44100)
    raw = p.read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt "
    fmt = struct.unpack("<IHHIIHH", raw[16:36])
    assert fmt == (16, 1, 1, 44100, 44100 * 3, 3, 24)
    assert raw[36:40] == b"data" and struct.unpack("<I", raw[40:44])[0] == 3003
    assert struct.unpack("<I", raw[4:8])[0] == len(raw) - 8
    q2, sr = wav.read_wav_pcm24(str(p))
    assert sr == 44100 and np.array_equal(q, q2)
    # python's own wave module agrees on the layout
    import wave
    with wave.open(str(p), "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 3, 44100, 1001)


def test_flag_helpers_follow_reference_semantics():
    a = ["--note", "72", "--volume", "abc", "--no-mlp", "--output"]
    assert preamp_bench.parse_flag(a, "--note", 60.0) == 72.0
    assert preamp_bench.parse_flag(a, "--volume", 0.6) == 0.6        # unparsable -> default (main.rs:101)
    assert preamp_bench.parse_flag(a, "--output", 1.0) == 1.0        # flag in last position has no value (main.rs:99)
    assert preamp_bench.has_flag(a, "--no-mlp") and not preamp_bench.has_flag(a, "--normalize")
    assert preamp_bench._as_u8(300.0) == 255 and preamp_bench._as_u8(-4.0) == 0 and preamp_bench._as_u8(60.9) == 60
    assert preamp_bench.to_dbfs(0.0) == -120.0 and abs(preamp_bench.to_dbfs(0.1) + 20.0) < 1e-12
    assert reed_renderer.midi_note_name(60) == "C4" and reed_renderer.midi_note_name(33) == "A1" and reed_renderer.midi_note_name(61) == "Cs4"


def test_reed_renderer_cli_rejects_bad_input(capsys):
    assert reed_renderer.main(["--note", "20"]) == 1
    assert "out of range" in capsys.readouterr().err
    assert reed_renderer.main(["--bogus"]) == 1
    assert reed_renderer.main(["--help"]) == 0


@pytest.mark.gpu
def test_cli_renders_match_oracle_wavs(tmp_path):
    """End to end through the CLI mirrors: the 24-bit files equal the oracle's samples pushed through the same quantiser
    (a +-1 LSB slack covers a 1e-9 sample difference landing on a rounding boundary)."""
    import oracle_lib as ol
    out = tmp_path / "r"
    assert reed_renderer.main(["-n", "45,72", "-v", "60,110", "-d", "0.25", "--output-dir", str(out)]) == 0
    for n, name in ((45, "A2"), (72, "C5")):
        for v in (60, 110):
            q, sr = wav.read_wav_pcm24(str(out / f"reed_{name}_v{v}.wav"))
            ref = ol.render_voices([ol.voice_job(midi=n, vel=v, dur=0.25)])[0]
            assert sr == 44100 and q.size == ref.size
            assert np.max(np.abs(q.astype(np.int64) - wav.pcm24_truncate(ref))) <= 1
    p = tmp_path / "pb.wav"
    assert preamp_bench.main(["render", "--note", "57", "--velocity", "90", "--duration", "0.2", "--tremolo-depth", "0.4",
                              "--volume", "0.9", "--normalize", "--output", str(p)]) == 0
    q, sr = wav.read_wav_pcm24(str(p))
    ref = ol.render_bench([ol.bench_job(midi=57, vel=90, dur=0.2, depth=0.4, volume=0.9)])[0]
    sc = wav.normalize_scale(ref, True)
    assert sr == 44100 and np.max(np.abs(q.astype(np.int64) - wav.pcm24_round(ref, sc))) <= 1


# ---- SMF front-end of render-midi (main.rs:1626-1716) ------------------------------------------------------------------------------
def _vlq(n):
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append(0x80 | (n & 0x7F))
        n >>= 7
    return bytes(reversed(out))


def _smf(tracks, division=480, fmt=1):
    body = b"MThd" + struct.pack(">IHHH", 6, fmt, len(tracks), division)
    for t in tracks:
        body += b"MTrk" + struct.pack(">I", len(t)) + t
    return body


def test_smf_parser_time_map_running_status_and_filters():
    from openwurli_b200 import smf
    tempo = lambda us: b"\xFF\x51\x03" + us.to_bytes(3, "big")
    eot = b"\x00\xFF\x2F\x00"
    # track 0: tempo 250000 at t=0 (240 BPM), note 60 on at tick 480, running-status note 64, vel-0 note-on = off, CC64 pedal, sysex
    t0 = (_vlq(0) + tempo(250_000) + _vlq(480) + b"\x90\x3C\x64" + _vlq(0) + b"\x40\x50" + _vlq(240) + b"\x3C\x00"
          + _vlq(0) + b"\xB0\x40\x7F" + _vlq(120) + b"\xB0\x40\x3F" + _vlq(0) + b"\xF0\x03\x01\x02\xF7" + _vlq(0) + b"\x80\x40\x00"
          + _vlq(0) + b"\xC0\x05" + eot)
    # track 1 keeps the default tempo (500000): the reference's time map is per track
    t1 = _vlq(960) + b"\x91\x30\x7F" + _vlq(960) + b"\x81\x30\x00" + eot
    data = _smf([t0, t1])
    ev = smf.timed_events(data)
    kinds = [(round(t, 9), k, n, v) for t, k, n, v in ev]
    assert kinds == [(0.25, "on", 60, 100), (0.25, "on", 64, 80), (0.375, "off", 60, 0), (0.375, "pedal", 1, 0),
                     (0.4375, "pedal", 0, 0), (0.4375, "off", 64, 0), (1.0, "on", 48, 127), (2.0, "off", 48, 0)]
    assert [e[2] for e in smf.timed_events(data, track_filter=1)] == [48, 48]
    assert smf.total_samples(ev, 2.0) == int(4.0 * 44100.0)
    # events apply at the first 64-sample chunk boundary at or after their time
    assert smf.chunk_of(0.0) == 0 and smf.chunk_of(64 / 44100.0) == 1 and smf.chunk_of(64 / 44100.0 + 1e-12) == 2
    ee = smf.engine_events(ev)
    assert ee[0][0] == smf.chunk_of(0.25) * 64 and ee[0][1] == 0 and abs(ee[0][3] - 100 / 127) < 1e-7
    with pytest.raises(smf.SmfError):
        smf.timed_events(_smf([t0], division=0xE728))   # SMPTE timing is rejected like main.rs:1633-1636
    with pytest.raises(smf.SmfError):
        smf.timed_events(b"RIFFxxxx")


def test_smf_parser_edge_cases_follow_midly_non_strict():
    """midly 0.5 as the reference builds it (no `strict` feature): meta and sysex events cancel running status, a system common / realtime
    status byte or a truncated event ends the track and keeps what was read, no exception escapes."""
    from openwurli_b200 import smf
    eot = b"\x00\xFF\x2F\x00"
    on = _vlq(0) + b"\x90\x3C\x64"
    # a data byte after a meta event has no running status any more: the track ends there, the note before it is kept
    t = on + _vlq(10) + b"\xFF\x01\x01\x41" + _vlq(0) + b"\x3E\x50" + _vlq(0) + b"\x90\x40\x64" + eot
    assert [(k, n) for _, k, n, _ in smf.timed_events(_smf([t]))] == [("on", 60)]
    # the same after a sysex
    t = on + _vlq(0) + b"\xF0\x02\x01\xF7" + _vlq(0) + b"\x3E\x50" + eot
    assert [(k, n) for _, k, n, _ in smf.timed_events(_smf([t]))] == [("on", 60)]
    # F8 (timing clock) / F2 (song position) cannot occur in an SMF: end of track, not a channel message with two data bytes
    for bad in (b"\xF8", b"\xF2\x01\x02"):
        t = on + _vlq(0) + bad + _vlq(0) + b"\x90\x40\x64" + eot
        assert [(k, n) for _, k, n, _ in smf.timed_events(_smf([t]))] == [("on", 60)]
    # truncated in the middle of a channel message, of a delta and of a meta payload
    for cut in (on + _vlq(0) + b"\x90\x40", on + b"\x81", on + _vlq(0) + b"\xFF\x51\x03\x07"):
        assert [(k, n) for _, k, n, _ in smf.timed_events(_smf([cut]))] == [("on", 60)]
    # a second, healthy track is unaffected by a broken first one
    ev = smf.timed_events(_smf([on + _vlq(0) + b"\xF8", _vlq(480) + b"\x91\x30\x7F" + eot]))
    assert [(k, n) for _, k, n, _ in ev] == [("on", 60), ("on", 48)]


@pytest.mark.gpu
def test_render_midi_cli_matches_oracle(tmp_path):
    """SMF file -> events -> render-midi voice manager + chain -> 24-bit WAV, against the oracle's restatement of cmd_render_midi."""
    import oracle_lib as ol
    from openwurli_b200 import smf
    eot = b"\x00\xFF\x2F\x00"
    trk = (_vlq(0) + b"\xFF\x51\x03" + (300_000).to_bytes(3, "big") + _vlq(0) + b"\x90\x3C\x64" + _vlq(120) + b"\x43\x50" + _vlq(120) + b"\xB0\x40\x7F"
           + _vlq(120) + b"\x90\x3C\x00" + _vlq(240) + b"\xB0\x40\x00" + _vlq(0) + b"\x80\x43\x00" + eot)
    mid, out = tmp_path / "a.mid", tmp_path / "a.wav"
    mid.write_bytes(_smf([trk], division=480, fmt=0))
    assert preamp_bench.main(["render-midi", "--midi", str(mid), "--output", str(out), "--tail", "0.2", "--volume", "0.7", "--speaker", "0.5"]) == 0
    q, sr = wav.read_wav_pcm24(str(out))
    ev = smf.timed_events(mid.read_bytes())
    n = smf.total_samples(ev, 0.2)
    assert sr == 44100 and q.size == n
    import ctypes as C
    from openwurli_b200 import _abi
    arr = (_abi.MidiEvent * len(ev))()
    for i, (t, kind, a, b) in enumerate(ev):
        code = 0 if kind == smf.NOTE_ON else (1 if kind == smf.NOTE_OFF else 2)
        arr[i] = _abi.MidiEvent(t, code, a if code != 2 else 0, b if code == 0 else (a if code == 2 else 0), 0, 0)
    ref = np.zeros(n)
    assert ol.lib().owo_render_midi(C.cast(arr, C.c_void_p), len(ev), n, 0.7, 0.5, 0, 0, ol.dptr(ref), None) == 0
    assert np.abs(ref).max() > 1e-3
    assert np.max(np.abs(q.astype(np.int64) - wav.pcm24_round(ref, 1.0))) <= 1
    # --engine: the same schedule through the plugin's WurliEngine instead
    out2 = tmp_path / "b.wav"
    assert preamp_bench.main(["render-midi", "--midi", str(mid), "--output", str(out2), "--tail", "0.2", "--volume", "0.7", "--speaker", "0.5", "--engine"]) == 0
    q2, _ = wav.read_wav_pcm24(str(out2))
    ref2 = ol.render_engines([ol.engine_job(smf.engine_events(ev), sr=44100.0, dur=n / 44100.0 + 0.5 / 44100.0, volume=0.7, depth=0.0, speaker=0.5,
                                            block=64, warm_up=True)])[0][:n].astype(np.float64)
    assert np.max(np.abs(q2.astype(np.int64) - wav.pcm24_round(ref2, 1.0))) <= 2


@pytest.mark.gpu
def test_render_poly_signals_match_oracle():
    """render-poly: shared-chain mix, sum of separately processed voices and the intermod residual against the oracle."""
    import oracle_lib as ol
    notes, vels, dur = [38, 59, 62], [45, 40, 40], 0.3
    fin, sep, res = preamp_bench.render_poly(notes, vels, duration=dur, volume=0.6, speaker_char=1.0, r_ldr=1e6)
    n = int(dur * 44100.0)
    voices = ol.render_voices([ol.voice_job(midi=nt, vel=v, dur=dur, mlp=True, seed=(nt * 2654435761 + i) & 0xFFFFFFFF)
                               for i, (nt, v) in enumerate(zip(notes, vels))])[:, :n]
    mix = np.zeros(n)
    for vb in voices:
        mix += vb
    rows = np.ascontiguousarray(np.vstack([mix[None, :], voices]))
    p = ol.bench_job(r_ldr=1e6, volume=0.6, speaker=1.0)
    arr = (ol.BenchJob * 4)(*[p] * 4)
    ref = np.zeros_like(rows)
    assert ol.lib().owo_chain_batch(ol.dptr(rows), n, 4, n, arr, 1, ol.dptr(ref), n, 4, 0) == 0
    rsep = np.zeros(n)
    for k in range(1, 4):
        rsep += ref[k]
    assert np.abs(fin - ref[0]).max() <= 1e-6 and np.abs(sep - rsep).max() <= 1e-6
    assert np.abs(res - (ref[0] - rsep)).max() <= 1e-6 and np.abs(res).max() > 1e-7   # shared nonlinearity: the residual is not zero


@pytest.mark.gpu
def test_alias_audit_cli_json_matches_the_reference_fixture(capsys):
    """`preamp-bench alias-audit --json` (main.rs:985-1017) for the three fixture notes in one device batch: same keys and number formats as the
    reference's hand-rolled JSON; f0 equal to the fixture's value, the regression gate of alias_audit_regression.rs on the printed numbers."""
    import json
    from openwurli_b200.cli import preamp_bench
    base = {e["note"]: e for e in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_alias_audit_v0_5_1.json")))["entries"]}
    assert preamp_bench.main(["alias-audit", "--notes", "72,84,91", "--velocity", "120", "--json"]) == 0
    out = capsys.readouterr().out
    docs = [json.loads(d + "}") for d in out.strip()[:-1].split("}\n")]   # three JSON documents, one after the other
    assert len(docs) == 3
    for note, d in zip((72, 84, 91), docs):
        assert list(d.keys()) == ["f0_hz", "h1_dbfs", "harmonic_dbc", "max_step_up_db", "max_step_up_from_harmonic", "hf_band_dbc"]
        assert d["f0_hz"] == base[note]["f0_hz"] and len(d["harmonic_dbc"]) == 12 and d["harmonic_dbc"][0] == 0.0
        assert d["max_step_up_db"] - base[note]["max_step_up_db"] <= 1.5 and d["hf_band_dbc"] - base[note]["hf_band_dbc"] <= 2.0
        assert d["max_step_up_from_harmonic"] == base[note]["max_step_up_from_harmonic"]
    assert preamp_bench.main(["alias-audit", "--note", "84"]) == 0
    text = capsys.readouterr().out
    assert text.startswith("Click-band alias audit\n  Stimulus:   note=84 vel=120 vol=0.50") and "(* = harmonics in plateau-detection band)" in text
