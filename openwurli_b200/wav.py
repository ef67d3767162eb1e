"""24-bit mono PCM WAV output in the two quantisations the reference tools use (both through hound 3.5's
16-byte-fmt integer PCM writer; hound is un-vendored, the canonical RIFF layout is restated here):

* `reed-renderer` (tools/reed-renderer/src/main.rs:110-126): clamp to [-1, 1], scale by 2^23 - 1, truncate toward zero.
* `preamp-bench render` (tools/preamp-bench/src/main.rs:941-957): x * scale * (2^23 - 1), round half away from zero,
  clamp to +-(2^23 - 1); `--normalize` sets scale = 0.7 / peak when peak > 0.7 (main.rs:503-507).
"""
import struct

import numpy as np

PCM24_MAX = (1 << 23) - 1


def pcm24_truncate(samples):
    """reed-renderer quantisation: `(clamp(s, -1, 1) * 8388607.0) as i32`."""
    s = np.clip(np.asarray(samples, dtype=np.float64), -1.0, 1.0) * float(PCM24_MAX)
    s = np.where(np.isnan(s), 0.0, s)  # Rust `NaN as i32` == 0
    return np.trunc(s).astype(np.int32)


def pcm24_round(samples, scale=1.0):
    """preamp-bench quantisation: `(s * scale * 8388607.0).round() as i32` then clamp (f64::round = half away from zero;
    the saturating `as i32` cast is covered by clamping in f64 first)."""
    s = np.asarray(samples, dtype=np.float64) * scale * float(PCM24_MAX)
    r = np.copysign(np.floor(np.abs(s) + 0.5), s)
    # floor(|s| + 0.5) differs from round-half-away only when |s| + 0.5 rounds up in f64 (|s| = 0.49999999999999994)
    r = np.where(np.abs(s) == 0.49999999999999994, 0.0, r)
    r = np.where(np.isnan(r), 0.0, r)
    return np.clip(r, -float(PCM24_MAX), float(PCM24_MAX)).astype(np.int32)


def normalize_scale(samples, normalize):
    """preamp-bench `--normalize` (main.rs:503-507)."""
    if not normalize:
        return 1.0
    peak = float(np.max(np.abs(samples))) if len(samples) else 0.0
    return 0.7 / peak if peak > 0.7 else 1.0


def pack_pcm24(q):
    """int32 -> little-endian 3-byte samples."""
    q = np.asarray(q, dtype=np.int32)
    b = q.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
    return np.ascontiguousarray(b).tobytes()


def write_wav_pcm24(path, q, sample_rate):
    """Canonical 44-byte-header RIFF/WAVE, PCM (format tag 1), 1 channel, 24 bit."""
    data = pack_pcm24(q)
    if len(data) & 1:
        pad = b"\0"
    else:
        pad = b""
    sr = int(sample_rate)  # `sample_rate as u32`
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data) + len(pad)) + b"WAVE"
    hdr += b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sr, sr * 3, 3, 24)
    hdr += b"data" + struct.pack("<I", len(data))
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(data)
        f.write(pad)


def read_wav_pcm24(path):
    """Inverse of write_wav_pcm24 (tests, round trips). Returns (int32 samples, sample_rate)."""
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE"
    pos, sr, data = 12, None, None
    while pos + 8 <= len(raw):
        tag, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + size]
        if tag == b"fmt ":
            fmt, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            assert (fmt, ch, bits) == (1, 1, 24)
        elif tag == b"data":
            data = body
        pos += 8 + size + (size & 1)
    b = np.frombuffer(data, dtype=np.uint8).reshape(-1, 3)
    q = b[:, 0].astype(np.int32) | (b[:, 1].astype(np.int32) << 8) | (b[:, 2].astype(np.int8).astype(np.int32) << 16)
    return q, sr


def write_reed_renderer_wav(path, samples, sample_rate=44100):
    write_wav_pcm24(path, pcm24_truncate(samples), sample_rate)


def write_preamp_bench_wav(path, samples, sample_rate, scale=1.0):
    write_wav_pcm24(path, pcm24_round(samples, scale), sample_rate)
