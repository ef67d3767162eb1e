"""Standard MIDI File front-end of `preamp-bench render-midi` (tools/preamp-bench/src/main.rs:1626-1716; the reference parses
with the un-vendored `midly` 0.5 crate, whose published SMF semantics are restated here):

* header `MThd` (format, n_tracks, division); only metrical timing (ticks per beat) is accepted, like main.rs:1630-1636;
* per track: variable-length deltas, running status, meta events (FF type len data), sysex (F0/F7 len data);
* time map exactly as the reference builds it: EVERY track starts at tempo 500 000 us/beat and follows only the tempo meta
  events found in that same track (main.rs:1655-1669), `time_s += (delta / ticks_per_beat) * (tempo / 1e6)` in f64;
* note-on with velocity 0 is a note-off; controller 64 with value >= 64 is pedal down, else pedal up; everything else is dropped;
* `--track N` keeps the notes of track N only but still walks every track (tempo is per track anyway);
* the merged list is sorted by time with a stable sort (ties keep track order, then file order).
"""
import struct

NOTE_ON, NOTE_OFF, PEDAL = "on", "off", "pedal"


class SmfError(ValueError):
    pass


def _vlq(data, pos):
    val = 0
    for _ in range(4):
        if pos >= len(data):
            raise SmfError("truncated variable-length quantity")
        b = data[pos]
        pos += 1
        val = (val << 7) | (b & 0x7F)
        if not b & 0x80:
            return val, pos
    raise SmfError("variable-length quantity longer than 4 bytes")


def parse_tracks(data):
    """-> (ticks_per_beat, [track]) with track = list of (delta_ticks, kind, a, b):
    kind 'tempo' (a = us per beat), 'on' (a = key, b = vel), 'off' (a = key), 'cc' (a = controller, b = value), or None."""
    if len(data) < 14 or data[:4] != b"MThd":
        raise SmfError("not a Standard MIDI File (no MThd)")
    hlen, fmt, ntrk, division = struct.unpack(">IHHH", data[4:14])
    if division & 0x8000:
        raise SmfError("Only metrical (ticks per beat) MIDI timing is supported")
    pos = 8 + hlen
    tracks = []
    while pos + 8 <= len(data) and len(tracks) < ntrk:
        tag, tlen = data[pos:pos + 4], struct.unpack(">I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + tlen]
        pos += 8 + tlen
        if tag != b"MTrk":
            continue  # alien chunk: skipped
        ev, p, status = [], 0, None
        # midly 0.5 without its `strict` feature (the reference's Cargo.toml enables none): a malformed event -- truncated data, a
        # data byte with no running status, a system common / realtime status (F1-F6, F8-FE), which an SMF cannot contain -- silently
        # ends the track, keeping what was read; meta and sysex events cancel running status
        try:
            while p < len(body):
                delta, p = _vlq(body, p)
                if p >= len(body):
                    break
                b0 = body[p]
                if b0 == 0xFF:  # meta
                    if p + 1 >= len(body):
                        break
                    mtype = body[p + 1]
                    mlen, q = _vlq(body, p + 2)
                    if q + mlen > len(body):
                        break
                    payload = body[q:q + mlen]
                    p = q + mlen
                    status = None
                    if mtype == 0x51 and mlen == 3:
                        ev.append((delta, "tempo", int.from_bytes(payload, "big"), 0))
                    else:
                        ev.append((delta, None, 0, 0))
                    if mtype == 0x2F:
                        break
                    continue
                if b0 in (0xF0, 0xF7):  # sysex / escape
                    slen, q = _vlq(body, p + 1)
                    if q + slen > len(body):
                        break
                    p = q + slen
                    status = None
                    ev.append((delta, None, 0, 0))
                    continue
                if b0 >= 0xF1:  # F1-F6, F8-FE: not allowed in a Standard MIDI File
                    break
                if b0 & 0x80:
                    status = b0
                    p += 1
                elif status is None:
                    break
                hi = status & 0xF0
                n_data = 1 if hi in (0xC0, 0xD0) else 2
                d = body[p:p + n_data]
                if len(d) < n_data:
                    break
                p += n_data
                if hi == 0x90:
                    ev.append((delta, "on", d[0] & 0x7F, d[1] & 0x7F))
                elif hi == 0x80:
                    ev.append((delta, "off", d[0] & 0x7F, d[1] & 0x7F))
                elif hi == 0xB0:
                    ev.append((delta, "cc", d[0] & 0x7F, d[1] & 0x7F))
                else:
                    ev.append((delta, None, 0, 0))
        except SmfError:
            pass  # a truncated variable-length quantity: the same lenient end of track
        tracks.append(ev)
    return float(division), tracks


def timed_events(data, track_filter=None):
    """The reference's `events` vector: [(time_s, kind, note, velocity)] sorted by time (stable)."""
    tpb, tracks = parse_tracks(data)
    out = []
    for ti, trk in enumerate(tracks):
        tempo, time_s = 500_000.0, 0.0
        emit = track_filter is None or track_filter == ti
        for delta, kind, a, b in trk:
            time_s += (float(delta) / tpb) * (tempo / 1_000_000.0)
            if kind == "tempo":
                tempo = float(a)
            elif not emit:
                continue
            elif kind == "on":
                out.append((time_s, NOTE_OFF, a, 0) if b == 0 else (time_s, NOTE_ON, a, b))
            elif kind == "off":
                out.append((time_s, NOTE_OFF, a, 0))
            elif kind == "cc" and a == 64:
                out.append((time_s, PEDAL, 1 if b >= 64 else 0, 0))
    out.sort(key=lambda e: e[0])
    return out


def total_samples(events, tail_seconds=2.0, sample_rate=44100.0):
    """main.rs:1719-1721: (last_event_time + tail) * BASE_SR, truncated."""
    if not events:
        return 0
    return int((events[-1][0] + tail_seconds) * sample_rate)


def chunk_of(time_s, chunk=64, sample_rate=44100.0):
    """First processing chunk whose start time `sample_pos / BASE_SR` is >= time_s (main.rs:1790-1795 applies an event at the
    first 64-sample chunk boundary at or after it)."""
    k = max(int(time_s * sample_rate / chunk), 0)
    while (k * chunk) / sample_rate < time_s:
        k += 1
    while k > 0 and ((k - 1) * chunk) / sample_rate >= time_s:
        k -= 1
    return k


def engine_events(events, block=64, sample_rate=44100.0):
    """The same schedule as WurliEngine events (sample index = start of the block the event is applied in; velocity as the f32
    the plugin passes, vel/127)."""
    import numpy as np
    from .api import NOTE_OFF as E_OFF, NOTE_ON as E_ON, SUSTAIN as E_SUS
    out = []
    for t, kind, a, b in events:
        s = chunk_of(t, block, sample_rate) * block
        if kind == NOTE_ON:
            out.append((s, E_ON, min(max(a, 33), 96), float(np.float32(b / 127.0))))
        elif kind == NOTE_OFF:
            out.append((s, E_OFF, min(max(a, 33), 96), 0.0))
        else:
            out.append((s, E_SUS, a, 0.0))
    return out
