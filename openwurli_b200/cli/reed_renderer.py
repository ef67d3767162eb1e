"""`reed-renderer` mirror (tools/reed-renderer/src/main.rs:11-107): same flags, same file names, same 24-bit WAVs --
but every (note, velocity) pair of the invocation is rendered in ONE device batch.

    python -m openwurli_b200.cli.reed_renderer --sweep -v 40,80,127 -d 2.0 --output-dir out/
"""
import math
import os
import sys

import numpy as np

from .. import api, wav

SAMPLE_RATE = 44100.0
MIDI_LO, MIDI_HI = 33, 96  # tables.rs:14-15
USAGE = """Usage: reed_renderer [OPTIONS]
  -n, --note <N[,N..]>       MIDI note(s) 33-96 (default 60)
  -v, --velocity <V[,V..]>   MIDI velocity 0-127 (default 100)
  -d, --duration <SECS>      seconds (default 1.0)
  -o, --output <FILE>        output file (single note/velocity only)
      --output-dir <DIR>     output directory (default ".")
      --sweep                notes 33,45,57,60,69,72,81,93,96
"""


def midi_note_name(midi):  # main.rs:128-135
    names = ["C", "Cs", "D", "Ds", "E", "F", "Fs", "G", "Gs", "A", "As", "B"]
    return f"{names[midi % 12]}{midi // 12 - 1}"


def main(argv=None):
    args = list(sys.argv[1:] if argv is None else argv)
    notes, velocities, duration, output_file, output_dir = [], [], 1.0, None, "."
    i = 0
    while i < len(args):
        a = args[i]
        if a in ("--note", "-n"):
            i += 1
            notes += [int(s.strip()) for s in args[i].split(",")]
        elif a in ("--velocity", "-v"):
            i += 1
            velocities += [int(s.strip()) for s in args[i].split(",")]
        elif a in ("--duration", "-d"):
            i += 1
            duration = float(args[i])
        elif a in ("--output", "-o"):
            i += 1
            output_file = args[i]
        elif a == "--output-dir":
            i += 1
            output_dir = args[i]
        elif a == "--sweep":
            notes = [33, 45, 57, 60, 69, 72, 81, 93, 96]
        elif a in ("--help", "-h"):
            sys.stderr.write(USAGE)
            return 0
        else:
            sys.stderr.write(f"Unknown argument: {a}\n{USAGE}")
            return 1
        i += 1
    notes = notes or [60]
    velocities = velocities or [100]
    for n in notes:
        if not MIDI_LO <= n <= MIDI_HI:
            sys.stderr.write(f"MIDI note {n} out of range ({MIDI_LO}-{MIDI_HI})\n")
            return 1
    pairs = [(n, v) for n in notes for v in velocities]
    # Voice::render_note (voice.rs:198-221): MLP off, default seed, attack noise on
    jobs = [api.voice_job(midi=n, velocity=v, sample_rate=SAMPLE_RATE, duration=duration) for n, v in pairs]
    out = api.render_voices(jobs)
    single = len(pairs) == 1
    for k, (n, v) in enumerate(pairs):
        name = midi_note_name(n)
        filename = output_file if (output_file and single) else f"{output_dir}/reed_{name}_v{v}.wav"
        samples = out[k]
        peak = float(np.max(np.abs(samples))) if samples.size else 0.0
        db = 20.0 * math.log10(peak) if peak > 0 else float("-inf")
        sys.stderr.write(f"Rendering MIDI {n} ({name}) vel={v} dur={duration}s → {filename}\n")
        sys.stderr.write(f"  Peak amplitude: {peak:.6f} ({db:.1f} dBFS)\n")
        os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
        wav.write_reed_renderer_wav(filename, samples, int(SAMPLE_RATE))
        sys.stderr.write(f"  Written: {filename}\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
