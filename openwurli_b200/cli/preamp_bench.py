"""`preamp-bench render` mirror (tools/preamp-bench/src/main.rs:371-553): same flags, defaults, report and 24-bit WAV.

    python -m openwurli_b200.cli.preamp_bench render --note 60 --velocity 100 --tremolo-depth 0.5 --output c4.wav

Batch extension (not in the reference): `--note` / `--velocity` accept comma-separated lists; all pairs are rendered in one
device batch and written to `<output stem>_n<NOTE>_v<VEL>.wav`.  `calibrate` prints the T5 columns of `run_calibrate`
(main.rs:1127-1260) for a notes x velocities grid through the on-device analysis reductions.
"""
import math
import os
import sys
import tempfile

import numpy as np

from .. import api, wav

BASE_SR = 44100.0


def parse_flag(args, flag, default):  # main.rs:98-105 (a malformed value falls back to the default)
    for i in range(max(len(args) - 1, 0)):
        if args[i] == flag:
            try:
                return float(args[i + 1])
            except ValueError:
                return default
    return default


def parse_flag_str(args, flag, default):  # main.rs:107-114
    for i in range(max(len(args) - 1, 0)):
        if args[i] == flag:
            return args[i + 1]
    return default


def has_flag(args, flag):  # main.rs:116-118
    return flag in args


def to_dbfs(val):  # main.rs:2241-2247
    return 20.0 * math.log10(val) if val > 1e-15 else -120.0


def _as_u8(x):  # Rust `f64 as u8`: saturating, NaN -> 0
    if x != x:
        return 0
    return int(min(max(math.trunc(x), 0), 255))


def _list_flag(args, flag, default):
    s = parse_flag_str(args, flag, None)
    if s is None or "," not in s:
        return [_as_u8(parse_flag(args, flag, default))]
    return [_as_u8(float(t)) for t in s.split(",")]


def cmd_render(args):
    notes = _list_flag(args, "--note", 60.0)
    velocities = _list_flag(args, "--velocity", 100.0)
    duration = parse_flag(args, "--duration", 2.0)
    r_ldr = parse_flag(args, "--ldr", 1_000_000.0)
    volume = parse_flag(args, "--volume", 0.60)
    speaker_char = parse_flag(args, "--speaker", 1.0)
    tremolo_depth = parse_flag(args, "--tremolo-depth", 0.0)
    sample_rate = parse_flag(args, "--sample-rate", BASE_SR)
    no_poweramp = has_flag(args, "--no-poweramp")
    # The power amplifier is the crate's other compile-time choice (`legacy-power-amp`, in the default features): behavioural model or the
    # melange 7-BJT solver.  Here it is a flag.  --no-rail-sag (main.rs:381, 481-483) is a no-op on the behavioural path
    # (power_amp.rs:255-257) and selects ideal rails on the melange one.
    no_rail_sag = has_flag(args, "--no-rail-sag")
    pa_name = parse_flag_str(args, "--power-amp", "behavioral")
    if pa_name not in ("behavioral", "melange"):
        sys.stderr.write(f"Unknown --power-amp {pa_name} (behavioral | melange)\n")
        return 1
    pa_model = api.PA_BEHAVIORAL if pa_name == "behavioral" else (api.PA_MELANGE_IDEAL_RAILS if no_rail_sag else api.PA_MELANGE)
    no_preamp = has_flag(args, "--no-preamp")
    no_attack_noise = has_flag(args, "--no-attack-noise")
    no_mlp = has_flag(args, "--no-mlp")
    normalize = has_flag(args, "--normalize")
    disp_scale = parse_flag(args, "--displacement-scale", 0.30) if has_flag(args, "--displacement-scale") else None
    output_path = parse_flag_str(args, "--output", os.path.join(tempfile.gettempdir(), "preamp_render.wav"))
    do_oversample = sample_rate < 88200.0
    # The reference picks the preamp at compile time (cargo feature `melange-preamp`; its default build is the legacy 8-node
    # solver, openwurli-dsp/Cargo.toml:10-19).  Here it is a flag; the default is the north-star melange 12-node model.
    model_name = parse_flag_str(args, "--preamp-model", "melange12")
    if model_name not in ("melange12", "legacy8"):
        sys.stderr.write(f"Unknown --preamp-model {model_name} (melange12 | legacy8)\n")
        return 1
    preamp_model = api.LEGACY8 if model_name == "legacy8" else api.MELANGE12

    pairs = [(n, v) for n in notes for v in velocities]
    jobs = [api.bench_job(note=n, velocity=v, duration=duration, ldr=r_ldr, volume=volume, speaker=speaker_char,
                          tremolo_depth=tremolo_depth, sample_rate=sample_rate, no_poweramp=no_poweramp, no_preamp=no_preamp,
                          no_attack_noise=no_attack_noise, no_mlp=no_mlp, displacement_scale=disp_scale) for n, v in pairs]
    out = api.render_bench(jobs, preamp_model=preamp_model, power_amp_model=pa_model)
    stem, ext = os.path.splitext(output_path)
    for k, (note, velocity) in enumerate(pairs):
        final_output = out[k]
        path = output_path if len(pairs) == 1 else f"{stem}_n{note}_v{velocity}{ext or '.wav'}"
        peak = float(np.max(np.abs(final_output))) if final_output.size else 0.0
        peak_dbfs = to_dbfs(peak)
        scale = wav.normalize_scale(final_output, normalize)
        if not normalize and peak > 1.0:
            sys.stderr.write(f"WARNING: Peak exceeds 0 dBFS ({peak_dbfs:.1f} dBFS) — consider reducing --volume\n")
        wav.write_preamp_bench_wav(path, final_output, sample_rate, scale)
        print("Render complete")
        print(f"  Note:      MIDI {note}")
        print(f"  Velocity:  {velocity}")
        print(f"  Duration:  {duration:.1f}s")
        if tremolo_depth > 0.0:
            print(f"  Tremolo:   depth={tremolo_depth:.2f}")
        else:
            print(f"  LDR:       {r_ldr:.0f} Ω (static)")
        print(f"  Volume:    {volume:.3f} (PA gain: 69x, headroom: 22V)")
        print(f"  Speaker:   {speaker_char:.1f}")
        if disp_scale is not None:
            print(f"  Disp scale: {disp_scale:.3f}")
        if no_preamp:
            print("  Preamp:    BYPASSED")
        if no_poweramp:
            print("  Power amp: BYPASSED")
        if normalize:
            print("  Normalize: ON (-3 dBFS ceiling)")
        if sample_rate != BASE_SR:
            print(f"  Sample rate: {sample_rate:.0f} Hz (oversample: {'on' if do_oversample else 'off'})")
        print(f"  Peak:      {peak_dbfs:.1f} dBFS (raw)")
        print(f"  Build:     openwurli_b200 (libowgpu ABI {api.lib().owg_abi_version()})")
        print(f"  Output:    {path}")
    return 0


def cmd_sensitivity(args):
    """`preamp-bench sensitivity` (main.rs:1313-1386): calibrate at every DS_AT_C4 of --ds-range; one device batch per DS value."""
    notes = _csv_u8_list(args, "--notes", "36,48,54,60,66,72,78,84")
    velocities = _csv_u8_list(args, "--velocities", "40,80,127")
    ds_values = []
    for t in parse_flag_str(args, "--ds-range", "0.50,0.55,0.60,0.65,0.70,0.75,0.80,0.85").split(","):
        try:
            ds_values.append(float(t.strip()))
        except ValueError:
            pass
    volume = parse_flag(args, "--volume", 0.40)
    speaker_char = parse_flag(args, "--speaker", 1.0)
    scale_mode = "zero-trim" if has_flag(args, "--zero-trim") else parse_flag_str(args, "--scale-mode", "track")
    output_path = parse_flag_str(args, "--output", os.path.join(tempfile.gettempdir(), "sensitivity.csv"))
    model = api.LEGACY8 if parse_flag_str(args, "--preamp-model", "melange12") == "legacy8" else api.MELANGE12
    sys.stderr.write(f"Sensitivity: {len(ds_values)} DS × {len(notes)} notes × {len(velocities)} vel = "
                     f"{len(ds_values) * len(notes) * len(velocities)} renders\n")
    lines = [CALIBRATE_HEADER]
    for ds in ds_values:
        if scale_mode == "freeze":
            cfg = api.calib_cfg(ds_at_c4=0.85)
        elif scale_mode == "zero-trim":
            cfg = api.calib_cfg(ds_at_c4=ds, zero_trim=True)
        else:
            cfg = api.calib_cfg(ds_at_c4=ds)
        rows = api.render_calibrate(notes, velocities, cfg, volume=volume, speaker=speaker_char, preamp_model=model)
        rows[:, 0] = ds  # the ds_at_c4 column is stamped with the sweep value (main.rs:1371-1374)
        lines += calibrate_csv_lines(notes, velocities, rows)[1:]
    with open(output_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    sys.stderr.write(f"Sensitivity: {len(lines) - 1} total rows → {output_path}\n")
    return 0


def midi_note_name(note):  # main.rs:666-672 (sharps with '#', unlike reed-renderer's file names)
    names = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]
    return f"{names[note % 12]}{note // 12 - 1}"


def _csv_u8_list(args, flag, default):  # parse_csv_list::<u8>: unparsable / out-of-range items are dropped
    out = []
    for t in parse_flag_str(args, flag, default).split(","):
        try:
            v = int(t.strip())
        except ValueError:
            continue
        if 0 <= v <= 255:
            out.append(v)
    return out


CALIBRATE_HEADER = ("midi,note_name,velocity,ds_at_c4,ds_actual,y_peak,t2_peak_db,t2_rms_db,t2_h2_h1_db,t3_peak_db,t3_rms_db,"
                    "t4_peak_db,t4_rms_db,t4_h2_h1_db,t5_peak_db,t5_rms_db,t5_h2_h1_db,proxy_db,trim_db,proxy_error_db,tanh_compression_db")


def calibrate_csv_lines(notes, velocities, rows):
    """write_calibrate_csv (main.rs:1262-1310): 4 decimals for the three scale columns, 2 for the dB columns."""
    lines = [CALIBRATE_HEADER]
    k = 0
    for n in notes:
        for v in velocities:
            r = rows[k]
            k += 1
            lines.append(f"{n},{midi_note_name(n)},{v}," + ",".join(f"{x:.4f}" for x in r[:3]) + "," + ",".join(f"{x:.2f}" for x in r[3:]))
    return lines


def cmd_calibrate(args):
    """`preamp-bench calibrate` (main.rs:1069-1104): same flags and CSV; all rows in one device batch."""
    notes = _csv_u8_list(args, "--notes", "36,40,44,48,52,56,60,64,68,72,76,80,84")
    velocities = _csv_u8_list(args, "--velocities", "40,80,127")
    cfg = api.calib_cfg(ds_at_c4=parse_flag(args, "--ds-at-c4", 0.75), ds_clamp_max=parse_flag(args, "--ds-clamp-max", 0.82),
                        zero_trim=has_flag(args, "--zero-trim"))
    volume = parse_flag(args, "--volume", 0.40)
    speaker_char = parse_flag(args, "--speaker", 1.0)
    output_path = parse_flag_str(args, "--output", os.path.join(tempfile.gettempdir(), "calibrate.csv"))
    model = api.LEGACY8 if parse_flag_str(args, "--preamp-model", "melange12") == "legacy8" else api.MELANGE12
    rows = api.render_calibrate(notes, velocities, cfg, volume=volume, speaker=speaker_char, preamp_model=model)
    with open(output_path, "w") as f:
        f.write("\n".join(calibrate_csv_lines(notes, velocities, rows)) + "\n")
    sys.stderr.write(f"Calibrate: {len(notes)} notes × {len(velocities)} velocities = {len(rows)} rows → {output_path}\n")
    return 0


def render_poly(notes, velocities, duration=3.0, volume=0.60, speaker_char=1.0, r_ldr=1_000_000.0, no_poweramp=False,
                preamp_model=api.MELANGE12):
    """The signal part of cmd_render_poly (main.rs:1396-1507): K voices rendered alone, summed voice by voice, the sum and every voice
    through its own chain B (static preamp in set-then-reset order).  Returns (final_output, separate_sum, residual)."""
    n = int(duration * BASE_SR) if duration > 0 else 0
    jobs = [api.voice_job(midi=nt, velocity=v, sample_rate=BASE_SR, duration=duration, mlp=True, seed=(nt * 2654435761 + i) & 0xFFFFFFFF)
            for i, (nt, v) in enumerate(zip(notes, velocities))]
    voices = api.render_voices(jobs)[:, :n]
    mix = np.zeros(n)
    for vb in voices:
        mix += vb                                         # sum_buf[j] += voice_buf[j] (main.rs:1451-1453)
    rows = np.ascontiguousarray(np.vstack([mix[None, :], voices]))
    p = api.bench_job(ldr=r_ldr, volume=volume, speaker=speaker_char, no_poweramp=no_poweramp, sample_rate=BASE_SR)
    out = api.chain_batch(rows, [p] * rows.shape[0], init_order=api.SET_THEN_RESET, preamp_model=preamp_model)
    final_output = out[0]
    separate_sum = np.zeros(n)
    for k in range(1, out.shape[0]):
        separate_sum += out[k]                            # separate_sum[i] += ... (main.rs:1499-1501)
    return final_output, separate_sum, final_output - separate_sum


def _peak_db(x):
    return to_dbfs(float(np.max(np.abs(x))) if x.size else 0.0)


def _rms_db(x):
    m = float(np.sum(x * x)) / x.size if x.size else 0.0
    return 10.0 * math.log10(m) if m > 0.0 else -120.0


def cmd_render_poly(args):
    """`preamp-bench render-poly` (main.rs:1396-1593): same flags, both WAV files and the intermod report."""
    notes = _csv_u8_list(args, "--notes", "38,59,62,66")
    vraw = _csv_u8_list(args, "--velocities", "45,40,40,40")
    duration = parse_flag(args, "--duration", 3.0)
    volume = parse_flag(args, "--volume", 0.60)
    speaker_char = parse_flag(args, "--speaker", 1.0)
    r_ldr = parse_flag(args, "--ldr", 1_000_000.0)
    no_poweramp = has_flag(args, "--no-poweramp")
    normalize = has_flag(args, "--normalize")
    output_path = parse_flag_str(args, "--output", os.path.join(tempfile.gettempdir(), "preamp_render_poly.wav"))
    model = api.LEGACY8 if parse_flag_str(args, "--preamp-model", "melange12") == "legacy8" else api.MELANGE12
    velocities = [vraw[i] if i < len(vraw) else (vraw[-1] if vraw else 80) for i in range(len(notes))]
    sys.stderr.write(f"Rendering {len(notes)} voices, {duration:.1f}s @ {BASE_SR:.0f} Hz...\n")
    final_output, separate_sum, residual = render_poly(notes, velocities, duration, volume, speaker_char, r_ldr, no_poweramp, model)
    n = final_output.size
    ms, me = int(0.2 * BASE_SR), int(min(2.0 * BASE_SR, float(n)))
    poly_peak, sep_peak, res_peak_db = _peak_db(final_output[ms:me]), _peak_db(separate_sum[ms:me]), _peak_db(residual[ms:me])
    poly_rms, sep_rms, res_rms = _rms_db(final_output[ms:me]), _rms_db(separate_sum[ms:me]), _rms_db(residual[ms:me])
    peak = float(np.max(np.abs(final_output))) if n else 0.0
    wav.write_preamp_bench_wav(output_path, final_output, BASE_SR, wav.normalize_scale(final_output, normalize))
    residual_path = output_path.replace(".wav", "_residual.wav")
    res_peak = float(np.max(np.abs(residual))) if n else 0.0
    wav.write_preamp_bench_wav(residual_path, residual, BASE_SR, 0.5 / res_peak if res_peak > 1e-10 else 1.0)
    print("Polyphonic render complete")
    print("  Notes:     [" + ", ".join(f'"{midi_note_name(x)} ({x})"' for x in notes) + "]")
    print(f"  Velocities: {velocities}")
    print(f"  Duration:  {duration:.1f}s")
    print(f"  Volume:    {volume:.3f} (audio taper: {volume * volume:.3f})")
    print(f"  Speaker:   {speaker_char:.1f}")
    print(f"  Peak:      {to_dbfs(peak):.1f} dBFS")
    print()
    print("  === INTERMOD ANALYSIS (0.2-2.0s window) ===")
    print(f"  Shared chain (poly):  peak={poly_peak:.1f} dBFS  rms={poly_rms:.1f} dBFS")
    print(f"  Separate chains (sum): peak={sep_peak:.1f} dBFS  rms={sep_rms:.1f} dBFS")
    print(f"  Residual (intermod):  peak={res_peak_db:.1f} dBFS  rms={res_rms:.1f} dBFS")
    print(f"  Intermod ratio:       {poly_rms - res_rms:.1f} dB below signal")
    print()
    margin = poly_rms - res_rms
    verdict = ("CLEAN — intermod negligible" if margin > 60.0 else "OK — intermod present but likely inaudible" if margin > 40.0 else
               "MARGINAL — intermod may be audible on revealing systems" if margin > 20.0 else "DIRTY — intermod clearly audible")
    print(f"  Verdict: {verdict}")
    print()
    print(f"  Output:    {output_path}")
    print(f"  Residual:  {residual_path} (normalized for listening)")
    return 0


def cmd_render_midi(args):
    """`render-midi --midi FILE` (main.rs:1603-1895): SMF front-end (smf.py), the tool's own voice manager and chain on the device
    (owg_render_midi).  `--engine` renders the same schedule through the plugin's WurliEngine instead (chain E: stealing crossfade,
    tremolo via --tremolo-depth, f32), which is not what the reference tool does."""
    from .. import smf
    midi_path = parse_flag_str(args, "--midi", "")
    if not midi_path:
        sys.stderr.write("Usage: preamp_bench render-midi --midi <file.mid> [--output <file.wav>] [--volume V] [--speaker S] [--tail T] [--track N]\n")
        return 1
    output_path = parse_flag_str(args, "--output", os.path.join(tempfile.gettempdir(), "preamp_render_midi.wav"))
    volume = parse_flag(args, "--volume", 0.60)
    speaker_char = parse_flag(args, "--speaker", 1.0)
    tail = parse_flag(args, "--tail", 2.0)
    depth = parse_flag(args, "--tremolo-depth", 0.0)
    track = int(parse_flag(args, "--track", 0.0)) if has_flag(args, "--track") else None
    model = api.LEGACY8 if parse_flag_str(args, "--preamp-model", "melange12") == "legacy8" else api.MELANGE12
    if has_flag(args, "--engine") and has_flag(args, "--no-poweramp"):
        sys.stderr.write("--no-poweramp is not available on the WurliEngine path\n")
        return 1
    try:
        events = smf.timed_events(open(midi_path, "rb").read(), track)
    except smf.SmfError as e:
        sys.stderr.write(f"{e}\n")
        return 1
    if not events:
        sys.stderr.write("No note events found in MIDI file\n")
        return 1
    n = smf.total_samples(events, tail, BASE_SR)
    if not has_flag(args, "--engine"):
        out = api.render_midi([events], volume=volume, speaker=speaker_char, no_poweramp=has_flag(args, "--no-poweramp"), tail=tail,
                              preamp_model=model)[0]
        peak = float(np.max(np.abs(out))) if out.size else 0.0
        if peak > 1.0:
            sys.stderr.write(f"WARNING: Peak exceeds 0 dBFS ({to_dbfs(peak):.1f} dBFS) — consider reducing --volume\n")
        wav.write_preamp_bench_wav(output_path, out, BASE_SR, 1.0)
        d = api.last_diag()
        print("MIDI render complete")
        print(f"  File:      {midi_path}")
        print(f"  Notes:     {int(d.nr_iter_hist[0])} note-ons")
        print(f"  Peak poly: {int(d.nr_iter_hist[3])} voices")
        print(f"  Duration:  {n / BASE_SR:.1f}s")
        print(f"  Volume:    {volume:.3f} (audio taper: {volume * volume:.3f})")
        print(f"  Speaker:   {speaker_char:.1f}")
        if has_flag(args, "--no-poweramp"):
            print("  Power amp: BYPASSED")
        print(f"  Peak:      {to_dbfs(peak):.1f} dBFS")
        print(f"  Output:    {output_path}")
        return 0
    job = api.engine_job(smf.engine_events(events, 64, BASE_SR), sample_rate=BASE_SR, duration=n / BASE_SR + 0.5 / BASE_SR, volume=volume,
                         tremolo_depth=depth, speaker_character=speaker_char, block_size=64, warm_up=True)
    out = api.render_engines([job], preamp_model=model)[0][:n].astype(np.float64)
    peak = float(np.max(np.abs(out))) if out.size else 0.0
    if peak > 1.0:
        sys.stderr.write(f"WARNING: Peak exceeds 0 dBFS ({to_dbfs(peak):.1f} dBFS) — consider reducing --volume\n")
    wav.write_preamp_bench_wav(output_path, out, BASE_SR, 1.0)
    d = api.last_diag()
    print("MIDI render complete")
    print(f"  File:      {midi_path}")
    print(f"  Notes:     {int(d.nr_iter_hist[0])} note-ons")
    print(f"  Peak poly: {int(d.nr_iter_hist[3])} voices")
    print(f"  Duration:  {n / BASE_SR:.1f}s")
    print(f"  Volume:    {volume:.3f}")
    print(f"  Speaker:   {speaker_char:.1f}")
    print(f"  Peak:      {to_dbfs(peak):.1f} dBFS")
    print(f"  Output:    {output_path}")
    return 0


USAGE = """Usage: preamp_bench <render|calibrate|sensitivity|render-midi|render-poly|alias-audit> [flags]
  alias-audit --note N | --notes a,b,c  --velocity V  --json
  sensitivity --notes a,b --velocities x,y --ds-range d1,d2 --scale-mode track|freeze|zero-trim --volume X --speaker C --output FILE
  render-poly --notes a,b,c --velocities x,y,z --duration S --volume X --speaker C --ldr OHM --no-poweramp --normalize --output FILE
  render-midi --midi FILE --output FILE --volume X --speaker C --tail S --track N --tremolo-depth D --preamp-model M
  render     --note N --velocity V --duration S --ldr OHM --volume X --speaker C --tremolo-depth D --sample-rate HZ
             --no-poweramp --no-rail-sag --power-amp behavioral|melange --no-preamp --no-attack-noise --no-mlp --normalize --displacement-scale DS --output FILE
             --preamp-model melange12|legacy8   (compile-time cargo feature in the reference)
  calibrate  --notes a,b,c --velocities x,y,z --ds-at-c4 D --ds-clamp-max M --volume X --speaker C --zero-trim --output FILE
"""


def alias_audit_results(notes, velocity=120):
    """alias_audit::run_with_note for several notes in ONE device batch (alias_audit.rs:106-204): render_stimulus as engine jobs
    (no warm-up, six 1024-sample settle blocks, note-on, 1.5 s at 44.1 kHz, volume 0.5, depth 0, character 0, MLP on), analysed on the device."""
    sr, total, settle = 44100.0, int(44100.0 * 1.5), 6 * 1024
    vel = float(np.float32(velocity) / np.float32(127.0))   # `velocity as f32 / 127.0`
    jobs = [api.engine_job([(settle, api.NOTE_ON, int(n), vel)], sample_rate=sr, duration=(settle + total + 0.5) / sr, volume=0.5, tremolo_depth=0.0,
                          speaker_character=0.0, mlp=True, block_size=1024, warm_up=False) for n in notes]
    rows = api.render_engines_alias(jobs, [api.note_hz(int(n)) for n in notes], analyze_seconds=0.5)
    return [api.AliasAuditResult(r) for r in rows]


def alias_audit_text(note, velocity, r, want_json):
    """The two output formats of cmd_alias_audit (main.rs:985-1066)."""
    if want_json:
        return ("{\n" + f'  "f0_hz": {r.f0_hz:.4f},\n  "h1_dbfs": {r.h1_dbfs:.3f},\n  "harmonic_dbc": [' + ", ".join(f"{v:.3f}" for v in r.harmonic_dbc) + "],\n"
                + f'  "max_step_up_db": {r.max_step_up_db:.3f},\n  "max_step_up_from_harmonic": {r.max_step_up_from_harmonic},\n  "hf_band_dbc": {r.hf_band_dbc:.3f}\n' + "}")
    lines = ["Click-band alias audit", f"  Stimulus:   note={note} vel={velocity} vol={0.5:.2f}",
             f"  Render:     {1.5:.2f}s @ {44100.0:.0f} Hz, analyzing last {0.5:.2f}s", "",
             f"  Detected f0:  {r.f0_hz:.3f} Hz  (nominal {api.note_hz(note):.3f} Hz)", f"  H1:           {r.h1_dbfs:.2f} dBFS", "",
             "  Harmonic envelope (dBc relative to H1):"]
    for i, dbc in enumerate(r.harmonic_dbc):
        marker = " *" if 6 <= i + 1 and i < 11 else "  "
        lines.append(f"    H{i + 1:<2} {marker} {dbc:>8.2f} dBc")
    lines += ["                  (* = harmonics in plateau-detection band)", "",
              f"  max_step_up_db:  {r.max_step_up_db:+.2f} dB  (worst rise: H{r.max_step_up_from_harmonic} \u2192 H{r.max_step_up_from_harmonic + 1})",
              "                   target: \u2264 0 dB (monotonic descent); gate: \u2264 +1.0",
              f"  hf_band_dbc:     {r.hf_band_dbc:.2f} dBc  (5\u201318 kHz RMS rel. to H1)",
              "                   target: track baseline; gate: \u2264 baseline + 2.0 dB"]
    return "\n".join(lines)


def cmd_alias_audit(args):
    """`preamp-bench alias-audit [--note N] [--velocity V] [--json]` (main.rs:985-1066); `--notes a,b,c` audits several notes in one batch."""
    want_json = has_flag(args, "--json")
    velocity = _as_u8(parse_flag(args, "--velocity", 120.0))
    notes = _csv_u8_list(args, "--notes", "") or [_as_u8(parse_flag(args, "--note", 84.0))]
    for n, r in zip(notes, alias_audit_results(notes, velocity)):
        print(alias_audit_text(n, velocity, r, want_json))
    return 0


def main(argv=None):
    args = list(sys.argv[1:] if argv is None else argv)
    if not args:
        sys.stderr.write(USAGE)
        return 1
    if args[0] == "render":
        return cmd_render(args[1:])
    if args[0] == "calibrate":
        return cmd_calibrate(args[1:])
    if args[0] == "render-midi":
        return cmd_render_midi(args[1:])
    if args[0] == "render-poly":
        return cmd_render_poly(args[1:])
    if args[0] == "sensitivity":
        return cmd_sensitivity(args[1:])
    if args[0] == "alias-audit":
        return cmd_alias_audit(args[1:])
    sys.stderr.write(f"Unknown subcommand: {args[0]} (only the batched render paths are mirrored)\n{USAGE}")
    return 1


if __name__ == "__main__":
    sys.exit(main())
