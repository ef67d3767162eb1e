"""The hand-written oracle against a MECHANICAL transliteration of the reference's generated solvers.

tools/transliterate_gen.py parses the Rust subset the melange code generator emits (gen_preamp.rs, gen_tremolo.rs) and prints the
same program as C++ -- no line restated by hand.  oracle/translit_check.cpp then drives both that transliteration and the hand
restatement the GPU parity tests are checked against (oracle/ow_preamp.hpp, ow_tremolo.hpp) with identical random states and long
trajectories and demands bit-identical outputs, next states and rebuilt matrices, with every guard path (BE fallback, voltage
damping, NaN reset, max-iteration exit, pnjlim's logarithm, per-sample matrix rebuilds) exercised.

The transliteration is generated from /root/reference (present in the build container, absent on the GPU box); the generated
headers live in oracle/_ref/ (git-ignored, never committed).  Without the reference and without prebuilt files the test skips."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/crates/openwurli-dsp/src"
EXE = os.path.join(ROOT, "oracle", "_ref", "translit_check")


def _build():
    if os.path.isdir(REF):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(EXE)


def test_hand_restatement_is_bit_identical_to_the_transliterated_reference_solvers():
    if not _build():
        pytest.skip("reference sources absent and no prebuilt oracle/_ref/translit_check")
    p = subprocess.run([EXE, "25000", "100000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:]
    r = json.loads(p.stdout.strip().splitlines()[-1])
    assert r["preamp_random_mismatches"] == 0 and r["preamp_mismatches"] == 0 and r["tremolo_mismatches"] == 0
    # the comparison is only worth something if the guard paths actually fired in the reference's own code
    for key in ("preamp_be_fallback", "preamp_voltage_damp", "preamp_nan_reset", "preamp_nr_max_iter", "preamp_pnjlim_ln", "preamp_rebuilds",
                "tremolo_be_fallback", "tremolo_nr_max_iter"):
        assert r[key] > 0, key


def test_transliterator_covers_the_whole_generated_files(tmp_path):
    """Every function of the three generated solvers is inside the tool's Rust subset except the stderr dump helper and the
    clock-seeded noise generators (noise is off in every parity run); gen_power_amp.rs (the melange power amplifier) too."""
    if not os.path.isdir(REF):
        pytest.skip("reference sources absent")
    for name in ("gen_preamp", "gen_tremolo", "gen_power_amp"):
        out = str(tmp_path / (name + ".hpp"))  # not oracle/_ref: rewriting those would make the built oracle look stale to make
        p = subprocess.run(["python3", os.path.join(ROOT, "tools", "transliterate_gen.py"), os.path.join(REF, name + ".rs"), name, out],
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        skipped = p.stderr.split("noise-only):")[-1].strip() if "noise-only" in p.stderr else ""
        assert skipped in ("", "dc_op_dump"), skipped
