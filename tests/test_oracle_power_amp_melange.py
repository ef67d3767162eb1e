"""SURVEY 8(f) #4, oracle side: the reference's default power amplifier (melange 7-BJT Class AB solver + behavioural rail sag,
crates/openwurli-dsp/src/power_amp.rs) restated as RailDynamics + adapter by hand over the MECHANICALLY transliterated gen_power_amp.rs
(oracle/_ref/gen_power_amp.hpp, see tests/test_transliterated_solvers.py), pinned by the reference's own unit tests (power_amp.rs:470-800).
The device side of this row is not built."""
import numpy as np
import pytest

import oracle_lib as O

L = O.lib()
SR = 44100.0
needs_solver = pytest.mark.skipif(not L.owo_have_melange_power_amp(), reason="oracle built without oracle/_ref/gen_power_amp.hpp (reference sources absent)")


def _rails(v_out):
    v = np.ascontiguousarray(v_out, dtype=np.float64)
    r = np.zeros((len(v), 2))
    assert L.owo_rail_dynamics(SR, O.dptr(v), len(v), O.dptr(r)) == 0
    return r


def _amp(x, rail_sag=True, toggle_off_at=-1):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y, r = np.zeros(len(x)), np.zeros((len(x), 2))
    resets = L.owo_power_amp_melange(SR, 1 if rail_sag else 0, O.dptr(x), len(x), O.dptr(y), O.dptr(r), toggle_off_at)
    assert resets >= 0
    return y, r, resets


def _sine(amp, freq, n, start=0):
    return amp * np.sin(2 * np.pi * freq * (start + np.arange(n)) / SR)


def test_rail_dynamics_unit_and_offsets():  # power_amp.rs:740-800 (no solver needed)
    r = _rails(np.zeros(int(SR) // 4))
    assert abs(r[0, 0] - 22.5) < 0.05 and abs(r[-1, 0] - 24.5) < 0.05 and abs(r[-1, 1] - 24.5) < 0.05
    r2 = _rails(np.concatenate([np.zeros(int(SR) // 4), np.full(int(SR * 0.3), 8.0)]))
    assert abs(r2[-1, 0] - 21.0) < 0.1          # 1 A on the positive rail: 24.5 - 3.5 V
    assert abs(r2[-1, 1] - 24.5) < 0.05         # the negative rail is untouched
    off = r2[-1] - 22.5
    assert -2.0 < off[0] < -1.0 and abs(off[1] - 2.0) < 0.05
    # asymmetric one-pole: the sag (8 ms) is faster than the recovery (15 ms)
    r3 = _rails(np.concatenate([np.full(int(SR * 0.5), 8.0), np.zeros(int(SR * 0.5))]))
    k = int(SR * 0.5)
    assert r3[k - 1, 0] < 21.2 and r3[-1, 0] > 24.4


@needs_solver
def test_closed_loop_gain_clipping_and_bounds():  # :493-560
    settle, measure = int(SR * 0.3), int(SR * 0.1)
    y, _, resets = _amp(_sine(0.001, 1000.0, settle + measure))
    gain_db = 20 * np.log10(np.abs(y[settle:]).max() / 0.001)
    assert 5.0 < gain_db < 20.0 and resets == 0, gain_db      # 69x / 22 V headroom = 3.14x = 9.9 dB
    y, _, _ = _amp(_sine(5.0, 100.0, int(SR * 0.2)))
    peak = np.abs(y[int(SR * 0.1) + 1:]).max()
    assert 0.85 < peak <= 1.0, peak
    for v in (0.0, 0.001, 0.01, 0.1, 0.5, 1.0, 5.0, -0.1, -1.0, -5.0):
        y, _, _ = _amp(np.full(101, v))
        assert np.isfinite(y[-1]) and abs(y[-1]) <= 1.0


@needs_solver
def test_crossover_reduced_by_feedback():  # :523-545
    n = int(SR * 0.3)
    y, _, _ = _amp(_sine(0.001, 440.0, n))
    s = y[n // 2 + 1:]
    i = np.arange(len(s))
    mag = lambda f: np.hypot(np.sum(s * np.cos(2 * np.pi * f * i / SR)), np.sum(s * np.sin(2 * np.pi * f * i / SR))) / len(s)
    assert 20 * np.log10(mag(3 * 440.0) / mag(440.0)) < -30.0


@needs_solver
def test_rail_sag_idle_load_recovery_and_toggle():  # :598-738
    _, r, _ = _amp(np.zeros(100), rail_sag=False)
    assert np.abs(r - 22.5).max() < 1e-9                                  # static bias when off
    _, r, _ = _amp(np.zeros(int(SR) // 4))
    assert abs(r[-1, 0] - 24.5) < 0.05 and abs(r[-1, 1] - 24.5) < 0.05    # idle rails
    idle = int(SR) // 10
    _, r, _ = _amp(np.concatenate([np.zeros(idle), _sine(0.20, 220.0, int(SR * 0.5))]))
    vp_idle, (vp_l, vn_l) = r[idle - 1, 0], r[-1]
    assert vp_l < vp_idle - 0.1 and vn_l < vp_idle - 0.1 and vp_l > 20.0 and vn_l > 20.0
    k = int(SR * 0.2)
    _, r, _ = _amp(np.concatenate([_sine(0.3, 110.0, k), np.zeros(k)]))
    assert r[k - 1, 0] < 24.0 and r[-1, 0] > r[k - 1, 0] + 0.5 and abs(r[-1, 0] - 24.5) < 0.05
    m = int(SR * 0.05)
    _, r, _ = _amp(np.concatenate([_sine(0.5, 220.0, m), [0.0]]), toggle_off_at=m)
    assert np.abs(r[m] - 22.5).max() < 1e-9 and r[m - 1, 0] != 22.5       # toggling off zeroes the offsets at once
