"""CPU checks of the kernel LOGIC: the device functions of libgooey_b200/csrc compiled for the host (tests/emu) and
driven exactly as the kernels drive them, against the oracle.  On the host both sides go through the same libm, so
the comparison is bit-exact; the GPU parity tests (test_voices_gpu.py) allow the 1e-5 of BASELINE.json."""
import numpy as np
import pytest

from libgooey_b200 import voices as V
import emu_lib as E
import oracle_lib as O
from workloads import drum_sweep_patches


@pytest.mark.parametrize("exact_tier", [True, False])
@pytest.mark.parametrize("mode", [0, 1])
def test_drum_sweep_bit_exact(exact_tier, mode):
    patches, vel, kinds = drum_sweep_patches(48, seed=0x600E7, exact_tier=exact_tier)
    trig = [(i, 0, float(vel[i])) for i in range(len(patches))]
    want = O.render_voices(patches, 16384, triggers=trig, threads=8)
    got, fast = E.render_voices(patches, 16384, triggers=trig, mode=mode)
    assert fast.sum() == (len(patches) if mode else 0)
    assert np.array_equal(got, want)


def _kit():
    return [V.patch(V.KICK, V.KICK_PRESETS["punch"]), V.patch(V.SNARE, V.SNARE_PRESETS["loose"]),
            V.patch(V.HIHAT, V.HIHAT_PRESETS["loose"]), V.patch(V.TOM, V.TOM_PRESETS["ring"], aux=1),
            V.patch(V.KICK, V.KICK_PRESETS["dirt"]), V.patch(V.SNARE, V.SNARE_PRESETS["hiss"]),
            V.patch(V.HIHAT, V.HIHAT_PRESETS["soft"]), V.patch(V.TOM, V.TOM_PRESETS["void"], aux=1)]


@pytest.mark.parametrize("n_calls", [1, 3, 7])
def test_retrigger_snapped_edits_split_path(n_calls):
    patches = _kit()
    n = len(patches)
    trig = [(i, 0, 0.5 + 0.05 * i) for i in range(n)]
    for i in range(n):
        for f in (5513, 11026, 30000, 30001, 52000):
            trig.append((i, f, 0.6))
    params = [(i, 9000, 4 if i % 4 == 0 else 1, 0.5, True) for i in range(n)] + [(3, 9000, 6, 0.9, True), (7, 12000, 5, 0.0, True)]
    want = O.render_voices(patches, 60000, triggers=trig, params=params)
    got, fast = E.render_voices(patches, 60000, triggers=trig, params=params, mode=1, n_calls=n_calls)
    assert (fast == n_calls).all()          # every call planned: parameters were snapped
    assert np.array_equal(got, want)


def test_gliding_parameters_fall_back_to_general_path_and_recover():
    patches = _kit()
    n = len(patches)
    trig = [(i, 0, 1.0) for i in range(n)] + [(i, 40000, 0.7) for i in range(n)]
    params = [(i, 3000, 0, 0.8, False) for i in range(n)]   # unsnapped edit: the smoother glides for ~150 ms
    want = O.render_voices(patches, 60000, triggers=trig, params=params)
    got, fast = E.render_voices(patches, 60000, triggers=trig, params=params, mode=1, n_calls=4)
    smoothed = np.array([p.instrument != V.TOM for p in patches])
    assert (fast[smoothed] == 3).all()      # the call containing the edit runs the general path, the others split
    assert (fast[~smoothed] == 4).all()     # Tom2 parameters are not smoothed
    assert np.array_equal(got, want)


def test_division_through_the_hoisted_reciprocal_equals_the_ieee_quotient():
    """gm::g_div_by (Markstein correction) as the additive oscillators and the pink-noise scale use it: bit-identical to `/`
    (host build of the same function; tests/test_gmath_gpu.py runs it on the device, with every 24-bit numerator for the noise scale)."""
    rng = np.random.default_rng(11)
    for b in (44100.0, 48000.0, 22050.0, 96000.0, 16777215.0):
        a = np.concatenate([rng.uniform(0, 1e3, 200000), rng.uniform(0, 3e9, 200000), np.exp(rng.uniform(-12, 30, 200000)),
                            np.arange(0, 1 << 24, 37), [0.0]]).astype(np.float32)
        got = E.math(0, a, np.full_like(a, b))
        assert np.array_equal(got.view(np.uint32), (a / np.float32(b)).astype(np.float32).view(np.uint32)), b


def test_front_end_sine_and_additive_triangle_stay_inside_their_error_budget():
    rng = np.random.default_rng(12)
    x = np.concatenate([rng.uniform(-7, 7, 200000), rng.uniform(0, 4e5, 200000), rng.uniform(0, 3e7, 200000)]).astype(np.float32)
    err = np.abs(E.math(1, x).astype(np.float64) - np.sin(x.astype(np.float64))).max()
    assert err <= 2.4e-7, err
    # the additive triangle (oscillator.rs:106-131) over two seconds of sample indices at drum pitches: fast against exact sine
    idx = rng.integers(0, 88200, 20000).astype(np.float32)
    freq = rng.uniform(30.0, 400.0, 20000).astype(np.float32)
    d = np.abs(E.math(2, idx, freq).astype(np.float64) - E.math(3, idx, freq).astype(np.float64)).max()
    print(f"g_sinf_fast err {err:.2e}; additive triangle, fast vs exact sine: {d:.2e}")
    assert d <= 6e-7, d


def test_bass_per_sample_tick_bit_exact():
    """bass_tick / bass_event (voices2.cuh: the tick bass_wave_kernel falls back to block by block, and the reference order its
    time-parallel blocks restate) against the oracle's BassSynth, host build: presets, retriggers, gliding and snapped edits."""
    patches = [V.patch(V.BASS, V.BASS_PRESETS[k]) for k in ("acid", "sub", "reese", "stab")] + [V.patch(V.BASS, V.BASS_PRESETS["reese"], tuning=0.7)]
    n = len(patches)
    trig = [(i, 0, 0.6 + 0.1 * i) for i in range(n)] + [(i, f, 0.8) for i in range(n) for f in (5513, 16539, 30000)]
    params = [(i, 9000, 6, 0.35, False) for i in range(n)] + [(i, 20000, 7, 0.9, True) for i in range(n)] + [(1, 12000, 13, 0.4, False), (3, 100, 2, 0.1, True)]
    want = O.render_voices(patches, 40000, triggers=trig, params=params)
    got, _ = E.render_voices(patches, 40000, triggers=trig, params=params, mode=0)
    assert np.abs(want).max() > 0.05
    assert np.array_equal(got, want)


@pytest.mark.parametrize("preset", [0, 1, 2, 3, 4])
def test_poly_synth_tick_bit_exact(preset):
    """poly_event / poly_tick (voices2.cuh, what slow_kernel<PolyV> runs) against the oracle's PolySynth, host build: chords, voice stealing
    (a seventh and eighth note), a parameter glide, release_all and a re-trigger during the release."""
    ev = [(0, 0, 60, 1.0), (0, 0, 64, 0.8), (0, 0, 67, 0.9), (3000, 0, 72, 0.7), (3000, 0, 76, 0.6), (3000, 0, 79, 0.5), (6000, 0, 84, 1.0), (6001, 0, 48, 0.9),
          (8000, 2, 2, 0.9), (8000, 2, 0, 0.7), (12000, 1, 0, 0.0), (15000, 0, 55, 1.0), (20000, 2, 13, 0.2)]
    want = E.poly_render(O.lib().orc_poly_render, preset, ev, 30000)
    got = E.poly_render(E.lib().emu_poly_render, preset, ev, 30000)
    assert np.abs(want).max() > 0.01
    assert np.array_equal(got, want)


def test_granulator_tick_bit_exact():
    """gran_event / gran_tick (voices2.cuh: the per-sample granulator path and the functions gran_wave_kernel's control lane runs) against the
    oracle's Granulator, host build: dense saturating cloud (slot stealing), random timing / amplitude, spray, reverse direction, glides."""
    rng = np.random.default_rng(5)
    n = 3 * 44100
    buf = (0.5 * np.sin(2 * np.pi * 220.0 * np.arange(n) / 44100.0) * (0.5 + 0.5 * np.sin(2 * np.pi * 0.7 * np.arange(n) / 44100.0)) + 0.1 * rng.uniform(-1, 1, n)).astype(np.float32)
    ev = [(0, 4, 7, 0.0), (0, 2, 4, 1.0), (0, 2, 1, 0.55), (0, 2, 2, 0.5), (0, 2, 3, 0.62), (0, 2, 6, 0.3), (0, 2, 5, 0.8), (0, 2, 9, 0.3), (0, 2, 10, 0.3), (0, 2, 7, 0.4),
          (0, 3, 0, 0.0), (0, 0, 0, 0.9), (9000, 2, 3, 0.4), (9000, 2, 11, 0.6), (20000, 2, 6, 0.9), (26000, 0, 0, 0.6)]
    want = E.gran_render(O.lib().orc_gran_render, buf, ev, 40000)
    got = E.gran_render(E.lib().emu_gran_render, buf, ev, 40000)
    assert np.abs(want).max() > 0.01
    assert np.array_equal(got, want)
