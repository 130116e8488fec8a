"""More of the reference's integration tests restated against the oracle (see tests/test_reference_suite_cpu.py): gain staging,
stereo output, granulator through the FFI.  Kept in a module of its own: the GPU run of the first module
(tests/test_zz_reference_suite_gpu.py) was verified on a B200 as it stands."""
import numpy as np
import pytest

import oracle_lib as O
from test_reference_suite_cpu import KICK, SNARE, HIHAT, TOM, BASS, FX_LIMITER, engine  # noqa: F401  (fixture)

SR = 44100.0
RENDER_FRAMES, SETTLE_FRAMES = 4096, 16384


def render_triggered(e, instruments):
    for i in instruments:
        e.trigger_instrument(i)
    return e.render(RENDER_FRAMES)


# ------------------------------------------------------------------------------------------------ tests/ffi_gain_staging.rs
def test_default_output_is_a_linear_sum_without_hidden_nonlinear_processing(engine):   # ffi_gain_staging.rs:91-121
    kick = render_triggered(engine(), [KICK])
    tom = render_triggered(engine(), [TOM])
    combined = render_triggered(engine(), [KICK, TOM])
    assert float(np.abs(combined).max()) > 0.01
    assert float(np.abs(combined - (kick + tom)).max()) < 1e-5


def test_master_gain_scales_the_dry_sum(engine):                 # ffi_gain_staging.rs:123-153
    quarter, half = engine(), engine()
    half.set_master_gain(0.5)
    quarter.render(SETTLE_FRAMES); half.render(SETTLE_FRAMES)
    q, h = render_triggered(quarter, [KICK]), render_triggered(half, [KICK])
    assert float(np.abs(q).max()) > 0.01
    assert float(np.abs(h - q * np.float32(2.0)).max()) < 5e-4


def test_master_gain_is_applied_before_the_optional_limiter(engine):   # ffi_gain_staging.rs:155-178
    dry_e, lim_e = engine(), engine()
    lim_e.set_global_effect_enabled(FX_LIMITER, True)
    dry, limited = render_triggered(dry_e, [KICK, TOM]), render_triggered(lim_e, [KICK, TOM])
    assert float(np.abs(limited - np.tanh(dry.astype(np.float64)).astype(np.float32)).max()) < 1e-6


def test_master_gain_clamps_and_ignores_non_finite_values(engine):     # ffi_gain_staging.rs:68-89, observed through the audio
    a, b, c = engine(), engine(), engine()
    a.set_master_gain(0.75)
    b.set_master_gain(0.75); b.set_master_gain(float("nan")); b.set_master_gain(float("inf"))
    c.set_master_gain(3.0)                                       # clamps to 2.0
    for e in (a, b, c):
        e.render(SETTLE_FRAMES)
    ra, rb, rc = (render_triggered(e, [KICK]) for e in (a, b, c))
    assert np.array_equal(ra, rb)
    d = engine(); d.set_master_gain(2.0); d.render(SETTLE_FRAMES)
    assert np.array_equal(rc, render_triggered(d, [KICK]))


def test_offline_bounce_snaps_to_the_configured_master_gain(engine):   # ffi_gain_staging.rs:180-214
    quarter, half = engine(), engine()
    half.set_master_gain(0.5)
    quarter.sequencer_set_step(0, True); half.sequencer_set_step(0, True)
    q, h = quarter.bounce_to_buffer(1), half.bounce_to_buffer(1)
    assert len(q) == len(h) and float(np.abs(q).max()) > 0.01
    assert float(np.abs(h - q * np.float32(2.0)).max()) < 1e-6


# ------------------------------------------------------------------------------------------------ tests/ffi_stereo.rs
def test_render_fills_an_interleaved_stereo_buffer(engine):      # ffi_stereo.rs:20-44
    e = engine()
    e.trigger_instrument(KICK)
    buf = e.render(512)
    assert buf.shape == (512, 2) and np.isfinite(buf).all()


def test_left_and_right_match_and_both_carry_audio_for_the_mono_path(engine):   # ffi_stereo.rs:46-83
    e = engine()
    e.trigger_instrument(KICK)
    buf = e.render(2048)
    assert float(np.abs(buf[:, 0]).max()) > 0.001 and float(np.abs(buf[:, 1]).max()) > 0.001
    assert np.array_equal(buf[:, 0], buf[:, 1])


def test_hard_left_pan_steers_audio_to_the_left_channel(engine):  # ffi_stereo.rs:114-146
    e = engine()
    e.set_instrument_pan(KICK, 0.0)
    e.render(1024)
    e.trigger_instrument(KICK)
    buf = e.render(2048).astype(np.float64)
    left, right = float((buf[:, 0] ** 2).sum()), float((buf[:, 1] ** 2).sum())
    assert left > 0.0 and right < left * 1e-3


def test_offline_bounce_downmixes_panned_audio_to_continuous_mono(engine):   # ffi_stereo.rs:148-183
    e = engine()
    e.sequencer_set_step(0, True)
    e.set_instrument_pan(KICK, 0.0)
    out = e.bounce_to_buffer(1).astype(np.float64)
    assert float(np.abs(out).max()) > 0.001
    even, odd = float((out[0::2] ** 2).sum()), float((out[1::2] ** 2).sum())
    assert odd > even * 0.25


# ------------------------------------------------------------------------------------------------ tests/ffi_granulator.rs
def sine_samples(seconds, hz):
    n = int(SR * seconds)
    return (np.sin(np.arange(n, dtype=np.float32) / np.float32(SR) * np.float32(hz) * np.float32(2 * np.pi)) * 0.5).astype(np.float32)


def test_granulator_is_silent_until_a_buffer_is_set_and_triggered(engine):      # ffi_granulator.rs:25-37: the silent placeholder buffer
    e = engine()
    e.granulator_trigger(1.0)
    assert float(np.abs(e.render(8192)).max()) < 1e-6


def test_granulator_set_buffer_rejects_invalid_inputs(engine):   # ffi_granulator.rs:59-88
    e = engine()
    assert not e.granulator_set_buffer(np.zeros(0, np.float32), SR)
    assert not e.granulator_set_buffer(sine_samples(0.1, 220.0), 0.0)
    assert not e.granulator_set_buffer(sine_samples(0.1, 220.0), float("nan"))
    assert e.granulator_set_buffer(sine_samples(0.1, 220.0), SR)


def test_granulator_trigger_and_render_produces_finite_nonzero_audio(engine):   # ffi_granulator.rs:171-202
    e = engine()
    assert e.granulator_set_buffer(sine_samples(1.0, 220.0), SR)
    e.granulator_set_seed(7)
    for p, v in [(8, 1.0), (4, 0.5), (7, 0.5)]:                 # GRANULATOR_PARAM_VOLUME, _DENSITY, _CLOUD_DURATION
        e.granulator_set_param(p, v)
    e.granulator_snap_params()
    e.granulator_trigger(1.0)
    buf = e.render(22050)
    assert np.isfinite(buf).all() and float(np.abs(buf).max()) > 1e-4
