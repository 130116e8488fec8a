"""The reference's inline unit tests of src/effects/*.rs (`#[cfg(test)] mod tests`), restated against the oracle's effect classes
(oracle/effects.hpp, oracle/prims.hpp) through `orc_fx_*`: each test cites the Rust test it follows and asserts what it asserts.
Getter-only tests (parameter clamping read back through `get_*`) and the oversampling-mode setters (no FFI for them on this path)
are left out.  f32 inputs are formed the way the Rust tests form them (`(i as f32 * 0.1).sin()` -> numpy float32 arithmetic)."""
import ctypes as c

import numpy as np
import pytest

import oracle_lib as O

LOWPASS, DELAY, SATURATION, COMPRESSOR, TILT, LIMITER, SPRING, WAVESHAPER, FBWS, PLATE = range(10)
SR = 44100.0
f32 = np.float32


class Fx:
    def __init__(self, kind, *ctor, sr=SR):
        L = O.lib()
        L.orc_fx_new.restype = c.c_void_p
        L.orc_fx_new.argtypes = [c.c_uint32, c.c_float, c.c_void_p]
        for name, args in [("orc_fx_free", [c.c_void_p]), ("orc_fx_set_param", [c.c_void_p, c.c_uint32, c.c_float]), ("orc_fx_set_bpm", [c.c_void_p, c.c_float]),
                           ("orc_fx_reset", [c.c_void_p]), ("orc_fx_process", [c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32]),
                           ("orc_fx_process_stereo", [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32])]:
            getattr(L, name).argtypes = args
            getattr(L, name).restype = None
        self.L = L
        arr = np.array(list(ctor) + [0.0], np.float32)
        self.h = L.orc_fx_new(kind, c.c_float(sr), arr.ctypes.data)
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_fx_free(self.h)
            self.h = None

    def set(self, p, v):
        self.L.orc_fx_set_param(self.h, p, c.c_float(v))

    def set_bpm(self, bpm):
        self.L.orc_fx_set_bpm(self.h, c.c_float(bpm))

    def reset(self):
        self.L.orc_fx_reset(self.h)

    def run(self, x):
        x = np.ascontiguousarray(np.atleast_1d(x), np.float32)
        out = np.empty_like(x)
        self.L.orc_fx_process(self.h, x.ctypes.data, out.ctypes.data, x.size)
        return out

    def one(self, x):
        return float(self.run([x])[0])

    def stereo(self, l, r):
        l = np.ascontiguousarray(l, np.float32); r = np.ascontiguousarray(r, np.float32)
        ol, orr = np.empty_like(l), np.empty_like(r)
        self.L.orc_fx_process_stereo(self.h, l.ctypes.data, r.ctypes.data, ol.ctypes.data, orr.ctypes.data, l.size)
        return ol, orr


def sin_i(n, k, amp=1.0):
    """(i as f32 * k).sin() * amp, in f32"""
    return (np.sin((np.arange(n, dtype=np.float32) * f32(k)).astype(np.float32)).astype(np.float32) * f32(amp)).astype(np.float32)


def tone(n, freq, sr=SR, amp=1.0):
    i = np.arange(n, dtype=np.float32)
    return (np.sin((f32(2.0) * f32(np.pi) * f32(freq) * i / f32(sr)).astype(np.float32)).astype(np.float32) * f32(amp)).astype(np.float32)


def impulse(n, width=100, level=1.0):
    x = np.zeros(n, np.float32)
    x[:width] = level
    return x


# ------------------------------------------------------------------------------------------------ effects/saturation.rs
def test_saturation_bypass_when_mix_zero():                      # saturation.rs test_bypass_when_mix_zero
    sat = Fx(SATURATION, 0.5, 0.5, 0.0)
    assert sat.one(0.5) == 0.5 and sat.one(-0.3) == f32(-0.3)


def test_saturation_soft_limiting():                             # test_soft_limiting
    out = Fx(SATURATION, 1.0, 0.0, 1.0).run(tone(4000, 1000.0))
    peak = float(np.abs(out[2000:]).max())
    assert 0.1 < peak < 1.0


def test_saturation_dc_stability_and_nan_protection():           # test_dc_stability, test_nan_protection
    assert np.isfinite(Fx(SATURATION, 0.5, 1.0, 1.0).run(np.full(44100, 0.5, np.float32))).all()
    assert Fx(SATURATION, 0.5, 0.5, 0.5).one(float("nan")) == 0.0


def test_saturation_reset_matches_fresh_instance():              # test_reset_matches_fresh_instance, test_nan_protection_resets_state
    reset = Fx(SATURATION, 0.5, 0.5, 1.0)
    reset.run(sin_i(1000, 0.1))
    reset.reset()
    x = sin_i(100, 0.27)
    assert np.array_equal(reset.run(x), Fx(SATURATION, 0.5, 0.5, 1.0).run(x))
    nan_reset = Fx(SATURATION, 0.5, 0.5, 1.0)
    nan_reset.run(sin_i(1000, 0.1))
    assert nan_reset.one(float("nan")) == 0.0
    assert nan_reset.one(0.5) == Fx(SATURATION, 0.5, 0.5, 1.0).one(0.5)


# ------------------------------------------------------------------------------------------------ effects/lowpass_filter.rs
def test_lowpass_basic_nan_and_stability():                      # test_filter_basic_processing, _nan_protection, _stability_at_high_resonance
    assert np.isfinite(Fx(LOWPASS, 1000.0, 0.0).run(np.ones(1000, np.float32))).all()
    assert np.isfinite(Fx(LOWPASS, 1000.0, 0.5).one(float("nan")))
    out = Fx(LOWPASS, 500.0, 0.95).run(impulse(44100))
    assert np.isfinite(out).all() and float(np.abs(out).max()) < 100.0


def test_lowpass_high_frequency_stability():                     # test_filter_high_frequency_stability
    out = Fx(LOWPASS, 20000.0, 0.95).run(sin_i(44100, 0.1, 0.5))
    assert np.isfinite(out).all() and float(np.abs(out).max()) < 10.0


def test_lowpass_parameter_changes_and_full_sweep_stay_stable():  # test_filter_control_thread_safety, test_filter_sweep_full_range
    f = Fx(LOWPASS, 1000.0, 0.5)
    for i in range(1000):
        f.set(0, 200.0 + i * 10.0)
        f.set(1, i / 2000.0)
        assert np.isfinite(f.one(1.0))
    f = Fx(LOWPASS, 20.0, 0.7)
    x = sin_i(44100, 0.05)
    for i in range(0, 44100, 49):                                # the sweep in 900 steps (one setter call per 49 samples)
        f.set(0, 20.0 * 1000.0 ** (i / 44100.0))
        out = f.run(x[i:i + 49])
        assert np.isfinite(out).all() and float(np.abs(out).max()) < 10.0


# ------------------------------------------------------------------------------------------------ effects/tilt_filter.rs
def test_tilt_passthrough_at_center():                           # test_passthrough_at_center
    f = Fx(TILT)
    f.run(np.zeros(4410, np.float32))
    assert abs(f.one(0.7) - 0.7) < 0.001


def test_tilt_lowpass_attenuates_high_freq():                    # test_lowpass_attenuates_high_freq
    f = Fx(TILT)
    f.set(0, 0.0)
    f.run(np.zeros(4410, np.float32))
    i = np.arange(4410, dtype=np.float32)
    x = np.sin((f32(2.0) * f32(np.pi) * f32(10000.0) * (i / f32(44100.0))).astype(np.float32)).astype(np.float32)
    assert float(np.abs(f.run(x)).max()) < 0.1


def test_tilt_highpass_attenuates_low_freq():                    # test_highpass_attenuates_low_freq
    f = Fx(TILT)
    f.set(0, 1.0)
    f.run(np.zeros(4410, np.float32))
    i = np.arange(4410, dtype=np.float32)
    x = np.sin((f32(2.0) * f32(np.pi) * f32(100.0) * (i / f32(44100.0))).astype(np.float32)).astype(np.float32)
    assert float(np.abs(f.run(x)).max()) < 0.1


def test_tilt_nan_sweep_and_high_resonance_stability():          # test_nan_protection, test_stability_full_sweep, test_stability_high_resonance
    f = Fx(TILT)
    f.set(0, 0.0)
    f.run(np.full(1000, 0.5, np.float32))
    assert np.isfinite(f.one(float("nan")))
    f = Fx(TILT)
    x = sin_i(44100, 0.05)
    for i in range(0, 44100, 49):
        f.set(0, i / 44100.0); f.set(1, 0.8)
        out = f.run(x[i:i + 49])
        assert np.isfinite(out).all() and float(np.abs(out).max()) < 10.0
    f = Fx(TILT)
    f.set(0, 0.1); f.set(1, 1.0)
    out = f.run(impulse(44100))
    assert np.isfinite(out).all() and float(np.abs(out).max()) < 100.0


# ------------------------------------------------------------------------------------------------ effects/compressor.rs
def test_compressor_bypass_when_mix_zero():                      # test_bypass_when_mix_zero
    comp = Fx(COMPRESSOR, -12.0, 4.0, 5.0, 100.0, 0.0)
    assert comp.one(0.5) == 0.5 and comp.one(-0.3) == f32(-0.3)


def test_compressor_gain_reduction_above_threshold():            # test_gain_reduction_above_threshold
    comp = Fx(COMPRESSOR, -20.0, 10.0, 0.1, 50.0, 1.0)
    comp.run(np.tile(np.array([0.8, -0.8], np.float32), 4000))
    out = abs(comp.one(0.8))
    assert 0.01 < out < 0.8


def test_compressor_no_reduction_below_threshold():              # test_no_reduction_below_threshold
    comp = Fx(COMPRESSOR, -1.0, 4.0, 5.0, 100.0, 1.0)
    comp.run(np.zeros(2000, np.float32))
    assert abs(comp.one(0.01) - 0.01) < 0.05


def test_compressor_nan_dc_and_reset():                          # test_nan_protection, test_dc_stability, test_compressor_oversampling_reset_matches_fresh
    assert Fx(COMPRESSOR, -12.0, 4.0, 5.0, 100.0, 0.5).one(float("nan")) == 0.0
    assert np.isfinite(Fx(COMPRESSOR, -12.0, 4.0, 5.0, 100.0, 1.0).run(np.full(44100, 0.5, np.float32))).all()
    reset = Fx(COMPRESSOR, -20.0, 10.0, 0.1, 50.0, 1.0)
    reset.run(sin_i(2000, 0.3, 0.9))
    reset.reset()
    x = sin_i(100, 0.27, 0.9)
    assert np.array_equal(reset.run(x), Fx(COMPRESSOR, -20.0, 10.0, 0.1, 50.0, 1.0).run(x))


# ------------------------------------------------------------------------------------------------ effects/delay.rs
EIGHTH, QUARTER = 3, 2


def test_delay_basic_processing_and_nan():                       # test_delay_basic_processing, test_delay_nan_protection
    assert np.isfinite(Fx(DELAY, EIGHTH, 120.0, 0.5, 0.5, 10000.0).run(np.ones(1000, np.float32))).all()
    assert np.isfinite(Fx(DELAY, EIGHTH, 120.0, 0.5, 0.5, 10000.0).one(float("nan")))


def test_delay_feedback_stability():                             # test_delay_feedback_stability
    out = Fx(DELAY, EIGHTH, 120.0, 0.95, 0.5, 5000.0).run(impulse(44100))
    assert np.isfinite(out).all() and float(np.abs(out).max()) < 100.0


def test_delay_filter_darkens_echoes():                          # test_delay_filter_darkens_echoes
    d = Fx(DELAY, EIGHTH, 120.0, 0.9, 1.0, 2000.0)
    d.one(1.0)
    ds = int(f32(0.25) * f32(44100.0))
    out = np.abs(d.run(np.zeros(ds * 3 - 1, np.float32)))       # out[k] is iteration i = k + 1
    first = float(out[ds - 3:ds + 2].max())
    second = float(out[2 * ds - 3:2 * ds + 2].max())
    assert first > 0.01 and second < first


def test_delay_reset_and_bpm_change():                           # test_delay_reset, test_delay_bpm_change_updates_time
    d = Fx(DELAY, QUARTER, 120.0, 0.5, 0.5, 10000.0)
    d.run(np.ones(44100, np.float32))
    d.reset()
    assert abs(d.one(0.0)) < 0.001
    d = Fx(DELAY, QUARTER, 120.0, 0.0, 1.0, 20000.0)
    d.set_bpm(60.0)
    assert np.isfinite(d.run(np.zeros(100, np.float32))).all()


# ------------------------------------------------------------------------------------------------ effects/plate_reverb.rs
def plate_ir(p, frames):
    x = np.zeros(frames, np.float32); x[0] = 1.0
    return p.stereo(x, x)


def test_plate_stable_at_max_decay():                            # stable_at_max_decay
    n = int(SR * 5.0)
    l, r = plate_ir(Fx(PLATE, 1.0, 1.0, 0.0), n)
    assert np.isfinite(l).all() and np.isfinite(r).all()
    assert float(np.abs(l).max()) < 4.0 and float(np.abs(r).max()) < 4.0
    one = int(SR)
    first = max(float(np.abs(l[:one]).max()), float(np.abs(r[:one]).max()))
    last = max(float(np.abs(l[n - one:]).max()), float(np.abs(r[n - one:]).max()))
    assert last <= first * 1.5


def test_plate_decay_time_is_sane():                             # decay_time_is_sane
    l, r = plate_ir(Fx(PLATE, 0.5, 1.0, 0.0), int(SR * 5.0))
    w = int(SR * 0.1)
    rms = [float(np.sqrt(((l[k:k + w].astype(np.float64) ** 2) + (r[k:k + w].astype(np.float64) ** 2)).sum() / (2.0 * len(l[k:k + w])))) for k in range(0, len(l), w)]
    th = max(rms) * 0.001
    t60 = next((i * 0.1 for i, v in enumerate(rms) if v < th), None)
    assert t60 is not None and 0.3 <= t60 <= 4.0
    for a, b in zip(rms[2:], rms[3:]):
        assert b <= a * 1.2


def test_plate_decorrelates_and_zero_width_collapses_to_mono():  # decorrelates_left_and_right, zero_width_collapses_to_mono
    l, r = plate_ir(Fx(PLATE, 0.7, 1.0, 0.3), int(SR))
    assert float(np.abs(l - r).max()) > 1e-3
    p = Fx(PLATE, 0.7, 1.0, 0.3)
    p.set(4, 0.0)                                                # PLATE_PARAM_WIDTH
    x = np.zeros(13230 + 22050, np.float32)
    x[:13230][np.arange(13230) % 1000 == 0] = 1.0
    x[13230:][np.arange(22050) % 1000 == 0] = 1.0
    l, r = p.stereo(x, x)
    assert float(np.abs(l[13230:] - r[13230:]).max()) < 1e-6


def test_plate_nan_reset_and_common_sample_rates():              # nan_input_produces_finite_output, reset_clears_state, constructs_and_runs_at_common_sample_rates
    p = Fx(PLATE, 0.8, 1.0, 0.2)
    l, r = p.stereo(np.full(1000, np.nan, np.float32), np.full(1000, np.inf, np.float32))
    assert np.isfinite(l).all() and np.isfinite(r).all()
    l, r = p.stereo(np.array([0.5], np.float32), np.array([0.5], np.float32))
    assert np.isfinite(l).all() and np.isfinite(r).all()
    p = Fx(PLATE, 0.9, 1.0, 0.0)
    plate_ir(p, 8192)
    p.reset()
    l, r = p.stereo(np.zeros(8192, np.float32), np.zeros(8192, np.float32))
    assert not l.any() and not r.any()
    for sr in (22050.0, 44100.0, 48000.0, 96000.0):
        p = Fx(PLATE, 1.0, 1.0, 0.0, sr=sr)
        p.set(5, 1.0)                                            # PLATE_PARAM_SIZE: max tank scale exercises buffer headroom
        l, r = plate_ir(p, 4096)
        assert np.isfinite(l).all() and np.isfinite(r).all()


# ------------------------------------------------------------------------------------------------ effects/waveshaper.rs
def test_waveshaper_bypasses():                                  # test_bypass_when_mix_zero, test_bypass_when_drive_one, test_zero_input
    for ws in (Fx(WAVESHAPER, 5.0, 0.0), Fx(WAVESHAPER, 1.0, 1.0)):
        assert ws.one(0.5) == 0.5 and ws.one(-0.3) == f32(-0.3)
    ws = Fx(WAVESHAPER, 5.0, 1.0)
    ws.run(np.zeros(20, np.float32))
    assert ws.one(0.0) == 0.0


def test_waveshaper_soft_clipping_and_gain_compensation():       # test_soft_clipping, test_gain_compensation_consistency
    ws = Fx(WAVESHAPER, 10.0, 1.0)
    ws.run(np.ones(20, np.float32))
    assert 0.3 < ws.one(1.0) < 0.8
    lo, hi = Fx(WAVESHAPER, 2.0, 1.0), Fx(WAVESHAPER, 10.0, 1.0)
    lo.run(np.full(100, 0.5, np.float32)); hi.run(np.full(100, 0.5, np.float32))
    assert lo.one(0.5) > 0.1 and hi.one(0.5) > 0.1


def test_waveshaper_reset_matches_fresh_instance():              # test_reset_matches_fresh_instance, test_nan_protection_resets_state
    reset = Fx(WAVESHAPER, 5.0, 1.0)
    reset.run(sin_i(1000, 0.1))
    reset.reset()
    x = sin_i(100, 0.27)
    assert np.array_equal(reset.run(x), Fx(WAVESHAPER, 5.0, 1.0).run(x))
    nan_reset = Fx(WAVESHAPER, 5.0, 1.0)
    nan_reset.run(sin_i(1000, 0.1))
    assert nan_reset.one(float("nan")) == 0.0
    assert nan_reset.one(0.5) == Fx(WAVESHAPER, 5.0, 1.0).one(0.5)


# ------------------------------------------------------------------------------------------------ effects/feedback_waveshaper.rs
def test_fbws_bypasses_and_zero_input():                         # test_bypass_when_mix_zero, test_bypass_when_drive_one, test_zero_input
    for ws in (Fx(FBWS, 5.0, 0.5, 2000.0, 0.0), Fx(FBWS, 1.0, 0.5, 2000.0, 1.0)):
        assert ws.one(0.5) == 0.5 and ws.one(-0.3) == f32(-0.3)
    assert Fx(FBWS, 5.0, 0.5, 2000.0, 1.0).one(0.0) == 0.0


def test_fbws_soft_clipping():                                   # test_soft_clipping
    full = float(np.abs(Fx(FBWS, 10.0, 0.0, 2000.0, 1.0).run(sin_i(2000, 0.3))).max())
    quiet = float(np.abs(Fx(FBWS, 10.0, 0.0, 2000.0, 1.0).run(sin_i(2000, 0.3, 0.1))).max())
    assert 0.5 < full < 1.1 and quiet < full


def test_fbws_feedback_changes_output_and_stays_bounded():       # test_feedback_changes_output, test_feedback_does_not_blow_up, test_extreme_settings_stability
    x = sin_i(1000, 0.1, 0.5)
    assert float(np.abs(Fx(FBWS, 5.0, 0.7, 2000.0, 1.0).run(x) - Fx(FBWS, 5.0, 0.0, 2000.0, 1.0).run(x)).max()) > 0.01
    sq = np.where(np.arange(44100) % 100 < 50, 0.8, -0.8).astype(np.float32)
    for ws in (Fx(FBWS, 10.0, 0.9, 2000.0, 1.0), Fx(FBWS, 100.0, 0.98, 2000.0, 1.0)):
        out = ws.run(sq)
        assert np.isfinite(out).all() and float(np.abs(out).max()) < 5.0


def test_fbws_dc_decays_nan_and_reset():                         # test_dc_stability, test_nan_protection, test_reset_clears_state
    ws = Fx(FBWS, 5.0, 0.8, 1000.0, 1.0)
    ws.run(np.full(44100, 0.3, np.float32))
    assert abs(float(ws.run(np.zeros(4410, np.float32))[-1])) < 0.01
    ws = Fx(FBWS, 5.0, 0.5, 2000.0, 1.0)
    ws.one(0.5); ws.one(0.3)
    assert ws.one(float("nan")) == 0.0 and np.isfinite(ws.one(0.5))
    ws = Fx(FBWS, 5.0, 0.8, 2000.0, 1.0)
    ws.run(np.full(1000, 0.5, np.float32))
    ws.reset()
    assert ws.one(0.5) == Fx(FBWS, 5.0, 0.8, 2000.0, 1.0).one(0.5)


def test_fbws_feedback_gain_compensation_and_transient():        # test_feedback_gain_compensation, test_transient_not_compressed_below_tail
    x = sin_i(4000, 0.1, 0.5)
    no_fb = Fx(FBWS, 5.0, 0.0, 2000.0, 1.0).run(x).astype(np.float64)
    fb = Fx(FBWS, 100.0, 0.98, 2000.0, 1.0).run(x).astype(np.float64)
    ratio = float(np.sqrt((fb ** 2).mean() / (no_fb ** 2).mean()))
    assert 1.0 < ratio < 5.0
    n = int(0.2 * SR); aw = int(0.005 * SR)
    i = np.arange(n, dtype=np.float32)
    amp = ((f32(1.0) - i / f32(n)) * f32(0.9) + f32(0.1)).astype(np.float32)
    out = np.abs(Fx(FBWS, 10.0, 0.0, 2000.0, 1.0).run((np.sin((i * f32(0.2)).astype(np.float32)).astype(np.float32) * amp).astype(np.float32)))
    assert float(out[:aw].max()) > float(out[n - aw + 1:].max())


# ------------------------------------------------------------------------------------------------ effects/limiter.rs
def test_limiter_is_a_scaled_tanh():                             # limiter.rs:24-41 (process), as tests/ffi_gain_staging.rs uses it
    L = O.lib()
    L.orc_limiter.restype = c.c_float
    L.orc_limiter.argtypes = [c.c_float, c.c_float]
    for th in (1.0, 0.7, 0.25):
        for x in (-2.0, -0.3, 0.0, 0.1, 0.9, 5.0):
            want = f32(np.tanh(f32(f32(x) * f32(f32(1.0) / f32(th))))) * f32(th)
            assert abs(L.orc_limiter(th, x) - float(want)) < 1e-6


# ================================================================================================ src/filters/*.rs, src/gen/*.rs
def filt(kind, x, a, b=0.0, cc=0.0, mode=0, sr=SR):
    L = O.lib()
    L.orc_filter_run.argtypes = [c.c_uint32, c.c_float, c.c_float, c.c_float, c.c_float, c.c_uint32, c.c_void_p, c.c_void_p, c.c_uint32]
    L.orc_filter_run.restype = None
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    L.orc_filter_run(kind, sr, a, b, cc, mode, x.ctypes.data, out.ctypes.data, x.size)
    return out


def test_svf_bandpass_settles_on_dc_and_all_outputs_respond():   # state_variable.rs test_svf_bandpass_output, test_svf_all_outputs
    assert abs(float(filt(0, np.ones(1000, np.float32), 1000.0, 2.0, mode=1)[-1])) < 0.1
    first = [float(filt(0, np.ones(1, np.float32), 1000.0, 1.0, mode=m)[0]) for m in (0, 1, 2)]
    assert any(v != 0.0 for v in first)


def rlp_rms(sr, cutoff, q, freq):                                # resonant_lowpass.rs response_rms
    n = int(sr)
    out = filt(2, tone(n, freq, sr), cutoff, q, sr=sr).astype(np.float64)
    return float(np.sqrt((out[n // 2:] ** 2).sum() / (n // 2)))


def test_resonant_lowpass_response():                            # lowpass_attenuates_frequencies_above_cutoff, resonance_boosts_response_near_cutoff
    assert rlp_rms(48000.0, 1000.0, 0.707, 100.0) > rlp_rms(48000.0, 1000.0, 0.707, 8000.0) * 10.0
    assert rlp_rms(48000.0, 1000.0, 4.0, 1000.0) > rlp_rms(48000.0, 1000.0, 0.5, 1000.0) * 4.0


def test_resonant_lowpass_stable_at_extreme_settings_and_sample_rates():   # remains_stable_at_extreme_settings_and_sample_rates
    for sr in (44100.0, 48000.0, 96000.0):
        x = sin_i(int(sr), 0.1)
        for cutoff in (20.0, 1000.0, 20000.0):
            for res in (0.5, 5.0, 10.0):
                out = filt(2, x, cutoff, res, sr=sr)
                assert np.isfinite(out).all() and float(np.abs(out).max()) < 100.0, (sr, cutoff, res)


def test_biquads_attenuate_dc():                                 # biquad_bandpass.rs test_bandpass_attenuates_dc, biquad_highpass.rs test_highpass_attenuates_dc
    assert abs(float(filt(3, np.ones(1000, np.float32), 1000.0, 1.0, 1.0)[-1])) < 0.1
    assert abs(float(filt(4, np.ones(2000, np.float32), 1000.0, 1.0)[-1])) < 0.1


def test_membrane_keeps_ringing_after_excitation():              # membrane_resonator.rs test_membrane_ringing
    x = np.zeros(101, np.float32); x[0] = 1.0
    assert float(np.abs(filt(5, x, 0.01)[1:]).max()) > 0.0001


def polyblep(square, n=44100, inc=100.0 / 44100.0):
    L = O.lib()
    L.orc_polyblep.argtypes = [c.c_int, c.c_double, c.c_void_p, c.c_uint32]
    L.orc_polyblep.restype = None
    out = np.empty(n, np.float32)
    L.orc_polyblep(int(square), inc, out.ctypes.data, n)
    return out


@pytest.mark.parametrize("square", [False, True])
def test_polyblep_range_and_energy(square):                      # polyblep.rs test_polyblep_{saw,square}_range, _not_silent
    out = polyblep(square, inc=float(np.float32(100.0) / np.float32(44100.0)))
    assert float(out.min()) >= -1.1 and float(out.max()) <= 1.1
    assert float((out.astype(np.float64) ** 2).sum()) > 1.0


def morph(morph_v, color, tone_v, n):
    L = O.lib()
    L.orc_morph_osc.argtypes = [c.c_float, c.c_float, c.c_float, c.c_float, c.c_float, c.c_void_p, c.c_uint32]
    L.orc_morph_osc.restype = None
    out = np.empty(n, np.float32)
    L.orc_morph_osc(SR, 440.0, morph_v, color, tone_v, out.ctypes.data, n)
    return out


def test_morph_osc_channels_and_range():                         # morph_osc.rs test_morph_osc_output_range, test_channel{1,2,3}_has_output, gated sine, colour
    out = morph(0.0, 60.0, 50.0, 1000)
    assert np.isfinite(out).all() and float(np.abs(out).max()) < 2.0
    for m in (-1.0, 0.0, 1.0):
        assert float(np.abs(morph(m, 60.0, 50.0, 100)).sum()) > 0.1
    assert float(np.abs(morph(1.0, 60.0, 99.0, 100)).sum()) > 0.1        # gate closed: the noise is still there
    assert float(np.abs(morph(1.0, 20.0, 50.0, 1000)).sum()) > 1.0 and float(np.abs(morph(1.0, 100.0, 50.0, 1000)).sum()) > 1.0


# ================================================================================================ src/instruments/{bass,poly_synth,granulator}.rs
BASS_DEFAULT = [0.24, 0.40, 0.80, 0.00, 0.00, 0.10, 0.15, 0.70, 0.85, 0.15, 0.08, 0.35, 0.10, 0.30, 0.80]      # BassConfig::default() == acid (bass.rs:188-208)
BASS_PRESETS = {"acid": BASS_DEFAULT,
                "sub": [0.18, 1.00, 0.15, 0.00, 0.00, 0.00, 0.70, 0.05, 0.10, 0.30, 0.20, 0.60, 0.15, 0.00, 0.85],
                "reese": [0.18, 0.30, 0.80, 0.80, 0.50, 0.05, 0.35, 0.30, 0.50, 0.40, 0.15, 0.55, 0.12, 0.60, 0.80],
                "stab": [0.30, 0.20, 0.90, 0.00, 0.00, 0.90, 0.20, 0.40, 0.90, 0.08, 0.05, 0.20, 0.08, 0.20, 0.80]}


def bass_render(params, frames):
    return O.render_voices([(4, 0, list(params))], frames, triggers=[(0, 0, 1.0)])[0]


def test_bass_synth_produces_audio_and_presets_differ():         # bass.rs test_bass_synth_produces_audio, test_bass_presets_produce_different_sounds
    assert float((bass_render(BASS_DEFAULT, 44100).astype(np.float64) ** 2).sum()) > 0.1
    outs = {k: bass_render(v, 22050) for k, v in BASS_PRESETS.items()}
    for k, o in outs.items():
        assert float((o.astype(np.float64) ** 2).sum()) > 0.01, k
    assert not np.array_equal(outs["acid"], outs["sub"])


def test_bass_synth_deactivates():                               # bass.rs test_bass_synth_deactivates: amp_decay 0.01, two seconds
    p = list(BASS_DEFAULT); p[11] = 0.01
    out = bass_render(p, 88200)
    assert float(np.abs(out[:2000]).max()) > 0.01 and not out[-22050:].any()      # an inactive voice returns exactly 0 (bass.rs:800-803)


@pytest.fixture
def eng():
    made = []

    def make():
        e = O.oracle_engine()
        e.set_master_gain(1.0)                                   # instrument-level thresholds; the centre pan leaves cos(pi/4) of the level
        e.render(16384)
        made.append(e)
        return e
    yield make
    for e in made:
        e.close()


def test_poly_synth_produces_audio_and_releases(eng):            # poly_synth.rs test_poly_synth_produces_audio, test_poly_synth_release
    e = eng()
    e.poly_trigger_notes([60])
    buf = e.render(4410)[:, 0].astype(np.float64)
    assert float((buf ** 2).sum()) > 0.1 * 0.5                   # the reference's 0.1 at unity, through the centre pan (x 0.5 in energy)
    e.poly_release()
    tail = e.render(441000)
    assert not tail[-44100:].any()


def test_poly_synth_voice_stealing_keeps_six_voices_sounding(eng):   # poly_synth.rs test_poly_synth_six_voices, _voice_stealing
    six, seven = eng(), eng()
    six.poly_trigger_notes(list(range(60, 66)))
    seven.poly_trigger_notes(list(range(60, 66)))
    seven.poly_trigger_notes([66])
    a, b = six.render(4410), seven.render(4410)
    assert np.isfinite(b).all() and float(np.abs(b).max()) > 0.01 and not np.array_equal(a, b)
    assert float(np.abs(b).max()) < 2.0 * float(np.abs(a).max())    # the seventh note replaced a voice, it did not add one


def gran_buffer(n=4410):                                         # granulator.rs test_buffer
    i = np.arange(n, dtype=np.float32)
    return (np.sin(((i / f32(44100.0)) * f32(440.0) * f32(2 * np.pi)).astype(np.float32)).astype(np.float32) * f32(0.5)).astype(np.float32)


def gran(eng, seed, params, velocity=1.0, buffer=None):
    e = eng()
    assert e.granulator_set_buffer(gran_buffer() if buffer is None else buffer, 44100.0)
    e.granulator_set_seed(seed)
    for p, v in params:
        e.granulator_set_param(p, v)
    e.granulator_snap_params()
    e.granulator_trigger(velocity)
    return e


G_GRAIN_LENGTH, G_DENSITY, G_CLOUD, G_VOLUME, G_RAND_TIMING, G_RAND_AMP, G_DRIVE = 1, 4, 7, 8, 9, 10, 11


def test_triggered_granulator_produces_finite_audio_and_same_seed_same_output(eng):   # granulator.rs triggered_granulator_produces_finite_audio, same_seed_produces_same_output
    out = gran(eng, 7, []).render(44100)
    assert np.isfinite(out).all() and float(np.abs(out).max()) > 0.001 * 0.7
    a, b = gran(eng, 99, [], 0.8).render(4096), gran(eng, 99, [], 0.8).render(4096)
    assert np.array_equal(a, b)


def test_dense_cloud_with_random_amp_remains_finite(eng):        # dense_cloud_with_random_amp_remains_finite
    out = gran(eng, 13, [(G_DENSITY, 1.0), (G_RAND_AMP, 1.0), (G_RAND_TIMING, 1.0), (G_CLOUD, 1.0)]).render(88200)
    assert np.isfinite(out).all() and float(np.abs(out).max()) > 0.001 * 0.7


def test_random_timing_preserves_average_density(eng):           # random_timing_preserves_average_density
    out = gran(eng, 101, [(G_DENSITY, 0.5), (G_GRAIN_LENGTH, 0.0), (G_RAND_TIMING, 1.0), (G_CLOUD, 1.0)]).render(88200)[:, 0]
    assert np.isfinite(out).all()
    audible_blocks = sum(float(np.abs(out[b * 4410:(b + 1) * 4410]).max()) > 1e-4 * 0.7 for b in range(20))
    assert audible_blocks >= 12


def test_soft_grain_stealing_does_not_run_away(eng):             # soft_grain_stealing_does_not_click
    out = gran(eng, 31, [(G_DENSITY, 1.0), (G_GRAIN_LENGTH, 1.0), (G_CLOUD, 1.0)]).render(88200)
    assert np.isfinite(out).all() and float(np.abs(out).max()) < 4.0


def test_drive_is_roughly_gain_neutral(eng):                     # drive_is_roughly_gain_neutral
    def peak(drive):
        out = gran(eng, 17, [(G_DENSITY, 0.4), (G_CLOUD, 0.6), (G_VOLUME, 1.0), (G_DRIVE, drive)]).render(44100)
        assert np.isfinite(out).all()
        return float(np.abs(out).max())
    dry, wet = peak(0.0), peak(1.0)
    assert wet <= dry * 1.25 and wet < 4.0


# ================================================================================================ utils/smoother.rs, frame.rs, mixer/graph.rs
def smoother(init, mn, mx, ops, sr=44100.0, ms=10.0):
    L = O.lib()
    L.orc_smoother_script.argtypes = [c.c_float] * 5 + [c.c_void_p, c.c_void_p, c.c_uint32, c.c_void_p]
    L.orc_smoother_script.restype = None
    codes = np.array([o[0] for o in ops], np.uint32)
    vals = np.array([o[1] for o in ops], np.float32)
    out = np.zeros(3, np.float32)
    L.orc_smoother_script(init, mn, mx, sr, ms, codes.ctypes.data, vals.ctypes.data, len(ops), out.ctypes.data)
    return float(out[0]), float(out[1]), bool(out[2])


SET_TARGET, SET_IMMEDIATE, SNAP, SET_NORMALIZED, SET_BIPOLAR, TICK = range(6)


def test_smoother_unit_tests():                                  # smoother.rs test_smoother_reaches_target ... test_snap
    cur, _, settled = smoother(0.0, 0.0, 1.0, [(SET_TARGET, 1.0), (TICK, 4410)])
    assert abs(cur - 1.0) < 0.001 and settled
    assert smoother(0.0, 0.0, 1.0, [(SET_IMMEDIATE, 1.0)]) == (1.0, 1.0, True)
    assert smoother(50.0, 20.0, 200.0, [(SET_TARGET, 300.0)])[1] == 200.0
    assert smoother(50.0, 20.0, 200.0, [(SET_TARGET, 10.0)])[1] == 20.0
    for n, want in [(0.5, 50.0), (0.0, 0.0), (1.0, 100.0)]:
        assert smoother(50.0, 0.0, 100.0, [(SET_NORMALIZED, n)])[1] == want
    for b, want in [(0.0, 50.0), (-1.0, 0.0), (1.0, 100.0)]:
        assert smoother(50.0, 0.0, 100.0, [(SET_BIPOLAR, b)])[1] == want
    assert smoother(0.0, 0.0, 1.0, [(SET_TARGET, 0.75)]) == (0.0, 0.75, False)
    assert smoother(0.0, 0.0, 1.0, [(SET_TARGET, 0.75), (SNAP, 0.0)]) == (0.75, 0.75, True)


def panned(x, pan):
    L = O.lib()
    L.orc_frame_panned.argtypes = [c.c_float, c.c_float, c.c_void_p]
    L.orc_frame_panned.restype = None
    out = np.zeros(2, np.float32)
    L.orc_frame_panned(x, pan, out.ctypes.data)
    return float(out[0]), float(out[1])


def test_stereo_frame_unit_tests():                              # frame.rs tests
    L = O.lib()
    L.orc_frame_downmix.argtypes = [c.c_float, c.c_float]
    L.orc_frame_downmix.restype = c.c_float
    assert L.orc_frame_downmix(-0.3, -0.3) == float(f32(-0.3))   # downmix_of_a_mono_frame_is_the_original_sample
    assert L.orc_frame_downmix(1.0, 0.0) == 0.5                  # downmix_averages_the_two_channels
    l, r = panned(0.8, 0.0)
    assert abs(l - 0.8) < 1e-6 and abs(r) < 1e-6                 # panned_hard_left_silences_right
    l, r = panned(0.8, 1.0)
    assert abs(l) < 1e-6 and abs(r - 0.8) < 1e-6                 # panned_hard_right_silences_left
    l, r = panned(1.0, 0.5)
    assert abs(l - r) < 1e-6 and abs(l - 2 ** -0.5) < 1e-6       # panned_center_is_equal_and_minus_three_db
    for pan in (0.0, 0.25, 0.5, 0.75, 1.0):                      # panned_preserves_power_across_sweep
        l, r = panned(0.6, pan)
        assert abs(l * l + r * r - 0.36) < 1e-5
    assert panned(0.5, -1.0) == panned(0.5, 0.0) and panned(0.5, 2.0) == panned(0.5, 1.0)   # panned_clamps_out_of_range


def graph_frame(gain=1.0, pan=0.5, mute=False, second_solo=False, l=1.0, r=1.0):
    L = O.lib()
    L.orc_graph_one_frame.argtypes = [c.c_float, c.c_float, c.c_int, c.c_int, c.c_float, c.c_float, c.c_void_p]
    L.orc_graph_one_frame.restype = None
    out = np.zeros(4, np.float32)
    L.orc_graph_one_frame(gain, pan, int(mute), int(second_solo), l, r, out.ctypes.data)
    return [float(v) for v in out]


def test_mixer_graph_unit_tests():                               # graph.rs mix_down_records_and_resets_track_peak, snap_strip_params_applies_current_targets_immediately
    assert graph_frame(l=0.25, r=-0.5) == [0.25, -0.5, 0.5, 0.0]
    assert graph_frame(gain=0.5, pan=0.0)[:2] == [0.5, 0.0]
    assert graph_frame(mute=True)[:2] == [0.0, 0.0]
    assert graph_frame(second_solo=True)[:2] == [0.0, 0.0]


# ================================================================================================ engine/sequencer.rs (swing), on the oracle AND on the product's host sequencer
def trigger_frames(swing, frames=6 * 88200):
    """Every step enabled at 120 BPM; the oracle's trigger table and the product's host-side schedule (csrc/engine.cuh HostSeq through
    gooey_b200_sequencer_schedule — no GPU involved) must agree bit for bit (tests/test_host_cpu.py); both are returned."""
    o = O.oracle_engine()
    o.set_bpm(120.0)
    o.set_swing(swing)
    for s in range(16):
        o.sequencer_set_instrument_step(0, s, True)
    fr, _ = O.trigger_table(o, 0, frames)
    o.close()
    from test_host_cpu import schedule
    got, _ = schedule(120.0, swing, [1] * 16, [1.0] * 16, frames)
    assert np.array_equal(got, fr)
    return fr.astype(np.int64)


def test_swing_delays_the_off_beats_and_preserves_the_tempo():   # sequencer.rs test_swing_timing_affects_triggers, test_swing_preserves_average_tempo
    straight, swung = trigger_frames(0.5), trigger_frames(0.75)
    k = 64                                                       # bar 5: the swing smoother (FFI set_swing glides, sequencer.rs:909) has settled
    straight_gap = straight[k + 1] - straight[k]
    assert swung[k + 1] - swung[k] > straight_gap                # the swung off-beat is late
    assert swung[k + 2] - swung[k + 1] < straight_gap            # and the step after it is short
    assert abs((swung[k + 2] - swung[k]) - (straight[k + 2] - straight[k])) <= 2
    assert abs((swung[k + 4] - swung[k]) - (straight[k + 4] - straight[k])) <= 4


# ================================================================================================ tests/effect_distortion_balance.rs
DB_SR, DB_N, DB_WARM, DB_BIN, DB_AMP = 48000.0, 8192, 8192, 37, 0.5


def db_input(total):
    fund = float(f32(DB_BIN) * f32(DB_SR) / f32(DB_N))                                     # fundamental_hz() in f32
    i = np.arange(total, dtype=np.float64)
    return (np.sin(2 * np.pi * fund * i / DB_SR).astype(np.float32) * f32(DB_AMP)).astype(np.float32)


def db_dry():
    i = np.arange(DB_WARM, DB_WARM + DB_N, dtype=np.float64)
    return (np.sin(2 * np.pi * DB_BIN * i / DB_N).astype(np.float32) * f32(DB_AMP)).astype(np.float32)


def rms(x):
    return float(np.sqrt((x.astype(np.float64) ** 2).mean()))


def bin_power(x, b):
    ph = 2 * np.pi * b / len(x) * np.arange(len(x))
    re, im = float((x.astype(np.float64) * np.cos(ph)).sum()), float(-(x.astype(np.float64) * np.sin(ph)).sum())
    return re * re + im * im


def harmonic_distortion(x):
    fund = max(bin_power(x, DB_BIN), 1e-30)
    return float(np.sqrt(sum(bin_power(x, DB_BIN * h) for h in range(2, 11) if DB_BIN * h < DB_N // 2) / fund))


def gain_db(p, dry):
    return 20.0 * np.log10(rms(p) / max(rms(dry), 1e-30))


def test_max_feedback_matches_saturation_gain_and_distortion():   # effect_distortion_balance.rs:90-112
    x, dry = db_input(DB_WARM + DB_N), db_dry()
    sat = Fx(SATURATION, 1.0, 0.5, 1.0, sr=DB_SR).run(x)[DB_WARM:]
    fb = Fx(FBWS, 100.0, 0.98, 2000.0, 1.0, sr=DB_SR).run(x)[DB_WARM:]
    assert abs(gain_db(fb, dry) - gain_db(sat, dry)) <= 1.5
    assert harmonic_distortion(fb) >= harmonic_distortion(sat) * 0.9


def test_mid_feedback_stays_near_mid_saturation_gain():          # effect_distortion_balance.rs:114-128
    x, dry = db_input(DB_WARM + DB_N), db_dry()
    sat = Fx(SATURATION, 0.5, 0.4, 1.0, sr=DB_SR).run(x)[DB_WARM:]
    fb = Fx(FBWS, 50.0, 0.49, 2000.0, 1.0, sr=DB_SR).run(x)[DB_WARM:]
    assert abs(gain_db(fb, dry) - gain_db(sat, dry)) <= 3.0


# ================================================================================================ tests/aliasing.rs
AL_SR, AL_N, AL_J = 48000.0, 8192, 367


def al_dt():
    return float(f32(AL_J) * f32(AL_SR) / f32(AL_N)) / AL_SR     # fundamental_hz() as f64 / SAMPLE_RATE as f64


def alias_to_signal(x, bins):                                    # aliasing.rs:46-57
    x = x.astype(np.float64)
    n = len(x)
    total_positive = (n * float((x ** 2).sum()) - float(x.sum()) ** 2) / 2.0
    signal = sum(bin_power(x, k) for k in bins)
    return max(total_positive - signal, 0.0) / max(signal, 1e-30)


def signal_bins(square):
    return [m * AL_J for m in range(1, AL_N) if m * AL_J <= AL_N // 2 and (not square or m % 2 == 1)]


def naive_wave(square):
    dt, phase, out = al_dt(), 0.0, np.empty(AL_N, np.float32)
    for i in range(AL_N):
        out[i] = (1.0 if phase < 0.5 else -1.0) if square else f32(2.0 * phase - 1.0)
        phase = (phase + dt) % 1.0
    return out


@pytest.mark.parametrize("square", [False, True])
def test_polyblep_suppresses_aliasing(square):                   # polyblep_saw_suppresses_aliasing, polyblep_square_suppresses_aliasing
    bins = signal_bins(square)
    naive = alias_to_signal(naive_wave(square), bins)
    clean = alias_to_signal(polyblep(square, AL_N, al_dt()), bins)
    assert naive > 0.02 and clean < naive * 0.25 and clean < 0.01


def test_oversampling_reduces_naive_square_aliasing():           # oversampling_reduces_naive_square_aliasing (the reconstructed half-band)
    L = O.lib()
    L.orc_oversample_square.argtypes = [c.c_int, c.c_double, c.c_void_p, c.c_uint32]
    L.orc_oversample_square.restype = None

    def render(mode, factor):
        out = np.empty(4096 + AL_N, np.float32)
        L.orc_oversample_square(mode, al_dt() / factor, out.ctypes.data, out.size)
        return out[4096:]
    bins = signal_bins(True)
    off, x4 = alias_to_signal(render(0, 1), bins), alias_to_signal(render(4, 4), bins)
    assert off > 0.02 and x4 < off * 0.5


# ================================================================================================ max_curve.rs, utils/blendable.rs
def test_max_curve_unit_tests():                                 # max_curve.rs test_max_curve_linear ... test_envelope_basic
    L = O.lib()
    mc = lambda p, cv: L.orc_max_curve(float(p), float(cv))
    for p in (0.0, 0.5, 1.0):
        assert abs(mc(p, 0.0) - p) < 0.001
    for cv in (-0.9, -0.5, 0.0, 0.5, 0.9):
        assert abs(mc(0.0, cv)) < 0.001 and abs(mc(1.0, cv) - 1.0) < 0.001
    assert mc(0.5, -0.83) > 0.5 and mc(0.5, 0.83) < 0.5
    L.orc_maxcurve_envelope.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32]
    L.orc_maxcurve_envelope.restype = None
    seg = np.array([1.0, 10.0, 0.0, 0.0, 100.0, 0.0], np.float32)       # (target, ms, curve) x 2
    times = np.array([0.0, 0.005, 0.01, 0.06], np.float64)
    out = np.zeros(4, np.float32)
    L.orc_maxcurve_envelope(seg.ctypes.data, times.ctypes.data, out.ctypes.data, 4)
    assert abs(out[0]) < 0.01 and abs(out[1] - 0.5) < 0.1 and abs(out[2] - 1.0) < 0.1 and abs(out[3] - 0.5) < 0.1


def blend2(corners, x, y):
    L = O.lib()
    L.orc_blend2.argtypes = [c.c_void_p, c.c_float, c.c_float, c.c_void_p]
    L.orc_blend2.restype = None
    cs = np.array(corners, np.float32).reshape(8)
    out = np.zeros(2, np.float32)
    L.orc_blend2(cs.ctypes.data, x, y, out.ctypes.data)
    return float(out[0]), float(out[1])


def test_preset_blender_unit_tests():                            # blendable.rs test_lerp_* (through the x axis), test_blend_at_corners ... test_uniform_blender
    ab = [(0.0, 10.0), (1.0, 20.0), (0.0, 10.0), (1.0, 20.0)]            # bottom: a -> b, top the same: blend(x, .) == a.lerp(b, x)
    assert blend2(ab, 0.0, 0.0) == (0.0, 10.0) and blend2(ab, 1.0, 0.0) == (1.0, 20.0) and blend2(ab, 0.5, 0.0) == (0.5, 15.0)
    sq = [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (1.0, 1.0)]                # bottom_left, bottom_right, top_left, top_right
    assert blend2(sq, 0.0, 0.0) == (0.0, 0.0) and blend2(sq, 1.0, 0.0) == (1.0, 0.0)
    assert blend2(sq, 0.0, 1.0) == (0.0, 1.0) and blend2(sq, 1.0, 1.0) == (1.0, 1.0)
    assert blend2(sq, 0.5, 0.5) == (0.5, 0.5)
    assert blend2(sq, -0.5, 1.5) == (0.0, 1.0)
    assert blend2([(0.5, 0.75)] * 4, 0.3, 0.7) == (0.5, 0.75)
