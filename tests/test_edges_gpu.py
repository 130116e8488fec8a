"""Edge cases the reference's tests care about, on the GPU path: empty and ragged shapes, extreme parameters, other
sample rates and tempi, NULL handling.  Tolerance 1e-5 of full scale."""
import ctypes

import numpy as np
import pytest

from libgooey_b200 import engine as G, voices as V, lib
import oracle_lib as O
import engine_scripts as S

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("frames", [1, 3, 31, 33, 1001, 8193])
def test_ragged_frame_counts(frames):
    patches = [V.patch(V.KICK, V.KICK_PRESETS["punch"]), V.patch(V.SNARE, V.SNARE_PRESETS["hiss"]), V.patch(V.HIHAT, V.HIHAT_PRESETS["loose"]),
               V.patch(V.TOM, V.TOM_PRESETS["ring"], aux=1), V.patch(V.BASS, V.BASS_PRESETS["reese"])]
    vel = np.full(5, 0.8, np.float32)
    b = V.VoiceBatch(patches, 44100.0)
    b.trigger_all(0, vel)
    got = b.render(frames)
    b.close()
    want = O.render_voices(patches, frames, triggers=[(i, 0, 0.8) for i in range(5)])
    assert got.shape == (5, frames)
    assert np.abs(got - want).max() <= TOL


def test_empty_batch_and_zero_frames():
    b = V.VoiceBatch([], 44100.0)
    assert b.render(16).shape == (0, 16)
    b.close()
    e = G.Engine()
    assert len(e.bounce_to_buffer(0)) == 0           # 0 bars: empty, non-NULL buffer (ffi.rs:7897-7915)
    e.close()


@pytest.mark.parametrize("level", [0.0, 1.0])
def test_parameter_extremes(level):
    """Every normalized parameter at its minimum / maximum (decay 0 skips the decay phase)."""
    patches = [V.patch(V.KICK, [level] * 18), V.patch(V.SNARE, [level] * 13 + [3.0 * level] + [level] * 5),
               V.patch(V.HIHAT, [level] * 5), V.patch(V.TOM, [100.0 * level] * 8, aux=1), V.patch(V.BASS, [level] * 15)]
    vel = np.ones(5, np.float32)
    b = V.VoiceBatch(patches, 44100.0)
    b.trigger_all(0, vel)
    b.trigger_all(7000, vel * 0.0)                   # velocity 0 retrigger
    got = b.render(12000)
    b.close()
    want = O.render_voices(patches, 12000, triggers=[(i, 0, 1.0) for i in range(5)] + [(i, 7000, 0.0) for i in range(5)])
    fin = np.isfinite(want)
    assert np.array_equal(fin, np.isfinite(got))
    assert (np.abs(got[fin] - want[fin]) / np.maximum(1.0, np.abs(want[fin]))).max() <= TOL


@pytest.mark.parametrize("sr", [48000.0, 22050.0])
def test_other_sample_rates(sr):
    patches = [V.patch(V.KICK, V.KICK_PRESETS["tight"]), V.patch(V.SNARE, V.SNARE_PRESETS["loose"]), V.patch(V.HIHAT, V.HIHAT_PRESETS["dark"]),
               V.patch(V.TOM, V.TOM_PRESETS["void"], aux=1)]
    vel = np.array([1.0, 0.9, 0.8, 0.7], np.float32)
    b = V.VoiceBatch(patches, sr)
    b.trigger_all(0, vel)
    got = b.render(20000)
    b.close()
    want = O.render_voices(patches, 20000, triggers=[(i, 0, float(vel[i])) for i in range(4)], sample_rate=sr)
    assert np.abs(got - want).max() <= TOL


def test_engines_with_different_tempi_in_one_batch_bounce():
    bpms = [90.0, 120.0, 174.0, 120.0]
    scripts = []
    for i, bpm in enumerate(bpms):
        def sc(e, i=i, bpm=bpm):
            S.pattern_engine(e, 40 + i, notes=False, graph=False)
            e.set_bpm(bpm)
        scripts.append(sc)
    engines = [G.Engine() for _ in bpms]
    for e, sc in zip(engines, scripts):
        sc(e)
    outs = G.batch_bounce(engines, 1)
    for e in engines:
        e.close()
    for i, sc in enumerate(scripts):
        o = O.oracle_engine(); sc(o); want = o.bounce_to_buffer(1); o.close()
        assert len(outs[i]) == len(want) == int(round(4 * 60.0 / bpms[i] * 44100.0))
        assert np.abs(outs[i] - want).max() <= TOL


def test_null_arguments_are_ignored_like_the_reference():
    from libgooey_b200 import LIB_PATH
    L = ctypes.CDLL(LIB_PATH)          # a private handle: raw prototypes here must not leak into the package's bindings
    c = ctypes
    L.gooey_engine_set_kick_param.argtypes = [c.c_void_p, c.c_uint32, c.c_float]
    L.gooey_engine_set_kick_param(None, 0, 0.5)                      # null engine: no-op
    L.gooey_engine_render.argtypes = [c.c_void_p, c.c_void_p, c.c_uint32]
    L.gooey_engine_render(None, None, 64)
    L.gooey_engine_bounce_to_buffer.restype = c.c_void_p
    L.gooey_engine_bounce_to_buffer.argtypes = [c.c_void_p, c.c_uint32, c.c_void_p]
    assert L.gooey_engine_bounce_to_buffer(None, 1, None) is None
    L.gooey_engine_get_bpm.restype = c.c_float
    L.gooey_engine_get_bpm.argtypes = [c.c_void_p]
    assert L.gooey_engine_get_bpm(None) == 120.0                      # sentinel (ffi.rs:3374-3377)
    e = G.Engine()
    e.set_kick_param(99, 0.5)                                         # unknown id: ignored
    e.sequencer_set_instrument_step(7, 0, True)                       # bad instrument: ignored
    e.sequencer_set_instrument_step(0, 99, True)                      # bad step: ignored
    L.gooey_engine_has_error.restype = c.c_bool
    L.gooey_engine_has_error.argtypes = [c.c_void_p]
    assert not L.gooey_engine_has_error(e._h)
    assert np.abs(e.bounce_to_buffer(1)).max() < 1e-3                 # nothing enabled: silence
    e.close()
