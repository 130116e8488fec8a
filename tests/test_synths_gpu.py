"""Poly synth ("chord oscillators") and granulator (row a22, BASELINE config C4 in miniature) through
the FFI engine on the GPU vs the oracle.  Tolerance 1e-5 of full scale."""
import numpy as np
import pytest

from libgooey_b200 import engine as G
import oracle_lib as O
import engine_scripts as S

pytestmark = pytest.mark.gpu
TOL = 1e-5


def c4_source(n=2 * 44100, seed=9):
    """SURVEY 8d C4 source formula (shortened): 220 Hz tone with a slow tremolo plus uniform noise."""
    i = np.arange(n, dtype=np.float64)
    rng = np.random.default_rng(seed)
    x = 0.5 * np.sin(2 * np.pi * 220.0 * i / 44100.0) * (0.5 + 0.5 * np.sin(2 * np.pi * 0.1 * i / 44100.0)) + 0.1 * rng.uniform(-1, 1, n)
    return x.astype(np.float32)


def both_render(script, frames_list=(20000, 24100)):
    o = O.oracle_engine(); g = G.Engine()
    script(o); script(g)
    a = np.concatenate([o.render(f) for f in frames_list])
    b = np.concatenate([g.render(f) for f in frames_list])
    o.close(); g.close()
    return b, a


@pytest.mark.parametrize("preset", [0, 1, 2, 3, 4])
def test_poly_chord_presets(preset):
    def script(e):
        e.poly_trigger_notes([48, 55, 64, 71], preset, 0.9)
    got, want = both_render(script)
    err = np.abs(got - want).max()
    print("poly preset", preset, "err", err, "peak", np.abs(want).max())
    assert np.abs(want).max() > 1e-3
    assert err <= TOL


def test_poly_retrigger_steals_voices_and_release():
    o = O.oracle_engine(); g = G.Engine()
    outs = []
    for e in (o, g):
        e.poly_set_param(2, 0.8)
        e.poly_trigger_notes([40, 47, 52, 56], 2, 1.0)
        a = e.render(6000)
        e.poly_trigger_notes([45, 52, 57, 60, 64, 69], 3, 0.7)      # 4 + 6 notes > 6 voices: steals the oldest
        b = e.render(9000)
        e.poly_release()
        c = e.render(12000)
        outs.append(np.concatenate([a, b, c]))
        e.close()
    err = np.abs(outs[0] - outs[1]).max()
    print("poly steal/release err", err)
    assert np.abs(outs[0]).max() > 1e-3
    assert err <= TOL


@pytest.mark.parametrize("seed", [1, 7])
def test_granulator_cloud_c4_style(seed):
    src = c4_source()
    rng = np.random.default_rng(seed)
    pitch, texture = float(rng.uniform(0.3, 0.7)), float(rng.random())

    def script(e):
        assert e.granulator_set_buffer(src, 44100.0)
        for p, v in [(4, 1.0), (1, 0.55), (2, 0.5), (3, pitch), (6, 0.3), (5, texture), (9, 0.3), (10, 0.3), (7, 1.0)]:
            e.granulator_set_param(p, v)
        e.granulator_set_seed(seed)
        e.granulator_snap_params()
        e.granulator_trigger(0.9)
    got, want = both_render(script, frames_list=(30000, 30000))
    err = np.abs(got - want).max()
    print("granulator err", err, "peak", np.abs(want).max())
    assert np.abs(want).max() > 0.01
    assert err <= TOL


def test_granulator_with_drive_inside_pattern_engine_bounce_and_shared_buffer():
    src = c4_source(44100)

    def script(e, share_from=None):
        S.pattern_engine(e, 4, notes=False)
        if share_from is None:
            assert e.granulator_set_buffer(src, 44100.0)
        else:
            assert e.granulator_share_buffer(share_from)
        e.granulator_set_param(4, 0.6); e.granulator_set_param(11, 0.5); e.granulator_set_param(1, 0.3)
        e.granulator_snap_params()
        e.granulator_trigger(1.0)
        e.poly_trigger_notes([50, 57, 62], 1, 0.8)
    o = O.oracle_engine(); script(o); want = o.bounce_to_buffer(1); o.close()
    g1 = G.Engine(); script(g1)
    g2 = G.Engine(); script(g2, share_from=g1)
    outs = G.batch_bounce([g1, g2], 1)
    g1.close(); g2.close()
    for got in outs:
        err = np.abs(got - want).max()
        print("engine + gran + poly err", err, "peak", np.abs(want).max())
        assert err <= TOL


def test_granulator_rejects_bad_buffers_like_the_reference():
    g = G.Engine()
    assert not g.granulator_set_buffer(np.array([0.0, np.nan], np.float32), 44100.0)
    assert not g.granulator_set_buffer(np.zeros(4, np.float32), 0.0)
    assert g.granulator_set_buffer(np.zeros(4, np.float32), 22050.0)
    g.close()
