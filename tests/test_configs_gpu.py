"""BASELINE.json configurations C3 / C4 / C5 in miniature, through the batch entry points, spot-checked against the oracle.
Full sizes are run by tools/engine_scale.py (results in profiles/README.md); here the batch is big enough to exercise the
launch geometry (several warps of engines, several pieces) and small enough for the oracle to finish in seconds.
Tolerance: 1e-5 of full scale, relative to max(1, |reference|) where the reference itself overflows (see below)."""
import numpy as np
import pytest

from libgooey_b200 import engine as G
import oracle_lib as O
import engine_scripts as S

pytestmark = pytest.mark.gpu
TOL = 1e-5


def compare(got, want):
    """The reference's Chamberlin SVF blows up for some random snare patches (high cutoff x low resonance) and its
    output overflows; parity then means: the same frames are non-finite, and finite ones agree relative to their size."""
    fin = np.isfinite(want)
    assert np.array_equal(fin, np.isfinite(got))
    return float((np.abs(got[fin] - want[fin]) / np.maximum(1.0, np.abs(want[fin]))).max()) if fin.any() else 0.0


def c3_script(e, i):
    S.random_voice_params(e, 1000 + i)
    S.pattern_engine(e, 2000 + i, swing=None if i % 2 == 0 else 0.4 + 0.3 * ((i * 37) % 100) / 100.0)


def test_c3_pattern_engines_with_mixer_graph_batch():
    n, bars = 72, 2
    engines = [G.Engine() for _ in range(n)]
    for i, e in enumerate(engines):
        c3_script(e, i)
    outs = G.batch_bounce(engines, bars)
    for e in engines:
        e.close()
    assert all(len(o) == 176400 for o in outs)
    worst = 0.0
    wants = O.bounce_many(c3_script, range(n), bars)             # every engine of the batch
    for i in range(n):
        want = wants[i]
        err = compare(outs[i], want)
        if i % 12 == 0:
            print(f"C3 engine {i}: err {err:.3e} peak {np.abs(want[np.isfinite(want)]).max():.3f}")
        worst = max(worst, err)
    print(f"C3: {n} engines checked, worst err {worst:.3e}")
    assert worst <= TOL


@pytest.mark.parametrize("plate", [False, True])
def test_c5_drum_bass_with_delay_reverb_tilt_chain_batch(plate):
    n = 40

    def script(e, i):
        S.random_voice_params(e, 500 + i)
        S.pattern_engine(e, 600 + i, notes=(i % 3 == 0), graph=(i % 2 == 0))
        S.fx_chain(e, 700 + i, plate=plate, spring=not plate or i % 2 == 0)
    engines = [G.Engine() for _ in range(n)]
    for i, e in enumerate(engines):
        script(e, i)
    outs = G.batch_bounce(engines, 1)
    for e in engines:
        e.close()
    worst = 0.0
    wants = O.bounce_many(script, range(n), 1)
    for i in range(n):
        err = compare(outs[i], wants[i])
        worst = max(worst, err)
    print(f"C5 plate={plate}: {n} engines checked, worst err {worst:.3e}")
    assert worst <= TOL


def test_c4_granulators_over_one_shared_source():
    n = 48
    rng = np.random.default_rng(3)
    src = (0.5 * np.sin(2 * np.pi * 220.0 * np.arange(3 * 44100) / 44100.0) * (0.5 + 0.5 * np.sin(2 * np.pi * 0.1 * np.arange(3 * 44100) / 44100.0))
           + 0.1 * rng.uniform(-1, 1, 3 * 44100)).astype(np.float32)
    pitch = rng.uniform(0.3, 0.7, n); tex = rng.random(n)

    def script(e, i, first=None):
        if first is None:
            assert e.granulator_set_buffer(src, 44100.0)
        else:
            assert e.granulator_share_buffer(first)
        for p, v in [(4, 1.0), (1, 0.55), (2, 0.5), (3, float(pitch[i])), (6, 0.3), (5, float(tex[i])), (9, 0.3), (10, 0.3), (7, 1.0)]:
            e.granulator_set_param(p, v)
        e.granulator_set_seed(i + 1)
        e.granulator_snap_params()
        e.granulator_trigger(1.0)
    engines = [G.Engine() for _ in range(n)]
    for i, e in enumerate(engines):
        script(e, i, first=None if i == 0 else engines[0])
    outs = G.batch_bounce(engines, 1)
    for e in engines:
        e.close()
    wants = O.bounce_many(lambda o, i: script(o, i), range(n), 1)
    worst = 0.0
    for i in range(n):
        want = wants[i]
        err = compare(outs[i], want)
        assert np.abs(want).max() > 0.01
        worst = max(worst, err)
    print(f"C4: {n} granulators checked, worst err {worst:.3e}")
    assert worst <= TOL
