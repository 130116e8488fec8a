"""GPU parity of the engine-level path (sequencer schedule -> voices -> strips -> mixer graph -> master -> global
effects -> bounce downmix) against the CPU oracle driven by the same FFI call script.  Tolerance 1e-5 (BASELINE.json)."""
import numpy as np
import pytest

from libgooey_b200 import engine as G
import oracle_lib as O
import engine_scripts as S

pytestmark = pytest.mark.gpu
TOL = 1e-5


def both(script, bars=1):
    o = O.oracle_engine()
    g = G.Engine()
    script(o)
    script(g)
    want = o.bounce_to_buffer(bars)
    got = g.bounce_to_buffer(bars)
    o.close(); g.close()
    assert got.shape == want.shape
    return got, want


def test_default_engine_kick_pattern_bounce():
    def script(e):
        for s in (0, 4, 8, 12):
            e.sequencer_set_instrument_step(S.KICK, s, True)
        e.sequencer_set_instrument_step(S.HIHAT, 2, True)
        e.sequencer_set_instrument_step(S.SNARE, 4, True)
    got, want = both(script)
    assert got.shape == (88200,)
    assert np.abs(want).max() > 0.01
    err = np.abs(got - want).max()
    print("default engine bounce err", err)
    assert err <= TOL


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_pattern_engine_with_graph_and_swing(seed):
    def script(e):
        S.random_voice_params(e, seed)
        S.pattern_engine(e, seed, swing=0.4 + 0.1 * seed)
    got, want = both(script, bars=2)
    err = np.abs(got - want).max()
    print("pattern engine err", err, "peak", np.abs(want).max())
    assert np.abs(want).max() > 0.01
    assert err <= TOL


@pytest.mark.parametrize("plate", [False, True])
def test_global_effect_chain(plate):
    def script(e):
        S.pattern_engine(e, 11, notes=False)
        S.fx_chain(e, 5, plate=plate, limiter=True)
    got, want = both(script, bars=1)
    err = np.abs(got - want).max()
    print("fx chain err", err, "peak", np.abs(want).max())
    assert err <= TOL


def test_batch_bounce_matches_single_engine_bounces():
    scripts = [lambda e, s=s: (S.random_voice_params(e, s), S.pattern_engine(e, s)) for s in range(8)]
    engines = [G.Engine() for _ in scripts]
    for e, sc in zip(engines, scripts):
        sc(e)
    outs = G.batch_bounce(engines, 1)
    for e in engines:
        e.close()
    for i, sc in enumerate(scripts):
        o = O.oracle_engine()
        sc(o)
        want = o.bounce_to_buffer(1)
        o.close()
        assert np.abs(outs[i] - want).max() <= TOL


def test_stereo_render_and_manual_trigger_mute_solo_pan():
    def script(e):
        e.set_instrument_pan(S.KICK, 0.1)
        e.set_instrument_gain(S.SNARE, 0.5)
        e.set_instrument_mute(S.HIHAT, True)
        e.set_master_gain(0.5)
        for i in range(5):
            e.trigger_instrument_with_velocity(i, 0.9)
    o = O.oracle_engine(); g = G.Engine()
    script(o); script(g)
    a = np.concatenate([o.render(3000), o.render(5000)])
    b = np.concatenate([g.render(3000), g.render(5000)])
    o.close(); g.close()
    assert np.abs(a).max() > 0.01
    assert np.abs(a - b).max() <= TOL
