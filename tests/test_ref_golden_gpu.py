"""CUDA path vs REFERENCE vectors (tests/golden/ref).  Skips with "parity unpinned" while a vector is absent."""
import os

import numpy as np
import pytest

import ref_golden as R

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_preset_kit_vs_reference():
    want = R.load("preset_kit.f32le")
    from libgooey_b200 import voices as V
    import golden_cases as GC
    patches, vel, names = GC.kit_patches()
    b = V.VoiceBatch(patches); b.trigger_all(0, vel); got = b.render(GC.KIT_FRAMES); b.close()
    err = np.abs(got - want.reshape(got.shape)).max(axis=1)
    assert err.max() <= TOL, dict(zip(names, err))


def test_sweep64_vs_reference():
    want = R.load("sweep64.f32le")
    from libgooey_b200 import voices as V
    raw, vel = R.sweep_voices()
    patches = [V.patch(i, p, aux=a) for (i, a, p) in raw]
    b = V.VoiceBatch(patches); b.trigger_all(0, vel); got = b.render(8192); b.close()
    assert np.abs(got - want.reshape(got.shape)).max() <= TOL


def test_c1_vs_reference():
    want = R.load("c1_kick.f32le")
    from libgooey_b200 import bounce as B
    e = B.Engine(44100.0); e.set_bpm(120.0); e.add_instrument("kick", B.KickDrum(44100.0))
    e.add_sequencer(B.Sequencer.with_pattern(120.0, 44100.0, [i == 0 for i in range(16)], "kick"))
    got = B.bounce_to_buffer(e, B.BounceLength.Samples(44100))
    assert np.abs(got - want).max() <= TOL


def test_engine_scripts_vs_reference():
    from libgooey_b200 import engine as G
    for f in sorted(os.listdir(R.SCRIPTS)):
        if not f.endswith(".calls"):
            continue
        name = f[:-6]
        want = R.load(f"engine_{name}.f32le")
        g = G.Engine(); bars = R.replay(g, name); got = g.bounce_to_buffer(bars); g.close()
        assert len(got) == len(want) and np.abs(got - want).max() <= TOL, name
