"""GPU parity for the drum voices: CUDA path (through the C ABI) vs the CPU oracle on the same patches.
Tolerance: 1e-5 max-abs relative to full scale (BASELINE.json north_star)."""
import numpy as np
import pytest

from libgooey_b200 import voices as V
import oracle_lib as O
from workloads import drum_sweep_patches

pytestmark = pytest.mark.gpu
TOL = 1e-5


def run_both(patches, frames, vel, params=()):
    b = V.VoiceBatch(patches, 44100.0)
    b.trigger_all(0, vel)
    for (v, f, p, x, s) in params:
        b.set_param(v, p, x, frame=f, snap=s)
    got = b.render(frames)
    b.close()
    want = O.render_voices(patches, frames, triggers=[(i, 0, float(vel[i])) for i in range(len(patches))], params=params, threads=8)
    return got, want


def report(got, want, kinds):
    err = np.abs(got - want).max(axis=1)
    for k in sorted(set(kinds)):
        print(f"instrument {k}: max err {err[np.array(kinds) == k].max():.3e}")
    return err


def test_presets_default_velocity():
    patches, kinds = [], []
    for name, p in V.KICK_PRESETS.items():
        patches.append(V.patch(V.KICK, p)); kinds.append(0)
    for name, p in V.SNARE_PRESETS.items():
        patches.append(V.patch(V.SNARE, p)); kinds.append(1)
    for name, p in V.HIHAT_PRESETS.items():
        patches.append(V.patch(V.HIHAT, p)); kinds.append(2)
    for name, p in V.TOM_PRESETS.items():
        patches.append(V.patch(V.TOM, p, aux=1)); kinds.append(3)
    patches.append(V.patch(V.TOM)); kinds.append(3)
    vel = np.linspace(0.3, 1.0, len(patches)).astype(np.float32)
    got, want = run_both(patches, 44100, vel)
    err = report(got, want, kinds)
    assert np.isfinite(got).all()
    assert err.max() <= TOL


@pytest.mark.parametrize("exact_tier", [True, False])
def test_random_drum_sweep(exact_tier):
    patches, vel, kinds = drum_sweep_patches(256, seed=0x600E7, exact_tier=exact_tier)
    got, want = run_both(patches, 22050, vel)
    err = report(got, want, kinds)
    assert np.isfinite(got).all()
    assert err.max() <= TOL


def test_retrigger_and_param_edit():
    patches = [V.patch(V.KICK, V.KICK_PRESETS["punch"]), V.patch(V.SNARE, V.SNARE_PRESETS["loose"]),
               V.patch(V.HIHAT, V.HIHAT_PRESETS["loose"]), V.patch(V.TOM, V.TOM_PRESETS["ring"], aux=1)]
    vel = np.array([0.9, 0.7, 1.0, 1.0], np.float32)
    b = V.VoiceBatch(patches, 44100.0)
    b.trigger_all(0, vel)
    trig = [(i, 0, float(vel[i])) for i in range(4)]
    params = [(0, 3000, 0, 0.8, False), (1, 3000, 0, 0.6, False), (2, 3000, 0, 0.2, False), (3, 3000, 0, 0.9, False),
              (0, 9000, 4, 0.5, True), (1, 9000, 10, 0.3, True)]
    for (v, f, p, x, s) in params:
        b.set_param(v, p, x, frame=f, snap=s)
    for i in range(4):
        b.trigger(i, 11026, 0.6)
        trig.append((i, 11026, 0.6))
    got = b.render(30000)
    b.close()
    want = O.render_voices(patches, 30000, triggers=trig, params=params)
    err = np.abs(got - want).max(axis=1)
    print("retrigger errs", err)
    assert err.max() <= TOL


def test_two_renders_continue_state():
    patches = [V.patch(V.KICK, V.KICK_PRESETS["dirt"]), V.patch(V.HIHAT, V.HIHAT_PRESETS["loose"])]
    vel = np.array([1.0, 1.0], np.float32)
    b = V.VoiceBatch(patches, 44100.0)
    b.trigger_all(0, vel)
    a1 = b.render(5000)
    a2 = b.render(7003)
    b.close()
    b2 = V.VoiceBatch(patches, 44100.0)
    b2.trigger_all(0, vel)
    whole = b2.render(12003)
    b2.close()
    two = np.concatenate([a1, a2], axis=1)
    assert np.abs(two[1] - whole[1]).max() <= 2e-6          # hi-hat: 2x2 scans of its high-passes
    assert np.abs(two[0] - whole[0]).max() <= 2e-6          # kick: half-band scans re-associate per 32-frame block


def test_wave_backend_matches_serial_backend(monkeypatch):
    """Kernel W (warp per voice, oversampler as prefix scans) against kernel C (reference operation order) on the
    same device state: the only difference allowed is the re-association noise of the half-band scans."""
    patches, vel, kinds = drum_sweep_patches(128, seed=7, exact_tier=False)

    def render():
        b = V.VoiceBatch(patches, 44100.0)
        b.trigger_all(0, vel)
        b.trigger_all(9000, vel)
        out = b.render(20000)
        b.close()
        return out

    monkeypatch.setenv("GOOEY_B200_BACKEND", "serial")
    ser = render()
    monkeypatch.delenv("GOOEY_B200_BACKEND")
    wav = render()
    err = np.abs(ser - wav).max(axis=1)
    for k in range(4):
        print(f"instrument {k}: max |wave - serial| = {err[np.array(kinds) == k].max():.3e}")
    assert np.isfinite(wav).all()
    assert err.max() <= 5e-6
