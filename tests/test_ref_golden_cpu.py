"""Oracle vs REFERENCE vectors (tests/golden/ref, produced by rust/examples/dump_golden.rs inside libgooey).
Every test skips with "parity unpinned" while the vector is absent; the scripts themselves are always checked."""
import ctypes
import os

import numpy as np

import oracle_lib as O
import ref_golden as R

TOL = 1e-5


def test_scripts_are_current_and_replayable():
    """The committed call scripts replay on the oracle and equal a fresh recording of the Python scripts."""
    import subprocess, sys, tempfile, shutil
    keep = {f: open(os.path.join(R.SCRIPTS, f)).read() for f in os.listdir(R.SCRIPTS)}
    subprocess.run([sys.executable, os.path.join(os.path.dirname(R.REF), "make_ref_scripts.py")], check=True, capture_output=True)
    for f, text in keep.items():
        assert open(os.path.join(R.SCRIPTS, f)).read() == text, f"{f} is stale: rerun tests/golden/make_ref_scripts.py"
    import golden_cases as GC
    for name, script in list(GC.ENGINE_CASES.items()) + list(GC.REF_CASES.items()):
        a = O.oracle_engine(); b = O.oracle_engine()
        script(a); bars = R.replay(b, name)
        wa, wb = a.bounce_to_buffer(1), b.bounce_to_buffer(bars)
        a.close(); b.close()
        assert np.array_equal(wa, wb), name


def test_default_hasher_values():
    want = R.load("hasher.u64le", np.uint64)
    got = np.array([O.lib().orc_siphash(0, 0, n, 1, 3) for n in range(32)], np.uint64)
    assert np.array_equal(got, want)


def test_halfband_impulse_responses():
    up = R.load("halfband_up8.f32le"); down = R.load("halfband_down8.f32le")
    L = O.lib()
    L.orc_halfband_impulse.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    gu = np.zeros(128, np.float32); gd = np.zeros(128, np.float32)
    L.orc_halfband_impulse(gu.ctypes.data, gd.ctypes.data, 64)
    assert np.abs(gu - up).max() <= 1e-6
    assert np.abs(gd - down).max() <= 1e-6


def test_oversampler_tanh():
    want = R.load("oversampler.f32le")
    L = O.lib()
    L.orc_oversampler_tanh.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_float]
    n = np.arange(256, dtype=np.float32)
    sig = (np.float32(0.8) * np.sin(np.float32(2.0) * np.float32(np.pi) * np.float32(1000.0) * n / np.float32(44100.0))).astype(np.float32)
    got = np.zeros(512, np.float32)
    L.orc_oversampler_tanh(sig.ctypes.data, got.ctypes.data, 256, 3.0)
    assert np.abs(got - want).max() <= TOL


def test_preset_kit_vs_reference():
    want = R.load("preset_kit.f32le")
    import golden_cases as GC
    patches, vel, names = GC.kit_patches()
    got = O.render_voices(patches, GC.KIT_FRAMES, triggers=[(i, 0, float(vel[i])) for i in range(len(patches))])
    err = np.abs(got - want.reshape(got.shape)).max(axis=1)
    assert err.max() <= TOL, dict(zip(names, err))


def test_sweep64_vs_reference():
    want = R.load("sweep64.f32le")
    patches, vel = R.sweep_voices()
    got = O.render_voices(patches, 8192, triggers=[(i, 0, float(vel[i])) for i in range(len(patches))], threads=os.cpu_count() or 1)
    assert np.abs(got - want.reshape(got.shape)).max() <= TOL


def test_engine_scripts_vs_reference():
    for f in sorted(os.listdir(R.SCRIPTS)):
        if not f.endswith(".calls"):
            continue
        name = f[:-6]
        want = R.load(f"engine_{name}.f32le")
        o = O.oracle_engine(); bars = R.replay(o, name); got = o.bounce_to_buffer(bars); o.close()
        assert len(got) == len(want) and np.abs(got - want).max() <= TOL, name
