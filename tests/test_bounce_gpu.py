"""Rust-API bounce path on the GPU (libgooey_b200.bounce) — the reference's own tests/bounce.rs restated, plus parity of
BASELINE.json config C1 (single kick voice, default params, 1 s at 44.1 kHz via bounce.rs) and of multi-instrument
engines against the oracle's RustEngine.  Tolerance 1e-5 max-abs relative to full scale (BASELINE.json)."""
import ctypes
import os

import numpy as np
import pytest

from libgooey_b200 import bounce as B
from libgooey_b200 import voices as V
import oracle_lib as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
SR = 44100.0


def setup_engine(pattern_on=(0, 4, 8, 12)):
    """tests/bounce.rs:5-24"""
    engine = B.Engine(SR)
    engine.set_bpm(120.0)
    engine.add_instrument("kick", B.KickDrum(SR))
    pattern = [i in pattern_on for i in range(16)]
    engine.add_sequencer(B.Sequencer.with_pattern(120.0, SR, pattern, "kick"))
    return engine


oracle_rust_bounce = O.rust_bounce


def test_bounce_correct_length():
    engine = setup_engine()
    assert len(B.bounce_to_buffer(engine, B.BounceLength.Bars(1))) == 88200
    assert len(B.bounce_to_buffer(engine, B.BounceLength.Bars(2))) == 176400


def test_bounce_beats_length():
    assert len(B.bounce_to_buffer(setup_engine(), B.BounceLength.Beats(2.0))) == 44100


def test_bounce_produces_audio():
    assert np.abs(B.bounce_to_buffer(setup_engine(), B.BounceLength.Bars(1))).max() > 0.01


def test_bounce_deterministic():
    a = B.bounce_to_buffer(setup_engine(), B.BounceLength.Bars(1))
    b = B.bounce_to_buffer(setup_engine(), B.BounceLength.Bars(1))
    assert np.array_equal(a, b)


def test_bounce_silent_when_no_pattern():
    assert np.abs(B.bounce_to_buffer(setup_engine(pattern_on=()), B.BounceLength.Bars(1))).max() < 0.001


def test_c1_single_kick_default_params_one_second_matches_oracle():
    """BASELINE.json configs[0]."""
    engine = setup_engine(pattern_on=(0,))
    got = B.bounce_to_buffer(engine, B.BounceLength.Samples(44100))
    want = oracle_rust_bounce([("kick", B.KickDrum(SR))], [("kick", [1] + [0] * 15, [1.0] * 16)], 44100)
    err = np.abs(got - want).max()
    print("C1 max|gpu-oracle| =", err, "peak", np.abs(want).max())
    assert np.abs(want).max() > 0.01
    assert err <= TOL


def test_second_bounce_of_the_same_engine_keeps_voice_state_like_the_reference():
    engine = setup_engine()
    B.bounce_to_buffer(engine, B.BounceLength.Bars(1))
    got = B.bounce_to_buffer(engine, B.BounceLength.Bars(1))
    # oracle: same engine bounced twice
    L = O.lib()
    c = ctypes
    oracle_rust_bounce([], [], 1)  # binds argtypes
    e = L.orc_rust_engine_new(SR)
    L.orc_rust_engine_add_instrument(e, b"kick", c.byref(B.KickDrum(SR).patch))
    en = np.array([i in (0, 4, 8, 12) for i in range(16)], np.uint8); ve = np.ones(16, np.float32)
    L.orc_rust_engine_add_sequencer(e, b"kick", en.ctypes.data, ve.ctypes.data, 16)
    want = np.zeros(88200, np.float32)
    L.orc_rust_engine_bounce_samples(e, 88200, want.ctypes.data)
    L.orc_rust_engine_bounce_samples(e, 88200, want.ctypes.data)
    L.orc_rust_engine_free(e)
    assert np.abs(got - want).max() <= TOL


def test_engine_batch_of_drum_kits_matches_oracle():
    rng = np.random.default_rng(5)
    n = 6
    batch = B.EngineBatch(n, SR)
    specs = []
    for i in range(n):
        e = batch[i]
        insts = [("kick", B.KickDrum(SR, ["tight", "punch", "loose", "dirt"][i % 4])), ("snare", B.SnareDrum(SR, ["tight", "loose", "hiss", "smack"][i % 4])),
                 ("hat", B.HiHat2(SR, ["short", "loose", "dark", "soft"][i % 4])), ("tom", B.Tom2(SR, ["derp", "ring", "brush", "void"][i % 4])), ("bass", B.BassSynth(SR))]
        seqs = []
        for name, inst in insts:
            e.add_instrument(name, inst)
            en = (rng.random(16) < 0.3).astype(np.uint8)
            ve = rng.uniform(0.3, 1.0, 16).astype(np.float32)
            e.add_sequencer(B.Sequencer.with_velocity_pattern(120.0, SR, list(zip(en.tolist(), ve.tolist())), name))
            seqs.append((name, en, ve))
        master = float(rng.uniform(0.2, 0.9))
        e.set_master_gain(master)
        limiter = i % 2 == 0
        if not limiter:
            e.clear_global_effects()
        specs.append((insts, seqs, master, limiter))
    got = batch.bounce(B.BounceLength.Bars(1))
    batch.close()
    for i, (insts, seqs, master, limiter) in enumerate(specs):
        want = oracle_rust_bounce(insts, seqs, 88200, master=master, limiter=limiter)
        err = np.abs(got[i] - want).max()
        print(f"engine {i}: err {err:.3e} peak {np.abs(want).max():.3f}")
        assert np.abs(want).max() > 0.01
        assert err <= TOL


def test_bounce_to_wav_16_and_24_bit(tmp_path):
    engine = setup_engine()
    buf = B.bounce_to_buffer(setup_engine(), B.BounceLength.Beats(1.0))
    for bits in (16, 24):
        path = os.path.join(tmp_path, f"k{bits}.wav")
        B.bounce_to_wav(setup_engine(), B.BounceLength.Beats(1.0), path, B.WavConfig(bits))
        raw = open(path, "rb").read()
        assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and len(raw) == 44 + len(buf) * bits // 8
        scale = np.float32(32767.0 if bits == 16 else 8388607.0)
        x = buf * scale
        want = np.where(x >= 0, np.floor(x + np.float32(0.5)), -np.floor(-x + np.float32(0.5))).astype(np.int64)
        data = np.frombuffer(raw[44:], np.uint8).reshape(-1, bits // 8).astype(np.int64)
        got = sum(data[:, k] << (8 * k) for k in range(bits // 8))
        got = np.where(got >= 1 << (bits - 1), got - (1 << bits), got)
        assert np.array_equal(got, want)
    with pytest.raises(B.GooeyError):
        B.bounce_to_wav(engine, B.BounceLength.Beats(1.0), os.path.join(tmp_path, "x.wav"), B.WavConfig(8))
