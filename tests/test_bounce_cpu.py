"""Host logic of the Rust-API mirror that needs no GPU: BounceLength arithmetic (bounce.rs:20-32) and the WAV writer's
quantisation (bounce.rs:96-130: `(s * scale).round() as iN`, saturating)."""
import os

import numpy as np
import pytest

from libgooey_b200 import bounce as B


def test_bounce_length_to_samples():
    assert B.BounceLength.Bars(1).to_samples(120.0, 44100.0) == 88200        # tests/bounce.rs:31-35
    assert B.BounceLength.Bars(2).to_samples(120.0, 44100.0) == 176400
    assert B.BounceLength.Beats(2.0).to_samples(120.0, 44100.0) == 44100     # tests/bounce.rs:38-45
    assert B.BounceLength.Samples(12345).to_samples(97.0, 48000.0) == 12345
    assert B.BounceLength.Bars(8).to_samples(120.0, 44100.0) == 705600       # C3
    # f64 arithmetic on f32-rounded tempo, round half away from zero
    assert B.BounceLength.Bars(1).to_samples(97.3, 44100.0) == int(np.floor(4.0 * (60.0 / float(np.float32(97.3))) * 44100.0 + 0.5))


@pytest.mark.parametrize("bits", [16, 24])
def test_wav_writer_quantisation_and_header(tmp_path, bits):
    x = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 1.5, -1.5, 1e-5, -1e-5, 0.25 + 1.0 / 65534.0, np.nan, 3e4, -3e4], np.float32)
    path = os.path.join(tmp_path, "t.wav")
    B.write_wav(path, x, 44100, bits)
    raw = open(path, "rb").read()
    bps = bits // 8
    assert raw[:4] == b"RIFF" and int.from_bytes(raw[4:8], "little") == 36 + len(x) * bps
    assert raw[8:16] == b"WAVEfmt " and int.from_bytes(raw[16:20], "little") == 16
    assert int.from_bytes(raw[20:22], "little") == 1 and int.from_bytes(raw[22:24], "little") == 1
    assert int.from_bytes(raw[24:28], "little") == 44100 and int.from_bytes(raw[28:32], "little") == 44100 * bps
    assert int.from_bytes(raw[32:34], "little") == bps and int.from_bytes(raw[34:36], "little") == bits
    assert raw[36:40] == b"data" and int.from_bytes(raw[40:44], "little") == len(x) * bps
    data = np.frombuffer(raw[44:], np.uint8).reshape(-1, bps).astype(np.int64)
    got = sum(data[:, k] << (8 * k) for k in range(bps))
    got = np.where(got >= 1 << (bits - 1), got - (1 << bits), got)
    scale = np.float32(32767.0 if bits == 16 else 8388607.0)
    want = []
    for s in x:
        v = np.float32(s) * scale
        if np.isnan(v):
            q = 0
        else:
            r = np.floor(float(v) + 0.5) if v >= 0 else -np.floor(-float(v) + 0.5)
            if bits == 16:
                q = int(min(max(r, -32768), 32767))
            else:
                q = int(min(max(r, -2 ** 31), 2 ** 31 - 1)) & 0xffffff        # hound keeps the low 24 bits of the i32
                q = q - (1 << 24) if q >= 1 << 23 else q
        want.append(q)
    assert got.tolist() == want


def test_wav_rejects_other_bit_depths(tmp_path):
    with pytest.raises(B.GooeyError):
        B.write_wav(os.path.join(tmp_path, "t.wav"), np.zeros(4, np.float32), 44100, 8)
