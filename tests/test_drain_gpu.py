"""Host-facing drains of the batch entry points: pitched f32 block with overlapped drain, 16-bit PCM quantised on the device
(bounce.rs:105-113 / ffi.rs:7968-7972: `(s * 32767).round() as i16`, half away from zero, saturating), batch WAV files, NUMA-placed
pinned buffers, and one batch spread over two devices (skipped on a one-GPU box)."""
import numpy as np
import pytest

from libgooey_b200 import engine as G, voices as V, lib, HostBuffer
import engine_scripts as S
from workloads import drum_sweep_patches

pytestmark = pytest.mark.gpu


def pcm16_ref(x):
    """f32 `(s * 32767.0).round() as i16` with Rust's rounding (half away from zero) and saturating cast."""
    r = (x.astype(np.float32) * np.float32(32767.0)).astype(np.float32)
    r = np.where(np.isnan(r), np.float32(0), np.sign(r) * np.floor(np.abs(r) + np.float32(0.5)))
    return np.clip(r, -32768, 32767).astype(np.int16)


def script(e, i):
    S.random_voice_params(e, 1000 + i)
    S.pattern_engine(e, 2000 + i, swing=None if i % 2 == 0 else 0.45 + 0.01 * i)
    if i % 3 == 0:
        S.fx_chain(e, 3000 + i, plate=(i % 2 == 0))


def make(n):
    es = [G.Engine() for _ in range(n)]
    for i, e in enumerate(es):
        script(e, i)
    return es


def test_voice_batch_pcm16_is_the_quantised_f32_render():
    patches, vel, _ = drum_sweep_patches(64, seed=11)
    b = V.VoiceBatch(patches); b.trigger_all(0, vel); f32 = b.render(20001); b.close()
    b = V.VoiceBatch(patches); b.trigger_all(0, vel); pcm = b.render_pcm16(20001); b.close()
    assert np.array_equal(pcm, pcm16_ref(f32))
    assert np.abs(pcm).max() > 1000


def test_block_bounce_equals_per_engine_bounce_and_pcm16():
    n = 10
    es = make(n); ref = G.batch_bounce(es, 1); [e.close() for e in es]
    es = make(n); blk = G.batch_bounce_host(es, 1); [e.close() for e in es]
    es = make(n); pcm = G.batch_bounce_pcm16(es, 1); [e.close() for e in es]
    assert blk.shape == (n, 88200)
    for i in range(n):
        assert np.array_equal(blk[i], ref[i]), i
    assert np.array_equal(pcm, pcm16_ref(blk))


def test_block_bounce_into_numa_placed_pinned_memory():
    n = 6
    hb = HostBuffer(n * 88200 * 4, device=0)
    out = hb.array((n, 88200), np.float32)
    es = make(n); got = G.batch_bounce_host(es, 1, out=out); [e.close() for e in es]
    es = make(n); ref = G.batch_bounce(es, 1); [e.close() for e in es]
    for i in range(n):
        assert np.array_equal(got[i], ref[i])
    print("numa node of the pinned buffer:", hb.numa_node)
    del out, got
    hb.close()


def test_batch_bounce_to_wav_matches_single_engine_files(tmp_path):
    n = 4
    es = make(n)
    paths = [tmp_path / f"b{i}.wav" for i in range(n)]
    G.batch_bounce_to_wav(es, 1, paths)
    [e.close() for e in es]
    for i in range(n):
        e = G.Engine(); script(e, i)
        assert e.bounce_to_wav(1, tmp_path / f"s{i}.wav")
        e.close()
        assert (tmp_path / f"b{i}.wav").read_bytes() == (tmp_path / f"s{i}.wav").read_bytes(), i


def test_one_batch_over_two_devices():
    if lib().gooey_b200_device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 8
    es = []
    for i in range(n):
        G.set_device(i % 2 if i < 4 else (0 if i < 6 else 1))      # first half interleaved (non-contiguous groups), second half contiguous
        e = G.Engine(); script(e, i); es.append(e)
    blk = G.batch_bounce_host(es, 1)
    pcm_engines = []
    for i in range(n):
        G.set_device(1 - (i % 2))
        e = G.Engine(); script(e, i); pcm_engines.append(e)
    pcm = G.batch_bounce_pcm16(pcm_engines, 1)
    [e.close() for e in es + pcm_engines]
    G.set_device(0)
    es = make(n); ref = G.batch_bounce(es, 1); [e.close() for e in es]
    for i in range(n):
        assert np.array_equal(blk[i], ref[i]), i
    assert np.array_equal(pcm, pcm16_ref(blk))
