"""Sample-playback sources on the CPU.

1. The reference's own unit tests of mixer/loop_channel.rs, mixer/stereo_buffer.rs, mixer/mod.rs and instruments/sampler.rs,
   restated against the oracle (oracle/loops.hpp) — the pins this path has (the reference holds no golden audio for it).
2. The product's device functions (libgooey_b200/csrc/loops.cuh, compiled for the host by tests/emu) against the oracle's
   loop mixer and sampler rack, bit for bit, on random playback states: forward / reverse varispeed, sample-rate conversion,
   interior and wrap-around loop windows, gliding faders and gates, tempo warp, layered and expiring sampler voices.
"""
import ctypes

import numpy as np
import pytest

import emu_lib as EMU
import oracle_lib as O

SR = 44100.0
c = ctypes


def ramp(frames):
    return np.arange(frames, dtype=np.float32)


def dc(value, frames):
    return np.full(frames, value, np.float32)


def tick(e, frames):
    """LoopMixer::tick on its own, `frames` times: [frames, 2]."""
    out = np.zeros((frames, 2), np.float32)
    L = O.lib()
    L.orc_engine_loop_mixer_tick.argtypes = [c.c_void_p, c.c_uint32, c.c_void_p]
    L.orc_engine_loop_mixer_tick.restype = None
    L.orc_engine_loop_mixer_tick(e._h, frames, out.ctypes.data)
    return out


def rack_tick(e, rack, frames):
    out = np.zeros((frames, 2), np.float32)
    L = O.lib()
    L.orc_engine_sampler_rack_tick.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32, c.c_void_p]
    L.orc_engine_sampler_rack_tick.restype = None
    L.orc_engine_sampler_rack_tick(e._h, rack, frames, out.ctypes.data)
    return out


def cursor(e, ch):
    L = O.lib()
    L.orc_engine_loop_get_cursor.argtypes = [c.c_void_p, c.c_uint32]
    L.orc_engine_loop_get_cursor.restype = c.c_double
    return L.orc_engine_loop_get_cursor(e._h, ch)


@pytest.fixture
def e():
    eng = O.oracle_engine()
    yield eng
    eng.close()


# ---- mixer/loop_channel.rs mod tests ------------------------------------------------------------------------------------
def test_silent_until_playing(e):                                   # :634
    assert e.loop_load(0, ramp(100), SR)
    assert np.array_equal(tick(e, 1), np.zeros((1, 2), np.float32))


def test_cursor_wraps_within_loop_window(e):                        # :642
    e.loop_load(0, ramp(10), SR)
    e.loop_set_start(0, 0.0); e.loop_set_end(0, 0.5); e.loop_set_playing(0, True)
    for _ in range(50):
        tick(e, 1)
        assert e.loop_get_position(0) < 0.5 + 1e-3


def test_set_buffer_starts_at_loop_start(e):                        # :656
    e.loop_set_start(0, 0.5); e.loop_set_end(0, 1.0)
    e.loop_load(0, ramp(100), SR)
    assert abs(e.loop_get_position(0) - 0.5) < 1e-3


def test_set_position_round_trips_and_clamps_into_window(e):        # :666, :674
    e.loop_load(0, ramp(100), SR)
    e.loop_set_position(0, 0.42)
    assert abs(e.loop_get_position(0) - 0.42) < 1e-2
    e.loop_set_start(0, 0.25); e.loop_set_end(0, 0.75)
    e.loop_set_position(0, 0.9)
    assert e.loop_get_position(0) <= 0.75 + 1e-6
    e.loop_set_position(0, 0.1)
    assert e.loop_get_position(0) >= 0.25 - 1e-6


def test_reverse_playback_stays_in_window(e):                       # :687
    e.loop_load(0, ramp(20), SR)
    e.loop_set_start(0, 0.25); e.loop_set_end(0, 0.75); e.loop_set_speed(0, -1.0); e.loop_set_playing(0, True)
    for _ in range(100):
        tick(e, 1)
        assert 0.25 <= e.loop_get_position(0) < 0.75 + 1e-2


def settle_then(e, script, frames):
    """The fader and the gate start settled at 1, so the first ticks are the raw reads."""
    script()
    return tick(e, frames)


def test_wrapped_window_plays_union(e):                             # :792 — exact values
    e.loop_set_start(0, 0.75); e.loop_set_end(0, 0.25)
    e.loop_load(0, ramp(8), SR)
    e.loop_set_playing(0, True)
    out = tick(e, 16)
    assert np.array_equal(out[:, 0], np.tile(np.float32([6, 7, 0, 1]), 4))


def test_non_wrapped_interior_window_exact_sequence(e):             # :808 — exact values
    e.loop_set_start(0, 0.25); e.loop_set_end(0, 0.5)
    e.loop_load(0, ramp(8), SR)
    e.loop_set_playing(0, True)
    out = tick(e, 16)
    assert np.array_equal(out[:, 0], np.tile(np.float32([2, 3]), 8))


def test_wrapped_window_reverse_stays_in_union(e):                  # :824
    e.loop_set_start(0, 0.7); e.loop_set_end(0, 0.3)
    e.loop_load(0, ramp(10), SR)
    e.loop_set_speed(0, -1.0); e.loop_set_playing(0, True)
    for _ in range(200):
        tick(e, 1)
        p = e.loop_get_position(0)
        assert not (0.3 + 1e-3 <= p <= 0.7 - 1e-3), p


def test_set_position_folds_into_wrapped_window(e):                 # :842
    e.loop_set_start(0, 0.7); e.loop_set_end(0, 0.3)
    e.loop_load(0, ramp(10), SR)
    for pos, want in ((0.1, 0.1), (0.9, 0.9), (0.45, 0.3), (0.55, 0.7)):
        e.loop_set_position(0, pos)
        assert abs(e.loop_get_position(0) - want) < 1e-6


def test_degenerate_wrapped_window_is_clamped(e):                   # :874
    e.loop_set_start(0, 0.9); e.loop_set_end(0, 0.1)
    e.loop_load(0, ramp(4), SR)
    e.loop_set_playing(0, True)
    out = tick(e, 200)
    assert np.isfinite(out).all() and np.isfinite(e.loop_get_position(0))


# ---- mixer/stereo_buffer.rs mod tests -----------------------------------------------------------------------------------
def test_from_interleaved_mono_and_stereo(e):                       # :267, :276
    e.loop_load(0, np.float32([0.1, 0.2, 0.3]), SR)
    e.loop_set_playing(0, True)
    out = tick(e, 3)
    assert np.array_equal(out[:, 0], out[:, 1]) and np.allclose(out[:, 0], [0.1, 0.2, 0.3], atol=1e-7)
    st = np.float32([[1.0, -1.0], [0.5, -0.5]])
    e.loop_load(1, st, SR)
    e.loop_set_playing(0, False); e.loop_set_playing(1, True)
    out = tick(e, 2)
    assert np.array_equal(out, st)


def test_non_finite_and_bad_rate_rejected(e):                       # :293
    assert not e.loop_load(0, np.float32([0.0, np.nan]), SR)
    assert not e.loop_load(0, np.float32([0.0, 1.0]), 0.0)
    assert not e.loop_load(7, np.float32([0.0, 1.0]), SR)


# ---- mixer/mod.rs mod tests ---------------------------------------------------------------------------------------------
def test_muted_channel_drops_out_after_smoothing(e):                # :497
    e.loop_load(0, dc(0.5, 64), SR); e.loop_set_playing(0, True)
    tick(e, 4096)
    assert abs(tick(e, 1)[0, 0]) > 0.1
    e.loop_set_mute(0, True)
    tick(e, 4096)
    assert abs(tick(e, 1)[0, 0]) < 5e-3


def test_solo_silences_other_channels(e):                           # :519
    for ch in (0, 1):
        e.loop_load(ch, dc(0.5, 64), SR); e.loop_set_playing(ch, True)
    e.loop_set_solo(0, True)
    tick(e, 4096)
    assert abs(tick(e, 1)[0, 0] - 0.5) < 0.05


def test_render_channel_frame_count_rejects_and_region(e):          # :555, :564, :576
    e.loop_load(0, dc(0.5, 4096), SR)
    assert e.loop_render(0, 1000, 512).shape == (1000, 2)
    assert e.loop_render(99, 100) is None and e.loop_render(1, 100) is None
    e.loop_load(0, ramp(400), SR)
    e.loop_set_start(0, 0.0); e.loop_set_end(0, 0.25)
    out = e.loop_render(0, 350)
    assert np.abs(out[:, 0] - (np.arange(350) % 100)).max() < 1e-3


def test_render_channel_applies_gain_and_ignores_mute_solo(e):      # :597, :612
    e.loop_load(0, dc(0.5, 4096), SR)
    e.loop_set_gain(0, 0.5)
    assert abs(e.loop_render(0, 256, 128)[0, 0] - 0.25) < 1e-3
    e.loop_set_gain(0, 1.0)
    e.loop_load(1, dc(-0.9, 4096), SR)
    e.loop_set_mute(0, True); e.loop_set_solo(1, True)
    assert abs(e.loop_render(0, 256, 128)[0, 0] - 0.5) < 1e-3


# ---- instruments/sampler.rs mod tests -----------------------------------------------------------------------------------
def test_sampler_stereo_buffer_is_interpolated_and_preserved(e):    # :335 (through a voice: position 0.5 needs increment 0.5)
    r = e.sampler_register()
    assert r == 0 and e.sampler_get_source_id(0) == 5 and e.sampler_get_source_id(1) == 0xFFFFFFFF
    long = np.zeros((200, 2), np.float32); long[:, 0] = np.arange(200); long[:, 1] = -np.arange(200)
    assert e.sampler_set_slot_buffer(0, 3, long, SR / 2)
    assert e.sampler_slot_is_loaded(0, 3) and e.sampler_slot_frames(0, 3) == 200 and e.sampler_slot_channels(0, 3) == 2
    assert e.sampler_slot_sample_rate(0, 3) == SR / 2
    assert e.sampler_trigger(0, 3, 1.0) and not e.sampler_trigger(0, 4, 1.0)
    out = rack_tick(e, 0, 101)
    # past the 32-frame fade-in the voice is the linear read at position f / 2
    assert np.allclose(out[80:101, 0], np.arange(80, 101) / 2, atol=1e-5) and np.allclose(out[80:101, 1], -np.arange(80, 101) / 2, atol=1e-5)


def test_rack_layers_and_steals_without_non_finite_audio(e):        # :343
    e.sampler_register()
    e.sampler_set_slot_buffer(0, 0, dc(0.5, 256), 22050.0)
    for _ in range(32 + 4):
        assert e.sampler_trigger(0, 0, 1.0)
    assert np.isfinite(rack_tick(e, 0, 32)).all()


# ---- device functions (host build) against the oracle, bit for bit --------------------------------------------------------
def emu_loop_mixer(chans, engine_sr, frames):
    """chans: 4 dicts (or None) with left, right, buf_sr, cursor, warp, start, end, speed, playing, gain (c, t), active (c, t)."""
    L = EMU.lib()
    fp = c.POINTER(c.c_float)
    left = (fp * 4)(); right = (fp * 4)()
    ln = np.zeros(4, np.uint32); bsr = np.zeros(4, np.float32); cur = np.zeros(4, np.float64); warp = np.ones(4, np.float64)
    st = np.zeros(4, np.float32); en = np.ones(4, np.float32); sp = np.ones(4, np.float32); pl = np.zeros(4, np.uint32)
    g = np.ones((4, 2), np.float32); a = np.ones((4, 2), np.float32)
    pres = np.zeros(4, np.uint32); wpp = np.ones(4, np.float64)
    keep = []
    for k, ch in enumerate(chans):
        if ch is None:
            continue
        if ch.get("left") is not None:
            l = np.ascontiguousarray(ch["left"], np.float32); r = np.ascontiguousarray(ch["right"], np.float32)
            keep += [l, r]
            left[k] = l.ctypes.data_as(fp); right[k] = r.ctypes.data_as(fp); ln[k] = len(l)
        bsr[k] = ch.get("buf_sr", 0.0); cur[k] = ch.get("cursor", 0.0); warp[k] = ch.get("warp", 1.0)
        st[k] = ch.get("start", 0.0); en[k] = ch.get("end", 1.0); sp[k] = ch.get("speed", 1.0); pl[k] = 1 if ch.get("playing") else 0
        g[k] = ch.get("gain", (1.0, 1.0)); a[k] = ch.get("active", (1.0, 1.0))
        pres[k] = 1 if ch.get("preserve") else 0; wpp[k] = ch.get("warp_pp", 1.0)
    ol = np.zeros(frames, np.float32); orr = np.zeros(frames, np.float32)
    p = lambda arr, t: arr.ctypes.data_as(c.POINTER(t))
    L.emu_loop_mixer(left, right, p(ln, c.c_uint32), p(bsr, c.c_float), p(cur, c.c_double), p(warp, c.c_double), p(st, c.c_float), p(en, c.c_float),
                     p(sp, c.c_float), p(pl, c.c_uint32), p(g, c.c_float), p(a, c.c_float), c.c_float(engine_sr), frames, p(ol, c.c_float), p(orr, c.c_float),
                     p(pres, c.c_uint32), p(wpp, c.c_double))
    return np.stack([ol, orr], 1), cur, g, a


def window_lo(start, end, n):
    L = EMU.lib()
    L.emu_window_lo.restype = c.c_double
    L.emu_window_lo.argtypes = [c.c_float, c.c_float, c.c_double]
    return L.emu_window_lo(start, end, float(n))


@pytest.mark.parametrize("seed", range(12))
def test_device_loop_mixer_matches_oracle_bit_for_bit(seed):
    rng = np.random.default_rng(1000 + seed)
    engine_sr = [44100.0, 48000.0, 22050.0][seed % 3]
    o = O.oracle_engine(engine_sr)
    chans = [None] * 4
    for k in range(4):
        if rng.random() < 0.2:
            continue                                   # channel left empty
        n = int(rng.integers(1, 3000)) if rng.random() < 0.9 else 1
        pcm = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        buf_sr = float(rng.choice([44100.0, 48000.0, 32000.0, 96000.0]))
        assert o.loop_load(k, pcm, buf_sr)
        start, end = (float(np.float32(x)) for x in rng.uniform(0, 1, 2))
        if rng.random() < 0.3:
            start, end = 0.0, 1.0
        speed = float(np.float32(rng.uniform(-4, 4))) if rng.random() < 0.7 else 1.0
        o.loop_set_start(k, start); o.loop_set_end(k, end); o.loop_set_speed(k, speed)
        o.loop_restart(k)
        playing = rng.random() < 0.85
        o.loop_set_playing(k, bool(playing))
        gain_t = float(np.float32(rng.uniform(0, 2))) if rng.random() < 0.6 else 1.0
        o.loop_set_gain(k, gain_t)
        muted = rng.random() < 0.25
        o.loop_set_mute(k, bool(muted))
        warp = 1.0
        if rng.random() < 0.4:
            src_bpm = float(np.float32(rng.uniform(70, 170)))
            o.loop_set_source_bpm(k, src_bpm); o.loop_set_pitch_mode(k, 1)
            warp = float(np.float32(133.0)) / src_bpm
        chans[k] = dict(left=pcm[:, 0], right=pcm[:, 1], buf_sr=buf_sr, cursor=window_lo(start, end, n), warp=warp, start=start, end=end, speed=speed,
                        playing=playing, gain=(1.0, gain_t), active=(1.0, 0.0 if muted else 1.0))
    o.set_bpm(133.0)
    frames = 6000
    want = tick(o, frames)
    got, cur, g, a = emu_loop_mixer([ch if ch is not None else None for ch in chans], engine_sr, frames)
    assert np.array_equal(got, want), np.abs(got - want).max()
    for k in range(4):
        if chans[k] is not None:
            assert cur[k] == cursor(o, k)
    o.close()


def test_device_loop_mixer_second_call_continues_state():
    """Two calls with the state carried over equal one long call (what engines_render does piece by piece and call by call)."""
    rng = np.random.default_rng(7)
    pcm = rng.uniform(-1, 1, (777, 2)).astype(np.float32)
    ch = dict(left=pcm[:, 0], right=pcm[:, 1], buf_sr=48000.0, cursor=window_lo(0.8, 0.3, 777), start=0.8, end=0.3, speed=-1.37, playing=True,
              gain=(1.0, 0.4), active=(1.0, 0.0))
    whole, _, _, _ = emu_loop_mixer([ch, None, None, None], 44100.0, 5000)
    a, cur, g, act = emu_loop_mixer([ch, None, None, None], 44100.0, 1234)
    ch2 = dict(ch, cursor=cur[0], gain=tuple(g[0]), active=tuple(act[0]))
    b, _, _, _ = emu_loop_mixer([ch2, None, None, None], 44100.0, 5000 - 1234)
    assert np.array_equal(np.concatenate([a, b]), whole)


@pytest.mark.parametrize("seed", range(6))
def test_device_sampler_rack_matches_oracle_bit_for_bit(seed):
    rng = np.random.default_rng(2000 + seed)
    engine_sr = [44100.0, 48000.0][seed % 2]
    o = O.oracle_engine(engine_sr)
    assert o.sampler_register() == 0
    pads = []
    for slot in range(5):
        n = int(rng.integers(1, 1500))
        chn = int(rng.integers(1, 3))
        pcm = rng.uniform(-1, 1, (n, chn)).astype(np.float32)
        sr = float(rng.choice([44100.0, 22050.0, 48000.0, 11025.0]))
        assert o.sampler_set_slot_buffer(0, slot, pcm, sr)
        pads.append((pcm, sr))
    n_hits = int(rng.integers(1, 33))                      # all at the start of the call: voices 0 .. n_hits - 1 in order
    hits = [(int(rng.integers(0, 5)), float(np.float32(rng.uniform(-0.2, 1.3)))) for _ in range(n_hits)]
    for slot, vel in hits:
        assert o.sampler_trigger(0, slot, vel)
    frames = 4000
    want = rack_tick(o, 0, frames)
    L = EMU.lib()
    fp = c.POINTER(c.c_float)
    ptrs = (fp * n_hits)()
    fr = np.zeros(n_hits, np.uint32); chs = np.zeros(n_hits, np.uint32); inc = np.zeros(n_hits, np.float64); vel = np.zeros(n_hits, np.float32)
    pos = np.zeros(n_hits, np.float64)
    for v, (slot, ve) in enumerate(hits):
        pcm, sr = pads[slot]
        ptrs[v] = pcm.ctypes.data_as(fp); fr[v] = pcm.shape[0]; chs[v] = pcm.shape[1]
        inc[v] = float(np.float32(sr)) / float(np.float32(engine_sr)); vel[v] = min(max(ve, 0.0), 1.0)
    ol = np.zeros(frames, np.float32); orr = np.zeros(frames, np.float32)
    p = lambda arr, t: arr.ctypes.data_as(c.POINTER(t))
    alive = L.emu_sampler_rack(n_hits, ptrs, p(fr, c.c_uint32), p(chs, c.c_uint32), p(inc, c.c_double), p(vel, c.c_float), p(pos, c.c_double), frames,
                               p(ol, c.c_float), p(orr, c.c_float))
    got = np.stack([ol, orr], 1)
    assert np.array_equal(got, want), np.abs(got - want).max()
    O.lib().orc_engine_sampler_active_voices.argtypes = [c.c_void_p, c.c_uint32]
    assert alive == O.lib().orc_engine_sampler_active_voices(o._h, 0)
    assert np.abs(want).max() > 0.05
    o.close()


def test_loop_source_reaches_the_engine_mix_on_the_loops_track():
    """ffi.rs:1296-1308: the loop mixer is graph source 4 (default track 3); a sampler rack sounds once it is routed."""
    o = O.oracle_engine()
    silent = o.render(256)
    assert np.abs(silent).max() == 0.0
    o.loop_load(0, dc(0.5, 64), SR); o.loop_set_playing(0, True)
    out = o.render(2048)
    assert abs(out[-1, 0] - 0.5 * 0.25) < 1e-3                 # unity strip, centre balance, master gain 0.25
    o.mixer_set_track_mute(3, True)
    assert abs(o.render(4096)[-1, 0]) < 1e-3
    assert o.sampler_register() == 0
    o.sampler_set_slot_buffer(0, 0, dc(0.8, 4000), SR)
    assert not o.mixer_route_source(6, 0)                      # rack 1 is not registered
    o.sampler_trigger(0, 0, 1.0)
    assert np.abs(o.render(512)).max() < 1e-3                  # registered racks start unrouted
    assert o.mixer_route_source(5, 0)
    out = o.render(512)
    assert abs(out[-1, 0] - 0.8 * 0.25) < 1e-3
    o.close()


# ---- mixer/wsola.rs: PitchMode::PreservePitch ------------------------------------------------------------------------------
def sine_pcm(seconds, hz, sr):
    n = int(np.float32(sr) * np.float32(seconds))
    x = np.sin((np.arange(n, dtype=np.float32) / np.float32(sr) * np.float32(hz) * np.float32(2 * np.pi)).astype(np.float32)).astype(np.float32)
    return x


def test_wsola_hop_window_and_constant_overlap_add():               # wsola.rs:480, :488 (through the product's table builder)
    L = EMU.lib()
    L.emu_wsola_window.restype = c.c_int
    L.emu_wsola_window.argtypes = [c.c_float, c.c_void_p, c.c_int]
    w = np.zeros(4096, np.float32)
    hop = L.emu_wsola_window(48000.0, w.ctypes.data, 4096)
    assert hop == 960
    assert np.abs(w[:hop] + w[hop:2 * hop] - 1.0).max() < 1e-4


def test_wsola_produces_finite_output_and_warp_consumes_source_faster(e):   # wsola.rs:497, :510 through the channel
    slow, fast = O.oracle_engine(48000.0), O.oracle_engine(48000.0)
    for eng, bpm in ((slow, 100.0), (fast, 200.0)):
        eng.loop_load(0, sine_pcm(4.0, 220.0, 48000.0), 48000.0)
        eng.loop_set_source_bpm(0, 100.0); eng.loop_set_pitch_mode(0, 2); eng.set_bpm(bpm); eng.loop_set_playing(0, True)
        out = tick(eng, 20 * 960)
        assert np.isfinite(out).all() and np.abs(out).max() > 0.5
    assert cursor(fast, 0) > cursor(slow, 0)
    slow.close(); fast.close()


def test_preserve_pitch_wrapped_window_is_finite_and_bounded(e):    # loop_channel.rs:911
    e.loop_set_start(0, 0.75); e.loop_set_end(0, 0.25)
    e.loop_load(0, sine_pcm(1.0, 220.0, SR), SR)
    e.loop_set_source_bpm(0, 120.0); e.loop_set_pitch_mode(0, 2); e.set_bpm(150.0); e.loop_set_playing(0, True)
    out = tick(e, 20000)
    assert np.isfinite(out).all() and np.abs(out).max() < 2.0
    p = e.loop_get_position(0)
    assert p >= 0.75 - 1e-6 or p < 0.25 + 1e-6


@pytest.mark.parametrize("seed", range(6))
def test_device_wsola_matches_oracle_bit_for_bit(seed):
    rng = np.random.default_rng(3000 + seed)
    engine_sr = [44100.0, 48000.0, 22050.0][seed % 3]
    o = O.oracle_engine(engine_sr)
    n = int(rng.integers(3000, 20000))
    t = np.arange(n) / 44100.0
    pcm = np.stack([0.5 * np.sin(2 * np.pi * 180 * t) + 0.2 * rng.uniform(-1, 1, n), 0.5 * np.sin(2 * np.pi * 271 * t + 1) + 0.2 * rng.uniform(-1, 1, n)], 1).astype(np.float32)
    buf_sr = float(rng.choice([44100.0, 48000.0, 32000.0]))
    start, end = (0.0, 1.0) if seed % 4 == 0 else (float(np.float32(x)) for x in rng.uniform(0, 1, 2))
    if seed == 5:
        start, end = 0.5, 0.52                             # too small for a grain: every hop restarts at the loop start
    speed = float(np.float32(rng.uniform(0.3, 2.0))) if seed % 2 else 1.0
    src_bpm = float(np.float32(rng.uniform(80, 160)))
    assert o.loop_load(0, pcm, buf_sr)
    o.loop_set_start(0, start); o.loop_set_end(0, end); o.loop_set_speed(0, speed); o.loop_restart(0)
    o.loop_set_source_bpm(0, src_bpm); o.loop_set_pitch_mode(0, 2); o.set_bpm(128.0); o.loop_set_playing(0, True)
    frames = 6000
    want = tick(o, frames)
    ch = dict(left=pcm[:, 0], right=pcm[:, 1], buf_sr=buf_sr, cursor=window_lo(start, end, n), start=start, end=end, speed=speed, playing=True,
              preserve=True, warp_pp=float(np.float32(128.0)) / src_bpm)
    got, cur, _, _ = emu_loop_mixer([ch, None, None, None], engine_sr, frames)
    assert np.abs(want).max() > 0.05
    assert np.array_equal(got, want), np.abs(got - want).max()
    assert cur[0] == cursor(o, 0)
    o.close()


# ---- queued swaps (loop_channel.rs:413-423, 249-276) ------------------------------------------------------------------------
def test_queued_swap_lands_at_first_division_boundary(e):          # :708
    e.loop_load(0, ramp(100), SR); e.loop_set_playing(0, True)
    assert e.loop_queue_swap(0, dc(1000.0, 100), SR, divisions=4)
    out = tick(e, 48)[:, 0]
    idx = int(np.argmax(out > 500.0))
    assert out[idx] > 500.0 and 24 <= idx <= 27
    assert e.loop_swaps_completed(0) == 1


def test_queued_swap_divisions_one_only_swaps_at_wrap(e):          # :734
    e.loop_load(0, ramp(100), SR); e.loop_set_playing(0, True)
    e.loop_queue_swap(0, dc(1000.0, 100), SR, divisions=1)
    assert (tick(e, 90)[:, 0] < 500.0).all() and e.loop_swaps_completed(0) == 0
    tick(e, 20)
    assert e.loop_swaps_completed(0) == 1


def test_cancel_drops_and_requeue_replaces_the_queued_swap(e):     # :754, :768
    e.loop_load(0, ramp(100), SR); e.loop_set_playing(0, True)
    e.loop_queue_swap(0, dc(1000.0, 100), SR, divisions=4)
    e.loop_cancel_queued_swap(0)
    assert (tick(e, 150)[:, 0] < 500.0).all() and e.loop_swaps_completed(0) == 0
    e.loop_restart(0)
    e.loop_queue_swap(0, dc(-1000.0, 100), SR, divisions=4)
    e.loop_queue_swap(0, dc(2000.0, 100), SR, divisions=4)
    tick(e, 40)
    assert e.loop_swaps_completed(0) == 1 and tick(e, 1)[0, 0] > 1500.0


def test_queued_swap_lands_in_wrapped_window(e):                   # :889
    e.loop_set_start(0, 0.7); e.loop_set_end(0, 0.3)
    e.loop_load(0, ramp(10), SR); e.loop_set_playing(0, True)
    e.loop_queue_swap(0, dc(1000.0, 10), SR, divisions=1)
    tick(e, 5)
    assert e.loop_swaps_completed(0) == 0
    tick(e, 4)
    assert e.loop_swaps_completed(0) == 1 and tick(e, 1)[0, 0] > 500.0


@pytest.mark.parametrize("seed", range(9))
def test_device_queued_swap_matches_oracle_bit_for_bit(seed):
    """The take lands on the same sample, the phrase restarts on the new buffer with ITS rate and tempo tag: direct, wrapped, Resample and
    PreservePitch channels."""
    rng = np.random.default_rng(4000 + seed)
    engine_sr = 44100.0
    mode = seed % 3                                             # pitch mode
    o = O.oracle_engine(engine_sr)
    # (a PreservePitch loop too short for a second grain restarts at its loop start every hop: the cursor never moves and the take never
    # lands — the reference's behaviour, so the stretched cases get a longer phrase and a grid of at least two divisions)
    n = int(rng.integers(9000, 14000)) if mode == 2 else int(rng.integers(2500, 6000))
    a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    nb = int(rng.integers(2500, 6000))
    b = (2.0 + rng.uniform(-1, 1, (nb, 2))).astype(np.float32)
    sr_a, sr_b = float(rng.choice([44100.0, 48000.0])), float(rng.choice([44100.0, 32000.0]))
    start, end = ((0.0, 1.0), (0.1, 0.8), (0.75, 0.3))[seed // 3]
    speed = 1.0 if mode == 2 else float(np.float32(rng.uniform(0.5, 1.5))) * (-1.0 if seed == 4 else 1.0)
    bpm_a, bpm_b = float(np.float32(rng.uniform(90, 150))), float(np.float32(rng.uniform(90, 150)))
    div = int(rng.integers(2 if mode == 2 else 1, 6))
    assert o.loop_load(0, a, sr_a)
    o.loop_set_start(0, start); o.loop_set_end(0, end); o.loop_set_speed(0, speed); o.loop_restart(0)
    o.loop_set_source_bpm(0, bpm_a); o.loop_set_pitch_mode(0, mode); o.set_bpm(120.0); o.loop_set_playing(0, True)
    assert o.loop_queue_swap(0, b, sr_b, bpm_b, div)
    frames = 9000
    want = tick(o, frames)
    assert o.loop_swaps_completed(0) == 1 and want[:, 0].max() > 1.5
    L = EMU.lib()
    fp = c.POINTER(c.c_float)
    bl, br = np.ascontiguousarray(b[:, 0]), np.ascontiguousarray(b[:, 1])
    ratio_a, ratio_b = float(np.float32(120.0)) / bpm_a, float(np.float32(120.0)) / bpm_b
    L.emu_loop_queue.argtypes = [c.c_int, fp, fp, c.c_uint32, c.c_float, c.c_double, c.c_double, c.c_uint32]
    L.emu_loop_queue(0, bl.ctypes.data_as(fp), br.ctypes.data_as(fp), nb, sr_b, ratio_b if mode == 1 else 1.0, ratio_b if mode else 1.0, div)
    ch = dict(left=a[:, 0], right=a[:, 1], buf_sr=sr_a, cursor=window_lo(start, end, n), start=start, end=end, speed=speed, playing=True,
              warp=ratio_a if mode == 1 else 1.0, preserve=mode == 2, warp_pp=ratio_a if mode else 1.0)
    got, cur, _, _ = emu_loop_mixer([ch, None, None, None], engine_sr, frames)
    L.emu_loop_swaps.restype = c.c_uint32
    assert L.emu_loop_swaps(0) == 1
    assert np.array_equal(got, want), (np.abs(got - want).max(), int(np.argmax((got != want).any(1))))
    assert cur[0] == cursor(o, 0)
    o.close()


# ---- sampler-rack step patterns on the transport (sampler.rs:232-310, ffi.rs:1139-1147, 1199-1210, clip_grid.rs:174-191) -------------
def oracle_hits(o, rack=0, cap=4096):
    L = O.lib()
    u32, f32 = c.c_uint32, c.c_float
    fr, sl, ve = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.float32)
    L.orc_engine_sampler_hit_log.argtypes = [c.c_void_p, u32, c.POINTER(u32), c.POINTER(u32), c.POINTER(f32), u32]
    L.orc_engine_sampler_hit_log.restype = u32
    n = L.orc_engine_sampler_hit_log(o._h, rack, fr.ctypes.data_as(c.POINTER(u32)), sl.ctypes.data_as(c.POINTER(u32)), ve.ctypes.data_as(c.POINTER(f32)), cap)
    return [(int(fr[i]), int(sl[i]), float(ve[i])) for i in range(n)]


def product_schedule(sr, bpm, swing, steps, transport_running, transport_beat, pending_beat, bounce, calls, cap=4096):
    """gooey_b200_sampler_schedule: the host code engines_render runs for a rack pattern (no device needed)."""
    from libgooey_b200._lib import lib
    L = lib()
    u8, u32, f32 = c.c_uint8, c.c_uint32, c.c_float
    en = (u8 * 16)(*[1 if s[0] else 0 for s in steps]); pd = (u8 * 16)(*[s[1] for s in steps]); ve = (f32 * 16)(*[s[2] for s in steps])
    cl = (u32 * len(calls))(*calls)
    of, op, ov = (u32 * cap)(), (u32 * cap)(), (f32 * cap)()
    beat = c.c_double(0.0)
    L.gooey_b200_sampler_schedule.restype = u32
    L.gooey_b200_sampler_schedule.argtypes = [f32, f32, f32, c.POINTER(u8), c.POINTER(u8), c.POINTER(f32), c.c_int, c.c_double, c.c_double, c.c_int,
                                              c.POINTER(u32), u32, c.POINTER(u32), c.POINTER(u32), c.POINTER(f32), u32, c.POINTER(c.c_double)]
    n = L.gooey_b200_sampler_schedule(sr, bpm, swing, en, pd, ve, int(transport_running), transport_beat, pending_beat, int(bounce), cl, len(calls),
                                      of, op, ov, cap, c.byref(beat))
    return [(int(of[i]), int(op[i]), float(ov[i])) for i in range(n)], beat.value


def random_steps(rng):
    return [(bool(rng.random() < 0.5), int(rng.integers(0, 16)), float(np.float32(rng.uniform(0.1, 1.0)))) for _ in range(16)]


@pytest.mark.parametrize("seed", range(8))
def test_rack_pattern_schedule_matches_the_oracle_exactly(seed):
    """Armed start on the running transport, quantised to the next sixteenth / quarter / bar after some rendering; hits over several calls."""
    rng = np.random.default_rng(5000 + seed)
    sr = [44100.0, 48000.0][seed % 2]
    bpm = float(np.float32(rng.uniform(70, 180)))
    swing = 0.5 if seed % 3 else float(np.float32(rng.uniform(0.4, 0.7)))
    steps = random_steps(rng)
    steps[0] = (True, steps[0][1], steps[0][2])
    o = O.oracle_engine(sr)
    if seed % 2:
        o.set_bpm(bpm)                                           # a rack registered later starts at the engine's current tempo (ffi.rs:6014-6018)
    assert o.sampler_register() == 0
    o.set_bpm(bpm); o.set_swing(swing)
    for i, (en, pad, vel) in enumerate(steps):
        assert o.sampler_set_step(0, i, en, pad, vel)
    assert o.sampler_get_step(0, 3) == (steps[3][0], steps[3][1], pytest.approx(steps[3][2]))
    pre = int(rng.integers(0, 30000))                        # frames rendered with the transport running before the pattern is armed
    o.sequencer_start()
    if pre:
        o.render(pre)
    q = seed % 3
    assert o.sampler_start_pattern(0, q) and not o.sampler_start_pattern(0, 3) and not o.sampler_start_pattern(1, q)
    pending = o.sampler_get_pending_start_beat(0)
    assert pending > 0.0 and not o.sampler_is_pattern_running(0)
    calls = [int(x) for x in rng.integers(1, 40000, 4)] + [int(4 * 60.0 / bpm * sr) + 20000]      # the last call reaches past a whole bar
    for n in calls:
        o.render(n)
    want = [(f - pre, s, v) for (f, s, v) in oracle_hits(o)]
    L = O.lib(); L.orc_engine_transport_beat.restype = c.c_double; L.orc_engine_transport_beat.argtypes = [c.c_void_p]
    beat_after_pre, _ = product_schedule(sr, bpm, swing, steps, True, 0.0, -1.0, False, [pre] if pre else [1])[1], None
    if not pre:
        beat_after_pre = 0.0
    got, beat = product_schedule(sr, bpm, swing, steps, True, beat_after_pre, pending, False, calls)
    assert len(want) > 2 and o.sampler_is_pattern_running(0) and o.sampler_get_pending_start_beat(0) == -1.0
    assert got == want
    assert beat == L.orc_engine_transport_beat(o._h)
    o.close()


def test_rack_pattern_waits_for_the_transport_and_restarts_with_every_bounce():
    """Armed with the transport stopped -> beat 0 -> fires on the first frame after sequencer_start; a bounce resets the rack's sequencer to step 0
    without touching the transport; stop_pattern / sequencer_stop silence it."""
    rng = np.random.default_rng(77)
    steps = random_steps(rng); steps[0] = (True, 2, 0.9)
    o = O.oracle_engine()
    o.sampler_register()
    for i, (en, pad, vel) in enumerate(steps):
        o.sampler_set_step(0, i, en, pad, vel)
    assert o.sampler_start_pattern(0, 2) and o.sampler_get_pending_start_beat(0) == 0.0
    o.render(3000)                                              # transport stopped: nothing happens
    assert oracle_hits(o) == [] and not o.sampler_is_pattern_running(0)
    o.sequencer_start()
    o.render(20000)
    first = oracle_hits(o)
    assert first[0] == (3000, 2, pytest.approx(0.9)) and o.sampler_is_pattern_running(0)
    got, _ = product_schedule(44100.0, 120.0, 0.5, steps, True, 0.0, 0.0, False, [20000])
    assert got == [(f - 3000, s, v) for (f, s, v) in first]
    n0 = len(first)
    o.bounce_to_buffer(1)                                       # 88 200 frames from step 0
    bounced = [(f - 23000, s, v) for (f, s, v) in oracle_hits(o)[n0:]]
    got, _ = product_schedule(44100.0, 120.0, 0.5, steps, True, 0.0, -1.0, True, [88200])
    assert got == bounced and bounced[0][0] == 0
    assert o.sampler_stop_pattern(0) and not o.sampler_is_pattern_running(0)
    n1 = len(oracle_hits(o))
    o.render(10000)
    assert len(oracle_hits(o)) == n1
    o.close()


@pytest.mark.parametrize("seed", range(4))
def test_device_rack_starts_pattern_hits_like_the_oracle(seed):
    """The device function that starts voices at the hit frames (free voice first, then the oldest), across kernel-like pieces, against the oracle
    playing the same pattern — bit for bit."""
    rng = np.random.default_rng(6000 + seed)
    sr = 44100.0
    o = O.oracle_engine(sr)
    o.set_bpm(170.0)
    o.sampler_register()
    pads = []
    for slot in range(6):
        n = int(rng.integers(200, 6000)) if seed != 3 else int(rng.integers(150000, 190000))      # seed 3: pads outlive 32 later hits -> stealing
        chn = int(rng.integers(1, 3))
        pcm = rng.uniform(-1, 1, (n, chn)).astype(np.float32)
        psr = float(rng.choice([44100.0, 22050.0, 48000.0]))
        o.sampler_set_slot_buffer(0, slot, pcm, psr)
        pads.append((pcm, psr))
    steps = [(bool(rng.random() < 0.8), int(rng.integers(0, 7)), float(np.float32(rng.uniform(0.2, 1.0)))) for _ in range(16)]   # pad 6 is empty: a hit on it is dropped
    for i, (en, pad, vel) in enumerate(steps):
        o.sampler_set_step(0, i, en, pad, vel)
    o.sampler_start_pattern(0, 0)
    o.sequencer_start()
    piece, n_pieces = 4096, 48
    want = rack_tick_with_engine(o, piece * n_pieces)
    hits = oracle_hits(o)
    L = EMU.lib()
    fp = c.POINTER(c.c_float)
    ptrs = (fp * 6)()
    fr = np.zeros(6, np.uint32); chs = np.zeros(6, np.uint32); inc = np.zeros(6, np.float64)
    for k, (pcm, psr) in enumerate(pads):
        ptrs[k] = pcm.ctypes.data_as(fp); fr[k] = pcm.shape[0]; chs[k] = pcm.shape[1]; inc[k] = float(np.float32(psr)) / float(np.float32(sr))
    hf = np.array([h[0] for h in hits], np.uint32); hs = np.array([h[1] for h in hits], np.uint32); hv = np.array([h[2] for h in hits], np.float32)
    ol = np.zeros(piece * n_pieces, np.float32); orr = np.zeros(piece * n_pieces, np.float32)
    p = lambda arr, t: arr.ctypes.data_as(c.POINTER(t))
    L.emu_sampler_rack_hits(6, ptrs, p(fr, c.c_uint32), p(chs, c.c_uint32), p(inc, c.c_double), len(hits), p(hf, c.c_uint32), p(hs, c.c_uint32), p(hv, c.c_float),
                            piece, n_pieces, p(ol, c.c_float), p(orr, c.c_float))
    got = np.stack([ol, orr], 1)
    assert len(hits) > 30 and np.abs(want).max() > 0.1
    assert np.array_equal(got, want), np.abs(got - want).max()
    o.close()


def rack_tick_with_engine(o, frames):
    """Rack 0's own stereo frames while the ENGINE renders `frames` frames (so that the pattern and the transport run)."""
    L = O.lib()
    L.orc_engine_capture_rack0.argtypes = [c.c_void_p, c.c_bool]; L.orc_engine_capture_rack0.restype = None
    L.orc_engine_rack0_capture.argtypes = [c.c_void_p, c.c_void_p, c.c_uint32]; L.orc_engine_rack0_capture.restype = c.c_uint32
    L.orc_engine_capture_rack0(o._h, True)
    o.render(frames)
    out = np.zeros((frames, 2), np.float32)
    assert L.orc_engine_rack0_capture(o._h, out.ctypes.data, frames) == frames
    L.orc_engine_capture_rack0(o._h, False)
    return out


# ---- edge cases of the device functions, bit for bit against the oracle --------------------------------------------------------------
EDGES = [
    dict(n=1, start=0.0, end=1.0, speed=1.0, mode=0),            # one-frame buffer: every read is that frame
    dict(n=1, start=0.0, end=1.0, speed=1.0, mode=2),
    dict(n=2, start=0.0, end=1.0, speed=-4.0, mode=0),
    dict(n=500, start=0.5, end=0.5, speed=1.0, mode=0),          # empty window: span clamps to one frame
    dict(n=500, start=0.5, end=0.5, speed=-0.7, mode=1),
    dict(n=500, start=0.5, end=0.5, speed=1.0, mode=2),
    dict(n=3000, start=0.0, end=1.0, speed=0.0, mode=0),         # frozen cursor
    dict(n=3000, start=0.2, end=0.9, speed=0.0, mode=2),         # WSOLA with a zero native step (floored at 1e-6)
    dict(n=3000, start=1.0, end=0.0, speed=1.3, mode=0),         # start past end at the extremes: the whole buffer as a wrap-around window
    dict(n=3000, start=1.0, end=0.0, speed=1.3, mode=2),
    dict(n=40000, start=0.0, end=1.0, speed=4.0, mode=2),        # fastest varispeed through the stretcher
    dict(n=7, start=0.9, end=0.1, speed=4.0, mode=1),            # window shorter than one step
]


@pytest.mark.parametrize("k", range(len(EDGES)))
def test_device_loop_channel_edge_cases_match_the_oracle(k):
    ed = EDGES[k]
    rng = np.random.default_rng(7000 + k)
    n, start, end, speed, mode = ed["n"], ed["start"], ed["end"], ed["speed"], ed["mode"]
    pcm = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    buf_sr, engine_sr, bpm, src_bpm = 48000.0, 44100.0, 133.0, 97.0
    o = O.oracle_engine(engine_sr)
    assert o.loop_load(0, pcm, buf_sr)
    o.loop_set_start(0, start); o.loop_set_end(0, end); o.loop_set_speed(0, speed); o.loop_restart(0)
    o.loop_set_source_bpm(0, src_bpm); o.loop_set_pitch_mode(0, mode); o.set_bpm(bpm); o.loop_set_gain(0, 0.6); o.loop_set_playing(0, True)
    frames = 5000
    want = tick(o, frames)
    ratio = float(np.float32(bpm)) / float(np.float32(src_bpm))
    ch = dict(left=pcm[:, 0], right=pcm[:, 1], buf_sr=buf_sr, cursor=window_lo(start, end, n), start=start, end=end, speed=speed, playing=True,
              warp=ratio if mode == 1 else 1.0, preserve=mode == 2, warp_pp=ratio if mode else 1.0, gain=(1.0, float(np.float32(0.6))))
    got, cur, _, _ = emu_loop_mixer([ch, None, None, None], engine_sr, frames)
    assert np.isfinite(want).all()
    assert np.array_equal(got, want), (np.abs(got - want).max(), int(np.argmax((got != want).any(1))))
    assert cur[0] == cursor(o, 0)
    o.close()
