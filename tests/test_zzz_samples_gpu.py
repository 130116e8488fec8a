"""Sample-playback sources on the GPU: the loop mixer and the sampler racks of the product, through the C ABI
(gooey_engine_loop_* / gooey_engine_sampler_*), against the oracle running the same call script.  Comparisons with drum voices in
the mix are held to the repo's audio bar (1e-5, the voices set it); where only the sample-playback path sounds the bar is EXACT_TOL:
its arithmetic is IEEE f32 / f64 adds, multiplies, divisions and fmod in the reference's order (tests/test_samples_cpu.py holds the
same device functions, compiled for the host, to bit-identity with the oracle), so anything above rounding noise is a logic error."""
import ctypes
import struct

import numpy as np
import pytest

from libgooey_b200 import engine as G
import oracle_lib as O
import engine_scripts as S
from test_ffi_surface_gpu import busy_pattern

pytestmark = pytest.mark.gpu
TOL = 1e-5
EXACT_TOL = 1e-6
c = ctypes


def pcm(seed, frames, channels=2):
    rng = np.random.default_rng(seed)
    t = np.arange(frames)[:, None] / 44100.0
    tone = 0.4 * np.sin(2 * np.pi * (110.0 * (1 + np.arange(channels))[None, :]) * t)
    return (tone + rng.uniform(-0.3, 0.3, (frames, channels))).astype(np.float32)


def both(script, run):
    o = O.oracle_engine()
    g = G.Engine()
    script(o); script(g)
    want, got = run(o), run(g)
    assert not g.has_error(), g.get_error_message()
    o.close(); g.close()
    return got, want


def test_loop_on_the_loops_track_with_the_kit_playing():
    def script(e):
        busy_pattern(e)
        e.loop_load(0, pcm(1, 3000), 48000.0)
        e.loop_set_start(0, 0.2); e.loop_set_end(0, 0.9); e.loop_set_speed(0, 0.8); e.loop_set_gain(0, 0.7)
        e.loop_set_playing(0, True)
        e.loop_load(2, pcm(2, 1777, channels=1), 32000.0)
        e.loop_set_playing(2, True)
        e.mixer_set_track_pan(3, 0.3)
        e.sequencer_start()
    got, want = both(script, lambda e: e.render(20000))
    assert np.abs(want).max() > 0.05
    assert np.abs(got - want).max() <= TOL


def test_loops_alone_through_the_whole_mix():
    """No voice sounds: the output is loop mixer -> Loops track strip -> master gain (settled), all exact f32 products."""
    def script(e):
        e.loop_load(1, pcm(3, 2500), 44100.0)
        e.loop_set_start(1, 0.8); e.loop_set_end(1, 0.3); e.loop_set_speed(1, -1.37)      # wrap-around window, reverse
        e.loop_set_playing(1, True)
        e.loop_load(3, pcm(4, 900), 96000.0)
        e.loop_set_gain(3, 1.6); e.loop_set_playing(3, True)
    got, want = both(script, lambda e: np.concatenate([e.render(5000), e.render(3000)]))
    assert np.abs(want).max() > 0.05
    assert np.abs(got - want).max() <= EXACT_TOL


def test_mute_solo_gates_and_edits_between_render_calls():
    def run(e):
        a = e.render(4000)
        e.loop_set_mute(0, True); e.loop_set_gain(1, 0.25); e.loop_set_speed(1, 2.5)
        b = e.render(4000)
        e.loop_set_solo(0, True)                        # solo wins over mute (mod.rs:63-72)
        e.loop_set_position(1, 0.5); e.loop_restart(0)
        cc = e.render(4000)
        e.loop_set_playing(1, False)
        d = e.render(1000)
        return np.concatenate([a, b, cc, d]), [e.loop_get_position(k) for k in range(4)]

    def script(e):
        e.loop_load(0, pcm(5, 2000), 44100.0); e.loop_set_playing(0, True)
        e.loop_load(1, pcm(6, 4000), 22050.0); e.loop_set_start(1, 0.25); e.loop_set_end(1, 0.75); e.loop_set_playing(1, True)
    (got, gpos), (want, wpos) = both(script, run)
    assert np.abs(want).max() > 0.05
    assert np.abs(got - want).max() <= EXACT_TOL
    assert np.allclose(gpos, wpos, atol=1e-6)


def test_resample_warp_follows_the_engine_tempo():
    def script(e):
        e.loop_load(0, pcm(7, 5000), 44100.0)
        e.loop_set_source_bpm(0, 100.0); e.loop_set_pitch_mode(0, 1)
        e.set_bpm(133.0)
        e.loop_set_playing(0, True)
    got, want = both(script, lambda e: (e.render(6000), e.loop_get_source_bpm(0), e.loop_get_pitch_mode(0)))
    assert got[1:] == want[1:] == (100.0, 1)
    assert np.abs(got[0] - want[0]).max() <= EXACT_TOL


def test_preserve_pitch_time_stretch_linear_and_wrapped_windows():
    """PitchMode::PreservePitch (mixer/wsola.rs): grain search, overlap-add, stretcher state across calls, and the events that drop it."""
    def script(e):
        e.loop_load(0, pcm(30, 20000), 48000.0)
        e.loop_set_source_bpm(0, 100.0); e.loop_set_pitch_mode(0, 2); e.set_bpm(140.0)
        e.loop_set_playing(0, True)
        e.loop_load(1, pcm(31, 9000), 44100.0)
        e.loop_set_start(1, 0.7); e.loop_set_end(1, 0.4); e.loop_restart(1)                  # wrap-around window
        e.loop_set_source_bpm(1, 120.0); e.loop_set_pitch_mode(1, 2); e.loop_set_speed(1, 1.2); e.loop_set_playing(1, True)

    def run(e):
        a = e.render(5000)
        b = e.render(3000)                              # the stretchers carry over (mid-hop)
        e.loop_set_position(0, 0.3)                     # drops channel 0's stretcher: re-seeded at the new cursor
        e.loop_set_speed(1, -1.0)                       # reverse: channel 1 falls back to the direct read (loop_channel.rs:184)
        cc = e.render(3000)
        e.loop_set_pitch_mode(0, 1); e.loop_set_speed(1, 0.9)
        d = e.render(2000)
        return np.concatenate([a, b, cc, d]), [e.loop_get_position(k) for k in (0, 1)], e.loop_get_pitch_mode(1)
    (got, gpos, gm), (want, wpos, wm) = both(script, run)
    assert gm == wm == 2
    assert np.abs(want).max() > 0.05
    assert np.abs(got - want).max() <= EXACT_TOL
    assert np.allclose(gpos, wpos, atol=1e-6)


def test_bounce_with_loops_many_pieces_and_a_global_chain():
    """One bar = 88 200 frames: the lead pieces, double-buffered rows, the time-parallel strips feeding the chain kernel."""
    def script(e):
        busy_pattern(e)
        S.fx_chain(e, 11)
        e.loop_load(0, pcm(8, 30000), 44100.0); e.loop_set_playing(0, True)
        e.loop_load(1, pcm(9, 7000, channels=1), 48000.0); e.loop_set_speed(1, -0.5); e.loop_set_playing(1, True)
    got, want = both(script, lambda e: e.bounce_to_buffer(1))
    assert got.shape == want.shape == (88200,)
    assert np.abs(got - want).max() <= TOL


def test_loop_through_a_track_rack_effect():
    def script(e):
        e.loop_load(0, pcm(10, 6000), 44100.0); e.loop_set_playing(0, True)
        slot = e.track_effect_add(3, 1)                  # delay on the Loops track
        e.track_effect_set_param(3, slot, 1, 0.5); e.track_effect_set_param(3, slot, 2, 0.5)
    got, want = both(script, lambda e: e.render(30000))
    assert np.abs(got - want).max() <= TOL


def test_offline_channel_render_and_the_float_wav(tmp_path):
    def script(e):
        e.loop_load(2, pcm(12, 4096), 48000.0)
        e.loop_set_start(2, 0.1); e.loop_set_end(2, 0.6); e.loop_set_gain(2, 0.5)
        e.loop_set_mute(2, True)                        # ignored by the offline render
    got, want = both(script, lambda e: (e.loop_render(2, 10000, 512), e.loop_render(0, 16)))
    assert got[1] is None and want[1] is None
    assert np.abs(got[0] - want[0]).max() <= EXACT_TOL
    g = G.Engine(); script(g)
    L = G.lib()
    L.gooey_engine_loop_render_to_wav.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32, c.c_uint32, c.c_char_p]
    L.gooey_engine_loop_render_to_wav.restype = c.c_bool
    path = tmp_path / "loop.wav"
    assert L.gooey_engine_loop_render_to_wav(g._h, 2, 10000, 512, str(path).encode())
    assert not L.gooey_engine_loop_render_to_wav(g._h, 1, 100, 0, str(path).encode())
    g.close()
    raw = path.read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE"
    fmt = raw.index(b"fmt ")
    tag, ch, rate, _, _, bits = struct.unpack("<HHIIHH", raw[fmt + 8:fmt + 24])
    assert (tag, ch, rate, bits) == (3, 2, 44100, 32)            # IEEE float stereo (ffi.rs:8032-8037)
    data = raw.index(b"data")
    n = struct.unpack("<I", raw[data + 4:data + 8])[0]
    assert n == 10000 * 2 * 4
    assert np.abs(np.frombuffer(raw[data + 8:data + 8 + n], np.float32).reshape(-1, 2) - want[0]).max() <= EXACT_TOL


def test_sampler_rack_layers_retriggers_and_carries_voices_across_calls():
    def run(e):
        r = e.sampler_register()
        assert r == 0 and e.mixer_route_source(5, 0) and not e.mixer_route_source(6, 0)
        e.sampler_set_slot_buffer(0, 0, pcm(20, 9000), 44100.0)
        e.sampler_set_slot_buffer(0, 1, pcm(21, 2500, channels=1), 22050.0)
        e.sampler_set_slot_buffer(0, 7, pcm(22, 700), 48000.0)
        info = (e.sampler_slot_is_loaded(0, 1), e.sampler_slot_frames(0, 1), e.sampler_slot_channels(0, 1), e.sampler_slot_sample_rate(0, 1),
                e.sampler_slot_is_loaded(0, 2), e.sampler_get_source_id(0), e.sampler_get_source_id(3))
        e.sampler_trigger(0, 0, 1.0); e.sampler_trigger(0, 1, 0.6)
        a = e.render(3000)
        e.sampler_trigger(0, 7, 0.9); e.sampler_trigger(0, 0, 0.5)         # layered on the voices still sounding
        b = e.render(3000)
        e.sampler_clear_slot(0, 0)                                           # stops its voices (sampler.rs:187-194)
        cc = e.render(6000)
        d = e.render(500)                                                    # everything has run out: silence
        return np.concatenate([a, b, cc, d]), info
    (got, ginfo), (want, winfo) = both(lambda e: None, run)
    assert ginfo == winfo == (True, 2500, 1, 22050.0, False, 5, 0xFFFFFFFF)
    assert np.abs(want[:3000]).max() > 0.05 and np.abs(want[-500:]).max() == 0.0
    assert np.abs(got - want).max() <= EXACT_TOL


def test_sampler_voice_stealing_takes_the_oldest_voice():
    def run(e):
        e.sampler_register(); e.sampler_register()
        e.mixer_route_source(6, 2)
        e.sampler_set_slot_buffer(1, 4, pcm(23, 20000, channels=1), 44100.0)
        e.sampler_set_slot_buffer(1, 5, pcm(24, 300), 44100.0)
        out = []
        for k in range(36):                                                  # 32 voices, then four steals
            e.sampler_trigger(1, 4 if k % 3 else 5, 0.2 + 0.02 * k)
            out.append(e.render(64))
        out.append(e.render(2000))
        return np.concatenate(out)
    got, want = both(lambda e: None, run)
    assert np.abs(want).max() > 0.05
    assert np.abs(got - want).max() <= EXACT_TOL


def test_batch_of_engines_with_and_without_sources():
    """Row pairs are handed out per engine and source: engines without loops, with loops, with racks, in one launch."""
    n = 40

    def script(e, i):
        if i % 2 == 0:
            busy_pattern(e)
        if i % 3 != 1:
            e.loop_load(i % 4, pcm(100 + i, 1000 + 37 * i), [44100.0, 48000.0][i % 2])
            e.loop_set_speed(i % 4, 0.5 + 0.1 * (i % 7)); e.loop_set_gain(i % 4, 0.3 + 0.04 * (i % 10))
            e.loop_set_playing(i % 4, True)
        if i % 5 == 0:
            r = e.sampler_register()
            e.mixer_route_source(5 + r, 1)
            e.sampler_set_slot_buffer(r, 2, pcm(200 + i, 5000, channels=1 + i % 2), 44100.0)
            e.sampler_trigger(r, 2, 0.8)
    engines = [G.Engine() for _ in range(n)]
    for i, e in enumerate(engines):
        script(e, i)
    got = G.batch_bounce(engines, 1)
    assert not any(e.has_error() for e in engines)
    for e in engines:
        e.close()
    for i in range(n):
        o = O.oracle_engine(); script(o, i)
        want = o.bounce_to_buffer(1)
        o.close()
        assert np.abs(got[i] - want).max() <= TOL, i
        if i % 3 != 1:
            assert np.abs(want).max() > 0.01


def test_engines_sharing_one_resident_loop_buffer():
    src = pcm(40, 12000)
    engines = [G.Engine() for _ in range(6)]
    assert engines[0].loop_load(1, src, 44100.0)
    for i, e in enumerate(engines):
        if i:
            assert e.loop_share_buffer(i % 4, engines[0], 1)
            assert not e.loop_share_buffer(4, engines[0], 1) and not e.loop_share_buffer(0, engines[0], 0)      # bad channel / empty source
        e.loop_set_speed(1 if i == 0 else i % 4, 0.6 + 0.15 * i); e.loop_set_playing(1 if i == 0 else i % 4, True)
    got = G.batch_render(engines, 9000)
    for e in engines:
        e.close()
    for i in range(6):
        o = O.oracle_engine()
        ch = 1 if i == 0 else i % 4
        o.loop_load(ch, src, 44100.0); o.loop_set_speed(ch, 0.6 + 0.15 * i); o.loop_set_playing(ch, True)
        want = o.render(9000)
        o.close()
        assert np.abs(want).max() > 0.05
        assert np.abs(got[i] - want).max() <= EXACT_TOL, i


def test_requests_for_parts_that_are_not_built_latch_the_sticky_error():
    L = G.lib()
    L.gooey_engine_loop_effect_add.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32]; L.gooey_engine_loop_effect_add.restype = c.c_int32
    for call in (lambda e: L.gooey_engine_loop_effect_add(e._h, 0, 1),):
        g = G.Engine()
        assert not g.has_error()
        call(g)
        assert g.has_error() and "not built" in g.get_error_message()
        g.close()
    g = G.Engine()
    assert not g.loop_load(0, np.float32([0.0, np.inf]), 44100.0) and not g.loop_load(4, pcm(1, 8), 44100.0) and not g.loop_load(0, pcm(1, 8), -1.0)
    assert not g.sampler_set_slot_buffer(0, 0, pcm(1, 8), 44100.0)               # rack not registered
    assert g.sampler_register() == 0
    assert not g.sampler_set_slot_buffer(0, 16, pcm(1, 8), 44100.0) and not g.sampler_set_slot_buffer(0, 0, pcm(1, 8, channels=3), 44100.0)
    assert not g.sampler_trigger(0, 0, 1.0)
    assert [g.sampler_register() for _ in range(4)] == [1, 2, 3, -1]
    assert not g.has_error()
    g.close()


# ---- added after the round's last GPU run (the tests above passed on a B200; these ran only against the oracle and the host build) ----
def test_queued_swaps_land_on_the_grid_and_carry_their_own_rate_and_tempo():
    """gooey_engine_loop_queue_swap (loop_channel.rs:413-423, 249-276): the take replaces the buffer at the first grid boundary, the count
    and the new tempo tag are visible afterwards, a cancelled take never lands, a second queue replaces the first."""
    def run(e):
        e.set_bpm(126.0)
        e.loop_load(0, pcm(50, 6000), 44100.0); e.loop_set_source_bpm(0, 110.0); e.loop_set_pitch_mode(0, 1); e.loop_set_playing(0, True)
        e.loop_load(1, pcm(51, 5000), 48000.0); e.loop_set_start(1, 0.7); e.loop_set_end(1, 0.2); e.loop_restart(1); e.loop_set_playing(1, True)
        e.loop_load(2, pcm(52, 20000), 44100.0); e.loop_set_source_bpm(2, 100.0); e.loop_set_pitch_mode(2, 2); e.loop_set_playing(2, True)
        a = e.render(1000)
        assert e.loop_queue_swap(0, pcm(53, 4000) * 0.5, 32000.0, 140.0, 4)
        assert e.loop_queue_swap(1, pcm(54, 3000, channels=1), 44100.0, 0.0, 1)
        assert e.loop_queue_swap(2, pcm(55, 15000), 44100.0, 90.0, 3)
        assert e.loop_queue_swap(3, pcm(56, 100), 44100.0, 0.0, 1)          # nothing plays on channel 3: it can never land
        b = e.render(9000)
        info1 = [e.loop_swaps_completed(k) for k in range(4)], e.loop_get_source_bpm(0), e.loop_get_source_bpm(2)
        e.loop_queue_swap(0, pcm(57, 2000), 44100.0, 0.0, 2); e.loop_cancel_queued_swap(0)
        e.loop_queue_swap(1, pcm(58, 2000), 44100.0, 0.0, 2); e.loop_queue_swap(1, pcm(59, 2500) * 0.25, 44100.0, 0.0, 2)
        cc = e.render(5000)
        info2 = [e.loop_swaps_completed(k) for k in range(4)]
        return np.concatenate([a, b, cc]), info1, info2
    (got, g1, g2), (want, w1, w2) = both(lambda e: None, run)
    assert g1 == w1 == ([1, 1, 1, 0], 140.0, 90.0)
    assert g2 == w2 == [1, 2, 1, 0]
    assert np.abs(want).max() > 0.05
    assert np.abs(got - want).max() <= EXACT_TOL


def test_sampler_rack_pattern_armed_on_the_transport_bounced_and_stopped():
    """The rack's 16-step pattern (ffi.rs:6173-6290): armed on the next quarter of the running transport, fires inside a render call, keeps
    running across calls, restarts from step 0 in a bounce, is silenced by stop_pattern; with the kit playing next to it."""
    def run(e):
        busy_pattern(e)
        assert e.sampler_register() == 0 and e.mixer_route_source(5, 2)
        e.set_bpm(140.0); e.set_swing(0.58)
        for slot, (n, ch, sr) in enumerate([(9000, 2, 44100.0), (2500, 1, 22050.0), (700, 2, 48000.0), (40000, 1, 44100.0)]):
            e.sampler_set_slot_buffer(0, slot, pcm(60 + slot, n, channels=ch), sr)
        for step in range(16):
            e.sampler_set_step(0, step, step % 3 != 1, (step * 5) % 6, 0.3 + 0.04 * step)          # pads 4 and 5 are empty: their hits are dropped
        info = [e.sampler_get_step(0, 7), e.sampler_start_pattern(0, 3), e.sampler_is_pattern_running(0)]
        e.sequencer_start()
        a = e.render(5000)
        assert e.sampler_start_pattern(0, 1)                                                      # the next quarter note
        info.append(e.sampler_get_pending_start_beat(0))
        b = e.render(30000)
        info += [e.sampler_is_pattern_running(0), e.sampler_get_pending_start_beat(0)]
        cc = e.render(12345)
        bounced = e.bounce_to_buffer(1)
        e.sequencer_start()                                                                       # the bounce stopped the sequencers
        d = e.render(4000)
        assert e.sampler_stop_pattern(0)
        f = e.render(3000)
        info.append(e.transport_get_beat_position())                                              # every rendered frame since sequencer_start, add by add
        return np.concatenate([a, b, cc, d, f]), bounced, info
    (got, gb, ginfo), (want, wb, winfo) = both(lambda e: None, run)
    assert ginfo == winfo and winfo[1] is False and winfo[4] is True and winfo[5] == -1.0
    assert np.abs(want).max() > 0.05 and np.abs(wb).max() > 0.05
    assert np.abs(got - want).max() <= TOL
    assert np.abs(gb - wb).max() <= TOL


def test_graph_layout_unroute_clear_and_reset_to_the_default():
    """gooey_engine_mixer_{unroute_source,get_source_route,clear_layout,reset_default_layout} and the strip getters (ffi.rs:6291-6320,
    6427-6455, 6472-6566): the audio after each layout edit and the values the getters report."""
    def run(e):
        busy_pattern(e)
        e.loop_load(0, pcm(70, 4000), 44100.0); e.loop_set_playing(0, True)
        e.mixer_set_track_gain(1, 1.6); e.mixer_set_track_pan(0, 0.3); e.mixer_set_track_mute(2, True)
        slot = e.track_effect_add(1, 1)
        e.track_effect_set_param(1, slot, 2, 0.5)
        e.sequencer_start()
        info = [[e.mixer_get_source_route(s) for s in range(6)], e.mixer_get_track_gain(1), e.mixer_get_track_pan(0), e.mixer_get_track_mute(2), e.mixer_get_track_solo(2)]
        a = e.render(6000)
        info.append((e.mixer_unroute_source(1), e.mixer_unroute_source(1), e.mixer_unroute_source(6), e.mixer_get_source_route(1)))
        b = e.render(6000)                                   # the bass is out of the mix
        e.mixer_clear_layout()
        info.append(([e.mixer_get_source_route(s) for s in range(5)], e.mixer_route_source(0, 0)))
        cc = e.render(3000)                                  # no tracks: silence
        info.append((e.mixer_add_track("only"), e.mixer_route_source(0, 0), e.mixer_route_source(4, 0)))
        d = e.render(6000)                                   # kit and loops on one fresh strip
        e.mixer_reset_default_layout()
        info.append(([e.mixer_get_source_route(s) for s in range(5)], e.mixer_get_track_gain(1), e.mixer_get_track_mute(2)))
        f = e.render(6000)                                   # default strips again, the delay rack is gone
        return np.concatenate([a, b, cc, d, f]), info
    (got, ginfo), (want, winfo) = both(lambda e: None, run)
    assert ginfo == winfo
    assert winfo[0] == [0, 1, 2, 3, 3, -1] and winfo[5] == (True, False, False, -1) and winfo[6] == ([-1] * 5, False) and winfo[8][0] == [0, 1, 2, 3, 3]
    assert np.abs(want[:6000]).max() > 0.05 and np.abs(want[12000:15000]).max() == 0.0 and np.abs(want[15000:]).max() > 0.05
    assert np.abs(got - want).max() <= TOL
