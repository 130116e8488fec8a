"""The reference's own integration tests of the sample-playback FFI, restated against the oracle: tests/loop_mixer.rs (the
tests that do not need per-channel effect chains), tests/loop_render_wav.rs (content; the WAV container is the product's writer,
checked on the GPU) and the pattern-free part of tests/sampler_rack.rs.  Same call sequences, same thresholds (file:line in each
test).  They pin the oracle's loop mixer — in particular its WSOLA time-stretch: tempo follows the BPM ratio, pitch does not."""
import numpy as np
import pytest

import oracle_lib as O

SR = 44100.0
OFF, RESAMPLE, PRESERVE = 0, 1, 2


def stereo_sine(seconds, hz):                                   # loop_mixer.rs:11-21
    frames = int(np.float32(SR) * np.float32(seconds))
    i = np.arange(frames, dtype=np.float32)
    s = (np.sin((i / np.float32(SR) * np.float32(hz) * np.float32(2 * np.pi)).astype(np.float32)) * np.float32(0.5)).astype(np.float32)
    return np.stack([s, s], 1)


def render_peak(e, frames):                                     # :24-31: peak of the second half of the INTERLEAVED buffer
    buf = e.render(frames).reshape(-1)
    assert np.isfinite(buf).all()
    return float(np.abs(buf[frames:]).max())


def zero_crossing_frequency(x, sr):                             # :312-318
    a, b = x[:-1], x[1:]
    crossings = int((((a <= 0.0) & (b > 0.0)) | ((a >= 0.0) & (b < 0.0))).sum())
    return (crossings / 2.0) / (len(x) / sr)


@pytest.fixture
def e():
    eng = O.oracle_engine(SR)
    yield eng
    eng.close()


def test_load_and_render_produces_audio(e):                     # :34
    assert e.loop_load(0, stereo_sine(0.5, 220.0), SR)
    e.loop_set_playing(0, True)
    assert render_peak(e, 8192) > 1e-3
    assert e.loop_get_position(0) > 0.0


def test_load_rejects_invalid_inputs(e):                        # :59 (the cases a ctypes caller can express)
    good = stereo_sine(0.1, 220.0)
    assert not e.loop_load(99, good, SR)
    assert not e.loop_load(0, good, 0.0) and not e.loop_load(0, good, float("nan"))
    bad = good.copy(); bad[3, 0] = np.inf
    assert not e.loop_load(0, bad, SR)
    assert e.loop_load(0, good, SR)


def test_mute_silences_a_channel(e):                            # :114
    e.loop_load(0, stereo_sine(0.5, 220.0), SR); e.loop_set_playing(0, True)
    audible = render_peak(e, 8192)
    assert audible > 1e-3
    e.loop_set_mute(0, True)
    assert render_peak(e, 8192) < audible * 0.02


def test_solo_isolates_channels(e):                             # :143
    for ch in range(2):
        e.loop_load(ch, stereo_sine(0.5, 220.0), SR); e.loop_set_playing(ch, True)
    both = render_peak(e, 8192)
    e.loop_set_solo(0, True)
    assert render_peak(e, 8192) < both


def test_master_gain_scales_loops(e):                           # :240
    e.loop_load(0, stereo_sine(0.5, 220.0), SR); e.loop_set_playing(0, True)
    e.set_master_gain(1.0)
    loud = render_peak(e, 8192)
    assert loud > 1e-2
    e.set_master_gain(0.0)
    render_peak(e, 8192)
    assert render_peak(e, 8192) < loud * 0.01


def position_after(mode, engine_bpm, frames):
    o = O.oracle_engine(SR)
    o.loop_load(0, stereo_sine(2.0, 220.0), SR)
    o.loop_set_source_bpm(0, 120.0); o.loop_set_pitch_mode(0, mode); o.set_bpm(engine_bpm); o.loop_set_playing(0, True)
    o.render(frames)
    pos = o.loop_get_position(0)
    o.close()
    return pos


def test_resample_mode_tempo_ratio_matches_bpm_ratio():         # :321
    baseline, warped = position_after(OFF, 120.0, 8192), position_after(RESAMPLE, 240.0, 8192)
    assert baseline > 0.0 and abs(warped / baseline - 2.0) < 0.05


def test_preserve_pitch_mode_tempo_ratio_matches_bpm_ratio():   # :356
    baseline, warped = position_after(PRESERVE, 120.0, 16384), position_after(PRESERVE, 240.0, 16384)
    assert baseline > 0.0 and abs(warped / baseline - 2.0) < 0.25


def test_preserve_pitch_holds_frequency_while_resample_shifts_it():   # :399
    def left(mode, frames):
        o = O.oracle_engine(SR)
        o.loop_load(0, stereo_sine(2.0, 440.0), SR)
        o.loop_set_source_bpm(0, 120.0); o.loop_set_pitch_mode(0, mode); o.set_bpm(180.0); o.loop_set_playing(0, True)
        out = o.render(frames)[:, 0]
        o.close()
        return out
    warmup, measure = 4096, 8192
    preserved_hz = zero_crossing_frequency(left(PRESERVE, warmup + measure)[warmup:], SR)
    resampled_hz = zero_crossing_frequency(left(RESAMPLE, warmup + measure)[warmup:], SR)
    assert abs(preserved_hz - 440.0) < 44.0, preserved_hz
    assert abs(resampled_hz - 660.0) < 66.0, resampled_hz


def test_preserve_pitch_bpm_change_mid_stream_holds_pitch(e):   # :442
    e.loop_load(0, stereo_sine(2.0, 440.0), SR)
    e.loop_set_source_bpm(0, 120.0); e.loop_set_pitch_mode(0, PRESERVE); e.set_bpm(120.0); e.loop_set_playing(0, True)
    for _ in range(8):
        e.render(512)
    for bpm in (120.0, 150.0, 200.0, 250.0, 200.0, 150.0, 90.0, 60.0, 120.0):
        e.set_bpm(bpm)
        x = np.concatenate([e.render(512)[:, 0] for _ in range(6)])
        hz = zero_crossing_frequency(x, SR)
        assert abs(hz - 440.0) < 44.0, (bpm, hz)


def test_preserve_pitch_varispeed_still_shifts_pitch_by_design(e):    # :496
    e.loop_load(0, stereo_sine(2.0, 440.0), SR)
    e.loop_set_pitch_mode(0, PRESERVE); e.loop_set_speed(0, 1.5); e.loop_set_playing(0, True)
    left = e.render(4096 + 8192)[:, 0]
    assert abs(zero_crossing_frequency(left[4096:], SR) - 660.0) < 66.0


def test_preserve_pitch_finite_across_loop_seam(e):             # :533
    e.loop_load(0, stereo_sine(0.05, 330.0), SR)
    e.loop_set_source_bpm(0, 120.0); e.loop_set_pitch_mode(0, PRESERVE); e.set_bpm(150.0); e.loop_set_playing(0, True)
    assert np.isfinite(e.render(int(SR))).all()


def test_queued_swap_lands_in_preserve_pitch_mode(e):           # :579
    first = stereo_sine(0.1, 220.0)
    assert e.loop_load(0, first, SR)
    e.loop_set_source_bpm(0, 120.0); e.set_bpm(140.0); e.loop_set_pitch_mode(0, PRESERVE); e.loop_set_playing(0, True)
    assert e.loop_get_pitch_mode(0) == PRESERVE
    render_peak(e, len(first))
    baseline = e.loop_swaps_completed(0)
    assert e.loop_queue_swap(0, stereo_sine(0.1, 330.0), SR, 120.0, 1)
    render_peak(e, len(first) * 4)
    assert e.loop_swaps_completed(0) > baseline


def test_queued_swap_preserves_source_bpm_tag(e):               # :634
    first = stereo_sine(0.1, 220.0)
    assert e.loop_load(0, first, SR)
    e.loop_set_source_bpm(0, 100.0); e.set_bpm(150.0); e.loop_set_pitch_mode(0, RESAMPLE); e.loop_set_playing(0, True)
    assert e.loop_get_source_bpm(0) == 100.0
    baseline = e.loop_swaps_completed(0)
    assert e.loop_queue_swap(0, stereo_sine(0.1, 330.0), SR, 128.0, 1)
    render_peak(e, len(first) * 4)
    assert e.loop_swaps_completed(0) > baseline
    assert e.loop_get_source_bpm(0) == 128.0


# ---- tests/loop_render_wav.rs (content of the rendered stem) ---------------------------------------------------------------
def test_render_exact_frame_count_gain_region_and_mute_solo(e):  # :78, :98, :119, :144
    e.loop_load(0, np.full((4096, 2), 0.5, np.float32), SR)
    out = e.loop_render(0, 1000, 512)
    assert out.shape == (1000, 2)
    e.loop_set_gain(0, 0.5)
    out = e.loop_render(0, 256, 128)
    assert np.abs(out - 0.25).max() < 1e-3                       # gain baked in from the first sample
    e.loop_set_gain(0, 1.0)
    ramp = np.repeat(np.arange(400, dtype=np.float32)[:, None], 2, 1)
    e.loop_load(1, ramp, SR)
    e.loop_set_start(1, 0.0); e.loop_set_end(1, 0.25)
    out = e.loop_render(1, 350)
    assert np.abs(out[:, 0] - (np.arange(350) % 100)).max() < 1e-3
    e.loop_set_mute(0, True); e.loop_set_solo(1, True)
    assert np.abs(e.loop_render(0, 256, 128) - 0.5).max() < 1e-3


def test_render_rejects_invalid_arguments(e):                    # :196
    assert e.loop_render(0, 100) is None                         # nothing loaded
    e.loop_load(0, np.full((64, 2), 0.5, np.float32), SR)
    assert e.loop_render(9, 100) is None
    assert e.loop_render(0, 100) is not None


# ---- tests/sampler_rack.rs (without the transport-armed step pattern) --------------------------------------------------------
def test_registration_has_a_fixed_limit(e):                      # :21
    for rack in range(4):
        assert e.sampler_register() == rack
        assert e.sampler_get_source_id(rack) == 5 + rack
    assert e.sampler_register() == -1


def test_registration_keeps_legacy_sources_and_rack_routes_after_a_default_graph_reset(e):   # :21 (last assert), :43
    rack = e.sampler_register()
    assert e.mixer_get_source_route(0) == 0                      # SOURCE_DRUMKIT still on track 0
    source = e.sampler_get_source_id(rack)
    assert e.mixer_get_source_route(source) == -1
    e.mixer_reset_default_layout()
    assert e.mixer_route_source(source, 3)
    assert e.mixer_get_source_route(source) == 3


def test_graph_layout_calls(e):                                  # ffi.rs:6291-6320, 6427-6455, 6472-6566
    assert [e.mixer_get_source_route(s) for s in range(6)] == [0, 1, 2, 3, 3, -1]
    e.mixer_set_track_gain(1, 1.7); e.mixer_set_track_pan(2, 0.2); e.mixer_set_track_mute(0, True); e.mixer_set_track_solo(3, True)
    assert e.mixer_get_track_gain(1) == pytest.approx(1.7) and e.mixer_get_track_pan(2) == pytest.approx(0.2)
    assert e.mixer_get_track_mute(0) and e.mixer_get_track_solo(3) and not e.mixer_get_track_mute(1)
    assert e.mixer_get_track_gain(9) == 1.0 and e.mixer_get_track_pan(9) == 0.5 and not e.mixer_get_track_solo(9)
    assert e.mixer_unroute_source(1) and not e.mixer_unroute_source(1) and not e.mixer_unroute_source(7)
    assert e.mixer_get_source_route(1) == -1
    e.mixer_clear_layout()
    assert [e.mixer_get_source_route(s) for s in range(5)] == [-1] * 5 and not e.mixer_route_source(0, 0)
    assert e.mixer_add_track("only") == 0 and e.mixer_route_source(0, 0)
    e.mixer_reset_default_layout()
    assert [e.mixer_get_source_route(s) for s in range(5)] == [0, 1, 2, 3, 3]
    assert e.mixer_get_track_gain(1) == 1.0 and not e.mixer_get_track_mute(0) and not e.mixer_get_track_solo(3)


def test_loaded_slot_can_be_routed_and_triggered(e):             # :56 (up to the step pattern)
    rack = e.sampler_register()
    assert e.mixer_route_source(e.sampler_get_source_id(rack), 3)
    assert e.sampler_set_slot_buffer(rack, 0, np.full(4096, 0.5, np.float32), SR)
    assert e.sampler_slot_is_loaded(rack, 0) and e.sampler_slot_frames(rack, 0) == 4096
    assert e.sampler_slot_channels(rack, 0) == 1 and e.sampler_slot_sample_rate(rack, 0) == SR
    assert e.sampler_trigger(rack, 0, 0.8)
    assert np.abs(e.render(256)).max() > 0.01
    assert e.sampler_clear_slot(rack, 0)
    assert not e.sampler_slot_is_loaded(rack, 0) and not e.sampler_trigger(rack, 0, 1.0)


def test_pattern_start_is_bar_quantized_and_never_seeks_the_clip_transport(e):   # sampler_rack.rs:227-268
    e.set_bpm(60.0)
    rack = e.sampler_register()
    assert e.mixer_route_source(e.sampler_get_source_id(rack), 3)
    assert e.sampler_set_slot_buffer(rack, 0, np.full(4096, 0.5, np.float32), SR)
    assert e.sampler_set_step(rack, 0, True, 0, 1.0)
    assert e.sampler_get_step(rack, 0) == (True, 0, 1.0)
    e.sequencer_start()
    e.render(int(SR / 10.0))                                     # beat 0.1
    assert e.sampler_start_pattern(rack, 2)                      # CLIP_QUANTIZE_BAR
    assert e.sampler_get_pending_start_beat(rack) == 4.0 and not e.sampler_is_pattern_running(rack)
    e.render(int((4.0 - 0.1) * SR))
    assert not e.sampler_is_pattern_running(rack)
    before = e.transport_get_beat_position()
    e.render(1)
    assert e.sampler_is_pattern_running(rack)
    assert np.abs(e.render(64)).max() > 0.001                    # step zero fires on the boundary
    assert e.transport_get_beat_position() > before
    assert e.sampler_stop_pattern(rack) and not e.sampler_is_pattern_running(rack)
    assert e.sampler_start_pattern(rack, 2) and e.sampler_get_pending_start_beat(rack) == 8.0
    assert e.sampler_cancel_pattern_start(rack) and e.sampler_get_pending_start_beat(rack) == -1.0


def test_the_products_host_schedule_fires_that_pattern_on_the_same_frame():
    """The same scenario through gooey_b200_sampler_schedule (the host code engines_render runs): armed at beat 0.1 for beat 4, the first hit
    lands on the first frame after (4 - 0.1) s of further rendering."""
    import ctypes as c
    from test_samples_cpu import product_schedule
    steps = [(i == 0, 0, 1.0) for i in range(16)]
    pre, wait = int(SR / 10.0), int((4.0 - 0.1) * SR)
    _, beat = product_schedule(SR, 60.0, 0.5, steps, True, 0.0, -1.0, False, [pre])
    hits, _ = product_schedule(SR, 60.0, 0.5, steps, True, beat, 4.0, False, [wait, 1, 64])
    assert hits == [(wait, 0, 1.0)]
