"""CPU checks of the host side: the C-ABI library loads and exports every symbol the headers declare (no compute
without a GPU), fails loudly without a device, and the host-resolved sequencer schedule is bit-exact."""
import ctypes
import os
import re

import numpy as np
import pytest

from libgooey_b200 import lib, GooeyError
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("gooey.h", "gooey_batch.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(gooey_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    L = lib()
    missing = [n for n in declared_symbols() if not hasattr(L, n)]
    assert not missing, missing
    assert len(declared_symbols()) > 50


def test_no_device_fails_loudly():
    L = lib()
    if L.gooey_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from libgooey_b200 import voices as V
    with pytest.raises(GooeyError):
        V.VoiceBatch([V.patch(V.KICK, V.KICK_PRESETS["tight"])])
    L.gooey_engine_new.restype = ctypes.c_void_p
    assert L.gooey_engine_new(ctypes.c_float(44100.0)) is None
    assert b"no CUDA device" in L.gooey_b200_last_error()


def schedule(bpm, swing, enabled, velocity, frames, sr=44100.0):
    L = lib()
    en = np.asarray(enabled, np.uint8)
    ve = np.asarray(velocity, np.float32)
    of = np.zeros(4096, np.uint32)
    ov = np.zeros(4096, np.float32)
    L.gooey_b200_sequencer_schedule.restype = ctypes.c_uint32
    L.gooey_b200_sequencer_schedule.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                                ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    n = L.gooey_b200_sequencer_schedule(sr, bpm, swing, en.ctypes.data, ve.ctypes.data, len(en), frames, of.ctypes.data, ov.ctypes.data, 4096)
    return of[:n].copy(), ov[:n].copy()


def test_default_swing_grid_is_k_times_5513():
    # sequencer.rs:331-346 pins samples_per_step = 5512.5 at 120 BPM / 44.1 kHz; round-half-away gives k * 5513
    fr, ve = schedule(120.0, 0.5, [1] * 16, [1.0] * 16, 8 * 88200)
    assert np.array_equal(fr, np.arange(len(fr), dtype=np.uint32) * 5513)
    assert len(fr) == 128


@pytest.mark.parametrize("bpm,swing", [(120.0, 0.5), (120.0, 0.62), (97.3, 0.41), (174.0, 0.7), (60.0, 0.0), (133.33, 1.0)])
def test_schedule_matches_oracle_sequencer_bit_exact(bpm, swing):
    rng = np.random.default_rng(int(bpm * 10 + swing * 100))
    en = (rng.random(16) < 0.5).astype(np.uint8)
    ve = rng.uniform(0.0, 1.0, 16).astype(np.float32)
    o = O.oracle_engine()
    o.set_bpm(bpm)
    o.set_swing(swing)
    for s in range(16):
        o.sequencer_set_instrument_step_settings(2, s, bool(en[s]), True, float(ve[s]), False, 0.0, 0.0, False, 0)
    want_f, want_v = O.trigger_table(o, 2, 8 * 88200)
    o.close()
    got_f, got_v = schedule(bpm, swing, en, ve, 8 * 88200)
    assert np.array_equal(got_f, want_f)
    assert np.array_equal(got_v.view(np.uint32), want_v.view(np.uint32))


def _chord(root, scale, degree, voicing, octave):
    import ctypes as c
    L = lib()
    L.gooey_b200_chord_notes.restype = c.c_uint32
    L.gooey_b200_chord_notes.argtypes = [c.c_uint32, c.c_uint32, c.c_uint32, c.c_uint32, c.c_int32, c.POINTER(c.c_uint8), c.c_uint32]
    out = (c.c_uint8 * 8)()
    n = L.gooey_b200_chord_notes(root, scale, degree, voicing, octave, out, 8)
    return list(out[:n])


def test_chord_tables_against_the_reference_unit_tests():
    """gooey_engine_poly_trigger_chord's note table (host-only view).  Known answers: music/voicing.rs:225-241 (Cmaj7 drop-2
    and shell voicings at octave 4), music/key.rs:270-276 (C major: I = Cmaj7, V = G7, vii = Bm7b5)."""
    assert _chord(0, 0, 0, 0, 4) == [60, 64, 67, 71]            # Cmaj7 root position
    assert _chord(0, 0, 0, 5, 4) == [55, 60, 64, 71]            # drop 2: G4 -> G3
    assert _chord(0, 0, 0, 8, 4) == [60, 64, 71]                # shell: root, 3rd, 7th
    assert _chord(0, 0, 4, 0, 4) == [67, 71, 74, 77]            # G7
    assert _chord(0, 0, 6, 0, 4) == [71, 74, 77, 81]            # Bm7b5
    assert _chord(0, 0, 13, 0, 4) == _chord(0, 0, 6, 0, 4)      # degree % 7
    assert _chord(9, 1, 0, 0, 3) == [57, 60, 64, 67]            # A natural minor: i7 = Am7 at octave 3
    assert _chord(0, 0, 0, 1, 4) == [64, 67, 71, 72]            # first inversion
    assert _chord(0, 0, 0, 3, 4) == [71, 72, 76, 79]            # third inversion
    assert _chord(0, 0, 0, 4, 4) == [60, 67, 76, 83]            # open: every other note up an octave
    assert _chord(0, 0, 0, 7, 4) == [60, 64, 79, 83]            # spread
    assert _chord(0, 0, 0, 9, 4) == [52, 67, 71]                # rootless: drop the root, 3rd down an octave
    assert _chord(0, 0, 0, 6, 4) == [60, 64, 67, 71]            # drop 3 needs five notes: unchanged
    assert _chord(0, 0, 0, 0, 40) == _chord(0, 0, 0, 0, 8)      # octave clamped to 0..8
    assert all(n <= 127 for n in _chord(11, 0, 0, 3, 8))
