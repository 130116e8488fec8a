"""The reference's own integration tests (tests/test_reference_suite_cpu.py: mute / solo, gains, volume-zero, panning, effect order,
mixer graph, instrument swap, sequencer-triggers switch, stereo effects, LFO routes — each citing the Rust test it restates), run
against the PRODUCT through the C ABI on the GPU, plus the product-vs-oracle parity of the sequencer-triggers switch
(`gooey_engine_set_sequencer_triggers_enabled`, ffi.rs:2188-2215).  Named to run last."""
import numpy as np
import pytest

from libgooey_b200 import engine as G
import oracle_lib as O
import test_reference_suite_cpu as R

pytestmark = pytest.mark.gpu


@pytest.fixture
def engine():
    made = []

    def make(sr=44100.0):
        e = G.Engine(sr)
        made.append(e)
        return e
    yield make
    for e in made:
        assert not e.has_error(), e.get_error_message()
        e.close()


for _name in dir(R):
    if _name.startswith("test_"):
        globals()[_name] = getattr(R, _name)      # the `engine` fixture of THIS module binds them to the product


def test_sequencer_triggers_switch_matches_the_oracle():
    def script(e):
        for s in (0, 4, 8, 12):
            e.sequencer_set_instrument_step(R.KICK, s, True)
        for s in (2, 10):
            e.sequencer_set_instrument_step_with_velocity(R.SNARE, s, True, 0.7)
        e.sequencer_start()
        out, midi = [], []
        out.append(e.render(30000)); midi.append(e.drain_midi_events())
        e.set_sequencer_triggers_enabled(False)
        out.append(e.render(30000)); midi.append(e.drain_midi_events())
        e.trigger_instrument_with_velocity(R.TOM, 0.9)
        out.append(e.render(12345)); midi.append(e.drain_midi_events())
        e.set_sequencer_triggers_enabled(True)
        out.append(e.render(30000)); midi.append(e.drain_midi_events())
        return np.concatenate(out), midi
    o, g = O.oracle_engine(), G.Engine()
    want, want_midi = script(o)
    got, got_midi = script(g)
    assert not g.has_error(), g.get_error_message()
    o.close(); g.close()
    assert want_midi[1] == [] and len(want_midi[0]) > 0 and len(want_midi[3]) > 0
    assert [[(c, o_) for c, _, o_ in m] for m in got_midi] == [[(c, o_) for c, _, o_ in m] for m in want_midi]      # channel and sample offset: exact
    assert np.abs(want).max() > 0.01
    assert float(np.abs(got - want).max()) <= 1e-5
