"""The settled-chain kernel (csrc/chain.cuh: tilt / delay / spring of whole warps of settled engines, pipelined ring reads)
against the general effect mixer (mix_kernel) it replaces: the two run the same arithmetic in the same order, so their
outputs must be IDENTICAL bit for bit, and against the oracle at the 1e-5 of BASELINE.json.  GOOEY_B200_NO_CHAIN_FAST=1
switches the kernel off; `gooey_b200_kernel_stat("chain_fast_kernel")` reports the engine-frames it took."""
import ctypes as c
import os

import numpy as np
import pytest

from libgooey_b200 import engine as G
from libgooey_b200._lib import lib
import oracle_lib as O
import engine_scripts as S

pytestmark = pytest.mark.gpu


def chain_units():
    L = lib()
    L.gooey_b200_kernel_stat.argtypes = [c.c_char_p, c.POINTER(c.c_uint64), c.POINTER(c.c_double), c.POINTER(c.c_double)]
    n, ms, vf = c.c_uint64(0), c.c_double(0), c.c_double(0)
    L.gooey_b200_kernel_stat(b"chain_fast_kernel", c.byref(n), c.byref(ms), c.byref(vf))
    return vf.value


def run(script, n, calls, fast, sr=44100.0):
    """calls: list of ("bounce", bars) / ("render", frames) / ("edit", fn(engines)); returns the list of outputs and the chain kernel's engine-frames."""
    if fast:
        os.environ.pop("GOOEY_B200_NO_CHAIN_FAST", None)
    else:
        os.environ["GOOEY_B200_NO_CHAIN_FAST"] = "1"
    try:
        lib().gooey_b200_kernel_stats_reset()
        engines = [G.Engine(sr) for _ in range(n)]
        for i, e in enumerate(engines):
            script(e, i)
        outs = []
        for kind, arg in calls:
            if kind == "bounce":
                outs.append(np.stack(G.batch_bounce(engines, arg)))
            elif kind == "render":
                outs.append(G.batch_render(engines, arg))
            else:
                arg(engines)
        for e in engines:
            e.close()
        return outs, chain_units()
    finally:
        os.environ.pop("GOOEY_B200_NO_CHAIN_FAST", None)


def c5_script(e, i):
    S.random_voice_params(e, 500 + i)
    S.pattern_engine(e, 600 + i, notes=(i % 3 == 0), graph=(i % 2 == 0))
    S.fx_chain(e, 700 + i)


def same(a, b):
    return all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))


def test_settled_chain_is_bit_identical_to_the_general_mixer_and_matches_the_oracle():
    n = 70                                   # two full warps of engines and a partial one
    calls = [("bounce", 1), ("bounce", 1)]   # the second bounce starts with every smoother settled
    fast, units = run(c5_script, n, calls, True)
    slow, none = run(c5_script, n, calls, False)
    assert none == 0.0
    assert units >= n * 88200 * 1.5          # all of the second bounce and most of the first (after the FFI edits have glided out)
    assert same(fast, slow)
    # oracle: the same two bounces, every engine
    want1 = O.bounce_many(c5_script, range(n), 1)
    worst = 0.0
    for i in range(n):
        w = want1[i]
        fin = np.isfinite(w)
        assert np.array_equal(fin, np.isfinite(fast[0][i]))
        worst = max(worst, float((np.abs(fast[0][i][fin] - w[fin]) / np.maximum(1.0, np.abs(w[fin]))).max()))
    print(f"chain kernel vs oracle, first bounce, {n} engines: worst err {worst:.3e}; engine-frames taken {units:.0f}")
    assert worst <= 1e-5


@pytest.mark.parametrize("sr", [22050.0, 48000.0])
def test_orders_subsets_limiter_stereo_and_other_sample_rates(sr):
    orders = [[7, 2, 0, 4, 1, 3, 8, 6, 9], [7, 2, 0, 6, 1, 3, 8, 4, 9], [1, 6, 4, 7, 2, 0, 3, 8, 9]]   # tilt/delay/spring permuted (ids 4, 1, 6)

    def script_for(group):
        def script(e, i):
            S.random_voice_params(e, 40 + i)
            S.pattern_engine(e, 80 + i, notes=False, graph=(i % 2 == 1))
            S.fx_chain(e, 90 + i, tilt=group != 1, delay=group != 2, spring=True, limiter=(group == 0))
            assert e.set_effect_order(orders[group])
            e.sequencer_start()
        return script
    for group in range(3):
        calls = [("render", 30000), ("render", 5000), ("render", 12345)]
        fast, units = run(script_for(group), 37, calls, True, sr)
        slow, none = run(script_for(group), 37, calls, False, sr)
        assert none == 0.0 and units >= 37 * (5000 + 12345)
        assert same(fast, slow), group
        assert np.abs(fast[2]).max() > 1e-3


def test_a_warp_with_one_unqualified_engine_stays_with_the_general_mixer():
    def script(e, i):
        c5_script(e, i)
        if i == 5:
            e.set_global_effect_param(S.FX_DELAY, 4, 1.0)       # ping-pong: not handled by the chain kernel
        if i == 40:
            e.set_global_effect_enabled(S.FX_PLATE, True)        # another effect kind in the chain
    calls = [("bounce", 1), ("bounce", 1)]
    fast, units = run(script, 96, calls, True)
    slow, none = run(script, 96, calls, False)
    assert none == 0.0
    assert 0 < units <= 32 * 2 * 88200      # only the third warp (engines 64..95) qualifies
    assert same(fast, slow)


def test_edits_between_renders_fall_back_and_come_back():
    def edit(engines):
        for i, e in enumerate(engines):
            e.set_global_effect_param(S.FX_TILT, 0, 0.3 + 0.4 * (i % 7) / 7.0)
            e.set_global_effect_param(S.FX_REVERB, 0, 0.6)
    def script(e, i):
        c5_script(e, i)
        e.sequencer_start()
    calls = [("render", 20000), ("edit", edit), ("render", 40000), ("render", 9999)]
    fast, units = run(script, 33, calls, True)
    slow, none = run(script, 33, calls, False)
    assert none == 0.0 and units >= 33 * 9999
    assert same(fast, slow)
    assert np.abs(fast[2]).max() > 1e-3
