"""ctypes access to oracle/_build/liboracle.so — the CPU restatement of the reference.

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        L = ctypes.CDLL(_SO)
        c = ctypes
        L.orc_siphash.restype = c.c_uint64
        L.orc_siphash.argtypes = [c.c_uint64, c.c_uint64, c.c_uint64, c.c_int, c.c_int]
        L.orc_hash_noise.restype = c.c_float
        L.orc_hash_noise.argtypes = [c.c_uint64]
        L.orc_max_curve.restype = c.c_float
        L.orc_max_curve.argtypes = [c.c_float, c.c_float]
        L.orc_smoother_coeff.restype = c.c_float
        L.orc_smoother_coeff.argtypes = [c.c_float, c.c_float]
        L.orc_render_voices.restype = c.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def render_voices(patches, frames, triggers=(), params=(), sample_rate=44100.0, threads=1):
    """triggers: (voice, frame, velocity); params: (voice, frame, ffi_param, value, snap)."""
    from libgooey_b200._lib import VoicePatch
    n = len(patches)
    arr = patches if isinstance(patches, ctypes.Array) else (VoicePatch * n)(*patches)
    ev = [(v, f, 0, 0, vel) for (v, f, vel) in triggers] + [(v, f, 2 if s else 1, p, x) for (v, f, p, x, s) in params]
    ev.sort(key=lambda e: (e[0], e[1]))
    m = len(ev)
    ev_v = np.array([e[0] for e in ev], np.uint32)
    ev_f = np.array([e[1] for e in ev], np.uint32)
    ev_k = np.array([e[2] for e in ev], np.uint32)
    ev_p = np.array([e[3] for e in ev], np.uint32)
    ev_x = np.array([e[4] for e in ev], np.float32)
    out = np.zeros((n, frames), np.float32)
    u32, f32 = ctypes.c_uint32, ctypes.c_float
    rc = lib().orc_render_voices(arr, n, f32(sample_rate), frames, m, _p(ev_v, u32), _p(ev_f, u32), _p(ev_k, u32), _p(ev_p, u32),
                                 _p(ev_x, f32), _p(out, f32), int(threads))
    assert rc == 0
    return out


def oracle_engine(sample_rate=44100.0):
    """The oracle's FfiEngine behind the same Python class the product uses (FFI-named methods)."""
    from libgooey_b200.engine import Engine
    return Engine(sample_rate, library=lib(), prefix="orc_engine_")


def trigger_table(engine, channel, frames, cap=4096):
    """Sequencer trigger frames / velocities of one channel after reset + start (bit-exact gate)."""
    L = lib()
    fr = np.zeros(cap, np.uint32)
    ve = np.zeros(cap, np.float32)
    L.orc_engine_trigger_table.restype = ctypes.c_uint32
    L.orc_engine_trigger_table.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    n = L.orc_engine_trigger_table(engine._h, channel, frames, fr.ctypes.data, ve.ctypes.data, cap)
    return fr[:n].copy(), ve[:n].copy()


def bounce_many(script, indices, bars):
    """Oracle bounce of engine i (configured by script(engine, i)) for every i, one engine per host thread (ctypes calls
    release the GIL, engines share nothing).  Returns {i: mono array}."""
    from concurrent.futures import ThreadPoolExecutor

    def one(i):
        o = oracle_engine()
        script(o, i)
        out = o.bounce_to_buffer(bars)
        o.close()
        return i, out
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        return dict(ex.map(one, list(indices)))
