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


class OrcPatch(ctypes.Structure):
    """Same layout as GooeyVoicePatch (include/gooey_batch.h); lets CPU-only callers avoid the product package."""
    _fields_ = [("instrument", ctypes.c_uint32), ("aux", ctypes.c_uint32), ("params", ctypes.c_float * 24)]


def patch_array(patches):
    """ctypes array of patches from either product VoicePatch structs or raw (instrument, aux, params) tuples."""
    if isinstance(patches, ctypes.Array):
        return patches
    if patches and isinstance(patches[0], tuple):
        arr = (OrcPatch * len(patches))()
        for i, (inst, aux, params) in enumerate(patches):
            arr[i].instrument = inst
            arr[i].aux = aux
            for j, x in enumerate(params):
                arr[i].params[j] = x
        return arr
    return (type(patches[0]) * len(patches))(*patches)


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def render_voices(patches, frames, triggers=(), params=(), sample_rate=44100.0, threads=1):
    """triggers: (voice, frame, velocity); params: (voice, frame, ffi_param, value, snap)."""
    n = len(patches)
    arr = patch_array(patches)
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


def rust_bounce(instruments, sequencers, samples, master=None, limiter=True, bpm=120.0, sample_rate=44100.0):
    """Oracle RustEngine (src/engine/mod.rs + src/bounce.rs restated): instruments = [(name, obj with .patch)],
    sequencers = [(name, enabled[16], velocity[16])]; returns the mono bounce of `samples` frames."""
    L = lib()
    c = ctypes
    SR = sample_rate
    L.orc_rust_engine_new.restype = c.c_void_p
    L.orc_rust_engine_new.argtypes = [c.c_float]
    L.orc_rust_engine_free.argtypes = [c.c_void_p]
    L.orc_rust_engine_add_instrument.argtypes = [c.c_void_p, c.c_char_p, c.c_void_p]
    L.orc_rust_engine_add_sequencer.argtypes = [c.c_void_p, c.c_char_p, c.c_void_p, c.c_void_p, c.c_uint32]
    L.orc_rust_engine_set_bpm.argtypes = [c.c_void_p, c.c_float]
    L.orc_rust_engine_set_master_gain.argtypes = [c.c_void_p, c.c_float]
    L.orc_rust_engine_clear_global_effects.argtypes = [c.c_void_p]
    L.orc_rust_engine_bounce_samples.argtypes = [c.c_void_p, c.c_uint32, c.c_void_p]
    e = L.orc_rust_engine_new(SR)
    L.orc_rust_engine_set_bpm(e, bpm)
    for name, inst in instruments:
        assert L.orc_rust_engine_add_instrument(e, name.encode(), c.byref(inst.patch)) == 0
    for name, en, ve in sequencers:
        en = np.asarray(en, np.uint8); ve = np.asarray(ve, np.float32)
        L.orc_rust_engine_add_sequencer(e, name.encode(), en.ctypes.data, ve.ctypes.data, len(en))
    if master is not None:
        L.orc_rust_engine_set_master_gain(e, master)
    if not limiter:
        L.orc_rust_engine_clear_global_effects(e)
    out = np.zeros(samples, np.float32)
    L.orc_rust_engine_bounce_samples(e, samples, out.ctypes.data)
    L.orc_rust_engine_free(e)
    return out
