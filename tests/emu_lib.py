"""ctypes access to tests/emu/libemu.so — host execution of the device functions (TEST INFRASTRUCTURE ONLY).

The kernels in libgooey_b200/csrc/kernels.cuh are thin index wrappers around __host__ __device__ functions; this
library compiles those same functions with g++ and drives them with loops that mirror the kernels, so that the CPU
suite can check planner/front/back logic against the oracle.  It is never loaded by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SRC = os.path.join(ROOT, "tests", "emu", "emu.cpp")
_SO = os.path.join(ROOT, "tests", "emu", "libemu.so")
_lib = None


def build():
    csrc = os.path.join(ROOT, "libgooey_b200", "csrc")
    deps = [_SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if os.path.exists(_SO) and all(os.path.getmtime(d) <= os.path.getmtime(_SO) for d in deps):
        return
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-shared", "-o", _SO, _SRC], check=True)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def render_voices(patches, frames, triggers=(), params=(), sample_rate=44100.0, mode=1, n_calls=1):
    """Same arguments as oracle_lib.render_voices.  mode 0 = general per-sample path, 1 = plan/front/back split.
    Returns (audio[n, frames], fast_calls[n])."""
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
    fast = np.zeros(n, np.int32)
    u32, f32 = ctypes.c_uint32, ctypes.c_float
    rc = lib().emu_render_voices(arr, n, f32(sample_rate), frames, m, _p(ev_v, u32), _p(ev_f, u32), _p(ev_k, u32), _p(ev_p, u32),
                                 _p(ev_x, f32), _p(out, f32), int(mode), int(n_calls), _p(fast, ctypes.c_int))
    assert rc == 0
    return out, fast


def math(kind, a, b=None, sample_rate=44100.0):
    """emu_math: 0 = division through the hoisted reciprocal, 1 = the front end's sine, 2 / 3 = additive triangle with that sine / the exact one."""
    a = np.ascontiguousarray(a, np.float32)
    b = a if b is None else np.ascontiguousarray(b, np.float32)
    out = np.empty_like(a)
    f32 = ctypes.c_float
    rc = lib().emu_math(int(kind), _p(a, f32), _p(b, f32), f32(sample_rate), _p(out, f32), a.size)
    assert rc == 0
    return out


def _events(ev):
    ev = sorted(ev, key=lambda e: e[0])
    fr = np.array([e[0] for e in ev], np.uint32)
    kd = np.array([e[1] for e in ev], np.uint32)
    a = np.array([e[2] for e in ev], np.float32)
    b = np.array([e[3] for e in ev], np.float32)
    return fr, kd, a, b


def poly_render(fn, preset, events, frames, sample_rate=44100.0):
    """events: (frame, kind, a, b); kind 0 note-on (a = midi note, b = velocity), 1 release_all, 2 set_param (a = id, b = value).
    fn = emu lib's emu_poly_render or the oracle's orc_poly_render (same signature)."""
    fr, kd, a, b = _events(events)
    out = np.zeros(frames, np.float32)
    u32, f32 = ctypes.c_uint32, ctypes.c_float
    fn.restype = None if fn.__name__.startswith("orc_") else ctypes.c_int
    fn(ctypes.c_uint32(preset), f32(sample_rate), ctypes.c_uint32(len(fr)), _p(fr, u32), _p(kd, u32), _p(a, f32), _p(b, f32), ctypes.c_uint32(frames), _p(out, f32))
    return out


def gran_render(fn, buf, events, frames, sample_rate=44100.0, buf_sr=44100.0):
    """events: (frame, kind, a, b); kind 0 trigger (b = velocity), 2 set_param, 3 snap_params, 4 set_seed (a)."""
    fr, kd, a, b = _events(events)
    buf = np.ascontiguousarray(buf, np.float32)
    out = np.zeros(frames, np.float32)
    u32, f32 = ctypes.c_uint32, ctypes.c_float
    fn.restype = None if fn.__name__.startswith("orc_") else ctypes.c_int
    fn(f32(sample_rate), _p(buf, f32), ctypes.c_uint32(buf.size), f32(buf_sr), ctypes.c_uint32(len(fr)), _p(fr, u32), _p(kd, u32), _p(a, f32), _p(b, f32),
       ctypes.c_uint32(frames), _p(out, f32))
    return out
