"""Deterministic synthetic workloads shared by tests and bench.py (seed 0x600E7)."""
import numpy as np

KICK, SNARE, HIHAT, TOM = 0, 1, 2, 3
# KickConfig::punch (src/instruments/kick.rs:257-350); the same table as libgooey_b200.voices.KICK_PRESETS["punch"] — kept here
# so that the CPU reference arm of bench.py can build the workload without importing the product package.
_KICK_PUNCH = [0.50, 0.20, 1.00, 0.20, 0.12, 0.60, 0.10, 0.85, 0.24, 1.00, 0.07, 0.11, 0.42, 0.20, 0.00, 0.47, 0.12, 0.02]


def drum_sweep_raw(n, seed=0x600E7, exact_tier=False):
    """C2 as plain data: [(instrument, aux, params), ...], velocities, kinds.  n patches cycling kick/snare/hihat/tom,
    every FFI-reachable parameter ~U[0,1), tuning fixed at neutral, one trigger at frame 0 with velocity U[0.3,1].
    exact_tier forces overdrive = 0 so the reconstructed half-band oversampler is never exercised."""
    class V:
        KICK, SNARE, HIHAT, TOM = KICK, SNARE, HIHAT, TOM
        KICK_PRESETS = {"punch": _KICK_PUNCH}

        @staticmethod
        def patch(instrument, params=(), aux=0):
            return (int(instrument), int(aux), [float(x) for x in params])
    return _drum_sweep(V, n, seed, exact_tier)


def drum_sweep_patches(n, seed=0x600E7, exact_tier=False):
    """drum_sweep_raw as GooeyVoicePatch structs of the product library."""
    from libgooey_b200 import voices as V
    return _drum_sweep(V, n, seed, exact_tier)


def _drum_sweep(V, n, seed, exact_tier):
    rng = np.random.default_rng(seed)
    patches, kinds = [], []
    vel = rng.uniform(0.3, 1.0, n).astype(np.float32)
    for i in range(n):
        k = i % 4
        u = rng.uniform(0.0, 1.0, 24).astype(np.float32)
        if k == V.KICK:
            p = list(V.KICK_PRESETS["punch"])
            # FFI-exposed: frequency, punch, sub, click, osc decay, pitch envelope amount, volume
            p[0], p[1], p[2], p[3], p[4], p[5], p[7] = (float(u[j]) for j in range(7))
            p[10], p[11], p[12] = float(u[8]), float(u[9]), float(u[10])     # noise layer (blend presets reach these)
            p[13] = 0.0 if exact_tier else float(u[11]) * 0.6                # overdrive
            patches.append(V.patch(V.KICK, p))
        elif k == V.SNARE:
            p = [float(x) for x in u[:19]]
            p[13] = float(int(u[13] * 4) % 4)                                # filter type 0..3
            p[16] = 0.0 if exact_tier else float(u[16])                      # overdrive
            patches.append(V.patch(V.SNARE, p))
        elif k == V.HIHAT:
            patches.append(V.patch(V.HIHAT, [float(x) for x in u[:5]]))
        else:
            patches.append(V.patch(V.TOM, [float(x) * 100.0 for x in u[:8]], aux=1))
        kinds.append(k)
    return patches, vel, kinds
