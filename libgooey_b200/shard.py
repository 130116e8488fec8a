"""Multi-GPU sharding of a batch: one process per GPU, contiguous shard per rank, NO data-path collective
(engines share nothing — `GooeyEngine` has no statics, src/ffi.rs:670-771).

Voices / engines are first interleaved by cost class (instrument type: a kick costs ~20x a hi-hat) so that every
contiguous shard carries the same mix, then rank r of N takes [r * ceil(n / N), (r + 1) * ceil(n / N)).
`torch.distributed` is used only by callers that want a barrier or a max-over-ranks timing; each rank drains its own
shard to its own pinned host buffer.
"""
import numpy as np


def balanced_order(kinds):
    """Permutation of range(len(kinds)) that round-robins the cost classes: position p holds an item of the class
    that is most behind its fair share, ties broken by class id — deterministic, identical on every rank."""
    kinds = np.asarray(kinds)
    classes = sorted(set(kinds.tolist()))
    queues = {k: list(np.nonzero(kinds == k)[0]) for k in classes}
    total = {k: len(queues[k]) for k in classes}
    taken = {k: 0 for k in classes}
    order = []
    n = len(kinds)
    for p in range(n):
        best, best_lag = None, None
        for k in classes:
            if taken[k] >= total[k]:
                continue
            lag = total[k] * (p + 1) / n - taken[k]
            if best is None or lag > best_lag + 1e-12:
                best, best_lag = k, lag
        order.append(int(queues[best][taken[best]]))
        taken[best] += 1
    return np.array(order, dtype=np.int64)


def shard_bounds(n, rank, world):
    """Contiguous block of rank `rank` (the last ranks may be short or empty when world does not divide n)."""
    per = -(-n // world)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def shard_indices(kinds, rank, world):
    """Indices (into the caller's batch) that rank `rank` of `world` renders."""
    order = balanced_order(kinds)
    lo, hi = shard_bounds(len(order), rank, world)
    return order[lo:hi]
