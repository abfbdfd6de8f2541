"""Deterministic synthetic embeddings, host side.

Counter-based: element (row, col) of matrix `seed` = splitmix64 hash -> sum of four 16-bit
uniforms, centred, times 2^-16 (every value exact in fp32).  Bit-identical to the device
generator behind `Index.add_synthetic` (csrc/common.cuh: synth_value), so a host-made query
batch and a device-made corpus belong to the same reproducible workload.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def synth_rows(seed: int, row0: int, n: int, dim: int, dup_every: int = 0, zero_every: int = 0) -> np.ndarray:
    """fp32 [n, dim]: rows row0..row0+n of synthetic matrix `seed`."""
    out = np.empty((n, dim), dtype=np.float32)
    key = _mix64(np.asarray([seed], dtype=np.uint64))[0]
    cols = np.arange(dim, dtype=np.uint64)
    step = max(1, (1 << 22) // max(dim, 1))
    for b in range(0, n, step):
        rows = np.arange(row0 + b, row0 + min(n, b + step), dtype=np.uint64)
        src = rows
        if dup_every and dup_every > 1:
            src = np.where(rows % np.uint64(dup_every) == np.uint64(dup_every - 1), rows - np.uint64(1), rows)
        with np.errstate(over="ignore"):
            h = _mix64((src[:, None] * np.uint64(dim) + cols[None, :]) ^ key)
        s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF))
             + ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + (h >> np.uint64(48)))
        blk = (s.astype(np.int64) - 131070).astype(np.float32) * np.float32(2.0 ** -16)
        if zero_every and zero_every > 1:
            blk[rows % np.uint64(zero_every) == np.uint64(zero_every - 1)] = 0.0
        out[b:b + len(rows)] = blk
    return out


TOPIC_SEED_OFFSET = 0x7091C5   # include/ragfin.h RAGFIN_TOPIC_SEED_OFFSET


def synth_topic_rows(seed: int, row0: int, n: int, dim: int, topic_rows: int, noise_shift: int = 3, rows_fn=None) -> np.ndarray:
    """fp32 [n, dim]: rows row0..row0+n of the "templated corpus" behind `Index.add_synthetic_topics`:
    row r = centre(r // topic_rows) + noise(r) * 2**-noise_shift, centre = row `topic` of synthetic matrix
    seed + TOPIC_SEED_OFFSET, noise = row r of synthetic matrix `seed`.  Exact in fp32, so bit-identical to the device.
    `rows_fn(seed, row0, n, dim)` may supply a faster generator with the same contract (the C oracle's in tests / bench)."""
    gen = rows_fn or synth_rows
    noise = gen(seed, row0, n, dim)
    t0, t1 = row0 // topic_rows, (row0 + max(n, 1) - 1) // topic_rows
    centres = gen(seed + TOPIC_SEED_OFFSET, t0, t1 - t0 + 1, dim)
    topic = (np.arange(row0, row0 + n, dtype=np.int64) // topic_rows) - t0
    return (centres[topic] + noise * np.float32(2.0 ** -noise_shift)).astype(np.float32)
