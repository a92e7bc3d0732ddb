"""Synthetic workloads of the BASELINE.json configs (generators in tools/datagen.c)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(os.path.dirname(_HERE), "tools", "libzscgen.so")
MIXED, TELEMETRY, TEXT, RANDOM = 0, 1, 2, 3
CANTERBURY_SIZES = [152089, 513216, 11150, 1029744, 38240, 426754, 481861, 24603, 3721, 4227, 125179]

_gen = None


def _lib():
    global _gen
    if _gen is None:
        if not os.path.exists(_LIB):
            raise RuntimeError(f"{_LIB} missing: run `make testlibs`")
        _gen = C.CDLL(_LIB)
        _gen.zscgen_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_int]
        _gen.zscgen_fill.restype = None
    return _gen


def fill(n, seed, kind, piece=1 << 20, threads=None, out=None):
    """n bytes of workload `kind`; piece k is generated from (seed, k) so any sub-range is reproducible."""
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    if threads is None:
        threads = min(32, os.cpu_count() or 1)
    if n:
        _lib().zscgen_fill(out.ctypes.data, n, seed, kind, piece, threads)
    return out


def mixed(n, seed=1):
    """Config 2: alternating text / telemetry(+5 % noise) segments of 4..64 KiB."""
    return fill(n, seed, MIXED)


def telemetry_buffers(count, size=262144, seed=1000):
    """Config 3: `count` independent telemetry-like buffers, buffer i seeded with seed + i."""
    return fill(count * size, seed, TELEMETRY, piece=size)


def canterbury_shaped(seed=0xC0FFEE):
    """Config 1 stand-in: 11 buffers with the Canterbury file sizes, alternating text / binary."""
    out = []
    for i, sz in enumerate(CANTERBURY_SIZES):
        out.append(fill(sz, seed + i, TEXT if i % 2 == 0 else MIXED, piece=1 << 30, threads=1))
    return out


def random_bytes(n, seed=5):
    return fill(n, seed, RANDOM)
