"""Test-side access to the checkers: oracle/_ref (the unmodified reference, compiled by oracle/Makefile),
oracle/libzsc_oracle.so (our CPU restatement) and tests/libzsc_cpuharness.so (host build of product logic)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libzsc_ref.so")
ORACLE_PATH = os.path.join(ROOT, "oracle", "libzsc_oracle.so")
HARNESS_PATH = os.path.join(ROOT, "tests", "libzsc_cpuharness.so")

import sys
sys.path.insert(0, ROOT)
from zsc_b200.capi import Zsc, _declare_zsc, u8p, u32p  # noqa: E402


def have_ref():
    return os.path.exists(REF_PATH)


_ref = None


def ref():
    """Zsc wrapper over the reference's own code."""
    global _ref
    if _ref is None:
        L = C.CDLL(REF_PATH, mode=C.RTLD_LOCAL)
        _declare_zsc(L)
        L.refprobe_sizeof_deflate_state.restype = C.c_uint32
        L.refprobe_sizeof_inflate_state.restype = C.c_uint32
        _ref = Zsc(L)
    return _ref


def ref_uncompress_batch(comp, comp_off, comp_len, raw_len, threads=None):
    """zsc_uncompress of n independent streams packed in `comp`, by the reference itself, one pthread per core
    (oracle/ref_probe.c:refprobe_batch).  Returns (rets, produced, out) with stream i at out[sum(raw_len[:i]):]."""
    L = ref().L
    n = len(comp_off)
    u64p, i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    L.refprobe_batch.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_void_p, u64p, u32p, C.c_void_p, u64p, u32p, u32p, i32p,
                                 C.c_uint32, C.c_int32, C.c_int32]
    comp = np.ascontiguousarray(comp, dtype=np.uint8)
    roff = np.concatenate([[0], np.cumsum(np.asarray(raw_len, dtype=np.uint64))]).astype(np.uint64)
    out = np.empty(int(roff[-1]) + 16, dtype=np.uint8)
    so = (C.c_uint64 * n)(*[int(v) for v in comp_off]); sl = (C.c_uint32 * n)(*[int(v) for v in comp_len])
    do = (C.c_uint64 * n)(*[int(v) for v in roff[:-1]]); dc = (C.c_uint32 * n)(*[int(v) for v in raw_len])
    dl = (C.c_uint32 * n)(); rt = (C.c_int32 * n)()
    L.refprobe_batch(1, threads or (os.cpu_count() or 4), n, comp.ctypes.data, so, sl, out.ctypes.data, do, dc, dl, rt, 0, 0, 0)
    return list(rt), list(dl), out[:int(roff[-1])]


_ora = None


def oracle():
    """Zsc wrapper over the CPU restatement oracle/zsc_oracle.c (uncompress, checksums, size checks)."""
    global _ora
    if _ora is None:
        L = C.CDLL(ORACLE_PATH, mode=C.RTLD_LOCAL)
        _declare_zsc(L, strict=False)
        L.ora_inflate_raw.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, u32p, C.POINTER(C.c_char_p)]
        _ora = Zsc(L)
    return _ora


_har = None
_har_n = None


def harness_narrow():
    """the harness compiled with the table geometry of narrow batches (10 / 8 root bits)"""
    global _har_n
    if _har_n is None:
        L = C.CDLL(HARNESS_PATH.replace(".so", "_n.so"), mode=C.RTLD_LOCAL)
        L.h_inflate.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, C.c_int, u32p]
        L.h_inflate_batched.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, C.c_int, u32p, C.c_uint32]
        L.h_inflate_spec.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, C.c_int, u32p, C.c_uint32]
        _har_n = L
    return _har_n


_har_w = None


def harness_wide():
    """the harness compiled with a second geometry of the speculative decoder (regions of 320 bits, 80 symbols per lane)"""
    global _har_w
    if _har_w is None:
        L = C.CDLL(HARNESS_PATH.replace(".so", "_w.so"), mode=C.RTLD_LOCAL)
        L.h_inflate.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, C.c_int, u32p]
        L.h_inflate_spec.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, C.c_int, u32p, C.c_uint32]
        _har_w = L
    return _har_w


def harness():
    global _har
    if _har is None:
        L = C.CDLL(HARNESS_PATH, mode=C.RTLD_LOCAL)
        L.h_inflate.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, C.c_int, u32p]
        L.h_inflate_batched.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, C.c_int, u32p, C.c_uint32]
        L.h_deflate_model.argtypes = [u8p, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), u8p, C.c_uint32, u32p, C.c_uint32, u32p]
        L.h_deflate_model.restype = C.c_uint32
        _har = L
    return _har


def h_inflate(comp, cap, wrap=1):
    comp = np.ascontiguousarray(comp, dtype=np.uint8)
    padded = np.zeros(len(comp) + 16, dtype=np.uint8)
    padded[:len(comp)] = comp
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    res = (C.c_uint32 * 7)()
    harness().h_inflate(padded.ctypes.data_as(u8p), len(comp), out.ctypes.data_as(u8p), cap, wrap, res)
    # the group form of the same decoder (fast batch decode + generic steps) must agree in every field and byte, with
    # the compact table geometry (wide batches) and with the wide roots (narrow batches)
    for H, group in ((harness(), 32), (harness(), 8), (harness_narrow(), 32), (harness_narrow(), 16)):
        out2 = np.zeros(max(cap, 1), dtype=np.uint8)
        res2 = (C.c_uint32 * 7)()
        H.h_inflate_batched(padded.ctypes.data_as(u8p), len(comp), out2.ctypes.data_as(u8p), cap, wrap, res2, group)
        assert list(res) == list(res2), (group, list(res), list(res2))
        assert np.array_equal(out[:res[2]], out2[:res2[2]])
    # the speculative warp decoder (inflate_spec.h, lanes one after the other), in both of its geometries, writing and counting
    for H in (harness_narrow(), harness_wide()):
        for opts in (0, 1):
            out2 = np.zeros(max(cap, 1), dtype=np.uint8)
            res2 = (C.c_uint32 * 7)()
            H.h_inflate_spec(padded.ctypes.data_as(u8p), len(comp), out2.ctypes.data_as(u8p), cap, wrap, res2, opts)
            assert list(res) == list(res2), ("spec", opts, list(res), list(res2))
            assert opts or np.array_equal(out[:res[2]], out2[:res2[2]])
    res3 = (C.c_uint32 * 7)()
    out3 = np.zeros(max(cap, 1), dtype=np.uint8)
    harness_narrow().h_inflate(padded.ctypes.data_as(u8p), len(comp), out3.ctypes.data_as(u8p), cap, wrap, res3)
    assert list(res) == list(res3) and np.array_equal(out[:res[2]], out3[:res3[2]])
    ret = C.c_int32(res[0]).value
    return ret, out[:res[2]].copy(), dict(reason=res[1], produced=res[2], consumed=res[3], data_errors=res[4],
                                          stored_check=res[5], have_check=res[6])


# level -> [mode, chain, nice, lazy, min_len, max_dist, force_type, wrap, zhdr, good, max_lazy]; must mirror zs_lz_params (engine.cu)
FAST_MAX_DIST = 32768 - 3 * 2048 - 272
#            level:  0    1    2    3    4    5    6    7    8     9
CHAIN_TAB = [0, 0, 4, 8, 48, 128, 160, 256, 384, 512]
NICE_TAB = [0, 258, 258, 258, 16, 32, 128, 128, 258, 258]
GOOD_TAB = [0, 258, 258, 258, 4, 8, 8, 8, 32, 32]
MAX_LAZY_TAB = [0, 258, 258, 258, 4, 16, 16, 32, 128, 258]


def lz_params(level, strategy=0, wrap=1, wbits=15):
    if level == -1:
        level = 6
    chain, nice, good, max_lazy = CHAIN_TAB[level], NICE_TAB[level], GOOD_TAB[level], MAX_LAZY_TAB[level]
    lazy = 0 if level < 2 else 1
    mode, min_len, force = 0, 3, -1
    if strategy == 2:
        mode = 2
    elif strategy == 3:
        mode, lazy = 1, 0
    elif strategy == 1:
        min_len = 6
    elif strategy == 4:
        force = 1
    if level == 0:
        mode, force = 2, 0
    lf = 0 if (strategy >= 2 or level < 2) else (1 if level < 6 else (2 if level == 6 else 3))
    hdr = ((8 + ((wbits - 8) << 4)) << 8) | (lf << 6)
    hdr += 31 - (hdr % 31)
    zhdr = (hdr >> 8) | ((hdr & 0xFF) << 8)
    max_dist = 1 << wbits
    if not (mode == 0 and chain > 0):
        max_dist = min(max_dist, FAST_MAX_DIST)     # single-candidate kernel: 32 KiB ring (deflate_lz.cu ZL_FAST_MAX_DIST)
    return [mode, chain, nice, lazy, min_len, max_dist, force, wrap, zhdr, good, max_lazy]


def model_deflate(src, max_block_len, level, strategy=0, wrap=1, wbits=15):
    """Bit-exact CPU prediction of what the GPU deflate path must emit (stream at a 16-byte aligned offset)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    n = len(src)
    padded = np.zeros(n + 64, dtype=np.uint8)
    padded[:n] = src
    cap = n + n // 4 + 4096
    out = np.zeros(cap, dtype=np.uint8)
    syms = np.zeros(max(n, 1), dtype=np.uint32)
    nsym = C.c_uint32(0)
    p = (C.c_int32 * 11)(*lz_params(level, strategy, wrap, wbits))
    sz = harness().h_deflate_model(padded.ctypes.data_as(u8p), n, max_block_len, p, out.ctypes.data_as(u8p), cap,
                                   syms.ctypes.data_as(u32p), len(syms), C.byref(nsym))
    return out[:sz].copy(), syms[:nsym.value].copy()
