"""The speculative warp decoder (zsc_b200/csrc/cuda/inflate_spec.h) against the one-thread decoder, on the CPU.

The kernel (zs_inflate_spec_kernel) runs the lane functions of inflate_spec.h on 32 lanes at once; the harness runs the same
functions with the lanes one after the other (zp_round_host).  Results must be those of zi_inflate — which the known-answer
vectors pin against the reference (tests/test_oracle.py, tests/test_host.py) — in every field and byte, for intact streams
(reference src/inflate.c:704-1404, src/inffast.c:76-314) and for corrupted ones, where a round must stop in front of the
first symbol the reference rejects (src/inffast.c:181-189,285-301) and the recovery of zsc_uncompress
(src/zsc_uncompr.c:103-127) must see the same cursor."""
import ctypes as C
import zlib

import numpy as np
import pytest

import refimpl
from zsc_b200 import datagen
from zsc_b200.capi import u8p


def _both(H, comp, cap, wrap):
    comp = np.frombuffer(bytes(comp), dtype=np.uint8)
    padded = np.zeros(len(comp) + 16, np.uint8)
    padded[:len(comp)] = comp
    outs, ress = [], []
    for spec in (0, 1):
        out = np.zeros(max(cap, 1), np.uint8)
        res = (C.c_uint32 * 7)()
        if spec:
            H.h_inflate_spec(padded.ctypes.data_as(u8p), len(comp), out.ctypes.data_as(u8p), cap, wrap, res, 0)
        else:
            H.h_inflate(padded.ctypes.data_as(u8p), len(comp), out.ctypes.data_as(u8p), cap, wrap, res)
        outs.append(out); ress.append(list(res))
    assert ress[0] == ress[1], (ress[0], ress[1])
    assert np.array_equal(outs[0][:ress[0][2]], outs[1][:ress[1][2]])
    return ress[0]


def _stats(H, reset=True):
    s = (C.c_uint64 * 16)()
    H.h_spec_stats(s, 1 if reset else 0)
    return list(s)


def _inputs():
    S = 65536
    rng = np.random.default_rng(7)
    text = (b"the quick brown fox jumps over the lazy dog, and the dog does not care; " * 40)
    words = [text[i:i + int(rng.integers(3, 12))] for i in rng.integers(0, len(text) - 12, 6000)]
    return {
        "mixed": datagen.mixed(S, seed=5).tobytes(),
        "telemetry": datagen.telemetry_buffers(1, S, seed=77).tobytes(),
        "text": b" ".join(words)[:S],
        "runs": bytes(np.repeat(rng.integers(0, 256, 700, dtype=np.uint8), rng.integers(1, 400, 700))[:S]),
        "noise": rng.integers(0, 256, S // 4, dtype=np.uint8).tobytes(),
    }


@pytest.mark.parametrize("geometry", ["narrow", "wide"])
def test_spec_decoder_equals_serial_decoder_on_intact_streams(geometry):
    H = refimpl.harness_narrow() if geometry == "narrow" else refimpl.harness_wide()
    _stats(H)
    for name, data in _inputs().items():
        for level in (1, 6, 9):
            for wbits in (15, -15, 12):
                co = zlib.compressobj(level, zlib.DEFLATED, wbits)
                comp = co.compress(data) + co.flush()
                r = _both(H, comp, len(data), 1 if wbits > 0 else 0)
                assert r[0] == 0 and r[2] == len(data), (name, level, wbits, r)
        # fixed codes and a stream with full-flush markers (sections), and a short output buffer
        co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
        comp = co.compress(data) + co.flush()
        assert _both(H, comp, len(data), 1)[0] == 0
        co = zlib.compressobj(6)
        comp = b"".join(co.compress(data[i:i + 9000]) + co.flush(zlib.Z_FULL_FLUSH) for i in range(0, len(data), 9000)) + co.flush()
        assert _both(H, comp, len(data), 1)[0] == 0
        assert _both(H, comp, len(data) // 2, 1)[0] == 0xFFFFFFFB      # Z_BUF_ERROR: the rounds stop where the room ends
    st = _stats(H)
    assert st[0] > 100 and st[2] > 50 * st[0], st      # rounds ran, and took more than 50 symbols each on average


@pytest.mark.parametrize("geometry", ["narrow", "wide"])
def test_spec_decoder_equals_serial_decoder_on_corrupted_streams(geometry):
    """bit flips, byte smashes and truncation: same return code, reason, counts, recovered bytes"""
    H = refimpl.harness_narrow() if geometry == "narrow" else refimpl.harness_wide()
    rng = np.random.default_rng(11)
    _stats(H)
    kinds = set()
    for name, data in _inputs().items():
        co = zlib.compressobj(6)
        step = 16000
        base = b"".join(co.compress(data[i:i + step]) + co.flush(zlib.Z_FULL_FLUSH) for i in range(0, len(data), step)) + co.flush()
        for trial in range(40):
            comp = bytearray(base)
            how = trial % 4
            if how == 0:
                for _ in range(int(rng.integers(1, 4))):
                    comp[int(rng.integers(2, len(comp)))] ^= 1 << int(rng.integers(0, 8))
            elif how == 1:
                p = int(rng.integers(2, len(comp) - 8))
                comp[p:p + 8] = rng.integers(0, 256, 8, dtype=np.uint8).tobytes()
            elif how == 2:
                comp = comp[:int(rng.integers(len(comp) // 4, len(comp)))]
            else:
                comp[int(rng.integers(2, len(comp)))] ^= 0xFF
            r = _both(H, comp, len(data) + 1000, 1)
            kinds.add(r[0])
    assert 0xFFFFFFFD in kinds                                      # data errors were among them
    assert _stats(H)[0] > 100
