"""The z_stream API (include/zsc/zlib.h; reference include/zsc/zlib.h:150-990, scenarios test/zlib_gtest.cpp:59-80):
deflate / inflate in pieces with every flush mode, preset dictionaries, Z_NEED_DICT, inflateSync, error codes.
The GPU library is driven through ctypes exactly like the reference library (oracle/_ref), and the two are checked
against each other and against Python's zlib in both directions."""
import ctypes as C
import zlib

import numpy as np
import pytest

import refimpl
from zsc_b200 import capi, datagen

Z_NO_FLUSH, Z_PARTIAL_FLUSH, Z_SYNC_FLUSH, Z_FULL_FLUSH, Z_FINISH, Z_BLOCK = 0, 1, 2, 3, 4, 5
VERSION = b"1.2.11"


class Driver:
    """z_stream calls on one library (ours or the reference's): work buffer, piecewise feeding, small output buffers"""

    def __init__(self, L):
        self.L = capi.declare_zstream(L)

    def _strm(self, work_size):
        self.work = (C.c_uint8 * work_size)()
        s = capi.ZStream()
        s.next_work = C.cast(self.work, capi.u8p)
        s.avail_work = work_size
        return s

    def deflate_pieces(self, data, pieces, flushes, out_chunk=4096, level=6, wbits=15, strategy=0, zdict=None, mem_level=8):
        """feeds data[pieces[i-1]:pieces[i]] with flushes[i], then Z_FINISH; -> (compressed bytes, return codes seen)"""
        ws = C.c_uint32(0)
        assert self.L.deflateWorkSize2(wbits, mem_level, C.byref(ws)) == 0
        s = self._strm(ws.value)
        assert self.L.deflateInit2_(C.byref(s), level, 8, wbits, mem_level, strategy, VERSION, C.sizeof(capi.ZStream)) == 0
        if zdict is not None:
            d = (C.c_uint8 * len(zdict)).from_buffer_copy(zdict)
            assert self.L.deflateSetDictionary(C.byref(s), C.cast(d, capi.u8p), len(zdict)) == 0
        src = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(bytes(data) or b"\0")
        out = bytearray()
        obuf = (C.c_uint8 * out_chunk)()
        pos = 0
        steps = list(zip(pieces, flushes)) + [(len(data), Z_FINISH)]
        for end, flush in steps:
            s.next_in = C.cast(C.byref(src, pos), capi.u8p)
            s.avail_in = end - pos
            pos = end
            guard = 0
            while True:
                guard += 1
                assert guard < 100000, "deflate makes no progress"
                s.next_out = C.cast(obuf, capi.u8p)
                s.avail_out = out_chunk
                r = self.L.deflate(C.byref(s), flush)
                out += bytes(obuf[:out_chunk - s.avail_out])
                assert r in (0, 1, -5), r
                if r == 1 or (flush != Z_FINISH and s.avail_in == 0 and s.avail_out != 0) or r == -5:
                    break
            assert s.avail_in == 0
        assert r == 1 and s.total_in == len(data) and s.total_out == len(out)
        adler = s.adler
        assert self.L.deflateEnd(C.byref(s)) == 0
        return bytes(out), adler

    def inflate_pieces(self, comp, in_chunk, out_chunk, wbits=15, zdict=None, expect_dict=False, sync=False, cap=1 << 26):
        """-> (output bytes, final return code, total_in, list of return codes)"""
        ws = C.c_uint32(0)
        assert self.L.inflateWorkSize2(wbits, C.byref(ws)) == 0
        s = self._strm(ws.value)
        assert self.L.inflateInit2_(C.byref(s), wbits, VERSION, C.sizeof(capi.ZStream)) == 0
        if zdict is not None and wbits < 0:
            d = (C.c_uint8 * len(zdict)).from_buffer_copy(zdict)
            assert self.L.inflateSetDictionary(C.byref(s), C.cast(d, capi.u8p), len(zdict)) == 0
        src = (C.c_uint8 * max(len(comp), 1)).from_buffer_copy(bytes(comp) or b"\0")
        obuf = (C.c_uint8 * out_chunk)()
        out, codes, pos, r = bytearray(), [], 0, 0
        saw_dict = False
        guard = 0
        while r not in (1,) and len(out) < cap:
            guard += 1
            assert guard < 200000 + 4 * (len(comp) // max(in_chunk, 1)), "inflate makes no progress"
            if s.avail_in == 0 and pos < len(comp):
                n = min(in_chunk, len(comp) - pos)
                s.next_in = C.cast(C.byref(src, pos), capi.u8p)
                s.avail_in = n
                pos += n
            s.next_out = C.cast(obuf, capi.u8p)
            s.avail_out = out_chunk
            r = self.L.inflate(C.byref(s), Z_NO_FLUSH)
            codes.append(r)
            out += bytes(obuf[:out_chunk - s.avail_out])
            if r == 2:
                saw_dict = True
                assert zdict is not None and s.adler == zlib.adler32(zdict)
                d = (C.c_uint8 * len(zdict)).from_buffer_copy(zdict)
                bad = (C.c_uint8 * len(zdict)).from_buffer_copy(bytes(reversed(zdict)))
                if bytes(reversed(zdict)) != zdict:
                    assert self.L.inflateSetDictionary(C.byref(s), C.cast(bad, capi.u8p), len(zdict)) == -3
                assert self.L.inflateSetDictionary(C.byref(s), C.cast(d, capi.u8p), len(zdict)) == 0
                continue
            if r == -3 and sync:
                rs = self.L.inflateSync(C.byref(s))
                codes.append(("sync", rs))
                while rs == -3 and pos < len(comp):
                    n = min(in_chunk, len(comp) - pos)
                    s.next_in = C.cast(C.byref(src, pos), capi.u8p); s.avail_in = n; pos += n
                    rs = self.L.inflateSync(C.byref(s))
                    codes.append(("sync", rs))
                if rs != 0:
                    break
                r = 0
                continue
            if r < 0 and not (r == -5 and (pos < len(comp) or s.avail_out == 0)):
                break
            if r == -5 and pos >= len(comp) and s.avail_in == 0 and s.avail_out != 0:
                break
        if expect_dict:
            assert saw_dict
        tin = s.total_in - 0
        left = s.avail_in
        assert self.L.inflateEnd(C.byref(s)) == 0
        return bytes(out), r, tin, codes, left


def ours():
    return Driver(capi.lib())


def theirs():
    return Driver(refimpl.ref().L)


# ------------------------------------------------------------------ symbols and sizes (no GPU needed)
def test_zstream_symbols_exported_and_work_sizes_are_the_reference_s():
    L = capi.lib()
    for name in capi.ZSTREAM_SYMBOLS:
        assert hasattr(L, name), name
    capi.declare_zstream(L)
    v = C.c_uint32(0)
    assert L.deflateWorkSize2(15, 8, C.byref(v)) == 0 and v.value == 333600
    assert L.inflateWorkSize2(15, C.byref(v)) == 0 and v.value == 39920
    assert L.deflateWorkSize2(9, 1, C.byref(v)) == 0 and v.value == 5920 + 2 * 512 + 4 * 512 + 2 * 256 + 4 * 128
    assert L.deflateWorkSize2(16, 8, C.byref(v)) == -2 and L.inflateWorkSize2(7, C.byref(v)) == -2
    # argument errors that are decided before any device is needed
    s = capi.ZStream()
    assert L.deflateInit2_(C.byref(s), 6, 8, 15, 8, 0, b"9.9", C.sizeof(capi.ZStream)) == -6          # Z_VERSION_ERROR
    assert L.deflateInit2_(C.byref(s), 6, 8, 15, 8, 0, VERSION, C.sizeof(capi.ZStream)) == -2           # no work buffer
    assert L.inflateInit2_(C.byref(s), 15, VERSION, 12) == -6
    assert L.deflate(C.byref(s), 0) == -2 and L.inflate(C.byref(s), 0) == -2 and L.deflateEnd(C.byref(s)) == -2


# ------------------------------------------------------------------ deflate
@pytest.mark.gpu
@pytest.mark.parametrize("level,wbits,strategy", [(6, 15, 0), (1, 15, 0), (9, -15, 0), (0, 15, 0), (6, 9, 0), (6, 15, 3), (6, 31, 0)])
def test_deflate_in_pieces_with_every_flush_mode(level, wbits, strategy):
    rng = np.random.default_rng(7 + level)
    x = datagen.fill(700001, 60 + level, datagen.MIXED).tobytes()
    cuts = sorted(set(int(v) for v in rng.integers(1, len(x), 9)))
    flushes = [(Z_NO_FLUSH, Z_SYNC_FLUSH, Z_FULL_FLUSH, Z_PARTIAL_FLUSH, Z_BLOCK, Z_NO_FLUSH)[i % 6] for i in range(len(cuts))]
    comp, adler = ours().deflate_pieces(x, cuts, flushes, out_chunk=3001, level=level, wbits=wbits, strategy=strategy)
    if wbits > 15:
        import gzip
        assert gzip.decompress(comp) == x and adler == zlib.crc32(x)
    else:
        assert zlib.decompress(comp, wbits) == x
        if wbits > 0:
            assert adler == zlib.adler32(x)
            # and through the reference's one-shot zsc_uncompress
            if refimpl.have_ref():
                rr, out, used = refimpl.ref().uncompress(np.frombuffer(comp, np.uint8), len(x), window_bits=wbits)
                assert rr == 0 and used == len(comp) and out.tobytes() == x
    if level and refimpl.have_ref() and wbits == 15 and strategy == 0:
        # the same call sequence on the reference: sizes comparable (ours pays 5 bytes per internal chunk)
        rcomp, _ = theirs().deflate_pieces(x, cuts, flushes, out_chunk=3001, level=level, wbits=wbits, strategy=strategy)
        assert len(comp) <= 1.03 * len(rcomp) + 64, (len(comp), len(rcomp))


@pytest.mark.gpu
def test_deflate_finish_in_one_call_and_empty_input():
    L = capi.lib()
    for x in (datagen.fill(300000, 3, datagen.TEXT).tobytes(), b"", b"a"):
        comp, adler = ours().deflate_pieces(x, [], [], out_chunk=len(x) + len(x) // 8 + 4096)
        assert zlib.decompress(comp) == x and adler == zlib.adler32(x)


@pytest.mark.gpu
def test_deflate_with_preset_dictionary_both_ways():
    zdict = datagen.fill(20000, 5, datagen.TEXT).tobytes()
    x = (zdict[3000:9000] + datagen.fill(50000, 6, datagen.TEXT).tobytes() + zdict[100:7000])
    comp, _ = ours().deflate_pieces(x, [10000], [Z_NO_FLUSH], zdict=zdict)
    assert (comp[1] & 0x20) and int.from_bytes(comp[2:6], "big") == zlib.adler32(zdict)          # FDICT + DICTID
    d = zlib.decompressobj(zdict=zdict)
    assert d.decompress(comp) == x
    plain, _ = ours().deflate_pieces(x, [10000], [Z_NO_FLUSH])
    assert len(comp) < len(plain)                                                                 # the dictionary is really used
    # our inflate: Z_NEED_DICT, wrong dictionary refused, right one accepted
    out, r, tin, codes, left = ours().inflate_pieces(comp, 5000, 7000, zdict=zdict, expect_dict=True)
    assert r == 1 and out == x and tin == len(comp)
    # a stream Python's zlib made with the dictionary
    c = zlib.compressobj(6, zlib.DEFLATED, 15, zdict=zdict)
    pcomp = c.compress(x) + c.flush()
    out, r, tin, codes, left = ours().inflate_pieces(pcomp, 999, 4096, zdict=zdict, expect_dict=True)
    assert r == 1 and out == x
    if refimpl.have_ref():
        out, r, tin, codes, left = theirs().inflate_pieces(comp, 5000, 7000, zdict=zdict, expect_dict=True)
        assert r == 1 and out == x                                                                # the reference's inflate takes our stream
    # raw deflate: dictionary set up front on both sides
    rcomp, _ = ours().deflate_pieces(x, [], [], wbits=-15, zdict=zdict)
    assert zlib.decompressobj(-15, zdict=zdict).decompress(rcomp) == x
    out, r, tin, codes, left = ours().inflate_pieces(rcomp, 4000, 4000, wbits=-15, zdict=zdict)
    assert r == 1 and out == x


# ------------------------------------------------------------------ inflate
@pytest.mark.gpu
@pytest.mark.parametrize("in_chunk,out_chunk", [(1 << 20, 1 << 20), (4096, 70000), (70000, 4096), (13, 100000), (50000, 7)])
def test_inflate_in_pieces(in_chunk, out_chunk):
    n = 400000 if min(in_chunk, out_chunk) < 100 else 1500000
    x = datagen.fill(n, 70, datagen.MIXED)
    if refimpl.have_ref():
        rc, comp = refimpl.ref().compress(x, 100000, 6)
        comp = comp.tobytes()
    else:
        comp = zlib.compress(x.tobytes(), 6)
    if out_chunk < 100:
        x = x[:30000]
        comp = zlib.compress(x.tobytes(), 6)
    garbage = b"\x55" * 37
    out, r, tin, codes, left = ours().inflate_pieces(comp + garbage, in_chunk, out_chunk)
    assert r == 1 and out == x.tobytes()
    assert tin == len(comp)                      # nothing behind the stream is consumed


@pytest.mark.gpu
def test_inflate_raw_small_window_and_stored():
    x = datagen.fill(200000, 71, datagen.RANDOM).tobytes() + bytes(50000)
    for wb in (-15, 9, 15):
        c = zlib.compressobj(6, zlib.DEFLATED, wb)
        comp = c.compress(x) + c.flush()
        out, r, tin, codes, left = ours().inflate_pieces(comp, 30000, 30000, wbits=wb)
        assert r == 1 and out == x and tin == len(comp)
    # a window-15 stream into a window-9 inflater is refused like the reference refuses it
    comp = zlib.compress(x)
    out, r, tin, codes, left = ours().inflate_pieces(comp, 30000, 30000, wbits=9)
    assert r == -3


@pytest.mark.gpu
def test_inflate_data_error_then_sync_matches_the_reference():
    if not refimpl.have_ref():
        pytest.skip("oracle/_ref not present")
    x = datagen.fill(600000, 72, datagen.MIXED)
    rc, comp = refimpl.ref().compress(x, 100000, 6)
    rng = np.random.default_rng(3)
    for _ in range(6):
        bad = bytearray(comp.tobytes())
        pos = int(rng.integers(10, len(bad) - 10))
        bad[pos] ^= 0x5A
        a = ours().inflate_pieces(bytes(bad), 1 << 20, 1 << 20, sync=True)
        b = theirs().inflate_pieces(bytes(bad), 1 << 20, 1 << 20, sync=True)
        assert a[0] == b[0], (pos, len(a[0]), len(b[0]))          # the same bytes come out
        assert (a[1] == 1) == (b[1] == 1)
    # truncated input: Z_BUF_ERROR when nothing more can be done, output so far intact
    out, r, tin, codes, left = ours().inflate_pieces(comp.tobytes()[:len(comp) // 2], 10000, 10000)
    out2, r2, tin2, codes2, left2 = theirs().inflate_pieces(comp.tobytes()[:len(comp) // 2], 10000, 10000)
    assert r == r2 == -5 and out == out2


@pytest.mark.gpu
def test_zstream_error_codes():
    D = ours()
    L = D.L
    ws = C.c_uint32(0)
    L.deflateWorkSize2(15, 8, C.byref(ws))
    s = D._strm(ws.value)
    assert L.deflateInit2_(C.byref(s), 10, 8, 15, 8, 0, VERSION, C.sizeof(capi.ZStream)) == -2       # bad level
    assert L.deflateInit2_(C.byref(s), 6, 8, 15, 8, 0, VERSION, C.sizeof(capi.ZStream)) == 0
    obuf = (C.c_uint8 * 100)()
    s.next_out = C.cast(obuf, capi.u8p); s.avail_out = 0
    assert L.deflate(C.byref(s), Z_NO_FLUSH) == -5                                                     # no output room
    s.avail_out = 100
    assert L.deflate(C.byref(s), 6) == -2                                                              # Z_TREES is not a deflate flush
    assert L.deflate(C.byref(s), Z_FINISH) == 1
    assert L.deflate(C.byref(s), Z_NO_FLUSH) == -2                                                     # after the end
    assert bytes(obuf[:100 - s.avail_out]) and zlib.decompress(bytes(obuf[:100 - s.avail_out])) == b""
    assert L.deflateEnd(C.byref(s)) == 0
    # too small a work buffer
    s2 = D._strm(1000)
    assert L.inflateInit2_(C.byref(s2), 15, VERSION, C.sizeof(capi.ZStream)) == -2
