"""Generates the committed golden fixtures from the REFERENCE ITSELF (oracle/_ref/libzsc_ref.so, built by
oracle/Makefile from /root/reference) and from the known-answer vectors in the reference's own tests.
Run here (the container that has /root/reference); the fixtures travel, the reference does not.

  infcover_vectors.json  raw-deflate / zlib / gzip byte strings of reference test/infcover.c
                         (:367-371, :399-411, :583-613, :643-658) with the expectation written there, and the
                         result of the reference's one-shot zsc_uncompress_gzip2 on each
  bad_headers.json       the hand-built bad zlib / gzip headers of reference test/zlib_gtest.cpp:1815-1918
  resync_vectors.json    corrupted streams whose recovery depends on what inflateSync resets (window emptied, distance
                         limit back to 32768)
  ref_streams.json       small seeded inputs, compressed by the reference at several levels / strategies /
                         section sizes: compressed bytes, sizes, adler32, crc32
  ref_sizes.json         the size-check functions over the (window_bits, mem_level, level, len) grid, and
                         sizeof(deflate_state) / sizeof(inflate_state)
"""
import ctypes as C
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from refimpl import ref  # noqa: E402
from zsc_b200 import datagen  # noqa: E402

REF_SRC = "/root/reference"
R = ref()


def hexbytes(s):
    return bytes(int(t, 16) for t in s.split())


def run_ref_uncompress(data, wbits, cap=70000):
    r, out, used = R.uncompress(np.frombuffer(data, dtype=np.uint8), cap, window_bits=wbits)
    return {"ret": int(r), "produced": int(len(out)), "consumed": int(used), "out_hex": out.tobytes().hex() if len(out) <= 600 else None,
            "out_adler": int(R.adler32(out)) if len(out) else 1}


def infcover():
    src = open(os.path.join(REF_SRC, "test", "infcover.c")).read()
    src = re.sub(r'"\s*\n\s*"', "", src)          # join split string literals
    vecs = []
    for m in re.finditer(r'\btry\("([0-9a-f ]*)",\s*"([^"]*)",\s*(-?\d+)\)', src):
        hx, what, err = m.group(1), m.group(2), int(m.group(3))
        wbits = 47 if err < 0 else -15
        vecs.append({"kind": "try", "hex": hx, "what": what, "expect_err": err, "window_bits": wbits,
                     "ref": run_ref_uncompress(hexbytes(hx), wbits)})
    for m in re.finditer(r'\binf\("([0-9a-f ]*)",\s*"([^"]*)",\s*(\d+),\s*(-?\d+),\s*(\d+),\s*(\w+)\)', src):
        hx, what, step, win, ln, ret = m.groups()
        win = int(win)
        if win in (1,):            # "bad window size": parameter error of inflateInit2, not a data vector
            continue
        wb = win if win != 0 else 15
        if win == -8:
            wb = -15               # the one-shot API checks the work buffer first; window 8 streams decode the same with 15
        vecs.append({"kind": "inf", "hex": hx, "what": what, "expect_ret_streaming": ret, "window_bits": wb,
                     "ref": run_ref_uncompress(hexbytes(hx), wb)})
    return vecs


def bad_headers():
    out = []
    def zl(method=8, wb=15, bump=0):
        h = (method + ((wb - 8) << 4)) << 8
        h |= 2 << 6
        h += 31 - (h % 31)
        h += bump
        return bytes([(h >> 8) & 0xFF, h & 0xFF])
    body = bytes.fromhex("63000000010001")          # fixed block, one literal 0, adler
    cases = [("incorrect header check", zl(bump=1) + body, 15), ("unknown compression method", zl(method=5) + body, 15),
             ("invalid window size", zl(wb=16) + body, 15), ("good", zl() + body, 15),
             ("gzip bad magic", bytes([31, 140, 8, 0, 0, 0, 0, 0, 0, 3]) + bytes.fromhex("0300") + bytes(8), 31),
             ("gzip bad method", bytes([31, 139, 7, 0, 0, 0, 0, 0, 0, 3]) + bytes.fromhex("0300") + bytes(8), 31),
             ("gzip bad flags", bytes([31, 139, 8, 0x80, 0, 0, 0, 0, 0, 3]) + bytes.fromhex("0300") + bytes(8), 31),
             ("gzip good empty", bytes([31, 139, 8, 0, 0, 0, 0, 0, 0, 3]) + bytes.fromhex("0300") + bytes(8), 31),
             ("zlib stream, gzip requested", zl() + body, 31), ("auto detect zlib", zl() + body, 47)]
    for what, data, wb in cases:
        out.append({"what": what, "hex": data.hex(), "window_bits": wb, "ref": run_ref_uncompress(data, wb)})
    return out


def resync_vectors():
    """Streams whose recovery path exercises what inflateSync resets (reference src/inflate.c:1547-1604 ->
    inflateReset :295-330: empty window, dmax back to 32768): a back-reference across a resynchronisation point, and a
    small-window header whose limit no longer applies behind one.  Built with Python's zlib, answered by the reference."""
    import zlib
    rng = np.random.default_rng(1234)
    out = []
    text = datagen.fill(6000, 41, datagen.TEXT).tobytes()
    # (a) sync flushes keep the window, so the section behind the marker refers back across it; a flipped bit in the
    #     first section sends the decoder to the marker, where those distances reach behind the new start
    c = zlib.compressobj(6)
    a = c.compress(text[:3000]) + c.flush(zlib.Z_SYNC_FLUSH)
    b = c.compress(text[1000:4000]) + c.flush(zlib.Z_FULL_FLUSH)
    d = c.compress(text[3000:]) + c.flush()
    for flip in (10, 40, len(a) // 2):
        bad = bytearray(a + b + d)
        bad[flip] ^= 0x04
        out.append({"what": f"back-reference across a resync point (flip at {flip})", "hex": bytes(bad).hex(), "window_bits": 15})
    # (b) header says window_bits 9; the body (raw deflate made with a 32 KiB window) has a first section with short
    #     distances, a full flush, then a block that repeats 1000 random bytes at distance 1000 (> 512)
    blob = rng.integers(0, 256, 1000, dtype=np.uint8).tobytes()
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    s1 = c.compress(b"abcabcabcabcabc" * 20) + c.flush(zlib.Z_FULL_FLUSH)
    s2 = c.compress(blob + blob) + c.flush()
    hdr = (8 + ((9 - 8) << 4)) << 8
    hdr += 31 - (hdr % 31)
    head = bytes([hdr >> 8, hdr & 0xFF])
    trailer = zlib.adler32(b"abcabcabcabcabc" * 20 + blob + blob).to_bytes(4, "big")
    good = head + s1 + s2 + trailer
    out.append({"what": "window_bits 9 header, distance 1000 behind it: too far back", "hex": good.hex(), "window_bits": 15})
    bad = bytearray(good)
    bad[2] |= 0x06                                   # block type 3 in the first block header: a certain data error
    out.append({"what": "window_bits 9 header, first section corrupted: the limit is 32768 again behind the resync point", "hex": bytes(bad).hex(), "window_bits": 15})
    for x in out:
        x["ref"] = run_ref_uncompress(bytes.fromhex(x["hex"]), x["window_bits"], cap=20000)
        x["ref"]["out_hex"] = None
    return out


def ref_streams():
    inputs = {
        "mixed20k": datagen.fill(20000, 11, datagen.MIXED),
        "text9k": datagen.fill(9000, 12, datagen.TEXT),
        "telem12k": datagen.fill(12288, 1000, datagen.TELEMETRY, piece=12288),
        "random3k": datagen.fill(3000, 5, datagen.RANDOM),
        "zeros5k": np.zeros(5000, np.uint8),
        "ff300": np.full(300, 255, np.uint8),
        "one": np.array([0x41], np.uint8),
        "empty": np.zeros(0, np.uint8),
        "abc_rep": np.frombuffer((b"abcabcabcabd" * 400)[:4099], dtype=np.uint8).copy(),
    }
    out = {"inputs": {}, "streams": []}
    for name, x in inputs.items():
        out["inputs"][name] = {"hex": x.tobytes().hex(), "adler32": int(R.adler32(x)), "crc32": int(R.crc32(x))}
        combos = [(1, 0, 100000, 15), (6, 0, 100000, 15), (9, 0, 100000, 15), (0, 0, 100000, 15), (6, 0, 4096, 15),
                  (6, 2, 100000, 15), (6, 3, 100000, 15), (6, 4, 100000, 15), (6, 1, 100000, 15), (6, 0, 100000, -15),
                  (6, 0, 100000, 31), (6, 0, 5000, 9)]
        for level, strat, mbl, wb in combos:
            r, c = R.compress(x, mbl, level, window_bits=wb, strategy=strat)
            out["streams"].append({"input": name, "level": level, "strategy": strat, "max_block_len": mbl, "window_bits": wb,
                                   "ret": int(r), "size": int(len(c)), "hex": c.tobytes().hex()})
    return out


def ref_sizes():
    rows = []
    for wb in list(range(8, 16)) + [-15, -9, 24, 31, 7, 16, 0, 500]:
        for ml in (0, 1, 5, 8, 9, 10):
            r, v = R.compress_work_size(wb, ml)
            rows.append({"fn": "cwork", "wb": wb, "ml": ml, "ret": int(r), "val": int(v) if r == 0 else None})
        r, v = R.uncompress_work_size(wb)
        rows.append({"fn": "uwork", "wb": wb, "ret": int(r), "val": int(v) if r == 0 else None})
    for n in (0, 1, 100, 4096, 152089, 1 << 20, (1 << 30)):
        for mbl in (1, 100, 100000, 262144, 1 << 30):
            if n // mbl > 1 << 22:
                continue
            for level in (0, 1, 6, 9):
                for wb, ml in ((15, 8), (15, 9), (9, 1), (-15, 8), (31, 8), (12, 8)):
                    r, v = R.max_output_size(n, mbl, level, wb, ml)
                    rows.append({"fn": "bound", "n": n, "mbl": mbl, "level": level, "wb": wb, "ml": ml, "ret": int(r),
                                 "val": int(v) if r == 0 else None})
    return {"rows": rows, "sizeof_deflate_state": int(R.L.refprobe_sizeof_deflate_state()),
            "sizeof_inflate_state": int(R.L.refprobe_sizeof_inflate_state()),
            "pinned_by_reference_log": {"alice_bound_L6": 152160, "alice_bound_L0": 173502, "note": "test/output/Test.log:26,262"}}


if __name__ == "__main__":
    which = sys.argv[1:] or ["infcover_vectors", "bad_headers", "resync_vectors", "ref_streams", "ref_sizes"]
    for name, fn in (("infcover_vectors", infcover), ("bad_headers", bad_headers), ("resync_vectors", resync_vectors), ("ref_streams", ref_streams), ("ref_sizes", ref_sizes)):
        if name not in which:
            continue
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=0)
        print("wrote", name)
