"""CPU-side tests (`-m "not gpu"`): the C-ABI library loads and exports every declared symbol, the pure
arithmetic entry points match the reference, argument validation returns the reference's codes before any
GPU work, and the product's __host__ __device__ logic (decode core, Huffman stage, offset algebra, LZ77
parse model) is correct when compiled for the host."""
import ctypes as C
import json
import os
import re
import zlib

import numpy as np
import pytest

import refimpl
from conftest import has_gpu
from zsc_b200 import capi, datagen, shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def load(name):
    return json.load(open(os.path.join(G, name)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    for s in capi.ZSCGPU_SYMBOLS + capi.ZSC_SYMBOLS:
        assert hasattr(L, s), s
    # every function name declared in the public headers is in the lists above (no undeclared drift)
    hdr = open(os.path.join(ROOT, "include", "zscgpu.h")).read()
    declared = set(re.findall(r"\b(zscgpu_[a-z0-9_]+)\s*\(", hdr))
    assert declared <= set(capi.ZSCGPU_SYMBOLS), declared - set(capi.ZSCGPU_SYMBOLS)
    hdr = open(os.path.join(ROOT, "include", "zsc", "zsc_pub.h")).read()
    declared = set(re.findall(r"\b(zsc_[a-z0-9_]+)\s*\(", hdr))
    assert declared <= set(capi.ZSC_SYMBOLS), declared - set(capi.ZSC_SYMBOLS)


def test_section_size_guess_of_the_one_pass_inflate():
    """zscgpu_guess_section_size: for every stream zsc_compress can make (sections of max_block_len, a shorter last
    one) with a round max_block_len and an exact capacity the guess is max_block_len; it never leaves the range the
    section count allows; it declines when nothing round fits."""
    L = capi.lib()
    f = L.zscgpu_guess_section_size
    f.argtypes = [C.c_uint64, C.c_uint32]; f.restype = C.c_uint64
    rng = np.random.default_rng(3)
    for mbl in (4096, 65536, 100000 // 32 * 32 * 8, 262144, 1 << 20, 3 << 18):
        for _ in range(200):
            k = int(rng.integers(8, 5000))
            last = int(rng.integers(1, mbl + 1))
            total = (k - 1) * mbl + last
            if total >= 1 << 32:
                continue
            g = f(total, k)
            assert g == 0 or ((k - 1) * g < total <= k * g)
            if mbl & (mbl - 1) == 0:
                assert g == mbl, (mbl, k, last, g)       # a power of two is always the roundest value of its range
    assert f((1 << 30), 4096) == 262144 and f((1 << 30) + 5, 4097) == 262144
    assert f(10, 2) == 0 and f(5, 8) == 0 and f(1000, 1) == 0
    assert f(8 * 1000 + 1, 9) == 0                      # the range [889, 1000] holds no multiple of 256
    # the decimal twin: a max_block_len like the reference's own Performance test uses (100 000, test/zlib_gtest.cpp:2400-2892)
    f10 = L.zscgpu_guess_section_size10
    f10.argtypes = [C.c_uint64, C.c_uint32]; f10.restype = C.c_uint64
    for mbl in (1000, 50000, 100000, 250000, 3000000):
        for _ in range(200):
            k = int(rng.integers(2, 3000))
            total = (k - 1) * mbl + int(rng.integers(1, mbl + 1))
            if total >= 1 << 32:
                continue
            g = f10(total, k)
            assert g == 0 or ((k - 1) * g < total <= k * g)
            if k >= 300:
                assert g == mbl, (mbl, k, total, g)     # (with few sections another equally round number may share the range)
    assert f10(1029744, 11) == 100000 and f10(152089, 2) == 100000 and f10(10, 2) == 0


def test_size_check_functions_match_reference_fixture():
    Z = capi.zsc()
    for row in load("ref_sizes.json")["rows"]:
        if row["fn"] == "cwork":
            r, v = Z.compress_work_size(row["wb"], row["ml"])
        elif row["fn"] == "uwork":
            r, v = Z.uncompress_work_size(row["wb"])
        else:
            r, v = Z.max_output_size(row["n"], row["mbl"], row["level"], row["wb"], row["ml"])
        assert r == row["ret"], row
        if r == 0:
            assert v == row["val"], row
    assert Z.max_output_size(152089, 100000, 6)[1] == 152160       # reference test/output/Test.log:26
    assert Z.max_output_size(152089, 100000, 0)[1] == 173502       # :262
    assert Z.compress_work_size()[1] == 333600 and Z.uncompress_work_size()[1] == 39920


def test_bounds_with_gzip_header_fields():
    """reference test/zlib_gtest.cpp:1403-1409: a gzip name adds strlen + 1 to the bound"""
    Z = capi.zsc()
    name = (C.c_uint8 * 6)(*b"Hello\0")
    gz = capi.GzHeader()
    r0, b0 = Z.max_output_size(10000, 1000, 6, 31, 8, gz)
    gz.name = C.cast(name, capi.u8p)
    r1, b1 = Z.max_output_size(10000, 1000, 6, 31, 8, gz)
    assert (r0, r1) == (0, 0) and b1 == b0 + 6
    # function work size <= compile-time macro for every legal parameter pair (zlib_gtest.cpp:1417-1435)
    for wb in range(9, 16):
        for ml in range(1, 10):
            macro = 6400 + (1 << wb) * 2 + (1 << wb) * 2 * 2 + (1 << (ml + 7)) * 2 + (1 << (ml + 6)) * 4
            assert Z.compress_work_size(wb, ml)[1] <= macro
        assert Z.uncompress_work_size(wb)[1] <= 7600 + (1 << wb)


def test_argument_errors_are_reported_before_any_gpu_work():
    """reference test/zlib_gtest.cpp:1439-1612 (ZSCCompressErrors / ZSCUncompressErrors)"""
    Z = capi.zsc()
    x = np.arange(1000, dtype=np.uint8)
    assert Z.compress(x, 100, 6, work_len=0)[0] == capi.Z_MEM_ERROR
    assert Z.compress(x, 100, 6, window_bits=500, dest_cap=2000, work_len=400000)[0] == capi.Z_STREAM_ERROR
    assert Z.compress(x, 100, 6, mem_level=10, dest_cap=2000, work_len=400000)[0] == capi.Z_STREAM_ERROR
    assert Z.compress(x, 100, 6, strategy=7, dest_cap=2000)[0] == capi.Z_STREAM_ERROR
    assert Z.compress(x, 100, 11, dest_cap=2000)[0] == capi.Z_STREAM_ERROR
    assert Z.uncompress(x, 1000, work_len=0)[0] == capi.Z_MEM_ERROR
    assert Z.uncompress(x, 1000, window_bits=500, work_len=50000)[0] == capi.Z_STREAM_ERROR
    gz = capi.GzHeader()
    assert Z.uncompress(x, 1000, window_bits=15, gz=gz)[0] == capi.Z_STREAM_ERROR     # header request on a zlib stream
    assert Z.compress(x, 100, 6, gz=gz, dest_cap=2000)[0] == capi.Z_STREAM_ERROR       # gzip header on a zlib stream
    L = capi.lib()
    assert L.zError(-3) == b"data error" and L.zError(-5) == b"buffer error" and L.zError(2) == b"need dictionary"
    assert L.zlibVersion().startswith(b"1.2.11")
    assert L.adler32_z(0, None, 0) == 1 and L.crc32_z(0, None, 0) == 0                 # src/adler32.c:82-84, src/crc32.c:507


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own zsc_compress2 on the host cores, oracle/_ref): one JSON line
    with the metric of the GPU arm, impl = reference, a cpu_baseline describing the run and an e2e equal to the value."""
    import subprocess
    import sys
    if not refimpl.have_ref():
        pytest.skip("oracle/_ref not present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "deflate_level1_input_GBps" and line["unit"] == "GB/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "u8"
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and 2.0 < line["ratio"] < 2.6


DEATH_CASES = [
    ("zsc_compress", "source", "L.zsc_compress(buf, C.byref(n), None, 100, 100, work, 400000, 6)"),
    ("zsc_compress", "dest", "L.zsc_compress(None, C.byref(n), buf, 100, 100, work, 400000, 6)"),
    ("zsc_compress", "dest_len", "L.zsc_compress(buf, None, buf, 100, 100, work, 400000, 6)"),
    ("zsc_compress", "work", "L.zsc_compress(buf, C.byref(n), buf, 100, 100, None, 400000, 6)"),
    ("zsc_compress", "max_block_len", "L.zsc_compress(buf, C.byref(n), buf, 100, 0, work, 400000, 6)"),
    ("zsc_uncompress", "source", "L.zsc_uncompress(buf, C.byref(n), None, C.byref(m), work, 400000)"),
    ("zsc_uncompress", "source_len", "L.zsc_uncompress(buf, C.byref(n), buf, None, work, 400000)"),
    ("zsc_uncompress", "dest", "L.zsc_uncompress(None, C.byref(n), buf, C.byref(m), work, 400000)"),
    ("zsc_uncompress", "dest_len", "L.zsc_uncompress(buf, None, buf, C.byref(m), work, 400000)"),
    ("zsc_uncompress", "work", "L.zsc_uncompress(buf, C.byref(n), buf, C.byref(m), None, 400000)"),
    ("zsc_compress_get_max_output_size", "size_out", "L.zsc_compress_get_max_output_size(1000, 100, 6, None)"),
    ("zsc_compress_get_min_work_buf_size", "size_out", "L.zsc_compress_get_min_work_buf_size(None)"),
    ("zsc_uncompress_get_min_work_buf_size", "size_out", "L.zsc_uncompress_get_min_work_buf_size(None)"),
]


@pytest.mark.parametrize("fn,param,call", DEATH_CASES, ids=[f"{c[0]}-{c[1]}" for c in DEATH_CASES])
def test_null_arguments_fire_the_assertion_naming_the_parameter(fn, param, call):
    """The reference's death tests (test/zlib_gtest.cpp:2105-2390): a NULL source / dest / dest_len / work /
    source_len / size_out, or max_block_len == 0, fires ZSC_ASSERT with the parameter in the message — before any
    GPU work, so this holds on a machine without one.  Each case runs in its own process."""
    import subprocess
    import sys
    code = (
        "import ctypes as C, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from zsc_b200 import capi\n"
        "L = capi.lib()\n"
        "buf = (C.c_uint8 * 4096)(); work = (C.c_uint8 * 400000)()\n"
        "n = C.c_uint32(4096); m = C.c_uint32(100)\n"
        "for f in ('zsc_compress', 'zsc_uncompress', 'zsc_compress_get_max_output_size', 'zsc_compress_get_min_work_buf_size', 'zsc_uncompress_get_min_work_buf_size'):\n"
        "    getattr(L, f).argtypes = None\n"
        f"{call}\n"
        "print('survived')\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "survived" not in r.stdout, (r.returncode, r.stdout, r.stderr)
    assert "Assertion" in r.stderr and re.search(rf"\b{param} != ", r.stderr), r.stderr


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_a_gpu(capfd):
    """With no B200 the codec entry points must fail loudly, not fall back to a CPU path."""
    Z = capi.zsc()
    x = np.arange(4096, dtype=np.uint8)
    r, out = Z.compress(x, 1000, 6)
    assert r == capi.Z_MEM_ERROR and len(out) == 0
    r, out, used = Z.uncompress(x, 4096)
    assert r == capi.Z_MEM_ERROR and len(out) == 0
    with pytest.raises(RuntimeError):
        capi.Engine(raw_bytes=1 << 20, comp_bytes=1 << 20)
    assert b"no CPU fallback" in capi.lib().zscgpu_last_error(None) or b"sm_" in capi.lib().zscgpu_last_error(None)


def test_checksum_combination_operators():
    L = capi.lib()
    rng = np.random.default_rng(1)
    data = rng.integers(0, 256, 200000, dtype=np.uint8).tobytes()
    for cut in (0, 1, 5552, 65521, 100000, 199999, 200000):
        a, b = data[:cut], data[cut:]
        assert L.zscgpu_adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(data)
        assert L.zscgpu_crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(data)
    # worst case for deferred modulo: all 0xFF, length not a multiple of 16
    ff = b"\xff" * 70001
    parts = [ff[i:i + 8751] for i in range(0, len(ff), 8751)]
    acc = 1
    for p in parts:
        acc = L.zscgpu_adler32_combine(acc, zlib.adler32(p), len(p))
    assert acc == zlib.adler32(ff)


# ---------------------------------------------------------------- product decode core on the host
def vec_bytes(x):
    return bytes(int(t, 16) for t in x["hex"].split()) if "kind" in x else bytes.fromhex(x["hex"])


def test_decode_core_on_known_answer_vectors():
    for x in load("infcover_vectors.json") + load("bad_headers.json") + load("resync_vectors.json"):
        if x["window_bits"] not in (15, -15):
            continue
        r, out, info = refimpl.h_inflate(np.frombuffer(vec_bytes(x), np.uint8), 70000, wrap=1 if x["window_bits"] == 15 else 0)
        ref = x["ref"]
        if ref["ret"] == 0 and x["window_bits"] == 15:
            assert info["have_check"] == 1 and info["stored_check"] == ref["out_adler"]
        if r == 0 and x["window_bits"] == 15 and info["have_check"] and info["stored_check"] != zlib.adler32(out.tobytes()):
            r = -3          # the data check is the caller's (a separate adler32 pass on the GPU, inflate.cu zs_inflate_check_kernel)
        assert (r, len(out), info["consumed"]) == (ref["ret"], ref["produced"], ref["consumed"]), x["what"]
        if ref["out_hex"] is not None:
            assert out.tobytes().hex() == ref["out_hex"]


def test_decode_core_on_reference_streams():
    g = load("ref_streams.json")
    inputs = {k: np.frombuffer(bytes.fromhex(v["hex"]), dtype=np.uint8) for k, v in g["inputs"].items()}
    for s in g["streams"]:
        if s["window_bits"] not in (15, -15, 9):
            continue
        x = inputs[s["input"]]
        comp = np.frombuffer(bytes.fromhex(s["hex"]), dtype=np.uint8)
        r, out, info = refimpl.h_inflate(comp, len(x) + 8, wrap=0 if s["window_bits"] < 0 else 1)
        assert r == 0 and info["consumed"] == len(comp) and np.array_equal(out, x), s["input"]
        # output one byte short -> Z_BUF_ERROR with the buffer filled (reference src/inflate.c:1400-1402)
        if len(x) > 1:
            r, out, info = refimpl.h_inflate(comp, len(x) - 1, wrap=0 if s["window_bits"] < 0 else 1)
            assert r == -5 and np.array_equal(out, x[:len(x) - 1])


# ---------------------------------------------------------------- deflate model (LZ parse + Huffman stage + framing)
@pytest.mark.parametrize("level,strategy", [(1, 0), (6, 0), (0, 0), (6, 2), (6, 3), (6, 4), (6, 1), (3, 0), (4, 0), (9, 0)])
def test_deflate_model_streams_inflate_through_the_oracle(level, strategy):
    O = refimpl.oracle()
    cases = [datagen.fill(300000, 31, datagen.MIXED), datagen.fill(70000, 1000, datagen.TELEMETRY, piece=70000),
             datagen.fill(20000, 5, datagen.RANDOM), np.zeros(100000, np.uint8), np.zeros(0, np.uint8),
             np.frombuffer(b"a", np.uint8), np.frombuffer(b"abcabcabcabc" * 50, np.uint8)]
    for x in cases:
        for mbl in (100000, 4096, 1 << 30):
            comp, syms = refimpl.model_deflate(x, mbl, level, strategy)
            r, out, used = O.uncompress(comp, len(x) + 4)
            assert r == 0 and used == len(comp) and np.array_equal(out, x), (len(x), mbl)
            assert np.array_equal(np.frombuffer(zlib.decompress(comp.tobytes()), np.uint8), x)
            # The reference's bound budgets 4 bytes per section; a truly independent incompressible section
            # costs 10 (stored header + marker) — the reference itself only fits because its loop merges
            # sections when the output chunk fills (SURVEY.md §7 "reference quirk"), see DESIGN.md.
            bound = capi.zsc().max_output_size(len(x), mbl, level)[1]
            if mbl >= 65536:
                assert len(comp) <= bound, (len(comp), bound)
            # sections are independently decodable: one full-flush marker per section boundary
            nsec = max(1, -(-len(x) // mbl))
            assert comp.tobytes().count(b"\x00\x00\xff\xff") >= nsec - 1


@pytest.mark.skipif(not refimpl.have_ref(), reason="oracle/_ref not built")
def test_deflate_model_ratio_within_two_percent_of_reference():
    R = refimpl.ref()
    for kind, seed, piece in ((datagen.MIXED, 1, 1 << 20), (datagen.TELEMETRY, 1000, 262144)):
        x = datagen.fill(2 << 20, seed, kind, piece=piece)
        for level in (1, 6):
            comp, _ = refimpl.model_deflate(x, 262144, level)
            rc, refc = R.compress(x, 262144, level)
            assert rc == 0 and len(comp) <= 1.02 * len(refc), (kind, level, len(comp), len(refc))
            r, out, used = R.uncompress(comp, len(x))
            assert r == 0 and used == len(comp) and np.array_equal(out, x)


def test_offset_algebra_scan_equals_sequential_walk():
    """zk_elem composition (one prefix scan) must give the same bit offsets as walking the blocks in order."""
    L = refimpl.harness()
    L.h_zk_apply_seq.argtypes = [refimpl.u32p] * 4 + [C.c_uint32, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.integers(1, 40))
        types = rng.integers(0, 4, n).astype(np.uint32)
        body = rng.integers(0, 400000, n).astype(np.uint32)
        inlen = rng.integers(0, 60000, n).astype(np.uint32)
        flags = rng.integers(0, 8, n).astype(np.uint32)
        offs = (C.c_uint64 * (n + 1))()
        end = C.c_uint64(0)
        x0 = int(rng.integers(0, 1000)) * 8
        L.h_zk_apply_seq(types.ctypes.data_as(refimpl.u32p), body.ctypes.data_as(refimpl.u32p), inlen.ctypes.data_as(refimpl.u32p),
                         flags.ctypes.data_as(refimpl.u32p), n, 1, x0, offs, C.byref(end))
        assert end.value == offs[n]


# ---------------------------------------------------------------- multi-GPU sharding logic (gloo, world_size 2)
def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mbl, level = 50000, 6
    x = datagen.fill(777777, 9, datagen.MIXED)                       # every rank regenerates the same input
    nsec = -(-len(x) // mbl)
    lo, hi = shard.partition(nsec, world)[rank]
    b0, b1 = shard.byte_range(lo, hi, mbl, len(x))
    # stand-in for the per-rank GPU engine (wrap=0, part bit 1 unless last): raw deflate, full flush per section
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    part = b""
    for s in range(lo, hi):
        c0, c1 = shard.byte_range(s, s + 1, mbl, len(x))
        part += co.compress(x[c0:c1].tobytes())
        part += co.flush(zlib.Z_FINISH if s == nsec - 1 else zlib.Z_FULL_FLUSH)
    mine = (part, zlib.adler32(x[b0:b1].tobytes()), b1 - b0)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        stream = shard.stitch([g[0] for g in gathered], [g[1] for g in gathered], [g[2] for g in gathered], level)
        ok = zlib.decompress(stream) == x.tobytes()
        r, out, used = refimpl.oracle().uncompress(np.frombuffer(stream, np.uint8), len(x))
        q.put(bool(ok and r == 0 and used == len(stream) and np.array_equal(out, x)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_stream_stitching_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_partition_covers_all_sections_contiguously():
    for n in (0, 1, 7, 4096, 4097):
        for w in (1, 2, 4, 8):
            parts = shard.partition(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1


def test_code_lengths_are_complete_and_length_limited():
    """zh_lengths: Kraft sum exactly 1 (the reference's inflate rejects anything else, src/inftrees.c:168-177),
    never longer than the limit, every used symbol coded — including Fibonacci-like tables that overflow 15 / 7 bits."""
    L = refimpl.harness()
    L.h_lengths.argtypes = [refimpl.u32p, C.c_int, C.c_int, refimpl.u8p]
    rng = np.random.default_rng(0)
    for trial in range(3000):
        n, mb = ((19, 7), (30, 15), (286, 15))[trial % 3]
        k = int(rng.integers(0, n + 1))
        f = np.zeros(n, np.uint32)
        idx = rng.choice(n, k, replace=False)
        mode = trial % 4
        if mode == 0:
            f[idx] = rng.integers(1, 5, k)
        elif mode == 1:
            f[idx] = (2 ** rng.integers(0, 13, k)).astype(np.uint32)
        elif mode == 2:
            f[idx] = rng.integers(1, 8192, k)
        else:
            a, b = 1, 1
            for i in idx[:40]:
                f[i] = min(a, 8000)
                a, b = b, a + b
        ln = np.zeros(n, np.uint8)
        L.h_lengths(f.ctypes.data_as(refimpl.u32p), n, mb, ln.ctypes.data_as(refimpl.u8p))
        used = ln[ln > 0].astype(np.int64)
        assert len(used) >= 2 and used.max() <= mb
        assert int(np.sum(1 << (mb - used))) == 1 << mb, (trial, n, mb)
        assert all(ln[i] > 0 for i in range(n) if f[i] > 0)
