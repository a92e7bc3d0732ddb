#!/usr/bin/env python
"""bench.py — benchmark of the zsc-b200 engine.  Headline: BASELINE.json configs[1].

Workload of the headline (per GPU): 1 GiB synthetic mixed text/binary (tools/datagen.c, seed 1 + rank), compressed
as ONE zsc_compress-shaped stream at level 1 with max_block_len = 256 KiB (4096 independently
decodable sections).  A "step" is one full pass of the deflate path over that buffer.

  value      input GB/s, inputs resident in HBM, CUDA events on the engine's stream, max over ranks
  e2e        the same pass through the host-buffer C-ABI call behind zsc_compress
             (zscgpu_compress_host: H2D of the 1 GiB, all kernels, D2H of the compressed stream), plus zsc_compress
             itself with pinned and with pageable caller buffers, plus the per-rank copy rates that bound it
  roofline   the LZ77 kernel (dominant): (N + C) algorithmic bytes / its event-timed duration
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the reference's own zsc_compress2 (oracle/_ref, compiled from /root/reference) on the
             host cores, bounded sample of the same workload

Side legs in the same JSON line (skipped with --no-side), each with its own parity check, roofline and CPU baseline:
  canterbury_shaped   configs[0]: the reference's Performance test shape through zsc_compress / zsc_uncompress (N = 1)
  deflate_levels_6_9  configs[2]: 4096 telemetry buffers at levels 6 and 9, buffers partitioned over the ranks
  inflate_batched     configs[3]: 16 GiB of reference-compressed streams, streams partitioned over the ranks
  checksums, strategies   configs[4]: adler32 / crc32 over 8 GiB sharded over the ranks + host combine; Z_HUFFMAN_ONLY / Z_RLE
  one_stream_over_ranks   one 2 GiB logical stream split over the GPUs by section range, stitched on the host (N > 1)

`--impl reference` times only the reference arm.  Launch with torchrun for --gpus > 1 (one rank per GPU; the headline
gives every rank its own 1 GiB — the path shards with no collective, scaling "weak" — the side legs split a fixed job).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

GIB = 1 << 30
SECTION = 262144
LEVEL = 1
METRIC = "deflate_level1_input_GBps"
WORKLOAD = "configs[1]: 1 GiB synthetic mixed text/binary per GPU, zlib level 1, max_block_len 256 KiB (4096 sections), one stream"


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the LZ kernel from the committed `ncu --set full` capture, if any."""
    p = os.path.join(ROOT, "profiles", "lz_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch_1GiB")
        except Exception:
            return None
    return None


def reference_arm(nbytes_per_job, jobs, threads, level=LEVEL, data=None):
    """Time the reference's zsc_compress2 over `jobs` independent slices on `threads` host threads."""
    from refimpl import REF_PATH
    from zsc_b200 import datagen
    L = C.CDLL(REF_PATH, mode=C.RTLD_LOCAL)
    u64p, u32p, i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    L.refprobe_batch.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_void_p, u64p, u32p, C.c_void_p, u64p, u32p, u32p, i32p,
                                 C.c_uint32, C.c_int32, C.c_int32]
    total = nbytes_per_job * jobs
    if data is None:
        data = datagen.mixed(total, seed=1)
    cap = nbytes_per_job + nbytes_per_job // 8 + 4096
    dst = np.empty(cap * jobs, dtype=np.uint8)
    so = (C.c_uint64 * jobs)(*[i * nbytes_per_job for i in range(jobs)])
    sl = (C.c_uint32 * jobs)(*[nbytes_per_job] * jobs)
    do = (C.c_uint64 * jobs)(*[i * cap for i in range(jobs)])
    dc = (C.c_uint32 * jobs)(*[cap] * jobs)
    dl = (C.c_uint32 * jobs)()
    rt = (C.c_int32 * jobs)()

    def once():
        t = time.perf_counter()
        L.refprobe_batch(0, threads, jobs, data.ctypes.data, so, sl, dst.ctypes.data, do, dc, dl, rt, SECTION, level, 0)
        dt = time.perf_counter() - t
        assert all(r == 0 for r in rt), "reference zsc_compress2 failed"
        return dt, sum(dl)
    return once, total


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = cores
    job = 4 << 20
    jobs = max(threads, min(256, (threads * 16 << 20) // job))
    once, total = reference_arm(job, jobs, threads)
    for _ in range(min(args.warmup, 1)):
        once()
    ts = []
    for _ in range(args.steps):
        dt, csz = once()
        ts.append(dt)
    t = sum(ts) / len(ts)
    v = total / 1e9 / t
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "level": LEVEL, "max_block_len": SECTION,
                   "note": "reference zsc_compress2 (unmodified, oracle/_ref) on host cores; each step = bounded sample"},
        "cpu_baseline": {"value": round(v, 4), "unit": "GB/s", "cores": threads, "kind": "reference",
                         "sample": f"{jobs} independent {job >> 20} MiB slices of the workload ({total >> 20} MiB), one pthread per core"},
        "e2e": {"value": round(v, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ratio": round(total / csz, 4),
    }
    print(json.dumps(line), flush=True)


def rank_range(n, rank, world):
    """contiguous share of n units for this rank (GPU g gets [g*ceil(n/G), (g+1)*ceil(n/G)))"""
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def ref_batch(kind, data, src_off, src_len, dst_cap, threads, level=6, mbl=SECTION, strategy=0):
    """n independent reference calls (kind 0 zsc_compress2, 1 zsc_uncompress) on `threads` host threads, timed"""
    from refimpl import REF_PATH
    L = C.CDLL(REF_PATH, mode=C.RTLD_LOCAL)
    u64p, u32p, i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    L.refprobe_batch.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_void_p, u64p, u32p, C.c_void_p, u64p, u32p, u32p, i32p,
                                 C.c_uint32, C.c_int32, C.c_int32]
    n = len(src_off)
    doff = np.concatenate([[0], np.cumsum(np.asarray(dst_cap, dtype=np.uint64))]).astype(np.uint64)
    dst = np.empty(int(doff[-1]) + 16, dtype=np.uint8)
    so = (C.c_uint64 * n)(*[int(v) for v in src_off]); sl = (C.c_uint32 * n)(*[int(v) for v in src_len])
    do = (C.c_uint64 * n)(*[int(v) for v in doff[:-1]]); dc = (C.c_uint32 * n)(*[int(v) for v in dst_cap])
    dl = (C.c_uint32 * n)(); rt = (C.c_int32 * n)()
    t = time.perf_counter()
    L.refprobe_batch(kind, threads, n, data.ctypes.data, so, sl, dst.ctypes.data, do, dc, dl, rt, mbl, level, strategy)
    dt = time.perf_counter() - t
    assert all(r == 0 for r in rt), "reference call failed"
    return dt, list(dl)


def inflate_config3(device, rank, world, barrier, allmax, peak):
    """BASELINE configs[3]: batched inflate of 16 GiB of reference-compressed streams — 512 unique 256 KiB buffers
    (half configs[1] mixed data, half configs[2] telemetry) compressed by the reference at levels 1/6/9 in equal
    thirds, replicated 128x in device memory = 65 536 independent zlib streams per pass, partitioned over the ranks by
    contiguous stream ranges (strong scaling: the 16 GiB are the whole job).  The reference is only the producer of the
    inputs here (as the config demands); the thing timed is zscgpu_inflate_batch's kernels."""
    from concurrent.futures import ThreadPoolExecutor
    from zsc_b200 import Engine, datagen
    import refimpl
    uniq, rep_all, S, slot = 512, 128, 262144, 160000
    r_lo, r_hi = rank_range(rep_all, rank, world)
    rep = r_hi - r_lo
    n = uniq * rep
    x = np.concatenate([datagen.mixed(uniq // 2 * S, seed=1), datagen.telemetry_buffers(uniq - uniq // 2, S, seed=1000)])
    E = Engine(raw_bytes=n * S + (1 << 20), comp_bytes=n * slot + (1 << 20), deflate_batch_max=uniq * S + (1 << 20),
               max_streams=max(n, 64), max_chunks=uniq + 16, device=device)
    try:
        if refimpl.have_ref():
            R = refimpl.ref()
            with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
                comps = list(ex.map(lambda i: R.compress(x[i * S:(i + 1) * S], S, (1, 6, 9)[i % 3])[1], range(uniq)))
            producer = "reference zsc_compress (oracle/_ref), levels 1/6/9 in thirds"
        else:
            E.upload(0, 0, x)
            st0 = Engine.make_streams([i * S for i in range(uniq)], [S] * uniq, [i * slot for i in range(uniq)], [slot] * uniq)
            res0 = E.deflate(st0, S, 6)
            comps = [E.download(1, i * slot, res0[i].produced) for i in range(uniq)]
            producer = "this engine at level 6 (reference library not built)"
        one = np.zeros(uniq * slot, np.uint8)
        offs1, off = [], 0
        for c in comps:
            one[off:off + len(c)] = c
            offs1.append((off, len(c)))
            off += (len(c) + 15) & ~15
        offs = []
        for r_ in range(rep):
            E.upload(1, r_ * off, one[:off])
            offs += [(r_ * off + o, l) for o, l in offs1]
        st = Engine.make_streams([i * S for i in range(n)], [S] * n, [o[0] for o in offs], [o[1] for o in offs])
        E.inflate_enqueue(st, 1)
        res = E.fetch(n)
        bad = sum(1 for r in res if r.ret != 0 or r.produced != S)
        same = all(bool(np.array_equal(E.download(0, k * uniq * S, uniq * S), x)) for k in (0, rep - 1))
        ts = []
        for _ in range(3):
            barrier(); E.sync()
            E.event(0); E.relaunch(); E.event(1); E.sync(); ts.append(allmax(E.elapsed_ms(0, 1)))
        t = sum(ts) / len(ts)
        csum1 = sum(len(c) for c in comps)
        n_all, csum = uniq * rep_all, csum1 * rep_all
        out = {"value": round(n_all * S / 1e6 / t, 2), "unit": "GB/s of output", "ms": round(t, 2), "streams": n_all, "output_bytes": n_all * S,
               "compressed_bytes": csum, "producer": producer, "launches_per_pass": 3, "scaling": "strong", "streams_per_rank": n,
               "parity": "all streams Z_OK with the trailer adler32 verified; first and last replica bit-exact with the inputs" if bad == 0 and same else f"MISMATCH ({bad} bad streams)",
               "roofline": {"bound": "hbm", "kernel": "zs_inflate_spec_kernel<false>" if n > 888 else "zs_inflate_pipe_kernel", "achieved": round((n_all * S + csum) / world / 1e6 / t, 1), "peak": peak, "unit": "GB/s",
                            "frac": round((n_all * S + csum) / world / 1e6 / t / peak, 5), "algorithmic_bytes": (n_all * S + csum) // world,
                            "note": "C + N per pass and GPU over the whole pass (decode, output adler32, check kernels)"}}
        if world == 1:
            # the same streams in narrower batches (a warp decodes one stream with all 32 lanes at every width, inflate_spec.h):
            # beyond 888 streams the wide build (28 streams per SM), up to 592 one warp per stream with the stream's window in
            # shared memory, up to 444 (and again up to 888, in two turns) with a second warp per stream that writes one round
            # while the first decodes the next (zs_inflate_pipe_kernel)
            nb = {}
            for k in (64, 512, 4096):
                if k > n:
                    continue
                stk = Engine.make_streams([i * S for i in range(k)], [S] * k, [o[0] for o in offs[:k]], [o[1] for o in offs[:k]])
                E.inflate_enqueue(stk, 1)
                resk = E.fetch(k)
                okk = all(r.ret == 0 and r.produced == S for r in resk) and bool(np.array_equal(E.download(0, 0, min(k, uniq) * S), x[:min(k, uniq) * S]))
                tk = []
                for _ in range(3):
                    E.sync(); E.event(0); E.relaunch(); E.event(1); E.sync(); tk.append(E.elapsed_ms(0, 1))
                nb[str(k)] = {"value": round(k * S / 1e6 / min(tk), 2), "ms": round(min(tk), 3), "parity": "bit-exact" if okk else "MISMATCH"}
            out["narrower_batches"] = dict(nb, unit="GB/s of output", kernel="the speculative warp decoder: zs_inflate_pipe_kernel at 64 streams (two warps per stream), zs_inflate_spec_kernel<true> at 512, <false> at 4096")
        if rank == 0 and refimpl.have_ref():
            cores = os.cpu_count() or 1
            packed = np.concatenate(comps)
            poff = np.concatenate([[0], np.cumsum([len(c) for c in comps])])[:-1]
            dt, prods = ref_batch(1, packed, poff, [len(c) for c in comps], [S] * uniq, cores)
            out["cpu_baseline"] = {"value": round(uniq * S / 1e9 / dt, 3), "unit": "GB/s of output", "cores": cores, "kind": "reference",
                                   "sample": f"the {uniq} unique streams ({uniq * S >> 20} MiB of output) through the reference's zsc_uncompress, one pthread per core"}
        return out
    finally:
        E.close()


def deflate_config2(device, rank, world, barrier, allmax, peak):
    """BASELINE configs[2]: 4096 independent 256 KiB telemetry-like buffers, one stream each, at levels 6 and 9 (hash
    chains + lazy parse, the chain kernel), device resident, partitioned over the ranks by contiguous buffer ranges
    (strong scaling); every stream goes back through the reference's zsc_uncompress, size against the reference's on
    the first 64 buffers of rank 0."""
    from zsc_b200 import Engine, datagen
    import refimpl
    nall, S, slot = 4096, 262144, 300000
    lo, hi = rank_range(nall, rank, world)
    nbuf = hi - lo
    n = nbuf * S
    x = datagen.fill(n, 1000 + lo, datagen.TELEMETRY, piece=S)       # buffer i is seeded with 1000 + i whatever the rank
    E = Engine(raw_bytes=n + (1 << 20), comp_bytes=nbuf * slot + (1 << 20), deflate_batch_max=n + (1 << 20), max_streams=nbuf, max_chunks=nbuf + 16, device=device)
    out = {}
    try:
        E.upload(0, 0, x)
        st = Engine.make_streams([i * S for i in range(nbuf)], [S] * nbuf, [i * slot for i in range(nbuf)], [slot] * nbuf)
        for level in (6, 9):
            E.deflate_enqueue(st, S, level)
            res = E.fetch(nbuf)
            assert all(r.ret == 0 for r in res)
            csize = sum(r.produced for r in res)
            barrier(); E.sync()
            E.event(0); E.relaunch(); E.event(1); E.sync()
            ms = allmax(E.elapsed_ms(0, 1))
            lz_ms = E.elapsed_ms(9, 10)
            row = {"value": round(nall * S / 1e6 / ms, 3), "unit": "GB/s", "ms": round(ms, 2), "ratio": round(n / csize, 4),
                   "roofline": {"bound": "hbm", "kernel": "zs_lzc_kernel", "achieved": round((n + csize) / 1e6 / lz_ms, 2), "peak": peak, "unit": "GB/s",
                                "frac": round((n + csize) / 1e6 / lz_ms / peak, 5), "kernel_ms": round(lz_ms, 2), "algorithmic_bytes": int(n + csize)}}
            if refimpl.have_ref():
                R = refimpl.ref()
                k = min(64, nbuf)
                ref_c = sum(len(R.compress(x[i * S:(i + 1) * S], S, level)[1]) for i in range(k)) if rank == 0 else 1
                ours = sum(res[i].produced for i in range(k))
                if rank == 0:
                    row["size_vs_reference"] = round(ours / ref_c, 4)
                # every stream of this rank through the reference's own zsc_uncompress (host threads, untimed)
                comp_all = E.download(1, 0, nbuf * slot)
                rets, prods, outb = refimpl.ref_uncompress_batch(comp_all, [i * slot for i in range(nbuf)], [r.produced for r in res], [S] * nbuf)
                okn = sum(1 for i in range(nbuf) if rets[i] == 0 and prods[i] == S)
                same = bool(np.array_equal(outb, x))
                row["parity"] = f"reference inflates {okn}/{nbuf} streams of rank 0" + (", all bytes equal the inputs" if same else ", BYTES MISMATCH")
            out[str(level)] = row
        res_line = {"workload": "configs[2]: 4096 x 256 KiB telemetry-like buffers, one zlib stream each", "scaling": "strong", "buffers_per_rank": nbuf, "levels": out}
        if rank == 0 and refimpl.have_ref():
            cores = os.cpu_count() or 1
            k = min(nbuf, max(cores, 32))
            dt, _ = ref_batch(0, x, [i * S for i in range(k)], [S] * k, [slot] * k, cores, level=6)
            res_line["cpu_baseline"] = {"value": round(k * S / 1e9 / dt, 4), "unit": "GB/s", "cores": cores, "kind": "reference", "level": 6,
                                        "sample": f"the first {k} buffers through the reference's zsc_compress2 at level 6, one pthread per core"}
        return res_line
    finally:
        E.close()


def canterbury_config0(device):
    """BASELINE configs[0], the reference's own Performance test (test/zlib_gtest.cpp:2400-2892): 11 Canterbury-shaped
    buffers, max_block_len 100 000, level 6, zsc_compress then zsc_uncompress — through the zsc_pub.h entry points
    themselves on ordinary (pageable) host buffers, timed per call with the host clock, next to the reference on one
    host thread in the same run.  At these sizes a call is launch- and latency-bound on the GPU; the line says so."""
    from zsc_b200 import datagen, zsc
    import refimpl
    Z = zsc()
    R = refimpl.ref() if refimpl.have_ref() else None
    bufs = datagen.canterbury_shaped()
    Z.compress(bufs[0], 100000, 6)                       # first call initialises the process-wide engine
    rows, tot = [], {"gpu_c": 0.0, "gpu_u": 0.0, "ref_c": 0.0, "ref_u": 0.0, "size": 0, "ref_size": 0, "bytes": 0}
    ok = True
    for x in bufs:
        t = time.perf_counter(); r, comp = Z.compress(x, 100000, 6); tc = time.perf_counter() - t
        t = time.perf_counter(); r2, back, used = Z.uncompress(comp, len(x)); tu = time.perf_counter() - t
        ok = ok and r == 0 and r2 == 0 and bool(np.array_equal(back, x))
        row = {"bytes": len(x), "compressed": len(comp), "compress_ms": round(tc * 1e3, 3), "uncompress_ms": round(tu * 1e3, 3)}
        tot["gpu_c"] += tc; tot["gpu_u"] += tu; tot["size"] += len(comp); tot["bytes"] += len(x)
        if R is not None:
            t = time.perf_counter(); rr, rcomp = R.compress(x, 100000, 6); rtc = time.perf_counter() - t
            t = time.perf_counter(); rr2, rback, rused = R.uncompress(comp, len(x)); rtu = time.perf_counter() - t     # the reference inflates OUR stream
            ok = ok and rr == 0 and rr2 == 0 and bool(np.array_equal(rback, x))
            row.update({"ref_compressed": len(rcomp), "ref_compress_ms": round(rtc * 1e3, 3), "ref_uncompress_ms": round(rtu * 1e3, 3)})
            tot["ref_c"] += rtc; tot["ref_u"] += rtu; tot["ref_size"] += len(rcomp)
        rows.append(row)
    out = {"workload": "configs[0]: 11 Canterbury-shaped buffers (2 810 784 B), max_block_len 100 000, level 6, zsc_compress + zsc_uncompress on pageable host buffers",
           "compress_MBps": round(tot["bytes"] / 1e6 / tot["gpu_c"], 1), "uncompress_MBps": round(tot["bytes"] / 1e6 / tot["gpu_u"], 1),
           "compressed_bytes": tot["size"], "parity": "round trip bit-exact; the reference inflates every stream" if ok else "MISMATCH", "buffers": rows}
    if R is not None:
        out["cpu_baseline"] = {"compress_MBps": round(tot["bytes"] / 1e6 / tot["ref_c"], 1), "uncompress_MBps": round(tot["bytes"] / 1e6 / tot["ref_u"], 1),
                               "cores": 1, "kind": "reference", "sample": "the same 11 calls through the reference on one host thread", "compressed_bytes": tot["ref_size"]}
        out["size_vs_reference"] = round(tot["size"] / tot["ref_size"], 4)
        out["verdict"] = ("compress: GPU %.1f x one host core; uncompress: GPU %.2f x one host core (calls of 4 KB - 1 MB: launch latency, the PCIe round "
                          "trip and one warp per 100 000-byte section bound the GPU side)" % (tot["ref_c"] / tot["gpu_c"], tot["ref_u"] / tot["gpu_u"]))
    return out


def checksums_config4(device, rank, world, barrier, allmax, peak):
    """BASELINE configs[4], first half: adler32 / crc32 over 8 GiB of uniform-random bytes (seed 5), sharded over the
    ranks in contiguous ranges and combined on the host (zscgpu_adler32_combine / zscgpu_crc32_combine); on one GPU the
    8 GiB are additionally checked as 8 shards + combine.  Bit-exact against Python's zlib on a strided sample and, for
    the whole buffer, against the value combined from independently computed pieces."""
    from zsc_b200 import Engine, datagen, shard
    import zlib
    total = 8 << 30
    lo, hi = rank_range(total >> 20, rank, world)
    n = (hi - lo) << 20
    E = Engine(raw_bytes=n + (1 << 20), comp_bytes=1 << 20, deflate_batch_max=1 << 20, max_streams=64, max_chunks=64, device=device)
    try:
        piece = 1 << 30
        host_a, host_c = 1, 0                                  # zlib's running values over this rank's whole range (host, untimed)
        for off in range(0, n, piece):
            m = min(piece, n - off)
            x = datagen.fill(m, 5 + ((lo << 20) + off) // piece, datagen.RANDOM)
            E.upload(0, off, x)
            host_a, host_c = zlib.adler32(x, host_a), zlib.crc32(x, host_c)
        ok = True
        res = {}
        for name, fn, enq in (("adler32", E.adler32, E.L.zscgpu_adler32_enqueue), ("crc32", E.crc32, E.L.zscgpu_crc32_enqueue)):
            v = fn(0, n)
            ts = []
            for _ in range(3):
                barrier(); E.sync()
                E.event(0); enq(E.h, 0, n); E.event(1); E.sync(); ts.append(allmax(E.elapsed_ms(0, 1)))
            t = min(ts)
            # the same value from 8 shards combined on the host
            cuts = [i * (n // 8) + (i % 3) for i in range(8)] + [n]
            acc = 1 if name == "adler32" else 0
            for a_, b_ in zip(cuts[:-1], cuts[1:]):
                acc = (shard.adler32_combine if name == "adler32" else shard.crc32_combine)(acc, fn(a_, b_ - a_), b_ - a_)
            ok = ok and acc == v and v == (host_a if name == "adler32" else host_c)
            res[name] = {"value": round(total / 1e6 / t, 1), "unit": "GB/s", "ms": round(t, 3), "rank0_value": int(v),
                         "roofline": {"bound": "hbm", "achieved": round(n / 1e6 / t, 1), "peak": peak, "unit": "GB/s", "frac": round(n / 1e6 / t / peak, 4),
                                      "algorithmic_bytes": n}}
        combined = None
        if world > 1:
            # the final combine of the sharded job: per-rank values gathered (12 bytes per rank) and folded on the host
            import torch
            import torch.distributed as dist
            mine = torch.tensor([res["adler32"]["rank0_value"], res["crc32"]["rank0_value"], n, host_a, host_c], dtype=torch.int64, device=f"cuda:{device}")
            allv = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine)
            ca, cc, ha, hc = 1, 0, 1, 0
            for v_ in allv:
                a_, c_, n_, ha_, hc_ = (int(t) for t in v_)
                ca, cc = shard.adler32_combine(ca, a_, n_), shard.crc32_combine(cc, c_, n_)
                ha, hc = shard.adler32_combine(ha, ha_, n_), shard.crc32_combine(hc, hc_, n_)
            ok = ok and (ca, cc) == (ha, hc)
            combined = {"adler32": ca, "crc32": cc}
        out = {"workload": "configs[4]: adler32 / crc32 over 8 GiB of random bytes, contiguous shards per rank + host combine", "scaling": "strong",
               "bytes_per_rank": n, "combined_over_ranks": combined,
               "parity": "bit-exact with zlib over the whole range; the same value from 8 shards combined on the host" + ("; per-rank values combined over the ranks equal zlib's" if world > 1 else "") if ok else "MISMATCH", **res}
        if rank == 0:
            import refimpl
            if refimpl.have_ref():
                R = refimpl.ref()
                y = datagen.fill(256 << 20, 5, datagen.RANDOM)
                t0 = time.perf_counter(); R.adler32(y); ta = time.perf_counter() - t0
                t0 = time.perf_counter(); R.crc32(y); tc = time.perf_counter() - t0
                out["cpu_baseline"] = {"adler32": round(len(y) / 1e9 / ta, 3), "crc32": round(len(y) / 1e9 / tc, 3), "unit": "GB/s", "cores": 1, "kind": "reference",
                                       "sample": "256 MiB through the reference's adler32_z / crc32_z on one host thread"}
        return out
    finally:
        E.close()


def strategies_config4(E, data, n, st, cap):
    """BASELINE configs[4], second half: the 1 GiB of configs[1] at Z_HUFFMAN_ONLY and Z_RLE (level 6), resident; size
    against the reference's on the first 64 MiB, a 64 MiB prefix of each stream through the reference's inflate."""
    import refimpl
    out = {}
    for name, strat in (("Z_HUFFMAN_ONLY", 2), ("Z_RLE", 3)):
        E.deflate_enqueue(st, SECTION, 6, strat)
        res = E.fetch(1)
        assert res[0].ret == 0
        ts = []
        for _ in range(2):
            E.event(0); E.relaunch(); E.event(1); E.sync(); ts.append(E.elapsed_ms(0, 1))
        row = {"value": round(n / 1e6 / min(ts), 2), "unit": "GB/s", "ms": round(min(ts), 2), "ratio": round(n / res[0].produced, 4)}
        if refimpl.have_ref():
            R = refimpl.ref()
            m = 64 << 20
            p = DeflateSample(E, data, m, strat)
            rc, refc = R.compress(data[:m], SECTION, 6, strategy=strat)
            rr, back, used = R.uncompress(p, m)
            row["size_vs_reference"] = round(len(p) / len(refc), 4)
            row["parity"] = "reference inflates the 64 MiB sample bit-exact" if rr == 0 and bool(np.array_equal(back, data[:m])) else "MISMATCH"
        out[name] = row
    return {"workload": "configs[4]: the 1 GiB of configs[1] at level 6, Z_HUFFMAN_ONLY / Z_RLE, device resident", **out}


def DeflateSample(E, data, m, strat):
    """the first m bytes of the resident buffer as their own stream at level 6 / `strat` -> compressed bytes"""
    from zsc_b200 import Engine
    st = Engine.make_streams([0], [m], [0], [m + (m >> 3) + 4096])
    r = E.deflate(st, SECTION, 6, strat)[0]
    assert r.ret == 0
    return E.download(1, 0, r.produced)


def one_stream_over_ranks(E, data_all_fn, rank, world, dist, local_rank):
    """north_star / SURVEY 8e: ONE logical stream split over the GPUs by contiguous section ranges; every rank deflates
    its range as a raw part, the parts are gathered on rank 0, concatenated behind a zlib header with the adler32s
    folded on the host (zscgpu_adler32_combine), and the stitched stream goes through the reference's zsc_uncompress."""
    import torch
    from zsc_b200 import Engine, shard
    import refimpl
    total = 2 << 30                                          # a 2 GiB logical stream, 8192 sections
    nsec = total // SECTION
    lo, hi = shard.partition(nsec, world)[rank]
    b0, b1 = shard.byte_range(lo, hi, SECTION, total)
    x = data_all_fn(b0, b1)
    n = b1 - b0
    E.upload(0, 0, x)
    cap = n + (n >> 3) + 4096
    st = Engine.make_streams([0], [n], [0], [cap])
    part = (1 if rank > 0 else 0) | (2 if rank < world - 1 else 0)
    t0 = time.perf_counter()
    r = E.deflate(st, SECTION, LEVEL, 0, 0, 15, part)[0]
    dt = time.perf_counter() - t0
    assert r.ret == 0
    comp = E.download(1, 0, r.produced)
    meta = torch.tensor([r.produced, r.check, n], dtype=torch.int64, device=f"cuda:{local_rank}")
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    sizes = [int(m[0]) for m in metas]
    mx = max(sizes)
    buf = torch.zeros(mx, dtype=torch.uint8, device=f"cuda:{local_rank}")
    buf[:r.produced] = torch.from_numpy(comp).to(buf.device)
    bufs = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, bufs, dst=0)                              # collection of the finished parts, not a data-path collective
    if rank != 0:
        return None
    parts = [bufs[i][:sizes[i]].cpu().numpy().tobytes() for i in range(world)]
    stream = np.frombuffer(shard.stitch(parts, [int(m[1]) for m in metas], [int(m[2]) for m in metas], LEVEL), np.uint8)
    res = {"workload": f"one {total >> 30} GiB logical stream ({nsec} sections) split over {world} engines by section range, parts stitched on the host",
           "compressed_bytes": int(len(stream)), "deflate_GBps_slowest_rank": round(n / 1e9 / dt, 2)}
    if refimpl.have_ref():
        whole = np.concatenate([data_all_fn(*shard.byte_range(l_, h_, SECTION, total)) for l_, h_ in shard.partition(nsec, world)])
        rr, out, used = refimpl.ref().uncompress(stream, total)
        res["parity"] = ("reference zsc_uncompress inflates the stitched stream bit-exact (adler32 trailer folded on the host)"
                         if rr == 0 and used == len(stream) and bool(np.array_equal(out, whole)) else f"MISMATCH (reference returned {rr})")
    return res


def bind_to_gpu_numa_node(gpu):
    """Run this rank (and first-touch its host buffers) on the CPU socket its GPU hangs off: with one rank per GPU the
    end-to-end path moves 1.5 GB per step and rank over PCIe, and buffers on the far socket cross the socket link."""
    try:
        try:
            import torch
            pr = torch.cuda.get_device_properties(gpu)     # CUDA's own numbering of the devices
            bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        except Exception:
            import pynvml
            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(gpu)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
            if len(bus.split(":")[0]) == 8:
                bus = bus[4:]                              # sysfs uses a 4-digit PCI domain
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_gpu(args, rank, world, local_rank):
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    from zsc_b200 import Engine, EngineConfig, datagen, DeflateParams, Result, lib, zsc
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peak, peak_src = measured_peak()
    n = GIB
    E = Engine(raw_bytes=n + (1 << 20), comp_bytes=n + (n >> 3) + (1 << 20), deflate_batch_max=n + (1 << 20),
               max_streams=8192, max_chunks=8192, device=local_rank)
    data = datagen.mixed(n, seed=1 + rank)
    cap = n + (n >> 3)
    dest = np.empty(cap, dtype=np.uint8)
    E.L.zscgpu_host_register(data.ctypes.data, data.nbytes)
    E.L.zscgpu_host_register(dest.ctypes.data, dest.nbytes)
    E.upload(0, 0, data)
    st = Engine.make_streams([0], [n], [0], [cap])

    # ---- warm-up ----
    E.deflate_enqueue(st, SECTION, LEVEL)
    res = E.fetch(1)
    assert res[0].ret == 0, f"deflate failed: {res[0].ret}"
    csize = res[0].produced
    for _ in range(max(args.warmup - 1, 0)):
        E.relaunch()
    E.sync()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = int(E.L.zscgpu_launch_total(E.h))
    # ---- timed: device-resident ----
    barrier(); E.sync()
    t_wall0 = time.perf_counter()
    step_ms, lz_ms, parts = [], [], []
    for _ in range(args.steps):
        E.event(0); E.relaunch(); E.event(1); E.sync()
        step_ms.append(E.elapsed_ms(0, 1))
        lz_ms.append(E.elapsed_ms(9, 10))
        parts.append([E.elapsed_ms(8 + i, 9 + i) for i in range(5)])
    E.sync(); barrier()
    t_wall = time.perf_counter() - t_wall0
    total_ms = allmax(sum(step_ms))
    # ---- timed: end to end through the host-buffer call (pinned caller buffers) ----
    e2e_steps = max(1, min(args.steps, 5))
    p = DeflateParams(SECTION, LEVEL, 0, 1, 15, 0)
    r1 = Result()
    launches_resident = int(E.L.zscgpu_launch_total(E.h)) - launches0
    E.L.zscgpu_compress_host(E.h, dest.ctypes.data, cap, data.ctypes.data, n, C.byref(p), 0, C.byref(r1))   # warm
    launches1 = int(E.L.zscgpu_launch_total(E.h))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rc = E.L.zscgpu_compress_host(E.h, dest.ctypes.data, cap, data.ctypes.data, n, C.byref(p), 0, C.byref(r1))
        assert rc == 0 and r1.ret == 0
    barrier()
    my_e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_s = allmax(my_e2e_s)
    launches = launches_resident + int(E.L.zscgpu_launch_total(E.h)) - launches1     # kernels inside the two timed regions
    # per-rank host<->device rates, measured alone on this rank's buffers while every rank does the same (names the e2e limiter)
    barrier(); t0 = time.perf_counter(); E.upload(0, 0, data); h2d_s = time.perf_counter() - t0
    barrier(); t0 = time.perf_counter(); E.L.zscgpu_download(E.h, 1, dest.ctypes.data, 0, int(r1.produced)); d2h_s = time.perf_counter() - t0
    copy_rates = {"h2d_GBps_per_rank_all_ranks_copying": round(n / 1e9 / allmax(h2d_s), 1),
                  "d2h_GBps_per_rank_all_ranks_copying": round(int(r1.produced) / 1e9 / allmax(d2h_s), 1),
                  "h2d_GBps_sum": round(world * n / 1e9 / allmax(h2d_s), 1)}
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- the zsc_pub.h call itself (zsc_compress on the process-wide engine), pageable and pinned caller buffers ----
    api = None
    try:
        cfg = EngineConfig()
        lib().zscgpu_default_config(C.byref(cfg))
        cfg.device = local_rank
        lib().zscgpu_global_init(C.byref(cfg))
        Z = zsc()
        wl = Z.compress_work_size()[1]
        work = np.empty(wl, np.uint8)
        src_pageable = data.copy() if world == 1 else None       # a fresh allocation nobody registered
        dst_pageable = np.empty(cap, np.uint8) if world == 1 else None
        api = {}
        for name, sbuf, dbuf in (("pinned", data, dest), ("pageable", src_pageable, dst_pageable)):
            if sbuf is None:
                continue
            dl = C.c_uint32(cap)
            args_ = (dbuf.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(dl), sbuf.ctypes.data_as(C.POINTER(C.c_uint8)), n, SECTION,
                     work.ctypes.data_as(C.POINTER(C.c_uint8)), wl, LEVEL)
            assert Z.L.zsc_compress(*args_) == 0                  # warm
            barrier(); t0 = time.perf_counter()
            for _ in range(2):
                dl.value = cap
                assert Z.L.zsc_compress(*args_) == 0
            barrier()
            api[name] = round(world * n / 1e9 / allmax((time.perf_counter() - t0) / 2), 3)
        api["what"] = "zsc_compress (zsc_pub.h) on the process-wide engine, 1 GiB per call and rank, caller buffers pinned / ordinary pageable memory"
        lib().zscgpu_global_shutdown()
    except Exception as ex:  # pragma: no cover
        api = {"error": repr(ex)}

    # ---- the way back: this run's own stream through the section-parallel inflate, resident and end to end ----
    inflate = None
    try:
        comp = dest[:r1.produced]
        back = np.empty(n, dtype=np.uint8)
        E.L.zscgpu_host_register(back.ctypes.data, back.nbytes)
        E.upload(1, 0, comp)
        st1 = Engine.make_streams([0], [n], [0], [len(comp)])
        ri = E.inflate_sectioned(st1, 1)                                     # warm
        assert ri.ret == 0 and ri.produced == n and ri.consumed == len(comp)
        barrier(); t0 = time.perf_counter()
        for _ in range(2):
            ri = E.inflate_sectioned(st1, 1)
        barrier(); inf_s = allmax((time.perf_counter() - t0) / 2)
        r2 = Result()
        E.L.zscgpu_uncompress_host(E.h, back.ctypes.data, n, comp.ctypes.data, len(comp), 1, C.byref(r2))   # warm
        barrier(); t0 = time.perf_counter()
        for _ in range(2):
            rc = E.L.zscgpu_uncompress_host(E.h, back.ctypes.data, n, comp.ctypes.data, len(comp), 1, C.byref(r2))
            assert rc == 0 and r2.ret == 0
        barrier(); inf_e2e_s = allmax((time.perf_counter() - t0) / 2)
        inflate = {"value": round(world * n / 1e9 / inf_s, 3), "unit": "GB/s of output",
                   "e2e": round(world * n / 1e9 / inf_e2e_s, 3),
                   "what": "the 1 GiB stream this run produced (4096 sections) back through zscgpu_inflate_sectioned / zscgpu_uncompress_host, per rank",
                   "parity": "bit-exact with the input" if bool(np.array_equal(back, data)) else "MISMATCH"}
        E.L.zscgpu_host_unregister(back.ctypes.data)
        del back
    except Exception as ex:  # pragma: no cover
        inflate = {"value": None, "error": repr(ex)}

    # ---- parity of this run's output: the whole 1 GiB stream (4096 sections, header, adler32 trailer) through the
    # reference's own zsc_uncompress, outside the timed regions; Python's zlib when oracle/_ref is absent ----
    parity = "unchecked"
    try:
        from refimpl import have_ref, ref
        if have_ref():
            rr, out, used = ref().uncompress(dest[:r1.produced], n)
            okp = rr == 0 and used == r1.produced and bool(np.array_equal(out, data))
            parity = f"reference zsc_uncompress inflates the whole stream ({-(-n // SECTION)}/{-(-n // SECTION)} sections) bit-exact" if okp else f"MISMATCH (reference returned {rr})"
            del out
        else:
            import zlib
            okp = zlib.decompress(dest[:r1.produced].tobytes()) == data.tobytes()
            parity = "python zlib inflates the whole stream bit-exact (oracle/_ref absent)" if okp else "MISMATCH"
    except Exception as ex:  # pragma: no cover
        parity = f"check failed: {ex!r}"
    parity_ok = allsum(0.0 if parity.startswith("MISMATCH") or parity.startswith("check failed") else 1.0)

    side = {}
    if not args.no_side:
        # configs[4] second half on the resident 1 GiB (N = 1), then the engine of the headline leg is released
        if world == 1:
            try:
                side["strategies"] = strategies_config4(E, data, n, st, cap)
            except Exception as ex:  # pragma: no cover
                side["strategies"] = {"error": repr(ex)}
        if world > 1:
            try:
                E.L.zscgpu_host_unregister(data.ctypes.data)
                side["one_stream_over_ranks"] = one_stream_over_ranks(
                    E, lambda b0, b1: _slice_mixed(datagen, 77, b0, b1), rank, world, dist, local_rank)
            except Exception as ex:  # pragma: no cover
                side["one_stream_over_ranks"] = {"error": repr(ex)}
    E.close()
    del dest
    if not args.no_side:
        for key, fn in (("deflate_levels_6_9", deflate_config2), ("inflate_batched", inflate_config3), ("checksums", checksums_config4)):
            try:
                side[key] = fn(local_rank, rank, world, barrier, allmax, peak)
            except Exception as ex:  # pragma: no cover
                side[key] = {"error": repr(ex)}
        if world == 1:
            try:
                side["canterbury_shaped"] = canterbury_config0(local_rank)
            except Exception as ex:  # pragma: no cover
                side["canterbury_shaped"] = {"error": repr(ex)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    ms_per_step = total_ms / args.steps
    value = world * n / 1e9 / (ms_per_step / 1e3)
    lz = sum(lz_ms) / len(lz_ms)
    alg_bytes = n + csize
    achieved = alg_bytes / 1e9 / (lz / 1e3)
    pk = [sum(p_[i] for p_ in parts) / len(parts) for i in range(5)]
    if world > 1:
        parity += f" (all {world} ranks)" if parity_ok == world else f" ({int(parity_ok)}/{world} ranks OK)"
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "level": LEVEL, "max_block_len": SECTION, "bytes_per_gpu": n,
                   "compressed_bytes": int(csize), "ratio": round(n / csize, 4),
                   "l2": "inputs (1 GiB) and symbol scratch (1.3 GiB) far exceed the 126 MB L2; no flush needed",
                   "parity": parity, "wall_s_timed_region": round(t_wall, 3),
                   "host_numa_node_of_rank0": numa},
        "roofline": {"bound": "hbm", "kernel": "zs_lz_kernel", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 5), "traffic": ncu_traffic(), "traffic_source": "profiles/lz_kernel_traffic.json (ncu --set full of this kernel; see its commit field)",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": round(lz, 3),
                     "kernel_share_of_step": round(lz / ms_per_step, 4),
                     "stage_ms": {"adler32": round(pk[0], 3), "lz77": round(pk[1], 3), "block_codes": round(pk[2], 3),
                                  "offsets": round(pk[3], 3), "bitpack": round(pk[4], 3)}},
        "e2e": {"value": round(world * n / 1e9 / e2e_s, 3), "unit": "GB/s", "h2d_bytes_per_step": world * n,
                "d2h_bytes_per_step": world * int(r1.produced), "steps": e2e_steps,
                "api": "zscgpu_compress_host (the call behind zsc_compress), pinned host buffers", "zsc_compress": api,
                "copy_rates": copy_rates},
        "gpu_launches": launches,
        "inflate": inflate,
        "clocks": sampler.summary(),
    }
    line.update(side)
    if world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            job = 4 << 20
            jobs = max(cores, min(256, (cores * 16 << 20) // job))
            once, total = reference_arm(job, jobs, cores, data=data[:job * jobs] if job * jobs <= n else None)
            dt, csz = once()
            line["cpu_baseline"] = {"value": round(total / 1e9 / dt, 4), "unit": "GB/s", "cores": cores, "kind": "reference",
                                    "sample": f"first {total >> 20} MiB of the workload as {jobs} independent {job >> 20} MiB zsc_compress2 calls, one pthread per core",
                                    "ratio": round(total / csz, 4)}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {ex!r}"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def _slice_mixed(datagen, seed, b0, b1):
    """bytes [b0, b1) of the mixed workload with this seed (tools/datagen.c generates 1 MiB pieces from (seed, piece index))"""
    piece = 1 << 20
    return datagen.fill(-(-b1 // piece) * piece, seed, datagen.MIXED)[b0:b1].copy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="zsc_b200", choices=["zsc_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", "--no-inflate-batch", dest="no_side", action="store_true",
                    help="skip the side legs (configs[0] Canterbury-shaped, configs[2] levels 6/9, configs[3] 16 GiB batched inflate, configs[4] checksums and strategies, one stream over N GPUs)")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
