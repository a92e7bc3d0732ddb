#!/usr/bin/env python
"""bench.py — headline benchmark of the zsc-b200 engine (BASELINE.json configs[1]).

Workload (per GPU): 1 GiB synthetic mixed text/binary (tools/datagen.c, seed 1 + rank), compressed
as ONE zsc_compress-shaped stream at level 1 with max_block_len = 256 KiB (4096 independently
decodable sections).  A "step" is one full pass of the deflate path over that buffer.

  value      input GB/s, inputs resident in HBM, CUDA events on the engine's stream, max over ranks
  e2e        the same pass through the host-buffer C-ABI call behind zsc_compress
             (zscgpu_compress_host: H2D of the 1 GiB, all kernels, D2H of the compressed stream)
  roofline   the LZ77 kernel (dominant): (N + C) algorithmic bytes / its event-timed duration
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the reference's own zsc_compress2 (oracle/_ref, compiled from /root/reference) on the
             host cores, bounded sample of the same workload

`--impl reference` times only the reference arm.  Launch with torchrun for --gpus > 1 (one rank per
GPU, independent 1 GiB per rank: the path shards with no collective, scaling is "weak").
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


def inflate_config4(device):
    """BASELINE configs[3]: batched inflate of 16 GiB of reference-compressed streams — 512 unique 256 KiB buffers
    (half configs[1] mixed data, half configs[2] telemetry) compressed by the reference at levels 1/6/9 in equal
    thirds, replicated 128x in device memory = 65 536 independent zlib streams per pass.  The reference is only the
    producer of the inputs here (as the config demands); the thing timed is zscgpu_inflate_batch's kernels."""
    from concurrent.futures import ThreadPoolExecutor
    from zsc_b200 import Engine, datagen
    import refimpl
    uniq, rep, S, slot = 512, 128, 262144, 160000
    n = uniq * rep
    x = np.concatenate([datagen.mixed(uniq // 2 * S, seed=1), datagen.telemetry_buffers(uniq - uniq // 2, S, seed=1000)])
    E = Engine(raw_bytes=n * S + (1 << 20), comp_bytes=n * slot + (1 << 20), deflate_batch_max=uniq * S + (1 << 20),
               max_streams=n, max_chunks=uniq + 16, device=device)
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
            E.event(0); E.relaunch(); E.event(1); E.sync(); ts.append(E.elapsed_ms(0, 1))
        t = sum(ts) / len(ts)
        csum = sum(len(c) for c in comps) * rep
        return {"value": round(n * S / 1e6 / t, 2), "unit": "GB/s of output", "ms": round(t, 2), "streams": n, "output_bytes": n * S,
                "compressed_bytes": csum, "producer": producer, "launches_per_pass": 3,
                "parity": "all streams Z_OK with the trailer adler32 verified; first and last replica bit-exact with the inputs" if bad == 0 and same else f"MISMATCH ({bad} bad streams)",
                "roofline": {"bound": "hbm", "achieved": round((n * S + csum) / 1e6 / t, 1), "unit": "GB/s", "algorithmic_bytes": n * S + csum}}
    finally:
        E.close()


def deflate_config2(device):
    """BASELINE configs[2]: 4096 independent 256 KiB telemetry-like buffers, one stream each, at levels 6 and 9 (hash
    chains + lazy parse, the chain kernel), device resident; ratio against the reference on the first 64 buffers."""
    from zsc_b200 import Engine, datagen
    import refimpl
    nbuf, S, slot = 4096, 262144, 300000
    n = nbuf * S
    x = datagen.telemetry_buffers(nbuf, S, seed=1000)
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
            E.event(0); E.relaunch(); E.event(1); E.sync()
            ms = E.elapsed_ms(0, 1)
            row = {"value": round(n / 1e6 / ms, 3), "unit": "GB/s", "ms": round(ms, 2), "ratio": round(n / csize, 4)}
            if refimpl.have_ref():
                R = refimpl.ref()
                k = 64
                ref_c = sum(len(R.compress(x[i * S:(i + 1) * S], S, level)[1]) for i in range(k))
                ours = sum(res[i].produced for i in range(k))
                row["size_vs_reference"] = round(ours / ref_c, 4)
                # every one of the 4096 streams through the reference's own zsc_uncompress (host threads, untimed)
                comp_all = E.download(1, 0, nbuf * slot)
                rets, prods, outb = refimpl.ref_uncompress_batch(comp_all, [i * slot for i in range(nbuf)], [r.produced for r in res], [S] * nbuf)
                okn = sum(1 for i in range(nbuf) if rets[i] == 0 and prods[i] == S)
                same = bool(np.array_equal(outb, x))
                row["parity"] = f"reference inflates {okn}/{nbuf} streams" + (", all bytes equal the inputs" if same else ", BYTES MISMATCH")
            out[str(level)] = row
        return {"workload": "configs[2]: 4096 x 256 KiB telemetry-like buffers, one zlib stream each", "levels": out}
    finally:
        E.close()


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
    from zsc_b200 import Engine, datagen, DeflateParams, Result
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

    # ---- warm-up (also the parity spot-check of this very run) ----
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
    # ---- timed: end to end through the host-buffer call ----
    e2e_steps = max(1, min(args.steps, 3))
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
    e2e_s = allmax((time.perf_counter() - t0) / e2e_steps)
    launches = launches_resident + int(E.L.zscgpu_launch_total(E.h)) - launches1     # kernels inside the two timed regions

    # ---- the way back: this run's own stream through the section-parallel inflate, resident and end to end ----
    inflate = None
    try:
        if world > 1:
            raise NotImplementedError
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
                   "what": "the 1 GiB stream this run produced (4096 sections) back through zscgpu_inflate_sectioned / zscgpu_uncompress_host",
                   "parity": "bit-exact with the input" if bool(np.array_equal(back, data)) else "MISMATCH"}
    except NotImplementedError:
        inflate = {"value": None, "note": "side measurement, taken at N=1 only"}
    except Exception as ex:  # pragma: no cover
        inflate = {"value": None, "error": repr(ex)}
    sampler.stop_flag = True
    sampler.join(timeout=2)
    levels = None
    if world == 1 and not args.no_inflate_batch:
        try:
            levels = deflate_config2(local_rank)
        except Exception as ex:  # pragma: no cover
            levels = {"error": repr(ex)}
    inflate4 = None
    if world == 1 and not args.no_inflate_batch:
        try:
            inflate4 = inflate_config4(local_rank)
        except Exception as ex:  # pragma: no cover
            inflate4 = {"value": None, "error": repr(ex)}

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

    if rank != 0:
        return
    ms_per_step = total_ms / args.steps
    value = world * n / 1e9 / (ms_per_step / 1e3)
    peak, peak_src = measured_peak()
    lz = sum(lz_ms) / len(lz_ms)
    alg_bytes = n + csize
    achieved = alg_bytes / 1e9 / (lz / 1e3)
    pk = [sum(p_[i] for p_ in parts) / len(parts) for i in range(5)]
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "level": LEVEL, "max_block_len": SECTION, "bytes_per_gpu": n,
                   "compressed_bytes": int(csize), "ratio": round(n / csize, 4),
                   "l2": "inputs (1 GiB) and symbol scratch (1.3 GiB) far exceed the 126 MB L2; no flush needed",
                   "parity": parity, "wall_s_timed_region": round(t_wall, 3),
                   "host_numa_node_of_rank0": numa},
        "roofline": {"bound": "hbm", "kernel": "zs_lz_kernel<false>", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 5), "traffic": ncu_traffic(), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": round(lz, 3),
                     "kernel_share_of_step": round(lz / ms_per_step, 4),
                     "stage_ms": {"adler32": round(pk[0], 3), "lz77": round(pk[1], 3), "block_codes": round(pk[2], 3),
                                  "offsets": round(pk[3], 3), "bitpack": round(pk[4], 3)}},
        "e2e": {"value": round(world * n / 1e9 / e2e_s, 3), "unit": "GB/s", "h2d_bytes_per_step": world * n,
                "d2h_bytes_per_step": world * int(r1.produced), "api": "zscgpu_compress_host (the call behind zsc_compress), pinned host buffers"},
        "gpu_launches": launches,
        "inflate": inflate,
        "inflate_batched": inflate4,
        "deflate_levels_6_9": levels,
        "clocks": sampler.summary(),
    }
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
    E.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="zsc_b200", choices=["zsc_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inflate-batch", action="store_true", help="skip the side measurements (configs[2] levels 6/9, configs[3] 16 GiB batched inflate)")
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
