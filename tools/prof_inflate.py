"""Inflate throughput probe: N reference-compressed 256 KiB streams (levels 1/6/9 in thirds), replicated.
usage: prof_inflate.py UNIQ REP   (build variants of the engine with tools/build_variant.sh and pick one with ZSC_B200_LIB)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from zsc_b200 import Engine, datagen
import refimpl
uniq = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 8
gs = [0]
S = 262144
n = uniq * rep
E = Engine(raw_bytes=n * S + (1 << 20), comp_bytes=n * 160000 + (1 << 20), deflate_batch_max=uniq * S + (1 << 20), max_streams=n, max_chunks=uniq + 16)
x = np.concatenate([datagen.mixed(uniq // 2 * S, seed=1), datagen.telemetry_buffers(uniq - uniq // 2, S, seed=1000)])
R = refimpl.ref() if refimpl.have_ref() else None
t = time.time()
comps = []
cache = f"/tmp/zsc_prof_inflate_{uniq}.npz"
if os.path.exists(cache):
    z = np.load(cache)
    comps = [z[f"c{i}"] for i in range(uniq)]
elif R is not None and uniq <= 1024:
    from concurrent.futures import ThreadPoolExecutor
    def one(i):
        return R.compress(x[i * S:(i + 1) * S], S, (1, 6, 9)[i % 3])[1]
    with ThreadPoolExecutor(os.cpu_count() or 4) as ex:      # ctypes releases the GIL inside the reference
        comps = list(ex.map(one, range(uniq)))
    np.savez(cache, **{f"c{i}": c for i, c in enumerate(comps)})
else:
    E.upload(0, 0, x)
    st = Engine.make_streams([i * S for i in range(uniq)], [S] * uniq, [i * 160000 for i in range(uniq)], [160000] * uniq)
    res = E.deflate(st, S, 6)
    for i in range(uniq):
        comps.append(E.download(1, i * 160000, res[i].produced))
print("compressed", uniq, "streams in %.1fs" % (time.time() - t), "avg", sum(len(c) for c in comps) / uniq, flush=True)
one_rep = np.zeros(uniq * 160000, np.uint8)
offs1, off = [], 0
for i in range(uniq):
    c = comps[i]
    one_rep[off:off + len(c)] = c
    offs1.append((off, len(c)))
    off += (len(c) + 15) & ~15
rep_bytes = off
offs = []
for r_ in range(rep):
    E.upload(1, r_ * rep_bytes, one_rep[:rep_bytes])
    offs += [(r_ * rep_bytes + o, l) for o, l in offs1]
st = Engine.make_streams([i * S for i in range(n)], [S] * n, [o[0] for o in offs], [o[1] for o in offs])
for g in gs:
    E.inflate_enqueue(st, 1)
    res = E.fetch(n)
    bad = sum(1 for r in res if r.ret != 0 or r.produced != S)
    back = E.download(0, (rep - 1) * uniq * S, uniq * S)
    ok = bool(np.array_equal(back, x))
    ts = []
    for _ in range(3):
        E.event(0); E.relaunch(); E.event(1); E.sync(); ts.append(E.elapsed_ms(0, 1))
    print("lib", os.environ.get("ZSC_B200_LIB", "default"), "G", g, "streams", n, "bad", bad, "bytes_ok", ok, "ms", [round(t, 2) for t in ts], "GB/s", round(n * S / 1e6 / min(ts), 2), flush=True)
import ctypes as _C
try:
    _L = _C.CDLL(os.environ.get("ZSC_B200_LIB") or os.path.join(ROOT, "zsc_b200", "libzsc_b200.so"))
    _f = _L.zs_inflate_spec_prof
    a = (_C.c_uint64 * 32)(); _f(a, 1); a = list(a)
    launches = 4
    names = ["setup", "phase1", "phase2", "chain+measure", "emit", "flush", "cursor", "old path"]
    rounds = max(a[8], 1)
    print("spec profile per round (cycles of lane 0):", {nm: round(a[i] / rounds) for i, nm in enumerate(names)}, "rounds/stream", round(a[8] / launches / n, 1), "sym/round", round(a[9] / rounds), "bytes/round", round(a[10] / rounds), "kernel cycles/stream", round(a[11] / launches / n))
    sw = max(a[20], 1)
    print("emit per sweep: sweeps/round", round(a[20] / rounds, 1), "level calc", round(a[13] / sw), "own copies", round(a[14] / sw), "long copies + sync", round(a[15] / sw), "levels", round(a[16] / sw, 2), "matches", round(a[17] / sw, 1), "reaching in", round(a[18] / sw, 1), "long", round(a[19] / sw, 2), "fetch + scan", round(a[12] / sw))
    if a[23]:
        print("decode set-up: broadcasts", round(a[7] / rounds), "bitmap", round(a[23] / rounds), "(the rest of 'setup': the lanes' readers)")
    if a[21] or a[22]:
        print("two warps per stream: the decoder waits for a free buffer", round(a[21] / rounds), "cycles per round, the writer for a chain", round(a[22] / rounds),
              "; writer warp: lifetime", round(a[24] / launches / n), "of which inside zp_round_emit", round(a[25] / launches / n), "per stream")
except AttributeError:
    pass
E.close()
