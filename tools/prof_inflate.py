"""Inflate throughput probe: N reference-compressed 256 KiB streams (levels 1/6/9 in thirds), replicated."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from zsc_b200 import Engine, datagen
import refimpl
uniq = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 8
S = 262144
n = uniq * rep
E = Engine(raw_bytes=n * S + (1 << 20), comp_bytes=n * 160000 + (1 << 20), deflate_batch_max=uniq * S + (1 << 20), max_streams=n, max_chunks=uniq + 16)
x = np.concatenate([datagen.mixed(uniq // 2 * S, seed=1), datagen.telemetry_buffers(uniq - uniq // 2, S, seed=1000)])
R = refimpl.ref() if refimpl.have_ref() else None
t = time.time()
comps = []
if R is not None and uniq <= 1024:
    import bench
    import ctypes as C
    L = C.CDLL(refimpl.REF_PATH, mode=C.RTLD_LOCAL)
    for i in range(uniq):
        r, c = R.compress(x[i * S:(i + 1) * S], S, (1, 6, 9)[i % 3])
        comps.append(c)
else:
    E.upload(0, 0, x)
    st = Engine.make_streams([i * S for i in range(uniq)], [S] * uniq, [i * 160000 for i in range(uniq)], [160000] * uniq)
    res = E.deflate(st, S, 6)
    for i in range(uniq):
        comps.append(E.download(1, i * 160000, res[i].produced))
print("compressed", uniq, "streams in %.1fs" % (time.time() - t), "avg", sum(len(c) for c in comps) / uniq)
offs, off = [], 0
buf = np.zeros(n * 160000, np.uint8)
for r_ in range(rep):
    for i in range(uniq):
        c = comps[i]
        buf[off:off + len(c)] = c
        offs.append((off, len(c)))
        off += (len(c) + 15) & ~15
E.upload(1, 0, buf[:off])
st = Engine.make_streams([i * S for i in range(n)], [S] * n, [o[0] for o in offs], [o[1] for o in offs])
E.inflate_enqueue(st, 1)
res = E.fetch(n)
bad = sum(1 for r in res if r.ret != 0 or r.produced != S)
back = E.download(0, 0, uniq * S)
ok = bool(np.array_equal(back, x))
ts = []
for _ in range(3):
    E.event(0); E.relaunch(); E.event(1); E.sync(); ts.append(E.elapsed_ms(0, 1))
print("streams", n, "bad", bad, "bytes_ok", ok, "ms", [round(t, 2) for t in ts], "GB/s", round(n * S / 1e6 / min(ts), 2))
E.close()
