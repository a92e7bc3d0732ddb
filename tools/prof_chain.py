"""Profiling driver for the chain kernel: N telemetry buffers of 256 KiB at one level (default 296 buffers, level 6)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zsc_b200 import Engine, datagen
nbuf = int(sys.argv[1]) if len(sys.argv) > 1 else 296
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
kind = sys.argv[3] if len(sys.argv) > 3 else "telemetry"
S = 262144
n = nbuf * S
E = Engine(raw_bytes=n + (1 << 20), comp_bytes=nbuf * 300000 + (1 << 20), deflate_batch_max=n + (1 << 20), max_streams=nbuf, max_chunks=nbuf + 16)
x = datagen.telemetry_buffers(nbuf, S, seed=1000) if kind == "telemetry" else datagen.mixed(n, seed=1)
E.upload(0, 0, x)
st = Engine.make_streams([i * S for i in range(nbuf)], [S] * nbuf, [i * 300000 for i in range(nbuf)], [300000] * nbuf)
E.deflate_enqueue(st, S, level)
res = E.fetch(nbuf)
csize = sum(r.produced for r in res)
E.event(0); E.relaunch(); E.event(1); E.sync()
ms = E.elapsed_ms(0, 1)
print(kind, "level", level, "buffers", nbuf, "ms", round(ms, 2), "lz_ms", round(E.elapsed_ms(9, 10), 2), "GB/s", round(n / 1e6 / ms, 2), "ratio", round(n / csize, 4), flush=True)
E.close()
