"""configs[2]: N independent 256 KiB telemetry-like buffers at levels 6 and 9 (one stream each), device resident."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zsc_b200 import Engine, datagen
nbuf = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
levels = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [6, 9]
S = 262144
n = nbuf * S
E = Engine(raw_bytes=n + (1 << 20), comp_bytes=nbuf * 300000 + (1 << 20), deflate_batch_max=n + (1 << 20), max_streams=nbuf, max_chunks=nbuf + 16)
x = datagen.telemetry_buffers(nbuf, S, seed=1000)
E.upload(0, 0, x)
st = Engine.make_streams([i * S for i in range(nbuf)], [S] * nbuf, [i * 300000 for i in range(nbuf)], [300000] * nbuf)
for level in levels:
    E.deflate_enqueue(st, S, level)
    res = E.fetch(nbuf)
    assert all(r.ret == 0 for r in res)
    csize = sum(r.produced for r in res)
    E.event(0); E.relaunch(); E.event(1); E.sync()
    ms = E.elapsed_ms(0, 1)
    print("level", level, "buffers", nbuf, "ms", round(ms, 2), "lz_ms", round(E.elapsed_ms(9, 10), 2), "GB/s", round(n / 1e6 / ms, 2), "ratio", round(n / csize, 4), flush=True)
E.close()
