"""adler32 / crc32 throughput on a resident buffer (default 4 GiB of random bytes), with a bit-exact check."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zsc_b200 import Engine, datagen
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 4096) << 20
E = Engine(raw_bytes=n + (1 << 20), comp_bytes=1 << 20, deflate_batch_max=1 << 20, max_streams=16, max_chunks=16)
piece = datagen.fill(256 << 20, 5, datagen.RANDOM)
for off in range(0, n, len(piece)):
    E.upload(0, off, piece[:min(len(piece), n - off)])
for off, m in ((0, 1 << 20), (5, (64 << 20) + 333), (0, 256 << 20)):
    got = E.crc32(off, m)
    exp = zlib.crc32(piece[off:off + m].tobytes())
    print("crc32 off", off, "len", m, "ok" if got == exp else f"MISMATCH {got:08x} {exp:08x}", flush=True)
for name, fn in (("adler32", E.L.zscgpu_adler32_enqueue), ("crc32", E.L.zscgpu_crc32_enqueue)):
    for m in (n, 1 << 30, 64 << 20):
        if m > n:
            continue
        fn(E.h, 0, m); E.sync()
        ts = []
        for _ in range(5):
            E.event(0); fn(E.h, 0, m); E.event(1); E.sync(); ts.append(E.elapsed_ms(0, 1))
        print(name, "bytes", m, "ms", round(min(ts), 3), "GB/s", round(m / 1e6 / min(ts), 1), flush=True)
E.close()
