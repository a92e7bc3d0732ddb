"""Profiling driver: one deflate pass over 296 chunks (2 per SM) so that ncu replays stay short."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zsc_b200 import Engine, datagen
level = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nchunks = int(sys.argv[2]) if len(sys.argv) > 2 else 296
n = nchunks * 262144
E = Engine(raw_bytes=n + (1 << 20), comp_bytes=n + (n >> 3) + (1 << 20), deflate_batch_max=n + (1 << 20), max_streams=16, max_chunks=4096)
x = datagen.mixed(n, seed=1)
E.upload(0, 0, x)
st = Engine.make_streams([0], [n], [0], [n + (n >> 3)])
E.deflate_enqueue(st, 262144, level)
r = E.fetch(1)[0]
E.event(0); E.relaunch(); E.event(1); E.sync()
print("ret", r.ret, "produced", r.produced, "ms", E.elapsed_ms(0, 1), "lz_ms", E.elapsed_ms(9, 10), "block_ms", E.elapsed_ms(10, 11), "offs_ms", E.elapsed_ms(11, 12), "enc_ms", E.elapsed_ms(12, 13), "GB/s", n / 1e6 / E.elapsed_ms(0, 1))
E.close()
