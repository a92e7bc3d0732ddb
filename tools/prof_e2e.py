import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zsc_b200 import Engine, datagen, DeflateParams, Result
n = 1 << 30
E = Engine(raw_bytes=n + (1 << 20), comp_bytes=n + (n >> 3) + (1 << 20), deflate_batch_max=n + (1 << 20), max_streams=8192, max_chunks=8192)
x = datagen.mixed(n, seed=1)
dest = np.empty(n + (n >> 3), np.uint8)
print("register", E.L.zscgpu_host_register(x.ctypes.data, x.nbytes), E.L.zscgpu_host_register(dest.ctypes.data, dest.nbytes))
def t(f, reps=3):
    f(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)
print("upload 1GiB ms", t(lambda: E.upload(0, 0, x)))
print("upload 64MiB ms", t(lambda: E.upload(0, 0, x[:64 << 20])))
print("download 468MB ms", t(lambda: E.L.zscgpu_download(E.h, 1, dest.ctypes.data, 0, 468000000)))
st = Engine.make_streams([0], [n], [0], [n + (n >> 3)])
E.deflate_enqueue(st, 262144, 1); E.fetch(1)
print("deflate 1GiB kernels ms", t(lambda: (E.relaunch(), E.sync())))
st2 = Engine.make_streams([0], [64 << 20], [0], [80 << 20])
def wave():
    E.deflate_enqueue(st2, 262144, 1); E.fetch(1)
print("deflate 64MiB enqueue+fetch ms", t(wave))
p = DeflateParams(262144, 1, 0, 1, 15, 0); r = Result()
print("compress_host ms", t(lambda: E.L.zscgpu_compress_host(E.h, dest.ctypes.data, len(dest), x.ctypes.data, n, C.byref(p), 0, C.byref(r))), r.ret, r.produced)
comp = dest[:r.produced].copy()
back = np.empty(n, np.uint8)
print("register", E.L.zscgpu_host_register(back.ctypes.data, back.nbytes), E.L.zscgpu_host_register(comp.ctypes.data, comp.nbytes))
E.upload(1, 0, comp)
st1 = Engine.make_streams([0], [n], [0], [len(comp)])
rs = E.inflate_sectioned(st1, 1)
print("inflate_sectioned resident ms", t(lambda: E.inflate_sectioned(st1, 1)), rs.ret, rs.produced, rs.consumed == len(comp))
r2 = Result()
print("uncompress_host ms", t(lambda: E.L.zscgpu_uncompress_host(E.h, back.ctypes.data, n, comp.ctypes.data, len(comp), 1, C.byref(r2))), r2.ret, r2.produced, bool(np.array_equal(back, x)))
E.close()
