"""zscgpu_uncompress_host end to end on the 1 GiB level-1 stream (4096 sections), pinned buffers; wave layouts swept via
ZSC_B200_UNC_WAVES / ZSC_B200_UNC_STREAMS and traced with ZSC_B200_TRACE (tuning builds, tools/build_variant.sh)."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zsc_b200 import Engine, datagen, DeflateParams, Result
n = 1 << 30
E = Engine(raw_bytes=n + (1 << 20), comp_bytes=n + (n >> 3) + (1 << 20), deflate_batch_max=n + (1 << 20), max_streams=8192, max_chunks=8192)
x = datagen.mixed(n, seed=1)
comp = np.empty(n + (n >> 3), np.uint8)
back = np.empty(n, np.uint8)
for a in (x, comp, back):
    E.L.zscgpu_host_register(a.ctypes.data, a.nbytes)
p = DeflateParams(262144, 1, 0, 1, 15, 0); r = Result()
assert E.L.zscgpu_compress_host(E.h, comp.ctypes.data, len(comp), x.ctypes.data, n, C.byref(p), 0, C.byref(r)) == 0 and r.ret == 0
clen = r.produced
for cfg in (sys.argv[1:] or ["default"]):
    for kv in cfg.split(","):
        if "=" in kv:
            k, v = kv.split("="); os.environ[k] = v
    ts = []
    for i in range(4):
        back[:4096] = 0
        t0 = time.perf_counter()
        rc = E.L.zscgpu_uncompress_host(E.h, back.ctypes.data, n, comp.ctypes.data, clen, 1, C.byref(r))
        ts.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0 and r.ret == 0 and r.produced == n, (rc, r.ret, r.produced)
    print(cfg, "ms", [round(t, 2) for t in ts], "GB/s", round(n / 1e6 / min(ts[1:]), 2), "bytes equal:", bool(np.array_equal(back, x)), flush=True)
    for kv in cfg.split(","):
        if "=" in kv:
            os.environ.pop(kv.split("=")[0], None)
E.close()
