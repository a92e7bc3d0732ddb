"""zscgpu_compress_host end to end (1 GiB, level 1, 256 KiB sections), sections per wave swept via ZSC_B200_WAVE_SECS (a tuning build: tools/build_variant.sh NAME -DZSC_TUNING)."""
import os, sys, time, ctypes as C, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zsc_b200 import Engine, datagen, DeflateParams, Result
n = 1 << 30
E = Engine(raw_bytes=n + (1 << 20), comp_bytes=n + (n >> 3) + (1 << 20), deflate_batch_max=n + (1 << 20), max_streams=8192, max_chunks=8192)
x = datagen.mixed(n, seed=1)
dest = np.empty(n + (n >> 3), np.uint8)
E.L.zscgpu_host_register(x.ctypes.data, x.nbytes); E.L.zscgpu_host_register(dest.ctypes.data, dest.nbytes)
p = DeflateParams(262144, 1, 0, 1, 15, 0); r = Result()
ref = None
for rounds in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["296"]):
    os.environ["ZSC_B200_WAVE_SECS"] = rounds
    ts = []
    for i in range(4):
        t0 = time.perf_counter()
        rc = E.L.zscgpu_compress_host(E.h, dest.ctypes.data, len(dest), x.ctypes.data, n, C.byref(p), 0, C.byref(r))
        ts.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0 and r.ret == 0, (rc, r.ret)
    h = zlib.adler32(dest[:r.produced].tobytes())
    if ref is None:
        ref = (r.produced, h)
        ok = zlib.decompress(dest[:r.produced].tobytes()) == x.tobytes()
        print("inflates to the input:", ok, flush=True)
    print("sections per wave", rounds, "ms", [round(t, 2) for t in ts], "GB/s", round(n / 1e6 / min(ts[1:]), 2), "produced", r.produced, "same bytes as first:", (r.produced, h) == ref, flush=True)
E.close()
