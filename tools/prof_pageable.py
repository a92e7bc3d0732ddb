"""zsc_compress (zsc_pub.h) on ordinary pageable caller buffers: 1 GiB, level 1, 256 KiB sections, on the process-wide engine.
Compare builds with ZSC_B200_LIB (tools/build_variant.sh)."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zsc_b200 import datagen
from zsc_b200.capi import zsc, lib, EngineConfig
n = 1 << 30
x = datagen.mixed(n, seed=1)
cfg = EngineConfig(); lib().zscgpu_default_config(C.byref(cfg)); cfg.device = 0
lib().zscgpu_global_init(C.byref(cfg))
Z = zsc()
cap = Z.max_output_size(n, 262144, 1)[1]
wl = Z.compress_work_size()[1]
work = np.empty(wl, np.uint8)
dst = np.empty(cap, np.uint8)
dl = C.c_uint32(cap)
u8 = C.POINTER(C.c_uint8)
args = (dst.ctypes.data_as(u8), C.byref(dl), x.ctypes.data_as(u8), n, 262144, work.ctypes.data_as(u8), wl, 1)
ts = []
for i in range(5):
    dl.value = cap
    t0 = time.perf_counter(); rc = Z.L.zsc_compress(*args); ts.append((time.perf_counter() - t0) * 1e3)
    assert rc == 0
print("cpus", os.cpu_count(), "zsc_compress pageable ms", [round(t, 1) for t in ts], "GB/s", round(n / 1e6 / min(ts[1:]), 2), "produced", dl.value)
lib().zscgpu_global_shutdown()
