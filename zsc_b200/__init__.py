"""zsc_b200 — Python binding (ctypes) of the B200 DEFLATE engine.

The product is the shared library ``libzsc_b200.so`` (host C + CUDA for sm_100a) that exports the
zsc_pub.h surface of abcouwer-jpl/zsc and the batched C-ABI of include/zscgpu.h.  This package only
loads it and mirrors both interfaces for tests and benchmarks; there is no Python or CPU codec here.
"""
from .capi import (Engine, EngineConfig, DeflateParams, Stream, Result, lib, zsc,  # noqa: F401
                   Z_OK, Z_STREAM_END, Z_NEED_DICT, Z_STREAM_ERROR, Z_DATA_ERROR, Z_MEM_ERROR,
                   Z_BUF_ERROR, Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED,
                   LIB_PATH)
