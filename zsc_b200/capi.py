"""ctypes mirror of include/zscgpu.h and include/zsc/zsc_pub.h (same names, same argument order).

Loading fails loudly when the library has not been built (``python -c 'import __graft_entry__ as g;
g.build()'`` or ``make``); engine creation fails loudly when there is no B200 — by design there is
no fallback path.
"""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.environ.get("ZSC_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libzsc_b200.so")   # the override is a tuning aid (tools/build_variant.sh)

Z_OK, Z_STREAM_END, Z_NEED_DICT = 0, 1, 2
Z_ERRNO, Z_STREAM_ERROR, Z_DATA_ERROR, Z_MEM_ERROR, Z_BUF_ERROR, Z_VERSION_ERROR = -1, -2, -3, -4, -5, -6
Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED = 0, 1, 2, 3, 4

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)


class EngineConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("raw_bytes", C.c_uint64), ("comp_bytes", C.c_uint64),
                ("deflate_batch_max", C.c_uint64), ("max_streams", C.c_uint32), ("max_chunks", C.c_uint32)]


class Stream(C.Structure):
    _fields_ = [("raw_off", C.c_uint64), ("raw_len", C.c_uint32), ("comp_len", C.c_uint32), ("comp_off", C.c_uint64)]


class Result(C.Structure):
    _fields_ = [("ret", C.c_int32), ("produced", C.c_uint32), ("consumed", C.c_uint32), ("check", C.c_uint32)]


class DeflateParams(C.Structure):
    _fields_ = [("max_block_len", C.c_uint32), ("level", C.c_int32), ("strategy", C.c_int32),
                ("wrap", C.c_int32), ("window_bits", C.c_int32), ("part", C.c_int32), ("hist_len", C.c_uint32)]


class GzHeader(C.Structure):
    _fields_ = [("text", C.c_int32), ("time", C.c_uint32), ("xflags", C.c_int32), ("os", C.c_int32),
                ("extra", u8p), ("extra_len", C.c_uint32), ("extra_max", C.c_uint32),
                ("name", u8p), ("name_max", C.c_uint32), ("comment", u8p), ("comm_max", C.c_uint32),
                ("hcrc", C.c_int32), ("done", C.c_int32)]


# every symbol include/zscgpu.h and include/zsc/zsc_pub.h declare; tests check they are all exported
ZSCGPU_SYMBOLS = [
    "zscgpu_default_config", "zscgpu_init", "zscgpu_destroy", "zscgpu_last_error", "zscgpu_build_info",
    "zscgpu_global_init", "zscgpu_global", "zscgpu_global_shutdown", "zscgpu_raw_ptr", "zscgpu_comp_ptr",
    "zscgpu_raw_capacity", "zscgpu_comp_capacity", "zscgpu_cuda_stream", "zscgpu_upload", "zscgpu_download",
    "zscgpu_upload_async", "zscgpu_download_async", "zscgpu_sync", "zscgpu_copy_within",
    "zscgpu_host_register", "zscgpu_host_unregister", "zscgpu_deflate_batch", "zscgpu_inflate_batch",
    "zscgpu_deflate_enqueue", "zscgpu_inflate_enqueue", "zscgpu_inflate_sectioned", "zscgpu_fetch_results", "zscgpu_relaunch",
    "zscgpu_last_launch_count", "zscgpu_launch_total", "zscgpu_guess_section_size", "zscgpu_guess_section_size10", "zscgpu_compress_host", "zscgpu_uncompress_host", "zscgpu_checksum_host",
    "zscgpu_adler32", "zscgpu_crc32", "zscgpu_adler32_enqueue", "zscgpu_crc32_enqueue",
    "zscgpu_event_record", "zscgpu_event_elapsed_ms", "zscgpu_debug_fetch_symbols",
    "zscgpu_adler32_combine", "zscgpu_crc32_combine",
    "zscgpu_inflate_stream_open", "zscgpu_inflate_stream_close", "zscgpu_inflate_stream_reset", "zscgpu_inflate_stream_step",
    "zscgpu_inflate_stream_set_dict",
]
ZSC_SYMBOLS = [
    "zsc_compress_get_min_work_buf_size", "zsc_compress_get_min_work_buf_size2",
    "zsc_compress_get_max_output_size", "zsc_compress_get_max_output_size_gzip",
    "zsc_compress_get_max_output_size2", "zsc_compress_get_max_output_size_gzip2",
    "zsc_compress", "zsc_compress_gzip", "zsc_compress2", "zsc_compress_gzip2",
    "zsc_uncompress_get_min_work_buf_size", "zsc_uncompress_get_min_work_buf_size2",
    "zsc_uncompress", "zsc_uncompress_gzip", "zsc_uncompress2", "zsc_uncompress_gzip2",
    "adler32", "adler32_z", "crc32", "crc32_z", "zError", "zlibVersion",
]
# the z_stream API (include/zsc/zlib.h)
ZSTREAM_SYMBOLS = [
    "deflateInit_", "deflateInit2_", "deflate", "deflateEnd", "deflateReset", "deflateSetDictionary", "deflateBoundNoStream",
    "deflateWorkSize", "deflateWorkSize2", "inflateInit_", "inflateInit2_", "inflate", "inflateEnd", "inflateReset",
    "inflateReset2", "inflateSetDictionary", "inflateSync", "inflateWorkSize", "inflateWorkSize2",
]


class ZStream(C.Structure):
    """z_stream of include/zsc/zlib_types_pub.h (same layout as the reference's)"""
    _fields_ = [("next_in", u8p), ("avail_in", C.c_uint32), ("total_in", C.c_uint32),
                ("next_out", u8p), ("avail_out", C.c_uint32), ("total_out", C.c_uint32),
                ("next_work", u8p), ("avail_work", C.c_uint32),
                ("msg", C.c_char_p), ("state", C.c_void_p), ("data_type", C.c_int32), ("adler", C.c_uint32), ("reserved", C.c_uint32)]


def declare_zstream(L):
    """argtypes of the z_stream API on a loaded library (ours or the reference's)"""
    I, U, zp = C.c_int32, C.c_uint32, C.POINTER(ZStream)
    L.deflateInit2_.argtypes = [zp, I, I, I, I, I, C.c_char_p, I]
    L.deflate.argtypes = [zp, I]
    L.deflateEnd.argtypes = [zp]
    L.deflateReset.argtypes = [zp]
    L.deflateSetDictionary.argtypes = [zp, u8p, U]
    L.deflateWorkSize2.argtypes = [I, I, u32p]
    L.inflateInit2_.argtypes = [zp, I, C.c_char_p, I]
    L.inflate.argtypes = [zp, I]
    L.inflateEnd.argtypes = [zp]
    L.inflateReset.argtypes = [zp]
    L.inflateSetDictionary.argtypes = [zp, u8p, U]
    L.inflateSync.argtypes = [zp]
    L.inflateWorkSize2.argtypes = [I, u32p]
    for n in ("deflateInit2_", "deflate", "deflateEnd", "deflateReset", "deflateSetDictionary", "deflateWorkSize2",
              "inflateInit2_", "inflate", "inflateEnd", "inflateReset", "inflateSetDictionary", "inflateSync", "inflateWorkSize2"):
        getattr(L, n).restype = I
    return L



class _Tolerant:
    """attribute proxy that ignores symbols a checker library does not export"""

    def __init__(self, L):
        object.__setattr__(self, "_L", L)

    def __getattr__(self, name):
        try:
            return getattr(self._L, name)
        except AttributeError:
            return _Missing()


class _Missing:
    def __setattr__(self, k, v):
        pass


def _declare_zsc(L, strict=True):
    """argtypes/restype of the zsc_pub.h surface on a loaded library (ours, the reference's, the oracle's)."""
    I, U = C.c_int32, C.c_uint32
    gz = C.POINTER(GzHeader)
    real = L
    if not strict:
        L = _Tolerant(L)
    L.zsc_compress_get_min_work_buf_size.argtypes = [u32p]
    L.zsc_compress_get_min_work_buf_size2.argtypes = [I, I, u32p]
    L.zsc_compress_get_max_output_size.argtypes = [U, U, I, u32p]
    L.zsc_compress_get_max_output_size_gzip.argtypes = [U, U, I, gz, u32p]
    L.zsc_compress_get_max_output_size2.argtypes = [U, U, I, I, I, u32p]
    L.zsc_compress_get_max_output_size_gzip2.argtypes = [U, U, I, I, I, gz, u32p]
    L.zsc_compress.argtypes = [u8p, u32p, u8p, U, U, u8p, U, I]
    L.zsc_compress_gzip.argtypes = [u8p, u32p, u8p, U, U, u8p, U, I, gz]
    L.zsc_compress2.argtypes = [u8p, u32p, u8p, U, U, u8p, U, I, I, I, I]
    L.zsc_compress_gzip2.argtypes = [u8p, u32p, u8p, U, U, u8p, U, I, I, I, I, gz]
    L.zsc_uncompress_get_min_work_buf_size.argtypes = [u32p]
    L.zsc_uncompress_get_min_work_buf_size2.argtypes = [I, u32p]
    L.zsc_uncompress.argtypes = [u8p, u32p, u8p, u32p, u8p, U]
    L.zsc_uncompress_gzip.argtypes = [u8p, u32p, u8p, u32p, u8p, U, gz]
    L.zsc_uncompress2.argtypes = [u8p, u32p, u8p, u32p, u8p, U, I]
    L.zsc_uncompress_gzip2.argtypes = [u8p, u32p, u8p, u32p, u8p, U, I, gz]
    for n in ZSC_SYMBOLS[:16]:
        getattr(L, n).restype = I
    L.adler32.argtypes = [U, u8p, U]; L.adler32.restype = U
    L.adler32_z.argtypes = [U, u8p, C.c_size_t]; L.adler32_z.restype = U
    L.crc32.argtypes = [U, u8p, U]; L.crc32.restype = U
    L.crc32_z.argtypes = [U, u8p, C.c_size_t]; L.crc32_z.restype = U
    L.zError.argtypes = [I]; L.zError.restype = C.c_char_p
    L.zlibVersion.argtypes = []; L.zlibVersion.restype = C.c_char_p
    return real


_lib = None


def lib():
    """The loaded product library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make` or __graft_entry__.build(); "
                               "zsc_b200 has no fallback implementation")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        _declare_zsc(L)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
        L.zscgpu_default_config.argtypes = [C.POINTER(EngineConfig)]
        L.zscgpu_init.argtypes = [C.POINTER(EngineConfig), C.POINTER(vp)]
        L.zscgpu_destroy.argtypes = [vp]; L.zscgpu_destroy.restype = None
        L.zscgpu_last_error.argtypes = [vp]; L.zscgpu_last_error.restype = C.c_char_p
        L.zscgpu_build_info.restype = C.c_char_p
        L.zscgpu_global_init.argtypes = [C.POINTER(EngineConfig)]
        L.zscgpu_global.restype = vp
        L.zscgpu_raw_ptr.argtypes = [vp]; L.zscgpu_raw_ptr.restype = vp
        L.zscgpu_comp_ptr.argtypes = [vp]; L.zscgpu_comp_ptr.restype = vp
        L.zscgpu_raw_capacity.argtypes = [vp]; L.zscgpu_raw_capacity.restype = u64
        L.zscgpu_comp_capacity.argtypes = [vp]; L.zscgpu_comp_capacity.restype = u64
        L.zscgpu_cuda_stream.argtypes = [vp]; L.zscgpu_cuda_stream.restype = vp
        for n in ("zscgpu_upload", "zscgpu_upload_async"):
            getattr(L, n).argtypes = [vp, i32, u64, vp, u64]
        for n in ("zscgpu_download", "zscgpu_download_async"):
            getattr(L, n).argtypes = [vp, i32, vp, u64, u64]
        L.zscgpu_sync.argtypes = [vp]
        L.zscgpu_copy_within.argtypes = [vp, i32, u64, u64, u64]
        L.zscgpu_host_register.argtypes = [vp, u64]
        L.zscgpu_host_unregister.argtypes = [vp]
        sp, rp, pp = C.POINTER(Stream), C.POINTER(Result), C.POINTER(DeflateParams)
        L.zscgpu_deflate_batch.argtypes = [vp, sp, u32, pp, rp]
        L.zscgpu_inflate_batch.argtypes = [vp, sp, u32, i32, rp]
        L.zscgpu_deflate_enqueue.argtypes = [vp, sp, u32, pp]
        L.zscgpu_inflate_enqueue.argtypes = [vp, sp, u32, i32]
        L.zscgpu_inflate_sectioned.argtypes = [vp, sp, i32, rp]
        L.zscgpu_launch_total.argtypes = [vp]
        L.zscgpu_launch_total.restype = C.c_ulonglong
        L.zscgpu_fetch_results.argtypes = [vp, u32, rp]
        L.zscgpu_relaunch.argtypes = [vp]
        L.zscgpu_last_launch_count.argtypes = [vp]; L.zscgpu_last_launch_count.restype = u32
        L.zscgpu_compress_host.argtypes = [vp, vp, u32, vp, u32, pp, u32, rp]
        L.zscgpu_uncompress_host.argtypes = [vp, vp, u32, vp, u32, i32, rp]
        L.zscgpu_checksum_host.argtypes = [vp, i32, u32, vp, u64, u32p]
        L.zscgpu_adler32.argtypes = [vp, u64, u64, u32, u32p]
        L.zscgpu_crc32.argtypes = [vp, u64, u64, u32, u32p]
        L.zscgpu_adler32_enqueue.argtypes = [vp, u64, u64]
        L.zscgpu_crc32_enqueue.argtypes = [vp, u64, u64]
        L.zscgpu_event_record.argtypes = [vp, i32]
        L.zscgpu_event_elapsed_ms.argtypes = [vp, i32, i32, C.POINTER(C.c_float)]
        L.zscgpu_debug_fetch_symbols.argtypes = [vp, u32, u32p, u32, u32p]
        L.zscgpu_adler32_combine.argtypes = [u32, u32, u64]; L.zscgpu_adler32_combine.restype = u32
        L.zscgpu_crc32_combine.argtypes = [u32, u32, u64]; L.zscgpu_crc32_combine.restype = u32
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(u8p)


class Zsc:
    """The zsc_pub.h calls on numpy buffers; works for our library and for the reference's (oracle)."""

    def __init__(self, L):
        self.L = L

    def compress_work_size(self, window_bits=15, mem_level=8):
        v = C.c_uint32(0)
        r = self.L.zsc_compress_get_min_work_buf_size2(window_bits, mem_level, C.byref(v))
        return r, v.value

    def uncompress_work_size(self, window_bits=15):
        v = C.c_uint32(0)
        r = self.L.zsc_uncompress_get_min_work_buf_size2(window_bits, C.byref(v))
        return r, v.value

    def max_output_size(self, source_len, max_block_len, level, window_bits=15, mem_level=8, gz=None):
        v = C.c_uint32(0)
        r = self.L.zsc_compress_get_max_output_size_gzip2(source_len, max_block_len, level, window_bits, mem_level,
                                                          C.byref(gz) if gz is not None else None, C.byref(v))
        return r, v.value

    def compress(self, src, max_block_len, level, window_bits=15, mem_level=8, strategy=0, dest_cap=None,
                 work_len=None, gz=None):
        """-> (ret, bytes) through zsc_compress_gzip2."""
        src = np.ascontiguousarray(src, dtype=np.uint8)
        if dest_cap is None:
            r, dest_cap = self.max_output_size(len(src), max_block_len, level, window_bits, mem_level, gz)
            if r != 0:
                dest_cap = len(src) + len(src) // 4 + 1024
        if work_len is None:
            r, work_len = self.compress_work_size(window_bits, mem_level)
            if r != 0:
                work_len = 400000
        dest = np.empty(max(dest_cap, 1), dtype=np.uint8)
        dest[:] = 0xA5
        work = np.empty(max(work_len, 1), dtype=np.uint8)
        dl = C.c_uint32(dest_cap)
        s = src if len(src) else np.zeros(1, dtype=np.uint8)
        r = self.L.zsc_compress_gzip2(_ptr(dest), C.byref(dl), _ptr(s), len(src), max_block_len, _ptr(work), work_len,
                                      level, window_bits, mem_level, strategy, C.byref(gz) if gz is not None else None)
        return r, dest[:dl.value].copy()

    def uncompress(self, comp, dest_cap, window_bits=15, work_len=None, gz=None):
        """-> (ret, bytes, consumed) through zsc_uncompress_gzip2."""
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        if work_len is None:
            r, work_len = self.uncompress_work_size(window_bits)
            if r != 0:
                work_len = 50000
        dest = np.empty(max(dest_cap, 1), dtype=np.uint8)
        dest[:] = 0x5A
        work = np.empty(max(work_len, 1), dtype=np.uint8)
        dl, sl = C.c_uint32(dest_cap), C.c_uint32(len(comp))
        c = comp if len(comp) else np.zeros(1, dtype=np.uint8)
        r = self.L.zsc_uncompress_gzip2(_ptr(dest), C.byref(dl), _ptr(c), C.byref(sl), _ptr(work), work_len, window_bits,
                                        C.byref(gz) if gz is not None else None)
        return r, dest[:dl.value].copy(), sl.value

    def adler32(self, data, init=1):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        return self.L.adler32_z(init, _ptr(data) if len(data) else _ptr(np.zeros(1, np.uint8)), len(data))

    def crc32(self, data, init=0):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        return self.L.crc32_z(init, _ptr(data) if len(data) else _ptr(np.zeros(1, np.uint8)), len(data))


def zsc():
    return Zsc(lib())


class Engine:
    """One zscgpu_engine: arenas fixed at construction, batched device-resident calls."""

    def __init__(self, raw_bytes=None, comp_bytes=None, deflate_batch_max=None, max_streams=None, max_chunks=None, device=0):
        self.L = lib()
        cfg = EngineConfig()
        self.L.zscgpu_default_config(C.byref(cfg))
        cfg.device = device
        if raw_bytes is not None:
            cfg.raw_bytes = raw_bytes
        if comp_bytes is not None:
            cfg.comp_bytes = comp_bytes
        if deflate_batch_max is not None:
            cfg.deflate_batch_max = deflate_batch_max
        elif raw_bytes is not None:
            cfg.deflate_batch_max = raw_bytes
        if max_streams is not None:
            cfg.max_streams = max_streams
        if max_chunks is not None:
            cfg.max_chunks = max_chunks
        self.cfg = cfg
        h = C.c_void_p()
        r = self.L.zscgpu_init(C.byref(cfg), C.byref(h))
        if r != 0:
            raise RuntimeError(f"zscgpu_init failed ({r}): {self.L.zscgpu_last_error(None).decode()}")
        self.h = h

    def close(self):
        if self.h:
            self.L.zscgpu_destroy(self.h)
            self.h = None

    def _ck(self, r):
        if r != 0:
            raise RuntimeError(f"zscgpu call failed ({r}): {self.L.zscgpu_last_error(self.h).decode()}")

    def upload(self, which, off, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        self._ck(self.L.zscgpu_upload(self.h, which, off, arr.ctypes.data, arr.nbytes))

    def download(self, which, off, n):
        out = np.empty(n, dtype=np.uint8)
        self._ck(self.L.zscgpu_download(self.h, which, out.ctypes.data, off, n))
        return out

    @staticmethod
    def make_streams(raw_offs, raw_lens, comp_offs, comp_lens):
        n = len(raw_offs)
        arr = (Stream * n)()
        for i in range(n):
            arr[i].raw_off, arr[i].raw_len = int(raw_offs[i]), int(raw_lens[i])
            arr[i].comp_off, arr[i].comp_len = int(comp_offs[i]), int(comp_lens[i])
        return arr

    def deflate(self, streams, max_block_len, level, strategy=0, wrap=1, window_bits=15, part=0):
        p = DeflateParams(max_block_len, level, strategy, wrap, window_bits, part)
        res = (Result * len(streams))()
        self._ck(self.L.zscgpu_deflate_batch(self.h, streams, len(streams), C.byref(p), res))
        return res

    def inflate(self, streams, wrap=1):
        res = (Result * len(streams))()
        self._ck(self.L.zscgpu_inflate_batch(self.h, streams, len(streams), wrap, res))
        return res

    def inflate_sectioned(self, stream, wrap=1):
        """one large stream (a 1-element stream array), its flush-delimited sections decoded in parallel"""
        res = Result()
        self._ck(self.L.zscgpu_inflate_sectioned(self.h, stream, wrap, C.byref(res)))
        return res

    def deflate_enqueue(self, streams, max_block_len, level, strategy=0, wrap=1, window_bits=15):
        p = DeflateParams(max_block_len, level, strategy, wrap, window_bits, 0)
        self._ck(self.L.zscgpu_deflate_enqueue(self.h, streams, len(streams), C.byref(p)))

    def inflate_enqueue(self, streams, wrap=1):
        self._ck(self.L.zscgpu_inflate_enqueue(self.h, streams, len(streams), wrap))

    def fetch(self, n):
        res = (Result * n)()
        self._ck(self.L.zscgpu_fetch_results(self.h, n, res))
        return res

    def relaunch(self):
        self._ck(self.L.zscgpu_relaunch(self.h))

    def sync(self):
        self._ck(self.L.zscgpu_sync(self.h))

    def event(self, slot):
        self._ck(self.L.zscgpu_event_record(self.h, slot))

    def elapsed_ms(self, a, b):
        ms = C.c_float(0)
        self._ck(self.L.zscgpu_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def adler32(self, off, n, init=1):
        v = C.c_uint32(0)
        self._ck(self.L.zscgpu_adler32(self.h, off, n, init, C.byref(v)))
        return v.value

    def crc32(self, off, n, init=0):
        v = C.c_uint32(0)
        self._ck(self.L.zscgpu_crc32(self.h, off, n, init, C.byref(v)))
        return v.value

    def symbols(self, chunk, cap):
        out = np.empty(cap, dtype=np.uint32)
        n = C.c_uint32(0)
        self._ck(self.L.zscgpu_debug_fetch_symbols(self.h, chunk, out.ctypes.data_as(u32p), cap, C.byref(n)))
        return out[:min(n.value, cap)], n.value
