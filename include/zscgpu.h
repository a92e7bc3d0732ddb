/* zscgpu.h — the C-ABI of the B200 DEFLATE engine.
 *
 * This is the thin layer the host C code (zsc_b200/csrc/host/zsc_api.c, the zsc_pub.h surface) calls;
 * it is also what a foreign-language binding (ctypes / cgo / JNI) would bind for batched,
 * device-resident work.  Plain pointers and sizes only; no CUDA or torch types.
 *
 * Memory model ("no dynamic allocation after init", reference README.md:37-44 and the
 * work-buffer carve-up at reference src/deflate.c:332-375): zscgpu_init() allocates every device
 * arena and every pinned control block once; no later call allocates or frees anything.
 *
 *   raw  arena : uncompressed bytes (deflate input / inflate output)
 *   comp arena : compressed bytes   (deflate output / inflate input)
 *   sym  arena : LZ77 symbol scratch, 4 B per input byte of the largest deflate batch
 *
 * What each entry point replaces in the reference:
 *   zscgpu_deflate_batch   the section loop of zsc_compress_gzip2 (src/zsc_compress.c:121-140) and
 *                          everything below it: deflate() -> deflate_fast / deflate_slow /
 *                          deflate_rle / deflate_huff / deflate_stored (src/deflate.c:1694-2245),
 *                          longest_match (:1400), fill_window (:1532), _tr_flush_block and the
 *                          Huffman stage (src/trees.c:426-993), zlib header/trailer
 *                          (src/deflate.c:1029-1057, :1284-1285)
 *   zscgpu_inflate_batch   the loop of zsc_uncompress_gzip2 (src/zsc_uncompr.c:103-127): inflate()
 *                          (src/inflate.c:704-1404), inflate_table (src/inftrees.c:60),
 *                          inflate_fast (src/inffast.c:76), inflateSync (src/inflate.c:1547)
 *   zscgpu_adler32 / zscgpu_crc32   adler32_z (src/adler32.c:56), crc32_z (src/crc32.c:502)
 *
 * All functions return 0 on success or a negative zscgpu_status; stream-level results use the
 * reference's ZlibReturn values (Z_OK 0, Z_DATA_ERROR -3, Z_BUF_ERROR -5 ...).
 *
 * Threading: the zsc_pub.h entry points and the one-shot host-buffer calls below (zscgpu_compress_host,
 * zscgpu_uncompress_host, zscgpu_checksum_host) may be called from any number of threads — they serialise on the
 * engine, as every call uses the whole arenas.  The batched, device-resident calls (enqueue / relaunch / fetch,
 * upload / download) belong to ONE thread per engine at a time: the descriptor arrays and result slots of an
 * engine hold one batch.  Concurrency beyond that is one engine per caller (zscgpu_init is cheap to repeat with
 * smaller arenas) or, across GPUs, one engine per device.
 */
#ifndef ZSCGPU_H
#define ZSCGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zscgpu_engine zscgpu_engine;

typedef enum {
    ZSCGPU_OK = 0,
    ZSCGPU_ERR_NO_DEVICE = -101,   /* no CUDA device / wrong architecture: there is NO CPU fallback */
    ZSCGPU_ERR_CUDA = -102,        /* a CUDA runtime call failed; see zscgpu_last_error() */
    ZSCGPU_ERR_CAPACITY = -103,    /* batch exceeds the arenas fixed at init */
    ZSCGPU_ERR_ARG = -104
} zscgpu_status;

typedef struct {
    int32_t  device;            /* CUDA device ordinal */
    uint64_t raw_bytes;         /* capacity of the raw arena */
    uint64_t comp_bytes;        /* capacity of the comp arena */
    uint64_t deflate_batch_max; /* largest number of input bytes one zscgpu_deflate_batch may take */
    uint32_t max_streams;       /* largest number of streams per batch */
    uint32_t max_chunks;        /* largest number of chunks (sections and their sub-chunks) per batch */
} zscgpu_config;

/* One stream of a batch.  Offsets are byte offsets into the arenas. */
typedef struct {
    uint64_t raw_off;   /* deflate: input,  inflate: output */
    uint32_t raw_len;   /* deflate: input length, inflate: output capacity */
    uint32_t comp_len;  /* deflate: output capacity, inflate: input length */
    uint64_t comp_off;  /* deflate: output, inflate: input */
} zscgpu_stream;

/* Per-stream results. */
typedef struct {
    int32_t  ret;       /* ZlibReturn of this stream */
    uint32_t produced;  /* bytes written (deflate: compressed, inflate: uncompressed) */
    uint32_t consumed;  /* bytes consumed of the input */
    uint32_t check;     /* adler32 (zlib) or crc32 (gzip) of the uncompressed data */
} zscgpu_result;

/* Parameters of a deflate batch; same meaning and validation as zsc_compress2's arguments. */
typedef struct {
    uint32_t max_block_len;  /* section size; a full-flush marker follows every section but the last */
    int32_t  level;          /* -1, 0..9 */
    int32_t  strategy;       /* ZlibStrategy */
    int32_t  wrap;           /* 0 raw deflate, 1 zlib (2-byte header + adler32), 2 gzip body: raw deflate,
                                header/trailer added by the host layer, crc32 reported in result.check */
    int32_t  window_bits;    /* 9..15 (0 = 15): largest match distance is 1 << window_bits */
    int32_t  part;           /* 0 whole streams.  For one logical stream sharded over several engines
                                (GPUs): bit 0 = not the first part (no stream header), bit 1 = not the last
                                part (the last section ends with a full-flush marker, no final block and no
                                trailer).  Parts concatenate byte-wise; result.check is the part's own
                                adler32, to be folded with zscgpu_adler32_combine. */
    uint32_t hist_len;       /* preset dictionary / history: this many bytes (<= 32768) directly in front of every
                                stream's raw_off belong to the match window of its first section (deflateSetDictionary
                                and the chunk-by-chunk deflate() of the streaming API; 0 otherwise) */
} zscgpu_deflate_params;

void zscgpu_default_config(zscgpu_config *cfg);
int  zscgpu_init(const zscgpu_config *cfg, zscgpu_engine **out);
void zscgpu_destroy(zscgpu_engine *e);
const char *zscgpu_last_error(const zscgpu_engine *e);   /* e may be NULL: last init error */
const char *zscgpu_build_info(void);                     /* e.g. "sm_100a cuda 12.9" */

/* Process-wide default engine used by the zsc_pub.h entry points.  zscgpu_global_init() is the
 * explicit "init" of the no-allocation-after-init model; if the application never calls it, the
 * first zsc_* call initialises an engine with zscgpu_default_config(). */
int  zscgpu_global_init(const zscgpu_config *cfg);
zscgpu_engine *zscgpu_global(void);
void zscgpu_global_shutdown(void);

/* Arena access (device pointers, for callers that produce or consume data on the GPU). */
void    *zscgpu_raw_ptr(zscgpu_engine *e);
void    *zscgpu_comp_ptr(zscgpu_engine *e);
uint64_t zscgpu_raw_capacity(const zscgpu_engine *e);
uint64_t zscgpu_comp_capacity(const zscgpu_engine *e);
void    *zscgpu_cuda_stream(zscgpu_engine *e);   /* the cudaStream_t every kernel is launched on */

/* Host <-> arena copies on the engine's stream; *_async returns after enqueueing. `which`: 0 raw, 1 comp. */
int zscgpu_upload(zscgpu_engine *e, int which, uint64_t off, const void *host, uint64_t n);
int zscgpu_download(zscgpu_engine *e, int which, void *host, uint64_t off, uint64_t n);
int zscgpu_upload_async(zscgpu_engine *e, int which, uint64_t off, const void *host, uint64_t n);
int zscgpu_download_async(zscgpu_engine *e, int which, void *host, uint64_t off, uint64_t n);
int zscgpu_sync(zscgpu_engine *e);
/* Device-to-device replicate inside an arena (bench: build config 4's 16x replicated streams). */
int zscgpu_copy_within(zscgpu_engine *e, int which, uint64_t dst_off, uint64_t src_off, uint64_t n);

/* Pinned host memory from the engine's fixed pool is not provided; callers may pin their own. */
int zscgpu_host_register(void *p, uint64_t n);
int zscgpu_host_unregister(void *p);

/* Batched codec calls: inputs and outputs are resident in the arenas.  Blocking: results are valid
 * on return.  n streams; res has n entries. */
int zscgpu_deflate_batch(zscgpu_engine *e, const zscgpu_stream *streams, uint32_t n,
                         const zscgpu_deflate_params *p, zscgpu_result *res);
int zscgpu_inflate_batch(zscgpu_engine *e, const zscgpu_stream *streams, uint32_t n,
                         int32_t wrap, zscgpu_result *res);

/* Split calls used by the benchmark to time with CUDA events: enqueue only / fetch results. */
int zscgpu_deflate_enqueue(zscgpu_engine *e, const zscgpu_stream *streams, uint32_t n,
                           const zscgpu_deflate_params *p);
int zscgpu_inflate_enqueue(zscgpu_engine *e, const zscgpu_stream *streams, uint32_t n, int32_t wrap);
/* One large stream, its sections inflated in parallel: zsc_compress ends every max_block_len section with a
 * full flush (reference src/zsc_compress.c:121-140), after which nothing refers back, so the sections decode
 * like independent streams once a scan for the flush markers and a size-only pass have located them.  Same
 * result as zscgpu_inflate_batch on that one stream (to which anything irregular falls back). */
int zscgpu_inflate_sectioned(zscgpu_engine *e, const zscgpu_stream *stream, int32_t wrap, zscgpu_result *result);
/* The section size zscgpu_inflate_sectioned tries first (one decode pass instead of two) for a stream of `sections`
 * flush-delimited sections whose data fills `total` bytes; 0 when no plausible uniform size exists.  Pure arithmetic. */
uint64_t zscgpu_guess_section_size(uint64_t total, uint32_t sections);
uint64_t zscgpu_guess_section_size10(uint64_t total, uint32_t sections);   /* ... the same with decimal roundness */
int zscgpu_fetch_results(zscgpu_engine *e, uint32_t n, zscgpu_result *res);
/* Re-launch the kernels of the last enqueue without rebuilding descriptors (bench inner loop). */
int zscgpu_relaunch(zscgpu_engine *e);
uint32_t zscgpu_last_launch_count(const zscgpu_engine *e);
/* kernels launched by this engine since zscgpu_init (bench.py reports the difference over its timed regions) */
unsigned long long zscgpu_launch_total(const zscgpu_engine *e);

/* One-shot calls on HOST buffers (what zsc_compress / zsc_uncompress / adler32 / crc32 use): copy in,
 * run a batch of one stream at offset 0 of the arenas, copy out.  comp_skip leaves room at the start
 * of dest for a wrapper the caller writes itself (gzip header). kind: 0 adler32, 1 crc32. */
int zscgpu_compress_host(zscgpu_engine *e, uint8_t *dest, uint32_t dest_cap, const uint8_t *src,
                         uint32_t src_len, const zscgpu_deflate_params *p, uint32_t comp_skip,
                         zscgpu_result *res);
int zscgpu_uncompress_host(zscgpu_engine *e, uint8_t *dest, uint32_t dest_cap, const uint8_t *src,
                           uint32_t src_len, int32_t wrap, zscgpu_result *res);
int zscgpu_checksum_host(zscgpu_engine *e, int kind, uint32_t init, const uint8_t *buf, uint64_t len,
                         uint32_t *out);

/* The z_stream API's inflate (reference include/zsc/zlib.h:303 inflate, :818 inflateSetDictionary, :856 inflateSync): the
 * decoder state of a stream lives in one of a fixed pool of device slots between calls.  A step stages `in_len` new
 * input bytes behind the `in_left` bytes the previous step left unread, decodes until the input runs out, `out_cap`
 * bytes are produced, the stream ends or an error stops it, and copies the produced bytes to `out`.
 * status: 0 more input needed, 1 output full, 2 stream end (trailer read: stored_check / have_check), 3 data error
 * (the stream waits for zscgpu_inflate_stream_sync), 4 preset dictionary needed (stored_check = its adler32),
 * 5 flush point found (sync only).  in_pos = bytes of the staged input (in_left + in_len) that were used up. */
typedef struct {
    uint32_t status, in_pos, produced, adler, stored_check, have_check;
} zscgpu_stream_step;
#define ZSCGPU_STREAM_IN_MAX 65536u
#define ZSCGPU_STREAM_OUT_MAX 32768u
int zscgpu_inflate_stream_open(zscgpu_engine *e, int32_t wrap, int32_t *slot);      /* wrap as for zscgpu_inflate_batch */
int zscgpu_inflate_stream_close(zscgpu_engine *e, int32_t slot);
int zscgpu_inflate_stream_reset(zscgpu_engine *e, int32_t slot, int32_t wrap);
int zscgpu_inflate_stream_step(zscgpu_engine *e, int32_t slot, const uint8_t *in, uint32_t in_len, uint32_t in_left,
                               uint8_t *out, uint32_t out_cap, int32_t sync, zscgpu_stream_step *res);
int zscgpu_inflate_stream_set_dict(zscgpu_engine *e, int32_t slot, const uint8_t *dict, uint32_t len);

/* Checksums over raw-arena bytes [off, off+len): value continues from `init` (adler: 1, crc: 0 to start). */
int zscgpu_adler32(zscgpu_engine *e, uint64_t off, uint64_t len, uint32_t init, uint32_t *out);
int zscgpu_crc32(zscgpu_engine *e, uint64_t off, uint64_t len, uint32_t init, uint32_t *out);
int zscgpu_adler32_enqueue(zscgpu_engine *e, uint64_t off, uint64_t len);
int zscgpu_crc32_enqueue(zscgpu_engine *e, uint64_t off, uint64_t len);

/* Event timing on the engine's stream (so a C or ctypes caller can time kernels on the launching
 * stream without a CUDA binding of its own).  Slots 0..7 are the caller's.  Slots 8..13 are recorded
 * by every deflate launch: 8 start, 9 after adler32, 10 after the LZ77 kernel, 11 after block
 * histogram + code construction, 12 after the offset scan, 13 after bit packing. */
int zscgpu_event_record(zscgpu_engine *e, int slot);
int zscgpu_event_elapsed_ms(zscgpu_engine *e, int slot_start, int slot_stop, float *ms);

/* Debug / test hooks: LZ77 symbols of chunk `chunk` of the last deflate batch. */
int zscgpu_debug_fetch_symbols(zscgpu_engine *e, uint32_t chunk, uint32_t *out, uint32_t cap,
                               uint32_t *nsym);

/* Host-side checksum combination (the reference removed adler32_combine / crc32_combine,
 * src/adler32.c:142-143, src/crc32.c:636-637; these follow from the definitions). */
uint32_t zscgpu_adler32_combine(uint32_t adler1, uint32_t adler2, uint64_t len2);
uint32_t zscgpu_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);

#ifdef __cplusplus
}
#endif
#endif
