/* common.cuh — shared device-side definitions of the B200 DEFLATE engine. */
#ifndef ZSC_COMMON_CUH
#define ZSC_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include "huff_build.h"

/* ---- geometry ---- */
#define ZS_BLOCK_SYMS   8192u        /* LZ77 symbols per deflate block */
#define ZS_CHUNK_MAX    262144u      /* largest chunk one LZ CTA owns; larger sections are sub-chunked
                                        with the preceding 32 KiB as dictionary */
#define ZS_WINDOW       32768u
#define ZS_MAX_MATCH    258u
#define ZS_MIN_MATCH    3u
#define ZS_TOO_FAR      4096u        /* length-3 matches farther than this cost more than literals
                                        (same constant as reference src/deflate.c:130) */

/* One unit of LZ77 work: a chunk of a section.  Host-built, read by every deflate kernel. */
struct ZsChunk {
    uint64_t raw_off;     /* first input byte (raw arena offset) */
    uint64_t sym_off;     /* first symbol slot (multiple of 4) */
    uint32_t len;         /* input bytes */
    uint32_t dict_len;    /* bytes before raw_off that belong to the same section (<= 32768) */
    uint32_t blk_base;    /* first block slot */
    uint32_t blk_cap;     /* block slots reserved = max(1, ceil(len / ZS_BLOCK_SYMS)) */
    uint32_t stream;      /* owning stream */
    uint32_t flags;       /* ZC_* */
};
enum { ZC_FIRST_OF_STREAM = 1, ZC_LAST_OF_SECTION = 2, ZC_LAST_OF_STREAM = 4 };

/* Per-stream control block (device). */
struct ZsStream {
    uint64_t raw_off;
    uint64_t comp_off;
    uint32_t raw_len;
    uint32_t comp_cap;
    uint32_t blk_first;   /* first block slot of the stream */
    uint32_t blk_count;   /* number of block slots of the stream */
    uint32_t chunk_first;
    uint32_t chunk_count;
};

/* LZ77 search parameters for one batch (derived from level/strategy in engine.cu). */
struct ZsLzParams {
    int32_t mode;        /* 0 hash search, 1 run-length (distance 1 only), 2 literals only */
    int32_t chain;       /* candidates tried beyond the first (0: the single-candidate kernel) */
    int32_t nice;        /* stop the search at this match length */
    int32_t lazy;        /* 1: defer a match when the next position has a longer one */
    int32_t min_len;     /* shortest match kept (3; 6 for Z_FILTERED) */
    int32_t force_type;  /* -1, ZH_STATIC (Z_FIXED) or ZH_STORED (level 0) */
    int32_t wrap;        /* 0 raw, 1 zlib, 2 gzip body */
    int32_t zhdr;        /* the two zlib header bytes, little-endian packed */
    int32_t max_dist;    /* 1 << window_bits */
    int32_t good;        /* chain kernel: behind a match at least this long a position searches with a quarter of the budget */
    int32_t max_lazy;    /* chain kernel: behind a taken match at least this long the next position is not searched deeper */
};

/* Adler-32 accumulators: sum of bytes and position-weighted sum, both already reduced mod 65521. */
struct ZsAdlerAcc { unsigned long long s1, s2; };

#define ZS_ADLER_BASE 65521u
#define ZS_STR2(x) #x
#define ZS_STR(x) ZS_STR2(x)

__device__ __forceinline__ uint32_t zs_lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ uint32_t zs_lanemask_gt() { uint32_t m; asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m)); return m; }

#define ZS_CUDA_CHECK(call)                                                          \
    do { cudaError_t _e = (call); if (_e != cudaSuccess) return zs_fail(e, _e, #call, __LINE__); } while (0)

#endif
