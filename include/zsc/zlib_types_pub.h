/* Public types of the zsc surface, as the B200 engine exports them.
 *
 * ABI-identical to reference include/zsc/zlib_types_pub.h (enum values :161-249, z_stream :254-275,
 * gz_header :281-296, sizing macros :67-121): a caller compiled against the reference headers links
 * against this library unchanged.  Only the one-shot zsc_* path is implemented on the GPU; z_stream
 * is declared for ABI completeness (the streaming state machine is out of scope).
 */
#ifndef ZLIB_TYPES_PUB_H
#define ZLIB_TYPES_PUB_H

#include "zsc/zsc_conf_global_types.h"

#ifndef Z_NULL
#define Z_NULL NULL
#endif

typedef U16 Pos;

/* ---- compile-time sizing (values pinned by the reference's Bounds test) ---- */
#define Z_DEFLATE_OUTPUT_BOUND(n) ((n) + (((n) + 7) >> 3) + (((n) + 63) >> 6) + 5 + 18 + 2)
#define Z_DEFLATE_OUTPUT_BOUND_BLOCKS(n, mbl) \
    (Z_DEFLATE_OUTPUT_BOUND((n)) + (Z_DEFLATE_OUTPUT_BOUND((n)) / (mbl) + 1) * 4)
#define Z_DEFLATE_STATE_SIZE 6400
#define Z_INFLATE_STATE_SIZE 7600
#define Z_COMPRESS_WORK_SIZE2(wbits, mlevel)                                         \
    (Z_DEFLATE_STATE_SIZE + (1 << (wbits)) * 2 * sizeof(U8) + (1 << (wbits)) * 2 * sizeof(Pos) + \
     (1 << ((mlevel) + 7)) * sizeof(Pos) + (1 << ((mlevel) + 6)) * (sizeof(U16) + 2))
#define Z_UNCOMPRESS_WORK_SIZE2(wbits) (Z_INFLATE_STATE_SIZE + (1 << (wbits)) * sizeof(U8))

enum { MAX_MEM_LEVEL = 9, DEF_MEM_LEVEL = 8, MAX_WBITS = 15, DEF_WBITS = MAX_WBITS };
enum { GZIP_CODE = 0x10 };

typedef enum {
    Z_NO_FLUSH = 0, Z_PARTIAL_FLUSH = 1, Z_SYNC_FLUSH = 2, Z_FULL_FLUSH = 3, Z_FINISH = 4,
    Z_BLOCK = 5, Z_TREES = 6
} ZlibFlush;

typedef enum {
    Z_OK = 0, Z_STREAM_END = 1, Z_NEED_DICT = 2, Z_ERRNO = -1, Z_STREAM_ERROR = -2,
    Z_DATA_ERROR = -3, Z_MEM_ERROR = -4, Z_BUF_ERROR = -5, Z_VERSION_ERROR = -6
} ZlibReturn;

enum { Z_NO_COMPRESSION = 0, Z_BEST_SPEED = 1, Z_BEST_COMPRESSION = 9, Z_DEFAULT_COMPRESSION = -1 };

typedef enum {
    Z_FILTERED = 1, Z_HUFFMAN_ONLY = 2, Z_RLE = 3, Z_FIXED = 4, Z_DEFAULT_STRATEGY = 0
} ZlibStrategy;

typedef enum { Z_BINARY = 0, Z_TEXT = 1, Z_ASCII = Z_TEXT, Z_UNKNOWN = 2 } ZlibDataType;
typedef enum { Z_DEFLATED = 8 } ZlibMethod;

struct internal_state;

typedef struct z_stream_s {
    const U8 *next_in;  U32 avail_in;  U32 total_in;
    U8 *next_out;       U32 avail_out; U32 total_out;
    U8 *next_work;      U32 avail_work;
    const U8 *msg;
    struct internal_state *state;
    ZlibDataType data_type;
    U32 adler;
    U32 reserved;
} z_stream;

typedef struct gz_header_s {
    I32 text;  U32 time;  I32 xflags;  I32 os;
    U8 *extra; U32 extra_len; U32 extra_max;
    U8 *name;  U32 name_max;
    U8 *comment; U32 comm_max;
    I32 hcrc;  I32 done;
} gz_header;

#endif
