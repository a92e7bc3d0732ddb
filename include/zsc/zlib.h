/* The slice of the reference's zlib.h this engine serves: checksums, the version / error-string helpers
 * (reference include/zsc/zlib.h:1153-1208, :120, :1262) and the z_stream API (reference :150-990, :1225-1256) with the
 * reference's memory model — the stream state is carved out of strm->next_work / avail_work, sized by
 * deflateWorkSize2 / inflateWorkSize2 (zsc_b200/csrc/host/zsc_stream.c).  Not served through z_stream: the gzip
 * wrapper on the inflate side (use zsc_uncompress_gzip*), deflateParams / deflateTune / deflatePrime / deflateSetHeader,
 * inflateBack*, inflatePrime, inflateGetHeader, inflateGetDictionary / deflateGetDictionary.
 */
#ifndef ZLIB_H
#define ZLIB_H

#include "zsc/zlib_types_pub.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ZLIB_VERSION "1.2.11.f-zsc-b200-v0"

const U8 *zlibVersion(void);
const U8 *zError(I32 err);

/* adler == 0 and buf == NULL return the initial value (1 / 0), as the reference does. */
U32 adler32(U32 adler, const U8 *buf, U32 len);
U32 adler32_z(U32 adler, const U8 *buf, z_size_t len);
U32 crc32(U32 crc, const U8 *buf, U32 len);
U32 crc32_z(U32 crc, const U8 *buf, z_size_t len);

/* ---- z_stream API: same prototypes, return codes and flush values as the reference ---- */
ZlibReturn deflateInit_(z_stream *strm, I32 level, const U8 *version, I32 stream_size);
ZlibReturn deflateInit2_(z_stream *strm, I32 level, ZlibMethod method, I32 windowBits, I32 memLevel,
                         ZlibStrategy strategy, const U8 *version, I32 stream_size);
ZlibReturn deflate(z_stream *strm, ZlibFlush flush);
ZlibReturn deflateEnd(z_stream *strm);
ZlibReturn deflateReset(z_stream *strm);
ZlibReturn deflateSetDictionary(z_stream *strm, const U8 *dictionary, U32 dictLength);
ZlibReturn deflateBoundNoStream(U32 sourceLen, I32 level, I32 windowBits, I32 memLevel, gz_header *gz_head, U32 *size_out);
ZlibReturn deflateWorkSize(U32 *size_out);
ZlibReturn deflateWorkSize2(I32 window_bits, I32 mem_level, U32 *size_out);
ZlibReturn inflateInit_(z_stream *strm, const U8 *version, I32 stream_size);
ZlibReturn inflateInit2_(z_stream *strm, I32 windowBits, const U8 *version, I32 stream_size);
ZlibReturn inflate(z_stream *strm, ZlibFlush flush);
ZlibReturn inflateEnd(z_stream *strm);
ZlibReturn inflateReset(z_stream *strm);
ZlibReturn inflateReset2(z_stream *strm, I32 windowBits);
ZlibReturn inflateSetDictionary(z_stream *strm, const U8 *dictionary, U32 dictLength);
ZlibReturn inflateSync(z_stream *strm);
ZlibReturn inflateWorkSize(U32 *size_out);
ZlibReturn inflateWorkSize2(I32 windowBits, U32 *size_out);
#define deflateInit(strm, level) deflateInit_((strm), (level), (const U8 *)ZLIB_VERSION, (I32)sizeof(z_stream))
#define inflateInit(strm) inflateInit_((strm), (const U8 *)ZLIB_VERSION, (I32)sizeof(z_stream))
#define deflateInit2(strm, level, method, windowBits, memLevel, strategy) \
    deflateInit2_((strm), (level), (method), (windowBits), (memLevel), (strategy), (const U8 *)ZLIB_VERSION, (I32)sizeof(z_stream))
#define inflateInit2(strm, windowBits) inflateInit2_((strm), (windowBits), (const U8 *)ZLIB_VERSION, (I32)sizeof(z_stream))

#ifdef __cplusplus
}
#endif
#endif
