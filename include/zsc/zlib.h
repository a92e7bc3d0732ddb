/* The slice of the reference's zlib.h that lies on the accelerated path: checksums and the
 * version / error-string helpers (reference include/zsc/zlib.h:1153-1208, :120, :1262).
 * The streaming z_stream API (deflateInit2 / deflate / inflate ...) is NOT exported by this
 * engine.
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

#ifdef __cplusplus
}
#endif
#endif
