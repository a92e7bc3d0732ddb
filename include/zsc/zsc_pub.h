/* zsc all-in-one API, B200 engine edition.
 *
 * Same 16 entry points, argument order, in/out conventions and return codes as reference
 * include/zsc/zsc_pub.h:86-411.  Each call compresses/decompresses one host buffer; the data path
 * runs entirely on the GPU (zscgpu.h); there is no CPU codec behind these functions.
 *
 *   compress:   dest/dest_len (in: capacity, out: bytes written), source/source_len,
 *               max_block_len = size of each independently decodable section (full-flush marker
 *               between sections), work/work_len = caller scratch (must be >= the size the
 *               *_get_min_work_buf_size functions report, else Z_MEM_ERROR)
 *   uncompress: dest_len and source_len are in/out (capacity/size in, produced/consumed out)
 *   returns:    Z_OK on success; Z_MEM_ERROR, Z_BUF_ERROR, Z_STREAM_ERROR, Z_DATA_ERROR as the
 *               reference (src/zsc_compress.c:83-159, src/zsc_uncompr.c:78-153)
 */
#ifndef ZSC_PUB_H
#define ZSC_PUB_H

#include "zsc/zsc_conf_global_types.h"
#include "zsc/zlib.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- buffer-size check functions (pure arithmetic; values pinned by the reference tests) ---- */
ZlibReturn zsc_compress_get_min_work_buf_size(U32 *size_out);
ZlibReturn zsc_compress_get_min_work_buf_size2(I32 window_bits, I32 mem_level, U32 *size_out);
ZlibReturn zsc_compress_get_max_output_size(U32 source_len, U32 max_block_len, I32 level, U32 *size_out);
ZlibReturn zsc_compress_get_max_output_size_gzip(U32 source_len, U32 max_block_len, I32 level,
                                                 gz_header *gz_header, U32 *size_out);
ZlibReturn zsc_compress_get_max_output_size2(U32 source_len, U32 max_block_len, I32 level,
                                             I32 window_bits, I32 mem_level, U32 *size_out);
ZlibReturn zsc_compress_get_max_output_size_gzip2(U32 source_len, U32 max_block_len, I32 level,
                                                  I32 window_bits, I32 mem_level,
                                                  gz_header *gz_header, U32 *size_out);
ZlibReturn zsc_uncompress_get_min_work_buf_size(U32 *size_out);
ZlibReturn zsc_uncompress_get_min_work_buf_size2(I32 window_bits, U32 *size_out);

/* ---- compress ---- */
ZlibReturn zsc_compress(U8 *dest, U32 *dest_len, const U8 *source, U32 source_len,
                        U32 max_block_len, U8 *work, U32 work_len, I32 level);
ZlibReturn zsc_compress_gzip(U8 *dest, U32 *dest_len, const U8 *source, U32 source_len,
                             U32 max_block_len, U8 *work, U32 work_len, I32 level,
                             gz_header *gz_header);
ZlibReturn zsc_compress2(U8 *dest, U32 *dest_len, const U8 *source, U32 source_len,
                         U32 max_block_len, U8 *work, U32 work_len, I32 level,
                         I32 window_bits, I32 mem_level, ZlibStrategy strategy);
ZlibReturn zsc_compress_gzip2(U8 *dest, U32 *dest_len, const U8 *source, U32 source_len,
                              U32 max_block_len, U8 *work, U32 work_len, I32 level,
                              I32 window_bits, I32 mem_level, ZlibStrategy strategy,
                              gz_header *gz_header);

/* ---- uncompress ---- */
ZlibReturn zsc_uncompress(U8 *dest, U32 *dest_len, const U8 *source, U32 *source_len,
                          U8 *work, U32 work_len);
ZlibReturn zsc_uncompress_gzip(U8 *dest, U32 *dest_len, const U8 *source, U32 *source_len,
                               U8 *work, U32 work_len, gz_header *gz_head);
ZlibReturn zsc_uncompress2(U8 *dest, U32 *dest_len, const U8 *source, U32 *source_len,
                           U8 *work, U32 work_len, I32 window_bits);
ZlibReturn zsc_uncompress_gzip2(U8 *dest, U32 *dest_len, const U8 *source, U32 *source_len,
                                U8 *work, U32 work_len, I32 window_bits, gz_header *gz_head);

#ifdef __cplusplus
}
#endif
#endif
