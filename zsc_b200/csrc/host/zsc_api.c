/* zsc_api.c — the zsc_pub.h surface in host C, in front of the B200 engine.
 *
 * Mirrors the reference's all-in-one layer (src/zsc_compress.c, src/zsc_uncompr.c) call for call:
 * same argument validation order, same ZSC_ASSERT on NULL arguments, same return codes, same
 * in/out length conventions.  What differs is everything below it: where the reference drives its
 * own deflate()/inflate() state machines on the CPU, these functions hand the whole buffer to the
 * GPU engine (zscgpu.h).  There is no CPU codec in this library: if no B200 is present the calls
 * fail with Z_MEM_ERROR and a ZSC_WARN naming the cause.
 *
 * The caller-supplied work buffer keeps its contract (it must be at least as large as the
 * *_get_min_work_buf_size functions report, else Z_MEM_ERROR) although the engine keeps its state
 * in device arenas fixed at init; the sizes reported are the reference's, bit for bit
 * (src/deflate.c:857-902, src/inflate.c:249-276), because integrators size static buffers by them.
 */
#include "zsc/zsc_conf_private.h"
#include "zsc/zsc_pub.h"
#include "zscgpu.h"

/* sizeof(deflate_state) / sizeof(inflate_state) of the reference (include/zsc/deflate.h:119-290,
 * include/zsc/inflate.h): they are part of the pinned work-buffer sizes. LP64 / ILP32 values. */
#define ZSC_REF_DEFLATE_STATE ((U32)(sizeof(void *) == 8 ? 5920u : 5828u))
#define ZSC_REF_INFLATE_STATE ((U32)(sizeof(void *) == 8 ? 7152u : 7120u))

ZSC_COMPILE_ASSERT(Z_DEFLATE_STATE_SIZE >= 5920, deflate_state_margin);
ZSC_COMPILE_ASSERT(Z_INFLATE_STATE_SIZE >= 7152, inflate_state_margin);

/* ------------------------------------------------------------------ helpers */

typedef struct { I32 wrap; I32 wbits; } zsc_wb;

/* window_bits decoding of deflateInit2_ (reference src/deflate.c:302-331) */
ZSC_PRIVATE ZlibReturn zsc_deflate_wbits(I32 window_bits, I32 mem_level, I32 level, I32 strategy, zsc_wb *out)
{
    I32 wrap = 1;
    if (level == Z_DEFAULT_COMPRESSION) level = 6;
    if (window_bits < 0) { wrap = 0; window_bits = -window_bits; }
    else if (window_bits > 15) { wrap = 2; window_bits -= 16; }
    if (mem_level < 1 || mem_level > MAX_MEM_LEVEL || window_bits < 8 || window_bits > 15 ||
        level < 0 || level > 9 || strategy < 0 || strategy > Z_FIXED || (window_bits == 8 && wrap != 1)) {
        return Z_STREAM_ERROR;
    }
    if (window_bits == 8) window_bits = 9;
    out->wrap = wrap; out->wbits = window_bits;
    return Z_OK;
}

ZSC_PRIVATE U32 zsc_strlen0(const U8 *s) { U32 n = 0; while (s[n]) n++; return n + 1; }

/* bitwise CRC-32 for the few bytes of a gzip HEADER (FHCRC); payload CRCs come from the GPU */
ZSC_PRIVATE U32 zsc_hdr_crc(const U8 *p, U32 n)
{
    U32 c = 0xFFFFFFFFu;
    for (U32 i = 0; i < n; i++) {
        c ^= p[i];
        for (I32 k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
    }
    return c ^ 0xFFFFFFFFu;
}

ZSC_PRIVATE zscgpu_engine *zsc_engine(const char *who)
{
    zscgpu_engine *e = zscgpu_global();
    if (e == Z_NULL) {
        ZSC_WARN2("In %s, the GPU engine is unavailable: %s", who, zscgpu_last_error(Z_NULL));
    }
    return e;
}

/* ------------------------------------------------------------------ size checks */

ZlibReturn zsc_compress_get_min_work_buf_size2(I32 window_bits, I32 mem_level, U32 *size_out)
{
    ZSC_ASSERT(size_out != Z_NULL);
    *size_out = U32_MAX;
    if (window_bits < 0) window_bits = -window_bits;
    else if (window_bits > 15) window_bits -= 16;
    if (window_bits == 8) window_bits = 9;
    if (mem_level < 1 || mem_level > MAX_MEM_LEVEL || window_bits < 8 || window_bits > 15) {
        ZSC_WARN2("In zsc_compress_get_min_work_buf_size2(), bad mem_level (%d) or window_bits (%d).",
                  mem_level, window_bits);
        return Z_STREAM_ERROR;
    }
    U32 wsize = 1u << window_bits, hsize = 1u << (mem_level + 7), lit = 1u << (mem_level + 6);
    *size_out = ZSC_REF_DEFLATE_STATE + wsize * 2u * (U32)sizeof(U8) + wsize * 2u * (U32)sizeof(Pos) +
                hsize * (U32)sizeof(Pos) + lit * ((U32)sizeof(U16) + 2u);
    return Z_OK;
}

ZlibReturn zsc_compress_get_min_work_buf_size(U32 *size_out)
{
    return zsc_compress_get_min_work_buf_size2(DEF_WBITS, DEF_MEM_LEVEL, size_out);
}

/* deflateBoundNoStream (reference src/deflate.c:761-849) */
ZSC_PRIVATE ZlibReturn zsc_bound(U32 source_len, I32 level, I32 window_bits, I32 mem_level,
                                 gz_header *gz_head, U32 *size_out)
{
    ZSC_ASSERT(size_out != Z_NULL);
    *size_out = U32_MAX;
    I32 wrap = 1;
    if (window_bits < 0) { wrap = 0; window_bits = -window_bits; }
    else if (window_bits > 15) { wrap = 2; window_bits -= 16; }
    if (mem_level < 1 || mem_level > MAX_MEM_LEVEL || window_bits < 8 || window_bits > 15 ||
        (window_bits == 8 && wrap != 1)) {
        return Z_STREAM_ERROR;
    }
    U32 wraplen = 0;
    if (wrap == 1) wraplen = 6 + 4;
    else if (wrap == 2) {
        wraplen = 18;
        if (gz_head != Z_NULL) {
            if (gz_head->extra != Z_NULL) wraplen += 2 + gz_head->extra_len;
            if (gz_head->name != Z_NULL) wraplen += zsc_strlen0(gz_head->name);
            if (gz_head->comment != Z_NULL) wraplen += zsc_strlen0(gz_head->comment);
            if (gz_head->hcrc) wraplen += 2;
        }
    }
    if (window_bits != 15 || mem_level + 7 != 15 || level == Z_NO_COMPRESSION) {
        *size_out = source_len + ((source_len + 7) >> 3) + ((source_len + 63) >> 6) + 5 + wraplen;
    } else {
        *size_out = source_len + (source_len >> 12) + (source_len >> 14) + (source_len >> 25) + 13 - 6 + wraplen;
    }
    return Z_OK;
}

ZlibReturn zsc_compress_get_max_output_size_gzip2(U32 source_len, U32 max_block_len, I32 level,
                                                  I32 window_bits, I32 mem_level,
                                                  gz_header *gz_header, U32 *size_out)
{
    U32 inter = U32_MAX;
    ZlibReturn err = zsc_bound(source_len, level, window_bits, mem_level, gz_header, &inter);
    if (err != Z_OK) {
        ZSC_WARN1("In zsc_compress_get_max_output_size_gzip2(), could not get deflate output bound, error %d.", err);
        return err;
    }
    ZSC_ASSERT(max_block_len != 0);
    /* every section costs a 4-byte full-flush marker (reference src/zsc_compress.c:219-230) */
    U32 extra = ((inter / max_block_len) + 1) * 4;
    err = zsc_bound(source_len + extra, level, window_bits, mem_level, gz_header, size_out);
    if (err != Z_OK) {
        ZSC_WARN1("In zsc_compress_get_max_output_size_gzip2(), could not recalculate deflate output bound, error %d.", err);
    }
    return err;
}

ZlibReturn zsc_compress_get_max_output_size2(U32 source_len, U32 max_block_len, I32 level,
                                             I32 window_bits, I32 mem_level, U32 *size_out)
{
    return zsc_compress_get_max_output_size_gzip2(source_len, max_block_len, level, window_bits, mem_level, Z_NULL, size_out);
}

ZlibReturn zsc_compress_get_max_output_size_gzip(U32 source_len, U32 max_block_len, I32 level,
                                                 gz_header *gz_header, U32 *size_out)
{
    return zsc_compress_get_max_output_size_gzip2(source_len, max_block_len, level, DEF_WBITS + GZIP_CODE,
                                                  DEF_MEM_LEVEL, gz_header, size_out);
}

ZlibReturn zsc_compress_get_max_output_size(U32 source_len, U32 max_block_len, I32 level, U32 *size_out)
{
    return zsc_compress_get_max_output_size2(source_len, max_block_len, level, DEF_WBITS, DEF_MEM_LEVEL, size_out);
}

ZlibReturn zsc_uncompress_get_min_work_buf_size2(I32 window_bits, U32 *size_out)
{
    ZSC_ASSERT(size_out != Z_NULL);
    if (window_bits < 0) window_bits = -window_bits;
    else if (window_bits < 48) window_bits &= 15;
    if (window_bits && (window_bits < 8 || window_bits > 15)) {
        ZSC_WARN1("Cannot determine working size for windowBits = %d", window_bits);
        return Z_STREAM_ERROR;
    }
    *size_out = ZSC_REF_INFLATE_STATE + (1u << window_bits) * (U32)sizeof(U8);
    return Z_OK;
}

ZlibReturn zsc_uncompress_get_min_work_buf_size(U32 *size_out)
{
    return zsc_uncompress_get_min_work_buf_size2(DEF_WBITS, size_out);
}

/* ------------------------------------------------------------------ compress */

ZlibReturn zsc_compress_gzip2(U8 *dest, U32 *dest_len, const U8 *source, U32 source_len,
                              U32 max_block_len, U8 *work, U32 work_len, I32 level,
                              I32 window_bits, I32 mem_level, ZlibStrategy strategy,
                              gz_header *gz_header)
{
    ZSC_ASSERT(source != Z_NULL);
    ZSC_ASSERT(dest != Z_NULL);
    ZSC_ASSERT(dest_len != Z_NULL);
    ZSC_ASSERT(work != Z_NULL);

    U32 dest_len_in = *dest_len;
    *dest_len = 0;

    U32 min_work = U32_MAX;
    ZlibReturn err = zsc_compress_get_min_work_buf_size2(window_bits, mem_level, &min_work);
    if (err != Z_OK) {
        ZSC_WARN1("In zsc_compress_gzip2(), could not get min work buf size, error %d.", err);
        return err;
    }
    if (work_len < min_work) {
        ZSC_WARN2("In zsc_compress_gzip2(), working memory (%u B) was smaller than required (%u B).", work_len, min_work);
        return Z_MEM_ERROR;
    }
    zsc_wb wb;
    err = zsc_deflate_wbits(window_bits, mem_level, level, (I32)strategy, &wb);
    if (err != Z_OK) {
        ZSC_WARN1("In zsc_compress_gzip2, could not deflateInit, error %d.", err);
        return err;
    }
    if (gz_header != Z_NULL && wb.wrap != 2) {
        /* deflateSetHeader refuses a header on a non-gzip stream (reference src/deflate.c:582-590) */
        ZSC_WARN1("In zsc_compress_gzip2(), could not set deflate header, error %d.", Z_STREAM_ERROR);
        return Z_STREAM_ERROR;
    }
    U32 bound = U32_MAX;
    err = zsc_compress_get_max_output_size_gzip2(source_len, max_block_len, level, window_bits, mem_level, gz_header, &bound);
    if (err != Z_OK) {
        ZSC_WARN1("In, zsc_compress_gzip2(), could not get deflate output bound, error %d.", err);
        return err;
    }
    ZSC_ASSERT(max_block_len != 0);

    zscgpu_engine *e = zsc_engine("zsc_compress_gzip2()");
    if (e == Z_NULL) return Z_MEM_ERROR;

    /* gzip wrapper: the header is written here, the deflate body and the CRC come from the GPU */
    U32 hlen = 0;
    if (wb.wrap == 2) {
        U32 need = 10;
        if (gz_header != Z_NULL) {
            if (gz_header->extra != Z_NULL) need += 2 + (gz_header->extra_len & 0xFFFF);
            if (gz_header->name != Z_NULL) need += zsc_strlen0(gz_header->name);
            if (gz_header->comment != Z_NULL) need += zsc_strlen0(gz_header->comment);
            if (gz_header->hcrc) need += 2;
        }
        if (dest_len_in < need + 8) {
            ZSC_WARN2("In zsc_compress_gzip2(), output buffer (%u bytes) was smaller than bound (%u bytes).", dest_len_in, bound);
            return Z_BUF_ERROR;
        }
        U8 xfl = (U8)(level == 9 ? 2 : (((I32)strategy >= Z_HUFFMAN_ONLY || (level < 2 && level != Z_DEFAULT_COMPRESSION)) ? 4 : 0));
        U8 *h = dest;
        h[0] = 31; h[1] = 139; h[2] = 8;
        if (gz_header == Z_NULL) {
            h[3] = 0; h[4] = h[5] = h[6] = h[7] = 0; h[8] = xfl; h[9] = 3 /* OS_CODE: unix */;
            hlen = 10;
        } else {
            h[3] = (U8)((gz_header->text ? 1 : 0) + (gz_header->hcrc ? 2 : 0) + (gz_header->extra == Z_NULL ? 0 : 4) +
                        (gz_header->name == Z_NULL ? 0 : 8) + (gz_header->comment == Z_NULL ? 0 : 16));
            h[4] = (U8)(gz_header->time & 0xff); h[5] = (U8)((gz_header->time >> 8) & 0xff);
            h[6] = (U8)((gz_header->time >> 16) & 0xff); h[7] = (U8)((gz_header->time >> 24) & 0xff);
            h[8] = xfl; h[9] = (U8)(gz_header->os & 0xff);
            hlen = 10;
            if (gz_header->extra != Z_NULL) {
                U32 xl = gz_header->extra_len & 0xFFFF;
                h[hlen++] = (U8)(xl & 0xff); h[hlen++] = (U8)(xl >> 8);
                zmemcpy(h + hlen, gz_header->extra, xl); hlen += xl;
            }
            if (gz_header->name != Z_NULL) { U32 n = zsc_strlen0(gz_header->name); zmemcpy(h + hlen, gz_header->name, n); hlen += n; }
            if (gz_header->comment != Z_NULL) { U32 n = zsc_strlen0(gz_header->comment); zmemcpy(h + hlen, gz_header->comment, n); hlen += n; }
            if (gz_header->hcrc) { U32 c = zsc_hdr_crc(h, hlen); h[hlen++] = (U8)(c & 0xff); h[hlen++] = (U8)((c >> 8) & 0xff); }
        }
    }

    zscgpu_deflate_params p;
    p.max_block_len = max_block_len; p.level = level; p.strategy = (I32)strategy;
    p.wrap = wb.wrap; p.window_bits = wb.wbits; p.part = 0; p.hist_len = 0;
    zscgpu_result res;
    U32 cap = dest_len_in;
    if (wb.wrap == 2) cap -= 8;                       /* room for CRC32 + ISIZE */
    int rc = zscgpu_compress_host(e, dest, cap, source, source_len, &p, hlen, &res);
    if (rc != 0) {
        ZSC_WARN2("In zsc_compress_gzip2(), the GPU engine failed (%d): %s", rc, zscgpu_last_error(e));
        return rc == ZSCGPU_ERR_ARG ? Z_STREAM_ERROR : Z_MEM_ERROR;
    }
    if (res.ret != Z_OK) {
        ZSC_WARN1("In zsc_compress_gzip2(), deflate loop ended with error code %d.", res.ret);
        if (dest_len_in < bound) {
            ZSC_WARN2("In zsc_compress_gzip2(), output buffer (%u bytes) was smaller than bound (%u bytes). "
                      "Output may not have fit in the buffer.", dest_len_in, bound);
        }
        if (res.ret == Z_BUF_ERROR) *dest_len = hlen + res.produced;      /* total_out: what was written before the room ran out (reference src/zsc_compress.c:140) */
        return (ZlibReturn)res.ret;
    }
    U32 total = hlen + res.produced;
    if (wb.wrap == 2) {
        U8 *t = dest + total;
        t[0] = (U8)(res.check & 0xff); t[1] = (U8)((res.check >> 8) & 0xff);
        t[2] = (U8)((res.check >> 16) & 0xff); t[3] = (U8)((res.check >> 24) & 0xff);
        t[4] = (U8)(source_len & 0xff); t[5] = (U8)((source_len >> 8) & 0xff);
        t[6] = (U8)((source_len >> 16) & 0xff); t[7] = (U8)((source_len >> 24) & 0xff);
        total += 8;
    }
    *dest_len = total;
    return Z_OK;
}

ZlibReturn zsc_compress2(U8 *dest, U32 *dest_len, const U8 *source, U32 source_len,
                         U32 max_block_len, U8 *work, U32 work_len, I32 level,
                         I32 window_bits, I32 mem_level, ZlibStrategy strategy)
{
    return zsc_compress_gzip2(dest, dest_len, source, source_len, max_block_len, work, work_len, level,
                              window_bits, mem_level, strategy, Z_NULL);
}

ZlibReturn zsc_compress_gzip(U8 *dest, U32 *dest_len, const U8 *source, U32 source_len,
                             U32 max_block_len, U8 *work, U32 work_len, I32 level, gz_header *gz_header)
{
    return zsc_compress_gzip2(dest, dest_len, source, source_len, max_block_len, work, work_len, level,
                              DEF_WBITS + GZIP_CODE, DEF_MEM_LEVEL, Z_DEFAULT_STRATEGY, gz_header);
}

ZlibReturn zsc_compress(U8 *dest, U32 *dest_len, const U8 *source, U32 source_len,
                        U32 max_block_len, U8 *work, U32 work_len, I32 level)
{
    return zsc_compress2(dest, dest_len, source, source_len, max_block_len, work, work_len, level,
                         DEF_WBITS, DEF_MEM_LEVEL, Z_DEFAULT_STRATEGY);
}

/* ------------------------------------------------------------------ uncompress */

/* Parse a gzip member header (RFC 1952; what the reference does in inflate()'s FLAGS..HCRC states,
 * src/inflate.c:786-954).  Returns the header length, 0 when truncated, U32_MAX when malformed. */
ZSC_PRIVATE U32 zsc_gzip_header(const U8 *s, U32 n, gz_header *gh)
{
    if (n < 2) return 0;
    if (s[0] != 31 || s[1] != 139) return U32_MAX;
    if (n < 4) return 0;
    if (s[2] != 8) return U32_MAX;                /* the reference checks method and flags as soon as it */
    U32 flg = s[3];                               /* has pulled them (src/inflate.c:786-801) */
    if (flg & 0xe0) return U32_MAX;
    if (n < 10) return 0;
    if (gh != Z_NULL) {
        gh->text = (I32)(flg & 1);
        gh->time = (U32)s[4] | ((U32)s[5] << 8) | ((U32)s[6] << 16) | ((U32)s[7] << 24);
        gh->xflags = s[8]; gh->os = s[9];
    }
    U32 p = 10;
    if (flg & 4) {
        if (p + 2 > n) return 0;
        U32 xl = (U32)s[p] | ((U32)s[p + 1] << 8);
        p += 2;
        if (p + xl > n) return 0;
        if (gh != Z_NULL) {
            gh->extra_len = xl;
            if (gh->extra != Z_NULL) zmemcpy(gh->extra, s + p, xl < gh->extra_max ? xl : gh->extra_max);
        }
        p += xl;
    } else if (gh != Z_NULL) gh->extra = Z_NULL;
    if (flg & 8) {
        U32 k = 0;
        for (;;) {
            if (p >= n) return 0;
            U8 c = s[p++];
            if (gh != Z_NULL && gh->name != Z_NULL && k < gh->name_max) gh->name[k++] = c;
            if (c == 0) break;
        }
    } else if (gh != Z_NULL) gh->name = Z_NULL;
    if (flg & 16) {
        U32 k = 0;
        for (;;) {
            if (p >= n) return 0;
            U8 c = s[p++];
            if (gh != Z_NULL && gh->comment != Z_NULL && k < gh->comm_max) gh->comment[k++] = c;
            if (c == 0) break;
        }
    } else if (gh != Z_NULL) gh->comment = Z_NULL;
    if (flg & 2) {
        if (p + 2 > n) return 0;
        U32 want = (U32)s[p] | ((U32)s[p + 1] << 8);
        if ((zsc_hdr_crc(s, p) & 0xFFFF) != want) return U32_MAX;
        p += 2;
    }
    if (gh != Z_NULL) { gh->hcrc = (I32)((flg >> 1) & 1); gh->done = 1; }
    return p;
}

ZlibReturn zsc_uncompress_gzip2(U8 *dest, U32 *dest_len, const U8 *source, U32 *source_len,
                                U8 *work, U32 work_len, I32 window_bits, gz_header *gz_head)
{
    ZSC_ASSERT(source != Z_NULL);
    ZSC_ASSERT(source_len != Z_NULL);
    ZSC_ASSERT(dest != Z_NULL);
    ZSC_ASSERT(dest_len != Z_NULL);
    ZSC_ASSERT(work != Z_NULL);

    U32 dest_len_in = *dest_len, source_len_in = *source_len;
    *dest_len = 0;
    *source_len = 0;

    U32 min_work = U32_MAX;
    ZlibReturn err = zsc_uncompress_get_min_work_buf_size2(window_bits, &min_work);
    if (err != Z_OK) {
        ZSC_WARN1("In zsc_uncompress_safe_gzip2(), could not get work buffer size, error %d.", err);
        return err;
    }
    if (work_len < min_work) {
        ZSC_WARN2("In zsc_uncompress_safe_gzip2(), work buffer (%u B) is smaller than required (%u B).", work_len, min_work);
        return Z_MEM_ERROR;
    }
    /* wrapper selection of inflateReset2 (reference src/inflate.c:333-368) */
    I32 wrap, wbits;
    if (window_bits < 0) { wrap = 0; wbits = -window_bits; }
    else { wrap = (window_bits >> 4) + 5; wbits = window_bits < 48 ? (window_bits & 15) : window_bits; }
    if (wbits == 0) wbits = 15;
    if (gz_head != Z_NULL) {
        /* inflateGetHeader (reference src/inflate.c:1488-1501): only gzip-capable streams */
        if ((wrap & 2) == 0) {
            ZSC_WARN1("In zsc_uncompress_safe_gzip2(), could not get header, error %d.", Z_STREAM_ERROR);
            return Z_STREAM_ERROR;
        }
        gz_head->done = 0;
    }
    zscgpu_engine *e = zsc_engine("zsc_uncompress_gzip2()");
    if (e == Z_NULL) return Z_MEM_ERROR;

    I32 is_gzip = 0;
    if (wrap & 2) {
        if ((wrap & 1) == 0) is_gzip = 1;                                  /* gzip only */
        else is_gzip = (source_len_in >= 2 && source[0] == 31 && source[1] == 139);   /* auto-detect */
    }
    zscgpu_result res;
    int rc;
    if (!is_gzip) {
        if (gz_head != Z_NULL) gz_head->done = -1;
        rc = zscgpu_uncompress_host(e, dest, dest_len_in, source, source_len_in, (wrap ? 1 : 0) | (wbits << 8), &res);
        if (rc != 0) {
            ZSC_WARN2("In zsc_uncompress_safe_gzip2(), the GPU engine failed (%d): %s", rc, zscgpu_last_error(e));
            return Z_MEM_ERROR;
        }
        *dest_len = res.produced;
        *source_len = res.consumed;
    } else {
        U32 hl = zsc_gzip_header(source, source_len_in, gz_head);
        if (hl == U32_MAX) { ZSC_WARN1("In zsc_uncompress_safe_gzip2(), inflate loop failed with error %d.", Z_DATA_ERROR); return Z_DATA_ERROR; }
        if (hl == 0) { ZSC_WARN1("In zsc_uncompress_safe_gzip2(), inflate loop failed with error %d.", Z_BUF_ERROR); return Z_BUF_ERROR; }
        rc = zscgpu_uncompress_host(e, dest, dest_len_in, source + hl, source_len_in - hl, wbits << 8, &res);
        if (rc != 0) {
            ZSC_WARN2("In zsc_uncompress_safe_gzip2(), the GPU engine failed (%d): %s", rc, zscgpu_last_error(e));
            return Z_MEM_ERROR;
        }
        *dest_len = res.produced;
        *source_len = hl + res.consumed;
        if (res.ret == Z_OK) {
            const U8 *t = source + hl + res.consumed;
            /* CHECK then LENGTH, each as soon as its four bytes are there (src/inflate.c:1322-1354) */
            U32 left = source_len_in - (hl + res.consumed);
            if (left < 4) { res.ret = Z_BUF_ERROR; *source_len = source_len_in; }
            else {
                U32 crc = 0;
                rc = res.produced ? zscgpu_checksum_host(e, 1, 0, dest, res.produced, &crc) : 0;
                U32 want = (U32)t[0] | ((U32)t[1] << 8) | ((U32)t[2] << 16) | ((U32)t[3] << 24);
                if (rc != 0) return Z_MEM_ERROR;
                if (crc != want) { res.ret = Z_DATA_ERROR; *source_len = source_len_in; }
                else if (left < 8) { res.ret = Z_BUF_ERROR; *source_len = source_len_in; }
                else {
                    U32 isize = (U32)t[4] | ((U32)t[5] << 8) | ((U32)t[6] << 16) | ((U32)t[7] << 24);
                    if (isize != res.produced) { res.ret = Z_DATA_ERROR; *source_len = source_len_in; }
                    else *source_len += 8;
                }
            }
        }
    }
    if (res.ret != Z_OK) {
        ZSC_WARN1("In zsc_uncompress_safe_gzip2(), inflate loop failed with error %d.", res.ret);
        return (ZlibReturn)res.ret;
    }
    return Z_OK;
}

ZlibReturn zsc_uncompress2(U8 *dest, U32 *dest_len, const U8 *source, U32 *source_len,
                           U8 *work, U32 work_len, I32 window_bits)
{
    return zsc_uncompress_gzip2(dest, dest_len, source, source_len, work, work_len, window_bits, Z_NULL);
}

ZlibReturn zsc_uncompress(U8 *dest, U32 *dest_len, const U8 *source, U32 *source_len, U8 *work, U32 work_len)
{
    return zsc_uncompress2(dest, dest_len, source, source_len, work, work_len, DEF_WBITS);
}

ZlibReturn zsc_uncompress_gzip(U8 *dest, U32 *dest_len, const U8 *source, U32 *source_len,
                               U8 *work, U32 work_len, gz_header *gz_head)
{
    return zsc_uncompress_gzip2(dest, dest_len, source, source_len, work, work_len, DEF_WBITS + GZIP_CODE, gz_head);
}

/* ------------------------------------------------------------------ checksums, version, errors */

U32 adler32_z(U32 adler, const U8 *buf, z_size_t len)
{
    if (buf == Z_NULL) return 1u;                         /* reference src/adler32.c:82-84 */
    zscgpu_engine *e = zsc_engine("adler32_z()");
    ZSC_ASSERT(e != Z_NULL);
    U32 out = adler;
    if (len == 0) return adler;
    int rc = zscgpu_checksum_host(e, 0, adler, buf, (uint64_t)len, &out);
    ZSC_ASSERT1(rc == 0, rc);
    (void)rc;
    return out;
}
U32 adler32(U32 adler, const U8 *buf, U32 len) { return adler32_z(adler, buf, len); }

U32 crc32_z(U32 crc, const U8 *buf, z_size_t len)
{
    if (buf == Z_NULL) return 0u;                         /* reference src/crc32.c:507 */
    zscgpu_engine *e = zsc_engine("crc32_z()");
    ZSC_ASSERT(e != Z_NULL);
    U32 out = crc;
    if (len == 0) return crc;
    int rc = zscgpu_checksum_host(e, 1, crc, buf, (uint64_t)len, &out);
    ZSC_ASSERT1(rc == 0, rc);
    (void)rc;
    return out;
}
U32 crc32(U32 crc, const U8 *buf, U32 len) { return crc32_z(crc, buf, len); }

const U8 *zlibVersion(void) { return (const U8 *)ZLIB_VERSION; }

const U8 *zError(I32 err)
{
    /* same table as the reference's z_errmsg (src/zutil.c:40-51), indexed by Z_NEED_DICT - err */
    static const char *const msg[10] = {
        "need dictionary", "stream end", "", "file error", "stream error", "data error",
        "insufficient memory", "buffer error", "incompatible version", "" };
    I32 i = Z_NEED_DICT - err;
    if (i < 0 || i > 9) i = 9;
    return (const U8 *)msg[i];
}
