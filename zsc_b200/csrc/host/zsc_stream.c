/* zsc_stream.c — the z_stream API of the reference (include/zsc/zlib.h:150-990: deflateInit2_ / deflate / deflateEnd /
 * deflateSetDictionary / deflateReset, inflateInit2_ / inflate / inflateEnd / inflateSetDictionary / inflateSync /
 * inflateReset) in host C, on top of the B200 engine.  SURVEY.md §8(f) row 4.
 *
 * Same memory model as the reference (src/deflate.c:265-382, src/inflate.c:249-330): no allocation; the stream's state
 * is carved out of the caller's work buffer (strm->next_work / avail_work, sized by deflateWorkSize2 / inflateWorkSize2,
 * which return the reference's numbers).  What differs is the machine behind it:
 *
 *   deflate   input is gathered in the work buffer behind up to 2^window_bits bytes of history; whenever that buffer is
 *             full, or the caller flushes or finishes, the gathered bytes go to the GPU as ONE chunk with the history as
 *             its preset dictionary (zscgpu_deflate_params.hist_len) and come back as raw deflate blocks that end on a
 *             byte boundary: with an empty stored block (what Z_SYNC_FLUSH emits; the window carries over, so nothing
 *             is lost but 5 bytes per chunk), with a full flush (history dropped), or with the final block.  The zlib /
 *             gzip wrapper, DICTID and the running adler32 / crc32 (folded with zscgpu_adler32_combine) are the host's.
 *   inflate   the decoder state lives in a device slot (zscgpu_inflate_stream_*); a call stages new input, decodes up
 *             to one window of output into the work buffer and hands it out as the caller's space allows.  Data errors
 *             leave the stream waiting for inflateSync; a preset-dictionary header returns Z_NEED_DICT with strm->adler
 *             set, inflateSetDictionary checks the dictionary's adler32 and installs it as history.
 *
 * Every byte is compressed / decompressed on the GPU; there is no CPU codec here.  Throughput per call is that of one
 * small chunk (a launch and a PCIe round trip): the z_stream API exists for drop-in completeness, the bulk path is
 * zsc_compress / zsc_uncompress and the batched zscgpu_* calls.
 */
#include <string.h>
#include "zsc/zsc_conf_private.h"
#include "zsc/zsc_pub.h"
#include "zsc/zlib.h"
#include "zscgpu.h"

#define ZST_DEFLATE_MAGIC 0x5A534431u   /* "ZSD1" */
#define ZST_INFLATE_MAGIC 0x5A534931u   /* "ZSI1" */

struct internal_state {
    U32 magic;
    z_stream *strm;
    /* ---- deflate ---- */
    I32 level, strategy, wrap, wbits;
    U32 status;                  /* 0 nothing emitted yet, 1 busy, 2 final chunk emitted */
    I32 last_flush;
    U8 *hist_area;               /* wsize bytes; the history occupies its last `hist` bytes, the new data follow directly */
    U32 wsize, hist, fill, cap_new;
    U8 *pend;                    /* compressed bytes not yet handed out */
    U32 pend_len, pend_pos, pend_cap;
    U32 check;                   /* running adler32 (zlib) / crc32 (gzip) of the input */
    U32 total_len;               /* gzip ISIZE */
    U32 dictid;
    I32 have_dict;
    /* ---- inflate ---- */
    I32 slot;
    I32 iwrap;                   /* wrap argument of the engine: 0 raw / 1 zlib | max window bits << 8 */
    U32 in_left;                 /* bytes staged on the device that the decoder has not used yet */
    I32 mode;                    /* 0 running, 1 ended, 2 data error (waits for inflateSync), 3 needs a dictionary */
    U32 out_cap;                 /* size of the hand-out buffer in the work area (2^window_bits, at most ZSCGPU_STREAM_OUT_MAX) */
    I32 gz_member;               /* gzip: 0 header not parsed yet */
};

ZSC_PRIVATE zscgpu_engine *zst_engine(const char *who)
{
    zscgpu_engine *e = zscgpu_global();
    if (e == Z_NULL) { ZSC_WARN2("In %s, the GPU engine is unavailable: %s", who, zscgpu_last_error(Z_NULL)); }
    return e;
}

ZlibReturn deflateWorkSize2(I32 window_bits, I32 mem_level, U32 *size_out) { return zsc_compress_get_min_work_buf_size2(window_bits, mem_level, size_out); }
ZlibReturn deflateWorkSize(U32 *size_out) { return zsc_compress_get_min_work_buf_size(size_out); }
ZlibReturn inflateWorkSize2(I32 windowBits, U32 *size_out) { return zsc_uncompress_get_min_work_buf_size2(windowBits, size_out); }
ZlibReturn inflateWorkSize(U32 *size_out) { return zsc_uncompress_get_min_work_buf_size(size_out); }

ZlibReturn deflateBoundNoStream(U32 sourceLen, I32 level, I32 windowBits, I32 memLevel, gz_header *gz_head, U32 *size_out)
{
    /* one section as large as the source: the bound of zsc_compress without flush markers (reference src/deflate.c:761-849) */
    return zsc_compress_get_max_output_size_gzip2(sourceLen, sourceLen ? sourceLen : 1u, level, windowBits, memLevel, gz_head, size_out);
}

ZSC_PRIVATE struct internal_state *zst_state(z_stream *strm, U32 magic)
{
    if (strm == Z_NULL || strm->state == Z_NULL || strm->state->magic != magic || strm->state->strm != strm) return Z_NULL;
    return strm->state;
}

/* ================================================================== deflate */

ZlibReturn deflateInit2_(z_stream *strm, I32 level, ZlibMethod method, I32 windowBits, I32 memLevel,
                         ZlibStrategy strategy, const U8 *version, I32 stream_size)
{
    static const U8 my_version[] = ZLIB_VERSION;
    if (version == Z_NULL || version[0] != my_version[0] || stream_size != (I32)sizeof(z_stream)) {
        ZSC_WARN("deflateInit version error.");
        return Z_VERSION_ERROR;
    }
    if (strm == Z_NULL) { ZSC_WARN("deflateInit stream error: null stream."); return Z_STREAM_ERROR; }
    U32 work_size = U32_MAX;
    if (strm->next_work == Z_NULL || deflateWorkSize2(windowBits, memLevel, &work_size) != Z_OK || strm->avail_work < work_size) {
        ZSC_WARN("deflateInit stream error: no or too small a work buffer.");
        return Z_STREAM_ERROR;
    }
    strm->msg = Z_NULL;
    if (level == Z_DEFAULT_COMPRESSION) level = 6;
    I32 wrap = 1;
    if (windowBits < 0) { wrap = 0; windowBits = -windowBits; }
    else if (windowBits > 15) { wrap = 2; windowBits -= 16; }
    if (memLevel < 1 || memLevel > MAX_MEM_LEVEL || method != Z_DEFLATED || windowBits < 8 || windowBits > 15 || level < 0 ||
        level > 9 || (I32)strategy < 0 || strategy > Z_FIXED || (windowBits == 8 && wrap != 1)) {
        ZSC_WARN("deflateInit() bad arguments.");
        return Z_STREAM_ERROR;
    }
    if (windowBits == 8) windowBits = 9;
    if (zst_engine("deflateInit2_()") == Z_NULL) return Z_MEM_ERROR;

    /* carve the work buffer: state, history area, new-data area, pending-output area (its size bounds a chunk's output:
       stored blocks cost 5 bytes per 65535, plus a marker or final block, plus wrapper bytes) */
    U8 *w = strm->next_work;
    U32 off = (U32)((8u - ((uintptr_t)w & 7u)) & 7u);
    struct internal_state *s = (struct internal_state *)(w + off);
    off += ((U32)sizeof(struct internal_state) + 7u) & ~7u;
    memset(s, 0, sizeof(*s));
    s->wsize = 1u << windowBits;
    const U32 rest = work_size - off - s->wsize - 96u;
    s->cap_new = (U32)(((unsigned long long)rest * 8u) / 17u) & ~15u;
    s->pend_cap = rest - s->cap_new;
    s->hist_area = w + off;
    s->pend = s->hist_area + s->wsize + s->cap_new;
    s->magic = ZST_DEFLATE_MAGIC; s->strm = strm;
    s->level = level; s->strategy = (I32)strategy; s->wrap = wrap; s->wbits = windowBits;
    strm->state = s;
    strm->next_work += work_size; strm->avail_work -= work_size;
    return deflateReset(strm);
}

ZlibReturn deflateInit_(z_stream *strm, I32 level, const U8 *version, I32 stream_size)
{
    return deflateInit2_(strm, level, Z_DEFLATED, MAX_WBITS, DEF_MEM_LEVEL, Z_DEFAULT_STRATEGY, version, stream_size);
}

ZlibReturn deflateReset(z_stream *strm)
{
    struct internal_state *s = zst_state(strm, ZST_DEFLATE_MAGIC);
    if (s == Z_NULL) { ZSC_WARN("deflateReset: bad state."); return Z_STREAM_ERROR; }
    strm->total_in = strm->total_out = 0;
    strm->msg = Z_NULL;
    strm->data_type = Z_UNKNOWN;
    s->status = 0; s->last_flush = Z_NO_FLUSH;
    s->hist = s->fill = 0;
    s->pend_len = s->pend_pos = 0;
    s->check = (s->wrap == 2) ? 0u : 1u;
    s->total_len = 0; s->dictid = 0; s->have_dict = 0;
    strm->adler = s->check;
    return Z_OK;
}

ZlibReturn deflateSetDictionary(z_stream *strm, const U8 *dictionary, U32 dictLength)
{
    struct internal_state *s = zst_state(strm, ZST_DEFLATE_MAGIC);
    if (s == Z_NULL || dictionary == Z_NULL) { ZSC_WARN("deflateSetDictionary: bad state or null dictionary."); return Z_STREAM_ERROR; }
    /* reference src/deflate.c:408-470: not for gzip, for zlib only before the first deflate(), for raw with no input pending */
    if (s->wrap == 2 || (s->wrap == 1 && s->status != 0) || s->fill != 0) return Z_STREAM_ERROR;
    zscgpu_engine *e = zst_engine("deflateSetDictionary()");
    if (e == Z_NULL) return Z_MEM_ERROR;
    if (s->wrap == 1) {
        U32 id = 1;
        if (dictLength && zscgpu_checksum_host(e, 0, 1u, dictionary, dictLength, &id) != 0) return Z_MEM_ERROR;
        s->dictid = id; s->have_dict = 1;
        strm->adler = id;
    }
    const U32 keep = dictLength < s->wsize ? dictLength : s->wsize;     /* only the tail can ever be referenced */
    memcpy(s->hist_area + s->wsize - keep, dictionary + (dictLength - keep), keep);
    s->hist = keep;
    return Z_OK;
}

/* hand pending output to the caller */
ZSC_PRIVATE void zst_drain(z_stream *strm, struct internal_state *s)
{
    U32 n = s->pend_len - s->pend_pos;
    if (n > strm->avail_out) n = strm->avail_out;
    if (n) {
        memcpy(strm->next_out, s->pend + s->pend_pos, n);
        strm->next_out += n; strm->avail_out -= n; strm->total_out += n;
        s->pend_pos += n;
    }
    if (s->pend_pos == s->pend_len) s->pend_pos = s->pend_len = 0;
}

/* compress the gathered bytes as one chunk; kind 0: ends with an empty stored block (window kept), 1: the same but the
   history is dropped (full flush), 2: final block and trailer */
ZSC_PRIVATE ZlibReturn zst_emit(z_stream *strm, struct internal_state *s, I32 kind)
{
    zscgpu_engine *e = zst_engine("deflate()");
    if (e == Z_NULL) return Z_MEM_ERROR;
    U32 pos = 0;
    if (s->status == 0) {
        if (s->wrap == 1) {
            /* zlib header (reference src/deflate.c:1029-1057), DICTID behind it when a dictionary was set */
            const I32 lf = (s->strategy >= Z_HUFFMAN_ONLY || s->level < 2) ? 0 : (s->level < 6 ? 1 : (s->level == 6 ? 2 : 3));
            U32 h = ((8u + ((U32)(s->wbits - 8) << 4)) << 8) | ((U32)lf << 6);
            if (s->have_dict) h |= 0x20u;
            h += 31u - (h % 31u);
            s->pend[pos++] = (U8)(h >> 8); s->pend[pos++] = (U8)h;
            if (s->have_dict) { s->pend[pos++] = (U8)(s->dictid >> 24); s->pend[pos++] = (U8)(s->dictid >> 16); s->pend[pos++] = (U8)(s->dictid >> 8); s->pend[pos++] = (U8)s->dictid; }
        } else if (s->wrap == 2) {
            /* minimal gzip header (reference src/deflate.c:1066-1086 with no gz_header set) */
            static const U8 gzh[10] = {31, 139, 8, 0, 0, 0, 0, 0, 0, 3};
            memcpy(s->pend, gzh, 10);
            s->pend[8] = (U8)(s->level == 9 ? 2 : ((s->strategy >= Z_HUFFMAN_ONLY || s->level < 2) ? 4 : 0));
            pos = 10;
        }
        s->status = 1;
    }
    if (s->fill == 0 && kind != 2) {
        /* nothing gathered: the flush is just the empty stored block (every chunk ends on a byte boundary) */
        static const U8 marker[5] = {0, 0, 0, 0xFF, 0xFF};
        memcpy(s->pend + pos, marker, 5);
        pos += 5;
    } else {
        zscgpu_deflate_params p;
        memset(&p, 0, sizeof(p));
        p.max_block_len = s->fill ? s->fill : 1u;
        p.level = s->level; p.strategy = s->strategy;
        p.wrap = (s->wrap == 2) ? 2 : 0;                  /* raw blocks; 2 also reports the chunk's crc32 */
        p.window_bits = s->wbits;
        p.part = 1 | (kind == 2 ? 0 : 2);
        p.hist_len = s->hist;
        zscgpu_result res;
        const U8 *src = s->hist_area + s->wsize - s->hist;
        const int rc = zscgpu_compress_host(e, s->pend + pos, s->pend_cap - pos - 8u, src, s->hist + s->fill, &p, 0, &res);
        if (rc != 0 || res.ret != Z_OK) {
            ZSC_WARN2("In deflate(), the GPU engine failed (%d / %d).", rc, rc == 0 ? res.ret : 0);
            return rc != 0 ? Z_MEM_ERROR : (ZlibReturn)res.ret;
        }
        pos += res.produced;
        if (s->fill) {
            s->check = (s->wrap == 2) ? zscgpu_crc32_combine(s->check, res.check, s->fill)
                                      : zscgpu_adler32_combine(s->check, res.check, s->fill);
        }
        s->total_len += s->fill;
        /* slide: the newest wsize bytes of history + chunk become the history of the next chunk */
        const U32 have = s->hist + s->fill, nh = have < s->wsize ? have : s->wsize;
        memmove(s->hist_area + s->wsize - nh, s->hist_area + s->wsize + s->fill - nh, nh);
        s->hist = nh; s->fill = 0;
    }
    if (kind == 1) s->hist = 0;
    if (kind == 2) {
        if (s->wrap == 1) {
            s->pend[pos++] = (U8)(s->check >> 24); s->pend[pos++] = (U8)(s->check >> 16); s->pend[pos++] = (U8)(s->check >> 8); s->pend[pos++] = (U8)s->check;
        } else if (s->wrap == 2) {
            for (I32 k = 0; k < 4; k++) s->pend[pos++] = (U8)(s->check >> (8 * k));
            for (I32 k = 0; k < 4; k++) s->pend[pos++] = (U8)(s->total_len >> (8 * k));
        }
        s->status = 2;
    }
    if (s->wrap != 0) strm->adler = s->check;
    s->pend_len = pos; s->pend_pos = 0;
    return Z_OK;
}

ZlibReturn deflate(z_stream *strm, ZlibFlush flush)
{
    struct internal_state *s = zst_state(strm, ZST_DEFLATE_MAGIC);
    if (s == Z_NULL || (I32)flush > Z_BLOCK || (I32)flush < 0) { ZSC_WARN("deflate: bad state or flush."); return Z_STREAM_ERROR; }
    if (strm->next_out == Z_NULL || (strm->avail_in != 0 && strm->next_in == Z_NULL) || (s->status == 2 && flush != Z_FINISH)) {
        strm->msg = (const U8 *)"stream error";
        return Z_STREAM_ERROR;
    }
    if (strm->avail_out == 0) { strm->msg = (const U8 *)"buffer error"; return Z_BUF_ERROR; }
    const I32 old_flush = s->last_flush;
    s->last_flush = (I32)flush;
    const U32 in0 = strm->avail_in, out0 = strm->avail_out;
    /* a repeated flush with nothing new is refused as the reference refuses it (src/deflate.c:1216-1224) */
    if (strm->avail_in == 0 && s->pend_len == 0 && s->fill == 0 && (I32)flush <= old_flush && flush != Z_FINISH && s->status == 1) {
        strm->msg = (const U8 *)"buffer error";
        return Z_BUF_ERROR;
    }
    for (;;) {
        zst_drain(strm, s);
        if (s->pend_len != 0) return Z_OK;                               /* the caller's buffer is full */
        if (s->status == 2) return Z_STREAM_END;
        U32 n = s->cap_new - s->fill;
        if (n > strm->avail_in) n = strm->avail_in;
        if (n) {
            memcpy(s->hist_area + s->wsize + s->fill, strm->next_in, n);
            strm->next_in += n; strm->avail_in -= n; strm->total_in += n;
            s->fill += n;
        }
        ZlibReturn r = Z_OK;
        if (s->fill == s->cap_new) r = zst_emit(strm, s, 0);
        else if (flush == Z_FINISH) r = zst_emit(strm, s, 2);
        else if (flush != Z_NO_FLUSH) {
            r = zst_emit(strm, s, flush == Z_FULL_FLUSH ? 1 : 0);
            if (r != Z_OK) return r;
            zst_drain(strm, s);
            return Z_OK;
        } else {
            if (in0 == strm->avail_in && out0 == strm->avail_out) { strm->msg = (const U8 *)"buffer error"; return Z_BUF_ERROR; }
            return Z_OK;
        }
        if (r != Z_OK) return r;
    }
}

ZlibReturn deflateEnd(z_stream *strm)
{
    struct internal_state *s = zst_state(strm, ZST_DEFLATE_MAGIC);
    if (s == Z_NULL) { ZSC_WARN("deflateEnd: bad state."); return Z_STREAM_ERROR; }
    const U32 status = s->status;
    const I32 busy = (status == 1) || s->fill != 0 || s->pend_len != 0;
    s->magic = 0;
    strm->state = Z_NULL;
    return busy ? Z_DATA_ERROR : Z_OK;
}

/* ================================================================== inflate */

ZlibReturn inflateReset2(z_stream *strm, I32 windowBits)
{
    struct internal_state *s = zst_state(strm, ZST_INFLATE_MAGIC);
    if (s == Z_NULL) { ZSC_WARN("inflateReset2: bad state."); return Z_STREAM_ERROR; }
    I32 wrap;
    if (windowBits < 0) { wrap = 0; windowBits = -windowBits; }
    else { wrap = (windowBits >> 4) + 5; if (windowBits < 48) windowBits &= 15; }
    if (windowBits && (windowBits < 8 || windowBits > 15)) { ZSC_WARN1("inflateReset2: bad window bits %d.", windowBits); return Z_STREAM_ERROR; }
    if (wrap & 2) { ZSC_WARN("inflateReset2: the gzip wrapper is served by zsc_uncompress_gzip*, not by the z_stream API of this engine."); return Z_STREAM_ERROR; }
    zscgpu_engine *e = zst_engine("inflateReset2()");
    if (e == Z_NULL) return Z_MEM_ERROR;
    s->iwrap = (wrap ? 1 : 0) | ((windowBits ? windowBits : 15) << 8);
    if (zscgpu_inflate_stream_reset(e, s->slot, s->iwrap) != 0) return Z_MEM_ERROR;
    s->wbits = windowBits ? windowBits : 15;
    s->out_cap = 1u << s->wbits;
    if (s->out_cap > ZSCGPU_STREAM_OUT_MAX) s->out_cap = ZSCGPU_STREAM_OUT_MAX;
    s->in_left = 0; s->mode = 0; s->pend_len = s->pend_pos = 0;
    s->check = 1u;
    strm->total_in = strm->total_out = 0;
    strm->msg = Z_NULL;
    strm->adler = (U32)(wrap & 1);
    return Z_OK;
}

ZlibReturn inflateReset(z_stream *strm)
{
    struct internal_state *s = zst_state(strm, ZST_INFLATE_MAGIC);
    if (s == Z_NULL) { ZSC_WARN("inflateReset: bad state."); return Z_STREAM_ERROR; }
    return inflateReset2(strm, (s->iwrap & 1) ? s->wbits : -s->wbits);
}

ZlibReturn inflateInit2_(z_stream *strm, I32 windowBits, const U8 *version, I32 stream_size)
{
    static const U8 my_version[] = ZLIB_VERSION;
    if (version == Z_NULL || version[0] != my_version[0] || stream_size != (I32)sizeof(z_stream)) {
        ZSC_WARN("inflateInit version error.");
        return Z_VERSION_ERROR;
    }
    if (strm == Z_NULL) { ZSC_WARN("inflateInit stream error: null stream."); return Z_STREAM_ERROR; }
    U32 work_size = U32_MAX;
    if (strm->next_work == Z_NULL || inflateWorkSize2(windowBits, &work_size) != Z_OK || strm->avail_work < work_size) {
        ZSC_WARN("inflateInit stream error: no or too small a work buffer.");
        return Z_STREAM_ERROR;
    }
    zscgpu_engine *e = zst_engine("inflateInit2_()");
    if (e == Z_NULL) return Z_MEM_ERROR;
    U8 *w = strm->next_work;
    U32 off = (U32)((8u - ((uintptr_t)w & 7u)) & 7u);
    struct internal_state *s = (struct internal_state *)(w + off);
    off += ((U32)sizeof(struct internal_state) + 7u) & ~7u;
    memset(s, 0, sizeof(*s));
    s->magic = ZST_INFLATE_MAGIC; s->strm = strm;
    s->pend = w + off;                                     /* the window area of the reference's layout: output not yet handed out */
    s->pend_cap = work_size - off;
    if (zscgpu_inflate_stream_open(e, 1 | (15 << 8), &s->slot) != 0) {
        ZSC_WARN1("inflateInit: %s", zscgpu_last_error(e));
        return Z_MEM_ERROR;
    }
    strm->state = s;
    const ZlibReturn r = inflateReset2(strm, windowBits);
    if (r != Z_OK) { (void)zscgpu_inflate_stream_close(e, s->slot); strm->state = Z_NULL; return r; }
    if (s->out_cap > s->pend_cap) s->out_cap = s->pend_cap & ~15u;
    strm->next_work += work_size; strm->avail_work -= work_size;
    return Z_OK;
}

ZlibReturn inflateInit_(z_stream *strm, const U8 *version, I32 stream_size)
{
    return inflateInit2_(strm, DEF_WBITS, version, stream_size);
}

ZlibReturn inflateEnd(z_stream *strm)
{
    struct internal_state *s = zst_state(strm, ZST_INFLATE_MAGIC);
    if (s == Z_NULL) { ZSC_WARN("inflateEnd: bad state."); return Z_STREAM_ERROR; }
    zscgpu_engine *e = zscgpu_global();
    if (e != Z_NULL) (void)zscgpu_inflate_stream_close(e, s->slot);
    s->magic = 0;
    strm->state = Z_NULL;
    return Z_OK;
}

ZlibReturn inflateSetDictionary(z_stream *strm, const U8 *dictionary, U32 dictLength)
{
    struct internal_state *s = zst_state(strm, ZST_INFLATE_MAGIC);
    if (s == Z_NULL || dictionary == Z_NULL) { ZSC_WARN("inflateSetDictionary: bad state or null dictionary."); return Z_STREAM_ERROR; }
    /* reference src/inflate.c:1446-1484: zlib streams only right after Z_NEED_DICT, raw streams at any time */
    if ((s->iwrap & 1) && s->mode != 3) return Z_STREAM_ERROR;
    zscgpu_engine *e = zst_engine("inflateSetDictionary()");
    if (e == Z_NULL) return Z_MEM_ERROR;
    if (s->mode == 3) {
        U32 id = 1;
        if (dictLength && zscgpu_checksum_host(e, 0, 1u, dictionary, dictLength, &id) != 0) return Z_MEM_ERROR;
        if (id != strm->adler) return Z_DATA_ERROR;
    }
    const U32 keep = dictLength < 32768u ? dictLength : 32768u;
    if (zscgpu_inflate_stream_set_dict(e, s->slot, dictionary + (dictLength - keep), keep) != 0) return Z_MEM_ERROR;
    if (s->mode == 3) s->mode = 0;
    return Z_OK;
}

ZSC_PRIVATE void zst_drain_out(z_stream *strm, struct internal_state *s)
{
    U32 n = s->pend_len - s->pend_pos;
    if (n > strm->avail_out) n = strm->avail_out;
    if (n) {
        memcpy(strm->next_out, s->pend + s->pend_pos, n);
        strm->next_out += n; strm->avail_out -= n; strm->total_out += n;
        s->pend_pos += n;
    }
    if (s->pend_pos == s->pend_len) s->pend_pos = s->pend_len = 0;
}

ZlibReturn inflate(z_stream *strm, ZlibFlush flush)
{
    struct internal_state *s = zst_state(strm, ZST_INFLATE_MAGIC);
    if (s == Z_NULL || strm->next_out == Z_NULL || (strm->next_in == Z_NULL && strm->avail_in != 0)) {
        ZSC_WARN("inflate: bad state or buffers.");
        return Z_STREAM_ERROR;
    }
    zscgpu_engine *e = zst_engine("inflate()");
    if (e == Z_NULL) return Z_MEM_ERROR;
    const U32 in0 = strm->avail_in, out0 = strm->avail_out;
    for (;;) {
        zst_drain_out(strm, s);
        if (s->pend_len != 0) break;                                      /* the caller's buffer is full */
        if (s->mode == 1) return Z_STREAM_END;
        if (s->mode == 2) { strm->msg = (const U8 *)"data error"; return Z_DATA_ERROR; }
        if (s->mode == 3) return Z_NEED_DICT;
        if (strm->avail_out == 0) break;
        U32 take = ZSCGPU_STREAM_IN_MAX - s->in_left;
        if (take > strm->avail_in) take = strm->avail_in;
        zscgpu_stream_step r;
        if (zscgpu_inflate_stream_step(e, s->slot, strm->next_in, take, s->in_left, s->pend, s->out_cap, 0, &r) != 0) {
            ZSC_WARN1("In inflate(), the GPU engine failed: %s", zscgpu_last_error(e));
            return Z_MEM_ERROR;
        }
        /* input: what the decoder left unread stays staged on the device only when it is the beginning of something it
           could not finish (more input needed); otherwise the unread bytes of THIS call are handed back to the caller, so
           that total_in is exact at the end of the stream and next_in points at what follows it */
        U32 left = s->in_left + take - r.in_pos;
        U32 back = 0;
        if (r.status != 0) { back = left < take ? left : take; left -= back; }
        strm->next_in += take - back; strm->avail_in -= take - back; strm->total_in += take - back;
        s->in_left = left;
        s->pend_len = r.produced; s->pend_pos = 0;
        if (r.produced) s->check = zscgpu_adler32_combine(s->check, r.adler, r.produced);
        if (s->iwrap & 1) strm->adler = s->check;
        if (r.status == 2) {
            if ((s->iwrap & 1) && r.have_check && r.stored_check != s->check) { s->mode = 2; strm->msg = (const U8 *)"incorrect data check"; }
            else s->mode = 1;
        } else if (r.status == 3) { s->mode = 2; strm->msg = (const U8 *)"invalid stream"; }
        else if (r.status == 4) { s->mode = 3; strm->adler = r.stored_check; }
        else if (r.status == 0 && r.produced == 0 && strm->avail_in == 0) break;      /* needs input the caller does not have */
        else if (r.status == 0 && r.produced == 0 && take == 0) break;                /* staging full of an unfinished item: cannot happen with 64 KiB */
    }
    if (s->pend_len == 0 && s->mode == 1) return Z_STREAM_END;
    if (in0 == strm->avail_in && out0 == strm->avail_out) {
        if (s->mode == 2) return Z_DATA_ERROR;
        if (s->mode == 3) return Z_NEED_DICT;
        strm->msg = (const U8 *)"buffer error";
        return Z_BUF_ERROR;
    }
    if (flush == Z_FINISH && s->mode == 0) return Z_BUF_ERROR;           /* reference src/inflate.c:1400-1402 */
    return Z_OK;
}

ZlibReturn inflateSync(z_stream *strm)
{
    struct internal_state *s = zst_state(strm, ZST_INFLATE_MAGIC);
    if (s == Z_NULL) { ZSC_WARN("In inflateSync(), bad state."); return Z_STREAM_ERROR; }
    if (strm->avail_in == 0 && s->in_left == 0) { ZSC_WARN("In inflateSync(), not enough input."); return Z_BUF_ERROR; }
    zscgpu_engine *e = zst_engine("inflateSync()");
    if (e == Z_NULL) return Z_MEM_ERROR;
    s->pend_len = s->pend_pos = 0;
    for (;;) {
        U32 take = ZSCGPU_STREAM_IN_MAX - s->in_left;
        if (take > strm->avail_in) take = strm->avail_in;
        zscgpu_stream_step r;
        if (zscgpu_inflate_stream_step(e, s->slot, strm->next_in, take, s->in_left, s->pend, 0, 1, &r) != 0) return Z_MEM_ERROR;
        U32 left = s->in_left + take - r.in_pos;
        U32 back = 0;
        if (r.status == 5) { back = left < take ? left : take; left -= back; }
        strm->next_in += take - back; strm->avail_in -= take - back; strm->total_in += take - back;
        s->in_left = left;
        if (r.status == 5) {
            /* restart on a new block with an empty window and a fresh check value (reference src/inflate.c:1590-1602) */
            s->mode = 0; s->check = 1u;
            return Z_OK;
        }
        if (strm->avail_in == 0) { ZSC_WARN("In inflateSync(), did not find 4 bytes."); return Z_DATA_ERROR; }
    }
}
