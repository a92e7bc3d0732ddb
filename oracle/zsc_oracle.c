/* zsc_oracle.c — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (libzsc_b200.so, zsc_b200/) includes, links or calls this file; it is used
 * by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline leg as the checker.
 *
 * What is restated (own words, plain C, every function cites the reference lines it follows):
 *   - adler32_z / crc32_z                         src/adler32.c:56-131, src/crc32.c:502-593
 *   - the eight buffer-size check functions       src/deflate.c:761-902, src/inflate.c:249-276,
 *                                                 src/zsc_compress.c:196-260, src/zsc_uncompr.c:32-42
 *   - zsc_uncompress[_gzip][2]: zlib/gzip header, block loop, dynamic-header decode, canonical
 *     Huffman decode, stored blocks, data check, inflateSync recovery
 *                                                 src/zsc_uncompr.c:44-154, src/inflate.c:704-1404,
 *                                                 :1523-1604, src/inftrees.c:60-358, src/inffast.c:76-314
 *   - zError / z_errmsg                           src/zutil.c:40-51,150
 * The decoder resolves codes bit by bit from per-length counts (the table layout of inftrees.c is an
 * optimisation, not part of the result); results, return codes and consumed/produced counts are
 * what the reference produces on the same input.
 *
 * What is NOT restated: deflate's match finder and tree builder.  Compressed bytes are not a
 * parity target (any valid stream within 2 % of the reference's size is correct), so for the
 * compress direction the checker is the reference itself: oracle/_ref/libzsc_ref.so, compiled by
 * oracle/Makefile from the sources under /root/reference (size comparison + its own inflate).
 *
 * Pinning: tests/test_oracle.py checks this file against (1) every raw-deflate known-answer vector
 * of reference test/infcover.c (tests/golden/infcover_vectors.json, harvested by
 * tests/golden/make_golden.py), (2) the hand-built bad headers of reference test/zlib_gtest.cpp
 * :1815-1918, (3) streams and checksums produced by oracle/_ref on seeded inputs
 * (tests/golden/ref_streams.json), (4) the pinned bound / work-size values of
 * reference test/output/Test.log:26-27,64,262.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef uint8_t U8;
typedef uint16_t U16;
typedef uint32_t U32;
typedef int32_t I32;

enum { Z_OK = 0, Z_STREAM_END = 1, Z_NEED_DICT = 2, Z_STREAM_ERROR = -2, Z_DATA_ERROR = -3, Z_MEM_ERROR = -4, Z_BUF_ERROR = -5 };

typedef struct {
    I32 text; U32 time; I32 xflags; I32 os; U8 *extra; U32 extra_len; U32 extra_max;
    U8 *name; U32 name_max; U8 *comment; U32 comm_max; I32 hcrc; I32 done;
} gz_header;

/* sizeof(deflate_state), sizeof(inflate_state) in the reference build (LP64); tests compare them with
 * refprobe_sizeof_* of oracle/_ref */
#define ORA_DEFLATE_STATE 5920u
#define ORA_INFLATE_STATE 7152u

/* ------------------------------------------------------------------ checksums */

/* src/adler32.c:56-131: s1 += byte, s2 += s1, both mod 65521; the reference defers the modulo to
 * every 5552 bytes (the largest n with 255 n (n+1)/2 + (n+1)(65520) < 2^32); same arithmetic here. */
U32 adler32_z(U32 adler, const U8 *buf, size_t len)
{
    U32 s1 = adler & 0xffff, s2 = (adler >> 16) & 0xffff;
    if (buf == NULL) return 1;                       /* src/adler32.c:82-84 */
    while (len > 0) {
        size_t n = len < 5552 ? len : 5552;
        len -= n;
        while (n--) { s1 += *buf++; s2 += s1; }
        s1 %= 65521u; s2 %= 65521u;
    }
    return (s2 << 16) | s1;
}
U32 adler32(U32 adler, const U8 *buf, U32 len) { return adler32_z(adler, buf, len); }

/* src/crc32.c:502-593: reflected CRC-32, polynomial 0xEDB88320, register pre- and post-inverted.
 * The reference uses eight hard-coded 256-entry tables; the table is computed here (same values
 * as crc_table[0], src/crc32.c:58-110) and the bytes are folded one at a time. */
static U32 ora_crc_table[256];
static int ora_crc_ready;
static void ora_crc_init(void)
{
    for (U32 n = 0; n < 256; n++) {
        U32 c = n;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        ora_crc_table[n] = c;
    }
    ora_crc_ready = 1;
}
U32 crc32_z(U32 crc, const U8 *buf, size_t len)
{
    if (buf == NULL) return 0;                       /* src/crc32.c:507 */
    if (!ora_crc_ready) ora_crc_init();
    crc ^= 0xffffffffu;
    while (len--) crc = ora_crc_table[(crc ^ *buf++) & 0xff] ^ (crc >> 8);
    return crc ^ 0xffffffffu;
}
U32 crc32(U32 crc, const U8 *buf, U32 len) { return crc32_z(crc, buf, len); }

/* ------------------------------------------------------------------ size checks */

/* src/deflate.c:857-902 (deflateWorkSize2) */
I32 zsc_compress_get_min_work_buf_size2(I32 window_bits, I32 mem_level, U32 *size_out)
{
    *size_out = 0xFFFFFFFFu;
    if (window_bits < 0) window_bits = -window_bits;
    else if (window_bits > 15) window_bits -= 16;
    if (window_bits == 8) window_bits = 9;
    if (mem_level < 1 || mem_level > 9 || window_bits < 8 || window_bits > 15) return Z_STREAM_ERROR;
    U32 w = 1u << window_bits;
    *size_out = ORA_DEFLATE_STATE + w * 2 + w * 2 * 2 + (1u << (mem_level + 7)) * 2 + (1u << (mem_level + 6)) * 4;
    return Z_OK;
}
I32 zsc_compress_get_min_work_buf_size(U32 *size_out) { return zsc_compress_get_min_work_buf_size2(15, 8, size_out); }

static U32 ora_cstrlen(const U8 *s) { U32 n = 1; while (*s++) n++; return n; }

/* src/deflate.c:761-849 (deflateBoundNoStream) */
static I32 ora_bound(U32 n, I32 level, I32 wbits, I32 mem_level, const gz_header *gh, U32 *out)
{
    *out = 0xFFFFFFFFu;
    I32 wrap = 1;
    if (wbits < 0) { wrap = 0; wbits = -wbits; }
    else if (wbits > 15) { wrap = 2; wbits -= 16; }
    if (mem_level < 1 || mem_level > 9 || wbits < 8 || wbits > 15 || (wbits == 8 && wrap != 1)) return Z_STREAM_ERROR;
    U32 wraplen = wrap == 0 ? 0 : wrap == 1 ? 10 : 18;
    if (wrap == 2 && gh != NULL) {
        if (gh->extra) wraplen += 2 + gh->extra_len;
        if (gh->name) wraplen += ora_cstrlen(gh->name);
        if (gh->comment) wraplen += ora_cstrlen(gh->comment);
        if (gh->hcrc) wraplen += 2;
    }
    if (wbits != 15 || mem_level != 8 || level == 0) *out = n + ((n + 7) >> 3) + ((n + 63) >> 6) + 5 + wraplen;
    else *out = n + (n >> 12) + (n >> 14) + (n >> 25) + 13 - 6 + wraplen;
    return Z_OK;
}
/* src/zsc_compress.c:207-236: bound, then again with 4 extra bytes per section */
I32 zsc_compress_get_max_output_size_gzip2(U32 n, U32 mbl, I32 level, I32 wbits, I32 mem_level, gz_header *gh, U32 *out)
{
    U32 first;
    I32 e = ora_bound(n, level, wbits, mem_level, gh, &first);
    if (e != Z_OK) return e;
    return ora_bound(n + (first / mbl + 1) * 4, level, wbits, mem_level, gh, out);
}
I32 zsc_compress_get_max_output_size2(U32 n, U32 mbl, I32 level, I32 wbits, I32 mem_level, U32 *out)
{ return zsc_compress_get_max_output_size_gzip2(n, mbl, level, wbits, mem_level, NULL, out); }
I32 zsc_compress_get_max_output_size_gzip(U32 n, U32 mbl, I32 level, gz_header *gh, U32 *out)
{ return zsc_compress_get_max_output_size_gzip2(n, mbl, level, 15 + 16, 8, gh, out); }
I32 zsc_compress_get_max_output_size(U32 n, U32 mbl, I32 level, U32 *out)
{ return zsc_compress_get_max_output_size2(n, mbl, level, 15, 8, out); }

/* src/inflate.c:249-276 (inflateWorkSize2) */
I32 zsc_uncompress_get_min_work_buf_size2(I32 wbits, U32 *out)
{
    if (wbits < 0) wbits = -wbits;
    else if (wbits < 48) wbits &= 15;
    if (wbits && (wbits < 8 || wbits > 15)) return Z_STREAM_ERROR;
    *out = ORA_INFLATE_STATE + (1u << wbits);
    return Z_OK;
}
I32 zsc_uncompress_get_min_work_buf_size(U32 *out) { return zsc_uncompress_get_min_work_buf_size2(15, out); }

/* ------------------------------------------------------------------ inflate */

typedef struct {
    const U8 *in; U32 in_len, in_pos;       /* next byte to pull */
    U32 hold; int bits;                     /* bit accumulator, LSB first (src/inflate.c NEEDBITS/DROPBITS :640-690) */
    U8 *out; U32 out_cap, out_pos, out_base;   /* out_base: output position of the last resynchronisation */
    U32 dmax;                               /* largest legal distance (window size from the header, :774-779) */
    const char *msg;
    int short_in, short_out;                /* why decoding stopped, when it did not finish */
} ora_strm;

/* pull n <= 16 bits; sets short_in when the input ends first */
static int ora_need(ora_strm *s, int n)
{
    while (s->bits < n) {
        if (s->in_pos >= s->in_len) { s->short_in = 1; return 0; }
        s->hold |= (U32)s->in[s->in_pos++] << s->bits;
        s->bits += 8;
    }
    return 1;
}
static U32 ora_bitsv(ora_strm *s, int n) { U32 v = s->hold & ((1u << n) - 1); s->hold >>= n; s->bits -= n; return v; }

typedef struct { U16 count[16]; U16 symbol[288]; } ora_code;

/* src/inftrees.c:107-177: count lengths, reject over-subscribed sets, and incomplete ones unless the
 * set is a single one-bit code of a distance/length alphabet.  Returns 0 ok, -1 invalid. */
static int ora_build(ora_code *h, const U8 *lens, int n, int is_codelens)
{
    U16 offs[16];
    int left = 1, max = 0;
    memset(h->count, 0, sizeof(h->count));
    for (int i = 0; i < n; i++) h->count[lens[i]]++;
    for (int l = 15; l >= 1; l--) if (h->count[l]) { max = l; break; }
    if (max == 0) { h->count[0] = 0; return is_codelens ? -1 : 1; }      /* no codes at all (:128-138) */
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= h->count[l]; if (left < 0) return -1; }
    if (left > 0 && (is_codelens || max != 1)) return -1;
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + h->count[l];
    for (int i = 0; i < n; i++) if (lens[i]) h->symbol[offs[lens[i]]++] = (U16)i;
    h->count[0] = 0;
    return 0;
}

/* canonical decode, one bit at a time: returns symbol, -1 invalid code, -2 out of input */
static int ora_decode(ora_strm *s, const ora_code *h)
{
    int code = 0, first = 0, index = 0;
    U32 save_pos = s->in_pos, save_hold = s->hold; int save_bits = s->bits;
    for (int len = 1; len <= 15; len++) {
        if (!ora_need(s, 1)) return -2;
        code |= (int)ora_bitsv(s, 1);
        int count = h->count[len];
        if (code - count < first) return h->symbol[index + (code - first)];
        index += count; first += count; first <<= 1; code <<= 1;
    }
    /* An unassigned code can only exist in a one-bit incomplete set (src/inftrees.c:168-177); the
       reference's table entry for it consumes that one bit (src/inftrees.c:329-349), no more. */
    s->in_pos = save_pos; s->hold = save_hold; s->bits = save_bits;
    if (ora_need(s, 1)) (void)ora_bitsv(s, 1);
    return -1;
}

static const U16 ora_lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const U8 ora_lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const U16 ora_dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const U8 ora_dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

/* LEN / MATCH states, src/inflate.c:1182-1321 and src/inffast.c:105-300.  0 end of block, <0 error */
static int ora_codes(ora_strm *s, const ora_code *lc, const ora_code *dc)
{
    for (;;) {
        int sym = ora_decode(s, lc);
        if (sym == -2) return -2;
        if (sym < 0) { s->msg = "invalid literal/length code"; return -1; }
        if (sym < 256) {
            if (s->out_pos >= s->out_cap) { s->short_out = 1; return -2; }
            s->out[s->out_pos++] = (U8)sym;
        } else if (sym == 256) return 0;
        else {
            sym -= 257;
            if (sym >= 29) { s->msg = "invalid literal/length code"; return -1; }
            if (!ora_need(s, ora_lext[sym])) return -2;
            U32 len = ora_lbase[sym] + ora_bitsv(s, ora_lext[sym]);
            int d = ora_decode(s, dc);
            if (d == -2) return -2;
            if (d < 0 || d >= 30) { s->msg = "invalid distance code"; return -1; }
            if (!ora_need(s, ora_dext[d])) return -2;
            U32 dist = ora_dbase[d] + ora_bitsv(s, ora_dext[d]);
            if (dist > s->dmax || dist > s->out_pos - s->out_base) { s->msg = "invalid distance too far back"; return -1; }   /* inffast.c:184-189, inflate.c:1280-1286 */
            while (len--) {
                if (s->out_pos >= s->out_cap) { s->short_out = 1; return -2; }
                s->out[s->out_pos] = s->out[s->out_pos - dist];
                s->out_pos++;
            }
        }
    }
}

/* one deflate block sequence until the final block; 0 done, -1 data error, -2 needs input/output */
static int ora_blocks(ora_strm *s)
{
    static const U8 order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    ora_code lc, dc;
    U8 lens[320];
    for (;;) {
        if (!ora_need(s, 3)) return -2;
        int last = (int)ora_bitsv(s, 1), type = (int)ora_bitsv(s, 2);       /* src/inflate.c:975-1009 */
        if (type == 0) {                                                   /* stored, :1010-1049 */
            ora_bitsv(s, s->bits & 7);
            if (!ora_need(s, 16)) return -2;
            U32 len = ora_bitsv(s, 16);
            if (!ora_need(s, 16)) return -2;
            U32 nlen = ora_bitsv(s, 16);
            if (len != (nlen ^ 0xffff)) { s->msg = "invalid stored block lengths"; return -1; }
            /* whole bytes still in the accumulator go back to the input */
            s->in_pos -= (U32)(s->bits >> 3); s->bits = 0; s->hold = 0;
            while (len--) {
                if (s->in_pos >= s->in_len) { s->short_in = 1; return -2; }
                if (s->out_pos >= s->out_cap) { s->short_out = 1; return -2; }
                s->out[s->out_pos++] = s->in[s->in_pos++];
            }
        } else if (type == 1) {                                            /* fixed, :122-206, :989-999 */
            int i = 0;
            for (; i < 144; i++) lens[i] = 8;
            for (; i < 256; i++) lens[i] = 9;
            for (; i < 280; i++) lens[i] = 7;
            for (; i < 288; i++) lens[i] = 8;
            ora_build(&lc, lens, 288, 0);
            for (i = 0; i < 32; i++) lens[i] = 5;
            ora_build(&dc, lens, 32, 0);
            int r = ora_codes(s, &lc, &dc);
            if (r) return r;
        } else if (type == 2) {                                            /* dynamic, :1050-1178 */
            if (!ora_need(s, 14)) return -2;
            int nlen = (int)ora_bitsv(s, 5) + 257, ndist = (int)ora_bitsv(s, 5) + 1, ncode = (int)ora_bitsv(s, 4) + 4;
            if (nlen > 286 || ndist > 30) { s->msg = "too many length or distance symbols"; return -1; }
            memset(lens, 0, 19);
            for (int i = 0; i < ncode; i++) { if (!ora_need(s, 3)) return -2; lens[order[i]] = (U8)ora_bitsv(s, 3); }
            ora_code cl;
            if (ora_build(&cl, lens, 19, 1)) { s->msg = "invalid code lengths set"; return -1; }
            int idx = 0;
            while (idx < nlen + ndist) {
                int sym = ora_decode(s, &cl);
                if (sym == -2) return -2;
                if (sym < 0) { s->msg = "invalid code lengths set"; return -1; }
                if (sym < 16) lens[idx++] = (U8)sym;
                else {
                    int rep, v = 0;
                    if (sym == 16) {
                        if (idx == 0) { s->msg = "invalid bit length repeat"; return -1; }
                        v = lens[idx - 1];
                        if (!ora_need(s, 2)) return -2;
                        rep = 3 + (int)ora_bitsv(s, 2);
                    } else if (sym == 17) { if (!ora_need(s, 3)) return -2; rep = 3 + (int)ora_bitsv(s, 3); }
                    else { if (!ora_need(s, 7)) return -2; rep = 11 + (int)ora_bitsv(s, 7); }
                    if (idx + rep > nlen + ndist) { s->msg = "invalid bit length repeat"; return -1; }
                    while (rep--) lens[idx++] = (U8)v;
                }
            }
            if (lens[256] == 0) { s->msg = "invalid code -- missing end-of-block"; return -1; }
            if (ora_build(&lc, lens, nlen, 0) < 0) { s->msg = "invalid literal/lengths set"; return -1; }
            if (ora_build(&dc, lens + nlen, ndist, 0) < 0) { s->msg = "invalid distances set"; return -1; }
            int r = ora_codes(s, &lc, &dc);
            if (r) return r;
        } else { s->msg = "invalid block type"; return -1; }
        if (last) return 0;
    }
}

/* src/inflate.c:1523-1545 (syncsearch): find 00 00 FF FF starting at byte `from` */
static U32 ora_sync(const U8 *in, U32 n, U32 from)
{
    U32 got = 0;
    for (U32 p = from; p < n; p++) {
        U32 b = in[p];
        if (b == (got < 2 ? 0u : 0xffu)) got++;
        else if (b) got = 0;
        else got = 4 - got;
        if (got == 4) return p + 1;
    }
    return n + 1;          /* not found */
}

/* gzip member header, src/inflate.c:786-954, parsed in the order the reference pulls it.
 * >0 header length; 0 input ended inside the header; -1 malformed: *pulled = bytes the reference has
 * pulled when it reports the error, *held = how many of them still sit un-dropped in its accumulator */
static I32 ora_gzip_header(const U8 *s, U32 n, gz_header *gh, const char **msg, U32 *pulled, U32 *held)
{
    *held = 2;
    if (n < 4) return 0;
    *pulled = 4;
    if (s[2] != 8) { *msg = "unknown compression method"; return -1; }
    U32 flg = s[3], p = 10;
    if (flg & 0xe0) { *msg = "unknown header flags set"; return -1; }
    if (n < 10) return 0;
    if (gh) { gh->text = (I32)(flg & 1); gh->time = s[4] | (s[5] << 8) | (s[6] << 16) | ((U32)s[7] << 24); gh->xflags = s[8]; gh->os = s[9]; }
    if (flg & 4) {
        if (p + 2 > n) return 0;
        U32 xl = s[p] | (s[p + 1] << 8); p += 2;
        if (p + xl > n) return 0;
        if (gh) { gh->extra_len = xl; if (gh->extra) memcpy(gh->extra, s + p, xl < gh->extra_max ? xl : gh->extra_max); }
        p += xl;
    } else if (gh) gh->extra = NULL;
    for (int f = 8; f <= 16; f <<= 1) {
        U8 *dst = !gh ? NULL : (f == 8 ? gh->name : gh->comment);
        U32 max = !gh ? 0 : (f == 8 ? gh->name_max : gh->comm_max), k = 0;
        if (flg & f) {
            for (;;) { if (p >= n) return 0; U8 c = s[p++]; if (dst && k < max) dst[k++] = c; if (!c) break; }
        } else if (gh) { if (f == 8) gh->name = NULL; else gh->comment = NULL; }
    }
    if (flg & 2) {
        if (p + 2 > n) return 0;
        *pulled = p + 2;
        if ((crc32_z(0, s, p) & 0xffff) != (U32)(s[p] | (s[p + 1] << 8))) { *msg = "header crc mismatch"; return -1; }
        p += 2;
    }
    if (gh) { gh->hcrc = (I32)((flg >> 1) & 1); gh->done = 1; }
    return (I32)p;
}

const char *ora_last_msg;       /* the reference's strm->msg of the last data error (test hook) */

/* src/zsc_uncompr.c:44-154 with everything beneath it */
I32 zsc_uncompress_gzip2(U8 *dest, U32 *dest_len, const U8 *source, U32 *source_len,
                         U8 *work, U32 work_len, I32 window_bits, gz_header *gz_head)
{
    U32 dcap = *dest_len, slen = *source_len, need = 0;
    (void)work;
    *dest_len = 0; *source_len = 0;
    ora_last_msg = NULL;
    I32 e = zsc_uncompress_get_min_work_buf_size2(window_bits, &need);
    if (e != Z_OK) return e;
    if (work_len < need) return Z_MEM_ERROR;
    I32 wrap, wbits;
    if (window_bits < 0) { wrap = 0; wbits = -window_bits; }
    else { wrap = (window_bits >> 4) + 5; wbits = window_bits < 48 ? (window_bits & 15) : window_bits; }
    if (gz_head != NULL) { if ((wrap & 2) == 0) return Z_STREAM_ERROR; gz_head->done = 0; }

    ora_strm s;
    memset(&s, 0, sizeof(s));
    s.in = source; s.in_len = slen; s.out = dest; s.out_cap = dcap;
    s.dmax = wbits ? (1u << wbits) : 32768u;
    int gzip = 0, errors = 0, rc = 0;
    U32 held = 0;                /* bytes the reference still holds un-dropped when it reports the current error */
    U32 check_from = 0;          /* output position the running data check restarts from after a sync */
    /* HEAD state, src/inflate.c:740-785: the first 16 bits decide gzip / zlib */
    if (wrap) {
        if (slen < 2) { *source_len = slen; return Z_BUF_ERROR; }
        s.in_pos = 2; held = 2;
        if ((wrap & 2) && source[0] == 31 && source[1] == 139) {
            const char *m = NULL; U32 pulled = 2;
            I32 hl = ora_gzip_header(source, slen, gz_head, &m, &pulled, &held);
            gzip = 1;
            if (hl == 0) { *source_len = slen; return Z_BUF_ERROR; }
            if (hl < 0) { s.msg = m; rc = -1; s.in_pos = pulled; }
            else { s.in_pos = (U32)hl; held = 0; }
        } else {
            U32 cmf = source[0], flg = source[1];
            if (gz_head) gz_head->done = -1;
            if (!(wrap & 1) || ((cmf << 8) + flg) % 31) { s.msg = "incorrect header check"; rc = -1; }
            else if ((cmf & 15) != 8) { s.msg = "unknown compression method"; rc = -1; }
            else if ((cmf >> 4) + 8 > (U32)(wbits ? wbits : 15)) { s.msg = "invalid window size"; rc = -1; }
            else if (flg & 0x20) { return Z_NEED_DICT; }     /* total_in is not updated on this return (:970-973) */
            else { s.dmax = 1u << ((cmf >> 4) + 8); held = 0; }
        }
    }
    /* the loop of src/zsc_uncompr.c:103-127: inflate; on a data error look for the next flush point */
    int terminal = 0;
    for (;;) {
        if (rc == 0) {
            rc = ora_blocks(&s);
            held = (rc == -1 && s.msg && strcmp(s.msg, "invalid stored block lengths") == 0) ? 4 : 0;
        }
        if (rc == 0) {
            /* CHECK / LENGTH states, src/inflate.c:1322-1354 */
            ora_bitsv(&s, s.bits & 7);
            s.in_pos -= (U32)(s.bits >> 3); s.bits = 0; s.hold = 0;
            if (gzip) {
                const U8 *t = s.in + s.in_pos;
                if (s.in_len - s.in_pos < 4) { s.in_pos = s.in_len; s.short_in = 1; rc = -2; }
                else if ((t[0] | (t[1] << 8) | (t[2] << 16) | ((U32)t[3] << 24)) != crc32_z(0, s.out + check_from, s.out_pos - check_from)) {
                    s.msg = "incorrect data check"; rc = -1; s.in_pos += 4; held = 4;
                } else if (s.in_len - s.in_pos < 8) { s.in_pos = s.in_len; s.short_in = 1; rc = -2; }
                else if ((t[4] | (t[5] << 8) | (t[6] << 16) | ((U32)t[7] << 24)) != s.out_pos) {
                    s.msg = "incorrect length check"; rc = -1; s.in_pos += 8; held = 4;
                } else s.in_pos += 8;
            } else if (wrap) {
                if (s.in_len - s.in_pos < 4) { s.in_pos = s.in_len; s.short_in = 1; rc = -2; }
                else {
                    const U8 *t = s.in + s.in_pos;
                    U32 want = ((U32)t[0] << 24) | (t[1] << 16) | (t[2] << 8) | t[3];
                    s.in_pos += 4;
                    if (want != adler32_z(1, s.out + check_from, s.out_pos - check_from)) { s.msg = "incorrect data check"; rc = -1; held = 4; }
                }
            }
        }
        if (rc != -1) break;
        errors++;
        ora_last_msg = s.msg;
        /* inflateSync, src/inflate.c:1547-1604 */
        U32 pulled = s.in_pos - (U32)(s.bits >> 3);
        s.bits = 0; s.hold = 0;
        if (held == 0 && pulled >= s.in_len) { s.in_pos = s.in_len; terminal = Z_BUF_ERROR; break; }   /* :1556 */
        U32 nx = ora_sync(s.in, s.in_len, pulled - held);
        held = 0;
        if (nx > s.in_len) { s.in_pos = s.in_len; terminal = Z_DATA_ERROR; break; }                    /* :1587 */
        s.in_pos = nx; rc = 0; check_from = s.out_pos;
        s.out_base = s.out_pos; s.dmax = 32768u;          /* inflateReset inside inflateSync: whave = 0, dmax = 32768 (src/inflate.c:301,324) */
    }
    *dest_len = s.out_pos;
    *source_len = s.in_pos - (U32)(s.bits >> 3);
    if (terminal) return terminal;
    if (rc == -2) return Z_BUF_ERROR;
    if (errors) return Z_DATA_ERROR;            /* src/zsc_uncompr.c:149-152 */
    return Z_OK;
}
I32 zsc_uncompress2(U8 *d, U32 *dl, const U8 *s, U32 *sl, U8 *w, U32 wl, I32 wbits) { return zsc_uncompress_gzip2(d, dl, s, sl, w, wl, wbits, NULL); }
I32 zsc_uncompress(U8 *d, U32 *dl, const U8 *s, U32 *sl, U8 *w, U32 wl) { return zsc_uncompress2(d, dl, s, sl, w, wl, 15); }
I32 zsc_uncompress_gzip(U8 *d, U32 *dl, const U8 *s, U32 *sl, U8 *w, U32 wl, gz_header *gh) { return zsc_uncompress_gzip2(d, dl, s, sl, w, wl, 15 + 16, gh); }

/* raw-deflate entry for the known-answer vectors of reference test/infcover.c (inflateInit2(-15) there) */
I32 ora_inflate_raw(const U8 *in, U32 in_len, U8 *out, U32 out_cap, U32 *produced, const char **msg)
{
    ora_strm s;
    memset(&s, 0, sizeof(s));
    s.in = in; s.in_len = in_len; s.out = out; s.out_cap = out_cap; s.dmax = 32768u;
    int rc = ora_blocks(&s);
    *produced = s.out_pos;
    *msg = s.msg;
    return rc == 0 ? Z_STREAM_END : rc == -1 ? Z_DATA_ERROR : Z_BUF_ERROR;
}

/* src/zutil.c:40-51 */
const char *zError(I32 err)
{
    static const char *const m[10] = {"need dictionary", "stream end", "", "file error", "stream error", "data error",
                                      "insufficient memory", "buffer error", "incompatible version", ""};
    I32 i = 2 - err;
    return m[(i < 0 || i > 9) ? 9 : i];
}
const char *zlibVersion(void) { return "1.2.11.f-abcouwer-safety-critical-v0 (oracle restatement)"; }
