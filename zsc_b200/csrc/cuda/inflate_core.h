/* inflate_core.h — one DEFLATE/zlib stream decoded by one thread (host + device).
 *
 * Covers what zsc_uncompress needs of the reference's inflate (src/inflate.c:704-1404: zlib header
 * :740-785, block types :975-1009, stored :1010-1049, dynamic header :1050-1178, data check
 * :1322-1342), inflate_table (src/inftrees.c:60-358), inflate_fast (src/inffast.c:76-314) and the
 * corruption recovery of zsc_uncompress_gzip2 (src/zsc_uncompr.c:103-127 with inflateSync,
 * src/inflate.c:1523-1604) — for the case zsc_uncompress always has: all input and all output
 * present, so there is no resumable state machine and no sliding window (matches are copied
 * straight out of the output buffer, as the reference itself does on that path, src/inflate.c:1380).
 *
 * The decode tables are NOT the reference's op/bits/val two-level tables: each alphabet gets one
 * direct-lookup table of 16-bit entries (10 / 8 index bits) plus the canonical-code arrays
 * (symbols sorted by length, per-length counts) that resolve the rare longer codes by
 * first-code comparison.  That keeps a stream's tables at 3.6 KiB, so that every warp of an SM can hold its
 * stream's tables in shared memory.
 *
 * Error behaviour follows the reference: any malformed header, code set, code, distance or
 * data check yields Z_DATA_ERROR (-3); running out of output room or input yields Z_BUF_ERROR (-5).
 * The functions are `__host__ __device__` so tests/ can run the GPU's exact decode logic on the
 * CPU against the reference's known-answer vectors.
 */
/* (inflate.cu includes this file twice, in two namespaces with two table geometries: ZI_REINCLUDE lifts the guard) */
#if !defined(ZSC_INFLATE_CORE_H) || defined(ZI_REINCLUDE)
#define ZSC_INFLATE_CORE_H

#include <stdint.h>

#ifdef __CUDACC__
#define ZID static __host__ __device__ __forceinline__
#else
#define ZID static inline
#endif

/* Table geometry.  The default (9 / 6 root bits, 192 second-level entries: 1.6 KB per stream) is the one wide batches
 * use, where shared memory decides how many streams an SM holds; narrow batches (one warp per stream, latency bound)
 * use 10 / 8 root bits (inflate.cu).  Results do not depend on the geometry. */
#ifndef ZI_LBITS
#define ZI_LBITS 9                   /* root bits of the literal/length table */
#define ZI_DBITS 6                   /* root bits of the distance table */
#define ZI_POOL 192                  /* second-level entries, shared by both alphabets of a block */
#endif

#define ZI_OK 0
#define ZI_NEED_DICT 2
#define ZI_DATA_ERROR (-3)
#define ZI_BUF_ERROR (-5)

/* fine-grained reason, mirrors the reference's strm->msg strings (src/inflate.c, src/inffast.c) */
enum {
    ZI_E_NONE = 0, ZI_E_HEADER_CHECK, ZI_E_METHOD, ZI_E_WINDOW, ZI_E_BLOCK_TYPE, ZI_E_STORED_LEN,
    ZI_E_TOO_MANY_SYMS, ZI_E_CODELEN_SET, ZI_E_BITLEN_REPEAT, ZI_E_NO_EOB, ZI_E_LITLEN_SET,
    ZI_E_DIST_SET, ZI_E_LITLEN_CODE, ZI_E_DIST_CODE, ZI_E_DIST_TOO_FAR, ZI_E_DATA_CHECK,
    ZI_E_INPUT_END, ZI_E_OUTPUT_FULL, ZI_E_NEED_DICT
};

/* Table entries (both alphabets): 0 = not resolved here (the canonical walk decides); bit 15 clear: sym | len << 9, a
 * whole code of len <= 15 bits; bit 15 set: a link, pool index | sub-table bits << 10 — the next bits of the input index a
 * second-level table in `pool` whose entries have the first form.  Sub-tables are handed out while the pool lasts; what
 * does not fit (or is wider than 7 bits) stays 0 and goes through the walk, so the pool can be small. */
typedef struct {                     /* hot: 1624 bytes per stream */
    uint16_t lit[1 << ZI_LBITS];
    uint16_t dist[1 << ZI_DBITS];
    uint16_t pool[ZI_POOL];
    uint16_t lcount[16];
    uint16_t dcount[16];
    uint32_t pool_used;
    uint32_t lfirst, lindex, dfirst, dindex;   /* where zi_decode's canonical walk stands after the lengths the direct
                                                  tables resolve (set with the tables; read by zi_fast_batch) */
} zi_tables;

typedef struct {                     /* cold: only the rare codes longer than the direct tables read these
                                        (global memory on the GPU) */
    uint16_t lsorted[288];
    uint16_t dsorted[32];
} zi_aux;

typedef struct {
    const uint8_t *in;      /* stream start */
    uint32_t in_len;
    uint32_t ip;            /* next input byte to load */
    uint64_t hold;
    uint32_t bits;
    uint8_t *out;
    uint32_t out_cap;
    uint32_t op;
    uint32_t pre;           /* device: the input word at [ip, ip + 4), loaded one refill ahead */
    uint32_t pv;            /* pre is valid */
} zi_io;

typedef struct {
    int32_t ret;            /* ZI_OK / ZI_DATA_ERROR / ZI_BUF_ERROR / ZI_NEED_DICT */
    int32_t reason;         /* first ZI_E_* seen */
    uint32_t produced;
    uint32_t consumed;
    uint32_t data_errors;   /* corrupted sections skipped */
    uint32_t stored_check;  /* adler32 from the trailer */
    uint32_t have_check;    /* trailer was reached */
    int32_t last_reason;    /* ZI_E_* of the most recent failure */
    uint32_t at_flush;      /* section mode: decoding stopped behind an empty non-final stored block (a flush point) */
} zi_result;

ZID void zi_refill(zi_io *io)
{
#ifdef __CUDA_ARCH__
    /* 4-byte aligned loads once the cursor is aligned; the arenas are padded, and bits past in_len
       are never consumed (zi_overrun is checked before anything derived from them is used) */
    if (io->bits <= 32) {
        if (io->pv) {
            /* the word was requested one refill ago; ask for the next one now */
            io->hold |= (uint64_t)io->pre << io->bits;
            io->bits += 32; io->ip += 4;
            io->pre = __ldg(reinterpret_cast<const uint32_t *>(io->in + io->ip));
        } else {
            const uint8_t *p = io->in + io->ip;
            while ((((uintptr_t)p) & 3) != 0 && io->bits <= 56) { io->hold |= (uint64_t)(*p++) << io->bits; io->bits += 8; io->ip++; }
            if (io->bits <= 32) {
                io->hold |= (uint64_t)__ldg(reinterpret_cast<const uint32_t *>(p)) << io->bits;
                io->bits += 32; io->ip += 4; p += 4;
            }
            if ((((uintptr_t)p) & 3) == 0) { io->pre = __ldg(reinterpret_cast<const uint32_t *>(p)); io->pv = 1; }
        }
    }
#else
    while (io->bits <= 56) {
        uint8_t b = io->ip < io->in_len ? io->in[io->ip] : 0;
        io->hold |= (uint64_t)b << io->bits; io->bits += 8; io->ip++;
    }
#endif
}
/* true when more bits were consumed than the input holds */
ZID int zi_overrun(const zi_io *io) { return (uint64_t)io->ip * 8 - io->bits > (uint64_t)io->in_len * 8; }
ZID uint32_t zi_peek(const zi_io *io, int n) { return (uint32_t)io->hold & ((1u << n) - 1u); }
ZID void zi_drop(zi_io *io, int n) { io->hold >>= n; io->bits -= (uint32_t)n; }
ZID uint32_t zi_take(zi_io *io, int n) { uint32_t v = zi_peek(io, n); zi_drop(io, n); return v; }
ZID uint32_t zi_take32(zi_io *io) { uint32_t v = (uint32_t)io->hold; zi_drop(io, 32); return v; }
/* move the cursor to byte `pos`, discarding buffered bits */
ZID void zi_seek(zi_io *io, uint32_t pos) { io->hold = 0; io->bits = 0; io->ip = pos; io->pv = 0; }
ZID uint32_t zi_consumed_bytes(const zi_io *io) { return (uint32_t)(((uint64_t)io->ip * 8 - io->bits + 7) >> 3); }

/* n bytes from a region that does not overlap the destination (stored blocks): eight loads in flight */
ZID void zi_copy_fwd(uint8_t *dst, const uint8_t *src, uint32_t n)
{
    uint32_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint8_t b0 = src[i], b1 = src[i + 1], b2 = src[i + 2], b3 = src[i + 3], b4 = src[i + 4], b5 = src[i + 5], b6 = src[i + 6], b7 = src[i + 7];
        dst[i] = b0; dst[i + 1] = b1; dst[i + 2] = b2; dst[i + 3] = b3; dst[i + 4] = b4; dst[i + 5] = b5; dst[i + 6] = b6; dst[i + 7] = b7;
    }
    for (; i < n; i++) dst[i] = src[i];
}

/* LZ77 copy of n bytes from `dist` back (may overlap forwards, as in the reference's inffast.c:259-272).
 * The loads of a step never depend on its stores: with dist >= 8 a step reads bytes written by earlier
 * steps only; shorter distances replicate the dist-byte pattern, read once. */
ZID void zi_copy_match(uint8_t *q, uint32_t dist, uint32_t n)
{
    const uint8_t *f = q - dist;
    if (dist >= 8) {
        uint32_t i = 0;
        for (; i + 8 <= n; i += 8) {
            uint8_t b0 = f[i], b1 = f[i + 1], b2 = f[i + 2], b3 = f[i + 3], b4 = f[i + 4], b5 = f[i + 5], b6 = f[i + 6], b7 = f[i + 7];
            q[i] = b0; q[i + 1] = b1; q[i + 2] = b2; q[i + 3] = b3; q[i + 4] = b4; q[i + 5] = b5; q[i + 6] = b6; q[i + 7] = b7;
        }
        uint8_t t[8];
        const uint32_t r = n - i;
        for (uint32_t k = 0; k < 8; k++) if (k < r) t[k] = f[i + k];
        for (uint32_t k = 0; k < 8; k++) if (k < r) q[i + k] = t[k];
    } else {
        uint64_t pat = 0;                       /* the dist-byte pattern in a register */
        for (uint32_t k = 0; k < 8; k++) if (k < dist) pat |= (uint64_t)f[k] << (8 * k);
        uint32_t sh = 0;
        const uint32_t wrap = 8 * dist;
        for (uint32_t i = 0; i < n; i++) { q[i] = (uint8_t)(pat >> sh); sh += 8; if (sh == wrap) sh = 0; }
    }
}

ZID uint32_t zi_rev(uint32_t v, int n)
{
#ifdef __CUDA_ARCH__
    return __brev(v) >> (32 - n);
#endif
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

ZID int zi_popc(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

/* Build one alphabet's tables from code lengths. kind 0: code-length/literal alphabets must be
 * complete; an incomplete set is tolerated only when its longest code is 1 bit and kind != 0
 * (same rule as the reference, src/inftrees.c:168-177).  Returns 0 or -1. */
ZID int zi_build(const uint8_t *lens, int n, int tbits, uint16_t *table, uint16_t *sorted,
                               uint16_t *count, int shift, int allow_incomplete)
{
    uint16_t offs[16];
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int i = 0; i < n; i++) count[lens[i]]++;
    for (int i = 0; i < (1 << tbits); i++) table[i] = 0;
    int maxl = 15;
    while (maxl > 0 && count[maxl] == 0) maxl--;
    if (maxl == 0) { count[0] = 0; return allow_incomplete ? 0 : -1; }   /* no codes: any use is an error */
    int left = 1;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= count[l]; if (left < 0) return -1; }
    if (left > 0 && (!allow_incomplete || maxl != 1)) return -1;
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int i = 0; i < n; i++) if (lens[i]) sorted[offs[lens[i]]++] = (uint16_t)i;
    count[0] = 0;
    /* direct table: canonical codes in (length, symbol) order */
    uint32_t code = 0; int k = 0;
    for (int l = 1; l <= maxl && l <= tbits; l++) {
        for (int c = 0; c < count[l]; c++, k++, code++) {
            uint32_t r = zi_rev(code, l);
            uint16_t e = (uint16_t)(sorted[k] | (l << shift));
            for (uint32_t j = r; j < (1u << tbits); j += (1u << l)) table[j] = e;
        }
        code <<= 1;
    }
    return 0;
}

/* Decode one symbol. Returns the symbol or -1 (invalid code). */
ZID int zi_decode(zi_io *io, const uint16_t *table, int tbits, const uint16_t *sorted,
                                const uint16_t *count, int shift)
{
    uint32_t e = table[zi_peek(io, tbits)];
    if (e) { zi_drop(io, (int)(e >> shift)); return (int)(e & ((1u << shift) - 1u)); }
    /* longer than tbits: canonical first-code walk over the bit-reversed prefix */
    uint32_t v = zi_rev(zi_peek(io, 15), 15);
    uint32_t first = 0, index = 0;
    for (int l = 1; l <= 15; l++) {
        uint32_t c = count[l];
        uint32_t code = v >> (15 - l);
        if (l > tbits && code - first < c) { zi_drop(io, l); return (int)sorted[index + (code - first)]; }
        index += c; first = (first + c) << 1;
    }
    return -1;
}

/* where zi_decode's canonical walk stands after the code lengths a direct table of tbits resolves */
ZID void zi_walk_start(const uint16_t *count, int tbits, uint32_t *first_out, uint32_t *index_out)
{
    uint32_t first = 0, index = 0;
    for (int k = 1; k <= tbits; k++) { const uint32_t c = count[k]; index += c; first = (first + c) << 1; }
    *first_out = first; *index_out = index;
}

/* direct entries of the codes of at most `root` bits (sorted / count as zi_build leaves them).  `complete`: the code leaves no
 * bit pattern unused, so every root slot is written here or, as the prefix of longer codes, by zi_fill_long: no need to clear
 * the table first (a thousand stores per block on one lane). */
ZID void zi_fill_root(const uint16_t *sorted, const uint16_t *count, int root, uint16_t *table, int complete)
{
    if (!complete) for (int i = 0; i < (1 << root); i++) table[i] = 0;
    uint32_t code = 0; int k = 0;
    for (int l = 1; l <= root; l++) {
        for (int c = 0; c < count[l]; c++, k++, code++) {
            const uint32_t r = zi_rev(code, l);
            const uint16_t e = (uint16_t)(sorted[k] | (l << 9));
            for (uint32_t j = r; j < (1u << root); j += (1u << l)) table[j] = e;
        }
        code <<= 1;
    }
}

/* second-level tables for the codes longer than `root` bits, as far as the pool lasts */
ZID void zi_fill_long(const uint16_t *sorted, const uint16_t *count, int root, uint16_t *table, uint16_t *pool, uint32_t *pool_used)
{
    uint32_t code = 0; int k = 0, any = 0;
    for (int l = 1; l <= root; l++) { code = (code + count[l]) << 1; k += count[l]; }
    for (int l = root + 1; l <= 15; l++) any += count[l];
    if (!any) return;
    const uint32_t rmask = (1u << root) - 1u;
    /* longest code under every root slot: lengths ascend, the last marker written wins */
    uint32_t c2 = code;
    for (int l = root + 1; l <= 15; l++) {
        for (int c = 0; c < count[l]; c++, c2++) table[zi_rev(c2, l) & rmask] = (uint16_t)(0x4000u | (uint32_t)(l - root));
        c2 <<= 1;
    }
    /* sub-tables are handed out in code order, each when the first code under its root slot comes by (a walk over the
       long codes, a few dozen, instead of one over all root slots) */
    for (int l = root + 1; l <= 15; l++) {
        for (int c = 0; c < count[l]; c++, k++, code++) {
            const uint32_t r = zi_rev(code, l), slot = r & rmask;
            uint32_t e = table[slot];
            if ((e & 0xC000u) == 0x4000u) {
                const uint32_t sb = e & 15u, need = 1u << sb;
                if (sb > 7u || *pool_used + need > ZI_POOL) e = 0;
                else {
                    for (uint32_t j = 0; j < need; j++) pool[*pool_used + j] = 0;
                    e = 0x8000u | (sb << 10) | *pool_used;
                    *pool_used += need;
                }
                table[slot] = (uint16_t)e;
            }
            if (!(e & 0x8000u)) continue;
            const uint32_t sb = (e >> 10) & 7u, off = e & 0x3FFu;
            const uint16_t v = (uint16_t)(sorted[k] | (l << 9));
            for (uint32_t j = r >> root; j < (1u << sb); j += (1u << (l - root))) pool[off + j] = v;
        }
        code <<= 1;
    }
}

/* counts, sorted symbols and both table levels of one alphabet; same accept / reject rule as zi_build */
ZID int zi_build2(const uint8_t *lens, int n, int root, uint16_t *table, uint16_t *pool, uint32_t *pool_used,
                  uint16_t *sorted, uint16_t *count, int allow_incomplete)
{
    uint16_t offs[16];
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int i = 0; i < n; i++) count[lens[i]]++;
    int maxl = 15;
    while (maxl > 0 && count[maxl] == 0) maxl--;
    if (maxl == 0) { count[0] = 0; for (int i = 0; i < (1 << root); i++) table[i] = 0; return allow_incomplete ? 0 : -1; }
    int left = 1;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= count[l]; if (left < 0) return -1; }
    if (left > 0 && (!allow_incomplete || maxl != 1)) return -1;
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int i = 0; i < n; i++) if (lens[i]) sorted[offs[lens[i]]++] = (uint16_t)i;
    count[0] = 0;
    zi_fill_root(sorted, count, root, table, left == 0);
    zi_fill_long(sorted, count, root, table, pool, pool_used);
    return 0;
}

/* Decode one symbol through both table levels, else by the canonical walk. Returns the symbol or -1 (invalid code). */
ZID int zi_decode2(zi_io *io, const uint16_t *table, int root, const uint16_t *pool, const uint16_t *sorted, const uint16_t *count)
{
    const uint32_t b = zi_peek(io, 15);
    uint32_t e = table[b & ((1u << root) - 1u)];
    if (e & 0x8000u) e = pool[(e & 0x3FFu) + ((b >> root) & ((1u << ((e >> 10) & 7u)) - 1u))];
    if (e) { zi_drop(io, (int)(e >> 9)); return (int)(e & 511u); }
    const uint32_t v = zi_rev(b, 15);
    uint32_t first = 0, index = 0;
    for (int l = 1; l <= 15; l++) {
        const uint32_t c = count[l];
        const uint32_t code = v >> (15 - l);
        if (l > root && code - first < c) { zi_drop(io, l); return (int)sorted[index + (code - first)]; }
        index += c; first = (first + c) << 1;
    }
    return -1;
}

ZID int zi_fail(zi_result *r, int ret, int reason)
{
    if (r->reason == ZI_E_NONE) r->reason = reason;
    r->last_reason = reason;
    r->ret = ret;
    return ret;
}

/* Block header: BFINAL / BTYPE, then either the stored-block length (cursor left on the first payload
 * byte) or the decode tables of a fixed / dynamic block.  Returns ZI_OK or the failure code. */
ZID int zi_block_head(zi_io *io, zi_tables *T, zi_aux *X, zi_result *res, uint32_t *last_out, uint32_t *type_out, uint32_t *stored_len)
{
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    zi_refill(io);
    uint32_t last = zi_take(io, 1), type = zi_take(io, 2);
    *last_out = last; *type_out = type;
    if (zi_overrun(io)) return zi_fail(res, ZI_BUF_ERROR, ZI_E_INPUT_END);
    if (type == 3) return zi_fail(res, ZI_DATA_ERROR, ZI_E_BLOCK_TYPE);
    if (type == 0) {
        zi_drop(io, (int)(io->bits & 7));
        zi_refill(io);
        uint32_t v = zi_take32(io);
        if (zi_overrun(io)) return zi_fail(res, ZI_BUF_ERROR, ZI_E_INPUT_END);
        uint32_t len = v & 0xFFFF;
        if (len != ((v >> 16) ^ 0xFFFF)) return zi_fail(res, ZI_DATA_ERROR, ZI_E_STORED_LEN);
        /* rewind the bit buffer to a byte cursor */
        zi_seek(io, io->ip - (io->bits >> 3));
        *stored_len = len;
        return ZI_OK;
    }
    {
            if (type == 1) {
                /* fixed code (RFC 1951 3.2.6): 32 five-bit distance codes (30 and 31 are invalid when
                   used, src/inflate.c:122-206), literal/length lengths 8/9/7/8 in closed form */
                uint8_t tmp[32];
                for (int i = 0; i < 32; i++) tmp[i] = 5;
                T->pool_used = 0;
                (void)zi_build2(tmp, 32, ZI_DBITS, T->dist, T->pool, &T->pool_used, X->dsorted, T->dcount, 1);
                for (int i = 0; i < 16; i++) T->lcount[i] = 0;
                T->lcount[7] = 24; T->lcount[8] = 152; T->lcount[9] = 112;
                int k = 0;
                for (int i = 256; i < 280; i++) X->lsorted[k++] = (uint16_t)i;
                for (int i = 0; i < 144; i++) X->lsorted[k++] = (uint16_t)i;
                for (int i = 280; i < 288; i++) X->lsorted[k++] = (uint16_t)i;
                for (int i = 144; i < 256; i++) X->lsorted[k++] = (uint16_t)i;
                zi_fill_root(X->lsorted, T->lcount, ZI_LBITS, T->lit, 1);   /* no code is longer than 9 bits */
            } else {
                zi_refill(io);
                uint32_t nlen = zi_take(io, 5) + 257, ndist = zi_take(io, 5) + 1, ncode = zi_take(io, 4) + 4;
                if (zi_overrun(io)) return zi_fail(res, ZI_BUF_ERROR, ZI_E_INPUT_END);
                if (nlen > 286 || ndist > 30) return zi_fail(res, ZI_DATA_ERROR, ZI_E_TOO_MANY_SYMS);
                uint8_t cl[19];
                for (int i = 0; i < 19; i++) cl[i] = 0;
                for (uint32_t i = 0; i < ncode; i++) { zi_refill(io); cl[order[i]] = (uint8_t)zi_take(io, 3); }
                if (zi_overrun(io)) return zi_fail(res, ZI_BUF_ERROR, ZI_E_INPUT_END);
                /* code-length code: a 7-bit direct table (128 entries) in the upper half of the literal
                   table (entries 256..383), which is dead until this header has been parsed (lens uses its
                   first 320 bytes) */
                uint16_t *cltab = T->lit + 256;
                uint16_t clsorted[19], clcount[16];
                if (zi_build(cl, 19, 7, cltab, clsorted, clcount, 5, 0)) return zi_fail(res, ZI_DATA_ERROR, ZI_E_CODELEN_SET);
                uint8_t *lens = (uint8_t *)T->lit;
                uint32_t i = 0, total = nlen + ndist;
                while (i < total) {
                    zi_refill(io);
                    int s = zi_decode(io, cltab, 7, clsorted, clcount, 5);
                    if (s < 0) return zi_fail(res, zi_overrun(io) ? ZI_BUF_ERROR : ZI_DATA_ERROR, zi_overrun(io) ? ZI_E_INPUT_END : ZI_E_CODELEN_SET);
                    if (s < 16) lens[i++] = (uint8_t)s;
                    else {
                        uint32_t rep, val = 0;
                        if (s == 16) {
                            if (i == 0) return zi_fail(res, ZI_DATA_ERROR, ZI_E_BITLEN_REPEAT);
                            val = lens[i - 1]; rep = 3 + zi_take(io, 2);
                        } else if (s == 17) rep = 3 + zi_take(io, 3);
                        else rep = 11 + zi_take(io, 7);
                        if (i + rep > total) return zi_fail(res, ZI_DATA_ERROR, ZI_E_BITLEN_REPEAT);
                        while (rep--) lens[i++] = (uint8_t)val;
                    }
                    if (zi_overrun(io)) return zi_fail(res, ZI_BUF_ERROR, ZI_E_INPUT_END);
                }
                if (lens[256] == 0) return zi_fail(res, ZI_DATA_ERROR, ZI_E_NO_EOB);
                /* distance tables first (they overwrite the code-length table), from the tail of lens */
                T->pool_used = 0;
                if (zi_build2(lens + nlen, (int)ndist, ZI_DBITS, T->dist, T->pool, &T->pool_used, X->dsorted, T->dcount, 1))
                    return zi_fail(res, ZI_DATA_ERROR, ZI_E_DIST_SET);
                /* literal/length: sorted + counts while lens is alive, then the tables over it */
                {
                    uint16_t offs[16];
                    for (int k = 0; k < 16; k++) T->lcount[k] = 0;
                    for (uint32_t k = 0; k < nlen; k++) T->lcount[lens[k]]++;
                    T->lcount[0] = 0;
                    int left = 1, maxl = 15;
                    for (int l = 1; l <= 15; l++) { left <<= 1; left -= T->lcount[l]; if (left < 0) return zi_fail(res, ZI_DATA_ERROR, ZI_E_LITLEN_SET); }
                    while (maxl > 0 && T->lcount[maxl] == 0) maxl--;
                    if (left > 0 && maxl != 1) return zi_fail(res, ZI_DATA_ERROR, ZI_E_LITLEN_SET);
                    offs[1] = 0;
                    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + T->lcount[l]);
                    for (uint32_t k = 0; k < nlen; k++) if (lens[k]) X->lsorted[offs[lens[k]]++] = (uint16_t)k;
                    zi_fill_root(X->lsorted, T->lcount, ZI_LBITS, T->lit, left == 0);
                    zi_fill_long(X->lsorted, T->lcount, ZI_LBITS, T->lit, T->pool, &T->pool_used);
                }
            }
    }
    zi_walk_start(T->lcount, ZI_LBITS, &T->lfirst, &T->lindex);
    zi_walk_start(T->dcount, ZI_DBITS, &T->dfirst, &T->dindex);
    return ZI_OK;
}

/* Scan for the next full-flush marker 00 00 FF FF at or after byte `from`; returns the position just
 * after it, or in_len + 1 when there is none (the reference's syncsearch, src/inflate.c:1523-1545). */
ZID uint32_t zi_sync(const uint8_t *in, uint32_t in_len, uint32_t from)
{
    uint32_t got = 0;     /* how much of 00 00 FF FF has been seen */
    for (uint32_t p = from; p < in_len; p++) {
        uint32_t b = in[p];
        if (b == (got < 2 ? 0u : 0xFFu)) got++;
        else if (b) got = 0;
        else got = 4 - got;          /* a zero where FF was due: the zero run restarts */
        if (got == 4) return p + 1;
    }
    return in_len + 1;
}

/* ---- the decoder as a state machine -------------------------------------------------------------
 * One call of zi_step() advances one stream by a bounded amount of work: a header, up to two symbols, or
 * up to 16 copied bytes.  On the GPU the 32 streams of a warp call it in lockstep, so lanes re-converge
 * after every step instead of each running its own nest of loops (thread-per-stream decoding is only
 * viable that way); on the host zi_inflate() simply loops over it.  Both run exactly this code.
 *
 * Recovery mirrors zsc_uncompress's loop (src/zsc_uncompr.c:103-127) around inflateSync
 * (src/inflate.c:1547-1604): after a data error the search for 00 00 FF FF starts at the first
 * byte the decoder has not pulled (`held` bytes earlier for the fields the reference still holds
 * un-dropped in its accumulator when it reports the error: the 16-bit zlib header, the 32-bit stored-block
 * LEN/NLEN, the 32-bit check value).  No bytes left and nothing held -> Z_BUF_ERROR; no marker ->
 * Z_DATA_ERROR with all input consumed; marker -> decoding continues behind it and output keeps
 * appending; a stream that needed any recovery ends as Z_DATA_ERROR. */
enum { ZM_HEAD = 0, ZM_BLOCK, ZM_SYM, ZM_COPY, ZM_STORED, ZM_TRAIL, ZM_RECOVER, ZM_DONE, ZM_DICT /* streaming: waiting for inflateSetDictionary */ };

typedef struct {
    zi_io io;
    zi_result res;
    zi_tables *T;
    zi_aux *X;
    int32_t state, wrap;
    uint32_t last, rem, dist, win, maxw, held;
    uint32_t base;          /* output position of the last resynchronisation: no distance may reach behind it */
    uint32_t hist;          /* streaming: valid bytes in front of out[0] (earlier calls' output or a preset dictionary) */
    uint32_t skip;          /* streaming: bits of in[0] an earlier call has already used */
    uint32_t opts;          /* ZI_OPT_*: section-parallel decoding of one stream (engine.cu zs_inflate_sectioned) */
} zi_mach;

#define ZI_OPT_COUNT_ONLY 1u        /* advance the output position without writing (sizes of the sections) */
#define ZI_OPT_STOP_AT_FLUSH 2u     /* end, successfully, behind the first empty non-final stored block */
#define ZI_OPT_STREAMING 8u         /* the z_stream API (zs_inflate_stream_kernel): a preset-dictionary header waits in ZM_DICT
                                       instead of ending the stream with Z_NEED_DICT (4 is taken by the kernels' section flag) */

ZID void zi_m_init(zi_mach *m, const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap, int wrap, zi_tables *T, zi_aux *X)
{
    m->io.in = in; m->io.in_len = in_len; m->io.ip = 0; m->io.hold = 0; m->io.bits = 0;
    m->io.out = out; m->io.out_cap = out_cap; m->io.op = 0; m->io.pre = 0; m->io.pv = 0;
    m->res.ret = ZI_OK; m->res.reason = ZI_E_NONE; m->res.produced = 0; m->res.consumed = 0;
    m->res.data_errors = 0; m->res.stored_check = 0; m->res.have_check = 0; m->res.last_reason = ZI_E_NONE;
    /* wrap: low byte 0 raw / 1 zlib; bits 8..15 = largest window_bits the caller accepts (0 = 15) */
    m->maxw = ((wrap >> 8) & 0xFF) ? (uint32_t)((wrap >> 8) & 0xFF) : 15u;
    m->wrap = wrap & 0xFF;
    m->win = 1u << m->maxw;
    m->T = T; m->X = X; m->last = 0; m->rem = 0; m->dist = 0; m->held = 0; m->opts = 0; m->base = 0; m->hist = 0; m->skip = 0;
    m->res.at_flush = 0;
    m->state = m->wrap == 1 ? ZM_HEAD : ZM_BLOCK;
}

ZID void zi_m_finish(zi_mach *m, int ret)
{
    zi_io *io = &m->io;
    m->res.produced = io->op;
    uint32_t c = zi_consumed_bytes(io);
    m->res.consumed = c > io->in_len ? io->in_len : c;
    if (ret == ZI_OK && m->res.data_errors) ret = ZI_DATA_ERROR;
    if (ret == ZI_NEED_DICT) m->res.consumed = 0;     /* the reference returns before updating total_in (src/inflate.c:970-973) */
    m->res.ret = ret;
    m->state = ZM_DONE;
}

/* a failure code from a step: data errors go to recovery, everything else ends the stream */
ZID void zi_m_fail(zi_mach *m, int r, uint32_t held)
{
    if (r == ZI_DATA_ERROR) { m->held = held; m->state = ZM_RECOVER; }
    else zi_m_finish(m, r);
}

ZID void zi_step(zi_mach *m)
{
    zi_io *io = &m->io;
    zi_result *res = &m->res;
    zi_tables *T = m->T;
    zi_aux *X = m->X;
    if (m->state == ZM_SYM) {
        /* up to two symbols: most are literals, and two keep the lanes of a warp busy between copies */
        for (int rep = 0; rep < 2 && m->state == ZM_SYM; rep++) {
            zi_refill(io);
            const int s = zi_decode2(io, T->lit, ZI_LBITS, T->pool, X->lsorted, T->lcount);
            if (zi_overrun(io)) { zi_m_fail(m, zi_fail(res, ZI_BUF_ERROR, ZI_E_INPUT_END), 0); break; }
            if (s < 0) { zi_m_fail(m, zi_fail(res, ZI_DATA_ERROR, ZI_E_LITLEN_CODE), 0); break; }
            if (s < 256) {
                if (io->op >= io->out_cap) { zi_m_fail(m, zi_fail(res, ZI_BUF_ERROR, ZI_E_OUTPUT_FULL), 0); break; }
                if (!(m->opts & ZI_OPT_COUNT_ONLY)) io->out[io->op] = (uint8_t)s;
                io->op++;
                continue;
            }
            if (s == 256) { m->state = m->last ? ZM_TRAIL : ZM_BLOCK; break; }
            if (s > 285) { zi_m_fail(m, zi_fail(res, ZI_DATA_ERROR, ZI_E_LITLEN_CODE), 0); break; }
            uint32_t c = (uint32_t)s - 257, len;
            if (c < 8) len = 3 + c;
            else if (c == 28) len = 258;
            else { uint32_t eb = (c - 4) >> 2; len = 3 + ((4 + (c & 3)) << eb) + zi_take(io, (int)eb); }
            zi_refill(io);
            const int d = zi_decode2(io, T->dist, ZI_DBITS, T->pool, X->dsorted, T->dcount);
            if (d < 0 || d > 29) {
                int ov = zi_overrun(io);
                zi_m_fail(m, zi_fail(res, ov ? ZI_BUF_ERROR : ZI_DATA_ERROR, ov ? ZI_E_INPUT_END : ZI_E_DIST_CODE), 0);
                break;
            }
            uint32_t dist;
            if (d < 4) dist = 1 + (uint32_t)d;
            else { uint32_t eb = ((uint32_t)d - 2) >> 1; dist = 1 + ((2 + ((uint32_t)d & 1)) << eb) + zi_take(io, (int)eb); }
            if (zi_overrun(io)) { zi_m_fail(m, zi_fail(res, ZI_BUF_ERROR, ZI_E_INPUT_END), 0); break; }
            if (dist > io->op - m->base + m->hist || dist > m->win) { zi_m_fail(m, zi_fail(res, ZI_DATA_ERROR, ZI_E_DIST_TOO_FAR), 0); break; }
            m->rem = len; m->dist = dist; m->state = ZM_COPY;
        }
    }
    if (m->state == ZM_COPY) {
        uint32_t room = io->out_cap - io->op;
        uint32_t n = m->rem < 16 ? m->rem : 16;
        if (n > room) n = room;
        if (!(m->opts & ZI_OPT_COUNT_ONLY)) zi_copy_match(io->out + io->op, m->dist, n);
        io->op += n; m->rem -= n;
        if (m->rem == 0) m->state = ZM_SYM;
        else if (n == 0 || io->op >= io->out_cap) zi_m_fail(m, zi_fail(res, ZI_BUF_ERROR, ZI_E_OUTPUT_FULL), 0);
        return;
    }
    if (m->state == ZM_STORED) {
        uint32_t pos = io->ip;
        uint32_t avail = pos < io->in_len ? io->in_len - pos : 0, room = io->out_cap - io->op;
        uint32_t n = m->rem < 16 ? m->rem : 16;
        if (n > avail) n = avail;
        if (n > room) n = room;
        if (!(m->opts & ZI_OPT_COUNT_ONLY)) zi_copy_fwd(io->out + io->op, io->in + pos, n);
        io->op += n; io->ip = pos + n; io->pv = 0; m->rem -= n;
        if (m->rem == 0) m->state = m->last ? ZM_TRAIL : ZM_BLOCK;
        else if (n == 0) zi_m_fail(m, zi_fail(res, ZI_BUF_ERROR, room == 0 ? ZI_E_OUTPUT_FULL : ZI_E_INPUT_END), 0);
        return;
    }
    if (m->state == ZM_HEAD) {
        zi_refill(io);
        uint32_t h = zi_take(io, 16);
        int r = ZI_OK;
        if (zi_overrun(io)) r = zi_fail(res, ZI_BUF_ERROR, ZI_E_INPUT_END);
        else if ((((h & 0xFF) << 8) | (h >> 8)) % 31) r = zi_fail(res, ZI_DATA_ERROR, ZI_E_HEADER_CHECK);
        else if ((h & 0xF) != 8) r = zi_fail(res, ZI_DATA_ERROR, ZI_E_METHOD);
        else if (((h >> 4) & 0xF) + 8 > m->maxw) r = zi_fail(res, ZI_DATA_ERROR, ZI_E_WINDOW);
        else if (h & 0x2000) {
            if (m->opts & ZI_OPT_STREAMING) {
                /* DICTID follows the header (src/inflate.c:957-973): report it and wait for the dictionary */
                zi_refill(io);
                const uint32_t t = zi_take32(io);
                if (zi_overrun(io)) r = zi_fail(res, ZI_BUF_ERROR, ZI_E_INPUT_END);
                else {
                    res->stored_check = ((t & 0xFF) << 24) | ((t & 0xFF00) << 8) | ((t >> 8) & 0xFF00) | (t >> 24);
                    m->win = 1u << (((h >> 4) & 0xF) + 8);
                    m->state = ZM_DICT;
                    return;
                }
            } else r = zi_fail(res, ZI_NEED_DICT, ZI_E_NEED_DICT);
        }
        if (r == ZI_OK) { m->win = 1u << (((h >> 4) & 0xF) + 8); m->state = ZM_BLOCK; }
        else zi_m_fail(m, r, 2);
        return;
    }
    if (m->state == ZM_BLOCK) {
        uint32_t type = 0, slen = 0;
        int r = zi_block_head(io, T, X, res, &m->last, &type, &slen);
        if (r != ZI_OK) { zi_m_fail(m, r, (r == ZI_DATA_ERROR && res->last_reason == ZI_E_STORED_LEN) ? 4u : 0u); return; }
        if (type == 0) {
            if (slen == 0 && !m->last && (m->opts & ZI_OPT_STOP_AT_FLUSH)) { res->at_flush = 1; zi_m_finish(m, ZI_OK); return; }
            m->rem = slen; m->state = slen ? ZM_STORED : (m->last ? ZM_TRAIL : ZM_BLOCK);
        } else m->state = ZM_SYM;
        return;
    }
    if (m->state == ZM_TRAIL) {
        /* final block done: byte-align, then the 4-byte data check of a zlib stream */
        zi_drop(io, (int)(io->bits & 7));
        if (m->wrap == 1) {
            zi_refill(io);
            uint32_t t = zi_take32(io);
            if (zi_overrun(io)) { zi_m_finish(m, zi_fail(res, ZI_BUF_ERROR, ZI_E_INPUT_END)); return; }
            if (res->data_errors) {
                /* the reference's running check restarted at the flush point it resynchronised to, so it
                   cannot match the whole-stream trailer: it reports a data error with those 4 bytes held */
                zi_m_fail(m, zi_fail(res, ZI_DATA_ERROR, ZI_E_DATA_CHECK), 4);
                return;
            }
            res->stored_check = ((t & 0xFF) << 24) | ((t & 0xFF00) << 8) | ((t >> 8) & 0xFF00) | (t >> 24);
            res->have_check = 1;
        }
        zi_m_finish(m, ZI_OK);
        return;
    }
    if (m->state == ZM_RECOVER) {
        res->data_errors++;
        uint32_t pos = io->ip - (io->bits >> 3);
        if (pos > io->in_len) pos = io->in_len;
        uint32_t held = m->held;
        m->held = 0;
        if (held == 0 && pos >= io->in_len) { zi_seek(io, io->in_len); zi_m_finish(m, ZI_BUF_ERROR); return; }
        uint32_t nx = zi_sync(io->in, io->in_len, pos - held);
        if (nx > io->in_len) { zi_seek(io, io->in_len); zi_m_finish(m, ZI_DATA_ERROR); return; }
        zi_seek(io, nx);
        /* inflateSync resets the stream (src/inflate.c:1593 -> inflateReset: whave = 0, dmax = 32768) and the next
           inflate() call measures distances from its own first output byte (src/inflate.c:1284-1295, inffast.c:190-200) */
        m->base = io->op;
        m->hist = 0;
        m->win = 32768u;
        m->state = ZM_BLOCK;
        return;
    }
}

/* ---- group form of the decoder: G lanes per stream (zs_inflate_group_kernel) ------------------------
 * The leader lane of a group runs the state machine above; inside a compressed block it first lets
 * zi_fast_batch decode up to G symbols into a record queue without touching the output, and the G lanes then
 * write the literals and perform the copies together.  Records use the LZ kernel's symbol format: literal
 * byte, or bit 31 | (len - 3) << 16 | (dist - 1).  zi_fast_batch is a pure accelerator: whatever it does not
 * take (see below) is left, unconsumed, to zi_step, so results are those of the one-thread decoder by
 * construction; tests/ runs both forms over the same vectors and asserts equality. */
#define ZI_BATCH 32

/* ---- the accelerated symbol decoder of the group kernel (zs_inflate_group_kernel) --------------------
 * base | extra bits << 16 of a length code (0..28) / distance code (0..29), RFC 1951 3.2.5 */
ZID uint32_t zi_lut_len(uint32_t c)
{
    const uint32_t eb = (c < 8 || c == 28) ? 0u : ((c - 4) >> 2);
    return (c < 8 ? 3u + c : c == 28 ? 258u : 3u + ((4u + (c & 3u)) << eb)) | (eb << 16);
}
ZID uint32_t zi_lut_dist(uint32_t d)
{
    const uint32_t eb = d < 4 ? 0u : ((d - 2) >> 1);
    return (d < 4 ? 1u + d : 1u + ((2u + (d & 1u)) << eb)) | (eb << 16);
}
/* On the GPU the decode tables, the base/extra-bits tables and the record queue of zi_fast_batch are in shared
 * memory.  Their 32-bit shared-space addresses are taken once and made opaque to the compiler: left to itself it
 * recomputes them from threadIdx for every symbol (two S2R and five arithmetic instructions, measured as 6 of the
 * 35 instructions of the literal path) because the loop is short of registers. */
#ifdef __CUDA_ARCH__
typedef uint32_t zi_sa;
static __device__ __forceinline__ zi_sa zi_sa_of(const void *p) { uint32_t v = (uint32_t)__cvta_generic_to_shared(p); asm volatile("" : "+r"(v)); return v; }
static __device__ __forceinline__ uint32_t zi_sa_ld16(zi_sa a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
static __device__ __forceinline__ uint32_t zi_sa_ld32(zi_sa a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
static __device__ __forceinline__ void zi_sa_st32(zi_sa a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v)); }
#else
typedef uintptr_t zi_sa;
static inline zi_sa zi_sa_of(const void *p) { return (zi_sa)p; }
static inline uint32_t zi_sa_ld16(zi_sa a) { return *(const uint16_t *)a; }
static inline uint32_t zi_sa_ld32(zi_sa a) { return *(const uint32_t *)a; }
static inline void zi_sa_st32(zi_sa a, uint32_t v) { *(uint32_t *)a = v; }
#endif

ZID uint32_t zi_fast_batch(zi_mach *m, const uint32_t *lut_len, const uint32_t *lut_dist, uint32_t *q, uint32_t maxn, uint32_t *vop)
{
    const zi_tables *T = m->T;
    const zi_sa lit_a = zi_sa_of(T->lit), dist_a = zi_sa_of(T->dist), pool_a = zi_sa_of(T->pool), len_a = zi_sa_of(lut_len), dl_a = zi_sa_of(lut_dist), q_a = zi_sa_of(q);
    uint32_t n = 0, op = m->io.op;
    /* how many symbols this batch may take: each needs at most 258 bytes of output room and pulls at most
       8 bytes of input (two refills); the last 16 input bytes are left to zi_step */
    uint32_t nmax = maxn;
    {
        const uint32_t room = m->io.out_cap - op;
        if (room < nmax * 258u) nmax = room / 258u;
        const uint64_t need = (uint64_t)m->io.ip + 16u;
        const uint32_t left = (uint64_t)m->io.in_len > need ? (uint32_t)(m->io.in_len - need) : 0u;
        if (left < nmax * 8u) nmax = left / 8u;
    }
    if (nmax == 0) { *vop = op; return 0; }
    zi_io io = m->io;                                    /* the cursor in registers for the whole batch */
    const uint32_t win = m->win, base = m->base - m->hist;   /* (wraps when hist > base: the comparison below is on the difference) */
    while (n < nmax) {
        zi_refill(&io);                                  /* >= 33 bits: a literal/length code and its extra bits */
        const uint64_t h = io.hold;
        const uint32_t b = io.bits;
        uint32_t l, sym;
        {
            uint32_t e = zi_sa_ld16(lit_a + 2u * ((uint32_t)h & ((1u << ZI_LBITS) - 1u)));
            if (e & 0x8000u) e = zi_sa_ld16(pool_a + 2u * ((e & 0x3FFu) + (((uint32_t)h >> ZI_LBITS) & ((1u << ((e >> 10) & 7u)) - 1u))));
            if (e) { l = e >> 9; sym = e & 511u; }
            else {
                const uint32_t v = zi_rev((uint32_t)h & 0x7FFFu, 15);
                uint32_t first = T->lfirst, index = T->lindex;
                l = 0; sym = 0;
                for (uint32_t k = ZI_LBITS + 1; k <= 15; k++) {
                    const uint32_t c = T->lcount[k], code = v >> (15 - k);
                    if (l == 0 && code - first < c) { l = k; sym = m->X->lsorted[index + (code - first)]; }
                    index += c; first = (first + c) << 1;
                }
                if (l == 0) break;
            }
        }
        if (sym < 256u) { io.hold = h >> l; io.bits = b - l; zi_sa_st32(q_a + 4u * n, sym); n++; op++; continue; }
        if (sym == 256u || sym > 285u) break;
        const uint32_t lb = zi_sa_ld32(len_a + 4u * (sym - 257u)), eb = lb >> 16;
        const uint32_t len = (lb & 0xFFFFu) + ((uint32_t)(h >> l) & ((1u << eb) - 1u));
        const uint32_t ip0 = io.ip, pre0 = io.pre, pv0 = io.pv;
        io.hold = h >> (l + eb); io.bits = b - (l + eb);
        zi_refill(&io);                                  /* >= 33 bits again: a distance code and its extra bits */
        const uint64_t h2 = io.hold;
        const uint32_t b2 = io.bits;
        uint32_t l2, d;
        {
            uint32_t e = zi_sa_ld16(dist_a + 2u * ((uint32_t)h2 & ((1u << ZI_DBITS) - 1u)));
            if (e & 0x8000u) e = zi_sa_ld16(pool_a + 2u * ((e & 0x3FFu) + (((uint32_t)h2 >> ZI_DBITS) & ((1u << ((e >> 10) & 7u)) - 1u))));
            if (e) { l2 = e >> 9; d = e & 511u; }
            else {
                const uint32_t v = zi_rev((uint32_t)h2 & 0x7FFFu, 15);
                uint32_t first = T->dfirst, index = T->dindex;
                l2 = 0; d = 31;
                for (uint32_t k = ZI_DBITS + 1; k <= 15; k++) {
                    const uint32_t c = T->dcount[k], code = v >> (15 - k);
                    if (l2 == 0 && code - first < c) { l2 = k; d = m->X->dsorted[index + (code - first)]; }
                    index += c; first = (first + c) << 1;
                }
            }
        }
        uint32_t dist = 0, eb2 = 0;
        if (l2 != 0 && d <= 29u) {
            const uint32_t db = zi_sa_ld32(dl_a + 4u * d);
            eb2 = db >> 16;
            dist = (db & 0xFFFFu) + ((uint32_t)(h2 >> l2) & ((1u << eb2) - 1u));
        }
        if (dist == 0 || dist > op - base || dist > win) {
            /* not ours: put the cursor back in front of the length code */
            io.hold = h; io.bits = b; io.ip = ip0; io.pre = pre0; io.pv = pv0;
            break;
        }
        io.hold = h2 >> (l2 + eb2); io.bits = b2 - (l2 + eb2);
        zi_sa_st32(q_a + 4u * n, 0x80000000u | ((len - 3u) << 16) | (dist - 1u)); n++;
        op += len;
    }
    m->io.hold = io.hold; m->io.bits = io.bits; m->io.ip = io.ip; m->io.pre = io.pre; m->io.pv = io.pv;
    *vop = op;
    return n;
}

/* stored block in one piece: how many bytes can move now (the copy itself is the caller's), then the
 * bookkeeping zi_step would have arrived at 16 bytes at a time */
ZID uint32_t zi_stored_plan(const zi_mach *m)
{
    const zi_io *io = &m->io;
    const uint32_t avail = io->ip < io->in_len ? io->in_len - io->ip : 0, room = io->out_cap - io->op;
    uint32_t n = m->rem;
    if (n > avail) n = avail;
    if (n > room) n = room;
    return n;
}
ZID void zi_stored_done(zi_mach *m, uint32_t n)
{
    zi_io *io = &m->io;
    io->op += n; io->ip += n; io->pv = 0; m->rem -= n;
    if (m->rem == 0) m->state = m->last ? ZM_TRAIL : ZM_BLOCK;
    else zi_m_fail(m, zi_fail(&m->res, ZI_BUF_ERROR, io->out_cap == io->op ? ZI_E_OUTPUT_FULL : ZI_E_INPUT_END), 0);
}

/* Whole stream on one thread: wrapper (wrap 1 = zlib, 0 = raw), blocks, trailer, corruption recovery.
 * The adler32 of the output is verified by the caller (a separate HBM-streaming pass on the GPU);
 * res->stored_check/have_check report the trailer. */
ZID void zi_inflate(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap,
                                  int wrap, zi_tables *T, zi_result *res)
{
    zi_mach m;
    zi_aux X;
    zi_m_init(&m, in, in_len, out, out_cap, wrap, T, &X);
    while (m.state != ZM_DONE) zi_step(&m);
    *res = m.res;
}

/* The group form run serially (host): same control flow as zs_inflate_group_kernel with the cooperative writes
 * replaced by plain loops.  Used by tests/ to pin the batch logic against zi_inflate on the CPU. */
ZID void zi_inflate_batched(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap,
                            int wrap, zi_tables *T, zi_result *res, uint32_t group)
{
    zi_mach m;
    zi_aux X;
    uint32_t q[ZI_BATCH], lut_len[32], lut_dist[32];
    for (uint32_t c = 0; c < 29; c++) lut_len[c] = zi_lut_len(c);
    for (uint32_t d = 0; d < 30; d++) lut_dist[d] = zi_lut_dist(d);
    zi_m_init(&m, in, in_len, out, out_cap, wrap, T, &X);
    while (m.state != ZM_DONE) {
        if (m.state == ZM_SYM) {
            uint32_t vop = 0, p = m.io.op;
            const uint32_t n = zi_fast_batch(&m, lut_len, lut_dist, q, group, &vop);
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t r = q[i];
                if (r >> 31) { const uint32_t len = ((r >> 16) & 0xFF) + 3, dist = (r & 0x7FFF) + 1; for (uint32_t k = 0; k < len; k++, p++) out[p] = out[p - dist]; }
                else out[p++] = (uint8_t)r;
            }
            if (n) m.io.op = vop;
            if (n < group) zi_step(&m);
        } else if (m.state == ZM_STORED) {
            const uint32_t n = zi_stored_plan(&m);
            for (uint32_t k = 0; k < n; k++) out[m.io.op + k] = in[m.io.ip + k];
            zi_stored_done(&m, n);
        } else zi_step(&m);
    }
    *res = m.res;
}

#endif
