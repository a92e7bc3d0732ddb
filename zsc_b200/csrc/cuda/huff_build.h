/* huff_build.h — per-block Huffman stage of the B200 deflate engine (host + device).
 *
 * Replaces, for one deflate block, the reference's build_tree / gen_bitlen / gen_codes /
 * build_bl_tree / send_all_trees and the stored-static-dynamic decision of _tr_flush_block
 * (reference src/trees.c:426-941).  It is NOT a transcription: code lengths come from a sort +
 * two-queue merge (no heap), the length limit is enforced on the length histogram, and the whole
 * dynamic header is rendered once into a dense bit string that the encode kernel copies with
 * shifts.  The bit-level FORMAT it emits is RFC 1951 (reference doc/rfc1951.txt), which is what the
 * reference's inflate (src/inflate.c:1050-1178) parses.
 *
 * The functions are `__host__ __device__` so that tests can run exactly the code the GPU runs.
 */
#ifndef ZSC_HUFF_BUILD_H
#define ZSC_HUFF_BUILD_H

#include <stdint.h>

#ifdef __CUDACC__
#define ZHD __host__ __device__
#else
#define ZHD
#endif

#define ZH_LCODES 286
#define ZH_LCODES_PAD 288
#define ZH_DCODES 30
#define ZH_DCODES_PAD 32
#define ZH_BLCODES 19
#define ZH_HDR_WORDS 96            /* 3072 bits: > worst-case dynamic header (see zh_build_block) */

enum { ZH_STORED = 0, ZH_STATIC = 1, ZH_DYNAMIC = 2, ZH_UNUSED = 3, ZH_STORED_CONT = 4 };

/* Per-block result consumed by the encode kernel. code entries: (bit-reversed code) | len << 16. */
typedef struct {
    uint32_t type;                    /* ZH_* */
    uint32_t hdr_bits;                /* valid bits in hdr[] (includes the 3-bit block header) */
    uint32_t body_bits;               /* hdr_bits + all symbol bits + end-of-block (types 1,2) */
    uint32_t nsym;                    /* symbols in this block */
    uint32_t in_start;                /* input offset (chunk-relative) where this block starts */
    uint32_t in_len;                  /* input bytes the block covers */
    uint32_t flags;                   /* ZB_* */
    uint32_t stored_total;            /* head of a merged stored run: bytes of the whole run (its LEN field) */
    uint64_t bitoff;                  /* absolute bit offset in the comp arena (offset pass) */
    uint64_t sym_off;                 /* offset of the first symbol in the sym arena */
    uint32_t hdr[ZH_HDR_WORDS];
    uint32_t lcode[ZH_LCODES_PAD];
    uint32_t dcode[ZH_DCODES_PAD];
} zh_block;

enum {
    ZB_FIRST_OF_STREAM = 1,   /* stream header bytes precede the block */
    ZB_LAST_OF_SECTION = 2,   /* full-flush marker follows the block */
    ZB_LAST_OF_STREAM = 4,    /* BFINAL set; trailer follows */
    ZB_STREAM_FAILED = 8      /* set by the offset pass when the stream does not fit: write nothing */
};

typedef struct {
    uint32_t key[ZH_LCODES_PAD];          /* (freq << 9) | symbol, sorted ascending */
    uint32_t w[2 * ZH_LCODES_PAD];        /* node weights: leaves then internal nodes */
    uint16_t parent[2 * ZH_LCODES_PAD];
    uint8_t  depth[2 * ZH_LCODES_PAD];
    uint8_t  llen[ZH_LCODES_PAD];
    uint8_t  dlen[ZH_DCODES_PAD];
    uint8_t  bllen[ZH_BLCODES + 1];
    uint16_t tok[ZH_LCODES_PAD + ZH_DCODES_PAD + 8];   /* RLE tokens: sym | extra << 8 */
    uint32_t blfreq[ZH_BLCODES + 1];
    uint32_t blcode[ZH_BLCODES + 1];
    uint32_t tmpfreq[ZH_LCODES_PAD];
} zh_scratch;

/* Chain search (levels 2..9, deflate_chain.cu; tests/cpu_harness.cpp mirrors it): tiles of ZC_TILE positions are
 * searched in ZC_ROUNDS rounds.  A round parses the tile with the match lengths known so far and lets the positions
 * the parse visits (and the lazy-evaluation candidates behind the matches it takes) follow their hash chains
 * zc_round_cap(r) candidates further; the last round goes to the end of the level's budget. */
#define ZC_TILE 2048u
#ifndef ZC_ROUNDS
#define ZC_ROUNDS 2
#define ZC_CAP0 32
#define ZC_CAP_SHIFT 2
#endif
/* candidates a position may examine beyond the first: the level's budget in full 256 KiB chunks, up to eight times that in
 * shorter ones (262144 / length, rounded down) — there the whole search costs microseconds, and small inputs and small
 * sections are the ones that lose most to a shallow search */
ZHD static inline int zc_budget(int chain, uint32_t chunk_len) { return chain * (int)(262144u / (chunk_len < 32768u ? 32768u : chunk_len)); }
ZHD static inline int zc_round_cap(int r) { return r + 1 < ZC_ROUNDS ? (ZC_CAP0 << (ZC_CAP_SHIFT * r)) : 0x7FFFFFFF; }

/* the same working arrays sized for the 30-symbol distance alphabet (a lane-per-block kernel keeps one per
   thread): the stage functions below are templates over the scratch type and behave identically on both */
typedef struct {
    uint32_t key[ZH_DCODES_PAD];
    uint32_t w[2 * ZH_DCODES_PAD];
    uint16_t parent[2 * ZH_DCODES_PAD];
    uint8_t  depth[2 * ZH_DCODES_PAD];
    uint32_t tmpfreq[ZH_DCODES_PAD];
} zh_small;

ZHD static inline int zh_extra_lbits(int c) { return (c < 8 || c == 28) ? 0 : ((c - 4) >> 2); }
ZHD static inline int zh_extra_dbits(int c) { return c < 4 ? 0 : ((c - 2) >> 1); }

ZHD static inline uint32_t zh_bitrev(uint32_t v, int n)
{
#ifdef __CUDA_ARCH__
    return n ? (__brev(v) >> (32 - n)) : 0u;
#else
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
#endif
}

/* Code lengths for freq[0..n), limited to maxbits, in three stages so that the block kernel can run the
 * sort on all its threads: prepare (keys = (freq << 9) | symbol of the used symbols; at least two symbols
 * end up with a code, as the reference guarantees for its trees, src/trees.c:595-610, so that the decoder
 * always sees a complete code), sort ascending, finish (two-queue merge, depths, length limit). */
template <class SC> ZHD static inline int zh_lengths_prepare(const uint32_t *freq_in, int n, SC *s, int *max_code_out)
{
    uint32_t *freq = s->tmpfreq;
    int m = 0, max_code = -1;
    for (int i = 0; i < n; i++) { freq[i] = freq_in[i]; if (freq[i]) { m++; max_code = i; } }
    while (m < 2) {
        int node = (max_code < 2) ? ++max_code : 0;
        if (freq[node] == 0) { freq[node] = 1; m++; }
        else { /* node 0 already used: take the next free one */
            int k = 0; while (freq[k]) k++;
            freq[k] = 1; m++; if (k > max_code) max_code = k;
        }
    }
    m = 0;
    for (int i = 0; i < n; i++) if (freq[i]) s->key[m++] = (freq[i] << 9) | (uint32_t)i;
    *max_code_out = max_code;
    return m;
}

ZHD static inline void zh_sort_keys(uint32_t *key, int m)
{
    /* shell sort, ascending by (freq, symbol); keys are unique */
    const int gaps[6] = {132, 57, 23, 10, 4, 1};
    for (int gi = 0; gi < 6; gi++) {
        int gap = gaps[gi];
        for (int i = gap; i < m; i++) {
            uint32_t v = key[i]; int j = i;
            while (j >= gap && key[j - gap] > v) { key[j] = key[j - gap]; j -= gap; }
            key[j] = v;
        }
    }
}

/* two-queue merge over the sorted keys: leaves [0,m), internal nodes [m, 2m-1) in creation (= weight)
 * order; fills s->w and s->parent, returns the node count (root = count - 1) */
template <class SC> ZHD static inline int zh_merge(int m, SC *s)
{
    for (int i = 0; i < m; i++) s->w[i] = s->key[i] >> 9;
    int a = 0, b = m, e = m;
    while ((m - a) + (e - b) > 1) {
        int x0, x1;
        if (a < m && (b >= e || s->w[a] <= s->w[b])) x0 = a++; else x0 = b++;
        if (a < m && (b >= e || s->w[a] <= s->w[b])) x1 = a++; else x1 = b++;
        s->w[e] = s->w[x0] + s->w[x1];
        s->parent[x0] = (uint16_t)e; s->parent[x1] = (uint16_t)e;
        e++;
    }
    return e;
}

/* Clipping leaf depths to maxbits over-subscribes the code by `excess` units of 2^-maxbits.  Each step
 * below moves one leaf down from the deepest level that still has room and makes a clipped leaf its
 * sibling, which removes exactly one unit (the repair idea of the reference's gen_bitlen overflow loop,
 * src/trees.c:474-507, restated on the Kraft sum); then lengths are handed out by frequency, longest to
 * rarest (s->depth[i] of the i-th sorted key). */
template <class SC> ZHD static inline void zh_repair(int maxbits, uint32_t *bl_count, SC *s)
{
    int64_t excess = -((int64_t)1 << maxbits);
    for (int l = 1; l <= maxbits; l++) excess += (int64_t)bl_count[l] << (maxbits - l);
    while (excess > 0) {
        int bits = maxbits - 1;
        while (bl_count[bits] == 0) bits--;
        bl_count[bits]--; bl_count[bits + 1] += 2; bl_count[maxbits]--;
        excess--;
    }
    int i = 0;
    for (int bits = maxbits; bits >= 1; bits--)
        for (uint32_t k = 0; k < bl_count[bits]; k++) { s->depth[i] = (uint8_t)bits; i++; }
}

template <class SC> ZHD static inline void zh_lengths_finish(int m, int n, int maxbits, uint8_t *len, SC *s)
{
    for (int i = 0; i < n; i++) len[i] = 0;
    int e = zh_merge(m, s);
    /* depths, root first; leaf depth clipped to maxbits with the clipped count recorded */
    uint32_t bl_count[16];
    for (int i = 0; i < 16; i++) bl_count[i] = 0;
    s->depth[e - 1] = 0;
    for (int i = e - 2; i >= m; i--) {
        int d = s->depth[s->parent[i]] + 1;
        s->depth[i] = (uint8_t)(d > 250 ? 250 : d);
    }
    int overflow = 0;
    for (int i = m - 1; i >= 0; i--) {
        int d = s->depth[s->parent[i]] + 1;
        if (d > maxbits) { d = maxbits; overflow++; }
        s->depth[i] = (uint8_t)d;
        bl_count[d]++;
    }
    if (overflow > 0) zh_repair(maxbits, bl_count, s);
    for (int i = 0; i < m; i++) len[s->key[i] & 0x1FF] = s->depth[i];
}

/* all three stages; returns the largest symbol index with a non-zero length */
template <class SC> ZHD static inline int zh_lengths(const uint32_t *freq_in, int n, int maxbits, uint8_t *len, SC *s)
{
    int max_code;
    int m = zh_lengths_prepare(freq_in, n, s, &max_code);
    zh_sort_keys(s->key, m);
    zh_lengths_finish(m, n, maxbits, len, s);
    return max_code;
}

/* Canonical codes (RFC 1951 3.2.2), stored bit-reversed for LSB-first output. out[i] = code | len << 16. */
ZHD static inline void zh_codes(const uint8_t *len, int n, uint32_t *out)
{
    uint32_t cnt[17], next[17];
    for (int i = 0; i <= 16; i++) cnt[i] = 0;
    for (int i = 0; i < n; i++) cnt[len[i]]++;
    cnt[0] = 0;
    uint32_t code = 0;
    for (int b = 1; b <= 15; b++) { code = (code + cnt[b - 1]) << 1; next[b] = code; }
    for (int i = 0; i < n; i++) {
        int l = len[i];
        out[i] = l ? (zh_bitrev(next[l]++, l) | ((uint32_t)l << 16)) : 0;
    }
}

typedef struct { uint32_t *w; uint32_t nbits; } zh_bitw;
ZHD static inline void zh_put(zh_bitw *bw, uint32_t v, int n)
{
    if (n == 0) return;
    uint32_t wi = bw->nbits >> 5, sh = bw->nbits & 31;
    bw->w[wi] |= v << sh;
    if (sh + (uint32_t)n > 32) bw->w[wi + 1] |= v >> (32 - sh);
    bw->nbits += (uint32_t)n;
}

/* RLE one code-length array into tokens (code-length alphabet 0..18). Returns new token count. */
ZHD static inline int zh_rle(const uint8_t *len, int n, uint16_t *tok, int nt, uint32_t *blfreq)
{
    int i = 0;
    while (i < n) {
        int v = len[i], j = i + 1;
        while (j < n && len[j] == v) j++;
        int run = j - i;
        if (v == 0) {
            while (run >= 11) { int r = run > 138 ? 138 : run; tok[nt++] = (uint16_t)(18 | ((r - 11) << 8)); blfreq[18]++; run -= r; }
            if (run >= 3) { tok[nt++] = (uint16_t)(17 | ((run - 3) << 8)); blfreq[17]++; run = 0; }
            while (run-- > 0) { tok[nt++] = 0; blfreq[0]++; }
        } else {
            tok[nt++] = (uint16_t)v; blfreq[v]++; run--;
            while (run >= 3) { int r = run > 6 ? 6 : run; tok[nt++] = (uint16_t)(16 | ((r - 3) << 8)); blfreq[16]++; run -= r; }
            while (run-- > 0) { tok[nt++] = (uint16_t)v; blfreq[v]++; }
        }
        i = j;
    }
    return nt;
}

typedef struct { int type, nl, nd, nbl, nt; uint32_t hdr_bits_est; uint64_t dyn, fix; } zh_decision;

/* From the code lengths and the symbol costs: RLE tokens, the code-length code, the dynamic header size and
 * the stored / static / dynamic decision (same rule as the reference, src/trees.c:902-934, on byte-rounded
 * sizes).  dyn_sym / fix_sym = bits of all symbols incl. end-of-block under the dynamic / fixed codes. */
ZHD static inline void zh_decide(int max_l, int max_d, uint64_t dyn_sym, uint64_t fix_sym, uint32_t in_len, int force,
                                 zh_scratch *s, zh_decision *D)
{
    const int bl_order[ZH_BLCODES] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    if (max_l < 256) max_l = 256;
    int nl = max_l + 1, nd = max_d + 1;
    for (int i = 0; i <= ZH_BLCODES; i++) s->blfreq[i] = 0;
    int nt = zh_rle(s->llen, nl, s->tok, 0, s->blfreq);
    nt = zh_rle(s->dlen, nd, s->tok, nt, s->blfreq);
    (void)zh_lengths(s->blfreq, ZH_BLCODES, 7, s->bllen, s);
    int nbl = ZH_BLCODES;
    while (nbl > 4 && s->bllen[bl_order[nbl - 1]] == 0) nbl--;
    uint32_t hdr = 3 + 5 + 5 + 4 + 3 * (uint32_t)nbl;
    for (int i = 0; i < nt; i++) {
        int sym = s->tok[i] & 0xFF;
        hdr += s->bllen[sym] + (sym == 16 ? 2 : sym == 17 ? 3 : sym == 18 ? 7 : 0);
    }
    uint64_t dyn = dyn_sym + hdr, fix = fix_sym + 3;
    uint64_t opt_lenb = (dyn + 7) >> 3, static_lenb = (fix + 7) >> 3;
    if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
    int type;
    if (force == ZH_STORED) type = ZH_STORED;
    else if ((uint64_t)in_len + 4 <= opt_lenb && in_len <= 65535u) type = ZH_STORED;   /* also under Z_FIXED */
    else if (force == ZH_STATIC) type = ZH_STATIC;
    else if (static_lenb == opt_lenb || hdr > 32u * ZH_HDR_WORDS - 64u) type = ZH_STATIC;
    else type = ZH_DYNAMIC;
    D->type = type; D->nl = nl; D->nd = nd; D->nbl = nbl; D->nt = nt; D->hdr_bits_est = hdr; D->dyn = dyn; D->fix = fix;
}

/* bits of one RLE token under the code-length code: (value, count) */
ZHD static inline uint32_t zh_tok_bits(uint16_t tok, const uint32_t *blcode, uint32_t *nbits)
{
    int sym = tok & 0xFF, ex = tok >> 8;
    uint32_t cl = blcode[sym] >> 16, v = blcode[sym] & 0xFFFF;
    int eb = sym == 16 ? 2 : sym == 17 ? 3 : sym == 18 ? 7 : 0;
    *nbits = cl + (uint32_t)eb;
    return v | ((uint32_t)ex << cl);
}

/* Decide the block type and render its header and code tables (serial form; the block kernel runs the same
 * steps spread over its threads and must produce identical results).
 *   lfreq[286] (with lfreq[256] already counting the end-of-block), dfreq[30]
 *   in_len: input bytes the block covers (stored-block cost), final_block: BFINAL
 *   force: -1 none, ZH_STATIC to force fixed codes (Z_FIXED), ZH_STORED to force stored (level 0)
 */
ZHD static inline void zh_build_block(const uint32_t *lfreq, const uint32_t *dfreq, uint32_t in_len,
                                      int final_block, int force, zh_block *out, zh_scratch *s)
{
    const int bl_order[ZH_BLCODES] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int max_l = zh_lengths(lfreq, ZH_LCODES, 15, s->llen, s);
    int max_d = zh_lengths(dfreq, ZH_DCODES, 15, s->dlen, s);
    uint64_t dyn = 0, fix = 0;
    for (int i = 0; i < ZH_LCODES; i++) {
        uint32_t f = lfreq[i];
        if (!f) continue;
        int ex = i >= 257 ? zh_extra_lbits(i - 257) : 0;
        int fl = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        dyn += (uint64_t)f * (uint32_t)(s->llen[i] + ex);
        fix += (uint64_t)f * (uint32_t)(fl + ex);
    }
    for (int i = 0; i < ZH_DCODES; i++) {
        uint32_t f = dfreq[i];
        if (!f) continue;
        int ex = zh_extra_dbits(i);
        dyn += (uint64_t)f * (uint32_t)(s->dlen[i] + ex);
        fix += (uint64_t)f * (uint32_t)(5 + ex);
    }
    zh_decision D;
    zh_decide(max_l, max_d, dyn, fix, in_len, force, s, &D);
    const int type = D.type, nl = D.nl, nd = D.nd, nbl = D.nbl, nt = D.nt;

    out->type = (uint32_t)type;
    out->in_len = in_len;
    for (int i = 0; i < ZH_HDR_WORDS; i++) out->hdr[i] = 0;
    zh_bitw bw; bw.w = out->hdr; bw.nbits = 0;
    if (type == ZH_STORED) {
        zh_put(&bw, (uint32_t)(final_block ? 1 : 0), 3);
        out->hdr_bits = 3;
        out->body_bits = 0;          /* stored payload is position dependent: see zk_elem_of_block */
        return;
    }
    if (type == ZH_STATIC) {
        for (int i = 0; i < ZH_LCODES_PAD; i++) s->llen[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
        for (int i = 0; i < ZH_DCODES_PAD; i++) s->dlen[i] = 5;
        zh_codes(s->llen, ZH_LCODES_PAD, out->lcode);
        zh_codes(s->dlen, ZH_DCODES_PAD, out->dcode);
        zh_put(&bw, (uint32_t)(final_block ? 1 : 0) | (1u << 1), 3);
        out->hdr_bits = 3;
        out->body_bits = (uint32_t)D.fix;
        return;
    }
    for (int i = nl; i < ZH_LCODES_PAD; i++) s->llen[i] = 0;
    for (int i = nd; i < ZH_DCODES_PAD; i++) s->dlen[i] = 0;
    zh_codes(s->llen, ZH_LCODES_PAD, out->lcode);
    zh_codes(s->dlen, ZH_DCODES_PAD, out->dcode);
    zh_codes(s->bllen, ZH_BLCODES, s->blcode);
    zh_put(&bw, (uint32_t)(final_block ? 1 : 0) | (2u << 1), 3);
    zh_put(&bw, (uint32_t)(nl - 257), 5);
    zh_put(&bw, (uint32_t)(nd - 1), 5);
    zh_put(&bw, (uint32_t)(nbl - 4), 4);
    for (int i = 0; i < nbl; i++) zh_put(&bw, s->bllen[bl_order[i]], 3);
    for (int i = 0; i < nt; i++) {
        uint32_t nb, v = zh_tok_bits(s->tok[i], s->blcode, &nb);
        zh_put(&bw, v, (int)nb);
    }
    out->hdr_bits = bw.nbits;
    out->body_bits = (uint32_t)D.dyn;
}

/* ---- symbol format shared by the LZ77 kernel, the histogram and the encoder ----
 * literal: byte value; match: bit 31 set, (len - 3) in bits 16..23, (dist - 1) in bits 0..14 */
#define ZS_MATCH 0x80000000u
ZHD static inline uint32_t zs_match(uint32_t len, uint32_t dist) { return ZS_MATCH | ((len - 3) << 16) | (dist - 1); }

ZHD static inline int zh_msb(uint32_t v)     /* position of the highest set bit, v != 0 */
{
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    int n = 31;
    while (!((v >> n) & 1)) n--;
    return n;
#endif
}
ZHD static inline int zs_len_code(uint32_t lc /* len - 3 */)
{
    if (lc < 8) return (int)lc;
    if (lc == 255) return 28;
    int n = zh_msb(lc);                  /* msb position, >= 3 */
    return ((n - 1) << 2) | (int)((lc >> (n - 2)) & 3);
}
ZHD static inline int zs_dist_code(uint32_t d /* dist - 1 */)
{
    if (d < 4) return (int)d;
    int n = zh_msb(d);
    return (n << 1) | (int)((d >> (n - 1)) & 1);
}

/* ---- bit-offset algebra for the offset pass -------------------------------------------------
 * Every output element maps a bit position x to  (al ? roundup8(x + a) + b : x + a).
 * Elements compose associatively, so block offsets come out of one prefix scan even though
 * stored blocks, full-flush markers and trailers byte-align the stream. */
typedef struct { uint64_t a, b; uint32_t al; uint32_t pad; } zk_elem;

ZHD static inline uint64_t zk_up8(uint64_t x) { return (x + 7) & ~(uint64_t)7; }
ZHD static inline zk_elem zk_ident(void) { zk_elem e; e.a = 0; e.b = 0; e.al = 0; e.pad = 0; return e; }
ZHD static inline zk_elem zk_compose(zk_elem f, zk_elem g)   /* apply f, then g */
{
    zk_elem r; r.pad = 0;
    if (!f.al) {
        if (!g.al) { r.a = f.a + g.a; r.b = 0; r.al = 0; }
        else { r.a = f.a + g.a; r.b = g.b; r.al = 1; }
    } else {
        if (!g.al) { r.a = f.a; r.b = f.b + g.a; r.al = 1; }
        else { r.a = f.a; r.b = zk_up8(f.b + g.a) + g.b; r.al = 1; }
    }
    return r;
}
ZHD static inline uint64_t zk_apply(zk_elem f, uint64_t x) { return f.al ? zk_up8(x + f.a) + f.b : x + f.a; }

/* The element of one block, including the stream header before the first block, the full-flush
 * marker after a section's last block and the trailer after the stream's last block.
 * wrap: 0 raw, 1 zlib, 2 gzip body (raw). */
ZHD static inline zk_elem zk_elem_of_block(uint32_t type, uint32_t body_bits, uint32_t in_len,
                                           uint32_t flags, int wrap)
{
    zk_elem e = zk_ident();
    if (type == ZH_UNUSED) return e;
    if ((flags & ZB_FIRST_OF_STREAM) && wrap == 1) e.a = 16;
    zk_elem body = zk_ident();
    if (type == ZH_STORED) { body.a = 3; body.b = 32 + 8ull * in_len; body.al = 1; }
    else if (type == ZH_STORED_CONT) body.a = 8ull * in_len;      /* payload only, already byte aligned */
    else body.a = body_bits;
    e = zk_compose(e, body);
    if (flags & ZB_LAST_OF_STREAM) {
        zk_elem t = zk_ident(); t.al = 1; t.b = (wrap == 1) ? 32 : 0;
        e = zk_compose(e, t);
    } else if (flags & ZB_LAST_OF_SECTION) {
        zk_elem t = zk_ident(); t.a = 3; t.b = 32; t.al = 1;
        e = zk_compose(e, t);
    }
    return e;
}

#endif
