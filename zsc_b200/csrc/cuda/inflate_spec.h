/* inflate_spec.h — speculative decoding of ONE deflate block by all the lanes of a warp (host + device).
 *
 * inflate_fast (reference src/inffast.c:76-314) is a serial loop: where a symbol starts is known only when the
 * one before it has been decoded.  One GPU lane runs that loop at a tenth of a host core's speed, which is what a
 * lone section or a narrow batch used to get (zi_fast_batch on the group's leader).  Here the other 31 lanes of
 * the warp stop waiting: a round covers ZP_NL regions of ZP_R input bits; lane k starts decoding at the first bit
 * of region k as if a literal/length code began there.  Deflate codes resynchronise: after a few symbols a lane
 * that started inside a code is, as a rule, on true symbol boundaries.  Every lane marks the bit positions at which
 * it began a literal/length code inside its own region (phase 1), then keeps decoding into the next regions until it
 * lands on a position the owner of that region marked (phase 2): from there on the two lanes would decode the same
 * symbols.  Lane 0 starts on a true boundary, so the chain lane 0 -> the lane it met -> ... is the serial decoder's
 * symbol sequence; lanes the chain skips decoded garbage, which is dropped.  What a lane decodes is a pure function
 * of the bit position (the tables are the block's), so a meeting point can never be a false one.
 *
 * The chain's symbols are then written 32 at a time (inflate_spec.inc): offsets from a prefix sum, literals at once,
 * matches in as many steps as the deepest dependency among the 32 has links — all into a ring of recent output in shared
 * memory, where a copy that reads what the copy before it wrote costs a shared-memory round trip instead of one through
 * the L2; the ring goes out to global memory in 16-byte pieces.
 *
 * Like zi_fast_batch this is a pure accelerator.  It takes only literals and valid matches whose codes the tables
 * resolve, that reach no further back than the stream allows, fit the output and end 16 bytes before the input
 * does; at anything else (end of block, a bad code or distance, the tail of a buffer) the round ends in front of that
 * symbol, or is dropped whole before a byte is written, and the generic zi_step — the code the known-answer tests
 * pin against the reference — takes over.  Results are therefore those of the one-thread decoder by construction;
 * tests/ runs zi_inflate_spec (the same lane functions, lanes one after the other) over every vector on the CPU.
 */
#if !defined(ZSC_INFLATE_SPEC_H) || defined(ZI_REINCLUDE)
#define ZSC_INFLATE_SPEC_H

#ifndef ZP_R
#define ZP_R 512u                    /* input bits per region (a multiple of 32) */
#define ZP_CAP 128u                  /* symbols a lane can hold per round */
#endif
#define ZP_RS 33u                    /* row stride of the symbol array (words): a column read hits 32 banks */
#ifndef ZP_RETRY
#define ZP_RETRY 4u                  /* fresh starts of a guessing lane that ran into something no stream has */
#endif
#ifndef ZP_OV
#define ZP_OV 0u                     /* input bits a lane decodes in front of its region, unrecorded, to find the symbol grid sooner */
#endif
#define ZP_NL 32u                    /* regions = lanes */
#ifndef ZP_LONG
#define ZP_LONG 16u                  /* matches at least this long are copied by the whole warp */
#endif
#define ZP_MARGIN 128u               /* input bits behind the round that stay with zi_step */

/* The symbols of a round: in the scratch itself (shared memory on the device), or — ZP_REC_GLOBAL, the build for wide batches —
 * in global memory behind a pointer, so that shared memory holds only tables and bitmap and more streams share an SM. */
#undef ZP_REC_ST
#undef ZP_REC_LD
#ifdef ZP_REC_GLOBAL
#define ZP_REC_ST(S, lane, i, v) ((S)->rec[(i) * ZP_RS + (lane)] = (v))
#define ZP_REC_LD(S, lane, i) ((S)->rec[(i) * ZP_RS + (lane)])
#else
#define ZP_REC_ST(S, lane, i, v) zi_sa_st32(zi_sa_of((S)->rec) + 4u * ((i) * ZP_RS + (lane)), (v))
#define ZP_REC_LD(S, lane, i) zi_sa_ld32(zi_sa_of((S)->rec) + 4u * ((i) * ZP_RS + (lane)))
#endif
typedef struct {
    uint32_t bitmap[ZP_NL * ZP_R / 32u];     /* bit b: some lane began a literal/length code b bits into the round, inside its own region */
#ifdef ZP_REC_GLOBAL
    uint32_t *rec;
#else
    uint32_t rec[ZP_CAP * ZP_RS];            /* symbol i of lane k at [i * ZP_RS + k]: literal byte, or bit 31 | (len - 3) << 16 | (dist - 1) */
#endif
} zp_scratch;

enum { ZP_RUNNING = 0, ZP_MERGED, ZP_SPAN_END, ZP_FULL, ZP_STOP };

typedef struct {
    uint64_t hold;          /* reader: bits not yet used, next word index, the word after (requested one refill ahead) */
    uint32_t bits, nw, pre;
    uint32_t rel;           /* bit position of the next symbol, relative to the round's first bit */
    uint32_t nrec, retry;
    uint32_t reason, mj, midx;      /* how the lane ended; ZP_MERGED: in region mj, on that lane's symbol midx */
    uint32_t a;             /* first symbol of the part the chain uses */
    uint32_t valid;
    uint32_t sum, reach;    /* output bytes of that part; how far its matches reach in front of its first byte */
    uint32_t start;         /* first output byte of that part */
} zp_lane;

/* what every lane of a round knows */
typedef struct {
    const uint8_t *wbase;   /* input, 4-byte aligned on the device */
    uint32_t guard;         /* host: bytes readable behind wbase */
    uint64_t bit0;          /* the round's first bit, counted from wbase */
    uint32_t nl;            /* regions in use */
    uint32_t win;           /* distance limit of the stream */
} zp_round;

ZID uint32_t zp_word(const zp_round *R, uint32_t wi)
{
#ifdef __CUDA_ARCH__
    return __ldg(reinterpret_cast<const uint32_t *>(R->wbase) + wi);
#else
    uint32_t v = 0;
    for (uint32_t b = 0; b < 4; b++) { const uint64_t i = (uint64_t)wi * 4 + b; if (i < R->guard) v |= (uint32_t)R->wbase[i] << (8 * b); }
    return v;
#endif
}

ZID void zp_lane_init(zp_lane *L, const zp_round *R, uint32_t lane)
{
    const uint32_t ov = lane ? ZP_OV : 0u;
    const uint64_t bit = R->bit0 + (uint64_t)lane * ZP_R - ov;
    const uint32_t wi = (uint32_t)(bit >> 5), sh = (uint32_t)bit & 31u;
    const uint64_t lo = zp_word(R, wi), hi = zp_word(R, wi + 1);
    L->hold = (lo | (hi << 32)) >> sh; L->bits = 64u - sh; L->nw = wi + 2; L->pre = zp_word(R, wi + 2);
    L->rel = lane * ZP_R - ov; L->nrec = 0; L->retry = 0; L->reason = ZP_RUNNING; L->mj = 0; L->midx = 0;
    L->a = 0; L->valid = 0; L->sum = 0; L->reach = 0; L->start = 0;
}

#define ZP_REFILL(L, R) do { if ((L)->bits <= 32u) { (L)->hold |= (uint64_t)(L)->pre << (L)->bits; (L)->bits += 32u; (L)->nw++; (L)->pre = zp_word((R), (L)->nw); } } while (0)

/* Phase 1 (own region, marking) or phase 2 (the regions behind, until a marked position) of one lane, at most `budget`
 * symbols of it (the lane stays ZP_RUNNING when the budget runs out). */
ZID void zp_run(zp_lane *L, const zp_round *R, const zi_tables *T, const zi_aux *X, const uint32_t *lut_len, const uint32_t *lut_dist,
                zp_scratch *S, uint32_t lane, int phase, uint32_t budget)
{
    const zi_sa lit_a = zi_sa_of(T->lit), dist_a = zi_sa_of(T->dist), pool_a = zi_sa_of(T->pool), len_a = zi_sa_of(lut_len), dl_a = zi_sa_of(lut_dist);
    const zi_sa bm_a = zi_sa_of(S->bitmap);
    if (L->reason != ZP_RUNNING) return;
    const uint32_t own_end = (lane + 1u) * ZP_R;
    for (;;) {
        const uint32_t rel = L->rel;
        bool lead_in = false;
        if (phase == 1) {
            if (rel >= own_end) return;                               /* on into phase 2 */
            if (L->nrec >= ZP_CAP) { L->reason = ZP_FULL; return; }
            lead_in = ZP_OV != 0u && rel < lane * ZP_R;               /* in front of the region: neither marked nor kept */
            if (!lead_in) {
                const zi_sa wa = bm_a + 4u * (rel >> 5);
                zi_sa_st32(wa, zi_sa_ld32(wa) | (1u << (rel & 31u)));    /* the words of a region are its lane's alone */
            }
        } else {
            const uint32_t j = rel / ZP_R;
            if (j >= R->nl) { L->reason = ZP_SPAN_END; return; }
            if ((zi_sa_ld32(bm_a + 4u * (rel >> 5)) >> (rel & 31u)) & 1u) {
                /* lane j began a symbol here: which one */
                uint32_t idx = 0;
                for (uint32_t w = j * (ZP_R / 32u); w < (rel >> 5); w++) idx += (uint32_t)zi_popc(zi_sa_ld32(bm_a + 4u * w));
                idx += (uint32_t)zi_popc(zi_sa_ld32(bm_a + 4u * (rel >> 5)) & ((1u << (rel & 31u)) - 1u));
                L->reason = ZP_MERGED; L->mj = j; L->midx = idx;
                return;
            }
            if (L->nrec >= ZP_CAP) { L->reason = ZP_FULL; return; }
        }
        ZP_REFILL(L, R);                                              /* >= 33 bits: a literal/length code and its extra bits */
        const uint64_t h = L->hold;
        const uint32_t b = L->bits;
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
                    if (l == 0 && code - first < c) { l = k; sym = X->lsorted[index + (code - first)]; }
                    index += c; first = (first + c) << 1;
                }
                if (l == 0) goto stop;
            }
        }
        uint32_t r, used;
        if (sym < 256u) { L->hold = h >> l; L->bits = b - l; r = sym; used = l; }
        else {
            if (sym == 256u || sym > 285u) goto stop;
            const uint32_t lb = zi_sa_ld32(len_a + 4u * (sym - 257u)), eb = lb >> 16;
            const uint32_t len = (lb & 0xFFFFu) + ((uint32_t)(h >> l) & ((1u << eb) - 1u));
            L->hold = h >> (l + eb); L->bits = b - (l + eb);
            ZP_REFILL(L, R);                                          /* >= 33 bits again: a distance code and its extra bits */
            const uint64_t h2 = L->hold;
            const uint32_t b2 = L->bits;
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
                        if (l2 == 0 && code - first < c) { l2 = k; d = X->dsorted[index + (code - first)]; }
                        index += c; first = (first + c) << 1;
                    }
                }
            }
            if (l2 == 0 || d > 29u) goto stop;
            const uint32_t db = zi_sa_ld32(dl_a + 4u * d), eb2 = db >> 16;
            const uint32_t dist = (db & 0xFFFFu) + ((uint32_t)(h2 >> l2) & ((1u << eb2) - 1u));
            if (dist > R->win) goto stop;
            L->hold = h2 >> (l2 + eb2); L->bits = b2 - (l2 + eb2);
            r = 0x80000000u | ((len - 3u) << 16) | (dist - 1u);
            used = l + eb + l2 + eb2;
        }
        if (!lead_in) { ZP_REC_ST(S, lane, L->nrec, r); L->nrec++; }
        L->rel = rel + used;
        if (--budget == 0) return;
        continue;
    stop:
        /* something the round does not take: the end of the block or an error if this lane is on the chain — or, in a
           lane that is still guessing, a sign that it is off the symbol grid.  Such a lane drops what it has, moves on
           one bit and tries again (the chain, should it come this way, finds no mark, decodes these symbols itself and
           stops here by itself). */
        if (phase == 1 && lane != 0u && L->retry < ZP_RETRY && rel + 1u < own_end) {
            L->retry++;
            for (uint32_t wd = lane * (ZP_R / 32u); wd < (lane + 1u) * (ZP_R / 32u); wd++) zi_sa_st32(bm_a + 4u * wd, 0u);
            const uint64_t bit = R->bit0 + rel + 1u;
            const uint32_t wi = (uint32_t)(bit >> 5), sh = (uint32_t)bit & 31u;
            const uint64_t lo = zp_word(R, wi), hi = zp_word(R, wi + 1);
            L->hold = (lo | (hi << 32)) >> sh; L->bits = 64u - sh; L->nw = wi + 2; L->pre = zp_word(R, wi + 2);
            L->rel = rel + 1u; L->nrec = 0;
            continue;
        }
        L->reason = ZP_STOP;
        return;
    }
}

/* output bytes of the symbols [a, nrec) of a lane, and how far their matches reach in front of the first of them */
ZID void zp_measure(zp_lane *L, const zp_scratch *S, uint32_t lane)
{
    uint32_t sum = 0, reach = 0;
    for (uint32_t i = L->a; i < L->nrec; i++) {
        const uint32_t r = ZP_REC_LD(S, lane, i);
        if (r >> 31) {
            const uint32_t len = ((r >> 16) & 0xFFu) + 3u, dist = (r & 0x7FFFu) + 1u;
            if (dist > sum && dist - sum > reach) reach = dist - sum;
            sum += len;
        } else sum++;
    }
    L->sum = sum; L->reach = reach;
}

/* ---- one round, lanes one after the other (host; the device form is zp_round_warp in inflate_spec.inc) ----
 * Returns the number of symbols taken (m's cursor and output position advanced); *stopped = the chain ended in front of
 * a symbol the round does not take. */
#ifndef __CUDA_ARCH__
#ifdef ZP_STATS
static uint64_t zp_stat[16];         /* rounds, rounds dropped, symbols, lanes used, waves, regions offered, long copies */
#endif
static inline uint32_t zp_round_host(zi_mach *m, const uint32_t *lut_len, const uint32_t *lut_dist, zp_scratch *S, uint32_t *stopped)
{
    zp_round R;
    zp_lane L[ZP_NL];
    *stopped = 0;
    const uint64_t pos0 = (uint64_t)m->io.ip * 8 - m->io.bits, end = (uint64_t)m->io.in_len * 8;
    if (pos0 + ZP_MARGIN + 2u * ZP_R > end) return 0;
    uint64_t nl64 = (end - pos0 - ZP_MARGIN) / ZP_R;
    R.nl = nl64 > ZP_NL ? ZP_NL : (uint32_t)nl64;
    R.wbase = m->io.in; R.guard = m->io.in_len; R.bit0 = pos0; R.win = m->win;
    const uint32_t op0 = m->io.op, room = m->io.out_cap - op0, floor_ = m->base - m->hist;
    const int count_only = (m->opts & ZI_OPT_COUNT_ONLY) != 0;
    for (uint32_t w = 0; w < ZP_NL * ZP_R / 32u; w++) S->bitmap[w] = 0;
    for (uint32_t k = 0; k < R.nl; k++) { zp_lane_init(&L[k], &R, k); zp_run(&L[k], &R, m->T, m->X, lut_len, lut_dist, S, k, 1, 0xFFFFFFFFu); }
    for (uint32_t k = 0; k < R.nl; k++) zp_run(&L[k], &R, m->T, m->X, lut_len, lut_dist, S, k, 2, 0xFFFFFFFFu);
    /* the chain */
    uint32_t c = 0, last;
    L[0].valid = 1; L[0].a = 0;
    for (;;) {
        last = c;
        if (L[c].reason != ZP_MERGED) break;
        const uint32_t j = L[c].mj;
        L[j].valid = 1; L[j].a = L[c].midx;
        c = j;
    }
    uint32_t total = 0, nsym = 0, bad = 0;
    for (uint32_t k = 0; k < R.nl; k++) {
        if (!L[k].valid) continue;
        zp_measure(&L[k], S, k);
        if (total + L[k].sum > room) {
            /* the output ends inside this round: the chain is cut behind the last lane whose symbols still fit */
            if (k == 0) bad = 1;
            for (uint32_t j = k; j < R.nl; j++) L[j].valid = 0;
            break;
        }
        last = k;
        L[k].start = op0 + total;
        if (L[k].reach > (uint32_t)(L[k].start - floor_)) bad = 1;
        total += L[k].sum; nsym += L[k].nrec - L[k].a;            /* (sums stay far below 2^32: 32 lanes x 96 symbols x 258) */
    }
#ifdef ZP_STATS
    zp_stat[0]++; zp_stat[5] += R.nl;
#endif
    if (bad || nsym == 0) {
#ifdef ZP_STATS
        zp_stat[1]++;
#endif
        return 0;
    }
    /* the chain's symbols in order (the device writes them 32 at a time through a ring in shared memory) */
    if (!count_only) for (uint32_t k = 0; k < R.nl; k++) {
        if (!L[k].valid) continue;
        uint32_t p = L[k].start;
        for (uint32_t i = L[k].a; i < L[k].nrec; i++) {
            const uint32_t r = S->rec[i * ZP_RS + k];
            if (r >> 31) { const uint32_t len = ((r >> 16) & 0xFFu) + 3u, dist = (r & 0x7FFFu) + 1u; for (uint32_t q = 0; q < len; q++, p++) m->io.out[p] = m->io.out[p - dist]; }
            else m->io.out[p++] = (uint8_t)r;
        }
    }
#ifdef ZP_STATS
    zp_stat[2] += nsym;
    for (uint32_t k = 0; k < R.nl; k++) zp_stat[3] += L[k].valid;
    zp_stat[8 + L[last].reason]++;
    for (uint32_t k = 0; k < R.nl; k++) if (L[k].valid && L[k].reason == ZP_MERGED && L[k].mj != k + 1) zp_stat[13]++;
    for (uint32_t k = 0; k < R.nl; k++) if (L[k].valid && L[k].reason == ZP_MERGED) zp_stat[14]++;
#endif
    /* the cursor behind the last symbol taken */
    const uint64_t pe = pos0 + L[last].rel;
    zi_seek(&m->io, (uint32_t)(pe >> 3));
    zi_refill(&m->io);
    zi_drop(&m->io, (int)(pe & 7u));
    m->io.op = op0 + total;
    *stopped = L[last].reason == ZP_STOP;
    return nsym;
}

/* Whole stream, the way zs_inflate_spec_kernel runs it: speculative rounds inside compressed blocks, zi_fast_batch and
 * zi_step for everything a round leaves. */
static inline void zi_inflate_spec(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap, int wrap, zi_tables *T, zi_result *res, uint32_t opts)
{
    zi_mach m;
    zi_aux X;
    uint32_t q[ZI_BATCH], lut_len[32], lut_dist[32];
    zp_scratch *S = (zp_scratch *)malloc(sizeof(zp_scratch));
    for (uint32_t c = 0; c < 29; c++) lut_len[c] = zi_lut_len(c);
    for (uint32_t d = 0; d < 30; d++) lut_dist[d] = zi_lut_dist(d);
    zi_m_init(&m, in, in_len, out, out_cap, wrap, T, &X);
    m.opts = opts;
    while (m.state != ZM_DONE) {
        if (m.state == ZM_SYM) {
            uint32_t stopped = 0;
            const uint32_t took = zp_round_host(&m, lut_len, lut_dist, S, &stopped);
            if (took && !stopped) continue;
            uint32_t vop = 0, p = m.io.op;
            const uint32_t n = zi_fast_batch(&m, lut_len, lut_dist, q, 32, &vop);
            if (!(opts & ZI_OPT_COUNT_ONLY)) for (uint32_t i = 0; i < n; i++) {
                const uint32_t r = q[i];
                if (r >> 31) { const uint32_t len = ((r >> 16) & 0xFF) + 3, dist = (r & 0x7FFF) + 1; for (uint32_t k = 0; k < len; k++, p++) out[p] = out[p - dist]; }
                else out[p++] = (uint8_t)r;
            }
            if (n) m.io.op = vop;
            if (n < 32) zi_step(&m);
        } else if (m.state == ZM_STORED) {
            const uint32_t n = zi_stored_plan(&m);
            if (!(opts & ZI_OPT_COUNT_ONLY)) for (uint32_t k = 0; k < n; k++) out[m.io.op + k] = in[m.io.ip + k];
            zi_stored_done(&m, n);
        } else zi_step(&m);
    }
    free(S);
    *res = m.res;
}
#endif

#endif
