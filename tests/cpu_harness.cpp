// Test-only host build of the product's __host__ __device__ logic (huff_build.h, inflate_core.h)
// plus a scalar model of the LZ77 kernel's parse (deflate_lz.cu) so that `-m "not gpu"` tests can
// exercise the format logic, and `-m gpu` tests can compare the GPU's symbol stream and bytes with
// a bit-exact prediction.  Nothing here is linked into libzsc_b200.so.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "huff_build.h"
#include "inflate_core.h"
#define ZP_STATS
#include "inflate_spec.h"

extern "C" {

int h_inflate(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap, int wrap, uint32_t *res7)
{
    zi_tables T;
    zi_result r;
    zi_inflate(in, in_len, out, out_cap, wrap, &T, &r);
    res7[0] = (uint32_t)r.ret; res7[1] = (uint32_t)r.reason; res7[2] = r.produced; res7[3] = r.consumed;
    res7[4] = r.data_errors; res7[5] = r.stored_check; res7[6] = r.have_check;
    return r.ret;
}

// the warp-per-stream form of the decoder (zi_sym_batch + cooperative writes), run serially
int h_inflate_batched(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap, int wrap, uint32_t *res7, uint32_t group)
{
    zi_tables T;
    zi_result r;
    zi_inflate_batched(in, in_len, out, out_cap, wrap, &T, &r, group);
    res7[0] = (uint32_t)r.ret; res7[1] = (uint32_t)r.reason; res7[2] = r.produced; res7[3] = r.consumed;
    res7[4] = r.data_errors; res7[5] = r.stored_check; res7[6] = r.have_check;
    return r.ret;
}

// the speculative warp decoder (inflate_spec.h), lanes one after the other
int h_inflate_spec(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap, int wrap, uint32_t *res7, uint32_t opts)
{
    zi_tables T;
    zi_result r;
    zi_inflate_spec(in, in_len, out, out_cap, wrap, &T, &r, opts);
    res7[0] = (uint32_t)r.ret; res7[1] = (uint32_t)r.reason; res7[2] = r.produced; res7[3] = r.consumed;
    res7[4] = r.data_errors; res7[5] = r.stored_check; res7[6] = r.have_check;
    return r.ret;
}
// zh_lengths itself (huff_build.h): the serial code-length build that the warp forms of deflate_huff.cu must equal
int h_zh_lengths(const uint32_t *freq, int n, int maxbits, uint8_t *len_out)
{
    static zh_scratch sc;
    return zh_lengths(freq, n, maxbits, len_out, &sc);
}
// leaf depths of the (unlimited) literal/length tree, by sorted leaf: what zs_merge_kernel hands to zs_block_kernel<1>
int h_zh_leaf_depths(const uint32_t *freq, int n, uint8_t *depth_out)
{
    static zh_scratch sc;
    int max_code;
    const int m = zh_lengths_prepare(freq, n, &sc, &max_code);
    zh_sort_keys(sc.key, m);
    const int e = zh_merge(m, &sc);
    sc.depth[e - 1] = 0;
    for (int i = e - 2; i >= 0; i--) { int d = sc.depth[sc.parent[i]] + 1; sc.depth[i] = (uint8_t)(d > 250 ? 250 : d); }
    for (int i = 0; i < m; i++) depth_out[i] = sc.depth[i];
    return m;
}
void h_spec_stats(uint64_t *out8, int reset) { for (int i = 0; i < 16; i++) { out8[i] = zp_stat[i]; if (reset) zp_stat[i] = 0; } }

// ---------------------------------------------------------------- LZ77 model (mirrors deflate_lz.cu)
struct LzP { int mode, chain, nice, lazy, min_len, max_dist, good, max_lazy; };
static const uint32_t TILE = 2048, WINDOW = 32768, NOHASH = 0xFFFF;
static int g_hash_bits = 15;
static uint64_t g_chain_steps = 0;   /* candidates visited beyond the first (a proxy for the chain kernel's work) */
extern "C" uint64_t h_chain_steps(void) { uint64_t v = g_chain_steps; g_chain_steps = 0; return v; }

static inline uint32_t ld32(const uint8_t *p, uint32_t q, uint32_t q_end)
{
    uint32_t v = 0;
    for (int k = 0; k < 4; k++) v |= (uint32_t)(q + k < q_end ? p[q + k] : 0) << (8 * k);
    return v;
}
static inline uint32_t hash3(uint32_t v)
{
    uint32_t h = ((v & 0xFFFFFFu) * 2654435761u) >> (32 - g_hash_bits);
    return h == 0x7FFF ? 0x7FFE : h;       /* 0x7FFF | flag would collide with the no-hash sentinel */
}

// data: pointer such that data[q] is the byte at position q (q = 0 is the 16-byte aligned base)
static void lz_chunk(const uint8_t *data, uint32_t a, uint32_t dict_len, uint32_t len, const LzP &P, std::vector<uint32_t> &sym,
                     std::vector<uint32_t> &blk_start, uint32_t block_syms)
{
    const uint32_t q_dict = a, q_start = a + dict_len, q_end = q_start + len;
    g_hash_bits = 14;
    std::vector<uint16_t> head(32768, 0);
    std::vector<uint16_t> t_dist(TILE), t_len(TILE + 32, 0);
    uint32_t carry = q_start;
    const uint32_t t_first = P.mode == 0 ? 0 : (q_start / TILE) * TILE;
    if (len == 0) blk_start.push_back(0);
    for (uint32_t t0 = t_first; t0 < q_end; t0 += TILE) {
        if (P.mode == 0) {
            for (uint32_t g = 0; g < TILE / 32; g++) {
                uint32_t h[32], d[32];
                for (uint32_t l = 0; l < 32; l++) {
                    uint32_t q = t0 + g * 32 + l;
                    h[l] = (q >= q_dict && q + 3 <= q_end) ? hash3(ld32(data, q, q_end)) : NOHASH;
                }
                for (uint32_t l = 0; l < 32; l++) {
                    uint32_t q = t0 + g * 32 + l;
                    d[l] = 0;
                    if (h[l] == NOHASH) continue;
                    /* candidates come from the table as it stood before this group of 32 positions,
                       unless one of the ZL_NEAR preceding lanes has the same hash */
                    d[l] = (q - head[h[l]]) & 0xFFFF;
                    if (l > 0 && h[l - 1] == h[l]) d[l] = 1;
                }
                for (uint32_t l = 0; l < 32; l++) {
                    uint32_t q = t0 + g * 32 + l;
                    t_dist[g * 32 + l] = (uint16_t)d[l];
                    if (h[l] == NOHASH) continue;
                    /* the lowest lane of each hash claims the slot */
                    bool below = false;
                    for (uint32_t j = 0; j < l; j++) if (h[j] == h[l]) below = true;
                    if (!below) head[h[l]] = (uint16_t)q;
                }
            }
        }
        if (t0 + TILE <= q_start) continue;
        for (uint32_t i = 0; i < TILE; i++) {
            const uint32_t q = t0 + i;
            uint32_t best = 0, bestd = 0;
            if (P.mode != 2 && q >= q_start && q + 3 <= q_end) {
                const uint32_t maxl = q_end - q < 258 ? q_end - q : 258;
                const uint32_t maxd = q - q_dict < (uint32_t)P.max_dist ? q - q_dict : (uint32_t)P.max_dist;
                auto mlen = [&](uint32_t dd) { uint32_t n = 0; while (n < maxl && data[q + n] == data[q + n - dd]) n++; return n; };
                if (P.mode == 1) { if (maxd >= 1) { best = mlen(1); bestd = 1; } }
                else {
                    const uint32_t d = t_dist[i];
                    if (d != 0 && d <= maxd) { best = mlen(d); bestd = d; }
                }
            }
            if (best < (uint32_t)P.min_len || (best == 3 && bestd > 4096)) { best = 0; bestd = 0; }
            t_len[i] = (uint16_t)best; t_dist[i] = (uint16_t)bestd;
        }
        // parse: next(p) walk from carry
        uint32_t s = carry - t0;
        while (s < TILE) {
            uint32_t L = t_len[s];
            bool take = L >= 3;
            if (take && P.lazy && t_len[s + 1] > L) take = false;
            uint32_t q = t0 + s;
            if (q >= q_start && q < q_end) {
                if ((sym.size() % block_syms) == 0) blk_start.push_back(q - q_start);
                sym.push_back(take ? zs_match(L, t_dist[s]) : data[q]);
            }
            s += take ? L : 1;
        }
        carry = t0 + s;
    }
}


// ---------------------------------------------------------------- chain search model (mirrors deflate_chain.cu, levels 2..9)
// Tiles of ZC_TILE positions.  Hash pass: per group of 32 positions, a position's first candidate is the nearest lower
// lane of the group with the same hash, else the head-table entry as it stood before the group; the highest lane of
// each hash then holds the slot, so every occurrence stays reachable through the links.  The links of a tile are
// published when the tile is searched, so a walk from the tile may follow links of positions >= t0 + ZC_TILE - 32768.
// Search: every position compares its first candidate; then ZC_ROUNDS rounds of { parse the tile with the lengths
// known so far; positions the parse visits, and the positions right behind the matches it takes (the lazy-evaluation
// candidates), walk their chains a bounded number of steps further }, the last round to the end of the budget.
static void lz_chunk_chain(const uint8_t *data, uint32_t a, uint32_t dict_len, uint32_t len, const LzP &P, std::vector<uint32_t> &sym,
                           std::vector<uint32_t> &blk_start, uint32_t block_syms)
{
    const uint32_t CT = ZC_TILE;
    const uint32_t q_dict = a, q_start = a + dict_len, q_end = q_start + len;
    g_hash_bits = 14;
    std::vector<uint16_t> head(32768, 0), prevd(WINDOW, 0), t_cand(CT), t_len(CT + 32, 0), t_dist(CT);
    std::vector<uint32_t> best(CT), bestd(CT), cur(CT);
    std::vector<int> budget(CT);
    std::vector<uint8_t> cut(CT), mark(CT + 1);
    uint32_t carry = q_start;
    if (len == 0) blk_start.push_back(0);
    for (uint32_t t0 = 0; t0 < q_end; t0 += CT) {
        for (uint32_t g = 0; g < CT / 32; g++) {
            uint32_t h[32], d[32];
            for (uint32_t l = 0; l < 32; l++) {
                const uint32_t q = t0 + g * 32 + l;
                h[l] = (q >= q_dict && q + 3 <= q_end) ? hash3(ld32(data, q, q_end)) : NOHASH;
            }
            for (uint32_t l = 0; l < 32; l++) {
                const uint32_t q = t0 + g * 32 + l;
                d[l] = 0;
                if (h[l] == NOHASH) continue;
                d[l] = (q - head[h[l]]) & 0xFFFF;
                for (int j = (int)l - 1; j >= 0; j--) if (h[j] == h[l]) { d[l] = l - (uint32_t)j; break; }
            }
            for (uint32_t l = 0; l < 32; l++) {
                const uint32_t q = t0 + g * 32 + l;
                t_cand[g * 32 + l] = (uint16_t)d[l];
                if (h[l] == NOHASH) continue;
                bool last = true;
                for (uint32_t j = l + 1; j < 32; j++) if (h[j] == h[l]) last = false;
                if (last) head[h[l]] = (uint16_t)q;
                prevd[q & (WINDOW - 1)] = (uint16_t)d[l];
            }
        }
        if (t0 + CT <= q_start) continue;
        auto MAXL = [&](uint32_t q) { return q_end - q < 258 ? q_end - q : 258u; };
        auto MAXD = [&](uint32_t q) { return q - q_dict < (uint32_t)P.max_dist ? q - q_dict : (uint32_t)P.max_dist; };
        auto mlen = [&](uint32_t q, uint32_t dd, uint32_t ml) { uint32_t n = 0; while (n < ml && data[q + n] == data[q + n - dd]) n++; return n; };
        auto walk = [&](uint32_t i, int steps) {
            const uint32_t q = t0 + i, ml = MAXL(q), md = MAXD(q);
            uint32_t d = cur[i];
            while (d != 0 && steps-- > 0 && budget[i] > 0) {
                const uint32_t c = q - d;
                if (c + WINDOW < t0 + CT) { d = 0; break; }             /* its link has been recycled */
                const uint32_t step = prevd[c & (WINDOW - 1)];
                if (step == 0) { d = 0; break; }
                d += step;
                g_chain_steps++;
                if (d > md) { d = 0; break; }
                budget[i]--;
                const uint32_t eff = best[i] < 2 ? 2 : best[i];
                if (data[q + eff] == data[q + eff - d] && data[q + eff - 1] == data[q + eff - 1 - d]) {
                    const uint32_t n = mlen(q, d, ml);
                    if (n > best[i]) { best[i] = n; bestd[i] = d; if (n >= (uint32_t)P.nice || n >= ml) { d = 0; break; } }
                }
            }
            if (budget[i] <= 0) d = 0;
            cur[i] = d;
        };
        for (uint32_t i = 0; i < CT; i++) {
            const uint32_t q = t0 + i;
            best[i] = bestd[i] = cur[i] = 0; budget[i] = zc_budget(P.chain, len); cut[i] = 0;
            if (!(q >= q_start && q + 3 <= q_end)) continue;
            const uint32_t ml = MAXL(q), md = MAXD(q), d = t_cand[i];
            if (d != 0 && d <= md) {
                const uint32_t n = mlen(q, d, ml);
                cur[i] = d;
                if (n > 0) { best[i] = n; bestd[i] = d; if (n >= (uint32_t)P.nice || n >= ml) cur[i] = 0; }
            }
        }
        auto filt = [&]() {
            for (uint32_t i = 0; i < CT; i++) {
                uint32_t b = best[i], bd = bestd[i];
                if (b < (uint32_t)P.min_len || (b == 3 && bd > 4096)) { b = 0; bd = 0; }
                t_len[i] = (uint16_t)b; t_dist[i] = (uint16_t)bd;
            }
        };
        for (int r = 0; r < ZC_ROUNDS; r++) {
            filt();
            for (uint32_t i = 0; i <= CT; i++) mark[i] = 0;
            uint32_t s = carry - t0;
            while (s < CT) {
                const uint32_t L = t_len[s];
                bool take = L >= 3;
                if (take && P.lazy && t_len[s + 1] > L) take = false;
                mark[s] = take ? 2 : 1;
                s += take ? L : 1;
            }
            uint32_t nwalk = 0;
            for (uint32_t i = 0; i < CT; i++) {
                if (cur[i] == 0) continue;
                bool need = mark[i] != 0;
                if (i > 0 && mark[i - 1]) {
                    const uint32_t Lp = t_len[i - 1];
                    if (mark[i - 1] == 2 && P.lazy && Lp < (uint32_t)P.max_lazy) need = true;
                    if (need && Lp >= (uint32_t)P.good && !cut[i]) { budget[i] >>= 2; cut[i] = 1; }   /* "good enough" before it: a quarter of the budget */
                }
                if (!need) continue;
                nwalk++;
                walk(i, zc_round_cap(r));
            }
            if (nwalk == 0) break;
        }
        filt();
        uint32_t s = carry - t0;
        while (s < CT) {
            const uint32_t L = t_len[s];
            bool take = L >= 3;
            if (take && P.lazy && t_len[s + 1] > L) take = false;
            const uint32_t q = t0 + s;
            if (q >= q_start && q < q_end) {
                if ((sym.size() % block_syms) == 0) blk_start.push_back(q - q_start);
                sym.push_back(take ? zs_match(L, t_dist[s]) : data[q]);
            }
            s += take ? L : 1;
        }
        carry = t0 + s;
    }
}

// ---------------------------------------------------------------- encoder model (mirrors deflate_huff.cu)
struct BitW {
    std::vector<uint8_t> out; uint64_t nbits = 0;
    void put(uint64_t v, uint32_t n) {
        for (uint32_t i = 0; i < n; i++) {
            if ((nbits & 7) == 0) out.push_back(0);
            out.back() |= (uint8_t)(((v >> i) & 1) << (nbits & 7));
            nbits++;
        }
    }
    void align() { nbits = (nbits + 7) & ~7ull; }
};

static void encode_block(BitW &bw, const zh_block &B, const uint32_t *sy, const uint8_t *in_bytes, int wrap, uint32_t zhdr, uint32_t adler)
{
    if ((B.flags & ZB_FIRST_OF_STREAM) && wrap == 1) bw.put(zhdr, 16);
    if (B.type == ZH_STORED) {
        bw.put(B.hdr[0] & 7, 3); bw.align();
        bw.put(B.stored_total & 0xFFFF, 16); bw.put((~B.stored_total) & 0xFFFF, 16);
        for (uint32_t i = 0; i < B.in_len; i++) bw.put(in_bytes[i], 8);
    } else if (B.type == ZH_STORED_CONT) {
        for (uint32_t i = 0; i < B.in_len; i++) bw.put(in_bytes[i], 8);
    } else {
        for (uint32_t i = 0; i < B.hdr_bits; i++) bw.put((B.hdr[i >> 5] >> (i & 31)) & 1, 1);
        for (uint32_t i = 0; i < B.nsym; i++) {
            uint32_t s = sy[i];
            if (s & ZS_MATCH) {
                uint32_t lc = (s >> 16) & 0xFF, d = s & 0x7FFF;
                int c = zs_len_code(lc);
                uint32_t e = B.lcode[257 + c];
                bw.put(e & 0xFFFF, e >> 16);
                int eb = zh_extra_lbits(c);
                if (eb) bw.put(lc & ((1u << eb) - 1), (uint32_t)eb);
                int dc = zs_dist_code(d);
                e = B.dcode[dc];
                bw.put(e & 0xFFFF, e >> 16);
                int db = zh_extra_dbits(dc);
                if (db) bw.put(d & ((1u << db) - 1), (uint32_t)db);
            } else {
                uint32_t e = B.lcode[s & 0xFF];
                bw.put(e & 0xFFFF, e >> 16);
            }
        }
        uint32_t e = B.lcode[256];
        bw.put(e & 0xFFFF, e >> 16);
    }
    if (B.flags & ZB_LAST_OF_STREAM) {
        bw.align();
        if (wrap == 1) { bw.put(adler >> 24, 8); bw.put((adler >> 16) & 0xFF, 8); bw.put((adler >> 8) & 0xFF, 8); bw.put(adler & 0xFF, 8); }
    } else if (B.flags & ZB_LAST_OF_SECTION) {
        bw.put(0, 3); bw.align(); bw.put(0xFFFF0000u, 32);
    }
}

static uint32_t adler32_ref(const uint8_t *p, uint64_t n)
{
    uint32_t a = 1, b = 0;
    for (uint64_t i = 0; i < n; i++) { a = (a + p[i]) % 65521; b = (b + a) % 65521; }
    return (b << 16) | a;
}

// Full model of one zscgpu deflate stream.  Returns compressed size (or 0 if cap too small).
// params: [mode, chain, nice, lazy, min_len, max_dist, force_type, wrap, zhdr, good, max_lazy]
uint32_t h_deflate_model(const uint8_t *src, uint32_t n, uint32_t max_block_len, const int32_t *params,
                         uint8_t *out, uint32_t cap, uint32_t *sym_out, uint32_t sym_cap, uint32_t *nsym_out)
{
    LzP P{params[0], params[1], params[2], params[3], params[4], params[5], params[9], params[10]};
    const int force = params[6], wrap = params[7];
    const uint32_t zhdr = (uint32_t)params[8];
    const uint32_t BS = getenv("ZS_TEST_BS") ? (uint32_t)atoi(getenv("ZS_TEST_BS")) : 8192, CHUNK = 262144;
    BitW bw;
    uint32_t adler = adler32_ref(src, n);
    // positions are modelled with the source at a 16-byte aligned address
    uint32_t pos = 0, total_sym = 0;
    static zh_scratch scratch;
    do {
        uint32_t sec = n - pos < max_block_len ? n - pos : max_block_len;
        uint32_t nsub = sec ? (sec + CHUNK - 1) / CHUNK : 1;
        for (uint32_t j = 0; j < nsub; j++) {
            uint32_t coff = j * CHUNK, clen = sec - coff < CHUNK ? sec - coff : CHUNK;
            uint32_t dict = coff < WINDOW ? coff : WINDOW;
            uint32_t cstart = pos + coff;
            uint32_t a = (cstart - dict) & 15;
            std::vector<uint32_t> sym, bstart;
            if (P.mode == 0 && P.chain > 0) lz_chunk_chain(src + (cstart - dict) - a, a, dict, clen, P, sym, bstart, BS);
            else lz_chunk(src + (cstart - dict) - a, a, dict, clen, P, sym, bstart, BS);
            uint32_t nblk = (uint32_t)((sym.size() + BS - 1) / BS);
            if (nblk == 0) nblk = 1;
            std::vector<zh_block> blks(nblk);
            for (uint32_t k = 0; k < nblk; k++) {
                uint32_t cnt = (uint32_t)sym.size() - k * BS < BS ? (uint32_t)sym.size() - k * BS : BS;
                uint32_t lf[ZH_LCODES_PAD] = {0}, df[ZH_DCODES_PAD] = {0};
                for (uint32_t i = 0; i < cnt; i++) {
                    uint32_t s = sym[k * BS + i];
                    if (s & ZS_MATCH) { lf[257 + zs_len_code((s >> 16) & 0xFF)]++; df[zs_dist_code(s & 0x7FFF)]++; }
                    else lf[s & 0xFF]++;
                }
                lf[256]++;
                uint32_t in_start = bstart[k], in_end = k + 1 < nblk ? bstart[k + 1] : clen;
                uint32_t flags = 0;
                if (pos == 0 && j == 0 && k == 0) flags |= ZB_FIRST_OF_STREAM;
                if (j == nsub - 1 && k == nblk - 1) { flags |= ZB_LAST_OF_SECTION; if (pos + sec >= n) flags |= ZB_LAST_OF_STREAM; }
                zh_block &B = blks[k];
                zh_build_block(lf, df, in_end - in_start, (flags & ZB_LAST_OF_STREAM) ? 1 : 0, force, &B, &scratch);
                B.nsym = cnt; B.flags = flags; B.in_start = in_start; B.stored_total = in_end - in_start;
            }
            // merge runs of stored blocks (mirrors zs_offset_kernel)
            {
                int head = -1; uint32_t total = 0;
                for (uint32_t k = 0; k < nblk; k++) {
                    zh_block &B = blks[k];
                    if (B.type == ZH_STORED) {
                        if (head >= 0 && total + B.in_len <= 65535u) {
                            B.type = ZH_STORED_CONT; total += B.in_len; blks[head].stored_total = total;
                            if (B.flags & ZB_LAST_OF_STREAM) blks[head].hdr[0] |= 1u;
                        } else { head = (int)k; total = B.in_len; }
                    } else head = -1;
                    if (B.flags & (ZB_LAST_OF_SECTION | ZB_LAST_OF_STREAM)) head = -1;
                }
            }
            for (uint32_t k = 0; k < nblk; k++)
                encode_block(bw, blks[k], sym.data() + k * BS, src + cstart + blks[k].in_start, wrap, zhdr, adler);
            for (size_t i = 0; i < sym.size() && total_sym < sym_cap; i++) sym_out[total_sym++] = sym[i];
        }
        pos += sec;
    } while (pos < n);
    *nsym_out = total_sym;
    if (bw.out.size() > cap) return 0;
    memcpy(out, bw.out.data(), bw.out.size());
    return (uint32_t)bw.out.size();
}

// code lengths for a frequency table (property tests: Kraft equality, length limit)
int h_lengths(const uint32_t *freq, int n, int maxbits, uint8_t *len)
{
    static zh_scratch s;
    return zh_lengths(freq, n, maxbits, len, &s);
}

// the zk_elem offset algebra, for property tests
void h_zk_apply_seq(const uint32_t *types, const uint32_t *body_bits, const uint32_t *in_len, const uint32_t *flags,
                    uint32_t n, int wrap, uint64_t x0, uint64_t *offs, uint64_t *end_by_scan)
{
    zk_elem run = zk_ident();
    uint64_t x = x0;
    for (uint32_t i = 0; i < n; i++) {
        zk_elem e = zk_elem_of_block(types[i], body_bits[i], in_len[i], flags[i], wrap);
        offs[i] = x;
        x = zk_apply(e, x);
        run = zk_compose(run, e);
    }
    offs[n] = x;
    *end_by_scan = zk_apply(run, x0);
}

}  // extern "C"
