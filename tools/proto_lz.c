/* Development tool (not product, not oracle): CPU model of the group-parallel LZ77 parse used by the
 * CUDA deflate kernels, to tune compression ratio against the reference before spending GPU time.
 * Usage: proto_lz <kind> <MiB> <section_len>   (links oracle/_ref/libzsc_ref.so for the reference sizes)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>

void zscgen_fill(uint8_t *out, uint64_t n, uint64_t seed, int kind, uint64_t piece, int threads);

typedef struct {
    int hash_bytes;   /* 3 or 4 */
    int hash_bits;    /* 14/15 */
    int ways;         /* 1 or 2 */
    int lazy;         /* 0 greedy, 1 lazy one-step */
    int chain;        /* extra prev-chain depth (0 = table only) */
    int block_syms;   /* symbols per deflate block */
    int min3_maxdist; /* max distance accepted for a length-3 match */
    int group;        /* group width (32) */
    int nice;         /* stop chain when len >= nice */
} params_t;

static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

static inline uint32_t hashf(const uint8_t *p, const params_t *P)
{
    uint32_t v = ld32(p);
    if (P->hash_bytes == 3) v &= 0xFFFFFF;
    return (v * 2654435761u) >> (32 - P->hash_bits);
}

/* ---------- Huffman cost ---------- */
static const int extra_lbits[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const int extra_dbits[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static const int bl_order[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
static int length_code[259], base_len[29];
static int dist_code_tab[32769];

static void init_tabs(void)
{
    int l = 3;
    for (int c = 0; c < 28; c++) { base_len[c] = l; for (int k = 0; k < (1 << extra_lbits[c]); k++) if (l <= 258) length_code[l++] = c; }
    length_code[258] = 28; base_len[28] = 258;
    int d = 1;
    for (int c = 0; c < 30; c++) for (int k = 0; k < (1 << extra_dbits[c]); k++) if (d <= 32768) dist_code_tab[d++] = c;
}

/* optimal-ish length-limited code lengths: two-queue Huffman, then zlib-like overflow repair */
static void build_lengths(const uint32_t *freq, int n, int maxbits, uint8_t *len)
{
    int idx[600], m = 0;
    for (int i = 0; i < n; i++) { len[i] = 0; if (freq[i]) idx[m++] = i; }
    if (m == 0) return;
    if (m == 1) { len[idx[0]] = 1; return; }
    /* sort by freq (stable by symbol) */
    for (int i = 1; i < m; i++) { int v = idx[i], j = i - 1; while (j >= 0 && freq[idx[j]] > freq[v]) { idx[j + 1] = idx[j]; j--; } idx[j + 1] = v; }
    uint64_t w[1200]; int parent[1200];
    for (int i = 0; i < m; i++) w[i] = freq[idx[i]];
    int a = 0, b = m, e = m;
    while ((m - a) + (e - b) > 1) {
        int x[2];
        for (int k = 0; k < 2; k++) {
            if (a < m && (b >= e || w[a] <= w[b])) x[k] = a++; else x[k] = b++;
        }
        w[e] = w[x[0]] + w[x[1]]; parent[x[0]] = e; parent[x[1]] = e; e++;
    }
    int depth[1200]; depth[e - 1] = 0;
    for (int i = e - 2; i >= 0; i--) depth[i] = depth[parent[i]] + 1;
    int bl_count[64] = {0}, overflow = 0;
    for (int i = 0; i < m; i++) { int dl = depth[i]; if (dl > maxbits) { dl = maxbits; overflow++; } bl_count[dl]++; }
    if (overflow) {
        /* Kraft repair as in the reference's gen_bitlen (src/trees.c:474-507) */
        do {
            int bits = maxbits - 1;
            while (bl_count[bits] == 0) bits--;
            bl_count[bits]--; bl_count[bits + 1] += 2; bl_count[maxbits]--;
            overflow -= 2;
        } while (overflow > 0);
    }
    /* assign: longest lengths to least frequent */
    int i = 0;
    for (int bits = maxbits; bits >= 1; bits--) for (int k = 0; k < bl_count[bits]; k++) len[idx[i++]] = (uint8_t)bits;
}

static uint64_t block_cost(const uint32_t *lf, const uint32_t *df, uint64_t stored_len)
{
    uint8_t ll[286], dl[30];
    uint32_t lfreq[286], dfreq[30];
    memcpy(lfreq, lf, sizeof(lfreq)); memcpy(dfreq, df, sizeof(dfreq));
    lfreq[256] = 1;
    build_lengths(lfreq, 286, 15, ll);
    build_lengths(dfreq, 30, 15, dl);
    uint64_t dyn = 0, fix = 0;
    for (int i = 0; i < 286; i++) {
        int ex = i >= 257 ? extra_lbits[i - 257] : 0;
        dyn += (uint64_t)lfreq[i] * (ll[i] + ex);
        int fl = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        fix += (uint64_t)lfreq[i] * (fl + ex);
    }
    for (int i = 0; i < 30; i++) { dyn += (uint64_t)dfreq[i] * (dl[i] + extra_dbits[i]); fix += (uint64_t)dfreq[i] * (5 + extra_dbits[i]); }
    int nl = 286, nd = 30;
    while (nl > 257 && ll[nl - 1] == 0) nl--;
    while (nd > 1 && dl[nd - 1] == 0) nd--;
    uint8_t seq[320]; int ns = 0;
    for (int i = 0; i < nl; i++) seq[ns++] = ll[i];
    for (int i = 0; i < nd; i++) seq[ns++] = dl[i];
    uint32_t bf[19] = {0}; int extra = 0;
    for (int i = 0; i < ns;) {
        int j = i; while (j < ns && seq[j] == seq[i]) j++;
        int run = j - i, v = seq[i];
        if (v == 0) {
            while (run >= 11) { int r = run > 138 ? 138 : run; bf[18]++; extra += 7; run -= r; }
            if (run >= 3) { bf[17]++; extra += 3; run = 0; }
            bf[0] += run;
        } else {
            bf[v]++; run--;
            while (run >= 3) { int r = run > 6 ? 6 : run; bf[16]++; extra += 2; run -= r; }
            bf[v] += run;
        }
        i = j;
    }
    uint8_t bll[19]; build_lengths(bf, 19, 7, bll);
    int nbl = 19; while (nbl > 4 && bll[bl_order[nbl - 1]] == 0) nbl--;
    uint64_t hdr = 5 + 5 + 4 + 3 * nbl + extra;
    for (int i = 0; i < 19; i++) hdr += (uint64_t)bf[i] * bll[i];
    dyn += hdr;
    uint64_t best = dyn < fix ? dyn : fix;
    best += 3;
    uint64_t stored = (stored_len + 5) * 8 + ((stored_len / 65535) * 5 * 8);
    return best < stored ? best : stored;
}

typedef struct { uint64_t bits; uint64_t nsym; uint64_t nmatch; uint64_t matchbytes; } result_t;

static result_t compress_section(const uint8_t *in, uint32_t n, const params_t *P)
{
    result_t R = {0, 0, 0, 0};
    uint32_t hsize = 1u << P->hash_bits;
    int32_t *table = (int32_t *)malloc(sizeof(int32_t) * hsize * P->ways);
    int32_t *prev = (int32_t *)malloc(sizeof(int32_t) * (n + 64));
    for (uint32_t i = 0; i < hsize * (uint32_t)P->ways; i++) table[i] = -1;
    uint16_t *mlen = (uint16_t *)calloc(n + 64, 2);
    uint16_t *mdist = (uint16_t *)calloc(n + 64, 2);
    int G = P->group;
    /* phase 1+2: candidates and match lengths for every position (group-parallel semantics) */
    for (uint32_t p0 = 0; p0 < n; p0 += G) {
        uint32_t h[64]; int32_t c[64][2];
        int cnt = (int)(n - p0 < (uint32_t)G ? n - p0 : (uint32_t)G);
        for (int i = 0; i < cnt; i++) {
            uint32_t p = p0 + i;
            c[i][0] = c[i][1] = -1;
            if (p + 4 > n) { h[i] = 0xFFFFFFFF; continue; }
            h[i] = hashf(in + p, P);
            int k = 0;
            for (int j = i - 1; j >= 0 && k < P->ways; j--) if (h[j] == h[i]) c[i][k++] = (int32_t)(p0 + j);
            for (int w = 0; w < P->ways && k < P->ways; w++) c[i][k++] = table[h[i] * P->ways + w];
        }
        for (int i = 0; i < cnt; i++) {
            if (h[i] == 0xFFFFFFFF) continue;
            uint32_t p = p0 + i;
            prev[p] = table[h[i] * P->ways];
            for (int w = P->ways - 1; w > 0; w--) table[h[i] * P->ways + w] = table[h[i] * P->ways + w - 1];
            table[h[i] * P->ways] = (int32_t)p;
        }
        for (int i = 0; i < cnt; i++) {
            uint32_t p = p0 + i;
            uint32_t best = 0, bestd = 0;
            uint32_t maxl = n - p < 258 ? n - p : 258;
            int tried = 0;
            int32_t cc = c[i][0];
            int way = 0, chain = P->chain;
            while (cc >= 0 && p - (uint32_t)cc <= 32768) {
                uint32_t l = 0;
                while (l < maxl && in[cc + l] == in[p + l]) l++;
                if (l > best) { best = l; bestd = p - (uint32_t)cc; }
                tried++;
                if ((int)best >= P->nice) break;
                if (P->ways > 1) { way++; if (way >= P->ways) break; cc = c[i][way]; }
                else { if (chain-- <= 0) break; int32_t nx = prev[cc]; if (nx >= cc) break; cc = nx; }
            }
            if (best >= 3 && !(best == 3 && bestd > (uint32_t)P->min3_maxdist)) { mlen[p] = (uint16_t)best; mdist[p] = (uint16_t)bestd; }
        }
    }
    /* phase 3: parse */
    uint32_t lf[286], df[30];
    memset(lf, 0, sizeof(lf)); memset(df, 0, sizeof(df));
    uint32_t bsyms = 0, bstart = 0, p = 0;
    while (p < n) {
        uint32_t L = mlen[p];
        int take = L >= 3;
        if (take && P->lazy && p + 1 < n && mlen[p + 1] > L) take = 0;
        if (take) {
            lf[257 + length_code[L]]++; df[dist_code_tab[mdist[p]]]++;
            R.nmatch++; R.matchbytes += L; p += L;
        } else { lf[in[p]]++; p++; }
        bsyms++; R.nsym++;
        if (bsyms == (uint32_t)P->block_syms || p >= n) {
            R.bits += block_cost(lf, df, p - bstart);
            memset(lf, 0, sizeof(lf)); memset(df, 0, sizeof(df));
            bsyms = 0; bstart = p;
        }
    }
    R.bits += 3 + 7 + 32; /* empty stored block marker, worst-case pad */
    free(table); free(prev); free(mlen); free(mdist);
    return R;
}

typedef int (*compress2_fn)(uint8_t *, uint32_t *, const uint8_t *, uint32_t, uint32_t, uint8_t *, uint32_t, int32_t, int32_t, int32_t, int);

int main(int argc, char **argv)
{
    int kind = argc > 1 ? atoi(argv[1]) : 0;
    uint64_t mib = argc > 2 ? (uint64_t)atoi(argv[2]) : 16;
    uint32_t sect = argc > 3 ? (uint32_t)atoi(argv[3]) : 262144;
    uint64_t n = mib << 20;
    init_tabs();
    uint8_t *in = (uint8_t *)malloc(n + 64);
    zscgen_fill(in, n, kind == 1 ? 1000 : 1, kind, kind == 1 ? 262144 : (1 << 20), 8);
    void *h = dlopen("oracle/_ref/libzsc_ref.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "%s\n", dlerror()); return 1; }
    compress2_fn c2 = (compress2_fn)dlsym(h, "zsc_compress2");
    uint8_t *work = (uint8_t *)malloc(400000), *out = (uint8_t *)malloc(n + n / 4 + 1024);
    int levels[3] = {1, 6, 9};
    uint32_t refsz[3];
    for (int li = 0; li < 3; li++) {
        uint32_t dl = (uint32_t)(n + n / 4 + 1024);
        int r = c2(out, &dl, in, (uint32_t)n, sect, work, 400000, levels[li], 15, 8, 0);
        refsz[li] = dl;
        printf("ref L%d: ret %d size %u ratio %.4f\n", levels[li], r, dl, (double)n / dl);
    }
    params_t configs[] = {
        /* hb  bits ways lazy chain bsyms  min3  grp nice */
        {3, 15, 1, 0, 0, 16383, 4096, 32, 258},
        {3, 15, 1, 1, 0, 16383, 4096, 32, 258},
        {4, 15, 1, 0, 0, 16383, 4096, 32, 258},
        {4, 15, 1, 1, 0, 16383, 4096, 32, 258},
        {4, 14, 1, 1, 0, 16383, 4096, 32, 258},
        {4, 15, 2, 1, 0, 16383, 4096, 32, 258},
        {3, 15, 2, 1, 0, 16383, 4096, 32, 258},
        {4, 15, 1, 1, 1, 16383, 4096, 32, 258},
        {4, 15, 1, 1, 3, 16383, 4096, 32, 258},
        {3, 15, 1, 1, 3, 16383, 4096, 32, 258},
        {3, 15, 1, 1, 3, 8192, 4096, 32, 258},
        {3, 15, 1, 1, 7, 16383, 4096, 32, 32},
        {3, 15, 1, 1, 31, 16383, 4096, 32, 128},
        {3, 15, 1, 1, 127, 16383, 4096, 32, 128},
        {3, 15, 1, 1, 4095, 16383, 4096, 32, 258},
    };
    for (unsigned ci = 0; ci < sizeof(configs) / sizeof(configs[0]); ci++) {
        params_t *P = &configs[ci];
        uint64_t bits = 0, nsym = 0, nm = 0;
        for (uint64_t off = 0; off < n; off += sect) {
            uint32_t len = (uint32_t)(n - off < sect ? n - off : sect);
            result_t r = compress_section(in + off, len, P);
            bits += r.bits; nsym += r.nsym; nm += r.nmatch;
        }
        uint64_t bytes = bits / 8 + 6;
        printf("hb%d bits%d ways%d lazy%d chain%-4d bs%-5d nice%-3d: size %llu ratio %.4f  vsL1 %+.2f%% vsL6 %+.2f%% vsL9 %+.2f%%  sym/B %.3f\n",
               P->hash_bytes, P->hash_bits, P->ways, P->lazy, P->chain, P->block_syms, P->nice,
               (unsigned long long)bytes, (double)n / bytes,
               100.0 * ((double)bytes / refsz[0] - 1), 100.0 * ((double)bytes / refsz[1] - 1), 100.0 * ((double)bytes / refsz[2] - 1),
               (double)nsym / n);
    }
    return 0;
}
