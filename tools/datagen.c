/* Synthetic workload generators for the BASELINE.json configs.
 *
 * Test/bench infrastructure only; not part of the product library.  Everything is deterministic
 * in (seed, length): PRNG = xorshift64* seeded through splitmix64.  Large buffers are generated in
 * independent 1 MiB pieces (piece k uses seed splitmix(seed + k)) so that pthreads can fill them
 * in parallel and so that any sub-range can be regenerated without the rest.
 *
 *   text      words drawn Zipf(s = 1.1) from a 4096-word vocabulary of random lowercase strings
 *             of length 2..12, separated by " " (80 %), ", " (10 %) or ".\n" (10 %)
 *   telemetry 64-byte records: u32 sequence counter, u32 ms timestamp (+9..11), 8 x i16 channels
 *             as bounded random walks (step -3..3), 4 status bytes from a 6-symbol alphabet with
 *             0.95 self-transition, 16 zero bytes, 8-byte ASCII tag cycling over 16 tags,
 *             u16 checksum of the record, u16 pad
 *   mixed     alternating text / telemetry(+5 % uniform byte noise) segments, length uniform in
 *             [4 KiB, 64 KiB]
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } rng_t;

static uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static void rng_seed(rng_t *r, uint64_t seed) { r->s = splitmix64(seed); if (!r->s) r->s = 1; }
static inline uint64_t rng_next(rng_t *r)
{
    uint64_t x = r->s;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    r->s = x;
    return x * 0x2545F4914F6CDD1Dull;
}
static inline uint32_t rng_u32(rng_t *r) { return (uint32_t)(rng_next(r) >> 32); }
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)(((uint64_t)rng_u32(r) * n) >> 32); }

/* ---- vocabulary (fixed for all seeds, built once) ---- */
#define VOCAB 4096
static char vocab_word[VOCAB][13];
static uint8_t vocab_len[VOCAB];
static uint32_t vocab_cdf[VOCAB];
static pthread_once_t vocab_once = PTHREAD_ONCE_INIT;

static void vocab_init(void)
{
    rng_t r; rng_seed(&r, 0x5EEDF00Dull);
    double tot = 0.0, acc = 0.0;
    for (int i = 0; i < VOCAB; i++) tot += 1.0 / pow((double)(i + 1), 1.1);
    for (int i = 0; i < VOCAB; i++) {
        int len = 2 + (int)rng_below(&r, 11);
        for (int k = 0; k < len; k++) vocab_word[i][k] = (char)('a' + rng_below(&r, 26));
        vocab_word[i][len] = 0;
        vocab_len[i] = (uint8_t)len;
        acc += 1.0 / pow((double)(i + 1), 1.1);
        double c = acc / tot * 4294967295.0;
        vocab_cdf[i] = (i == VOCAB - 1) ? 0xFFFFFFFFu : (uint32_t)c;
    }
}

static size_t gen_text(uint8_t *out, size_t n, rng_t *r)
{
    size_t o = 0;
    while (o < n) {
        uint32_t u = rng_u32(r);
        int lo = 0, hi = VOCAB - 1;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (vocab_cdf[mid] < u) lo = mid + 1; else hi = mid; }
        int len = vocab_len[lo];
        for (int k = 0; k < len && o < n; k++) out[o++] = (uint8_t)vocab_word[lo][k];
        uint32_t sep = rng_below(r, 10);
        if (sep < 8) { if (o < n) out[o++] = ' '; }
        else if (sep == 8) { if (o < n) out[o++] = ','; if (o < n) out[o++] = ' '; }
        else { if (o < n) out[o++] = '.'; if (o < n) out[o++] = '\n'; }
    }
    return o;
}

typedef struct {
    uint32_t seq, ts; int16_t ch[8]; uint8_t st[4]; uint32_t tag;
} telem_t;

static const char telem_tags[16][9] = {
    "ATT_CTRL", "PWR_BUS1", "PWR_BUS2", "THERM_A0", "THERM_B1", "RW_SPEED", "GYRO_XYZ", "STAR_TRK",
    "COMM_SBD", "COMM_XBD", "PROP_TNK", "PAYLD_01", "PAYLD_02", "FSW_HLTH", "MEM_SCRB", "TIME_SVC" };

static void telem_init(telem_t *t, rng_t *r)
{
    t->seq = rng_u32(r) & 0xFFFFF; t->ts = rng_u32(r) & 0x3FFFFFFF;
    for (int i = 0; i < 8; i++) t->ch[i] = (int16_t)((int)rng_below(r, 4001) - 2000);
    for (int i = 0; i < 4; i++) t->st[i] = (uint8_t)rng_below(r, 6);
    t->tag = rng_below(r, 16);
}

static void telem_record(uint8_t *rec, telem_t *t, rng_t *r)
{
    static const uint8_t st_alpha[6] = { 0x00, 0x01, 0x03, 0x10, 0x80, 0xFF };
    t->seq++; t->ts += 9 + rng_below(r, 3);
    memcpy(rec, &t->seq, 4); memcpy(rec + 4, &t->ts, 4);
    for (int i = 0; i < 8; i++) {
        int v = t->ch[i] + (int)rng_below(r, 7) - 3;
        if (v > 8000) v = 8000;
        if (v < -8000) v = -8000;
        t->ch[i] = (int16_t)v;
        memcpy(rec + 8 + 2 * i, &t->ch[i], 2);
    }
    for (int i = 0; i < 4; i++) {
        if (rng_below(r, 100) >= 95) t->st[i] = (uint8_t)rng_below(r, 6);
        rec[24 + i] = st_alpha[t->st[i]];
    }
    memset(rec + 28, 0, 16);
    memcpy(rec + 44, telem_tags[t->tag], 8);
    t->tag = (t->tag + 1) & 15;
    uint16_t ck = 0;
    for (int i = 0; i < 52; i++) ck = (uint16_t)(ck + rec[i]);
    memcpy(rec + 52, &ck, 2);
    rec[54] = 0; rec[55] = 0;
    /* bytes 56..63: a second small status block so the record is 64 B: 2 counters + zero pad */
    uint16_t sub = (uint16_t)(t->seq & 0xFF);
    memcpy(rec + 56, &sub, 2); memset(rec + 58, 0, 6);
}

static size_t gen_telemetry(uint8_t *out, size_t n, rng_t *r, int noise_pct)
{
    telem_t t; telem_init(&t, r);
    uint8_t rec[64];
    size_t o = 0;
    while (o < n) {
        telem_record(rec, &t, r);
        if (noise_pct > 0)
            for (int i = 0; i < 64; i++)
                if (rng_below(r, 100) < (uint32_t)noise_pct) rec[i] = (uint8_t)rng_u32(r);
        size_t c = n - o < 64 ? n - o : 64;
        memcpy(out + o, rec, c);
        o += c;
    }
    return o;
}

static void gen_mixed_piece(uint8_t *out, size_t n, uint64_t seed, int start_text)
{
    rng_t r; rng_seed(&r, seed);
    size_t o = 0; int text = start_text;
    while (o < n) {
        size_t seg = 4096 + rng_below(&r, 61441);
        if (seg > n - o) seg = n - o;
        if (text) gen_text(out + o, seg, &r); else gen_telemetry(out + o, seg, &r, 5);
        o += seg; text = !text;
    }
}

/* kind: 0 mixed, 1 telemetry (per-buffer, see zscgen_telemetry_buffers), 2 text, 3 uniform random */
typedef struct { uint8_t *out; uint64_t n; uint64_t seed; int kind; uint64_t piece; volatile uint64_t *next; uint64_t npieces; } gjob_t;

static void gen_piece(uint8_t *out, uint64_t n, uint64_t seed, int kind, uint64_t k)
{
    rng_t r;
    switch (kind) {
    case 0: gen_mixed_piece(out, n, seed + 0x100000000ull * (k + 1), (int)(k & 1)); break;
    case 1: rng_seed(&r, seed + k); gen_telemetry(out, n, &r, 0); break;
    case 2: rng_seed(&r, seed + 0x100000000ull * (k + 1)); gen_text(out, n, &r); break;
    default:
        rng_seed(&r, seed + 0x100000000ull * (k + 1));
        { uint64_t o = 0; for (; o + 8 <= n; o += 8) { uint64_t v = rng_next(&r); memcpy(out + o, &v, 8); }
          for (; o < n; o++) out[o] = (uint8_t)rng_u32(&r); }
    }
}

static void *gworker(void *arg)
{
    gjob_t *j = (gjob_t *)arg;
    for (;;) {
        uint64_t k = __sync_fetch_and_add(j->next, 1);
        if (k >= j->npieces) break;
        uint64_t off = k * j->piece;
        uint64_t len = j->n - off < j->piece ? j->n - off : j->piece;
        gen_piece(j->out + off, len, j->seed, j->kind, k);
    }
    return 0;
}

/* Fill out[0..n) with workload `kind`, in pieces of `piece` bytes (piece k seeded from seed and k).
 * For kind 1 (telemetry) piece = the buffer size and piece k uses seed + k, which is config 3's
 * "seed = 1000 + idx" when seed = 1000. */
void zscgen_fill(uint8_t *out, uint64_t n, uint64_t seed, int kind, uint64_t piece, int threads)
{
    pthread_once(&vocab_once, vocab_init);
    if (piece == 0) piece = 1u << 20;
    volatile uint64_t next = 0;
    gjob_t j = { out, n, seed, kind, piece, &next, (n + piece - 1) / piece };
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int t = 0; t < threads; t++) pthread_create(&th[t], 0, gworker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
}
