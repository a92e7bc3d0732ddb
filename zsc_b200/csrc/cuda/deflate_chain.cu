/* deflate_chain.cu — LZ77 match finding with hash chains + lazy parse (levels 2..9), one CTA per chunk (sm_100a).
 *
 * Takes the place of the reference's deflate_slow / longest_match / INSERT_STRING chains (reference
 * src/deflate.c:1400-1518, :1989-2122) for a whole batch of independent chunks.  Not a port: a serial deflate only
 * searches where its parse arrives, and that dependence is what this kernel re-creates in parallel form.
 *
 *   - the chunk streams through a 64 KiB shared-memory ring (32 KiB window + the tiles in flight), staged by
 *     cp.async.bulk copies that one thread queues and an mbarrier completes (lz_common.cuh);
 *   - 16 worker warps hash the coming tile (3-byte hash; __match_any_sync finds, per group of 32 positions, the nearest
 *     lower lane with the same hash and the highest lane of every hash), the hasher warp walks the head table one tile
 *     ahead: first candidate of a position = that nearest lower lane, else the table entry; the highest lane of each
 *     hash takes the slot, so every occurrence stays reachable through the link array;
 *   - a tile of ZC_TILE positions is then searched in ZC_ROUNDS rounds.  Every position compares its first candidate.
 *     A round parses the tile (greedy / one-step lazy, pointer doubling as in deflate_lz.cu) with the lengths known so
 *     far, collects the positions the parse visits plus the positions right behind the matches it takes (the
 *     lazy-evaluation candidates, unless the match is already max_lazy long) into a work list, and the 512 worker
 *     threads follow the chains of the listed positions zc_round_cap(r) candidates further (a quarter of the budget
 *     behind a match that is already `good`).  Positions the parse never visits are never searched deeper — on the
 *     BASELINE workloads a third of the positions walk at all, 7-9 candidates per input byte at level 6 instead of
 *     37-44 when every position walks, at a smaller output size;
 *   - the final parse emits 32-bit symbols exactly like deflate_lz.cu.
 *
 * The reference's budget counts only candidates that pass its quick reject (the MISRA rewrite of longest_match moved
 * the decrement behind the `continue`, src/deflate.c:1462-1469,1505), so its "128" reaches much deeper than stock
 * zlib's; the budgets in engine.cu:zs_lz_params count every candidate and are sized so that the output stays within
 * 1 % of the reference's at every level (tests/test_gpu.py ratio gates).
 *
 * Candidates are verified byte for byte, so stale or aliased table entries cost ratio, never correctness; every step
 * is deterministic and tests/cpu_harness.cpp:lz_chunk_chain predicts the symbol stream bit for bit.
 */
#include "lz_common.cuh"

#define ZC_GROUPS (ZC_TILE / 32)
#define ZC_GPW (ZC_GROUPS / ZL_WORKER_WARPS)       /* consecutive groups per worker warp (one block) */
#define ZC_BLOCK (ZC_GPW * 32)
#define ZC_RING 65536u
#define ZC_HASH_BITS 14u
#define ZC_NOHASH 0xFFFFFFFFu

struct ZcSmem {
    uint32_t ring32[ZC_RING / 4 + ZL_MIRROR / 4];
    uint16_t head[1u << ZC_HASH_BITS];
    uint16_t prevd[ZS_WINDOW];        /* distance from a position to the previous one with the same hash (0 = none) */
    uint32_t t_hash[2][ZC_TILE];      /* workers -> hasher: hash | lower-lane distance << 16 | highest-lane flag << 21 */
    uint16_t t_cand[2][ZC_TILE];      /* hasher -> workers: first-candidate distance */
    uint16_t best[ZC_TILE + 32];      /* longest match found so far (+ zero sentinel) */
    uint16_t bestd[ZC_TILE];
    uint16_t cur[ZC_TILE];            /* distance of the last candidate examined; 0 = the walk is over */
    int16_t budget[ZC_TILE];
    uint16_t t_exit[ZC_TILE];
    uint16_t t_bexit[ZC_TILE];
    uint16_t list[ZC_TILE];           /* positions that walk in this round */
    uint32_t m_any[ZC_GROUPS];        /* parse starts of the round, one bit per position */
    uint32_t m_take[ZC_GROUPS];       /* ... that take a match */
    uint32_t cutbits[ZC_GROUPS];      /* budget already quartered */
    uint32_t g_cnt[ZC_GROUPS];
    uint32_t g_off[ZC_GROUPS];
    uint16_t b_entry[ZL_WORKER_WARPS];
    uint32_t carry, nsym, list_n;
    unsigned long long stage_bar;
};

/* match length a parse may use: the kept length, or 0 (too short / a 3-byte match too far away) */
__device__ __forceinline__ uint32_t zc_flen(const ZcSmem &S, uint32_t i, uint32_t min_len)
{
    uint32_t L = S.best[i];
    if (L < min_len || (L == 3 && S.bestd[i] > ZS_TOO_FAR)) L = 0;
    return L;
}

/* workers: hashes of one tile.  A warp covers one group of 32 consecutive positions per step. */
__device__ __forceinline__ void zc_hash_tile(const uint32_t *ring32, uint32_t *t_hash, uint32_t t0, uint32_t q_dict, uint32_t q_end, uint32_t wtid)
{
    const uint32_t lane = wtid & 31, lt = zs_lanemask_lt(), gt = zs_lanemask_gt();
    for (uint32_t i = wtid; i < ZC_TILE; i += ZL_WORKERS) {
        const uint32_t q = t0 + i;
        uint32_t h = ZC_NOHASH;
        if (q >= q_dict && q + 3 <= q_end) h = zl_hash<ZC_HASH_BITS>(zl_ld32<ZC_RING>(ring32, q));
        const uint32_t m = __match_any_sync(0xFFFFFFFFu, h);
        const uint32_t lower = m & lt;
        const uint32_t dl = lower ? lane - (31u - (uint32_t)__clz((int)lower)) : 0u;
        const uint32_t last = (m & gt) == 0 ? 1u : 0u;
        t_hash[i] = (h == ZC_NOHASH) ? ZC_NOHASH : (h | (dl << 16) | (last << 21));
    }
}

/* hasher warp: head-table pass over one tile, groups of 32 positions in order */
__device__ __forceinline__ void zc_hasher_tile(ZcSmem &S, const uint32_t *t_hash, uint16_t *t_cand, uint32_t t0, uint32_t lane)
{
#pragma unroll 4
    for (uint32_t g = 0; g < ZC_GROUPS; g++) {
        const uint32_t i = g * 32 + lane, q = t0 + i;
        const uint32_t v = t_hash[i];
        const bool valid = (v != ZC_NOHASH);
        const uint32_t hs = valid ? (v & 0x3FFFu) : 0u;
        const uint32_t old = S.head[hs];                                   /* table as it stood before the group */
        if (valid && ((v >> 21) & 1u)) S.head[hs] = (uint16_t)q;           /* the highest lane of each hash takes it */
        const uint32_t dl = (v >> 16) & 31u;
        t_cand[i] = (uint16_t)(valid ? (dl ? dl : ((q - old) & 0xFFFFu)) : 0u);
        __syncwarp();
    }
}

/* Parse of the tile from S.carry with the lengths known now (the E phases of deflate_lz.cu).
 * Fills jn (symbol length per position of this warp's groups) and marks (parse starts, per group, lane-uniform). */
__device__ __forceinline__ void zc_parse(ZcSmem &S, const ZsLzParams &P, uint32_t t0, uint32_t wtid, uint32_t (&jn)[ZC_GPW], uint32_t (&marks)[ZC_GPW])
{
    const uint32_t lane = wtid & 31, ww = wtid >> 5;
    const uint32_t b0 = ww * ZC_BLOCK;
    /* E1: exit function per group (pointer doubling), composed per block */
#pragma unroll
    for (int g = 0; g < ZC_GPW; g++) {
        const uint32_t i = b0 + g * 32 + lane;
        const uint32_t L = zc_flen(S, i, (uint32_t)P.min_len);
        uint32_t Ln = __shfl_down_sync(0xFFFFFFFFu, L, 1);
        if (lane == 31) Ln = zc_flen(S, i + 1, (uint32_t)P.min_len);      /* the sentinel behind the tile is 0 */
        bool take = L >= 3;
        if (take && P.lazy && Ln > L) take = false;
        const uint32_t n = take ? L : 1u;
        jn[g] = n;
        uint32_t j = lane + n;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const uint32_t jj = __shfl_sync(0xFFFFFFFFu, j, j & 31);
            if (j < 32) j = jj;
        }
        S.t_exit[i] = (uint16_t)(b0 + g * 32 + j);
    }
    __syncwarp();
#pragma unroll
    for (int g = 0; g < ZC_GPW; g++) {
        const uint32_t i = b0 + g * 32 + lane;
        uint32_t e = S.t_exit[i];
#pragma unroll
        for (int hop = g + 1; hop < ZC_GPW; hop++) if (e < b0 + ZC_BLOCK) e = S.t_exit[e];
        S.t_bexit[i] = (uint16_t)e;
    }
    zl_bar_workers();
    /* E2: hop from block to block */
    if (wtid == 0) {
        uint32_t s = S.carry - t0;
#pragma unroll 4
        for (uint32_t b = 0; b < ZL_WORKER_WARPS; b++) {
            uint32_t e = ZL_NONE;
            if (s < (b + 1) * ZC_BLOCK) { e = s; s = S.t_bexit[s]; }
            S.b_entry[b] = (uint16_t)e;
        }
        S.list_n = 0;
        S.g_cnt[0] = s;                                   /* where the parse leaves the tile (tile relative) */
    }
    zl_bar_workers();
    /* E3: mark parse starts */
    uint32_t s_in = S.b_entry[ww];
#pragma unroll
    for (int g = 0; g < ZC_GPW; g++) {
        uint32_t m = 0;
        if (s_in != ZL_NONE && s_in < b0 + (g + 1) * 32) { m = 1u << (s_in - (b0 + g * 32)); s_in = S.t_exit[s_in]; }
        uint32_t j = lane + jn[g];
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const uint32_t contrib = (((m >> lane) & 1u) && j < 32) ? (1u << j) : 0u;
            m |= __reduce_or_sync(0xFFFFFFFFu, contrib);
            const uint32_t jj = __shfl_sync(0xFFFFFFFFu, j, j & 31);
            if (j < 32) j = jj;
        }
        marks[g] = m;
    }
}

/* Follow the chain of tile position i up to `cap` candidates further.  One step = one candidate: its link and its
 * reject byte are requested together (one shared-memory round trip on the chain); the full comparison runs only when
 * the two bytes behind the best match agree.  The listed positions are dealt out statically, one walk per thread and
 * pass: handing them out dynamically (a lane takes the next position when its walk ends) and a burst loop over the
 * active lanes measured 20-25 % slower at every budget (profiles/r02e_chain_variants.log) — the walks that decide a
 * round's duration are the few that run to the end of the budget, and those are on the critical path either way. */
__device__ __forceinline__ void zc_walk_one(ZcSmem &S, const ZsLzParams &P, uint32_t i, uint32_t t0, int cap, uint32_t q_dict, uint32_t q_end)
{
    const uint32_t q = t0 + i;
    const uint32_t ml = min(ZS_MAX_MATCH, q_end - q);
    const uint32_t md = min((uint32_t)P.max_dist, q - q_dict);
    const uint32_t lim = i + ZS_WINDOW - ZC_TILE;
    uint32_t d = S.cur[i], best = S.best[i], bestd = S.bestd[i];
    int bud = S.budget[i];
    int steps = cap < bud ? cap : bud;
    bud -= steps;
    uint32_t eff = best < 2 ? 2u : best;
    uint32_t tail = zl_ld8<ZC_RING>(S.ring32, q + eff), tail1 = zl_ld8<ZC_RING>(S.ring32, q + eff - 1);
    uint32_t step = (d != 0 && d <= lim) ? (uint32_t)S.prevd[(q - d) & (ZS_WINDOW - 1)] : 0u;
    while (steps > 0) {
        if (step == 0) { d = 0; break; }
        d += step;
        if (d > md) { d = 0; break; }
        steps--;
        step = d <= lim ? (uint32_t)S.prevd[(q - d) & (ZS_WINDOW - 1)] : 0u;
        const uint32_t rb = zl_ld8<ZC_RING>(S.ring32, q + eff - d);
        if (rb == tail && zl_ld8<ZC_RING>(S.ring32, q + eff - 1 - d) == tail1) {
            const uint32_t n = zl_match_len<ZC_RING>(S.ring32, q, d, ml);
            if (n > best) {
                best = n; bestd = d;
                if (n >= (uint32_t)P.nice || n >= ml) { d = 0; break; }
                eff = best < 2 ? 2u : best;
                tail = zl_ld8<ZC_RING>(S.ring32, q + eff); tail1 = zl_ld8<ZC_RING>(S.ring32, q + eff - 1);
            }
        }
    }
    bud += steps;
    if (bud <= 0) d = 0;
    S.cur[i] = (uint16_t)d; S.budget[i] = (int16_t)bud; S.best[i] = (uint16_t)best; S.bestd[i] = (uint16_t)bestd;
}

__global__ void __launch_bounds__(ZL_THREADS, 1)
zs_lzc_kernel(const uint8_t *__restrict__ raw, const ZsChunk *__restrict__ chunks,
              uint32_t *__restrict__ sym, uint32_t *__restrict__ chunk_nsym,
              uint32_t *__restrict__ blk_in_start, ZsLzParams P)
{
    extern __shared__ __align__(16) unsigned char zc_smem_raw[];
    ZcSmem &S = *reinterpret_cast<ZcSmem *>(zc_smem_raw);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_hasher = (warp == 0);
    const uint32_t wtid = tid - 32, ww = warp - 1;                 /* worker thread / warp index */
    const ZsChunk cd = chunks[blockIdx.x];
    const uint64_t src_addr = (uint64_t)(raw + cd.raw_off) - cd.dict_len;
    const uint32_t a = (uint32_t)(src_addr & 15);
    const uint8_t *gbase = (const uint8_t *)(src_addr - a);       /* q = 0 */
    const uint32_t q_dict = a, q_start = a + cd.dict_len, q_end = q_start + cd.len;
    const uint32_t q_end16 = (q_end + 15u) & ~15u;
    uint32_t *out_sym = sym + cd.sym_off;

    for (uint32_t i = tid; i < (1u << ZC_HASH_BITS) / 2; i += ZL_THREADS) ((uint32_t *)S.head)[i] = 0;
    for (uint32_t i = tid; i < ZS_WINDOW / 2; i += ZL_THREADS) ((uint32_t *)S.prevd)[i] = 0;
    for (uint32_t i = tid; i < ZC_TILE + 32; i += ZL_THREADS) S.best[i] = 0;
    if (tid == 0) { S.carry = q_start; S.nsym = 0; S.list_n = 0; if (cd.len == 0) blk_in_start[cd.blk_base] = 0; zl_mbar_init(&S.stage_bar, 1); }
    __syncthreads();

    const uint32_t ntiles = (q_end + ZC_TILE - 1) / ZC_TILE;
    uint32_t loaded = 0, stage_phase = 0;

    /* ---- prologue: stage tiles 0 and 1, hash them, head-table pass of tile 0 ---- */
    {
        const uint32_t need = min(q_end16, 2 * ZC_TILE + ZL_LOOKAHEAD);
        if (need > loaded) {
            if (tid == 0) zl_stage_bulk<ZC_RING>(S.ring32, gbase, loaded, need, &S.stage_bar);
            zl_mbar_wait(&S.stage_bar, stage_phase & 1u);
            stage_phase++;
            loaded = need;
        }
    }
    if (!is_hasher) {
        zc_hash_tile(S.ring32, S.t_hash[0], 0, q_dict, q_end, wtid);
        if (ntiles > 1) zc_hash_tile(S.ring32, S.t_hash[1], ZC_TILE, q_dict, q_end, wtid);
    }
    __syncthreads();
    if (is_hasher && ntiles > 0) zc_hasher_tile(S, S.t_hash[0], S.t_cand[0], 0, lane);
    __syncthreads();

    for (uint32_t k = 0; k < ntiles; k++) {
        const uint32_t t0 = k * ZC_TILE;
        const uint32_t need = min(q_end16, t0 + 3 * ZC_TILE + ZL_LOOKAHEAD);
        const bool staging = need > loaded;
        if (is_hasher) {
            if (staging && lane == 0) zl_stage_bulk<ZC_RING>(S.ring32, gbase, loaded, need, &S.stage_bar);
            if (k + 1 < ntiles) zc_hasher_tile(S, S.t_hash[(k + 1) & 1], S.t_cand[(k + 1) & 1], t0 + ZC_TILE, lane);
        } else {
            const uint16_t *cand = S.t_cand[k & 1];
            const bool live = (t0 + ZC_TILE > q_start);            /* not a dictionary-only tile */
            /* ---- publish the links of this tile: walks from it may follow links of positions >= t0 + ZC_TILE - 32768 ---- */
            for (uint32_t i = wtid; i < ZC_TILE; i += ZL_WORKERS) S.prevd[(t0 + i) & (ZS_WINDOW - 1)] = cand[i];
            if (live) {
                /* ---- first candidates: 16 bytes per lane, long matches extended once per run of equal distances
                   (exactly the per-position match lengths; see deflate_lz.cu) ---- */
                const uint32_t le = zs_lanemask_lt() | (1u << lane), gt = zs_lanemask_gt();
                for (uint32_t i = wtid; i < ZC_TILE; i += ZL_WORKERS) {
                    const uint32_t q = t0 + i;
                    uint32_t d = 0, limit = 0;
                    if (q >= q_start && q + 3 <= q_end) {
                        d = cand[i];
                        if (d > min((uint32_t)P.max_dist, q - q_dict)) d = 0;
                        limit = q_end - q;
                    }
                    uint32_t best = 0;
                    if (d) best = zl_match16<ZC_RING>(S.ring32, q, d);
                    const bool lng = d != 0 && best == 16 && limit > 16;
                    const uint32_t lm = __ballot_sync(0xFFFFFFFFu, lng);
                    if (lm) {
                        const uint32_t dprev = __shfl_up_sync(0xFFFFFFFFu, d, 1);
                        const bool prev_l = lane > 0 && ((lm >> (lane - 1)) & 1u);
                        const bool head = lng && (!prev_l || d != dprev);
                        const uint32_t hm = __ballot_sync(0xFFFFFFFFu, head);
                        const uint32_t hl = 31u - (uint32_t)__clz((int)(hm & le));
                        const uint32_t stop = (hm | ~lm) & gt;
                        const uint32_t nexth = stop ? (uint32_t)__ffs((int)stop) - 1u : 32u;
                        uint32_t ext = 0;
                        if (head) ext = 16u + zl_match_ext<ZC_RING>(S.ring32, q + 16, d, min(limit, ZS_MAX_MATCH + (nexth - lane - 1u)) - 16u);
                        const uint32_t e = __shfl_sync(0xFFFFFFFFu, ext, hl & 31u);
                        if (lng) best = e - (lane - hl);
                    }
                    const uint32_t ml = min(limit, ZS_MAX_MATCH);
                    best = min(best, ml);
                    S.best[i] = (uint16_t)best;
                    S.bestd[i] = (uint16_t)(best ? d : 0u);
                    S.cur[i] = (uint16_t)((d != 0 && !(best >= (uint32_t)P.nice || best >= ml)) ? d : 0u);
                    S.budget[i] = (int16_t)zc_budget(P.chain, cd.len);
                }
                if (wtid < ZC_GROUPS) S.cutbits[wtid] = 0;
                zl_bar_workers();

                uint32_t jn[ZC_GPW], marks[ZC_GPW];
                const uint32_t b0 = ww * ZC_BLOCK;
                /* ---- rounds: parse, list the positions that deserve a deeper search, search ---- */
                for (int r = 0; r < ZC_ROUNDS; r++) {
                    zc_parse(S, P, t0, wtid, jn, marks);
                    /* publish the marks: a position's rule looks at the position before it */
#pragma unroll
                    for (int g = 0; g < ZC_GPW; g++) {
                        const uint32_t tk = __ballot_sync(0xFFFFFFFFu, ((marks[g] >> lane) & 1u) && jn[g] >= 3u);
                        if (lane == 0) { S.m_any[ww * ZC_GPW + g] = marks[g]; S.m_take[ww * ZC_GPW + g] = tk; }
                    }
                    zl_bar_workers();
#pragma unroll
                    for (int g = 0; g < ZC_GPW; g++) {
                        const uint32_t gi = ww * ZC_GPW + g, i = b0 + g * 32 + lane;
                        const uint32_t many = marks[g], mtake = S.m_take[gi];
                        /* marks of the position before: bit lane-1 of this group, or bit 31 of the previous group */
                        uint32_t pany, ptake;
                        if (lane > 0) { pany = (many >> (lane - 1)) & 1u; ptake = (mtake >> (lane - 1)) & 1u; }
                        else if (gi > 0) { pany = S.m_any[gi - 1] >> 31; ptake = S.m_take[gi - 1] >> 31; }
                        else { pany = 0; ptake = 0; }
                        bool need = (many >> lane) & 1u;
                        bool cutnow = false;
                        const bool alive = S.cur[i] != 0;
                        if (alive && pany) {
                            const uint32_t Lp = zc_flen(S, i - 1, (uint32_t)P.min_len);
                            if (ptake && P.lazy && Lp < (uint32_t)P.max_lazy) need = true;
                            if (need && Lp >= (uint32_t)P.good && !((S.cutbits[gi] >> lane) & 1u)) cutnow = true;
                        }
                        need = need && alive;
                        if (cutnow) S.budget[i] = (int16_t)(S.budget[i] >> 2);
                        const uint32_t cm = __ballot_sync(0xFFFFFFFFu, cutnow);
                        const uint32_t nm = __ballot_sync(0xFFFFFFFFu, need);
                        if (lane == 0 && cm) S.cutbits[gi] |= cm;
                        if (nm) {
                            uint32_t base = 0;
                            if (lane == 0) base = atomicAdd(&S.list_n, (uint32_t)__popc(nm));
                            base = __shfl_sync(0xFFFFFFFFu, base, 0);
                            if (need) S.list[base + __popc(nm & zs_lanemask_lt())] = (uint16_t)i;
                        }
                    }
                    zl_bar_workers();
                    const uint32_t n = S.list_n;
                    if (n == 0) break;
                    const int cap = zc_round_cap(r);
                    for (uint32_t j = wtid; j < n; j += ZL_WORKERS) zc_walk_one(S, P, S.list[j], t0, cap, q_dict, q_end);
                    zl_bar_workers();
                }

                /* ---- final parse, symbols ---- */
                zc_parse(S, P, t0, wtid, jn, marks);
                const uint32_t leave = S.g_cnt[0];
                zl_bar_workers();
                if (wtid == 0) S.carry = t0 + leave;
                uint32_t vmask[ZC_GPW], val[ZC_GPW];
#pragma unroll
                for (int g = 0; g < ZC_GPW; g++) {
                    const uint32_t i = b0 + g * 32 + lane, q = t0 + i;
                    const bool v = ((marks[g] >> lane) & 1u) && q >= q_start && q < q_end;
                    vmask[g] = __ballot_sync(0xFFFFFFFFu, v);
                    val[g] = (jn[g] >= 3) ? zs_match(jn[g], S.bestd[i]) : zl_ld8<ZC_RING>(S.ring32, q);
                    if (lane == 0) S.g_cnt[ww * ZC_GPW + g] = __popc(vmask[g]);
                }
                zl_bar_workers();
                /* scan of group counts (worker warp 0) */
                if (ww == 0) {
                    const uint32_t c0 = S.g_cnt[lane * 2], c1 = S.g_cnt[lane * 2 + 1];
                    const uint32_t s = c0 + c1;
                    uint32_t inc = s;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += t; }
                    const uint32_t base = S.nsym + inc - s;
                    S.g_off[lane * 2] = base; S.g_off[lane * 2 + 1] = base + c0;
                    __syncwarp();
                    if (lane == 31) S.nsym = S.nsym + inc;
                }
                zl_bar_workers();
#pragma unroll
                for (int g = 0; g < ZC_GPW; g++) {
                    const uint32_t q = t0 + b0 + g * 32 + lane;
                    if ((vmask[g] >> lane) & 1u) {
                        const uint32_t idx = S.g_off[ww * ZC_GPW + g] + __popc(vmask[g] & zs_lanemask_lt());
                        out_sym[idx] = val[g];
                        if ((idx & (ZS_BLOCK_SYMS - 1)) == 0) blk_in_start[cd.blk_base + idx / ZS_BLOCK_SYMS] = q - q_start;
                    }
                }
            }
            /* ---- hashes of tile k+2 (its bytes were queued at the top; wait until they have landed) ---- */
            if (staging) { if (lane == 0) zl_mbar_wait(&S.stage_bar, stage_phase & 1u); __syncwarp(); }   /* one poll per warp */
            if (k + 2 < ntiles) zc_hash_tile(S.ring32, S.t_hash[k & 1], t0 + 2 * ZC_TILE, q_dict, q_end, wtid);
        }
        if (staging) { stage_phase++; loaded = need; }
        __syncthreads();
    }
    if (tid == 0) chunk_nsym[blockIdx.x] = S.nsym;
}

static_assert(ZC_GROUPS == 64, "group-count scan assumes 64 groups per tile");
static_assert(ZC_GPW * ZL_WORKER_WARPS == ZC_GROUPS, "groups must divide evenly over the worker warps");
static_assert(ZS_WINDOW + 3u * ZC_TILE + ZL_LOOKAHEAD <= ZC_RING, "chain ring too small");
static_assert(sizeof(ZcSmem) <= 227u * 1024u, "chain kernel shared memory");

extern "C" size_t zs_lzc_smem_bytes(void) { return sizeof(ZcSmem); }

extern "C" cudaError_t zs_lzc_launch(cudaStream_t st, uint32_t nchunks, const uint8_t *raw,
                                     const ZsChunk *chunks, uint32_t *sym, uint32_t *chunk_nsym,
                                     uint32_t *blk_in_start, ZsLzParams P)
{
    if (nchunks == 0) return cudaSuccess;
    cudaFuncSetAttribute(zs_lzc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZcSmem));
    zs_lzc_kernel<<<nchunks, ZL_THREADS, sizeof(ZcSmem), st>>>(raw, chunks, sym, chunk_nsym, blk_in_start, P);
    return cudaGetLastError();
}
