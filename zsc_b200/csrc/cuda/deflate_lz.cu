/* deflate_lz.cu — LZ77 match finding + parse, one CTA per chunk (sm_100a).
 *
 * Takes the place of the reference's fill_window / INSERT_STRING / longest_match /
 * deflate_fast / deflate_slow / deflate_rle / deflate_huff (reference src/deflate.c:1400-2245) for a
 * whole batch of independent chunks.  The algorithm is re-designed for a GPU rather than ported:
 *
 *   - the chunk streams through a 64 KiB shared-memory ring (32 KiB history + tiles in flight);
 *   - work advances in tiles of ZL_TILE positions through a two-stage, warp-specialised pipeline:
 *       hasher warp (warp 0), one tile ahead: walks the tile 32 positions at a time; every lane reads
 *          the head table for its 3-byte hash (-> candidate distance), the highest lane of each hash
 *          then claims the slot (shuffle + store + read-back, no __match_any_sync: that instruction
 *          costs ~15 cycles per distinct value), chain links are recorded for the deeper levels;
 *       16 worker warps: stage the next tiles into the ring and hash them (A/B), compare every
 *          position with its candidate(s) (D), and resolve the parse (E): greedy / one-step-lazy
 *          parsing is a pure function next(p) of the per-position results, so it is resolved with
 *          pointer doubling inside each 32-position group (warp shuffles), composed per 128-position
 *          block, one short serial hop per block, then ballot/popc compaction;
 *   - symbols go to the sym arena as 32-bit words (see ZS_MATCH in huff_build.h); the block
 *     histogram, code construction and bit packing are separate kernels (deflate_huff.cu).
 *
 * Candidate positions are always verified byte-for-byte, so stale or aliased head-table entries
 * can cost ratio but never correctness.  Every step is deterministic: tests/cpu_harness.cpp holds
 * a scalar model that predicts the symbol stream bit for bit.
 */
#include "lz_common.cuh"

#define ZL_TILE 2048
#define ZL_GROUPS (ZL_TILE / 32)
#define ZL_GPW (ZL_GROUPS / ZL_WORKER_WARPS)       /* consecutive groups per worker warp (one block) */
#define ZL_BLOCK (ZL_GPW * 32)                     /* 128 positions */
/* The single-candidate kernel (level 1, Z_RLE, Z_HUFFMAN_ONLY; levels 2..9 are deflate_chain.cu) uses a 32 KiB ring
 * and a 14-bit head table so that two CTAs share an SM (2 x 97 KB of shared memory, 56 registers): the second CTA's
 * warps fill the issue slots the first leaves empty at its barriers and shared-memory round trips.  A 32 KiB
 * ring holds the three staged tiles plus the window, which bounds match distances to ZL_FAST_MAX_DIST. */
#define ZL_RING 32768u
#define ZL_HASH_BITS 14u
#define ZL_FAST_MAX_DIST (32768u - 3u * ZL_TILE - ZL_LOOKAHEAD)   /* 26352: see zs_lz_fast_max_dist() */
#define ZL_NOHASH 0xFFFFu
#define ZL_NOTFIRST 0x8000u                      /* t_hash flag: a lower lane of the group has the same hash */

struct ZlSmem {
    uint32_t ring32[ZL_RING / 4 + ZL_MIRROR / 4];   /* + a mirror of the first bytes: multi-word reads never wrap */
    uint16_t head[1u << ZL_HASH_BITS];
    uint16_t t_hash[2][ZL_TILE];      /* double buffered: written by workers, read by the hasher */
    uint16_t t_cand[2][ZL_TILE];      /* double buffered: written by the hasher, read by workers */
    uint16_t t_dist[ZL_TILE];         /* best distance per position */
    uint16_t t_len[ZL_TILE + 32];     /* best length per position (+ zero sentinel) */
    uint16_t t_exit[ZL_TILE];         /* first parse start beyond the position's group */
    uint16_t t_bexit[ZL_TILE];        /* first parse start beyond the position's 128-block */
    uint16_t b_entry[ZL_WORKER_WARPS];
    uint32_t g_cnt[ZL_GROUPS];
    uint32_t g_off[ZL_GROUPS];
    uint32_t carry;                   /* absolute q of the next parse start */
    uint32_t nsym;                    /* symbols emitted so far */
    unsigned long long stage_bar;     /* mbarrier of the bulk-copy staging (lz_common.cuh) */
};

/* workers: 3-byte hashes of one tile (positions i, i + 512, ...: a warp covers one group of 32), 15 bits +
 * ZL_NOTFIRST when a lower lane of the group has the same hash.  Finding those duplicates here, on 16 warps
 * and off the hasher's serial walk, lets the hasher store without any conflict handling.  The search is a
 * slot-claiming loop in a 256-entry per-warp scratch: lanes store (rest of hash, lane) at slot hash & 255 and
 * read back; equal hashes meet in one slot and settle on their lowest lane, lanes that lost the slot to a
 * different hash go round again.  Exact and independent of which colliding store the hardware keeps. */
template <uint32_t RING, uint32_t HASH_BITS> __device__ __forceinline__ void zl_hash_tile(const uint32_t *ring32, uint16_t *t_hash, uint16_t *scratch, uint32_t t0, uint32_t q_dict, uint32_t q_end, uint32_t wtid)
{
    const uint32_t lane = wtid & 31;
    uint16_t *sc = scratch + (wtid >> 5) * 256;
    for (uint32_t i = wtid; i < ZL_TILE; i += ZL_WORKERS) {
        uint32_t q = t0 + i, h = ZL_NOHASH;
        if (q >= q_dict && q + 3 <= q_end) h = zl_hash<HASH_BITS>(zl_ld32<RING>(ring32, q));
        const bool valid = (h != ZL_NOHASH);
        const uint32_t hp = __shfl_up_sync(0xFFFFFFFFu, h, 1);
        bool notfirst = valid && lane > 0 && hp == h;          /* runs: decided by the neighbour alone */
        bool un = valid && !notfirst;
        const uint32_t slot = h & 255u, tag = ((h >> 8) << 5) | lane;
        while (__any_sync(0xFFFFFFFFu, un)) {
            if (un) sc[slot] = (uint16_t)tag;
            __syncwarp();
            uint32_t r;
            for (;;) {
                r = un ? (uint32_t)sc[slot] : tag;
                const bool higher_won = un && (r >> 5) == (tag >> 5) && (r & 31u) > lane;
                if (!__any_sync(0xFFFFFFFFu, higher_won)) break;
                if (higher_won) sc[slot] = (uint16_t)tag;
                __syncwarp();
            }
            if (un && (r >> 5) == (tag >> 5)) { notfirst = (r & 31u) != lane; un = false; }
            __syncwarp();
        }
        t_hash[i] = (uint16_t)(h | ((valid && notfirst) ? ZL_NOTFIRST : 0u));
    }
}

/* hasher warp: head-table pass over one tile, groups of 32 positions in order.
 * Candidate of a position = 1 if the previous lane has the same hash (runs), else the head-table entry as it
 * stood before the group; afterwards the lowest lane of every hash holds the slot.  The workers have flagged
 * every other lane ZL_NOTFIRST, so the stores of a group never collide and the walk is a plain in-order
 * stream of shared-memory loads and stores with nothing to wait for. */
__device__ __forceinline__ void zl_hasher_tile(ZlSmem &S, const uint16_t *t_hash, uint16_t *t_cand, uint32_t t0, uint32_t lane)
{
#pragma unroll 4
    for (uint32_t g = 0; g < ZL_GROUPS; g++) {
        const uint32_t i = g * 32 + lane, q = t0 + i;
        const uint32_t h16 = t_hash[i];
        const bool valid = (h16 != ZL_NOHASH);
        const uint32_t hs = valid ? (h16 & 0x7FFFu) : 0;
        const uint32_t old = S.head[hs];                                    /* table as it stood before the group */
        if (valid && !(h16 & ZL_NOTFIRST)) S.head[hs] = (uint16_t)q;        /* the first lane of each hash claims it */
        t_cand[i] = (uint16_t)old;                                          /* the workers turn it into a distance */
        __syncwarp();
    }
}

__global__ void __launch_bounds__(ZL_THREADS, 2)
zs_lz_kernel(const uint8_t *__restrict__ raw, const ZsChunk *__restrict__ chunks,
             uint32_t *__restrict__ sym, uint32_t *__restrict__ chunk_nsym,
             uint32_t *__restrict__ blk_in_start, ZsLzParams P)
{
    extern __shared__ __align__(16) unsigned char zl_smem_raw[];
    ZlSmem &S = *reinterpret_cast<ZlSmem *>(zl_smem_raw);
    constexpr uint32_t RING = ZL_RING, HASH_BITS = ZL_HASH_BITS;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_hasher = (warp == 0);
    const uint32_t wtid = tid - 32, ww = warp - 1;                 /* worker thread / warp index */
    const ZsChunk cd = chunks[blockIdx.x];
    const uint64_t src_addr = (uint64_t)(raw + cd.raw_off) - cd.dict_len;
    const uint32_t a = (uint32_t)(src_addr & 15);
    const uint8_t *gbase = (const uint8_t *)(src_addr - a);       /* q = 0 */
    const uint32_t q_dict = a, q_start = a + cd.dict_len, q_end = q_start + cd.len;
    const uint32_t q_end16 = (q_end + 15u) & ~15u;                 /* staging moves whole 16-byte pieces (the arenas are padded) */
    uint32_t *out_sym = sym + cd.sym_off;
    const bool hashing = (P.mode == 0);

    for (uint32_t i = tid; i < (1u << HASH_BITS) / 2; i += ZL_THREADS) ((uint32_t *)S.head)[i] = 0;
    for (uint32_t i = tid; i < ZL_TILE + 32; i += ZL_THREADS) S.t_len[i] = 0;
    if (tid == 0) { S.carry = q_start; S.nsym = 0; if (cd.len == 0) blk_in_start[cd.blk_base] = 0; zl_mbar_init(&S.stage_bar, 1); }
    __syncthreads();

    /* tiles are [t_first + k * ZL_TILE, ...); with hashing the dictionary is walked too */
    const uint32_t t_first = hashing ? 0 : (q_start / ZL_TILE) * ZL_TILE;
    const uint32_t ntiles = q_end > t_first ? (q_end - t_first + ZL_TILE - 1) / ZL_TILE : 0;
    /* without hashing (Z_RLE, Z_HUFFMAN_ONLY, level 0) nothing farther back than the byte before the first position is ever
       read: staging starts just below the first tile (a staging step must stay below the ring size, lz_common.cuh) */
    uint32_t loaded = hashing ? 0 : (t_first >= 16u ? t_first - 16u : 0u);
    uint32_t stage_phase = 0;                                      /* completed phases of the staging barrier */

    /* ---- prologue: stage tiles 0 and 1 (bulk copy issued by one thread), hash them, head-table pass of tile 0 ---- */
    {
        const uint32_t need = min(q_end16, t_first + 2 * ZL_TILE + ZL_LOOKAHEAD);
        if (need > loaded) {
            if (tid == 0) zl_stage_bulk<RING>(S.ring32, gbase, loaded, need, &S.stage_bar);
            zl_mbar_wait(&S.stage_bar, stage_phase & 1u);
            stage_phase++;
            loaded = need;
        }
    }
    if (!is_hasher && hashing) {
        zl_hash_tile<RING, HASH_BITS>(S.ring32, S.t_hash[0], S.t_exit, t_first, q_dict, q_end, wtid);
        if (ntiles > 1) zl_hash_tile<RING, HASH_BITS>(S.ring32, S.t_hash[1], S.t_exit, t_first + ZL_TILE, q_dict, q_end, wtid);
    }
    __syncthreads();
    if (is_hasher && hashing && ntiles > 0) zl_hasher_tile(S, S.t_hash[0], S.t_cand[0], t_first, lane);
    __syncthreads();

    for (uint32_t k = 0; k < ntiles; k++) {
        const uint32_t t0 = t_first + k * ZL_TILE;
        /* ---- A: stage tile k+2 (only ring bytes older than the window of tile k are replaced): one thread queues the
           bulk copies, the workers wait for the bytes right before they hash that tile (phase B) ---- */
        const uint32_t need = min(q_end16, t0 + 3 * ZL_TILE + ZL_LOOKAHEAD);
        const bool staging = need > loaded;
        if (is_hasher) {
            if (staging && lane == 0) zl_stage_bulk<RING>(S.ring32, gbase, loaded, need, &S.stage_bar);
            /* one tile ahead of the workers */
            if (hashing && k + 1 < ntiles)
                zl_hasher_tile(S, S.t_hash[(k + 1) & 1], S.t_cand[(k + 1) & 1], t0 + ZL_TILE, lane);
        } else {
            const bool live = (t0 + ZL_TILE > q_start);           /* not a dictionary-only tile */
            if (live) {
                /* ---- D: match lengths ---- */
                const uint16_t *cand = S.t_cand[k & 1];
                if (P.mode != 2) {
                    /* One candidate per position.  First every lane compares 16 bytes (no divergence); that
                       settles all short matches.  Lanes that matched all 16 are long matches, and neighbouring
                       long lanes with the same distance are inside the same match: only the first lane of
                       such a run extends the comparison, the others subtract their offset from its extent.
                       Results are exactly the per-position match lengths. */
                    const uint32_t le = zs_lanemask_lt() | (1u << lane), gt = zs_lanemask_gt();
                    for (uint32_t i = wtid; i < ZL_TILE; i += ZL_WORKERS) {
                        const uint32_t q = t0 + i;
                        uint32_t d = 0, limit = 0;
                        const uint32_t h16 = hashing ? (uint32_t)S.t_hash[k & 1][i] : ZL_NOHASH;
                        const uint32_t hp16 = __shfl_up_sync(0xFFFFFFFFu, h16, 1);
                        if (q >= q_start && q + 3 <= q_end) {
                            if (P.mode == 1) d = 1u;
                            else {
                                /* same hash as the previous lane: a run, candidate distance 1; else the table entry */
                                const bool same = lane > 0 && hp16 != ZL_NOHASH && ((hp16 ^ h16) & 0x7FFFu) == 0;
                                d = (h16 == ZL_NOHASH) ? 0u : (same ? 1u : ((q - (uint32_t)cand[i]) & 0xFFFFu));
                            }
                            if (d > min(min((uint32_t)P.max_dist, ZL_FAST_MAX_DIST), q - q_dict)) d = 0;
                            limit = q_end - q;
                        }
                        uint32_t best = 0;
                        if (d) best = zl_match16<RING>(S.ring32, q, d);
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
                            if (head) ext = 16u + zl_match_ext<RING>(S.ring32, q + 16, d, min(limit, ZS_MAX_MATCH + (nexth - lane - 1u)) - 16u);
                            const uint32_t e = __shfl_sync(0xFFFFFFFFu, ext, hl & 31u);
                            if (lng) best = e - (lane - hl);
                        }
                        best = min(best, min(limit, ZS_MAX_MATCH));
                        uint32_t bestd = d;
                        if (best < (uint32_t)P.min_len || (best == 3 && bestd > ZS_TOO_FAR)) { best = 0; bestd = 0; }
                        S.t_len[i] = (uint16_t)best;
                        S.t_dist[i] = (uint16_t)bestd;
                    }
                } else {
                    for (uint32_t i = wtid; i < ZL_TILE; i += ZL_WORKERS) { S.t_len[i] = 0; S.t_dist[i] = 0; }
                }
                zl_bar_workers();

                /* ---- E1: exit function per group (pointer doubling), composed per 128-block ---- */
                uint32_t jn[ZL_GPW];
                const uint32_t b0 = ww * ZL_BLOCK;                 /* this warp's block, tile relative */
#pragma unroll
                for (int g = 0; g < ZL_GPW; g++) {
                    uint32_t i = b0 + g * 32 + lane;
                    uint32_t L = S.t_len[i];
                    bool take = L >= 3;
                    if (take && P.lazy && S.t_len[i + 1] > L) take = false;
                    uint32_t n = take ? L : 1u;
                    jn[g] = n;
                    uint32_t j = lane + n;
#pragma unroll
                    for (int r = 0; r < 5; r++) {
                        uint32_t jj = __shfl_sync(0xFFFFFFFFu, j, j & 31);
                        if (j < 32) j = jj;
                    }
                    S.t_exit[i] = (uint16_t)(b0 + g * 32 + j);
                }
                __syncwarp();
#pragma unroll
                for (int g = 0; g < ZL_GPW; g++) {
                    uint32_t i = b0 + g * 32 + lane;
                    uint32_t e = S.t_exit[i];
#pragma unroll
                    for (int hop = g + 1; hop < ZL_GPW; hop++) if (e < b0 + ZL_BLOCK) e = S.t_exit[e];
                    S.t_bexit[i] = (uint16_t)e;
                }
                zl_bar_workers();
                /* ---- E2: hop from block to block ---- */
                if (wtid == 0) {
                    uint32_t s = S.carry - t0;
#pragma unroll 4
                    for (uint32_t b = 0; b < ZL_WORKER_WARPS; b++) {
                        uint32_t e = ZL_NONE;
                        if (s < (b + 1) * ZL_BLOCK) { e = s; s = S.t_bexit[s]; }
                        S.b_entry[b] = (uint16_t)e;
                    }
                    S.carry = t0 + s;
                }
                zl_bar_workers();
                /* ---- E3: mark parse starts, count ---- */
                uint32_t vmask[ZL_GPW], val[ZL_GPW];
                uint32_t s_in = S.b_entry[ww];
#pragma unroll
                for (int g = 0; g < ZL_GPW; g++) {
                    uint32_t i = b0 + g * 32 + lane, q = t0 + i;
                    uint32_t marks = 0;
                    if (s_in != ZL_NONE && s_in < b0 + (g + 1) * 32) { marks = 1u << (s_in - (b0 + g * 32)); s_in = S.t_exit[s_in]; }
                    uint32_t n = jn[g];
                    uint32_t j = lane + n;
#pragma unroll
                    for (int r = 0; r < 5; r++) {
                        uint32_t contrib = (((marks >> lane) & 1u) && j < 32) ? (1u << j) : 0u;
                        marks |= __reduce_or_sync(0xFFFFFFFFu, contrib);
                        uint32_t jj = __shfl_sync(0xFFFFFFFFu, j, j & 31);
                        if (j < 32) j = jj;
                    }
                    bool v = ((marks >> lane) & 1u) && q >= q_start && q < q_end;
                    vmask[g] = __ballot_sync(0xFFFFFFFFu, v);
                    val[g] = (n >= 3) ? zs_match(n, S.t_dist[i]) : zl_ld8<RING>(S.ring32, q);
                    if (lane == 0) S.g_cnt[ww * ZL_GPW + g] = __popc(vmask[g]);
                }
                zl_bar_workers();
                /* ---- scan of group counts (worker warp 0) ---- */
                if (ww == 0) {
                    uint32_t c0 = S.g_cnt[lane * 2], c1 = S.g_cnt[lane * 2 + 1];
                    uint32_t s = c0 + c1, inc = s;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += t; }
                    uint32_t base = S.nsym + inc - s;
                    S.g_off[lane * 2] = base; S.g_off[lane * 2 + 1] = base + c0;
                    __syncwarp();
                    if (lane == 31) S.nsym = S.nsym + inc;
                }
                zl_bar_workers();
                /* ---- write symbols ---- */
#pragma unroll
                for (int g = 0; g < ZL_GPW; g++) {
                    uint32_t q = t0 + b0 + g * 32 + lane;
                    if ((vmask[g] >> lane) & 1u) {
                        uint32_t idx = S.g_off[ww * ZL_GPW + g] + __popc(vmask[g] & zs_lanemask_lt());
                        out_sym[idx] = val[g];
                        if ((idx & (ZS_BLOCK_SYMS - 1)) == 0) blk_in_start[cd.blk_base + idx / ZS_BLOCK_SYMS] = q - q_start;
                    }
                }
            }
            /* ---- B: hashes of tile k+2 (its bytes were queued in A; wait until they have landed) ---- */
            if (staging) { if (lane == 0) zl_mbar_wait(&S.stage_bar, stage_phase & 1u); __syncwarp(); }   /* one poll per warp */
            if (hashing && k + 2 < ntiles) zl_hash_tile<RING, HASH_BITS>(S.ring32, S.t_hash[k & 1], S.t_exit, t0 + 2 * ZL_TILE, q_dict, q_end, wtid);
        }
        if (staging) { stage_phase++; loaded = need; }
        __syncthreads();
    }
    if (tid == 0) chunk_nsym[blockIdx.x] = S.nsym;
}

static_assert(ZL_GROUPS == 64, "group-count scan assumes 64 groups per tile");
static_assert(ZL_GPW * ZL_WORKER_WARPS == ZL_GROUPS, "groups must divide evenly over the worker warps");

extern "C" size_t zs_lz_smem_bytes(void) { return sizeof(ZlSmem); }
/* longest match distance the single-candidate kernel can represent: its ring must hold the window of the tile
 * being searched and the three tiles staged ahead of it */
extern "C" uint32_t zs_lz_fast_max_dist(void) { return ZL_FAST_MAX_DIST; }
static_assert(ZL_FAST_MAX_DIST + 3u * ZL_TILE + ZL_LOOKAHEAD <= ZL_RING, "fast ring too small");

extern "C" cudaError_t zs_lz_launch(cudaStream_t st, uint32_t nchunks, const uint8_t *raw,
                                    const ZsChunk *chunks, uint32_t *sym, uint32_t *chunk_nsym,
                                    uint32_t *blk_in_start, ZsLzParams P)
{
    if (nchunks == 0) return cudaSuccess;
    cudaFuncSetAttribute(zs_lz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZlSmem));
    zs_lz_kernel<<<nchunks, ZL_THREADS, sizeof(ZlSmem), st>>>(raw, chunks, sym, chunk_nsym, blk_in_start, P);
    return cudaGetLastError();
}
