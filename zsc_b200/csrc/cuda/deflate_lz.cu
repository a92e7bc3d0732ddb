/* deflate_lz.cu — LZ77 match finding + parse, one CTA per chunk (sm_100a).
 *
 * Takes the place of the reference's fill_window / INSERT_STRING / longest_match /
 * deflate_fast / deflate_slow / deflate_rle / deflate_huff (reference src/deflate.c:1400-2245) for a
 * whole batch of independent chunks.  The algorithm is re-designed for a GPU rather than ported:
 *
 *   - the chunk streams through a 64 KiB shared-memory ring (32 KiB history + tile + lookahead);
 *   - a tile of ZL_TILE positions is handled in barrier-separated phases:
 *       B  every position gets its 3-byte hash (all threads)
 *       C  one warp walks the tile 32 positions at a time: head-table lookup, __match_any_sync to
 *          find same-hash peers inside the group, head/chain update  -> first candidate per position
 *       D  every position compares against its candidate(s) (all threads; chain walk for the
 *          deeper levels) -> best (length, distance) per position
 *       E  the parse (greedy or one-step lazy) is a pure function next(p) of the per-position
 *          results, so it is resolved with pointer doubling inside each 32-position group
 *          (warp shuffles), one short serial hop per group, and ballot/popc compaction;
 *   - symbols go to the sym arena as 32-bit words (see ZS_MATCH in huff_build.h); the block
 *     histogram, code construction and bit packing are separate kernels (deflate_huff.cu).
 *
 * Candidate positions are always verified byte-for-byte, so stale or aliased head-table entries
 * can cost ratio but never correctness.
 */
#include "common.cuh"

#define ZL_THREADS 512
#define ZL_WARPS (ZL_THREADS / 32)
#define ZL_TILE 2048
#define ZL_GROUPS (ZL_TILE / 32)
#define ZL_GPW (ZL_GROUPS / ZL_WARPS)      /* groups per warp */
#define ZL_RING 65536u
#define ZL_RING_MASK 0xFFFFu
#define ZL_LOOKAHEAD 272u                  /* >= 258 + 3, multiple of 16 */
#define ZL_HASH_BITS 15
#define ZL_NOHASH 0xFFFFu

struct ZlSmem {
    uint32_t ring32[ZL_RING / 4];
    uint16_t head[1 << ZL_HASH_BITS];
    uint16_t t_hash[ZL_TILE];
    uint16_t t_dist[ZL_TILE];
    uint16_t t_len[ZL_TILE + 32];
    uint16_t t_exit[ZL_TILE];
    uint16_t g_entry[ZL_GROUPS];
    uint32_t g_cnt[ZL_GROUPS];
    uint32_t g_off[ZL_GROUPS];
    uint32_t carry;       /* absolute q of the next parse start */
    uint32_t nsym;        /* symbols emitted so far */
};
struct ZlSmemChain {
    ZlSmem s;
    uint16_t prevd[ZS_WINDOW];   /* distance from a position to the previous one with the same hash */
};

__device__ __forceinline__ uint32_t zl_ld32(const uint32_t *ring32, uint32_t q)
{
    uint32_t i = (q >> 2) & (ZL_RING / 4 - 1);
    uint32_t w0 = ring32[i], w1 = ring32[(i + 1) & (ZL_RING / 4 - 1)];
    return __funnelshift_r(w0, w1, (q & 3) * 8);
}
__device__ __forceinline__ uint32_t zl_ld8(const uint32_t *ring32, uint32_t q)
{
    return ((const uint8_t *)ring32)[q & ZL_RING_MASK];
}
__device__ __forceinline__ uint32_t zl_hash(uint32_t v)
{
    return ((v & 0xFFFFFFu) * 2654435761u) >> (32 - ZL_HASH_BITS);
}

/* length of the common prefix of the strings at q and q - d, at most maxl */
__device__ __forceinline__ uint32_t zl_match_len(const uint32_t *ring32, uint32_t q, uint32_t d, uint32_t maxl)
{
    uint32_t l = 0;
    while (l < maxl) {
        uint32_t x = zl_ld32(ring32, q + l) ^ zl_ld32(ring32, q + l - d);
        if (x) { l += (uint32_t)(__ffs((int)x) - 1) >> 3; break; }
        l += 4;
    }
    return l < maxl ? l : maxl;
}

template <bool CHAIN>
__global__ void __launch_bounds__(ZL_THREADS, 1)
zs_lz_kernel(const uint8_t *__restrict__ raw, const ZsChunk *__restrict__ chunks,
             uint32_t *__restrict__ sym, uint32_t *__restrict__ chunk_nsym,
             uint32_t *__restrict__ blk_in_start, ZsLzParams P)
{
    extern __shared__ __align__(16) unsigned char zl_smem_raw[];
    ZlSmem &S = *reinterpret_cast<ZlSmem *>(zl_smem_raw);
    uint16_t *prevd = CHAIN ? reinterpret_cast<ZlSmemChain *>(zl_smem_raw)->prevd : nullptr;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ZsChunk cd = chunks[blockIdx.x];
    const uint64_t src_addr = (uint64_t)(raw + cd.raw_off) - cd.dict_len;
    const uint32_t a = (uint32_t)(src_addr & 15);
    const uint8_t *gbase = (const uint8_t *)(src_addr - a);       /* q = 0 */
    const uint32_t q_dict = a, q_start = a + cd.dict_len, q_end = q_start + cd.len;
    uint32_t *out_sym = sym + cd.sym_off;

    for (uint32_t i = tid; i < (1u << ZL_HASH_BITS) / 2; i += ZL_THREADS) ((uint32_t *)S.head)[i] = 0;
    if (CHAIN) for (uint32_t i = tid; i < ZS_WINDOW / 2; i += ZL_THREADS) ((uint32_t *)prevd)[i] = 0;
    for (uint32_t i = tid; i < ZL_TILE + 32; i += ZL_THREADS) S.t_len[i] = 0;
    if (tid == 0) { S.carry = q_start; S.nsym = 0; if (cd.len == 0) blk_in_start[cd.blk_base] = 0; }
    __syncthreads();

    uint32_t loaded = 0;   /* ring holds q in [0, loaded) (uniform across the CTA) */
    const uint32_t t_first = (P.mode == 0) ? 0 : (q_start / ZL_TILE) * ZL_TILE;

    for (uint32_t t0 = t_first; t0 < q_end; t0 += ZL_TILE) {
        /* ---- A: stream input into the ring (16-byte vectors; bytes for the ragged end) ---- */
        {
            uint32_t need = min(q_end, t0 + ZL_TILE + ZL_LOOKAHEAD);
            if (loaded < t0 && P.mode != 0) loaded = (t0 > ZS_WINDOW ? t0 - ZS_WINDOW : 0) & ~15u;
            uint32_t full_end = need & ~15u;             /* vectors entirely inside the data */
            for (uint32_t q = loaded + tid * 16; q < full_end; q += ZL_THREADS * 16) {
                uint4 v = __ldg(reinterpret_cast<const uint4 *>(gbase + q));
                *reinterpret_cast<uint4 *>(&((uint8_t *)S.ring32)[q & ZL_RING_MASK]) = v;
            }
            uint32_t tail0 = max(loaded, full_end);
            for (uint32_t q = tail0 + tid; q < need; q += ZL_THREADS)
                ((uint8_t *)S.ring32)[q & ZL_RING_MASK] = __ldg(gbase + q);
            /* zero a few bytes past the very end so 4-byte compares read defined data */
            if (need == q_end && tid < 8) ((uint8_t *)S.ring32)[(q_end + tid) & ZL_RING_MASK] = 0;
            loaded = (need == q_end) ? need : full_end;
        }
        __syncthreads();

        if (P.mode == 0) {
            /* ---- B: hashes ---- */
            for (uint32_t i = tid; i < ZL_TILE; i += ZL_THREADS) {
                uint32_t q = t0 + i;
                uint32_t h = ZL_NOHASH;
                if (q >= q_dict && q + 3 <= q_end) h = zl_hash(zl_ld32(S.ring32, q));
                S.t_hash[i] = (uint16_t)h;
            }
            __syncthreads();
            /* ---- C: head-table pass, one warp, groups in order ---- */
            if (warp == 0) {
                const uint32_t lt = zs_lanemask_lt(), gt = zs_lanemask_gt();
#pragma unroll 4
                for (uint32_t g = 0; g < ZL_GROUPS; g++) {
                    uint32_t i = g * 32 + lane, q = t0 + i;
                    uint32_t h = S.t_hash[i];
                    bool valid = (h != ZL_NOHASH);
                    uint32_t peers = __match_any_sync(0xFFFFFFFFu, h);
                    uint32_t d = 0;
                    if (valid) {
                        uint32_t old = S.head[h];
                        uint32_t below = peers & lt;
                        d = below ? (lane - (31u - (uint32_t)__clz((int)below))) : ((q - old) & 0xFFFFu);
                        if (!(peers & gt)) S.head[h] = (uint16_t)q;
                        if (CHAIN) prevd[q & (ZS_WINDOW - 1)] = (uint16_t)d;
                    }
                    S.t_dist[i] = (uint16_t)d;
                    __syncwarp();
                }
            }
            __syncthreads();
        }
        if (t0 + ZL_TILE <= q_start) continue;      /* dictionary-only tile */

        /* ---- D: match lengths ---- */
        if (P.mode != 2) {
            for (uint32_t i = tid; i < ZL_TILE; i += ZL_THREADS) {
                uint32_t q = t0 + i;
                uint32_t best = 0, bestd = 0;
                if (q >= q_start && q + 3 <= q_end) {
                    uint32_t maxl = min(ZS_MAX_MATCH, q_end - q);
                    uint32_t maxd = min((uint32_t)P.max_dist, q - q_dict);
                    if (P.mode == 1) {
                        if (maxd >= 1) { best = zl_match_len(S.ring32, q, 1, maxl); bestd = 1; }
                    } else {
                        uint32_t d = S.t_dist[i];
                        int budget = P.chain;
                        while (d != 0 && d <= maxd) {
                            /* cheap reject: the byte that would extend the best match must agree */
                            if (best < 3 || zl_ld8(S.ring32, q + best) == zl_ld8(S.ring32, q + best - d)) {
                                uint32_t l = zl_match_len(S.ring32, q, d, maxl);
                                if (l > best) { best = l; bestd = d; if (l >= (uint32_t)P.nice || l >= maxl) break; }
                            }
                            if (!CHAIN || budget-- <= 0) break;
                            uint32_t c = q - d;
                            if (c + ZS_WINDOW < t0 + ZL_TILE) break;     /* its link was recycled */
                            uint32_t step = prevd[c & (ZS_WINDOW - 1)];
                            if (step == 0) break;
                            d += step;
                        }
                    }
                    if (best < (uint32_t)P.min_len || (best == 3 && bestd > ZS_TOO_FAR)) { best = 0; bestd = 0; }
                }
                S.t_len[i] = (uint16_t)best;
                S.t_dist[i] = (uint16_t)bestd;
            }
        } else {
            for (uint32_t i = tid; i < ZL_TILE; i += ZL_THREADS) { S.t_len[i] = 0; S.t_dist[i] = 0; }
        }
        __syncthreads();

        /* ---- E1: per-group exit function by pointer doubling ---- */
        uint32_t jn[ZL_GPW];
#pragma unroll
        for (int k = 0; k < ZL_GPW; k++) {
            uint32_t g = warp + k * ZL_WARPS, i = g * 32 + lane;
            uint32_t L = S.t_len[i];
            bool take = L >= 3;
            if (take && P.lazy && S.t_len[i + 1] > L) take = false;
            uint32_t n = take ? L : 1u;
            jn[k] = n;
            uint32_t j = lane + n;
#pragma unroll
            for (int r = 0; r < 5; r++) {
                uint32_t jj = __shfl_sync(0xFFFFFFFFu, j, j & 31);
                if (j < 32) j = jj;
            }
            S.t_exit[i] = (uint16_t)(g * 32 + j);
        }
        __syncthreads();
        /* ---- E2: hop from group to group ---- */
        if (tid == 0) {
            uint32_t s = S.carry - t0;
            for (uint32_t g = 0; g < ZL_GROUPS; g++) {
                uint32_t e = 0xFFFFu;
                if (s < (g + 1) * 32) { e = s; s = S.t_exit[s]; }
                S.g_entry[g] = (uint16_t)e;
            }
            S.carry = t0 + s;
        }
        __syncthreads();
        /* ---- E3: mark parse starts, count ---- */
        uint32_t vmask[ZL_GPW], val[ZL_GPW];
#pragma unroll
        for (int k = 0; k < ZL_GPW; k++) {
            uint32_t g = warp + k * ZL_WARPS, i = g * 32 + lane, q = t0 + i;
            uint32_t e = S.g_entry[g];
            uint32_t marks = (e != 0xFFFFu) ? (1u << (e - g * 32)) : 0u;
            uint32_t n = jn[k];
            uint32_t j = lane + n;
#pragma unroll
            for (int r = 0; r < 5; r++) {
                uint32_t contrib = (((marks >> lane) & 1u) && j < 32) ? (1u << j) : 0u;
                marks |= __reduce_or_sync(0xFFFFFFFFu, contrib);
                uint32_t jj = __shfl_sync(0xFFFFFFFFu, j, j & 31);
                if (j < 32) j = jj;
            }
            bool v = ((marks >> lane) & 1u) && q >= q_start && q < q_end;
            vmask[k] = __ballot_sync(0xFFFFFFFFu, v);
            val[k] = (n >= 3) ? zs_match(n, S.t_dist[i]) : zl_ld8(S.ring32, q);
            if (lane == 0) S.g_cnt[g] = __popc(vmask[k]);
        }
        __syncthreads();
        /* ---- scan of group counts (warp 0) ---- */
        if (warp == 0) {
            uint32_t c0 = S.g_cnt[lane * 2], c1 = S.g_cnt[lane * 2 + 1];
            uint32_t s = c0 + c1, inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += t; }
            uint32_t base = S.nsym + inc - s;
            S.g_off[lane * 2] = base; S.g_off[lane * 2 + 1] = base + c0;
            __syncwarp();
            if (lane == 31) S.nsym = S.nsym + inc;
        }
        __syncthreads();
        /* ---- write symbols ---- */
#pragma unroll
        for (int k = 0; k < ZL_GPW; k++) {
            uint32_t g = warp + k * ZL_WARPS, q = t0 + g * 32 + lane;
            if ((vmask[k] >> lane) & 1u) {
                uint32_t idx = S.g_off[g] + __popc(vmask[k] & zs_lanemask_lt());
                out_sym[idx] = val[k];
                if ((idx & (ZS_BLOCK_SYMS - 1)) == 0) blk_in_start[cd.blk_base + idx / ZS_BLOCK_SYMS] = q - q_start;
            }
        }
        /* the next tile's phase A only touches ring bytes older than the window; phases B..E are
           separated from this tile's reads by the barrier after A */
    }
    __syncthreads();
    if (tid == 0) chunk_nsym[blockIdx.x] = S.nsym;
}

static_assert(ZL_GROUPS == 64, "group-count scan assumes 64 groups per tile");
static_assert(ZL_GPW * ZL_WARPS == ZL_GROUPS, "groups must divide evenly over the warps");

extern "C" size_t zs_lz_smem_bytes(int chain) { return chain ? sizeof(ZlSmemChain) : sizeof(ZlSmem); }

extern "C" cudaError_t zs_lz_launch(cudaStream_t st, int chain, uint32_t nchunks, const uint8_t *raw,
                                    const ZsChunk *chunks, uint32_t *sym, uint32_t *chunk_nsym,
                                    uint32_t *blk_in_start, ZsLzParams P)
{
    if (nchunks == 0) return cudaSuccess;
    if (chain) {
        cudaFuncSetAttribute(zs_lz_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZlSmemChain));
        zs_lz_kernel<true><<<nchunks, ZL_THREADS, sizeof(ZlSmemChain), st>>>(raw, chunks, sym, chunk_nsym, blk_in_start, P);
    } else {
        cudaFuncSetAttribute(zs_lz_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZlSmem));
        zs_lz_kernel<false><<<nchunks, ZL_THREADS, sizeof(ZlSmem), st>>>(raw, chunks, sym, chunk_nsym, blk_in_start, P);
    }
    return cudaGetLastError();
}
