/* deflate_huff.cu — block histogram + code construction, output-offset scan, bit packing (sm_100a).
 *
 * The GPU form of the reference's Huffman stage (src/trees.c): _tr_tally (include/zsc/deflate.h:338-354)
 * becomes a per-block histogram over the symbol words the LZ kernel wrote; build_tree / gen_codes /
 * send_all_trees become zh_build_block (huff_build.h); compress_block + send_bits + flush_pending
 * (src/trees.c:948-993, :292-304, src/deflate.c:927) become a kernel in which every thread packs a run
 * of 32 symbols at a bit offset that comes from a prefix sum, into a shared-memory image of the
 * block that is then written to its final position in one coalesced pass.  Block, section and
 * stream offsets come from one scan (zk_elem algebra) — nothing is compacted after the fact.
 */
#include "common.cuh"

/* ======================= K2: histogram + codes ======================= */
#define ZB_THREADS 128
#define ZB_WARPS (ZB_THREADS / 32)

struct ZbSmem {
    uint32_t lfreq[ZB_WARPS][ZH_LCODES_PAD];
    uint32_t dfreq[ZB_WARPS][ZH_DCODES_PAD];
    zh_scratch sc;          /* literal/length tree, then code-length tree, tokens, lengths */
    zh_scratch sc2;         /* distance tree (built by one thread while the literal tree is in progress) */
    zh_block blk;
    uint32_t bl_count[16];
    uint32_t cnt_l[16], cnt_d[16], next_l[16], next_d[16];
    uint32_t red[2][ZB_WARPS];
    uint32_t scan[ZB_WARPS];
    uint32_t ccnt[9][16];
    uint16_t lct[256];      /* len - 3 -> 257 + length code */
    zh_decision D;
    int m, max_l, max_d, overflow;
};

/* What the three block kernels hand to each other, per block slot (global memory):
 * zs_block_kernel<0> leaves the sorted keys, the reduced histograms and the distance code lengths;
 * zs_merge_kernel adds the leaf depths of the literal/length tree; zs_block_kernel<1> reads it all back. */
struct __align__(16) ZbScratch {
    uint32_t key[ZH_LCODES_PAD];
    uint32_t lfreq[ZH_LCODES_PAD];
    uint32_t dfreq[ZH_DCODES_PAD];
    uint8_t depth[ZH_LCODES_PAD];           /* depth of the i-th sorted leaf in the literal/length tree (not yet limited) */
    uint8_t dlen[ZH_DCODES_PAD];
    int32_t m, max_l, max_d, pad;
};

/* zh_lengths (huff_build.h) for an alphabet of n <= 30 symbols on one warp, lane = symbol: the same dummy symbols, the
 * same sort order (frequency, then symbol), the same two-queue merge (the leaf on ties), the same repair of lengths
 * beyond maxbits — the same lengths.  Weights, parents and depths live in registers and travel by shuffles; `sw` is 32
 * words of shared scratch of this warp.  Returns the largest symbol with a code; *len_out = the length of symbol `lane`. */
__device__ __forceinline__ int zb_lengths_warp(uint32_t f, int n, int maxbits, uint32_t *sw, uint32_t *len_out)
{
    const uint32_t FULL = 0xFFFFFFFFu;
    const int lane = (int)(threadIdx.x & 31);
    if (lane >= n) f = 0;
    uint32_t usedm = __ballot_sync(FULL, f != 0);
    int m = __popc(usedm);
    int max_code = usedm ? 31 - __clz((int)usedm) : -1;
    while (m < 2) {                                         /* zh_lengths_prepare: at least two symbols get a code */
        const int node = (max_code < 2) ? ++max_code : 0;
        if (lane == node) f = 1;
        usedm |= 1u << node; m++;
    }
    const uint32_t key = (f << 9) | (uint32_t)lane;
    int rank = 0;
    for (int j = 0; j < n; j++) { const uint32_t kj = __shfl_sync(FULL, key, j); rank += (((usedm >> j) & 1u) && kj < key) ? 1 : 0; }
    if (f) sw[rank] = key;
    __syncwarp();
    const uint32_t skey = lane < m ? sw[lane] : 0u;         /* lane i: the i-th smallest key */
    __syncwarp();
    const uint32_t wl = skey >> 9;                          /* leaf weights; wi: lane e holds internal node e */
    uint32_t wi = 0;
    int pl = 0, pi = 0;                                     /* parent (an internal node's number) of leaf `lane` / internal node `lane` */
    int a = 0, b = 0, e = 0;
    for (int it = 0; it < m - 1; it++) {
        uint32_t sum, la = __shfl_sync(FULL, wl, a & 31), ib = __shfl_sync(FULL, wi, b & 31);
        if (a < m && (b >= e || la <= ib)) { if (lane == a) pl = e; sum = la; a++; } else { if (lane == b) pi = e; sum = ib; b++; }
        la = __shfl_sync(FULL, wl, a & 31); ib = __shfl_sync(FULL, wi, b & 31);
        if (a < m && (b >= e || la <= ib)) { if (lane == a) pl = e; sum += la; a++; } else { if (lane == b) pi = e; sum += ib; b++; }
        if (lane == e) wi = sum;
        e++;
    }
    uint32_t di = 0;                                        /* depth of internal node `lane`; the root (e - 1) has 0 */
    for (int i = e - 2; i >= 0; i--) { const int p = __shfl_sync(FULL, pi, i); const uint32_t d = __shfl_sync(FULL, di, p) + 1u; if (lane == i) di = d; }
    uint32_t dl = __shfl_sync(FULL, di, pl) + 1u;
    const bool isleaf = lane < m;
    const bool ov = isleaf && dl > (uint32_t)maxbits;
    if (ov) dl = (uint32_t)maxbits;
    if (__ballot_sync(FULL, ov)) {                          /* zh_repair, bl_count[l] on lane l */
        uint32_t cnt = 0;
        for (int bits = 1; bits <= maxbits; bits++) { const uint32_t c = (uint32_t)__popc(__ballot_sync(FULL, isleaf && dl == (uint32_t)bits)); if (lane == bits) cnt = c; }
        int ex = (lane >= 1 && lane <= maxbits) ? (int)(cnt << (maxbits - lane)) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ex += __shfl_xor_sync(FULL, ex, o);
        int excess = ex - (1 << maxbits);
        while (excess > 0) {
            const uint32_t nz = __ballot_sync(FULL, cnt != 0 && lane >= 1 && lane < maxbits);
            const int bits = 31 - __clz((int)nz);
            if (lane == bits) cnt--;
            if (lane == bits + 1) cnt += 2;
            if (lane == maxbits) cnt--;
            excess--;
        }
        uint32_t dnew = 0, run = 0;
        for (int bits = maxbits; bits >= 1; bits--) { const uint32_t c = __shfl_sync(FULL, cnt, bits); if ((uint32_t)lane >= run && (uint32_t)lane < run + c) dnew = (uint32_t)bits; run += c; }
        dl = dnew;
    }
    sw[lane] = 0;
    __syncwarp();
    if (isleaf) sw[skey & 0x1Fu] = dl;
    __syncwarp();
    *len_out = sw[lane];
    __syncwarp();
    return max_code;
}

/* Sorts key[0, m), m <= 128 * EPT, ascending into key[] (padded with 0xFFFFFFFF to a multiple of four); A: 128 * EPT
 * words of shared scratch.  Thread t of warp w holds the elements w * 32 * EPT + q * 32 + lane, q < EPT. */
template <int EPT>
__device__ __forceinline__ void zb_sort_keys(uint32_t *A, uint32_t *key, int m, uint32_t tid)
{
    static_assert(ZB_THREADS == 128, "four warps");
    constexpr uint32_t N = 128u * EPT;
    const uint32_t lane = tid & 31, base = (tid >> 5) * 32u * EPT + lane;
    uint32_t x[EPT];
#pragma unroll
    for (int q = 0; q < EPT; q++) { const uint32_t i = base + 32u * q; x[q] = (int)i < m ? key[i] : 0xFFFFFFFFu; }
#pragma unroll
    for (uint32_t kk = 2; kk <= N; kk <<= 1) {
#pragma unroll
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            if (j >= 32u * EPT) {                            /* the partner is in another warp */
                __syncthreads();
#pragma unroll
                for (int q = 0; q < EPT; q++) A[base + 32u * q] = x[q];
                __syncthreads();
#pragma unroll
                for (int q = 0; q < EPT; q++) {
                    const uint32_t i = base + 32u * q, y = A[i ^ j];
                    x[q] = (((i & j) == 0u) == ((i & kk) == 0u)) ? min(x[q], y) : max(x[q], y);
                }
            } else if (j >= 32u) {                           /* in another register of this thread */
#pragma unroll
                for (int q = 0; q < EPT; q++)
                    if ((q & (int)(j >> 5)) == 0) {
                        const int r = q | (int)(j >> 5);
                        const uint32_t lo = min(x[q], x[r]), hi = max(x[q], x[r]);
                        const bool up = ((base + 32u * q) & kk) == 0u;
                        x[q] = up ? lo : hi; x[r] = up ? hi : lo;
                    }
            } else {                                         /* in another lane */
#pragma unroll
                for (int q = 0; q < EPT; q++) {
                    const uint32_t i = base + 32u * q, y = __shfl_xor_sync(0xFFFFFFFFu, x[q], (int)j);
                    x[q] = (((i & j) == 0u) == ((i & kk) == 0u)) ? min(x[q], y) : max(x[q], y);
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < EPT; q++) { const uint32_t i = base + 32u * q; if ((int)i < ((m + 3) & ~3)) key[i] = x[q]; }
}

/* The serial recipe of zh_build_block (huff_build.h) spread over the CTA: key collection, sort, leaf depths,
 * costs, canonical codes and the header bit string run on all threads; the distance tree and the 19-symbol
 * code-length tree are built by one warp each (zb_lengths_warp).  The two-queue merge of the literal/length tree — an inherently serial
 * walk of up to 285 steps — is taken out into zs_merge_kernel, where every lane merges a different block
 * (in here it kept one lane busy and 127 waiting for 41 % of a block's time): PHASE 0 is everything before
 * it, PHASE 1 everything after.  Results are identical to the serial form (tests compare the GPU stream with
 * the host model bit for bit). */
template <int PHASE>
__global__ void __launch_bounds__(ZB_THREADS)
zs_block_kernel(const ZsChunk *__restrict__ chunks, const uint32_t *__restrict__ blk_chunk,
                const uint32_t *__restrict__ sym, const uint32_t *__restrict__ chunk_nsym,
                const uint32_t *__restrict__ blk_in_start, zh_block *__restrict__ blocks,
                uint4 *__restrict__ blk_meta, ZbScratch *__restrict__ scratch, uint32_t *__restrict__ used /* [0] = count, then slots */,
                ZsLzParams P, uint32_t slot0 /* first block slot of this launch */)
{
    __shared__ __align__(16) ZbSmem S;
    const uint32_t b = slot0 + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t c = blk_chunk[b];
    const ZsChunk cd = chunks[c];
    const uint32_t k = b - cd.blk_base;
    const uint32_t nsym = chunk_nsym[c];
    uint32_t nblk = (nsym + ZS_BLOCK_SYMS - 1) / ZS_BLOCK_SYMS;
    if (nblk == 0) nblk = 1;
    if (k >= nblk) { if (PHASE == 0 && tid == 0) { blocks[b].type = ZH_UNUSED; blocks[b].flags = 0; blk_meta[b] = make_uint4(ZH_UNUSED, 0, 0, 0); } return; }
    const uint32_t cnt = min(ZS_BLOCK_SYMS, nsym - k * ZS_BLOCK_SYMS);
    const uint32_t *bs = sym + cd.sym_off + (uint64_t)k * ZS_BLOCK_SYMS;
    ZbScratch &X = scratch[b];
    int m;
  if (PHASE == 0) {
    /* ---- histogram ---- */
    for (uint32_t i = tid; i < ZB_WARPS * ZH_LCODES_PAD; i += ZB_THREADS) (&S.lfreq[0][0])[i] = 0;
    for (uint32_t i = tid; i < ZB_WARPS * ZH_DCODES_PAD; i += ZB_THREADS) (&S.dfreq[0][0])[i] = 0;
    if (tid < 16) { S.bl_count[tid] = 0; S.cnt_l[tid] = 0; S.cnt_d[tid] = 0; }
    if (tid == 0) { S.m = 0; S.max_l = 0; S.overflow = 0; }
    for (uint32_t i = tid; i < 256; i += ZB_THREADS) S.lct[i] = (uint16_t)(257 + zs_len_code(i));
    __syncthreads();
    /* four 16-byte loads in flight per thread, then branch-light counting (the symbol arena is 16-byte
       aligned per block: sym_off and ZS_BLOCK_SYMS are multiples of four) */
    {
        const uint4 *bs4 = reinterpret_cast<const uint4 *>(bs);
        for (uint32_t i0 = tid * 4; i0 < cnt; i0 += ZB_THREADS * 16) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + (uint32_t)u * ZB_THREADS * 4;
                v[u] = make_uint4(0, 0, 0, 0);
                if (i < cnt) v[u] = __ldg(bs4 + (i >> 2));
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + (uint32_t)u * ZB_THREADS * 4;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const uint32_t sv = e == 0 ? v[u].x : e == 1 ? v[u].y : e == 2 ? v[u].z : v[u].w;
                    if (i + (uint32_t)e < cnt) {
                        const bool mt = (sv & ZS_MATCH) != 0;
                        const uint32_t d = sv & 0x7FFFu;
                        /* literal: its byte; match: 257 + length code, from a table (ten instructions otherwise) */
                        const uint32_t li = mt ? (uint32_t)S.lct[(sv >> 16) & 0xFFu] : (sv & 0xFFu);
                        const uint32_t nd_ = 31u - (uint32_t)__clz((int)(d | 2u));
                        uint32_t dcode = (nd_ << 1) | ((d >> (nd_ - 1u)) & 1u);
                        dcode = d < 2u ? d : dcode;
                        atomicAdd(&S.lfreq[warp][li], 1u);
                        if (mt) atomicAdd(&S.dfreq[warp][dcode], 1u);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < ZH_LCODES_PAD; i += ZB_THREADS) {
        uint32_t v = 0;
        for (int w = 0; w < ZB_WARPS; w++) v += S.lfreq[w][i];
        v += (i == 256 ? 1u : 0u);
        S.lfreq[0][i] = v;
        S.sc.llen[i] = 0;
        /* ---- literal/length keys (any order: they are sorted next) ---- */
        if (v && i < ZH_LCODES) { int slot = atomicAdd(&S.m, 1); S.sc.key[slot] = (v << 9) | i; atomicMax(&S.max_l, (int)i); }
    }
    if (tid < ZH_DCODES_PAD) {
        uint32_t v = 0;
        for (int w = 0; w < ZB_WARPS; w++) v += S.dfreq[w][tid];
        S.dfreq[0][tid] = v;
    }
    __syncthreads();
    /* fewer than two used symbols (an empty block): the serial routine adds the dummy symbols */
    if (S.m < 2) { if (tid == 0) S.m = zh_lengths_prepare(S.lfreq[0], ZH_LCODES, &S.sc, &S.max_l); __syncthreads(); }
    m = S.m;
    /* ---- the distance tree on one warp (zs_merge_kernel had it as a one-thread job in local memory, the longer of its
            two jobs); the other warps go on to the sort ---- */
    if (warp == ZB_WARPS - 1) {
        uint32_t dl;
        const int md = zb_lengths_warp(S.dfreq[0][lane], ZH_DCODES, 15, S.sc2.tmpfreq, &dl);
        X.dlen[lane] = (uint8_t)dl;
        if (lane == 0) X.max_d = md;
    }
    /* ---- bitonic sort of the keys (unique) on all threads, 128 * EPT of them padded with 0xFFFFFFFF, in registers:
            strides below 32 by shuffles, strides inside a thread by register swaps, only the strides across warps through
            shared memory (3 of the 45 steps of 512 keys) ---- */
    if (m <= 128) zb_sort_keys<1>(S.sc.w, S.sc.key, m, tid);
    else if (m <= 256) zb_sort_keys<2>(S.sc.w, S.sc.key, m, tid);
    else zb_sort_keys<4>(S.sc.w, S.sc.key, m, tid);
    __syncthreads();
    /* ---- the hand-over (the distance tree, a serial job of one thread, is built in zs_merge_kernel where every
       lane has one to build; in here it kept 127 threads waiting for 37 % of the kernel's time) ---- */
    if (tid == 32) { X.m = m; X.max_l = S.max_l; }
    for (uint32_t i = tid; i < ZH_LCODES_PAD; i += ZB_THREADS) { X.key[i] = S.sc.key[i]; X.lfreq[i] = S.lfreq[0][i]; }
    if (tid < ZH_DCODES_PAD) X.dfreq[tid] = S.dfreq[0][tid];
    if (tid == 0) used[1 + atomicAdd(&used[0], 1u)] = b;
    return;
  } else {
    /* ---- PHASE 1: read the hand-over back ---- */
    if (tid < 16) { S.bl_count[tid] = 0; S.cnt_l[tid] = 0; S.cnt_d[tid] = 0; }
    if (tid == 0) { S.m = X.m; S.max_l = X.max_l; S.max_d = X.max_d; S.overflow = 0; }
    for (uint32_t i = tid; i < ZH_LCODES_PAD; i += ZB_THREADS) { S.sc.key[i] = X.key[i]; S.lfreq[0][i] = X.lfreq[i]; S.sc.llen[i] = 0; }
    if (tid < ZH_DCODES_PAD) { S.dfreq[0][tid] = X.dfreq[tid]; S.sc.dlen[tid] = X.dlen[tid]; }
    __syncthreads();
    m = S.m;
  }
    /* ---- leaf depths (from zs_merge_kernel), clipped to 15 ---- */
    for (int i = (int)tid; i < m; i += ZB_THREADS) {
        uint32_t d = X.depth[i];
        if (d > 15) { d = 15; atomicAdd(&S.overflow, 1); }
        S.sc.depth[i] = (uint8_t)d;
        atomicAdd(&S.bl_count[d], 1u);
    }
    __syncthreads();
    if (S.overflow > 0) { if (tid == 0) zh_repair(15, S.bl_count, &S.sc); __syncthreads(); }
    for (int i = (int)tid; i < m; i += ZB_THREADS) S.sc.llen[S.sc.key[i] & 0x1FF] = S.sc.depth[i];
    __syncthreads();
    /* ---- symbol costs under the dynamic and the fixed code ---- */
    {
        uint32_t dyn = 0, fix = 0;
        for (uint32_t i = tid; i < ZH_LCODES; i += ZB_THREADS) {
            uint32_t f = S.lfreq[0][i];
            if (f) {
                uint32_t ex = i >= 257 ? (uint32_t)zh_extra_lbits((int)i - 257) : 0u;
                uint32_t fl = i < 144 ? 8u : i < 256 ? 9u : i < 280 ? 7u : 8u;
                dyn += f * (S.sc.llen[i] + ex); fix += f * (fl + ex);
            }
        }
        if (tid < ZH_DCODES) {
            uint32_t f = S.dfreq[0][tid];
            if (f) { uint32_t ex = (uint32_t)zh_extra_dbits((int)tid); dyn += f * (S.sc.dlen[tid] + ex); fix += f * (5u + ex); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { dyn += __shfl_down_sync(0xFFFFFFFFu, dyn, o); fix += __shfl_down_sync(0xFFFFFFFFu, fix, o); }
        if (lane == 0) { S.red[0][warp] = dyn; S.red[1][warp] = fix; }
    }
    __syncthreads();
    const uint32_t in_start = blk_in_start[b];
    const uint32_t in_end = (k + 1 < nblk) ? blk_in_start[b + 1] : cd.len;
    uint32_t flags = 0;
    if (k == 0 && (cd.flags & ZC_FIRST_OF_STREAM)) flags |= ZB_FIRST_OF_STREAM;
    if (k == nblk - 1 && (cd.flags & ZC_LAST_OF_SECTION)) flags |= ZB_LAST_OF_SECTION;
    if (k == nblk - 1 && (cd.flags & ZC_LAST_OF_STREAM)) flags |= ZB_LAST_OF_STREAM;
    const uint32_t final_block = (flags & ZB_LAST_OF_STREAM) ? 1u : 0u;
    /* ---- RLE of the code lengths into code-length-alphabet tokens, in parallel (same tokens as zh_rle):
            runs of equal lengths -> tokens per run -> prefix sums -> every run writes its own tokens ---- */
    const int nl = max(S.max_l, 256) + 1, nd = S.max_d + 1, nseq = nl + nd;
    uint16_t *run_start = S.sc2.parent;                 /* the distance tree's scratch is free now */
    uint8_t *run_val = S.sc2.depth;
    if (tid <= ZH_BLCODES) S.sc.blfreq[tid] = 0;
    {
        const int i0 = (int)tid * 3;
        uint32_t v[3]; bool st[3]; int nst = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int i = i0 + j;
            v[j] = 0; st[j] = false;
            if (i < nseq) {
                v[j] = i < nl ? S.sc.llen[i] : S.sc.dlen[i - nl];
                const uint32_t pv = (i == 0 || i == nl) ? 0xFFFFu : (i - 1 < nl ? S.sc.llen[i - 1] : S.sc.dlen[i - 1 - nl]);
                st[j] = (pv != v[j]);
                nst += st[j] ? 1 : 0;
            }
        }
        uint32_t inc = (uint32_t)nst;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += t; }
        if (lane == 31) S.scan[warp] = inc;
        __syncthreads();
        uint32_t r = inc - (uint32_t)nst;
        for (uint32_t w = 0; w < warp; w++) r += S.scan[w];
        uint32_t nruns = 0;
        for (uint32_t w = 0; w < ZB_WARPS; w++) nruns += S.scan[w];
#pragma unroll
        for (int j = 0; j < 3; j++) if (st[j]) { run_start[r] = (uint16_t)(i0 + j); run_val[r] = (uint8_t)v[j]; r++; }
        if (tid == 0) run_start[nruns] = (uint16_t)nseq;
        __syncthreads();
        /* tokens per run (the runs this thread found), then their offsets */
        uint32_t r0 = r - (uint32_t)nst, ntk = 0;
        for (uint32_t q = r0; q < r0 + (uint32_t)nst; q++) {
            uint32_t run = (uint32_t)run_start[q + 1] - run_start[q];
            if (run_val[q] == 0) { uint32_t rem = run % 138; ntk += run / 138 + (rem >= 3 ? 1u : rem); }
            else { run--; uint32_t rem = run % 6; ntk += 1 + run / 6 + (rem >= 3 ? 1u : rem); }
        }
        __syncthreads();                                 /* S.scan is reused */
        inc = ntk;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += t; }
        if (lane == 31) S.scan[warp] = inc;
        __syncthreads();
        uint32_t t = inc - ntk;
        for (uint32_t w = 0; w < warp; w++) t += S.scan[w];
        uint32_t total = 0;
        for (uint32_t w = 0; w < ZB_WARPS; w++) total += S.scan[w];
        if (tid == 0) S.D.nt = (int)total;
        for (uint32_t q = r0; q < r0 + (uint32_t)nst; q++) {
            uint32_t run = (uint32_t)run_start[q + 1] - run_start[q];
            const uint32_t val = run_val[q];
            if (val == 0) {
                while (run >= 11) { uint32_t rr = run > 138 ? 138 : run; S.sc.tok[t++] = (uint16_t)(18 | ((rr - 11) << 8)); atomicAdd(&S.sc.blfreq[18], 1u); run -= rr; }
                if (run >= 3) { S.sc.tok[t++] = (uint16_t)(17 | ((run - 3) << 8)); atomicAdd(&S.sc.blfreq[17], 1u); run = 0; }
                if (run) { atomicAdd(&S.sc.blfreq[0], run); while (run-- > 0) S.sc.tok[t++] = 0; }
            } else {
                S.sc.tok[t++] = (uint16_t)val; run--;
                uint32_t lits = 1;
                while (run >= 3) { uint32_t rr = run > 6 ? 6 : run; S.sc.tok[t++] = (uint16_t)(16 | ((rr - 3) << 8)); atomicAdd(&S.sc.blfreq[16], 1u); run -= rr; }
                lits += run;
                while (run-- > 0) S.sc.tok[t++] = (uint16_t)val;
                atomicAdd(&S.sc.blfreq[val], lits);
            }
        }
    }
    __syncthreads();
    /* ---- code-length tree (19 symbols) on one thread, header size on all, decision ---- */
    if (warp == 0) {
        uint32_t bl;
        (void)zb_lengths_warp(lane < ZH_BLCODES ? S.sc.blfreq[lane] : 0u, ZH_BLCODES, 7, S.sc.tmpfreq, &bl);
        if (lane < ZH_BLCODES) S.sc.bllen[lane] = (uint8_t)bl;
        __syncwarp();
    }
    if (tid == 0) {
        const int bl_order[ZH_BLCODES] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        int nbl = ZH_BLCODES;
        while (nbl > 4 && S.sc.bllen[bl_order[nbl - 1]] == 0) nbl--;
        S.D.nbl = nbl; S.D.nl = nl; S.D.nd = nd;
        zh_codes(S.sc.bllen, ZH_BLCODES, S.sc.blcode);
    }
    __syncthreads();
    {
        const int nt_all = S.D.nt;
        uint32_t hb = 0;
        for (int t = (int)tid; t < nt_all; t += ZB_THREADS) { uint32_t nb; (void)zh_tok_bits(S.sc.tok[t], S.sc.blcode, &nb); hb += nb; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hb += __shfl_down_sync(0xFFFFFFFFu, hb, o);
        __syncthreads();
        if (lane == 0) S.scan[warp] = hb;
        __syncthreads();
        if (tid == 0) {
            uint64_t dyn = 0, fix = 0;
            for (int w = 0; w < ZB_WARPS; w++) { dyn += S.red[0][w]; fix += S.red[1][w]; }
            uint32_t hdr = 3 + 5 + 5 + 4 + 3 * (uint32_t)S.D.nbl;
            for (int w = 0; w < ZB_WARPS; w++) hdr += S.scan[w];
            /* same decision rule as zh_decide (reference src/trees.c:902-934) */
            dyn += hdr; fix += 3;
            const uint32_t in_len = in_end - in_start;
            uint64_t opt_lenb = (dyn + 7) >> 3, static_lenb = (fix + 7) >> 3;
            if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
            int type;
            if (P.force_type == ZH_STORED) type = ZH_STORED;
            else if ((uint64_t)in_len + 4 <= opt_lenb && in_len <= 65535u) type = ZH_STORED;
            else if (P.force_type == ZH_STATIC) type = ZH_STATIC;
            else if (static_lenb == opt_lenb || hdr > 32u * ZH_HDR_WORDS - 64u) type = ZH_STATIC;
            else type = ZH_DYNAMIC;
            S.D.type = type; S.D.hdr_bits_est = hdr; S.D.dyn = dyn; S.D.fix = fix;
        }
    }
    for (uint32_t i = tid; i < ZH_HDR_WORDS; i += ZB_THREADS) S.blk.hdr[i] = 0;
    __syncthreads();
    const int type = S.D.type, nbl = S.D.nbl, nt = S.D.nt;
    if (type != ZH_STORED) {
        /* ---- final code lengths, their histogram ---- */
        for (uint32_t i = tid; i < ZH_LCODES_PAD; i += ZB_THREADS) {
            uint32_t l = type == ZH_STATIC ? (i < 144 ? 8u : i < 256 ? 9u : i < 280 ? 7u : 8u) : ((int)i < nl ? S.sc.llen[i] : 0u);
            S.sc.llen[i] = (uint8_t)l;
            if (l) atomicAdd(&S.cnt_l[l], 1u);
        }
        if (tid < ZH_DCODES_PAD) {
            uint32_t l = type == ZH_STATIC ? 5u : ((int)tid < nd ? S.sc.dlen[tid] : 0u);
            S.sc.dlen[tid] = (uint8_t)l;
            if (l) atomicAdd(&S.cnt_d[l], 1u);
        }
        __syncthreads();
        if (tid == 0 || tid == 32) {
            const uint32_t *cn = tid == 0 ? S.cnt_l : S.cnt_d;
            uint32_t *nx = tid == 0 ? S.next_l : S.next_d;
            uint32_t code = 0, prev = 0;
            for (int bits = 1; bits <= 15; bits++) { code = (code + prev) << 1; nx[bits] = code; prev = cn[bits]; }
        }
        __syncthreads();
        /* ---- canonical codes: next_code[len] + rank among the earlier symbols of the same length.
                Ranks come from ballots inside each chunk of 32 symbols plus per-length chunk prefixes ---- */
        uint32_t myrank[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const uint32_t ch = warp + (uint32_t)q * ZB_WARPS;
            if (ch < 9) {
                const uint32_t l = S.sc.llen[ch * 32 + lane];
                for (uint32_t bits = 1; bits <= 15; bits++) {
                    const uint32_t mm = __ballot_sync(0xFFFFFFFFu, l == bits);
                    if (l == bits) myrank[q] = __popc(mm & zs_lanemask_lt());
                    if (lane == bits) S.ccnt[ch][bits] = __popc(mm);
                }
            }
        }
        __syncthreads();
        if (tid >= 1 && tid <= 15) { uint32_t run = 0; for (int ch = 0; ch < 9; ch++) { uint32_t t = S.ccnt[ch][tid]; S.ccnt[ch][tid] = run; run += t; } }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const uint32_t ch = warp + (uint32_t)q * ZB_WARPS;
            if (ch < 9) {
                const uint32_t i = ch * 32 + lane, l = S.sc.llen[i];
                S.blk.lcode[i] = l ? (zh_bitrev(S.next_l[l] + S.ccnt[ch][l] + myrank[q], (int)l) | (l << 16)) : 0u;
            }
        }
        if (tid < ZH_DCODES_PAD) {
            const uint32_t l = S.sc.dlen[tid];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < tid; j++) rank += (S.sc.dlen[j] == l) ? 1u : 0u;
            S.blk.dcode[tid] = l ? (zh_bitrev(S.next_d[l] + rank, (int)l) | (l << 16)) : 0u;
        }
    }
    /* ---- header bit string ---- */
    if (type == ZH_DYNAMIC) {
        const uint32_t base = 3 + 5 + 5 + 4 + 3 * (uint32_t)nbl;
        if (tid == 0) {
            const int bl_order[ZH_BLCODES] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            zh_bitw bw; bw.w = S.blk.hdr; bw.nbits = 0;
            zh_put(&bw, final_block | (2u << 1), 3);
            zh_put(&bw, (uint32_t)(nl - 257), 5);
            zh_put(&bw, (uint32_t)(nd - 1), 5);
            zh_put(&bw, (uint32_t)(nbl - 4), 4);
            for (int i = 0; i < nbl; i++) zh_put(&bw, S.sc.bllen[bl_order[i]], 3);
        }
        /* tokens: every thread owns a contiguous run, offsets from a block-wide prefix sum */
        const int per = (nt + ZB_THREADS - 1) / ZB_THREADS;
        const int t0 = min(nt, (int)tid * per), t1 = min(nt, t0 + per);
        uint32_t mine = 0;
        for (int t = t0; t < t1; t++) { uint32_t nb; (void)zh_tok_bits(S.sc.tok[t], S.sc.blcode, &nb); mine += nb; }
        uint32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += v; }
        if (lane == 31) S.scan[warp] = inc;
        __syncthreads();                                   /* also orders thread 0's zh_put before the atomics */
        uint32_t pos = base + inc - mine;
        for (uint32_t w = 0; w < warp; w++) pos += S.scan[w];
        for (int t = t0; t < t1; t++) {
            uint32_t nb, v = zh_tok_bits(S.sc.tok[t], S.sc.blcode, &nb);
            const uint32_t wi = pos >> 5, sh = pos & 31;
            atomicOr(&S.blk.hdr[wi], v << sh);
            if (sh + nb > 32) atomicOr(&S.blk.hdr[wi + 1], v >> (32 - sh));
            pos += nb;
        }
        if (tid == ZB_THREADS - 1) { S.blk.hdr_bits = pos; S.blk.body_bits = (uint32_t)S.D.dyn; }
    } else if (tid == 0) {
        S.blk.hdr[0] = final_block | ((type == ZH_STATIC ? 1u : 0u) << 1);
        S.blk.hdr_bits = 3;
        S.blk.body_bits = type == ZH_STATIC ? (uint32_t)S.D.fix : 0u;
    }
    if (tid == 0) {
        S.blk.type = (uint32_t)type;
        S.blk.in_len = in_end - in_start;
        S.blk.nsym = cnt;
        S.blk.in_start = in_start;
        S.blk.flags = flags;
        S.blk.stored_total = in_end - in_start;
        S.blk.bitoff = 0;
        S.blk.sym_off = cd.sym_off + (uint64_t)k * ZS_BLOCK_SYMS;
    }
    __syncthreads();
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&S.blk);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&blocks[b]);
    for (uint32_t i = tid; i < sizeof(zh_block) / 4; i += ZB_THREADS) dst[i] = src[i];
    /* what the offset pass needs, 16 bytes per block instead of a walk over the large records */
    if (tid == 0) blk_meta[b] = make_uint4(S.blk.type, S.blk.body_bits, S.blk.in_len, S.blk.flags);
}

/* ======================= K2m: the serial tree work, one block per lane ======================= */
#define ZMG_THREADS 64

/* The one inherently serial job of a block, done where every lane has one to do: the two-queue merge of the
 * literal/length tree.  Leaves are the sorted keys (weight = key >> 9 <= 8193, 16 bits); they are first copied from the
 * hand-over into shared memory — in the loop they would be dependent global loads, one L2 round trip per merge step —
 * next to the weights of the internal nodes, which the merge both appends and consumes in order; one column per lane.  Same picks as zh_merge (huff_build.h): the smaller head of the two queues, the leaf on
 * ties.  (The distance tree and the code-length tree, 30 and 19 symbols, are built by a warp each inside
 * zs_block_kernel<0> / <1>: zb_lengths_warp.) */
__global__ void __launch_bounds__(ZMG_THREADS)
zs_merge_kernel(ZbScratch *__restrict__ scratch, const uint32_t *__restrict__ used)
{
    extern __shared__ uint16_t zmg_w[];                    /* [2][ZH_LCODES_PAD][ZMG_THREADS]: leaves, internal nodes */
    const uint32_t col = threadIdx.x;
    const uint32_t t = blockIdx.x * ZMG_THREADS + col;
    if (t >= used[0]) return;
    ZbScratch &X = scratch[used[1 + t]];
    const int m = X.m;
    uint16_t *wl = zmg_w + col;
    uint16_t *w = zmg_w + ZH_LCODES_PAD * ZMG_THREADS + col;
    {
        const uint4 *k4 = reinterpret_cast<const uint4 *>(X.key);
#pragma unroll 4
        for (int j = 0; j < (m + 3) >> 2; j++) {
            const uint4 q = k4[j];
            wl[(4 * j + 0) * ZMG_THREADS] = (uint16_t)(q.x >> 9); wl[(4 * j + 1) * ZMG_THREADS] = (uint16_t)(q.y >> 9);
            wl[(4 * j + 2) * ZMG_THREADS] = (uint16_t)(q.z >> 9); wl[(4 * j + 3) * ZMG_THREADS] = (uint16_t)(q.w >> 9);
        }
    }
    /* A node that leaves its queue leaves a free slot behind: its parent's number goes there (shared memory; as scattered
       16-bit stores to global memory the links were what this kernel waited for).  Afterwards the slots of the internal
       nodes turn from parents into depths, root first, and the depths of the leaves go out four to a word. */
    int a = 0, b = 0, e = 0;                               /* leaves taken, internal nodes taken, internal nodes made */
    uint32_t la = wl[0], ib = 0xFFFFFFFFu;                 /* heads of the two queues */
    for (int it = 0; it < m - 1; it++) {
        uint32_t sum;
        if (a < m && la <= ib) { sum = la; wl[a * ZMG_THREADS] = (uint16_t)e; a++; la = a < m ? wl[a * ZMG_THREADS] : 0xFFFFFFFFu; }
        else { sum = ib; w[b * ZMG_THREADS] = (uint16_t)e; b++; ib = b < e ? w[b * ZMG_THREADS] : 0xFFFFFFFFu; }
        if (a < m && la <= ib) { sum += la; wl[a * ZMG_THREADS] = (uint16_t)e; a++; la = a < m ? wl[a * ZMG_THREADS] : 0xFFFFFFFFu; }
        else { sum += ib; w[b * ZMG_THREADS] = (uint16_t)e; b++; ib = b < e ? w[b * ZMG_THREADS] : 0xFFFFFFFFu; }
        w[e * ZMG_THREADS] = (uint16_t)sum;
        if (b == e) ib = sum;                              /* the queue was empty: the new node is its head */
        e++;
    }
    w[(e - 1) * ZMG_THREADS] = 0;                          /* the root */
    for (int i = e - 2; i >= 0; i--) { const uint32_t p = w[i * ZMG_THREADS]; w[i * ZMG_THREADS] = (uint16_t)(w[p * ZMG_THREADS] + 1u); }
    uint32_t *d32 = reinterpret_cast<uint32_t *>(X.depth);
    for (int i = 0; i < m; i += 4) {
        uint32_t v = 0;
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i + u < m) { const uint32_t d = (uint32_t)w[(uint32_t)wl[(i + u) * ZMG_THREADS] * ZMG_THREADS] + 1u; v |= (d > 255u ? 255u : d) << (8 * u); }
        d32[i >> 2] = v;
    }
}

/* ======================= K2b: stored-run merging (thread per chunk), offsets (one CTA per stream) ======================= */
#define ZM_THREADS 128
#define ZO_THREADS_MAX 1024

/* Merge runs of adjacent stored blocks of a chunk into one stored block of <= 65535 bytes: the first keeps
 * the header with the run's length, the others become payload-only continuations.  Keeps incompressible data
 * at 5 bytes per 64 KiB, inside the reference's output bound. */
__global__ void __launch_bounds__(ZM_THREADS)
zs_stored_merge_kernel(const ZsChunk *__restrict__ chunks, uint32_t nchunks, zh_block *__restrict__ blocks, uint4 *__restrict__ blk_meta)
{
    const uint32_t c = blockIdx.x * ZM_THREADS + threadIdx.x;
    if (c >= nchunks) return;
    const uint32_t base = chunks[c].blk_base, capn = chunks[c].blk_cap;
    uint32_t head = 0xFFFFFFFFu, total = 0;
    for (uint32_t k = 0; k < capn; k++) {
        uint4 mt = blk_meta[base + k];                      /* type, body_bits, in_len, flags */
        if (mt.x == ZH_UNUSED) break;
        if (mt.x == ZH_STORED) {
            if (head != 0xFFFFFFFFu && total + mt.z <= 65535u) {
                mt.x = ZH_STORED_CONT;
                blk_meta[base + k] = mt;
                blocks[base + k].type = ZH_STORED_CONT;
                total += mt.z;
                blocks[head].stored_total = total;
                if (mt.w & ZB_LAST_OF_STREAM) blocks[head].hdr[0] |= 1u;
            } else { head = base + k; total = mt.z; }
        } else head = 0xFFFFFFFFu;
        if (mt.w & (ZB_LAST_OF_SECTION | ZB_LAST_OF_STREAM)) head = 0xFFFFFFFFu;
    }
}

__device__ __forceinline__ zk_elem zk_shfl_up(zk_elem e, int o)
{
    zk_elem r;
    r.a = __shfl_up_sync(0xFFFFFFFFu, e.a, o); r.b = __shfl_up_sync(0xFFFFFFFFu, e.b, o);
    r.al = __shfl_up_sync(0xFFFFFFFFu, e.al, o); r.pad = 0;
    return r;
}

/* PHASE 2: a stream's whole offset scan on one CTA.  PHASE 0 / 1: one large stream cut into gridDim.y parts of consecutive
 * block slots (the one CTA of a 1 GiB stream spent 0.2 ms on its 131 072 strided 16-byte records — one SM's load/store
 * unit, a sector per request): phase 0 leaves the composed element of every part in part_total, phase 1 repeats the scan
 * of its part, puts the parts before it in front and writes the offsets; part 0 also writes the stream's results. */
#define ZO_PARTS_MAX 64
template <int PHASE>
__global__ void __launch_bounds__(ZO_THREADS_MAX)
zs_offset_kernel(const ZsStream *__restrict__ streams, uint4 *__restrict__ blk_meta, uint64_t *__restrict__ blk_bitoff,
                 const ZsAdlerAcc *__restrict__ adler_acc, uint32_t *__restrict__ comp32,
                 int32_t *__restrict__ res_ret, uint32_t *__restrict__ res_produced,
                 uint32_t *__restrict__ res_check, ZsLzParams P, zk_elem *__restrict__ part_total)
{
    __shared__ zk_elem wpart[32];
    __shared__ zk_elem s_pre;
    __shared__ int s_fail;
    const uint32_t sidx = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t part = PHASE == 2 ? 0u : blockIdx.y, nparts = PHASE == 2 ? 1u : gridDim.y;
    const ZsStream st = streams[sidx];
    const uint32_t per_part = (st.blk_count + nparts - 1) / nparts;
    const uint32_t p_lo = min(st.blk_count, part * per_part), p_hi = min(st.blk_count, p_lo + per_part);
    const uint32_t per = (p_hi - p_lo + nthr - 1) / nthr;
    const uint32_t lo = min(p_hi, p_lo + tid * per), hi = min(p_hi, lo + per);
    /* a thread's records are read eight at a time (the loop is a chain of compositions) */
    zk_elem mine = zk_ident();
    for (uint32_t i = lo; i < hi; i += 8) {
        uint4 mt[8];
#pragma unroll
        for (uint32_t u = 0; u < 8; u++) mt[u] = (i + u < hi) ? blk_meta[st.blk_first + i + u] : make_uint4(ZH_UNUSED, 0, 0, 0);
#pragma unroll
        for (uint32_t u = 0; u < 8; u++)
            if (i + u < hi) mine = zk_compose(mine, zk_elem_of_block(mt[u].x, mt[u].y, mt[u].z, mt[u].w, P.wrap));
    }
    /* exclusive scan of the per-thread elements (composition is associative, not commutative) */
    zk_elem inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { zk_elem t = zk_shfl_up(inc, o); if ((int)lane >= o) inc = zk_compose(t, inc); }
    if (lane == 31) wpart[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t nw = nthr >> 5;
        zk_elem w = lane < nw ? wpart[lane] : zk_ident();
        zk_elem wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { zk_elem t = zk_shfl_up(wi, o); if ((int)lane >= o) wi = zk_compose(t, wi); }
        zk_elem wex = zk_shfl_up(wi, 1);
        if (lane == 0) wex = zk_ident();
        wpart[lane] = wex;                                   /* exclusive prefix of the warp totals */
        if (lane == 31) {
            if (PHASE == 0) part_total[sidx * ZO_PARTS_MAX + part] = wi;
            else {
                zk_elem pre = zk_ident(), all = wi;          /* the parts before this one; the whole stream */
                if (PHASE == 1) {
                    for (uint32_t q = 0; q < part; q++) pre = zk_compose(pre, part_total[sidx * ZO_PARTS_MAX + q]);
                    all = zk_compose(pre, wi);
                    for (uint32_t q = part + 1; q < nparts; q++) all = zk_compose(all, part_total[sidx * ZO_PARTS_MAX + q]);
                }
                s_pre = pre;
                const uint64_t x0 = st.comp_off * 8ull;
                const uint64_t xe = zk_apply(all, x0);
                const uint64_t bytes = (xe - x0 + 7) >> 3;
                const int fail = bytes > st.comp_cap;
                s_fail = fail;
                if (part == 0) {
                    res_ret[sidx] = fail ? -5 /* Z_BUF_ERROR */ : 0;
                    res_produced[sidx] = fail ? 0u : (uint32_t)bytes;
                    uint32_t a = (uint32_t)(adler_acc[sidx].s1 % ZS_ADLER_BASE), bsum = (uint32_t)(adler_acc[sidx].s2 % ZS_ADLER_BASE);
                    /* adler32 of the stream = (1 + sum bytes, len + weighted sum) mod 65521 */
                    a = (a + 1) % ZS_ADLER_BASE;
                    bsum = (uint32_t)((bsum + (uint64_t)st.raw_len) % ZS_ADLER_BASE);
                    res_check[sidx] = (bsum << 16) | a;
                    if (!fail && (xe & 31)) comp32[xe >> 5] = 0;
                }
            }
        }
    }
    if (PHASE == 0) return;
    __syncthreads();
    zk_elem before = zk_shfl_up(inc, 1);
    if (lane == 0) before = zk_ident();
    before = zk_compose(s_pre, zk_compose(wpart[warp], before));
    const int fail = s_fail;
    uint64_t x = zk_apply(before, st.comp_off * 8ull);
    for (uint32_t i0 = lo; i0 < hi; i0 += 8) {
        uint4 m8[8];
#pragma unroll
        for (uint32_t u = 0; u < 8; u++) m8[u] = (i0 + u < hi) ? blk_meta[st.blk_first + i0 + u] : make_uint4(ZH_UNUSED, 0, 0, 0);
#pragma unroll
        for (uint32_t u = 0; u < 8; u++) {
            const uint32_t i = i0 + u;
            uint4 mt = m8[u];
            if (mt.x == ZH_UNUSED) continue;                 /* (also the records beyond hi) */
            blk_bitoff[st.blk_first + i] = x;
            if (fail) { mt.w |= ZB_STREAM_FAILED; blk_meta[st.blk_first + i] = mt; } else comp32[x >> 5] = 0;
            x = zk_apply(zk_elem_of_block(mt.x, mt.y, mt.z, mt.w & ~(uint32_t)ZB_STREAM_FAILED, P.wrap), x);
        }
    }
}

/* ======================= K3: bit packing, one CTA per block ======================= */
#define ZE_THREADS 256
#ifndef ZE_SUB
#define ZE_SUB 8                                    /* symbols per thread and round (16: no gain, measured) */
#endif
#define ZE_ROUNDS (ZS_BLOCK_SYMS / (ZE_THREADS * ZE_SUB))
#define ZE_STAGE_WORDS 12480                        /* 16 + 3072 + 8192*48 + 15 + 47 bits, rounded up */

struct ZeSmem {
    uint32_t stage[ZE_STAGE_WORDS];
    uint32_t lcode[ZH_LCODES_PAD];
    uint32_t dcode[ZH_DCODES_PAD];
    uint32_t tab[516];          /* code bits | count << 24 of: a literal byte [0, 256); a match length - 3 with its extra bits [256, 512); no symbol [512] */
    uint32_t wsum[2][ZE_THREADS / 32];
    uint32_t total_bits;
};

static_assert(ZS_BLOCK_SYMS == ZE_THREADS * ZE_SUB * ZE_ROUNDS && ZE_SUB % 4 == 0, "ZE_ROUNDS rounds of ZE_SUB symbols per thread");

__device__ __forceinline__ void ze_or_bits(uint32_t *stage, uint32_t pos, uint64_t v, uint32_t n)
{
    /* n <= 33 bits at local bit position pos, shared-memory atomics (used at piece boundaries) */
    if (n == 0) return;
    uint32_t w = pos >> 5, sh = pos & 31;
    unsigned long long t = (unsigned long long)v << sh;        /* n + sh <= 64 */
    atomicOr(&stage[w], (uint32_t)t);
    if (sh + n > 32) atomicOr(&stage[w + 1], (uint32_t)(t >> 32));
}

/* distance part of a match: code + extra bits (<= 28 bits) and their count */
__device__ __forceinline__ void ze_dist_bits(const ZeSmem &S, uint32_t s, uint32_t &v1, uint32_t &n1)
{
    const uint32_t d = s & 0x7FFF;
    uint32_t dc, eb, ev;
    if (d < 4) { dc = d; eb = 0; ev = 0; }
    else { uint32_t n = 31u - (uint32_t)__clz((int)d); dc = (n << 1) | ((d >> (n - 1)) & 1u); eb = n - 1; ev = d & ((1u << eb) - 1u); }
    const uint32_t e = S.dcode[dc];
    const uint32_t cl = e >> 16;
    v1 = (e & 0xFFFFu) | (ev << cl); n1 = cl + eb;
}

/* write one byte of the comp arena; words shared with neighbouring blocks go through atomics */
__device__ __forceinline__ void ze_put_byte(uint8_t *comp, uint64_t pos, uint32_t v, uint64_t w_first, uint64_t w_last)
{
    uint64_t w = pos >> 2;
    if (w == w_first || w == w_last) atomicOr(reinterpret_cast<uint32_t *>(comp) + w, v << (8 * (uint32_t)(pos & 3)));
    else comp[pos] = (uint8_t)v;
}

__global__ void __launch_bounds__(ZE_THREADS, 4)
zs_encode_kernel(const zh_block *__restrict__ blocks, const uint32_t *__restrict__ blk_chunk,
                 const ZsChunk *__restrict__ chunks, const uint32_t *__restrict__ sym,
                 const uint8_t *__restrict__ raw, uint8_t *__restrict__ comp,
                 const uint32_t *__restrict__ res_check, const uint4 *__restrict__ blk_meta,
                 const uint64_t *__restrict__ blk_bitoff, ZsLzParams P)
{
    extern __shared__ __align__(16) unsigned char ze_smem_raw[];
    ZeSmem &S = *reinterpret_cast<ZeSmem *>(ze_smem_raw);
    const uint32_t b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const zh_block *bp = &blocks[b];
    const uint4 mt = blk_meta[b];                          /* type, body_bits, in_len, flags */
    const uint32_t type = mt.x, flags = mt.w;
    if (type == ZH_UNUSED || (flags & ZB_STREAM_FAILED)) return;
    const ZsChunk cd = chunks[blk_chunk[b]];
    const uint64_t x = blk_bitoff[b];
    const zk_elem el = zk_elem_of_block(type, mt.y, mt.z, flags, P.wrap);
    const uint64_t xe = zk_apply(el, x);
    const uint64_t w_first = x >> 5, w_last = (xe & 31) ? (xe >> 5) : ~0ull;
    const uint32_t hdr_bits = bp->hdr_bits;
    const uint32_t pre_bits = ((flags & ZB_FIRST_OF_STREAM) && P.wrap == 1) ? 16u : 0u;
    uint32_t *comp32 = reinterpret_cast<uint32_t *>(comp);

    /* suffix value: full-flush marker (00 00 FF FF) or big-endian adler32 */
    uint32_t suf_bits = 0, suf_val = 0, suf_pre = 0;
    if (flags & ZB_LAST_OF_STREAM) { if (P.wrap == 1) { suf_bits = 32; suf_val = __byte_perm(res_check[cd.stream], 0, 0x0123); } }
    else if (flags & ZB_LAST_OF_SECTION) { suf_pre = 3; suf_bits = 32; suf_val = 0xFFFF0000u; }

    if (type == ZH_STORED || type == ZH_STORED_CONT) {
        /* [stream header][3-bit block header][pad][LEN][NLEN][bytes][suffix]: byte-granular writes.
           A continuation of a merged run is payload (+ suffix) only and starts byte aligned. */
        const uint32_t in_len = mt.z;
        const uint32_t run_len = bp->stored_total;
        const uint64_t hb = x + pre_bits;                 /* bit position of the block header */
        const uint64_t d0 = type == ZH_STORED ? (zk_up8(hb + 3) >> 3) + 4 : (x >> 3);   /* first payload byte */
        if (tid == 0) {
            if (type == ZH_STORED) {
                /* leading bits: zlib header (if any) + 3 header bits, then zero pad */
                uint64_t v = (pre_bits ? (uint64_t)(uint32_t)P.zhdr : 0ull) | ((uint64_t)(bp->hdr[0] & 7u) << pre_bits);
                uint32_t sh = (uint32_t)(x & 7);
                v <<= sh;
                for (uint64_t p = x >> 3; p < d0 - 4; p++) { ze_put_byte(comp, p, (uint32_t)(v & 0xFF), w_first, w_last); v >>= 8; }
                ze_put_byte(comp, d0 - 4, run_len & 0xFF, w_first, w_last);
                ze_put_byte(comp, d0 - 3, (run_len >> 8) & 0xFF, w_first, w_last);
                ze_put_byte(comp, d0 - 2, (~run_len) & 0xFF, w_first, w_last);
                ze_put_byte(comp, d0 - 1, ((~run_len) >> 8) & 0xFF, w_first, w_last);
            }
            uint64_t p = d0 + in_len;
            if (suf_pre) { ze_put_byte(comp, p, 0, w_first, w_last); p++; }
            for (uint32_t i = 0; i < suf_bits / 8; i++) ze_put_byte(comp, p + i, (suf_val >> (8 * i)) & 0xFF, w_first, w_last);
        }
        const uint8_t *src = raw + cd.raw_off + bp->in_start;
        for (uint32_t i = tid; i < in_len; i += ZE_THREADS) ze_put_byte(comp, d0 + i, src[i], w_first, w_last);
        return;
    }

    /* ---- compressed block: build the bit image in shared memory ---- */
    const uint32_t lbase = (uint32_t)(x & 31);            /* local bit position of the element start */
    const uint32_t nsym = bp->nsym;
    for (uint32_t i = tid; i < ZH_LCODES_PAD; i += ZE_THREADS) S.lcode[i] = bp->lcode[i];
    if (tid < ZH_DCODES_PAD) S.dcode[tid] = bp->dcode[tid];
    const uint32_t total_words = (uint32_t)(((xe - (x & ~31ull)) + 31) >> 5);
    for (uint32_t i = tid; i < total_words + 1 && i < ZE_STAGE_WORDS; i += ZE_THREADS) S.stage[i] = 0;
    __syncthreads();
    {
        /* length symbol -> packed code + extra bits */
        uint32_t lc = tid;                                 /* ZE_THREADS == 256 */
        int c = zs_len_code(lc);
        uint32_t e = S.lcode[257 + c];
        uint32_t cl = e >> 16;
        uint32_t eb = (uint32_t)zh_extra_lbits(c);
        uint32_t ev = eb ? (lc & ((1u << eb) - 1u)) : 0u;
        S.tab[256 + lc] = ((e & 0xFFFFu) | (ev << cl)) | ((cl + eb) << 24);
        const uint32_t el = S.lcode[lc];                   /* literal byte lc */
        S.tab[lc] = (el & 0xFFFFu) | ((el >> 16) << 24);
        if (tid == 0) S.tab[512] = 0;
    }
    /* stream header + block header bits */
    if (tid == 0 && pre_bits) ze_or_bits(S.stage, lbase, (uint32_t)P.zhdr, 16);
    {
        const uint32_t hpos = lbase + pre_bits;
        const uint32_t hwords = (hdr_bits + 31) >> 5;
        for (uint32_t i = tid; i < hwords; i += ZE_THREADS) {
            uint32_t nb = min(32u, hdr_bits - i * 32);
            uint32_t v = bp->hdr[i];
            if (nb < 32) v &= (1u << nb) - 1u;
            ze_or_bits(S.stage, hpos + i * 32, v, nb);
        }
    }
    __syncthreads();

    /* ---- the block in ZE_ROUNDS rounds of ZE_SUB symbols per thread.  A round computes the code bits of its symbols ONCE
            and keeps them in registers (two words per symbol: length part | bits << 24, distance part under a sentinel
            bit), a block-wide prefix sum of the bit counts gives every thread its position, then the thread packs its
            run: only the first and the last word of a run can be shared with a neighbour ---- */
    const uint32_t *bs = sym + bp->sym_off;
    const uint32_t sympos = lbase + pre_bits + hdr_bits;
    uint32_t base = sympos;                              /* bit position of the round's first symbol */
#pragma unroll 1
    for (uint32_t r = 0; r < ZE_ROUNDS; r++) {
        const uint32_t s0 = (r * ZE_THREADS + tid) * ZE_SUB;
        if (r * ZE_THREADS * ZE_SUB >= nsym) break;      /* uniform */
        uint32_t ca[ZE_SUB], cb[ZE_SUB];
        uint32_t mybits = 0;
        {
            uint4 v[ZE_SUB / 4];
#pragma unroll
            for (int i = 0; i < ZE_SUB / 4; i++) {
                v[i] = make_uint4(0, 0, 0, 0);
                if (s0 + i * 4 < nsym) v[i] = __ldg(reinterpret_cast<const uint4 *>(bs + s0) + i);
            }
#pragma unroll
            for (int i = 0; i < ZE_SUB; i++) {
                const uint32_t sv = (i & 3) == 0 ? v[i >> 2].x : (i & 3) == 1 ? v[i >> 2].y : (i & 3) == 2 ? v[i >> 2].z : v[i >> 2].w;
                /* first part: one table for literals and match lengths, read by every lane; the distance part only by matches */
                const bool ok = s0 + i < nsym, mt = ok && (sv & ZS_MATCH);
                const uint32_t t = S.tab[!ok ? 512u : (sv & ZS_MATCH) ? 256u + ((sv >> 16) & 0xFFu) : (sv & 0xFFu)];
                uint32_t c = 1u;
                if (mt) { uint32_t v1, n1; ze_dist_bits(S, sv, v1, n1); c = v1 | (1u << n1); mybits += n1; }
                ca[i] = t; cb[i] = c;
                mybits += t >> 24;
            }
        }
        uint32_t inc = mybits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += t; }
        if (lane == 31) S.wsum[r & 1][warp] = inc;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (uint32_t w = 0; w < ZE_THREADS / 32; w++) { const uint32_t t = S.wsum[r & 1][w]; total += t; if (w < warp) wbase += t; }
        uint32_t pos = base + wbase + inc - mybits;
        base += total;
        if (mybits) {
            uint32_t w = pos >> 5;
            uint32_t fill = pos & 31;
            unsigned long long acc = 0;
            bool first = true;
#pragma unroll
            for (int i = 0; i < ZE_SUB; i++) {
                acc |= (unsigned long long)(ca[i] & 0xFFFFFFu) << fill; fill += ca[i] >> 24;
                if (fill >= 32) {
                    if (first) { atomicOr(&S.stage[w], (uint32_t)acc); first = false; } else S.stage[w] = (uint32_t)acc;
                    w++; acc >>= 32; fill -= 32;
                }
                if (cb[i] > 1u) {
                    const uint32_t n1 = 31u - (uint32_t)__clz((int)cb[i]);
                    acc |= (unsigned long long)(cb[i] ^ (1u << n1)) << fill; fill += n1;
                    if (fill >= 32) {
                        if (first) { atomicOr(&S.stage[w], (uint32_t)acc); first = false; } else S.stage[w] = (uint32_t)acc;
                        w++; acc >>= 32; fill -= 32;
                    }
                }
            }
            if (fill) atomicOr(&S.stage[w], (uint32_t)acc);
        }
    }
    if (tid == 0) S.total_bits = base - sympos;
    __syncthreads();
    if (tid == 0) {
        /* end-of-block, then the suffix */
        uint32_t p = sympos + S.total_bits;
        uint32_t e = S.lcode[256];
        ze_or_bits(S.stage, p, e & 0xFFFFu, e >> 16);
        p += e >> 16;
        if (flags & (ZB_LAST_OF_STREAM | ZB_LAST_OF_SECTION)) {
            /* positions are relative to a word boundary, so byte alignment is preserved */
            p += suf_pre;
            p = (p + 7) & ~7u;
            ze_or_bits(S.stage, p, suf_val, suf_bits);
        }
    }
    __syncthreads();
    /* ---- flush: interior words are owned by this block; boundary words are OR-ed ---- */
    for (uint32_t i = tid; i < total_words; i += ZE_THREADS) {
        uint64_t w = w_first + i;
        uint32_t v = S.stage[i];
        if (w == w_first || w == w_last) atomicOr(&comp32[w], v);
        else comp32[w] = v;
    }
}

extern "C" size_t zs_encode_smem_bytes(void) { return sizeof(ZeSmem); }
extern "C" size_t zs_block_scratch_bytes(void) { return sizeof(ZbScratch); }
extern "C" size_t zs_offset_part_bytes(void) { return sizeof(zk_elem) * ZO_PARTS_MAX; }

/* The block stage (histogram + sort, tree merges, codes + header) of the block slots [slot0, slot0 + nslots). */
extern "C" cudaError_t zs_block_stage_launch(cudaStream_t st, uint32_t slot0, uint32_t nslots,
                                             const ZsChunk *chunks, const uint32_t *blk_chunk, const uint32_t *sym,
                                             const uint32_t *chunk_nsym, const uint32_t *blk_in_start, zh_block *blocks,
                                             ZsLzParams P, void *blk_meta_v, void *blk_scratch_v, uint32_t *blk_used)
{
    if (nslots == 0) return cudaSuccess;
    uint4 *blk_meta = reinterpret_cast<uint4 *>(blk_meta_v);
    ZbScratch *scratch = reinterpret_cast<ZbScratch *>(blk_scratch_v);
    cudaMemsetAsync(blk_used, 0, 4, st);
    zs_block_kernel<0><<<nslots, ZB_THREADS, 0, st>>>(chunks, blk_chunk, sym, chunk_nsym, blk_in_start, blocks, blk_meta, scratch, blk_used, P, slot0);
    {
        const size_t smem = sizeof(uint16_t) * 2 * ZH_LCODES_PAD * ZMG_THREADS;
        cudaFuncSetAttribute(zs_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        zs_merge_kernel<<<(nslots + ZMG_THREADS - 1) / ZMG_THREADS, ZMG_THREADS, smem, st>>>(scratch, blk_used);
    }
    zs_block_kernel<1><<<nslots, ZB_THREADS, 0, st>>>(chunks, blk_chunk, sym, chunk_nsym, blk_in_start, blocks, blk_meta, scratch, blk_used, P, slot0);
    return cudaGetLastError();
}

/* Everything behind the block stage: stored-run merging, the offset scan, bit packing. */
extern "C" cudaError_t zs_huff_launch(cudaStream_t st, uint32_t nblk_slots, uint32_t nstreams,
                                      const ZsChunk *chunks, const uint32_t *blk_chunk,
                                      const ZsStream *streams, const uint32_t *sym,
                                      zh_block *blocks, const ZsAdlerAcc *adler_acc,
                                      const uint8_t *raw, uint8_t *comp, int32_t *res_ret,
                                      uint32_t *res_produced, uint32_t *res_check, ZsLzParams P,
                                      cudaEvent_t ev_after_offset,
                                      uint32_t nchunks, void *blk_meta_v, unsigned long long *blk_bitoff_v, void *off_part_v /* zs_offset_part_bytes() */)
{
    if (nblk_slots == 0 || nstreams == 0) return cudaSuccess;
    uint4 *blk_meta = reinterpret_cast<uint4 *>(blk_meta_v);
    uint64_t *blk_bitoff = reinterpret_cast<uint64_t *>(blk_bitoff_v);
    zs_stored_merge_kernel<<<(nchunks + ZM_THREADS - 1) / ZM_THREADS, ZM_THREADS, 0, st>>>(chunks, nchunks, blocks, blk_meta);
    /* few streams with many blocks each: wide CTAs; many small streams: narrow ones */
    const uint32_t othreads = (nblk_slots / nstreams >= 1024u) ? ZO_THREADS_MAX : 128u;
    zk_elem *part_total = reinterpret_cast<zk_elem *>(off_part_v);
    if (nstreams == 1 && nblk_slots >= 16384u && part_total) {
        /* one large stream: its slots in parts of about 2048 */
        const dim3 grid(1, min((uint32_t)ZO_PARTS_MAX, nblk_slots / 2048u));
        zs_offset_kernel<0><<<grid, 256, 0, st>>>(streams, blk_meta, blk_bitoff, adler_acc, reinterpret_cast<uint32_t *>(comp), res_ret, res_produced, res_check, P, part_total);
        zs_offset_kernel<1><<<grid, 256, 0, st>>>(streams, blk_meta, blk_bitoff, adler_acc, reinterpret_cast<uint32_t *>(comp), res_ret, res_produced, res_check, P, part_total);
    } else
        zs_offset_kernel<2><<<nstreams, othreads, 0, st>>>(streams, blk_meta, blk_bitoff, adler_acc,
                                                           reinterpret_cast<uint32_t *>(comp), res_ret, res_produced, res_check, P, part_total);
    if (ev_after_offset) cudaEventRecord(ev_after_offset, st);
    cudaFuncSetAttribute(zs_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZeSmem));
    cudaFuncSetAttribute(zs_encode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    zs_encode_kernel<<<nblk_slots, ZE_THREADS, sizeof(ZeSmem), st>>>(blocks, blk_chunk, chunks, sym, raw, comp, res_check, blk_meta, blk_bitoff, P);
    return cudaGetLastError();
}
