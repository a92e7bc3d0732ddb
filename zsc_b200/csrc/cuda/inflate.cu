/* inflate.cu — batched inflate: a group of lanes per stream, decode tables in shared memory (sm_100a).
 *
 * GPU form of zsc_uncompress's hot loop (reference src/zsc_uncompr.c:103-127 -> inflate /
 * inflate_fast / inflate_table).  Streams are independent, so a batch is spread one stream per warp (or
 * half warp when the batch is wide): the group's leader lane decodes symbols into a small queue, the
 * whole group writes them out, and the generic state machine of inflate_core.h handles every header, block
 * boundary, buffer end and error exactly as the CPU tests pin it.  The data check (adler32 of the output
 * against the trailer, reference src/inflate.c:1322-1342) is a second, HBM-streaming pass over the output.
 */
#include <stdlib.h>
#include "common.cuh"
#include "inflate_core.h"

#ifndef ZS_INFLATE_WARP_MAX
#define ZS_INFLATE_WARP_MAX 6144u                /* streams in a batch up to which each gets a whole warp */
#endif

/* ======================= a group of G lanes per stream ======================= */
/* The leader lane of a group runs the state machine.  Inside a compressed block it decodes up to G symbols
 * into a record queue with zi_fast_batch — a pure accelerator: it only takes symbols whose codes resolve
 * from the stream's shared-memory tables and that are valid, fit the output and cannot run past the input;
 * at anything else (end of block, a bad code, the last bytes of a buffer) it stops in front of that symbol
 * without consuming it and the generic zi_step, the code the CPU tests pin against the reference, takes
 * over for one step.  The G lanes then write the batch together (zw_emit).  Groups of one warp synchronise
 * with their own lane masks only, so they progress independently.  G = 32 is one warp per stream. */
#define ZW_THREADS 128

struct ZwLut { uint32_t len[32]; uint32_t dist[32]; };     /* base | extra bits << 16 (RFC 1951 3.2.5) */

template <int G> struct ZwStream {
    zi_tables T;
    zi_aux X;
    uint32_t q[G];
};

/* The records of one batch written by the G lanes of a group: output offsets from a prefix sum of the
 * lengths, all literals at once, every match that reads nothing of this batch by its own lane, then the
 * remaining matches in order, each copied by all lanes (distances shorter than the copy repeat their pattern;
 * a copy longer than its distance >= G proceeds in G-byte steps that read what the previous step wrote). */
template <int G>
__device__ __forceinline__ void zw_emit(uint8_t *out, uint32_t base, const uint32_t *q, uint32_t n, uint32_t gl, uint32_t gmask, uint32_t gshift)
{
    const uint32_t r = gl < n ? q[gl] : 0u;
    const bool is_match = gl < n && (r >> 31);
    const uint32_t olen = gl < n ? (is_match ? ((r >> 16) & 0xFFu) + 3u : 1u) : 0u;
    uint32_t inc = olen;
#pragma unroll
    for (int o = 1; o < G; o <<= 1) { const uint32_t t = __shfl_up_sync(gmask, inc, o, G); if ((int)gl >= o) inc += t; }
    const uint32_t pos = inc - olen;
    if (gl < n && !is_match) out[base + pos] = (uint8_t)r;
    /* matches whose source lies entirely before this batch's output depend on nothing written here: every
       lane copies its own, all at once (most matches of a batch; their lengths are short) */
    const uint32_t dist = (r & 0x7FFFu) + 1u;
    const bool own = is_match && dist >= pos + olen && olen <= 32u;
    if (own) {
        uint8_t *dst = out + base + pos;
        const uint8_t *src = dst - dist;
        /* the source ends before this batch begins, so no load depends on a store of the copy: the three bytes every
           match has go in one round trip, the rest four at a time (byte after byte the loop was one L2 round trip
           per byte and held a quarter of the kernel's stall samples) */
        {
            const uint8_t b0 = src[0], b1 = src[1], b2 = src[2];
            dst[0] = b0; dst[1] = b1; dst[2] = b2;
        }
        uint32_t k = 3;
        for (; k + 4 <= olen; k += 4) {
            const uint8_t b0 = src[k], b1 = src[k + 1], b2 = src[k + 2], b3 = src[k + 3];
            dst[k] = b0; dst[k + 1] = b1; dst[k + 2] = b2; dst[k + 3] = b3;
        }
        for (; k < olen; k++) dst[k] = src[k];
    }
    uint32_t mm = __ballot_sync(gmask, is_match && !own) >> gshift;
    __syncwarp(gmask);
    while (mm) {
        const int j = __ffs((int)mm) - 1;
        mm &= mm - 1;
        const uint32_t p = base + __shfl_sync(gmask, pos, j, G);
        const uint32_t L = __shfl_sync(gmask, olen, j, G);
        const uint32_t D = (__shfl_sync(gmask, r, j, G) & 0x7FFFu) + 1u;
        uint8_t *dst = out + p;
        const uint8_t *src = dst - D;
        if (D >= L) {
            for (uint32_t k = gl; k < L; k += G) dst[k] = src[k];
        } else if (D >= (uint32_t)G) {
            for (uint32_t k0 = 0; k0 < L; k0 += G) {
                const uint32_t k = k0 + gl;
                if (k < L) dst[k] = src[k];
                __syncwarp(gmask);
            }
        } else {
            uint32_t k = gl, km = gl % D;
            const uint32_t step = (uint32_t)G % D;
            for (; k < L; k += G) { dst[k] = src[km]; km += step; if (km >= D) km -= D; }
        }
        __syncwarp(gmask);
    }
}

template <int G>
__global__ void __launch_bounds__(ZW_THREADS, 8)
zs_inflate_group_kernel(uint32_t n, const ZsStream *__restrict__ streams, const uint8_t *__restrict__ comp,
                        uint8_t *__restrict__ raw, int32_t wrap, int32_t *__restrict__ ret,
                        uint32_t *__restrict__ produced, uint32_t *__restrict__ consumed,
                        uint32_t *__restrict__ aux /* [2n]: stored check, flags */)
{
    constexpr int GROUPS = ZW_THREADS / G;
    extern __shared__ __align__(16) unsigned char zw_smem_raw[];
    ZwStream<G> *W = reinterpret_cast<ZwStream<G> *>(zw_smem_raw);
    ZwLut &lut = *reinterpret_cast<ZwLut *>(zw_smem_raw + sizeof(ZwStream<G>) * GROUPS);
    if (threadIdx.x < 29) {
        const uint32_t c = threadIdx.x;
        lut.len[c] = zi_lut_len(c);
    }
    if (threadIdx.x >= 32 && threadIdx.x < 62) {
        const uint32_t d = threadIdx.x - 32;
        lut.dist[d] = zi_lut_dist(d);
    }
    __syncthreads();
    const uint32_t g = threadIdx.x / G, gl = threadIdx.x % G;
    const uint32_t gshift = (threadIdx.x & 31u) - gl;                      /* first lane of the group within its warp */
    const uint32_t gmask = (G == 32 ? 0xFFFFFFFFu : ((1u << (G & 31)) - 1u) << gshift);
    const uint32_t s = blockIdx.x * GROUPS + g;
    if (s >= n) return;
    ZwStream<G> &w = W[g];
    const ZsStream st = streams[s];
    const uint8_t *in = comp + st.comp_off;
    uint8_t *out = raw + st.raw_off;
    /* per-stream options of section-parallel decoding ride in ZsStream.chunk_first (unused by inflate):
       ZI_OPT_* | 4 = this stream continues another one: no zlib header in front of its first block (the trailer,
       if the batch's wrap has one, still follows its final block) */
    const uint32_t sopt = st.chunk_first;
    zi_mach m;
    zi_m_init(&m, in, st.comp_cap, out, st.raw_len, wrap, &w.T, &w.X);
    m.opts = sopt & 3u;
    if ((sopt & 4u) && m.state == ZM_HEAD) m.state = ZM_BLOCK;
    const bool count_only = (sopt & ZI_OPT_COUNT_ONLY) != 0;
    for (;;) {
        const int state = __shfl_sync(gmask, m.state, 0, G);
        if (state == ZM_DONE) break;
        if (state == ZM_SYM) {
            uint32_t cnt = 0, base = 0, vop = 0;
            if (gl == 0) { base = m.io.op; cnt = zi_fast_batch(&m, lut.len, lut.dist, w.q, G, &vop); }
            cnt = __shfl_sync(gmask, cnt, 0, G);
            base = __shfl_sync(gmask, base, 0, G);
            if (cnt) {
                __syncwarp(gmask);                           /* the leader's queue writes are visible to the group */
                if (!count_only) zw_emit<G>(out, base, w.q, cnt, gl, gmask, gshift);
                if (gl == 0) m.io.op = vop;
            }
            if (cnt < (uint32_t)G) {
                /* the fast decoder stopped in front of something: one generic step */
                if (gl == 0) zi_step(&m);
                __syncwarp(gmask);
            }
        } else if (state == ZM_STORED) {
            uint32_t cnt = 0, from = 0, to = 0;
            if (gl == 0) { cnt = zi_stored_plan(&m); from = m.io.ip; to = m.io.op; }
            cnt = __shfl_sync(gmask, cnt, 0, G);
            from = __shfl_sync(gmask, from, 0, G);
            to = __shfl_sync(gmask, to, 0, G);
            if (!count_only) for (uint32_t k = gl; k < cnt; k += G) out[to + k] = in[from + k];
            __syncwarp(gmask);
            if (gl == 0) zi_stored_done(&m, cnt);
        } else {
            if (gl == 0) zi_step(&m);
            __syncwarp(gmask);
        }
    }
    if (gl == 0) {
        ret[s] = m.res.ret;
        produced[s] = m.res.produced;
        consumed[s] = m.res.consumed;
        aux[2 * s] = m.res.stored_check;
        aux[2 * s + 1] = m.res.have_check | (m.res.data_errors ? 2u : 0u) | (m.res.at_flush ? 4u : 0u);
    }
}

template <int G>
static cudaError_t zs_inflate_group_launch(cudaStream_t st, uint32_t n, const ZsStream *streams, const uint8_t *comp, uint8_t *raw, int32_t wrap,
                                           int32_t *ret, uint32_t *produced, uint32_t *consumed, uint32_t *aux)
{
    constexpr int GROUPS = ZW_THREADS / G;
    const size_t smem = sizeof(ZwStream<G>) * GROUPS + sizeof(ZwLut);
    cudaFuncSetAttribute(zs_inflate_group_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(zs_inflate_group_kernel<G>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    zs_inflate_group_kernel<G><<<(n + GROUPS - 1) / GROUPS, ZW_THREADS, smem, st>>>(n, streams, comp, raw, wrap, ret, produced, consumed, aux);
    return cudaGetLastError();
}

__global__ void zs_inflate_check_kernel(uint32_t n, const ZsAdlerAcc *__restrict__ acc, const uint32_t *__restrict__ produced,
                                        const uint32_t *__restrict__ aux, int32_t wrap, int32_t *__restrict__ ret,
                                        uint32_t *__restrict__ check)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t a = (uint32_t)((acc[s].s1 + 1) % ZS_ADLER_BASE);
    uint32_t b = (uint32_t)((acc[s].s2 + produced[s]) % ZS_ADLER_BASE);
    uint32_t v = (b << 16) | a;
    check[s] = v;
    if ((wrap & 0xFF) == 1 && ret[s] == 0 && (aux[2 * s + 1] & 1u) && aux[2 * s] != v) ret[s] = -3;   /* incorrect data check */
}

extern "C" cudaError_t zs_adler_streams_launch(cudaStream_t st, uint32_t n, uint32_t max_len, const uint8_t *raw,
                                               const ZsStream *streams, const uint32_t *produced, ZsAdlerAcc *acc);

extern "C" cudaError_t zs_inflate_launch(cudaStream_t st, uint32_t n, const ZsStream *streams, const uint8_t *comp,
                                         uint8_t *raw, int32_t wrap, int32_t *ret, uint32_t *produced,
                                         uint32_t *consumed, uint32_t *check, uint32_t *aux, ZsAdlerAcc *acc,
                                         uint32_t max_raw_len, int with_check, uint32_t *counter, int sms)
{
    if (n == 0) return cudaSuccess;
    (void)counter; (void)sms;
    /* a warp per stream while that fills the machine (148 SMs x 32 warps), two streams per warp beyond: both
       leaders of a warp decode at the same time, the decode cost per symbol halves.  Smaller groups measured
       slower (8 lanes: 40 GB/s, 4 lanes: 25 GB/s against 54 GB/s at 65 536 streams): shared memory holds 64 streams per
       SM whatever the group size, so fewer lanes per stream only mean fewer warps to hide latency with.  Also measured
       and dropped (round 2, profiles/r02_exp_inflate_lanes_*.log): a lane per stream on a few decoder warps in lockstep
       feeding writer warps through double-buffered queues — 4.4 instead of 15 warp-instructions per output byte, but
       33-41 GB/s at 16 384 streams and 32 GB/s at 65 536: with 64 streams per SM there are too few decoder warps, and
       every lockstep step pays the literal, the match and the refill path one after the other (1900 cycles a symbol). */
    cudaError_t ge = n <= ZS_INFLATE_WARP_MAX ? zs_inflate_group_launch<32>(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux)
                                              : zs_inflate_group_launch<16>(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux);
    if (ge != cudaSuccess) return ge;
    if (!with_check) return cudaSuccess;               /* section passes: the caller checks the whole stream */
    cudaMemsetAsync(acc, 0, sizeof(ZsAdlerAcc) * n, st);
    cudaError_t ce = zs_adler_streams_launch(st, n, max_raw_len, raw, streams, produced, acc);
    if (ce != cudaSuccess) return ce;
    zs_inflate_check_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, acc, produced, aux, wrap, ret, check);
    return cudaGetLastError();
}

