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

/* ======================= a lane per stream decodes, writer warps write =======================
 * Wide batches (thousands of streams).  Huffman decoding never looks at the output, so it can run ahead of the
 * writing: the 64 lanes of two decoder warps each own one stream and decode, in lockstep, bursts of up to ZX_Q symbols
 * into that stream's record queue (zi_fast_batch — the same accelerator the group kernel uses, but on 32 lanes of a
 * warp at once instead of one); the other warps of the CTA are writers: while the decoders fill one half of the
 * double-buffered queues, each writer warp empties the other half of the streams it owns (zw_emit<32>: offsets by
 * prefix sum, literals at once, matches coalesced).  A stream's batches are always written by the same warp, in order,
 * so a match can read what an earlier batch wrote.  Everything the accelerator does not take (block ends, the last
 * bytes of a buffer, errors, recovery) is done by the generic zi_step on the decoder lane itself, as in every other
 * form of the decoder — it writes its few bytes directly, and therefore only runs once the stream's previous batch
 * has been written (the lane sits out one phase).  Stored blocks are handed to the writers as one copy.
 * Finished lanes take the next stream of the batch from a global counter, so ragged batches stay balanced.
 * Results are those of the one-thread decoder by construction (zi_fast_batch and zi_step are the code tests/ pins). */
#define ZX_DEC_WARPS 2
#define ZX_STREAMS (ZX_DEC_WARPS * 32)
#define ZX_WR_WARPS 14
#define ZX_THREADS ((ZX_DEC_WARPS + ZX_WR_WARPS) * 32)
#define ZX_Q 32

struct ZxStream {
    zi_tables T;
    zi_aux X;
    uint32_t q[2][ZX_Q];
};
struct ZxCtl {                        /* what a decoder lane hands to the writer of its stream, per phase parity */
    uint8_t *out[2];                  /* the stream's output buffer (nullptr: a count-only stream, nothing is written) */
    const uint8_t *in[2];             /* ... and input (stored blocks are copied from it) */
    uint32_t cnt[2], base[2];         /* records in q[par] and the output position of the first */
    uint32_t st_n[2], st_from[2], st_to[2];   /* a stored-block copy that precedes the records */
};

__global__ void __launch_bounds__(ZX_THREADS, 1)
zs_inflate_lanes_kernel(uint32_t n, const ZsStream *__restrict__ streams, const uint8_t *__restrict__ comp,
                        uint8_t *__restrict__ raw, int32_t wrap, int32_t *__restrict__ ret,
                        uint32_t *__restrict__ produced, uint32_t *__restrict__ consumed,
                        uint32_t *__restrict__ aux, uint32_t *__restrict__ next_stream)
{
    extern __shared__ __align__(16) unsigned char zx_smem_raw[];
    ZxStream *W = reinterpret_cast<ZxStream *>(zx_smem_raw);
    ZxCtl *ctl = reinterpret_cast<ZxCtl *>(zx_smem_raw + sizeof(ZxStream) * ZX_STREAMS);
    ZwLut &lut = *reinterpret_cast<ZwLut *>(zx_smem_raw + sizeof(ZxStream) * ZX_STREAMS + sizeof(ZxCtl) * ZX_STREAMS);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool decoder = warp < ZX_DEC_WARPS;
    if (tid < 29) lut.len[tid] = zi_lut_len(tid);
    if (tid >= 32 && tid < 62) lut.dist[tid - 32] = zi_lut_dist(tid - 32);
    if (tid < ZX_STREAMS) { ZxCtl &c = ctl[tid]; c.cnt[0] = c.cnt[1] = 0; c.st_n[0] = c.st_n[1] = 0; }
    __syncthreads();

    zi_mach m;
    bool have = false, more = true;   /* this lane owns a stream / the batch may still have unclaimed streams */
    uint32_t sid = 0;
    uint32_t last_emit = 0xFFFFFFF0u; /* phase in which this lane last handed a batch to its writer */
    uint8_t *s_out = nullptr;        /* output and input of the lane's stream */
    const uint8_t *s_in = nullptr;
    m.state = ZM_DONE;

    for (uint32_t phase = 0;; phase++) {
        const uint32_t par = phase & 1u;
        int alive = 0;
        if (decoder) {
            ZxStream &w = W[tid];
            ZxCtl &c = ctl[tid];
            bool emitted = false;
            c.cnt[par] = 0; c.st_n[par] = 0;          /* this half was drained during the previous phase */
            /* a bounded number of machine steps per phase; a lane stops early once it has handed over a batch */
            for (int it = 0; it < 8; it++) {
                if (!have) {
                    if (!more) break;
                    sid = atomicAdd(next_stream, 1u);
                    if (sid >= n) { more = false; break; }
                    const ZsStream st = streams[sid];
                    zi_m_init(&m, comp + st.comp_off, st.comp_cap, raw + st.raw_off, st.raw_len, wrap, &w.T, &w.X);
                    const uint32_t sopt = st.chunk_first;              /* section options, see zs_inflate_group_kernel */
                    m.opts = sopt & 3u;
                    if ((sopt & 4u) && m.state == ZM_HEAD) m.state = ZM_BLOCK;
                    s_out = (sopt & ZI_OPT_COUNT_ONLY) ? nullptr : raw + st.raw_off; s_in = comp + st.comp_off;
                    have = true;
                }
                const int state = m.state;
                if (state == ZM_DONE) {
                    ret[sid] = m.res.ret; produced[sid] = m.res.produced; consumed[sid] = m.res.consumed;
                    aux[2 * sid] = m.res.stored_check;
                    aux[2 * sid + 1] = m.res.have_check | (m.res.data_errors ? 2u : 0u) | (m.res.at_flush ? 4u : 0u);
                    have = false;
                    if (emitted) break;                /* the next stream would reuse this half of the queue */
                    continue;
                }
                if (state == ZM_SYM) {
                    if (emitted) break;
                    uint32_t vop = 0;
                    const uint32_t base = m.io.op;
                    const uint32_t cnt = zi_fast_batch(&m, lut.len, lut.dist, w.q[par], ZX_Q, &vop);
                    if (cnt) {
                        m.io.op = vop;
                        c.cnt[par] = cnt; c.base[par] = base; c.out[par] = s_out; c.in[par] = s_in;
                        emitted = true; last_emit = phase;
                        if (cnt == ZX_Q) break;
                    }
                    if (cnt < ZX_Q) {
                        /* the accelerator stopped in front of something: one generic step, which writes by itself and so
                           needs every earlier batch of this stream written (handed over in phase p: written by the end of p + 1) */
                        if (emitted || last_emit + 1 == phase) break;
                        zi_step(&m);
                    }
                    continue;
                }
                if (state == ZM_STORED) {
                    if (emitted) break;
                    const uint32_t k = zi_stored_plan(&m);
                    c.st_n[par] = k; c.st_from[par] = m.io.ip; c.st_to[par] = m.io.op;
                    zi_stored_done(&m, k);
                    if (k) { c.out[par] = s_out; c.in[par] = s_in; emitted = true; last_emit = phase; }
                    continue;
                }
                if (state == ZM_COPY) {
                    if (emitted || last_emit + 1 == phase) break;      /* writes by itself, see above */
                    zi_step(&m);
                    continue;
                }
                zi_step(&m);                           /* header, block header, trailer, recovery: no output */
            }
            alive = (have || emitted || more) ? 1 : 0;
        } else if (phase > 0) {
            /* writers: the half the decoders filled in the previous phase */
            const uint32_t pp = par ^ 1u;
            for (uint32_t s = warp - ZX_DEC_WARPS; s < ZX_STREAMS; s += ZX_WR_WARPS) {
                const ZxCtl &c = ctl[s];
                const uint32_t sn = c.st_n[pp], cnt = c.cnt[pp];
                if (!(sn | cnt)) continue;
                uint8_t *out = c.out[pp];
                if (out == nullptr) continue;          /* count-only streams write nothing */
                if (sn) {
                    const uint8_t *src = c.in[pp] + c.st_from[pp];
                    uint8_t *dst = out + c.st_to[pp];
                    for (uint32_t k = lane; k < sn; k += 32) dst[k] = src[k];
                    __syncwarp();
                }
                if (cnt) zw_emit<32>(out, c.base[pp], W[s].q[pp], cnt, lane, 0xFFFFFFFFu, 0u);
            }
        }
        if (!__syncthreads_or(alive)) break;
    }
}

static cudaError_t zs_inflate_lanes_launch(cudaStream_t st, uint32_t n, const ZsStream *streams, const uint8_t *comp, uint8_t *raw, int32_t wrap,
                                           int32_t *ret, uint32_t *produced, uint32_t *consumed, uint32_t *aux, uint32_t *next_stream, int sms)
{
    const size_t smem = sizeof(ZxStream) * ZX_STREAMS + sizeof(ZxCtl) * ZX_STREAMS + sizeof(ZwLut);
    cudaFuncSetAttribute(zs_inflate_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemsetAsync(next_stream, 0, sizeof(uint32_t), st);
    uint32_t grid = (n + ZX_STREAMS - 1) / ZX_STREAMS;
    if (grid > (uint32_t)sms) grid = (uint32_t)sms;
    zs_inflate_lanes_kernel<<<grid, ZX_THREADS, smem, st>>>(n, streams, comp, raw, wrap, ret, produced, consumed, aux, next_stream);
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

#ifndef ZS_INFLATE_LANES_MIN
#define ZS_INFLATE_LANES_MIN 6145u               /* streams in a batch from which the lane-per-stream kernel takes over */
#endif
extern "C" cudaError_t zs_inflate_launch(cudaStream_t st, uint32_t n, const ZsStream *streams, const uint8_t *comp,
                                         uint8_t *raw, int32_t wrap, int32_t *ret, uint32_t *produced,
                                         uint32_t *consumed, uint32_t *check, uint32_t *aux, ZsAdlerAcc *acc,
                                         uint32_t max_raw_len, int with_check, uint32_t *counter, int sms)
{
    if (n == 0) return cudaSuccess;
    uint32_t lanes_min = ZS_INFLATE_LANES_MIN;
#ifdef ZSC_TUNING
    if (getenv("ZSC_B200_INFLATE_LANES_MIN")) lanes_min = (uint32_t)atoi(getenv("ZSC_B200_INFLATE_LANES_MIN"));
#endif
    /* a warp per stream while that fills the machine (148 SMs x 32 warps), two streams per warp beyond: both
       leaders of a warp decode at the same time, the decode cost per symbol halves.  Smaller groups measured
       slower (8 lanes: 40 GB/s, 4 lanes: 25 GB/s against 54 GB/s at 65 536 streams): shared memory holds 64 streams per
       SM whatever the group size, so fewer lanes per stream only mean fewer warps to hide latency with. */
    cudaError_t ge = n >= lanes_min ? zs_inflate_lanes_launch(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux, counter, sms)
                     : n <= ZS_INFLATE_WARP_MAX ? zs_inflate_group_launch<32>(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux)
                                              : zs_inflate_group_launch<16>(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux);
    if (ge != cudaSuccess) return ge;
    if (!with_check) return cudaSuccess;               /* section passes: the caller checks the whole stream */
    cudaMemsetAsync(acc, 0, sizeof(ZsAdlerAcc) * n, st);
    cudaError_t ce = zs_adler_streams_launch(st, n, max_raw_len, raw, streams, produced, acc);
    if (ce != cudaSuccess) return ce;
    zs_inflate_check_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, acc, produced, aux, wrap, ret, check);
    return cudaGetLastError();
}

