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
#include <stdint.h>
#include "common.cuh"

#define ZW_THREADS 128
struct ZwLut { uint32_t len[32]; uint32_t dist[32]; };     /* base | extra bits << 16 (RFC 1951 3.2.5) */

/* The decoder core and the group kernel exist twice: `zw` with the compact tables (9 / 6 root bits, 1.6 KB per stream:
 * 128 streams share an SM, for wide batches) and `zn` with wide roots (10 / 8 bits: fewer second-level lookups, for
 * batches of at most a warp per stream, where latency decides).  Same code, same results. */
namespace zw {
#include "inflate_core.h"
#include "inflate_group.inc"
}
#undef ZI_LBITS
#undef ZI_DBITS
#undef ZI_POOL
#define ZI_LBITS 10
#define ZI_DBITS 8
#define ZI_POOL 256
#define ZI_REINCLUDE
namespace zn {
#include "inflate_core.h"
#include "inflate_group.inc"
#include "inflate_spec.h"
#include "inflate_spec.inc"
}
/* The speculative warp decoder (inflate_spec.h) exists twice as well: `zn` keeps each stream's window and the symbols of a
 * round in shared memory (55 KB per stream, 4 per SM, 75 KB and 3 or 2 per SM with a second warp: few streams, latency
 * decides); `zm` keeps only tables and bitmap there (5 KB), the symbols in a scratch in global memory (17 KB per stream,
 * written once and read once per round: L2 traffic) and writes straight to global memory: 28 streams per SM — the batch fills
 * the machine.  Same code, same results.  (Measured: symbols in shared memory, 15 KB per stream and 14 per SM: 55 instead of
 * 73 GB/s at 4096 streams; regions of 320 / 384 / 640 / 768 bits, 24 / 32 streams per SM: profiles/README.md.) */
#ifndef ZP_WIDE_R
#define ZP_WIDE_R 512u
#define ZP_WIDE_CAP 128u
#endif
#undef ZP_R
#undef ZP_CAP
#define ZP_R ZP_WIDE_R
#define ZP_CAP ZP_WIDE_CAP
#define ZP_NO_PROFILE
#define ZP_REC_GLOBAL
namespace zm {
#include "inflate_core.h"
#include "inflate_group.inc"
#include "inflate_spec.h"
#include "inflate_spec.inc"
}
#undef ZP_NO_PROFILE
#undef ZP_REC_GLOBAL
#undef ZI_REINCLUDE
using namespace zw;                                          /* the streaming kernel below uses the compact geometry */

#define ZSI_HIST 32768u                           /* history kept in front of a streaming slot's output staging */

#ifndef ZS_INFLATE_WARP_MAX
#define ZS_INFLATE_WARP_MAX 6144u                /* streams in a batch up to which each gets a whole warp */
#define ZS_INFLATE_G16_MAX 12288u
#define ZS_INFLATE_RING_MAX 640u                 /* streams in a batch up to which the speculative warp decoder keeps each stream's window in shared memory (4 x 148 fit at once) */
#define ZS_INFLATE_SPEC_MAX 12288u               /* ... and up to which it is used at all */                /* ... and up to which each gets half a warp; beyond, a quarter */
#endif

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

/* ======================= one z_stream, resumable (the streaming API: inflate / inflateSync) =======================
 * The decoder state of a stream lives in a device slot between calls (machine, tables, up to ZSI_IN staged input
 * bytes, 32 KiB of history in front of an output staging buffer).  A call runs the same state machine as everything
 * else (zi_step) on one lane until the staged input runs out, the output staging is full, the stream ends, or an
 * error stops it.  zi_step assumes all input and output present, so every step is taken on a copy of the machine and
 * thrown away when it ends in "input ended" / "output full": a step never consumes half a symbol or half a header,
 * which makes the machine resumable without touching the code the known-answer tests pin.  A data error is reported
 * and the machine waits in ZM_RECOVER for inflateSync (the one-shot path recovers by itself; the z_stream API leaves
 * that to the caller: reference src/inflate.c:1356-1360, :1547). */
struct ZsInfSlot {
    zi_mach m;
    zi_tables T;
    zi_aux X;
};
struct ZsInfStepArgs {
    ZsInfSlot *slot;
    uint8_t *in;            /* staging: in_have bytes */
    uint8_t *out;           /* ZSI_HIST bytes of history, then the output staging */
    uint32_t in_have, out_cap;
    int32_t mode;           /* 0 inflate, 1 inflateSync (look for a flush point), 2 reset (wrap in out_cap), 3 dictionary of out_cap bytes installed */
    uint32_t *res;          /* [8]: status, input position reached, produced, adler32 of the produced bytes, stored check / DICTID, have_check, new history length */
};
enum { ZSI_NEED_INPUT = 0, ZSI_OUTPUT_FULL = 1, ZSI_END = 2, ZSI_DATA_ERROR = 3, ZSI_NEED_DICT = 4, ZSI_SYNC_FOUND = 5 };

__device__ __forceinline__ void zsi_move_down(uint8_t *dst, const uint8_t *src, uint32_t n, uint32_t lane)
{
    /* overlapping move towards lower addresses: every round loads before it stores */
    for (uint32_t k0 = 0; k0 < n; k0 += 32) {
        const uint32_t k = k0 + lane;
        uint8_t b = 0;
        if (k < n) b = src[k];
        __syncwarp();
        if (k < n) dst[k] = b;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32, 1) zs_inflate_stream_kernel(ZsInfStepArgs a)
{
    __shared__ zi_tables T;
    __shared__ zi_aux X;
    __shared__ uint32_t sh[4];
    const uint32_t lane = threadIdx.x;
    ZsInfSlot *slot = a.slot;
    if (a.mode == 2) {
        if (lane == 0) {
            zi_mach m;
            zi_m_init(&m, a.in, 0, a.out + ZSI_HIST, 0, (int)a.out_cap, nullptr, nullptr);
            m.opts = ZI_OPT_STREAMING;
            slot->m = m;
        }
        return;
    }
    if (a.mode == 3) {
        if (lane == 0) {
            slot->m.hist = a.out_cap;                       /* the bytes lie right in front of the output staging */
            if (slot->m.state == ZM_DICT) slot->m.state = ZM_BLOCK;
        }
        return;
    }
    for (uint32_t i = lane; i < sizeof(zi_tables) / 4; i += 32) ((uint32_t *)&T)[i] = ((const uint32_t *)&slot->T)[i];
    for (uint32_t i = lane; i < sizeof(zi_aux) / 4; i += 32) ((uint32_t *)&X)[i] = ((const uint32_t *)&slot->X)[i];
    __syncwarp();
    uint8_t *out = a.out + ZSI_HIST;
    if (lane == 0) {
        zi_mach m = slot->m;
        m.T = &T; m.X = &X;
        m.io.in = a.in; m.io.in_len = a.in_have; m.io.pv = 0;
        m.io.out = out; m.io.out_cap = a.out_cap; m.io.op = 0; m.base = 0;
        if (m.skip && a.in_have) { zi_refill(&m.io); zi_drop(&m.io, (int)m.skip); m.skip = 0; }   /* the first staged byte was partly used */
        uint32_t status = ZSI_NEED_INPUT;
        if (a.mode == 1 && m.state != ZM_DONE) {
            /* inflateSync: search the staged input for 00 00 FF FF (zi_sync, the reference's syncsearch); the bits still
               buffered are dropped first, as the reference's byte-aligned restart does */
            uint32_t pos = m.io.ip - (m.io.bits >> 3);
            if (pos > m.io.in_len) pos = m.io.in_len;
            const uint32_t held = m.state == ZM_RECOVER ? m.held : 0u;
            const uint32_t nx = zi_sync(m.io.in, m.io.in_len, pos >= held ? pos - held : 0u);
            if (nx > m.io.in_len) {
                /* not here: everything but a possible marker prefix is used up */
                zi_seek(&m.io, m.io.in_len > 3 ? m.io.in_len - 3 : 0);
                m.held = 0; m.state = ZM_RECOVER;
                status = ZSI_DATA_ERROR;
            } else {
                zi_seek(&m.io, nx);
                m.held = 0; m.hist = 0; m.win = 32768u; m.state = ZM_BLOCK;
                status = ZSI_SYNC_FOUND;
            }
        } else {
            for (;;) {
                if (m.state == ZM_DONE) { status = (m.res.ret == ZI_OK) ? ZSI_END : ZSI_DATA_ERROR; break; }
                if (m.state == ZM_DICT) { status = ZSI_NEED_DICT; break; }
                if (m.state == ZM_RECOVER) { status = ZSI_DATA_ERROR; break; }
                zi_mach snap = m;
                zi_step(&m);
                if (m.state == ZM_DONE && m.res.ret == ZI_BUF_ERROR) {
                    const int why = m.res.last_reason;
                    m = snap;                                   /* the step did not happen */
                    status = (why == ZI_E_OUTPUT_FULL) ? ZSI_OUTPUT_FULL : ZSI_NEED_INPUT;
                    break;
                }
            }
        }
        /* hand back the whole bytes still in the bit buffer; a partial byte stays in it */
        uint32_t pos = m.io.ip - (m.io.bits >> 3);
        if (pos > m.io.in_len) pos = m.io.in_len;
        if (m.state == ZM_RECOVER && status == ZSI_DATA_ERROR && a.mode == 0) {
            const uint32_t held = m.held < pos ? m.held : pos;  /* inflateSync starts its search at the bytes the reference still holds */
            pos -= held; m.held = held;
            m.io.hold = 0; m.io.bits = 0;
        } else {
            /* a partly used byte stays staged (the cursor never runs ahead of the input); its used bits are skipped next time */
            const uint32_t keep = m.io.bits & 7u;
            if (keep && pos > 0) { pos -= 1; m.skip = 8u - keep; }
            m.io.hold = 0; m.io.bits = 0;
        }
        m.io.ip = 0; m.io.pv = 0;
        const uint32_t valid = m.io.op - m.base + m.hist;
        const uint32_t nh = valid < ZSI_HIST ? valid : ZSI_HIST;
        sh[0] = pos; sh[1] = m.io.op; sh[2] = nh; sh[3] = status;
        a.res[0] = status; a.res[1] = pos; a.res[2] = m.io.op;
        a.res[4] = m.res.stored_check; a.res[5] = m.res.have_check; a.res[6] = nh; a.res[7] = (uint32_t)m.res.ret;
        m.hist = nh; m.base = 0;
        if (m.state == ZM_RECOVER && a.mode == 0) m.io.ip = m.held;      /* the search of inflateSync starts `held` bytes into the leftover... */
        slot->m = m;
    }
    __syncwarp();
    const uint32_t pos = sh[0], produced = sh[1], nh = sh[2];
    /* adler32 of the produced bytes (as a stream of its own; the host folds it into the running value) */
    {
        unsigned long long s1 = 0, s2 = 0;
        for (uint32_t i = lane; i < produced; i += 32) { const uint32_t b = out[i]; s1 += b; s2 += (unsigned long long)(produced - i) * b; }
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_down_sync(0xFFFFFFFFu, s1, o); s2 += __shfl_down_sync(0xFFFFFFFFu, s2, o); }
        if (lane == 0) {
            const uint32_t A = (uint32_t)((s1 + 1) % ZS_ADLER_BASE), B = (uint32_t)((s2 + produced) % ZS_ADLER_BASE);
            a.res[3] = (B << 16) | A;
        }
    }
    /* tables back to the slot; unread input to the front of the staging; the newest history in front of the output staging */
    for (uint32_t i = lane; i < sizeof(zi_tables) / 4; i += 32) ((uint32_t *)&slot->T)[i] = ((const uint32_t *)&T)[i];
    for (uint32_t i = lane; i < sizeof(zi_aux) / 4; i += 32) ((uint32_t *)&slot->X)[i] = ((const uint32_t *)&X)[i];
    if (pos) zsi_move_down(a.in, a.in + pos, a.in_have - pos, lane);
    if (produced) zsi_move_down(out - nh, out + produced - nh, nh, lane);
}

extern "C" cudaError_t zs_inflate_stream_launch(cudaStream_t st, void *slot, uint8_t *in, uint8_t *out, uint32_t in_have, uint32_t out_cap, int32_t mode, uint32_t *res)
{
    ZsInfStepArgs a;
    a.slot = (ZsInfSlot *)slot; a.in = in; a.out = out; a.in_have = in_have; a.out_cap = out_cap; a.mode = mode; a.res = res;
    zs_inflate_stream_kernel<<<1, 32, 0, st>>>(a);
    return cudaGetLastError();
}
extern "C" size_t zs_inflate_stream_slot_bytes(void) { return sizeof(ZsInfSlot); }

extern "C" size_t zs_inflate_spec_scratch_bytes(void) { return (size_t)ZP_WIDE_CAP * 33u * 4u; }   /* per stream, of the wide build */

/* zscgpu_init: the inflate kernels are large; loaded here, not inside the first call (module loading is lazy) */
extern "C" cudaError_t zs_inflate_preload(void)
{
    cudaFuncAttributes fa;
    cudaError_t ce;
    if ((ce = cudaFuncGetAttributes(&fa, zn::zs_inflate_pipe_kernel<2>)) != cudaSuccess) return ce;
    if ((ce = cudaFuncGetAttributes(&fa, zn::zs_inflate_pipe_kernel<3>)) != cudaSuccess) return ce;
    if ((ce = cudaFuncGetAttributes(&fa, zn::zs_inflate_spec_kernel<true>)) != cudaSuccess) return ce;
    if ((ce = cudaFuncGetAttributes(&fa, zm::zs_inflate_spec_kernel<false>)) != cudaSuccess) return ce;
    if ((ce = cudaFuncGetAttributes(&fa, zw::zs_inflate_group_kernel<8>)) != cudaSuccess) return ce;
    if ((ce = cudaFuncGetAttributes(&fa, zs_inflate_stream_kernel)) != cudaSuccess) return ce;
    return cudaFuncGetAttributes(&fa, zs_inflate_check_kernel);
}

extern "C" cudaError_t zs_adler_streams_launch(cudaStream_t st, uint32_t n, uint32_t max_len, const uint8_t *raw,
                                               const ZsStream *streams, const uint32_t *produced, ZsAdlerAcc *acc);

extern "C" cudaError_t zs_inflate_launch(cudaStream_t st, uint32_t n, const ZsStream *streams, const uint8_t *comp,
                                         uint8_t *raw, int32_t wrap, int32_t *ret, uint32_t *produced,
                                         uint32_t *consumed, uint32_t *check, uint32_t *aux, ZsAdlerAcc *acc,
                                         uint32_t max_raw_len, int with_check, uint32_t *counter /* zi_aux[n] */, int sms /* < 0: one wave of several */,
                                         uint32_t *spec_scratch /* [n] x zs_inflate_spec_scratch_bytes() */)
{
    if (n == 0) return cudaSuccess;
    const int wide_hint = sms < 0;                     /* the batch is one wave of a larger job: the machine is shared, no window in shared memory */
    static_assert(sizeof(zw::zi_aux) == 640 && sizeof(zn::zi_aux) == 640, "engine.cu sizes the zi_aux pool with 640 bytes per stream");
    /* a warp per stream while that fills the machine (148 SMs x 32 warps); beyond, several streams per warp: their
       leaders decode at the same time, and the two-level tables (1.6 KB per stream) let 128 streams share an SM */
    uint32_t g = n <= ZS_INFLATE_WARP_MAX ? 32u : (n <= ZS_INFLATE_G16_MAX ? 16u : 8u);
#ifdef ZSC_TUNING
    if (getenv("ZSC_B200_INFLATE_G")) g = (uint32_t)atoi(getenv("ZSC_B200_INFLATE_G"));
#endif
    /* A warp per stream with all of its lanes decoding (inflate_spec.inc).  4 / 3: two warps per stream, one decoding ahead of
       the one that writes, while two / three streams per SM hold the batch (or hold it in two turns); 1: one warp, the stream's
       window in shared memory, while four per SM hold it; 2: the wide build, 28 streams per SM.  (0: the group kernels, one
       decoding lane per group of 8 / 16 / 32: without a scratch, and in tuning builds — the wide build passed them at every
       batch size.) */
    const uint32_t nsm = sms > 0 ? (uint32_t)sms : 0u;
    int spec = 2;
    if (!wide_hint && nsm) spec = n <= 2u * nsm ? 4 : n <= 3u * nsm ? 3 : n <= 4u * nsm ? 1 : n <= 6u * nsm ? 3 : 2;
    if (!spec_scratch && spec == 2) spec = 0;
#ifdef ZSC_TUNING
    if (getenv("ZSC_B200_INFLATE_SPEC")) spec = atoi(getenv("ZSC_B200_INFLATE_SPEC"));
#endif
    cudaError_t ge = spec == 4 ? zn::zs_inflate_pipe_launch<3>(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux, reinterpret_cast<zn::zi_aux *>(counter))
                   : spec == 3 ? zn::zs_inflate_pipe_launch<2>(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux, reinterpret_cast<zn::zi_aux *>(counter))
                   : spec == 1 ? zn::zs_inflate_spec_launch<true>(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux, reinterpret_cast<zn::zi_aux *>(counter))
                   : spec == 2 ? zm::zs_inflate_spec_launch<false>(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux, reinterpret_cast<zm::zi_aux *>(counter), spec_scratch)
                   : g == 32 ? zn::zs_inflate_group_launch<32>(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux, reinterpret_cast<zn::zi_aux *>(counter))
                   : g == 16 ? zn::zs_inflate_group_launch<16>(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux, reinterpret_cast<zn::zi_aux *>(counter))
                             : zw::zs_inflate_group_launch<8>(st, n, streams, comp, raw, wrap, ret, produced, consumed, aux, reinterpret_cast<zw::zi_aux *>(counter));
    if (ge != cudaSuccess) return ge;
    if (!with_check) return cudaSuccess;               /* section passes: the caller checks the whole stream */
    cudaMemsetAsync(acc, 0, sizeof(ZsAdlerAcc) * n, st);
    cudaError_t ce = zs_adler_streams_launch(st, n, max_raw_len, raw, streams, produced, acc);
    if (ce != cudaSuccess) return ce;
    zs_inflate_check_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, acc, produced, aux, wrap, ret, check);
    return cudaGetLastError();
}

