/* lz_common.cuh — device helpers shared by the two LZ77 kernels (deflate_lz.cu: one candidate per position;
 * deflate_chain.cu: hash chains): ring-buffer loads, the 3-byte hash, match-length comparisons, and the bulk-copy
 * (cp.async.bulk + mbarrier) staging of input bytes into the shared-memory ring. */
#ifndef ZSC_LZ_COMMON_CUH
#define ZSC_LZ_COMMON_CUH
#include "common.cuh"

#define ZL_WORKER_WARPS 16
#define ZL_WORKERS (ZL_WORKER_WARPS * 32)          /* 512 worker threads */
#define ZL_THREADS (ZL_WORKERS + 32)               /* + the hasher warp */
#define ZL_MIRROR 32u                               /* bytes of the ring start repeated after its end */
#define ZL_LOOKAHEAD 272u                          /* >= 258 + 3, multiple of 16 */
#define ZL_NONE 0xFFFFu

__device__ __forceinline__ void zl_bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(ZL_WORKERS) : "memory"); }

template <uint32_t RING> __device__ __forceinline__ uint32_t zl_ld32(const uint32_t *ring32, uint32_t q)
{
    uint32_t i = (q >> 2) & (RING / 4 - 1);
    uint32_t w0 = ring32[i], w1 = ring32[i + 1];
    return __funnelshift_r(w0, w1, (q & 3) * 8);
}
template <uint32_t RING> __device__ __forceinline__ uint32_t zl_ld8(const uint32_t *ring32, uint32_t q)
{
    return ((const uint8_t *)ring32)[q & (RING - 1)];
}
template <uint32_t HASH_BITS> __device__ __forceinline__ uint32_t zl_hash(uint32_t v)
{
    uint32_t h = ((v & 0xFFFFFFu) * 2654435761u) >> (32 - HASH_BITS);
    return h == 0x7FFFu ? 0x7FFEu : h;          /* 0x7FFF | ZL_NOTFIRST would collide with ZL_NOHASH */
}

/* length of the common prefix of the strings at q and q - d, at most maxl */
template <uint32_t RING> __device__ __forceinline__ uint32_t zl_match_len(const uint32_t *ring32, uint32_t q, uint32_t d, uint32_t maxl)
{
    uint32_t l = 0;
    while (l < maxl) {
        uint32_t x = zl_ld32<RING>(ring32, q + l) ^ zl_ld32<RING>(ring32, q + l - d);
        if (x) { l += (uint32_t)(__ffs((int)x) - 1) >> 3; break; }
        l += 4;
    }
    return l < maxl ? l : maxl;
}

/* common prefix of the strings at q and q - d, looking at 16 bytes only (branch-free: 0..16) */
template <uint32_t RING> __device__ __forceinline__ uint32_t zl_match16(const uint32_t *ring32, uint32_t q, uint32_t d)
{
    const uint32_t qb = q - d;
    const uint32_t ia = (q >> 2) & (RING / 4 - 1), ib = (qb >> 2) & (RING / 4 - 1);
    const uint32_t sa = (q & 3) * 8, sb = (qb & 3) * 8;
    uint32_t a[5], b[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { a[k] = ring32[ia + k]; b[k] = ring32[ib + k]; }
    uint32_t len = 16;
#pragma unroll
    for (int k = 3; k >= 0; k--) {
        const uint32_t x = __funnelshift_r(a[k], a[k + 1], sa) ^ __funnelshift_r(b[k], b[k + 1], sb);
        if (x) len = 4 * k + ((uint32_t)(__ffs((int)x) - 1) >> 3);
    }
    return len;
}

/* same, eight bytes per step (three aligned words per side, two funnel shifts) */
template <uint32_t RING> __device__ __forceinline__ uint32_t zl_match_ext(const uint32_t *ring32, uint32_t q, uint32_t d, uint32_t limit)
{
    uint32_t l = 0;
    while (l < limit) {
        const uint32_t qa = q + l, qb = qa - d;
        const uint32_t ia = (qa >> 2) & (RING / 4 - 1), ib = (qb >> 2) & (RING / 4 - 1);
        const uint32_t a0 = ring32[ia], a1 = ring32[ia + 1], a2 = ring32[ia + 2];
        const uint32_t b0 = ring32[ib], b1 = ring32[ib + 1], b2 = ring32[ib + 2];
        const uint32_t sa = (qa & 3) * 8, sb = (qb & 3) * 8;
        const uint32_t x0 = __funnelshift_r(a0, a1, sa) ^ __funnelshift_r(b0, b1, sb);
        const uint32_t x1 = __funnelshift_r(a1, a2, sa) ^ __funnelshift_r(b1, b2, sb);
        if (x0) { l += (uint32_t)(__ffs((int)x0) - 1) >> 3; break; }
        if (x1) { l += 4 + ((uint32_t)(__ffs((int)x1) - 1) >> 3); break; }
        l += 8;
    }
    return l < limit ? l : limit;
}


/* ---- bulk-copy staging: one elected thread moves whole 16-byte-aligned pieces of the input into the ring with
 * cp.async.bulk (the 1-D form of the TMA copy: no tensor map) and the consumers wait on an mbarrier that counts the
 * bytes.  The worker warps issue no load/store instructions for staging at all. */
__device__ __forceinline__ uint32_t zl_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void zl_mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(zl_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void zl_mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(zl_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void zl_mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tZL_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra ZL_DONE;\n\tbra ZL_WAIT;\n\tZL_DONE:\n\t}"
                 ::"r"(zl_smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void zl_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(zl_smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(zl_smem_addr(bar)) : "memory");
}
/* Called by ONE thread: queue the copies of input positions [from, to) (both multiples of 16) into the ring, the ring's
 * first ZL_MIRROR bytes also into the mirror behind its end, and tell the barrier how many bytes to expect. */
template <uint32_t RING> __device__ __forceinline__ void zl_stage_bulk(uint32_t *ring32, const uint8_t *gbase, uint32_t from, uint32_t to, unsigned long long *bar)
{
    uint8_t *ring8 = (uint8_t *)ring32;
    uint32_t bytes = to - from;
    const uint32_t wrap = (from | (RING - 1)) + 1;                 /* next multiple of RING above `from` */
    if (to <= wrap && (from & (RING - 1)) >= ZL_MIRROR) {          /* the usual case: one piece, clear of the ring's ends */
        zl_mbar_expect_tx(bar, bytes);
        zl_bulk_g2s(ring8 + (from & (RING - 1)), gbase + from, bytes, bar);
        return;
    }
    for (uint32_t h = 0; h < ZL_MIRROR; h += 16) {
        const uint32_t m0 = (from & ~(RING - 1)) + h, m1 = wrap + h;   /* positions whose ring offset is h */
        if ((m0 >= from && m0 < to) || (m1 >= from && m1 < to)) bytes += 16;
    }
    zl_mbar_expect_tx(bar, bytes);
    const uint32_t e1 = to < wrap ? to : wrap;
    zl_bulk_g2s(ring8 + (from & (RING - 1)), gbase + from, e1 - from, bar);
    if (to > wrap) zl_bulk_g2s(ring8, gbase + wrap, to - wrap, bar);
    for (uint32_t h = 0; h < ZL_MIRROR; h += 16) {
        const uint32_t m0 = (from & ~(RING - 1)) + h, m1 = wrap + h;
        if (m0 >= from && m0 < to) zl_bulk_g2s(ring8 + RING + h, gbase + m0, 16, bar);
        else if (m1 >= from && m1 < to) zl_bulk_g2s(ring8 + RING + h, gbase + m1, 16, bar);
    }
}

#endif
