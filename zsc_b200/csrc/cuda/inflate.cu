/* inflate.cu — batched inflate: one thread per stream, decode tables in shared memory (sm_100a).
 *
 * GPU form of zsc_uncompress's hot loop (reference src/zsc_uncompr.c:103-127 -> inflate /
 * inflate_fast / inflate_table).  Streams are independent, so the batch is spread one stream per
 * thread; the 32 threads of a warp keep their 32 sets of tables (1856 B each, inflate_core.h) in
 * shared memory and decode in lockstep, which works well because reference-compressed streams end
 * their blocks on the same symbol count.  The data check (adler32 of the output against the
 * trailer, reference src/inflate.c:1322-1342) is a second, HBM-streaming pass over the output.
 */
#include "common.cuh"
#include "inflate_core.h"

#define ZI_THREADS 32

__global__ void __launch_bounds__(ZI_THREADS)
zs_inflate_kernel(uint32_t n, const ZsStream *__restrict__ streams, const uint8_t *__restrict__ comp,
                  uint8_t *__restrict__ raw, int32_t wrap, int32_t *__restrict__ ret,
                  uint32_t *__restrict__ produced, uint32_t *__restrict__ consumed,
                  uint32_t *__restrict__ aux /* [2n]: stored check, flags */, zi_aux *__restrict__ cold, zi_tables *__restrict__ tabs)
{
    const uint32_t s = blockIdx.x * ZI_THREADS + threadIdx.x;
#ifdef ZI_TABLES_IN_SMEM
    extern __shared__ __align__(16) unsigned char zi_smem_raw[];
    zi_tables *T = reinterpret_cast<zi_tables *>(zi_smem_raw) + threadIdx.x;
#else
    /* decode tables live in global memory (they stay L2 / L1 resident): without a shared-memory footprint the
       SM holds ~20 warps = 640 streams, and it is that parallelism, not table latency, that sets throughput */
    zi_tables *T = tabs + (s < n ? s : 0);
#endif
    /* first-level tables (codes of <= 7 bits, distance table): 384 B per stream in shared memory */
    __shared__ zi_fast fast[ZI_THREADS];
    zi_fast *F = &fast[threadIdx.x];
    zi_mach m;
    if (s < n) {
        const ZsStream st = streams[s];
        zi_m_init(&m, comp + st.comp_off, st.comp_cap, raw + st.raw_off, st.raw_len, wrap, T, cold + s, F);
    } else {
        m.state = ZM_DONE;
    }
    /* lockstep: every lane advances its own stream by one bounded step, then the warp re-converges */
    while (__any_sync(0xFFFFFFFFu, m.state != ZM_DONE)) {
        if (m.state != ZM_DONE) zi_step(&m);
    }
    if (s >= n) return;
    ret[s] = m.res.ret;
    produced[s] = m.res.produced;
    consumed[s] = m.res.consumed;
    aux[2 * s] = m.res.stored_check;
    aux[2 * s + 1] = m.res.have_check | (m.res.data_errors ? 2u : 0u);
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
                                         uint32_t max_raw_len, void *cold, unsigned long long tabs_off)
{
    if (n == 0) return cudaSuccess;
#ifdef ZI_TABLES_IN_SMEM
    const size_t smem = sizeof(zi_tables) * ZI_THREADS;
    cudaFuncSetAttribute(zs_inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#else
    const size_t smem = 0;
#endif
    zs_inflate_kernel<<<(n + ZI_THREADS - 1) / ZI_THREADS, ZI_THREADS, smem, st>>>(n, streams, comp, raw, wrap, ret, produced, consumed, aux, reinterpret_cast<zi_aux *>(cold), reinterpret_cast<zi_tables *>(reinterpret_cast<uint8_t *>(cold) + (size_t)tabs_off));
    cudaMemsetAsync(acc, 0, sizeof(ZsAdlerAcc) * n, st);
    cudaError_t ce = zs_adler_streams_launch(st, n, max_raw_len, raw, streams, produced, acc);
    if (ce != cudaSuccess) return ce;
    zs_inflate_check_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, acc, produced, aux, wrap, ret, check);
    return cudaGetLastError();
}
