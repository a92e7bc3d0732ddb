/* checksum.cu — Adler-32 and CRC-32 as HBM-streaming kernels with log-depth combination (sm_100a).
 *
 * Results are bit-identical to the reference's adler32_z (src/adler32.c:56-131) and crc32_z
 * (src/crc32.c:502-593); the method is not.  The reference walks bytes serially (16-way unrolled,
 * modulo every 5552 bytes; slicing-by-4 tables).  Here every piece of the buffer is summed
 * independently with 128-bit loads and the pieces are combined algebraically:
 *
 *   Adler:  a = 1 + sum b_i,  b = N + sum (N - i) b_i   (mod 65521)
 *           a piece starting at stream index p contributes (S1, (N - p) S1 - sum o b_o), o = offset
 *           in the piece; dp4a forms the byte sums and the 0..3-weighted sums four bytes at a time.
 *   CRC:    raw CRCs (zero init, no final xor) of the pieces are multiplied by x^(8 * bytes after
 *           the piece) mod P and XOR-ed; the init/final inversions are folded in on the host side
 *           of the engine (zscgpu_crc32).  The reference removed crc32_combine (src/crc32.c:636);
 *           the operator used here follows from CRC linearity.
 */
#include "common.cuh"

#define ZA_THREADS 256
#define ZA_PIECE 65536u               /* bytes per CTA */

/* Sums over one piece [p, p + n) whose first byte has weight W (= bytes from it to the end of the
 * stream, inclusive).  Adds (S1 mod 65521, S2 mod 65521) into acc with 64-bit atomics. */
__device__ __forceinline__ void za_piece(const uint8_t *__restrict__ p, uint32_t n, uint64_t W, ZsAdlerAcc *acc)
{
    __shared__ unsigned long long red[3][ZA_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t head = min(n, (uint32_t)((16 - ((uint64_t)p & 15)) & 15));
    const uint32_t nvec = (n - head) >> 4;
    const uint32_t tail0 = head + nvec * 16;
    uint32_t a_s = 0, a_j = 0;            /* sum of bytes; sum of (offset in vector) * byte */
    unsigned long long a_os = 0;          /* sum of (vector offset) * (vector byte sum) */
    const uint4 *vp = reinterpret_cast<const uint4 *>(p + head);
    for (uint32_t v = tid; v < nvec; v += ZA_THREADS) {
        uint4 x = __ldg(vp + v);
        uint32_t s0 = __dp4a(x.x, 0x01010101u, 0u), s1 = __dp4a(x.y, 0x01010101u, 0u);
        uint32_t s2 = __dp4a(x.z, 0x01010101u, 0u), s3 = __dp4a(x.w, 0x01010101u, 0u);
        uint32_t j = __dp4a(x.x, 0x03020100u, 0u);
        j = __dp4a(x.y, 0x07060504u, j);
        j = __dp4a(x.z, 0x0B0A0908u, j);
        j = __dp4a(x.w, 0x0F0E0D0Cu, j);
        uint32_t s = s0 + s1 + s2 + s3;
        a_s += s; a_j += j;
        a_os += (unsigned long long)(head + v * 16) * s;
    }
    /* ragged head and tail bytes */
    for (uint32_t o = tid; o < head; o += ZA_THREADS) { uint32_t b = p[o]; a_s += b; a_os += (unsigned long long)o * b; }
    for (uint32_t o = tail0 + tid; o < n; o += ZA_THREADS) { uint32_t b = p[o]; a_s += b; a_os += (unsigned long long)o * b; }
    unsigned long long r0 = a_s, r1 = a_os + a_j, r2 = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { r0 += __shfl_down_sync(0xFFFFFFFFu, r0, o); r1 += __shfl_down_sync(0xFFFFFFFFu, r1, o); }
    if (lane == 0) { red[0][warp] = r0; red[1][warp] = r1; }
    __syncthreads();
    if (tid == 0) {
        r0 = 0; r1 = 0;
        for (int w = 0; w < ZA_THREADS / 32; w++) { r0 += red[0][w]; r1 += red[1][w]; }
        unsigned long long S1 = r0 % ZS_ADLER_BASE;
        unsigned long long neg = r1 % ZS_ADLER_BASE;                       /* sum o * b_o */
        r2 = ((W % ZS_ADLER_BASE) * S1 + ZS_ADLER_BASE - neg) % ZS_ADLER_BASE;
        atomicAdd(&acc->s1, S1);
        atomicAdd(&acc->s2, r2);
    }
}

/* one CTA per (chunk, 64 KiB piece): deflate batches */
__global__ void __launch_bounds__(ZA_THREADS)
zs_adler_chunks_kernel(const uint8_t *__restrict__ raw, const ZsChunk *__restrict__ chunks,
                       const ZsStream *__restrict__ streams, ZsAdlerAcc *__restrict__ acc, uint32_t pieces_per_chunk)
{
    const uint32_t c = blockIdx.x / pieces_per_chunk, k = blockIdx.x % pieces_per_chunk;
    const ZsChunk cd = chunks[c];
    const uint32_t off = k * ZA_PIECE;
    if (off >= cd.len) return;
    const ZsStream st = streams[cd.stream];
    const uint32_t n = min(ZA_PIECE, cd.len - off);
    const uint64_t idx = cd.raw_off + off - st.raw_off;            /* stream index of the first byte */
    za_piece(raw + cd.raw_off + off, n, (uint64_t)st.raw_len - idx, &acc[cd.stream]);
}

/* one CTA per (stream, 64 KiB piece) of inflate output; produced[] is known only on the device */
__global__ void __launch_bounds__(ZA_THREADS)
zs_adler_streams_kernel(const uint8_t *__restrict__ raw, const ZsStream *__restrict__ streams,
                        const uint32_t *__restrict__ produced, ZsAdlerAcc *__restrict__ acc, uint32_t pieces_per_stream)
{
    const uint32_t s = blockIdx.x / pieces_per_stream, k = blockIdx.x % pieces_per_stream;
    const uint32_t len = produced[s];
    const uint64_t off = (uint64_t)k * ZA_PIECE;
    if (off >= len) return;
    const uint32_t n = (uint32_t)min((uint64_t)ZA_PIECE, len - off);
    za_piece(raw + streams[s].raw_off + off, n, len - off, &acc[s]);
}

/* one CTA per 64 KiB piece of a flat buffer: standalone checksum */
__global__ void __launch_bounds__(ZA_THREADS)
zs_adler_flat_kernel(const uint8_t *__restrict__ p, uint64_t len, ZsAdlerAcc *__restrict__ acc)
{
    for (uint64_t piece = blockIdx.x; piece * ZA_PIECE < len; piece += gridDim.x) {
        const uint64_t off = piece * ZA_PIECE;
        const uint32_t n = (uint32_t)min((uint64_t)ZA_PIECE, len - off);
        za_piece(p + off, n, len - off, acc);
        __syncthreads();
    }
}

/* ------------------------------- CRC-32 ------------------------------- */
#define ZC_POLY 0xEDB88320u
#define ZC_THREADS 256
#define ZC_SUB 256u                     /* bytes per thread */
#define ZC_PIECE (ZC_THREADS * ZC_SUB)  /* bytes per CTA */

#define ZF_THREADS 1024
#define ZF_ROW 512u                     /* bytes one warp loads at a time: 32 lanes x 16 bytes */
#define ZF_MIN_LEN (1u << 20)           /* shorter buffers take zs_crc_flat_kernel */

__device__ uint32_t zc_table[8][256];   /* slicing tables, filled once by zs_crc_init_kernel */
__device__ uint32_t zc_x2n[32];         /* x^(2^k) mod P, reflected */
__device__ uint32_t zc_ftab[4][256];    /* fold tables: state byte k -> the state 512 zero bytes later (zs_crc_fold_kernel) */
__device__ uint32_t zc_ffin[128];       /* x^(8 * (512 - 16 lane - 4 k)) mod P: a row's word to the end of that row */

__host__ __device__ inline uint32_t zc_multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ ZC_POLY : b >> 1;
    }
    return p;
}
/* x^(n * 2^k) mod P */
__device__ inline uint32_t zc_x2nmodp(uint64_t n, uint32_t k)
{
    uint32_t p = 1u << 31;
    while (n) { if (n & 1) p = zc_multmodp(zc_x2n[k & 31], p); n >>= 1; k++; }
    return p;
}

__global__ void zs_crc_init_kernel()
{
    uint32_t t = threadIdx.x;
    if (t < 256) {
        uint32_t c = t;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ ZC_POLY : c >> 1;
        zc_table[0][t] = c;
    }
    __syncthreads();
    if (t < 256) {
        uint32_t c = zc_table[0][t];
        for (int k = 1; k < 8; k++) { c = zc_table[0][c & 0xFF] ^ (c >> 8); zc_table[k][t] = c; }
    }
    if (t == 0) {
        uint32_t p = 1u << 30;            /* x^1 */
        zc_x2n[0] = p;
        for (int n = 1; n < 32; n++) { p = zc_multmodp(p, p); zc_x2n[n] = p; }
    }
    __syncthreads();
    if (t < 256) {
        const uint32_t xs = zc_x2nmodp(ZF_ROW, 3);
        for (int k = 0; k < 4; k++) zc_ftab[k][t] = zc_multmodp(xs, t << (8 * k));
    }
    if (t < 128) zc_ffin[t] = zc_x2nmodp(ZF_ROW - 16u * (t >> 2) - 4u * (t & 3u), 3);
}

__global__ void __launch_bounds__(ZC_THREADS)
zs_crc_flat_kernel(const uint8_t *__restrict__ p, uint64_t len, uint32_t *__restrict__ acc)
{
    __shared__ uint32_t T[8][256];
    __shared__ uint32_t red[ZC_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < 8 * 256; i += ZC_THREADS) (&T[0][0])[i] = (&zc_table[0][0])[i];
    __syncthreads();
    for (uint64_t piece = blockIdx.x; piece * ZC_PIECE < len; piece += gridDim.x) {
        const uint64_t off = piece * ZC_PIECE + (uint64_t)tid * ZC_SUB;
        uint32_t contrib = 0;
        if (off < len) {
            const uint32_t n = (uint32_t)min((uint64_t)ZC_SUB, len - off);
            const uint8_t *q = p + off;
            uint32_t c = 0, i = 0;
            if ((((uint64_t)q) & 15) == 0) {
                for (; i + 16 <= n; i += 16) {
                    uint4 v = __ldg(reinterpret_cast<const uint4 *>(q + i));
                    uint32_t w0 = v.x ^ c, w1 = v.y;
                    c = T[7][w0 & 0xFF] ^ T[6][(w0 >> 8) & 0xFF] ^ T[5][(w0 >> 16) & 0xFF] ^ T[4][w0 >> 24] ^
                        T[3][w1 & 0xFF] ^ T[2][(w1 >> 8) & 0xFF] ^ T[1][(w1 >> 16) & 0xFF] ^ T[0][w1 >> 24];
                    w0 = v.z ^ c; w1 = v.w;
                    c = T[7][w0 & 0xFF] ^ T[6][(w0 >> 8) & 0xFF] ^ T[5][(w0 >> 16) & 0xFF] ^ T[4][w0 >> 24] ^
                        T[3][w1 & 0xFF] ^ T[2][(w1 >> 8) & 0xFF] ^ T[1][(w1 >> 16) & 0xFF] ^ T[0][w1 >> 24];
                }
            }
            for (; i < n; i++) c = T[0][(c ^ q[i]) & 0xFF] ^ (c >> 8);
            const uint64_t after = len - off - n;
            contrib = after ? zc_multmodp(zc_x2nmodp(after, 3), c) : c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib ^= __shfl_down_sync(0xFFFFFFFFu, contrib, o);
        if (lane == 0) red[warp] = contrib;
        __syncthreads();
        if (tid == 0) {
            uint32_t r = 0;
            for (int w = 0; w < ZC_THREADS / 32; w++) r ^= red[w];
            atomicXor(acc, r);
        }
        __syncthreads();
    }
}

/* a * b mod P without branches (same product as zc_multmodp) */
__device__ __forceinline__ uint32_t zc_mul(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
#pragma unroll
    for (int i = 31; i >= 0; i--) {
        p ^= b & (0u - ((a >> i) & 1u));
        b = (b >> 1) ^ (ZC_POLY & (0u - (b & 1u)));
    }
    return p;
}

/* raw CRC of n < 512 bytes at p[off..], moved to the end of a buffer of len bytes */
__device__ uint32_t zc_small(const uint8_t *__restrict__ p, uint64_t off, uint32_t n, uint64_t len)
{
    uint32_t c = 0;
    for (uint32_t i = 0; i < n; i++) c = zc_table[0][(c ^ p[off + i]) & 0xFF] ^ (c >> 8);
    const uint64_t after = len - off - n;
    return after ? zc_multmodp(zc_x2nmodp(after, 3), c) : c;
}

/* CRC-32 of a large buffer at HBM speed.  The CRC register is linear in the message, so the buffer is read
 * as it lies in memory — a warp loads rows of 512 bytes, lane l the 16 bytes at 16 l: perfectly coalesced —
 * and every (lane, word-of-four) pair keeps its own register for the sub-message made of its words with
 * zeros everywhere else.  Between two of its words lie 508 foreign bytes, so one step is
 * c = F512(c ^ word): four table lookups, exactly the cost of slicing-by-4, with tables built for a
 * 512-byte advance.  The tables are replicated per lane in shared memory ([entry][lane], 128 KiB), so the
 * 32 lookups of a warp instruction hit 32 different banks whatever the data.  The last row of a region
 * is advanced only to the region's end (x^(8 n) constants per lane and word), the 128 registers are XOR-ed,
 * and the region's CRC is moved to the end of the buffer with x^(8 * bytes after), a product the lanes form
 * together (one factor per bit of the distance, five shuffle steps). */
__global__ void __launch_bounds__(ZF_THREADS, 1)
zs_crc_fold_kernel(const uint8_t *__restrict__ p, uint64_t len, uint32_t head, uint64_t rows, uint32_t R,
                   uint32_t *__restrict__ acc)
{
    extern __shared__ __align__(16) uint32_t zf_tab[];             /* [4 * 256][32] */
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll 4
    for (uint32_t j = 0; j < 32; j++) { const uint32_t idx = j * ZF_THREADS + tid; zf_tab[idx] = (&zc_ftab[0][0])[idx >> 5]; }
    __syncthreads();
    const uint32_t *tl = zf_tab + lane;
#define ZF_STEP(c, w) { const uint32_t v_ = (c) ^ (w); \
        (c) = tl[(v_ & 0xFFu) * 32u] ^ tl[(256u + ((v_ >> 8) & 0xFFu)) * 32u] ^ tl[(512u + ((v_ >> 16) & 0xFFu)) * 32u] ^ tl[(768u + (v_ >> 24)) * 32u]; }
    const uint4 *base = reinterpret_cast<const uint4 *>(p + head);
    const uint64_t nwarps = (uint64_t)gridDim.x * (ZF_THREADS / 32);
    for (uint64_t piece = (uint64_t)blockIdx.x * (ZF_THREADS / 32) + warp; piece * R < rows; piece += nwarps) {
        const uint64_t r0 = piece * R;
        const uint32_t nr = (uint32_t)min((uint64_t)R, rows - r0);
        const uint4 *q = base + r0 * 32 + lane;
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, r = 0;
        for (; r + 4 < nr; r += 4) {
            const uint4 v0 = __ldg(q + (size_t)r * 32), v1 = __ldg(q + (size_t)(r + 1) * 32);
            const uint4 v2 = __ldg(q + (size_t)(r + 2) * 32), v3 = __ldg(q + (size_t)(r + 3) * 32);
            ZF_STEP(c0, v0.x) ZF_STEP(c1, v0.y) ZF_STEP(c2, v0.z) ZF_STEP(c3, v0.w)
            ZF_STEP(c0, v1.x) ZF_STEP(c1, v1.y) ZF_STEP(c2, v1.z) ZF_STEP(c3, v1.w)
            ZF_STEP(c0, v2.x) ZF_STEP(c1, v2.y) ZF_STEP(c2, v2.z) ZF_STEP(c3, v2.w)
            ZF_STEP(c0, v3.x) ZF_STEP(c1, v3.y) ZF_STEP(c2, v3.z) ZF_STEP(c3, v3.w)
        }
        for (; r + 1 < nr; r++) {
            const uint4 v = __ldg(q + (size_t)r * 32);
            ZF_STEP(c0, v.x) ZF_STEP(c1, v.y) ZF_STEP(c2, v.z) ZF_STEP(c3, v.w)
        }
        const uint4 v = __ldg(q + (size_t)(nr - 1) * 32);
        uint32_t x = zc_mul(zc_ffin[lane * 4 + 0], c0 ^ v.x) ^ zc_mul(zc_ffin[lane * 4 + 1], c1 ^ v.y) ^
                     zc_mul(zc_ffin[lane * 4 + 2], c2 ^ v.z) ^ zc_mul(zc_ffin[lane * 4 + 3], c3 ^ v.w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x ^= __shfl_xor_sync(0xFFFFFFFFu, x, o);
        const uint64_t after = len - head - (r0 + nr) * ZF_ROW;
        uint32_t f = ((after >> lane) & 1u) ? zc_x2n[(lane + 3) & 31] : 0x80000000u;
        if (after >> 32) { if ((after >> (lane + 32)) & 1u) f = zc_mul(f, zc_x2n[(lane + 35) & 31]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) f = zc_mul(f, __shfl_xor_sync(0xFFFFFFFFu, f, o));
        if (lane == 0) atomicXor(acc, zc_mul(f, x));
    }
#undef ZF_STEP
    if (blockIdx.x == 0 && warp < 2) {
        /* the unaligned head (< 16 bytes) and the bytes behind the last whole row (< 512) */
        const uint64_t toff = head + rows * ZF_ROW;
        uint32_t x = 0;
        if (warp == 0) { const uint64_t o = toff + 16ull * lane; if (o < len) x = zc_small(p, o, (uint32_t)min((uint64_t)16, len - o), len); }
        else if (lane == 0 && head) x = zc_small(p, 0, head, len);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x ^= __shfl_xor_sync(0xFFFFFFFFu, x, o);
        if (lane == 0 && x) atomicXor(acc, x);
    }
}

/* acc[0] = raw CRC of the buffer; finish: crc = x^(8 len) * (init ^ ~0) ^ raw ^ ~0 */
__global__ void zs_crc_finish_kernel(uint32_t *acc, uint64_t len, uint32_t init)
{
    uint32_t pre = init ^ 0xFFFFFFFFu;
    uint32_t r = acc[0] ^ (len ? zc_multmodp(zc_x2nmodp(len, 3), pre) : pre);
    acc[1] = r ^ 0xFFFFFFFFu;
}

extern "C" cudaError_t zs_adler_chunks_launch(cudaStream_t st, uint32_t nchunks, const uint8_t *raw,
                                              const ZsChunk *chunks, const ZsStream *streams, ZsAdlerAcc *acc)
{
    if (nchunks == 0) return cudaSuccess;
    const uint32_t ppc = ZS_CHUNK_MAX / ZA_PIECE;
    zs_adler_chunks_kernel<<<nchunks * ppc, ZA_THREADS, 0, st>>>(raw, chunks, streams, acc, ppc);
    return cudaGetLastError();
}
extern "C" cudaError_t zs_adler_streams_launch(cudaStream_t st, uint32_t n, uint32_t max_len, const uint8_t *raw,
                                               const ZsStream *streams, const uint32_t *produced, ZsAdlerAcc *acc)
{
    uint32_t pps = (max_len + ZA_PIECE - 1) / ZA_PIECE;
    if (n == 0 || pps == 0) return cudaSuccess;
    if ((uint64_t)n * pps > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    zs_adler_streams_kernel<<<n * pps, ZA_THREADS, 0, st>>>(raw, streams, produced, acc, pps);
    return cudaGetLastError();
}
extern "C" cudaError_t zs_adler_flat_launch(cudaStream_t st, const uint8_t *p, uint64_t len, ZsAdlerAcc *acc, int sms)
{
    uint64_t pieces = (len + ZA_PIECE - 1) / ZA_PIECE;
    uint32_t grid = (uint32_t)(pieces < (uint64_t)sms * 16 ? (pieces ? pieces : 1) : (uint64_t)sms * 16);
    zs_adler_flat_kernel<<<grid, ZA_THREADS, 0, st>>>(p, len, acc);
    return cudaGetLastError();
}
extern "C" cudaError_t zs_crc_init_launch(cudaStream_t st)
{
    zs_crc_init_kernel<<<1, 256, 0, st>>>();
    return cudaGetLastError();
}
extern "C" cudaError_t zs_crc_flat_launch(cudaStream_t st, const uint8_t *p, uint64_t len, uint32_t init, uint32_t *acc2, int sms)
{
    if (len >= ZF_MIN_LEN) {
        /* rows of 512 bytes behind the 16-byte alignment point, R rows per warp and pass, passes balanced */
        const uint32_t head = (uint32_t)((16 - ((uintptr_t)p & 15)) & 15);
        const uint64_t rows = (len - head) / ZF_ROW;
        const uint64_t W = (uint64_t)sms * (ZF_THREADS / 32);
        const uint64_t passes = (rows + W * 128 - 1) / (W * 128);
        uint64_t R = (rows + W * passes - 1) / (W * passes);
        if (R < 16) R = 16;
        const uint64_t warps = (rows + R - 1) / R;
        const uint32_t grid = (uint32_t)((warps + ZF_THREADS / 32 - 1) / (ZF_THREADS / 32) < (uint64_t)sms ? (warps + ZF_THREADS / 32 - 1) / (ZF_THREADS / 32) : (uint64_t)sms);
        const int smem = 4 * 256 * 32 * 4;
        cudaFuncSetAttribute(zs_crc_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        zs_crc_fold_kernel<<<grid, ZF_THREADS, smem, st>>>(p, len, head, rows, (uint32_t)R, acc2);
    } else {
        uint64_t pieces = (len + ZC_PIECE - 1) / ZC_PIECE;
        uint32_t grid = (uint32_t)(pieces < (uint64_t)sms * 8 ? (pieces ? pieces : 1) : (uint64_t)sms * 8);
        zs_crc_flat_kernel<<<grid, ZC_THREADS, 0, st>>>(p, len, acc2);
    }
    zs_crc_finish_kernel<<<1, 1, 0, st>>>(acc2, len, init);
    return cudaGetLastError();
}
