/* Test infrastructure: compiled INTO oracle/_ref/libzsc_ref.so next to the unmodified reference
 * sources.  Reports the reference's private struct sizes (they enter the pinned work-buffer sizes,
 * reference src/deflate.c:886-899 and src/inflate.c:270-273) and offers a multi-threaded driver
 * that calls the reference's own zsc_compress2 / zsc_uncompress on independent buffers, one pthread
 * per core, used ONLY as the CPU baseline arm of bench.py.  No reference code is restated here. */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "zsc/zsc_pub.h"
#include "zsc/deflate.h"
#include "zsc/inflate.h"

U32 refprobe_sizeof_deflate_state(void) { return (U32)sizeof(deflate_state); }
U32 refprobe_sizeof_inflate_state(void) { return (U32)sizeof(inflate_state); }
U32 refprobe_sizeof_z_stream(void) { return (U32)sizeof(z_stream); }

typedef struct {
    int kind;                 /* 0 compress, 1 uncompress */
    const U8 *src; const unsigned long long *src_off; const U32 *src_len;
    U8 *dst; const unsigned long long *dst_off; const U32 *dst_cap; U32 *dst_len;
    I32 *ret;
    U32 n; U32 max_block_len; I32 level; I32 strategy;
    volatile U32 *next;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    U32 wlen = 0;
    if (j->kind == 0) zsc_compress_get_min_work_buf_size(&wlen);
    else zsc_uncompress_get_min_work_buf_size(&wlen);
    U8 *work = (U8 *)malloc(wlen);
    for (;;) {
        U32 i = __sync_fetch_and_add(j->next, 1);
        if (i >= j->n) break;
        U32 dl = j->dst_cap[i];
        if (j->kind == 0) {
            j->ret[i] = zsc_compress2(j->dst + j->dst_off[i], &dl, j->src + j->src_off[i],
                                      j->src_len[i], j->max_block_len, work, wlen, j->level,
                                      DEF_WBITS, DEF_MEM_LEVEL, (ZlibStrategy)j->strategy);
        } else {
            U32 sl = j->src_len[i];
            j->ret[i] = zsc_uncompress(j->dst + j->dst_off[i], &dl, j->src + j->src_off[i], &sl,
                                       work, wlen);
        }
        j->dst_len[i] = dl;
    }
    free(work);
    return 0;
}

/* Runs n independent reference calls over `threads` pthreads. */
int refprobe_batch(int kind, int threads, U32 n, const U8 *src, const unsigned long long *src_off,
                   const U32 *src_len, U8 *dst, const unsigned long long *dst_off,
                   const U32 *dst_cap, U32 *dst_len, I32 *ret, U32 max_block_len, I32 level,
                   I32 strategy)
{
    volatile U32 next = 0;
    job_t j = { kind, src, src_off, src_len, dst, dst_off, dst_cap, dst_len, ret, n,
                max_block_len, level, strategy, &next };
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int t = 0; t < threads; t++) pthread_create(&th[t], 0, worker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
    return 0;
}
