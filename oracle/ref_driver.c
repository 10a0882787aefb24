/*
 * ref_driver.c -- restates, in C, the call sequence the reference's Haskell
 * shim performs around the codec (TEST INFRASTRUCTURE / CPU BASELINE ONLY).
 *
 * It is compiled twice by oracle/Makefile:
 *   - with -DDRV_USE_REFERENCE together with the reference's own, unmodified
 *     /root/reference/cbits/lz4.c  ->  oracle/_ref/libreflz4.so
 *     (cpu_baseline.kind == "reference"; the file is read where it lies, never
 *     copied into this repository);
 *   - together with oracle/lz4_oracle.c  ->  oracle/liboracle.so
 *     (cpu_baseline.kind == "port").
 *
 * Call sequences restated (reference file:line):
 *   compressChunk        src/Streamly/Internal/LZ4.hs:226-281   bound -> alloc -> compress_fast_continue -> header pokes
 *   decompressChunk      src/Streamly/Internal/LZ4.hs:290-336   header read -> checks -> decompress_safe_continue
 *   compressChunksD      src/Streamly/Internal/LZ4.hs:353-394   one ctx per stream, arrays strictly in order
 *   decompressChunksRawD src/Streamly/Internal/LZ4.hs:539-567   one decode ctx per stream, previous OUTPUT is the dictionary
 *   independent mode     test/Main.hs:57-65                     fresh ctx pair per block
 *   header layout        src/Streamly/Internal/LZ4.hs:177-207, :86-94  [compLen LE32][uncompLen LE32 iff BlockHasSize]
 *
 * Arrays of one stream must not be address-adjacent (SURVEY.md section 5
 * quirk 1); the driver returns DRV_E_ADJACENT instead of silently letting the
 * codec switch to prefix mode.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#ifdef DRV_USE_REFERENCE
#include "lz4.h"      /* found with -I/root/reference/cbits */
typedef LZ4_stream_t        cstream_t;
typedef LZ4_streamDecode_t  dstream_t;
#define C_CREATE()                LZ4_createStream()
#define C_FREE(s)                 LZ4_freeStream(s)
#define D_CREATE()                LZ4_createStreamDecode()
#define D_FREE(s)                 LZ4_freeStreamDecode(s)
#define C_BOUND(n)                LZ4_compressBound(n)
#define C_CONT(s,src,dst,n,cap,a) LZ4_compress_fast_continue(s,(const char*)(src),(char*)(dst),n,cap,a)
#define D_CONT(s,src,dst,n,cap)   LZ4_decompress_safe_continue(s,(const char*)(src),(char*)(dst),n,cap)
#define DRV_KIND 1
#else
typedef struct ora_cstream cstream_t;
typedef struct ora_dstream dstream_t;
cstream_t* ora_cstream_create(void); void ora_cstream_free(cstream_t*);
dstream_t* ora_dstream_create(void); void ora_dstream_free(dstream_t*);
int ora_compress_bound(int);
int ora_compress_continue(cstream_t*, const uint8_t*, uint8_t*, int, int, int);
int ora_decompress_continue(dstream_t*, const uint8_t*, uint8_t*, int, int);
#define C_CREATE()                ora_cstream_create()
#define C_FREE(s)                 ora_cstream_free(s)
#define D_CREATE()                ora_dstream_create()
#define D_FREE(s)                 ora_dstream_free(s)
#define C_BOUND(n)                ora_compress_bound(n)
#define C_CONT(s,src,dst,n,cap,a) ora_compress_continue(s,src,dst,n,cap,a)
#define D_CONT(s,src,dst,n,cap)   ora_decompress_continue(s,src,dst,n,cap)
#define DRV_KIND 0
#endif

#define DRV_E_ADJACENT  (-100)
#define DRV_E_TOOBIG    (-101)
#define DRV_E_CODEC     (-102)
#define DRV_E_HEADER    (-103)

int drv_kind(void) { return DRV_KIND; }   /* 1 = the reference's lz4.c, 0 = the restatement */

static void put_le32(uint8_t* p, int32_t v)
{ uint32_t u = (uint32_t)v; p[0] = (uint8_t)u; p[1] = (uint8_t)(u >> 8); p[2] = (uint8_t)(u >> 16); p[3] = (uint8_t)(u >> 24); }
static int32_t get_le32(const uint8_t* p)
{ return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24)); }

typedef struct {
    /* one job = one stream = arrays [first, last) processed in order through one ctx */
    const uint8_t* const* src; const int32_t* src_len;
    uint8_t* const* dst; const int32_t* dst_cap; int32_t* out_len;
    const int32_t* stream_first; int n_streams;
    int accel; int meta;           /* header bytes: 8, 4 or 0 */
    int max_block;                 /* decode capacity for meta==4; compress guard */
    int decode;
    volatile int next;             /* work counter */
    int status;
    pthread_mutex_t mu;
} job_t;

static int run_compress_stream(job_t* j, int s)
{
    cstream_t* ctx = C_CREATE();
    int rc = 0;
    if (!ctx) return DRV_E_CODEC;
    for (int i = j->stream_first[s]; i < j->stream_first[s + 1]; i++) {
        int n = j->src_len[i];
        int bound, r;
        if (i > j->stream_first[s] && j->src[i - 1] + j->src_len[i - 1] == j->src[i]) { rc = DRV_E_ADJACENT; break; }
        if (j->max_block > 0 && n > j->max_block) { rc = DRV_E_TOOBIG; j->out_len[i] = 0; break; }   /* LZ4.hs:237-241 */
        bound = C_BOUND(n);                                                /* LZ4.hs:244 */
        if (bound <= 0 || j->dst_cap[i] < bound + j->meta) { rc = DRV_E_TOOBIG; j->out_len[i] = 0; break; }
        r = C_CONT(ctx, j->src[i], j->dst[i] + j->meta, n, bound, j->accel);   /* LZ4.hs:254-256 */
        if (r <= 0) { rc = DRV_E_CODEC; j->out_len[i] = 0; break; }        /* LZ4.hs:257-260 */
        if (j->meta == 8) put_le32(j->dst[i] + 4, n);                      /* LZ4.hs:261 */
        if (j->meta >= 4) put_le32(j->dst[i], r);                          /* LZ4.hs:262 */
        j->out_len[i] = r;
    }
    C_FREE(ctx);
    return rc;
}

static int run_decompress_stream(job_t* j, int s)
{
    dstream_t* ctx = D_CREATE();
    int rc = 0;
    if (!ctx) return DRV_E_CODEC;
    for (int i = j->stream_first[s]; i < j->stream_first[s + 1]; i++) {
        const uint8_t* a = j->src[i];
        int avail = j->src_len[i] - j->meta;
        int comp_len, cap, r;
        if (avail < 0) { rc = DRV_E_HEADER; j->out_len[i] = -1; break; }
        comp_len = j->meta >= 4 ? get_le32(a) : avail;                     /* LZ4.hs:303 */
        cap = j->meta == 8 ? get_le32(a + 4) : j->max_block;               /* LZ4.hs:189-198 */
        /* LZ4.hs:309-318, with quirk 2 fixed: the real array length bounds the read */
        if (comp_len <= 0 || comp_len > avail || cap < 0 || cap > j->dst_cap[i]) { rc = DRV_E_HEADER; j->out_len[i] = -1; break; }
        r = D_CONT(ctx, a + j->meta, j->dst[i], comp_len, cap);            /* LZ4.hs:322-324 */
        j->out_len[i] = r;
        if (r < 0) { rc = DRV_E_CODEC; break; }                            /* LZ4.hs:325-330 */
    }
    D_FREE(ctx);
    return rc;
}

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    for (;;) {
        int s, rc;
        pthread_mutex_lock(&j->mu);
        s = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (s >= j->n_streams) break;
        rc = j->decode ? run_decompress_stream(j, s) : run_compress_stream(j, s);
        if (rc) { pthread_mutex_lock(&j->mu); if (!j->status) j->status = rc; pthread_mutex_unlock(&j->mu); }
    }
    return NULL;
}

static int run(job_t* j, int threads)
{
    pthread_t th[256];
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads > j->n_streams) threads = j->n_streams > 0 ? j->n_streams : 1;
    j->next = 0; j->status = 0;
    pthread_mutex_init(&j->mu, NULL);
    if (threads == 1) worker(j);
    else {
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, j);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    pthread_mutex_destroy(&j->mu);
    return j->status;
}

/*
 * Compress n arrays.  Stream s owns arrays stream_first[s] .. stream_first[s+1]-1
 * and gets ONE fresh ctx (linked mode = few long streams; independent mode =
 * n streams of one array each).  dst[i] receives [header][LZ4 block];
 * out_len[i] = LZ4 payload length (0 on failure).  Streams are distributed
 * over `threads` pthreads.  Returns 0 or the first DRV_E_* seen.
 */
int drv_compress(const uint8_t* const* src, const int32_t* src_len, int n,
                 const int32_t* stream_first, int n_streams,
                 uint8_t* const* dst, const int32_t* dst_cap, int32_t* out_len,
                 int accel, int meta, int max_block, int threads)
{
    job_t j; memset(&j, 0, sizeof j); (void)n;
    j.src = src; j.src_len = src_len; j.dst = dst; j.dst_cap = dst_cap; j.out_len = out_len;
    j.stream_first = stream_first; j.n_streams = n_streams;
    j.accel = accel; j.meta = meta; j.max_block = max_block; j.decode = 0;
    return run(&j, threads);
}

/*
 * Decompress n framed arrays ([header][block], exactly one block per array).
 * dst[i] must have room for the header's uncompLen (meta==8) or max_block
 * (meta==4).  out_len[i] = bytes produced, < 0 on failure.
 */
int drv_decompress(const uint8_t* const* src, const int32_t* src_len, int n,
                   const int32_t* stream_first, int n_streams,
                   uint8_t* const* dst, const int32_t* dst_cap, int32_t* out_len,
                   int meta, int max_block, int threads)
{
    job_t j; memset(&j, 0, sizeof j); (void)n;
    j.src = src; j.src_len = src_len; j.dst = dst; j.dst_cap = dst_cap; j.out_len = out_len;
    j.stream_first = stream_first; j.n_streams = n_streams;
    j.meta = meta; j.max_block = max_block; j.decode = 1;
    return run(&j, threads);
}
