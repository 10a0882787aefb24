/*
 * ffi_replay.c -- the exact foreign-call sequences of haskell/Streamly/Internal/LZ4/B200.hs, from plain C.
 *
 * GHC is not available where this repository is built, so the Haskell shim cannot be compiled there.  This program
 * pins its contract at the C ABI instead: it marshals arguments exactly as the shim does (Int64 offsets, Int32
 * lengths, CInt counts, arrays staged at 16-byte aligned offsets with a 16-byte gap through b200lz4_gather_host,
 * stream table [0, n] with one persistent handle in linked mode, NULL table in independent mode, capacities taken
 * from the uncompLen header field, outputs sliced at dst_off / out_len, errors read with b200lz4_ctx_last_error) and
 * runs the three pipelines of haskell/benchmark/MainB200.hs:
 *
 *   compressChunksD      newSession -> [cstreamCreate] -> per batch: compressBound*, stage, compress_batch, sliceOut
 *   resizeChunksD        accumulate pieces, b200lz4_reframe per accumulated range, keep the rest
 *   decompressChunksRawD newSession -> [dstreamCreate] -> per batch: caps from headers, stage, decompress_batch, sliceOut
 *
 * over several batches of one stream (so the persistent stream state crosses calls), linked and independent, and
 * checks the round trip.  Exit code 0 = every call behaved as the shim expects.
 *
 *   gcc -O2 -Iinclude examples/ffi_replay.c -Lstreamly_lz4_b200 -lb200lz4 -Wl,-rpath,$PWD/streamly_lz4_b200 -o build/ffi_replay
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b200lz4.h"

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } } while (0)

typedef struct { uint8_t* p; int n; } arr_t;                    /* one Haskell `Array Word8` (separately allocated, pageable) */

static int align16(int n) { return (n + 15) / 16 * 16; }

static void fill(uint8_t* p, size_t n, uint32_t x)
{
    static const char* const words[8] = {"stream ", "array ", "block ", "header ", "lz4 ", "the ", "of ", "chunk\n"};
    size_t i = 0;
    while (i < n) {
        x = x * 1664525u + 1013904223u;
        if ((x >> 29) == 0) { p[i++] = (uint8_t)(x >> 8); continue; }          /* some noise */
        const char* w = words[(x >> 24) & 7];
        for (size_t k = 0; w[k] && i < n; k++) p[i++] = (uint8_t)w[k];
    }
}

/* Session of B200.hs */
typedef struct { b200lz4_ctx* ctx; uint8_t* src; int64_t src_cap; uint8_t* dst; int64_t dst_cap; } session_t;

static int ensure(uint8_t** p, int64_t* cap, int64_t need)
{
    if (need <= *cap) return 0;
    if (*p) b200lz4_host_free(*p);
    *cap = need + need / 4;
    *p = (uint8_t*)b200lz4_host_alloc((size_t)*cap);
    return *p ? 0 : 1;
}

/* stage: offsets = scanl (+ align16 (n + 16)), one b200lz4_gather_host call */
static int stage(session_t* s, const arr_t* a, int n, int64_t* off, int32_t* len, int64_t* src_bytes)
{
    int64_t at = 0;
    const void** ptrs = (const void**)malloc(sizeof(void*) * (size_t)(n ? n : 1));
    for (int i = 0; i < n; i++) { off[i] = at; len[i] = a[i].n; ptrs[i] = a[i].p; at += align16(a[i].n + 16); }
    if (ensure(&s->src, &s->src_cap, at + 16)) { free(ptrs); return 1; }
    int rc = b200lz4_gather_host(s->src, ptrs, off, len, n, 0);
    free(ptrs);
    *src_bytes = n ? off[n - 1] + len[n - 1] : 0;
    return rc;
}

/* compressBatch of B200.hs; outputs appended to `out` (fresh exact-size arrays, like sliceOut) */
static int compress_batch(session_t* s, b200lz4_cstream* strm, int speed, int meta, const arr_t* a, int n, arr_t* out)
{
    int64_t bounds = 0;
    for (int i = 0; i < n; i++) {
        int b = b200lz4_compress_bound(a[i].n);
        CHECK(b > 0, "compressChunk: compressed length <= 0.");
        bounds += b + meta;
    }
    CHECK(!ensure(&s->dst, &s->dst_cap, bounds), "host_alloc failed");
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)n); int32_t* len = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int64_t* dst_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1)); int32_t* out_len = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int64_t src_bytes = 0;
    CHECK(!stage(s, a, n, off, len, &src_bytes), "stage: %s", b200lz4_last_error());
    int32_t first[2] = {0, n};
    b200lz4_cstream* table[1] = {strm};
    int rc = b200lz4_compress_batch(s->ctx, s->src, src_bytes, off, len, n, strm ? first : NULL, strm ? 1 : 0, strm ? table : NULL,
                                    speed, meta, s->dst, s->dst_cap, dst_off, out_len);
    CHECK(rc == 0, "compressChunk: c_compressFastContinue failed. %s", b200lz4_ctx_last_error(s->ctx));
    for (int i = 0; i < n; i++) {
        out[i].n = (int)(dst_off[i + 1] - dst_off[i]);
        out[i].p = (uint8_t*)malloc((size_t)out[i].n + 1);
        memcpy(out[i].p, s->dst + dst_off[i], (size_t)out[i].n);
        CHECK(out[i].n == meta + out_len[i], "header + payload != slice");
    }
    free(off); free(len); free(dst_off); free(out_len);
    return 0;
}

static int decompress_batch(session_t* s, b200lz4_dstream* strm, int meta, int max_block, const arr_t* a, int n, arr_t* out)
{
    int64_t caps = 0;
    for (int i = 0; i < n; i++) {
        int32_t u = max_block;
        if (meta == 8 && a[i].n >= 8) { memcpy(&u, a[i].p + 4, 4); if (u < 0) u = 0; }      /* uncompLen LE32 (little-endian host) */
        caps += u;
    }
    CHECK(!ensure(&s->dst, &s->dst_cap, caps + 64), "host_alloc failed");
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)n); int32_t* len = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int64_t* dst_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1)); int32_t* out_len = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int64_t src_bytes = 0;
    CHECK(!stage(s, a, n, off, len, &src_bytes), "stage: %s", b200lz4_last_error());
    int32_t first[2] = {0, n};
    b200lz4_dstream* table[1] = {strm};
    int rc = b200lz4_decompress_batch(s->ctx, s->src, src_bytes, off, len, n, strm ? first : NULL, strm ? 1 : 0, strm ? table : NULL,
                                      meta, max_block, s->dst, s->dst_cap, dst_off, out_len);
    CHECK(rc == 0, "decompressChunk: c_decompressSafeContinue failed. %s", b200lz4_ctx_last_error(s->ctx));
    for (int i = 0; i < n; i++) {
        out[i].n = out_len[i];
        out[i].p = (uint8_t*)malloc((size_t)out[i].n + 1);
        memcpy(out[i].p, s->dst + dst_off[i], (size_t)out[i].n);
    }
    free(off); free(len); free(dst_off); free(out_len);
    return 0;
}

static int run(int independent)
{
    enum { N = 37, BATCH = 8, META = 8, PIECE = 50000 };
    session_t cs = {0}, ds = {0};
    CHECK(b200lz4_ctx_create(0, &cs.ctx) == 0, "b200lz4_ctx_create failed: %s", b200lz4_last_error());
    CHECK(b200lz4_ctx_create(0, &ds.ctx) == 0, "b200lz4_ctx_create failed: %s", b200lz4_last_error());
    b200lz4_cstream* cstrm = NULL; b200lz4_dstream* dstrm = NULL;
    if (!independent) {
        CHECK(b200lz4_cstream_create(cs.ctx, &cstrm) == 0, "b200lz4_cstream_create failed: %s", b200lz4_last_error());
        CHECK(b200lz4_dstream_create(ds.ctx, &dstrm) == 0, "b200lz4_dstream_create failed: %s", b200lz4_last_error());
    }
    /* the input stream: N arrays of uneven sizes, each its own allocation */
    arr_t in[N], comp[N];
    int64_t total = 0, ctotal = 0;
    for (int i = 0; i < N; i++) {
        in[i].n = (i % 7 == 3) ? 0 : 20000 + 3571 * i;
        in[i].p = (uint8_t*)malloc((size_t)in[i].n + 1);
        fill(in[i].p, (size_t)in[i].n, 7u + (uint32_t)i);
        total += in[i].n;
    }
    /* compressChunksD: BFill up to BATCH arrays, flush, BDrain */
    for (int b = 0; b < N; b += BATCH) {
        int n = N - b < BATCH ? N - b : BATCH;
        if (compress_batch(&cs, cstrm, 1, META, in + b, n, comp + b)) return 1;
    }
    /* write the stream, read it back in PIECE-byte pieces, resizeChunksD over b200lz4_reframe */
    for (int i = 0; i < N; i++) ctotal += comp[i].n;
    uint8_t* file = (uint8_t*)malloc((size_t)ctotal + 1);
    { int64_t at = 0; for (int i = 0; i < N; i++) { memcpy(file + at, comp[i].p, (size_t)comp[i].n); at += comp[i].n; } }
    arr_t framed[N]; int nframed = 0;
    uint8_t* acc = (uint8_t*)malloc((size_t)ctotal + PIECE); int64_t have = 0;
    for (int64_t at = 0; at < ctotal; at += PIECE) {
        int64_t piece = ctotal - at < PIECE ? ctotal - at : PIECE;
        memcpy(acc + have, file + at, (size_t)piece); have += piece;                       /* Array.spliceTwo */
        int64_t boff[64]; int32_t blen[64]; int64_t nfound = 0, used = 0; int ended = 0;
        int rc = b200lz4_reframe(acc, have, META, 0, boff, blen, 64, &nfound, &used, &ended);
        CHECK(rc == 0, "resizeChunksD: %s", b200lz4_last_error());
        for (int64_t k = 0; k < nfound; k++) {
            CHECK(nframed < N, "too many blocks");
            framed[nframed].n = blen[k]; framed[nframed].p = (uint8_t*)malloc((size_t)blen[k] + 1);
            memcpy(framed[nframed].p, acc + boff[k], (size_t)blen[k]); nframed++;
        }
        memmove(acc, acc + used, (size_t)(have - used)); have -= used;
    }
    CHECK(have == 0, "resizeChunksD: Incomplete block");
    CHECK(nframed == N, "re-framing found %d of %d blocks", nframed, N);
    for (int i = 0; i < N; i++) CHECK(framed[i].n == comp[i].n && memcmp(framed[i].p, comp[i].p, (size_t)comp[i].n) == 0, "re-framed block %d differs", i);
    /* decompressChunksRawD, different batch size so that the decode state crosses calls at other places */
    arr_t back[N];
    for (int b = 0; b < N; b += 5) {
        int n = N - b < 5 ? N - b : 5;
        if (decompress_batch(&ds, dstrm, META, 0, framed + b, n, back + b)) return 1;
    }
    for (int i = 0; i < N; i++) CHECK(back[i].n == in[i].n && memcmp(back[i].p, in[i].p, (size_t)in[i].n) == 0, "array %d does not round-trip", i);
    /* a failing batch must leave its text in the ctx (error path of the shim) */
    {
        arr_t bad = { (uint8_t*)malloc(16), 16 }, out1;
        memset(bad.p, 0xFF, 16);
        int64_t off = 0, doff[2]; int32_t len = 16, ol = 0; int64_t sb = 0;
        CHECK(!stage(&ds, &bad, 1, &off, &len, &sb), "stage");
        int rc = b200lz4_decompress_batch(ds.ctx, ds.src, sb, &off, &len, 1, NULL, 0, NULL, META, 0, ds.dst, ds.dst_cap, doff, &ol);
        CHECK(rc == B200LZ4_E_BLOCK && ol < 0 && strlen(b200lz4_ctx_last_error(ds.ctx)) > 0, "a corrupt block must fail with a message");
        (void)out1; free(bad.p);
    }
    float h2d, k, d2h;
    b200lz4_last_timing(cs.ctx, &h2d, &k, &d2h);
    printf("%s: %d arrays, %lld -> %lld bytes, re-framed from %d-byte pieces, round trip identical (last compress call: h2d %.3f ms, kernels %.3f ms, d2h %.3f ms)\n",
           independent ? "independent" : "linked     ", N, (long long)total, (long long)ctotal, PIECE, h2d, k, d2h);
    if (cstrm) b200lz4_cstream_free(cstrm);
    if (dstrm) b200lz4_dstream_free(dstrm);
    b200lz4_host_free(cs.src); b200lz4_host_free(cs.dst); b200lz4_host_free(ds.src); b200lz4_host_free(ds.dst);
    b200lz4_ctx_destroy(cs.ctx); b200lz4_ctx_destroy(ds.ctx);
    return 0;
}

int main(void)
{
    if (run(0)) return 1;
    if (run(1)) return 1;
    return 0;
}
