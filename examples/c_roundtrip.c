/*
 * c_roundtrip.c -- libb200lz4.so from plain C, through include/b200lz4.h only.
 *
 * (1) batched interface: compress a buffer as independent 64 KiB blocks, decompress, compare;
 * (2) legacy interface: the exact call sequence of the reference's compressChunk / decompressChunk
 *     (src/Streamly/Internal/LZ4.hs:226-336) on one linked stream: LZ4_createStream, LZ4_compressBound,
 *     LZ4_compress_fast_continue, header pokes, LZ4_decompress_safe_continue.
 *
 *   gcc -O2 -Iinclude examples/c_roundtrip.c -Lstreamly_lz4_b200 -lb200lz4 -Wl,-rpath,$PWD/streamly_lz4_b200 -o build/c_roundtrip
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b200lz4.h"

static void fill(uint8_t* p, size_t n)
{   /* compressible, not trivial: words drawn from a small vocabulary */
    static const char* const words[8] = {"stream ", "array ", "block ", "header ", "lz4 ", "the ", "of ", "chunk\n"};
    uint32_t x = 12345;
    size_t i = 0;
    while (i < n) {
        x = x * 1664525u + 1013904223u;
        const char* w = words[(x >> 24) & 7];
        for (size_t k = 0; w[k] && i < n; k++) p[i++] = (uint8_t)w[k];
    }
}

int main(void)
{
    const int kBlock = 65536, kBlocks = 64;
    const int64_t total = (int64_t)kBlock * kBlocks;
    b200lz4_ctx* ctx = NULL;
    if (b200lz4_ctx_create(0, &ctx) != B200LZ4_OK) { fprintf(stderr, "ctx: %s\n", b200lz4_last_error()); return 2; }

    /* ---- (1) batched */
    uint8_t* src = (uint8_t*)b200lz4_host_alloc((size_t)total);
    const int64_t cap = (int64_t)kBlocks * (B200LZ4_HDR_BOTH + b200lz4_compress_bound(kBlock));
    uint8_t* comp = (uint8_t*)b200lz4_host_alloc((size_t)cap);
    uint8_t* back = (uint8_t*)b200lz4_host_alloc((size_t)total);
    fill(src, (size_t)total);
    int64_t off[64], coff[65], boff[65]; int32_t len[64], clen[64], blen[64], flen[64];
    for (int i = 0; i < kBlocks; i++) { off[i] = (int64_t)i * kBlock; len[i] = kBlock; }
    int rc = b200lz4_compress_batch(ctx, src, total, off, len, kBlocks, NULL, 0, NULL, 1, B200LZ4_HDR_BOTH, comp, cap, coff, clen);
    if (rc != B200LZ4_OK) { fprintf(stderr, "compress: %s\n", b200lz4_last_error()); return 3; }
    for (int i = 0; i < kBlocks; i++) flen[i] = (int32_t)(coff[i + 1] - coff[i]);
    rc = b200lz4_decompress_batch(ctx, comp, coff[kBlocks], coff, flen, kBlocks, NULL, 0, NULL, B200LZ4_HDR_BOTH, 0, back, total, boff, blen);
    if (rc != B200LZ4_OK || memcmp(src, back, (size_t)total) != 0) { fprintf(stderr, "decompress: %s\n", b200lz4_last_error()); return 4; }
    float h2d, k, d2h;
    b200lz4_last_timing(ctx, &h2d, &k, &d2h);
    printf("batched : %lld -> %lld bytes, round trip identical (last call: h2d %.3f ms, kernels %.3f ms, d2h %.3f ms)\n",
           (long long)total, (long long)coff[kBlocks], h2d, k, d2h);

    /* ---- (2) legacy symbols, one linked stream */
    LZ4_stream_t* cs = LZ4_createStream();
    LZ4_streamDecode_t* ds = LZ4_createStreamDecode();
    if (!cs || !ds) { fprintf(stderr, "legacy create failed\n"); return 5; }
    const int bound = LZ4_compressBound(kBlock);
    char* cbuf = (char*)malloc((size_t)bound + 8);
    char* outs[4];
    long long ctot = 0;
    for (int i = 0; i < 4; i++) {
        const char* in = (const char*)src + (size_t)i * kBlock;
        int n = LZ4_compress_fast_continue(cs, in, cbuf + 8, kBlock, bound, 1);
        if (n <= 0) { fprintf(stderr, "LZ4_compress_fast_continue failed\n"); return 6; }
        int32_t hdr[2] = { n, kBlock };                          /* [compLen LE32][uncompLen LE32] (little-endian host) */
        memcpy(cbuf, hdr, 8);
        outs[i] = (char*)malloc(kBlock);                           /* decode outputs stay alive: the next block's dictionary */
        int m = LZ4_decompress_safe_continue(ds, cbuf + 8, outs[i], n, kBlock);
        if (m != kBlock || memcmp(outs[i], in, kBlock) != 0) { fprintf(stderr, "legacy round trip failed at block %d\n", i); return 7; }
        ctot += n + 8;
    }
    printf("legacy  : 4 linked blocks, %d -> %lld bytes, round trip identical\n", 4 * kBlock, ctot);
    for (int i = 0; i < 4; i++) free(outs[i]);
    free(cbuf);
    LZ4_freeStream(cs); LZ4_freeStreamDecode(ds);
    b200lz4_host_free(src); b200lz4_host_free(comp); b200lz4_host_free(back);
    b200lz4_ctx_destroy(ctx);
    return 0;
}
