/*
 * datagen.c -- seeded synthetic corpora for tests and bench (no network, so the
 * reference's Canterbury corpora, download-corpora.sh:6-49, are replaced by
 * these; shapes follow SURVEY.md section 8(d) and the QuickCheck generators of
 * test/Main.hs:33-55).
 *
 * Every 64 KiB segment k of a corpus depends only on (kind, seed, k), so any
 * byte range can be produced independently (and in parallel) and the CPU and
 * GPU arms of a benchmark see identical bytes.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#define SEG 65536u

enum {
    GEN_TEXT = 0,      /* word-vocabulary text */
    GEN_RANDOM = 1,    /* uniform bytes, incompressible */
    GEN_SPARSE01 = 2,  /* bytes in {0,1}, P(1) = 1/16 */
    GEN_RECORDS = 3,   /* 48-byte record repeated, one noisy byte per record */
    GEN_MIXED = 4,     /* cycle text / random / sparse01 / records per 64 KiB segment */
    GEN_BITS01 = 5,    /* uniform over {0,1}              (test/Main.hs:33-41) */
    GEN_BIASED01 = 6,  /* 0 w.p. 0.9, 1 w.p. 0.1          (test/Main.hs:44-47) */
    GEN_ZERO = 7
};

static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
typedef struct { uint64_t s; } rng_t;
static inline uint64_t next64(rng_t* r) { r->s += 0x9E3779B97F4A7C15ULL; return mix64(r->s); }

#define VOCAB 2048
typedef struct { uint8_t len[VOCAB]; char w[VOCAB][12]; } vocab_t;

static void build_vocab(vocab_t* v, uint64_t seed)
{
    static const char letters[] = "eeeeeeetttttaaaaooooiiiinnnnsssshhhrrrddllcumwfgypbvkjxqz";
    rng_t r = { mix64(seed ^ 0x766F636162ULL) };
    for (int i = 0; i < VOCAB; i++) {
        uint64_t x = next64(&r);
        int L = 2 + (int)(x % 9);           /* 2..10 */
        v->len[i] = (uint8_t)L;
        for (int k = 0; k < L; k++) { x = next64(&r); v->w[i][k] = letters[x % (sizeof(letters) - 1)]; }
    }
}

static void seg_text(const vocab_t* v, rng_t* r, uint8_t* out)
{
    uint32_t p = 0;
    while (p < SEG) {
        uint64_t x = next64(r);
        /* skewed word choice: cube of a uniform picks low indices often */
        double u = (double)(x >> 11) * (1.0 / 9007199254740992.0);
        int idx = (int)(u * u * u * VOCAB);
        const char* w = v->w[idx]; int L = v->len[idx];
        for (int k = 0; k < L && p < SEG; k++) out[p++] = (uint8_t)w[k];
        if (p < SEG) {
            uint32_t c = (uint32_t)(x & 63);
            if (c == 0) { out[p++] = '.'; if (p < SEG) out[p++] = '\n'; }
            else if (c < 4) { out[p++] = ','; if (p < SEG) out[p++] = ' '; }
            else out[p++] = ' ';
        }
    }
}

static void seg_random(rng_t* r, uint8_t* out)
{
    for (uint32_t p = 0; p < SEG; p += 8) { uint64_t x = next64(r); memcpy(out + p, &x, 8); }
}

/* one byte per 4 random bits: value 1 iff the nibble is < thr (thr/16 probability) */
static void seg_nibble01(rng_t* r, uint8_t* out, unsigned thr)
{
    for (uint32_t p = 0; p < SEG; p += 16) {
        uint64_t x = next64(r);
        for (int k = 0; k < 16; k++) { out[p + k] = ((x & 15) < thr) ? 1 : 0; x >>= 4; }
    }
}

static void seg_biased(rng_t* r, uint8_t* out)   /* P(1) = 0.1 */
{
    for (uint32_t p = 0; p < SEG; p += 4) {
        uint64_t x = next64(r);
        for (int k = 0; k < 4; k++) { out[p + k] = ((x & 0xFFFF) < 6554) ? 1 : 0; x >>= 16; }
    }
}

static void seg_records(rng_t* r, uint8_t* out)
{
    uint8_t rec[48];
    for (int k = 0; k < 48; k += 8) { uint64_t x = next64(r); memcpy(rec + k, &x, 8); }
    uint32_t p = 0;
    while (p < SEG) {
        uint64_t x = next64(r);
        uint32_t room = SEG - p, L = room < 48 ? room : 48;
        memcpy(out + p, rec, L);
        uint32_t at = (uint32_t)(x % 48);
        if (at < L) out[p + at] = (uint8_t)(x >> 32);
        p += L;
    }
}

static void gen_segment(int kind, uint64_t seed, uint64_t k, const vocab_t* v, uint8_t* out)
{
    rng_t r = { mix64(seed * 0x100000001B3ULL + k) };
    if (kind == GEN_MIXED) kind = (int)(k & 3);     /* text, random, sparse01, records */
    switch (kind) {
    case GEN_TEXT: seg_text(v, &r, out); break;
    case GEN_RANDOM: seg_random(&r, out); break;
    case GEN_SPARSE01: seg_nibble01(&r, out, 1); break;
    case GEN_RECORDS: seg_records(&r, out); break;
    case GEN_BITS01: seg_nibble01(&r, out, 8); break;
    case GEN_BIASED01: seg_biased(&r, out); break;
    default: memset(out, 0, SEG); break;
    }
}

/* Fill out[0..len) with bytes [offset, offset+len) of corpus (kind, seed). */
void b200gen_fill(int kind, uint64_t seed, uint64_t offset, uint64_t len, uint8_t* out)
{
    vocab_t* v = NULL;
    uint8_t* tmp = NULL;
    if (kind == GEN_TEXT || kind == GEN_MIXED) { v = (vocab_t*)malloc(sizeof *v); build_vocab(v, seed); }
    uint64_t pos = offset, end = offset + len;
    while (pos < end) {
        uint64_t k = pos / SEG, seg_start = k * SEG;
        if (pos == seg_start && end - pos >= SEG) {
            gen_segment(kind, seed, k, v, out + (pos - offset));
            pos += SEG;
        } else {
            if (!tmp) tmp = (uint8_t*)malloc(SEG);
            gen_segment(kind, seed, k, v, tmp);
            uint64_t a = pos - seg_start, b = (end - seg_start < SEG) ? end - seg_start : SEG;
            memcpy(out + (pos - offset), tmp + a, (size_t)(b - a));
            pos = seg_start + b;
        }
    }
    free(tmp); free(v);
}
