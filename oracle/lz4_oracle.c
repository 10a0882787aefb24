/*
 * lz4_oracle.c -- CPU restatement of the LZ4 block codec path that
 * streamly-lz4 drives (TEST INFRASTRUCTURE ONLY).
 *
 * This file is a checker.  Nothing in the product path (libb200lz4.so, the
 * host mirror, bench.py's GPU arm) may call, link or import it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may.  It is a from-the-spec restatement of what the reference computes,
 * written against SURVEY.md section 8(A)/(B); it is NOT a copy of cbits/lz4.c
 * and deliberately has a different shape (index arithmetic instead of pointer
 * arithmetic, one explicit probe routine, a byte-exact common-prefix counter).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 * byte-for-byte against the reference's own vendored LZ4 1.9.3
 * (/root/reference/cbits/lz4.c compiled unmodified into oracle/_ref/ by
 * oracle/Makefile) on seeded inputs, and against the committed golden
 * fixtures under tests/golden/ that were generated from that build.
 *
 * What is restated (reference file:line):
 *   ora_compress_bound        LZ4_COMPRESSBOUND            cbits/lz4.h:170-171, cbits/lz4.c:674
 *   ora_cstream_*             LZ4_createStream/initStream  cbits/lz4.c:1423-1451, state cbits/lz4.h:595-603
 *   ora_compress_continue     LZ4_compress_fast_continue   cbits/lz4.c:1565-1637 (ext-dict branch :1607-1636)
 *     -> encode_block         LZ4_compress_generic_validated(limitedOutput, byU32, usingExtDict,
 *                             {noDictIssue,dictSmall})     cbits/lz4.c:851-1240
 *     -> hash5                LZ4_hash5 / LZ4_hashPosition cbits/lz4.c:706-722
 *     -> common_prefix        LZ4_count                    cbits/lz4.c:603-626
 *     -> renorm               LZ4_renormDictT              cbits/lz4.c:1545-1562
 *   ora_dstream_* / ora_decompress_continue
 *                             LZ4_decompress_safe_continue cbits/lz4.c:2322-2359
 *     -> decode_block         LZ4_decompress_generic(endOnInputSize, decode_full_block,
 *                             {noDict,usingExtDict})       cbits/lz4.c:1737-2165 (safe loop :1929-2151)
 *
 * Canonical dictionary semantics: consecutive arrays
 * of a stream are NEVER address-adjacent (Haskell arrays are separately
 * allocated), so the compressor is always in external-dictionary mode with the
 * immediately preceding array as the dictionary, and the decoder is always in
 * forceExtDict mode with the immediately preceding OUTPUT as dictionary.  The
 * prefix-mode branches (cbits/lz4.c:1600-1605, :2334-2346) are intentionally
 * not restated; oracle/ref_driver.c refuses adjacent arrays so the real
 * reference never takes them either.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORA_HASH_ENTRIES 4096u          /* LZ4_HASH_SIZE_U32, cbits/lz4.h:578-580 */
#define ORA_MAX_DISTANCE 65535u         /* LZ4_DISTANCE_MAX, cbits/lz4.h:556-558 */
#define ORA_MAX_INPUT    0x7E000000     /* LZ4_MAX_INPUT_SIZE, cbits/lz4.h:170 */
#define ORA_MIN_MATCH    4
#define ORA_LAST_LITERALS 5
#define ORA_MFLIMIT      12
#define ORA_MIN_LENGTH   13             /* MFLIMIT+1, cbits/lz4.c:219-221 */
#define ORA_ACCEL_MAX    65537          /* LZ4_ACCELERATION_MAX, cbits/lz4.c:57 */

typedef struct ora_cstream {
    uint32_t table[ORA_HASH_ENTRIES];   /* stream positions (byU32) */
    uint32_t offset;                    /* currentOffset */
    const uint8_t* dict;                /* previous array */
    uint32_t dict_len;                  /* its length (NOT clamped to 64 KiB) */
} ora_cstream;

typedef struct ora_dstream {
    const uint8_t* prev_out;            /* previous output array */
    size_t prev_len;
} ora_dstream;

int ora_compress_bound(int n)
{
    if ((unsigned)n > (unsigned)ORA_MAX_INPUT) return 0;
    return n + n / 255 + 16;
}

ora_cstream* ora_cstream_create(void) { return (ora_cstream*)calloc(1, sizeof(ora_cstream)); }
void ora_cstream_free(ora_cstream* s) { free(s); }
ora_dstream* ora_dstream_create(void) { return (ora_dstream*)calloc(1, sizeof(ora_dstream)); }
void ora_dstream_free(ora_dstream* s) { free(s); }

/* introspection for tests: the hash table is part of the observable state in
 * linked mode */
const uint32_t* ora_cstream_table(const ora_cstream* s) { return s->table; }
uint32_t ora_cstream_offset(const ora_cstream* s) { return s->offset; }

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

/* 12-bit bucket from the 5 bytes at p (x86-64 little-endian variant of
 * LZ4_hash5, cbits/lz4.c:706-716).  Only p[0..4] influence the result. */
static inline uint32_t hash5(const uint8_t* p)
{
    uint64_t five = (uint64_t)rd32(p) | ((uint64_t)p[4] << 32);
    return (uint32_t)(((five << 24) * 889523592379ULL) >> 52);
}

/* number of equal leading bytes of a[0..] and b[0..], at most `cap`
 * (what LZ4_count returns, cbits/lz4.c:603-626) */
static inline uint32_t common_prefix(const uint8_t* a, const uint8_t* b, uint32_t cap)
{
    uint32_t k = 0;
    while (k < cap && a[k] == b[k]) k++;
    return k;
}

/* length field continuation: v >= 15 already put in the nibble */
static inline uint8_t* put_len_ext(uint8_t* op, uint32_t rest)
{
    while (rest >= 255) { *op++ = 255; rest -= 255; }
    *op++ = (uint8_t)rest;
    return op;
}

typedef struct {
    const uint8_t* src; int32_t n;
    const uint8_t* dict; uint32_t dict_len;
    uint32_t start;              /* stream index of src[0] */
    int dict_small; uint32_t low_index;   /* prefixIdxLimit */
    uint32_t* table;
} enc_t;

/* Resolve a table entry into bytes.  Returns pointer to candidate, sets
 * *floor_out to the lowest address catch-up may reach and *in_dict. */
static inline const uint8_t* locate(const enc_t* e, uint32_t idx, const uint8_t** floor_out, int* in_dict)
{
    if (idx < e->start) {               /* cbits/lz4.c:985-989 */
        *in_dict = 1; *floor_out = e->dict;
        return e->dict + e->dict_len - (e->start - idx);
    }
    *in_dict = 0; *floor_out = e->src;  /* cbits/lz4.c:990-993 */
    return e->src + (idx - e->start);
}

/* One table probe at block position `pos`: read bucket, overwrite it with the
 * current index, and say whether the old entry is an acceptable 4-byte match
 * (cbits/lz4.c:960-1012 and :1159-1189 share this shape). */
static inline int probe(enc_t* e, int32_t pos, const uint8_t** cand, const uint8_t** floor_out,
                        int* in_dict, uint32_t* dist)
{
    uint32_t h = hash5(e->src + pos);
    uint32_t cur = e->start + (uint32_t)pos;
    uint32_t old = e->table[h];
    e->table[h] = cur;
    if (e->dict_small && old < e->low_index) return 0;      /* :1001 / :1187 */
    if (old + ORA_MAX_DISTANCE < cur) return 0;             /* :1003-1006 / :1188 */
    *cand = locate(e, old, floor_out, in_dict);
    if (rd32(*cand) != rd32(e->src + pos)) return 0;        /* :1009 / :1189 */
    *dist = cur - old;
    return 1;
}

static int encode_block(enc_t* e, uint8_t* dst, int cap, int accel)
{
    const uint8_t* const src = e->src;
    const int32_t n = e->n;
    uint8_t* op = dst;
    uint8_t* const oend = dst + cap;
    int32_t anchor = 0;
    const int32_t last_probe = n - ORA_MFLIMIT + 1;     /* mflimitPlusOne as an index, :883 */
    const int32_t match_cap = n - ORA_LAST_LITERALS;    /* matchlimit, :884 */

    if (n >= ORA_MIN_LENGTH) {                          /* :921 */
        int32_t ip;
        e->table[hash5(src)] = e->start;                /* :924 */
        ip = 1;                                         /* :925 */
        for (;;) {
            const uint8_t *cand = NULL, *floor = NULL; int in_dict = 0; uint32_t dist = 0;
            int32_t at = ip, step = 1;
            int32_t tick = accel << 6;                  /* searchMatchNb, :958 */
            int hit = 0;
            /* ---- forward search, :959-1014 ---- */
            for (;;) {
                int32_t next = at + step;
                step = tick++ >> 6;                     /* :967 */
                if (next > last_probe) goto tail;       /* :969 */
                if (probe(e, at, &cand, &floor, &in_dict, &dist)) { hit = 1; break; }
                at = next;
            }
            (void)hit;
            ip = at;
            /* ---- catch up, :1019 ---- */
            while (ip > anchor && cand > floor && src[ip - 1] == cand[-1]) { ip--; cand--; }
            {   /* ---- literals, :1022-1046 ---- */
                uint32_t lit = (uint32_t)(ip - anchor);
                uint8_t* token = op++;
                if (op + lit + (2 + 1 + ORA_LAST_LITERALS) + lit / 255 > oend) return 0;   /* :1024-1027 */
                if (lit >= 15) { *token = 0xF0; op = put_len_ext(op, lit - 15); }
                else *token = (uint8_t)(lit << 4);
                memcpy(op, src + anchor, lit); op += lit;
                for (;;) {   /* ---- _next_match, :1048-1136 ---- */
                    uint32_t mlen;
                    op[0] = (uint8_t)dist; op[1] = (uint8_t)(dist >> 8); op += 2;     /* :1068 */
                    if (in_dict) {                      /* :1078-1090 */
                        uint32_t room_dict = (uint32_t)((e->dict + e->dict_len) - cand);   /* dictEnd - match */
                        int32_t lim = ip + (int32_t)(room_dict < (uint32_t)(match_cap - ip) ? room_dict
                                                                                            : (uint32_t)(match_cap - ip));
                        mlen = common_prefix(src + ip + 4, cand + 4, (uint32_t)(lim - (ip + 4) > 0 ? lim - (ip + 4) : 0));
                        ip += (int32_t)mlen + 4;
                        if (ip == lim) {                /* ran off the dictionary end: continue at block start */
                            uint32_t more = common_prefix(src + lim, src, (uint32_t)(match_cap - lim));
                            mlen += more; ip += (int32_t)more;
                        }
                    } else {                            /* :1092-1094 */
                        mlen = common_prefix(src + ip + 4, cand + 4, (uint32_t)(match_cap - (ip + 4)));
                        ip += (int32_t)mlen + 4;
                    }
                    if (op + (1 + ORA_LAST_LITERALS) + (mlen + 240) / 255 > oend) return 0;   /* :1097-1121 */
                    if (mlen >= 15) { *token += 15; op = put_len_ext(op, mlen - 15); }        /* :1123-1135 */
                    else *token += (uint8_t)mlen;
                    anchor = ip;
                    if (ip >= last_probe) goto tail;    /* :1143 */
                    e->table[hash5(src + ip - 2)] = e->start + (uint32_t)(ip - 2);   /* :1146 */
                    /* immediate re-test at ip, :1159-1196 */
                    if (!probe(e, ip, &cand, &floor, &in_dict, &dist)) break;
                    token = op++; *token = 0;
                }
            }
            ip++;                                       /* :1200 */
        }
    }
tail:
    {   /* ---- last literals, :1204-1231 ---- */
        uint32_t run = (uint32_t)(n - anchor);
        if (op + run + 1 + ((run + 255 - 15) / 255) > oend) return 0;     /* :1207-1217 */
        if (run >= 15) { *op++ = 0xF0; op = put_len_ext(op, run - 15); }
        else *op++ = (uint8_t)(run << 4);
        memcpy(op, src + anchor, run); op += run;
    }
    return (int)(op - dst);
}

/* LZ4_renormDictT, cbits/lz4.c:1545-1562 */
static void renorm(ora_cstream* s, int next)
{
    if (s->offset + (uint32_t)next > 0x80000000u) {
        uint32_t delta = s->offset - 65536u;
        const uint8_t* dict_end = s->dict + s->dict_len;
        for (uint32_t i = 0; i < ORA_HASH_ENTRIES; i++)
            s->table[i] = (s->table[i] < delta) ? 0 : s->table[i] - delta;
        s->offset = 65536u;
        if (s->dict_len > 65536u) s->dict_len = 65536u;
        s->dict = dict_end - s->dict_len;
    }
}

/* LZ4_compress_fast_continue restricted to non-adjacent arrays
 * (cbits/lz4.c:1565-1637).  Returns compressed size, 0 on failure. */
int ora_compress_continue(ora_cstream* s, const uint8_t* src, uint8_t* dst, int n, int cap, int accel)
{
    enc_t e; int r;
    if (n < 0 || (unsigned)n > (unsigned)ORA_MAX_INPUT) return 0;       /* :1262 */
    renorm(s, n);                                                       /* :1576 */
    if (accel < 1) accel = 1;                                           /* :1577 */
    if (accel > ORA_ACCEL_MAX) accel = ORA_ACCEL_MAX;                   /* :1578 */
    if (s->dict_len >= 1 && s->dict_len <= 3) { s->dict_len = 0; s->dict = src; }   /* :1581-1587 */
    if (n == 0) {                                                       /* :1263-1273 */
        if (cap <= 0) return 0;
        dst[0] = 0;
        s->dict = src; s->dict_len = 0;                                 /* :1633-1634 */
        return 1;
    }
    e.src = src; e.n = n; e.dict = s->dict; e.dict_len = s->dict_len;
    e.start = s->offset; e.table = s->table;
    e.dict_small = (s->dict_len < 65536u) && (s->dict_len < s->offset); /* :1627 */
    e.low_index = s->offset - s->dict_len;                              /* :879 */
    s->offset += (uint32_t)n;                                           /* :918 */
    r = encode_block(&e, dst, cap, accel);
    s->dict = src; s->dict_len = (uint32_t)n;                           /* :1633-1634 */
    return r;
}

/* ------------------------------------------------------------------ */
/* Decoder: safe sequence loop semantics (cbits/lz4.c:1929-2151).      */

static int decode_block(const uint8_t* src, int src_len, uint8_t* dst, int cap,
                        const uint8_t* dict, size_t dict_len)
{
    int ip = 0, op = 0;
    const int check_offset = dict_len < 65536;                          /* :1764 */
    if (src == NULL) return -1;
    if (cap == 0) return (src_len == 1 && src[0] == 0) ? 0 : -1;        /* :1781-1785 */
    if (src_len == 0) return -1;                                        /* :1787 */
    for (;;) {
        uint32_t token = src[ip++];
        size_t len = token >> 4;
        size_t dist; long from;
        if (len == 15) {                                                /* :1977-1983, reader :1707-1729 */
            uint32_t s;
            if (ip >= src_len - 15) goto bad;                           /* initial_error */
            do {
                s = src[ip++]; len += s;
                if (ip >= src_len - 15) break;                          /* loop_error: length so far is used */
            } while (s == 255);
        }
        /* end rule, :1991-2047 */
        if ((long)op + (long)len > (long)cap - ORA_MFLIMIT || (long)ip + (long)len > (long)src_len - (2 + 1 + ORA_LAST_LITERALS)) {
            if ((long)ip + (long)len != (long)src_len || (long)op + (long)len > (long)cap) goto bad;
            memmove(dst + op, src + ip, len);
            ip += (int)len; op += (int)len;
            break;
        }
        memcpy(dst + op, src + ip, len); ip += (int)len; op += (int)len;
        dist = (size_t)src[ip] | ((size_t)src[ip + 1] << 8); ip += 2;   /* :2055 */
        len = token & 15;
        if (len == 15) {                                                /* :2062-2067 */
            uint32_t s;
            do {
                s = src[ip++]; len += s;
                if (ip >= src_len - ORA_LAST_LITERALS + 1) goto bad;    /* loop_error -> error */
            } while (s == 255);
        }
        len += ORA_MIN_MATCH;
        from = (long)op - (long)dist;
        if (check_offset && from + (long)dict_len < 0) goto bad;        /* :2073 */
        if ((long)op + (long)len > (long)cap - ORA_LAST_LITERALS) goto bad;   /* :2076-2078, :2139 */
        if (dist == 0) goto bad;   /* format violation; the reference's behaviour here is copy-garbage (:2122-2130) */
        {   /* byte-forward copy, possibly starting in the dictionary tail (:2075-2100, :2103-2150) */
            size_t k;
            for (k = 0; k < len; k++) {
                long f = from + (long)k;
                dst[op + (long)k] = (f < 0) ? dict[(long)dict_len + f] : dst[f];
            }
            op += (int)len;
        }
    }
    return op;
bad:
    return -ip - 1;                                                     /* :2162-2163 */
}

/* LZ4_decompress_safe_continue restricted to non-adjacent outputs
 * (cbits/lz4.c:2322-2359: first call :2327-2333, otherwise :2347-2356). */
int ora_decompress_continue(ora_dstream* s, const uint8_t* src, uint8_t* dst, int src_len, int cap)
{
    int r;
    if (s->prev_len == 0) r = decode_block(src, src_len, dst, cap, NULL, 0);
    else r = decode_block(src, src_len, dst, cap, s->prev_out, s->prev_len);
    if (r <= 0) return r;                                               /* :2331, :2353 */
    s->prev_out = dst; s->prev_len = (size_t)r;
    return r;
}


/* ---------------------------------------------------------------------------
 * XXH32, restated from its published definition (xxHash is bundled with lz4's
 * frame library, lib/xxhash.c, which is NOT part of the reference tree: the
 * reference's frame parser stops at a stub, src/Streamly/Internal/LZ4.hs:602).
 * Pinned by the known answers in tests/test_oracle.py (empty input, "a", "abc",
 * a 39-byte sentence, and the frame descriptors 64 40 -> A7, 64 70 -> B9).
 */
static inline uint32_t xrotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
uint32_t ora_xxh32(const uint8_t* p, size_t len, uint32_t seed)
{
    const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    const uint8_t* const end = p + len;
    uint32_t h;
    if (len >= 16) {
        uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        const uint8_t* const limit = end - 16;
        do {
            v1 = xrotl(v1 + rd32(p) * P2, 13) * P1;
            v2 = xrotl(v2 + rd32(p + 4) * P2, 13) * P1;
            v3 = xrotl(v3 + rd32(p + 8) * P2, 13) * P1;
            v4 = xrotl(v4 + rd32(p + 12) * P2, 13) * P1;
            p += 16;
        } while (p <= limit);
        h = xrotl(v1, 1) + xrotl(v2, 7) + xrotl(v3, 12) + xrotl(v4, 18);
    } else {
        h = seed + P5;
    }
    h += (uint32_t)len;
    while (p + 4 <= end) { h = xrotl(h + rd32(p) * P3, 17) * P4; p += 4; }
    while (p < end) { h = xrotl(h + (*p) * P5, 11) * P1; p++; }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}
