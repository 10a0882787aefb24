/* CPU-only study (development aid, not product, not oracle): what does a SPECULATIVE mid-block start of LZ4's greedy
 * parse cost when it is judged by its OUTPUT rather than by its table state?
 *
 * The reference's encoder (cbits/lz4.c:851-1240, restated with citations in oracle/lz4_oracle.c) is one serial chain per
 * block: every probe reads and overwrites the position table.  DESIGN.md section 7 records that a parse started W bytes
 * before a split point with an empty table almost never reaches the TRUE table state.  This tool asks the finer question:
 *   - where does the speculative parse first share a sequence boundary with the true parse (sync),
 *   - how many table buckets differ there in a way that can still matter (live differences),
 *   - how many of the true parse's sequences after sync are NOT produced by the speculative parse (damage), in how many
 *     separate runs, and how long the parses stay apart per run,
 *   - and whether a cheap check predicts the damage: for every live-different bucket take its FIRST access by the
 *     speculative parse after sync and ask whether the true value would have led to the same accept/reject decision
 *     (a bucket whose first access decides the same is healed by the overwrite that follows).
 * Independent blocks, no dictionary (config 2 / few large blocks).  Built and driven by tools/spec_split_study.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HASH_ENTRIES 4096
#define MAX_DIST 65535
#define MFLIMIT 12
#define LAST_LITERALS 5
#define MIN_LENGTH 13
#define EMPTY (-1)

typedef struct { int32_t start, anchor_before, end, dist; } seq_t;   /* match [start, end), literals [anchor_before, start) */

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t hash5(const uint8_t* p)
{
    uint64_t five = (uint64_t)rd32(p) | ((uint64_t)p[4] << 32);
    return (uint32_t)(((five << 24) * 889523592379ULL) >> 52);
}

typedef struct {
    const uint8_t* src; int32_t n; int32_t origin;     /* the parse treats src[origin] as its first byte */
    int32_t table[HASH_ENTRIES];
    /* first access of every bucket at or after `watch_from` (set before parsing): position and the value read */
    int32_t watch_from;
    int32_t first_pos[HASH_ENTRIES], first_old[HASH_ENTRIES];
    uint8_t first_is_probe[HASH_ENTRIES];              /* 0: plain insert (ip-2), value read does not matter */
    /* snapshot of the table when the anchor first equals snap_at (right after a match ended there) */
    int32_t snap_at; int snapped; int32_t snap[HASH_ENTRIES];
} parser_t;

static inline void note(parser_t* P, uint32_t h, int32_t pos, int32_t old, int is_probe)
{
    if (pos >= P->watch_from && P->first_pos[h] == EMPTY) { P->first_pos[h] = pos; P->first_old[h] = old; P->first_is_probe[h] = (uint8_t)is_probe; }
}

/* decision of a probe at `pos` if the bucket held `old` */
static inline int accepts(const uint8_t* src, int32_t pos, int32_t old)
{
    if (old == EMPTY) return 0;
    if (old + MAX_DIST < pos) return 0;
    return rd32(src + old) == rd32(src + pos);
}

static inline int probe(parser_t* P, int32_t pos, int32_t* cand)
{
    const uint32_t h = hash5(P->src + pos);
    const int32_t old = P->table[h];
    note(P, h, pos, old, 1);
    P->table[h] = pos;
    if (!accepts(P->src, pos, old)) return 0;
    *cand = old;
    return 1;
}

/* greedy parse of src[origin, n) as the reference does it (positions are block positions); returns the number of sequences */
static int parse(parser_t* P, int accel, seq_t* out, int max_out)
{
    const uint8_t* const src = P->src;
    const int32_t n = P->n;
    const int32_t last_probe = n - MFLIMIT + 1, match_cap = n - LAST_LITERALS;
    int32_t anchor = P->origin, ip;
    int cnt = 0;
    if (accel < 1) accel = 1;
    if (accel > 65537) accel = 65537;
    if (n - P->origin < MIN_LENGTH) return 0;
    P->table[hash5(src + P->origin)] = P->origin;
    ip = P->origin + 1;
    for (;;) {
        int32_t cand = 0, at = ip, step = 1, tick = accel << 6;
        for (;;) {
            const int32_t next = at + step;
            step = tick++ >> 6;
            if (next > last_probe) return cnt;
            if (probe(P, at, &cand)) break;
            at = next;
        }
        ip = at;
        while (ip > anchor && cand > P->origin && src[ip - 1] == src[cand - 1]) { ip--; cand--; }
        for (;;) {
            int32_t k = 0;
            const int32_t cap = match_cap - (ip + 4);
            while (k < cap && src[ip + 4 + k] == src[cand + 4 + k]) k++;
            if (cnt < max_out) { out[cnt].start = ip; out[cnt].anchor_before = anchor; out[cnt].end = ip + 4 + k; out[cnt].dist = ip - cand; }
            cnt++;
            ip += 4 + k;
            anchor = ip;
            if (!P->snapped && anchor >= P->snap_at && P->snap_at >= 0 && anchor == P->snap_at) { memcpy(P->snap, P->table, sizeof P->snap); P->snapped = 1; }
            if (ip >= last_probe) return cnt;
            {
                const uint32_t h2 = hash5(src + ip - 2);
                note(P, h2, ip - 2, P->table[h2], 0);
                P->table[h2] = ip - 2;
            }
            if (!probe(P, ip, &cand)) break;
        }
        ip++;
    }
}

static void parser_init(parser_t* P, const uint8_t* src, int32_t n, int32_t origin, int32_t fill, int32_t watch_from, int32_t snap_at)
{
    P->src = src; P->n = n; P->origin = origin;
    for (int i = 0; i < HASH_ENTRIES; i++) { P->table[i] = fill; P->first_pos[i] = EMPTY; P->first_old[i] = EMPTY; P->first_is_probe[i] = 0; }
    P->watch_from = watch_from; P->snap_at = snap_at; P->snapped = 0;
}

/* results:
 * r[0] sync position (-1: none)          r[1] true sequences after sync        r[2] of those, not produced by the speculative parse
 * r[3] runs of consecutive missing ones  r[4] bytes covered by the longest run r[5] live-different buckets at sync
 * r[6] of those, buckets whose first access by the speculative parse decides differently under the true value
 * r[7] position of the earliest such access (-1: none)   r[8] position of the first missing true sequence (-1: none)
 * r[9] bytes from split point to sync   r[10] bytes covered by all runs together */
int spec_split_study(const uint8_t* src, int32_t n, int accel, int32_t split, int32_t warm, int64_t* r)
{
    const int max_seq = n / 4 + 16;
    seq_t* T = (seq_t*)malloc(sizeof(seq_t) * (size_t)max_seq);
    seq_t* S = (seq_t*)malloc(sizeof(seq_t) * (size_t)max_seq);
    parser_t* pt = (parser_t*)malloc(sizeof(parser_t));
    parser_t* ps = (parser_t*)malloc(sizeof(parser_t));
    int32_t origin = split - warm;
    int nt, nsq, i, j;
    if (origin < 0) origin = 0;
    for (i = 0; i < 11; i++) r[i] = -1;

    /* pass 1: both parses, find the first shared match end >= split */
    parser_init(pt, src, n, 0, 0, n, -1);             /* a fresh reference table is all zeros = "position 0" */
    nt = parse(pt, accel, T, max_seq);
    parser_init(ps, src, n, origin, EMPTY, n, -1);
    nsq = parse(ps, accel, S, max_seq);
    int32_t sync = -1; int it = 0, is = 0;
    for (i = 0, j = 0; i < nt && j < nsq;) {
        if (T[i].end < split) { i++; continue; }
        if (S[j].end < split) { j++; continue; }
        if (T[i].end == S[j].end) { sync = T[i].end; it = i + 1; is = j + 1; break; }
        if (T[i].end < S[j].end) i++; else j++;
    }
    r[0] = sync;
    if (sync < 0) { free(T); free(S); free(pt); free(ps); return 0; }
    r[9] = sync - split;

    /* pass 2: tables at sync, first accesses of the speculative parse after sync */
    parser_init(pt, src, n, 0, 0, n, sync);
    parse(pt, accel, T, max_seq);
    parser_init(ps, src, n, origin, EMPTY, sync - 2, sync);      /* the ip-2 insert right after sync counts as an access */
    parse(ps, accel, S, max_seq);
    {
        int64_t live = 0, bad = 0; int32_t earliest = -1;
        for (i = 0; i < HASH_ENTRIES; i++) {
            const int32_t a = pt->snap[i], b = ps->snap[i];
            const int dead_a = (a + MAX_DIST < sync), dead_b = (b == EMPTY) || (b + MAX_DIST < sync);
            if (a == b || (dead_a && dead_b)) continue;
            live++;
            if (ps->first_pos[i] == EMPTY || !ps->first_is_probe[i]) continue;        /* never probed again, or overwritten blindly first */
            {
                const int32_t pos = ps->first_pos[i];
                const int da = accepts(src, pos, a), db = accepts(src, pos, ps->first_old[i]);
                if (da != db || (da && a != ps->first_old[i])) { bad++; if (earliest < 0 || pos < earliest) earliest = pos; }
            }
        }
        r[5] = live; r[6] = bad; r[7] = earliest;
    }

    /* damage: true sequences after sync that the speculative parse does not contain */
    {
        int64_t after = 0, missing = 0, runs = 0, longest = 0, covered = 0; int32_t first_missing = -1;
        int in_run = 0; int32_t run_from = 0;
        j = is;
        for (i = it; i < nt; i++) {
            after++;
            while (j < nsq && S[j].start < T[i].start) j++;
            const int same = j < nsq && S[j].start == T[i].start && S[j].end == T[i].end && S[j].dist == T[i].dist && S[j].anchor_before == T[i].anchor_before;
            if (!same) {
                missing++;
                if (first_missing < 0) first_missing = T[i].start;
                if (!in_run) { in_run = 1; runs++; run_from = T[i].anchor_before; }
            } else if (in_run) {
                in_run = 0;
                if (T[i].anchor_before - run_from > longest) longest = T[i].anchor_before - run_from;
                covered += T[i].anchor_before - run_from;
            }
        }
        if (in_run) { if (n - run_from > longest) longest = n - run_from; covered += n - run_from; }
        /* sequences only the speculative parse has count as damage too */
        if ((int64_t)(nsq - is) > after - missing) {
            const int64_t extra = (int64_t)(nsq - is) - (after - missing);
            if (missing == 0) { runs = 1; first_missing = S[is].start; }
            missing += extra;
        }
        r[1] = after; r[2] = missing; r[3] = runs; r[4] = longest; r[8] = first_missing; r[10] = covered;
    }
    free(T); free(S); free(pt); free(ps);
    return 0;
}

/* the true parse alone: (literal length, match length, distance) triples, for validation against the oracle's bytes */
int spec_true_parse(const uint8_t* src, int32_t n, int accel, int32_t* out, int max_seq)
{
    seq_t* T = (seq_t*)malloc(sizeof(seq_t) * (size_t)max_seq);
    parser_t* pt = (parser_t*)malloc(sizeof(parser_t));
    int nt, i;
    parser_init(pt, src, n, 0, 0, n, -1);
    nt = parse(pt, accel, T, max_seq);
    for (i = 0; i < nt && i < max_seq; i++) { out[3 * i] = T[i].start - T[i].anchor_before; out[3 * i + 1] = T[i].end - T[i].start; out[3 * i + 2] = T[i].dist; }
    free(T); free(pt);
    return nt;
}
