/* CPU PROTOTYPE (development aid; not product, not oracle): a byte-exact LZ4 block compressor whose greedy parse is
 * SPLIT ACROSS WORKERS -- speculate, verify, repair.  It exists to answer, with the reference's bytes as the judge, whether
 * the one-serial-chain-per-block (and per linked stream) limit of csrc/compress.cu (DESIGN.md sections 4.1, 7) can be lifted.
 *
 * The reference's encoder (cbits/lz4.c:851-1240, driven by LZ4_compress_fast_continue :1565-1637; restated with
 * citations in oracle/lz4_oracle.c) reads and overwrites a 4096-entry position table at every probe, so its output is
 * defined by a serial recurrence.  Observation (tools/spec_split_study.c): the recurrence forgets.  A parse started
 * W >= 64 KiB before a split point with an EMPTY table meets the true parse at a match end within a few bytes of the
 * split, and from there on the two differ in a handful of short stretches at most, although their tables are rarely
 * identical.  That is enough for an exact algorithm:
 *
 *   phase 1 (parallel, one worker per unit)   a unit is a SEGMENT of a large independent block, or a BLOCK of a linked
 *       stream.  Worker k parses from W bytes (or w blocks) before its unit with an empty table and keeps, for its unit:
 *       its sequences, a log of its table accesses (position, value read, probe or blind insert), and a snapshot of
 *       its table at the point where the true parse can join (segments: its first match end inside the unit; linked
 *       blocks: the block start, where every parse begins anew).
 *   phase 2 (units in order; the inner loops are data-parallel: 4096-bucket compares, log scans)
 *       the TRUE state (table + position of the last match end) enters unit k from unit k-1.
 *       sync:    advance the true parse one sequence at a time until its match end is also a match end of worker k.
 *       verify:  the speculative parse from there on is exact unless some table READ returns a value that changes a
 *                decision.  Only the FIRST access of a bucket after sync can read a pre-sync value, so: D = buckets whose
 *                true and speculative values differ (and are not both out of reach); scan worker k's log forward; a blind
 *                insert or a probe that decides the same under the true value resolves its bucket (the overwrite that
 *                follows makes both tables equal there); the first probe that decides differently is the divergence.
 *                D empty => everything worker k produced from here on is exact.
 *       accept:  every speculative sequence completed before the divergence is appended; the true table is brought
 *                forward by replaying those log entries (a per-bucket max: parallel).
 *       repair:  from the last accepted match end the true parse runs serially until it syncs again; D is then
 *                updated from the buckets either parse touched in between (no second full compare).
 *   phase 3: the sequence lists are encoded (the existing emitter's job).
 *
 * Soundness: by induction over worker k's accesses after sync -- an access reads either a value written after sync by an
 * access already shown equal in both parses, or the pre-sync value, which is the first-access case that was checked.
 * What a kernel would NOT need (checked here on every linked test stream, stats [17]-[20]): the access log -- it follows
 * from a worker's sequences plus one number per sequence, the catch-up length (reconstruct_and_compare), and the values
 * read at first accesses are the snapshot's entries; the sequential log scan -- first access per bucket is a minimum over
 * access order and every differing bucket is judged independently (verify_block_logfree); the log replay -- a unit that
 * verifies to its end is taken over by a 4096-wide merge of the worker's final table.
 * The statistics returned say how much was left to the serial part.  Driven by tools/specparse_proto.py and
 * tests/test_specparse_proto.py (bytes compared with the oracle's).
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
#define NSTATS 24

typedef struct { int32_t start, anchor_before, end, dist; int32_t log_idx; int32_t floor_hit; int32_t catchup; } seq_t;
typedef struct { int32_t pos, old; int32_t probe; } acc_t;

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t hash5(const uint8_t* p)
{
    uint64_t five = (uint64_t)rd32(p) | ((uint64_t)p[4] << 32);
    return (uint32_t)(((five << 24) * 889523592379ULL) >> 52);
}

typedef struct {
    const uint8_t* src; int accel;      /* src: the whole (contiguous) stream; every position below indexes it */
    /* geometry of the block being parsed: the block, where its parse begins (the block start; later for a speculative
     * mid-block start), the floor of a catch-up inside the block, the dictionary = the previous block (cbits/lz4.c:1607-1636) */
    int32_t blk_lo, blk_hi, begin, floor_blk, dict_lo; int dict_small;
    int32_t table[HASH_ENTRIES];
    int32_t anchor; int started;
    seq_t* seqs; int nseq, cap_seq;
    acc_t* log; int nlog, cap_log; int logging;
    int32_t snap_from; int snapped; int32_t snap[HASH_ENTRIES]; int32_t snap_log_idx, snap_seq_idx;
    int64_t bytes_parsed, accesses; int ended_in_tail;
} parser_t;

/* can a probe at `pos` still use the table value `old` at all? (cbits/lz4.c:1001 / :1187 prefixIdxLimit, :1003-1006 / :1188) */
static inline int reachable(const parser_t* G, int32_t pos, int32_t old)
{
    if (old == EMPTY) return 0;
    if (G->dict_small && old < G->dict_lo) return 0;
    if (old + MAX_DIST < pos) return 0;
    return 1;
}
static inline int accepts(const parser_t* G, int32_t pos, int32_t old)
{
    return reachable(G, pos, old) && rd32(G->src + old) == rd32(G->src + pos);        /* :1009 / :1189 */
}
static inline void log_access(parser_t* P, int32_t pos, int32_t old, int probe)
{
    P->accesses++;
    if (!P->logging) return;
    if (P->nlog == P->cap_log) { P->cap_log = P->cap_log ? 2 * P->cap_log : 4096; P->log = (acc_t*)realloc(P->log, sizeof(acc_t) * (size_t)P->cap_log); }
    P->log[P->nlog].pos = pos; P->log[P->nlog].old = old; P->log[P->nlog].probe = probe; P->nlog++;
}
static inline int probe(parser_t* P, int32_t pos, int32_t* cand)
{
    const uint32_t h = hash5(P->src + pos);
    const int32_t old = P->table[h];
    log_access(P, pos, old, 1);
    P->table[h] = pos;
    if (!accepts(P, pos, old)) return 0;
    *cand = old;
    return 1;
}
static inline void push_seq(parser_t* P, int32_t start, int32_t anchor, int32_t end, int32_t dist, int floor_hit, int32_t catchup)
{
    if (P->nseq == P->cap_seq) { P->cap_seq = P->cap_seq ? 2 * P->cap_seq : 1024; P->seqs = (seq_t*)realloc(P->seqs, sizeof(seq_t) * (size_t)P->cap_seq); }
    seq_t* s = &P->seqs[P->nseq++];
    s->start = start; s->anchor_before = anchor; s->end = end; s->dist = dist; s->log_idx = P->nlog; s->floor_hit = floor_hit; s->catchup = catchup;
}

/* Runs the greedy parse of the current block from the saved state (a match end, or the block's beginning) until the
 * first match end >= stop_at.  returns 0: stopped at a match end (P->anchor), 1: the block's tail was reached (its last
 * literals start at P->anchor). */
static int run(parser_t* P, int32_t stop_at)
{
    const uint8_t* const src = P->src;
    const int32_t hi = P->blk_hi;
    const int32_t last_probe = hi - MFLIMIT + 1, match_cap = hi - LAST_LITERALS;
    int32_t anchor = P->anchor, ip = anchor, cand = 0;
    const int32_t entered_at = anchor;
    int floor_hit = 0; int32_t catchup = 0;
    if (!P->started) {
        if (hi - P->begin < MIN_LENGTH) { P->ended_in_tail = 1; return 1; }     /* :921 */
        P->started = 1;
        log_access(P, P->begin, P->table[hash5(src + P->begin)], 0);           /* :924 */
        P->table[hash5(src + P->begin)] = P->begin;
        ip = P->begin + 1;                                                       /* :925 */
        goto search;
    }
after_match:
    if (ip >= last_probe) { P->anchor = anchor; P->bytes_parsed += hi - entered_at; P->ended_in_tail = 1; return 1; }     /* :1143 */
    {
        const uint32_t h2 = hash5(src + ip - 2);                                 /* :1146 */
        log_access(P, ip - 2, P->table[h2], 0);
        P->table[h2] = ip - 2;
    }
    floor_hit = 0; catchup = 0;
    if (probe(P, ip, &cand)) goto match;                                        /* :1159-1196, no catch-up */
    ip++;                                                                        /* :1200 */
search:
    {
        int32_t at = ip, step = 1, tick = P->accel << 6;                         /* :957-958 */
        for (;;) {
            const int32_t next = at + step;
            step = tick++ >> 6;                                                  /* :967 */
            if (next > last_probe) { P->anchor = anchor; P->bytes_parsed += hi - entered_at; P->ended_in_tail = 1; return 1; }    /* :969 */
            if (probe(P, at, &cand)) break;
            at = next;
        }
        ip = at;
        {   /* catch up, :1019: never below the block start for a candidate inside the block, never below the dictionary's start */
            const int32_t floor = cand >= P->blk_lo ? P->floor_blk : P->dict_lo;
            while (ip > anchor && cand > floor && src[ip - 1] == src[cand - 1]) { ip--; cand--; }
            catchup = at - ip;
            /* the true parse's floor is the block start: a speculative catch-up that stopped AT its own origin may be short */
            floor_hit = (cand >= P->blk_lo && P->floor_blk > P->blk_lo && ip > anchor && cand == floor && src[ip - 1] == src[cand - 1]);
        }
    }
match:
    {
        /* match length: bytes of the stream are contiguous here, so "runs off the dictionary's end and continues at the
         * block's start" (:1078-1090) is the same comparison as inside the block (:1092-1094) */
        int32_t k = 0;
        const int32_t cap = match_cap - (ip + 4);
        while (k < cap && src[ip + 4 + k] == src[cand + 4 + k]) k++;
        push_seq(P, ip, anchor, ip + 4 + k, ip - cand, floor_hit, catchup);
        ip += 4 + k;
        anchor = ip;
        if (!P->snapped && anchor >= P->snap_from) {
            memcpy(P->snap, P->table, sizeof P->snap); P->snapped = 1; P->snap_log_idx = P->nlog; P->snap_seq_idx = P->nseq - 1;
        }
        if (anchor >= stop_at) { P->anchor = anchor; P->bytes_parsed += anchor - entered_at; return 0; }
        goto after_match;
    }
}

static parser_t* parser_new(const uint8_t* src, int accel, int32_t fill)
{
    parser_t* P = (parser_t*)calloc(1, sizeof(parser_t));
    P->src = src; P->accel = accel < 1 ? 1 : accel > 65537 ? 65537 : accel;        /* :1577-1578 */
    for (int i = 0; i < HASH_ENTRIES; i++) P->table[i] = fill;
    P->snap_from = INT32_MAX;
    return P;
}
/* the next block of the stream: [lo, hi) with the previous block [prev_lo, lo) as dictionary (LZ4_compress_fast_continue) */
static void parser_set_block(parser_t* P, int32_t lo, int32_t hi, int32_t prev_lo)
{
    int32_t dict_len = lo - prev_lo;
    if (dict_len >= 1 && dict_len <= 3) dict_len = 0;                               /* :1581-1587 */
    P->blk_lo = lo; P->blk_hi = hi; P->begin = lo; P->floor_blk = lo;
    P->dict_lo = lo - dict_len;
    P->dict_small = (dict_len < 65536) && (dict_len < lo);                          /* :1627 (currentOffset == lo) */
    P->anchor = lo; P->started = 0; P->ended_in_tail = 0;
}
static void parser_free(parser_t* P) { if (P) { free(P->seqs); free(P->log); free(P); } }


/* The access log of a whole-block parse is implied by its OUTPUT plus one number per sequence (the catch-up length, i.e.
 * where the accepting probe stood): blind insert at the block's beginning, then per sequence the insert at (previous
 * end - 2), the re-test at the previous end and, unless that one produced the sequence, the probes of the search run in
 * closed form (cbits/lz4.c:957-969).  The VALUES read are not needed: the first access of a bucket after the sync point
 * reads the snapshot's entry by definition.  Returns the number of positions where the reconstruction differs from the
 * logged accesses (0 = the log need not be stored). */
static int64_t reconstruct_and_compare(const parser_t* P)
{
    const int32_t last_probe = P->blk_hi - MFLIMIT + 1;
    int64_t bad = 0; int L = 0;
#define EXPECT(p, pr) do { if (L >= P->nlog || P->log[L].pos != (p) || P->log[L].probe != (pr)) bad++; L++; } while (0)
    if (P->blk_hi - P->begin < MIN_LENGTH) return P->nlog ? 1 : 0;
    EXPECT(P->begin, 0);
    int32_t e = P->begin;               /* end of the previous sequence (the block's beginning before the first) */
    for (int i = 0; i <= P->nseq; i++) {
        const int have = i < P->nseq;
        int from_retest = 0;
        if (i > 0) {
            if (e >= last_probe) break;                                  /* :1143: the tail begins, nothing more is touched */
            EXPECT(e - 2, 0);
            EXPECT(e, 1);
            from_retest = have && P->seqs[i].start == e && P->seqs[i].anchor_before == e && P->seqs[i].catchup == 0;
        }
        if (!from_retest) {
            const int32_t target = have ? P->seqs[i].start + P->seqs[i].catchup : INT32_MAX;
            int32_t at = e + 1, step = 1, tick = P->accel << 6;
            for (;;) {
                const int32_t next = at + step;
                step = tick++ >> 6;
                if (next > last_probe) { if (have) bad++; break; }      /* the run ends in the tail */
                EXPECT(at, 1);
                if (at == target) break;
                if (at > target) { bad++; break; }
                at = next;
            }
        }
        if (have) e = P->seqs[i].end;
    }
#undef EXPECT
    if (L != P->nlog) bad++;
    return bad;
}

/* ---- encoder of a sequence list (what emitter_main does on the device) ---- */
static uint8_t* put_ext(uint8_t* op, uint32_t rest) { while (rest >= 255) { *op++ = 255; rest -= 255; } *op++ = (uint8_t)rest; return op; }
static int encode(const uint8_t* src, int32_t hi, const seq_t* s, int ns, int32_t tail_from, uint8_t* dst)
{
    uint8_t* op = dst;
    for (int i = 0; i < ns; i++) {
        const uint32_t lit = (uint32_t)(s[i].start - s[i].anchor_before), m = (uint32_t)(s[i].end - s[i].start - 4);
        uint8_t* token = op++;
        if (lit >= 15) { *token = 0xF0; op = put_ext(op, lit - 15); } else *token = (uint8_t)(lit << 4);
        memcpy(op, src + s[i].anchor_before, lit); op += lit;
        op[0] = (uint8_t)s[i].dist; op[1] = (uint8_t)(s[i].dist >> 8); op += 2;
        if (m >= 15) { *token += 15; op = put_ext(op, m - 15); } else *token += (uint8_t)m;
    }
    {
        const uint32_t run_len = (uint32_t)(hi - tail_from);
        if (run_len >= 15) { *op++ = 0xF0; op = put_ext(op, run_len - 15); } else *op++ = (uint8_t)(run_len << 4);
        memcpy(op, src + tail_from, run_len); op += run_len;
    }
    return (int)(op - dst);
}

/* stats (sums over the call):
 * [0] units (segments / blocks)         [1] bytes parsed by all workers (warm-ups included)
 * [2] longest single worker (bytes)     [3] bytes parsed serially in phase 2 (sync steps + repairs)
 * [4] sequences parsed serially         [5] full 4096-bucket compares      [6] divergences found
 * [7] log entries scanned by verifications and difference updates          [8] log entries replayed into tables
 * [9] sequences accepted from workers   [10] total sequences               [11] syncs that needed no serial step
 * [12] table accesses of all workers    [13] of the busiest worker         [14] of the serial part of phase 2
 * [15] of a plain serial parse (the baseline the critical path is compared with)     [16] verifications
 * [20] units (or rests of units) taken over by a 4096-wide merge of the worker's final table instead of a log replay
 * [18] (linked) block-level checks also run in their log-free, kernel-shaped form   [19] of those, results that differ from the log scan's
 * [17] (linked) positions where the access log RECONSTRUCTED from a worker's sequences + catch-up lengths differs from the logged one */

static inline int bucket_differs(const parser_t* T, const int32_t* Stab, uint32_t b)
{
    const int32_t a = T->table[b], c = Stab[b];
    return !(a == c || (!reachable(T, T->anchor, a) && !reachable(T, T->anchor, c)));
}


/* The block-level check in the shape a kernel would run it, without any log: every access position follows from the
 * worker's sequences (+ catch-up lengths); "first access per bucket" is a minimum over access order (atomicMin on the
 * device); a differing bucket whose first access is a probe that decides differently under the true value marks a
 * divergence, the earliest one wins.  `Ttab` = true table, `snap` = the worker's table, both at the block's beginning.
 * returns the order index of the earliest deciding-differently access, -1 if the whole block is exact. */
static int verify_block_logfree(const parser_t* S, const parser_t* G, const int32_t* Ttab)
{
    const int32_t last_probe = S->blk_hi - MFLIMIT + 1;
    int32_t* pos = (int32_t*)malloc(sizeof(int32_t) * (size_t)(S->nlog + 8));
    uint8_t* isprobe = (uint8_t*)malloc((size_t)(S->nlog + 8));
    int n = 0;
    /* (1) access positions in order: embarrassingly parallel over sequences on the device, closed-form inside a run */
    if (S->blk_hi - S->begin >= MIN_LENGTH) {
        int32_t e = S->begin;
        pos[n] = S->begin; isprobe[n++] = 0;
        for (int i = 0; i <= S->nseq; i++) {
            const int have = i < S->nseq;
            int from_retest = 0;
            if (i > 0) {
                if (e >= last_probe) break;
                pos[n] = e - 2; isprobe[n++] = 0;
                pos[n] = e; isprobe[n++] = 1;
                from_retest = have && S->seqs[i].start == e && S->seqs[i].anchor_before == e && S->seqs[i].catchup == 0;
            }
            if (!from_retest) {
                const int32_t target = have ? S->seqs[i].start + S->seqs[i].catchup : INT32_MAX;
                int32_t at = e + 1, step = 1, tick = S->accel << 6;
                for (;;) {
                    const int32_t next = at + step;
                    step = tick++ >> 6;
                    if (next > last_probe) break;
                    pos[n] = at; isprobe[n++] = 1;
                    if (at >= target) break;
                    at = next;
                }
            }
            if (have) e = S->seqs[i].end;
        }
    }
    /* (2) first access per bucket (atomicMin of the order index) */
    int32_t first[HASH_ENTRIES];
    for (int b = 0; b < HASH_ENTRIES; b++) first[b] = INT32_MAX;
    for (int i = 0; i < n; i++) { const uint32_t b = hash5(S->src + pos[i]); if (i < first[b]) first[b] = i; }
    /* (3) every differing bucket, independently */
    int bad = INT32_MAX;
    for (uint32_t b = 0; b < HASH_ENTRIES; b++) {
        const int32_t a = Ttab[b], c = S->snap[b];
        if (a == c || (!reachable(G, G->blk_lo, a) && !reachable(G, G->blk_lo, c))) continue;
        const int i = first[b];
        if (i == INT32_MAX || !isprobe[i]) continue;
        const int da = accepts(G, pos[i], a), dc = accepts(G, pos[i], c);
        if ((da != dc || (da && a != c)) && i < bad) bad = i;
    }
    free(pos); free(isprobe);
    return bad == INT32_MAX ? -1 : bad;
}

/* Phase 2 for one unit: the true parser T stands at a match end (or at the beginning of its block: T->started == 0)
 * inside the unit of worker S; both have the same block geometry.  Runs until T->anchor >= to or the block's tail.
 * `j0`: index of the first worker sequence that may serve as a sync point; `at_block_start`: the sync is the block's
 * beginning itself (linked blocks), the worker's snapshot was taken there.  returns 1 if the tail was reached. */
static int merge_unit(parser_t* T, parser_t* S, int32_t to, int j0, int at_block_start, int64_t* stats)
{
    const uint8_t* const src = T->src;
    int32_t Stab[HASH_ENTRIES];
    uint8_t differs[HASH_ENTRIES];
    int stab_valid = 0, stab_log = 0, have_diff = 0, unresolved = 0, spec_applied = 0;
    int j = j0, tail = 0, serial_since_sync_try = 0, first = at_block_start;
    T->nlog = 0;
    while (!tail && T->anchor < to) {
        int L0;
        if (first) {
            L0 = 0;                         /* the block's beginning: every parse starts anew here, no sequence to match */
        } else {
            /* sync: is the true match end also a match end of the worker (at or after its snapshot)? */
            while (j < S->nseq && S->seqs[j].end < T->anchor) j++;
            if (!(j < S->nseq && S->seqs[j].end == T->anchor)) {
                if (j >= S->nseq) { tail = run(T, to); continue; }      /* the worker has nothing further */
                tail = run(T, T->anchor + 1);                           /* one sequence of the true parse */
                serial_since_sync_try++;
                continue;
            }
            L0 = S->seqs[j].log_idx;
        }
        if (!serial_since_sync_try) stats[11]++;
        serial_since_sync_try = 0;
        /* speculative table at this sync point: the snapshot brought forward by the worker's own log */
        if (!stab_valid) { memcpy(Stab, S->snap, sizeof Stab); stab_log = S->snap_log_idx; stab_valid = 1; }
        if (L0 > stab_log) stats[8] += L0 - stab_log;
        for (; stab_log < L0; stab_log++) Stab[hash5(src + S->log[stab_log].pos)] = S->log[stab_log].pos;
        if (!have_diff) {                   /* first sync in this unit: compare all 4096 buckets (data-parallel) */
            unresolved = 0;
            for (uint32_t b = 0; b < HASH_ENTRIES; b++) { differs[b] = (uint8_t)bucket_differs(T, Stab, b); unresolved += differs[b]; }
            have_diff = 1;
            stats[5]++;
        } else {                            /* later syncs: only buckets either parse touched since the tables were last reconciled */
            for (int pass = 0; pass < 2; pass++) {
                const acc_t* lg = pass ? T->log : S->log;
                const int x0 = pass ? 0 : spec_applied, x1 = pass ? T->nlog : L0;
                for (int x = x0; x < x1; x++) {
                    const uint32_t b = hash5(src + lg[x].pos);
                    const uint8_t d = (uint8_t)bucket_differs(T, Stab, b);
                    unresolved += (int)d - (int)differs[b]; differs[b] = d;
                }
                if (x1 > x0) stats[7] += x1 - x0;
            }
        }
        T->nlog = 0;
        /* verify: scan the worker's log until every differing bucket is resolved or one decides differently */
        stats[16]++;
        int bad = -1, L = L0;
        for (; L < S->nlog && unresolved > 0; L++) {
            const acc_t* e = &S->log[L];
            const uint32_t b = hash5(src + e->pos);
            if (!differs[b]) continue;
            differs[b] = 0; unresolved--;
            if (e->probe) {
                const int da = accepts(T, e->pos, T->table[b]), dc = accepts(T, e->pos, e->old);
                if (da != dc || (da && T->table[b] != e->old)) { bad = L; break; }
            }
        }
        stats[7] += L - L0;
        if (first && S->logging) {           /* linked blocks: the kernel-shaped, log-free check must find the same divergence */
            const int lf = verify_block_logfree(S, T, T->table);
            stats[18]++;
            if (lf != bad) stats[19]++;
        }
        /* accept the worker's sequences completed before the divergence (all of them if there is none) */
        const int j_first = first ? 0 : j + 1;
        int j_last = j_first - 1;
        for (int q = j_first; q < S->nseq; q++) {
            if (bad >= 0 && S->seqs[q].log_idx > bad) break;
            if (S->seqs[q].floor_hit) { bad = bad < 0 ? S->seqs[q].log_idx : bad; break; }
            j_last = q;
        }
        spec_applied = L0;
        if (bad < 0) {
            /* everything the worker produced from this sync point on is exact: it ran to its first match end >= `to`, or into
             * the block's tail.  No replay is needed then: every access after the sync point has a position >= (sync - 2)
             * (the insert at :1146 is the lowest; everything before the sync point lies at least 4 bytes below it), so the
             * worker's FINAL table tells which buckets it touched since, and with what: a 4096-wide merge. */
            const int32_t threshold = first ? T->blk_lo : T->anchor - 2;
            for (int q = j_first; q < S->nseq; q++) push_seq(T, S->seqs[q].start, S->seqs[q].anchor_before, S->seqs[q].end, S->seqs[q].dist, 0, S->seqs[q].catchup);
            stats[9] += S->nseq - j_first;
            for (int b = 0; b < HASH_ENTRIES; b++) if (S->table[b] != EMPTY && S->table[b] >= threshold) T->table[b] = S->table[b];
            stats[20]++;
            if (S->nseq > j_first) { T->anchor = S->seqs[S->nseq - 1].end; T->started = 1; }
            if (S->ended_in_tail) { tail = 1; T->started = 1; T->ended_in_tail = 1; }
            break;
        }
        if (j_last >= j_first) {
            for (int q = j_first; q <= j_last; q++) push_seq(T, S->seqs[q].start, S->seqs[q].anchor_before, S->seqs[q].end, S->seqs[q].dist, 0, S->seqs[q].catchup);
            stats[9] += j_last - j_first + 1;
            const int L1 = S->seqs[j_last].log_idx;
            for (int x = L0; x < L1; x++) T->table[hash5(src + S->log[x].pos)] = S->log[x].pos;      /* per-bucket max: parallel */
            stats[8] += L1 - L0;
            T->anchor = S->seqs[j_last].end;
            T->started = 1;
            spec_applied = L1;
            j = j_last;
            first = 0;
        }
        stats[6]++;
        tail = run(T, T->anchor + 1);       /* repair: the true parse takes the diverging step itself */
        first = 0;
        serial_since_sync_try = 1;
    }
    return tail;
}

/* ---- one large independent block, split into segments ---- */
int specparse_compress(const uint8_t* src, int32_t n, int accel, int32_t seg, int32_t warm, uint8_t* dst, int64_t* stats)
{
    int K = seg > 0 ? (int)((n + seg - 1) / seg) : 1;
    if (K < 1) K = 1;
    parser_t** W = (parser_t**)calloc((size_t)K, sizeof(parser_t*));
    for (int i = 0; i < NSTATS; i++) stats[i] = 0;
    {   /* baseline: the plain serial parse */
        parser_t* B = parser_new(src, accel, 0);
        parser_set_block(B, 0, n, 0);
        run(B, n + 1);
        stats[15] = B->accesses;
        parser_free(B);
    }
    stats[0] = K;

    /* ---- phase 1: speculative workers (independent of each other: this loop is the parallel part) ---- */
    for (int k = 1; k < K; k++) {
        const int32_t from = (int32_t)((int64_t)k * seg), to = (k + 1 < K) ? (int32_t)((int64_t)(k + 1) * seg) : n;
        const int32_t origin = from - warm > 0 ? from - warm : 0;
        parser_t* P = W[k] = parser_new(src, accel, origin == 0 ? 0 : EMPTY);
        parser_set_block(P, 0, n, 0);
        P->begin = origin; P->floor_blk = origin; P->anchor = origin;
        P->snap_from = from;
        /* warm-up [origin, from) without a log, then the unit with one: the snapshot is taken at the first match end >= from
         * and the merge only reads log entries from the snapshot on */
        if (from - 1 > origin && run(P, from - 1)) { /* ran into the tail during the warm-up: nothing to offer */ }
        P->logging = 1;
        if (!P->ended_in_tail) run(P, to);
        stats[1] += P->bytes_parsed; if (P->bytes_parsed > stats[2]) stats[2] = P->bytes_parsed;
        stats[12] += P->accesses; if (P->accesses > stats[13]) stats[13] = P->accesses;
    }

    /* ---- phase 2: the true parse, segment by segment ---- */
    parser_t* T = parser_new(src, accel, 0);        /* zero table = a fresh LZ4_stream_t */
    parser_set_block(T, 0, n, 0);
    int tail = run(T, K > 1 ? seg : n + 1);         /* segment 0 is the true parse by construction */
    stats[1] += T->bytes_parsed; if (T->bytes_parsed > stats[2]) stats[2] = T->bytes_parsed;
    stats[12] += T->accesses; if (T->accesses > stats[13]) stats[13] = T->accesses;
    T->bytes_parsed = 0; T->accesses = 0;
    T->logging = 1;                                 /* from here on the true parse notes which buckets its serial steps touch */
    const int seg0_seqs = T->nseq;
    for (int k = 1; k < K && !tail; k++) {
        const int32_t to = (k + 1 < K) ? (int32_t)((int64_t)(k + 1) * seg) : n;
        parser_t* S = W[k];
        tail = merge_unit(T, S, to, S->snapped ? S->snap_seq_idx : S->nseq, 0, stats);
    }
    if (!tail) tail = run(T, n + 1);
    stats[3] = T->bytes_parsed;
    stats[14] = T->accesses;
    stats[10] = T->nseq;
    stats[4] = T->nseq - stats[9] - seg0_seqs;
    const int out = encode(src, n, T->seqs, T->nseq, T->anchor, dst);
    for (int k = 1; k < K; k++) parser_free(W[k]);
    free(W);
    parser_free(T);
    return out;
}

/* ---- a LINKED stream (one LZ4_stream_t, dictionary = previous block: the reference's own mode) ----
 * src: the stream's blocks back to back (positions = the reference's stream indices as long as the stream stays below
 * 2 GiB: no LZ4_renormDictT here); block i = [off[i], off[i+1]).  A worker per block, warmed up over the `warm_blocks`
 * blocks before it.  dst: compressBound-sized slot per block at dst_off[i]; out_len[i] = compressed size. */
int specparse_compress_linked(const uint8_t* src, const int32_t* off, int nblocks, int accel, int warm_blocks,
                              uint8_t* dst, const int64_t* dst_off, int32_t* out_len, int64_t* stats)
{
    parser_t** W = (parser_t**)calloc((size_t)nblocks, sizeof(parser_t*));
    for (int i = 0; i < NSTATS; i++) stats[i] = 0;
    stats[0] = nblocks;
    {   /* baseline: the plain serial parse of the stream */
        parser_t* B = parser_new(src, accel, 0);
        for (int i = 0; i < nblocks; i++) { parser_set_block(B, off[i], off[i + 1], i ? off[i - 1] : 0); run(B, off[i + 1] + 1); }
        stats[15] = B->accesses;
        parser_free(B);
    }
    /* ---- phase 1 ---- */
    for (int k = 1; k < nblocks; k++) {
        const int first = k - warm_blocks > 0 ? k - warm_blocks : 0;
        parser_t* P = W[k] = parser_new(src, accel, first == 0 ? 0 : EMPTY);
        for (int i = first; i < k; i++) { parser_set_block(P, off[i], off[i + 1], i ? off[i - 1] : 0); run(P, off[i + 1] + 1); }
        P->nseq = 0;
        parser_set_block(P, off[k], off[k + 1], off[k - 1]);
        memcpy(P->snap, P->table, sizeof P->snap); P->snapped = 1; P->snap_log_idx = 0; P->snap_seq_idx = 0;
        P->logging = 1;
        run(P, off[k + 1] + 1);
        stats[17] += reconstruct_and_compare(P);
        stats[1] += P->bytes_parsed; if (P->bytes_parsed > stats[2]) stats[2] = P->bytes_parsed;
        stats[12] += P->accesses; if (P->accesses > stats[13]) stats[13] = P->accesses;
    }
    /* ---- phase 2 ---- */
    parser_t* T = parser_new(src, accel, 0);
    int64_t serial_seqs = 0;
    for (int k = 0; k < nblocks; k++) {
        parser_set_block(T, off[k], off[k + 1], k ? off[k - 1] : 0);
        T->nseq = 0;
        if (k == 0) {
            run(T, off[1] + 1);                     /* block 0 is the true parse by construction */
            stats[1] += T->bytes_parsed; if (T->bytes_parsed > stats[2]) stats[2] = T->bytes_parsed;
            stats[12] += T->accesses; if (T->accesses > stats[13]) stats[13] = T->accesses;
            T->bytes_parsed = 0; T->accesses = 0;
            T->logging = 1;
        } else {
            const int64_t acc0 = stats[9];
            int tail = merge_unit(T, W[k], off[k + 1], 0, 1, stats);
            if (!tail) run(T, off[k + 1] + 1);
            serial_seqs += T->nseq - (stats[9] - acc0);
        }
        stats[10] += T->nseq;
        out_len[k] = encode(src, off[k + 1], T->seqs, T->nseq, T->anchor, dst + dst_off[k]);
    }
    stats[3] = T->bytes_parsed;
    stats[14] = T->accesses;
    stats[4] = serial_seqs;
    for (int k = 1; k < nblocks; k++) parser_free(W[k]);
    free(W);
    parser_free(T);
    return 0;
}
