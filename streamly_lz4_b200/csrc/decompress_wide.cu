// decompress_wide.cu -- LZ4 decoder for FEW streams (linked streams, a handful of large blocks): one stream per SM.
//
// Why a second kernel.  The narrow kernel (decompress.cu) gives a stream one parser warp and one copier warp, and a
// linked stream is one chain (block N's matches may read block N-1's output, cbits/lz4.c:2322-2359, :2075-2100), so
// config 4's 128 streams occupy 256 warps of a machine that holds 9 472.  Measured (profiles/r2a_*): the chain is
// PARSER-bound -- with the copier's copies switched off it runs only 10-40 % faster.  But a block's token chain does not
// depend on any other block; only match COPIES do.  So here
//
//   8 PARSER warps   parse blocks N, N+1, .. N+7 of the stream at the same time (the unchanged parse_block, all safe-
//                    decoder rules; the one rule that needs the previous block's length, cbits/lz4.c:2073, is deferred:
//                    the parser reports how far before its block a match reaches, the dispatcher compares).  Each
//                    parser writes 32-descriptor batches into its own ring in GLOBAL memory (256 batches, a whole
//                    64 KiB block: L2-resident) and a 16-byte batch header into a shared-memory ring.
//   7 COPIER warps   take batches strictly in stream order from a dispatcher (a critical section: which parser's ring
//                    is next, the batch's absolute output position, flow control) but EXECUTE them out of order.  The
//                    stream's last 128 KiB of output live in a shared-memory ring indexed by stream position, with one
//                    READY bit per byte (lap parity, so bits never need clearing): literals are copied at once, a
//                    match waits (spinning on the bits) only for the bytes it really reads, then publishes its own.
//                    Matches that are ready together are copied lane-parallel; a dependency chain inside a batch
//                    (records, runs) degrades to one warp-cooperative copy per link, as in the narrow kernel.  Every
//                    warp flushes its own batch to global memory with 128-bit stores.
//
// Flow control: a batch may only be handed out while its end is less than 60 KiB ahead of the in-order completion
// frontier (the ring holds 128 KiB, matches reach 64 KiB back); long sequences are cut into pieces of 16 KiB by the
// parser so that every batch is small against that window.
#include <cstdlib>
#include <mutex>
#include "decode_common.cuh"

namespace b200lz4 {

using namespace dec;

namespace {

constexpr int kWP = kWideParsers;                 // parser warps
constexpr int kWC = 7;                            // copier warps
constexpr int kWThreads = (kWP + kWC) * 32;
constexpr uint32_t kWOut = 131072, kWM = kWOut - 1;      // output ring (bytes), indexed by stream position
constexpr int kWBitWords = kWOut / 32;            // one ready bit per ring byte
constexpr int kWR = kWideRingBatches;             // batches per parser ring
constexpr int kWStage = 1536;                     // per copier: staging of one batch's compressed bytes
constexpr int kTickets = 64;                      // completion flags (ring)
constexpr int kRunAhead = 61440;                  // a batch may end at most this far beyond the completion frontier
constexpr uint32_t kPiece = 16384;                // long sequences are cut into pieces of this many output bytes
constexpr uint32_t kSeedBase = 65536;             // stream position of the first output byte after (re)seeding

struct WCtl {                                     // dispatcher state (shared memory; guarded by `lock` unless noted)
    uint32_t lock, finished, stream_open, reseed_pending, persist_pending, next_iter;
    int s, b0, b1, cur_b;
    uint32_t base;                                // stream position of the current block's first output byte
    uint32_t dict_len;                            // length of the last successful block (cbits/lz4.c:2353-2355)
    uint32_t valid_lo;                            // lowest stream position a match may read
    uint32_t T, F, fpos;                          // tickets handed out / completed in order; position below which all output is final
    const uint8_t* last_out; int last_len; int pad_;
    uint32_t rd[kWP];                             // next batch of each parser ring to hand out
    uint32_t wr_pub[kWP];                         // batches published (written by the parser, lock-free)
    uint32_t cons[kWP];                           // batches retired (read by the parser, lock-free)
    uint32_t done[kTickets];                      // ticket t complete <=> done[t % 64] == t + 1 (written by copiers, lock-free)
    uint32_t tend[kTickets], tn[kTickets], tj[kTickets];
};

__device__ __forceinline__ uint32_t vld(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void vst(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_or(uint32_t a, uint32_t m) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(m) : "memory"); }
__device__ __forceinline__ void red_and(uint32_t a, uint32_t m) { asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(a), "r"(m) : "memory"); }
__device__ __forceinline__ uint4 ld_cg_128(const uint4* p)
{ uint4 v; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }

// ---------------------------------------------------------------- ready bits ----
// Bit i of word (pos >> 5) & 4095 belongs to ring byte pos & 131071.  A byte written on lap L = pos >> 17 sets its bit
// to L & 1; a reader that wants position pos expects (pos >> 17) & 1.  What the previous lap left behind has the other
// parity, so bits are never cleared; the run-ahead rule keeps a writer from lapping a reader.
__device__ __forceinline__ uint32_t bit_addr(uint32_t bits_s, uint32_t pos) { return bits_s + (((pos >> 5) & (kWBitWords - 1)) << 2); }

__device__ __forceinline__ void bits_set(uint32_t bits_s, uint32_t pos, uint32_t n)          // per lane, short ranges
{
    const uint32_t end = pos + n;
    while (pos != end) {
        const uint32_t lo = pos & 31u, cnt = min(32u - lo, end - pos);
        const uint32_t mask = (cnt == 32u ? 0xFFFFFFFFu : ((1u << cnt) - 1u)) << lo;
        if ((pos >> 17) & 1u) red_or(bit_addr(bits_s, pos), mask); else red_and(bit_addr(bits_s, pos), ~mask);
        pos += cnt;
    }
}
__device__ __forceinline__ bool bits_ready(uint32_t bits_s, uint32_t pos, uint32_t n)        // per lane, short ranges
{
    const uint32_t end = pos + n;
    bool ok = true;
    while (pos != end) {
        const uint32_t lo = pos & 31u, cnt = min(32u - lo, end - pos);
        const uint32_t mask = (cnt == 32u ? 0xFFFFFFFFu : ((1u << cnt) - 1u)) << lo;
        const uint32_t want = ((pos >> 17) & 1u) ? 0xFFFFFFFFu : 0u;
        ok = ok && ((((lds32(bit_addr(bits_s, pos)) ^ ~want) & mask)) == mask);
        pos += cnt;
    }
    return ok;
}
// the same for long ranges, all lanes together (words strided over the lanes; position arithmetic may wrap at 2^32)
__device__ __forceinline__ void bits_set_coop(uint32_t bits_s, uint32_t pos, uint32_t n)
{
    if (n == 0) return;
    const uint32_t last = pos + n - 1u, w0 = pos >> 5;
    const uint32_t count = (((last >> 5) - w0) & 0x07FFFFFFu) + 1u;
    for (uint32_t k = lane_id(); k < count; k += 32) {
        const uint32_t w = w0 + k;
        const uint32_t lo = (k == 0) ? (pos & 31u) : 0u, hi = (k == count - 1u) ? (last & 31u) : 31u;
        const uint32_t mask = (hi - lo == 31u) ? 0xFFFFFFFFu : (((1u << (hi - lo + 1u)) - 1u) << lo);
        const uint32_t a = bits_s + ((w & (kWBitWords - 1)) << 2);
        if ((w >> 12) & 1u) red_or(a, mask); else red_and(a, ~mask);
    }
}
__device__ __forceinline__ bool bits_ready_coop(uint32_t bits_s, uint32_t pos, uint32_t n)
{
    bool ok = true;
    if (n) {
        const uint32_t last = pos + n - 1u, w0 = pos >> 5;
        const uint32_t count = (((last >> 5) - w0) & 0x07FFFFFFu) + 1u;
        for (uint32_t k = lane_id(); k < count; k += 32) {
            const uint32_t w = w0 + k;
            const uint32_t lo = (k == 0) ? (pos & 31u) : 0u, hi = (k == count - 1u) ? (last & 31u) : 31u;
            const uint32_t mask = (hi - lo == 31u) ? 0xFFFFFFFFu : (((1u << (hi - lo + 1u)) - 1u) << lo);
            const uint32_t want = ((w >> 12) & 1u) ? 0xFFFFFFFFu : 0u;
            ok = ok && (((lds32(bits_s + ((w & (kWBitWords - 1)) << 2)) ^ ~want) & mask) == mask);
        }
    }
    return __all_sync(kFull, ok);
}

// ------------------------------------------------------------- ring <-> global ----
__device__ __forceinline__ uint32_t rix(uint32_t out_s, uint32_t pos) { return out_s + (pos & kWM); }

// n bytes of global memory -> ring positions [pos, pos + n); all lanes, identical arguments
template <bool kReadOnly>
__device__ __forceinline__ void copy_g2r(uint32_t out_s, uint32_t pos, const uint8_t* src, uint32_t n)
{
    const uint32_t lane = lane_id();
    uint32_t head = (4u - (pos & 3u)) & 3u;
    if (head > n) head = n;
    if (lane < head) sts8(rix(out_s, pos + lane), kReadOnly ? (uint32_t)__ldg(src + lane) : (uint32_t)src[lane]);
    pos += head; src += head; n -= head;
    const uint32_t nw = n >> 2;
    for (uint32_t w = lane; w < nw; w += 32)
        sts32(rix(out_s, pos + 4u * w), kReadOnly ? ldg_u32_unaligned(src + 4u * w) : ld_u32_unaligned(src + 4u * w));
    const uint32_t done = nw << 2, tail = n - done;
    if (lane < tail) sts8(rix(out_s, pos + done + lane), kReadOnly ? (uint32_t)__ldg(src + done + lane) : (uint32_t)src[done + lane]);
}

// ring positions [pos, pos + n) -> global memory at gdst (128-bit stores; the ring is read in aligned words)
__device__ __forceinline__ void flush_r2g(uint32_t out_s, uint8_t* gdst, uint32_t pos, uint32_t n)
{
    const uint32_t lane = lane_id();
    uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(gdst) & 15)) & 15);
    if (head > n) head = n;
    if (lane < head) gdst[lane] = (uint8_t)lds8(rix(out_s, pos + lane));
    gdst += head; pos += head; n -= head;
    const uint32_t nvec = n >> 4;
    uint4* gv = reinterpret_cast<uint4*>(gdst);
    if ((pos & 15u) == 0) {
        for (uint32_t v = lane; v < nvec; v += 32) gv[v] = lds128(rix(out_s, pos + 16u * v));
    } else {
        const uint32_t sh = (pos & 3u) * 8u;
        for (uint32_t v = lane; v < nvec; v += 32) {
            const uint32_t a = (pos + 16u * v) & ~3u;
            const uint32_t x0 = lds32(rix(out_s, a)), x1 = lds32(rix(out_s, a + 4)), x2 = lds32(rix(out_s, a + 8)), x3 = lds32(rix(out_s, a + 12));
            const uint32_t x4 = sh ? lds32(rix(out_s, a + 16)) : 0u;
            gv[v] = make_uint4(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh), __funnelshift_r(x3, x4, sh));
        }
    }
    const uint32_t done = nvec << 4, tail = n - done;
    if (lane < tail) gdst[done + lane] = (uint8_t)lds8(rix(out_s, pos + done + lane));
}

// One match of at most 64 bytes copied by the whole warp (two bytes per lane).  Its source -- the dist bytes below dst,
// or the len bytes from dst - dist on -- is complete; an overlapping match (dist < len) repeats those dist bytes:
// byte i comes from source byte i mod dist, through a 16-bit fixed-point reciprocal that is exact for i < 64.
__device__ __forceinline__ void coop_short_match(uint32_t out_s, uint32_t dst, uint32_t len, uint32_t dist)
{
    const uint32_t lane = lane_id(), i1 = lane + 32u;
    const uint32_t inv = (dist < len) ? (uint32_t)(65536.0f * __frcp_rn((float)dist)) + 2u : 0u;
    const uint32_t src = dst - dist;
    const uint32_t k0 = lane - ((lane * inv) >> 16) * dist, k1 = i1 - ((i1 * inv) >> 16) * dist;
    if (lane < len) sts8(rix(out_s, dst + lane), lds8(rix(out_s, src + k0)));
    if (i1 < len) sts8(rix(out_s, dst + i1), lds8(rix(out_s, src + k1)));
}

// A long match piece inside the ring, all lanes: rounds of one period each (the period's source is complete before
// the round starts); periods below 32 are replicated from registers.
__device__ __forceinline__ void coop_long_match(uint32_t out_s, uint32_t dst, uint32_t len, uint32_t dist)
{
    const uint32_t lane = lane_id();
    const uint32_t src = dst - dist;
    if (dist >= 32u) {
        for (uint32_t done = 0; done < len; done += dist) {
            const uint32_t n = min(dist, len - done);
            for (uint32_t i = lane; i < n; i += 32) sts8(rix(out_s, dst + done + i), lds8(rix(out_s, src + done + i)));
            __syncwarp();
        }
    } else {
        const uint32_t pat = lds8(rix(out_s, src + (lane < dist ? lane : 0u)));
        uint32_t k = lane % dist;
        const uint32_t adv = 32u % dist;
        for (uint32_t i0 = 0; i0 < len; i0 += 32) {
            const uint32_t bb = __shfl_sync(kFull, pat, k);
            if (i0 + lane < len) sts8(rix(out_s, dst + i0 + lane), bb);
            k += adv; if (k >= dist) k -= dist;
        }
    }
}

// (re)start the ring at stream position kSeedBase: ready bits for lap 0 -- the lower half (which may hold the
// dictionary, right-aligned below kSeedBase) reads as complete, the upper half as not yet written -- then the
// dictionary bytes themselves.
__device__ void wseed(uint32_t out_s, uint32_t bits_s, const uint8_t* tail_src, uint32_t kept)
{
    const uint32_t lane = lane_id();
    for (uint32_t i = lane; i < kWBitWords / 4; i += 32) {
        const uint32_t v = (i < kWBitWords / 8) ? 0u : 0xFFFFFFFFu;
        sts128(bits_s + 16u * i, v, v, v, v);
    }
    if (kept) copy_g2r<false>(out_s, kSeedBase - kept, tail_src, kept);
    __threadfence_block();
    __syncwarp();
}

// ------------------------------------------------------------------ parser sink ----
struct WParser {
    static constexpr bool kWide = true;
    static constexpr int kRing = kInRing;
    uint32_t ring_s;            // shared address of this parser's input ring
    const uint8_t* gbase;       // global address of ring-space position 0
    int end, issued_end, ready_end, cur_start;
    uint4* gring;               // this parser's descriptor ring (global memory)
    uint32_t hdr_s;             // shared address of its header ring: {count | flags, first output position, end position, reach}
    uint32_t* wr_pub; uint32_t* cons;
    uint32_t batch;             // batches published so far
    int fill, flags;
    int op_start, op_end;       // block-relative output range of the batch being filled
    int need;                   // bytes before the block start the furthest-reaching match reads (deferred cbits/lz4.c:2073)

    __device__ __forceinline__ uint32_t at(int p) const { return lds8(ring_s + ((uint32_t)p & (kRing - 1))); }
    __device__ __forceinline__ void begin()
    {   // the ring slot of this batch must have been retired
        while (batch - vld(cons) >= (uint32_t)kWR) __nanosleep(200);
        fill = 0;
    }
    __device__ __forceinline__ uint4* slot() const { return gring + (size_t)(batch & (kWR - 1)) * 32; }
    __device__ __forceinline__ void put(uint32_t rank, uint32_t x, uint32_t y, uint32_t z, uint32_t w) { slot()[fill + rank] = make_uint4(x, y, z, w); }
    __device__ __forceinline__ void advance(int n, int op_after) { fill += n; op_end = op_after; }
    __device__ __forceinline__ void push(uint32_t x, uint32_t y, uint32_t z, uint32_t w, int op_after)
    {
        if (lane_id() == 0) slot()[fill] = make_uint4(x, y, z, w);
        fill++; op_end = op_after;
    }
    __device__ __forceinline__ void note_reach(int r) { need = r > need ? r : need; }
    __device__ void publish(int extra, int result, int ip)
    {
        uint32_t cf = (uint32_t)(fill | flags | extra);
        if ((extra & kEndBlock) && result < 0) cf |= (uint32_t)kFailed;
        if (extra & kEndBlock) cp_async_wait_all();      // nothing of this block's input may land after the next block starts
        __threadfence();                                 // descriptor stores (several lanes) before the header
        __syncwarp();
        if (lane_id() == 0) {
            sts128(hdr_s + 16u * (batch & (kWR - 1)), cf, (uint32_t)op_start, (uint32_t)op_end, (extra & kEndBlock) ? (uint32_t)need : 0u);
            __threadfence_block();
            vst(wr_pub, batch + 1);
        }
        batch++;
        flags = 0; op_start = op_end; cur_start = ip;
        begin();
    }
    // a long sequence: pieces of at most kPiece output bytes, one batch each (descriptor: literal start, literals, match bytes, offset)
    __device__ void push_bulk(uint32_t lit_src, uint32_t lit, uint32_t mlen, uint32_t dist, int op_seq, int ip_after)
    {
        int op = op_seq;
        while (lit) {
            const uint32_t n = lit < kPiece ? lit : kPiece;
            op_start = op;
            push(lit_src, n, 0u, 0u, op + (int)n);
            publish(kBulk, 0, ip_after);
            lit_src += n; lit -= n; op += (int)n;
        }
        while (mlen) {
            const uint32_t n = mlen < kPiece ? mlen : kPiece;
            op_start = op;
            push(0u, 0u, n, dist, op + (int)n);
            publish(kBulk, 0, ip_after);
            mlen -= n; op += (int)n;
        }
    }
    __device__ __forceinline__ void top_up(int ip)
    {   // only this warp reads its ring, and it never looks back more than a window: 2.5 KiB ahead always fits 4 KiB
        while (issued_end < end && issued_end - ip < 2048) {
            const int p = issued_end + 16 * (int)lane_id();
            if (p < end) cp_async_16(ring_s + ((uint32_t)p & (kRing - 1)), gbase + p);
            cp_async_commit();
            issued_end += kFill;
        }
    }
    __device__ void ensure(int ip, int need_pos)
    {
        if (need_pos <= ready_end) return;
        if (ip >= issued_end) {                         // jumped over everything requested so far: restart at ip
            cp_async_wait_all();
            issued_end = ip & ~(kFill - 1);
            ready_end = issued_end;
        }
        top_up(ip);
        cp_async_wait_all();
        __syncwarp();
        ready_end = issued_end;
    }
};

__device__ void wparser_main(const DecompressArgs& a, int j, WCtl* ctl, uint32_t ring_s, uint32_t hdr_s, uint4* gring)
{
    WParser P;
    P.ring_s = ring_s; P.gbase = nullptr; P.end = P.issued_end = P.ready_end = P.cur_start = 0;
    P.gring = gring; P.hdr_s = hdr_s; P.wr_pub = &ctl->wr_pub[j]; P.cons = &ctl->cons[j];
    P.batch = 0; P.fill = 0; P.flags = 0; P.op_start = P.op_end = 0; P.need = 0;
    for (int s = blockIdx.x; s < a.n_streams; s += gridDim.x) {
        const int b0 = a.stream_first ? a.stream_first[s] : a.first_block + s;
        const int b1 = a.stream_first ? a.stream_first[s + 1] : b0 + 1;
        for (int b = b0 + j; b < b1; b += kWP) {        // block (b - b0) of a stream belongs to parser (b - b0) % 8
            const BlockGeom g = block_geom(a, b);
            P.flags = kBegin; P.need = 0; P.op_start = P.op_end = 0;
            int r = -1;
            if (g.ok) {
                const int skew = (int)(reinterpret_cast<uintptr_t>(g.payload) & 15);
                P.gbase = g.payload - skew;
                r = parse_block(P, skew, g.comp_len, g.cap, 0u);
            }
            P.publish(kEndBlock, r, 0);
        }
    }
}

// ------------------------------------------------------------------ dispatcher ----
__device__ __forceinline__ void wlock(WCtl* c)
{
    if (lane_id() == 0) while (atomicCAS(&c->lock, 0u, 1u) != 0u) __nanosleep(64);
    __syncwarp();
    __threadfence_block();
}
__device__ __forceinline__ void wunlock(WCtl* c)
{
    __threadfence_block();
    __syncwarp();
    if (lane_id() == 0) atomicExch(&c->lock, 0u);
}
// advance the in-order completion frontier over finished tickets; retire their ring slots (lock held; every lane runs
// the same code on the same shared-memory words)
__device__ __forceinline__ void wadvance(WCtl* c)
{
    uint32_t F = vld(&c->F);
    const uint32_t T = vld(&c->T);
    while (F != T && vld(&c->done[F & (kTickets - 1)]) == F + 1) {
        const uint32_t k = F & (kTickets - 1);
        vst(&c->fpos, vld(&c->tend[k]));
        vst(&c->cons[vld(&c->tj[k])], vld(&c->tn[k]) + 1);
        F++;
    }
    vst(&c->F, F);
}
__device__ __forceinline__ void wdrain(WCtl* c)
{
    for (;;) { wadvance(c); if (vld(&c->F) == vld(&c->T)) break; __nanosleep(100); }
    __threadfence_block();
}

__device__ void wcopier_main(const DecompressArgs& a, WCtl* ctl, uint32_t out_s, uint32_t bits_s, uint32_t hdrs_s,
                             uint32_t stage_s, const uint4* garena)
{
    const uint32_t lane = lane_id();
    volatile WCtl* const c = ctl;        // every field access below is a real shared-memory access
    int cached_blk = -1; uint8_t* blk_out = nullptr; const uint8_t* blk_gbase = nullptr;
    for (;;) {
        wlock(ctl);
        // ---- stream bookkeeping: close / reseed / open
        bool fin = false;
        for (;;) {
            if (c->finished) { fin = true; break; }
            DState* st = (a.states && c->s >= 0) ? reinterpret_cast<DState*>(a.states[c->s]) : nullptr;
            if (c->reseed_pending) {
                // a block of this stream failed: what it wrote into the ring is garbage; restart the ring from the last
                // good output (which stays the dictionary, cbits/lz4.c:2353), taken from global memory
                wdrain(ctl);
                uint32_t kept = 0, dl = 0;
                if (c->last_out) {
                    dl = (uint32_t)c->last_len; kept = dl < 65536u ? dl : 65536u;
                    wseed(out_s, bits_s, c->last_out + c->last_len - kept, kept);
                } else if (st && st->prev_len) {
                    dl = st->prev_len; kept = st->kept;
                    wseed(out_s, bits_s, st->tail + 65536 - kept, kept);
                } else wseed(out_s, bits_s, nullptr, 0u);
                c->base = kSeedBase; c->fpos = kSeedBase; c->dict_len = dl; c->valid_lo = kSeedBase - kept;
                c->reseed_pending = 0u;
            }
            if (c->stream_open) break;
            wdrain(ctl);
            if (c->persist_pending) {        // keep the reachable tail of the last output for the next call
                const uint32_t kept = c->last_len < 65536 ? (uint32_t)c->last_len : 65536u;
                warp_copy_rw(st->tail + 65536 - kept, c->last_out + c->last_len - kept, kept);
                __syncwarp();
                if (lane == 0) { st->prev_len = (uint32_t)c->last_len; st->kept = kept; }
                c->persist_pending = 0u;
            }
            const int s = (int)blockIdx.x + (int)c->next_iter * (int)gridDim.x;
            if (s >= a.n_streams) { c->finished = 1u; fin = true; break; }
            c->next_iter = c->next_iter + 1;
            const int b0 = a.stream_first ? a.stream_first[s] : a.first_block + s;
            const int b1 = a.stream_first ? a.stream_first[s + 1] : b0 + 1;
            const DState* ns = a.states ? reinterpret_cast<const DState*>(a.states[s]) : nullptr;
            uint32_t kept = 0, dl = 0;
            if (ns && ns->prev_len) { dl = ns->prev_len; kept = ns->kept; wseed(out_s, bits_s, ns->tail + 65536 - kept, kept); }
            else wseed(out_s, bits_s, nullptr, 0u);
            c->s = s; c->b0 = b0; c->b1 = b1; c->cur_b = b0;
            c->last_out = nullptr; c->last_len = 0;
            c->base = kSeedBase; c->fpos = kSeedBase; c->dict_len = dl; c->valid_lo = kSeedBase - kept;
            c->stream_open = b0 < b1 ? 1u : 0u;
            __syncwarp();
        }
        if (fin) { wunlock(ctl); break; }

        // ---- hand out the next batch of the stream (this warp takes it)
        const int blk = c->cur_b;
        const uint32_t j = (uint32_t)(blk - c->b0) % (uint32_t)kWP;
        const uint32_t n = c->rd[j];
        while (c->wr_pub[j] == n) { wadvance(ctl); __nanosleep(40); }      // (retiring slots may be what the parser waits for)
        __threadfence_block();
        const uint4 h = lds128(hdrs_s + (j * kWR + (n & (kWR - 1))) * 16u);
        const uint32_t base = c->base, valid_lo = c->valid_lo;
        const uint32_t e = base + h.z;
        for (;;) {
            wadvance(ctl);
            if ((int)(e - c->fpos) <= kRunAhead && c->T - c->F < (uint32_t)(kTickets - 8)) break;
            __nanosleep(40);
        }
        const uint32_t t = c->T;
        {
            const uint32_t k = t & (kTickets - 1);
            c->tend[k] = e; c->tj[k] = j; c->tn[k] = n;
            c->T = t + 1; c->rd[j] = n + 1;
        }
        const int cf = (int)h.x;
        const uint32_t op_start = h.y;
        if (cf & kEndBlock) {
            int r = (cf & kFailed) ? -1 : (int)h.z;
            const uint32_t dl = c->dict_len;
            if (r >= 0 && dl < 65536u && h.w > dl) r = -1;                  // cbits/lz4.c:2073, deferred: a match reached below the dictionary
            if (lane == 0) a.out_len[blk] = r;
            if (r > 0) {                                                    // cbits/lz4.c:2353-2355
                c->base = base + (uint32_t)r; c->dict_len = (uint32_t)r;
                c->valid_lo = base + (uint32_t)r - ((uint32_t)r < 65536u ? (uint32_t)r : 65536u);
                c->last_out = a.dst + a.dst_off[blk]; c->last_len = r;
            } else if (r < 0 && blk + 1 < c->b1) c->reseed_pending = 1u;
            c->cur_b = blk + 1;
            if (blk + 1 == c->b1) { c->stream_open = 0u; c->persist_pending = (a.states && c->last_out) ? 1u : 0u; }
        }
        wunlock(ctl);

        // ---- execute it
        const int cnt = cf & kCountMask;
        uint4 d = make_uint4(0, 0, 0, 0);
        if ((int)lane < cnt) d = ld_cg_128(garena + ((size_t)j * kWR + (n & (kWR - 1))) * 32 + lane);
        if (blk != cached_blk) {
            const BlockGeom g = block_geom(a, blk);
            blk_out = g.out; blk_gbase = g.payload - (reinterpret_cast<uintptr_t>(g.payload) & 15);
            cached_blk = blk;
        }
        if (cf & kBulk) {
            // one piece of a long sequence
            const uint32_t lit_src = __shfl_sync(kFull, d.x, 0), lit = __shfl_sync(kFull, d.y, 0);
            const uint32_t mlen = __shfl_sync(kFull, d.z, 0), dist = __shfl_sync(kFull, d.w, 0);
            const uint32_t pos = base + op_start;
            uint8_t* gout = blk_out + op_start;
            if (lit) {
                copy_g2r<true>(out_s, pos, blk_gbase + lit_src, lit);
                __threadfence_block();
                __syncwarp();
                bits_set_coop(bits_s, pos, lit);
                warp_copy_ro(gout, blk_gbase + lit_src, lit);
            }
            if (mlen) {
                const uint32_t m_pos = pos + lit, from = m_pos - dist;
                if ((int)(from - valid_lo) >= 0) {                          // (else: the block is rejected at its end; nothing to wait for)
                    const uint32_t need_n = mlen < dist ? mlen : dist;
                    while (!bits_ready_coop(bits_s, from, need_n)) __nanosleep(64);
                    __threadfence_block();
                    coop_long_match(out_s, m_pos, mlen, dist);
                }
                __threadfence_block();
                __syncwarp();
                bits_set_coop(bits_s, m_pos, mlen);
                flush_r2g(out_s, gout + lit, m_pos, mlen);
            }
        } else if (cnt) {
            // up to 32 short sequences, one per lane
            const bool is_seq = (int)lane < cnt;
            const uint32_t lit = d.y & 0xFFu, mlen = (d.y >> 8) & 0xFFu, dist = d.y >> 16;
            const uint32_t lit_pos = base + d.z, m_pos = lit_pos + lit, from = m_pos - dist;
            const uint32_t s_pos = __shfl_sync(kFull, lit_pos, 0);
            const uint32_t e_pos = __shfl_sync(kFull, m_pos + mlen, cnt - 1);
            // the batch's compressed bytes (its literals lie between the first literal start and the last literal end):
            // one coalesced 128-bit pass into this warp's staging buffer
            const uint32_t lo = __shfl_sync(kFull, d.x, 0) & ~15u;
            const uint32_t hi = __shfl_sync(kFull, d.x + lit, cnt - 1);
            const uint32_t nvec = (hi - lo + 15u) >> 4;
            const bool staged = nvec <= (uint32_t)(kWStage / 16);
            if (staged) {
                const uint4* gv = reinterpret_cast<const uint4*>(blk_gbase + lo);
                for (uint32_t v = lane; v < nvec; v += 32) { const uint4 x = ldg_na_u128(gv + v); sts128(stage_s + 16u * v, x.x, x.y, x.z, x.w); }
            }
            __syncwarp();
            if (is_seq && lit) {
                if (staged) {
                    const uint32_t sl = stage_s + (d.x - lo);
                    for (uint32_t i = 0; i < lit; i++) sts8(rix(out_s, lit_pos + i), lds8(sl + i));
                } else {
                    const uint8_t* gp = blk_gbase + d.x;
                    for (uint32_t i = 0; i < lit; i++) sts8(rix(out_s, lit_pos + i), (uint32_t)__ldg(gp + i));
                }
            }
            __threadfence_block();
            if (is_seq && lit) bits_set(bits_s, lit_pos, lit);
            // matches: whichever are ready, round after round
            bool pend = is_seq && mlen != 0;
            const uint32_t need_n = mlen < dist ? mlen : dist;
            const bool doomed = (int)(from - valid_lo) < 0;                 // reads below the dictionary: the block will be rejected
            for (;;) {
                const bool rdy = pend && (doomed || bits_ready(bits_s, from, need_n));
                const uint32_t rb = __ballot_sync(kFull, rdy);
                if (rb) {
                    __threadfence_block();
                    if (__popc(rb) <= 4) {                                  // a chain: one warp-wide copy per link
                        uint32_t todo = rb;
                        while (todo) {
                            const int l = __ffs(todo) - 1; todo &= todo - 1;
                            const uint32_t c_dst = __shfl_sync(kFull, m_pos, l), c_len = __shfl_sync(kFull, mlen, l);
                            const uint32_t c_dist = __shfl_sync(kFull, dist, l);
                            const bool c_doomed = __shfl_sync(kFull, (int)doomed, l) != 0;
                            if (!c_doomed) coop_short_match(out_s, c_dst, c_len, c_dist);
                        }
                    } else if (rdy && !doomed) {                            // many at once: every lane copies its own
                        for (uint32_t i = 0; i < mlen; i++) sts8(rix(out_s, m_pos + i), lds8(rix(out_s, from + i)));
                    }
                    __threadfence_block();
                    __syncwarp();
                    if (rdy) bits_set(bits_s, m_pos, mlen);
                    pend = pend && !rdy;
                }
                if (!__ballot_sync(kFull, pend)) break;
                if (!rb) __nanosleep(32);
            }
            __syncwarp();
            flush_r2g(out_s, blk_out + (s_pos - base), s_pos, e_pos - s_pos);
        }
        // ---- report completion (in-order retirement happens in the dispatcher)
        __threadfence_block();
        __syncwarp();
        if (lane == 0) c->done[t & (kTickets - 1)] = t + 1;
    }
}

constexpr size_t kWSmem = 128 /* alignment slack */ + kWOut + kWBitWords * 4 + kWP * kInRing + kWP * kWR * 16 + kWC * kWStage + sizeof(WCtl);

__global__ void __launch_bounds__(kWThreads, 1)
decompress_kernel_wide(DecompressArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_dyn[];
    uint8_t* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    uint8_t* out_ring = base;
    uint8_t* bits = out_ring + kWOut;
    uint8_t* in_rings = bits + kWBitWords * 4;
    uint8_t* hdrs = in_rings + kWP * kInRing;
    uint8_t* stages = hdrs + kWP * kWR * 16;
    WCtl* ctl = reinterpret_cast<WCtl*>(stages + kWC * kWStage);
    for (uint32_t i = threadIdx.x; i < sizeof(WCtl) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ctl)[i] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) ctl->s = -1;
    __syncthreads();
    const int warp = (int)(threadIdx.x >> 5);
    uint4* garena = a.wide_arena + (size_t)blockIdx.x * (kWideArenaPerCta / 16);
    uint32_t out_s, bits_s;         // laundered so that the compiler keeps them in registers
    asm volatile("mov.u32 %0, %1;" : "=r"(out_s) : "r"(smem_u32(out_ring)));
    asm volatile("mov.u32 %0, %1;" : "=r"(bits_s) : "r"(smem_u32(bits)));
    if (warp < kWP)
        wparser_main(a, warp, ctl, smem_u32(in_rings) + (uint32_t)warp * kInRing, smem_u32(hdrs) + (uint32_t)warp * kWR * 16u,
                     garena + (size_t)warp * kWR * 32);
    else
        wcopier_main(a, ctl, out_s, bits_s, smem_u32(hdrs), smem_u32(stages) + (uint32_t)(warp - kWP) * kWStage, garena);
}

}  // namespace

cudaError_t launch_decompress_wide(const DecompressArgs& a, int sm_count, cudaStream_t stream)
{
    static std::mutex mu;
    static bool configured[64] = {false};
    int dev = 0; cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!configured[dev]) {
            e = cudaFuncSetAttribute(decompress_kernel_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWSmem);
            if (e != cudaSuccess) return e;
            configured[dev] = true;
        }
    }
    int grid = a.n_streams < sm_count ? a.n_streams : sm_count;
    if (grid > a.wide_ctas) grid = a.wide_ctas;
    if (grid <= 0) return cudaErrorInvalidValue;
    decompress_kernel_wide<<<grid, kWThreads, kWSmem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace b200lz4
