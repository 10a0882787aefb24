// compress.cu -- byte-exact LZ4 block compressor for sm_100a.
//
// Reproduces LZ4_compress_fast_continue in external-dictionary mode
// (cbits/lz4.c:1565-1637 -> LZ4_compress_generic_validated :851-1240 with
// limitedOutput/byU32/usingExtDict) bit for bit, but is organised for a SIMT machine:
//
//   * one WARP owns one stream (independent mode: one block) and keeps the 16 KiB
//     position table in shared memory for the whole stream;
//   * the serial "probe, overwrite, test" recurrence of the match finder
//     (cbits/lz4.c:959-1014) is evaluated 32 probes at a time: the probe positions of
//     a search run follow a closed-form schedule (step_k = (acc*64 + k - 1) >> 6), so
//     lane l speculatively evaluates probe j0+l; same-bucket forwarding inside the
//     window is resolved with __match_any_sync (a lane's candidate is the nearest
//     lower lane with the same hash, else the table), the lowest accepting lane wins,
//     lanes up to the winner commit their table writes in order (last writer per
//     bucket wins), later lanes are discarded;
//   * the post-match re-test at ip (cbits/lz4.c:1159-1196) rides in the same window
//     as probe index -1 of the next run;
//   * catch-up, match-length counting and literal copies are warp-parallel.
#include "common.cuh"
#include "kernels.h"

namespace b200lz4 {

namespace {

struct BlockIn {
    const uint8_t* src; int n;
    const uint8_t* dict_end;    // one past the last dictionary byte (valid iff dict_len > 0)
    uint32_t dict_len;
    uint32_t start;             // currentOffset before this block
    uint8_t* dst; int cap;
    int accel;
};

// offset (from the run start) and step of probe j >= -1 of a search run.
__device__ __forceinline__ void probe_schedule(long long j, int accel, long long& off, int& step)
{
    if (j <= 0) { off = j; step = 1; return; }       // j == -1: the re-test position; j == 0: run start
    long long m = j - 1, q = m >> 6, r = m & 63;
    off = 1 + m * accel + 32 * q * (q - 1) + q * r;
    step = accel + (int)q;
}

// common-prefix length of a[0..cap) and b[0..cap); a, b read-only global. Warp-parallel,
// 4 bytes per lane per round. (What LZ4_count returns, cbits/lz4.c:603-626.)
__device__ __forceinline__ uint32_t warp_common_prefix(const uint8_t* a, const uint8_t* b, uint32_t cap)
{
    const uint32_t lane = lane_id();
    uint32_t base = 0;
    for (;;) {
        uint32_t at = base + lane * 4;
        uint32_t nb = 0;            // equal bytes in this lane's chunk
        bool stop = true;
        if (at < cap) {
            uint32_t room = cap - at;
            if (room >= 4) {
                uint32_t x = ldg_u32_unaligned(a + at) ^ ldg_u32_unaligned(b + at);
                nb = x ? ((uint32_t)(__ffs(x) - 1) >> 3) : 4u;
                stop = (nb < 4);
            } else {
                while (nb < room && __ldg(a + at + nb) == __ldg(b + at + nb)) nb++;
                stop = true;
            }
        }
        uint32_t sb = __ballot_sync(kFull, stop);
        if (sb) {
            int l = __ffs(sb) - 1;
            uint32_t res = __shfl_sync(kFull, at + nb, l);
            return res < cap ? res : cap;
        }
        base += 128;
    }
}

__device__ __forceinline__ void emit_len_ext(uint8_t* op, uint32_t rest, uint32_t n_ext)
{   // n_ext = rest/255 + 1 bytes: 255,255,...,rest%255
    const uint32_t lane = lane_id();
    for (uint32_t i = lane; i < n_ext; i += 32) op[i] = (i + 1 < n_ext) ? 255 : (uint8_t)(rest % 255);
}

// Encode one block. Returns compressed size (0 = does not fit `cap`).
// `table` is this warp's shared-memory table; on return it holds the reference's table state.
__device__ int encode_block(const BlockIn& in, uint32_t* table)
{
    const uint32_t lane = lane_id();
    const uint8_t* const src = in.src;
    const int n = in.n;
    uint8_t* const dst = in.dst;
    const uint32_t S = in.start;
    const bool dict_small = (in.dict_len < 65536u) && (in.dict_len < S);   // cbits/lz4.c:1627
    const uint32_t low_index = S - in.dict_len;                             // prefixIdxLimit, :879
    const int mfl = n - kMfLimit + 1;          // mflimitPlusOne as an index
    const int mlimit = n - kLastLiterals;      // matchlimit
    long long op = 0;
    int anchor = 0;

    if (n >= kMinLength) {
        if (lane == 0) { uint2 v = ldg_5bytes(src); table[hash5(v.x, v.y)] = S; }   // :924
        __syncwarp();
        int run_start = 1;          // :925
        bool retest = false;
        for (;;) {
            // ---------------- speculative probe window(s) ----------------
            long long jbase = retest ? -1 : 0;
            int mpos = 0; uint32_t midx = 0; bool found = false;
            for (;;) {
                long long j = jbase + lane, off; int step;
                probe_schedule(j, in.accel, off, step);
                long long pos64 = (long long)run_start + off;
                bool valid = (j < 0) || (pos64 + step <= (long long)mfl);      // :969
                uint32_t h = 0, seq = 0, cur = 0;
                int pos = valid ? (int)pos64 : 0;
                if (valid) {
                    uint2 v = ldg_5bytes(src + pos);
                    seq = v.x; h = hash5(v.x, v.y); cur = S + (uint32_t)pos;
                }
                uint32_t key = valid ? h : (0x1000u + lane);
                uint32_t peers = __match_any_sync(kFull, key);
                uint32_t lower = peers & lanemask_lt();
                int from_lane = lower ? (31 - __clz(lower)) : (int)lane;
                uint32_t fwd = __shfl_sync(kFull, cur, from_lane);
                bool ok = false; uint32_t m = 0;
                if (valid) {
                    m = lower ? fwd : table[h];
                    bool reach = !(dict_small && m < low_index) && (m + kMaxDistance >= cur);   // :1001-1006
                    if (reach) {
                        const uint8_t* c = (m < S) ? (in.dict_end - (S - m)) : (src + (m - S));   // :985-993
                        ok = (ldg_u32_unaligned(c) == seq);                                     // :1009
                    }
                }
                uint32_t okb = __ballot_sync(kFull, ok);
                uint32_t vb = __ballot_sync(kFull, valid);
                int nvalid = __popc(vb);
                int w = okb ? (__ffs(okb) - 1) : 32;
                int ncommit = min(w + 1, nvalid);
                if ((int)lane < ncommit) {              // ordered commit: last writer per bucket
                    uint32_t grp = peers & (ncommit >= 32 ? kFull : ((1u << ncommit) - 1u));
                    if ((31 - __clz(grp)) == (int)lane) table[h] = cur;     // :998
                }
                __syncwarp();
                if (w < 32) { found = true; mpos = __shfl_sync(kFull, pos, w); midx = __shfl_sync(kFull, m, w); break; }
                if (nvalid < 32) break;                 // ran into mflimit: last literals
                jbase += 32;
            }
            if (!found) break;

            // ---------------- one sequence ----------------
            int ip = mpos;
            const uint32_t dist = (S + (uint32_t)ip) - midx;
            const bool in_dict = midx < S;
            const uint8_t* cand = in_dict ? (in.dict_end - (S - midx)) : (src + (midx - S));
            {   // catch up, :1019
                long long room_c = in_dict ? (long long)in.dict_len - (long long)(S - midx) : (long long)(midx - S);
                long long maxback = min((long long)(ip - anchor), room_c);
                long long back = 0;
                while (back < maxback) {
                    long long k = back + lane + 1;
                    bool eq = (k <= maxback) && (__ldg(src + ip - k) == __ldg(cand - k));
                    uint32_t b = __ballot_sync(kFull, eq);
                    int run = (b == kFull) ? 32 : (__ffs(~b) - 1);
                    back += run;
                    if (run < 32) break;
                }
                ip -= (int)back; cand -= back;
            }
            const uint32_t lit = (uint32_t)(ip - anchor);
            // match length, :1076-1095
            uint32_t mcode;
            if (in_dict) {
                uint32_t room_dict = (uint32_t)(in.dict_end - cand);
                uint32_t lim = min(room_dict, (uint32_t)(mlimit - ip));         // limit - ip
                mcode = warp_common_prefix(src + ip + 4, cand + 4, lim - 4);
                if (4 + mcode == lim) {
                    int at = ip + (int)lim;
                    mcode += warp_common_prefix(src + at, src, (uint32_t)(mlimit - at));
                }
            } else {
                mcode = warp_common_prefix(src + ip + 4, cand + 4, (uint32_t)(mlimit - (ip + 4)));
            }
            // limitedOutput guards, :1024-1027 and :1097-1121
            if (op + 1 + lit + (2 + 1 + kLastLiterals) + lit / 255 > in.cap) return 0;
            {
                uint8_t* tok = dst + op;
                long long o = op + 1;
                if (lit >= 15) { uint32_t rest = lit - 15, ne = rest / 255 + 1; emit_len_ext(dst + o, rest, ne); o += ne; }
                warp_copy_ro(dst + o, src + anchor, lit); o += lit;
                if (lane == 0) { dst[o] = (uint8_t)dist; dst[o + 1] = (uint8_t)(dist >> 8); }   // :1068
                o += 2;
                if (o + (1 + kLastLiterals) + (mcode + 240) / 255 > in.cap) return 0;
                if (mcode >= 15) { uint32_t rest = mcode - 15, ne = rest / 255 + 1; emit_len_ext(dst + o, rest, ne); o += ne; }
                if (lane == 0) *tok = (uint8_t)((min(lit, 15u) << 4) | min(mcode, 15u));
                op = o;
            }
            ip += 4 + (int)mcode;
            anchor = ip;
            if (ip >= mfl) break;                                                            // :1143
            if (lane == 0) { uint2 v = ldg_5bytes(src + ip - 2); table[hash5(v.x, v.y)] = S + (uint32_t)(ip - 2); }   // :1146
            __syncwarp();
            run_start = ip + 1;      // re-test at ip is probe -1 of the next run (:1159-1200)
            retest = true;
        }
    }
    {   // last literals, :1204-1231
        uint32_t run = (uint32_t)(n - anchor);
        if (op + run + 1 + ((run + 255 - 15) / 255) > in.cap) return 0;
        long long o = op + 1;
        if (run >= 15) { uint32_t rest = run - 15, ne = rest / 255 + 1; emit_len_ext(dst + o, rest, ne); o += ne; }
        if (lane == 0) dst[op] = (uint8_t)(min(run, 15u) << 4);
        warp_copy_ro(dst + o, src + anchor, run); o += run;
        op = o;
    }
    return (int)op;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
compress_kernel(CompressArgs a)
{
    extern __shared__ uint32_t smem_tables[];
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    uint32_t* table = smem_tables + warp * kHashEntries;
    uint32_t* counter = &a.scratch->work_counter[0];
    const int accel = a.accel < 1 ? 1 : (a.accel > kAccelMax ? kAccelMax : a.accel);   // :1577-1578

    for (;;) {
        int s = 0;
        if (lane == 0) s = (int)atomicAdd(counter, 1u);
        s = __shfl_sync(kFull, s, 0);
        if (s >= a.n_streams) break;
        const int b0 = a.stream_first ? a.stream_first[s] : s;
        const int b1 = a.stream_first ? a.stream_first[s + 1] : s + 1;
        CState* st = a.states ? reinterpret_cast<CState*>(a.states[s]) : nullptr;

        uint32_t offset = 0, dict_len = 0;
        const uint8_t* dict_end = nullptr;
        {   // table: zero (fresh LZ4_initStream, :1443-1451) or restored
            uint4* t4 = reinterpret_cast<uint4*>(table);
            if (st) {
                const uint4* g4 = reinterpret_cast<const uint4*>(st->table);
                for (int i = lane; i < kHashEntries / 4; i += 32) t4[i] = g4[i];
                offset = st->offset; dict_len = st->dict_len;
                dict_end = st->dict_buf + dict_len;
            } else {
                for (int i = lane; i < kHashEntries / 4; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
            }
            __syncwarp();
        }
        const uint8_t* last_src = nullptr; int last_n = -1;
        for (int b = b0; b < b1; b++) {
            const uint8_t* src = a.src + a.src_off[b];
            const int n = a.src_len[b];
            uint8_t* slot = a.dst + a.dst_off[b];
            int r = 0;
            if (n >= 0 && n <= kMaxInput) {                                                  // :1262
                const int bound = n + n / 255 + 16;
                const int cap = a.dst_cap ? (a.dst_cap[b] - a.header) : bound;
                if (offset + (uint32_t)n > 0x80000000u) {                                    // LZ4_renormDictT, :1545-1562
                    uint32_t delta = offset - 65536u;
                    for (int i = lane; i < kHashEntries; i += 32) { uint32_t v = table[i]; table[i] = v < delta ? 0u : v - delta; }
                    offset = 65536u;
                    if (dict_len > 65536u) dict_len = 65536u;
                    __syncwarp();
                }
                if (dict_len >= 1 && dict_len <= 3) dict_len = 0;                            // :1581-1587
                if (n == 0) {                                                                // :1263-1273
                    if (cap >= 1) { if (lane == 0) slot[a.header] = 0; r = 1; }
                } else if (cap > 0) {
                    BlockIn in{src, n, dict_end, dict_len, offset, slot + a.header, cap, accel};
                    offset += (uint32_t)n;                                                   // :918
                    r = encode_block(in, table);
                    __syncwarp();
                }
                dict_end = src + n; dict_len = (uint32_t)n;                                  // :1633-1634
                last_src = src; last_n = n;
            }
            if (lane == 0) {
                a.out_len[b] = r;
                if (a.header >= 4) { uint32_t v = (uint32_t)r; for (int k = 0; k < 4; k++) slot[k] = (uint8_t)(v >> (8 * k)); }       // LZ4.hs:262
                if (a.header == 8) { uint32_t v = (uint32_t)n; for (int k = 0; k < 4; k++) slot[4 + k] = (uint8_t)(v >> (8 * k)); }   // LZ4.hs:261
            }
        }
        if (st) {   // persist the stream (what the reference keeps in LZ4_stream_t + the live previous array)
            uint4* g4 = reinterpret_cast<uint4*>(st->table);
            const uint4* t4 = reinterpret_cast<const uint4*>(table);
            for (int i = lane; i < kHashEntries / 4; i += 32) g4[i] = t4[i];
            if (last_n >= 0) {
                uint32_t keep = (uint32_t)last_n <= st->dict_cap ? (uint32_t)last_n : 0u;   // host sizes dict_buf; 0 only on misuse
                if (keep) warp_copy_ro(st->dict_buf, last_src, keep);
                if (lane == 0) { st->offset = offset; st->dict_len = keep; }
            } else if (lane == 0) {
                st->offset = offset; st->dict_len = dict_len;
            }
        }
        __syncwarp();
    }
    // last CTA out resets the work counter so the scratch stays zeroed for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t done = atomicAdd(&a.scratch->work_counter[1], 1u);
        if (done == gridDim.x - 1) { a.scratch->work_counter[0] = 0; a.scratch->work_counter[1] = 0; __threadfence(); }
    }
}

}  // namespace

constexpr int kCompressWarps = 4;

cudaError_t launch_compress(const CompressArgs& a, cudaStream_t stream)
{
    static int sm_counts[64] = {0};     // per device: SM count, 0 = kernel not configured there yet
    const size_t smem = kCompressWarps * kHashEntries * sizeof(uint32_t);
    int dev = 0; cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!sm_counts[dev]) {
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(compress_kernel<kCompressWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        sm_counts[dev] = n;
    }
    const int sm_count = sm_counts[dev];
    if (a.n_streams <= 0) return cudaSuccess;
    int ctas_per_sm = 3;                         // 3 x 64 KiB of tables per SM
    int max_ctas = sm_count * ctas_per_sm;
    int want = (a.n_streams + kCompressWarps - 1) / kCompressWarps;
    int grid = want < max_ctas ? want : max_ctas;
    compress_kernel<kCompressWarps><<<grid, kCompressWarps * 32, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace b200lz4
