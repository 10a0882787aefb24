// decompress.cu -- LZ4 block decoder for sm_100a.
//
// Semantics: LZ4_decompress_safe_continue restricted to non-adjacent outputs
// (cbits/lz4.c:2322-2359), i.e. LZ4_decompress_generic(endOnInputSize,
// decode_full_block, {noDict | usingExtDict with the previous output})
// (cbits/lz4.c:1737-2165; the safe loop :1929-2151 defines the accept/reject rules).
//
// Organisation: one WARP per stream (independent mode: one block).  The compressed
// bytes are staged through a per-warp shared-memory window filled with coalesced
// 128-bit loads; tokens / offsets / length bytes are parsed from shared memory with
// warp-uniform control flow (length-extension runs are summed 32 bytes at a time with
// a ballot), literals and matches are copied cooperatively by the 32 lanes, with
// overlapping matches (offset < length) handled as a periodic source.  The last KiB of
// output is mirrored in a shared-memory ring so that matches with small offsets are served
// from shared memory rather than through an L2 round trip to bytes just stored.
#include "common.cuh"
#include "kernels.h"

namespace b200lz4 {

namespace {

constexpr int kWinBytes = 512;          // per-warp staging window of the compressed stream
constexpr int kRingBytes = 1024;        // per-warp ring of the most recent output bytes
constexpr uint32_t kRingReach = kRingBytes - 64;   // matches with offset <= this read the ring
constexpr int kDecWarps = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v)
{ asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

// Staging window over the compressed payload: 512 bytes of it live in shared memory.
struct Window {
    uint32_t win_s;             // shared address of the window
    const uint8_t* src;         // payload
    int src_len;
    int base;                   // payload index held at win_s (16-byte aligned in global space; may be slightly negative)

    __device__ __forceinline__ void load(int ip)
    {
        const uint32_t lane = lane_id();
        base = ip - (int)(reinterpret_cast<uintptr_t>(src + ip) & 15);
        const int at = base + 16 * (int)lane;
        __syncwarp();
        if (at < src_len) sts128(win_s + 16 * lane, __ldg(reinterpret_cast<const uint4*>(src + at)));
        __syncwarp();
    }
    // make [ip, ip + need) readable (need <= 64)
    __device__ __forceinline__ void ensure(int ip, int need)
    {
        if ((uint32_t)(ip - base) > (uint32_t)(kWinBytes - need)) load(ip);
    }
    __device__ __forceinline__ uint32_t at(int ip) const { return lds8(win_s + (uint32_t)(ip - base)); }
    __device__ __forceinline__ bool holds(int ip, uint32_t len) const
    { return (uint32_t)(ip - base) + len <= (uint32_t)kWinBytes; }
};

// Decode one block.  dict_end/dict_len: previous output of the stream (dict_len == 0: none).
// ring_s: shared address of this warp's output ring; ring[q & (kRingBytes-1)] mirrors dst[q] for
// ring_lo <= q < op, so that matches with small offsets (the common case on compressible data) are
// served from shared memory instead of an L2 round trip to bytes this warp has just stored.
__device__ int decode_block(Window& w, const uint32_t ring_s, uint8_t* dst, int cap, const uint8_t* dict_end, uint32_t dict_len)
{
    const uint32_t lane = lane_id();
    const int src_len = w.src_len;
    const bool check_offset = dict_len < 65536u;                       // cbits/lz4.c:1764
    int ip = 0, op = 0;
    int ring_lo = 0;
    if (cap == 0) {                                                    // :1781-1785
        if (src_len != 1) return -1;
        w.load(0);
        return w.at(0) == 0 ? 0 : -1;
    }
    if (src_len == 0) return -1;                                       // :1787
    w.load(0);
    for (;;) {
        w.ensure(ip, 64);
        const uint32_t token = w.at(ip); ip++;
        uint32_t len = token >> 4;
        if (len == 15) {                                               // :1977-1983 with reader :1707-1729
            const int lim = src_len - 15;
            if (ip >= lim) return -ip - 1;                             // initial_error
            for (;;) {
                w.ensure(ip, 32);
                const int p = ip + (int)lane;
                const uint32_t s = (p < lim) ? w.at(p) : 0u;
                const bool stop = (p >= lim - 1) || (s != 255u);       // reading position lim-1 ends the run (loop_error keeps the sum)
                const uint32_t sb = __ballot_sync(kFull, stop);
                if (sb) {
                    const int t = __ffs(sb) - 1;
                    len += 255u * (uint32_t)t + __shfl_sync(kFull, s, t);
                    ip += t + 1;
                    break;
                }
                len += 255u * 32u; ip += 32;
            }
            if (len > 0x7FFFFFFFu) return -ip - 1;                     // cannot be a valid run; keeps the int arithmetic below exact
        }
        // end rule, :1991-2047
        const bool last = ((long long)op + len > (long long)cap - kMfLimit) || ((long long)ip + len > (long long)src_len - (2 + 1 + kLastLiterals));
        if (last && ((long long)ip + len != (long long)src_len || (long long)op + len > (long long)cap)) return -ip - 1;
        if (len) {
            if (len <= 64 && w.holds(ip, len)) {
                for (uint32_t i = lane; i < len; i += 32) {
                    const uint32_t b = w.at(ip + (int)i);
                    dst[op + (int)i] = (uint8_t)b;
                    sts8(ring_s + ((uint32_t)(op + (int)i) & (kRingBytes - 1)), b);
                }
            } else {
                warp_copy_ro(dst + op, w.src + ip, len);
                ring_lo = op + (int)len;                               // the ring does not mirror this run
            }
            ip += (int)len; op += (int)len;
        }
        if (last) break;
        w.ensure(ip, 34);
        const uint32_t dist = w.at(ip) | (w.at(ip + 1) << 8); ip += 2;  // :2055
        uint32_t mlen = token & 15;
        if (mlen == 15) {                                              // :2062-2067
            const int lim = src_len - kLastLiterals + 1;
            for (;;) {
                w.ensure(ip, 32);
                const int p = ip + (int)lane;
                const uint32_t s = (p < src_len) ? w.at(p) : 0u;
                const bool stop = (p >= src_len) || (s != 255u);
                const uint32_t sb = __ballot_sync(kFull, stop);
                if (sb) {
                    const int t = __ffs(sb) - 1;
                    if (ip + t + 1 >= lim) return -(ip + t + 1) - 1;   // loop_error
                    mlen += 255u * (uint32_t)t + __shfl_sync(kFull, s, t);
                    ip += t + 1;
                    break;
                }
                if (ip + 32 >= lim) return -(ip + 32) - 1;
                mlen += 255u * 32u; ip += 32;
            }
        }
        mlen += kMinMatch;
        const int from = op - (int)dist;
        if (check_offset && (long long)from + (long long)dict_len < 0) return -ip - 1;   // :2073
        if ((long long)op + mlen > (long long)cap - kLastLiterals) return -ip - 1;        // :2076-2078, :2139
        if (dist == 0) return -ip - 1;     // format violation (the reference copies garbage here, :2122-2130)
        __syncwarp();                      // earlier stores of this warp (global and ring) are ordered before the loads below
        uint8_t* out = dst + op;
        if (from >= ring_lo && dist <= kRingReach) {
            // ---- source is in the ring
            const uint32_t rsrc = (uint32_t)from, rdst = (uint32_t)op;
            if (dist >= 32) {
                for (uint32_t i0 = 0; i0 < mlen; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    if (i < mlen) {
                        const uint32_t b = lds8(ring_s + ((rsrc + i) & (kRingBytes - 1)));
                        out[i] = (uint8_t)b;
                        sts8(ring_s + ((rdst + i) & (kRingBytes - 1)), b);
                    }
                    if (dist < mlen) __syncwarp();             // later rounds read what this round wrote
                }
            } else {                                           // short period: the source is dist bytes repeated
                // the pattern is taken into registers first: a long run would overwrite its ring slots
                const uint32_t pat = lds8(ring_s + ((rsrc + (lane < dist ? lane : 0u)) & (kRingBytes - 1)));
                uint32_t k = lane % dist;
                const uint32_t adv = 32 % dist;
                for (uint32_t i0 = 0; i0 < mlen; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    const uint32_t b = __shfl_sync(kFull, pat, k);
                    if (i < mlen) {
                        out[i] = (uint8_t)b;
                        sts8(ring_s + ((rdst + i) & (kRingBytes - 1)), b);
                    }
                    k += adv; if (k >= dist) k -= dist;
                }
            }
        } else if (from >= 0) {
            // ---- source in this block's output, read back through L2
            const uint8_t* m = dst + from;
            if (dist >= 32) {
                for (uint32_t i0 = 0; i0 < mlen; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    if (i < mlen) {
                        const uint32_t b = m[i];
                        out[i] = (uint8_t)b;
                        sts8(ring_s + ((uint32_t)(op + (int)i) & (kRingBytes - 1)), b);
                    }
                    if (dist < mlen) __syncwarp();
                }
            } else {
                uint32_t k = lane % dist;
                const uint32_t adv = 32 % dist;
                for (uint32_t i = lane; i < mlen; i += 32) {
                    const uint32_t b = m[k];
                    out[i] = (uint8_t)b;
                    sts8(ring_s + ((uint32_t)(op + (int)i) & (kRingBytes - 1)), b);
                    k += adv; if (k >= dist) k -= dist;
                }
            }
        } else if (dist >= 32) {                               // starts in the previous output (:2075-2100)
            for (uint32_t i0 = 0; i0 < mlen; i0 += 32) {
                const uint32_t i = i0 + lane;
                if (i < mlen) {
                    const int f = from + (int)i;
                    const uint32_t b = (f < 0) ? dict_end[f] : dst[f];
                    out[i] = (uint8_t)b;
                    sts8(ring_s + ((uint32_t)(op + (int)i) & (kRingBytes - 1)), b);
                }
                __syncwarp();
            }
        } else {                                               // dist < 32 and op < dist: a handful of bytes, serial
            if (lane == 0) for (uint32_t i = 0; i < mlen; i++) {
                const int f = from + (int)i;
                const uint32_t b = (f < 0) ? dict_end[f] : dst[f];
                out[i] = (uint8_t)b;
                sts8(ring_s + ((uint32_t)(op + (int)i) & (kRingBytes - 1)), b);
            }
        }
        op += (int)mlen;
    }
    return op;
}

__device__ __forceinline__ int read_le32(const uint8_t* p)
{ return (int)((uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24)); }

__global__ void __launch_bounds__(kDecWarps * 32)
decompress_kernel(DecompressArgs a)
{
    __shared__ uint4 windows[kDecWarps][kWinBytes / 16];
    __shared__ uint4 rings[kDecWarps][kRingBytes / 16];
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    uint32_t* counter = &a.scratch->work_counter[2];
    Window w;
    w.win_s = smem_u32(windows[warp]);
    const uint32_t ring_s = smem_u32(rings[warp]);

    for (;;) {
        int s = 0;
        if (lane == 0) s = (int)atomicAdd(counter, 1u);
        s = __shfl_sync(kFull, s, 0);
        if (s >= a.n_streams) break;
        const int b0 = a.stream_first ? a.stream_first[s] : a.first_block + s;
        const int b1 = a.stream_first ? a.stream_first[s + 1] : b0 + 1;
        DState* st = a.states ? reinterpret_cast<DState*>(a.states[s]) : nullptr;
        const uint8_t* dict_end = nullptr; uint32_t dict_len = 0;
        if (st && st->prev_len) { dict_len = st->prev_len; dict_end = st->tail + 65536; }
        const uint8_t* last_out = nullptr; int last_len = 0;
        for (int b = b0; b < b1; b++) {
            const uint8_t* arr = a.src + a.src_off[b];
            const int alen = a.src_len[b];
            const int avail = alen - a.header;
            uint8_t* out = a.dst + a.dst_off[b];
            int r = -1;
            if (avail >= 0) {
                int comp_len = a.header >= 4 ? read_le32(arr) : avail;                 // LZ4.hs:303
                int cap = a.header == 8 ? read_le32(arr + 4) : a.max_block;            // LZ4.hs:189-198
                bool ok = comp_len > 0 && comp_len == avail && cap >= 0;               // LZ4.hs:309-318 (array length bounds the read)
                if (a.dst_cap && cap > a.dst_cap[b]) ok = false;
                if (ok) {
                    w.src = arr + a.header; w.src_len = comp_len; w.base = 0;
                    r = decode_block(w, ring_s, out, cap, dict_end, dict_len);
                    __syncwarp();
                }
            }
            if (lane == 0) a.out_len[b] = r;
            if (r > 0) { dict_end = out + r; dict_len = (uint32_t)r; last_out = out; last_len = r; }   // cbits/lz4.c:2353-2355
        }
        if (st && last_out) {           // keep the reachable tail of the last output for the next call
            uint32_t kept = last_len < 65536 ? (uint32_t)last_len : 65536u;
            const uint8_t* from = last_out + last_len - kept;
            uint8_t* to = st->tail + 65536 - kept;
            __syncwarp();
            for (uint32_t i = lane; i < kept; i += 32) to[i] = from[i];
            if (lane == 0) { st->prev_len = (uint32_t)last_len; st->kept = kept; }
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t done = atomicAdd(&a.scratch->work_counter[3], 1u);
        if (done == gridDim.x - 1) { a.scratch->work_counter[2] = 0; a.scratch->work_counter[3] = 0; __threadfence(); }
    }
}

}  // namespace

cudaError_t launch_decompress(const DecompressArgs& a, cudaStream_t stream)
{
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0; cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
    }
    if (a.n_streams <= 0) return cudaSuccess;
    int max_ctas = sm_count * 16;
    int want = (a.n_streams + kDecWarps - 1) / kDecWarps;
    int grid = want < max_ctas ? want : max_ctas;
    decompress_kernel<<<grid, kDecWarps * 32, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace b200lz4
