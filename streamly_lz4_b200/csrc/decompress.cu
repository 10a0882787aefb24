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
// overlapping matches (offset < length) handled as a periodic source.
#include "common.cuh"
#include "kernels.h"

namespace b200lz4 {

namespace {

constexpr int kWinBytes = 512;          // per-warp staging window
constexpr int kDecWarps = 4;

struct Window {
    const uint8_t* smem;        // this warp's window (kWinBytes)
    uint4* smem4;
    const uint8_t* src;         // payload
    int src_len;
    long long base;             // payload index held in smem[0] (16-byte aligned in global space)

    __device__ __forceinline__ void load(int ip)
    {
        const uint32_t lane = lane_id();
        uintptr_t g = reinterpret_cast<uintptr_t>(src + ip);
        base = (long long)ip - (long long)(g & 15);
        long long at = base + 16 * (long long)lane;
        __syncwarp();
        if (at < (long long)src_len)
            smem4[lane] = __ldg(reinterpret_cast<const uint4*>(src + at));
        __syncwarp();
    }
    __device__ __forceinline__ void ensure(int ip, int need)
    {
        if ((long long)ip + need > base + kWinBytes || (long long)ip < base) load(ip);
    }
    __device__ __forceinline__ uint32_t at(int ip) const { return smem[(long long)ip - base]; }
    __device__ __forceinline__ bool holds(int ip, uint32_t len) const
    { return (long long)ip >= base && (long long)ip + (long long)len <= base + kWinBytes; }
};

// Decode one block.  dict_end/dict_len: previous output of the stream (dict_len == 0: none).
__device__ int decode_block(Window& w, uint8_t* dst, int cap, const uint8_t* dict_end, uint32_t dict_len)
{
    const uint32_t lane = lane_id();
    const int src_len = w.src_len;
    const bool check_offset = dict_len < 65536u;                       // cbits/lz4.c:1764
    int ip = 0;
    long long op = 0;
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
                int p = ip + (int)lane;
                uint32_t s = (p < lim) ? w.at(p) : 0u;
                bool stop = (p >= lim - 1) || (s != 255u);             // reading position lim-1 ends the run (loop_error keeps the sum)
                uint32_t sb = __ballot_sync(kFull, stop);
                if (sb) {
                    int t = __ffs(sb) - 1;
                    uint32_t last = __shfl_sync(kFull, s, t);
                    len += 255u * (uint32_t)t + last;
                    ip += t + 1;
                    break;
                }
                len += 255u * 32u; ip += 32;
            }
        }
        // end rule, :1991-2047
        if (op + (long long)len > (long long)cap - kMfLimit || (long long)ip + (long long)len > (long long)src_len - (2 + 1 + kLastLiterals)) {
            if ((long long)ip + (long long)len != (long long)src_len || op + (long long)len > (long long)cap) return -ip - 1;
            if (w.holds(ip, len)) { for (uint32_t i = lane; i < len; i += 32) dst[op + i] = (uint8_t)w.at(ip + (int)i); }
            else warp_copy_ro(dst + op, w.src + ip, len);
            op += len;
            break;
        }
        if (len) {
            if (w.holds(ip, len)) { for (uint32_t i = lane; i < len; i += 32) dst[op + i] = (uint8_t)w.at(ip + (int)i); }
            else warp_copy_ro(dst + op, w.src + ip, len);
            ip += (int)len; op += len;
        }
        w.ensure(ip, 34);
        const uint32_t dist = w.at(ip) | (w.at(ip + 1) << 8); ip += 2;  // :2055
        uint32_t mlen = token & 15;
        if (mlen == 15) {                                              // :2062-2067
            const int lim = src_len - kLastLiterals + 1;
            for (;;) {
                w.ensure(ip, 32);
                int p = ip + (int)lane;
                uint32_t s = (p < src_len) ? w.at(p) : 0u;
                bool stop = (p >= src_len) || (s != 255u);
                uint32_t sb = __ballot_sync(kFull, stop);
                if (sb) {
                    int t = __ffs(sb) - 1;
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
        const long long from = op - (long long)dist;
        if (check_offset && from + (long long)dict_len < 0) return -ip - 1;      // :2073
        if (op + (long long)mlen > (long long)cap - kLastLiterals) return -ip - 1; // :2076-2078, :2139
        if (dist == 0) return -ip - 1;     // format violation (the reference copies garbage here, :2122-2130)
        __syncwarp();                      // earlier stores of this warp are ordered before the loads below
        uint8_t* out = dst + op;
        if (from >= 0) {
            const uint8_t* m = dst + from;
            if (dist >= mlen) {                                // no overlap
                for (uint32_t i = lane; i < mlen; i += 32) out[i] = m[i];
            } else if (dist >= 32) {                           // overlap, period >= one round
                for (uint32_t i0 = 0; i0 < mlen; i0 += 32) {
                    uint32_t i = i0 + lane;
                    if (i < mlen) out[i] = m[i];
                    __syncwarp();
                }
            } else {                                           // short period: source is dist bytes repeated
                uint32_t k = lane % dist, adv = 32 % dist;
                for (uint32_t i = lane; i < mlen; i += 32) {
                    out[i] = m[k];
                    k += adv; if (k >= dist) k -= dist;
                }
            }
        } else if (dist >= 32) {                               // starts in the previous output (:2075-2100)
            for (uint32_t i0 = 0; i0 < mlen; i0 += 32) {
                uint32_t i = i0 + lane;
                if (i < mlen) { long long f = from + i; out[i] = (f < 0) ? dict_end[f] : dst[f]; }
                __syncwarp();
            }
        } else {                                               // dist < 32 and op < dist: a handful of bytes, serial
            if (lane == 0) for (uint32_t i = 0; i < mlen; i++) { long long f = from + i; out[i] = (f < 0) ? dict_end[f] : dst[f]; }
        }
        op += mlen;
    }
    return (int)op;
}

__device__ __forceinline__ int read_le32(const uint8_t* p)
{ return (int)((uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24)); }

__global__ void __launch_bounds__(kDecWarps * 32)
decompress_kernel(DecompressArgs a)
{
    __shared__ uint4 windows[kDecWarps][kWinBytes / 16];
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    uint32_t* counter = &a.scratch->work_counter[2];
    Window w;
    w.smem4 = windows[warp];
    w.smem = reinterpret_cast<const uint8_t*>(windows[warp]);

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
                    r = decode_block(w, out, cap, dict_end, dict_len);
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
