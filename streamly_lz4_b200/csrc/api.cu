// api.cu -- the C ABI of libb200lz4.so (see include/b200lz4.h): contexts, staging,
// the batched host/device entry points, re-framing and the legacy LZ4_* aliases.
// There is no CPU codec in this library: every compress/decompress call runs the
// sm_100a kernels, and fails with B200LZ4_E_CUDA when no such device is usable.
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <chrono>
#include <vector>

#include "b200lz4.h"
#include "kernels.h"

using namespace b200lz4;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }
int fail_cuda(cudaError_t e, const char* what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return B200LZ4_E_CUDA;
}
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail_cuda(e__, #call); } while (0)

inline int bound_of(int n) { return ((unsigned)n > (unsigned)kMaxInput) ? 0 : n + n / 255 + 16; }
inline int64_t align16(int64_t v) { return (v + 15) & ~int64_t(15); }

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail(B200LZ4_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocMapped | cudaHostAllocPortable);    // device-visible: kernels mirror sizes into it
        if (e != cudaSuccess) { p = nullptr; return fail(B200LZ4_E_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

constexpr int kMaxChunks = 6;           // pipeline depth of the host batch calls (H2D + D2H + chunk streams must fit the
                                        // 8 hardware queues of the default CUDA_DEVICE_MAX_CONNECTIONS, else chunks serialize)
constexpr int kKernelStreams = kMaxChunks;   // one stream per chunk: chunk kernels must be able to overlap

struct b200lz4_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;                  // H2D + bookkeeping
    cudaStream_t kstream[kKernelStreams] = {};      // chunk kernels (run concurrently)
    cudaStream_t dstream = nullptr;                 // D2H
    Scratch* scratch = nullptr;                     // kMaxChunks records: one work counter per in-flight kernel
    DevBuf d_src, d_slots, d_out, d_desc, d_wide;   // d_wide: descriptor rings of decompress_kernel_wide (one slice per chunk)
    PinBuf h_desc;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_h2d[kMaxChunks] = {}, ev_k[kMaxChunks] = {}, ev_k0 = nullptr, ev_d0 = nullptr, ev_d1 = nullptr;
    float t_h2d = 0, t_kernel = 0, t_d2h = 0;
    int64_t launches = 0;
    std::string err;                                // text of the last failure of a call on this ctx (any thread)
};

struct b200lz4_cstream {
    b200lz4_ctx* ctx; CState* d_state; uint8_t* d_dict; uint32_t dict_cap; uint32_t last_len;
};
struct b200lz4_dstream {
    b200lz4_ctx* ctx; DState* d_state;
};

namespace {

// carve typed arrays out of one host/device descriptor block
struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) { size_t o = off; off = (off + bytes + 15) & ~size_t(15); return o; }
};

int check_blocks(const int64_t* off, const int32_t* len, int n, int64_t bytes)
{
    for (int i = 0; i < n; i++) {
        if (len[i] < 0 || off[i] < 0 || off[i] + (int64_t)len[i] > bytes)
            return fail(B200LZ4_E_ARG, "block " + std::to_string(i) + " lies outside the source buffer");
    }
    return 0;
}
int check_streams(const int32_t* first, int n_streams, int n_blocks)
{
    if (!first) return 0;
    if (n_streams < 0 || (n_streams == 0 && n_blocks != 0)) return fail(B200LZ4_E_ARG, "n_streams");
    if (n_streams == 0) return 0;
    if (first[0] != 0 || first[n_streams] != n_blocks) return fail(B200LZ4_E_ARG, "stream_first must start at 0 and end at n_blocks");
    for (int s = 0; s < n_streams; s++) if (first[s] > first[s + 1]) return fail(B200LZ4_E_ARG, "stream_first must be non-decreasing");
    return 0;
}

int set_dict_fields(CState* d_state, uint8_t* buf, uint32_t cap, cudaStream_t st)
{
    struct { uint32_t dict_cap; uint32_t pad; uint8_t* dict_buf; } f{cap, 0, buf};
    static_assert(offsetof(CState, dict_buf) == offsetof(CState, dict_cap) + 8, "CState layout");
    CU(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(d_state) + offsetof(CState, dict_cap), &f, sizeof f, cudaMemcpyHostToDevice, st));
    return 0;
}

// grow a stream's previous-array buffer, preserving its content
int cstream_reserve(b200lz4_cstream* s, uint32_t need)
{
    if (need <= s->dict_cap) return 0;
    b200lz4_ctx* c = s->ctx;
    uint32_t cap = need < 65536u ? 65536u : need;
    uint8_t* nb = nullptr;
    cudaError_t e = cudaMalloc(&nb, (size_t)cap + 16);
    if (e != cudaSuccess) return fail(B200LZ4_E_NOMEM, std::string("cudaMalloc(dict): ") + cudaGetErrorString(e));
    if (s->d_dict && s->last_len) CU(cudaMemcpyAsync(nb, s->d_dict, s->last_len, cudaMemcpyDeviceToDevice, c->stream));
    int rc = set_dict_fields(s->d_state, nb, cap, c->stream);
    if (rc) { cudaFree(nb); return rc; }
    CU(cudaStreamSynchronize(c->stream));
    if (s->d_dict) cudaFree(s->d_dict);
    s->d_dict = nb; s->dict_cap = cap;
    return 0;
}

int sync_all(b200lz4_ctx* c)
{
    CU(cudaStreamSynchronize(c->stream));
    for (auto& k : c->kstream) CU(cudaStreamSynchronize(k));
    CU(cudaStreamSynchronize(c->dstream));
    return 0;
}

// ---- pipeline planning ------------------------------------------------------
// A host batch is cut into up to kMaxChunks chunks of whole streams (independent mode: whole
// blocks).  Chunk k's H2D copy, its kernels and its D2H copy run on three different CUDA
// streams, so that transfers in both directions overlap the kernels of neighbouring chunks;
// the chunk kernels themselves run concurrently on kKernelStreams streams because a codec
// kernel is bound by per-block latency, not by the number of blocks it is given.
struct Chunk { int b0, b1, s0, s1; int64_t lo, hi; };

int plan_chunks(const int64_t* off, const int32_t* len, int n, const int32_t* first, int ns, Chunk* out)
{
    int64_t total = 0;
    for (int i = 0; i < n; i++) total += len[i];
    // Chunk k becomes ready when its H2D copy lands and finishes one block latency later, so the end of the
    // call is "last byte arrives + block latency + D2H of the last chunk": keep the last chunk small.
    static const double kShare[kMaxChunks] = {0.14, 0.20, 0.20, 0.20, 0.18, 0.08};
    const bool small = total < (int64_t(48) << 20);
    int k = 0, unit = 0;
    const int units = first ? ns : n;
    int64_t done = 0;
    while (unit < units) {
        Chunk c; c.s0 = unit; c.b0 = first ? first[unit] : unit;
        double upto = 0;
        for (int q = 0; q <= k; q++) upto += kShare[q];
        const int64_t target = (small || k == kMaxChunks - 1) ? total : (int64_t)(upto * (double)total);
        while (unit < units && (done < target || k == kMaxChunks - 1 || small)) {
            const int u0 = first ? first[unit] : unit, u1 = first ? first[unit + 1] : unit + 1;
            for (int i = u0; i < u1; i++) done += len[i];
            unit++;
        }
        c.s1 = unit; c.b1 = first ? first[unit] : unit;
        c.lo = INT64_MAX; c.hi = 0;
        for (int i = c.b0; i < c.b1; i++) { c.lo = std::min(c.lo, off[i]); c.hi = std::max(c.hi, off[i] + (int64_t)len[i]); }
        if (c.lo > c.hi) c.lo = c.hi = 0;
        out[k++] = c;
    }
    return k;
}

int compress_host(b200lz4_ctx* c, const void* src, int64_t src_bytes,
                  const int64_t* src_off, const int32_t* src_len, int n,
                  const int32_t* stream_first, int n_streams, b200lz4_cstream* const* streams,
                  int accel, int header, const int32_t* block_cap,
                  void* dst, int64_t dst_cap, int64_t* dst_off, int32_t* out_len)
{
    if (!c) return fail(B200LZ4_E_ARG, "ctx is NULL");
    if (n < 0 || (n > 0 && (!src_off || !src_len || !dst_off || !out_len))) return fail(B200LZ4_E_ARG, "NULL descriptor array");
    if (header != 0 && header != 4 && header != 8) return fail(B200LZ4_E_ARG, "header_mode must be 0, 4 or 8");
    if (streams && !stream_first) return fail(B200LZ4_E_ARG, "streams given without stream_first");
    if (n == 0) { if (dst_off) dst_off[0] = 0; return 0; }
    if (!src || !dst) return fail(B200LZ4_E_ARG, "NULL data buffer");
    int rc;
    if ((rc = check_blocks(src_off, src_len, n, src_bytes))) return rc;
    if ((rc = check_streams(stream_first, n_streams, n))) return rc;
    CU(cudaSetDevice(c->device));
    const int ns = stream_first ? n_streams : n;

    // stream state: make sure each persistent stream can keep its last array
    if (streams) {
        for (int s = 0; s < n_streams; s++) {
            if (!streams[s] || streams[s]->ctx != c) return fail(B200LZ4_E_ARG, "stream handle belongs to another ctx");
            if (stream_first[s + 1] > stream_first[s]) {
                uint32_t need = (uint32_t)src_len[stream_first[s + 1] - 1];
                if ((rc = cstream_reserve(streams[s], need))) return rc;
            }
        }
    }
    Chunk chunks[kMaxChunks];
    const int nchunks = plan_chunks(src_off, src_len, n, stream_first, ns, chunks);

    // descriptor block
    Carver cv;
    const size_t o_src_off = cv.take(sizeof(int64_t) * n);
    const size_t o_slot_off = cv.take(sizeof(int64_t) * n);
    const size_t o_src_len = cv.take(sizeof(int32_t) * n);
    const size_t o_cap = cv.take(sizeof(int32_t) * n);
    const size_t o_first = cv.take(sizeof(int32_t) * (ns + 1));
    const size_t o_states = cv.take(sizeof(void*) * ns);
    const size_t o_upload_end = cv.off;
    const size_t o_out_len = cv.take(sizeof(int32_t) * n);
    const size_t o_out_off = cv.take(sizeof(int64_t) * (n + nchunks));
    const size_t desc_bytes = cv.off;
    if ((rc = c->h_desc.ensure(desc_bytes))) return rc;
    if ((rc = c->d_desc.ensure(desc_bytes))) return rc;
    uint8_t* hd = static_cast<uint8_t*>(c->h_desc.p);
    uint8_t* dd = static_cast<uint8_t*>(c->d_desc.p);
    int64_t* h_slot_off = reinterpret_cast<int64_t*>(hd + o_slot_off);
    int32_t* h_cap = reinterpret_cast<int32_t*>(hd + o_cap);
    memcpy(hd + o_src_off, src_off, sizeof(int64_t) * n);
    memcpy(hd + o_src_len, src_len, sizeof(int32_t) * n);
    int64_t slots_total = 0;
    for (int i = 0; i < n; i++) {
        int cap = block_cap ? block_cap[i] : bound_of(src_len[i]);
        if (cap < 0) cap = 0;
        h_cap[i] = cap + header;
        h_slot_off[i] = slots_total;
        slots_total += align16((int64_t)cap + header + 16);
    }
    if (stream_first) memcpy(hd + o_first, stream_first, sizeof(int32_t) * (ns + 1));
    if (streams) { void** hs = reinterpret_cast<void**>(hd + o_states); for (int s = 0; s < ns; s++) hs[s] = streams[s]->d_state; }

    if ((rc = c->d_src.ensure((size_t)src_bytes + 64))) return rc;
    if ((rc = c->d_slots.ensure((size_t)slots_total + 64))) return rc;
    if ((rc = c->d_out.ensure((size_t)slots_total + 64))) return rc;

    const uint8_t* hsrc = static_cast<const uint8_t*>(src);
    uint8_t* d_src = static_cast<uint8_t*>(c->d_src.p);
    uint8_t* d_slots = static_cast<uint8_t*>(c->d_slots.p);
    uint8_t* d_out = static_cast<uint8_t*>(c->d_out.p);
    int64_t* d_out_off = reinterpret_cast<int64_t*>(dd + o_out_off);
    int32_t* d_out_len = reinterpret_cast<int32_t*>(dd + o_out_len);

    cudaStream_t st = c->stream;
    CU(cudaEventRecord(c->ev[0], st));
    CU(cudaMemcpyAsync(dd, hd, o_upload_end, cudaMemcpyHostToDevice, st));
    for (int k = 0; k < nchunks; k++) {
        const Chunk& ch = chunks[k];
        if (ch.hi > ch.lo) CU(cudaMemcpyAsync(d_src + ch.lo, hsrc + ch.lo, (size_t)(ch.hi - ch.lo), cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(c->ev_h2d[k], st));
        cudaStream_t ks = c->kstream[k % kKernelStreams];
        CU(cudaStreamWaitEvent(ks, c->ev_h2d[k], 0));
        if (k == 0) CU(cudaEventRecord(c->ev_k0, ks));
        CompressArgs a{};
        a.src = d_src;
        a.src_off = reinterpret_cast<const int64_t*>(dd + o_src_off);
        a.src_len = reinterpret_cast<const int32_t*>(dd + o_src_len);
        a.n_blocks = n;
        a.first_block = ch.b0;
        a.stream_first = stream_first ? reinterpret_cast<const int32_t*>(dd + o_first) + ch.s0 : nullptr;
        a.n_streams = ch.s1 - ch.s0;
        a.states = streams ? reinterpret_cast<void* const*>(dd + o_states) + ch.s0 : nullptr;
        a.dst = d_slots;
        a.dst_off = reinterpret_cast<const int64_t*>(dd + o_slot_off);
        a.dst_cap = block_cap ? reinterpret_cast<const int32_t*>(dd + o_cap) : nullptr;
        a.out_len = d_out_len;
        a.accel = accel; a.header = header; a.scratch = c->scratch + k;
        CU(launch_compress(a, ks));
        const int nk = ch.b1 - ch.b0;
        // the scan kernel mirrors sizes and offsets into the pinned descriptor block (device-visible host memory)
        CompactArgs g{d_slots, a.dst_off + ch.b0, d_out_len + ch.b0, nk, header, d_out + h_slot_off[ch.b0], d_out_off + ch.b0 + k,
                      reinterpret_cast<int64_t*>(hd + o_out_off) + ch.b0 + k, reinterpret_cast<int32_t*>(hd + o_out_len) + ch.b0};
        CU(launch_compact(g, ks));
        c->launches += kernel_launches_per_compress() + kernel_launches_per_compact();
        if (k == nchunks - 1) CU(cudaEventRecord(c->ev[2], ks));
        CU(cudaEventRecord(c->ev_k[k], ks));
    }
    CU(cudaEventRecord(c->ev[1], st));

    // drain: as each chunk's sizes arrive, send its compacted bytes home
    const int32_t* h_out_len = reinterpret_cast<const int32_t*>(hd + o_out_len);
    const int64_t* h_out_off = reinterpret_cast<const int64_t*>(hd + o_out_off);
    int64_t base = 0;
    int result = 0;
    CU(cudaEventRecord(c->ev_d0, c->dstream));
    for (int k = 0; k < nchunks; k++) {
        const Chunk& ch = chunks[k];
        CU(cudaEventSynchronize(c->ev_k[k]));
        const int nk = ch.b1 - ch.b0;
        const int64_t* off_k = h_out_off + ch.b0 + k;
        const int64_t total_k = off_k[nk];
        for (int i = 0; i < nk; i++) dst_off[ch.b0 + i] = base + off_k[i];
        if (base + total_k > dst_cap) { result = fail(B200LZ4_E_NOMEM, "dst_cap too small"); break; }
        if (total_k) CU(cudaMemcpyAsync(static_cast<uint8_t*>(dst) + base, d_out + h_slot_off[ch.b0], (size_t)total_k,
                                        cudaMemcpyDeviceToHost, c->dstream));
        base += total_k;
    }
    dst_off[n] = base;
    CU(cudaEventRecord(c->ev_d1, c->dstream));
    if ((rc = sync_all(c))) return rc;
    if (result) return result;
    memcpy(out_len, h_out_len, sizeof(int32_t) * n);
    cudaEventElapsedTime(&c->t_h2d, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&c->t_kernel, c->ev_k0, c->ev[2]);
    cudaEventElapsedTime(&c->t_d2h, c->ev_d0, c->ev_d1);
    if (getenv("B200LZ4_DEBUG")) {          // device timeline of the pipeline, ms since the call's first event
        float t = 0;
        for (int k = 0; k < nchunks; k++) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, c->ev[0], c->ev_h2d[k]); cudaEventElapsedTime(&b, c->ev[0], c->ev_k[k]);
            fprintf(stderr, "[b200lz4] chunk %d: blocks %d..%d  h2d done %.2f  kernels done %.2f\n", k, chunks[k].b0, chunks[k].b1, a, b);
        }
        cudaEventElapsedTime(&t, c->ev[0], c->ev_d1);
        fprintf(stderr, "[b200lz4] d2h done %.2f\n", t);
    }
    if (streams) for (int s = 0; s < n_streams; s++)
        if (stream_first[s + 1] > stream_first[s]) streams[s]->last_len = (uint32_t)src_len[stream_first[s + 1] - 1];
    for (int i = 0; i < n; i++) if (out_len[i] <= 0) return fail(B200LZ4_E_BLOCK, "block " + std::to_string(i) + " failed to compress");
    return 0;
}

inline int32_t le32(const uint8_t* p)
{ return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24)); }

int decompress_host(b200lz4_ctx* c, const void* src, int64_t src_bytes,
                    const int64_t* src_off, const int32_t* src_len, int n,
                    const int32_t* stream_first, int n_streams, b200lz4_dstream* const* streams,
                    int header, int max_block,
                    void* dst, int64_t dst_cap, int64_t* dst_off, int32_t* out_len)
{
    if (!c) return fail(B200LZ4_E_ARG, "ctx is NULL");
    if (n < 0 || (n > 0 && (!src_off || !src_len || !dst_off || !out_len))) return fail(B200LZ4_E_ARG, "NULL descriptor array");
    if (header != 0 && header != 4 && header != 8) return fail(B200LZ4_E_ARG, "header_mode must be 0, 4 or 8");
    if (header != 8 && max_block < 0) return fail(B200LZ4_E_ARG, "max_block");
    if (streams && !stream_first) return fail(B200LZ4_E_ARG, "streams given without stream_first");
    if (n == 0) { if (dst_off) dst_off[0] = 0; return 0; }
    if (!src || !dst) return fail(B200LZ4_E_ARG, "NULL data buffer");
    int rc;
    if ((rc = check_blocks(src_off, src_len, n, src_bytes))) return rc;
    if ((rc = check_streams(stream_first, n_streams, n))) return rc;
    CU(cudaSetDevice(c->device));
    const int ns = stream_first ? n_streams : n;
    if (streams) for (int s = 0; s < n_streams; s++)
        if (!streams[s] || streams[s]->ctx != c) return fail(B200LZ4_E_ARG, "stream handle belongs to another ctx");
    Chunk chunks[kMaxChunks];
    const int nchunks = plan_chunks(src_off, src_len, n, stream_first, ns, chunks);

    Carver cv;
    const size_t o_src_off = cv.take(sizeof(int64_t) * n);
    const size_t o_slot_off = cv.take(sizeof(int64_t) * n);
    const size_t o_src_len = cv.take(sizeof(int32_t) * n);
    const size_t o_cap = cv.take(sizeof(int32_t) * n);
    const size_t o_first = cv.take(sizeof(int32_t) * (ns + 1));
    const size_t o_states = cv.take(sizeof(void*) * ns);
    const size_t o_upload_end = cv.off;
    const size_t o_out_len = cv.take(sizeof(int32_t) * n);
    const size_t o_out_off = cv.take(sizeof(int64_t) * (n + nchunks));
    const size_t desc_bytes = cv.off;
    if ((rc = c->h_desc.ensure(desc_bytes))) return rc;
    if ((rc = c->d_desc.ensure(desc_bytes))) return rc;
    uint8_t* hd = static_cast<uint8_t*>(c->h_desc.p);
    uint8_t* dd = static_cast<uint8_t*>(c->d_desc.p);
    int64_t* h_slot_off = reinterpret_cast<int64_t*>(hd + o_slot_off);
    int32_t* h_cap = reinterpret_cast<int32_t*>(hd + o_cap);
    memcpy(hd + o_src_off, src_off, sizeof(int64_t) * n);
    memcpy(hd + o_src_len, src_len, sizeof(int32_t) * n);
    // capacities: the header's uncompLen (BlockHasSize) or the configured maximum (LZ4.hs:189-198)
    const uint8_t* hsrc = static_cast<const uint8_t*>(src);
    int64_t slots_total = 0;
    for (int i = 0; i < n; i++) {
        int cap = max_block;
        if (header == 8) cap = src_len[i] >= 8 ? le32(hsrc + src_off[i] + 4) : 0;
        if (cap < 0) cap = 0;
        h_cap[i] = cap;
        h_slot_off[i] = slots_total;
        slots_total += (header == 8) ? (int64_t)cap : align16((int64_t)cap);
    }
    const bool contiguous = (header == 8);     // slots are already exactly the output layout
    if (contiguous && slots_total > dst_cap) return fail(B200LZ4_E_NOMEM, "dst_cap too small: need " + std::to_string(slots_total));
    if (stream_first) memcpy(hd + o_first, stream_first, sizeof(int32_t) * (ns + 1));
    if (streams) { void** hs = reinterpret_cast<void**>(hd + o_states); for (int s = 0; s < ns; s++) hs[s] = streams[s]->d_state; }

    if ((rc = c->d_src.ensure((size_t)src_bytes + 64))) return rc;
    if ((rc = c->d_slots.ensure((size_t)slots_total + 64))) return rc;
    if (!contiguous && (rc = c->d_out.ensure((size_t)slots_total + 64))) return rc;
    // descriptor rings for chunks the wide kernel may take (few streams): one slice of one CTA-arena per stream, per chunk
    int wide_first[kMaxChunks + 1] = {0};
    for (int k = 0; k < nchunks; k++) {
        const int nsk = chunks[k].s1 - chunks[k].s0;
        wide_first[k + 1] = wide_first[k] + (nsk < kWideMaxCtas ? nsk : kWideMaxCtas);
    }
    if ((rc = c->d_wide.ensure((size_t)wide_first[nchunks] * kWideArenaPerCta))) return rc;
    uint8_t* d_src = static_cast<uint8_t*>(c->d_src.p);
    uint8_t* d_slots = static_cast<uint8_t*>(c->d_slots.p);
    uint8_t* d_out = static_cast<uint8_t*>(c->d_out.p);
    int64_t* d_out_off = reinterpret_cast<int64_t*>(dd + o_out_off);
    int32_t* d_out_len = reinterpret_cast<int32_t*>(dd + o_out_len);

    cudaStream_t st = c->stream;
    CU(cudaEventRecord(c->ev[0], st));
    CU(cudaMemcpyAsync(dd, hd, o_upload_end, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(c->ev_d0, c->dstream));
    for (int k = 0; k < nchunks; k++) {
        const Chunk& ch = chunks[k];
        if (ch.hi > ch.lo) CU(cudaMemcpyAsync(d_src + ch.lo, hsrc + ch.lo, (size_t)(ch.hi - ch.lo), cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(c->ev_h2d[k], st));
        cudaStream_t ks = c->kstream[k % kKernelStreams];
        CU(cudaStreamWaitEvent(ks, c->ev_h2d[k], 0));
        if (k == 0) CU(cudaEventRecord(c->ev_k0, ks));
        DecompressArgs a{};
        a.src = d_src;
        a.src_off = reinterpret_cast<const int64_t*>(dd + o_src_off);
        a.src_len = reinterpret_cast<const int32_t*>(dd + o_src_len);
        a.n_blocks = n;
        a.first_block = ch.b0;
        a.stream_first = stream_first ? reinterpret_cast<const int32_t*>(dd + o_first) + ch.s0 : nullptr;
        a.n_streams = ch.s1 - ch.s0;
        a.states = streams ? reinterpret_cast<void* const*>(dd + o_states) + ch.s0 : nullptr;
        a.dst = d_slots;
        a.dst_off = reinterpret_cast<const int64_t*>(dd + o_slot_off);
        a.dst_cap = reinterpret_cast<const int32_t*>(dd + o_cap);
        a.out_len = d_out_len;
        a.header = header; a.max_block = max_block; a.scratch = c->scratch + k;
        a.wide_arena = reinterpret_cast<uint4*>(static_cast<uint8_t*>(c->d_wide.p) + (size_t)wide_first[k] * kWideArenaPerCta);
        a.wide_ctas = wide_first[k + 1] - wide_first[k];
        CU(launch_decompress(a, ks));
        c->launches += kernel_launches_per_decompress();
        const int nk = ch.b1 - ch.b0;
        if (!contiguous) {
            CompactArgs g{d_slots, a.dst_off + ch.b0, d_out_len + ch.b0, nk, 0, d_out + h_slot_off[ch.b0], d_out_off + ch.b0 + k, nullptr, nullptr};
            CU(launch_compact(g, ks));
            c->launches += kernel_launches_per_compact();
            CU(cudaMemcpyAsync(hd + o_out_off + sizeof(int64_t) * (ch.b0 + k), dd + o_out_off + sizeof(int64_t) * (ch.b0 + k),
                               sizeof(int64_t) * (nk + 1), cudaMemcpyDeviceToHost, ks));
        }
        CU(cudaMemcpyAsync(hd + o_out_len + sizeof(int32_t) * ch.b0, dd + o_out_len + sizeof(int32_t) * ch.b0,
                           sizeof(int32_t) * nk, cudaMemcpyDeviceToHost, ks));
        if (k == nchunks - 1) CU(cudaEventRecord(c->ev[2], ks));
        CU(cudaEventRecord(c->ev_k[k], ks));
        if (contiguous) {   // sizes are known from the headers: the D2H copy can be queued right away
            const int64_t lo = h_slot_off[ch.b0];
            const int64_t hi = (ch.b1 < n) ? h_slot_off[ch.b1] : slots_total;
            CU(cudaStreamWaitEvent(c->dstream, c->ev_k[k], 0));
            if (hi > lo) CU(cudaMemcpyAsync(static_cast<uint8_t*>(dst) + lo, d_slots + lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost, c->dstream));
        }
    }
    CU(cudaEventRecord(c->ev[1], st));
    int result = 0;
    if (contiguous) {
        memcpy(dst_off, h_slot_off, sizeof(int64_t) * n); dst_off[n] = slots_total;
    } else {
        const int64_t* h_out_off = reinterpret_cast<const int64_t*>(hd + o_out_off);
        int64_t base = 0;
        for (int k = 0; k < nchunks; k++) {
            const Chunk& ch = chunks[k];
            CU(cudaEventSynchronize(c->ev_k[k]));
            const int nk = ch.b1 - ch.b0;
            const int64_t* off_k = h_out_off + ch.b0 + k;
            const int64_t total_k = off_k[nk];
            for (int i = 0; i < nk; i++) dst_off[ch.b0 + i] = base + off_k[i];
            if (base + total_k > dst_cap) { result = fail(B200LZ4_E_NOMEM, "dst_cap too small"); break; }
            if (total_k) CU(cudaMemcpyAsync(static_cast<uint8_t*>(dst) + base, d_out + h_slot_off[ch.b0], (size_t)total_k,
                                            cudaMemcpyDeviceToHost, c->dstream));
            base += total_k;
        }
        dst_off[n] = base;
    }
    CU(cudaEventRecord(c->ev_d1, c->dstream));
    if ((rc = sync_all(c))) return rc;
    if (result) return result;
    cudaEventElapsedTime(&c->t_h2d, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&c->t_kernel, c->ev_k0, c->ev[2]);
    cudaEventElapsedTime(&c->t_d2h, c->ev_d0, c->ev_d1);
    memcpy(out_len, hd + o_out_len, sizeof(int32_t) * n);
    for (int i = 0; i < n; i++) if (out_len[i] < 0) return fail(B200LZ4_E_BLOCK, "block " + std::to_string(i) + " failed to decompress");
    return 0;
}

}  // namespace

// ============================================================== C ABI ====

extern "C" {

int b200lz4_compress_bound(int n) { return bound_of(n); }
int b200lz4_version(void) { return B200LZ4_VERSION_NUMBER; }
const char* b200lz4_last_error(void) { return g_err.c_str(); }

int b200lz4_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

int b200lz4_ctx_create(int device, b200lz4_ctx** out)
{
    if (!out) return fail(B200LZ4_E_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(B200LZ4_E_CUDA, "no CUDA device: this library has no CPU path"); }
    if (device < 0 || device >= n) return fail(B200LZ4_E_ARG, "device index out of range");
    int major = 0;
    CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return fail(B200LZ4_E_CUDA, "device is not compute capability 10.x (kernels are built for sm_100a only)");
    CU(cudaSetDevice(device));
    b200lz4_ctx* c = new (std::nothrow) b200lz4_ctx();
    if (!c) return fail(B200LZ4_E_NOMEM, "out of host memory");
    c->device = device;
    cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int i = 0; i < kKernelStreams && e2 == cudaSuccess; i++) e2 = cudaStreamCreateWithFlags(&c->kstream[i], cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaStreamCreateWithFlags(&c->dstream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&c->scratch, sizeof(Scratch) * kMaxChunks);
    if (e2 == cudaSuccess) e2 = cudaMemset(c->scratch, 0, sizeof(Scratch) * kMaxChunks);
    for (int i = 0; i < 4 && e2 == cudaSuccess; i++) e2 = cudaEventCreate(&c->ev[i]);
    for (int i = 0; i < kMaxChunks && e2 == cudaSuccess; i++) {
        e2 = cudaEventCreate(&c->ev_h2d[i]);
        if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->ev_k[i]);
    }
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->ev_k0);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->ev_d0);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->ev_d1);
    if (e2 != cudaSuccess) { b200lz4_ctx_destroy(c); return fail_cuda(e2, "ctx_create"); }
    *out = c;
    return 0;
}

void b200lz4_ctx_destroy(b200lz4_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->d_src.release(); c->d_slots.release(); c->d_out.release(); c->d_desc.release(); c->d_wide.release(); c->h_desc.release();
    if (c->scratch) cudaFree(c->scratch);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_h2d) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_k) if (e) cudaEventDestroy(e);
    if (c->ev_k0) cudaEventDestroy(c->ev_k0);
    if (c->ev_d0) cudaEventDestroy(c->ev_d0);
    if (c->ev_d1) cudaEventDestroy(c->ev_d1);
    for (auto& k : c->kstream) if (k) cudaStreamDestroy(k);
    if (c->dstream) cudaStreamDestroy(c->dstream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void* b200lz4_host_alloc(size_t bytes)
{
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail_cuda(e, "cudaMallocHost"); return nullptr; }
    return p;
}
void* b200lz4_host_alloc_wc(size_t bytes)
{
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocWriteCombined | cudaHostAllocPortable);
    if (e != cudaSuccess) { fail_cuda(e, "cudaHostAlloc(write-combined)"); return nullptr; }
    return p;
}
void* b200lz4_ctx_host_alloc(b200lz4_ctx* c, size_t bytes)
{
    if (!c) { fail(B200LZ4_E_ARG, "ctx is NULL"); return nullptr; }
    if (cudaSetDevice(c->device) != cudaSuccess) { fail(B200LZ4_E_CUDA, "cudaSetDevice"); return nullptr; }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { fail_cuda(e, "cudaHostAlloc"); return nullptr; }
    return p;
}
void b200lz4_host_free(void* p) { if (p) cudaFreeHost(p); }

int b200lz4_last_timing(b200lz4_ctx* c, float* h2d, float* kern, float* d2h)
{
    if (!c) return fail(B200LZ4_E_ARG, "ctx is NULL");
    if (h2d) *h2d = c->t_h2d;
    if (kern) *kern = c->t_kernel;
    if (d2h) *d2h = c->t_d2h;
    return 0;
}
int64_t b200lz4_launch_count(b200lz4_ctx* c) { return c ? c->launches : 0; }
const char* b200lz4_ctx_last_error(b200lz4_ctx* c) { return c ? c->err.c_str() : ""; }

// Parallel gather of separately allocated (pageable) arrays into one staging buffer: what the Haskell shim / api.py do
// before every batch call.  One contiguous range of arrays per thread, balanced by bytes.
int b200lz4_gather_host(void* dst, const void* const* src_ptrs, const int64_t* dst_off, const int32_t* len, int n, int threads)
{
    if (n < 0 || (n > 0 && (!dst || !src_ptrs || !dst_off || !len))) return fail(B200LZ4_E_ARG, "NULL argument");
    if (n == 0) return 0;
    int64_t total = 0;
    for (int i = 0; i < n; i++) { if (len[i] < 0 || (len[i] > 0 && !src_ptrs[i])) return fail(B200LZ4_E_ARG, "bad array"); total += len[i]; }
    unsigned hw = std::thread::hardware_concurrency();
    int t = threads > 0 ? threads : (int)(hw ? (hw > 16 ? 16 : hw) : 4);
    if (total < (int64_t(4) << 20)) t = 1;
    if (t > n) t = n;
    auto work = [&](int lo, int hi) {
        uint8_t* d = static_cast<uint8_t*>(dst);
        for (int i = lo; i < hi; i++) if (len[i]) memcpy(d + dst_off[i], src_ptrs[i], (size_t)len[i]);
    };
    if (t <= 1) { work(0, n); return 0; }
    std::vector<std::thread> pool;
    int lo = 0; int64_t done = 0;
    for (int k = 0; k < t; k++) {
        const int64_t target = total * (k + 1) / t;
        int hi = lo;
        while (hi < n && (done < target || k == t - 1)) done += len[hi++];
        if (hi > lo) pool.emplace_back(work, lo, hi);
        lo = hi;
    }
    for (auto& th : pool) th.join();
    return 0;
}

// Measurement aid (bench.py's copy ceiling): one plain H2D copy of h2d_bytes and one D2H copy of d2h_bytes on the ctx's
// two copy streams at once, no kernels; returns after both are complete.  Host buffers should be page-locked.
int b200lz4_copy_probe(b200lz4_ctx* c, const void* h_src, int64_t h2d_bytes, void* h_dst, int64_t d2h_bytes)
{
    if (!c || h2d_bytes < 0 || d2h_bytes < 0) return fail(B200LZ4_E_ARG, "bad argument");
    CU(cudaSetDevice(c->device));
    int rc;
    if (h2d_bytes && (rc = c->d_src.ensure((size_t)h2d_bytes + 64))) return rc;
    if (d2h_bytes && (rc = c->d_out.ensure((size_t)d2h_bytes + 64))) return rc;
    if (h2d_bytes) CU(cudaMemcpyAsync(c->d_src.p, h_src, (size_t)h2d_bytes, cudaMemcpyHostToDevice, c->stream));
    if (d2h_bytes) CU(cudaMemcpyAsync(h_dst, c->d_out.p, (size_t)d2h_bytes, cudaMemcpyDeviceToHost, c->dstream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->dstream));
    return 0;
}

int b200lz4_cstream_create(b200lz4_ctx* c, b200lz4_cstream** out)
{
    if (!c || !out) return fail(B200LZ4_E_ARG, "NULL argument");
    *out = nullptr;
    CU(cudaSetDevice(c->device));
    b200lz4_cstream* s = new (std::nothrow) b200lz4_cstream{c, nullptr, nullptr, 0, 0};
    if (!s) return fail(B200LZ4_E_NOMEM, "out of host memory");
    cudaError_t e = cudaMalloc(&s->d_state, sizeof(CState));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_state, 0, sizeof(CState), c->stream);     // LZ4_initStream, cbits/lz4.c:1443-1451
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { if (s->d_state) cudaFree(s->d_state); delete s; return fail_cuda(e, "cstream_create"); }
    *out = s;
    return 0;
}
void b200lz4_cstream_free(b200lz4_cstream* s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->d_state) cudaFree(s->d_state);
    if (s->d_dict) cudaFree(s->d_dict);
    delete s;
}
int b200lz4_cstream_peek(b200lz4_cstream* s, uint32_t* table, uint32_t* offset)
{
    if (!s) return fail(B200LZ4_E_ARG, "NULL stream");
    CU(cudaSetDevice(s->ctx->device));
    CU(cudaStreamSynchronize(s->ctx->stream));
    if (table) CU(cudaMemcpy(table, s->d_state->table, sizeof(uint32_t) * kHashEntries, cudaMemcpyDeviceToHost));
    if (offset) CU(cudaMemcpy(offset, &s->d_state->offset, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}
int b200lz4_dstream_create(b200lz4_ctx* c, b200lz4_dstream** out)
{
    if (!c || !out) return fail(B200LZ4_E_ARG, "NULL argument");
    *out = nullptr;
    CU(cudaSetDevice(c->device));
    b200lz4_dstream* s = new (std::nothrow) b200lz4_dstream{c, nullptr};
    if (!s) return fail(B200LZ4_E_NOMEM, "out of host memory");
    cudaError_t e = cudaMalloc(&s->d_state, sizeof(DState));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_state, 0, 16, c->stream);                 // calloc, cbits/lz4.c:2265-2270
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { if (s->d_state) cudaFree(s->d_state); delete s; return fail_cuda(e, "dstream_create"); }
    *out = s;
    return 0;
}
void b200lz4_dstream_free(b200lz4_dstream* s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->d_state) cudaFree(s->d_state);
    delete s;
}

int b200lz4_compress_batch(b200lz4_ctx* c, const void* src, int64_t src_bytes,
                           const int64_t* src_off, const int32_t* src_len, int n_blocks,
                           const int32_t* stream_first, int n_streams, b200lz4_cstream* const* streams,
                           int acceleration, int header_mode,
                           void* dst, int64_t dst_cap, int64_t* dst_off, int32_t* out_len)
{
    const int rc = compress_host(c, src, src_bytes, src_off, src_len, n_blocks, stream_first, n_streams, streams,
                                 acceleration, header_mode, nullptr, dst, dst_cap, dst_off, out_len);
    if (c && rc) c->err = g_err;
    return rc;
}

int b200lz4_decompress_batch(b200lz4_ctx* c, const void* src, int64_t src_bytes,
                             const int64_t* src_off, const int32_t* src_len, int n_blocks,
                             const int32_t* stream_first, int n_streams, b200lz4_dstream* const* streams,
                             int header_mode, int max_block,
                             void* dst, int64_t dst_cap, int64_t* dst_off, int32_t* out_len)
{
    const int rc = decompress_host(c, src, src_bytes, src_off, src_len, n_blocks, stream_first, n_streams, streams,
                                   header_mode, max_block, dst, dst_cap, dst_off, out_len);
    if (c && rc) c->err = g_err;
    return rc;
}

}  // extern "C"

// ------------------------------------------------------------ several devices
// One batch striped over the GPUs of one box: contiguous ranges of blocks (independent mode) or of
// whole streams (linked mode), balanced by source bytes, one host thread and one b200lz4_ctx per device, no exchange step.

struct b200lz4_mctx {
    std::vector<b200lz4_ctx*> ctx;
    std::string err;
};

namespace {

struct Stripe { int u0, u1, b0, b1; };

// contiguous unit ranges with about equal source bytes
std::vector<Stripe> plan_stripes(const int32_t* len, int n, const int32_t* first, int ns, int parts)
{
    const int units = first ? ns : n;
    int64_t total = 0;
    for (int i = 0; i < n; i++) total += len[i];
    std::vector<Stripe> out;
    int u = 0; int64_t done = 0;
    for (int k = 0; k < parts; k++) {
        Stripe s; s.u0 = u; s.b0 = first ? first[u] : u;
        const int64_t target = total * (k + 1) / parts;
        while (u < units && (done < target || k == parts - 1)) {
            const int a = first ? first[u] : u, b = first ? first[u + 1] : u + 1;
            for (int i = a; i < b; i++) done += len[i];
            u++;
        }
        s.u1 = u; s.b1 = first ? first[u] : u;
        out.push_back(s);
    }
    return out;
}

template <class Call>
int run_striped(b200lz4_mctx* m, const int64_t* src_off, const int32_t* src_len, int n,
                const int32_t* stream_first, int n_streams, const std::vector<int64_t>& region, /* n + 1 prefix of worst-case output */
                int64_t* dst_off, int32_t* out_len, Call call)
{
    const int parts = (int)m->ctx.size();
    const std::vector<Stripe> st = plan_stripes(src_len, n, stream_first, n_streams, parts);
    std::vector<int> rcs(parts, 0);
    std::vector<std::string> errs(parts);
    std::vector<std::thread> pool;
    for (int d = 0; d < parts; d++) {
        const Stripe s = st[d];
        if (s.b1 <= s.b0) continue;
        pool.emplace_back([&, d, s]() {
            const int nb = s.b1 - s.b0;
            int64_t lo = INT64_MAX, hi = 0;
            for (int i = s.b0; i < s.b1; i++) { lo = std::min(lo, src_off[i]); hi = std::max(hi, src_off[i] + (int64_t)src_len[i]); }
            if (lo > hi) lo = hi = 0;
            std::vector<int64_t> off(nb), doff(nb + 1);
            for (int i = 0; i < nb; i++) off[i] = src_off[s.b0 + i] - lo;
            std::vector<int32_t> first;
            if (stream_first) { first.resize(s.u1 - s.u0 + 1); for (int u = s.u0; u <= s.u1; u++) first[u - s.u0] = stream_first[u] - s.b0; }
            rcs[d] = call(m->ctx[d], lo, hi - lo, off.data(), src_len + s.b0, nb, stream_first ? first.data() : nullptr, s.u1 - s.u0,
                          region[s.b0], region[s.b1] - region[s.b0], doff.data(), out_len + s.b0);
            if (rcs[d]) errs[d] = g_err;
            if (rcs[d] == 0 || rcs[d] == B200LZ4_E_BLOCK) for (int i = 0; i < nb; i++) dst_off[s.b0 + i] = region[s.b0] + doff[i];
            if ((rcs[d] == 0 || rcs[d] == B200LZ4_E_BLOCK) && s.b1 == n) dst_off[n] = region[s.b0] + doff[nb];
        });
    }
    for (auto& t : pool) t.join();
    int rc = 0;
    for (int d = 0; d < parts; d++) {
        if (rcs[d] && (rc == 0 || rc == B200LZ4_E_BLOCK)) { rc = rcs[d]; m->err = "device " + std::to_string(m->ctx[d]->device) + ": " + errs[d]; }
    }
    if (rc) g_err = m->err;
    return rc;
}

}  // namespace

extern "C" {

int b200lz4_mctx_create(const int* devices, int n, b200lz4_mctx** out)
{
    if (!out) return fail(B200LZ4_E_ARG, "out is NULL");
    *out = nullptr;
    std::vector<int> devs;
    if (devices) { if (n <= 0) return fail(B200LZ4_E_ARG, "no devices"); devs.assign(devices, devices + n); }
    else {
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { cudaGetLastError(); return fail(B200LZ4_E_CUDA, "no CUDA device: this library has no CPU path"); }
        if (n > 0 && n < cnt) cnt = n;
        for (int d = 0; d < cnt; d++) devs.push_back(d);
    }
    b200lz4_mctx* m = new (std::nothrow) b200lz4_mctx();
    if (!m) return fail(B200LZ4_E_NOMEM, "out of host memory");
    for (int d : devs) {
        b200lz4_ctx* c = nullptr;
        const int rc = b200lz4_ctx_create(d, &c);
        if (rc) { for (auto* x : m->ctx) b200lz4_ctx_destroy(x); delete m; return rc; }
        m->ctx.push_back(c);
    }
    *out = m;
    return 0;
}
void b200lz4_mctx_destroy(b200lz4_mctx* m)
{
    if (!m) return;
    for (auto* c : m->ctx) b200lz4_ctx_destroy(c);
    delete m;
}
int b200lz4_mctx_size(b200lz4_mctx* m) { return m ? (int)m->ctx.size() : 0; }
b200lz4_ctx* b200lz4_mctx_ctx(b200lz4_mctx* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }
const char* b200lz4_mctx_last_error(b200lz4_mctx* m) { return m ? m->err.c_str() : ""; }

int b200lz4_compress_batch_multi(b200lz4_mctx* m, const void* src, int64_t src_bytes,
                                 const int64_t* src_off, const int32_t* src_len, int n_blocks,
                                 const int32_t* stream_first, int n_streams,
                                 int acceleration, int header_mode,
                                 void* dst, int64_t dst_cap, int64_t* dst_off, int32_t* out_len)
{
    if (!m || m->ctx.empty()) return fail(B200LZ4_E_ARG, "mctx is NULL");
    if (n_blocks < 0 || (n_blocks > 0 && (!src_off || !src_len || !dst_off || !out_len || !src || !dst))) return fail(B200LZ4_E_ARG, "NULL argument");
    if (header_mode != 0 && header_mode != 4 && header_mode != 8) return fail(B200LZ4_E_ARG, "header_mode must be 0, 4 or 8");
    if (n_blocks == 0) { if (dst_off) dst_off[0] = 0; return 0; }
    int rc;
    if ((rc = check_blocks(src_off, src_len, n_blocks, src_bytes))) return rc;
    if ((rc = check_streams(stream_first, n_streams, n_blocks))) return rc;
    std::vector<int64_t> region(n_blocks + 1, 0);
    for (int i = 0; i < n_blocks; i++) region[i + 1] = region[i] + header_mode + bound_of(src_len[i]);
    if (region[n_blocks] > dst_cap) return fail(B200LZ4_E_NOMEM, "dst_cap too small: need " + std::to_string(region[n_blocks]));
    const uint8_t* hs = static_cast<const uint8_t*>(src);
    uint8_t* hd = static_cast<uint8_t*>(dst);
    return run_striped(m, src_off, src_len, n_blocks, stream_first, n_streams, region, dst_off, out_len,
        [&](b200lz4_ctx* c, int64_t lo, int64_t bytes, const int64_t* off, const int32_t* len, int nb, const int32_t* first, int ns,
            int64_t r0, int64_t rcap, int64_t* doff, int32_t* olen) {
            return compress_host(c, hs + lo, bytes, off, len, nb, first, ns, nullptr, acceleration, header_mode, nullptr,
                                 hd + r0, rcap, doff, olen);
        });
}

int b200lz4_decompress_batch_multi(b200lz4_mctx* m, const void* src, int64_t src_bytes,
                                   const int64_t* src_off, const int32_t* src_len, int n_blocks,
                                   const int32_t* stream_first, int n_streams,
                                   int header_mode, int max_block,
                                   void* dst, int64_t dst_cap, int64_t* dst_off, int32_t* out_len)
{
    if (!m || m->ctx.empty()) return fail(B200LZ4_E_ARG, "mctx is NULL");
    if (n_blocks < 0 || (n_blocks > 0 && (!src_off || !src_len || !dst_off || !out_len || !src || !dst))) return fail(B200LZ4_E_ARG, "NULL argument");
    if (header_mode != 0 && header_mode != 4 && header_mode != 8) return fail(B200LZ4_E_ARG, "header_mode must be 0, 4 or 8");
    if (header_mode != 8 && max_block < 0) return fail(B200LZ4_E_ARG, "max_block");
    if (n_blocks == 0) { if (dst_off) dst_off[0] = 0; return 0; }
    int rc;
    if ((rc = check_blocks(src_off, src_len, n_blocks, src_bytes))) return rc;
    if ((rc = check_streams(stream_first, n_streams, n_blocks))) return rc;
    const uint8_t* hs = static_cast<const uint8_t*>(src);
    uint8_t* hd = static_cast<uint8_t*>(dst);
    std::vector<int64_t> region(n_blocks + 1, 0);
    for (int i = 0; i < n_blocks; i++) {
        int cap = max_block;
        if (header_mode == 8) cap = src_len[i] >= 8 ? le32(hs + src_off[i] + 4) : 0;
        if (cap < 0) cap = 0;
        region[i + 1] = region[i] + (header_mode == 8 ? (int64_t)cap : align16((int64_t)cap));
    }
    if (region[n_blocks] > dst_cap) return fail(B200LZ4_E_NOMEM, "dst_cap too small: need " + std::to_string(region[n_blocks]));
    return run_striped(m, src_off, src_len, n_blocks, stream_first, n_streams, region, dst_off, out_len,
        [&](b200lz4_ctx* c, int64_t lo, int64_t bytes, const int64_t* off, const int32_t* len, int nb, const int32_t* first, int ns,
            int64_t r0, int64_t rcap, int64_t* doff, int32_t* olen) {
            return decompress_host(c, hs + lo, bytes, off, len, nb, first, ns, nullptr, header_mode, max_block,
                                   hd + r0, rcap, doff, olen);
        });
}

}  // extern "C"

extern "C" {

// ---------------------------------------------------------------- device path

size_t b200lz4_cstate_bytes(void) { return sizeof(CState); }
size_t b200lz4_dstate_bytes(void) { return sizeof(DState); }
size_t b200lz4_scratch_bytes(void) { return kScratchBytes; }

int b200lz4_cstate_set_dict(void* d_state, void* d_dict_buf, uint32_t dict_cap, void* cuda_stream)
{
    if (!d_state) return fail(B200LZ4_E_ARG, "NULL state");
    return set_dict_fields(static_cast<CState*>(d_state), static_cast<uint8_t*>(d_dict_buf), dict_cap,
                           static_cast<cudaStream_t>(cuda_stream));
}

int b200lz4_compress_dev(const void* d_src, const int64_t* d_src_off, const int32_t* d_src_len, int n_blocks,
                         const int32_t* d_stream_first, int n_streams, void* const* d_states,
                         void* d_dst, const int64_t* d_dst_off, const int32_t* d_dst_cap, int32_t* d_out_len,
                         int acceleration, int header_mode, void* d_scratch, void* cuda_stream)
{
    if (n_blocks < 0 || !d_scratch) return fail(B200LZ4_E_ARG, "bad argument");
    if (header_mode != 0 && header_mode != 4 && header_mode != 8) return fail(B200LZ4_E_ARG, "header_mode must be 0, 4 or 8");
    if (n_blocks == 0) return 0;
    CompressArgs a{};
    a.src = static_cast<const uint8_t*>(d_src); a.src_off = d_src_off; a.src_len = d_src_len; a.n_blocks = n_blocks;
    a.stream_first = d_stream_first; a.n_streams = d_stream_first ? n_streams : n_blocks; a.states = d_states;
    a.dst = static_cast<uint8_t*>(d_dst); a.dst_off = d_dst_off; a.dst_cap = d_dst_cap; a.out_len = d_out_len;
    a.accel = acceleration; a.header = header_mode; a.scratch = static_cast<Scratch*>(d_scratch);
    CU(launch_compress(a, static_cast<cudaStream_t>(cuda_stream)));
    return 0;
}

int b200lz4_decompress_dev(const void* d_src, const int64_t* d_src_off, const int32_t* d_src_len, int n_blocks,
                           const int32_t* d_stream_first, int n_streams, void* const* d_states,
                           void* d_dst, const int64_t* d_dst_off, const int32_t* d_dst_cap, int32_t* d_out_len,
                           int header_mode, int max_block, void* d_scratch, void* cuda_stream)
{
    if (n_blocks < 0 || !d_scratch) return fail(B200LZ4_E_ARG, "bad argument");
    if (header_mode != 0 && header_mode != 4 && header_mode != 8) return fail(B200LZ4_E_ARG, "header_mode must be 0, 4 or 8");
    if (n_blocks == 0) return 0;
    DecompressArgs a{};
    a.src = static_cast<const uint8_t*>(d_src); a.src_off = d_src_off; a.src_len = d_src_len; a.n_blocks = n_blocks;
    a.stream_first = d_stream_first; a.n_streams = d_stream_first ? n_streams : n_blocks; a.states = d_states;
    a.dst = static_cast<uint8_t*>(d_dst); a.dst_off = d_dst_off; a.dst_cap = d_dst_cap; a.out_len = d_out_len;
    a.header = header_mode; a.max_block = max_block; a.scratch = static_cast<Scratch*>(d_scratch);
    a.wide_arena = reinterpret_cast<uint4*>(static_cast<uint8_t*>(d_scratch) + sizeof(Scratch)); a.wide_ctas = kWideMaxCtas;
    CU(launch_decompress(a, static_cast<cudaStream_t>(cuda_stream)));
    return 0;
}

int b200lz4_compact_dev(const void* d_slots, const int64_t* d_slot_off, const int32_t* d_len, int n_blocks,
                        int header_mode, void* d_out, int64_t* d_out_off, void* d_scratch, void* cuda_stream)
{
    (void)d_scratch;
    if (n_blocks < 0) return fail(B200LZ4_E_ARG, "bad argument");
    CompactArgs g{static_cast<const uint8_t*>(d_slots), d_slot_off, d_len, n_blocks, header_mode,
                  static_cast<uint8_t*>(d_out), d_out_off, nullptr, nullptr};
    CU(launch_compact(g, static_cast<cudaStream_t>(cuda_stream)));
    return 0;
}

// ------------------------------------------------------------------ checksums
int b200lz4_xxh32_dev(const void* d_buf, const int64_t* d_off, const int32_t* d_len, int n, uint32_t seed, uint32_t* d_out, void* cuda_stream)
{
    if (n < 0 || (n > 0 && (!d_buf || !d_off || !d_len || !d_out))) return fail(B200LZ4_E_ARG, "NULL argument");
    CU(launch_xxh32(static_cast<const uint8_t*>(d_buf), d_off, d_len, n, seed, d_out, static_cast<cudaStream_t>(cuda_stream)));
    return 0;
}

int b200lz4_xxh32_batch(b200lz4_ctx* c, const void* src, int64_t src_bytes, const int64_t* off, const int32_t* len, int n,
                        uint32_t seed, uint32_t* out)
{
    if (!c) return fail(B200LZ4_E_ARG, "ctx is NULL");
    if (n < 0 || (n > 0 && (!off || !len || !out))) return fail(B200LZ4_E_ARG, "NULL argument");
    if (n == 0) return 0;
    int rc;
    if ((rc = check_blocks(off, len, n, src_bytes))) return rc;
    CU(cudaSetDevice(c->device));
    Carver cv;
    const size_t o_off = cv.take(sizeof(int64_t) * n), o_len = cv.take(sizeof(int32_t) * n), o_up = cv.off, o_out = cv.take(sizeof(uint32_t) * n);
    if ((rc = c->h_desc.ensure(cv.off))) return rc;
    if ((rc = c->d_desc.ensure(cv.off))) return rc;
    if ((rc = c->d_src.ensure((size_t)src_bytes + 64))) return rc;
    uint8_t* hd = static_cast<uint8_t*>(c->h_desc.p); uint8_t* dd = static_cast<uint8_t*>(c->d_desc.p);
    memcpy(hd + o_off, off, sizeof(int64_t) * n); memcpy(hd + o_len, len, sizeof(int32_t) * n);
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(dd, hd, o_up, cudaMemcpyHostToDevice, st));
    if (src_bytes) CU(cudaMemcpyAsync(c->d_src.p, src, (size_t)src_bytes, cudaMemcpyHostToDevice, st));
    CU(launch_xxh32(static_cast<const uint8_t*>(c->d_src.p), reinterpret_cast<const int64_t*>(dd + o_off), reinterpret_cast<const int32_t*>(dd + o_len),
                    n, seed, reinterpret_cast<uint32_t*>(dd + o_out), st));
    c->launches += 1;
    CU(cudaMemcpyAsync(hd + o_out, dd + o_out, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(out, hd + o_out, sizeof(uint32_t) * n);
    return 0;
}

// ------------------------------------------------------------------ re-frame
// resizeChunksD's `process` (src/Streamly/Internal/LZ4.hs:459-484) applied to one contiguous range.
int b200lz4_reframe(const void* buf, int64_t len, int header_mode, int has_end_mark,
                    int64_t* block_off, int32_t* block_len, int64_t max_blocks,
                    int64_t* n_found, int64_t* consumed, int* ended)
{
    if (!n_found || !consumed || len < 0 || (len > 0 && !buf)) return fail(B200LZ4_E_ARG, "NULL argument");
    if (header_mode != 4 && header_mode != 8) return fail(B200LZ4_E_ARG, "header_mode must be 4 or 8");
    const uint8_t* p = static_cast<const uint8_t*>(buf);
    int64_t at = 0, k = 0;
    if (ended) *ended = 0;
    *n_found = 0; *consumed = 0;
    while (k < max_blocks) {
        int64_t rest = len - at;
        if (rest < 4) break;                                            // LZ4.hs:461-462
        int32_t comp = le32(p + at);
        if (has_end_mark && comp == 0) {                                // LZ4.hs:464-466 -> RFooter; footer is the 4 bytes themselves
            at += 4; if (ended) *ended = 1; break;
        }
        if (rest <= header_mode) break;                                 // LZ4.hs:468-469
        if (comp <= 0) { *n_found = k; *consumed = at; return fail(B200LZ4_E_FRAME, "block header with compLen <= 0"); }
        int64_t required = (int64_t)comp + header_mode;                 // LZ4.hs:474
        if (rest < required) break;                                     // LZ4.hs:477-478 (RAccumulate)
        if (block_off) block_off[k] = at;
        if (block_len) block_len[k] = (int32_t)required;
        k++; at += required;                                            // LZ4.hs:475-476, :479-484
    }
    *n_found = k; *consumed = at;
    return 0;
}

int b200lz4_reframe_dev(const void* d_buf, int64_t len, int header_mode, int has_end_mark,
                        int64_t* d_block_off, int32_t* d_block_len, int64_t max_blocks,
                        int64_t* d_result, void* cuda_stream)
{
    if (!d_result || len < 0 || (len > 0 && !d_buf) || max_blocks < 0 || (max_blocks > 0 && (!d_block_off || !d_block_len)))
        return fail(B200LZ4_E_ARG, "NULL argument");
    if (header_mode != 4 && header_mode != 8) return fail(B200LZ4_E_ARG, "header_mode must be 4 or 8");
    CU(launch_reframe(static_cast<const uint8_t*>(d_buf), len, header_mode, has_end_mark, d_block_off, d_block_len, max_blocks,
                      d_result, static_cast<cudaStream_t>(cuda_stream)));
    return 0;
}

// ------------------------------------------------------------ legacy aliases

struct LZ4_stream_u { b200lz4_cstream* s; };
struct LZ4_streamDecode_u { b200lz4_dstream* s; };

static std::mutex g_legacy_mu;
static b200lz4_ctx* g_legacy_ctx = nullptr;
static b200lz4_ctx* legacy_ctx()
{
    if (!g_legacy_ctx) {
        int dev = 0;
        if (const char* e = getenv("B200LZ4_DEVICE")) dev = atoi(e);
        if (b200lz4_ctx_create(dev, &g_legacy_ctx) != 0) {
            fprintf(stderr, "libb200lz4: %s\n", g_err.c_str());
            return nullptr;
        }
    }
    return g_legacy_ctx;
}

LZ4_stream_t* LZ4_createStream(void)
{
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    b200lz4_ctx* c = legacy_ctx();
    if (!c) return nullptr;
    LZ4_stream_u* h = new (std::nothrow) LZ4_stream_u{nullptr};
    if (!h) return nullptr;
    if (b200lz4_cstream_create(c, &h->s) != 0) { delete h; return nullptr; }
    return h;
}
int LZ4_freeStream(LZ4_stream_t* h)
{
    if (!h) return 0;
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    b200lz4_cstream_free(h->s); delete h;
    return 0;
}
LZ4_streamDecode_t* LZ4_createStreamDecode(void)
{
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    b200lz4_ctx* c = legacy_ctx();
    if (!c) return nullptr;
    LZ4_streamDecode_u* h = new (std::nothrow) LZ4_streamDecode_u{nullptr};
    if (!h) return nullptr;
    if (b200lz4_dstream_create(c, &h->s) != 0) { delete h; return nullptr; }
    return h;
}
int LZ4_freeStreamDecode(LZ4_streamDecode_t* h)
{
    if (!h) return 0;
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    b200lz4_dstream_free(h->s); delete h;
    return 0;
}
int LZ4_compressBound(int n) { return bound_of(n); }

int LZ4_compress_fast_continue(LZ4_stream_t* h, const char* src, char* dst, int srcSize, int dstCapacity, int acceleration)
{
    if (!h || !h->s || srcSize < 0 || !dst) return 0;
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    int64_t off = 0, dst_off[2] = {0, 0}; int32_t len = srcSize, out_len = 0, cap = dstCapacity, first[2] = {0, 1};
    static const char empty = 0;
    std::vector<char> tmp((size_t)bound_of(srcSize) + 16);
    b200lz4_cstream* ss = h->s;
    int rc = compress_host(ss->ctx, srcSize ? src : &empty, srcSize, &off, &len, 1, first, 1, &ss,
                           acceleration, 0, &cap, tmp.data(), (int64_t)tmp.size(), dst_off, &out_len);
    if (rc != 0 || out_len <= 0 || out_len > dstCapacity) return 0;
    memcpy(dst, tmp.data(), (size_t)out_len);
    return out_len;
}

int LZ4_decompress_safe_continue(LZ4_streamDecode_t* h, const char* src, char* dst, int srcSize, int dstCapacity)
{
    if (!h || !h->s || !src || srcSize < 0 || dstCapacity < 0) return -1;
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    int64_t off = 0, dst_off[2] = {0, 0}; int32_t len = srcSize, out_len = -1, first[2] = {0, 1};
    std::vector<char> tmp((size_t)dstCapacity + 16);
    b200lz4_dstream* ss = h->s;
    int rc = decompress_host(ss->ctx, src, srcSize, &off, &len, 1, first, 1, &ss, 0, dstCapacity,
                             tmp.data(), (int64_t)tmp.size(), dst_off, &out_len);
    if (rc != 0 && rc != B200LZ4_E_BLOCK) return -1;
    if (out_len > 0) memcpy(dst, tmp.data() + dst_off[0], (size_t)out_len);
    return out_len;
}

}  // extern "C"
