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
constexpr int kMaxGroups = 16;          // streamed compress call: block groups (events, work counters, arrival flags)
constexpr int kMaxSegs = 16;            // ... and segments a block is cut into on its way to / from the device
constexpr size_t kFlagBytes = 1024 + kMaxGroups * kMaxSegs * sizeof(uint32_t);   // ctx->d_flags: see b200lz4_ctx

struct b200lz4_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;                  // H2D + bookkeeping
    cudaStream_t kstream[kKernelStreams] = {};      // chunk kernels (run concurrently)
    cudaStream_t dstream = nullptr;                 // D2H
    Scratch* scratch = nullptr;                     // kMaxChunks records: one work counter per in-flight kernel
    DevBuf d_src, d_slots, d_out, d_desc, d_wide;   // d_wide: descriptor rings of decompress_kernel_wide (one slice per chunk)
    PinBuf h_desc;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_h2d[kMaxGroups] = {}, ev_k[kMaxGroups] = {}, ev_k0 = nullptr, ev_d0 = nullptr, ev_d1 = nullptr;
    // streamed compress call (compress_host_streamed): per-group arrival flags + finder cycle counters in HBM, the
    // constants the copy stream writes into the flags, and what earlier calls measured (the copy schedule is tuned by it)
    uint32_t* d_flags = nullptr;            // [kMaxGroups] u32 flags, then at byte 64: u64 busy cycles, u64 bytes; at byte 1024:
                                            // [kMaxGroups][kMaxSegs] u32 segment counters of the streamed decompress call
    uint32_t* h_const = nullptr;            // pinned: h_const[i] == i
    double est_ns_per_byte = 0;             // finder time per input byte (0 = not measured yet)
    double est_h2d_gbs = 0;                 // H2D bandwidth of the last streamed call
    int clock_khz = 0;
    int streamed_calls = 0;
    cudaStream_t fstream = nullptr;         // arrival flags are raised from here (set_flag kernels behind per-piece events), so that
                                            // the copy stream's pieces run back to back
    std::vector<cudaEvent_t> ev_piece;
    int n_kgood = -1;                       // kernel streams verified not to share a hardware queue with the copy streams
                                            // or with each other (-1: not probed yet); they come first in kstream[]
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
    if (c->fstream) CU(cudaStreamSynchronize(c->fstream));
    return 0;
}

// ---- pipeline planning ------------------------------------------------------
// A host batch is cut into up to kMaxChunks chunks of whole streams (independent mode: whole
// blocks).  Chunk k's H2D copy, its kernels and its D2H copy run on three different CUDA
// streams, so that transfers in both directions overlap the kernels of neighbouring chunks;
// the chunk kernels themselves run concurrently on kKernelStreams streams because a codec
// kernel is bound by per-block latency, not by the number of blocks it is given.
struct Chunk { int b0, b1, s0, s1; int64_t lo, hi; };

int plan_chunks(const int64_t* off, const int32_t* len, int n, const int32_t* first, int ns, Chunk* out, bool early_start = false)
{
    int64_t total = 0;
    for (int i = 0; i < n; i++) total += len[i];
    // Chunk k becomes ready when its H2D copy lands and finishes one block latency later, so the end of the
    // call is "last byte arrives + block latency + D2H of the last chunk": keep the last chunk small.
    // early_start (mirrored decompress: the output leaves from inside the kernels, so the call ends one PCIe-bound stretch
    // after the FIRST chunk has landed): a small first chunk instead of a small last one
    static const double kShareTail[kMaxChunks] = {0.14, 0.20, 0.20, 0.20, 0.18, 0.08};
    static const double kShareHead[kMaxChunks] = {0.03, 0.09, 0.18, 0.24, 0.24, 0.22};
    const double* kShare = early_start ? kShareHead : kShareTail;
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

// ---- hardware-queue aliasing ---------------------------------------------------
// CUDA multiplexes streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware work queues (8 unless the variable is set before the
// context exists; the library's load-time constructor below asks for 32).  Entries of one queue are dispatched in order, so
// anything queued behind a kernel's same-stream successor waits for that kernel to END -- harmless for the plain pipeline,
// fatal for the streamed call, whose kernels wait for copies that are queued later on the copy stream.  Which streams share
// a queue is not documented, so it is MEASURED once per ctx: a kernel that spins on a flag (bounded by a timeout), a
// successor behind it, then a write of the flag through the other stream; if the spinning kernel times out, the two
// streams share a queue.  Kernel streams that alias the H2D stream, the D2H stream or an already accepted kernel stream are
// replaced by fresh ones (up to 40 tries).
__attribute__((constructor)) void ask_for_more_hardware_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
// (CUDA_MODULE_LOADING is deliberately left alone: probe_stream_aliasing loads the few kernels that matter by hand.)

int streams_alias(b200lz4_ctx* c, cudaStream_t spinner, cudaStream_t other, bool other_writes_by_kernel, bool* aliased)
{
    uint32_t* flag = c->d_flags + 40;       // bytes 160.. of the flag block: not used by the calls themselves
    uint32_t* result = c->d_flags + 44;
    CU(cudaMemsetAsync(flag, 0, 32, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(launch_alias_probe(flag, result, (long long)c->clock_khz * 15, spinner));     // 15 ms
    if (other_writes_by_kernel) CU(launch_set_flag(flag, 1u, other));
    else CU(cudaMemcpyAsync(flag, c->h_const + 1, sizeof(uint32_t), cudaMemcpyHostToDevice, other));
    CU(cudaStreamSynchronize(spinner));
    CU(cudaStreamSynchronize(other));
    uint32_t r[2] = {0, 0};
    CU(cudaMemcpy(r, result, sizeof r, cudaMemcpyDeviceToHost));
    *aliased = (r[0] != 1u);
    return 0;
}

int probe_stream_aliasing(b200lz4_ctx* c)
{
    if (c->n_kgood >= 0) return 0;
    // every kernel a streamed call launches is loaded NOW: a lazy first-time load in the middle of the call waits for
    // running kernels to end, and those wait for copies the blocked host thread has not queued yet
    CU(preload_compress_kernels());
    CU(preload_compact_kernels());
    {   // ... including whatever the runtime itself uses for memset and small copies: run each once
        CU(cudaMemsetAsync(c->d_flags, 0, kFlagBytes, c->stream));
        CU(cudaMemcpyAsync(c->d_flags + 48, c->h_const + 1, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        CU(launch_set_flag(c->d_flags + 48, 0u, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    std::vector<cudaStream_t> good, rejected;
    int next_existing = 0, tries = 0, rc;
    while ((int)good.size() < kKernelStreams && tries < 40) {
        cudaStream_t cand;
        if (next_existing < kKernelStreams) cand = c->kstream[next_existing++];
        else CU(cudaStreamCreateWithFlags(&cand, cudaStreamNonBlocking));
        tries++;
        bool bad = false;
        if ((rc = streams_alias(c, cand, c->stream, false, &bad))) return rc;
        if (!bad && (rc = streams_alias(c, cand, c->dstream, false, &bad))) return rc;
        for (size_t j = 0; j < good.size() && !bad; j++) if ((rc = streams_alias(c, good[j], cand, true, &bad))) return rc;
        (bad ? rejected : good).push_back(cand);
    }
    while (next_existing < kKernelStreams) rejected.push_back(c->kstream[next_existing++]);
    // the flag stream must not sit behind a waiting kernel either
    for (int t = 0; t < 16 && !c->fstream; t++) {
        cudaStream_t cand;
        CU(cudaStreamCreateWithFlags(&cand, cudaStreamNonBlocking));
        bool bad = false;
        for (size_t j = 0; j < good.size() && !bad; j++) if ((rc = streams_alias(c, good[j], cand, true, &bad))) return rc;
        if (bad) cudaStreamDestroy(cand); else c->fstream = cand;
    }
    c->n_kgood = (int)good.size();
    int k = 0;
    for (cudaStream_t x : good) c->kstream[k++] = x;
    for (cudaStream_t x : rejected) { if (k < kKernelStreams) c->kstream[k++] = x; else cudaStreamDestroy(x); }
    if (getenv("B200LZ4_DEBUG")) fprintf(stderr, "[b200lz4] hardware queues: %d of %d kernel streams independent of the copy streams after %d tries\n", c->n_kgood, kKernelStreams, tries);
    return 0;
}

// One segment of every block of a group: `rows` ranges of `width` bytes at constant pitches, as one pitched copy.  (The copy
// engine moves a pitched copy row by row: measured 9 % slower than one plain copy of the same bytes on an idle link and
// 15 % slower while the opposite direction is busy, which is why segments are not made smaller than they have to be.)
int copy_rows(uint8_t* dst, size_t dpitch, const uint8_t* src, size_t spitch, size_t width, size_t rows, cudaMemcpyKind kind, cudaStream_t st)
{
    if (rows == 0 || width == 0) return 0;
    if (rows == 1 || (width == dpitch && width == spitch)) CU(cudaMemcpyAsync(dst, src, width * rows, kind, st));
    else CU(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st));
    return 0;
}

// ---- streamed compress call ---------------------------------------------------
// A batch of equally long independent blocks at a constant pitch (what compressChunks over fixed-size reads produces)
// does not have to arrive block by block.  The end of the plain pipeline above is "last byte lands + one BLOCK latency"
// (a 640 000-byte block keeps one finder busy for ~7 ms, a third of the whole transfer).  Here blocks are cut into S
// segments and block groups into 2-D copies (one segment of every block of a group per copy), each followed by a 4-byte
// copy that bumps the group's arrival flag; the group's kernel is launched as soon as its FIRST segment has landed and
// its finders wait for later segments only if they get there before the data (compress.cu: kStreamed).  Pieces are sent
// in deadline order, key(g, s) = g + s * w: w = 0 is the plain block order, a large w sends segment 0 of every block
// first.  w is chosen so that a group's segments arrive at the pace its finders consume them, from what the previous
// streamed call on this ctx measured (finder cycles per byte, H2D rate); then the end of the call is "last byte lands +
// one SEGMENT latency + the last group's D2H".  The device layout is private to the call (128-byte aligned block starts,
// 128 bytes of gap), so no L1 sector is read before it is complete.
struct StreamShape { int L; int64_t P; bool ok; };
constexpr double kStreamSpread = 1.5;    // measured: a little later is better than a little early (the D2H copies of finished groups
                                        // share the link, and a starved finder only waits while an early piece delays everyone else's)
constexpr int kStreamedGaveUp = -1000;  // internal: compress_host retries through the plain pipeline

StreamShape streamed_shape(const int64_t* off, const int32_t* len, int n, const int32_t* stream_first, const int32_t* block_cap)
{
    StreamShape r{0, 0, false};
    const bool off_env = getenv("B200LZ4_NO_STREAMED") != nullptr;      // (read per call: bench.py times both pipelines in one process)
    if (off_env || stream_first || block_cap || n < 96) return r;
    const int L = len[0];
    if (L < 65536 || (int64_t)L * n < (int64_t(96) << 20)) return r;
    const int64_t P = off[1] - off[0];
    if (P < L) return r;
    for (int i = 0; i < n; i++) {
        if (off[i] != off[0] + P * i) return r;
        if (i < n - 1 ? len[i] != L : (len[i] <= 0 || len[i] > L)) return r;
    }
    r.L = L; r.P = P; r.ok = true;
    return r;
}

int compress_host_streamed(b200lz4_ctx* c, const void* src, const int64_t* src_off, const int32_t* src_len, int n,
                           const StreamShape& shp, int accel, int header,
                           void* dst, int64_t dst_cap, int64_t* dst_off, int32_t* out_len)
{
    int rc;
    const int L = shp.L;
    const int nks = c->n_kgood;             // kernel streams that cannot block the copy stream (>= 2, checked by the caller)
    const int64_t DP = ((int64_t)L + 127) / 128 * 128 + 128;       // device pitch
    int64_t total = 0;
    for (int i = 0; i < n; i++) total += src_len[i];

    // ---- schedule parameters
    const char* env_w = getenv("B200LZ4_STREAM_W");      // measurement overrides (read per call)
    const char* env_g = getenv("B200LZ4_STREAM_G");
    const char* env_s = getenv("B200LZ4_STREAM_S");
    const char* env_f = getenv("B200LZ4_STREAM_FLAG");
    const bool flag_by_kernel = env_f && env_f[0] == 'k';
    const bool flag_inline_copy = env_f && env_f[0] == 'c';
    const bool flag_separate = !flag_by_kernel && !flag_inline_copy && c->fstream != nullptr;
    const double ns_per_byte = c->est_ns_per_byte > 0 ? c->est_ns_per_byte : 12.0;       // first call: ~85 MB/s per finder
    const double h2d_gbs = c->est_h2d_gbs > 0 ? c->est_h2d_gbs : 50.0;
    const double t_block_ms = ns_per_byte * 1.1 * L * 1e-6, t_h2d_ms = (double)total / (h2d_gbs * 1e6);
    const double rho = t_block_ms / t_h2d_ms;
    int G = env_g ? atoi(env_g) : (rho <= 0.45 ? 2 * nks : nks);        // a group runs behind group g - nks on its stream
    if (G < 1) G = 1;
    if (G > kMaxGroups) G = kMaxGroups;
    // segments of ~160 000 bytes (rows of a pitched copy: smaller ones cost copy-engine efficiency while the D2H direction is
    // busy, larger ones lengthen the tail, which is one segment's worth of finder time): 4 for 640 000-byte blocks, 16 from 2.5 MB
    int S = env_s ? atoi(env_s) : (int)std::min<int64_t>(kMaxSegs, std::max<int64_t>(4, ((int64_t)L + 80000) / 160000));
    if (S < 1) S = 1;
    if (S > kMaxSegs) S = kMaxSegs;
    const int seg = (int)((((int64_t)L + S - 1) / S + 127) / 128 * 128);
    S = (L + seg - 1) / seg;
    const int per_group = (n + G - 1) / G;
    G = (n + per_group - 1) / per_group;
    // a group's S pieces are (S - 1) * w key units apart, the whole schedule spans (G - 1) + (S - 1) * w units in t_h2d:
    // the last piece of a group should land when its finders get there, (S - 1) / S of a block time after the first
    double w = (double)G;
    if (S - rho * (S - 1) > 0.01) w = kStreamSpread * rho * (G - 1) / (S - rho * (S - 1));
    if (env_w) w = atof(env_w);
    if (w > G) w = G;

    // ---- descriptor block (device offsets are the private layout)
    Carver cv;
    const size_t o_src_off = cv.take(sizeof(int64_t) * n);
    const size_t o_slot_off = cv.take(sizeof(int64_t) * n);
    const size_t o_src_len = cv.take(sizeof(int32_t) * n);
    const size_t o_upload_end = cv.off;
    const size_t o_out_len = cv.take(sizeof(int32_t) * n);
    const size_t o_out_off = cv.take(sizeof(int64_t) * (n + G));
    const size_t desc_bytes = cv.off;
    if ((rc = c->h_desc.ensure(desc_bytes))) return rc;
    if ((rc = c->d_desc.ensure(desc_bytes))) return rc;
    uint8_t* hd = static_cast<uint8_t*>(c->h_desc.p);
    uint8_t* dd = static_cast<uint8_t*>(c->d_desc.p);
    int64_t* h_src_off = reinterpret_cast<int64_t*>(hd + o_src_off);
    int64_t* h_slot_off = reinterpret_cast<int64_t*>(hd + o_slot_off);
    memcpy(hd + o_src_len, src_len, sizeof(int32_t) * n);
    int64_t slots_total = 0;
    for (int i = 0; i < n; i++) {
        h_src_off[i] = DP * i;
        h_slot_off[i] = slots_total;
        slots_total += align16((int64_t)bound_of(src_len[i]) + header + 16);
    }
    if ((rc = c->d_src.ensure((size_t)(DP * n) + 256))) return rc;
    if ((rc = c->d_slots.ensure((size_t)slots_total + 64))) return rc;
    if ((rc = c->d_out.ensure((size_t)slots_total + 64))) return rc;
    const uint8_t* hsrc = static_cast<const uint8_t*>(src);
    uint8_t* d_src = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(c->d_src.p) + 127) & ~uintptr_t(127));
    uint8_t* d_slots = static_cast<uint8_t*>(c->d_slots.p);
    uint8_t* d_out = static_cast<uint8_t*>(c->d_out.p);
    int64_t* d_out_off = reinterpret_cast<int64_t*>(dd + o_out_off);
    int32_t* d_out_len = reinterpret_cast<int32_t*>(dd + o_out_len);
    unsigned long long* d_busy = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(c->d_flags) + 64);

    // ---- pieces in deadline order
    struct Piece { int g, s; double key; };
    std::vector<Piece> pieces;
    pieces.reserve((size_t)G * S);
    for (int g = 0; g < G; g++) for (int s2 = 0; s2 < S; s2++) pieces.push_back({g, s2, g + s2 * w});
    std::stable_sort(pieces.begin(), pieces.end(), [](const Piece& x, const Piece& y) { return x.key < y.key; });

    cudaStream_t st = c->stream;
    CU(cudaEventRecord(c->ev[0], st));
    CU(cudaMemsetAsync(c->d_flags, 0, 256, st));
    CU(cudaMemcpyAsync(dd, hd, o_upload_end, cudaMemcpyHostToDevice, st));
    bool first_kernel = true;
    for (const Piece& pc : pieces) {
        const int b0 = pc.g * per_group, b1 = std::min(n, b0 + per_group);
        const int lo = pc.s * seg, width = std::min(seg, L - lo);
        int rows = b1 - b0;
        if (b1 == n && src_len[n - 1] < L) {        // the batch's last block may be shorter: its share of the segment goes separately
            rows--;
            const int wl = std::min(width, src_len[n - 1] - lo);
            if (wl > 0) CU(cudaMemcpyAsync(d_src + DP * (n - 1) + lo, hsrc + src_off[n - 1] + lo, (size_t)wl, cudaMemcpyHostToDevice, st));
        }
        if (rows > 0 && (rc = copy_rows(d_src + DP * b0 + lo, (size_t)DP, hsrc + src_off[b0] + lo, (size_t)shp.P, (size_t)width, (size_t)rows,
                                        cudaMemcpyHostToDevice, st))) return rc;
        if (flag_separate) {                // the flag is raised from another stream: the next piece starts right behind this one
            const size_t pi = (size_t)(&pc - pieces.data());
            while (c->ev_piece.size() <= pi) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->ev_piece.push_back(e); }
            CU(cudaEventRecord(c->ev_piece[pi], st));
            CU(cudaStreamWaitEvent(c->fstream, c->ev_piece[pi], 0));
            CU(launch_set_flag(c->d_flags + pc.g, (uint32_t)(pc.s + 1), c->fstream));
            c->launches++;
        }
        else if (flag_by_kernel) CU(launch_set_flag(c->d_flags + pc.g, (uint32_t)(pc.s + 1), st));
        else CU(cudaMemcpyAsync(c->d_flags + pc.g, c->h_const + (pc.s + 1), sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        if (pc.s != 0) continue;
        // first segment of the group is on its way: queue the group's kernels behind it
        CU(cudaEventRecord(c->ev_h2d[pc.g], st));
        cudaStream_t ks = c->kstream[pc.g % nks];
        CU(cudaStreamWaitEvent(ks, c->ev_h2d[pc.g], 0));
        if (first_kernel) { CU(cudaEventRecord(c->ev_k0, ks)); first_kernel = false; }
        CompressArgs a{};
        a.src = d_src;
        a.src_off = reinterpret_cast<const int64_t*>(dd + o_src_off);
        a.src_len = reinterpret_cast<const int32_t*>(dd + o_src_len);
        a.n_blocks = n;
        a.first_block = b0;
        a.n_streams = b1 - b0;
        a.dst = d_slots;
        a.dst_off = reinterpret_cast<const int64_t*>(dd + o_slot_off);
        a.out_len = d_out_len;
        a.accel = accel; a.header = header; a.scratch = c->scratch + pc.g;
        a.arrived = c->d_flags + pc.g; a.seg_bytes = seg; a.busy_cycles = d_busy;
        CU(launch_compress(a, ks));
        CompactArgs cg{d_slots, a.dst_off + b0, d_out_len + b0, b1 - b0, header, d_out + h_slot_off[b0], d_out_off + b0 + pc.g,
                       reinterpret_cast<int64_t*>(hd + o_out_off) + b0 + pc.g, reinterpret_cast<int32_t*>(hd + o_out_len) + b0};
        CU(launch_compact(cg, ks));
        c->launches += kernel_launches_per_compress() + kernel_launches_per_compact();
        CU(cudaEventRecord(c->ev_k[pc.g], ks));
    }
    CU(cudaEventRecord(c->ev[1], st));
    const bool dbg = getenv("B200LZ4_DEBUG") != nullptr;
    if (dbg) { fprintf(stderr, "[b200lz4] streamed: %zu pieces queued (G %d S %d seg %d w %.3f)\n", pieces.size(), G, S, seg, w); fflush(stderr); }

    // drain: as each group's sizes arrive, send its compacted bytes home
    const int32_t* h_out_len = reinterpret_cast<const int32_t*>(hd + o_out_len);
    const int64_t* h_out_off = reinterpret_cast<const int64_t*>(hd + o_out_off);
    int64_t base = 0;
    int result = 0;
    CU(cudaEventRecord(c->ev_d0, c->dstream));
    for (int g = 0; g < G; g++) {
        const int b0 = g * per_group, b1 = std::min(n, b0 + per_group);
        CU(cudaEventSynchronize(c->ev_k[g]));
        if (dbg) { fprintf(stderr, "[b200lz4] streamed: group %d done\n", g); fflush(stderr); }
        const int nk = b1 - b0;
        const int64_t* off_k = h_out_off + b0 + g;
        const int64_t total_k = off_k[nk];
        for (int i = 0; i < nk; i++) dst_off[b0 + i] = base + off_k[i];
        if (base + total_k > dst_cap) { result = fail(B200LZ4_E_NOMEM, "dst_cap too small"); break; }
        if (total_k) CU(cudaMemcpyAsync(static_cast<uint8_t*>(dst) + base, d_out + h_slot_off[b0], (size_t)total_k,
                                        cudaMemcpyDeviceToHost, c->dstream));
        base += total_k;
    }
    dst_off[n] = base;
    CU(cudaEventRecord(c->ev_d1, c->dstream));
    if ((rc = sync_all(c))) return rc;
    if (result) return result;
    memcpy(out_len, h_out_len, sizeof(int32_t) * n);
    cudaEventElapsedTime(&c->t_h2d, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&c->t_kernel, c->ev_k0, c->ev_k[G - 1]);
    cudaEventElapsedTime(&c->t_d2h, c->ev_d0, c->ev_d1);
    // what this call measured tunes the next one's schedule
    unsigned long long busy[3] = {0, 0, 0};
    CU(cudaMemcpy(busy, d_busy, sizeof busy, cudaMemcpyDeviceToHost));
    if (busy[2]) { fail(B200LZ4_E_CUDA, "streamed compress: " + std::to_string(busy[2]) + " blocks gave up waiting for their input segments"); return kStreamedGaveUp; }
    if (busy[1] && c->clock_khz > 0) {
        const double ns = (double)busy[0] / (double)busy[1] / ((double)c->clock_khz * 1e-6);
        c->est_ns_per_byte = c->est_ns_per_byte > 0 ? 0.5 * c->est_ns_per_byte + 0.5 * ns : ns;
    }
    if (c->t_h2d > 0) {
        const double gbs = (double)total / ((double)c->t_h2d * 1e6);
        c->est_h2d_gbs = c->est_h2d_gbs > 0 ? 0.5 * c->est_h2d_gbs + 0.5 * gbs : gbs;
    }
    c->streamed_calls++;
    if (getenv("B200LZ4_DEBUG")) {
        fprintf(stderr, "[b200lz4] streamed: G %d S %d seg %d w %.3f rho %.3f  est %.2f ns/B  h2d %.1f GB/s\n", G, S, seg, w, rho, c->est_ns_per_byte, c->est_h2d_gbs);
        float t = 0;
        for (int g = 0; g < G; g++) {
            float x = 0, y = 0;
            cudaEventElapsedTime(&x, c->ev[0], c->ev_h2d[g]); cudaEventElapsedTime(&y, c->ev[0], c->ev_k[g]);
            fprintf(stderr, "[b200lz4] group %d: first segment landed %.2f  kernels done %.2f\n", g, x, y);
        }
        cudaEventElapsedTime(&t, c->ev[0], c->ev[1]); fprintf(stderr, "[b200lz4] h2d done %.2f\n", t);
        cudaEventElapsedTime(&t, c->ev[0], c->ev_d1); fprintf(stderr, "[b200lz4] d2h done %.2f\n", t);
    }
    for (int i = 0; i < n; i++) if (out_len[i] <= 0) return fail(B200LZ4_E_BLOCK, "block " + std::to_string(i) + " failed to compress");
    return 0;
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
    if (!streams) {
        const StreamShape shp = streamed_shape(src_off, src_len, n, stream_first, block_cap);
        if (shp.ok) {
            if ((rc = probe_stream_aliasing(c))) return rc;
            if (c->n_kgood >= 2) {
                rc = compress_host_streamed(c, src, src_off, src_len, n, shp, accel, header, dst, dst_cap, dst_off, out_len);
                if (rc != kStreamedGaveUp) return rc;
                // finders gave up waiting for their input (should not happen on verified streams): never again on this
                // ctx, and this batch goes through the plain pipeline
                c->n_kgood = 0;
                if (getenv("B200LZ4_DEBUG")) fprintf(stderr, "[b200lz4] streamed call gave up (%s): plain pipeline from now on\n", g_err.c_str());
            }
        }
    }

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

int decompress_host_streamed(b200lz4_ctx* c, const void* src, const int64_t* src_off, const int32_t* src_len, int n,
                             int L, int cap_last, void* dst, int64_t dst_cap, int64_t* dst_off, int32_t* out_len);

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
    const uint8_t* hsrc = static_cast<const uint8_t*>(src);
    // How the output goes home (B200LZ4_DECODE_OUT): "copy" = D2H copies behind each chunk's kernel; "pieces" = segment by
    // segment while the blocks are being decoded (decompress_host_streamed; equally large independent blocks; the default
    // where it applies); "mirror" = independent blocks with their sizes in the headers, page-locked destination: the copier
    // warps store every output byte to the host buffer as well as to HBM (decompress.cu: kMirror) and no D2H copy follows.
    // Measured on config 2 (1.07 GB out, 0.77 GB in): copy 27.7 ms, mirror 27.0 ms (the posted writes of 1 678 CTAs reach
    // 40 GB/s and slow the H2D copy from 15 to 18.5 ms), pieces 26.4 ms -- so mirror stays opt-in.
    const bool no_streamed = getenv("B200LZ4_NO_STREAMED") != nullptr;
    const char* out_env = getenv("B200LZ4_DECODE_OUT");
    uint8_t* mirror_dst = nullptr;
    const char* dwide_env = getenv("B200LZ4_DWIDE");            // (tests force the wide kernel with it: that one has no mirror)
    if (out_env && out_env[0] == 'm' && header == 8 && !stream_first && !streams && !(dwide_env && dwide_env[0] == '1')) {
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, dst) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
            mirror_dst = static_cast<uint8_t*>(pa.devicePointer);
        else cudaGetLastError();
    }
    // equally large independent blocks with their size in the header: the output leaves segment by segment
    if (!mirror_dst && !no_streamed && !(out_env && out_env[0] == 'c') && header == 8 && !stream_first && !streams && n >= 96 && src_len[0] >= 8) {
        const int L = le32(hsrc + src_off[0] + 4);
        bool ok = L >= 65536 && (int64_t)L * n >= (int64_t(96) << 20);
        int cap_last = 0;
        for (int i = 0; i < n && ok; i++) {
            const int cap = src_len[i] >= 8 ? le32(hsrc + src_off[i] + 4) : -1;
            if (i < n - 1) ok = (cap == L); else { ok = (cap > 0 && cap <= L); cap_last = cap; }
        }
        if (ok) return decompress_host_streamed(c, src, src_off, src_len, n, L, cap_last, dst, dst_cap, dst_off, out_len);
    }
    Chunk chunks[kMaxChunks];
    const int nchunks = plan_chunks(src_off, src_len, n, stream_first, ns, chunks, mirror_dst != nullptr);

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
    if ((rc = c->d_slots.ensure((size_t)slots_total + 64 + 16))) return rc;
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
    if (mirror_dst) d_slots += reinterpret_cast<uintptr_t>(mirror_dst) & 15;      // same address modulo 16 on both sides: 128-bit stores line up
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
        a.host_dst = mirror_dst;
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
        if (contiguous && !mirror_dst) {   // sizes are known from the headers: the D2H copy can be queued right away
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

// ---- streamed decompress call --------------------------------------------------
// The mirror image of compress_host_streamed for the OUTPUT side: decompression of equally large independent blocks is
// bound by the D2H copy (1.40 x the bytes of the H2D copy on config 2), and in the plain pipeline the first byte cannot
// leave before a whole chunk has arrived AND one block latency has passed.  Here every group's kernel counts, per segment of
// the (equal) block capacity, the blocks whose flushed output has passed it (decompress.cu: kPublish); the last one raises
// a flag in page-locked host memory, the calling thread polls the flags and queues that segment of every block of the
// group as one 2-D D2H copy.  The copy engine then runs from "first group landed + one SEGMENT latency" to the end.
int decompress_host_streamed(b200lz4_ctx* c, const void* src, const int64_t* src_off, const int32_t* src_len, int n,
                             int L, int cap_last, void* dst, int64_t dst_cap, int64_t* dst_off, int32_t* out_len)
{
    int rc;
    const int header = 8;
    const int64_t out_total = (int64_t)L * (n - 1) + cap_last;
    if (out_total > dst_cap) return fail(B200LZ4_E_NOMEM, "dst_cap too small: need " + std::to_string(out_total));
    const char* env_g = getenv("B200LZ4_STREAM_G");
    const char* env_s = getenv("B200LZ4_STREAM_S");
    int G = env_g ? atoi(env_g) : 12;
    if (G < 1) G = 1;
    if (G > kMaxGroups) G = kMaxGroups;
    int S = env_s ? atoi(env_s) : 8;
    if (S < 1) S = 1;
    if (S > kMaxSegs) S = kMaxSegs;
    const int seg = (int)((((int64_t)L + S - 1) / S + 127) / 128 * 128);
    S = (L + seg - 1) / seg;
    const int per_group = (n + G - 1) / G;
    G = (n + per_group - 1) / per_group;
    if ((rc = probe_stream_aliasing(c))) return rc;     // nothing can deadlock here, but kernels queued behind another group's would start late
    const int nks = c->n_kgood >= 2 ? c->n_kgood : kKernelStreams;

    Carver cv;
    const size_t o_src_off = cv.take(sizeof(int64_t) * n);
    const size_t o_slot_off = cv.take(sizeof(int64_t) * n);
    const size_t o_src_len = cv.take(sizeof(int32_t) * n);
    const size_t o_cap = cv.take(sizeof(int32_t) * n);
    const size_t o_upload_end = cv.off;
    const size_t o_out_len = cv.take(sizeof(int32_t) * n);
    const size_t o_ready = cv.take(sizeof(uint32_t) * kMaxGroups * kMaxSegs);
    const size_t desc_bytes = cv.off;
    if ((rc = c->h_desc.ensure(desc_bytes))) return rc;
    if ((rc = c->d_desc.ensure(desc_bytes))) return rc;
    uint8_t* hd = static_cast<uint8_t*>(c->h_desc.p);
    uint8_t* dd = static_cast<uint8_t*>(c->d_desc.p);
    int64_t* h_slot_off = reinterpret_cast<int64_t*>(hd + o_slot_off);
    int32_t* h_cap = reinterpret_cast<int32_t*>(hd + o_cap);
    volatile uint32_t* ready = reinterpret_cast<volatile uint32_t*>(hd + o_ready);
    memcpy(hd + o_src_off, src_off, sizeof(int64_t) * n);
    memcpy(hd + o_src_len, src_len, sizeof(int32_t) * n);
    int64_t src_hi = 0;
    for (int i = 0; i < n; i++) {
        h_cap[i] = i < n - 1 ? L : cap_last;
        h_slot_off[i] = (int64_t)L * i;
        dst_off[i] = (int64_t)L * i;
        src_hi = std::max(src_hi, src_off[i] + (int64_t)src_len[i]);
    }
    dst_off[n] = out_total;
    for (int i = 0; i < kMaxGroups * kMaxSegs; i++) ready[i] = 0;
    if ((rc = c->d_src.ensure((size_t)src_hi + 64))) return rc;
    if ((rc = c->d_slots.ensure((size_t)out_total + 64))) return rc;
    const uint8_t* hsrc = static_cast<const uint8_t*>(src);
    uint8_t* d_src = static_cast<uint8_t*>(c->d_src.p);
    uint8_t* d_slots = static_cast<uint8_t*>(c->d_slots.p);
    uint32_t* d_counts = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(c->d_flags) + 1024);

    cudaStream_t st = c->stream;
    CU(cudaEventRecord(c->ev[0], st));
    CU(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * kMaxGroups * kMaxSegs, st));
    CU(cudaMemcpyAsync(dd, hd, o_upload_end, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(c->ev_d0, c->dstream));
    for (int g = 0; g < G; g++) {
        const int b0 = g * per_group, b1 = std::min(n, b0 + per_group);
        int64_t lo = INT64_MAX, hi = 0;
        for (int i = b0; i < b1; i++) { lo = std::min(lo, src_off[i]); hi = std::max(hi, src_off[i] + (int64_t)src_len[i]); }
        if (hi > lo) CU(cudaMemcpyAsync(d_src + lo, hsrc + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(c->ev_h2d[g], st));
        cudaStream_t ks = c->kstream[g % nks];
        CU(cudaStreamWaitEvent(ks, c->ev_h2d[g], 0));
        if (g == 0) CU(cudaEventRecord(c->ev_k0, ks));
        DecompressArgs a{};
        a.src = d_src;
        a.src_off = reinterpret_cast<const int64_t*>(dd + o_src_off);
        a.src_len = reinterpret_cast<const int32_t*>(dd + o_src_len);
        a.n_blocks = n;
        a.first_block = b0;
        a.n_streams = b1 - b0;
        a.dst = d_slots;
        a.dst_off = reinterpret_cast<const int64_t*>(dd + o_slot_off);
        a.dst_cap = reinterpret_cast<const int32_t*>(dd + o_cap);
        a.out_len = reinterpret_cast<int32_t*>(hd + o_out_len);     // page-locked, device-visible: no small D2H copies that would queue behind the pieces
        a.header = header; a.max_block = 0; a.scratch = c->scratch + g;
        a.seg_count = d_counts + g * kMaxSegs;
        a.host_ready = reinterpret_cast<uint32_t*>(hd + o_ready) + g * kMaxSegs;
        a.seg_bytes = seg; a.n_segs = S;
        CU(launch_decompress(a, ks));
        c->launches += kernel_launches_per_decompress();
        CU(cudaEventRecord(c->ev_k[g], ks));
    }
    CU(cudaEventRecord(c->ev[1], st));

    // drain: queue every (group, segment) piece as soon as its flag is up
    int next_s[kMaxGroups] = {0};
    int remaining = G * S;
    uint8_t* hdst = static_cast<uint8_t*>(dst);
    const bool dbg = getenv("B200LZ4_DEBUG") != nullptr;
    struct PieceLog { int g, s; double issued_ms; cudaEvent_t done; };
    std::vector<PieceLog> plog;
    const auto t_host0 = std::chrono::steady_clock::now();
    auto send = [&](int g, int sgm) -> int {
        const int b0 = g * per_group, b1 = std::min(n, b0 + per_group);
        const int lo = sgm * seg, width = std::min(seg, L - lo);
        int rows = b1 - b0;
        if (b1 == n && cap_last < L) {
            rows--;
            const int wl = std::min(width, cap_last - lo);
            if (wl > 0) CU(cudaMemcpyAsync(hdst + (int64_t)L * (n - 1) + lo, d_slots + (int64_t)L * (n - 1) + lo, (size_t)wl, cudaMemcpyDeviceToHost, c->dstream));
        }
        if (rows > 0) return copy_rows(hdst + (int64_t)L * b0 + lo, (size_t)L, d_slots + (int64_t)L * b0 + lo, (size_t)L, (size_t)width, (size_t)rows,
                                       cudaMemcpyDeviceToHost, c->dstream);
        return 0;
    };
    long spins = 0;
    while (remaining) {
        bool any = false;
        for (int g = 0; g < G; g++) {
            while (next_s[g] < S && ready[g * kMaxSegs + next_s[g]]) {
                if ((rc = send(g, next_s[g]))) return rc;
                if (dbg) {
                    PieceLog pl{g, next_s[g], std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count(), nullptr};
                    cudaEventCreate(&pl.done); cudaEventRecord(pl.done, c->dstream);
                    plog.push_back(pl);
                }
                next_s[g]++; remaining--; any = true;
            }
        }
        if (any) { spins = 0; continue; }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if (++spins % 4096 == 0) {          // a faulted or finished kernel must not leave this loop spinning
            bool all_done = true;
            for (int g = 0; g < G; g++) {
                const cudaError_t q = cudaEventQuery(c->ev_k[g]);
                if (q == cudaErrorNotReady) { all_done = false; break; }
                if (q != cudaSuccess) return fail_cuda(q, "streamed decompress");
            }
            if (all_done) {                 // every block has ended, so every flag is up: one last look
                bool missing = false;
                for (int g = 0; g < G; g++) for (int k = next_s[g]; k < S; k++) if (!ready[g * kMaxSegs + k]) missing = true;
                if (missing) return fail(B200LZ4_E_CUDA, "streamed decompress: kernels ended without raising every segment flag");
            }
        }
    }
    CU(cudaEventRecord(c->ev_d1, c->dstream));
    if ((rc = sync_all(c))) return rc;
    cudaEventElapsedTime(&c->t_h2d, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&c->t_kernel, c->ev_k0, c->ev_k[G - 1]);
    cudaEventElapsedTime(&c->t_d2h, c->ev_d0, c->ev_d1);
    memcpy(out_len, hd + o_out_len, sizeof(int32_t) * n);
    if (getenv("B200LZ4_DEBUG")) {
        fprintf(stderr, "[b200lz4] streamed decompress: G %d S %d seg %d\n", G, S, seg);
        float t = 0;
        for (int g = 0; g < G; g++) {
            float x = 0, y = 0;
            cudaEventElapsedTime(&x, c->ev[0], c->ev_h2d[g]); cudaEventElapsedTime(&y, c->ev[0], c->ev_k[g]);
            fprintf(stderr, "[b200lz4] group %d: input landed %.2f  kernel done %.2f\n", g, x, y);
        }
        cudaEventElapsedTime(&t, c->ev[0], c->ev_d1); fprintf(stderr, "[b200lz4] d2h done %.2f\n", t);
        for (auto& pl : plog) {     // (host time counts from the moment the drain loop started)
            float x = 0; cudaEventElapsedTime(&x, c->ev[0], pl.done);
            fprintf(stderr, "[b200lz4] piece g%d s%d: queued at host +%.2f ms, copied by %.2f\n", pl.g, pl.s, pl.issued_ms, x);
            cudaEventDestroy(pl.done);
        }
    }
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
    if (e2 == cudaSuccess) e2 = cudaMalloc(&c->scratch, sizeof(Scratch) * kMaxGroups);
    if (e2 == cudaSuccess) e2 = cudaMemset(c->scratch, 0, sizeof(Scratch) * kMaxGroups);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&c->d_flags, kFlagBytes);
    if (e2 == cudaSuccess) e2 = cudaMemset(c->d_flags, 0, kFlagBytes);
    if (e2 == cudaSuccess) e2 = cudaHostAlloc(reinterpret_cast<void**>(&c->h_const), 64 * sizeof(uint32_t), cudaHostAllocPortable);
    if (e2 == cudaSuccess) for (uint32_t i = 0; i < 64; i++) c->h_const[i] = i;
    if (e2 == cudaSuccess) e2 = cudaDeviceGetAttribute(&c->clock_khz, cudaDevAttrClockRate, device);
    for (int i = 0; i < 4 && e2 == cudaSuccess; i++) e2 = cudaEventCreate(&c->ev[i]);
    for (int i = 0; i < kMaxGroups && e2 == cudaSuccess; i++) {
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
    if (c->d_flags) cudaFree(c->d_flags);
    if (c->h_const) cudaFreeHost(c->h_const);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_h2d) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_k) if (e) cudaEventDestroy(e);
    if (c->ev_k0) cudaEventDestroy(c->ev_k0);
    if (c->ev_d0) cudaEventDestroy(c->ev_d0);
    if (c->ev_d1) cudaEventDestroy(c->ev_d1);
    for (auto& k : c->kstream) if (k) cudaStreamDestroy(k);
    if (c->fstream) cudaStreamDestroy(c->fstream);
    for (auto& e : c->ev_piece) if (e) cudaEventDestroy(e);
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
