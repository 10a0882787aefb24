// Development probe: can a running kernel see a device-memory flag that a copy stream writes behind a bulk H2D copy?
// Matrix: how the flag is written (4-byte cudaMemcpyAsync from pinned / pageable memory) x how it is polled.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned ld_acq_sys(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_vol(const unsigned* p) { unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_rlx_gpu(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// mode 0: acquire.sys  1: volatile  2: relaxed.gpu ; out[0] = value seen, out[1] = kilo-cycles waited, out[2] = data word seen after the flag
__global__ void poll_kernel(const unsigned* flag, const unsigned* data, int mode, unsigned want, long long timeout_cycles, unsigned* out)
{
    const long long t0 = clock64();
    unsigned v = 0;
    for (;;) {
        v = mode == 0 ? ld_acq_sys(flag) : mode == 1 ? ld_vol(flag) : ld_rlx_gpu(flag);
        if (v >= want || clock64() - t0 > timeout_cycles) break;
        __nanosleep(1000);
    }
    if (threadIdx.x == 0) { out[0] = v; out[1] = (unsigned)((clock64() - t0) >> 10); out[2] = __ldg(data); }
}

int main()
{
    const size_t big = size_t(256) << 20;
    unsigned char *h_big, *d_big; unsigned *d_flag, *d_out, *h_const, h_out[3];
    CK(cudaHostAlloc((void**)&h_big, big, cudaHostAllocPortable));
    memset(h_big, 0x5A, big);
    CK(cudaMalloc((void**)&d_big, big));
    CK(cudaMalloc((void**)&d_flag, 256)); CK(cudaMalloc((void**)&d_out, 256));
    CK(cudaHostAlloc((void**)&h_const, 256, cudaHostAllocPortable));
    for (int i = 0; i < 64; i++) h_const[i] = i;
    unsigned pageable[64]; for (int i = 0; i < 64; i++) pageable[i] = i;
    cudaStream_t sc, sk; CK(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking));
    cudaEvent_t ev; CK(cudaEventCreate(&ev));
    const char* wnames[] = {"memcpy4 pinned", "memcpy4 pageable", "2D copy + memcpy4 pinned, kernel launched behind an event on the copy stream"};
    const char* pnames[] = {"ld.acquire.sys", "ld.volatile", "ld.relaxed.gpu"};
    for (int wm = 0; wm < 3; wm++) for (int pm = 0; pm < 3; pm++) {
        CK(cudaMemset(d_flag, 0, 256)); CK(cudaMemset(d_out, 0, 256)); CK(cudaMemset(d_big, 0, big));
        CK(cudaDeviceSynchronize());
        if (wm < 2) {
            poll_kernel<<<1, 32, 0, sk>>>(d_flag, (const unsigned*)(d_big + big - 4), pm, 1, 600000000LL, d_out);      // ~0.3 s
            CK(cudaMemcpyAsync(d_big, h_big, big, cudaMemcpyHostToDevice, sc));
            CK(cudaMemcpyAsync(d_flag, wm == 0 ? h_const + 1 : pageable + 1, 4, cudaMemcpyHostToDevice, sc));
        } else {
            // the library's pattern: segment 0 (2-D), flag = 1, event -> kernel; then segment 1 (2-D), flag = 2 while the kernel polls for 2
            const size_t pitch = 640000, rows = 200, width = 80000;
            CK(cudaMemcpy2DAsync(d_big, pitch, h_big, pitch, width, rows, cudaMemcpyHostToDevice, sc));
            CK(cudaMemcpyAsync(d_flag, h_const + 1, 4, cudaMemcpyHostToDevice, sc));
            CK(cudaEventRecord(ev, sc));
            CK(cudaStreamWaitEvent(sk, ev, 0));
            poll_kernel<<<1, 32, 0, sk>>>(d_flag, (const unsigned*)(d_big + pitch * (rows - 1) + width + width - 4), pm, 2, 600000000LL, d_out);
            for (int rep = 0; rep < 20; rep++) CK(cudaMemcpy2DAsync(d_big + width, pitch, h_big + width, pitch, width, rows, cudaMemcpyHostToDevice, sc));
            CK(cudaMemcpyAsync(d_flag, h_const + 2, 4, cudaMemcpyHostToDevice, sc));
        }
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h_out, d_out, 12, cudaMemcpyDeviceToHost));
        printf("%-36s | %-16s : flag seen %u after %u kcycles, data word %08x %s\n", wnames[wm], pnames[pm], h_out[0], h_out[1], h_out[2],
               h_out[0] >= (wm == 2 ? 2u : 1u) ? (h_out[2] == 0x5A5A5A5Au ? "OK" : "FLAG OK, DATA STALE") : "TIMED OUT");
        fflush(stdout);
    }
    return 0;
}
