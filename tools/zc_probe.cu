// Development probe: H2D by an SM gather kernel reading page-locked host memory, against the copy engine (plain and pitched),
// with and without a D2H copy running in the opposite direction.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int U>
__global__ void gather_rows(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t dpitch16, size_t spitch16, size_t width16, size_t rows)
{
    // rows x width16 vectors; consecutive threads take consecutive vectors of a row; U loads in flight per thread
    const size_t total = rows * width16;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
        uint4 v[U];
        #pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t i = i0 + u * stride;
            if (i < total) { const size_t r = i / width16, c = i - r * width16; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + r * spitch16 + c)); }
        }
        #pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t i = i0 + u * stride;
            if (i < total) { const size_t r = i / width16, c = i - r * width16; dst[r * dpitch16 + c] = v[u]; }
        }
    }
}

int main()
{
    const size_t total = size_t(1) << 30, back = size_t(768) << 20;
    unsigned char *h_src, *h_dst, *d_a, *d_b;
    CK(cudaHostAlloc((void**)&h_src, total, cudaHostAllocPortable));
    CK(cudaHostAlloc((void**)&h_dst, back, cudaHostAllocPortable));
    memset(h_src, 0x3C, total); memset(h_dst, 0, back);
    CK(cudaMalloc((void**)&d_a, total + 4096)); CK(cudaMalloc((void**)&d_b, back));
    cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t pitch = 640000, width = 80000, rows = 1677, segs = 8;
    for (int duplex = 0; duplex < 2; duplex++) {
        for (int mode = 0; mode < 8; mode++) {
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaDeviceSynchronize());
                if (duplex) for (int k = 0; k < 6; k++) CK(cudaMemcpyAsync(h_dst + k * (back / 6), d_b + k * (back / 6), back / 6, cudaMemcpyDeviceToHost, s2));
                CK(cudaEventRecord(e0, s1));
                if (mode == 0) CK(cudaMemcpyAsync(d_a, h_src, rows * pitch, cudaMemcpyHostToDevice, s1));
                else if (mode == 1) for (size_t sg = 0; sg < segs; sg++) CK(cudaMemcpy2DAsync(d_a + sg * width, pitch, h_src + sg * width, pitch, width, rows, cudaMemcpyHostToDevice, s1));
                else if (mode == 2) for (size_t sg = 0; sg < 4; sg++) CK(cudaMemcpy2DAsync(d_a + sg * 2 * width, pitch, h_src + sg * 2 * width, pitch, 2 * width, rows, cudaMemcpyHostToDevice, s1));
                else {
                    const int ctas = mode == 3 ? 8 : mode == 4 ? 16 : mode == 5 ? 32 : mode == 6 ? 64 : 148;
                    for (size_t sg = 0; sg < segs; sg++)
                        gather_rows<8><<<ctas, 256, 0, s1>>>((uint4*)(d_a + sg * width), (const uint4*)(h_src + sg * width), pitch / 16, pitch / 16, width / 16, rows);
                }
                CK(cudaEventRecord(e1, s1));
                CK(cudaDeviceSynchronize());
                float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            const char* names[] = {"copy engine, one plain copy", "copy engine, 8 pitched copies (80000-byte rows)", "copy engine, 4 pitched copies (160000-byte rows)",
                                   "gather kernel, 8 CTAs", "gather kernel, 16 CTAs", "gather kernel, 32 CTAs", "gather kernel, 64 CTAs", "gather kernel, 148 CTAs"};
            printf("%-7s %-52s %7.2f ms  %6.1f GB/s\n", duplex ? "duplex" : "alone", names[mode], best, rows * pitch / best / 1e6);
            fflush(stdout);
        }
    }
    // the other direction: D2H by SM stores into page-locked host memory (same kernel, operands swapped), against the copy engine
    for (int duplex = 0; duplex < 2; duplex++) {
        for (int mode = 0; mode < 6; mode++) {
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaDeviceSynchronize());
                if (duplex) for (int k = 0; k < 6; k++) CK(cudaMemcpyAsync(d_a + k * (back / 6), h_src + k * (back / 6), back / 6, cudaMemcpyHostToDevice, s2));
                CK(cudaEventRecord(e0, s1));
                const size_t rows2 = back / pitch;
                if (mode == 0) CK(cudaMemcpyAsync(h_dst, d_b, rows2 * pitch, cudaMemcpyDeviceToHost, s1));
                else if (mode == 1) for (size_t sg = 0; sg < segs; sg++) CK(cudaMemcpy2DAsync(h_dst + sg * width, pitch, d_b + sg * width, pitch, width, rows2, cudaMemcpyDeviceToHost, s1));
                else {
                    const int ctas = mode == 2 ? 8 : mode == 3 ? 16 : mode == 4 ? 32 : 64;
                    for (size_t sg = 0; sg < segs; sg++)
                        gather_rows<8><<<ctas, 256, 0, s1>>>((uint4*)(h_dst + sg * width), (const uint4*)(d_b + sg * width), pitch / 16, pitch / 16, width / 16, rows2);
                }
                CK(cudaEventRecord(e1, s1));
                CK(cudaDeviceSynchronize());
                float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            const char* names[] = {"D2H copy engine, one plain copy", "D2H copy engine, 8 pitched copies", "D2H store kernel, 8 CTAs", "D2H store kernel, 16 CTAs",
                                   "D2H store kernel, 32 CTAs", "D2H store kernel, 64 CTAs"};
            printf("%-7s %-52s %7.2f ms  %6.1f GB/s\n", duplex ? "duplex" : "alone", names[mode], best, (back / pitch) * pitch / best / 1e6);
            fflush(stdout);
        }
    }
    unsigned v = 0; CK(cudaMemcpy(&v, d_a + 12345 * 16, 4, cudaMemcpyDeviceToHost));
    printf("check word %08x\n", v);
    return 0;
}
