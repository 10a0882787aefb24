// xxh32.cu -- XXH32 checksums of many byte ranges (LZ4 frame format: header checksum, block checksums, content checksum).
//
// The reference stops at a stub (src/Streamly/Internal/LZ4.hs:602 `assertHeaderChecksum = satisfy (const True)`, :631-640
// rejects every checksum flag); SURVEY.md section 8f rank 3 asks for the real thing.  xxHash is not part of the
// reference tree (lz4's frame library bundles it); this follows the published XXH32 definition:
//   >= 16 bytes: four accumulators v1..v4 = seed + P1 + P2, seed + P2, seed, seed - P1 consume 16-byte stripes
//                (v = rotl(v + word * P2, 13) * P1), h = rotl(v1,1) + rotl(v2,7) + rotl(v3,12) + rotl(v4,18); else h = seed + P5;
//   h += length; remaining words: h = rotl(h + w * P3, 17) * P4; remaining bytes: h = rotl(h + b * P5, 11) * P1;
//   avalanche: h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16.
//
// One warp per range: every iteration stages 512 bytes (32 lanes x 16 bytes, coalesced) in shared memory, then lanes 0..3
// each advance one accumulator over the 32 stripes -- the four multiply-rotate chains are the only parallelism a single
// range offers, so the throughput comes from many ranges (blocks) in flight.
#include "common.cuh"
#include "kernels.h"

namespace b200lz4 {

namespace {

constexpr uint32_t kP1 = 2654435761u, kP2 = 2246822519u, kP3 = 3266489917u, kP4 = 668265263u, kP5 = 374761393u;
__device__ __forceinline__ uint32_t rotl(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

constexpr int kXWarps = 4;

__global__ void __launch_bounds__(kXWarps * 32)
xxh32_kernel(const uint8_t* __restrict__ buf, const int64_t* __restrict__ off, const int32_t* __restrict__ len, int n,
             uint32_t seed, uint32_t* __restrict__ out)
{
    __shared__ uint32_t stage[kXWarps][128];            // 512 bytes per warp
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (int b = blockIdx.x * kXWarps + (int)warp; b < n; b += gridDim.x * kXWarps) {
        const uint8_t* p = buf + off[b];
        const uint32_t L = len[b] > 0 ? (uint32_t)len[b] : 0u;
        uint32_t v = 0;                                 // lanes 0..3: accumulator v1..v4
        if (lane == 0) v = seed + kP1 + kP2; else if (lane == 1) v = seed + kP2; else if (lane == 2) v = seed; else if (lane == 3) v = seed - kP1;
        const uint32_t stripes = L >> 4;
        for (uint32_t s0 = 0; s0 < stripes; s0 += 32) {
            const uint32_t s = s0 + lane;
            if (s < stripes) {
                const uint8_t* q = p + (size_t)s * 16;
                stage[warp][4 * lane + 0] = ldg_u32_unaligned(q); stage[warp][4 * lane + 1] = ldg_u32_unaligned(q + 4);
                stage[warp][4 * lane + 2] = ldg_u32_unaligned(q + 8); stage[warp][4 * lane + 3] = ldg_u32_unaligned(q + 12);
            }
            __syncwarp();
            if (lane < 4) {
                const uint32_t cnt = min(32u, stripes - s0);
                for (uint32_t k = 0; k < cnt; k++) v = rotl(v + stage[warp][4 * k + lane] * kP2, 13) * kP1;
            }
            __syncwarp();
        }
        // lane 0 finishes
        const uint32_t v1 = __shfl_sync(kFull, v, 0), v2 = __shfl_sync(kFull, v, 1), v3 = __shfl_sync(kFull, v, 2), v4 = __shfl_sync(kFull, v, 3);
        if (lane == 0) {
            uint32_t h = (L >= 16) ? rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18) : seed + kP5;
            h += L;
            uint32_t at = stripes << 4;
            for (; at + 4 <= L; at += 4) h = rotl(h + ldg_u32_unaligned(p + at) * kP3, 17) * kP4;
            for (; at < L; at++) h = rotl(h + (uint32_t)__ldg(p + at) * kP5, 11) * kP1;
            h ^= h >> 15; h *= kP2; h ^= h >> 13; h *= kP3; h ^= h >> 16;
            out[b] = h;
        }
        __syncwarp();
    }
}

}  // namespace

cudaError_t launch_xxh32(const uint8_t* buf, const int64_t* off, const int32_t* len, int n, uint32_t seed, uint32_t* out, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    int sm_count = 0;
    cudaError_t e = device_sm_count(&sm_count); if (e != cudaSuccess) return e;
    const int want = (n + kXWarps - 1) / kXWarps, max_ctas = sm_count * 8;
    xxh32_kernel<<<want < max_ctas ? want : max_ctas, kXWarps * 32, 0, stream>>>(buf, off, len, n, seed, out);
    return cudaGetLastError();
}

}  // namespace b200lz4
