// Ordered stream compaction of the survivors of a verdict array (north star: "compact the surviving segments and paths
// with warp-level scans"): out_idx[0 .. count) = ascending indices i with flags[i] == keep.
// Three small launches: per-CTA survivor counts (ballot + popc), an exclusive scan of the CTA counts by one CTA
// (warp shuffles), then each CTA re-derives its lanes' ranks with ballot scans and scatters the indices.
#include "common.cuh"

namespace ppnet {

constexpr int kCmpThreads = 256;
constexpr int kCmpPerThread = 4;
constexpr int kCmpTile = kCmpThreads * kCmpPerThread;      // 1024 flags per CTA

__global__ void __launch_bounds__(kCmpThreads)
compact_count_kernel(const uint8_t* __restrict__ flags, int64_t n, uint8_t keep, int64_t* __restrict__ block_cnt) {
    __shared__ int wsum[kCmpThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kCmpTile;
    int c = 0;
#pragma unroll
    for (int r = 0; r < kCmpPerThread; ++r) {
        const int64_t i = base + r * kCmpThreads + threadIdx.x;
        c += __popc(__ballot_sync(0xffffffffu, i < n && flags[i] == keep));   // same value in every lane of the warp
    }
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kCmpThreads / 32; ++w) t += wsum[w];
        block_cnt[blockIdx.x] = t;
    }
}

// exclusive scan in place; block_cnt[n_blocks] and *out_count receive the total
__global__ void __launch_bounds__(1024)
compact_scan_kernel(int64_t* __restrict__ block_cnt, int64_t n_blocks, int64_t* __restrict__ out_count) {
    __shared__ int64_t wtot[32];
    __shared__ int64_t carry_s, chunk_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < n_blocks; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const int64_t v = i < n_blocks ? block_cnt[i] : 0;
        int64_t inc = v;                                       // inclusive scan inside the warp
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int64_t o = __shfl_up_sync(0xffffffffu, inc, s);
            if (lane >= s) inc += o;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        if (warp == 0) {                                       // exclusive scan of the 32 warp totals
            const int64_t t = wtot[lane];
            int64_t ti = t;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int64_t o = __shfl_up_sync(0xffffffffu, ti, s);
                if (lane >= s) ti += o;
            }
            wtot[lane] = ti - t;
            if (lane == 31) chunk_s = ti;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        if (i < n_blocks) block_cnt[i] = carry + wtot[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + chunk_s;
        __syncthreads();
    }
    if (threadIdx.x == 0) { block_cnt[n_blocks] = carry_s; *out_count = carry_s; }
}

__global__ void __launch_bounds__(kCmpThreads)
compact_scatter_kernel(const uint8_t* __restrict__ flags, int64_t n, uint8_t keep, const int64_t* __restrict__ block_off,
                       int64_t* __restrict__ out_idx) {
    __shared__ int wcnt[kCmpPerThread][kCmpThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * kCmpTile;
    unsigned bal[kCmpPerThread];
    bool mine[kCmpPerThread];
#pragma unroll
    for (int r = 0; r < kCmpPerThread; ++r) {
        const int64_t i = base + r * kCmpThreads + threadIdx.x;
        mine[r] = i < n && flags[i] == keep;
        bal[r] = __ballot_sync(0xffffffffu, mine[r]);
        if (lane == 0) wcnt[r][warp] = __popc(bal[r]);
    }
    __syncthreads();
    const int64_t off = block_off[blockIdx.x];
#pragma unroll
    for (int r = 0; r < kCmpPerThread; ++r) {
        if (!mine[r]) continue;
        int before = 0;                                        // survivors of this CTA in earlier rows / earlier warps of this row
        for (int rr = 0; rr < r; ++rr)
            for (int w = 0; w < kCmpThreads / 32; ++w) before += wcnt[rr][w];
        for (int w = 0; w < warp; ++w) before += wcnt[r][w];
        out_idx[off + before + __popc(bal[r] & ((1u << lane) - 1u))] = base + r * kCmpThreads + threadIdx.x;
    }
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int64_t ppnet_compact_workspace_elems(int64_t n) { return (n + kCmpTile - 1) / kCmpTile + 1; }

extern "C" int ppnet_compact_u8(const uint8_t* flags, int64_t n, uint8_t keep, int64_t* out_idx, int64_t* out_count,
                                int64_t* workspace, void* stream) {
    PPNET_REQUIRE(n >= 0, "compact: negative n");
    PPNET_REQUIRE(out_count, "compact: out_count is null");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { PPNET_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st)); return PPNET_OK; }
    PPNET_REQUIRE(flags && out_idx && workspace, "compact: null pointer");
    const int64_t nb = (n + kCmpTile - 1) / kCmpTile;
    PPNET_REQUIRE(nb <= 2147483647LL, "compact: too many elements for one launch");
    compact_count_kernel<<<(unsigned)nb, kCmpThreads, 0, st>>>(flags, n, keep, workspace);
    PPNET_LAUNCH_CHECK("compact_count_kernel");
    compact_scan_kernel<<<1, 1024, 0, st>>>(workspace, nb, out_count);
    PPNET_LAUNCH_CHECK("compact_scan_kernel");
    compact_scatter_kernel<<<(unsigned)nb, kCmpThreads, 0, st>>>(flags, n, keep, workspace, out_idx);
    PPNET_LAUNCH_CHECK("compact_scatter_kernel");
    return PPNET_OK;
}
