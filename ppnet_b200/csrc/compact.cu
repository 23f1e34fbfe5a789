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

// ---- the same compaction for bit-packed verdicts: survivor i <=> bit i is clear in every given array ---------------
// (the free segments under A11 AND A12 AND the DDA).  One word (32 segments) per thread: popc + shuffle scan for the
// counts; the scatter walks a warp's 32 words one at a time, each lane writing at most one index (coalesced).
constexpr int kCbThreads = 256;

__device__ __forceinline__ uint32_t free_word(const uint32_t* a, const uint32_t* b, const uint32_t* c, int64_t w, int64_t n_words,
                                              int64_t n) {
    if (w >= n_words) return 0u;
    uint32_t v = a[w];
    if (b) v |= b[w];
    if (c) v |= c[w];
    v = ~v;
    const int64_t left = n - (w << 5);                         // bits beyond n are not segments
    if (left < 32) v &= (1u << left) - 1u;
    return v;
}

__global__ void __launch_bounds__(kCbThreads)
compact_bits_count_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, const uint32_t* __restrict__ c,
                          int64_t n_words, int64_t n, int64_t* __restrict__ block_cnt) {
    __shared__ int wsum[kCbThreads / 32];
    const int64_t w = (int64_t)blockIdx.x * kCbThreads + threadIdx.x;
    int v = __popc(free_word(a, b, c, w, n_words, n));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < kCbThreads / 32; ++k) t += wsum[k];
        block_cnt[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kCbThreads)
compact_bits_scatter_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, const uint32_t* __restrict__ c,
                            int64_t n_words, int64_t n, const int64_t* __restrict__ block_off, int32_t idx_base,
                            int32_t* __restrict__ out_idx) {
    __shared__ int wsum[kCbThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * kCbThreads + threadIdx.x;
    const uint32_t v = free_word(a, b, c, w, n_words, n);
    const int mine = __popc(v);
    int inc = mine;                                            // inclusive scan of the warp's 32 word counts
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= s) inc += o;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int64_t off = block_off[blockIdx.x];
    for (int k = 0; k < warp; ++k) off += wsum[k];
    const int excl = inc - mine;
    for (int k = 0; k < 32; ++k) {                             // word k of this warp: lane l writes index 32 w_k + l if free
        const uint32_t vk = __shfl_sync(0xffffffffu, v, k);
        if (vk == 0u) continue;
        const int ek = __shfl_sync(0xffffffffu, excl, k);
        PPNET_ASSERT(!((vk >> lane) & 1u) || off + ek + __popc(vk & ((1u << lane) - 1u)) < n);
        if ((vk >> lane) & 1u)
            out_idx[off + ek + __popc(vk & ((1u << lane) - 1u))] = idx_base + (int32_t)(((w - lane + k) << 5) + lane);
    }
}

// short flag arrays (per-map flags of one slice: the valid paths): ONE CTA, ballot scans with a running carry
__global__ void __launch_bounds__(1024)
compact_u8_onecta_kernel(const uint8_t* __restrict__ flags, int64_t n, uint8_t keep, int32_t idx_base,
                         int32_t* __restrict__ out_idx, int64_t* __restrict__ out_count) {
    __shared__ int wcnt[32];
    __shared__ int64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < n; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const bool mine = i < n && flags[i] == keep;
        const unsigned bal = __ballot_sync(0xffffffffu, mine);
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, tile = 0;
        for (int w = 0; w < 32; ++w) { const int c = wcnt[w]; before += w < warp ? c : 0; tile += c; }
        const int64_t carry = carry_s;
        PPNET_ASSERT(!mine || carry + before + __popc(bal & ((1u << lane) - 1u)) < n);
        if (mine) out_idx[carry + before + __popc(bal & ((1u << lane) - 1u))] = idx_base + (int32_t)i;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tile;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_count = carry_s;
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_compact_u8_i32(const uint8_t* flags, int64_t n, uint8_t keep, int32_t idx_base, int32_t* out_idx,
                                    int64_t* out_count, void* stream) {
    PPNET_REQUIRE(n >= 0 && n + (int64_t)idx_base <= 2147483647LL, "compact_u8_i32: indices must fit an int32");
    PPNET_REQUIRE(out_count && (n == 0 || (flags && out_idx)), "compact_u8_i32: null pointer");
    compact_u8_onecta_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(flags, n, keep, idx_base, out_idx, out_count);
    PPNET_LAUNCH_CHECK("compact_u8_onecta_kernel");
    return PPNET_OK;
}

extern "C" int64_t ppnet_compact_bits_workspace_elems(int64_t n) { return ((n + 31) / 32 + kCbThreads - 1) / kCbThreads + 1; }

extern "C" int ppnet_compact_bits(const uint32_t* a, const uint32_t* b, const uint32_t* c, int64_t n, int32_t idx_base,
                                  int32_t* out_idx, int64_t* out_count, int64_t* workspace, void* stream) {
    PPNET_REQUIRE(n >= 0 && n + (int64_t)idx_base <= 2147483647LL, "compact_bits: indices must fit an int32");
    PPNET_REQUIRE(out_count, "compact_bits: out_count is null");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { PPNET_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st)); return PPNET_OK; }
    PPNET_REQUIRE(a && out_idx && workspace, "compact_bits: null pointer");
    const int64_t n_words = (n + 31) / 32, nb = (n_words + kCbThreads - 1) / kCbThreads;
    compact_bits_count_kernel<<<(unsigned)nb, kCbThreads, 0, st>>>(a, b, c, n_words, n, workspace);
    PPNET_LAUNCH_CHECK("compact_bits_count_kernel");
    compact_scan_kernel<<<1, 1024, 0, st>>>(workspace, nb, out_count);
    PPNET_LAUNCH_CHECK("compact_scan_kernel");
    compact_bits_scatter_kernel<<<(unsigned)nb, kCbThreads, 0, st>>>(a, b, c, n_words, n, workspace, idx_base, out_idx);
    PPNET_LAUNCH_CHECK("compact_bits_scatter_kernel");
    return PPNET_OK;
}

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
