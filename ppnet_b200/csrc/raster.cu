// Bit-packed occupancy maps and the integer DDA grid check.
//
//   A15  plot_obstacles  EDaGe-PP/Path.py:36-49 -- restated GEOMETRICALLY (matplotlib/JPEG/PIL dither are
//        unpinned third-party code): pixel (row i, col j) is an obstacle iff its centre
//        (j + 0.5, i + 0.5) lies inside a disk (x, y, r + inflate), evaluated in float64 as
//        rn(rn(dx^2) + rn(dy^2)) <= rn((r + inflate)^2).  1 bit per pixel: bit (j & 31) of word j >> 5.
//   DDA  new functionality (the reference has no occupancy-grid lookup): endpoints snapped with
//        the A4 rule (rint = half-to-even), then an all-integer walk of the major axis.
//
// raster: one CTA per map, bitmap built in shared memory with atomicOr on word masks (one (circle, row)
// span per lane), then streamed out with 16-byte stores.
// DDA:    one CTA per (map, chunk); the map's bitmap is pulled into shared memory with ONE bulk async
// copy (cp.async.bulk -> UBLKCP, completion on an mbarrier: 6 272 B at R = 224, 131 072 B at R = 1024);
// one warp per segment, 32 cells per round, __ballot_sync picks the first blocked cell and ends the walk.
#include "common.cuh"
#include "raster.cuh"

namespace ppnet {

// ------------------------------------------------------------------------------------------ raster
__global__ void __launch_bounds__(256)
raster_kernel(const double* __restrict__ obs, const int32_t* __restrict__ obs_cnt, int omax, int R, int W,
              double inflate, uint32_t* __restrict__ bits) {
    extern __shared__ __align__(16) uint32_t bm[];            // [R][W]
    const int64_t m = blockIdx.x;
    const int words = R * W;
    for (int i = threadIdx.x; i < words; i += blockDim.x) bm[i] = 0u;
    __syncthreads();
    const int cnt = min(obs_cnt[m], omax);
    const double* mo = obs + (size_t)m * omax * 3;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int k = warp; k < cnt; k += nwarps)
        raster_disk_warp(bm, R, W, mo[3 * k], mo[3 * k + 1], __dadd_rn(mo[3 * k + 2], inflate));
    __syncthreads();
    store_bitmap(bm, bits + (size_t)m * words, words);
}

// ------------------------------------------------------------------------------------------ DDA
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int floordiv32(int num, int den) {              // den > 0
    return num >= 0 ? num / den : -((den - 1 - num) / den);
}
__device__ __forceinline__ long long floordiv64(long long num, long long den) {
    return num >= 0 ? num / den : -((den - 1 - num) / den);
}

constexpr int kDdaThreads = 256;
constexpr int kCoordClamp = 1 << 29;

__device__ __forceinline__ int snap(float v, bool& bad) {
    if (!(v == v)) { bad = true; return 0; }
    const double r = rint((double)v);                                      // A4 rule, step 1, offset 0
    return (int)fmin(fmax(r, -(double)kCoordClamp), (double)kCoordClamp);
}

__global__ void __launch_bounds__(kDdaThreads)
dda_kernel(const uint32_t* __restrict__ bits, int R, int W, const float* __restrict__ segs,
           const int64_t* __restrict__ seg_off, int64_t segs_per_map, int chunk,
           uint8_t* __restrict__ verdict, int32_t* __restrict__ first_hit) {
    extern __shared__ __align__(128) unsigned char dsm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm);                      // 16 B header
    uint32_t* bm = reinterpret_cast<uint32_t*>(dsm + 16);
    const int m = blockIdx.x;
    const int64_t lo = seg_off ? seg_off[m] : (int64_t)m * segs_per_map;
    const int64_t hi = seg_off ? seg_off[m + 1] : lo + segs_per_map;
    const int64_t base = lo + (int64_t)blockIdx.y * chunk;
    if (base >= hi) return;
    const int64_t end = min(hi, base + (int64_t)chunk);
    const uint32_t bytes = (uint32_t)(R * W * 4);

    if ((bytes & 15u) == 0) {
        // regular case: one bulk async copy (TMA engine, no tensor map needed for a contiguous block)
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(bm)), "l"(bits + (size_t)m * R * W), "r"(bytes), "r"(smem_u32(bar))
                         : "memory");
        }
        // every thread waits for phase 0 of the barrier (HW sleep, not a spin on memory)
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(smem_u32(bar)) : "memory");
        }
    } else {
        // odd-sized bitmaps (R*W not a multiple of 4 words) cannot use the bulk engine: plain loads
        const uint32_t* src = bits + (size_t)m * R * W;
        for (int i = threadIdx.x; i < R * W; i += kDdaThreads) bm[i] = __ldg(src + i);
        __syncthreads();
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int nwarps = kDdaThreads / 32;
    for (int64_t b0 = base + 32 * warp; b0 < end; b0 += 32 * nwarps) {
        // each lane loads its own segment (512 B coalesced per warp)
        const int64_t mine = b0 + lane;
        int x0 = 0, y0 = 0, dx = 0, dy = 0;
        bool bad = false;
        if (mine < end) {
            const float4 s = __ldg(reinterpret_cast<const float4*>(segs) + mine);
            x0 = snap(s.x, bad); y0 = snap(s.y, bad);
            dx = snap(s.z, bad) - x0; dy = snap(s.w, bad) - y0;          // |d| <= 2^31 fits after the clamp
        }
        int my_hit = 0, my_first = -1;
        const int nb = (int)min((int64_t)32, end - b0);
        for (int j = 0; j < nb; ++j) {
            const int sx0 = __shfl_sync(0xffffffffu, x0, j), sy0 = __shfl_sync(0xffffffffu, y0, j);
            const int sdx = __shfl_sync(0xffffffffu, dx, j), sdy = __shfl_sync(0xffffffffu, dy, j);
            const bool sbad = __shfl_sync(0xffffffffu, (int)bad, j) != 0;
            int first = -1;
            if (sbad) first = 0;                                           // NaN coordinate: blocked at k = 0
            else {
                const long long adx = llabs((long long)sdx), ady = llabs((long long)sdy);
                const long long n = adx > ady ? adx : ady;
                const bool small = n <= 16384;
                for (long long k0 = 0; k0 <= n; k0 += 32) {
                    const long long k = k0 + lane;
                    bool blocked = false;
                    if (k <= n) {
                        long long cx = sx0, cy = sy0;
                        if (n) {
                            if (small) {
                                cx += floordiv32(2 * (int)k * sdx + (int)n, 2 * (int)n);
                                cy += floordiv32(2 * (int)k * sdy + (int)n, 2 * (int)n);
                            } else {
                                cx += floordiv64(2 * k * sdx + n, 2 * n);
                                cy += floordiv64(2 * k * sdy + n, 2 * n);
                            }
                        }
                        if (cx < 0 || cx >= R || cy < 0 || cy >= R) blocked = true;
                        else blocked = (bm[(int)cy * W + ((int)cx >> 5)] >> ((int)cx & 31)) & 1u;
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, blocked);
                    if (bal) { first = (int)k0 + (__ffs(bal) - 1); break; }   // early exit for the whole warp
                }
            }
            if (lane == j) { my_hit = first >= 0; my_first = first; }
        }
        if (mine < end) {
            verdict[mine] = (uint8_t)my_hit;
            if (first_hit) first_hit[mine] = my_first;
        }
    }
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_raster_circles_bits(const double* obs, const int32_t* obs_cnt, int32_t omax, int64_t n_maps,
                                         int32_t resolution, double inflate, uint32_t* bits, void* stream) {
    PPNET_REQUIRE(n_maps >= 0 && resolution > 0 && omax >= 0, "raster: bad sizes");
    if (n_maps == 0) return PPNET_OK;
    PPNET_REQUIRE(obs_cnt && bits && (obs || omax == 0), "raster: null pointer");
    const int W = (resolution + 31) / 32;
    const size_t smem = (size_t)resolution * W * 4;
    PPNET_REQUIRE(smem <= 220 * 1024, "raster: resolution too large for a shared-memory bitmap (max ~1300)");
    PPNET_REQUIRE((reinterpret_cast<uintptr_t>(bits) & 15) == 0, "raster: bits must be 16-byte aligned");
    if (smem > 48 * 1024)
        PPNET_CUDA(cudaFuncSetAttribute(raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    raster_kernel<<<(unsigned)n_maps, 256, smem, (cudaStream_t)stream>>>(obs, obs_cnt, omax, resolution, W, inflate, bits);
    PPNET_LAUNCH_CHECK("raster_kernel");
    return PPNET_OK;
}

extern "C" int ppnet_dda_gridcheck(const uint32_t* bits, int32_t resolution, int64_t n_maps, const float* segs_xy,
                                   int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map, uint8_t* verdict,
                                   int32_t* first_hit, void* stream) {
    PPNET_REQUIRE(n_maps >= 0 && n_segs >= 0 && resolution > 0, "dda: bad sizes");
    if (n_maps == 0 || n_segs == 0) return PPNET_OK;
    PPNET_REQUIRE(bits && segs_xy && verdict, "dda: null pointer");
    PPNET_REQUIRE(seg_off || segs_per_map * n_maps == n_segs, "dda: bad uniform grouping");
    PPNET_REQUIRE(seg_off == nullptr || segs_per_map > 0, "dda: pass the longest row in segs_per_map with a CSR");
    const int W = (resolution + 31) / 32;
    const size_t bm_bytes = (size_t)resolution * W * 4;
    PPNET_REQUIRE((reinterpret_cast<uintptr_t>(bits) & 15) == 0 && (reinterpret_cast<uintptr_t>(segs_xy) & 15) == 0,
                  "dda: bits and segs must be 16-byte aligned");
    const size_t smem = bm_bytes + 16;
    PPNET_REQUIRE(smem <= 220 * 1024, "dda: resolution too large for a shared-memory bitmap");
    // big bitmaps: amortise the staging over every segment of the map; small ones: more CTAs in flight
    const int chunk = bm_bytes >= 64 * 1024 ? 8192 : 1024;
    const int64_t chunks = (segs_per_map + chunk - 1) / chunk;
    PPNET_REQUIRE(chunks <= 65535, "dda: too many segments in one map");
    if (smem > 48 * 1024)
        PPNET_CUDA(cudaFuncSetAttribute(dda_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)n_maps, (unsigned)chunks);
    dda_kernel<<<grid, kDdaThreads, smem, (cudaStream_t)stream>>>(bits, resolution, W, segs_xy, seg_off, segs_per_map,
                                                                  chunk, verdict, first_hit);
    PPNET_LAUNCH_CHECK("dda_kernel");
    return PPNET_OK;
}
