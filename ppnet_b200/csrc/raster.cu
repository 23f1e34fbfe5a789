// Bit-packed occupancy maps and the integer DDA grid check.
//
//   A15  plot_obstacles  EDaGe-PP/Path.py:36-49 -- restated GEOMETRICALLY (matplotlib/JPEG/PIL dither are
//        unpinned third-party code): pixel (row i, col j) is an obstacle iff its centre
//        (j + 0.5, i + 0.5) lies inside a disk (x, y, r + inflate), evaluated in float64 as
//        rn(rn(dx^2) + rn(dy^2)) <= rn((r + inflate)^2).  1 bit per pixel: bit (j & 31) of word j >> 5.
//
// One CTA per map, bitmap built in shared memory with atomicOr on word masks (one (circle, row) span per
// lane), then streamed out with 16-byte stores.  (The DDA that consumes these maps is in dda.cu.)
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
