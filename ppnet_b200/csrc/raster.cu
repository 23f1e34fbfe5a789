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

// ------------------------------------------------------------------------------------------ compose
// A15's return value + MapGenerate.py:111-113: map image f32[3][R][R], 1 = free (white), 0 = obstacle (black),
// optionally + the placed corridor mask, then A16's two 7x7 red stamps.
__global__ void bits_to_image_kernel(const uint32_t* __restrict__ bits, int R, int W, const float* __restrict__ add,
                                     float* __restrict__ img, int64_t m0) {
    const int64_t m = m0 + blockIdx.y;
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= R * R) return;
    const int i = px / R, j = px % R;
    const uint32_t w = bits[((size_t)m * R + i) * W + (j >> 5)];
    float v = ((w >> (j & 31)) & 1u) ? 0.0f : 1.0f;
    for (int ch = 0; ch < 3; ++ch) {
        const size_t o = (((size_t)m * 3 + ch) * R + i) * R + j;
        img[o] = add ? __fadd_rn(v, add[o]) : v;                       // map_img + path_space
    }
}

// A16  process_map.add_init_end_single  EDaGe-PP/process_map.py:119-145: (255, 0, 0) on the cells
// (round(p_r) + dj, round(p_c) + dk), dj, dk in -3..3, clipped to the image; round = half-to-even.
__global__ void add_init_end_kernel(float* __restrict__ img, int R, const double* __restrict__ init,
                                    const double* __restrict__ endp, int64_t n) {
    const int64_t m = blockIdx.x;
    if (m >= n) return;
    const int t = threadIdx.x;                                         // 2 points x 49 cells
    if (t >= 98) return;
    const double* p = (t < 49 ? init : endp) + 2 * m;
    const int c = t % 49, dj = c / 7 - 3, dk = c % 7 - 3;
    const double r0 = rint(p[0]), r1 = rint(p[1]);
    if (!(fabs(r0) < 1e9 && fabs(r1) < 1e9)) return;                   // NaN / huge: nothing lands in the image
    const int i = (int)r0 + dj, j = (int)r1 + dk;
    if (i < 0 || i >= R || j < 0 || j >= R) return;
    float* base = img + (size_t)m * 3 * R * R + (size_t)i * R + j;
    base[0] = 255.0f;
    base[(size_t)R * R] = 0.0f;
    base[(size_t)2 * R * R] = 0.0f;
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_bits_to_image(const uint32_t* bits, int32_t resolution, int64_t n_maps, const float* add,
                                   float* image, void* stream) {
    PPNET_REQUIRE(n_maps >= 0 && resolution > 0, "bits_to_image: bad sizes");
    if (n_maps == 0) return PPNET_OK;
    PPNET_REQUIRE(bits && image, "bits_to_image: null pointer");
    const int W = (resolution + 31) / 32;
    for (int64_t m0 = 0; m0 < n_maps; m0 += 65535) {            // grid.y limit: chunks of 65535 maps
        dim3 grid((unsigned)((resolution * resolution + 255) / 256), (unsigned)(n_maps - m0 < 65535 ? n_maps - m0 : 65535));
        bits_to_image_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(bits, resolution, W, add, image, m0);
        PPNET_LAUNCH_CHECK("bits_to_image_kernel");
    }
    return PPNET_OK;
}

extern "C" int ppnet_add_init_end(float* image, int32_t resolution, const double* init, const double* end,
                                  int64_t n_maps, void* stream) {
    PPNET_REQUIRE(n_maps >= 0 && resolution > 0, "add_init_end: bad sizes");
    if (n_maps == 0) return PPNET_OK;
    PPNET_REQUIRE(image && init && end, "add_init_end: null pointer");
    add_init_end_kernel<<<(unsigned)n_maps, 128, 0, (cudaStream_t)stream>>>(image, resolution, init, end, n_maps);
    PPNET_LAUNCH_CHECK("add_init_end_kernel");
    return PPNET_OK;
}

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
