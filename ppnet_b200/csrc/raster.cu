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

// ------------------------------------------------------------------------------------------ canvas model (A15 pin)
// plot_obstacles (EDaGe-PP/Path.py:36-49) draws on matplotlib's default 6.4 x 4.8 in figure at dpi 90 -- a 576 x 432
// canvas whose axes occupy [0.125, 0.9] x [0.11, 0.88] of the figure -- then crops rows 53:383 / cols 73:517 and resizes
// the 330 x 444 crop to R x R (torchvision Resize, bilinear, no antialias in the reference's torchvision 0.12).  In
// canvas pixels the data -> pixel map is px = 72 + x / size_w * 446.4, py = 51.84 + y / size_h * 332.64 (y axis inverted
// by ax.axis(ymin=size[1], ymax=0)), so a data circle is an ELLIPSE with semi-axes r * 446.4 / size_w and
// r * 332.64 / size_h.  This mode restates exactly that geometry: canvas pixel black iff its centre is inside an
// ellipse; bilinear resize of the binary crop (ATen's align_corners = False index rule, float32); pixel occupied iff the
// resized value < 0.5.  What it leaves out is what cannot be pinned (anti-aliasing, JPEG, the Floyd-Steinberg dither).
constexpr int kCanvasRows = 330, kCanvasCols = 444, kCanvasW = (kCanvasCols + 31) / 32;   // the crop [53:383, 73:517]

__global__ void __launch_bounds__(256)
raster_canvas_kernel(const double* __restrict__ obs, const int32_t* __restrict__ obs_cnt, int omax, double size_w, double size_h,
                     int R, int W, double inflate, uint32_t* __restrict__ bits) {
    __shared__ uint32_t cv[kCanvasRows * kCanvasW];           // 1 = black canvas pixel
    const int64_t m = blockIdx.x;
    for (int i = threadIdx.x; i < kCanvasRows * kCanvasW; i += blockDim.x) cv[i] = 0u;
    __syncthreads();
    const int cnt = min(obs_cnt[m], omax);
    const double* mo = obs + (size_t)m * omax * 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const double sx = 446.4 / size_w, sy = 332.64 / size_h;
    for (int k = warp; k < cnt; k += nwarps) {
        const double rr = mo[3 * k + 2] + inflate;
        const double cx = 72.0 + mo[3 * k] * sx, cy = 51.84 + mo[3 * k + 1] * sy, ax = rr * sx, ay = rr * sy;
        if (!(rr > 0.0) || !(cx == cx) || !(cy == cy) || isinf(cx) || isinf(cy) || isinf(rr)) continue;
        const int v0 = max(0, (int)floor(fmax(cy - ay - 1.0 - 53.0, -1.0))), v1 = min(kCanvasRows, (int)ceil(fmin(cy + ay + 1.0 - 53.0, 1e6)));
        for (int v = v0 + lane; v < v1; v += 32) {
            const double ny = ((double)(53 + v) + 0.5 - cy) / ay;
            const double q = 1.0 - ny * ny;
            if (!(q >= 0.0)) continue;
            const double hw = ax * sqrt(q);
            // candidate span from the real-arithmetic interval, ends settled with the rule itself
            auto inside = [&](int u) { const double nx = ((double)(73 + u) + 0.5 - cx) / ax; return nx * nx + ny * ny <= 1.0; };
            int u0 = (int)fmax(fmin(ceil(cx - hw - 0.5 - 73.0), 1e6), -2.0), u1 = (int)fmax(fmin(floor(cx + hw - 0.5 - 73.0), 1e6), -2.0);
            while (u0 > -2 && inside(u0 - 1)) --u0;
            while (u0 <= u1 && !inside(u0)) ++u0;
            while (u1 < kCanvasCols + 1 && inside(u1 + 1)) ++u1;
            while (u1 >= u0 && !inside(u1)) --u1;
            u0 = max(u0, 0);
            u1 = min(u1, kCanvasCols - 1);
            if (u0 <= u1) or_span(cv + v * kCanvasW, u0, u1);
        }
    }
    __syncthreads();
    // bilinear resize of the (white = 1) crop, ATen upsample_bilinear2d index rule (align_corners = False), then < 0.5
    const float sch = (float)kCanvasRows / (float)R, scw = (float)kCanvasCols / (float)R;
    auto white = [&](int v, int u) { return ((cv[v * kCanvasW + (u >> 5)] >> (u & 31)) & 1u) ? 0.0f : 1.0f; };
    for (int task = warp; task < R * W; task += nwarps) {
        const int i = task / W, j = (task % W) * 32 + lane;
        bool occ = false;
        if (j < R) {
            const float fy = fmaxf(__fsub_rn(__fmul_rn(sch, (float)i + 0.5f), 0.5f), 0.0f);
            const float fx = fmaxf(__fsub_rn(__fmul_rn(scw, (float)j + 0.5f), 0.5f), 0.0f);
            const int y0 = min((int)fy, kCanvasRows - 1), x0 = min((int)fx, kCanvasCols - 1);
            const int y1 = min(y0 + 1, kCanvasRows - 1), x1 = min(x0 + 1, kCanvasCols - 1);
            const float ly1 = __fsub_rn(fy, (float)y0), lx1 = __fsub_rn(fx, (float)x0);
            const float ly0 = __fsub_rn(1.0f, ly1), lx0 = __fsub_rn(1.0f, lx1);
            const float top = __fadd_rn(__fmul_rn(lx0, white(y0, x0)), __fmul_rn(lx1, white(y0, x1)));
            const float bot = __fadd_rn(__fmul_rn(lx0, white(y1, x0)), __fmul_rn(lx1, white(y1, x1)));
            occ = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot)) < 0.5f;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, occ);
        if (lane == 0) bits[(size_t)m * R * W + task] = word;
    }
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

extern "C" int ppnet_raster_canvas_bits(const double* obs, const int32_t* obs_cnt, int32_t omax, int64_t n_maps, double size_w,
                                        double size_h, int32_t resolution, double inflate, uint32_t* bits, void* stream) {
    PPNET_REQUIRE(n_maps >= 0 && resolution > 0 && omax >= 0 && size_w > 0 && size_h > 0, "raster_canvas: bad sizes");
    if (n_maps == 0) return PPNET_OK;
    PPNET_REQUIRE(obs_cnt && bits && (obs || omax == 0), "raster_canvas: null pointer");
    raster_canvas_kernel<<<(unsigned)n_maps, 256, 0, (cudaStream_t)stream>>>(obs, obs_cnt, omax, size_w, size_h, resolution,
                                                                            (resolution + 31) / 32, inflate, bits);
    PPNET_LAUNCH_CHECK("raster_canvas_kernel");
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
