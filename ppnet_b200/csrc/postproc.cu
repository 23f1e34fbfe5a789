// "Next" rows of the scope table: the label-mask rasteriser and the post-hoc path extraction that feed /
// follow the hot path.
//
//   N2  process_map.generate_gen_path   EDaGe-PP/process_map.py:148-163   every 5th label point -> 255
//   N3  process_map.extract_path        EDaGe-PP/process_map.py:293-365   greedy 8-neighbour walk on a heat-map
// (N1 / the mask half of N2 -- torchvision rotate + affine of a corridor mask -- is ppnet_mask_rigid in grid.cu.)
#include <math_constants.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace ppnet {

// one thread per (map, sampled point)
__global__ void path_mask_kernel(const double* __restrict__ pathpt, int np, int stride, int R, uint8_t* __restrict__ out,
                                 int64_t m0) {
    const int64_t m = m0 + blockIdx.y;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) * stride;        // step % stride == 0
    if (k >= np) return;
    const double* p = pathpt + ((size_t)m * np + k) * 2;
    const double r0 = rint(p[0]), c0 = rint(p[1]);                         // int(np.round(.)): half-to-even
    if (r0 > 0.0 && r0 < (double)R && c0 > 0.0 && c0 < (double)R)          // strict: row / column 0 are never painted
        out[((size_t)m * R + (int)r0) * R + (int)c0] = 255;
}

// one warp per image.  walk[] holds the accepted points in down-sampled coordinates; lanes share the history scan.
__global__ void __launch_bounds__(128)
extract_path_kernel(const float* __restrict__ mask, int h, int w, const double* __restrict__ init_state,
                    const double* __restrict__ end_state, double ds, int64_t n, int max_len, double* __restrict__ out,
                    int32_t* __restrict__ out_len, uint8_t* __restrict__ ok) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * 4 + warp;
    if (m >= n) return;
    const float* mk = mask + (size_t)m * h * w;
    double2* walk = reinterpret_cast<double2*>(out) + (size_t)m * (max_len + 2) + 1;   // slot 0 is init_state
    const double i0 = init_state[2 * m], i1 = init_state[2 * m + 1], e0s = end_state[2 * m], e1s = end_state[2 * m + 1];
    double n0 = __ddiv_rn(i0, ds), n1 = __ddiv_rn(i1, ds);                 // next_point = init / down_sample_rate
    const double e0 = __ddiv_rn(e0s, ds), e1 = __ddiv_rn(e1s, ds);
    const int mo0[8] = {0, 0, 1, -1, 1, 1, -1, -1}, mo1[8] = {1, -1, 0, 0, 1, -1, 1, -1};
    int L = 0;
    bool success = false;
    while (L < max_len) {
        float val[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {                                      // mask value at each of the 8 neighbours (0 outside)
            const double r = rint(__dadd_rn(n0, (double)mo0[i])), c = rint(__dadd_rn(n1, (double)mo1[i]));
            val[i] = (r >= 0.0 && r < (double)h && c >= 0.0 && c < (double)w) ? __ldg(mk + (int)r * w + (int)c) : 0.0f;
        }
        bool accepted = false;
        double c0 = 0.0, c1 = 0.0;
        for (;;) {
            int ci = 0;                                                    // first index of the maximum
            float best = val[0];
#pragma unroll
            for (int i = 1; i < 8; ++i) if (val[i] > best) { best = val[i]; ci = i; }
            if (!(best > 0.0f)) break;
            c0 = __dadd_rn(n0, (double)mo0[ci]); c1 = __dadd_rn(n1, (double)mo1[ci]);
            bool seen = false;                                             // visited, or within 1.5 of anything but the last two points
            for (int i = lane; i < L; i += 32) {
                const double2 q = walk[i];
                const double dx = __dsub_rn(c0, q.x), dy = __dsub_rn(c1, q.y);
                const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                seen = seen || (c0 == q.x && c1 == q.y) || (d <= 1.5 && i < L - 2);
            }
            if (__any_sync(0xffffffffu, seen)) {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (i == ci) val[i] = 0.0f;   // candidate_v[candidate_i] = 0
                continue;
            }
            accepted = true;
            break;
        }
        if (!accepted) break;                                              // 'inference failed'
        if (lane == 0) walk[L] = make_double2(c0, c1);
        __syncwarp();
        ++L;
        n0 = c0; n1 = c1;
        const double dx = __dsub_rn(c0, e0), dy = __dsub_rn(c1, e1);
        if (__dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) <= 2.5) { success = true; break; }
    }
    __syncwarp();
    if (success) {                                                         // [init_state, ds * walk ..., end_state]
        for (int i = lane; i < L; i += 32) { double2 q = walk[i]; q.x = __dmul_rn(q.x, ds); q.y = __dmul_rn(q.y, ds); walk[i] = q; }
        if (lane == 0) {
            walk[-1] = make_double2(i0, i1);
            walk[L] = make_double2(e0s, e1s);
        }
    }
    if (lane == 0) { out_len[m] = success ? L + 2 : 0; ok[m] = success ? 1 : 0; }
}

// N4  gerated_by_planners.generated_by_planners, the two label masks (EDaGe-PP/gerated_by_planners.py:88-157).
// One thread per ray: 360 rays around each end point, then +-normal rays at 100 samples per edge; a ray is K sub-pixel
// steps of 1/224 px (K = round(clearance / 2 * 224) = 502 at the reference's clearance).  Cells are stored in the
// orientation of the saved files (row = y, col = x).
__global__ void planner_mask_kernel(const double* __restrict__ wp, const int64_t* __restrict__ path_off, int K, int R,
                                    int pps, uint8_t* __restrict__ space, uint8_t* __restrict__ pathm, int64_t m0) {
    const int64_t m = m0 + blockIdx.y;
    const int64_t lo = path_off[m], L = path_off[m + 1] - lo;
    if (L < 2) return;
    const double2* p = reinterpret_cast<const double2*>(wp) + lo;
    const int n_cap = 720, n_band = (int)(L - 1) * pps * 2;
    const int ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_cap + n_band) return;
    const double step_img = 1.0 / 224.0;
    double ox, oy, dx, dy;
    if (ray < n_cap) {
        const bool at_end = ray >= 360;
        const int l = at_end ? ray - 360 : ray;
        const double2 a = at_end ? p[L - 1] : p[0], b = at_end ? p[0] : p[L - 1];
        double d0 = __dsub_rn(a.x, b.x), d1 = __dsub_rn(a.y, b.y);
        const double nn = __dsqrt_rn(__fma_rn(d1, d1, __dmul_rn(d0, d0)));            // np.linalg.norm (ddot)
        d0 = __ddiv_rn(d0, nn); d1 = __ddiv_rn(d1, nn);
        const double rad = __dmul_rn(__ddiv_rn((double)l, 180.0), 3.14159265358979323846);
        const double c = cos(rad), s = sin(rad);
        ox = a.x; oy = a.y;
        dx = __dadd_rn(__dmul_rn(c, d0), __dmul_rn(-s, d1));                          // coord_rotation (1-D np.dot)
        dy = __dadd_rn(__dmul_rn(s, d0), __dmul_rn(c, d1));
    } else {
        const int t = ray - n_cap;
        const int l = t / (2 * pps), j = (t % (2 * pps)) >> 1;
        const bool neg = t & 1;
        const double2 a = p[l], b = p[l + 1];
        double d0 = __dsub_rn(b.x, a.x), d1 = __dsub_rn(b.y, a.y);
        const double nn = __dsqrt_rn(__fma_rn(d1, d1, __dmul_rn(d0, d0)));
        d0 = __ddiv_rn(d0, nn); d1 = __ddiv_rn(d1, nn);
        const double ex = __dsub_rn(a.x, b.x), ey = __dsub_rn(a.y, b.y);
        const double step = __ddiv_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey))), (double)pps);   // scipy euclidean / 100
        const double js = __dmul_rn((double)j, step);
        ox = __dadd_rn(a.x, __dmul_rn(js, d0)); oy = __dadd_rn(a.y, __dmul_rn(js, d1));  // waypoint j of the edge
        if (!neg) {                                                                   // the waypoint itself -> mask_path (once)
            const double x = rint(ox), y = rint(oy);
            if (x > 0.0 && x < (double)(R - 1) && y > 0.0 && y < (double)(R - 1)) pathm[((size_t)m * R + (int)y) * R + (int)x] = 1;
        }
        dx = d1; dy = -d0;                                                            // dir = [dir[1], -dir[0]]
        if (neg) { dx = -dx; dy = -dy; }
    }
    int px = -1, py = -1;
    for (int k = 0; k < K; ++k) {
        const double ks = __dmul_rn((double)k, step_img);
        const double x = rint(__dadd_rn(ox, __dmul_rn(ks, dx))), y = rint(__dadd_rn(oy, __dmul_rn(ks, dy)));
        if (x > 0.0 && x < (double)(R - 1) && y > 0.0 && y < (double)(R - 1)) {
            const int ix = (int)x, iy = (int)y;
            if (ix != px || iy != py) { space[((size_t)m * R + iy) * R + ix] = 1; px = ix; py = iy; }   // ~224 steps per cell
        }
    }
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_planner_masks(const double* wp, const int64_t* path_off, int64_t n_paths, int64_t max_len,
                                   double clearance, int32_t resolution, int32_t points_per_seg, uint8_t* mask_space,
                                   uint8_t* mask_path, void* stream) {
    PPNET_REQUIRE(n_paths >= 0 && resolution > 1 && points_per_seg > 0 && max_len >= 0, "planner_masks: bad sizes");
    if (n_paths == 0) return PPNET_OK;
    PPNET_REQUIRE(wp && path_off && mask_space && mask_path, "planner_masks: null pointer");
    const size_t bytes = (size_t)n_paths * resolution * resolution;
    PPNET_CUDA(cudaMemsetAsync(mask_space, 0, bytes, (cudaStream_t)stream));
    PPNET_CUDA(cudaMemsetAsync(mask_path, 0, bytes, (cudaStream_t)stream));
    const int K = (int)nearbyint(clearance / 2.0 / (1.0 / 224.0));                    // round(clearance / 2 / step_img)
    const int64_t rays = 720 + std::max<int64_t>(max_len - 1, 0) * points_per_seg * 2;
    for (int64_t m0 = 0; m0 < n_paths; m0 += 65535) {           // grid.y limit: chunks of 65535 solutions
        dim3 grid((unsigned)((rays + 127) / 128), (unsigned)std::min<int64_t>(65535, n_paths - m0));
        planner_mask_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(wp, path_off, K, resolution, points_per_seg, mask_space,
                                                                    mask_path, m0);
        PPNET_LAUNCH_CHECK("planner_mask_kernel");
    }
    return PPNET_OK;
}

extern "C" int ppnet_path_mask(const double* pathpt, int32_t np, int64_t n_maps, int32_t stride, int32_t resolution,
                               uint8_t* out, void* stream) {
    PPNET_REQUIRE(n_maps >= 0 && np >= 0 && stride > 0 && resolution > 0, "path_mask: bad sizes");
    if (n_maps == 0) return PPNET_OK;
    PPNET_REQUIRE(pathpt && out, "path_mask: null pointer");
    PPNET_CUDA(cudaMemsetAsync(out, 0, (size_t)n_maps * resolution * resolution, (cudaStream_t)stream));
    const int pts = (np + stride - 1) / stride;
    if (pts == 0) return PPNET_OK;
    for (int64_t m0 = 0; m0 < n_maps; m0 += 65535) {            // grid.y limit: chunks of 65535 maps
        dim3 grid((unsigned)((pts + 127) / 128), (unsigned)std::min<int64_t>(65535, n_maps - m0));
        path_mask_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(pathpt, np, stride, resolution, out, m0);
        PPNET_LAUNCH_CHECK("path_mask_kernel");
    }
    return PPNET_OK;
}

extern "C" int ppnet_extract_path(const float* mask, int32_t h, int32_t w, const double* init_state, const double* end_state,
                                  double down_sample_rate, int64_t n, int32_t max_len, double* out, int32_t* out_len,
                                  uint8_t* ok, void* stream) {
    PPNET_REQUIRE(n >= 0 && h > 0 && w > 0 && max_len > 0 && down_sample_rate > 0, "extract_path: bad sizes");
    if (n == 0) return PPNET_OK;
    PPNET_REQUIRE(mask && init_state && end_state && out && out_len && ok, "extract_path: null pointer");
    extract_path_kernel<<<(unsigned)((n + 3) / 4), 128, 0, (cudaStream_t)stream>>>(mask, h, w, init_state, end_state,
                                                                                down_sample_rate, n, max_len, out, out_len, ok);
    PPNET_LAUNCH_CHECK("extract_path_kernel");
    return PPNET_OK;
}
