// A4  Path.coord_euclidean2image   EDaGe-PP/Path.py:378-386   -- the grid-index rule
// A5  Path.free_space_bydirection  EDaGe-PP/Path.py:397-404   -- the float ray-march, and the four
//     driver loops of Path.path_space :113-134.
// Both bit-exact: float64 divide, add, round-half-to-even (rint), no contraction.
#include <algorithm>

#include "common.cuh"

namespace ppnet {

__device__ __forceinline__ int grid_cell(double v, double step, double off) {
    // int(np.round(v / step_len + mapoffset))
    return (int)rint(__dadd_rn(__ddiv_rn(v, step), off));
}

__global__ void grid_index_kernel(const double* __restrict__ pts, int64_t n, double step, double off,
                                  int32_t* __restrict__ idx) {
    // two values per thread: 16 B loads, 8 B stores
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(pts + i));
        int2 o;
        o.x = grid_cell(v.x, step, off);
        o.y = grid_cell(v.y, step, off);
        *reinterpret_cast<int2*>(idx + i) = o;
    } else if (i < n) {
        idx[i] = grid_cell(pts[i], step, off);
    }
}

// one thread per ray; rays of one corridor paint the same W x H byte image (benign same-value races)
__global__ void corridor_paint_kernel(const double* __restrict__ x0, const double* __restrict__ dir,
                                      const double* __restrict__ step_num, int rays_per_path, double step,
                                      double off, int W, int H, uint8_t value, uint8_t* __restrict__ space, int64_t p0) {
    const int64_t p = p0 + blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rays_per_path) return;
    const size_t ri = ((size_t)p * rays_per_path + r) * 2;
    const double px = x0[ri], py = x0[ri + 1], dx = dir[ri], dy = dir[ri + 1];
    const int ns = (int)rint(step_num[p]);                       // range(int(np.round(step_num)))
    uint8_t* img = space + (size_t)p * W * H;
    for (int i = 0; i < ns; ++i) {
        // x_init + i * dir : integer i times the f64 vector, then ONE add (not an accumulated +=)
        const double vx = __dadd_rn(px, __dmul_rn((double)i, dx));
        const double vy = __dadd_rn(py, __dmul_rn((double)i, dy));
        const int cx = grid_cell(vx, step, off), cy = grid_cell(vy, step, off);
        if (0 < cx && cx < W && 0 < cy && cy < H) img[(size_t)cx * H + cy] = value;   // strict: index 0 is rejected
        else break;                                              // `return space` ends the ray
    }
}

// A7 (mask part) / N1: the rigid resampling torchvision applies to a corridor mask -- T.RandomRotation(degrees=(d, d))
// (nearest, about the image centre, fill 0) followed by T.functional.affine(translate=(tx, ty)) (nearest) and a crop
// to the top-left Ro x Ro (Path.py:160-161, 175-178; MapGenerate.py:102-106).  Two nearest-neighbour passes, so
// the two roundings are kept separate:  (i, j) -> (rint(i - ty), rint(j - tx)) -> rotate by d about (Ws-1)/2 -> rint.
// Identical pixels to the reference on all golden corridors (float64 here, float32 grids there).
__global__ void mask_rigid_kernel(const uint8_t* __restrict__ src, int Ws, const double* __restrict__ angle_deg,
                                  const double* __restrict__ translate, int Ro, uint8_t* __restrict__ out, int64_t m0) {
    const int64_t m = m0 + blockIdx.y;
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= Ro * Ro) return;
    const int i = px / Ro, j = px % Ro;
    const double tx = translate[2 * m], ty = translate[2 * m + 1];
    const double j2 = rint((double)j - tx), i2 = rint((double)i - ty);
    uint8_t v = 0;
    if (j2 >= 0.0 && j2 < (double)Ws && i2 >= 0.0 && i2 < (double)Ws) {
        const double c = 0.5 * (double)(Ws - 1);
        const double th = angle_deg[m] / 180.0 * 3.14159265358979323846;
        double sn, cs;
        sincos(th, &sn, &cs);
        const double x = j2 - c, y = i2 - c;
        const double js = rint(cs * x - sn * y + c), is = rint(sn * x + cs * y + c);
        if (js >= 0.0 && js < (double)Ws && is >= 0.0 && is < (double)Ws) v = src[((size_t)m * Ws + (int)is) * Ws + (int)js];
    }
    out[((size_t)m * Ro + i) * Ro + j] = v;
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_mask_rigid(const uint8_t* src, int32_t Ws, const double* angle_deg, const double* translate,
                                int64_t n, int32_t Ro, uint8_t* out, void* stream) {
    PPNET_REQUIRE(n >= 0 && Ws > 0 && Ro > 0, "mask_rigid: bad sizes");
    if (n == 0) return PPNET_OK;
    PPNET_REQUIRE(src && angle_deg && translate && out, "mask_rigid: null pointer");
    for (int64_t m0 = 0; m0 < n; m0 += 65535) {                 // grid.y limit: launch in chunks of 65535 masks
        dim3 grid((unsigned)((Ro * Ro + 255) / 256), (unsigned)std::min<int64_t>(65535, n - m0));
        mask_rigid_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, Ws, angle_deg, translate, Ro, out, m0);
        PPNET_LAUNCH_CHECK("mask_rigid_kernel");
    }
    return PPNET_OK;
}

extern "C" int ppnet_grid_index_f64(const double* pts, int64_t n_values, double map_size, double resolution,
                                    double mapoffset, int32_t* idx, void* stream) {
    PPNET_REQUIRE(n_values >= 0, "grid_index: negative n");
    if (n_values == 0) return PPNET_OK;
    PPNET_REQUIRE(pts && idx, "grid_index: null pointer");
    PPNET_REQUIRE((reinterpret_cast<uintptr_t>(pts) & 15) == 0 && (reinterpret_cast<uintptr_t>(idx) & 7) == 0,
                  "grid_index: pts must be 16-byte and idx 8-byte aligned");
    const double step = map_size / resolution;                   // host IEEE divide == numpy's
    const int64_t pairs = (n_values + 1) / 2;
    grid_index_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pts, n_values, step,
                                                                                        mapoffset, idx);
    PPNET_LAUNCH_CHECK("grid_index_kernel");
    return PPNET_OK;
}

extern "C" int ppnet_corridor_paint(const double* x0, const double* dir, const double* step_num,
                                    int64_t n_paths, int32_t rays_per_path, double map_size,
                                    double resolution, double mapoffset, int32_t W, int32_t H,
                                    uint8_t value, uint8_t* space, void* stream) {
    PPNET_REQUIRE(n_paths >= 0 && rays_per_path >= 0 && W > 0 && H > 0, "corridor_paint: bad sizes");
    if (n_paths == 0 || rays_per_path == 0) return PPNET_OK;
    PPNET_REQUIRE(x0 && dir && step_num && space, "corridor_paint: null pointer");
    const double step = map_size / resolution;
    for (int64_t p0 = 0; p0 < n_paths; p0 += 65535) {           // grid.y limit: chunks of 65535 corridors
        dim3 grid((unsigned)((rays_per_path + 127) / 128), (unsigned)std::min<int64_t>(65535, n_paths - p0));
        corridor_paint_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x0, dir, step_num, rays_per_path, step,
                                                                       mapoffset, W, H, value, space, p0);
        PPNET_LAUNCH_CHECK("corridor_paint_kernel");
    }
    return PPNET_OK;
}
