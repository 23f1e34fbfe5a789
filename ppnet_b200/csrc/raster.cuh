// Shared-memory disk rasteriser used by raster.cu (stand-alone) and generate.cu (fused).
// Rule (geometric restatement of plot_obstacles, EDaGe-PP/Path.py:36-49 -- parity unpinned):
// pixel (row i, col j) is set iff rn(rn(dx^2) + rn(dy^2)) <= rn(rr^2), dx = (j+0.5)-x, dy = (i+0.5)-y.
#pragma once
#include "common.cuh"

namespace ppnet {

__device__ __forceinline__ bool px_inside(double ox, double dy2, double r2, int j) {
    const double dx = __dsub_rn(__dadd_rn((double)j, 0.5), ox);
    return __dadd_rn(__dmul_rn(dx, dx), dy2) <= r2;
}

__device__ __forceinline__ void or_span(uint32_t* row, int j0, int j1) {   // inclusive [j0, j1]
    const int w0 = j0 >> 5, w1 = j1 >> 5;
    for (int w = w0; w <= w1; ++w) {
        uint32_t mask = 0xffffffffu;
        if (w == w0) mask &= 0xffffffffu << (j0 & 31);
        if (w == w1) mask &= 0xffffffffu >> (31 - (j1 & 31));
        atomicOr(row + w, mask);
    }
}

// Row range [i0, i1) a disk can touch (empty for degenerate disks)
__device__ __forceinline__ void disk_rows(int R, double ox, double oy, double rr, int& i0, int& i1) {
    i0 = i1 = 0;
    if (!(rr > 0.0) || !(ox == ox) || !(oy == oy) || isinf(rr) || isinf(ox) || isinf(oy)) return;
    const double lo_f = floor(oy - rr - 1.0), hi_f = ceil(oy + rr + 1.0);
    i0 = lo_f < 0.0 ? 0 : (lo_f > (double)R ? R : (int)lo_f);
    i1 = hi_f > (double)R ? R : (hi_f < 0.0 ? 0 : (int)hi_f);
}

// One (disk, row) task.  The inside set of a row is an interval (every rounded op is monotone in |dx|); in real
// arithmetic it is the integers of [a, b], a = ox - h - 0.5, b = ox + h - 0.5, h = sqrt(r2 - dy2), and the rounded rule
// can only disagree with the real one for a pixel whose |dx| is within ~2 sqrt(2^-53) rr of h.  h comes from the SFU
// (float sqrt.approx, relative error < 3e-7 with the conversion), so when both a and b are farther than
// tol = 1e-3 + 1e-5 rr from every integer, [ceil a, floor b] IS the rounded rule's interval (tol/2 exceeds both the
// error of h and the disagreement band); otherwise (~4 tol of the rows, and every non-finite / clamped case, where the
// comparisons below are false) the ends are settled pixel by pixel with the exact rule.
__device__ __forceinline__ void raster_disk_row(uint32_t* bm, int R, int W, double ox, double oy, double rr, int i) {
    const double r2 = __dmul_rn(rr, rr);
    const double dy = __dsub_rn(__dadd_rn((double)i, 0.5), oy);
    const double dy2 = __dmul_rn(dy, dy);
    if (!(dy2 <= r2)) return;                              // even dx = 0 fails
    float hwf;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(hwf) : "f"((float)(r2 - dy2)));
    const double hw = (double)hwf;
    const double af = ox - hw - 0.5, bf = ox + hw - 0.5;
    const double a = ceil(af), b = floor(bf);
    const double tol = 1e-3 + 1e-5 * rr, da = a - af, db = bf - b;        // da, db in [0, 1)
    const bool sure = da > tol && da < 1.0 - tol && db > tol && db < 1.0 - tol && hw < 1e6;
    // clamp to [-2, R + 1] on the integer side: the conversion saturates (+-inf) and maps NaN to 0 (then `sure` is false)
    int j0 = max(-2, min(R + 1, __double2int_rz(a))), j1 = max(-2, min(R + 1, __double2int_rz(b)));
    if (!sure) {
        while (j0 > -2 && px_inside(ox, dy2, r2, j0 - 1)) --j0;
        while (j0 <= j1 && !px_inside(ox, dy2, r2, j0)) ++j0;
        while (j1 < R + 1 && px_inside(ox, dy2, r2, j1 + 1)) ++j1;
        while (j1 >= j0 && !px_inside(ox, dy2, r2, j1)) --j1;
    }
    j0 = max(j0, 0);
    j1 = min(j1, R - 1);
    if (j0 <= j1) or_span(bm + i * W, j0, j1);
}

// One warp rasterises one disk into bm[R][W] (shared memory): one row span per lane.
__device__ __forceinline__ void raster_disk_warp(uint32_t* bm, int R, int W, double ox, double oy, double rr) {
    const int lane = threadIdx.x & 31;
    int i0, i1;
    disk_rows(R, ox, oy, rr, i0, i1);
    for (int i = i0 + lane; i < i1; i += 32) raster_disk_row(bm, R, W, ox, oy, rr, i);
}

__device__ __forceinline__ void store_bitmap(const uint32_t* bm, uint32_t* dst, int words) {
    if ((words & 3) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(bm);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (int i = threadIdx.x; i < words / 4; i += blockDim.x) d4[i] = s4[i];
    } else {
        for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = bm[i];
    }
}

}  // namespace ppnet
