// A14  MapGenerate.generate_map_randomly -- the obstacle clearance verdict (today's 76 % CPU hotspot)
//      EDaGe-PP/MapGenerate.py:132-143, float64, bit-exact.
//
// One CTA per map.  The map's odd-indexed path points (`if i % 2`, :139) are staged once into
// shared memory as double2 (8 KB at Np = 1000); each warp then owns candidates round-robin: lanes
// stride over the staged points, keep min of rn(rn(dx^2) + rn(dy^2)) -- scipy's euclidean is
// un-fused -- and a shuffle-min finishes it.  One square root per candidate: correctly rounded sqrt
// is monotone, so sqrt(min d^2) == min sqrt(d^2) bit for bit.  Accepted circles are compacted in
// candidate order with a ballot/popc scan (the reference appends in loop order).
#include <math_constants.h>

#include "common.cuh"

namespace ppnet {

constexpr int kClrThreads = 256;
constexpr int kClrWarps = kClrThreads / 32;

// shared with generate.cu: min squared distance of q to the staged odd points, warp-cooperative
__device__ __forceinline__ double warp_min_d2(const double2* __restrict__ pts, int n, double q0, double q1) {
    const int lane = threadIdx.x & 31;
    double m0 = CUDART_INF, m1 = CUDART_INF;
    int i = lane;
    for (; i + 32 < n; i += 64) {                 // two independent chains
        const double2 a = pts[i], b = pts[i + 32];
        const double ax = __dsub_rn(a.x, q0), ay = __dsub_rn(a.y, q1);
        const double bx = __dsub_rn(b.x, q0), by = __dsub_rn(b.y, q1);
        m0 = fmin(m0, __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
        m1 = fmin(m1, __dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)));
    }
    if (i < n) {
        const double2 a = pts[i];
        const double ax = __dsub_rn(a.x, q0), ay = __dsub_rn(a.y, q1);
        m0 = fmin(m0, __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
    }
    double m = fmin(m0, m1);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, s));
    return m;
}

__global__ void __launch_bounds__(kClrThreads)
clearance_kernel(const double* __restrict__ pathpt, int np, const double* __restrict__ cand, int O,
                 double M, double R, double c, uint8_t* __restrict__ accept, double* __restrict__ out,
                 int32_t* __restrict__ out_cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* pts = reinterpret_cast<double2*>(smem_raw);               // [n_odd]
    const int n_odd = np / 2;
    uint8_t* acc_s = reinterpret_cast<uint8_t*>(pts + n_odd);          // [O]
    const int64_t m = blockIdx.x;
    const double2* src = reinterpret_cast<const double2*>(pathpt + (size_t)m * np * 2);
    for (int i = threadIdx.x; i < n_odd; i += kClrThreads) pts[i] = __ldg(src + 2 * i + 1);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double thr_c = __dmul_rn(__ddiv_rn(c, M), R);                // c / M * R   (:142)
    const double* mc = cand + (size_t)m * O * 3;
    for (int j = warp; j < O; j += kClrWarps) {
        const double q0 = __dmul_rn(__ddiv_rn(mc[3 * j], M), R);       // coord / M * R   (:134)
        const double q1 = __dmul_rn(__ddiv_rn(mc[3 * j + 1], M), R);
        const double rimg = __dmul_rn(__ddiv_rn(mc[3 * j + 2], M), R); // :136
        const double m2 = warp_min_d2(pts, n_odd, q0, q1);
        // min(dis) of an empty list raises in the reference; we define "no points" as accept
        const bool ok = __dsqrt_rn(m2) > __dadd_rn(rimg, thr_c);       // :142
        if (lane == 0) acc_s[j] = ok ? 1 : 0;
    }
    __syncthreads();
    // ordered compaction by warp 0
    if (warp == 0) {
        int base = 0;
        for (int j0 = 0; j0 < O; j0 += 32) {
            const int j = j0 + lane;
            const bool ok = j < O && acc_s[j];
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (j < O && accept) accept[(size_t)m * O + j] = ok ? 1 : 0;
            if (ok && out) {
                const int k = base + __popc(bal & ((1u << lane) - 1));
                double* o = out + ((size_t)m * O + k) * 3;
                o[0] = __dmul_rn(__ddiv_rn(mc[3 * j + 1], M), R);      // [coord_img[1], coord_img[0], r]  (:143)
                o[1] = __dmul_rn(__ddiv_rn(mc[3 * j], M), R);
                o[2] = __dmul_rn(__ddiv_rn(mc[3 * j + 2], M), R);
            }
            base += __popc(bal);
        }
        if (lane == 0 && out_cnt) out_cnt[m] = base;
    }
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_clearance_filter_f64(const double* pathpt, int32_t np, const double* cand, int32_t O,
                                          int64_t n_maps, double map_size, double resolution,
                                          double clearance, uint8_t* accept, double* out,
                                          int32_t* out_cnt, void* stream) {
    PPNET_REQUIRE(n_maps >= 0 && np >= 0 && O >= 0, "clearance_filter: negative sizes");
    if (n_maps == 0) return PPNET_OK;
    PPNET_REQUIRE(cand || O == 0, "clearance_filter: cand is null");
    PPNET_REQUIRE(pathpt || np == 0, "clearance_filter: pathpt is null");
    PPNET_REQUIRE((reinterpret_cast<uintptr_t>(pathpt) & 15) == 0, "clearance_filter: pathpt must be 16-byte aligned");
    const size_t smem = sizeof(double2) * (size_t)(np / 2) + (size_t)O + 16;
    PPNET_REQUIRE(smem <= 200 * 1024, "clearance_filter: np/O too large for shared memory");
    if (smem > 48 * 1024)
        PPNET_CUDA(cudaFuncSetAttribute(clearance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    clearance_kernel<<<(unsigned)n_maps, kClrThreads, smem, (cudaStream_t)stream>>>(
        pathpt, np, cand, O, map_size, resolution, clearance, accept, out, out_cnt);
    PPNET_LAUNCH_CHECK("clearance_kernel");
    return PPNET_OK;
}
