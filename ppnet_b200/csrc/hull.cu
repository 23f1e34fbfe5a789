// A6  Path.convexhull  EDaGe-PP/Path.py:388-395  ("convehull" pruning): 2-D convex hull of the path's
// integer grid cells.  scipy/Qhull on integer input returns exactly the strict corners in CCW order;
// a strict monotone chain (pop while cross <= 0) gives the same vertex set, so that is
// what runs here: one CTA per path -- bitonic sort of packed (x, y) keys in shared memory, then one
// thread walks the chain (~2N steps).  Output starts at the lexicographically smallest point.
#include "common.cuh"

namespace ppnet {

constexpr int kHullThreads = 256;

__device__ __forceinline__ long long crossz(long long ox, long long oy, long long ax, long long ay, long long bx,
                                            long long by) {
    return (ax - ox) * (by - oy) - (ay - oy) * (bx - ox);
}

__global__ void __launch_bounds__(kHullThreads)
hull_kernel(const int32_t* __restrict__ pts, int np, int npow2, int hmax, int32_t* __restrict__ hull,
            int32_t* __restrict__ hull_cnt) {
    extern __shared__ __align__(16) unsigned char hraw[];
    unsigned long long* key = reinterpret_cast<unsigned long long*>(hraw);        // [npow2]
    int2* stack = reinterpret_cast<int2*>(key + npow2);                            // [np + 1]
    const int64_t p = blockIdx.x;
    const int2* src = reinterpret_cast<const int2*>(pts) + (size_t)p * np;
    for (int i = threadIdx.x; i < npow2; i += kHullThreads) {
        unsigned long long k = ~0ull;                                              // padding sorts last
        if (i < np) {
            const int2 v = src[i];
            k = ((unsigned long long)((uint32_t)v.x ^ 0x80000000u) << 32) | (uint32_t)((uint32_t)v.y ^ 0x80000000u);
        }
        key[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow2; i += kHullThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = key[i], b = key[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { key[i] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {
        auto X = [&](int i) { return (long long)(int32_t)((uint32_t)(key[i] >> 32) ^ 0x80000000u); };
        auto Y = [&](int i) { return (long long)(int32_t)((uint32_t)key[i] ^ 0x80000000u); };
        int n = 0;                                                                 // stack size
        // lower hull
        for (int i = 0; i < np; ++i) {
            const long long x = X(i), y = Y(i);
            while (n >= 2 && crossz(stack[n - 2].x, stack[n - 2].y, stack[n - 1].x, stack[n - 1].y, x, y) <= 0) --n;
            stack[n++] = make_int2((int)x, (int)y);
        }
        // upper hull (the last point of the lower hull is the first of the upper one)
        const int lower_n = n;
        for (int i = np - 2; i >= 0; --i) {
            const long long x = X(i), y = Y(i);
            while (n > lower_n && crossz(stack[n - 2].x, stack[n - 2].y, stack[n - 1].x, stack[n - 1].y, x, y) <= 0) --n;
            stack[n++] = make_int2((int)x, (int)y);
        }
        if (n > 1) --n;                                                            // last == first
        // degenerate inputs: all points equal -> 1 vertex; collinear -> the two extremes
        if (n == 2 && stack[0].x == stack[1].x && stack[0].y == stack[1].y) n = 1;
        const int cnt = min(n, hmax);
        for (int i = 0; i < cnt; ++i) {
            hull[((size_t)p * hmax + i) * 2] = stack[i].x;
            hull[((size_t)p * hmax + i) * 2 + 1] = stack[i].y;
        }
        hull_cnt[p] = n;                                                           // > hmax tells the caller it overflowed
    }
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_hull2d_i32(const int32_t* pts, int32_t np, int64_t n_paths, int32_t hmax, int32_t* hull,
                                int32_t* hull_cnt, void* stream) {
    PPNET_REQUIRE(n_paths >= 0 && np > 0 && hmax > 0, "hull2d: bad sizes");
    if (n_paths == 0) return PPNET_OK;
    PPNET_REQUIRE(pts && hull && hull_cnt, "hull2d: null pointer");
    PPNET_REQUIRE((reinterpret_cast<uintptr_t>(pts) & 7) == 0, "hull2d: pts must be 8-byte aligned");
    int npow2 = 1;
    while (npow2 < np) npow2 <<= 1;
    const size_t smem = sizeof(unsigned long long) * (size_t)npow2 + sizeof(int2) * (size_t)(np + 2);
    PPNET_REQUIRE(smem <= 200 * 1024, "hull2d: too many points per path for shared memory");
    if (smem > 48 * 1024)
        PPNET_CUDA(cudaFuncSetAttribute(hull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hull_kernel<<<(unsigned)n_paths, kHullThreads, smem, (cudaStream_t)stream>>>(pts, np, npow2, hmax, hull, hull_cnt);
    PPNET_LAUNCH_CHECK("hull_kernel");
    return PPNET_OK;
}
