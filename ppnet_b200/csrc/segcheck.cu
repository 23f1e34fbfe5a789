// Segment-vs-circles collision verdicts, both reference flavours, bit-exact.
//
//   A11  process_map.collision_check_circle_edge   EDaGe-PP/process_map.py:383-425   (float64)
//   A12  neuralplanner.collision_check_circle_edge experiments/MPNet/neuralplanner.py:43-69
//        + steerTo :86-92, feasibility_check :96-102, lvc :123-138                   (float32)
//
// Layout: one CTA owns one (map, chunk-of-segments).  The map's circles are staged once into shared
// memory in "decision form" {o.x, o.y, thr, T}: centre already rounded through float32 (the
// reference builds it with torch.tensor([ox, oy])), thr = r + clearance/2 in the flavour's
// precision, and T = the exact image of thr under sqrt (sqrt_rn(q) < thr <=> q < T), which removes
// the square root from the vertex test without changing a single verdict.  One thread walks one
// segment over the staged circles (shared-memory broadcast reads, no bank conflicts) and stops at
// its first hit -- the reference's early `return True`.
//
// The expensive part of the edge test (2 sqrt + 4 div to normalise p-s and p-e) only matters when
// |dis| < thr; and even then it can only come out negative when the foot of the perpendicular lies
// between s and e.  `tt` below is the un-normalised projection of (o - s) on d; when it is outside
// [0, |d|^2] by a margin ~1e7 ulp wide, normalised (p-s).(p-e) is within 1e-6 of +1 and the exact
// computation is skipped.  Everything inside the margin runs the reference's operation sequence
// verbatim (same roundings, same NaN behaviour).
#include <math_constants.h>

#include "common.cuh"

namespace ppnet {

template <typename T>
struct Circle {   // 32 B (double) / 16 B (float): one LDS.128 pair / one LDS.128
    T ox, oy, thr, T2;
};

template <typename T> struct FP;
template <> struct FP<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt_(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double abs_(double a) { return fabs(a); }
    static constexpr double kMargin = 1e-9;
};
template <> struct FP<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt_(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float abs_(float a) { return fabsf(a); }
    static constexpr float kMargin = 1e-3f;
};

// np.dot / np.linalg.norm^2 of the flavour: f64 follows dot_mode, f32 is always un-fused
template <typename T, int MODE> struct Dot;
template <int MODE> struct Dot<double, MODE> {
    static __device__ __forceinline__ double f(double a0, double a1, double b0, double b1) {
        return dot2<MODE>(a0, a1, b0, b1);
    }
};
template <int MODE> struct Dot<float, MODE> {
    static __device__ __forceinline__ float f(float a0, float a1, float b0, float b1) {
        return dot2f(a0, a1, b0, b1);
    }
};

// Per-segment state that does not depend on the circle
template <typename T>
struct SegState {
    T s0, s1, e0, e1;   // (x, y) after the flavour's swap
    T d0, d1, L, n0, n1;
    T mag;              // |s|+|e| magnitude for the skip margin
    bool oob;
};

template <typename T, int MODE, bool SWAP>
__device__ __forceinline__ SegState<T> seg_setup(T a0, T a1, T b0, T b1, T bound) {
    using F = FP<T>;
    SegState<T> g;
    // bounds test on the raw inputs, *before* the swap (process_map.py:384-387 / neuralplanner.py:44-47)
    g.oob = (a0 < T(0)) || (a1 > bound) || (b0 < T(0)) || (b1 > bound);
    if (SWAP) { g.s0 = a1; g.s1 = a0; g.e0 = b1; g.e1 = b0; }
    else      { g.s0 = a0; g.s1 = a1; g.e0 = b0; g.e1 = b1; }
    g.d0 = F::sub(g.e0, g.s0);
    g.d1 = F::sub(g.e1, g.s1);
    g.L = F::sqrt_(Dot<T, MODE>::f(g.d0, g.d1, g.d0, g.d1));     // np.linalg.norm(dir)
    g.n0 = F::div(g.d1, g.L);                                    // [dir[1], -dir[0]] / norm
    g.n1 = F::div(-g.d0, g.L);
    g.mag = F::abs_(g.s0) + F::abs_(g.s1) + F::abs_(g.e0) + F::abs_(g.e1) + T(1);
    return g;
}

// One (segment, circle) pair.  Returns true on collision.
template <typename T, int MODE>
__device__ __forceinline__ bool pair_hit(const SegState<T>& g, const Circle<T>& c) {
    using F = FP<T>;
    // vertex test on e only:  euclidean(e, o) < thr   <=>   rn(v0^2 + v1^2) < T2   (un-fused, scipy)
    const T v0 = F::sub(g.e0, c.ox), v1 = F::sub(g.e1, c.oy);
    const T q = F::add(F::mul(v0, v0), F::mul(v1, v1));
    if (q < c.T2) return true;
    // signed offset from the line;  |dis| is invariant under the reference's `dir = -dir` flips
    const T q0 = F::sub(c.ox, g.s0), q1 = F::sub(c.oy, g.s1);
    const T dis = Dot<T, MODE>::f(g.n0, g.n1, q0, q1);
    const T a = F::abs_(dis);
    if (!(a < c.thr)) return false;        // `dis < size + clearance/2 and ...` short-circuits
    // foot clearly beyond either end => normalised (p-s).(p-e) ~ +1, cannot be < 0
    {
        const T tt = q0 * g.d0 + q1 * g.d1;
        const T m = F::kMargin * (g.mag + F::abs_(c.ox) + F::abs_(c.oy));
        if (tt > g.L * (g.L + m) || tt < -(g.L * m)) return false;
    }
    // exact tail: projection = o + |dis| * dir(flipped)  ==  o - dis * n   (sign-symmetric roundings)
    const T p0 = F::sub(c.ox, F::mul(dis, g.n0)), p1 = F::sub(c.oy, F::mul(dis, g.n1));
    T u0 = F::sub(p0, g.s0), u1 = F::sub(p1, g.s1);
    const T nu = F::sqrt_(Dot<T, MODE>::f(u0, u1, u0, u1));
    u0 = F::div(u0, nu); u1 = F::div(u1, nu);
    T w0 = F::sub(p0, g.e0), w1 = F::sub(p1, g.e1);
    const T nw = F::sqrt_(Dot<T, MODE>::f(w0, w1, w0, w1));
    w0 = F::div(w0, nw); w1 = F::div(w1, nw);
    return Dot<T, MODE>::f(u0, u1, w0, w1) < T(0);
}

template <typename T>
__device__ __forceinline__ Circle<T> make_circle(const double* __restrict__ o, double clearance) {
    Circle<T> c;
    // torch.tensor([ox, oy]) => float32 centre in BOTH flavours (process_map.py:396, neuralplanner.py:53)
    c.ox = (T)(float)o[0];
    c.oy = (T)(float)o[1];
    // size + clearance/2 in Python floats (f64); the f32 flavour compares it as f32 (NEP 50)
    c.thr = (T)__dadd_rn(o[2], __ddiv_rn(clearance, 2.0));
    c.T2 = sqrt_lt_threshold(c.thr);
    return c;
}

constexpr int kCircTile = 128;       // circles staged per pass
constexpr int kSegThreads = 256;
template <typename T> struct SegCfg;      // segments per thread: registers (f64) vs ILP (f32)
template <> struct SegCfg<double> { static constexpr int kPerThread = 1; };
template <> struct SegCfg<float> { static constexpr int kPerThread = 2; };

// ---- exact-safe spatial culling -------------------------------------------------------------------------
// A circle can only collide with a segment whose bounding box meets the circle's box inflated by thr:
//   vertex test  |e - o| < thr            => o is within thr of the endpoint e,
//   edge test    |dis| < thr and the foot of the perpendicular strictly between s and e
//                                         => o is within thr of a point of the segment.
// Both boxes are further inflated by a margin thousands of ulps wide (kCull * magnitude), so a pair that the
// grid separates is separated by far more than any rounding of the reference's operation sequence could
// bridge (for the edge test the foot then lies beyond an end by >= the margin, where normalised
// (p-s).(p-e) is within 1e-6 of +1).  Culled pairs are exactly the pairs whose verdict is "no".
// The staged circles are binned into an 8x8 grid over [0, bound]^2 (cell lists in shared memory); a segment
// only walks the lists of the cells its box touches.  The bin function is monotone, so two intersecting
// intervals always share a bin.  Long segments (box > kMaxQueryCells bins) and non-finite coordinates take
// the plain loop over all staged circles.
constexpr int kGridN = 8;
constexpr int kGridCells = kGridN * kGridN;
constexpr int kMaxQueryCells = 12;
template <typename T> struct Cull;
template <> struct Cull<double> { static constexpr double k = 1e-9; };
template <> struct Cull<float> { static constexpr float k = 1e-3f; };

__device__ __forceinline__ int bin_of(float v, float scale) {          // monotone non-decreasing in v
    return (int)fminf(fmaxf(v * scale, 0.0f), (float)(kGridN - 1));
}

template <typename T> struct Vec4;   // 4 coordinates of one segment
template <> struct Vec4<double> {
    static __device__ __forceinline__ void load(const double* p, double& a, double& b, double& c, double& d) {
        const double2 x = __ldg(reinterpret_cast<const double2*>(p));
        const double2 y = __ldg(reinterpret_cast<const double2*>(p) + 1);
        a = x.x; b = x.y; c = y.x; d = y.y;
    }
};
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float& a, float& b, float& c, float& d) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(p));
        a = x.x; b = x.y; c = x.z; d = x.w;
    }
};

// grid = (n_maps, chunks_per_map_max)
template <typename T, int MODE, bool SWAP, bool STEER>
__global__ void __launch_bounds__(kSegThreads)
segcheck_kernel(const T* __restrict__ pts, const int64_t* __restrict__ seg_off, int64_t segs_per_map,
                const double* __restrict__ obs, const int32_t* __restrict__ obs_cnt, int omax,
                double clearance, T bound, uint8_t* __restrict__ verdict, uint8_t* __restrict__ steer) {
    constexpr int kSegPerThread = SegCfg<T>::kPerThread;
    constexpr int kSegChunk = kSegThreads * kSegPerThread;
    const int m = blockIdx.x;
    const int64_t lo = seg_off ? seg_off[m] : (int64_t)m * segs_per_map;
    const int64_t hi = seg_off ? seg_off[m + 1] : lo + segs_per_map;
    const int64_t base = lo + (int64_t)blockIdx.y * kSegChunk;
    if (base >= hi) return;                       // whole CTA exits together

    __shared__ Circle<T> sc[kCircTile];
    __shared__ uint8_t cell_idx[kGridCells][kCircTile];
    __shared__ int cell_cnt[kGridCells];
    const int cnt = min(obs_cnt[m], omax);
    const bool use_grid = bound > T(0) && bound < T(1e30);
    const float bscale = use_grid ? (float)kGridN / (float)bound : 0.0f;
    const double* __restrict__ mobs = obs + (size_t)m * omax * 3;

    SegState<T> g[kSegPerThread];
    bool live[kSegPerThread], hit[kSegPerThread];
#pragma unroll
    for (int k = 0; k < kSegPerThread; ++k) {
        const int64_t i = base + threadIdx.x + (int64_t)k * kSegThreads;
        live[k] = i < hi;
        hit[k] = false;
        if (live[k]) {
            T a0, a1, b0, b1;
            Vec4<T>::load(pts + 4 * i, a0, a1, b0, b1);
            g[k] = seg_setup<T, MODE, SWAP>(a0, a1, b0, b1, bound);
            hit[k] = g[k].oob;                    // `return True` before any circle is looked at
        }
    }

    for (int t0 = 0; t0 < cnt; t0 += kCircTile) {
        const int nt = min(kCircTile, cnt - t0);
        __syncthreads();
        if (threadIdx.x < kGridCells) cell_cnt[threadIdx.x] = 0;
        __syncthreads();
        if (threadIdx.x < nt) {
            const Circle<T> c = make_circle<T>(mobs + 3 * (t0 + threadIdx.x), clearance);
            sc[threadIdx.x] = c;
            // circles that can never answer "hit" (thr <= 0 or NaN, non-finite centre) are not binned at all
            if (use_grid && c.thr > T(0) && FP<T>::abs_(c.ox) < T(1e30) && FP<T>::abs_(c.oy) < T(1e30)) {
                const T h = c.thr + Cull<T>::k * (FP<T>::abs_(c.ox) + FP<T>::abs_(c.oy) + c.thr + bound);
                const int x0 = bin_of((float)(c.ox - h), bscale), x1 = bin_of((float)(c.ox + h), bscale);
                const int y0 = bin_of((float)(c.oy - h), bscale), y1 = bin_of((float)(c.oy + h), bscale);
                for (int by = y0; by <= y1; ++by)
                    for (int bx = x0; bx <= x1; ++bx) {
                        const int cell = by * kGridN + bx;
                        cell_idx[cell][atomicAdd(&cell_cnt[cell], 1)] = (uint8_t)threadIdx.x;
                    }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSegPerThread; ++k) {
            if (!live[k] || hit[k]) continue;
            const SegState<T>& q = g[k];
            bool brute = !use_grid;
            int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
            if (!brute) {
                const T ms = Cull<T>::k * (q.mag + bound);
                const T lox = (q.s0 < q.e0 ? q.s0 : q.e0) - ms, hix = (q.s0 < q.e0 ? q.e0 : q.s0) + ms;
                const T loy = (q.s1 < q.e1 ? q.s1 : q.e1) - ms, hiy = (q.s1 < q.e1 ? q.e1 : q.s1) + ms;
                // NaN / inf coordinates fail this test and take the plain loop
                if (!(FP<T>::abs_(lox) < T(1e30) && FP<T>::abs_(hix) < T(1e30) && FP<T>::abs_(loy) < T(1e30) &&
                      FP<T>::abs_(hiy) < T(1e30))) brute = true;
                else {
                    x0 = bin_of((float)lox, bscale); x1 = bin_of((float)hix, bscale);
                    y0 = bin_of((float)loy, bscale); y1 = bin_of((float)hiy, bscale);
                    if ((x1 - x0 + 1) * (y1 - y0 + 1) > kMaxQueryCells) brute = true;
                }
            }
            if (brute) { x0 = x1 = y0 = y1 = 0; }                  // one pseudo-bin holding every staged circle
            for (int by = y0; by <= y1 && !hit[k]; ++by)
                for (int bx = x0; bx <= x1 && !hit[k]; ++bx) {
                    const int cell = by * kGridN + bx;
                    const int nc = brute ? nt : cell_cnt[cell];
                    for (int t = 0; t < nc; ++t) {
                        const int j = brute ? t : (int)cell_idx[cell][t];
                        if (pair_hit<T, MODE>(q, sc[j])) { hit[k] = true; break; }
                    }
                }
        }
    }

#pragma unroll
    for (int k = 0; k < kSegPerThread; ++k) {
        const int64_t i = base + threadIdx.x + (int64_t)k * kSegThreads;
        if (!live[k]) continue;
        if (verdict) verdict[i] = hit[k] ? 1 : 0;
        if (STEER && steer) {
            // steerTo: dist = euclidean(start, end) in f32 (un-fused); 0 iff dist > 0 and blocked
            using F = FP<T>;
            const T x = F::sub(g[k].s0, g[k].e0), y = F::sub(g[k].s1, g[k].e1);
            const T dist = F::sqrt_(F::add(F::mul(x, x), F::mul(y, y)));
            steer[i] = (dist > T(0) && hit[k]) ? 0 : 1;
        }
    }
}

template <typename T>
static int check_common(const T* pts, int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map,
                        int64_t n_maps, const double* obs, const int32_t* obs_cnt, int32_t omax,
                        int64_t* chunks) {
    PPNET_REQUIRE(n_segs >= 0 && n_maps >= 0, "segcheck: negative sizes");
    PPNET_REQUIRE(n_maps <= 2147483647LL, "segcheck: too many maps for one launch");
    if (n_segs == 0 || n_maps == 0) { *chunks = 0; return PPNET_OK; }
    PPNET_REQUIRE(pts && obs_cnt, "segcheck: null pointer");
    PPNET_REQUIRE(omax >= 0 && (omax == 0 || obs), "segcheck: obs is null but omax > 0");
    PPNET_REQUIRE(seg_off || segs_per_map > 0, "segcheck: need seg_off or segs_per_map");
    PPNET_REQUIRE(seg_off || segs_per_map * n_maps == n_segs,
                  "segcheck: uniform grouping needs n_segs == n_maps * segs_per_map");
    PPNET_REQUIRE((reinterpret_cast<uintptr_t>(pts) & 15) == 0, "segcheck: pts must be 16-byte aligned");
    return PPNET_OK;
}

// largest per-map segment count decides grid.x; with a CSR we cannot know it without a device
// read, so the caller passes max_segs_per_map through segs_per_map when seg_off != NULL.
template <typename T, bool SWAP, bool STEER>
static int launch(const T* pts, int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map,
                  int64_t n_maps, const double* obs, const int32_t* obs_cnt, int32_t omax,
                  double clearance, double bound, int32_t dot_mode, uint8_t* verdict, uint8_t* steer,
                  cudaStream_t st) {
    int64_t dummy = 1;
    int rc = check_common(pts, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, &dummy);
    if (rc != PPNET_OK || dummy == 0) return rc;
    constexpr int kSegChunk = kSegThreads * SegCfg<T>::kPerThread;
    const int64_t per_map = segs_per_map > 0 ? segs_per_map : n_segs;
    const int64_t chunks = (per_map + kSegChunk - 1) / kSegChunk;
    PPNET_REQUIRE(chunks <= 65535, "segcheck: more than 65535*512 segments in one map");
    dim3 grid((unsigned)n_maps, (unsigned)chunks);
    if (dot_mode == PPNET_DOT_UNFUSED)
        segcheck_kernel<T, PPNET_DOT_UNFUSED, SWAP, STEER><<<grid, kSegThreads, 0, st>>>(
            pts, seg_off, segs_per_map, obs, obs_cnt, omax, clearance, (T)bound, verdict, steer);
    else
        segcheck_kernel<T, PPNET_DOT_FUSED_SKX, SWAP, STEER><<<grid, kSegThreads, 0, st>>>(
            pts, seg_off, segs_per_map, obs, obs_cnt, omax, clearance, (T)bound, verdict, steer);
    PPNET_LAUNCH_CHECK("segcheck_kernel");
    return PPNET_OK;
}

// ------------------------------------------------------------------------------------------------
// feasibility_check / lvc: one warp per path, circles of the path's map staged per warp in smem.
// ------------------------------------------------------------------------------------------------
constexpr int kPathWarps = 4;
constexpr int kPathCirc = 64;        // circles staged per pass per warp

__device__ __forceinline__ bool steer_blocked(const float* a, const float* b, const Circle<float>* sc,
                                              int nt, float bound, bool first_tile, bool& dist_pos,
                                              SegState<float>& g) {
    // evaluates one tile of circles; caller ORs tiles
    if (first_tile) {
        g = seg_setup<float, 0, false>(a[0], a[1], b[0], b[1], bound);
        const float x = __fsub_rn(a[0], b[0]), y = __fsub_rn(a[1], b[1]);
        dist_pos = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))) > 0.0f;
        if (g.oob) return true;
    }
    for (int j = 0; j < nt; ++j)
        if (pair_hit<float, 0>(g, sc[j])) return true;
    return false;
}

// blocked(i, j) for waypoint pair (a, b) against all circles of the map: lanes cooperate on staging,
// each lane evaluates ITS OWN pair.  `active` lanes have a pair; returns per-lane steerTo (0/1).
__device__ __forceinline__ int warp_steer(bool active, const float* a, const float* b,
                                          const double* __restrict__ mobs, int cnt, double clearance,
                                          float bound, Circle<float>* sc) {
    const int lane = threadIdx.x & 31;
    bool blocked = false, dist_pos = false;
    SegState<float> g;
    float aa[2] = {0.f, 0.f}, bb[2] = {0.f, 0.f};
    if (active) { aa[0] = a[0]; aa[1] = a[1]; bb[0] = b[0]; bb[1] = b[1]; }
    bool first = true;
    if (cnt == 0 && active) blocked = steer_blocked(aa, bb, sc, 0, bound, true, dist_pos, g);
    for (int t0 = 0; t0 < cnt; t0 += kPathCirc) {
        const int nt = min(kPathCirc, cnt - t0);
        __syncwarp();
        for (int j = lane; j < nt; j += 32) sc[j] = make_circle<float>(mobs + 3 * (t0 + j), clearance);
        __syncwarp();
        if (active && !blocked) blocked = steer_blocked(aa, bb, sc, nt, bound, first, dist_pos, g);
        first = false;
    }
    return (active && dist_pos && blocked) ? 0 : 1;
}

__global__ void __launch_bounds__(kPathWarps * 32)
feasible_kernel(const float* __restrict__ wp, const int64_t* __restrict__ path_off,
                const int32_t* __restrict__ path_map, int64_t n_paths, const double* __restrict__ obs,
                const int32_t* __restrict__ obs_cnt, int omax, double clearance, float bound,
                uint8_t* __restrict__ feasible, int32_t* __restrict__ n_checked) {
    __shared__ Circle<float> sc_all[kPathWarps][kPathCirc];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * kPathWarps + warp;
    if (p >= n_paths) return;
    const int m = path_map[p];
    const int cnt = min(obs_cnt[m], omax);
    const double* mobs = obs + (size_t)m * omax * 3;
    const int64_t lo = path_off[p], hi = path_off[p + 1];
    const int64_t n_edges = hi - lo - 1;
    int first_block = -1;                                  // first blocked edge (reference stops there)
    for (int64_t e0 = 0; e0 < n_edges && first_block < 0; e0 += 32) {
        const int64_t e = e0 + lane;
        const bool active = e < n_edges;
        const float* a = wp + 2 * (lo + (active ? e : 0));
        const int st = warp_steer(active, a, a + 2, mobs, cnt, clearance, bound, sc_all[warp]);
        const unsigned bal = __ballot_sync(0xffffffffu, active && st == 0);
        if (bal) first_block = (int)e0 + (__ffs(bal) - 1);
    }
    if (lane == 0) {
        feasible[p] = first_block < 0 ? 1 : 0;
        if (n_checked) n_checked[p] = first_block < 0 ? (int32_t)max((int64_t)0, n_edges) : first_block + 1;
    }
}

// lvc (neuralplanner.py:123-138).  The reference restarts the double loop from i = 0 after every
// contraction; steerTo is a pure function of its two points, so every pair it re-tests gives the
// answer it gave before and the scan can simply continue at i + 1 on the contracted list -- the
// output list is identical (tests compare against the literally recursive oracle).
// For a fixed i the candidates j = len-1 .. i+2 are independent: 32 lanes test 32 of them at once,
// the winner is the LARGEST free j (the first the reference would find scanning from the far end).
__global__ void __launch_bounds__(kPathWarps * 32)
lvc_kernel(const float* __restrict__ wp, const int64_t* __restrict__ path_off,
           const int32_t* __restrict__ path_map, int64_t n_paths, const double* __restrict__ obs,
           const int32_t* __restrict__ obs_cnt, int omax, double clearance, float bound,
           float* __restrict__ out_wp, int32_t* __restrict__ out_len) {
    __shared__ Circle<float> sc_all[kPathWarps][kPathCirc];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * kPathWarps + warp;
    if (p >= n_paths) return;
    const int m = path_map[p];
    const int cnt = min(obs_cnt[m], omax);
    const double* mobs = obs + (size_t)m * omax * 3;
    const int64_t lo = path_off[p];
    int len = (int)(path_off[p + 1] - lo);
    float* cur = out_wp + 2 * lo;                          // contracted in place in the output
    for (int k = lane; k < 2 * len; k += 32) cur[k] = wp[2 * lo + k];
    __syncwarp();
    for (int i = 0; i < len - 1; ++i) {
        int found = -1;
        for (int jt = len - 1; jt > i + 1 && found < 0; jt -= 32) {
            const int j = jt - lane;
            const bool active = j > i + 1;
            const int st = warp_steer(active, cur + 2 * i, cur + 2 * (active ? j : i), mobs, cnt,
                                      clearance, bound, sc_all[warp]);
            const unsigned bal = __ballot_sync(0xffffffffu, active && st == 1);
            if (bal) found = jt - (__ffs(bal) - 1);        // lowest lane = largest j
        }
        if (found >= 0) {                                  // drop waypoints i+1 .. found-1
            const int shift = found - (i + 1);
            for (int k0 = 2 * (i + 1); k0 < 2 * (len - shift); k0 += 32) {
                const int k = k0 + lane;
                float v = 0.f;
                if (k < 2 * (len - shift)) v = cur[k + 2 * shift];
                __syncwarp();
                if (k < 2 * (len - shift)) cur[k] = v;
                __syncwarp();
            }
            len -= shift;
        }
    }
    if (lane == 0) out_len[p] = len;
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_segcheck_edage_f64(const double* pts_rc, int64_t n_segs, const int64_t* seg_off,
                                        int64_t segs_per_map, int64_t n_maps, const double* obs,
                                        const int32_t* obs_cnt, int32_t omax, double clearance,
                                        double bound, int32_t dot_mode, uint8_t* verdict, void* stream) {
    PPNET_REQUIRE(dot_mode == PPNET_DOT_FUSED_SKX || dot_mode == PPNET_DOT_UNFUSED, "bad dot_mode");
    PPNET_REQUIRE(verdict || n_segs == 0, "segcheck: verdict is null");
    return launch<double, true, false>(pts_rc, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax,
                                       clearance, bound, dot_mode, verdict, nullptr, (cudaStream_t)stream);
}

extern "C" int ppnet_segcheck_mpnet_f32(const float* pts_xy, int64_t n_segs, const int64_t* seg_off,
                                        int64_t segs_per_map, int64_t n_maps, const double* obs,
                                        const int32_t* obs_cnt, int32_t omax, double clearance,
                                        double bound, uint8_t* verdict, uint8_t* steer, void* stream) {
    PPNET_REQUIRE(verdict || steer || n_segs == 0, "segcheck: both outputs are null");
    return launch<float, false, true>(pts_xy, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax,
                                      clearance, bound, 0, verdict, steer, (cudaStream_t)stream);
}

extern "C" int ppnet_path_feasible_f32(const float* wp, const int64_t* path_off, const int32_t* path_map,
                                       int64_t n_paths, const double* obs, const int32_t* obs_cnt,
                                       int32_t omax, double clearance, double bound, uint8_t* feasible,
                                       int32_t* n_checked, void* stream) {
    PPNET_REQUIRE(n_paths >= 0, "feasible: negative n_paths");
    if (n_paths == 0) return PPNET_OK;
    PPNET_REQUIRE(wp && path_off && path_map && obs_cnt && feasible, "feasible: null pointer");
    PPNET_REQUIRE(omax == 0 || obs, "feasible: obs is null");
    const unsigned grid = (unsigned)((n_paths + kPathWarps - 1) / kPathWarps);
    feasible_kernel<<<grid, kPathWarps * 32, 0, (cudaStream_t)stream>>>(
        wp, path_off, path_map, n_paths, obs, obs_cnt, omax, clearance, (float)bound, feasible, n_checked);
    PPNET_LAUNCH_CHECK("feasible_kernel");
    return PPNET_OK;
}

extern "C" int ppnet_lvc_f32(const float* wp, const int64_t* path_off, const int32_t* path_map,
                             int64_t n_paths, const double* obs, const int32_t* obs_cnt, int32_t omax,
                             double clearance, double bound, float* out_wp, int32_t* out_len,
                             void* stream) {
    PPNET_REQUIRE(n_paths >= 0, "lvc: negative n_paths");
    if (n_paths == 0) return PPNET_OK;
    PPNET_REQUIRE(wp && path_off && path_map && obs_cnt && out_wp && out_len, "lvc: null pointer");
    PPNET_REQUIRE(omax == 0 || obs, "lvc: obs is null");
    const unsigned grid = (unsigned)((n_paths + kPathWarps - 1) / kPathWarps);
    lvc_kernel<<<grid, kPathWarps * 32, 0, (cudaStream_t)stream>>>(
        wp, path_off, path_map, n_paths, obs, obs_cnt, omax, clearance, (float)bound, out_wp, out_len);
    PPNET_LAUNCH_CHECK("lvc_kernel");
    return PPNET_OK;
}
