// Shared device helpers of the segment-vs-circles verdict kernels (segcheck.cu: one flavour per launch;
// verdict.cu: A11 + A12 fused on one read of the segments).  See segcheck.cu for the derivations.
#pragma once
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

constexpr int kCircTile = 128;       // circles staged per pass (one bit each in a 128-bit candidate mask)
constexpr int kSegThreads = 128;

// ---- the fast path: exact-safe culling + division-free decisions ----------------------------------------
// Notation: u = unit roundoff of the flavour, Mg = |s|_1 + |e|_1 + |o|_1 + thr + 1 (bounds every length in
// the pair), E = 50 u Mg = the reach of the reference's accumulated rounding (its `dis` is within 7u Mg of
// the real-arithmetic value of the same expression on the same rounded d = e-s, q = o-s; the sign of its
// normalised (p-s).(p-e) is certain once the foot is >= 4x the error of p away from both ends), and
// delta = eps Mg with eps >= 10x E/Mg plus the error of the approximate L used below.
//   far:      |cross(q,d)| > (thr + delta) L                  => |dis| > thr + E       => reference says no
//   outside:  q.d < -delta L  or  q.d > |d|^2 + delta L       => foot beyond an end    => reference says no
//   inside:   |cross| < (thr - delta) L and delta L < q.d < |d|^2 - delta L           => reference says yes
// Anything else (a fraction ~eps of the near pairs) replays the reference's operation sequence verbatim, and
// so does every pair with a non-finite / huge / degenerate operand.  The filter arithmetic may round any
// way it likes (its own error is part of delta).
// Culling: circle boxes inflated by thr + eps*mc and segment-piece boxes inflated by eps*ms are binned
// separably (32 x-bins, 32 y-bins, one 128-bit circle mask per bin); box intersection on a grid is
// separable, so  cand = OR_x(xmask) & OR_y(ymask)  is exactly the set of circles whose box meets the
// piece's box.  A culled circle is farther than thr + delta (in the max norm, hence Euclidean) from every
// point of the segment: the vertex test fails, and either the foot is inside (|dis| >= thr + delta), or
// outside by >= E (sign certain), or within E of an end (|dis| >= thr + delta - E/2): the reference says no.
template <typename T> struct Filt;
template <> struct Filt<double> {
    static constexpr double eps = 4e-6;      // L comes from a float sqrt (rel. 1.2e-7): 30x over it
    static constexpr double lim = 1e15, tiny = 1e-30;
};
template <> struct Filt<float> {
    static constexpr float eps = 5e-5f;      // E/Mg = 50 * 2^-24 = 3e-6
    static constexpr float lim = 1e8f, tiny = 1e-20f;
};
constexpr int kBins = 32;

struct Mask128 {
    uint32_t w[4];
};
__device__ __forceinline__ int bin_clamp(float b) {                     // monotone non-decreasing in b
    return (int)fminf(fmaxf(b, 0.0f), (float)(kBins - 1));
}

// Per-segment state of the fast path (no division, no exact square root)
template <typename T>
struct FastSeg {
    T s0, s1, e0, e1;   // (x, y) after the flavour's swap
    T d0, d1, L2, L;    // d = e - s (rounded once, as the reference does), |d|^2, approximate |d|
    T es;               // eps * (|s|_1 + |e|_1 + 1)
    bool oob, verbatim;
};

template <typename T, bool SWAP>
__device__ __forceinline__ FastSeg<T> fast_setup(T a0, T a1, T b0, T b1, T bound) {
    using F = FP<T>;
    FastSeg<T> g;
    g.oob = (a0 < T(0)) || (a1 > bound) || (b0 < T(0)) || (b1 > bound);
    if (SWAP) { g.s0 = a1; g.s1 = a0; g.e0 = b1; g.e1 = b0; }
    else      { g.s0 = a0; g.s1 = a1; g.e0 = b0; g.e1 = b1; }
    g.d0 = F::sub(g.e0, g.s0);
    g.d1 = F::sub(g.e1, g.s1);
    g.L2 = g.d0 * g.d0 + g.d1 * g.d1;
    g.L = (T)sqrtf((float)g.L2);
    const T ms = F::abs_(g.s0) + F::abs_(g.s1) + F::abs_(g.e0) + F::abs_(g.e1) + T(1);
    g.es = Filt<T>::eps * ms;
    g.verbatim = !(ms < Filt<T>::lim) || !(g.L2 > Filt<T>::tiny);      // NaN / inf / huge / degenerate
    return g;
}

// The reference's edge test, operation for operation (process_map.py:401-417 / neuralplanner.py:55-68)
template <typename T, int MODE>
__device__ __noinline__ bool edge_exact(T s0, T s1, T e0, T e1, T ox, T oy, T thr) {
    using F = FP<T>;
    const T d0 = F::sub(e0, s0), d1 = F::sub(e1, s1);
    const T L = F::sqrt_(Dot<T, MODE>::f(d0, d1, d0, d1));
    const T n0 = F::div(d1, L), n1 = F::div(-d0, L);
    const T q0 = F::sub(ox, s0), q1 = F::sub(oy, s1);
    const T dis = Dot<T, MODE>::f(n0, n1, q0, q1);
    if (!(F::abs_(dis) < thr)) return false;
    const T p0 = F::sub(ox, F::mul(dis, n0)), p1 = F::sub(oy, F::mul(dis, n1));
    T u0 = F::sub(p0, s0), u1 = F::sub(p1, s1);
    const T nu = F::sqrt_(Dot<T, MODE>::f(u0, u1, u0, u1));
    u0 = F::div(u0, nu); u1 = F::div(u1, nu);
    T w0 = F::sub(p0, e0), w1 = F::sub(p1, e1);
    const T nw = F::sqrt_(Dot<T, MODE>::f(w0, w1, w0, w1));
    w0 = F::div(w0, nw); w1 = F::div(w1, nw);
    return Dot<T, MODE>::f(u0, u1, w0, w1) < T(0);
}

template <typename T, int MODE>
__device__ __forceinline__ bool fast_pair(const FastSeg<T>& g, const Circle<T>& c, T em, bool exact_only) {
    using F = FP<T>;
    // vertex test on e only, exact:  euclidean(e, o) < thr  <=>  rn(rn(v0^2) + rn(v1^2)) < T2
    const T v0 = F::sub(g.e0, c.ox), v1 = F::sub(g.e1, c.oy);
    if (F::add(F::mul(v0, v0), F::mul(v1, v1)) < c.T2) return true;
    if (!exact_only) {
        const T q0 = F::sub(c.ox, g.s0), q1 = F::sub(c.oy, g.s1);
        const T ac = F::abs_(q0 * g.d1 - q1 * g.d0);
        const T del = g.es + em;
        if (ac > (c.thr + del) * g.L) return false;                    // far
        const T tt = q0 * g.d0 + q1 * g.d1;
        const T dl = del * g.L;
        if (tt < -dl || tt > g.L2 + dl) return false;                  // foot beyond an end
        if (ac < (c.thr - del) * g.L && tt > dl && tt < g.L2 - dl) return true;   // inside
    }
    return edge_exact<T, MODE>(g.s0, g.s1, g.e0, g.e1, c.ox, c.oy, c.thr);
}

// The same decisions with every test computed up front and combined with predicates: in a 32-pair round some lane
// nearly always reaches the last test, so the early exits of `fast_pair` save no warp instruction, while each of them is a
// divergence point (BSSY / BSYNC + a branch to resolve) in front of dependent arithmetic.  Only the verbatim replay stays a
// branch.  NaN / inf operands make every filter comparison false, as above.
template <typename T, int MODE>
__device__ __forceinline__ bool fast_pair_bf(T s0, T s1, T e0, T e1, T L, T es, T ox, T oy, T thr, T T2, T em) {
    using F = FP<T>;
    const T v0 = F::sub(e0, ox), v1 = F::sub(e1, oy);
    const bool vert = F::add(F::mul(v0, v0), F::mul(v1, v1)) < T2;     // exact vertex test
    const T d0 = F::sub(e0, s0), d1 = F::sub(e1, s1);
    const T L2 = d0 * d0 + d1 * d1;
    const T q0 = F::sub(ox, s0), q1 = F::sub(oy, s1);
    const T ac = F::abs_(q0 * d1 - q1 * d0);
    const T tt = q0 * d0 + q1 * d1;
    const T del = es + em, dl = del * L;
    const bool far = ac > (thr + del) * L;
    const bool out = tt < -dl || tt > L2 + dl;
    const bool in = ac < (thr - del) * L && tt > dl && tt < L2 - dl;
    if (vert || far || out || in) return vert || (!far && !out && in);
    return edge_exact<T, MODE>(s0, s1, e0, e1, ox, oy, thr);
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

}  // namespace ppnet
