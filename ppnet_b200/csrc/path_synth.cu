// Target-path synthesis: what PathGroup.generate does for every accepted Path, batched over paths.
//
//   A1  PathSeg.random / translation / gradient / length      EDaGe-PP/PathSeg.py:10-58
//   A2  Path.generate + transform (C1 chaining) + plot         EDaGe-PP/Path.py:78-98, 224-316
//   A3  Path.draw_boundary                                      EDaGe-PP/Path.py:318-356
//   A5  the ray set Path.path_space paints                      EDaGe-PP/Path.py:113-134  (paint itself: grid.cu)
//   A7  Path.space_normalization, point part                    EDaGe-PP/Path.py:157-193
//   A8  Path.search_isle                                        EDaGe-PP/Path.py:502-537
//   A9  Path.set_obstacles                                      EDaGe-PP/Path.py:463-500
// (A4 grid index and A6 hull are the kernels of grid.cu / hull.cu, called from the launcher below.)
//
// The reference draws from the global MT19937 streams; here every draw is Philox keyed by the global path id
// (STREAM_PATH for A1, STREAM_PATH_OBST for A9), so a bank is reproducible for any sharding.  In parity mode the
// caller supplies the reference's own draws and every output is comparable to the reference (1e-5 relative:
// np.polyfit is LAPACK least squares, not bit-stable; integer outputs are exact).
//
// The degree-d least-squares fit over the FIXED abscissae x_k = k/100 (k < 1000) is done in the discrete
// orthogonal (Gram) basis of those abscissae: phi_{j+1} = (t - alpha_j) phi_j - beta_j phi_{j-1}, t = (x-a)/b.
// The recurrence constants and the phi_j -> monomial conversion are computed once on the host in long double
// (cond ~1: no normal-equation squaring) and travel as kernel parameters; a warp accumulates the d+1 inner
// products <y, phi_j> of one curve piece.
#include <math_constants.h>

#include <mutex>

#include "common.cuh"

namespace ppnet {

constexpr int kMaxOrder = 7;
constexpr int kFitN = 1000;          // PathSeg.py:22
constexpr int kSegSamples = 100;     // Path.plot / PathSeg.length
constexpr int kBndSamples = 50;      // Path.draw_boundary
constexpr double kSegLenRange = 7.0, kMinLen = 0.0;   // PathSeg.py:5-6

struct FitBasis {
    double a, b;                                   // t = (x - a) / b
    double alpha[kMaxOrder + 1], beta[kMaxOrder + 1], inv_norm[kMaxOrder + 1];
    double G[kMaxOrder + 1][kMaxOrder + 1];        // G[j][i] = coefficient of x^i in phi_j(t(x))
};

static FitBasis make_basis(int order) {
    FitBasis B{};
    typedef long double ld;
    const int N = kFitN;
    const ld a = (ld)4.995L, b = (ld)4.995L;
    B.a = (double)a; B.b = (double)b;
    static ld t[kFitN], p0[kFitN], p1[kFitN], p2[kFitN];
    for (int k = 0; k < N; ++k) { t[k] = ((ld)((double)k / 100.0) - a) / b; p0[k] = 0; p1[k] = 1; }
    ld G[kMaxOrder + 2][kMaxOrder + 2] = {};       // monomial coefficients (in x) of phi_j
    G[0][0] = 1;
    ld norm_prev = 1;
    for (int j = 0; j <= order; ++j) {
        ld nn = 0, ta = 0;
        for (int k = 0; k < N; ++k) { nn += p1[k] * p1[k]; ta += t[k] * p1[k] * p1[k]; }
        const ld alpha = ta / nn, beta = j == 0 ? 0 : nn / norm_prev;
        B.alpha[j] = (double)alpha; B.beta[j] = (double)beta; B.inv_norm[j] = (double)(1 / nn);
        // phi_{j+1} = (t - alpha) phi_j - beta phi_{j-1},  t = x/b - a/b
        for (int i = 0; i <= j + 1; ++i) {
            ld c = 0;
            if (i >= 1) c += G[j][i - 1] / b;
            if (i <= j) c += (-a / b - alpha) * G[j][i];
            if (j >= 1 && i <= j - 1) c -= beta * G[j - 1][i];
            G[j + 1][i] = c;
        }
        for (int k = 0; k < N; ++k) { p2[k] = (t[k] - alpha) * p1[k] - beta * p0[k]; p0[k] = p1[k]; p1[k] = p2[k]; }
        norm_prev = nn;
    }
    for (int j = 0; j <= order; ++j)
        for (int i = 0; i <= order; ++i) B.G[j][i] = (double)G[j][i];
    return B;
}

// np.polyval: Horner from the highest power, un-fused (coefficients highest power first)
__device__ __forceinline__ double polyval(const double* p, int n, double x) {
    double y = 0.0;
    for (int i = 0; i < n; ++i) y = __dadd_rn(__dmul_rn(y, x), p[i]);
    return y;
}
// Path.coord_rotation on one column: OpenBLAS dgemm accumulates with FMA on AVX2/AVX-512 hosts
__device__ __forceinline__ void rot2p(double c, double s, double x0, double x1, double& r0, double& r1) {
    r0 = __fma_rn(-s, x1, __dmul_rn(c, x0));
    r1 = __fma_rn(c, x1, __dmul_rn(s, x0));
}
__device__ __forceinline__ double dotp(double a0, double a1, double b0, double b1) {     // np.dot, SkylakeX ddot
    return __fma_rn(a1, b1, __dmul_rn(a0, b0));
}
__device__ __forceinline__ int cell_of(double v, double step, double off) {              // A4
    return (int)rint(__dadd_rn(__ddiv_rn(v, step), off));
}
__device__ __forceinline__ double u53_at(uint2 key, uint32_t stream, uint64_t unit, uint32_t block, int half) {
    const uint4 r = Philox::gen(key, make_uint4(block, stream, (uint32_t)unit, (uint32_t)(unit >> 32)));
    return half ? u53(r.z, r.w) : u53(r.x, r.y);
}

// ---------------------------------------------------------------------------------------------------------
// A1: one warp per curve piece
// draw layout of path g (STREAM_PATH): block 0 (x,y) -> the PathGroup "forced straight" draw;
//   piece i: block 1 + 502 i: (x,y) -> is_straight draw, (z,w) -> EndPoint draw; blocks 2 + 502 i + j, j < 500:
//   y[2j] = (x,y), y[2j+1] = (z,w).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pathseg_kernel(ppnet_path_params P, FitBasis B) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t sid = (int64_t)blockIdx.x * 4 + warp;
    const int S = P.seg_num, D = P.poly_order, NC = D + 1;
    if (sid >= P.n_paths * S) return;
    const int64_t p = sid / S;
    const int i = (int)(sid % S);
    const uint64_t g = (uint64_t)(P.path0 + p);
    const uint2 key = make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32));
    const uint32_t blk0 = 1u + 502u * (uint32_t)i;

    // is_straight = forced or random(1) < 0.2   (PathSeg.py:19; PathGenerate.py:36 draws the forced flag)
    bool straight;
    if (P.in_straight) straight = P.in_straight[sid] != 0;
    else {
        const bool forced = P.force_straight ? P.force_straight[p] != 0
                                             : !(u53_at(key, STREAM_PATH, g, 0u, 0) > 0.01);
        straight = forced || u53_at(key, STREAM_PATH, g, blk0, 0) < 0.2;
    }
    double poly[kMaxOrder + 1];
    double endp;
    if (P.in_poly) {                                       // PathSeg.random(poly, endpoint)
        for (int d = 0; d < NC; ++d) poly[d] = P.in_poly[sid * NC + d];
        endp = P.in_uend[sid];
    } else {
        double m[kMaxOrder + 1];
        for (int d = 0; d < NC; ++d) m[d] = 0.0;
        for (int k = lane; k < kFitN; k += 32) {
            double u;
            if (P.in_y) u = P.in_y[sid * kFitN + k];
            else u = u53_at(key, STREAM_PATH, g, blk0 + 1u + (uint32_t)(k >> 1), k & 1);
            const double y = __dsub_rn(__dmul_rn(u, 10.0), 5.0);                          // random(1000)*10 - 5
            const double t = ((double)k / 100.0 - B.a) / B.b;
            double pm = 0.0, pc = 1.0;
            for (int d = 0; d < NC; ++d) {
                m[d] = fma(y, pc, m[d]);
                const double pn = (t - B.alpha[d]) * pc - B.beta[d] * pm;
                pm = pc; pc = pn;
            }
        }
        for (int d = 0; d < NC; ++d)
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) m[d] += __shfl_xor_sync(0xffffffffu, m[d], sft);
        // monomial coefficients, highest power first (np.polyfit order)
        for (int d = 0; d < NC; ++d) {
            double c = 0.0;
            for (int j = 0; j < NC; ++j) c = fma(m[j] * B.inv_norm[j], B.G[j][D - d], c);
            poly[d] = c;
        }
        const double u_end = P.in_uend ? P.in_uend[sid] : u53_at(key, STREAM_PATH, g, blk0, 1);
        endp = __dadd_rn(__dmul_rn(u_end, kSegLenRange - kMinLen), kMinLen);
    }
    poly[D] = 0.0;                                         // :28
    if (straight) for (int d = 0; d < NC - 2; ++d) poly[d] = 0.0;                         // :29-31
    double pder[kMaxOrder + 1];
    for (int d = 0; d < D; ++d) pder[d] = __dmul_rn(poly[d], (double)(D - d));           // np.polyder
    const double y_end = polyval(poly, NC, endp);
    // length(): 99 chords between the 100 samples + the chord to the end point (:49-58)
    double len = 0.0;
    for (int k = lane; k < kSegSamples; k += 32) {
        const double x0 = __dmul_rn((double)k / 100.0, endp), y0 = polyval(poly, NC, x0);
        double x1, y1;
        if (k + 1 < kSegSamples) { x1 = __dmul_rn((double)(k + 1) / 100.0, endp); y1 = polyval(poly, NC, x1); }
        else { x1 = endp; y1 = y_end; }
        const double dx = __dsub_rn(x1, x0), dy = __dsub_rn(y1, y0);
        len += __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) len += __shfl_xor_sync(0xffffffffu, len, sft);
    if (lane == 0) {
        if (i == 0) {                                      // Path.is_straight: only the PathGroup draw forces a whole path
            bool forced = false;
            if (P.force_straight) forced = P.force_straight[p] != 0;
            else if (!P.in_straight) forced = !(u53_at(key, STREAM_PATH, g, 0u, 0) > 0.01);
            P.path_straight[p] = forced ? 1 : 0;
        }
        for (int d = 0; d < NC; ++d) P.poly[sid * NC + d] = poly[d];
        P.endpoint[sid] = endp;
        P.is_straight[sid] = straight ? 1 : 0;
        P.seg_trans_local[sid * 2] = endp;
        P.seg_trans_local[sid * 2 + 1] = y_end;
        P.grad_st[sid] = polyval(pder, D, 0.0);
        P.grad_end[sid] = polyval(pder, D, endp);
        P.seg_length[sid] = len;
    }
}

// ---------------------------------------------------------------------------------------------------------
// A2 + A3 + ray set + A4 cells: one CTA per path
// ---------------------------------------------------------------------------------------------------------
constexpr int kChainThreads = 128;
constexpr int kMaxSeg = 64;

__global__ void __launch_bounds__(kChainThreads)
chain_kernel(ppnet_path_params P) {
    __shared__ double rot[kMaxSeg], cs[kMaxSeg], sn[kMaxSeg], tr[kMaxSeg][2];
    __shared__ double red[kChainThreads / 32];
    const int64_t p = blockIdx.x;
    const int S = P.seg_num, D = P.poly_order, NC = D + 1, Np = kSegSamples * S, Nb = 2 * kBndSamples * S + 2 * kBndSamples;
    const double* poly = P.poly + p * S * NC;
    const double* endp = P.endpoint + p * S;
    const double* tl = P.seg_trans_local + p * S * 2;
    if (threadIdx.x == 0) {
        double acc = 0.0;                                  // angle_abs: running sum of atan(GradEnd_{i-1}) - atan(GradSt_i)
        rot[0] = 0.0; cs[0] = 1.0; sn[0] = 0.0;
        for (int i = 1; i < S; ++i) {
            acc = __dadd_rn(acc, __dsub_rn(atan(P.grad_end[p * S + i - 1]), atan(P.grad_st[p * S + i])));
            rot[i] = acc; cs[i] = cos(acc); sn[i] = sin(acc);
        }
        double t0 = 0.0, t1 = 0.0;                         // translation_seg: sum_{k<i} R(rot_k) T_k  (k = 0 un-rotated)
        for (int i = 0; i < S; ++i) {
            tr[i][0] = t0; tr[i][1] = t1;
            double a0 = tl[2 * i], a1 = tl[2 * i + 1];
            if (i != 0) rot2p(cs[i], sn[i], tl[2 * i], tl[2 * i + 1], a0, a1);
            t0 = __dadd_rn(t0, a0); t1 = __dadd_rn(t1, a1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += kChainThreads) {
        P.seg_rot[p * S + i] = rot[i];
        P.seg_trans[(p * S + i) * 2] = tr[i][0];
        P.seg_trans[(p * S + i) * 2 + 1] = tr[i][1];
    }
    auto transform = [&](int i, double x, double y, double& ox, double& oy) {          // Path.point_transform
        if (i != 0) rot2p(cs[i], sn[i], x, y, ox, oy); else { ox = x; oy = y; }
        ox = __dadd_rn(ox, tr[i][0]); oy = __dadd_rn(oy, tr[i][1]);
    };
    // SegPoint (:86-91)
    double* sp = P.segpoint_raw + p * (S + 1) * 2;
    for (int i = threadIdx.x; i <= S; i += kChainThreads) {
        double x = 0.0, y = 0.0;
        if (i > 0) transform(i - 1, endp[i - 1], polyval(poly + (i - 1) * NC, NC, endp[i - 1]), x, y);
        sp[2 * i] = x; sp[2 * i + 1] = y;
    }
    double Ex, Ey;
    transform(S - 1, endp[S - 1], polyval(poly + (S - 1) * NC, NC, endp[S - 1]), Ex, Ey);
    // PathPoint (plot(), :256-260) + A4 cells at offset R (convexhull(), :390) + Length (:93-94)
    const double stepA4 = __ddiv_rn(P.map_size, P.resolution);                           // MapSize / Resolution
    double* pp = P.pathpoint_raw + p * Np * 2;
    int32_t* cells = P.cells + p * Np * 2;
    auto path_pt = [&](int idx, double& x, double& y) {
        const int i = idx / kSegSamples, k = idx % kSegSamples;
        const double xs = __dmul_rn((double)k / 100.0, endp[i]);
        transform(i, xs, polyval(poly + i * NC, NC, xs), x, y);
    };
    double len = 0.0;
    for (int idx = threadIdx.x; idx < Np; idx += kChainThreads) {
        double x, y;
        path_pt(idx, x, y);
        pp[2 * idx] = x; pp[2 * idx + 1] = y;
        cells[2 * idx] = cell_of(x, stepA4, P.resolution);
        cells[2 * idx + 1] = cell_of(y, stepA4, P.resolution);
        if (idx + 1 < Np) {
            double x1, y1;
            path_pt(idx + 1, x1, y1);
            const double dx = __dsub_rn(x, x1), dy = __dsub_rn(y, y1);
            len += __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) len += __shfl_xor_sync(0xffffffffu, len, sft);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = len;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kChainThreads / 32; ++w) t += red[w];
        P.length[p] = t;
    }
    // boundary (:318-333): 50 samples per piece, unit normal R(rot).[y', -1] / |.|
    const double hc = __dmul_rn(0.5, P.clearance);
    const double step_len = __dmul_rn(__ddiv_rn(1.0, P.resolution), P.map_size);         // 1 / Resolution * map_size
    double* up = P.up + p * S * kBndSamples * 2;
    double* upd = P.up_dir + p * S * kBndSamples * 2;
    double* down = P.down + p * S * kBndSamples * 2;
    double* bp = P.boundary_raw + p * Nb * 2;
    double* rx = P.ray_x0 + p * Nb * 2;
    double* rd = P.ray_dir + p * Nb * 2;
    auto up_pt = [&](int i, int k, double& ux, double& uy, double& n0, double& n1, double& dxo, double& dyo) {
        const double xs = __dmul_rn((double)k / 50.0, endp[i]);
        double pder[kMaxOrder + 1];
        for (int d = 0; d < D; ++d) pder[d] = __dmul_rn(poly[i * NC + d], (double)(D - d));
        const double ys = polyval(poly + i * NC, NC, xs), yd = polyval(pder, D, xs);
        double px, py;
        transform(i, xs, ys, px, py);
        rot2p(cs[i], sn[i], yd, -1.0, n0, n1);             // Rotation_0 == False == 0: cos 1, sin 0
        const double nn = __dsqrt_rn(dotp(n0, n1, n0, n1));
        n0 = __ddiv_rn(n0, nn); n1 = __ddiv_rn(n1, nn);
        ux = __dsub_rn(px, __dmul_rn(hc, n0)); uy = __dsub_rn(py, __dmul_rn(hc, n1));
        dxo = __dadd_rn(px, __dmul_rn(hc, n0)); dyo = __dadd_rn(py, __dmul_rn(hc, n1));
    };
    const int nside = S * kBndSamples;
    for (int idx = threadIdx.x; idx < nside; idx += kChainThreads) {
        const int i = idx / kBndSamples, k = idx % kBndSamples;
        double ux, uy, n0, n1, dx, dy;
        up_pt(i, k, ux, uy, n0, n1, dx, dy);
        up[2 * idx] = ux; up[2 * idx + 1] = uy;
        upd[2 * idx] = n0; upd[2 * idx + 1] = n1;
        down[2 * idx] = dx; down[2 * idx + 1] = dy;
        // BoundaryPoint = [init reversed | up | end | down reversed]  (:345-356)
        bp[2 * (kBndSamples + idx)] = ux; bp[2 * (kBndSamples + idx) + 1] = uy;
        const int di = 2 * kBndSamples + nside + (nside - 1 - idx);
        bp[2 * di] = dx; bp[2 * di + 1] = dy;
        // rays (:127-134): up rays follow the 100 cap rays, then the down rays
        const int ru = 2 * kBndSamples + idx, rdn = 2 * kBndSamples + nside + idx;
        rx[2 * ru] = ux; rx[2 * ru + 1] = uy;
        rd[2 * ru] = __dmul_rn(step_len, n0); rd[2 * ru + 1] = __dmul_rn(step_len, n1);
        rx[2 * rdn] = dx; rx[2 * rdn + 1] = dy;
        rd[2 * rdn] = __dmul_rn(step_len, __dmul_rn(-1.0, n0)); rd[2 * rdn + 1] = __dmul_rn(step_len, __dmul_rn(-1.0, n1));
    }
    // caps (:334-343): 50 points each, rotating up[0][0] about the origin / up[S-1][49] about the end point
    for (int i = threadIdx.x; i < 2 * kBndSamples; i += kChainThreads) {
        const bool is_end = i >= kBndSamples;
        const int k = is_end ? i - kBndSamples : i;
        double ux, uy, n0, n1, dx, dy, cx, cy;
        if (!is_end) {
            up_pt(0, 0, ux, uy, n0, n1, dx, dy);
            const double a = __dmul_rn(__ddiv_rn(CUDART_PI, 50.0), (double)(k + 1));
            rot2p(cos(a), sin(a), ux, uy, cx, cy);
            P.cap_init[(p * kBndSamples + k) * 2] = cx; P.cap_init[(p * kBndSamples + k) * 2 + 1] = cy;
            const int bi = kBndSamples - 1 - k;            // init reversed
            bp[2 * bi] = cx; bp[2 * bi + 1] = cy;
            // ray (:120-122): dir = -step_len * p / |p|
            const double nn = __dsqrt_rn(dotp(cx, cy, cx, cy));
            rx[2 * k] = cx; rx[2 * k + 1] = cy;
            rd[2 * k] = __ddiv_rn(__dmul_rn(-step_len, cx), nn); rd[2 * k + 1] = __ddiv_rn(__dmul_rn(-step_len, cy), nn);
        } else {
            up_pt(S - 1, kBndSamples - 1, ux, uy, n0, n1, dx, dy);
            const double a = __dmul_rn(__ddiv_rn(-CUDART_PI, 50.0), (double)(k + 1));
            rot2p(cos(a), sin(a), __dsub_rn(ux, Ex), __dsub_rn(uy, Ey), cx, cy);
            cx = __dadd_rn(cx, Ex); cy = __dadd_rn(cy, Ey);
            P.cap_end[(p * kBndSamples + k) * 2] = cx; P.cap_end[(p * kBndSamples + k) * 2 + 1] = cy;
            const int bi = kBndSamples + nside + k;
            bp[2 * bi] = cx; bp[2 * bi + 1] = cy;
            // ray (:123-126): dir = step_len * (E - p) / |E - p|
            const double v0 = __dsub_rn(Ex, cx), v1 = __dsub_rn(Ey, cy);
            const double nn = __dsqrt_rn(dotp(v0, v1, v0, v1));
            const int ri = kBndSamples + k;
            rx[2 * ri] = cx; rx[2 * ri + 1] = cy;
            rd[2 * ri] = __ddiv_rn(__dmul_rn(step_len, v0), nn); rd[2 * ri + 1] = __ddiv_rn(__dmul_rn(step_len, v1), nn);
        }
    }
    if (threadIdx.x == 0) P.step_num[p] = __ddiv_rn(__dmul_rn(0.8, P.clearance), step_len);   // dis = 0.8 * c / step_len
}

// ---------------------------------------------------------------------------------------------------------
// A7 (point part): one CTA per path
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kChainThreads)
normalize_kernel(ppnet_path_params P) {
    __shared__ double sh[4];                               // cos, sin, shift0, shift1
    const int64_t p = blockIdx.x;
    const int S = P.seg_num, Np = kSegSamples * S, Nb = 2 * kBndSamples * S + 2 * kBndSamples;
    const double R = P.resolution;
    const int H = min(P.hull_cnt[p], P.hmax);
    if (threadIdx.x == 0) {
        const double* e = P.segpoint_raw + (p * (S + 1) + S) * 2;
        // Rotation = atan(Ey / Ex) / pi * 180 + angle(-135)   (:158-159)
        const double rotation = __dadd_rn(__dmul_rn(__ddiv_rn(atan(__ddiv_rn(e[1], e[0])), CUDART_PI), 180.0), -135.0);
        const double rad = __dmul_rn(__ddiv_rn(-rotation, 180.0), CUDART_PI);
        const double c = cos(rad), s = sin(rad);
        double m0 = 0.0, m1 = 0.0;
        for (int i = 0; i < H; ++i) {                      // hull about (R, R), then its vertex mean (:162-168)
            double r0, r1;
            rot2p(c, s, __dsub_rn((double)P.hull_raw[(p * P.hmax + i) * 2], R), __dsub_rn((double)P.hull_raw[(p * P.hmax + i) * 2 + 1], R), r0, r1);
            m0 += __dadd_rn(r0, R); m1 += __dadd_rn(r1, R);
        }
        m0 /= (double)max(H, 1); m1 /= (double)max(H, 1);
        sh[0] = c; sh[1] = s;
        sh[2] = __dsub_rn(__dmul_rn(R, 0.5), m0); sh[3] = __dsub_rn(__dmul_rn(R, 0.5), m1);
        P.rotation[p] = rotation;
        P.translation[2 * p] = sh[3]; P.translation[2 * p + 1] = sh[2];                  // stored swapped (:171)
    }
    __syncthreads();
    const double c = sh[0], s = sh[1], t0 = sh[2], t1 = sh[3];
    for (int i = threadIdx.x; i < P.hmax; i += kChainThreads) {
        double h0 = 0.0, h1 = 0.0;
        if (i < H) {
            rot2p(c, s, __dsub_rn((double)P.hull_raw[(p * P.hmax + i) * 2], R), __dsub_rn((double)P.hull_raw[(p * P.hmax + i) * 2 + 1], R), h0, h1);
            h0 = __dadd_rn(__dadd_rn(h0, R), t0); h1 = __dadd_rn(__dadd_rn(h1, R), t1);
        }
        P.hull[(p * P.hmax + i) * 2] = h0; P.hull[(p * P.hmax + i) * 2 + 1] = h1;
    }
    const double stepA4 = __ddiv_rn(P.map_size, R);
    auto norm = [&](const double* src, double* dst, int n) {                            // rotate -> A4 at offset R -> + shift
        for (int i = threadIdx.x; i < n; i += kChainThreads) {
            double r0, r1;
            rot2p(c, s, src[2 * i], src[2 * i + 1], r0, r1);
            dst[2 * i] = __dadd_rn((double)cell_of(r0, stepA4, R), t0);
            dst[2 * i + 1] = __dadd_rn((double)cell_of(r1, stepA4, R), t1);
        }
    };
    norm(P.segpoint_raw + p * (S + 1) * 2, P.segpoint_img + p * (S + 1) * 2, S + 1);
    norm(P.pathpoint_raw + p * Np * 2, P.pathpoint + p * Np * 2, Np);
    norm(P.boundary_raw + p * Nb * 2, P.boundary + p * Nb * 2, Nb);
}

// ---------------------------------------------------------------------------------------------------------
// A8: one CTA per path, one warp per hull edge
// ---------------------------------------------------------------------------------------------------------
constexpr int kIsleThreads = 256;
constexpr int kMaxHull = 128;

__global__ void __launch_bounds__(kIsleThreads)
isle_kernel(ppnet_path_params P) {
    __shared__ int slot[kMaxHull][2];
    const int64_t p = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Np = kSegSamples * P.seg_num;
    const int H = min(min(P.hull_cnt[p], P.hmax), kMaxHull);
    const double2* pp = reinterpret_cast<const double2*>(P.pathpoint) + p * Np;
    const double2* hull = reinterpret_cast<const double2*>(P.hull) + p * P.hmax;
    const double step_len = __dmul_rn(__ddiv_rn(1.0, P.resolution), P.map_size);
    const double long_edge = __ddiv_rn(5.0, step_len);
    const double thr = (double)(int)rint(__dmul_rn(__ddiv_rn(P.clearance, step_len), P.width_coef));   // int(np.round(c/step*width_coef))
    const bool straight_path = P.path_straight[p] != 0;   // path_obstacles(): a straight Path gets no isles (:149-150)
    for (int i = warp; i < H; i += kIsleThreads / 32) {
        const int j = (i == H - 1) ? 0 : i + 1;
        int lo = -1, hi = -1;
        const double2 a = hull[i], b = hull[j];
        const double e0 = __dsub_rn(b.x, a.x), e1 = __dsub_rn(b.y, a.y);
        if (!straight_path && __dsqrt_rn(dotp(e0, e1, e0, e1)) > long_edge) {
            // nearest path point to each end of the edge: first minimum of the rounded norms
            double d0 = CUDART_INF, d1 = CUDART_INF;
            int i0 = 0x7fffffff, i1 = 0x7fffffff;
            for (int k = lane; k < Np; k += 32) {
                const double2 q = pp[k];
                const double x0 = __dsub_rn(q.x, a.x), y0 = __dsub_rn(q.y, a.y);
                const double x1 = __dsub_rn(q.x, b.x), y1 = __dsub_rn(q.y, b.y);
                const double n0 = __dsqrt_rn(dotp(x0, y0, x0, y0)), n1 = __dsqrt_rn(dotp(x1, y1, x1, y1));
                if (n0 < d0) { d0 = n0; i0 = k; }
                if (n1 < d1) { d1 = n1; i1 = k; }
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                const double od0 = __shfl_xor_sync(0xffffffffu, d0, sft), od1 = __shfl_xor_sync(0xffffffffu, d1, sft);
                const int oi0 = __shfl_xor_sync(0xffffffffu, i0, sft), oi1 = __shfl_xor_sync(0xffffffffu, i1, sft);
                if (od0 < d0 || (od0 == d0 && oi0 < i0)) { d0 = od0; i0 = oi0; }
                if (od1 < d1 || (od1 == d1 && oi1 < i1)) { d1 = od1; i1 = oi1; }
            }
            const int l = min(i0, i1), h = max(i0, i1);
            if (h > l) {
                const double2 b0 = pp[l], bl = pp[h - 1];
                double v0 = __dsub_rn(b0.x, bl.x), v1 = __dsub_rn(b0.y, bl.y);
                const double nv = __dsqrt_rn(dotp(v0, v1, v0, v1));
                v0 = __ddiv_rn(v0, nv); v1 = __ddiv_rn(v1, nv);
                const double m0 = v1, m1 = -v0;            // dir = [dir[1], -dir[0]]
                int first = -1;                            // first point farther than thr from the chord
                for (int k0 = l; k0 < h && first < 0; k0 += 32) {
                    const int k = k0 + lane;
                    bool over = false;
                    if (k < h) {
                        const double2 q = pp[k];
                        over = fabs(dotp(__dsub_rn(q.x, b0.x), __dsub_rn(q.y, b0.y), m0, m1)) > thr;   // NaN: never
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, over);
                    if (bal) first = k0 + __ffs(bal) - 1;
                }
                if (first >= 0) {
                    const double2 q = pp[first];
                    if (q.x != bl.x || q.y != bl.y) { lo = l; hi = h; }                 // (p != boundary[-1]).any()
                }
            }
        }
        if (lane == 0) { slot[i][0] = lo; slot[i][1] = hi; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                // the reference appends in hull-edge order
        int n = 0;
        for (int i = 0; i < H; ++i)
            if (slot[i][0] >= 0) {
                if (n < P.hmax) { P.isle[(p * P.hmax + n) * 2] = slot[i][0]; P.isle[(p * P.hmax + n) * 2 + 1] = slot[i][1]; }
                ++n;
            }
        P.isle_cnt[p] = n;
    }
}

// ---------------------------------------------------------------------------------------------------------
// A9: one warp per path.  The scalar bookkeeping is float32 where the reference computes on float32 tensors
// (radius, motion, their running sum -- torch.rand(1) products), float64 elsewhere.
// draws: torch.rand(1) -> u24; draw t of path g = word (t & 3) of Philox block (t >> 2), STREAM_PATH_OBST.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
obstacles_kernel(ppnet_path_params P) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * 4 + warp;
    if (p >= P.n_paths) return;
    const int Np = kSegSamples * P.seg_num;
    const double2* pp = reinterpret_cast<const double2*>(P.pathpoint) + p * Np;
    const uint64_t g = (uint64_t)(P.path0 + p);
    const uint2 key = make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32));
    const double c_px = __dmul_rn(__ddiv_rn(P.clearance, P.map_size), P.resolution);   // c / MapSize * Resolution
    const float c_px_f = (float)c_px;
    const float size_clearance = (float)__dmul_rn(c_px, 1.1);
    const int n_isle = min(P.isle_cnt[p], P.hmax);
    const int n_in = P.in_obst_rand ? P.in_obst_rand_cnt[p] : 0x7fffffff;
    int used = 0, n_obs = 0, status = 0;
    auto draw = [&]() -> float {
        float v;
        if (P.in_obst_rand) v = used < n_in ? P.in_obst_rand[p * P.max_obst_rand + used] : 0.5f;
        else {
            const uint4 r = Philox::gen(key, make_uint4((uint32_t)(used >> 2), STREAM_PATH_OBST, (uint32_t)g, (uint32_t)(g >> 32)));
            const uint32_t w = (used & 3) == 0 ? r.x : (used & 3) == 1 ? r.y : (used & 3) == 2 ? r.z : r.w;
            v = u24(w);
        }
        if (used >= n_in) status |= 2;                      // ran out of supplied draws
        ++used;
        return v;
    };
    for (int t = 0; t < n_isle; ++t) {
        const int lo = P.isle[(p * P.hmax + t) * 2], hi = P.isle[(p * P.hmax + t) * 2 + 1];
        const int n = hi - lo;
        const double2 b0 = pp[lo], bl = pp[hi - 1];
        const double ce0 = __ddiv_rn(__dadd_rn(b0.x, bl.x), 2.0), ce1 = __ddiv_rn(__dadd_rn(b0.y, bl.y), 2.0);
        double t0 = __dsub_rn(b0.x, bl.x), t1 = __dsub_rn(b0.y, bl.y);
        const double nt = __dsqrt_rn(dotp(t0, t1, t0, t1));
        t0 = __ddiv_rn(t0, nt); t1 = __ddiv_rn(t1, nt);                                  // dir_tangent
        double n0 = t1, n1 = -t0;                                                        // dir_normal
        const double2 mid = pp[lo + n / 2];
        if (!(dotp(__dsub_rn(mid.x, ce0), __dsub_rn(mid.y, ce1), n0, n1) < 0.0)) { n0 = -n0; n1 = -n1; }
        // size_max = 2 * max |(p - isle[0]) . n|, peak = first argmax
        double dmax = -1.0;
        int imax = 0x7fffffff;
        for (int k = lane; k < n; k += 32) {
            const double2 q = pp[lo + k];
            const double d = fabs(dotp(__dsub_rn(q.x, b0.x), __dsub_rn(q.y, b0.y), n0, n1));
            if (d > dmax) { dmax = d; imax = k; }
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, dmax, sft);
            const int oi = __shfl_xor_sync(0xffffffffu, imax, sft);
            if (od > dmax || (od == dmax && oi < imax)) { dmax = od; imax = oi; }
        }
        const double size_max = __dmul_rn(dmax, 2.0);
        const float size_max_f = (float)size_max;
        const double2 peak = pp[lo + (imax == 0x7fffffff ? 0 : imax)];
        float sum = 0.0f, size_pre = 0.0f;
        int n_here = 0;
        double co0 = 0.0, co1 = 0.0;
        for (int it = 0; (n_here == 0 ? 0.0 : (double)sum) < size_max; ++it) {
            if (it >= P.max_obst_iter) { status |= 1; break; }                          // the reference would spin on
            const bool first = n_here == 0;
            const float radius = __fdiv_rn(__fmul_rn(draw(), size_max_f), 2.0f);
            const float rn = first ? 1.0f : draw();
            float motion = __fmul_rn(rn, __fadd_rn(__fadd_rn(radius, size_pre), first ? size_clearance : 0.0f));
            const float alt = __fsub_rn(radius, first ? 0.0f : sum);
            if (alt > motion) motion = alt;                                              // max(motion, radius - sum(obs_size))
            const double bx = first ? peak.x : co0, by = first ? peak.y : co1;
            co0 = __dadd_rn(bx, __dmul_rn((double)motion, n0));
            co1 = __dadd_rn(by, __dmul_rn((double)motion, n1));
            if (!first) {
                const float k = __fdiv_rn(__fmul_rn(__fdiv_rn(__fsub_rn(draw(), 0.5f), 0.5f), radius), 2.0f);
                co0 = __dadd_rn(co0, __dmul_rn((double)k, t0));
                co1 = __dadd_rn(co1, __dmul_rn((double)k, t1));
            }
            double m2 = CUDART_INF;                         // clearance clamp against the odd-indexed path points (:486-491)
            for (int k = 1 + 2 * lane; k < Np; k += 64) {
                const double2 q = pp[k];
                const double dx = __dsub_rn(q.x, co0), dy = __dsub_rn(q.y, co1);
                m2 = fmin(m2, __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) m2 = fmin(m2, __shfl_xor_sync(0xffffffffu, m2, sft));
            const double mind = __dsqrt_rn(m2);
            double r_out = (double)radius;
            bool clamped = false;
            if (mind < (double)__fadd_rn(radius, c_px_f)) { r_out = __dsub_rn(mind, c_px); clamped = true; }
            if (r_out > 0.0) {
                sum = n_here == 0 ? motion : __fadd_rn(sum, motion);
                size_pre = clamped ? (float)r_out : radius;
                ++n_here;
                if (n_obs < P.pomax && lane == 0) {
                    double* o = P.obs + (p * P.pomax + n_obs) * 3;
                    o[0] = co1; o[1] = co0; o[2] = r_out;                                // [coord[1], coord[0], radius]
                }
                if (n_obs >= P.pomax) status |= 4;
                ++n_obs;
            }
        }
    }
    if (lane == 0) {
        if (P.hull_cnt[p] > P.hmax || P.isle_cnt[p] > P.hmax) status |= 8;   // hull / isle capacity overflow: unusable bank entry
        P.obs_cnt[p] = min(n_obs, P.pomax);
        P.obst_rand_used[p] = used;
        P.status[p] = status;
    }
}

__global__ void neg_kernel(const double* __restrict__ a, double* __restrict__ b, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = -a[i];
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_path_synthesize(const ppnet_path_params* pp_, void* stream) {
    PPNET_REQUIRE(pp_, "path_synthesize: null params");
    const ppnet_path_params& P = *pp_;
    cudaStream_t st = (cudaStream_t)stream;
    PPNET_REQUIRE(P.n_paths >= 0, "path_synthesize: negative n_paths");
    if (P.n_paths == 0) return PPNET_OK;
    PPNET_REQUIRE(P.seg_num >= 1 && P.seg_num <= kMaxSeg, "path_synthesize: seg_num must be in 1..%d", kMaxSeg);
    PPNET_REQUIRE(P.poly_order >= 2 && P.poly_order <= kMaxOrder, "path_synthesize: poly_order must be in 2..%d", kMaxOrder);
    PPNET_REQUIRE(P.hmax >= 3 && P.hmax <= kMaxHull && P.pomax >= 1 && P.max_obst_iter >= 1, "path_synthesize: bad hmax/pomax/max_obst_iter");
    PPNET_REQUIRE(P.resolution > 0 && P.map_size > 0, "path_synthesize: bad resolution / map_size");
    PPNET_REQUIRE(P.poly && P.endpoint && P.is_straight && P.path_straight && P.seg_trans_local && P.grad_st && P.grad_end && P.seg_length &&
                  P.seg_rot && P.seg_trans && P.segpoint_raw && P.pathpoint_raw && P.length && P.cells && P.up && P.up_dir &&
                  P.down && P.cap_init && P.cap_end && P.boundary_raw && P.ray_x0 && P.ray_dir && P.step_num && P.hull_raw &&
                  P.hull_cnt && P.rotation && P.translation && P.hull && P.segpoint_img && P.pathpoint && P.boundary &&
                  P.isle && P.isle_cnt && P.obs && P.obs_cnt && P.obst_rand_used && P.status && P.neg_rotation_ws,
                  "path_synthesize: every output pointer except space_raw is required");
    PPNET_REQUIRE(!P.in_poly || P.in_uend, "path_synthesize: in_poly needs in_uend (= the end points)");
    PPNET_REQUIRE(!P.in_obst_rand || (P.in_obst_rand_cnt && P.max_obst_rand > 0), "path_synthesize: in_obst_rand needs counts");
    static std::mutex mu;
    static FitBasis basis[kMaxOrder + 1];
    static bool have[kMaxOrder + 1] = {};
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!have[P.poly_order]) { basis[P.poly_order] = make_basis(P.poly_order); have[P.poly_order] = true; }
    }
    const int64_t n_seg = P.n_paths * P.seg_num;
    pathseg_kernel<<<(unsigned)((n_seg + 3) / 4), 128, 0, st>>>(P, basis[P.poly_order]);
    PPNET_LAUNCH_CHECK("pathseg_kernel");
    chain_kernel<<<(unsigned)P.n_paths, kChainThreads, 0, st>>>(P);
    PPNET_LAUNCH_CHECK("chain_kernel");
    const int Np = kSegSamples * P.seg_num, Nb = 2 * kBndSamples * P.seg_num + 2 * kBndSamples;
    if (P.space_raw) {                                     // A5: the corridor, painted on a 2R x 2R canvas at offset R
        const int W2 = 2 * (int)P.resolution;
        PPNET_CUDA(cudaMemsetAsync(P.space_raw, 0, (size_t)P.n_paths * W2 * W2, st));
        int rc = ppnet_corridor_paint(P.ray_x0, P.ray_dir, P.step_num, P.n_paths, Nb, P.map_size, P.resolution, P.resolution,
                                      W2, W2, 255, P.space_raw, stream);
        if (rc != PPNET_OK) return rc;
    }
    if (P.in_hull) {                                       // parity: the reference's own vertex order (Qhull's start vertex is arbitrary)
        PPNET_REQUIRE(P.in_hull_cnt, "path_synthesize: in_hull needs in_hull_cnt");
        PPNET_CUDA(cudaMemcpyAsync(P.hull_raw, P.in_hull, sizeof(int32_t) * 2 * (size_t)P.hmax * P.n_paths, cudaMemcpyDeviceToDevice, st));
        PPNET_CUDA(cudaMemcpyAsync(P.hull_cnt, P.in_hull_cnt, sizeof(int32_t) * (size_t)P.n_paths, cudaMemcpyDeviceToDevice, st));
    } else {
        int rc = ppnet_hull2d_i32(P.cells, Np, P.n_paths, P.hmax, P.hull_raw, P.hull_cnt, stream);   // A6
        if (rc != PPNET_OK) return rc;
    }
    normalize_kernel<<<(unsigned)P.n_paths, kChainThreads, 0, st>>>(P);
    PPNET_LAUNCH_CHECK("normalize_kernel");
    if (P.space_raw && P.space) {                          // A7 mask: rotate by -Rotation, translate by Translation, crop R x R
        neg_kernel<<<(unsigned)((P.n_paths + 127) / 128), 128, 0, st>>>(P.rotation, P.neg_rotation_ws, P.n_paths);
        PPNET_LAUNCH_CHECK("neg_kernel");
        int rc = ppnet_mask_rigid(P.space_raw, 2 * (int)P.resolution, P.neg_rotation_ws, P.translation, P.n_paths,
                                  (int)P.resolution, P.space, stream);
        if (rc != PPNET_OK) return rc;
    }
    isle_kernel<<<(unsigned)P.n_paths, kIsleThreads, 0, st>>>(P);
    PPNET_LAUNCH_CHECK("isle_kernel");
    obstacles_kernel<<<(unsigned)((P.n_paths + 3) / 4), 128, 0, st>>>(P);
    PPNET_LAUNCH_CHECK("obstacles_kernel");
    return PPNET_OK;
}
