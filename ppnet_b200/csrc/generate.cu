// The fused map generator: MapGenerate.generate's inner block + generate_map_randomly, one CTA per map.
//
//   A10  Path.boundary_check                 EDaGe-PP/Path.py:100-111      placement rejection
//   A13  MapGenerate.generate placement      EDaGe-PP/MapGenerate.py:63-93 rigid transform -> label
//   A14  MapGenerate.generate_map_randomly   EDaGe-PP/MapGenerate.py:126-151 obstacle draws + clearance verdict
//   A15  plot_obstacles (geometric restatement, raster.cuh) -> bit-packed occupancy map
//
// Per map g (global index, Philox counters keyed by g => any sharding gives identical output):
//   1. warp 0 runs the rejection loop 32 tries at a time: try t draws (angle, t0, t1), rotates the
//      target path's hull vertices and tests 0 <= h' < R; the first passing try wins (= the sequential
//      `while` of the reference).  In parity mode the draws come from the caller instead.
//   2. all threads rotate/translate the 1000 path points + 11 segment points (label, 16.2 KB -- the
//      dominant HBM stream, written with 16-byte stores) and keep the odd-indexed points in shared memory.
//   3. candidate circles: drawn (Philox, one thread per candidate) by warps 1.. while warp 0 is busy with step 1.
//      The staged odd points are grouped into boxes of 16
//      consecutive points; the (candidate, box) pairs are spread over the CTA (warp = candidate, lane = box,
//      no cross-lane reduction): a pair is skipped when an f32 test on the outward-rounded box puts the
//      candidate farther from it than the threshold plus a 1e-5 (R + |centre|) margin (those points cannot be the
//      ones that decide `min(dis) > r_px + c*R/M`), otherwise its points are evaluated with the reference's exact
//      un-fused arithmetic and folded into the candidate's minimum by a shared-memory atomicMin on the bit
//      pattern.  One sqrt per candidate; ordered compaction by ballot scan.
//   4. the path-hugging obstacles of the target path are placed with the same rigid transform and appended.
//   5. optional: all obstacles rasterised into a shared-memory bitmap -- (disk, row) span tasks spread flat over
//      the CTA through a prefix over the disks' row counts -- and streamed out.
// The target-path bank (16 KB / path) is read through L2; nothing else is read from HBM.
#include <math_constants.h>

#include "common.cuh"
#include "raster.cuh"

namespace ppnet {

constexpr int kGenThreads = 128;
constexpr int kGenWarps = kGenThreads / 32;
constexpr int kBlkPts = 16;

// Path.coord_rotation (Path.py:271-274) = np.dot(2x2, 2xN): OpenBLAS dgemm accumulates with FMA on
// AVX2/AVX-512 hosts: out = fma(r01, x1, r00 * x0).   (waypoint parity is 1e-5; verdict parity of
// boundary_check is exact except within an ulp of 0 / R)
__device__ __forceinline__ void rot2(double c, double s, double x0, double x1, double& r0, double& r1) {
    r0 = __fma_rn(-s, x1, __dmul_rn(c, x0));
    r1 = __fma_rn(c, x1, __dmul_rn(s, x0));
}

// boundary_check(angle_arg, [ta, tb]):  h' = Rot(angle_arg/180*pi).(h - R/2) + t + R/2 ; ok iff all in [0, R)
__device__ __forceinline__ bool hull_inside(const double2* __restrict__ hull, int H, double angle_arg, double ta,
                                            double tb, double R) {
    const double off = __dmul_rn(R, 0.5);
    const double th = __dmul_rn(__ddiv_rn(angle_arg, 180.0), CUDART_PI);
    double s, c;
    sincos(th, &s, &c);
    bool ok = true;
    for (int i = 0; i < H; ++i) {
        const double2 h = hull[i];
        double r0, r1;
        rot2(c, s, __dsub_rn(h.x, off), __dsub_rn(h.y, off), r0, r1);
        const double h0 = __dadd_rn(__dadd_rn(r0, ta), off), h1 = __dadd_rn(__dadd_rn(r1, tb), off);
        ok = ok && !(h0 < 0.0 || h0 >= R || h1 < 0.0 || h1 >= R);
    }
    return ok;
}

// draws of placement try t of map g:  block 2t -> (angle, t0), block 2t+1 -> (t1, -)
__device__ __forceinline__ void draw_placement(uint2 key, uint64_t g, uint32_t t, double R, double& angle, int& t0,
                                               int& t1) {
    const uint4 a = Philox::gen(key, make_uint4(2u * t, STREAM_PLACE, (uint32_t)g, (uint32_t)(g >> 32)));
    const uint4 b = Philox::gen(key, make_uint4(2u * t + 1u, STREAM_PLACE, (uint32_t)g, (uint32_t)(g >> 32)));
    angle = __dsub_rn(__dmul_rn(u53(a.x, a.y), 360.0), 180.0);                   // random([1])*360 - 180
    const double half = __dmul_rn(R, 0.5);
    t0 = (int)__dsub_rn(__dmul_rn(u53(a.z, a.w), R), half);                      // np.array(.., dtype=int): trunc
    t1 = (int)__dsub_rn(__dmul_rn(u53(b.x, b.y), R), half);
}

// candidate j of map g: block j -> (x, y) ; block O + j/2, half j&1 -> r      (map units)
__device__ __forceinline__ void draw_candidate(uint2 key, uint64_t g, int j, int O, double M, double osize, double& x,
                                               double& y, double& r) {
    const uint4 a = Philox::gen(key, make_uint4((uint32_t)j, STREAM_OBST, (uint32_t)g, (uint32_t)(g >> 32)));
    const uint4 b = Philox::gen(key, make_uint4((uint32_t)(O + (j >> 1)), STREAM_OBST, (uint32_t)g, (uint32_t)(g >> 32)));
    x = __dmul_rn(u53(a.x, a.y), M);                                             // random(O) * MapSize
    y = __dmul_rn(u53(a.z, a.w), M);
    r = __dmul_rn((j & 1) ? u53(b.z, b.w) : u53(b.x, b.y), osize);               // random(O) * ObstacleSize
}

__device__ __forceinline__ double warp_min_d2_gen(const double2* __restrict__ pts, int n, double q0, double q1) {
    const int lane = threadIdx.x & 31;
    double m0 = CUDART_INF, m1 = CUDART_INF;
    int i = lane;
    for (; i + 32 < n; i += 64) {
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

struct GenShared {
    double angle;
    int t0, t1, tries, n_acc;
};

#ifndef PPNET_GEN_MINB
#define PPNET_GEN_MINB 8
#endif
__global__ void __launch_bounds__(kGenThreads, PPNET_GEN_MINB)
generate_kernel(ppnet_gen_params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: odd points [np/2] double2 | hull [hmax] double2 | bitmap [R*W padded to 4] words (optional) |
    //         obstacle triples [O + pomax] x 3 doubles (candidate k at slot k, path obstacle k at slot O + k) |
    //         per-candidate threshold / cull radius / running min d^2 [O] each | raster task prefix | accept flags [O]
    const int n_odd = P.np / 2;
    const int O = P.obstacles_num;
    const int omax_out = O + P.pomax;
    int W = 0, words = 0;
    if (P.out_bits) { W = ((int)P.resolution + 31) / 32; words = (int)P.resolution * W; }
    const int n_blk = (n_odd + kBlkPts - 1) / kBlkPts;       // boxes of kBlkPts consecutive odd points
    double2* odd = reinterpret_cast<double2*>(smem_raw);
    double2* hull = odd + n_odd;
    double4* box = reinterpret_cast<double4*>(hull + P.hmax + (P.hmax & 1));    // (row_lo, row_hi, col_lo, col_hi), 32-B aligned
    float4* boxf = reinterpret_cast<float4*>(box);                             // the boxes as stored: f32, rounded outward
    uint32_t* bm = reinterpret_cast<uint32_t*>(box + n_blk);
    double* sobs = reinterpret_cast<double*>(bm + ((words + 3) & ~3));
    double* thr_s = sobs + 3 * omax_out;                                       // [O] r_px + c_px
    double* cull_s = thr_s + O;                                                // [O] squared cull radius
    unsigned long long* m2_s = reinterpret_cast<unsigned long long*>(cull_s + O);   // [O] bits of min d^2 (non-negative doubles order like integers)
    int* task_s = reinterpret_cast<int*>(m2_s + O);                            // [omax_out + 1] prefix of raster row tasks
    int* row0_s = task_s + omax_out + 1;                                       // [omax_out] first row of each disk
    uint8_t* acc_s = reinterpret_cast<uint8_t*>(row0_s + omax_out);
    __shared__ GenShared sh;

    const int64_t lm = blockIdx.x;                        // local map index
    const uint64_t g = (uint64_t)(P.map0 + lm);           // global map index
    const int j = (int)((g / (uint64_t)P.reps) % (uint64_t)P.n_bank);   // target path (MapGenerate.py:68)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint2 key = make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32));
    const double R = P.resolution;
    const int H = min(P.bank_hull_cnt[j], P.hmax);

    for (int i = threadIdx.x; i < H; i += kGenThreads)
        hull[i] = reinterpret_cast<const double2*>(P.bank_hull)[(size_t)j * P.hmax + i];
    __syncthreads();
    const double M = P.map_size;
    const double thr_c = __dmul_rn(__ddiv_rn(P.clearance, M), R);             // c / M * R

    // ---- 1. placement: rejection loop, 32 tries per round (first passing try wins) -------------------
    //         ... while the other warps draw the candidate circles (neither depends on the other) and clear the bitmap
    if (warp != 0) {
        for (int i = threadIdx.x - 32; i < words; i += kGenThreads - 32) bm[i] = 0u;
        for (int k = threadIdx.x - 32; k < O; k += kGenThreads - 32) {        // one Philox draw per candidate (:128-130)
            double x, y, r;
            if (P.in_cand) {
                const double* c = P.in_cand + ((size_t)lm * O + k) * 3;
                x = c[0]; y = c[1]; r = c[2];
            } else {
                draw_candidate(key, g, k, O, M, P.obstacle_size, x, y, r);
            }
            double* o = sobs + 3 * k;                      // [coord_img[1], coord_img[0], radius_img]  (:134-136, :143)
            const double q1 = __dmul_rn(__ddiv_rn(y, M), R), q0 = __dmul_rn(__ddiv_rn(x, M), R);
            const double rimg = __dmul_rn(__ddiv_rn(r, M), R);
            o[0] = q1; o[1] = q0; o[2] = rimg;
            const double thr = __dadd_rn(rimg, thr_c);
            // squared cull radius for the f32 box test: (thr + margin)^2, the margin ~100x the float rounding of the
            // coordinates involved (boxes are rounded outward, so only the centre's and the subtractions' rounding count)
            const double cr = thr + 1e-5 * (R + fabs(q0) + fabs(q1) + fabs(thr));
            thr_s[k] = thr;
            cull_s[k] = cr * cr * (1.0 + 1e-5);
            m2_s[k] = 0x7ff0000000000000ull;               // +inf
        }
    }
    if (warp == 0) {
        int tries = 0;
        double angle = 0.0;
        int t0 = 0, t1 = 0;
        if (P.bank_hull_cnt[j] > P.hmax) {
            // the bank's hull was truncated (ppnet_hull2d_i32 reports the true vertex count): a placement test on a partial
            // hull could accept a path that leaves the image, so the map is flagged invalid (tries = 0) instead
        } else if (P.in_angle) {                          // parity mode: the caller supplies the draws
            angle = P.in_angle[lm];
            t0 = P.in_trans[2 * lm];
            t1 = P.in_trans[2 * lm + 1];
            // MapGenerate.py:66  boundary_check(-angle, [translation[1], translation[0]])
            const bool ok = hull_inside(hull, H, -angle, (double)t1, (double)t0, R);
            tries = ok ? 1 : 0;
        } else {
            for (int base = 0; base < P.max_tries && tries == 0; base += 32) {
                const int t = base + lane;
                double a;
                int x0, x1;
                draw_placement(key, g, (uint32_t)t, R, a, x0, x1);
                const bool ok = t < P.max_tries && hull_inside(hull, H, -a, (double)x1, (double)x0, R);
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (bal) {
                    const int src = __ffs(bal) - 1;
                    angle = __shfl_sync(0xffffffffu, a, src);
                    t0 = __shfl_sync(0xffffffffu, x0, src);
                    t1 = __shfl_sync(0xffffffffu, x1, src);
                    tries = base + src + 1;
                }
            }
        }
        if (lane == 0) { sh.angle = angle; sh.t0 = t0; sh.t1 = t1; sh.tries = tries; }
    }
    __syncthreads();
    const int tries = sh.tries;
    const bool valid = tries > 0;
    if (threadIdx.x == 0) {
        if (P.out_angle) P.out_angle[lm] = sh.angle;
        if (P.out_trans) { P.out_trans[2 * lm] = sh.t0; P.out_trans[2 * lm + 1] = sh.t1; }
        if (P.out_tries) P.out_tries[lm] = tries;
        if (P.out_valid) P.out_valid[lm] = valid ? 1 : 0;
    }
    if (!valid) {                                         // retry budget exhausted: flagged, not silently dropped
        if (threadIdx.x == 0) {
            if (P.out_obs_cnt) P.out_obs_cnt[lm] = 0;
            if (P.out_rand_cnt) P.out_rand_cnt[lm] = 0;
            if (P.counters) { atomicAdd(P.counters + 0, 1ull); atomicAdd(P.counters + 3, (unsigned long long)P.max_tries); }
        }
        if (P.out_bits) store_bitmap(bm, P.out_bits + (size_t)lm * words, words);
        return;
    }

    // ---- 2. rigid transform of the label points (MapGenerate.py:70-80) ---------------------------------
    const double off = __dmul_rn(R, 0.5);
    const double th = __dmul_rn(__ddiv_rn(-sh.angle, 180.0), CUDART_PI);      // -angle/180*pi
    double sn, cs;
    sincos(th, &sn, &cs);
    const double tr = (double)sh.t1, tc = (double)sh.t0;                      // + [translation[1], translation[0]]
    {
        const double2* src = reinterpret_cast<const double2*>(P.bank_pathpt) + (size_t)j * P.np;
        double2* dst = P.out_pathpt ? reinterpret_cast<double2*>(P.out_pathpt) + (size_t)lm * P.np : nullptr;
        for (int i = threadIdx.x; i < P.np; i += kGenThreads) {
            const double2 p = __ldg(src + i);
            double r0, r1;
            rot2(cs, sn, __dsub_rn(p.x, off), __dsub_rn(p.y, off), r0, r1);
            double2 q;
            q.x = __dadd_rn(__dadd_rn(r0, off), tr);
            q.y = __dadd_rn(__dadd_rn(r1, off), tc);
            if (dst) dst[i] = q;
            if (i & 1) odd[i >> 1] = q;
        }
        if (P.out_segpt) {
            const double2* ssrc = reinterpret_cast<const double2*>(P.bank_segpt) + (size_t)j * P.nseg1;
            double2* sdst = reinterpret_cast<double2*>(P.out_segpt) + (size_t)lm * P.nseg1;
            for (int i = threadIdx.x; i < P.nseg1; i += kGenThreads) {
                const double2 p = __ldg(ssrc + i);
                double r0, r1;
                rot2(cs, sn, __dsub_rn(p.x, off), __dsub_rn(p.y, off), r0, r1);
                double2 q;
                q.x = __dadd_rn(__dadd_rn(r0, off), tr);
                q.y = __dadd_rn(__dadd_rn(r1, off), tc);
                sdst[i] = q;
            }
        }
    }
    __syncthreads();

    // ---- 3. candidate circles: clearance verdict (MapGenerate.py:132-143) --------------------------------
    for (int b = threadIdx.x; b < n_blk; b += kGenThreads) {                  // boxes of the staged odd points
        double4 bb = make_double4(CUDART_INF, -CUDART_INF, CUDART_INF, -CUDART_INF);
        const int e = min(n_odd, (b + 1) * kBlkPts);
        for (int i = b * kBlkPts; i < e; ++i) {
            const double2 p = odd[i];
            bb.x = fmin(bb.x, p.x); bb.y = fmax(bb.y, p.x); bb.z = fmin(bb.z, p.y); bb.w = fmax(bb.w, p.y);
        }
        boxf[b] = make_float4(__double2float_rd(bb.x), __double2float_ru(bb.y), __double2float_rd(bb.z), __double2float_ru(bb.w));
    }
    __syncthreads();
    // (candidate, box) pairs over the whole CTA: a pair whose box is farther than the cull radius cannot hold
    // the point that decides `min(dis) > r_px + c_px`; the others evaluate their 16 points with the reference's
    // un-fused arithmetic and fold into the candidate's minimum with one shared-memory atomicMin.
    for (int k = warp; k < O; k += kGenWarps) {            // warp <-> candidate, lane <-> box: no index arithmetic
        const double q1 = sobs[3 * k], q0 = sobs[3 * k + 1];
        const float q0f = (float)q0, q1f = (float)q1, cull2 = __double2float_ru(cull_s[k]);
        for (int b = lane; b < n_blk; b += 32) {
            const float4 bb = boxf[b];
            const float ex = fmaxf(fmaxf(bb.x - q0f, q0f - bb.y), 0.0f), ey = fmaxf(fmaxf(bb.z - q1f, q1f - bb.w), 0.0f);
            if (ex * ex + ey * ey > cull2) continue;       // NaN never culls
            double m2 = CUDART_INF;
            const int e = min(n_odd, (b + 1) * kBlkPts);
            for (int i = b * kBlkPts; i < e; ++i) {
                const double2 p = odd[i];
                const double ax = __dsub_rn(p.x, q0), ay = __dsub_rn(p.y, q1);     // scipy euclidean: un-fused
                m2 = fmin(m2, __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
            }
            atomicMin(m2_s + k, (unsigned long long)__double_as_longlong(m2));
        }
    }
    __syncthreads();
    // verdict + ordered compaction of the accepted candidates (the reference appends in loop order)
    if (warp == 0) {
        int base = 0;
        double* gout = P.out_obs ? P.out_obs + (size_t)lm * omax_out * 3 : nullptr;
        for (int k0 = 0; k0 < O; k0 += 32) {
            const int k = k0 + lane;
            // sqrt is monotone: sqrt(min d^2) == min sqrt(d^2); skipped boxes only hold points with d > thr   (:142)
            const bool ok = k < O && __dsqrt_rn(__longlong_as_double((long long)m2_s[k])) > thr_s[k];
            if (k < O) acc_s[k] = ok ? 1 : 0;
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (ok && gout) {
                const int d = base + __popc(bal & ((1u << lane) - 1));
                const double* o = sobs + 3 * k;
                gout[3 * d] = o[0]; gout[3 * d + 1] = o[1]; gout[3 * d + 2] = o[2];
            }
            base += __popc(bal);
        }
        if (lane == 0) sh.n_acc = base;
    }
    __syncthreads();
    const int n_acc = sh.n_acc;

    // ---- 4. path-hugging obstacles of the target path (MapGenerate.py:83-89) ---------------------------
    const int n_po = min(P.bank_obs_cnt ? P.bank_obs_cnt[j] : 0, P.pomax);
    for (int k = threadIdx.x; k < n_po; k += kGenThreads) {
        const double* ob = P.bank_obs + ((size_t)j * P.pomax + k) * 3;
        double r0, r1;
        rot2(cs, sn, __dsub_rn(ob[1], off), __dsub_rn(ob[0], off), r0, r1);   // coord = [obs[1], obs[0]] - R/2
        r0 = __dadd_rn(__dadd_rn(r0, off), tr);
        r1 = __dadd_rn(__dadd_rn(r1, off), tc);
        if (P.out_obs) {
            double* o = P.out_obs + ((size_t)lm * omax_out + n_acc + k) * 3;
            o[0] = r1; o[1] = r0; o[2] = ob[2];                                // [coord[1], coord[0], r]
        }
        if (P.out_bits) { double* o = sobs + 3 * (O + k); o[0] = r1; o[1] = r0; o[2] = ob[2]; }
    }
    if (threadIdx.x == 0) {
        if (P.out_obs_cnt) P.out_obs_cnt[lm] = n_acc + n_po;
        if (P.out_rand_cnt) P.out_rand_cnt[lm] = n_acc;
        if (P.counters) {
            atomicAdd(P.counters + 0, 1ull);
            atomicAdd(P.counters + 1, 1ull);
            atomicAdd(P.counters + 2, (unsigned long long)n_acc);
            atomicAdd(P.counters + 3, (unsigned long long)tries);
        }
    }

    // ---- 5. occupancy bitmap ----------------------------------------------------------------------------
    if (P.out_bits) {
        __syncthreads();
        // (disk, row) tasks flattened over the CTA: every thread gets the same number of row spans whatever the radii
        const int n_disk = O + n_po;
        PPNET_ASSERT(n_disk <= omax_out);
        for (int k = threadIdx.x; k < n_disk; k += kGenThreads) {
            int i0 = 0, i1 = 0;
            if (k >= O || acc_s[k])                        // rejected candidates paint nothing
                disk_rows((int)R, sobs[3 * k], sobs[3 * k + 1], __dadd_rn(sobs[3 * k + 2], P.raster_inflate), i0, i1);
            row0_s[k] = i0;
            task_s[k] = i1 - i0;
        }
        __syncthreads();
        if (warp == 0) {                                   // exclusive prefix of the row counts
            int carry = 0;
            for (int k0 = 0; k0 < n_disk; k0 += 32) {
                const int k = k0 + lane;
                const int v = k < n_disk ? task_s[k] : 0;
                int inc = v;
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, inc, sft);
                    if (lane >= sft) inc += o;
                }
                if (k < n_disk) task_s[k] = carry + inc - v;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) task_s[n_disk] = carry;
        }
        __syncthreads();
        const int total = task_s[n_disk];
        // a contiguous run of tasks per thread: the disk index is found once (binary search), then advances rarely
        const int per = (total + kGenThreads - 1) / kGenThreads;
        const int t_lo = min(total, (int)threadIdx.x * per), t_hi = min(total, t_lo + per);
        if (t_lo < t_hi) {
            int k = 0, hi = n_disk;                        // largest k with task_s[k] <= t_lo
            while (hi - k > 1) { const int mid = (k + hi) >> 1; if (task_s[mid] <= t_lo) k = mid; else hi = mid; }
            for (int t = t_lo; t < t_hi; ++t) {
                while (t >= task_s[k + 1]) ++k;
                PPNET_ASSERT(k < n_disk && row0_s[k] + (t - task_s[k]) >= 0 && row0_s[k] + (t - task_s[k]) < (int)R);
                raster_disk_row(bm, (int)R, W, sobs[3 * k], sobs[3 * k + 1], __dadd_rn(sobs[3 * k + 2], P.raster_inflate),
                                row0_s[k] + (t - task_s[k]));
            }
        }
        __syncthreads();
        store_bitmap(bm, P.out_bits + (size_t)lm * words, words);
    }
}

// stand-alone A10 for the Python `Path.boundary_check` shim and the parity tests
__global__ void boundary_check_kernel(const double* __restrict__ hull, const int32_t* __restrict__ hull_cnt, int hmax,
                                      const int32_t* __restrict__ path_idx, const double* __restrict__ angle_arg,
                                      const double* __restrict__ trans_arg, int64_t n, double R,
                                      uint8_t* __restrict__ ok, double* __restrict__ hull_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = path_idx ? path_idx[i] : 0;
    const int H = min(hull_cnt[j], hmax);
    const double2* h = reinterpret_cast<const double2*>(hull) + (size_t)j * hmax;
    const double ta = trans_arg[2 * i], tb = trans_arg[2 * i + 1];
    ok[i] = hull_inside(h, H, angle_arg[i], ta, tb, R) ? 1 : 0;
    if (hull_out) {
        const double off = __dmul_rn(R, 0.5);
        const double th = __dmul_rn(__ddiv_rn(angle_arg[i], 180.0), CUDART_PI);
        double s, c;
        sincos(th, &s, &c);
        for (int k = 0; k < H; ++k) {
            double r0, r1;
            rot2(c, s, __dsub_rn(h[k].x, off), __dsub_rn(h[k].y, off), r0, r1);
            hull_out[((size_t)i * hmax + k) * 2] = __dadd_rn(__dadd_rn(r0, ta), off);
            hull_out[((size_t)i * hmax + k) * 2 + 1] = __dadd_rn(__dadd_rn(r1, tb), off);
        }
    }
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_generate_maps(const ppnet_gen_params* p, void* stream) {
    PPNET_REQUIRE(p, "generate_maps: null params");
    PPNET_REQUIRE(p->n_maps >= 0 && p->n_bank > 0 && p->np >= 0 && p->nseg1 >= 0 && p->hmax > 0 && p->pomax >= 0,
                  "generate_maps: bad sizes");
    if (p->n_maps == 0) return PPNET_OK;
    PPNET_REQUIRE(p->bank_pathpt && p->bank_hull && p->bank_hull_cnt, "generate_maps: bank pointers are null");
    PPNET_REQUIRE(p->nseg1 == 0 || p->bank_segpt || !p->out_segpt, "generate_maps: bank_segpt is null");
    PPNET_REQUIRE(p->pomax == 0 || (p->bank_obs && p->bank_obs_cnt), "generate_maps: bank_obs is null");
    PPNET_REQUIRE(p->reps > 0 && p->obstacles_num >= 0 && p->max_tries > 0, "generate_maps: bad reps/O/max_tries");
    PPNET_REQUIRE(p->resolution > 0 && p->resolution <= 4096 && p->resolution == (double)(int)p->resolution,
                  "generate_maps: resolution must be a positive integer");
    PPNET_REQUIRE((p->in_angle == nullptr) == (p->in_trans == nullptr), "generate_maps: in_angle and in_trans go together");
    const int R = (int)p->resolution, W = (R + 31) / 32;
    size_t smem = sizeof(double2) * (size_t)(p->np / 2 + p->hmax + 1) + 32 * (size_t)((p->np / 2 + kBlkPts - 1) / kBlkPts) +
                  sizeof(double) * 3 * (size_t)(p->obstacles_num + p->pomax) + 24 * (size_t)p->obstacles_num +
                  8 * (size_t)(p->obstacles_num + p->pomax + 2) + (size_t)p->obstacles_num + 64;
    if (p->out_bits) smem += (size_t)(((R * W) + 3) & ~3) * 4;
    PPNET_REQUIRE(smem <= 220 * 1024, "generate_maps: shared memory budget exceeded (%zu bytes)", smem);
    if (smem > 48 * 1024)
        PPNET_CUDA(cudaFuncSetAttribute(generate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    generate_kernel<<<(unsigned)p->n_maps, kGenThreads, smem, (cudaStream_t)stream>>>(*p);
    PPNET_LAUNCH_CHECK("generate_kernel");
    return PPNET_OK;
}

extern "C" int ppnet_boundary_check(const double* hull, const int32_t* hull_cnt, int32_t hmax, const int32_t* path_idx,
                                    const double* angle_arg, const double* trans_arg, int64_t n, double resolution,
                                    uint8_t* ok, double* hull_out, void* stream) {
    PPNET_REQUIRE(n >= 0 && hmax > 0, "boundary_check: bad sizes");
    if (n == 0) return PPNET_OK;
    PPNET_REQUIRE(hull && hull_cnt && angle_arg && trans_arg && ok, "boundary_check: null pointer");
    boundary_check_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        hull, hull_cnt, hmax, path_idx, angle_arg, trans_arg, n, resolution, ok, hull_out);
    PPNET_LAUNCH_CHECK("boundary_check_kernel");
    return PPNET_OK;
}
