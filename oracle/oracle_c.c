/* TEST INFRASTRUCTURE ONLY -- plain-C restatement (oracle) of PPNet's EDaGe-PP hot path.
 *
 * This is the checker / CPU baseline, never the product.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs load it (oracle/_build/liboracle.so).
 * It restates the same reference lines as oracle/ppnet_oracle.py (which is pinned against the
 * real reference by tests/golden), one individually rounded IEEE operation per reference
 * operation.  Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).  Single-threaded per
 * call; oracle/c_oracle.py fans ranges out over host threads (ctypes drops the GIL).
 * Contraction MUST stay off; the only fused operations are the explicit fma() calls that model
 * OpenBLAS' SkylakeX ddot (dot_mode 0).  Citations: files under the PPNet tree.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DOT_FUSED_SKX 0
#define DOT_UNFUSED 1

static inline double dot2(double a0, double a1, double b0, double b1, int mode) {
    if (mode == DOT_FUSED_SKX) return fma(a1, b1, a0 * b0);
    return a0 * b0 + a1 * b1;
}

/* ---- A11 process_map.collision_check_circle_edge (EDaGe-PP/process_map.py:383-425) --------
 * pts_rc: (s_r, s_c, e_r, e_c) as the reference receives them; obs rows [x, y, r]. */
static int segcheck_f64_one(const double* p, const double* obs, int cnt, double clearance,
                            double bound, int mode) {
    double s_r = p[0], s_c = p[1], e_r = p[2], e_c = p[3];
    if (s_r < 0 || s_c > bound) return 1;                    /* :384-387 */
    if (e_r < 0 || e_c > bound) return 1;
    double s0 = s_c, s1 = s_r, e0 = e_c, e1 = e_r;           /* :388-389 */
    double d0 = e0 - s0, d1 = e1 - s1;
    double L = sqrt(dot2(d0, d1, d0, d1, mode));             /* :391 */
    double n0 = d1 / L, n1 = (-d0) / L;
    for (int k = 0; k < cnt; ++k) {
        double o0 = (double)(float)obs[3 * k], o1 = (double)(float)obs[3 * k + 1];   /* :396 */
        double thr = obs[3 * k + 2] + clearance / 2;
        double v0 = e0 - o0, v1 = e1 - o1;
        if (sqrt(v0 * v0 + v1 * v1) < thr) return 1;          /* :397 */
        double dis = dot2(n0, n1, o0 - s0, o1 - s1, mode);    /* :406 */
        if (dis > 0) { n0 = -n0; n1 = -n1; }                  /* :407-408 */
        double a = fabs(dis);
        double p0 = o0 + a * n0, p1 = o1 + a * n1;            /* :410 */
        double u0 = p0 - s0, u1 = p1 - s1;
        double nu = sqrt(dot2(u0, u1, u0, u1, mode));
        u0 = u0 / nu; u1 = u1 / nu;
        double w0 = p0 - e0, w1 = p1 - e1;
        double nw = sqrt(dot2(w0, w1, w0, w1, mode));
        w0 = w0 / nw; w1 = w1 / nw;
        if (a < thr && dot2(u0, u1, w0, w1, mode) < 0) return 1;   /* :415 */
    }
    return 0;
}

void orc_segcheck_f64(const double* pts_rc, const int32_t* seg_map, const double* obs,
                      const int32_t* obs_cnt, int omax, double clearance, double bound,
                      int dot_mode, long n, uint8_t* verdict) {
    for (long i = 0; i < n; ++i) {
        int m = seg_map[i];
        verdict[i] = (uint8_t)segcheck_f64_one(pts_rc + 4 * i, obs + (size_t)m * omax * 3,
                                               obs_cnt[m], clearance, bound, dot_mode);
    }
}

/* ---- A12 experiments/MPNet/neuralplanner.py:43-69, all float32 --------------------------- */
/* cmp64 = 0: NumPy >= 2 (NEP 50): the float32 offsets are compared with float32(size + clearance/2);
 * cmp64 = 1: NumPy 1.x (the reference's requirements.txt era): np.float32 < Python float promotes to float64. */
static int segcheck_f32_cmp(const float* p, const double* obs, int cnt, double clearance,
                            double bound, int cmp64) {
    float s0 = p[0], s1 = p[1], e0 = p[2], e1 = p[3];
    float fb = (float)bound;        /* NEP 50: python scalar compares in f32 (0 and 224 exact) */
    if (s0 < 0.0f || s1 > fb) return 1;
    if (e0 < 0.0f || e1 > fb) return 1;
    float d0 = e0 - s0, d1 = e1 - s1;
    float L = sqrtf(d0 * d0 + d1 * d1);
    float n0 = d1 / L, n1 = (-d0) / L;
    for (int k = 0; k < cnt; ++k) {
        float o0 = (float)obs[3 * k], o1 = (float)obs[3 * k + 1];
        double thr64 = obs[3 * k + 2] + clearance / 2;
        float thr = (float)thr64;                               /* f64 sum, one rounding to f32 */
        float v0 = e0 - o0, v1 = e1 - o1;
        float dv = sqrtf(v0 * v0 + v1 * v1);
        if (cmp64 ? ((double)dv < thr64) : (dv < thr)) return 1;
        float q0 = o0 - s0, q1 = o1 - s1;
        float dis = n0 * q0 + n1 * q1;
        if (dis > 0.0f) { n0 = -n0; n1 = -n1; }
        float a = fabsf(dis);
        float p0 = o0 + a * n0, p1 = o1 + a * n1;
        float u0 = p0 - s0, u1 = p1 - s1;
        float nu = sqrtf(u0 * u0 + u1 * u1);
        u0 = u0 / nu; u1 = u1 / nu;
        float w0 = p0 - e0, w1 = p1 - e1;
        float nw = sqrtf(w0 * w0 + w1 * w1);
        w0 = w0 / nw; w1 = w1 / nw;
        if ((cmp64 ? ((double)a < thr64) : (a < thr)) && (u0 * w0 + u1 * w1) < 0.0f) return 1;
    }
    return 0;
}
static int segcheck_f32_one(const float* p, const double* obs, int cnt, double clearance, double bound) {
    return segcheck_f32_cmp(p, obs, cnt, clearance, bound, 0);
}

static int steer_one(const float* a, const float* b, const double* obs, int cnt, double clearance,
                     double bound) {
    /* neuralplanner.py:86-92 */
    float d0 = a[0] - b[0], d1 = a[1] - b[1];
    float dist = sqrtf(d0 * d0 + d1 * d1);
    if (dist > 0.0f) {
        float p[4] = {a[0], a[1], b[0], b[1]};
        if (segcheck_f32_one(p, obs, cnt, clearance, bound)) return 0;
    }
    return 1;
}

void orc_segcheck_f32(const float* pts_xy, const int32_t* seg_map, const double* obs,
                      const int32_t* obs_cnt, int omax, double clearance, double bound, long n,
                      uint8_t* verdict, uint8_t* steer) {
    for (long i = 0; i < n; ++i) {
        int m = seg_map[i];
        const double* ob = obs + (size_t)m * omax * 3;
        verdict[i] = (uint8_t)segcheck_f32_one(pts_xy + 4 * i, ob, obs_cnt[m], clearance, bound);
        if (steer) steer[i] = (uint8_t)steer_one(pts_xy + 4 * i, pts_xy + 4 * i + 2, ob, obs_cnt[m],
                                                 clearance, bound);
    }
}

void orc_segcheck_f32_cmp(const float* pts_xy, const int32_t* seg_map, const double* obs,
                          const int32_t* obs_cnt, int omax, double clearance, double bound, int cmp64,
                          long n, uint8_t* verdict) {
    for (long i = 0; i < n; ++i) {
        int m = seg_map[i];
        verdict[i] = (uint8_t)segcheck_f32_cmp(pts_xy + 4 * i, obs + (size_t)m * omax * 3, obs_cnt[m],
                                               clearance, bound, cmp64);
    }
}

/* feasibility_check neuralplanner.py:96-102; paths in CSR form */
void orc_feasible(const float* wp, const int64_t* path_off, const int32_t* path_map,
                  const double* obs, const int32_t* obs_cnt, int omax, double clearance,
                  double bound, long n_paths, uint8_t* feasible, int64_t* n_checked) {
    for (long p = 0; p < n_paths; ++p) {
        int m = path_map[p];
        const double* ob = obs + (size_t)m * omax * 3;
        uint8_t ok = 1;
        int64_t checked = 0;
        for (int64_t i = path_off[p]; i + 1 < path_off[p + 1]; ++i) {
            ++checked;
            if (!steer_one(wp + 2 * i, wp + 2 * (i + 1), ob, obs_cnt[m], clearance, bound)) {
                ok = 0;
                break;
            }
        }
        feasible[p] = ok;
        if (n_checked) n_checked[p] = checked;
    }
}

/* lvc neuralplanner.py:123-138, literally recursive (tail call => loop with restart) */
void orc_lvc(const float* wp, const int64_t* path_off, const int32_t* path_map, const double* obs,
             const int32_t* obs_cnt, int omax, double clearance, double bound, long n_paths,
             float* out_wp, int32_t* out_len) {
    for (long p = 0; p < n_paths; ++p) {
        int m = path_map[p];
        const double* ob = obs + (size_t)m * omax * 3;
        int64_t base = path_off[p];
        int len = (int)(path_off[p + 1] - base);
        float* cur = out_wp + 2 * base;
        memcpy(cur, wp + 2 * base, sizeof(float) * 2 * (size_t)len);
        for (;;) {
            int found = 0;
            for (int i = 0; i < len - 1 && !found; ++i)
                for (int j = len - 1; j > i + 1; --j)
                    if (steer_one(cur + 2 * i, cur + 2 * j, ob, obs_cnt[m], clearance, bound) == 1) {
                        memmove(cur + 2 * (i + 1), cur + 2 * j, sizeof(float) * 2 * (size_t)(len - j));
                        len = i + 1 + (len - j);
                        found = 1;
                        break;
                    }
            if (!found) break;
        }
        out_len[p] = len;
    }
}

/* ---- A14 MapGenerate.generate_map_randomly (EDaGe-PP/MapGenerate.py:132-143) ------------- */
void orc_clearance_filter(const double* pathpt, int np_, const double* cand, int O, double M,
                          double R, double c, long n_maps, uint8_t* accept, double* out,
                          int32_t* out_cnt) {
    for (long m = 0; m < n_maps; ++m) {
        const double* pp = pathpt + (size_t)m * np_ * 2;
        int k = 0;
        for (int j = 0; j < O; ++j) {
            const double* it = cand + ((size_t)m * O + j) * 3;
            double q0 = it[0] / M * R, q1 = it[1] / M * R, rimg = it[2] / M * R;
            double m2 = INFINITY;
            for (int i = 1; i < np_; i += 2) {               /* `if i % 2` :139 */
                double dx = pp[2 * i] - q0, dy = pp[2 * i + 1] - q1;
                double d2 = dx * dx + dy * dy;
                if (d2 < m2) m2 = d2;
            }
            int ok = sqrt(m2) > rimg + c / M * R;            /* :142 */
            accept[(size_t)m * O + j] = (uint8_t)ok;
            if (ok) {
                double* o = out + ((size_t)m * O + k) * 3;
                o[0] = q1; o[1] = q0; o[2] = rimg;          /* :143 */
                ++k;
            }
        }
        out_cnt[m] = k;
    }
}

/* ---- A4 Path.coord_euclidean2image (EDaGe-PP/Path.py:378-386) ---------------------------- */
void orc_grid_index(const double* pts, long n, double map_size, double resolution, double off,
                    int64_t* out) {
    double step = map_size / resolution;
    for (long i = 0; i < 2 * n; ++i) out[i] = (int64_t)nearbyint(pts[i] / step + off);
}

/* ---- A5 Path.free_space_bydirection (EDaGe-PP/Path.py:397-404) ---------------------------
 * paints space[W*H] (row-major [cx][cy]) with 255 for n_rays rays; returns cells painted
 * (counting repaints). */
long orc_corridor_paint(const double* x0, const double* dir, const double* step_num, long n_rays,
                        double map_size, double resolution, double off, int W, int H,
                        uint8_t* space) {
    double step = map_size / resolution;
    long painted = 0;
    for (long r = 0; r < n_rays; ++r) {
        long ns = (long)nearbyint(step_num[r]);
        for (long i = 0; i < ns; ++i) {
            double vx = x0[2 * r] + (double)i * dir[2 * r];
            double vy = x0[2 * r + 1] + (double)i * dir[2 * r + 1];
            int64_t cx = (int64_t)nearbyint(vx / step + off);
            int64_t cy = (int64_t)nearbyint(vy / step + off);
            if (0 < cx && cx < W && 0 < cy && cy < H) {
                space[cx * H + cy] = 255;
                ++painted;
            } else
                break;
        }
    }
    return painted;
}

/* ---- A10 Path.boundary_check (EDaGe-PP/Path.py:100-111), dgemm modelled as FMA chain ------ */
int orc_boundary_check(const double* hull, int H, double angle_deg, double t0, double t1,
                       double R, double* out) {
    double off = R / 2, th = angle_deg / 180 * M_PI;
    double c = cos(th), s = sin(th);
    int ok = 1;
    for (int i = 0; i < H; ++i) {
        double x0 = hull[2 * i] - off, x1 = hull[2 * i + 1] - off;
        double r0 = fma(-s, x1, c * x0), r1 = fma(c, x1, s * x0);
        double h0 = (r0 + t0) + off, h1 = (r1 + t1) + off;
        if (out) { out[2 * i] = h0; out[2 * i + 1] = h1; }
        if (h0 < 0 || h0 >= R || h1 < 0 || h1 >= R) ok = 0;
    }
    return ok;
}

/* ---- geometric raster (A15 restated; parity unpinned) + integer DDA (new functionality) --- */
void orc_raster_circles_bits(const double* obs, const int32_t* obs_cnt, int omax, long n_maps,
                             int R, double inflate, uint32_t* bits) {
    int W = (R + 31) / 32;
    for (long m = 0; m < n_maps; ++m) {
        uint32_t* b = bits + (size_t)m * R * W;
        memset(b, 0, sizeof(uint32_t) * (size_t)R * W);
        for (int k = 0; k < obs_cnt[m]; ++k) {
            const double* o = obs + ((size_t)m * omax + k) * 3;
            double rr = o[2] + inflate;
            if (!(rr > 0) || !isfinite(rr) || !isfinite(o[0]) || !isfinite(o[1])) continue;
            double r2 = rr * rr;
            /* bounding box clamped to the image in double (a cast of +-1e300 to long is undefined) */
            double fi0 = floor(o[1] - rr - 1), fi1 = ceil(o[1] + rr + 1);
            double fj0 = floor(o[0] - rr - 1), fj1 = ceil(o[0] + rr + 1);
            long i0 = fi0 < 0 ? 0 : (fi0 > R ? R : (long)fi0), i1 = fi1 > R ? R : (fi1 < 0 ? 0 : (long)fi1);
            long j0 = fj0 < 0 ? 0 : (fj0 > R ? R : (long)fj0), j1 = fj1 > R ? R : (fj1 < 0 ? 0 : (long)fj1);
            for (long i = i0; i < i1; ++i) {
                double dy = ((double)i + 0.5) - o[1];
                for (long j = j0; j < j1; ++j) {
                    double dx = ((double)j + 0.5) - o[0];
                    if (dx * dx + dy * dy <= r2) b[i * W + (j >> 5)] |= 1u << (j & 31);
                }
            }
        }
    }
}

static inline int64_t snap_clamp(float v) {             /* A4 rule at step 1 / offset 0, clamped to +-2^29 */
    double r = nearbyint((double)v);
    if (r < -536870912.0) r = -536870912.0;
    if (r > 536870912.0) r = 536870912.0;
    return (int64_t)r;
}

static inline int64_t fdiv(int64_t a, int64_t b) {      /* floor division, b > 0 */
    int64_t q = a / b, r = a % b;
    return (r != 0 && r < 0) ? q - 1 : q;
}

void orc_dda_gridcheck(const uint32_t* bits, int R, const float* segs_xy, const int32_t* seg_map,
                       long n, uint8_t* verdict, int32_t* first_hit) {
    int W = (R + 31) / 32;
    for (long i = 0; i < n; ++i) {
        const uint32_t* b = bits + (size_t)seg_map[i] * R * W;
        const float* sg = segs_xy + 4 * i;
        if (sg[0] != sg[0] || sg[1] != sg[1] || sg[2] != sg[2] || sg[3] != sg[3]) {   /* NaN: blocked at k = 0 */
            verdict[i] = 1;
            if (first_hit) first_hit[i] = 0;
            continue;
        }
        int64_t x0 = snap_clamp(sg[0]), y0 = snap_clamp(sg[1]), x1 = snap_clamp(sg[2]), y1 = snap_clamp(sg[3]);
        int64_t dx = x1 - x0, dy = y1 - y0;
        int64_t nn = llabs(dx) > llabs(dy) ? llabs(dx) : llabs(dy);
        int hit = 0;
        int32_t fh = -1;
        for (int64_t k = 0; k <= nn; ++k) {
            int64_t cx = x0, cy = y0;
            if (nn) {
                cx += fdiv(2 * k * dx + nn, 2 * nn);
                cy += fdiv(2 * k * dy + nn, 2 * nn);
            }
            if (cx < 0 || cx >= R || cy < 0 || cy >= R || ((b[cy * W + (cx >> 5)] >> (cx & 31)) & 1u)) {
                hit = 1;
                fh = (int32_t)k;
                break;
            }
        }
        verdict[i] = (uint8_t)hit;
        if (first_hit) first_hit[i] = fh;
    }
}

