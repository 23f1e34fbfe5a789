// Segment-vs-circles collision verdicts, both reference flavours, bit-exact.
//
//   A11  process_map.collision_check_circle_edge   EDaGe-PP/process_map.py:383-425   (float64)
//   A12  neuralplanner.collision_check_circle_edge experiments/MPNet/neuralplanner.py:43-69
//        + steerTo :86-92, feasibility_check :96-102, lvc :123-138                   (float32)
//
// The batched verdict kernel itself lives in verdict.cu (one kernel for A11, A12 and both fused); this file keeps the
// entry points of the one-flavour forms and the per-path kernels (feasibility_check, lvc).  The notes below describe the
// decision forms and margins that kernel uses (helpers in segcheck.cuh).
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
#include <cstdlib>

#include "segcheck.cuh"

namespace ppnet {

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

namespace ppnet {
int verdict_a11_only(const double*, int64_t, const int64_t*, int64_t, int64_t, const double*, const int32_t*, int32_t, double, double,
                     int32_t, uint8_t*, cudaStream_t);
int verdict_a12_only(const float*, int64_t, const int64_t*, int64_t, int64_t, const double*, const int32_t*, int32_t, double, double,
                     int32_t, uint8_t*, uint8_t*, uint8_t*, cudaStream_t);
}

extern "C" int ppnet_segcheck_edage_f64(const double* pts_rc, int64_t n_segs, const int64_t* seg_off,
                                        int64_t segs_per_map, int64_t n_maps, const double* obs,
                                        const int32_t* obs_cnt, int32_t omax, double clearance,
                                        double bound, int32_t dot_mode, uint8_t* verdict, void* stream) {
    PPNET_REQUIRE(dot_mode == PPNET_DOT_FUSED_SKX || dot_mode == PPNET_DOT_UNFUSED, "bad dot_mode");
    PPNET_REQUIRE(verdict || n_segs == 0, "segcheck: verdict is null");
    int64_t dummy = 1;
    int rc = check_common(pts_rc, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, &dummy);
    if (rc != PPNET_OK || dummy == 0) return rc;
    return verdict_a11_only(pts_rc, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, clearance, bound, dot_mode, verdict,
                            (cudaStream_t)stream);
}

extern "C" int ppnet_segcheck_mpnet_f32(const float* pts_xy, int64_t n_segs, const int64_t* seg_off,
                                        int64_t segs_per_map, int64_t n_maps, const double* obs,
                                        const int32_t* obs_cnt, int32_t omax, double clearance,
                                        double bound, uint8_t* verdict, uint8_t* steer, void* stream) {
    PPNET_REQUIRE(verdict || steer || n_segs == 0, "segcheck: both outputs are null");
    int64_t dummy = 1;
    int rc = check_common(pts_xy, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, &dummy);
    if (rc != PPNET_OK || dummy == 0) return rc;
    // more than one circle tile: later tiles continue from the stored verdict bytes
    PPNET_REQUIRE(verdict || omax <= kCircTile, "segcheck: verdict may only be null when omax <= 128");
    return verdict_a12_only(pts_xy, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, clearance, bound, PPNET_CMP_F32_NEP50,
                            verdict, steer, nullptr, (cudaStream_t)stream);
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
