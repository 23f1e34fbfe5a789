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
#include "segcheck.cuh"

namespace ppnet {

// ---- kernel -----------------------------------------------------------------------------------------------
// grid = (n_maps, chunks_per_map_max).  One CTA owns `chunk` consecutive segments of one map.
//  1. staging (once per CTA): circles -> decision form; per axis two prefix tables over the 32 bins,
//     LT[b] = circles whose box ends before bin b, GT[b] = circles whose box starts after bin b, so the set
//     of circles meeting a bin range [b0, b1] is ~(LT[b0] | GT[b1]): four 128-bit loads per segment piece.
//  2. each warp takes batches of 32 segments, one per lane: setup + candidate mask (converged code);
//  3. the (segment, candidate circle) pairs of the whole batch go through a per-warp shared-memory queue and
//     are evaluated 32 at a time whatever their owner: a lane with 9 candidates no longer holds back 31 lanes
//     with one.  Hits are OR-ed into a per-warp mask; a pair whose segment already hit is skipped.
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kQueueCap = 512;                   // pairs per round (16 per lane)
constexpr int kTakeMax = kQueueCap / 32;

// resident CTAs per SM the compiler must allow: f64 needs its 80 registers (spills cost more than occupancy gives),
// f32 is best at 72 registers and 7 CTAs (measured: 5 / 6 / 7 / 8 / 9 / 10 CTAs -> 0.319 / 0.290 / 0.273 / 0.279 / 0.297 / 0.314 ms)
template <typename T> struct SegOcc;
template <> struct SegOcc<double> { static constexpr int kMinBlocks = 6; };
template <> struct SegOcc<float> { static constexpr int kMinBlocks = 7; };

template <typename T>
struct __align__(16) SegSlot {                   // what a pair needs of its segment (d, |d|^2 are recomputed)
    T s0, s1, e0, e1, L, es;                     // L == 0 marks a verbatim-only segment
};

template <typename T, int MODE, bool SWAP, bool STEER>
__global__ void __launch_bounds__(kSegThreads, SegOcc<T>::kMinBlocks)
segcheck_kernel(const T* __restrict__ pts, const int64_t* __restrict__ seg_off, int64_t segs_per_map, int chunk,
                const double* __restrict__ obs, const int32_t* __restrict__ obs_cnt, int omax,
                double clearance, T bound, uint8_t* __restrict__ verdict, uint8_t* __restrict__ steer) {
    const int m = blockIdx.x;
    const int64_t lo = seg_off ? seg_off[m] : (int64_t)m * segs_per_map;
    const int64_t hi = seg_off ? seg_off[m + 1] : lo + segs_per_map;
    const int64_t base = lo + (int64_t)blockIdx.y * chunk;
    if (base >= hi) return;                       // whole CTA exits together
    const int64_t end = min(hi, base + (int64_t)chunk);

    __shared__ Circle<T> sc[kCircTile];
    __shared__ T sem[kCircTile];                  // eps * (|o|_1 + thr)
    __shared__ uint4 edge_lo[2][kBins], edge_hi[2][kBins];   // circles whose box starts / ends in this bin (x, y)
    __shared__ uint4 LT[2][kBins], GT[2][kBins];
    __shared__ uint32_t live_mask[4];             // circles that can ever answer "hit" (thr > 0)
    __shared__ uint32_t odd_mask[4];              // ... of those, the ones that always take the verbatim path
    __shared__ SegSlot<T> slot[kSegWarps][32];
    __shared__ uint16_t queue[kSegWarps][kQueueCap];
    __shared__ uint32_t hitmask[kSegWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cnt = min(obs_cnt[m], omax);
    const bool use_grid = bound > T(0) && bound < T(1e6);
    const float bscale = use_grid ? (float)kBins / (float)bound : 0.0f;
    const double* __restrict__ mobs = obs + (size_t)m * omax * 3;

    // first segment of this lane: its load is in flight while the circles are staged
    int64_t i = base + 32 * warp + lane;
    T a0 = T(0), a1 = T(0), b0 = T(0), b1 = T(0);
    if (i < end) Vec4<T>::load(pts + 4 * i, a0, a1, b0, b1);

    for (int t0 = 0; t0 == 0 || t0 < cnt; t0 += kCircTile) {
        const int nt = max(0, min(kCircTile, cnt - t0));
        const bool first_tile = t0 == 0, last_tile = t0 + kCircTile >= cnt;
        __syncthreads();
        for (int t = threadIdx.x; t < 4 * kBins; t += kSegThreads) {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            (t < 2 * kBins ? &edge_lo[0][0] : &edge_hi[0][0] - 2 * kBins)[t] = z;
        }
        if (threadIdx.x < 4) { live_mask[threadIdx.x] = 0u; odd_mask[threadIdx.x] = 0u; }
        __syncthreads();
        if (threadIdx.x < nt) {
            const int j = threadIdx.x;
            const Circle<T> c = make_circle<T>(mobs + 3 * (t0 + j), clearance);
            sc[j] = c;
            const uint32_t bit = 1u << (j & 31);
            const int w = j >> 5;
            if (c.thr > T(0)) {                   // thr <= 0 or NaN: neither test can ever be true
                atomicOr(&live_mask[w], bit);
                const T mc = FP<T>::abs_(c.ox) + FP<T>::abs_(c.oy) + c.thr;
                const T em = Filt<T>::eps * mc;
                sem[j] = em;
                if (!(mc < Filt<T>::lim)) {
                    atomicOr(&odd_mask[w], bit);  // NaN / inf / huge: never culled, always verbatim
                } else if (use_grid) {
                    const T h = c.thr + em;
                    const int x0 = bin_clamp((float)(c.ox - h) * bscale - 2e-3f), x1 = bin_clamp((float)(c.ox + h) * bscale + 2e-3f);
                    const int y0 = bin_clamp((float)(c.oy - h) * bscale - 2e-3f), y1 = bin_clamp((float)(c.oy + h) * bscale + 2e-3f);
                    atomicOr(reinterpret_cast<uint32_t*>(&edge_lo[0][x0]) + w, bit);
                    atomicOr(reinterpret_cast<uint32_t*>(&edge_hi[0][x1]) + w, bit);
                    atomicOr(reinterpret_cast<uint32_t*>(&edge_lo[1][y0]) + w, bit);
                    atomicOr(reinterpret_cast<uint32_t*>(&edge_hi[1][y1]) + w, bit);
                }
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 4 * kBins; t += kSegThreads) {   // prefix tables: one (which, axis, bin) entry per thread
            const int bin = t % kBins, axis = (t / kBins) & 1, which = t / (2 * kBins);
            uint4 acc = make_uint4(0u, 0u, 0u, 0u);
            if (which == 0) {
                for (int b = 0; b < bin; ++b) { const uint4 t = edge_hi[axis][b]; acc.x |= t.x; acc.y |= t.y; acc.z |= t.z; acc.w |= t.w; }
                LT[axis][bin] = acc;
            } else {
                for (int b = bin + 1; b < kBins; ++b) { const uint4 t = edge_lo[axis][b]; acc.x |= t.x; acc.y |= t.y; acc.z |= t.z; acc.w |= t.w; }
                GT[axis][bin] = acc;
            }
        }
        __syncthreads();
        const uint32_t lv0 = live_mask[0], lv1 = live_mask[1], lv2 = live_mask[2], lv3 = live_mask[3];
        const uint32_t od0 = odd_mask[0], od1 = odd_mask[1], od2 = odd_mask[2], od3 = odd_mask[3];
        if (!first_tile) { i = base + 32 * warp + lane; if (i < end) Vec4<T>::load(pts + 4 * i, a0, a1, b0, b1); }

        // warp-uniform loop over this warp's batches of 32 segments
        for (int64_t batch = base + 32 * warp; batch < end; batch += kSegThreads) {
            const bool have = i < end;
            const FastSeg<T> q = fast_setup<T, SWAP>(a0, a1, b0, b1, bound);
            const int64_t cur = i;
            i += kSegThreads;
            if (i < end) Vec4<T>::load(pts + 4 * i, a0, a1, b0, b1);      // next batch's load overlaps this one's work
            bool hit = !have || (first_tile ? q.oob : (verdict[cur] != 0));   // later tiles continue from the stored verdict
            uint32_t c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
            if (!hit && nt > 0) {
                if (q.verbatim || !use_grid) {
                    c0 = lv0; c1 = lv1; c2 = lv2; c3 = lv3;
                } else {
                    // pieces <= 3 bins long (one box for most short segments); each box is inflated by the margin + float slop
                    const float sx = (float)q.s0 * bscale, sy = (float)q.s1 * bscale;
                    const float dx = (float)q.d0 * bscale, dy = (float)q.d1 * bscale;
                    const float mb = (float)q.es * bscale + 2e-3f;
                    const int np_ = min(16, 1 + (int)(fmaxf(fabsf(dx), fabsf(dy)) * (1.0f / 12.0f)));
                    const float inv = 1.0f / (float)np_;
                    uint32_t n0 = ~0u, n1 = ~0u, n2 = ~0u, n3 = ~0u;       // circles culled by EVERY piece
                    for (int pc = 0; pc < np_; ++pc) {
                        const float ta = (float)pc * inv, tb = (float)(pc + 1) * inv;
                        const float ax = sx + dx * ta, bx = sx + dx * tb, ay = sy + dy * ta, by = sy + dy * tb;
                        const int x0 = bin_clamp(fminf(ax, bx) - mb), x1 = bin_clamp(fmaxf(ax, bx) + mb);
                        const int y0 = bin_clamp(fminf(ay, by) - mb), y1 = bin_clamp(fmaxf(ay, by) + mb);
                        const uint4 p = LT[0][x0], r = GT[0][x1], u = LT[1][y0], v = GT[1][y1];
                        n0 &= p.x | r.x | u.x | v.x; n1 &= p.y | r.y | u.y | v.y;
                        n2 &= p.z | r.z | u.z | v.z; n3 &= p.w | r.w | u.w | v.w;
                    }
                    c0 = lv0 & (~n0 | od0); c1 = lv1 & (~n1 | od1); c2 = lv2 & (~n2 | od2); c3 = lv3 & (~n3 | od3);
                }
            }
            {
                SegSlot<T> sl;
                sl.s0 = q.s0; sl.s1 = q.s1; sl.e0 = q.e0; sl.e1 = q.e1; sl.es = q.es;
                sl.L = q.verbatim ? T(0) : q.L;
                slot[warp][lane] = sl;
            }
            if (lane == 0) hitmask[warp] = 0u;
            // rounds of <= kQueueCap pairs (one round unless some lane has > 16 candidates)
            while (__any_sync(0xffffffffu, (c0 | c1 | c2 | c3) != 0u)) {
                const int mine = min(kTakeMax, __popc(c0) + __popc(c1) + __popc(c2) + __popc(c3));
                int off = mine;                                            // inclusive warp scan
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, off, sft);
                    if (lane >= sft) off += t;
                }
                const int total = __shfl_sync(0xffffffffu, off, 31);
                off -= mine;
                __syncwarp();
                {   // one short loop per mask word (simple body) instead of one loop choosing the word every time
                    uint16_t* qp = &queue[warp][off];
                    const uint16_t tag = (uint16_t)(lane << 8);
                    int room = mine;
#define PPNET_DRAIN(cw, basej)                                                                  \
                    for (int t = min(room, __popc(cw)); t > 0; --t, --room) {                   \
                        const int b = __ffs(cw) - 1;                                            \
                        cw &= cw - 1;                                                           \
                        *qp++ = (uint16_t)(tag | (basej + b));                                  \
                    }
                    PPNET_DRAIN(c0, 0)
                    PPNET_DRAIN(c1, 32)
                    PPNET_DRAIN(c2, 64)
                    PPNET_DRAIN(c3, 96)
#undef PPNET_DRAIN
                }
                __syncwarp();
                for (int k = lane; k < total; k += 32) {
                    const int e = queue[warp][k];
                    const int owner = e >> 8, j = e & 127;
                    if ((*reinterpret_cast<volatile uint32_t*>(&hitmask[warp]) >> owner) & 1u) continue;
                    const SegSlot<T> sl = slot[warp][owner];
                    FastSeg<T> g;
                    g.s0 = sl.s0; g.s1 = sl.s1; g.e0 = sl.e0; g.e1 = sl.e1; g.L = sl.L; g.es = sl.es;
                    g.d0 = FP<T>::sub(g.e0, g.s0);
                    g.d1 = FP<T>::sub(g.e1, g.s1);
                    g.L2 = g.d0 * g.d0 + g.d1 * g.d1;
                    const bool odd = ((j < 64 ? (j < 32 ? od0 : od1) : (j < 96 ? od2 : od3)) >> (j & 31)) & 1u;
                    if (fast_pair<T, MODE>(g, sc[j], sem[j], !(sl.L > T(0)) || odd)) atomicOr(&hitmask[warp], 1u << owner);
                }
                __syncwarp();
            }
            __syncwarp();
            if (have) {
                hit = hit || ((hitmask[warp] >> lane) & 1u);
                if (verdict && (first_tile || hit)) verdict[cur] = hit ? 1 : 0;     // later tiles read it back
                if (STEER && steer && last_tile) {
                    // steerTo: dist = euclidean(start, end) in f32 (un-fused); 0 iff dist > 0 and blocked.
                    // sqrt(x) > 0  <=>  x > 0  (and NaN stays false), so the root itself is not needed.
                    using F = FP<T>;
                    const T x = F::sub(q.s0, q.e0), y = F::sub(q.s1, q.e1);
                    const T d2 = F::add(F::mul(x, x), F::mul(y, y));
                    steer[cur] = (d2 > T(0) && hit) ? 0 : 1;
                }
            }
            __syncwarp();
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

// largest per-map segment count decides grid.y; with a CSR we cannot know it without a device
// read, so the caller passes max_segs_per_map through segs_per_map when seg_off != NULL.
template <typename T, bool SWAP, bool STEER>
static int launch(const T* pts, int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map,
                  int64_t n_maps, const double* obs, const int32_t* obs_cnt, int32_t omax,
                  double clearance, double bound, int32_t dot_mode, uint8_t* verdict, uint8_t* steer,
                  cudaStream_t st) {
    int64_t dummy = 1;
    int rc = check_common(pts, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, &dummy);
    if (rc != PPNET_OK || dummy == 0) return rc;
    // more than one circle tile: later tiles continue from the verdict the earlier ones stored
    PPNET_REQUIRE(verdict || omax <= kCircTile, "segcheck: verdict may only be null when omax <= 128");
    const int64_t per_map = segs_per_map > 0 ? segs_per_map : n_segs;
    // one CTA per map when there are enough maps to fill the machine (staging amortised over the whole map);
    // otherwise cut maps into chunks until ~8 CTAs per SM exist
    int64_t chunk = 8192;
    while (chunk > kSegThreads && n_maps * ((per_map + chunk - 1) / chunk) < 8 * kNumSMs) chunk >>= 1;
    const int64_t chunks = (per_map + chunk - 1) / chunk;
    PPNET_REQUIRE(chunks <= 65535, "segcheck: more than 65535*8192 segments in one map");
    dim3 grid((unsigned)n_maps, (unsigned)chunks);
    if (dot_mode == PPNET_DOT_UNFUSED)
        segcheck_kernel<T, PPNET_DOT_UNFUSED, SWAP, STEER><<<grid, kSegThreads, 0, st>>>(
            pts, seg_off, segs_per_map, (int)chunk, obs, obs_cnt, omax, clearance, (T)bound, verdict, steer);
    else
        segcheck_kernel<T, PPNET_DOT_FUSED_SKX, SWAP, STEER><<<grid, kSegThreads, 0, st>>>(
            pts, seg_off, segs_per_map, (int)chunk, obs, obs_cnt, omax, clearance, (T)bound, verdict, steer);
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
