// Fused segment verdicts: A11 (process_map.collision_check_circle_edge, EDaGe-PP/process_map.py:383-425, float64)
// and A12 (neuralplanner.collision_check_circle_edge, experiments/MPNet/neuralplanner.py:43-69, float32) on ONE read
// of the candidate segments, verdicts returned bit-packed (one ballot word per 32 segments) and/or as bytes.
//
// Input is the A11 layout pts_rc[N][4] = (s_row, s_col, e_row, e_col) float64.  The A12 flavour sees the same
// geometric segment the way a float32 caller would hold it: (x, y) = (float32(col), float32(row)) -- the cast + swap
// that bench.py / the host pipeline used to do on the CPU and upload as a second 16 B/segment array.  Each flavour's
// verdict is bit-identical to running its own kernel (segcheck.cu) on its own array: the per-pair decisions are the
// same `fast_pair` / `edge_exact` code, evaluated per flavour on that flavour's operands; what is shared is everything
// that does not depend on the last bits -- the segment load, the separable bin culling (done once, with the wider
// float32 margins plus the float32 rounding of the endpoints), the per-warp pair queue and the output ballots.
//
// What changed against segcheck.cu's one-flavour kernel (round-1 profile: shared-memory wavefronts were ~90 % of the
// kernel's SM cycles, 4.0x / 3.0x bank-conflict replay on the AoS slot / circle records, 10x on the hit-mask atomics):
//   * slots and circles are SoA; the queue is owner-major, so the 32 pairs of a round read CONSECUTIVE slot words
//     (duplicates broadcast) -- conflict-free by construction; circle words are 8 B each (random j: ~1.5x ideal).
//   * hits are combined with redux.sync (one instruction per round and flavour) into warp-uniform registers instead
//     of shared-memory atomics; "owner already hit" is a register test.
//   * "exact only" circles / segments are encoded in the data (margin = +inf / L = NaN fall through every filter
//     comparison into `edge_exact`), no flag words in the pair loop.
//   * (second half of round 2) an inner-disk grid resolves the segments whose END point lies well inside a circle before
//     any culling or pair (see `inner` below); the pair decisions are predicated instead of branched and both flavours run
//     back to back (`fast_pair_bf`); the approximate |d| and 1 / pieces come from one MUFU each; the batches after the
//     next are prefetched into L2.  DESIGN.md 4.2 has the measurements, including what was tried and dropped.
#include "segcheck.cuh"

#ifndef PPNET_VBRANCHFREE
#define PPNET_VBRANCHFREE 1
#endif

namespace ppnet {

#ifndef PPNET_VTHREADS
#define PPNET_VTHREADS 128
#endif
constexpr int kVThreads = PPNET_VTHREADS;
constexpr int kVWarps = kVThreads / 32;
constexpr int kVQueue = 512;                   // pairs per round of the queue (16 per lane)
constexpr int kVTake = kVQueue / 32;
constexpr int kIG = 64;                        // cells per axis of the inner-disk grid (see `inner` in the kernel)
constexpr int kIGMinSegs = 256;                // CTAs with fewer segments skip the grid
constexpr int kIGRows = 8;                     // (circle, row) fill tasks per circle and sweep

// smallest float >= t (t finite or inf): (double)x < t  <=>  x < ceil_f32(t) for every float x
__device__ __forceinline__ float ceil_f32(double t) {
    float f = (float)t;                                      // round to nearest
    if ((double)f < t) f = __int_as_float(__float_as_int(f) + (f >= 0.0f ? 1 : -1));
    return f;
}

// one MUFU each; relative error <= 2^-22: the margins that absorb the approximate |d| (eps64 = 4e-6, eps32 = 5e-5, see
// segcheck.cuh) carry 10x over it, and the piece boundaries of the culling query carry 2e-3 bins of slop
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// pair test of one flavour on that flavour's operands (same code path as segcheck.cu's kernel)
template <typename T, int MODE>
__device__ __forceinline__ bool flavour_pair(T s0, T s1, T e0, T e1, T L, T es, T ox, T oy, T thr, T T2, T em) {
#if PPNET_VBRANCHFREE
    // L = NaN (verbatim-only segment) or em = +inf (odd circle) make every filter comparison false -> edge_exact
    return fast_pair_bf<T, MODE>(s0, s1, e0, e1, L, es, ox, oy, thr, T2, em);
#else
    FastSeg<T> g;
    g.s0 = s0; g.s1 = s1; g.e0 = e0; g.e1 = e1; g.L = L; g.es = es;
    g.d0 = FP<T>::sub(e0, s0);
    g.d1 = FP<T>::sub(e1, s1);
    g.L2 = g.d0 * g.d0 + g.d1 * g.d1;
    Circle<T> c;
    c.ox = ox; c.oy = oy; c.thr = thr; c.T2 = T2;
    return fast_pair<T, MODE>(g, c, em, false);
#endif
}

#ifndef PPNET_VPF2
#define PPNET_VPF2 2
#endif
template <typename TIN> struct VOcc;
#ifndef PPNET_VPREFETCH
#define PPNET_VPREFETCH 1
#endif
#ifndef PPNET_VOCC
#define PPNET_VOCC 6
#endif
template <> struct VOcc<double> { static constexpr int kMinBlocks = PPNET_VOCC; };
template <> struct VOcc<float> { static constexpr int kMinBlocks = 7; };

// bit-packed output helpers: bit (i & 31) of word (i >> 5) belongs to global segment i.  `i0` = index of lane 0.
__device__ __forceinline__ uint32_t vbits_load(const uint32_t* w, int64_t n_words, int64_t i0) {
    const int64_t k = i0 >> 5;
    const int sh = (int)(i0 & 31);
    uint32_t v = w[k] >> sh;
    if (sh && k + 1 < n_words) v |= w[k + 1] << (32 - sh);
    return v;
}
__device__ __forceinline__ void vbits_or(uint32_t* w, int64_t i0, uint32_t m, bool exclusive, bool first) {
    const int64_t k = i0 >> 5;
    const int sh = (int)(i0 & 31);
    if (exclusive) {                                         // this warp owns the whole word
        PPNET_ASSERT(sh == 0);
        if (first) w[k] = m;
        else if (m) w[k] |= m;
        return;
    }
    if (m << sh) atomicOr(w + k, m << sh);                   // words were zeroed by the launcher
    if (sh && (m >> (32 - sh))) atomicOr(w + k + 1, m >> (32 - sh));
}

// TIN = double: pts = (s_row, s_col, e_row, e_col), flavours DO64 (A11) and/or DO32 (A12 on the cast + swap);
// TIN = float:  pts = (s_x, s_y, e_x, e_y), DO32 only (A12 as the reference receives it).
template <int MODE, bool DO64, bool DO32, typename TIN>
__global__ void __launch_bounds__(kVThreads, VOcc<TIN>::kMinBlocks)
verdict_kernel(const TIN* __restrict__ pts, const int64_t* __restrict__ seg_off, int64_t segs_per_map, int chunk,
               const double* __restrict__ obs, const int32_t* __restrict__ obs_cnt, int omax, double clearance,
               TIN bound, int cmp64, uint8_t* __restrict__ v64, uint8_t* __restrict__ v32, uint32_t* __restrict__ b64,
               uint32_t* __restrict__ b32, uint8_t* __restrict__ steer, int exclusive_words, int64_t n_words) {
    const int m = blockIdx.x;
    const int64_t lo = seg_off ? seg_off[m] : (int64_t)m * segs_per_map;
    const int64_t hi = seg_off ? seg_off[m + 1] : lo + segs_per_map;
    const int64_t base = lo + (int64_t)blockIdx.y * chunk;
    if (base >= hi) return;                       // whole CTA exits together
    const int64_t end = min(hi, base + (int64_t)chunk);

    // circles, SoA (decision form; see segcheck.cu)
    __shared__ float2 c_oxy[kCircTile];           // centre, exactly float32 in both flavours
    __shared__ double c_thr64[DO64 ? kCircTile : 1], c_T64[DO64 ? kCircTile : 1];
    __shared__ float2 c_t32[DO32 ? kCircTile : 1];   // (thr, T) of the float32 flavour
    __shared__ float2 c_em[kCircTile];            // (eps64 * mc, eps32 * mc), +inf = exact only
    __shared__ uint4 edge_lo[2][kBins], edge_hi[2][kBins];
    __shared__ uint4 LT[2][kBins], GT[2][kBins];
    __shared__ __align__(16) uint32_t live_mask[4], odd_mask[4];
    // per-warp slots, SoA: what a pair needs of its segment
    __shared__ TIN sl_s0[kVWarps][32], sl_s1[kVWarps][32], sl_e0[kVWarps][32], sl_e1[kVWarps][32];
    __shared__ float2 sl_a[kVWarps][32];          // (L, es) of the float64 flavour, as floats (NaN L = verbatim)
    __shared__ float2 sl_b[kVWarps][32];          // (L, es) of the float32 flavour
    __shared__ uint16_t queue[kVWarps][kVQueue];
    // Inner-disk grid: bit (iy, ix) set = the whole cell [ix, ix+1] x [iy, iy+1] * bound/kIG lies inside the disk
    // |p - o| <= thr - mg of some circle of the tile.  The reference returns True as soon as ONE circle passes the vertex
    // test euclidean(e, o) < thr (process_map.py:397, neuralplanner.py:54), whatever s is and whatever the other circles do
    // (no earlier iteration can raise), so a segment whose END point falls in a set cell is blocked in both flavours and
    // never enters the pair queue (42 % of the config-2 segments have e inside a disk; the 64 x 64 grid catches 33 %, and
    // they are the segments with the most candidates).  mg = 2 eps32 (2 bound + |o|_1 + thr + 1) + 2e-5 bound covers the
    // reference's rounding of the distance (<= 50 u32 Mg, e in [0, bound)^2), the float32 rounding of e and of thr in
    // the float32 flavour, and the float arithmetic of the cell index / the fill below (each <= 1e-6 bound).
    __shared__ uint32_t inner[kIG][kIG / 32];
    __shared__ float c_rin[kCircTile];            // thr - mg, <= 0: circle takes no part (dead, odd, too small, no grid)
    __shared__ int inner_any;                     // some circle of the tile marks cells: segments look their end point up
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cnt = min(obs_cnt[m], omax);
    const bool use_grid = bound > TIN(0) && bound < TIN(1e6);
    const float bscale = use_grid ? (float)kBins / (float)bound : 0.0f;
#ifndef PPNET_VINNER
#define PPNET_VINNER 1
#endif
    // (bound > 1e-6 keeps every product below finite; a CTA with few segments would not amortise the fill)
    const bool use_inner = PPNET_VINNER != 0 && use_grid && (float)bound > 1e-6f && (end - base) >= kIGMinSegs;
    const float gs = use_inner ? (float)kIG / (float)bound : 0.0f, cw = (float)bound / (float)kIG;
    const double* __restrict__ mobs = obs + (size_t)m * omax * 3;

    // CTA-relative 32-bit indices from here on (chunk <= 8192)
    const int n_here = (int)(end - base);
    pts += 4 * base;
    if (v64) v64 += base;
    if (v32) v32 += base;
    if (steer) steer += base;
    int i = 32 * warp + lane;
    TIN a0 = TIN(0), a1 = TIN(0), b0 = TIN(0), b1 = TIN(0);
#if PPNET_VPREFETCH
    if (i < n_here) Vec4<TIN>::load(pts + 4 * i, a0, a1, b0, b1);
#endif
#if PPNET_VPF2
    for (int k = 1; k <= PPNET_VPF2; ++k)
        if (i + k * kVThreads < n_here) asm volatile("prefetch.global.L2 [%0];" ::"l"(pts + 4 * (i + k * kVThreads)));
#endif

    for (int t0 = 0; t0 == 0 || t0 < cnt; t0 += kCircTile) {
        const int nt = max(0, min(kCircTile, cnt - t0));
        const bool first_tile = t0 == 0, last_tile = t0 + kCircTile >= cnt;
        const bool wide = nt > 64;                                 // live / candidate words 2, 3 are zero otherwise
        __syncthreads();
        for (int t = threadIdx.x; t < 4 * kBins; t += kVThreads) {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            (t < 2 * kBins ? &edge_lo[0][0] : &edge_hi[0][0] - 2 * kBins)[t] = z;
        }
        if (threadIdx.x < 4) { live_mask[threadIdx.x] = 0u; odd_mask[threadIdx.x] = 0u; }
        if (threadIdx.x == 0) inner_any = 0;
        for (int t = threadIdx.x; t < kIG * kIG / 32; t += kVThreads) (&inner[0][0])[t] = 0u;
        __syncthreads();
        for (int j = threadIdx.x; j < nt; j += kVThreads) {
            const double* o = mobs + 3 * (t0 + j);
            const float oxf = (float)o[0], oyf = (float)o[1];            // torch.tensor([ox, oy]) -> float32 centre
            const double thr = __dadd_rn(o[2], __dmul_rn(clearance, 0.5));   // size + clearance/2 (Python floats; x / 2 == x * 0.5 exactly)
            // float32 flavour: NumPy >= 2 compares the float32 offset with float32(thr) (NEP 50); NumPy 1.x promotes
            // the offset to float64, which for float32 x is  x < ceil_f32(thr)
            const float thrf = cmp64 ? ceil_f32(thr) : (float)thr;
            c_oxy[j] = make_float2(oxf, oyf);
            if (DO64) { c_thr64[j] = thr; c_T64[j] = sqrt_lt_threshold(thr); }
            if (DO32) c_t32[j] = make_float2(thrf, sqrt_lt_threshold(thrf));
            const uint32_t bit = 1u << (j & 31);
            const int w = j >> 5;
            const bool live = DO64 ? thr > 0.0 : thrf > 0.0f;            // thrf > 0 implies thr > 0
            float2 em = make_float2(CUDART_INF_F, CUDART_INF_F);
            float rin = -1.0f;
            if (live) {
                atomicOr(&live_mask[w], bit);
                const double mc = fabs((double)oxf) + fabs((double)oyf) + fabs(thr);
                if (mc < Filt<double>::lim) em.x = (float)(Filt<double>::eps * mc);
                if (mc < (double)Filt<float>::lim) em.y = Filt<float>::eps * (float)mc;
                const bool odd = DO32 ? !(mc < (double)Filt<float>::lim) : !(mc < Filt<double>::lim);
                if (odd) {
                    atomicOr(&odd_mask[w], bit);                         // never culled, always verbatim
                } else if (use_grid) {
                    // box of the wider (float32) margin, + the float32 rounding of the threshold itself
                    const float h = (float)thr * 1.000001f + (DO32 ? em.y : em.x);
                    const int x0 = bin_clamp((oxf - h) * bscale - 2e-3f), x1 = bin_clamp((oxf + h) * bscale + 2e-3f);
                    const int y0 = bin_clamp((oyf - h) * bscale - 2e-3f), y1 = bin_clamp((oyf + h) * bscale + 2e-3f);
                    atomicOr(reinterpret_cast<uint32_t*>(&edge_lo[0][x0]) + w, bit);
                    atomicOr(reinterpret_cast<uint32_t*>(&edge_hi[0][x1]) + w, bit);
                    atomicOr(reinterpret_cast<uint32_t*>(&edge_lo[1][y0]) + w, bit);
                    atomicOr(reinterpret_cast<uint32_t*>(&edge_hi[1][y1]) + w, bit);
                    if (use_inner) {
                        rin = (float)thr * 0.999999f -
                              (2.0f * Filt<float>::eps * (2.0f * (float)bound + (float)mc + 1.0f) + 2e-5f * (float)bound);
                        if (rin > 0.70711f * cw) inner_any = 1;    // a disk of radius < cw / sqrt(2) cannot hold a cell
                        else rin = -1.0f;
                    }
                }
            }
            c_em[j] = em;
            c_rin[j] = rin;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 4 * kBins; t += kVThreads) {   // prefix tables: one (which, axis, bin) entry per thread
            const int bin = t % kBins, axis = (t / kBins) & 1, which = t / (2 * kBins);
            uint4 acc = make_uint4(0u, 0u, 0u, 0u);
            if (which == 0) {
                for (int b = 0; b < bin; ++b) { const uint4 q = edge_hi[axis][b]; acc.x |= q.x; acc.y |= q.y; acc.z |= q.z; acc.w |= q.w; }
                LT[axis][bin] = acc;
            } else {
                for (int b = bin + 1; b < kBins; ++b) { const uint4 q = edge_lo[axis][b]; acc.x |= q.x; acc.y |= q.y; acc.z |= q.z; acc.w |= q.w; }
                GT[axis][bin] = acc;
            }
        }
        if (use_inner && inner_any) {
            // inner-disk grid: kIGRows (circle, cell row) tasks per circle and sweep; every rounding shrinks the span
            static_assert(kIG == 64, "one 64-bit span mask per row");
            for (int t = threadIdx.x; t < kIGRows * nt; t += kVThreads) {
                const int j = t / kIGRows;
                const float rin = c_rin[j];
                if (!(rin > 0.0f)) continue;
                const float2 oc = c_oxy[j];
                const int iy_lo = (int)fmaxf(floorf((oc.y - rin) * gs), 0.0f);
                const int iy_hi = (int)fminf(floorf((oc.y + rin) * gs), (float)(kIG - 1));
                for (int iy = iy_lo + t % kIGRows; iy <= iy_hi; iy += kIGRows) {
                    const float yl = (float)iy * cw;
                    const float fy = fmaxf(fabsf(yl - oc.y), fabsf(yl + cw - oc.y));   // farthest y of the row from the centre
                    const float w2 = rin * rin - fy * fy;
                    if (!(w2 > 0.0f)) continue;
                    const float hw = sqrt_approx(w2) * 0.999999f;                    // (2 ulp; the factor is 8 ulp)
                    const int ix0 = max((int)ceilf((oc.x - hw) * gs + 1e-3f), 0);
                    const int ix1 = min((int)floorf((oc.x + hw) * gs - 1e-3f) - 1, kIG - 1);
                    if (ix0 > ix1) continue;
                    PPNET_ASSERT(iy >= 0 && iy < kIG && ix0 >= 0 && ix1 < kIG);
                    const unsigned long long span = ((2ull << ix1) - 1ull) & ~((1ull << ix0) - 1ull);   // bits ix0..ix1
                    if ((uint32_t)span) atomicOr(&inner[iy][0], (uint32_t)span);
                    if ((uint32_t)(span >> 32)) atomicOr(&inner[iy][1], (uint32_t)(span >> 32));
                }
            }
        }
        __syncthreads();
#if PPNET_VPREFETCH
        if (!first_tile) { i = 32 * warp + lane; if (i < n_here) Vec4<TIN>::load(pts + 4 * i, a0, a1, b0, b1); }
#else
        i = 32 * warp + lane;
#endif

        const bool look_inner = use_inner && inner_any != 0;
        for (int batch = 32 * warp; batch < n_here; batch += kVThreads) {
            const bool have = i < n_here;
            const int cur = i;
#if !PPNET_VPREFETCH
            if (have) Vec4<TIN>::load(pts + 4 * i, a0, a1, b0, b1);
#endif
            // ---- per-segment setup (both flavours), then the next batch's load goes out
            bool oob64 = false, oob32 = false, verb64 = false, verb32 = false;
            TIN s0, s1, e0, e1;                                   // (x, y) of the input precision
            if (DO64) {                                           // TIN = double, (row, col) -> swap
                oob64 = (a0 < TIN(0)) || (a1 > bound) || (b0 < TIN(0)) || (b1 > bound);   // process_map.py:384-387, raw (r, c)
                s0 = a1; s1 = a0; e0 = b1; e1 = b0;
            } else if (sizeof(TIN) == 8) {
                s0 = a1; s1 = a0; e0 = b1; e1 = b0;
            } else {
                s0 = a0; s1 = a1; e0 = b0; e1 = b1;
            }
            float L64 = 0.f, es64 = 0.f, L32 = 0.f, es32 = 0.f;
            float fs0 = (float)s0, fs1 = (float)s1, fe0 = (float)e0, fe1 = (float)e1;
            // |s|_1 + |e|_1 + 1 in float serves both flavours' margins (they carry 10x slack; float overflow -> inf -> verbatim)
            const float msf = fabsf(fs0) + fabsf(fs1) + fabsf(fe0) + fabsf(fe1) + 1.0f;
            if (DO64) {
                // d = e - s rounded once in double (as the reference does), everything after it in float: the approximate
                // |d| (relative error ~2e-7, eps64 = 4e-6 is 20x over it) and the margin
                const float fd0 = (float)__dsub_rn((double)e0, (double)s0), fd1 = (float)__dsub_rn((double)e1, (double)s1);
                const float L2 = fd0 * fd0 + fd1 * fd1;
                verb64 = !(msf < 1e15f) || !(L2 > 1e-30f);
                L64 = verb64 ? CUDART_NAN_F : sqrt_approx(L2);
                es64 = (float)Filt<double>::eps * 1.000001f * msf;
            }
            if (DO32) {
                // neuralplanner.py:44-47 on (x, y): s[0] < 0 or s[1] > 224 ...
                oob32 = (fs0 < 0.0f) || (fs1 > (float)bound) || (fe0 < 0.0f) || (fe1 > (float)bound);
                const float d0 = __fsub_rn(fe0, fs0), d1 = __fsub_rn(fe1, fs1);
                const float L2 = d0 * d0 + d1 * d1;
                verb32 = !(msf < Filt<float>::lim) || !(L2 > Filt<float>::tiny);
                L32 = verb32 ? CUDART_NAN_F : sqrt_approx(L2);
                es32 = Filt<float>::eps * msf;
            }
            i += kVThreads;
#if PPNET_VPREFETCH
            if (i < n_here) Vec4<TIN>::load(pts + 4 * i, a0, a1, b0, b1);     // next batch's load overlaps this one's work
#endif
#if PPNET_VPF2
            // ... and the batches after it are pulled into L2 (no registers): a batch is short since the inner-disk grid decides a
            // third of the segments up front, and one batch of work no longer covers the DRAM latency
            if (i + PPNET_VPF2 * kVThreads < n_here)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pts + 4 * (i + PPNET_VPF2 * kVThreads)));
#endif
            // state of this segment so far (later circle tiles continue from the stored verdict)
            bool hit64 = oob64, hit32 = oob32;
            if (!first_tile) {
                uint32_t p64 = 0u, p32 = 0u;
                if (DO64) p64 = v64 ? (have && v64[cur] != 0) : ((vbits_load(b64, n_words, base + batch) >> lane) & 1u);
                if (DO32) p32 = v32 ? (have && v32[cur] != 0) : ((vbits_load(b32, n_words, base + batch) >> lane) & 1u);
                hit64 = p64 != 0u; hit32 = p32 != 0u;
            }
            if (look_inner) {                                      // end point inside a disk with margin: blocked, both flavours
                const float gx = fe0 * gs, gy = fe1 * gs;
                if (gx >= 0.0f && gx < (float)kIG && gy >= 0.0f && gy < (float)kIG) {   // (false for NaN)
                    const int ix = (int)gx, iy = (int)gy;
                    PPNET_ASSERT(ix >= 0 && ix < kIG && iy >= 0 && iy < kIG);
                    if (PPNET_VINNER == 1 && (inner[iy][ix >> 5] >> (ix & 31)) & 1u) hit64 = hit32 = true;
                }
            }
            const bool open = have && ((DO64 && !hit64) || (DO32 && !hit32));
            uint32_t c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
            if (open && nt > 0) {
                const uint4 lv = *reinterpret_cast<const uint4*>(live_mask);     // broadcast reads
                if (verb64 || verb32 || !use_grid) {
                    c0 = lv.x; c1 = lv.y; c2 = lv.z; c3 = lv.w;
                } else {
                    // one query for both flavours: pieces <= 3 bins long, each box inflated by the float32 margin, the
                    // float32 rounding of the endpoints (<= 2^-24 |coord|) and float slop
                    const float sx = fs0 * bscale, sy = fs1 * bscale;
                    const float dx = (fe0 - fs0) * bscale, dy = (fe1 - fs1) * bscale;
                    const float mb = (DO32 ? es32 * 1.01f : es64) * bscale + 2e-3f;
                    const int np_ = min(16, 1 + (int)(fmaxf(fabsf(dx), fabsf(dy)) * (1.0f / 12.0f)));
                    const float inv = rcp_approx((float)np_);
                    uint32_t n0 = ~0u, n1 = ~0u, n2 = ~0u, n3 = ~0u;       // circles culled by EVERY piece
                    if (wide) {
                        for (int pc = 0; pc < np_; ++pc) {
                            const float ta = (float)pc * inv, tb = (float)(pc + 1) * inv;
                            const float ax = sx + dx * ta, bx = sx + dx * tb, ay = sy + dy * ta, by = sy + dy * tb;
                            const int x0 = bin_clamp(fminf(ax, bx) - mb), x1 = bin_clamp(fmaxf(ax, bx) + mb);
                            const int y0 = bin_clamp(fminf(ay, by) - mb), y1 = bin_clamp(fmaxf(ay, by) + mb);
                            const uint4 p = LT[0][x0], r = GT[0][x1], u = LT[1][y0], v = GT[1][y1];
                            n0 &= p.x | r.x | u.x | v.x; n1 &= p.y | r.y | u.y | v.y;
                            n2 &= p.z | r.z | u.z | v.z; n3 &= p.w | r.w | u.w | v.w;
                        }
                    } else {                                               // <= 64 circles in this tile: two words, 8-byte loads
                        for (int pc = 0; pc < np_; ++pc) {
                            const float ta = (float)pc * inv, tb = (float)(pc + 1) * inv;
                            const float ax = sx + dx * ta, bx = sx + dx * tb, ay = sy + dy * ta, by = sy + dy * tb;
                            const int x0 = bin_clamp(fminf(ax, bx) - mb), x1 = bin_clamp(fmaxf(ax, bx) + mb);
                            const int y0 = bin_clamp(fminf(ay, by) - mb), y1 = bin_clamp(fmaxf(ay, by) + mb);
                            const uint2 p = *reinterpret_cast<const uint2*>(&LT[0][x0]), r = *reinterpret_cast<const uint2*>(&GT[0][x1]);
                            const uint2 u = *reinterpret_cast<const uint2*>(&LT[1][y0]), v = *reinterpret_cast<const uint2*>(&GT[1][y1]);
                            n0 &= p.x | r.x | u.x | v.x; n1 &= p.y | r.y | u.y | v.y;
                        }
                    }
                    const uint4 od = *reinterpret_cast<const uint4*>(odd_mask);
                    c0 = lv.x & (~n0 | od.x); c1 = lv.y & (~n1 | od.y); c2 = lv.z & (~n2 | od.z); c3 = lv.w & (~n3 | od.w);
                }
            }
            sl_s0[warp][lane] = s0; sl_s1[warp][lane] = s1; sl_e0[warp][lane] = e0; sl_e1[warp][lane] = e1;
            if (DO64) sl_a[warp][lane] = make_float2(L64, es64);
            if (DO32) sl_b[warp][lane] = make_float2(L32, es32);
            // warp-uniform hit masks (bit = owner lane); segments that are already decided start set
            uint32_t hm64 = DO64 ? __ballot_sync(0xffffffffu, hit64 || !have) : ~0u;
            uint32_t hm32 = DO32 ? __ballot_sync(0xffffffffu, hit32 || !have) : ~0u;
            while (__any_sync(0xffffffffu, (c0 | c1 | c2 | c3) != 0u)) {       // (c2 = c3 = 0 when !wide)
                const int mine = min(kVTake, __popc(c0) + __popc(c1) + (wide ? __popc(c2) + __popc(c3) : 0));
                int off = mine;                                            // inclusive warp scan
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, off, sft);
                    if (lane >= sft) off += t;
                }
                const int total = __shfl_sync(0xffffffffu, off, 31);
                off -= mine;
                __syncwarp();
                {
                    uint16_t* qp = &queue[warp][off];
                    const uint16_t tag = (uint16_t)(lane << 8);
                    int room = mine;
#define PPNET_DRAIN(cw, basej)                                                                  \
                    for (int t = min(room, __popc(cw)); t > 0; --t, --room) {                   \
                        const int b = __ffs(cw) - 1;                                            \
                        cw &= cw - 1;                                                           \
                        PPNET_ASSERT(qp < &queue[warp][0] + kVQueue && basej + b < nt);         \
                        *qp++ = (uint16_t)(tag | (basej + b));                                  \
                    }
                    PPNET_DRAIN(c0, 0)
                    PPNET_DRAIN(c1, 32)
                    if (wide) {
                        PPNET_DRAIN(c2, 64)
                        PPNET_DRAIN(c3, 96)
                    }
#undef PPNET_DRAIN
                }
                __syncwarp();
                for (int k0 = 0; k0 < total; k0 += 32) {                   // uniform trip count: redux below is warp-wide
                    const int k = k0 + lane;
                    uint32_t h64 = 0u, h32 = 0u;
                    if (k < total) {
                        const int e = queue[warp][k];
                        const int owner = e >> 8, j = e & 127;
                        PPNET_ASSERT(total <= kVQueue && owner < 32 && j < nt);
                        const uint32_t obit = 1u << owner;
                        const bool need64 = DO64 && !(hm64 & obit), need32 = DO32 && !(hm32 & obit);
                        if (need64 || need32) {
                            const TIN q_s0 = sl_s0[warp][owner], q_s1 = sl_s1[warp][owner];
                            const TIN q_e0 = sl_e0[warp][owner], q_e1 = sl_e1[warp][owner];
                            const float2 oc = c_oxy[j];
                            const float2 em = c_em[j];
#if PPNET_VBRANCHFREE
                            // both flavours straight through (independent arithmetic, no divergence point between them); a
                            // flavour whose owner is already decided just drops its result
                            bool r64 = false, r32 = false;
                            if (DO64) {
                                const float2 la = sl_a[warp][owner];
                                r64 = flavour_pair<double, MODE>((double)q_s0, (double)q_s1, (double)q_e0, (double)q_e1, (double)la.x,
                                                                 (double)la.y, (double)oc.x, (double)oc.y, c_thr64[DO64 ? j : 0],
                                                                 c_T64[DO64 ? j : 0], (double)em.x);
                            }
                            if (DO32) {
                                const float2 lb = sl_b[warp][owner];
                                const float2 tt = c_t32[DO32 ? j : 0];
                                r32 = flavour_pair<float, 0>((float)q_s0, (float)q_s1, (float)q_e0, (float)q_e1, lb.x, lb.y, oc.x, oc.y,
                                                             tt.x, tt.y, em.y);
                            }
                            if (need64 && r64) h64 = obit;
                            if (need32 && r32) h32 = obit;
#else
                            if (need64) {
                                const float2 la = sl_a[warp][owner];
                                if (flavour_pair<double, MODE>((double)q_s0, (double)q_s1, (double)q_e0, (double)q_e1, (double)la.x,
                                                               (double)la.y, (double)oc.x, (double)oc.y, c_thr64[DO64 ? j : 0],
                                                               c_T64[DO64 ? j : 0], (double)em.x))
                                    h64 = obit;
                            }
                            if (need32) {
                                const float2 lb = sl_b[warp][owner];
                                const float2 tt = c_t32[DO32 ? j : 0];
                                if (flavour_pair<float, 0>((float)q_s0, (float)q_s1, (float)q_e0, (float)q_e1, lb.x, lb.y, oc.x, oc.y,
                                                           tt.x, tt.y, em.y))
                                    h32 = obit;
                            }
#endif
                        }
                    }
                    if (DO64) hm64 |= __reduce_or_sync(0xffffffffu, h64);
                    if (DO32) hm32 |= __reduce_or_sync(0xffffffffu, h32);
                }
                __syncwarp();
            }
            // ---- results: bytes and / or one ballot word per flavour
            const uint32_t valid = __ballot_sync(0xffffffffu, have);
            if (DO64) {                                            // hm64 started from the bounds test / stored verdict
                const bool h = (hm64 >> lane) & 1u;
                if (v64 && have && (first_tile || h)) v64[cur] = h ? 1 : 0;
                if (b64 && lane == 0) vbits_or(b64, base + batch, hm64 & valid, exclusive_words != 0, first_tile);
            }
            if (DO32) {
                const bool h = (hm32 >> lane) & 1u;
                if (v32 && have && (first_tile || h)) v32[cur] = h ? 1 : 0;
                if (steer && have && last_tile) {
                    // steerTo (neuralplanner.py:86-92): dist = euclidean(start, end) in f32 (un-fused); 0 iff dist > 0 and
                    // blocked.  sqrt(x) > 0 <=> x > 0 (NaN stays false), so the root itself is not needed.
                    const float x = __fsub_rn(fs0, fe0), y = __fsub_rn(fs1, fe1);
                    const float d2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
                    steer[cur] = (d2 > 0.0f && h) ? 0 : 1;
                }
                if (b32 && lane == 0) vbits_or(b32, base + batch, hm32 & valid, exclusive_words != 0, first_tile);
            }
            __syncwarp();
        }
    }
}

}  // namespace ppnet

using namespace ppnet;

template <bool DO64, bool DO32, typename TIN>
static int launch_verdict(const TIN* pts, int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map, int64_t n_maps,
                          const double* obs, const int32_t* obs_cnt, int32_t omax, double clearance, double bound,
                          int32_t dot_mode, int32_t cmp_mode, uint8_t* v64, uint8_t* v32, uint32_t* b64, uint32_t* b32,
                          uint8_t* steer, cudaStream_t st) {
    const int64_t per_map = segs_per_map > 0 ? segs_per_map : n_segs;
    int64_t chunk = 8192;
    while (chunk > kVThreads && n_maps * ((per_map + chunk - 1) / chunk) < 8 * kNumSMs) chunk >>= 1;
    const int64_t chunks = (per_map + chunk - 1) / chunk;
    PPNET_REQUIRE(chunks <= 65535, "verdict: more than 65535*8192 segments in one map");
    const int64_t n_words = (n_segs + 31) / 32;
    // every 32-segment batch owns its output word when rows start on multiples of 32; otherwise words are shared
    // between warps / CTAs: zero them here and OR with atomics in the kernel
    const int exclusive = (seg_off == nullptr && segs_per_map % 32 == 0) ? 1 : 0;
    if (!exclusive) {
        if (b64) PPNET_CUDA(cudaMemsetAsync(b64, 0, 4 * (size_t)n_words, st));
        if (b32) PPNET_CUDA(cudaMemsetAsync(b32, 0, 4 * (size_t)n_words, st));
    }
    dim3 grid((unsigned)n_maps, (unsigned)chunks);
    if (dot_mode == PPNET_DOT_UNFUSED)
        verdict_kernel<PPNET_DOT_UNFUSED, DO64, DO32, TIN><<<grid, kVThreads, 0, st>>>(
            pts, seg_off, segs_per_map, (int)chunk, obs, obs_cnt, omax, clearance, (TIN)bound, cmp_mode, v64, v32, b64, b32,
            steer, exclusive, n_words);
    else
        verdict_kernel<PPNET_DOT_FUSED_SKX, DO64, DO32, TIN><<<grid, kVThreads, 0, st>>>(
            pts, seg_off, segs_per_map, (int)chunk, obs, obs_cnt, omax, clearance, (TIN)bound, cmp_mode, v64, v32, b64, b32,
            steer, exclusive, n_words);
    PPNET_LAUNCH_CHECK("verdict_kernel");
    return PPNET_OK;
}

extern "C" int ppnet_verdict_fused(const double* pts_rc, int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map,
                                   int64_t n_maps, const double* obs, const int32_t* obs_cnt, int32_t omax,
                                   double clearance, double bound, int32_t dot_mode, int32_t cmp_mode,
                                   uint8_t* verdict_f64, uint8_t* verdict_f32, uint32_t* vbits_f64, uint32_t* vbits_f32,
                                   void* stream) {
    PPNET_REQUIRE(n_segs >= 0 && n_maps >= 0, "verdict: negative sizes");
    PPNET_REQUIRE(n_maps <= 2147483647LL, "verdict: too many maps for one launch");
    PPNET_REQUIRE(dot_mode == PPNET_DOT_FUSED_SKX || dot_mode == PPNET_DOT_UNFUSED, "verdict: bad dot_mode");
    PPNET_REQUIRE(cmp_mode == PPNET_CMP_F32_NEP50 || cmp_mode == PPNET_CMP_F64_NUMPY1, "verdict: bad cmp_mode");
    if (n_segs == 0 || n_maps == 0) return PPNET_OK;
    PPNET_REQUIRE(pts_rc && obs_cnt, "verdict: null pointer");
    PPNET_REQUIRE(omax >= 0 && (omax == 0 || obs), "verdict: obs is null but omax > 0");
    PPNET_REQUIRE(seg_off ? segs_per_map > 0 : segs_per_map * n_maps == n_segs,
                  "verdict: uniform grouping needs n_segs == n_maps * segs_per_map; a CSR needs the longest row in segs_per_map");
    PPNET_REQUIRE((reinterpret_cast<uintptr_t>(pts_rc) & 15) == 0, "verdict: pts must be 16-byte aligned");
    const bool want64 = verdict_f64 || vbits_f64, want32 = verdict_f32 || vbits_f32;
    PPNET_REQUIRE(want64 || want32, "verdict: every output is null");
    cudaStream_t st = (cudaStream_t)stream;
    if (want64 && want32)
        return launch_verdict<true, true, double>(pts_rc, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, clearance,
                                                  bound, dot_mode, cmp_mode, verdict_f64, verdict_f32, vbits_f64, vbits_f32, nullptr, st);
    if (want64)
        return launch_verdict<true, false, double>(pts_rc, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, clearance,
                                                   bound, dot_mode, cmp_mode, verdict_f64, nullptr, vbits_f64, nullptr, nullptr, st);
    return launch_verdict<false, true, double>(pts_rc, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, clearance,
                                               bound, dot_mode, cmp_mode, nullptr, verdict_f32, nullptr, vbits_f32, nullptr, st);
}

// ---- the one-flavour entry points run the same kernel (one flavour switched off) -------------------------------------
namespace ppnet {
int verdict_a11_only(const double* pts_rc, int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map, int64_t n_maps,
                     const double* obs, const int32_t* obs_cnt, int32_t omax, double clearance, double bound, int32_t dot_mode,
                     uint8_t* verdict, cudaStream_t st) {
    return launch_verdict<true, false, double>(pts_rc, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, clearance, bound,
                                               dot_mode, PPNET_CMP_F32_NEP50, verdict, nullptr, nullptr, nullptr, nullptr, st);
}
int verdict_a12_only(const float* pts_xy, int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map, int64_t n_maps,
                     const double* obs, const int32_t* obs_cnt, int32_t omax, double clearance, double bound, int32_t cmp_mode,
                     uint8_t* verdict, uint8_t* steer, uint8_t* scratch, cudaStream_t st) {
    // more than one circle tile needs the verdict bytes as the state between tiles: `scratch` stands in when the caller
    // only asked for steer
    return launch_verdict<false, true, float>(pts_xy, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, clearance, bound,
                                              PPNET_DOT_FUSED_SKX, cmp_mode, nullptr, verdict ? verdict : scratch, nullptr, nullptr,
                                              steer, st);
}
}  // namespace ppnet
