// Integer DDA grid check against bit-packed occupancy maps (new functionality: the reference has no
// occupancy-grid lookup).  Endpoints are snapped with the A4 rule (Path.coord_euclidean2image,
// EDaGe-PP/Path.py:378-386: rint = round-half-to-even; here step 1, offset 0), then the walk is all-integer:
//     n = max(|dx|, |dy|),  cell_k = (x0 + floor((2 k dx + n) / 2n),  y0 + floor((2 k dy + n) / 2n)),  k = 0..n
// verdict = first k whose cell is outside [0,R)^2 or occupied.
//
// One CTA per (map, chunk).  The map's bitmap (6 272 B at R = 224, 131 072 B at R = 1024) is pulled into
// shared memory by ONE bulk async copy (cp.async.bulk -> UBLKCP on the TMA engine, completion on an mbarrier)
// while the CTA's first segments are already being loaded (coalesced float4).
//  1. park: every thread snaps its segments, tests the START cell, and resolves on the spot what needs no walk
//     (NaN coordinate, start outside the map or on an occupied cell -- about half of all segments on dense
//     maps).  The survivors are COMPACTED into a shared-memory stage with a warp ballot scan (one shared atomic
//     per warp), already in walk form {bit address, last step, minor-axis remainder increments}.
//  2. walk: the major axis moves one cell per step, the minor axis carries the remainder of the floor() above
//     (no division; straight-line predicated code).  A lane that finishes (occupied cell, left the map, end
//     reached) pulls the NEXT parked segment instead of idling until the slowest lane of its warp is done:
//     __ballot_sync finds the idle lanes and ranks them, the warp claims stage entries 64 at a time.
//     Walk lengths are wildly uneven, so this keeps ~all lanes busy where a fixed lane<->segment mapping kept
//     a quarter of them busy.
//  3. flush: results (first blocked step, -1 = free) are written back coalesced.
#include "common.cuh"

namespace ppnet {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one-lane shared-memory counter bump.  Plain atomicAdd() under `if (lane == 0)` makes the compiler wrap its warp-aggregation
// sequence (vote, find-leader, popc, shuffle) around an instruction only one lane executes: ~12 instructions per claim.
__device__ __forceinline__ int smem_add(int* p, int v) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ long long floordiv64(long long num, long long den) {   // den > 0
    return num >= 0 ? num / den : -((den - 1 - num) / den);
}

#ifndef PPNET_DDA_THREADS
#define PPNET_DDA_THREADS 256
#endif
#ifndef PPNET_DDA_PER
#define PPNET_DDA_PER 4
#endif
constexpr int kDdaThreads = PPNET_DDA_THREADS;
constexpr int kDdaPerThread = PPNET_DDA_PER;
constexpr int kDdaStage = kDdaThreads * kDdaPerThread;   // segments per stage (32 KB of walk records)
constexpr int kDdaClaim = 64;                // stage entries a warp claims at a time
constexpr float kCoordClampF = 536870912.0f; // 2^29
constexpr int kLaneWalkMaxN = 1 << 28;       // remainders stay inside int32

// A4 rule at step 1 / offset 0: rintf is round-half-to-even and exact (floats >= 2^23 are integers already), so
// it equals rint((double)v); the clamp to +-2^29 is exact in float.
__device__ __forceinline__ int snap(float v) { return (int)fminf(fmaxf(rintf(v), -kCoordClampF), kCoordClampF); }

// over-long segments (n > 2^28; only reachable with coordinates far outside the map): division per cell, but the
// walk leaves the map within R steps
__device__ __noinline__ int walk_slow(const uint32_t* __restrict__ bm, int R, int W, int x0, int y0, int dx, int dy) {
    const long long n = max(abs((long long)dx), abs((long long)dy));
    for (long long k = 0; k <= n; ++k) {
        const long long cx = x0 + floordiv64(2 * k * dx + n, 2 * n), cy = y0 + floordiv64(2 * k * dy + n, 2 * n);
        if (cx < 0 || cx >= R || cy < 0 || cy >= R) return (int)k;
        if ((bm[(int)cy * W + ((int)cx >> 5)] >> ((int)cx & 31)) & 1u) return (int)k;
    }
    return -1;
}

// segment sources: float32 (s_x, s_y, e_x, e_y) as the entry point has always taken them, or the A11 array
// (s_row, s_col, e_row, e_col) float64 read directly -- the same cast + swap the fused verdict kernel applies, so the
// float32 copy of the segments never has to exist (host pipeline: one upload, three verdicts).
__device__ __forceinline__ float4 load_seg(const float* segs, int64_t i) {
    return __ldg(reinterpret_cast<const float4*>(segs) + i);
}
__device__ __forceinline__ float4 load_seg(const double* segs, int64_t i) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(segs) + 2 * i);
    const double2 b = __ldg(reinterpret_cast<const double2*>(segs) + 2 * i + 1);
    return make_float4((float)a.y, (float)a.x, (float)b.y, (float)b.x);
}

// FIRST = the caller wants the first blocked step.  Without it (verdict only) a segment whose END cell is outside the map
// or occupied is blocked whatever happens before (cell_n = the end cell; leaving the map earlier is blocked too), so it is
// resolved at park time like a blocked start cell: a third fewer segments walk, and the ones that no longer do are the
// partial walks.
// MINB = resident CTAs the register allocation aims at: 4 for small bitmaps (64 registers; measured at config 2: no hint
// 0.246, 3 -> 0.259, 4 -> 0.232, 5 -> 0.244 ms), 1 for bitmaps that fill shared memory on their own (84 registers, no
// pressure: config 5 0.349 -> 0.240 ms).
template <typename TS, bool FIRST, int MINB>
__global__ void __launch_bounds__(kDdaThreads, MINB)
dda_kernel(const uint32_t* __restrict__ bits, int R, int W, const TS* __restrict__ segs,
           const int64_t* __restrict__ seg_off, int64_t segs_per_map, int chunk,
           uint8_t* __restrict__ verdict, int32_t* __restrict__ first_hit, uint32_t* __restrict__ vbits,
           int exclusive_words) {
    extern __shared__ __align__(128) unsigned char dsm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm);                      // 16 B header: mbarrier, claim counter, parked count
    int* next = reinterpret_cast<int*>(dsm + 8);
    int* parked = reinterpret_cast<int*>(dsm + 12);
    uint32_t* bm = reinterpret_cast<uint32_t*>(dsm + 16);
    const int m = blockIdx.x;
    const int64_t lo = seg_off ? seg_off[m] : (int64_t)m * segs_per_map;
    const int64_t hi = seg_off ? seg_off[m + 1] : lo + segs_per_map;
    const int64_t base = lo + (int64_t)blockIdx.y * chunk;
    if (base >= hi) return;
    const int64_t end = min(hi, base + (int64_t)chunk);
    const uint32_t bytes = (uint32_t)(R * W * 4);
    int4* stage = reinterpret_cast<int4*>(dsm + 16 + ((bytes + 15u) & ~15u));   // walk records of the parked segments (2 x int4 each)
    uint16_t* sidx = reinterpret_cast<uint16_t*>(stage + 2 * kDdaStage);       // their slot in the stage's result array
    int32_t* res = reinterpret_cast<int32_t*>(sidx + kDdaStage);                // first blocked step per segment, -1 free
    const bool bulk = (bytes & 15u) == 0;

    if (bulk) {
        // regular case: one bulk async copy (TMA engine; a contiguous block needs no tensor map)
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(bm)), "l"(bits + (size_t)m * R * W), "r"(bytes), "r"(smem_u32(bar))
                         : "memory");
        }
    } else {
        // odd-sized bitmaps (R*W not a multiple of 4 words) cannot use the bulk engine: plain loads
        const uint32_t* src = bits + (size_t)m * R * W;
        for (int i = threadIdx.x; i < R * W; i += kDdaThreads) bm[i] = __ldg(src + i);
    }

    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int RS = W * 32;                                                 // bits per bitmap row
    for (int64_t st0 = base; st0 < end; st0 += kDdaStage) {
        const int ns = (int)min((int64_t)kDdaStage, end - st0);
        // ---- 1. park ----------------------------------------------------------------------------------------
        float4 sv[kDdaPerThread];
#pragma unroll
        for (int j = 0; j < kDdaPerThread; ++j) {                          // all loads in flight before the first use
            const int t = threadIdx.x + j * kDdaThreads;
            sv[j] = t < ns ? load_seg(segs, st0 + t) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (threadIdx.x == 0) { *next = 0; *parked = 0; }
        __syncthreads();                                                   // orders the barrier init / plain bitmap loads / counters
        if (st0 == base && bulk) {
            uint32_t done = 0;                                             // wait for phase 0 (HW sleep, no spin on memory)
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(smem_u32(bar)) : "memory");
            }
        }
#pragma unroll
        for (int j = 0; j < kDdaPerThread; ++j) {
            const int t = threadIdx.x + j * kDdaThreads;
            const float4 s = sv[j];
            const bool have = t < ns;
            const bool nan = !(s.x == s.x && s.y == s.y && s.z == s.z && s.w == s.w);
            const int x0 = snap(s.x), y0 = snap(s.y);
            const int dx = snap(s.z) - x0, dy = snap(s.w) - y0;            // |d| <= 2^30 after the clamp
            const bool inside = (unsigned)x0 < (unsigned)R && (unsigned)y0 < (unsigned)R;
            const int a = y0 * RS + x0;
            bool blocked0 = nan || !inside;                                // NaN coordinate or start outside the map: blocked at k = 0
            if (!blocked0) blocked0 = (bm[a >> 5] >> (a & 31)) & 1u;
            if (!FIRST && !blocked0) {                                     // verdict only: the end cell decides as well
                const int x1 = x0 + dx, y1 = y0 + dy;                      // (|x0|, |dx| <= 2^30: no overflow)
                const bool in1 = (unsigned)x1 < (unsigned)R && (unsigned)y1 < (unsigned)R;
                const int a1 = in1 ? y1 * RS + x1 : 0;
                blocked0 = !in1 || ((bm[a1 >> 5] >> (a1 & 31)) & 1u);
            }
            const int adx = abs(dx), ady = abs(dy);
            const int n = max(adx, ady);
            int r0 = blocked0 ? 0 : (n == 0 ? -1 : -2);                    // -2: needs a walk
            if (r0 == -2 && n > kLaneWalkMaxN) r0 = walk_slow(bm, R, W, x0, y0, dx, dy);
            const bool walk = have && r0 == -2;
            if (have && r0 != -2) res[t] = r0;
            // compaction of the survivors: ballot scan inside the warp, one shared atomic per warp
            const unsigned wm = __ballot_sync(0xffffffffu, walk);
            int wb = 0;
            if (lane == 0 && wm) wb = smem_add(parked, __popc(wm));
            wb = __shfl_sync(0xffffffffu, wb, 0);
            if (walk) {
                const bool xmaj = adx >= ady;
                const int dM = xmaj ? dx : dy, cM = xmaj ? x0 : y0;
                const int kM = dM > 0 ? R - cM : cM + 1;                   // first k whose major coordinate is outside
                const int kend = min(n, kM - 1);
                const int dm = xmaj ? dy : dx;                             // minor-axis delta, |dm| <= n
                const int pos = wb + __popc(wm & lt);
                PPNET_ASSERT(pos >= 0 && pos < kDdaStage && a >= 0 && a < R * RS);
                // the record is stored in the walk's own terms: parking runs with every lane busy, the refill below with a
                // third of them, so whatever can be derived here is.  Remainder of (2 k d_minor + n) mod 2n kept as a count-up
                // for either sign: for d_minor < 0 the mirrored remainder 2n - 1 - r starts at n - 1 and wraps exactly when r
                // would drop below 0.
                const int sM = (dM > 0 ? 1 : -1) * (xmaj ? 1 : RS);                 // address step of the major axis
                const int sm = (dm < 0 ? -1 : 1) * (xmaj ? RS : 1);                 // ... of the minor axis when it moves
                stage[2 * pos] = make_int4(a, kend | ((xmaj ? y0 : x0) << 16), n - (dm < 0 ? 1 : 0),
                                           kM - 1 < n ? kend + 1 : -1);             // leaves by the major axis next (blocked) / end (free)
                stage[2 * pos + 1] = make_int4(2 * abs(dm), 2 * n, sM, sm);
                sidx[pos] = (uint16_t)t;
            }
        }
        __syncthreads();
        const int np_ = *parked;

        // ---- 2. walk: lanes pull parked segments until the stage is empty -----------------------------------
        bool live = false, exhausted = np_ == 0;
        int32_t* res_ptr = res;                                            // where this lane's current segment reports
        int res_end = -1;                                                  // its result if the walk reaches kend unblocked
        int k = 0, kend = 0, a = 0, r = 0, n2 = 0, inc = 0, stepM = 0, stepm = 0, stepc = 0, cm = 0;
        const uint32_t bm_s = smem_u32(bm);
        int wnext = 0, wend = 0;                                           // this warp's claimed range (warp-uniform)
        for (;;) {
            const unsigned need = __ballot_sync(0xffffffffu, !live);
            if (need && !exhausted) {
                if (wnext >= wend) {
                    // guided self-scheduling: big claims first, small ones near the end (short tail per warp)
                    const int c = max(8, min(kDdaClaim, (np_ - wend) >> 4));
                    int b = 0;
                    if (lane == 0) b = smem_add(next, c);
                    b = __shfl_sync(0xffffffffu, b, 0);
                    wnext = b;
                    wend = min(b + c, np_);
                    exhausted = b >= np_;
                }
                if (!exhausted) {
                    const int my = wnext + __popc(need & lt);
                    if (!live && my < wend) {
                        PPNET_ASSERT(my >= 0 && my < np_ && sidx[my] < ns);
                        const int4 q = stage[2 * my], w = stage[2 * my + 1];
                        res_ptr = res + sidx[my];
                        a = q.x;
                        kend = q.y & 0xffff;
                        cm = q.y >> 16;
                        r = q.z;
                        res_end = q.w;
                        inc = w.x;                                 // 2 |d_minor|
                        n2 = w.y;
                        stepM = w.z;
                        stepm = w.w;
                        stepc = w.w < 0 ? -1 : 1;
                        k = 0;
                        live = true;
                    }
                    wnext = min(wend, wnext + __popc(need));
                }
            }
            if (!__any_sync(0xffffffffu, live)) {
                if (exhausted) break;
                continue;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {                                  // 4 cells between two warp votes; straight-line code
                const bool in = live && (unsigned)cm < (unsigned)R;       // minor coordinate still inside?
                uint32_t word = 0xffffffffu;
                PPNET_ASSERT(!in || (a >= 0 && (a >> 5) < R * W));
                if (in) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(bm_s + ((uint32_t)(a >> 5) << 2)));
                const bool blocked = (word >> (a & 31)) & 1u;              // occupied, or outside by the minor axis
                const bool done = live && (blocked || k == kend);
                if (done) *res_ptr = blocked ? k : res_end;
                live = live && !done;
                ++k;                                                       // (a finished lane's state is dead; advancing it is harmless)
                r += inc;
                const bool wrap = r >= n2;                                 // the minor axis moves one cell
                r -= wrap ? n2 : 0;
                a += stepM + (wrap ? stepm : 0);
                cm += wrap ? stepc : 0;
            }
        }
        __syncthreads();
        // ---- 3. flush the stage's results (coalesced) ----------------------------------------------------------
        for (int t0 = 0; t0 < ns; t0 += kDdaThreads) {                   // uniform trip count: the ballot below is warp-wide
            const int t = t0 + threadIdx.x;
            const int rr = t < ns ? res[t] : -1;
            if (t < ns) {
                if (verdict) verdict[st0 + t] = (uint8_t)(rr >= 0);
                if (first_hit) first_hit[st0 + t] = rr;
            }
            if (vbits) {                                                   // bit (i & 31) of word (i >> 5) = segment i
                const uint32_t m = __ballot_sync(0xffffffffu, rr >= 0);
                if (lane == 0 && t < ns) {
                    const int64_t i0 = st0 + t, k = i0 >> 5;
                    const int sh = (int)(i0 & 31);
                    if (exclusive_words) vbits[k] = m;
                    else {                                                 // shared words were zeroed by the launcher
                        if (m << sh) atomicOr(vbits + k, m << sh);
                        if (sh && (m >> (32 - sh))) atomicOr(vbits + k + 1, m >> (32 - sh));
                    }
                }
            }
        }
        __syncthreads();                                                   // res / stage are rewritten by the next stage
    }
}

}  // namespace ppnet

using namespace ppnet;

template <typename TS>
static int launch_dda(const uint32_t* bits, int32_t resolution, int64_t n_maps, const TS* segs, int64_t n_segs,
                      const int64_t* seg_off, int64_t segs_per_map, uint8_t* verdict, int32_t* first_hit, uint32_t* vbits,
                      void* stream) {
    PPNET_REQUIRE(n_maps >= 0 && n_segs >= 0 && resolution > 0, "dda: bad sizes");
    if (n_maps == 0 || n_segs == 0) return PPNET_OK;
    PPNET_REQUIRE(bits && segs && (verdict || vbits), "dda: null pointer");
    PPNET_REQUIRE(seg_off || segs_per_map * n_maps == n_segs, "dda: bad uniform grouping");
    PPNET_REQUIRE(seg_off == nullptr || segs_per_map > 0, "dda: pass the longest row in segs_per_map with a CSR");
    const int W = (resolution + 31) / 32;
    const size_t bm_bytes = (size_t)resolution * W * 4;
    PPNET_REQUIRE((reinterpret_cast<uintptr_t>(bits) & 15) == 0 && (reinterpret_cast<uintptr_t>(segs) & 15) == 0,
                  "dda: bits and segs must be 16-byte aligned");
    const size_t smem = ((bm_bytes + 15) & ~(size_t)15) + 16 + (size_t)kDdaStage * (32 + 2 + 4);
    PPNET_REQUIRE(smem <= 220 * 1024, "dda: resolution too large for a shared-memory bitmap");
    // big bitmaps: amortise the staging over every segment of the map; small ones: more CTAs in flight
    const int chunk = bm_bytes >= 64 * 1024 ? 8192 : kDdaStage;
    const int64_t chunks = (segs_per_map + chunk - 1) / chunk;
    PPNET_REQUIRE(chunks <= 65535, "dda: too many segments in one map");
    cudaStream_t st = (cudaStream_t)stream;
    // a 32-segment flush group owns its output word when every row starts on a multiple of 32
    const int exclusive = (seg_off == nullptr && segs_per_map % 32 == 0) ? 1 : 0;
    if (vbits && !exclusive) PPNET_CUDA(cudaMemsetAsync(vbits, 0, 4 * (size_t)((n_segs + 31) / 32), st));
    dim3 grid((unsigned)n_maps, (unsigned)chunks);
#define PPNET_DDA_LAUNCH(F, B)                                                                                              \
    do {                                                                                                                    \
        if (smem > 48 * 1024)                                                                                               \
            PPNET_CUDA(cudaFuncSetAttribute(dda_kernel<TS, F, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dda_kernel<TS, F, B><<<grid, kDdaThreads, smem, st>>>(bits, resolution, W, segs, seg_off, segs_per_map, chunk, verdict,   \
                                                              F ? first_hit : nullptr, vbits, exclusive);                  \
    } while (0)
    const bool big = bm_bytes >= 64 * 1024;                   // at most two such CTAs fit an SM
    if (first_hit) { if (big) PPNET_DDA_LAUNCH(true, 1); else PPNET_DDA_LAUNCH(true, 4); }
    else           { if (big) PPNET_DDA_LAUNCH(false, 1); else PPNET_DDA_LAUNCH(false, 4); }
#undef PPNET_DDA_LAUNCH
    PPNET_LAUNCH_CHECK("dda_kernel");
    return PPNET_OK;
}

extern "C" int ppnet_dda_gridcheck(const uint32_t* bits, int32_t resolution, int64_t n_maps, const float* segs_xy,
                                   int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map, uint8_t* verdict,
                                   int32_t* first_hit, void* stream) {
    PPNET_REQUIRE(verdict || n_segs == 0 || n_maps == 0, "dda: verdict is null");
    return launch_dda<float>(bits, resolution, n_maps, segs_xy, n_segs, seg_off, segs_per_map, verdict, first_hit, nullptr, stream);
}

extern "C" int ppnet_dda_gridcheck_rc64(const uint32_t* bits, int32_t resolution, int64_t n_maps, const double* segs_rc,
                                        int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map, uint8_t* verdict,
                                        int32_t* first_hit, uint32_t* vbits, void* stream) {
    return launch_dda<double>(bits, resolution, n_maps, segs_rc, n_segs, seg_off, segs_per_map, verdict, first_hit, vbits, stream);
}
