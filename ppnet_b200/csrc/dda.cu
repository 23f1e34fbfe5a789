// Integer DDA grid check against bit-packed occupancy maps (new functionality: the reference has no
// occupancy-grid lookup).  Endpoints are snapped with the A4 rule (Path.coord_euclidean2image,
// EDaGe-PP/Path.py:378-386: rint = round-half-to-even; here step 1, offset 0), then the walk is all-integer:
//     n = max(|dx|, |dy|),  cell_k = (x0 + floor((2 k dx + n) / 2n),  y0 + floor((2 k dy + n) / 2n)),  k = 0..n
// verdict = first k whose cell is outside [0,R)^2 or occupied.
//
// One CTA per (map, chunk).  The map's bitmap (6 272 B at R = 224, 131 072 B at R = 1024) is pulled into
// shared memory by ONE bulk async copy (cp.async.bulk -> UBLKCP on the TMA engine, completion on an mbarrier).
// A warp owns a batch of 32 segments, one per lane.  Each lane walks its own segment with an exact
// incremental form of the floor() above (remainder accumulators, no division); the warp leaves the loop as
// soon as __ballot_sync says every lane is blocked or finished.  When only a few long segments survive
// (dense maps: most lanes hit early), the stragglers are finished cooperatively: 32 cells per round across the
// lanes, __ballot_sync picks the first blocked cell.
#include "common.cuh"

namespace ppnet {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ long long floordiv64(long long num, long long den) {   // den > 0
    return num >= 0 ? num / den : -((den - 1 - num) / den);
}

constexpr int kDdaThreads = 256;
constexpr int kCoordClamp = 1 << 29;
constexpr int kLaneWalkMaxN = 1 << 28;       // remainders stay inside int32
constexpr int kStragglers = 4;               // <= this many live lanes ...
constexpr int kCoopMinRemaining = 96;        // ... with more than this many cells left: finish them cooperatively

__device__ __forceinline__ int snap(float v, bool& bad) {
    if (!(v == v)) { bad = true; return 0; }
    const double r = rint((double)v);                                      // A4 rule, step 1, offset 0
    return (int)fmin(fmax(r, -(double)kCoordClamp), (double)kCoordClamp);
}

__device__ __forceinline__ bool cell_blocked(const uint32_t* __restrict__ bm, int R, int W, int cx, int cy) {
    if ((unsigned)cx >= (unsigned)R || (unsigned)cy >= (unsigned)R) return true;
    return (bm[cy * W + (cx >> 5)] >> (cx & 31)) & 1u;
}

__global__ void __launch_bounds__(kDdaThreads)
dda_kernel(const uint32_t* __restrict__ bits, int R, int W, const float* __restrict__ segs,
           const int64_t* __restrict__ seg_off, int64_t segs_per_map, int chunk,
           uint8_t* __restrict__ verdict, int32_t* __restrict__ first_hit) {
    extern __shared__ __align__(128) unsigned char dsm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm);                      // 16 B header
    uint32_t* bm = reinterpret_cast<uint32_t*>(dsm + 16);
    const int m = blockIdx.x;
    const int64_t lo = seg_off ? seg_off[m] : (int64_t)m * segs_per_map;
    const int64_t hi = seg_off ? seg_off[m + 1] : lo + segs_per_map;
    const int64_t base = lo + (int64_t)blockIdx.y * chunk;
    if (base >= hi) return;
    const int64_t end = min(hi, base + (int64_t)chunk);
    const uint32_t bytes = (uint32_t)(R * W * 4);

    if ((bytes & 15u) == 0) {
        // regular case: one bulk async copy (TMA engine; a contiguous block needs no tensor map)
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(bm)), "l"(bits + (size_t)m * R * W), "r"(bytes), "r"(smem_u32(bar))
                         : "memory");
        }
    }

    // while the copy is in flight: every lane loads and snaps its first segment
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int nwarps = kDdaThreads / 32;

    if ((bytes & 15u) == 0) {
        uint32_t done = 0;                                                 // wait for phase 0 (HW sleep, no spin on memory)
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(smem_u32(bar)) : "memory");
        }
    } else {
        // odd-sized bitmaps (R*W not a multiple of 4 words) cannot use the bulk engine: plain loads
        const uint32_t* src = bits + (size_t)m * R * W;
        for (int i = threadIdx.x; i < R * W; i += kDdaThreads) bm[i] = __ldg(src + i);
        __syncthreads();
    }

    for (int64_t b0 = base + 32 * warp; b0 < end; b0 += 32 * nwarps) {
        const int64_t mine = b0 + lane;                                    // 512 B coalesced per warp
        int x0 = 0, y0 = 0, dx = 0, dy = 0, n = 0;
        bool bad = false, have = mine < end;
        if (have) {
            const float4 s = __ldg(reinterpret_cast<const float4*>(segs) + mine);
            x0 = snap(s.x, bad); y0 = snap(s.y, bad);
            dx = snap(s.z, bad) - x0; dy = snap(s.w, bad) - y0;            // |d| <= 2^30 after the clamp
            n = max(abs(dx), abs(dy));
        }
        int first = -1;                                                    // first blocked k of MY segment
        bool live = have;
        if (have && bad) { first = 0; live = false; }                      // NaN coordinate: blocked at k = 0
        const bool lane_walk_ok = n <= kLaneWalkMaxN;

        // ---- phase 1: one lane per segment, exact incremental walk -------------------------------------
        int k = 0, cx = x0, cy = y0;
        int rx = n, ry = n;                                                // remainders of (2k d + n) mod 2n
        const int n2 = 2 * n, dx2 = 2 * dx, dy2 = 2 * dy;
        for (;;) {
            const unsigned alive = __ballot_sync(0xffffffffu, live && lane_walk_ok);
            if (!alive) break;                                             // every lane blocked or finished
            if (__popc(alive) <= kStragglers) {                            // few survivors with a long way to go?
                const int rem = (live && lane_walk_ok) ? n - k : 0;
                if (__reduce_max_sync(0xffffffffu, rem) > kCoopMinRemaining) break;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {                                  // 4 cells between two warp votes
                if (live && lane_walk_ok) {
                    if (cell_blocked(bm, R, W, cx, cy)) { first = k; live = false; }
                    else if (k == n) { live = false; }                    // reached the end cell: free
                    else {
                        ++k;
                        rx += dx2; ry += dy2;
                        if (rx >= n2) { rx -= n2; ++cx; } else if (rx < 0) { rx += n2; --cx; }
                        if (ry >= n2) { ry -= n2; ++cy; } else if (ry < 0) { ry += n2; --cy; }
                    }
                }
            }
        }

        // ---- phase 2: stragglers (and over-long segments), 32 cells per round across the warp ----------
        unsigned todo = __ballot_sync(0xffffffffu, live);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const int sx0 = __shfl_sync(0xffffffffu, x0, j), sy0 = __shfl_sync(0xffffffffu, y0, j);
            const long long sdx = __shfl_sync(0xffffffffu, dx, j), sdy = __shfl_sync(0xffffffffu, dy, j);
            const long long sn = __shfl_sync(0xffffffffu, n, j);
            const long long kstart = __shfl_sync(0xffffffffu, k, j);       // cells < kstart were already free
            int f = -1;
            for (long long k0 = kstart; k0 <= sn; k0 += 32) {
                const long long kk = k0 + lane;
                bool blocked = false;
                if (kk <= sn) {
                    long long ccx = sx0, ccy = sy0;
                    if (sn) {
                        ccx += floordiv64(2 * kk * sdx + sn, 2 * sn);
                        ccy += floordiv64(2 * kk * sdy + sn, 2 * sn);
                    }
                    blocked = (ccx < 0 || ccx >= R || ccy < 0 || ccy >= R) ? true
                              : (bool)((bm[(int)ccy * W + ((int)ccx >> 5)] >> ((int)ccx & 31)) & 1u);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, blocked);
                if (bal) { f = (int)k0 + (__ffs(bal) - 1); break; }        // early exit for the whole warp
            }
            if (lane == j) { first = f; live = false; }
        }

        if (have) {
            verdict[mine] = (uint8_t)(first >= 0);
            if (first_hit) first_hit[mine] = first;
        }
    }
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_dda_gridcheck(const uint32_t* bits, int32_t resolution, int64_t n_maps, const float* segs_xy,
                                   int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map, uint8_t* verdict,
                                   int32_t* first_hit, void* stream) {
    PPNET_REQUIRE(n_maps >= 0 && n_segs >= 0 && resolution > 0, "dda: bad sizes");
    if (n_maps == 0 || n_segs == 0) return PPNET_OK;
    PPNET_REQUIRE(bits && segs_xy && verdict, "dda: null pointer");
    PPNET_REQUIRE(seg_off || segs_per_map * n_maps == n_segs, "dda: bad uniform grouping");
    PPNET_REQUIRE(seg_off == nullptr || segs_per_map > 0, "dda: pass the longest row in segs_per_map with a CSR");
    const int W = (resolution + 31) / 32;
    const size_t bm_bytes = (size_t)resolution * W * 4;
    PPNET_REQUIRE((reinterpret_cast<uintptr_t>(bits) & 15) == 0 && (reinterpret_cast<uintptr_t>(segs_xy) & 15) == 0,
                  "dda: bits and segs must be 16-byte aligned");
    const size_t smem = bm_bytes + 16;
    PPNET_REQUIRE(smem <= 220 * 1024, "dda: resolution too large for a shared-memory bitmap");
    // big bitmaps: amortise the staging over every segment of the map; small ones: more CTAs in flight
    const int chunk = bm_bytes >= 64 * 1024 ? 8192 : 1024;
    const int64_t chunks = (segs_per_map + chunk - 1) / chunk;
    PPNET_REQUIRE(chunks <= 65535, "dda: too many segments in one map");
    if (smem > 48 * 1024)
        PPNET_CUDA(cudaFuncSetAttribute(dda_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)n_maps, (unsigned)chunks);
    dda_kernel<<<grid, kDdaThreads, smem, (cudaStream_t)stream>>>(bits, resolution, W, segs_xy, seg_off, segs_per_map,
                                                                  chunk, verdict, first_hit);
    PPNET_LAUNCH_CHECK("dda_kernel");
    return PPNET_OK;
}
