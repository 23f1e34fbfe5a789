// Counter-based sampling (Philox4x32-10): uniform draws and the GMM waypoint sampler.
//
//   A17  GMM  EDaGe-PP/GMM.py:7-16:  mean ~ U(0, mean_range)^{KxD}, std ~ U(0, std_range)^{KxD},
//        w ~ U(0,1)^K (normalised by Categorical); Distribution.sample([N]) -> f32[N, D].
//
// The reference draws from torch's global Mersenne-Twister stream; that stream is inherently serial.
// Here every draw is a pure function of (seed, stream id, unit index, block), so any sharding of the
// index range over GPUs reproduces the same numbers.  Parity with the reference is statistical
// (tests: chi^2 on component frequencies, KS on both marginals) -- SURVEY 8(a) A17.
#include <algorithm>

#include "common.cuh"

namespace ppnet {

// out[u][k], k < per_unit : 53-bit doubles; block = k/2, words (x,y) for even k, (z,w) for odd k
__global__ void uniform_kernel(uint2 key, uint32_t stream_id, uint64_t unit0, int64_t n_units, int per_unit,
                               double* __restrict__ out) {
    const int64_t pairs_per_unit = (per_unit + 1) / 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_units * pairs_per_unit) return;
    const int64_t u = t / pairs_per_unit;
    const uint32_t b = (uint32_t)(t % pairs_per_unit);
    const uint64_t g = unit0 + (uint64_t)u;
    const uint4 r = Philox::gen(key, make_uint4(b, stream_id, (uint32_t)g, (uint32_t)(g >> 32)));
    double* o = out + u * per_unit + 2 * (int64_t)b;
    o[0] = u53(r.x, r.y);
    if (2 * (int64_t)b + 1 < per_unit) o[1] = u53(r.z, r.w);
}

// GMM.__init__: all parameters are torch.rand (24-bit floats) times a range, float32
__global__ void gmm_param_kernel(uint2 key, int K, int D, float mean_range, float std_range,
                                 float* __restrict__ mean, float* __restrict__ stdv, float* __restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * D) return;
    const uint4 r = Philox::gen(key, make_uint4((uint32_t)i, STREAM_GMM_PARAM, 0u, 0u));
    mean[i] = __fmul_rn(u24(r.x), mean_range);
    stdv[i] = __fmul_rn(u24(r.y), std_range);
    if (i < K) w[i] = u24(r.z);
}

constexpr int kGmmMaxK = 64;

// one thread per sample.  block 0 of the sample's counter: x -> component, (y, z) -> Box-Muller pair
// for dims 0/1; dims >= 2 take further blocks (4 normals each).
// Box-Muller on (0,1] x [0,1): u1 = (wa + 1) / 2^32 never 0.  SFU transcendentals (MUFU.LG2 / SQRT / SIN / COS,
// |error| ~ 2^-21): sampling parity is statistical, and the argument of sin / cos is folded into [-pi, pi) where the
// hardware approximations are at their best: cos(2 pi u) = -cos(2 pi u - pi), sin(2 pi u) = -sin(2 pi u - pi)
__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, float& z0, float& z1) {
    const float u1 = ((float)(wa >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float u2 = u24(wb);
    float rad;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(rad) : "f"(-2.0f * __logf(u1)));
    float sn, cs;
    __sincosf(6.283185307179586f * u2 - 3.141592653589793f, &sn, &cs);
    z0 = rad * -cs;
    z1 = rad * -sn;
}

// kPlanar: D == 2 (the reference's only use, MapGenerate.py / GMM.py): parameters as float2 in shared memory, one
// 8-byte store per sample, no per-dimension loop.
template <bool kPlanar>
__global__ void __launch_bounds__(256)
gmm_sample_kernel(uint2 key, uint64_t sample0, int64_t n, int K, int D, const float* __restrict__ mean,
                  const float* __restrict__ stdv, const float* __restrict__ w, float* __restrict__ out,
                  int32_t* __restrict__ comp) {
    __shared__ float cdf[kGmmMaxK];
    __shared__ __align__(8) float s_mean[kGmmMaxK * 4], s_std[kGmmMaxK * 4];
    if (threadIdx.x == 0) {                               // normalised inclusive CDF, serial f32 sum
        float tot = 0.f;
        for (int k = 0; k < K; ++k) tot = __fadd_rn(tot, w[k]);
        float acc = 0.f;
        for (int k = 0; k < K; ++k) { acc = __fadd_rn(acc, w[k]); cdf[k] = __fdiv_rn(acc, tot); }
        // entries the search must never count: the last one (k <= K - 1) and the padding up to 15
        if (K <= 16) for (int k = K - 1; k < 16; ++k) cdf[k] = 2.0f;
    }
    const bool cache = D <= 4;
    if (cache) for (int i = threadIdx.x; i < K * D; i += blockDim.x) { s_mean[i] = mean[i]; s_std[i] = stdv[i]; }
    __syncthreads();
    // persistent grid-stride loop: the CDF / parameter staging above is paid once per CTA, not once per 256 samples
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t g = sample0 + (uint64_t)i;
        uint4 r = Philox::gen(key, make_uint4(0u, STREAM_GMM_SAMPLE, (uint32_t)g, (uint32_t)(g >> 32)));
        const float uc = u24(r.x);
        // inverse CDF: the CDF is non-decreasing (weights >= 0, as torch's Categorical requires), so the first k with
        // uc < cdf[k] is the count of entries <= uc among the first K - 1: four probes of the padded table, else a scan
        int k = 0;
        if (K <= 16) {
            k += (uc >= cdf[k + 7]) << 3;
            k += (uc >= cdf[k + 3]) << 2;
            k += (uc >= cdf[k + 1]) << 1;
            k += (uc >= cdf[k]);
        } else {
            for (int j = 0; j < K - 1; ++j) k += (uc >= cdf[j]);
        }
        if (comp) comp[i] = k;
        if (kPlanar) {
            float z0, z1;
            box_muller(r.y, r.z, z0, z1);
            const float2 mu = reinterpret_cast<const float2*>(s_mean)[k], sd = reinterpret_cast<const float2*>(s_std)[k];
            reinterpret_cast<float2*>(out)[i] = make_float2(mu.x + sd.x * z0, mu.y + sd.y * z1);
            continue;
        }
        uint32_t wa = r.y, wb = r.z;
        for (int d = 0; d < D; d += 2) {
            if (d >= 2) {
                const int q = (d - 2) >> 1;                   // pair index among the extra blocks
                if ((q & 1) == 0) r = Philox::gen(key, make_uint4(1u + (uint32_t)(q >> 1), STREAM_GMM_SAMPLE,
                                                                 (uint32_t)g, (uint32_t)(g >> 32)));
                wa = (q & 1) ? r.z : r.x;
                wb = (q & 1) ? r.w : r.y;
            }
            float z0, z1;
            box_muller(wa, wb, z0, z1);
            const float m0 = cache ? s_mean[k * D + d] : mean[k * D + d];
            const float s0 = cache ? s_std[k * D + d] : stdv[k * D + d];
            out[i * D + d] = m0 + s0 * z0;
            if (d + 1 < D) {
                const float m1 = cache ? s_mean[k * D + d + 1] : mean[k * D + d + 1];
                const float s1 = cache ? s_std[k * D + d + 1] : stdv[k * D + d + 1];
                out[i * D + d + 1] = m1 + s1 * z1;
            }
        }
    }
}

// Candidate segments of map g (the workload SURVEY 8(d) config 2 describes: s ~ U(0, R)^2, e = s + N(0, sigma^2) per
// axis), drawn on the device so that generator-mode callers upload nothing.  Segment k of map g is a pure function of
// (seed, g, k): Philox block 2k -> start (two 53-bit uniforms), block 2k+1 -> Box-Muller offset in float64.
__global__ void propose_segments_kernel(uint2 key, uint64_t map0, int64_t n_maps, int64_t spm, double R, double sigma,
                                        double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_maps * spm) return;
    const uint64_t g = map0 + (uint64_t)(t / spm);
    const uint32_t k = (uint32_t)(t % spm);
    const uint4 a = Philox::gen(key, make_uint4(2u * k, STREAM_SEGS, (uint32_t)g, (uint32_t)(g >> 32)));
    const uint4 b = Philox::gen(key, make_uint4(2u * k + 1u, STREAM_SEGS, (uint32_t)g, (uint32_t)(g >> 32)));
    const double s0 = __dmul_rn(u53(a.x, a.y), R), s1 = __dmul_rn(u53(a.z, a.w), R);
    const double u1 = u53(b.x, b.y) + (1.0 / 9007199254740992.0);          // (0, 1]
    const double rad = __dmul_rn(sqrt(-2.0 * log(u1)), sigma);
    double sn, cs;
    sincospi(2.0 * u53(b.z, b.w), &sn, &cs);
    double2* o = reinterpret_cast<double2*>(out + 4 * t);
    o[0] = make_double2(s0, s1);
    o[1] = make_double2(__dadd_rn(s0, __dmul_rn(rad, cs)), __dadd_rn(s1, __dmul_rn(rad, sn)));
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_propose_segments(uint64_t seed, uint64_t map0, int64_t n_maps, int64_t segs_per_map, double resolution,
                                      double sigma, double* segs_rc, void* stream) {
    PPNET_REQUIRE(n_maps >= 0 && segs_per_map >= 0 && segs_per_map <= 1073741823LL, "propose_segments: bad sizes");
    if (n_maps == 0 || segs_per_map == 0) return PPNET_OK;
    PPNET_REQUIRE(segs_rc && (reinterpret_cast<uintptr_t>(segs_rc) & 15) == 0, "propose_segments: output must be 16-byte aligned");
    const int64_t n = n_maps * segs_per_map;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    propose_segments_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(key, map0, n_maps, segs_per_map,
                                                                                            resolution, sigma, segs_rc);
    PPNET_LAUNCH_CHECK("propose_segments_kernel");
    return PPNET_OK;
}

static inline uint2 make_key(uint64_t seed) { return make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)); }

extern "C" int ppnet_uniform_f64(uint64_t seed, uint32_t stream_id, uint64_t unit0, int64_t n_units,
                                 int32_t per_unit, double* out, void* stream) {
    PPNET_REQUIRE(n_units >= 0 && per_unit >= 0, "uniform: negative sizes");
    if (n_units == 0 || per_unit == 0) return PPNET_OK;
    PPNET_REQUIRE(out, "uniform: out is null");
    const int64_t total = n_units * ((per_unit + 1) / 2);
    uniform_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(make_key(seed), stream_id,
                                                                                     unit0, n_units, per_unit, out);
    PPNET_LAUNCH_CHECK("uniform_kernel");
    return PPNET_OK;
}

extern "C" int ppnet_gmm_params(uint64_t seed, int32_t order, int32_t dim, float mean_range, float std_range,
                                float* mean, float* stdv, float* weights, void* stream) {
    PPNET_REQUIRE(order > 0 && dim > 0, "gmm_params: order and dim must be positive");
    PPNET_REQUIRE(mean && stdv && weights, "gmm_params: null pointer");
    gmm_param_kernel<<<(order * dim + 127) / 128, 128, 0, (cudaStream_t)stream>>>(make_key(seed), order, dim,
                                                                                 mean_range, std_range, mean, stdv,
                                                                                 weights);
    PPNET_LAUNCH_CHECK("gmm_param_kernel");
    return PPNET_OK;
}

extern "C" int ppnet_gmm_sample(uint64_t seed, uint64_t sample0, int64_t n, int32_t order, int32_t dim,
                                const float* mean, const float* stdv, const float* weights, float* out,
                                int32_t* comp, void* stream) {
    PPNET_REQUIRE(n >= 0, "gmm_sample: negative n");
    PPNET_REQUIRE(order > 0 && order <= kGmmMaxK, "gmm_sample: order must be in 1..%d", kGmmMaxK);
    PPNET_REQUIRE(dim > 0 && dim <= 1026, "gmm_sample: bad dim");
    if (n == 0) return PPNET_OK;
    PPNET_REQUIRE(mean && stdv && weights && out, "gmm_sample: null pointer");
    const int64_t ctas = std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 8);
    if (dim == 2 && (reinterpret_cast<uintptr_t>(out) & 7) == 0)
        gmm_sample_kernel<true><<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(make_key(seed), sample0, n, order, dim, mean,
                                                                                  stdv, weights, out, comp);
    else
        gmm_sample_kernel<false><<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(make_key(seed), sample0, n, order, dim, mean,
                                                                                   stdv, weights, out, comp);
    PPNET_LAUNCH_CHECK("gmm_sample_kernel");
    return PPNET_OK;
}
