// 64-bit content digest of per-map output arrays, for the cross-rank identity proof (SURVEY 4 tier 5 / 8(e)): the dataset
// a run produces must not depend on how the global map range was split over ranks or launches.  Every 32-bit word is
// mixed with its position and the GLOBAL map index it belongs to, and the mixed words are SUMMED modulo 2^64 --
// associative and commutative, so digest([a, c)) = digest([a, b)) + digest([b, c)) for any split, on any number of GPUs.
#include "common.cuh"

namespace ppnet {

__host__ __device__ inline uint64_t mix64(uint64_t x) {          // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

// one CTA per unit (map); only the first rows[u] * row_words words of a unit count when `rows` is given (obstacle
// sets: rows beyond obs_cnt are unspecified)
__global__ void __launch_bounds__(256)
digest_kernel(const uint32_t* __restrict__ data, int64_t words_per_unit, uint64_t unit0, const int32_t* __restrict__ rows,
              int row_words, uint64_t salt, unsigned long long* __restrict__ acc) {
    __shared__ unsigned long long wsum[8];
    const int64_t u = blockIdx.x;
    const uint64_t g = unit0 + (uint64_t)u;
    int64_t nw = words_per_unit;
    if (rows) nw = min(nw, (int64_t)max(rows[u], 0) * row_words);
    const uint32_t* p = data + u * words_per_unit;
    const uint64_t key = mix64(g * 0x9E3779B97F4A7C15ull + salt);
    unsigned long long s = 0;
    for (int64_t k = threadIdx.x; k < nw; k += 256) s += mix64((uint64_t)__ldg(p + k) ^ mix64(key + (uint64_t)k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        atomicAdd(acc, t);
    }
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_digest_u32(const uint32_t* data, int64_t words_per_unit, int64_t n_units, uint64_t unit0,
                                const int32_t* rows, int32_t row_words, uint64_t salt, uint64_t* acc, void* stream) {
    PPNET_REQUIRE(n_units >= 0 && words_per_unit >= 0 && n_units <= 2147483647LL, "digest: bad sizes");
    PPNET_REQUIRE(acc, "digest: acc is null");
    if (n_units == 0 || words_per_unit == 0) return PPNET_OK;
    PPNET_REQUIRE(data, "digest: data is null");
    PPNET_REQUIRE(!rows || row_words > 0, "digest: row_words must be positive with rows");
    digest_kernel<<<(unsigned)n_units, 256, 0, (cudaStream_t)stream>>>(data, words_per_unit, unit0, rows, row_words, salt,
                                                                      (unsigned long long*)acc);
    PPNET_LAUNCH_CHECK("digest_kernel");
    return PPNET_OK;
}
