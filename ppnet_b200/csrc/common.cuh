// ppnet_b200 -- shared device/host helpers for the sm_100a kernels.
//
// All bit-exact kernels are compiled with -fmad=false: the compiler never contracts a*b+c.
// Where the reference itself is fused (OpenBLAS SkylakeX ddot / dgemm) the code says so with an
// explicit __fma_rn.  Division and square root are IEEE (nvcc default -prec-div/-prec-sqrt).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ppnet_b200.h"

namespace ppnet {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define PPNET_REQUIRE(cond, ...)                \
    do {                                        \
        if (!(cond)) {                          \
            ppnet::set_error(__VA_ARGS__);      \
            return PPNET_E_INVALID;             \
        }                                       \
    } while (0)

#define PPNET_CUDA(call)                                                              \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) {                                                      \
            ppnet::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),  \
                             __FILE__, __LINE__);                                     \
            return PPNET_E_CUDA;                                                      \
        }                                                                             \
    } while (0)

#define PPNET_LAUNCH_CHECK(name)                                                       \
    do {                                                                               \
        cudaError_t _e = cudaGetLastError();                                           \
        if (_e != cudaSuccess) {                                                       \
            ppnet::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
            return PPNET_E_CUDA;                                                       \
        }                                                                              \
        ppnet::count_launch();                                                         \
    } while (0)

constexpr int kNumSMs = 148;   // B200

// Bounds / protocol checks of our own on every shared-memory queue, stage and scatter index (compute-sanitizer is not
// available on the GPU pool): compiled in by `make debug` (-DPPNET_DEBUG_BOUNDS -> libppnet_b200_dbg.so), the GPU
// parity tests are then run against that library (scripts/gpu_debug_bounds.sh); compiled out of the product.
#ifdef PPNET_DEBUG_BOUNDS
#include <cassert>
#define PPNET_ASSERT(cond) assert(cond)
#else
#define PPNET_ASSERT(cond) ((void)0)
#endif

// ---- the reference's 2-vector dot product (np.dot -> OpenBLAS ddot), see SURVEY 8(c) ----------
template <int MODE>
__device__ __forceinline__ double dot2(double a0, double a1, double b0, double b1) {
    if (MODE == PPNET_DOT_FUSED_SKX) return __fma_rn(a1, b1, __dmul_rn(a0, b0));
    return __dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1));
}
// float32 np.dot is un-fused under every OpenBLAS kernel tested
__device__ __forceinline__ float dot2f(float a0, float a1, float b0, float b1) {
    return __fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1));
}

// Smallest x with sqrt_rn(x) >= t  (so that  sqrt_rn(q) < t  <=>  q < T  for every q >= 0, exactly:
// correctly rounded sqrt is monotone).  t <= 0 or NaN -> 0 (never true for q >= 0 / NaN).
__device__ __forceinline__ double sqrt_lt_threshold(double t) {
    if (!(t > 0.0)) return 0.0;
    if (isinf(t)) return t;
    double x = __dmul_rn(t, t);
    if (isinf(x)) return x;   // t*t overflows: every finite q has sqrt(q) < t only if ...; keep inf
    // walk to the boundary (a handful of steps at most)
    while (x > 0.0 && sqrt(__longlong_as_double(__double_as_longlong(x) - 1)) >= t)
        x = __longlong_as_double(__double_as_longlong(x) - 1);
    while (sqrt(x) < t) x = __longlong_as_double(__double_as_longlong(x) + 1);
    return x;
}
__device__ __forceinline__ float sqrt_lt_threshold(float t) {
    if (!(t > 0.0f)) return 0.0f;
    if (isinf(t)) return t;
    float x = __fmul_rn(t, t);
    if (isinf(x)) return x;
    while (x > 0.0f && __fsqrt_rn(__int_as_float(__float_as_int(x) - 1)) >= t)
        x = __int_as_float(__float_as_int(x) - 1);
    while (__fsqrt_rn(x) < t) x = __int_as_float(__float_as_int(x) + 1);
    return x;
}

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based: reproducible for any sharding -------
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline uint4 gen(uint2 key, uint4 c) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
            uint32_t hi0 = __umulhi(M0, c.x), hi1 = __umulhi(M1, c.z);
#else
            uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c.x) >> 32);
            uint32_t hi1 = (uint32_t)(((uint64_t)M1 * c.z) >> 32);
#endif
            uint32_t lo0 = M0 * c.x, lo1 = M1 * c.z;
            c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
            key.x += W0;
            key.y += W1;
        }
        return c;
    }
};
// 53-bit double in [0,1) from two 32-bit words -- the same construction as numpy's random_sample
__host__ __device__ inline double u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// 24-bit float in [0,1) -- the same construction as torch.rand(float32) on CPU
__host__ __device__ inline float u24(uint32_t a) { return (float)(a >> 8) * (1.0f / 16777216.0f); }

// Philox sub-stream ids (counter word .y); word .x = draw block, (.z,.w) = 64-bit unit index
enum : uint32_t {
    STREAM_PLACE = 1,      // placement tries of map g: block t -> (angle, t0 ; t1 = next pair)
    STREAM_OBST = 2,       // obstacle candidates of map g
    STREAM_GMM_PARAM = 3,  // GMM parameters
    STREAM_GMM_SAMPLE = 4, // GMM samples
    STREAM_UNIFORM = 5,    // ppnet_uniform_f64
    STREAM_PATH = 6,       // path synthesis draws (A1)
    STREAM_PATH_OBST = 7,  // set_obstacles draws (A9)
    STREAM_SEGS = 8,       // candidate segment proposals of map g (ppnet_propose_segments)
};

}  // namespace ppnet
