// torch.ops.ppnet_b200.* -- dispatcher registration of the hot-path entry points ("thin PyTorch C++ / C-ABI
// extension").  Every op checks its tensors, switches to their device, takes torch's current stream on that device and
// calls the C ABI of include/ppnet_b200.h; nothing is computed here and there is no CPU implementation (CUDA key only).
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include "../../../include/ppnet_b200.h"

namespace {

void need(const at::Tensor& t, at::ScalarType dt, const char* name) {
    TORCH_CHECK(t.is_cuda(), "ppnet_b200: ", name, " must be a CUDA tensor (no CPU fallback)");
    TORCH_CHECK(t.scalar_type() == dt, "ppnet_b200: ", name, " must be ", dt, ", got ", t.scalar_type());
    TORCH_CHECK(t.is_contiguous(), "ppnet_b200: ", name, " must be contiguous");
}
void same_device(const at::Tensor& a, const at::Tensor& b, const char* name) {
    TORCH_CHECK(a.device() == b.device(), "ppnet_b200: ", name, " is on ", b.device(), ", expected ", a.device());
}
void ok(int rc, const char* what) { TORCH_CHECK(rc == PPNET_OK, what, " failed (", rc, "): ", ppnet_last_error()); }
void* stream() { return (void*)at::cuda::getCurrentCUDAStream().stream(); }

struct Grouping {
    const int64_t* off;
    int64_t spm;
};
// uniform grouping (seg_off undefined) or a CSR whose longest row the caller states (no device read here)
Grouping grouping(int64_t n, int64_t m, const c10::optional<at::Tensor>& seg_off, int64_t max_segs_per_map, const at::Tensor& ref) {
    if (!seg_off.has_value()) {
        TORCH_CHECK(m > 0 && n % m == 0, "ppnet_b200: uniform grouping needs n_segs divisible by n_maps (or pass seg_off)");
        return {nullptr, n / m};
    }
    need(*seg_off, at::kLong, "seg_off");
    same_device(ref, *seg_off, "seg_off");
    TORCH_CHECK(seg_off->numel() == m + 1, "ppnet_b200: seg_off must have n_maps + 1 entries");
    TORCH_CHECK(max_segs_per_map > 0, "ppnet_b200: a CSR needs max_segs_per_map (the longest row)");
    return {seg_off->data_ptr<int64_t>(), max_segs_per_map};
}

at::Tensor segcheck_edage_f64(const at::Tensor& pts_rc, const at::Tensor& obs, const at::Tensor& obs_cnt, double clearance,
                              double bound, int64_t dot_mode, const c10::optional<at::Tensor>& seg_off, int64_t max_segs_per_map) {
    need(pts_rc, at::kDouble, "pts_rc"); need(obs, at::kDouble, "obs"); need(obs_cnt, at::kInt, "obs_cnt");
    same_device(pts_rc, obs, "obs"); same_device(pts_rc, obs_cnt, "obs_cnt");
    c10::cuda::CUDAGuard guard(pts_rc.device());
    const int64_t n = pts_rc.size(0), m = obs.size(0);
    const Grouping g = grouping(n, m, seg_off, max_segs_per_map, pts_rc);
    at::Tensor out = at::empty({n}, pts_rc.options().dtype(at::kByte));
    ok(ppnet_segcheck_edage_f64(pts_rc.data_ptr<double>(), n, g.off, g.spm, m, obs.data_ptr<double>(), obs_cnt.data_ptr<int32_t>(),
                                (int32_t)obs.size(1), clearance, bound, (int32_t)dot_mode, out.data_ptr<uint8_t>(), stream()),
       "ppnet_segcheck_edage_f64");
    return out;
}

at::Tensor segcheck_mpnet_f32(const at::Tensor& pts_xy, const at::Tensor& obs, const at::Tensor& obs_cnt, double clearance,
                              double bound, const c10::optional<at::Tensor>& seg_off, int64_t max_segs_per_map) {
    need(pts_xy, at::kFloat, "pts_xy"); need(obs, at::kDouble, "obs"); need(obs_cnt, at::kInt, "obs_cnt");
    same_device(pts_xy, obs, "obs"); same_device(pts_xy, obs_cnt, "obs_cnt");
    c10::cuda::CUDAGuard guard(pts_xy.device());
    const int64_t n = pts_xy.size(0), m = obs.size(0);
    const Grouping g = grouping(n, m, seg_off, max_segs_per_map, pts_xy);
    at::Tensor out = at::empty({n}, pts_xy.options().dtype(at::kByte));
    ok(ppnet_segcheck_mpnet_f32(pts_xy.data_ptr<float>(), n, g.off, g.spm, m, obs.data_ptr<double>(), obs_cnt.data_ptr<int32_t>(),
                                (int32_t)obs.size(1), clearance, bound, out.data_ptr<uint8_t>(), nullptr, stream()),
       "ppnet_segcheck_mpnet_f32");
    return out;
}

// -> (bits64, bits32): int32[ceil(N / 32)] each, bit (i & 31) of word (i >> 5) = segment i
std::tuple<at::Tensor, at::Tensor> verdict_fused(const at::Tensor& pts_rc, const at::Tensor& obs, const at::Tensor& obs_cnt,
                                                 double clearance, double bound, int64_t dot_mode, int64_t cmp_mode,
                                                 const c10::optional<at::Tensor>& seg_off, int64_t max_segs_per_map) {
    need(pts_rc, at::kDouble, "pts_rc"); need(obs, at::kDouble, "obs"); need(obs_cnt, at::kInt, "obs_cnt");
    same_device(pts_rc, obs, "obs"); same_device(pts_rc, obs_cnt, "obs_cnt");
    c10::cuda::CUDAGuard guard(pts_rc.device());
    const int64_t n = pts_rc.size(0), m = obs.size(0);
    const Grouping g = grouping(n, m, seg_off, max_segs_per_map, pts_rc);
    at::Tensor b64 = at::empty({(n + 31) / 32}, pts_rc.options().dtype(at::kInt)), b32 = at::empty_like(b64);
    ok(ppnet_verdict_fused(pts_rc.data_ptr<double>(), n, g.off, g.spm, m, obs.data_ptr<double>(), obs_cnt.data_ptr<int32_t>(),
                           (int32_t)obs.size(1), clearance, bound, (int32_t)dot_mode, (int32_t)cmp_mode, nullptr, nullptr,
                           (uint32_t*)b64.data_ptr<int32_t>(), (uint32_t*)b32.data_ptr<int32_t>(), stream()),
       "ppnet_verdict_fused");
    return {b64, b32};
}

at::Tensor dda_gridcheck_rc64(const at::Tensor& bits, const at::Tensor& segs_rc, const c10::optional<at::Tensor>& seg_off,
                              int64_t max_segs_per_map) {
    need(bits, at::kInt, "bits"); need(segs_rc, at::kDouble, "segs_rc");
    same_device(bits, segs_rc, "segs_rc");
    TORCH_CHECK(bits.dim() == 3, "ppnet_b200: bits must be [M, R, ceil(R/32)]");
    c10::cuda::CUDAGuard guard(bits.device());
    const int64_t n = segs_rc.size(0), m = bits.size(0);
    const Grouping g = grouping(n, m, seg_off, max_segs_per_map, bits);
    at::Tensor words = at::empty({(n + 31) / 32}, bits.options());
    ok(ppnet_dda_gridcheck_rc64((const uint32_t*)bits.data_ptr<int32_t>(), (int32_t)bits.size(1), m, segs_rc.data_ptr<double>(), n, g.off,
                                g.spm, nullptr, nullptr, (uint32_t*)words.data_ptr<int32_t>(), stream()),
       "ppnet_dda_gridcheck_rc64");
    return words;
}

at::Tensor dda_gridcheck(const at::Tensor& bits, const at::Tensor& segs_xy, const c10::optional<at::Tensor>& seg_off,
                         int64_t max_segs_per_map) {
    need(bits, at::kInt, "bits"); need(segs_xy, at::kFloat, "segs_xy");
    same_device(bits, segs_xy, "segs_xy");
    TORCH_CHECK(bits.dim() == 3, "ppnet_b200: bits must be [M, R, ceil(R/32)]");
    c10::cuda::CUDAGuard guard(bits.device());
    const int64_t n = segs_xy.size(0), m = bits.size(0);
    const Grouping g = grouping(n, m, seg_off, max_segs_per_map, bits);
    at::Tensor v = at::empty({n}, bits.options().dtype(at::kByte));
    ok(ppnet_dda_gridcheck((const uint32_t*)bits.data_ptr<int32_t>(), (int32_t)bits.size(1), m, segs_xy.data_ptr<float>(), n, g.off, g.spm,
                           v.data_ptr<uint8_t>(), nullptr, stream()),
       "ppnet_dda_gridcheck");
    return v;
}

// survivors of bit-packed verdicts -> (idx int32[n], count int64[1])
std::tuple<at::Tensor, at::Tensor> compact_bits(const at::Tensor& a, const c10::optional<at::Tensor>& b,
                                                const c10::optional<at::Tensor>& c, int64_t n, int64_t idx_base) {
    need(a, at::kInt, "a");
    if (b.has_value()) { need(*b, at::kInt, "b"); same_device(a, *b, "b"); TORCH_CHECK(b->numel() == a.numel(), "ppnet_b200: b has a different length"); }
    if (c.has_value()) { need(*c, at::kInt, "c"); same_device(a, *c, "c"); TORCH_CHECK(c->numel() == a.numel(), "ppnet_b200: c has a different length"); }
    TORCH_CHECK((n + 31) / 32 == a.numel(), "ppnet_b200: n does not match the number of words");
    c10::cuda::CUDAGuard guard(a.device());
    at::Tensor idx = at::empty({n}, a.options()), cnt = at::empty({1}, a.options().dtype(at::kLong));
    at::Tensor ws = at::empty({ppnet_compact_bits_workspace_elems(n)}, a.options().dtype(at::kLong));
    ok(ppnet_compact_bits((const uint32_t*)a.data_ptr<int32_t>(), b.has_value() ? (const uint32_t*)b->data_ptr<int32_t>() : nullptr,
                          c.has_value() ? (const uint32_t*)c->data_ptr<int32_t>() : nullptr, n, (int32_t)idx_base, idx.data_ptr<int32_t>(),
                          cnt.data_ptr<int64_t>(), ws.data_ptr<int64_t>(), stream()),
       "ppnet_compact_bits");
    return {idx, cnt};
}

at::Tensor grid_index_f64(const at::Tensor& pts, double map_size, double resolution, double mapoffset) {
    need(pts, at::kDouble, "pts");
    c10::cuda::CUDAGuard guard(pts.device());
    at::Tensor idx = at::empty(pts.sizes(), pts.options().dtype(at::kInt));
    ok(ppnet_grid_index_f64(pts.data_ptr<double>(), pts.numel(), map_size, resolution, mapoffset, idx.data_ptr<int32_t>(), stream()),
       "ppnet_grid_index_f64");
    return idx;
}

// -> (accept u8[M, O], out f64[M, O, 3], out_cnt i32[M])
std::tuple<at::Tensor, at::Tensor, at::Tensor> clearance_filter_f64(const at::Tensor& pathpt, const at::Tensor& cand, double map_size,
                                                                    double resolution, double clearance) {
    need(pathpt, at::kDouble, "pathpt"); need(cand, at::kDouble, "cand");
    same_device(pathpt, cand, "cand");
    TORCH_CHECK(pathpt.dim() == 3 && cand.dim() == 3 && pathpt.size(0) == cand.size(0), "ppnet_b200: pathpt [M, Np, 2], cand [M, O, 3]");
    c10::cuda::CUDAGuard guard(pathpt.device());
    const int64_t m = pathpt.size(0), O = cand.size(1);
    at::Tensor acc = at::empty({m, O}, pathpt.options().dtype(at::kByte)), out = at::zeros({m, O, 3}, pathpt.options());
    at::Tensor cnt = at::empty({m}, pathpt.options().dtype(at::kInt));
    ok(ppnet_clearance_filter_f64(pathpt.data_ptr<double>(), (int32_t)pathpt.size(1), cand.data_ptr<double>(), (int32_t)O, m, map_size,
                                  resolution, clearance, acc.data_ptr<uint8_t>(), out.data_ptr<double>(), cnt.data_ptr<int32_t>(), stream()),
       "ppnet_clearance_filter_f64");
    return {acc, out, cnt};
}

at::Tensor raster_circles_bits(const at::Tensor& obs, const at::Tensor& obs_cnt, int64_t resolution, double inflate) {
    need(obs, at::kDouble, "obs"); need(obs_cnt, at::kInt, "obs_cnt");
    same_device(obs, obs_cnt, "obs_cnt");
    c10::cuda::CUDAGuard guard(obs.device());
    at::Tensor bits = at::empty({obs.size(0), resolution, (resolution + 31) / 32}, obs.options().dtype(at::kInt));
    ok(ppnet_raster_circles_bits(obs.data_ptr<double>(), obs_cnt.data_ptr<int32_t>(), (int32_t)obs.size(1), obs.size(0), (int32_t)resolution,
                                 inflate, (uint32_t*)bits.data_ptr<int32_t>(), stream()),
       "ppnet_raster_circles_bits");
    return bits;
}

at::Tensor gmm_sample(int64_t seed, int64_t sample0, int64_t n, const at::Tensor& mean, const at::Tensor& stdv, const at::Tensor& weights) {
    need(mean, at::kFloat, "mean"); need(stdv, at::kFloat, "std"); need(weights, at::kFloat, "weights");
    same_device(mean, stdv, "std"); same_device(mean, weights, "weights");
    c10::cuda::CUDAGuard guard(mean.device());
    at::Tensor out = at::empty({n, mean.size(1)}, mean.options());
    ok(ppnet_gmm_sample((uint64_t)seed, (uint64_t)sample0, n, (int32_t)mean.size(0), (int32_t)mean.size(1), mean.data_ptr<float>(),
                        stdv.data_ptr<float>(), weights.data_ptr<float>(), out.data_ptr<float>(), nullptr, stream()),
       "ppnet_gmm_sample");
    return out;
}

// generator-mode segment source -> f64[n_maps * segs_per_map, 4]; `like` only supplies the device
at::Tensor propose_segments(const at::Tensor& like, int64_t seed, int64_t map0, int64_t n_maps, int64_t segs_per_map, double resolution,
                            double sigma) {
    TORCH_CHECK(like.is_cuda(), "ppnet_b200: `like` must be a CUDA tensor (it selects the device)");
    c10::cuda::CUDAGuard guard(like.device());
    at::Tensor out = at::empty({n_maps * segs_per_map, 4}, like.options().dtype(at::kDouble));
    ok(ppnet_propose_segments((uint64_t)seed, (uint64_t)map0, n_maps, segs_per_map, resolution, sigma, out.data_ptr<double>(), stream()),
       "ppnet_propose_segments");
    return out;
}

}  // namespace

TORCH_LIBRARY(ppnet_b200, m) {
    m.def("segcheck_edage_f64(Tensor pts_rc, Tensor obs, Tensor obs_cnt, float clearance, float bound=224., int dot_mode=0, "
          "Tensor? seg_off=None, int max_segs_per_map=0) -> Tensor");
    m.def("segcheck_mpnet_f32(Tensor pts_xy, Tensor obs, Tensor obs_cnt, float clearance, float bound=224., Tensor? seg_off=None, "
          "int max_segs_per_map=0) -> Tensor");
    m.def("verdict_fused(Tensor pts_rc, Tensor obs, Tensor obs_cnt, float clearance, float bound=224., int dot_mode=0, int cmp_mode=0, "
          "Tensor? seg_off=None, int max_segs_per_map=0) -> (Tensor, Tensor)");
    m.def("dda_gridcheck(Tensor bits, Tensor segs_xy, Tensor? seg_off=None, int max_segs_per_map=0) -> Tensor");
    m.def("dda_gridcheck_rc64(Tensor bits, Tensor segs_rc, Tensor? seg_off=None, int max_segs_per_map=0) -> Tensor");
    m.def("compact_bits(Tensor a, Tensor? b, Tensor? c, int n, int idx_base=0) -> (Tensor, Tensor)");
    m.def("grid_index_f64(Tensor pts, float map_size, float resolution, float mapoffset) -> Tensor");
    m.def("clearance_filter_f64(Tensor pathpt, Tensor cand, float map_size, float resolution, float clearance) -> (Tensor, Tensor, Tensor)");
    m.def("raster_circles_bits(Tensor obs, Tensor obs_cnt, int resolution, float inflate=0.) -> Tensor");
    m.def("gmm_sample(int seed, int sample0, int n, Tensor mean, Tensor std, Tensor weights) -> Tensor");
    m.def("propose_segments(Tensor like, int seed, int map0, int n_maps, int segs_per_map, float resolution=224., float sigma=15.) -> Tensor");
}

// CUDA key only: a CPU tensor finds no kernel and the dispatcher raises -- there is no CPU path
TORCH_LIBRARY_IMPL(ppnet_b200, CUDA, m) {
    m.impl("segcheck_edage_f64", segcheck_edage_f64);
    m.impl("segcheck_mpnet_f32", segcheck_mpnet_f32);
    m.impl("verdict_fused", verdict_fused);
    m.impl("dda_gridcheck", dda_gridcheck);
    m.impl("dda_gridcheck_rc64", dda_gridcheck_rc64);
    m.impl("compact_bits", compact_bits);
    m.impl("grid_index_f64", grid_index_f64);
    m.impl("clearance_filter_f64", clearance_filter_f64);
    m.impl("raster_circles_bits", raster_circles_bits);
    m.impl("gmm_sample", gmm_sample);
    m.impl("propose_segments", propose_segments);
}
