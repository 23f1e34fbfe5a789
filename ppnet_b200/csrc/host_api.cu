// Host-buffer entry points (the "e2e" boundary): HOST pointers in, HOST pointers out.
//
// A ppnet_ctx owns two CUDA streams and a grow-only device arena per pipeline slot.  Segment batches
// are cut at map boundaries into slices; slice k+1's host->device copy overlaps slice k's kernel and
// slice k-1's device->host copy (double buffering).  Host buffers may be pageable (copies then stage
// through the driver) or pinned (cudaHostRegister / torch pin_memory -- true async DMA).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace ppnet {

struct Slot {
    cudaStream_t st = nullptr;
    void* buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[6] = {0, 0, 0, 0, 0, 0};
};

struct Ctx {
    int device = 0;
    Slot slot[2];
    int64_t h2d_bytes = 0, d2h_bytes = 0;
};

static int slot_reserve(Slot& s, int i, size_t bytes) {
    if (bytes <= s.cap[i]) return PPNET_OK;
    if (s.buf[i]) {
        PPNET_CUDA(cudaStreamSynchronize(s.st));
        PPNET_CUDA(cudaFree(s.buf[i]));
        s.buf[i] = nullptr;
        s.cap[i] = 0;
    }
    const size_t want = bytes + bytes / 4 + 256;
    if (cudaMalloc(&s.buf[i], want) != cudaSuccess) {
        cudaGetLastError();
        set_error("host api: cudaMalloc of %zu bytes failed", want);
        return PPNET_E_NOMEM;
    }
    s.cap[i] = want;
    return PPNET_OK;
}

// slices of whole maps with at most ~max_segs segments each (at least one map)
struct Slice { int64_t m0, m1, s0, s1; };
static std::vector<Slice> make_slices(int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map,
                                      int64_t n_maps, int64_t max_segs) {
    std::vector<Slice> out;
    int64_t m = 0;
    while (m < n_maps) {
        const int64_t s0 = seg_off ? seg_off[m] : m * segs_per_map;
        int64_t m1 = m + 1;
        auto end_of = [&](int64_t mm) { return seg_off ? seg_off[mm] : mm * segs_per_map; };
        while (m1 < n_maps && end_of(m1 + 1) - s0 <= max_segs) ++m1;
        out.push_back({m, m1, s0, end_of(m1)});
        m = m1;
    }
    (void)n_segs;
    return out;
}

template <typename T, typename Launch>
static int segcheck_host(Ctx* c, const T* pts, int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map,
                         int64_t n_maps, const double* obs, const int32_t* obs_cnt, int32_t omax,
                         uint8_t* verdict, uint8_t* steer, Launch launch) {
    PPNET_REQUIRE(c, "host api: null context");
    PPNET_REQUIRE(n_segs >= 0 && n_maps >= 0, "host api: negative sizes");
    if (n_segs == 0 || n_maps == 0) return PPNET_OK;
    PPNET_REQUIRE(pts && obs_cnt && (verdict || steer), "host api: null pointer");
    PPNET_REQUIRE(seg_off || segs_per_map * n_maps == n_segs, "host api: bad uniform grouping");
    PPNET_CUDA(cudaSetDevice(c->device));
    const int64_t kSliceSegs = 1 << 20;
    std::vector<Slice> sl = make_slices(n_segs, seg_off, segs_per_map, n_maps, kSliceSegs);
    std::vector<int64_t> rel;     // slice-relative CSR offsets (host scratch, kept alive until sync)
    std::vector<std::vector<int64_t>> rel_keep(sl.size());
    for (size_t k = 0; k < sl.size(); ++k) {
        Slot& s = c->slot[k & 1];
        const Slice& q = sl[k];
        const int64_t ns = q.s1 - q.s0, nm = q.m1 - q.m0;
        int rc;
        if ((rc = slot_reserve(s, 0, sizeof(T) * 4 * (size_t)ns)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 1, sizeof(double) * 3 * (size_t)omax * nm)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 2, sizeof(int32_t) * (size_t)nm)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 3, (size_t)ns)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 4, (size_t)ns)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 5, sizeof(int64_t) * (size_t)(nm + 1))) != PPNET_OK) return rc;
        PPNET_CUDA(cudaMemcpyAsync(s.buf[0], pts + 4 * q.s0, sizeof(T) * 4 * (size_t)ns, cudaMemcpyHostToDevice, s.st));
        if (omax > 0)
            PPNET_CUDA(cudaMemcpyAsync(s.buf[1], obs + (size_t)q.m0 * omax * 3, sizeof(double) * 3 * (size_t)omax * nm,
                                       cudaMemcpyHostToDevice, s.st));
        PPNET_CUDA(cudaMemcpyAsync(s.buf[2], obs_cnt + q.m0, sizeof(int32_t) * (size_t)nm, cudaMemcpyHostToDevice, s.st));
        c->h2d_bytes += (int64_t)(sizeof(T) * 4 * ns + sizeof(double) * 3 * (size_t)omax * nm + 4 * nm);
        const int64_t* d_off = nullptr;
        int64_t spm = segs_per_map;
        if (seg_off) {
            rel_keep[k].resize(nm + 1);
            int64_t longest = 1;
            for (int64_t i = 0; i <= nm; ++i) rel_keep[k][i] = seg_off[q.m0 + i] - q.s0;
            for (int64_t i = 0; i < nm; ++i) longest = std::max(longest, rel_keep[k][i + 1] - rel_keep[k][i]);
            PPNET_CUDA(cudaMemcpyAsync(s.buf[5], rel_keep[k].data(), sizeof(int64_t) * (size_t)(nm + 1),
                                       cudaMemcpyHostToDevice, s.st));
            c->h2d_bytes += 8 * (nm + 1);
            d_off = (const int64_t*)s.buf[5];
            spm = longest;
        }
        rc = launch((const T*)s.buf[0], ns, d_off, spm, nm, (const double*)s.buf[1], (const int32_t*)s.buf[2],
                    verdict ? (uint8_t*)s.buf[3] : nullptr, steer ? (uint8_t*)s.buf[4] : nullptr, (void*)s.st);
        if (rc != PPNET_OK) return rc;
        if (verdict) PPNET_CUDA(cudaMemcpyAsync(verdict + q.s0, s.buf[3], (size_t)ns, cudaMemcpyDeviceToHost, s.st));
        if (steer) PPNET_CUDA(cudaMemcpyAsync(steer + q.s0, s.buf[4], (size_t)ns, cudaMemcpyDeviceToHost, s.st));
        c->d2h_bytes += (verdict ? ns : 0) + (steer ? ns : 0);
    }
    PPNET_CUDA(cudaStreamSynchronize(c->slot[0].st));
    PPNET_CUDA(cudaStreamSynchronize(c->slot[1].st));
    return PPNET_OK;
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_ctx_create(int32_t device, void** ctx) {
    PPNET_REQUIRE(ctx, "ctx_create: null out pointer");
    PPNET_CUDA(cudaSetDevice(device));
    Ctx* c = new Ctx();
    c->device = device;
    for (int i = 0; i < 2; ++i) PPNET_CUDA(cudaStreamCreateWithFlags(&c->slot[i].st, cudaStreamNonBlocking));
    *ctx = c;
    return PPNET_OK;
}

extern "C" int ppnet_ctx_destroy(void* ctx) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return PPNET_OK;
    cudaSetDevice(c->device);
    for (int i = 0; i < 2; ++i) {
        if (c->slot[i].st) { cudaStreamSynchronize(c->slot[i].st); cudaStreamDestroy(c->slot[i].st); }
        for (int j = 0; j < 6; ++j) if (c->slot[i].buf[j]) cudaFree(c->slot[i].buf[j]);
    }
    delete c;
    return PPNET_OK;
}

extern "C" int ppnet_ctx_bytes(void* ctx, int64_t* h2d, int64_t* d2h) {
    Ctx* c = (Ctx*)ctx;
    PPNET_REQUIRE(c, "ctx_bytes: null context");
    if (h2d) *h2d = c->h2d_bytes;
    if (d2h) *d2h = c->d2h_bytes;
    return PPNET_OK;
}

extern "C" int ppnet_segcheck_edage_f64_host(void* ctx, const double* pts_rc, int64_t n_segs, const int64_t* seg_off,
                                             int64_t segs_per_map, int64_t n_maps, const double* obs,
                                             const int32_t* obs_cnt, int32_t omax, double clearance, double bound,
                                             int32_t dot_mode, uint8_t* verdict) {
    return segcheck_host<double>((Ctx*)ctx, pts_rc, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, verdict,
                                 nullptr,
                                 [&](const double* p, int64_t ns, const int64_t* off, int64_t spm, int64_t nm,
                                     const double* o, const int32_t* oc, uint8_t* v, uint8_t*, void* st) {
                                     return ppnet_segcheck_edage_f64(p, ns, off, spm, nm, o, oc, omax, clearance, bound,
                                                                     dot_mode, v, st);
                                 });
}

extern "C" int ppnet_segcheck_mpnet_f32_host(void* ctx, const float* pts_xy, int64_t n_segs, const int64_t* seg_off,
                                             int64_t segs_per_map, int64_t n_maps, const double* obs,
                                             const int32_t* obs_cnt, int32_t omax, double clearance, double bound,
                                             uint8_t* verdict, uint8_t* steer) {
    return segcheck_host<float>((Ctx*)ctx, pts_xy, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, verdict,
                                steer,
                                [&](const float* p, int64_t ns, const int64_t* off, int64_t spm, int64_t nm,
                                    const double* o, const int32_t* oc, uint8_t* v, uint8_t* s, void* st) {
                                    return ppnet_segcheck_mpnet_f32(p, ns, off, spm, nm, o, oc, omax, clearance, bound,
                                                                    v, s, st);
                                });
}

extern "C" int ppnet_clearance_filter_f64_host(void* ctx, const double* pathpt, int32_t np, const double* cand,
                                               int32_t O, int64_t n_maps, double map_size, double resolution,
                                               double clearance, uint8_t* accept, double* out, int32_t* out_cnt) {
    Ctx* c = (Ctx*)ctx;
    PPNET_REQUIRE(c, "host api: null context");
    PPNET_REQUIRE(n_maps >= 0 && np >= 0 && O >= 0, "host api: negative sizes");
    if (n_maps == 0) return PPNET_OK;
    PPNET_REQUIRE(pathpt && (cand || O == 0), "host api: null pointer");
    PPNET_CUDA(cudaSetDevice(c->device));
    const int64_t kSliceMaps = 4096;
    int k = 0;
    for (int64_t m0 = 0; m0 < n_maps; m0 += kSliceMaps, ++k) {
        Slot& s = c->slot[k & 1];
        const int64_t nm = std::min(kSliceMaps, n_maps - m0);
        const size_t b_pp = sizeof(double) * 2 * (size_t)np * nm, b_cd = sizeof(double) * 3 * (size_t)O * nm;
        int rc;
        if ((rc = slot_reserve(s, 0, b_pp)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 1, b_cd)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 2, (size_t)O * nm)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 3, b_cd)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 4, sizeof(int32_t) * (size_t)nm)) != PPNET_OK) return rc;
        PPNET_CUDA(cudaMemcpyAsync(s.buf[0], pathpt + (size_t)m0 * np * 2, b_pp, cudaMemcpyHostToDevice, s.st));
        if (O) PPNET_CUDA(cudaMemcpyAsync(s.buf[1], cand + (size_t)m0 * O * 3, b_cd, cudaMemcpyHostToDevice, s.st));
        c->h2d_bytes += (int64_t)(b_pp + b_cd);
        rc = ppnet_clearance_filter_f64((const double*)s.buf[0], np, (const double*)s.buf[1], O, nm, map_size,
                                        resolution, clearance, (uint8_t*)s.buf[2], (double*)s.buf[3],
                                        (int32_t*)s.buf[4], (void*)s.st);
        if (rc != PPNET_OK) return rc;
        if (accept && O) PPNET_CUDA(cudaMemcpyAsync(accept + (size_t)m0 * O, s.buf[2], (size_t)O * nm, cudaMemcpyDeviceToHost, s.st));
        if (out && O) PPNET_CUDA(cudaMemcpyAsync(out + (size_t)m0 * O * 3, s.buf[3], b_cd, cudaMemcpyDeviceToHost, s.st));
        if (out_cnt) PPNET_CUDA(cudaMemcpyAsync(out_cnt + m0, s.buf[4], sizeof(int32_t) * (size_t)nm, cudaMemcpyDeviceToHost, s.st));
        c->d2h_bytes += (int64_t)((accept ? O * nm : 0) + (out ? b_cd : 0) + (out_cnt ? 4 * nm : 0));
    }
    PPNET_CUDA(cudaStreamSynchronize(c->slot[0].st));
    PPNET_CUDA(cudaStreamSynchronize(c->slot[1].st));
    return PPNET_OK;
}
