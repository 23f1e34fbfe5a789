// Host-buffer entry points (the "e2e" boundary): HOST pointers in, HOST pointers out.
//
// A ppnet_ctx owns two CUDA streams and a grow-only device arena per pipeline slot.  Segment batches
// are cut at map boundaries into slices; slice k+1's host->device copy overlaps slice k's kernel and
// slice k-1's device->host copy (double buffering).  Host buffers may be pageable (copies then stage
// through the driver) or pinned (cudaHostRegister / torch pin_memory -- true async DMA).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace ppnet {

constexpr int kSlotBufs = 12;   // see generate_host_run for the buffer ids
struct Slot {
    cudaStream_t st = nullptr;
    void* buf[kSlotBufs] = {};
    size_t cap[kSlotBufs] = {};
};

struct Ctx {
    int device = 0;
    Slot slot[2];
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    unsigned long long* pinned_cnt = nullptr;    // pinned staging for per-slice counters: a device->host copy into
    size_t pinned_cap = 0;                       // pageable memory would block the issuing thread and serialise slices
    char* pinned_small = nullptr;                // pinned staging for the small per-map outputs (one copy per slice
    size_t pinned_small_cap = 0;                 // instead of six tiny ones: a tiny DMA costs ~6 us of engine time)
    // generate / generate-and-check pipeline: one stream per direction + one for the kernels, kPipe slices in flight.
    // Uploads run back to back on st_up and downloads back to back on st_down (each PCIe direction stays busy), the
    // kernels of slice k wait for its upload, its download waits for its kernels; events guard the reuse of a slot.
    static constexpr int kPipe = 3;
    Slot pslot[kPipe];                           // buffers only (their .st stays null)
    cudaStream_t st_up = nullptr, st_cmp = nullptr, st_down = nullptr;
    cudaEvent_t ev_up[kPipe] = {}, ev_cmp[kPipe] = {}, ev_cnt[kPipe] = {}, ev_down[kPipe] = {};
};

// An entry point that fails half-way may still have copies into the caller's host buffers in flight on its streams: let
// them land before the caller sees the error (and possibly frees those buffers).
static int drain_on_error(Ctx* c, int rc) {
    if (rc != PPNET_OK && c) {
        for (int i = 0; i < 2; ++i) if (c->slot[i].st) cudaStreamSynchronize(c->slot[i].st);
        for (cudaStream_t st : {c->st_up, c->st_cmp, c->st_down}) if (st) cudaStreamSynchronize(st);
        cudaGetLastError();
    }
    return rc;
}

static int slot_reserve(Slot& s, int i, size_t bytes) {
    if (bytes <= s.cap[i]) return PPNET_OK;
    if (s.buf[i]) {
        if (s.st) PPNET_CUDA(cudaStreamSynchronize(s.st));
        else PPNET_CUDA(cudaDeviceSynchronize());          // pipeline slots are shared by three streams
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
static int segcheck_host_run(Ctx* c, const T* pts, int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map,
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

template <typename T, typename Launch>
static int segcheck_host(Ctx* c, const T* pts, int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map, int64_t n_maps,
                         const double* obs, const int32_t* obs_cnt, int32_t omax, uint8_t* verdict, uint8_t* steer, Launch launch) {
    return drain_on_error(c, segcheck_host_run<T>(c, pts, n_segs, seg_off, segs_per_map, n_maps, obs, obs_cnt, omax, verdict, steer, launch));
}

}  // namespace ppnet

using namespace ppnet;

extern "C" int ppnet_ctx_create(int32_t device, void** ctx) {
    PPNET_REQUIRE(ctx, "ctx_create: null out pointer");
    PPNET_CUDA(cudaSetDevice(device));
    Ctx* c = new Ctx();
    c->device = device;
    for (int i = 0; i < 2; ++i)
        if (cudaStreamCreateWithFlags(&c->slot[i].st, cudaStreamNonBlocking) != cudaSuccess) {
            set_error("ctx_create: cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
            ppnet_ctx_destroy(c);                            // frees whatever was created
            return PPNET_E_CUDA;
        }
    *ctx = c;
    return PPNET_OK;
}

extern "C" int ppnet_ctx_destroy(void* ctx) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return PPNET_OK;
    cudaSetDevice(c->device);
    if (c->pinned_cnt) cudaFreeHost(c->pinned_cnt);
    if (c->pinned_small) cudaFreeHost(c->pinned_small);
    for (cudaStream_t st : {c->st_up, c->st_cmp, c->st_down})
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (int i = 0; i < Ctx::kPipe; ++i) {
        for (cudaEvent_t e : {c->ev_up[i], c->ev_cmp[i], c->ev_cnt[i], c->ev_down[i]}) if (e) cudaEventDestroy(e);
        for (int j = 0; j < kSlotBufs; ++j) if (c->pslot[i].buf[j]) cudaFree(c->pslot[i].buf[j]);
    }
    for (int i = 0; i < 2; ++i) {
        if (c->slot[i].st) { cudaStreamSynchronize(c->slot[i].st); cudaStreamDestroy(c->slot[i].st); }
        for (int j = 0; j < kSlotBufs; ++j) if (c->slot[i].buf[j]) cudaFree(c->slot[i].buf[j]);
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

static int clearance_host_run(Ctx* c, const double* pathpt, int32_t np, const double* cand,
                              int32_t O, int64_t n_maps, double map_size, double resolution,
                              double clearance, uint8_t* accept, double* out, int32_t* out_cnt) {
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

extern "C" int ppnet_clearance_filter_f64_host(void* ctx, const double* pathpt, int32_t np, const double* cand,
                                               int32_t O, int64_t n_maps, double map_size, double resolution,
                                               double clearance, uint8_t* accept, double* out, int32_t* out_cnt) {
    return drain_on_error((Ctx*)ctx, clearance_host_run((Ctx*)ctx, pathpt, np, cand, O, n_maps, map_size, resolution, clearance, accept,
                                                        out, out_cnt));
}

// ------------------------------------------------------------------------------------------------
// generate / dda / gmm with host buffers
// ------------------------------------------------------------------------------------------------
namespace ppnet {
struct Bank {          // device copy of a target-path bank, owned by the context user
    int device = 0;
    double *pathpt = nullptr, *segpt = nullptr, *hull = nullptr, *obs = nullptr;
    int32_t *hull_cnt = nullptr, *obs_cnt = nullptr;
    int32_t n_bank = 0, np = 0, nseg1 = 0, hmax = 0, pomax = 0;
};
template <typename T>
static int upload(T** dst, const T* src, size_t n) {
    *dst = nullptr;
    if (n == 0) return PPNET_OK;
    if (cudaMalloc((void**)dst, n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); set_error("bank: cudaMalloc failed"); return PPNET_E_NOMEM; }
    PPNET_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return PPNET_OK;
}
}  // namespace ppnet

extern "C" int ppnet_bank_upload(int32_t device, const double* pathpt, const double* segpt, const double* hull,
                                 const int32_t* hull_cnt, const double* obs, const int32_t* obs_cnt, int32_t n_bank,
                                 int32_t np, int32_t nseg1, int32_t hmax, int32_t pomax, void** bank) {
    PPNET_REQUIRE(bank && pathpt && hull && hull_cnt, "bank_upload: null pointer");
    PPNET_REQUIRE(n_bank > 0 && np >= 0 && nseg1 >= 0 && hmax > 0 && pomax >= 0, "bank_upload: bad sizes");
    PPNET_REQUIRE(pomax == 0 || (obs && obs_cnt), "bank_upload: obs is null");
    PPNET_CUDA(cudaSetDevice(device));
    Bank* b = new Bank();
    b->device = device; b->n_bank = n_bank; b->np = np; b->nseg1 = nseg1; b->hmax = hmax; b->pomax = pomax;
    int rc;
    if ((rc = upload(&b->pathpt, pathpt, (size_t)n_bank * np * 2)) != PPNET_OK ||
        (rc = upload(&b->segpt, segpt, segpt ? (size_t)n_bank * nseg1 * 2 : 0)) != PPNET_OK ||
        (rc = upload(&b->hull, hull, (size_t)n_bank * hmax * 2)) != PPNET_OK ||
        (rc = upload(&b->hull_cnt, hull_cnt, (size_t)n_bank)) != PPNET_OK ||
        (rc = upload(&b->obs, obs, (size_t)n_bank * pomax * 3)) != PPNET_OK ||
        (rc = upload(&b->obs_cnt, obs_cnt, pomax ? (size_t)n_bank : 0)) != PPNET_OK) {
        ppnet_bank_free(b);                                  // a partly built bank does not leak its device arrays
        return rc;
    }
    *bank = b;
    return PPNET_OK;
}

extern "C" int ppnet_bank_free(void* bank) {
    Bank* b = (Bank*)bank;
    if (!b) return PPNET_OK;
    cudaSetDevice(b->device);
    cudaFree(b->pathpt); cudaFree(b->segpt); cudaFree(b->hull); cudaFree(b->hull_cnt); cudaFree(b->obs); cudaFree(b->obs_cnt);
    delete b;
    return PPNET_OK;
}

// `p` carries the generation settings and HOST output pointers (bank_* / in_* fields are ignored: the bank
// comes from `bank`, draws from Philox).  Maps are produced in slices on two streams: slice k's device->host
// copies overlap slice k+1's host->device copies and kernels.  With `io`, every slice also uploads (or draws on the
// device) its candidate segments and runs the verdict kernels against the maps it has just generated -- the obstacle
// sets and bitmaps never leave the device in between (ppnet_generate_and_check_host).
//
// One-array mode (io->segs_xy_f32 == NULL): ONE float64 upload per segment feeds all three verdicts -- the fused
// A11 + A12 kernel and the DDA read it directly and apply the float32 cast + swap themselves; verdicts come back
// bit-packed and / or as bytes; the survivors (free segments, valid maps) are compacted on the device and only
// `count` indices travel.  Their counts are known one slice late, so slice k's index lists are copied after slice
// k+1's kernels have been enqueued (the device never waits for the host).
static int generate_host_run(Ctx* c, Bank* b, const ppnet_gen_params* p, const ppnet_pipeline_io* io) {
    const int64_t spm = io ? io->segs_per_map : 0;
    const bool propose = io && !io->segs_rc_f64 && io->propose_sigma > 0.0;
    const bool one_array = io && !io->segs_xy_f32;
    const bool want_bits_out = io && (io->vbits_f64 || io->vbits_f32 || io->vbits_dda);
    const bool want_free = io && io->free_idx && io->free_count;
    const bool want_valid = io && io->valid_idx && io->valid_count;
    const bool any64 = io && (io->verdict_f64 || io->vbits_f64), any32 = io && (io->verdict_f32 || io->vbits_f32);
    const bool anydda = io && (io->verdict_dda || io->vbits_dda);
    if (io) {
        PPNET_REQUIRE(spm > 0, "generate_and_check_host: segs_per_map must be positive");
        PPNET_REQUIRE(!(io->segs_rc_f64 && io->propose_sigma > 0.0), "generate_and_check_host: segs_rc_f64 and propose_sigma are exclusive");
        PPNET_REQUIRE(!any64 || io->segs_rc_f64 || propose, "generate_and_check_host: the A11 verdict needs segs_rc_f64 (or a device proposal)");
        PPNET_REQUIRE(!(any32 || anydda) || io->segs_xy_f32 || io->segs_rc_f64 || propose,
                      "generate_and_check_host: the A12 / DDA verdicts need segments");
        PPNET_REQUIRE(one_array || !(want_bits_out || want_free),
                      "generate_and_check_host: bit-packed verdicts / survivor lists need the one-array mode (segs_xy_f32 == NULL)");
        PPNET_REQUIRE(!propose || one_array, "generate_and_check_host: a device proposal excludes segs_xy_f32");
        PPNET_REQUIRE(!io->out_segs_rc || propose, "generate_and_check_host: out_segs_rc is the output of a device proposal");
        PPNET_REQUIRE(!want_free || any64 || any32 || anydda, "generate_and_check_host: free_idx needs at least one verdict");
        PPNET_REQUIRE((io->free_idx == nullptr) == (io->free_count == nullptr) && (io->valid_idx == nullptr) == (io->valid_count == nullptr),
                      "generate_and_check_host: index lists come with their counts");
        PPNET_REQUIRE(p->n_maps * spm <= 2147483647LL, "generate_and_check_host: more than 2^31 segments in one call");
    }
    const int R = (int)p->resolution, W = (R + 31) / 32;
    const int64_t O = p->obstacles_num, oo = O + b->pomax;
    static const int64_t slice_env = getenv("PPNET_HOST_SLICE") ? atoll(getenv("PPNET_HOST_SLICE")) : 0;   // dev knob
    // measured (10 k maps per call, two calls in flight): 1024 / 2048 / 4096 maps per slice -> 8.0 / 7.6 / 7.2 ms per call;
    // a third of the call per slice keeps the three pipeline stages busy for smaller calls
    const int64_t auto_slice = std::min<int64_t>(4096, std::max<int64_t>(1024, ((p->n_maps + 2) / 3 + 511) / 512 * 512));
    const int64_t kSlice = slice_env > 0 ? slice_env : auto_slice;
    // per-map byte sizes of the outputs, slot buffer ids: 0 pathpt, 1 segpt, 2 obs, 3 bits, 4 small ints, 5 counters,
    // 6 segments f64, 7 segments f32, 8 / 9 / 10 verdict bytes, 11 verdict words x3 + lists + workspace
    const size_t b_pp = sizeof(double) * 2 * (size_t)b->np, b_sp = sizeof(double) * 2 * (size_t)b->nseg1;
    const size_t b_ob = sizeof(double) * 3 * (size_t)oo, b_bt = (size_t)R * W * 4;
    const size_t b_small = 8 /*angle*/ + 8 /*trans*/ + 4 * 3 /*obs_cnt, rand_cnt, tries*/ + 4 /*valid, padded*/;
    const int64_t n_slices = (p->n_maps + kSlice - 1) / kSlice;
    // pinned staging per slice: 4 generator counters + free count + valid count (6 words, padded to 8)
    const size_t cnt_words = 8 * (size_t)std::max<int64_t>(n_slices, 1);
    if (c->pinned_cap < cnt_words) {
        if (c->pinned_cnt) cudaFreeHost(c->pinned_cnt);
        c->pinned_cnt = nullptr; c->pinned_cap = 0;
        PPNET_CUDA(cudaHostAlloc((void**)&c->pinned_cnt, sizeof(unsigned long long) * (cnt_words + 64), cudaHostAllocDefault));
        c->pinned_cap = cnt_words + 64;
    }
    unsigned long long* cnt_keep = c->pinned_cnt;
    for (size_t i = 0; i < cnt_words; ++i) cnt_keep[i] = 0ull;
    const bool want_small = p->out_angle || p->out_trans || p->out_obs_cnt || p->out_rand_cnt || p->out_tries || p->out_valid;
    if (want_small && c->pinned_small_cap < b_small * (size_t)p->n_maps) {
        if (c->pinned_small) cudaFreeHost(c->pinned_small);
        c->pinned_small = nullptr; c->pinned_small_cap = 0;
        const size_t want = b_small * (size_t)p->n_maps + b_small * (size_t)p->n_maps / 4 + 4096;
        PPNET_CUDA(cudaHostAlloc((void**)&c->pinned_small, want, cudaHostAllocDefault));
        c->pinned_small_cap = want;
    }
    constexpr int kPipe = Ctx::kPipe;
    if (!c->st_up) {
        PPNET_CUDA(cudaStreamCreateWithFlags(&c->st_up, cudaStreamNonBlocking));
        PPNET_CUDA(cudaStreamCreateWithFlags(&c->st_cmp, cudaStreamNonBlocking));
        PPNET_CUDA(cudaStreamCreateWithFlags(&c->st_down, cudaStreamNonBlocking));
        for (int i = 0; i < kPipe; ++i)
            for (cudaEvent_t* e : {&c->ev_up[i], &c->ev_cmp[i], &c->ev_cnt[i], &c->ev_down[i]})
                PPNET_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    cudaStream_t s_up = c->st_up, s_cmp = c->st_cmp, s_down = c->st_down;
    const bool need_bits = p->out_bits || anydda;
    const bool dev_words = one_array && (want_bits_out || want_free);
    int64_t free_total = 0, valid_total = 0;

    // second half of a slice: its index lists, once its counts have landed in the pinned staging area
    auto finish_lists = [&](int64_t k) -> int {
        if (!want_free && !want_valid) return PPNET_OK;                   // (ev_down was recorded right after the bulk copies)
        Slot& s = c->pslot[k % kPipe];
        PPNET_CUDA(cudaEventSynchronize(c->ev_cnt[k % kPipe]));
        const int64_t nm = std::min(kSlice, p->n_maps - k * kSlice), ns = nm * spm;
        const size_t words = (size_t)((ns + 31) / 32);
        char* wbase = (char*)s.buf[11];
        if (want_free) {
            const int64_t cnt = (int64_t)cnt_keep[8 * k + 4];
            const int32_t* d_idx = (const int32_t*)(wbase + 3 * 4 * words + 64);
            if (cnt > 0) {
                PPNET_CUDA(cudaMemcpyAsync(io->free_idx + free_total, d_idx, 4 * (size_t)cnt, cudaMemcpyDeviceToHost, s_down));
                c->d2h_bytes += 4 * cnt;
            }
            free_total += cnt;
        }
        if (want_valid) {
            const int64_t cnt = (int64_t)cnt_keep[8 * k + 5];
            const int32_t* d_idx = (const int32_t*)(wbase + 3 * 4 * words + 64 + (want_free ? 4 * (size_t)ns : 0));
            if (cnt > 0) {
                PPNET_CUDA(cudaMemcpyAsync(io->valid_idx + valid_total, d_idx, 4 * (size_t)cnt, cudaMemcpyDeviceToHost, s_down));
                c->d2h_bytes += 4 * cnt;
            }
            valid_total += cnt;
        }
        PPNET_CUDA(cudaEventRecord(c->ev_down[k % kPipe], s_down));       // slot k % kPipe is free for slice k + kPipe
        return PPNET_OK;
    };

    static const bool dbg = getenv("PPNET_HOST_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_lists = 0.0;
    int64_t k = 0;
    for (int64_t m0 = 0; m0 < p->n_maps; m0 += kSlice, ++k) {
        const int sb = (int)(k % kPipe);
        Slot& s = c->pslot[sb];
        const int64_t nm = std::min(kSlice, p->n_maps - m0);
        const int64_t ns = nm * spm;
        const size_t words = (size_t)((ns + 31) / 32);
        int rc;
        if ((rc = slot_reserve(s, 0, b_pp * nm)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 1, b_sp * nm + 16)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 2, b_ob * nm + 16)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 3, b_bt * nm + 16)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 4, b_small * nm + 64)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 5, 64)) != PPNET_OK) return rc;
        if (io) {
            if ((io->segs_rc_f64 || propose) && (rc = slot_reserve(s, 6, 32 * (size_t)ns)) != PPNET_OK) return rc;
            if (io->segs_xy_f32 && (rc = slot_reserve(s, 7, 16 * (size_t)ns)) != PPNET_OK) return rc;
            for (int v = 8; v <= 10; ++v)
                if ((rc = slot_reserve(s, v, (size_t)ns)) != PPNET_OK) return rc;
            // words x3 | counts (64 B) | free list | valid list | compaction workspace
            const size_t ws_elems = (size_t)ppnet_compact_bits_workspace_elems(ns);
            if ((rc = slot_reserve(s, 11, 3 * 4 * words + 64 + (want_free ? 4 * (size_t)ns : 0) + 4 * (size_t)nm + 64 + 8 * ws_elems)) != PPNET_OK)
                return rc;
            // uploads: back to back on the upload stream; the slot's segment buffers were last read by slice k - kPipe
            if (k >= kPipe) PPNET_CUDA(cudaStreamWaitEvent(s_up, c->ev_cmp[sb], 0));
            if (io->segs_rc_f64) {
                PPNET_CUDA(cudaMemcpyAsync(s.buf[6], io->segs_rc_f64 + 4 * m0 * spm, 32 * (size_t)ns, cudaMemcpyHostToDevice, s_up));
                c->h2d_bytes += 32 * ns;
            }
            if (io->segs_xy_f32) {
                PPNET_CUDA(cudaMemcpyAsync(s.buf[7], io->segs_xy_f32 + 4 * m0 * spm, 16 * (size_t)ns, cudaMemcpyHostToDevice, s_up));
                c->h2d_bytes += 16 * ns;
            }
            PPNET_CUDA(cudaEventRecord(c->ev_up[sb], s_up));
        }
        // kernels: wait for this slice's upload and for the download of the slice that used the slot's output buffers
        if (io) PPNET_CUDA(cudaStreamWaitEvent(s_cmp, c->ev_up[sb], 0));
        if (k >= kPipe) PPNET_CUDA(cudaStreamWaitEvent(s_cmp, c->ev_down[sb], 0));
        if (propose &&
            (rc = ppnet_propose_segments(p->seed, (uint64_t)(p->map0 + m0), nm, spm, p->resolution, io->propose_sigma,
                                         (double*)s.buf[6], (void*)s_cmp)) != PPNET_OK) return rc;
        char* small = (char*)s.buf[4];
        double* d_angle = (double*)small;
        int32_t* d_trans = (int32_t*)(small + 8 * nm);
        int32_t* d_ocnt = (int32_t*)(small + 16 * nm);
        int32_t* d_rcnt = (int32_t*)(small + 20 * nm);
        int32_t* d_tries = (int32_t*)(small + 24 * nm);
        uint8_t* d_valid = (uint8_t*)(small + 28 * nm);
        PPNET_CUDA(cudaMemsetAsync(s.buf[5], 0, 32, s_cmp));
        ppnet_gen_params q = *p;
        q.bank_pathpt = b->pathpt; q.bank_segpt = b->segpt; q.bank_hull = b->hull; q.bank_hull_cnt = b->hull_cnt;
        q.bank_obs = b->obs; q.bank_obs_cnt = b->obs_cnt;
        q.n_bank = b->n_bank; q.np = b->np; q.nseg1 = b->nseg1; q.hmax = b->hmax; q.pomax = b->pomax;
        q.map0 = p->map0 + m0; q.n_maps = nm;
        q.out_pathpt = p->out_pathpt ? (double*)s.buf[0] : nullptr;
        q.out_segpt = p->out_segpt ? (double*)s.buf[1] : nullptr;
        q.out_obs = (double*)s.buf[2];
        q.out_bits = need_bits ? (uint32_t*)s.buf[3] : nullptr;
        q.out_angle = d_angle; q.out_trans = d_trans; q.out_obs_cnt = d_ocnt; q.out_rand_cnt = d_rcnt;
        q.out_tries = d_tries; q.out_valid = d_valid;
        q.counters = (unsigned long long*)s.buf[5];
        if ((rc = ppnet_generate_maps(&q, (void*)s_cmp)) != PPNET_OK) return rc;
        char* wbase = io ? (char*)s.buf[11] : nullptr;
        uint32_t* w64 = (uint32_t*)wbase;
        uint32_t* w32 = (uint32_t*)(wbase + 4 * words);
        uint32_t* wdd = (uint32_t*)(wbase + 8 * words);
        int64_t* d_counts = (int64_t*)(wbase + 12 * words);                   // [0] free, [1] valid (64 B, 8-aligned: words*12 % 8 == 0 or 4)
        if (io && ((12 * words) & 7)) d_counts = (int64_t*)(wbase + 12 * words + 4);
        int32_t* d_free = (int32_t*)(wbase + 12 * words + 64);
        int32_t* d_vidx = (int32_t*)(wbase + 12 * words + 64 + (want_free ? 4 * (size_t)ns : 0));
        int64_t* d_ws = (int64_t*)(((uintptr_t)(d_vidx + nm) + 63) & ~(uintptr_t)63);
        if (io && one_array) {
            if ((any64 || any32) &&
                (rc = ppnet_verdict_fused((const double*)s.buf[6], ns, nullptr, spm, nm, (const double*)s.buf[2], d_ocnt, (int32_t)oo,
                                          io->clearance_px, io->bound, io->dot_mode, io->cmp_mode,
                                          io->verdict_f64 ? (uint8_t*)s.buf[8] : nullptr, io->verdict_f32 ? (uint8_t*)s.buf[9] : nullptr,
                                          (any64 && dev_words) ? w64 : nullptr, (any32 && dev_words) ? w32 : nullptr,
                                          (void*)s_cmp)) != PPNET_OK) return rc;
            if (anydda &&
                (rc = ppnet_dda_gridcheck_rc64((const uint32_t*)s.buf[3], R, nm, (const double*)s.buf[6], ns, nullptr, spm,
                                               io->verdict_dda ? (uint8_t*)s.buf[10] : nullptr, nullptr, dev_words ? wdd : nullptr,
                                               (void*)s_cmp)) != PPNET_OK) return rc;
            if (want_free) {
                const uint32_t* arr[3]; int na = 0;
                if (any64) arr[na++] = w64;
                if (any32) arr[na++] = w32;
                if (anydda) arr[na++] = wdd;
                if ((rc = ppnet_compact_bits(arr[0], na > 1 ? arr[1] : nullptr, na > 2 ? arr[2] : nullptr, ns, (int32_t)(m0 * spm), d_free,
                                             d_counts, d_ws, (void*)s_cmp)) != PPNET_OK) return rc;
            }
        } else if (io) {
            if (io->verdict_f64 &&
                (rc = ppnet_segcheck_edage_f64((const double*)s.buf[6], ns, nullptr, spm, nm, (const double*)s.buf[2], d_ocnt,
                                               (int32_t)oo, io->clearance_px, io->bound, io->dot_mode, (uint8_t*)s.buf[8],
                                               (void*)s_cmp)) != PPNET_OK) return rc;
            if (io->verdict_f32 &&
                (rc = ppnet_segcheck_mpnet_f32((const float*)s.buf[7], ns, nullptr, spm, nm, (const double*)s.buf[2], d_ocnt,
                                               (int32_t)oo, io->clearance_px, io->bound, (uint8_t*)s.buf[9], nullptr,
                                               (void*)s_cmp)) != PPNET_OK) return rc;
            if (io->verdict_dda &&
                (rc = ppnet_dda_gridcheck((const uint32_t*)s.buf[3], R, nm, (const float*)s.buf[7], ns, nullptr, spm,
                                          (uint8_t*)s.buf[10], nullptr, (void*)s_cmp)) != PPNET_OK) return rc;
        }
        if (want_valid &&
            (rc = ppnet_compact_u8_i32(d_valid, nm, 1, (int32_t)m0, d_vidx, d_counts + 1, (void*)s_cmp)) != PPNET_OK) return rc;
#define PPNET_D2H(host, devp, bytes)                                                                        \
    if (host) {                                                                                             \
        PPNET_CUDA(cudaMemcpyAsync((char*)(host), devp, (bytes), cudaMemcpyDeviceToHost, s_down));            \
        c->d2h_bytes += (int64_t)(bytes);                                                                   \
    }
        PPNET_CUDA(cudaEventRecord(c->ev_cmp[sb], s_cmp));
        PPNET_CUDA(cudaStreamWaitEvent(s_down, c->ev_cmp[sb], 0));
        if (want_free || want_valid) {          // the counts first: the host needs them to size the list copies
            PPNET_CUDA(cudaMemcpyAsync(cnt_keep + 8 * k + 4, d_counts, 16, cudaMemcpyDeviceToHost, s_down));
            c->d2h_bytes += 16;
        }
        if (p->counters)
            PPNET_CUDA(cudaMemcpyAsync(cnt_keep + 8 * k, s.buf[5], 32, cudaMemcpyDeviceToHost, s_down));
        PPNET_CUDA(cudaEventRecord(c->ev_cnt[sb], s_down));
        PPNET_D2H(p->out_pathpt ? (char*)p->out_pathpt + b_pp * m0 : nullptr, s.buf[0], b_pp * nm);
        PPNET_D2H(p->out_segpt ? (char*)p->out_segpt + b_sp * m0 : nullptr, s.buf[1], b_sp * nm);
        PPNET_D2H(p->out_obs ? (char*)p->out_obs + b_ob * m0 : nullptr, s.buf[2], b_ob * nm);
        PPNET_D2H(p->out_bits ? (char*)p->out_bits + b_bt * m0 : nullptr, s.buf[3], b_bt * nm);
        // angle | trans | obs_cnt | rand_cnt | tries | valid of this slice: ONE copy into pinned staging, scattered below
        PPNET_D2H(want_small ? c->pinned_small + b_small * (size_t)m0 : nullptr, small, 29 * (size_t)nm);
        if (io) {
            PPNET_D2H(io->verdict_f64 ? io->verdict_f64 + m0 * spm : nullptr, s.buf[8], (size_t)ns);
            PPNET_D2H(io->verdict_f32 ? io->verdict_f32 + m0 * spm : nullptr, s.buf[9], (size_t)ns);
            PPNET_D2H(io->verdict_dda ? io->verdict_dda + m0 * spm : nullptr, s.buf[10], (size_t)ns);
            // slices start on word boundaries: kSlice * spm is a multiple of 32
            PPNET_D2H(io->vbits_f64 ? io->vbits_f64 + (m0 * spm) / 32 : nullptr, w64, 4 * words);
            PPNET_D2H(io->vbits_f32 ? io->vbits_f32 + (m0 * spm) / 32 : nullptr, w32, 4 * words);
            PPNET_D2H(io->vbits_dda ? io->vbits_dda + (m0 * spm) / 32 : nullptr, wdd, 4 * words);
            PPNET_D2H(io->out_segs_rc ? io->out_segs_rc + 4 * m0 * spm : nullptr, s.buf[6], 32 * (size_t)ns);
        }
        if (!want_free && !want_valid) PPNET_CUDA(cudaEventRecord(c->ev_down[sb], s_down));
        const double tl = now();
        if (k > 0 && (rc = finish_lists(k - 1)) != PPNET_OK) return rc;
        t_lists += now() - tl;
    }
    const double t_issued = now();
    if (k > 0) { int rc = finish_lists(k - 1); if (rc != PPNET_OK) return rc; }
    PPNET_CUDA(cudaStreamSynchronize(s_down));                           // every slice ends on the download stream
    PPNET_CUDA(cudaStreamSynchronize(s_cmp));
    PPNET_CUDA(cudaStreamSynchronize(s_up));
    if (dbg)
        fprintf(stderr, "[ppnet host] %lld slices: issue %.2f ms (of which waiting for list counts %.2f), final sync %.2f ms\n",
                (long long)k, t_issued - t_begin, t_lists, now() - t_issued);
    if (p->counters)
        for (int64_t q = 0; q < n_slices; ++q)
            for (int i = 0; i < 4; ++i) p->counters[i] += cnt_keep[8 * q + i];
    if (want_small)
        for (int64_t m0 = 0; m0 < p->n_maps; m0 += kSlice) {
            const size_t nm = (size_t)std::min(kSlice, p->n_maps - m0);
            const char* blk = c->pinned_small + b_small * (size_t)m0;
            if (p->out_angle) memcpy(p->out_angle + m0, blk, 8 * nm);
            if (p->out_trans) memcpy(p->out_trans + 2 * m0, blk + 8 * nm, 8 * nm);
            if (p->out_obs_cnt) memcpy(p->out_obs_cnt + m0, blk + 16 * nm, 4 * nm);
            if (p->out_rand_cnt) memcpy(p->out_rand_cnt + m0, blk + 20 * nm, 4 * nm);
            if (p->out_tries) memcpy(p->out_tries + m0, blk + 24 * nm, 4 * nm);
            if (p->out_valid) memcpy(p->out_valid + m0, blk + 28 * nm, nm);
        }
    if (want_free) *io->free_count = free_total;
    if (want_valid) *io->valid_count = valid_total;
    return PPNET_OK;
}

static int generate_host_impl(Ctx* c, Bank* b, const ppnet_gen_params* p, const ppnet_pipeline_io* io) {
    PPNET_REQUIRE(c && b && p, "generate_maps_host: null argument");
    PPNET_REQUIRE(p->n_maps >= 0, "generate_maps_host: negative n_maps");
    PPNET_REQUIRE(!p->in_angle && !p->in_trans && !p->in_cand, "generate_maps_host: caller-supplied draws need the device API");
    PPNET_CUDA(cudaSetDevice(c->device));
    return drain_on_error(c, generate_host_run(c, b, p, io));
}

extern "C" int ppnet_generate_maps_host(void* ctx, void* bank, const ppnet_gen_params* p) {
    return generate_host_impl((Ctx*)ctx, (Bank*)bank, p, nullptr);
}

extern "C" int ppnet_generate_and_check_host(void* ctx, void* bank, const ppnet_gen_params* p,
                                             const ppnet_pipeline_io* io) {
    PPNET_REQUIRE(io, "generate_and_check_host: null io");
    return generate_host_impl((Ctx*)ctx, (Bank*)bank, p, io);
}

static int dda_host_run(Ctx* c, const uint32_t* bits, int32_t resolution, int64_t n_maps,
                        const float* segs_xy, int64_t n_segs, const int64_t* seg_off,
                        int64_t segs_per_map, uint8_t* verdict, int32_t* first_hit) {
    PPNET_REQUIRE(c, "dda_host: null context");
    PPNET_REQUIRE(n_maps >= 0 && n_segs >= 0 && resolution > 0, "dda_host: bad sizes");
    if (n_maps == 0 || n_segs == 0) return PPNET_OK;
    PPNET_REQUIRE(bits && segs_xy && verdict, "dda_host: null pointer");
    PPNET_REQUIRE(seg_off || segs_per_map * n_maps == n_segs, "dda_host: bad uniform grouping");
    PPNET_CUDA(cudaSetDevice(c->device));
    const int W = (resolution + 31) / 32;
    const size_t bm = (size_t)resolution * W * 4;
    std::vector<Slice> sl = make_slices(n_segs, seg_off, segs_per_map, n_maps, 1 << 20);
    std::vector<std::vector<int64_t>> rel_keep(sl.size());
    for (size_t k = 0; k < sl.size(); ++k) {
        Slot& s = c->slot[k & 1];
        const Slice& q = sl[k];
        const int64_t ns = q.s1 - q.s0, nm = q.m1 - q.m0;
        int rc;
        if ((rc = slot_reserve(s, 0, 16 * (size_t)ns)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 1, bm * nm)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 3, (size_t)ns)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 4, 4 * (size_t)ns)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 5, 8 * (size_t)(nm + 1))) != PPNET_OK) return rc;
        PPNET_CUDA(cudaMemcpyAsync(s.buf[0], segs_xy + 4 * q.s0, 16 * (size_t)ns, cudaMemcpyHostToDevice, s.st));
        PPNET_CUDA(cudaMemcpyAsync(s.buf[1], (const char*)bits + bm * q.m0, bm * nm, cudaMemcpyHostToDevice, s.st));
        c->h2d_bytes += (int64_t)(16 * ns + bm * nm);
        const int64_t* d_off = nullptr;
        int64_t spm = segs_per_map;
        if (seg_off) {
            rel_keep[k].resize(nm + 1);
            int64_t longest = 1;
            for (int64_t i = 0; i <= nm; ++i) rel_keep[k][i] = seg_off[q.m0 + i] - q.s0;
            for (int64_t i = 0; i < nm; ++i) longest = std::max(longest, rel_keep[k][i + 1] - rel_keep[k][i]);
            PPNET_CUDA(cudaMemcpyAsync(s.buf[5], rel_keep[k].data(), 8 * (size_t)(nm + 1), cudaMemcpyHostToDevice, s.st));
            d_off = (const int64_t*)s.buf[5];
            spm = longest;
        }
        rc = ppnet_dda_gridcheck((const uint32_t*)s.buf[1], resolution, nm, (const float*)s.buf[0], ns, d_off, spm,
                                 (uint8_t*)s.buf[3], first_hit ? (int32_t*)s.buf[4] : nullptr, (void*)s.st);
        if (rc != PPNET_OK) return rc;
        PPNET_CUDA(cudaMemcpyAsync(verdict + q.s0, s.buf[3], (size_t)ns, cudaMemcpyDeviceToHost, s.st));
        if (first_hit) PPNET_CUDA(cudaMemcpyAsync(first_hit + q.s0, s.buf[4], 4 * (size_t)ns, cudaMemcpyDeviceToHost, s.st));
        c->d2h_bytes += ns + (first_hit ? 4 * ns : 0);
    }
    PPNET_CUDA(cudaStreamSynchronize(c->slot[0].st));
    PPNET_CUDA(cudaStreamSynchronize(c->slot[1].st));
    return PPNET_OK;
}

extern "C" int ppnet_dda_gridcheck_host(void* ctx, const uint32_t* bits, int32_t resolution, int64_t n_maps,
                                        const float* segs_xy, int64_t n_segs, const int64_t* seg_off,
                                        int64_t segs_per_map, uint8_t* verdict, int32_t* first_hit) {
    return drain_on_error((Ctx*)ctx, dda_host_run((Ctx*)ctx, bits, resolution, n_maps, segs_xy, n_segs, seg_off, segs_per_map, verdict,
                                                  first_hit));
}

static int gmm_host_run(Ctx* c, uint64_t seed, uint64_t sample0, int64_t n, int32_t order, int32_t dim,
                        const float* mean, const float* stdv, const float* weights, float* out) {
    PPNET_REQUIRE(c && mean && stdv && weights && (out || n == 0), "gmm_sample_host: null argument");
    PPNET_REQUIRE(n >= 0 && order > 0 && dim > 0, "gmm_sample_host: bad sizes");
    if (n == 0) return PPNET_OK;
    PPNET_CUDA(cudaSetDevice(c->device));
    const int64_t kSlice = 1 << 21;
    int k = 0;
    for (int64_t i0 = 0; i0 < n; i0 += kSlice, ++k) {
        Slot& s = c->slot[k & 1];
        const int64_t ni = std::min(kSlice, n - i0);
        int rc;
        if ((rc = slot_reserve(s, 0, 4 * (size_t)ni * dim)) != PPNET_OK) return rc;
        if ((rc = slot_reserve(s, 1, 4 * (size_t)order * (2 * dim + 1) + 64)) != PPNET_OK) return rc;
        float* d_mean = (float*)s.buf[1];
        float* d_std = d_mean + order * dim;
        float* d_w = d_std + order * dim;
        PPNET_CUDA(cudaMemcpyAsync(d_mean, mean, 4 * (size_t)order * dim, cudaMemcpyHostToDevice, s.st));
        PPNET_CUDA(cudaMemcpyAsync(d_std, stdv, 4 * (size_t)order * dim, cudaMemcpyHostToDevice, s.st));
        PPNET_CUDA(cudaMemcpyAsync(d_w, weights, 4 * (size_t)order, cudaMemcpyHostToDevice, s.st));
        c->h2d_bytes += 4 * order * (2 * dim + 1);
        rc = ppnet_gmm_sample(seed, sample0 + (uint64_t)i0, ni, order, dim, d_mean, d_std, d_w, (float*)s.buf[0], nullptr,
                              (void*)s.st);
        if (rc != PPNET_OK) return rc;
        PPNET_CUDA(cudaMemcpyAsync(out + i0 * dim, s.buf[0], 4 * (size_t)ni * dim, cudaMemcpyDeviceToHost, s.st));
        c->d2h_bytes += 4 * ni * dim;
    }
    PPNET_CUDA(cudaStreamSynchronize(c->slot[0].st));
    PPNET_CUDA(cudaStreamSynchronize(c->slot[1].st));
    return PPNET_OK;
}

extern "C" int ppnet_gmm_sample_host(void* ctx, uint64_t seed, uint64_t sample0, int64_t n, int32_t order, int32_t dim,
                                     const float* mean, const float* stdv, const float* weights, float* out) {
    return drain_on_error((Ctx*)ctx, gmm_host_run((Ctx*)ctx, seed, sample0, n, order, dim, mean, stdv, weights, out));
}
