/* ppnet_b200 -- C ABI of the B200-native EDaGe-PP hot path (drop-in boundary).
 *
 * The PPNet reference has no FFI layer: its boundary is a set of Python functions/methods
 * (SURVEY.md 8(b)).  Each entry point below replaces the numerical body of the reference
 * function cited beside it; the modules under `ppnet_b200/` keep the reference's Python names/signatures and
 * binds these symbols with ctypes (see INTEGRATION.md for the stub a PPNet maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - `*_dev` style (default): every pointer is a DEVICE pointer owned by the caller, no hidden
 *     allocation, work is enqueued on `stream` (a cudaStream_t passed as void*), no sync.
 *   - `*_host` entry points take HOST pointers, stage through an internal per-handle arena
 *     (pinned + device), copy H2D, launch, copy D2H and synchronise before returning.
 *   - return 0 on success, a negative PPNET_E_* code otherwise; never throws, never exits.
 *     ppnet_last_error() gives a thread-local message.
 *   - thread-safe for distinct streams / distinct handles.
 *   - Segments are grouped by map in CSR form: map m owns segments [seg_off[m], seg_off[m+1]).
 *     If seg_off is NULL the grouping is uniform: map m owns [m*segs_per_map, (m+1)*segs_per_map).
 *   - Obstacles: obs[M][omax][3] = (x, y, r) float64, obs_cnt[M] valid rows per map.
 */
#ifndef PPNET_B200_H
#define PPNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPNET_OK 0
#define PPNET_E_INVALID (-1)   /* bad argument */
#define PPNET_E_CUDA (-2)      /* CUDA runtime error (see ppnet_last_error) */
#define PPNET_E_NOMEM (-3)

#define PPNET_DOT_FUSED_SKX 0  /* np.dot == fma(a1,b1, a0*b0): OpenBLAS SkylakeX ddot (AVX-512 hosts) */
#define PPNET_DOT_UNFUSED 1    /* np.dot == a0*b0 + a1*b1:     OpenBLAS Haswell/Zen ddot */

#define PPNET_CMP_F32_NEP50 0   /* A12 compares its float32 offsets with float32(thr): NumPy >= 2 (NEP 50)        */
#define PPNET_CMP_F64_NUMPY1 1  /* ... promoted to float64 against the Python-float thr: NumPy 1.x (torch 1.11 era) */

const char* ppnet_last_error(void);
int ppnet_version(void);
/* number of kernel launches this library has enqueued since load (bench.py's gpu_launches) */
int64_t ppnet_launch_count(void);
/* sizeof the parameter structs as compiled (0: ppnet_gen_params, 1: ppnet_path_params, 2: ppnet_pipeline_io): layout
 * guard for bindings */
int64_t ppnet_sizeof_params(int32_t which);

/* ---- A11  process_map.collision_check_circle_edge(s, e, obs, clearance)
 *      EDaGe-PP/process_map.py:383-425.  pts_rc[N][4] = (s_row, s_col, e_row, e_col) exactly as
 *      the reference receives them (it swaps to (x, y) itself).  verdict[N] = 1 collision.     */
int ppnet_segcheck_edage_f64(const double* pts_rc, int64_t n_segs, const int64_t* seg_off,
                             int64_t segs_per_map, int64_t n_maps, const double* obs,
                             const int32_t* obs_cnt, int32_t omax, double clearance, double bound,
                             int32_t dot_mode, uint8_t* verdict, void* stream);

/* ---- A12  neuralplanner.collision_check_circle_edge(s, e, idx) / steerTo(start, end, idx)
 *      experiments/MPNet/neuralplanner.py:43-69, 86-92.  pts_xy[N][4] = (s_x, s_y, e_x, e_y)
 *      float32.  verdict and steer may each be NULL.  steer[i] = 0 blocked / 1 free.           */
int ppnet_segcheck_mpnet_f32(const float* pts_xy, int64_t n_segs, const int64_t* seg_off,
                             int64_t segs_per_map, int64_t n_maps, const double* obs,
                             const int32_t* obs_cnt, int32_t omax, double clearance, double bound,
                             uint8_t* verdict, uint8_t* steer, void* stream);

/* ---- A11 + A12 fused on ONE read of the segments (the hot path's verdict step).  pts_rc as for A11; the A12 flavour
 *      runs on (x, y) = (float32(col), float32(row)) of the same segment -- the cast + swap a float32 caller holds.
 *      Any of the four outputs may be NULL: verdict_* are bytes, vbits_* are bit-packed (bit i & 31 of word i >> 5 =
 *      segment i; ceil(n_segs / 32) words).  Each flavour is bit-identical to its own entry point above.
 *      cmp_mode selects the NumPy generation of the A12 threshold comparison (neuralplanner.py:54,66).            */
int ppnet_verdict_fused(const double* pts_rc, int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map,
                        int64_t n_maps, const double* obs, const int32_t* obs_cnt, int32_t omax, double clearance,
                        double bound, int32_t dot_mode, int32_t cmp_mode, uint8_t* verdict_f64, uint8_t* verdict_f32,
                        uint32_t* vbits_f64, uint32_t* vbits_f32, void* stream);

/* ---- A12  feasibility_check(path, idx)  neuralplanner.py:96-102
 *      waypoints wp[total][2] f32, path p owns [path_off[p], path_off[p+1]), uses obstacle set
 *      path_map[p].  feasible[p] in {0,1}; n_checked[p] (may be NULL) = steerTo calls the
 *      reference would have made (it stops at the first blocked edge).                          */
int ppnet_path_feasible_f32(const float* wp, const int64_t* path_off, const int32_t* path_map,
                            int64_t n_paths, const double* obs, const int32_t* obs_cnt,
                            int32_t omax, double clearance, double bound, uint8_t* feasible,
                            int32_t* n_checked, void* stream);

/* ---- A12  lvc(path, idx)  neuralplanner.py:123-138 (lazy vertex contraction)
 *      out_wp has the same layout/offsets as wp; out_len[p] = contracted length.                */
int ppnet_lvc_f32(const float* wp, const int64_t* path_off, const int32_t* path_map,
                  int64_t n_paths, const double* obs, const int32_t* obs_cnt, int32_t omax,
                  double clearance, double bound, float* out_wp, int32_t* out_len, void* stream);

/* ---- A14  MapGenerate.generate_map_randomly clearance verdict  EDaGe-PP/MapGenerate.py:132-143
 *      pathpt[M][np][2] (row, col) float64; cand[M][O][3] = (x, y, r) in map units.
 *      accept[M][O]; out[M][O][3] = accepted [col_px, row_px, r_px] compacted in order (rows >= out_cnt[m]
 *      are left untouched); out_cnt[M].                                                                               */
int ppnet_clearance_filter_f64(const double* pathpt, int32_t np, const double* cand, int32_t O,
                               int64_t n_maps, double map_size, double resolution, double clearance,
                               uint8_t* accept, double* out, int32_t* out_cnt, void* stream);

/* ---- A4   Path.coord_euclidean2image  EDaGe-PP/Path.py:378-386 (the grid-index rule)
 *      idx = int(rint(v / (map_size/resolution) + mapoffset)), half-to-even.                     */
int ppnet_grid_index_f64(const double* pts, int64_t n_values, double map_size, double resolution,
                         double mapoffset, int32_t* idx, void* stream);

/* ---- A5   Path.free_space_bydirection + the four driver loops of Path.path_space
 *      EDaGe-PP/Path.py:113-134, 397-404.  n_paths corridors, rays_per_path rays each:
 *      x0/dir[n_paths][rays][2], step_num[n_paths]; space[n_paths][W][H] uint8 must be zeroed by
 *      the caller; painted cells are set to `value`.                                             */
int ppnet_corridor_paint(const double* x0, const double* dir, const double* step_num,
                         int64_t n_paths, int32_t rays_per_path, double map_size, double resolution,
                         double mapoffset, int32_t W, int32_t H, uint8_t value, uint8_t* space,
                         void* stream);

/* ---- A7 (mask) / N1: torchvision's rigid resampling of a corridor mask: RandomRotation(degrees=(d, d)) (nearest,
 *      about the image centre) then functional.affine(translate=(tx, ty)), cropped to the top-left Ro x Ro
 *      (EDaGe-PP/Path.py:160-161, 175-178; MapGenerate.py:102-106).  src[n][Ws][Ws] -> out[n][Ro][Ro] uint8.      */
int ppnet_mask_rigid(const uint8_t* src, int32_t Ws, const double* angle_deg, const double* translate, int64_t n,
                     int32_t Ro, uint8_t* out, void* stream);

/* ---- A6   Path.convexhull  EDaGe-PP/Path.py:388-395 (scipy.spatial.ConvexHull on integer cells)
 *      pts[P][np][2] int32 -> hull[P][hmax][2] CCW from the lexicographically smallest vertex,
 *      strict corners only; hull_cnt[P] (a value > hmax means the output was truncated).          */
int ppnet_hull2d_i32(const int32_t* pts, int32_t np, int64_t n_paths, int32_t hmax, int32_t* hull,
                     int32_t* hull_cnt, void* stream);

/* ---- A10  Path.boundary_check(angle, translation)  EDaGe-PP/Path.py:100-111
 *      hull[B][hmax][2] float64 (row, col), hull_cnt[B]; query i uses hull path_idx[i] (NULL -> 0),
 *      angle_arg[i] (degrees, as passed to the method) and trans_arg[i][2].  ok[i] in {0,1};
 *      hull_out[n][hmax][2] (may be NULL) receives the transformed vertices.                      */
int ppnet_boundary_check(const double* hull, const int32_t* hull_cnt, int32_t hmax,
                         const int32_t* path_idx, const double* angle_arg, const double* trans_arg,
                         int64_t n, double resolution, uint8_t* ok, double* hull_out, void* stream);

/* ---- counter-based sampling (Philox4x32-10); the reference draws from global MT19937 streams
 *      (np.random / torch.rand), which cannot be sharded; parity is distributional.
 *      out[u][k] = k-th uniform double in [0,1) of unit (unit0 + u) in sub-stream stream_id.     */
int ppnet_uniform_f64(uint64_t seed, uint32_t stream_id, uint64_t unit0, int64_t n_units,
                      int32_t per_unit, double* out, void* stream);

/* ---- A17  GMM(order, dim, mean_range, std_range)  EDaGe-PP/GMM.py:7-16
 *      gmm_params draws mean/std [order][dim] and weights [order] (float32, torch.rand-style);
 *      gmm_sample = Distribution.sample([n]) -> out[n][dim] float32, comp[n] (may be NULL).       */
int ppnet_gmm_params(uint64_t seed, int32_t order, int32_t dim, float mean_range, float std_range,
                     float* mean, float* stdv, float* weights, void* stream);
int ppnet_gmm_sample(uint64_t seed, uint64_t sample0, int64_t n, int32_t order, int32_t dim,
                     const float* mean, const float* stdv, const float* weights, float* out,
                     int32_t* comp, void* stream);

/* ---- A15  plot_obstacles  EDaGe-PP/Path.py:36-49, restated geometrically (parity unpinned):
 *      bits[M][R][ceil(R/32)] uint32, bit (j&31) of word (j>>5) of row i set iff the centre of
 *      pixel (row i, col j) is inside a disk (x, y, r + inflate).                                 */
int ppnet_raster_circles_bits(const double* obs, const int32_t* obs_cnt, int32_t omax,
                              int64_t n_maps, int32_t resolution, double inflate, uint32_t* bits,
                              void* stream);

/* ---- A15, second mode: the canvas model of the same function (matplotlib's default 576 x 432 canvas at dpi 90, data ->
 *      pixel px = 72 + x / size_w * 446.4, py = 51.84 + y / size_h * 332.64, crop [53:383, 73:517], bilinear resize of the
 *      330 x 444 crop to R x R, occupied iff the resized value < 0.5).  obs (x, y, r) in the units of `size` (the
 *      reference passes size = resolution).  Same bit layout as above.  Quantifies the +-1 px gap between the pinned
 *      geometry and the centre-in-disk rule (anti-aliasing / JPEG / dither still cannot be pinned).                  */
int ppnet_raster_canvas_bits(const double* obs, const int32_t* obs_cnt, int32_t omax, int64_t n_maps, double size_w,
                             double size_h, int32_t resolution, double inflate, uint32_t* bits, void* stream);

/* ---- A15 return value + MapGenerate.py:111: image[M][3][R][R] float32, 1 = free / 0 = obstacle from the bit-packed
 *      map, plus `add` (same shape, the placed corridor mask `path_space`; may be NULL).                         */
int ppnet_bits_to_image(const uint32_t* bits, int32_t resolution, int64_t n_maps, const float* add, float* image,
                        void* stream);

/* ---- A16  process_map.add_init_end_single(image, init, end)  EDaGe-PP/process_map.py:119-145
 *      image[M][3][R][R] float32 in place; init/end[M][2] (row, col): 7x7 (255, 0, 0) stamps, clipped.           */
int ppnet_add_init_end(float* image, int32_t resolution, const double* init, const double* end, int64_t n_maps,
                       void* stream);

/* ---- integer DDA grid check (new functionality; endpoints snapped with the A4 rule).
 *      segs_xy[N][4] float32 pixel coordinates grouped by map (CSR: pass the longest row in
 *      segs_per_map).  verdict[i] = 1 if a visited cell is occupied or outside [0,R)^2;
 *      first_hit[i] (may be NULL) = step index of that cell, -1 when free.                        */
int ppnet_dda_gridcheck(const uint32_t* bits, int32_t resolution, int64_t n_maps,
                        const float* segs_xy, int64_t n_segs, const int64_t* seg_off,
                        int64_t segs_per_map, uint8_t* verdict, int32_t* first_hit, void* stream);

/* the same walk on the A11 array read directly: segs_rc[N][4] = (s_row, s_col, e_row, e_col) float64, walked as
 * (x, y) = (float32(col), float32(row)) -- identical verdicts to ppnet_dda_gridcheck on that cast + swap.
 * verdict (bytes) and vbits (bit-packed, ceil(n_segs / 32) words) may each be NULL, not both.                     */
int ppnet_dda_gridcheck_rc64(const uint32_t* bits, int32_t resolution, int64_t n_maps, const double* segs_rc,
                             int64_t n_segs, const int64_t* seg_off, int64_t segs_per_map, uint8_t* verdict,
                             int32_t* first_hit, uint32_t* vbits, void* stream);

/* ---- A10 + A13 + A14 (+ A15): the fused map generator, one launch for n_maps maps.
 *      MapGenerate.generate inner block EDaGe-PP/MapGenerate.py:58-124 and generate_map_randomly
 *      :126-151.  Map g = map0 + i uses target path (g / reps) % n_bank  (index rule :68).        */
typedef struct ppnet_gen_params {
    /* target-path bank, device pointers */
    const double* bank_pathpt;     /* [n_bank][np][2]     Path.PathPoint  (row, col)              */
    const double* bank_segpt;      /* [n_bank][nseg1][2]  Path.SegPointImage                      */
    const double* bank_hull;       /* [n_bank][hmax][2]   Path.ConvexHull                         */
    const int32_t* bank_hull_cnt;  /* [n_bank]                                                    */
    const double* bank_obs;        /* [n_bank][pomax][3]  Path.obstacles [x, y, r] (may be NULL)  */
    const int32_t* bank_obs_cnt;   /* [n_bank]                                                    */
    int32_t n_bank, np, nseg1, hmax, pomax;
    /* what to generate */
    int64_t map0, n_maps;
    int32_t reps;                  /* maps per target path before moving on (= P in the reference) */
    int32_t obstacles_num;         /* O  */
    int32_t max_tries;             /* placement retry budget (reference guard: 10^6)               */
    int32_t reserved0;
    double resolution, map_size, obstacle_size, clearance, raster_inflate;
    uint64_t seed;
    /* optional caller-supplied draws (parity mode); NULL -> Philox keyed by the global map index  */
    const double* in_angle;        /* [n_maps]      angle as drawn (degrees)                      */
    const int32_t* in_trans;       /* [n_maps][2]   translation as drawn [t0, t1]                 */
    const double* in_cand;         /* [n_maps][O][3] (x, y, r) in map units                       */
    /* outputs (any may be NULL) */
    double* out_angle;             /* [n_maps]                                                    */
    int32_t* out_trans;            /* [n_maps][2]                                                 */
    double* out_segpt;             /* [n_maps][nseg1][2]                                          */
    double* out_pathpt;            /* [n_maps][np][2]                                             */
    double* out_obs;               /* [n_maps][O + pomax][3] accepted random first, then path obs */
    int32_t* out_obs_cnt;          /* [n_maps]                                                    */
    int32_t* out_rand_cnt;         /* [n_maps] accepted random obstacles                          */
    uint32_t* out_bits;            /* [n_maps][R][ceil(R/32)]                                     */
    int32_t* out_tries;            /* [n_maps] placement tries used, 0 = budget exhausted         */
    uint8_t* out_valid;            /* [n_maps]                                                    */
    unsigned long long* counters;  /* [4] += maps_done, valid_paths, obstacles_accepted, tries    */
} ppnet_gen_params;
int ppnet_generate_maps(const ppnet_gen_params* params, void* stream);

/* ---- A1 + A2 + A3 + A5 + A6 + A7 + A8 + A9: target-path synthesis, what PathGroup.generate does per accepted
 *      Path (EDaGe-PP/PathGenerate.py:33-50 -> PathSeg.py:10-58, Path.py:78-98, 113-193, 224-356, 388-395,
 *      463-537).  All pointers are device pointers; every output except space_raw is required.
 *      S = seg_num, C = poly_order + 1, Np = 100 S, Nb = 100 S + 100.                                          */
typedef struct ppnet_path_params {
    int64_t path0, n_paths;        /* global path ids [path0, path0 + n_paths): the Philox unit                 */
    int32_t seg_num, poly_order;   /* S <= 64, order 2..7 (reference: 10 and 4)                                 */
    int32_t hmax, pomax;           /* hull / isle capacity (<= 128), path-obstacle capacity                     */
    int32_t max_obst_iter;         /* A9 guard per isle (the reference loops until it succeeds)                 */
    int32_t max_obst_rand;         /* row length of in_obst_rand                                                */
    double clearance, map_size, resolution;
    double width_coef;             /* Path.search_isle(width_coef): isle depth threshold, reference default 0.2 */
    uint64_t seed;
    /* optional caller-supplied draws (parity mode); NULL -> Philox keyed by the global path id                 */
    const uint8_t* force_straight; /* [n]        PathGroup's 1 % forced-straight flag                           */
    const uint8_t* in_straight;    /* [n][S]     resolved PathSeg.is_straight                                   */
    const double* in_y;            /* [n][S][1000] np.random.random(1000) of PathSeg.random                     */
    const double* in_uend;         /* [n][S]     np.random.random(1) of EndPoint (the end point itself with in_poly) */
    const double* in_poly;         /* [n][S][C]  PathSeg.random(poly, endpoint): skip the fit                   */
    const float* in_obst_rand;     /* [n][max_obst_rand] torch.rand(1) values consumed by set_obstacles         */
    const int32_t* in_obst_rand_cnt; /* [n]                                                                     */
    const int32_t* in_hull;        /* [n][hmax][2] raw hull cells in the reference's vertex order               */
    const int32_t* in_hull_cnt;    /* [n]                                                                       */
    /* A1 outputs */
    double* poly;                  /* [n][S][C]  highest power first                                            */
    double* endpoint;              /* [n][S]                                                                    */
    uint8_t* is_straight;          /* [n][S]                                                                    */
    uint8_t* path_straight;        /* [n]        Path.is_straight                                               */
    double* seg_trans_local;       /* [n][S][2]  PathSeg.translation(): (EndPoint, polyval(EndPoint))           */
    double* grad_st;               /* [n][S]                                                                    */
    double* grad_end;              /* [n][S]                                                                    */
    double* seg_length;            /* [n][S]                                                                    */
    /* A2 outputs */
    double* seg_rot;               /* [n][S]     PathSeg.Rotation after transform()                             */
    double* seg_trans;             /* [n][S][2]  PathSeg.Translation after transform()                          */
    double* segpoint_raw;          /* [n][S+1][2] Path.SegPoint (map units)                                     */
    double* pathpoint_raw;         /* [n][Np][2] Path.PathPoint before normalisation                            */
    double* length;                /* [n]        Path.Length                                                    */
    int32_t* cells;                /* [n][Np][2] coord_euclidean2image(PathPoint, mapoffset = R)                */
    /* A3 outputs */
    double* up;                    /* [n][S][50][2] upboundary.point                                            */
    double* up_dir;                /* [n][S][50][2] upboundary.direction (downboundary.direction = -up_dir)     */
    double* down;                  /* [n][S][50][2] downboundary.point                                          */
    double* cap_init;              /* [n][50][2] initboundary                                                   */
    double* cap_end;               /* [n][50][2] endboundary                                                    */
    double* boundary_raw;          /* [n][Nb][2] BoundaryPoint before normalisation                             */
    /* A5: the rays path_space paints, in its order (init cap, end cap, up, down)                              */
    double* ray_x0;                /* [n][Nb][2]                                                                */
    double* ray_dir;               /* [n][Nb][2]                                                                */
    double* step_num;              /* [n]        0.8 * clearance / step_len                                     */
    uint8_t* space_raw;            /* [n][2R][2R] painted corridor (optional)                                   */
    uint8_t* space;                /* [n][R][R]  Path.Space: the corridor after space_normalization (with space_raw) */
    /* A6 */
    int32_t* hull_raw;             /* [n][hmax][2]                                                              */
    int32_t* hull_cnt;             /* [n]                                                                       */
    /* A7 */
    double* rotation;              /* [n]        Path.Rotation (degrees)                                        */
    double* translation;           /* [n][2]     Path.Translation as stored (swapped)                           */
    double* neg_rotation_ws;       /* [n]        workspace (-rotation, the angle handed to the mask rotation)   */
    double* hull;                  /* [n][hmax][2] Path.ConvexHull                                              */
    double* segpoint_img;          /* [n][S+1][2] Path.SegPointImage                                            */
    double* pathpoint;             /* [n][Np][2] Path.PathPoint (row, col)                                      */
    double* boundary;              /* [n][Nb][2] Path.BoundaryPoint                                             */
    /* A8 */
    int32_t* isle;                 /* [n][hmax][2] (lo, hi): isle = PathPoint[lo:hi]                            */
    int32_t* isle_cnt;             /* [n]                                                                       */
    /* A9 */
    double* obs;                   /* [n][pomax][3] Path.obstacles [x, y, r]                                    */
    int32_t* obs_cnt;              /* [n]                                                                       */
    int32_t* obst_rand_used;       /* [n]        torch.rand(1) draws consumed                                   */
    int32_t* status;               /* [n] bit0: max_obst_iter hit, bit1: supplied draws exhausted, bit2: pomax overflow,
                                      bit3: hull / isle capacity (hmax) overflow -- the bank entry is unusable           */
} ppnet_path_params;
int ppnet_path_synthesize(const ppnet_path_params* params, void* stream);

/* ---- "next" rows ------------------------------------------------------------------------------
 * N2  process_map.generate_gen_path  EDaGe-PP/process_map.py:148-163: out[M][R][R] uint8, 255 at round(p) of every
 *     `stride`-th (5th) label point with 0 < row, col < R, 0 elsewhere (the PNG encoding stays with the caller).   */
int ppnet_path_mask(const double* pathpt, int32_t np, int64_t n_maps, int32_t stride, int32_t resolution, uint8_t* out,
                    void* stream);
/* N3  process_map.extract_path  EDaGe-PP/process_map.py:293-365: greedy 8-neighbour walk on the down-sampled heat-map
 *     mask[n][h][w] float32 from init_state/ds towards end_state/ds.  out[n][max_len + 2][2] = init_state,
 *     ds * walk..., end_state; out_len[n] = number of rows (0 on failure); ok[n].  The reference's 1 s wall-clock
 *     timeout becomes the step budget max_len.                                                                    */
int ppnet_extract_path(const float* mask, int32_t h, int32_t w, const double* init_state, const double* end_state,
                       double down_sample_rate, int64_t n, int32_t max_len, double* out, int32_t* out_len, uint8_t* ok,
                       void* stream);

/* N4  gerated_by_planners.generated_by_planners  EDaGe-PP/gerated_by_planners.py:88-157: the corridor and path label
 *     masks of planner solutions.  wp[total][2] (x, y), solution i owns [path_off[i], path_off[i+1]) (max_len = the
 *     longest); mask_space / mask_path [n][R][R] uint8 (row = y, col = x), 1 where the reference paints.            */
int ppnet_planner_masks(const double* wp, const int64_t* path_off, int64_t n_paths, int64_t max_len, double clearance,
                        int32_t resolution, int32_t points_per_seg, uint8_t* mask_space, uint8_t* mask_path, void* stream);

/* ---- ordered stream compaction of the survivors of a verdict array (free segments, feasible paths ...):
 *      out_idx[0 .. *out_count) = ascending i with flags[i] == keep.  workspace: ppnet_compact_workspace_elems(n)
 *      int64 elements of device memory (per-CTA counts; no hidden allocation).                                     */
int64_t ppnet_compact_workspace_elems(int64_t n);
int ppnet_compact_u8(const uint8_t* flags, int64_t n, uint8_t keep, int64_t* out_idx, int64_t* out_count,
                     int64_t* workspace, void* stream);

/* the same for bit-packed verdicts: survivors = segments whose bit is CLEAR in every given array (b, c may be NULL),
 * e.g. free under A11 and A12 and the DDA.  out_idx int32[n] (first *out_count valid, ascending).                  */
int64_t ppnet_compact_bits_workspace_elems(int64_t n);
int ppnet_compact_bits(const uint32_t* a, const uint32_t* b, const uint32_t* c, int64_t n, int32_t idx_base,
                       int32_t* out_idx, int64_t* out_count, int64_t* workspace, void* stream);
/* short flag arrays (per-map flags, e.g. the valid paths of a slice): one CTA, out_idx int32 = idx_base + i.      */
int ppnet_compact_u8_i32(const uint8_t* flags, int64_t n, uint8_t keep, int32_t idx_base, int32_t* out_idx,
                         int64_t* out_count, void* stream);

/* ---- device-side segment source (north star: "segment proposal"): the config-2 candidate segments of maps
 *      [map0, map0 + n_maps), s ~ U(0, R)^2, e = s + N(0, sigma^2) per axis, as segs_rc f64[n_maps * segs_per_map][4].
 *      Segment k of global map g is a pure function of (seed, g, k) (Philox4x32-10), so any sharding gives the same
 *      bytes and generator-mode callers upload nothing.                                                          */
int ppnet_propose_segments(uint64_t seed, uint64_t map0, int64_t n_maps, int64_t segs_per_map, double resolution,
                           double sigma, double* segs_rc, void* stream);

/* ---- 64-bit content digest of per-map arrays (cross-rank identity proof): *acc += sum over units u and 32-bit words k
 *      of mix(word ^ mix(mix(global_unit_index, salt) + k)) modulo 2^64 -- additive over any split of the unit range.
 *      data[n_units][words_per_unit] uint32 (any dtype viewed as words); with rows != NULL only the first
 *      rows[u] * row_words words of unit u count.  acc is a DEVICE uint64.                                         */
int ppnet_digest_u32(const uint32_t* data, int64_t words_per_unit, int64_t n_units, uint64_t unit0,
                     const int32_t* rows, int32_t row_words, uint64_t salt, uint64_t* acc, void* stream);

/* ---- N1 dataset writer (host code): MapGenerate.generate_map_randomly's record file, EDaGe-PP/MapGenerate.py:144-149.
 *      One JSON line per problem {"Index", "Init", "End", "Length", "Obstacles"}, byte for byte what json.dumps writes
 *      (floats as Python's repr).  HOST pointers: index[n], init / end [n][2], length[n], obs[n][omax][3] (first obs_cnt[n]
 *      rows).  Appends to `path` when append != 0.                                                                   */
int ppnet_write_problems_jsonl(const char* path, int32_t append, int64_t n, const int64_t* index, const double* init,
                               const double* end, const double* length, const double* obs, const int32_t* obs_cnt,
                               int32_t omax, int64_t* out_bytes);

/* ---- host-buffer boundary (e2e): HOST pointers, copies inside, synchronous on return.          */
int ppnet_ctx_create(int32_t device, void** ctx);
int ppnet_ctx_destroy(void* ctx);
int ppnet_ctx_bytes(void* ctx, int64_t* h2d_bytes, int64_t* d2h_bytes);
int ppnet_segcheck_edage_f64_host(void* ctx, const double* pts_rc, int64_t n_segs,
                                  const int64_t* seg_off, int64_t segs_per_map, int64_t n_maps,
                                  const double* obs, const int32_t* obs_cnt, int32_t omax,
                                  double clearance, double bound, int32_t dot_mode, uint8_t* verdict);
int ppnet_segcheck_mpnet_f32_host(void* ctx, const float* pts_xy, int64_t n_segs,
                                  const int64_t* seg_off, int64_t segs_per_map, int64_t n_maps,
                                  const double* obs, const int32_t* obs_cnt, int32_t omax,
                                  double clearance, double bound, uint8_t* verdict, uint8_t* steer);
int ppnet_clearance_filter_f64_host(void* ctx, const double* pathpt, int32_t np, const double* cand,
                                    int32_t O, int64_t n_maps, double map_size, double resolution,
                                    double clearance, uint8_t* accept, double* out, int32_t* out_cnt);

/* Target-path bank resident on the device (host arrays in, opaque handle out).                    */
int ppnet_bank_upload(int32_t device, const double* pathpt, const double* segpt, const double* hull,
                      const int32_t* hull_cnt, const double* obs, const int32_t* obs_cnt,
                      int32_t n_bank, int32_t np, int32_t nseg1, int32_t hmax, int32_t pomax,
                      void** bank);
int ppnet_bank_free(void* bank);
/* MapGenerate.generate with HOST outputs: `params` carries the settings and host out_* pointers
 * (its bank_* and in_* fields are ignored; counters, if set, is a host uint64[4] accumulator).    */
int ppnet_generate_maps_host(void* ctx, void* bank, const ppnet_gen_params* params);
/* MapGenerate.generate + the verdicts on the freshly generated maps in ONE host call: slice by slice the candidate
 * segments (uniform grouping, segs_per_map per map, HOST pointers) are uploaded, the maps generated, A11 / A12 / DDA
 * run against them on the device, and labels + verdicts downloaded; uploads, kernels and downloads of neighbouring
 * slices overlap.  NULL verdict pointers skip that check.                                                        */
typedef struct ppnet_pipeline_io {
    const double* segs_rc_f64;     /* [n_maps * segs_per_map][4] (s_row, s_col, e_row, e_col), for verdict_f64  */
    const float* segs_xy_f32;      /* [n_maps * segs_per_map][4] (s_x, s_y, e_x, e_y), for verdict_f32 / _dda;
                                      NULL (with segs_rc_f64 or a device proposal): the float32 flavours run on the
                                      device-side cast + swap (x, y) = (float32(col), float32(row)) -- ONE upload   */
    int64_t segs_per_map;
    double clearance_px, bound;
    int32_t dot_mode, cmp_mode;    /* PPNET_DOT_*, PPNET_CMP_* */
    uint8_t* verdict_f64;          /* A11, one byte per segment */
    uint8_t* verdict_f32;          /* A12 */
    uint8_t* verdict_dda;          /* integer DDA vs the bit-packed map */
    /* ---- round 2 (appended; zero-initialise the struct to get the round-1 behaviour) ----
     * bit-packed verdicts: bit (i & 31) of word (i >> 5) = segment i of this call; ceil(n/32) words each; they need
     * the one-array mode (segs_xy_f32 == NULL). */
    uint32_t* vbits_f64;
    uint32_t* vbits_f32;
    uint32_t* vbits_dda;
    /* survivors, compacted on the device with warp scans: ascending segment indices (of this call) that are free
     * under EVERY verdict requested above, and ascending map indices (relative to map0) with a valid placement */
    int32_t* free_idx;             /* [n_maps * segs_per_map] capacity */
    int64_t* free_count;           /* [1] */
    int32_t* valid_idx;            /* [n_maps] capacity */
    int64_t* valid_count;          /* [1] */
    /* device-side segment source: when segs_rc_f64 == NULL and propose_sigma > 0 the candidate segments are drawn on
     * the device (ppnet_propose_segments with params->seed, the global map index and the resolution): nothing is
     * uploaded.  out_segs_rc (may be NULL) receives them. */
    double propose_sigma;
    double* out_segs_rc;           /* [n_maps * segs_per_map][4] */
} ppnet_pipeline_io;
int ppnet_generate_and_check_host(void* ctx, void* bank, const ppnet_gen_params* params, const ppnet_pipeline_io* io);
int ppnet_dda_gridcheck_host(void* ctx, const uint32_t* bits, int32_t resolution, int64_t n_maps,
                             const float* segs_xy, int64_t n_segs, const int64_t* seg_off,
                             int64_t segs_per_map, uint8_t* verdict, int32_t* first_hit);
int ppnet_gmm_sample_host(void* ctx, uint64_t seed, uint64_t sample0, int64_t n, int32_t order,
                          int32_t dim, const float* mean, const float* stdv, const float* weights,
                          float* out);

#ifdef __cplusplus
}
#endif
#endif /* PPNET_B200_H */
