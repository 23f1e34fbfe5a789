/* ppnet_b200 -- C ABI of the B200-native EDaGe-PP hot path (drop-in boundary).
 *
 * The PPNet reference has no FFI layer: its boundary is a set of Python functions/methods
 * (SURVEY.md 8(b)).  Each entry point below replaces the numerical body of the reference
 * function cited beside it; `ppnet_b200/*.py` keeps the reference's Python names/signatures and
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

const char* ppnet_last_error(void);
int ppnet_version(void);
/* number of kernel launches this library has enqueued since load (bench.py's gpu_launches) */
int64_t ppnet_launch_count(void);

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
 *      accept[M][O]; out[M][O][3] = accepted [col_px, row_px, r_px] compacted in order;
 *      out_cnt[M].                                                                               */
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

#ifdef __cplusplus
}
#endif
#endif /* PPNET_B200_H */
