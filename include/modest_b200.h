/*
 * libmodest_b200 -- C ABI of the B200-native MODEST seed-label hot path.
 *
 * Every entry point replaces one numeric stage of the reference's
 * generate_cluster_mask/ programs (paths below are relative to that directory of
 * YurongYou/MODEST).  The only FFI the reference itself has on this path is the pybind11
 * module `iou3d_nms_cuda` (utils/iou3d_nms/src/iou3d_nms_api.cpp:11-17); its five names are
 * re-exported by modest_b200/generate_cluster_mask/utils/iou3d_nms on top of the
 * `modest_boxes_*` / `modest_nms_*` functions below.  The other stages are Python + SciPy /
 * scikit-learn in the reference; their replacements are bound with ctypes
 * (modest_b200/_lib.py; INTEGRATION.md shows the stub a maintainer would add upstream).
 *
 * Conventions
 *   - plain C: pointers, sizes, scalars.  No torch / C++ types cross this boundary.
 *   - pointers named d_* are DEVICE pointers, h_* HOST pointers.  The caller owns all memory,
 *     including the scratch `d_ws` (size it with the matching *_workspace_bytes call); the
 *     library never allocates device memory and never frees anything.
 *   - `stream` is a cudaStream_t passed as void*.  Work is enqueued asynchronously on it; no
 *     entry point synchronises unless its comment says so.
 *   - return value: MODEST_OK (0) or a negative MODEST_ERR_* code; modest_last_error() gives
 *     the text for the calling thread.  Nothing exits the process (the reference's op calls
 *     exit(), utils/iou3d_nms/src/iou3d_nms.cpp:14-38) and nothing throws.
 *   - ragged batches: scan s owns rows [off[s], off[s+1]) of a concatenated array.
 */
#ifndef MODEST_B200_H_
#define MODEST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MODEST_OK 0
#define MODEST_ERR_ARG (-1)       /* bad argument / workspace too small */
#define MODEST_ERR_CUDA (-2)      /* a CUDA runtime call or launch failed */
#define MODEST_ERR_CAPACITY (-3)  /* a fixed-capacity device buffer overflowed */

#define MODEST_ABI_VERSION 1

int modest_abi_version(void);
const char* modest_last_error(void);
/* Number of kernel launches this library has enqueued since process start (bench.py's
 * `gpu_launches`). */
int64_t modest_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Stage C+D: persistence-point (PP) score.
 * Replaces count_neighbors() + compute_ephe_score() (pre_compute_pp_score.py:54-75) and the
 * cKDTree builds at :188-190.
 *
 * For every query point q of scan s and every traversal t of that scan:
 *     count[q][t] = #{ h in traversal t : dx*dx + dy*dy + dz*dz <= radius*radius }  (f64, seq.)
 *     P = count / (sum_t count + 1e-8);   H = -sum_t P ln(P + 1e-8) / ln(T_s);   pp = (float)H
 *
 *   d_query_xyz  (NQ,3) f32, all scans concatenated; scan s owns rows [d_q_off[s], d_q_off[s+1])
 *   d_hist_xyz   (NH,3) f32, all traversals of all scans concatenated;
 *                traversal g owns rows [d_h_off[g], d_h_off[g+1])
 *   d_trav_off   (n_scans+1) i32: scan s owns traversals [d_trav_off[s], d_trav_off[s+1])
 *   n_trav_total = d_trav_off[n_scans];  max_trav_points = max_g (d_h_off[g+1]-d_h_off[g]);
 *   max_query_points = max_s (d_q_off[s+1]-d_q_off[s])            (host copies, for grid sizing)
 *   grid_dim     cells per side of the per-scan 2-D hash grid over x,y (0 -> default 512)
 *   d_counts     optional (may be NULL): i32, scan s holds an (N_s, T_s) row-major block at
 *                d_count_off[s] (i64 element offsets, n_scans+1 entries).  When NULL the counts
 *                live in the workspace only.
 *   d_pp         (NQ) f32 out.
 * ------------------------------------------------------------------------------------------ */
size_t modest_pp_workspace_bytes(int n_scans, int64_t n_query_total, int64_t n_count_total,
                                 int grid_dim);
int modest_pp_score_batch(const float* d_query_xyz, const int64_t* d_q_off,
                          const float* d_hist_xyz, const int64_t* d_h_off,
                          const int32_t* d_trav_off, int n_scans, int n_trav_total,
                          int64_t n_query_total, int64_t n_count_total,
                          int64_t max_query_points, int64_t max_trav_points, double radius,
                          int grid_dim, int32_t* d_counts, const int64_t* d_count_off,
                          float* d_pp, void* d_ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MODEST_B200_H_ */
