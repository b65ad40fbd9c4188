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

#define MODEST_ABI_VERSION 2

int modest_abi_version(void);
const char* modest_last_error(void);
/* Number of kernel launches this library has enqueued since process start (bench.py's
 * `gpu_launches`). */
int64_t modest_launch_count(void);

/* Stage A plumbing of the streaming engine (load_velo_scan, utils/pointcloud_utils.py:61, for every
 * frame of pre_compute_pp_score.py:133-150): n host -> device copies enqueued on `stream` in one call.
 * h_src[i] must be pinned host memory (the copies are asynchronous), d_dst[i] device memory of at
 * least n_bytes[i] bytes.  The three arrays are host arrays. */
int modest_upload_frames(const void* const* h_src, void* const* d_dst, const int64_t* n_bytes, int n,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage B: bring scan frames into the fixed frame.
 * Replaces transform_points() (utils/pointcloud_utils.py:11-19: [p,1] @ Tr^T in float32) and,
 * optionally, remove_center() (pre_compute_pp_score.py:48-52, nuScenes history frames).
 *   d_in (M,point_stride) f32; frame f owns rows [d_frame_off[f], d_frame_off[f+1])
 *   d_T  (n_frames,16) f32 row-major 4x4 (the float32 matrix get_relative_pose returns)
 *   h_center_box host float[4] {x_lo, x_hi, y_lo, y_hi}: rows with x_lo <= x < x_hi and
 *        y_lo <= y < y_hi are removed when remove_center != 0 (written as NaN so that every
 *        later distance test rejects them; row positions are preserved)
 *   d_out (M,3) f32
 * ------------------------------------------------------------------------------------------ */
int modest_transform_frames_batch(const float* d_in, int point_stride,
                                  const int64_t* d_frame_off, const float* d_T, int n_frames,
                                  int64_t max_frame_points, int remove_center,
                                  const float* h_center_box, float* d_out, void* stream);

/* Stage B for the streaming engine (pre_compute_pp_score.py:125-167): raw frames
 * stay in a device cache (one allocation per velodyne/*.bin) and are addressed by pointer.
 * d_jobs is an array of n_jobs 96-byte records, built on the host and copied to the device:
 *     struct { const float* src; int64_t dst_row; int32_t n; int32_t flags; float T[16]; int64_t pad; }
 * Job j reads n rows of src_stride floats at src and writes rows dst_row .. dst_row+n-1 of d_out
 * (out_stride floats per row): [x,y,z,1] @ T^T rounded like transform_frames_batch; flags & 1
 * applies remove_center (NaN rows) with h_center_box; flags & 2 copies the rows unchanged
 * (columns beyond src_stride are zero) -- used to assemble a batch's raw (N,4) scans. */
int modest_transform_gather_batch(const void* d_jobs, int n_jobs, int src_stride, int out_stride,
                                  int64_t max_frame_points, const float* h_center_box, float* d_out,
                                  void* stream);

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
 *
 * Two history passes produce the same bits:
 *   tiled (default)  -- needs the HOST copies of the offset tables (h_q_off, h_h_off, h_trav_off:
 *                the arrays the caller built before uploading them), at most 32 traversals per
 *                scan and grid_dim % 8 == 0.  The batch is cut into groups of consecutive scans
 *                of about `group_points` history points (0 = default 2 M) so that a group's
 *                intermediate stays L2-resident; `bin_records` float4 slots of the workspace hold
 *                it: take the value modest_pp_bin_records() returns for the same tables.
 *   global hash      -- any input; taken when a host table is NULL, bin_records == 0,
 *                group_points < 0, or a scan has more than 32 traversals.
 * ------------------------------------------------------------------------------------------ */
int64_t modest_pp_bin_records(const int64_t* h_h_off, const int32_t* h_trav_off, int n_scans,
                              int64_t group_points);
size_t modest_pp_workspace_bytes(int n_scans, int64_t n_query_total, int64_t n_count_total,
                                 int grid_dim, int64_t bin_records);
int modest_pp_score_batch(const float* d_query_xyz, const int64_t* d_q_off,
                          const float* d_hist_xyz, const int64_t* d_h_off,
                          const int32_t* d_trav_off, int n_scans, int n_trav_total,
                          int64_t n_query_total, int64_t n_count_total,
                          int64_t max_query_points, int64_t max_trav_points, double radius,
                          int grid_dim, int32_t* d_counts, const int64_t* d_count_off,
                          float* d_pp, const int64_t* h_q_off, const int64_t* h_h_off,
                          const int32_t* h_trav_off, int64_t group_points, int64_t bin_records,
                          void* d_ws, size_t ws_bytes, void* stream);

/* Timing hook for the history pass (tiled: tile query + per group count / plan / scatter /
 * join kernels; global hash: pp_count_kernel): with n_slots > 0 every modest_pp_score_batch call
 * records a CUDA event pair around it on the launching stream, in a ring of n_slots pairs (n_slots = 0 switches it off).  After synchronising the
 * stream, modest_pp_profile_read writes the durations [ms] of the most recent launches
 * (newest first) into h_ms and returns how many it wrote. */
int modest_pp_profile_enable(int n_slots);
int modest_pp_profile_read(float* h_ms, int max_out);

/* ------------------------------------------------------------------------------------------
 * Stage E: RANSAC ground plane.
 * Replaces estimate_plane() (utils/pointcloud_utils.py:44-65), i.e. sklearn's
 * RANSACRegressor().fit(xy, z) with all defaults (trial loop: sklearn/linear_model/_ransac.py
 * :447-560, an un-vendored dependency of the reference).
 *
 * Step 1, modest_plane_candidates_batch: keeps, in order, the points with z < max_hs and
 *   x_lo < x < x_hi, y_lo < y < y_hi (pointcloud_utils.py:45-49) and computes the residual
 *   threshold MAD(z) in float32 (_ransac.py:396-398).
 *     d_ptc (NP,point_stride) f32 rows [x,y,z,...]; scan s owns rows [d_off[s], d_off[s+1])
 *     d_cand (NP,3) f32 out: scan s's candidates are packed at row d_off[s]
 *     d_n_cand (n_scans) i32 out;  d_thr (n_scans) f32 out
 * Step 2, modest_ransac_fit_batch: scores `max_trials` minimal-set hypotheses per scan,
 *   replays sklearn's sequential accept / dynamic-early-stop rules over them, refits on the
 *   consensus set and emits the plane [a,b,c,d] (c > 0) of pointcloud_utils.py:53-62.
 *     d_triples  (n_scans,max_trials,3) i32 indices into each scan's candidate list -- drawn by
 *                the caller (parity: numpy's global RandomState, like sklearn) -- or NULL to
 *                draw them on the device (throughput mode): trial h of scan s uses a
 *                counter-based hash of (seed, d_scan_keys[s], h), so a scan's minimal sets --
 *                and with them its labels -- do not depend on its slot in the batch or on how
 *                the scans are sharded over GPUs
 *     d_scan_keys (n_scans) i64 per-scan key (the scan id) or NULL (key = s)
 *     d_plane    (n_scans,4) f64 out (NaN when no valid consensus set exists)
 *     d_model    (n_scans,3) f64 out or NULL: the float32 coef_[0], coef_[1], intercept_
 *     d_info     (n_scans,4) i32 out: n_candidates, n_trials_ consumed, best trial, n_inliers
 *     d_triples_out  optional (n_scans,max_trials,3) i32: the triples actually used
 *     d_inlier_mask  optional (NP) u8: consensus mask per candidate (packed like d_cand)
 * ------------------------------------------------------------------------------------------ */
int modest_plane_candidates_batch(const float* d_ptc, int point_stride, const int64_t* d_off,
                                  int n_scans, float max_hs, float x_lo, float x_hi, float y_lo,
                                  float y_hi, float* d_cand, int32_t* d_n_cand, float* d_thr,
                                  void* stream);
size_t modest_ransac_workspace_bytes(int n_scans, int max_trials);
int modest_ransac_fit_batch(const float* d_cand, const int64_t* d_off, const int32_t* d_n_cand,
                            const float* d_thr, int n_scans, int64_t max_points,
                            const int32_t* d_triples, uint64_t seed, const int64_t* d_scan_keys,
                            int max_trials, double* d_plane, double* d_model, int32_t* d_info,
                            int32_t* d_triples_out, uint8_t* d_inlier_mask, void* d_ws,
                            size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * road-plane files for OpenPCDet's ground-truth sampling.
 * Replaces the numeric part of extract_ransac() (data_preprocessing/RANSAC.py:9-68): rect-camera
 * coordinates in f64, candidates min_h < y < max_h, -10 < z < 70, -20 < x < 20 (in order),
 * sklearn RANSACRegressor().fit((x,z), y) in float64, plane [a, -1, b, c] / |(a,-1,b)|; fewer than
 * 5 candidates give the script's default [0, -1, 0, 1.65].  Same two-step protocol as stage E:
 *   modest_road_candidates_batch: d_calib (S,21) f64 as for the box stage; d_cand (NP,3) f64 out,
 *       rows (x, z, y) packed at d_off[s]; d_n_cand (S) i32 out; d_thr (S) f64 out (MAD of y)
 *   modest_road_plane_fit_batch:  d_triples (S,max_trials,3) i32 or NULL (device draws);
 *       d_plane (S,4) f64 out; d_info (S,4) i32 out; workspace as modest_ransac_workspace_bytes
 * ------------------------------------------------------------------------------------------ */
int modest_road_candidates_batch(const float* d_ptc, int point_stride, const int64_t* d_off,
                                 const double* d_calib, int n_scans, double min_h, double max_h,
                                 double* d_cand, int32_t* d_n_cand, double* d_thr, void* stream);
int modest_road_plane_fit_batch(const double* d_cand, const int64_t* d_off,
                                const int32_t* d_n_cand, const double* d_thr, int n_scans,
                                int64_t max_points, const int32_t* d_triples, uint64_t seed,
                                int max_trials, double* d_plane, int32_t* d_info, void* d_ws,
                                size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stages F+G: ground removal and range gate, compacted.
 * Replaces above_plane()/distance_to_plane() (utils/pointcloud_utils.py:68-81) and the
 * limit_range product of generate_mask.py:57-65.  A point survives iff
 *     NOT( (p.n + d)/|n| < offset  AND  only_x_lo < x < only_x_hi AND only_y_lo < y < only_y_hi )
 *     AND lim_x_lo < x <= lim_x_hi AND lim_y_lo < y <= lim_y_hi
 *   d_planes      (n_scans,4) f64;  d_pp (NP) f32 PP scores
 *   h_only_range  host float[4] {x_lo,x_hi,y_lo,y_hi} or NULL (the reference's only_range=None)
 *   h_limit_range host float[4]
 *   d_kept        (NP,4) f32 out: surviving points as (x,y,z,pp), original order, packed at
 *                 row d_off[s];  d_kept_idx (NP) i32 out: their indices within the scan;
 *   d_n_kept      (n_scans) i32 out;  d_mask (NP) u8 out or NULL: the boolean final_mask
 * ------------------------------------------------------------------------------------------ */
int modest_ground_mask_batch(const float* d_ptc, int point_stride, const int64_t* d_off,
                             const float* d_pp, const double* d_planes, int n_scans,
                             double offset, const float* h_only_range,
                             const float* h_limit_range, float* d_kept, int32_t* d_kept_idx,
                             int32_t* d_n_kept, uint8_t* d_mask, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage H: mutual-kNN AND radius graph with L1 PP-score weights.
 * Replaces precompute_affinity_matrix(ptc, pp, 'radius_mutual_knn', 'l1', n_neighbors, radius)
 * (utils/clustering_utils.py:32-48; sklearn kneighbors_graph / radius_neighbors_graph).
 * Edge (i,j), i != j:  j among the n_neighbors nearest other points of i, i among those of j,
 * and dx*dx+dy*dy+dz*dz <= radius*radius, all in sequential f64 on the f32 coordinates;
 * weight = |pp_i - pp_j| in f32.  Output is a fixed-width adjacency: row i of scan s lives at
 * row d_off[s]+i and holds d_nbr_cnt entries (unordered) of d_nbr (neighbour index within the
 * kept list) and d_nbr_w.  *d_flags (one i32) gets bit0/bit1 set if distance ties made a
 * k-th neighbour ambiguous / overflowed a row (never on duplicate-free data).
 *   partition_eps / d_nbr_eps_cnt  optional (pass a negative value / NULL to switch off): the
 *                 caller announces the DBSCAN radius it will use; every row then lists the edges
 *                 with (double)w <= partition_eps first and d_nbr_eps_cnt (NP) i32 receives their
 *                 number (-1 for rows left unpartitioned).  The order of the entries of a row
 *                 carries no other meaning (the reference's CSR is rebuilt sorted by column).
 *                 With partitioned rows d_nbr_w may be NULL: only the eps-edges are then written
 *                 (rows of d_nbr_eps_cnt entries) -- all DBSCAN needs (n_neighbors <= 96).
 * ------------------------------------------------------------------------------------------ */
size_t modest_graph_workspace_bytes(int n_scans, int64_t n_points_total, int n_neighbors,
                                    int grid_dim);
int modest_affinity_graph_batch(const float* d_kept, const int64_t* d_off,
                                const int32_t* d_n_kept, int n_scans, int64_t n_points_total,
                                int64_t max_points, int n_neighbors, double radius,
                                int grid_dim, int32_t* d_nbr, float* d_nbr_w,
                                int32_t* d_nbr_cnt, double partition_eps,
                                int32_t* d_nbr_eps_cnt, int32_t* d_flags, void* d_ws,
                                size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * the other rectangle fitters of get_obj() (utils/pointcloud_utils.py:88-165,
 * 219-275) and get_lowest_point_rect() (:278-290), one cluster per call.
 *   modest_fit_rectangle: d_xz (n,2) f64 rect (x,z) of the cluster; method 0 min_zx_area_fit
 *     (minimum_bounding_rectangle), 1 PCA (PCA_rectangle), 2 variance_to_edge (variance_rectangle;
 *     needs the search-angle tables of modest_filter_and_fit_batch and a workspace);
 *     d_out 11 doubles: corners (4,2) row-major, angle, area, status (0 ok).
 *   modest_lowest_point_rect: max rect-y of the rows of d_rect (n,3) strictly inside the footprint
 *     centred (cx,cz) with half sizes l/2, w/2 rotated by ry (pass numpy's cos(ry), sin(ry));
 *     *d_bottom = that y, NaN when no point is inside.
 * ------------------------------------------------------------------------------------------ */
size_t modest_fit_rectangle_workspace_bytes(int n, int n_angles);
int modest_fit_rectangle(const double* d_xz, int n, int method, const double* d_trig,
                         const double* d_angles, int n_angles, double* d_out, void* d_ws,
                         size_t ws_bytes, void* stream);
int modest_lowest_point_rect(const double* d_rect, int n, double cx, double cz, double cos_ry,
                             double sin_ry, double l, double w, double* d_bottom, void* stream);

/* ------------------------------------------------------------------------------------------
 * the non-default graph / affinity types of precompute_affinity_matrix()
 * (utils/clustering_utils.py:16-31,49-56), one scan per call, exact brute force (these types are
 * not on the seed-label path).
 *   modest_knn_bruteforce: the n_neighbors nearest neighbours of every point at any distance
 *     (sklearn kneighbors_graph: sequential f64 squared distances, self excluded);
 *     d_knn (n, n_neighbors) i32, d_knn_cnt (n) i32 (= min(n_neighbors, n-1)), d_rk2 (n) f64 the
 *     k-th squared distance; d2_max = an upper bound of every squared distance (bounding box);
 *     *d_flags |= 2 when exact ties at the k-th distance were cut to the first k by index.
 *   modest_radius_graph: radius_neighbors_graph (d2 <= radius^2, self excluded) in two calls:
 *     counts (d_counts (n) i64, d_indices NULL), then -- after the caller's prefix sum -- the
 *     fill (d_indptr (n+1) i64, d_indices i32, ascending within a row).
 *   modest_edge_affinity: the float32 edge weights of :42-56 for a CSR pattern, widened to f64:
 *     kind 0 'l1' |pp_r - pp_j|, 1 'exp' exp((pp_r - pp_j)^2), 2 '3d_l2_distance' the norm of the
 *     difference of the first `width` columns of the point rows.
 * ------------------------------------------------------------------------------------------ */
int modest_knn_bruteforce(const float* d_pts, int point_stride, int n, int n_neighbors, double d2_max,
                          int32_t* d_knn, int32_t* d_knn_cnt, double* d_rk2, int32_t* d_flags,
                          void* stream);
int modest_radius_graph(const float* d_pts, int point_stride, int n, double radius,
                        int64_t* d_counts, const int64_t* d_indptr, int32_t* d_indices, void* stream);
int modest_edge_affinity(const float* d_pts, int point_stride, int width, const float* d_pp,
                         const int64_t* d_indptr, const int32_t* d_indices, int n, int kind,
                         double* d_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage I: DBSCAN on that graph.
 * Replaces sklearn.cluster.DBSCAN(metric='precomputed', eps, min_samples).fit(graph).labels_
 * (generate_mask.py:75-81): neighbourhood(i) = {j : edge, (double)w <= eps} + {i}; core iff
 * its size >= min_samples; clusters = connected components of core-core edges, numbered by
 * their smallest member index; a border point takes the smallest cluster id among its core
 * neighbours; everything else -1.
 *   d_nbr_eps_cnt optional (NULL = scan the weights): the prefix counts written by
 *                 modest_affinity_graph_batch called with partition_eps == eps
 *   d_labels_kept (NP) i32 out: label per kept point;  d_labels_full (NP) i32 out: label per
 *   original point (-1 for removed points), i.e. the `labels` array of generate_mask.py:76-81
 *   d_n_clusters  (n_scans) i32 out
 * ------------------------------------------------------------------------------------------ */
size_t modest_dbscan_workspace_bytes(int64_t n_points_total);
int modest_dbscan_batch(const int64_t* d_off, const int32_t* d_n_kept,
                        const int32_t* d_kept_idx, int n_scans, int64_t n_points_total,
                        int64_t max_points, int n_neighbors, const int32_t* d_nbr,
                        const float* d_nbr_w, const int32_t* d_nbr_cnt,
                        const int32_t* d_nbr_eps_cnt, double eps,
                        int min_samples, int32_t* d_labels_kept, int32_t* d_labels_full,
                        int32_t* d_n_clusters, void* d_ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stages J-M: cluster filtering, box fitting, volume gate, final labels.
 * Replaces filter_labels()/is_valid_cluster() (utils/clustering_utils.py:94-135, given the
 * second plane), Calibration.project_velo_to_rect (utils/kitti_util.py:293-329),
 * get_obj(..., 'closeness_to_edge') (utils/pointcloud_utils.py:167-216,278-317) and the volume
 * gate + label compaction of generate_mask.py:92-103.
 *   d_labels          (NP) i32 DBSCAN labels per original point (-1 = none), d_n_clusters (S)
 *   d_planes          (S,4) f64: the SECOND plane (filter_labels' own RANSAC fit)
 *   d_calib           (S,21) f64: Tr_velo_to_cam (3x4 row-major) then R0_rect (3x3)
 *   d_rect_in         optional (NP,3) f64: rect-camera coordinates of every point; when given
 *                     they are used instead of projecting d_ptc with d_calib (the operator-level
 *                     get_obj() receives rect coordinates directly)
 *   h_gates           host double[8]: min_points, max_min_height, min_max_height,
 *                     percentile/100 evaluated in float32, min_percentile_pp_score,
 *                     min_volume, max_volume, d0 (closeness floor, 1e-2)
 *   d_trig            (4,n_angles) f64: cos(a), sin(a), cos(a+pi/2), sin(a+pi/2) for the search
 *                     headings a -- a constant table the caller evaluates with the host libm so
 *                     that it holds the same values numpy gives the reference
 *   d_angles          (2,n_angles) f64: a and a+pi/2
 *   d_labels_filtered (NP) i32 out: filter_labels() result (0 = background, 1..K)
 *   d_labels_final    (NP) i32 out: labels after the volume gate, compacted (the seg .npy)
 *   d_boxes           (S,max_boxes,8) f64 out: t.x t.y t.z l w h ry volume per surviving box
 *   d_n_boxes, d_n_valid (S) i32 out;  d_flags (1) i32 out: bit2 cluster capacity, bit3 empty
 *                     footprint, bit4 box capacity exceeded
 * ------------------------------------------------------------------------------------------ */
size_t modest_filter_workspace_bytes(int n_scans, int64_t n_points_total, int max_clusters);
int modest_filter_and_fit_batch(const float* d_ptc, int point_stride, const int64_t* d_off,
                                const float* d_pp, const int32_t* d_labels,
                                const int32_t* d_n_clusters, const double* d_planes,
                                const double* d_calib, const double* d_rect_in, int n_scans,
                                int64_t n_points_total,
                                int64_t max_points, int max_clusters, int max_boxes,
                                const double* h_gates, const double* d_trig,
                                const double* d_angles, int n_angles,
                                int32_t* d_labels_filtered, int32_t* d_labels_final,
                                double* d_boxes, int32_t* d_n_boxes, int32_t* d_n_valid,
                                int32_t* d_flags, void* d_ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * PP-score percentile of the scan points inside each box.
 * Replaces the numeric part of filter_by_ppscore() (combine_labels.py:42-60): a point is
 * inside when its rect (x,z), moved into the box frame, lies strictly inside (-l/2,l/2) x
 * (-w/2,w/2) and t.y - h < y <= t.y; the percentile is numpy's float32 'linear' rule.
 *   d_boxes (S,max_boxes,8) f64 rows t.x t.y t.z l w h ry (8th column ignored), d_n_boxes (S)
 *   d_box_trig optional (S,max_boxes,2) f64: cos(ry), sin(ry) evaluated by the caller with the
 *           host libm (what numpy gives the reference); NULL -> evaluated on the device
 *   q_f32   percentile/100 evaluated in float32 by the caller
 *   d_percentile (S,max_boxes) f32 out, d_count (S,max_boxes) i32 out (0 -> box holds no point)
 * Rect coordinates come from d_rect_in when given, else from d_ptc + d_calib.
 * ------------------------------------------------------------------------------------------ */
size_t modest_box_pp_workspace_bytes(int n_scans, int64_t n_points_total, int64_t max_points,
                                     int max_boxes);
int modest_box_pp_percentile_batch(const float* d_ptc, int point_stride, const int64_t* d_off,
                                   const float* d_pp, const double* d_calib,
                                   const double* d_rect_in, const double* d_boxes,
                                   const double* d_box_trig, const int32_t* d_n_boxes, int n_scans,
                                   int64_t n_points_total, int64_t max_points, int max_boxes,
                                   double q_f32, float* d_percentile, int32_t* d_count,
                                   void* d_ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage N: rotated BEV IoU / overlap / NMS on boxes [x, y, z, dx, dy, dz, heading] (f32).
 * Drop-in for the reference's pybind11 module iou3d_nms_cuda
 * (utils/iou3d_nms/src/iou3d_nms_api.cpp:11-17; host wrappers iou3d_nms.cpp:48-188;
 * kernels iou3d_nms_kernel.cu:236-372):
 *   boxes_iou_bev_gpu     -> modest_boxes_iou_bev      out (num_a,num_b) f32
 *   boxes_overlap_bev_gpu -> modest_boxes_overlap_bev
 *   nms_gpu               -> modest_nms_bev    boxes sorted by score; writes the kept indices
 *   nms_normal_gpu        -> modest_nms_normal (axis-aligned IoU)
 * The NMS calls return the number kept through *h_num_out and therefore synchronise `stream`
 * (as the reference does with its blocking cudaMemcpy); d_keep (device, i64) and h_keep
 * (host, i64) are both optional.
 * modest_seed_nms_batch is the batched form of objs_nms() (utils/pointcloud_utils.py:320-344,
 * use_score_rank=False): per scan, K x K IoU of [t.x,t.z,0,l,w,h,-ry] (rounded to f32), visit
 * boxes by descending self-IoU (ties: larger index first), suppress IoU > thresh.
 *   d_boxes (S,max_boxes,8) f64 rows as written by modest_filter_and_fit_batch
 *   d_iou_or_null (S,max_boxes,max_boxes) f32 out (optional), d_iou_ws same shape scratch
 *   d_keep (S,max_boxes) u8 out
 * ------------------------------------------------------------------------------------------ */
int modest_boxes_iou_bev(const float* d_boxes_a, int num_a, const float* d_boxes_b, int num_b,
                         float* d_iou, void* stream);
int modest_boxes_overlap_bev(const float* d_boxes_a, int num_a, const float* d_boxes_b,
                             int num_b, float* d_overlap, void* stream);
size_t modest_nms_workspace_bytes(int n);
int modest_nms_bev(const float* d_boxes, int n, float thresh, int64_t* d_keep, int64_t* h_keep,
                   int* h_num_out, void* d_ws, size_t ws_bytes, void* stream);
int modest_nms_normal(const float* d_boxes, int n, float thresh, int64_t* d_keep,
                      int64_t* h_keep, int* h_num_out, void* d_ws, size_t ws_bytes,
                      void* stream);
int modest_seed_nms_batch(const double* d_boxes, const int32_t* d_n_boxes, int n_scans,
                          int max_boxes, float thresh, float* d_iou_or_null, float* d_iou_ws,
                          uint8_t* d_keep, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage O (host): field-of-view gate + KITTI label text.
 * Replaces is_within_fov() and objs2label() (utils/pointcloud_utils.py:347-379, with
 * compute_box_3d / project_to_image, utils/kitti_util.py:383-389,405-478).  All pointers are
 * HOST pointers.  h_boxes (n,8) f64 rows t.x t.y t.z l w h ry volume; h_keep_in optional
 * (n) u8; h_P the 3x4 P2 matrix; h_scores optional (n) -> the `with_score` line format.
 * Writes '\n'-joined "%.4f" lines without trailing newline (NUL-terminated) into h_text.
 * ------------------------------------------------------------------------------------------ */
int modest_kitti_labels_host(const double* h_boxes, int n, const uint8_t* h_keep_in,
                             const double* h_P, int fov_only, int image_h, int image_w,
                             const char* obj_type, const double* h_scores, char* h_text,
                             size_t text_cap, size_t* h_len, int* h_n_out,
                             uint8_t* h_kept_out);
/* Batched form for the streaming engine: scan s owns h_n_boxes[s] rows at
 * h_boxes + s*max_boxes*8, keep flags at h_keep + s*max_boxes (optional), P2 at h_P + s*12.
 * The texts are written back to back into h_text; h_text_off (n_scans+1, i64) gets their offsets. */
int modest_kitti_labels_batch_host(const double* h_boxes, const int32_t* h_n_boxes,
                                   const uint8_t* h_keep, int n_scans, int max_boxes,
                                   const double* h_P, int fov_only, int image_h, int image_w,
                                   const char* obj_type, char* h_text, size_t text_cap,
                                   int64_t* h_text_off);

#ifdef __cplusplus
}
#endif
#endif /* MODEST_B200_H_ */
