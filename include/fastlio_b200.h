/* fastlio_b200.h -- C ABI of the B200-native FAST-LIO2 measurement-update path.
 *
 * The reference (hku-mars/FAST_LIO) has no FFI layer: the hot path is reached through two
 * C++ class APIs used by src/laserMapping.cpp.  This header is the extern "C" boundary that
 * sits underneath drop-in facades of those two classes (include/ikd-Tree/ikd_Tree.h and
 * include/IKFoM_toolkit/esekfom/esekfom.hpp in this repository); each entry point cites
 * the reference interface it replaces (paths relative to the reference tree).
 *
 * Conventions: opaque handles; caller-owned HOST buffers unless a name says _device;
 * every function returns an int status (0 = ok, <0 = error, see FL_ERR_*) except the
 * counting queries; no exceptions cross the boundary; fl_last_error() returns a
 * thread-local message.  One CUDA stream per map handle; handles are thread-compatible
 * (serialise calls on one handle), distinct handles are independent.
 *
 * Point layout everywhere: 4 floats (x, y, z, intensity) -- the fields of
 * pcl::PointXYZINormal (include/common_lib.h:37) that the path reads.
 * State layout (26 doubles): pos(3) rot(x,y,z,w) offset_R_L_I(x,y,z,w) offset_T_L_I(3)
 * vel(3) bg(3) ba(3) grav(3)  == state_ikfom (include/use-ikfom.hpp:12-21), quaternions in
 * Eigen coeffs() order.  Covariance: 23 x 23 doubles, row-major, DOF order of state_ikfom.
 */
#ifndef FASTLIO_B200_H
#define FASTLIO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FL_OK 0
#define FL_ERR_CUDA (-1)
#define FL_ERR_ARG (-2)
#define FL_ERR_NCCL (-3)
#define FL_ERR_STATE (-4)
#define FL_ERR_CAPACITY (-5)

typedef struct fl_map fl_map_t;        /* replaces KD_TREE<PointType>          include/ikd-Tree/ikd_Tree.h:48-341 */
typedef struct fl_filter fl_filter_t;  /* replaces esekfom::esekf<state_ikfom,12,input_ikfom> + h_share_model */

/* Same layout as the oracle's per-pass log; used by the parity tests. */
typedef struct fl_pass_log {
    int searched, valid, effct, converged;
    double res_sum;
    double HtH[144];
    double Hth[12];
    double x_after[26];
} fl_pass_log_t;

const char* fl_last_error(void);
/* Page-lock a caller-owned, long-lived host buffer (e.g. the scan buffer reused every scan) so that
 * fl_filter_update / fl_map_add_points copy from it by DMA; unregister before freeing it. */
int fl_host_register(const void* ptr, unsigned long long bytes);
int fl_host_unregister(const void* ptr);
int fl_device_count(void);
int fl_version(void);

/* ------------------------------------------------------------------ map: KD_TREE<PointType> */
/* KD_TREE::KD_TREE(delete_param, balance_param, box_length)          ikd_Tree.h:309, ikd_Tree.cpp:9-18
 * (the two rebuild criteria have no counterpart: leaves are re-packed by fl_map_rebuild / automatically) */
int fl_map_create(fl_map_t** out, int device, float downsample_size);
int fl_map_destroy(fl_map_t* m);   /* drops the caller's reference; filters / scans created on the map keep it alive until they go */
/* KD_TREE::set_downsample_param                                      ikd_Tree.h:319-322 */
int fl_map_set_downsample(fl_map_t* m, float downsample_size);
/* KD_TREE::Build(PointVector)                                        ikd_Tree.cpp:409-423 */
int fl_map_build(fl_map_t* m, const float* pts_xyzi, int n);
/* KD_TREE::size() / validnum()                                       ikd_Tree.cpp:70-97, 140-163 */
int fl_map_size(fl_map_t* m);
int fl_map_validnum(fl_map_t* m);
/* KD_TREE::Nearest_Search, batched over nq queries, k <= 5           ikd_Tree.cpp:426-461
 * out_pts: nq*k*4 floats (ascending distance), out_d2: nq*k floats, out_cnt: nq ints.  Safe for
 * concurrent host callers on one handle (internally serialised). */
int fl_map_knn(fl_map_t* m, const float* q_xyzi, int nq, int k, float* out_pts, float* out_d2, int* out_cnt);
/* KD_TREE::Add_Points(PointVector&, bool downsample_on) -> int       ikd_Tree.cpp:478-573
 * returns the reference's return value (>= 0) or an error (< 0) */
int fl_map_add_points(fl_map_t* m, const float* pts_xyzi, int n, int downsample_on);
/* KD_TREE::Delete_Point_Boxes(vector<BoxPointType>&) -> int          ikd_Tree.cpp:632-658
 * boxes6: nb * (min xyz, max xyz); returns the number of points invalidated or an error (< 0) */
int fl_map_delete_boxes(fl_map_t* m, const float* boxes6, int nb);
/* KD_TREE::Add_Point_Boxes(vector<BoxPointType>&)                     ikd_Tree.cpp:576-603 (Add_by_range :854-934)
 * points that Delete_Point_Boxes removed, lie in the boxes and have not been overwritten by later inserts come back
 * (points removed by the down-sampling of Add_Points do not, as in the reference); returns how many, or an error (< 0) */
int fl_map_add_boxes(fl_map_t* m, const float* boxes6, int nb);
/* KD_TREE::acquire_removed_points(PointVector&)                       ikd_Tree.cpp:661-676 (called at laserMapping.cpp:225)
 * the points removed by Delete_Point_Boxes since the previous call (the reference hands them over when it rebuilds the
 * subtree, this map at once); the first call starts the record and returns 0.  Returns the number of points (writes at
 * most cap of them), or an error (< 0).  Call with cap >= the return value of the deletes since the last call. */
int fl_map_acquire_removed(fl_map_t* m, float* out_xyzi, int cap);
/* KD_TREE::flatten(Root_Node, Storage, NOT_RECORD): all valid points  ikd_Tree.cpp:1627-1658
 * returns the number of valid points (writes at most cap of them) or an error (< 0) */
int fl_map_flatten(fl_map_t* m, float* out_xyzi, int cap);
/* KD_TREE::tree_range()                                              ikd_Tree.cpp:100-137 */
int fl_map_tree_range(fl_map_t* m, float* box6);
/* KD_TREE::Rebuild of the whole tree (ikd_Tree.cpp:736-764): re-sorts all valid points into fresh leaves */
int fl_map_rebuild(fl_map_t* m);
/* introspection: [0] main leaves [1] overflow leaves [2] internal levels [3] rebuilds so far */
int fl_map_stats(fl_map_t* m, int* out4);
/* The k-NN fast path: a hashed directory of cubic cells over the same points (no reference counterpart; the results are
 * those of KD_TREE::Nearest_Search either way).  on = 0 answers every query through the BVH walk; cell_size <= 0 picks
 * 2 x downsample_size.  Takes effect immediately (the directory is re-listed). */
int fl_map_set_cell_directory(fl_map_t* m, int on, float cell_size);
/* [0] cells [1] external buckets [2] crowded cells (queries touching them use the BVH walk) [3] table capacity
 * [4] directory re-lists triggered by inserts [5] queries answered by the BVH walk since the last call (-1: directory off) */
int fl_map_dir_stats(fl_map_t* m, int* out6);

/* ------------------------------------------------------------------ filter: esekf + h_share_model */
/* esekf::esekf + esekf::init_dyn_share(f, f_x, f_w, h_share_model, maximum_iteration, limit)
 *                                                                   esekfom.hpp:238-254, laserMapping.cpp:826-828 */
int fl_filter_create(fl_filter_t** out, fl_map_t* map, int max_points);
int fl_filter_destroy(fl_filter_t* f);
/* maximum_iter, limit[23], extrinsic_est_en (laserMapping.cpp:739,789) */
int fl_filter_set_params(fl_filter_t* f, int max_iter, const double* limit23, int extrinsic_est_en);
/* 1 (default): the gain of esekfom.hpp:1782-1809 through one 6x6 (12x12 with extrinsic estimation) solve,
 *    algebraically identical;
 * 0: the information form with two 23x23 inversions exactly as written in the reference (validation) */
int fl_filter_set_solver(fl_filter_t* f, int mode);
/* kNN of the search passes: 1 (default) one lane per scan point through the map's cell directory, the BVH walk for
 * whatever that cannot prove; 0 one warp per scan point through the BVH walk only -- same neighbours, same distances */
int fl_filter_set_search(fl_filter_t* f, int mode);
/* 1 (default): the whole update in ONE persistent kernel launch (h_share_model fused with the search, the Kalman step
 * in the kernel's solver block); 0: the two-kernels-per-pass chain it grew out of (A/B; solver mode 0 always uses it) */
int fl_filter_set_fused(fl_filter_t* f, int on);
/* esekf::update_iterated_dyn_share_modified(R, solve_time) with feats_down_body bound
 *                                                                   esekfom.hpp:1619-1931, laserMapping.cpp:638-754, :960
 * x26 / P: in = kf.get_x()/get_P() before the update, out = after.  solve_time_s (may be NULL)
 * is incremented by the device time of the update, like the reference's out-parameter. */
int fl_filter_update(fl_filter_t* f, const float* body_xyzi, int nq, double* x26, double* P, double R, double* solve_time_s);
/* map_incremental() (laserMapping.cpp:427-474) without leaving the device: classifies the bound scan with the
 * updated state and the cached neighbours, then performs the two Add_Points calls (:470-471).
 * out3 (may be NULL): [0] |PointToAdd| [1] |PointNoNeedDownsample| [2] return value of Add_Points(PointToAdd, true) */
int fl_filter_map_incremental(fl_filter_t* f, double filter_size_map_min, int flg_EKF_inited, int* out3);
/* Nearest_Points after the update (laserMapping.cpp:102, read by map_incremental :438-460) */
int fl_filter_get_nearest(fl_filter_t* f, float* out_pts, int* out_cnt, int nq);
/* point_selected_surf after the update (laserMapping.cpp:76) */
int fl_filter_get_selected(fl_filter_t* f, unsigned char* out, int nq);
/* per-pass H^T H, H^T h, effct_feat_num, total_residual, state -- for parity tests / the reference's debug log */
int fl_filter_get_pass_logs(fl_filter_t* f, fl_pass_log_t* out, int cap);
/* pieces of fl_filter_update for pipelines that keep the scan resident in HBM */
int fl_filter_upload_scan(fl_filter_t* f, const float* body_xyzi, int nq);
int fl_filter_upload_state(fl_filter_t* f, const double* x26, const double* P, double R);
int fl_filter_run(fl_filter_t* f);                 /* enqueue all passes, asynchronous */
int fl_filter_download_state(fl_filter_t* f, double* x26, double* P, int* n_pass);   /* synchronises */
int fl_filter_sync(fl_filter_t* f);
/* device time in milliseconds of `reps` back-to-back resident updates from the uploaded state
 * (CUDA events on the handle's stream); optionally flushes L2 between repetitions */
int fl_filter_time_resident(fl_filter_t* f, int reps, int flush_l2, float* ms_total);
/* wall-clock seconds of `reps` consecutive fl_filter_update(body, nq, copy of x26, copy of P, R) calls issued from
 * native code (the per-scan cost a C++ caller such as laserMapping.cpp sees); x26_out / P_out (may be NULL): last result */
int fl_filter_time_e2e(fl_filter_t* f, const float* body_xyzi, int nq, const double* x26, const double* P, double R, int reps,
                       double* seconds, double* x26_out, double* P_out);
/* device time of `reps` launches of the dominant kernel alone (k_search: the kNN of the first pass of an update) */
int fl_filter_time_search_pass(fl_filter_t* f, int reps, int flush_l2, float* ms_total);
int fl_filter_gpu_launches(fl_filter_t* f);
/* clock64() stamps of the last on-device Kalman step (tuning aid; layout in scripts/profile_once.py) */
int fl_filter_debug_prof(fl_filter_t* f, long long* out16);

/* ------------------------------------------------------------------ scan front end
 * The two steps that produce feats_down_body, kept in HBM on the map's device and stream so that a scan goes
 * raw -> de-skewed -> down-sampled -> update -> map_incremental with one upload. */
typedef struct fl_scan fl_scan_t;      /* replaces the feats_undistort / feats_down_body clouds   laserMapping.cpp:122-124 */
int fl_scan_create(fl_scan_t** out, fl_map_t* map);
int fl_scan_destroy(fl_scan_t* s);
/* Measures.lidar (common_lib.h:47-58): n x (x,y,z,intensity) + PointType::curvature = offset time in ms */
int fl_scan_upload(fl_scan_t* s, const float* xyzi, const float* offset_ms, int n);
/* ImuProcess::UndistortPcl, the sort (:234) and the backward pass (:312-346)         IMU_Processing.hpp:216-346
 * imu_pose22: IMUpose, n_pose x Pose6D (msg/Pose6D.msg: offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9] row-major),
 * filled by the caller's forward propagation (:244-301, esekf::predict stays on the host);
 * x26_end: kf_state.get_x() after the last predict (:303).  Points stay in HBM, time-sorted (stable). */
int fl_scan_undistort(fl_scan_t* s, const double* imu_pose22, int n_pose, const double* x26_end);
/* downSizeFilterSurf.setInputCloud(feats_undistort); .filter(*feats_down_body)         laserMapping.cpp:904-905
 * = pcl::VoxelGrid<PointType> with leaf (l,l,l) (laserMapping.cpp:811): centroid of every occupied cell, output in
 * ascending cell index.  Returns feats_down_size (>= 0) or an error (< 0). */
int fl_scan_voxel_downsample(fl_scan_t* s, float leaf_size);
/* which 0: the raw / de-skewed cloud, 1: the down-sampled cloud; returns the cloud's size (writes at most cap points) */
int fl_scan_download(fl_scan_t* s, int which, float* out_xyzi, int cap);
/* fl_filter_update on the down-sampled cloud of `s` without a host hop */
int fl_filter_update_scan(fl_filter_t* f, fl_scan_t* s, double* x26, double* P, double R, double* solve_time_s);

/* ------------------------------------------------------------------ local-map cube
 * lasermap_fov_segment()                                            laserMapping.cpp:229-277
 * LocalMap_Points / Localmap_Initialized (:229-230) live in the handle; cube_len = cube_side_length (:774),
 * det_range = mapping/det_range (:775). */
typedef struct fl_localmap fl_localmap_t;
int fl_localmap_create(fl_localmap_t** out, double cube_len, float det_range);
int fl_localmap_destroy(fl_localmap_t* l);
/* One call per scan with pos_lid (:236).  Returns |cub_needrm| (0..3); boxes6_out (may be NULL, room for 3 boxes)
 * receives cub_needrm; with map != NULL also runs ikdtree.Delete_Point_Boxes(cub_needrm) (:275) and stores
 * kdtree_delete_counter in *n_deleted (may be NULL). */
int fl_localmap_segment(fl_localmap_t* l, fl_map_t* map, const double* pos_lid, float* boxes6_out, int* n_deleted);
int fl_localmap_get(fl_localmap_t* l, float* box6);   /* LocalMap_Points as (min xyz, max xyz) */

/* ------------------------------------------------------------------ multi-GPU (no reference counterpart)
 * scan points are sharded across ranks, the map is replicated, the 92 normal-equation doubles
 * are all-reduced once per pass (NCCL over NVLink) and every rank solves redundantly. */
int fl_comm_unique_id(void* out128);
int fl_filter_comm_init(fl_filter_t* f, int nranks, int rank, const void* unique_id128);
int fl_filter_set_shard(fl_filter_t* f, int q_begin, int q_end);
/* Same exchange without NCCL, fused into the residual kernel: each rank exports its mailbox as a 64-byte CUDA-IPC
 * handle (fl_filter_p2p_handle), the application all-gathers the handles, fl_filter_p2p_connect maps the peers.
 * Per pass every rank stores its 92 sums straight into its peers' mailboxes over NVLink as epoch-tagged 8-byte words
 * (the data is the flag) and adds the slots of its own mailbox in rank order: every rank ends with the same bits. */
int fl_filter_p2p_handle(fl_filter_t* f, void* out64);
int fl_filter_p2p_connect(fl_filter_t* f, int nranks, int rank, const void* handles /* nranks x 64 bytes */);

#ifdef __cplusplus
}
#endif
#endif /* FASTLIO_B200_H */
