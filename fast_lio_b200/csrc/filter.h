// Host-side owner of the per-scan iterated-EKF measurement update on the device: the B200
// counterpart of esekfom::esekf<state_ikfom,12,input_ikfom>::update_iterated_dyn_share_modified
// (reference include/IKFoM_toolkit/esekfom/esekfom.hpp:1619-1931) with h_share_model
// (reference src/laserMapping.cpp:638-754) bound as a fused device measurement model.
#pragma once
#include "lie.cuh"
#include "map.h"

namespace fl {

// Per-pass record, same layout as the oracle's OraclePassLog (tests compare them field by field).
struct PassLog {
    int searched, valid, effct, converged;
    double res_sum;
    double HtH[144];
    double Hth[12];
    double x_after[XLEN];
};

// Device-resident control block of one update (what the reference keeps in locals of
// update_iterated_dyn_share_modified and in members x_, P_ of esekf).
struct FilterCtl {
    int iter;            // loop variable i, starts at -1            (esekfom.hpp:1633)
    int t;               // converged-step counter                   (esekfom.hpp:1624)
    int converge;        // dyn_share.converge: run the kNN on the next pass
    int done;            // the update has returned
    int n_pass;          // passes executed so far
    int max_iter;        // maximum_iter
    int extrinsic_est;   // extrinsic_est_en (laserMapping.cpp:739)
    int error;           // device-side failure: 1 singular system, 2 a peer never delivered its sums, 3 a block gave up waiting
    int ticket;          // k_residual: blocks finished so far (the last one reduces and solves)
    int gen;             // k_update: passes published so far (the workers of the next pass spin on it)
    double R;            // LASER_POINT_COV passed as R
    double limit[NDOF];
    double x[XLEN];
    double x_prop[XLEN];
    double P[NDOF * NDOF];
    double P_prop[NDOF * NDOF];
    long long prof[16];  // clock64() stamps of the last solve_pass (thread 0), for tuning
    double x_search[XLEN];   // the state the last searching pass used (Nearest_Points belong to it)
    FilterCtl* host_mirror;  // page-locked, device-mapped copy on the host: the pass that ends the update stores the result there
};

struct ScanView {
    const float4* body;      // [Q] body-frame points (x, y, z, intensity)      feats_down_body
    float4* nearest;         // [Q * 5] neighbours of the last search pass      Nearest_Points
    int* nearest_cnt;        // [Q]
    unsigned char* selected; // [Q] point_selected_surf (persists across passes, trap T3)
    float4* normvec;         // [Q] (n, pd2)                                    normvec
    float4* plane;           // [Q] pabcd of the last search pass's plane fit (reused by the passes that do not search)
    double* srange;          // [Q] sqrt(|p_body|) of the score (laserMapping.cpp:681), cached by the searching passes
    int q_begin, q_end;      // this rank's shard of the scan
    int Q;
};

struct NcclApi;

// Peer-memory exchange of the per-pass sums (fused into k_residual's solver block): every rank owns a
// mailbox [2 epoch parities][nranks][96] values, each two epoch-tagged 8-byte words; peers store into it over NVLink.
constexpr int P2P_MAX_RANKS = 8;
struct P2PState {
    double* peer_mail[P2P_MAX_RANKS];                 // mailbox of rank r (mapped through CUDA IPC)
    unsigned long long* peer_flag[P2P_MAX_RANKS];     // its flag array
    unsigned long long* peer_bar[P2P_MAX_RANKS];      // rank r's barrier slots (k_p2p_barrier: aligns the ranks before a timed step)
    unsigned long long epoch;                         // passes exchanged so far (identical on all ranks)
    unsigned long long bar_epoch;                     // barriers passed so far
    int nranks, rank;
};

class Filter {
public:
    Filter(Map* map, int max_points);
    ~Filter();
    int init();
    int set_params(int max_iter, const double* limit23, int extrinsic_est_en);
    void set_search_mode(int mode) { search_mode_ = mode; }
    void set_pdl(bool on) { pdl_ = on; }
    void set_solver(int mode) { solver_ = mode; }       // 1 (default): one ne x ne solve; 0: the reference's two 23x23 inversions, literally

    // whole update with host buffers (scan H2D, state H2D, passes, state D2H)
    int update(const float* body_xyzi, int nq, double* x26, double* P, double R, double* solve_time_s);
    // same with feats_down_body already in HBM on the map's device (ScanFrontEnd::down_device())
    int update_device(const float4* d_body, int nq, double* x26, double* P, double R, double* solve_time_s);
    // pieces, for device-resident benchmarking / pipelines
    int upload_scan(const float* body_xyzi, int nq);
    int set_scan_device(const float4* d_body, int nq);
    int upload_state(const double* x26, const double* P, double R, bool snapshot = true);
    int restore_state();                               // device-side copy of the last uploaded state -> control block
    int run_passes();                                  // enqueue every pass on the stream (no sync)
    int launch_search_only();
    int launch_residual_only();
    int download_state(double* x26, double* P, int* n_pass);
    int sync();

    // map_incremental (laserMapping.cpp:427-474) on the device: classify every scan point with the final
    // state and its cached neighbours, then Add_Points(PointToAdd, true) + Add_Points(PointNoNeedDownsample, false)
    int map_incremental(double filter_size_map_min, int ekf_inited, int* n_to_add, int* n_no_downsample, int* added);
    int get_nearest(float* out_pts, int* out_cnt, int nq);
    // multi-GPU: Nearest_Points of the points outside this rank's shard (searched by their own rank during the update) are
    // recomputed here, with the state of the last searching pass, before anything reads the whole scan's neighbours
    int complete_neighbours();
    int get_selected(unsigned char* out, int nq);
    int get_pass_logs(PassLog* out, int cap, int* n);
    // multi-GPU: scan points sharded across ranks, map replicated, one all-reduce per pass
    int comm_init(int nranks, int rank, const void* nccl_unique_id_128);
    int set_shard(int q_begin, int q_end);             // default: the whole scan
    // fused all-reduce over NVLink peer memory instead of NCCL (one kernel per pass)
    int p2p_local_handle(void* out64);
    int p2p_connect(int nranks, int rank, const void* handles64);
    int p2p_barrier();                                 // device-side rendezvous of all ranks on the stream (no-op on one rank)

    // 1 (default): the whole update in one persistent fused kernel (k_update, update.cuh); 0: the two-kernels-per-pass chain
    // (k_search / k_search_c + k_residual) it grew out of, kept for A/B and for solver mode 0
    void set_fused(bool on) { fused_ = on; }
    bool fused() const { return fused_ && solver_ == 1; }
    int gpu_launches() const { return launches_; }
    const long long* host_ns() const { return host_ns_; }
    const float4* nearest_device() const { return scan_.nearest; }
    const ScanView& scan() const { return scan_; }
    cudaStream_t stream() const { return map_->stream(); }
    Map* map() const { return map_; }
    const FilterCtl* ctl_device() const { return ctl_.as<FilterCtl>(); }

private:
    int reserve(int nq);
    int update_any(const float* body_xyzi, const float4* d_body, int nq, double* x26, double* P, double R, double* solve_time_s);
    Map* map_;
    int max_points_;
    int max_iter_ = 4;
    double limit_[NDOF];
    int extrinsic_est_ = 0;
    int solver_ = 1;
    bool mirror_ = true;               // the last pass stores the result into the page-locked host block itself
    bool pdl_ = true;                  // programmatic dependent launch between the kernels of a scan
    int search_occ_ = 5;               // resident k_search blocks per SM the kernel is compiled for
    int search_mode_ = 1;              // 1 (default): one lane per query through the cell directory (k_search_c); 0: one warp per query through the BVH (k_search)
    ScanView scan_;
    DeviceBuffer srange_;
    DeviceBuffer body_, nearest_, nearest_cnt_, selected_, normvec_, plane_, partials_, red_, ctl_, ctl0_, logs_;
    DeviceBuffer mi_world_, mi_flag_add_, mi_flag_no_, mi_list_add_, mi_list_no_, mi_tmp_, mi_counts_;
    FilterCtl* h_ctl_ = nullptr;       // pinned staging
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    int sms_ = 0, search_grid_max_ = 0, max_resid_grid_ = 0, resid_grid_ = 1;
    bool fused_ = true;
    DeviceBuffer pub_;                 // k_update's publication block
    unsigned launch_nonce_ = 0;
    int upd_capacity_[2] = {0, 0};     // co-resident k_update<EXTR> blocks on this device
    int launch_update(int max_passes, int mode, int search_only);
    int launches_ = 0;
    long long host_ns_[4] = {0, 0, 0, 0};
    bool shard_set_ = false;
    bool neighbours_complete_ = true;
    // NCCL (resolved lazily with dlopen so that single-GPU use needs no NCCL at all)
    NcclApi* nccl_ = nullptr;
    void* comm_ = nullptr;
    int nranks_ = 1, rank_ = 0;
    DeviceBuffer mailbox_, p2p_;
    void* peer_ptr_[P2P_MAX_RANKS] = {nullptr};
    bool p2p_on_ = false;
};

}  // namespace fl
