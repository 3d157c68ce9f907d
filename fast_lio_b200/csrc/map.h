// Host-side owner of the device point map (the B200 counterpart of KD_TREE<PointType>,
// reference include/ikd-Tree/ikd_Tree.h:48-341).  All methods return fl::Status.
#pragma once
#include <vector>

#include "map.cuh"

namespace fl {

struct DeviceBuffer {
    void* ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t want);      // grow-only (x1.5), contents NOT preserved
    void release();
    template <class T> T* as() const { return static_cast<T*>(ptr); }
};

class Map {
public:
    Map(int device, float downsample_size);
    ~Map();
    int init();

    // KD_TREE::Build (ikd_Tree.cpp:409-423).  pts: n x (x, y, z, intensity), host memory.
    int build(const float* pts_xyzi, int n);
    // same, from points already resident on this device
    int build_device(const float4* d_pts_xyzi, int n);
    // KD_TREE::Nearest_Search, batched (ikd_Tree.cpp:426-461); host buffers
    int knn(const float* q_xyzi, int nq, int k, float* out_pts, float* out_d2, int* out_cnt);
    // KD_TREE::Delete_Point_Boxes (ikd_Tree.cpp:632-658); returns the number of points invalidated in *deleted
    int delete_boxes(const float* boxes6, int nb, int* deleted);
    // KD_TREE::Add_Points (ikd_Tree.cpp:478-573); *added = the reference's return value
    int add_points(const float* pts_xyzi, int n, bool downsample_on, int* added);
    int add_points_device(const float4* d_pts_xyzi, int n, bool downsample_on, int* added);
    // KD_TREE::Add_Point_Boxes (ikd_Tree.cpp:576-603) / acquire_removed_points (:661-676)
    int add_boxes(const float* boxes6, int nb, int* revived);
    int acquire_removed(float* out_xyzi, int cap, int* n_out);
    // all valid points, unordered (flatten(Root_Node, ..., NOT_RECORD), ikd_Tree.cpp:1627-1658)
    int flatten(float* out_xyzi, int cap, int* n_out);
    // re-sort every valid point into fresh, evenly filled leaves (ikd-Tree's Rebuild, ikd_Tree.cpp:736-764)
    int rebuild();
    // re-list every live slot in the hashed cell directory (map.cuh); done by build / rebuild, and when inserts crowd it
    int build_directory();
    void set_cell_directory(bool on, float cell_size) { dir_enabled_ = on; cell_override_ = cell_size; }
    int dir_rebuild_count() const { return n_dir_rebuilds_; }
    int dir_stats(int* out6) const;
    // recompute every AABB from the valid points (after deletions)
    int refit();

    int size() const { return n_valid_ + n_tomb_; }       // KD_TREE::size()  (valid + lazily deleted)
    int validnum() const { return n_valid_; }             // KD_TREE::validnum()
    void set_downsample(float v) { downsample_ = v; }
    float downsample() const { return downsample_; }
    int tree_range(float* box6);

    const MapView& view() const { return v_; }
    cudaStream_t stream() const { return stream_; }
    int device() const { return device_; }
    int rebuild_count() const { return n_rebuilds_; }
    int overflow_leaves() const;

    int fill_ = 24;            // slots populated per leaf at (re)build time; the rest absorb inserts
    int min_pool_ = 65536;     // minimum overflow-pool size in leaves (32 MB)
    float rebuild_overflow_frac_ = 0.05f;   // rebuild when overflow leaves exceed this fraction of main leaves

private:
    int ensure_capacity(int n_points);
    int build_from_sorted(const float4* d_src, int n);
    int insert_device(const float4* d_pts, int n);
    int maybe_rebuild();

    int device_;
    float downsample_;
    cudaStream_t stream_ = nullptr;
    MapView v_;
    int n_valid_ = 0, n_tomb_ = 0, n_rebuilds_ = 0;
    bool built_ = false;

    DeviceBuffer pts_, payload_, next_, counters_, dir_tab_, dir_lists_, removed_, ins_slots_, dir_fix_;
    bool record_removed_ = false;
    int n_removed_ = 0;
    bool dir_enabled_ = true;
    float cell_override_ = 0.f;           // 0: cell edge = 2 x downsample size
    size_t dir_min_cap_ = 0, dir_min_pool_ = 0;
    int n_dir_rebuilds_ = 0;
    DeviceBuffer ebox_[MAX_LEVELS];
    DeviceBuffer segid_, segtab_[2], bbox_;        // k-d partition build scratch
    DeviceBuffer src_, keys_in_, keys_out_, vals_in_, vals_out_, cub_tmp_, scratch_, scratch2_, scratch3_;
    int* h_counters_ = nullptr;     // pinned mirror of the device counters
};

}  // namespace fl
