// Scan front end: the steps between the raw LiDAR points and the measurement update, kept in HBM.
//   sort by offset time + per-point de-skew     ImuProcess::UndistortPcl   src/IMU_Processing.hpp:232-234, 312-346
//   voxel-grid down-sampling                    pcl::VoxelGrid::filter     src/laserMapping.cpp:904-905
// and the sliding local-map cube that produces the delete boxes (host arithmetic only)
//   LocalMapCube                                lasermap_fov_segment()     src/laserMapping.cpp:229-277
#pragma once
#include "map.h"

namespace fl {

constexpr int POSE_DOUBLES = 22;   // msg/Pose6D.msg: offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9]

class ScanFrontEnd {
public:
    explicit ScanFrontEnd(Map* map) : map_(map) {}
    ~ScanFrontEnd();
    int init();
    // raw scan: n x (x,y,z,intensity) and the per-point offset time in ms (PointType::curvature), host memory
    int upload(const float* xyzi, const float* offset_ms, int n);
    // stable sort by offset time (:234), then backward propagation of every point to the frame end (:312-346).
    // poses: n_pose x 22 doubles = IMUpose; x26_end: kf_state.get_x() after the last predict (:303)
    int undistort(const double* poses, int n_pose, const double* x26_end);
    // feats_undistort -> feats_down_body; *n_out = number of occupied voxels (one 4-byte read-back)
    int voxel_downsample(float leaf, int* n_out);
    int download(int which, float* out_xyzi, int cap, int* n);   // which 0: raw / undistorted, 1: down-sampled
    const float4* down_device() const { return down_.as<float4>(); }
    int down_count() const { return n_down_; }
    int raw_count() const { return n_raw_; }
    Map* map() const { return map_; }

private:
    Map* map_;
    int n_raw_ = 0, n_down_ = 0;
    DeviceBuffer raw_, raw_alt_, time_, time_alt_, down_, keys_, keys_alt_, vals_, vals_alt_, heads_, pos_, cub_tmp_, ctl_, poses_;
    int* h_count_ = nullptr;        // pinned
};

// lasermap_fov_segment() without its globals: LocalMap_Points (:229) and Localmap_Initialized (:230) live here.
class LocalMapCube {
public:
    LocalMapCube(double cube_len, float det_range) : cube_len_(cube_len), det_range_(det_range) {}
    // returns the number of delete boxes written to boxes6 (<= 3, each min xyz / max xyz) -- cub_needrm
    int slide(const double pos_lid[3], float* boxes6);
    bool initialized() const { return init_; }
    void get(float* box6) const;

private:
    double cube_len_;
    float det_range_;
    float lo_[3] = {0, 0, 0}, hi_[3] = {0, 0, 0};
    bool init_ = false;
};

}  // namespace fl
