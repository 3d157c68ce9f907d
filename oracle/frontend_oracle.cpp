// TEST INFRASTRUCTURE ONLY -- CPU restatement of the three steps that sit immediately in front of the
// measurement update.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg may load this; the product (fast_lio_b200/) never does.
//
//   oracle_fov_segment   lasermap_fov_segment()              src/laserMapping.cpp:229-277
//   oracle_undistort     ImuProcess::UndistortPcl, backward  src/IMU_Processing.hpp:232-234, 312-346
//                        pass only (the forward IMU propagation that fills IMUpose stays with the caller)
//   oracle_voxelgrid     pcl::VoxelGrid<PointXYZINormal>::filter as called at src/laserMapping.cpp:904-905.
//                        PCL is a third-party dependency that is NOT in /root/reference (README.md:73
//                        asks for PCL >= 1.8); this restates the published algorithm of
//                        pcl/filters/impl/voxel_grid.hpp (applyFilter, PCL 1.8-1.12: getMinMax3D,
//                        min_b/div_b/divb_mul, idx = ijk . divb_mul, sort by idx, CentroidPoint per cell,
//                        output in ascending idx) -- "parity unpinned": no PCL build exists here to pin it.
//                        PCL sorts with std::sort, which leaves the order INSIDE a cell unspecified; the
//                        float sums here run in ascending input index (a stable sort).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct V3 { double v[3]; double& operator[](int i) { return v[i]; } double operator[](int i) const { return v[i]; } };
struct Q4 { double x, y, z, w; };
struct M3 { double m[9]; };

inline V3 cross(const V3& a, const V3& b) { return V3{{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}}; }
inline V3 add(const V3& a, const V3& b) { return V3{{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }
inline V3 sub(const V3& a, const V3& b) { return V3{{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
inline V3 scale(const V3& a, double s) { return V3{{a[0] * s, a[1] * s, a[2] * s}}; }
// Eigen::QuaternionBase::_transformVector: uv = 2 * (q.vec x v); v + w*uv + q.vec x uv
inline V3 qrot(const Q4& q, const V3& v) {
    V3 qv{{q.x, q.y, q.z}};
    V3 uv = cross(qv, v);
    uv = add(uv, uv);
    return add(add(v, scale(uv, q.w)), cross(qv, uv));
}
inline Q4 conj(const Q4& q) { return Q4{-q.x, -q.y, -q.z, q.w}; }
inline V3 mv(const M3& a, const V3& x) {
    V3 r;
    for (int i = 0; i < 3; i++) r[i] = a.m[i * 3 + 0] * x[0] + a.m[i * 3 + 1] * x[1] + a.m[i * 3 + 2] * x[2];
    return r;
}
inline M3 mm(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[i * 3 + 0] * b.m[0 * 3 + j] + a.m[i * 3 + 1] * b.m[1 * 3 + j] + a.m[i * 3 + 2] * b.m[2 * 3 + j];
    return r;
}
// Exp(ang_vel, dt)  include/so3_math.h:37-58
inline M3 exp_rodrigues(const V3& w, double dt) {
    M3 eye = {{1, 0, 0, 0, 1, 0, 0, 0, 1}};
    double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (!(n > 0.0000001)) return eye;
    V3 a{{w[0] / n, w[1] / n, w[2] / n}};
    M3 K = {{0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0}};
    double ang = n * dt, s = std::sin(ang), c1 = 1.0 - std::cos(ang);
    M3 cK;
    for (int i = 0; i < 9; i++) cK.m[i] = c1 * K.m[i];
    M3 cKK = mm(cK, K);
    M3 r;
    for (int i = 0; i < 9; i++) r.m[i] = (eye.m[i] + s * K.m[i]) + cKK.m[i];
    return r;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// lasermap_fov_segment (laserMapping.cpp:229-277).  State: LocalMap_Points (:229) + Localmap_Initialized (:230).
// boxes: up to 3 x (min xyz, max xyz) -- cub_needrm (:100).  Returns |cub_needrm|.
struct OracleLocalMap {
    float vmin[3], vmax[3];
    int initialized;
};

int oracle_fov_segment(OracleLocalMap* lm, const double* pos_lid, double cube_len, float det_range, float* boxes) {
    const float MOV_THRESHOLD = 1.5f;                                    // :78
    if (!lm->initialized) {                                              // :238-245
        for (int i = 0; i < 3; i++) {
            lm->vmin[i] = float(pos_lid[i] - cube_len / 2.0);
            lm->vmax[i] = float(pos_lid[i] + cube_len / 2.0);
        }
        lm->initialized = 1;
        return 0;
    }
    float edge[3][2];
    bool need_move = false;
    for (int i = 0; i < 3; i++) {                                        // :248-252
        edge[i][0] = float(std::fabs(pos_lid[i] - double(lm->vmin[i])));
        edge[i][1] = float(std::fabs(pos_lid[i] - double(lm->vmax[i])));
        if (edge[i][0] <= MOV_THRESHOLD * det_range || edge[i][1] <= MOV_THRESHOLD * det_range) need_move = true;
    }
    if (!need_move) return 0;
    float nmin[3], nmax[3];
    memcpy(nmin, lm->vmin, sizeof(nmin));
    memcpy(nmax, lm->vmax, sizeof(nmax));
    float mov_dist = float(std::max((cube_len - 2.0 * MOV_THRESHOLD * det_range) * 0.5 * 0.9, double(det_range * (MOV_THRESHOLD - 1))));   // :256
    int nb = 0;
    for (int i = 0; i < 3; i++) {                                        // :257-270
        float bmin[3], bmax[3];
        memcpy(bmin, lm->vmin, sizeof(bmin));
        memcpy(bmax, lm->vmax, sizeof(bmax));
        if (edge[i][0] <= MOV_THRESHOLD * det_range) {
            nmax[i] -= mov_dist;
            nmin[i] -= mov_dist;
            bmin[i] = lm->vmax[i] - mov_dist;
        } else if (edge[i][1] <= MOV_THRESHOLD * det_range) {
            nmax[i] += mov_dist;
            nmin[i] += mov_dist;
            bmax[i] = lm->vmin[i] + mov_dist;
        } else {
            continue;
        }
        memcpy(boxes + nb * 6, bmin, sizeof(bmin));
        memcpy(boxes + nb * 6 + 3, bmax, sizeof(bmax));
        nb++;
    }
    memcpy(lm->vmin, nmin, sizeof(nmin));
    memcpy(lm->vmax, nmax, sizeof(nmax));
    return nb;
}

// ---------------------------------------------------------------------------------------------
// UndistortPcl, backward pass (IMU_Processing.hpp:312-346).
//   pts    n x 4 float (x,y,z,intensity), ALREADY sorted by offset time (:234), compensated in place
//   t_ms   n float: PointType::curvature = offset time in milliseconds
//   poses  n_pose x 22 doubles: Pose6D (msg/Pose6D.msg) offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9]
//   x26    state at the frame end (imu_state after the last predict, :303)
void oracle_undistort(float* pts, const float* t_ms, int n, const double* poses, int n_pose, const double* x26) {
    if (n == 0) return;                                                  // :310
    V3 pos_end{{x26[0], x26[1], x26[2]}};
    Q4 rot{x26[3], x26[4], x26[5], x26[6]}, offR{x26[7], x26[8], x26[9], x26[10]};
    V3 offT{{x26[11], x26[12], x26[13]}};
    int ip = n - 1;                                                      // it_pcl = end() - 1
    for (int kp = n_pose - 1; kp >= 1; kp--) {                           // :314
        const double* head = poses + (size_t)(kp - 1) * 22;
        const double* tail = poses + (size_t)kp * 22;
        M3 R_imu;
        memcpy(R_imu.m, head + 13, sizeof(R_imu.m));
        V3 vel{{head[7], head[8], head[9]}}, pos{{head[10], head[11], head[12]}};
        V3 acc{{tail[1], tail[2], tail[3]}}, gyr{{tail[4], tail[5], tail[6]}};
        for (; t_ms[ip] / double(1000) > head[0]; ip--) {                // :325
            double dt = t_ms[ip] / double(1000) - head[0];
            M3 R_i = mm(R_imu, exp_rodrigues(gyr, dt));
            V3 P_i{{pts[ip * 4 + 0], pts[ip * 4 + 1], pts[ip * 4 + 2]}};
            // pos + vel*dt + 0.5*acc*dt*dt - pos_end, evaluated left to right per coefficient (:335)
            V3 T_ei;
            for (int c = 0; c < 3; c++) T_ei[c] = ((pos[c] + vel[c] * dt) + ((0.5 * acc[c]) * dt) * dt) - pos_end[c];
            V3 inner = add(mv(R_i, add(qrot(offR, P_i), offT)), T_ei);
            V3 comp = qrot(conj(offR), sub(qrot(conj(rot), inner), offT));   // :336
            pts[ip * 4 + 0] = float(comp[0]);
            pts[ip * 4 + 1] = float(comp[1]);
            pts[ip * 4 + 2] = float(comp[2]);
            if (ip == 0) break;                                          // :343
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pcl::VoxelGrid<PointT>::applyFilter with leaf (l,l,l), downsample_all_data_ = true, min_points_per_voxel_ = 0,
// no filter field, dense input.  Only the fields the path keeps (x,y,z,intensity) are produced.
// out must hold n points.  Returns the number of output points.
int oracle_voxelgrid(const float* pts, int n, float leaf, float* out) {
    if (n == 0) return 0;
    const float inv = 1.0f / leaf;                                       // inverse_leaf_size_ = 1 / leaf_size_
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; i++)                                          // getMinMax3D
        for (int c = 0; c < 3; c++) {
            mn[c] = std::min(mn[c], pts[i * 4 + c]);
            mx[c] = std::max(mx[c], pts[i * 4 + c]);
        }
    int64_t d[3];
    for (int c = 0; c < 3; c++) d[c] = int64_t((mx[c] - mn[c]) * inv) + 1;
    if (d[0] * d[1] * d[2] > int64_t(INT32_MAX)) {                       // "Leaf size is too small": output = input
        memcpy(out, pts, sizeof(float) * 4 * (size_t)n);
        return n;
    }
    int min_b[3], max_b[3], div_b[3];
    for (int c = 0; c < 3; c++) {
        min_b[c] = int(std::floor(mn[c] * inv));
        max_b[c] = int(std::floor(mx[c] * inv));
        div_b[c] = max_b[c] - min_b[c] + 1;
    }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    struct Cell { unsigned idx; int pt; };
    std::vector<Cell> cells((size_t)n);
    for (int i = 0; i < n; i++) {
        int ijk[3];
        for (int c = 0; c < 3; c++) ijk[c] = int(std::floor(pts[i * 4 + c] * inv) - float(min_b[c]));
        cells[i] = Cell{unsigned(ijk[0] * mul[0] + ijk[1] * mul[1] + ijk[2] * mul[2]), i};
    }
    std::stable_sort(cells.begin(), cells.end(), [](const Cell& a, const Cell& b) { return a.idx < b.idx; });
    int total = 0;
    for (size_t first = 0; first < cells.size();) {
        size_t last = first + 1;
        while (last < cells.size() && cells[last].idx == cells[first].idx) ++last;
        // CentroidPoint: AccumulatorXYZ (Vector3f sum, / n) and AccumulatorIntensity (float sum, / n)
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (size_t li = first; li < last; li++)
            for (int c = 0; c < 4; c++) s[c] += pts[cells[li].pt * 4 + c];
        const float cnt = float(last - first);
        for (int c = 0; c < 4; c++) out[total * 4 + c] = s[c] / cnt;
        total++;
        first = last;
    }
    return total;
}

}  // extern "C"
