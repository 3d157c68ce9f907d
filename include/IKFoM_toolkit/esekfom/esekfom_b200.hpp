// Drop-in adaptor: esekfom::esekf_b200<state, process_noise_dof, input> derives from the
// reference's own esekfom::esekf (include/IKFoM_toolkit/esekfom/esekfom.hpp:105-2004, which stays
// on the include path: predict(), get_x(), change_x() ... are inherited untouched) and re-routes
//
//     update_iterated_dyn_share_modified(double R, double& solve_time)        esekfom.hpp:1619-1931
//
// to the fused device path (fl_filter_update): h_share_model (laserMapping.cpp:638-754) is not
// called back point by point any more -- its arithmetic runs inside the kernels against the device
// map.  Changes in laserMapping.cpp (shown in INTEGRATION.md), 4 lines:
//
//     esekfom::esekf_b200<state_ikfom, 12, input_ikfom> kf;                   // :131
//     kf.bind_map(ikdtree.handle(), extrinsic_est_en);                        // after :828
//     kf.bind_scan(feats_down_body);                                          // before :960
//     kf.fetch_nearest(Nearest_Points);                                       // after :960 (map_incremental reads it)
//
// Requires state == state_ikfom (include/use-ikfom.hpp:12-21): members pos, rot, offset_R_L_I,
// offset_T_L_I, vel, bg, ba, grav with Eigen-style coefficient access.
#pragma once
#include <vector>

#include <IKFoM_toolkit/esekfom/esekfom.hpp>   // the reference's own header (stays on the include path)
#include <stdio.h>

#include "../../fastlio_b200.h"

namespace esekfom {

template <typename state, int process_noise_dof, typename input = state, typename measurement = state, int measurement_noise_dof = 0>
class esekf_b200 : public esekf<state, process_noise_dof, input, measurement, measurement_noise_dof> {
    typedef esekf<state, process_noise_dof, input, measurement, measurement_noise_dof> base;

public:
    typedef typename base::cov cov;
    typedef typename base::scalar_type scalar_type;

    esekf_b200(const state& x = state(), const cov& P = cov::Identity()) : base(x, P) {}
    ~esekf_b200() { if (filter_) fl_filter_destroy(filter_); }

    // same signature as the reference (esekfom.hpp:238); records maximum_iter / limit for the device path
    template <class F, class FX, class FW, class H>
    void init_dyn_share(F f_in, FX f_x_in, FW f_w_in, H h_dyn_share_in, int maximum_iteration, scalar_type limit_vector[state::DOF]) {
        base::init_dyn_share(f_in, f_x_in, f_w_in, h_dyn_share_in, maximum_iteration, limit_vector);
        max_iter_ = maximum_iteration;
        for (int i = 0; i < 23; i++) limit_[i] = limit_vector[i];
        push_params();
    }

    int bind_map(fl_map_t* map, bool extrinsic_est_en, int max_points = 100000) {
        extrinsic_est_ = extrinsic_est_en;
        if (filter_) { fl_filter_destroy(filter_); filter_ = nullptr; }
        int rc = fl_filter_create(&filter_, map, max_points);
        if (rc == FL_OK) rc = push_params();
        return rc;
    }

    // feats_down_body: any container of points with x, y, z, intensity (pcl::PointCloud<PointType>::points)
    template <class Cloud>
    void bind_scan(const Cloud& cloud) {
        scan_.resize(cloud->points.size() * 4);
        for (size_t i = 0; i < cloud->points.size(); i++) {
            scan_[4 * i] = cloud->points[i].x; scan_[4 * i + 1] = cloud->points[i].y;
            scan_[4 * i + 2] = cloud->points[i].z; scan_[4 * i + 3] = cloud->points[i].intensity;
        }
    }

    // esekfom.hpp:1619
    void update_iterated_dyn_share_modified(double R, double& solve_time) {
        if (!filter_) { base::update_iterated_dyn_share_modified(R, solve_time); return; }   // not bound: the reference's own path
        double x[26], P[23 * 23];
        pack(x, P);
        if (fl_filter_update(filter_, scan_.data(), (int)(scan_.size() / 4), x, P, R, &solve_time) != FL_OK) {
            fprintf(stderr, "esekf_b200: %s\n", fl_last_error());
            return;                                                  // state untouched, like an invalid measurement
        }
        unpack(x, P);
    }

    // The same update on a cloud that fl_scan_voxel_downsample left in HBM (no host hop for the points).
    void update_from(fl_scan_t* scan, double R, double& solve_time) {
        if (!filter_ || !scan) { fprintf(stderr, "esekf_b200::update_from: bind_map() first\n"); return; }
        double x[26], P[23 * 23];
        pack(x, P);
        if (fl_filter_update_scan(filter_, scan, x, P, R, &solve_time) != FL_OK) {
            fprintf(stderr, "esekf_b200: %s\n", fl_last_error());
            return;
        }
        unpack(x, P);
    }

    // map_incremental() (laserMapping.cpp:427-474) on the device, from the state and neighbours of the last update;
    // returns the value of ikdtree.Add_Points(PointToAdd, true) (:470) or < 0
    int map_incremental(double filter_size_map_min, bool flg_EKF_inited) {
        int out[3] = {0, 0, 0};
        if (!filter_) return FL_ERR_STATE;
        int rc = fl_filter_map_incremental(filter_, filter_size_map_min, flg_EKF_inited ? 1 : 0, out);
        return rc == FL_OK ? out[2] : rc;
    }

    // Nearest_Points of the last search pass (laserMapping.cpp:102) for map_incremental (:438-460)
    template <class PointVector>
    void fetch_nearest(std::vector<PointVector>& Nearest_Points) {
        const int nq = (int)(scan_.size() / 4);
        std::vector<float> pts((size_t)nq * 20);
        std::vector<int> cnt(nq);
        Nearest_Points.resize(nq);
        if (!filter_ || fl_filter_get_nearest(filter_, pts.data(), cnt.data(), nq) != FL_OK) return;
        for (int i = 0; i < nq; i++) {
            Nearest_Points[i].clear();
            for (int j = 0; j < cnt[i]; j++) {
                typename PointVector::value_type p;
                p.x = pts[(size_t)i * 20 + 4 * j]; p.y = pts[(size_t)i * 20 + 4 * j + 1]; p.z = pts[(size_t)i * 20 + 4 * j + 2];
                p.intensity = pts[(size_t)i * 20 + 4 * j + 3];
                Nearest_Points[i].push_back(p);
            }
        }
    }

    fl_filter_t* handle() const { return filter_; }

private:
    // state_ikfom <-> the 26-double layout of fastlio_b200.h
    void pack(double* x, double* P) const {
        const state& s = this->get_x();
        for (int i = 0; i < 3; i++) { x[i] = s.pos[i]; x[11 + i] = s.offset_T_L_I[i]; x[14 + i] = s.vel[i]; x[17 + i] = s.bg[i]; x[20 + i] = s.ba[i]; x[23 + i] = s.grav[i]; }
        for (int i = 0; i < 4; i++) { x[3 + i] = s.rot.coeffs()[i]; x[7 + i] = s.offset_R_L_I.coeffs()[i]; }
        const cov& Pm = this->get_P();
        for (int i = 0; i < 23; i++) for (int j = 0; j < 23; j++) P[i * 23 + j] = Pm(i, j);
    }
    void unpack(const double* x, const double* P) {
        state out = this->get_x();
        for (int i = 0; i < 3; i++) { out.pos[i] = x[i]; out.offset_T_L_I[i] = x[11 + i]; out.vel[i] = x[14 + i]; out.bg[i] = x[17 + i]; out.ba[i] = x[20 + i]; out.grav.vec[i] = x[23 + i]; }
        for (int i = 0; i < 4; i++) { out.rot.coeffs()[i] = x[3 + i]; out.offset_R_L_I.coeffs()[i] = x[7 + i]; }
        cov Pout;
        for (int i = 0; i < 23; i++) for (int j = 0; j < 23; j++) Pout(i, j) = P[i * 23 + j];
        this->change_x(out);
        this->change_P(Pout);
    }
    int push_params() { return filter_ ? fl_filter_set_params(filter_, max_iter_, limit_, extrinsic_est_ ? 1 : 0) : FL_OK; }
    fl_filter_t* filter_ = nullptr;
    std::vector<float> scan_;
    int max_iter_ = 4;
    double limit_[23] = {0};
    bool extrinsic_est_ = false;
};

}  // namespace esekfom
