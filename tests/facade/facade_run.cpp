// EXECUTES the two drop-in facades on a GPU (tests/test_gpu_boundary.py builds and runs it): the calls are the ones
// src/laserMapping.cpp makes -- KD_TREE<PointType>::Build / Nearest_Search / Add_Points / Delete_Point_Boxes / size /
// validnum / acquire_removed_points and esekf::init_dyn_share / change_x / change_P /
// update_iterated_dyn_share_modified / get_x / get_P -- on data the Python side wrote, results written back for comparison
// with the ctypes path and the CPU oracle.
//   usage: facade_run <in.bin> <out.bin>
//   in : int32 n_map, n_scan, max_iter, extrinsic; float32 map[n_map*4], scan[n_scan*4]; float64 x[26], P[529], R
//   out: float64 x[26], P[529], solve_time; int32 size, validnum, add_ret, del_ret, n_removed, nearest_total, n_knn;
//        float32 knn_d2[n_knn*5], knn_pts[n_knn*5*4]
#include <ikd-Tree/ikd_Tree.h>
#include <IKFoM_toolkit/esekfom/esekfom_b200.hpp>

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

typedef pcl::PointXYZINormal PointType;
typedef KD_TREE<PointType>::PointVector PointVector;
struct Cloud { std::vector<PointType> points; };

template <class T> static bool rd(FILE* f, T* p, size_t n) { return fread(p, sizeof(T), n, f) == n; }
template <class T> static void wr(FILE* f, const T* p, size_t n) { fwrite(p, sizeof(T), n, f); }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: facade_run in out\n"); return 2; }
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) { perror("in"); return 2; }
    int hdr[4];
    if (!rd(fi, hdr, 4)) return 2;
    const int n_map = hdr[0], n_scan = hdr[1], max_iter = hdr[2], extr = hdr[3];
    std::vector<float> map((size_t)n_map * 4), scan((size_t)n_scan * 4);
    double x[26], P[529], R;
    if (!rd(fi, map.data(), map.size()) || !rd(fi, scan.data(), scan.size()) || !rd(fi, x, 26) || !rd(fi, P, 529) || !rd(fi, &R, 1)) return 2;
    fclose(fi);

    // ---- KD_TREE<PointType> ikdtree;  (laserMapping.cpp:120, :913, :919)
    KD_TREE<PointType> ikdtree(0.5f, 0.6f, 0.5f);
    if (!ikdtree.ok()) { fprintf(stderr, "no device map: %s\n", KD_TREE<PointType>::last_error()); return 3; }
    ikdtree.set_downsample_param(0.5f);
    PointVector cloud(n_map);
    for (int i = 0; i < n_map; i++) { cloud[i].x = map[4 * i]; cloud[i].y = map[4 * i + 1]; cloud[i].z = map[4 * i + 2]; cloud[i].intensity = map[4 * i + 3]; }
    ikdtree.Build(cloud);
    if (ikdtree.Root_Node == nullptr) return 4;                              // :909
    // ---- Nearest_Search, one point at a time as h_share_model does (:670)
    const int n_knn = n_map < 16 ? n_map : 16;
    std::vector<float> knn_d2((size_t)n_knn * 5, 0.f), knn_pts((size_t)n_knn * 20, 0.f);
    for (int i = 0; i < n_knn; i++) {
        PointType q = cloud[(size_t)i * 7 % n_map];
        q.x += 0.11f; q.y -= 0.07f;
        PointVector near;
        std::vector<float> d2;
        ikdtree.Nearest_Search(q, 5, near, d2);
        for (size_t j = 0; j < near.size(); j++) {
            knn_d2[(size_t)i * 5 + j] = d2[j];
            knn_pts[((size_t)i * 5 + j) * 4] = near[j].x; knn_pts[((size_t)i * 5 + j) * 4 + 1] = near[j].y;
            knn_pts[((size_t)i * 5 + j) * 4 + 2] = near[j].z; knn_pts[((size_t)i * 5 + j) * 4 + 3] = near[j].intensity;
        }
    }
    // ---- esekf (laserMapping.cpp:131, :826-828, :950-961)
    esekfom::esekf_b200<state_ikfom, 12, input_ikfom> kf;
    state_ikfom s;
    for (int i = 0; i < 3; i++) { s.pos[i] = x[i]; s.offset_T_L_I[i] = x[11 + i]; s.vel[i] = x[14 + i]; s.bg[i] = x[17 + i]; s.ba[i] = x[20 + i]; s.grav.vec[i] = x[23 + i]; }
    for (int i = 0; i < 4; i++) { s.rot.coeffs()[i] = x[3 + i]; s.offset_R_L_I.coeffs()[i] = x[7 + i]; }
    esekfom::esekf_b200<state_ikfom, 12, input_ikfom>::cov Pm;
    for (int i = 0; i < 23; i++) for (int j = 0; j < 23; j++) Pm(i, j) = P[i * 23 + j];
    kf.change_x(s);
    kf.change_P(Pm);
    double epsi[23];
    for (int i = 0; i < 23; i++) epsi[i] = 0.001;
    kf.init_dyn_share(0, 0, 0, 0, max_iter, epsi);
    if (kf.bind_map(ikdtree.handle(), extr != 0, n_scan > 0 ? n_scan : 1) != FL_OK) { fprintf(stderr, "bind_map: %s\n", fl_last_error()); return 5; }
    auto feats_down_body = std::make_shared<Cloud>();
    feats_down_body->points.resize(n_scan);
    for (int i = 0; i < n_scan; i++) {
        feats_down_body->points[i].x = scan[4 * i]; feats_down_body->points[i].y = scan[4 * i + 1];
        feats_down_body->points[i].z = scan[4 * i + 2]; feats_down_body->points[i].intensity = scan[4 * i + 3];
    }
    kf.bind_scan(feats_down_body);
    double solve_time = 0.0;
    kf.update_iterated_dyn_share_modified(R, solve_time);
    const state_ikfom& so = kf.get_x();
    double xo[26], Po[529];
    for (int i = 0; i < 3; i++) { xo[i] = so.pos[i]; xo[11 + i] = so.offset_T_L_I[i]; xo[14 + i] = so.vel[i]; xo[17 + i] = so.bg[i]; xo[20 + i] = so.ba[i]; xo[23 + i] = so.grav[i]; }
    for (int i = 0; i < 4; i++) { xo[3 + i] = so.rot.coeffs()[i]; xo[7 + i] = so.offset_R_L_I.coeffs()[i]; }
    for (int i = 0; i < 23; i++) for (int j = 0; j < 23; j++) Po[i * 23 + j] = kf.get_P()(i, j);
    std::vector<PointVector> Nearest_Points;
    kf.fetch_nearest(Nearest_Points);                                        // map_incremental reads it (:438-460)
    int nearest_total = 0;
    for (const auto& v : Nearest_Points) nearest_total += (int)v.size();
    // ---- map maintenance (:225, :275, :470-471)
    PointVector history;
    ikdtree.acquire_removed_points(history);
    PointVector to_add(cloud.begin(), cloud.begin() + (n_map < 500 ? n_map : 500));
    for (auto& p : to_add) { p.x += 0.05f; p.z += 0.02f; }
    const int add_ret = ikdtree.Add_Points(to_add, true);
    std::vector<BoxPointType> boxes(1);
    for (int a = 0; a < 3; a++) { boxes[0].vertex_min[a] = -3.f; boxes[0].vertex_max[a] = 3.f; }
    const int del_ret = ikdtree.Delete_Point_Boxes(boxes);
    ikdtree.acquire_removed_points(history);
    const int out_i[7] = {ikdtree.size(), ikdtree.validnum(), add_ret, del_ret, (int)history.size(), nearest_total, n_knn};
    if (ikdtree.failed()) { fprintf(stderr, "a KD_TREE call failed: %s\n", KD_TREE<PointType>::last_error()); return 6; }

    FILE* fo = fopen(argv[2], "wb");
    if (!fo) { perror("out"); return 2; }
    wr(fo, xo, 26); wr(fo, Po, 529); wr(fo, &solve_time, 1); wr(fo, out_i, 7);
    wr(fo, knn_d2.data(), knn_d2.size()); wr(fo, knn_pts.data(), knn_pts.size());
    fclose(fo);
    printf("facade_run ok: size %d validnum %d add %d del %d removed %d nearest %d solve %.3f ms\n", out_i[0], out_i[1], out_i[2], out_i[3], out_i[4], out_i[5], 1e3 * solve_time);
    return 0;
}
