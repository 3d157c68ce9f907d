// Compile-and-link check of the two drop-in facades (run by tests/test_facade_compile.py).
#include <ikd-Tree/ikd_Tree.h>
#include <IKFoM_toolkit/esekfom/esekfom_b200.hpp>
#include <memory>
struct Cloud { std::vector<pcl::PointXYZINormal> points; };
int main() {
    KD_TREE<pcl::PointXYZINormal> tree(0.5f, 0.6f, 0.5f);       // needs a GPU at run time; here we only link
    KD_TREE<pcl::PointXYZINormal>::PointVector v(3), out;
    std::vector<float> d;
    std::vector<BoxPointType> boxes(1);
    if (tree.handle()) {
        tree.Build(v); tree.Nearest_Search(v[0], 5, out, d); tree.Add_Points(v, true); tree.Delete_Point_Boxes(boxes);
        tree.flatten(tree.Root_Node, out, NOT_RECORD);
    }
    esekfom::esekf_b200<state_ikfom, 12, input_ikfom> kf;
    double lim[23] = {0}, st = 0;
    kf.init_dyn_share(0, 0, 0, 0, 3, lim);
    auto cloud = std::make_shared<Cloud>();
    kf.bind_scan(cloud);
    if (tree.handle()) kf.bind_map(tree.handle(), false);
    kf.update_iterated_dyn_share_modified(0.001, st);
    std::vector<KD_TREE<pcl::PointXYZINormal>::PointVector> nearest;
    kf.fetch_nearest(nearest);
    // the resident front end: raw cloud -> de-skew -> voxel grid -> update -> map_incremental, plus the local-map cube
    fl_scan_t* scan = nullptr;
    fl_localmap_t* cube = nullptr;
    if (tree.handle() && fl_scan_create(&scan, tree.handle()) == FL_OK) {
        float xyzi[4] = {1, 2, 3, 4}, t_ms[1] = {0};
        double pose[22] = {0}, x26[26] = {0};
        fl_scan_upload(scan, xyzi, t_ms, 1);
        fl_scan_undistort(scan, pose, 1, x26);
        fl_scan_voxel_downsample(scan, 0.5f);
        kf.update_from(scan, 0.001, st);
        kf.map_incremental(0.5, true);
        fl_scan_destroy(scan);
    }
    if (fl_localmap_create(&cube, 1000.0, 100.0f) == FL_OK) {
        double pos[3] = {0, 0, 0};
        int deleted = 0;
        fl_localmap_segment(cube, tree.handle(), pos, nullptr, &deleted);
        fl_localmap_destroy(cube);
    }
    return tree.size() + tree.validnum() + (int)nearest.size();
}
