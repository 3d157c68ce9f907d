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
    return tree.size() + tree.validnum() + (int)nearest.size();
}
