// ============================================================================
// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/fastlio_oracle.cpp header).
//
// extern "C" wrapper around the reference's OWN, UNMODIFIED ikd-Tree
// (/root/reference/include/ikd-Tree/ikd_Tree.{h,cpp}, submodule pinned at
// e2e3f4e in /root/reference/.SUBMODULES.json).  The reference sources are
// compiled where they lie (see oracle/Makefile, target _ref/libikdtree_ref.so);
// nothing from them is copied into this repository.  Only this wrapper and the
// PCL/Eigen shim header (oracle/shim/pcl/point_types.h) are ours.
//
// Used to (1) pin the kNN / Add_Points / Delete_Point_Boxes semantics of the
// CUDA map against the real thing and (2) as the "reference" CPU baseline.
// ============================================================================
#include <ikd_Tree.h>

#include <cstring>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef pcl::PointXYZINormal PointType;
typedef KD_TREE<PointType> Tree;
typedef Tree::PointVector PointVector;

static inline PointType mk(const float* p4) {
    PointType p;
    p.x = p4[0]; p.y = p4[1]; p.z = p4[2]; p.intensity = p4[3];
    return p;
}

extern "C" {

// KD_TREE is ~80 MB (trap T10): always heap-allocate.
void* ref_kdtree_create(float delete_param, float balance_param, float box_length) {
    return new Tree(delete_param, balance_param, box_length);
}
void ref_kdtree_destroy(void* h) { delete static_cast<Tree*>(h); }
void ref_kdtree_set_downsample(void* h, float v) { static_cast<Tree*>(h)->set_downsample_param(v); }

void ref_kdtree_build(void* h, const float* pts4, int n) {
    PointVector v(n);
    for (int i = 0; i < n; i++) v[i] = mk(&pts4[size_t(i) * 4]);
    static_cast<Tree*>(h)->Build(v);
}

int ref_kdtree_size(void* h) { return static_cast<Tree*>(h)->size(); }
int ref_kdtree_validnum(void* h) {
    Tree* t = static_cast<Tree*>(h);
    int v = t->validnum();
    for (int tries = 0; v < 0 && tries < 100000; tries++) { usleep(100); v = t->validnum(); }   // -1 while a root rebuild holds the lock
    return v;
}

// Same signature as oracle_knn_fn (oracle/fastlio_oracle.cpp).
int ref_kdtree_knn1(void* h, const float* q_xyz, int k, float* out_pts4, float* out_d2) {
    Tree* t = static_cast<Tree*>(h);
    PointType q; q.x = q_xyz[0]; q.y = q_xyz[1]; q.z = q_xyz[2];
    PointVector near;
    std::vector<float> d2;
    t->Nearest_Search(q, k, near, d2);
    int cnt = int(near.size());
    for (int i = 0; i < cnt; i++) {
        out_pts4[i * 4 + 0] = near[i].x; out_pts4[i * 4 + 1] = near[i].y;
        out_pts4[i * 4 + 2] = near[i].z; out_pts4[i * 4 + 3] = near[i].intensity;
        out_d2[i] = d2[i];
    }
    return cnt;
}

void ref_kdtree_knn(void* h, const float* q4, int nq, int k, float* out_pts4, float* out_d2, int* out_cnt, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < nq; i++)
        out_cnt[i] = ref_kdtree_knn1(h, &q4[size_t(i) * 4], k, &out_pts4[size_t(i) * k * 4], &out_d2[size_t(i) * k]);
}

int ref_kdtree_add(void* h, const float* pts4, int n, int downsample_on) {
    PointVector v(n);
    for (int i = 0; i < n; i++) v[i] = mk(&pts4[size_t(i) * 4]);
    return static_cast<Tree*>(h)->Add_Points(v, downsample_on != 0);
}

int ref_kdtree_delete_boxes(void* h, const float* boxes6, int nb) {
    std::vector<BoxPointType> b(nb);
    for (int i = 0; i < nb; i++) {
        for (int a = 0; a < 3; a++) { b[i].vertex_min[a] = boxes6[i * 6 + a]; b[i].vertex_max[a] = boxes6[i * 6 + 3 + a]; }
    }
    return static_cast<Tree*>(h)->Delete_Point_Boxes(b);
}

// All currently valid points (flatten(Root_Node, ..., NOT_RECORD) skips deleted ones).
// Returns the count; writes at most cap points.
int ref_kdtree_flatten(void* h, float* out4, int cap) {
    Tree* t = static_cast<Tree*>(h);
    // let a pending background rebuild settle so that flatten sees a stable tree
    (void)ref_kdtree_validnum(h);
    PointVector v;
    t->flatten(t->Root_Node, v, NOT_RECORD);
    int n = int(v.size());
    for (int i = 0; i < n && i < cap; i++) {
        out4[size_t(i) * 4 + 0] = v[i].x; out4[size_t(i) * 4 + 1] = v[i].y;
        out4[size_t(i) * 4 + 2] = v[i].z; out4[size_t(i) * 4 + 3] = v[i].intensity;
    }
    return n;
}

}  // extern "C"
