// ============================================================================
// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/fastlio_oracle.cpp header).
//
// Port ("kind": "port") of the static part of the reference's ikd-Tree:
//   KD_TREE::BuildTree      include/ikd-Tree/ikd_Tree.cpp:679-733
//   KD_TREE::Search         include/ikd-Tree/ikd_Tree.cpp:1062-1244
//   KD_TREE::Nearest_Search include/ikd-Tree/ikd_Tree.cpp:426-461
//   calc_dist/calc_box_dist include/ikd-Tree/ikd_Tree.cpp:1683-1709
//   MANUAL_HEAP/PointType_CMP include/ikd-Tree/ikd_Tree.h:93-201
// Array-based nodes instead of heap nodes; no delete flags, no rebuild thread
// (the static configs never mutate the tree).  It exists (a) as the kNN
// back-end of the oracle on machines where oracle/_ref could not be built and
// (b) as an independent cross-check of oracle/_ref.  The mutation semantics
// (Add_Points / Delete_Point_Boxes) are restated in tests/semantics.py.
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct P4 { float x, y, z, w; };

struct Node {
    P4 point;
    int axis;
    int left, right;            // -1 == nullptr
    float lo[3], hi[3];         // node_range_{x,y,z}
};

struct Tree {
    std::vector<Node> nodes;
    int root = -1;
};

struct Cand { float d; P4 p; };
// PointType_CMP::operator<  (ikd_Tree.h:102-108)
inline bool cmp_less(const Cand& a, const Cand& b) {
    if (std::fabs(a.d - b.d) < 1e-10) return a.p.x < b.p.x;
    return a.d < b.d;
}

// MANUAL_HEAP (max-heap on cmp_less), ikd_Tree.h:111-201
struct Heap {
    Cand h[16]; int n = 0; int cap;
    explicit Heap(int c) : cap(c) {}
    void push(const Cand& c) {
        if (n >= cap) return;
        int i = n; h[n++] = c;
        Cand tmp = h[i];
        while (i > 0) { int a = (i - 1) / 2; if (cmp_less(h[a], tmp)) { h[i] = h[a]; i = a; } else break; }
        h[i] = tmp;
    }
    void pop() {
        if (!n) return;
        h[0] = h[n - 1]; n--;
        int i = 0, l = 1; Cand tmp = h[0];
        while (l < n) {
            if (l + 1 < n && cmp_less(h[l], h[l + 1])) l++;
            if (cmp_less(tmp, h[l])) { h[i] = h[l]; i = l; l = 2 * i + 1; } else break;
        }
        h[i] = tmp;
    }
};

inline float calc_dist(const float* q, const P4& p) {                 // ikd_Tree.cpp:1683-1688
    return (q[0] - p.x) * (q[0] - p.x) + (q[1] - p.y) * (q[1] - p.y) + (q[2] - p.z) * (q[2] - p.z);
}
inline float calc_box_dist(const Tree& t, int n, const float* q) {    // ikd_Tree.cpp:1691-1709
    if (n < 0) return INFINITY;
    const Node& nd = t.nodes[n];
    float m = 0.0f;
    for (int a = 0; a < 3; a++) {
        if (q[a] < nd.lo[a]) m += (q[a] - nd.lo[a]) * (q[a] - nd.lo[a]);
        if (q[a] > nd.hi[a]) m += (q[a] - nd.hi[a]) * (q[a] - nd.hi[a]);
    }
    return m;
}

int build(Tree& t, std::vector<P4>& s, int l, int r) {                // BuildTree
    if (l > r) return -1;
    int mid = (l + r) >> 1;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = l; i <= r; i++) {
        const float c[3] = {s[i].x, s[i].y, s[i].z};
        for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], c[a]); mx[a] = std::max(mx[a], c[a]); }
    }
    int axis = 0;
    float range[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
    for (int a = 1; a < 3; a++) if (range[a] > range[axis]) axis = a;
    auto cmp = [axis](const P4& a, const P4& b) {
        return axis == 0 ? a.x < b.x : (axis == 1 ? a.y < b.y : a.z < b.z);
    };
    std::nth_element(s.begin() + l, s.begin() + mid, s.begin() + r + 1, cmp);
    int id = int(t.nodes.size());
    t.nodes.push_back(Node());
    t.nodes[id].point = s[mid];
    t.nodes[id].axis = axis;
    for (int a = 0; a < 3; a++) { t.nodes[id].lo[a] = mn[a]; t.nodes[id].hi[a] = mx[a]; }   // == Update()'s subtree range
    int L = build(t, s, l, mid - 1);
    int Rr = build(t, s, mid + 1, r);
    t.nodes[id].left = L; t.nodes[id].right = Rr;
    return id;
}

void search(const Tree& t, int n, int k, const float* q, Heap& hp) {   // Search (max_dist = INFINITY)
    if (n < 0) return;
    const Node& nd = t.nodes[n];
    float dist = calc_dist(q, nd.point);
    if (hp.n < k || dist < hp.h[0].d) {
        if (hp.n >= k) hp.pop();
        hp.push(Cand{dist, nd.point});
    }
    float dl = calc_box_dist(t, nd.left, q), dr = calc_box_dist(t, nd.right, q);
    if (hp.n < k || (dl < hp.h[0].d && dr < hp.h[0].d)) {
        if (dl <= dr) {
            search(t, nd.left, k, q, hp);
            if (hp.n < k || dr < hp.h[0].d) search(t, nd.right, k, q, hp);
        } else {
            search(t, nd.right, k, q, hp);
            if (hp.n < k || dl < hp.h[0].d) search(t, nd.left, k, q, hp);
        }
    } else {
        if (dl < hp.h[0].d) search(t, nd.left, k, q, hp);
        if (dr < hp.h[0].d) search(t, nd.right, k, q, hp);
    }
}

}  // namespace

extern "C" {

void* port_kdtree_build(const float* pts4, int n) {
    Tree* t = new Tree();
    std::vector<P4> s(n);
    memcpy(s.data(), pts4, sizeof(P4) * size_t(n));
    t->nodes.reserve(n);
    t->root = build(*t, s, 0, n - 1);
    return t;
}
void port_kdtree_destroy(void* h) { delete static_cast<Tree*>(h); }

// same signature as oracle_knn_fn
int port_kdtree_knn1(void* h, const float* q_xyz, int k, float* out_pts4, float* out_d2) {
    const Tree& t = *static_cast<Tree*>(h);
    Heap hp(2 * k);
    search(t, t.root, k, q_xyz, hp);
    int found = std::min(k, hp.n);
    for (int i = found - 1; i >= 0; i--) {          // Nearest_Search: results nearest-first
        out_pts4[i * 4 + 0] = hp.h[0].p.x; out_pts4[i * 4 + 1] = hp.h[0].p.y;
        out_pts4[i * 4 + 2] = hp.h[0].p.z; out_pts4[i * 4 + 3] = hp.h[0].p.w;
        out_d2[i] = hp.h[0].d;
        hp.pop();
    }
    return found;
}

void port_kdtree_knn(void* h, const float* q4, int nq, int k, float* out_pts4, float* out_d2, int* out_cnt, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < nq; i++)
        out_cnt[i] = port_kdtree_knn1(h, &q4[size_t(i) * 4], k, &out_pts4[size_t(i) * k * 4], &out_d2[size_t(i) * k]);
}

}  // extern "C"
