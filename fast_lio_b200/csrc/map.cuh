// Device-side view of the point map and the warp-cooperative traversals over it.
//
// The reference's ikd-Tree (include/ikd-Tree/ikd_Tree.{h,cpp}) is a pointer-chasing
// binary k-d tree with 176-byte nodes (ikd_Tree.h:59-82) and ~33 dependent node visits
// per query.  Here the map is a flattened, implicit 32-ary bounding-volume hierarchy over
// a k-d partition of the points:
//
//   level 0   leaf buckets: 32 slots of float4 (x, y, z, flag) = 512 B, one coalesced
//             warp load.  At (re)build time the points are split top-down at the median
//             of the longest axis -- the rule of KD_TREE::BuildTree (ikd_Tree.cpp:679-733)
//             -- until each cell holds `fill` points; the free slots absorb inserts.
//   level k   entity e of level k (a leaf for k = 0, an internal node otherwise) has its
//             AABB in ebox[k][e] (two float4: lo, hi).  Node j of level k+1 owns entities
//             32j .. 32j+31 of level k: five binary k-d levels collapse into one 32-wide
//             node, there are no child pointers, and a warp tests all children of a node
//             with one box per lane.
//
// A query is served by one warp: every lane holds the query, the k best candidates live
// one per lane (lanes 0..k-1, ascending), candidates are ranked with the hardware warp
// reduction (redux.sync) and the traversal state of each level lives in registers (the
// recursion over the <= 7 levels is unrolled at compile time): no stack in memory.
#pragma once
#include "common.cuh"

namespace fl {

struct MapView {
    float4* pts;                     // [leaf_cap * 32]  (x, y, z, as_float(flag)); flag 1 = valid
    float* payload;                  // [leaf_cap * 32]  intensity of the point in that slot
    int* next;                       // [leaf_cap]       overflow chain of a leaf, -1 = none
    float4* ebox[MAX_LEVELS];        // [count * 2]      AABB (lo, hi) of each entity of level k
    int count[MAX_LEVELS + 1];       // entities per level; count[n_levels] == 1 (the root)
    int n_levels;                    // number of internal levels (>= 1)
    int n_main;                      // leaves addressed by the implicit tree (== count[0])
    int leaf_cap;                    // allocated leaves (main + overflow pool)
};

__device__ __forceinline__ bool slot_valid(const float4& p) { return __float_as_int(p.w) == 1; }

// ----------------------------------------------------------------------------- k-best list
// Lane j < K holds the j-th best (distance, slot); other lanes hold +inf.  Mirrors MANUAL_HEAP
// + PointType_CMP (ikd_Tree.h:93-201) in effect: a candidate enters only if strictly closer
// than the current k-th best (ikd_Tree.cpp:1088); `w` caches that k-th best, warp-uniform.
struct KBest {
    float d;
    int idx;
    float w;
    int n;          // entries filled so far (warp-uniform)
    __device__ __forceinline__ void init() { d = INFINITY; idx = -1; w = INFINITY; n = 0; }
    // nd, nidx warp-uniform, nd < w
    __device__ __forceinline__ void insert(float nd, int nidx, int lane) {
        const float up_d = __shfl_up_sync(FULL, d, 1);
        const int up_i = __shfl_up_sync(FULL, idx, 1);
        if (lane < KNN_K && nd < d) {
            const bool take_prev = lane > 0 && nd < up_d;
            d = take_prev ? up_d : nd;
            idx = take_prev ? up_i : nidx;
        }
        w = __shfl_sync(FULL, d, KNN_K - 1);
    }
};

// Squared distances are non-negative floats, so their bit patterns order like the values and
// +inf (0x7f800000) can serve as the "no candidate" marker of the integer warp reductions.
constexpr unsigned INF_BITS = 0x7f800000u;

// Visit one leaf bucket (and its overflow chain): each lane scores one slot, then the
// (at most K) improving candidates are extracted in ascending order.
__device__ __forceinline__ void knn_leaf(const MapView& m, int leaf, float qx, float qy, float qz,
                                         KBest& kb, int lane) {
    while (leaf >= 0) {
        const float4 p = __ldg(&m.pts[leaf * LEAF + lane]);
        const int nxt = __ldg(&m.next[leaf]);
        unsigned key = slot_valid(p) ? __float_as_uint(sq_dist3(qx, qy, qz, p.x, p.y, p.z)) : INF_BITS;
        if (kb.n == 0) {
            // empty list (the first leaf of a query): the r-th smallest goes straight to lane r -- no merge
            unsigned best = INF_BITS;
#pragma unroll
            for (int r = 0; r < KNN_K; r++) {
                best = __reduce_min_sync(FULL, key);
                if (best == INF_BITS) break;
                const int src = __ffs(__ballot_sync(FULL, key == best)) - 1;
                if (lane == r) { kb.d = __uint_as_float(best); kb.idx = leaf * LEAF + src; }
                if (lane == src) key = INF_BITS;
                kb.n = r + 1;
            }
            if (kb.n == KNN_K) kb.w = __uint_as_float(best);
            leaf = nxt;
            continue;
        }
#pragma unroll 1
        for (int it = 0; it < KNN_K; it++) {
            const unsigned best = __reduce_min_sync(FULL, key);
            if (best >= __float_as_uint(kb.w)) break;       // nothing strictly closer than the k-th best is left (covers the marker)
            const int src = __ffs(__ballot_sync(FULL, key == best)) - 1;
            kb.insert(__uint_as_float(best), leaf * LEAF + src, lane);
            if (lane == src) key = INF_BITS;
        }
        leaf = nxt;
    }
}

// Visit node `node` of level L (its children are entities of level L-1).  Children are
// taken nearest-first and re-tested against the shrinking k-th best distance after each
// return -- the pruning rule of KD_TREE::Search (ikd_Tree.cpp:1097-1243).  The ordering key is
// the box distance with its low 5 mantissa bits traded for the lane id; a child is skipped only
// when even that rounded-DOWN distance is not below the k-th best, so pruning never drops a
// child that could matter (it may visit one whose distance ties the bound within 2^-18).
template <int L>
__device__ __forceinline__ void knn_node(const MapView& m, int node, float qx, float qy, float qz,
                                         KBest& kb, int lane) {
    const int e = node * FAN + lane;
    unsigned key = 0xffffffffu;
    if (e < m.count[L - 1]) {
        const float4 lo = __ldg(&m.ebox[L - 1][2 * e]);
        const float4 hi = __ldg(&m.ebox[L - 1][2 * e + 1]);
        key = (__float_as_uint(box_dist3(qx, qy, qz, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z)) & ~31u) | (unsigned)lane;
    }
#pragma unroll 1
    while (true) {
        const unsigned best = __reduce_min_sync(FULL, key);
        if ((best & ~31u) >= __float_as_uint(kb.w)) break;
        const int c = best & 31;
        if (lane == c) key = 0xffffffffu;
        if constexpr (L == 1) knn_leaf(m, node * FAN + c, qx, qy, qz, kb, lane);
        else knn_node<L - 1>(m, node * FAN + c, qx, qy, qz, kb, lane);
    }
}

// The root owns up to 64 entities (two per lane), which saves the two nearly empty top levels a
// strict 32-ary hierarchy would have (1 M points: 41 667 leaves -> 1 303 -> 41 -> root).
constexpr int ROOT_FAN = 64;

template <int L>
__device__ __forceinline__ void knn_root(const MapView& m, float qx, float qy, float qz, KBest& kb, int lane) {
    const int cnt = m.count[L - 1];
    unsigned k0 = 0xffffffffu, k1 = 0xffffffffu;
    if (lane < cnt) {
        const float4 lo = __ldg(&m.ebox[L - 1][2 * lane]), hi = __ldg(&m.ebox[L - 1][2 * lane + 1]);
        k0 = (__float_as_uint(box_dist3(qx, qy, qz, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z)) & ~63u) | (unsigned)lane;
    }
    if (lane + 32 < cnt) {
        const float4 lo = __ldg(&m.ebox[L - 1][2 * (lane + 32)]), hi = __ldg(&m.ebox[L - 1][2 * (lane + 32) + 1]);
        k1 = (__float_as_uint(box_dist3(qx, qy, qz, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z)) & ~63u) | (unsigned)(lane + 32);
    }
#pragma unroll 1
    while (true) {
        const unsigned best = __reduce_min_sync(FULL, min(k0, k1));
        if ((best & ~63u) >= __float_as_uint(kb.w)) break;
        const int c = best & 63;
        if (lane == (c & 31)) { if (c < 32) k0 = 0xffffffffu; else k1 = 0xffffffffu; }
        if constexpr (L == 1) knn_leaf(m, c, qx, qy, qz, kb, lane);
        else knn_node<L - 1>(m, c, qx, qy, qz, kb, lane);
    }
}

// Exact k-nearest-neighbour search for one query by one warp.
__device__ __forceinline__ void knn_query(const MapView& m, float qx, float qy, float qz, KBest& kb, int lane) {
    kb.init();
    switch (m.n_levels) {
        case 1: knn_root<1>(m, qx, qy, qz, kb, lane); break;
        case 2: knn_root<2>(m, qx, qy, qz, kb, lane); break;
        case 3: knn_root<3>(m, qx, qy, qz, kb, lane); break;
        case 4: knn_root<4>(m, qx, qy, qz, kb, lane); break;
        case 5: knn_root<5>(m, qx, qy, qz, kb, lane); break;
        default: knn_root<6>(m, qx, qy, qz, kb, lane); break;
    }
}

// ----------------------------------------------------------------------------- box query
// Half-open membership test of Search_by_range / Delete_by_range (ikd_Tree.cpp:796,1263):
//   vertex_min <= p < vertex_max  on every axis.
__device__ __forceinline__ bool in_box(const float4& p, const float* bmin, const float* bmax) {
    return bmin[0] <= p.x && bmax[0] > p.x && bmin[1] <= p.y && bmax[1] > p.y && bmin[2] <= p.z && bmax[2] > p.z;
}
// AABB-vs-box rejection, the negation of ikd_Tree.cpp:1253-1258
__device__ __forceinline__ bool box_overlaps(const float4& lo, const float4& hi, const float* bmin, const float* bmax) {
    if (bmax[0] <= lo.x || bmin[0] > hi.x) return false;
    if (bmax[1] <= lo.y || bmin[1] > hi.y) return false;
    if (bmax[2] <= lo.z || bmin[2] > hi.z) return false;
    return true;
}

// Functor interface: f.leaf(leaf_index) is called warp-uniformly for every main leaf whose
// AABB overlaps the box (the functor walks the overflow chain itself).
template <int L, class F>
__device__ __forceinline__ void box_node(const MapView& m, int node, const float* bmin, const float* bmax, F& f, int lane) {
    const int e = node * FAN + lane;
    bool hit = false;
    if (e < m.count[L - 1]) {
        const float4 lo = __ldg(&m.ebox[L - 1][2 * e]);
        const float4 hi = __ldg(&m.ebox[L - 1][2 * e + 1]);
        hit = box_overlaps(lo, hi, bmin, bmax);
    }
    unsigned mask = __ballot_sync(FULL, hit);
    while (mask) {
        const int c = __ffs(mask) - 1;
        mask &= mask - 1;
        if constexpr (L == 1) f.leaf(node * FAN + c);
        else box_node<L - 1, F>(m, node * FAN + c, bmin, bmax, f, lane);
    }
}
template <int L, class F>
__device__ __forceinline__ void box_root(const MapView& m, const float* bmin, const float* bmax, F& f, int lane) {
    const int cnt = m.count[L - 1];
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int e = lane + 32 * half;
        bool hit = false;
        if (e < cnt) {
            const float4 lo = __ldg(&m.ebox[L - 1][2 * e]), hi = __ldg(&m.ebox[L - 1][2 * e + 1]);
            hit = box_overlaps(lo, hi, bmin, bmax);
        }
        unsigned mask = __ballot_sync(FULL, hit);
        while (mask) {
            const int c = __ffs(mask) - 1 + 32 * half;
            mask &= mask - 1;
            if constexpr (L == 1) f.leaf(c);
            else box_node<L - 1, F>(m, c, bmin, bmax, f, lane);
        }
    }
}
template <class F>
__device__ __forceinline__ void box_query(const MapView& m, const float* bmin, const float* bmax, F& f, int lane) {
    switch (m.n_levels) {
        case 1: box_root<1, F>(m, bmin, bmax, f, lane); break;
        case 2: box_root<2, F>(m, bmin, bmax, f, lane); break;
        case 3: box_root<3, F>(m, bmin, bmax, f, lane); break;
        case 4: box_root<4, F>(m, bmin, bmax, f, lane); break;
        case 5: box_root<5, F>(m, bmin, bmax, f, lane); break;
        default: box_root<6, F>(m, bmin, bmax, f, lane); break;
    }
}

}  // namespace fl

// ============================================================================= thread-per-query traversal
// Same tree, same arithmetic, same result as knn_query -- but one THREAD per query.  A warp-per-
// query walk spends most of its instructions on lanes whose child box is pruned (32 boxes tested,
// ~1.3 useful) and on warp-wide ranking; with one query per lane every lane does useful work and
// the only loss is divergence between neighbouring queries.  MEASURED (B200, 30k queries vs 1M points):
// 112.7 us against 39.6 us for the warp-per-query kernel -- every load instruction touches 32 different
// 16-byte sectors (32 L1 wavefronts), which costs more than the idle lanes of the cooperative walk.
// Kept as a selectable alternative (fl_filter_set_search) and as evidence for the design choice.
namespace fl {

struct TKBest {                      // ascending, replicated per thread
    float d[KNN_K];
    int idx[KNN_K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < KNN_K; i++) { d[i] = INFINITY; idx[i] = -1; }
    }
    __device__ __forceinline__ float w() const { return d[KNN_K - 1]; }
    __device__ __forceinline__ void insert(float nd, int nidx) {       // nd < d[K-1]
#pragma unroll
        for (int i = KNN_K - 1; i > 0; i--) {
            const bool shift = nd < d[i - 1];
            const bool here = !shift && nd < d[i];
            d[i] = shift ? d[i - 1] : (here ? nd : d[i]);
            idx[i] = shift ? idx[i - 1] : (here ? nidx : idx[i]);
        }
        if (nd < d[0]) { d[0] = nd; idx[0] = nidx; }
    }
};

__device__ __forceinline__ void tknn_leaf(const MapView& m, int leaf, float qx, float qy, float qz, TKBest& kb) {
    while (leaf >= 0) {
        const float4* base = m.pts + (size_t)leaf * LEAF;
#pragma unroll 4
        for (int s = 0; s < LEAF; s++) {
            const float4 p = __ldg(&base[s]);
            if (slot_valid(p)) {
                const float d = sq_dist3(qx, qy, qz, p.x, p.y, p.z);
                if (d < kb.w()) kb.insert(d, leaf * LEAF + s);
            }
        }
        leaf = __ldg(&m.next[leaf]);
    }
}

// children of `node` at level L (entities of level L-1), visited in ascending (distance, index)
// order; the runner-up of a scan is remembered so that the usual "nearest child, then nothing
// else qualifies" case needs a single pass over the boxes.
template <int L>
__device__ __forceinline__ void tknn_node(const MapView& m, int first, int n_child, float qx, float qy, float qz, TKBest& kb) {
    const float4* boxes = m.ebox[L - 1] + 2 * (size_t)first;
    unsigned long long last = 0ull;          // keys are > 0: (distance bits + 1) << 6 | child
    unsigned long long runner = ~0ull;
    bool have_runner = false;
    while (true) {
        unsigned long long best = ~0ull;
        if (have_runner) {
            best = runner; have_runner = false;
            // the runner-up's box distance has not changed; only the bound has
            if (__uint_as_float((unsigned)((best >> 6) - 1ull)) >= kb.w()) break;   // and every other child is farther
        } else {
            runner = ~0ull;
#pragma unroll 4
            for (int c = 0; c < n_child; c++) {
                const float4 lo = __ldg(&boxes[2 * c]), hi = __ldg(&boxes[2 * c + 1]);
                const float d = box_dist3(qx, qy, qz, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z);
                if (d < kb.w()) {
                    const unsigned long long key = (((unsigned long long)__float_as_uint(d) + 1ull) << 6) | (unsigned long long)c;
                    if (key > last) {
                        if (key < best) { runner = best; best = key; }
                        else if (key < runner) runner = key;
                    }
                }
            }
            if (best == ~0ull) break;
            have_runner = runner != ~0ull;
        }
        last = best;
        const int c = (int)(best & 63ull);
        if constexpr (L == 1) tknn_leaf(m, first + c, qx, qy, qz, kb);
        else {
            const int child = first + c;
            const int cnt = min(FAN, m.count[L - 2] - child * FAN);
            tknn_node<L - 1>(m, child * FAN, cnt, qx, qy, qz, kb);
        }
    }
}

__device__ __forceinline__ void tknn_query(const MapView& m, float qx, float qy, float qz, TKBest& kb) {
    kb.init();
    const int top = m.count[m.n_levels - 1];          // the root owns up to 64 entities
    switch (m.n_levels) {
        case 1: tknn_node<1>(m, 0, top, qx, qy, qz, kb); break;
        case 2: tknn_node<2>(m, 0, top, qx, qy, qz, kb); break;
        case 3: tknn_node<3>(m, 0, top, qx, qy, qz, kb); break;
        case 4: tknn_node<4>(m, 0, top, qx, qy, qz, kb); break;
        case 5: tknn_node<5>(m, 0, top, qx, qy, qz, kb); break;
        default: tknn_node<6>(m, 0, top, qx, qy, qz, kb); break;
    }
}

}  // namespace fl

