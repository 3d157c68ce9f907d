// Device point map: build / refit / search / box-delete / incremental insert.
// B200-native replacement for the subset of KD_TREE<PointType> that FAST-LIO's
// laserMapping.cpp calls (reference include/ikd-Tree/ikd_Tree.cpp).  See map.cuh for the
// memory layout and DESIGN.md for the rationale.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <vector>

#include "map.h"

namespace fl {

// ============================================================================= errors
static thread_local char g_last_error[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_last_error; }

// ============================================================================= DeviceBuffer
int DeviceBuffer::reserve(size_t want) {
    if (want <= bytes) return FL_OK;
    size_t grow = std::max(want, bytes + bytes / 2);
    if (ptr) { FL_CUDA(cudaFree(ptr)); ptr = nullptr; bytes = 0; }
    FL_CUDA(cudaMalloc(&ptr, grow));
    bytes = grow;
    return FL_OK;
}
void DeviceBuffer::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr; bytes = 0;
}

// counters_ layout
enum Counter { C_LEAF_USED = 0, C_DELETED, C_ADDED, C_GROUPS, C_NINSERT, C_TOMB, C_ERROR, C_COMPACT,
               C_DIR_CELLS, C_DIR_POOL, C_DIR_CROWDED, C_DIR_ERROR, C_DIR_WALKED, C_REMOVED, C_REVIVED, C_NFIX, C_COUNT = 16 };

// ============================================================================= kernels
// ----------------------------------------------------------------------------- k-d partition build
// Top-down, level-synchronous: every level (1) bounds each segment, (2) sorts all points by
// (segment, coordinate on the segment's longest axis) with one radix sort, (3) splits each
// segment at a leaf boundary.  A segment owns a contiguous range of sorted points [p0, p1) and a
// contiguous range of leaves [a, b); splits are placed on 32^k-aligned leaf boundaries so that
// the five binary levels below any 32-wide node are exactly its children.
struct Segment { int p0, p1, a, b; };

__device__ __forceinline__ unsigned flip_float(float f) {
    const unsigned u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
// split leaf of a segment covering leaves [a, b), b - a >= 2
// pmax = leaves under one child of the root: above it the root's (up to 64) children are split
// evenly instead of on powers of 32
__host__ __device__ __forceinline__ int split_leaf(int a, int b, long long pmax) {
    const int span = b - a;
    long long p = 1;
    while (p * FAN < span && p * FAN <= pmax) p *= FAN;
    const int nch = (int)((span + p - 1) / p);
    return a + (int)((nch / 2) * p);
}

__global__ void k_kd_init(int n, unsigned* __restrict__ idx, int* __restrict__ segid, Segment* seg, int n_leaves) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { idx[i] = (unsigned)i; segid[i] = 0; }
    if (i == 0) { seg[0].p0 = 0; seg[0].p1 = n; seg[0].a = 0; seg[0].b = n_leaves; }
}
__global__ void k_kd_bbox_init(float* __restrict__ bbox, int nseg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nseg * 6) bbox[i] = (i % 6) < 3 ? INFINITY : -INFINITY;
}
__global__ void k_kd_bbox(const float4* __restrict__ src, const unsigned* __restrict__ idx, const int* __restrict__ segid,
                          int n, float* __restrict__ bbox) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = r < n;
    int s = -1;
    float v[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (live) {
        s = segid[r];
        const float4 p = src[idx[r]];
        v[0] = v[3] = p.x; v[1] = v[4] = p.y; v[2] = v[5] = p.z;
    }
    // positions are ordered by segment: usually the whole warp shares one segment
    const int s0 = __shfl_sync(FULL, s, 0);
    if (__all_sync(FULL, s == s0)) {
        if (s0 < 0) return;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int a = 0; a < 3; a++) {
                v[a] = fminf(v[a], __shfl_xor_sync(FULL, v[a], o));
                v[3 + a] = fmaxf(v[3 + a], __shfl_xor_sync(FULL, v[3 + a], o));
            }
        }
        if (lane < 3) atomic_min_float(&bbox[s0 * 6 + lane], lane == 0 ? v[0] : (lane == 1 ? v[1] : v[2]));
        else if (lane < 6) atomic_max_float(&bbox[s0 * 6 + lane], lane == 3 ? v[3] : (lane == 4 ? v[4] : v[5]));
    } else if (live) {
#pragma unroll
        for (int a = 0; a < 3; a++) { atomic_min_float(&bbox[s * 6 + a], v[a]); atomic_max_float(&bbox[s * 6 + 3 + a], v[3 + a]); }
    }
}
__global__ void k_kd_keys(const float4* __restrict__ src, const unsigned* __restrict__ idx, const int* __restrict__ segid, int n,
                          const float* __restrict__ bbox, unsigned long long* __restrict__ keys, unsigned* __restrict__ vals) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int s = segid[r];
    const float* b = &bbox[s * 6];
    // longest axis, first one wins ties (ikd_Tree.cpp:704-707)
    const float ex = b[3] - b[0], ey = b[4] - b[1], ez = b[5] - b[2];
    int axis = 0; float best = ex;
    if (ey > best) { best = ey; axis = 1; }
    if (ez > best) { axis = 2; }
    const unsigned id = idx[r];
    const float4 p = src[id];
    const float c = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
    keys[r] = ((unsigned long long)(unsigned)s << 32) | flip_float(c);
    vals[r] = id;
}
__global__ void k_kd_split(const Segment* __restrict__ in, Segment* __restrict__ out, int nseg, int fill, long long pmax) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const Segment g = in[s];
    Segment l = g, r;
    r.p0 = r.p1 = g.p1; r.a = r.b = g.b;
    if (g.b - g.a >= 2) {
        const int m = split_leaf(g.a, g.b, pmax);
        long long cut = (long long)g.p0 + (long long)(m - g.a) * fill;
        if (cut > g.p1) cut = g.p1;
        l.p1 = (int)cut; l.b = m;
        r.p0 = (int)cut; r.p1 = g.p1; r.a = m; r.b = g.b;
    }
    out[2 * s] = l;
    out[2 * s + 1] = r;
}
__global__ void k_kd_assign(const unsigned long long* __restrict__ sorted_keys, int n, const Segment* __restrict__ next_seg,
                            int* __restrict__ segid) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int s = (int)(sorted_keys[r] >> 32);
    segid[r] = 2 * s + (r >= next_seg[2 * s].p1 ? 1 : 0);
}

// Scatter the partitioned points into their leaf buckets; clear every other slot, the overflow
// pool and the chains.
__global__ void k_clear_leaves(MapView m) {
    const long long total = (long long)m.leaf_cap * LEAF;
    for (long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x; slot < total;
         slot += (long long)gridDim.x * blockDim.x) {
        m.pts[slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
        m.payload[slot] = 0.f;
        if ((slot % LEAF) == 0) m.next[slot / LEAF] = -1;
    }
}
__global__ void k_fill_leaves(MapView m, const float4* __restrict__ src, const unsigned* __restrict__ idx,
                              const int* __restrict__ segid, const Segment* __restrict__ seg, int n) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const Segment g = seg[segid[r]];
    const float4 p = src[idx[r]];
    const long long slot = (long long)g.a * LEAF + (r - g.p0);
    m.pts[slot] = make_float4(p.x, p.y, p.z, __int_as_float(1));
    m.payload[slot] = p.w;
}

// One warp per main leaf: AABB of the valid points of the leaf and of its overflow chain.
__global__ void k_refit_leaves(MapView m) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int leaf = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; leaf < m.n_main; leaf += warps) {
        float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
        int l = leaf;
        while (l >= 0) {
            const float4 p = m.pts[l * LEAF + lane];
            if (slot_valid(p)) {
                lx = fminf(lx, p.x); ly = fminf(ly, p.y); lz = fminf(lz, p.z);
                hx = fmaxf(hx, p.x); hy = fmaxf(hy, p.y); hz = fmaxf(hz, p.z);
            }
            l = m.next[l];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(FULL, lx, o)); ly = fminf(ly, __shfl_xor_sync(FULL, ly, o));
            lz = fminf(lz, __shfl_xor_sync(FULL, lz, o)); hx = fmaxf(hx, __shfl_xor_sync(FULL, hx, o));
            hy = fmaxf(hy, __shfl_xor_sync(FULL, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(FULL, hz, o));
        }
        if (lane == 0) {
            m.ebox[0][2 * leaf] = make_float4(lx, ly, lz, 0.f);
            m.ebox[0][2 * leaf + 1] = make_float4(hx, hy, hz, 0.f);
        }
    }
}

// One warp per entity of level k (k >= 1): union of its <= 32 children boxes.
__global__ void k_refit_level(MapView m, int k) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < m.count[k]; e += warps) {
        const int c = e * FAN + lane;
        float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
        if (c < m.count[k - 1]) {
            const float4 lo = m.ebox[k - 1][2 * c], hi = m.ebox[k - 1][2 * c + 1];
            lx = lo.x; ly = lo.y; lz = lo.z; hx = hi.x; hy = hi.y; hz = hi.z;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(FULL, lx, o)); ly = fminf(ly, __shfl_xor_sync(FULL, ly, o));
            lz = fminf(lz, __shfl_xor_sync(FULL, lz, o)); hx = fmaxf(hx, __shfl_xor_sync(FULL, hx, o));
            hy = fmaxf(hy, __shfl_xor_sync(FULL, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(FULL, hz, o));
        }
        if (lane == 0) {
            m.ebox[k][2 * e] = make_float4(lx, ly, lz, 0.f);
            m.ebox[k][2 * e + 1] = make_float4(hx, hy, hz, 0.f);
        }
    }
}

// ----------------------------------------------------------------------------- cell directory (map.cuh)
// find the entry of `key`, claiming a free one if the cell is new; returns its table index (0xffffffff: table full)
__device__ __forceinline__ unsigned dir_claim(const CellDir& D, unsigned long long key, int* counters, bool* fresh = nullptr) {
    unsigned s = cell_slot(key, D.cap);
    for (unsigned probes = 0; probes < D.cap; probes++) {
        const unsigned long long old = atomicCAS(&D.tab[s].key, 0ull, key);
        if (old == 0ull) { atomicAdd(&counters[C_DIR_CELLS], 1); if (fresh) *fresh = true; return s; }
        if (old == key) return s;
        s = (s + 1 == D.cap) ? 0u : s + 1;
    }
    atomicExch(&counters[C_DIR_ERROR], 1);      // table full (the host sizes it so that this cannot happen)
    return 0xffffffffu;
}
__device__ __forceinline__ unsigned dir_find(const CellDir& D, unsigned long long key) {
    unsigned s = cell_slot(key, D.cap);
    for (unsigned probes = 0; probes < D.cap; probes++) {
        const unsigned long long k = D.tab[s].key;
        if (k == key) return s;
        if (k == 0ull) break;
        s = (s + 1 == D.cap) ? 0u : s + 1;
    }
    return 0xffffffffu;
}
// the 27 cells whose halo lists a point belongs to: neighbour t of its own cell
__device__ __forceinline__ unsigned long long halo_key(const CellDir& D, const float4& p, int t) {
    return cell_key(cell_coord(p.x, D.inv_cell) + t % 3 - 1, cell_coord(p.y, D.inv_cell) + (t / 3) % 3 - 1, cell_coord(p.z, D.inv_cell) + t / 9 - 1);
}
__device__ __forceinline__ bool same_cell(const CellDir& D, const float4& a, const float4& b) {
    return cell_coord(a.x, D.inv_cell) == cell_coord(b.x, D.inv_cell) && cell_coord(a.y, D.inv_cell) == cell_coord(b.y, D.inv_cell) &&
           cell_coord(a.z, D.inv_cell) == cell_coord(b.z, D.inv_cell);
}

// (re)list, three passes so that no thread ever waits for another:
//   1. every live slot claims its 27 cells and counts itself in each (cnt_cap holds a plain count);
//   2. every cell gets room for its count + 25 % (a multiple of 4 indices) out of the list pool, counts restart;
//   3. every live slot appends its index to its 27 lists.
__global__ void k_halo_count(MapView m, int n_leaf_used, int* counters) {
    const long long total = (long long)n_leaf_used * LEAF;
    for (long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x; slot < total; slot += (long long)gridDim.x * blockDim.x) {
        const float4 p = m.pts[slot];
        if (!slot_valid(p)) continue;
#pragma unroll 1
        for (int t = 0; t < 27; t++) {
            const unsigned e = dir_claim(m.dir, halo_key(m.dir, p, t), counters);
            if (e != 0xffffffffu) atomicAdd(&m.dir.tab[e].cnt_cap, 1u);
        }
    }
}
__global__ void k_halo_alloc(MapView m, int* counters) {
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < m.dir.cap; e += gridDim.x * blockDim.x) {
        CellEntry& E = m.dir.tab[e];
        if (E.key == 0ull) continue;
        const int cnt = (int)E.cnt_cap;
        if (cnt > HALO_MAX) { E.start = -1; E.cnt_cap = 0u; atomicAdd(&counters[C_DIR_CROWDED], 1); continue; }
        const int room = (cnt + max(8, cnt / 4) + 3) & ~3;
        const int start = atomicAdd(&counters[C_DIR_POOL], room);
        if (start + room > m.dir.lists_cap) { E.start = -1; E.cnt_cap = 0u; atomicExch(&counters[C_DIR_ERROR], 1); continue; }
        E.start = start;
        E.cnt_cap = (unsigned)room << 16;
    }
}
__global__ void k_halo_fill(MapView m, int n_leaf_used) {
    const long long total = (long long)n_leaf_used * LEAF;
    for (long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x; slot < total; slot += (long long)gridDim.x * blockDim.x) {
        const float4 p = m.pts[slot];
        if (!slot_valid(p)) continue;
#pragma unroll 1
        for (int t = 0; t < 27; t++) {
            const unsigned e = dir_find(m.dir, halo_key(m.dir, p, t));
            if (e == 0xffffffffu) continue;
            CellEntry& E = m.dir.tab[e];
            if (E.start < 0) continue;
            const unsigned old = atomicAdd(&E.cnt_cap, 1u);
            const int pos = (int)(old & 0xffffu), room = (int)(old >> 16);
            if (pos < room) m.dir.lists[E.start + pos] = (int)slot;
        }
    }
}
// incremental, two kernels (claim, then append) so that nobody waits for a list that is still being set up.
// One warp per inserted point, lane t < 27 its t-th cell.  slots[i] < 0: the point could not be placed.
__global__ void __launch_bounds__(256) k_halo_claim(MapView m, const float4* __restrict__ pts, const int* __restrict__ slots, int n, int* counters) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        if (slots[i] < 0 || lane >= 27) continue;
        bool fresh = false;
        const unsigned e = dir_claim(m.dir, halo_key(m.dir, pts[i], lane), counters, &fresh);
        if (e == 0xffffffffu || !fresh) continue;
        CellEntry& E = m.dir.tab[e];
        const int start = atomicAdd(&counters[C_DIR_POOL], HALO_NEW_CAP);
        if (start + HALO_NEW_CAP > m.dir.lists_cap) { E.start = -1; E.cnt_cap = 0u; atomicExch(&counters[C_DIR_ERROR], 1); continue; }
        E.start = start;
        E.cnt_cap = (unsigned)HALO_NEW_CAP << 16;
    }
}
__global__ void __launch_bounds__(256) k_halo_append(MapView m, const float4* __restrict__ pts, const int* __restrict__ slots, int n, int* counters,
                                                      unsigned* __restrict__ fix, int fix_cap) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const int slot = slots[i];
        if (slot < 0 || lane >= 27) continue;
        const unsigned e = dir_find(m.dir, halo_key(m.dir, pts[i], lane));
        if (e == 0xffffffffu) { atomicExch(&counters[C_DIR_ERROR], 1); continue; }
        CellEntry& E = m.dir.tab[e];
        if (E.start < 0) continue;                                   // already over-full
        const unsigned old = atomicAdd(&E.cnt_cap, 1u);
        const int pos = (int)(old & 0xffffu), room = (int)(old >> 16);
        if (pos < room) m.dir.lists[E.start + pos] = slot;
        else {                                                       // no room: queue the cell for k_halo_fix (a new, larger list)
            atomicSub(&E.cnt_cap, 1u);
            if (atomicExch(&E.start, -1) >= 0) {
                const int at = atomicAdd(&counters[C_NFIX], 1);
                if (at < fix_cap) fix[at] = e; else atomicAdd(&counters[C_DIR_CROWDED], 1);
            }
        }
    }
}

// Lists that ran out of room are made anew, larger, by one warp each: the live points of the cell's 3x3x3 block are found through
// the BVH (box query over the block, exact membership by cell coordinate), counted, then listed.  The old list is abandoned in the
// pool (the next global re-list packs it away).  A cell whose block holds more than HALO_MAX points stays with the BVH walk.
struct HaloGather {
    const MapView& m; int lane; int cx, cy, cz; int* out; int cnt = 0;
    __device__ HaloGather(const MapView& m_, int lane_, int cx_, int cy_, int cz_, int* out_) : m(m_), lane(lane_), cx(cx_), cy(cy_), cz(cz_), out(out_) {}
    __device__ __forceinline__ void leaf(int l) {
        const float inv = m.dir.inv_cell;
        while (l >= 0) {
            const int slot = l * LEAF + lane;
            const float4 p = m.pts[slot];
            const bool in = slot_valid(p) && abs(cell_coord(p.x, inv) - cx) <= 1 && abs(cell_coord(p.y, inv) - cy) <= 1 && abs(cell_coord(p.z, inv) - cz) <= 1;
            const unsigned mask = __ballot_sync(FULL, in);
            if (out && in) { const int at = cnt + __popc(mask & ((1u << lane) - 1)); out[at] = slot; }
            cnt += __popc(mask);
            l = m.next[l];
        }
    }
};
__global__ void __launch_bounds__(256) k_halo_fix(MapView m, const unsigned* __restrict__ fix, int fix_cap, int* counters) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int n = min(counters[C_NFIX], fix_cap);
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        CellEntry& E = m.dir.tab[fix[i]];
        const unsigned long long key = E.key;
        const int cx = (int)((key >> 42) & 0x1fffffu) - CELL_OFF, cy = (int)((key >> 21) & 0x1fffffu) - CELL_OFF, cz = (int)(key & 0x1fffffu) - CELL_OFF;
        const float c = m.dir.cell, pad = 1e-3f * c;
        const float bmin[3] = {(cx - 1) * c - pad, (cy - 1) * c - pad, (cz - 1) * c - pad};
        const float bmax[3] = {(cx + 2) * c + pad, (cy + 2) * c + pad, (cz + 2) * c + pad};
        HaloGather count(m, lane, cx, cy, cz, nullptr);
        box_query(m, bmin, bmax, count, lane);
        const int cnt = count.cnt;
        int start = -1, room = 0;
        if (cnt <= HALO_MAX) {
            room = (cnt + max(16, cnt / 2) + 3) & ~3;
            if (lane == 0) start = atomicAdd(&counters[C_DIR_POOL], room);
            start = __shfl_sync(FULL, start, 0);
            if (start + room > m.dir.lists_cap) { start = -1; if (lane == 0) atomicExch(&counters[C_DIR_ERROR], 1); }
        }
        if (start < 0) { if (lane == 0) atomicAdd(&counters[C_DIR_CROWDED], 1); continue; }      // stays with the BVH walk
        HaloGather fill(m, lane, cx, cy, cz, m.dir.lists + start);
        box_query(m, bmin, bmax, fill, lane);
        __syncwarp();
        if (lane == 0) { E.cnt_cap = ((unsigned)room << 16) | (unsigned)fill.cnt; __threadfence(); E.start = start; }
    }
}

// Batched Nearest_Search: one lane per query (cell directory), BVH walk by the warp for what that cannot prove.
__global__ void __launch_bounds__(128) k_knn_batch(MapView m, const float4* __restrict__ q, int nq, int k,
                                                    float4* __restrict__ out_pts, float* __restrict__ out_d2,
                                                    int* __restrict__ out_cnt) {
    __shared__ WalkPool pool;
    if (threadIdx.x == 0) pool.n[0] = pool.n[1] = 0;
    __syncthreads();
    int phase = 0;
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x; base < nq; base += stride) {          // block-uniform trip count (knn_block has barriers)
        const int i = base + threadIdx.x;
        const bool active = i < nq;
        float4 qq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) qq = __ldg(&q[i]);
        TBest kb;
        knn_block(m, active, qq.x, qq.y, qq.z, kb, pool, phase);
        if (active) {
            float4 p[KNN_K];
            const int cnt = knn_fetch(m, kb, p);
#pragma unroll
            for (int j = 0; j < KNN_K; j++) {
                if (j < k) { out_pts[(size_t)i * k + j] = p[j]; out_d2[(size_t)i * k + j] = kb.d[j]; }
            }
            out_cnt[i] = min(k, cnt);
        }
        __syncthreads();
    }
}

// Delete_Point_Boxes: every slot tests itself against the boxes (half-open, ikd_Tree.cpp:796).
// A flat pass over the leaf array is bandwidth-trivial on HBM3e (16 B per slot) and needs
// no tree descent, no lazy flags and no push-down.
__global__ void k_delete_boxes(MapView m, const float* __restrict__ boxes, int nb, int n_leaf_used, int* counters,
                               float4* __restrict__ removed, int removed_cap) {
    const long long total = (long long)n_leaf_used * LEAF;
    const int lane = threadIdx.x & 31;
    int local = 0;
    for (long long base = (blockIdx.x * (long long)blockDim.x + threadIdx.x) - lane; base < total; base += (long long)gridDim.x * blockDim.x) {
        const long long slot = base + lane;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        bool hit = false;
        if (slot < total) {
            p = m.pts[slot];
            if (slot_valid(p)) for (int b = 0; b < nb && !hit; b++) hit = in_box(p, &boxes[b * 6], &boxes[b * 6 + 3]);
            if (hit) { m.pts[slot].w = __int_as_float(SLOT_TOMB); local++; }
        }
        if (removed) {                                     // history for acquire_removed_points (ikd_Tree.cpp:661-676)
            const unsigned mask = __ballot_sync(FULL, hit);
            int off = 0;
            if (lane == 0 && mask) off = atomicAdd(&counters[C_REMOVED], __popc(mask));
            off = __shfl_sync(FULL, off, 0);
            const int at = off + __popc(mask & ((1u << lane) - 1));
            if (hit && at < removed_cap) { p.w = m.payload[slot]; removed[at] = p; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL, local, o);
    if (lane == 0 && local) atomicAdd(&counters[C_DELETED], local);
}
// Add_Point_Boxes (ikd_Tree.cpp:576-603 -> Add_by_range :854-934): points deleted by Delete_Point_Boxes that lie in the boxes
// and have not been overwritten since come back (points removed by down-sampling do not, as in the reference)
__global__ void k_revive_boxes(MapView m, const float* __restrict__ boxes, int nb, int n_leaf_used, int* counters) {
    const long long total = (long long)n_leaf_used * LEAF;
    int local = 0;
    for (long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x; slot < total; slot += (long long)gridDim.x * blockDim.x) {
        const float4 p = m.pts[slot];
        if (__float_as_int(p.w) != SLOT_TOMB) continue;
        bool hit = false;
        for (int b = 0; b < nb && !hit; b++) hit = in_box(p, &boxes[b * 6], &boxes[b * 6 + 3]);
        if (hit) { m.pts[slot].w = __int_as_float(SLOT_VALID); local++; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&counters[C_REVIVED], local);
}

// Compaction of every valid point (x, y, z, intensity) -- flatten() and the input of rebuild().
__global__ void k_compact(MapView m, int n_leaf_used, float4* __restrict__ out, int* counters) {
    const long long total = (long long)n_leaf_used * LEAF;
    const int lane = threadIdx.x & 31;
    for (long long base = (blockIdx.x * (long long)blockDim.x + threadIdx.x) - lane; base < total;
         base += (long long)gridDim.x * blockDim.x) {
        const long long slot = base + lane;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        bool v = false;
        if (slot < total) { p = m.pts[slot]; v = slot_valid(p); }
        const unsigned mask = __ballot_sync(FULL, v);
        int off = 0;
        if (lane == 0 && mask) off = atomicAdd(&counters[C_COMPACT], __popc(mask));
        off = __shfl_sync(FULL, off, 0);
        if (v) {
            p.w = m.payload[slot];
            out[off + __popc(mask & ((1u << lane) - 1))] = p;
        }
    }
}

// ----------------------------------------------------------------------------- Add_Points
// Voxel of a point exactly as Add_Points computes it (ikd_Tree.cpp:491-499): float division,
// float floor, float multiply.
struct VoxelBox { float bmin[3], bmax[3], mid[3]; };
__device__ __forceinline__ void voxel_box(const float4& p, float ds, VoxelBox& vb) {
    const float c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        vb.bmin[a] = __fmul_rn(floorf(__fdiv_rn(c[a], ds)), ds);
        vb.bmax[a] = __fadd_rn(vb.bmin[a], ds);
        // mid = min + (max - min) / 2.0  evaluated in double, stored to float (ikd_Tree.cpp:497-499)
        vb.mid[a] = (float)((double)vb.bmin[a] + (double)__fsub_rn(vb.bmax[a], vb.bmin[a]) / 2.0);
    }
}
__device__ __forceinline__ unsigned long long voxel_key(const float4& p, float ds) {
    unsigned long long key = 0;
    const float c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float f = floorf(__fdiv_rn(c[a], ds)) + 1048576.0f;
        f = f < 0.f ? 0.f : (f > 2097151.f ? 2097151.f : f);
        key = (key << 21) | (unsigned long long)(unsigned)f;
    }
    return key;
}
__global__ void k_voxel_keys(const float4* __restrict__ pts, int n, float ds, unsigned long long* __restrict__ keys,
                             unsigned* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = voxel_key(pts[i], ds); vals[i] = (unsigned)i; }
}
__global__ void k_group_heads(const unsigned long long* __restrict__ keys, int n, int* __restrict__ group_start, int* counters) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n && (r == 0 || keys[r] != keys[r - 1])) group_start[atomicAdd(&counters[C_GROUPS], 1)] = r;
}

struct BoxScan {       // pass 1: count the valid points inside the voxel and find the one closest to its centre
    const MapView& m; const VoxelBox& vb; int lane;
    int count = 0; float best_d = INFINITY; int best_slot = -1;
    __device__ BoxScan(const MapView& m_, const VoxelBox& vb_, int lane_) : m(m_), vb(vb_), lane(lane_) {}
    __device__ __forceinline__ void leaf(int l) {
        while (l >= 0) {
            const int slot = l * LEAF + lane;
            const float4 p = m.pts[slot];
            const bool in = slot_valid(p) && in_box(p, vb.bmin, vb.bmax);
            const unsigned mask = __ballot_sync(FULL, in);
            if (mask) {
                count += __popc(mask);
                const float d = in ? sq_dist3(p.x, p.y, p.z, vb.mid[0], vb.mid[1], vb.mid[2]) : INFINITY;
                const unsigned key = in ? __float_as_uint(d) : 0xffffffffu;
                const unsigned best = __reduce_min_sync(FULL, key);
                if (__uint_as_float(best) < best_d) {       // strict: the first one found wins ties (ikd_Tree.cpp:508)
                    best_d = __uint_as_float(best);
                    best_slot = l * LEAF + (__ffs(__ballot_sync(FULL, key == best)) - 1);
                }
            }
            l = m.next[l];
        }
    }
};
struct BoxKill {       // pass 2: invalidate every valid point inside the voxel except `keep`
    const MapView& m; const VoxelBox& vb; int lane; int keep;
    __device__ BoxKill(const MapView& m_, const VoxelBox& vb_, int lane_, int keep_) : m(m_), vb(vb_), lane(lane_), keep(keep_) {}
    __device__ __forceinline__ void leaf(int l) {
        while (l >= 0) {
            const int slot = l * LEAF + lane;
            const float4 p = m.pts[slot];
            if (slot_valid(p) && in_box(p, vb.bmin, vb.bmax) && slot != keep) m.pts[slot].w = __int_as_float(SLOT_TOMB_DS);
            l = m.next[l];
        }
    }
};

__device__ __forceinline__ bool same_point(const float4& a, const float4& b) {   // ikd_Tree.cpp:1676-1680, EPSS = 1e-6
    return fabs((double)a.x - (double)b.x) < 1e-6 && fabs((double)a.y - (double)b.y) < 1e-6 && fabs((double)a.z - (double)b.z) < 1e-6;
}

// One warp per touched voxel.  Reproduces the sequential per-point semantics of
// Add_Points(downsample_on = true) (ikd_Tree.cpp:489-521) for all batch points that fall
// into the voxel, in their original order (the sort is stable).
__global__ void __launch_bounds__(256) k_downsample_resolve(MapView m, const float4* __restrict__ batch,
                                                             const unsigned long long* __restrict__ keys,
                                                             const unsigned* __restrict__ vals, int n,
                                                             const int* __restrict__ group_start, float ds,
                                                             float4* __restrict__ insert_list, int* counters) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int n_groups = counters[C_GROUPS];
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
        const int r0 = group_start[g];
        const unsigned long long key = keys[r0];
        const float4 first = batch[vals[r0]];
        VoxelBox vb;
        voxel_box(first, ds, vb);
        BoxScan scan(m, vb, lane);
        box_query(m, vb.bmin, vb.bmax, scan, lane);
        // ---- sequential resolution (warp-uniform scalar work)
        int cur_count = scan.count;
        float cur_d = scan.best_d;
        int cur_slot = scan.best_slot;        // >= 0: existing map point; -1: none / a batch point
        int cur_batch = -1;                   // index into batch[] if the current best is a new point
        float4 cur_pt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cur_slot >= 0) cur_pt = m.pts[cur_slot];
        bool modified = false;
        int added = 0;
        for (int r = r0; r < n && keys[r] == key; r++) {
            const int bi = (int)vals[r];
            const float4 p = batch[bi];
            const float d_new = sq_dist3(p.x, p.y, p.z, vb.mid[0], vb.mid[1], vb.mid[2]);
            const bool cur_wins = cur_count > 0 && cur_d < d_new;       // tmp_dist < min_dist
            const bool same = !cur_wins || same_point(p, cur_pt);
            if (cur_count > 1 || same) {
                modified = true;
                added++;
                if (!cur_wins) { cur_d = d_new; cur_slot = -1; cur_batch = bi; cur_pt = p; }
                cur_count = 1;
            }
        }
        if (modified) {
            const int keep = cur_slot;            // existing survivor stays in place (== delete + re-add of the same point)
            BoxKill kill(m, vb, lane, keep);
            box_query(m, vb.bmin, vb.bmax, kill, lane);
            if (lane == 0) {
                const int killed = scan.count - (keep >= 0 ? 1 : 0);
                if (killed) atomicAdd(&counters[C_TOMB], killed);
                if (keep < 0) insert_list[atomicAdd(&counters[C_NINSERT], 1)] = batch[cur_batch];
                atomicAdd(&counters[C_ADDED], added);
            }
        }
    }
}

// One warp per new point: descend towards the nearest child box to the home leaf, claim a free slot there or
// in its overflow chain (allocating a chain leaf from the pool if needed), publish the point
// and grow the AABBs on the root path with float atomics ("partial refit").
__global__ void __launch_bounds__(256) k_insert(MapView m, const float4* __restrict__ pts, int n, int* counters, int* __restrict__ slot_out) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const float4 p = pts[i];
        int node = 0;
        for (int k = m.n_levels - 1; k >= 0; k--) {
            const bool root = (k == m.n_levels - 1);
            unsigned key = 0xffffffffu;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                if (half == 1 && !root) break;
                const int c = lane + 32 * half;
                const int e = root ? c : node * FAN + c;
                if (e < m.count[k]) {
                    const float4 lo = m.ebox[k][2 * e], hi = m.ebox[k][2 * e + 1];
                    // empty entities have inverted boxes (distance +inf): they are the last resort
                    const float d = box_dist3(p.x, p.y, p.z, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z);
                    key = min(key, (min(__float_as_uint(d), 0x7f800000u) & ~63u) | (unsigned)c);
                }
            }
            const unsigned best = __reduce_min_sync(FULL, key);
            node = (root ? 0 : node * FAN) + (int)(best & 63u);
        }
        const int home = node;
        int leaf = home;
        bool placed = false;
        while (!placed) {
            const float4 cur = m.pts[leaf * LEAF + lane];
            const int w = __float_as_int(cur.w);
            unsigned freem = __ballot_sync(FULL, w == SLOT_FREE || w == SLOT_TOMB || w == SLOT_TOMB_DS);      // never used, or a deleted point's
            while (freem && !placed) {
                const int s = __ffs(freem) - 1;
                const int seen = __shfl_sync(FULL, w, s);
                int old = 0;
                if (lane == 0) old = atomicCAS((int*)&m.pts[leaf * LEAF + s].w, seen, SLOT_BUSY);
                old = __shfl_sync(FULL, old, 0);
                if (old == seen) {
                    if (lane == 0) {
                        m.payload[leaf * LEAF + s] = p.w;
                        m.pts[leaf * LEAF + s] = make_float4(p.x, p.y, p.z, __int_as_float(SLOT_VALID));
                        slot_out[i] = leaf * LEAF + s;
                    }
                    placed = true;
                } else {
                    freem &= ~(1u << s);
                }
            }
            if (placed) break;
            int nxt = 0;
            if (lane == 0) {
                nxt = atomicAdd(&m.next[leaf], 0);
                if (nxt < 0) {
                    int fresh = atomicAdd(&counters[C_LEAF_USED], 1);
                    if (fresh >= m.leaf_cap) { atomicExch(&counters[C_ERROR], 1); nxt = -2; }
                    else {
                        int old = atomicCAS(&m.next[leaf], -1, fresh);
                        nxt = (old == -1) ? fresh : old;       // lost the race: follow the winner (the fresh leaf stays unused)
                    }
                }
            }
            nxt = __shfl_sync(FULL, nxt, 0);
            if (nxt < 0) break;                                    // pool exhausted (host guarantees this cannot happen)
            leaf = nxt;
        }
        if (!placed) { if (lane == 0) slot_out[i] = -1; continue; }
        // grow the boxes of the home leaf and of its ancestors
        int e = home;
        const float c3[3] = {p.x, p.y, p.z};
        for (int k = 0; k < m.n_levels; k++) {
            if (lane < 3) atomic_min_float(&((float*)&m.ebox[k][2 * e])[lane], c3[lane]);
            else if (lane < 6) atomic_max_float(&((float*)&m.ebox[k][2 * e + 1])[lane - 3], c3[lane - 3]);
            e /= FAN;
        }
    }
}

// ============================================================================= host
static inline int blocks_for(long long threads, int block, int cap = 148 * 16) {
    long long b = (threads + block - 1) / block;
    return (int)std::max<long long>(1, std::min<long long>(b, cap));
}

Map::Map(int device, float downsample_size) : device_(device), downsample_(downsample_size) { memset(&v_, 0, sizeof(v_)); }

Map::~Map() {
    cudaSetDevice(device_);
    pts_.release(); payload_.release(); next_.release(); counters_.release(); dir_tab_.release(); dir_lists_.release(); removed_.release(); ins_slots_.release(); dir_fix_.release();
    for (int k = 0; k < MAX_LEVELS; k++) ebox_[k].release();
    segid_.release(); segtab_[0].release(); segtab_[1].release(); bbox_.release();
    src_.release(); keys_in_.release(); keys_out_.release(); vals_in_.release(); vals_out_.release();
    cub_tmp_.release(); scratch_.release(); scratch2_.release(); scratch3_.release();
    if (h_counters_) cudaFreeHost(h_counters_);
    if (stream_) cudaStreamDestroy(stream_);
}

int Map::init() {
    FL_CUDA(cudaSetDevice(device_));
    FL_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    FL_CHECK(counters_.reserve(sizeof(int) * C_COUNT));
    FL_CUDA(cudaMemsetAsync(counters_.ptr, 0, sizeof(int) * C_COUNT, stream_));
    FL_CUDA(cudaMallocHost(&h_counters_, sizeof(int) * C_COUNT));
    memset(h_counters_, 0, sizeof(int) * C_COUNT);
    if (const char* e = getenv("FASTLIO_B200_NO_CELLDIR")) dir_enabled_ = !(e[0] == '1');      // A/B: BVH walk only
    if (const char* e = getenv("FASTLIO_B200_CELL")) cell_override_ = (float)atof(e);           // A/B: cell edge in metres
    // an empty map: one empty leaf under a one-level root, so that every kernel is well defined
    FL_CHECK(src_.reserve(sizeof(float4)));
    return build_from_sorted(src_.as<float4>(), 0);
}

int Map::ensure_capacity(int n_points) {
    const int n_main = std::max(1, (n_points + fill_ - 1) / fill_);
    // overflow pool: half the main leaves, never less than 64k leaves (32 MB) -- enough for
    // any single Add_Points batch to take one fresh leaf per point in the worst case
    const long long want = (long long)n_main + std::max<long long>(n_main / 2, min_pool_);
    if (want > 0x7fffffff / LEAF) { set_last_error("map too large: %d points", n_points); return FL_ERR_CAPACITY; }
    if (want > v_.leaf_cap || !pts_.ptr) {
        FL_CHECK(pts_.reserve(sizeof(float4) * LEAF * (size_t)want));
        FL_CHECK(payload_.reserve(sizeof(float) * LEAF * (size_t)want));
        FL_CHECK(next_.reserve(sizeof(int) * (size_t)want));
        v_.leaf_cap = (int)std::min<size_t>({pts_.bytes / (sizeof(float4) * LEAF), payload_.bytes / (sizeof(float) * LEAF),
                                              next_.bytes / sizeof(int)});
        v_.pts = pts_.as<float4>(); v_.payload = payload_.as<float>(); v_.next = next_.as<int>();
    }
    // level geometry
    v_.n_main = n_main;
    v_.count[0] = n_main;
    int k = 0;
    while (true) {
        const int c = v_.count[k];
        FL_CHECK(ebox_[k].reserve(sizeof(float4) * 2 * (size_t)c));
        v_.ebox[k] = ebox_[k].as<float4>();
        k++;
        if (c <= ROOT_FAN) { v_.count[k] = 1; break; }          // the root owns up to 64 entities of level k-1
        v_.count[k] = (c + FAN - 1) / FAN;
        if (k >= MAX_LEVELS) { set_last_error("too many tree levels"); return FL_ERR_CAPACITY; }
    }
    v_.n_levels = k;
    return FL_OK;
}

int Map::refit() {
    FL_CUDA(cudaSetDevice(device_));
    k_refit_leaves<<<blocks_for((long long)v_.n_main * 32, 256), 256, 0, stream_>>>(v_);
    for (int k = 1; k < v_.n_levels; k++)
        k_refit_level<<<blocks_for((long long)v_.count[k] * 32, 256), 256, 0, stream_>>>(v_, k);
    FL_CUDA(cudaGetLastError());
    return FL_OK;
}

static int kd_depth(int span, long long pmax) {
    if (span <= 1) return 0;
    const int m = split_leaf(0, span, pmax);
    return 1 + std::max(kd_depth(m, pmax), kd_depth(span - m, pmax));
}

// d_src: n points (x, y, z, intensity) on the device, any order.
int Map::build_from_sorted(const float4* d_src, int n) {
    FL_CUDA(cudaSetDevice(device_));
    FL_CHECK(ensure_capacity(n));
    k_clear_leaves<<<blocks_for((long long)v_.leaf_cap * LEAF, 256), 256, 0, stream_>>>(v_);
    if (n > 0) {
        const int L = v_.n_main;
        long long pmax = 1;
        for (int k = 1; k < v_.n_levels; k++) pmax *= FAN;            // leaves under one child of the root
        const int depth = kd_depth(L, pmax);
        const size_t max_seg = (size_t)1 << depth;
        FL_CHECK(keys_in_.reserve(sizeof(unsigned long long) * (size_t)n));
        FL_CHECK(keys_out_.reserve(sizeof(unsigned long long) * (size_t)n));
        FL_CHECK(vals_in_.reserve(sizeof(unsigned) * (size_t)n));
        FL_CHECK(vals_out_.reserve(sizeof(unsigned) * (size_t)n));
        FL_CHECK(segid_.reserve(sizeof(int) * (size_t)n));
        FL_CHECK(segtab_[0].reserve(sizeof(Segment) * max_seg));
        FL_CHECK(segtab_[1].reserve(sizeof(Segment) * max_seg));
        FL_CHECK(bbox_.reserve(sizeof(float) * 6 * max_seg));
        unsigned* idx = vals_out_.as<unsigned>();          // current order of the points
        unsigned* idx_tmp = vals_in_.as<unsigned>();
        int* segid = segid_.as<int>();
        const int nb = blocks_for(n, 256, 1 << 30);
        k_kd_init<<<nb, 256, 0, stream_>>>(n, idx, segid, segtab_[0].as<Segment>(), L);
        int cur = 0;
        for (int lvl = 0; lvl < depth; lvl++) {
            const int nseg = 1 << lvl;
            k_kd_bbox_init<<<blocks_for((long long)nseg * 6, 256, 1 << 30), 256, 0, stream_>>>(bbox_.as<float>(), nseg);
            k_kd_bbox<<<nb, 256, 0, stream_>>>(d_src, idx, segid, n, bbox_.as<float>());
            k_kd_keys<<<nb, 256, 0, stream_>>>(d_src, idx, segid, n, bbox_.as<float>(), keys_in_.as<unsigned long long>(), idx_tmp);
            size_t tmp = 0;
            const int end_bit = 32 + std::max(1, lvl);
            FL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys_in_.as<unsigned long long>(), keys_out_.as<unsigned long long>(),
                                                    idx_tmp, idx, n, 0, end_bit, stream_));
            FL_CHECK(cub_tmp_.reserve(tmp));
            tmp = cub_tmp_.bytes;
            FL_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp_.ptr, tmp, keys_in_.as<unsigned long long>(), keys_out_.as<unsigned long long>(),
                                                    idx_tmp, idx, n, 0, end_bit, stream_));
            k_kd_split<<<blocks_for(nseg, 256, 1 << 30), 256, 0, stream_>>>(segtab_[cur].as<Segment>(), segtab_[cur ^ 1].as<Segment>(), nseg, fill_, pmax);
            k_kd_assign<<<nb, 256, 0, stream_>>>(keys_out_.as<unsigned long long>(), n, segtab_[cur ^ 1].as<Segment>(), segid);
            cur ^= 1;
        }
        k_fill_leaves<<<nb, 256, 0, stream_>>>(v_, d_src, idx, segid, segtab_[cur].as<Segment>(), n);
    }
    FL_CUDA(cudaGetLastError());
    h_counters_[C_LEAF_USED] = v_.n_main;
    FL_CUDA(cudaMemcpyAsync(&counters_.as<int>()[C_LEAF_USED], &h_counters_[C_LEAF_USED], sizeof(int), cudaMemcpyHostToDevice, stream_));
    FL_CHECK(refit());
    FL_CUDA(cudaStreamSynchronize(stream_));
    n_valid_ = n;
    n_tomb_ = 0;
    built_ = true;
    return build_directory();
}

// (Re)list the cell directory over the live slots.  Pass 1 counts into a table sized from the point count (re-sized when the
// cells turn out to be more numerous than guessed: load factor below ~0.6); the list pool is then sized from the counts.
int Map::build_directory() {
    FL_CUDA(cudaSetDevice(device_));
    if (!dir_enabled_) { v_.dir.cap = 0; return FL_OK; }
    const float cell = cell_override_ > 0.f ? cell_override_ : (downsample_ > 0.f ? 2.f * downsample_ : 1.f);
    const int used = h_counters_[C_LEAF_USED];
    size_t want_cap = std::max<size_t>(dir_min_cap_, std::max<size_t>(16384, (size_t)n_valid_ * 4));
    int* d_cnt = counters_.as<int>();
    const int nb = blocks_for((long long)used * LEAF, 256);
    for (int attempt = 0; attempt < 6; attempt++) {
        if (want_cap > 0xfffffff0ull / 2) { set_last_error("cell directory too large"); return FL_ERR_CAPACITY; }
        FL_CHECK(dir_tab_.reserve(sizeof(CellEntry) * want_cap));
        v_.dir.tab = dir_tab_.as<CellEntry>();
        v_.dir.cap = (unsigned)(dir_tab_.bytes / sizeof(CellEntry));
        v_.dir.cell = cell;
        v_.dir.inv_cell = 1.0f / cell;
        v_.dir.n_walked = &d_cnt[C_DIR_WALKED];
        FL_CUDA(cudaMemsetAsync(dir_tab_.ptr, 0, sizeof(CellEntry) * (size_t)v_.dir.cap, stream_));
        FL_CUDA(cudaMemsetAsync(&d_cnt[C_DIR_CELLS], 0, sizeof(int) * 4, stream_));      // CELLS, POOL, CROWDED, ERROR
        k_halo_count<<<nb, 256, 0, stream_>>>(v_, used, d_cnt);
        FL_CUDA(cudaGetLastError());
        FL_CUDA(cudaMemcpyAsync(&h_counters_[C_DIR_CELLS], &d_cnt[C_DIR_CELLS], sizeof(int) * 4, cudaMemcpyDeviceToHost, stream_));
        FL_CUDA(cudaStreamSynchronize(stream_));
        const size_t cells = (size_t)h_counters_[C_DIR_CELLS];
        if (h_counters_[C_DIR_ERROR] || cells * 10 > (size_t)v_.dir.cap * 6) { want_cap = std::max(want_cap * 2, cells * 5 / 2 + 16384); continue; }
        // lists: 27 listings per point + 25 % + 8 per cell (rounded to 4), plus room for the lists inserts will create
        const size_t pool = (size_t)n_valid_ * 27 + (size_t)n_valid_ * 27 / 4 + cells * 12 + std::max<size_t>(dir_min_pool_, std::max<size_t>((size_t)HALO_NEW_CAP * 262144, (size_t)n_valid_ * 6));
        if (pool > 0x7ffffff0ull) { set_last_error("cell directory: list pool too large"); return FL_ERR_CAPACITY; }
        FL_CHECK(dir_lists_.reserve(sizeof(int) * pool));
        v_.dir.lists = dir_lists_.as<int>();
        v_.dir.lists_cap = (int)std::min<size_t>(dir_lists_.bytes / sizeof(int), 0x7ffffff0ull);
        // the slack at the end of a list is read (and ignored) by the 16-byte loads of the search: keep it defined
        FL_CUDA(cudaMemsetAsync(dir_lists_.ptr, 0, sizeof(int) * (size_t)v_.dir.lists_cap, stream_));
        k_halo_alloc<<<blocks_for(v_.dir.cap, 256), 256, 0, stream_>>>(v_, d_cnt);
        k_halo_fill<<<nb, 256, 0, stream_>>>(v_, used);
        FL_CUDA(cudaGetLastError());
        FL_CUDA(cudaMemcpyAsync(&h_counters_[C_DIR_CELLS], &d_cnt[C_DIR_CELLS], sizeof(int) * 4, cudaMemcpyDeviceToHost, stream_));
        FL_CUDA(cudaStreamSynchronize(stream_));
        if (h_counters_[C_DIR_ERROR]) { dir_min_pool_ = std::max<size_t>(dir_min_pool_ * 2, (size_t)HALO_NEW_CAP * 262144); continue; }
        return FL_OK;
    }
    set_last_error("cell directory: could not size the table");
    return FL_ERR_CAPACITY;
}

int Map::dir_stats(int* out6) const {
    // [5]: queries answered by the BVH walk since the last call (0 when the directory is off: every query walks)
    int walked = 0;
    if (v_.dir.cap) {
        cudaSetDevice(device_);
        cudaMemcpyAsync(&walked, &counters_.as<int>()[C_DIR_WALKED], sizeof(int), cudaMemcpyDeviceToHost, stream_);
        cudaMemsetAsync(&counters_.as<int>()[C_DIR_WALKED], 0, sizeof(int), stream_);
        cudaStreamSynchronize(stream_);
    }
    out6[0] = h_counters_[C_DIR_CELLS]; out6[1] = h_counters_[C_DIR_POOL]; out6[2] = h_counters_[C_DIR_CROWDED];
    out6[3] = (int)v_.dir.cap; out6[4] = n_dir_rebuilds_; out6[5] = v_.dir.cap ? walked : -1;
    return FL_OK;
}

int Map::build_device(const float4* d_pts_xyzi, int n) {
    if (n < 0) { set_last_error("build: n < 0"); return FL_ERR_ARG; }
    return build_from_sorted(d_pts_xyzi, n);
}

int Map::build(const float* pts_xyzi, int n) {
    if (n < 0 || (n > 0 && !pts_xyzi)) { set_last_error("build: bad arguments"); return FL_ERR_ARG; }
    FL_CUDA(cudaSetDevice(device_));
    FL_CHECK(src_.reserve(sizeof(float4) * (size_t)std::max(n, 1)));
    if (n > 0) FL_CUDA(cudaMemcpyAsync(src_.ptr, pts_xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, stream_));
    return build_from_sorted(src_.as<float4>(), n);
}

int Map::knn(const float* q_xyzi, int nq, int k, float* out_pts, float* out_d2, int* out_cnt) {
    if (nq < 0 || k < 1 || k > KNN_K) { set_last_error("knn: k must be in [1, %d]", KNN_K); return FL_ERR_ARG; }
    if (nq == 0) return FL_OK;
    FL_CUDA(cudaSetDevice(device_));
    const size_t qb = sizeof(float4) * (size_t)nq, pb = sizeof(float4) * (size_t)nq * k, db = sizeof(float) * (size_t)nq * k, cb = sizeof(int) * (size_t)nq;
    FL_CHECK(scratch_.reserve(qb + pb + db + cb));
    char* base = scratch_.as<char>();
    float4* d_q = (float4*)base; float4* d_p = (float4*)(base + qb); float* d_d = (float*)(base + qb + pb); int* d_c = (int*)(base + qb + pb + db);
    FL_CUDA(cudaMemcpyAsync(d_q, q_xyzi, qb, cudaMemcpyHostToDevice, stream_));
    k_knn_batch<<<blocks_for(nq, 128, 1 << 20), 128, 0, stream_>>>(v_, d_q, nq, k, d_p, d_d, d_c);
    FL_CUDA(cudaGetLastError());
    FL_CUDA(cudaMemcpyAsync(out_pts, d_p, pb, cudaMemcpyDeviceToHost, stream_));
    FL_CUDA(cudaMemcpyAsync(out_d2, d_d, db, cudaMemcpyDeviceToHost, stream_));
    FL_CUDA(cudaMemcpyAsync(out_cnt, d_c, cb, cudaMemcpyDeviceToHost, stream_));
    FL_CUDA(cudaStreamSynchronize(stream_));
    return FL_OK;
}

int Map::overflow_leaves() const { return h_counters_ ? h_counters_[C_LEAF_USED] - v_.n_main : 0; }

int Map::delete_boxes(const float* boxes6, int nb, int* deleted) {
    if (deleted) *deleted = 0;
    if (nb < 0 || (nb > 0 && !boxes6)) { set_last_error("delete_boxes: bad arguments"); return FL_ERR_ARG; }
    if (nb == 0) return FL_OK;
    FL_CUDA(cudaSetDevice(device_));
    FL_CHECK(scratch_.reserve(sizeof(float) * 6 * (size_t)nb));
    FL_CUDA(cudaMemcpyAsync(scratch_.ptr, boxes6, sizeof(float) * 6 * (size_t)nb, cudaMemcpyHostToDevice, stream_));
    FL_CUDA(cudaMemsetAsync(&counters_.as<int>()[C_DELETED], 0, sizeof(int), stream_));
    const int used = h_counters_[C_LEAF_USED];
    if (record_removed_) {      // room for the worst case on top of what is already recorded
        const size_t want = (size_t)n_removed_ + (size_t)n_valid_;
        if (want * sizeof(float4) > removed_.bytes) {
            DeviceBuffer bigger;
            FL_CHECK(bigger.reserve(sizeof(float4) * (want + want / 2)));
            if (n_removed_) FL_CUDA(cudaMemcpyAsync(bigger.ptr, removed_.ptr, sizeof(float4) * (size_t)n_removed_, cudaMemcpyDeviceToDevice, stream_));
            FL_CUDA(cudaStreamSynchronize(stream_));
            removed_.release();
            removed_ = bigger;
        }
    }
    k_delete_boxes<<<blocks_for((long long)used * LEAF, 256), 256, 0, stream_>>>(v_, scratch_.as<float>(), nb, used, counters_.as<int>(),
                                                                                   record_removed_ ? removed_.as<float4>() : nullptr,
                                                                                   (int)std::min<size_t>(removed_.bytes / sizeof(float4), 0x7fffffff));
    FL_CUDA(cudaGetLastError());
    FL_CUDA(cudaMemcpyAsync(&h_counters_[C_DELETED], &counters_.as<int>()[C_DELETED], sizeof(int), cudaMemcpyDeviceToHost, stream_));
    FL_CUDA(cudaStreamSynchronize(stream_));
    const int d = h_counters_[C_DELETED];
    if (record_removed_) n_removed_ += d;
    if (d > 0) {
        n_valid_ -= d; n_tomb_ += d;
        FL_CHECK(refit());          // tighten every AABB ("box-delete by refit")
        FL_CHECK(maybe_rebuild());
    }
    if (deleted) *deleted = d;
    return FL_OK;
}

// KD_TREE::acquire_removed_points (ikd_Tree.cpp:661-676): the points Delete_Point_Boxes removed since the last call.  Recording
// starts with the first call (the reference's caller asks before every box delete, laserMapping.cpp:273-275).
int Map::acquire_removed(float* out_xyzi, int cap, int* n_out) {
    FL_CUDA(cudaSetDevice(device_));
    const int n = n_removed_;
    if (n_out) *n_out = n;
    if (!record_removed_) {
        record_removed_ = true;
        FL_CUDA(cudaMemsetAsync(&counters_.as<int>()[C_REMOVED], 0, sizeof(int), stream_));
        return FL_OK;
    }
    if (n > 0 && out_xyzi && cap > 0)
        FL_CUDA(cudaMemcpyAsync(out_xyzi, removed_.ptr, sizeof(float4) * (size_t)std::min(n, cap), cudaMemcpyDeviceToHost, stream_));
    FL_CUDA(cudaMemsetAsync(&counters_.as<int>()[C_REMOVED], 0, sizeof(int), stream_));
    FL_CUDA(cudaStreamSynchronize(stream_));
    n_removed_ = 0;
    return FL_OK;
}

// KD_TREE::Add_Point_Boxes (ikd_Tree.cpp:576-603)
int Map::add_boxes(const float* boxes6, int nb, int* revived) {
    if (revived) *revived = 0;
    if (nb < 0 || (nb > 0 && !boxes6)) { set_last_error("add_boxes: bad arguments"); return FL_ERR_ARG; }
    if (nb == 0) return FL_OK;
    FL_CUDA(cudaSetDevice(device_));
    FL_CHECK(scratch_.reserve(sizeof(float) * 6 * (size_t)nb));
    FL_CUDA(cudaMemcpyAsync(scratch_.ptr, boxes6, sizeof(float) * 6 * (size_t)nb, cudaMemcpyHostToDevice, stream_));
    FL_CUDA(cudaMemsetAsync(&counters_.as<int>()[C_REVIVED], 0, sizeof(int), stream_));
    const int used = h_counters_[C_LEAF_USED];
    k_revive_boxes<<<blocks_for((long long)used * LEAF, 256), 256, 0, stream_>>>(v_, scratch_.as<float>(), nb, used, counters_.as<int>());
    FL_CUDA(cudaGetLastError());
    FL_CUDA(cudaMemcpyAsync(&h_counters_[C_REVIVED], &counters_.as<int>()[C_REVIVED], sizeof(int), cudaMemcpyDeviceToHost, stream_));
    FL_CUDA(cudaStreamSynchronize(stream_));
    const int r = h_counters_[C_REVIVED];
    if (r > 0) {
        n_valid_ += r; n_tomb_ = std::max(0, n_tomb_ - r);
        FL_CHECK(refit());          // the boxes were tightened when the points went away
    }
    if (revived) *revived = r;
    return FL_OK;
}

int Map::flatten(float* out_xyzi, int cap, int* n_out) {
    FL_CUDA(cudaSetDevice(device_));
    FL_CHECK(src_.reserve(sizeof(float4) * (size_t)std::max(1, n_valid_)));
    FL_CUDA(cudaMemsetAsync(&counters_.as<int>()[C_COMPACT], 0, sizeof(int), stream_));
    const int used = h_counters_[C_LEAF_USED];
    k_compact<<<blocks_for((long long)used * LEAF, 256), 256, 0, stream_>>>(v_, used, src_.as<float4>(), counters_.as<int>());
    FL_CUDA(cudaGetLastError());
    FL_CUDA(cudaMemcpyAsync(&h_counters_[C_COMPACT], &counters_.as<int>()[C_COMPACT], sizeof(int), cudaMemcpyDeviceToHost, stream_));
    FL_CUDA(cudaStreamSynchronize(stream_));
    const int n = h_counters_[C_COMPACT];
    if (n != n_valid_) { set_last_error("flatten: %d valid points on device, host expected %d", n, n_valid_); return FL_ERR_STATE; }
    if (n_out) *n_out = n;
    if (out_xyzi && cap > 0) {
        FL_CUDA(cudaMemcpyAsync(out_xyzi, src_.ptr, sizeof(float4) * (size_t)std::min(n, cap), cudaMemcpyDeviceToHost, stream_));
        FL_CUDA(cudaStreamSynchronize(stream_));
    }
    return FL_OK;
}

int Map::rebuild() {
    FL_CUDA(cudaSetDevice(device_));
    FL_CHECK(src_.reserve(sizeof(float4) * (size_t)std::max(1, n_valid_)));
    FL_CUDA(cudaMemsetAsync(&counters_.as<int>()[C_COMPACT], 0, sizeof(int), stream_));
    const int used = h_counters_[C_LEAF_USED];
    k_compact<<<blocks_for((long long)used * LEAF, 256), 256, 0, stream_>>>(v_, used, src_.as<float4>(), counters_.as<int>());
    FL_CUDA(cudaGetLastError());
    n_rebuilds_++;
    return build_from_sorted(src_.as<float4>(), n_valid_);
}

int Map::maybe_rebuild() {
    const int overflow = h_counters_[C_LEAF_USED] - v_.n_main;
    const bool too_chained = overflow > std::max(64, (int)(rebuild_overflow_frac_ * v_.n_main));
    const bool too_sparse = n_tomb_ > 1024 && n_tomb_ > n_valid_;       // ikd-Tree's delete criterion (alpha_del = 0.5)
    if (too_chained || too_sparse) return rebuild();
    return FL_OK;
}

int Map::insert_device(const float4* d_pts, int n) {
    if (n <= 0) return FL_OK;
    // worst case one fresh overflow leaf per point: guarantee the pool can take the batch
    if ((long long)h_counters_[C_LEAF_USED] + n > v_.leaf_cap) {
        min_pool_ = std::max(min_pool_, n + 1024);
        FL_CHECK(rebuild());            // re-packs the leaves and re-sizes the pool
    }
    if (v_.dir.cap) {
        // the insert kernels cope with a full table / an exhausted list pool (they flag it, the directory is re-listed below);
        // re-list beforehand only when the table could get so full that probing degenerates
        const size_t cells = (size_t)h_counters_[C_DIR_CELLS] + (size_t)n * 27;
        if (cells * 10 > (size_t)v_.dir.cap * 9) {
            dir_min_cap_ = std::max(dir_min_cap_, cells * 2 + 16384);
            n_dir_rebuilds_++;
            FL_CHECK(build_directory());
        }
    }
    FL_CHECK(ins_slots_.reserve(sizeof(int) * (size_t)n));
    k_insert<<<blocks_for((long long)n * 32, 256), 256, 0, stream_>>>(v_, d_pts, n, counters_.as<int>(), ins_slots_.as<int>());
    if (v_.dir.cap) {
        k_halo_claim<<<blocks_for((long long)n * 32, 256), 256, 0, stream_>>>(v_, d_pts, ins_slots_.as<int>(), n, counters_.as<int>());
        FL_CHECK(dir_fix_.reserve(sizeof(unsigned) * 65536));
        FL_CUDA(cudaMemsetAsync(&counters_.as<int>()[C_NFIX], 0, sizeof(int), stream_));
        k_halo_append<<<blocks_for((long long)n * 32, 256), 256, 0, stream_>>>(v_, d_pts, ins_slots_.as<int>(), n, counters_.as<int>(),
                                                                                 dir_fix_.as<unsigned>(), 65536);
        k_halo_fix<<<296, 256, 0, stream_>>>(v_, dir_fix_.as<unsigned>(), 65536, counters_.as<int>());
    }
    FL_CUDA(cudaGetLastError());
    FL_CUDA(cudaMemcpyAsync(h_counters_, counters_.ptr, sizeof(int) * C_COUNT, cudaMemcpyDeviceToHost, stream_));
    FL_CUDA(cudaStreamSynchronize(stream_));
    if (h_counters_[C_ERROR]) { set_last_error("insert: overflow pool exhausted"); return FL_ERR_CAPACITY; }
    n_valid_ += n;
    // directory: out of room, or too many over-full cells (their queries walk the BVH) -> re-list the live slots
    if (v_.dir.cap) {
        const bool crowded = h_counters_[C_DIR_CROWDED] > std::max(64, h_counters_[C_DIR_CELLS] / 1024);
        if (h_counters_[C_DIR_ERROR]) {
            dir_min_cap_ = std::max(dir_min_cap_, (size_t)v_.dir.cap + (size_t)v_.dir.cap / 2);
            dir_min_pool_ = std::max<size_t>(dir_min_pool_ * 2, (size_t)HALO_NEW_CAP * 262144);
        }
        if (h_counters_[C_DIR_ERROR] || crowded) { n_dir_rebuilds_++; FL_CHECK(build_directory()); }
    }
    return FL_OK;
}

int Map::add_points_device(const float4* d_pts, int n, bool downsample_on, int* added) {
    if (added) *added = 0;
    if (n < 0) { set_last_error("add_points: n < 0"); return FL_ERR_ARG; }
    if (n == 0) return FL_OK;
    FL_CUDA(cudaSetDevice(device_));
    if (!downsample_on) {                                   // ikd_Tree.cpp:549-568: plain inserts, return value 0
        FL_CHECK(insert_device(d_pts, n));
        return maybe_rebuild();
    }
    FL_CHECK(keys_in_.reserve(sizeof(unsigned long long) * (size_t)n));
    FL_CHECK(keys_out_.reserve(sizeof(unsigned long long) * (size_t)n));
    FL_CHECK(vals_in_.reserve(sizeof(unsigned) * (size_t)n));
    FL_CHECK(vals_out_.reserve(sizeof(unsigned) * (size_t)n));
    FL_CHECK(scratch2_.reserve(sizeof(int) * (size_t)n));          // group starts
    FL_CHECK(scratch3_.reserve(sizeof(float4) * (size_t)n));       // insert list
    int* d_cnt = counters_.as<int>();
    FL_CUDA(cudaMemsetAsync(&d_cnt[C_ADDED], 0, sizeof(int) * 4, stream_));   // ADDED, GROUPS, NINSERT, TOMB
    k_voxel_keys<<<blocks_for(n, 256, 1 << 30), 256, 0, stream_>>>(d_pts, n, downsample_, keys_in_.as<unsigned long long>(), vals_in_.as<unsigned>());
    size_t tmp = 0;
    FL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys_in_.as<unsigned long long>(), keys_out_.as<unsigned long long>(),
                                            vals_in_.as<unsigned>(), vals_out_.as<unsigned>(), n, 0, 63, stream_));
    FL_CHECK(cub_tmp_.reserve(tmp));
    tmp = cub_tmp_.bytes;
    FL_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp_.ptr, tmp, keys_in_.as<unsigned long long>(), keys_out_.as<unsigned long long>(),
                                            vals_in_.as<unsigned>(), vals_out_.as<unsigned>(), n, 0, 63, stream_));
    k_group_heads<<<blocks_for(n, 256, 1 << 30), 256, 0, stream_>>>(keys_out_.as<unsigned long long>(), n, scratch2_.as<int>(), d_cnt);
    k_downsample_resolve<<<blocks_for((long long)n * 32, 256), 256, 0, stream_>>>(
        v_, d_pts, keys_out_.as<unsigned long long>(), vals_out_.as<unsigned>(), n, scratch2_.as<int>(), downsample_,
        scratch3_.as<float4>(), d_cnt);
    FL_CUDA(cudaGetLastError());
    FL_CUDA(cudaMemcpyAsync(h_counters_, counters_.ptr, sizeof(int) * C_COUNT, cudaMemcpyDeviceToHost, stream_));
    FL_CUDA(cudaStreamSynchronize(stream_));
    const int n_ins = h_counters_[C_NINSERT], n_tomb = h_counters_[C_TOMB];
    if (added) *added = h_counters_[C_ADDED];
    n_valid_ -= n_tomb; n_tomb_ += n_tomb;
    FL_CHECK(insert_device(scratch3_.as<float4>(), n_ins));
    // tombstoned slots are reused by later inserts; they stop counting once re-occupied
    n_tomb_ = std::max(0, n_tomb_ - n_ins);
    return maybe_rebuild();
}

int Map::add_points(const float* pts_xyzi, int n, bool downsample_on, int* added) {
    if (added) *added = 0;
    if (n < 0 || (n > 0 && !pts_xyzi)) { set_last_error("add_points: bad arguments"); return FL_ERR_ARG; }
    if (n == 0) return FL_OK;
    FL_CUDA(cudaSetDevice(device_));
    FL_CHECK(scratch_.reserve(sizeof(float4) * (size_t)n));
    FL_CUDA(cudaMemcpyAsync(scratch_.ptr, pts_xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, stream_));
    return add_points_device(scratch_.as<float4>(), n, downsample_on, added);
}

int Map::tree_range(float* box6) {
    FL_CUDA(cudaSetDevice(device_));
    // union of the top-level entity boxes
    const int k = v_.n_levels - 1;
    const int c = v_.count[k];
    std::vector<float4> h(2 * (size_t)c);
    FL_CUDA(cudaMemcpyAsync(h.data(), v_.ebox[k], sizeof(float4) * 2 * (size_t)c, cudaMemcpyDeviceToHost, stream_));
    FL_CUDA(cudaStreamSynchronize(stream_));
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < c; i++) {
        lo[0] = std::min(lo[0], h[2 * i].x); lo[1] = std::min(lo[1], h[2 * i].y); lo[2] = std::min(lo[2], h[2 * i].z);
        hi[0] = std::max(hi[0], h[2 * i + 1].x); hi[1] = std::max(hi[1], h[2 * i + 1].y); hi[2] = std::max(hi[2], h[2 * i + 1].z);
    }
    if (n_valid_ == 0) for (int a = 0; a < 3; a++) lo[a] = hi[a] = 0.f;        // memset(&range, 0, ...) ikd_Tree.cpp:114
    for (int a = 0; a < 3; a++) { box6[a] = lo[a]; box6[3 + a] = hi[a]; }
    return FL_OK;
}

}  // namespace fl
