// EXPERIMENTAL -- round-2 prototype, NOT part of libfastlio_b200.so and NOT validated on a GPU yet (written after the
// round-1 GPU budget was spent; scripts/build_experimental.py only proves that it compiles for sm_100a).
//
// "Hash grid of leaf buckets": the same 32-slot float4 buckets and overflow chains as the product's map (map.cuh), but
// the 32-ary BVH above them is replaced by a hashed directory of cubic cells.  A bucket's bounding box is implicit (the
// cell), so the k-NN walk needs no box loads: lane l < 27 owns the neighbour cell (dx, dy, dz), scores it from the
// query's offset inside its own cell, probes the directory, and the warp then visits the occupied neighbours nearest
// first with the same k-best machinery as k_search (knn_leaf).  After the 27-cell block, the result is final when the
// k-th squared distance is strictly below the squared distance to the block's faces (tests/cell_directory_model.py
// pins that rule on the CPU); otherwise shells of cells are added ring by ring.
//
// Why: DESIGN.md §6c -- on the benchmark map the 5-NN ball cuts 5.3 cells of 2 m (2.9 occupied); the BVH walk costs 508
// warp instructions per query, most of them box tests and child selection in the three levels above the leaves.
#include <cfloat>
#include <cstdarg>
#include <cstdio>

#include <cub/cub.cuh>

#include "../common.cuh"
#include "../map.cuh"

namespace fl {
void set_last_error(const char* fmt, ...) {        // the prototype library is standalone
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr);
}
}  // namespace fl

namespace flx {
using namespace fl;

constexpr unsigned long long EMPTY_KEY = ~0ull;
constexpr int CELL_BIAS = 1 << 20;                 // 21 bits per axis

struct CellDir {
    unsigned long long* keys;   // [cap] packed cell coordinates, EMPTY_KEY = free
    int* first;                 // [cap] first bucket of the cell
    int log2cap;
    float cell, inv;
    int lo[3], hi[3];           // cell bounds of the map: rings stop once they cover them
};

__host__ __device__ __forceinline__ unsigned long long pack_cell(int ix, int iy, int iz) {
    return ((unsigned long long)(unsigned)(ix + CELL_BIAS) << 42) | ((unsigned long long)(unsigned)(iy + CELL_BIAS) << 21) |
           (unsigned long long)(unsigned)(iz + CELL_BIAS);
}
__device__ __forceinline__ unsigned slot_of(unsigned long long key, int log2cap) {
    return (unsigned)((key * 0x9E3779B97F4A7C15ull) >> (64 - log2cap));
}
__device__ __forceinline__ int dir_find(const CellDir& d, unsigned long long key) {
    const unsigned mask = (1u << d.log2cap) - 1u;
    unsigned s = slot_of(key, d.log2cap);
    while (true) {
        const unsigned long long k = __ldg(&d.keys[s]);
        if (k == key) return __ldg(&d.first[s]);
        if (k == EMPTY_KEY) return -1;
        s = (s + 1) & mask;
    }
}

// ------------------------------------------------------------------------------------------------ build
__global__ void k_cell_keys(const float4* __restrict__ pts, int n, float inv, unsigned long long* keys, int* idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    keys[i] = pack_cell((int)floorf(__fmul_rn(p.x, inv)), (int)floorf(__fmul_rn(p.y, inv)), (int)floorf(__fmul_rn(p.z, inv)));
    idx[i] = i;
}
// heads[i] = number of buckets the run starting at i needs (0 if i is not the first point of its cell)
__global__ void k_cell_runs(const unsigned long long* __restrict__ keys, int n, int* __restrict__ need) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i > 0 && keys[i] == keys[i - 1]) { need[i] = 0; return; }
    int j = i + 1;
    while (j < n && keys[j] == keys[i]) j++;
    need[i] = (j - i + LEAF - 1) / LEAF;
}
__global__ void k_cell_fill(const float4* __restrict__ src, const unsigned long long* __restrict__ keys, const int* __restrict__ idx,
                            const int* __restrict__ need, const int* __restrict__ base, int n, MapView m, CellDir d, int* n_buckets) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == n - 1) *n_buckets = base[i] + need[i];
    if (need[i] == 0) return;
    const unsigned long long key = keys[i];
    int j = i, b = base[i];
    while (j < n && keys[j] == key) {                 // one thread per cell: cells hold a handful of points
        const int take = min(LEAF, n - j);
        int filled = 0;
        for (int s = 0; s < LEAF; s++) {
            float4 slot = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
            float pay = 0.f;
            if (s < take && keys[j + s] == key) {
                const float4 p = src[idx[j + s]];
                slot = make_float4(p.x, p.y, p.z, __int_as_float(1));
                pay = p.w;
                filled++;
            }
            m.pts[(size_t)b * LEAF + s] = slot;
            m.payload[(size_t)b * LEAF + s] = pay;
        }
        j += filled;
        const bool more = j < n && keys[j] == key;
        m.next[b] = more ? b + 1 : -1;
        b++;
    }
    // publish the cell in the directory (linear probing, CAS on the key)
    const unsigned mask = (1u << d.log2cap) - 1u;
    unsigned s = slot_of(key, d.log2cap);
    while (true) {
        const unsigned long long prev = atomicCAS(&d.keys[s], EMPTY_KEY, key);
        if (prev == EMPTY_KEY || prev == key) { d.first[s] = base[i]; break; }
        s = (s + 1) & mask;
    }
}

// ------------------------------------------------------------------------------------------------ k-NN
// squared distance from the query to the neighbour cell (dx, dy, dz in -1..1) of its own cell: per axis the gap to the
// shared face, 0 on the axes where the neighbour is the own slab.  `lo` = q - c0*cell, `hi` = (c0+1)*cell - q.
__device__ __forceinline__ float cell_gap2(int d, float lo, float hi) {
    const float g = d < 0 ? lo : (d > 0 ? hi : 0.f);
    return g * g;
}

__global__ void __launch_bounds__(256) k_cell_knn(MapView m, CellDir d, const float4* __restrict__ q, int nq, int k,
                                                  float4* __restrict__ out_pts, float* __restrict__ out_d2, int* __restrict__ out_cnt,
                                                  int* __restrict__ ring_stats) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nq; i += warps) {
        const float4 qq = __ldg(&q[i]);
        const int cx = (int)floorf(__fmul_rn(qq.x, d.inv)), cy = (int)floorf(__fmul_rn(qq.y, d.inv)), cz = (int)floorf(__fmul_rn(qq.z, d.inv));
        // gaps to the faces of the own cell, shrunk by a relative 2^-20 so that rounding in c*cell can never prune a
        // cell that holds a closer point (exact anyway when the cell size is a power of two)
        const float sh = 1.0f - 9.5367431640625e-07f;
        const float lox = fmaxf(0.f, (qq.x - cx * d.cell) * sh), hix = fmaxf(0.f, ((cx + 1) * d.cell - qq.x) * sh);
        const float loy = fmaxf(0.f, (qq.y - cy * d.cell) * sh), hiy = fmaxf(0.f, ((cy + 1) * d.cell - qq.y) * sh);
        const float loz = fmaxf(0.f, (qq.z - cz * d.cell) * sh), hiz = fmaxf(0.f, ((cz + 1) * d.cell - qq.z) * sh);
        // beyond this ring every cell of the map has been seen
        const int max_ring = max(max(max(cx - d.lo[0], d.hi[0] - cx), max(cy - d.lo[1], d.hi[1] - cy)), max(max(cz - d.lo[2], d.hi[2] - cz), 1));
        KBest kb;
        kb.init();
        // ---- ring 0 and 1: one neighbour cell per lane
        unsigned key = 0xffffffffu;
        int bucket = -1;
        if (lane < 27) {
            const int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
            bucket = dir_find(d, pack_cell(cx + dx, cy + dy, cz + dz));
            if (bucket >= 0) {
                const float g2 = cell_gap2(dx, lox, hix) + cell_gap2(dy, loy, hiy) + cell_gap2(dz, loz, hiz);
                key = (__float_as_uint(g2) & ~31u) | (unsigned)lane;
            }
        }
#pragma unroll 1
        while (true) {
            const unsigned best = __reduce_min_sync(FULL, key);
            if ((best & ~31u) >= __float_as_uint(kb.w)) break;       // also ends the loop when only markers are left
            const int c = best & 31;
            const int leaf = __shfl_sync(FULL, bucket, c);
            if (lane == c) key = 0xffffffffu;
            knn_leaf(m, leaf, qq.x, qq.y, qq.z, kb, lane);
        }
        // ---- are the k best final?  Everything outside the (2r+1)^3 block is at least `g` away.
        int r = 1;
        while (true) {
            const float g = fminf(fminf(fminf(lox, hix), fminf(loy, hiy)), fminf(loz, hiz)) + (float)r * d.cell * sh;
            if ((kb.n == KNN_K || kb.w < INFINITY) && kb.w < g * g) break;
            if (r >= max_ring) break;
            r++;
            // shell of ring r: (2r+1)^3 - (2r-1)^3 cells, 32 per step; the rare slow path, kept simple
            const int side = 2 * r + 1, total = side * side * side;
            for (int base = 0; base < total; base += 32) {
                const int t = base + lane;
                int leaf = -1;
                if (t < total) {
                    const int dx = t % side - r, dy = (t / side) % side - r, dz = t / (side * side) - r;
                    if (max(abs(dx), max(abs(dy), abs(dz))) == r) leaf = dir_find(d, pack_cell(cx + dx, cy + dy, cz + dz));
                }
                unsigned todo = __ballot_sync(FULL, leaf >= 0);
                while (todo) {
                    const int c = __ffs(todo) - 1;
                    todo &= todo - 1;
                    knn_leaf(m, __shfl_sync(FULL, leaf, c), qq.x, qq.y, qq.z, kb, lane);
                }
            }
        }
        if (ring_stats && lane == 0) atomicAdd(&ring_stats[min(r, 7)], 1);
        const int cnt = min(k, __popc(__ballot_sync(FULL, lane < KNN_K && kb.idx >= 0)));
        if (lane < k) {
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kb.idx >= 0) { p = m.pts[kb.idx]; p.w = m.payload[kb.idx]; }
            out_pts[(size_t)i * k + lane] = p;
            out_d2[(size_t)i * k + lane] = kb.d;
        }
        if (lane == 0) out_cnt[i] = cnt;
    }
}

struct CellMap {
    int device = 0;
    float cell = 2.0f;
    cudaStream_t stream = nullptr;
    MapView view;
    CellDir dir;
    int n_buckets = 0, n_points = 0;
    void *pts = nullptr, *payload = nullptr, *next = nullptr, *keys = nullptr, *first = nullptr;
};

}  // namespace flx

#define FLX_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return -1; } } while (0)

extern "C" {

int flx_cellmap_create(flx::CellMap** out, int device, float cell) {
    flx::CellMap* m = new flx::CellMap();
    m->device = device; m->cell = cell;
    memset(&m->view, 0, sizeof(m->view));
    FLX_CUDA(cudaSetDevice(device));
    FLX_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    *out = m;
    return 0;
}

int flx_cellmap_build(flx::CellMap* m, const float* pts_xyzi, int n) {
    using namespace flx;
    FLX_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    const size_t N = (size_t)(n > 0 ? n : 1);
    float4* src; unsigned long long *k0, *k1; int *i0, *i1, *need, *base, *d_nb;
    FLX_CUDA(cudaMalloc(&src, sizeof(float4) * N)); FLX_CUDA(cudaMalloc(&k0, 8 * N)); FLX_CUDA(cudaMalloc(&k1, 8 * N));
    FLX_CUDA(cudaMalloc(&i0, 4 * N)); FLX_CUDA(cudaMalloc(&i1, 4 * N)); FLX_CUDA(cudaMalloc(&need, 4 * N)); FLX_CUDA(cudaMalloc(&base, 4 * N));
    FLX_CUDA(cudaMalloc(&d_nb, 4));
    FLX_CUDA(cudaMemcpyAsync(src, pts_xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
    // worst case one bucket per point; a real implementation sizes this from the run lengths
    const size_t cap_buckets = N;
    for (void* p : {m->pts, m->payload, m->next, m->keys, m->first}) if (p) cudaFree(p);
    FLX_CUDA(cudaMalloc(&m->pts, sizeof(float4) * LEAF * cap_buckets)); FLX_CUDA(cudaMalloc(&m->payload, sizeof(float) * LEAF * cap_buckets));
    FLX_CUDA(cudaMalloc(&m->next, sizeof(int) * cap_buckets));
    int log2cap = 4;
    while ((1ull << log2cap) < 2 * N) log2cap++;
    FLX_CUDA(cudaMalloc(&m->keys, 8ull << log2cap)); FLX_CUDA(cudaMalloc(&m->first, 4ull << log2cap));
    FLX_CUDA(cudaMemsetAsync(m->keys, 0xff, 8ull << log2cap, st));
    m->view.pts = (float4*)m->pts; m->view.payload = (float*)m->payload; m->view.next = (int*)m->next;
    m->dir.keys = (unsigned long long*)m->keys; m->dir.first = (int*)m->first; m->dir.log2cap = log2cap;
    m->dir.cell = m->cell; m->dir.inv = 1.0f / m->cell;
    for (int a = 0; a < 3; a++) { m->dir.lo[a] = 0; m->dir.hi[a] = 0; }
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            const int c = (int)floorf(pts_xyzi[4 * (size_t)i + a] * m->dir.inv);
            if (i == 0 || c < m->dir.lo[a]) m->dir.lo[a] = c;
            if (i == 0 || c > m->dir.hi[a]) m->dir.hi[a] = c;
        }
    m->n_points = n;
    if (n > 0) {
        const int nb = (n + 255) / 256;
        k_cell_keys<<<nb, 256, 0, st>>>(src, n, m->dir.inv, k0, i0);
        size_t tmp = 0; void* d_tmp = nullptr;
        FLX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k0, k1, i0, i1, n, 0, 63, st));
        size_t tmp2 = 0;
        FLX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, need, base, n, st));
        if (tmp2 > tmp) tmp = tmp2;
        FLX_CUDA(cudaMalloc(&d_tmp, tmp));
        FLX_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp, k0, k1, i0, i1, n, 0, 63, st));
        k_cell_runs<<<nb, 256, 0, st>>>(k1, n, need);
        FLX_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp, need, base, n, st));
        k_cell_fill<<<nb, 256, 0, st>>>(src, k1, i1, need, base, n, m->view, m->dir, d_nb);
        FLX_CUDA(cudaGetLastError());
        FLX_CUDA(cudaMemcpyAsync(&m->n_buckets, d_nb, 4, cudaMemcpyDeviceToHost, st));
        FLX_CUDA(cudaStreamSynchronize(st));
        cudaFree(d_tmp);
    }
    for (void* p : {(void*)src, (void*)k0, (void*)k1, (void*)i0, (void*)i1, (void*)need, (void*)base, (void*)d_nb}) cudaFree(p);
    return 0;
}

// returns device milliseconds of the k-NN kernel in *ms (may be NULL); ring_stats8 (may be NULL): queries settled at ring r
int flx_cellmap_knn(flx::CellMap* m, const float* q_xyzi, int nq, int k, float* out_pts, float* out_d2, int* out_cnt, float* ms, int* ring_stats8) {
    using namespace flx;
    if (k < 1 || k > KNN_K) return -2;
    FLX_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    const size_t Q = (size_t)(nq > 0 ? nq : 1);
    float4 *d_q, *d_p; float* d_d; int *d_c, *d_r;
    FLX_CUDA(cudaMalloc(&d_q, 16 * Q)); FLX_CUDA(cudaMalloc(&d_p, 16 * Q * k)); FLX_CUDA(cudaMalloc(&d_d, 4 * Q * k)); FLX_CUDA(cudaMalloc(&d_c, 4 * Q));
    FLX_CUDA(cudaMalloc(&d_r, 32));
    FLX_CUDA(cudaMemsetAsync(d_r, 0, 32, st));
    FLX_CUDA(cudaMemcpyAsync(d_q, q_xyzi, 16 * (size_t)nq, cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1;
    FLX_CUDA(cudaEventCreate(&e0)); FLX_CUDA(cudaEventCreate(&e1));
    if (nq > 0 && m->n_points > 0) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
        long long want = ((long long)nq * 32 + 255) / 256;
        int grid = (int)(want < (long long)sms * 5 ? want : (long long)sms * 5);
        FLX_CUDA(cudaEventRecord(e0, st));
        k_cell_knn<<<grid, 256, 0, st>>>(m->view, m->dir, d_q, nq, k, d_p, d_d, d_c, d_r);
        FLX_CUDA(cudaEventRecord(e1, st));
        FLX_CUDA(cudaGetLastError());
    } else {
        FLX_CUDA(cudaMemsetAsync(d_c, 0, 4 * Q, st));
        FLX_CUDA(cudaEventRecord(e0, st)); FLX_CUDA(cudaEventRecord(e1, st));
    }
    FLX_CUDA(cudaMemcpyAsync(out_pts, d_p, 16 * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
    FLX_CUDA(cudaMemcpyAsync(out_d2, d_d, 4 * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
    FLX_CUDA(cudaMemcpyAsync(out_cnt, d_c, 4 * (size_t)nq, cudaMemcpyDeviceToHost, st));
    if (ring_stats8) FLX_CUDA(cudaMemcpyAsync(ring_stats8, d_r, 32, cudaMemcpyDeviceToHost, st));
    FLX_CUDA(cudaStreamSynchronize(st));
    if (ms) { float t = 0.f; cudaEventElapsedTime(&t, e0, e1); *ms = t; }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (void* p : {(void*)d_q, (void*)d_p, (void*)d_d, (void*)d_c, (void*)d_r}) cudaFree(p);
    return 0;
}

int flx_cellmap_destroy(flx::CellMap* m) {
    if (!m) return 0;
    cudaSetDevice(m->device);
    for (void* p : {m->pts, m->payload, m->next, m->keys, m->first}) if (p) cudaFree(p);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return 0;
}

}  // extern "C"
