// Scan front end kernels (see scan.h).  Everything runs on the map's stream, so a scan goes
// raw -> de-skewed -> down-sampled -> iEKF update -> map_incremental without leaving HBM.
#include <cfloat>
#include <cmath>
#include <cstring>

#include <cub/cub.cuh>

#include "common.cuh"
#include "lie.cuh"
#include "scan.h"

namespace fl {

void set_last_error(const char* fmt, ...);

namespace {

// ------------------------------------------------------------------------------------------------ de-skew
struct EndState {
    D3 pos, offT;
    Q4 rot_inv, offR, offR_inv;
};

// Exp(ang_vel, dt) as the reference evaluates it (include/so3_math.h:37-58): Rodrigues with the
// normalised axis, I + sin(a) K + ((1 - cos a) K) K.
__device__ __forceinline__ M33 rodrigues(const D3& w, double dt) {
    const double n = norm3(w);
    if (!(n > 0.0000001)) return eye33();
    const D3 a = d3(w.x / n, w.y / n, w.z / n);
    const M33 K = hat3(a);
    const double ang = n * dt;
    double s, c;
    sincos(ang, &s, &c);
    const double c1 = 1.0 - c;
    M33 cK;
#pragma unroll
    for (int i = 0; i < 9; i++) cK.m[i] = c1 * K.m[i];
    const M33 cKK = mul33(cK, K);
    const M33 I = eye33();
    M33 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.m[i] = (I.m[i] + s * K.m[i]) + cKK.m[i];
    return r;
}

// One point, one IMU segment [head, tail] (IMU_Processing.hpp:327-341).
__device__ __forceinline__ void compensate(float4& p, double t, const double* head, const double* tail, const EndState& e) {
    const double dt = t - head[0];
    M33 R_imu;
#pragma unroll
    for (int i = 0; i < 9; i++) R_imu.m[i] = head[13 + i];
    const D3 vel = ld3(head + 7), pos = ld3(head + 10), acc = ld3(tail + 1), gyr = ld3(tail + 4);
    const M33 R_i = mul33(R_imu, rodrigues(gyr, dt));
    const D3 P_i = d3(p.x, p.y, p.z);
    const D3 T_ei = d3(((pos.x + vel.x * dt) + ((0.5 * acc.x) * dt) * dt) - e.pos.x,
                       ((pos.y + vel.y * dt) + ((0.5 * acc.y) * dt) * dt) - e.pos.y,
                       ((pos.z + vel.z * dt) + ((0.5 * acc.z) * dt) * dt) - e.pos.z);
    const D3 inner = mul33v(R_i, qrot(e.offR, P_i) + e.offT) + T_ei;
    const D3 out = qrot(e.offR_inv, qrot(e.rot_inv, inner) - e.offT);
    p.x = float(out.x); p.y = float(out.y); p.z = float(out.z);
}

// The reference sweeps points and IMU segments backwards together (:314-345).  For time-sorted points that
// is: point i belongs to the LAST segment kp whose head is strictly older than the point; points older than
// every head stay untouched.  One quirk is kept: the sweep `break`s on the first point and then re-tests
// it against every earlier segment, so point 0 is compensated once per earlier segment that is older than it.
__global__ void k_undistort(float4* __restrict__ pts, const float* __restrict__ t_ms, int n,
                            const double* __restrict__ poses, int n_pose, const double* __restrict__ x_end) {
    extern __shared__ double s_pose[];
    for (int i = threadIdx.x; i < n_pose * POSE_DOUBLES; i += blockDim.x) s_pose[i] = poses[i];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EndState e;
    e.pos = ld3(x_end);
    e.rot_inv = qconj(ldq(x_end + 3));
    e.offR = ldq(x_end + 7);
    e.offR_inv = qconj(e.offR);
    e.offT = ld3(x_end + 11);
    const double t = double(t_ms[i]) / double(1000);
    float4 p = pts[i];
    bool touched = false;
    for (int kp = n_pose - 1; kp >= 1; kp--) {
        const double* head = s_pose + (kp - 1) * POSE_DOUBLES;
        if (t > head[0]) {
            compensate(p, t, head, head + POSE_DOUBLES, e);
            touched = true;
            if (i != 0) break;
        }
    }
    if (touched) pts[i] = p;
}

// ------------------------------------------------------------------------------------------------ voxel grid
struct VgCtl {
    float mn[3], mx[3];
    int total;          // number of output points
    int passthrough;    // PCL's "leaf size too small" exit: output = input
};

__global__ void k_vg_reset(VgCtl* c) {
    if (threadIdx.x < 3) { c->mn[threadIdx.x] = FLT_MAX; c->mx[threadIdx.x] = -FLT_MAX; }
    if (threadIdx.x == 3) { c->total = 0; c->passthrough = 0; }
}

// getMinMax3D (pcl/common/impl/common.hpp) over a dense cloud
__global__ void k_vg_minmax(const float4* __restrict__ pts, int n, VgCtl* c) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    // one set of atomics per BLOCK: all of them hit the same six words, so every atomic that is saved is latency saved
    __shared__ float s_mn[8][3], s_mx[8][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) { s_mn[warp][a] = mn[a]; s_mx[warp][a] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int a = threadIdx.x;
        float lo = s_mn[0][a], hi = s_mx[0][a];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) { lo = fminf(lo, s_mn[w][a]); hi = fmaxf(hi, s_mx[w][a]); }
        atomic_min_float(&c->mn[a], lo);
        atomic_max_float(&c->mx[a], hi);
    }
}

struct VgGrid {
    float inv;
    int min_b[3], mul[3];
    bool overflow;
};
// pcl::VoxelGrid::applyFilter: leaf-size check, min_b_/div_b_/divb_mul_
__device__ __forceinline__ VgGrid vg_grid(const VgCtl* c, float leaf) {
    VgGrid g;
    g.inv = __fdiv_rn(1.0f, leaf);
    long long d[3];
    int div_b[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float mn = c->mn[a], mx = c->mx[a];
        d[a] = (long long)(__fmul_rn(__fsub_rn(mx, mn), g.inv)) + 1;
        g.min_b[a] = int(floorf(__fmul_rn(mn, g.inv)));
        div_b[a] = int(floorf(__fmul_rn(mx, g.inv))) - g.min_b[a] + 1;
    }
    g.overflow = d[0] * d[1] * d[2] > (long long)INT_MAX;
    g.mul[0] = 1; g.mul[1] = div_b[0]; g.mul[2] = div_b[0] * div_b[1];
    return g;
}

__global__ void k_vg_keys(const float4* __restrict__ pts, int n, float leaf, VgCtl* c, unsigned* __restrict__ keys, int* __restrict__ vals) {
    const VgGrid g = vg_grid(c, leaf);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) c->passthrough = g.overflow ? 1 : 0;
    if (i >= n) return;
    const float4 p = pts[i];
    const int i0 = int(__fsub_rn(floorf(__fmul_rn(p.x, g.inv)), float(g.min_b[0])));
    const int i1 = int(__fsub_rn(floorf(__fmul_rn(p.y, g.inv)), float(g.min_b[1])));
    const int i2 = int(__fsub_rn(floorf(__fmul_rn(p.z, g.inv)), float(g.min_b[2])));
    // in passthrough mode the "cell" is the point itself, which turns the rest of the pipeline into a copy
    keys[i] = g.overflow ? unsigned(i) : unsigned(i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2]);
    vals[i] = i;
}

__global__ void k_vg_heads(const unsigned* __restrict__ keys, int n, int* __restrict__ heads) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) heads[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// CentroidPoint per occupied cell: float sums in ascending input index (the radix sort is stable), then / count.
// One thread per cell walks its run; raw scans put a handful of points into a cell.
__global__ void k_vg_centroid(const float4* __restrict__ pts, const unsigned* __restrict__ keys, const int* __restrict__ vals,
                              const int* __restrict__ heads, const int* __restrict__ pos, int n, float4* __restrict__ out, VgCtl* c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == n - 1) c->total = pos[i] + heads[i];
    if (!heads[i]) return;
    const unsigned key = keys[i];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int j = i;
    do {
        const float4 p = pts[vals[j]];
        sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); si = __fadd_rn(si, p.w);
        j++;
    } while (j < n && keys[j] == key);
    const float cnt = float(j - i);
    out[pos[i]] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), __fdiv_rn(si, cnt));
}

}  // namespace

// ================================================================================================ ScanFrontEnd
ScanFrontEnd::~ScanFrontEnd() {
    cudaSetDevice(map_->device());
    DeviceBuffer* all[] = {&raw_, &raw_alt_, &time_, &time_alt_, &down_, &keys_, &keys_alt_, &vals_, &vals_alt_, &heads_, &pos_, &cub_tmp_, &ctl_, &poses_};
    for (DeviceBuffer* b : all) b->release();
    if (h_count_) cudaFreeHost(h_count_);
}

int ScanFrontEnd::init() {
    FL_CUDA(cudaSetDevice(map_->device()));
    FL_CHECK(ctl_.reserve(sizeof(VgCtl)));
    FL_CUDA(cudaMallocHost(&h_count_, sizeof(int)));
    return FL_OK;
}

int ScanFrontEnd::upload(const float* xyzi, const float* offset_ms, int n) {
    if (n < 0 || (n > 0 && (!xyzi || !offset_ms))) { set_last_error("scan upload: bad arguments"); return FL_ERR_ARG; }
    FL_CUDA(cudaSetDevice(map_->device()));
    const size_t m = (size_t)std::max(1, n);
    FL_CHECK(raw_.reserve(sizeof(float4) * m));
    FL_CHECK(time_.reserve(sizeof(float) * m));
    if (n > 0) {
        FL_CUDA(cudaMemcpyAsync(raw_.ptr, xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, map_->stream()));
        FL_CUDA(cudaMemcpyAsync(time_.ptr, offset_ms, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, map_->stream()));
    }
    n_raw_ = n;
    n_down_ = 0;
    return FL_OK;
}

int ScanFrontEnd::undistort(const double* poses, int n_pose, const double* x26_end) {
    if (n_pose < 0 || (n_pose > 0 && !poses) || !x26_end) { set_last_error("undistort: bad arguments"); return FL_ERR_ARG; }
    const int n = n_raw_;
    if (n == 0) return FL_OK;
    FL_CUDA(cudaSetDevice(map_->device()));
    cudaStream_t st = map_->stream();
    // sort(pcl_out.points.begin(), pcl_out.points.end(), time_list)  (:234) -- stable here
    FL_CHECK(raw_alt_.reserve(sizeof(float4) * (size_t)n));
    FL_CHECK(time_alt_.reserve(sizeof(float) * (size_t)n));
    size_t tmp = 0;
    FL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, time_.as<float>(), time_alt_.as<float>(), raw_.as<float4>(), raw_alt_.as<float4>(), n, 0, 32, st));
    FL_CHECK(cub_tmp_.reserve(tmp));
    FL_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp_.ptr, tmp, time_.as<float>(), time_alt_.as<float>(), raw_.as<float4>(), raw_alt_.as<float4>(), n, 0, 32, st));
    std::swap(raw_, raw_alt_);
    std::swap(time_, time_alt_);
    if (n_pose < 2) return FL_OK;                       // no segment: the backward sweep has nothing to walk
    const size_t np = (size_t)n_pose * POSE_DOUBLES;
    const size_t smem = sizeof(double) * np;
    if (smem > 200 * 1024) { set_last_error("undistort: %d IMU poses do not fit shared memory", n_pose); return FL_ERR_CAPACITY; }
    // a few KB from pageable host memory: the runtime stages them before returning, the caller's arrays are free at once
    FL_CHECK(poses_.reserve(sizeof(double) * (np + XLEN)));
    FL_CUDA(cudaMemcpyAsync(poses_.ptr, poses, sizeof(double) * np, cudaMemcpyHostToDevice, st));
    FL_CUDA(cudaMemcpyAsync(poses_.as<double>() + np, x26_end, sizeof(double) * XLEN, cudaMemcpyHostToDevice, st));
    if (smem > 48 * 1024) FL_CUDA(cudaFuncSetAttribute(k_undistort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int block = 128;
    k_undistort<<<(n + block - 1) / block, block, smem, st>>>(raw_.as<float4>(), time_.as<float>(), n, poses_.as<double>(), n_pose,
                                                              poses_.as<double>() + (size_t)n_pose * POSE_DOUBLES);
    FL_CUDA(cudaGetLastError());
    return FL_OK;
}

int ScanFrontEnd::voxel_downsample(float leaf, int* n_out) {
    if (n_out) *n_out = 0;
    if (!(leaf > 0.f)) { set_last_error("voxel_downsample: leaf size must be > 0"); return FL_ERR_ARG; }
    const int n = n_raw_;
    n_down_ = 0;
    if (n == 0) return FL_OK;
    FL_CUDA(cudaSetDevice(map_->device()));
    cudaStream_t st = map_->stream();
    FL_CHECK(down_.reserve(sizeof(float4) * (size_t)n));
    FL_CHECK(keys_.reserve(sizeof(unsigned) * (size_t)n));
    FL_CHECK(keys_alt_.reserve(sizeof(unsigned) * (size_t)n));
    FL_CHECK(vals_.reserve(sizeof(int) * (size_t)n));
    FL_CHECK(vals_alt_.reserve(sizeof(int) * (size_t)n));
    FL_CHECK(heads_.reserve(sizeof(int) * (size_t)n));
    FL_CHECK(pos_.reserve(sizeof(int) * (size_t)n));
    VgCtl* ctl = ctl_.as<VgCtl>();
    const int block = 256, grid = (n + block - 1) / block;
    k_vg_reset<<<1, 32, 0, st>>>(ctl);
    k_vg_minmax<<<std::min(grid, 148), block, 0, st>>>(raw_.as<float4>(), n, ctl);
    k_vg_keys<<<grid, block, 0, st>>>(raw_.as<float4>(), n, leaf, ctl, keys_.as<unsigned>(), vals_.as<int>());
    size_t tmp_sort = 0, tmp_scan = 0;
    FL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys_.as<unsigned>(), keys_alt_.as<unsigned>(), vals_.as<int>(), vals_alt_.as<int>(), n, 0, 32, st));
    FL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, heads_.as<int>(), pos_.as<int>(), n, st));
    FL_CHECK(cub_tmp_.reserve(std::max(tmp_sort, tmp_scan)));
    FL_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp_.ptr, tmp_sort, keys_.as<unsigned>(), keys_alt_.as<unsigned>(), vals_.as<int>(), vals_alt_.as<int>(), n, 0, 32, st));
    k_vg_heads<<<grid, block, 0, st>>>(keys_alt_.as<unsigned>(), n, heads_.as<int>());
    FL_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp_.ptr, tmp_scan, heads_.as<int>(), pos_.as<int>(), n, st));
    k_vg_centroid<<<grid, block, 0, st>>>(raw_.as<float4>(), keys_alt_.as<unsigned>(), vals_alt_.as<int>(), heads_.as<int>(), pos_.as<int>(), n,
                                          down_.as<float4>(), ctl);
    FL_CUDA(cudaGetLastError());
    FL_CUDA(cudaMemcpyAsync(h_count_, &ctl->total, sizeof(int), cudaMemcpyDeviceToHost, st));
    FL_CUDA(cudaStreamSynchronize(st));
    n_down_ = *h_count_;
    if (n_out) *n_out = n_down_;
    return FL_OK;
}

int ScanFrontEnd::download(int which, float* out_xyzi, int cap, int* n) {
    const int have = which == 0 ? n_raw_ : n_down_;
    if (n) *n = have;
    if (which != 0 && which != 1) { set_last_error("scan download: which must be 0 or 1"); return FL_ERR_ARG; }
    const int take = std::min(have, cap);
    if (take <= 0) return FL_OK;
    if (!out_xyzi) { set_last_error("scan download: null buffer"); return FL_ERR_ARG; }
    FL_CUDA(cudaSetDevice(map_->device()));
    const void* src = which == 0 ? raw_.ptr : down_.ptr;
    FL_CUDA(cudaMemcpyAsync(out_xyzi, src, sizeof(float4) * (size_t)take, cudaMemcpyDeviceToHost, map_->stream()));
    FL_CUDA(cudaStreamSynchronize(map_->stream()));
    return FL_OK;
}

// ================================================================================================ LocalMapCube
// The arithmetic keeps the reference's types: cube corners are float (BoxPointType, ikd_Tree.h:42-45), the
// LiDAR position and cube_len are double, MOV_THRESHOLD (1.5f) and DET_RANGE are float (laserMapping.cpp:77-78).
int LocalMapCube::slide(const double pos[3], float* boxes6) {
    const float margin = 1.5f * det_range_;
    if (!init_) {                                      // :238-245: first call only centres the cube
        for (int a = 0; a < 3; a++) {
            lo_[a] = float(pos[a] - cube_len_ / 2.0);
            hi_[a] = float(pos[a] + cube_len_ / 2.0);
        }
        init_ = true;
        return 0;
    }
    enum Dir { STAY, TO_LOW, TO_HIGH } dir[3];
    bool near_edge = false;
    for (int a = 0; a < 3; a++) {
        const float to_lo = float(std::fabs(pos[a] - double(lo_[a])));
        const float to_hi = float(std::fabs(pos[a] - double(hi_[a])));
        dir[a] = to_lo <= margin ? TO_LOW : (to_hi <= margin ? TO_HIGH : STAY);   // the low face wins (:259-264)
        near_edge = near_edge || dir[a] != STAY;
    }
    if (!near_edge) return 0;
    const float step = float(std::max((cube_len_ - 2.0 * 1.5f * det_range_) * 0.5 * 0.9, double(det_range_ * (1.5f - 1))));   // :256
    int nb = 0;
    float new_lo[3], new_hi[3];
    for (int a = 0; a < 3; a++) {
        new_lo[a] = lo_[a]; new_hi[a] = hi_[a];
        if (dir[a] == STAY) continue;
        float* b = boxes6 + nb * 6;                    // the slab the cube leaves behind, spanning the OLD cube on the other axes
        for (int c = 0; c < 3; c++) { b[c] = lo_[c]; b[3 + c] = hi_[c]; }
        if (dir[a] == TO_LOW) {
            new_hi[a] = hi_[a] - step; new_lo[a] = lo_[a] - step;
            b[a] = hi_[a] - step;
        } else {
            new_hi[a] = hi_[a] + step; new_lo[a] = lo_[a] + step;
            b[3 + a] = lo_[a] + step;
        }
        nb++;
    }
    for (int a = 0; a < 3; a++) { lo_[a] = new_lo[a]; hi_[a] = new_hi[a]; }
    return nb;
}

void LocalMapCube::get(float* box6) const {
    for (int a = 0; a < 3; a++) { box6[a] = lo_[a]; box6[3 + a] = hi_[a]; }
}

}  // namespace fl
