// Fused measurement kernel + on-device iterated-EKF solve for FAST-LIO2's per-scan update.
//
//   k_measure : h_share_model (reference src/laserMapping.cpp:638-754) for all scan points:
//               body->world transform, k=5 nearest-neighbour search in the device map (search
//               passes only), 5-point plane fit (esti_plane, include/common_lib.h:225-257),
//               residual gating, Jacobian row -- and, instead of materialising h_x (m x 12) and
//               h (m), the FP64 normal equations H^T H (12x12) / H^T h (12) that
//               update_iterated_dyn_share_modified consumes (esekfom.hpp:1784,1804), reduced with
//               warp shuffles and one deterministic per-block partial.
//   k_solve   : esekfom.hpp:1651-1927 -- boxminus, manifold congruences on P, the Kalman gain
//               algebra, boxplus, convergence bookkeeping, final covariance -- in one thread block,
//               so that the whole multi-pass update runs without a host round trip.
#include <dlfcn.h>

#include <algorithm>

#include "filter.h"

namespace fl {

constexpr int PSTRIDE = 96;          // doubles per partial row (NRED = 92 padded)
constexpr int MEASURE_THREADS = 256;
constexpr int SOLVE_THREADS = 512;
constexpr int MAX_LOGS = 16;

__device__ __forceinline__ int tri12(int a, int b) { return a * 12 - (a * (a - 1)) / 2 + (b - a); }   // a <= b

// ============================================================================= esti_plane
// float32 least squares A n = -1 on the 5 neighbours, column-pivoted Householder QR -- the
// algorithm the reference gets from Eigen (common_lib.h:241, colPivHouseholderQr().solve) --
// followed by the reference's normalisation and 0.1 m point-to-plane check (common_lib.h:243-256).
__device__ __forceinline__ bool esti_plane_dev(float pabcd[4], const float (&pt)[KNN_K][3], float threshold) {
    constexpr int rows = KNN_K, cols = 3, size = 3;
    float qr[rows][cols];
    float c[rows];
#pragma unroll
    for (int j = 0; j < rows; j++) { qr[j][0] = pt[j][0]; qr[j][1] = pt[j][1]; qr[j][2] = pt[j][2]; c[j] = -1.0f; }
    float hc[size];
    int transp[size];
    float nu[cols], nd[cols];
#pragma unroll
    for (int k = 0; k < cols; k++) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < rows; i++) s += qr[i][k] * qr[i][k];
        nd[k] = sqrtf(s); nu[k] = nd[k];
    }
    const float eps = 1.1920929e-07f;
    const float maxn = fmaxf(nu[0], fmaxf(nu[1], nu[2]));
    const float threshold_helper = (maxn * eps) * (maxn * eps) / float(rows);
    const float norm_downdate_threshold = sqrtf(eps);
    int nonzero_pivots = size;
#pragma unroll
    for (int k = 0; k < size; k++) {
        int biggest = k; float bn = nu[k];
#pragma unroll
        for (int j = k + 1; j < cols; j++) if (nu[j] > bn) { bn = nu[j]; biggest = j; }
        if (nonzero_pivots == size && bn * bn < threshold_helper * float(rows - k)) nonzero_pivots = k;
        transp[k] = biggest;
        if (k != biggest) {
#pragma unroll
            for (int j = k + 1; j < cols; j++) {
                if (j == biggest) {
#pragma unroll
                    for (int i = 0; i < rows; i++) { float t = qr[i][k]; qr[i][k] = qr[i][j]; qr[i][j] = t; }
                    float t = nu[k]; nu[k] = nu[j]; nu[j] = t;
                    t = nd[k]; nd[k] = nd[j]; nd[j] = t;
                }
            }
        }
        float tailSq = 0.f;
#pragma unroll
        for (int i = k + 1; i < rows; i++) tailSq += qr[i][k] * qr[i][k];
        const float c0 = qr[k][k];
        float tau, beta;
        if (tailSq <= 1.17549435e-38f) {
            tau = 0.f; beta = c0;
#pragma unroll
            for (int i = k + 1; i < rows; i++) qr[i][k] = 0.f;
        } else {
            beta = sqrtf(c0 * c0 + tailSq);
            if (c0 >= 0.f) beta = -beta;
#pragma unroll
            for (int i = k + 1; i < rows; i++) qr[i][k] = qr[i][k] / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        hc[k] = tau;
        qr[k][k] = beta;
        if (tau != 0.f) {
#pragma unroll
            for (int j = k + 1; j < cols; j++) {
                float tmp = 0.f;
#pragma unroll
                for (int i = k + 1; i < rows; i++) tmp += qr[i][k] * qr[i][j];
                tmp += qr[k][j];
                qr[k][j] -= tau * tmp;
#pragma unroll
                for (int i = k + 1; i < rows; i++) qr[i][j] -= tau * qr[i][k] * tmp;
            }
        }
#pragma unroll
        for (int j = k + 1; j < cols; j++) {
            if (nu[j] != 0.f) {
                float temp = fabsf(qr[k][j]) / nu[j];
                temp = (1.f + temp) * (1.f - temp);
                temp = temp < 0.f ? 0.f : temp;
                const float r = nu[j] / nd[j];
                const float temp2 = temp * r * r;
                if (temp2 <= norm_downdate_threshold) {
                    float s = 0.f;
#pragma unroll
                    for (int i = k + 1; i < rows; i++) s += qr[i][j] * qr[i][j];
                    nd[j] = sqrtf(s); nu[j] = nd[j];
                } else {
                    nu[j] *= sqrtf(temp);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < size; k++) {
        if (k < nonzero_pivots) {
            const float tau = hc[k];
            if (tau != 0.f) {
                float tmp = 0.f;
#pragma unroll
                for (int i = k + 1; i < rows; i++) tmp += qr[i][k] * c[i];
                tmp += c[k];
                c[k] -= tau * tmp;
#pragma unroll
                for (int i = k + 1; i < rows; i++) c[i] -= tau * qr[i][k] * tmp;
            }
        }
    }
#pragma unroll
    for (int i = size - 1; i >= 0; i--) {
        if (i < nonzero_pivots) {
            float s = c[i];
#pragma unroll
            for (int j = i + 1; j < size; j++) if (j < nonzero_pivots) s -= qr[i][j] * c[j];
            c[i] = s / qr[i][i];
        }
    }
    int perm[3] = {0, 1, 2};
#pragma unroll
    for (int k = 0; k < size; k++) {
        // swap(perm[k], perm[transp[k]]) with static indexing
#pragma unroll
        for (int j = 0; j < size; j++) if (j == transp[k] && j != k) { int t = perm[k]; perm[k] = perm[j]; perm[j] = t; }
    }
    float x[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < size; i++) {
        if (i < nonzero_pivots) {
#pragma unroll
            for (int j = 0; j < 3; j++) if (perm[i] == j) x[j] = c[i];
        }
    }
    const float n = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    pabcd[0] = x[0] / n; pabcd[1] = x[1] / n; pabcd[2] = x[2] / n;
    pabcd[3] = (float)(1.0 / (double)n);
    bool ok = true;
#pragma unroll
    for (int j = 0; j < rows; j++)
        if (fabsf(pabcd[0] * pt[j][0] + pabcd[1] * pt[j][1] + pabcd[2] * pt[j][2] + pabcd[3]) > threshold) ok = false;
    return ok;
}

// ============================================================================= measurement model
struct PoseS {            // the part of the state h_share_model reads
    Q4 rot, offR; D3 pos, offT;
};
__device__ __forceinline__ PoseS load_pose(const double* x) {
    PoseS p; p.rot = ldq(x + X_ROT); p.offR = ldq(x + X_OFFR); p.pos = ld3(x + X_POS); p.offT = ld3(x + X_OFFT);
    return p;
}
// laserMapping.cpp:656-661
__device__ __forceinline__ void body_to_world(const PoseS& s, const float4& pb, float& wx, float& wy, float& wz) {
    D3 p_body = d3(pb.x, pb.y, pb.z);
    D3 g = qrot(s.rot, qrot(s.offR, p_body) + s.offT) + s.pos;
    wx = (float)g.x; wy = (float)g.y; wz = (float)g.z;
}
// laserMapping.cpp:723-751: one row of h_x (12 wide) and the entry of h
template <bool EXTR>
__device__ __forceinline__ void jacobian_row(const PoseS& s, const float4& pb, const float4& nv, double* h, double& z) {
    D3 p_be = d3(pb.x, pb.y, pb.z);
    D3 p_this = qrot(s.offR, p_be) + s.offT;
    D3 n = d3(nv.x, nv.y, nv.z);
    D3 C = qrot(qconj(s.rot), n);
    D3 A = mul33v(hat3(p_this), C);
    h[0] = n.x; h[1] = n.y; h[2] = n.z; h[3] = A.x; h[4] = A.y; h[5] = A.z;
    if (EXTR) {
        D3 B = mul33v(mul33(hat3(p_be), qmat(qconj(s.offR))), C);
        h[6] = B.x; h[7] = B.y; h[8] = B.z; h[9] = C.x; h[10] = C.y; h[11] = C.z;
    }
    z = -(double)nv.w;
}

// Per-point part of h_share_model after the search (laserMapping.cpp:674-692).
// Returns true when the point contributes a row.
template <bool EXTR>
__device__ __forceinline__ bool measure_point(const ScanView& sc, int q, const PoseS& s, double* h, double& z, float& absres) {
    if (!sc.selected[q]) return false;                                   // :674
    sc.selected[q] = 0;                                                  // :677
    const float4 pb = __ldg(&sc.body[q]);
    float wx, wy, wz;
    body_to_world(s, pb, wx, wy, wz);
    float pn[KNN_K][3];
#pragma unroll
    for (int j = 0; j < KNN_K; j++) { const float4 p = sc.nearest[(size_t)q * KNN_K + j]; pn[j][0] = p.x; pn[j][1] = p.y; pn[j][2] = p.z; }
    float pabcd[4];
    if (!esti_plane_dev(pabcd, pn, 0.1f)) return false;                  // :678
    const float pd2 = pabcd[0] * wx + pabcd[1] * wy + pabcd[2] * wz + pabcd[3];          // :680
    const D3 p_body = d3(pb.x, pb.y, pb.z);
    const float score = (float)(1 - 0.9 * fabs((double)pd2) / sqrt(norm3(p_body)));     // :681 (T8)
    if (!((double)score > 0.9)) return false;                            // :683
    sc.selected[q] = 1;
    const float4 nv = make_float4(pabcd[0], pabcd[1], pabcd[2], pd2);
    sc.normvec[q] = nv;
    absres = fabsf(pd2);                                                 // res_last
    jacobian_row<EXTR>(s, pb, nv, h, z);
    return true;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Fold the rows held by the 32 lanes into the lane-distributed accumulators: output o of the
// NRED-vector lives in lane (o & 31), register acc[o >> 5].
template <bool EXTR>
__device__ __forceinline__ void warp_accumulate(bool contrib, const double* h, double z, float absres, double (&acc)[3], int lane) {
    constexpr int NC = EXTR ? 12 : 6;
    const unsigned any = __ballot_sync(FULL, contrib);
    if (!any) return;
#pragma unroll
    for (int a = 0; a < NC; a++) {
#pragma unroll
        for (int b = a; b < NC; b++) {
            const int o = a * 12 - (a * (a - 1)) / 2 + (b - a);
            const double v = warp_sum(contrib ? h[a] * h[b] : 0.0);
            if (lane == (o & 31)) acc[o >> 5] += v;
        }
    }
#pragma unroll
    for (int a = 0; a < NC; a++) {
        const int o = 78 + a;
        const double v = warp_sum(contrib ? h[a] * z : 0.0);
        if (lane == (o & 31)) acc[o >> 5] += v;
    }
    {
        const double v = (double)__popc(any);
        if (lane == (90 & 31)) acc[90 >> 5] += v;
        const double r = warp_sum(contrib ? (double)absres : 0.0);
        if (lane == (91 & 31)) acc[91 >> 5] += r;
    }
}

// One launch per iEKF pass.  Search passes: one warp per scan point (kNN), then the per-point
// tail thread-parallel over the points the warp has served.  Non-search passes: one thread per
// point.  Which of the two runs is decided on the device (ctl->converge), so the launch sequence
// of a scan is fixed and needs no host synchronisation.
template <bool EXTR>
__global__ void __launch_bounds__(MEASURE_THREADS) k_measure(MapView m, ScanView sc, const FilterCtl* __restrict__ ctl,
                                                              double* __restrict__ partials) {
    if (ctl->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_per_block = MEASURE_THREADS / 32;
    const int gwarp = blockIdx.x * warps_per_block + warp;
    const int nwarps = gridDim.x * warps_per_block;
    const bool search = ctl->converge != 0;
    const PoseS s = load_pose(ctl->x);
    double acc[3] = {0.0, 0.0, 0.0};
    const int q0 = sc.q_begin, q1 = sc.q_end;

    if (search) {
        int mine = -1;          // the query this lane will run the tail for
        int n_mine = 0;
        for (int q = q0 + gwarp; q < q1; q += nwarps) {
            const float4 pb = __ldg(&sc.body[q]);
            float wx, wy, wz;
            body_to_world(s, pb, wx, wy, wz);
            KBest kb;
            knn_query(m, wx, wy, wz, kb, lane);
            int myidx = -1;
#pragma unroll
            for (int j = 0; j < KNN_K; j++) if (lane == j) myidx = kb.idx[j];
            if (lane < KNN_K) {
                float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                if (myidx >= 0) { p = m.pts[myidx]; p.w = m.payload[myidx]; }
                sc.nearest[(size_t)q * KNN_K + lane] = p;
            }
            if (lane == 0) {
                int cnt = 0;
#pragma unroll
                for (int j = 0; j < KNN_K; j++) cnt += kb.idx[j] >= 0 ? 1 : 0;
                sc.nearest_cnt[q] = cnt;
                // laserMapping.cpp:671
                sc.selected[q] = (cnt < KNN_K) ? 0 : (kb.d[KNN_K - 1] > 5.0f ? 0 : 1);
            }
            if (lane == n_mine) mine = q;
            n_mine++;
            if (n_mine == 32) {
                __syncwarp();
                double h[12]; double z = 0.0; float ar = 0.f;
                const bool contrib = measure_point<EXTR>(sc, mine, s, h, z, ar);
                warp_accumulate<EXTR>(contrib, h, z, ar, acc, lane);
                n_mine = 0; mine = -1;
            }
        }
        if (__any_sync(FULL, n_mine > 0)) {
            __syncwarp();
            double h[12]; double z = 0.0; float ar = 0.f;
            bool contrib = false;
            if (lane < n_mine) contrib = measure_point<EXTR>(sc, mine, s, h, z, ar);
            warp_accumulate<EXTR>(contrib, h, z, ar, acc, lane);
        }
    } else {
        for (int base = q0 + gwarp * 32; base < q1; base += nwarps * 32) {
            const int q = base + lane;
            double h[12]; double z = 0.0; float ar = 0.f;
            bool contrib = false;
            if (q < q1) contrib = measure_point<EXTR>(sc, q, s, h, z, ar);
            warp_accumulate<EXTR>(contrib, h, z, ar, acc, lane);
        }
    }
    // deterministic block partial: warps -> shared -> fixed-order sum
    __shared__ double wacc[MEASURE_THREADS / 32][PSTRIDE];
#pragma unroll
    for (int j = 0; j < 3; j++) wacc[warp][lane + 32 * j] = acc[j];
    __syncthreads();
    if (threadIdx.x < PSTRIDE) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < MEASURE_THREADS / 32; w++) v += wacc[w][threadIdx.x];
        partials[(size_t)blockIdx.x * PSTRIDE + threadIdx.x] = v;
    }
}

// ============================================================================= solve
// Gauss-Jordan with (logical) partial pivoting on an n x nc augmented system [A | B] held in
// shared memory; on return  A^{-1} B  is read through row_of[]:  X[k][j] = a[row_of[k]][n + j] / a[row_of[k]][k].
__device__ bool gj_eliminate(double* a, int n, int nc, int ld, double* colbuf, int* row_of, int* used, int* s_piv) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid < n) used[tid] = 0;
    __syncthreads();
    bool ok = true;
    for (int k = 0; k < n; k++) {
        if (tid < 32) {
            double best = -1.0; int bi = -1;
            if (tid < n && !used[tid]) { best = fabs(a[tid * ld + k]); bi = tid; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(FULL, best, o);
                const int oi = __shfl_xor_sync(FULL, bi, o);
                if (ob > best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
            }
            if (tid == 0) { *s_piv = (best > 0.0) ? bi : -1; }
        }
        __syncthreads();
        const int p = *s_piv;
        if (p < 0) { ok = false; break; }
        const double pv = a[p * ld + k];
        if (tid < n) colbuf[tid] = (tid == p) ? 0.0 : a[tid * ld + k] / pv;
        if (tid == 0) { used[p] = 1; row_of[k] = p; }
        __syncthreads();
        for (int e = tid; e < n * nc; e += nt) {
            const int i = e / nc, j = e % nc;
            const double f = colbuf[i];
            if (f != 0.0) a[i * ld + j] -= f * a[p * ld + j];
        }
        __syncthreads();
    }
    return ok;
}


struct SolveShared {
    double red[PSTRIDE];
    double HTH[144];
    double Hth[12];
    double P[NDOF * NDOF];        // P_propagated after the manifold congruence (esekfom.hpp:1657-1699)
    double L[NDOF * NDOF];        // scratch, then L_ of the final covariance step
    double aug[NDOF * 2 * NDOF];  // augmented system for Gauss-Jordan
    double Kx[NDOF * 12];         // K_x[:, 0:12]   (columns 12..22 are zero in every branch)
    double Kh[NDOF];
    double dx[NDOF], dx_new[NDOF], dxu[NDOF];
    double J[2][9];
    double M2[4];
    double xnew[XLEN];
    double colbuf[32];
    double rows[22 * 13];         // small-m branch: [h_x row (12) | h]
    double PHt[NDOF * 22];        // small-m branch
    double T[22 * 13];            // small-m branch: S^{-1} [h_x | h]
    int row_of[32], used[32], piv, m_rows, finish;
    int row_idx[22];
};

// rows {3..5, 6..8, 21..22} of dst := J * (same rows of src), first `ncols` columns
__device__ void apply_rows(double* dst, const double* src, const double* J3, const double* J6, const double* M2, int ncols, int ld) {
    const int i = threadIdx.x;
    if (i < ncols) {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int idx = b == 0 ? 3 : 6;
            const double* J = b == 0 ? J3 : J6;
            const double v0 = src[idx * ld + i], v1 = src[(idx + 1) * ld + i], v2 = src[(idx + 2) * ld + i];
            dst[idx * ld + i] = J[0] * v0 + J[1] * v1 + J[2] * v2;
            dst[(idx + 1) * ld + i] = J[3] * v0 + J[4] * v1 + J[5] * v2;
            dst[(idx + 2) * ld + i] = J[6] * v0 + J[7] * v1 + J[8] * v2;
        }
        const double a = src[21 * ld + i], bq = src[22 * ld + i];
        dst[21 * ld + i] = M2[0] * a + M2[1] * bq;
        dst[22 * ld + i] = M2[2] * a + M2[3] * bq;
    }
}
// columns {3..5, 6..8, 21..22} of the 23x23 `mat` := (row block) * J^T, for every row
__device__ void apply_cols(double* mat, const double* J3, const double* J6, const double* M2) {
    const int i = threadIdx.x;
    if (i < NDOF) {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int idx = b == 0 ? 3 : 6;
            const double* J = b == 0 ? J3 : J6;
            const double v0 = mat[i * NDOF + idx], v1 = mat[i * NDOF + idx + 1], v2 = mat[i * NDOF + idx + 2];
            mat[i * NDOF + idx] = J[0] * v0 + J[1] * v1 + J[2] * v2;
            mat[i * NDOF + idx + 1] = J[3] * v0 + J[4] * v1 + J[5] * v2;
            mat[i * NDOF + idx + 2] = J[6] * v0 + J[7] * v1 + J[8] * v2;
        }
        const double a = mat[i * NDOF + 21], bq = mat[i * NDOF + 22];
        mat[i * NDOF + 21] = a * M2[0] + bq * M2[1];
        mat[i * NDOF + 22] = a * M2[2] + bq * M2[3];
    }
}
// thread 0: the three congruence blocks for the tangent vector d (esekfom.hpp:1659-1690 / 1836-1876)
__device__ void make_congruence(const double* d, const double* x_now, const double* x_prop, double (*J)[9], double* M2) {
    for (int b = 0; b < 2; b++) {
        const int idx = b == 0 ? 3 : 6;
        M33 Jm = transpose33(A_matrix(d3(d[idx], d[idx + 1], d[idx + 2])));        // T5: A_matrix(dx)^T
        for (int i = 0; i < 9; i++) J[b][i] = Jm.m[i];
    }
    S2_congruence(ld3(x_now + X_GRAV), ld3(x_prop + X_GRAV), d[21], d[22], M2);
}

// mode 0: reduce the block partials and solve (single GPU);
// mode 1: reduce only -> red_g (an all-reduce over the ranks follows);
// mode 2: solve from red_g.
__global__ void __launch_bounds__(SOLVE_THREADS) k_solve(FilterCtl* ctl, const double* __restrict__ partials, int n_partials,
                                                          double* red_g, int mode, ScanView sc, PassLog* logs, int solver) {
    __shared__ SolveShared S;
    const int tid = threadIdx.x, nt = blockDim.x;
    constexpr int n = NDOF;
    if (ctl->done) return;
    // ------------------------------------------------------------------ reduce
    if (mode != 2) {
        const int o = tid >> 2, sub = tid & 3;
        double v = 0.0;
        if (o < NRED) for (int b = sub; b < n_partials; b += 4) v += partials[(size_t)b * PSTRIDE + o];
        v += __shfl_xor_sync(FULL, v, 1);
        v += __shfl_xor_sync(FULL, v, 2);
        if (o < NRED && sub == 0) { S.red[o] = v; if (mode == 1) red_g[o] = v; }
        if (mode == 1) return;
    } else {
        if (tid < NRED) S.red[tid] = red_g[tid];
    }
    __syncthreads();
    if (tid < 144) { const int a = tid / 12, b = tid % 12; S.HTH[tid] = S.red[a <= b ? tri12(a, b) : tri12(b, a)]; }
    if (tid < 12) S.Hth[tid] = S.red[78 + tid];
    const int effct = (int)(S.red[90] + 0.5);
    const int it = ctl->iter, max_iter = ctl->max_iter, n_pass = ctl->n_pass;
    const int searched = ctl->converge;
    const int t_in = ctl->t;
    const double R = ctl->R;
    PassLog* lg = (logs && n_pass < MAX_LOGS) ? &logs[n_pass] : nullptr;
    __syncthreads();
    if (lg) {
        if (tid < 144) lg->HtH[tid] = S.HTH[tid];
        if (tid < 12) lg->Hth[tid] = S.Hth[tid];
        if (tid == 0) { lg->searched = searched; lg->effct = effct; lg->res_sum = S.red[91]; lg->valid = effct >= 1; }
    }
    // ------------------------------------------------------------------ invalid pass (laserMapping.cpp:708-713, esekfom.hpp:1638-1641)
    if (effct < 1) {
        if (tid == 0) {
            if (lg) { lg->converged = searched; for (int i = 0; i < XLEN; i++) lg->x_after[i] = ctl->x[i]; }
            ctl->n_pass = n_pass + 1;
            ctl->iter = it + 1;
            if (it + 1 >= max_iter) ctl->done = 1;
        }
        return;
    }
    // ------------------------------------------------------------------ esekfom.hpp:1651-1699
    if (tid == 0) {
        state_boxminus(ctl->x, ctl->x_prop, S.dx);
        for (int i = 0; i < n; i++) S.dx_new[i] = S.dx[i];
        make_congruence(S.dx, ctl->x, ctl->x_prop, S.J, S.M2);
        for (int b = 0; b < 2; b++) {
            const int idx = b == 0 ? 3 : 6;
            const double* J = S.J[b];
            const double v0 = S.dx_new[idx], v1 = S.dx_new[idx + 1], v2 = S.dx_new[idx + 2];
            S.dx_new[idx] = J[0] * v0 + J[1] * v1 + J[2] * v2;
            S.dx_new[idx + 1] = J[3] * v0 + J[4] * v1 + J[5] * v2;
            S.dx_new[idx + 2] = J[6] * v0 + J[7] * v1 + J[8] * v2;
        }
        const double d0 = S.M2[0] * S.dx_new[21] + S.M2[1] * S.dx_new[22];
        const double d1 = S.M2[2] * S.dx_new[21] + S.M2[3] * S.dx_new[22];
        S.dx_new[21] = d0; S.dx_new[22] = d1;
    }
    for (int e = tid; e < n * n; e += nt) S.P[e] = ctl->P_prop[e];
    __syncthreads();
    // The reference interleaves row/column products per block (rows3, cols3, rows6, cols6, rows21,
    // cols21); the blocks act on disjoint index sets, so P := T P T^T either way.
    apply_rows(S.P, S.P, S.J[0], S.J[1], S.M2, n, n);
    __syncthreads();
    apply_cols(S.P, S.J[0], S.J[1], S.M2);
    __syncthreads();

    bool ok = true;
    if (effct < n) {
        // -------------------------------------------------------------- small-m branch, esekfom.hpp:1715-1744 (T6)
        //   K = P H^T (H P H^T / R + I)^{-1} / R ;  K_h = K h ;  K_x = K H
        // the m (< 23) Jacobian rows are rebuilt, in point order, from what k_measure left per point
        if (tid == 0) {
            int mrows = 0;
            for (int q = sc.q_begin; q < sc.q_end && mrows < 22; q++) if (sc.selected[q]) S.row_idx[mrows++] = q;
            S.m_rows = mrows;
        }
        __syncthreads();
        const int mr = S.m_rows;
        if (tid < mr) {
            const PoseS ps = load_pose(ctl->x);
            double h[12]; double z;
            for (int a = 0; a < 12; a++) h[a] = 0.0;
            const int q = S.row_idx[tid];
            if (ctl->extrinsic_est) jacobian_row<true>(ps, sc.body[q], sc.normvec[q], h, z);
            else jacobian_row<false>(ps, sc.body[q], sc.normvec[q], h, z);
            for (int a = 0; a < 12; a++) S.rows[tid * 13 + a] = h[a];
            S.rows[tid * 13 + 12] = z;
        }
        __syncthreads();
        for (int e = tid; e < n * mr; e += nt) {                 // PHt = P H^T  (23 x m)
            const int i = e / mr, r = e % mr;
            double v = 0.0;
            for (int k = 0; k < 12; k++) v += S.P[i * n + k] * S.rows[r * 13 + k];
            S.PHt[i * 22 + r] = v;
        }
        __syncthreads();
        const int ld = mr + 13;
        for (int e = tid; e < mr * ld; e += nt) {                // [H P H^T / R + I | h_x | h]
            const int r = e / ld, c2 = e % ld;
            double v;
            if (c2 < mr) {
                v = 0.0;
                for (int k = 0; k < 12; k++) v += S.rows[r * 13 + k] * S.PHt[k * 22 + c2];
                v = v / R + (r == c2 ? 1.0 : 0.0);
            } else {
                v = S.rows[r * 13 + (c2 - mr)];
            }
            S.aug[r * ld + c2] = v;
        }
        __syncthreads();
        ok = gj_eliminate(S.aug, mr, ld, ld, S.colbuf, S.row_of, S.used, &S.piv);
        if (ok) {
            for (int e = tid; e < mr * 13; e += nt) {            // T = S^{-1} [h_x | h]
                const int k = e / 13, j = e % 13;
                const int pr = S.row_of[k];
                S.T[e] = S.aug[pr * ld + mr + j] / S.aug[pr * ld + k];
            }
            __syncthreads();
            for (int e = tid; e < n * 13; e += nt) {             // [K_x | K_h] = PHt T / R
                const int i = e / 13, j = e % 13;
                double v = 0.0;
                for (int k = 0; k < mr; k++) v += S.PHt[i * 22 + k] * S.T[k * 13 + j];
                v /= R;
                if (j < 12) S.Kx[i * 12 + j] = v; else S.Kh[i] = v;
            }
        }
    } else if (solver == 0) {
        // -------------------------------------------------------------- information form exactly as the reference, esekfom.hpp:1782-1809
        //   P_temp = (P/R)^{-1};  P_temp[0:12,0:12] += H^T H;  P_inv = P_temp^{-1}
        for (int e = tid; e < n * 2 * n; e += nt) {
            const int i = e / (2 * n), j = e % (2 * n);
            S.aug[e] = j < n ? S.P[i * n + j] / R : (j - n == i ? 1.0 : 0.0);
        }
        __syncthreads();
        ok = gj_eliminate(S.aug, n, 2 * n, 2 * n, S.colbuf, S.row_of, S.used, &S.piv);
        if (ok) {
            for (int e = tid; e < n * n; e += nt) {
                const int k = e / n, j = e % n;
                const int pr = S.row_of[k];
                double v = S.aug[pr * 2 * n + n + j] / S.aug[pr * 2 * n + k];
                if (k < 12 && j < 12) v += S.HTH[k * 12 + j];
                S.L[e] = v;
            }
            __syncthreads();
            for (int e = tid; e < n * 2 * n; e += nt) {
                const int i = e / (2 * n), j = e % (2 * n);
                S.aug[e] = j < n ? S.L[i * n + j] : (j - n == i ? 1.0 : 0.0);
            }
            __syncthreads();
            ok = gj_eliminate(S.aug, n, 2 * n, 2 * n, S.colbuf, S.row_of, S.used, &S.piv);
        }
        if (ok) {
            // K_h = P_inv[:, 0:12] H^T h ;  K_x[:, 0:12] = P_inv[:, 0:12] H^T H
            for (int e = tid; e < n * 13; e += nt) {
                const int i = e / 13, j = e % 13;
                const int pr = S.row_of[i];
                const double piv = S.aug[pr * 2 * n + i];
                double v = 0.0;
                for (int a = 0; a < 12; a++) v += (S.aug[pr * 2 * n + n + a] / piv) * (j < 12 ? S.HTH[a * 12 + j] : S.Hth[a]);
                if (j < 12) S.Kx[i * 12 + j] = v; else S.Kh[i] = v;
            }
        }
    } else {
        // -------------------------------------------------------------- same gain through one 12x12 solve.
        // With A = P/R and E = [I_12; 0]:  (A^{-1} + E M E^T)^{-1} E = A E (I + M A_11)^{-1}, hence
        //   [K_h | K_x[:, 0:12]] = (P[:, 0:12] / R) (I + H^T H P_11 / R)^{-1} [H^T h | H^T H]
        // -- algebraically identical to esekfom.hpp:1782-1809 without the two 23x23 inversions.
        const int ld = 25;
        for (int e = tid; e < 12 * ld; e += nt) {
            const int r = e / ld, c2 = e % ld;
            double v;
            if (c2 < 12) {
                v = 0.0;
                for (int k = 0; k < 12; k++) v += S.HTH[r * 12 + k] * (S.P[k * n + c2] / R);
                v += (r == c2 ? 1.0 : 0.0);
            } else if (c2 == 12) v = S.Hth[r];
            else v = S.HTH[r * 12 + (c2 - 13)];
            S.aug[e] = v;
        }
        __syncthreads();
        ok = gj_eliminate(S.aug, 12, ld, ld, S.colbuf, S.row_of, S.used, &S.piv);
        if (ok) {
            for (int e = tid; e < 12 * 13; e += nt) {
                const int k = e / 13, j = e % 13;
                const int pr = S.row_of[k];
                S.T[e] = S.aug[pr * ld + 12 + j] / S.aug[pr * ld + k];      // column 0: for H^T h, 1..12: for H^T H
            }
            __syncthreads();
            for (int e = tid; e < n * 13; e += nt) {
                const int i = e / 13, j = e % 13;
                double v = 0.0;
                for (int a = 0; a < 12; a++) v += (S.P[i * n + a] / R) * S.T[a * 13 + j];
                if (j == 0) S.Kh[i] = v; else S.Kx[i * 12 + (j - 1)] = v;
            }
        }
    }
    __syncthreads();
    if (!ok) {
        if (tid == 0) { ctl->error = 1; ctl->done = 1; ctl->n_pass = n_pass + 1; }
        return;
    }
    // ------------------------------------------------------------------ esekfom.hpp:1815-1832
    if (tid < n) {
        double v = S.Kh[tid];
        for (int j = 0; j < 12; j++) v += S.Kx[tid * 12 + j] * S.dx_new[j];
        S.dxu[tid] = v - S.dx_new[tid];                       // K_h + (K_x - I) dx_new
    }
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < XLEN; i++) S.xnew[i] = ctl->x[i];
        state_boxplus(S.xnew, S.dxu);
        int converge = 1;
        for (int i = 0; i < n; i++) if (fabs(S.dxu[i]) > ctl->limit[i]) { converge = 0; break; }
        int t = t_in;
        if (converge) t++;
        if (!t && it == max_iter - 2) converge = 1;            // T2: force a re-search on the last pass
        const int finish = (t > 1 || it == max_iter - 1) ? 1 : 0;
        S.finish = finish;
        for (int i = 0; i < XLEN; i++) ctl->x[i] = S.xnew[i];
        ctl->t = t;
        ctl->converge = converge;
        ctl->n_pass = n_pass + 1;
        ctl->iter = it + 1;
        if (finish) ctl->done = 1;
        if (lg) { lg->converged = converge; for (int i = 0; i < XLEN; i++) lg->x_after[i] = S.xnew[i]; }
        if (finish) make_congruence(S.dxu, S.xnew, ctl->x_prop, S.J, S.M2);
    }
    __syncthreads();
    if (!S.finish) {
        // the reference leaves P_ = congruence-transformed P_propagated between passes
        for (int e = tid; e < n * n; e += nt) ctl->P[e] = S.P[e];
        return;
    }
    // ------------------------------------------------------------------ final covariance, esekfom.hpp:1834-1927
    for (int e = tid; e < n * n; e += nt) S.L[e] = S.P[e];
    __syncthreads();
    apply_rows(S.L, S.P, S.J[0], S.J[1], S.M2, n, n);          // L rows from P rows
    __syncthreads();
    apply_rows(S.Kx, S.Kx, S.J[0], S.J[1], S.M2, 12, 12);      // K_x rows, first 12 columns
    apply_cols(S.L, S.J[0], S.J[1], S.M2);
    __syncthreads();
    apply_cols(S.P, S.J[0], S.J[1], S.M2);
    __syncthreads();
    for (int e = tid; e < n * n; e += nt) {                     // P_ = L_ - K_x[:, 0:12] P_[0:12, :]
        const int i = e / n, j = e % n;
        double v = 0.0;
        for (int a = 0; a < 12; a++) v += S.Kx[i * 12 + a] * S.P[a * n + j];
        ctl->P[e] = S.L[e] - v;
    }
}

__global__ void k_init_ctl(FilterCtl* ctl) {
    const int tid = threadIdx.x;
    for (int e = tid; e < NDOF * NDOF; e += blockDim.x) ctl->P_prop[e] = ctl->P[e];
    if (tid < XLEN) ctl->x_prop[tid] = ctl->x[tid];
    if (tid == 0) { ctl->iter = -1; ctl->t = 0; ctl->converge = 1; ctl->done = 0; ctl->n_pass = 0; ctl->error = 0; }
}

// ============================================================================= NCCL (lazy)
struct NcclUniqueId { char internal[128]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi* load_nccl() {
    static NcclApi api;
    if (api.lib) return &api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) { set_last_error("NCCL: cannot dlopen libnccl.so.2: %s", dlerror()); return nullptr; }
    api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void**, int, NcclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
    api.CommDestroy = (int (*)(void*))dlsym(api.lib, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
        set_last_error("NCCL: missing symbols in libnccl");
        api.lib = nullptr;
        return nullptr;
    }
    return &api;
}
int nccl_unique_id(void* out128) {
    NcclApi* api = load_nccl();
    if (!api) return FL_ERR_NCCL;
    NcclUniqueId id;
    int rc = api->GetUniqueId(&id);
    if (rc != 0) { set_last_error("ncclGetUniqueId failed: %d", rc); return FL_ERR_NCCL; }
    memcpy(out128, &id, 128);
    return FL_OK;
}

// ============================================================================= Filter (host)
Filter::Filter(Map* map, int max_points) : map_(map), max_points_(max_points) {
    memset(&scan_, 0, sizeof(scan_));
    for (int i = 0; i < NDOF; i++) limit_[i] = 0.001;      // epsi, laserMapping.cpp:826-827
}
Filter::~Filter() {
    cudaSetDevice(map_->device());
    if (comm_ && nccl_) nccl_->CommDestroy(comm_);
    body_.release(); nearest_.release(); nearest_cnt_.release(); selected_.release(); normvec_.release();
    partials_.release(); red_.release(); ctl_.release(); logs_.release();
    if (h_ctl_) cudaFreeHost(h_ctl_);
}

int Filter::init() {
    FL_CUDA(cudaSetDevice(map_->device()));
    FL_CHECK(ctl_.reserve(sizeof(FilterCtl)));
    FL_CHECK(red_.reserve(sizeof(double) * PSTRIDE));
    FL_CHECK(logs_.reserve(sizeof(PassLog) * MAX_LOGS));
    FL_CUDA(cudaMallocHost(&h_ctl_, sizeof(FilterCtl)));
    memset(h_ctl_, 0, sizeof(FilterCtl));
    FL_CUDA(cudaMemsetAsync(ctl_.ptr, 0, sizeof(FilterCtl), stream()));
    FL_CUDA(cudaMemsetAsync(logs_.ptr, 0, sizeof(PassLog) * MAX_LOGS, stream()));
    // persistent grid: as many blocks as can be co-resident
    int dev = map_->device(), sms = 0, occ_a = 0, occ_b = 0;
    FL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, k_measure<false>, MEASURE_THREADS, 0));
    FL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, k_measure<true>, MEASURE_THREADS, 0));
    grid_ = sms * std::max(1, std::min(occ_a, occ_b));
    FL_CHECK(partials_.reserve(sizeof(double) * PSTRIDE * (size_t)grid_));
    FL_CHECK(reserve(std::max(1, max_points_)));
    return FL_OK;
}

int Filter::reserve(int nq) {
    FL_CHECK(body_.reserve(sizeof(float4) * (size_t)nq));
    FL_CHECK(nearest_.reserve(sizeof(float4) * KNN_K * (size_t)nq));
    FL_CHECK(nearest_cnt_.reserve(sizeof(int) * (size_t)nq));
    FL_CHECK(selected_.reserve((size_t)nq));
    FL_CHECK(normvec_.reserve(sizeof(float4) * (size_t)nq));
    scan_.nearest = nearest_.as<float4>();
    scan_.nearest_cnt = nearest_cnt_.as<int>();
    scan_.selected = selected_.as<unsigned char>();
    scan_.normvec = normvec_.as<float4>();
    return FL_OK;
}

int Filter::set_params(int max_iter, const double* limit23, int extrinsic_est_en) {
    if (max_iter < 1 || max_iter + 1 > MAX_LOGS) { set_last_error("set_params: max_iter must be in [1, %d]", MAX_LOGS - 1); return FL_ERR_ARG; }
    max_iter_ = max_iter;
    if (limit23) for (int i = 0; i < NDOF; i++) limit_[i] = limit23[i];
    extrinsic_est_ = extrinsic_est_en ? 1 : 0;
    return FL_OK;
}

int Filter::set_scan_device(const float4* d_body, int nq) {
    if (nq < 0) { set_last_error("scan: nq < 0"); return FL_ERR_ARG; }
    FL_CUDA(cudaSetDevice(map_->device()));
    FL_CHECK(reserve(std::max(1, nq)));
    scan_.body = d_body;
    scan_.Q = nq;
    if (!shard_set_) { scan_.q_begin = 0; scan_.q_end = nq; }
    // per-scan state of the reference's globals: point_selected_surf is rewritten for every point
    // on the first (always searching) pass, Nearest_Points likewise
    return FL_OK;
}

int Filter::upload_scan(const float* body_xyzi, int nq) {
    if (nq < 0 || (nq > 0 && !body_xyzi)) { set_last_error("scan: bad arguments"); return FL_ERR_ARG; }
    FL_CUDA(cudaSetDevice(map_->device()));
    FL_CHECK(reserve(std::max(1, nq)));
    if (nq > 0) FL_CUDA(cudaMemcpyAsync(body_.ptr, body_xyzi, sizeof(float4) * (size_t)nq, cudaMemcpyHostToDevice, stream()));
    return set_scan_device(body_.as<float4>(), nq);
}

int Filter::set_shard(int q_begin, int q_end) {
    if (q_begin < 0 || q_end < q_begin) { set_last_error("set_shard: bad range"); return FL_ERR_ARG; }
    scan_.q_begin = q_begin; scan_.q_end = q_end;
    shard_set_ = true;
    return FL_OK;
}

int Filter::upload_state(const double* x26, const double* P, double R) {
    FL_CUDA(cudaSetDevice(map_->device()));
    memcpy(h_ctl_->x, x26, sizeof(double) * XLEN);
    memcpy(h_ctl_->P, P, sizeof(double) * NDOF * NDOF);
    for (int i = 0; i < NDOF; i++) h_ctl_->limit[i] = limit_[i];
    h_ctl_->R = R;
    h_ctl_->max_iter = max_iter_;
    h_ctl_->extrinsic_est = extrinsic_est_;
    // x, x_prop, P are contiguous with the header: copy [0, offsetof(P_prop))
    FL_CUDA(cudaMemcpyAsync(ctl_.ptr, h_ctl_, offsetof(FilterCtl, P_prop), cudaMemcpyHostToDevice, stream()));
    return FL_OK;
}

int Filter::run_passes() {
    FL_CUDA(cudaSetDevice(map_->device()));
    if (scan_.q_end > scan_.Q) { set_last_error("shard exceeds the scan"); return FL_ERR_ARG; }
    FilterCtl* ctl = ctl_.as<FilterCtl>();
    cudaStream_t st = stream();
    k_init_ctl<<<1, 256, 0, st>>>(ctl);
    launches_ = 1;
    const MapView& mv = map_->view();
    for (int pass = 0; pass <= max_iter_; pass++) {
        if (extrinsic_est_) k_measure<true><<<grid_, MEASURE_THREADS, 0, st>>>(mv, scan_, ctl, partials_.as<double>());
        else k_measure<false><<<grid_, MEASURE_THREADS, 0, st>>>(mv, scan_, ctl, partials_.as<double>());
        launches_++;
        if (nranks_ > 1) {
            k_solve<<<1, SOLVE_THREADS, 0, st>>>(ctl, partials_.as<double>(), grid_, red_.as<double>(), 1, scan_, logs_.as<PassLog>(), solver_);
            int rc = nccl_->AllReduce(red_.ptr, red_.ptr, NRED, /*ncclDouble*/ 8, /*ncclSum*/ 0, comm_, st);
            if (rc != 0) { set_last_error("ncclAllReduce failed: %d", rc); return FL_ERR_NCCL; }
            k_solve<<<1, SOLVE_THREADS, 0, st>>>(ctl, partials_.as<double>(), grid_, red_.as<double>(), 2, scan_, logs_.as<PassLog>(), solver_);
            launches_ += 2;
        } else {
            k_solve<<<1, SOLVE_THREADS, 0, st>>>(ctl, partials_.as<double>(), grid_, red_.as<double>(), 0, scan_, logs_.as<PassLog>(), solver_);
            launches_++;
        }
    }
    FL_CUDA(cudaGetLastError());
    return FL_OK;
}

int Filter::sync() {
    FL_CUDA(cudaSetDevice(map_->device()));
    FL_CUDA(cudaStreamSynchronize(stream()));
    return FL_OK;
}

int Filter::download_state(double* x26, double* P, int* n_pass) {
    FL_CUDA(cudaSetDevice(map_->device()));
    FL_CUDA(cudaMemcpyAsync(h_ctl_, ctl_.ptr, offsetof(FilterCtl, P_prop), cudaMemcpyDeviceToHost, stream()));
    FL_CUDA(cudaStreamSynchronize(stream()));
    if (h_ctl_->error) { set_last_error("update: singular system on device"); return FL_ERR_STATE; }
    if (x26) memcpy(x26, h_ctl_->x, sizeof(double) * XLEN);
    if (P) memcpy(P, h_ctl_->P, sizeof(double) * NDOF * NDOF);
    if (n_pass) *n_pass = h_ctl_->n_pass;
    return FL_OK;
}

int Filter::update(const float* body_xyzi, int nq, double* x26, double* P, double R, double* solve_time_s) {
    if (!x26 || !P) { set_last_error("update: null state"); return FL_ERR_ARG; }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    FL_CUDA(cudaSetDevice(map_->device()));
    if (solve_time_s) { FL_CUDA(cudaEventCreate(&e0)); FL_CUDA(cudaEventCreate(&e1)); FL_CUDA(cudaEventRecord(e0, stream())); }
    FL_CHECK(upload_scan(body_xyzi, nq));
    FL_CHECK(upload_state(x26, P, R));
    FL_CHECK(run_passes());
    if (solve_time_s) FL_CUDA(cudaEventRecord(e1, stream()));
    FL_CHECK(download_state(x26, P, nullptr));
    if (solve_time_s) {
        float ms = 0.f;
        FL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        *solve_time_s += ms * 1e-3;                       // the reference accumulates into solve_time (esekfom.hpp:1926)
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    return FL_OK;
}

int Filter::get_nearest(float* out_pts, int* out_cnt, int nq) {
    if (nq > scan_.Q) { set_last_error("get_nearest: nq exceeds the bound scan"); return FL_ERR_ARG; }
    FL_CUDA(cudaSetDevice(map_->device()));
    if (out_pts) FL_CUDA(cudaMemcpyAsync(out_pts, scan_.nearest, sizeof(float4) * KNN_K * (size_t)nq, cudaMemcpyDeviceToHost, stream()));
    if (out_cnt) FL_CUDA(cudaMemcpyAsync(out_cnt, scan_.nearest_cnt, sizeof(int) * (size_t)nq, cudaMemcpyDeviceToHost, stream()));
    FL_CUDA(cudaStreamSynchronize(stream()));
    return FL_OK;
}
int Filter::get_selected(unsigned char* out, int nq) {
    if (nq > scan_.Q) { set_last_error("get_selected: nq exceeds the bound scan"); return FL_ERR_ARG; }
    FL_CUDA(cudaSetDevice(map_->device()));
    FL_CUDA(cudaMemcpyAsync(out, scan_.selected, (size_t)nq, cudaMemcpyDeviceToHost, stream()));
    FL_CUDA(cudaStreamSynchronize(stream()));
    return FL_OK;
}
int Filter::get_pass_logs(PassLog* out, int cap, int* n) {
    FL_CUDA(cudaSetDevice(map_->device()));
    FL_CUDA(cudaMemcpyAsync(h_ctl_, ctl_.ptr, offsetof(FilterCtl, P_prop), cudaMemcpyDeviceToHost, stream()));
    FL_CUDA(cudaStreamSynchronize(stream()));
    const int np = std::min(std::min(h_ctl_->n_pass, MAX_LOGS), cap);
    if (np > 0) {
        FL_CUDA(cudaMemcpyAsync(out, logs_.ptr, sizeof(PassLog) * (size_t)np, cudaMemcpyDeviceToHost, stream()));
        FL_CUDA(cudaStreamSynchronize(stream()));
    }
    if (n) *n = np;
    return FL_OK;
}

int Filter::comm_init(int nranks, int rank, const void* id128) {
    if (nranks < 1 || rank < 0 || rank >= nranks) { set_last_error("comm_init: bad rank/size"); return FL_ERR_ARG; }
    nranks_ = nranks; rank_ = rank;
    if (nranks == 1) return FL_OK;
    FL_CUDA(cudaSetDevice(map_->device()));
    nccl_ = load_nccl();
    if (!nccl_) return FL_ERR_NCCL;
    NcclUniqueId id;
    memcpy(&id, id128, 128);
    int rc = nccl_->CommInitRank(&comm_, nranks, id, rank);
    if (rc != 0) { set_last_error("ncclCommInitRank failed: %d (%s)", rc, nccl_->GetErrorString ? nccl_->GetErrorString(rc) : "?"); return FL_ERR_NCCL; }
    return FL_OK;
}

}  // namespace fl
