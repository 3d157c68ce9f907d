// k_update -- the whole iterated-EKF measurement update of one scan in ONE persistent kernel launch.
//
//   worker blocks (1 .. gridDim.x-1), one lane per scan point, every pass:
//       h_share_model (reference src/laserMapping.cpp:638-754) fused end to end -- body->world transform, k = 5 nearest
//       neighbours (cell directory + BVH walk, map.cuh) on the passes that search, 5-point plane fit straight from the
//       registers that hold the neighbours (esti_plane, include/common_lib.h:225-257), residual gating, Jacobian row --
//       folded into the FP64 normal equations H^T H / H^T h with one deterministic partial per block.  The rows never
//       reach memory; the neighbours are written once (map_incremental reads them, laserMapping.cpp:438-460).
//   solver block (0):
//       update_iterated_dyn_share_modified (reference include/IKFoM_toolkit/esekfom/esekfom.hpp:1619-1931) for the same
//       pass: while the workers measure it prepares everything that depends only on the state (x [-] x_prop, the manifold
//       congruence T P_prop T^T, :1651-1699); when their tickets are in it reduces the partials in a fixed order, forms
//       the gain in one warp with the system in registers, applies [+], decides convergence and PUBLISHES the new pose
//       (release store of a generation counter) -- the workers of the next pass spin on that counter, there is no
//       kernel boundary between passes.  Covariance bookkeeping and the pass log happen after the publication, off the
//       critical path.
//
// Gain.  The reference's information form (:1782-1809)   K_h = (H^T H + (P/R)^-1)^-1 H^T h,  K_x = (...)^-1 H^T H
// is evaluated through the matrix-inversion lemma on the only block H touches (ne = 6 columns, 12 with extrinsic
// estimation):   [K_h | K_x[:, :ne]] = (P[:, :ne] / R) (I + H^T H P_11 / R)^-1 [H^T h | H^T H]   -- no 23x23 inverse.
// Only  dx_ = K_h + (K_x - I) dx_new  is needed to advance the state, i.e. ONE extra right-hand side
//   v = (I + H^T H P_11 / R)^-1 (H^T h + H^T H dx_new[:ne]),   dx_ = (P[:, :ne] / R) v - dx_new,
// and K_x itself only for the covariance of the pass that ends the update (:1834-1927), which collapses to
//   P_final = T2 (P - (P[:, :ne] / R) W P[:ne, :]) T2^T,   W = (I + H^T H P_11 / R)^-1 H^T H,   T2 = congruence at dx_.
// The small-m branch of the reference (:1715-1744, fewer than 23 rows) is the same gain by the same lemma; this kernel
// uses the one form for every m >= 1 (which also makes the multi-GPU solve independent of how the rows are sharded).
// fl_filter_set_solver(0) runs the reference's two formulas literally through the legacy kernels (validation).
#pragma once

namespace fl {

constexpr int UPD_THREADS = 256;
constexpr int UPD_WARPS = UPD_THREADS / 32;

struct UpdArgs {
    MapView m;
    ScanView sc;
    FilterCtl* ctl;
    double* partials;      // [worker blocks][PSTRIDE]
    double* red_g;         // [PSTRIDE] sums of this rank (mode 1 out, mode 3 in)
    PassLog* logs;
    P2PState* p2p;
    unsigned long long* pub;   // publication block (32 words, own 256-byte line pair): the pose of the next pass as tagged words
    unsigned nonce;        // unique per launch: stale words of an earlier launch can never look current
    int mode;              // 0: single GPU; 1: workers + reduction only (ncclAllReduce follows); 2: peer-memory exchange; 3: solver only, sums from red_g
    int max_passes;        // passes this launch may run (persistent: max_iter + 1; NCCL chain: 1)
    int search_only;       // 1: the kNN phase of one searching pass alone (neighbours + gate), for timing
    int dbg;               // tuning switches (FASTLIO_B200_DBG)
    int pose_from_search;  // search_only: transform with ctl->x_search (the state of the last searching pass) instead of ctl->x
};

struct SolverSm {
    double Pp[NDOF * NDOF];        // P_propagated (fixed for the update)
    double Pt[NDOF * NDOF];        // T P_prop T^T of the current pass   (esekfom.hpp:1657-1699)
    double W1[NDOF * NDOF];        // scratch
    double Y[12 * NDOF];           // W P[:ne, :]
    double Wm[12 * 13];            // row k: [v_k | W_k,0..ne-1]
    double HTH[144];
    double wred[UPD_WARPS][PSTRIDE];
    double red[PSTRIDE];
    double x[XLEN], xprop[XLEN], xnew[XLEN];
    double dx[NDOF], dxn[NDOF], dxu[NDOF], limit[NDOF];
    double J[2][9], M2[4];
    double Bprop[6];               // S2_Bx(x_propagated.grav): fixed for the update
    double R;
    int iter, t, converge, done, n_pass, max_iter, error;
    int effct, ok, finish, searched, late;
    int row_k[12];                 // pivot row -> elimination step
};
template <bool EXTR> struct WorkerSm {
    static constexpr int STAGE = 32 * RowStage<EXTR>::RS + 96;
    double stage[UPD_WARPS][STAGE];
    double wred[UPD_WARPS][PSTRIDE];
    WalkPool walks;
    unsigned pose_bits[28];
    unsigned flags;
    int abort;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
constexpr long long SPIN_LIMIT = 3000000000ll;       // ~1.5 s of SM clocks: a wedged peer must not hang the GPU

// Publication of a pass's result to the worker blocks WITHOUT a fence: every 8-byte word carries 32 bits of payload and a
// 32-bit tag (launch nonce, pass number) -- the data is the flag.  Words 0..27: the 14 doubles of the pose h_share_model
// reads (pos, rot, offset_R_L_I, offset_T_L_I = x[0..13]) as halves; word 28: bit 0 converge, bit 1 done.
constexpr int PUB_WORDS = 29;
__device__ __forceinline__ unsigned pub_tag(unsigned nonce, int pass) { return (nonce << 8) | (unsigned)(pass & 0xff); }
__device__ __forceinline__ void pub_store(unsigned long long* p, unsigned tag, unsigned payload) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(((unsigned long long)tag << 32) | payload) : "memory");
}
__device__ __forceinline__ unsigned long long pub_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// warp 0 of the solver block: x[0..13] and the flags, tagged for `pass`
__device__ __forceinline__ void pub_publish(unsigned long long* pub, unsigned tag, const double* x, int converge, int done, int lane) {
    if (lane < 28) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(x[lane >> 1]);
        pub_store(pub + lane, tag, (lane & 1) ? (unsigned)(bits >> 32) : (unsigned)bits);
    } else if (lane == 28) {
        pub_store(pub + 28, tag, (converge ? 1u : 0u) | (done ? 2u : 0u));
    }
}

// ============================================================================= workers
// Everything of h_share_model for one scan point (laserMapping.cpp:650-692), one thread per point.  Block-collective on a
// searching pass (the queries the directory cannot prove are walked through the BVH by all warps of the block, knn_block).  On a searching pass the five neighbours go from the search's registers
// straight into the plane fit; they are stored once (map_incremental reads them, laserMapping.cpp:438-460).  Returns true when
// the point contributes a row.
template <bool EXTR>
__device__ __forceinline__ bool measure_fused(const MapView& m, const ScanView& sc, int q, bool active, const PoseS& s, bool searched,
                                              bool search_only, WalkPool& walks, int& phase, double* h, double& z, float& absres) {
    float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
    float wx = 0.f, wy = 0.f, wz = 0.f;
    D3 p_this = d3(0.0, 0.0, 0.0);
    if (active) {
        pb = __ldg(&sc.body[q]);
        p_this = qrot(s.offR, d3(pb.x, pb.y, pb.z)) + s.offT;               // :656-661 (also the lever arm of the Jacobian, :733)
        const D3 g = qrot(s.rot, p_this) + s.pos;
        wx = (float)g.x; wy = (float)g.y; wz = (float)g.z;
    }
    bool sel = false;
    float pabcd[4] = {0.f, 0.f, 0.f, 0.f};
    if (searched) {                                                         // :667
        TBest kb;
        knn_block(m, active, wx, wy, wz, kb, walks, phase);                        // :670 (block-wide: two barriers)
        if (active) {
            float4 p[KNN_K];
            const int cnt = knn_fetch(m, kb, p);
#pragma unroll
            for (int j = 0; j < KNN_K; j++) sc.nearest[(size_t)q * KNN_K + j] = p[j];
            sc.nearest_cnt[q] = cnt;
            sel = cnt >= KNN_K && !(kb.d[KNN_K - 1] > 5.0f);                // :671
            if (search_only) { sc.selected[q] = sel ? 1 : 0; return false; }
            if (sel) {
                float pn[KNN_K][3];
#pragma unroll
                for (int j = 0; j < KNN_K; j++) { pn[j][0] = p[j].x; pn[j][1] = p[j].y; pn[j][2] = p[j].z; }
                sel = esti_plane_dev(pabcd, pn, 0.1f);                      // :678
                if (sel) sc.plane[q] = make_float4(pabcd[0], pabcd[1], pabcd[2], pabcd[3]);
            }
        }
    } else if (active) {
        // a pass that does not search fits the plane to the SAME five neighbours (Nearest_Points persists, T3) and only
        // points whose fit and score succeeded last time are still selected: the fit is reused, not recomputed
        sel = sc.selected[q] != 0;                                          // :674
        if (sel) { const float4 pl = sc.plane[q]; pabcd[0] = pl.x; pabcd[1] = pl.y; pabcd[2] = pl.z; pabcd[3] = pl.w; }
    }
    bool contrib = false;
    if (active && sel) {
        const float pd2 = pabcd[0] * wx + pabcd[1] * wy + pabcd[2] * wz + pabcd[3];             // :680
        // sqrt(p_body.norm()) depends on the point alone: computed on the searching passes, cached for the others
        double srange;
        if (searched) { srange = sqrt(norm3(d3(pb.x, pb.y, pb.z))); sc.srange[q] = srange; }
        else srange = sc.srange[q];
        const float score = (float)(1 - 0.9 * fabs((double)pd2) / srange);                     // :681 (T8)
        if ((double)score > 0.9) {                                                              // :683
            contrib = true;
            absres = fabsf(pd2);                                                                // res_last
            jacobian_row_at<EXTR>(s, pb, p_this, make_float4(pabcd[0], pabcd[1], pabcd[2], pd2), h, z);    // :723-751
        }
    }
    if (active) sc.selected[q] = contrib ? 1 : 0;                                               // :677, :685
    return contrib;
}

// ============================================================================= solver
// once per launch: the control block into shared memory
__device__ void sol_load(SolverSm& S, const FilterCtl* ctl) {
    const int tid = threadIdx.x;
    for (int e = tid; e < NDOF * NDOF; e += UPD_THREADS) S.Pp[e] = __ldcg(&ctl->P_prop[e]);
    if (tid < XLEN) { S.x[tid] = __ldcg(&ctl->x[tid]); S.xprop[tid] = __ldcg(&ctl->x_prop[tid]); }
    if (tid >= 32 && tid < 32 + NDOF) S.limit[tid - 32] = __ldcg(&ctl->limit[tid - 32]);
    if (tid == 64) {
        S.R = __ldcg(&ctl->R);
        S.iter = __ldcg(&ctl->iter); S.t = __ldcg(&ctl->t); S.converge = __ldcg(&ctl->converge); S.done = __ldcg(&ctl->done);
        S.n_pass = __ldcg(&ctl->n_pass); S.max_iter = __ldcg(&ctl->max_iter); S.error = 0; S.late = 0;
    }
    __syncthreads();
    if (tid == 0) S2_Bx(ld3(S.xprop + X_GRAV), S.Bprop);
    __syncthreads();
}

// the state-only half of a pass (esekfom.hpp:1651-1699): dx = x [-] x_prop, dx_new, the congruence blocks, Pt = T P_prop T^T
__device__ void sol_prepare(SolverSm& S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int n = NDOF;
    if (lane == 0) {
        if (warp < 2) {                 // SO3 blocks: rot (idx 3), offset_R_L_I (idx 6)
            const int idx = warp == 0 ? 3 : 6, xo = warp == 0 ? X_ROT : X_OFFR;
            const D3 l = so3_log(qmul(qconj(ldq(S.xprop + xo)), ldq(S.x + xo)));                    // SOn.hpp:237-239
            const M33 J = transpose33(A_matrix(l));                                                 // T5
#pragma unroll 1
            for (int i = 0; i < 9; i++) S.J[warp][i] = J.m[i];
            const D3 seg = mul33v(J, l);
            S.dx[idx] = l.x; S.dx[idx + 1] = l.y; S.dx[idx + 2] = l.z;
            S.dxn[idx] = seg.x; S.dxn[idx + 1] = seg.y; S.dxn[idx + 2] = seg.z;
        } else if (warp == 2) {         // S2 block: grav (idx 21)
            double d0, d1;
            S2_boxminus_B(ld3(S.x + X_GRAV), ld3(S.xprop + X_GRAV), S.Bprop, d0, d1);
            S2_congruence_B(ld3(S.x + X_GRAV), ld3(S.xprop + X_GRAV), S.Bprop, d0, d1, S.M2);
            S.dx[21] = d0; S.dx[22] = d1;
            S.dxn[21] = S.M2[0] * d0 + S.M2[1] * d1;
            S.dxn[22] = S.M2[2] * d0 + S.M2[3] * d1;
        }
    }
    if (warp == 3 && lane < 15) {       // vect blocks: pos, offset_T_L_I, vel, bg, ba
        const int b = lane / 3, c = lane % 3;
        const int dof = b == 0 ? 0 : 9 + 3 * (b - 1);
        const int xo = b == 0 ? X_POS : (b == 1 ? X_OFFT : (b == 2 ? X_VEL : (b == 3 ? X_BG : X_BA)));
        const double d = S.x[xo + c] - S.xprop[xo + c];
        S.dx[dof + c] = d; S.dxn[dof + c] = d;
    }
    __syncthreads();
    // W1 = P_prop T^T  (columns of the SO3 / S2 blocks), then Pt = T W1 (their rows): T P T^T as :1659-1699 apply it block by block
#pragma unroll 1
    for (int e = tid; e < n * n; e += UPD_THREADS) {
        const int i = e / n, j = e - i * n;
        const double* row = &S.Pp[i * n];
        double v;
        if (j >= 3 && j < 9) {
            const int b = j >= 6, base = 3 + 3 * b, r = j - base;
            v = S.J[b][3 * r] * row[base] + S.J[b][3 * r + 1] * row[base + 1] + S.J[b][3 * r + 2] * row[base + 2];
        } else if (j >= 21) {
            const int r = j - 21;
            v = S.M2[2 * r] * row[21] + S.M2[2 * r + 1] * row[22];
        } else v = row[j];
        S.W1[e] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int e = tid; e < n * n; e += UPD_THREADS) {
        const int i = e / n, j = e - i * n;
        double v;
        if (i >= 3 && i < 9) {
            const int b = i >= 6, base = 3 + 3 * b, r = i - base;
            v = S.J[b][3 * r] * S.W1[base * n + j] + S.J[b][3 * r + 1] * S.W1[(base + 1) * n + j] + S.J[b][3 * r + 2] * S.W1[(base + 2) * n + j];
        } else if (i >= 21) {
            const int r = i - 21;
            v = S.M2[2 * r] * S.W1[21 * n + j] + S.M2[2 * r + 1] * S.W1[22 * n + j];
        } else v = S.W1[e];
        S.Pt[e] = v;
    }
    __syncthreads();
}

// fixed-order reduction of the workers' block partials into S.red
__device__ void sol_reduce(SolverSm& S, const double* __restrict__ partials, int nwork) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int b = warp; b < nwork; b += 8 * UPD_WARPS) {          // eight rows per step: 24 loads in flight, then summed in row order
        double v[8][3];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int bb = b + k * UPD_WARPS;
            const double* row = partials + (size_t)bb * PSTRIDE;
            v[k][0] = v[k][1] = v[k][2] = 0.0;
            if (bb < nwork) { v[k][0] = __ldcg(&row[lane]); v[k][1] = __ldcg(&row[lane + 32]); v[k][2] = __ldcg(&row[lane + 64]); }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) { a0 += v[k][0]; a1 += v[k][1]; a2 += v[k][2]; }
    }
    S.wred[warp][lane] = a0; S.wred[warp][lane + 32] = a1; S.wred[warp][lane + 64] = a2;
    __syncthreads();
    if (tid < PSTRIDE) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < UPD_WARPS; w++) v += S.wred[w][tid];
        S.red[tid] = v;
        if (tid < 78) {                       // H^T H in full 12 x 12 form for the gain (tid -> (a <= b) of the packed upper triangle)
            int a = 0, rem = tid;
            while (rem >= 12 - a) { rem -= 12 - a; a++; }
            const int b = a + rem;
            S.HTH[a * 12 + b] = v; S.HTH[b * 12 + a] = v;
        }
    }
    __syncthreads();
}
// the same expansion for sums that did not come through sol_reduce (peer exchange, NCCL)
__device__ void sol_expand(SolverSm& S) {
    const int tid = threadIdx.x;
    if (tid < 144) { const int a = tid / 12, b = tid - a * 12; S.HTH[tid] = S.red[a <= b ? tri12(a, b) : tri12(b, a)]; }
    __syncthreads();
}

// all-reduce of S.red over the peer mailboxes (see k_residual / DESIGN.md section 5): every rank stores its sums into its slot
// of every rank's mailbox as epoch-tagged words and adds the slots of its own mailbox in rank order
__device__ void sol_exchange(SolverSm& S, P2PState* p2p) {
    const int nr = p2p->nranks, me = p2p->rank;
    const unsigned long long epoch = p2p->epoch + 1;
    const unsigned long long tag = (epoch & 0xffffffffull) << 32;
    const int par = (int)(epoch & 1ull);
    for (int idx = threadIdx.x; idx < nr * PSTRIDE; idx += UPD_THREADS) {
        const int r = idx / PSTRIDE, o = idx - r * PSTRIDE;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(S.red[o]);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(p2p->peer_mail[r]) + (((size_t)par * nr + me) * PSTRIDE + o) * 2;
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(tag | (bits & 0xffffffffull)) : "memory");
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(tag | (bits >> 32)) : "memory");
    }
    __syncthreads();                                  // S.red has been sent before it is overwritten with the sum
    if (threadIdx.x < PSTRIDE) {
        const unsigned long long* mail = reinterpret_cast<const unsigned long long*>(p2p->peer_mail[me]) + (size_t)par * nr * PSTRIDE * 2;
        double v = 0.0;
        bool late = false;
        for (int r = 0; r < nr && !late; r++) {
            const unsigned long long* src = mail + ((size_t)r * PSTRIDE + threadIdx.x) * 2;
            unsigned long long lo = 0, hi = 0;
            const long long t0 = clock64();
            while (true) {
                asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(lo) : "l"(src) : "memory");
                asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(hi) : "l"(src + 1) : "memory");
                if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
                if (clock64() - t0 > SPIN_LIMIT) { late = true; break; }
            }
            v += __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
        }
        S.red[threadIdx.x] = v;
        if (late) S.late = 1;
    }
    if (threadIdx.x == 0) p2p->epoch = epoch;
    __syncthreads();
}

// Gauss-Jordan with partial pivoting on [A | B], ne rows, one COLUMN per lane, rows in registers.  The pivot row is scaled to a
// unit pivot as it is chosen, so after the last step the solution row of unknown k is physical row `r` with row_k[r] == k.
template <int N>
__device__ __forceinline__ bool gj_cols(double (&c)[N], int* row_k, int lane) {
    unsigned used = 0;
    bool ok = true;
#pragma unroll 1
    for (int k = 0; k < N; k++) {
        int p = 0; double best = -1.0;
#pragma unroll
        for (int r = 0; r < N; r++) {
            const double v = fabs(c[r]);
            const bool cand = !((used >> r) & 1u) && v > best;
            best = cand ? v : best; p = cand ? r : p;
        }
        p = __shfl_sync(FULL, p, k);
        best = __shfl_sync(FULL, best, k);
        if (!(best > 0.0)) ok = false;
        double apj = 0.0;
#pragma unroll
        for (int r = 0; r < N; r++) apj = (r == p) ? c[r] : apj;
        const double inv = 1.0 / __shfl_sync(FULL, apj, k);
        apj *= inv;                                        // the scaled pivot row, this lane's column
#pragma unroll
        for (int r = 0; r < N; r++) {
            const double ck = __shfl_sync(FULL, c[r], k);
            if (r == p) c[r] = apj;
            else if (lane > k) c[r] -= ck * apj;
        }
        used |= 1u << p;
        if (lane == 0) row_k[p] = k;
    }
    __syncwarp();
    return ok;
}

// The H-dependent half of a pass (esekfom.hpp:1782-1834).  The critical path -- gain, dx_, convergence, [+] on the pose the
// measurement model reads, publication -- runs in WARP 0 ALONE, with the system in registers and no block barrier; everything
// else ([+] on velocity / biases / gravity, the log, the covariance bookkeeping and, on the pass that ends the update, the final
// covariance) follows the publication.
template <bool EXTR>
__device__ void sol_pass(SolverSm& S, FilterCtl* ctl, PassLog* logs, unsigned long long* pub, unsigned tag_next) {
    constexpr int NE = EXTR ? 12 : 6;
    constexpr int n = NDOF;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double Rinv = 1.0 / S.R;
    if (warp == 0) {
        const int effct = (int)(S.red[90] + 0.5);
        bool ok = true;
        int finish = 0;
        if (effct >= 1 && !S.late) {
            // lane j <= 2 NE holds column j of [A | rhs | H^T H] as  base + H^T H u:
            //   j < NE   : A = I + H^T H P_11 / R          u = P_11[:, j] / R       base = e_j
            //   j == NE  : H^T h + H^T H dx_new[:ne]       u = dx_new[:ne]          base = H^T h
            //   j > NE   : H^T H                           u = e_(j - NE - 1)       base = 0
            double c[NE], u[NE];
#pragma unroll
            for (int k = 0; k < NE; k++) {
                const double pk = S.Pt[k * n + (lane < NE ? lane : 0)] * Rinv;
                u[k] = lane < NE ? pk : (lane == NE ? S.dxn[k] : (k == lane - NE - 1 ? 1.0 : 0.0));
            }
#pragma unroll
            for (int r = 0; r < NE; r++) {
                double v = lane < NE ? (r == lane ? 1.0 : 0.0) : (lane == NE ? S.red[78 + r] : 0.0);
#pragma unroll
                for (int k = 0; k < NE; k++) v = fma(S.HTH[r * 12 + k], u[k], v);
                c[r] = lane <= 2 * NE ? v : 0.0;
            }
            ok = gj_cols<NE>(c, S.row_k, lane);
            if (lane == 0) ctl->prof[4] = clock64();
            if (lane >= NE && lane <= 2 * NE) {
#pragma unroll
                for (int r = 0; r < NE; r++) S.Wm[S.row_k[r] * 13 + (lane - NE)] = c[r];
            }
            __syncwarp();
            // dx_ = K_h + (K_x - I) dx_new = (P[:, :ne] / R) v - dx_new                                  (:1815)
            double d = 0.0;
            if (lane < n) {
#pragma unroll
                for (int a = 0; a < NE; a++) d = fma(S.Pt[lane * n + a] * Rinv, S.Wm[a * 13], d);
                d -= S.dxn[lane];
                S.dxu[lane] = d;
            }
            const unsigned over = __ballot_sync(FULL, lane < n && fabs(d) > S.limit[lane]);              // :1818-1825
            int converge = over ? 0 : 1;
            int t = S.t;
            if (converge) t++;
            if (!t && S.iter == S.max_iter - 2) converge = 1;               // T2: force a re-search on the last pass (:1829-1832)
            finish = (t > 1 || S.iter == S.max_iter - 1) ? 1 : 0;           // :1834
            __syncwarp();
            if (lane == 0) ctl->prof[6] = clock64();
            if (ok) {
                // x_.boxplus(dx_) on the pose (:1817): lanes 0 / 1 the two rotations (one instruction stream), lanes 2..7 pos, offset_T
                if (lane < 2) {
                    const int idx = lane == 0 ? 3 : 6, xo = lane == 0 ? X_ROT : X_OFFR;
                    stq(S.xnew + xo, qmul(ldq(S.x + xo), so3_exp(d3(S.dxu[idx], S.dxu[idx + 1], S.dxu[idx + 2]))));   // SOn.hpp:233-236
                } else if (lane < 8) {
                    const int b = (lane - 2) / 3, cc = (lane - 2) % 3;
                    const int dof = b == 0 ? 0 : 9, xo = b == 0 ? X_POS : X_OFFT;
                    S.xnew[xo + cc] = S.x[xo + cc] + S.dxu[dof + cc];                                                  // vect.hpp:117-119
                }
                __syncwarp();
                if (lane == 0) ctl->prof[10] = clock64();
                pub_publish(pub, tag_next, S.xnew, converge, finish, lane);
                if (lane == 0) { ctl->prof[5] = clock64(); S.searched = S.converge; S.t = t; S.converge = converge; }
            }
        }
        if (lane == 0) { S.effct = effct; S.ok = ok ? 1 : 0; S.finish = finish; }
        if (effct < 1 || !ok || S.late) {
            // ---- invalid pass (laserMapping.cpp:708-713, esekfom.hpp:1638-1641: `continue`) or a device-side failure
            if (lane == 0) {
                if (effct < 1 && !S.late) { S.iter++; if (S.iter >= S.max_iter) S.done = 1; S.searched = S.converge; }
                else { S.error = S.late ? (S.late == 2 ? 3 : 2) : 1; S.done = 1; }     // 2: a peer never delivered; 3: the workers never reported; 1: singular system
            }
            __syncwarp();
            pub_publish(pub, tag_next, S.x, S.converge, S.done, lane);
        }
    }
    __syncthreads();
    PassLog* lg = (logs && S.n_pass < MAX_LOGS) ? &logs[S.n_pass] : nullptr;
    if (S.effct < 1 || !S.ok || S.late) {
        if (tid == 0) {
            if (lg && S.effct < 1 && !S.late) {
                lg->searched = S.searched; lg->effct = 0; lg->res_sum = 0.0; lg->valid = 0; lg->converged = S.converge;
                for (int i = 0; i < XLEN; i++) lg->x_after[i] = S.x[i];
            }
            S.n_pass++;
            ctl->iter = S.iter; ctl->n_pass = S.n_pass; ctl->done = S.done; ctl->error = S.error;
        }
        __syncthreads();
        return;
    }
    // ------------------------------------------------------------------ after the publication
    const int finish = S.finish;
    // [+] on the rest of the state; on the last pass also the congruence blocks at dx_ (:1836-1876)
    if (warp == 1 && lane < 2 && finish) {
        const int idx = lane == 0 ? 3 : 6;
        const M33 J = transpose33(A_matrix(d3(S.dxu[idx], S.dxu[idx + 1], S.dxu[idx + 2])));
#pragma unroll 1
        for (int i = 0; i < 9; i++) S.J[lane][i] = J.m[i];
    }
    if (warp == 2 && lane == 0) {
        const D3 g = S2_boxplus(ld3(S.x + X_GRAV), S.dxu[21], S.dxu[22]);                           // S2.hpp:136-142
        st3(S.xnew + X_GRAV, g);
        if (finish) S2_congruence_B(g, ld3(S.xprop + X_GRAV), S.Bprop, S.dxu[21], S.dxu[22], S.M2);
    }
    if (warp == 3 && lane < 9) {
        const int b = lane / 3, c = lane % 3;
        const int dof = 12 + 3 * b, xo = b == 0 ? X_VEL : (b == 1 ? X_BG : X_BA);
        S.xnew[xo + c] = S.x[xo + c] + S.dxu[dof + c];
    }
    if (lg) {
        if (tid >= 64 && tid < 64 + 144) lg->HtH[tid - 64] = S.HTH[tid - 64];
        if (tid >= 224 && tid < 236) lg->Hth[tid - 224] = S.red[78 + tid - 224];
        if (tid == 255) { lg->searched = S.searched; lg->effct = S.effct; lg->res_sum = S.red[91]; lg->valid = 1; lg->converged = S.converge; }
    }
    if (finish && warp >= 4) {
        // final covariance (:1834-1927):  P = T2 (Pt - (Pt[:, :ne] / R) W Pt[:ne, :]) T2^T.  The part that needs neither the new
        // state nor the congruence at dx_ is formed by warps 4..7 while warps 1..3 are still busy with those.
        const int t4 = tid - 128;
#pragma unroll 1
        for (int e = t4; e < NE * n; e += 128) {
            const int a = e / n, j = e - a * n;
            double v = 0.0;
#pragma unroll
            for (int b = 0; b < NE; b++) v = fma(S.Wm[a * 13 + 1 + b], S.Pt[b * n + j], v);
            S.Y[e] = v;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
        for (int e = t4; e < n * n; e += 128) {
            const int i = e / n, j = e - i * n;
            double v = 0.0;
#pragma unroll
            for (int a = 0; a < NE; a++) v = fma(S.Pt[i * n + a] * Rinv, S.Y[a * n + j], v);
            S.W1[e] = S.Pt[e] - v;
        }
    }
    __syncthreads();
    if (tid < XLEN) { ctl->x[tid] = S.xnew[tid]; if (lg) lg->x_after[tid] = S.xnew[tid]; }
    if (tid == 32) { ctl->t = S.t; ctl->converge = S.converge; ctl->iter = S.iter + 1; ctl->n_pass = S.n_pass + 1; ctl->done = finish; }
    if (!finish) {
        // the reference leaves P_ = congruence-transformed P_propagated between passes
#pragma unroll 1
        for (int e = tid; e < n * n; e += UPD_THREADS) ctl->P[e] = S.Pt[e];
    } else {
#pragma unroll 1
        for (int e = tid; e < n * n; e += UPD_THREADS) {            // rows
            const int i = e / n, j = e - i * n;
            double v;
            if (i >= 3 && i < 9) {
                const int b = i >= 6, base = 3 + 3 * b, r = i - base;
                v = S.J[b][3 * r] * S.W1[base * n + j] + S.J[b][3 * r + 1] * S.W1[(base + 1) * n + j] + S.J[b][3 * r + 2] * S.W1[(base + 2) * n + j];
            } else if (i >= 21) {
                const int r = i - 21;
                v = S.M2[2 * r] * S.W1[21 * n + j] + S.M2[2 * r + 1] * S.W1[22 * n + j];
            } else v = S.W1[e];
            S.Pt[e] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int e = tid; e < n * n; e += UPD_THREADS) {            // columns
            const int i = e / n, j = e - i * n;
            const double* row = &S.Pt[i * n];
            double v;
            if (j >= 3 && j < 9) {
                const int b = j >= 6, base = 3 + 3 * b, r = j - base;
                v = S.J[b][3 * r] * row[base] + S.J[b][3 * r + 1] * row[base + 1] + S.J[b][3 * r + 2] * row[base + 2];
            } else if (j >= 21) {
                const int r = j - 21;
                v = S.M2[2 * r] * row[21] + S.M2[2 * r + 1] * row[22];
            } else v = row[j];
            ctl->P[e] = v;
        }
    }
    __syncthreads();
    if (tid < XLEN) S.x[tid] = S.xnew[tid];
    if (tid == 32) { S.iter++; S.n_pass++; S.done = finish; }
    __syncthreads();
}

// ============================================================================= the kernel
template <bool EXTR>
__global__ void __launch_bounds__(UPD_THREADS, 2) k_update(UpdArgs a) {
    __shared__ __align__(16) unsigned char smem_raw[sizeof(SolverSm) > sizeof(WorkerSm<EXTR>) ? sizeof(SolverSm) : sizeof(WorkerSm<EXTR>)];
    FilterCtl* ctl = a.ctl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwork = (int)gridDim.x - 1;
    pdl_wait();
    pdl_launch();
    if (blockIdx.x > 0) {
        // ------------------------------------------------------------------ worker block
        WorkerSm<EXTR>& Wk = *reinterpret_cast<WorkerSm<EXTR>*>(smem_raw);
        const int wb = (int)blockIdx.x - 1;
        if (tid == 0) { Wk.abort = 0; Wk.walks.n[0] = Wk.walks.n[1] = 0; }
        __syncthreads();
        int walk_phase = 0;
        for (int p = 0; p < a.max_passes; p++) {
            PoseS s;
            bool searched;
            if (p == 0) {
                // the state this launch starts from is in the control block (uploaded / left by the previous launch)
                if (__ldcg(&ctl->done) && !a.search_only) return;      // (a search-only launch may follow a finished update)
                searched = __ldcg(&ctl->converge) != 0 || a.search_only;      // dyn_share.converge (laserMapping.cpp:667)
                s = load_pose(a.pose_from_search ? ctl->x_search : ctl->x);
            } else {
                // later passes: wait for the solver block's publication of pass p (tagged words, see pub_publish)
                const unsigned tag = pub_tag(a.nonce, p);
                if (tid == 0) {
                    const long long t0 = clock64();
                    while ((unsigned)(pub_load(a.pub + 28) >> 32) != tag) {
                        if (clock64() - t0 > SPIN_LIMIT) { Wk.abort = 1; atomicExch(&ctl->error, 3); break; }
                        __nanosleep((a.dbg & 1) ? 2000 : 100);
                    }
                }
                __syncthreads();
                if (Wk.abort) return;
                if (tid < PUB_WORDS) {
                    unsigned long long w = pub_load(a.pub + tid);
                    const long long t0 = clock64();
                    while ((unsigned)(w >> 32) != tag && clock64() - t0 < SPIN_LIMIT) w = pub_load(a.pub + tid);
                    if (tid < 28) Wk.pose_bits[tid] = (unsigned)w; else Wk.flags = (unsigned)w;
                }
                __syncthreads();
                if (Wk.flags & 2u) return;
                searched = (Wk.flags & 1u) != 0;
                double x14[14];
#pragma unroll
                for (int i = 0; i < 14; i++) x14[i] = __longlong_as_double((long long)(((unsigned long long)Wk.pose_bits[2 * i + 1] << 32) | Wk.pose_bits[2 * i]));
                s.pos = d3(x14[X_POS], x14[X_POS + 1], x14[X_POS + 2]);
                s.rot.x = x14[X_ROT]; s.rot.y = x14[X_ROT + 1]; s.rot.z = x14[X_ROT + 2]; s.rot.w = x14[X_ROT + 3];
                s.offR.x = x14[X_OFFR]; s.offR.y = x14[X_OFFR + 1]; s.offR.z = x14[X_OFFR + 2]; s.offR.w = x14[X_OFFR + 3];
                s.offT = d3(x14[X_OFFT], x14[X_OFFT + 1], x14[X_OFFT + 2]);
            }
            double acc[3] = {0.0, 0.0, 0.0};
            double* stage = Wk.stage[warp];
            // tiles of UPD_THREADS consecutive points, dealt round-robin to the worker blocks -- the same tiles every pass (a point's
            // cached neighbours, plane and flag are only ever touched by its own thread)
            const int q0 = a.sc.q_begin, q1 = a.sc.q_end;
            for (int tile = q0 + wb * UPD_THREADS; tile < q1; tile += nwork * UPD_THREADS) {
                const int q = tile + tid;
                double h[12]; double z = 0.0; float ar = 0.f;
                const bool contrib = measure_fused<EXTR>(a.m, a.sc, q, q < q1, s, searched, a.search_only != 0, Wk.walks, walk_phase, h, z, ar);
                if (!a.search_only) warp_accumulate<EXTR>(contrib, h, z, ar, acc, lane, stage);
            }
            if (a.search_only) return;
#pragma unroll
            for (int j = 0; j < 3; j++) Wk.wred[warp][lane + 32 * j] = acc[j];
            __syncthreads();
            if (tid < PSTRIDE) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < UPD_WARPS; w++) v += Wk.wred[w][tid];
                __stcg(&a.partials[(size_t)wb * PSTRIDE + tid], v);
            }
            __syncthreads();                                // the block's partial is written ...
            if (tid == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(&ctl->ticket) : "memory");      // ... and ordered before the ticket (release)
            if (a.mode == 1) return;
        }
        return;
    }
    // ------------------------------------------------------------------ solver block
    if (a.search_only) return;
    SolverSm& S = *reinterpret_cast<SolverSm*>(smem_raw);
    sol_load(S, ctl);
    int p = 0;
    for (; p < a.max_passes && !S.done; p++) {
        if (tid == 0) ctl->prof[0] = clock64();
        if (S.converge && tid < XLEN) ctl->x_search[tid] = S.x[tid];      // this pass searches: Nearest_Points will belong to this state
        if (a.mode != 1) sol_prepare(S);                 // overlaps the workers' measurement
        if (tid == 0) ctl->prof[8] = clock64();
        if (a.mode != 3) {
            if (tid == 0) {
                const long long t0 = clock64();
                while (ld_acquire(&ctl->ticket) < nwork * (p + 1)) {        // tickets only grow within a launch
                    if (clock64() - t0 > SPIN_LIMIT) { S.late = 2; break; }
                }
            }
            __syncthreads();
            if (tid == 0) ctl->prof[9] = clock64();
            sol_reduce(S, a.partials, nwork);
            if (a.mode == 1) {
                if (tid < NRED) a.red_g[tid] = S.red[tid];
                if (tid == 0) ctl->ticket = 0;
                return;
            }
        } else {
            if (tid < PSTRIDE) S.red[tid] = tid < NRED ? a.red_g[tid] : 0.0;
            __syncthreads();
            sol_expand(S);
        }
        if (a.mode == 2) { sol_exchange(S, a.p2p); sol_expand(S); }
        if (tid == 0) ctl->prof[1] = clock64();
        sol_pass<EXTR>(S, ctl, a.logs, a.pub, pub_tag(a.nonce, p + 1));
        if (tid == 0) ctl->prof[7] = clock64();
    }
    if (tid == 0) { ctl->error = S.error | ctl->error; ctl->ticket = 0; }
    mirror_result(ctl);
}

}  // namespace fl
