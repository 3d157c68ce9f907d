// standalone check of the shared-memory Gauss-Jordan routines (run on the GPU box)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../fast_lio_b200/csrc/gj.cuh"
using namespace fl;
namespace fl { void set_last_error(const char*, ...) {} }

__global__ void k_block(const double* in, double* out, int n, int nrhs, int* okflag) {
    __shared__ double a[23 * 46];
    __shared__ int row_of[32];
    const int nc = n + nrhs, ld = nc;
    for (int e = threadIdx.x; e < n * nc; e += blockDim.x) a[e] = in[e];
    __syncthreads();
    bool ok = gj_eliminate(a, n, nc, ld, row_of);
    if (threadIdx.x == 0) *okflag = ok;
    for (int e = threadIdx.x; e < n * nrhs; e += blockDim.x) { int k = e / nrhs, j = e % nrhs; out[e] = a[row_of[k] * ld + n + j] / a[row_of[k] * ld + k]; }
}
__global__ void k_time(const double* in, int n, int nrhs, long long* t) {
    __shared__ double a[12 * 32];
    __shared__ int row_of[32];
    const int nc = n + nrhs, ld = nc;
    for (int rep = 0; rep < 3; rep++) {
        for (int e = threadIdx.x; e < n * nc; e += blockDim.x) a[e] = in[e];
        __syncthreads();
        long long t0 = clock64();
        if (threadIdx.x < 32) { if (n == 6) gj_warp_reg<6>(a, nc, ld, row_of, threadIdx.x); else gj_warp_reg<12>(a, nc, ld, row_of, threadIdx.x); }
        __syncthreads();
        long long t1 = clock64();
        for (int e = threadIdx.x; e < n * nc; e += blockDim.x) a[e] = in[e];
        __syncthreads();
        long long t2 = clock64();
        if (threadIdx.x < 32) gj_warp(a, n, nc, ld, row_of, threadIdx.x);
        __syncthreads();
        long long t3 = clock64();
        if (threadIdx.x == 0) { t[rep * 2] = t1 - t0; t[rep * 2 + 1] = t3 - t2; }
    }
}
__global__ void k_warp(const double* in, double* out, int n, int nrhs, int* okflag) {
    __shared__ double a[12 * 32];
    __shared__ int row_of[32];
    const int nc = n + nrhs, ld = nc;
    for (int e = threadIdx.x; e < n * nc; e += blockDim.x) a[e] = in[e];
    __syncthreads();
    if (threadIdx.x < 32) { bool ok = n == 6 ? gj_warp_reg<6>(a, nc, ld, row_of, threadIdx.x) : (n == 12 ? gj_warp_reg<12>(a, nc, ld, row_of, threadIdx.x) : gj_warp(a, n, nc, ld, row_of, threadIdx.x)); if (threadIdx.x == 0) *okflag = ok; }
    __syncthreads();
    for (int e = threadIdx.x; e < n * nrhs; e += blockDim.x) { int k = e / nrhs, j = e % nrhs; out[e] = a[row_of[k] * ld + n + j] / a[row_of[k] * ld + k]; }
}
static void host_solve(std::vector<double> a, int n, int nrhs, std::vector<double>& x) {
    int nc = n + nrhs;
    for (int k = 0; k < n; k++) {
        int p = k; for (int r = k; r < n; r++) if (fabs(a[r*nc+k]) > fabs(a[p*nc+k])) p = r;
        for (int j = 0; j < nc; j++) std::swap(a[k*nc+j], a[p*nc+j]);
        double pv = a[k*nc+k]; for (int j = 0; j < nc; j++) a[k*nc+j] /= pv;
        for (int i = 0; i < n; i++) if (i != k) { double f = a[i*nc+k]; for (int j = 0; j < nc; j++) a[i*nc+j] -= f * a[k*nc+j]; }
    }
    x.resize(n * nrhs);
    for (int k = 0; k < n; k++) for (int j = 0; j < nrhs; j++) x[k*nrhs+j] = a[k*nc+n+j];
}
int main() {
    srand(3);
    int cases[4][3] = {{23, 23, 256}, {23, 23, 512}, {6, 7, 256}, {12, 13, 256}};
    for (int c = 0; c < 4; c++) {
        int n = cases[c][0], nrhs = cases[c][1], nt = cases[c][2], nc = n + nrhs;
        std::vector<double> a(n * nc);
        for (int i = 0; i < n; i++) for (int j = 0; j < nc; j++) a[i*nc+j] = (rand() / (double)RAND_MAX - 0.5) * (i < 6 ? 1e5 : 1.0) + (i == j ? (i < 6 ? 3e5 : 3.0) : 0.0);
        std::vector<double> ref; host_solve(a, n, nrhs, ref);
        double *din, *dout; int* dok;
        cudaMalloc(&din, a.size() * 8); cudaMalloc(&dout, n * nrhs * 8); cudaMalloc(&dok, 4);
        cudaMemcpy(din, a.data(), a.size() * 8, cudaMemcpyHostToDevice);
        std::vector<double> got(n * nrhs);
        for (int which = 0; which < 2; which++) {
            if (which == 1 && n > 12) continue;
            if (which == 0) k_block<<<1, nt>>>(din, dout, n, nrhs, dok); else k_warp<<<1, nt>>>(din, dout, n, nrhs, dok);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(got.data(), dout, got.size() * 8, cudaMemcpyDeviceToHost);
            double md = 0, mr = 0; for (size_t i = 0; i < got.size(); i++) { md = fmax(md, fabs(got[i] - ref[i])); mr = fmax(mr, fabs(ref[i])); }
            printf("n=%d nrhs=%d nt=%d %s: err=%s maxdiff=%.3e (max |x|=%.3e)\n", n, nrhs, nt, which ? "warp" : "block", cudaGetErrorString(e), md, mr);
        }
        if (n <= 12) { long long* dt; cudaMalloc(&dt, 64); k_time<<<1, 256>>>(din, n, nrhs, dt); long long ht[6]; cudaMemcpy(ht, dt, 48, cudaMemcpyDeviceToHost); printf("  cycles reg/smem: rep0 %lld / %lld  rep1 %lld / %lld  rep2 %lld / %lld\n", ht[0], ht[1], ht[2], ht[3], ht[4], ht[5]); }
    }
    return 0;
}
