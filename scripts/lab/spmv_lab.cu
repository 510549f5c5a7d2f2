// spmv_lab.cu — micro-benchmark of SpMV kernel variants on the fine-level (7-point, 128^3) and a wide coarse-like matrix.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o spmv_lab spmv_lab.cu && ./spmv_lab
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct M { int n; long long nnz; int *rp, *col; double* val; };

static M make_stencil(int nx, int ny, int nz, int reach) {  // (2*reach+1)-wide stencil in x plus +-1 in y, z planes scaled by reach
    std::vector<int> rp(1, 0), col; std::vector<double> val;
    int n = nx * ny * nz;
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        std::vector<int> c;
        for (int dk = -1; dk <= 1; ++dk) for (int dj = -1; dj <= 1; ++dj) for (int di = -reach; di <= reach; ++di) {
            if (reach == 0 && (abs(dk) + abs(dj) + abs(di)) > 1) continue;
            if (reach == 0 && false) continue;
            int ii = i + di, jj = j + dj, kk = k + dk;
            if (reach == 0) { /* 7-point */ }
            else if (abs(dk) + abs(dj) > 1 && reach < 3) continue;
            if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz) continue;
            c.push_back(ii + nx * (jj + ny * kk));
        }
        if (reach == 0) {
            c.clear();
            int id = i + nx * (j + ny * k);
            if (k > 0) c.push_back(id - nx * ny); if (j > 0) c.push_back(id - nx); if (i > 0) c.push_back(id - 1);
            c.push_back(id);
            if (i < nx - 1) c.push_back(id + 1); if (j < ny - 1) c.push_back(id + nx); if (k < nz - 1) c.push_back(id + nx * ny);
        }
        std::sort(c.begin(), c.end());
        for (int x : c) { col.push_back(x); val.push_back(1.0 / (1 + (x % 7))); }
        rp.push_back((int)col.size());
    }
    M m; m.n = n; m.nnz = (long long)col.size();
    CK(cudaMalloc(&m.rp, sizeof(int) * (n + 1))); CK(cudaMalloc(&m.col, sizeof(int) * m.nnz)); CK(cudaMalloc(&m.val, sizeof(double) * m.nnz));
    CK(cudaMemcpy(m.rp, rp.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.col, col.data(), sizeof(int) * m.nnz, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.val, val.data(), sizeof(double) * m.nnz, cudaMemcpyHostToDevice));
    return m;
}

// ---- A: streaming ceiling: read val + col, write n doubles ----
__global__ void k_stream(long long nnz, const double* __restrict__ val, const int* __restrict__ col, double* y, int n) {
    double acc = 0.;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) acc += val[k] * (double)col[k];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = acc;
}
// ---- D: CSR scalar ----
__global__ void k_scalar(int n, const int* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x, double* y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double acc = 0.;
        for (int k = rp[i]; k < rp[i + 1]; ++k) acc += val[k] * x[col[k]];
        y[i] = acc;
    }
}
// ---- E: CSR vector, G lanes per row, UN unrolled independent loads ----
template <int G, int UN>
__global__ void k_vec(int n, const int* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x, double* y) {
    const int t = threadIdx.x, gl = t & (G - 1);
    const int rpb = blockDim.x / G;
    for (int i = blockIdx.x * rpb + t / G; i < n + rpb; i += gridDim.x * rpb) {
        double acc = 0.;
        if (i < n) {
            const int lo = rp[i], hi = rp[i + 1];
            int k = lo + gl;
            for (; k + (UN - 1) * G < hi; k += UN * G) {
                double v[UN]; int c[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) { v[u] = val[k + u * G]; c[u] = col[k + u * G]; }
                double xv[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) xv[u] = x[c[u]];
#pragma unroll
                for (int u = 0; u < UN; ++u) acc += v[u] * xv[u];
            }
            for (; k < hi; k += G) acc += val[k] * x[col[k]];
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, G);
        if (i < n && gl == 0) y[i] = acc;
    }
}
// ---- E2: CSR vector with a fully predicated unrolled body (no serial tail) ----
template <int G, int UN>
__global__ void k_vecp(int n, const int* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x, double* y) {
    const int t = threadIdx.x, gl = t & (G - 1);
    const int rpb = blockDim.x / G;
    for (int i = blockIdx.x * rpb + t / G; i < n + rpb; i += gridDim.x * rpb) {
        double acc = 0.;
        if (i < n) {
            const int lo = rp[i], hi = rp[i + 1];
            for (int k = lo + gl; k < hi; k += UN * G) {
                double v[UN], xv[UN]; int c[UN]; bool ok[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) { ok[u] = k + u * G < hi; v[u] = ok[u] ? val[k + u * G] : 0.; c[u] = ok[u] ? col[k + u * G] : i; }
#pragma unroll
                for (int u = 0; u < UN; ++u) xv[u] = x[c[u]];
#pragma unroll
                for (int u = 0; u < UN; ++u) if (ok[u]) acc += v[u] * xv[u];
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, G);
        if (i < n && gl == 0) y[i] = acc;
    }
}
// ---- B: block-staged ordered (production kernel shape), U loads per thread, ROWS rows per block ----
template <int BLOCK, int U>
__global__ void __launch_bounds__(BLOCK) k_staged(int n, const int* __restrict__ rp_, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x, double* y) {
    constexpr int CAP = BLOCK * U;
    __shared__ double prod[CAP];
    __shared__ int rp[BLOCK + 1];
    const int t = threadIdx.x;
    const int ntiles = (n + BLOCK - 1) / BLOCK;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int r0 = tile * BLOCK, nr = min(BLOCK, n - r0);
        __syncthreads();
        for (int q = t; q <= nr; q += BLOCK) rp[q] = rp_[r0 + q];
        __syncthreads();
        const int kbeg = rp[0], kend = rp[nr];
        const int lo = (t < nr) ? rp[t] : 0, hi = (t < nr) ? rp[t + 1] : 0;
        double acc = 0.;
        for (int c = kbeg; c < kend; c += CAP) {
            const int ce = min(c + CAP, kend);
            double v[U]; int cc[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { int k = c + t + u * BLOCK; bool ok = k < ce; v[u] = ok ? val[k] : 0.; cc[u] = ok ? col[k] : -1; }
#pragma unroll
            for (int u = 0; u < U; ++u) if (cc[u] >= 0) prod[t + u * BLOCK] = v[u] * x[cc[u]];
            __syncthreads();
            const int b0 = max(lo, c), b1 = min(hi, ce);
            for (int k = b0; k < b1; ++k) acc += prod[k - c];
            __syncthreads();
        }
        if (t < nr) y[r0 + t] = acc;
    }
}
// ---- F: warp-tile ordered: each warp stages the nnz slice of its 32 rows in its own smem, no block barriers ----
template <int WARPS, int U>
__global__ void __launch_bounds__(WARPS * 32) k_warptile(int n, const int* __restrict__ rp_, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x, double* y) {
    constexpr int CAP = 32 * U;
    __shared__ double prod[WARPS][CAP];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nw = gridDim.x * WARPS, gw = blockIdx.x * WARPS + w;
    const int ntiles = (n + 31) / 32;
    for (int tile = gw; tile < ntiles; tile += nw) {
        const int r0 = tile * 32, nr = min(32, n - r0);
        const int lo = (lane < nr) ? rp_[r0 + lane] : 0, hi = (lane < nr) ? rp_[r0 + lane + 1] : 0;
        const int kbeg = __shfl_sync(0xffffffffu, lo, 0);
        const int kend = __shfl_sync(0xffffffffu, hi, nr - 1);
        double acc = 0.;
        for (int c = kbeg; c < kend; c += CAP) {
            const int ce = min(c + CAP, kend);
            double v[U]; int cc[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { int k = c + lane + u * 32; bool ok = k < ce; v[u] = ok ? val[k] : 0.; cc[u] = ok ? col[k] : -1; }
#pragma unroll
            for (int u = 0; u < U; ++u) if (cc[u] >= 0) prod[w][lane + u * 32] = v[u] * x[cc[u]];
            __syncwarp();
            const int b0 = max(lo, c), b1 = min(hi, ce);
            for (int k = b0; k < b1; ++k) acc += prod[w][k - c];
            __syncwarp();
        }
        if (lane < nr) y[r0 + lane] = acc;
    }
}

template <class F>
static double timeit(F f, int reps = 20) {
    for (int i = 0; i < 3; ++i) f();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    CK(cudaGetLastError());
    return ms / reps * 1e3;
}

static void run(const char* name, const M& m) {
    double *x, *y;
    CK(cudaMalloc(&x, sizeof(double) * m.n)); CK(cudaMalloc(&y, sizeof(double) * (m.n + 4096)));
    std::vector<double> hx(m.n); for (int i = 0; i < m.n; ++i) hx[i] = 1.0 + (i % 13) * 0.01;
    CK(cudaMemcpy(x, hx.data(), sizeof(double) * m.n, cudaMemcpyHostToDevice));
    const double bytes = 12.0 * m.nnz + 20.0 * m.n;
    printf("== %s: n=%d nnz=%lld (%.1f/row) algorithmic %.1f MB\n", name, m.n, m.nnz, (double)m.nnz / m.n, bytes / 1e6);
    auto rep = [&](const char* k, double us) { printf("  %-28s %8.1f us  %7.1f GB/s\n", k, us, bytes / us / 1e3); };
    const int SM = 148;
    rep("A stream 1184x256", timeit([&] { k_stream<<<SM * 8, 256>>>(m.nnz, m.val, m.col, y, m.n); }));
    rep("A stream 2368x256", timeit([&] { k_stream<<<SM * 16, 256>>>(m.nnz, m.val, m.col, y, m.n); }));
    rep("D scalar", timeit([&] { k_scalar<<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("B staged 256x8", timeit([&] { k_staged<256, 8><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("B staged 256x4", timeit([&] { k_staged<256, 4><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("B staged 128x8", timeit([&] { k_staged<128, 8><<<SM * 16, 128>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("B staged 128x16", timeit([&] { k_staged<128, 16><<<SM * 12, 128>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("B staged 512x8", timeit([&] { k_staged<512, 8><<<SM * 4, 512>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("F warptile 8w x8", timeit([&] { k_warptile<8, 8><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("F warptile 8w x16", timeit([&] { k_warptile<8, 16><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("F warptile 4w x8", timeit([&] { k_warptile<4, 8><<<SM * 16, 128>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E vec G4 U2", timeit([&] { k_vec<4, 2><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E vec G8 U1", timeit([&] { k_vec<8, 1><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E vec G8 U2", timeit([&] { k_vec<8, 2><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E vec G8 U4", timeit([&] { k_vec<8, 4><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E vec G16 U2", timeit([&] { k_vec<16, 2><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E vec G16 U4", timeit([&] { k_vec<16, 4><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E2 vecp G2 U4", timeit([&] { k_vecp<2, 4><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E2 vecp G4 U4", timeit([&] { k_vecp<4, 4><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E2 vecp G4 U8", timeit([&] { k_vecp<4, 8><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E2 vecp G8 U4", timeit([&] { k_vecp<8, 4><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E2 vecp G8 U8", timeit([&] { k_vecp<8, 8><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E2 vecp G16 U4", timeit([&] { k_vecp<16, 4><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E2 vecp G16 U8", timeit([&] { k_vecp<16, 8><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("D scalar 2368 blocks", timeit([&] { k_scalar<<<SM * 16, 128>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E vec G32 U2", timeit([&] { k_vec<32, 2><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    rep("E vec G32 U4", timeit([&] { k_vec<32, 4><<<SM * 8, 256>>>(m.n, m.rp, m.col, m.val, x, y); }));
    cudaFree(x); cudaFree(y);
}

int main() {
    { M m = make_stencil(128, 128, 128, 0); run("fine 7-point 128^3", m); }
    { M m = make_stencil(64, 128, 128, 2); run("coarse-like (~25/row) 1M rows", m); }
    { M m = make_stencil(32, 128, 128, 3); run("coarse-like (~60/row) 0.5M rows", m); }
    { M m = make_stencil(64, 64, 64, 6); run("coarse-like (~110/row) 0.26M rows", m); }
    return 0;
}
