// reorder_lab.cu — does a locality ordering of the coarse AMG unknowns (Morton code of the aggregate positions) speed up the
// production gather SpMV? Runs the production lane mapping (k_spmv_vec: G lanes per row, UN loads in flight per lane) on a matrix
// file, one and three systems. Feed it the natural and the permuted dump of the same level (scripts/lab/dump_levels.py --morton).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o reorder_lab reorder_lab.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
struct alignas(32) D4 { double a, b, c, d; };
template <int K> struct Cell;
template <> struct Cell<1> {
    static constexpr int S = 1;
    static __device__ __forceinline__ void ld(const double* p, size_t i, double (&v)[1]) { v[0] = p[i]; }
    static __device__ __forceinline__ void st(double* p, size_t i, const double (&v)[1]) { p[i] = v[0]; }
};
template <> struct Cell<3> {
    static constexpr int S = 4;
    static __device__ __forceinline__ void ld(const double* p, size_t i, double (&v)[3]) { const D4 q = reinterpret_cast<const D4*>(p)[i]; v[0] = q.a; v[1] = q.b; v[2] = q.c; }
    static __device__ __forceinline__ void st(double* p, size_t i, const double (&v)[3]) { D4 q; q.a = v[0]; q.b = v[1]; q.c = v[2]; q.d = 0.; reinterpret_cast<D4*>(p)[i] = q; }
};
template <int G, int UN, int K>
__global__ void __launch_bounds__(256, (K == 1 ? 8 : 6)) k_gather(int n, const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                                                                  const double* __restrict__ x, double* __restrict__ y) {
    const int t = threadIdx.x, gl = t & (G - 1);
    constexpr int RPB = 256 / G;
    const int ngroups = (n + RPB - 1) / RPB;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int i = grp * RPB + t / G;
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.;
        if (i < n) {
            const int lo = rowptr[i], hi = rowptr[i + 1];
            for (int q = lo + gl; q < hi; q += UN * G) {
                double v[UN], xv[UN][K]; int cc[UN]; bool ok[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) { ok[u] = q + u * G < hi; v[u] = ok[u] ? val[q + u * G] : 0.; cc[u] = ok[u] ? col[q + u * G] : 0; }
#pragma unroll
                for (int u = 0; u < UN; ++u) Cell<K>::ld(x, cc[u], xv[u]);
#pragma unroll
                for (int u = 0; u < UN; ++u) if (ok[u]) {
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] += v[u] * xv[u][k];
                }
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o, G);
        }
        if (i < n && gl == 0) Cell<K>::st(y, i, acc);
    }
}
struct Mat { int n; long long nnz; int *rp, *col; double* val; };
static Mat load(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { printf("cannot open %s\n", path); exit(1); }
    long long hdr[2];
    if (fread(hdr, 8, 2, f) != 2) exit(1);
    Mat m; m.n = (int)hdr[0]; m.nnz = hdr[1];
    std::vector<int> rp(m.n + 1), col(m.nnz); std::vector<double> val(m.nnz);
    if (fread(rp.data(), 4, rp.size(), f) != rp.size() || fread(col.data(), 4, col.size(), f) != col.size() || fread(val.data(), 8, val.size(), f) != val.size()) exit(1);
    fclose(f);
    CK(cudaMalloc(&m.rp, 4 * rp.size())); CK(cudaMalloc(&m.col, 4 * col.size())); CK(cudaMalloc(&m.val, 8 * val.size()));
    CK(cudaMemcpy(m.rp, rp.data(), 4 * rp.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(m.col, col.data(), 4 * col.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.val, val.data(), 8 * val.size(), cudaMemcpyHostToDevice));
    return m;
}
template <class F> static float timeit(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    CK(cudaGetLastError());
    return ms / reps * 1e3f;
}
template <int K, int G, int UN> static void run(const Mat& m, int reps) {
    constexpr int S = Cell<K>::S;
    const size_t nx = (size_t)m.n * S;
    double *x, *y;
    CK(cudaMalloc(&x, 8 * nx)); CK(cudaMalloc(&y, 8 * nx));
    std::vector<double> hx(nx);
    for (size_t i = 0; i < nx; ++i) hx[i] = 0.001 * (double)((i * 2654435761u) % 1000) - 0.5;
    CK(cudaMemcpy(x, hx.data(), 8 * nx, cudaMemcpyHostToDevice));
    const double alg = 12. * m.nnz + 4. * m.n + 16. * K * m.n;
    const int RPB = 256 / G, ngroups = (m.n + RPB - 1) / RPB;
    const int gg = std::min(ngroups, 148 * (K == 1 ? 8 : 6));
    float us = timeit([&] { k_gather<G, UN, K><<<gg, 256>>>(m.n, m.rp, m.col, m.val, x, y); }, reps);
    printf("  K=%d G=%2d UN=%d: %6.1f us = %4.0f GB/s\n", K, G, UN, us, alg / us * 1e-3);
    cudaFree(x); cudaFree(y);
}
int main(int argc, char** argv) {
    if (argc < 2) { printf("usage: reorder_lab matrix.bin [reps]\n"); return 1; }
    const int reps = argc > 2 ? atoi(argv[2]) : 50;
    Mat m = load(argv[1]);
    printf("%s: %d rows, %lld entries (%.1f per row)\n", argv[1], m.n, m.nnz, (double)m.nnz / m.n);
    run<1, 4, 4>(m, reps); run<1, 8, 4>(m, reps); run<1, 16, 2>(m, reps); run<1, 8, 2>(m, reps);
    run<3, 4, 2>(m, reps); run<3, 8, 2>(m, reps); run<3, 16, 2>(m, reps); run<3, 8, 4>(m, reps); run<3, 16, 1>(m, reps);
    return 0;
}
