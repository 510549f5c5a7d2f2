// small.cu — single-launch solvers for systems small enough that ONE thread block can hold the whole solve
// (north star kernel group 4: "shared-memory-staged coarse levels").
//
// The reference's own meshes (1 008 and 8 001 cells) and every AMG level under them are launch bound on the multi-kernel path:
// a 50-iteration BiCGSTAB in reference-order mode is 500 launches of a few microseconds, a SIMPLE iteration on channel_flow.msh
// 11 000 launches. Here the whole loop of src/linear_algebra.rs:247-269 runs in ONE launch: the five work vectors live in shared
// memory (up to 5 000 rows; in L1-cached global scratch above), the matrix streams through the block's L1, block barriers
// separate the phases and the dot products follow nalgebra's eight-accumulator order — the arithmetic is, operation for
// operation, that of bicgstab_reference_order() in linalg.cu, so the results are bit-identical to it and to the oracle.
// Likewise the lexicographic Gauss-Seidel (the intended formula of :219-246): all sweeps in one launch, the dataflow of
// k_gs_sweep with its ready flags in shared memory (a dependency hop costs a shared-memory round trip instead of an L2 one).
#include "linalg.cuh"

#include <algorithm>

namespace orc {

constexpr int SMALL_T = 1024;
constexpr int SMALL_WARPS = SMALL_T / 32;
constexpr size_t SMALL_SMEM_MAX = 200 * 1024;   // dynamic shared memory the block may ask for (227 KB per SM minus static)

// nalgebra's dot (base/blas.rs), see k_dot_ref in linalg.cu: eight interleaved sequential accumulators over chunks of eight,
// res += acc0+acc4; res += acc1+acc5; res += acc2+acc6; res += acc3+acc7; then the tail in order. Called by a whole warp; the
// result is valid in lane 0. `b == nullptr` stands for a vector of ones (r_hat_0).
__device__ __forceinline__ double warp_dot_ref(const double* a, const double* b, int n) {
    const int lane = threadIdx.x & 31;
    const int m = n - (n % 8);
    double acc = 0.;
    if (lane < 8) {
        int i = lane;
        for (; i + 56 < m; i += 64) {  // eight loads in flight, then the dependent chain (same order as the plain loop)
            double pr[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) pr[u] = a[i + 8 * u] * (b ? b[i + 8 * u] : 1.);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += pr[u];
        }
        for (; i < m; i += 8) acc += a[i] * (b ? b[i] : 1.);
    }
    const double a0 = __shfl_sync(0xffffffffu, acc, 0), a1 = __shfl_sync(0xffffffffu, acc, 1), a2 = __shfl_sync(0xffffffffu, acc, 2),
                 a3 = __shfl_sync(0xffffffffu, acc, 3), a4 = __shfl_sync(0xffffffffu, acc, 4), a5 = __shfl_sync(0xffffffffu, acc, 5),
                 a6 = __shfl_sync(0xffffffffu, acc, 6), a7 = __shfl_sync(0xffffffffu, acc, 7);
    double res = 0.;
    if (lane == 0) {
        res += a0 + a4; res += a1 + a5; res += a2 + a6; res += a3 + a7;
        for (int i = m; i < n; ++i) res += a[i] * (b ? b[i] : 1.);
    }
    return res;
}

// y_i = sum_k a_ik x_k in ascending k (thread per row, the order of k_spmv)
__device__ __forceinline__ double row_dot(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, const double* x, int i) {
    double acc = 0.;
    const int lo = rowptr[i], hi = rowptr[i + 1];
    for (int q = lo; q < hi; ++q) acc += val[q] * x[col[q]];
    return acc;
}

// BiCGSTAB (src/linear_algebra.rs:247-269), unguarded, r_hat_0 = 1, exactly `iterations` iterations, one block.
// `work`: 5 n doubles of global scratch, used when the vectors do not fit into shared memory (vec_in_smem == 0).
__global__ void __launch_bounds__(SMALL_T, 1) k_bicgstab_small(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                               const double* __restrict__ val, const double* __restrict__ b, double* x,
                                                               double* work, int vec_in_smem, unsigned long long iterations) {
    extern __shared__ __align__(16) double sm_vec[];
    __shared__ double s_rho, s_alpha, s_ts, s_tt, s_beta;
    double* base = vec_in_smem ? sm_vec : work;
    double *r = base, *p = base + n, *nu = base + 2 * (size_t)n, *s = base + 3 * (size_t)n, *tv = base + 4 * (size_t)n;
    const int t = threadIdx.x, wid = t >> 5;
    for (int i = t; i < n; i += SMALL_T) {  // r = b - A x; p = r  (:250-252)
        const double ri = b[i] - row_dot(rowptr, col, val, x, i);
        r[i] = ri; p[i] = ri;
    }
    __syncthreads();
    if (wid == 0) { const double v = warp_dot_ref(r, nullptr, n); if (t == 0) s_rho = v; }  // rho = r . r_hat_0  (:253)
    __syncthreads();
    for (unsigned long long it = 0; it < iterations; ++it) {
        for (int i = t; i < n; i += SMALL_T) nu[i] = row_dot(rowptr, col, val, p, i);       // nu = A p  (:256)
        __syncthreads();
        if (wid == 0) { const double v = warp_dot_ref(nu, nullptr, n); if (t == 0) s_alpha = s_rho / v; }  // alpha = rho / (r_hat_0 . nu)  (:257)
        __syncthreads();
        const double alpha = s_alpha;
        for (int i = t; i < n; i += SMALL_T) s[i] = r[i] - alpha * nu[i];                   // s = r - alpha nu  (:259)
        __syncthreads();
        for (int i = t; i < n; i += SMALL_T) tv[i] = row_dot(rowptr, col, val, s, i);       // t = A s  (:260)
        __syncthreads();
        if (wid == 0) { const double v = warp_dot_ref(tv, s, n); if (t == 0) s_ts = v; }    // omega = (t . s) / (t . t)  (:261)
        if (wid == 1) { const double v = warp_dot_ref(tv, tv, n); if (t == 32) s_tt = v; }
        __syncthreads();
        const double omega = s_ts / s_tt;
        for (int i = t; i < n; i += SMALL_T) {
            const double si = s[i];
            const double h = x[i] + alpha * p[i];   // h = x + alpha p      (:258)
            x[i] = h + omega * si;                  // x = h + omega s      (:262)
            r[i] = si - omega * tv[i];              // r = s - omega t      (:263)
        }
        __syncthreads();
        if (wid == 0) {                             // rho, beta = rho / rho_prev * alpha / omega  (:265-266)
            const double v = warp_dot_ref(r, nullptr, n);
            if (t == 0) { const double rho_prev = s_rho; s_rho = v; s_beta = v / rho_prev * alpha / omega; }
        }
        __syncthreads();
        const double beta = s_beta;
        for (int i = t; i < n; i += SMALL_T) p[i] = r[i] + beta * (p[i] - omega * nu[i]);   // p = r + beta (p - omega nu)  (:267)
        __syncthreads();
    }
}

bool small_enabled() {   // ORC_B200_SMALL=0 switches every one-block kernel off (A/B runs, tests of the multi-launch paths)
    static const bool off = [] { const char* e = getenv("ORC_B200_SMALL"); return e && atoi(e) == 0; }();
    return !off;
}
bool small_solve_ok(const Ctx& c, const DCsr& A) {
    return small_enabled() && c.exact_order && A.nrows == A.ncols && A.nrows > 0 && A.nrows <= ORC_AUTO_EXACT_MAX_ROWS;
}

void bicgstab_small(Ctx& c, const DCsr& A, const double* b, double* x, uint64_t iterations) {
    const int n = (int)A.nrows;
    const size_t need = 5 * (size_t)n * sizeof(double);
    const bool in_smem = need <= SMALL_SMEM_MAX;
    DBuf<double> work;
    if (!in_smem) work.alloc(&c, 5 * (size_t)n);
    static bool attr_set = false;
    if (!attr_set) {
        ORC_CUDA(cudaFuncSetAttribute(k_bicgstab_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM_MAX));
        attr_set = true;
    }
    ProfScope ps(c, PC_BICG, (double)iterations * (2. * (12. * (double)A.nnz + 20. * n) + 104. * n), -1., (int64_t)A.nnz * 8 + 5);
    k_bicgstab_small<<<1, SMALL_T, in_smem ? need : 0, c.stream>>>(n, A.rowptr, A.col, A.val, b, x, work.p, in_smem ? 1 : 0,
                                                                  (unsigned long long)iterations);
    c.after_launch("k_bicgstab_small");
}

// Lexicographic Gauss-Seidel / SOR, all sweeps in one launch (one block). x and the per-row sweep counters live in shared memory.
// Warp w owns the row chunks w, w + 32, ... (ascending), walks the rows of a chunk in order and, before it reads x_j of a stored
// neighbour outside its chunk, waits until row j has finished THIS sweep (j < i) or the PREVIOUS one (j > i) — exactly the values
// the sequential sweep sees. Needs a structurally symmetric pattern (so that row j > i cannot overtake row i within a sweep).
// Deadlock-free: waits point to lower chunks of the same sweep or to the previous sweep, and every warp takes its chunks in
// ascending order. The in-row sum is k_gs_sweep's ordered sum, so the result is bit-identical to it.
//
// What bounds it is the dependency chain i-1 -> i (a y-fastest 2-D mesh: 63 rows per column, 127 columns), so (i) a chunk is 64
// rows — about one mesh column, the 32 warps then work on 32 columns at once, each one row behind its left neighbour — and
// (ii) nothing on the chain may wait for the matrix: the rows are fetched four at a time (eight entry slots per row, one lane
// per slot), one group ahead of the row being solved, so the L2 latency of (col, val, b) overlaps the chain instead of sitting
// three times in every hop (measured: 2 100 cycles per row without the prefetch). Rows longer than eight entries take a
// generic path with direct loads.
constexpr int GSS_CHUNK = 64;
struct GsSub {          // one lane's share of a group of four rows: entry slot (lane & 7) of row (lane >> 3)
    int j;              // column, -1 = no entry in this slot
    double v;           // value
    int len;            // entries of the row (valid in every lane of the row's group)
    int lo;             // first entry of the row
    double bi;          // right-hand side of the row
};
__device__ __forceinline__ GsSub gs_load_sub(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                                             const double* __restrict__ b, int rs, int r1, int lane) {
    GsSub d;
    d.j = -1; d.v = 0.; d.len = 0; d.lo = 0; d.bi = 0.;
    const int row = rs + (lane >> 3), slot = lane & 7;
    if (row < r1) {
        d.lo = rowptr[row];
        d.len = rowptr[row + 1] - d.lo;
        d.bi = b[row];
        if (slot < d.len && d.len <= 8) { d.j = col[d.lo + slot]; d.v = val[d.lo + slot]; }
    }
    return d;
}
__global__ void __launch_bounds__(SMALL_T, 1) k_gs_small(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                         const double* __restrict__ val, const double* __restrict__ b, double* x, double w,
                                                         double one_minus_w, int sweeps, int* flags) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    volatile double* xs = reinterpret_cast<volatile double*>(sm_raw);
    volatile unsigned short* done = reinterpret_cast<volatile unsigned short*>(sm_raw + (size_t)n * sizeof(double));
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    for (int i = t; i < n; i += SMALL_T) { xs[i] = x[i]; done[i] = 0; }
    __syncthreads();
    const int nchunks = (n + GSS_CHUNK - 1) / GSS_CHUNK;
    bool bail = false;
    auto wait_for = [&](int j, int i, int r0, int r1, int sweep) {   // row j's value as the sequential sweep sees it at row i
        if (j >= r0 && j < r1) return;   // rows of this chunk are this warp's own: program order covers them
        const int need = (j < i) ? sweep : sweep - 1;
        long long spins = 0;
        while ((int)done[j] < need) {
            if (++spins > (1ll << 26)) { atomicOr(flags, DF_SPIN); bail = true; break; }
        }
        __threadfence_block();
    };
    for (int sweep = 1; sweep <= sweeps && !bail; ++sweep) {
        for (int chunk = wid; chunk < nchunks && !bail; chunk += SMALL_WARPS) {
            const int r0 = chunk * GSS_CHUNK, r1 = min(n, r0 + GSS_CHUNK);
            GsSub cur = gs_load_sub(rowptr, col, val, b, r0, r1, lane);
            for (int rs = r0; rs < r1 && !bail; rs += 4) {
                GsSub nxt = gs_load_sub(rowptr, col, val, b, rs + 4, r1, lane);   // in flight while the four rows below are solved
                for (int r = 0; r < 4; ++r) {
                    const int i = rs + r;
                    if (i >= r1) break;
                    const int src = 8 * r;
                    const int len = __shfl_sync(0xffffffffu, cur.len, src);
                    const double bi = __shfl_sync(0xffffffffu, cur.bi, src);
                    double sum = 0., aii = 0.;
                    bool have_diag = false;
                    if (len <= 8) {
                        double pr = 0.;
                        if ((lane >> 3) == r && cur.j >= 0 && cur.j != i) {
                            wait_for(cur.j, i, r0, r1, sweep);
                            pr = cur.v * xs[cur.j];
                        }
                        for (int l = 0; l < len; ++l) {  // ordered sum, identical in every lane
                            const double pv = __shfl_sync(0xffffffffu, pr, src + l);
                            const double vv = __shfl_sync(0xffffffffu, cur.v, src + l);
                            const int jj = __shfl_sync(0xffffffffu, cur.j, src + l);
                            if (jj != i) sum += pv; else { aii = vv; have_diag = true; }
                        }
                    } else {  // long row: lanes across the entries, direct loads (the path of k_gs_sweep)
                        const int lo = __shfl_sync(0xffffffffu, cur.lo, src), hi = lo + len;
                        for (int base = lo; base < hi; base += 32) {
                            const int k = base + lane;
                            double pr = 0., vk = 0.;
                            int j = -1;
                            if (k < hi) {
                                j = col[k]; vk = val[k];
                                if (j != i) { wait_for(j, i, r0, r1, sweep); pr = vk * xs[j]; }
                            }
                            const int cnt = min(32, hi - base);
                            for (int l = 0; l < cnt; ++l) {
                                const double pv = __shfl_sync(0xffffffffu, pr, l);
                                const double vv = __shfl_sync(0xffffffffu, vk, l);
                                const int jj = __shfl_sync(0xffffffffu, j, l);
                                if (jj != i) sum += pv; else { aii = vv; have_diag = true; }
                            }
                        }
                    }
                    bail = __any_sync(0xffffffffu, bail);
                    if (lane == 0) {
                        if (!have_diag) {
                            atomicOr(flags, DF_MISSING_ENTRY);
                        } else {
                            const double xi = xs[i] * one_minus_w + w * (bi - sum) / aii;
                            if (xi != xi) atomicOr(flags, DF_GS_NAN);
                            xs[i] = xi;
                        }
                        __threadfence_block();
                        done[i] = (unsigned short)sweep;
                    }
                    __syncwarp();
                    if (bail) break;
                }
                cur = nxt;
            }
        }
    }
    __syncthreads();
    for (int i = t; i < n; i += SMALL_T) x[i] = xs[i];
}

bool gs_small_ok(const Ctx& c, const DCsr& A, uint64_t sweeps) {
    const size_t need = (size_t)A.nrows * (sizeof(double) + sizeof(unsigned short));
    return small_enabled() && A.nrows == A.ncols && A.nrows > 0 && need <= SMALL_SMEM_MAX && sweeps < 65535 && A.sym == 1;
}

void gauss_seidel_small(Ctx& c, const DCsr& A, const double* b, double* x, double w, double one_minus_w, uint64_t sweeps) {
    const int n = (int)A.nrows;
    const size_t need = (size_t)n * (sizeof(double) + sizeof(unsigned short));
    static bool attr_set = false;
    if (!attr_set) {
        ORC_CUDA(cudaFuncSetAttribute(k_gs_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM_MAX));
        attr_set = true;
    }
    k_gs_small<<<1, SMALL_T, need, c.stream>>>(n, A.rowptr, A.col, A.val, b, x, w, one_minus_w, (int)sweeps, c.d_flags);
    c.after_launch("k_gs_small");
}

}  // namespace orc
