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

// Lexicographic Gauss-Seidel / SOR (the intended formula of src/linear_algebra.rs:219-246), all sweeps in ONE launch of one block,
// LEVEL SCHEDULED: level(i) = 1 + max level(j) over the stored lower neighbours j < i. For a structurally symmetric pattern the
// rows of one level do not store each other and every upper neighbour of a row sits in a later level, so running the levels in
// order — rows of a level in parallel, a block barrier between levels — gives every row exactly the values the sequential sweep
// sees (new x_j for j < i, old for j > i): bit-identical to k_gs_sweep and to the oracle. On the reference's couette mesh
// (y-fastest numbering, 127 x 63 cells) that is 189 levels of ~42 rows instead of a chain of 8 001 dependent rows; the ticketed
// dataflow sweep needs ~2 000 cycles per row there because every hop waits for the matrix row from L2.
// The levels are found in the same launch (monotone relaxation in shared memory over each thread's cached lower neighbours,
// as many rounds as there are levels), rows are bucketed by level, and a thread fetches the matrix row of its NEXT level into
// registers before the barrier of the current one, so the per-level critical path is shared-memory reads, the ordered row sum
// and one division. x lives in shared memory for the whole solve.
constexpr int GSL_T = 512;             // threads: two prefetched rows (8 entries each) per thread need ~100 registers
constexpr int GSL_LOWER = 4;           // lower neighbours cached in registers during the level search (more: re-read from global)
constexpr int GSL_SLOTS = 8;           // row entries prefetched into registers (longer rows: direct loads)
struct GslRow {
    int i, len, lo;
    int j[GSL_SLOTS];
    double v[GSL_SLOTS];
    double bi;
};
__device__ __forceinline__ void gsl_fetch(GslRow& r, int i, const int* __restrict__ rowptr, const int* __restrict__ col,
                                          const double* __restrict__ val, const double* __restrict__ b) {
    r.i = i;
    r.len = 0; r.lo = 0; r.bi = 0.;
    if (i < 0) return;
    r.lo = rowptr[i];
    r.len = rowptr[i + 1] - r.lo;
    r.bi = b[i];
    if (r.len <= GSL_SLOTS) {
#pragma unroll
        for (int q = 0; q < GSL_SLOTS; ++q) {
            const bool ok = q < r.len;
            r.j[q] = ok ? col[r.lo + q] : -1;
            r.v[q] = ok ? val[r.lo + q] : 0.;
        }
    }
}
// x_i of the sweep from the prefetched row (ascending k: the ordered sum of the reference)
__device__ __forceinline__ void gsl_solve(const GslRow& r, const int* __restrict__ col, const double* __restrict__ val, double* xs, double w,
                                          double one_minus_w, int* flags) {
    const int i = r.i;
    double sum = 0., aii = 0.;
    bool have_diag = false;
    if (r.len <= GSL_SLOTS) {
#pragma unroll
        for (int q = 0; q < GSL_SLOTS; ++q) {
            if (q < r.len) {
                if (r.j[q] != i) sum += r.v[q] * xs[r.j[q]]; else { aii = r.v[q]; have_diag = true; }
            }
        }
    } else {
        for (int q = r.lo; q < r.lo + r.len; ++q) {
            const int j = col[q];
            const double v = val[q];
            if (j != i) sum += v * xs[j]; else { aii = v; have_diag = true; }
        }
    }
    if (!have_diag) {
        atomicOr(flags, DF_MISSING_ENTRY);
    } else {
        const double xi = xs[i] * one_minus_w + w * (r.bi - sum) / aii;
        if (xi != xi) atomicOr(flags, DF_GS_NAN);
        xs[i] = xi;
    }
}
__global__ void __launch_bounds__(GSL_T, 1) k_gs_levels(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                        const double* __restrict__ val, const double* __restrict__ b, double* x, double w,
                                                        double one_minus_w, int sweeps, int* flags) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    double* xs = reinterpret_cast<double*>(sm_raw);                       // n doubles
    int* order = reinterpret_cast<int*>(xs + n);                          // n ints: rows bucketed by level
    int* lvl = order + n;                                                 // n ints: level of row i (1-based)
    int* start = lvl + n;                                                 // n + 2 ints: first position of level L in `order`
    __shared__ int s_nlev;
    const int t = threadIdx.x;
    for (int i = t; i < n; i += GSL_T) xs[i] = x[i];
    if (t == 0) s_nlev = 0;
    // ---- levels, batch by batch in row order: the lower neighbours of a batch's rows are final or inside the batch, so a
    // monotone relaxation inside the batch (levels only grow, any update order converges) settles in as many rounds as the
    // longest dependency chain inside the batch; every thread keeps the lower neighbours of its one row in registers ----
    for (int base = 0; base < n; base += GSL_T) {
        const int i = base + t;
        int low0 = -1, low1 = -1, low2 = -1, low3 = -1, nlow = 0;
        if (i < n) {
            for (int q = rowptr[i]; q < rowptr[i + 1]; ++q) {
                const int j = col[q];
                if (j < i) {
                    if (nlow == 0) low0 = j; else if (nlow == 1) low1 = j; else if (nlow == 2) low2 = j; else if (nlow == 3) low3 = j;
                    ++nlow;
                }
            }
            lvl[i] = 1;
        }
        __syncthreads();
        for (int round = 0; round <= GSL_T; ++round) {
            int changed = 0;
            if (i < n && nlow > 0) {
                volatile int* lv = lvl;
                int L = 1;
                if (nlow <= GSL_LOWER) {
                    if (low0 >= 0) L = max(L, 1 + lv[low0]);
                    if (low1 >= 0) L = max(L, 1 + lv[low1]);
                    if (low2 >= 0) L = max(L, 1 + lv[low2]);
                    if (low3 >= 0) L = max(L, 1 + lv[low3]);
                } else {
                    for (int q = rowptr[i]; q < rowptr[i + 1]; ++q) { const int j = col[q]; if (j < i) L = max(L, 1 + lv[j]); }
                }
                if (L > lv[i]) { lv[i] = L; changed = 1; }
            }
            if (!__syncthreads_or(changed)) break;
        }
    }
    // ---- bucket the rows by level (counting sort with an atomic cursor per level: any order inside a level is valid, its rows are
    // independent, so the result does not depend on it) ----
    for (int i = t; i < n + 2; i += GSL_T) start[i] = 0;
    __syncthreads();
    int mx = 0;
    for (int i = t; i < n; i += GSL_T) { atomicAdd(&start[lvl[i] + 1], 1); mx = max(mx, lvl[i]); }
    atomicMax(&s_nlev, mx);
    __syncthreads();
    const int nlev = s_nlev;
    // the count of level L sits in start[L + 1]: an inclusive running sum turns start[L] into the first position of level L
    if (t == 0) { int run = 0; for (int L = 1; L <= nlev + 1; ++L) { run += start[L]; start[L] = run; } }
    __syncthreads();
    for (int i = t; i < n; i += GSL_T) order[atomicAdd(&start[lvl[i]], 1)] = i;
    __syncthreads();
    // the cursors now hold the END of each level == the start of the next one: shift back by one level
    if (t == 0) { int prev = 0; for (int L = 1; L <= nlev; ++L) { const int e = start[L]; start[L] = prev; prev = e; } start[nlev + 1] = prev; }
    __syncthreads();
    // ---- the sweeps ----
    GslRow cur, nxt;
    auto row_of = [&](int L, int k) { const int p = start[L] + k; return (L <= nlev && p < start[L + 1]) ? order[p] : -1; };
    gsl_fetch(cur, row_of(1, t), rowptr, col, val, b);
    for (int sweep = 0; sweep < sweeps; ++sweep) {
        for (int L = 1; L <= nlev; ++L) {
            // the row this thread solves in the next level (the next sweep's first level after the last one), fetched before it is needed
            const int Ln = (L < nlev) ? L + 1 : 1;
            gsl_fetch(nxt, (L < nlev || sweep + 1 < sweeps) ? row_of(Ln, t) : -1, rowptr, col, val, b);
            if (cur.i >= 0) gsl_solve(cur, col, val, xs, w, one_minus_w, flags);
            for (int k = t + GSL_T; k < start[L + 1] - start[L]; k += GSL_T) {   // levels wider than the block: direct fetches
                gsl_fetch(cur, row_of(L, k), rowptr, col, val, b);
                gsl_solve(cur, col, val, xs, w, one_minus_w, flags);
            }
            __syncthreads();
            cur = nxt;
        }
    }
    for (int i = t; i < n; i += GSL_T) x[i] = xs[i];
}

static size_t gs_levels_smem(int64_t n) { return (size_t)n * (sizeof(double) + 3 * sizeof(int)) + 2 * sizeof(int); }
bool gs_small_ok(const Ctx& c, const DCsr& A, uint64_t sweeps) {
    return small_enabled() && A.nrows == A.ncols && A.nrows > 0 && gs_levels_smem(A.nrows) <= SMALL_SMEM_MAX && sweeps < (1u << 30) && A.sym == 1;
}

void gauss_seidel_small(Ctx& c, const DCsr& A, const double* b, double* x, double w, double one_minus_w, uint64_t sweeps) {
    const int n = (int)A.nrows;
    static bool attr_set = false;
    if (!attr_set) {
        ORC_CUDA(cudaFuncSetAttribute(k_gs_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM_MAX));
        attr_set = true;
    }
    k_gs_levels<<<1, GSL_T, gs_levels_smem(n), c.stream>>>(n, A.rowptr, A.col, A.val, b, x, w, one_minus_w, (int)sweeps, c.d_flags);
    c.after_launch("k_gs_levels");
}

}  // namespace orc
