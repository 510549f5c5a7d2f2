// staged_lab.cu — lab harness for the row-block staged SpMV of the AMG coarse levels (round 2).
//
// The coarse-level gather kernels are bound by the L1 data pipe: a warp-wide gather costs ~2 cycles per distinct 128-byte line
// it touches (B300_MICROARCH.md "L1tex wavefront queue"), 0.5 (one system) to 0.7 (three systems) lines per stored entry. A block
// of R consecutive rows touches 5-10 x fewer distinct lines than its warp instructions do one by one, so the block stages the
// lines of x it needs in shared memory ONCE (cp.async, 16-byte chunks, whole 128-byte lines: perfectly coalesced), double
// buffered against the computation of the previous block, and the per-entry access becomes a shared-memory read through a
// 16-bit local index (10 instead of 12 bytes per entry from HBM).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o staged_lab staged_lab.cu
//   ./staged_lab level1.bin [R] [reps]
// File: int64 n, int64 nnz, int32 rowptr[n+1], int32 col[nnz], double val[nnz] (scripts/lab/dump_levels.py).
#include <cuda_runtime.h>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

// ---- baseline: the production gather kernel (linalg.cu k_spmv_vec, EP_NONE) ----
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
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o, G);
        if (i < n && gl == 0) Cell<K>::st(y, i, acc);
    }
}

// ---- staging plan: per block of R rows the sorted list of distinct 128-byte lines of x its entries touch, and per entry the 16-bit
// local index (line rank * CPL + column within the line). CPL = columns per line: 4 (32-byte cells, K = 3) or 16 (doubles, K = 1) ----
constexpr int PLAN_T = 256;
constexpr int BM_WORDS = 4096;   // bitmap words: a block's line span may reach 131 072 lines
template <int CPL>
__global__ void __launch_bounds__(PLAN_T) k_plan(int n, int R, const int* __restrict__ rowptr, const int* __restrict__ col, int cap_lines, int* __restrict__ nlines,
                                                 int* __restrict__ lines, unsigned short* __restrict__ lidx, int* __restrict__ stats /* [0] overflow, [1] max lines */) {
    __shared__ unsigned int bitmap[BM_WORDS];
    __shared__ int prefix[BM_WORDS];
    __shared__ int s_mn, s_mx, s_warp[PLAN_T / 32];
    const int t = threadIdx.x, b = blockIdx.x;
    const int r0 = b * R, r1 = min(n, r0 + R);
    const int lo = rowptr[r0], hi = rowptr[r1];
    if (t == 0) { s_mn = INT_MAX; s_mx = -1; }
    __syncthreads();
    int mn = INT_MAX, mx = -1;
    for (int k = lo + t; k < hi; k += PLAN_T) { const int l = col[k] / CPL; mn = min(mn, l); mx = max(mx, l); }
    for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((t & 31) == 0) { atomicMin(&s_mn, mn); atomicMax(&s_mx, mx); }
    __syncthreads();
    mn = s_mn; mx = s_mx;
    if (hi == lo) { if (t == 0) nlines[b] = 0; return; }
    const int span = mx - mn + 1;
    if (span > BM_WORDS * 32) { if (t == 0) { nlines[b] = -1; atomicExch(stats, 1); } return; }
    const int words = (span + 31) >> 5;
    for (int w = t; w < words; w += PLAN_T) bitmap[w] = 0u;
    __syncthreads();
    for (int k = lo + t; k < hi; k += PLAN_T) { const int l = col[k] / CPL - mn; atomicOr(&bitmap[l >> 5], 1u << (l & 31)); }
    __syncthreads();
    // exclusive prefix of the word popcounts: thread t owns words [t * per, (t + 1) * per)
    const int per = (words + PLAN_T - 1) / PLAN_T;
    int mine = 0;
    for (int w = t * per; w < min(words, (t + 1) * per); ++w) mine += __popc(bitmap[w]);
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((t & 31) >= o) incl += v; }
    if ((t & 31) == 31) s_warp[t >> 5] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < (t >> 5); ++w) wbase += s_warp[w];
    int total = 0;
    for (int w = 0; w < PLAN_T / 32; ++w) total += s_warp[w];
    int run = wbase + incl - mine;
    for (int w = t * per; w < min(words, (t + 1) * per); ++w) { prefix[w] = run; run += __popc(bitmap[w]); }
    __syncthreads();
    if (t == 0) { nlines[b] = total <= cap_lines ? total : -1; atomicMax(stats + 1, total); if (total > cap_lines) atomicExch(stats, 1); }
    if (total > cap_lines) return;
    for (int w = t; w < words; w += PLAN_T) {
        unsigned int bits = bitmap[w];
        int o = prefix[w];
        while (bits) { const int bit = __ffs(bits) - 1; lines[(size_t)b * cap_lines + o++] = mn + (w << 5) + bit; bits &= bits - 1; }
    }
    for (int k = lo + t; k < hi; k += PLAN_T) {
        const int c = col[k], l = c / CPL - mn;
        const int rank = prefix[l >> 5] + __popc(bitmap[l >> 5] & ((1u << (l & 31)) - 1u));
        lidx[k] = (unsigned short)(rank * CPL + c % CPL);
    }
}

// ---- the staged kernel ----
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <int K> struct StageGeom {
    static constexpr int CPL = (K == 1) ? 16 : 4;   // columns per 128-byte line
};
// 16-byte chunk q of the staged footprint -> its slot in shared memory. K = 3: the two halves of a 32-byte cell are swapped on
// odd lines, so that the first halves (x_u, x_v) of random cells spread over all eight 16-byte bank groups instead of four.
template <int K> __device__ __forceinline__ int swz(int q) { return K == 3 ? (q ^ ((q >> 3) & 1)) : q; }

template <int G, int UN, int K, int T, int IDS>
__global__ void __launch_bounds__(T, 2) k_staged(int n, int ncols, int R, int nblk, const int* __restrict__ rowptr, const unsigned short* __restrict__ lidx,
                                                 const double* __restrict__ val, int cap_lines, const int* __restrict__ nlines, const int* __restrict__ lines,
                                                 const double* __restrict__ x, double* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int CPL = StageGeom<K>::CPL;
    const int t = threadIdx.x, gl = t & (G - 1);
    const size_t buf_bytes = (size_t)cap_lines * 128;
    const char* xb = reinterpret_cast<const char*>(x);
    const long long xbytes = (long long)ncols * Cell<K>::S * 8;

    int ids[IDS];   // line ids of the chunks this thread copies for the NEXT block to be staged
    auto load_ids = [&](int b) {
        const int L = (b < nblk) ? nlines[b] : 0;
#pragma unroll
        for (int j = 0; j < IDS; ++j) {
            const int c = t + j * T;   // chunk index: line rank c >> 3, 16-byte chunk c & 7
            ids[j] = ((c >> 3) < L) ? lines[(size_t)b * cap_lines + (c >> 3)] : -1;
        }
    };
    auto stage = [&](unsigned char* buf) {
#pragma unroll
        for (int j = 0; j < IDS; ++j) {
            if (ids[j] < 0) continue;
            const int c = t + j * T;
            const long long off = (long long)ids[j] * 128 + (c & 7) * 16;
            if (off + 16 <= xbytes) cp_async16(buf + (size_t)swz<K>(c) * 16, xb + off);
            else if (off + 8 <= xbytes) cp_async8(buf + (size_t)swz<K>(c) * 16, xb + off);   // K = 1 only: the last double of an odd-length vector
        }
    };

    int b = blockIdx.x;
    load_ids(b);
    stage(smem);
    cp_async_commit();
    load_ids(b + gridDim.x);
    for (int it = 0; b < nblk; b += gridDim.x, ++it) {
        cp_async_wait_all();
        __syncthreads();   // block b is staged; every thread is done with the buffer the next stage overwrites
        unsigned char* cur = smem + (size_t)(it & 1) * buf_bytes;
        stage(smem + (size_t)((it + 1) & 1) * buf_bytes);   // block b + grid (ids were loaded one iteration ago)
        cp_async_commit();
        load_ids(b + 2 * gridDim.x);
        const int r0 = b * R, r1 = min(n, r0 + R);
        for (int i = r0 + t / G; i < r1; i += T / G) {
            const int lo = rowptr[i], hi = rowptr[i + 1];
            double acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = 0.;
            for (int q = lo + gl; q < hi; q += UN * G) {
                double v[UN], xv[UN][K]; int li[UN]; bool ok[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) { ok[u] = q + u * G < hi; v[u] = ok[u] ? val[q + u * G] : 0.; li[u] = ok[u] ? (int)lidx[q + u * G] : 0; }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    if (K == 1) {
                        xv[u][0] = reinterpret_cast<const double*>(cur)[li[u]];
                    } else {
                        const int q0 = 2 * li[u];
                        const double2 a = *reinterpret_cast<const double2*>(cur + (size_t)swz<K>(q0) * 16);
                        const double w = *reinterpret_cast<const double*>(cur + (size_t)swz<K>(q0 + 1) * 16);
                        xv[u][0] = a.x; xv[u][1 % K] = a.y; xv[u][2 % K] = w;
                    }
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) if (ok[u]) {
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] += v[u] * xv[u][k];
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o, G);
            if (gl == 0) Cell<K>::st(y, i, acc);
        }
    }
    cp_async_wait_all();
}

// ---- v2: everything the compute phase touches is in shared memory. Per row block the producer thread fetches the block's slice of
// (val, lidx, rowptr) with three bulk copies (cp.async.bulk + mbarrier: the slice of a row block is contiguous), all threads fetch
// the footprint lines of x with 16-byte cp.async, both one block ahead of the computation (two stages). One block per SM. ----
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar)) : "memory");
}
struct Stage2 {   // byte offsets of one stage inside the dynamic shared memory
    size_t xf, val, lidx, rp, bytes;
};
__host__ __device__ inline Stage2 stage2_layout(int cap_lines, int ecap, int R) {
    Stage2 s;
    s.xf = 0;
    s.val = (size_t)cap_lines * 128;
    s.lidx = s.val + (size_t)(ecap + 2) * 8;
    s.rp = s.lidx + (((size_t)(ecap + 16) * 2 + 15) & ~(size_t)15);
    s.bytes = (s.rp + (((size_t)(R + 4) * 4 + 15) & ~(size_t)15) + 127) & ~(size_t)127;
    return s;
}
template <int G, int UN, int K, int T, int IDS>
__global__ void __launch_bounds__(T, 1) k_staged2(int n, int ncols, int R, int nblk, const int* __restrict__ rowptr, const unsigned short* __restrict__ lidx,
                                                  const double* __restrict__ val, int cap_lines, int ecap, const int* __restrict__ nlines,
                                                  const int* __restrict__ lines, const double* __restrict__ x, double* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full[2];
    const Stage2 L = stage2_layout(cap_lines, ecap, R);
    const int t = threadIdx.x, gl = t & (G - 1);
    const char* xb = reinterpret_cast<const char*>(x);
    const long long xbytes = (long long)ncols * Cell<K>::S * 8;
    if (t == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();

    int ids[IDS];
    auto load_ids = [&](int b) {
        const int Lb = (b < nblk) ? nlines[b] : 0;
#pragma unroll
        for (int j = 0; j < IDS; ++j) { const int c = t + j * T; ids[j] = ((c >> 3) < Lb) ? lines[(size_t)b * cap_lines + (c >> 3)] : -1; }
    };
    auto stage = [&](int b, int s) {   // block b -> stage s
        unsigned char* base = smem + (size_t)s * L.bytes;
#pragma unroll
        for (int j = 0; j < IDS; ++j) {
            if (ids[j] < 0) continue;
            const int c = t + j * T;
            const long long off = (long long)ids[j] * 128 + (c & 7) * 16;
            if (off + 16 <= xbytes) cp_async16(base + L.xf + (size_t)swz<K>(c) * 16, xb + off);
            else if (off + 8 <= xbytes) cp_async8(base + L.xf + (size_t)swz<K>(c) * 16, xb + off);
        }
        if (t == 0 && b < nblk) {
            const int r0 = b * R, r1 = min(n, r0 + R);
            const int lo = rowptr[r0], hi = rowptr[r1];
            const int v0 = lo & ~1, v1 = (hi + 1) & ~1;      // 16-byte aligned slice of val
            const int l0 = lo & ~7, l1 = (hi + 7) & ~7;      // ... of lidx
            const unsigned int bv = (unsigned int)(v1 - v0) * 8, bl = (unsigned int)(l1 - l0) * 2, br = (unsigned int)((r1 - r0 + 1 + 3) & ~3) * 4;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&full[s], bv + bl + br);
            if (bv) bulk_g2s(base + L.val, val + v0, bv, &full[s]);
            if (bl) bulk_g2s(base + L.lidx, lidx + l0, bl, &full[s]);
            bulk_g2s(base + L.rp, rowptr + r0, br, &full[s]);
        }
    };

    int b = blockIdx.x;
    load_ids(b);
    stage(b, 0);
    cp_async_commit();
    load_ids(b + gridDim.x);
    for (int it = 0; b < nblk; b += gridDim.x, ++it) {
        const int s = it & 1;
        cp_async_wait_all();
        mbar_wait(&full[s], (unsigned int)((it >> 1) & 1));
        __syncthreads();   // block b is complete in stage s; every thread is done with stage s ^ 1
        stage(b + gridDim.x, s ^ 1);
        cp_async_commit();
        load_ids(b + 2 * gridDim.x);
        const unsigned char* base = smem + (size_t)s * L.bytes;
        const unsigned char* cur = base + L.xf;
        const double* sval = reinterpret_cast<const double*>(base + L.val);
        const unsigned short* slid = reinterpret_cast<const unsigned short*>(base + L.lidx);
        const int* srp = reinterpret_cast<const int*>(base + L.rp);
        const int r0 = b * R, r1 = min(n, r0 + R);
        const int lo0 = srp[0];
        const int vbase = lo0 & ~1, lbase = lo0 & ~7;
        for (int i = r0 + t / G; i < r1; i += T / G) {
            const int lo = srp[i - r0], hi = srp[i - r0 + 1];
            double acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = 0.;
            for (int q = lo + gl; q < hi; q += UN * G) {
                double v[UN], xv[UN][K]; int li[UN]; bool ok[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) { ok[u] = q + u * G < hi; v[u] = ok[u] ? sval[q + u * G - vbase] : 0.; li[u] = ok[u] ? (int)slid[q + u * G - lbase] : 0; }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    if (K == 1) {
                        xv[u][0] = reinterpret_cast<const double*>(cur)[li[u]];
                    } else {
                        const int q0 = 2 * li[u];
                        const double2 a = *reinterpret_cast<const double2*>(cur + (size_t)swz<K>(q0) * 16);
                        const double w = *reinterpret_cast<const double*>(cur + (size_t)swz<K>(q0 + 1) * 16);
                        xv[u][0] = a.x; xv[u][1 % K] = a.y; xv[u][2 % K] = w;
                    }
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) if (ok[u]) {
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] += v[u] * xv[u][k];
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o, G);
            if (gl == 0) Cell<K>::st(y, i, acc);
        }
    }
    cp_async_wait_all();
}

struct Mat { int n; long long nnz; int *rp, *col; double* val; std::vector<int> hrp; };
static Mat load(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { printf("cannot open %s\n", path); exit(1); }
    long long n, nnz;
    if (fread(&n, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1) exit(1);
    std::vector<int> rp(n + 1), col(nnz); std::vector<double> val(nnz);
    if (fread(rp.data(), 4, n + 1, f) != (size_t)n + 1 || fread(col.data(), 4, nnz, f) != (size_t)nnz || fread(val.data(), 8, nnz, f) != (size_t)nnz) exit(1);
    fclose(f);
    Mat m; m.n = (int)n; m.nnz = nnz; m.hrp = rp;
    CK(cudaMalloc(&m.rp, 4 * (n + 1) + 64)); CK(cudaMalloc(&m.col, 4 * nnz + 64)); CK(cudaMalloc(&m.val, 8 * nnz + 64));
    CK(cudaMemcpy(m.rp, rp.data(), 4 * (n + 1), cudaMemcpyHostToDevice)); CK(cudaMemcpy(m.col, col.data(), 4 * nnz, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.val, val.data(), 8 * nnz, cudaMemcpyHostToDevice));
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

template <int K, int G>
static void run(const Mat& m, int R, int reps, int cap_lines) {
    constexpr int S = Cell<K>::S, UN = (K == 1) ? 4 : 2, CPL = StageGeom<K>::CPL, T = 512;
    const int n = m.n;
    const size_t nx = (size_t)n * S;
    double *x, *y0, *y1;
    CK(cudaMalloc(&x, 8 * nx)); CK(cudaMalloc(&y0, 8 * nx)); CK(cudaMalloc(&y1, 8 * nx));
    std::vector<double> hx(nx);
    for (size_t i = 0; i < nx; ++i) hx[i] = 0.001 * (double)((i * 2654435761u) % 1000) - 0.5;
    CK(cudaMemcpy(x, hx.data(), 8 * nx, cudaMemcpyHostToDevice));
    CK(cudaMemset(y0, 0, 8 * nx)); CK(cudaMemset(y1, 0xff, 8 * nx));
    const int nblk = (n + R - 1) / R;
    int *nlines, *lines, *stats; unsigned short* lidx;
    CK(cudaMalloc(&nlines, 4 * nblk)); CK(cudaMalloc(&lines, 4 * (size_t)nblk * cap_lines)); CK(cudaMalloc(&lidx, 2 * m.nnz + 64)); CK(cudaMalloc(&stats, 8));
    CK(cudaMemset(stats, 0, 8));
    float us_plan = timeit([&] { k_plan<CPL><<<nblk, PLAN_T>>>(n, R, m.rp, m.col, cap_lines, nlines, lines, lidx, stats); }, 3);
    int hstats[2];
    CK(cudaMemcpy(hstats, stats, 8, cudaMemcpyDeviceToHost));
    std::vector<int> hnl(nblk);
    CK(cudaMemcpy(hnl.data(), nlines, 4 * nblk, cudaMemcpyDeviceToHost));
    long long tot = 0; for (int v : hnl) tot += v > 0 ? v : 0;
    printf("  K=%d G=%d R=%d: plan %.0f us, lines/entry %.4f, max lines/block %d (cap %d)%s\n", K, G, R, us_plan, (double)tot / m.nnz, hstats[1], cap_lines,
           hstats[0] ? "  ** OVERFLOW: staged kernel skipped **" : "");
    const double alg = 12. * m.nnz + 4. * n + 16. * K * n;
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int RPB = 256 / G, ngroups = (n + RPB - 1) / RPB;
    const int gg = std::min(ngroups, sms * (K == 1 ? 8 : 6));
    float us_g = timeit([&] { k_gather<G, UN, K><<<gg, 256>>>(n, m.rp, m.col, m.val, x, y0); }, reps);
    printf("     gather: %.1f us = %.0f GB/s\n", us_g, alg / us_g * 1e-3);
    if (hstats[0]) return;
    constexpr int IDS = 6;
    std::vector<double> h0(nx), h1(nx);
    CK(cudaMemcpy(h0.data(), y0, 8 * nx, cudaMemcpyDeviceToHost));
    if (cap_lines * 8 <= IDS * T) {
        const size_t smem = 2 * (size_t)cap_lines * 128;
        auto kern = k_staged<G, UN, K, T, IDS>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
        const int grid = std::min(nblk, sms * std::max(occ, 1));
        float us_s = timeit([&] { kern<<<grid, T, smem>>>(n, n, R, nblk, m.rp, lidx, m.val, cap_lines, nlines, lines, x, y1); }, reps);
        CK(cudaMemcpy(h1.data(), y1, 8 * nx, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < nx; ++i) if (memcmp(&h0[i], &h1[i], 8) != 0) ++bad;
        printf("     staged: %.1f us = %.0f GB/s (%d CTAs/SM, grid %d, smem %zu KB)  %s (%zu of %zu differ)\n", us_s, alg / us_s * 1e-3, occ, grid, smem / 1024,
               bad ? "MISMATCH" : "bit-identical to gather", bad, nx);
    }
    // ---- v2 ----
    {
        constexpr int T2 = 1024, IDS2 = 8;   // cap_lines * 8 <= IDS2 * T2
        int maxe = 0;
        for (int b = 0; b < nblk; ++b) { const int r0 = b * R, r1 = std::min(n, r0 + R); maxe = std::max(maxe, m.hrp[r1] - m.hrp[r0]); }
        const int ecap = (maxe + 16 + 15) & ~15;
        const int cap2 = std::min(cap_lines, hstats[1] + 0);
        const Stage2 L = stage2_layout(cap_lines, ecap, R);
        const size_t smem2 = 2 * L.bytes;
        if (smem2 > 220 * 1024 || cap_lines * 8 > IDS2 * T2) { printf("     v2: does not fit (%zu KB, ecap %d)\n", smem2 / 1024, ecap); }
        else {
            auto k2 = k_staged2<G, UN, K, T2, IDS2>;
            CK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            CK(cudaMemset(y1, 0xff, 8 * nx));
            const int grid2 = std::min(nblk, sms);
            float us2 = timeit([&] { k2<<<grid2, T2, smem2>>>(n, n, R, nblk, m.rp, lidx, m.val, cap_lines, ecap, nlines, lines, x, y1); }, reps);
            CK(cudaMemcpy(h1.data(), y1, 8 * nx, cudaMemcpyDeviceToHost));
            size_t bad2 = 0;
            for (size_t i = 0; i < nx; ++i) if (memcmp(&h0[i], &h1[i], 8) != 0) ++bad2;
            printf("     v2    : %.1f us = %.0f GB/s (1 CTA/SM x %d threads, smem %zu KB, ecap %d)  %s (%zu differ)\n", us2, alg / us2 * 1e-3, T2, smem2 / 1024, ecap,
                   bad2 ? "MISMATCH" : "bit-identical to gather", bad2);
            (void)cap2;
        }
    }
    cudaFree(x); cudaFree(y0); cudaFree(y1); cudaFree(nlines); cudaFree(lines); cudaFree(lidx); cudaFree(stats);
}

int main(int argc, char** argv) {
    if (argc < 2) { printf("usage: staged_lab matrix.bin [reps]\n"); return 1; }
    const int reps = argc > 2 ? atoi(argv[2]) : 50;
    Mat m = load(argv[1]);
    const double avg = (double)m.nnz / m.n;
    printf("%s: %d rows, %lld entries (%.1f per row)\n", argv[1], m.n, m.nnz, avg);
    const int cap = argc > 3 ? atoi(argv[3]) : 384;
    for (int R : {32, 64, 128, 256}) {
        if (avg < 32.) { run<1, 4>(m, R, reps, cap); run<3, 4>(m, R, reps, cap); }
        else { run<1, 8>(m, R, reps, cap); run<3, 8>(m, R, reps, cap); }
    }
    return 0;
}
