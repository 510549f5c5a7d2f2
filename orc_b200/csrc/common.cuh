// common.cuh — context, error plumbing, stream-ordered device buffers, deterministic block reductions.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/orc_b200.h"

namespace orc {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define ORC_CUDA(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
            throw ::orc::Error(ORC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                                               std::to_string(__LINE__) + ")");                                \
    } while (0)

#define ORC_REQUIRE(cond, code, msg)                      \
    do {                                                  \
        if (!(cond)) throw ::orc::Error((code), (msg));   \
    } while (0)

// Device-side status word shared by all kernels of a context (bit flags, sticky until cleared).
enum DevFlag : int {
    DF_NAN_JACOBI = 1,       // linear_algebra.rs:192-196
    DF_JACOBI_HUGE = 2,      // linear_algebra.rs:214-216
    DF_GS_NAN = 4,           // linear_algebra.rs:240-242
    DF_MG_NAN = 8,           // linear_algebra.rs:103-105
    DF_SPIN = 16,            // dataflow kernel exceeded its spin bound
    DF_MISSING_ENTRY = 32,   // CsrMatrix::get on an un-stored entry (lib.rs:664-666)
    DF_UNSUPPORTED_BC = 64,  // face zone type outside {2,3,4,5,7,10} reached on the path
    DF_CONVERGED = 128,      // Jacobi convergence latch (not an error)
    DF_SINGULAR = 256        // least-squares gradient: try_inverse().unwrap() on a singular normal matrix (solver.rs:855, 945)
};

// Per-kernel-class device timing (CUDA events on the context stream around every launch of the class), switched on by
// bench.py for the roofline leg: achieved bytes/s = sum of algorithmic bytes / sum of event durations.
// PC_BICG brackets a WHOLE BiCGSTAB call (all iterations of one solve on one level) with ONE event pair: the launches inside run
// back to back, unperturbed by events, so this is the in-situ throughput of the solver loop. It is never sub-sampled.
enum ProfClass : int { PC_SPMV = 0, PC_VECTOR, PC_ASSEMBLY, PC_RESTRICT, PC_GALERKIN, PC_SCALE, PC_OTHER, PC_BICG, PC_COUNT };
struct KernelProf {
    bool enabled = false;
    unsigned class_mask = ~0u;   // classes that are timed (bit = ProfClass)
    unsigned sample_every = 1;   // time every n-th launch of a class only (events cost ~5 us each on short kernels)
    uint64_t seen[8] = {};
    struct Rec { cudaEvent_t a, b; int cls; double bytes, ref_bytes; int64_t key; };
    struct Detail { double ms = 0, bytes = 0; uint64_t count = 0; };
    std::map<int64_t, Detail> detail;   // SpMV launches by (entries of the matrix) * 8 + systems per launch
    std::map<int64_t, int64_t> rows_of; // entries -> rows of that matrix
    std::vector<Rec> pool;
    size_t used = 0;
    double ms[PC_COUNT] = {}, bytes[PC_COUNT] = {}, ref_bytes[PC_COUNT] = {};  // ref_bytes: the same launches counted in the reference's units
    uint64_t count[PC_COUNT] = {};
};

struct Ctx {
    KernelProf prof;
    bool exact_order = false;  // set per solve from orc_settings.reduction_mode
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    uint64_t launches = 0;
    int* d_flags = nullptr;          // device status word
    double* d_scal = nullptr;        // 64 device scalars (BiCGSTAB rho/alpha/omega/beta, norms, ...)
    double* d_partials = nullptr;    // block partial sums, 4 lanes x kMaxBlocks
    unsigned int* d_counter = nullptr;  // last-block-done counters (zeroed, self-resetting)
    cudaMemPool_t pool = nullptr;
    static constexpr int kMaxBlocks = 4096;
    // Fused reductions are organised in VIRTUAL blocks: the rows (or elements) a partial sum covers, and the order in which
    // partials are combined, depend on the problem size only — not on how many thread blocks a kernel variant can keep
    // resident. A launch maps virtual blocks round-robin onto its real blocks; 3552 = lcm(8, 6) * 148 divides evenly for
    // kernels that fit 8 or 6 blocks per SM. Results are therefore identical across kernel variants (1 or 3 systems per
    // launch) and across devices.
    static constexpr int kVirtualBlocks = 3552;
    static constexpr int kPartialLanes = 8;

    // Caching device allocator. Every solve re-creates the AMG hierarchy (the reference rebuilds it per call), whose buffer
    // sizes change slightly from solve to solve; all work is issued on ONE stream, so a freed block can be handed out again
    // immediately (stream order protects it). Sizes are rounded up to 1/8 of the enclosing power of two so that nearby
    // sizes share blocks; a request takes the smallest cached block that wastes at most 25 %.
    std::multimap<size_t, void*> cache_free;
    std::unordered_map<void*, size_t> cache_live;
    size_t cache_bytes = 0;
    static size_t round_size(size_t b) {
        if (b < 512) return 512;
        size_t p2 = 1;
        while (p2 < b) p2 <<= 1;
        size_t step = p2 >> 4;  // b is in (p2/2, p2]: 8 classes
        return (b + step - 1) / step * step;
    }
    void cache_release_all() {
        cudaStreamSynchronize(stream);
        for (auto& kv : cache_free) { cudaFree(kv.second); cache_bytes -= kv.first; }
        cache_free.clear();
    }
    void* alloc(size_t bytes) {
        const size_t r = round_size(bytes);
        auto it = cache_free.lower_bound(r);
        if (it != cache_free.end() && it->first <= r + r / 4) {
            void* p = it->second;
            cache_live[p] = it->first;
            cache_free.erase(it);
            return p;
        }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, r);
        if (e != cudaSuccess) {  // give the cached blocks back to the driver and retry once
            cudaGetLastError();
            cache_release_all();
            e = cudaMalloc(&p, r);
        }
        if (e != cudaSuccess) throw Error(ORC_E_CUDA, std::string("cudaMalloc(") + std::to_string(r) + " bytes): " + cudaGetErrorString(e));
        cache_live[p] = r;
        cache_bytes += r;
        return p;
    }
    void free(void* p) {
        if (!p) return;
        auto it = cache_live.find(p);
        if (it == cache_live.end()) return;
        cache_free.emplace(it->second, p);
        cache_live.erase(it);
    }
    template <class T>
    T* alloc_n(size_t n) { return static_cast<T*>(alloc(n * sizeof(T))); }
    void sync() { ORC_CUDA(cudaStreamSynchronize(stream)); }
    int read_flags() {
        int f = 0;
        ORC_CUDA(cudaMemcpyAsync(&f, d_flags, sizeof(int), cudaMemcpyDeviceToHost, stream));
        sync();
        return f;
    }
    void clear_flags() { ORC_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int), stream)); }
    int prof_begin(int cls, double bytes, double ref_bytes = -1., int64_t key = -1) {
        if (!prof.enabled) return -1;
        if (!((prof.class_mask >> cls) & 1u)) return -1;
        if (prof.sample_every > 1 && cls != PC_BICG && (prof.seen[cls]++ % prof.sample_every) != 0) return -1;
        if (prof.used == prof.pool.size()) {
            KernelProf::Rec r;
            ORC_CUDA(cudaEventCreate(&r.a));
            ORC_CUDA(cudaEventCreate(&r.b));
            prof.pool.push_back(r);
        }
        KernelProf::Rec& r = prof.pool[prof.used];
        r.cls = cls; r.bytes = bytes; r.ref_bytes = ref_bytes < 0. ? bytes : ref_bytes; r.key = key;
        ORC_CUDA(cudaEventRecord(r.a, stream));
        return (int)prof.used++;
    }
    void prof_end(int id) {
        if (id >= 0) ORC_CUDA(cudaEventRecord(prof.pool[id].b, stream));
    }
    void prof_resolve() {  // call after a stream synchronisation point
        if (!prof.used) return;
        ORC_CUDA(cudaStreamSynchronize(stream));
        for (size_t k = 0; k < prof.used; ++k) {
            float t = 0.f;
            cudaEventElapsedTime(&t, prof.pool[k].a, prof.pool[k].b);
            prof.ms[prof.pool[k].cls] += t; prof.bytes[prof.pool[k].cls] += prof.pool[k].bytes; prof.ref_bytes[prof.pool[k].cls] += prof.pool[k].ref_bytes;
            prof.count[prof.pool[k].cls]++;
            if (prof.pool[k].key >= 0) { auto& d = prof.detail[prof.pool[k].key]; d.ms += t; d.bytes += prof.pool[k].bytes; d.count++; }
        }
        prof.used = 0;
    }
    void prof_reset() {
        prof_resolve();
        for (int k = 0; k < PC_COUNT; ++k) { prof.ms[k] = 0; prof.bytes[k] = 0; prof.ref_bytes[k] = 0; prof.count[k] = 0; }
        prof.detail.clear(); prof.rows_of.clear();
    }
    void after_launch(const char* what) {
        ++launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) throw Error(ORC_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    }
};

// RAII stream-ordered buffer
template <class T>
struct DBuf {
    Ctx* ctx = nullptr;
    T* p = nullptr;
    size_t n = 0;
    DBuf() {}
    DBuf(Ctx* c, size_t n_) : ctx(c), n(n_) { p = c->alloc_n<T>(n_); }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : ctx(o.ctx), p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) { reset(); ctx = o.ctx; p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DBuf() { reset(); }
    void reset() { if (p && ctx) ctx->free(p); p = nullptr; n = 0; }
    void alloc(Ctx* c, size_t n_) { reset(); ctx = c; n = n_; p = c->alloc_n<T>(n_); }
    void zero() { if (n) ORC_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), ctx->stream)); }
    void upload(const T* h) { if (n) ORC_CUDA(cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream)); }
    void download(T* h) const { if (n) ORC_CUDA(cudaMemcpyAsync(h, p, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream)); }
    operator T*() const { return p; }
};

// Where the unknowns of a matrix sit in space. A Multigrid solve uses it to STORE its coarse levels along a space-filling curve
// (linalg.cu: "locality ordering"): the gather SpMVs of those levels are bound by the distinct 128-byte lines a warp instruction
// touches, and the aggregates' own numbering scatters a row's columns over 2-3 x more lines than a Morton order does.
// Row I of the matrix is at (x, y, z)[min(I << shift, n - 1)]: a coarse row of the "Strongest" restriction collects the pushes of
// the fine rows 2I and 2I+1 (linear_algebra.rs:52-53), so its place is that of fine row 2I, `shift` levels up.
struct PosHint {
    const double *x = nullptr, *y = nullptr, *z = nullptr;  // device planes of the mesh mirror (or the matrix's own copy)
    int64_t n = 0;
    int shift = 0;
    double lo[3] = {0., 0., 0.}, inv[3] = {0., 0., 0.};     // quantisation: (pos - lo) * inv lies in [0, 1024)
    bool on() const { return x != nullptr && n > 0; }
};

// Device CSR (nalgebra-sparse CsrMatrix<f64>): int32 offsets/indices, fp64 values. The pattern arrays
// may be shared between matrices (all five mesh matrices share one pattern); `val` is always owned.
struct DCsr {
    Ctx* ctx = nullptr;
    int64_t nrows = 0, ncols = 0, nnz = 0;
    int* rowptr = nullptr;
    int* col = nullptr;
    int* diag = nullptr;   // index of the stored (i,i) entry or -1; built on demand
    double* val = nullptr;
    bool own_pattern = true;
    bool own_diag = true;
    int sym = -1;          // structural symmetry: -1 unknown, 0 no, 1 yes
    int full_diag = -1;    // every row stores its diagonal: -1 unknown
    int max_row = -1;      // longest row: -1 unknown (filled by the Galerkin product)
    int simplex = -1;      // the FINE matrix this one descends from has rows of <= 5 entries (triangles / tetrahedra): -1 unknown.
                           // Picks the ticket shape of the restriction kernel (linalg.cu: dfr_rows_for)
    PosHint hint;          // optional: positions of the unknowns
    double* hint_own = nullptr;  // 3 n doubles behind `hint` when the matrix owns them (handles detached from their mesh)
    ~DCsr() {
        if (!ctx) return;
        if (own_pattern) { ctx->free(rowptr); ctx->free(col); }
        if (own_diag) ctx->free(diag);
        ctx->free(val);
        ctx->free(hint_own);
    }
};

// ---- deterministic reductions --------------------------------------------------------------------
// Block partial sums are combined in a fixed order (lane-strided serial sums, then a shuffle tree), so a
// given grid size always produces the same bits. The order differs from nalgebra's 8-accumulator dot
// (SURVEY.md §8c); that difference is part of the documented solve tolerance, not of assembly parity.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
// sum over the block; result valid in thread 0. `sh` must hold 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.;
    if (wid == 0) {
        r = (lane < (int)((blockDim.x + 31) >> 5)) ? sh[lane] : 0.;
        r = warp_sum(r);
    }
    return r;
}
// NQ sums at once behind one barrier pair; the tree of every quantity is block_sum's. Valid in thread 0. `sh`: 32 * NQ doubles.
template <int NQ>
__device__ __forceinline__ void block_sum_n(double (&v)[NQ], double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NQ; ++q) v[q] = warp_sum(v[q]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) sh[q * 32 + wid] = v[q];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            double r = (lane < (int)((blockDim.x + 31) >> 5)) ? sh[q * 32 + lane] : 0.;
            v[q] = warp_sum(r);
        }
    }
}
__device__ __forceinline__ double block_max(double v, double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = -INFINITY;
    if (wid == 0) {
        r = (lane < (int)((blockDim.x + 31) >> 5)) ? sh[lane] : -INFINITY;
        r = warp_max(r);
    }
    return r;
}
// Returns true in every thread of the LAST block to arrive (counter self-resets for the next launch).
__device__ __forceinline__ bool last_block_done(unsigned int* counter) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) *counter = 0u;
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}
// Fixed-order sum of `n` partials by one block (call from the last block only). Valid in thread 0.
__device__ __forceinline__ double sum_partials(const double* part, int n, double* sh) {
    double a = 0.;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += __ldcg(part + i);
    return block_sum(a, sh);
}
__device__ __forceinline__ double max_partials(const double* part, int n, double* sh) {
    double a = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a = fmax(a, __ldcg(part + i));
    return block_max(a, sh);
}

// ---- virtual-block reductions (Ctx::kVirtualBlocks) --------------------------------------------------------------
// A real block works through its virtual blocks without a block barrier in between (a barrier per virtual block drains
// the load pipeline three or four times per launch: measured +7 us on a 40 us SpMV): every warp parks its partial of
// virtual block `lv` in shared memory; after the loop each virtual block's warp partials are combined with block_sum's
// second-stage tree — the published value is bit-identical to block_sum over that virtual block.
constexpr int kMaxLocalVb = 24;   // virtual blocks per real block (3552 / 148 at one resident block per SM)
constexpr int kWarpsPerBlock = 8; // all reducing kernels run 256-thread blocks
template <int NQ>
__device__ __forceinline__ void vb_park(double* park, int lv, double (&v)[NQ]) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const double r = warp_sum(v[q]);
        if (lane == 0) park[(lv * NQ + q) * kWarpsPerBlock + wid] = r;
    }
}
template <int NQ>
__device__ __forceinline__ void vb_publish(const double* park, int n_local, int first_vb, int vb_stride, double* partials) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    for (int item = wid; item < n_local * NQ; item += kWarpsPerBlock) {
        const int lv = item / NQ, q = item - lv * NQ;
        double r = (lane < kWarpsPerBlock) ? park[item * kWarpsPerBlock + lane] : 0.;
        r = warp_sum(r);
        if (lane == 0) partials[q * Ctx::kMaxBlocks + first_vb + lv * vb_stride] = r;
    }
}
// Fixed-order totals of NQ partial lanes (last block only): thread t adds elements t, t + 256, ... in ascending order — the
// loads of a batch of eight are issued together — then one block tree for all quantities. Valid in thread 0.
template <int NQ>
__device__ __forceinline__ void sum_partials_n(const double* partials, int n, double (&T)[NQ], double* sh) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const double* part = partials + q * Ctx::kMaxBlocks;
        double a = 0.;
        for (int base = threadIdx.x; base < n; base += 8 * (int)blockDim.x) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int i = base + u * (int)blockDim.x; v[u] = (i < n) ? __ldcg(part + i) : 0.; }
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int i = base + u * (int)blockDim.x; if (i < n) a += v[u]; }
        }
        T[q] = a;
    }
    block_sum_n<NQ>(T, sh);
}

struct ProfScope {  // RAII: times everything enqueued on the stream during its lifetime as one record of class `cls`
    Ctx& c;
    int id;
    ProfScope(Ctx& c_, int cls, double bytes, double ref_bytes = -1., int64_t key = -1) : c(c_), id(c_.prof_begin(cls, bytes, ref_bytes, key)) {}
    ~ProfScope() { try { c.prof_end(id); } catch (...) {} }
};

inline int grid_for(int64_t n, int block, int cap = Ctx::kMaxBlocks) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

}  // namespace orc
