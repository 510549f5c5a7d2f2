// dist.cu — NCCL communicator (resolved with dlopen) and the halo exchange of a row-partitioned mesh.
#include "dist.cuh"

#include <dlfcn.h>

#include <cstring>

namespace orc {

namespace {
struct ncclUniqueId_ { char internal[128]; };
typedef int (*fn_get_uid)(ncclUniqueId_*);
typedef int (*fn_init_rank)(void**, int, ncclUniqueId_, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_group)(void);
typedef int (*fn_sendrecv)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);
struct Nccl {
    void* lib = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    fn_sendrecv send = nullptr, recv = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_errstr errstr = nullptr;
};
Nccl& nccl() {
    static Nccl n;
    if (n.lib) return n;
    n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!n.lib) n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.lib) n.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!n.lib) throw Error(ORC_E_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* name) {
        void* p = dlsym(n.lib, name);
        if (!p) throw Error(ORC_E_NCCL, std::string("libnccl lacks ") + name);
        return p;
    };
    n.get_uid = (fn_get_uid)sym("ncclGetUniqueId");
    n.init_rank = (fn_init_rank)sym("ncclCommInitRank");
    n.destroy = (fn_destroy)sym("ncclCommDestroy");
    n.group_start = (fn_group)sym("ncclGroupStart");
    n.group_end = (fn_group)sym("ncclGroupEnd");
    n.send = (fn_sendrecv)sym("ncclSend");
    n.recv = (fn_sendrecv)sym("ncclRecv");
    n.allreduce = (fn_allreduce)sym("ncclAllReduce");
    n.errstr = (fn_errstr)sym("ncclGetErrorString");
    return n;
}
void nccl_check(int rc, const char* what) {
    if (rc != 0) throw Error(ORC_E_NCCL, std::string(what) + ": " + nccl().errstr(rc));
}
constexpr int kNcclFloat64 = 8;  // ncclDataType_t::ncclFloat64
}  // namespace

void Comm::unique_id(char out128[128]) {
    ncclUniqueId_ id;
    nccl_check(nccl().get_uid(&id), "ncclGetUniqueId");
    memcpy(out128, id.internal, 128);
}
void Comm::init(Ctx& c, int rank_, int nranks_, const char id128[128]) {
    ORC_REQUIRE(nranks_ >= 1 && rank_ >= 0 && rank_ < nranks_, ORC_E_INVALID, "bad rank / nranks");
    rank = rank_; nranks = nranks_;
    if (nranks == 1) return;
    ncclUniqueId_ id;
    memcpy(id.internal, id128, 128);
    ORC_CUDA(cudaSetDevice(c.device));
    nccl_check(nccl().init_rank(&comm, nranks, id, rank), "ncclCommInitRank");
}
void Comm::destroy() {
    if (comm) nccl().destroy(comm);
    comm = nullptr;
}
void Comm::allreduce(Ctx& c, double* dev, int count, int op) {
    if (!active()) return;
    ProfScope ps(c, PC_OTHER, 0.);
    nccl_check(nccl().allreduce(dev, dev, (size_t)count, kNcclFloat64, op, comm, c.stream), "ncclAllReduce");
    ++c.launches;
}

// -------------------------------------------------------------------------------------------------
__global__ void k_halo_pack(int count, const int* __restrict__ idx, const double* __restrict__ x, double* __restrict__ out) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) out[k] = x[idx[k]];
}

// cells of four doubles (three systems + padding): one 32-byte gather per packed cell
struct alignas(32) HaloCell { double a, b, c, d; };
__global__ void k_halo_pack4(int count, const int* __restrict__ idx, const double* __restrict__ x, double* __restrict__ out) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x)
        reinterpret_cast<HaloCell*>(out)[k] = reinterpret_cast<const HaloCell*>(x)[idx[k]];
}

void Halo::build(Ctx& c, const PartPlan& p) {
    n_loc = p.n_lo + p.n_own + p.n_hi;
    own_lo = p.n_lo; own_hi = p.n_lo + p.n_own;
    nbr.assign(p.nbr_rank.begin(), p.nbr_rank.end());
    send_ptr.assign(p.send_ptr.begin(), p.send_ptr.end());
    recv_begin.assign(p.recv_begin.begin(), p.recv_begin.end());
    recv_count.assign(p.recv_count.begin(), p.recv_count.end());
    const size_t ns = p.send_idx.size();
    send_idx.alloc(&c, std::max<size_t>(ns, 1));
    if (ns) ORC_CUDA(cudaMemcpyAsync(send_idx.p, p.send_idx.data(), ns * sizeof(int), cudaMemcpyHostToDevice, c.stream));
    sendbuf.alloc(&c, std::max<size_t>(ns * kMaxFields, 1));
    c.sync();
}

void Halo::exchange(Ctx& c, Comm& comm, double* const* fields, int nfields) {
    if (!comm.active() || nbr.empty()) return;
    ORC_REQUIRE(nfields >= 1 && nfields <= kMaxFields, ORC_E_INVALID, "halo exchange: too many fields");
    ProfScope ps(c, PC_OTHER, 0.);
    const int ns = send_ptr.back();
    for (int f = 0; f < nfields; ++f) {
        k_halo_pack<<<grid_for(std::max(ns, 1), 256, c.sm_count * 4), 256, 0, c.stream>>>(ns, send_idx, fields[f], sendbuf.p + (size_t)f * ns);
        c.after_launch("k_halo_pack");
    }
    Nccl& n = nccl();
    nccl_check(n.group_start(), "ncclGroupStart");
    for (int f = 0; f < nfields; ++f)
        for (size_t q = 0; q < nbr.size(); ++q) {
            const int cnt = send_ptr[q + 1] - send_ptr[q];
            if (cnt > 0) nccl_check(n.send(sendbuf.p + (size_t)f * ns + send_ptr[q], (size_t)cnt, kNcclFloat64, nbr[q], comm.comm, c.stream), "ncclSend");
            if (recv_count[q] > 0) nccl_check(n.recv(fields[f] + recv_begin[q], (size_t)recv_count[q], kNcclFloat64, nbr[q], comm.comm, c.stream), "ncclRecv");
        }
    nccl_check(n.group_end(), "ncclGroupEnd");
    ++c.launches;
}

void Halo::exchange_cells(Ctx& c, Comm& comm, double* x, int stride) {
    if (stride == 1) { exchange(c, comm, x); return; }
    ORC_REQUIRE(stride == 4 && kMaxFields >= 4, ORC_E_INTERNAL, "halo exchange: cells hold 1 or 4 doubles");
    if (!comm.active() || nbr.empty()) return;
    ProfScope ps(c, PC_OTHER, 0.);
    const int ns = send_ptr.back();
    k_halo_pack4<<<grid_for(std::max(ns, 1), 256, c.sm_count * 4), 256, 0, c.stream>>>(ns, send_idx, x, sendbuf.p);
    c.after_launch("k_halo_pack4");
    Nccl& n = nccl();
    nccl_check(n.group_start(), "ncclGroupStart");
    for (size_t q = 0; q < nbr.size(); ++q) {
        const int cnt = send_ptr[q + 1] - send_ptr[q];
        if (cnt > 0) nccl_check(n.send(sendbuf.p + 4 * (size_t)send_ptr[q], 4 * (size_t)cnt, kNcclFloat64, nbr[q], comm.comm, c.stream), "ncclSend");
        if (recv_count[q] > 0) nccl_check(n.recv(x + 4 * (size_t)recv_begin[q], 4 * (size_t)recv_count[q], kNcclFloat64, nbr[q], comm.comm, c.stream), "ncclRecv");
    }
    nccl_check(n.group_end(), "ncclGroupEnd");
    ++c.launches;
}

}  // namespace orc
