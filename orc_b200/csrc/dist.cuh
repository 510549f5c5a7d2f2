// dist.cuh — multi-GPU plumbing of the SIMPLE loop: one process per GPU, NCCL over NVLink/NVSwitch.
// C1 halo exchange of cell fields (ncclSend/ncclRecv, one group per exchange), C2 small allreduces for the BiCGSTAB
// scalars and the iteration scalars, C4 partition-lagged recurrences. AMG below the fine level is built per partition.
#pragma once
#include "linalg.cuh"
#include "mesh_host.hpp"

namespace orc {

// NCCL is resolved at run time (dlopen "libnccl.so.2"): a single-GPU process never loads it, and inside a torch process
// the already-loaded NCCL of torch is reused.
struct Comm {
    int rank = 0, nranks = 1;
    void* comm = nullptr;  // ncclComm_t
    bool active() const { return nranks > 1 && comm != nullptr; }
    static void unique_id(char out128[128]);
    void init(Ctx& c, int rank_, int nranks_, const char id128[128]);
    void destroy();
    void allreduce(Ctx& c, double* dev, int count, int op /* 0 sum, 2 max, 3 min */);
};

// device-side exchange plan of one partition (from PartPlan)
struct Halo {
    int64_t n_loc = 0, own_lo = 0, own_hi = 0;
    std::vector<int> nbr, send_ptr, recv_begin, recv_count;
    DBuf<int> send_idx;       // local ids of owned cells to pack, all neighbours back to back
    DBuf<double> sendbuf;     // kMaxFields * total send count
    static constexpr int kMaxFields = 4;
    void build(Ctx& c, const PartPlan& p);
    // exchanges up to kMaxFields vectors of length n_loc in one NCCL group: owned values -> the neighbours' halo slots
    void exchange(Ctx& c, Comm& comm, double* const* fields, int nfields);
    void exchange(Ctx& c, Comm& comm, double* f0) { double* f[1] = {f0}; exchange(c, comm, f, 1); }
    // one vector of `stride`-double cells (stride 1 = plain vector, 4 = a batch of three systems, linalg.cu Cell<3>)
    void exchange_cells(Ctx& c, Comm& comm, double* x, int stride);
};

struct DistEnv {  // what a distributed solve needs besides the matrices
    Comm* comm = nullptr;
    Halo* halo = nullptr;
    bool on() const { return comm && halo && comm->active(); }
};

// iterative_solve over a row-partitioned matrix (rows = local cells, halo rows empty; columns local ids incl. halo).
// BiCGSTAB runs globally (halo exchange before every SpMV, allreduce for every scalar); Multigrid = global BiCGSTAB
// pre-smoothing + the reference's multigrid_solve applied to THIS rank's diagonal block (partition-local aggregates).
void iterative_solve_dist(Ctx& c, DistEnv& env, DCsr& a, const double* b, double* x, const SolveParams& sp, MgTrace* trace, int K = 1);

}  // namespace orc
