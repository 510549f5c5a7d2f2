"""Multi-GPU worker (run under torchrun, one rank per GPU): solves the same case partitioned over the ranks and on rank 0
alone, and prints the deviations as one JSON line. Used by tests/test_gpu_multi.py and by hand:

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc_b200  # noqa: E402
from orc_b200 import synthetic as syn  # noqa: E402
from orc_b200.settings import SolutionMethod, VelocityInterpolation  # noqa: E402

RHO, MU = 1000.0, 1e-3


def gather_owned(local, info, n_global, device):
    """all ranks' owned values -> the global vector (on every rank)."""
    world = dist.get_world_size()
    sizes = [None] * world
    dist.all_gather_object(sizes, (info["g0"], info["g1"]))
    out = np.zeros(n_global)
    for r, (g0, g1) in enumerate(sizes):
        t = torch.zeros(g1 - g0, dtype=torch.float64, device=device)
        if r == dist.get_rank():
            t.copy_(torch.from_numpy(local))
        dist.broadcast(t, r)
        out[g0:g1] = t.cpu().numpy()
    return out


def run_case(ctx, gmesh, settings, iters, device):
    rank, world = dist.get_rank(), dist.get_world_size()
    part = gmesh.partition(rank, world)
    info = part.partition_info()
    st = orc_b200.SteadySolver(part, settings, RHO, MU, ctx)
    st.set_fields(*(np.zeros(info["n_own"]) for _ in range(4)))
    rep = st.iterate(iters)
    fields = [gather_owned(f, info, info["n_global"], device) for f in st.get_fields()]
    st.close()
    return fields, rep


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = orc_b200.Context(local_rank)
    ctx.comm_init()
    shape = tuple(int(x) for x in os.environ.get("ORC_MGPU_SHAPE", "12,8,8").split(","))
    gmesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(*shape)))
    syn.channel_bcs(gmesh)
    n = gmesh.n_cells
    out = {"world": world, "cells": n}
    cases = {
        # no recurrence in the assembly and a global BiCGSTAB: the partitioned run equals the single-GPU run up to the
        # summation order of the dot products
        "bicgstab_linear_weighted": (orc_b200.NumericalSettings(velocity_interpolation=VelocityInterpolation.LinearWeighted,
                                                                  matrix_solver=orc_b200.MatrixSolverSettings(solver_type=SolutionMethod.BiCGSTAB)), 3),
        # reference defaults: Rhie-Chow (partition-lagged diagonals, C4) + Multigrid (per-partition aggregates)
        "multigrid_rhie_chow": (orc_b200.NumericalSettings(), 3),
    }
    only = os.environ.get("ORC_MGPU_CASE")
    single = orc_b200.Context(local_rank) if rank == 0 else None   # a context without a communicator: the single-GPU path
    for name, (settings, iters) in cases.items():
        if only and name != only:
            continue
        fields, rep = run_case(ctx, gmesh, settings, iters, device)
        res = {"u_avg": rep["u_avg"], "p_corr": rep["pressure_correction"], "finite": bool(all(np.isfinite(f).all() for f in fields))}
        if rank == 0:
            u, v, w, p = (np.zeros(n) for _ in range(4))
            reps = []
            orc_b200.solve_steady(gmesh, u, v, w, p, settings, RHO, MU, iters, 1, ctx=single, on_report=reps.append)
            vel = np.sqrt(sum(np.linalg.norm(f) ** 2 for f in (u, v, w)))
            res["dev_vs_single"] = {c: float(np.linalg.norm(a - b) / (vel if c != "p" else np.linalg.norm(b)))
                                    for c, a, b in zip("uvwp", fields, (u, v, w, p))}
            res["u_avg_single"] = reps[-1]["u_avg"]
            res["p_corr_single"] = reps[-1]["pressure_correction"]
            np.savez(os.path.join(ROOT, "gpurun_out", f"mgpu_{name}_w{world}.npz") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else
                     f"/tmp/mgpu_{name}_w{world}.npz", u=fields[0], v=fields[1], w=fields[2], p=fields[3])
        out[name] = res
        dist.barrier()
    # the partitioned run against the ORACLE with its partition emulation (lagged diagonals at the cut + per-block aggregates)
    from mgpu_check import partition_parity
    par = partition_parity(ctx, device, shape=shape, iters=3)
    if rank == 0:
        out["oracle_partition_parity"] = par
        print("MGPU_RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
