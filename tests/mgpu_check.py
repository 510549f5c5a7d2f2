"""Partitioned-run parity check shared by tests/mgpu_worker.py and bench.py --gpus N (checker leg, outside the timed region):
a small hex channel is solved partitioned over all ranks on the GPUs, and on rank 0 by the oracle with its partition emulation
(oracle.set_partition: diagonals of cells in another partition lag by one exchange, Multigrid coarse correction per partition
block — the two documented deviations of the multi-GPU path) and without it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RHO, MU = 1000.0, 1e-3


def gather_owned(local, info, n_global, device):
    """all ranks' owned values -> the global vector (on every rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    sizes = [None] * world
    dist.all_gather_object(sizes, (info["g0"], info["g1"]))
    out = np.zeros(n_global)
    for r, (g0, g1) in enumerate(sizes):
        t = torch.zeros(g1 - g0, dtype=torch.float64, device=device)
        if r == dist.get_rank():
            t.copy_(torch.from_numpy(np.ascontiguousarray(local)))
        dist.broadcast(t, r)
        out[g0:g1] = t.cpu().numpy()
    return out, [s[0] for s in sizes] + [sizes[-1][1]]


def partition_parity(ctx, device, shape=None, iters=3, settings=None, oracle_settings=None, inner_iterations=None):
    """Returns (on rank 0) {"cells", "world", "vs_oracle_partitioned": {u, v, w, p}, "vs_oracle_single": {...}}: relative L2
    deviations of the partitioned GPU fields (velocity components against the norm of the velocity field, p against ||p||)."""
    import torch.distributed as dist
    import orc_b200
    from orc_b200 import synthetic as syn
    rank, world = dist.get_rank(), dist.get_world_size()
    shape = shape or (12, 8, 4 * world)
    arrays = syn.hex_box(*shape)
    gmesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(arrays))
    syn.channel_bcs(gmesh)
    settings = settings or orc_b200.NumericalSettings()
    if inner_iterations:
        settings.matrix_solver.iterations = inner_iterations
    part = gmesh.partition(rank, world)
    info = part.partition_info()
    st = orc_b200.SteadySolver(part, settings, RHO, MU, ctx)
    st.set_fields(*(np.zeros(info["n_own"]) for _ in range(4)))
    st.iterate(iters)
    fields, cuts = [], None
    for f in st.get_fields():
        g, cuts = gather_owned(f, info, info["n_global"], device)
        fields.append(g)
    st.close()
    if rank != 0:
        return None
    from oracle import pyoracle as po
    om = po.Mesh.from_arrays(*syn.mesh_args(arrays))
    syn.channel_bcs(om)
    n = om.n_cells
    z = np.zeros(n)
    out = {"cells": n, "world": world, "iterations": iters, "cuts": [int(c) for c in cuts]}
    os_ = oracle_settings or po.Settings()
    if inner_iterations:
        os_.iterations = inner_iterations
    out["inner_iterations"] = int(os_.iterations)
    def dev(got, ref):
        vel = np.sqrt(sum(np.linalg.norm(b) ** 2 for b in ref[:3]))
        return {name: float(np.linalg.norm(a - b) / (vel if name != "p" else np.linalg.norm(b))) for name, a, b in zip("uvwp", got, ref)}

    refs = {}
    for key, c, rho in (("vs_oracle_partitioned", cuts, RHO), ("vs_oracle_single", None, RHO), ("perturbed", cuts, RHO * (1.0 + 2.3e-16))):
        po.set_partition(c)
        try:
            refs[key] = om.solve_steady(z, z, z, z, os_, rho, MU, iters, 0)[:4]
        finally:
            po.set_partition(None)
    out["vs_oracle_partitioned"] = dev(fields, refs["vs_oracle_partitioned"])
    out["vs_oracle_single"] = dev(fields, refs["vs_oracle_single"])
    # the yardstick: how far the ORACLE moves when the density changes by one ulp (the reference's unguarded solvers amplify rounding)
    out["oracle_sensitivity_to_1ulp_of_rho"] = dev(refs["perturbed"], refs["vs_oracle_partitioned"])
    return out
