"""Per-level timings on the REAL AMG hierarchy of an n^3 hex channel: SpMV, one BiCGSTAB iteration, restriction build,
Galerkin product. Usage: python scripts/level_bench.py [n] [reps]. Prints one line per level (algorithmic GB/s for SpMV)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orc_b200
from orc_b200 import synthetic as syn
from orc_b200 import discretization as disc, linear_algebra as la

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(n, n, n)))
syn.channel_bcs(mesh)
ctx = orc_b200.default_context()
a_di, bu, bv, bw = disc.build_momentum_diffusion_matrix(mesh, 1e-3)
a0, b0 = a_di.jacobi_scale(bu)   # what multigrid_solve sees on the fine level (linear_algebra.rs:157-168)


def timed(f, k=3):
    best = 1e30
    out = None
    for _ in range(k):
        ctx.synchronize()
        t = time.perf_counter()
        out = f()
        ctx.synchronize()
        best = min(best, time.perf_counter() - t)
    return best * 1e3, out


A = a0
lvl = 0
while True:
    nr, nc, nnz = A.dims
    alg = 12.0 * nnz + 20.0 * nr
    ms = la.bench_spmv(A, reps)
    msb = la.bench_bicgstab(A, reps)
    print(f"level {lvl}: rows {nr} nnz {nnz} ({nnz / nr:.1f}/row) alg {alg / 1e6:.1f} MB  spmv {ms * 1e3:.1f} us = {alg / ms / 1e6:.0f} GB/s"
          f"  bicgstab-iter {msb * 1e3:.1f} us (2 spmv + 3 vec; spmv share {2 * ms / msb:.2f})", flush=True)
    alg3 = 12.0 * nnz + 4.0 * nr + 48.0 * nr
    ms3 = la.bench_spmv(A, reps, systems=3)
    msb3 = la.bench_bicgstab(A, reps, systems=3)
    print(f"         3 systems in lockstep: alg {alg3 / 1e6:.1f} MB  spmv {ms3 * 1e3:.1f} us = {alg3 / ms3 / 1e6:.0f} GB/s ({3 * ms / ms3:.2f}x vs 3 launches)"
          f"  bicgstab-iter {msb3 * 1e3:.1f} us ({3 * msb / msb3:.2f}x vs 3 solves)", flush=True)
    for k in (1, 3):   # split of one BiCGSTAB iteration by kernel class (events around every launch: adds a few us per launch)
        ctx.prof_enable(True)
        la.bench_bicgstab(A, reps, systems=k)
        pr = ctx.prof_get()
        ctx.prof_enable(False)
        print(f"         K={k} per iteration: spmv {pr['spmv'][0] / (reps + 3) * 1e3:.1f} us ({pr['spmv'][2]} launches), "
              f"vector {pr['vector'][0] / (reps + 3) * 1e3:.1f} us ({pr['vector'][2]} launches)", flush=True)
    if lvl == 3:
        break
    t_r, t_g, Ac = la.bench_amg_setup(A, 3)
    print(f"   setup of the next level (device time): restriction {t_r:.2f} ms  galerkin {t_g:.2f} ms", flush=True)
    A = Ac   # the recursion coarsens the UNSCALED product (linear_algebra.rs:110); the smoother scales its own copy
    lvl += 1
