"""Device time of the greedy restriction (k_strongest_dataflow) and the Galerkin product per AMG level on differently shaped meshes.
Usage: python scripts/lab/restriction_shapes.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc_b200
from orc_b200 import synthetic as syn
from orc_b200 import discretization as disc, linear_algebra as la
from cases import smooth_fields

ctx = orc_b200.default_context()
CASES = [("hex 128^3", lambda: syn.hex_box(128, 128, 128), {}, False),
         ("tet 48^3", lambda: syn.tet_box(48, 48, 48), dict(momentum=orc_b200.MomentumDiscretization.TVD, limiter=orc_b200.TVD_UMIST), True),
         ("tet 80^3", lambda: syn.tet_box(80, 80, 80), dict(momentum=orc_b200.MomentumDiscretization.TVD, limiter=orc_b200.TVD_UMIST), True),
         ("tet 150x150x19 (one rank's slab of the 8-GPU config-5 run)", lambda: syn.tet_box(150, 150, 19, lz=0.001 * 19 / 150),
          dict(momentum=orc_b200.MomentumDiscretization.TVD, limiter=orc_b200.TVD_UMIST), True)]
CASES.append(("hex 256x256x32 slab (one rank's share of the 8-GPU weak-scaling box)", lambda: syn.hex_box(256, 256, 32, lx=0.008, ly=0.002, lz=0.00025), {}, False))
only = sys.argv[1:] 
for name, make, kw, f3d in CASES:
    if only and not any(o in name for o in only):
        continue
    mesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(make()))
    syn.channel_bcs(mesh, fully_3d=f3d)
    u, v, w, p = smooth_fields(mesh.export())
    a_di, *_ = disc.build_momentum_diffusion_matrix(mesh, 1e-3, ctx)
    a = [disc.initialize_momentum_matrix(mesh, ctx) for _ in range(3)]
    disc.build_momentum_advection_matrices(*a, a_di, mesh, u, v, w, p, orc_b200.NumericalSettings(**kw), 1000.0)
    A, _ = a[0].jacobi_scale(np.ones(mesh.n_cells))
    line = f"{name}: "
    for lvl in range(3):
        nr, _, nnz = A.dims
        la.bench_amg_setup(A, 2)                      # the two trial builds of the ticket-shape tuner (linalg.cu: DfrTune)
        t_r, t_g, Ac = la.bench_amg_setup(A, 3)
        line += f"[L{lvl}: {nr} rows, restriction {t_r:.2f} ms = {1e6 * t_r / nr:.1f} ns/row, galerkin {t_g:.2f} ms] "
        A = Ac
    print(line, flush=True)
    del mesh, a, a_di, A
