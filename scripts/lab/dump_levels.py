"""Dump the AMG coarse levels of the n^3 hex channel's momentum matrix (GPU-built hierarchy) as binary CSR files for the CUDA labs.
Usage: python scripts/lab/dump_levels.py [n] [outdir]. File: int64 n, int64 nnz, int32 rowptr[n+1], int32 col[nnz], double val[nnz]."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import orc_b200
from orc_b200 import synthetic as syn
from orc_b200 import discretization as disc, linear_algebra as la
from cases import smooth_fields

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
out = sys.argv[2] if len(sys.argv) > 2 else "/tmp"
mesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(n, n, n)))
syn.channel_bcs(mesh)
ctx = orc_b200.default_context()
u, v, w, p = smooth_fields(mesh.export())
a_di, *_ = disc.build_momentum_diffusion_matrix(mesh, 1e-3, ctx)
a = [disc.initialize_momentum_matrix(mesh, ctx) for _ in range(3)]
disc.build_momentum_advection_matrices(*a, a_di, mesh, u, v, w, p, orc_b200.NumericalSettings(), 1000.0)
x, levels = la.multigrid_trace(a[0], np.ones(mesh.n_cells), np.zeros(mesh.n_cells), iteration_count=1)
for l, (_, A) in enumerate(levels):
    rp, co, va = A.arrays()
    path = os.path.join(out, f"lvl{l + 1}.bin")
    with open(path, "wb") as f:
        np.array([A.dims[0], A.dims[2]], np.int64).tofile(f)
        rp.astype(np.int32).tofile(f); co.astype(np.int32).tofile(f); va.tofile(f)
    print(path, A.dims, flush=True)
