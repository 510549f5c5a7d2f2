"""Dump the AMG coarse levels of the n^3 hex channel's momentum matrix (GPU-built hierarchy) as binary CSR files for the CUDA labs.
Usage: python scripts/lab/dump_levels.py [n] [outdir] [--morton]. File: int64 n, int64 nnz, int32 rowptr[n+1], int32 col[nnz], double val[nnz].
--morton also writes lvl<l>_morton.bin: the same matrix with rows and columns renumbered along the Morton code of the aggregate
positions (position of coarse row I = position of the fine row 2 I, quantised to 10 bits per axis on the bounding box), columns
sorted within the rows (scripts/lab/reorder_lab.cu)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import orc_b200
from orc_b200 import synthetic as syn
from orc_b200 import discretization as disc, linear_algebra as la
from cases import smooth_fields

MORTON = "--morton" in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if len(args) > 0 else 128
out = args[1] if len(args) > 1 else "/tmp"
mesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(n, n, n)))
syn.channel_bcs(mesh)
ctx = orc_b200.default_context()
u, v, w, p = smooth_fields(mesh.export())
a_di, *_ = disc.build_momentum_diffusion_matrix(mesh, 1e-3, ctx)
a = [disc.initialize_momentum_matrix(mesh, ctx) for _ in range(3)]
disc.build_momentum_advection_matrices(*a, a_di, mesh, u, v, w, p, orc_b200.NumericalSettings(), 1000.0)
x, levels = la.multigrid_trace(a[0], np.ones(mesh.n_cells), np.zeros(mesh.n_cells), iteration_count=1)
def write(path, nrows, rp, co, va):
    with open(path, "wb") as f:
        np.array([nrows, co.size], np.int64).tofile(f)
        rp.astype(np.int32).tofile(f); co.astype(np.int32).tofile(f); va.tofile(f)
    print(path, nrows, co.size, flush=True)


def morton_bbox(c, bits=10):
    lo, hi = c.min(axis=0), c.max(axis=0)
    q = np.minimum(((c - lo) / (hi - lo + 1e-300) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    key = np.zeros(c.shape[0], np.int64)
    for b in range(bits):
        for a in range(3):
            key |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return key


cent = mesh.export()["cell_centroid"]
for l, (_, A) in enumerate(levels):
    rp, co, va = A.arrays()
    write(os.path.join(out, f"lvl{l + 1}.bin"), A.dims[0], rp, co, va)
    if MORTON:
        import scipy.sparse as sp
        m = A.dims[0]
        cent = cent[np.minimum(2 * np.arange(m), cent.shape[0] - 1)]
        perm = np.argsort(morton_bbox(cent), kind="stable")
        B = sp.csr_matrix((va, co, rp), shape=(m, m))[perm][:, perm].tocsr()
        B.sort_indices()
        write(os.path.join(out, f"lvl{l + 1}_morton.bin"), m, B.indptr, B.indices, B.data)
