"""Profiling target: `iters` SIMPLE iterations on an n^3 hex channel (defaults of the reference). Used under ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orc_b200
from orc_b200 import synthetic as syn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(n, n, n)))
syn.channel_bcs(mesh)
st = orc_b200.SteadySolver(mesh, orc_b200.NumericalSettings(), 1000.0, 1e-3)
st.set_fields(*(np.zeros(mesh.n_cells) for _ in range(4)))
for k in range(iters):
    rep = st.iterate(1)
    print(k, rep["u_avg"], rep["pressure_correction"], st.phase_ms(), flush=True)
print("launches", orc_b200.default_context().launch_count())
