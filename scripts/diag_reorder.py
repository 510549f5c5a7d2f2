"""Trajectory of the bench workload with the coarse-level locality ordering off / on: per-iteration ||p'|| and mean u, until the
reference algorithm's own divergence. Usage: python scripts/diag_reorder.py [n] [iterations]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orc_b200
from orc_b200 import synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
mesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(n, n, n)))
syn.channel_bcs(mesh)
for mode in ("0", "1", "2", "3"):
    os.environ["ORC_B200_REORDER"] = mode
    s = orc_b200.NumericalSettings(pressure_relaxation=1e-4)
    st = orc_b200.SteadySolver(mesh, s, 1000.0, 1e-3)
    st.set_fields(*(np.zeros(mesh.n_cells) for _ in range(4)))
    out = []
    try:
        for k in range(iters):
            r = st.iterate(1); out.append(f"{r['pressure_correction']:.6e}/{r['u_avg']:.6e}")
    except orc_b200.OrcError as e:
        out.append(str(e)[:40])
    print(f"n={n} reorder={mode}:", " ".join(out), flush=True)
    del st
