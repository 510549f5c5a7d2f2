import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orc_b200
from orc_b200 import synthetic as syn
n = int(sys.argv[1]); iters = int(sys.argv[2])
mesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(n, n, n)))
syn.channel_bcs(mesh)
for pr in [float(x) for x in sys.argv[3:]]:
    s = orc_b200.NumericalSettings(pressure_relaxation=pr)
    st = orc_b200.SteadySolver(mesh, s, 1000.0, 1e-3)
    st.set_fields(*(np.zeros(mesh.n_cells) for _ in range(4)))
    out = []
    try:
        for k in range(iters):
            r = st.iterate(1); out.append(f"{r['pressure_correction']:.1e}/{r['u_avg']:.3e}")
    except orc_b200.OrcError as e:
        out.append(str(e)[:40])
    print(n, pr, " ".join(out), flush=True)
    st.close()
