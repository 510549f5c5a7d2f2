import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orc_b200
from orc_b200 import synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
def run(tag, iters=6, shape=None, ext=(0.004, 0.001, 0.001), **kw):
    shape = shape or (n, n, n)
    mesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(*shape, *ext)))
    syn.channel_bcs(mesh)
    s = orc_b200.NumericalSettings()
    for k, v in kw.items():
        if hasattr(s.matrix_solver, k): setattr(s.matrix_solver, k, v)
        else: setattr(s, k, v)
    st = orc_b200.SteadySolver(mesh, s, 1000.0, 1e-3)
    st.set_fields(*(np.zeros(mesh.n_cells) for _ in range(4)))
    out = []
    try:
        for k in range(iters):
            r = st.iterate(1); out.append(f"{r['pressure_correction']:.1e}/{r['u_avg']:.2e}")
    except orc_b200.OrcError as e:
        out.append(str(e)[:40])
    print(tag, shape, ext, kw, " ".join(out), flush=True)
run("default")
run("it100", iterations=100)
run("it200", iterations=200)
run("prelax1e-3", pressure_relaxation=1e-3)
run("cube", ext=(0.001, 0.001, 0.001))
run("long", shape=(4 * n, n // 2, n // 2))
run("frozen", assembly_mode=orc_b200.settings.AssemblyMode.Frozen)
