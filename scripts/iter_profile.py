"""Per-iteration kernel-class times (CUDA events) for `iters` SIMPLE iterations on an n^3 hex channel."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orc_b200
from orc_b200 import synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(n, n, n)))
syn.channel_bcs(mesh)
ctx = orc_b200.default_context()
st = orc_b200.SteadySolver(mesh, orc_b200.NumericalSettings(), 1000.0, 1e-3)
st.set_fields(*(np.zeros(mesh.n_cells) for _ in range(4)))
prev = {k: 0.0 for k in ("momentum_assembly", "momentum_solves", "pressure_assembly", "pressure_solve", "correction")}
for k in range(iters):
    ctx.prof_enable(True)
    t0 = time.perf_counter()
    rep = st.iterate(1)
    wall = time.perf_counter() - t0
    pr = ctx.prof_get()
    ph = st.phase_ms()
    d = {a: round(ph[a] - prev[a], 1) for a in ph}
    prev = ph
    print(f"iter {k}: wall {wall*1e3:.0f} ms | classes " + " ".join(f"{a}={v[0]:.0f}ms/{v[2]}" for a, v in pr.items()) + f" | phases {d} | levels {st.level_sizes()} | u_avg {rep['u_avg']:.3e} |p'| {rep['pressure_correction']:.3e}", flush=True)
