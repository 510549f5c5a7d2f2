"""Diagnostic: per-field deviation of the GPU solve_steady from the oracle (not a test)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import orc_b200
from orc_b200 import synthetic as syn
from oracle import pyoracle as po
from cases import make_pair, settings_pair

def run(arrays, iters, bcs, **kw):
    pm, om = make_pair(po, arrays)
    for m in (pm, om): bcs(m)
    ps, os_ = settings_pair(po, **kw)
    n = pm.n_cells
    u, v, w, p = (np.zeros(n) for _ in range(4))
    orc_b200.solve_steady(pm, u, v, w, p, ps, 1000.0, 1e-3, iters, 0)
    uo, vo, wo, pp, _, _ = om.solve_steady(np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n), os_, 1000.0, 1e-3, iters, 0)
    vel = np.sqrt(np.linalg.norm(uo)**2 + np.linalg.norm(vo)**2 + np.linalg.norm(wo)**2)
    out = []
    for name, a, b in zip("uvwp", (u, v, w, p), (uo, vo, wo, pp)):
        out.append(f"{name}: own {np.linalg.norm(a-b)/np.linalg.norm(b):.2e} vec {np.linalg.norm(a-b)/(vel if name!='p' else np.linalg.norm(b)):.2e} |{name}|={np.linalg.norm(b):.2e}")
    print(kw, iters, " | ".join(out), flush=True)

for kw in (dict(solver_type=3, iterations=20), dict(solver_type=1, iterations=30), dict(solver_type=2, iterations=10), dict(solver_type=2, iterations=50)):
    for iters in (1, 3, 10, 40):
        run(syn.hex_box(12, 8, 6), iters, syn.channel_bcs, **kw)
