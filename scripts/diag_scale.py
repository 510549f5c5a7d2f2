import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import orc_b200
from orc_b200 import synthetic as syn
from oracle import pyoracle as po
from cases import make_pair, settings_pair
for shape, iters in (((24, 24, 24), 4), ((32, 32, 32), 3), ((64, 64, 64), 0)):
    arrays = syn.hex_box(*shape)
    pm, om = make_pair(po, arrays)
    for m in (pm, om): syn.channel_bcs(m)
    ps, os_ = settings_pair(po)
    n = pm.n_cells
    reps = []
    u, v, w, p = (np.zeros(n) for _ in range(4))
    orc_b200.solve_steady(pm, u, v, w, p, ps, 1000.0, 1e-3, max(iters, 4), 1, on_report=reps.append)
    print(shape, "gpu u_avg", " ".join(f"{r['u_avg']:.3e}" for r in reps), "|p'|", " ".join(f"{r['pressure_correction']:.2e}" for r in reps), flush=True)
    if iters:
        uo, vo, wo, pp, orep, _ = om.solve_steady(np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n), os_, 1000.0, 1e-3, iters, 1)
        print(shape, "ora u_avg", " ".join(f"{r[1]:.3e}" for r in orep), "|p'|", " ".join(f"{r[8]:.2e}" for r in orep), flush=True)
