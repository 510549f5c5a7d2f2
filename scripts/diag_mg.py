import os, sys
import numpy as np, scipy.sparse as sp
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import orc_b200
from orc_b200 import linear_algebra as la
from oracle import pyoracle as po
from cases import GOLDEN
name = sys.argv[1] if len(sys.argv) > 1 else "channel_flow"
k = np.load(os.path.join(GOLDEN, f"kat_{name}.npz"))
rp, co = k["rowptr"], k["col"]
n = rp.size - 1
ctx = orc_b200.default_context()
for key, bkey in (("a_u", "b_u"), ("pc_a", "pc_b")):
    a = sp.csr_matrix((k[key], co, rp), shape=(n, n))
    g = la.CsrMatrix.from_scipy(a, ctx); o = po.Csr.from_arrays(n, n, rp, co, k[key])
    b = k[bkey]
    for its in (5, 50):
        xg = np.zeros(n); la.iterative_solve(g, b, xg, its, 3, 0.5, 1e-3, 1)
        xo = po.iterative_solve(o, b, np.zeros(n), its, po.BICGSTAB, 0.5, 1e-3, 1)
        print(key, "bicgstab", its, np.linalg.norm(xg - xo) / np.linalg.norm(xo))
    try:
        x, glev = la.multigrid_trace(g, b, np.zeros(n), iteration_count=50)
    except Exception as e:
        print(key, "GPU MG failed:", e); x = None
        x5, glev = la.multigrid_trace(g, b, np.zeros(n), iteration_count=2)
    xo, olev = po.multigrid_trace(o, b, np.zeros(n), iterations=50 if x is not None else 2)
    for l, ((gr, ga), (orr, oa)) in enumerate(zip(glev, olev)):
        grp, gco, gva = gr.arrays(); orp, oco, ova = orr.arrays()
        same_r = np.array_equal(grp, orp) and np.array_equal(gco, oco) and np.array_equal(gva, ova)
        arp, aco, ava = ga.arrays(); brp, bco, bva = oa.arrays()
        same_p = np.array_equal(arp, brp) and np.array_equal(aco, bco)
        empty = int((np.diff(brp) == 0).sum())
        print(key, "level", l + 1, "dims", ga.dims, "R same", same_r, "A pattern same", same_p, "A values same", same_p and np.array_equal(ava, bva), "empty rows", empty,
              "max val diff", (np.abs(ava - bva).max() if same_p else None))
    if x is not None:
        print(key, "MG solution rel diff", np.linalg.norm(x - xo) / np.linalg.norm(xo))
