"""The other cell types the reference accepts (README "Supported cell types"): triangular prisms and pyramids (five faces, triangles
and quads in one mesh) and polyhedra with six-node polygon faces, through the whole loop against the oracle. Host geometry, pattern, level
schedule and the TGRID round trip of these meshes are checked on the CPU (tests/test_host_logic.py); here: three SIMPLE iterations
through `orc_solve_steady` with reference-order reductions, fields within the north star's 1e-8 of the oracle's.
(Named zzz so that it runs last: added after the round's GPU budget was spent; the device code is generic over the cell -> face
lists and had run on 4- and 6-face cells only.)"""
import numpy as np
import pytest

import orc_b200
from orc_b200 import synthetic as syn
from cases import make_pair, settings_pair

pytestmark = pytest.mark.gpu
RHO, MU = 1000.0, 1e-3
CASES = [("wedge", lambda: syn.wedge_box(6, 4, 3), dict(solver_type=2)),
         ("wedge-umist", lambda: syn.wedge_box(6, 4, 3), dict(solver_type=2, momentum=3, limiter=4)),
         ("polyhedra", lambda: syn.poly_box(8, 4, 3), dict(solver_type=2)),
         ("pyramid", lambda: syn.pyramid_box(4, 3, 2), dict(solver_type=2))]


@pytest.mark.parametrize("name,gen,kw", CASES, ids=[c[0] for c in CASES])
def test_wedge_and_polyhedral_meshes_match_the_oracle(oracle, name, gen, kw):
    pm, om = make_pair(oracle, gen())
    for m in (pm, om):
        syn.channel_bcs(m, fully_3d=True)
    ps, os_ = settings_pair(oracle, reference_order=True, **kw)
    n = pm.n_cells
    u, v, w, p = (np.zeros(n) for _ in range(4))
    orc_b200.solve_steady(pm, u, v, w, p, ps, RHO, MU, 3, 0)
    z = np.zeros(n)
    uo, vo, wo, po_ = om.solve_steady(z, z, z, z, os_, RHO, MU, 3, 0)[:4]
    vel = np.sqrt(sum(np.linalg.norm(b) ** 2 for b in (uo, vo, wo)))
    errs = {c: float(np.linalg.norm(a - b) / (np.linalg.norm(b) if c == "p" else vel)) for c, a, b in zip("uvwp", (u, v, w, p), (uo, vo, wo, po_))}
    print(name, errs, "bit-identical:", all(np.array_equal(a, b) for a, b in zip((u, v, w, p), (uo, vo, wo, po_))))
    for c, e in errs.items():
        assert np.isfinite(e) and e <= 1e-8, (c, e)
