"""End-to-end parity of solve_steady (src/solver.rs:26-244) against the oracle on the same mesh and settings.
Bar from the north star: converged u/v/w/p <= 1e-8 relative L2."""
import numpy as np
import pytest

import orc_b200
from orc_b200 import synthetic as syn
from cases import make_pair, settings_pair
from conftest import rel_l2

pytestmark = pytest.mark.gpu
RHO, MU = 1000.0, 1e-3


def run_both(oracle, arrays, iters, bcs, **kw):
    pm, om = make_pair(oracle, arrays)
    for m in (pm, om):
        bcs(m)
    ps, os_ = settings_pair(oracle, **kw)
    n = pm.n_cells
    u, v, w, p = (np.zeros(n) for _ in range(4))
    reports = []
    orc_b200.solve_steady(pm, u, v, w, p, ps, RHO, MU, iters, 1, on_report=reports.append)
    uo, vo, wo, po_, orep, _ = om.solve_steady(np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n), os_, RHO, MU, iters, 1)
    return (u, v, w, p), (uo, vo, wo, po_), reports, orep


@pytest.mark.parametrize("kw,tol", [
    (dict(solver_type=3, iterations=20), 1e-8),                          # BiCGSTAB
    (dict(solver_type=1, iterations=30), 1e-10),                         # Jacobi
    (dict(solver_type=0, iterations=10), 1e-10),                         # Gauss-Seidel (intended semantics)
    (dict(solver_type=2, iterations=10), 1e-8),                          # Multigrid (reference default method)
    (dict(solver_type=2, iterations=10, momentum=3, limiter=3), 1e-8),   # + TVD QUICK
])
def test_hex_channel_fields_match_oracle(oracle, kw, tol):
    g, o, reports, orep = run_both(oracle, syn.hex_box(12, 8, 6), 3, syn.channel_bcs, **kw)
    for name, a, b in zip("uvwp", g, o):
        assert np.isfinite(a).all()
        assert rel_l2(a, b) <= tol, (name, rel_l2(a, b))
    assert len(reports) == 3 == len(orep)
    assert abs(reports[-1]["u_avg"] - orep[-1][1]) <= 1e-8 * abs(orep[-1][1])


def test_tet_channel_fields_match_oracle(oracle):
    g, o, _, _ = run_both(oracle, syn.tet_box(5, 4, 3), 2, lambda m: syn.channel_bcs(m, fully_3d=True), solver_type=2, iterations=8,
                          momentum=3, limiter=4)
    for name, a, b in zip("uvwp", g, o):
        assert rel_l2(a, b) <= 1e-8, (name, rel_l2(a, b))


def test_resident_solver_equals_one_shot(oracle):
    arrays = syn.hex_box(10, 6, 4)
    pm, _ = make_pair(oracle, arrays)
    syn.channel_bcs(pm)
    ps, _ = settings_pair(oracle, iterations=8)
    n = pm.n_cells
    u, v, w, p = (np.zeros(n) for _ in range(4))
    orc_b200.solve_steady(pm, u, v, w, p, ps, RHO, MU, 3, 0)
    st = orc_b200.SteadySolver(pm, ps, RHO, MU)
    st.set_fields(*(np.zeros(n) for _ in range(4)))
    st.iterate(1); st.iterate(2)
    for a, b in zip((u, v, w, p), st.get_fields()):
        assert np.array_equal(a, b)   # same kernels, same grid sizes: deterministic
    assert len(st.level_sizes()) == 4
    assert sum(st.phase_ms().values()) > 0
