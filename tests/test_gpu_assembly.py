"""GPU parity of the assembly kernels (discretization.rs + solver.rs helpers) against the oracle, per call with identical
inputs, through the C ABI. Bar: coefficients <= 1e-12 relative — in practice bit-exact, because the kernels
keep the reference's operation order and are compiled without FMA contraction."""
import numpy as np
import pytest

import orc_b200
from orc_b200 import discretization as disc
from orc_b200 import linear_algebra as la
from orc_b200 import synthetic as syn
from cases import make_pair, settings_pair, smooth_fields
from conftest import max_rel

pytestmark = pytest.mark.gpu

MESHES = {
    "hex_10x7x5": lambda: syn.hex_box(10, 7, 5),
    "hex_16x16x1": lambda: syn.hex_box(16, 16, 1),
    "tet_5x4x3": lambda: syn.tet_box(5, 4, 3),
}
RHO, MU = 1000.0, 1e-3


def csr_vals(g, o):
    grp, gco, gva = g.arrays()
    orp, oco, ova = o.arrays()
    assert np.array_equal(grp, orp) and np.array_equal(gco, oco), "sparsity pattern differs"
    return gva, ova


def setup(oracle, name, fully_3d=False, velocity_inlet=False):
    arrays = MESHES[name]()
    pm, om = make_pair(oracle, arrays)
    for m in (pm, om):
        syn.channel_bcs(m, fully_3d=fully_3d)
        if velocity_inlet:
            m.set_zone("INLET", 10, 0.0, (1e-3, 1e-4, 0.0))
            m.set_zone("WALL", 3, 0.0, (2e-4, 0.0, 1e-4))
    return pm, om


@pytest.mark.parametrize("name", list(MESHES))
def test_diffusion_and_init_matrices(oracle, ctx, name):
    pm, om = setup(oracle, name, velocity_inlet=True)
    ga, gbu, gbv, gbw = disc.build_momentum_diffusion_matrix(pm, MU, ctx)
    oa, obu, obv, obw = om.build_momentum_diffusion(MU)
    gv, ov = csr_vals(ga, oa)
    assert np.array_equal(gv, ov)
    assert np.array_equal(gbu, obu) and np.array_equal(gbv, obv) and np.array_equal(gbw, obw)
    gv, ov = csr_vals(disc.initialize_momentum_matrix(pm, ctx), om.init_momentum_matrix())
    assert np.array_equal(gv, ov)


@pytest.mark.parametrize("name", list(MESHES))
def test_pressure_gradient_reproduces_float_times_vector_quirk(oracle, ctx, name):
    pm, om = setup(oracle, name)
    _, _, _, p = smooth_fields(pm.export())
    g = disc.calculate_pressure_gradient(pm, p, ctx)
    o = om.pressure_gradient(p)
    assert np.array_equal(g, o)
    assert np.array_equal(g[:, 2], g[:, 1])  # Q1: .z receives the .y sum


SCHEMES = [
    dict(),                                                                  # defaults: CD1, SecondOrder, RhieChow
    dict(momentum=0),                                                        # UD
    dict(momentum=3, limiter=2), dict(momentum=3, limiter=3), dict(momentum=3, limiter=4),  # TVD LUD / QUICK / UMIST
    dict(velocity_interpolation=0, pressure_interpolation=0),                # Linear / Linear
    dict(velocity_interpolation=1, pressure_interpolation=1),                # LinearWeighted / LinearWeighted
]


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("kw", SCHEMES, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()) or "default")
def test_momentum_and_pressure_assembly(oracle, ctx, name, kw):
    """Two consecutive assemblies (the second starts from the diagonals the first one left: the Q2 recurrence state),
    each followed by the pressure-correction assembly and the correction step."""
    pm, om = setup(oracle, name, velocity_inlet=(kw.get("momentum") == 0))
    ps, os_ = settings_pair(oracle, **kw)
    u, v, w, p = smooth_fields(pm.export())
    g_di, *_ = disc.build_momentum_diffusion_matrix(pm, MU, ctx)
    o_di, *_ = om.build_momentum_diffusion(MU)
    g_a = [disc.initialize_momentum_matrix(pm, ctx) for _ in range(3)]
    o_a = [om.init_momentum_matrix() for _ in range(3)]
    for sweep in range(2):
        gb = disc.build_momentum_advection_matrices(*g_a, g_di, pm, u, v, w, p, ps, RHO)
        ob = om.build_momentum_advection(*o_a, o_di, u, v, w, p, os_, RHO)
        for k in range(3):
            gv, ov = csr_vals(g_a[k], o_a[k])
            assert max_rel(gv, ov) <= 1e-12, (sweep, k, max_rel(gv, ov))
            assert np.array_equal(gv, ov), (sweep, k)
            assert np.array_equal(gb[k], ob[k]), (sweep, "b", k)
        assert np.allclose(gb[3][0], ob[3][0], rtol=1e-12, atol=0) and gb[3][1] == ob[3][1] and gb[3][2] == ob[3][2]
        gpa, gpb = disc.build_pressure_correction_matrices(pm, u, v, w, p, *g_a, ps, RHO)
        opa, opb = om.build_pressure_correction(*o_a, u, v, w, p, os_, RHO)
        gv, ov = csr_vals(gpa, opa)
        assert np.array_equal(gv, ov) and np.array_equal(gpb, opb)
        pp = 1e-3 * np.cos(np.arange(u.size))
        gu, gvv, gw, gp, gn = disc.apply_pressure_correction(pm, *g_a, pp, u, v, w, p, ps)
        ou, ovv, ow, op_, on = om.apply_pressure_correction(*o_a, pp, u, v, w, p, os_)
        assert np.array_equal(gu, ou) and np.array_equal(gvv, ovv) and np.array_equal(gw, ow) and np.array_equal(gp, op_)
        assert np.allclose(gn, on, rtol=1e-13, atol=0)
        u, v, w, p = ou, ovv, ow, op_


def test_frozen_mode_differs_only_through_the_diagonals(oracle, ctx):
    """Frozen assembly is the documented deviation: first sweep from diag == 1 everywhere equals the exact mode only where
    no lower neighbour exists; b (pressure source) is identical in both modes."""
    from orc_b200.settings import AssemblyMode
    pm, om = setup(oracle, "hex_10x7x5")
    ps, os_ = settings_pair(oracle)
    ps.assembly_mode = AssemblyMode.Frozen
    u, v, w, p = smooth_fields(pm.export())
    g_di, *_ = disc.build_momentum_diffusion_matrix(pm, MU, ctx)
    o_di, *_ = om.build_momentum_diffusion(MU)
    g_a = [disc.initialize_momentum_matrix(pm, ctx) for _ in range(3)]
    o_a = [om.init_momentum_matrix() for _ in range(3)]
    gb = disc.build_momentum_advection_matrices(*g_a, g_di, pm, u, v, w, p, ps, RHO)
    ob = om.build_momentum_advection(*o_a, o_di, u, v, w, p, os_, RHO)
    assert np.array_equal(gb[0], ob[0])
    gv, ov = csr_vals(g_a[0], o_a[0])
    rp, _ = pm.pattern()
    assert np.array_equal(gv[rp[0]:rp[1]], ov[rp[0]:rp[1]])   # cell 0 has no lower neighbour
    assert not np.array_equal(gv, ov)


def test_unsupported_schemes_map_to_status_codes(oracle, ctx):
    pm, _ = setup(oracle, "hex_16x16x1")
    n = pm.n_cells
    z = np.zeros(n)
    for kw in (dict(momentum=2), dict(pressure_interpolation=2), dict(velocity_interpolation=3), dict(gradient=1)):
        ps, _ = settings_pair(oracle, **kw)
        with pytest.raises(orc_b200.OrcError) as e:
            orc_b200.solve_steady(pm, z.copy(), z.copy(), z.copy(), z.copy(), ps, RHO, MU, 1, 1)
        assert e.value.code == orc_b200._lib.E_UNSUPPORTED
    pm.set_zone("SYM", 36, 0.0, (0, 0, 0))  # Outflow: a BC the path panics on
    ps, _ = settings_pair(oracle)
    with pytest.raises(orc_b200.OrcError) as e:
        orc_b200.solve_steady(pm, z.copy(), z.copy(), z.copy(), z.copy(), ps, RHO, MU, 1, 1)
    assert e.value.code == orc_b200._lib.E_UNSUPPORTED
