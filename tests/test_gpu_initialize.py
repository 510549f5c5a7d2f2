"""Flow initialisation, the step before the SIMPLE loop (src/solver.rs:246-352, 414-509, 703-770),
against the oracle's restatement: the Laplace system bit-exact, the fields bit-identical with reference-order reductions
and within the documented tolerance with the fused reductions."""
import numpy as np
import pytest

import orc_b200
from orc_b200 import synthetic as syn
from orc_b200 import discretization as disc
from orc_b200.settings import ReductionMode
from cases import make_pair, load_mesh_arrays, couette_bcs
from conftest import rel_l2

pytestmark = pytest.mark.gpu
RHO, MU = 1000.0, 1e-3


def csr_equal(g, o):
    grp, gco, gva = g.arrays()
    orp, oco, ova = o.arrays()
    assert np.array_equal(grp, orp) and np.array_equal(gco, oco)
    assert np.array_equal(gva, ova), np.abs(gva - ova).max()


@pytest.mark.parametrize("shape", [(8, 6, 4), (12, 8, 6)])
def test_pressure_laplace_system_bit_exact(oracle, shape):
    pm, om = make_pair(oracle, syn.hex_box(*shape))
    for m in (pm, om):
        syn.channel_bcs(m)
    ga, gb = disc.build_pressure_laplace(pm)
    oa, ob = om.build_pressure_laplace()
    csr_equal(ga, oa)
    assert np.array_equal(gb, ob)


def test_pressure_laplace_on_tets_bit_exact(oracle):
    pm, om = make_pair(oracle, syn.tet_box(4, 3, 3))
    for m in (pm, om):
        syn.channel_bcs(m)
    ga, gb = disc.build_pressure_laplace(pm)
    oa, ob = om.build_pressure_laplace()
    csr_equal(ga, oa)
    assert np.array_equal(gb, ob)


def test_check_boundary_conditions(oracle):
    pm, om = make_pair(oracle, syn.hex_box(6, 4, 3))
    with pytest.raises(orc_b200.OrcError) as e:          # all zones are stationary walls by default: nothing is set
        orc_b200.check_boundary_conditions(pm)
    assert "You must set boundary conditions" in e.value.message
    for m in (pm, om):
        syn.channel_bcs(m)
    assert orc_b200.check_boundary_conditions(pm) == om.check_boundary_conditions() == orc_b200.SystemConstraintType.PressureOnly
    for m in (pm, om):
        m.set_zone("WALL", 3, 0.0, (1e-3, 0.0, 0.0))      # a moving wall: velocity BC + two pressure BCs
    assert orc_b200.check_boundary_conditions(pm) == om.check_boundary_conditions() == orc_b200.SystemConstraintType.Hybrid


@pytest.mark.parametrize("iters", [8, 40])
def test_initialize_flow_bit_identical_in_reference_order(oracle, iters):
    pm, om = make_pair(oracle, syn.hex_box(10, 6, 4))
    for m in (pm, om):
        syn.channel_bcs(m)
    g = orc_b200.initialize_flow(pm, MU, RHO, iters, reduction_mode=ReductionMode.ReferenceOrder)
    o = om.initialize_flow(MU, RHO, iters)
    for c, a, b in zip("uvwp", g, o):
        assert np.array_equal(a, b), (c, rel_l2(a, b))


def test_initialize_flow_on_the_reference_couette_mesh(oracle):
    """tests.rs:84-86 initialises the flow on couette_flow_128x64x1.msh before solve_steady (with 1000 inner iterations; 30
    here: the unguarded BiCGSTAB of the reference amplifies rounding to O(1) long before that)."""
    arrays = load_mesh_arrays("couette_flow_128x64x1")
    pm, om = make_pair(oracle, arrays)
    for m in (pm, om):
        couette_bcs(m)
    g = orc_b200.initialize_flow(pm, MU, RHO, 30, reduction_mode=ReductionMode.ReferenceOrder)
    o = om.initialize_flow(MU, RHO, 30)
    for c, a, b in zip("uvwp", g, o):
        assert np.array_equal(a, b), (c, rel_l2(a, b))


def test_initialize_flow_fast_reductions_and_lockstep(oracle, monkeypatch):
    """Fused reductions: p is bit-exact up to the Jacobi norms (no effect on the values), u, v, w agree with the oracle to the
    solve tolerance; the lockstep u/v/w solve equals three sequential solves bit for bit."""
    pm, om = make_pair(oracle, syn.hex_box(10, 6, 4))
    for m in (pm, om):
        syn.channel_bcs(m)
    o = om.initialize_flow(MU, RHO, 6)
    g = orc_b200.initialize_flow(pm, MU, RHO, 6)
    vel = np.sqrt(sum(np.linalg.norm(b) ** 2 for b in o[:3]))
    assert np.array_equal(g[3], o[3])
    for a, b in zip(g[:3], o[:3]):
        assert np.isfinite(a).all() and np.linalg.norm(a - b) <= 1e-8 * vel
    monkeypatch.setenv("ORC_B200_BATCH", "0")
    s = orc_b200.initialize_flow(pm, MU, RHO, 6)
    for a, b in zip(g, s):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name,walls,moving,dp_dx,u_wall", [("channel_flow", ("WALL",), None, 5.0, 0.0),
                                                            ("couette_flow_128x64x1", ("TOP_WALL", "BOTTOM_WALL"), "TOP_WALL", 10.0, 5e-4)])
def test_initialize_flow_against_committed_golden(name, walls, moving, dp_dx, u_wall):
    """The reference's example meshes with the BCs of src/tests.rs:60-76 against tests/golden/kat_init_<name>.npz: the Laplace
    system and, with reference-order reductions, the initial fields — bit for bit, no oracle involved at run time."""
    import os
    from cases import GOLDEN
    k = np.load(os.path.join(GOLDEN, f"kat_init_{name}.npz"))
    m = orc_b200.Mesh.from_arrays(*syn.mesh_args(load_mesh_arrays(name)))
    couette_bcs(m, u_wall=u_wall, dp_dx=dp_dx, wall_zones=walls, moving=moving)
    a, b = disc.build_pressure_laplace(m)
    assert np.array_equal(a.arrays()[2], k["laplace_val"]) and np.array_equal(b, k["laplace_b"])
    assert orc_b200.check_boundary_conditions(m) == int(k["constraint_type"])
    g = orc_b200.initialize_flow(m, MU, RHO, int(k["iters"]), reduction_mode=ReductionMode.ReferenceOrder)
    for c, x in zip("uvwp", g):
        assert np.array_equal(x, k[c]), (c, rel_l2(x, k[c]))
