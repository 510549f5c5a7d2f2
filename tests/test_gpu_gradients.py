"""Least-squares gradients (src/solver.rs:803-869, 903-947), the velocity-inlet flow initialisation (initialize_flow_new /
initialize_velocity_field, :354-410, 511-696) and write_gradients (src/io.rs:623-662) against the oracle, through the C ABI.
Everything here is per-cell arithmetic in the reference's operation order (nalgebra's dense kernels restated): bit-exact."""
import os

import numpy as np
import pytest

import orc_b200
from orc_b200 import discretization as disc
from orc_b200 import synthetic as syn
from orc_b200.settings import GradientReconstructionMethods as G
from cases import make_pair, settings_pair, smooth_fields, load_mesh_arrays, couette_bcs
from conftest import rel_l2

pytestmark = pytest.mark.gpu
RHO, MU = 1000.0, 1e-3
MESHES = {"hex_10x7x5": lambda: syn.hex_box(10, 7, 5), "tet_5x4x3": lambda: syn.tet_box(5, 4, 3), "hex_16x16x1": lambda: syn.hex_box(16, 16, 1),
          "hex_16x16x1_exact": lambda: syn.hex_box(16, 16, 1, jitter=0.0)}   # exactly axis aligned: no z differences between neighbours


def setup(oracle, name, velocity_inlet=False):
    pm, om = make_pair(oracle, MESHES[name]())
    for m in (pm, om):
        syn.channel_bcs(m, fully_3d=True)
        if velocity_inlet:
            m.set_zone("INLET", 10, 0.0, (1e-3, 1e-4, 0.0))
    return pm, om


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("scheme", [0, 2], ids=["green_gauss", "least_squares"])
def test_gradients_of_every_cell_bit_exact(oracle, ctx, name, scheme):
    pm, om = setup(oracle, name, velocity_inlet=True)
    u, v, w, p = smooth_fields(pm.export())
    gp, gu = disc.calculate_gradients(pm, u, v, w, p, scheme, ctx)
    op, ou = om.gradients(u, v, w, p, gradient=scheme)
    assert np.array_equal(gp, op), np.abs(gp - op).max()
    assert np.array_equal(gu, ou), np.abs(gu - ou).max()


def test_least_squares_gradient_is_exact_for_linear_fields_inside(ctx):
    """Known answer: away from the boundary (where the reference feeds the boundary VALUE instead of a difference) the
    least-squares fit reproduces the gradient of a linear field."""
    pm = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(6, 5, 4)))
    syn.channel_bcs(pm, fully_3d=True)
    ex = pm.export()
    cc = ex["cell_centroid"]
    p = 3 * cc[:, 0] + 2 * cc[:, 1] - cc[:, 2] + 1
    u, v, w = 2 * cc[:, 0], -cc[:, 1], 0.5 * cc[:, 2] + cc[:, 0]
    gp, gu = disc.calculate_gradients(pm, u, v, w, p, G.LeastSquares, ctx)
    co, cf, c1 = ex["cell_face_offsets"], ex["cell_face_indices"], ex["face_c1"]
    inside = [i for i in range(pm.n_cells) if all(c1[f] >= 0 for f in cf[co[i]:co[i + 1]])]
    assert len(inside) == 4 * 3 * 2
    assert np.abs(gp[inside] - [3, 2, -1]).max() < 1e-9
    assert np.abs(gu[inside] - [[2, 0, 0], [0, -1, 0], [1, 0, 0.5]]).max() < 1e-9


@pytest.mark.parametrize("name", ["hex_10x7x5", "tet_5x4x3"])
@pytest.mark.parametrize("kw", [dict(gradient=2), dict(gradient=2, momentum=3, limiter=4), dict(gradient=2, velocity_interpolation=1, pressure_interpolation=1)],
                         ids=["defaults", "tvd_umist", "linear_weighted"])
def test_assembly_with_least_squares_gradients(oracle, ctx, name, kw):
    """GradientReconstructionMethods::LeastSquares through the whole assembly (Rhie-Chow and SecondOrder read grad p, TVD reads
    grad u): two momentum assemblies + the pressure-correction system, bit-exact."""
    pm, om = setup(oracle, name)
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
            assert np.array_equal(g_a[k].arrays()[2], o_a[k].arrays()[2]), (sweep, k)
            assert np.array_equal(gb[k], ob[k]), (sweep, "b", k)
    gpa, gpb = disc.build_pressure_correction_matrices(pm, u, v, w, p, *g_a, ps, RHO)
    opa, opb = om.build_pressure_correction(*o_a, u, v, w, p, os_, RHO)
    assert np.array_equal(gpa.arrays()[2], opa.arrays()[2]) and np.array_equal(gpb, opb)


def test_solve_steady_with_least_squares_gradients(oracle):
    """Three SIMPLE iterations with LeastSquares gradient reconstruction, reference-order reductions: bit-identical fields."""
    pm, om = setup(oracle, "hex_10x7x5")
    ps, os_ = settings_pair(oracle, reference_order=True, gradient=2)
    n = pm.n_cells
    u, v, w, p = (np.zeros(n) for _ in range(4))
    orc_b200.solve_steady(pm, u, v, w, p, ps, RHO, MU, 3, 0)
    z = np.zeros(n)
    uo, vo, wo, po_, _, _ = om.solve_steady(z, z, z, z, os_, RHO, MU, 3, 0)
    for c, a, b in zip("uvwp", (u, v, w, p), (uo, vo, wo, po_)):
        assert np.array_equal(a, b), (c, rel_l2(a, b))


@pytest.mark.parametrize("name", list(MESHES))
def test_velocity_potential_system_and_gradient(oracle, ctx, name):
    """The psi system of initialize_velocity_field and the least-squares velocity (column selection on the one-cell-thick mesh)."""
    pm, om = setup(oracle, name, velocity_inlet=True)
    ga, gb = disc.build_velocity_potential(pm, ctx)
    oa, ob = om.build_velocity_potential()
    for x, y in zip(ga.arrays(), oa.arrays()):
        assert np.array_equal(x, y)
    assert np.array_equal(gb, ob)
    psi = np.cos(np.arange(pm.n_cells) * 0.37) * 1e-4
    gu, gv, gw = disc.potential_gradient(pm, psi, ctx)
    og = om.potential_gradient(psi)
    assert np.array_equal(gu, og[:, 0]) and np.array_equal(gv, og[:, 1]) and np.array_equal(gw, og[:, 2])
    if name == "hex_16x16x1_exact":
        assert not gw.any()   # no z differences between neighbours: the z column is dropped, w stays zero


@pytest.mark.parametrize("name", ["hex_10x7x5", "tet_5x4x3"])
@pytest.mark.parametrize("kind", ["velocity_only", "pressure_only", "hybrid"])
def test_initialize_flow_new_matches_oracle(oracle, name, kind):
    """initialize_flow_new (src/solver.rs:354-410) on the three constraint systems: VelocityOnly (one pressure outlet) runs
    initialize_velocity_field, PressureOnly and Hybrid (overlapping match arm) only initialise the pressure. Default (Auto)
    reductions = reference order at this size: bit-identical."""
    pm, om = make_pair(oracle, MESHES[name]())
    for m in (pm, om):
        syn.channel_bcs(m, fully_3d=True)
        if kind in ("velocity_only", "hybrid"):
            m.set_zone("INLET", 10, 0.0, (1e-3, 0.0, 0.0))
        if kind == "hybrid":
            m.set_zone("WALL", 4, 0.5, (0.0, 0.0, 0.0))   # a second pressure boundary: > 1 pressure BCs + a velocity BC
    expect = {"velocity_only": 1, "pressure_only": 0, "hybrid": 2}[kind]
    assert orc_b200.check_boundary_conditions(pm) == om.check_boundary_conditions() == expect
    g = orc_b200.initialize_flow_new(pm, MU, RHO, 10)
    o = om.initialize_flow_new(MU, RHO, 10)
    for c, a, b in zip("uvwp", g, o):
        assert np.array_equal(a, b), (c, rel_l2(a, b))
    if kind == "velocity_only":
        assert g[0].any() and not g[3].any()
    else:
        assert g[3].any() and not (g[0].any() or g[1].any() or g[2].any())


def test_velocity_inlet_driver_of_the_reference_from_a_cold_start(oracle):
    """src/tests.rs:153-209 (what the reference's current main() runs): VelocityInlet / PressureOutlet channel on
    couette_flow_128x64x1.msh, initialize_flow_new, then solve_steady. With the reference's Multigrid solver the first iteration
    ends in "Multigrid diverged" in both implementations (zero right-hand sides); with the Jacobi solver the cold
    start + three SIMPLE iterations are bit-identical to the oracle."""
    pm, om = make_pair(oracle, load_mesh_arrays("couette_flow_128x64x1"))
    for m in (pm, om):
        couette_bcs(m, u_wall=0.0, dp_dx=0.0)
        m.set_zone("INLET", 10, 0.0, (1e-3, 0.0, 0.0))
    g = orc_b200.initialize_flow_new(pm, MU, RHO, 1000)
    o = om.initialize_flow_new(MU, RHO, 1000)
    for c, a, b in zip("uvwp", g, o):
        assert np.array_equal(a, b), (c, rel_l2(a, b))
    ps, os_ = settings_pair(oracle, reference_order=True, solver_type=1, iterations=30)
    u, v, w, p = (a.copy() for a in g)
    orc_b200.solve_steady(pm, u, v, w, p, ps, RHO, MU, 3, 0)
    uo, vo, wo, po_, _, _ = om.solve_steady(*o, os_, RHO, MU, 3, 0)
    for c, a, b in zip("uvwp", (u, v, w, p), (uo, vo, wo, po_)):
        assert np.array_equal(a, b), (c, rel_l2(a, b))


def test_write_gradients_format(tmp_path, oracle, ctx):
    """write_gradients (src/io.rs:623-662): `centroid \\t (nine values, ) \\t (three values, )` with the separators the reference
    leaves in place, Rust's `{:.Ne}` formatting; the numbers are the device gradients (== the oracle's)."""
    pm, om = setup(oracle, "hex_10x7x5")
    u, v, w, p = smooth_fields(pm.export())
    path = os.path.join(tmp_path, "gradients.csv")
    orc_b200.write_gradients(pm, u, v, w, p, path, 3, G.GreenGaussCellBased, ctx)
    op, ou = om.gradients(u, v, w, p, gradient=0)
    lines = open(path).read().splitlines()
    assert len(lines) == pm.n_cells
    from orc_b200.io import format_gradient_line
    cc = pm.export()["cell_centroid"]
    for i in (0, 17, pm.n_cells - 1):
        assert lines[i] == format_gradient_line(cc[i], ou[i].ravel(), op[i], 3)
    centroid, vg, pg = lines[0].split("\t")
    assert vg.startswith("(") and vg.endswith(", )") and vg.count(", ") == 9 and pg.count(", ") == 3
    assert all("e" in tok for tok in vg[1:-3].split(", "))


@pytest.mark.parametrize("kind", ["hex", "tet", "2d", "tgrid_file"])
def test_device_geometry_is_bit_identical_to_the_host_pass(ctx, tmp_path, kind):
    """The geometry pass of read_mesh (src/io.rs:289-438) on the device, from the node coordinates: face normal / centroid / area
    (triangle fan), cell centroid / volume — same operator order as the host pass, so every value is bit-identical. The TGRID file
    case writes faces whose first cell is missing (flipped normals, io.rs:332-337)."""
    if kind == "hex":
        m = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(40, 30, 20)))
    elif kind == "tet":
        m = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.tet_box(12, 9, 7)))
    elif kind == "2d":
        m = orc_b200.Mesh.from_arrays(*syn.mesh_args(syn.hex_box(16, 16, 1)))
    else:
        a = syn.hex_box(6, 5, 4)
        swap = a["c1"] == 0                      # boundary faces: write them as (0, cell) so that the reader flips them
        a["c0"], a["c1"] = np.where(swap, 0, a["c0"]), np.where(swap, a["c0"], a["c1"])
        path = os.path.join(tmp_path, "box.msh")
        syn.write_tgrid(path, a)
        m = orc_b200.read_mesh(path)
    host = m.export()
    dev = m.geometry_on_device(ctx)
    for k in ("face_area", "face_normal", "face_centroid", "cell_volume", "cell_centroid"):
        assert np.array_equal(dev[k], host[k]), k
    assert dev["device_ms"] > 0
