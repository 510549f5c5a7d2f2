"""GPU parity on the reference's own example meshes (BASELINE.json configs 1-2): against the committed golden outputs
(tests/golden/kat_*.npz, produced by the oracle in the build container) and against the oracle run live — including a
CONVERGED solution, which is what the north star's 1e-8 field bar is stated for."""
import os

import numpy as np
import pytest

import orc_b200
from orc_b200 import discretization as disc
from orc_b200 import synthetic as syn
from cases import GOLDEN, couette_bcs, figure_analytical, figure_misfit, load_figure, load_mesh_arrays, make_pair, settings_pair
from conftest import max_rel, rel_l2

pytestmark = pytest.mark.gpu
RHO, MU = 1000.0, 1e-3
CASES = [("channel_flow", ("WALL",), None, 5.0, 0.0), ("couette_flow_128x64x1", ("TOP_WALL", "BOTTOM_WALL"), "TOP_WALL", 10.0, 5e-4)]


def product_mesh(name, walls, moving, dp_dx, u_wall):
    m = orc_b200.Mesh.from_arrays(*syn.mesh_args(load_mesh_arrays(name)))
    couette_bcs(m, u_wall=u_wall, dp_dx=dp_dx, wall_zones=walls, moving=moving)
    return m


def settings_from(k):
    from orc_b200 import settings as S
    mom, lim, pint, vint = (int(x) for x in k["settings"])
    return orc_b200.NumericalSettings(momentum=S.MomentumDiscretization(mom), limiter=S.TvdLimiter(lim),
                                      pressure_interpolation=S.PressureInterpolation(pint), velocity_interpolation=S.VelocityInterpolation(vint),
                                      reduction_mode=S.ReductionMode.ReferenceOrder)


@pytest.mark.parametrize("name,walls,moving,dp_dx,u_wall", CASES)
def test_assembly_from_golden_fields_is_bit_exact(ctx, name, walls, moving, dp_dx, u_wall):
    """Inputs (u, v, w, p after a few iterations) and outputs (all coefficient arrays) come from the golden file."""
    k = np.load(os.path.join(GOLDEN, f"kat_{name}.npz"))
    m = product_mesh(name, walls, moving, dp_dx, u_wall)
    s = settings_from(k)
    rp, co = m.pattern()
    assert np.array_equal(rp, k["rowptr"]) and np.array_equal(co, k["col"])          # sparsity pattern: bit-exact
    a_di, *_ = disc.build_momentum_diffusion_matrix(m, MU, ctx)
    assert np.array_equal(a_di.arrays()[2], k["a_di"])
    mats = [disc.initialize_momentum_matrix(m, ctx) for _ in range(3)]
    bu, bv, bw, pe = disc.build_momentum_advection_matrices(*mats, a_di, m, k["u"], k["v"], k["w"], k["p"], s, RHO)
    for g, key in zip(mats, ("a_u", "a_v", "a_w")):
        assert max_rel(g.arrays()[2], k[key]) <= 1e-12
        assert np.array_equal(g.arrays()[2], k[key])
    assert np.array_equal(bu, k["b_u"]) and np.array_equal(bv, k["b_v"]) and np.array_equal(bw, k["b_w"])
    assert np.allclose(pe, k["peclet"], rtol=1e-12, atol=0)
    pa, pb = disc.build_pressure_correction_matrices(m, k["u"], k["v"], k["w"], k["p"], *mats, s, RHO)
    assert np.array_equal(pa.arrays()[2], k["pc_a"]) and np.array_equal(pb, k["pc_b"])


@pytest.mark.parametrize("name,walls,moving,dp_dx,u_wall", CASES)
def test_fields_after_a_few_iterations_are_bit_identical_to_golden(name, walls, moving, dp_dx, u_wall):
    """On these meshes the momentum systems converge to machine precision well before the 50th BiCGSTAB iteration and the
    reference's unguarded recurrences (Q8) then run on rounding noise: the result depends on the last bit of every dot
    product (measured: 4e-3 relative between two summation orders on channel_flow.msh). Parity on the
    reference's own cases is therefore demonstrated in ORC_REDUCE_REFERENCE_ORDER mode, where every reduction follows
    nalgebra's accumulation order and the fields are BIT-IDENTICAL to the reference's CPU path."""
    k = np.load(os.path.join(GOLDEN, f"kat_{name}.npz"))
    m = product_mesh(name, walls, moving, dp_dx, u_wall)
    n = m.n_cells
    u, v, w, p = (np.zeros(n) for _ in range(4))
    reps = []
    orc_b200.solve_steady(m, u, v, w, p, settings_from(k), RHO, MU, int(k["iters"]), 1, on_report=reps.append)
    for c, a in zip("uvwp", (u, v, w, p)):
        assert np.array_equal(a, k[c]), (c, rel_l2(a, k[c]))
    assert np.allclose([r["u_avg"] for r in reps], k["reports"][:, 1], rtol=1e-12)


def test_converged_poiseuille_fields_match_oracle_and_analytical(oracle):
    """120 SIMPLE iterations on channel_flow.msh (config 2: TVD, Rhie-Chow, SecondOrder, AMG/BiCGSTAB): the mean has settled to
    5 digits. Reference-order mode: converged fields bit-identical to the oracle (bar: <= 1e-8 relative L2), and within the
    reference's own 10 % validation threshold of the analytical profile (src/tests.rs:111-151)."""
    arrays = load_mesh_arrays("channel_flow")
    pm, om = make_pair(oracle, arrays)
    for m in (pm, om):
        couette_bcs(m, u_wall=0.0, dp_dx=5.0, wall_zones=("WALL",), moving=None)
    ps, os_ = settings_pair(oracle, reference_order=True, momentum=3, limiter=4)
    n = pm.n_cells
    u, v, w, p = (np.zeros(n) for _ in range(4))
    orc_b200.solve_steady(pm, u, v, w, p, ps, RHO, MU, 120, 0)
    z = np.zeros(n)
    uo, vo, wo, po_, _, _ = om.solve_steady(z, z, z, z, os_, RHO, MU, 120, 0)
    for c, a, b in zip("uvwp", (u, v, w, p), (uo, vo, wo, po_)):
        assert rel_l2(a, b) <= 1e-8, (c, rel_l2(a, b))
        assert np.array_equal(a, b), c
    mean_exact = -(1e-3 ** 2) / (12 * 1e-3) * 5.0
    assert abs(u.mean() - mean_exact) < 0.1 * abs(mean_exact) and abs(u.min() + 6.25e-4) < 0.1 * 6.25e-4


def test_default_settings_reproduce_the_reference_on_its_own_mesh(oracle):
    """The DEFAULT reduction mode (ORC_REDUCE_AUTO) on the same case: a 1008-cell mesh is far below ORC_AUTO_EXACT_MAX_ROWS, so
    the solve runs in the reference's summation order and must do exactly what the oracle does — 120 iterations without a
    divergence status, fields bit-identical, the analytical Poiseuille mean and extremum within the reference's 10 %
    (src/tests.rs:111-151). (With ORC_REDUCE_FAST forced this mesh ends in "Multigrid diverged": it is exactly axis aligned, the
    v / w right-hand sides are rounding noise and the unguarded BiCGSTAB divides noise by noise — DESIGN.md §5.)"""
    pm, om = make_pair(oracle, load_mesh_arrays("channel_flow"))
    for m in (pm, om):
        couette_bcs(m, u_wall=0.0, dp_dx=5.0, wall_zones=("WALL",), moving=None)
    from orc_b200 import settings as S
    ps = orc_b200.NumericalSettings(momentum=S.MomentumDiscretization.TVD, limiter=S.TVD_UMIST)
    assert ps.reduction_mode == S.ReductionMode.Auto
    n = pm.n_cells
    u, v, w, p = (np.zeros(n) for _ in range(4))
    orc_b200.solve_steady(pm, u, v, w, p, ps, RHO, MU, 120, 0)   # raises OrcError on any divergence status
    z = np.zeros(n)
    uo, vo, wo, po_, _, _ = om.solve_steady(z, z, z, z, oracle.Settings(momentum=oracle.TVD, limiter=oracle.PSI_UMIST), RHO, MU, 120, 0)
    for c, a, b in zip("uvwp", (u, v, w, p), (uo, vo, wo, po_)):
        assert np.array_equal(a, b), (c, rel_l2(a, b))
    mean_exact = -(1e-3 ** 2) / (12 * 1e-3) * 5.0
    assert abs(u.mean() - mean_exact) < 0.1 * abs(mean_exact) and abs(u.min() + 6.25e-4) < 0.1 * 6.25e-4


def test_couette_validation_case_of_the_reference_converged(oracle):
    """The reference's "couette_flow" validation case (src/main.rs:85-102 -> src/tests.rs:44-151) on its own mesh: 60 SIMPLE
    iterations from rest, TVD-UMIST / SecondOrder / Rhie-Chow / Multigrid. With reference-order reductions the GPU fields are
    bit-identical to the oracle's after all 60 iterations, and bulk / minimum / maximum velocity meet the reference's 10 %
    criterion against the analytical Couette-Poiseuille profile; the throughput mode (fused reductions) meets it too."""
    from orc_b200 import settings as S
    from test_oracle_kats import couette_analytical, reference_compare
    arrays = load_mesh_arrays("couette_flow_128x64x1")
    pm, om = make_pair(oracle, arrays)
    for m in (pm, om):
        couette_bcs(m, u_wall=5e-4, dp_dx=10.0)
    n = pm.n_cells
    z = np.zeros(n)
    uo, vo, wo, po_, _, _ = om.solve_steady(z, z, z, z, oracle.Settings(momentum=oracle.TVD, limiter=oracle.PSI_UMIST), RHO, MU, 60, 0)
    avg, lo, hi = couette_analytical(5e-4, 10.0, MU)
    for mode in (S.ReductionMode.ReferenceOrder, S.ReductionMode.Fast):
        ps = orc_b200.NumericalSettings(momentum=S.MomentumDiscretization.TVD, limiter=S.TVD_UMIST, reduction_mode=mode)
        u, v, w, p = (np.zeros(n) for _ in range(4))
        orc_b200.solve_steady(pm, u, v, w, p, ps, RHO, MU, 60, 0)
        for got, exact in ((u.mean(), avg), (u.min(), lo), (u.max(), hi)):
            assert reference_compare(got, exact, 0.1) and abs(got - exact) < 0.1 * abs(exact), (mode, got, exact)
        if mode == S.ReductionMode.ReferenceOrder:
            for c, a, b in zip("uvwp", (u, v, w, p), (uo, vo, wo, po_)):
                assert np.array_equal(a, b), (c, rel_l2(a, b))
        else:
            vel = np.sqrt(sum(np.linalg.norm(b) ** 2 for b in (uo, vo, wo)))
            print("fused reductions vs oracle after 60 iterations:", [float(np.linalg.norm(a - b) / vel) for a, b in zip((u, v, w), (uo, vo, wo))])


def test_product_reproduces_the_couette_profile_real_orc_published():
    """The CUDA path against the one output of REAL ORC available here — the scatter plot examples/couette_flow_velocity_profile.png,
    digitised to a fraction of a pixel (tests/golden/digitise_reference_figures.py; 1 px = 8.7e-7 m/s = 0.1 % of the velocity range).
    Same case as the figure (src/main.rs:85-102 with dp/dx = 5: TVD-UMIST, SecondOrder, Rhie-Chow, Multigrid / BiCGSTAB), default
    reduction mode, 200 SIMPLE iterations from rest through orc_solve_steady: the per-level mid-range of u must sit on the figure's
    marker blobs within 0.5 px rms / 1 px max and reproduce their inlet-to-outlet spread — bounds the analytical profile misses by
    an order of magnitude (4.9 px rms), i.e. the test sees ORC's own discretisation error, not just "a parabola"."""
    from orc_b200 import settings as S
    fig = load_figure("couette_flow_velocity_profile")
    m = orc_b200.Mesh.from_arrays(*syn.mesh_args(load_mesh_arrays("couette_flow_128x64x1")))
    couette_bcs(m, u_wall=float(fig["u_wall"]), dp_dx=float(fig["dp_dx"]))
    n = m.n_cells
    u, v, w, p = (np.zeros(n) for _ in range(4))
    ps = orc_b200.NumericalSettings(momentum=S.MomentumDiscretization.TVD, limiter=S.TVD_UMIST)
    orc_b200.solve_steady(m, u, v, w, p, ps, float(fig["rho"]), float(fig["mu"]), 200, 0)
    cell_y = m.export()["cell_centroid"][:, 1]
    rms, worst, spread = figure_misfit(fig, u, cell_y)
    print(f"product vs the figure of real ORC: mid-range rms {rms:.2f} px, max {worst:.2f} px, spread rms {spread:.2f} px")
    assert rms <= 0.5 and worst <= 1.0 and spread <= 2.5, (rms, worst, spread)
    rms_a, worst_a, _ = figure_misfit(fig, figure_analytical(fig, cell_y), cell_y)
    assert rms_a >= 4.0 and worst_a >= 8.0, (rms_a, worst_a)
