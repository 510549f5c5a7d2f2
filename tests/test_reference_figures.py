"""CPU tests against the only outputs of REAL ORC that exist in this container: the four figures the reference ships
(examples/{couette,channel}_flow_{velocity_profile,contour_plots}.png, README.md "Validation"), digitised by
tests/golden/digitise_reference_figures.py into tests/golden/fig_*.npz. They show three fields of a converged run on
couette_flow_128x64x1.msh (Couette figures: 127 x 63 cells) and channel_flow.msh (channel figures: 16 x 63 cells) — u over y of
every cell, the pressure contours, and the du/dy contours of `write_gradients` — at a resolution of 0.07-0.2 % of each field's
range. tests/golden/kat_fig_<case>.npz holds the oracle's fields after 600 / 1000 SIMPLE iterations from rest on the same two
cases (src/main.rs:64-102 / src/tests.rs:44-108: TVD-UMIST, SecondOrder, Rhie-Chow, Multigrid);
the checks below put them through the product's host-side writers (orc_b200.io) and a matplotlib-free restatement of the reference's
plotting script (examples/plot_output.py:121-219), and compare with the pixels of the figures.
(The CUDA path is compared with the Couette profile figure in tests/test_gpu_golden.py.)"""
import os

import numpy as np
import pytest

from cases import (GOLDEN, contour_misfit, couette_bcs, figure_analytical, figure_misfit, load_figure, load_mesh_arrays, plot_script_inputs,
                   plot_script_read)
from orc_b200 import synthetic as syn

CASES = ["couette", "channel"]


def _case(oracle, case):
    fig = load_figure(f"{case}_flow_velocity_profile")
    m = oracle.Mesh.from_arrays(*syn.mesh_args(load_mesh_arrays(str(fig["mesh"]))))
    couette_bcs(m, u_wall=float(fig["u_wall"]), dp_dx=float(fig["dp_dx"]), wall_zones=tuple(str(z) for z in fig["walls"]),
                moving=str(fig["moving"]) or None)
    k = np.load(os.path.join(GOLDEN, f"kat_fig_{case}.npz"))
    return fig, m, k


@pytest.mark.parametrize("case", CASES)
def test_converged_oracle_fields_land_on_the_figures_of_real_orc(oracle, tmp_path, case):
    """Velocity profile: per-level mid-range within 0.6 px rms / 1.2 px max of the marker blobs, the rows' inlet-to-outlet spread
    within 2 px rms. Pressure contours: every band edge within 1.5 px rms / 4 px max (Couette: 1 px = 1.4e-6 m = 0.07 % of the
    pressure range). du/dy contours: Couette within 0.5 px rms / 1 px max — the +-0.5 px of a fill that is not anti-aliased (1 px =
    0.18 % of the du/dy range); channel (16 cells along x, an earlier version of the plotting script) within 1.3 px rms / 2 px max.
    Measured, px: Couette 0.47 / 1.00 / 1.11, 0.98 / 3.08, 0.28 / 0.59; channel 0.27 / 0.58 / 1.73, 0.27 / 0.52, 1.02 / 1.62."""
    fig, m, k = _case(oracle, case)
    cc = m.export()["cell_centroid"]
    rms, worst, spread = figure_misfit(fig, k["u"], cc[:, 1])
    assert rms <= 0.6 and worst <= 1.2 and spread <= 2.0, (rms, worst, spread)
    _, gu = m.gradients(k["u"], k["v"], k["w"], k["p"], 0)
    data, grad = plot_script_inputs(cc, k["u"], k["v"], k["w"], k["p"], gu)
    # the data file as the product's write_data puts it on disk (host-side mirror of src/io.rs:572-591) reads back the same
    import orc_b200
    pm = orc_b200.Mesh.from_arrays(*syn.mesh_args(load_mesh_arrays(str(fig["mesh"]))))
    orc_b200.write_data(pm, k["u"], k["v"], k["w"], k["p"], str(tmp_path / "run.csv"))
    assert np.array_equal(plot_script_read(open(tmp_path / "run.csv").readlines(), [])[0], data)
    (p_rms, p_max, p_n), (g_rms, g_max, g_n) = contour_misfit(load_figure(f"{case}_flow_contour_plots"), data, grad)
    print(f"{case}: profile {rms:.2f} / {worst:.2f} / {spread:.2f} px, pressure {p_rms:.2f} / {p_max:.2f} px ({p_n} edges), "
          f"du/dy {g_rms:.2f} / {g_max:.2f} px ({g_n} edges)")
    assert p_n >= 200 and g_n >= 250
    assert p_rms <= 1.5 and p_max <= 4.0, (p_rms, p_max)
    assert (g_rms <= 0.5 and g_max <= 1.0) if case == "couette" else (g_rms <= 1.3 and g_max <= 2.0), (g_rms, g_max)


def test_the_figures_resolve_orc_from_the_exact_solution_and_from_another_file_format(oracle):
    """Resolving power of the fixtures (Couette case): the exact solution of the continuous problem — parabolic u, linear p, linear
    du/dy — put through the same writers and plotting pipeline misses the figures by several times the bounds above, and so do the
    oracle's fields when the centroids are NOT rounded to three significant digits (the `{:.2e}` of `impl Display for Vector`,
    src/lib.rs:551-556, which moves the pressure band edges by up to 5 px on the right half of the figure, where x >= 1e-3)."""
    fig, m, k = _case(oracle, "couette")
    cont = load_figure("couette_flow_contour_plots")
    cc = m.export()["cell_centroid"]
    x, y, h = cc[:, 0], cc[:, 1], 1e-3
    u_exact = figure_analytical(fig, y)
    gu_exact = np.zeros((y.size, 3, 3))
    gu_exact[:, 0, 1] = float(fig["u_wall"]) / h + 1.0 / (2.0 * float(fig["mu"])) * float(fig["dp_dx"]) * (2.0 * y - h)
    p_exact = -float(fig["dp_dx"]) * 0.002 + float(fig["dp_dx"]) * x
    zero = np.zeros_like(y)
    rms, worst, _ = figure_misfit(fig, u_exact, y)
    assert rms >= 4.0 and worst >= 8.0
    (p_rms, p_max, _), (g_rms, g_max, _) = contour_misfit(cont, *plot_script_inputs(cc, u_exact, zero, zero, p_exact, gu_exact))
    assert p_rms >= 5.0 and p_max >= 10.0 and g_rms >= 1.0 and g_max >= 3.0, (p_rms, p_max, g_rms, g_max)
    _, gu = m.gradients(k["u"], k["v"], k["w"], k["p"], 0)
    data, grad = plot_script_inputs(cc, k["u"], k["v"], k["w"], k["p"], gu)
    data[:, 0], data[:, 1] = x, y
    (p_rms, p_max, _), _ = contour_misfit(cont, data, grad)
    assert p_rms >= 1.7 and p_max >= 5.0, (p_rms, p_max)


@pytest.mark.parametrize("case", CASES)
def test_the_committed_fields_are_the_oracles_converged_state(oracle, case):
    """kat_fig_<case>.npz is what tests/golden/digitise_reference_figures.py --converged wrote: 600 / 1000 iterations from rest.
    Repeating them here would cost two minutes; instead the oracle continues from the committed fields for 8 iterations (the first
    of which sees unit Rhie-Chow diagonals, src/solver.rs:43-45) and must stay put: u moves by less than 0.1 px of the profile figure."""
    fig, m, k = _case(oracle, case)
    u = m.solve_steady(k["u"], k["v"], k["w"], k["p"], oracle.Settings(momentum=oracle.TVD, limiter=oracle.PSI_UMIST),
                       float(fig["rho"]), float(fig["mu"]), 8, 0)[0]
    assert np.abs(u - k["u"]).max() <= 0.1 * float(fig["u_per_px"]), np.abs(u - k["u"]).max() / float(fig["u_per_px"])
