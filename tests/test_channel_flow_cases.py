"""The mirror of the reference's validation cases (orc_b200/channel_flow.py <- src/tests.rs:1-236). The host glue is checked here on
the CPU: the analytical-profile file and its return values, the reference's `compare`, and the whole case function with stand-ins
for the three device calls (flow initialisation, solve_steady, the gradients of write_gradients) — files, printed lines, verdict.
The device calls themselves have their own GPU tests (test_gpu_initialize.py, test_gpu_steady.py, test_gpu_gradients.py)."""
import os

import numpy as np
import pytest

import orc_b200
from orc_b200 import channel_flow as cf
from orc_b200 import discretization as disc
from orc_b200 import synthetic as syn
from cases import GOLDEN, load_figure, load_mesh_arrays, plot_script_read, contour_misfit, figure_misfit


def test_analytical_profile_file_and_extrema(tmp_path):
    """src/tests.rs:18-42 for the case of src/main.rs:85-102 (moving wall 5e-4, dp/dx = 10): 128 lines `y,u` in `{:.3e}`; the
    returned mean / minimum / maximum are the ones tests/test_oracle_kats.py quotes."""
    path = tmp_path / "a.csv"
    avg, lo, hi = cf.write_couette_flow_analytical_profile(str(path), cf.ChannelFlowParameters(5e-4, 10.0, 1e-3, 1000.0), 1e-3)
    assert abs(avg - (-5.8333e-4)) < 1e-7 and abs(lo - (-1.0125e-3)) < 1e-9 and hi == 5e-4
    lines = path.read_text().splitlines()
    assert len(lines) == 128 and lines[0] == "0.000e0,0.000e0" and lines[64] == "5.000e-4,-1.000e-3"
    y, u = np.array([[float(t) for t in ln.split(",")] for ln in lines]).T
    assert np.allclose(u, 5e-4 * y / 1e-3 + 1 / 2e-3 * 10.0 * (y ** 2 - 1e-3 * y), rtol=0, atol=1e-6)          # y and u are rounded to four digits independently


def test_compare_is_the_reference_ratio_test():
    assert cf.compare(-5.58e-4, -5.83e-4, 0.1) and cf.compare(4.6e-4, 5e-4, 0.1) and not cf.compare(4.4e-4, 5e-4, 0.1)
    assert cf.compare(-1.0, 1.0, 0.1)          # as written: max / min - 1 = -2 for values of opposite sign


def test_couette_case_end_to_end_with_stand_ins_for_the_device(oracle, tmp_path, monkeypatch, capsys):
    """solve_channel_flow on the mesh file of the reference (written back from the committed connectivity), the figure's parameters
    (dp/dx = 5): the stand-in for solve_steady delivers the committed converged fields, so the run files must plot onto the
    figures of real ORC and the reference's own 10 % validation must print "passed"."""
    arrays = dict(load_mesh_arrays("couette_flow_128x64x1"))
    arrays["n_cells"] = 8001
    examples = tmp_path / "examples"
    examples.mkdir()
    syn.write_tgrid(str(examples / "couette_flow_128x64x1.msh"), arrays)
    k = np.load(os.path.join(GOLDEN, "kat_fig_couette.npz"))
    calls = []

    def fake_initialize_flow(mesh, mu, rho, iteration_count, ctx=None, reduction_mode=0):
        calls.append(("initialize_flow", mu, rho, iteration_count))
        return tuple(np.zeros(mesh.n_cells) for _ in range(4))

    def fake_solve_steady(mesh, u, v, w, p, numerics, rho, mu, iteration_count, reporting_interval, ctx=None, on_report=None):
        z = mesh.zones()
        calls.append(("solve_steady", iteration_count, reporting_interval, dict(zip(z["names"], zip(z["types"].tolist(), z["scalar"].tolist())))))
        for a, c in zip((u, v, w, p), "uvwp"):
            a[:] = k[c]

    def fake_gradients(mesh, u, v, w, p, scheme=0, ctx=None):
        om = oracle.Mesh.from_arrays(*syn.mesh_args(arrays))
        return om.gradients(u, v, w, p, int(scheme))

    monkeypatch.setattr(orc_b200.solver, "initialize_flow", fake_initialize_flow)
    monkeypatch.setattr(orc_b200.solver, "solve_steady", fake_solve_steady)
    monkeypatch.setattr(disc, "calculate_gradients", fake_gradients)
    numerics = orc_b200.NumericalSettings(momentum=orc_b200.MomentumDiscretization.TVD, limiter=orc_b200.TVD_UMIST)
    u, v, w, p, passed = cf.solve_channel_flow(600, 0, cf.ChannelFlowParameters(5e-4, 5.0, 1e-3, 1000.0), numerics, "couette_flow", 0.1,
                                               examples_dir=str(examples))
    out = capsys.readouterr().out
    assert passed and "couette_flow validation passed." in out and "**FAIL**" not in out
    assert " U_min: \tCFD = -3.95e-4; Analytical = -4.00e-4; Error =   -1.2%" in out
    assert calls[0] == ("initialize_flow", 1e-3, 1000.0, 1000) and calls[1][:3] == ("solve_steady", 600, 1)
    zones = calls[1][3]
    assert zones["INLET"] == (4, -0.01) and zones["OUTLET"] == (5, 0.0) and zones["TOP_WALL"][0] == 3 and zones["PERIODIC_+Z"][0] == 7
    for name in ("couette_flow.csv", "couette_flow_gradients.csv", "couette_flow_analytical.csv"):
        assert (examples / name).exists()
    data, grad = plot_script_read(open(examples / "couette_flow.csv").readlines(), open(examples / "couette_flow_gradients.csv").readlines())
    (p_rms, p_max, _), (g_rms, g_max, _) = contour_misfit(load_figure("couette_flow_contour_plots"), data, grad)
    assert p_rms <= 1.5 and p_max <= 4.0 and g_rms <= 0.5 and g_max <= 1.0
    # a second call finds the data file and restarts from it instead of initialising (src/tests.rs:84-86)
    calls.clear()
    cf.solve_channel_flow(1, 1, cf.ChannelFlowParameters(5e-4, 5.0, 1e-3, 1000.0), numerics, "couette_flow", 0.1, examples_dir=str(examples))
    assert [c[0] for c in calls] == ["solve_steady"]


def test_velocity_inlet_case_with_stand_ins_for_the_device(oracle, tmp_path, monkeypatch, capsys):
    """solve_channel_flow_velocity_inlet (src/tests.rs:153-236, the case of the reference's current `main`): VelocityInlet zone with
    its vector, initialize_flow_new for the cold start, the two run files and the three printed lines."""
    arrays = dict(load_mesh_arrays("couette_flow_128x64x1"))
    arrays["n_cells"] = 8001
    examples = tmp_path / "examples"
    examples.mkdir()
    syn.write_tgrid(str(examples / "couette_flow_128x64x1.msh"), arrays)
    seen = {}

    def fake_initialize_flow_new(mesh, mu, rho, iteration_count, ctx=None, reduction_mode=2):
        seen["init"] = (mu, rho, iteration_count)
        return tuple(np.full(mesh.n_cells, 1e-3 if k == 0 else 0.0) for k in range(4))

    def fake_solve_steady(mesh, u, v, w, p, numerics, rho, mu, iteration_count, reporting_interval, ctx=None, on_report=None):
        z = mesh.zones()
        k = z["names"].index("INLET")
        seen["inlet"] = (int(z["types"][k]), tuple(z["vector"][k]))
        seen["solve"] = (iteration_count, reporting_interval)
        u *= np.linspace(0.5, 1.5, u.size)

    def fake_gradients(mesh, u, v, w, p, scheme=0, ctx=None):
        return oracle.Mesh.from_arrays(*syn.mesh_args(arrays)).gradients(u, v, w, p, int(scheme))

    monkeypatch.setattr(orc_b200.solver, "initialize_flow_new", fake_initialize_flow_new)
    monkeypatch.setattr(orc_b200.solver, "solve_steady", fake_solve_steady)
    monkeypatch.setattr(disc, "calculate_gradients", fake_gradients)
    u, v, w, p = cf.solve_channel_flow_velocity_inlet(5, 2, orc_b200.NumericalSettings(), "channel_flow_velocity_inlet", 0.0, 1e-3, 0.001, 1000.0,
                                                      examples_dir=str(examples))
    out = capsys.readouterr().out
    assert seen["init"] == (0.001, 1000.0, 1000) and seen["solve"] == (5, 2) and seen["inlet"] == (10, (1e-3, 0.0, 0.0))
    assert " U_mean:\tCFD = 1.00e-3" in out and " U_min: \tCFD = 5.00e-4" in out and " U_max: \tCFD = 1.50e-3" in out
    assert (examples / "channel_flow_velocity_inlet.csv").exists() and (examples / "channel_flow_velocity_inlet_gradients.csv").exists()
    assert np.array_equal(orc_b200.read_data(str(examples / "channel_flow_velocity_inlet.csv"))[0], u)     # `{:e}` round-trips every bit
