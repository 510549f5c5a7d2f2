"""Digitises the two velocity-profile figures the reference ships (outputs of REAL ORC runs, README.md "Validation"):

    /root/reference/examples/couette_flow_velocity_profile.png     (moving top wall 5e-4 m/s, p_inlet = -0.01 Pa: dp/dx = 5)
    /root/reference/examples/channel_flow_velocity_profile.png     (both walls at rest, mu = 0.1, p_inlet = 10 Pa: dp/dx = -5000)

and writes tests/golden/fig_<name>.npz. These figures are the only output of the real reference that exists in this container
(there is no Rust toolchain, so ORC itself cannot run here); `examples/plot_output.py:204-219` made them with
`ax.scatter(y, u)` over ALL cells of couette_flow_128x64x1.msh (127 x 63 cells: 63 distinct y levels) at 300 dpi, so every y level
shows as the union of 127 discs of one diameter centred on u(x_j, y). At the pixel column of a level the neighbouring levels' discs
do not reach (level spacing 21.6 px, disc radius ~15.5 px), hence the blue run of that column is the union of the intervals
[u_j - r, u_j + r]: its middle is the MID-RANGE of u over the row of cells and its length is 2 r + (max_j u_j - min_j u_j).
The analytical line drawn over the markers has the same colour and passes inside them; it does not move the run's ends.

What is stored per level: y, the mid-range velocity `u_mid` and the run length `run_px` (sub-pixel: linear interpolation of the
0.5 coverage crossing of the anti-aliased edge), plus the axis calibration (tick-mark pixel positions, m/s per pixel).
One pixel is 8.7e-7 m/s on the Couette figure (0.1 % of its velocity range), 6.0e-6 m/s on the channel figure.

Run here (needs /root/reference and PIL):   python tests/golden/digitise_reference_figures.py [--check ITERATIONS]
--check also runs the oracle on the case of each figure and prints the misfit (DESIGN.md section 2 quotes these numbers).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF = "/root/reference/examples"
# name, x ticks (first, last) in m, u ticks (top, bottom) in m/s: read off the tick labels of the figures
FIGURES = {
    "couette_flow_velocity_profile": dict(x_ticks=(0.0, 1e-3), u_ticks=(4e-4, -4e-4), u_wall=5e-4, dp_dx=5.0, mu=1e-3, rho=1000.0),
    "channel_flow_velocity_profile": dict(x_ticks=(0.0, 1e-3), u_ticks=(6e-3, 0.0), u_wall=0.0, dp_dx=-5000.0, mu=0.1, rho=1000.0),
}


def _runs(v):
    """Centres of the runs of consecutive integers in v."""
    out, start = [], 0
    for k in range(1, len(v) + 1):
        if k == len(v) or v[k] != v[k - 1] + 1:
            out.append(float(np.mean(v[start:k])))
            start = k
    return out


def y_levels():
    """The 63 distinct cell-centre heights of couette_flow_128x64x1.msh (oracle geometry, io.rs:289-438)."""
    from oracle import pyoracle as po
    from cases import load_mesh_arrays
    from orc_b200 import synthetic as syn
    m = po.Mesh.from_arrays(*syn.mesh_args(load_mesh_arrays("couette_flow_128x64x1")))
    cc = m.export()["cell_centroid"]
    return np.unique(np.round(cc[:, 1], 10))


def digitise(name, spec, levels):
    from PIL import Image
    a = np.asarray(Image.open(os.path.join(REF, name + ".png")).convert("RGB")).astype(float)
    # marker colour is matplotlib's C0 = (31, 119, 180): the red channel is the anti-aliased coverage; text / frame / ticks are grey
    cov = np.clip((255.0 - a[:, :, 0]) / (255.0 - 31.0), 0.0, 1.0)
    cov[(np.abs(a[:, :, 0] - a[:, :, 1]) < 12) & (np.abs(a[:, :, 1] - a[:, :, 2]) < 12)] = 0.0
    dark = a.sum(axis=2) < 200
    frame_cols = _runs(np.where(dark.sum(axis=0) > 800)[0])
    frame_rows = _runs(np.where(dark.sum(axis=1) > 1000)[0])
    left, bottom = int(frame_cols[0]), int(frame_rows[-1])
    # tick marks: short dark segments just outside the frame
    xt = _runs(np.where(dark[bottom + 4:bottom + 14, :].sum(axis=0) >= 8)[0])
    ut = _runs(np.where(dark[:, left - 12:left - 2].sum(axis=1) >= 8)[0])
    x0, x1 = xt[0], xt[-1]
    r_top, r_bot = ut[0], ut[-1]
    u_top, u_bot = spec["u_ticks"]
    u_per_px = (u_top - u_bot) / (r_bot - r_top)
    out = []
    for y in levels:
        c = int(round(x0 + (y - spec["x_ticks"][0]) / (spec["x_ticks"][1] - spec["x_ticks"][0]) * (x1 - x0)))
        prof = cov[:, c - 1:c + 2].mean(axis=1)
        rows = np.where(prof > 0.5)[0]
        run = max(np.split(rows, np.where(np.diff(rows) > 1)[0] + 1), key=len)
        t, b = int(run[0]), int(run[-1])
        te = t - (prof[t] - 0.5) / (prof[t] - prof[t - 1])
        be = b + (prof[b] - 0.5) / (prof[b] - prof[b + 1])
        out.append((y, u_top - ((te + be) / 2 - r_top) * u_per_px, be - te))
    out = np.array(out)
    return dict(y=out[:, 0], u_mid=out[:, 1], run_px=out[:, 2], u_per_px=np.float64(u_per_px), x_tick_px=np.array(xt), u_tick_px=np.array(ut),
                u_wall=np.float64(spec["u_wall"]), dp_dx=np.float64(spec["dp_dx"]), mu=np.float64(spec["mu"]), rho=np.float64(spec["rho"]))


def misfit(fig, u, cell_y):
    """(rms, max) of figure mid-range minus the field's mid-range per level, and the rms of (run length - shortest run) minus the
    field's spread per level, all in pixels of the figure."""
    lev = np.abs(cell_y[:, None] - fig["y"][None, :]).argmin(axis=1)
    lo = np.array([u[lev == k].min() for k in range(fig["y"].size)])
    hi = np.array([u[lev == k].max() for k in range(fig["y"].size)])
    px = float(fig["u_per_px"])
    d = (fig["u_mid"] - (lo + hi) / 2) / px
    s = (fig["run_px"] - fig["run_px"].min()) - (hi - lo) / px
    return float(np.sqrt((d ** 2).mean())), float(np.abs(d).max()), float(np.sqrt((s ** 2).mean()))


def analytical(fig, y, h=1e-3):
    """write_couette_flow_analytical_profile, src/tests.rs:18-31."""
    return float(fig["u_wall"]) * y / h + 1.0 / (2.0 * float(fig["mu"])) * float(fig["dp_dx"]) * (y ** 2 - h * y)


def oracle_run(fig, iterations, momentum="umist"):
    from oracle import pyoracle as po
    from cases import couette_bcs, load_mesh_arrays
    from orc_b200 import synthetic as syn
    m = po.Mesh.from_arrays(*syn.mesh_args(load_mesh_arrays("couette_flow_128x64x1")))
    couette_bcs(m, u_wall=float(fig["u_wall"]), dp_dx=float(fig["dp_dx"]))
    z = np.zeros(m.n_cells)
    s = po.Settings(momentum=po.TVD, limiter=po.PSI_UMIST) if momentum == "umist" else po.Settings()
    u = m.solve_steady(z, z, z, z, s, float(fig["rho"]), float(fig["mu"]), iterations, 0)[0]
    return u, m.export()["cell_centroid"][:, 1]


if __name__ == "__main__":
    levels = y_levels()
    for name, spec in FIGURES.items():
        fig = digitise(name, spec, levels)
        np.savez(os.path.join(HERE, f"fig_{name}.npz"), **fig)
        ana = analytical(fig, fig["y"])
        d = (fig["u_mid"] - ana) / float(fig["u_per_px"])
        print(f"{name}: {levels.size} levels, {float(fig['u_per_px']):.3e} m/s per pixel, run length {fig['run_px'].min():.1f}..{fig['run_px'].max():.1f} px, "
              f"figure - analytical: rms {np.sqrt((d ** 2).mean()):.2f} px, max {np.abs(d).max():.2f} px")
        if "--check" in sys.argv:
            its = int(sys.argv[sys.argv.index("--check") + 1])
            for mom in ("umist", "cd1"):
                u, cy = oracle_run(fig, its, mom)
                print(f"  oracle, {its} iterations from rest, {mom}: mid-range rms %.2f px, max %.2f px, spread rms %.2f px" % misfit(fig, u, cy))
