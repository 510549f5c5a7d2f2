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

The two contour figures of the same runs (examples/<case>_flow_contour_plots.png, `plot_output.py:131-199`) hold two more fields:
the upper panel is `contourf` of the pressure, the lower one of du/dy as `write_gradients` stored it (element [0, 1] of the 3 x 3
tensor of `<case>_gradients.csv`). The script read both from the TEXT files (`write_data`: centroids with `{:.2e}`, i.e. three
significant digits; values with `{:e}`), interpolated them linearly over the Delaunay triangulation of the rounded centroids onto a
200 x 200 grid and filled between "nice" levels; the fills are not anti-aliased, so a band edge is known to +-0.5 px. What is stored:
for a set of pixel rows of the pressure panel (those no quiver arrow touches) the x of every band edge, for a set of pixel columns
of the du/dy panel the y of every band edge, both in metres through the tick calibration, each with its level — identified through
the colour bar (band colours and tick marks), not assumed. One pixel is 1.4e-6 m in x (6.9e-6 Pa at dp/dx = 5, 0.07 % of the
pressure range) and 1.9e-6 m in y (0.0094 1/s, 0.18 % of the du/dy range) on the Couette figure.

Run here (needs /root/reference and PIL):   python tests/golden/digitise_reference_figures.py [--check ITERATIONS] [--converged]
--check also runs the oracle on the case of each figure and prints the misfit (DESIGN.md section 2 quotes these numbers);
--converged writes tests/golden/kat_fig_<case>.npz: the oracle's fields after 600 (Couette, 8001 cells) / 1000 (channel, 1008 cells)
SIMPLE iterations from rest, about two minutes of one core, which tests/test_reference_figures.py compares with the figures.
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
# The Couette figures show 127 x 63 cells (couette_flow_128x64x1.msh); the channel figures show 16 x 63 (the frame of the contour
# panels spans x = 6.25e-5 .. 1.9375e-3, the cell centres of channel_flow.msh), with the same 63 heights.
FIGURES = {
    "couette_flow_velocity_profile": dict(x_ticks=(0.0, 1e-3), u_ticks=(4e-4, -4e-4), u_wall=5e-4, dp_dx=5.0, mu=1e-3, rho=1000.0,
                                          mesh="couette_flow_128x64x1", walls=("TOP_WALL", "BOTTOM_WALL"), moving="TOP_WALL", iterations=600),
    "channel_flow_velocity_profile": dict(x_ticks=(0.0, 1e-3), u_ticks=(6e-3, 0.0), u_wall=0.0, dp_dx=-5000.0, mu=0.1, rho=1000.0,
                                          mesh="channel_flow", walls=("WALL",), moving=None, iterations=1000),
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
                u_wall=np.float64(spec["u_wall"]), dp_dx=np.float64(spec["dp_dx"]), mu=np.float64(spec["mu"]), rho=np.float64(spec["rho"]),
                mesh=np.str_(spec["mesh"]), walls=np.array(spec["walls"]), moving=np.str_(spec["moving"] or ""))


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


def oracle_run(spec, iterations, momentum="umist", all_fields=False):
    from oracle import pyoracle as po
    from cases import couette_bcs, load_mesh_arrays
    from orc_b200 import synthetic as syn
    m = po.Mesh.from_arrays(*syn.mesh_args(load_mesh_arrays(spec["mesh"])))
    couette_bcs(m, u_wall=spec["u_wall"], dp_dx=spec["dp_dx"], wall_zones=spec["walls"], moving=spec["moving"])
    z = np.zeros(m.n_cells)
    s = po.Settings(momentum=po.TVD, limiter=po.PSI_UMIST) if momentum == "umist" else po.Settings()
    u, v, w, p = m.solve_steady(z, z, z, z, s, spec["rho"], spec["mu"], iterations, 0)[:4]
    return (u, v, w, p) if all_fields else (u, m.export()["cell_centroid"][:, 1])


# name, panel tick labels read off the figure: pressure colour bar (top, bottom, step), du/dy colour bar (top, bottom, step),
# x ticks (first, last) and y ticks (top, bottom) of the panels in m
CONTOURS = {
    "couette_flow_contour_plots": dict(p_bar=(0.0, -0.010, 0.001), g_bar=(3.0, -2.25, 0.25), x_ticks=(0.25e-3, 1.75e-3), y_ticks=(8e-4, 2e-4)),
    "channel_flow_contour_plots": dict(p_bar=(10.0, 0.0, 1.0), g_bar=(22.5, -22.5, 2.5), x_ticks=(0.25e-3, 1.75e-3), y_ticks=(8e-4, 2e-4)),
}


def _colour_runs(line):
    """[(first, last, (r, g, b))] of the runs of one colour along a line of pixels; runs shorter than 3 px are dropped."""
    out, s = [], 0
    for i in range(1, len(line) + 1):
        if i == len(line) or tuple(line[i]) != tuple(line[s]):
            if i - s >= 3:
                out.append((s, i - 1, tuple(int(c) for c in line[s])))
            s = i
    return out


def _band_edges(line, offset, bands):
    """(pixel position, level) of every edge between two colour-bar colours along a line of pixels."""
    runs = _colour_runs(line)
    out = []
    for (s0, e0, c0), (s1, e1, c1) in zip(runs[:-1], runs[1:]):
        if c0 in bands and c1 in bands:
            shared = set(bands[c0]) & set(bands[c1])
            if len(shared) == 1:
                out.append(((e0 + s1) / 2 + offset, shared.pop()))
    return out


def digitise_contours(name, spec):
    from PIL import Image
    a = np.asarray(Image.open(os.path.join(REF, name + ".png")).convert("RGB")).astype(int)
    blk = a.max(axis=2) < 60
    fc = _runs(np.where(blk.sum(axis=0) > 400)[0])           # left, right of the panels; left, right of the colour bars
    fr = _runs(np.where(blk.sum(axis=1) > 1200)[0])          # top, bottom of the upper panel; top, bottom of the lower panel
    left, right, cb0, cb1 = [int(round(v)) for v in fc[:4]]
    panels = []
    for (top, bot), (v_top, v_bot, step) in zip(((fr[0], fr[1]), (fr[2], fr[3])), (spec["p_bar"], spec["g_bar"])):
        top, bot = int(round(top)), int(round(bot))
        xt = _runs(np.where(blk[bot + 4:bot + 12, left:right + 1].sum(axis=0) >= 6)[0] + left)
        yt = _runs(np.where(blk[top:bot + 1, left - 11:left - 3].sum(axis=1) >= 6)[0] + top)
        ct = _runs(np.where(blk[top - 2:bot + 3, cb1 + 4:cb1 + 12].sum(axis=1) >= 6)[0] + top - 2)
        bar = [(s + top + 3, e + top + 3, c) for s, e, c in _colour_runs(a[top + 3:bot - 2, (cb0 + cb1) // 2])]
        val = lambda r, ct=ct, v_top=v_top, v_bot=v_bot: v_top + (r - ct[0]) / (ct[-1] - ct[0]) * (v_bot - v_top)
        bands = {}
        for k, (s, e, c) in enumerate(bar):       # the first / last band run up to the frame of the bar
            hi = val(s - 0.5) if k > 0 else val(top)
            lo = val(e + 0.5) if k < len(bar) - 1 else val(bot)
            bands[c] = (round(lo / step) * step, round(hi / step) * step)
        x_of = lambda c, xt=xt: spec["x_ticks"][0] + (c - xt[0]) / (xt[-1] - xt[0]) * (spec["x_ticks"][1] - spec["x_ticks"][0])
        y_of = lambda r, yt=yt: spec["y_ticks"][0] + (r - yt[0]) / (yt[-1] - yt[0]) * (spec["y_ticks"][1] - spec["y_ticks"][0])
        panels.append(dict(top=top, bot=bot, bands=bands, x_of=x_of, y_of=y_of,
                           m_per_px_x=abs(x_of(1) - x_of(0)), m_per_px_y=abs(y_of(1) - y_of(0))))
    P, G = panels
    # pressure panel: rows no quiver arrow touches (black pixels), up to 24 of them spread over the height
    inner = a[P["top"] + 3:P["bot"] - 2, left + 3:right - 2]
    free = np.where((inner.max(axis=2) < 60).sum(axis=1) == 0)[0] + P["top"] + 3
    p_pts = [(P["y_of"](r), P["x_of"](px), lv) for r in free[np.linspace(0, free.size - 1, min(24, free.size)).astype(int)] for px, lv in _band_edges(a[r, left + 3:right - 2], left + 3, P["bands"])]
    g_pts = [(G["x_of"](c), G["y_of"](px), lv) for c in range(left + 40, right - 40, 100)
             for px, lv in _band_edges(a[G["top"] + 3:G["bot"] - 2, c], G["top"] + 3, G["bands"])]
    p_pts, g_pts = np.array(p_pts), np.array(g_pts)
    return dict(p_y=p_pts[:, 0], p_x=p_pts[:, 1], p_level=p_pts[:, 2], g_x=g_pts[:, 0], g_y=g_pts[:, 1], g_level=g_pts[:, 2],
                p_m_per_px=np.float64(P["m_per_px_x"]), g_m_per_px=np.float64(G["m_per_px_y"]))


if __name__ == "__main__":
    levels = y_levels()
    for name, spec in FIGURES.items():
        fig = digitise(name, spec, levels)
        np.savez(os.path.join(HERE, f"fig_{name}.npz"), **fig)
        ana = analytical(fig, fig["y"])
        d = (fig["u_mid"] - ana) / float(fig["u_per_px"])
        print(f"{name}: {levels.size} levels, {float(fig['u_per_px']):.3e} m/s per pixel, run length {fig['run_px'].min():.1f}..{fig['run_px'].max():.1f} px, "
              f"figure - analytical: rms {np.sqrt((d ** 2).mean()):.2f} px, max {np.abs(d).max():.2f} px")
        cname = name.replace("velocity_profile", "contour_plots")
        cont = digitise_contours(cname, CONTOURS[cname])
        np.savez(os.path.join(HERE, f"fig_{cname}.npz"), **cont)
        print(f"{cname}: {cont['p_x'].size} pressure band edges on {np.unique(cont['p_y']).size} rows (levels {np.unique(cont['p_level']).size}), "
              f"{cont['g_y'].size} du/dy band edges on {np.unique(cont['g_x']).size} columns (levels {np.unique(cont['g_level']).size})")
        if "--converged" in sys.argv:
            u, v, w, p = oracle_run(spec, spec["iterations"], "umist", all_fields=True)
            np.savez_compressed(os.path.join(HERE, f"kat_fig_{name.split('_')[0]}.npz"), u=u, v=v, w=w, p=p, iters=spec["iterations"])
        if "--check" in sys.argv:
            its = int(sys.argv[sys.argv.index("--check") + 1])
            for mom in ("umist", "cd1"):
                u, cy = oracle_run(spec, its, mom)
                print(f"  oracle, {its} iterations from rest, {mom}: mid-range rms %.2f px, max %.2f px, spread rms %.2f px" % misfit(fig, u, cy))
