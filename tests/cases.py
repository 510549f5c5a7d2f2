"""Shared case builders for the parity tests: the same arrays / settings go to the product and to the oracle."""
import os

import numpy as np

from orc_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_mesh_arrays(name):
    """Connectivity of one of the reference's example meshes (tests/golden/make_golden.py wrote it from /root/reference)."""
    z = np.load(os.path.join(GOLDEN, f"mesh_{name}.npz"), allow_pickle=False)
    m = {k: z[k] for k in z.files}
    m["dims"] = int(m["dims"])
    m["zone_names"] = [str(s) for s in m["zone_names"]]
    return m


def make_pair(oracle, arrays):
    """(product mesh, oracle mesh) from one set of arrays."""
    import orc_b200
    return orc_b200.Mesh.from_arrays(*syn.mesh_args(arrays)), oracle.Mesh.from_arrays(*syn.mesh_args(arrays))


def couette_bcs(mesh, u_wall=5e-4, dp_dx=10.0, wall_zones=("TOP_WALL", "BOTTOM_WALL"), moving="TOP_WALL"):
    """BCs of src/tests.rs:60-76 (solve_channel_flow): walls, PressureInlet p = -dp_dx * L, PressureOutlet 0, Symmetry sides."""
    for zname in wall_zones:
        mesh.set_zone(zname, 3, 0.0, (u_wall if zname == moving else 0.0, 0.0, 0.0))
    mesh.set_zone("INLET", 4, -dp_dx * 0.002, (0.0, 0.0, 0.0))
    mesh.set_zone("OUTLET", 5, 0.0, (0.0, 0.0, 0.0))
    mesh.set_zone("PERIODIC_-Z", 7, 0.0, (0.0, 0.0, 0.0))
    mesh.set_zone("PERIODIC_+Z", 7, 0.0, (0.0, 0.0, 0.0))


def settings_pair(oracle, reference_order=False, **kw):
    """(product NumericalSettings, oracle Settings) with identical values. Keys use the oracle's names.
    reference_order=True selects ORC_REDUCE_REFERENCE_ORDER on the product side (bit-identical solves)."""
    import orc_b200
    from orc_b200 import settings as S
    o = oracle.Settings(**kw)
    ms = orc_b200.MatrixSolverSettings(solver_type=S.SolutionMethod(o.solver_type), iterations=o.iterations, relaxation=o.relaxation,
                                       relative_convergence_threshold=o.threshold, preconditioner=S.PreconditionMethod(o.preconditioner))
    p = orc_b200.NumericalSettings(momentum=S.MomentumDiscretization(o.momentum), limiter=S.TvdLimiter(o.limiter),
                                   pressure_interpolation=S.PressureInterpolation(o.pressure_interpolation),
                                   velocity_interpolation=S.VelocityInterpolation(o.velocity_interpolation),
                                   gradient_reconstruction=S.GradientReconstructionMethods(o.gradient),
                                   pressure_relaxation=o.pressure_relaxation, momentum_relaxation=o.momentum_relaxation, matrix_solver=ms,
                                   mg_smoother=S.SolutionMethod(o.mg_smoother), mg_levels=o.mg_levels,
                                   gs_mode=S.GaussSeidelMode.Lexicographic if o.gs_intended else S.GaussSeidelMode.ReferencePanic)
    p.reduction_mode = S.ReductionMode.ReferenceOrder if reference_order else S.ReductionMode.Fast
    return p, o


def smooth_fields(mesh_export, seed=0, scale_u=1e-3, scale_p=1e-2):
    """Smooth, non-trivial u, v, w, p on the cell centroids (plus seeded noise) for per-call assembly parity."""
    cc = mesh_export["cell_centroid"]
    rng = np.random.default_rng(seed)
    x, y, z = cc[:, 0] / 0.004, cc[:, 1] / 0.001, cc[:, 2] / 0.001
    u = scale_u * (np.sin(3 * x) + y * (1 - y) + 0.1 * rng.standard_normal(x.size))
    v = scale_u * 0.3 * (np.cos(2 * y + x) + 0.1 * rng.standard_normal(x.size))
    w = scale_u * 0.2 * (np.sin(z + 2 * x) + 0.1 * rng.standard_normal(x.size))
    p = scale_p * (1 - x + 0.2 * np.sin(4 * y) + 0.05 * rng.standard_normal(x.size))
    return u, v, w, p


def load_figure(name):
    """Digitised velocity profile of one of the figures the reference ships (real ORC output; made from
    /root/reference/examples/<name>.png by tests/golden/digitise_reference_figures.py, which explains what is stored)."""
    z = np.load(os.path.join(GOLDEN, f"fig_{name}.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def figure_misfit(fig, u, cell_y):
    """A field against a digitised figure, in PIXELS of the figure: (rms, max) of the figure's per-level mid-range minus the
    field's mid-range over the row of cells at that height, and the rms of the blue run's excess length over the shortest run
    minus the field's spread over the row (the scatter plot overlays all cells of a row, see the digitiser)."""
    lev = np.abs(np.asarray(cell_y)[:, None] - fig["y"][None, :]).argmin(axis=1)
    lo = np.array([u[lev == k].min() for k in range(fig["y"].size)])
    hi = np.array([u[lev == k].max() for k in range(fig["y"].size)])
    px = float(fig["u_per_px"])
    d = (fig["u_mid"] - (lo + hi) / 2) / px
    s = (fig["run_px"] - fig["run_px"].min()) - (hi - lo) / px
    return float(np.sqrt((d ** 2).mean())), float(np.abs(d).max()), float(np.sqrt((s ** 2).mean()))


def figure_analytical(fig, y, h=1e-3):
    """write_couette_flow_analytical_profile, src/tests.rs:18-31."""
    return float(fig["u_wall"]) * y / h + 1.0 / (2.0 * float(fig["mu"])) * float(fig["dp_dx"]) * (y ** 2 - h * y)


# ---- the reference's plotting pipeline (examples/plot_output.py:131-199), restated without matplotlib --------------------------
def plot_script_read(data_lines, gradient_lines):
    """What plot_output.py reads back from the two text files of a run: (x, y, p) of `<case>.csv` through its regular expression
    (:139-148) and (x, y, du/dy) of `<case>_gradients.csv` through its splitting (:154-163; du/dy is element [0, 1] of the nine).
    The centroids come back with three significant digits (`{:.2e}`, src/lib.rs:551-556)."""
    import re
    flt = "[\\d|\\.|e|\\-]+"
    vec = f"\\(({flt}),\\s+({flt}),\\s+({flt})\\)"
    pattern = re.compile(f"{vec}\\t{vec}\\t({flt})")
    data, grad = [], []
    for line in data_lines:
        m = pattern.match(line)
        if m:
            g = [float(t) for t in m.groups()]
            data.append((g[0], g[1], g[6]))
    for line in gradient_lines:
        centroid, vel_grad, _ = [s.split(", ") for s in line.rstrip("\n").replace("(", "").replace(")", "").split("\t")]
        grad.append((float(centroid[0]), float(centroid[1]), float(np.reshape(np.array(vel_grad[:9]), (3, 3))[0, 1])))
    return np.array(data), np.array(grad)


def plot_script_inputs(cell_centroid, u, v, w, p, grad_u):
    """plot_script_read of the lines the PRODUCT's host-side writers produce for these fields (orc_b200.io: write_data's line format
    and format_gradient_line with 7 decimals, as src/tests.rs:97-107 calls them); no device needed."""
    from orc_b200 import io as oio
    data = [f"{oio._vector_display(*cell_centroid[i])}\t({oio._rust_exp(u[i])}, {oio._rust_exp(v[i])}, {oio._rust_exp(w[i])})\t{oio._rust_exp(p[i])}"
            for i in range(len(u))]
    grad = [oio.format_gradient_line(cell_centroid[i], np.asarray(grad_u[i]).ravel(), (0.0, 0.0, 0.0), 7) for i in range(len(u))]
    out = plot_script_read(data, grad)
    assert out[0].shape[0] == len(u)
    return out


def interpolate_to_grid(x, y, z, n=200):
    """plot_output.py:121-129: linear interpolation over the Delaunay triangulation onto an n x n grid spanning the points."""
    from scipy.interpolate import LinearNDInterpolator
    xl, yl = np.linspace(x.min(), x.max(), n), np.linspace(y.min(), y.max(), n)
    return xl, yl, LinearNDInterpolator(np.c_[x, y], z)(*np.meshgrid(xl, yl))


def contour_misfit(cont, data, grad):
    """Band edges of the digitised contour figure against the level crossings of the gridded fields, in PIXELS of the figure:
    ((rms, max, n) of the pressure panel along its rows, (rms, max, n) of the du/dy panel along its columns)."""
    def crossings(t, f, level):
        k = np.where((f[:-1] - level) * (f[1:] - level) < 0)[0]
        return [t[i] + (level - f[i]) / (f[i + 1] - f[i]) * (t[i + 1] - t[i]) for i in k]

    out = []
    for (pts, along, across, level, m_per_px) in ((data, cont["p_y"], cont["p_x"], cont["p_level"], float(cont["p_m_per_px"])),
                                                 (grad, cont["g_x"], cont["g_y"], cont["g_level"], float(cont["g_m_per_px"]))):
        pressure_panel = pts is data
        xl, yl, grid = interpolate_to_grid(pts[:, 0], pts[:, 1], pts[:, 2])
        d = []
        for pos in np.unique(along):
            axis = yl if pressure_panel else xl
            j = min(max(np.searchsorted(axis, pos) - 1, 0), axis.size - 2)
            t = (pos - axis[j]) / (axis[j + 1] - axis[j])
            line = (1 - t) * grid[j] + t * grid[j + 1] if pressure_panel else (1 - t) * grid[:, j] + t * grid[:, j + 1]
            sel = along == pos
            for where, lv in zip(across[sel], level[sel]):
                c = crossings(xl if pressure_panel else yl, line, lv)
                if c:
                    d.append((min(c, key=lambda q: abs(q - where)) - where) / m_per_px)
        d = np.array(d)
        out.append((float(np.sqrt((d ** 2).mean())), float(np.abs(d).max()), int(d.size)))
    return tuple(out)
