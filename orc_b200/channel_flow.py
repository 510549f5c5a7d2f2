"""The reference's two validation cases (`pub mod channel_flow`, src/tests.rs:1-236) over this package's mirrors of the calls they
make: same names, arguments, boundary conditions, files and printed lines, so that `src/main.rs:64-113` reads the same here:

    from orc_b200 import NumericalSettings, MomentumDiscretization, TVD_UMIST
    from orc_b200.channel_flow import ChannelFlowParameters, solve_channel_flow
    solve_channel_flow(iteration_count, reporting_interval,
                       ChannelFlowParameters(top_wall_velocity=5e-4, dp_dx=10., mu=0.001, rho=1000.),
                       NumericalSettings(momentum=MomentumDiscretization.TVD, limiter=TVD_UMIST), "couette_flow", 0.1)

The hot path (initialize_flow, solve_steady, the gradients of write_gradients) runs on the GPU through the C ABI; this module is
host glue only. `examples_dir` / `mesh_path` default to the reference's relative paths (it is run from its repository root)."""
import dataclasses
import os

import numpy as np

from . import io as _io
from . import solver as _solver
from .mesh import FaceConditionTypes


@dataclasses.dataclass
class ChannelFlowParameters:  # src/tests.rs:11-16
    top_wall_velocity: float
    dp_dx: float
    mu: float
    rho: float


def write_couette_flow_analytical_profile(output_path, parameters, channel_height):
    """src/tests.rs:18-42: 128 lines `y,u` (`{:.3e}` each) of the Couette-Poiseuille profile, -> (u_avg, u_min, u_max)."""
    n = 128
    with open(output_path, "w") as f:
        for i in range(n):
            y = float(i) / float(n) * channel_height
            u = parameters.top_wall_velocity * y / channel_height + 1.0 / (2.0 * parameters.mu) * parameters.dp_dx * (y ** 2 - channel_height * y)
            f.write(f"{_io._rust_exp(y, 3)},{_io._rust_exp(u, 3)}\n")
    u_extremum = -(2.0 * parameters.mu * parameters.top_wall_velocity - channel_height ** 2 * parameters.dp_dx) ** 2 \
        / (8.0 * channel_height ** 2 * parameters.dp_dx * parameters.mu)
    u_avg = parameters.top_wall_velocity / 2.0 - channel_height ** 2 / (12.0 * parameters.mu) * parameters.dp_dx
    u_max = max(max(parameters.top_wall_velocity, 0.0), u_extremum)
    u_min = min(min(parameters.top_wall_velocity, 0.0), u_extremum)
    return u_avg, u_min, u_max


def compare(value_1, value_2, tolerance):
    """src/tests.rs:115-117, as written (a ratio of the larger to the smaller value, signs included)."""
    return max(value_1, value_2) / min(value_1, value_2) - 1.0 < tolerance


def _set_common_zones(mesh, top_wall_velocity):
    mesh.get_face_zone("TOP_WALL").zone_type = FaceConditionTypes.Wall
    mesh.get_face_zone("TOP_WALL").vector_value = (top_wall_velocity, 0.0, 0.0)
    mesh.get_face_zone("BOTTOM_WALL").zone_type = FaceConditionTypes.Wall
    mesh.get_face_zone("OUTLET").zone_type = FaceConditionTypes.PressureOutlet
    mesh.get_face_zone("OUTLET").scalar_value = 0.0
    mesh.get_face_zone("PERIODIC_-Z").zone_type = FaceConditionTypes.Symmetry
    mesh.get_face_zone("PERIODIC_+Z").zone_type = FaceConditionTypes.Symmetry


def _rust_sci(x, width, precision):   # `{:>W.Pe}`
    return _io._rust_exp(x, precision).rjust(width)


def solve_channel_flow(iteration_count, reporting_interval, flow_parameters, numerics, name, validation_threshold,
                       examples_dir="./examples", mesh_path=None, ctx=None):
    """src/tests.rs:44-151. Returns (u, v, w, p, passed) on top of what the reference does (it returns nothing)."""
    channel_height, dx = 0.001, 0.002
    mesh = _io.read_mesh(mesh_path or os.path.join(examples_dir, "couette_flow_128x64x1.msh"))
    _set_common_zones(mesh, flow_parameters.top_wall_velocity)
    mesh.get_face_zone("INLET").zone_type = FaceConditionTypes.PressureInlet
    mesh.get_face_zone("INLET").scalar_value = -flow_parameters.dp_dx * dx
    data_path = os.path.join(examples_dir, f"{name}.csv")
    analytical_path = os.path.join(examples_dir, f"{name}_analytical.csv")
    gradients_path = os.path.join(examples_dir, f"{name}_gradients.csv")
    try:
        u, v, w, p = _io.read_data(data_path)
    except OSError:
        u, v, w, p = _solver.initialize_flow(mesh, flow_parameters.mu, flow_parameters.rho, 1000, ctx=ctx)
    u, v, w, p = (np.ascontiguousarray(a, dtype=np.float64) for a in (u, v, w, p))
    _solver.solve_steady(mesh, u, v, w, p, numerics, flow_parameters.rho, flow_parameters.mu, iteration_count, max(reporting_interval, 1), ctx=ctx)
    _io.write_data(mesh, u, v, w, p, data_path)
    _io.write_gradients(mesh, u, v, w, p, gradients_path, 7, numerics.gradient_reconstruction, ctx)
    u_mean_a, u_min_a, u_max_a = write_couette_flow_analytical_profile(analytical_path, flow_parameters, channel_height)
    u_mean, u_min, u_max = float(np.sum(u) / u.size), float(u.min()), float(u.max())
    ok = []
    for label, got, exact in ((" U_mean:\t", u_mean, u_mean_a), (" U_min: \t", u_min, u_min_a), (" U_max: \t", u_max, u_max_a)):
        ok.append(compare(got, exact, validation_threshold))
        print(f"{label}CFD = {_rust_sci(got, 8, 2)}; Analytical = {_rust_sci(exact, 8, 2)}; Error = {(got / exact - 1.0) * 100.0:>6.1f}%")
        if not ok[-1]:
            print("**FAIL**")
    passed = all(ok)
    print(f"{name} validation passed." if passed else f"{name} validation failed.")
    return u, v, w, p, passed


def solve_channel_flow_velocity_inlet(iteration_count, reporting_interval, numerics, name, top_wall_velocity, inlet_velocity, mu, rho,
                                      examples_dir="./examples", mesh_path=None, ctx=None):
    """src/tests.rs:153-236 (what the reference's `main` runs today, src/main.rs:104-113). Returns (u, v, w, p)."""
    mesh = _io.read_mesh(mesh_path or os.path.join(examples_dir, "couette_flow_128x64x1.msh"))
    _set_common_zones(mesh, top_wall_velocity)
    mesh.get_face_zone("INLET").zone_type = FaceConditionTypes.VelocityInlet
    mesh.get_face_zone("INLET").vector_value = (inlet_velocity, 0.0, 0.0)
    data_path = os.path.join(examples_dir, f"{name}.csv")
    gradients_path = os.path.join(examples_dir, f"{name}_gradients.csv")
    try:
        u, v, w, p = _io.read_data(data_path)
    except OSError:
        u, v, w, p = _solver.initialize_flow_new(mesh, mu, rho, 1000, ctx=ctx)
    u, v, w, p = (np.ascontiguousarray(a, dtype=np.float64) for a in (u, v, w, p))
    _solver.solve_steady(mesh, u, v, w, p, numerics, rho, mu, iteration_count, max(reporting_interval, 1), ctx=ctx)
    _io.write_data(mesh, u, v, w, p, data_path)
    _io.write_gradients(mesh, u, v, w, p, gradients_path, 7, numerics.gradient_reconstruction, ctx)
    for label, value in ((" U_mean:\t", float(np.sum(u) / u.size)), (" U_min: \t", float(u.min())), (" U_max: \t", float(u.max()))):
        print(f"{label}CFD = {_rust_sci(value, 5, 2)}")
    return u, v, w, p


if __name__ == "__main__":
    # what the reference's binary does today (src/main.rs:49-122): `orc [iteration_count=10] [reporting_interval=0]`, run from a checkout
    # of the reference so that ./examples/couette_flow_128x64x1.msh is found
    import sys
    import time
    from .settings import NumericalSettings
    _start = time.time()
    _iterations = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    _interval = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    solve_channel_flow_velocity_inlet(_iterations, _interval, NumericalSettings(), "channel_flow_velocity_inlet", 0.0, 1e-3, 0.001, 1000.0)
    print(f"Complete in {int(time.time() - _start)}s.")
