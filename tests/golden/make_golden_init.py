"""Golden fixtures of the flow initialisation (src/solver.rs:246-352, 414-509), produced by the ORACLE on the committed
connectivity of the reference's example meshes (tests/golden/mesh_*.npz — no reference tree needed):

    python tests/golden/make_golden_init.py

kat_init_<name>.npz: the Laplace system of initialize_pressure_field (values, right-hand side) and u, v, w, p of
initialize_flow(mesh, mu, rho, ITERS) with the boundary conditions of src/tests.rs:60-76. Like kat_<name>.npz they pin the oracle
against accidental change and give the GPU tests a file-based comparison; they are only as good as the oracle's fidelity."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po  # noqa: E402
from orc_b200 import synthetic as syn  # noqa: E402
from cases import couette_bcs, load_mesh_arrays  # noqa: E402

ITERS = 30
CASES = [("channel_flow", ("WALL",), None, 5.0, 0.0), ("couette_flow_128x64x1", ("TOP_WALL", "BOTTOM_WALL"), "TOP_WALL", 10.0, 5e-4)]


def main():
    for name, walls, moving, dp_dx, u_wall in CASES:
        m = po.Mesh.from_arrays(*syn.mesh_args(load_mesh_arrays(name)))
        couette_bcs(m, u_wall=u_wall, dp_dx=dp_dx, wall_zones=walls, moving=moving)
        a, b = m.build_pressure_laplace()
        _, _, va = a.arrays()
        u, v, w, p = m.initialize_flow(1e-3, 1000.0, ITERS)
        out = os.path.join(HERE, f"kat_init_{name}.npz")
        np.savez_compressed(out, laplace_val=va, laplace_b=b, u=u, v=v, w=w, p=p, iters=np.int64(ITERS),
                            constraint_type=np.int64(m.check_boundary_conditions()))
        print(name, m.n_cells, "cells ->", out, os.path.getsize(out) // 1024, "KiB", "finite:", bool(np.isfinite(u).all()))


if __name__ == "__main__":
    main()
