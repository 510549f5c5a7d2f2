"""Generates the golden fixtures under tests/golden/. Run in the build container (needs /root/reference for the example
meshes; the GPU box has no reference tree, so nothing under tests/ reads it at run time):

    python tests/golden/make_golden.py

1. mesh_<name>.npz  — connectivity of the reference's example meshes (examples/*.msh) as TGRID-style arrays
   (node coordinates as parsed from the file, face nodes, c0/c1 1-based with 0 = none, zone ids/types/names), read with
   the oracle's restatement of io.rs::read_mesh. Geometry is NOT stored: both implementations recompute it.
2. kat_<name>.npz   — outputs of the ORACLE (oracle/, the CPU restatement of the reference path) on those meshes:
   assembled coefficients after one assembly, and fields after a few SIMPLE iterations at the reference's default
   settings. They pin the oracle against accidental change and give the GPU tests a second, file-based comparison.
   They are only as good as the oracle's fidelity ("parity unpinned" below 1e-3, see oracle/orc_oracle.hpp).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po  # noqa: E402
from cases import couette_bcs  # noqa: E402

REF = "/root/reference/examples"
MESHES = ["2D_3x6", "3D_1x3", "3x3_cube", "couette_flow_8x8x1", "channel_flow", "couette_flow_128x64x1"]


def tgrid_arrays(path):
    """Parse the (10 / (13 sections with the same rules as the reader (hex ids, zone name = last word of the last comment)."""
    xyz, faces, zones = [], [], {}
    zone_name, dims = "", 3
    with open(path) as f:
        lines = [l.rstrip("\r\n") for l in f]
    i = 0
    import re
    while i < len(lines):
        line = lines[i]
        tok = line.split()
        if not tok:
            i += 1
            continue
        if tok[0] == "(0":
            zone_name = line.rsplit(" ", 1)[1]
            while zone_name.endswith('")'):
                zone_name = zone_name[:-2]
        elif tok[0] == "(2":
            dims = int(tok[1].rstrip(")"))
        elif tok[0] == "(10" and tok[1] != "(0":
            i += 1
            while not lines[i].startswith(")"):
                if lines[i] != "(":
                    v = [float(x) for x in lines[i].split()]
                    if len(v) == dims:
                        xyz.append(v + [0.0] * (3 - dims))
                i += 1
        elif tok[0] == "(13" and tok[1] != "(0":
            items = [int(x, 16) for x in re.findall(r"([0-9a-z]+)", line)]
            zid, bc = items[1], items[4]
            zones.setdefault(zid, (bc, zone_name))
            i += 1
            while not lines[i].startswith(")"):
                if lines[i] != "(":
                    v = [int(x, 16) for x in lines[i].split()]
                    faces.append((v[:-2], v[-2], v[-1], zid))
                i += 1
        i += 1
    offs = np.cumsum([0] + [len(f[0]) for f in faces]).astype(np.int64)
    nodes = np.array([n - 1 for f in faces for n in f[0]], dtype=np.int64)
    zid = sorted(zones)
    return dict(dims=np.int64(dims), xyz=np.array(xyz), face_node_offsets=offs, face_nodes=nodes,
                c0=np.array([f[1] for f in faces], np.int64), c1=np.array([f[2] for f in faces], np.int64),
                face_zone=np.array([f[3] for f in faces], np.int64), zone_ids=np.array(zid, np.int64),
                zone_types=np.array([zones[z][0] for z in zid], np.int64), zone_names=np.array([zones[z][1] for z in zid]))


def main():
    for name in MESHES:
        a = tgrid_arrays(os.path.join(REF, name + ".msh"))
        # the arrays must reproduce the oracle's reader bit for bit
        m_file = po.Mesh.read(os.path.join(REF, name + ".msh"))
        m_arr = po.Mesh.from_arrays(int(a["dims"]), a["xyz"], a["face_node_offsets"], a["face_nodes"], a["c0"], a["c1"], a["face_zone"],
                                    a["zone_ids"], a["zone_types"], [str(s) for s in a["zone_names"]])
        ea, eb = m_file.export(), m_arr.export()
        for k in ea:
            assert np.array_equal(ea[k], eb[k]), (name, k)
        np.savez_compressed(os.path.join(HERE, f"mesh_{name}.npz"), **a)
        print("wrote mesh", name, m_file.counts())

    # known-answer outputs of the oracle on the two configs of BASELINE.json that ship as meshes
    for name, walls, moving, dp_dx, u_wall, kw, iters in (
            ("channel_flow", ("WALL",), None, 5.0, 0.0, dict(momentum=po.TVD, limiter=po.PSI_QUICK), 3),
            ("couette_flow_128x64x1", ("TOP_WALL", "BOTTOM_WALL"), "TOP_WALL", 10.0, 5e-4, dict(), 2)):
        a = np.load(os.path.join(HERE, f"mesh_{name}.npz"))
        m = po.Mesh.from_arrays(int(a["dims"]), a["xyz"], a["face_node_offsets"], a["face_nodes"], a["c0"], a["c1"], a["face_zone"],
                                a["zone_ids"], a["zone_types"], [str(s) for s in a["zone_names"]])
        couette_bcs(m, u_wall=u_wall, dp_dx=dp_dx, wall_zones=walls, moving=moving)
        n = m.n_cells
        s = po.Settings(**kw)
        z = np.zeros(n)
        u, v, w, p, rep, _ = m.solve_steady(z, z, z, z, s, 1000.0, 1e-3, iters, 1)
        # one assembly from those fields (per-call parity input/outputs)
        a_di, bud, bvd, bwd = m.build_momentum_diffusion(1e-3)
        mats = [m.init_momentum_matrix() for _ in range(3)]
        bu, bv, bw, pe = m.build_momentum_advection(*mats, a_di, u, v, w, p, s, 1000.0)
        pa, pb = m.build_pressure_correction(*mats, u, v, w, p, s, 1000.0)
        np.savez_compressed(os.path.join(HERE, f"kat_{name}.npz"), iters=np.int64(iters), u=u, v=v, w=w, p=p, reports=rep,
                            a_di=a_di.arrays()[2], a_u=mats[0].arrays()[2], a_v=mats[1].arrays()[2], a_w=mats[2].arrays()[2],
                            b_u=bu, b_v=bv, b_w=bw, peclet=np.array(pe), pc_a=pa.arrays()[2], pc_b=pb,
                            rowptr=a_di.arrays()[0], col=a_di.arrays()[1],
                            settings=np.array([s.momentum, s.limiter, s.pressure_interpolation, s.velocity_interpolation], np.int64),
                            bc=np.array([dp_dx, u_wall]))
        print("wrote kat", name, "u_avg per iteration", rep[:, 1])


if __name__ == "__main__":
    main()
