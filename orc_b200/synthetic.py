"""Synthetic box meshes for the measurement configs: TGRID-style connectivity arrays that are handed,
unchanged, to both the product (`Mesh.from_arrays`) and the oracle, plus a TGRID ASCII writer so that the reader
(`io.read_mesh`, src/io.rs:32-515 of the reference) can be exercised on meshes of any size.

Conventions (the ones every shipped 3-D example mesh follows):
  * node ids 0-based in the arrays, cell ids 1-based with 0 = "no cell" (exactly as in a `(13 ...)` section);
  * c0 is always present, boundary faces have c1 = 0, interior faces have c0 < c1;
  * unit((n2 - n1) x (n1 - n0)) points OUT of c0;
  * node coordinates are multiplied by (1 + jitter * N(0,1)) (seeded): on an exactly axis-aligned mesh b_v == b_w == 0 and
    the reference's unguarded BiCGSTAB divides 0/0 on the first iteration (SURVEY.md §7.4 hard part 2).
"""
import numpy as np

ZONE_NAMES = ["FLUID", "INLET", "OUTLET", "WALL", "SYM"]
ZONE_IDS = np.array([2, 3, 4, 5, 6], dtype=np.int64)
ZONE_TYPES = np.array([2, 3, 3, 3, 3], dtype=np.int64)  # interior + "wall" defaults, like the reference's files


def _jitter(xyz, jitter, seed):
    if jitter:
        rng = np.random.default_rng(seed)
        xyz = xyz * (1.0 + jitter * rng.standard_normal(xyz.shape))
    return np.ascontiguousarray(xyz)


def hex_box(nx, ny, nz, lx=0.004, ly=0.001, lz=0.001, jitter=1e-7, seed=0):
    """Structured hex channel, cells numbered x-fastest; interior faces sorted by (c0, direction +x,+y,+z), then the
    boundary zones INLET (x-), OUTLET (x+), WALL (y-, y+), SYM (z-, z+)."""
    n = (nx, ny, nz)
    xs = np.linspace(0.0, lx, nx + 1)
    ys = np.linspace(0.0, ly, ny + 1)
    zs = np.linspace(0.0, lz, nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    xyz = _jitter(np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1), jitter, seed)
    nstr = np.array([1, nx + 1, (nx + 1) * (ny + 1)], dtype=np.int64)   # node strides
    cstr = np.array([1, nx, nx * ny], dtype=np.int64)                    # cell strides

    def face_block(axis, pos, outward_sign):
        """All faces normal to `axis` at node-plane index `pos` (array over the two other axes).
        Returns node quads ordered so that the TGRID normal is +axis (sign=+1) or -axis (sign=-1), and the grid
        coordinates (i, j, k) of the plane's lower-corner nodes."""
        b, c = (axis + 1) % 3, (axis + 2) % 3
        rb, rc = np.arange(n[b], dtype=np.int64), np.arange(n[c], dtype=np.int64)
        RC, RB = np.meshgrid(rc, rb, indexing="ij")
        idx = [None, None, None]
        idx[axis] = np.broadcast_to(pos, RB.shape) if np.ndim(pos) else np.full(RB.shape, pos, dtype=np.int64)
        idx[b], idx[c] = RB, RC
        p = idx[0] * nstr[0] + idx[1] * nstr[1] + idx[2] * nstr[2]
        if outward_sign > 0:   # p, p+c, p+c+b, p+b : (n2-n1) x (n1-n0) = b x c = +axis
            quad = np.stack([p, p + nstr[c], p + nstr[c] + nstr[b], p + nstr[b]], axis=-1)
        else:                  # p, p+b, p+b+c, p+c : c x b = -axis
            quad = np.stack([p, p + nstr[b], p + nstr[b] + nstr[c], p + nstr[c]], axis=-1)
        return quad.reshape(-1, 4), [a.ravel() for a in idx]

    quads, c0s, c1s, dirs = [], [], [], []
    for axis in range(3):
        if n[axis] < 2:
            continue
        for pos in range(1, n[axis]):  # interior planes: c0 = cell below, c1 = cell above, normal +axis
            quad, idx = face_block(axis, pos, +1)
            cell_lo = [idx[0].copy(), idx[1].copy(), idx[2].copy()]
            cell_lo[axis] = cell_lo[axis] - 1
            c0 = cell_lo[0] * cstr[0] + cell_lo[1] * cstr[1] + cell_lo[2] * cstr[2]
            quads.append(quad); c0s.append(c0); c1s.append(c0 + cstr[axis]); dirs.append(np.full(c0.shape, axis, dtype=np.int64))
    if quads:
        quad_i = np.concatenate(quads); c0_i = np.concatenate(c0s); c1_i = np.concatenate(c1s); dir_i = np.concatenate(dirs)
        order = np.lexsort((dir_i, c0_i))
        quad_i, c0_i, c1_i = quad_i[order], c0_i[order], c1_i[order]
    else:
        quad_i = np.zeros((0, 4), np.int64); c0_i = np.zeros(0, np.int64); c1_i = np.zeros(0, np.int64)
    face_nodes = [quad_i]
    c0 = [c0_i + 1]
    c1 = [c1_i + 1]
    zone = [np.full(c0_i.shape, ZONE_IDS[0], dtype=np.int64)]
    for axis, side, zid in ((0, 0, 3), (0, 1, 4), (1, 0, 5), (1, 1, 5), (2, 0, 6), (2, 1, 6)):
        pos = 0 if side == 0 else n[axis]
        quad, idx = face_block(axis, pos, -1 if side == 0 else +1)
        cell = [idx[0].copy(), idx[1].copy(), idx[2].copy()]
        cell[axis] = np.zeros_like(cell[axis]) if side == 0 else np.full_like(cell[axis], n[axis] - 1)
        cb = cell[0] * cstr[0] + cell[1] * cstr[1] + cell[2] * cstr[2]
        face_nodes.append(quad); c0.append(cb + 1); c1.append(np.zeros_like(cb)); zone.append(np.full(cb.shape, zid, dtype=np.int64))
    face_nodes = np.concatenate(face_nodes)
    nf = face_nodes.shape[0]
    return dict(dims=3, xyz=xyz, face_node_offsets=np.arange(0, 4 * nf + 1, 4, dtype=np.int64), face_nodes=face_nodes.ravel(),
                c0=np.concatenate(c0), c1=np.concatenate(c1), face_zone=np.concatenate(zone), zone_ids=ZONE_IDS.copy(),
                zone_types=ZONE_TYPES.copy(), zone_names=list(ZONE_NAMES), n_cells=nx * ny * nz, shape=(nx, ny, nz), extent=(lx, ly, lz))


def hex_box_window(nx, ny, nz, z0, z1, lx=0.004, ly=0.001, lz=0.001, jitter=1e-7, seed=0):
    """The cells with z-index in [z0, z1) of hex_box(nx, ny, nz, ...), as a mesh of its own whose node coordinates are the
    GLOBAL jittered coordinates (same seeded stream), so geometry is bit-identical to the global mesh for every cell that is
    not on an artificial cut plane. Used by the multi-GPU bench: rank r builds only its slab + two layers per side.
    Returns (arrays, id_offset) with id_offset = global id of the window's cell 0."""
    assert 0 <= z0 < z1 <= nz
    m = hex_box(nx, ny, z1 - z0, lx, ly, lz * (z1 - z0) / nz, jitter=0.0)
    xs = np.linspace(0.0, lx, nx + 1); ys = np.linspace(0.0, ly, ny + 1); zs = np.linspace(0.0, lz, nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    xyz = _jitter(np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1), jitter, seed)
    plane = (nx + 1) * (ny + 1)
    m["xyz"] = np.ascontiguousarray(xyz[z0 * plane:(z1 + 1) * plane])
    return m, z0 * nx * ny


def slab_partition(nx, ny, nz, rank, nranks, layers=2, **box):
    """z-slab partition of hex_box(nx, ny, nz, **box): (window arrays, cuts in window numbering, id_offset, n_global)."""
    plane = nx * ny
    zc = [((nz * r // nranks) if r < nranks else nz) for r in range(nranks + 1)]
    z0, z1 = max(0, zc[rank] - layers), min(nz, zc[rank + 1] + layers)
    arrays, off = hex_box_window(nx, ny, nz, z0, z1, **box)
    cuts = [min(max((z - z0) * plane, 0), (z1 - z0) * plane) for z in zc]
    return arrays, cuts, off, nx * ny * nz


def tet_box(nx, ny, nz, lx=0.004, ly=0.001, lz=0.001, jitter=1e-7, seed=0):
    """The same jittered lattice with every hex split into 6 tetrahedra (Kuhn), cells numbered hex-major."""
    xs = np.linspace(0.0, lx, nx + 1); ys = np.linspace(0.0, ly, ny + 1); zs = np.linspace(0.0, lz, nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    xyz0 = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    xyz = _jitter(xyz0, jitter, seed)
    nstr = np.array([1, nx + 1, (nx + 1) * (ny + 1)], dtype=np.int64)
    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = (I * nstr[0] + J * nstr[1] + K * nstr[2]).ravel()
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    tets = np.empty((base.size, 6, 4), dtype=np.int64)
    for t, (a, b, c) in enumerate(perms):
        tets[:, t, 0] = base
        tets[:, t, 1] = base + nstr[a]
        tets[:, t, 2] = base + nstr[a] + nstr[b]
        tets[:, t, 3] = base + nstr[a] + nstr[b] + nstr[c]
    tets = tets.reshape(-1, 4)
    nt = tets.shape[0]
    # the 4 faces of every tet, keyed by their sorted node triple
    tri = np.stack([tets[:, [1, 2, 3]], tets[:, [0, 2, 3]], tets[:, [0, 1, 3]], tets[:, [0, 1, 2]]], axis=1).reshape(-1, 3)
    owner = np.repeat(np.arange(nt, dtype=np.int64), 4)
    key = np.sort(tri, axis=1)
    order = np.lexsort((owner, key[:, 2], key[:, 1], key[:, 0]))
    key, tri, owner = key[order], tri[order], owner[order]
    first = np.ones(key.shape[0], dtype=bool)
    first[1:] = np.any(key[1:] != key[:-1], axis=1)
    start = np.flatnonzero(first)
    count = np.diff(np.append(start, key.shape[0]))
    f_nodes = tri[start]
    f_c0 = owner[start]                                  # smaller cell id first (owner sorted within a key group)
    f_c1 = np.where(count == 2, owner[np.minimum(start + 1, owner.size - 1)], -1)
    # orient: unit((n2-n1) x (n1-n0)) must point out of c0
    P = xyz[f_nodes]
    nrm = np.cross(P[:, 2] - P[:, 1], P[:, 1] - P[:, 0])
    cc = xyz[tets[f_c0]].mean(axis=1)
    flip = np.einsum("ij,ij->i", nrm, P.mean(axis=1) - cc) < 0
    f_nodes[flip] = f_nodes[flip][:, ::-1]
    interior = f_c1 >= 0
    oi = np.flatnonzero(interior)
    oi = oi[np.lexsort((f_c1[oi], f_c0[oi]))]
    ob = np.flatnonzero(~interior)
    cen0 = xyz0[f_nodes[ob]].mean(axis=1)
    eps = 1e-9 * max(lx, ly, lz)
    zid = np.full(ob.size, -1, dtype=np.int64)
    zid[np.abs(cen0[:, 0]) < eps] = 3
    zid[np.abs(cen0[:, 0] - lx) < eps] = 4
    zid[(np.abs(cen0[:, 1]) < eps) | (np.abs(cen0[:, 1] - ly) < eps)] = 5
    zid[(np.abs(cen0[:, 2]) < eps) | (np.abs(cen0[:, 2] - lz) < eps)] = 6
    assert (zid > 0).all()
    ob = ob[np.lexsort((f_c0[ob], zid))]
    zid = np.sort(zid, kind="stable")
    sel = np.concatenate([oi, ob])
    fn = f_nodes[sel]
    nf = fn.shape[0]
    return dict(dims=3, xyz=xyz, face_node_offsets=np.arange(0, 3 * nf + 1, 3, dtype=np.int64), face_nodes=fn.ravel(),
                c0=f_c0[sel] + 1, c1=np.where(f_c1[sel] >= 0, f_c1[sel] + 1, 0),
                face_zone=np.concatenate([np.full(oi.size, ZONE_IDS[0], dtype=np.int64), zid]), zone_ids=ZONE_IDS.copy(),
                zone_types=ZONE_TYPES.copy(), zone_names=list(ZONE_NAMES), n_cells=nt, shape=(nx, ny, nz), extent=(lx, ly, lz))


def tet_box_window(nx, ny, nz, z0, z1, lx=0.004, ly=0.001, lz=0.001, jitter=1e-7, seed=0):
    """The tets of the hexes with z-index in [z0, z1) of tet_box(nx, ny, nz, ...), as a mesh of its own on the GLOBAL jittered
    node coordinates (same seeded stream). Numbering is hex-major, so the window's cells are a contiguous range of the global
    mesh and keep their relative order; so do the faces of every cell that does not touch an artificial cut plane (interior faces
    are sorted by (c0, c1), boundary faces by (zone, c0): both orders survive the shift). Faces on the cut planes land in the SYM
    zone of the window; they belong to the outermost hex layer only, which a partition never owns or reads (two layers per side).
    Returns (arrays, id_offset) with id_offset = global id of the window's cell 0."""
    assert 0 <= z0 < z1 <= nz
    m = tet_box(nx, ny, z1 - z0, lx, ly, lz * (z1 - z0) / nz, jitter=0.0)
    xs = np.linspace(0.0, lx, nx + 1); ys = np.linspace(0.0, ly, ny + 1); zs = np.linspace(0.0, lz, nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    xyz = _jitter(np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1), jitter, seed)
    plane = (nx + 1) * (ny + 1)
    m["xyz"] = np.ascontiguousarray(xyz[z0 * plane:(z1 + 1) * plane])
    return m, 6 * z0 * nx * ny


def tet_slab_partition(nx, ny, nz, rank, nranks, layers=2, **box):
    """z-slab partition of tet_box(nx, ny, nz, **box): (window arrays, cuts in window numbering, id_offset, n_global)."""
    plane = 6 * nx * ny
    zc = [((nz * r // nranks) if r < nranks else nz) for r in range(nranks + 1)]
    z0, z1 = max(0, zc[rank] - layers), min(nz, zc[rank + 1] + layers)
    arrays, off = tet_box_window(nx, ny, nz, z0, z1, **box)
    cuts = [min(max((z - z0) * plane, 0), (z1 - z0) * plane) for z in zc]
    return arrays, cuts, off, 6 * nx * ny * nz


def _from_face_list(base, faces, n_cells):
    """Mesh dict from [(node loop, c0, c1, zone)] (cells 1-based, 0 = none): interior faces get c0 < c1 (loop reversed when the
    cells swap, so that the TGRID normal keeps pointing out of c0) and are sorted by (c0, c1); boundary zones follow in id order."""
    fixed = []
    for loop, c0, c1, zone in faces:
        if c1 != 0 and c0 > c1:
            loop, c0, c1 = [loop[0]] + loop[:0:-1], c1, c0
        fixed.append((list(loop), c0, c1, zone))
    interior = sorted((f for f in fixed if f[2] != 0), key=lambda f: (f[1], f[2]))
    boundary = sorted((f for f in fixed if f[2] == 0), key=lambda f: f[3])            # stable: keeps the order within a zone
    allf = interior + boundary
    offsets = np.zeros(len(allf) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(f[0]) for f in allf])
    out = dict(base)
    out.update(face_node_offsets=offsets, face_nodes=np.array([v for f in allf for v in f[0]], dtype=np.int64),
               c0=np.array([f[1] for f in allf], dtype=np.int64), c1=np.array([f[2] for f in allf], dtype=np.int64),
               face_zone=np.array([f[3] for f in allf], dtype=np.int64), n_cells=n_cells)
    return out


def _face_list(m):
    fo, fn = m["face_node_offsets"], m["face_nodes"]
    return [([int(v) for v in fn[fo[q]:fo[q + 1]]], int(m["c0"][q]), int(m["c1"][q]), int(m["face_zone"][q])) for q in range(m["c0"].size)]


def wedge_box(nx, ny, nz, **box):
    """Every hex of hex_box cut along the diagonal (i, j) - (i+1, j+1) into two triangular prisms (TGRID cell type 6): hex c gives
    wedge 2c (the half holding node (i+1, j)) and 2c + 1 (the half holding (i, j+1)). Triangular and quadrilateral faces in one
    mesh: the z-normal faces are split, the others are handed to the wedge they bound, one new quad per hex separates the two."""
    m = hex_box(nx, ny, nz, **box)
    sx, sy = 1, nx + 1
    plane = (nx + 1) * (ny + 1)

    def ij(node):
        r = node % plane
        return r % (nx + 1), r // (nx + 1)

    def half(cell, loop):        # which wedge of hex `cell` (1-based) owns a face with these nodes: 0 -> 2c, 1 -> 2c + 1
        c = cell - 1
        i, j = c % nx, (c // nx) % ny
        pts = {ij(v) for v in loop}
        if (i + 1, j) in pts:
            return 0
        assert (i, j + 1) in pts
        return 1

    faces = []
    for loop, c0, c1, zone in _face_list(m):
        pts = [ij(v) for v in loop]
        if len(set(pts)) == 4:                                   # z-normal quad: two triangles along the hex diagonal
            k = next(q for q in range(4) if pts[q] == (min(p[0] for p in pts), min(p[1] for p in pts)))
            a, b, c, d = (loop[(k + q) % 4] for q in range(4))   # a = (i, j), c = (i+1, j+1)
            for tri in ([a, b, c], [a, c, d]):
                w0 = 2 * (c0 - 1) + half(c0, tri) + 1
                w1 = 2 * (c1 - 1) + half(c1, tri) + 1 if c1 else 0
                faces.append((tri, w0, w1, zone))
        else:
            w0 = 2 * (c0 - 1) + half(c0, loop) + 1
            w1 = 2 * (c1 - 1) + half(c1, loop) + 1 if c1 else 0
            faces.append((loop, w0, w1, zone))
    for c in range(nx * ny * nz):                                # the cut: normal (n2 - n1) x (n1 - n0) points from 2c to 2c + 1
        i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
        p = i * sx + j * sy + k * plane
        faces.append(([p, p + sx + sy, p + sx + sy + plane, p + plane], 2 * c + 1, 2 * c + 2, int(ZONE_IDS[0])))
    return _from_face_list(m, faces, 2 * nx * ny * nz)


def pyramid_box(nx, ny, nz, **box):
    """Every hex of hex_box cut into six pyramids (TGRID cell type 5) with a new apex node at the mean of its corners: pyramid
    6c + s stands on side s (x-, x+, y-, y+, z-, z+) of hex c; its base is the hex's quad face, its four triangles (hex edge +
    apex) are shared with the pyramids on the adjacent sides."""
    m = hex_box(nx, ny, nz, **box)
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    n_old = m["xyz"].shape[0]
    ncell = nx * ny * nz
    corner = lambda c, d: (c % nx + d[0]) * sx + ((c // nx) % ny + d[1]) * sy + (c // (nx * ny) + d[2]) * sz
    apex_xyz = np.array([np.mean([m["xyz"][corner(c, (a, b, g))] for a in (0, 1) for b in (0, 1) for g in (0, 1)], axis=0) for c in range(ncell)])
    xyz = np.vstack([m["xyz"], apex_xyz])

    def side(cell, loop):            # which side of hex `cell` (1-based) a quad with these nodes is
        c = cell - 1
        for axis in range(3):
            for hi in (0, 1):
                want = {corner(c, tuple(hi if q == axis else b[q] for q in range(3))) for b in ((0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1))}
                if set(loop) == want:
                    return 2 * axis + hi
        raise AssertionError("face does not bound the cell")

    faces = []
    for loop, c0, c1, zone in _face_list(m):
        p0 = 6 * (c0 - 1) + side(c0, loop) + 1
        p1 = 6 * (c1 - 1) + side(c1, loop) + 1 if c1 else 0
        faces.append((loop, p0, p1, zone))
    base_centre = lambda c, s_: np.mean([xyz[corner(c, tuple((s_ % 2) if q == s_ // 2 else b[q] for q in range(3)))]
                                         for b in ((0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1))], axis=0)
    for c in range(ncell):
        apex = n_old + c
        for a1 in range(3):
            for a2 in range(a1 + 1, 3):
                a3 = 3 - a1 - a2                                     # the edge runs along a3 where sides (a1, h1) and (a2, h2) meet
                for h1 in (0, 1):
                    for h2 in (0, 1):
                        d0, d1 = [0, 0, 0], [0, 0, 0]
                        d0[a1] = d1[a1] = h1
                        d0[a2] = d1[a2] = h2
                        d1[a3] = 1
                        s0, s1 = 2 * a1 + h1, 2 * a2 + h2           # s0 < s1: pyramid 6c + s0 is c0
                        tri = [corner(c, tuple(d0)), corner(c, tuple(d1)), apex]
                        nrm = np.cross(xyz[tri[2]] - xyz[tri[1]], xyz[tri[1]] - xyz[tri[0]])
                        if np.dot(nrm, base_centre(c, s1) - base_centre(c, s0)) < 0:
                            tri = [tri[1], tri[0], tri[2]]
                        faces.append((tri, 6 * c + s0 + 1, 6 * c + s1 + 1, int(ZONE_IDS[0])))
    out = _from_face_list(m, faces, 6 * ncell)
    out["xyz"] = np.ascontiguousarray(xyz)
    return out


def poly_box(nx, ny, nz, **box):
    """Pairs of hexes (2m, 2m + 1 along x; nx even) merged into one polyhedral cell (TGRID cell type 7): the quad between them goes,
    the two coplanar quads they show to a neighbour (or to a boundary zone) become ONE six-node polygon, so every pair of cells
    still shares exactly one face. The loop of a polygon starts so that its first three nodes span a corner (the reference takes
    the face normal from the first three nodes, src/io.rs:322-326)."""
    assert nx % 2 == 0
    m = hex_box(nx, ny, nz, **box)
    xyz = m["xyz"]
    merged = lambda cell: ((cell - 1) % nx) // 2 + (nx // 2) * ((cell - 1) // nx) + 1 if cell else 0
    groups = {}
    for loop, c0, c1, zone in _face_list(m):
        a, b = merged(c0), merged(c1)
        if a == b:
            continue
        groups.setdefault((a, b, zone), []).append(loop)

    def join(l1, l2):
        e1 = {(l1[q], l1[(q + 1) % len(l1)]) for q in range(len(l1))}
        e2 = {(l2[q], l2[(q + 1) % len(l2)]) for q in range(len(l2))}
        shared = {(u, v) for (u, v) in e1 if (v, u) in e2}
        assert len(shared) == 1
        nxt = {u: v for (u, v) in (e1 | e2) if (u, v) not in shared and (v, u) not in shared}
        start = next(iter(nxt))
        loop, v = [start], nxt[start]
        while v != start:
            loop.append(v)
            v = nxt[v]
        corner = lambda q: np.linalg.norm(np.cross(xyz[loop[(q + 2) % len(loop)]] - xyz[loop[(q + 1) % len(loop)]],
                                                   xyz[loop[(q + 1) % len(loop)]] - xyz[loop[q]]))
        q = max(range(len(loop)), key=corner)
        return loop[q:] + loop[:q]

    faces = []
    for (a, b, zone), loops in groups.items():
        assert len(loops) in (1, 2)
        faces.append((loops[0] if len(loops) == 1 else join(*loops), a, b, zone))
    return _from_face_list(m, faces, nx * ny * nz // 2)


def channel_bcs(mesh, inlet_pressure=-0.01, fully_3d=False):
    """BCs of the synthetic channel, set by zone NAME like src/tests.rs:60-76:
    INLET PressureInlet, OUTLET PressureOutlet 0, WALL Wall (no slip), SYM Symmetry (or Wall for a fully 3-D flow).
    `mesh` is anything with set_zone(name, type, scalar, vector): the product Mesh and the oracle Mesh both qualify."""
    mesh.set_zone("INLET", 4, inlet_pressure, (0.0, 0.0, 0.0))
    mesh.set_zone("OUTLET", 5, 0.0, (0.0, 0.0, 0.0))
    mesh.set_zone("WALL", 3, 0.0, (0.0, 0.0, 0.0))
    mesh.set_zone("SYM", 3 if fully_3d else 7, 0.0, (0.0, 0.0, 0.0))


def mesh_args(m):
    """Positional arguments of Mesh.from_arrays (product and oracle share the signature)."""
    return (m["dims"], m["xyz"], m["face_node_offsets"], m["face_nodes"], m["c0"], m["c1"], m["face_zone"], m["zone_ids"],
            m["zone_types"], m["zone_names"])


def write_tgrid(path, m):
    """Write the arrays as a TGRID ASCII file in the subset the reference reader accepts (uniform tri/quad sections, or mixed
    sections of face type 0 whose lines carry no node count — the reference takes it from the line length —, hex integers, one
    `(0 "... NAME")` comment before each face section)."""
    xyz, fo, fn = m["xyz"], m["face_node_offsets"], m["face_nodes"]
    c0, c1, fz = m["c0"], m["c1"], m["face_zone"]
    nn, nf, nc = xyz.shape[0], c0.size, int(m["n_cells"])
    with open(path, "w") as f:
        f.write('(0 "Created by: orc_b200.synthetic")\n(2 3)\n(0 "Node Section")\n')
        f.write(f"(10 (0 1 {nn:x} 0 3))\n(10 (1 1 {nn:x} 1 3)\n(\n")
        for x, y, z in xyz:
            f.write(f"{float(x)!r} {float(y)!r} {float(z)!r}\n")
        f.write("))\n")
        f.write(f"(12 (0 1 {nc:x} 0 0))\n(12 (7 1 {nc:x} 1 4))\n(13 (0 1 {nf:x} 0 0))\n")
        start = 0
        for zid, ztype, name in zip(m["zone_ids"], m["zone_types"], m["zone_names"]):
            idx = np.flatnonzero(fz == zid)
            if idx.size == 0:
                continue
            assert idx[0] == start and idx[-1] == start + idx.size - 1, "faces of a zone must be contiguous"
            counts = np.diff(fo)[idx]
            k = int(counts[0]) if np.all(counts == counts[0]) and counts[0] <= 4 else 0     # 0: mixed section, the node count is the line length (io.rs:231-235)
            f.write(f'(0 "Faces of zone {name}")\n(13 ({int(zid):x} {start + 1:x} {start + idx.size:x} {int(ztype):x} {k:x})(\n')
            for q in idx:
                nodes = " ".join(f"{int(v) + 1:x}" for v in fn[fo[q]:fo[q + 1]])
                f.write(f"{nodes} {int(c0[q]):x} {int(c1[q]):x}\n")
            f.write(")\n)\n")
            start += idx.size
