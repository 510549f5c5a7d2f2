"""Partition-contiguous renumbering for arbitrary meshes. The multi-GPU path cuts the mesh into
contiguous cell-index ranges (rank r owns [cuts[r], cuts[r+1]), `Mesh.partition`); that is a good partition only if the
numbering is spatially coherent. For a mesh that is not (a mesher's numbering, or a shuffled one) the cells are renumbered
FIRST — recursive coordinate bisection on the cell centroids, part sizes equal to the even cuts of `Mesh.partition` — and
the renumbered connectivity is what is handed to both the product and the oracle (the reference's sequential recurrences
depend on the numbering, Q18: a renumbered mesh is simply another mesh).

Host-side numpy: the hot path never sees this module."""
import numpy as np


def even_cuts(n, nparts):
    """The cut points of Mesh.partition (mesh_host.hpp: even_cuts): coarse rows i/2 stay aligned with the owners."""
    return np.array([n if r == nparts else (n * r // nparts) & ~1 for r in range(nparts + 1)], dtype=np.int64)


def rcb_order(centroids, nparts):
    """new_to_old: a permutation of the cells such that part r = new ids [cuts[r], cuts[r+1]) is one box of a recursive
    coordinate bisection (longest extent of the current box, split at the count the cuts ask for). Inside a part the old
    relative order is kept."""
    c = np.asarray(centroids, dtype=np.float64)
    n = c.shape[0]
    cuts = even_cuts(n, nparts)
    sizes = np.diff(cuts)
    out = np.empty(n, dtype=np.int64)

    def split(ids, p0, p1, at):
        if p1 - p0 == 1:
            out[at:at + ids.size] = np.sort(ids)
            return
        pm = (p0 + p1) // 2
        nleft = int(sizes[p0:pm].sum())
        pts = c[ids]
        axis = int(np.argmax(pts.max(axis=0) - pts.min(axis=0))) if ids.size else 0
        # stable: ties along the axis are broken by the old index
        order = np.lexsort((ids, pts[:, axis]))
        split(ids[order[:nleft]], p0, pm, at)
        split(ids[order[nleft:]], pm, p1, at + nleft)

    split(np.arange(n, dtype=np.int64), 0, nparts, 0)
    return out


def renumber_cells(arrays, new_to_old):
    """The same mesh with cell `new_to_old[k]` renamed k. `arrays`: the dict of synthetic.hex_box / tet_box (TGRID-style:
    c0 / c1 are 1-based cell ids, 0 = no cell). Faces, nodes and zones are untouched: every cell keeps its faces in the same
    (ascending face index) order, so per-cell sums keep their order too."""
    new_to_old = np.asarray(new_to_old, dtype=np.int64)
    n = int(arrays["n_cells"])
    assert new_to_old.size == n and np.array_equal(np.sort(new_to_old), np.arange(n))
    old_to_new = np.empty(n, dtype=np.int64)
    old_to_new[new_to_old] = np.arange(n, dtype=np.int64)
    lut = np.concatenate([[0], old_to_new + 1])      # 1-based ids, 0 stays "no cell"
    out = dict(arrays)
    out["c0"] = lut[np.asarray(arrays["c0"], dtype=np.int64)]
    out["c1"] = lut[np.asarray(arrays["c1"], dtype=np.int64)]
    return out


def rcb_renumber(arrays, nparts, centroids=None):
    """(renumbered arrays, new_to_old). Centroids default to the product's own host geometry of the mesh (no GPU involved)."""
    if centroids is None:
        from .mesh import Mesh
        from . import synthetic as syn
        centroids = Mesh.from_arrays(*syn.mesh_args(arrays)).export()["cell_centroid"]
    new_to_old = rcb_order(centroids, nparts)
    return renumber_cells(arrays, new_to_old), new_to_old


def halo_cells(mesh, nparts):
    """Total number of halo cells over all ranks of Mesh.partition(rank, nparts): the volume of one halo exchange."""
    total = 0
    for r in range(nparts):
        info = mesh.partition(r, nparts).partition_info()
        total += info["n_lo"] + info["n_hi"]
    return total
