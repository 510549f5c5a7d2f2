"""CPU analysis (no GPU): distinct 128-byte lines per warp instruction of the coarse-level gather SpMV under the production lane
mapping (G lanes per row, 32/G rows per warp), for the natural numbering of the aggregates and for locality orderings of them.
The hierarchy is the oracle's (bit-identical to the GPU's). Usage: python scripts/lab/reorder_analysis.py [n]"""
import os, sys, time
import numpy as np
import scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as oracle
from orc_b200 import synthetic as syn
from cases import smooth_fields

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
kind = sys.argv[2] if len(sys.argv) > 2 else "hex"
oracle.build()
arr = syn.hex_box(n, n, n) if kind == "hex" else syn.tet_box(n, n, n)
om = oracle.Mesh.from_arrays(*syn.mesh_args(arr))
syn.channel_bcs(om)
ex = om.export()
u, v, w, p = smooth_fields(ex)
o_di, *_ = om.build_momentum_diffusion(1e-3)
o_a = [om.init_momentum_matrix() for _ in range(3)]
om.build_momentum_advection(*o_a, o_di, u, v, w, p, oracle.Settings(), 1000.0)
N = om.n_cells
t = time.time()
x, levels = oracle.multigrid_trace(o_a[0], np.ones(N), np.zeros(N), iterations=1)
print("hierarchy", time.time() - t, "s")
cc = ex["cell_centroid"]


def lines_per_entry(A, G, cpl):
    A = A.tocsr(); A.sort_indices()
    rp, col = A.indptr, A.indices
    nr = A.shape[0]
    row = np.repeat(np.arange(nr), np.diff(rp))
    pos = np.arange(col.size) - rp[row]
    instr = (row // (32 // G)).astype(np.int64) * 4096 + pos // G
    key = instr * (1 << 22) + (col // cpl)
    return np.unique(key).size / col.size


def morton(c, bits=10):
    q = np.zeros((c.shape[0], 3), np.int64)
    for a in range(3):   # rank-quantise per axis
        r = np.argsort(np.argsort(c[:, a], kind="stable"), kind="stable")
        q[:, a] = r * (1 << bits) // c.shape[0]
    key = np.zeros(c.shape[0], np.int64)
    for b in range(bits):
        for a in range(3):
            key |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return key


def bricks(c, shape):
    """rank-quantised grid, brick-major order with bricks of `shape` cells (x fastest inside)."""
    m = c.shape[0]
    side = round(m ** (1 / 3))
    q = [np.minimum((np.argsort(np.argsort(c[:, a], kind="stable"), kind="stable") * side // m), side - 1) for a in range(3)]
    bx, by, bz = shape
    key = ((q[2] // bz) * 4096 + (q[1] // by)) * 4096 + (q[0] // bx)
    key = key * 64 + ((q[2] % bz) * by + (q[1] % by)) * bx + (q[0] % bx)
    return key


cent = cc
for l, (R, A) in enumerate(levels):
    Rs, As = R.to_scipy().tocsr(), A.to_scipy().tocsr()
    cnt = np.asarray(Rs.sum(axis=1)).ravel()   # R entries are 1: members per aggregate
    cent = (Rs @ cent) / np.maximum(np.asarray((Rs != 0).sum(axis=1)).ravel(), 1)[:, None]
    m = As.shape[0]
    epr = As.nnz / m
    G = 4 if epr < 24 else 8
    print(f"level {l + 1}: {m} rows, {epr:.1f} entries/row, G={G}")
    for name, keyf in [("natural", None), ("morton", lambda c: morton(c)), ("brick 1x2x2", lambda c: bricks(c, (1, 2, 2))),
                       ("brick 2x2x1", lambda c: bricks(c, (2, 2, 1))), ("brick 4x2x2", lambda c: bricks(c, (4, 2, 2))), ("brick 2x2x2", lambda c: bricks(c, (2, 2, 2))),
                       ("brick 4x4x4", lambda c: bricks(c, (4, 4, 4)))]:
        if keyf is None:
            B = As
        else:
            perm = np.argsort(keyf(cent), kind="stable")
            B = As[perm][:, perm]
        print(f"   {name:12s}: K=1 {lines_per_entry(B, G, 16):.3f}  K=3 {lines_per_entry(B, G, 4):.3f} lines/entry;  G=8: K=3 {lines_per_entry(B, 8, 4):.3f}  G=16: {lines_per_entry(B, 16, 4):.3f} G=32: {lines_per_entry(B, 32, 4):.3f}")


# ---- the cheap variant a device implementation would use: position of coarse row I = position of fine row 2 I, Morton code of
# the position quantised on the bounding box (bits per axis chosen from the number of distinct rows per axis ~ cube root).
def morton_bbox(c, bits):
    lo, hi = c.min(axis=0), c.max(axis=0)
    q = np.minimum(((c - lo) / (hi - lo + 1e-300) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    key = np.zeros(c.shape[0], np.int64)
    for b in range(bits):
        for a in range(3):
            key |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return key


print("---- cheap variant: position of row I = position of fine row 2I; bounding-box Morton")
cent = cc
for l, (R, A) in enumerate(levels):
    As = A.to_scipy().tocsr()
    m = As.shape[0]
    cent = cent[np.minimum(2 * np.arange(m), cent.shape[0] - 1)]
    G = 4 if As.nnz / m < 24 else 8
    for bits in (5, 7, 10):
        perm = np.argsort(morton_bbox(cent, bits), kind="stable")
        B = As[perm][:, perm]
        print(f"level {l + 1} bits {bits}: K=1 {lines_per_entry(B, G, 16):.3f}  K=3 {lines_per_entry(B, G, 4):.3f}")
