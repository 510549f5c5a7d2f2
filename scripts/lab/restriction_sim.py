"""CPU discrete-event model of k_strongest_dataflow's schedule (tickets in index order, 8 warps x DFR_ROWS rows per block, a row
waits for the lower touchers of its candidate) on a real matrix: where does the time go on the Kuhn-split tet slabs?
Usage: python scripts/lab/restriction_sim.py nx ny nz [hex|tet]"""
import heapq, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as oracle
from orc_b200 import synthetic as syn
from cases import smooth_fields

nx, ny, nz = (int(a) for a in sys.argv[1:4])
kind = sys.argv[4] if len(sys.argv) > 4 else "tet"
oracle.build()
arr = syn.hex_box(nx, ny, nz) if kind == "hex" else syn.tet_box(nx, ny, nz)
om = oracle.Mesh.from_arrays(*syn.mesh_args(arr))
syn.channel_bcs(om, fully_3d=(kind == "tet"))
u, v, w, p = smooth_fields(om.export())
o_di, *_ = om.build_momentum_diffusion(1e-3)
o_a = [om.init_momentum_matrix() for _ in range(3)]
st = oracle.Settings(momentum=3, limiter=4) if kind == "tet" else oracle.Settings()
om.build_momentum_advection(*o_a, o_di, u, v, w, p, st, 1000.0)
A = o_a[0]
b = np.ones(A.dims[0])
As, _ = A.jacobi_scale(b)
rp, col, val = As.arrays()
n = As.dims[0]
print(f"{kind} {nx}x{ny}x{nz}: {n} rows, {col.size / n:.1f} entries per row", flush=True)

# ---- the sequential greedy + the waits the kernel performs ----
WARPS = int(os.environ.get("SIM_WARPS", "8"))
ROWS = int(os.environ.get("SIM_ROWS", "4"))        # rows a warp handles one after the other per ticket
BLOCKS = 148 * (64 // WARPS)                       # 64 resident warps per SM
T_TICKET = float(os.environ.get("SIM_TICKET", "1.0"))   # us: block barrier + atomic ticket round trip
T_ROW, T_CAND, T_VIS, T_DEC = 1.0, 0.5, 0.7, 0.3     # us: row scan (3 dependent loads), candidate row, poll latency after a decision, publish
combined = np.zeros(n, bool)
pick = np.full(n, -1, np.int64)
tdec = np.zeros(n)
waits_on = np.full(n, -1, np.int64)      # the row whose decision released row i (critical-path link)
blocks = [(0.0, b_, -1) for b_ in range(BLOCKS)]
seq_link = np.full(n, -1, np.int64)      # the row that set the warp clock row i started from
heapq.heapify(blocks)
BCH = WARPS * ROWS
rpl, coll, vall = rp.tolist(), col.tolist(), val.tolist()
t0 = time.time()
nretry = 0
for c0 in range(0, n, BCH):
    tfree, bid, lastrow = heapq.heappop(blocks)
    wclock = [tfree + T_TICKET] * WARPS
    wprev = [lastrow] * WARPS      # the row that set this warp's clock
    for q in range(ROWS):
        for wv in range(WARPS):
            i = c0 + q * WARPS + wv
            if i >= n:
                continue
            t = wclock[wv] + T_ROW
            lo, hi = rpl[i], rpl[i + 1]
            cand = sorted((vall[k], k) for k in range(lo, hi) if coll[k] != i)
            chosen = -1
            link = -1
            for vv, k in cand:
                j = coll[k]
                # the kernel skips j when it already SEES combined[j]; a hint that is not visible yet costs a retry
                t += T_CAND
                taken = False
                tw = t
                lnk = -1
                for kk in range(rpl[j], rpl[j + 1]):
                    r = coll[kk]
                    if r < i and r != j:
                        if pick[r] == j:
                            # early exit: as soon as the taker is decided
                            if tdec[r] + T_VIS > t: tw_t = tdec[r] + T_VIS
                            else: tw_t = t
                            taken = True; tw = tw_t; lnk = r
                            break
                if taken:
                    if tw > t: link = lnk
                    t = max(t, tw); nretry += 1
                    continue
                for kk in range(rpl[j], rpl[j + 1]):
                    r = coll[kk]
                    if r < i and r != j and tdec[r] + T_VIS > tw:
                        tw = tdec[r] + T_VIS; lnk = r
                if tw > t: link = lnk
                t = max(t, tw)
                chosen = j
                break
            if chosen >= 0:
                combined[chosen] = True
            pick[i] = chosen
            t += T_DEC
            tdec[i] = t
            waits_on[i] = link
            seq_link[i] = wprev[wv]
            wclock[wv] = t
            wprev[wv] = i
    wl = int(np.argmax(wclock))
    heapq.heappush(blocks, (wclock[wl], bid, wprev[wl]))
total = tdec.max()
print(f"warps/block {WARPS}, rows per warp per ticket {ROWS}, ticket cost {T_TICKET} us -> simulated: {total / 1e3:.2f} ms = {total * 1e3 / n:.1f} ns per row ({time.time() - t0:.0f} s of CPU), retries {nretry}")
# critical path: a row's time was set either by the decision it waited for (dep) or by its warp's clock (seq: the previous row
# of the warp, or the row that freed the block)
i = int(np.argmax(tdec)); ndep = nseq = 0; dist_dep = []; dist_seq = []
while i >= 0:
    if waits_on[i] >= 0:
        dist_dep.append(i - int(waits_on[i])); i = int(waits_on[i]); ndep += 1
    else:
        j = int(seq_link[i])
        if j >= 0: dist_seq.append(i - j)
        i = j; nseq += 1
print(f"critical chain: {ndep} dependency links + {nseq} warp/block-sequence links")
for name, d in (("dependency", dist_dep), ("sequence", dist_seq)):
    if d:
        vals, cnts = np.unique(np.array(d), return_counts=True)
        order = np.argsort(-cnts)[:8]
        print(f"  {name} link distances (rows back: count):", ", ".join(f"{int(vals[o])}: {int(cnts[o])}" for o in order))
if kind == "tet":
    L = 6 * nx
    print("line: t(first row decided) .. t(last row decided) [us]")
    for ln in list(range(0, 6)) + list(range(40, 50)) + list(range(200, 204)):
        if (ln + 1) * L <= n:
            seg = tdec[ln * L:(ln + 1) * L]
            print(f"  line {ln}: {seg[0]:.1f} .. {seg[-1]:.1f}   (quarter points: {seg[L // 4]:.1f} {seg[L // 2]:.1f} {seg[3 * L // 4]:.1f})")
