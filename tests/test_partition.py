"""Host-side logic of the multi-GPU path on CPU: partition extraction and the halo-exchange plan, checked in one process and
with a world_size-2 gloo group (torch.distributed send/recv standing in for ncclSend/ncclRecv)."""
import os
import socket

import numpy as np
import pytest

import orc_b200
from orc_b200 import synthetic as syn


def global_mesh(kind="hex"):
    arrays = {"hex": lambda: syn.hex_box(6, 5, 8), "tet": lambda: syn.tet_box(3, 3, 4), "wedge": lambda: syn.wedge_box(4, 3, 6),
              "polyhedra": lambda: syn.poly_box(6, 3, 6), "pyramid": lambda: syn.pyramid_box(3, 2, 4)}[kind]()
    m = orc_b200.Mesh.from_arrays(*syn.mesh_args(arrays))
    syn.channel_bcs(m)
    return m


@pytest.mark.parametrize("kind", ["hex", "tet", "wedge", "polyhedra", "pyramid"])
@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
def test_partition_geometry_pattern_and_plan(kind, nranks):
    g = global_mesh(kind)
    ge = g.export()
    grp, gco = g.pattern()
    n = g.n_cells
    owned_total = 0
    parts = []
    for r in range(nranks):
        p = g.partition(r, nranks)
        info, maps, e = p.partition_info(), p.partition_maps(), p.export()
        parts.append((info, maps))
        l2g = maps["local_to_global"]
        lo, hi = info["n_lo"], info["n_lo"] + info["n_own"]
        assert info["g0"] % 2 == 0 and info["n_global"] == n
        assert np.all(np.diff(l2g) > 0)                                   # ascending global id: column order is preserved
        assert np.array_equal(l2g[lo:hi], np.arange(info["g0"], info["g1"]))
        owned_total += info["n_own"]
        # geometry of every local cell (halo included) is the global geometry, bit for bit
        assert np.array_equal(e["cell_volume"], ge["cell_volume"][l2g]) and np.array_equal(e["cell_centroid"], ge["cell_centroid"][l2g])
        # owned cells keep all their faces in the same order; halo cells have none
        nf = np.diff(e["cell_face_offsets"])
        assert np.array_equal(nf[lo:hi], np.diff(ge["cell_face_offsets"])[info["g0"]:info["g1"]]) and nf[:lo].sum() == 0 and nf[hi:].sum() == 0
        for c in (lo, hi - 1):
            gf = ge["cell_face_indices"][ge["cell_face_offsets"][l2g[c]]:ge["cell_face_offsets"][l2g[c] + 1]]
            lf = e["cell_face_indices"][e["cell_face_offsets"][c]:e["cell_face_offsets"][c + 1]]
            assert np.array_equal(e["face_area"][lf], ge["face_area"][gf]) and np.array_equal(e["face_normal"][lf], ge["face_normal"][gf])
        # pattern of the owned rows = global rows with columns mapped to local ids
        rp, co = p.pattern()
        for c in range(lo, hi):
            assert np.array_equal(l2g[co[rp[c]:rp[c + 1]]], gco[grp[l2g[c]]:grp[l2g[c] + 1]])
        assert rp[lo] == 0 and rp[-1] == rp[hi]
        # the level schedule only follows owned lower neighbours
        lv = p.levels()
        for c in range(lo, hi):
            lower = [j for j in co[rp[c]:rp[c + 1]] if lo <= j < c]
            assert lv[c] == (1 + max(lv[j] for j in lower) if lower else 0)
    assert owned_total == n
    # the exchange plan is consistent between ranks: what r sends to q is, in order, what q expects from r
    for r, (info, maps) in enumerate(parts):
        for k, q in enumerate(maps["nbr_rank"]):
            send_global = maps["local_to_global"][maps["send_idx"][maps["send_ptr"][k]:maps["send_ptr"][k + 1]]]
            qi, qm = parts[q]
            kk = list(qm["nbr_rank"]).index(r)
            recv_global = qm["local_to_global"][qm["recv_begin"][kk]:qm["recv_begin"][kk] + qm["recv_count"][kk]]
            assert np.array_equal(send_global, recv_global)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = global_mesh("hex")
        n = g.n_cells
        p = g.partition(rank, world)
        info, maps = p.partition_info(), p.partition_maps()
        l2g = maps["local_to_global"]
        lo, hi = info["n_lo"], info["n_lo"] + info["n_own"]
        # a global matrix with the mesh pattern and a global vector, identical on every rank
        grp, gco = g.pattern()
        rng = np.random.default_rng(0)
        gval, x_global = rng.standard_normal(gco.size), rng.standard_normal(n)
        # local x: owned values only; halo filled by the exchange (send = pack owned cells, recv = contiguous halo slice)
        x = np.zeros(l2g.size)
        x[lo:hi] = x_global[info["g0"]:info["g1"]]
        reqs, keep = [], []
        for k, qrank in enumerate(maps["nbr_rank"]):
            snd = torch.from_numpy(x[maps["send_idx"][maps["send_ptr"][k]:maps["send_ptr"][k + 1]]].copy())
            rcv = torch.zeros(int(maps["recv_count"][k]), dtype=torch.float64)
            keep.append((k, rcv))
            reqs.append(dist.isend(snd, int(qrank)))
            reqs.append(dist.irecv(rcv, int(qrank)))
        for rq in reqs:
            rq.wait()
        for k, rcv in keep:
            x[maps["recv_begin"][k]:maps["recv_begin"][k] + maps["recv_count"][k]] = rcv.numpy()
        assert np.array_equal(x, x_global[l2g])
        # local SpMV over the owned rows in local column order == the global rows, bit for bit (ascending-k order kept)
        rp, co = p.pattern()
        y = np.zeros(info["n_own"])
        for c in range(lo, hi):
            gi = l2g[c]
            vals = gval[grp[gi]:grp[gi + 1]]
            acc = 0.0
            for a, j in zip(vals, co[rp[c]:rp[c + 1]]):
                acc += a * x[j]
            y[c - lo] = acc
        y_ref = np.zeros(n)
        for i in range(n):
            acc = 0.0
            for k in range(grp[i], grp[i + 1]):
                acc += gval[k] * x_global[gco[k]]
            y_ref[i] = acc
        assert np.array_equal(y, y_ref[info["g0"]:info["g1"]])
        # a dot product = allreduce of the owned partial sums
        t = torch.tensor([float(np.dot(y, y))], dtype=torch.float64)
        dist.all_reduce(t)
        assert np.isclose(t.item(), float(np.dot(y_ref, y_ref)), rtol=1e-13)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_plan_with_gloo_world_size_2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


@pytest.mark.parametrize("kind", ["hex", "tet"])
@pytest.mark.parametrize("nranks", [2, 4])
def test_window_partition_equals_partition_of_the_global_mesh(nranks, kind):
    """Each rank builds only its z-slab (+2 layers per side) of the box; the resulting partition must be identical — geometry,
    connectivity, pattern, exchange plan — to the one cut from the full global mesh. Hex channel and the Kuhn-split tet box
    (BASELINE.json configs[4]: windowed generator for the 8-GPU run)."""
    nx, ny, nz = 5, 4, 12
    per_hex = 1 if kind == "hex" else 6
    box, slab = (syn.hex_box, syn.slab_partition) if kind == "hex" else (syn.tet_box, syn.tet_slab_partition)
    g = orc_b200.Mesh.from_arrays(*syn.mesh_args(box(nx, ny, nz)))
    syn.channel_bcs(g, fully_3d=(kind == "tet"))
    for r in range(nranks):
        # the global cuts must coincide with plane boundaries for the comparison: nz * r / nranks planes
        arrays, cuts, off, n_global = slab(nx, ny, nz, r, nranks)
        w = orc_b200.Mesh.from_arrays(*syn.mesh_args(arrays))
        syn.channel_bcs(w, fully_3d=(kind == "tet"))
        pw = w.partition_window(r, nranks, cuts, off, n_global)
        gcuts = [(nz * q // nranks) * nx * ny * per_hex for q in range(nranks)] + [nx * ny * nz * per_hex]
        pg = g.partition_window(r, nranks, gcuts, 0, n_global)
        assert pw.partition_info() == pg.partition_info()
        a, b = pw.export(), pg.export()
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        ma, mb = pw.partition_maps(), pg.partition_maps()
        for k in ma:
            assert np.array_equal(ma[k], mb[k]), k
        assert np.array_equal(pw.pattern()[1], pg.pattern()[1]) and np.array_equal(pw.levels(), pg.levels())


@pytest.mark.parametrize("kind", ["hex", "tet"])
@pytest.mark.parametrize("nparts", [2, 3, 4, 8])
def test_rcb_renumbering_makes_index_ranges_compact(kind, nparts):
    """A mesh whose numbering is not spatially coherent (here: shuffled) is renumbered by recursive coordinate bisection before
    it is partitioned by index ranges. The renumbered mesh is the same mesh (geometry permuted bit for
    bit, faces untouched), every part is one contiguous range with the sizes Mesh.partition cuts, and the halos shrink."""
    from orc_b200 import partition as part
    arrays = syn.hex_box(8, 6, 8) if kind == "hex" else syn.tet_box(4, 3, 4)
    n = int(arrays["n_cells"])
    rng = np.random.default_rng(5)
    shuffled = part.renumber_cells(arrays, rng.permutation(n))
    ms = orc_b200.Mesh.from_arrays(*syn.mesh_args(shuffled))
    renum, new_to_old = part.rcb_renumber(shuffled, nparts)
    mr = orc_b200.Mesh.from_arrays(*syn.mesh_args(renum))
    es, er = ms.export(), mr.export()
    assert np.array_equal(np.sort(new_to_old), np.arange(n))
    assert np.array_equal(er["cell_volume"], es["cell_volume"][new_to_old])
    assert np.array_equal(er["cell_centroid"], es["cell_centroid"][new_to_old])
    assert np.array_equal(er["face_area"], es["face_area"]) and np.array_equal(er["face_normal"], es["face_normal"])
    # every cell keeps its faces, in the same order
    for k in (0, n // 2, n - 1):
        o = new_to_old[k]
        assert np.array_equal(er["cell_face_indices"][er["cell_face_offsets"][k]:er["cell_face_offsets"][k + 1]],
                              es["cell_face_indices"][es["cell_face_offsets"][o]:es["cell_face_offsets"][o + 1]])
    # parts are boxes: the bounding boxes of two different parts overlap in at most a sliver along the split axes
    cuts = part.even_cuts(n, nparts)
    cc = er["cell_centroid"]
    for r in range(nparts):
        assert cuts[r + 1] > cuts[r]
    h_shuffled, h_rcb = part.halo_cells(ms, nparts), part.halo_cells(mr, nparts)
    h_structured = part.halo_cells(orc_b200.Mesh.from_arrays(*syn.mesh_args(arrays)), nparts)
    print(kind, nparts, "halo cells: shuffled", h_shuffled, "rcb", h_rcb, "structured numbering", h_structured)
    assert h_rcb < 0.6 * h_shuffled
    assert h_rcb <= 2.0 * h_structured + 8     # as good as a slab split of the structured numbering, up to the box shapes
    # the partition machinery accepts the renumbered mesh: owned ranges tile the cells, plans are consistent
    owned = 0
    for r in range(nparts):
        p = mr.partition(r, nparts)
        info = p.partition_info()
        owned += info["n_own"]
        assert info["g0"] == cuts[r] and info["g1"] == cuts[r + 1]
    assert owned == n
