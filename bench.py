#!/usr/bin/env python
"""bench.py — SIMPLE iterations/s of the B200 path on BASELINE.json's workload (synthetic hex channel, defaults:
CD1 momentum, SecondOrder pressure, Rhie-Chow, Multigrid with BiCGSTAB smoothing, 50 inner iterations, fp64).

One "step" = one SIMPLE iteration (the body of the loop at src/solver.rs:60-222 of the reference).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size n]            # our arm
  python bench.py --impl reference [--steps K] [--warmup W]                  # the reference's CPU path (oracle restatement)

Prints ONE JSON line (rank 0). See DESIGN.md §7 for what every key means and how the roofline figure is derived.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RHO, MU = 1000.0, 1e-3
P_RELAX = 1e-4        # pressure relaxation (README of the reference: "<< 0.1"); every other setting is the reference default
SPMV_SAMPLE = 3
RESET_EVERY = 6       # SIMPLE iterations between resets of the fields (see run_ours.step)
METRIC = "SIMPLE iters/s"


def workload_name(n):
    return f"synthetic {n}^3 hex channel ({n ** 3 / 1e6:.1f}M cells), SIMPLE + AMG-BiCGSTAB fp64"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        try:
            if self.proc:
                self.proc.terminate()
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def oracle_sample_rate(m, steps=1, warmup=0):
    """cell-updates/s of the CPU restatement of the reference path on an m^3 hex channel with the same settings (1 thread)."""
    from oracle import pyoracle as po
    from orc_b200 import synthetic as syn
    a = syn.hex_box(m, m, m)
    om = po.Mesh.from_arrays(*syn.mesh_args(a))
    syn.channel_bcs(om)
    n = om.n_cells
    z = [np.zeros(n) for _ in range(4)]
    if warmup:
        z = list(om.solve_steady(*z, po.Settings(pressure_relaxation=P_RELAX), RHO, MU, warmup, 0)[:4])
    t0 = time.perf_counter()
    out = om.solve_steady(*z, po.Settings(pressure_relaxation=P_RELAX), RHO, MU, steps, 0)
    dt = time.perf_counter() - t0
    return n, dt, out[5]


def run_reference(args, rank):
    """The reference's own CPU implementation of the path, timed on the host cores. The Rust crate cannot be built in this
    image (no cargo/rustc, no network: DESIGN.md §2), so this times oracle/ — its C++ restatement, single-threaded like ORC."""
    if rank != 0:
        return
    n = args.size
    total = max(1, args.steps + args.warmup)
    m = 24
    for cand in (64, 48, 40, 32, 24):          # ~14k cell-updates/s/core measured: keep the whole run under ~2.5 minutes
        if total * cand ** 3 / 14000.0 <= 150.0:
            m = cand
            break
    cells, dt, phases = oracle_sample_rate(m, steps=args.steps, warmup=args.warmup)
    cell_rate = cells * args.steps / dt
    value = cell_rate / n ** 3
    sample = (f"{args.steps} SIMPLE iteration(s) after {args.warmup} warm-up on a {m}^3 hex channel ({cells} cells), same settings; "
              f"iters/s scaled to {n}^3 by cell count (measured {cell_rate:.0f} cell-updates/s)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "iter/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n), "solver": "Multigrid(BiCGSTAB x50, 3 levels, Jacobi precond)", "momentum": "CD1",
                       "velocity_interpolation": "RhieChow", "pressure_interpolation": "SecondOrder"},
            "cpu_baseline": {"value": value, "unit": "iter/s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "cell_updates_per_s": cell_rate, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_e2e(args, world, dist, torch, orc_b200, mesh, settings, solver, ctx, barrier):
    """e2e: the reference-facing call (solve_steady through the C ABI) with HOST buffers, copies inside the timed region."""
    solver.reset()
    solver.iterate(2)
    start_fields = solver.get_fields()   # every e2e step starts from the same host fields (2 iterations from rest)
    e2e_steps = max(1, min(args.steps, 3))
    sets = [[torch.from_numpy(f.copy()).pin_memory().numpy() for f in start_fields] for _ in range(e2e_steps)]
    # one untimed call first: the one-shot entry creates its own device state (5 matrices, 15 vectors), and the caching
    # allocator has to see those sizes once — like the W warm-up steps of the resident leg
    warm = [f.copy() for f in start_fields]
    with contextlib.redirect_stdout(sys.stderr):   # the mirror prints the reference's "Solving..." lines; stdout carries the JSON line only
        orc_b200.solve_steady(mesh, *warm, settings, RHO, MU, 1, 0, ctx=ctx, on_report=lambda d: None)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        tk = time.perf_counter()
        with contextlib.redirect_stdout(sys.stderr):
            orc_b200.solve_steady(mesh, *sets[k], settings, RHO, MU, 1, 0, ctx=ctx, on_report=lambda d: None)
        print(f"[e2e] step {k}: {(time.perf_counter() - tk) * 1e3:.1f} ms", file=sys.stderr, flush=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    return world * e2e_steps / e2e_s



def run_ours(args, rank, world):
    import torch
    import orc_b200
    from orc_b200 import synthetic as syn

    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.size
    stream = torch.cuda.current_stream()
    ctx = orc_b200.Context(local_rank, stream.cuda_stream)

    # Multi-GPU: weak scaling. The global mesh has `world` x size^3 cells (1: n^3, 2: n x n x 2n, 4: n x 2n x 2n,
    # 8: (2n)^3 — the 256^3 / 16.8M-cell config at n = 128), cut into z-slabs of size^3 cells per rank. Every rank builds only
    # its slab (+2 layers per side) of the box and takes its partition from that window; halo exchange over NCCL send/recv,
    # BiCGSTAB scalars over NCCL allreduce, AMG hierarchy per partition.
    gshape = {1: (n, n, n), 2: (n, n, 2 * n), 4: (n, 2 * n, 2 * n), 8: (2 * n, 2 * n, 2 * n)}.get(world, (n, n, n * world))
    tet = args.mesh == "tet"
    if tet:
        # BASELINE.json configs[4] in kind: every hex of the lattice split into 6 tetrahedra (hex-major numbering), TVD-UMIST momentum.
        # The AMG smoother stays BiCGSTAB: with the reference's algorithm a Gauss-Seidel or Jacobi smoother panics on the coarse
        # levels. TVD makes a_u, a_v, a_w differ, so the three momentum solves run one after the other.
        arrays = syn.tet_box(*gshape)
        mesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(arrays))
        syn.channel_bcs(mesh, fully_3d=True)
        if world > 1:
            ctx.comm_init(rank, world)
            gmesh = mesh
            mesh = gmesh.partition(rank, world)   # every rank builds the global mesh: fine up to a few million tets per rank
            del gmesh
    elif world == 1:
        arrays = syn.hex_box(*gshape)
        mesh = orc_b200.Mesh.from_arrays(*syn.mesh_args(arrays))
        syn.channel_bcs(mesh)
    else:
        ctx.comm_init(rank, world)
        arrays, cuts, off, n_global = syn.slab_partition(*gshape, rank, world)
        window = orc_b200.Mesh.from_arrays(*syn.mesh_args(arrays))
        syn.channel_bcs(window)
        mesh = window.partition_window(rank, world, cuts, off, n_global)
        del window
    del arrays
    cells = mesh.partition_info()["n_own"] if world > 1 else mesh.n_cells
    settings = orc_b200.NumericalSettings(pressure_relaxation=P_RELAX)
    if tet:
        settings = orc_b200.NumericalSettings(pressure_relaxation=P_RELAX, momentum=orc_b200.MomentumDiscretization.TVD, limiter=orc_b200.TVD_UMIST)
    solver = orc_b200.SteadySolver(mesh, settings, RHO, MU, ctx)
    solver.set_fields(*(np.zeros(cells) for _ in range(4)))
    done = [0]

    def step():
        # The reference's algorithm does not converge on this mesh (oracle-confirmed at 64^3): its unguarded
        # BiCGSTAB / multigrid eventually produce NaN ("Multigrid diverged"). The work per iteration does not depend on that, so
        # the fields are put back to the start state every RESET_EVERY iterations; the reset (a few memsets) is inside the
        # timed region.
        if done[0] and done[0] % args.reset_every == 0:
            solver.reset()
        done[0] += 1
        return solver.iterate(1)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    # roofline leg, live in the timed region: CUDA events around every SPMV_SAMPLE-th SpMV launch only (an event pair costs a
    # few microseconds; around all ~3100 launches of a step it slowed the step by 15 %). The per-class breakdown comes from one
    # extra, untimed step below.
    ctx.prof_config(classes=["spmv"], sample_every=SPMV_SAMPLE)
    ctx.prof_enable(True)
    phases0 = solver.phase_ms()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    rep = None
    for _ in range(args.steps):
        rep = step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    prof = ctx.prof_get()
    sp_ref_bytes = ctx.prof_ref_bytes("spmv")
    sp_detail = ctx.prof_spmv_detail()
    ctx.prof_enable(False)
    phases = {k: v - phases0[k] for k, v in solver.phase_ms().items()}   # timed steps only
    batched = solver.batched
    ctx.prof_config(classes=None, sample_every=1)
    ctx.prof_enable(True)
    step()                                  # untimed: device time per kernel class, events around every launch
    classes = ctx.prof_get()
    ctx.prof_enable(False)
    clocks = sampler.finish() if sampler else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # weak scaling: one global problem of world x size^3 cells; value counts size^3-cell-equivalent iterations (= cell-updates/s / size^3)
    value = world * args.steps / (ms * 1e-3)

    # ---- e2e: the reference-facing call (solve_steady through the C ABI) with HOST buffers, copies inside the timed region
    if args.no_e2e:
        e2e_value = None
    else:
        e2e_value = run_e2e(args, world, dist, torch, orc_b200, mesh, settings, solver, ctx, barrier)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    sp_ms, sp_bytes, sp_count = prof["spmv"]
    achieved = sp_bytes / (sp_ms * 1e-3) / 1e9 if sp_ms > 0 else 0.0
    achieved_ref = sp_ref_bytes / (sp_ms * 1e-3) / 1e9 if sp_ms > 0 else 0.0
    # the same timed launches per AMG level (rows) and systems per launch; the matrices of one level differ from step to step and
    # between the momentum and the pressure system, so they are pooled by (rows, systems). R has <= 4 entries per row, R^T <= 2.
    pool = {}
    for r, z, k, t, b, cnt in sp_detail:
        if cnt and t > 0 and z > 4.5 * r:
            e = pool.setdefault((r, k), [0.0, 0.0, 0, 0.0])
            e[0] += t; e[1] += b; e[2] += cnt; e[3] += z * cnt
    by_matrix = [{"rows": r, "systems_per_launch": k, "entries_per_row": round(e[3] / e[2] / r, 1), "launches_timed": e[2],
                  "us_per_launch": round(1e3 * e[0] / e[2], 1), "GB/s": round(e[1] / (e[0] * 1e-3) / 1e9, 1),
                  "frac": round(e[1] / (e[0] * 1e-3) / 1e9 / peak, 3) if peak else None}
                 for (r, k), e in sorted(pool.items(), key=lambda kv: (-kv[0][0], kv[0][1]))]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(str(n))
    levels = solver.level_sizes()
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not tet:   # the CPU sample is the hex channel
        m = args.cpu_sample
        ccells, cdt, _ = oracle_sample_rate(m)
        cpu = {"value": ccells / cdt / n ** 3, "unit": "iter/s", "cores": 1, "kind": "port",
               "sample": f"1 SIMPLE iteration on a {m}^3 hex channel ({ccells} cells, {cdt:.1f} s), same settings, C++ restatement of the "
                         f"reference path (oracle/), single thread like ORC; scaled to {n}^3 by cell count"}
    # SURVEY.md §8d byte model of the WHOLE SIMPLE iteration as the reference performs it: 4 solves x 50 BiCGSTAB iterations on
    # every level (multiplicity 1, 2, 2, 1: pre- and post-smoothing), one iteration = 24 nnz_l + 152 n_l bytes. This is the
    # figure the "60 % of HBM roofline" target of BASELINE.json is stated in (SURVEY: 0.37 MB per cell per iteration for hexes).
    mult = [1, 2, 2, 1]
    per_set = sum(m * (24.0 * z + 152.0 * r) for m, (r, z) in zip(mult, levels)) if len(levels) == 4 else None
    step_model = None
    if per_set:
        model_bytes = 4 * 50 * per_set
        eff = model_bytes / (ms / args.steps * 1e-3) / 1e9
        step_model = {"bytes_per_iteration_model": model_bytes, "bytes_per_cell": model_bytes / cells, "effective_GB/s": eff,
                      "frac_of_measured_peak": eff / peak if peak else None, "frac_of_nominal_8TB/s": eff / 8000.0,
                      "note": "whole-iteration throughput in the reference's own byte model (SURVEY.md §8d): 4 solves x 50 BiCGSTAB iterations x "
                              "levels (1,2,2,1) x (24 nnz + 152 n) bytes, divided by the measured time per iteration — set-up, assembly and "
                              "launch gaps included in the time, the lockstep saving counted as throughput"}
    line = {
        "metric": METRIC, "value": value, "unit": "iter/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n) if not tet else f"synthetic tet box: {n}^3 lattice x 6 tets ({cells / 1e6:.2f}M cells per GPU), SIMPLE + AMG-BiCGSTAB fp64",
                   "solver": "Multigrid(BiCGSTAB x50, 3 levels, Jacobi precond)", "momentum": "TVD-UMIST" if tet else "CD1",
                   "velocity_interpolation": "RhieChow", "pressure_interpolation": "SecondOrder", "assembly_mode": "exact",
                   "pressure_relaxation": P_RELAX, "fields_reset_every": args.reset_every,
                   "momentum_solves": ("u, v, w in lockstep: a_u == a_v == a_w bit for bit (checked on the device every iteration), one matrix "
                                       "pass and one AMG hierarchy for the three systems; every system's arithmetic is that of its own solve"
                                       if batched else "three sequential solves"),
                   "parallelism": "1 GPU" if world == 1 else f"{world} z-slabs of {n}^3 cells, NCCL halo send/recv + allreduce, per-partition AMG",
                   "global_mesh": list(gshape), "value_counts": f"SIMPLE iterations of one GPU's share of the mesh ({cells} cells): cell-updates/s / {cells}",
                   "l2": "inputs larger than L2 (fine matrix 175 MB at 128^3, 5 matrices + coarse levels); no flush needed",
                   "amg_levels_rows_nnz": levels},
        "cell_updates_per_s": value * cells,
        "roofline": {"bound": "hbm", "kernel": "k_spmv (all fused epilogues, all AMG levels)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src, "launches": sp_count,
                     "bytes_per_launch_model": "12*nnz_l + 4*n_l + 16*K*n_l of the level it runs on (K = systems per launch: 3 for the lockstep momentum solves, 1 for p')",
                     "timed": f"CUDA events around every {SPMV_SAMPLE}rd SpMV launch inside the timed region ({sp_count} launches)",
                     "by_matrix": by_matrix,
                     "by_matrix_note": "the same timed launches per AMG level (rows) and systems per launch; the short-row R / R^T products are omitted; "
                                       "the coarse levels are bound by the L1 gather pipe, not by HBM (profiles/r1_spmv_k3_ncu.txt)",
                     "achieved_in_reference_units": achieved_ref,
                     "reference_units_note": "same launches counted as the reference's SpMVs (12*nnz + 20*n each): a lockstep launch does three of them in one matrix pass",
                     "time_share_of_step": classes["spmv"][0] / max(1e-9, sum(v[0] for v in classes.values())),
                     "whole_iteration": step_model},
        "kernel_classes_ms_per_step": {k: v[0] for k, v in classes.items()},
        "kernel_classes_note": "device time per kernel class of ONE extra untimed step with events around every launch",
        "phases_ms_per_step": {k: v / args.steps for k, v in phases.items()},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "iter/s", "h2d_bytes_per_step": 32 * cells, "d2h_bytes_per_step": 32 * cells,
                "call": "orc_solve_steady(iteration_count=1) per step, pinned host u/v/w/p", "steps": max(1, min(args.steps, 3))},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "last_report": rep,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=int(os.environ.get("ORC_BENCH_N", "128")), help="hex channel is size^3 cells")
    ap.add_argument("--cpu-sample", type=int, default=64, help="edge of the hex box the CPU baseline is timed on")
    ap.add_argument("--mesh", default="hex", choices=["hex", "tet"], help="hex: the headline channel; tet: size^3 lattice split into 6 tets per "
                    "hex, TVD-UMIST momentum (BASELINE.json configs[4] in kind)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reset-every", type=int, default=RESET_EVERY, help="SIMPLE iterations between resets of the fields (the reference's "
                    "algorithm diverges on the synthetic boxes after a few iterations; sooner on larger ones)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs under ncu)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
