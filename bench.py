#!/usr/bin/env python
"""bench.py — Newton steps/s of the linearised-PNP hot path (EAFE assembly + AMG-FGMRES solve).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # CPU reference arm (oracle, 1 core)

A "step" is one full Newton step (benchmarks/pnp_exact_solution/main.cpp:334-345: assemble the
Jacobian with the EAFE overwrite and the rhs, AMG setup, FGMRES solve, u += du, residual norms)
started from the Dirichlet interpolant of the pnp_exact_solution physics on an n^3 Kuhn box mesh
of [-1,1]^3 (BASELINE.json configs[4]; default n=256 = 50.9 M dof, the north-star size). Every
timed step does the same work: the state is reset to the initial guess first.
  value : steps/s with the initial guess already in HBM (device-to-device reset)
  e2e   : steps/s through the C ABI with HOST buffers: pnpb200_set_solution (H2D) +
          pnpb200_newton_step + pnpb200_get_solution (D2H) inside the timed region
Timing: CUDA events on the handle's stream, max over ranks; inputs (matrix ~18 GB) exceed L2.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "newton_steps_per_s"
UNIT = "Newton steps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def byte_model(stats, nv, nc, krylov_its, restart, kcycle_levels=0, shortcuts=False, stages=False):
    """Algorithmic bytes of one Newton step, SURVEY.md §8(d), from the run's actual level sizes
    and iteration count (fp64 = 8 B, int32 = 4 B; every array touched once per logical pass).
    shortcuts=False is the SURVEY model (a cycle = 3 full matrix sweeps per level, a matvec = 1);
    shortcuts=True counts what the [L | D | U] schedule of this library actually has to read:
    forward sweep from zero = L, residual after it = U, backward sweep = full row, and the matvec
    after a cycle = L (2.5 instead of 4 matrix passes per cycle + matvec)."""
    N = 3 * nv
    rows = [stats.level_rows[l] for l in range(stats.n_levels)]
    nnzb = [stats.level_nnzb[l] for l in range(stats.n_levels)]
    b_mat = [nnzb[l] * 76 + (rows[l] + 1) * 4 for l in range(len(rows))]
    b_spmv = [b_mat[l] + 2 * 3 * rows[l] * 8 for l in range(len(rows))]
    # half-row pass: the off-diagonal blocks split evenly between L and U
    b_half = [(nnzb[l] - rows[l]) // 2 * 76 + (rows[l] + 1) * 4 + 2 * 3 * rows[l] * 8 for l in range(len(rows))]
    sweeps3 = [(2 * b_half[l] + b_spmv[l]) if shortcuts else 3 * b_spmv[l] for l in range(len(rows))]
    matvec = [b_half[l] if shortcuts else b_spmv[l] for l in range(len(rows))]
    b_asm = nc * 16 + nv * 24 + N * 8 + nv * 16 + N * 8 * 2 + nnzb[0] * 72
    b_eafe = 2 * (nnzb[0] * 12 + nv * 24)
    b_res = nc * 16 + nv * 24 + N * 8 * 2 + nv * 16 + N * 8 * 3
    b_setup = sum(2 * b_mat[l] + b_mat[l + 1] + rows[l] * 76 for l in range(len(rows) - 1))
    # cycles per preconditioner application on each level: the K-cycle (kcycle_levels > 0) runs two
    # cycles + two matvecs + ~12 vector passes per visit of a K level
    nl = len(rows)
    cyc = [1] * nl
    b_vc = 0
    for l in range(nl - 1):
        if l >= 1:
            if l <= kcycle_levels:
                cyc[l] = 2 * cyc[l - 1]
                b_vc += cyc[l - 1] * (2 * matvec[l] + 12 * 3 * rows[l] * 8)
            else:
                cyc[l] = cyc[l - 1]
        b_vc += cyc[l] * (sweeps3[l] + 2 * 3 * (rows[l] + rows[l + 1]) * 8)
    b_kry = 0
    for it in range(max(krylov_its, 0)):
        j = it % restart
        b_kry += matvec[0] + b_vc + (2 * (j + 1) + 2) * N * 8
    total = b_asm + b_eafe + 2 * b_res + b_setup + b_kry + 4 * N * 8
    if stages:
        # the library's stage timers: "assemble" = Jacobian + EAFE + the rhs residual, "residual" = the
        # residual after the update (the one the next step re-uses as its rhs)
        return {"assemble": b_asm + b_eafe + b_res, "amg_setup": b_setup, "krylov": b_kry, "residual": b_res, "total": total}
    return total


CPU_EST_S = {16: 0.3, 24: 1.2, 32: 3.5, 40: 7.0, 48: 13.0, 56: 22.0, 64: 36.0}     # oracle, 1 core, first Newton step (measured, r01/r02)


def cpu_sample_n_for(steps, budget_s, cap):
    """largest sample box whose `steps` full Newton steps fit the wall budget of the CPU arm"""
    if cap < min(CPU_EST_S):
        return cap
    best = min(CPU_EST_S)
    for n in sorted(CPU_EST_S):
        if n <= cap and steps * CPU_EST_S[n] <= budget_s:
            best = n
    return best


def cpu_reference_leg(sample_n, steps=1):
    """The reference path on the host: the oracle (the reference's OWN compiled element kernels when
    oracle/_ref is present + the restated DOLFIN assembler / FASP UA-AMG V(1,1) + pvfgmres of bsr.dat),
    serial like the reference (CMakeLists.txt links FASP without OpenMP), first Newton step of the same
    physics on a sample_n^3 box. Nothing is extrapolated: returns the measured seconds of every step,
    the dof count and the Krylov iterations of that box."""
    from oracle import oracle as O
    used_ref = False
    if O.ref_available():
        try:
            O.use_ref_kernels(True); used_ref = True
        except Exception:
            used_ref = False
    ts, its = [], None
    for _ in range(steps):
        t0 = time.perf_counter()
        _, res = O.newton_exact(sample_n, sample_n, sample_n, 2.0, 2.0, 2.0, max_newton=1)
        ts.append(time.perf_counter() - t0)
        its = res.krylov_its[0]
    if used_ref:
        O.use_ref_kernels(False)
    dof = 3 * (sample_n + 1) ** 3
    return ts, dof, its, used_ref


def workload_name(n):
    return "pnp_exact_solution physics, %d^3 Kuhn box of [-1,1]^3 (%.2f M dof), Newton step 1 from the Dirichlet interpolant" % (n, 3 * (n + 1) ** 3 / 1e6)


def multi_gpu_parity(pkg, problems, partition_fn, dist, rank, world, dev, it, amg):
    """Correctness gate of the N>1 path, OUTSIDE the timed region: the partitioned
    Newton step (halo exchange + allreduce over NCCL) against the single-GPU path on the same global
    mesh, same bars as tests/test_gpu_multi.py: initial residual norms 1e-12, three Newton steps'
    residuals 2e-3 (different aggregates per partition -> different but equally valid preconditioner),
    solution 1e-4. Raises on failure; the result goes into the JSON line."""
    import numpy as np
    import torch
    nx, ny, nz, lx, ly, lz = 24, 10, max(24, 4 * world), 2.0, 0.8, 2.0
    P = partition_fn(nx, ny, nz, lx, ly, lz, rank, world)
    g = pkg.PNP(P["xyz"], P["cells"], device=dev)
    uid = [pkg.capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g.comm_init(uid[0], rank, world); g.set_partition(P)
    if "color" in P:
        g.set_coloring(P["color"], 4)
    f = problems.exact_solution_fields(P["xyz"][0::3])
    g.set_coefficients(f["eps"], f["fixed"], f["diff"], f["val"], f["reac"])
    g.set_dirichlet((3 * P["bc_vertices"][:, None] + np.arange(3)).ravel().astype(np.int32))
    g.set_solution(f["uu0"])
    r0 = g.residual()
    hist = [g.newton_step(it, amg) for _ in range(3)]
    uu = g.get_solution()[:3 * P["n_own"]]
    gd = (3 * np.asarray(P["gids"])[:P["n_own"], None] + np.arange(3)).ravel()
    g.close()
    # every rank checks its owned part against the single-GPU solution computed by rank 0
    out = [None]
    if rank == 0:
        s = pkg.PNP(box=(nx, ny, nz, lx, ly, lz), device=dev)
        pr = problems.exact_solution_problem(nx, ny, nz, lx, ly, lz)
        s.set_coefficients(pr["eps"], pr["fixed"], pr["diff"], pr["val"], pr["reac"]); s.set_dirichlet(pr["bc_dofs"]); s.set_solution(pr["uu0"])
        q0 = s.residual()
        ref = [s.newton_step(it, amg) for _ in range(3)]
        out = [(q0, ref, s.get_solution())]
        s.close()
    dist.broadcast_object_list(out, src=0)
    q0, ref, uu_ref = out[0]
    res0 = max(abs(r0[0] - q0[0]) / q0[0], abs(r0[1] - q0[1]) / q0[1])
    step = max(abs(h[1] - r[1]) / r[1] for h, r in zip(hist, ref))
    sol = float(np.abs(uu - uu_ref[gd]).max())
    ok = res0 <= 1e-12 and step <= 2e-3 and sol <= 1e-4 and all(h[0] > 0 for h in hist) and all(r[0] > 0 for r in ref)
    t = torch.tensor([res0, step, sol, 0.0 if ok else 1.0], device="cuda:%d" % dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res0, step, sol, bad = [float(v) for v in t.tolist()]
    rep = {"mesh": "%dx%dx%d box, %d ranks" % (nx, ny, nz, world), "initial_residual_rel_diff": res0, "newton_step_residual_rel_diff": step,
           "solution_max_abs_diff": sol, "krylov_its_partitioned": [h[0] for h in hist], "krylov_its_single_gpu": [r[0] for r in ref],
           "bars": {"initial_residual": 1e-12, "newton_step_residual": 2e-3, "solution": 1e-4}, "ok": bad == 0.0}
    if bad != 0.0:
        raise RuntimeError("multi-GPU parity gate failed: %r" % (rep,))
    return rep


_JSON_FD = None


def claim_stdout():
    """Only the JSON line may reach stdout: libraries write there too (NCCL prints its version
    banner on fd 1 when the communicator comes up), so fd 1 is pointed at stderr for the run and the
    result line goes to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, line)
    else:
        os.write(_JSON_FD, line)


def stats_nnzb0(g):
    return float(g.stats().level_nnzb[0])


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--box-n", dest="n", type=int, default=int(os.environ.get("PNPB200_BENCH_N", "256")))
    ap.add_argument("--cpu-sample-n", type=int, default=64,
                    help="largest sample box of the CPU legs (the reference arm shrinks it until its K steps fit --cpu-budget-s)")
    ap.add_argument("--cpu-budget-s", type=float, default=300.0, help="wall budget of the timed steps of --impl reference")
    ap.add_argument("--cycle", default=os.environ.get("PNPB200_BENCH_CYCLE", "NA"),
                    help="AMG_cycle_type: V (bsr.dat) or NA (FASP nonlinear AMLI = K-cycle; needed above ~20 M dof where V exceeds itsolver_maxit)")
    ap.add_argument("--profile", action="store_true",
                    help="short run for ncu: 1 warm-up + 1 timed step, no e2e / cpu_baseline / probe legs")
    ap.add_argument("--maxit", type=int, default=int(os.environ.get("PNPB200_BENCH_MAXIT", "100")))
    ap.add_argument("--scaling", default=os.environ.get("PNPB200_BENCH_SCALING", "weak"), choices=["weak", "strong"],
                    help="N > 1: weak = n^3 vertices per GPU (the global box grows with N), strong = the n^3 box split over the N GPUs")
    ap.add_argument("--partition", default=os.environ.get("PNPB200_BENCH_PARTITION", "slabs"), choices=["slabs", "bricks"],
                    help="N > 1: z-slabs (weak: box n x n x nN) or px x py x pz bricks (weak: box n px x n py x n pz), csrc/partition.cpp")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.n
    nv, nc = (n + 1) ** 3, 6 * n ** 3
    config = {"workload": workload_name(n), "n": n, "dofs": 3 * nv, "cells": nc,
              "solver": "bsr.dat (UA-AMG %s-cycle(1,1) mc-GS + vFGMRES(30), tol 1e-6, maxit %d)" % (args.cycle, args.maxit),
              "amg_cycle_type": args.cycle, "l2_policy": "inputs_larger_than_l2"}

    if args.impl == "reference":
        # CPU arm: the reference path on the host cores, UN-EXTRAPOLATED. Every timed step is one full
        # first Newton step of the same physics on an m^3 sample box, m the largest box whose K steps fit
        # the wall budget; value = measured steps/s AT THAT m (config.n = m says so). The reference's
        # solver configuration is bsr.dat's V-cycle, which is what the oracle runs.
        if rank != 0:
            return 0
        W, K = max(args.warmup, 0), max(args.steps, 1)
        m = cpu_sample_n_for(K, args.cpu_budget_s, args.cpu_sample_n)
        for _ in range(min(W, 3)):
            cpu_reference_leg(16)                # warm-up: page in the libraries / element kernels (small box)
        ts, dof, its, used_ref = cpu_reference_leg(m, steps=K)
        sec = sum(ts) / len(ts)
        val = 1.0 / sec
        cfg_ref = {"workload": workload_name(m) + " [bounded sample of the %d^3 workload]" % n, "n": m, "dofs": dof, "cells": 6 * m ** 3,
                   "solver": "bsr.dat (UA-AMG V-cycle(1,1) GS + vFGMRES(30), tol 1e-6, maxit 100)", "amg_cycle_type": "V",
                   "bench_workload_n": n}
        sample = ("%d timed first Newton steps on a %d^3 box (%d dof, %d Krylov its), mean %.2f s per step, serial (1 core of %d) like the reference; "
                  "not extrapolated%s" % (K, m, dof, its, sec, os.cpu_count() or 0, "; reference element kernels (oracle/_ref)" if used_ref else ""))
        emit({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
              "warmup": W, "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg_ref, "extrapolated": False,
              "krylov_its_per_step": its, "us_per_dof": 1e6 * sec / dof, "seconds_per_step_each": ts,
              "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference" if used_ref else "port", "sample": sample},
              # context only, never the headline: what the measured us/dof would mean at the bench size if the
              # cost were linear in dof (it is not: V-cycle iterations grow with n and exceed maxit at 256^3)
              "linear_in_dof_projection": {"n": n, "steps_per_s": val * dof / float(3 * nv), "note": "projection, not a measurement"},
              "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return 0

    import numpy as np
    import torch
    from pnpb200_loader import load_package
    from importlib import import_module
    pkg = load_package()
    problems = import_module("modular_pnp_b200.problems")

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    torch.cuda.set_device(dev)

    cyc = {"V": 1, "W": 2, "A": 3, "NA": 4}[args.cycle]
    it, amg = pkg.default_params(maxit=args.maxit, cycle_type=cyc)
    parity = None
    partition = import_module("modular_pnp_b200.partition")
    grid_of = (lambda nx, ny, nz: (1, 1, world)) if args.partition == "slabs" else (lambda nx, ny, nz: partition.brick_grid(world, nx, ny, nz))

    def part_fn(nx, ny, nz, lx, ly, lz, r, w):
        return partition.box_brick_partition(nx, ny, nz, lx, ly, lz, r, w, grid=grid_of(nx, ny, nz))
    if world > 1 and not args.profile:
        parity = multi_gpu_parity(pkg, problems, part_fn, dist, rank, world, dev, it, amg)
    if world == 1:
        g = pkg.PNP(box=(n, n, n, 2.0, 2.0, 2.0), device=dev)
        pr = problems.exact_solution_problem(n, n, n, 2.0, 2.0, 2.0)
        g.set_coefficients(pr["eps"], pr["fixed"], pr["diff"], pr["val"], pr["reac"])
        g.set_dirichlet(pr["bc_dofs"])
        n_local_dofs = 3 * nv
        global_dofs, part_desc = 3 * nv, "1 GPU"
    else:
        # block rows (vertices) partitioned by the library's C++ partitioner (csrc/partition.cpp), one part per GPU;
        # ghost vertices by halo exchange over NCCL. weak: every GPU owns ~n^3 vertices (slabs: the box is
        # n x n x nN on [-1,1]^2 x [-N,N]; bricks: n px x n py x n pz); strong: the n^3 box of [-1,1]^3 is split.
        if args.scaling == "strong":
            gn, gl = (n, n, n), (2.0, 2.0, 2.0)
        elif args.partition == "slabs":
            gn, gl = (n, n, n * world), (2.0, 2.0, 2.0 * world)
        else:
            px, py, pz = partition.brick_grid(world, n, n, n)
            gn, gl = (n * px, n * py, n * pz), (2.0 * px, 2.0 * py, 2.0 * pz)
        part = part_fn(gn[0], gn[1], gn[2], gl[0], gl[1], gl[2], rank, world)
        g = pkg.PNP(part["xyz"], part["cells"], device=dev)
        uid = [pkg.capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(uid[0], rank, world)
        g.set_partition(part)
        g.set_coloring(part["color"], 4)       # (i + j + k) mod 4 of the global lattice: 4 colour launches per sweep, like one GPU
        pr = problems.exact_solution_fields(part["xyz"][0::3])
        g.set_coefficients(pr["eps"], pr["fixed"], pr["diff"], pr["val"], pr["reac"])
        g.set_dirichlet((3 * part["bc_vertices"][:, None] + np.arange(3)[None, :]).ravel().astype(np.int32))
        n_local_dofs = 3 * part["n_loc"]
        global_dofs = 3 * (gn[0] + 1) * (gn[1] + 1) * (gn[2] + 1)
        part_desc = "%d GPUs, %s of a %dx%dx%d box (%.1f M dof, %s scaling), grid %s" % (
            world, args.partition, gn[0], gn[1], gn[2], global_dofs / 1e6, args.scaling, part["grid"])
        del part
    uu0_host = torch.from_numpy(pr["uu0"]).pin_memory()
    uu_out_host = torch.empty_like(uu0_host).pin_memory()
    uu0_dev = uu0_host.to("cuda:%d" % dev)
    stream = torch.cuda.ExternalStream(g.stream, device=dev)
    L, h = g.L, g.h
    import ctypes as C

    def step_resident():
        L.pnpb200_set_solution_device(h, C.c_void_p(uu0_dev.data_ptr()))
        return g.newton_step(it, amg)

    def step_e2e():
        L.pnpb200_set_solution(h, uu0_host.numpy())
        r = g.newton_step(it, amg)
        L.pnpb200_get_solution(h, uu_out_host.numpy())
        return r

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches, its, last = 0, [], None
        e0.record(stream)
        for _ in range(K):
            last = fn()
            launches += g.stats().launches
            its.append(last[0])
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda:%d" % dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, its, last

    W, K = max(args.warmup, 3), max(args.steps, 1)
    if args.profile:
        W, K = 1, 1
    L.pnpb200_set_solution_device(h, C.c_void_p(uu0_dev.data_ptr()))
    init_res = g.residual()                 # ||F(u0)||: the residual after the step is reported relative to it
    for _ in range(W):
        step_resident()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    if args.profile:                # ncu --profile-from-start off: only the timed step is captured
        torch.cuda.cudart().cudaProfilerStart()
    ms, launches, its, last = timed(step_resident, K)
    if args.profile:
        torch.cuda.cudart().cudaProfilerStop()
    clocks = sampler.stop() if rank == 0 else None
    if args.profile:
        if rank == 0:
            emit(({"profile_run": True, "ms_per_step": ms / K, "krylov_its": its, "gpu_launches": launches,
                              "config": config}))
        g.close()
        return 0
    step_e2e()
    ms_e2e, _, _, _ = timed(step_e2e, K)
    stats = g.stats()
    halo = g.halo_counts() if world > 1 else None      # exchanges issued by this rank: (peer-memory kernel, NCCL send/recv)
    kits = its[-1]

    # roofline of the dominant kernel. profiles/README.md (ncu launch list): the multicolour
    # Gauss-Seidel colour launches are the largest share of the device time. One "launch" of the
    # roofline line = one full backward sweep over the fine-level matrix as the cycle runs it
    # (gs_stream_kernel<S_GSD>: all colour launches, also writes delta = x_new - x_old), timed live
    # with CUDA events on the library's stream. The other passes of a cycle are listed next to it.
    peak, peak_src = peaks()
    spmv_ms, spmv_bytes = g.probe_kernel(0, 20)
    gs_ms, gs_bytes = g.probe_kernel(9, 10)
    gs_bytes += 3 * nv * 8                     # the delta vector it writes
    passes = {}
    for w, name in [(7, "forward sweep from zero (L only)"), (8, "residual after it (U only)"),
                    (10, "A*z after the backward sweep (L only)"), (1, "full residual r = b - A x")]:
        pm, pb = g.probe_kernel(w, 10)
        passes[name] = {"ms": pm, "full_row_algorithmic_gbs": pb / pm / 1e6}
    achieved = gs_bytes / gs_ms / 1e6
    # values per block the solve's matrix passes read on the fine level: 7 when the PNP block structure was packed
    # (entries (1,2), (2,1) structurally zero: 60 instead of 76 bytes per block), else 9
    fmt = g.value_format(0)
    gs_kernel_bytes = gs_bytes - (16.0 * stats_nnzb0(g) if fmt == 7 else 0.0)
    klev = 3 if cyc != 1 else 0
    kk = kits if kits > 0 else stats.krylov_its
    b_step = byte_model(stats, nv, nc, kk, it.restart, kcycle_levels=klev)
    b_sched = byte_model(stats, nv, nc, kk, it.restart, kcycle_levels=klev, shortcuts=True)
    b_strict = byte_model(stats, nv, nc, kk, it.restart, kcycle_levels=0)      # SURVEY 8(d) as written: one V(1,1) per iteration
    b_stage = byte_model(stats, nv, nc, kk, it.restart, kcycle_levels=klev, stages=True)
    step_gbs = b_step / (ms / K) / 1e6
    stage_ms = {"assemble": stats.ms_assemble, "amg_setup": stats.ms_setup, "krylov": stats.ms_krylov, "residual": stats.ms_residual}
    stage_frac = {k: (b_stage[k] / stage_ms[k] / 1e6 / peak if stage_ms[k] > 0 else None) for k in stage_ms}
    l2_0, _ = init_res
    # traffic: dram bytes of one sweep (dram__bytes_read.sum + dram__bytes_write.sum over its colour
    # launches) from the committed single-pass ncu capture taken at the same n (tools/ncu_traffic.sh)
    traffic, traffic_note = None, None
    for tp in (os.path.join(ROOT, "profiles", "r02_traffic_gsd_n%d.json" % n), os.path.join(ROOT, "profiles", "r02_traffic_gsd_n256.json"),
               os.path.join(ROOT, "profiles", "r01_traffic_gsd_n256.json")):
        if not os.path.exists(tp):
            continue
        try:
            with open(tp) as f:
                tj = json.load(f)
            ratio = tj["gs_sweep"]["dram_bytes_per_sweep"] / tj["gs_sweep"]["algorithmic_bytes"]
            traffic_note = "ncu at n=%d (%s): %.4g B dram per sweep = %.2fx algorithmic (x re-read once per colour, partial sectors)" % (
                tj.get("n"), os.path.basename(tp), tj["gs_sweep"]["dram_bytes_per_sweep"], ratio)
            ncol_cap = len(tj["gs_sweep"].get("per_launch", [])) or None
            if tj.get("n") == n and (ncol_cap is None or ncol_cap == stats.level_colors[0]) and tj.get("values_per_block", 9) == fmt:
                traffic = tj["gs_sweep"]["dram_bytes_per_sweep"]
            elif ncol_cap is not None and ncol_cap != stats.level_colors[0]:
                traffic_note += "; captured with %d colour launches per sweep (graph colouring, PNPB200_BOX_COLORING=0), this run sweeps %d colours (closed-form BoxMesh colouring: x is re-read %d instead of %d times) -> traffic not re-measured, null" % (
                    ncol_cap, stats.level_colors[0], stats.level_colors[0] - 1, ncol_cap - 1)
            break
        except Exception:
            pass

    if rank == 0:
        # weak scaling: one unit of work = a Newton step on one n^3-box-equivalent (the N-GPU job
        # steps an N-times larger mesh, so the whole-job aggregate is N * global steps/s)
        # strong scaling: the job steps the same n^3 box, value = literal steps/s
        units = world if args.scaling == "weak" else 1
        value = units * K / (ms / 1e3)
        e2e_v = units * K / (ms_e2e / 1e3)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": dict(config, parallelism=part_desc + ("" if world == 1 else ": halo exchange at every distributed AMG level (peer-memory kernel over NVLink, NCCL send/recv as fallback: see halo_exchanges) + fp64 NCCL allreduce, aggregates local to a rank, coarse levels below 40 000 global rows replicated on every rank"),
                              global_dofs=global_dofs, global_steps_per_s=K / (ms / 1e3)),
               "krylov_its_per_step": its, "krylov_status_ok": bool(kits > 0), "amg_cycle_type": args.cycle,
               "l2_residual_after_step": None if last is None else last[1],
               "relative_residual_after_step": None if last is None else last[1] / l2_0,
               "multi_gpu_parity": parity,
               "halo_exchanges": None if halo is None else {"peer_memory_kernel": halo[0], "nccl_send_recv": halo[1], "note": "rank 0, whole run"},
               "levels": [[stats.level_rows[l], stats.level_nnzb[l], stats.level_colors[l]] for l in range(stats.n_levels)],
               "stage_ms_last_step": stage_ms,
               "stage_roofline_frac": dict(stage_frac, note="SURVEY 8(d) algorithmic bytes of the stage / its time / peak; assemble = B_asmJ + B_eafe + one B_res (the rhs), residual = one B_res"),
               "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": n_local_dofs * 8 * world, "d2h_bytes_per_step": n_local_dofs * 8 * world},
               "gpu_launches": launches,
               "clocks": clocks,
               "roofline": {"bound": "hbm", "kernel": "gs_stream_kernel<S_GSD> (csrc/gs_stream.cu; bsrt_row_kernel<M_GSD> with PNPB200_GS_STREAM=0): one backward multicolour Gauss-Seidel sweep of the fine level, writes x and delta (%d colour launches)" % stats.level_colors[0],
                            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                            "traffic_note": traffic_note, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": gs_bytes, "ms_per_launch": gs_ms,
                            "block_format": {"values_per_block": fmt, "bytes_per_block_read": 60 if fmt == 7 else 76,
                                             "kernel_bytes_per_launch": gs_kernel_bytes, "kernel_bytes_achieved": gs_kernel_bytes / gs_ms / 1e6,
                                             "kernel_bytes_frac": gs_kernel_bytes / gs_ms / 1e6 / peak,
                                             "note": "achieved / frac above are on SURVEY 8(d)'s figure (76 B per stored 3x3 block); "
                                                     "the solve reads a 7-value copy of every block where the PNP structure allows it (no cation-anion "
                                                     "coupling: entries (1,2) and (2,1) are zero on every level), i.e. 60 B per block: kernel_bytes_* "
                                                     "count what the kernel has to read in that format, and `traffic` is to be compared with kernel_bytes_per_launch"},
                            "other_passes": passes,
                            "spmv": {"kernel": "gs_stream_kernel<S_SPMV> (y = A x, fine level)", "achieved": spmv_bytes / spmv_ms / 1e6,
                                     "frac": spmv_bytes / spmv_ms / 1e6 / peak, "ms": spmv_ms, "algorithmic_bytes": spmv_bytes},
                            "whole_step": {"algorithmic_bytes": b_step, "achieved": step_gbs, "frac": step_gbs / peak,
                                           "note": "algorithmic_bytes = SURVEY 8(d) model (3 full sweeps + 1 matvec per cycle); scheduled_bytes = what the [L|D|U] schedule reads (2.5 passes)",
                                           "scheduled_bytes": b_sched, "scheduled_frac": b_sched / (ms / K) / 1e6 / peak,
                                           "strict_v11_bytes": b_strict, "strict_v11_frac": b_strict / (ms / K) / 1e6 / peak,
                                           "format_note": "all three byte counts use SURVEY 8(d)'s 76 B per block; with the 7-value block copy (roofline.block_format) the matrix passes of the solve read 60 B per block",
                                           "strict_note": "strict_v11 = the 8(d) formula as written (ONE V(1,1) cycle per Krylov iteration) on this run's levels and iteration count; algorithmic_bytes adds the K-cycle's repeated coarse-level visits"}}}
        # the step must actually have solved the linear system and reduced the residual
        if not (kits > 0 and last is not None and last[1] == last[1] and last[1] < 1e3 * l2_0):
            out["invalid"] = "Krylov status %r / residual %r after the step" % (kits, None if last is None else last[1])
        # CPU baseline on the box's host cores, N=1 only: ONE measured first Newton step of the oracle on the
        # sample box (not extrapolated), and the GPU timed on the SAME box right here, so that a same-config
        # pair exists
        if world == 1:
            try:
                m = args.cpu_sample_n
                ts, dof, cits, used_ref = cpu_reference_leg(m)
                sec = ts[0]
                out["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": 1, "kind": "reference" if used_ref else "port",
                                       "host_cores_available": os.cpu_count(), "n": m, "dofs": dof, "us_per_dof": 1e6 * sec / dof,
                                       "sample": "ONE measured first Newton step of the oracle on a %d^3 box (%d dof, V-cycle, %d Krylov its): %.2f s on 1 core "
                                                 "(serial like the reference); value = steps/s AT THAT SIZE, not extrapolated to %d^3%s" % (
                                                     m, dof, cits, sec, n, "; reference element kernels (oracle/_ref)" if used_ref else "")}
                g.close(); g = None

                def gpu_same_box(mm):
                    """GPU on the mm^3 box through the C ABI from host buffers: bsr.dat's V-cycle (what the CPU leg runs) and the bench's cycle"""
                    gs = pkg.PNP(box=(mm, mm, mm, 2.0, 2.0, 2.0), device=dev)
                    prs = problems.exact_solution_problem(mm, mm, mm, 2.0, 2.0, 2.0)
                    gs.set_coefficients(prs["eps"], prs["fixed"], prs["diff"], prs["val"], prs["reac"]); gs.set_dirichlet(prs["bc_dofs"])
                    itv, amgv = pkg.default_params(maxit=args.maxit, cycle_type=1)          # bsr.dat: V-cycle, like the CPU leg
                    same = {}
                    for nm, (i_, a_) in (("V", (itv, amgv)), (args.cycle, (it, amg))):
                        for _ in range(3):
                            gs.set_solution(prs["uu0"]); r_ = gs.newton_step(i_, a_)
                        t0 = time.perf_counter()
                        for _ in range(5):
                            gs.set_solution(prs["uu0"]); r_ = gs.newton_step(i_, a_)     # host buffers in, like e2e
                        same[nm] = {"gpu_steps_per_s": 5.0 / (time.perf_counter() - t0), "gpu_krylov_its": r_[0]}
                    gs.close()
                    return same
                same = gpu_same_box(m)
                out["cpu_same_n"] = {"n": m, "dofs": dof, "cpu_steps_per_s": 1.0 / sec, "cpu_krylov_its": cits, "cpu_cycle": "V",
                                     "gpu": same, "ratio_same_config_V": same["V"]["gpu_steps_per_s"] * sec,
                                     "note": "GPU: host clock around 5 x (pnpb200_set_solution from host memory + pnpb200_newton_step), same box, same bsr.dat V-cycle as the CPU leg"}
                # seconds per step and us per dof of the CPU path NEXT TO the GPU's at the same n, n in {32, 64}
                table = [{"n": m, "dofs": dof, "cpu_s_per_step": sec, "cpu_us_per_dof": 1e6 * sec / dof, "cpu_krylov_its": cits,
                          "gpu_s_per_step_V": 1.0 / same["V"]["gpu_steps_per_s"], "gpu_us_per_dof_V": 1e6 / same["V"]["gpu_steps_per_s"] / dof,
                          "gpu_krylov_its_V": same["V"]["gpu_krylov_its"]}]
                for m2 in (32,):
                    if m2 >= m:
                        continue
                    ts2, dof2, cits2, _ = cpu_reference_leg(m2)
                    same2 = gpu_same_box(m2)
                    table.insert(0, {"n": m2, "dofs": dof2, "cpu_s_per_step": ts2[0], "cpu_us_per_dof": 1e6 * ts2[0] / dof2, "cpu_krylov_its": cits2,
                                     "gpu_s_per_step_V": 1.0 / same2["V"]["gpu_steps_per_s"], "gpu_us_per_dof_V": 1e6 / same2["V"]["gpu_steps_per_s"] / dof2,
                                     "gpu_krylov_its_V": same2["V"]["gpu_krylov_its"]})
                out["cpu_same_n"]["table"] = table
            except Exception as e:  # the oracle is a checker; its absence must not hide the GPU number
                out.setdefault("cpu_baseline", {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "unavailable: %r" % (e,)})
        emit(out)
    if g is not None:
        g.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
