"""Which stage of a partitioned Newton step is not reproducible run to run?  (development aid, 2+ GPUs under torchrun)

Repeats the SAME first Newton step and hashes, per repetition and per rank: the assembled matrix / rhs, every
level's aggregates, sweep order and matrix (columns sorted inside a row, so storage order does not count), one
cycle applied to a fixed vector, and the step's Krylov count / residual. Prints the first stage whose hash set has
more than one member.   N=48 REPS=100 torchrun --nproc-per-node 2 tools/det_stages.py
"""
import os, sys, hashlib
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnpb200_loader import load_package
from importlib import import_module
pkg = load_package()
part_mod = import_module("modular_pnp_b200.partition"); prob = import_module("modular_pnp_b200.problems")
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n = int(os.environ.get("N", "48")); reps = int(os.environ.get("REPS", "60"))
cyc = int(os.environ.get("CYCLE", "4"))
P = part_mod.box_brick_partition(n, n, n * world, 2.0, 2.0, 2.0 * world, rank, world, grid=(1, 1, world))
g = pkg.PNP(P["xyz"], P["cells"], device=rank)
if world > 1:
    uid = [pkg.capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g.comm_init(uid[0], rank, world); g.set_partition(P)
    if os.environ.get("COLOR", "1") == "1":
        g.set_coloring(P["color"], 4)
f = prob.exact_solution_fields(P["xyz"][0::3])
g.set_coefficients(f["eps"], f["fixed"], f["diff"], f["val"], f["reac"])
g.set_dirichlet((3 * P["bc_vertices"][:, None] + np.arange(3)).ravel().astype(np.int32))
it, amg = pkg.default_params(cycle_type=cyc)
rng = np.random.default_rng(7 + rank)
rvec = rng.standard_normal(g.N)
light = os.environ.get("LIGHT", "0") == "1"
if os.environ.get("SHORTCUTS", "1") == "0":
    g.set_sweep_shortcuts(False)


def hx(*arrs):
    m = hashlib.sha1()
    for a in arrs:
        m.update(np.ascontiguousarray(a).tobytes())
    return m.hexdigest()[:10]


def canon(IA, JA, val):
    val = val.reshape(-1, 9)
    row = np.repeat(np.arange(len(IA) - 1), np.diff(IA))
    o = np.lexsort((JA, row))
    return hx(IA, JA[o], val[o]), hx(JA)


seen = {}
firsts = {}
for rep in range(reps):
    g.set_solution(f["uu0"])
    rec = {}
    if not light:
        g.assemble()
        IA, JA, val = g.get_matrix(); b = g.get_rhs()
        rec["asm"] = hx(IA, JA, val, b)
    st, l2, mx = g.newton_step(it, amg)
    s = g.stats()
    rec["its"] = "%d" % st
    rec["l2"] = "%.17e" % l2
    for l in range(1, 0 if light else s.n_levels):
        try:
            a, na = g.level_aggregates(l - 1)
            rec["agg%d" % (l - 1)] = hx(a)
        except Exception as e:
            rec["agg%d" % (l - 1)] = "n/a"
        try:
            rec["ord%d" % l] = hx(g.level_order(l))
        except Exception:
            rec["ord%d" % l] = "n/a"
        try:
            m = g.level_matrix(l)
            c, j = canon(*m)
            rec["A%d" % l] = c; rec["A%d_colorder" % l] = j
        except Exception:
            rec["A%d" % l] = "n/a"
    if not light:
        try:
            rec["cycle"] = hx(g.apply_vcycle(rvec))
        except Exception as e:
            rec["cycle"] = "n/a"
    for k, v in rec.items():
        seen.setdefault(k, {}).setdefault(v, []).append(rep)
out = [None] * world
if world > 1:
    dist.all_gather_object(out, seen)
else:
    out = [seen]
if rank == 0:
    for r, d in enumerate(out):
        for k, vals in d.items():
            if len(vals) > 1:
                desc = ", ".join("%s:%s" % (v, (reps_ if len(reps_) < 6 else "%d reps" % len(reps_))) for v, reps_ in vals.items())
                print("RANK%d %-14s %d distinct: %s" % (r, k, len(vals), desc), flush=True)
            else:
                print("RANK%d %-14s stable" % (r, k), flush=True)
g.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
