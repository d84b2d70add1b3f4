"""2+ GPUs: the same Newton step repeated from the same state must give the same bits every time (iteration count,
residual norms) whichever halo path carries the exchanges.  torchrun ... tools/mg_determinism.py [n] [reps]"""
import os, sys, struct
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pnpb200_loader import load_package
from importlib import import_module
pkg = load_package()
part_mod = import_module("modular_pnp_b200.partition"); prob = import_module("modular_pnp_b200.problems")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
P = part_mod.box_brick_partition(n, n, n * world, 2.0, 2.0, 2.0 * world, rank, world, grid=(1, 1, world))
g = pkg.PNP(P["xyz"], P["cells"], device=rank)
uid = [pkg.capi.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
g.comm_init(uid[0], rank, world); g.set_partition(P)
if "color" in P:
    g.set_coloring(P["color"], 4)
f = prob.exact_solution_fields(P["xyz"][0::3])
g.set_coefficients(f["eps"], f["fixed"], f["diff"], f["val"], f["reac"])
g.set_dirichlet((3 * P["bc_vertices"][:, None] + np.arange(3)).ravel().astype(np.int32))
it, amg = pkg.default_params(cycle_type=int(os.environ.get("DET_CYCLE", "4")))
if os.environ.get("DET_SHORTCUTS") == "0":
    g.set_sweep_shortcuts(0)
out = []
for r in range(reps):
    g.set_solution(f["uu0"])
    st, l2, mx = g.newton_step(it, amg)
    out.append((st, struct.pack("<d", l2).hex(), struct.pack("<d", mx).hex()))
if rank == 0:
    print("halo path counts (peer kernel, nccl):", g.halo_counts())
    print("DISTINCT", len(set(out)), "of", len(out), [o[0] for o in out], sorted(set(o[1] for o in out)))
g.close()
dist.barrier(); dist.destroy_process_group()
