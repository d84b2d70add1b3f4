"""Per-kernel device time of ONE resident Newton step (development aid, 1 GPU): CUPTI activity records through
torch.profiler — no replay, no serialisation, so the totals add up to the step time. A cross-check of the
ncu launch lists (which are cold-cache and serialised).   python tools/kineto_step.py 256 > table.txt"""
import os, sys, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from importlib import import_module
from pnpb200_loader import load_package
import torch
from torch.profiler import profile, ProfilerActivity

pkg = load_package()
problems = import_module("modular_pnp_b200.problems")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = pkg.PNP(box=(n, n, n, 2.0, 2.0, 2.0))
pr = problems.exact_solution_problem(n, n, n, 2.0, 2.0, 2.0)
g.set_coefficients(pr["eps"], pr["fixed"], pr["diff"], pr["val"], pr["reac"]); g.set_dirichlet(pr["bc_dofs"])
it, amg = pkg.default_params(cycle_type=4)
for _ in range(2):
    g.set_solution(pr["uu0"]); r = g.newton_step(it, amg)
torch.cuda.synchronize()
g.set_solution(pr["uu0"])
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    r = g.newton_step(it, amg)
    torch.cuda.synchronize()
s = g.stats()
tot = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        nm = ev.name
        tot[nm][0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        tot[nm][1] += 1
allus = sum(v[0] for v in tot.values())
print("n=%d krylov its %d; stage ms: assemble %.1f setup %.1f krylov %.1f residual %.1f; kernels+memops %.1f ms in %d records" % (
    n, r[0], s.ms_assemble, s.ms_setup, s.ms_krylov, s.ms_residual, allus / 1e3, sum(v[1] for v in tot.values())))
print("%-110s %8s %7s %6s" % ("kernel", "ms", "calls", "share"))
for nm, (us, cnt) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:60]:
    print("%-110s %8.2f %7d %5.1f%%" % (nm[:110], us / 1e3, cnt, 100.0 * us / allus))
g.close()
