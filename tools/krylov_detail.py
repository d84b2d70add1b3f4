"""Krylov-stage breakdown of one Newton step (development aid): pnpb200_set_timing_detail brackets the
cycle, the matvec and the orthogonalisation of every iteration with CUDA events (adds host syncs)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from importlib import import_module
from pnpb200_loader import load_package

pkg = load_package()
problems = import_module("modular_pnp_b200.problems")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = pkg.PNP(box=(n, n, n, 2.0, 2.0, 2.0))
pr = problems.exact_solution_problem(n, n, n, 2.0, 2.0, 2.0)
g.set_coefficients(pr["eps"], pr["fixed"], pr["diff"], pr["val"], pr["reac"])
g.set_dirichlet(pr["bc_dofs"])
it, amg = pkg.default_params(cycle_type=4)
for rep in range(3):
    g.set_solution(pr["uu0"])
    g.L.pnpb200_set_timing_detail(g.h, 1 if rep == 2 else 0)
    st, l2, mx = g.newton_step(it, amg)
    s = g.stats()
    print("step %d: krylov its %d, assemble %.1f, setup %.1f, krylov %.1f (cycle %.1f, matvec %.1f, ortho %.1f), residual %.1f ms" % (
        rep, st, s.ms_assemble, s.ms_setup, s.ms_krylov, s.ms_vcycle, s.ms_spmv, s.ms_ortho, s.ms_residual), flush=True)
for lvl in range(0, 4):
    for w, nm in [(7, "gs0"), (8, "resU"), (9, "gsd"), (10, "Ldelta"), (0, "spmv")]:
        ms, by = g.probe_kernel(100 * lvl + w, 10)
        print("  level %d %-7s %.3f ms" % (lvl, nm, ms))
g.close()
