"""Per-phase wall time of the AMG setup of one Newton step (development aid): AMG print_level 3 makes
amg.cu bracket every setup phase with a stream synchronise + host clock."""
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
for rep in range(2):
    g.set_solution(pr["uu0"])
    if rep == 1:
        it, amg = pkg.default_params(cycle_type=4, print_level=0)
        amg.print_level = 3
    st, l2, mx = g.newton_step(it, amg)
    s = g.stats()
    print("step %d: krylov %d, assemble %.1f ms, setup %.1f ms, krylov %.1f ms, residual %.1f ms" % (rep, st, s.ms_assemble, s.ms_setup, s.ms_krylov, s.ms_residual), flush=True)
g.close()
