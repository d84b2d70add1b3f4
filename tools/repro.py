import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from pnpb200_loader import load_package
pkg=load_package()
from importlib import import_module
pr=import_module("modular_pnp_b200.problems").exact_solution_problem(12,12,12,2.,2.,2.)
g=pkg.PNP(box=(12,12,12,2.,2.,2.))
g.set_coefficients(pr["eps"],pr["fixed"],pr["diff"],pr["val"],pr["reac"]); g.set_dirichlet(pr["bc_dofs"]); g.set_solution(pr["uu0"])
it,amg=pkg.default_params(cycle_type=4)
for k in range(3):
    print(g.newton_step(it,amg), flush=True)
