"""Bitwise reproducibility of the fine-level matrix passes (development aid, 1 GPU).  N=64 REPS=300 python tools/repeat_pass.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnpb200_loader import load_package
from importlib import import_module
pkg = load_package()
prob = import_module("modular_pnp_b200.problems")
n = int(os.environ.get("N", "64")); reps = int(os.environ.get("REPS", "200"))
g = pkg.PNP(box=(n, n, n, 2.0, 2.0, 2.0))
f = prob.exact_solution_problem(n, n, n, 2.0, 2.0, 2.0)
g.set_coefficients(f["eps"], f["fixed"], f["diff"], f["val"], f["reac"])
g.set_dirichlet(f["bc_dofs"])
it, amg = pkg.default_params(cycle_type=4)
g.set_solution(f["uu0"])
print("newton", g.newton_step(it, amg), flush=True)
names = {12: "GS back", 0: "spmv", 1: "residual", 2: "GS fwd", 7: "GS0 fwd from zero", 8: "RESU", 9: "GSD back+delta", 10: "LDELTA"}
for which in [int(w) for w in os.environ.get("WHICH", "9,12,2,7").split(",")]:
    out = g.debug_repeat_pass(which, reps)
    print("pass %2d %-18s bad reps %d / %d; first: rep %d, %d entries, positions %s" %
          (which, names[which], out[0], reps, out[17], out[1], list(out[2:2 + min(15, out[1])])), flush=True)
g.close()
