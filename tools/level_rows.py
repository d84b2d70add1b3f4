"""Row-length histogram of the coarse levels (development aid): python tools/level_rows.py 128"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from importlib import import_module
from pnpb200_loader import load_package
pkg = load_package()
problems = import_module("modular_pnp_b200.problems")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = pkg.PNP(box=(n, n, n, 2.0, 2.0, 2.0))
pr = problems.exact_solution_problem(n, n, n, 2.0, 2.0, 2.0)
g.set_coefficients(pr["eps"], pr["fixed"], pr["diff"], pr["val"], pr["reac"]); g.set_dirichlet(pr["bc_dofs"])
it, amg = pkg.default_params(cycle_type=4)
g.set_solution(pr["uu0"]); print("newton", g.newton_step(it, amg))
s = g.stats()
for l in range(1, min(4, s.n_levels)):
    IA, JA, val = g.level_matrix(l)
    ln = np.diff(IA)
    h = np.bincount(ln)
    cum = np.cumsum(h) / len(ln)
    print("level %d: rows %d, blocks %d, mean %.2f, max %d; share of rows with <= 16: %.3f, <= 24: %.3f, <= 32: %.3f; blocks in rows > 16: %.3f, > 32: %.3f" % (
        l, len(ln), ln.sum(), ln.mean(), ln.max(), cum[min(16, len(cum) - 1)], cum[min(24, len(cum) - 1)], cum[min(32, len(cum) - 1)],
        ln[ln > 16].sum() / ln.sum(), ln[ln > 32].sum() / ln.sum()))
    print("   histogram:", {int(k): int(v) for k, v in enumerate(h) if v})
g.close()
