"""Iteration counts / timings of the PNP-Stokes block solver on the synthetic block system (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import scipy.sparse as sp
from importlib import import_module
from pnpb200_loader import load_package
pkg = load_package()
problems = import_module("modular_pnp_b200.problems")
from helpers import bsr_to_csr

VCYC = int(os.environ.get("BLOCK_VEL_CYCLE", "1"))      # AMG_cycle_type_v: 1 = V (ns.dat), 4 = NA (K-cycle)
for a in sys.argv[1:]:
    n = int(a)
    pr = problems.exact_solution_problem(n, n, n, 2.0, 2.0, 2.0)
    g = pkg.PNP(box=(n, n, n, 2.0, 2.0, 2.0))
    g.set_coefficients(pr["eps"], pr["fixed"], pr["diff"], pr["val"], pr["reac"]); g.set_dirichlet(pr["bc_dofs"]); g.set_solution(pr["uu0"])
    g.assemble()
    IA, JA, val = g.get_matrix(); g.close()
    sIA, sJA, sval = bsr_to_csr(IA, JA, val)
    App = sp.csr_matrix((sval, sJA, sIA), shape=(3 * (n + 1) ** 3,) * 2); App.eliminate_zeros()
    for coupling in (0.05, 0.5):
        S = problems.pnp_stokes_block_system(App, n, coupling=coupling)
        t0 = time.time()
        params = list(pkg.capi.pnp_stokes_params())
        params[4].param_v.cycle_type = VCYC
        st, x, stats = pkg.capi.solve_pnp_stokes(S["blocks"], S["b"], S["n_vel"], S["n_pres"], params)
        dt = time.time() - t0
        rel = np.linalg.norm(S["b"] - S["A"] @ x) / np.linalg.norm(S["b"])
        print("n=%d coupling=%.2f dofs pnp %d vel %d pres %d | status %d %s | true relres %.2e | %.2f s" % (
            n, coupling, App.shape[0], S["n_vel"], S["n_pres"], st, stats, rel, dt), flush=True)
