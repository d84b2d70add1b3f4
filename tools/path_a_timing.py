"""Path A of INTEGRATION.md timed (development aid, 1 GPU): fasp_solver_dbsr_krylov_amg called with HOST structs, as
benchmarks/pnp_exact_solution/linear_pnp.cpp:101 does once per Newton step. First call = upload + colouring + layout +
solve; following calls bring new values on the same pattern.   python tools/path_a_timing.py 64 [96 ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from importlib import import_module
from pnpb200_loader import load_package
pkg = load_package()
problems = import_module("modular_pnp_b200.problems")
for n in [int(a) for a in sys.argv[1:]] or [64]:
    g = pkg.PNP(box=(n, n, n, 2.0, 2.0, 2.0))
    pr = problems.exact_solution_problem(n, n, n, 2.0, 2.0, 2.0)
    g.set_coefficients(pr["eps"], pr["fixed"], pr["diff"], pr["val"], pr["reac"]); g.set_dirichlet(pr["bc_dofs"])
    g.set_solution(pr["uu0"]); g.assemble()
    IA, JA, val = g.get_matrix(); b = g.get_rhs()
    it, amg = pkg.default_params(cycle_type=1)                 # bsr.dat as the reference runs it
    t0 = time.perf_counter(); st_fused = g.solve(it, amg); t_fused = time.perf_counter() - t0
    g.close()
    pkg.lib().pnpb200_release_host_cache()
    ts = []
    for k in range(5):
        t0 = time.perf_counter()
        x, st = pkg.capi.fasp_solver_dbsr_krylov_amg(IA, JA, val, b, it, amg)
        ts.append(time.perf_counter() - t0)
    print("n=%d (%d dof, %.2f GB of values): device-resident solve %.1f ms (%d its); host-struct entry: first call %.1f ms, "
          "following calls %s ms (%d its)" % (n, 3 * (n + 1) ** 3, val.nbytes / 1e9, 1e3 * t_fused, st_fused, 1e3 * ts[0],
                                               ", ".join("%.1f" % (1e3 * t) for t in ts[1:]), st), flush=True)
    pkg.lib().pnpb200_release_host_cache()
