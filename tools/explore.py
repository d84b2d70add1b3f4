"""Scaling exploration on one GPU (development aid, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pnpb200_loader import load_package
pkg = load_package()
from importlib import import_module
problems = import_module("modular_pnp_b200.problems")

def run(n, steps=2, lx=2.0, ly=2.0, lz=2.0, reuse=False, **kw):
    t0 = time.time()
    g = pkg.PNP(box=(n, n, n, lx, ly, lz))
    t1 = time.time()
    pr = problems.exact_solution_problem(n, n, n, lx, ly, lz)
    g.set_coefficients(pr["eps"], pr["fixed"], pr["diff"], pr["val"], pr["reac"])
    g.set_dirichlet(pr["bc_dofs"])
    g.set_solution(pr["uu0"])
    t2 = time.time()
    it, amg = pkg.default_params(**kw)
    g.set_timing_detail(True)
    g.set_hierarchy_reuse(reuse)
    l2_0, mx0 = g.residual()
    print("n=%d dof=%d create %.2fs coef %.2fs  r0 %.4e" % (n, g.N, t1 - t0, t2 - t1, l2_0), flush=True)
    for k in range(steps):
        t = time.time()
        st, l2, mx = g.newton_step(it, amg)
        dt = time.time() - t
        s = g.stats()
        print("  step %d: krylov %d relres %.3e wall %.3fs | asm %.1f setup %.1f krylov %.1f res %.1f ms | vcycle %.1f spmv %.1f ortho %.1f | launches %d"
              % (k + 1, st, l2 / l2_0, dt, s.ms_assemble, s.ms_setup, s.ms_krylov, s.ms_residual, s.ms_vcycle, s.ms_spmv, s.ms_ortho, s.launches), flush=True)
    s = g.stats()
    print("  levels:", [(s.level_rows[l], s.level_nnzb[l], s.level_colors[l]) for l in range(s.n_levels)])
    for w, name in [(0, "spmv"), (1, "residual"), (2, "gs_sweep"), (3, "multidot16"), (4, "res_asm"), (5, "jac_asm"),
                    (7, "gs0_fwd_L"), (8, "res_U"), (9, "gsd_back"), (10, "apply_Ldelta")]:
        ms, by = g.probe_kernel(w, 5)
        print("  probe %-10s %.3f ms  %.1f GB/s (algorithmic %.2f GB)" % (name, ms, by / ms / 1e6, by / 1e9))
    for lvl in range(1, s.n_levels - 1):
        for w, name in [(0, "spmv"), (2, "gs_sweep"), (7, "gs0_fwd_L"), (8, "res_U"), (9, "gsd_back"), (10, "apply_Ldelta")]:
            ms, by = g.probe_kernel(100 * lvl + w, 20)
            print("  L%d probe %-10s %.3f ms  %.1f GB/s (algorithmic %.3f GB)" % (lvl, name, ms, by / ms / 1e6, by / 1e9))
    import torch
    free, tot = torch.cuda.mem_get_info()
    print("  mem used %.1f GB" % ((tot - free) / 2**30))
    g.close()

if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        if "=" in a:
            k, v = a.split("="); kw[k] = float(v) if "." in v or "e" in v else int(v)
    for a in sys.argv[1:]:
        if "=" not in a:
            run(int(a), **kw)
