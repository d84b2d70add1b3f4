"""Row-kernel probes on one GPU (development aid): builds the n^3 box problem, runs one Newton
step (hierarchy + block inverses), then times the requested kernel probes. Under
`ncu --profile-from-start off` only the probes are captured (cudaProfilerStart/Stop)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from importlib import import_module
from pnpb200_loader import load_package

pkg = load_package()
problems = import_module("modular_pnp_b200.problems")


def main():
    n = int(sys.argv[1])
    probes = [int(a) for a in sys.argv[2].split(",")]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    g = pkg.PNP(box=(n, n, n, 2.0, 2.0, 2.0))
    pr = problems.exact_solution_problem(n, n, n, 2.0, 2.0, 2.0)
    g.set_coefficients(pr["eps"], pr["fixed"], pr["diff"], pr["val"], pr["reac"])
    g.set_dirichlet(pr["bc_dofs"])
    g.set_solution(pr["uu0"])
    it, amg = pkg.default_params(cycle_type=4)
    g.residual()
    st, l2, mx = g.newton_step(it, amg)
    print("n=%d krylov %d" % (n, st), flush=True)
    import torch
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    names = {0: "spmv", 1: "residual", 2: "gs_sweep", 7: "gs0_fwd_L", 8: "res_U", 9: "gsd_back", 10: "apply_Ldelta"}
    for w in probes:
        ms, by = g.probe_kernel(w, reps)
        print("  probe %-12s %.3f ms  %.1f GB/s (algorithmic %.2f GB)" % (names.get(w, str(w)), ms, by / ms / 1e6, by / 1e9), flush=True)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    g.close()


if __name__ == "__main__":
    main()
