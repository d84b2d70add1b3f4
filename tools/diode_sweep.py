"""BASELINE configs[1] end to end on one GPU: the I-V sweep of benchmarks/pnp_diode/main.cpp:96-189 on the C ABI.

For every bias of the sweep (-0.3 ... +0.5 V, step 0.1): the adaptivity loop of main.cpp:139-189 — damped Newton
(pnp_newton_solver.h:195-284; device-resident fast path) with the block ILU(2)-FGMRES of linear_pnp.cpp:103-109, the
entropy indicator and the marking rules of Mesh_Refiner (src/mesh_refiner.cpp:73-209: tolerance 1e-2 per cell, growth
factor 1.05 through mark_for_refinement_with_target_size) on the device, refinement of the marked cells by the caller
(here a conforming centroid split standing in for dolfin::adapt), a NEW handle on the refined mesh, the P1-interpolated
solution as the warm start (adapt(Function, mesh), main.cpp:184). Prints one table row per (bias, mesh level).

    python tools/diode_sweep.py [--levels 3] [--precond ilu|amg] > profiles/r02_diode_sweep.txt
"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from importlib import import_module
from pnpb200_loader import load_package

pkg = load_package()
pr = import_module("modular_pnp_b200.problems")


class Backend:
    def __init__(self, g, it, amg):
        self.g, self.it, self.amg, self.n_dofs = g, it, amg, g.N
        self.krylov = []

    def set_solution(self, uu): self.g.set_solution(np.ascontiguousarray(uu, dtype=np.float64))
    def residual(self): return self.g.residual()
    def solve(self):
        st = self.g.solve_update(self.it, self.amg); self.krylov.append(st); return st
    def ranges(self): return self.g.update_ranges()
    def trial(self, alpha): return self.g.trial_residual(alpha)


def kuhn_box(nx, ny, nz, lx, ly, lz):
    g = pkg.PNP(box=(nx, ny, nz, lx, ly, lz))
    xyz, cells = g.get_mesh(); g.close()
    return xyz.reshape(-1, 3).copy(), cells.reshape(-1, 4).astype(np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=3, help="mesh levels per bias (the benchmark stops by size: 250 000 cells)")
    ap.add_argument("--precond", default="ilu", choices=["ilu", "amg"])
    ap.add_argument("--max-newton", type=int, default=250)
    ap.add_argument("--volts", default="-0.3,-0.2,-0.1,0.0,0.1,0.2,0.3,0.4,0.5")
    args = ap.parse_args()
    X0, C0 = kuhn_box(40, 4, 4, 2.0, 0.2, 0.2)                    # benchmarks/pnp_diode/domain.dat:14-20
    it, amg = pkg.default_params(maxit=500)                       # pnp_diode/bsr.dat: itsolver_maxit 500
    print("# pnp_diode I-V sweep on the device: damped Newton (max %d, rel 1e-7, max-res 1e-10), %s, entropy tolerance 1e-2 per cell, growth 1.05"
          % (args.max_newton, "block ILU(2)-FGMRES" if args.precond == "ilu" else "BSR UA-AMG-FGMRES"))
    print("%6s %5s %8s %8s %7s %9s %10s %10s %11s %11s %8s %7s" % ("volts", "level", "vertices", "cells", "newton", "converged", "backtracks", "krylov/it",
                                                                  "rel resid", "max resid", "marked", "sec"))
    for v in [float(t) for t in args.volts.split(",")]:
        X, C, uu = X0.copy(), C0.copy(), None
        for level in range(args.levels):
            t0 = time.time()
            cells = np.sort(C, axis=1).astype(np.int32)
            f = pr.diode_fields(X[:, 0], v)
            g = pkg.PNP(X.ravel(), cells.ravel(), 0)
            g.set_coefficients(f["eps"], f["fixed"], f["diff"], f["val"], f["reac"])
            lo, hi = X[:, 0].min(), X[:, 0].max()
            bcv = np.nonzero((X[:, 0] < lo + 1e-10) | (X[:, 0] > hi - 1e-10))[0]
            g.set_dirichlet((3 * bcv[:, None] + np.arange(3)[None, :]).ravel().astype(np.int32))
            g.set_form_variant(pr.DIODE_POISSON_SCALE, True)
            if args.precond == "ilu":
                g.set_preconditioner("ilu", 2)
            start = f["uu0"] if uu is None else uu
            if uu is not None:                                    # the warm start carries the exact contact values
                s2 = start.reshape(-1, 3).copy(); s2[bcv] = f["uu0"].reshape(-1, 3)[bcv]; start = s2.ravel()
            be = Backend(g, it, amg)
            hist, conv = pr.damped_newton(be, start, max_newton=args.max_newton)
            sol = g.get_solution().reshape(-1, 3)
            # Mesh_Refiner::multilevel_refinement with max_elements = floor(1.05 * cells) (main.cpp:178-179)
            nc = len(C)
            g.entropy_indicator()
            marked, cnt, tol = g.mark_cells(1.0e-2)
            if nc + 3 * cnt > int(np.floor(1.05 * nc)):           # too aggressive: proportional marking, target shrunk by 0.95 per try (:133-147)
                target = int(np.floor(1.05 * nc))
                while True:
                    target = int(round(0.95 * target))
                    marked, cnt, tol = g.mark_cells(1.0e-2, target_size=max(target, nc + 1))
                    if nc + 3 * cnt <= int(np.floor(1.05 * nc)) or target <= nc:
                        break
            g.close()
            kr = [k for k in be.krylov if k > 0]
            print("%6.2f %5d %8d %8d %7d %9s %10d %10.1f %11.3e %11.3e %8d %7.2f" % (
                v, level, len(X), nc, len(hist), conv, sum(h[4] for h in hist), (sum(kr) / len(kr)) if kr else float("nan"),
                hist[-1][2] if hist else float("nan"), hist[-1][3] if hist else float("nan"), cnt, time.time() - t0), flush=True)
            if cnt == 0:
                break
            split, keep = C[marked], C[~marked]
            new_ids = len(X) + np.arange(len(split))
            X = np.vstack([X, X[split].mean(axis=1)])
            uu = np.vstack([sol, sol[split].mean(axis=1)]).ravel()
            subs = []
            for drop in range(4):
                t = split.copy(); t[:, drop] = new_ids; subs.append(t)
            C = np.vstack([keep] + subs)


if __name__ == "__main__":
    main()
