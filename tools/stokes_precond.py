"""PNP-Stokes block solver on the device-assembled RT1 x DG0 system (first Newton step of the pnp_stokes problem):
iteration counts of the outer / PNP / Stokes / velocity solves by mesh size, with the velocity block solved by an inner
AMG-FGMRES (ns.dat's _v set, default) and by ONE AMG cycle (PNPB200_STOKES_VEL_KRYLOV=0).
    python tools/stokes_precond.py [nx,ny,nz ...]   -> profiles/r02_stokes_precond.txt"""
import os, sys, time
from importlib import import_module
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pnpb200_loader import load_package
pkg = load_package()
problems = import_module("modular_pnp_b200.problems")
import scipy.sparse as sp

grids = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(10, 2, 2), (20, 3, 3), (30, 5, 5), (50, 5, 5), (60, 10, 10)]
print("grid            dofs   mode      outer  pnp_its  stokes_its  vel_solves  vel_its  relres     seconds")
for grid in grids:
    lengths = (10.0, 1.0, 1.0)
    h = pkg.capi.PNP(box=tuple(grid) + tuple(lengths))
    xyz, cells = h.get_mesh(); h.close()
    S = pkg.capi.PNPStokes(xyz, cells)
    cs, cf, fc = S.topology()
    pr = problems.pnp_stokes_problem(xyz, cs, cf, fc, lengths[0])
    S.set_parameters(pr["par"]); S.set_dirichlet(pr["bc_dofs"]); S.set_solution(pr["cc"], pr["uu"], pr["pp"])
    S.assemble()
    IA, JA, val, b = S.get_system()
    A = sp.csr_matrix((val, JA, IA), shape=(S.ndof, S.ndof))
    npnp = 3 * S.nv
    blocks = [A[:npnp, :npnp].tocsr(), A[:npnp, npnp:].tocsr(), A[npnp:, :npnp].tocsr(), A[npnp:, npnp:].tocsr()]
    for B in blocks:
        B.sort_indices()
    for mode in ("1", "0"):
        os.environ["PNPB200_STOKES_VEL_KRYLOV"] = mode
        t = time.time()
        st, xs, stats = pkg.capi.solve_pnp_stokes(blocks, b, S.nf, S.nc)
        dt = time.time() - t
        rr = np.linalg.norm(b - A @ xs) / np.linalg.norm(b)
        print("%-14s %7d  %-8s %5d %8d %11d %11d %8d  %.2e  %7.2f" % ("%dx%dx%d" % grid, S.ndof, "krylov_v" if mode == "1" else "cycle_v", st,
              stats["pnp_its"], stats["stokes_its"], stats["velocity_solves"], stats["velocity_its"], rr, dt), flush=True)
    S.close()
