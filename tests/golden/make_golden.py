"""Generates tests/golden/element_kernels.npz from the REFERENCE's own compiled element kernels
(oracle/_ref/libpnpref.so, built by oracle/Makefile from /root/reference/include/EAFE.h and
benchmarks/pnp_exact_solution/vector_linear_pnp_forms.h). Run in the authoring container only
(needs /root/reference); the vectors travel with the repo.
    python tests/golden/make_golden.py [main] [diode] [marker] [newton] [ns]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    ref = C.CDLL(O.REF_PATH)
    f8 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ref.ref_eafe_tabulate.argtypes = [f8] * 6
    ref.ref_jac_tabulate.argtypes = [f8] * 6
    ref.ref_res_tabulate.argtypes = [f8] * 8
    rng = np.random.default_rng(1234)
    n = 64
    X = np.zeros((n, 12)); UU = np.zeros((n, 12)); EPS = np.zeros((n, 4)); D = np.zeros((n, 12)); Q = np.zeros((n, 12))
    FC = np.zeros((n, 4)); RE = np.zeros((n, 12)); ETA = np.zeros((n, 4)); AL = np.zeros((n, 4)); BE = np.zeros((n, 4)); GA = np.zeros((n, 4))
    A_eafe = np.zeros((n, 16)); A_jac = np.zeros((n, 144)); b_res = np.zeros((n, 12))
    base = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1.0])
    for t in range(n):
        X[t] = (base + 0.2 * rng.uniform(-1, 1, 12)) * rng.uniform(0.01, 2.0)
        UU[t] = rng.uniform(-1, 1, 12); EPS[t] = rng.uniform(0.5, 2, 4); D[t] = rng.uniform(0.1, 2, 12)
        Q[t] = rng.uniform(-2, 2, 12); FC[t] = rng.uniform(-1, 1, 4); RE[t] = rng.uniform(-1, 1, 12)
        ETA[t] = rng.uniform(-1, 1, 4); AL[t] = rng.uniform(0.1, 2, 4); GA[t] = rng.uniform(0, 1, 4)
        # potentials on a 1/64 lattice: every edge difference is 0 or >= 1/64 (well-conditioned Bernoulli)
        BE[t] = ETA[t] + np.round(rng.uniform(-3, 3, 4) * 64) / 64
        if t % 8 == 0:
            BE[t, 1] = BE[t, 0] + (ETA[t, 1] - ETA[t, 0])      # an exactly-zero edge difference: |d| < DOLFIN_EPS branch
        ref.ref_eafe_tabulate(A_eafe[t], ETA[t], AL[t], BE[t], GA[t], X[t])
        ref.ref_jac_tabulate(A_jac[t], UU[t], EPS[t], D[t], Q[t], X[t])
        ref.ref_res_tabulate(b_res[t], UU[t], EPS[t], FC[t], D[t], Q[t], RE[t], X[t])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "element_kernels.npz"), X=X, UU=UU, EPS=EPS, D=D, Q=Q, FC=FC,
                        RE=RE, ETA=ETA, AL=AL, BE=BE, GA=GA, A_eafe=A_eafe, A_jac=A_jac, b_res=b_res)
    # global fixture: the oracle assembler DRIVEN BY THE REFERENCE KERNELS on a 4x2x2 box
    O.use_ref_kernels(True)
    for grid in ((4, 2, 2), (8, 4, 4)):
        P = O.Problem(*grid)
        uu = P.exact + 0.1 * np.random.default_rng(1234).uniform(-1, 1, P.N)
        uu[0::3] = np.round(uu[0::3] * 64) / 64
        A, b = P.assemble(uu, use_eafe=True)
        A2, _ = P.assemble(uu, use_eafe=False)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "assembled_%dx%dx%d.npz" % grid), uu=uu, IA=P.IA, JA=P.JA,
                            A_eafe=A, A_galerkin=A2, b=b)
    O.use_ref_kernels(False)
    print("golden vectors written")


def diode():
    """golden vectors of the DIODE variant of the forms from the reference's own compiled kernels
    (oracle/_ref/libpnpref_diode.so <- benchmarks/pnp_diode/vector_linear_pnp_forms.h)"""
    ref = C.CDLL(O.REF_DIODE_PATH)
    f8 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ref.ref_jac_tabulate.argtypes = [f8] * 6
    ref.ref_res_tabulate.argtypes = [f8] * 8
    ref.ref_set_poisson_scale.argtypes = [C.c_double]
    rng = np.random.default_rng(4321)
    n = 48
    ps = 37.5
    ref.ref_set_poisson_scale(ps)
    X = np.zeros((n, 12)); UU = np.zeros((n, 12)); EPS = np.zeros((n, 4)); D = np.zeros((n, 12)); Q = np.zeros((n, 12))
    FC = np.zeros((n, 4)); RE = np.zeros((n, 12)); A_jac = np.zeros((n, 144)); b_res = np.zeros((n, 12))
    base = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1.0])
    for t in range(n):
        X[t] = (base + 0.2 * rng.uniform(-1, 1, 12)) * rng.uniform(0.01, 2.0)
        UU[t] = rng.uniform(-1, 1, 12); EPS[t] = rng.uniform(0.5, 2, 4); D[t] = rng.uniform(0.1, 2, 12)
        Q[t] = rng.uniform(-2, 2, 12); FC[t] = rng.uniform(-1, 1, 4); RE[t] = rng.uniform(-1, 1, 12)
        if t % 3 == 0:       # log-densities straddling the -50 cutoff of conc(w)
            UU[t, 4:8] = -50.0 + rng.uniform(-3, 3, 4)
        if t % 3 == 1:
            UU[t, 8:12] = -50.0 + rng.uniform(-1, 1, 4); UU[t, 0:4] = 20.0 * rng.uniform(-1, 1, 4)
        ref.ref_jac_tabulate(A_jac[t], UU[t], EPS[t], D[t], Q[t], X[t])
        ref.ref_res_tabulate(b_res[t], UU[t], EPS[t], FC[t], D[t], Q[t], RE[t], X[t])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "element_kernels_diode.npz"), poisson_scale=ps, X=X, UU=UU, EPS=EPS,
                        D=D, Q=Q, FC=FC, RE=RE, A_jac=A_jac, b_res=b_res)
    print("diode golden vectors written")


def marker():
    """golden vectors of the DG0 entropy indicator from the reference's own compiled kernels
    (oracle/_ref/libpnpref_marker.so <- include/poisson_cell_marker.h): L-form value with the entropy
    vector bound to zero as in src/mesh_refiner.cpp:250-251, and the DG0 mass entry"""
    ref = C.CDLL(os.path.join(os.path.dirname(O.REF_PATH), "libpnpref_marker.so"))
    f8 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ref.ref_marker_mass.argtypes = [f8, f8]
    ref.ref_marker_tabulate.argtypes = [f8] * 6
    rng = np.random.default_rng(2468)
    n = 64
    X = np.zeros((n, 12)); MU = np.zeros((n, 4)); LW = np.zeros((n, 4)); DF = np.zeros((n, 4))
    L = np.zeros((n, 1)); M = np.zeros((n, 1))
    zero12 = np.zeros(12)
    base = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1.0])
    for t in range(n):
        X[t] = (base + 0.2 * rng.uniform(-1, 1, 12)) * rng.uniform(0.01, 2.0)
        MU[t] = rng.uniform(-3, 3, 4); LW[t] = rng.uniform(-4, 2, 4); DF[t] = rng.uniform(0.1, 2, 4)
        ref.ref_marker_tabulate(L[t], MU[t], zero12, LW[t], DF[t], X[t])
        ref.ref_marker_mass(M[t], X[t])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "marker_kernels.npz"), X=X, MU=MU, LW=LW, DF=DF, L=L[:, 0], M=M[:, 0])
    print("marker golden vectors written")


def newton():
    """Newton history of BASELINE configs[0] (pnp_exact_solution, 50x10x10 box of [-1,1]x[-0.2,0.2]^2,
    bsr.dat, main.cpp:317-356) from the oracle DRIVEN BY THE REFERENCE'S OWN ELEMENT KERNELS."""
    import json
    O.use_ref_kernels(True)
    uu, res = O.newton_exact(50, 10, 10)
    O.use_ref_kernels(False)
    k = res.newton_its
    P = O.Problem(50, 10, 10)
    out = {"grid": [50, 10, 10], "lengths": [2.0, 0.4, 0.4], "newton_its": int(k), "converged": int(res.converged),
           "initial_residual": float(res.initial_residual), "initial_max_residual": float(res.initial_max_residual),
           "rel_res": [float(res.rel_res[i]) for i in range(k)], "max_res": [float(res.max_res[i]) for i in range(k)],
           "krylov_its": [int(res.krylov_its[i]) for i in range(k)],
           "max_nodal_error": float(np.abs(uu - P.exact).max()),
           "solution_checksum": [float(np.sum(uu[c::3])) for c in range(3)]}
    with open(os.path.join(ROOT, "tests", "golden", "newton_history_50x10x10.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("newton history written:", out["newton_its"], out["krylov_its"])


def ns():
    """golden vectors of the PNP-Stokes forms (P1^3 x RT1 x DG0) from the reference's own compiled kernels
    (oracle/_ref/libpnpref_ns.so <- benchmarks/pnp_stokes/vector_linear_pnp_ns_forms.h): element tensors of random
    cells / facet pairs, and the assembled mixed system of a 4x2x2 box driven by those kernels."""
    ref = C.CDLL(O.REF_NS_PATH)
    f8 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ref.ref_ns_cell_a.argtypes = [f8] * 5
    ref.ref_ns_cell_L.argtypes = [f8] * 6
    ref.ref_ns_facet_a.argtypes = [f8] * 6 + [C.c_int] * 2
    ref.ref_ns_facet_L.argtypes = [f8] * 7 + [C.c_int] * 2
    rng = np.random.default_rng(1357)
    n = 32
    X0 = np.zeros((n, 12)); X1 = np.zeros((n, 12)); CC = np.zeros((n, 24)); UU = np.zeros((n, 8)); PP = np.zeros((n, 2))
    PAR = np.zeros((n, 8)); F = np.zeros((n, 2), dtype=np.int32)
    A_cell = np.zeros((n, 289)); b_cell = np.zeros((n, 17)); A_facet = np.zeros((n, 1156)); b_facet = np.zeros((n, 34))
    base = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1.0])
    for t in range(n):
        x0 = ((base + 0.2 * rng.uniform(-1, 1, 12)) * rng.uniform(0.05, 2.0)).reshape(4, 3)
        if t % 2:
            x0[[1, 2]] = x0[[2, 1]]                      # negatively oriented cells too (detJ < 0)
        f0, f1 = int(rng.integers(4)), int(rng.integers(4))
        face = [i for i in range(4) if i != f0]
        x1 = np.zeros((4, 3))
        x1[[i for i in range(4) if i != f1]] = x0[face]  # same vertex order on both sides (UFC ordering)
        x1[f1] = 2.0 * x0[face].mean(0) - x0[f0] + 0.1 * rng.uniform(-1, 1, 3) * np.abs(x0).max()
        X0[t] = x0.reshape(-1); X1[t] = x1.reshape(-1); F[t] = (f0, f1)
        CC[t] = rng.uniform(-1.5, 1.5, 24); UU[t] = rng.uniform(-1, 1, 8); PP[t] = rng.uniform(-1, 1, 2)
        PAR[t] = O.NS_PAR_DEFAULT * rng.uniform(0.5, 2.0, 8)
        ref.ref_ns_cell_a(A_cell[t], CC[t, :12].copy(), UU[t, :4].copy(), PAR[t], X0[t])
        ref.ref_ns_cell_L(b_cell[t], CC[t, :12].copy(), UU[t, :4].copy(), PP[t, :1].copy(), PAR[t], X0[t])
        ref.ref_ns_facet_a(A_facet[t], CC[t], UU[t], PAR[t], X0[t], X1[t], f0, f1)
        ref.ref_ns_facet_L(b_facet[t], CC[t], UU[t], PP[t], PAR[t], X0[t], X1[t], f0, f1)
    P = O.Problem(4, 2, 2, lx=2.0, ly=1.0, lz=1.0)
    cs, cf, fc = O.ns_topology(P.cells)
    nv, nc, nf = P.xyz.size // 3, cs.shape[0], fc.shape[0]
    cc = rng.uniform(-1, 1, 3 * nv); uu = rng.uniform(-1, 1, nf); pp = rng.uniform(-1, 1, nc)
    bc = np.flatnonzero(np.repeat(np.abs(np.abs(P.xyz.reshape(-1, 3)[:, 0]) - 1.0) < 1e-12, 3)).astype(np.int32)
    A, b = O.ns_assemble(P.xyz, cs, cf, fc, cc, uu, pp, bc=bc, use_ref=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ns_kernels.npz"), X0=X0, X1=X1, F=F, CC=CC, UU=UU, PP=PP, PAR=PAR,
                        A_cell=A_cell, b_cell=b_cell, A_facet=A_facet, b_facet=b_facet,
                        box=np.array([4, 2, 2]), lengths=np.array([2.0, 1.0, 1.0]), cc=cc, uu=uu, pp=pp, bc=bc,
                        cells_sorted=cs, cell_facets=cf, facet_cells=fc,
                        IA=A.indptr.astype(np.int32), JA=A.indices.astype(np.int32), A=A.data, b=b)
    print("PNP-Stokes golden vectors written")


if __name__ == "__main__":
    which = sys.argv[1:] or ["main", "diode", "marker"]
    for w in which:
        {"main": main, "diode": diode, "marker": marker, "newton": newton, "ns": ns}[w]()
