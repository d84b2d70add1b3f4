"""-m "not gpu": host side of the PNP-Stokes block path: the three-in-one ns.dat
reader (benchmarks/pnp_stokes/main.cpp:85-86), fasp_dcsr_getblk (linear_pnp_ns.cpp:150-157)
and the synthetic block system the GPU tests solve."""
import ctypes as C
from importlib import import_module

import numpy as np


def test_ns_param_input_reads_the_three_parameter_sets(pkg):
    it, it_pnp, amg_pnp, it_ns, amg_ns = pkg.capi.pnp_stokes_params()
    # bcsr.dat: outer vFGMRES(10), block upper triangular (33)
    assert (it.itsolver_type, it.precond_type, it.restart, it.maxit) == (6, 33, 10, 100) and abs(it.tol - 1e-6) < 1e-20
    # ns.dat: plain keys -> Stokes iteration; _v / _p -> velocity / pressure blocks
    assert (it_ns.itsolver_type, it_ns.precond_type, it_ns.maxit, it_ns.restart) == (5, 10, 100, 100)
    assert (it_ns.itsolver_type_v, it_ns.precond_type_v, it_ns.pre_maxit_v, it_ns.pre_restart_v) == (6, 2, 100, 100)
    assert (it_ns.itsolver_type_p, it_ns.precond_type_p) == (6, 2) and abs(it_ns.pre_tol_p - 1e-6) < 1e-20
    v, p = amg_ns.param_v, amg_ns.param_p
    assert (v.AMG_type, v.cycle_type, v.coarse_dof, v.aggregation_type, v.max_aggregation, v.smooth_order) == (3, 1, 100, 1, 20, 1)
    assert (p.AMG_type, p.cycle_type, p.coarse_dof, p.aggregation_type, p.max_aggregation, p.smooth_order) == (3, 1, 50, 1, 100, 0)
    assert abs(v.strong_coupled - 0.08) < 1e-15 and abs(p.strong_coupled - 0.04) < 1e-15
    # bsr.dat (PNP block) is untouched by the suffix handling
    assert (it_pnp.restart, amg_pnp.coarse_dof) == (30, 100)


def test_getblk_extracts_blocks_in_parent_entry_order(pkg):
    import scipy.sparse as sp
    L = pkg.lib()
    rng = np.random.default_rng(5)
    A = sp.random(40, 40, density=0.2, random_state=1, format="csr") + sp.identity(40, format="csr")
    A.sort_indices()
    parent = pkg.capi.numpy_to_csr(A.indptr, A.indices, A.data, 40)
    Is = np.ascontiguousarray(rng.permutation(40)[:17], dtype=np.int32)
    Js = np.ascontiguousarray(rng.permutation(40)[:23], dtype=np.int32)
    m = pkg.capi.dCSRmat()
    assert L.fasp_dcsr_getblk(C.byref(parent), Is, Js, len(Is), len(Js), C.byref(m)) == 0
    assert (m.row, m.col) == (17, 23)
    ia = np.ctypeslib.as_array(m.IA, shape=(18,)).copy()
    ja = np.ctypeslib.as_array(m.JA, shape=(m.nnz,)).copy()
    va = np.ctypeslib.as_array(m.val, shape=(m.nnz,)).copy()
    L.fasp_dcsr_free(C.byref(m))
    B = sp.csr_matrix((va, ja, ia), shape=(17, 23)).toarray()
    assert np.array_equal(B, A.toarray()[np.ix_(Is, Js)])


def test_synthetic_pnp_stokes_system_is_a_saddle_point_problem(oracle):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from helpers import bsr_to_csr
    problems = import_module("modular_pnp_b200.problems")
    n = 5
    Ass, nu, nq = problems.mac_stokes_block(n)
    assert (nu, nq) == (3 * (n + 1) * n * n, n ** 3) and Ass.shape == (nu + nq, nu + nq)
    assert abs(Ass - Ass.T).max() == 0.0
    Auu, Aqu = Ass[:nu, :nu], Ass[nu:, :nu]
    assert spl.eigsh(Auu.tocsc(), k=1, which="SA", return_eigenvectors=False)[0] > 0.0      # velocity block SPD
    assert abs(np.ones(nq) @ Aqu).max() < 1e-14                                               # discrete Gauss theorem, closed box
    P = oracle.Problem(n, n, n, 2.0, 2.0, 2.0)
    A, _ = P.assemble(P.uu0)
    sIA, sJA, sval = bsr_to_csr(P.IA, P.JA, A)
    App = sp.csr_matrix((sval, sJA, sIA), shape=(P.N, P.N)); App.eliminate_zeros()
    S = problems.pnp_stokes_block_system(App, n, coupling=0.5)
    App2, Aps, Asp, Ass2 = S["blocks"]
    assert Aps.shape == (P.N, nu + nq) and Asp.shape == (nu + nq, P.N)
    # ion rows feel the velocity, the velocity rows feel the potential; Dirichlet rows / wall faces stay decoupled
    rows = np.unique(Aps.nonzero()[0]); cols = np.unique(Asp.nonzero()[1])
    assert set(rows % 3) == {1, 2} and set(cols % 3) == {0}
    assert not np.any(P.bcflag[rows]) and not np.any(P.bcflag[cols])
    assert Aps[:, nu:].nnz == 0 and Asp[nu:, :].nnz == 0
    x = spl.spsolve(S["A"].tocsc(), S["b"])
    assert np.linalg.norm(x - S["x_star"]) <= 1e-9 * np.linalg.norm(S["x_star"])


def test_reference_include_pattern_compiles_and_links(pkg, tmp_path):
    """The reference reaches FASP / fasp4ns through `extern "C" { #include "fasp.h" ... }`
    (benchmarks/pnp_stokes/linear_pnp_ns.h:13-18) and sets its parameter blocks up as in
    benchmarks/pnp_stokes/main.cpp:55-86. With include/pnpb200/compat on the include path that code
    compiles against this library's headers and links libpnpb200.so (host-side calls only here)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "modular-pnp_b200", "host")
    src = tmp_path / "main.cpp"
    src.write_text(r'''
#include <cstdio>
extern "C" {
  #include "fasp.h"
  #include "fasp_functs.h"
  #include "fasp4ns.h"
  #include "fasp4ns_functs.h"
}
int main(int argc, char** argv) {
  input_param inpar; itsolver_param itpar; AMG_param amgpar; ILU_param ilupar;
  fasp_param_input(argv[1], &inpar);
  fasp_param_init(&inpar, &itpar, &amgpar, &ilupar, NULL);
  input_param pnp_inpar; itsolver_param pnp_itpar; AMG_param pnp_amgpar; ILU_param pnp_ilupar; Schwarz_param pnp_schpar;
  fasp_param_input(argv[2], &pnp_inpar);
  fasp_param_init(&pnp_inpar, &pnp_itpar, &pnp_amgpar, &pnp_ilupar, &pnp_schpar);
  input_ns_param ns_inpar; itsolver_ns_param ns_itpar; AMG_ns_param ns_amgpar; ILU_param ns_ilupar; Schwarz_param ns_schpar;
  fasp_ns_param_input(argv[3], &ns_inpar);
  fasp_ns_param_init(&ns_inpar, &ns_itpar, &ns_amgpar, &ns_ilupar, &ns_schpar);
  block_dCSRmat A; A.brow = 2; A.bcol = 2; A.blocks = NULL;
  ivector dofs; fasp_ivec_alloc(4, &dofs);
  dvector b, x; fasp_dvec_alloc(3, &b); fasp_dvec_alloc(3, &x);
  INT status = fasp_solver_bdcsr_krylov_pnp_stokes(&A, &b, &x, &itpar, &pnp_itpar, &pnp_amgpar, &ns_itpar, &ns_amgpar, 2, 1);
  printf("%d %d %d %d %d %d\n", (int)itpar.restart, (int)pnp_itpar.restart, (int)ns_itpar.restart, (int)ns_amgpar.param_p.coarse_dof,
         dofs.row, (int)status);
  fasp_dvec_free(&b); fasp_dvec_free(&x);
  return 0;
}
''')
    exe = str(tmp_path / "main")
    cc = subprocess.run(["g++", "-std=c++11", "-I", os.path.join(root, "include", "pnpb200", "compat"), str(src), "-o", exe,
                         "-L", os.path.join(root, "modular-pnp_b200"), "-lpnpb200",
                         "-Wl,-rpath," + os.path.join(root, "modular-pnp_b200")], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-3000:]
    out = subprocess.run([exe, os.path.join(host, "bcsr.dat"), os.path.join(host, "bsr.dat"), os.path.join(host, "ns.dat")],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    # restart values of the three files, the pressure block's coarse size, and ERROR_INPUT_PAR for the empty block matrix
    assert out.stdout.split() == ["10", "30", "100", "50", "4", "-13"]


def test_block_preconditioner_restated_with_exact_block_solves(oracle):
    """The algorithm of csrc/block_solver.cu restated with scipy (test infrastructure): outer GMRES
    with the block upper-triangular preconditioner, Stokes block by GMRES preconditioned with
    [A_uu^-1 ; diagonal SIMPLE Schur complement]. With exact velocity / PNP block solves in place of
    the AMG cycles this is the best the GPU path can do; it pins that the scheme itself converges in a
    handful of outer steps on the synthetic system (the GPU tests then check the real thing)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from helpers import bsr_to_csr
    problems = import_module("modular_pnp_b200.problems")
    n = 6
    P = oracle.Problem(n, n, n, 2.0, 2.0, 2.0)
    A, _ = P.assemble(P.uu0)
    sIA, sJA, sval = bsr_to_csr(P.IA, P.JA, A)
    App = sp.csr_matrix((sval, sJA, sIA), shape=(P.N, P.N)); App.eliminate_zeros()
    S = problems.pnp_stokes_block_system(App, n, coupling=0.5)
    App, Aps, Asp, Ass = S["blocks"]
    nu, Afull, b = S["n_vel"], S["A"], S["b"]
    np_, ns = App.shape[0], Ass.shape[0]
    luP, luU = spl.splu(App.tocsc()), spl.splu(Ass[:nu, :nu].tocsc())
    Aqu, Auq = Ass[nu:, :nu], Ass[:nu, nu:]
    sd = Ass[nu:, nu:].diagonal() - (Aqu @ sp.diags(1.0 / Ass[:nu, :nu].diagonal()) @ Auq).diagonal()
    assert np.all(sd < 0.0)                       # -(penalty + B D^-1 B^T): negative definite diagonal
    inner = {"its": 0, "solves": 0}

    def m_stokes(r):
        zu = luU.solve(r[:nu])
        return np.concatenate([zu, (r[nu:] - Aqu @ zu) / sd])

    def stokes_solve(r):
        def cb(_):
            inner["its"] += 1
        z, info = spl.gmres(Ass, r, M=spl.LinearOperator((ns, ns), matvec=m_stokes), rtol=1e-6, restart=31, maxiter=4,
                            callback=cb, callback_type="pr_norm")
        inner["solves"] += 1
        assert info == 0
        return z

    def m_upper(r):
        zs = stokes_solve(r[np_:])
        return np.concatenate([luP.solve(r[:np_] - Aps @ zs), zs])

    outer = {"its": 0}

    def cb(_):
        outer["its"] += 1
    x, info = spl.gmres(Afull, b, M=spl.LinearOperator(Afull.shape, matvec=m_upper), rtol=1e-8, restart=10, maxiter=10,
                        callback=cb, callback_type="pr_norm")
    assert info == 0 and outer["its"] <= 8
    assert np.linalg.norm(b - Afull @ x) <= 1e-8 * np.linalg.norm(b)
    assert np.linalg.norm(x - S["x_star"]) <= 1e-6 * np.linalg.norm(S["x_star"])
    assert inner["its"] / inner["solves"] <= 40   # diagonal Schur complement: mesh-size independent for a stable pair
