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
