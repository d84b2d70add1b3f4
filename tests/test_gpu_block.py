"""-m gpu: the PNP-Stokes 2x2 block solver through the C ABI
(fasp_solver_bdcsr_krylov_pnp_stokes, benchmarks/pnp_stokes/linear_pnp_ns.cpp:183-193).
fasp4ns' source is not in the reference tree and the reference holds no golden data for this
path ("parity unpinned"), so the checker is a sparse DIRECT solve (scipy, test infrastructure)
of the same block system: the PNP block is the Jacobian this library assembles on the GPU, the
Stokes block a MAC saddle-point matrix, the coupling blocks synthetic (problems.py)."""
import ctypes as C
from importlib import import_module

import numpy as np
import pytest

from helpers import bsr_to_csr, gpu_problem

pytestmark = pytest.mark.gpu


def _system(oracle, pkg, n, coupling):
    import scipy.sparse as sp
    problems = import_module("modular_pnp_b200.problems")
    P = oracle.Problem(n, n, n, 2.0, 2.0, 2.0)
    g = gpu_problem(pkg, P)
    g.assemble()
    IA, JA, val = g.get_matrix()
    g.close()
    sIA, sJA, sval = bsr_to_csr(IA, JA, val)
    App = sp.csr_matrix((sval, sJA, sIA), shape=(P.N, P.N))
    App.eliminate_zeros()
    return problems.pnp_stokes_block_system(App, n, coupling=coupling)


@pytest.mark.parametrize("n,coupling", [(6, 0.05), (10, 0.5)])
def test_block_solve_matches_direct_solve(oracle, pkg, gpu_lib, n, coupling):
    import scipy.sparse.linalg as spl
    S = _system(oracle, pkg, n, coupling)
    params = list(pkg.capi.pnp_stokes_params())
    params[0].tol = 1e-9                      # outer tolerance (bcsr.dat: 1e-6) tightened for the comparison
    st, x, stats = pkg.capi.solve_pnp_stokes(S["blocks"], S["b"], S["n_vel"], S["n_pres"], params)
    assert st > 0, (st, stats)
    A, b = S["A"], S["b"]
    rel = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
    assert rel <= 1e-9 * 1.01 and abs(rel - stats["relres"]) <= 1e-3 * rel + 1e-14     # the reported residual is the true one
    x_ref = spl.spsolve(A.tocsc(), b)
    assert np.linalg.norm(x - x_ref) <= 1e-6 * np.linalg.norm(x_ref)
    # block upper-triangular preconditioning with near-exact block solves: a handful of outer steps
    assert stats["outer_its"] == st and st <= 12
    assert stats["pnp_solves"] == st and stats["stokes_solves"] == st
    assert stats["pnp_its"] >= st and stats["stokes_its"] >= st


def test_bcsr_dat_defaults_and_initial_guess(oracle, pkg, gpu_lib):
    S = _system(oracle, pkg, 6, 0.05)
    A, b = S["A"], S["b"]
    st, x, stats = pkg.capi.solve_pnp_stokes(S["blocks"], b, S["n_vel"], S["n_pres"])      # bcsr.dat: tol 1e-6, restart 10
    assert st > 0 and np.linalg.norm(b - A @ x) <= 1e-6 * np.linalg.norm(b)
    # restarting from the converged iterate needs no iteration (x is an in/out initial guess)
    st2, x2, _ = pkg.capi.solve_pnp_stokes(S["blocks"], b, S["n_vel"], S["n_pres"], x0=x)
    assert st2 == 0 and np.array_equal(x2, x)


def test_blocks_extracted_with_getblk_from_a_mixed_numbering(oracle, pkg, gpu_lib):
    """linear_pnp_ns.cpp:150-157: the four blocks are cut out of ONE mixed-space matrix with
    fasp_dcsr_getblk and index lists; the parent's entry order survives, so block rows are not
    column-sorted. Same answer as with the sorted blocks."""
    import scipy.sparse as sp
    L = pkg.lib()
    S = _system(oracle, pkg, 6, 0.05)
    A, b = S["A"].tocsr(), S["b"]
    N = A.shape[0]
    np_ = S["blocks"][0].shape[0]
    perm = np.random.default_rng(3).permutation(N)            # mixed numbering: new index of block-ordered dof i
    Pm = sp.csr_matrix((np.ones(N), (perm, np.arange(N))), shape=(N, N))
    Amix = (Pm @ A @ Pm.T).tocsr()
    Amix.sort_indices()
    parent = pkg.capi.numpy_to_csr(Amix.indptr, Amix.indices, Amix.data, N)
    pnp_dofs = np.ascontiguousarray(perm[:np_], dtype=np.int32)
    stokes_dofs = np.ascontiguousarray(perm[np_:], dtype=np.int32)
    blocks = []
    mats = [pkg.capi.dCSRmat() for _ in range(4)]
    for m, (Is, Js) in zip(mats, [(pnp_dofs, pnp_dofs), (pnp_dofs, stokes_dofs), (stokes_dofs, pnp_dofs), (stokes_dofs, stokes_dofs)]):
        assert L.fasp_dcsr_getblk(C.byref(parent), Is, Js, len(Is), len(Js), C.byref(m)) == 0
        ia = np.ctypeslib.as_array(m.IA, shape=(m.row + 1,)).copy()
        ja = np.ctypeslib.as_array(m.JA, shape=(max(m.nnz, 1),))[:m.nnz].copy()
        va = np.ctypeslib.as_array(m.val, shape=(max(m.nnz, 1),))[:m.nnz].copy()
        blocks.append(sp.csr_matrix((va, ja, ia), shape=(m.row, m.col)))
        L.fasp_dcsr_free(C.byref(m))
    unsorted = any(np.any(np.diff(B.indices[B.indptr[i]:B.indptr[i + 1]]) < 0) for B in blocks for i in range(B.shape[0]))
    assert unsorted
    params = list(pkg.capi.pnp_stokes_params()); params[0].tol = 1e-9
    st, x, _ = pkg.capi.solve_pnp_stokes(blocks, b, S["n_vel"], S["n_pres"], params)
    st_ref, x_ref, _ = pkg.capi.solve_pnp_stokes(S["blocks"], b, S["n_vel"], S["n_pres"], params)
    assert st == st_ref and st > 0
    # same system, entries summed in a different order: equal up to the solver tolerance
    assert np.linalg.norm(x - x_ref) <= 1e-7 * np.linalg.norm(x_ref)
    assert np.linalg.norm(b - S["A"] @ x) <= 1.01e-9 * np.linalg.norm(b)


def test_block_solver_argument_errors(oracle, pkg, gpu_lib):
    S = _system(oracle, pkg, 6, 0.05)
    st, _, _ = pkg.capi.solve_pnp_stokes(S["blocks"], S["b"], S["n_vel"] + 1, S["n_pres"])
    assert st == -13                           # ERROR_INPUT_PAR: n_vel + n_pres != rows of the Stokes block
    st, _, _ = pkg.capi.solve_pnp_stokes(S["blocks"], S["b"][:-1], S["n_vel"], S["n_pres"])
    assert st == -13


def test_host_mirror_linear_pnp_ns_binary(oracle, pkg, gpu_lib, tmp_path):
    """modular-pnp_b200/host/linear_pnp_ns.hpp (the C++ mirror of Linear_PNP_NS: get_dofs_fasp,
    setup_fasp_linear_algebra, fasp_solve) on a mixed-numbered system handed over as DOLFIN would:
    one CSR matrix, one rhs, the sorted dof list of each of the 5 solution components."""
    import os
    import subprocess
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    S = _system(oracle, pkg, 6, 0.5)
    A, b = S["A"].tocsr(), S["b"]
    N, np_, nu = A.shape[0], S["blocks"][0].shape[0], S["n_vel"]
    # component of every block-ordered dof: PNP dofs interleaved (3 per vertex), then velocity, then pressure
    comp = np.concatenate([np.arange(np_) % 3, np.full(nu, 3), np.full(N - np_ - nu, 4)])
    # mixed numbering: random, but each component's dofs keep their relative order (the lists are sorted)
    rng = np.random.default_rng(11)
    slots = rng.permutation(N)
    new_of = np.empty(N, dtype=np.int64)
    dof_lists = []
    start = 0
    for c in range(5):
        idx = np.nonzero(comp == c)[0]
        mine = np.sort(slots[start:start + len(idx)]); start += len(idx)
        new_of[idx] = mine
        dof_lists.append(mine.astype(np.int32))
    Pm = sp.csr_matrix((np.ones(N), (new_of, np.arange(N))), shape=(N, N))
    Amix = (Pm @ A @ Pm.T).tocsr(); Amix.sort_indices()
    bmix = Pm @ b
    sysfile, outfile = str(tmp_path / "system.bin"), str(tmp_path / "update.bin")
    with open(sysfile, "wb") as f:
        np.array([N, Amix.nnz, 5], dtype=np.int32).tofile(f)
        Amix.indptr.astype(np.int32).tofile(f); Amix.indices.astype(np.int32).tofile(f)
        Amix.data.astype(np.float64).tofile(f); bmix.astype(np.float64).tofile(f)
        for d in dof_lists:
            np.array([len(d)], dtype=np.int32).tofile(f); d.tofile(f)
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "modular-pnp_b200", "host", "pnp_stokes_block")
    out = subprocess.run([exe, sysfile, outfile], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    status = int([l for l in out.stdout.splitlines() if l.startswith("BLOCK_SOLVE_STATUS")][0].split()[1])
    assert 0 < status <= 12
    upd = np.fromfile(outfile, dtype=np.float64)
    assert np.linalg.norm(bmix - Amix @ upd) <= 1.01e-6 * np.linalg.norm(bmix)          # bcsr.dat tolerance
    x_ref = spl.spsolve(Amix.tocsc(), bmix)
    assert np.linalg.norm(upd - x_ref) <= 1e-4 * np.linalg.norm(x_ref)
