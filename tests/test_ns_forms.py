"""-m "not gpu": the PNP-Stokes forms (benchmarks/pnp_stokes/vector_linear_pnp_ns_forms.ufl, mixed space
P1^3 x RT1 x DG0 with interior-facet DG terms). Three statements of the same element / row arithmetic are pinned to the
golden vectors made by the REFERENCE's own compiled kernels (tests/golden/ns_kernels.npz <- make_golden.py ns <-
vector_linear_pnp_ns_forms.h:16355-23898):
  * the oracle's quadrature-loop restatement (oracle/ns_elements.cpp),
  * the oracle's global cell + facet assembly loop,
  * the PRODUCT's row-wise closed-form code (modular-pnp_b200/csrc/ns_forms.cuh, what the CUDA kernels call), compiled
    for the host by tests/host_shims/ns_rows_host.cpp — so the arithmetic of the device path is checked without a GPU.
Bars: 1e-12 of the largest entry (double rounding; the quadrature rules are restated from closed forms, not from the
header's 15-digit tables)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "ns_kernels.npz"))


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_oracle_element_kernels_match_the_reference_vectors(oracle):
    L = oracle.lib()
    n = G["X0"].shape[0]
    assert n >= 32 and (np.linalg.det((G["X0"].reshape(n, 4, 3)[:, 1:] - G["X0"].reshape(n, 4, 3)[:, :1])) < 0).any()
    for t in range(n):
        x0, x1 = G["X0"][t].copy(), G["X1"][t].copy()
        cc, uu, pp, par = G["CC"][t].copy(), G["UU"][t].copy(), G["PP"][t].copy(), G["PAR"][t].copy()
        f0, f1 = int(G["F"][t, 0]), int(G["F"][t, 1])
        A = np.zeros(289); b = np.zeros(17); Af = np.zeros(1156); bf = np.zeros(34)
        L.orc_ns_cell_a(A, cc[:12].copy(), uu[:4].copy(), par, x0)
        L.orc_ns_cell_L(b, cc[:12].copy(), uu[:4].copy(), pp[:1].copy(), par, x0)
        L.orc_ns_facet_a(Af, par, x0, x1, f0, f1)
        L.orc_ns_facet_L(bf, uu, par, x0, x1, f0, f1)
        assert _rel(A, G["A_cell"][t]) < 1e-12 and _rel(b, G["b_cell"][t]) < 1e-12, t
        assert _rel(Af, G["A_facet"][t]) < 1e-12 and _rel(bf, G["b_facet"][t]) < 1e-12, t


def test_facet_tensor_structure():
    """what the row-wise product code relies on: the DG terms couple velocity dofs only, and the macro tensor is
    symmetric (the three dS terms of the form are a symmetric interior-penalty triple)"""
    for t in range(G["X0"].shape[0]):
        A = G["A_facet"][t].reshape(34, 34)
        rt = [12, 13, 14, 15, 29, 30, 31, 32]
        mask = np.ones((34, 34), bool); mask[np.ix_(rt, rt)] = False
        assert np.abs(A[mask]).max() == 0.0
        assert np.abs(A - A.T).max() <= 1e-12 * np.abs(A).max()


def _box_state(oracle):
    nx, ny, nz = (int(v) for v in G["box"])
    lx, ly, lz = (float(v) for v in G["lengths"])
    P = oracle.Problem(nx, ny, nz, lx, ly, lz)
    return P, G["cc"].copy(), G["uu"].copy(), G["pp"].copy(), G["bc"].copy()


def test_numpy_topology_is_the_fixture_topology(oracle):
    P, *_ = _box_state(oracle)
    cs, cf, fc = oracle.ns_topology(P.cells)
    assert (cs == G["cells_sorted"]).all() and (cf == G["cell_facets"]).all() and (fc == G["facet_cells"]).all()
    # every facet has one or two cells, every interior facet appears in exactly two cells
    cnt = np.bincount(cf.reshape(-1), minlength=fc.shape[0])
    assert ((cnt == 2) == (fc[:, 1] >= 0)).all() and ((cnt == 1) == (fc[:, 1] < 0)).all()


def test_oracle_global_assembly_matches_the_reference_driven_assembly(oracle):
    import scipy.sparse as sp
    P, cc, uu, pp, bc = _box_state(oracle)
    cs, cf, fc = oracle.ns_topology(P.cells)
    A, b = oracle.ns_assemble(P.xyz, cs, cf, fc, cc, uu, pp, bc=bc)
    Ag = sp.csr_matrix((G["A"], G["JA"], G["IA"]), shape=A.shape)
    assert abs(A - Ag).max() < 1e-12 * abs(Ag).max() and _rel(b, G["b"]) < 1e-12
    if os.path.exists(oracle.REF_NS_PATH):       # authoring container: the reference kernels themselves, live
        Ar, br = oracle.ns_assemble(P.xyz, cs, cf, fc, cc, uu, pp, bc=bc, use_ref=True)
        assert abs(A - Ar).max() < 1e-12 * abs(Ar).max() and _rel(b, br) < 1e-12


@pytest.fixture(scope="module")
def row_shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "ns_rows_host.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-x", "c++", "-I" + os.path.join(ROOT, "modular-pnp_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_shims", "ns_rows_host.cpp"), "-o", so])
    return C.CDLL(so)


def _run_rows(shim, xyz, cs, cf, fc, cc, uu, pp, par, ia, ja, bc_flags, matrix=True):
    nv, nc, nf = xyz.size // 3, cs.shape[0], fc.shape[0]
    order = np.argsort(cs.reshape(-1), kind="stable")
    vcell = (order // 4).astype(np.int32)
    ptr = np.zeros(nv + 1, dtype=np.int64); np.add.at(ptr, cs.reshape(-1) + 1, 1); ptr = np.cumsum(ptr).astype(np.int32)
    ndof = 3 * nv + nf + nc
    mom = np.zeros(40 * nc); val = np.zeros(ja.size); rhs = np.zeros(ndof)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    keep = [np.ascontiguousarray(a) for a in (xyz, cs, cf, fc, ptr, vcell, cc, uu, pp, par, ia, ja, bc_flags)]
    shim.shim_ns_assemble(nv, nc, nf, *[p(a) for a in keep], p(mom), p(val), p(rhs), 1 if matrix else 0)
    return val, rhs


def test_product_row_code_matches_the_reference_driven_assembly(oracle, row_shim):
    import scipy.sparse as sp
    P, cc, uu, pp, bc = _box_state(oracle)
    cs, cf, fc = oracle.ns_topology(P.cells)
    ndof = G["IA"].size - 1
    flags = np.zeros(ndof, dtype=np.uint8); flags[bc] = 1
    val, rhs = _run_rows(row_shim, P.xyz, cs, cf, fc, cc, uu, pp, oracle.NS_PAR_DEFAULT.copy(), G["IA"], G["JA"], flags)
    assert _rel(val, G["A"]) < 1e-12 and _rel(rhs, G["b"]) < 1e-12
    # residual-only pass leaves the matrix alone and gives the same right-hand side
    val2, rhs2 = _run_rows(row_shim, P.xyz, cs, cf, fc, cc, uu, pp, oracle.NS_PAR_DEFAULT.copy(), G["IA"], G["JA"], flags, matrix=False)
    assert (val2 == 0).all() and (rhs2 == rhs).all()


def test_product_row_code_on_a_jittered_mesh_with_other_parameters(oracle, row_shim):
    """irregular cells (both orientations), random parameters: product rows vs the oracle's element loop"""
    import scipy.sparse as sp
    P = oracle.Problem(3, 3, 2, 1.5, 1.0, 0.7)
    rng = np.random.default_rng(5)
    xyz = P.xyz + 0.05 * rng.uniform(-1, 1, P.xyz.size)
    perm = rng.permutation(P.nv)                      # renumber the vertices: sorted cells get both orientations
    xyz2 = np.zeros_like(xyz); xyz2.reshape(-1, 3)[perm] = xyz.reshape(-1, 3)
    cells = perm[P.cells].astype(np.int32)
    cs, cf, fc = oracle.ns_topology(cells)
    nv, nc, nf = P.nv, cs.shape[0], fc.shape[0]
    x4 = xyz2.reshape(-1, 3)[cs]
    det = np.linalg.det(x4[:, 1:] - x4[:, :1])
    assert (det > 0).any() and (det < 0).any()
    cc = rng.uniform(-1, 1, 3 * nv); uu = rng.uniform(-1, 1, nf); pp = rng.uniform(-1, 1, nc)
    par = oracle.NS_PAR_DEFAULT * rng.uniform(0.5, 2, 8)
    A, b = oracle.ns_assemble(xyz2, cs, cf, fc, cc, uu, pp, par=par)
    flags = np.zeros(A.shape[0], dtype=np.uint8)
    val, rhs = _run_rows(row_shim, xyz2, cs, cf, fc, cc, uu, pp, par, A.indptr.astype(np.int32), A.indices.astype(np.int32), flags)
    assert _rel(val, A.data) < 1e-12 and _rel(rhs, b) < 1e-12


def test_constant_velocity_is_conforming_and_divergence_free(oracle, pkg):
    """RT1 interpolant of a constant field on the library's basis: one value per facet whichever side computes it
    (the UFC ordering makes s_i sign(detJ) n_i a global facet normal), zero divergence rows, zero DG jump terms"""
    from importlib import import_module
    problems = import_module("modular_pnp_b200.problems")
    P = oracle.Problem(3, 2, 2, 1.0, 1.0, 1.0)
    cs, cf, fc = oracle.ns_topology(P.cells)
    w = np.array([0.3, -1.1, 0.7])
    uu, mismatch = problems.rt_interpolate_constant(P.xyz, cs, cf, fc, w)
    assert mismatch < 1e-13
    nv, nc, nf = P.nv, cs.shape[0], fc.shape[0]
    A, b = oracle.ns_assemble(P.xyz, cs, cf, fc, np.zeros(3 * nv), uu, np.zeros(nc))
    o_u, o_p = 3 * nv, 3 * nv + nf
    assert np.abs(b[o_p:]).max() < 1e-13                       # - div(uu) q
    x = np.zeros(A.shape[0]); x[o_u:o_p] = uu
    assert np.abs((A @ x)[o_p:]).max() < 1e-13                 # div u = 0 through the matrix as well
    interior = fc[:, 1] >= 0
    # velocity rows of L: viscous volume term vanishes (sym grad = 0), DG jump terms vanish -> only the electric force,
    # which is zero for cc = 0 gradients (grad phi = 0)
    assert np.abs(b[o_u:o_p]).max() < 1e-12
