"""-m "not gpu": pins the CPU oracle.
  * element kernels against golden vectors produced by the REFERENCE's own compiled kernels
    (tests/golden/make_golden.py; include/EAFE.h:4712-4910, forms.h:7917-8360, :8385-8644),
    and live against oracle/_ref when it is present;
  * the reference's known-answer problem for EAFE (tests/eafe_tests/test_eafe.cpp:39-97,195,272);
  * structural properties of the restated DOLFIN/FASP slice."""
import ctypes as C
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_element_kernels_match_reference_golden(oracle):
    g = np.load(os.path.join(GOLD, "element_kernels.npz"))
    L = oracle.lib()
    for t in range(g["X"].shape[0]):
        A = np.zeros(16); L.orc_eafe_element(A, g["ETA"][t].copy(), g["AL"][t].copy(), g["BE"][t].copy(), g["GA"][t].copy(), g["X"][t].copy())
        assert _rel(A, g["A_eafe"][t]) <= 1e-13
        J = np.zeros(144); L.orc_jac_element(J, g["UU"][t].copy(), g["EPS"][t].copy(), g["D"][t].copy(), g["Q"][t].copy(), g["X"][t].copy())
        assert _rel(J, g["A_jac"][t]) <= 1e-13
        b = np.zeros(12); L.orc_res_element(b, g["UU"][t].copy(), g["EPS"][t].copy(), g["FC"][t].copy(), g["D"][t].copy(), g["Q"][t].copy(), g["RE"][t].copy(), g["X"][t].copy())
        assert _rel(b, g["b_res"][t]) <= 1e-13


def test_element_kernels_match_reference_live(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libpnpref.so not built (needs /root/reference)")
    ref = C.CDLL(oracle.REF_PATH)
    f8 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ref.ref_eafe_tabulate.argtypes = [f8] * 6; ref.ref_jac_tabulate.argtypes = [f8] * 6; ref.ref_res_tabulate.argtypes = [f8] * 8
    L = oracle.lib()
    rng = np.random.default_rng(99)
    for _ in range(100):
        x = (np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1.0]) + 0.2 * rng.uniform(-1, 1, 12)) * rng.uniform(0.01, 2)
        uu = rng.uniform(-1, 1, 12); eps = rng.uniform(.5, 2, 4); D = rng.uniform(.1, 2, 12); q = rng.uniform(-2, 2, 12)
        fc = rng.uniform(-1, 1, 4); re = rng.uniform(-1, 1, 12)
        A = np.zeros(144); B = np.zeros(144)
        L.orc_jac_element(A, uu, eps, D, q, x); ref.ref_jac_tabulate(B, uu, eps, D, q, x)
        assert _rel(A, B) <= 1e-13
        a = np.zeros(12); b = np.zeros(12)
        L.orc_res_element(a, uu, eps, fc, D, q, re, x); ref.ref_res_tabulate(b, uu, eps, fc, D, q, re, x)
        assert _rel(a, b) <= 1e-13


@pytest.mark.parametrize("grid", [(4, 2, 2), (8, 4, 4)])
def test_assembly_matches_reference_driven_golden(oracle, grid):
    """global matrix assembled with the restated kernels == the one assembled by driving the
    reference's own kernels (golden, 4x2x2 and 8x4x4 boxes, EAFE and Galerkin)"""
    g = np.load(os.path.join(GOLD, "assembled_%dx%dx%d.npz" % grid))
    P = oracle.Problem(*grid)
    assert np.array_equal(P.IA, g["IA"]) and np.array_equal(P.JA, g["JA"])
    A, b = P.assemble(g["uu"], use_eafe=True)
    assert _rel(A, g["A_eafe"]) <= 1e-13 and _rel(b, g["b"]) <= 1e-13
    A2, _ = P.assemble(g["uu"], use_eafe=False)
    assert _rel(A2, g["A_galerkin"]) <= 1e-13


def test_box_mesh_and_pattern(oracle):
    P = oracle.Problem(5, 3, 2)
    assert P.nv == 6 * 4 * 3 and P.nc == 6 * 5 * 3 * 2
    cells = P.cells.reshape(-1, 4)
    assert np.all(np.diff(cells, axis=1) > 0)                 # UFC ordering
    # interior vertex of a Kuhn mesh has 15 pattern entries
    ix, iy, iz = 2, 1, 1
    v = iz * 6 * 4 + iy * 6 + ix
    assert P.IA[v + 1] - P.IA[v] == 15
    # Dirichlet set = the two x-faces (main.cpp:293-302)
    assert P.n_bc_vertices == 2 * 4 * 3
    x = P.xyz[0::3]
    assert np.all((np.abs(x[P.bcflag[0::3] == 1]) - 1.0) == 0.0)
    # volumes sum to the box volume
    X = P.xyz.reshape(-1, 3)[cells]
    vol = np.abs(np.linalg.det(X[:, 1:] - X[:, :1])).sum() / 6
    assert abs(vol - 2.0 * 0.4 * 0.4) < 1e-13


def test_eafe_column_sums_and_edge_formula(oracle):
    """EAFE.h:4876-4879 (zero column sums) and SURVEY App. A.3: A_ij = omega_ij * bern_ij"""
    P = oracle.Problem(4, 3, 3)
    rng = np.random.default_rng(5)
    eta = rng.uniform(-1, 1, P.nv); beta = eta + rng.uniform(-2, 2, P.nv); alpha = rng.uniform(.2, 2, P.nv)
    gamma = np.zeros(P.nv)
    val = np.zeros(P.nnzb)
    oracle.lib().orc_eafe_matrix(P.nv, P.xyz, P.nc, P.cells, P.IA, P.JA, eta, alpha, beta, gamma, val)
    rows = np.repeat(np.arange(P.nv), np.diff(P.IA))
    colsum = np.zeros(P.nv); np.add.at(colsum, P.JA, val)
    assert np.abs(colsum).max() <= 1e-13 * np.abs(val).max()
    one = np.ones(P.nv); zero = np.zeros(P.nv); om = np.zeros(P.nnzb)
    oracle.lib().orc_eafe_matrix(P.nv, P.xyz, P.nc, P.cells, P.IA, P.JA, zero, one, zero, gamma, om)   # = P1 stiffness
    psi = eta - beta
    d = psi[P.JA] - psi[rows]
    B = np.where(np.abs(d) < 3e-16, 1 - 0.5 * d, d / np.expm1(d))
    pred = om * 0.5 * (alpha[rows] + alpha[P.JA]) * np.exp(psi[P.JA]) * B * np.exp(beta[P.JA])
    off = rows != P.JA
    assert np.abs(pred[off] - val[off]).max() <= 1e-12 * np.abs(val).max()


@pytest.mark.parametrize("direction", ["x", "y"])
def test_eafe_known_answer_problem(oracle, direction):
    """tests/eafe_tests/test_eafe.cpp: -div(e^beta (grad u + u grad beta)) = 1 on [0,1]^2 x [0,0.03],
    alpha = 1, eta = beta (beta = x, then beta = -y^2), Dirichlet data from the analytic solution;
    pass criterion int (u_h-u)^2 / int u^2 <= 1e-6 (:195, :272)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    m = 40
    nv, nc = C.c_int(), C.c_int()
    L = oracle.lib()
    L.orc_box_mesh_counts(m, m, 3, C.byref(nv), C.byref(nc))
    nv, nc = nv.value, nc.value
    xyz = np.zeros(3 * nv); cells = np.zeros(4 * nc, dtype=np.int32)
    L.orc_box_mesh(m, m, 3, 1.0, 1.0, 3.0 / m, xyz, cells)
    X = xyz.reshape(-1, 3) + np.array([0.5, 0.5, 1.5 / m])      # BoxMesh(p0=(0,0,0), p1=(1,1,3/m))
    xyz = np.ascontiguousarray(X.ravel())
    IA = np.zeros(nv + 1, dtype=np.int32)
    nnz = L.orc_block_pattern_count(nv, nc, cells, IA)
    JA = np.zeros(nnz, dtype=np.int32)
    L.orc_block_pattern_fill(nv, nc, cells, IA, JA)
    c = X[:, 0] if direction == "x" else X[:, 1]
    if direction == "x":
        beta = c.copy(); bL, bR, gl, gr = 0.0, 1.0, 2.0, 3.0
    else:
        beta = -c * c; bL, bR, gl, gr = 0.0, -1.0, 4.0, 2.0
    u_exact = -0.5 * np.exp(-beta) * (c - 0.0) * (c - 1.0) + gl * np.exp(bL - beta) * (1.0 - c) + gr * np.exp(bR - beta) * (c - 0.0)
    val = np.zeros(nnz)
    L.orc_eafe_matrix(nv, xyz, nc, cells, IA, JA, beta, np.ones(nv), beta, np.zeros(nv), val)
    A = sp.csr_matrix((val, JA, IA), shape=(nv, nv)).tolil()
    # rhs int 1*v and P1 mass matrix
    cv = cells.reshape(-1, 4)
    vol = np.abs(np.linalg.det(X[cv][:, 1:] - X[cv][:, :1])) / 6
    b = np.zeros(nv); np.add.at(b, cv.ravel(), np.repeat(vol / 4, 4))
    bc = (c < 1e-12) | (c > 1 - 1e-12)
    for i in np.nonzero(bc)[0]:
        A.rows[i] = [int(i)]; A.data[i] = [1.0]
    b[bc] = u_exact[bc]
    u = spla.spsolve(A.tocsr(), b)
    Ml = np.array([[2, 1, 1, 1], [1, 2, 1, 1], [1, 1, 2, 1], [1, 1, 1, 2]]) / 20.0
    def l2sq(w):
        wc = w[cv]
        return float(np.einsum("c,ci,ij,cj->", vol, wc, Ml, wc))
    assert l2sq(u - u_exact) / l2sq(u_exact) <= 1e-6


def test_csr_bsr_roundtrip_and_spmv(oracle):
    from helpers import bsr_to_csr
    P = oracle.Problem(5, 3, 3)
    A, b = P.assemble(P.uu0)
    sIA, sJA, sval = bsr_to_csr(P.IA, P.JA, A)
    L = oracle.lib()
    bIA = np.zeros(P.nv + 1, dtype=np.int32)
    L.orc_csr_to_bsr3_count(3 * P.nv, sIA, sJA, bIA)
    bJA = np.zeros(bIA[-1], dtype=np.int32); bval = np.zeros(9 * bIA[-1])
    L.orc_csr_to_bsr3_fill(3 * P.nv, sIA, sJA, sval, bIA, bJA, bval)
    assert np.array_equal(bIA, P.IA) and np.array_equal(bJA, P.JA) and np.array_equal(bval, A)
    import scipy.sparse as sp
    S = sp.csr_matrix((sval, sJA, sIA), shape=(P.N, P.N))
    x = np.random.default_rng(1).uniform(-1, 1, P.N)
    assert np.abs(oracle.bsr_spmv(P.IA, P.JA, A, x) - S @ x).max() <= 1e-12 * np.abs(S @ x).max()


def test_vmb_aggregation_and_galerkin_product(oracle):
    import scipy.sparse as sp
    from helpers import bsr_to_csr
    P = oracle.Problem(8, 4, 4)
    A, _ = P.assemble(P.uu0)
    agg, na = oracle.vmb_aggregate(P.IA, P.JA, A, 0.08, 20)
    assert agg.min() >= -1 and agg.max() == na - 1 and na < P.nv / 2
    sizes = np.bincount(agg[agg >= 0], minlength=na)
    assert sizes.min() >= 1
    cIA, cJA, cval = oracle.rap_agg(P.IA, P.JA, A, agg, na)
    sIA, sJA, sval = bsr_to_csr(P.IA, P.JA, A)
    S = sp.csr_matrix((sval, sJA, sIA), shape=(P.N, P.N))
    keep = np.nonzero(agg >= 0)[0]
    Pm = sp.csr_matrix((np.ones(3 * len(keep)), ((3 * keep[:, None] + np.arange(3)).ravel(),
                                                   (3 * agg[keep][:, None] + np.arange(3)).ravel())), shape=(P.N, 3 * na))
    Cd = (Pm.T @ S @ Pm).toarray()
    cs = bsr_to_csr(cIA, cJA, cval)
    Co = sp.csr_matrix((cs[2], cs[1], cs[0]), shape=(3 * na, 3 * na)).toarray()
    assert np.abs(Cd - Co).max() <= 1e-12 * np.abs(Cd).max()


def test_pvfgmres_amg_converges_and_newton_counts(oracle):
    P = oracle.Problem(20, 4, 4)
    A, b = P.assemble(P.uu0)
    prm = oracle.default_param()
    x, its = oracle.solver_dbsr_krylov_amg(P.IA, P.JA, A, b, prm)
    assert 0 < its < 100
    r = b - oracle.bsr_spmv(P.IA, P.JA, A, x)
    assert np.linalg.norm(r) <= 1.0001e-6 * np.linalg.norm(b)
    # K-cycle / dense coarse / CGS mirror modes solve the same system
    for kw in (dict(cycle_type=4), dict(coarse_dense=1), dict(ortho_cgs=1)):
        x2, its2 = oracle.solver_dbsr_krylov_amg(P.IA, P.JA, A, b, oracle.default_param(**kw))
        assert 0 < its2 < 100 and np.abs(x2 - x).max() <= 1e-4 * np.abs(x).max()
    # Newton on the reference's refine_exact starting grid (20x4x4): converges, < 15 steps
    uu, res = oracle.newton_exact(20, 4, 4)
    assert res.converged == 1 and res.newton_its < 15
    assert res.rel_res[res.newton_its - 1] < 1e-8 or res.max_res[res.newton_its - 1] < 1e-10
    assert np.abs(uu - P.exact).max() < 0.05      # discretisation error of the manufactured solution


def test_diode_form_variant_matches_reference_golden(oracle):
    """pnp_diode forms (poisson_scale, exp cutoff at -50): restated kernels vs golden vectors from
    the reference's compiled benchmarks/pnp_diode/vector_linear_pnp_forms.h:8388-9755"""
    g = np.load(os.path.join(GOLD, "element_kernels_diode.npz"))
    L = oracle.lib()
    oracle.set_form_variant(float(g["poisson_scale"]), True)
    try:
        for t in range(g["X"].shape[0]):
            J = np.zeros(144); L.orc_jac_element(J, g["UU"][t].copy(), g["EPS"][t].copy(), g["D"][t].copy(), g["Q"][t].copy(), g["X"][t].copy())
            assert _rel(J, g["A_jac"][t]) <= 1e-13
            b = np.zeros(12); L.orc_res_element(b, g["UU"][t].copy(), g["EPS"][t].copy(), g["FC"][t].copy(), g["D"][t].copy(), g["Q"][t].copy(), g["RE"][t].copy(), g["X"][t].copy())
            assert _rel(b, g["b_res"][t]) <= 1e-13
    finally:
        oracle.set_form_variant(1.0, False)


def test_error_norms_against_closed_form(oracle):
    """src/error.cpp:58-103 restated: for e = (a.x + c) in every component on the box the L2 / H1
    norms have closed forms (P1 represents linear functions exactly)."""
    P = oracle.Problem(6, 3, 3)
    x, y, z = P.xyz[0::3], P.xyz[1::3], P.xyz[2::3]
    e = 0.3 * x - 0.2 * y + 0.5 * z + 0.1
    uu = P.exact.copy()
    for k in range(3):
        uu[k::3] += e
    l2, h1 = P.error_norms(uu)
    lx, ly, lz = 2.0, 0.4, 0.4
    vol = lx * ly * lz
    int_e2 = vol * (0.1 ** 2 + 0.3 ** 2 * lx ** 2 / 12 + 0.2 ** 2 * ly ** 2 / 12 + 0.5 ** 2 * lz ** 2 / 12)
    semi = vol * (0.3 ** 2 + 0.2 ** 2 + 0.5 ** 2)
    assert abs(l2 - np.sqrt(3 * int_e2)) <= 1e-12 * l2
    assert abs(h1 - np.sqrt(3 * int_e2 + 3 * semi)) <= 1e-12 * h1


def test_entropy_indicator_matches_reference_golden(oracle):
    """DG0 entropy indicator (Mesh_Refiner::compute_entropy_error_vector, src/mesh_refiner.cpp:212-271):
    the restated element value equals the reference's compiled poisson_cell_marker kernels
    (include/poisson_cell_marker.h:5203-5231, 5255-5422; golden vectors from oracle/_ref)."""
    g = np.load(os.path.join(GOLD, "marker_kernels.npz"))
    L = oracle.lib()
    for t in range(len(g["L"])):
        v = L.orc_marker_element(g["MU"][t].copy(), g["LW"][t].copy(), g["DF"][t].copy(), g["X"][t].copy())
        ref = g["L"][t] / g["M"][t]          # mass_lumping_solver: rhs / row sum (src/mesh_refiner.cpp:320-330)
        assert abs(v - ref) <= 1e-12 * abs(ref), (t, v, ref)
    # global: constant log-densities and a linear potential give D exp(u) q^2 |grad phi|^2 per species
    P = oracle.Problem(6, 3, 2)
    uu = np.zeros(P.N)
    x = P.xyz[0::3]
    uu[0::3] = 0.7 * x; uu[1::3] = -0.3; uu[2::3] = 0.4
    eta = P.entropy_indicator(uu)
    expect = P.diff[1] * np.exp(-0.3) * (0.7 * P.val[1]) ** 2 + P.diff[2] * np.exp(0.4) * (0.7 * P.val[2]) ** 2
    assert eta.shape == (P.nc,) and np.abs(eta - expect).max() <= 1e-12 * expect


def test_diode_damped_newton_on_the_oracle(oracle):
    """pnp_diode physics (benchmarks/pnp_diode/diode.h) + the benchmark's damped Newton
    (pnp_newton_solver.h:195-284) on the CPU oracle at zero bias: growth-limited first steps with one
    backtrack each, then full steps, converged to the benchmark's tolerances (rel 1e-7)."""
    from helpers import OracleNewtonBackend, diode_problem
    P, pr = diode_problem(oracle, 40, 4, 4, 0.0)
    left, right = pr.diode_contacts(0.0)
    x = P.xyz[0::3]
    assert np.allclose(P.uu0.reshape(-1, 3)[np.abs(x + 1) < 1e-12], left) and np.allclose(P.uu0.reshape(-1, 3)[np.abs(x - 1) < 1e-12], right)
    assert abs(pr.DIODE_BETA - 38.6817) < 1e-3 and abs(P.fixed[np.argmin(x)] - 1.0) < 1e-15 and abs(P.fixed[np.argmax(x)] + 1.0) < 1e-15
    prm = oracle.default_param(maxit=500)
    oracle.set_form_variant(pr.DIODE_POISSON_SCALE, True)
    try:
        hist, conv = pr.damped_newton(OracleNewtonBackend(oracle, P, prm), P.uu0, max_newton=60)
    finally:
        oracle.set_form_variant(1.0, False)
    assert conv and 20 <= len(hist) <= 40
    assert all(h[0] > 0 for h in hist)
    assert hist[0][1] < 1e-6 and hist[0][4] == 1            # the first update is cut back by the growth limit
    assert hist[-1][1] == 1.0 and hist[-1][2] < 1e-7        # full steps at the end


def test_newton_history_matches_reference_driven_golden(oracle):
    """BASELINE configs[0] (50x10x10, bsr.dat): the oracle with the RESTATED element kernels walks the
    same Newton path as the golden history generated by driving the reference's own compiled kernels
    (tests/golden/make_golden.py newton): same Newton count (8), same Krylov counts, residual
    histories equal far below the linear-solver tolerance. The GPU path is compared with this
    oracle run in tests/test_gpu_solver.py::test_newton_exact_solution_config."""
    import json
    g = json.load(open(os.path.join(GOLD, "newton_history_50x10x10.json")))
    uu, res = oracle.newton_exact(*g["grid"], *g["lengths"])
    assert res.newton_its == g["newton_its"] == 8 and res.converged == g["converged"] == 1
    assert abs(res.initial_residual - g["initial_residual"]) <= 1e-12 * g["initial_residual"]
    for k in range(res.newton_its):
        assert abs(res.krylov_its[k] - g["krylov_its"][k]) <= 1
        tol = 1e-6 if g["rel_res"][k] > 1e-6 else 1e-2      # the last steps sit at the linear-solver tolerance
        assert abs(res.rel_res[k] - g["rel_res"][k]) <= tol * g["rel_res"][k]
    # main.cpp:317-319: rel 1e-8 / max-res 1e-10
    assert res.rel_res[res.newton_its - 1] < 1e-8 or res.max_res[res.newton_its - 1] < 1e-10
    for c in range(3):
        assert abs(np.sum(uu[c::3]) - g["solution_checksum"][c]) <= 1e-6 * max(1.0, abs(g["solution_checksum"][c]))
