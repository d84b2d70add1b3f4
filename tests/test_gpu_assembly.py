"""-m gpu: CUDA assembly (through the C ABI) against the CPU oracle on the same seeded inputs.
Bar: BSR pattern identical; values within 1e-12 of the row maximum (north_star: 1e-12 relative;
row scaling per SURVEY §7 because 8 of 15 stiffness entries per row cancel on Kuhn meshes)."""
import numpy as np
import pytest

from helpers import gpu_problem, perturbed_state, row_scaled_error

pytestmark = pytest.mark.gpu

GRIDS = [(4, 2, 2), (8, 4, 4), (20, 6, 5)]


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("use_eafe", [True, False])
def test_matrix_and_rhs_parity(oracle, pkg, gpu_lib, grid, use_eafe):
    P = oracle.Problem(*grid)
    uu = perturbed_state(P, quantize=True)
    A_ref, b_ref = P.assemble(uu, use_eafe=use_eafe)
    g = gpu_problem(pkg, P, uu)
    g.use_eafe(use_eafe)
    g.assemble()
    IA, JA, val = g.get_matrix()
    assert np.array_equal(IA, P.IA) and np.array_equal(JA, P.JA)
    err = row_scaled_error(IA, val, A_ref)
    assert err <= 1e-12, "matrix row-scaled error %.3e" % err
    b = g.get_rhs()
    assert np.abs(b - b_ref).max() <= 1e-12 * np.abs(b_ref).max()
    # Dirichlet rows are exactly identity rows, rhs exactly zero there
    bc = np.nonzero(P.bcflag)[0]
    assert np.all(b[bc] == 0.0)
    V = val.reshape(-1, 3, 3)
    rows = np.repeat(np.arange(P.nv), np.diff(IA))
    for d in bc[:50]:
        i, k = divmod(int(d), 3)
        sel = rows == i
        blk = V[sel, k, :]
        diag = JA[sel] == i
        assert blk[diag, k] == 1.0 and np.abs(blk).sum() == 1.0
    g.close()


@pytest.mark.parametrize("grid", [(4, 2, 2), (8, 4, 4)])
def test_matrix_and_rhs_against_committed_golden(oracle, pkg, gpu_lib, grid):
    """the same comparison against the COMMITTED fixtures (tests/golden/assembled_*.npz: the oracle
    assembler driven by the reference's own compiled element kernels), EAFE and Galerkin"""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "assembled_%dx%dx%d.npz" % grid))
    P = oracle.Problem(*grid)
    uu = gold["uu"]
    assert np.array_equal(uu, perturbed_state(P, quantize=True))      # the seeded input of the test above
    g = gpu_problem(pkg, P, uu)
    for use_eafe, key in ((True, "A_eafe"), (False, "A_galerkin")):
        g.use_eafe(use_eafe)
        g.assemble()
        IA, JA, val = g.get_matrix()
        assert np.array_equal(IA, gold["IA"]) and np.array_equal(JA, gold["JA"])
        err = row_scaled_error(IA, val, gold[key])
        assert err <= 2e-12, "%s row-scaled error %.3e" % (key, err)
        b = g.get_rhs()
        assert np.abs(b - gold["b"]).max() <= 1e-12 * np.abs(gold["b"]).max()
    g.close()


def test_matrix_parity_unquantized_is_conditioning_limited(oracle, pkg, gpu_lib):
    """Continuous random potentials: the reference's own d/(exp(d)-1) carries eps/|d| relative
    error, so parity is checked against 1e-12 + 8 eps/|d_min| per row."""
    P = oracle.Problem(8, 4, 4)
    uu = perturbed_state(P, quantize=False)
    A_ref, _ = P.assemble(uu, use_eafe=True)
    g = gpu_problem(pkg, P, uu)
    g.assemble()
    IA, JA, val = g.get_matrix()
    rows = np.repeat(np.arange(P.nv), np.diff(IA))
    d = np.abs(uu[0::3][JA] - uu[0::3][rows])
    d = d[d > 3e-16]
    tol = 1e-12 + 8 * 2.2e-16 / d.min()
    err = row_scaled_error(IA, val, A_ref)
    assert err <= tol, "row-scaled error %.3e > %.3e" % (err, tol)
    g.close()


@pytest.mark.parametrize("grid", GRIDS)
def test_residual_norms(oracle, pkg, gpu_lib, grid):
    P = oracle.Problem(*grid)
    for uu in (P.uu0, perturbed_state(P)):
        _, l2, li = P.residual(uu)
        g = gpu_problem(pkg, P, uu)
        a, b = g.residual()
        assert abs(a - l2) <= 1e-12 * l2 and abs(b - li) <= 1e-12 * li
        g.close()


def test_full_size_box_invariants(oracle, pkg, gpu_lib):
    """64^3 box (0.82 M dof): size-independent properties instead of the oracle —
    EAFE (k,k) column sums vanish on interior columns, (k,1-k) coupling blocks are zero,
    Dirichlet rows are identity rows, J*1 restricted to the potential row of phi-phi block is 0."""
    n = 64
    P = oracle.Problem(n, n, n, 2.0, 2.0, 2.0)
    uu = perturbed_state(P)
    g = gpu_problem(pkg, P, uu)
    g.assemble()
    IA, JA, val = g.get_matrix()
    V = val.reshape(-1, 3, 3)
    assert np.all(V[:, 1, 2] == 0.0) and np.all(V[:, 2, 1] == 0.0)
    rows = np.repeat(np.arange(P.nv), np.diff(IA))
    interior_rows = P.bcflag[0::3][rows] == 0
    for k in (1, 2):
        colsum = np.zeros(P.nv)
        np.add.at(colsum, JA[interior_rows], V[interior_rows, k, k])
        # columns all of whose rows are interior
        touched_bc = np.zeros(P.nv, dtype=bool)
        touched_bc[JA[~interior_rows]] = True
        scale = np.abs(V[:, k, k]).max()
        assert np.abs(colsum[~touched_bc]).max() <= 1e-12 * scale
    # permittivity stiffness annihilates constants on interior rows
    rowsum = np.zeros(P.nv)
    np.add.at(rowsum, rows, V[:, 0, 0])
    assert np.abs(rowsum[P.bcflag[0::3] == 0]).max() <= 1e-12 * np.abs(V[:, 0, 0]).max()
    g.close()


@pytest.mark.parametrize("use_eafe", [True, False])
def test_diode_form_variant_parity(oracle, pkg, gpu_lib, use_eafe):
    """diode-like input — poisson_scale != 1, diffusivity jump, |q phi| up to 20
    (stresses the Bernoulli branch) and log-densities below the -50 cutoff of conc(w)."""
    P = oracle.Problem(10, 4, 4)
    rng = np.random.default_rng(77)
    x = P.xyz[0::3]
    uu = np.zeros(P.N)
    uu[0::3] = np.round(20.0 * np.tanh(3 * x) * 64) / 64 + np.round(rng.uniform(-0.5, 0.5, P.nv) * 64) / 64
    uu[1::3] = -26.0 * (1 + np.tanh(4 * x)) + 0.1 * rng.uniform(-1, 1, P.nv)      # down to -52
    uu[2::3] = -26.0 * (1 - np.tanh(4 * x)) + 0.1 * rng.uniform(-1, 1, P.nv)
    P.diff[1::3] = np.where(x > 0, 1.0, 0.1); P.diff[2::3] = np.where(x > 0, 2.0, 0.1)
    ps = 12.5
    oracle.set_form_variant(ps, True)
    try:
        A_ref, b_ref = P.assemble(uu, use_eafe=use_eafe)
    finally:
        oracle.set_form_variant(1.0, False)
    g = gpu_problem(pkg, P, uu)
    g.set_form_variant(ps, True)
    g.use_eafe(use_eafe)
    g.assemble()
    IA, JA, val = g.get_matrix()
    assert row_scaled_error(IA, val, A_ref) <= 1e-12
    assert np.abs(g.get_rhs() - b_ref).max() <= 1e-12 * np.abs(b_ref).max()
    g.close()


def test_error_norms_and_total_charge(oracle, pkg, gpu_lib):
    """L2/H1 error functionals (src/error.cpp:58-103) and the nodal total charge
    (linear_pnp.cpp:322-351) on the device vs the oracle / numpy."""
    P = oracle.Problem(12, 5, 4)
    uu = perturbed_state(P)
    g = gpu_problem(pkg, P, uu)
    l2, h1 = g.error_norms(P.exact)
    l2_ref, h1_ref = P.error_norms(uu)
    assert abs(l2 - l2_ref) <= 1e-12 * l2_ref and abs(h1 - h1_ref) <= 1e-12 * h1_ref
    q = g.total_charge()
    q_ref = P.fixed + P.val[1] * np.exp(uu[1::3]) + P.val[2] * np.exp(uu[2::3])
    assert np.abs(q - q_ref).max() <= 1e-13 * np.abs(q_ref).max()
    g.close()


def test_entropy_indicator_and_cell_marking(oracle, pkg, gpu_lib):
    """the DG0 entropy indicator (Mesh_Refiner::compute_entropy_error_vector,
    src/mesh_refiner.cpp:212-271) and the two marking rules (:155-209) on the device, on a Kuhn box
    and on an irregular (jittered, centroid-split) mesh, vs the oracle / the rules restated in numpy."""
    from helpers import GeneralProblem, irregular_mesh
    P1 = oracle.Problem(12, 5, 4)
    xyz, cells = irregular_mesh(oracle, 10, 4, 4)
    P2 = GeneralProblem(oracle, xyz, cells)
    for P in (P1, P2):
        rng = np.random.default_rng(77)
        uu = P.exact + 0.2 * rng.uniform(-1, 1, P.N)
        if P is P1:
            g = gpu_problem(pkg, P, uu)
        else:
            g = pkg.PNP(P.xyz, P.cells, 0)
            g.set_coefficients(P.eps, P.fixed, P.diff, P.val, P.reac)
            g.set_dirichlet(np.nonzero(P.bcflag)[0].astype(np.int32))
            g.set_solution(uu)
        eta = g.entropy_indicator()
        eta_ref = np.zeros(P.nc)
        oracle.lib().orc_entropy_indicator(P.nv, P.xyz, P.nc, P.cells, uu, P.diff, P.val, eta_ref)
        assert eta.shape == (P.nc,) and (eta_ref > 0).all()
        assert np.abs(eta - eta_ref).max() <= 1e-12 * np.abs(eta_ref).max()
        # mark_for_refinement: eta > tol
        tol = float(np.median(eta))
        m, cnt, used = g.mark_cells(tol)
        assert used == tol and cnt == int((eta > tol).sum()) and np.array_equal(m, eta > tol)
        # mark_for_refinement_with_target_size
        srt = np.sort(eta)
        for target in (P.nc + P.nc // 5, P.nc // 2, 10 * P.nc):
            permissible = 1 if P.nc > target else target - P.nc
            idx = min(P.nc if permissible > P.nc else P.nc - permissible, P.nc - 1)
            t_ref = max(srt[idx], 1e-6)
            m, cnt, used = g.mark_cells(1e-6, target)
            assert used == t_ref and np.array_equal(m, eta > t_ref) and cnt == int((eta > t_ref).sum())
        g.close()


@pytest.mark.parametrize("use_eafe", [True, False])
def test_zero_row_gets_unit_diagonal(oracle, pkg, gpu_lib, use_eafe):
    """SURVEY a8, PDE::EigenMatrix_to_dCSRmat (src/pde.cpp:596-621): a scalar row whose stored values
    are all zero gets diag = 1. Such rows appear when an ion's diffusivity vanishes around a vertex:
    every entry of that ion's row carries a factor D (Galerkin: forms.ufl:30-33; EAFE: the edge
    average of alpha, EAFE.h:4857-4867). Vertex i and all its neighbours get D_1 = 0 -> row (i, 1) is
    all-zero before the fix and exactly e_i after it, on the device as in the oracle."""
    P = oracle.Problem(8, 4, 4)
    X = P.xyz.reshape(-1, 3)
    i = int(np.argmin(np.abs(X).sum(axis=1)))                     # an interior vertex
    nbrs = P.JA[P.IA[i]:P.IA[i + 1]]
    assert len(nbrs) == 15
    saved = P.diff.copy()
    P.diff[3 * nbrs + 1] = 0.0
    uu = perturbed_state(P, quantize=True)
    try:
        A_ref, b_ref = P.assemble(uu, use_eafe=use_eafe)
        g = gpu_problem(pkg, P, uu)
    finally:
        P.diff[:] = saved
    g.use_eafe(use_eafe)
    g.assemble()
    IA, JA, val = g.get_matrix()
    V = val.reshape(-1, 3, 3); R = A_ref.reshape(-1, 3, 3)
    sl = slice(IA[i], IA[i + 1])
    row, row_ref = V[sl, 1, :], R[sl, 1, :]
    d = int(np.nonzero(JA[sl] == i)[0][0])
    expect = np.zeros_like(row); expect[d, 1] = 1.0
    assert np.array_equal(row_ref, expect), "oracle: the zero row must come out as e_i"
    assert np.array_equal(row, expect), "device: the zero row must come out as e_i"
    # the other two rows of the vertex and all other rows are untouched by the fix
    assert np.abs(V[sl, 0, :]).sum() > 0 and np.abs(V[sl, 2, :]).sum() > 0
    assert row_scaled_error(IA, val, A_ref) <= 1e-12
    g.close()


def test_caller_dof_numbering(oracle, pkg, gpu_lib):
    """SURVEY a14 / pnpb200_set_dof_map: a DOLFIN-style renumbering (nodes permuted, the three dofs of a node
    kept together, P1 coefficient spaces with their own permutation) and a fully general dof permutation.
    Every dof-indexed array crosses the ABI in the caller's numbering; results equal the oracle's (vertex
    numbering) after mapping back: rhs 1e-12, two Newton steps' residual norms 1e-12 / 2e-3 (as elsewhere)."""
    P = oracle.Problem(10, 4, 4)
    rng = np.random.default_rng(5)
    perm_v = rng.permutation(P.nv)
    block_map = (3 * perm_v[:, None] + np.arange(3)[None, :]).ravel()          # vertex_to_dof_map of a node-reordered space
    general_map = rng.permutation(P.N)
    smap = rng.permutation(P.nv)
    uu = perturbed_state(P, quantize=True)
    _, b_ref = P.assemble(uu, use_eafe=True)
    _, l2_ref, mx_ref = P.residual(uu)
    for v2d in (block_map, general_map):
        def ext(a):                     # internal (3v + k) -> caller's numbering
            out = np.empty_like(a); out[v2d] = a; return out

        def sext(a):
            out = np.empty_like(a); out[smap] = a; return out
        g = pkg.PNP(P.xyz, P.cells, 0)
        g.set_dof_map(v2d, smap)
        g.set_coefficients(sext(P.eps), sext(P.fixed), ext(P.diff), ext(P.val), ext(P.reac))
        g.set_dirichlet(v2d[np.nonzero(P.bcflag)[0]].astype(np.int32))
        g.set_solution(ext(uu))
        l2, mx = g.residual()
        assert abs(l2 - l2_ref) <= 1e-12 * l2_ref and abs(mx - mx_ref) <= 1e-12 * mx_ref
        g.assemble()
        b = g.get_rhs()
        assert np.abs(b[v2d] - b_ref).max() <= 1e-12 * np.abs(b_ref).max()
        assert np.array_equal(g.get_solution(), ext(uu))
        it, amg = pkg.default_params()
        st, l2n, _ = g.newton_step(it, amg)
        assert st > 0
        un = g.get_solution()[v2d]                       # back to vertex numbering
        _, l2o, _ = P.residual(un)
        assert abs(l2n - l2o) <= 1e-10 * l2o
        g.close()
