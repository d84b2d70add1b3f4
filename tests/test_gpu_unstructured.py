"""-m gpu: the general (unstructured) mesh path — SURVEY §8f-1 / BASELINE configs[2]: after
adaptive refinement the mesh is no longer a Kuhn box, so pattern, edge table, colouring and
hierarchy must come from an arbitrary (xyz, cells) pair. Parity against the oracle on a jittered
box with centroid-split cells."""
import numpy as np
import pytest

from helpers import GeneralProblem, irregular_mesh, row_scaled_error

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def irregular(oracle):
    xyz, cells = irregular_mesh(oracle, 12, 5, 4)
    return GeneralProblem(oracle, xyz, cells)


def _gpu(pkg, P, uu):
    g = pkg.PNP(P.xyz, P.cells, 0)
    g.set_coefficients(P.eps, P.fixed, P.diff, P.val, P.reac)
    g.set_dirichlet(np.nonzero(P.bcflag)[0].astype(np.int32))
    g.set_solution(uu)
    return g


@pytest.mark.parametrize("use_eafe", [True, False])
def test_irregular_mesh_matrix_parity(oracle, pkg, gpu_lib, irregular, use_eafe):
    P = irregular
    rng = np.random.default_rng(1234)
    uu = P.exact + 0.1 * rng.uniform(-1, 1, P.N)
    uu[0::3] = np.round(uu[0::3] * 64) / 64
    A_ref, b_ref = P.assemble(uu, use_eafe)
    g = _gpu(pkg, P, uu)
    g.use_eafe(use_eafe)
    g.assemble()
    IA, JA, val = g.get_matrix()
    assert np.array_equal(IA, P.IA) and np.array_equal(JA, P.JA)
    assert np.diff(IA).max() > 15                      # genuinely irregular rows
    assert row_scaled_error(IA, val, A_ref) <= 1e-12
    assert np.abs(g.get_rhs() - b_ref).max() <= 1e-12 * np.abs(b_ref).max()
    g.close()


def test_irregular_mesh_newton_steps(oracle, pkg, gpu_lib, irregular):
    """two Newton steps: GPU newton_step vs oracle assemble + FASP-restated solve + update"""
    P = irregular
    g = _gpu(pkg, P, P.uu0)
    it, amg = pkg.default_params()
    prm = oracle.default_param()
    uu = P.uu0.copy()
    _, r0, _ = P.residual(uu)
    a0, _ = g.residual()
    assert abs(a0 - r0) <= 1e-12 * r0
    for _ in range(2):
        st, l2, mx = g.newton_step(it, amg)
        assert st > 0
        A, b = P.assemble(uu, True)
        dx, its = oracle.solver_dbsr_krylov_amg(P.IA, P.JA, A, b, prm)
        assert its > 0
        uu = uu + dx
        _, l2_ref, mx_ref = P.residual(uu)
        assert abs(l2 - l2_ref) <= 2e-3 * l2_ref
    assert np.abs(g.get_solution() - uu).max() <= 1e-4
    g.close()
