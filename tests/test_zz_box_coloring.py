"""-m gpu: the closed-form 4-colouring of BoxMesh (mesh_setup.cu: box_coloring) against the
Jones-Plassmann colouring of the same mesh handed over as an explicit (xyz, cells) mesh: same
matrix, same Newton step up to the linear-solver tolerance, fewer colour launches per sweep."""
import numpy as np
import pytest

from helpers import gpu_problem

pytestmark = pytest.mark.gpu


def test_box_mesh_uses_the_four_colouring_and_solves_the_same_step(oracle, pkg, gpu_lib):
    from importlib import import_module
    problems = import_module("modular_pnp_b200.problems")
    n = (24, 12, 10)
    P = oracle.Problem(*n, 2.0, 1.0, 0.8)
    it, amg = pkg.default_params()
    # explicit mesh: graph colouring
    ge = gpu_problem(pkg, P)
    ge.assemble()
    IAe, JAe, vale = ge.get_matrix()
    ste, l2e, mxe = ge.newton_step(it, amg)
    se = ge.stats()
    uue = ge.get_solution()
    ge.close()
    # the same box generated on the device
    gb = pkg.PNP(box=(n[0], n[1], n[2], 2.0, 1.0, 0.8))
    xyz, cells = gb.get_mesh()
    assert len(cells) == len(P.cells) and np.abs(xyz - P.xyz).max() <= 1e-14
    gb.set_coefficients(P.eps, P.fixed, P.diff, P.val, P.reac)
    gb.set_dirichlet(np.nonzero(P.bcflag)[0].astype(np.int32))
    gb.set_solution(P.uu0)
    gb.assemble()
    IAb, JAb, valb = gb.get_matrix()
    stb, l2b, mxb = gb.newton_step(it, amg)
    sb = gb.stats()
    uub = gb.get_solution()
    gb.close()
    # the assembled matrix does not depend on the sweep order
    assert np.array_equal(IAb, IAe) and np.array_equal(JAb, JAe)
    assert np.abs(valb - vale).max() <= 1e-12 * np.abs(vale).max()
    assert sb.level_colors[0] <= se.level_colors[0] and sb.level_colors[0] >= 4
    assert ste > 0 and stb > 0 and abs(stb - ste) <= 4
    assert abs(l2b - l2e) <= 1e-3 * l2e and np.abs(uub - uue).max() <= 1e-4


def test_full_size_solve_true_residual_through_the_row_id_matvec(pkg, gpu_lib):
    """96^3 box (2.7 M dof): the K-cycle-preconditioned FGMRES solve with every shortcut on (L-only /
    U-only passes, A z from the sweep delta, coarse-level CUDA graph, 4-colour sweeps), checked with
    size-independent properties through an independent path: the row-id-space matvec of the C ABI
    (pnpb200_apply_matrix) — true residual below the tolerance, linearity of the operator."""
    from importlib import import_module
    problems = import_module("modular_pnp_b200.problems")
    n = 96
    pr = problems.exact_solution_problem(n, n, n, 2.0, 2.0, 2.0)
    g = pkg.PNP(box=(n, n, n, 2.0, 2.0, 2.0))
    g.set_coefficients(pr["eps"], pr["fixed"], pr["diff"], pr["val"], pr["reac"])
    g.set_dirichlet(pr["bc_dofs"])
    g.set_solution(pr["uu0"])
    g.assemble()
    b = g.get_rhs()
    it, amg = pkg.default_params(cycle_type=4)
    its = g.solve(it, amg)
    assert 0 < its <= 40, "solver status %d" % its
    du = g.get_update()
    r = b - g.apply_matrix(du)
    assert np.linalg.norm(r) <= 1.001 * it.tol * np.linalg.norm(b)
    rng = np.random.default_rng(4)
    x, y = rng.uniform(-1, 1, g.N), rng.uniform(-1, 1, g.N)
    lhs = g.apply_matrix(2.0 * x + y)
    rhs = 2.0 * g.apply_matrix(x) + g.apply_matrix(y)
    assert np.abs(lhs - rhs).max() <= 1e-12 * np.abs(rhs).max()
    s = g.stats()
    assert s.n_levels >= 4 and 4 <= s.level_colors[0] <= 12
    g.close()
