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
