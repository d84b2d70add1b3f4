"""Known-answer Newton problem (the idea of the reference's tests/pnp_tests/test_pnp_eafe.cpp:
source terms manufactured so that a chosen nodal field solves the DISCRETE problem exactly;
pass = Newton converges in < 15 steps and the nodal error is at round-off of the Newton tolerance).
Runs on the CPU oracle (not gpu) and through the C ABI on the GPU (gpu)."""
import numpy as np
import pytest

from helpers import manufactured_sources


def _target(P):
    x = P.xyz[0::3]; y = P.xyz[1::3]; z = P.xyz[2::3]
    uu = P.exact.copy()
    uu[0::3] += 0.3 * np.cos(2.0 * y) * x
    uu[1::3] += 0.2 * np.sin(3.0 * z + y)
    uu[2::3] -= 0.2 * x * y
    return uu


def _newton(step, residual, max_newton=15, rel_tol=1e-8, max_tol=1e-10):
    l2_0, mx = residual()
    it, rel = 1, 1.0
    while it < max_newton + 1 and rel > rel_tol and mx > max_tol:
        st, l2, mx = step()
        assert st > 0
        rel = l2 / l2_0
        it += 1
    return it - 1, rel


def test_oracle_newton_recovers_manufactured_solution(oracle):
    P = oracle.Problem(10, 5, 4)
    uu_star = _target(P)
    P.fixed, P.reac = manufactured_sources(oracle, P, uu_star)
    # Dirichlet values on the x-faces = the target; start from the target's face values + linear interpolant
    uu = P.uu0.copy()
    bc = P.bcflag.astype(bool)
    uu[bc] = uu_star[bc]
    _, r_star, _ = P.residual(uu_star)
    _, r_0, _ = P.residual(uu)
    assert r_star <= 1e-12 * r_0                       # uu* is an exact discrete root
    prm = oracle.default_param(tol=1e-10)
    state = {"uu": uu}
    def step():
        A, b = P.assemble(state["uu"])
        dx, its = oracle.solver_dbsr_krylov_amg(P.IA, P.JA, A, b, prm)
        state["uu"] = state["uu"] + dx
        _, l2, mx = P.residual(state["uu"])
        return its, l2, mx
    n_it, rel = _newton(step, lambda: P.residual(state["uu"])[1:])
    assert n_it < 15 and rel < 1e-8
    assert np.abs(state["uu"] - uu_star).max() < 1e-7


@pytest.mark.gpu
def test_gpu_newton_recovers_manufactured_solution(oracle, pkg, gpu_lib):
    P = oracle.Problem(16, 6, 5)
    uu_star = _target(P)
    P.fixed, P.reac = manufactured_sources(oracle, P, uu_star)
    uu = P.uu0.copy()
    bc = P.bcflag.astype(bool)
    uu[bc] = uu_star[bc]
    g = pkg.PNP(P.xyz, P.cells, 0)
    g.set_coefficients(P.eps, P.fixed, P.diff, P.val, P.reac)
    g.set_dirichlet(np.nonzero(P.bcflag)[0].astype(np.int32))
    g.set_solution(uu_star)
    assert g.residual()[0] <= 1e-11 * P.residual(uu)[1]
    g.set_solution(uu)
    it, amg = pkg.default_params(tol=1e-10)
    n_it, rel = _newton(lambda: g.newton_step(it, amg), g.residual)
    assert n_it < 15 and rel < 1e-8
    assert np.abs(g.get_solution() - uu_star).max() < 1e-7
    g.close()
