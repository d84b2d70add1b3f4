"""-m "not gpu": the C-ABI library loads, exports every symbol include/pnpb200/*.h declares, its
host-side FASP utilities behave like the reference expects, and — on a box without a GPU — the
compute entry points fail loudly with ERROR_DEVICE instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    names = []
    for hdr in ("pnpb200.h", "pnpb200_ns.h", "fasp_compat.h"):
        src = open(os.path.join(ROOT, "include", "pnpb200", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"typedef struct[^;{]*\{.*?\}[^;]*;", "", src, flags=re.S)
        for m in re.finditer(r"\b((?:fasp|pnpb200)_\w+)\s*\(", src):
            names.append(m.group(1))
    return sorted(set(names))


def test_every_declared_symbol_is_exported(pkg):
    L = pkg.lib()
    names = _declared_functions()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, "declared in include/ but not exported: %s" % missing


def test_param_input_reads_the_reference_values(pkg):
    it, amg = pkg.default_params()
    assert (it.itsolver_type, it.maxit, it.restart, it.stop_type) == (6, 100, 30, 1) and abs(it.tol - 1e-6) < 1e-20
    assert (amg.AMG_type, amg.cycle_type, amg.max_levels, amg.coarse_dof) == (3, 1, 20, 100)
    assert (amg.smoother, amg.presmooth_iter, amg.postsmooth_iter, amg.aggregation_type) == (2, 1, 1, 2)
    assert abs(amg.strong_coupled - 0.08) < 1e-15 and amg.max_aggregation == 20 and abs(amg.tentative_smooth - 0.67) < 1e-15


def test_host_vector_helpers(pkg):
    L = pkg.lib()
    v = pkg.capi.dvector()
    L.fasp_dvec_alloc(7, C.byref(v))
    assert v.row == 7 and all(v.val[i] == 0.0 for i in range(7))
    L.fasp_dvec_set(7, C.byref(v), 2.5)
    assert all(v.val[i] == 2.5 for i in range(7))
    L.fasp_dvec_free(C.byref(v))
    assert v.row == 0 and not v.val


def test_no_cpu_fallback(pkg):
    L = pkg.lib()
    if L.pnpb200_device_count() > 0:
        return      # on a GPU box the -m gpu tests cover the compute path
    h = C.c_void_p()
    xyz = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1.0]); cells = np.array([0, 1, 2, 3], dtype=np.int32)
    rc = L.pnpb200_create(C.byref(h), 4, xyz, 1, cells, 0)
    assert rc == -120 and not h.value          # ERROR_DEVICE: fails loudly, no fallback
    IA = np.array([0, 1], dtype=np.int32); JA = np.array([0], dtype=np.int32); val = np.eye(3).ravel()
    it, amg = pkg.default_params()
    x, st = pkg.capi.fasp_solver_dbsr_krylov_amg(IA, JA, val, np.ones(3), it, amg)
    assert st == -120


def test_problem_generator_matches_oracle(oracle, pkg):
    from importlib import import_module
    pr = import_module("modular_pnp_b200.problems").exact_solution_problem(10, 4, 3)
    P = oracle.Problem(10, 4, 3)
    for k in ("eps", "fixed", "diff", "val", "reac", "uu0", "exact"):
        assert np.abs(getattr(P, k) - pr[k]).max() <= 1e-14
    assert np.array_equal(np.nonzero(P.bcflag)[0], pr["bc_dofs"])
