"""ctypes binding of libpnpb200.so — the harness side of the C ABI declared in
include/pnpb200/pnpb200.h and include/pnpb200/fasp_compat.h (used by tests/ and bench.py;
C/C++ callers include the headers directly). There is no CPU fallback: if the CUDA library is
missing, loading fails loudly."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PNPB200_LIB") or os.path.join(_HERE, "libpnpb200.so")   # PNPB200_LIB: variant builds (development)

INT = C.c_int
SHORT = C.c_short
REAL = C.c_double


class dCSRmat(C.Structure):
    _fields_ = [("row", INT), ("col", INT), ("nnz", INT), ("IA", C.POINTER(INT)), ("JA", C.POINTER(INT)),
                ("val", C.POINTER(REAL))]


class dBSRmat(C.Structure):
    _fields_ = [("ROW", INT), ("COL", INT), ("NNZ", INT), ("nb", INT), ("storage_manner", INT),
                ("val", C.POINTER(REAL)), ("IA", C.POINTER(INT)), ("JA", C.POINTER(INT))]


class dvector(C.Structure):
    _fields_ = [("row", INT), ("val", C.POINTER(REAL))]


class input_param(C.Structure):
    _fields_ = [("print_level", SHORT), ("output_type", SHORT), ("inifile", C.c_char * 256), ("workdir", C.c_char * 256),
                ("problem_num", INT), ("solver_type", SHORT), ("precond_type", SHORT), ("stop_type", SHORT),
                ("itsolver_tol", REAL), ("itsolver_maxit", INT), ("restart", INT),
                ("ILU_type", SHORT), ("ILU_lfil", INT), ("ILU_droptol", REAL), ("ILU_relax", REAL), ("ILU_permtol", REAL),
                ("Schwarz_mmsize", INT), ("Schwarz_maxlvl", INT), ("Schwarz_type", INT), ("Schwarz_blksolver", INT),
                ("AMG_type", SHORT), ("AMG_levels", SHORT), ("AMG_cycle_type", SHORT), ("AMG_smoother", SHORT),
                ("AMG_smooth_order", SHORT), ("AMG_relaxation", REAL), ("AMG_polynomial_degree", SHORT),
                ("AMG_presmooth_iter", SHORT), ("AMG_postsmooth_iter", SHORT), ("AMG_coarse_dof", INT), ("AMG_tol", REAL),
                ("AMG_maxit", INT), ("AMG_ILU_levels", SHORT), ("AMG_coarse_solver", SHORT), ("AMG_coarse_scaling", SHORT),
                ("AMG_amli_degree", SHORT), ("AMG_nl_amli_krylov_type", SHORT), ("AMG_Schwarz_levels", INT),
                ("AMG_coarsening_type", SHORT), ("AMG_aggregation_type", SHORT), ("AMG_interpolation_type", SHORT),
                ("AMG_strong_threshold", REAL), ("AMG_truncation_threshold", REAL), ("AMG_max_row_sum", REAL),
                ("AMG_aggressive_level", INT), ("AMG_aggressive_path", INT), ("AMG_pair_number", INT),
                ("AMG_quality_bound", REAL), ("AMG_strong_coupled", REAL), ("AMG_max_aggregation", INT),
                ("AMG_tentative_smooth", REAL), ("AMG_smooth_filter", SHORT)]


class itsolver_param(C.Structure):
    _fields_ = [("itsolver_type", SHORT), ("precond_type", SHORT), ("stop_type", SHORT), ("maxit", INT), ("tol", REAL),
                ("restart", INT), ("print_level", SHORT)]


class AMG_param(C.Structure):
    _fields_ = [("AMG_type", SHORT), ("print_level", SHORT), ("maxit", INT), ("tol", REAL), ("max_levels", SHORT),
                ("coarse_dof", INT), ("cycle_type", SHORT), ("quality_bound", REAL), ("smoother", SHORT),
                ("smooth_order", SHORT), ("presmooth_iter", SHORT), ("postsmooth_iter", SHORT), ("relaxation", REAL),
                ("polynomial_degree", SHORT), ("coarse_solver", SHORT), ("coarse_scaling", SHORT), ("amli_degree", SHORT),
                ("amli_coef", C.POINTER(REAL)), ("nl_amli_krylov_type", SHORT), ("coarsening_type", SHORT),
                ("aggregation_type", SHORT), ("interpolation_type", SHORT), ("strong_threshold", REAL),
                ("max_row_sum", REAL), ("truncation_threshold", REAL), ("aggressive_level", INT), ("aggressive_path", INT),
                ("pair_number", INT), ("strong_coupled", REAL), ("max_aggregation", INT), ("tentative_smooth", REAL),
                ("smooth_filter", SHORT), ("ILU_levels", SHORT), ("ILU_type", SHORT), ("ILU_lfil", INT),
                ("ILU_droptol", REAL), ("ILU_relax", REAL), ("ILU_permtol", REAL), ("Schwarz_levels", INT),
                ("Schwarz_mmsize", INT), ("Schwarz_maxlvl", INT), ("Schwarz_type", INT), ("Schwarz_blksolver", INT)]


class ILU_param(C.Structure):
    _fields_ = [("print_level", SHORT), ("ILU_type", SHORT), ("ILU_lfil", INT), ("ILU_droptol", REAL),
                ("ILU_relax", REAL), ("ILU_permtol", REAL)]


class Schwarz_param(C.Structure):
    _fields_ = [("print_level", SHORT), ("Schwarz_type", SHORT), ("Schwarz_maxlvl", INT), ("Schwarz_mmsize", INT),
                ("Schwarz_blksolver", INT)]


class block_dCSRmat(C.Structure):
    _fields_ = [("brow", INT), ("bcol", INT), ("blocks", C.POINTER(C.POINTER(dCSRmat)))]


class input_ns_param(C.Structure):
    _fields_ = [("ns", input_param), ("v", input_param), ("p", input_param)]


class itsolver_ns_param(C.Structure):
    _fields_ = [("itsolver_type", SHORT), ("precond_type", SHORT), ("stop_type", SHORT), ("maxit", INT), ("tol", REAL),
                ("restart", INT), ("print_level", SHORT),
                ("itsolver_type_v", SHORT), ("precond_type_v", SHORT), ("pre_maxit_v", INT), ("pre_tol_v", REAL),
                ("pre_restart_v", INT),
                ("itsolver_type_p", SHORT), ("precond_type_p", SHORT), ("pre_maxit_p", INT), ("pre_tol_p", REAL),
                ("pre_restart_p", INT)]


class AMG_ns_param(C.Structure):
    _fields_ = [("param_v", AMG_param), ("param_p", AMG_param)]


class pnpb200_stats(C.Structure):
    _fields_ = [("n_levels", C.c_int), ("level_rows", C.c_int * 32), ("level_nnzb", C.c_int * 32),
                ("level_colors", C.c_int * 32), ("krylov_its", C.c_int), ("launches", C.c_int),
                ("ms_assemble", C.c_double), ("ms_setup", C.c_double), ("ms_krylov", C.c_double),
                ("ms_residual", C.c_double), ("ms_spmv", C.c_double), ("ms_vcycle", C.c_double), ("ms_ortho", C.c_double)]


_f8 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i4 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

_lib = None


def lib():
    """Load libpnpb200.so (built by __graft_entry__.build() / make -C modular-pnp_b200)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libpnpb200.so is missing (%s): build it with __graft_entry__.build(); "
                           "pnp-b200 has no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.pnpb200_version.restype = C.c_char_p
    L.pnpb200_device_count.restype = C.c_int
    L.pnpb200_create.argtypes = [C.POINTER(H), C.c_int, _f8, C.c_int, _i4, C.c_int]
    L.pnpb200_create_box.argtypes = [C.POINTER(H), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
    L.pnpb200_get_mesh.argtypes = [H, C.c_void_p, C.c_void_p]
    L.pnpb200_num_vertices.argtypes = [H]
    L.pnpb200_num_cells.argtypes = [H]
    L.pnpb200_probe_kernel.argtypes = [H, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.pnpb200_get_value_format.argtypes = [H, C.c_int]
    L.pnpb200_release_host_cache.argtypes = []
    L.pnpb200_release_host_cache.restype = None
    L.pnpb200_debug_repeat_pass.argtypes = [H, C.c_int, C.c_int, _i4]
    L.pnpb200_destroy.argtypes = [H]
    L.pnpb200_destroy.restype = None
    L.pnpb200_get_stream.argtypes = [H]
    L.pnpb200_get_stream.restype = C.c_void_p
    L.pnpb200_set_coefficients.argtypes = [H, _f8, _f8, _f8, _f8, _f8]
    L.pnpb200_set_dirichlet.argtypes = [H, C.c_int, _i4]
    L.pnpb200_set_solution.argtypes = [H, _f8]
    L.pnpb200_get_solution.argtypes = [H, _f8]
    L.pnpb200_set_solution_device.argtypes = [H, C.c_void_p]
    L.pnpb200_get_solution_device.argtypes = [H, C.c_void_p]
    L.pnpb200_use_eafe.argtypes = [H, C.c_int]
    L.pnpb200_set_form_variant.argtypes = [H, C.c_double, C.c_int]
    for f in ("pnpb200_num_dofs", "pnpb200_num_block_rows", "pnpb200_num_blocks"):
        getattr(L, f).argtypes = [H]
    L.pnpb200_assemble.argtypes = [H]
    L.pnpb200_get_matrix.argtypes = [H, C.POINTER(dBSRmat)]
    L.pnpb200_get_rhs.argtypes = [H, C.POINTER(dvector)]
    L.pnpb200_residual.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.pnpb200_solve.argtypes = [H, C.POINTER(itsolver_param), C.POINTER(AMG_param), C.POINTER(C.c_int)]
    L.pnpb200_get_update.argtypes = [H, _f8]
    L.pnpb200_test_solver.argtypes = [H, _f8, _f8, C.POINTER(itsolver_param), C.POINTER(AMG_param), C.POINTER(C.c_int)]
    L.pnpb200_newton_step.argtypes = [H, C.POINTER(itsolver_param), C.POINTER(AMG_param), C.POINTER(C.c_int),
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.pnpb200_set_hierarchy_reuse.argtypes = [H, C.c_int]
    L.pnpb200_solve_update.argtypes = [H, C.POINTER(itsolver_param), C.POINTER(AMG_param), C.POINTER(C.c_int)]
    L.pnpb200_update_ranges.argtypes = [H, _f8]
    L.pnpb200_trial_residual.argtypes = [H, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.pnpb200_set_sweep_shortcuts.argtypes = [H, C.c_int]
    L.pnpb200_error_norms.argtypes = [H, _f8, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.pnpb200_total_charge.argtypes = [H, _f8]
    L.pnpb200_entropy_indicator.argtypes = [H, _f8]
    L.pnpb200_mark_cells.argtypes = [H, C.c_double, C.c_long, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"),
                                     C.POINTER(C.c_long), C.POINTER(C.c_double)]
    L.pnpb200_get_stats.argtypes = [H, C.POINTER(pnpb200_stats)]
    L.pnpb200_set_timing_detail.argtypes = [H, C.c_int]
    L.pnpb200_get_level_matrix.argtypes = [H, C.c_int, C.POINTER(dBSRmat)]
    L.pnpb200_get_level_aggregates.argtypes = [H, C.c_int, _i4, C.POINTER(C.c_int)]
    L.pnpb200_get_level_order.argtypes = [H, C.c_int, _i4]
    L.pnpb200_apply_vcycle.argtypes = [H, _f8, _f8]
    L.pnpb200_apply_matrix.argtypes = [H, _f8, _f8]
    L.pnpb200_comm_unique_id.argtypes = [C.c_char_p]
    L.pnpb200_comm_init.argtypes = [H, C.c_char_p, C.c_int, C.c_int]
    L.pnpb200_set_partition.argtypes = [H, C.c_int, C.c_int, _i4, _i4, _i4, _i4]
    L.pnpb200_set_coloring.argtypes = [H, C.c_int, _i4]
    L.pnpb200_set_dof_map.argtypes = [H, C.c_void_p, C.c_void_p]
    L.pnpb200_set_preconditioner.argtypes = [H, C.c_int, C.c_int]
    L.pnpb200_get_ilu_factors.argtypes = [H, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pnpb200_apply_ilu.argtypes = [H, _f8, _f8]
    PP = C.POINTER(C.c_void_p)
    L.pnpb200_partition_box.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int] * 4 + [PP]
    L.pnpb200_partition_mesh.argtypes = [C.c_int, _f8, C.c_int, _i4, C.c_int, C.c_int, PP]
    L.pnpb200_partition_sizes.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 6
    L.pnpb200_partition_arrays.argtypes = [C.c_void_p] + [PP] * 9
    L.pnpb200_partition_free.argtypes = [C.c_void_p]
    L.pnpb200_partition_free.restype = None
    # FASP-compatible surface
    L.fasp_param_input.argtypes = [C.c_char_p, C.POINTER(input_param)]
    L.fasp_param_input.restype = None
    L.fasp_param_input_init.argtypes = [C.POINTER(input_param)]
    L.fasp_param_input_init.restype = None
    L.fasp_param_init.argtypes = [C.POINTER(input_param), C.POINTER(itsolver_param), C.POINTER(AMG_param),
                                  C.POINTER(ILU_param), C.c_void_p]
    L.fasp_param_init.restype = None
    L.fasp_param_amg_init.argtypes = [C.POINTER(AMG_param)]
    L.fasp_param_amg_init.restype = None
    L.fasp_param_solver_init.argtypes = [C.POINTER(itsolver_param)]
    L.fasp_param_solver_init.restype = None
    L.fasp_format_dcsr_dbsr.argtypes = [C.POINTER(dCSRmat), INT]
    L.fasp_format_dcsr_dbsr.restype = dBSRmat
    L.fasp_solver_dbsr_krylov_amg.argtypes = [C.POINTER(dBSRmat), C.POINTER(dvector), C.POINTER(dvector),
                                              C.POINTER(itsolver_param), C.POINTER(AMG_param)]
    L.fasp_solver_dbsr_krylov_amg.restype = INT
    L.fasp_solver_dbsr_krylov_ilu.argtypes = [C.POINTER(dBSRmat), C.POINTER(dvector), C.POINTER(dvector),
                                              C.POINTER(itsolver_param), C.POINTER(ILU_param)]
    L.fasp_solver_dbsr_krylov_ilu.restype = INT
    L.fasp_solver_dcsr_krylov.argtypes = [C.POINTER(dCSRmat), C.POINTER(dvector), C.POINTER(dvector),
                                          C.POINTER(itsolver_param)]
    L.fasp_solver_dcsr_krylov.restype = INT
    L.fasp_ns_param_input.argtypes = [C.c_char_p, C.POINTER(input_ns_param)]
    L.fasp_ns_param_input.restype = None
    L.fasp_ns_param_init.argtypes = [C.POINTER(input_ns_param), C.POINTER(itsolver_ns_param), C.POINTER(AMG_ns_param),
                                     C.POINTER(ILU_param), C.POINTER(Schwarz_param)]
    L.fasp_ns_param_init.restype = None
    L.fasp_solver_bdcsr_krylov_pnp_stokes.argtypes = [C.POINTER(block_dCSRmat), C.POINTER(dvector), C.POINTER(dvector),
                                                      C.POINTER(itsolver_param), C.POINTER(itsolver_param),
                                                      C.POINTER(AMG_param), C.POINTER(itsolver_ns_param),
                                                      C.POINTER(AMG_ns_param), INT, INT]
    L.fasp_solver_bdcsr_krylov_pnp_stokes.restype = INT
    L.pnpb200_last_block_solve_stats.argtypes = [C.POINTER(C.c_int * 5), C.POINTER(C.c_double)]
    L.fasp_dcsr_getblk.argtypes = [C.POINTER(dCSRmat), _i4, _i4, INT, INT, C.POINTER(dCSRmat)]
    L.fasp_dcsr_getblk.restype = SHORT
    L.fasp_dcsr_free.argtypes = [C.POINTER(dCSRmat)]
    L.fasp_dcsr_free.restype = None
    L.fasp_blas_dbsr_mxv.argtypes = [C.POINTER(dBSRmat), _f8, _f8]
    L.fasp_blas_dbsr_mxv.restype = None
    L.fasp_blas_dbsr_aAxpy.argtypes = [REAL, C.POINTER(dBSRmat), _f8, _f8]
    L.fasp_blas_dbsr_aAxpy.restype = None
    L.fasp_dbsr_free.argtypes = [C.POINTER(dBSRmat)]
    L.fasp_dbsr_free.restype = None
    L.fasp_dvec_alloc.argtypes = [INT, C.POINTER(dvector)]
    L.fasp_dvec_alloc.restype = None
    L.fasp_dvec_set.argtypes = [INT, C.POINTER(dvector), REAL]
    L.fasp_dvec_set.restype = None
    L.fasp_dvec_free.argtypes = [C.POINTER(dvector)]
    L.fasp_dvec_free.restype = None
    # include/pnpb200/pnpb200_ns.h
    L.pnpb200_ns_create.argtypes = [C.POINTER(H), C.c_int, _f8, C.c_int, _i4]
    L.pnpb200_ns_destroy.argtypes = [H]
    L.pnpb200_ns_destroy.restype = None
    L.pnpb200_ns_sizes.argtypes = [H] + [C.POINTER(C.c_int)] * 4
    L.pnpb200_ns_topology.argtypes = [H, _i4, _i4, _i4]
    L.pnpb200_ns_set_parameters.argtypes = [H, _f8]
    L.pnpb200_ns_set_dirichlet.argtypes = [H, C.c_int, _i4]
    L.pnpb200_ns_set_solution.argtypes = [H, _f8, _f8, _f8]
    L.pnpb200_ns_get_solution.argtypes = [H, _f8, _f8, _f8]
    L.pnpb200_ns_assemble.argtypes = [H]
    L.pnpb200_ns_get_system.argtypes = [H, C.POINTER(dCSRmat), C.POINTER(dvector)]
    L.pnpb200_ns_residual.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.pnpb200_ns_add_update.argtypes = [H, _f8, C.c_double]
    _lib = L
    return L


DEFAULT_BSR_DAT = os.path.join(_HERE, "host", "bsr.dat")


def default_params(dat_file=None, **overrides):
    """(itsolver_param, AMG_param) from a FASP .dat file — by default the repo's copy of the
    parameter VALUES of benchmarks/pnp_exact_solution/bsr.dat."""
    L = lib()
    inp, it, amg, ilu = input_param(), itsolver_param(), AMG_param(), ILU_param()
    L.fasp_param_input((dat_file or DEFAULT_BSR_DAT).encode(), C.byref(inp))
    L.fasp_param_init(C.byref(inp), C.byref(it), C.byref(amg), C.byref(ilu), None)
    for k, v in overrides.items():
        if hasattr(it, k):
            setattr(it, k, v)
        elif hasattr(amg, k):
            setattr(amg, k, v)
        else:
            raise KeyError(k)
    return it, amg


DEFAULT_BCSR_DAT = os.path.join(_HERE, "host", "bcsr.dat")
DEFAULT_NS_DAT = os.path.join(_HERE, "host", "ns.dat")


def pnp_stokes_params(bcsr_dat=None, bsr_dat=None, ns_dat=None):
    """The five parameter blocks of benchmarks/pnp_stokes/main.cpp:55-86: outer iteration
    (bcsr.dat), PNP block (bsr.dat), Stokes block (ns.dat)."""
    L = lib()
    inp, it, amg, ilu = input_param(), itsolver_param(), AMG_param(), ILU_param()
    L.fasp_param_input((bcsr_dat or DEFAULT_BCSR_DAT).encode(), C.byref(inp))
    L.fasp_param_init(C.byref(inp), C.byref(it), C.byref(amg), C.byref(ilu), None)
    it_pnp, amg_pnp = default_params(bsr_dat)
    ns_in, it_ns, amg_ns, ilu_ns, sch_ns = input_ns_param(), itsolver_ns_param(), AMG_ns_param(), ILU_param(), Schwarz_param()
    L.fasp_ns_param_input((ns_dat or DEFAULT_NS_DAT).encode(), C.byref(ns_in))
    L.fasp_ns_param_init(C.byref(ns_in), C.byref(it_ns), C.byref(amg_ns), C.byref(ilu_ns), C.byref(sch_ns))
    return it, it_pnp, amg_pnp, it_ns, amg_ns


def numpy_to_csr(IA, JA, val, ncol):
    """dCSRmat view over numpy arrays (arrays must outlive the struct)."""
    IA = np.ascontiguousarray(IA, dtype=np.int32); JA = np.ascontiguousarray(JA, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    m = dCSRmat()
    m.row = len(IA) - 1; m.col = int(ncol); m.nnz = len(JA)
    m.IA = IA.ctypes.data_as(C.POINTER(INT)); m.JA = JA.ctypes.data_as(C.POINTER(INT))
    m.val = val.ctypes.data_as(C.POINTER(REAL))
    m._keep = (IA, JA, val)
    return m


def solve_pnp_stokes(blocks, b, n_vel, n_pres, params=None, x0=None):
    """fasp_solver_bdcsr_krylov_pnp_stokes on four scipy-CSR-like blocks (objects with indptr,
    indices, data, shape) [A_pp, A_ps, A_sp, A_ss]; returns (status, x, stats dict)."""
    L = lib()
    it, it_pnp, amg_pnp, it_ns, amg_ns = params or pnp_stokes_params()
    mats = [numpy_to_csr(B.indptr, B.indices, B.data, B.shape[1]) for B in blocks]
    arr = (C.POINTER(dCSRmat) * 4)(*[C.pointer(m) for m in mats])
    A = block_dCSRmat(); A.brow = A.bcol = 2; A.blocks = C.cast(arr, C.POINTER(C.POINTER(dCSRmat)))
    x = np.zeros(len(b)) if x0 is None else np.array(x0, dtype=np.float64)
    bv, xv = numpy_to_dvec(b), numpy_to_dvec(x)
    st = L.fasp_solver_bdcsr_krylov_pnp_stokes(C.byref(A), C.byref(bv), C.byref(xv), C.byref(it), C.byref(it_pnp),
                                               C.byref(amg_pnp), C.byref(it_ns), C.byref(amg_ns), int(n_vel), int(n_pres))
    out, rr = (C.c_int * 5)(), C.c_double()
    L.pnpb200_last_block_solve_stats(C.byref(out), C.byref(rr))
    stats = {"outer_its": out[0], "pnp_solves": out[1], "pnp_its": out[2], "stokes_solves": out[3], "stokes_its": out[4],
             "relres": rr.value}
    vout = (C.c_int * 2)()
    L.pnpb200_last_block_velocity_stats(C.byref(vout))
    stats["velocity_solves"], stats["velocity_its"] = vout[0], vout[1]
    return st, xv._keep, stats


def _bsr_to_numpy(m, free=True):
    n, nnz = m.ROW, m.NNZ
    IA = np.ctypeslib.as_array(m.IA, shape=(n + 1,)).copy()
    JA = np.ctypeslib.as_array(m.JA, shape=(nnz,)).copy()
    val = np.ctypeslib.as_array(m.val, shape=(9 * nnz,)).copy()
    if free:
        lib().fasp_dbsr_free(C.byref(m))
    return IA.astype(np.int32), JA.astype(np.int32), val


def numpy_to_bsr(IA, JA, val):
    """dBSRmat view over numpy arrays (arrays must outlive the struct)."""
    IA = np.ascontiguousarray(IA, dtype=np.int32); JA = np.ascontiguousarray(JA, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    m = dBSRmat()
    m.ROW = m.COL = len(IA) - 1; m.NNZ = len(JA); m.nb = 3; m.storage_manner = 0
    m.IA = IA.ctypes.data_as(C.POINTER(INT)); m.JA = JA.ctypes.data_as(C.POINTER(INT))
    m.val = val.ctypes.data_as(C.POINTER(REAL))
    m._keep = (IA, JA, val)
    return m


def numpy_to_dvec(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    v = dvector(); v.row = len(x); v.val = x.ctypes.data_as(C.POINTER(REAL)); v._keep = x
    return v


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed with FASP status %d" % (what, rc))


class PNP:
    """Device-resident linearised-PNP problem: the Python mirror of the host-side
    Linear_PNP/PDE surface (modular-pnp_b200/host/linear_pnp.hpp)."""

    def __init__(self, xyz=None, cells=None, device=0, box=None):
        self.L = lib()
        self.h = C.c_void_p()
        if box is not None:
            nx, ny, nz, lx, ly, lz = box
            _check(self.L.pnpb200_create_box(C.byref(self.h), nx, ny, nz, lx, ly, lz, device), "pnpb200_create_box")
            self.nv, self.nc = self.L.pnpb200_num_vertices(self.h), self.L.pnpb200_num_cells(self.h)
        else:
            xyz = np.ascontiguousarray(xyz, dtype=np.float64).ravel()
            cells = np.ascontiguousarray(cells, dtype=np.int32).ravel()
            self.nv, self.nc = len(xyz) // 3, len(cells) // 4
            _check(self.L.pnpb200_create(C.byref(self.h), self.nv, xyz, self.nc, cells, device), "pnpb200_create")
        self.N = 3 * self.nv

    def get_mesh(self):
        xyz = np.zeros(3 * self.nv); cells = np.zeros(4 * self.nc, dtype=np.int32)
        _check(self.L.pnpb200_get_mesh(self.h, xyz.ctypes.data, cells.ctypes.data), "get_mesh")
        return xyz, cells

    def set_partition(self, part):
        nb = len(part["nbr"])
        pad = lambda a: np.ascontiguousarray(a if len(a) else np.zeros(1), dtype=np.int32)
        _check(self.L.pnpb200_set_partition(self.h, int(part["n_own"]), nb, pad(part["nbr"]), pad(part["send_ptr"]),
                                            pad(part["send_idx"]), pad(part["recv_ptr"])), "set_partition")

    def set_dof_map(self, vertex_to_dof=None, vertex_to_scalar_dof=None):
        a = None if vertex_to_dof is None else np.ascontiguousarray(vertex_to_dof, dtype=np.int32)
        b = None if vertex_to_scalar_dof is None else np.ascontiguousarray(vertex_to_scalar_dof, dtype=np.int32)
        _check(self.L.pnpb200_set_dof_map(self.h, None if a is None else a.ctypes.data, None if b is None else b.ctypes.data), "set_dof_map")

    def set_preconditioner(self, kind="amg", ilu_lfil=2):
        _check(self.L.pnpb200_set_preconditioner(self.h, {"amg": 0, "ilu": 1}[kind], int(ilu_lfil)), "set_preconditioner")

    def ilu_factors(self):
        nb = C.c_int()
        _check(self.L.pnpb200_get_ilu_factors(self.h, C.byref(nb), None, None, None, None), "get_ilu_factors")
        n = self.nv
        IA = np.zeros(n + 1, dtype=np.int32); JA = np.zeros(nb.value, dtype=np.int32)
        val = np.zeros(9 * nb.value); dinv = np.zeros(9 * n)
        _check(self.L.pnpb200_get_ilu_factors(self.h, C.byref(nb), IA.ctypes.data, JA.ctypes.data, val.ctypes.data, dinv.ctypes.data), "get_ilu_factors")
        return IA, JA, val, dinv

    def apply_ilu(self, r):
        z = np.zeros(self.N)
        _check(self.L.pnpb200_apply_ilu(self.h, np.ascontiguousarray(r, dtype=np.float64), z), "apply_ilu")
        return z

    def set_coloring(self, color, n_colors=None):
        color = np.ascontiguousarray(color, dtype=np.int32)
        _check(self.L.pnpb200_set_coloring(self.h, int(color.max()) + 1 if n_colors is None else int(n_colors), color), "set_coloring")

    def halo_counts(self):
        out = (C.c_longlong * 2)()
        self.L.pnpb200_comm_halo_counts(self.h, out)
        return int(out[0]), int(out[1])

    def comm_init(self, unique_id, rank, nranks):
        _check(self.L.pnpb200_comm_init(self.h, bytes(unique_id), rank, nranks), "comm_init")

    def probe_kernel(self, which, reps=10):
        ms, by = C.c_double(), C.c_double()
        _check(self.L.pnpb200_probe_kernel(self.h, which, reps, C.byref(ms), C.byref(by)), "probe_kernel")
        return ms.value, by.value

    def value_format(self, level=0):
        return self.L.pnpb200_get_value_format(self.h, level)

    def debug_repeat_pass(self, which, reps=100):
        out = np.zeros(18, dtype=np.int32)
        _check(self.L.pnpb200_debug_repeat_pass(self.h, which, reps, out), "debug_repeat_pass")
        return out

    def close(self):
        if self.h:
            self.L.pnpb200_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return self.L.pnpb200_get_stream(self.h)

    def set_coefficients(self, eps, fixed, diff, val, reac):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (eps, fixed, diff, val, reac)]
        _check(self.L.pnpb200_set_coefficients(self.h, *a), "set_coefficients")

    def set_dirichlet(self, dofs):
        dofs = np.ascontiguousarray(dofs, dtype=np.int32)
        _check(self.L.pnpb200_set_dirichlet(self.h, len(dofs), dofs), "set_dirichlet")

    def set_solution(self, uu):
        _check(self.L.pnpb200_set_solution(self.h, np.ascontiguousarray(uu, dtype=np.float64)), "set_solution")

    def get_solution(self):
        uu = np.zeros(self.N)
        _check(self.L.pnpb200_get_solution(self.h, uu), "get_solution")
        return uu

    def set_form_variant(self, poisson_scale=1.0, conc_cutoff=False):
        _check(self.L.pnpb200_set_form_variant(self.h, float(poisson_scale), 1 if conc_cutoff else 0), "set_form_variant")

    def use_eafe(self, on=True):
        _check(self.L.pnpb200_use_eafe(self.h, 1 if on else 0), "use_eafe")

    def assemble(self):
        _check(self.L.pnpb200_assemble(self.h), "assemble")

    def get_matrix(self):
        m = dBSRmat()
        _check(self.L.pnpb200_get_matrix(self.h, C.byref(m)), "get_matrix")
        return _bsr_to_numpy(m)

    def get_rhs(self):
        b = np.zeros(self.N)
        v = numpy_to_dvec(b)
        _check(self.L.pnpb200_get_rhs(self.h, C.byref(v)), "get_rhs")
        return v._keep

    def residual(self):
        a, b = C.c_double(), C.c_double()
        _check(self.L.pnpb200_residual(self.h, C.byref(a), C.byref(b)), "residual")
        return a.value, b.value

    def solve(self, it, amg):
        st = C.c_int()
        _check(self.L.pnpb200_solve(self.h, C.byref(it), C.byref(amg) if amg is not None else None, C.byref(st)), "solve")
        return st.value

    def get_update(self):
        du = np.zeros(self.N)
        _check(self.L.pnpb200_get_update(self.h, du), "get_update")
        return du

    def test_solver(self, target, it, amg):
        x = np.zeros(self.N); st = C.c_int()
        _check(self.L.pnpb200_test_solver(self.h, np.ascontiguousarray(target, dtype=np.float64), x, C.byref(it),
                                          C.byref(amg), C.byref(st)), "test_solver")
        return x, st.value

    def newton_step(self, it, amg):
        st, a, b = C.c_int(), C.c_double(), C.c_double()
        _check(self.L.pnpb200_newton_step(self.h, C.byref(it), C.byref(amg), C.byref(st), C.byref(a), C.byref(b)),
               "newton_step")
        return st.value, a.value, b.value

    def error_norms(self, exact):
        a, b = C.c_double(), C.c_double()
        _check(self.L.pnpb200_error_norms(self.h, np.ascontiguousarray(exact, dtype=np.float64), C.byref(a), C.byref(b)), "error_norms")
        return a.value, b.value

    def total_charge(self):
        out = np.zeros(self.nv)
        _check(self.L.pnpb200_total_charge(self.h, out), "total_charge")
        return out

    def entropy_indicator(self):
        """Mesh_Refiner::compute_entropy_error_vector (src/mesh_refiner.cpp:212-271): one value per cell"""
        eta = np.zeros(self.nc)
        _check(self.L.pnpb200_entropy_indicator(self.h, eta), "entropy_indicator")
        return eta

    def mark_cells(self, tol, target_size=0):
        """Mesh_Refiner::mark_for_refinement[_with_target_size]: (marker[nc] bool, count, tolerance used)"""
        m = np.zeros(self.nc, dtype=np.uint8); cnt = C.c_long(); used = C.c_double()
        _check(self.L.pnpb200_mark_cells(self.h, float(tol), int(target_size), m, C.byref(cnt), C.byref(used)), "mark_cells")
        return m.astype(bool), cnt.value, used.value

    def solve_update(self, it, amg=None):
        st = C.c_int()
        _check(self.L.pnpb200_solve_update(self.h, C.byref(it), C.byref(amg) if amg is not None else None, C.byref(st)), "solve_update")
        return st.value

    def update_ranges(self):
        r = np.zeros(6)
        _check(self.L.pnpb200_update_ranges(self.h, r), "update_ranges")
        return (r[0], r[1]), (r[2], r[3]), (r[4], r[5])

    def trial_residual(self, alpha):
        a, b = C.c_double(), C.c_double()
        _check(self.L.pnpb200_trial_residual(self.h, float(alpha), C.byref(a), C.byref(b)), "trial_residual")
        return a.value, b.value

    def set_hierarchy_reuse(self, on):
        _check(self.L.pnpb200_set_hierarchy_reuse(self.h, 1 if on else 0), "set_hierarchy_reuse")

    def set_sweep_shortcuts(self, on):
        _check(self.L.pnpb200_set_sweep_shortcuts(self.h, 1 if on else 0), "set_sweep_shortcuts")

    def set_timing_detail(self, on):
        _check(self.L.pnpb200_set_timing_detail(self.h, 1 if on else 0), "set_timing_detail")

    def stats(self):
        s = pnpb200_stats()
        _check(self.L.pnpb200_get_stats(self.h, C.byref(s)), "get_stats")
        return s

    def level_matrix(self, l):
        m = dBSRmat()
        _check(self.L.pnpb200_get_level_matrix(self.h, l, C.byref(m)), "get_level_matrix")
        return _bsr_to_numpy(m)

    def level_aggregates(self, l):
        n = self.stats().level_rows[l]
        agg = np.zeros(n, dtype=np.int32); na = C.c_int()
        _check(self.L.pnpb200_get_level_aggregates(self.h, l, agg, C.byref(na)), "get_level_aggregates")
        return agg, na.value

    def level_order(self, l):
        n = self.stats().level_rows[l]
        o = np.zeros(n, dtype=np.int32)
        _check(self.L.pnpb200_get_level_order(self.h, l, o), "get_level_order")
        return o

    def apply_vcycle(self, r):
        z = np.zeros(self.N)
        _check(self.L.pnpb200_apply_vcycle(self.h, np.ascontiguousarray(r, dtype=np.float64), z), "apply_vcycle")
        return z

    def apply_matrix(self, x):
        y = np.zeros(self.N)
        _check(self.L.pnpb200_apply_matrix(self.h, np.ascontiguousarray(x, dtype=np.float64), y), "apply_matrix")
        return y


def _partition_to_dict(L, p):
    """copy the arrays of a pnpb200_partition object into the dict layout of partition.py, then free it"""
    try:
        sz = [C.c_int() for _ in range(6)]
        _check(L.pnpb200_partition_sizes(p, *[C.byref(v) for v in sz]), "partition_sizes")
        n_own, n_loc, n_cells, n_nbr, n_send, n_bc = [v.value for v in sz]
        ptr = [C.c_void_p() for _ in range(9)]
        _check(L.pnpb200_partition_arrays(p, *[C.byref(v) for v in ptr]), "partition_arrays")

        def arr(v, ctype, dtype, count):
            if count == 0 or not v.value:
                return np.zeros(0, dtype=dtype)
            return np.ctypeslib.as_array(C.cast(v, C.POINTER(ctype)), shape=(count,)).astype(dtype, copy=True)
        d = dict(n_own=n_own, n_loc=n_loc,
                 xyz=arr(ptr[0], C.c_double, np.float64, 3 * n_loc), cells=arr(ptr[1], C.c_int, np.int32, 4 * n_cells),
                 gids=arr(ptr[2], C.c_longlong, np.int64, n_loc), nbr=arr(ptr[3], C.c_int, np.int32, n_nbr),
                 send_ptr=arr(ptr[4], C.c_int, np.int32, n_nbr + 1), send_idx=arr(ptr[5], C.c_int, np.int32, n_send),
                 recv_ptr=arr(ptr[6], C.c_int, np.int32, n_nbr + 1), bc_vertices=arr(ptr[8], C.c_int, np.int32, n_bc))
        col = arr(ptr[7], C.c_int, np.int32, n_loc)
        if len(col):
            d["color"] = col
        return d
    finally:
        L.pnpb200_partition_free(p)


def partition_box(nx, ny, nz, lx, ly, lz, px, py, pz, rank):
    """pnpb200_partition_box (csrc/partition.cpp): bricks of dolfin::BoxMesh, closed form, host only"""
    L = lib()
    p = C.c_void_p()
    _check(L.pnpb200_partition_box(nx, ny, nz, lx, ly, lz, px, py, pz, rank, C.byref(p)), "partition_box")
    return _partition_to_dict(L, p)


def partition_mesh(xyz, cells, nranks, rank):
    """pnpb200_partition_mesh: recursive coordinate bisection of any conforming tetrahedral mesh, host only"""
    L = lib()
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).ravel(); cells = np.ascontiguousarray(cells, dtype=np.int32).ravel()
    p = C.c_void_p()
    _check(L.pnpb200_partition_mesh(len(xyz) // 3, xyz, len(cells) // 4, cells, nranks, rank, C.byref(p)), "partition_mesh")
    return _partition_to_dict(L, p)


def comm_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().pnpb200_comm_unique_id(buf), "comm_unique_id")
    return buf.raw


def fasp_solver_dbsr_krylov_amg(IA, JA, val, b, it, amg, x0=None):
    """Reference-facing call: host arrays in, GPU solve, host vector out."""
    A = numpy_to_bsr(IA, JA, val)
    bv = numpy_to_dvec(b)
    x = np.zeros(len(b)) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
    xv = numpy_to_dvec(x)
    st = lib().fasp_solver_dbsr_krylov_amg(C.byref(A), C.byref(bv), C.byref(xv), C.byref(it), C.byref(amg))
    return xv._keep, st


def fasp_solver_dcsr_krylov(IA, JA, val, b, it):
    """tests/eafe_tests/test_eafe.cpp:170 entry: scalar CSR (n % 3 == 0), unpreconditioned FGMRES on the GPU"""
    IA = np.ascontiguousarray(IA, dtype=np.int32); JA = np.ascontiguousarray(JA, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    A = dCSRmat(); A.row = A.col = len(IA) - 1; A.nnz = len(JA)
    A.IA = IA.ctypes.data_as(C.POINTER(INT)); A.JA = JA.ctypes.data_as(C.POINTER(INT)); A.val = val.ctypes.data_as(C.POINTER(REAL))
    bv = numpy_to_dvec(b); xv = numpy_to_dvec(np.zeros(len(b)))
    st = lib().fasp_solver_dcsr_krylov(C.byref(A), C.byref(bv), C.byref(xv), C.byref(it))
    return xv._keep, st


def fasp_solver_dbsr_krylov_ilu(IA, JA, val, b, it, lfil=2):
    A = numpy_to_bsr(IA, JA, val)
    bv = numpy_to_dvec(b); xv = numpy_to_dvec(np.zeros(len(b)))
    ilu = ILU_param()
    ilu.ILU_type = 2; ilu.ILU_lfil = lfil          # benchmarks/pnp_diode/bsr.dat:34-38
    st = lib().fasp_solver_dbsr_krylov_ilu(C.byref(A), C.byref(bv), C.byref(xv), C.byref(it), C.byref(ilu))
    return xv._keep, st


def fasp_format_dcsr_dbsr(IA, JA, val, nb=3):
    IA = np.ascontiguousarray(IA, dtype=np.int32); JA = np.ascontiguousarray(JA, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    A = dCSRmat(); A.row = A.col = len(IA) - 1; A.nnz = len(JA)
    A.IA = IA.ctypes.data_as(C.POINTER(INT)); A.JA = JA.ctypes.data_as(C.POINTER(INT))
    A.val = val.ctypes.data_as(C.POINTER(REAL))
    B = lib().fasp_format_dcsr_dbsr(C.byref(A), nb)
    if not B.IA:
        raise RuntimeError("fasp_format_dcsr_dbsr failed")
    n, nnz = B.ROW, B.NNZ
    out = (np.ctypeslib.as_array(B.IA, shape=(n + 1,)).copy(), np.ctypeslib.as_array(B.JA, shape=(nnz,)).copy(),
           np.ctypeslib.as_array(B.val, shape=(nb * nb * nnz,)).copy())
    lib().fasp_dbsr_free(C.byref(B))
    return out


class PNPStokes:
    """Device assembly of the PNP-Stokes Newton system (include/pnpb200/pnpb200_ns.h): mixed space P1^3 x RT1 x DG0,
    dofs 3 v + k | 3 nv + facet | 3 nv + n_facets + cell (benchmarks/pnp_stokes/vector_linear_pnp_ns_forms.ufl)."""

    def __init__(self, xyz, cells):
        L = lib()
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1)
        cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1)
        self.nv, self.nc = xyz.size // 3, cells.size // 4
        self.h = C.c_void_p()
        _check(L.pnpb200_ns_create(C.byref(self.h), self.nv, xyz, self.nc, cells), "pnpb200_ns_create")
        nf, nif, nd, nnz = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.pnpb200_ns_sizes(self.h, C.byref(nf), C.byref(nif), C.byref(nd), C.byref(nnz))
        self.nf, self.n_interior, self.ndof, self.nnz = nf.value, nif.value, nd.value, nnz.value

    def close(self):
        if getattr(self, "h", None):
            lib().pnpb200_ns_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def topology(self):
        cs = np.zeros(4 * self.nc, dtype=np.int32); cf = np.zeros(4 * self.nc, dtype=np.int32); fc = np.zeros(2 * self.nf, dtype=np.int32)
        _check(lib().pnpb200_ns_topology(self.h, cs, cf, fc), "pnpb200_ns_topology")
        return cs.reshape(-1, 4), cf.reshape(-1, 4), fc.reshape(-1, 2)

    def set_parameters(self, par8):
        _check(lib().pnpb200_ns_set_parameters(self.h, np.ascontiguousarray(par8, dtype=np.float64)), "pnpb200_ns_set_parameters")

    def set_dirichlet(self, dofs):
        d = np.ascontiguousarray(dofs, dtype=np.int32)
        _check(lib().pnpb200_ns_set_dirichlet(self.h, len(d), d), "pnpb200_ns_set_dirichlet")

    def set_solution(self, cc, uu, pp):
        _check(lib().pnpb200_ns_set_solution(self.h, np.ascontiguousarray(cc, dtype=np.float64), np.ascontiguousarray(uu, dtype=np.float64),
                                             np.ascontiguousarray(pp, dtype=np.float64)), "pnpb200_ns_set_solution")

    def get_solution(self):
        cc, uu, pp = np.zeros(3 * self.nv), np.zeros(self.nf), np.zeros(self.nc)
        _check(lib().pnpb200_ns_get_solution(self.h, cc, uu, pp), "pnpb200_ns_get_solution")
        return cc, uu, pp

    def assemble(self):
        _check(lib().pnpb200_ns_assemble(self.h), "pnpb200_ns_assemble")

    def get_system(self):
        """(IA, JA, val, b) of the mixed system as numpy arrays"""
        IA = np.zeros(self.ndof + 1, dtype=np.int32); JA = np.zeros(self.nnz, dtype=np.int32); val = np.zeros(self.nnz)
        b = np.zeros(self.ndof)
        A = numpy_to_csr(IA, JA, val, self.ndof); bv = numpy_to_dvec(b)
        A._keep = (IA, JA, val); A.IA = IA.ctypes.data_as(C.POINTER(INT)); A.JA = JA.ctypes.data_as(C.POINTER(INT))
        A.val = val.ctypes.data_as(C.POINTER(REAL)); bv.val = b.ctypes.data_as(C.POINTER(REAL))
        _check(lib().pnpb200_ns_get_system(self.h, C.byref(A), C.byref(bv)), "pnpb200_ns_get_system")
        return IA, JA, val, b

    def residual(self):
        l2, li = C.c_double(), C.c_double()
        _check(lib().pnpb200_ns_residual(self.h, C.byref(l2), C.byref(li)), "pnpb200_ns_residual")
        return l2.value, li.value

    def add_update(self, du, alpha=1.0):
        _check(lib().pnpb200_ns_add_update(self.h, np.ascontiguousarray(du, dtype=np.float64), float(alpha)), "pnpb200_ns_add_update")
