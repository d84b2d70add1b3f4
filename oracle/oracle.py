"""TEST INFRASTRUCTURE ONLY — ctypes binding of the CPU oracle (oracle/liboracle.so).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product library (libpnpb200.so) never loads it.
See oracle/pnp_oracle.h for what each function restates (reference file:line).
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PNP_ORACLE_LIB: another build of the same sources (the sanitizer build, `make -C oracle san`)
LIB_PATH = os.environ.get("PNP_ORACLE_LIB") or os.path.join(_HERE, "liboracle.so")
REF_PATH = os.path.join(_HERE, "_ref", "libpnpref.so")

_f8 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i4 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u1 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


class SolverParam(C.Structure):
    _fields_ = [("tol", C.c_double), ("maxit", C.c_int), ("restart", C.c_int),
                ("max_levels", C.c_int), ("coarse_dof", C.c_int),
                ("strong_coupled", C.c_double), ("max_aggregation", C.c_int),
                ("tentative_smooth", C.c_double), ("amg_tol", C.c_double),
                ("presmooth", C.c_int), ("postsmooth", C.c_int), ("print_level", C.c_int),
                ("coarse_dense", C.c_int), ("ortho_cgs", C.c_int), ("cycle_type", C.c_int), ("kcycle_levels", C.c_int)]


class NewtonResult(C.Structure):
    _fields_ = [("newton_its", C.c_int), ("converged", C.c_int),
                ("initial_residual", C.c_double), ("initial_max_residual", C.c_double),
                ("rel_res", C.c_double * 32), ("max_res", C.c_double * 32),
                ("krylov_its", C.c_int * 32),
                ("t_assemble", C.c_double), ("t_solve", C.c_double), ("t_residual", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle` (or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        L.orc_eafe_element.argtypes = [_f8] * 6
        L.orc_jac_element.argtypes = [_f8] * 6
        L.orc_res_element.argtypes = [_f8] * 8
        L.orc_use_ref_kernels.argtypes = [C.c_int, C.c_char_p]
        L.orc_use_ref_kernels.restype = C.c_int
        L.orc_set_form_variant.argtypes = [C.c_double, C.c_int]
        L.orc_set_form_variant.restype = None
        L.orc_box_mesh_counts.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] * 2
        L.orc_box_mesh.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [_f8, _i4]
        L.orc_dirichlet_vertices.argtypes = [C.c_int, _f8, C.c_int, _i4, C.c_int, _u1]
        L.orc_dirichlet_vertices.restype = C.c_int
        L.orc_block_pattern_count.argtypes = [C.c_int, C.c_int, _i4, _i4]
        L.orc_block_pattern_count.restype = C.c_int
        L.orc_block_pattern_fill.argtypes = [C.c_int, C.c_int, _i4, _i4, _i4]
        L.orc_assemble.argtypes = [C.c_int, _f8, C.c_int, _i4, _i4, _i4] + [_f8] * 6 + [_u1, C.c_int, _f8, _f8]
        L.orc_residual.argtypes = [C.c_int, _f8, C.c_int, _i4] + [_f8] * 6 + [_u1, _f8,
                                                                             C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_eafe_matrix.argtypes = [C.c_int, _f8, C.c_int, _i4, _i4, _i4, _f8, _f8, _f8, _f8, _f8]
        L.orc_marker_element.argtypes = [_f8] * 4
        L.orc_marker_element.restype = C.c_double
        L.orc_entropy_indicator.argtypes = [C.c_int, _f8, C.c_int, _i4, _f8, _f8, _f8, _f8]
        L.orc_error_norms.argtypes = [C.c_int, _f8, C.c_int, _i4, _f8, _f8, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_ns_cell_a.argtypes = [_f8] * 5
        L.orc_ns_cell_L.argtypes = [_f8] * 6
        L.orc_ns_facet_a.argtypes = [_f8] * 4 + [C.c_int] * 2
        L.orc_ns_facet_L.argtypes = [_f8] * 5 + [C.c_int] * 2
        L.orc_ns_assemble.argtypes = [C.c_int, _f8, C.c_int, _i4, _i4, C.c_int, _i4, _f8, _f8, _f8, _f8, C.c_char_p,
                                      _i4, _i4, _f8, _f8]
        L.orc_ns_assemble.restype = C.c_long
        L.orc_csr_to_bsr3_count.argtypes = [C.c_int, _i4, _i4, _i4]
        L.orc_csr_to_bsr3_fill.argtypes = [C.c_int, _i4, _i4, _f8, _i4, _i4, _f8]
        L.orc_bsr_spmv.argtypes = [C.c_int, _i4, _i4, _f8, C.c_double, _f8, C.c_double, _f8]
        L.orc_bsr_diaginv.argtypes = [C.c_int, _i4, _i4, _f8, _f8]
        L.orc_bsr_gs.argtypes = [C.c_int, _i4, _i4, _f8, _f8, _f8, _f8, C.c_void_p, C.c_int]
        L.orc_vmb_aggregate.argtypes = [C.c_int, _i4, _i4, _f8, C.c_double, C.c_int, _i4]
        L.orc_vmb_aggregate.restype = C.c_int
        L.orc_rap_agg_count.argtypes = [C.c_int, _i4, _i4, _i4, C.c_int, _i4]
        L.orc_rap_agg_count.restype = C.c_int
        L.orc_rap_agg_fill.argtypes = [C.c_int, _i4, _i4, _f8, _i4, C.c_int, _i4, _i4, _f8]
        L.orc_default_param.argtypes = [C.POINTER(SolverParam)]
        L.orc_amg_setup.argtypes = [C.c_int, _i4, _i4, _f8, C.POINTER(SolverParam), C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_amg_setup.restype = C.c_void_p
        L.orc_amg_nlevels.argtypes = [C.c_void_p]
        L.orc_amg_nlevels.restype = C.c_int
        L.orc_amg_level_info.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3
        L.orc_amg_level_matrix.argtypes = [C.c_void_p, C.c_int, _i4, _i4, _f8]
        L.orc_amg_level_agg.argtypes = [C.c_void_p, C.c_int, _i4]
        L.orc_amg_vcycle.argtypes = [C.c_void_p, _f8, _f8]
        L.orc_amg_free.argtypes = [C.c_void_p]
        L.orc_pvfgmres.argtypes = [C.c_int, _i4, _i4, _f8, _f8, _f8, C.c_void_p, C.POINTER(SolverParam),
                                   C.c_void_p, C.c_int]
        L.orc_pvfgmres.restype = C.c_int
        L.orc_solver_dbsr_krylov_amg.argtypes = [C.c_int, _i4, _i4, _f8, _f8, _f8, C.POINTER(SolverParam)]
        L.orc_solver_dbsr_krylov_amg.restype = C.c_int
        L.orc_ilu_setup.argtypes = [C.c_int, _i4, _i4, _f8, C.c_int]
        L.orc_ilu_setup.restype = C.c_void_p
        L.orc_ilu_nnz.argtypes = [C.c_void_p]
        L.orc_ilu_factors.argtypes = [C.c_void_p, _i4, _i4, _f8, _f8]
        L.orc_ilu_apply.argtypes = [C.c_void_p, _f8, _f8]
        L.orc_ilu_free.argtypes = [C.c_void_p]
        L.orc_solver_dbsr_krylov_ilu.argtypes = [C.c_int, _i4, _i4, _f8, _f8, _f8, C.POINTER(SolverParam), C.c_int]
        L.orc_solver_dbsr_krylov_ilu.restype = C.c_int
        L.orc_exact_problem.argtypes = [C.c_int, _f8] + [_f8] * 6 + [C.c_void_p]
        L.orc_newton_exact.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [C.POINTER(SolverParam), C.c_int,
                                                                          C.c_double, C.c_double, C.c_int,
                                                                          C.c_void_p, C.POINTER(NewtonResult)]
        L.orc_newton_exact.restype = C.c_int
        _lib = L
    return _lib


def ref_available():
    return os.path.exists(REF_PATH)


REF_DIODE_PATH = os.path.join(_HERE, "_ref", "libpnpref_diode.so")


def set_form_variant(poisson_scale=1.0, conc_cutoff=False):
    lib().orc_set_form_variant(float(poisson_scale), 1 if conc_cutoff else 0)


def use_ref_kernels(on=True, diode=False):
    rc = lib().orc_use_ref_kernels(1 if on else 0, (REF_DIODE_PATH if diode else REF_PATH).encode())
    if rc != 0:
        raise RuntimeError("reference kernels unavailable (oracle/_ref/libpnpref.so)")


def default_param(**kw):
    p = SolverParam()
    lib().orc_default_param(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Problem:
    """Box-mesh PNP problem data (pnp_exact_solution physics unless overridden)."""

    def __init__(self, nx, ny, nz, lx=2.0, ly=0.4, lz=0.4):
        L = lib()
        nv, nc = C.c_int(), C.c_int()
        L.orc_box_mesh_counts(nx, ny, nz, C.byref(nv), C.byref(nc))
        self.nv, self.nc = nv.value, nc.value
        self.N = 3 * self.nv
        self.xyz = np.zeros(3 * self.nv)
        self.cells = np.zeros(4 * self.nc, dtype=np.int32)
        L.orc_box_mesh(nx, ny, nz, lx, ly, lz, self.xyz, self.cells)
        self.eps = np.zeros(self.nv); self.fixed = np.zeros(self.nv)
        self.diff = np.zeros(self.N); self.val = np.zeros(self.N); self.reac = np.zeros(self.N)
        self.uu0 = np.zeros(self.N); self.exact = np.zeros(self.N)
        L.orc_exact_problem(self.nv, self.xyz, self.eps, self.fixed, self.diff, self.val, self.reac,
                            self.uu0, self.exact.ctypes.data)
        vflag = np.zeros(self.nv, dtype=np.uint8)
        self.n_bc_vertices = L.orc_dirichlet_vertices(self.nv, self.xyz, self.nc, self.cells, 0, vflag)
        self.bcflag = np.repeat(vflag, 3).astype(np.uint8)
        self.IA = np.zeros(self.nv + 1, dtype=np.int32)
        self.nnzb = L.orc_block_pattern_count(self.nv, self.nc, self.cells, self.IA)
        self.JA = np.zeros(self.nnzb, dtype=np.int32)
        L.orc_block_pattern_fill(self.nv, self.nc, self.cells, self.IA, self.JA)

    def assemble(self, uu, use_eafe=True):
        A = np.zeros(9 * self.nnzb); b = np.zeros(self.N)
        lib().orc_assemble(self.nv, self.xyz, self.nc, self.cells, self.IA, self.JA, np.ascontiguousarray(uu),
                           self.eps, self.fixed, self.diff, self.val, self.reac, self.bcflag,
                           1 if use_eafe else 0, A, b)
        return A, b

    def error_norms(self, uu, exact=None):
        a, b = C.c_double(), C.c_double()
        lib().orc_error_norms(self.nv, self.xyz, self.nc, self.cells, np.ascontiguousarray(uu),
                              np.ascontiguousarray(self.exact if exact is None else exact), C.byref(a), C.byref(b))
        return a.value, b.value

    def entropy_indicator(self, uu):
        """Mesh_Refiner::compute_entropy_error_vector (src/mesh_refiner.cpp:212-271), one value per cell"""
        eta = np.zeros(self.nc)
        lib().orc_entropy_indicator(self.nv, self.xyz, self.nc, self.cells, np.ascontiguousarray(uu), self.diff, self.val, eta)
        return eta

    def residual(self, uu):
        b = np.zeros(self.N); l2, li = C.c_double(), C.c_double()
        lib().orc_residual(self.nv, self.xyz, self.nc, self.cells, np.ascontiguousarray(uu), self.eps,
                           self.fixed, self.diff, self.val, self.reac, self.bcflag, b, C.byref(l2), C.byref(li))
        return b, l2.value, li.value


def bsr_spmv(IA, JA, val, x, alpha=1.0, beta=0.0, y=None):
    n = len(IA) - 1
    y = np.zeros(3 * n) if y is None else np.ascontiguousarray(y, dtype=np.float64).copy()
    lib().orc_bsr_spmv(n, IA, JA, val, alpha, np.ascontiguousarray(x), beta, y)
    return y


def bsr_diaginv(IA, JA, val):
    n = len(IA) - 1
    d = np.zeros(9 * n)
    lib().orc_bsr_diaginv(n, IA, JA, val, d)
    return d


def bsr_gs(IA, JA, val, dinv, b, x, order=None):
    n = len(IA) - 1
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    if order is None:
        lib().orc_bsr_gs(n, IA, JA, val, dinv, b, x, None, 0)
    else:
        order = np.ascontiguousarray(order, dtype=np.int32)
        lib().orc_bsr_gs(n, IA, JA, val, dinv, b, x, order.ctypes.data, len(order))
    return x


def vmb_aggregate(IA, JA, val, strong=0.08, max_agg=20):
    n = len(IA) - 1
    agg = np.zeros(n, dtype=np.int32)
    na = lib().orc_vmb_aggregate(n, IA, JA, val, strong, max_agg, agg)
    return agg, na


def rap_agg(IA, JA, val, agg, nagg):
    n = len(IA) - 1
    agg = np.ascontiguousarray(agg, dtype=np.int32)
    cIA = np.zeros(nagg + 1, dtype=np.int32)
    nnz = lib().orc_rap_agg_count(n, IA, JA, agg, nagg, cIA)
    cJA = np.zeros(nnz, dtype=np.int32); cval = np.zeros(9 * nnz)
    lib().orc_rap_agg_fill(n, IA, JA, val, agg, nagg, cIA, cJA, cval)
    return cIA, cJA, cval


class AMG:
    def __init__(self, IA, JA, val, param, ext_agg=None, ext_nagg=None, ext_order=None):
        self.n = len(IA) - 1
        self._keep = []
        nl = 0; pa = pn = po = None
        if ext_agg is not None:
            nl = len(ext_agg) + 1
            aggs = [np.ascontiguousarray(a, dtype=np.int32) for a in ext_agg]
            self._keep += aggs
            arr = (C.c_void_p * nl)(*([a.ctypes.data for a in aggs] + [None]))
            nagg = np.ascontiguousarray(list(ext_nagg) + [0], dtype=np.int32)
            self._keep += [arr, nagg]
            pa, pn = C.cast(arr, C.c_void_p), nagg.ctypes.data
        if ext_order is not None:
            ords = [None if o is None else np.ascontiguousarray(o, dtype=np.int32) for o in ext_order]
            self._keep += ords
            nlev = nl if nl else len(ords)
            oarr = (C.c_void_p * max(nlev, 1))(*[None if o is None else o.ctypes.data for o in ords][:max(nlev, 1)])
            self._keep.append(oarr)
            po = C.cast(oarr, C.c_void_p)
        self.param = param
        self.h = lib().orc_amg_setup(self.n, IA, JA, val, C.byref(param), nl, pa, pn, po)

    def nlevels(self):
        return lib().orc_amg_nlevels(self.h)

    def level_info(self, l):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().orc_amg_level_info(self.h, l, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def level_matrix(self, l):
        n, nnz, _ = self.level_info(l)
        IA = np.zeros(n + 1, dtype=np.int32); JA = np.zeros(nnz, dtype=np.int32); val = np.zeros(9 * nnz)
        lib().orc_amg_level_matrix(self.h, l, IA, JA, val)
        return IA, JA, val

    def level_agg(self, l):
        n, _, na = self.level_info(l)
        agg = np.zeros(n, dtype=np.int32)
        lib().orc_amg_level_agg(self.h, l, agg)
        return agg, na

    def vcycle(self, r):
        z = np.zeros(3 * self.n)
        lib().orc_amg_vcycle(self.h, np.ascontiguousarray(r), z)
        return z

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_amg_free(self.h); self.h = None


def pvfgmres(IA, JA, val, b, param, amg=None, x0=None, hist_len=0):
    n = len(IA) - 1
    x = np.zeros(3 * n) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
    hist = np.zeros(max(hist_len, 1))
    its = lib().orc_pvfgmres(n, IA, JA, val, np.ascontiguousarray(b), x, amg.h if amg else None,
                             C.byref(param), hist.ctypes.data if hist_len else None, hist_len)
    return x, its, hist[:max(0, min(its, hist_len))]


def solver_dbsr_krylov_amg(IA, JA, val, b, param):
    n = len(IA) - 1
    x = np.zeros(3 * n)
    its = lib().orc_solver_dbsr_krylov_amg(n, IA, JA, val, np.ascontiguousarray(b), x, C.byref(param))
    return x, its


class ILU:
    """block ILU(k) of a BSR(3) matrix (oracle/ilu_restated.cpp): factors and z = U^{-1} L^{-1} r"""

    def __init__(self, IA, JA, val, lfil):
        self.n = len(IA) - 1
        self.h = lib().orc_ilu_setup(self.n, np.ascontiguousarray(IA, dtype=np.int32), np.ascontiguousarray(JA, dtype=np.int32),
                                     np.ascontiguousarray(val, dtype=np.float64), int(lfil))

    def factors(self):
        nnz = lib().orc_ilu_nnz(self.h)
        IA = np.zeros(self.n + 1, dtype=np.int32); JA = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(9 * nnz); dinv = np.zeros(9 * self.n)
        lib().orc_ilu_factors(self.h, IA, JA, val, dinv)
        return IA, JA, val, dinv

    def apply(self, r):
        z = np.zeros(3 * self.n)
        lib().orc_ilu_apply(self.h, np.ascontiguousarray(r, dtype=np.float64), z)
        return z

    def __del__(self):
        try:
            lib().orc_ilu_free(self.h)
        except Exception:
            pass


def solver_dbsr_krylov_ilu(IA, JA, val, b, param, lfil=2):
    n = len(IA) - 1
    x = np.zeros(3 * n)
    its = lib().orc_solver_dbsr_krylov_ilu(n, IA, JA, val, np.ascontiguousarray(b), x, C.byref(param), int(lfil))
    return x, its


def newton_exact(nx, ny, nz, lx=2.0, ly=0.4, lz=0.4, param=None, max_newton=15, rel_tol=1e-8,
                 max_tol=1e-10, use_eafe=True):
    param = param or default_param()
    res = NewtonResult()
    nv = (nx + 1) * (ny + 1) * (nz + 1)
    uu = np.zeros(3 * nv)
    lib().orc_newton_exact(nx, ny, nz, lx, ly, lz, C.byref(param), max_newton, rel_tol, max_tol,
                           1 if use_eafe else 0, uu.ctypes.data, C.byref(res))
    return uu, res


# ---- PNP-Stokes mixed system (benchmarks/pnp_stokes; oracle/ns_elements.cpp) --------------------------------------
REF_NS_PATH = os.path.join(_HERE, "_ref", "libpnpref_ns.so")
NS_PAR_DEFAULT = np.array([1.0, 1.0, 1.0, 1.0, -1.0, 0.1, 1.0, 1.0])      # benchmarks/pnp_stokes/main.cpp:123-132


def ns_topology(cells):
    """UFC ordering + facet numbering restated with numpy (dolfin::Mesh::order / init(2)): cells sorted ascending,
    facet i of a cell opposite its i-th vertex, facets numbered in lexicographic order of their vertex triples,
    facet_cells = (lower cell, higher cell or -1)."""
    cs = np.sort(np.asarray(cells, dtype=np.int64).reshape(-1, 4), axis=1)
    nc = cs.shape[0]
    tri = np.stack([np.delete(cs, i, axis=1) for i in range(4)], axis=1).reshape(-1, 3)      # row 4 c + i
    uniq, inv = np.unique(tri, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    nf = uniq.shape[0]
    cell_facets = inv.reshape(nc, 4).astype(np.int32)
    facet_cells = -np.ones((nf, 2), dtype=np.int32)
    for c in range(nc):
        for i in range(4):
            f = cell_facets[c, i]
            facet_cells[f, 0 if facet_cells[f, 0] < 0 else 1] = c
    return np.ascontiguousarray(cs.astype(np.int32)), cell_facets, facet_cells


def ns_assemble(xyz, cells_sorted, cell_facets, facet_cells, cc, uu, pp, par=None, bc=None, use_ref=False):
    """dolfin::assemble(a), assemble(L) + DirichletBC::apply restated for the mixed P1^3 x RT1 x DG0 system
    (src/pde.cpp:542-555); returns (scipy CSR, rhs). use_ref drives the reference's own compiled kernels."""
    import scipy.sparse as sp
    par = NS_PAR_DEFAULT if par is None else np.ascontiguousarray(par, dtype=np.float64)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64); nv = xyz.size // 3
    cells_sorted = np.ascontiguousarray(cells_sorted, dtype=np.int32); nc = cells_sorted.size // 4
    cell_facets = np.ascontiguousarray(cell_facets, dtype=np.int32)
    facet_cells = np.ascontiguousarray(facet_cells, dtype=np.int32); nf = facet_cells.size // 2
    nif = int((facet_cells.reshape(-1, 2)[:, 1] >= 0).sum())
    ndof = 3 * nv + nf + nc
    n = 289 * nc + 1156 * nif
    row = np.zeros(n, dtype=np.int32); col = np.zeros(n, dtype=np.int32); val = np.zeros(n); rhs = np.zeros(ndof)
    got = lib().orc_ns_assemble(nv, xyz.reshape(-1), nc, cells_sorted.reshape(-1), cell_facets.reshape(-1), nf,
                                facet_cells.reshape(-1), np.ascontiguousarray(cc, dtype=np.float64),
                                np.ascontiguousarray(uu, dtype=np.float64), np.ascontiguousarray(pp, dtype=np.float64), par,
                                REF_NS_PATH.encode() if use_ref else None, row, col, val, rhs)
    if got != n:
        raise RuntimeError("orc_ns_assemble failed (%d)" % got)
    A = sp.coo_matrix((val, (row, col)), shape=(ndof, ndof)).tocsr()
    A.sum_duplicates(); A.sort_indices()
    if bc is not None and len(bc):
        keep = np.ones(ndof); keep[bc] = 0.0
        A = (sp.diags(keep) @ A + sp.diags(1.0 - keep)).tocsr()
        rhs[bc] = 0.0
    return A, rhs
