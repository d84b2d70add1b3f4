"""Shared builders for the parity tests: the same seeded inputs go to the CPU oracle and to the
CUDA path (through the C ABI)."""
import numpy as np


def perturbed_state(P, seed=1234, amp=0.1, quantize=False):
    """uu = exact + amp*U(-1,1) per dof (SURVEY §8d kernel-level parity input). With
    quantize=True the potential is snapped to multiples of 1/64 so that every EAFE edge
    difference d is either 0 or >= 1/64: the reference formula d/(exp(d)-1) is then
    well-conditioned and 1e-12 is attainable across exp() implementations."""
    rng = np.random.default_rng(seed)
    uu = P.exact + amp * rng.uniform(-1.0, 1.0, P.N)
    if quantize:
        uu[0::3] = np.round(uu[0::3] * 64.0) / 64.0
    return uu


def gpu_problem(pkg, P, uu=None, device=0):
    g = pkg.PNP(P.xyz, P.cells, device)
    g.set_coefficients(P.eps, P.fixed, P.diff, P.val, P.reac)
    g.set_dirichlet(np.nonzero(P.bcflag)[0].astype(np.int32))
    g.set_solution(P.uu0 if uu is None else uu)
    return g


def row_scaled_error(IA, val_a, val_b):
    """max over blocks of |a-b| / max|row of b| for BSR(3) value arrays (per-entry
    relative error is meaningless where stiffness entries cancel)."""
    n = len(IA) - 1
    A = val_a.reshape(-1, 3, 3); B = val_b.reshape(-1, 3, 3)
    worst = 0.0
    rowmax = np.zeros((n, 3))
    cnt = np.diff(IA)
    rows = np.repeat(np.arange(n), cnt)
    absB = np.abs(B).max(axis=2)            # (nnzb, 3): per scalar row inside the block
    np.maximum.at(rowmax, rows, absB)
    err = np.abs(A - B).max(axis=2)         # (nnzb, 3)
    scale = rowmax[rows]
    scale[scale == 0] = 1.0
    return float((err / scale).max())


def bsr_to_csr(IA, JA, val):
    """expand BSR(3) to scalar CSR (all 9 entries of every block kept)"""
    n = len(IA) - 1
    cnt = np.diff(IA)
    sIA = np.zeros(3 * n + 1, dtype=np.int32)
    sIA[1:] = np.cumsum(np.repeat(3 * cnt, 3))
    sJA = np.zeros(9 * len(JA), dtype=np.int32)
    sval = np.zeros(9 * len(JA))
    V = val.reshape(-1, 3, 3)
    for i in range(n):
        p0, p1 = IA[i], IA[i + 1]
        cols = (3 * JA[p0:p1][:, None] + np.arange(3)[None, :]).ravel()
        for r in range(3):
            s = sIA[3 * i + r]
            sJA[s:s + len(cols)] = cols
            sval[s:s + len(cols)] = V[p0:p1, r, :].ravel()
    return sIA, sJA, sval


class GeneralProblem:
    """PNP problem on an arbitrary conforming tetrahedral mesh (oracle side): same fields as
    oracle.Problem but built from (xyz, cells) — the unstructured path an adaptively refined mesh
    takes (benchmarks/pnp_refine_exact/main.cpp:242-268)."""

    def __init__(self, O, xyz, cells):
        import ctypes as C
        L = O.lib()
        self.O = O
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64).ravel()
        self.cells = np.ascontiguousarray(cells, dtype=np.int32).ravel()
        self.nv, self.nc = len(self.xyz) // 3, len(self.cells) // 4
        self.N = 3 * self.nv
        self.eps = np.zeros(self.nv); self.fixed = np.zeros(self.nv)
        self.diff = np.zeros(self.N); self.val = np.zeros(self.N); self.reac = np.zeros(self.N)
        self.uu0 = np.zeros(self.N); self.exact = np.zeros(self.N)
        L.orc_exact_problem(self.nv, self.xyz, self.eps, self.fixed, self.diff, self.val, self.reac, self.uu0,
                            self.exact.ctypes.data)
        vflag = np.zeros(self.nv, dtype=np.uint8)
        L.orc_dirichlet_vertices(self.nv, self.xyz, self.nc, self.cells, 0, vflag)
        self.bcflag = np.repeat(vflag, 3).astype(np.uint8)
        self.IA = np.zeros(self.nv + 1, dtype=np.int32)
        self.nnzb = L.orc_block_pattern_count(self.nv, self.nc, self.cells, self.IA)
        self.JA = np.zeros(self.nnzb, dtype=np.int32)
        L.orc_block_pattern_fill(self.nv, self.nc, self.cells, self.IA, self.JA)

    def assemble(self, uu, use_eafe=True):
        A = np.zeros(9 * self.nnzb); b = np.zeros(self.N)
        self.O.lib().orc_assemble(self.nv, self.xyz, self.nc, self.cells, self.IA, self.JA, np.ascontiguousarray(uu),
                                  self.eps, self.fixed, self.diff, self.val, self.reac, self.bcflag,
                                  1 if use_eafe else 0, A, b)
        return A, b

    def residual(self, uu):
        import ctypes as C
        b = np.zeros(self.N); l2, li = C.c_double(), C.c_double()
        self.O.lib().orc_residual(self.nv, self.xyz, self.nc, self.cells, np.ascontiguousarray(uu), self.eps, self.fixed,
                                  self.diff, self.val, self.reac, self.bcflag, b, C.byref(l2), C.byref(li))
        return b, l2.value, li.value


def irregular_mesh(O, nx, ny, nz, seed=11, refine_fraction=0.25, jitter=0.15):
    """Kuhn box with jittered interior vertices and a random quarter of the cells split at their
    centroid (1 -> 4, conforming): irregular row lengths, vertex degrees and cell volumes."""
    P = O.Problem(nx, ny, nz)
    rng = np.random.default_rng(seed)
    X = P.xyz.reshape(-1, 3).copy()
    cells = P.cells.reshape(-1, 4).astype(np.int64)
    h = np.array([2.0 / nx, 0.4 / ny, 0.4 / nz])
    interior = np.all((np.abs(X) < np.array([1.0, 0.2, 0.2]) - 1e-12), axis=1)
    X[interior] += jitter * h * rng.uniform(-1, 1, (int(interior.sum()), 3))
    pick = rng.uniform(size=len(cells)) < refine_fraction
    keep = cells[~pick]
    split = cells[pick]
    cent = X[split].mean(axis=1)
    new_ids = len(X) + np.arange(len(split))
    X = np.vstack([X, cent])
    subs = []
    for drop in range(4):
        t = split.copy()
        t[:, drop] = new_ids
        subs.append(t)
    cells = np.vstack([keep] + subs)
    cells = np.sort(cells, axis=1)          # UFC ordering (ascending vertex ids)
    return np.ascontiguousarray(X.ravel()), np.ascontiguousarray(cells.astype(np.int32).ravel())


def manufactured_sources(O, P, uu_star):
    """Source-manufactured problem in the spirit of tests/pnp_tests/test_pnp_eafe.cpp:107-138: choose
    the P1 fields fixed_charge / reaction such that the given nodal vector uu_star is an EXACT root
    of the discrete residual L (forms.ufl:36-43): M fc = -r_phi(uu*), M reac_k = -r_k(uu*), with M the
    P1 mass matrix. Returns (fixed, reac)."""
    import ctypes as C
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    zero_nv, zero_N = np.zeros(P.nv), np.zeros(P.N)
    r = np.zeros(P.N); l2, li = C.c_double(), C.c_double()
    O.lib().orc_residual(P.nv, P.xyz, P.nc, P.cells, np.ascontiguousarray(uu_star), P.eps, zero_nv, P.diff, P.val, zero_N,
                         np.zeros(P.N, dtype=np.uint8), r, C.byref(l2), C.byref(li))
    X = P.xyz.reshape(-1, 3); cv = P.cells.reshape(-1, 4)
    vol = np.abs(np.linalg.det(X[cv][:, 1:] - X[cv][:, :1])) / 6
    Ml = (np.ones((4, 4)) + np.eye(4)) / 20.0
    rows = np.repeat(cv, 4, axis=1).ravel(); cols = np.tile(cv, (1, 4)).ravel()
    vals = (vol[:, None, None] * Ml[None]).ravel()
    M = sp.csr_matrix((vals, (rows, cols)), shape=(P.nv, P.nv))
    lu = spla.splu(M.tocsc())
    fixed = -lu.solve(r[0::3])
    reac = np.zeros(P.N)
    reac[1::3] = -lu.solve(r[1::3]); reac[2::3] = -lu.solve(r[2::3])
    return fixed, reac


class OracleNewtonBackend:
    """damped_newton backend (modular-pnp_b200/problems.py) on the CPU oracle: assemble + restated
    FASP BSR-AMG FGMRES for the update, residual-only evaluations for the trial iterates."""

    def __init__(self, O, P, prm, use_eafe=True):
        self.O, self.P, self.prm, self.use_eafe = O, P, prm, use_eafe
        self.n_dofs = P.N
        self.uu = None; self.base = None; self.du = None

    def set_solution(self, uu):
        self.uu = np.ascontiguousarray(uu, dtype=np.float64).copy()

    def residual(self):
        _, l2, mx = self.P.residual(self.uu)
        return l2, mx

    def solve(self):
        A, b = self.P.assemble(self.uu, self.use_eafe)
        dx, its = self.O.solver_dbsr_krylov_amg(self.P.IA, self.P.JA, A, b, self.prm)
        self.base = self.uu.copy(); self.du = dx
        return its

    def ranges(self):
        c = self.base + self.du
        return (self.base.min(), self.base.max()), (self.du.min(), self.du.max()), (c.min(), c.max())

    def trial(self, alpha):
        self.uu = self.base + alpha * self.du
        return self.residual()


def diode_problem(O, nx=40, ny=4, nz=4, voltage=0.1):
    """oracle.Problem carrying the pnp_diode physics (benchmarks/pnp_diode/diode.h, domain.dat 40x4x4
    on 2 x 0.2 x 0.2); the caller switches the form variant (poisson_scale, exp cutoff) on and off"""
    import sys
    from importlib import import_module
    from pnpb200_loader import load_package
    load_package()
    pr = import_module("modular_pnp_b200.problems")
    P = O.Problem(nx, ny, nz, 2.0, 0.2, 0.2)
    f = pr.diode_fields(P.xyz[0::3], voltage)
    P.eps, P.fixed, P.diff, P.val, P.reac, P.uu0 = f["eps"], f["fixed"], f["diff"], f["val"], f["reac"], f["uu0"]
    return P, pr


class GpuNewtonBackend:
    """damped_newton backend on the device-resident fast path of the C ABI: pnpb200_solve_update /
    pnpb200_update_ranges / pnpb200_trial_residual (the solution never returns to the host)."""

    def __init__(self, g, it, amg):
        self.g, self.it, self.amg = g, it, amg
        self.n_dofs = g.N

    def set_solution(self, uu):
        self.g.set_solution(np.ascontiguousarray(uu, dtype=np.float64))

    def residual(self):
        return self.g.residual()

    def solve(self):
        return self.g.solve_update(self.it, self.amg)

    def ranges(self):
        return self.g.update_ranges()

    def trial(self, alpha):
        return self.g.trial_residual(alpha)
