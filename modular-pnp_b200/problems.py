"""Problem data of the reference's benchmark drivers as flat numpy arrays (what the L5 mains
feed the hot path). pnp_exact_solution: benchmarks/pnp_exact_solution/main.cpp:24-135,293-302."""
import numpy as np

PI = 3.14159265359          # main.cpp:24 (truncated on purpose)


def box_vertices(nx, ny, nz, lx, ly, lz):
    """x-coordinates etc. of dolfin::BoxMesh(-L/2, +L/2) vertices in id order (src/domain.cpp:306-308)"""
    x = -lx / 2 + np.arange(nx + 1, dtype=np.float64) * (lx / 2 - (-lx / 2)) / nx
    y = -ly / 2 + np.arange(ny + 1, dtype=np.float64) * (ly / 2 - (-ly / 2)) / ny
    z = -lz / 2 + np.arange(nz + 1, dtype=np.float64) * (lz / 2 - (-lz / 2)) / nz
    return x, y, z


def exact_solution_problem(nx, ny, nz, lx=2.0, ly=0.4, lz=0.4):
    """Returns dict(eps, fixed, diff, val, reac, uu0, exact, bc_dofs) for the blocked dof order
    3*vertex + component on the box mesh (only x matters: the manufactured solution is 1-D)."""
    x1, _, _ = box_vertices(nx, ny, nz, lx, ly, lz)
    reps = (ny + 1) * (nz + 1)
    E, cref, perm = 1.0, 1.0, 1.0
    s = np.stack([x1 * E + np.sin(PI * x1), np.log(cref) - x1 * x1 * E, np.log(cref) + x1 * x1 * E], axis=1)
    ds = np.stack([E + PI * np.cos(PI * x1), -2.0 * x1 * E, 2.0 * x1 * E], axis=1)
    dds = np.stack([-(PI * PI) * np.sin(PI * x1), -2.0 * E * np.ones_like(x1), 2.0 * E * np.ones_like(x1)], axis=1)
    D = np.array([0.0, 1.0, 1.0])
    fixed1 = -perm * dds[:, 0] - (np.exp(s[:, 1]) - np.exp(s[:, 2]))
    reac1 = np.zeros((nx + 1, 3))
    reac1[:, 1] = -D[1] * np.exp(s[:, 1]) * (ds[:, 1] * (ds[:, 1] + ds[:, 0]) + (dds[:, 1] + dds[:, 0]))
    reac1[:, 2] = -D[2] * np.exp(s[:, 2]) * (ds[:, 2] * (ds[:, 2] - ds[:, 0]) + (dds[:, 2] - dds[:, 0]))
    left = np.array([-1.0 * E + np.sin(PI * -1.0), np.log(cref) - 1.0 * E, np.log(cref) + 1.0 * E])
    right = np.array([1.0 * E + np.sin(PI * 1.0), np.log(cref) - 1.0 * E, np.log(cref) + 1.0 * E])
    xmin, xmax = x1.min(), x1.max()
    dist = 1.0 / (xmax - xmin)
    init1 = left[None, :] * ((xmax - x1) * dist)[:, None]
    init1 = init1 + right[None, :] * ((x1 - xmin) * dist)[:, None]
    nv = (nx + 1) * reps
    tile = lambda a: np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1)))
    out = dict(
        eps=np.full(nv, perm), fixed=tile(fixed1),
        diff=np.ascontiguousarray(np.tile(D, nv)), val=np.ascontiguousarray(np.tile(np.array([0.0, 1.0, -1.0]), nv)),
        reac=tile(reac1).ravel(), uu0=tile(init1).ravel(), exact=tile(s).ravel())
    ix = np.arange(nv) % (nx + 1)
    bcv = np.nonzero((ix == 0) | (ix == nx))[0]
    out["bc_dofs"] = (3 * bcv[:, None] + np.arange(3)[None, :]).ravel().astype(np.int32)
    return out


def exact_solution_fields(x):
    """Same physics for an arbitrary list of vertex x-coordinates (one entry per local vertex of a
    partitioned mesh): returns dict(eps, fixed, diff, val, reac, uu0, exact) in dof order 3*v+k.
    The Dirichlet interpolant uses the global x-range [-1, 1] of the benchmark boxes."""
    x = np.asarray(x, dtype=np.float64)
    nv = x.shape[0]
    E, cref, perm = 1.0, 1.0, 1.0
    s = np.stack([x * E + np.sin(PI * x), np.log(cref) - x * x * E, np.log(cref) + x * x * E], axis=1)
    ds = np.stack([E + PI * np.cos(PI * x), -2.0 * x * E, 2.0 * x * E], axis=1)
    dds = np.stack([-(PI * PI) * np.sin(PI * x), -2.0 * E * np.ones_like(x), 2.0 * E * np.ones_like(x)], axis=1)
    D = np.array([0.0, 1.0, 1.0])
    reac = np.zeros((nv, 3))
    reac[:, 1] = -D[1] * np.exp(s[:, 1]) * (ds[:, 1] * (ds[:, 1] + ds[:, 0]) + (dds[:, 1] + dds[:, 0]))
    reac[:, 2] = -D[2] * np.exp(s[:, 2]) * (ds[:, 2] * (ds[:, 2] - ds[:, 0]) + (dds[:, 2] - dds[:, 0]))
    left = np.array([-1.0 * E + np.sin(PI * -1.0), np.log(cref) - 1.0 * E, np.log(cref) + 1.0 * E])
    right = np.array([1.0 * E + np.sin(PI * 1.0), np.log(cref) - 1.0 * E, np.log(cref) + 1.0 * E])
    xmin, xmax = -1.0, 1.0
    dist = 1.0 / (xmax - xmin)
    init = left[None, :] * ((xmax - x) * dist)[:, None]
    init = init + right[None, :] * ((x - xmin) * dist)[:, None]
    return dict(eps=np.full(nv, perm), fixed=-perm * dds[:, 0] - (np.exp(s[:, 1]) - np.exp(s[:, 2])),
                diff=np.ascontiguousarray(np.tile(D, nv)), val=np.ascontiguousarray(np.tile(np.array([0.0, 1.0, -1.0]), nv)),
                reac=np.ascontiguousarray(reac.ravel()), uu0=np.ascontiguousarray(init.ravel()), exact=np.ascontiguousarray(s.ravel()))


# ---- pnp_diode (BASELINE configs[1]): benchmarks/pnp_diode/diode.h ------------------------------
_Q = 1.60217662e-19; _KB = 1.38064852e-23; _T = 3e+2
DIODE_BETA = _Q / (_T * _KB)                       # thermodynamic_beta, diode.h:17
_EPS0 = 8.854187817e-12
_REF_LEN = 1e-5; _REF_DENS = 1.5e+22; _REF_DIFF = 2.87e+1; _REF_PERM = 1.17e+1
DIODE_PERMITTIVITY_FACTOR = _REF_PERM * _EPS0 / (_Q * DIODE_BETA * _REF_DENS * _REF_LEN * _REF_LEN)   # diode.h:25-26
DIODE_POISSON_SCALE = 1.0 / DIODE_PERMITTIVITY_FACTOR    # Poisson_Scale_Expression, diode.h:132-137 (the interpolated
                                                         # function overrides the constant of pnp_newton_solver.h:60)
_MAJ = 1.5e+22; _MIN = 6.6e+9


def diode_contacts(voltage_drop):
    """left_contact / right_contact (diode.h:66-79): (phi, log c1, log c2) at x = -1 and x = +1"""
    left = np.array([0.5 * voltage_drop * DIODE_BETA, np.log(_MIN / _REF_DENS), np.log(_MAJ / _REF_DENS)])
    right = np.array([-0.5 * voltage_drop * DIODE_BETA, np.log(_MAJ / _REF_DENS), np.log(_MIN / _REF_DENS)])
    return left, right


def diode_fields(x, voltage_drop):
    """Coefficients and initial guess of the diode benchmark for a list of vertex x-coordinates:
    dict(eps, fixed, diff, val, reac, uu0) in dof order 3*v+k (diode.h:36-64, Initial_Guess :84-116)."""
    x = np.asarray(x, dtype=np.float64)
    nv = x.shape[0]
    junction = np.abs(x) < 0.05
    ramp = 10.0 * (x + 0.05)
    d1 = np.where(x < 0.0, 1.07e+1, 1.09e+1); d2 = np.where(x < 0.0, 2.69e+1, 2.87e+1)
    d1 = np.where(junction, 1.07e+1 + ramp * (1.09e+1 - 1.07e+1), d1) / _REF_DIFF
    d2 = np.where(junction, 2.69e+1 + ramp * (2.87e+1 - 2.69e+1), d2) / _REF_DIFF
    fixed = np.where(x < 0.0, _MAJ, -_MAJ)
    fixed = np.where(junction, _MAJ + ramp * (-2.0 * _MAJ), fixed) / _REF_DENS
    left, right = diode_contacts(voltage_drop)
    lin = 0.5 * (left[None, :] * (1.0 - x)[:, None] + right[None, :] * (x + 1.0)[:, None])
    uu0 = lin.copy()
    for k in (1, 2):
        step = np.where(x < 0.0, left[k], right[k])
        step = np.where(junction, left[k] + ramp * (right[k] - left[k]), step)
        uu0[:, k] = 0.5 * lin[:, k] + 0.5 * step
    diff = np.zeros((nv, 3)); diff[:, 1] = d1; diff[:, 2] = d2
    return dict(eps=np.full(nv, 1.0), fixed=np.ascontiguousarray(fixed), diff=np.ascontiguousarray(diff.ravel()),
                val=np.ascontiguousarray(np.tile(np.array([0.0, 1.0, -1.0]), nv)), reac=np.zeros(3 * nv),
                uu0=np.ascontiguousarray(uu0.ravel()))



# ---- pnp_stokes (BASELINE configs[3]): synthetic 2x2 block system for the block solver --------
def mac_stokes_block(n, length=2.0, mu=1.0, pressure_penalty=1e-2):
    """Stokes saddle-point block [A_uu A_uq; A_qu A_qq] on an n^3 staggered (MAC) grid of a cube
    of edge `length`: face-normal velocities (like the reference's RT1 dofs: one per face), one
    pressure per cell (DG0), no-slip walls. A_uu = mu * (component-wise 7-point Laplacians, scaled
    by the cell volume; wall-normal faces are identity rows), A_qu = discrete divergence (face
    areas), A_uq = its transpose, A_qq = -pressure_penalty * volume / mu (removes the constant-
    pressure null space, like the penalty terms of vector_linear_pnp_ns_forms.ufl:38-40).
    Returns (A_ss as scipy CSR with velocity dofs first, n_vel, n_pres). Dof order: x-faces
    (i fastest, then j, k), y-faces, z-faces, then cells."""
    import scipy.sparse as sp
    h = length / n
    I = lambda m: sp.identity(m, format="csr")

    def t_normal(m):        # m+1 face positions along the face normal; the two walls are Dirichlet
        T = sp.diags([-np.ones(m), 2.0 * np.ones(m + 1), -np.ones(m)], [-1, 0, 1], format="lil")
        for w in (0, m):
            T[w, :] = 0.0; T[:, w] = 0.0
        return T.tocsr()

    def t_tangent(m):       # m cell-centred positions, wall by reflection (ghost = -inside)
        d = 2.0 * np.ones(m); d[0] = d[-1] = 3.0
        return sp.diags([-np.ones(m - 1), d, -np.ones(m - 1)], [-1, 0, 1], format="csr")

    def wall_mask(m):       # 1 on interior face positions
        w = np.ones(m + 1); w[0] = w[-1] = 0.0
        return w

    def d_normal(m):        # cells x faces difference along the normal, wall faces dropped
        D = sp.diags([-np.ones(m), np.ones(m)], [0, 1], shape=(m, m + 1), format="csr")
        return D @ sp.diags(wall_mask(m))

    kron3 = lambda az, ay, ax: sp.kron(az, sp.kron(ay, ax, format="csr"), format="csr")   # x fastest
    scale = mu * h                              # (1/h^2) * h^3
    Tn, Tt, In, It = t_normal(n), t_tangent(n), I(n + 1), I(n)
    Mn = sp.diags(wall_mask(n))
    wall_id = lambda az, ay, ax: kron3(az, ay, ax)
    # x-faces: (n+1) x n x n
    Lx = scale * (kron3(It, It, Tn) + kron3(It, Tt, Mn) + kron3(Tt, It, Mn)) + wall_id(It, It, In - Mn)
    Ly = scale * (kron3(It, Tn, It) + kron3(It, Mn, Tt) + kron3(Tt, Mn, It)) + wall_id(It, In - Mn, It)
    Lz = scale * (kron3(Tn, It, It) + kron3(Mn, It, Tt) + kron3(Mn, Tt, It)) + wall_id(In - Mn, It, It)
    Auu = sp.block_diag([Lx, Ly, Lz], format="csr")
    Dn = d_normal(n)
    B = (h * h) * sp.hstack([kron3(It, It, Dn), kron3(It, Dn, It), kron3(Dn, It, It)], format="csr")
    nq = n ** 3
    C = -(pressure_penalty * h ** 3 / mu) * sp.identity(nq, format="csr")
    Ass = sp.bmat([[Auu, -B.T], [-B, C]], format="csr")
    Ass.eliminate_zeros(); Ass.sort_indices()
    return Ass, Auu.shape[0], nq


def pnp_stokes_block_system(App, n, length=2.0, mu=1.0, coupling=0.05, seed=7):
    """Synthetic coupled system of the shape benchmarks/pnp_stokes/linear_pnp_ns.cpp:150-157 builds:
    App = the assembled PNP Jacobian on the n^3 box (scipy CSR, 3 dofs per vertex), A_ss = MAC
    Stokes block of the same box, A_ps = ion rows x velocity columns (the convective coupling
    (u . grad c) v couples a vertex with the faces of its cubes), A_sp = velocity rows x potential
    columns (the electric body force). Coupling entries: coupling * h * U(-1, 1), deterministic.
    Returns dict(blocks=[App, Aps, Asp, Ass], n_vel, n_pres, b, A) with A the assembled full
    matrix (for the direct-solve check of the tests)."""
    import scipy.sparse as sp
    h = length / n
    Ass, nu, nq = mac_stokes_block(n, length, mu)
    nvtx = (n + 1) ** 3
    assert App.shape == (3 * nvtx, 3 * nvtx)
    rng = np.random.default_rng(seed)
    ones = lambda m, k: sp.diags([np.ones(m - abs(k))], [k], shape=(m + 1, m), format="csr") if k <= 0 else None
    # vertex <-> cell incidence along one axis: vertex i touches cells i-1 and i
    v2c = lambda m: sp.diags([np.ones(m), np.ones(m)], [0, -1], shape=(m + 1, m), format="csr")
    # face <-> cell incidence along the face normal (face i between cells i-1 and i), identity tangentially
    kron3 = lambda az, ay, ax: sp.kron(az, sp.kron(ay, ax, format="csr"), format="csr")
    V2C = kron3(v2c(n), v2c(n), v2c(n))                               # vertices x cells
    F2C = sp.vstack([kron3(sp.identity(n), sp.identity(n), v2c(n)),   # x-faces x cells
                     kron3(sp.identity(n), v2c(n), sp.identity(n)),
                     kron3(v2c(n), sp.identity(n), sp.identity(n))], format="csr")
    adj = (V2C @ F2C.T).tocsr()                                        # vertices x faces pattern
    adj.data[:] = 1.0
    # interior faces only (wall faces carry a homogeneous Dirichlet velocity)
    interior = np.abs(Ass[:nu, :nu].diagonal() - 1.0) > 1e-14
    adj = (adj @ sp.diags(interior.astype(np.float64))).tocsr(); adj.eliminate_zeros()
    def coupling_block(rows_comp):      # 3*nvtx x nu, entries on component rows `rows_comp`
        blocks = []
        for k in range(3):
            if k in rows_comp:
                M = adj.copy(); M.data = coupling * h * rng.uniform(-1.0, 1.0, M.nnz)
            else:
                M = sp.csr_matrix(adj.shape)
            blocks.append(M)
        # interleave: row 3*v + k
        P = [sp.csr_matrix((np.ones(nvtx), (3 * np.arange(nvtx) + k, np.arange(nvtx))), shape=(3 * nvtx, nvtx)) for k in range(3)]
        return sum(P[k] @ blocks[k] for k in range(3)).tocsr()
    # Dirichlet rows of the PNP block stay identity rows: no coupling on them
    pdiag_only = np.asarray((App != 0).sum(axis=1)).ravel() == 1
    keep = sp.diags((~pdiag_only).astype(np.float64))
    Aps_u = (keep @ coupling_block({1, 2})).tocsr()
    Aps = sp.hstack([Aps_u, sp.csr_matrix((3 * nvtx, nq))], format="csr")
    Asp_u = (coupling_block({0}).T @ keep).tocsr()                    # velocity rows x PNP columns (potential, free dofs)
    Asp = sp.vstack([Asp_u, sp.csr_matrix((nq, 3 * nvtx))], format="csr")
    for M in (Aps, Asp):
        M.eliminate_zeros(); M.sort_indices()
    App = App.tocsr(); App.sort_indices()
    A = sp.bmat([[App, Aps], [Asp, Ass]], format="csr")
    xs = rng.uniform(-1.0, 1.0, A.shape[0])
    xs[3 * nvtx:3 * nvtx + nu][~interior] = 0.0
    b = A @ xs
    return dict(blocks=[App, Aps, Asp, Ass], n_vel=nu, n_pres=nq, b=b, A=A, x_star=xs)


class NewtonStatus:
    """src/newton_status.cpp:11-74"""

    def __init__(self, max_iterations, initial_residual, relative_tol, max_tol):
        self.iteration = 1; self.max_iterations = max_iterations
        self.initial_residual = initial_residual; self.relative_residual = 1.0
        self.max_residual = None; self.relative_tol = relative_tol; self.max_tol = max_tol

    def needs_to_iterate(self):
        return (self.iteration < self.max_iterations + 1 and self.relative_residual > self.relative_tol
                and self.max_residual > self.max_tol)

    def converged(self):
        return not (self.iteration > self.max_iterations) and (self.relative_residual < self.relative_tol
                                                               or self.max_residual < self.max_tol)


def damped_newton(backend, uu0, max_newton=250, rel_tol=1e-7, max_tol=1e-10, log=None):
    """The diode benchmark's Newton loop with growth limiting and residual backtracking
    (benchmarks/pnp_diode/pnp_newton_solver.h:171-284), on a backend that offers
        n_dofs, set_solution(uu), residual() -> (l2, max), solve() -> status (Newton update du from the
        current solution; the solution is NOT changed), ranges() -> (min/max of the solution, of du and
        of solution + du), trial(alpha) -> (l2, max) of the residual at solution_base + alpha * du.
    Every trial iterate is base + alpha * du, so the <= 50 backtracking evaluations per step are
    residual-only passes. Returns (history, converged): history rows (krylov status, alpha, relative
    residual, max residual, backtracks)."""
    N = float(backend.n_dofs)
    backend.set_solution(uu0)
    l2, mx = backend.residual()
    newton = NewtonStatus(max_newton, l2 / N, rel_tol, max_tol)
    newton.max_residual = mx
    hist, fails = [], 0
    while newton.needs_to_iterate():
        previous_residual = backend.residual()[0] / N
        status = backend.solve()
        if status < 0:
            fails += 1
            if fails > 10:
                break
        (pmin, pmax), (umin, umax), (cmin, cmax) = backend.ranges()
        alpha = 1.0
        residual_check = backend.trial(1.0)[0] / N          # the reference applies the full update first (:200, :217)
        prev_range = pmax - pmin + 1e-12
        if (cmax - cmin) > 1.1 * prev_range:                 # :229-252, increase_tolerance = 0.1
            alpha = 0.1 * prev_range / (umax - umin + 1e-12)
            if alpha * umax > alpha * umin + 0.1 * prev_range:
                alpha *= (alpha * umax - alpha * umin) * 0.1 * 1e-4
        k = 0
        while k < 50 and (residual_check > 1.00001 * previous_residual or residual_check != residual_check):
            k += 1
            residual_check = backend.trial(alpha * 0.5 ** k)[0] / N      # :256-268
        l2, mx = backend.residual()
        newton.relative_residual = (l2 / N) / newton.initial_residual
        newton.max_residual = mx
        newton.iteration += 1
        hist.append((status, alpha * 0.5 ** k if k else 1.0, newton.relative_residual, mx, k))
        if log:
            log(hist[-1])
    return hist, newton.converged()


def rt_interpolate_constant(xyz, cells_sorted, cell_facets, facet_cells, w):
    """Facet dofs of the lowest-order Raviart-Thomas interpolant of the constant field w on the basis of
    include/pnpb200/pnpb200_ns.h (phi_i = s_i (x - v_i) / detJ, s = (-1, +1, -1, +1)): the dof of facet i of a cell is
    s_i sign(detJ) 2 |F_i| (w . n_i), n_i the cell's outward unit normal. Returns (uu[n_facets], largest disagreement
    between the two cells of an interior facet) — the disagreement is zero when the cells are UFC-ordered."""
    X = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    cs = np.asarray(cells_sorted).reshape(-1, 4); cf = np.asarray(cell_facets).reshape(-1, 4)
    nf = np.asarray(facet_cells).reshape(-1, 2).shape[0]
    uu = np.full(nf, np.nan); mismatch = 0.0
    w = np.asarray(w, dtype=np.float64)
    for c in range(cs.shape[0]):
        v = X[cs[c]]
        detJ = np.linalg.det(v[1:] - v[0])
        for i in range(4):
            a, b, d = v[[j for j in range(4) if j != i]]
            n2 = np.cross(b - a, d - a)                    # |n2| = 2 |F|
            if np.dot(n2, a - v[i]) < 0:
                n2 = -n2
            val = (1.0 if i & 1 else -1.0) * np.sign(detJ) * np.dot(w, n2)
            f = cf[c, i]
            if np.isnan(uu[f]):
                uu[f] = val
            else:
                mismatch = max(mismatch, abs(uu[f] - val))
    return uu, mismatch


def pnp_stokes_problem(xyz, cells_sorted, cell_facets, facet_cells, lx):
    """Data of benchmarks/pnp_stokes/main.cpp:123-153 on a box of length lx in x, in the dof numbering of
    include/pnpb200/pnpb200_ns.h: parameters (permittivity 1, D 1 / 1, valencies +1 / -1, mu 0.1, penalties 1 / 1),
    initial guess (cation, anion, potential) linear in x from (0, -2.30258509299, 1) to (-2.30258509299, 0, -1)
    (main.cpp:147), Dirichlet dofs of the Newton update.
    Boundary conditions: the PNP components on the two x faces as Linear_PNP_NS::init_BC (linear_pnp_ns.cpp:287-304).
    The reference's velocity / pressure conditions (a DirichletBC on the single corner point, one on a facet marker of
    the DG0 space, :295-318) select no RT1 / DG0 dof with DOLFIN's topological search, which leaves the constant
    velocity fields in the kernel of the Newton matrix; this driver closes the system with the wall condition
    u.n = 0 (all boundary-facet velocity dofs; the initial velocity is 0 instead of (1, 0, 0)) and pins the pressure of
    cell 0."""
    X = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    nv = X.shape[0]
    fc = np.asarray(facet_cells).reshape(-1, 2)
    nf, nc = fc.shape[0], np.asarray(cells_sorted).reshape(-1, 4).shape[0]
    left = np.array([0.0, -2.30258509299, 1.0]); right = np.array([-2.30258509299, 0.0, -1.0])
    t = (X[:, 0] + lx / 2) / lx
    cc = (left[None, :] * (1 - t)[:, None] + right[None, :] * t[:, None]).reshape(-1)
    on_x = np.flatnonzero(np.abs(np.abs(X[:, 0]) - lx / 2) < 1e-5 * lx)
    bc_pnp = (3 * on_x[:, None] + np.arange(3)[None, :]).reshape(-1)
    bc_vel = 3 * nv + np.flatnonzero(fc[:, 1] < 0)
    bc_pres = np.array([3 * nv + nf])
    return dict(par=np.array([1.0, 1.0, 1.0, 1.0, -1.0, 0.1, 1.0, 1.0]), cc=cc, uu=np.zeros(nf), pp=np.zeros(nc),
                bc_dofs=np.concatenate([bc_pnp, bc_vel, bc_pres]).astype(np.int32),
                pnp_dofs=np.arange(3 * nv, dtype=np.int32), velocity_dofs=(3 * nv + np.arange(nf)).astype(np.int32),
                pressure_dofs=(3 * nv + nf + np.arange(nc)).astype(np.int32))
