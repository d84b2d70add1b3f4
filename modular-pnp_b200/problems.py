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
