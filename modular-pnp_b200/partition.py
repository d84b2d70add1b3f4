"""Row (vertex) partition of the box meshes across ranks — host-side plumbing of SURVEY §8e.

Slabs of vertex planes along z (the slowest index of dolfin::BoxMesh numbering, so every slab is a
contiguous range of global vertex ids). Local numbering: owned vertices first (natural order),
then the ghost plane of the lower neighbour, then the ghost plane of the upper neighbour. Local
cells = every cell that touches an owned vertex (ghost cells are duplicated, so assembly needs no
communication beyond the halo of the solution vector)."""
import numpy as np


def plane_split(nz, nranks):
    planes = nz + 1
    return [int(round(r * planes / nranks)) for r in range(nranks + 1)]


def box_slab_partition(nx, ny, nz, lx, ly, lz, rank, nranks):
    p = plane_split(nz, nranks)
    z0, z1 = p[rank], p[rank + 1]
    if z1 - z0 < 1:
        raise ValueError("every rank needs at least one vertex plane")
    ps = (nx + 1) * (ny + 1)
    has_lo, has_hi = rank > 0, rank < nranks - 1
    n_own = (z1 - z0) * ps
    n_lo = ps if has_lo else 0
    n_hi = ps if has_hi else 0
    n_loc = n_own + n_lo + n_hi
    # global ids of the local vertices, in local order
    gids = [np.arange(z0 * ps, z1 * ps, dtype=np.int64)]
    if has_lo:
        gids.append(np.arange((z0 - 1) * ps, z0 * ps, dtype=np.int64))
    if has_hi:
        gids.append(np.arange(z1 * ps, (z1 + 1) * ps, dtype=np.int64))
    gids = np.concatenate(gids)
    # coordinates (dolfin::BoxMesh formula, src/domain.cpp:306-308)
    ix = gids % (nx + 1); iy = (gids // (nx + 1)) % (ny + 1); iz = gids // ps
    xyz = np.empty((n_loc, 3))
    xyz[:, 0] = -lx / 2 + ix.astype(np.float64) * (lx / 2 - (-lx / 2)) / nx
    xyz[:, 1] = -ly / 2 + iy.astype(np.float64) * (ly / 2 - (-ly / 2)) / ny
    xyz[:, 2] = -lz / 2 + iz.astype(np.float64) * (lz / 2 - (-lz / 2)) / nz
    # cells of the cube layers that touch an owned plane (every Kuhn tet spans both planes of its layer)
    k0, k1 = max(z0 - 1, 0), min(z1, nz)
    kk, jj, ii = np.meshgrid(np.arange(k0, k1), np.arange(ny), np.arange(nx), indexing="ij")
    v0 = (kk * ps + jj * (nx + 1) + ii).ravel().astype(np.int64)
    sx, sxy = nx + 1, ps
    v1, v2, v3, v4 = v0 + 1, v0 + sx, v0 + 1 + sx, v0 + sxy
    v5, v6, v7 = v4 + 1, v4 + sx, v4 + 1 + sx
    tets = np.stack([np.stack(t, axis=1) for t in ((v0, v1, v3, v7), (v0, v1, v5, v7), (v0, v4, v5, v7),
                                                   (v0, v2, v3, v7), (v0, v4, v6, v7), (v0, v2, v6, v7))], axis=1)
    gcells = tets.reshape(-1, 4)
    # global -> local
    def to_local(g):
        loc = np.full(g.shape, -1, dtype=np.int64)
        own = (g >= z0 * ps) & (g < z1 * ps)
        loc[own] = g[own] - z0 * ps
        if has_lo:
            m = (g >= (z0 - 1) * ps) & (g < z0 * ps)
            loc[m] = n_own + (g[m] - (z0 - 1) * ps)
        if has_hi:
            m = (g >= z1 * ps) & (g < (z1 + 1) * ps)
            loc[m] = n_own + n_lo + (g[m] - z1 * ps)
        return loc
    cells = to_local(gcells)
    assert cells.min() >= 0
    nbr, send_ptr, send_idx, recv_ptr = [], [0], [], [n_own]
    if has_lo:      # lower neighbour needs my first owned plane; its values fill my lower ghost plane
        nbr.append(rank - 1); send_idx.append(np.arange(0, ps)); send_ptr.append(send_ptr[-1] + ps); recv_ptr.append(recv_ptr[-1] + ps)
    if has_hi:
        nbr.append(rank + 1); send_idx.append(np.arange(n_own - ps, n_own)); send_ptr.append(send_ptr[-1] + ps); recv_ptr.append(recv_ptr[-1] + ps)
    return dict(n_own=n_own, n_loc=n_loc, gids=gids, xyz=np.ascontiguousarray(xyz.ravel()),
                cells=np.ascontiguousarray(cells.astype(np.int32).ravel()),
                nbr=np.array(nbr, dtype=np.int32), send_ptr=np.array(send_ptr, dtype=np.int32),
                send_idx=(np.concatenate(send_idx) if send_idx else np.zeros(0)).astype(np.int32),
                recv_ptr=np.array(recv_ptr, dtype=np.int32),
                bc_vertices=np.nonzero((ix == 0) | (ix == nx))[0].astype(np.int32),
                # proper 4-colouring of the Kuhn mesh graph from the GLOBAL lattice position (the same
                # classes on every rank; see box_coloring in csrc/mesh_setup.cu): not consumed yet, the
                # slabs are still coloured as general graphs on the device
                color=((ix + iy + iz) % 4).astype(np.int32))


# ---- the C++ partitioner of the library (csrc/partition.cpp): what bench.py and the tests use ----------
def brick_grid(nranks, nx, ny, nz):
    """px * py * pz = nranks with the most cube-like bricks for an nx x ny x nz box (minimal halo
    surface); ties resolved towards cutting z, then y (x is the contiguous index of the numbering)."""
    best, grid = None, (1, 1, nranks)
    for px in range(1, nranks + 1):
        if nranks % px:
            continue
        for py in range(1, nranks // px + 1):
            if (nranks // px) % py:
                continue
            pz = nranks // (px * py)
            if px > nx + 1 or py > ny + 1 or pz > nz + 1:
                continue
            bx, by, bz = (nx + 1) / px, (ny + 1) / py, (nz + 1) / pz
            surf = min(px - 1, 2) * by * bz + min(py - 1, 2) * bx * bz + min(pz - 1, 2) * bx * by     # an interior brick has two faces per cut direction
            key = (surf, px, py)
            if best is None or key < best:
                best, grid = key, (px, py, pz)
    return grid


def box_brick_partition(nx, ny, nz, lx, ly, lz, rank, nranks, grid=None):
    """pnpb200_partition_box: the rank's brick of the BoxMesh (same dict layout as box_slab_partition,
    plus the closed-form 4-colouring); grid = (px, py, pz), default brick_grid(). (1, 1, nranks) gives
    exactly the z-slabs of box_slab_partition."""
    from . import capi
    px, py, pz = grid or brick_grid(nranks, nx, ny, nz)
    assert px * py * pz == nranks
    d = capi.partition_box(nx, ny, nz, lx, ly, lz, px, py, pz, rank)
    d["grid"] = (px, py, pz)
    return d


def mesh_rcb_partition(xyz, cells, rank, nranks):
    """pnpb200_partition_mesh: recursive coordinate bisection of any conforming tetrahedral mesh"""
    from . import capi
    return capi.partition_mesh(xyz, cells, nranks, rank)
