"""-m "not gpu": the unstructured path on a REAL DOLFIN mesh — benchmarks/pnp_pb/mesh1.xml.gz, the one mesh the
reference ships as data (mshr: a box with a spherical hole, 6 542 vertices, 33 373 cells, vertex degrees 3 ... 40);
fixture tests/golden/mshr_sphere_mesh.npz (tests/golden/make_mesh_fixture.py). Checks what does not need a GPU: the
block pattern and the oracle's assembled Newton system keep their structural properties on it, and the product's C++
partitioner (recursive coordinate bisection, csrc/partition.cpp) cuts it into parts that cover it exactly once with
matching halo lists."""
import os
from importlib import import_module

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mesh():
    d = np.load(os.path.join(ROOT, "tests", "golden", "mshr_sphere_mesh.npz"))
    return d["xyz"], d["cells"]


def test_pattern_and_oracle_system_on_the_mshr_mesh(oracle, mesh):
    from helpers import GeneralProblem
    xyz, cells = mesh
    P = GeneralProblem(oracle, xyz, cells)
    assert (P.nv, P.nc) == (6542, 33373)
    # pattern = vertex adjacency + diagonal, symmetric, ascending columns (src/pde.cpp:542-555 through DOLFIN's builder)
    edges = set()
    for a in range(4):
        for b in range(a + 1, 4):
            edges.update(zip(np.minimum(cells[:, a], cells[:, b]).tolist(), np.maximum(cells[:, a], cells[:, b]).tolist()))
    assert P.nnzb == P.nv + 2 * len(edges)
    rows = np.repeat(np.arange(P.nv), np.diff(P.IA))
    pairs = set(zip(rows.tolist(), P.JA.tolist()))
    assert all((j, i) in pairs for (i, j) in list(pairs)[:20000]) and all((i, i) in pairs for i in range(P.nv))
    assert all(np.all(np.diff(P.JA[P.IA[i]:P.IA[i + 1]]) > 0) for i in range(0, P.nv, 7))
    assert np.diff(P.IA).max() > 16          # rows longer than the streamed kernels' 16 blocks: the general row kernels' case
    rng = np.random.default_rng(3)
    uu = P.exact + 0.05 * rng.uniform(-1, 1, P.N)
    bc_v = P.bcflag[0::3].astype(bool)
    assert 0 < bc_v.sum() < P.nv
    for use_eafe in (True, False):
        A, b = P.assemble(uu, use_eafe=use_eafe)
        V = A.reshape(-1, 3, 3)
        assert np.isfinite(A).all() and np.isfinite(b).all()
        # Dirichlet rows are exactly e_i with a zero rhs (src/pde.cpp:433-461, linear_pnp.cpp:78-80)
        diag = P.JA == rows
        for k in range(3):
            r = V[bc_v[rows], k, :]
            assert np.array_equal(r[diag[bc_v[rows]]][:, k], np.ones(bc_v.sum()))
            assert np.abs(r).sum() == bc_v.sum()
        assert np.all(b[P.bcflag.astype(bool)] == 0.0)
        # the Poisson block eps * (grad phi_j, grad phi_i): rows of a stiffness matrix sum to zero (constants are in its
        # kernel), on every vertex that is not a Dirichlet vertex; its diagonal is positive
        s = np.zeros(P.nv); np.add.at(s, rows, V[:, 0, 0])
        scale = np.zeros(P.nv); np.maximum.at(scale, rows, np.abs(V[:, 0, 0]))
        assert np.abs(s[~bc_v]).max() <= 1e-12 * scale.max()
        assert np.all(V[diag, 0, 0] > 0)
        # no cation-anion coupling in the linearised PNP system: entries (1,2) and (2,1) of every block are zero
        # (what the 7-value block format of the device solve relies on)
        assert np.all(V[:, 1, 2] == 0.0) and np.all(V[:, 2, 1] == 0.0)


@pytest.mark.parametrize("world", [2, 5])
def test_rcb_partition_of_the_mshr_mesh(pkg, mesh, world):
    part_mod = import_module("modular_pnp_b200.partition")
    xyz, cells = mesh
    nv = len(xyz)
    gcells = np.sort(cells, axis=1)
    parts = [part_mod.mesh_rcb_partition(xyz.ravel(), cells.ravel(), r, world) for r in range(world)]
    owner = -np.ones(nv, dtype=np.int64)
    for r, P in enumerate(parts):
        own = P["gids"][:P["n_own"]]
        assert np.all(np.diff(own) > 0) and np.all(owner[own] == -1)
        owner[own] = r
        assert np.array_equal(P["xyz"].reshape(-1, 3), xyz[P["gids"]])
    assert np.all(owner >= 0)
    sizes = np.bincount(owner, minlength=world)
    assert sizes.max() <= 1.1 * sizes.mean() + 8                    # bisection by vertex count: balanced
    touch = [np.zeros(len(gcells), dtype=bool) for _ in range(world)]
    oc = owner[gcells]
    for r in range(world):
        touch[r] = (oc == r).any(axis=1)
    for r, P in enumerate(parts):
        lc = P["gids"][P["cells"].reshape(-1, 4)]
        assert np.all(np.diff(lc, axis=1) > 0)                      # UFC order: ascending global ids
        assert sorted(map(tuple, lc.tolist())) == sorted(map(tuple, gcells[touch[r]].tolist()))
        assert P["recv_ptr"][0] == P["n_own"] and P["recv_ptr"][-1] == P["n_loc"] and np.all(np.diff(P["nbr"]) > 0)
        for k, nb in enumerate(P["nbr"]):
            gh = P["gids"][P["recv_ptr"][k]:P["recv_ptr"][k + 1]]
            assert np.all(owner[gh] == nb) and np.all(np.diff(gh) > 0)
            Q = parts[nb]
            kk = int(np.nonzero(Q["nbr"] == r)[0][0])
            sent = Q["gids"][Q["send_idx"][Q["send_ptr"][kk]:Q["send_ptr"][kk + 1]]]
            assert np.array_equal(sent, gh)
