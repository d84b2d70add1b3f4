"""-m "not gpu": the N>1 host logic on the CPU with world_size-2 gloo processes: slab partition,
ghost layout, halo lists and the "ghost rows are identity rows" local systems. Each rank assembles
its LOCAL matrix with the CPU oracle, refreshes ghosts through torch.distributed send/recv driven
by the same lists the CUDA path hands to NCCL, and the distributed SpMV / residual norms must
equal the single-process ones."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, ctypes as C
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from oracle import oracle as O
    from pnpb200_loader import load_package
    from importlib import import_module
    load_package()
    part_mod = import_module("modular_pnp_b200.partition"); prob = import_module("modular_pnp_b200.problems")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    nx, ny, nz, lx, ly, lz = 6, 3, 7, 2.0, 0.8, 1.6
    MODE = os.environ.get("PNPB200_TEST_PARTITION", "pyslabs")
    if MODE == "pyslabs":      # the Python reference partitioner (z-slabs)
        P = part_mod.box_slab_partition(nx, ny, nz, lx, ly, lz, rank, world)
    elif MODE == "rcb":        # csrc/partition.cpp: recursive coordinate bisection of the global arrays
        G0 = O.Problem(nx, ny, nz, lx, ly, lz)
        P = part_mod.mesh_rcb_partition(G0.xyz, G0.cells, rank, world)
        ixg = np.rint((P["xyz"][0::3] + lx / 2) * nx / lx).astype(np.int64)
        P["bc_vertices"] = np.nonzero((ixg == 0) | (ixg == nx))[0].astype(np.int32)
    else:                      # csrc/partition.cpp: bricks in closed form, cut along z / x / y
        grid = {"slabs": (1, 1, world), "xbricks": (world, 1, 1), "ybricks": (1, world, 1)}[MODE]
        P = part_mod.box_brick_partition(nx, ny, nz, lx, ly, lz, rank, world, grid=grid)
    n_own, n_loc = P["n_own"], P["n_loc"]
    L = O.lib()
    IA = np.zeros(n_loc + 1, dtype=np.int32)
    nnzb = L.orc_block_pattern_count(n_loc, len(P["cells"]) // 4, P["cells"], IA)
    JA = np.zeros(nnzb, dtype=np.int32); L.orc_block_pattern_fill(n_loc, len(P["cells"]) // 4, P["cells"], IA, JA)
    f = prob.exact_solution_fields(P["xyz"][0::3])
    bc = np.zeros(3 * n_loc, dtype=np.uint8)
    bc[(3 * P["bc_vertices"][:, None] + np.arange(3)).ravel()] = 1
    bc[3 * n_own:] = 1                                   # ghost rows = identity rows
    # global reference on every rank (small)
    G = O.Problem(nx, ny, nz, lx, ly, lz)
    rng = np.random.default_rng(7)
    uu_g = G.exact + 0.1 * rng.uniform(-1, 1, G.N)
    x_g = rng.uniform(-1, 1, G.N)
    A_g, b_g = G.assemble(uu_g)
    y_g = O.bsr_spmv(G.IA, G.JA, A_g, x_g)
    gd = (3 * P["gids"][:, None] + np.arange(3)).ravel()
    def halo(vec):
        """refresh ghost entries of a local dof vector with the send/recv lists of the partition"""
        reqs, bufs = [], []
        for k, nb in enumerate(P["nbr"]):
            si = P["send_idx"][P["send_ptr"][k]:P["send_ptr"][k + 1]]
            sb = torch.from_numpy(np.ascontiguousarray(vec.reshape(-1, 3)[si].ravel()))
            rb = torch.zeros(3 * (P["recv_ptr"][k + 1] - P["recv_ptr"][k]), dtype=torch.float64)
            reqs.append(dist.isend(sb, int(nb))); reqs.append(dist.irecv(rb, int(nb))); bufs.append((k, rb, sb))
        for r in reqs: r.wait()
        for k, rb, _ in bufs:
            vec[3 * P["recv_ptr"][k]:3 * P["recv_ptr"][k + 1]] = rb.numpy()
    # solution vector: owned part set, ghosts via halo exchange
    uu = np.zeros(3 * n_loc); uu[:3 * n_own] = uu_g[gd[:3 * n_own]]
    halo(uu)
    assert np.array_equal(uu, uu_g[gd]), "halo lists inconsistent"
    A = np.zeros(9 * nnzb); b = np.zeros(3 * n_loc)
    L.orc_assemble(n_loc, P["xyz"], len(P["cells"]) // 4, P["cells"], IA, JA, uu, f["eps"], f["fixed"], f["diff"], f["val"], f["reac"], bc, 1, A, b)
    x = np.zeros(3 * n_loc); x[:3 * n_own] = x_g[gd[:3 * n_own]]
    halo(x)
    y = O.bsr_spmv(IA, JA, A, x)
    own = gd[:3 * n_own]
    assert np.abs(y[:3 * n_own] - y_g[own]).max() <= 1e-12 * np.abs(y_g).max(), "distributed SpMV differs"
    assert np.abs(b[:3 * n_own] - b_g[own]).max() <= 1e-12 * np.abs(b_g).max(), "distributed rhs differs"
    assert np.all(b[3 * n_own:] == 0.0)
    t = torch.tensor([float((b[:3 * n_own] ** 2).sum())], dtype=torch.float64); dist.all_reduce(t)
    assert abs(t.item() ** 0.5 - np.linalg.norm(b_g)) <= 1e-12 * np.linalg.norm(b_g)
    m = torch.tensor([float(np.abs(b[:3 * n_own]).max())], dtype=torch.float64); dist.all_reduce(m, op=dist.ReduceOp.MAX)
    assert abs(m.item() - np.abs(b_g).max()) <= 1e-15
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok")
''') % ROOT


import pytest


@pytest.mark.parametrize("mode", ["pyslabs", "slabs", "xbricks", "ybricks", "rcb"])
def test_world2_partition_halo_and_distributed_spmv(tmp_path, mode):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", OMP_NUM_THREADS="1", PNPB200_TEST_PARTITION=mode)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


def test_partition_colouring_is_proper_and_the_same_on_every_rank(oracle, pkg):
    """the (i + j + k) mod 4 colouring the partition carries (global lattice position) is a proper
    colouring of every rank's local mesh graph, ghosts included, and a vertex two ranks share gets the
    same colour on both"""
    import numpy as np
    from importlib import import_module
    part_mod = import_module("modular_pnp_b200.partition")
    nx, ny, nz, world = 5, 4, 9, 3
    seen = {}
    for rank in range(world):
        P = part_mod.box_slab_partition(nx, ny, nz, 2.0, 1.0, 1.6, rank, world)
        col, cells = P["color"], P["cells"].reshape(-1, 4)
        assert col.shape == (P["n_loc"],) and col.min() >= 0 and col.max() <= 3
        # the four vertices of every cell carry the four colours (balanced triangulation)
        assert np.all(np.sort(col[cells], axis=1) == np.arange(4)[None, :])
        for g, c in zip(P["gids"], col):
            assert seen.setdefault(int(g), int(c)) == int(c)
    assert len(seen) == (nx + 1) * (ny + 1) * (nz + 1)
    # and it is the colouring the device uses for a BoxMesh: same formula on the global numbering
    G = oracle.Problem(nx, ny, nz, 2.0, 1.0, 1.6)
    gcol = np.array([seen[g] for g in range(G.nv)])
    gc = G.cells.reshape(-1, 4)
    assert np.all(np.sort(gcol[gc], axis=1) == np.arange(4)[None, :])


def test_cpp_partitioner_covers_the_mesh_exactly_once(oracle, pkg):
    """pnpb200_partition_box (bricks) and pnpb200_partition_mesh (RCB): every vertex is owned by exactly one
    rank; every cell appears on every rank that owns one of its vertices and nowhere else; the send list of
    rank a for rank b is the ghost list of b from a, in the same order (ascending global id)."""
    import numpy as np
    from importlib import import_module
    part_mod = import_module("modular_pnp_b200.partition")
    nx, ny, nz, lx, ly, lz = 7, 5, 6, 2.0, 1.0, 1.6
    G = oracle.Problem(nx, ny, nz, lx, ly, lz)
    gcells = np.sort(G.cells.reshape(-1, 4), axis=1)
    for world, maker in [(8, lambda r: part_mod.box_brick_partition(nx, ny, nz, lx, ly, lz, r, 8, grid=(2, 2, 2))),
                         (6, lambda r: part_mod.box_brick_partition(nx, ny, nz, lx, ly, lz, r, 6, grid=(3, 1, 2))),
                         (5, lambda r: part_mod.mesh_rcb_partition(G.xyz, G.cells, r, 5))]:
        parts = [maker(r) for r in range(world)]
        owner = -np.ones(G.nv, dtype=np.int64)
        for r, P in enumerate(parts):
            own = P["gids"][:P["n_own"]]
            assert np.all(np.diff(own) > 0) and np.all(owner[own] == -1)
            owner[own] = r
            assert np.allclose(P["xyz"].reshape(-1, 3), G.xyz.reshape(-1, 3)[P["gids"]], rtol=0, atol=1e-15)
        assert np.all(owner >= 0)
        sizes = np.bincount(owner, minlength=world)
        assert sizes.max() <= 1.35 * sizes.mean() + 8              # balanced parts
        cell_owner_sets = [set(owner[c]) for c in gcells]
        for r, P in enumerate(parts):
            lc = P["gids"][P["cells"].reshape(-1, 4)]
            assert np.all(np.diff(lc, axis=1) > 0)                  # UFC order: ascending global ids
            expect = np.array([c for c, os_ in zip(gcells, cell_owner_sets) if r in os_])
            assert sorted(map(tuple, lc)) == sorted(map(tuple, expect))
            # ghosts grouped by owner (ascending), ascending gid inside; recv ranges tile [n_own, n_loc)
            assert P["recv_ptr"][0] == P["n_own"] and P["recv_ptr"][-1] == P["n_loc"] and np.all(np.diff(P["nbr"]) > 0)
            for k, nb in enumerate(P["nbr"]):
                gh = P["gids"][P["recv_ptr"][k]:P["recv_ptr"][k + 1]]
                assert np.all(owner[gh] == nb) and np.all(np.diff(gh) > 0)
                Q = parts[nb]
                kk = int(np.nonzero(Q["nbr"] == r)[0][0])
                sent = Q["gids"][Q["send_idx"][Q["send_ptr"][kk]:Q["send_ptr"][kk + 1]]]
                assert np.array_equal(sent, gh)
