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
    P = part_mod.box_slab_partition(nx, ny, nz, lx, ly, lz, rank, world)
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


def test_world2_partition_halo_and_distributed_spmv(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", OMP_NUM_THREADS="1")
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
