"""Distributed layout: host-side logic on CPU with a world_size-2 gloo group; the NCCL path on >= 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from aspect_b200 import synthetic
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for level in (1, 2, 3):
    L = synthetic.BoxLevel(3, level, None, (2, 1, 1), rank, world)
    for fam in "up":
        lists = L.exchange_lists(fam)
        gid = L.global_node_ids(fam)
        mine = {q: gid[v] for q, v in lists.items()}
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        for q, g in mine.items():                      # both sides list the same shared nodes in the same order
            ok &= np.array_equal(g, everyone[q][rank])
        idx, ranks = L.sharers(fam)
        w = np.ones(len(gid)); w[idx] = 1.0 / (ranks >= 0).sum(axis=1)
        tot = [None] * world
        dist.all_gather_object(tot, (gid, w))
        allg = np.concatenate([t[0] for t in tot]); allw = np.concatenate([t[1] for t in tot])
        u, inv = np.unique(allg, return_inverse=True)
        ok &= np.allclose(np.bincount(inv, weights=allw), 1.0)      # the inner-product weights of every DoF sum to one
        n_global = np.prod(L.grid_u if fam == "u" else L.grid_p)
        ok &= len(u) == n_global
# the 96-tree spherical shell (curved cells): 48 trees per rank, interfaces between trees of different orientation
from aspect_b200 import shell
for level in (0, 1):
    L = shell.ShellLevel(level, rank=rank, world=world)
    for fam in "up":
        lists, gid, w = L.exchange_lists(fam), L.global_node_ids(fam), 1.0 / L.sharer_counts(fam)
        everyone = [None] * world
        dist.all_gather_object(everyone, {q: gid[v] for q, v in lists.items()})
        for q, v in lists.items():
            ok &= np.array_equal(gid[v], everyone[q][rank])
        tot = [None] * world
        dist.all_gather_object(tot, (gid, w))
        allg = np.concatenate([t[0] for t in tot]); allw = np.concatenate([t[1] for t in tot])
        u, inv = np.unique(allg, return_inverse=True)
        ok &= np.allclose(np.bincount(inv, weights=allw), 1.0) and len(u) == L.n_global_nodes[fam]
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_partition_layout_world2_gloo(tmp_path):
    worker = tmp_path / "worker.py"
    worker.write_text(_GLOO_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(worker), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_single_rank_multi_tree_mesh_is_consistent():
    """All trees on one rank: no interfaces, DoF counts of the global lattice, same cell->DoF structure per tree."""
    from aspect_b200 import synthetic
    L = synthetic.BoxLevel(3, 2, None, (2, 2, 1), 0, 1)
    assert L.n_cells == 4 * 64 and L.n_u == 3 * 17 * 17 * 9 and L.n_p == 9 * 9 * 5
    assert L.exchange_lists("u") == {} and L.exchange_lists("p") == {}
    one = synthetic.BoxLevel(3, 2)
    assert np.array_equal(L.cell_dofs_u[:64], one.cell_dofs_u)       # the first tree is numbered like a single-tree box


@pytest.mark.gpu
def test_two_gpu_solve_matches_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(ROOT, "tools", "dist_check.py"), "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "-> OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_two_gpu_shell_solve_matches_one_gpu():
    """Curved cells over two ranks: 48 trees of the spherical shell each, halo exchange over the tree interfaces."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29521", os.path.join(ROOT, "tools", "dist_check.py"), "2", "shell"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "-> OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def _gids(L, dim=3):
    gu = L.global_node_ids("u")
    return (gu[:, None] * dim + np.arange(dim)[None, :]).ravel(), L.global_node_ids("p")


@pytest.mark.gpu
@pytest.mark.parametrize("rep,world", [((2, 1, 1), 2), ((2, 2, 1), 4)])
def test_subdomains_on_one_gpu_sum_to_the_global_operators(rep, world):
    """The partition on ONE device (runs on a 1-GPU box): the subdomains of all `world` ranks live in `world` single-rank
    contexts, the host plays the halo exchange (sum of the per-rank partial sums over the partition-independent DoF numbers,
    = what k_exchange_p2p / NCCL do between the ranks), and the result must be the operator of the global mesh held by one more
    context: active Stokes operator, A block and pressure mass matrix of every level, inverse diagonals.  What this does not
    cover is the transport itself (tests above, >= 2 GPUs; tools/dist_check.py inside bench.py at N > 1)."""
    from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic
    from util import rel
    R, avg = 2, "harmonic average"

    def context(rank, nranks):
        hier = synthetic.BoxHierarchy(3, R, None, rep, rank, nranks)
        eta_q, _, ps = benchmarks.nsinker_problem(hier, 8, 1e4, avg)
        hier.rank, hier.world = 0, 1                    # the handler sees one rank: no communicator, operators leave partial sums
        h = StokesMatrixFreeHandler(3, device=0)
        h.setup_dofs(hier, "zero")
        h.assemble(eta_q, avg, ps)
        h.build_preconditioner()
        return hier, h

    hg, Hg = context(0, 1)
    subs = [context(r, world) for r in range(world)]
    assert sum(s[0].active.n_cells for s in subs) == hg.active.n_cells

    def check(level_of, apply, family, salt, diag=None):
        Lg = level_of(hg)
        gid_g = _gids(Lg)[family]
        n_glob = int(gid_g.max()) + 1
        x_glob = np.sin(0.37 * np.arange(n_glob) + salt)
        cons_g = np.zeros(n_glob, bool)
        if family == 0:
            cons_g[gid_g[Lg.constrained_u("zero").astype(np.int64)]] = True
            x_glob[cons_g] = 0.0
        y_ref = np.zeros(n_glob); y_ref[gid_g] = apply(Hg, x_glob[gid_g], Lg)
        y_sum = np.zeros(n_glob)
        for hs, Hs in subs:
            Ls = level_of(hs)
            gid = _gids(Ls)[family]
            if family == 0:                             # each rank constrains exactly the global boundary DoFs it holds
                mine = np.zeros(n_glob, bool); mine[gid[Ls.constrained_u("zero").astype(np.int64)]] = True
                assert np.array_equal(mine[gid], cons_g[gid])
            y = apply(Hs, x_glob[gid], Ls)
            assert len(np.unique(gid)) == len(gid)
            y_sum[gid] += y
        free = ~cons_g
        assert rel(y_sum[free], y_ref[free]) < 1e-12, (family, salt)

    def vec(H, x, block, l):
        return H.initialize_dof_vector(block, l).from_host(x), H.initialize_dof_vector(block, l)

    def stokes(H, xu, xp):
        src, dst = H.initialize_dof_vector(), H.initialize_dof_vector()
        src[0].from_host(xu); src[1].from_host(xp)
        H.stokes_matrix.vmult(dst, src)
        return dst[0].to_host(), dst[1].to_host()

    # active Stokes operator: both blocks of the result
    from aspect_b200 import SMF_ACTIVE
    Lg = hg.active
    gu_g, gp_g = _gids(Lg)
    nu_glob, np_glob = int(gu_g.max()) + 1, int(gp_g.max()) + 1
    xu_glob, xp_glob = np.sin(0.37 * np.arange(nu_glob) + 1.0), np.sin(0.37 * np.arange(np_glob) + 2.0)
    cons = np.zeros(nu_glob, bool); cons[gu_g[Lg.constrained_u("zero").astype(np.int64)]] = True
    xu_glob[cons] = 0.0
    yu_ref, yp_ref = np.zeros(nu_glob), np.zeros(np_glob)
    a, b = stokes(Hg, xu_glob[gu_g], xp_glob[gp_g]); yu_ref[gu_g] = a; yp_ref[gp_g] = b
    yu_sum, yp_sum = np.zeros(nu_glob), np.zeros(np_glob)
    for hs, Hs in subs:
        gu, gp = _gids(hs.active)
        a, b = stokes(Hs, xu_glob[gu], xp_glob[gp])
        yu_sum[gu] += a; yp_sum[gp] += b
    assert rel(yu_sum[~cons], yu_ref[~cons]) < 1e-12 and rel(yp_sum, yp_ref) < 1e-12

    # level operators and diagonals
    for l in range(R + 1):
        def A(H, x, L, l=l):
            s, d = vec(H, x, 0, l); H.mg_matrices_A_block[l].vmult(d, s); return d.to_host()

        def M(H, x, L, l=l):
            s, d = vec(H, x, 1, l); H.mg_matrices_Schur_complement[l].vmult(d, s); return d.to_host()
        check(lambda h, l=l: h.levels[l], A, 0, 10.0 + l)
        check(lambda h, l=l: h.levels[l], M, 1, 20.0 + l)
        # diagonals: the per-rank diagonals are partial sums as well (the distributed set-up exchanges them before inverting)
        check(lambda h, l=l: h.levels[l], lambda H, x, L, l=l: 1.0 / H.get_matrix_diagonal_inverse(l, 0), 0, 0.0)
        check(lambda h, l=l: h.levels[l], lambda H, x, L, l=l: 1.0 / H.get_matrix_diagonal_inverse(l, 1), 1, 0.0)
    for _, H in subs + [(hg, Hg)]:
        H.close()
