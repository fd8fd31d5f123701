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
