"""Host logic of the lattice path (aspect_b200/csrc/schur_lattice.cu: detect_lattice) through the host-only ABI utility
smf_detect_box_lattice: no GPU needed.  The library must recover the box from the cell -> vertex maps alone -- whatever the DoF
numbering and the cell order -- and refuse anything that is not a box."""
import ctypes

import numpy as np

from aspect_b200 import _lib, synthetic


def _detect(cell_dofs_p, n_p):
    lib = _lib.load()
    cd = np.ascontiguousarray(cell_dofs_p, dtype=np.uint32)
    dims = np.zeros(3, dtype=np.int32)
    xyz = np.zeros((len(cd), 3), dtype=np.int32)
    lib.smf_detect_box_lattice.restype = ctypes.c_int
    rc = lib.smf_detect_box_lattice(_lib.ptr(cd), ctypes.c_int64(len(cd)), ctypes.c_int64(n_p), _lib.ptr(dims), _lib.ptr(xyz))
    return rc, dims, xyz


def test_box_levels_are_detected_with_their_cell_coordinates():
    for reps, level in (((1, 1, 1), 3), ((3, 2, 1), 2), ((2, 2, 2), 1)):
        L = synthetic.BoxLevel(3, level, None, reps)
        rc, dims, xyz = _detect(L.cell_dofs_p, L.n_p)
        assert rc == 1
        assert tuple(dims) == tuple(np.asarray(reps) * (1 << level) + 1)
        assert np.array_equal(xyz, L.cells)                      # the driver's own integer cell coordinates


def test_detection_does_not_depend_on_numbering_or_cell_order():
    L = synthetic.BoxLevel(3, 2, None, (2, 1, 1))
    rng = np.random.default_rng(5)
    perm_dofs = rng.permutation(L.n_p).astype(np.uint32)          # renumber the DoFs ...
    perm_cells = rng.permutation(L.n_cells)                       # ... and shuffle the cells
    cd = perm_dofs[L.cell_dofs_p.astype(np.int64)][perm_cells]
    rc, dims, xyz = _detect(cd, L.n_p)
    assert rc == 1 and tuple(dims) == (9, 5, 5)
    assert np.array_equal(xyz, L.cells[perm_cells])


def test_non_boxes_are_refused():
    L = synthetic.BoxLevel(3, 2)
    keep = ~((L.cells[:, 0] >= 2) & (L.cells[:, 1] >= 2))         # an L-shaped prism: one quadrant of columns removed
    cd = L.cell_dofs_p[keep]
    used = np.unique(cd)
    remap = np.full(L.n_p, -1, dtype=np.int64); remap[used] = np.arange(len(used))
    rc, _, _ = _detect(remap[cd.astype(np.int64)].astype(np.uint32), len(used))
    assert rc == 0
    # a box whose cells carry a rotated local vertex order is not the lattice the kernels assume either
    cd2 = L.cell_dofs_p.copy()
    cd2[0] = cd2[0][[1, 3, 0, 2, 5, 7, 4, 6]]
    rc, _, _ = _detect(cd2, L.n_p)
    assert rc == 0
    # shell levels (curved, 96 trees of different orientation) must fall through as well
    from aspect_b200 import shell
    S = shell.ShellLevel(1)
    rc, _, _ = _detect(S.cell_dofs_p, S.n_p)
    assert rc == 0
