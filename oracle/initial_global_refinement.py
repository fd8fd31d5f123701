"""TEST INFRASTRUCTURE (oracle): tests/initial_global_refinement.prm of the reference -- a second adaptive golden whose mesh follows
from refinement functions alone, with everything muparser_temperature_example does not have: ANISOTROPIC cells (a 1.2 x 1 box of
5 x 5 coarse cells: 0.24 x 0.2), FIVE cell sizes in one mesh, and DIFFERENT boundary conditions per face (tangential on the right,
zero velocity on the other three).

`Initial global refinement = 4` is four rounds of "flag every cell, let the plugins unflag, refine" (source/simulator/core.cc:2229-2247);
the `maximum refinement function` (4 for y >= 0.751, 2 for y >= 0.251, else 0, evaluated at the cell centre) clears the flag of cells
that have reached their level, p4est balances 2:1 over faces and corners.  Result, from the bottom: level 0 up to y = 0.2, level 1 to
0.3, level 2 to 0.7, level 3 (balance) to 0.75, level 4 above: 5 + 10 + 160 + 80 + 1,600 = 1,855 cells on 5 levels.
Golden (tests/initial_global_refinement/screen-output): 1,855 cells, 15,440 + 1,968 DoFs, 23+0 GMRES iterations of the default solver
(block GMG, `project to Q1 only viscosity`, tolerance 1e-7); velocity DoFs here: 15,290 = 15,440 - 2 x 75 coarse-fine edges (deal.II's
identity-constrained duplicates, see muparser_example.py).

Physics: `simple` material (viscosity 1, rho = 1 - 2e-5 (T - 1)), vertical gravity of magnitude 1 (the default), `perturbed box`
initial temperature 1 + sin(pi x / 1.2) sin(pi y) / 10 (source/initial_temperature/box.cc:30-48) with the boundary temperatures of
`Boundary temperature model = box` on the left (1) and right (0) faces, pressure scaling 1 / (0.01 x 1).
"""
import numpy as np

from . import adaptive as oa
from . import fe

EXTENTS = np.array([1.2, 1.0])
REPS = (5, 5)
COARSE = EXTENTS / np.array(REPS)
BC = {(0, 0): "zero", (0, 1): "tangential", (1, 0): "zero", (1, 1): "zero"}      # boundary indicators 0, 1, 2, 3 = left, right, bottom, top


def maximum_refinement_level(y):
    return 4 if y >= 0.751 else (2 if y >= 0.251 else 0)


def mesh():
    leaves = {(0, (i, j)) for i in range(REPS[0]) for j in range(REPS[1])}
    for _ in range(4):
        marked = {c for c in leaves if c[0] < maximum_refinement_level((c[1][1] + 0.5) * COARSE[1] / (1 << c[0]))}
        leaves = oa.AdaptiveBox._refine(2, leaves, marked, REPS)
    return oa.AdaptiveBox(2, leaves, REPS, COARSE)


def problem(box):
    """(viscosity per cell, velocity right-hand side before condensation, pressure scaling)."""
    t = fe.tensor_tables(2)
    unit = COARSE / (2 << box.Lmax)
    keys = np.array(sorted(box.node_u, key=box.node_u.get))
    x, y = keys[:, 0] * unit[0], keys[:, 1] * unit[1]
    T = 1.0 + np.sin(np.pi * x / EXTENTS[0]) * np.sin(np.pi * y / EXTENTS[1]) / 10.0
    T[keys[:, 0] == 0] = 1.0                                              # `Left temperature` (default 1)
    T[keys[:, 0] == REPS[0] * (2 << box.Lmax)] = 0.0                      # `Right temperature` (default 0)
    for k, w in box.hanging_u.items():
        T[k] = sum(v * T[m] for m, v in w)
    T_q = T[box.cell_dofs_u[:, :9] // 2] @ t["phi2"].T
    rho = 1.0 - 2e-5 * (T_q - 1.0)
    f = np.zeros(box.n_u)
    for m, cell in enumerate(box.cells):
        np.add.at(f, box.cell_dofs_u[m, 9:18], (-rho[m] * t["w"] * float(np.prod(box.cell_size(cell)))) @ t["phi2"])
    return np.ones(box.n_cells), f, 1.0 / 0.01


def solve():
    box = mesh()
    eta, f, ps = problem(box)
    ops = box.assemble(eta, ps, BC)
    rhs = np.concatenate([ops["Cu"].T @ f, np.zeros(box.n_p)])
    return box, oa.block_gmg_solve(box, eta, rhs, BC, ps, 1e-7, 100, 1000)
