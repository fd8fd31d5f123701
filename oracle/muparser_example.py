"""TEST INFRASTRUCTURE (oracle): tests/muparser_temperature_example.prm of the reference, the golden that pins the ADAPTIVE
local-smoothing path (hanging nodes, refinement edges, interface operators, copy indices) without any error indicator.

The test runs the default Stokes solver -- block GMG with `project to Q1 only viscosity` (helper_functions.cc:2813-2862) -- on a
5 x 1 box of 1,000 km coarse cells, three times: on the uniform mesh (global refinement 3) and after each of two initial adaptive
refinement steps whose ONLY strategy is `minimum refinement function` = 5 where depth <= 150 km, else 0 (minimum level 1).  With
no error indicator every cell is flagged for coarsening and the function alone decides (mesh_refinement/
minimum_refinement_function.cc: cells below their minimum level are refined, cells at it keep their level), p4est's 2:1 balance
does the rest; the resulting meshes are functions of depth only:
    mesh 1: level 3 everywhere                                                        (320 cells)
    mesh 2: level 4 above 875 km, 3 above 750 km, 2 below                             (260 cells on 5 levels)
    mesh 3: level 5 above 875 km, 4 above 750 km, 3 above 500 km, 2 below             (920 cells on 6 levels)
Golden (tests/muparser_temperature_example/screen-output): 2,754+369 / 2,546+347 / 8,322+1,091 velocity + pressure DoFs and
6 / 9 / 9 GMRES iterations.  deal.II counts the vertex of the fine cells that sits on the midpoint of a coarse edge as a DoF of
its own, constrained to the coarse edge's DoF with weight one; this restatement (and the product's driver) identifies the two,
so its velocity count is lower by dim x (number of coarse-fine edges) -- 120 and 280 here -- and nothing else differs.

Start vector: the tests here solve from a ZERO guess with the tolerance 1e-6 |b| (7 / 9 / 10 iterations).  The reference starts
from its initial pressure field -- the adiabatic (lithostatic) pressure, initial_conditions.cc:499-554 -- and measures the tolerance
against |b_u - B^T p_0| (local_smoothing.cc:1001-1025); with the adiabatic profile of this prm (surface temperature 273 K against a
1673 K mantle: p_0 balances the weight to 2.8 %) the restatement then needs 8 / 10 / 11.  In both variants the first GMRES step
makes no progress (the residual is a gradient field, which the block-triangular preconditioner turns into a pressure residual of the
same size: 1, 1, 0.16, 1e-2, ...), and the reference's counts are those of the same history WITHOUT that step; why its first step is
productive is not known.  Within +-1 from a zero guess, within 2 with the reference's start.

Physics: `simple` material (viscosity 1e21, rho = 3300 (1 - 2e-5 (T - 273))), vertical gravity 9.81, tangential velocity on all
sides, the half-space-cooling initial temperature of the prm interpolated to the Q2 nodes with the boundary temperatures (top 273,
bottom 1673) and the hanging-node constraints distributed, pressure scaling = viscosity / (0.01 x depth) (geometry_model/box.cc:252-255).
"""
import numpy as np
from scipy.special import erfc

from . import adaptive as oa
from . import fe

H = 1e6                                  # edge of a coarse cell: X extent 5e6 / 5 repetitions, Y extent 1e6 / 1
REPS = (5, 1)
REQUIRED = {                             # level as a function of the cell's upper edge y1 (the cells are aligned with these depths)
    1: lambda y1: 3,
    2: lambda y1: 4 if y1 > 0.875e6 else (3 if y1 > 0.75e6 else 2),
    3: lambda y1: 5 if y1 > 0.875e6 else (4 if y1 > 0.75e6 else (3 if y1 > 0.5e6 else 2)),
}


def mesh(which):
    """The active mesh of solve `which` (1, 2, 3) as an AdaptiveBox."""
    leaves, todo = set(), [(0, (i, 0)) for i in range(REPS[0])]
    while todo:
        l, ijk = todo.pop()
        if REQUIRED[which]((ijk[1] + 1) * H / (1 << l)) > l:
            todo += oa._children((l, ijk), 2)
        else:
            leaves.add((l, ijk))
    return oa.AdaptiveBox(2, leaves, REPS, H)


def initial_temperature(x, y):
    """`Initial temperature model/Function` of the prm (half-space cooling, plate age x / vsub left of the trench)."""
    ymax, xtrm, vsub, ageop, Tm, Ts, kappa = 1.0e6, 2.2e6, 7.927e-10, 9.46e14, 1673.0, 273.0, 1e-6
    plate = (x > 0.0) & (x <= xtrm)
    age = np.where(plate, x / vsub, ageop)
    cooled = Ts + (Tm - Ts) * (1.0 - erfc((ymax - y) / (2.0 * np.sqrt(kappa * age))))
    return np.where(x > 0.0, cooled, Tm)


def problem(box):
    """(viscosity per cell, velocity right-hand side before condensation, pressure scaling)."""
    t = fe.tensor_tables(2)
    unit = H / (2 << box.Lmax)                                 # lattice unit of the Q2 node keys
    keys = np.array(sorted(box.node_u, key=box.node_u.get))
    T = initial_temperature(keys[:, 0] * unit, keys[:, 1] * unit)
    T[keys[:, 1] == REPS[1] * (2 << box.Lmax)] = 273.0         # `Boundary temperature model = box`: top, bottom
    T[keys[:, 1] == 0] = 1673.0
    for k, w in box.hanging_u.items():                         # constraints.distribute (initial_conditions.cc:232-233)
        T[k] = sum(v * T[m] for m, v in w)
    T_q = T[box.cell_dofs_u[:, :9] // 2] @ t["phi2"].T
    rho = 3300.0 * (1.0 - 2e-5 * (T_q - 273.0))
    f = np.zeros(box.n_u)
    for m, cell in enumerate(box.cells):
        vol = float(np.prod(box.cell_size(cell)))
        np.add.at(f, box.cell_dofs_u[m, 9:18], (rho[m] * (-9.81) * t["w"] * vol) @ t["phi2"])
    return np.full(box.n_cells, 1e21), f, 1e21 / (0.01 * 1e6)


def solve(which):
    box = mesh(which)
    eta, f, ps = problem(box)
    ops = box.assemble(eta, ps, "tangential")
    rhs = np.concatenate([ops["Cu"].T @ f, np.zeros(box.n_p)])
    return box, oa.block_gmg_solve(box, eta, rhs, "tangential", ps, 1e-6, 100, 200)
