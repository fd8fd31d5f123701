"""Generate the operator-level golden vectors under tests/golden/ (run here, commit the output).

Source of truth: the oracle's ASSEMBLED sparse Q2-Q1 operators (oracle/assembled.py: textbook weak form, no
matrix-free code), which reproduce the reference's screen-output goldens (tests/test_oracle_goldens.py).  Each
fixture holds seeded inputs and the exact outputs of StokesOperator / ABlockOperator / BTBlockOperator /
MassMatrixOperator ::vmult, the inverse diagonals, one prolongation and one restriction, for a small mesh.
    python tools/make_operator_goldens.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import coeff, mesh, solver  # noqa: E402

CASES = [("box2d_r2_tangential_harmonic", 2, 2, "tangential", "harmonic", 1e5),
         ("box2d_r3_zero_none", 2, 3, "zero", "none", 3.7),
         ("box3d_r1_zero_none", 3, 1, "zero", "none", 1.0),
         ("box3d_r2_zero_harmonic", 3, 2, "zero", "harmonic", 2.5),
         ("box3d_r2_tangential_none", 3, 2, "tangential", "none", 1.0)]


# curved cells: the spherical shell of tests/hollow_sphere_gmg (12 coarse cells, refined once), geometry = the restated
# SphericalManifold / MappingQ(4) rules (oracle/shell.py); the fixtures also lock that geometry (inverse Jacobians, JxW)
SHELL_CASES = [("shell12_r1_zero_none", 12, 1, 0.5, 1.0, "none", 3.7),
               ("shell12_r1_zero_harmonic", 12, 1, 0.5, 1.0, "harmonic", 571.0)]


def shell_fixtures():
    from oracle import shell as oshell
    for name, n_coarse, R, r0, r1, avg, ps in SHELL_CASES:
        rng = np.random.default_rng(sum(map(ord, name)))
        hier = oshell.ShellHierarchy(R, r0, r1, n_coarse=n_coarse)
        act = hier.active
        eta_q = np.exp(rng.uniform(-4, 4, size=(act.n_cells, 27)))
        active, levels, emin, emax = coeff.viscosity_tables(hier, eta_q, avg)
        be = solver.AssembledBackend(hier, active, levels, ps, "zero")
        xu, xp = rng.uniform(-1, 1, act.n_u), rng.uniform(-1, 1, act.n_p)
        xu[act.constrained_u("zero")] = 0.0
        yu, yp = be.stokes(xu, xp)
        out = {"n_coarse": n_coarse, "R": R, "inner_radius": r0, "outer_radius": r1, "averaging": avg, "pressure_scaling": ps,
               "eta_q": eta_q, "xu": xu, "xp": xp, "stokes_u": yu, "stokes_p": yp, "A_active": be.A_active(xu), "BT": be.BT(xp),
               "mass_active": be.M_active(xp), "active_table": active, "eta_min_max": np.array([emin, emax]),
               "jinv_active": act.jinv, "jxw_active": act.jxw, "cell_dofs_u": act.cell_dofs_u.astype(np.uint32)}
        for l in range(R + 1):
            L = hier.levels[l]
            zu, zp = rng.uniform(-1, 1, L.n_u), rng.uniform(-1, 1, L.n_p)
            zu[L.constrained_u("zero")] = 0.0
            out[f"level{l}_xu"], out[f"level{l}_xp"] = zu, zp
            out[f"level{l}_A"], out[f"level{l}_mass"] = be.apply_A(l, zu), be.apply_M(l, zp)
            out[f"level{l}_dinv_A"], out[f"level{l}_dinv_M"] = be.dinv_A(l), be.dinv_M(l)
            if l > 0:
                C = hier.levels[l - 1]
                cu, cp = rng.uniform(-1, 1, C.n_u), rng.uniform(-1, 1, C.n_p)
                cu[C.constrained_u("zero")] = 0.0
                out[f"level{l}_coarse_u"], out[f"level{l}_coarse_p"] = cu, cp
                out[f"level{l}_prolongate_u"], out[f"level{l}_prolongate_p"] = be.prolongate_u(l, cu), be.prolongate_p(l, cp)
                out[f"level{l}_restrict_u"], out[f"level{l}_restrict_p"] = be.restrict_u(l, zu), be.restrict_p(l, zp)
        path = os.path.join(ROOT, "tests", "golden", f"curved_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB")


def main():
    if "--shell-only" not in sys.argv:
        box_fixtures()
    shell_fixtures()


def box_fixtures():
    for name, dim, R, bc, avg, ps in CASES:
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        hier = mesh.Hierarchy(dim, R)
        act = hier.active
        eta_q = np.exp(rng.uniform(-4, 4, size=(act.n_cells, 3 ** dim)))
        active, levels, emin, emax = coeff.viscosity_tables(hier, eta_q, avg)
        be = solver.AssembledBackend(hier, active, levels, ps, bc)
        xu, xp = rng.uniform(-1, 1, act.n_u), rng.uniform(-1, 1, act.n_p)
        xu[act.constrained_u(bc)] = 0.0
        yu, yp = be.stokes(xu, xp)
        out = {"dim": dim, "R": R, "bc": bc, "averaging": avg, "pressure_scaling": ps, "eta_q": eta_q,
               "xu": xu, "xp": xp, "stokes_u": yu, "stokes_p": yp, "A_active": be.A_active(xu), "BT": be.BT(xp),
               "mass_active": be.M_active(xp)}
        for l in range(R + 1):
            L = hier.levels[l]
            zu, zp = rng.uniform(-1, 1, L.n_u), rng.uniform(-1, 1, L.n_p)
            zu[L.constrained_u(bc)] = 0.0
            out[f"level{l}_xu"], out[f"level{l}_xp"] = zu, zp
            out[f"level{l}_A"], out[f"level{l}_mass"] = be.apply_A(l, zu), be.apply_M(l, zp)
            out[f"level{l}_dinv_A"], out[f"level{l}_dinv_M"] = be.dinv_A(l), be.dinv_M(l)
            if l > 0:
                C = hier.levels[l - 1]
                cu, cp = rng.uniform(-1, 1, C.n_u), rng.uniform(-1, 1, C.n_p)
                cu[C.constrained_u(bc)] = 0.0
                out[f"level{l}_coarse_u"], out[f"level{l}_coarse_p"] = cu, cp
                out[f"level{l}_prolongate_u"], out[f"level{l}_prolongate_p"] = be.prolongate_u(l, cu), be.prolongate_p(l, cp)
                out[f"level{l}_restrict_u"], out[f"level{l}_restrict_p"] = be.restrict_u(l, zu), be.restrict_p(l, zp)
        path = os.path.join(ROOT, "tests", "golden", f"operators_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
