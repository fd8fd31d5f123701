"""Sensitivity of the spherical_nsinker_gmg_gc iteration count (reference: 106, restatement: 110) to the level transfer of the
viscosity: the coarse level table as arithmetic / volume-weighted / harmonic / geometric mean of the children or the first child.
CPU only (oracle/); ~45 s per variant.  Result of round 2: 110 / 108 / 193 / 120 / 116."""
import sys
import numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import assembled, coeff, shell as oshell, solver
o = oshell.ShellHierarchy(1)
L = o.active
eo, ro = coeff.nsinker_shell(L.xq)
active, levels, _, _ = coeff.viscosity_tables(o, eo, "harmonic")
xm, det = L.midpoint_geometry()
ps = coeff.pressure_scaling_shell(L, coeff.nsinker_shell(xm)[0], det)
g = -L.xq / np.linalg.norm(L.xq, axis=-1, keepdims=True)
f = assembled.rhs_force(L, coeff.average(ro, "harmonic")[..., None] * g)
f[L.constrained_u("zero")] = 0.0
rhs = np.concatenate([f, np.zeros(L.n_p)])
ch = o.children(1)
vol = L.jxw.sum(axis=1)
variants = {
  "arithmetic (baseline)": levels[0],
  "volume weighted": ((levels[1][ch, 0] * vol[ch]).sum(1) / vol[ch].sum(1))[:, None],
  "harmonic of children": (8.0 / (1.0 / levels[1][ch, 0]).sum(1))[:, None],
  "geometric of children": np.exp(np.log(levels[1][ch, 0]).mean(1))[:, None],
  "first child": levels[1][ch[:, 0], 0][:, None],
}
for name, l0 in variants.items():
    lv = [l0, levels[1]]
    be = solver.AssembledBackend(o, active, lv, ps, "zero")
    out = solver.BlockGMGStokes(be).solve(rhs, tol_rel=1e-6, restart=200, max_its=2000)
    print(name, out["iterations"], flush=True)
