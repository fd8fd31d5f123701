import numpy as np, sys
sys.path.insert(0,'tests')
from oracle import mesh as omesh, solver as osolver
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic
from util import oracle_backend, rel, seeded
R=2
hier = synthetic.BoxHierarchy(2,R)
avg="harmonic average only viscosity"
eta_q, rhs, ps = benchmarks.solcx_problem(hier, averaging=avg)
h = StokesMatrixFreeHandler(2); h.setup_dofs(hier,"tangential"); h.assemble(eta_q, avg, ps); h.build_preconditioner()
h.parameters["Linear solver tolerance"]=1e-7
h.setup_smoothers()
ob,*_ = oracle_backend(omesh.Hierarchy(2,R), eta_q, avg, ps, "tangential")
s = osolver.BlockGMGStokes(ob)
for l in range(R+1):
    for w,mg in ((0,s.mg_A),(1,s.mg_S)):
        e = h.eigenvalue_estimates(l,w); sm = mg.smoothers[l]
        print(l,w,e, sm.min_eig_est, 1.2*sm.max_eig_est, sm.cg_iterations)
act=hier.active
xu, xp = seeded(act.n_u, 5), seeded(act.n_p, 6)
xu[act.constrained_u("tangential").astype(np.int64)] = 0.0
src = h.initialize_dof_vector(); dst = h.initialize_dof_vector()
src[0].from_host(xu); src[1].from_host(xp)
h.prec_A.vmult(dst[0], src[0]); print("vcycle A", rel(dst[0].to_host(), s.mg_A.vmult(xu)))
h.prec_Schur.vmult(dst[1], src[1]); print("vcycle S", rel(dst[1].to_host(), s.mg_S.vmult(xp)))
sol,out = h.solve(rhs); o = s.solve(rhs, tol_rel=1e-7)
print(out, o["iterations"], o["tolerance"])
