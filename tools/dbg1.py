import numpy as np, sys
sys.path.insert(0,'tests')
from oracle import mesh as omesh, solver as osolver
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic
from util import oracle_backend, rel
hier = synthetic.BoxHierarchy(3,2)
avg="harmonic average only viscosity"
eta_q, rhs, ps = benchmarks.nsinker_problem(hier, averaging=avg, n_sinkers=4)
h = StokesMatrixFreeHandler(3); h.setup_dofs(hier,"zero"); h.assemble(eta_q, avg, ps); h.build_preconditioner()
h.parameters["Linear solver tolerance"]=1e-6
sol,out = h.solve(rhs)
ob,*_ = oracle_backend(omesh.Hierarchy(3,2), eta_q, avg, ps, "zero")
s = osolver.BlockGMGStokes(ob); hist=[]
o = s.solve(rhs, tol_rel=1e-6, restart=100, max_its=1000, history=hist)
nu = hier.active.n_u
print("its", out["iterations"], o["iterations"], "res", out["final_residual"], o["residual"], "tol", out["tolerance"], o["tolerance"])
print("rel u", rel(sol[:nu], o["x"][:nu]), "rel p", rel(sol[nu:], o["x"][nu:]))
d = sol[nu:]-o["x"][nu:]
print("p diff mean", d.mean(), "std", d.std(), "|p|", np.linalg.norm(o["x"][nu:]))
# oracle vs oracle with perturbed rhs 1e-16
o2 = osolver.BlockGMGStokes(ob).solve(rhs*(1+1e-15), tol_rel=1e-6, restart=100, max_its=1000)
print("oracle self-sensitivity", rel(o2["x"], o["x"]))
for l in range(3):
    for w,mg in ((0,s.mg_A),(1,s.mg_S)):
        e = h.eigenvalue_estimates(l,w); sm = mg.smoothers[l]
        print(l,w,e, sm.min_eig_est, 1.2*sm.max_eig_est, sm.cg_iterations)
