// benchmarks/nsinker/nsinker_gmg.prm through the C++ host side (aspect_b200/cpp/stokes_matrix_free_b200.h): unit cube,
// `Initial global refinement = R`, n sinkers (nsinker.cc:64-76, 199-314), viscosity ratio 1e4,
// `Material averaging = harmonic average only viscosity`, zero velocity on all faces, gravity (0, 0, -1), tolerance 1e-6.
// Prints the lines of the reference's screen-output; `nsinker_box 2 4` is the mesh of tests/nsinker_gmg_gc (21 iterations).
//   g++ -std=c++17 -O2 -I include -I aspect_b200/cpp examples/nsinker_box.cc -L aspect_b200 -lsmf_b200 -Wl,-rpath,'$ORIGIN/../aspect_b200' -o examples/nsinker_box
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <exception>

#include "synthetic_box.h"

static const double centers[8][3] = {
  {2.4257829890e-01, 1.3469574514e-02, 3.8313885004e-01}, {4.1465269048e-01, 6.7768972864e-02, 9.9312692973e-01},
  {4.8430804651e-01, 7.6533776604e-01, 3.1833815403e-02}, {3.0935481671e-02, 9.3264044027e-01, 8.8787953411e-01},
  {5.9132973039e-01, 4.7877868473e-01, 8.3335433660e-01}, {1.8633519681e-01, 7.3565270739e-01, 1.1505317181e-01},
  {6.9865863058e-01, 3.5560411138e-01, 6.3830000658e-01}, {9.0821050755e-01, 2.9400041480e-01, 2.6497158886e-01}};

int main(int argc, char **argv)
{
  const int R = argc > 1 ? std::atoi(argv[1]) : 2, n_sinkers = std::min(8, argc > 2 ? std::atoi(argv[2]) : 4), dim = 3;
  try
    {
      using namespace aspect_b200;
      const HierarchyArrays hier = make_box_hierarchy(dim, R, "zero");
      const LevelArrays &act = hier.active;
      std::printf("Number of active cells: %lld (on %d levels)\n", (long long)act.n_cells, R + 1);
      std::printf("Number of degrees of freedom: %lld (%lld+%lld)\n", (long long)(act.n_dofs_u + act.n_dofs_p), (long long)act.n_dofs_u, (long long)act.n_dofs_p);

      // the material-model plugin stays on the host: viscosity and density at the quadrature points
      const std::vector<double> x = box_quadrature_points(dim, R);
      const std::size_t npts = (std::size_t)act.n_cells * 27;
      std::vector<double> eta(npts), force(npts * 3, 0.0);
      const double s = std::sqrt(1e4), delta = 200.0, omega = 0.1;
      for (std::size_t i = 0; i < npts; ++i)
        {
          double chi = 1.0;
          for (int k = 0; k < n_sinkers; ++k)
            {
              double d2 = 0;
              for (int d = 0; d < 3; ++d) d2 += (x[3 * i + d] - centers[k][d]) * (x[3 * i + d] - centers[k][d]);
              const double t = std::max(0.0, std::sqrt(d2) - omega / 2.0);
              chi *= 1.0 - std::exp(-delta * t * t);
            }
          eta[i] = (s - 1.0 / s) * (1.0 - chi) + 1.0 / s;
          force[3 * i + 2] = -10.0 * (1.0 - chi);           // density * gravity, gravity = (0, 0, -1)
        }

      StokesMatrixFreeHandlerB200 handler(dim);
      handler.parameters.linear_stokes_solver_tolerance = 1e-6;
      handler.parameters.n_cheap_stokes_solver_steps = 200;
      handler.parameters.n_expensive_stokes_solver_steps = 0;
      handler.setup_dofs(hier);
      handler.assemble(eta, SMF_AVG_HARMONIC, /*pressure scaling (nsinker.cc:320-331)*/ 1.0);
      handler.build_preconditioner();
      auto rhs_u = handler.initialize_dof_vector(0);
      handler.assemble_stokes_rhs(force, *rhs_u);
      std::vector<double> rhs((std::size_t)(act.n_dofs_u + act.n_dofs_p), 0.0), sol;
      rhs_u->to_host(rhs.data());
      const SolverOutputs out = handler.solve(rhs, sol);

      std::printf("   Solving Stokes system (%s)...\n", handler.name().c_str());
      std::printf("     GMG coarse size A: %lld, coarse size S: %lld\n", (long long)hier.levels[0].n_dofs_u, (long long)hier.levels[0].n_dofs_p);
      std::printf("     GMG levels: %d\n", R + 1);
      std::printf("     Viscosity range: %g - %g\n", handler.minimum_viscosity, handler.maximum_viscosity);
      std::printf("     Stokes solver (GMRES): %d+%d iterations.\n", out.iterations, out.iterations_expensive);
      std::printf("     Schur complement preconditioner: %d+0 iterations.\n", out.n_S_preconditioner);
      std::printf("     A block preconditioner: %d+0 iterations.\n", out.n_A_preconditioner);
      std::printf("     converged: %s, final residual %.6e\n", out.converged ? "yes" : "no", out.final_linear_residual);
      double umax = 0;
      for (std::int64_t i = 0; i < act.n_dofs_u; ++i) umax = std::max(umax, std::fabs(sol[(std::size_t)i]));
      std::printf("     max |velocity component|: %.12e\n", umax);
      return out.converged ? 0 : 2;
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
}
