// Quick-turnaround operator benchmark through the C++ host side (no Python, host set-up in seconds): nsinker box at global
// refinement R, timings of the Stokes / A / B^T / mass operators, the two V-cycles and one solve, wall clock around
// smf_synchronize after warm-up.  For kernel work on the GPU box:   ./examples/bench_operators 7 20
//   g++ -std=c++17 -O2 -I include -I aspect_b200/cpp examples/bench_operators.cc -L aspect_b200 -lsmf_b200 -Wl,-rpath,'$ORIGIN/../aspect_b200' -o examples/bench_operators
// (numbers for the record come from bench.py, which times with CUDA events inside the solves; this is the lab bench)
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <exception>

#include "synthetic_box.h"

static const double centers[8][3] = {
  {2.4257829890e-01, 1.3469574514e-02, 3.8313885004e-01}, {4.1465269048e-01, 6.7768972864e-02, 9.9312692973e-01},
  {4.8430804651e-01, 7.6533776604e-01, 3.1833815403e-02}, {3.0935481671e-02, 9.3264044027e-01, 8.8787953411e-01},
  {5.9132973039e-01, 4.7877868473e-01, 8.3335433660e-01}, {1.8633519681e-01, 7.3565270739e-01, 1.1505317181e-01},
  {6.9865863058e-01, 3.5560411138e-01, 6.3830000658e-01}, {9.0821050755e-01, 2.9400041480e-01, 2.6497158886e-01}};

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
  const int R = argc > 1 ? std::atoi(argv[1]) : 5, reps = argc > 2 ? std::atoi(argv[2]) : 20, dim = 3;
  const int averaging = argc > 3 ? std::atoi(argv[3]) : SMF_AVG_HARMONIC;      // 0 = none (27 values per cell on the active level)
  try
    {
      using namespace aspect_b200;
      double t0 = now();
      const HierarchyArrays hier = make_box_hierarchy(dim, R, "zero");
      const LevelArrays &act = hier.active;
      const std::vector<double> x = box_quadrature_points(dim, R);
      const std::size_t npts = (std::size_t)act.n_cells * 27;
      std::vector<double> eta(npts), force(npts * 3, 0.0);
      const double s = std::sqrt(1e4);
      for (std::size_t i = 0; i < npts; ++i)
        {
          double chi = 1.0;
          for (int k = 0; k < 8; ++k)
            {
              double d2 = 0;
              for (int d = 0; d < 3; ++d) d2 += (x[3 * i + d] - centers[k][d]) * (x[3 * i + d] - centers[k][d]);
              const double t = std::max(0.0, std::sqrt(d2) - 0.05);
              chi *= 1.0 - std::exp(-200.0 * t * t);
            }
          eta[i] = (s - 1.0 / s) * (1.0 - chi) + 1.0 / s;
          force[3 * i + 2] = -10.0 * (1.0 - chi);
        }
      const double t_host = now() - t0;
      t0 = now();
      StokesMatrixFreeHandlerB200 h(dim);
      h.parameters.linear_stokes_solver_tolerance = 1e-6;
      h.parameters.n_cheap_stokes_solver_steps = 200;
      h.parameters.n_expensive_stokes_solver_steps = 0;
      h.setup_dofs(hier);
      h.assemble(eta, averaging, 1.0);
      h.build_preconditioner();
      const double t_setup = now() - t0;

      auto su = h.initialize_dof_vector(0), sp = h.initialize_dof_vector(1), du = h.initialize_dof_vector(0), dp = h.initialize_dof_vector(1);
      std::vector<double> hu((std::size_t)act.n_dofs_u), hp((std::size_t)act.n_dofs_p);
      for (std::size_t i = 0; i < hu.size(); ++i) hu[i] = std::sin(0.37 * (double)i);
      for (std::size_t i = 0; i < hp.size(); ++i) hp[i] = std::cos(0.11 * (double)i);
      for (std::uint32_t c : act.constrained_u) hu[c] = 0.0;
      su->from_host(hu.data()); sp->from_host(hp.data());
      BlockVector src(*su, *sp), dst(*du, *dp);

      const auto sync = [&]() { h.synchronize(); };
      const auto time_it = [&](const char *name, const std::function<void()> &fn, double dofs) {
        for (int i = 0; i < 3; ++i) fn();
        sync();
        const double a = now();
        for (int i = 0; i < reps; ++i) fn();
        sync();
        const double t = (now() - a) / reps;
        std::printf("  %-22s %9.3f ms  %8.2f GDoF/s\n", name, 1e3 * t, dofs / t / 1e9);
        return t;
      };
      std::printf("nsinker box R=%d: %lld cells, %lld + %lld DoFs; host arrays %.1f s, device set-up %.1f s\n", R, (long long)act.n_cells,
                  (long long)act.n_dofs_u, (long long)act.n_dofs_p, t_host, t_setup);
      const double nu = (double)act.n_dofs_u, np = (double)act.n_dofs_p;
      time_it("stokes_vmult", [&]() { h.stokes_matrix.vmult(dst, src); }, nu + np);
      time_it("A_block (active)", [&]() { h.A_block_matrix.vmult(*du, *su); }, nu);
      time_it("BT_block", [&]() { h.BT_block.vmult(dst, src); }, nu);
      time_it("Schur mass (active)", [&]() { h.Schur_complement_block_matrix.vmult(*dp, *sp); }, np);
      for (int l = R; l >= std::max(0, R - 2); --l)
        {
          auto lu = h.initialize_dof_vector(0, l), ld = h.initialize_dof_vector(0, l);
          *lu = 1.0;
          char name[32];
          std::snprintf(name, sizeof name, "A_block level %d", l);
          time_it(name, [&]() { h.mg_matrices_A_block[(std::size_t)l].vmult(*ld, *lu); }, (double)hier.levels[(std::size_t)l].n_dofs_u);
        }
      // one solve (sets the smoothers up, which the V-cycles below need)
      auto rhs_u = h.initialize_dof_vector(0);
      h.assemble_stokes_rhs(force, *rhs_u);
      std::vector<double> rhs((std::size_t)(act.n_dofs_u + act.n_dofs_p), 0.0), sol;
      rhs_u->to_host(rhs.data());
      double a = now();
      SolverOutputs out = h.solve(rhs, sol);
      const double t_first = now() - a;
      a = now();
      std::fill(sol.begin(), sol.end(), 0.0);
      out = h.solve(rhs, sol);
      std::printf("  solve: %d iterations, %.3f s (first call %.3f s), converged %d, host buffers\n", out.iterations, now() - a, t_first, (int)out.converged);
      time_it("prec_A (V-cycle)", [&]() { h.prec_A.vmult(*du, *su); }, nu);
      time_it("prec_Schur (V-cycle)", [&]() { h.prec_Schur.vmult(*dp, *sp); }, np);
      return 0;
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
}
