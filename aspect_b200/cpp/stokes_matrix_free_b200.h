// C++ host side of the B200 block-GMG Stokes path: the mirror of ASPECT's StokesMatrixFreeHandler over the C ABI.
//
// Same life cycle, names and operator objects as the reference
//   /root/reference/include/aspect/simulator/solver/stokes_matrix_free.h:46-98   (setup_dofs / assemble / build_preconditioner)
//   /root/reference/include/aspect/simulator/solver/interface.h:36-101            (solve -> SolverOutputs, name())
//   stokes_matrix_free_local_smoothing.h:211-236 (stokes_matrix, A_block_matrix, BT_block, Schur_complement_block_matrix,
//                                                 mg_matrices_A_block, mg_matrices_Schur_complement, prec_A, prec_Schur)
// so that the ASPECT-side adapter of INTEGRATION.md is this class with the deal.II objects as the source of the arrays.
// Header only, C++17, no dependency besides include/smf_b200.h; errors become exceptions (the reference throws, too);
// SMF_NO_CONVERGENCE is reported through SolverOutputs::converged like SolverControl::NoConvergence is caught at
// local_smoothing.cc:1317.  aspect_b200/handler.py is the same mirror in Python (used by the tests and the bench).
#ifndef ASPECT_B200_STOKES_MATRIX_FREE_B200_H
#define ASPECT_B200_STOKES_MATRIX_FREE_B200_H

#include <cstdint>
#include <cstdlib>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "smf_b200.h"

namespace aspect_b200
{
  /** what one level of setup_dofs() hands over: the arrays MatrixFree holds (cell->DoF maps, constrained DoFs, geometry) */
  struct LevelArrays
  {
    std::int64_t n_cells = 0, n_dofs_u = 0, n_dofs_p = 0;
    std::vector<std::uint32_t> cell_dofs_u, cell_dofs_p, constrained_u, constrained_p;
    std::vector<double> cell_size;                        // Cartesian levels
    std::vector<double> inverse_jacobians, jxw;           // curved levels (geometry = 1)
  };

  /** the hierarchy: levels 0..L-1, the active mesh, and the parent->children tables of levels 1..L-1 */
  struct HierarchyArrays
  {
    int dim = 3;
    std::vector<LevelArrays> levels;
    LevelArrays active;
    std::vector<std::vector<std::int32_t>> children;      // children[l] for l >= 1 ([n_cells(l-1)][2^dim])
  };

  /** StokesSolver::SolverOutputs (interface.h:36-58) plus the counters the reference prints */
  struct SolverOutputs
  {
    double initial_nonlinear_residual = 0, final_linear_residual = 0, pressure_normalization_adjustment = 0;
    int iterations = 0, iterations_expensive = 0, n_A_preconditioner = 0, n_S_preconditioner = 0;
    bool converged = false;
  };

  class StokesMatrixFreeHandlerB200;

  /** LinearAlgebra::distributed::Vector<double> analogue in HBM */
  class DeviceVector
  {
  public:
    DeviceVector(smf_ctx *ctx, std::int64_t n) : ctx(ctx), n(n)
    {
      if (smf_vec_alloc(ctx, n, &ptr) != SMF_OK) throw std::runtime_error(smf_last_error(ctx));
    }
    ~DeviceVector() { if (ptr) smf_vec_free(ctx, ptr); }
    DeviceVector(const DeviceVector &) = delete;
    DeviceVector &operator=(const DeviceVector &) = delete;
    DeviceVector &operator=(double v) { check(smf_vec_set(ctx, ptr, v, n)); return *this; }
    void from_host(const double *h) { check(smf_vec_h2d(ctx, ptr, h, n)); check(smf_synchronize(ctx)); }
    void to_host(double *h) const { check(smf_vec_d2h(ctx, h, ptr, n)); }
    std::int64_t size() const { return n; }
    double *data() { return ptr; }
    const double *data() const { return ptr; }
  private:
    void check(int rc) const { if (rc != SMF_OK) throw std::runtime_error(smf_last_error(ctx)); }
    smf_ctx *ctx; std::int64_t n; double *ptr = nullptr;
  };
  using BlockVector = std::pair<DeviceVector &, DeviceVector &>;          // (velocity, pressure)

  /** duck-typed operator object with vmult(dst, src), as the Krylov driver and BlockSchurPreconditioner expect */
  template <class Dst, class Src> struct Operator
  {
    std::function<void(Dst &, const Src &)> apply;
    void vmult(Dst &dst, const Src &src) const { apply(dst, src); }
  };

  class StokesMatrixFreeHandlerB200
  {
  public:
    /** "Solver parameters / Stokes solver parameters" (source/simulator/parameters.cc:439-584), reference defaults */
    struct Parameters
    {
      double linear_stokes_solver_tolerance = 1e-7;
      int n_cheap_stokes_solver_steps = 1000;
      int n_expensive_stokes_solver_steps = 1000;
      int stokes_gmres_restart_length = 100;
      bool use_idr = false;                  // "Krylov method for cheap solver steps = IDR(s)"
      int idr_s_parameter = 2;
      double linear_solver_A_block_tolerance = 1e-2;
      double linear_solver_S_block_tolerance = 1e-6;
      /** "Stokes GMG type = global coarsening" (parameters.cc:448-453): on uniformly refined meshes the coarsening sequence is the
          sequence of refinement levels and this hierarchy serves both types (handler.py::_check_gmg_type); name() follows. */
      bool global_coarsening = false;
    } parameters;

    explicit StokesMatrixFreeHandlerB200(int dim, int device = 0) : dim(dim), device(device) {}
    ~StokesMatrixFreeHandlerB200() { if (ctx) smf_destroy(ctx); }
    StokesMatrixFreeHandlerB200(const StokesMatrixFreeHandlerB200 &) = delete;
    StokesMatrixFreeHandlerB200 &operator=(const StokesMatrixFreeHandlerB200 &) = delete;

    std::string name() const { return parameters.global_coarsening ? "GMG-GC" : "GMG"; }

    /** setup_dofs() (local_smoothing.cc:1450-1851): index maps, constraints, geometry, transfers of every level */
    void setup_dofs(const HierarchyArrays &h)
    {
      if (ctx) { smf_destroy(ctx); ctx = nullptr; }
      n_levels = (int)h.levels.size();
      smf_config cfg{dim, 2, n_levels, device};
      if (smf_create(&cfg, &ctx) != SMF_OK || !ctx)
        throw std::runtime_error("smf_create failed: a CUDA device is required (there is no CPU fallback)");
      char id[SMF_UNIQUE_ID_BYTES] = {0};
      check(smf_comm_init(ctx, 0, 1, id));
      if (const char *e = std::getenv("SMF_WARP_BRICK")) check(smf_set_option(ctx, SMF_OPT_WARP_BRICK_KERNEL, std::atoi(e)));   // kernel experiments
      if (const char *e = std::getenv("SMF_COARSE_TAIL")) check(smf_set_option(ctx, SMF_OPT_COARSE_TAIL, std::atoi(e)));
      for (int level = 0; level <= n_levels; ++level)
        {
          const LevelArrays &L = level < n_levels ? h.levels[level] : h.active;
          const bool curved = !L.inverse_jacobians.empty();
          smf_level_mesh m{L.n_cells, L.n_dofs_u, L.n_dofs_p, L.cell_dofs_u.data(), L.cell_dofs_p.data(),
                           (std::int64_t)L.constrained_u.size(), L.constrained_u.data(),
                           (std::int64_t)L.constrained_p.size(), L.constrained_p.data(),
                           curved ? 1 : 0, L.cell_size.empty() ? nullptr : L.cell_size.data(),
                           curved ? L.inverse_jacobians.data() : nullptr, curved ? L.jxw.data() : nullptr};
          check(smf_set_level(ctx, level < n_levels ? level : SMF_ACTIVE, &m));
        }
      for (int level = 1; level < n_levels; ++level) check(smf_set_transfer(ctx, level, h.children[level].data()));
      check(smf_set_copy_indices(ctx, nullptr, nullptr));
      n_u = h.active.n_dofs_u; n_p = h.active.n_dofs_p;
      level_sizes.clear();
      for (const LevelArrays &L : h.levels) level_sizes.emplace_back(L.n_dofs_u, L.n_dofs_p);
      make_operators();
    }

    /** assemble() = evaluate_material_model(): the plugin's viscosities at the velocity quadrature points of the active
        cells ([n_cells][3^dim]); averaging: SMF_AVG_* (`Material averaging`); the array work runs on the device */
    void assemble(const std::vector<double> &viscosity_q, int averaging, double pressure_scaling_, bool is_compressible = false)
    {
      double mm[2];
      check(smf_fill_cell_data(ctx, viscosity_q.data(), averaging, pressure_scaling_, is_compressible ? 1 : 0, mm));
      minimum_viscosity = mm[0]; maximum_viscosity = mm[1]; pressure_scaling = pressure_scaling_;
    }

    /** build_preconditioner(): operator diagonals of every level (local_smoothing.cc:1855-1867) */
    void build_preconditioner() { check(smf_compute_diagonals(ctx)); }

    /** velocity block of the right-hand side from the plugins' force density at the quadrature points (assemblers/stokes.cc:387-396) */
    void assemble_stokes_rhs(const std::vector<double> &force_q, DeviceVector &rhs_u) { check(smf_assemble_rhs(ctx, force_q.data(), rhs_u.data())); }

    /** solve() (interface.h:63-101; local_smoothing.cc:758-1446) with host vectors [u | p]; solution_vector: in = initial
        guess, out = solution (pressure in physical units) */
    SolverOutputs solve(const std::vector<double> &system_rhs, std::vector<double> &solution_vector)
    {
      solution_vector.resize((std::size_t)(n_u + n_p), 0.0);
      if (pressure_scaling != 1.0)
        for (std::int64_t i = n_u; i < n_u + n_p; ++i) solution_vector[(std::size_t)i] /= pressure_scaling;     // local_smoothing.cc:975
      smf_solve_params prm{parameters.linear_stokes_solver_tolerance, parameters.n_cheap_stokes_solver_steps,
                           parameters.stokes_gmres_restart_length, parameters.use_idr ? 1 : 0, parameters.idr_s_parameter,
                           parameters.n_expensive_stokes_solver_steps, parameters.linear_solver_A_block_tolerance,
                           parameters.linear_solver_S_block_tolerance, pressure_scaling != 1.0 ? 1 : 0, 0, 0.0};
      smf_solve_result res{};
      const int rc = smf_solve_host(ctx, &prm, solution_vector.data(), solution_vector.data() + n_u, system_rhs.data(),
                                    system_rhs.data() + n_u, &res);
      if (rc != SMF_OK && rc != SMF_NO_CONVERGENCE) throw std::runtime_error(smf_last_error(ctx));
      SolverOutputs out;
      out.initial_nonlinear_residual = res.initial_residual; out.final_linear_residual = res.final_residual;
      out.pressure_normalization_adjustment = res.pressure_normalization_adjustment;
      out.iterations = res.iterations; out.iterations_expensive = res.iterations_expensive;
      out.n_A_preconditioner = res.n_A_preconditioner; out.n_S_preconditioner = res.n_S_preconditioner;
      out.converged = rc == SMF_OK;
      return out;
    }

    std::unique_ptr<DeviceVector> initialize_dof_vector(int block, int level = SMF_ACTIVE) const
    {
      const auto sz = level == SMF_ACTIVE ? std::make_pair(n_u, n_p) : level_sizes.at((std::size_t)level);
      return std::make_unique<DeviceVector>(ctx, block == 0 ? sz.first : sz.second);
    }
    /** wait for the work queued on the context's stream */
    void synchronize() const { check(smf_synchronize(ctx)); }
    /** user constraints of MGConstrainedDoFs on a multigrid level (no-normal-flux on a curved boundary,
        local_smoothing.cc:1719-1789) or lines of the active constraints (level = SMF_ACTIVE): closed lines u_k = sum w_j u_mj */
    void set_constraint_lines(int level, int family, const std::vector<std::uint32_t> &dof, const std::vector<std::uint32_t> &row_start,
                              const std::vector<std::uint32_t> &master, const std::vector<double> &weight)
    { check(smf_set_constraints(ctx, level, family, (std::int64_t)dof.size(), dof.data(), row_start.data(), master.data(), weight.data())); }
    /** masters' entries of the constrained level operator's diagonal (MatrixFreeTools::compute_diagonal), after build_preconditioner() */
    void set_inverse_diagonal_entries(int level, int which, const std::vector<std::uint32_t> &dofs, const std::vector<double> &inverse_diagonal)
    { check(smf_set_inverse_diagonal_entries(ctx, level, which, (std::int64_t)dofs.size(), dofs.data(), inverse_diagonal.data())); }
    std::size_t get_cell_data_memory_consumption() const { return (std::size_t)smf_device_bytes(ctx); }
    std::int64_t n_velocity_dofs() const { return n_u; }
    std::int64_t n_pressure_dofs() const { return n_p; }
    double minimum_viscosity = 0, maximum_viscosity = 0, pressure_scaling = 1;

    // the operator objects of stokes_matrix_free_local_smoothing.h:211-236
    Operator<BlockVector, BlockVector> stokes_matrix, BT_block, preconditioner_cheap;
    Operator<DeviceVector, DeviceVector> A_block_matrix, Schur_complement_block_matrix, prec_A, prec_Schur;
    std::vector<Operator<DeviceVector, DeviceVector>> mg_matrices_A_block, mg_matrices_Schur_complement;

  private:
    void check(int rc) const { if (rc != SMF_OK) throw std::runtime_error(ctx ? smf_last_error(ctx) : "smf: no context"); }
    void make_operators()
    {
      stokes_matrix.apply = [this](BlockVector &d, const BlockVector &s) { check(smf_vmult_stokes(ctx, d.first.data(), d.second.data(), s.first.data(), s.second.data())); };
      BT_block.apply = [this](BlockVector &d, const BlockVector &s) { check(smf_vmult_BT(ctx, d.first.data(), s.second.data())); };
      preconditioner_cheap.apply = [this](BlockVector &d, const BlockVector &s) { check(smf_precondition(ctx, d.first.data(), d.second.data(), s.first.data(), s.second.data())); };
      A_block_matrix.apply = [this](DeviceVector &d, const DeviceVector &s) { check(smf_vmult_A(ctx, SMF_ACTIVE, d.data(), s.data())); };
      Schur_complement_block_matrix.apply = [this](DeviceVector &d, const DeviceVector &s) { check(smf_vmult_mass(ctx, SMF_ACTIVE, d.data(), s.data())); };
      prec_A.apply = [this](DeviceVector &d, const DeviceVector &s) { check(smf_vcycle_A(ctx, d.data(), s.data())); };
      prec_Schur.apply = [this](DeviceVector &d, const DeviceVector &s) { check(smf_vcycle_S(ctx, d.data(), s.data())); };
      mg_matrices_A_block.clear(); mg_matrices_Schur_complement.clear();
      for (int l = 0; l < n_levels; ++l)
        {
          mg_matrices_A_block.push_back({[this, l](DeviceVector &d, const DeviceVector &s) { check(smf_vmult_A(ctx, l, d.data(), s.data())); }});
          mg_matrices_Schur_complement.push_back({[this, l](DeviceVector &d, const DeviceVector &s) { check(smf_vmult_mass(ctx, l, d.data(), s.data())); }});
        }
    }
    int dim, device, n_levels = 0;
    smf_ctx *ctx = nullptr;
    std::int64_t n_u = 0, n_p = 0;
    std::vector<std::pair<std::int64_t, std::int64_t>> level_sizes;
  };
}  // namespace aspect_b200
#endif
