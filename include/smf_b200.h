/* smf_b200.h -- C ABI of the B200-native matrix-free GMG Stokes path.
 *
 * Drop-in boundary for ASPECT's `Stokes solver type = block GMG`
 * (StokesMatrixFreeHandler, /root/reference/include/aspect/simulator/solver/stokes_matrix_free.h:46-98).
 * Every entry point names the reference interface it replaces (paths relative to the ASPECT tree).
 * Plain pointers and sizes only; all functions return 0 (SMF_OK) or an error code, never abort.
 * One context per rank/GPU, single caller thread, all calls stream-ordered.
 *
 * Conventions
 *  - level = 0..n_levels-1 are multigrid levels (MGLevelObject index); level = SMF_ACTIVE (-1) is the
 *    active mesh (the operators held in stokes_matrix / A_block_matrix / BT_block /
 *    Schur_complement_block_matrix, stokes_matrix_free_local_smoothing.h:211-216).
 *  - cell->DoF maps are what MatrixFree stores: lexicographic node order (x fastest),
 *    component-major [c0: 3^dim][c1: 3^dim]..., 32-bit indices into the rank-local [owned | ghost] vector.
 *    The library never assumes a numbering beyond "the dim components of one node are consecutive DoFs"
 *    (true for FESystem(FE_Q(2),dim)); it verifies that and returns SMF_ERR_UNSUPPORTED otherwise.
 *  - vectors passed to *device* entry points are device pointers to doubles (smf_vec_alloc or any CUDA
 *    allocation on the context's device); *_host entry points take host pointers and copy.
 *  - number type: double everywhere (GMGNumberType, stokes_matrix_free.h:40).
 */
#ifndef SMF_B200_H
#define SMF_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SMF_OK 0
#define SMF_ERR_INVALID 1       /* bad argument / call order                          */
#define SMF_ERR_CUDA 2          /* CUDA runtime error (see smf_last_error)            */
#define SMF_ERR_UNSUPPORTED 3   /* feature outside the implemented path                */
#define SMF_NO_CONVERGENCE 4    /* solver hit its iteration limit (SolverControl::NoConvergence) */
#define SMF_ACTIVE (-1)

typedef struct smf_ctx smf_ctx;

typedef struct {
  int dim;        /* 2 or 3 */
  int degree;     /* Stokes velocity polynomial degree; only 2 (Q2-Q1) */
  int n_levels;   /* triangulation.n_global_levels() */
  int device;     /* CUDA device ordinal */
} smf_config;

/* create_matrix_free_solver + constructor (source/simulator/solver/stokes_matrix_free.cc:29-62) */
int smf_create(const smf_config *cfg, smf_ctx **out);
void smf_destroy(smf_ctx *ctx);
/* message of the last failing call (replaces the C++ exception text) */
const char *smf_last_error(const smf_ctx *ctx);
/* implementation switches (all on by default); set before smf_set_level */
#define SMF_OPT_STRUCTURED_BRICKS 0   /* conflict-free staging layout for batches that are 4x4x2 Morton bricks */
#define SMF_OPT_CUDA_GRAPHS 1         /* replay the coarse-level part of the V-cycles from CUDA graphs */
#define SMF_OPT_MAX_GRID_BLOCKS 4      /* > 0: cap on the grid of the persistent kernels (testing aid; default 0 = one resident wave) */
#define SMF_OPT_PIPELINED_KERNEL 3     /* structured levels: persistent, cp.async-pipelined velocity kernel (default 1) */
#define SMF_OPT_P2P_EXCHANGE 5        /* halo exchange as one kernel over NVLink peer memory (default 1); 0 = pack + NCCL send/recv + accumulate */
#define SMF_OPT_CURVED_SUMFAC 6        /* curved cells (geometry = 1): sum-factorised kernel (default 1); 0 = the general per-cell restatement kernel */
#define SMF_OPT_CURVED_TMA 7           /* curved cells: geometry stream staged in shared memory by cp.async.bulk (TMA) + mbarrier (default 0) */
#define SMF_OPT_WARP_BRICK_KERNEL 8    /* structured levels, A block / Stokes: one warp per brick with the three velocity components of a cell in one thread
                                         (no cross-warp stress exchange, no block barriers) instead of one warp per component */
#define SMF_OPT_COARSE_TAIL 9          /* Schur-complement V-cycle: the levels that fit one thread block's shared memory run in ONE kernel launch
                                         (default 1; one rank, one viscosity per cell on those levels) */
#define SMF_OPT_NONSYMMETRIC_A_SOLVER 10 /* `Force nonsymmetric A block solver` (parameters.cc): BiCGStab instead of CG for the A block in the expensive phase
                                         (block_stokes_preconditioner.h:119-132); implied by smf_set_free_surface_faces */
#define SMF_OPT_LATTICE_SCHUR 11        /* Schur-complement V-cycle on box levels: level vectors in lattice order inside the cycle, node-centred 27-point mass
                                         stencil with the Chebyshev update fused, one CUDA graph per cycle (default 1; 3D, Cartesian, one viscosity per cell) */
#define SMF_OPT_LATTICE_VELOCITY 12     /* experimental (default 0): A block of structured box levels on lattice-ordered copies of the vectors, bricks staged by
                                         TMA box loads (cp.async.bulk.tensor, x de-interleaved by the element stride) instead of node lists + cp.async */
#define SMF_OPT_REPLICATE_COARSE_A 13   /* several ranks, box subdomains (default 1): the coarse levels of prec_A (global mesh up to 70,000 cells) are computed by every
                                         rank on the GLOBAL mesh in a private child context -- one all-reduce per cycle instead of 9 halo exchanges per level */
#define SMF_OPT_ORTHO_SINGLE_PASS 2   /* GMRES: CGS2 with the second projection formed from the Gram matrix (default 1); 0 = two sweeps */
int smf_set_option(smf_ctx *ctx, int option, int value);
/* run all work on this CUDA stream (cudaStream_t); default: a stream owned by the context */
int smf_set_stream(smf_ctx *ctx, void *cuda_stream);

/* One level of setup_dofs() (stokes_matrix_free_local_smoothing.cc:1450-1851): what the per-level
 * MatrixFree<dim,double> object holds (dof_info.dof_indices, constrained DoFs, mapping_info). */
typedef struct {
  int64_t n_cells;
  int64_t n_dofs_u, n_dofs_p;        /* local vector sizes (owned + ghost) */
  const uint32_t *cell_dofs_u;       /* [n_cells][dim*3^dim] */
  const uint32_t *cell_dofs_p;       /* [n_cells][2^dim] */
  int64_t n_constrained_u;           /* homogeneous constraints: Dirichlet / no-normal-flux on axis-aligned
                                        faces / MGConstrainedDoFs boundary + refinement-edge indices */
  const uint32_t *constrained_u;
  int64_t n_constrained_p;
  const uint32_t *constrained_p;
  int geometry;                      /* 0: Cartesian, one cell size for the level (cell_size[dim]);
                                        1: general cells, per-quadrature-point geometry below: curved cells (3D: the
                                           sum-factorised kernel; 2D: the general per-cell kernel) and the active level of an
                                           adaptive mesh (cells of different sizes in one launch)               */
  const double *cell_size;
  /* geometry = 1: what MatrixFree's MappingInfo holds for the 3^dim-point Gauss rule of the cell loops
   * (matrix_free_operators.cc:384-385), in the cell order of cell_dofs_*, q lexicographic (x fastest):
   * inverse_jacobians [n_cells][3^dim][dim][dim] with entry [r][d] = d xi_r / d x_d (FEEvaluation's inverse_jacobian:
   * grad_x phi = J^-T grad_xi phi), jxw [n_cells][3^dim] = det(J) * quadrature weight.  Both are COPIED.
   * Adds the G term of SURVEY 8(d) to the algorithmic bytes: 8 (dim^2 + 1) 3^dim bytes per cell and application. */
  const double *inverse_jacobians;
  const double *jxw;
} smf_level_mesh;
int smf_set_level(smf_ctx *ctx, int level, const smf_level_mesh *mesh);

/* MGTransferMatrixFree::build (local_smoothing.cc:1843-1850): parent->children table of level `fine_level`
 * ([n_cells(fine_level-1)][2^dim] cell ids on fine_level, deal.II child order).  On adaptively refined meshes a parent that
 * is not refined has -1 in all its entries (it takes no part in the transfer). */
int smf_set_transfer(smf_ctx *ctx, int fine_level, const int32_t *children);
/* copy_to_mg / copy_from_mg index map: level-(n_levels-1) DoF of every active DoF (NULL = identity). */
int smf_set_copy_indices(smf_ctx *ctx, const uint32_t *active_to_level_u, const uint32_t *active_to_level_p);

/* ---- adaptively refined meshes (hanging nodes, local smoothing with refinement edges) ---------------------------------
 * What setup_dofs() builds on an adaptive mesh and the matrix-free operators / the V-cycle apply:
 *
 * smf_set_constraints: the closed lines u_k = sum_j w_j u_{m_j} of the level's AffineConstraints that are not plain
 *   homogeneous Dirichlet rows -- the hanging-node constraints of DoFTools::make_hanging_node_constraints
 *   (local_smoothing.cc:1508, :1580; ACTIVE level) -- in CSR form: line i constrains constrained_dof[i] to the masters
 *   master_dof[row_start[i] .. row_start[i+1]) with weights weight[..].  Masters are unconstrained DoFs (the lines are
 *   closed; Dirichlet masters dropped).  The operators then act as MatrixFreeOperators::Base does: read_dof_values
 *   resolves the lines, distribute_local_to_global applies their transpose, constrained rows get dst = src
 *   (A_mf = C^T A C + I_constrained).  family 0: velocity, 1: pressure.  n_lines = 0 clears.
 *   level = SMF_ACTIVE: hanging nodes and / or no-normal-flux lines of the active constraints.  level >= 0: the USER
 *   CONSTRAINTS of MGConstrainedDoFs on that multigrid level -- no-normal-flux on a curved boundary,
 *   VectorTools::compute_no_normal_flux_constraints_on_level (stokes_matrix_free_local_smoothing.cc:1719-1789): the level
 *   operator and its smoother then work with the lines, smf_prolongate distributes the COARSE level's lines on its source
 *   before it interpolates (MGTransferMatrixFree::prolongate), and the masters' diagonal entries come in through
 *   smf_set_inverse_diagonal_entries (below).
 * smf_set_refinement_edges: MGConstrainedDoFs::get_refinement_edge_indices of one level (:1733-1773).  The level operators
 *   then treat these DoFs like Dirichlet DoFs (A_l = (I - E) A_bc (I - E) + E, as Base::vmult with edge_constrained_indices),
 *   and the V-cycles add the interface products vmult_interface_down / vmult_interface_up (mg::Matrix edge matrices,
 *   :879-911).  Levels keep only the boundary indices in smf_level_mesh::constrained_u.
 * smf_set_level_copy_indices: MGTransferMatrixFree's copy_indices of one level: active DoF active_idx[i] <-> level DoF
 *   level_idx[i] (copy_to_mg / copy_from_mg on an adaptive mesh).  Setting them on any level switches the V-cycles to the
 *   adaptive algorithm (defects of all levels filled from the active vector, restriction ADDS, solutions of all levels
 *   copied back). */
int smf_set_constraints(smf_ctx *ctx, int level, int family, int64_t n_lines, const uint32_t *constrained_dof,
                        const uint32_t *row_start, const uint32_t *master_dof, const double *weight);
int smf_set_refinement_edges(smf_ctx *ctx, int level, int64_t n_u, const uint32_t *edge_u, int64_t n_p, const uint32_t *edge_p);
int smf_set_level_copy_indices(smf_ctx *ctx, int level, int family, int64_t n, const uint32_t *active_idx, const uint32_t *level_idx);

/* ---- distributed runs: one context per MPI rank / GPU ----------------------------------------------------------
 * Replaces LinearAlgebra::distributed::Vector::update_ghost_values / compress(add) around every cell loop
 * (matrix_free_operators.cc:626, 724, 856, 1063; local_smoothing.cc:645, 746, 1407) and the MPI_Allreduce of every
 * inner product.  Vectors hold every DoF the rank's cells touch; copies of a shared DoF are kept equal on all ranks.
 * smf_comm_get_unique_id: rank 0 creates the id, the host broadcasts it (MPI_Bcast / torch.distributed), every rank
 * passes it to smf_comm_init (collective; NCCL over NVLink). */
#define SMF_UNIQUE_ID_BYTES 128
int smf_comm_get_unique_id(char *id);
int smf_comm_init(smf_ctx *ctx, int rank, int world, const char *id);
/* Utilities::MPI::Partitioner of one level and DoF family (0: velocity, 1: pressure): for every neighbour rank
 * (ascending) the shared nodes, given by their first local DoF index, in an order both sides agree on (e.g. by
 * global index).  global_dof_ids (optional, [n_dofs]): partition-independent DoF numbers, used for the Chebyshev
 * eigenvalue start vector.  Collective.  Call after smf_set_level and before smf_set_transfer. */
typedef struct {
  int n_neighbors;
  const int *neighbor_ranks;
  const int64_t *counts;
  const uint32_t *const *nodes;
  const int64_t *global_dof_ids;
} smf_exchange;
int smf_set_exchange(smf_ctx *ctx, int level, int family, const smf_exchange *x);
/* compress(add) + update_ghost_values: every copy of a shared DoF := sum over the ranks (device vector). Collective. */
int smf_exchange_add(smf_ctx *ctx, int level, int family, double *vec);

/* OperatorCellData (include/aspect/simulator/solver/matrix_free_operators.h:91-208) +
 * set_cell_data (matrix_free_operators.cc:353-359, 779-783, 966-970): the table is COPIED to the device.
 * viscosity is [n_cells][n_per_cell] in the cell order of smf_set_level, n_per_cell = 1 or 3^dim. */
typedef struct {
  const double *viscosity;
  int n_per_cell;
  double pressure_scaling;
  int is_compressible;
} smf_cell_data;
int smf_set_cell_data(smf_ctx *ctx, int level, const smf_cell_data *data);

/* The array work of evaluate_material_model() on the device: viscosity_q = the material plugin's output at
 * the 3^dim velocity quadrature points of every active cell, host array [n_cells][3^dim].  Averaging over the points of a
 * cell (source/material_model/interface.cc:523-650; 0 none, 1 arithmetic, 2 harmonic, 3 geometric; 4 = no averaging, DGQ1 space), cellwise projection
 * onto DGQ0 (source/utilities.cc:2738-2812), the active table (matrix_free_operators.cc:227-246: n_q values per cell without
 * averaging, else 1), min / max after averaging (:248-262) and the level tables by interpolate_to_mg = mean of the children
 * (local_smoothing.cc:211-294).  Needs smf_set_level on all levels and smf_set_transfer; uniformly refined meshes.
 * Equivalent to smf_set_cell_data on the active mesh and on every level. */
#define SMF_AVG_NONE 0
#define SMF_AVG_ARITHMETIC 1
#define SMF_AVG_HARMONIC 2
#define SMF_AVG_GEOMETRIC 3
#define SMF_AVG_PROJECT_Q1 4   /* `project to Q1 [only viscosity]`: DGQ1 projection space, 3^dim values per cell on the active mesh and on every level,
                                 clamped to the range of the raw values (matrix_free_operators.cc:306-325, local_smoothing.cc:99-111, :274-289) */
int smf_fill_cell_data(smf_ctx *ctx, const double *viscosity_q, int averaging, double pressure_scaling, int is_compressible,
                       double *min_max_out);
/* host copy [n_cells][n_per_cell] of a level's table as the kernels see it */
int smf_get_cell_data(smf_ctx *ctx, int level, double *host_out, int *n_per_cell_out);

/* The remaining fields of OperatorCellData (matrix_free_operators.h:128-196), ACTIVE level only as in the reference
 * (local_smoothing.cc:296-612; the level operators never see them, :308-314): Newton-derivative and prescribed-dilation
 * tables, [n_cells][n_q][n_sym] or [n_cells][n_q] in the cell order of smf_set_level, n_q = 3^dim, symmetric tensors in
 * deal.II's unrolled order 00, 11, (22,) 01, (02, 12).  Tables are COPIED.  They enter StokesOperator::local_apply
 * (matrix_free_operators.cc:400-446, :464-538); enable_prescribed_dilation also makes the A block deviatoric
 * (:999-1005).  Passing NULL clears them.  The free-surface face term (:563-602) has its own entry point below
 * (apply_stabilization_free_surface_faces in this struct must stay 0). */
typedef struct {
  int enable_newton_derivatives;
  int symmetrize_newton_system;
  int average_newton_factors;
  int enable_prescribed_dilation;
  int apply_stabilization_free_surface_faces;
  const double *strain_rate_table;
  const double *newton_factor_wrt_strain_rate_table;
  const double *newton_factor_wrt_pressure_table;
  const double *dilation_lhs_term_table;
  const double *dilation_derivative_wrt_strain_rate_table;
  const double *dilation_derivative_wrt_pressure_table;
} smf_newton_data;
int smf_set_newton_data(smf_ctx *ctx, const smf_newton_data *data);

/* OperatorCellData::apply_stabilization_free_surface_faces + free_surface_stabilization_term_table
 * (matrix_free_operators.h:181-196; filled by local_smoothing.cc:528-612): the boundary faces with a free-surface indicator,
 * given as (active cell, local face number 2 d + side in deal.II's order), and for each of them the vector
 * g = pressure perturbation * g_hat at the 3^(dim-1) face Gauss points ([n_faces][3^(dim-1)][dim], q lexicographic over the
 * two tangential directions, lower direction fastest).  StokesOperator::local_apply_boundary_face (matrix_free_operators.cc:
 * 563-602) then adds  -int_face (g . u)(n . v)  to the velocity rows of the ACTIVE Stokes operator, which makes the A block
 * non-symmetric: the expensive phase solves it with BiCGStab instead of CG (block_stokes_preconditioner.h:119-132).
 * Cartesian cells (geometry = 0).  n_faces = 0 clears. */
int smf_set_free_surface_faces(smf_ctx *ctx, int64_t n_faces, const uint32_t *cells, const int32_t *face_numbers,
                               const double *stabilization_term_table);

/* ---- device vectors ------------------------------------------------------------------------- */
int smf_vec_alloc(smf_ctx *ctx, int64_t n, double **dev_ptr);   /* initialize_dof_vector */
int smf_vec_free(smf_ctx *ctx, double *dev_ptr);
int smf_vec_h2d(smf_ctx *ctx, double *dev, const double *host, int64_t n);
int smf_vec_d2h(smf_ctx *ctx, double *host, const double *dev, int64_t n);
int smf_vec_set(smf_ctx *ctx, double *dev, double value, int64_t n);   /* Vector::operator=(value) */
int smf_vec_copy(smf_ctx *ctx, double *dst, const double *src, int64_t n);   /* Vector::operator=(const Vector &), device to device */
int smf_synchronize(smf_ctx *ctx);

/* ---- operators (device pointers) ------------------------------------------------------------ */
/* StokesOperator::vmult (matrix_free_operators.cc:606-627, local_apply :376-547), active level */
int smf_vmult_stokes(smf_ctx *ctx, double *dst_u, double *dst_p, const double *src_u, const double *src_p);
/* StokesMatrixFreeHandler::correct_stokes_rhs (stokes_matrix_free_local_smoothing.cc:617-754): rhs -= Stokes(u0) with
   u0 read by read_dof_values_plain (boundary values are kept), nothing is added to constrained rows.  u0 holds the
   inhomogeneous boundary values and zeros elsewhere (AffineConstraints::distribute of the zero vector, :632-642). */
int smf_correct_stokes_rhs(smf_ctx *ctx, double *rhs_u, double *rhs_p, const double *u0_u, const double *u0_p);
/* ABlockOperator::vmult (matrix_free_operators.cc:1056-1064, :1029-1052) */
int smf_vmult_A(smf_ctx *ctx, int level, double *dst, const double *src);
/* BTBlockOperator::vmult (matrix_free_operators.cc:717-725, :668-701), active level */
int smf_vmult_BT(smf_ctx *ctx, double *dst_u, const double *src_p);
/* MassMatrixOperator::vmult (matrix_free_operators.cc:849-857, :787-845) */
int smf_vmult_mass(smf_ctx *ctx, int level, double *dst, const double *src);
/* host-buffer variant of the Stokes vmult (H2D, kernel, D2H on the context's stream, then synchronise) */
int smf_vmult_stokes_host(smf_ctx *ctx, double *dst_u, double *dst_p, const double *src_u, const double *src_p);

/* build_preconditioner(): ABlockOperator::compute_diagonal / MassMatrixOperator::compute_diagonal on every
 * level (local_smoothing.cc:1855-1867, matrix_free_operators.cc:1068-1098, :861-908) */
int smf_compute_diagonals(smf_ctx *ctx);
/* get_matrix_diagonal_inverse() (local_smoothing.cc:794, :822): which = 0 (A) or 1 (mass); host output */
int smf_get_inverse_diagonal(smf_ctx *ctx, int level, int which, double *host_out);
/* Overwrite entries of a level's inverse diagonal (after smf_compute_diagonals, before the smoothers are set up).
 * smf_compute_diagonals returns the diagonal of the operator WITHOUT the constraint lines of smf_set_constraints; on a level
 * that smooths with such lines (no-normal-flux on a curved boundary: user constraints of MGConstrainedDoFs,
 * stokes_matrix_free_local_smoothing.cc:1719-1789) the masters' entries of diag(C^T A C) -- what
 * MatrixFreeTools::compute_diagonal returns in the reference (matrix_free_operators.cc:1068-1098) -- come from the caller:
 * the ASPECT adapter passes deal.II's values, handler.py obtains them by probing the level operator with coloured unit
 * vectors.  values = INVERSE diagonal entries (> 0). */
int smf_set_inverse_diagonal_entries(smf_ctx *ctx, int level, int which, int64_t n, const uint32_t *idx, const double *values);

/* PreconditionChebyshev set-up + estimate_eigenvalues on every level, both hierarchies
 * (local_smoothing.cc:774-849).  Must follow smf_compute_diagonals. */
int smf_setup_smoothers(smf_ctx *ctx);
/* (lambda_min_estimate, lambda_max_estimate*1.2, cg iterations) per level; which = 0 (A) / 1 (mass) */
int smf_get_eigenvalue_estimates(smf_ctx *ctx, int level, int which, double *min_max_its);

/* MGTransfer prolongate / restrict_and_add (which = 0 velocity, 1 pressure), level vectors.  On several ranks smf_restrict_add ends in
 * the halo exchange of the coarse vector (compress(add)): dst must be zero on DoFs shared with other ranks when it is called (the
 * V-cycles zero it), otherwise the already consistent values there are summed once more. */
int smf_prolongate(smf_ctx *ctx, int which, int to_level, double *dst_fine, const double *src_coarse);
int smf_restrict_add(smf_ctx *ctx, int which, int from_level, double *dst_coarse, const double *src_fine);

/* PreconditionMG::vmult = one V-cycle (local_smoothing.cc:893-916): prec_A / prec_Schur */
int smf_vcycle_A(smf_ctx *ctx, double *dst_u, const double *src_u);
int smf_vcycle_S(smf_ctx *ctx, double *dst_p, const double *src_p);
/* BlockSchurPreconditioner::vmult, cheap variant (block_stokes_preconditioner.h:220-255) */
int smf_precondition(smf_ctx *ctx, double *dst_u, double *dst_p, const double *src_u, const double *src_p);

/* ---- solve() ------------------------------------------------------------------------------- */
typedef struct {
  double linear_solver_tolerance;        /* "Linear solver tolerance" (parameters.cc:491)              */
  int n_cheap_stokes_solver_steps;       /* "Number of cheap Stokes solver steps" (:515)               */
  int gmres_restart_length;              /* "GMRES solver restart length" (:538)                       */
  int krylov_method;                     /* "Krylov method for cheap solver steps" (:469): 0 = GMRES, 1 = IDR(s) */
  int idr_s_parameter;                   /* "IDR(s) parameter" (:483)                                  */
  int n_expensive_stokes_solver_steps;   /* "Maximum number of expensive Stokes solver steps" (:531)   */
  double linear_solver_A_block_tolerance;   /* "Linear solver A block tolerance" (:546)                */
  double linear_solver_S_block_tolerance;   /* "Linear solver S block tolerance" (:584)                */
  int rescale_pressure;                  /* 1: the returned pressure is multiplied by the pressure scaling (local_smoothing.cc:1421) */
  int normalize_pressure;                /* 1: Simulator::normalize_pressure with the weights of smf_set_pressure_normalization   */
  double surface_pressure;               /* "Surface pressure" (target of the surface normalisation)   */
} smf_solve_params;

typedef struct {
  int iterations;                        /* outer (cheap) iterations                                   */
  int n_A_preconditioner;                /* InverseVelocityBlock::n_iterations                         */
  int n_S_preconditioner;                /* SchurApproximation::n_iterations                           */
  double initial_residual;               /* outputs.initial_nonlinear_residual                         */
  double final_residual;                 /* outputs.final_linear_residual                              */
  double tolerance;                      /* absolute tolerance used (local_smoothing.cc:1024-1025)     */
  double setup_seconds;                  /* smoother set-up + eigenvalue estimation (host wall clock)  */
  double solve_seconds;                  /* Krylov iteration (host wall clock, synchronised)           */
  int iterations_expensive;              /* FGMRES iterations of the expensive phase ("cheap+expensive") */
  int n_A_expensive;                     /* inner CG iterations with the A block (expensive phase)     */
  int n_S_expensive;                     /* inner CG iterations with the pressure mass matrix          */
  double pressure_normalization_adjustment;   /* outputs.pressure_normalization_adjustment (0 without normalisation) */
} smf_solve_result;

/* StokesMatrixFreeHandler::solve() (local_smoothing.cc:758-1446) on device vectors: the cheap phase (GMRES or IDR(s) with
 * one V-cycle per block) and, if it fails or is skipped, the expensive phase (FGMRES with inner CG solves, :1317-1350).
 * rhs_{u,p}: right-hand side; sol_{u,p}: in = initial guess (pressure already divided by the pressure
 * scaling, constrained entries zero), out = solution.  Returns SMF_OK or SMF_NO_CONVERGENCE. */
int smf_solve(smf_ctx *ctx, const smf_solve_params *prm, double *sol_u, double *sol_p,
              const double *rhs_u, const double *rhs_p, smf_solve_result *result);
/* same with host buffers (the reference-facing plugin call: H2D of rhs/guess, solve, D2H of the solution) */
int smf_solve_host(smf_ctx *ctx, const smf_solve_params *prm, double *sol_u, double *sol_p,
                   const double *rhs_u, const double *rhs_p, smf_solve_result *result);

/* ---- callers either side of the solve ------------------------------------------------ */
/* Velocity block of the Stokes right-hand side, f_i = int force . phi_i (source/simulator/assemblers/stokes.cc:387-396;
 * cell loop source/simulator/assembly.cc:761-889): force_q = density * gravity from the plugins, host array
 * [n_cells][3^dim][dim]; rhs_u = device vector.  Constrained rows receive nothing; on several ranks the per-rank partial
 * sums are exchanged (compress(add)). */
int smf_assemble_rhs(smf_ctx *ctx, const double *force_q, double *rhs_u);
/* Simulator::normalize_pressure (source/simulator/helper_functions.cc:765-870): weights[i] = integral of the i-th pressure
 * shape function over the top boundary (`Pressure normalization = surface`) or over this rank's cells (`volume`),
 * host array [n_p]; NULL switches the normalisation off. */
int smf_set_pressure_normalization(smf_ctx *ctx, const double *weights);
/* sol_p *= pressure_scaling (local_smoothing.cc:1421), then, if normalize, sol_p += adjustment with
 * adjustment = -(integral of p) / area + surface_pressure (helper_functions.cc:843-850); returns the adjustment
 * (SolverOutputs::pressure_normalization_adjustment). */
int smf_postprocess_pressure(smf_ctx *ctx, double *sol_p, double pressure_scaling, int normalize, double surface_pressure,
                             double *adjustment_out);

/* Device-time profiler: the reference's "Execute solver timings" switch (local_smoothing.cc:60-65, :1119-1260)
 * and TimerOutput sections, done with CUDA event pairs on the context's stream around every launch.
 * mode 0 = off, 1 = operator kernels of the finest level only (Stokes, A), 2 = every launch.
 * smf_profile_read synchronises and returns accumulated milliseconds and launch counts per bucket:
 * 0 Stokes vmult, 1 A finest level, 2 A coarser levels, 3 pressure mass, 4 B^T, 5 MG transfer,
 * 6 Chebyshev vector updates, 7 GMRES orthogonalisation, 8 other vector kernels, 9 halo exchanges. */
#define SMF_PROFILE_BUCKETS 10
int smf_profile_enable(smf_ctx *ctx, int mode);
int smf_profile_read(smf_ctx *ctx, double *ms_out, int64_t *count_out);

/* bookkeeping the bench harness reports */
int64_t smf_kernel_launches(const smf_ctx *ctx);     /* kernels launched by this context so far */
int64_t smf_device_bytes(const smf_ctx *ctx);        /* get_*_memory_consumption() analogue      */
int64_t smf_exchange_count(const smf_ctx *ctx);      /* halo exchanges performed by this context so far */
/* The lattice detection of the Schur V-cycle as a host-only utility (no CUDA call, no context): given the pressure part of a level's
 * cell -> DoF map ([n_cells][8], 3D, lexicographic vertex order), decide whether the cells form a Cartesian box and, if so, return
 * the number of pressure nodes per direction and the integer coordinates of every cell ([n_cells][3]).  Returns 1 (box), 0 (not a box). */
int smf_detect_box_lattice(const uint32_t *cell_dofs_p, int64_t n_cells, int64_t n_dofs_p, int32_t *nodes_per_direction, int32_t *cell_xyz);
int smf_lattice_schur_active(const smf_ctx *ctx);    /* 1: prec_Schur runs on the lattice path (SMF_OPT_LATTICE_SCHUR; decided by smf_setup_smoothers) */
int smf_lattice_replicated_levels(const smf_ctx *ctx);
int smf_replicated_A_levels(const smf_ctx *ctx);          /* several ranks: coarse levels of prec_A computed redundantly on the global mesh (SMF_OPT_REPLICATE_COARSE_A) */   /* several ranks: coarse levels of the lattice path that every rank computes on the GLOBAL lattice (no halo exchange) */
int smf_exchange_transport(const smf_ctx *ctx);      /* 0 = single rank, 1 = NCCL send/recv, 2 = fused peer-memory kernel (CUDA IPC over NVLink) */
/* measured FP64 pipe peak of this device (dependent-chain-free DFMA micro-kernel, best of 5 launches, CUDA events): the
 * denominator of the FP64 fractions bench.py prints next to the HBM fractions */
int smf_measure_fp64_peak(smf_ctx *ctx, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif
