// Host driver of the block-GMG Stokes solve on device-resident vectors, plus the operator side of the C ABI.
//
// Mirrors /root/reference/source/simulator/solver/stokes_matrix_free_local_smoothing.cc:758-1313 (solve()):
// Chebyshev smoother set-up and eigenvalue estimation (:774-849), coarse "solver" = level-0 smoother (:855-860),
// V-cycles prec_A / prec_Schur (:893-916), tolerance (:977-1041), cheap GMRES (:1277-1301), and
// /root/reference/include/aspect/simulator/solver/block_stokes_preconditioner.h:220-255, with the deal.II 9.6
// semantics restated in SURVEY.md App. A (PreconditionChebyshev, Multigrid::level_v_step, SolverGMRES).
// Control flow stays on the host (as in the reference); every vector lives in HBM; one stream sync per outer
// iteration (Hessenberg column), two per CG step of the eigenvalue estimation.
#include "smf_internal.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <random>

using namespace smf;

namespace smf {
int launch_diag(smf_ctx *ctx, LevelDev &L);

#define TRY(x) do { int e__ = (x); if (e__) return e__; } while (0)

static double now_s()
{ return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// extreme eigenvalues of the symmetric tridiagonal (diag, off) by Sturm bisection
static void tridiag_extreme(const std::vector<double> &d, const std::vector<double> &o, double &lmin, double &lmax)
{
  const int n = (int)d.size();
  if (n == 0) { lmin = lmax = 1.0; return; }
  if (n == 1) { lmin = lmax = d[0]; return; }
  double lo = d[0], hi = d[0];
  for (int i = 0; i < n; ++i) {
    const double r = (i > 0 ? std::fabs(o[i - 1]) : 0.0) + (i < n - 1 ? std::fabs(o[i]) : 0.0);
    lo = std::min(lo, d[i] - r); hi = std::max(hi, d[i] + r);
  }
  auto count_below = [&](double x) {   // number of eigenvalues < x
    int cnt = 0; double q = d[0] - x;
    if (q < 0) ++cnt;
    for (int i = 1; i < n; ++i) {
      if (q == 0.0) q = 1e-300;
      q = d[i] - x - o[i - 1] * o[i - 1] / q;
      if (q < 0) ++cnt;
    }
    return cnt;
  };
  auto kth = [&](int k) {   // k-th smallest (0-based)
    double a = lo, b = hi;
    for (int it = 0; it < 200; ++it) {
      const double m = 0.5 * (a + b);
      if (m == a || m == b) break;
      if (count_below(m) > k) b = m; else a = m;
    }
    return 0.5 * (a + b);
  };
  lmin = kth(0); lmax = kth(n - 1);
}

constexpr int64_t GRAPH_MAX_CELLS = 70000;   // levels up to this size are replayed from a CUDA graph (the 49,152-cell level of the refinement-3 shell included)

struct MGSide {           // one of the two hierarchies
  bool velocity;
  std::vector<Smoother> *sm;
  std::vector<MGVectors> *v;
  int steps;              // SmootherRelaxation steps (1 for A, 4 for Schur: local_smoothing.cc:775-776, :803-804)
};

static int64_t lsize(smf_ctx *ctx, const MGSide &S, int l) { return S.velocity ? ctx->levels[l].n_u : ctx->levels[l].n_p; }
static const double *ldinv(smf_ctx *ctx, const MGSide &S, int l) { return S.velocity ? ctx->levels[l].dinv_A : ctx->levels[l].dinv_M; }
static int lapply_raw(smf_ctx *ctx, const MGSide &S, int l, double *dst, const double *src, bool dst_is_zero = false)
{ return S.velocity ? launch_vmult_A(ctx, ctx->levels[l], dst, src, dst_is_zero) : launch_vmult_mass(ctx, ctx->levels[l], dst, src, dst_is_zero); }
static EdgeSet &ledge(smf_ctx *ctx, const MGSide &S, int l) { return S.velocity ? ctx->levels[l].edge_u : ctx->levels[l].edge_p; }
// Level operator of local smoothing.  Without refinement edges: the cell loop.  With them (MatrixFreeOperators::Base with
// edge_constrained_indices): the edge entries of src are parked and zeroed around the loop, which then yields
// A_bc (I - E) x on all rows; the edge rows become x_E (the level operator A_l = (I - E) A_bc (I - E) + E of the smoothers) or,
// with_interface_down, A_bc (I - E) x + x_E = (A_l + vmult_interface_down) x, what the V-cycle's residual needs
static int lapply(smf_ctx *ctx, const MGSide &S, int l, double *dst, const double *src, bool dst_is_zero = false, bool with_interface_down = false)
{
  EdgeSet &E = ledge(ctx, S, l);
  if (E.n == 0) return lapply_raw(ctx, S, l, dst, src, dst_is_zero);
  TRY(edge_pre(ctx, E, src));
  TRY(lapply_raw(ctx, S, l, dst, src, dst_is_zero));
  return edge_post(ctx, E, dst, src, with_interface_down);
}
// operator application into the level's t vector: the memset is skipped when the last cheb_update left t zeroed
static int lapply_t(smf_ctx *ctx, const MGSide &S, int l, const double *src)
{
  MGVectors &V = (*S.v)[l];
  const bool z = V.t_zero;
  V.t_zero = false;
  return lapply(ctx, S, l, V.t, src, z);
}

// PreconditionChebyshev::estimate_eigenvalues (App. A.2): Jacobi-preconditioned CG on (global index mod 11) - mean
static int estimate_eigenvalues(smf_ctx *ctx, const MGSide &S, int l)
{
  Smoother &sm = (*S.sm)[l];
  MGVectors &V = (*S.v)[l];
  const int64_t n = lsize(ctx, S, l);
  const double *dinv = ldinv(ctx, S, l);
  // mean of i mod 11 over 0..n-1
  const int64_t full = n / 11, rem = n % 11;
  const double mean = (double)(full * 55 + rem * (rem - 1) / 2) / (double)n;
  double *r = V.b, *z = V.t, *p = V.x, *v = V.xo;
  const ExchangeDev &E = S.velocity ? ctx->levels[l].ex_u : ctx->levels[l].ex_p;
  DotWeights dw(ctx, E.weights);
  if (E.gid) TRY(start_vector_mod11(ctx, E, r, n));     // partition-independent global DoF numbers
  else TRY(vec_mod11(ctx, r, 0, mean, n));
  double res2;
  TRY(vec_norm2(ctx, r, n, &res2));
  std::vector<double> diag, off;
  int it = 0;
  double rz_old = 0, alpha_old = 0, beta = 0;
  const double tol = 1e-7;
  if (std::sqrt(res2) > tol && sm.eig_cg_n_iterations > 0)
    while (true) {
      ++it;
      TRY(vec_pointwise(ctx, z, 1.0, dinv, r, n));
      double rz;
      TRY(vec_dot(ctx, r, z, n, &rz));
      if (it > 1) { beta = rz / rz_old; TRY(vec_sadd(ctx, p, beta, 1.0, z, n)); }
      else TRY(vec_copy(ctx, p, z, n));
      TRY(lapply(ctx, S, l, v, p));
      double pv;
      TRY(vec_dot(ctx, p, v, n, &pv));
      const double alpha = rz / pv;
      TRY(vec_axpy(ctx, r, -alpha, v, n));
      TRY(vec_norm2(ctx, r, n, &res2));
      if (it > 1) { diag.push_back(1.0 / alpha + beta / alpha_old); off.push_back(std::sqrt(beta) / alpha_old); }
      else diag.push_back(1.0 / alpha);
      rz_old = rz; alpha_old = alpha;
      if (it >= sm.eig_cg_n_iterations || std::sqrt(res2) <= tol) break;
    }
  double lmin = 1.0, lmax = 1.0;
  tridiag_extreme(diag, off, lmin, lmax);
  sm.cg_its = it;
  sm.lambda_min = lmin;
  sm.lambda_max = 1.2 * lmax;
  const double alpha = sm.smoothing_range > 1.0 ? sm.lambda_max / sm.smoothing_range : std::min(0.9 * sm.lambda_max, lmin);
  sm.delta = 0.5 * (sm.lambda_max - alpha);
  sm.theta = 0.5 * (sm.lambda_max + alpha);
  V.t_zero = false;         // (t served as the CG's z vector)
  return SMF_OK;
}

// Chebyshev iterations after the first update; x holds the current iterate, xo the previous one (if have_old)
static int cheb_tail(smf_ctx *ctx, const MGSide &S, int l, double *&x, double *&xo, const double *b, double *t, bool have_old)
{
  const Smoother &sm = (*S.sm)[l];
  const int64_t n = lsize(ctx, S, l);
  if (sm.degree < 2 || std::fabs(sm.delta) < 1e-40) return SMF_OK;
  double rhok = sm.delta / sm.theta;
  const double sigma = sm.theta / sm.delta;
  MGVectors &V = (*S.v)[l];
  for (int k = 0; k < sm.degree - 1; ++k) {
    TRY(lapply_t(ctx, S, l, x));
    const double rhokp = 1.0 / (2.0 * sigma - rhok);
    const double f1 = rhokp * rhok, f2 = 2.0 * rhokp / sm.delta;
    rhok = rhokp;
    TRY(cheb_update(ctx, x, xo, ldinv(ctx, S, l), b, t, f1, f2, have_old, n));
    V.t_zero = true;
    std::swap(x, xo);
    have_old = true;
  }
  return SMF_OK;
}
// PreconditionChebyshev::vmult: zero initial guess, degree-1 operator applications
static int cheb_vmult(smf_ctx *ctx, const MGSide &S, int l, double *&x, double *&xo, const double *b, double *t)
{
  const Smoother &sm = (*S.sm)[l];
  TRY(cheb_first(ctx, x, ldinv(ctx, S, l), b, 1.0 / sm.theta, lsize(ctx, S, l)));
  return cheb_tail(ctx, S, l, x, xo, b, t, false);
}
// PreconditionChebyshev::step: non-zero initial guess, degree operator applications
static int cheb_step(smf_ctx *ctx, const MGSide &S, int l, double *&x, double *&xo, const double *b, double *t)
{
  const Smoother &sm = (*S.sm)[l];
  const int64_t n = lsize(ctx, S, l);
  TRY(lapply_t(ctx, S, l, x));
  // x_new = x + (1/theta) D^-1 (b - A x): the update kernel with f1 = 0, written to xo, then swapped
  TRY(cheb_update(ctx, x, xo, ldinv(ctx, S, l), b, t, 0.0, 1.0 / sm.theta, false, n));
  (*S.v)[l].t_zero = true;
  std::swap(x, xo);
  return cheb_tail(ctx, S, l, x, xo, b, t, true);
}

static int level_v_step_maybe_graph(smf_ctx *ctx, const MGSide &S, int l);

// Multigrid::level_v_step (App. A.3); defect of level l is V[l].b, the result ends in V[l].res
static int level_v_step(smf_ctx *ctx, const MGSide &S, int l)
{
  MGVectors &V = (*S.v)[l];
  const int which = S.velocity ? 0 : 1;
  VecBucket vb(ctx, B_SMOOTH_VEC);
  double *x = V.x, *xo = V.xo;
  // pre-smoothing: SmootherRelaxation::apply = vmult + (steps-1) x step; on level 0 this is the coarse solver
  TRY(cheb_vmult(ctx, S, l, x, xo, V.b, V.t));
  for (int s = 1; s < S.steps; ++s) TRY(cheb_step(ctx, S, l, x, xo, V.b, V.t));
  if (l > 0) {
    const int64_t n = lsize(ctx, S, l);
    MGVectors &C = (*S.v)[l - 1];
    TRY(lapply_t(ctx, S, l, x));
    TRY(vec_set(ctx, C.b, 0.0, lsize(ctx, S, l - 1)));
    TRY(launch_restrict_add_residual(ctx, which, l, C.b, V.b, V.t, n));      // coarse defect = R (defect - A x)
    if (!S.velocity && l - 1 == ctx->tail_level) TRY(launch_schur_tail(ctx, l - 1));      // levels 0..l-1 in one launch
    else if (S.velocity && l - 1 == ctx->replica_top) { TRY(replica_cycle(ctx, C.b, C.x)); C.res = C.x; }      // several ranks: levels 0..l-1 on the global mesh, no exchanges
    else TRY(level_v_step_maybe_graph(ctx, S, l - 1));
    TRY(launch_prolongate(ctx, which, l, x, C.res, true));                   // x += P x_coarse
    for (int s = 0; s < S.steps; ++s) TRY(cheb_step(ctx, S, l, x, xo, V.b, V.t));   // post-smoothing
  }
  V.res = x;
  return SMF_OK;
}

// Multigrid::level_v_step on an adaptively refined mesh (local smoothing, SURVEY App. A.3 with edge matrices):
// the defect of every level was filled by copy_to_mg, so the restriction ADDS; the residual carries the edge_out term
// (vmult_interface_down, fused into the level's cell loop), and before post-smoothing the defect loses
// vmult_interface_up(x) = (I - E) A_bc (E x) (one more cell loop on the edge part of the iterate).
static int level_v_step_adaptive(smf_ctx *ctx, const MGSide &S, int l)
{
  MGVectors &V = (*S.v)[l];
  const int which = S.velocity ? 0 : 1;
  VecBucket vb(ctx, B_SMOOTH_VEC);
  const int64_t n = lsize(ctx, S, l);
  double *x = V.x, *xo = V.xo;
  V.t_zero = false;
  TRY(cheb_vmult(ctx, S, l, x, xo, V.b, V.t));
  for (int s = 1; s < S.steps; ++s) TRY(cheb_step(ctx, S, l, x, xo, V.b, V.t));
  if (l > 0) {
    MGVectors &C = (*S.v)[l - 1];
    V.t_zero = false;
    TRY(lapply(ctx, S, l, V.t, x, false, true));                              // t = (A_l + interface_down) x
    TRY(vec_residual(ctx, V.t, V.b, n));                                     // t = defect - t
    TRY(launch_restrict_add(ctx, which, l, C.b, V.t));                       // onto the copy_to_mg part of the coarse defect
    TRY(level_v_step_adaptive(ctx, S, l - 1));
    TRY(launch_prolongate(ctx, which, l, x, C.res, true));
    EdgeSet &E = ledge(ctx, S, l);
    if (E.n) {
      if (!V.e) { SMF_CUDA(cudaMalloc(&V.e, sizeof(double) * n)); ctx->bytes += sizeof(double) * n; }
      TRY(edge_extract(ctx, E, V.e, x, n));                                  // e = E x
      TRY(lapply_raw(ctx, S, l, V.t, V.e, false));                           // t = A_bc e
      TRY(edge_subtract_off(ctx, E, V.b, V.t, n));                           // defect -= (I - E) t
    }
    V.t_zero = false;
    for (int s = 0; s < S.steps; ++s) TRY(cheb_step(ctx, S, l, x, xo, V.b, V.t));
  }
  V.res = x;
  return SMF_OK;
}

// PreconditionMG::vmult on an adaptive mesh: copy_to_mg fills the defects of ALL levels from the active vector, copy_from_mg
// collects the solutions of all levels (MGTransferMatrixFree copy indices, smf_set_level_copy_indices)
static int vcycle_adaptive(smf_ctx *ctx, const MGSide &S, double *dst, const double *src)
{
  const int nl = ctx->cfg.n_levels;
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  for (int l = 0; l < nl; ++l) {
    MGVectors &V = (*S.v)[l];
    TRY(vec_set(ctx, V.b, 0.0, lsize(ctx, S, l)));
    TRY(copy_pairs(ctx, S.velocity ? ctx->levels[l].copy_u : ctx->levels[l].copy_p, V.b, const_cast<double *>(src), true));
  }
  TRY(level_v_step_adaptive(ctx, S, nl - 1));
  TRY(vec_set(ctx, dst, 0.0, S.velocity ? A.n_u : A.n_p));
  for (int l = 0; l < nl; ++l)
    TRY(copy_pairs(ctx, S.velocity ? ctx->levels[l].copy_u : ctx->levels[l].copy_p, (*S.v)[l].res, dst, false));
  return SMF_OK;
}

// The adaptive cycle issues ~40 launches per level plus a dozen list kernels (edges, copy indices), nearly all of them tiny: on
// BASELINE cfg 4 (38 k cells) a solve was pure launch latency.  All buffers and coefficients are fixed between smoother set-ups,
// so the whole cycle is captured once into a CUDA graph between two staging copies (the caller's vectors change from call to
// call); the first call after a set-up runs eagerly (it allocates the edge scratch vectors).
static int vcycle_adaptive_maybe_graph(smf_ctx *ctx, const MGSide &S, double *dst, const double *src)
{
  if (!ctx->use_graphs || ctx->prof_mode != 0 || ctx->world > 1) return vcycle_adaptive(ctx, S, dst, src);
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  const int64_t n = S.velocity ? A.n_u : A.n_p;
  double *&in = S.velocity ? ctx->ad_in_u : ctx->ad_in_p, *&out = S.velocity ? ctx->ad_out_u : ctx->ad_out_p;
  bool &warm = S.velocity ? ctx->ad_warm_u : ctx->ad_warm_p;
  void *&exec = S.velocity ? ctx->graph_A : ctx->graph_S;
  bool &stale = S.velocity ? ctx->graph_A_stale : ctx->graph_S_stale;
  int64_t &n_launch = S.velocity ? ctx->graph_A_launches : ctx->graph_S_launches;
  if (!in) {
    SMF_CUDA(cudaMalloc(&in, sizeof(double) * n)); SMF_CUDA(cudaMalloc(&out, sizeof(double) * n));
    ctx->bytes += 2 * sizeof(double) * n;
  }
  if (!warm) { warm = true; stale = true; return vcycle_adaptive(ctx, S, dst, src); }
  VecBucket vb(ctx, B_VEC);
  TRY(vec_copy(ctx, in, src, n));
  if (stale) {
    cudaGraph_t graph = nullptr;
    SMF_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    const int64_t l0 = ctx->launches;
    const int rc = vcycle_adaptive(ctx, S, out, in);
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return ctx->fail(SMF_ERR_CUDA, std::string("cudaStreamEndCapture (adaptive V-cycle): ") + cudaGetErrorString(ce));
    n_launch = ctx->launches - l0;
    ctx->launches = l0;
    bool updated = false;
    if (exec) {
      cudaGraphExecUpdateResultInfo info;
      updated = cudaGraphExecUpdate((cudaGraphExec_t)exec, graph, &info) == cudaSuccess;
      if (!updated) { cudaGetLastError(); cudaGraphExecDestroy((cudaGraphExec_t)exec); exec = nullptr; }
    }
    if (!updated) {
      cudaGraphExec_t ge = nullptr;
      const cudaError_t ie = cudaGraphInstantiate(&ge, graph, 0);
      if (ie != cudaSuccess) { cudaGraphDestroy(graph); return ctx->fail(SMF_ERR_CUDA, std::string("cudaGraphInstantiate (adaptive V-cycle): ") + cudaGetErrorString(ie)); }
      exec = ge;
    }
    cudaGraphDestroy(graph);
    stale = false;
  }
  SMF_CUDA(cudaGraphLaunch((cudaGraphExec_t)exec, ctx->stream));
  ctx->launches += n_launch;
  return vec_copy(ctx, dst, out, n);
}

// Levels <= ctx->graph_level are small: their part of the V-cycle is launch-latency bound, so it is captured once per
// smoother set-up into a CUDA graph (all buffers and coefficients are fixed between set-ups) and replayed.
static int level_v_step_maybe_graph(smf_ctx *ctx, const MGSide &S, int l)
{
  const bool profiling_inside = ctx->prof_mode == 2 || (ctx->prof_mode == 1 && l == ctx->cfg.n_levels - 1);
  // the coarse tail of the Schur V-cycle: levels 0..l in one launch (coarse_tail.cu); inside a graph it is one node
  if (!ctx->use_graphs || l != ctx->graph_level || profiling_inside) return level_v_step(ctx, S, l);
  if (S.velocity && ctx->replica_top >= 0) return level_v_step(ctx, S, l);      // (the replicated part brings its own graph; an all-reduce sits above it)
  void *&exec = S.velocity ? ctx->graph_A : ctx->graph_S;
  bool &stale = S.velocity ? ctx->graph_A_stale : ctx->graph_S_stale;
  int64_t &n_launch = S.velocity ? ctx->graph_A_launches : ctx->graph_S_launches;
  int64_t &n_exch = S.velocity ? ctx->graph_A_exchanges : ctx->graph_S_exchanges;
  if (stale) {
    cudaGraph_t graph = nullptr;
    SMF_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    const int64_t l0 = ctx->launches, x0 = ctx->exchanges;
    const int rc = level_v_step(ctx, S, l);
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return ctx->fail(SMF_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    n_launch = ctx->launches - l0;
    ctx->launches = l0;
    n_exch = ctx->exchanges - x0;
    ctx->exchanges = x0;
    bool updated = false;
    if (exec) {
      cudaGraphExecUpdateResultInfo info;
      updated = cudaGraphExecUpdate((cudaGraphExec_t)exec, graph, &info) == cudaSuccess;
      if (!updated) { cudaGetLastError(); cudaGraphExecDestroy((cudaGraphExec_t)exec); exec = nullptr; }
    }
    if (!updated) {
      cudaGraphExec_t ge = nullptr;
      const cudaError_t ie = cudaGraphInstantiate(&ge, graph, 0);
      if (ie != cudaSuccess) { cudaGraphDestroy(graph); return ctx->fail(SMF_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
      exec = ge;
    }
    cudaGraphDestroy(graph);
    stale = false;
  }
  SMF_CUDA(cudaGraphLaunch((cudaGraphExec_t)exec, ctx->stream));
  ctx->launches += n_launch;
  ctx->exchanges += n_exch;
  return SMF_OK;
}

static int vcycle(smf_ctx *ctx, bool velocity, double *dst, const double *src)
{
  if (!ctx->smoothers_ready) return ctx->fail(SMF_ERR_INVALID, "call smf_setup_smoothers before applying a V-cycle");
  MGSide S{velocity, velocity ? &ctx->smoother_A : &ctx->smoother_S, velocity ? &ctx->mgv_A : &ctx->mgv_S, velocity ? 1 : 4};
  if (ctx->adaptive) return vcycle_adaptive_maybe_graph(ctx, S, dst, src);
  if (!velocity && ctx->lat_state == 1 && ctx->use_lattice) return lattice_vcycle(ctx, dst, src);      // box levels: schur_lattice.cu
  const int L = ctx->cfg.n_levels - 1;
  const int64_t n = lsize(ctx, S, L);
  const uint32_t *map = velocity ? ctx->copy_u : ctx->copy_p;
  MGVectors &V = (*S.v)[L];
  // Identity copy_to_mg / copy_from_mg and a finest level outside the CUDA graph: no copies at all.  The defect is
  // only read, so src serves as V.b; the smoothers alternate between x and xo, and the number of updates on the
  // level is known, so the buffer the result will end in is replaced by dst for the duration of the cycle.
  const Smoother &sm = (*S.sm)[L];
  const bool replayed = ctx->use_graphs && L == ctx->graph_level && ctx->prof_mode != 2 && !(ctx->prof_mode == 1 && L == ctx->cfg.n_levels - 1) &&
                        !(velocity && ctx->replica_top >= 0);
  if (!map && dst != src && !replayed && L > 0 && sm.degree >= 2 && std::fabs(sm.delta) >= 1e-40) {
    const int swaps = (sm.degree - 1) + (S.steps - 1) * sm.degree + S.steps * sm.degree;
    double *const b0 = V.b, *const x0 = V.x, *const xo0 = V.xo;
    V.b = const_cast<double *>(src);
    if (swaps & 1) V.xo = dst; else V.x = dst;
    const int rc = level_v_step(ctx, S, L);
    double *res = V.res;
    V.b = b0; V.x = x0; V.xo = xo0;
    if (rc) return rc;
    if (res != dst) TRY(vec_copy(ctx, dst, res, n));
    return SMF_OK;
  }
  // copy_to_mg
  if (map) TRY(vec_gather(ctx, V.b, src, map, n, true));
  else TRY(vec_copy(ctx, V.b, src, n));
  TRY(level_v_step_maybe_graph(ctx, S, L));
  // copy_from_mg
  if (map) TRY(vec_gather(ctx, dst, V.res, map, n, false));
  else TRY(vec_copy(ctx, dst, V.res, n));
  return SMF_OK;
}

static int ensure_tmp(smf_ctx *ctx)
{
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  if (!ctx->tmp_u) {
    SMF_CUDA(cudaMalloc(&ctx->tmp_u, sizeof(double) * (A.n_u + A.n_p)));
    ctx->tmp_p = ctx->tmp_u + A.n_u;
    SMF_CUDA(cudaMalloc(&ctx->tmp_u2, sizeof(double) * (A.n_u + A.n_p)));
    ctx->tmp_p2 = ctx->tmp_u2 + A.n_u;
    ctx->bytes += 2 * sizeof(double) * (A.n_u + A.n_p);
  }
  return SMF_OK;
}

static int vcycle(smf_ctx *ctx, bool velocity, double *dst, const double *src);

// deal.II SolverCG from a zero start vector, preconditioned by the V-cycle: the inner solves of the expensive phase
// (InverseVelocityBlock::vmult with do_solve_A, block_stokes_preconditioner.h:108-135; SchurApproximation::vmult with
// do_solve_Schur_complement, :312-345).  Iterations = SolverControl::last_step().
static int inner_cg(smf_ctx *ctx, bool velocity, double *x, const double *b, double tol, int max_its, int *its_out)
{
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  const int64_t n = velocity ? A.n_u : A.n_p;
  if (!ctx->cg_buf) {
    SMF_CUDA(cudaMalloc(&ctx->cg_buf, sizeof(double) * 3 * A.n_u));
    ctx->bytes += sizeof(double) * 3 * A.n_u;
  }
  double *g = ctx->cg_buf, *d = g + A.n_u, *h = d + A.n_u;
  DotWeights dw(ctx, velocity ? A.ex_u.weights : A.ex_p.weights);
  *its_out = 0;
  TRY(vec_set(ctx, x, 0.0, n));
  TRY(vec_copy(ctx, g, b, n));
  TRY(vec_scale(ctx, g, -1.0, n));                    // g = A x - b
  double r2;
  TRY(vec_norm2(ctx, g, n, &r2));
  if (std::sqrt(r2) <= tol) return SMF_OK;
  double gh = 0.0;
  for (int it = 1;; ++it) {
    TRY(vcycle(ctx, velocity, h, g));
    if (it > 1) {
      double beta = gh;
      TRY(vec_dot(ctx, g, h, n, &gh));
      beta = gh / beta;
      TRY(vec_sadd(ctx, d, beta, -1.0, h, n));        // d = -h + beta d
    } else {
      TRY(vec_copy(ctx, d, h, n));
      TRY(vec_scale(ctx, d, -1.0, n));
      TRY(vec_dot(ctx, g, h, n, &gh));
    }
    TRY(velocity ? launch_vmult_A(ctx, A, h, d) : launch_vmult_mass(ctx, A, h, d));
    double dAd;
    TRY(vec_dot(ctx, d, h, n, &dAd));
    const double alpha = gh / dAd;
    TRY(vec_axpy(ctx, x, alpha, d, n));
    TRY(vec_axpy(ctx, g, alpha, h, n));
    TRY(vec_norm2(ctx, g, n, &r2));
    *its_out = it;
    if (std::sqrt(r2) <= tol) return SMF_OK;
    if (it >= max_its) return SMF_NO_CONVERGENCE;
  }
}

// deal.II SolverBicgstab(AdditionalData(exact_residual = false)) from a zero start vector, preconditioned by the A-block
// V-cycle: InverseVelocityBlock::vmult when the A block is flagged non-symmetric (block_stokes_preconditioner.h:119-132: the
// free-surface stabilisation, or `Force nonsymmetric A block solver`).  Restated from deal.II 9.6 (solver_bicgstab.h, iterate()):
// the early-success check after the first half step, the residual norm updated by add_and_dot, breakdown -> restart from the
// current iterate.  Iterations = SolverControl::last_step().
static int inner_bicgstab(smf_ctx *ctx, double *x, const double *b, double tol, int max_its, int *its_out)
{
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  const int64_t n = A.n_u;
  if (!ctx->bicg_buf) {
    SMF_CUDA(cudaMalloc(&ctx->bicg_buf, sizeof(double) * 7 * n));
    ctx->bytes += sizeof(double) * 7 * n;
  }
  double *r = ctx->bicg_buf, *rbar = r + n, *p = rbar + n, *y = p + n, *z = y + n, *t = z + n, *v = t + n;
  DotWeights dw(ctx, A.ex_u.weights);
  const double breakdown = 1e-10;                   // SolverBicgstab::AdditionalData default
  *its_out = 0;
  TRY(vec_set(ctx, x, 0.0, n));
  int step = 0;
  while (true) {                                    // solve(): iterate() again after a breakdown
    TRY(launch_vmult_A(ctx, A, r, x));
    TRY(vec_sadd(ctx, r, -1.0, 1.0, b, n));         // r = b - A x
    TRY(vec_copy(ctx, rbar, r, n));
    double r2;
    TRY(vec_norm2(ctx, r, n, &r2));
    double res = std::sqrt(r2);
    *its_out = step;
    if (res <= tol) return SMF_OK;
    if (step >= max_its) return SMF_NO_CONVERGENCE;
    double alpha = 1.0, rho = 1.0, omega = 1.0;
    const int first = step + 1;
    bool broke = false;
    while (true) {
      ++step;
      double rhobar;
      if (step == first) rhobar = res * res; else TRY(vec_dot(ctx, r, rbar, n, &rhobar));
      if (std::fabs(rhobar) < breakdown) { broke = true; break; }
      const double beta = rhobar * alpha / (rho * omega);
      rho = rhobar;
      if (step == first) TRY(vec_copy(ctx, p, r, n));
      else { TRY(vec_sadd(ctx, p, beta, 1.0, r, n)); TRY(vec_axpy(ctx, p, -beta * omega, v, n)); }
      TRY(vcycle(ctx, true, y, p));
      TRY(launch_vmult_A(ctx, A, v, y));
      double rbar_v;
      TRY(vec_dot(ctx, rbar, v, n, &rbar_v));
      if (std::fabs(rbar_v) < breakdown) { broke = true; break; }
      alpha = rho / rbar_v;
      TRY(vec_axpy(ctx, r, -alpha, v, n));
      TRY(vec_norm2(ctx, r, n, &r2));
      res = std::sqrt(r2);
      *its_out = step;
      if (res <= tol) { TRY(vec_axpy(ctx, x, alpha, y, n)); return SMF_OK; }      // early success
      TRY(vcycle(ctx, true, z, r));
      TRY(launch_vmult_A(ctx, A, t, z));
      double t_r, t_t;
      TRY(vec_dot(ctx, t, r, n, &t_r));
      TRY(vec_norm2(ctx, t, n, &t_t));
      if (t_t < breakdown) { broke = true; break; }
      omega = t_r / t_t;
      TRY(vec_axpy(ctx, x, alpha, y, n));
      TRY(vec_axpy(ctx, x, omega, z, n));
      TRY(vec_axpy(ctx, r, -omega, t, n));
      TRY(vec_norm2(ctx, r, n, &r2));
      res = std::sqrt(r2);
      if (res <= tol) return SMF_OK;
      if (step >= max_its) return SMF_NO_CONVERGENCE;
    }
    if (!broke) break;
  }
  return SMF_OK;
}

// BlockSchurPreconditioner::vmult (block_stokes_preconditioner.h:220-255): cheap variant = one V-cycle per block;
// expensive variant = CG solves with the active-level mass matrix / A block, preconditioned by the V-cycles
static int precondition(smf_ctx *ctx, double *dst_u, double *dst_p, const double *src_u, const double *src_p,
                        const smf_solve_params *expensive = nullptr)
{
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  TRY(ensure_tmp(ctx));
  if (!expensive) {
    TRY(vcycle(ctx, false, dst_p, src_p));            // S^-1 (SchurApproximation::vmult :346-351)
    ctx->n_S_prec++;
  } else {
    double r2;
    { DotWeights dw(ctx, A.ex_p.weights); TRY(vec_norm2(ctx, src_p, A.n_p, &r2)); }
    const double nrm = std::sqrt(r2);
    if (nrm > 1e-50) {
      int its = 0;
      const int rc = inner_cg(ctx, false, dst_p, src_p, nrm * expensive->linear_solver_S_block_tolerance, 100, &its);
      ctx->n_S_exp += its;
      if (rc == SMF_NO_CONVERGENCE) return ctx->fail(SMF_NO_CONVERGENCE, "iterative (bottom right) solver did not converge in BlockSchurPreconditioner::vmult");
      TRY(rc);
    } else TRY(vec_set(ctx, dst_p, 0.0, A.n_p));
  }
  TRY(vec_scale(ctx, dst_p, -1.0, A.n_p));
  TRY(launch_vmult_BT(ctx, A, ctx->tmp_u, dst_p));  // tmp = B^T z_p
  TRY(vec_sadd(ctx, ctx->tmp_u, -1.0, 1.0, src_u, A.n_u));   // tmp = src_u - tmp
  if (!expensive) {
    TRY(vcycle(ctx, true, dst_u, ctx->tmp_u));        // A^-1 (InverseVelocityBlock::vmult :148-152)
    ctx->n_A_prec++;
  } else {
    double r2;
    { DotWeights dw(ctx, A.ex_u.weights); TRY(vec_norm2(ctx, ctx->tmp_u, A.n_u, &r2)); }
    int its = 0;
    // sim.stokes_A_block_is_symmetric() (local_smoothing.cc:1071): false with the free-surface stabilisation or when forced
    const bool nonsymmetric = ctx->force_nonsymmetric_A || A.fs_faces.n > 0;
    const int rc = nonsymmetric ? inner_bicgstab(ctx, dst_u, ctx->tmp_u, std::sqrt(r2) * expensive->linear_solver_A_block_tolerance, 10000, &its)
                                : inner_cg(ctx, true, dst_u, ctx->tmp_u, std::sqrt(r2) * expensive->linear_solver_A_block_tolerance, 10000, &its);
    ctx->n_A_exp += its;
    if (rc == SMF_NO_CONVERGENCE) return ctx->fail(SMF_NO_CONVERGENCE, "iterative (top left) solver did not converge in BlockSchurPreconditioner::vmult");
    TRY(rc);
  }
  return SMF_OK;
}

static int ensure_krylov(smf_ctx *ctx, int n_vectors)
{
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  const int64_t n = A.n_u + A.n_p;
  if (ctx->krylov_n != n) {
    for (double *p : ctx->krylov) cudaFree(p);
    ctx->krylov.clear();
    ctx->krylov_n = n;
  }
  while ((int)ctx->krylov.size() < n_vectors) {
    double *p = nullptr;
    SMF_CUDA(cudaMalloc(&p, sizeof(double) * n));
    ctx->krylov.push_back(p);
    ctx->bytes += sizeof(double) * n;
  }
  return SMF_OK;
}

// SolverGMRES with right preconditioning and delayed classical Gram-Schmidt (App. A.5)
static int gmres(smf_ctx *ctx, const smf_solve_params *prm, double *x_u, double *x_p, const double *b_u, const double *b_p,
                 double tol, smf_solve_result *res)
{
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  const int64_t nu = A.n_u, np = A.n_p, n = nu + np;
  if (prm->gmres_restart_length > SMF_MAX_RESTART) return ctx->fail(SMF_ERR_UNSUPPORTED, "GMRES solver restart length above 200");
  const int restart = std::max(1, prm->gmres_restart_length);
  const int max_its = prm->n_cheap_stokes_solver_steps;
  TRY(ensure_krylov(ctx, restart + 3));
  DotWeights dw(ctx, ctx->block_weights);
  std::vector<double *> &K = ctx->krylov;
  double *p = K[restart + 1], *xb = K[restart + 2];   // p: scratch block vector; xb: contiguous copy of the iterate
  TRY(vec_copy(ctx, xb, x_u, nu));
  TRY(vec_copy(ctx, xb + nu, x_p, np));
  const double eps = 2.220446049250313e-16;
  std::vector<double> H((size_t)(restart + 1) * restart, 0.0), gamma(restart + 1), ci(restart), si(restart), h(restart + 3), hh(restart + 3);
  int its = 0;
  // deal.II 9.6's default (delayed classical Gram-Schmidt, DCGS2) orthogonalises twice per iteration; its
  // numerics are those of CGS2, restated here as "always re-orthogonalise".  Single-pass CGS loses the
  // orthogonality of the basis on badly scaled problems (SolCx, pressure scaling 1e5) and stagnates.
  // As in DCGS2 the second projection costs no second sweep over the basis: it is formed from the Gram matrix
  // of the basis (vec_multi_dot_gram), so one iteration streams the basis twice (dots, update).
  bool re_orth = true;
  const bool single_pass = ctx->ortho_single_pass != 0;
  int state = 0;   // 0 iterate, 1 success, 2 failure
  double rho = 0;
  bool first = true;
  while (state == 0) {
    // p = b - A x
    TRY(launch_vmult_stokes(ctx, A, p, p + nu, xb, xb + nu));
    TRY(vec_sadd(ctx, p, -1.0, 1.0, b_u, nu));
    TRY(vec_sadd(ctx, p + nu, -1.0, 1.0, b_p, np));
    double r2;
    TRY(vec_norm2(ctx, p, n, &r2));
    rho = std::sqrt(r2);
    if (first) { res->initial_residual = rho; first = false; }
    if (rho <= tol) { state = 1; break; }
    if (its >= max_its) { state = 2; break; }
    std::fill(gamma.begin(), gamma.end(), 0.0);
    gamma[0] = rho;
    TRY(vec_copy(ctx, K[0], p, n));
    TRY(vec_scale(ctx, K[0], 1.0 / rho, n));
    int dim = 0;
    for (int j = 0; j < restart && state == 0; ++j) {
      ++its;
      double *vv = K[j + 1];
      TRY(precondition(ctx, p, p + nu, K[j], K[j] + nu));
      TRY(launch_vmult_stokes(ctx, A, vv, vv + nu, p, p + nu));
      dim = j + 1;
      const bool consider = !re_orth && (its % 5 == 0);
      VecBucket vb(ctx, B_ORTHO);
      if (single_pass) TRY(vec_multi_dot_gram(ctx, K.data(), dim, vv, n));
      else TRY(vec_multi_dot(ctx, K.data(), dim, vv, n, nullptr, false));
      TRY(vec_multi_axpy_from_dots(ctx, vv, dim, n, h.data()));   // h[0..dim-1], |vv_old|^2, |vv_new|^2
      const double norm_start = std::sqrt(h[dim]);
      double nrm = std::sqrt(h[dim + 1]);
      if (consider && !(nrm > 10.0 * norm_start * std::sqrt(eps))) re_orth = true;
      if (re_orth && !single_pass) {
        TRY(vec_multi_dot(ctx, K.data(), dim, vv, n, nullptr, false));
        TRY(vec_multi_axpy_from_dots(ctx, vv, dim, n, hh.data()));
        for (int i = 0; i < dim; ++i) h[i] += hh[i];
        nrm = std::sqrt(hh[dim + 1]);
      }
      h[dim] = nrm;
      if (nrm != 0.0) TRY(vec_scale(ctx, vv, 1.0 / nrm, n));
      // Givens rotations
      for (int i = 0; i < j; ++i) {
        const double t = h[i];
        h[i] = ci[i] * t + si[i] * h[i + 1];
        h[i + 1] = -si[i] * t + ci[i] * h[i + 1];
      }
      const double s = std::sqrt(h[j] * h[j] + h[j + 1] * h[j + 1]);
      ci[j] = h[j] / s; si[j] = h[j + 1] / s;
      h[j] = s;
      gamma[j + 1] = -si[j] * gamma[j];
      gamma[j] *= ci[j];
      for (int i = 0; i < dim; ++i) H[(size_t)i * restart + j] = h[i];
      rho = std::fabs(gamma[dim]);
      if (rho <= tol) state = 1;
      else if (its >= max_its) state = 2;
    }
    // x += P^-1 (V y), H y = gamma
    std::vector<double> y(dim);
    for (int i = dim - 1; i >= 0; --i) {
      double s = gamma[i];
      for (int k = i + 1; k < dim; ++k) s -= H[(size_t)i * restart + k] * y[k];
      y[i] = s / H[(size_t)i * restart + i];
    }
    TRY(vec_set(ctx, p, 0.0, n));
    TRY(vec_multi_add(ctx, p, K.data(), y.data(), dim, n));
    double *z = K[dim];       // v_dim is not part of the combination any more: reuse its storage
    TRY(precondition(ctx, z, z + nu, p, p + nu));
    TRY(vec_axpy(ctx, xb, 1.0, z, n));
  }
  TRY(vec_copy(ctx, x_u, xb, nu));
  TRY(vec_copy(ctx, x_p, xb + nu, np));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  res->iterations = its;
  res->final_residual = rho;
  return state == 1 ? SMF_OK : SMF_NO_CONVERGENCE;
}


// SolverFGMRES(AdditionalData(restart)) with the expensive preconditioner (local_smoothing.cc:1333-1350).  The basis
// vectors v_j and the preconditioned vectors z_j are allocated as the iteration proceeds (K[2 + 2j], K[3 + 2j]).
static int fgmres(smf_ctx *ctx, const smf_solve_params *prm, double *x_u, double *x_p, const double *b_u, const double *b_p,
                  double tol, smf_solve_result *res)
{
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  const int64_t nu = A.n_u, np = A.n_p, n = nu + np;
  if (prm->gmres_restart_length > SMF_MAX_RESTART) return ctx->fail(SMF_ERR_UNSUPPORTED, "GMRES solver restart length above 200");
  const int restart = std::max(1, prm->gmres_restart_length);
  const int max_its = prm->n_expensive_stokes_solver_steps;
  TRY(ensure_krylov(ctx, 4));
  DotWeights dw(ctx, ctx->block_weights);
  std::vector<double *> &K = ctx->krylov;
  double *xb = K[0], *p = K[1];
  TRY(vec_copy(ctx, xb, x_u, nu));
  TRY(vec_copy(ctx, xb + nu, x_p, np));
  std::vector<double> H((size_t)(restart + 1) * restart, 0.0), gamma(restart + 1), ci(restart), si(restart), h(restart + 3), hh(restart + 3);
  std::vector<double *> V, Z;
  int its = 0, state = 0;
  double rho = 0;
  while (state == 0) {
    TRY(launch_vmult_stokes(ctx, A, p, p + nu, xb, xb + nu));
    TRY(vec_sadd(ctx, p, -1.0, 1.0, b_u, nu));
    TRY(vec_sadd(ctx, p + nu, -1.0, 1.0, b_p, np));
    double r2;
    TRY(vec_norm2(ctx, p, n, &r2));
    rho = std::sqrt(r2);
    if (its == 0 && res->iterations == 0) res->initial_residual = rho;
    if (rho <= tol) { state = 1; break; }
    if (its >= max_its) { state = 2; break; }
    std::fill(gamma.begin(), gamma.end(), 0.0);
    gamma[0] = rho;
    V.assign(1, K[2]); Z.clear();
    TRY(vec_copy(ctx, V[0], p, n));
    TRY(vec_scale(ctx, V[0], 1.0 / rho, n));
    int dim = 0;
    for (int j = 0; j < restart && state == 0; ++j) {
      ++its;
      TRY(ensure_krylov(ctx, 2 + 2 * (j + 1) + 1));
      Z.push_back(K[3 + 2 * j]);
      V.push_back(K[2 + 2 * (j + 1)]);
      double *vv = V[j + 1];
      TRY(precondition(ctx, Z[j], Z[j] + nu, V[j], V[j] + nu, prm));
      TRY(launch_vmult_stokes(ctx, A, vv, vv + nu, Z[j], Z[j] + nu));
      dim = j + 1;
      {
        VecBucket vb(ctx, B_ORTHO);
        TRY(vec_multi_dot(ctx, V.data(), dim, vv, n, nullptr, false));
        TRY(vec_multi_axpy_from_dots(ctx, vv, dim, n, h.data()));
        TRY(vec_multi_dot(ctx, V.data(), dim, vv, n, nullptr, false));
        TRY(vec_multi_axpy_from_dots(ctx, vv, dim, n, hh.data()));
      }
      for (int i = 0; i < dim; ++i) h[i] += hh[i];
      const double nrm = std::sqrt(hh[dim + 1]);
      h[dim] = nrm;
      if (nrm != 0.0) TRY(vec_scale(ctx, vv, 1.0 / nrm, n));
      for (int i = 0; i < j; ++i) {
        const double t = h[i];
        h[i] = ci[i] * t + si[i] * h[i + 1];
        h[i + 1] = -si[i] * t + ci[i] * h[i + 1];
      }
      const double s = std::sqrt(h[j] * h[j] + h[j + 1] * h[j + 1]);
      ci[j] = h[j] / s; si[j] = h[j + 1] / s;
      h[j] = s;
      gamma[j + 1] = -si[j] * gamma[j];
      gamma[j] *= ci[j];
      for (int i = 0; i < dim; ++i) H[(size_t)i * restart + j] = h[i];
      rho = std::fabs(gamma[dim]);
      if (rho <= tol) state = 1;
      else if (its >= max_its) state = 2;
    }
    std::vector<double> y(dim);
    for (int i = dim - 1; i >= 0; --i) {
      double s = gamma[i];
      for (int k = i + 1; k < dim; ++k) s -= H[(size_t)i * restart + k] * y[k];
      y[i] = s / H[(size_t)i * restart + i];
    }
    TRY(vec_multi_add(ctx, xb, Z.data(), y.data(), dim, n));     // x += Z y
  }
  TRY(vec_copy(ctx, x_u, xb, nu));
  TRY(vec_copy(ctx, x_p, xb + nu, np));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  res->iterations_expensive = its;
  res->final_residual = rho;
  return state == 1 ? SMF_OK : SMF_NO_CONVERGENCE;
}

// Shadow vectors of SolverIDR: Q_1..Q_{s-1} are filled with std::normal_distribution<double>(0,1) values drawn from a
// default-seeded std::mt19937, entry by entry over the locally owned DoFs (velocity block, then pressure block), the
// stream continuing from one vector to the next.  nvcc's host compiler is g++, so this IS deal.II's stream (libstdc++).
// The values only depend on (n, s): they are kept on the device between solves.  With several ranks -- or whenever
// global DoF numbers were supplied -- the entries are a hash of the partition-independent global DoF number instead
// (every copy of a shared DoF must get the same value, and a P-rank run then repeats the 1-rank run).
static uint64_t mix64(uint64_t z)
{ z += 0x9e3779b97f4a7c15ull; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
static int idr_shadow_vectors(smf_ctx *ctx, int s, int64_t nu, int64_t np)
{
  const int64_t n = nu + np;
  if (ctx->idr_normals && ctx->idr_normals_n == (int64_t)(s - 1) * n) return SMF_OK;
  cudaFree(ctx->idr_normals); ctx->idr_normals = nullptr; ctx->idr_normals_n = 0;
  if (s < 2) return SMF_OK;
  std::vector<double> host((size_t)(s - 1) * n);
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  if (ctx->world == 1 && !(A.ex_u.gid && A.ex_p.gid)) {
    std::mt19937 rng;
    std::normal_distribution<double> normal(0.0, 1.0);
    for (double &v : host) v = normal(rng);
  } else {
    if (!A.ex_u.gid || !A.ex_p.gid) return ctx->fail(SMF_ERR_INVALID, "IDR(s) on several ranks needs the global DoF numbers (smf_set_exchange)");
    std::vector<int64_t> gid(n);
    SMF_CUDA(cudaMemcpy(gid.data(), A.ex_u.gid, sizeof(int64_t) * nu, cudaMemcpyDeviceToHost));
    SMF_CUDA(cudaMemcpy(gid.data() + nu, A.ex_p.gid, sizeof(int64_t) * np, cudaMemcpyDeviceToHost));
    for (int i = 0; i < s - 1; ++i)
      for (int64_t k = 0; k < n; ++k) {
        const uint64_t key = mix64((uint64_t)gid[k] * 2 + (k >= nu ? 1 : 0) + ((uint64_t)(i + 1) << 56));
        const double u1 = ((double)(mix64(key) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        const double u2 = ((double)(mix64(key ^ 0x5bf03635ull) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        host[(size_t)i * n + k] = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
      }
  }
  SMF_CUDA(cudaMalloc(&ctx->idr_normals, sizeof(double) * host.size()));
  SMF_CUDA(cudaMemcpy(ctx->idr_normals, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice));
  ctx->idr_normals_n = (int64_t)host.size();
  ctx->bytes += sizeof(double) * host.size();
  return SMF_OK;
}

// SolverIDR(AdditionalData(s)) with the cheap preconditioner (local_smoothing.cc:1286-1297; deal.II 9.6 semantics in
// SURVEY App. A.5): s + 1 preconditioner and operator applications per complete step, early exit inside the step
static int idr(smf_ctx *ctx, const smf_solve_params *prm, double *x_u, double *x_p, const double *b_u, const double *b_p,
               double tol, smf_solve_result *res)
{
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  const int64_t nu = A.n_u, np = A.n_p, n = nu + np;
  const int s = std::max(1, prm->idr_s_parameter);
  const int max_its = prm->n_cheap_stokes_solver_steps;
  TRY(ensure_krylov(ctx, 3 * s + 4));
  DotWeights dw(ctx, ctx->block_weights);
  std::vector<double *> &K = ctx->krylov;
  double *xb = K[0], *r = K[1], *v = K[2], *uhat = K[3];
  std::vector<double *> G(K.begin() + 4, K.begin() + 4 + s), U(K.begin() + 4 + s, K.begin() + 4 + 2 * s),
      Q(K.begin() + 4 + 2 * s, K.begin() + 4 + 3 * s);
  TRY(vec_copy(ctx, xb, x_u, nu));
  TRY(vec_copy(ctx, xb + nu, x_p, np));
  TRY(launch_vmult_stokes(ctx, A, r, r + nu, xb, xb + nu));
  TRY(vec_sadd(ctx, r, -1.0, 1.0, b_u, nu));
  TRY(vec_sadd(ctx, r + nu, -1.0, 1.0, b_p, np));
  double r2;
  TRY(vec_norm2(ctx, r, n, &r2));
  double rho = std::sqrt(r2);
  res->initial_residual = rho;
  int step = 0, state = rho <= tol ? 1 : 0;
  if (state == 0) {
    TRY(idr_shadow_vectors(ctx, s, nu, np));
    for (int i = 0; i < s; ++i) {
      TRY(vec_set(ctx, G[i], 0.0, n));
      TRY(vec_set(ctx, U[i], 0.0, n));
      if (i == 0) TRY(vec_copy(ctx, Q[0], r, n));
      else TRY(vec_copy(ctx, Q[i], ctx->idr_normals + (size_t)(i - 1) * n, n));
      for (int j = 0; j < i; ++j) {      // v = Q_j (Q_j . q) / (q . q); q -= v
        double a, b;
        TRY(vec_dot(ctx, Q[j], Q[i], n, &a));
        TRY(vec_norm2(ctx, Q[i], n, &b));
        TRY(vec_axpy(ctx, Q[i], -a / b, Q[j], n));
      }
      if (i != 0) {
        double b;
        TRY(vec_norm2(ctx, Q[i], n, &b));
        TRY(vec_scale(ctx, Q[i], 1.0 / std::sqrt(b), n));
      }
    }
  }
  std::vector<double> M((size_t)s * s, 0.0), phi(s + 1), gamma(s), Mk((size_t)s * s), rhs(s);
  for (int i = 0; i < s; ++i) M[(size_t)i * s + i] = 1.0;
  double omega = 1.0;
  auto solve_sub = [&](int k) {      // M(k:s, k:s) gamma = phi(k:s): Gaussian elimination with partial pivoting
    const int m = s - k;
    for (int i = 0; i < m; ++i) { rhs[i] = phi[k + i]; for (int j = 0; j < m; ++j) Mk[(size_t)i * m + j] = M[(size_t)(k + i) * s + k + j]; }
    for (int c = 0; c < m; ++c) {
      int piv = c;
      for (int i = c + 1; i < m; ++i) if (std::fabs(Mk[(size_t)i * m + c]) > std::fabs(Mk[(size_t)piv * m + c])) piv = i;
      if (piv != c) { for (int j = 0; j < m; ++j) std::swap(Mk[(size_t)c * m + j], Mk[(size_t)piv * m + j]); std::swap(rhs[c], rhs[piv]); }
      for (int i = c + 1; i < m; ++i) {
        const double f = Mk[(size_t)i * m + c] / Mk[(size_t)c * m + c];
        for (int j = c; j < m; ++j) Mk[(size_t)i * m + j] -= f * Mk[(size_t)c * m + j];
        rhs[i] -= f * rhs[c];
      }
    }
    for (int i = m - 1; i >= 0; --i) {
      double t = rhs[i];
      for (int j = i + 1; j < m; ++j) t -= Mk[(size_t)i * m + j] * gamma[j];
      gamma[i] = t / Mk[(size_t)i * m + i];
    }
  };
  while (state == 0) {
    ++step;
    {
      VecBucket vb(ctx, B_ORTHO);
      TRY(vec_multi_dot(ctx, Q.data(), s, r, n, phi.data(), true));       // phi_i = Q_i . r
    }
    bool early_exit = false;
    for (int k = 0; k < s; ++k) {
      solve_sub(k);
      std::vector<double> mg(s - k);
      for (int j = 0; j < s - k; ++j) mg[j] = -gamma[j];
      TRY(vec_copy(ctx, v, r, n));
      TRY(vec_multi_add(ctx, v, G.data() + k, mg.data(), s - k, n));      // v = r - sum gamma_j G_{k+j}
      TRY(precondition(ctx, uhat, uhat + nu, v, v + nu));
      if (omega != 1.0) TRY(vec_scale(ctx, uhat, omega, n));
      TRY(vec_multi_add(ctx, uhat, U.data() + k, gamma.data(), s - k, n));   // uhat = omega P^-1 v + sum gamma_j U_{k+j}
      TRY(launch_vmult_stokes(ctx, A, G[k], G[k] + nu, uhat, uhat + nu));
      if (k > 0) {                    // orthogonalise G_k against Q_0..Q_{k-1}, same combination on uhat
        double alpha;
        TRY(vec_dot(ctx, Q[0], G[k], n, &alpha));
        alpha /= M[0];
        for (int i = 1; i < k; ++i) {
          TRY(vec_axpy(ctx, G[k], -alpha, G[i - 1], n));
          TRY(vec_axpy(ctx, uhat, -alpha, U[i - 1], n));
          TRY(vec_dot(ctx, Q[i], G[k], n, &alpha));
          alpha /= M[(size_t)i * s + i];
        }
        TRY(vec_axpy(ctx, G[k], -alpha, G[k - 1], n));
        TRY(vec_axpy(ctx, uhat, -alpha, U[k - 1], n));
      }
      {                               // column k of M: M(i, k) = Q_i . G_k, i >= k
        std::vector<double> col(s - k + 1);
        VecBucket vb(ctx, B_ORTHO);
        TRY(vec_multi_dot(ctx, Q.data() + k, s - k, G[k], n, col.data(), true));
        for (int i = k; i < s; ++i) M[(size_t)i * s + k] = col[i - k];
      }
      const double beta = phi[k] / M[(size_t)k * s + k];
      TRY(vec_axpy(ctx, r, -beta, G[k], n));
      TRY(vec_axpy(ctx, xb, beta, uhat, n));
      std::swap(U[k], uhat);
      TRY(vec_norm2(ctx, r, n, &r2));
      rho = std::sqrt(r2);
      if (rho <= tol) { state = 1; early_exit = true; break; }
      if (step >= max_its) { state = 2; early_exit = true; break; }
      if (k + 1 < s) {
        for (int i = 0; i <= k; ++i) phi[i] = 0.0;
        for (int i = k + 1; i < s; ++i) phi[i] -= beta * M[(size_t)i * s + k];
      }
    }
    if (early_exit) break;
    TRY(precondition(ctx, uhat, uhat + nu, r, r + nu));
    TRY(launch_vmult_stokes(ctx, A, v, v + nu, uhat, uhat + nu));
    double vr, vv;
    TRY(vec_dot(ctx, v, r, n, &vr));
    TRY(vec_norm2(ctx, v, n, &vv));
    omega = vr / vv;
    TRY(vec_axpy(ctx, r, -omega, v, n));
    TRY(vec_norm2(ctx, r, n, &r2));
    rho = std::sqrt(r2);
    TRY(vec_axpy(ctx, xb, omega, uhat, n));
    if (rho <= tol) state = 1;
    else if (step >= max_its) state = 2;
  }
  TRY(vec_copy(ctx, x_u, xb, nu));
  TRY(vec_copy(ctx, x_p, xb + nu, np));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  res->iterations = step;
  res->final_residual = rho;
  return state == 1 ? SMF_OK : SMF_NO_CONVERGENCE;
}

}  // namespace smf

extern "C" {

int smf_vmult_stokes(smf_ctx *ctx, double *dst_u, double *dst_p, const double *src_u, const double *src_p)
{ if (!ctx) return SMF_ERR_INVALID; return launch_vmult_stokes(ctx, ctx->lev(SMF_ACTIVE), dst_u, dst_p, src_u, src_p); }
int smf_vmult_A(smf_ctx *ctx, int level, double *dst, const double *src)
{
  if (!ctx) return SMF_ERR_INVALID;
  if (level < SMF_ACTIVE || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  if (level >= 0 && ctx->levels[level].edge_u.n) {       // level operator with refinement-edge rows: A_l = (I - E) A_bc (I - E) + E
    MGSide S{true, nullptr, nullptr, 1};
    return lapply(ctx, S, level, dst, src);
  }
  return launch_vmult_A(ctx, ctx->lev(level), dst, src);
}
int smf_vmult_BT(smf_ctx *ctx, double *dst_u, const double *src_p)
{ if (!ctx) return SMF_ERR_INVALID; return launch_vmult_BT(ctx, ctx->lev(SMF_ACTIVE), dst_u, src_p); }
int smf_vmult_mass(smf_ctx *ctx, int level, double *dst, const double *src)
{
  if (!ctx) return SMF_ERR_INVALID;
  if (level < SMF_ACTIVE || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  if (level >= 0 && ctx->levels[level].edge_p.n) {
    MGSide S{false, nullptr, nullptr, 4};
    return lapply(ctx, S, level, dst, src);
  }
  return launch_vmult_mass(ctx, ctx->lev(level), dst, src);
}

// StokesMatrixFreeHandler::correct_stokes_rhs (local_smoothing.cc:617-754): rhs -= Stokes(u0), u0 read with
// read_dof_values_plain (boundary values kept), constrained rows receive nothing
int smf_correct_stokes_rhs(smf_ctx *ctx, double *rhs_u, double *rhs_p, const double *u0_u, const double *u0_p)
{
  if (!ctx) return SMF_ERR_INVALID;
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  TRY(ensure_tmp(ctx));
  TRY(launch_vmult_stokes(ctx, A, ctx->tmp_u, ctx->tmp_p, u0_u, u0_p, true));
  TRY(vec_axpy(ctx, rhs_u, -1.0, ctx->tmp_u, A.n_u));
  TRY(vec_axpy(ctx, rhs_p, -1.0, ctx->tmp_p, A.n_p));
  return SMF_OK;
}

int smf_vmult_stokes_host(smf_ctx *ctx, double *dst_u, double *dst_p, const double *src_u, const double *src_p)
{
  if (!ctx) return SMF_ERR_INVALID;
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  TRY(ensure_tmp(ctx));
  SMF_CUDA(cudaMemcpyAsync(ctx->tmp_u, src_u, sizeof(double) * A.n_u, cudaMemcpyHostToDevice, ctx->stream));
  SMF_CUDA(cudaMemcpyAsync(ctx->tmp_p, src_p, sizeof(double) * A.n_p, cudaMemcpyHostToDevice, ctx->stream));
  TRY(launch_vmult_stokes(ctx, A, ctx->tmp_u2, ctx->tmp_p2, ctx->tmp_u, ctx->tmp_p));
  SMF_CUDA(cudaMemcpyAsync(dst_u, ctx->tmp_u2, sizeof(double) * A.n_u, cudaMemcpyDeviceToHost, ctx->stream));
  SMF_CUDA(cudaMemcpyAsync(dst_p, ctx->tmp_p2, sizeof(double) * A.n_p, cudaMemcpyDeviceToHost, ctx->stream));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  return SMF_OK;
}

int smf_compute_diagonals(smf_ctx *ctx)
{
  if (!ctx) return SMF_ERR_INVALID;
  for (int l = 0; l < ctx->cfg.n_levels; ++l) {
    LevelDev &L = ctx->levels[l];
    if (L.n_cells == 0 || !L.has_cell_data) return ctx->fail(SMF_ERR_INVALID, "level mesh / cell data missing");
    TRY(launch_diag(ctx, L));
  }
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  if (A.n_cells && A.has_cell_data) TRY(launch_diag(ctx, A));
  ctx->smoothers_ready = false;
  return SMF_OK;
}

int smf_get_inverse_diagonal(smf_ctx *ctx, int level, int which, double *host_out)
{
  if (!ctx || !host_out) return SMF_ERR_INVALID;
  if (level < SMF_ACTIVE || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  LevelDev &L = ctx->lev(level);
  if (!L.has_diag) return ctx->fail(SMF_ERR_INVALID, "diagonals not computed");
  return smf_vec_d2h(ctx, host_out, which == 0 ? L.dinv_A : L.dinv_M, which == 0 ? L.n_u : L.n_p);
}

int smf_set_inverse_diagonal_entries(smf_ctx *ctx, int level, int which, int64_t n, const uint32_t *idx, const double *values)
{
  if (!ctx || n < 0 || (n > 0 && (!idx || !values)) || (which != 0 && which != 1)) return SMF_ERR_INVALID;
  if (level < SMF_ACTIVE || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  LevelDev &L = ctx->lev(level);
  if (!L.has_diag) return ctx->fail(SMF_ERR_INVALID, "call smf_compute_diagonals first");
  const int64_t nd = which == 0 ? L.n_u : L.n_p;
  double *d = which == 0 ? L.dinv_A : L.dinv_M;
  // set-up path: one round trip of the vector (the entries are few: masters of constraint lines)
  std::vector<double> h((size_t)nd);
  TRY(smf_vec_d2h(ctx, h.data(), d, nd));
  for (int64_t i = 0; i < n; ++i) {
    if (idx[i] >= nd) return ctx->fail(SMF_ERR_INVALID, "smf_set_inverse_diagonal_entries: index out of range");
    if (!(values[i] > 0.0)) return ctx->fail(SMF_ERR_INVALID, "smf_set_inverse_diagonal_entries: entries of a positive definite operator's diagonal must be positive");
    h[idx[i]] = values[i];
  }
  TRY(smf_vec_h2d(ctx, d, h.data(), nd));
  ctx->smoothers_ready = false;
  return SMF_OK;
}

int smf_setup_smoothers(smf_ctx *ctx)
{
  if (!ctx) return SMF_ERR_INVALID;
  const int nl = ctx->cfg.n_levels;
  for (int l = 0; l < nl; ++l) {
    LevelDev &L = ctx->levels[l];
    if (!L.has_diag) return ctx->fail(SMF_ERR_INVALID, "call smf_compute_diagonals first");
    if (l > 0 && !ctx->transfers[l].set) return ctx->fail(SMF_ERR_INVALID, "transfer missing");
    for (int side = 0; side < 2; ++side) {
      MGVectors &V = side == 0 ? ctx->mgv_A[l] : ctx->mgv_S[l];
      const int64_t n = side == 0 ? L.n_u : L.n_p;
      if (!V.x) {
        SMF_CUDA(cudaMalloc(&V.x, sizeof(double) * n)); SMF_CUDA(cudaMalloc(&V.b, sizeof(double) * n));
        SMF_CUDA(cudaMalloc(&V.t, sizeof(double) * n)); SMF_CUDA(cudaMalloc(&V.xo, sizeof(double) * n));
        ctx->bytes += 4 * sizeof(double) * n;
      }
      Smoother &sm = side == 0 ? ctx->smoother_A[l] : ctx->smoother_S[l];
      if (l > 0) { sm.degree = 4; sm.smoothing_range = 15.0; sm.eig_cg_n_iterations = 10; }
      else { sm.degree = 8; sm.smoothing_range = 1e-3; sm.eig_cg_n_iterations = 100; }
      MGSide S{side == 0, side == 0 ? &ctx->smoother_A : &ctx->smoother_S, side == 0 ? &ctx->mgv_A : &ctx->mgv_S, side == 0 ? 1 : 4};
      TRY(estimate_eigenvalues(ctx, S, l));
    }
  }
  ctx->smoothers_ready = true;
  ctx->graph_A_stale = ctx->graph_S_stale = true;
  ctx->graph_level = -1;
  for (int l = 0; l < nl; ++l)
    if (ctx->levels[l].n_cells <= GRAPH_MAX_CELLS) ctx->graph_level = l;
  ctx->tail_level = std::min(schur_tail_top_level(ctx), nl - 2);            // (never the finest level: its defect is the caller's vector)
  TRY(lattice_prepare(ctx));
  TRY(replica_prepare(ctx));
  return SMF_OK;
}

int smf_get_eigenvalue_estimates(smf_ctx *ctx, int level, int which, double *out)
{
  if (!ctx || !out || level < 0 || level >= ctx->cfg.n_levels) return SMF_ERR_INVALID;
  const Smoother &sm = which == 0 ? ctx->smoother_A[level] : ctx->smoother_S[level];
  out[0] = sm.lambda_min; out[1] = sm.lambda_max; out[2] = sm.cg_its;
  return SMF_OK;
}

int smf_prolongate(smf_ctx *ctx, int which, int to_level, double *dst, const double *src)
{ if (!ctx) return SMF_ERR_INVALID; return launch_prolongate(ctx, which, to_level, dst, src, false); }
int smf_restrict_add(smf_ctx *ctx, int which, int from_level, double *dst, const double *src)
{ if (!ctx) return SMF_ERR_INVALID; return launch_restrict_add(ctx, which, from_level, dst, src); }

int smf_profile_enable(smf_ctx *ctx, int mode)
{
  if (!ctx || mode < 0 || mode > 2) return SMF_ERR_INVALID;
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto &e : ctx->prof_live) { ctx->prof_pool.push_back(e.a); ctx->prof_pool.push_back(e.b); }
  ctx->prof_live.clear();
  for (int b = 0; b < B_COUNT; ++b) { ctx->prof_ms[b] = 0.0; ctx->prof_cnt[b] = 0; }
  ctx->prof_mode = mode;
  return SMF_OK;
}

int smf_profile_read(smf_ctx *ctx, double *ms_out, int64_t *count_out)
{
  if (!ctx || !ms_out || !count_out) return SMF_ERR_INVALID;
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto &e : ctx->prof_live) {
    float ms = 0.f;
    SMF_CUDA(cudaEventElapsedTime(&ms, e.a, e.b));
    ctx->prof_ms[e.bucket] += ms; ctx->prof_cnt[e.bucket]++;
    ctx->prof_pool.push_back(e.a); ctx->prof_pool.push_back(e.b);
  }
  ctx->prof_live.clear();
  for (int b = 0; b < B_COUNT; ++b) { ms_out[b] = ctx->prof_ms[b]; count_out[b] = ctx->prof_cnt[b]; }
  return SMF_OK;
}

// Simulator::normalize_pressure (source/simulator/helper_functions.cc:765-870) and the rescaling of the pressure to
// physical units (local_smoothing.cc:1421) on the device.  weights[i] = integral of the i-th pressure shape function
// over the top boundary ("surface") or over this rank's cells ("volume"); the adapter computes them once per mesh.
int smf_set_pressure_normalization(smf_ctx *ctx, const double *weights)
{
  if (!ctx) return SMF_ERR_INVALID;
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  if (A.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set the active level mesh first");
  cudaFree(ctx->pnorm_weights); ctx->pnorm_weights = nullptr; ctx->pnorm_area = 0.0;
  if (!weights) return SMF_OK;
  SMF_CUDA(cudaMalloc(&ctx->pnorm_weights, sizeof(double) * A.n_p));
  SMF_CUDA(cudaMemcpy(ctx->pnorm_weights, weights, sizeof(double) * A.n_p, cudaMemcpyHostToDevice));
  double area = 0.0;
  for (int64_t i = 0; i < A.n_p; ++i) area += weights[i];
  if (ctx->world > 1) {      // Utilities::MPI::sum of the per-rank integrals
    double *d = nullptr;
    SMF_CUDA(cudaMalloc(&d, sizeof(double)));
    SMF_CUDA(cudaMemcpyAsync(d, &area, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    TRY(allreduce_sum(ctx, d, 1));
    SMF_CUDA(cudaMemcpyAsync(&area, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SMF_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
  }
  if (!(area > 0.0)) return ctx->fail(SMF_ERR_INVALID, "the area / volume to integrate the pressure over is zero or negative");
  ctx->pnorm_area = area;
  return SMF_OK;
}

int smf_postprocess_pressure(smf_ctx *ctx, double *sol_p, double pressure_scaling, int normalize, double surface_pressure,
                             double *adjustment_out)
{
  if (!ctx || !sol_p) return SMF_ERR_INVALID;
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  if (pressure_scaling != 1.0) TRY(vec_scale(ctx, sol_p, pressure_scaling, A.n_p));
  double adjustment = 0.0;
  if (normalize) {
    if (!ctx->pnorm_weights) return ctx->fail(SMF_ERR_INVALID, "call smf_set_pressure_normalization first");
    DotWeights dw(ctx, nullptr);          // the weights already partition the integral between the ranks
    double integral;
    TRY(vec_dot(ctx, ctx->pnorm_weights, sol_p, A.n_p, &integral));
    adjustment = -integral / ctx->pnorm_area + surface_pressure;
    TRY(vec_shift(ctx, sol_p, adjustment, A.n_p));
  }
  if (adjustment_out) *adjustment_out = adjustment;
  return SMF_OK;
}

// right-hand side of the velocity block from the plugins' force density at the quadrature points (host array
// [n_cells][3^dim][dim] = density * gravity), assembled on the device into rhs_u (device vector)
int smf_assemble_rhs(smf_ctx *ctx, const double *force_q, double *rhs_u)
{
  if (!ctx || !force_q || !rhs_u) return SMF_ERR_INVALID;
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  if (A.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set the active level mesh first");
  const size_t nb = sizeof(double) * (size_t)A.n_cells * (A.dim == 3 ? 27 : 9) * A.dim;
  double *d = nullptr;
  SMF_CUDA(cudaMalloc(&d, nb));
  SMF_CUDA(cudaMemcpyAsync(d, force_q, nb, cudaMemcpyHostToDevice, ctx->stream));
  const int rc = launch_rhs(ctx, A, d, rhs_u);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  return rc;
}

int smf_vcycle_A(smf_ctx *ctx, double *dst, const double *src) { if (!ctx) return SMF_ERR_INVALID; return vcycle(ctx, true, dst, src); }
int smf_vcycle_S(smf_ctx *ctx, double *dst, const double *src) { if (!ctx) return SMF_ERR_INVALID; return vcycle(ctx, false, dst, src); }
int smf_precondition(smf_ctx *ctx, double *dst_u, double *dst_p, const double *src_u, const double *src_p)
{ if (!ctx) return SMF_ERR_INVALID; return precondition(ctx, dst_u, dst_p, src_u, src_p); }

int smf_solve(smf_ctx *ctx, const smf_solve_params *prm, double *sol_u, double *sol_p,
              const double *rhs_u, const double *rhs_p, smf_solve_result *result)
{
  if (!ctx || !prm || !result) return SMF_ERR_INVALID;
  if (prm->krylov_method != 0 && prm->krylov_method != 1) return ctx->fail(SMF_ERR_INVALID, "krylov_method: 0 = GMRES, 1 = IDR(s)");
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  if (A.n_cells == 0 || !A.has_cell_data) return ctx->fail(SMF_ERR_INVALID, "active level not set");
  std::memset(result, 0, sizeof(*result));
  ctx->err.clear();
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  const double t0 = now_s();
  // the reference rebuilds the smoothers (incl. eigenvalue estimation) at the start of every solve (:768-849)
  TRY(smf_setup_smoothers(ctx));
  TRY(ensure_tmp(ctx));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  const double t1 = now_s();
  if (ctx->h2d_pending) {      // smf_solve_host: the vectors arrive on the copy stream
    ctx->h2d_pending = false;
    SMF_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->h2d_event, 0));
    TRY(launch_set_entries(ctx, sol_u, A.constrained_u, A.n_constrained_u, 0.0));
    TRY(launch_set_entries(ctx, sol_p, A.constrained_p, A.n_constrained_p, 0.0));
  }
  // tolerance (:1016-1025): residual with the velocity part of the guess zeroed
  const int64_t nu = A.n_u, np = A.n_p;
  TRY(ensure_krylov(ctx, 2));
  double *g = ctx->krylov[0], *r = ctx->krylov[1];
  TRY(vec_set(ctx, g, 0.0, nu));
  TRY(vec_copy(ctx, g + nu, sol_p, np));
  TRY(launch_vmult_stokes(ctx, A, r, r + nu, g, g + nu));
  TRY(vec_sadd(ctx, r, -1.0, 1.0, rhs_u, nu));
  double ru2, rp2;
  if (ctx->world > 1 && !ctx->block_weights) {
    if (!A.ex_u.weights || !A.ex_p.weights) return ctx->fail(SMF_ERR_INVALID, "smf_set_exchange missing for the active level");
    SMF_CUDA(cudaMalloc(&ctx->block_weights, sizeof(double) * (nu + np)));
    SMF_CUDA(cudaMemcpyAsync(ctx->block_weights, A.ex_u.weights, sizeof(double) * nu, cudaMemcpyDeviceToDevice, ctx->stream));
    SMF_CUDA(cudaMemcpyAsync(ctx->block_weights + nu, A.ex_p.weights, sizeof(double) * np, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  { DotWeights dw(ctx, A.ex_u.weights); TRY(vec_norm2(ctx, r, nu, &ru2)); }
  { DotWeights dw(ctx, A.ex_p.weights); TRY(vec_norm2(ctx, rhs_p, np, &rp2)); }
  const double tol = prm->linear_solver_tolerance * std::sqrt(ru2 + rp2);
  result->tolerance = tol;
  ctx->n_A_prec = ctx->n_S_prec = ctx->n_A_exp = ctx->n_S_exp = 0;
  // step 1a: the cheap phase (skipped when no cheap steps are allowed, :1270-1272)
  int rc = SMF_NO_CONVERGENCE;
  if (prm->n_cheap_stokes_solver_steps > 0)
    rc = prm->krylov_method == 0 ? gmres(ctx, prm, sol_u, sol_p, rhs_u, rhs_p, tol, result) : idr(ctx, prm, sol_u, sol_p, rhs_u, rhs_p, tol, result);
  // step 1b: FGMRES with inner CG solves, starting from the last cheap iterate (:1317-1350)
  if (rc == SMF_NO_CONVERGENCE && prm->n_expensive_stokes_solver_steps > 0)
    rc = fgmres(ctx, prm, sol_u, sol_p, rhs_u, rhs_p, tol, result);
  const double t2 = now_s();
  result->n_A_preconditioner = ctx->n_A_prec;
  result->n_S_preconditioner = ctx->n_S_prec;
  result->n_A_expensive = ctx->n_A_exp;
  result->n_S_expensive = ctx->n_S_exp;
  result->setup_seconds = t1 - t0;
  result->solve_seconds = t2 - t1;
  if (rc == SMF_NO_CONVERGENCE && ctx->err.empty()) ctx->err = "iterative Stokes solver did not converge (SolverControl::NoConvergence)";
  // back to physical units, pressure normalisation (local_smoothing.cc:1421, :1441-1443)
  if ((rc == SMF_OK || rc == SMF_NO_CONVERGENCE) && (prm->rescale_pressure || prm->normalize_pressure)) {
    const std::string keep = ctx->err;
    TRY(smf_postprocess_pressure(ctx, sol_p, prm->rescale_pressure ? A.ps : 1.0, prm->normalize_pressure, prm->surface_pressure,
                                 &result->pressure_normalization_adjustment));
    ctx->err = keep;
  }
  return rc;
}

int smf_solve_host(smf_ctx *ctx, const smf_solve_params *prm, double *sol_u, double *sol_p,
                   const double *rhs_u, const double *rhs_p, smf_solve_result *result)
{
  if (!ctx) return SMF_ERR_INVALID;
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  const int64_t nu = A.n_u, np = A.n_p;
  if (ctx->io_n != 2 * (nu + np)) {      // persistent staging of the host-facing vectors (a17: the H2D/D2H point)
    cudaFree(ctx->io_buf); ctx->io_buf = nullptr; ctx->io_n = 0;
    SMF_CUDA(cudaMalloc(&ctx->io_buf, sizeof(double) * 2 * (nu + np)));
    ctx->io_n = 2 * (nu + np);
    ctx->bytes += sizeof(double) * 2 * (nu + np);
  }
  double *buf = ctx->io_buf;
  double *su = buf, *sp = buf + nu, *bu = buf + nu + np, *bp = bu + nu;
  // The right-hand side and the guess travel on a copy stream while the solve's first phase -- smoother set-up and eigenvalue
  // estimation, which need neither -- runs on the context's stream; smf_solve waits for the event before it touches them and
  // then zeroes the constrained entries of the guess (get_current_constraints().set_zero(guess), local_smoothing.cc:973: a guess
  // that carries boundary values would put b - x into the constrained rows and change the initial residual).
  if (!ctx->copy_stream) {
    SMF_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    SMF_CUDA(cudaEventCreateWithFlags(&ctx->h2d_event, cudaEventDisableTiming));
  }
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));      // (the staging buffers may still be read by work queued before this call)
  SMF_CUDA(cudaMemcpyAsync(bu, rhs_u, sizeof(double) * nu, cudaMemcpyHostToDevice, ctx->copy_stream));
  SMF_CUDA(cudaMemcpyAsync(bp, rhs_p, sizeof(double) * np, cudaMemcpyHostToDevice, ctx->copy_stream));
  SMF_CUDA(cudaMemcpyAsync(su, sol_u, sizeof(double) * nu, cudaMemcpyHostToDevice, ctx->copy_stream));
  SMF_CUDA(cudaMemcpyAsync(sp, sol_p, sizeof(double) * np, cudaMemcpyHostToDevice, ctx->copy_stream));
  SMF_CUDA(cudaEventRecord(ctx->h2d_event, ctx->copy_stream));
  ctx->h2d_pending = true;
  const int rc = smf_solve(ctx, prm, su, sp, bu, bp, result);
  if (ctx->h2d_pending) { cudaStreamSynchronize(ctx->copy_stream); ctx->h2d_pending = false; }      // (smf_solve failed before it consumed the event)
  if (rc == SMF_OK || rc == SMF_NO_CONVERGENCE) {
    cudaMemcpyAsync(sol_u, su, sizeof(double) * nu, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(sol_p, sp, sizeof(double) * np, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  return rc;
}

}  // extern "C"
