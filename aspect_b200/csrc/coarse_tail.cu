// The coarse tail of the Schur-complement V-cycle in ONE kernel launch.
//
// Below a few thousand cells every kernel of the V-cycle is pure launch latency: the pressure-mass hierarchy
// (Multigrid with 4 Chebyshev(4) sweeps before and after the coarse correction, local_smoothing.cc:800-825, :893-916) issues
// ~68 dependent launches per level and V-cycle -- 18,000 launches and 235 ms per refinement-7 solve, half of a refinement-5
// solve -- for vectors of a few hundred to a few thousand doubles.  This kernel runs levels 0..k of that hierarchy inside a
// single thread block: the four level vectors of every level (x, x_old, defect, t) live in shared memory (levels 0..4 of a 3D
// box: 5,802 pressure DoFs, 186 kB), the level operators are the same arithmetic as mass_kernel (per cell (ps^2 vol / eta)
// M1 x M1 x M1 on the 2^dim nodal values, accumulated with shared-memory atomics), the smoother is the same three-term
// recurrence as cheb_first / cheb_update / cheb_tail (solver.cu), the transfers use the patch tables of smf_set_transfer with
// their owner bits.  Block-wide barriers replace kernel boundaries (~0.2 us instead of ~3 us each).
// Semantics: Multigrid::level_v_step + SmootherRelaxation + MGCoarseGridApplySmoother as in solver.cu: level_v_step.
#include "smf_internal.h"

namespace smf {

constexpr int TAIL_MAX_LEVELS = 6;
constexpr int TAIL_THREADS = 1024;

struct TailLevel {
  int n_cells, n_dofs, n_parents;
  const uint32_t *cell_nodes;     // [n_cells][2^dim] (index | bit 31 = constrained: not supported here, checked on the host)
  const double *eta;              // per cell (level tables with one value per cell)
  const double *dinv;
  const uint32_t *patch;          // transfer (l-1 -> l): [n_parents][3^dim] fine node | bit 31 writes (prolongation) | bit 30 contributes (restriction)
  const uint32_t *coarse_nodes;   // cell_nodes of level l-1
  double scale;                   // ps^2 * cell volume
  double theta, delta; int degree;
  int off;                        // offset of this level's vectors in shared memory (doubles): x, xo, b, t each n_dofs
};
struct TailParams {
  int n_levels, steps;
  TailLevel lv[TAIL_MAX_LEVELS];
  const double *defect;           // defect of level n_levels-1 (global memory)
  double *solution;               // solution of level n_levels-1 (global memory)
};

template <int DIM> __device__ __forceinline__ void tail_mass(const TailLevel &L, const double *x, double *t)
{
  constexpr int NP = 1 << DIM;
  for (int c = threadIdx.x; c < L.n_cells; c += blockDim.x) {
    double p[NP];
    uint32_t idx[NP];
#pragma unroll
    for (int a = 0; a < NP; ++a) { idx[a] = L.cell_nodes[(size_t)c * NP + a] & NODE_IDX_MASK; p[a] = x[idx[a]]; }
    // M1 = [2 1; 1 2] / 6 along every direction
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const int s = 1 << d;
#pragma unroll
      for (int a = 0; a < NP; ++a)
        if (!(a & s)) {
          const double u = p[a], v = p[a + s];
          p[a] = (2.0 * u + v) * (1.0 / 6.0); p[a + s] = (u + 2.0 * v) * (1.0 / 6.0);
        }
    }
    const double coef = L.scale / L.eta[c];
#pragma unroll
    for (int a = 0; a < NP; ++a) atomicAdd(t + idx[a], coef * p[a]);
  }
}

// one Chebyshev sweep on level l: from_zero -> x = p(A) b from a zero start (degree - 1 operator applications), else one
// more sweep from the current x (degree applications); the result ends in the array the returned pointer names
template <int DIM> __device__ double *tail_cheb(const TailLevel &L, double *x, double *xo, const double *b, double *t, bool from_zero)
{
  const int n = L.n_dofs;
  bool have_old;
  if (from_zero) {
    const double it = 1.0 / L.theta;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { x[i] = it * L.dinv[i] * b[i]; t[i] = 0.0; }
    have_old = false;
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) t[i] = 0.0;
    __syncthreads();
    tail_mass<DIM>(L, x, t);
    __syncthreads();
    const double f2 = 1.0 / L.theta;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const double xi = x[i]; xo[i] = xi + f2 * L.dinv[i] * (b[i] - t[i]); t[i] = 0.0; }
    double *s = x; x = xo; xo = s;
    have_old = true;
  }
  __syncthreads();
  if (L.degree < 2 || fabs(L.delta) < 1e-40) return x;
  double rhok = L.delta / L.theta;
  const double sigma = L.theta / L.delta;
  for (int k = 0; k < L.degree - 1; ++k) {
    tail_mass<DIM>(L, x, t);
    __syncthreads();
    const double rhokp = 1.0 / (2.0 * sigma - rhok);
    const double f1 = rhokp * rhok, f2 = 2.0 * rhokp / L.delta;
    rhok = rhokp;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double xi = x[i];
      double v = xi + f2 * L.dinv[i] * (b[i] - t[i]);
      v += have_old ? f1 * (xi - xo[i]) : f1 * xi;
      xo[i] = v;
      t[i] = 0.0;
    }
    double *s = x; x = xo; xo = s;
    have_old = true;
    __syncthreads();
  }
  return x;
}

__device__ __forceinline__ double tail_p1(int r, int j) { return r == 0 ? (j == 0 ? 1.0 : 0.0) : (r == 1 ? 0.5 : (j == 1 ? 1.0 : 0.0)); }

template <int DIM>
__global__ void __launch_bounds__(TAIL_THREADS, 1) schur_tail_kernel(const TailParams P)
{
  extern __shared__ double sm[];
  constexpr int NPATCH = DIM == 3 ? 27 : 9, NC = 1 << DIM;
  const int top = P.n_levels - 1;
  double *xs[TAIL_MAX_LEVELS];          // where the current iterate of every level lives
  // defect of the top level
  {
    const TailLevel &L = P.lv[top];
    double *b = sm + L.off + 2 * L.n_dofs;
    for (int i = threadIdx.x; i < L.n_dofs; i += blockDim.x) b[i] = P.defect[i];
  }
  __syncthreads();
  for (int l = top; l >= 0; --l) {
    const TailLevel &L = P.lv[l];
    double *x = sm + L.off, *xo = x + L.n_dofs, *b = xo + L.n_dofs, *t = b + L.n_dofs;
    // pre-smoothing (on level 0: the coarse solver): SmootherRelaxation::apply = vmult + (steps - 1) x step
    double *cur = tail_cheb<DIM>(L, x, xo, b, t, true);
    for (int s = 1; s < P.steps; ++s) cur = tail_cheb<DIM>(L, cur, cur == x ? xo : x, b, t, false);
    xs[l] = cur;
    if (l == 0) break;
    // residual t = b - A x, restricted into the zeroed defect of the coarser level
    tail_mass<DIM>(L, cur, t);          // (t was left zeroed by the smoother)
    const TailLevel &C = P.lv[l - 1];
    double *bc = sm + C.off + 2 * C.n_dofs;
    for (int i = threadIdx.x; i < C.n_dofs; i += blockDim.x) bc[i] = 0.0;
    __syncthreads();
    for (int i = threadIdx.x; i < L.n_dofs; i += blockDim.x) t[i] = b[i] - t[i];
    __syncthreads();
    for (int w = threadIdx.x; w < L.n_parents * NC; w += blockDim.x) {
      const int par = w / NC, j = w % NC;
      double acc = 0.0;
      for (int s = 0; s < NPATCH; ++s) {
        const uint32_t pe = L.patch[(size_t)par * NPATCH + s];
        if (!((pe >> 30) & 1u)) continue;
        double wt = 1.0;
        int r = s, jj = j;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { wt *= tail_p1(r % 3, jj & 1); r /= 3; jj >>= 1; }
        if (wt != 0.0) acc = fma(wt, t[pe & NODE_IDX_MASK], acc);
      }
      if (acc != 0.0) atomicAdd(bc + (L.coarse_nodes[(size_t)par * NC + j] & NODE_IDX_MASK), acc);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L.n_dofs; i += blockDim.x) t[i] = 0.0;
    __syncthreads();
  }
  for (int l = 1; l <= top; ++l) {
    const TailLevel &L = P.lv[l], &C = P.lv[l - 1];
    double *x = sm + L.off, *xo = x + L.n_dofs, *b = xo + L.n_dofs, *t = b + L.n_dofs;
    double *cur = xs[l];
    const double *xc = xs[l - 1];
    // x += P x_coarse: the parent that owns a fine node writes it
    for (int w = threadIdx.x; w < L.n_parents * NPATCH; w += blockDim.x) {
      const int par = w / NPATCH, s = w % NPATCH;
      const uint32_t pe = L.patch[(size_t)par * NPATCH + s];
      if (!(pe >> 31)) continue;
      double acc = 0.0;
      for (int j = 0; j < NC; ++j) {
        double wt = 1.0;
        int r = s, jj = j;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { wt *= tail_p1(r % 3, jj & 1); r /= 3; jj >>= 1; }
        if (wt != 0.0) acc = fma(wt, xc[C.n_dofs ? (L.coarse_nodes[(size_t)par * NC + j] & NODE_IDX_MASK) : 0], acc);
      }
      cur[pe & NODE_IDX_MASK] += acc;
    }
    __syncthreads();
    for (int s = 0; s < P.steps; ++s) cur = tail_cheb<DIM>(L, cur, cur == x ? xo : x, b, t, false);
    xs[l] = cur;
  }
  {
    const TailLevel &L = P.lv[top];
    for (int i = threadIdx.x; i < L.n_dofs; i += blockDim.x) P.solution[i] = xs[top][i];
  }
}

// Largest level the tail may cover: everything must fit one block's shared memory, one viscosity per cell, no constrained
// pressure DoFs, one rank, no refinement edges.  Returns -1 when the tail is not applicable.
int schur_tail_top_level(smf_ctx *ctx)
{
  if (!ctx->use_coarse_tail || ctx->world > 1 || ctx->adaptive) return -1;
  int dev = 0, max_smem = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int64_t doubles = 0;
  int top = -1;
  for (int l = 0; l < ctx->cfg.n_levels && l < TAIL_MAX_LEVELS; ++l) {
    const LevelDev &L = ctx->levels[l];
    if (L.n_eta != 1 || L.n_constrained_p || L.geo || L.edge_p.n || L.n_cells > 40000) break;
    if (l > 0 && !ctx->transfers[l].patch_p) break;
    doubles += 4 * L.n_p;
    if (doubles * (int64_t)sizeof(double) > (int64_t)max_smem - 1024) break;
    top = l;
  }
  return top;
}

// defect = mgv_S[top].b, solution -> mgv_S[top].x (and .res)
int launch_schur_tail(smf_ctx *ctx, int top)
{
  TailParams P{};
  P.n_levels = top + 1; P.steps = 4;
  int64_t off = 0;
  for (int l = 0; l <= top; ++l) {
    const LevelDev &L = ctx->levels[l];
    const Smoother &sm = ctx->smoother_S[l];
    TailLevel &T = P.lv[l];
    T.n_cells = (int)L.n_cells; T.n_dofs = (int)L.n_p;
    T.cell_nodes = L.cell_nodes_p; T.eta = L.eta; T.dinv = L.dinv_M;
    double vol = 1.0;
    for (int d = 0; d < L.dim; ++d) vol *= L.h[d];
    T.scale = L.ps * L.ps * vol;
    T.theta = sm.theta; T.delta = sm.delta; T.degree = sm.degree;
    T.off = (int)off;
    off += 4 * L.n_p;
    if (l > 0) { T.patch = ctx->transfers[l].patch_p; T.n_parents = (int)ctx->transfers[l].n_parents; T.coarse_nodes = ctx->levels[l - 1].cell_nodes_p; }
  }
  MGVectors &V = ctx->mgv_S[top];
  P.defect = V.b; P.solution = V.x;
  const size_t smem = (size_t)off * sizeof(double);
  static size_t configured2 = 0, configured3 = 0;
  size_t &configured = ctx->cfg.dim == 3 ? configured3 : configured2;
  if (smem > configured) {
    if (ctx->cfg.dim == 3) SMF_CUDA(cudaFuncSetAttribute(schur_tail_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else SMF_CUDA(cudaFuncSetAttribute(schur_tail_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  {
    ProfScope prof(ctx, B_MASS);
    if (ctx->cfg.dim == 3) schur_tail_kernel<3><<<1, TAIL_THREADS, smem, ctx->stream>>>(P);
    else schur_tail_kernel<2><<<1, TAIL_THREADS, smem, ctx->stream>>>(P);
  }
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  V.res = V.x;
  V.t_zero = false;
  return SMF_OK;
}

}  // namespace smf
