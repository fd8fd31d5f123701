// The coarse tail of the Schur-complement V-cycle in ONE kernel launch.
//
// Below a few thousand cells every kernel of the V-cycle is pure launch latency: the pressure-mass hierarchy
// (Multigrid with 4 Chebyshev(4) sweeps before and after the coarse correction, local_smoothing.cc:800-825, :893-916) issues
// ~68 dependent launches per level and V-cycle -- 18,000 launches and 235 ms per refinement-7 solve, half of a refinement-5
// solve -- for vectors of a few hundred to a few thousand doubles.  This kernel runs levels 0..k of that hierarchy inside a
// single thread block: the four level vectors of every level (x, x_old, defect, t) live in shared memory (levels 0..4 of a 3D
// box: 5,802 pressure DoFs, 186 kB), the level operators are the same arithmetic as mass_kernel (per cell (ps^2 vol / eta)
// M1 x M1 x M1 on the 2^dim nodal values, evaluated node by node from the cells around a node: no atomics, the Chebyshev update of an iteration fused into the same pass), the smoother is the same three-term
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
  int off;                        // offset (bytes) of this level's block in shared memory, layout below
};
struct TailParams {
  int n_levels, steps, scratch_off;
  TailLevel lv[TAIL_MAX_LEVELS];
  const double *defect;           // defect of level n_levels-1 (global memory)
  double *solution;               // solution of level n_levels-1 (global memory)
};

// shared-memory block of one level: x, xo, b, t, dinv [n_dofs] doubles | coef [n_cells] doubles | (16-byte aligned) nodes
// [n_cells][NP] u16 | adj [n_dofs][NP] u16 (cell * NP + local node of the cells around a node, padded with the index of a zero
// slot of the scratch so that every node sums exactly NP entries)
__host__ __device__ inline size_t tail_align16(size_t b) { return (b + 15) & ~(size_t)15; }
template <int DIM> __host__ __device__ inline size_t tail_nodes_offset(int n_cells, int n_dofs) { return tail_align16((size_t)(5 * n_dofs + n_cells) * 8); }
template <int DIM> __host__ __device__ inline size_t tail_level_bytes(int n_cells, int n_dofs)
{ return tail_align16(tail_nodes_offset<DIM>(n_cells, n_dofs) + (size_t)(n_cells + n_dofs) * (1 << DIM) * 2); }
template <int DIM> struct TailView {
  static constexpr int NP = 1 << DIM;
  double *x, *xo, *b, *t, *dinv, *coef;
  unsigned short *nodes, *adj;
  __device__ TailView(unsigned char *base, const TailLevel &L)
  {
    x = reinterpret_cast<double *>(base + L.off); xo = x + L.n_dofs; b = xo + L.n_dofs; t = b + L.n_dofs; dinv = t + L.n_dofs;
    coef = dinv + L.n_dofs;
    nodes = reinterpret_cast<unsigned short *>(base + L.off + tail_nodes_offset<DIM>(L.n_cells, L.n_dofs));
    adj = nodes + (size_t)L.n_cells * NP;
  }
};
template <int DIM> struct TailIdx;      // NP 16-bit indices in one vector load
template <> struct TailIdx<3> { uint4 v; __device__ __forceinline__ int get(int k) const { const unsigned w = k < 2 ? v.x : (k < 4 ? v.y : (k < 6 ? v.z : v.w)); return (k & 1) ? (int)(w >> 16) : (int)(w & 0xffffu); } };
template <> struct TailIdx<2> { uint2 v; __device__ __forceinline__ int get(int k) const { const unsigned w = k < 2 ? v.x : v.y; return (k & 1) ? (int)(w >> 16) : (int)(w & 0xffffu); } };
template <int DIM> __device__ __forceinline__ TailIdx<DIM> tail_load_idx(const unsigned short *p)
{
  TailIdx<DIM> r;
  if constexpr (DIM == 3) r.v = *reinterpret_cast<const uint4 *>(p); else r.v = *reinterpret_cast<const uint2 *>(p);
  return r;
}

// Level operator in two passes without atomics and with short dependency chains (one SM runs this: latency is everything):
// pass 1, one thread per cell: the cell's NP node indices in ONE vector load, NP independent loads of x,
// y[cell][a] = (ps^2 vol / eta) (M1 x M1 x M1 p)_a (M1 = [2 1; 1 2] / 6; the arithmetic of mass_kernel); pass 2, one thread per
// node: its NP adjacency entries in one vector load, NP independent loads of y, summed as a fixed tree.
template <int DIM> __device__ __forceinline__ void tail_cells(const TailLevel &L, const TailView<DIM> &V, const double *x, double *y)
{
  constexpr int NP = 1 << DIM;
  for (int c = threadIdx.x; c < L.n_cells; c += blockDim.x) {
    const TailIdx<DIM> nd = tail_load_idx<DIM>(V.nodes + c * NP);
    double p[NP];
#pragma unroll
    for (int a = 0; a < NP; ++a) p[a] = x[nd.get(a)];
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
    const double coef = V.coef[c];
#pragma unroll
    for (int a = 0; a < NP; ++a) y[c * NP + a] = coef * p[a];
  }
  if (threadIdx.x == blockDim.x - 1) y[L.n_cells * NP] = 0.0;      // the zero slot the padded adjacency entries point to
}
template <int DIM> __device__ __forceinline__ double tail_node(const TailView<DIM> &V, const double *y, int i)
{
  constexpr int NP = 1 << DIM;
  const TailIdx<DIM> ad = tail_load_idx<DIM>(V.adj + i * NP);
  double v[NP];
#pragma unroll
  for (int s = 0; s < NP; ++s) v[s] = y[ad.get(s)];
#pragma unroll
  for (int w = 1; w < NP; w *= 2)
#pragma unroll
    for (int s = 0; s < NP; s += 2 * w) v[s] += v[s + w];
  return v[0];
}

// one Chebyshev sweep: from_zero -> x = p(A) b from a zero start (degree - 1 operator applications), else one more sweep
// from the current iterate (degree applications).  The vector update of an iteration is fused into the node pass of the
// operator.  Returns the array that holds the result.
template <int DIM> __device__ double *tail_cheb(const TailLevel &L, const TailView<DIM> &V, double *y, double *x, double *xo, bool from_zero)
{
  const int n = L.n_dofs;
  bool have_old;
  if (from_zero) {
    const double it = 1.0 / L.theta;
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = it * V.dinv[i] * V.b[i];
    have_old = false;
  } else {
    tail_cells<DIM>(L, V, x, y);
    __syncthreads();
    const double f2 = 1.0 / L.theta;
    for (int i = threadIdx.x; i < n; i += blockDim.x) xo[i] = x[i] + f2 * V.dinv[i] * (V.b[i] - tail_node<DIM>(V, y, i));
    double *s = x; x = xo; xo = s;
    have_old = true;
  }
  __syncthreads();
  if (L.degree < 2 || fabs(L.delta) < 1e-40) return x;
  double rhok = L.delta / L.theta;
  const double sigma = L.theta / L.delta;
  for (int k = 0; k < L.degree - 1; ++k) {
    const double rhokp = 1.0 / (2.0 * sigma - rhok);
    const double f1 = rhokp * rhok, f2 = 2.0 * rhokp / L.delta;
    rhok = rhokp;
    tail_cells<DIM>(L, V, x, y);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double xi = x[i];
      double v = xi + f2 * V.dinv[i] * (V.b[i] - tail_node<DIM>(V, y, i));
      v += have_old ? f1 * (xi - xo[i]) : f1 * xi;
      xo[i] = v;
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
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int NPATCH = DIM == 3 ? 27 : 9, NC = 1 << DIM, NP = NC;
  const int top = P.n_levels - 1;
  double *xs[TAIL_MAX_LEVELS];          // where the current iterate of every level lives
  double *y = reinterpret_cast<double *>(smraw + P.scratch_off);      // [n_cells][NP] of the largest level: per-cell results
  // stage the static data of every level and build the node -> cells adjacency (padded with the zero slot y[n_cells * NP] of the
  // level, sorted: fixed summation order); the scratch y doubles as the fill counters while the lists are built
  int *cnt = reinterpret_cast<int *>(y);
  for (int l = 0; l <= top; ++l) {
    const TailLevel &L = P.lv[l];
    TailView<DIM> V(smraw, L);
    for (int i = threadIdx.x; i < L.n_dofs; i += blockDim.x) V.dinv[i] = L.dinv[i];
    for (int c = threadIdx.x; c < L.n_cells; c += blockDim.x) V.coef[c] = L.scale / L.eta[c];
    for (int w = threadIdx.x; w < L.n_cells * NP; w += blockDim.x) V.nodes[w] = (unsigned short)(L.cell_nodes[w] & NODE_IDX_MASK);
    for (int w = threadIdx.x; w < L.n_dofs * NP; w += blockDim.x) V.adj[w] = (unsigned short)(L.n_cells * NP);
  }
  {
    const TailLevel &L = P.lv[top];
    TailView<DIM> V(smraw, L);
    for (int i = threadIdx.x; i < L.n_dofs; i += blockDim.x) V.b[i] = P.defect[i];
  }
  for (int l = 0; l <= top; ++l) {
    const TailLevel &L = P.lv[l];
    TailView<DIM> V(smraw, L);
    for (int i = threadIdx.x; i < L.n_dofs; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (int w = threadIdx.x; w < L.n_cells * NP; w += blockDim.x) {
      const int node = V.nodes[w];
      const int slot = atomicAdd(&cnt[node], 1);
      V.adj[node * NP + slot] = (unsigned short)w;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L.n_dofs; i += blockDim.x) {
      unsigned short *a = V.adj + i * NP;
      for (int p = 1; p < NP; ++p) { const unsigned short key = a[p]; int q = p - 1; while (q >= 0 && a[q] > key) { a[q + 1] = a[q]; --q; } a[q + 1] = key; }
    }
    __syncthreads();
  }

  for (int l = top; l >= 0; --l) {
    const TailLevel &L = P.lv[l];
    TailView<DIM> V(smraw, L);
    // pre-smoothing (on level 0: the coarse solver): SmootherRelaxation::apply = vmult + (steps - 1) x step
    double *cur = tail_cheb<DIM>(L, V, y, V.x, V.xo, true);
    for (int s = 1; s < P.steps; ++s) cur = tail_cheb<DIM>(L, V, y, cur, cur == V.x ? V.xo : V.x, false);
    xs[l] = cur;
    if (l == 0) break;
    // residual t = b - A x, restricted into the zeroed defect of the coarser level
    const TailLevel &C = P.lv[l - 1];
    TailView<DIM> VC(smraw, C);
    for (int i = threadIdx.x; i < C.n_dofs; i += blockDim.x) VC.b[i] = 0.0;
    tail_cells<DIM>(L, V, cur, y);
    __syncthreads();
    for (int i = threadIdx.x; i < L.n_dofs; i += blockDim.x) V.t[i] = V.b[i] - tail_node<DIM>(V, y, i);
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
        if (wt != 0.0) acc = fma(wt, V.t[pe & NODE_IDX_MASK], acc);
      }
      if (acc != 0.0) atomicAdd(VC.b + VC.nodes[par * NC + j], acc);
    }
    __syncthreads();
  }
  for (int l = 1; l <= top; ++l) {
    const TailLevel &L = P.lv[l], &C = P.lv[l - 1];
    TailView<DIM> V(smraw, L), VC(smraw, C);
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
        if (wt != 0.0) acc = fma(wt, xc[VC.nodes[par * NC + j]], acc);
      }
      cur[pe & NODE_IDX_MASK] += acc;
    }
    __syncthreads();
    for (int s = 0; s < P.steps; ++s) cur = tail_cheb<DIM>(L, V, y, cur, cur == V.x ? V.xo : V.x, false);
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
  size_t bytes = 0;
  int top = -1;
  for (int l = 0; l < ctx->cfg.n_levels && l < TAIL_MAX_LEVELS; ++l) {
    const LevelDev &L = ctx->levels[l];
    if (L.n_eta != 1 || L.n_constrained_p || L.geo || L.edge_p.n || L.n_cells > ctx->tail_max_cells || L.n_p > 60000 || L.n_cells * 8 > 60000) break;
    if (L.n_p > L.n_cells * (1 << L.dim) * 2) break;          // (the scratch of the level also holds n_p fill counters)
    if (l > 0 && !ctx->transfers[l].patch_p) break;
    bytes += L.dim == 3 ? tail_level_bytes<3>((int)L.n_cells, (int)L.n_p) : tail_level_bytes<2>((int)L.n_cells, (int)L.n_p);
    if (bytes + ((size_t)L.n_cells * (1 << L.dim) + 2) * 8 > (size_t)max_smem - 1024) break;      // + the per-cell scratch of the largest level
    top = l;
  }
  return top;
}

// defect = mgv_S[top].b, solution -> mgv_S[top].x (and .res)
int launch_schur_tail(smf_ctx *ctx, int top)
{
  TailParams P{};
  P.n_levels = top + 1; P.steps = 4;
  size_t off = 0;
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
    off += L.dim == 3 ? tail_level_bytes<3>((int)L.n_cells, (int)L.n_p) : tail_level_bytes<2>((int)L.n_cells, (int)L.n_p);
    if (l > 0) { T.patch = ctx->transfers[l].patch_p; T.n_parents = (int)ctx->transfers[l].n_parents; T.coarse_nodes = ctx->levels[l - 1].cell_nodes_p; }
  }
  MGVectors &V = ctx->mgv_S[top];
  P.defect = V.b; P.solution = V.x;
  P.scratch_off = (int)off;
  const size_t smem = off + ((size_t)ctx->levels[top].n_cells * (1 << ctx->cfg.dim) + 2) * sizeof(double);
  static size_t configured2[MAX_DEVICES] = {0}, configured3[MAX_DEVICES] = {0};
  size_t &configured = (ctx->cfg.dim == 3 ? configured3 : configured2)[ctx->cfg.device % MAX_DEVICES];
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
