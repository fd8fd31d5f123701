// Adaptively refined meshes on the device: hanging-node constraints of the active level and the refinement-edge handling
// of local smoothing.
//
// Reference: /root/reference/source/simulator/solver/stokes_matrix_free_local_smoothing.cc:1508, :1580 (hanging-node
// constraints), :1733-1773 (MGConstrainedDoFs refinement-edge indices), :879-911 (MGInterfaceOperator, set_edge_matrices) and
// the deal.II semantics of MatrixFreeOperators::Base: read_dof_values resolves the constraint lines,
// distribute_local_to_global applies their transpose, constrained rows get dst = src; with edge_constrained_indices the
// level operator is A_l = (I - E) A_bc (I - E) + E and vmult_interface_down / _up are E A_bc (I - E) / (I - E) A_bc E.
// (tests/test_gpu_adaptive.py checks every piece against the CPU restatement of these semantics.)
//
// The cell kernels stay untouched: a hanging node is an ordinary node of its cell.  Around the cell loop
//   hanging_pre   src[k] <- sum_j w_j src[m_j]   (the old entry is parked; a constrained entry of src is never read otherwise)
//   hanging_post  dst[m_j] += w_j dst[k], dst[k] <- parked src[k] (unit diagonal), src[k] restored
// which is C^T A C + I_c term by term (the cell-wise resolution of deal.II adds the same products in another order).
// The level operators park and zero the refinement-edge entries the same way; the interface products re-use the cell loop
// of the level: (A_l + down) x = A_bc (I - E) x + E x comes out of ONE loop (edge_post with add), up x = (I - E) A_bc (E x).
#include "smf_internal.h"

namespace smf {

__global__ void k_hang_pre(double *src, const uint32_t *dof, const uint32_t *start, const uint32_t *master, const double *w, double *save, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = dof[i];
  save[i] = src[k];
  double s = 0.0;
  for (uint32_t j = start[i]; j < start[i + 1]; ++j) s = fma(w[j], src[master[j]], s);     // (masters are never constrained: no race)
  src[k] = s;
}
__global__ void k_hang_post(double *dst, double *src, const uint32_t *dof, const uint32_t *start, const uint32_t *master, const double *w,
                            const double *save, int64_t n, int unit_diagonal, int restore)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = dof[i];
  const double v = dst[k];
  for (uint32_t j = start[i]; j < start[i + 1]; ++j) atomicAdd(dst + master[j], w[j] * v);
  dst[k] = unit_diagonal ? save[i] : 0.0;
  if (restore) src[k] = save[i];
}
__global__ void k_hang_restore(double *src, const uint32_t *dof, const double *save, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) src[dof[i]] = save[i];
}
__global__ void k_edge_pre(double *src, const uint32_t *idx, double *save, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  save[i] = src[idx[i]];
  src[idx[i]] = 0.0;
}
__global__ void k_edge_post(double *dst, double *src, const uint32_t *idx, const double *save, int64_t n, int add)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t e = idx[i];
  dst[e] = add ? dst[e] + save[i] : save[i];
  src[e] = save[i];
}
__global__ void k_take_entries(double *y, const double *x, const uint32_t *idx, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[idx[i]] = x[idx[i]];
}
__global__ void k_add_entries(double *b, const double *y, const uint32_t *idx, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) b[idx[i]] += y[idx[i]];
}
__global__ void k_copy_pairs(double *level_vec, double *active_vec, const uint32_t *ia, const uint32_t *il, int64_t n, int to_level)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (to_level) level_vec[il[i]] = active_vec[ia[i]];
  else active_vec[ia[i]] = level_vec[il[i]];
}

static inline unsigned blocks(int64_t n) { return (unsigned)((n + 255) / 256); }

int hanging_pre(smf_ctx *ctx, ConstraintLines &H, const double *src)
{
  if (H.n == 0) return SMF_OK;
  k_hang_pre<<<blocks(H.n), 256, 0, ctx->stream>>>(const_cast<double *>(src), H.dof, H.start, H.master, H.weight, H.save, H.n);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}
int hanging_post(smf_ctx *ctx, ConstraintLines &H, double *dst, const double *src, bool unit_diagonal)
{
  if (H.n == 0) return SMF_OK;
  k_hang_post<<<blocks(H.n), 256, 0, ctx->stream>>>(dst, const_cast<double *>(src), H.dof, H.start, H.master, H.weight, H.save, H.n,
                                                    unit_diagonal ? 1 : 0, src ? 1 : 0);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}
int hanging_restore(smf_ctx *ctx, ConstraintLines &H, const double *src)
{
  if (H.n == 0) return SMF_OK;
  k_hang_restore<<<blocks(H.n), 256, 0, ctx->stream>>>(const_cast<double *>(src), H.dof, H.save, H.n);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}
int edge_pre(smf_ctx *ctx, EdgeSet &E, const double *src)
{
  if (E.n == 0) return SMF_OK;
  k_edge_pre<<<blocks(E.n), 256, 0, ctx->stream>>>(const_cast<double *>(src), E.idx, E.save, E.n);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}
int edge_post(smf_ctx *ctx, EdgeSet &E, double *dst, const double *src, bool add)
{
  if (E.n == 0) return SMF_OK;
  k_edge_post<<<blocks(E.n), 256, 0, ctx->stream>>>(dst, const_cast<double *>(src), E.idx, E.save, E.n, add ? 1 : 0);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}
int edge_extract(smf_ctx *ctx, const EdgeSet &E, double *xe, const double *x, int64_t n)
{
  SMF_CUDA(cudaMemsetAsync(xe, 0, sizeof(double) * n, ctx->stream));
  if (E.n == 0) return SMF_OK;
  k_take_entries<<<blocks(E.n), 256, 0, ctx->stream>>>(xe, x, E.idx, E.n);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}
int edge_subtract_off(smf_ctx *ctx, const EdgeSet &E, double *b, const double *y, int64_t n)
{
  if (int e = vec_axpy(ctx, b, -1.0, y, n)) return e;          // b -= y everywhere ...
  if (E.n == 0) return SMF_OK;
  k_add_entries<<<blocks(E.n), 256, 0, ctx->stream>>>(b, y, E.idx, E.n);      // ... and back on the edge rows
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}
int copy_pairs(smf_ctx *ctx, const CopyPairs &C, double *level_vec, double *active_vec, bool to_level)
{
  if (C.n == 0) return SMF_OK;
  k_copy_pairs<<<blocks(C.n), 256, 0, ctx->stream>>>(level_vec, active_vec, C.active, C.level, C.n, to_level ? 1 : 0);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}

static void free_lines(ConstraintLines &H)
{ cudaFree(H.dof); cudaFree(H.start); cudaFree(H.master); cudaFree(H.weight); cudaFree(H.save); H = ConstraintLines(); }
static void free_edge(EdgeSet &E) { cudaFree(E.idx); cudaFree(E.save); E = EdgeSet(); }
static void free_pairs(CopyPairs &C) { cudaFree(C.active); cudaFree(C.level); C = CopyPairs(); }
void free_adaptive(LevelDev &L)
{ free_lines(L.hang_u); free_lines(L.hang_p); free_edge(L.edge_u); free_edge(L.edge_p); free_pairs(L.copy_u); free_pairs(L.copy_p); }

template <class T> static int to_device(smf_ctx *ctx, T **dev, const T *host, int64_t n)
{
  SMF_CUDA(cudaMalloc((void **)dev, sizeof(T) * (size_t)(n > 0 ? n : 1)));
  if (n > 0) SMF_CUDA(cudaMemcpy(*dev, host, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice));
  ctx->bytes += (int64_t)sizeof(T) * n;
  return SMF_OK;
}
}  // namespace smf

using namespace smf;

extern "C" {

int smf_set_constraints(smf_ctx *ctx, int level, int family, int64_t n_lines, const uint32_t *constrained_dof,
                        const uint32_t *row_start, const uint32_t *master_dof, const double *weight)
{
  if (!ctx || (family != 0 && family != 1) || n_lines < 0) return SMF_ERR_INVALID;
  if (level < SMF_ACTIVE || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  LevelDev &L = ctx->lev(level);
  if (L.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set the level mesh first");
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  ConstraintLines &H = family == 0 ? L.hang_u : L.hang_p;
  free_lines(H);
  if (n_lines == 0) return SMF_OK;
  if (!constrained_dof || !row_start || !master_dof || !weight) return ctx->fail(SMF_ERR_INVALID, "smf_set_constraints: missing array");
  const int64_t n_dofs = family == 0 ? L.n_u : L.n_p, nnz = row_start[n_lines];
  const std::vector<uint8_t> &cons = family == 0 ? L.h_cons_u : L.h_cons_p;
  std::vector<uint8_t> is_line((size_t)n_dofs, 0);
  for (int64_t i = 0; i < n_lines; ++i) {
    if (constrained_dof[i] >= n_dofs || row_start[i + 1] < row_start[i]) return ctx->fail(SMF_ERR_INVALID, "smf_set_constraints: line out of range");
    if (!cons.empty() && cons[constrained_dof[i]]) return ctx->fail(SMF_ERR_INVALID, "smf_set_constraints: a DoF of smf_level_mesh::constrained_* cannot carry a line as well");
    is_line[constrained_dof[i]] = 1;
  }
  for (int64_t j = 0; j < nnz; ++j)
    if (master_dof[j] >= n_dofs || is_line[master_dof[j]]) return ctx->fail(SMF_ERR_INVALID, "smf_set_constraints: lines must be closed (masters unconstrained)");
  int e;
  if ((e = to_device(ctx, &H.dof, constrained_dof, n_lines))) return e;
  if ((e = to_device(ctx, &H.start, row_start, n_lines + 1))) return e;
  if ((e = to_device(ctx, &H.master, master_dof, nnz))) return e;
  if ((e = to_device(ctx, &H.weight, weight, nnz))) return e;
  SMF_CUDA(cudaMalloc(&H.save, sizeof(double) * (size_t)n_lines));
  H.n = n_lines;
  L.has_diag = false;
  ctx->smoothers_ready = false;
  return SMF_OK;
}

int smf_set_refinement_edges(smf_ctx *ctx, int level, int64_t n_u, const uint32_t *edge_u, int64_t n_p, const uint32_t *edge_p)
{
  if (!ctx || n_u < 0 || n_p < 0) return SMF_ERR_INVALID;
  if (level < 0 || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  LevelDev &L = ctx->levels[level];
  if (L.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set the level mesh first");
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  free_edge(L.edge_u); free_edge(L.edge_p);
  for (int64_t i = 0; i < n_u; ++i) if (edge_u[i] >= L.n_u) return ctx->fail(SMF_ERR_INVALID, "refinement-edge index out of range");
  for (int64_t i = 0; i < n_p; ++i) if (edge_p[i] >= L.n_p) return ctx->fail(SMF_ERR_INVALID, "refinement-edge index out of range");
  int e;
  if (n_u) {
    if ((e = to_device(ctx, &L.edge_u.idx, edge_u, n_u))) return e;
    SMF_CUDA(cudaMalloc(&L.edge_u.save, sizeof(double) * (size_t)n_u));
    L.edge_u.n = n_u;
  }
  if (n_p) {
    if ((e = to_device(ctx, &L.edge_p.idx, edge_p, n_p))) return e;
    SMF_CUDA(cudaMalloc(&L.edge_p.save, sizeof(double) * (size_t)n_p));
    L.edge_p.n = n_p;
  }
  L.has_diag = false;
  ctx->smoothers_ready = false;
  return SMF_OK;
}

int smf_set_level_copy_indices(smf_ctx *ctx, int level, int family, int64_t n, const uint32_t *active_idx, const uint32_t *level_idx)
{
  if (!ctx || (family != 0 && family != 1) || n < 0) return SMF_ERR_INVALID;
  if (level < 0 || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  LevelDev &L = ctx->levels[level], &A = ctx->lev(SMF_ACTIVE);
  if (L.n_cells == 0 || A.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set the level meshes first");
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  CopyPairs &C = family == 0 ? L.copy_u : L.copy_p;
  free_pairs(C);
  const int64_t na = family == 0 ? A.n_u : A.n_p, nl = family == 0 ? L.n_u : L.n_p;
  for (int64_t i = 0; i < n; ++i)
    if (active_idx[i] >= na || level_idx[i] >= nl) return ctx->fail(SMF_ERR_INVALID, "copy index out of range");
  int e;
  if ((e = to_device(ctx, &C.active, active_idx, n))) return e;
  if ((e = to_device(ctx, &C.level, level_idx, n))) return e;
  C.n = n;
  ctx->adaptive = true;
  ctx->graph_A_stale = ctx->graph_S_stale = true;
  return SMF_OK;
}

}  // extern "C"
