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

// StokesOperator::local_apply_boundary_face (matrix_free_operators.cc:563-602) on Cartesian cells: one thread per face and
// face node I (3^(dim-1) nodes, all components): dst_I -= n_c int_face (g . u) phi_I dA with the 3^(dim-1)-point Gauss rule.
// Source values are read like in the cell loop (constrained DoFs read 0 unless plain), constrained rows receive nothing.
template <int DIM>
__global__ void k_free_surface_faces(const uint32_t *cells, const int32_t *face_no, const double *table, const uint32_t *cell_nodes_u,
                                     const double *src, double *dst, int64_t n_faces, double hx, double hy, double hz, int plain)
{
  constexpr int NF = DIM == 3 ? 9 : 3, NN = DIM == 3 ? 27 : 9;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_faces * NF) return;
  const int64_t f = t / NF; const int I = (int)(t % NF);
  const int fn = face_no[f], d = fn / 2, side = fn % 2;
  const double h[3] = {hx, hy, hz};
  int tang[2] = {0, 0}, nt = 0;
  double area = 1.0;
  for (int e = 0; e < DIM; ++e) if (e != d) { tang[nt++] = e; area *= h[e]; }
  const double s5 = 0.7745966692414834, gx[3] = {0.5 - 0.5 * s5, 0.5, 0.5 + 0.5 * s5}, gwt[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
  auto shape = [](int i, double x) { return i == 0 ? (1 - x) * (1 - 2 * x) : (i == 1 ? 4 * x * (1 - x) : x * (2 * x - 1)); };
  auto local_node = [&](int a) {       // lexicographic cell index of face node a (tangential digits, lower direction fastest)
    int idx[3] = {0, 0, 0};
    idx[d] = side == 0 ? 0 : 2;
    idx[tang[0]] = a % 3;
    if (DIM == 3) idx[tang[1]] = a / 3;
    return idx[0] + 3 * idx[1] + 9 * idx[2];
  };
  const uint32_t *cn = cell_nodes_u + (size_t)cells[f] * NN;
  // values of u at the face nodes
  double u[NF][DIM];
  for (int a = 0; a < NF; ++a) {
    const uint32_t e = cn[local_node(a)];
    for (int c = 0; c < DIM; ++c) {
      const bool cons = (e >> (NODE_FLAG_SHIFT + c)) & 1u;
      u[a][c] = (cons && !plain) ? 0.0 : src[(e & NODE_IDX_MASK) + c];
    }
  }
  double acc = 0.0;
  for (int q = 0; q < NF; ++q) {
    const int q0 = q % 3, q1 = DIM == 3 ? q / 3 : 0;
    double gu = 0.0;
    for (int a = 0; a < NF; ++a) {
      const double ph = shape(a % 3, gx[q0]) * (DIM == 3 ? shape(a / 3, gx[q1]) : 1.0);
      for (int c = 0; c < DIM; ++c) gu = fma(table[((size_t)f * NF + q) * DIM + c] * ph, u[a][c], gu);
    }
    const double w = gwt[q0] * (DIM == 3 ? gwt[q1] : 1.0) * area;
    acc = fma(gu * w, shape(I % 3, gx[q0]) * (DIM == 3 ? shape(I / 3, gx[q1]) : 1.0), acc);
  }
  const uint32_t e = cn[local_node(I)];
  if (!((e >> (NODE_FLAG_SHIFT + d)) & 1u)) atomicAdd(dst + (e & NODE_IDX_MASK) + d, side == 0 ? acc : -acc);     // -(g . u) n, n = -+ e_d
}

int free_surface_faces_apply(smf_ctx *ctx, const LevelDev &L, double *dst_u, const double *src_u, bool plain)
{
  const FaceSet &F = L.fs_faces;
  if (F.n == 0) return SMF_OK;
  const int NF = L.dim == 3 ? 9 : 3;
  const unsigned g = (unsigned)((F.n * NF + 127) / 128);
  if (L.dim == 3) k_free_surface_faces<3><<<g, 128, 0, ctx->stream>>>(F.cells, F.face_no, F.table, L.cell_nodes_u, src_u, dst_u, F.n, L.h[0], L.h[1], L.h[2], plain ? 1 : 0);
  else k_free_surface_faces<2><<<g, 128, 0, ctx->stream>>>(F.cells, F.face_no, F.table, L.cell_nodes_u, src_u, dst_u, F.n, L.h[0], L.h[1], 1.0, plain ? 1 : 0);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
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
static void free_faces(FaceSet &F) { cudaFree(F.cells); cudaFree(F.face_no); cudaFree(F.table); F = FaceSet(); }
void free_adaptive(LevelDev &L)
{ free_lines(L.hang_u); free_lines(L.hang_p); free_edge(L.edge_u); free_edge(L.edge_p); free_pairs(L.copy_u); free_pairs(L.copy_p); free_faces(L.fs_faces); }

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
  lattice_free(ctx);
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

int smf_set_free_surface_faces(smf_ctx *ctx, int64_t n_faces, const uint32_t *cells, const int32_t *face_numbers, const double *table)
{
  if (!ctx || n_faces < 0) return SMF_ERR_INVALID;
  LevelDev &L = ctx->lev(SMF_ACTIVE);
  if (L.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set the active level mesh first");
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  free_faces(L.fs_faces);
  if (n_faces == 0) return SMF_OK;
  if (!cells || !face_numbers || !table) return ctx->fail(SMF_ERR_INVALID, "smf_set_free_surface_faces: missing array");
  if (L.geo) return ctx->fail(SMF_ERR_UNSUPPORTED, "free-surface stabilisation faces on curved cells (geometry = 1) are not implemented");
  for (int64_t i = 0; i < n_faces; ++i)
    if (cells[i] >= L.n_cells || face_numbers[i] < 0 || face_numbers[i] >= 2 * L.dim) return ctx->fail(SMF_ERR_INVALID, "smf_set_free_surface_faces: face out of range");
  const int NF = L.dim == 3 ? 9 : 3;
  int e;
  if ((e = to_device(ctx, &L.fs_faces.cells, cells, n_faces))) return e;
  if ((e = to_device(ctx, &L.fs_faces.face_no, face_numbers, n_faces))) return e;
  if ((e = to_device(ctx, &L.fs_faces.table, table, n_faces * NF * L.dim))) return e;
  L.fs_faces.n = n_faces;
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
  lattice_free(ctx);
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
  lattice_free(ctx);
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
