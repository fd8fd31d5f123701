// Context life cycle and set-up side of the C ABI: what StokesMatrixFreeHandler::setup_dofs() and
// evaluate_material_model() hand to the operators
// (/root/reference/source/simulator/solver/stokes_matrix_free_local_smoothing.cc:1450-1851, :173-294).
// The host code here turns the MatrixFree-style cell->DoF maps into the batched layout the kernels use
// (DESIGN.md "Data layout"): per 64-cell batch a list of unique nodes (private nodes first) and 16-bit
// local slots per cell, lane-transposed so a warp reads them coalesced.
#include "smf_internal.h"
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <chrono>

using namespace smf;

namespace smf {
int launch_diag(smf_ctx *ctx, LevelDev &L);

template <class T> static int upload(smf_ctx *ctx, T **dev, const std::vector<T> &host, int64_t *bytes)
{
  const size_t nb = std::max<size_t>(host.size(), 1) * sizeof(T);
  SMF_CUDA(cudaMalloc((void **)dev, nb));
  if (!host.empty()) SMF_CUDA(cudaMemcpyAsync(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  *bytes += (int64_t)nb;
  ctx->bytes += (int64_t)nb;
  return SMF_OK;
}

static void free_level(LevelDev &L)
{
  cudaFree(L.batch_info); cudaFree(L.node_list_u); cudaFree(L.node_slot_u); cudaFree(L.sb_lists); cudaFree(L.node_list_p); cudaFree(L.cell_loc_u); cudaFree(L.cell_loc_p);
  cudaFree(L.cell_nodes_u); cudaFree(L.cell_nodes_p); cudaFree(L.constrained_u); cudaFree(L.constrained_p);
  cudaFree(L.eta); cudaFree(L.dinv_A); cudaFree(L.dinv_M); cudaFree(L.geo); cudaFree(L.cell_nodes_u_t); cudaFree(L.cell_nodes_p_t);
  cudaFree(L.nd.eps); cudaFree(L.nd.deta); cudaFree(L.nd.detadp); cudaFree(L.nd.dil_lhs); cudaFree(L.nd.ddil); cudaFree(L.nd.ddildp);
  free_exchange(L.ex_u); free_exchange(L.ex_p);
  free_adaptive(L);
  vlat_free(L);
  L = LevelDev();
}

// Build the batched staging layout for one DoF family.
//   nodes: [n_cells][npc] node entries (index | flags)
static void build_batches(const std::vector<uint32_t> &nodes, int64_t n_cells, int npc, int64_t n_index_space,
                          const int *round_of, int n_rounds,
                          std::vector<uint32_t> &node_list, std::vector<uint16_t> &cell_loc,
                          std::vector<int> &n_nodes, std::vector<int> &n_excl, int &max_nodes, bool &safe)
{
  const int64_t nb = (n_cells + BATCH_CELLS - 1) / BATCH_CELLS;
  std::vector<uint8_t> touch((size_t)n_index_space, 0);
  std::vector<std::vector<uint32_t>> uniq((size_t)nb);
  for (int64_t b = 0; b < nb; ++b) {
    const int64_t c0 = b * BATCH_CELLS, c1 = std::min<int64_t>(n_cells, c0 + BATCH_CELLS);
    std::vector<uint32_t> &u = uniq[(size_t)b];
    u.assign(nodes.begin() + c0 * npc, nodes.begin() + c1 * npc);
    std::sort(u.begin(), u.end(), [](uint32_t a, uint32_t b2) { return (a & NODE_IDX_MASK) < (b2 & NODE_IDX_MASK); });
    u.erase(std::unique(u.begin(), u.end(), [](uint32_t a, uint32_t b2) { return (a & NODE_IDX_MASK) == (b2 & NODE_IDX_MASK); }), u.end());
    for (uint32_t e : u) {
      uint8_t &t = touch[e & NODE_IDX_MASK];
      if (t < 2) ++t;
    }
  }
  max_nodes = 1;
  for (auto &u : uniq) max_nodes = std::max<int>(max_nodes, (int)u.size());
  max_nodes = (max_nodes + 3) & ~3;
  node_list.assign((size_t)nb * max_nodes, 0);
  cell_loc.assign((size_t)nb * KSUB * npc * SUB, 0);
  n_nodes.assign((size_t)nb, 0);
  n_excl.assign((size_t)nb, 0);
  safe = true;
  std::vector<int> stamp_k;
  std::vector<uint32_t> stamp_mask;
  for (int64_t b = 0; b < nb; ++b) {
    std::vector<uint32_t> &u = uniq[(size_t)b];
    // private nodes first (each group keeps ascending index order -> coalesced gathers and flushes)
    std::stable_partition(u.begin(), u.end(), [&](uint32_t e) { return touch[e & NODE_IDX_MASK] == 1; });
    int ne = 0;
    for (uint32_t e : u) ne += touch[e & NODE_IDX_MASK] == 1;
    n_nodes[(size_t)b] = (int)u.size();
    n_excl[(size_t)b] = ne;
    std::copy(u.begin(), u.end(), node_list.begin() + b * max_nodes);
    // slot lookup: sorted copy of (index, slot)
    std::vector<std::pair<uint32_t, int>> look(u.size());
    for (size_t s = 0; s < u.size(); ++s) look[s] = {u[s] & NODE_IDX_MASK, (int)s};
    std::sort(look.begin(), look.end());
    const int64_t c0 = b * BATCH_CELLS, c1 = std::min<int64_t>(n_cells, c0 + BATCH_CELLS);
    stamp_k.assign(u.size(), -1);
    stamp_mask.assign(u.size(), 0u);
    for (int64_t cell = c0; cell < c1; ++cell) {
      const int k = (int)((cell - c0) / SUB), lane = (int)((cell - c0) % SUB);
      int slots[27];
      for (int n = 0; n < npc; ++n) {
        const uint32_t idx = nodes[cell * npc + n] & NODE_IDX_MASK;
        auto it = std::lower_bound(look.begin(), look.end(), std::make_pair(idx, 0));
        const int slot = slots[n] = it->second;
        cell_loc[(((size_t)b * KSUB + k) * npc + n) * SUB + lane] = (uint16_t)slot;
        // inside one (sub-batch, round) every (cell, node) must hit a distinct slot
        if (stamp_k[slot] != k) { stamp_k[slot] = k; stamp_mask[slot] = 0u; }
        if (stamp_mask[slot] & (1u << round_of[n])) safe = false;
      }
      for (int n = 0; n < npc; ++n) {
        if (stamp_mask[slots[n]] & (1u << round_of[n])) {
          // the same cell hits one slot twice in one round (degenerate cell)
          for (int n2 = 0; n2 < n; ++n2) if (slots[n2] == slots[n] && round_of[n2] == round_of[n]) safe = false;
        }
        stamp_mask[slots[n]] |= 1u << round_of[n];
      }
    }
  }
}

// Structured-brick staging layout of the velocity nodes (3D): succeeds when every batch is a full 4x4x2 Morton brick
// whose cell maps are consistent with a 9x9x5 lattice.  node_list[b][sb_slot] = node entry | NODE_SHARED_BIT.
// Pressure nodes of structured bricks in lattice order (5 x 5 x 3): entry | NODE_SHARED_BIT, bit 31 = constrained.
static bool build_structured3_p(const std::vector<uint32_t> &nodes_p, int64_t n_cells, int64_t n_index_space, std::vector<uint32_t> &plist)
{
  const int64_t nb = n_cells / BATCH_CELLS;
  constexpr int NL = 5 * 5 * 3;
  plist.assign((size_t)nb * SB_PLIST, NODE_INVALID);
  std::vector<uint8_t> touch((size_t)n_index_space, 0);
  std::vector<int32_t> seen_at((size_t)n_index_space, -1);
  std::vector<uint32_t> batch_nodes;
  for (int64_t b = 0; b < nb; ++b) {
    uint32_t *L = plist.data() + (size_t)b * SB_PLIST;
    batch_nodes.clear();
    bool ok = true;
    for (int l = 0; l < BATCH_CELLS && ok; ++l) {
      const uint32_t *cn = nodes_p.data() + ((size_t)b * BATCH_CELLS + l) * 8;
      const int x = sb_x(l), y = sb_y(l), z = sb_z(l);
      for (int n = 0; n < 8; ++n) {
        const int id = (x + (n & 1)) + 5 * ((y + ((n >> 1) & 1)) + 5 * (z + (n >> 2)));
        const uint32_t idx = cn[n] & NODE_IDX_MASK;
        if (L[id] == NODE_INVALID) {
          if (seen_at[idx] >= 0) { ok = false; break; }
          L[id] = cn[n]; seen_at[idx] = id; batch_nodes.push_back(idx);
        } else if ((L[id] & NODE_IDX_MASK) != idx) { ok = false; break; }
      }
    }
    for (uint32_t idx : batch_nodes) { seen_at[idx] = -1; if (touch[idx] < 2) ++touch[idx]; }
    if (!ok) return false;
  }
  for (int64_t b = 0; b < nb; ++b)
    for (int i = 0; i < NL; ++i) {
      uint32_t &e = plist[(size_t)b * SB_PLIST + i];
      if (touch[e & NODE_IDX_MASK] > 1) e |= NODE_SHARED_BIT;
    }
  return true;
}

static bool build_structured3(const std::vector<uint32_t> &nodes, int64_t n_cells, int64_t n_index_space,
                              std::vector<uint32_t> &node_list, std::vector<uint16_t> &node_slot)
{
  if (n_cells % BATCH_CELLS != 0) return false;
  const int64_t nb = n_cells / BATCH_CELLS;
  constexpr int NL = 9 * 9 * 5;
  std::vector<uint32_t> lat((size_t)nb * NL);
  std::vector<uint8_t> touch((size_t)n_index_space, 0);
  std::vector<int32_t> seen_at((size_t)n_index_space, -1);     // lattice id of a node inside the current batch
  std::vector<uint32_t> batch_nodes;
  for (int64_t b = 0; b < nb; ++b) {
    uint32_t *L = lat.data() + (size_t)b * NL;
    std::fill(L, L + NL, NODE_INVALID);
    batch_nodes.clear();
    bool ok = true;
    for (int l = 0; l < BATCH_CELLS && ok; ++l) {
      const uint32_t *cn = nodes.data() + ((size_t)b * BATCH_CELLS + l) * 27;
      const int x = sb_x(l), y = sb_y(l), z = sb_z(l);
      for (int n = 0; n < 27; ++n) {
        const int id = (2 * x + n % 3) + 9 * ((2 * y + (n / 3) % 3) + 9 * (2 * z + n / 9));
        const uint32_t idx = cn[n] & NODE_IDX_MASK;
        if (L[id] == NODE_INVALID) {
          if (seen_at[idx] >= 0) { ok = false; break; }       // the same node at two lattice points
          L[id] = cn[n]; seen_at[idx] = id; batch_nodes.push_back(idx);
        } else if ((L[id] & NODE_IDX_MASK) != idx) { ok = false; break; }
      }
    }
    for (uint32_t idx : batch_nodes) { seen_at[idx] = -1; if (touch[idx] < 2) ++touch[idx]; }
    if (!ok) return false;
  }
  // list entries: nodes private to the brick first, then the shared ones, each group in ascending DoF-index order
  // (coalesced global access; the flush is warp-uniform except at the group boundary), each with its staging slot.
  // Padding entries carry brick facts: [405] = number of private nodes, [406] = 1 if any DoF of the brick is constrained.
  node_list.assign((size_t)nb * SB_LIST, NODE_INVALID);
  node_slot.assign((size_t)nb * SB_LIST, 0);
  std::vector<std::pair<uint32_t, uint16_t>> ent(NL);
  for (int64_t b = 0; b < nb; ++b) {
    const uint32_t *L = lat.data() + (size_t)b * NL;
    int m = 0;
    uint32_t n_private = 0, any_cons = 0;
    for (int Z = 0; Z < 5; ++Z)
      for (int Y = 0; Y < 9; ++Y)
        for (int X = 0; X < 9; ++X) {
          const uint32_t e = L[X + 9 * (Y + 9 * Z)];
          const bool shared = touch[e & NODE_IDX_MASK] > 1;
          n_private += shared ? 0 : 1;
          if (e >> NODE_FLAG_SHIFT) any_cons = 1;
          ent[m++] = {e | (shared ? NODE_SHARED_BIT : 0u), (uint16_t)sb_slot(X, Y, Z)};
        }
    std::sort(ent.begin(), ent.end(), [](const std::pair<uint32_t, uint16_t> &a, const std::pair<uint32_t, uint16_t> &c) {
      const uint32_t sa = a.first & NODE_SHARED_BIT, sc = c.first & NODE_SHARED_BIT;
      if (sa != sc) return sa < sc;
      return (a.first & NODE_IDX_MASK) < (c.first & NODE_IDX_MASK); });
    for (int i = 0; i < NL; ++i) {
      node_list[(size_t)b * SB_LIST + i] = ent[i].first;
      node_slot[(size_t)b * SB_LIST + i] = ent[i].second;
    }
    node_list[(size_t)b * SB_LIST + 405] = n_private;
    node_list[(size_t)b * SB_LIST + 406] = any_cons;
  }
  return true;
}
}  // namespace smf

extern "C" void smf_comm_destroy_internal(smf_ctx *ctx);

extern "C" {

int smf_create(const smf_config *cfg, smf_ctx **out)
{
  if (!cfg || !out) return SMF_ERR_INVALID;
  *out = nullptr;
  if ((cfg->dim != 2 && cfg->dim != 3) || cfg->degree != 2 || cfg->n_levels < 1) return SMF_ERR_UNSUPPORTED;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device >= ndev) {
    cudaGetLastError();
    return SMF_ERR_CUDA;   // no CPU fallback: the product path needs a CUDA device
  }
  if (cudaSetDevice(cfg->device) != cudaSuccess) return SMF_ERR_CUDA;
  smf_ctx *ctx = new smf_ctx();
  ctx->cfg = *cfg;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SMF_ERR_CUDA; }
  ctx->own_stream = true;
  ctx->levels.resize(cfg->n_levels + 1);
  ctx->transfers.resize(cfg->n_levels);
  ctx->smoother_A.resize(cfg->n_levels);
  ctx->smoother_S.resize(cfg->n_levels);
  ctx->mgv_A.resize(cfg->n_levels);
  ctx->mgv_S.resize(cfg->n_levels);
  *out = ctx;
  return SMF_OK;
}

void smf_destroy(smf_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &L : ctx->levels) free_level(L);
  for (auto &T : ctx->transfers) { cudaFree(T.patch_u); cudaFree(T.patch_p); cudaFree(T.tb_u); cudaFree(T.children); }
  for (auto *mg : {&ctx->mgv_A, &ctx->mgv_S})
    for (auto &v : *mg) { cudaFree(v.x); cudaFree(v.b); cudaFree(v.t); cudaFree(v.xo); cudaFree(v.e); }
  for (double *p : ctx->krylov) cudaFree(p);
  cudaFree(ctx->copy_u); cudaFree(ctx->copy_p); cudaFree(ctx->scratch);
  cudaFree(ctx->io_buf); cudaFree(ctx->block_weights); cudaFree(ctx->pnorm_weights); cudaFree(ctx->cg_buf); cudaFree(ctx->bicg_buf); cudaFree(ctx->idr_normals);
  smf_comm_destroy_internal(ctx);
  cudaFree(ctx->tmp_u); cudaFree(ctx->tmp_u2);   // tmp_p / tmp_p2 point into the same allocations
  if (ctx->host_pinned) cudaFreeHost(ctx->host_pinned);
  for (auto &e : ctx->prof_live) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
  lattice_free(ctx);
  cudaFree(ctx->ad_in_u); cudaFree(ctx->ad_out_u); cudaFree(ctx->ad_in_p); cudaFree(ctx->ad_out_p);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->h2d_event) cudaEventDestroy(ctx->h2d_event);
  if (ctx->graph_A) cudaGraphExecDestroy((cudaGraphExec_t)ctx->graph_A);
  if (ctx->graph_S) cudaGraphExecDestroy((cudaGraphExec_t)ctx->graph_S);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  delete ctx;
}

const char *smf_last_error(const smf_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int smf_set_option(smf_ctx *ctx, int option, int value)
{
  if (!ctx) return SMF_ERR_INVALID;
  if (option == SMF_OPT_STRUCTURED_BRICKS) ctx->allow_structured = value ? 1 : 0;
  else if (option == SMF_OPT_CUDA_GRAPHS) { ctx->use_graphs = value ? 1 : 0; ctx->graph_A_stale = ctx->graph_S_stale = true; }
  else if (option == SMF_OPT_ORTHO_SINGLE_PASS) ctx->ortho_single_pass = value ? 1 : 0;
  else if (option == SMF_OPT_PIPELINED_KERNEL) ctx->use_pipelined = value ? 1 : 0;
  else if (option == SMF_OPT_CURVED_SUMFAC) ctx->use_curved_fast = value ? 1 : 0;
  else if (option == SMF_OPT_CURVED_TMA) ctx->use_curved_tma = value ? 1 : 0;
  else if (option == SMF_OPT_MAX_GRID_BLOCKS) ctx->max_grid_blocks = value > 0 ? value : 0;
  else if (option == SMF_OPT_P2P_EXCHANGE) ctx->use_p2p = value ? 1 : 0;
  else if (option == SMF_OPT_WARP_BRICK_KERNEL) ctx->use_warp_brick = value ? 1 : 0;
  else if (option == SMF_OPT_LATTICE_VELOCITY) ctx->use_vlat = value ? 1 : 0;
  else if (option == SMF_OPT_REPLICATE_COARSE_A) { ctx->use_replica = value ? 1 : 0; lattice_free(ctx); ctx->smoothers_ready = false; }
  else if (option == SMF_OPT_LATTICE_SCHUR) { ctx->use_lattice = value ? 1 : 0; lattice_free(ctx); ctx->smoothers_ready = false; }
  else if (option == SMF_OPT_NONSYMMETRIC_A_SOLVER) ctx->force_nonsymmetric_A = value ? 1 : 0;
  else if (option == SMF_OPT_COARSE_TAIL) { ctx->use_coarse_tail = value ? 1 : 0; if (value > 1) ctx->tail_max_cells = value; ctx->smoothers_ready = false; ctx->graph_A_stale = ctx->graph_S_stale = true; }
  else return ctx->fail(SMF_ERR_INVALID, "unknown option");
  return SMF_OK;
}

int smf_set_stream(smf_ctx *ctx, void *cuda_stream)
{
  if (!ctx) return SMF_ERR_INVALID;
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return SMF_OK;
}

int smf_set_level(smf_ctx *ctx, int level, const smf_level_mesh *m)
{
  if (!ctx || !m) return SMF_ERR_INVALID;
  if (level < SMF_ACTIVE || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  if (m->geometry != 0 && m->geometry != 1) return ctx->fail(SMF_ERR_UNSUPPORTED, "level geometry: 0 (Cartesian) or 1 (per-quadrature-point inverse Jacobians and JxW)");
  if (m->n_cells < 1 || !m->cell_dofs_u || !m->cell_dofs_p) return ctx->fail(SMF_ERR_INVALID, "incomplete level mesh");
  if (m->geometry == 0 && !m->cell_size) return ctx->fail(SMF_ERR_INVALID, "Cartesian level mesh without cell_size");
  if (m->geometry == 1 && (!m->inverse_jacobians || !m->jxw)) return ctx->fail(SMF_ERR_INVALID, "curved level mesh without inverse_jacobians / jxw");
  if (m->n_dofs_u >= (int64_t)NODE_IDX_MASK || m->n_dofs_p >= (int64_t)NODE_IDX_MASK)
    return ctx->fail(SMF_ERR_UNSUPPORTED, "more than 2^28 - 1 local DoFs per rank and block (28-bit node index of the batch lists)");
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  LevelDev &L = ctx->lev(level);
  lattice_free(ctx);
  cudaFree(ctx->ad_in_u); cudaFree(ctx->ad_out_u); cudaFree(ctx->ad_in_p); cudaFree(ctx->ad_out_p);      // staging of the adaptive V-cycle graphs
  ctx->ad_in_u = ctx->ad_out_u = ctx->ad_in_p = ctx->ad_out_p = nullptr;
  ctx->ad_warm_u = ctx->ad_warm_p = false;
  ctx->graph_A_stale = ctx->graph_S_stale = true;
  ctx->bytes -= L.bytes;
  free_level(L);
  const int dim = ctx->cfg.dim, NN = dim == 3 ? 27 : 9, NP = dim == 3 ? 8 : 4;
  L.dim = dim; L.n_cells = m->n_cells; L.n_u = m->n_dofs_u; L.n_p = m->n_dofs_p;
  for (int d = 0; d < dim; ++d) L.h[d] = m->cell_size ? m->cell_size[d] : 1.0;
  if (m->geometry == 1) {
    // [cell][q][r][d], [cell][q]  ->  blocks of 64 cells, rows q * (dim*dim + 1) + k (blk_index): consecutive cells (one
    // per thread) read consecutive doubles and a block's rows are contiguous
    const int D2 = dim * dim, G = D2 + 1;
    std::vector<double> geo((size_t)NN * G * blk_cells(L.n_cells), 1.0);
    for (int64_t cell = 0; cell < L.n_cells; ++cell)
      for (int q = 0; q < NN; ++q) {
        const double *ji = m->inverse_jacobians + ((size_t)cell * NN + q) * D2;
        for (int k = 0; k < D2; ++k) geo[blk_index(cell, q * G + k, NN * G)] = ji[k];
        const double w = m->jxw[(size_t)cell * NN + q];
        if (!(w > 0.0)) return ctx->fail(SMF_ERR_INVALID, "jxw must be positive (inverted cell?)");
        geo[blk_index(cell, q * G + D2, NN * G)] = w;
      }
    if (int e = upload(ctx, &L.geo, geo, &L.bytes)) return e;
  }

  // constraint bitmaps
  std::vector<uint8_t> cu((size_t)L.n_u, 0), cp((size_t)L.n_p, 0);
  for (int64_t i = 0; i < m->n_constrained_u; ++i) {
    if (m->constrained_u[i] >= (uint64_t)L.n_u) return ctx->fail(SMF_ERR_INVALID, "constrained velocity index out of range");
    cu[m->constrained_u[i]] = 1;
  }
  for (int64_t i = 0; i < m->n_constrained_p; ++i) {
    if (m->constrained_p[i] >= (uint64_t)L.n_p) return ctx->fail(SMF_ERR_INVALID, "constrained pressure index out of range");
    cp[m->constrained_p[i]] = 1;
  }
  // node maps (component 0 index | per-component constraint flags); verify the interleaved-components layout
  std::vector<uint32_t> nodes_u((size_t)L.n_cells * NN), nodes_p((size_t)L.n_cells * NP);
  for (int64_t cell = 0; cell < L.n_cells; ++cell) {
    const uint32_t *cd = m->cell_dofs_u + cell * dim * NN;
    for (int n = 0; n < NN; ++n) {
      const uint32_t i0 = cd[n];
      if ((int64_t)i0 + dim > L.n_u) return ctx->fail(SMF_ERR_INVALID, "velocity DoF index out of range");
      uint32_t flags = 0;
      for (int c = 0; c < dim; ++c) {
        if (cd[c * NN + n] != i0 + c)
          return ctx->fail(SMF_ERR_UNSUPPORTED, "cell_dofs_u: the components of a node are not consecutive DoFs");
        if (cu[i0 + c]) flags |= 1u << (NODE_FLAG_SHIFT + c);
      }
      nodes_u[cell * NN + n] = i0 | flags;
    }
    const uint32_t *pd = m->cell_dofs_p + cell * NP;
    for (int n = 0; n < NP; ++n) {
      if (pd[n] >= (uint64_t)L.n_p) return ctx->fail(SMF_ERR_INVALID, "pressure DoF index out of range");
      nodes_p[cell * NP + n] = pd[n] | (cp[pd[n]] ? (1u << 31) : 0u);
    }
  }
  // batches
  int round_u[27], round_p[8];
  for (int n = 0; n < NN; ++n)
    round_u[n] = dim == 3 ? ((n % 3) == 2) + 2 * (((n / 3) % 3) == 2) + 4 * ((n / 9) == 2) : ((n % 3) == 2) + 2 * ((n / 3) == 2);
  for (int n = 0; n < NP; ++n) round_p[n] = n;
  std::vector<uint32_t> nl_u, nl_p;
  std::vector<uint16_t> loc_u, loc_p;
  std::vector<int> nn_u, ne_u, nn_p, ne_p;
  bool safe_u = true, safe_p = true;
  build_batches(nodes_u, L.n_cells, NN, L.n_u, round_u, 1 << dim, nl_u, loc_u, nn_u, ne_u, L.max_nodes_u, safe_u);
  build_batches(nodes_p, L.n_cells, NP, L.n_p, round_p, NP, nl_p, loc_p, nn_p, ne_p, L.max_nodes_p, safe_p);
  if (L.max_nodes_u > 65535 || L.max_nodes_p > 65535) return ctx->fail(SMF_ERR_UNSUPPORTED, "batch too large for 16-bit slots");
  std::vector<uint16_t> slot_u(2, 0);
  L.list_stride_u = L.max_nodes_u;
  if (dim == 3 && ctx->allow_structured) {
    std::vector<uint32_t> nl_s;
    if (build_structured3(nodes_u, L.n_cells, L.n_u, nl_s, slot_u)) {
      L.structured = 1;
      nl_u.swap(nl_s);
      L.max_nodes_u = SB_SLOTS;
      L.list_stride_u = SB_LIST;
      for (size_t b = 0; b < nn_u.size(); ++b) { nn_u[b] = 9 * 9 * 5; ne_u[b] = 0; }
      loc_u.assign(2, 0);                    // slots are computed in the kernel
      safe_u = true;
    }
  }
  L.safe_rounds = (safe_u && safe_p) ? 1 : 0;
  L.n_batches = (int)nn_u.size();
  std::vector<int> info((size_t)L.n_batches * BATCH_INFO, 0);
  for (int b = 0; b < L.n_batches; ++b) {
    info[b * BATCH_INFO + 0] = nn_u[b]; info[b * BATCH_INFO + 1] = ne_u[b];
    info[b * BATCH_INFO + 2] = nn_p[b]; info[b * BATCH_INFO + 3] = ne_p[b];
    info[b * BATCH_INFO + 4] = (int)std::min<int64_t>(BATCH_CELLS, L.n_cells - (int64_t)b * BATCH_CELLS);
  }
  std::vector<uint32_t> cons_u(m->constrained_u, m->constrained_u + m->n_constrained_u);
  std::vector<uint32_t> cons_p(m->constrained_p, m->constrained_p + m->n_constrained_p);
  L.n_constrained_u = m->n_constrained_u; L.n_constrained_p = m->n_constrained_p;
  int e;
  if ((e = upload(ctx, &L.batch_info, info, &L.bytes))) return e;
  if ((e = upload(ctx, &L.node_list_u, nl_u, &L.bytes))) return e;
  if ((e = upload(ctx, &L.node_slot_u, slot_u, &L.bytes))) return e;
  std::vector<uint32_t> plist;
  if (L.structured) L.h_sb_list = nl_u;      // (smf_set_transfer derives the owner bits of the brick-wise transfers from it)
  if (L.structured && build_structured3_p(nodes_p, L.n_cells, L.n_p, plist)) {
    // entries and slots of a brick side by side, then its pressure entries: one contiguous block per brick for the pipelined kernel
    std::vector<unsigned char> sbl((size_t)L.n_batches * SB_BLOCK_BYTES, 0);
    for (int b = 0; b < L.n_batches; ++b) {
      unsigned char *blk = sbl.data() + (size_t)b * SB_BLOCK_BYTES;
      std::memcpy(blk, nl_u.data() + (size_t)b * SB_LIST, SB_LIST * 4);
      std::memcpy(blk + SB_LIST * 4, slot_u.data() + (size_t)b * SB_LIST, SB_LIST * 2);
      std::memcpy(blk + SB_LIST_BYTES, plist.data() + (size_t)b * SB_PLIST, SB_PLIST * 4);
    }
    if ((e = upload(ctx, &L.sb_lists, sbl, &L.bytes))) return e;
  }
  if ((e = upload(ctx, &L.node_list_p, nl_p, &L.bytes))) return e;
  if ((e = upload(ctx, &L.cell_loc_u, loc_u, &L.bytes))) return e;
  if ((e = upload(ctx, &L.cell_loc_p, loc_p, &L.bytes))) return e;
  if ((e = upload(ctx, &L.cell_nodes_u, nodes_u, &L.bytes))) return e;
  if ((e = upload(ctx, &L.cell_nodes_p, nodes_p, &L.bytes))) return e;
  if (L.geo) {
    std::vector<uint32_t> tu((size_t)blk_cells(L.n_cells) * NN, 0), tp((size_t)blk_cells(L.n_cells) * NP, 0);
    for (int64_t cell = 0; cell < L.n_cells; ++cell) {
      for (int n = 0; n < NN; ++n) tu[blk_index(cell, n, NN)] = nodes_u[(size_t)cell * NN + n];
      for (int n = 0; n < NP; ++n) tp[blk_index(cell, n, NP)] = nodes_p[(size_t)cell * NP + n];
    }
    if ((e = upload(ctx, &L.cell_nodes_u_t, tu, &L.bytes))) return e;
    if ((e = upload(ctx, &L.cell_nodes_p_t, tp, &L.bytes))) return e;
  }
  if ((e = upload(ctx, &L.constrained_u, cons_u, &L.bytes))) return e;
  if ((e = upload(ctx, &L.constrained_p, cons_p, &L.bytes))) return e;
  SMF_CUDA(cudaMalloc(&L.dinv_A, sizeof(double) * L.n_u));
  SMF_CUDA(cudaMalloc(&L.dinv_M, sizeof(double) * L.n_p));
  L.bytes += sizeof(double) * (L.n_u + L.n_p);
  ctx->bytes += sizeof(double) * (L.n_u + L.n_p);
  ctx->smoothers_ready = false;
  // host copies of the raw maps are needed once more by smf_set_transfer
  L.h_nodes_u = std::move(nodes_u);
  L.h_nodes_p = std::move(nodes_p);
  L.h_cons_u = std::move(cu);
  L.h_cons_p = std::move(cp);
  return SMF_OK;
}

int smf_set_transfer(smf_ctx *ctx, int fine_level, const int32_t *children)
{
  if (!ctx || !children) return SMF_ERR_INVALID;
  if (fine_level < 1 || fine_level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "fine_level out of range");
  LevelDev &F = ctx->levels[fine_level], &C = ctx->levels[fine_level - 1];
  if (F.n_cells == 0 || C.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set both levels before their transfer");
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  const int dim = ctx->cfg.dim, NCH = 1 << dim;
  TransferDev &T = ctx->transfers[fine_level];
  cudaFree(T.patch_u); cudaFree(T.patch_p); cudaFree(T.tb_u); cudaFree(T.children);
  T = TransferDev();
  lattice_free(ctx);
  T.n_parents = C.n_cells;
  T.h_children.assign(children, children + (size_t)C.n_cells * NCH);
  SMF_CUDA(cudaMalloc(&T.children, sizeof(int32_t) * (size_t)C.n_cells * NCH));      // (smf_fill_cell_data: interpolate_to_mg of the DGQ0 viscosity)
  SMF_CUDA(cudaMemcpy(T.children, children, sizeof(int32_t) * (size_t)C.n_cells * NCH, cudaMemcpyHostToDevice));
  // bit 31: this parent writes the node in prolongation (first local parent that touches it);
  // bit 30: this parent contributes the node in restriction (same, and no lower rank owns the node)
  auto build = [&](int npd, const std::vector<uint32_t> &fine_nodes, int64_t n_fine_index, const std::vector<uint8_t> &not_owned,
                   std::vector<uint32_t> &patch) -> bool {
    const int npp = 2 * npd - 1;
    int npatch = 1, nloc = 1;
    for (int d = 0; d < dim; ++d) { npatch *= npp; nloc *= npd; }
    patch.assign((size_t)T.n_parents * npatch, 0);
    std::vector<uint8_t> seen((size_t)n_fine_index, 0);
    for (int64_t par = 0; par < T.n_parents; ++par) {
      if (children[par * NCH] < 0) {       // adaptive mesh: this parent is not refined -- its patch stays empty (no owner bits)
        for (int ch = 1; ch < NCH; ++ch) if (children[par * NCH + ch] >= 0) return false;
        continue;
      }
      for (int s = 0; s < npatch; ++s) {
        int r = s, child = 0, a = 0, mul = 1;
        for (int d = 0; d < dim; ++d) {
          const int sd = r % npp; r /= npp;
          const int o = sd > (npd - 1) ? 1 : 0;
          child |= o << d;
          a += (sd - (npd - 1) * o) * mul;
          mul *= npd;
        }
        const int32_t fc = children[par * NCH + child];
        if (fc < 0 || fc >= F.n_cells) return false;
        const uint32_t idx = fine_nodes[(size_t)fc * nloc + a] & NODE_IDX_MASK;
        uint32_t e = idx;
        if (!seen[idx]) {
          seen[idx] = 1;
          e |= 1u << 31;
          if (not_owned.empty() || !not_owned[idx]) e |= 1u << 30;
        }
        patch[(size_t)par * npatch + s] = e;
      }
    }
    return true;
  };
  std::vector<uint32_t> pu, pp;
  if (!build(3, F.h_nodes_u, F.n_u, F.h_not_owned_u, pu) || !build(2, F.h_nodes_p, F.n_p, F.h_not_owned_p, pp))
    return ctx->fail(SMF_ERR_INVALID, "children table: out-of-range entries, or a parent with only some of its children");
  int64_t bytes = 0;
  int e;
  if ((e = upload(ctx, &T.patch_u, pu, &bytes))) return e;
  if ((e = upload(ctx, &T.patch_p, pp, &bytes))) return e;
  // Structured fine level in 3D: transfer blocks per fine brick (a 4x4x2 brick of fine cells = 2x2x1 parents).
  if (dim == 3 && F.structured && F.sb_lists && !F.h_sb_list.empty()) {
    std::vector<int32_t> parent_of((size_t)F.n_cells, -1), child_no((size_t)F.n_cells, -1);
    for (int64_t par = 0; par < T.n_parents; ++par)
      for (int ch = 0; ch < 8; ++ch) {
        if (children[par * 8 + ch] < 0) continue;
        parent_of[children[par * 8 + ch]] = (int32_t)par; child_no[children[par * 8 + ch]] = ch;
      }
    const int nbr = F.n_batches;
    std::vector<unsigned char> tb((size_t)nbr * TB_BYTES, 0);
    std::vector<uint8_t> seen((size_t)F.n_u, 0);
    bool ok = true;
    for (int b = 0; b < nbr && ok; ++b) {
      uint32_t *cl = reinterpret_cast<uint32_t *>(tb.data() + (size_t)b * TB_BYTES);
      unsigned char *ob = tb.data() + (size_t)b * TB_BYTES + TB_COARSE * 4;
      for (int i = 0; i < TB_COARSE; ++i) cl[i] = NODE_INVALID;
      for (int q = 0; q < 4 && ok; ++q) {          // parent q = qx + 2 qy of the brick: fine lanes 8 q .. 8 q + 7
        const int32_t par = parent_of[(size_t)b * BATCH_CELLS + 8 * q];
        for (int l = 0; l < 8; ++l) {
          const size_t fc = (size_t)b * BATCH_CELLS + 8 * q + l;
          if (parent_of[fc] != par || child_no[fc] != l) ok = false;
        }
        if (!ok || par < 0) { ok = false; break; }
        const int qx = q & 1, qy = q >> 1;
        for (int n = 0; n < 27; ++n) {
          const int id = (2 * qx + n % 3) + 5 * ((2 * qy + (n / 3) % 3) + 5 * (n / 9));
          const uint32_t ce = C.h_nodes_u[(size_t)par * 27 + n];
          if (cl[id] == NODE_INVALID) cl[id] = ce;
          else if ((cl[id] & NODE_IDX_MASK) != (ce & NODE_IDX_MASK)) ok = false;
        }
      }
      const uint32_t *fl = F.h_sb_list.data() + (size_t)b * SB_LIST;
      for (int j = 0; j < 405; ++j) {
        const uint32_t idx = fl[j] & NODE_IDX_MASK;
        unsigned char o = 0;
        if (!seen[idx]) {
          seen[idx] = 1;
          o |= 1;
          if (F.h_not_owned_u.empty() || !F.h_not_owned_u[idx]) o |= 2;
        }
        ob[j] = o;
      }
    }
    if (ok && (e = upload(ctx, &T.tb_u, tb, &bytes))) return e;
  }
  T.set = true;
  return SMF_OK;
}

int smf_set_copy_indices(smf_ctx *ctx, const uint32_t *au, const uint32_t *ap)
{
  if (!ctx) return SMF_ERR_INVALID;
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  if (A.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set the active level first");
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  cudaFree(ctx->copy_u); cudaFree(ctx->copy_p);
  ctx->copy_u = ctx->copy_p = nullptr;
  lattice_free(ctx);
  int64_t bytes = 0;
  int e;
  if (au) { std::vector<uint32_t> v(au, au + A.n_u); if ((e = upload(ctx, &ctx->copy_u, v, &bytes))) return e; }
  if (ap) { std::vector<uint32_t> v(ap, ap + A.n_p); if ((e = upload(ctx, &ctx->copy_p, v, &bytes))) return e; }
  return SMF_OK;
}

}  // extern "C"

namespace smf {
// raw [cell][n_eta] -> lane-transposed [cell/32][n_eta][32], padded with 1.0
__global__ void transpose_eta_kernel(const double *raw, double *out, int64_t n_cells, int n_eta, int64_t n_padded_cells)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_padded_cells * n_eta) return;
  const int64_t grp = t / ((int64_t)n_eta * SUB);
  const int rem = (int)(t % ((int64_t)n_eta * SUB));
  const int q = rem / SUB, lane = rem % SUB;
  const int64_t cell = grp * SUB + lane;
  out[t] = cell < n_cells ? raw[cell * n_eta + q] : 1.0;
}

// ---- coefficient-table pipeline on the device ---------------------------------------------------
// The array work of evaluate_material_model(): cell-wise averaging of the plugin's quadrature-point viscosities
// (material_model/interface.cc:523-650, unweighted over the points), cellwise L2 projection onto DGQ0 = JxW-weighted
// mean (utilities.cc:2738-2812; matrix_free_operators.cc:227-246), min / max after averaging (:248-262), and
// interpolate_to_mg of the DGQ0 vector = mean of the 2^dim children, level by level (local_smoothing.cc:220-226).
// One thread per cell; per-block extrema are finished on the host.
__host__ __device__ constexpr double gw1(int i) { return i == 1 ? 8.0 / 18.0 : 5.0 / 18.0; }      // 3-point Gauss weights on [0,1]
__global__ void k_average_viscosity(const double *raw_q, int64_t n_cells, int nq, int op, const double *geo, double *active_raw, double *dg0,
                                    double *blk_min, double *blk_max)
{
  __shared__ double s_min[256], s_max[256];
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double vmin = 1e300, vmax = -1e300;
  if (cell < n_cells) {
    const double *v = raw_q + cell * nq;
    const int dim = nq == 27 ? 3 : 2;
    double avg = 0.0;
    if (op == 1) { for (int q = 0; q < nq; ++q) avg += v[q]; avg /= nq; }
    else if (op == 2) {
      bool zero = false;
      for (int q = 0; q < nq; ++q) { zero = zero || v[q] == 0.0; avg += 1.0 / v[q]; }
      avg = zero ? 0.0 : 1.0 / (avg / nq);
    } else if (op == 3) { avg = 1.0; for (int q = 0; q < nq; ++q) avg *= pow(v[q], 1.0 / nq); }
    double proj = 0.0, wsum = 0.0;
    for (int q = 0; q < nq; ++q) {
      const double o = op == 0 ? v[q] : avg;
      // Cartesian: reference-cell weights; curved cells: JxW of the cell (entry dim*dim of the geometry table)
      const double w = geo ? geo[blk_index(cell, q * (dim * dim + 1) + dim * dim, nq * (dim * dim + 1))]
                           : (dim == 3 ? gw1(q % 3) * gw1((q / 3) % 3) * gw1(q / 9) : gw1(q % 3) * gw1(q / 3));
      proj += w * o; wsum += w;
      vmin = fmin(vmin, o); vmax = fmax(vmax, o);
      if (op == 0) active_raw[cell * nq + q] = o;
    }
    if (geo) proj /= wsum;                              // (the reference-cell weights sum to 1)
    if (op != 0) active_raw[cell] = proj;
    dg0[cell] = proj;
  }
  s_min[threadIdx.x] = vmin; s_max[threadIdx.x] = vmax;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { s_min[threadIdx.x] = fmin(s_min[threadIdx.x], s_min[threadIdx.x + s]); s_max[threadIdx.x] = fmax(s_max[threadIdx.x], s_max[threadIdx.x + s]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { blk_min[blockIdx.x] = s_min[0]; blk_max[blockIdx.x] = s_max[0]; }
}
__global__ void k_coarsen_dg0(const double *fine, const int32_t *children, int nch, double *coarse, int64_t n_parents)
{
  const int64_t par = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (par >= n_parents) return;
  double s = 0.0;
  for (int c = 0; c < nch; ++c) s += fine[children[par * nch + c]];
  coarse[par] = s / nch;
}
// ---- `Material averaging = project to Q1 [only viscosity]`: DGQ1 projection space (local_smoothing.cc:99-111) ----
// k_project_q1: cellwise L2 projection (JxW-weighted; Utilities::project_cellwise) of the RAW quadrature-point viscosities onto
// the 2^dim Q1 shape functions: (Phi^T W Phi) c = Phi^T W v, solved per cell by Cholesky; extrema of the raw values per block.
// k_eval_q1: the DGQ1 function back at the quadrature points, clamped to the global [min, max] (matrix_free_operators.cc:
// 306-325, local_smoothing.cc:274-289).  k_coarsen_dg1: interpolate_to_mg of the DGQ1 vector -- FE_DGQ(1) is interpolatory,
// its restriction is not additive: the parent's nodal value at vertex k is child k's own nodal value k.
__device__ __forceinline__ double q1_shape_at(int dim, int q, int i)
{
  const double s = 0.7745966692414834, gx[3] = {0.5 - 0.5 * s, 0.5, 0.5 + 0.5 * s};
  double v = 1.0;
  for (int d = 0; d < dim; ++d) { const double x = gx[q % 3]; q /= 3; v *= ((i >> d) & 1) ? x : 1.0 - x; }
  return v;
}
__global__ void k_project_q1(const double *raw_q, int64_t n_cells, int nq, const double *geo, double *coef, double *blk_min, double *blk_max)
{
  __shared__ double s_min[128], s_max[128];
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double vmin = 1e300, vmax = -1e300;
  if (cell < n_cells) {
    const int dim = nq == 27 ? 3 : 2, nn = 1 << dim;
    double M[8][8], b[8];
    for (int i = 0; i < nn; ++i) { b[i] = 0.0; for (int j = 0; j < nn; ++j) M[i][j] = 0.0; }
    for (int q = 0; q < nq; ++q) {
      const double v = raw_q[cell * nq + q];
      vmin = fmin(vmin, v); vmax = fmax(vmax, v);
      const double w = geo ? geo[blk_index(cell, q * (dim * dim + 1) + dim * dim, nq * (dim * dim + 1))]
                           : (dim == 3 ? gw1(q % 3) * gw1((q / 3) % 3) * gw1(q / 9) : gw1(q % 3) * gw1(q / 3));
      double phi[8];
      for (int i = 0; i < nn; ++i) phi[i] = q1_shape_at(dim, q, i);
      for (int i = 0; i < nn; ++i) { b[i] += w * v * phi[i]; for (int j = 0; j <= i; ++j) M[i][j] += w * phi[i] * phi[j]; }
    }
    for (int j = 0; j < nn; ++j) {                 // Cholesky M = L L^T (lower triangle in place)
      for (int k = 0; k < j; ++k) M[j][j] -= M[j][k] * M[j][k];
      M[j][j] = sqrt(M[j][j]);
      for (int i = j + 1; i < nn; ++i) { for (int k = 0; k < j; ++k) M[i][j] -= M[i][k] * M[j][k]; M[i][j] /= M[j][j]; }
    }
    for (int i = 0; i < nn; ++i) { for (int k = 0; k < i; ++k) b[i] -= M[i][k] * b[k]; b[i] /= M[i][i]; }
    for (int i = nn - 1; i >= 0; --i) { for (int k = i + 1; k < nn; ++k) b[i] -= M[k][i] * b[k]; b[i] /= M[i][i]; }
    for (int i = 0; i < nn; ++i) coef[cell * nn + i] = b[i];
  }
  s_min[threadIdx.x] = vmin; s_max[threadIdx.x] = vmax;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { s_min[threadIdx.x] = fmin(s_min[threadIdx.x], s_min[threadIdx.x + s]); s_max[threadIdx.x] = fmax(s_max[threadIdx.x], s_max[threadIdx.x + s]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { blk_min[blockIdx.x] = s_min[0]; blk_max[blockIdx.x] = s_max[0]; }
}
__global__ void k_eval_q1(const double *coef, int64_t n_cells, int nq, double lo, double hi, double *table)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cells * nq) return;
  const int64_t cell = t / nq; const int q = (int)(t % nq);
  const int dim = nq == 27 ? 3 : 2, nn = 1 << dim;
  double v = 0.0;
  for (int i = 0; i < nn; ++i) v += coef[cell * nn + i] * q1_shape_at(dim, q, i);
  table[t] = fmin(fmax(v, lo), hi);
}
__global__ void k_coarsen_dg1(const double *fine, const int32_t *children, int nch, double *coarse, int64_t n_parents)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_parents * nch) return;
  const int64_t par = t / nch; const int k = (int)(t % nch);
  coarse[t] = fine[(int64_t)children[par * nch + k] * nch + k];
}
}  // namespace smf

extern "C" {

int smf_set_cell_data(smf_ctx *ctx, int level, const smf_cell_data *d)
{
  if (!ctx || !d || !d->viscosity) return SMF_ERR_INVALID;
  if (level < SMF_ACTIVE || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  LevelDev &L = ctx->lev(level);
  if (L.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set the level mesh first");
  const int NN = L.dim == 3 ? 27 : 9;
  if (d->n_per_cell != 1 && d->n_per_cell != NN) return ctx->fail(SMF_ERR_INVALID, "n_per_cell must be 1 or 3^dim");
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  const int64_t padded = (int64_t)L.n_batches * BATCH_CELLS;
  const int64_t n_out = padded * d->n_per_cell;
  if (L.eta && L.n_eta != d->n_per_cell) { cudaFree(L.eta); L.eta = nullptr; }
  if (!L.eta) {
    SMF_CUDA(cudaMalloc(&L.eta, sizeof(double) * n_out));
    L.bytes += sizeof(double) * n_out; ctx->bytes += sizeof(double) * n_out;
  }
  double *raw = nullptr;
  SMF_CUDA(cudaMalloc(&raw, sizeof(double) * L.n_cells * d->n_per_cell));
  SMF_CUDA(cudaMemcpyAsync(raw, d->viscosity, sizeof(double) * L.n_cells * d->n_per_cell, cudaMemcpyHostToDevice, ctx->stream));
  transpose_eta_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, ctx->stream>>>(raw, L.eta, L.n_cells, d->n_per_cell, padded);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(raw);
  L.n_eta = d->n_per_cell; L.ps = d->pressure_scaling; L.compressible = d->is_compressible;
  L.has_cell_data = true; L.has_diag = false;
  ctx->smoothers_ready = false;
  return SMF_OK;
}

// table of one level from a raw device array [n_cells][n_per_cell] (shared by smf_set_cell_data / smf_fill_cell_data)
static int install_table(smf_ctx *ctx, LevelDev &L, const double *raw_dev, int n_per_cell, double ps, int compressible)
{
  const int64_t padded = (int64_t)L.n_batches * BATCH_CELLS;
  const int64_t n_out = padded * n_per_cell;
  if (L.eta && L.n_eta != n_per_cell) { cudaFree(L.eta); L.eta = nullptr; }
  if (!L.eta) {
    SMF_CUDA(cudaMalloc(&L.eta, sizeof(double) * n_out));
    L.bytes += sizeof(double) * n_out; ctx->bytes += sizeof(double) * n_out;
  }
  transpose_eta_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, ctx->stream>>>(raw_dev, L.eta, L.n_cells, n_per_cell, padded);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  L.n_eta = n_per_cell; L.ps = ps; L.compressible = compressible;
  L.has_cell_data = true; L.has_diag = false;
  return SMF_OK;
}

int smf_fill_cell_data(smf_ctx *ctx, const double *viscosity_q, int averaging, double pressure_scaling, int is_compressible,
                       double *min_max_out)
{
  if (!ctx || !viscosity_q) return SMF_ERR_INVALID;
  if (averaging < 0 || averaging > 4) return ctx->fail(SMF_ERR_UNSUPPORTED, "averaging: 0 none, 1 arithmetic, 2 harmonic, 3 geometric, 4 project to Q1 (pick largest / log average are rejected by the reference's GMG path as well)");
  const int nl = ctx->cfg.n_levels;
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  if (A.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set the level meshes first");
  if (A.n_cells != ctx->levels[nl - 1].n_cells) return ctx->fail(SMF_ERR_UNSUPPORTED, "active mesh differs from the finest level (adaptive refinement)");
  for (int l = 1; l < nl; ++l)
    if (!ctx->transfers[l].children) return ctx->fail(SMF_ERR_INVALID, "smf_set_transfer must precede smf_fill_cell_data");
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  const int nq = A.dim == 3 ? 27 : 9, nch = 1 << A.dim;
  const int64_t nc = A.n_cells;
  if (averaging == SMF_AVG_PROJECT_Q1) {
    // DGQ1 tables: n_q values per cell on the active mesh and on every level
    const unsigned pb = (unsigned)((nc + 127) / 128);
    double *raw = nullptr, *coef = nullptr, *ext = nullptr, *table = nullptr;
    SMF_CUDA(cudaMalloc(&raw, sizeof(double) * nc * nq));
    SMF_CUDA(cudaMalloc(&coef, sizeof(double) * nc * nch));
    SMF_CUDA(cudaMalloc(&table, sizeof(double) * nc * nq));
    SMF_CUDA(cudaMalloc(&ext, sizeof(double) * 2 * pb));
    SMF_CUDA(cudaMemcpyAsync(raw, viscosity_q, sizeof(double) * nc * nq, cudaMemcpyHostToDevice, ctx->stream));
    k_project_q1<<<pb, 128, 0, ctx->stream>>>(raw, nc, nq, A.geo, coef, ext, ext + pb);
    ctx->launches++;
    std::vector<double> h_ext(2 * (size_t)pb);
    SMF_CUDA(cudaMemcpyAsync(h_ext.data(), ext, sizeof(double) * 2 * pb, cudaMemcpyDeviceToHost, ctx->stream));
    SMF_CUDA(cudaStreamSynchronize(ctx->stream));
    double mn = 1e300, mx = -1e300;
    for (unsigned b = 0; b < pb; ++b) { mn = std::min(mn, h_ext[b]); mx = std::max(mx, h_ext[pb + b]); }
    // the clamp uses the GLOBAL range (Utilities::MPI::min / max, matrix_free_operators.cc:258-259)
    if (int e = allreduce_minmax(ctx, &mn, &mx)) { cudaFree(raw); cudaFree(coef); cudaFree(table); cudaFree(ext); return e; }
    int rc = SMF_OK;
    double *fine = coef;
    std::vector<double *> tmp;
    for (int l = nl - 1; l >= 0 && rc == SMF_OK; --l) {
      LevelDev &L = ctx->levels[l];
      k_eval_q1<<<(unsigned)((L.n_cells * nq + 255) / 256), 256, 0, ctx->stream>>>(fine, L.n_cells, nq, mn, mx, table);
      ctx->launches++;
      if (l == nl - 1) rc = install_table(ctx, A, table, nq, pressure_scaling, is_compressible);
      if (rc == SMF_OK) rc = install_table(ctx, L, table, nq, pressure_scaling, is_compressible);
      if (l > 0 && rc == SMF_OK) {
        LevelDev &C = ctx->levels[l - 1];
        double *coarse = nullptr;
        if (cudaMalloc(&coarse, sizeof(double) * C.n_cells * nch) != cudaSuccess) { rc = ctx->fail(SMF_ERR_CUDA, "cudaMalloc failed"); break; }
        tmp.push_back(coarse);
        k_coarsen_dg1<<<(unsigned)((C.n_cells * nch + 255) / 256), 256, 0, ctx->stream>>>(fine, ctx->transfers[l].children, nch, coarse, C.n_cells);
        ctx->launches++;
        fine = coarse;
      }
    }
    cudaStreamSynchronize(ctx->stream);
    if (rc == SMF_OK && cudaGetLastError() != cudaSuccess) rc = ctx->fail(SMF_ERR_CUDA, "smf_fill_cell_data: kernel failure");
    cudaFree(raw); cudaFree(coef); cudaFree(table); cudaFree(ext);
    for (double *p : tmp) cudaFree(p);
    if (rc != SMF_OK) return rc;
    if (min_max_out) { min_max_out[0] = mn; min_max_out[1] = mx; }
    ctx->smoothers_ready = false;
    return SMF_OK;
  }
  const int n_act = averaging == 0 ? nq : 1;
  const unsigned blocks = (unsigned)((nc + 255) / 256);
  double *raw = nullptr, *act = nullptr, *dg0 = nullptr, *ext = nullptr;
  SMF_CUDA(cudaMalloc(&raw, sizeof(double) * nc * nq));
  SMF_CUDA(cudaMalloc(&act, sizeof(double) * nc * n_act));
  SMF_CUDA(cudaMalloc(&dg0, sizeof(double) * nc));
  SMF_CUDA(cudaMalloc(&ext, sizeof(double) * 2 * blocks));
  SMF_CUDA(cudaMemcpyAsync(raw, viscosity_q, sizeof(double) * nc * nq, cudaMemcpyHostToDevice, ctx->stream));
  k_average_viscosity<<<blocks, 256, 0, ctx->stream>>>(raw, nc, nq, averaging, A.geo, act, dg0, ext, ext + blocks);
  ctx->launches++;
  int rc = install_table(ctx, A, act, n_act, pressure_scaling, is_compressible);
  double *fine = dg0;
  std::vector<double *> tmp;
  for (int l = nl - 1; l >= 0 && rc == SMF_OK; --l) {
    LevelDev &L = ctx->levels[l];
    rc = install_table(ctx, L, fine, 1, pressure_scaling, is_compressible);
    if (l > 0 && rc == SMF_OK) {
      LevelDev &C = ctx->levels[l - 1];
      double *coarse = nullptr;
      if (cudaMalloc(&coarse, sizeof(double) * C.n_cells) != cudaSuccess) { rc = ctx->fail(SMF_ERR_CUDA, "cudaMalloc failed"); break; }
      tmp.push_back(coarse);
      k_coarsen_dg0<<<(unsigned)((C.n_cells + 255) / 256), 256, 0, ctx->stream>>>(fine, ctx->transfers[l].children, nch, coarse, C.n_cells);
      ctx->launches++;
      fine = coarse;
    }
  }
  std::vector<double> h_ext(2 * (size_t)blocks);
  if (rc == SMF_OK && cudaMemcpyAsync(h_ext.data(), ext, sizeof(double) * 2 * blocks, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = SMF_ERR_CUDA;
  cudaStreamSynchronize(ctx->stream);
  if (rc == SMF_OK && cudaGetLastError() != cudaSuccess) rc = ctx->fail(SMF_ERR_CUDA, "smf_fill_cell_data: kernel failure");
  cudaFree(raw); cudaFree(act); cudaFree(dg0); cudaFree(ext);
  for (double *p : tmp) cudaFree(p);
  if (rc != SMF_OK) return rc;
  {
    double mn = 1e300, mx = -1e300;
    for (unsigned b = 0; b < blocks; ++b) { mn = std::min(mn, h_ext[b]); mx = std::max(mx, h_ext[blocks + b]); }
    if ((rc = allreduce_minmax(ctx, &mn, &mx))) return rc;       // "Viscosity range" is global (Utilities::MPI::min / max)
    if (min_max_out) { min_max_out[0] = mn; min_max_out[1] = mx; }
  }
  ctx->smoothers_ready = false;
  return SMF_OK;
}

// un-transposed copy of a level's viscosity table [n_cells][n_per_cell] (tests, post-processing)
int smf_get_cell_data(smf_ctx *ctx, int level, double *host_out, int *n_per_cell_out)
{
  if (!ctx || !host_out) return SMF_ERR_INVALID;
  if (level < SMF_ACTIVE || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  LevelDev &L = ctx->lev(level);
  if (!L.has_cell_data) return ctx->fail(SMF_ERR_INVALID, "cell data not set");
  const int64_t padded = (int64_t)L.n_batches * BATCH_CELLS;
  std::vector<double> t((size_t)padded * L.n_eta);
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  SMF_CUDA(cudaMemcpy(t.data(), L.eta, sizeof(double) * t.size(), cudaMemcpyDeviceToHost));
  for (int64_t cell = 0; cell < L.n_cells; ++cell)
    for (int q = 0; q < L.n_eta; ++q)
      host_out[cell * L.n_eta + q] = t[((size_t)(cell / BATCH_CELLS) * L.n_eta + q) * BATCH_CELLS + cell % BATCH_CELLS];
  if (n_per_cell_out) *n_per_cell_out = L.n_eta;
  return SMF_OK;
}

int smf_set_newton_data(smf_ctx *ctx, const smf_newton_data *d)
{
  if (!ctx) return SMF_ERR_INVALID;
  LevelDev &L = ctx->lev(SMF_ACTIVE);
  if (L.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set the active level mesh first");
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  NewtonDev &N = L.nd;
  cudaFree(N.eps); cudaFree(N.deta); cudaFree(N.detadp); cudaFree(N.dil_lhs); cudaFree(N.ddil); cudaFree(N.ddildp);
  N = NewtonDev();
  L.has_diag = false;                 // (prescribed dilation makes the A block deviatoric)
  ctx->smoothers_ready = false;
  if (!d) return SMF_OK;
  if (d->apply_stabilization_free_surface_faces)
    return ctx->fail(SMF_ERR_UNSUPPORTED, "free-surface stabilisation faces (matrix_free_operators.cc:563-602) are not implemented");
  const int nq = L.dim == 3 ? 27 : 9, ns = L.dim * (L.dim + 1) / 2;
  auto up = [&](double **dev, const double *host, int per_q) -> int {
    if (!host) return ctx->fail(SMF_ERR_INVALID, "smf_set_newton_data: a table required by the enabled flags is missing");
    const size_t nb = sizeof(double) * (size_t)L.n_cells * nq * per_q;
    SMF_CUDA(cudaMalloc((void **)dev, nb));
    SMF_CUDA(cudaMemcpy(*dev, host, nb, cudaMemcpyHostToDevice));
    ctx->bytes += (int64_t)nb;
    return SMF_OK;
  };
  int e;
  if (d->enable_newton_derivatives) {
    if ((e = up(&N.eps, d->strain_rate_table, ns))) return e;
    if ((e = up(&N.deta, d->newton_factor_wrt_strain_rate_table, ns))) return e;
    if ((e = up(&N.detadp, d->newton_factor_wrt_pressure_table, 1))) return e;
  }
  if (d->enable_prescribed_dilation) {
    if ((e = up(&N.dil_lhs, d->dilation_lhs_term_table, 1))) return e;
    if (d->enable_newton_derivatives) {
      if ((e = up(&N.ddil, d->dilation_derivative_wrt_strain_rate_table, ns))) return e;
      if ((e = up(&N.ddildp, d->dilation_derivative_wrt_pressure_table, 1))) return e;
    }
  }
  N.newton = d->enable_newton_derivatives ? 1 : 0;
  N.symmetrize = d->symmetrize_newton_system ? 1 : 0;
  N.average = d->average_newton_factors ? 1 : 0;
  N.dilation = d->enable_prescribed_dilation ? 1 : 0;
  return SMF_OK;
}

int smf_vec_alloc(smf_ctx *ctx, int64_t n, double **p)
{
  if (!ctx || !p || n < 0) return SMF_ERR_INVALID;
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  SMF_CUDA(cudaMalloc(p, sizeof(double) * std::max<int64_t>(n, 1)));
  SMF_CUDA(cudaMemsetAsync(*p, 0, sizeof(double) * std::max<int64_t>(n, 1), ctx->stream));
  ctx->bytes += sizeof(double) * n;
  return SMF_OK;
}
int smf_vec_free(smf_ctx *ctx, double *p)
{
  if (!ctx) return SMF_ERR_INVALID;
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  SMF_CUDA(cudaFree(p));
  return SMF_OK;
}
int smf_vec_h2d(smf_ctx *ctx, double *dev, const double *host, int64_t n)
{
  if (!ctx) return SMF_ERR_INVALID;
  SMF_CUDA(cudaMemcpyAsync(dev, host, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  return SMF_OK;
}
int smf_vec_d2h(smf_ctx *ctx, double *host, const double *dev, int64_t n)
{
  if (!ctx) return SMF_ERR_INVALID;
  SMF_CUDA(cudaMemcpyAsync(host, dev, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  return SMF_OK;
}
int smf_vec_copy(smf_ctx *ctx, double *dst, const double *src, int64_t n)
{ if (!ctx) return SMF_ERR_INVALID; return vec_copy(ctx, dst, src, n); }
int smf_vec_set(smf_ctx *ctx, double *dev, double value, int64_t n)
{
  if (!ctx || (!dev && n > 0)) return SMF_ERR_INVALID;
  return smf::vec_set(ctx, dev, value, n);
}
int smf_synchronize(smf_ctx *ctx)
{
  if (!ctx) return SMF_ERR_INVALID;
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  return SMF_OK;
}
int64_t smf_kernel_launches(const smf_ctx *ctx) { return ctx ? ctx->launches : 0; }
int64_t smf_device_bytes(const smf_ctx *ctx) { return ctx ? ctx->bytes : 0; }
int64_t smf_exchange_count(const smf_ctx *ctx) { return ctx ? ctx->exchanges : 0; }

}  // extern "C"
