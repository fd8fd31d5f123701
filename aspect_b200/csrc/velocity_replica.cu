// Replicated coarse levels of the A-block V-cycle on several ranks (coarse-level agglomeration, SURVEY 2.2 C6, in its
// communication-free form).
//
// On P ranks every level of prec_A exchanges its interface sums 9 times per cycle, and below ~10^4 cells per rank an exchange
// costs the same ~25-30 us whatever its size (a neighbourhood barrier of the ranks): at 8 GPUs the five coarsest levels
// accounted for 45 of the 72 exchanges of a cycle and did almost no arithmetic.  When the subdomains are boxes of a global box
// lattice (schur_lattice.cu has located them from the partition-independent DoF numbers), those levels do not need to be
// distributed at all: every rank builds the GLOBAL level meshes itself -- cell -> DoF maps in lexicographic lattice numbering,
// the boundary faces constrained per component as the local levels say, viscosity tables gathered by one all-reduce per
// assembly -- in a private child context of this same library (world = 1, same stream), and replaces the recursion below
// level `top` by:  scatter the restricted defect (owned entries) into the global vector, ONE all-reduce, the child's V-cycle
// (levels 0..top, one CUDA graph, no communication), gather the rank's part of the result (shared DoFs take their owner's copy:
// one exchange, so that copies stay bitwise equal).  The child's smoothers take the parent's eigenvalue estimates (same matrices).
#include "smf_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>

using namespace smf;

namespace smf {

#define TRY(x) do { int e__ = (x); if (e__) return e__; } while (0)

struct ReplicaLevel {
  int gcx = 0, gcy = 0, gcz = 0;          // global cells per direction
  int64_t n_cells_global = 0;
  uint32_t *cellmap = nullptr;            // [local cells] index of the cell in the child's cell order
  double *geta = nullptr;                 // [global cells] viscosity table
};
struct Replica {
  smf_ctx *child = nullptr;
  int top = -1;
  int64_t n_global = 0, n_local = 0;      // velocity DoFs of level `top`: global / this rank
  double *gb = nullptr, *gx = nullptr;    // global defect / correction
  unsigned char *not_owned = nullptr;     // [n_local] 1: a lower rank owns the DoF
  std::vector<ReplicaLevel> lv;
  std::vector<double> h_eta;              // host staging of a viscosity table
  bool child_ready = false;
};

__global__ void k_rep_scatter_owned(double *g, const double *v, const int64_t *gid, const unsigned char *not_owned, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !not_owned[i]) g[gid[i]] = v[i];
}
__global__ void k_rep_gather(double *v, const double *g, const int64_t *gid, const unsigned char *not_owned, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = not_owned[i] ? 0.0 : g[gid[i]];
}
__global__ void k_rep_eta(double *geta, const double *eta, const uint32_t *cellmap, int64_t n_cells)
{
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_cells) geta[cellmap[c]] = eta[c];      // (n_eta = 1: the lane-transposed table is the plain cell order)
}

void replica_free(smf_ctx *ctx)
{
  Replica *R = (Replica *)ctx->replica;
  if (!R) return;
  if (R->child) smf_destroy(R->child);
  cudaFree(R->gb); cudaFree(R->gx); cudaFree(R->not_owned);
  for (auto &l : R->lv) { cudaFree(l.cellmap); cudaFree(l.geta); }
  delete R;
  ctx->replica = nullptr;
  ctx->replica_top = -1;
}

// cell index of global cell (x, y, z) in the child's order: 4 x 4 x 2 bricks in the lane order of the brick kernels when the
// level divides into bricks, lexicographic otherwise
static int64_t child_cell_index(const ReplicaLevel &L, int x, int y, int z)
{
  if (L.gcx % 4 == 0 && L.gcy % 4 == 0 && L.gcz % 2 == 0) {
    const int bx = x >> 2, by = y >> 2, bz = z >> 1, lx = x & 3, ly = y & 3, lz = z & 1;
    const int lane = (lx & 1) | ((ly & 1) << 1) | (lz << 2) | ((lx >> 1) << 3) | ((ly >> 1) << 4);
    return ((int64_t)(bz * (L.gcy / 4) + by) * (L.gcx / 4) + bx) * 32 + lane;
  }
  return ((int64_t)z * L.gcy + y) * L.gcx + x;
}

// local checks + what the ranks have to agree on: votes[0] = ok, votes[1 + 2 k], votes[2 + 2 k] = nodes inside boundary face
// f with component c constrained / not constrained (k = 3 f + c)
static bool replica_local(smf_ctx *ctx, int top, std::vector<std::vector<int>> &cell_xyz, double *votes)
{
  for (int l = 0; l <= top; ++l) {
    LevelDev &L = ctx->levels[l];
    const LatLevel &T = ctx->lat[l];
    if (L.geo || L.n_eta != 1 || L.hang_u.n || L.edge_u.n || !L.ex_u.set || !L.ex_u.gid) return false;
    int pd[3];
    if (!lattice_cell_coords(L.h_nodes_p, L.n_cells, L.n_p, pd, cell_xyz[l])) return false;
    if (pd[0] != T.nx || pd[1] != T.ny || pd[2] != T.nz) return false;
    const int nx = 2 * pd[0] - 1, ny = 2 * pd[1] - 1, nz = 2 * pd[2] - 1;
    const int GX = 2 * T.gnx - 1, GY = 2 * T.gny - 1, GZ = 2 * T.gnz - 1, ox = 2 * T.ox, oy = 2 * T.oy, oz = 2 * T.oz;
    if ((int64_t)nx * ny * nz * 3 != L.n_u || L.ex_u.h_gid.size() != (size_t)L.n_u) return false;
    // the partition-independent velocity DoF numbers must be the lexicographic ones of the global node lattice; the constraint
    // flags must be whole boundary faces per component
    for (int64_t c = 0; c < L.n_cells; ++c)
      for (int n = 0; n < 27; ++n) {
        const uint32_t e = L.h_nodes_u[(size_t)c * 27 + n];
        const int X = 2 * cell_xyz[l][c * 3] + n % 3 + ox, Y = 2 * cell_xyz[l][c * 3 + 1] + (n / 3) % 3 + oy, Z = 2 * cell_xyz[l][c * 3 + 2] + n / 9 + oz;
        const int64_t g = 3 * (((int64_t)Z * GY + Y) * GX + X);
        const uint32_t d0 = e & NODE_IDX_MASK;
        for (int k = 0; k < 3; ++k) if (L.ex_u.h_gid[d0 + k] != g + k) return false;
        if (l == top) {
          const int p[3] = {X, Y, Z}, hi[3] = {GX - 1, GY - 1, GZ - 1};
          int on = 0;
          for (int d = 0; d < 3; ++d) on += (p[d] == 0) + (p[d] == hi[d]);
          const uint32_t f = e >> NODE_FLAG_SHIFT;
          if (on == 0 && f) return false;                       // constrained DoF inside the domain
          if (on == 1)                                          // a node inside exactly one boundary face votes for that face
            for (int d = 0; d < 3; ++d)
              for (int s = 0; s < 2; ++s)
                if (p[d] == (s ? hi[d] : 0))
                  for (int k = 0; k < 3; ++k) votes[1 + 2 * (3 * (2 * d + s) + k) + (((f >> k) & 1u) ? 0 : 1)] += 1.0;
        }
      }
  }
  return true;
}

static int replica_build(smf_ctx *ctx)
{
  const int nl = ctx->cfg.n_levels;
  ctx->replica_top = -1;
  // candidate: the largest level whose GLOBAL mesh stays inside one CUDA graph of the child (and is replicated on the Schur side)
  const char *env = getenv("SMF_REPLICATE_A_CELLS");
  const int64_t max_cells = env ? atoll(env) : 70000;
  int top = -1;
  if (ctx->world > 1 && ctx->lat_state == 1 && ctx->cfg.dim == 3 && !ctx->adaptive)
    for (int l = 0; l <= ctx->lat_repl && l < nl - 1; ++l) {
      const LatLevel &T = ctx->lat[l];
      if ((int64_t)(T.gnx - 1) * (T.gny - 1) * (T.gnz - 1) > max_cells) break;
      top = l;
    }
  std::vector<std::vector<int>> cell_xyz(std::max(top + 1, 1));
  std::vector<double> votes(37, 0.0);
  const bool ok_local = top >= 0 && replica_local(ctx, top, cell_xyz, votes.data());
  // collective decision (every rank with world > 1 and a usable lattice path gets here; `top` depends on global sizes only)
  if (ctx->world > 1 && ctx->lat_state == 1) {
    votes[0] = ok_local ? 1.0 : 0.0;
    double *d = nullptr;
    SMF_CUDA(cudaMalloc(&d, sizeof(double) * 37));
    SMF_CUDA(cudaMemcpyAsync(d, votes.data(), sizeof(double) * 37, cudaMemcpyHostToDevice, ctx->stream));
    TRY(allreduce_sum(ctx, d, 37));
    SMF_CUDA(cudaMemcpyAsync(votes.data(), d, sizeof(double) * 37, cudaMemcpyDeviceToHost, ctx->stream));
    SMF_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
  }
  if (top < 0 || votes[0] != (double)ctx->world) return SMF_OK;
  int mask = 0;
  for (int k = 0; k < 18; ++k) {
    if (votes[1 + 2 * k] > 0.0 && votes[2 + 2 * k] > 0.0) return SMF_OK;      // a face that is only partly constrained
    if (votes[1 + 2 * k] > 0.0) mask |= 1 << k;
  }
  {
    // every boundary node of every replicated level (edges and corners included) must carry exactly the union of its faces' patterns
    double good = 1.0;
    for (int l = 0; l <= top && good > 0.0; ++l) {
      const LevelDev &L = ctx->levels[l];
      const LatLevel &T = ctx->lat[l];
      const int GX = 2 * T.gnx - 1, GY = 2 * T.gny - 1, GZ = 2 * T.gnz - 1;
      for (int64_t c = 0; c < L.n_cells && good > 0.0; ++c)
        for (int n = 0; n < 27; ++n) {
          const int p[3] = {2 * (cell_xyz[l][c * 3] + T.ox) + n % 3, 2 * (cell_xyz[l][c * 3 + 1] + T.oy) + (n / 3) % 3, 2 * (cell_xyz[l][c * 3 + 2] + T.oz) + n / 9};
          const int hi[3] = {GX - 1, GY - 1, GZ - 1};
          uint32_t expect = 0;
          for (int d = 0; d < 3; ++d) {
            if (p[d] == 0) expect |= (mask >> (3 * (2 * d))) & 7u;
            if (p[d] == hi[d]) expect |= (mask >> (3 * (2 * d + 1))) & 7u;
          }
          if ((L.h_nodes_u[(size_t)c * 27 + n] >> NODE_FLAG_SHIFT) != expect) { good = 0.0; break; }
        }
    }
    double *d = nullptr;
    SMF_CUDA(cudaMalloc(&d, sizeof(double)));
    SMF_CUDA(cudaMemcpyAsync(d, &good, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    TRY(allreduce_sum(ctx, d, 1));
    SMF_CUDA(cudaMemcpyAsync(&good, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SMF_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    if (good != (double)ctx->world) return SMF_OK;
  }
  // ---- the child context: global meshes of levels 0..top ----
  Replica *R = new Replica();
  ctx->replica = R;
  R->top = top;
  R->lv.resize(top + 1);
  smf_config cfg{3, ctx->cfg.degree, top + 1, ctx->cfg.device};
  if (smf_create(&cfg, &R->child) != SMF_OK) { replica_free(ctx); return SMF_OK; }
  smf_set_stream(R->child, ctx->stream);
  smf_set_option(R->child, SMF_OPT_CUDA_GRAPHS, ctx->use_graphs);
  for (int pass = 0; pass < 2; ++pass)        // levels, then the child's "active" mesh = level top
    for (int l = pass == 0 ? 0 : top; l <= top; ++l) {
      ReplicaLevel &G = R->lv[l];
      const LatLevel &T = ctx->lat[l];
      G.gcx = T.gnx - 1; G.gcy = T.gny - 1; G.gcz = T.gnz - 1;
      G.n_cells_global = (int64_t)G.gcx * G.gcy * G.gcz;
      const int GX = 2 * G.gcx + 1, GY = 2 * G.gcy + 1, GZ = 2 * G.gcz + 1, PX = G.gcx + 1, PY = G.gcy + 1;
      std::vector<uint32_t> cu((size_t)G.n_cells_global * 81), cp((size_t)G.n_cells_global * 8), cons;
      for (int z = 0; z < G.gcz; ++z)
        for (int y = 0; y < G.gcy; ++y)
          for (int x = 0; x < G.gcx; ++x) {
            const int64_t c = child_cell_index(G, x, y, z);
            for (int n = 0; n < 27; ++n) {
              const int64_t node = ((int64_t)(2 * z + n / 9) * GY + (2 * y + (n / 3) % 3)) * GX + (2 * x + n % 3);
              for (int k = 0; k < 3; ++k) cu[(size_t)c * 81 + k * 27 + n] = (uint32_t)(3 * node + k);
            }
            for (int n = 0; n < 8; ++n)
              cp[(size_t)c * 8 + n] = (uint32_t)(((int64_t)(z + (n >> 2)) * PY + (y + ((n >> 1) & 1))) * PX + (x + (n & 1)));
          }
      for (int Z = 0; Z < GZ; ++Z)
        for (int Y = 0; Y < GY; ++Y)
          for (int X = 0; X < GX; ++X) {
            const int p[3] = {X, Y, Z}, hi[3] = {GX - 1, GY - 1, GZ - 1};
            int f = 0;
            for (int d = 0; d < 3; ++d) {
              if (p[d] == 0) f |= (mask >> (3 * (2 * d))) & 7;
              if (p[d] == hi[d]) f |= (mask >> (3 * (2 * d + 1))) & 7;
            }
            for (int k = 0; k < 3; ++k) if ((f >> k) & 1) cons.push_back((uint32_t)(3 * (((int64_t)Z * GY + Y) * GX + X) + k));
          }
      smf_level_mesh m{};
      m.n_cells = G.n_cells_global; m.n_dofs_u = (int64_t)3 * GX * GY * GZ; m.n_dofs_p = (int64_t)PX * PY * (G.gcz + 1);
      m.cell_dofs_u = cu.data(); m.cell_dofs_p = cp.data();
      m.n_constrained_u = (int64_t)cons.size(); m.constrained_u = cons.data();
      m.geometry = 0; m.cell_size = ctx->levels[l].h;
      if (smf_set_level(R->child, pass == 0 ? l : SMF_ACTIVE, &m) != SMF_OK) { ctx->err = std::string("replicated level: ") + smf_last_error(R->child); replica_free(ctx); return SMF_ERR_INVALID; }
      if (pass == 1) { R->n_global = m.n_dofs_u; break; }
      // local cell -> child cell
      std::vector<uint32_t> cm((size_t)ctx->levels[l].n_cells);
      for (int64_t c = 0; c < ctx->levels[l].n_cells; ++c)
        cm[c] = (uint32_t)child_cell_index(G, cell_xyz[l][c * 3] + T.ox, cell_xyz[l][c * 3 + 1] + T.oy, cell_xyz[l][c * 3 + 2] + T.oz);
      SMF_CUDA(cudaMalloc(&G.cellmap, sizeof(uint32_t) * cm.size()));
      SMF_CUDA(cudaMemcpy(G.cellmap, cm.data(), sizeof(uint32_t) * cm.size(), cudaMemcpyHostToDevice));
      SMF_CUDA(cudaMalloc(&G.geta, sizeof(double) * G.n_cells_global));
    }
  for (int l = 1; l <= top; ++l) {
    const ReplicaLevel &C = R->lv[l - 1], &F = R->lv[l];
    std::vector<int32_t> ch((size_t)C.n_cells_global * 8);
    for (int z = 0; z < C.gcz; ++z)
      for (int y = 0; y < C.gcy; ++y)
        for (int x = 0; x < C.gcx; ++x)
          for (int k = 0; k < 8; ++k)
            ch[(size_t)child_cell_index(C, x, y, z) * 8 + k] = (int32_t)child_cell_index(F, 2 * x + (k & 1), 2 * y + ((k >> 1) & 1), 2 * z + (k >> 2));
    if (smf_set_transfer(R->child, l, ch.data()) != SMF_OK) { ctx->err = std::string("replicated transfer: ") + smf_last_error(R->child); replica_free(ctx); return SMF_ERR_INVALID; }
  }
  smf_set_copy_indices(R->child, nullptr, nullptr);
  LevelDev &LT = ctx->levels[top];
  R->n_local = LT.n_u;
  SMF_CUDA(cudaMalloc(&R->gb, sizeof(double) * R->n_global)); SMF_CUDA(cudaMalloc(&R->gx, sizeof(double) * R->n_global));
  SMF_CUDA(cudaMalloc(&R->not_owned, LT.n_u));
  std::vector<unsigned char> no((size_t)LT.n_u, 0);
  for (int64_t i = 0; i < LT.n_u && i < (int64_t)LT.h_not_owned_u.size(); ++i) no[i] = LT.h_not_owned_u[i];
  SMF_CUDA(cudaMemcpy(R->not_owned, no.data(), no.size(), cudaMemcpyHostToDevice));
  ctx->bytes += 2 * sizeof(double) * R->n_global;
  ctx->replica_top = top;
  return SMF_OK;
}

// called by smf_setup_smoothers (after lattice_prepare): build once per mesh, then refresh tables and smoother parameters
int replica_prepare(smf_ctx *ctx)
{
  if (!ctx->use_replica || ctx->world == 1) return SMF_OK;
  if (!ctx->replica_tried) {
    ctx->replica_tried = true;
    TRY(replica_build(ctx));
  }
  Replica *R = (Replica *)ctx->replica;
  if (!R || ctx->replica_top < 0) return SMF_OK;
  for (int l = 0; l <= R->top; ++l) {
    LevelDev &L = ctx->levels[l];
    ReplicaLevel &G = R->lv[l];
    SMF_CUDA(cudaMemsetAsync(G.geta, 0, sizeof(double) * G.n_cells_global, ctx->stream));
    k_rep_eta<<<(unsigned)((L.n_cells + 255) / 256), 256, 0, ctx->stream>>>(G.geta, L.eta, G.cellmap, L.n_cells);
    TRY(allreduce_sum(ctx, G.geta, (int)G.n_cells_global));      // every cell belongs to one rank: a gather
    R->h_eta.resize((size_t)G.n_cells_global);
    SMF_CUDA(cudaMemcpyAsync(R->h_eta.data(), G.geta, sizeof(double) * G.n_cells_global, cudaMemcpyDeviceToHost, ctx->stream));
    SMF_CUDA(cudaStreamSynchronize(ctx->stream));
    smf_cell_data cd{R->h_eta.data(), 1, L.ps, L.compressible};
    if (smf_set_cell_data(R->child, l, &cd) != SMF_OK || (l == R->top && smf_set_cell_data(R->child, SMF_ACTIVE, &cd) != SMF_OK))
      return ctx->fail(SMF_ERR_INVALID, std::string("replicated cell data: ") + smf_last_error(R->child));
  }
  if (smf_compute_diagonals(R->child) != SMF_OK) return ctx->fail(SMF_ERR_CUDA, std::string("replicated diagonals: ") + smf_last_error(R->child));
  if (!R->child_ready) {      // first time: vectors, graphs and the child's own estimates
    if (smf_setup_smoothers(R->child) != SMF_OK) return ctx->fail(SMF_ERR_CUDA, std::string("replicated smoothers: ") + smf_last_error(R->child));
    R->child_ready = true;
  }
  // the parent's eigenvalue estimates are those of the same global matrices (distributed operator): the replicated levels
  // smooth with exactly the parameters the distributed levels would have used
  for (int l = 0; l <= R->top; ++l) R->child->smoother_A[l] = ctx->smoother_A[l];
  R->child->smoothers_ready = true;
  R->child->graph_A_stale = true;
  return SMF_OK;
}

// coarse correction of the A-block cycle below level top + 1: b = restricted defect of level `top` on this rank (complete on
// shared DoFs), x <- this rank's part of the global V-cycle's result
int replica_cycle(smf_ctx *ctx, const double *b, double *x)
{
  Replica *R = (Replica *)ctx->replica;
  const LevelDev &L = ctx->levels[R->top];
  const unsigned nb = (unsigned)((R->n_local + 255) / 256);
  {
    ProfScope prof(ctx, B_TRANSFER);
    SMF_CUDA(cudaMemsetAsync(R->gb, 0, sizeof(double) * R->n_global, ctx->stream));
    k_rep_scatter_owned<<<nb, 256, 0, ctx->stream>>>(R->gb, b, L.ex_u.gid, R->not_owned, R->n_local);
    ctx->launches++;
  }
  TRY(allreduce_sum(ctx, R->gb, (int)R->n_global));
  const int64_t l0 = smf_kernel_launches(R->child);
  if (smf_vcycle_A(R->child, R->gx, R->gb) != SMF_OK) return ctx->fail(SMF_ERR_CUDA, std::string("replicated V-cycle: ") + smf_last_error(R->child));
  ctx->launches += smf_kernel_launches(R->child) - l0;
  {
    ProfScope prof(ctx, B_TRANSFER);
    k_rep_gather<<<nb, 256, 0, ctx->stream>>>(x, R->gx, L.ex_u.gid, R->not_owned, R->n_local);
    ctx->launches++;
  }
  SMF_CUDA(cudaGetLastError());
  // The child's cell kernels accumulate with unordered REDs, so two ranks' copies of the global correction differ in the last bits;
  // the copies of a shared DoF must stay bitwise equal, so every rank takes the owner's value: the others contributed zero above.
  return exchange_add(ctx, ctx->levels[R->top].ex_u, x);
}

}  // namespace smf
