// Internal declarations of the B200 matrix-free GMG Stokes library (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/smf_b200.h"

namespace smf {

constexpr int BATCH_CELLS = 32;   // cells staged per thread block (a Morton 4x4x2 / 8x4 brick)
constexpr int SUB = 32;           // cells per sub-batch = lanes of a warp
constexpr int KSUB = BATCH_CELLS / SUB;
constexpr uint32_t NODE_IDX_MASK = 0x0fffffffu;   // low 28 bits: DoF index of component 0 of the node
constexpr uint32_t NODE_SHARED_BIT = 1u << 28;    // structured bricks: the node is also touched by another batch
constexpr uint32_t NODE_INVALID = 0xffffffffu;    // structured bricks: padding slot of the staging layout
constexpr uint32_t NODE_FLAG_SHIFT = 29;          // bits 29..31: component c of this node is constrained

// ---- structured bricks (3D) ------------------------------------------------------------------------------
// When the 32 cells of a batch form a 4x4x2 brick in Morton order (always the case on uniformly refined levels),
// its 9x9x5 node lattice is staged in shared memory in a layout that needs no per-cell slot table and is free of
// bank conflicts: nodes are grouped by the parity of their lattice coordinates (X, Y, Z) -- for a fixed local node
// of the cell, all 32 cells hit the same parity class at (x, y, z) + const -- and inside a class the strides are
// chosen so that the 16 lanes of a half-warp (a 64-bit shared-memory access is served one half-warp at a time; in
// Morton order the lane bits are x0 y0 z0 x1 | y1) fall into 16 different 8-byte banks:
//   offset = ox * 1 + oy * SY + oz * SZ  with (SY, SZ) mod 16 = (8, 4) or (4, 8)  ->  bank bits x0:1 x1:2 y0:SY z0:SZ
//   class (X even, Y even): SY  8, SZ 52   (5 x 5 nodes per plane)
//   class (X odd,  Y even): SY  4, SZ 24   (4 x 5)
//   class (X even, Y odd ): SY  8, SZ 36   (5 x 4)
//   class (X odd,  Y odd ): SY  4, SZ 24   (4 x 4)
__host__ __device__ constexpr int sb_variant(int px, int py) { return px + 2 * py; }
__host__ __device__ constexpr int sb_sy(int v) { return (v & 1) ? 4 : 8; }
__host__ __device__ constexpr int sb_sz(int v) { return v == 0 ? 52 : (v == 2 ? 36 : 24); }
__host__ __device__ constexpr int sb_nx(int v) { return (v & 1) ? 4 : 5; }
__host__ __device__ constexpr int sb_ny(int v) { return (v & 2) ? 4 : 5; }
__host__ __device__ constexpr int sb_class_size(int v, int pz)
{ return sb_sz(v) * (pz ? 1 : 2) + sb_sy(v) * (sb_ny(v) - 1) + sb_nx(v); }
__host__ __device__ constexpr int sb_class_base(int v, int pz)
{
  int b = 0;
  for (int z = 0; z < 2; ++z)
    for (int w = 0; w < 4; ++w) {
      if (z == pz && w == v) return b;
      b += sb_class_size(w, z);
    }
  return b;
}
__host__ __device__ constexpr int sb_offset(int v, int ox, int oy, int oz) { return ox + sb_sy(v) * oy + sb_sz(v) * oz; }
// slot of lattice point (X, Y, Z), 0 <= X, Y < 9, 0 <= Z < 5
__host__ __device__ constexpr int sb_slot(int X, int Y, int Z)
{ return sb_class_base(sb_variant(X & 1, Y & 1), Z & 1) + sb_offset(sb_variant(X & 1, Y & 1), X >> 1, Y >> 1, Z >> 1); }
constexpr int SB_SLOTS = 612;   // = sb_class_base(3, 1) + sb_class_size(3, 1)
constexpr int SB_LIST = 408;    // node-list entries per structured brick (9 * 9 * 5 = 405, padded)
constexpr int SB_LIST_BYTES = SB_LIST * 6;   // pipelined kernel: per brick [SB_LIST] uint32 entries, then [SB_LIST] uint16 slots
constexpr int SB_PLIST = 76;                 // pressure nodes of a brick in lattice order (5 x 5 x 3 = 75, padded)
constexpr int SB_BLOCK_BYTES = SB_LIST_BYTES + SB_PLIST * 4;   // stride of a brick in LevelDev::sb_lists
static_assert(SB_LIST_BYTES % 16 == 0 && SB_BLOCK_BYTES % 16 == 0, "node lists are copied in 16-byte chunks");
static_assert(sb_class_base(3, 1) + sb_class_size(3, 1) == SB_SLOTS, "structured brick layout size");
// brick coordinates of lane l (Morton order inside the batch: x0 y0 z0 x1 y1)
__host__ __device__ constexpr int sb_x(int l) { return (l & 1) | (((l >> 3) & 1) << 1); }
__host__ __device__ constexpr int sb_y(int l) { return ((l >> 1) & 1) | (((l >> 4) & 1) << 1); }
__host__ __device__ constexpr int sb_z(int l) { return (l >> 2) & 1; }

// per-batch header (ints): n_nodes_u, n_excl_u, n_nodes_p, n_excl_p, n_cells, pad x3
constexpr int BATCH_INFO = 8;

// halo exchange of one level and DoF family (comm.cu)
struct ExchangeDev {
  bool set = false;
  int ncomp = 1;
  std::vector<int> ranks;                     // neighbour ranks, ascending
  std::vector<int64_t> counts, offsets;       // shared nodes per neighbour / start in the buffers (in nodes)
  int64_t n_send = 0;                         // total shared-node entries (= sum of counts)
  uint32_t *send_idx = nullptr;               // [n_send] first DoF of each shared node
  double *send_buf = nullptr, *recv_buf = nullptr;   // [n_send * ncomp]
  int64_t n_if = 0;                           // distinct interface nodes
  uint32_t *if_idx = nullptr; int *if_ptr = nullptr, *if_src = nullptr;   // rank-ordered accumulation tables
  double *weights = nullptr;                  // [n_dofs] 1 / (ranks sharing the DoF); nullptr on one rank
  int64_t *gid = nullptr; double gid_mean = 0.0;     // global DoF numbers (Chebyshev start vector)
  std::vector<int64_t> h_gid;                        // host copy (pressure: schur_lattice.cu locates the rank's box in the global lattice; velocity: levels up to 8 M DoFs, velocity_replica.cu)
  // fused peer-memory transport (comm.cu): neighbour table, {epoch, counters}, this rank's inbox (2 parity buffers)
  void *p2p_tab = nullptr, *p2p_state = nullptr; double *p2p_inbox = nullptr;
};

// Newton / prescribed-dilation tables of OperatorCellData (active level only; matrix_free_operators.h:128-196)
struct NewtonDev {
  int newton = 0, symmetrize = 0, average = 0, dilation = 0;
  double *eps = nullptr, *deta = nullptr, *detadp = nullptr, *dil_lhs = nullptr, *ddil = nullptr, *ddildp = nullptr;
};

// Per-cell tables read by one thread per cell (curved-cell geometry, transposed cell->DoF maps) are stored in blocks of
// GEO_BLOCK cells: [cell / 64][row][cell % 64].  Consecutive threads read consecutive words, and everything a thread
// block needs is one contiguous chunk (n_rows x 512 B) instead of n_rows rows that lie n_cells words apart (which costs
// a DRAM page and a TLB entry per row).
constexpr int GEO_BLOCK = 64;
__host__ __device__ inline int64_t blk_index(int64_t cell, int row, int n_rows)
{ return ((cell / GEO_BLOCK) * n_rows + row) * GEO_BLOCK + (cell % GEO_BLOCK); }
__host__ __device__ inline int64_t blk_cells(int64_t n_cells) { return (n_cells + GEO_BLOCK - 1) / GEO_BLOCK * GEO_BLOCK; }

constexpr int MAX_DEVICES = 64;        // per-device caches of function attributes / occupancy (indexed by the device ordinal)
constexpr int SMF_MAX_RESTART = 200;   // longest GMRES / FGMRES restart length (nsinker_spherical_shell/gmg.prm uses 200); sizes the multi-dot scratch

// closed constraint lines u_k = sum_j w_j u_{m_j} (hanging nodes) of one DoF family, CSR on the device (adaptive.cu)
struct ConstraintLines {
  int64_t n = 0;
  uint32_t *dof = nullptr, *start = nullptr, *master = nullptr;
  double *weight = nullptr, *save = nullptr;       // save: [n] the constrained entries of src while the cell loop runs
};
// refinement-edge indices of one level and family + scratch
struct EdgeSet { int64_t n = 0; uint32_t *idx = nullptr; double *save = nullptr; };
// free-surface boundary faces of the active level (adaptive.cu: free_surface_faces_apply)
struct FaceSet { int64_t n = 0; uint32_t *cells = nullptr; int32_t *face_no = nullptr; double *table = nullptr; };
struct CopyPairs { int64_t n = 0; uint32_t *active = nullptr, *level = nullptr; };

// Lattice-ordered mirror of the velocity vectors of a structured box level: [component][x parity][Z][Y][X / 2], row pitch Xp
// (even, so that every row starts 16-byte aligned), for the kernel that stages its bricks with TMA box loads (velocity_lat_kernel)
struct VLat {
  int nx = 0, ny = 0, nz = 0, Xp = 0;
  int64_t n_nodes = 0, n_lat = 0;           // nx * ny * nz;  6 * nz * ny * Xp
  uint32_t *lat2dof = nullptr;              // [n_nodes] first DoF of every lattice node (host numbering)
  int *brick_org = nullptr;                 // [n_batches][4]: lattice origin of the brick, bit d / 3 + d: on the low / high domain face of direction d
  int cons_mask = 0;                        // bit 3 f + c: component c is constrained on boundary face f = 2 d + side
  double *src = nullptr, *dst = nullptr;    // lattice-ordered copies of the operand / the result
};

struct LevelDev {
  int dim = 0;
  int64_t n_cells = 0, n_u = 0, n_p = 0;
  int n_batches = 0;
  int max_nodes_u = 0, max_nodes_p = 0;      // padded staging sizes (nodes per batch)
  int safe_rounds = 1;                        // 1: scatter rounds are collision-free inside every sub-batch
  int structured = 0;                         // 1: every batch is a structured brick (velocity staging by sb_slot)
  int *batch_info = nullptr;                  // [n_batches][BATCH_INFO]
  uint32_t *node_list_u = nullptr;            // [n_batches][list_stride_u]  ascending DoF index (general: exclusive nodes first)
  uint16_t *node_slot_u = nullptr;            // [n_batches][list_stride_u]  structured bricks: staging slot of each list entry
  int list_stride_u = 0;                      // entries per batch in node_list_u (general path: = max_nodes_u)
  unsigned char *sb_lists = nullptr;          // structured bricks: [n_batches][SB_BLOCK_BYTES] velocity entries + slots, pressure entries (velocity_sb_kernel)
  uint32_t *node_list_p = nullptr;            // [n_batches][max_nodes_p]
  uint16_t *cell_loc_u = nullptr;             // [n_batches][KSUB][3^dim][SUB]
  uint16_t *cell_loc_p = nullptr;             // [n_batches][KSUB][2^dim][SUB]
  uint32_t *cell_nodes_u = nullptr;           // [n_cells][3^dim]   (index | flags), raw map
  uint32_t *cell_nodes_p = nullptr;           // [n_cells][2^dim]
  uint32_t *constrained_u = nullptr; int64_t n_constrained_u = 0;
  uint32_t *constrained_p = nullptr; int64_t n_constrained_p = 0;
  double h[3] = {1, 1, 1};
  uint32_t *cell_nodes_u_t = nullptr, *cell_nodes_p_t = nullptr;   // curved cells: transposed maps, rows = nodes of a blk_index table (one thread per cell reads coalesced)
  double *geo = nullptr;                      // curved cells (geometry = 1): inverse Jacobian (r-major) and JxW, rows q * (dim*dim + 1) + k of a blk_index table
  // cell data
  double *eta = nullptr; int n_eta = 0; double ps = 1.0; int compressible = 0; bool has_cell_data = false;
  NewtonDev nd;
  // inverse diagonals
  double *dinv_A = nullptr, *dinv_M = nullptr; bool has_diag = false;
  int64_t bytes = 0;
  std::vector<uint32_t> h_nodes_u, h_nodes_p;   // host copies of the raw maps (needed by smf_set_transfer)
  std::vector<uint32_t> h_sb_list;              // host copy of the structured velocity node lists [n_batches][SB_LIST]
  std::vector<uint8_t> h_cons_u, h_cons_p;      // host constraint bitmaps (needed by smf_set_exchange)
  std::vector<uint8_t> h_not_owned_u, h_not_owned_p;   // DoFs whose owner is a lower rank (set by smf_set_exchange)
  ExchangeDev ex_u, ex_p;
  // adaptive meshes (adaptive.cu)
  ConstraintLines hang_u, hang_p;               // ACTIVE level: hanging-node lines
  EdgeSet edge_u, edge_p;                       // levels: refinement-edge indices
  CopyPairs copy_u, copy_p;                     // levels: copy_to_mg / copy_from_mg pairs
  FaceSet fs_faces;                             // ACTIVE level: free-surface stabilisation faces
  VLat *vlat = nullptr;                         // SMF_OPT_LATTICE_VELOCITY: lattice-ordered mirror (velocity_lattice.cu)
};

constexpr int TB_COARSE = 76;      // coarse velocity nodes under a fine brick: 5 x 5 x 3 lattice = 75, padded
constexpr int TB_BYTES = TB_COARSE * 4 + SB_LIST + 8;   // + one owner byte per entry of the fine brick's node list (+ padding: 720 bytes)
static_assert(TB_BYTES % 16 == 0, "transfer blocks are read in 16-byte chunks");
struct TransferDev {              // level l-1 -> l
  int64_t n_parents = 0;
  // structured fine level: per fine brick [TB_COARSE] uint32 coarse node entries (lattice order), then [SB_LIST] bytes
  // in the order of the brick's node list: bit 0 = this brick writes the node in prolongation, bit 1 = contributes it
  // in restriction (velocity transfers, prolongate_sb_kernel / restrict_sb_kernel)
  unsigned char *tb_u = nullptr;
  uint32_t *patch_u = nullptr;    // [n_parents][5^dim] fine node (component-0 DoF) | bit31 = this parent owns the node
  uint32_t *patch_p = nullptr;    // [n_parents][3^dim]
  int32_t *children = nullptr;    // [n_parents][2^dim] fine cells of every coarse cell (device copy)
  std::vector<int32_t> h_children;   // host copy (the lattice detection of schur_lattice.cu checks the child order against it)
  bool set = false;
};

// Lattice view of one level of the Schur-complement V-cycle (schur_lattice.cu): the pressure nodes of a Cartesian box level
// form an nx x ny x nz lattice; inside the V-cycle the level vectors are stored in lattice order (x fastest), so the
// pressure mass matrix is a 27-point node-centred stencil with one coefficient per cell -- no atomics, a fixed summation
// order, and the Chebyshev update fused into the same kernel.
struct LatLevel {
  int nx = 0, ny = 0, nz = 0;      // pressure nodes per direction (cells: nx-1, ny-1, nz-1)
  int64_t n = 0;                   // nx * ny * nz = n_p of the level
  int ox = 0, oy = 0, oz = 0;      // several ranks: position of this rank's box in the global lattice of the level
  int gnx = 0, gny = 0, gnz = 0;   // ... and the size of that lattice (0: unknown)
  uint32_t *dof2lat = nullptr;     // [n] lattice index of every pressure DoF of the level
  uint32_t *cell2lat = nullptr;    // [n_cells] lattice index of every cell
  double *coef = nullptr;          // [cells] ps^2 vol / (216 eta), lattice order
  double *dinv = nullptr;          // [n] inverse diagonal, lattice order
  double *x = nullptr, *xo = nullptr, *b = nullptr, *t = nullptr;
  // several ranks: interface nodes keep partial sums until the exchange; bit 0 = shared with another rank, bit 1 = owned by a lower rank
  unsigned char *mask = nullptr;
  uint32_t *if_nodes = nullptr; int64_t n_if = 0;    // lattice indices of the distinct interface nodes
  ExchangeDev ex;                                    // twin of LevelDev::ex_p with lattice indices (owns send_idx / if_idx only)
};

struct Smoother {                 // PreconditionChebyshev state of one level
  int degree = 4;
  double smoothing_range = 15.0;
  int eig_cg_n_iterations = 10;
  double lambda_min = 0, lambda_max = 0, theta = 1, delta = 0;
  int cg_its = 0;
};

// device-time buckets of the optional profiler (smf_profile_*): the split SURVEY.md 8(d) asks for
enum Bucket { B_STOKES = 0, B_A_FINE, B_A_COARSE, B_MASS, B_BT, B_TRANSFER, B_SMOOTH_VEC, B_ORTHO, B_VEC, B_EXCHANGE, B_COUNT };
struct ProfEvent { int bucket; cudaEvent_t a, b; };

struct MGVectors {                // per level: solution, defect, t, and a Chebyshev scratch
  double *x = nullptr, *b = nullptr, *t = nullptr, *xo = nullptr;
  double *e = nullptr;            // adaptive meshes: the edge part of the iterate (vmult_interface_up), allocated on demand
  double *res = nullptr;          // where the last level_v_step left the solution (x or xo)
  bool t_zero = false;            // t is known to be all zero (left so by cheb_update): the next operator application skips its memset
};

}  // namespace smf

struct smf_ctx {
  smf_config cfg{};
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  std::vector<smf::LevelDev> levels;     // index 0..n_levels-1, plus [n_levels] = active
  std::vector<smf::TransferDev> transfers;   // index l = transfer (l-1 -> l)
  std::vector<smf::Smoother> smoother_A, smoother_S;
  std::vector<smf::MGVectors> mgv_A, mgv_S;
  uint32_t *copy_u = nullptr, *copy_p = nullptr;      // active -> finest level index maps (nullptr = identity)
  bool smoothers_ready = false;
  bool adaptive = false;                              // per-level copy indices set: adaptive V-cycle (solver.cu)
  int64_t launches = 0;
  int64_t bytes = 0;
  double *scratch = nullptr; int64_t scratch_n = 0;   // reduction scratch
  double *host_pinned = nullptr; int64_t host_pinned_n = 0;
  // Krylov workspace
  std::vector<double *> krylov; int64_t krylov_n = 0;
  double *tmp_u = nullptr, *tmp_p = nullptr, *tmp_u2 = nullptr, *tmp_p2 = nullptr;
  double *io_buf = nullptr; int64_t io_n = 0;         // device staging of smf_solve_host
  cudaStream_t copy_stream = nullptr; cudaEvent_t h2d_event = nullptr; bool h2d_pending = false;   // smf_solve_host: the H2D copies overlap the smoother set-up
  int n_A_prec = 0, n_S_prec = 0;
  int n_A_exp = 0, n_S_exp = 0;                       // inner CG iterations of the expensive phase
  double *pnorm_weights = nullptr; double pnorm_area = 0.0;   // pressure normalisation: integral of every pressure shape function over the surface / volume
  double *bicg_buf = nullptr;                         // [7][n_u] scratch of the inner BiCGStab solve
  int force_nonsymmetric_A = 0;                       // `Force nonsymmetric A block solver` (SMF_OPT_NONSYMMETRIC_A_SOLVER)
  double *cg_buf = nullptr;                           // [3][n_u] scratch of the inner CG solves
  double *idr_normals = nullptr; int64_t idr_normals_n = 0;   // IDR(s) shadow-vector entries (kept between solves)
  // distributed run (comm.cu): one subdomain per rank
  int rank = 0, world = 1;
  void *comm = nullptr;                               // ncclComm_t
  void *p2p = nullptr;                                // P2PState (comm.cu): inbox arena + mapped peer arenas; nullptr = NCCL transport
  int use_p2p = 1;                                    // fused peer-memory halo exchange when the IPC mapping is available
  int64_t exchanges = 0;
  const double *dot_weights = nullptr;                // weights the inner products use right now (nullptr: none)
  double *block_weights = nullptr;                    // [n_u + n_p] of the active level
  int md_vec2 = 0;                                    // the pointer table of the last multi_dot allows 16-byte loads
  // CUDA graphs of the coarse part of the two V-cycles (level_v_step(graph_level) on the fixed MG buffers)
  void *graph_A = nullptr, *graph_S = nullptr;        // cudaGraphExec_t
  bool graph_A_stale = true, graph_S_stale = true;    // smoother coefficients changed since the capture
  int64_t graph_A_launches = 0, graph_S_launches = 0; // kernels inside one replay
  int64_t graph_A_exchanges = 0, graph_S_exchanges = 0;   // halo exchanges inside one replay
  int graph_level = -1;                               // finest level inside the graphs (-1: graphs unused)
  double *ad_in_u = nullptr, *ad_out_u = nullptr, *ad_in_p = nullptr, *ad_out_p = nullptr;   // adaptive meshes: staging of the graph-replayed V-cycles
  bool ad_warm_u = false, ad_warm_p = false;          // the first adaptive cycle after a mesh change runs eagerly (allocations)
  int use_graphs = 1;
  int ortho_single_pass = 1;                          // GMRES: re-orthogonalise through the Gram matrix (1) or with a second pass over the basis (0)
  int max_grid_blocks = 0;                            // > 0: cap on the persistent grids (tests: forces several bricks per block)
  int use_curved_tma = 0;                             // curved cells: geometry staged by cp.async.bulk + mbarrier (SMF_OPT_CURVED_TMA)
  int use_curved_fast = 1;                            // curved cells: sum-factorised one-thread-per-cell kernel (0: the general restatement kernel)
  int use_coarse_tail = 1;                            // Schur V-cycle: levels 0..tail_level in one single-block kernel (coarse_tail.cu)
  int tail_max_cells = 1024;                          // largest level (cells) the coarse tail takes (SMF_OPT_COARSE_TAIL value > 1 overrides)
  int tail_level = -1;                                // set by smf_setup_smoothers (-1: not applicable)
  // lattice path of the Schur V-cycle (schur_lattice.cu)
  std::vector<smf::LatLevel> lat;                     // per level; empty = not built
  int lat_state = 0;                                  // 0 unknown, 1 usable, -1 the hierarchy is not a box lattice
  int use_lattice = 1;                                // SMF_OPT_LATTICE_SCHUR
  uint32_t *lat_act = nullptr;                        // [n_p active] lattice index on the finest level of every active pressure DoF
  std::vector<smf::LatLevel> glat;                    // several ranks: levels 0..lat_repl as GLOBAL lattices, computed redundantly by every rank
  int lat_repl = -1;                                  // (no halo exchange on those levels; one all-reduce of the restricted defect per cycle)
  std::vector<void *> graph_L_seg;                    // cudaGraphExec_t of the segments of the lattice V-cycle (split at the all-reduce)
  bool lat_capturing = false;
  void *lat_capture_segs = nullptr;                   // std::vector<cudaGraph_t> * of the capture in progress
  void *graph_L = nullptr; bool graph_L_stale = true; // (unused: kept for layout) 
  int64_t graph_L_launches = 0, graph_L_exchanges = 0;
  int lat_tail = -1;                                  // levels 0..lat_tail of the lattice cycle run in one thread block
  double *lat_res = nullptr;                          // buffer the captured cycle leaves its result in
  void *replica = nullptr;                            // velocity_replica.cu: replicated coarse levels of the A-block cycle (child context)
  int replica_top = -1;                               // levels 0..replica_top of prec_A are computed redundantly on the global mesh (-1: none)
  bool replica_tried = false;
  int use_replica = 1;                                // SMF_OPT_REPLICATE_COARSE_A
  int use_vlat = 0;                                   // SMF_OPT_LATTICE_VELOCITY: A block of structured box levels through the TMA-staged lattice kernel
  int use_warp_brick = 0;                             // structured levels: one-warp-per-brick velocity kernel, three components per thread (SMF_OPT_WARP_BRICK_KERNEL)
  int use_pipelined = 1;                              // structured levels: persistent cp.async-pipelined velocity kernel (tests may switch it off)
  int allow_structured = 1;                           // structured-brick fast path of the velocity kernels (tests may switch it off)
  // profiler: 0 off, 1 operator kernels of the finest level only, 2 every launch
  int prof_mode = 0;
  int vec_bucket = smf::B_VEC;                        // bucket the vector kernels are booked under right now
  std::vector<smf::ProfEvent> prof_live;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[smf::B_COUNT] = {0};
  int64_t prof_cnt[smf::B_COUNT] = {0};

  smf::LevelDev &lev(int level) { return levels[level < 0 ? cfg.n_levels : level]; }
  int fail(int code, const std::string &msg) { err = msg; return code; }
};

#define SMF_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ctx->fail(SMF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));     \
  } while (0)

namespace smf {
// RAII pair of CUDA events around one launch on the context's stream (no-op unless profiling is on)
struct ProfScope {
  smf_ctx *ctx; cudaEvent_t b = nullptr;
  ProfScope(smf_ctx *c, int bucket) : ctx(c)
  {
    if (c->prof_mode == 0 || (c->prof_mode == 1 && bucket != B_STOKES && bucket != B_A_FINE)) return;
    cudaEvent_t a;
    auto get = [&](cudaEvent_t &e) {
      if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
      else cudaEventCreate(&e);
    };
    get(a); get(b);
    cudaEventRecord(a, c->stream);
    c->prof_live.push_back({bucket, a, b});
  }
  ~ProfScope() { if (b) cudaEventRecord(b, ctx->stream); }
};
struct DotWeights {     // inner products in this scope weight every DoF with 1 / (number of ranks sharing it)
  smf_ctx *ctx; const double *saved;
  DotWeights(smf_ctx *c, const double *w) : ctx(c), saved(c->dot_weights) { c->dot_weights = w; }
  ~DotWeights() { ctx->dot_weights = saved; }
};
struct VecBucket {      // books the vector kernels launched in this scope under `bucket`
  smf_ctx *ctx; int saved;
  VecBucket(smf_ctx *c, int bucket) : ctx(c), saved(c->vec_bucket) { c->vec_bucket = bucket; }
  ~VecBucket() { ctx->vec_bucket = saved; }
};
// operators.cu
int launch_vmult_A(smf_ctx *ctx, LevelDev &L, double *dst, const double *src, bool dst_is_zero = false);
int launch_vmult_stokes(smf_ctx *ctx, LevelDev &L, double *dst_u, double *dst_p, const double *src_u, const double *src_p, bool plain = false);
int launch_vmult_BT(smf_ctx *ctx, LevelDev &L, double *dst_u, const double *src_p);
int launch_vmult_mass(smf_ctx *ctx, LevelDev &L, double *dst, const double *src, bool dst_is_zero = false);
int launch_diag(smf_ctx *ctx, LevelDev &L);
int launch_set_entries(smf_ctx *ctx, double *vec, const uint32_t *idx, int64_t n, double value);
int launch_rhs(smf_ctx *ctx, LevelDev &L, const double *force_q_dev, double *rhs_u);
int launch_prolongate(smf_ctx *ctx, int which, int to_level, double *dst, const double *src, bool add);
int launch_restrict_add(smf_ctx *ctx, int which, int from_level, double *dst, const double *src);
int launch_restrict_add_residual(smf_ctx *ctx, int which, int from_level, double *dst, const double *b, double *t, int64_t n_fine);   // dst += R (b - t)
// adaptive.cu: hanging-node lines around a cell loop; refinement-edge handling of the level operators
int hanging_pre(smf_ctx *ctx, ConstraintLines &H, const double *src);                // save src[k]; src[k] = sum_j w_j src[m_j]
int hanging_post(smf_ctx *ctx, ConstraintLines &H, double *dst, const double *src, bool unit_diagonal);   // dst[m_j] += w_j dst[k]; dst[k] = saved (or 0); src[k] = saved
int hanging_restore(smf_ctx *ctx, ConstraintLines &H, const double *src);                                                   // src[k] = saved
int edge_pre(smf_ctx *ctx, EdgeSet &E, const double *src);                           // save src[e]; src[e] = 0
int edge_post(smf_ctx *ctx, EdgeSet &E, double *dst, const double *src, bool add);   // dst[e] = saved (add: += saved); src[e] = saved
int edge_extract(smf_ctx *ctx, const EdgeSet &E, double *xe, const double *x, int64_t n);        // xe = E x
int edge_subtract_off(smf_ctx *ctx, const EdgeSet &E, double *b, const double *y, int64_t n);   // b -= (I - E) y
int copy_pairs(smf_ctx *ctx, const CopyPairs &C, double *level_vec, double *active_vec, bool to_level);
int free_surface_faces_apply(smf_ctx *ctx, const LevelDev &L, double *dst_u, const double *src_u, bool plain);    // dst_u -= int_face (g . u)(n . v)
void free_adaptive(LevelDev &L);
// coarse_tail.cu
int schur_tail_top_level(smf_ctx *ctx);
int launch_schur_tail(smf_ctx *ctx, int top);
// velocity_lattice.cu / operators.cu
int vlat_prepare(smf_ctx *ctx, LevelDev &L);                         // builds L.vlat if the level qualifies (else leaves it null)
void vlat_free(LevelDev &L);
int launch_velocity_lat(smf_ctx *ctx, LevelDev &L, double *dst, const double *src);      // A block: dst (zeroed) += A src, both in host numbering
// velocity_replica.cu
int replica_prepare(smf_ctx *ctx);                                   // after lattice_prepare, in smf_setup_smoothers (collective)
int replica_cycle(smf_ctx *ctx, const double *b, double *x);         // coarse correction of prec_A below level replica_top + 1
void replica_free(smf_ctx *ctx);
// schur_lattice.cu
bool lattice_cell_coords(const std::vector<uint32_t> &nodes_p, int64_t n_cells, int64_t n_p, int dims[3], std::vector<int> &cell_xyz);
int lattice_prepare(smf_ctx *ctx);                                   // detection (once per mesh) + coefficients / diagonals (every smoother set-up)
int lattice_vcycle(smf_ctx *ctx, double *dst, const double *src);    // prec_Schur on the lattice
void lattice_free(smf_ctx *ctx);
// comm.cu
void free_exchange(ExchangeDev &E);
int exchange_add(smf_ctx *ctx, ExchangeDev &E, double *vec);
int allreduce_sum(smf_ctx *ctx, double *dev, int count);
int allreduce_minmax(smf_ctx *ctx, double *mn, double *mx);    // host values; collective
int start_vector_mod11(smf_ctx *ctx, const ExchangeDev &E, double *x, int64_t n);
// vector kernels (vecops.cu)
int vec_set(smf_ctx *ctx, double *x, double v, int64_t n);
int vec_shift(smf_ctx *ctx, double *x, double v, int64_t n);                                 // x += v (Vector::add(scalar))
int vec_copy(smf_ctx *ctx, double *y, const double *x, int64_t n);
int vec_scale(smf_ctx *ctx, double *x, double a, int64_t n);
int vec_axpy(smf_ctx *ctx, double *y, double a, const double *x, int64_t n);               // y += a x
int vec_sadd(smf_ctx *ctx, double *y, double s, double a, const double *x, int64_t n);     // y = s y + a x
int vec_pointwise(smf_ctx *ctx, double *z, double a, const double *d, const double *x, int64_t n);   // z = a d.x
int vec_dot(smf_ctx *ctx, const double *x, const double *y, int64_t n, double *host_result);   // synchronises
int vec_norm2(smf_ctx *ctx, const double *x, int64_t n, double *host_result);                  // synchronises
// h_j = V_j . w (j < k) and w.w; results stay on the device (and go to host_out when sync is set)
int vec_multi_dot(smf_ctx *ctx, double *const *V_host_ptrs, int k, const double *w, int64_t n, double *host_out, bool sync);
// as vec_multi_dot (results stay on the device), and V[k-1] is dotted with the basis in the same pass: leaves the
// re-orthogonalised (CGS2-equivalent) coefficients h = h1 - (V^T V - I) h1 on the device for vec_multi_axpy_from_dots
int vec_multi_dot_gram(smf_ctx *ctx, double *const *V_host_ptrs, int k, const double *w, int64_t n);
// w -= sum_j h_j V_j with the h of the preceding vec_multi_dot; host_out = {h_0..h_{k-1}, |w_old|^2, |w_new|^2}
int vec_multi_axpy_from_dots(smf_ctx *ctx, double *w, int k, int64_t n, double *host_out);
int vec_multi_add(smf_ctx *ctx, double *x, double *const *V_host_ptrs, const double *coef_host, int k, int64_t n);
int vec_mod11(smf_ctx *ctx, double *x, int64_t first, double mean, int64_t n);
int cheb_first(smf_ctx *ctx, double *x, const double *dinv, const double *b, double inv_theta, int64_t n);
int cheb_update(smf_ctx *ctx, double *x, double *xo, const double *dinv, const double *b, double *t,
                double f1, double f2, bool use_xo, int64_t n);   // xo <- (1+f1) x - f1 xo + f2 dinv (b - t); t <- 0; caller swaps
int vec_residual(smf_ctx *ctx, double *t, const double *b, int64_t n);   // t = b - t
int vec_gather(smf_ctx *ctx, double *y, const double *x, const uint32_t *idx, int64_t n, bool scatter);
}  // namespace smf
