// Distributed vectors: one context per rank/GPU, one subdomain (a range of coarse-cell trees) per rank.
//
// Replaces what deal.II's LinearAlgebra::distributed::Vector / Utilities::MPI::Partitioner do around every cell loop
// of the reference (update_ghost_values before, compress(add) after: matrix_free_operators.cc:626, 724, 856, 1063;
// explicit calls stokes_matrix_free_local_smoothing.cc:645, 746, 1407) and the MPI_Allreduce inside every dot product.
//
// Layout: every rank stores ALL DoFs its cells touch; a DoF on a subdomain interface therefore exists on every rank
// that touches it, and all copies are kept equal ("consistent").  Point-wise vector updates preserve that by
// construction; an operator application produces per-rank partial sums on the interface, which one exchange turns
// into the full sum on every copy (compress(add) and update_ghost_values in a single step).  The sum is taken in
// ascending rank order on every rank, so the copies stay BITWISE equal.  Inner products weight every DoF with
// 1 / (number of ranks sharing it) and finish with an all-reduce.
//
// Transport of the interface values: ONE kernel per exchange over NVLink peer memory (CUDA IPC).  Every rank owns an
// "inbox arena"; the kernel packs the rank's partial sums straight into the neighbours' inboxes (remote stores),
// publishes an epoch flag per neighbour, waits for the neighbours' flags and accumulates in rank order -- pack,
// transfer, synchronisation and accumulation fused, ~one launch latency instead of pack + NCCL group + unpack.  The
// inboxes are double-buffered by epoch parity (a neighbour can be at most one exchange ahead), and the epoch lives
// in device memory, so the kernel can sit inside the CUDA graphs of the coarse V-cycle levels.  If the IPC mapping is
// not available the same exchange runs as pack -> NCCL send/recv group -> accumulate.  Inner products: ncclAllReduce.
// NCCL is bound at run time (dlopen) so that the single-GPU library has no NCCL dependency.
#include "smf_internal.h"
#include <algorithm>
#include <dlfcn.h>
#include <map>
#include <nccl.h>

namespace smf {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api(std::string *why)
{
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // the copy PyTorch already loaded wins (same SONAME); otherwise the system library
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
      api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
      api.Send = (decltype(api.Send))dlsym(h, "ncclSend");
      api.Recv = (decltype(api.Recv))dlsym(h, "ncclRecv");
      api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.Send && api.Recv && api.GroupStart &&
          api.GroupEnd && api.GetErrorString)
        api.handle = h;
    }
  }
  if (!api.handle) {
    if (why) *why = "libnccl.so.2 could not be loaded (needed for more than one rank)";
    return nullptr;
  }
  return &api;
}

#define SMF_NCCL(call)                                                                            \
  do {                                                                                            \
    ncclResult_t r__ = (call);                                                                    \
    if (r__ != ncclSuccess)                                                                       \
      return ctx->fail(SMF_ERR_CUDA, std::string(#call) + ": " + nccl_api(nullptr)->GetErrorString(r__)); \
  } while (0)

// ---- kernels ---------------------------------------------------------------------------------------------------
__global__ void k_pack(double *buf, const double *vec, const uint32_t *idx, int64_t n_nodes, int ncomp)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes * ncomp) return;
  const int64_t i = t / ncomp;
  const int c = (int)(t - i * ncomp);
  buf[t] = vec[idx[i] + c];
}
// every copy of an interface DoF := sum of all ranks' partial values, added in ascending rank order.
// if_src[e] >= 0: position (node) in the receive buffer; -1: this rank's own value.  Bit (29 + c) of if_idx marks a
// constrained component: its copies already agree (unit diagonal of MatrixFreeOperators::Base) and are left alone.
__global__ void k_accumulate(double *vec, const double *recv, const uint32_t *if_idx, const int *if_ptr, const int *if_src,
                             int64_t n_if, int ncomp)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_if * ncomp) return;
  const int64_t i = t / ncomp;
  const int c = (int)(t - i * ncomp);
  const uint32_t e = if_idx[i];
  if ((e >> (NODE_FLAG_SHIFT + c)) & 1u) return;
  double *p = vec + (e & NODE_IDX_MASK) + c;
  double sum = 0.0;
  for (int k = if_ptr[i]; k < if_ptr[i + 1]; ++k) {
    const int s = if_src[k];
    sum += (s < 0) ? *p : recv[(int64_t)s * ncomp + c];
  }
  *p = sum;
}
// ---- fused pack / push / wait / accumulate over peer memory ---------------------------------------------------
constexpr int P2P_MAX_NB = 26;
constexpr int P2P_THREADS = 1024;       // the loops are chains of dependent loads: as many threads as elements, up to one block per SM
constexpr size_t P2P_FLAG_BYTES = 1 << 16;          // flag slots at the start of every arena
struct P2PNb {
  double *remote;                    // the neighbour's inbox segment for this rank (parity 0), mapped through CUDA IPC
  int64_t remote_stride;             // elements between the two parity buffers of the neighbour's inbox
  unsigned long long *remote_flag;   // the neighbour's flag slot for this rank
  unsigned long long *local_flag;    // this rank's flag slot for the neighbour
  int64_t seg_end;                   // end of the neighbour's segment in the send list (elements)
  int64_t seg_start;
};
struct P2PTab { P2PNb nb[P2P_MAX_NB]; int n_nb; };

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{ unsigned long long v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v)
{ asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// state[0] = epoch (number of completed exchanges), [1] = blocks through the push, [2] = push-complete epoch,
// [3] = blocks through the accumulation.  The grid is one co-resident wave (blocks wait for each other).
__global__ void __launch_bounds__(P2P_THREADS) k_exchange_p2p(double *vec, const uint32_t *send_idx, int64_t n_send_el, int ncomp, const P2PTab *tab,
                                                       const double *inbox, int64_t inbox_stride, const uint32_t *if_idx, const int *if_ptr,
                                                       const int *if_src, int64_t n_if_el, unsigned long long *state)
{
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) s_epoch = ld_volatile_u64(state);
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  const int64_t par = (int64_t)(epoch & 1ull);
  const int n_nb = tab->n_nb;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gstride = (int64_t)gridDim.x * blockDim.x;
  // push: this rank's partial sums go straight into the neighbours' inboxes
  for (int64_t t = gtid; t < n_send_el; t += gstride) {
    const int64_t i = t / ncomp;
    const int c = (int)(t - i * ncomp);
    const double v = vec[send_idx[i] + c];
    int k = 0;
    while (k + 1 < n_nb && t >= tab->nb[k].seg_end) ++k;
    const P2PNb &nb = tab->nb[k];
    nb.remote[par * nb.remote_stride + (t - nb.seg_start)] = v;
  }
  // (one fence per block: the barrier orders the block's stores before thread 0's release at system scope)
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_acq_rel_sys();
    const unsigned long long prev = atomicAdd(&state[1], 1ull);
    if (prev == gridDim.x - 1) {          // last block: every push of this rank is visible system-wide
      fence_acq_rel_sys();
      for (int k = 0; k < n_nb; ++k) st_volatile_u64(tab->nb[k].remote_flag, epoch + 1);
      state[1] = 0;
      if (gridDim.x > 1) { __threadfence(); st_volatile_u64(&state[2], epoch + 1); }
    } else {
      // wait until all pushes of this rank are done (the accumulation overwrites vec)
      while (ld_volatile_u64(&state[2]) < epoch + 1) {}
    }
  }
  // wait for every neighbour's data
  if (threadIdx.x >= 32 && threadIdx.x < 32 + n_nb) {
    while (ld_volatile_u64(tab->nb[threadIdx.x - 32].local_flag) < epoch + 1) {}
    fence_acq_rel_sys();
  }
  __syncthreads();
  // accumulate in ascending rank order (bitwise equal copies on all ranks)
  const double *in = inbox + par * inbox_stride;
  for (int64_t t = gtid; t < n_if_el; t += gstride) {
    const int64_t i = t / ncomp;
    const int c = (int)(t - i * ncomp);
    const uint32_t e = if_idx[i];
    if ((e >> (NODE_FLAG_SHIFT + c)) & 1u) continue;
    double *p = vec + (e & NODE_IDX_MASK) + c;
    double sum = 0.0;
    for (int k = if_ptr[i]; k < if_ptr[i + 1]; ++k) {
      const int s = if_src[k];
      sum += (s < 0) ? *p : __ldcg(in + (int64_t)s * ncomp + c);
    }
    *p = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(&state[3], 1ull);
    if (prev == gridDim.x - 1) { state[3] = 0; __threadfence(); st_volatile_u64(state, epoch + 1); }
  }
}

struct P2PState {
  char *arena = nullptr;
  size_t bytes = 0, used = P2P_FLAG_BYTES;
  int n_flags = 0;
  std::vector<char *> peer;          // [world] base of every rank's arena in this process (own rank: arena)
  int max_blocks = 148;
};

__global__ void k_mod11_gid(double *x, const int64_t *gid, double mean, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = (double)(gid[i] % 11) - mean;
}

void free_exchange(ExchangeDev &E)
{
  cudaFree(E.send_idx); cudaFree(E.send_buf); cudaFree(E.recv_buf); cudaFree(E.if_idx); cudaFree(E.if_ptr); cudaFree(E.if_src);
  cudaFree(E.weights); cudaFree(E.gid); cudaFree(E.p2p_tab); cudaFree(E.p2p_state);
  E = ExchangeDev();
}

// one exchange: pack -> grouped send/recv with every neighbour -> rank-ordered accumulation
int exchange_add(smf_ctx *ctx, ExchangeDev &E, double *vec)
{
  if (ctx->world == 1 || E.n_send == 0) return SMF_OK;
  NcclApi *api = nccl_api(&ctx->err);
  if (!api || !ctx->comm) return ctx->fail(SMF_ERR_INVALID, "exchange without a communicator (smf_comm_init)");
  ProfScope prof(ctx, B_EXCHANGE);
  const int64_t n = E.n_send * E.ncomp;
  if (E.p2p_tab) {
    const P2PState *P = (const P2PState *)ctx->p2p;
    const int64_t work = std::max(n, E.n_if * E.ncomp);
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((work + P2P_THREADS - 1) / P2P_THREADS, P->max_blocks));
    k_exchange_p2p<<<blocks, P2P_THREADS, 0, ctx->stream>>>(vec, E.send_idx, n, E.ncomp, (const P2PTab *)E.p2p_tab, E.p2p_inbox, n, E.if_idx,
                                                    E.if_ptr, E.if_src, E.n_if * E.ncomp, (unsigned long long *)E.p2p_state);
    ctx->launches++;
    ctx->exchanges++;
    SMF_CUDA(cudaGetLastError());
    return SMF_OK;
  }
  k_pack<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(E.send_buf, vec, E.send_idx, E.n_send, E.ncomp);
  ctx->launches++;
  SMF_NCCL(api->GroupStart());
  for (size_t k = 0; k < E.ranks.size(); ++k) {
    const size_t cnt = (size_t)E.counts[k] * E.ncomp, off = (size_t)E.offsets[k] * E.ncomp;
    SMF_NCCL(api->Send(E.send_buf + off, cnt, ncclDouble, E.ranks[k], (ncclComm_t)ctx->comm, ctx->stream));
    SMF_NCCL(api->Recv(E.recv_buf + off, cnt, ncclDouble, E.ranks[k], (ncclComm_t)ctx->comm, ctx->stream));
  }
  SMF_NCCL(api->GroupEnd());
  const int64_t m = E.n_if * E.ncomp;
  k_accumulate<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(vec, E.recv_buf, E.if_idx, E.if_ptr, E.if_src, E.n_if, E.ncomp);
  ctx->launches++;
  ctx->exchanges++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}

// sum of `count` doubles at `dev` over all ranks (in place, stream ordered)
int allreduce_sum(smf_ctx *ctx, double *dev, int count)
{
  if (ctx->world == 1) return SMF_OK;
  NcclApi *api = nccl_api(&ctx->err);
  if (!api || !ctx->comm) return ctx->fail(SMF_ERR_INVALID, "all-reduce without a communicator (smf_comm_init)");
  SMF_NCCL(api->AllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  return SMF_OK;
}

// global minimum / maximum of a pair of host values (Utilities::MPI::min / max, matrix_free_operators.cc:258-259): every
// rank puts its pair into its own slot of a zeroed table, the table is summed over the ranks, then reduced on the host
int allreduce_minmax(smf_ctx *ctx, double *mn, double *mx)
{
  if (ctx->world == 1) return SMF_OK;
  std::vector<double> tab((size_t)2 * ctx->world, 0.0);
  tab[2 * ctx->rank] = *mn; tab[2 * ctx->rank + 1] = *mx;
  double *d = nullptr;
  SMF_CUDA(cudaMalloc(&d, sizeof(double) * tab.size()));
  SMF_CUDA(cudaMemcpyAsync(d, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, ctx->stream));
  const int e = allreduce_sum(ctx, d, (int)tab.size());
  if (e) { cudaFree(d); return e; }
  SMF_CUDA(cudaMemcpyAsync(tab.data(), d, sizeof(double) * tab.size(), cudaMemcpyDeviceToHost, ctx->stream));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d);
  for (int r = 0; r < ctx->world; ++r) { *mn = std::min(*mn, tab[2 * r]); *mx = std::max(*mx, tab[2 * r + 1]); }
  return SMF_OK;
}

int start_vector_mod11(smf_ctx *ctx, const ExchangeDev &E, double *x, int64_t n)
{
  k_mod11_gid<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(x, E.gid, E.gid_mean, n);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}

}  // namespace smf

using namespace smf;

extern "C" {

int smf_comm_get_unique_id(char *id)
{
  if (!id) return SMF_ERR_INVALID;
  NcclApi *api = nccl_api(nullptr);
  if (!api) return SMF_ERR_UNSUPPORTED;
  ncclUniqueId u;
  if (api->GetUniqueId(&u) != ncclSuccess) return SMF_ERR_CUDA;
  memcpy(id, u.internal, SMF_UNIQUE_ID_BYTES);
  return SMF_OK;
}

int smf_comm_init(smf_ctx *ctx, int rank, int world, const char *id)
{
  if (!ctx || world < 1 || rank < 0 || rank >= world) return SMF_ERR_INVALID;
  ctx->rank = rank; ctx->world = world;
  if (world == 1) return SMF_OK;
  if (!id) return ctx->fail(SMF_ERR_INVALID, "unique id missing");
  NcclApi *api = nccl_api(&ctx->err);
  if (!api) return SMF_ERR_UNSUPPORTED;
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  ncclUniqueId u;
  memcpy(u.internal, id, SMF_UNIQUE_ID_BYTES);
  ncclComm_t comm = nullptr;
  SMF_NCCL(api->CommInitRank(&comm, world, u, rank));
  ctx->comm = comm;
  // inbox arena of the fused peer-memory exchange: allocate, publish the IPC handle, map every peer's arena.
  // The decision to use it is collective (all ranks or none).
  double ok = 1.0;
  P2PState *P = nullptr;
  const char *env = getenv("SMF_P2P");
  if (ctx->use_p2p && !(env && env[0] == '0') && world <= P2P_MAX_NB + 1) {
    P = new P2PState();
    const char *mb = getenv("SMF_P2P_ARENA_MB");
    P->bytes = (size_t)(mb ? std::max(8, atoi(mb)) : 256) << 20;
    cudaIpcMemHandle_t mine;
    std::vector<cudaIpcMemHandle_t> all(world);
    char *d_all = nullptr;
    if (cudaMalloc(&P->arena, P->bytes) != cudaSuccess || cudaMemsetAsync(P->arena, 0, P->bytes, ctx->stream) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine, P->arena) != cudaSuccess || cudaMalloc(&d_all, sizeof(mine) * world) != cudaSuccess)
      ok = 0.0;
    cudaGetLastError();
    if (d_all) {
      cudaMemcpyAsync(d_all + sizeof(mine) * rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream);
      SMF_NCCL(api->AllGather(d_all + sizeof(mine) * rank, d_all, sizeof(mine), ncclChar, comm, ctx->stream));
      SMF_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(mine) * world, cudaMemcpyDeviceToHost, ctx->stream));
      SMF_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(d_all);
    } else {      // keep the collective sequence identical on every rank
      return ctx->fail(SMF_ERR_CUDA, "cudaMalloc failed while setting up the peer-memory exchange");
    }
    P->peer.assign(world, nullptr);
    P->peer[rank] = P->arena;
    for (int q = 0; q < world && ok == 1.0; ++q) {
      if (q == rank) continue;
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0.0; cudaGetLastError(); }
      P->peer[q] = (char *)ptr;
    }
    int per_sm = 0;
    cudaDeviceProp prop;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_exchange_p2p, P2P_THREADS, 0) == cudaSuccess &&
        cudaGetDeviceProperties(&prop, ctx->cfg.device) == cudaSuccess && per_sm > 0)
      P->max_blocks = prop.multiProcessorCount;      // one co-resident wave, one block per SM
  } else ok = 0.0;
  {      // all ranks or none
    double *d = nullptr;
    SMF_CUDA(cudaMalloc(&d, sizeof(double)));
    SMF_CUDA(cudaMemcpyAsync(d, &ok, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SMF_NCCL(api->AllReduce(d, d, 1, ncclDouble, ncclMin, comm, ctx->stream));
    SMF_CUDA(cudaMemcpyAsync(&ok, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SMF_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
  }
  if (ok == 1.0) ctx->p2p = P;
  else if (P) {
    for (int q = 0; q < (int)P->peer.size(); ++q) if (q != rank && P->peer[q]) cudaIpcCloseMemHandle(P->peer[q]);
    cudaFree(P->arena);
    delete P;
  }
  if (!ctx->p2p) ctx->use_graphs = 0;      // (the NCCL transport is not captured into the V-cycle graphs)
  return SMF_OK;
}

int smf_detect_box_lattice(const uint32_t *cell_dofs_p, int64_t n_cells, int64_t n_dofs_p, int32_t *nodes_per_direction, int32_t *cell_xyz)
{
  if (!cell_dofs_p || n_cells < 1 || n_dofs_p < 1 || n_dofs_p >= (int64_t)NODE_IDX_MASK) return 0;
  for (int64_t i = 0; i < n_cells * 8; ++i) if (cell_dofs_p[i] >= (uint64_t)n_dofs_p) return 0;
  std::vector<uint32_t> nodes(cell_dofs_p, cell_dofs_p + n_cells * 8);
  std::vector<int> xyz;
  int dims[3];
  if (!lattice_cell_coords(nodes, n_cells, n_dofs_p, dims, xyz)) return 0;
  if (nodes_per_direction) for (int d = 0; d < 3; ++d) nodes_per_direction[d] = dims[d];
  if (cell_xyz) for (size_t i = 0; i < xyz.size(); ++i) cell_xyz[i] = xyz[i];
  return 1;
}
int smf_lattice_schur_active(const smf_ctx *ctx) { return ctx && ctx->use_lattice && ctx->lat_state == 1 ? 1 : 0; }
int smf_lattice_replicated_levels(const smf_ctx *ctx) { return ctx && ctx->use_lattice && ctx->lat_state == 1 ? ctx->lat_repl + 1 : 0; }
int smf_replicated_A_levels(const smf_ctx *ctx) { return ctx ? ctx->replica_top + 1 : 0; }
int smf_exchange_transport(const smf_ctx *ctx) { return !ctx || ctx->world == 1 ? 0 : (ctx->p2p ? 2 : 1); }

void smf_comm_destroy_internal(smf_ctx *ctx)
{
  if (ctx->p2p) {
    P2PState *P = (P2PState *)ctx->p2p;
    for (int q = 0; q < (int)P->peer.size(); ++q) if (q != ctx->rank && P->peer[q]) cudaIpcCloseMemHandle(P->peer[q]);
    cudaFree(P->arena);
    delete P;
    ctx->p2p = nullptr;
  }
  if (ctx->comm) {
    NcclApi *api = nccl_api(nullptr);
    if (api) api->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
}

int smf_set_exchange(smf_ctx *ctx, int level, int family, const smf_exchange *x)
{
  if (!ctx || !x || (family != 0 && family != 1)) return SMF_ERR_INVALID;
  if (level < SMF_ACTIVE || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  LevelDev &L = ctx->lev(level);
  if (L.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "set the level mesh first");
  SMF_CUDA(cudaSetDevice(ctx->cfg.device));
  ExchangeDev &E = family == 0 ? L.ex_u : L.ex_p;
  lattice_free(ctx);
  free_exchange(E);
  const int ncomp = family == 0 ? L.dim : 1;
  const int64_t n_dofs = family == 0 ? L.n_u : L.n_p;
  const std::vector<uint8_t> &cons = family == 0 ? L.h_cons_u : L.h_cons_p;
  E.ncomp = ncomp;
  // ---- neighbour lists ----
  std::vector<uint32_t> send_idx;
  std::map<uint32_t, std::vector<std::pair<int, int>>> sharers;     // first DoF of the node -> (rank, position in recv buffer)
  int64_t off = 0;
  for (int k = 0; k < x->n_neighbors; ++k) {
    const int q = x->neighbor_ranks[k];
    if (q < 0 || q >= ctx->world || q == ctx->rank) return ctx->fail(SMF_ERR_INVALID, "bad neighbour rank");
    if (k > 0 && q <= x->neighbor_ranks[k - 1]) return ctx->fail(SMF_ERR_INVALID, "neighbour ranks must be ascending");
    E.ranks.push_back(q); E.counts.push_back(x->counts[k]); E.offsets.push_back(off);
    for (int64_t i = 0; i < x->counts[k]; ++i) {
      const uint32_t d0 = x->nodes[k][i];
      if ((int64_t)d0 + ncomp > n_dofs) return ctx->fail(SMF_ERR_INVALID, "exchange index out of range");
      send_idx.push_back(d0);
      sharers[d0].push_back({q, (int)(off + i)});
    }
    off += x->counts[k];
  }
  E.n_send = off;
  // ---- rank-ordered accumulation tables, weights, ownership ----
  std::vector<uint32_t> if_idx;
  std::vector<int> if_ptr(1, 0), if_src;
  std::vector<double> weights((size_t)n_dofs, 1.0);
  std::vector<uint8_t> &not_owned = family == 0 ? L.h_not_owned_u : L.h_not_owned_p;
  not_owned.assign((size_t)n_dofs, 0);
  for (auto &kv : sharers) {
    const uint32_t d0 = kv.first;
    auto &lst = kv.second;       // ascending rank (neighbours were ascending)
    uint32_t flags = 0;
    for (int c = 0; c < ncomp; ++c)
      if (!cons.empty() && cons[d0 + c]) flags |= 1u << (NODE_FLAG_SHIFT + c);
    if_idx.push_back(d0 | flags);
    bool own_done = false;
    for (auto &pr : lst) {
      if (!own_done && pr.first > ctx->rank) { if_src.push_back(-1); own_done = true; }
      if_src.push_back(pr.second);
    }
    if (!own_done) if_src.push_back(-1);
    if_ptr.push_back((int)if_src.size());
    const double w = 1.0 / (double)(lst.size() + 1);
    for (int c = 0; c < ncomp; ++c) { weights[d0 + c] = w; if (lst.front().first < ctx->rank) not_owned[d0 + c] = 1; }
  }
  E.n_if = (int64_t)if_idx.size();
  auto up = [&](auto **dev, const auto &host) -> int {
    using T = typename std::remove_reference<decltype(host[0])>::type;
    const size_t nb = std::max<size_t>(host.size(), 1) * sizeof(T);
    SMF_CUDA(cudaMalloc((void **)dev, nb));
    if (!host.empty()) SMF_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    ctx->bytes += (int64_t)nb;
    return SMF_OK;
  };
  int e;
  if (E.n_send > 0) {
    if ((e = up(&E.send_idx, send_idx))) return e;
    if ((e = up(&E.if_idx, if_idx))) return e;
    if ((e = up(&E.if_ptr, if_ptr))) return e;
    if ((e = up(&E.if_src, if_src))) return e;
    SMF_CUDA(cudaMalloc(&E.send_buf, sizeof(double) * E.n_send * ncomp));
    SMF_CUDA(cudaMalloc(&E.recv_buf, sizeof(double) * E.n_send * ncomp));
    ctx->bytes += 2 * (int64_t)sizeof(double) * E.n_send * ncomp;
  }
  if (ctx->world > 1) {
    if ((e = up(&E.weights, weights))) return e;
  }
  // ---- peer-memory transport: inbox in this rank's arena, addresses traded with the neighbours (collective) ----
  if (ctx->p2p && E.n_send > 0) {
    P2PState *P = (P2PState *)ctx->p2p;
    NcclApi *api = nccl_api(&ctx->err);
    const int nn = (int)E.ranks.size();
    const size_t par_bytes = sizeof(double) * (size_t)E.n_send * ncomp;
    const size_t inbox_off = (P->used + 255) & ~(size_t)255;
    if (nn > P2P_MAX_NB) return ctx->fail(SMF_ERR_UNSUPPORTED, "too many neighbour ranks for the peer-memory exchange");
    if (inbox_off + 2 * par_bytes > P->bytes || (size_t)(P->n_flags + nn) * 8 > P2P_FLAG_BYTES)
      return ctx->fail(SMF_ERR_INVALID, "peer-memory inbox arena too small: set SMF_P2P_ARENA_MB");
    P->used = inbox_off + 2 * par_bytes;
    const int flag0 = P->n_flags;
    P->n_flags += nn;
    E.p2p_inbox = (double *)(P->arena + inbox_off);
    std::vector<int64_t> meta_send(3 * nn), meta_recv(3 * nn);
    for (int k = 0; k < nn; ++k) {
      meta_send[3 * k + 0] = (int64_t)(inbox_off + sizeof(double) * (size_t)E.offsets[k] * ncomp);   // where neighbour k writes (bytes)
      meta_send[3 * k + 1] = (int64_t)E.n_send * ncomp;                                              // parity stride (elements)
      meta_send[3 * k + 2] = (int64_t)(flag0 + k) * 8;                                               // flag slot neighbour k raises (bytes)
    }
    int64_t *d_meta = nullptr;
    SMF_CUDA(cudaMalloc(&d_meta, sizeof(int64_t) * 6 * nn));
    SMF_CUDA(cudaMemcpyAsync(d_meta, meta_send.data(), sizeof(int64_t) * 3 * nn, cudaMemcpyHostToDevice, ctx->stream));
    SMF_NCCL(api->GroupStart());
    for (int k = 0; k < nn; ++k) {
      SMF_NCCL(api->Send(d_meta + 3 * k, 3, ncclInt64, E.ranks[k], (ncclComm_t)ctx->comm, ctx->stream));
      SMF_NCCL(api->Recv(d_meta + 3 * nn + 3 * k, 3, ncclInt64, E.ranks[k], (ncclComm_t)ctx->comm, ctx->stream));
    }
    SMF_NCCL(api->GroupEnd());
    SMF_CUDA(cudaMemcpyAsync(meta_recv.data(), d_meta + 3 * nn, sizeof(int64_t) * 3 * nn, cudaMemcpyDeviceToHost, ctx->stream));
    SMF_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_meta);
    P2PTab tab;
    memset(&tab, 0, sizeof(tab));
    tab.n_nb = nn;
    for (int k = 0; k < nn; ++k) {
      char *base = P->peer[E.ranks[k]];
      tab.nb[k].remote = (double *)(base + meta_recv[3 * k + 0]);
      tab.nb[k].remote_stride = meta_recv[3 * k + 1];
      tab.nb[k].remote_flag = (unsigned long long *)(base + meta_recv[3 * k + 2]);
      tab.nb[k].local_flag = (unsigned long long *)(P->arena + (size_t)(flag0 + k) * 8);
      tab.nb[k].seg_start = E.offsets[k] * ncomp;
      tab.nb[k].seg_end = (E.offsets[k] + E.counts[k]) * ncomp;
    }
    SMF_CUDA(cudaMalloc(&E.p2p_tab, sizeof(P2PTab)));
    SMF_CUDA(cudaMemcpy(E.p2p_tab, &tab, sizeof(P2PTab), cudaMemcpyHostToDevice));
    SMF_CUDA(cudaMalloc(&E.p2p_state, 4 * sizeof(unsigned long long)));
    SMF_CUDA(cudaMemset(E.p2p_state, 0, 4 * sizeof(unsigned long long)));
  }
  // ---- global DoF numbers for the Chebyshev start vector ((global index mod 11) - mean, App. A.2) ----
  if (x->global_dof_ids) {
    std::vector<int64_t> gid(x->global_dof_ids, x->global_dof_ids + n_dofs);
    if ((e = up(&E.gid, gid))) return e;
    if (family == 1 || n_dofs <= 8000000) E.h_gid = gid;      // (velocity: coarse levels only, for velocity_replica.cu)
    double acc[2] = {0.0, 0.0};
    for (int64_t i = 0; i < n_dofs; ++i) { acc[0] += weights[i] * (double)(gid[i] % 11); acc[1] += weights[i]; }
    if (ctx->world > 1) {          // collective: every rank calls smf_set_exchange for the same (level, family)
      double *d = nullptr;
      SMF_CUDA(cudaMalloc(&d, 2 * sizeof(double)));
      SMF_CUDA(cudaMemcpyAsync(d, acc, 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      if ((e = allreduce_sum(ctx, d, 2))) { cudaFree(d); return e; }
      SMF_CUDA(cudaMemcpyAsync(acc, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      SMF_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(d);
    }
    E.gid_mean = acc[0] / acc[1];
  }
  E.set = true;
  ctx->smoothers_ready = false;
  return SMF_OK;
}

int smf_exchange_add(smf_ctx *ctx, int level, int family, double *vec)
{
  if (!ctx || !vec || (family != 0 && family != 1)) return SMF_ERR_INVALID;
  if (level < SMF_ACTIVE || level >= ctx->cfg.n_levels) return ctx->fail(SMF_ERR_INVALID, "level out of range");
  LevelDev &L = ctx->lev(level);
  return exchange_add(ctx, family == 0 ? L.ex_u : L.ex_p, vec);
}

}  // extern "C"
