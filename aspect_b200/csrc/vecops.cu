// Vector kernels for the Chebyshev smoothers and the Krylov driver (HBM-streaming, FP64).
// They replace the LinearAlgebra::distributed::Vector operations deal.II's PreconditionChebyshev,
// SolverCG and SolverGMRES perform between operator applications.
#include "smf_internal.h"

namespace smf {

constexpr int VT = 256;                 // threads per block
constexpr int RED_BLOCKS = 148 * 4;     // persistent reduction grid: 4 blocks per SM
constexpr int MD_CHUNK = SMF_MAX_RESTART;   // max basis vectors per multi-dot pass (= longest GMRES restart length)

static inline unsigned ew_blocks(int64_t n)
{
  int64_t b = (n + VT * 2 - 1) / (VT * 2);
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (unsigned)b;
}

#define GRID_STRIDE(i, n) for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

__global__ void k_set(double *x, double v, int64_t n) { GRID_STRIDE(i, n) x[i] = v; }
__global__ void k_copy(double *y, const double *x, int64_t n) { GRID_STRIDE(i, n) y[i] = x[i]; }
__global__ void k_scale(double *x, double a, int64_t n) { GRID_STRIDE(i, n) x[i] *= a; }
__global__ void k_axpy(double *y, double a, const double *x, int64_t n) { GRID_STRIDE(i, n) y[i] += a * x[i]; }
__global__ void k_sadd(double *y, double s, double a, const double *x, int64_t n) { GRID_STRIDE(i, n) y[i] = s * y[i] + a * x[i]; }
__global__ void k_pointwise(double *z, double a, const double *d, const double *x, int64_t n) { GRID_STRIDE(i, n) z[i] = a * d[i] * x[i]; }
__global__ void k_residual(double *t, const double *b, int64_t n) { GRID_STRIDE(i, n) t[i] = b[i] - t[i]; }
__global__ void k_cheb_first(double *x, const double *dinv, const double *b, double it, int64_t n)
{ GRID_STRIDE(i, n) x[i] = it * dinv[i] * b[i]; }
// xo <- (1+f1) x - f1 xo + f2 dinv (b - t)   (PreconditionChebyshev vector_updates, iteration_index >= 1)
// t is left zeroed: the next operator application into t then needs no separate memset pass
__global__ void k_cheb_update(double *x, double *xo, const double *dinv, const double *b, double *t,
                              double f1, double f2, int use_xo, int64_t n)
{
  GRID_STRIDE(i, n) {
    const double xi = x[i];
    double v = xi + f2 * dinv[i] * (b[i] - t[i]);
    if (use_xo) v += f1 * (xi - xo[i]);
    else v += f1 * xi;
    xo[i] = v;
    t[i] = 0.0;
  }
}
__global__ void k_shift(double *x, double v, int64_t n) { GRID_STRIDE(i, n) x[i] += v; }
__global__ void k_gather(double *y, const double *x, const uint32_t *idx, int64_t n) { GRID_STRIDE(i, n) y[i] = x[idx[i]]; }
__global__ void k_scatter(double *y, const double *x, const uint32_t *idx, int64_t n) { GRID_STRIDE(i, n) y[idx[i]] = x[i]; }
__global__ void k_mod11(double *x, int64_t first, double mean, int64_t n) { GRID_STRIDE(i, n) x[i] = (double)((i + first) % 11) - mean; }

#define LAUNCH_EW(kernel, n, ...)                                             \
  do {                                                                        \
    if ((n) > 0) {                                                            \
      ProfScope prof__(ctx, ctx->vec_bucket);                                 \
      kernel<<<ew_blocks(n), VT, 0, ctx->stream>>>(__VA_ARGS__);              \
      ctx->launches++;                                                        \
      SMF_CUDA(cudaGetLastError());                                           \
    }                                                                         \
  } while (0)

int vec_set(smf_ctx *ctx, double *x, double v, int64_t n) { LAUNCH_EW(k_set, n, x, v, n); return SMF_OK; }
int vec_shift(smf_ctx *ctx, double *x, double v, int64_t n) { LAUNCH_EW(k_shift, n, x, v, n); return SMF_OK; }
int vec_copy(smf_ctx *ctx, double *y, const double *x, int64_t n) { LAUNCH_EW(k_copy, n, y, x, n); return SMF_OK; }
int vec_scale(smf_ctx *ctx, double *x, double a, int64_t n) { LAUNCH_EW(k_scale, n, x, a, n); return SMF_OK; }
int vec_axpy(smf_ctx *ctx, double *y, double a, const double *x, int64_t n) { LAUNCH_EW(k_axpy, n, y, a, x, n); return SMF_OK; }
int vec_sadd(smf_ctx *ctx, double *y, double s, double a, const double *x, int64_t n) { LAUNCH_EW(k_sadd, n, y, s, a, x, n); return SMF_OK; }
int vec_pointwise(smf_ctx *ctx, double *z, double a, const double *d, const double *x, int64_t n) { LAUNCH_EW(k_pointwise, n, z, a, d, x, n); return SMF_OK; }
int vec_residual(smf_ctx *ctx, double *t, const double *b, int64_t n) { LAUNCH_EW(k_residual, n, t, b, n); return SMF_OK; }
int cheb_first(smf_ctx *ctx, double *x, const double *dinv, const double *b, double inv_theta, int64_t n)
{ LAUNCH_EW(k_cheb_first, n, x, dinv, b, inv_theta, n); return SMF_OK; }
int cheb_update(smf_ctx *ctx, double *x, double *xo, const double *dinv, const double *b, double *t,
                double f1, double f2, bool use_xo, int64_t n)
{ LAUNCH_EW(k_cheb_update, n, x, xo, dinv, b, t, f1, f2, use_xo ? 1 : 0, n); return SMF_OK; }
int vec_gather(smf_ctx *ctx, double *y, const double *x, const uint32_t *idx, int64_t n, bool scatter)
{
  if (scatter) LAUNCH_EW(k_scatter, n, y, x, idx, n);
  else LAUNCH_EW(k_gather, n, y, x, idx, n);
  return SMF_OK;
}
int vec_mod11(smf_ctx *ctx, double *x, int64_t first, double mean, int64_t n) { LAUNCH_EW(k_mod11, n, x, first, mean, n); return SMF_OK; }

// ---- reductions: warp shuffles -> shared -> one partial per block -> last block sums in fixed order ----------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[j] = sum_i V[j][i] * w[i], j < k; out[k] = w.w.  A block keeps a chunk of w in registers (MD_W doubles per
// thread) and walks the whole basis over it in groups of MD_GROUP vectors, all loads of a group in flight at once;
// each group ends in a warp reduction into the warp's own shared accumulators, so w and V are read exactly once.
// partial[block][j], then the last block sums the partials in a fixed order (deterministic).
constexpr int MD_W = 8;        // doubles of w per thread and chunk (as MD_W / 2 16-byte loads)
constexpr int MD_GROUP = 4;
// two consecutive doubles at p[i], p[i+1] (zero past the end); one 16-byte load when the vector allows it
__device__ __forceinline__ double2 load2(const double *p, int64_t i, int64_t n, bool vec2)
{
  if (vec2 && i + 1 < n) return *reinterpret_cast<const double2 *>(p + i);
  double2 r;
  r.x = (i < n) ? p[i] : 0.0;
  r.y = (i + 1 < n) ? p[i + 1] : 0.0;
  return r;
}
// TWO: a second right-hand side w2 is dotted with the same basis in the same pass (out[k + 1 + j] = V[j] . w2): the
// Gram-matrix column the single-pass re-orthogonalisation of the GMRES driver needs (vec_multi_dot_gram below).
template <bool TWO>
__global__ void __launch_bounds__(VT, 2) k_multi_dot(const double *const *V, int k, const double *w, const double *w2, int64_t n,
                                                      double *partial, unsigned int *counter, double *out, int vec2,
                                                      const double *wgt)
{
  constexpr int NR = TWO ? 2 : 1;
  __shared__ double s_acc[VT / 32][NR * (MD_CHUNK + 4)];
  __shared__ const double *s_v[MD_CHUNK];
  __shared__ bool s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = VT / 32;
  for (int j = threadIdx.x; j < nw * NR * (MD_CHUNK + 4); j += VT) (&s_acc[0][0])[j] = 0.0;
  for (int j = threadIdx.x; j < k; j += VT) s_v[j] = V[j];
  __syncthreads();
  constexpr int NL = MD_W / 2;
  const int64_t chunk = (int64_t)VT * MD_W;
  const int nout = NR * k + 1;
  for (int64_t base = (int64_t)blockIdx.x * chunk; base < n; base += (int64_t)gridDim.x * chunk) {
    double2 ww[NL], ww2[TWO ? NL : 1];
    double wn = 0.0;
#pragma unroll
    for (int r = 0; r < NL; ++r) {
      const int64_t i = base + ((int64_t)r * VT + threadIdx.x) * 2;
      ww[r] = load2(w, i, n, vec2);
      if (TWO) ww2[r] = load2(w2, i, n, vec2);
      const double x0 = ww[r].x, y0 = ww[r].y;
      if (wgt) {       // distributed run: every shared DoF counts 1 / (number of ranks that hold it)
        const double2 g = load2(wgt, i, n, vec2);
        ww[r].x *= g.x; ww[r].y *= g.y;
        if (TWO) { ww2[r].x *= g.x; ww2[r].y *= g.y; }
      }
      wn = fma(ww[r].x, x0, fma(ww[r].y, y0, wn));
    }
    for (int g = 0; g < k; g += MD_GROUP) {
      double2 vv[MD_GROUP][NL];
#pragma unroll
      for (int jj = 0; jj < MD_GROUP; ++jj) {
        const double *v = s_v[min(g + jj, k - 1)];
#pragma unroll
        for (int r = 0; r < NL; ++r) vv[jj][r] = load2(v, base + ((int64_t)r * VT + threadIdx.x) * 2, n, vec2);
      }
#pragma unroll
      for (int jj = 0; jj < MD_GROUP; ++jj) {
        double a = 0.0;
#pragma unroll
        for (int r = 0; r < NL; ++r) a = fma(vv[jj][r].x, ww[r].x, fma(vv[jj][r].y, ww[r].y, a));
        a = warp_sum(a);
        if (lane == 0 && g + jj < k) s_acc[warp][g + jj] += a;
        if (TWO) {
          double a2 = 0.0;
#pragma unroll
          for (int r = 0; r < NL; ++r) a2 = fma(vv[jj][r].x, ww2[r].x, fma(vv[jj][r].y, ww2[r].y, a2));
          a2 = warp_sum(a2);
          if (lane == 0 && g + jj < k) s_acc[warp][k + 1 + g + jj] += a2;
        }
      }
    }
    wn = warp_sum(wn);
    if (lane == 0) s_acc[warp][k] += wn;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nout; j += VT) {
    double a = 0.0;
    for (int q = 0; q < nw; ++q) a += s_acc[q][j];
    partial[(size_t)blockIdx.x * nout + j] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    // one warp per result; lanes stride over the blocks' partials, then a shuffle tree: the order is fixed
    for (int j = warp; j < nout; j += nw) {
      double a = 0.0;
      for (unsigned bI = lane; bI < gridDim.x; bI += 32) a += partial[(size_t)bI * nout + j];
      a = warp_sum(a);
      if (lane == 0) out[j] = a;
    }
    if (threadIdx.x == 0) *counter = 0;
  }
}

// w -= sum_j coef[j] V[j]; out[0] = |w_new|^2
constexpr int MA_UNROLL = 8;
__global__ void __launch_bounds__(VT) k_multi_axpy(double *w, const double *const *V, const double *coef, int k, int64_t n,
                                                    double *partial, unsigned int *counter, double *out, int vec2,
                                                    const double *wgt)
{
  __shared__ double s_coef[MD_CHUNK];
  __shared__ const double *s_v[MD_CHUNK];
  __shared__ double s_acc[VT / 32];
  __shared__ bool s_last;
  for (int j = threadIdx.x; j < k; j += VT) { s_coef[j] = coef[j]; s_v[j] = V[j]; }
  __syncthreads();
  double nrm = 0.0;
  const int64_t stride = (int64_t)gridDim.x * VT * 2;
  for (int64_t i = ((int64_t)blockIdx.x * VT + threadIdx.x) * 2; i < n; i += stride) {
    double2 ww = load2(w, i, n, vec2);
    int j = 0;
    for (; j + MA_UNROLL <= k; j += MA_UNROLL) {
      double2 vv[MA_UNROLL];
#pragma unroll
      for (int jj = 0; jj < MA_UNROLL; ++jj) vv[jj] = load2(s_v[j + jj], i, n, vec2);
#pragma unroll
      for (int jj = 0; jj < MA_UNROLL; ++jj) { const double cj = s_coef[j + jj]; ww.x = fma(-cj, vv[jj].x, ww.x); ww.y = fma(-cj, vv[jj].y, ww.y); }
    }
    for (; j < k; ++j) {
      const double2 vv = load2(s_v[j], i, n, vec2);
      const double cj = s_coef[j];
      ww.x = fma(-cj, vv.x, ww.x); ww.y = fma(-cj, vv.y, ww.y);
    }
    if (vec2 && i + 1 < n) *reinterpret_cast<double2 *>(w + i) = ww;
    else { w[i] = ww.x; if (i + 1 < n) w[i + 1] = ww.y; }
    if (wgt) {
      const double2 g = load2(wgt, i, n, vec2);
      nrm = fma(ww.x * g.x, ww.x, fma(ww.y * g.y, ww.y, nrm));
    } else {
      nrm = fma(ww.x, ww.x, fma(ww.y, ww.y, nrm));
    }
  }
  nrm = warp_sum(nrm);
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = nrm;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int q = 0; q < VT / 32; ++q) a += s_acc[q];
    partial[blockIdx.x] = a;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double a = 0.0;
    for (unsigned bI = 0; bI < gridDim.x; ++bI) a += partial[bI];
    out[0] = a;
    *counter = 0;
  }
}

// scratch layout (doubles): [0, SCR_OUT) results | [SCR_OUT, SCR_OUT + RED_BLOCKS*SCR_PART) partials | then 2 counters |
// device pointer table | Gram-matrix defect E = V^T V - I of the current Krylov basis (GRAM_LD x GRAM_LD)
constexpr int SCR_OUT = 2 * MD_CHUNK + 16;                   // results: [0, 2k] dots, [OUT_NRM] the norm multi_axpy returns
constexpr int OUT_NRM = 2 * MD_CHUNK + 4;
constexpr int SCR_PART = 2 * MD_CHUNK + 1;      // partials per block
constexpr int GRAM_LD = MD_CHUNK + 4;
static int ensure_scratch(smf_ctx *ctx)
{
  if (ctx->scratch) return SMF_OK;
  const int64_t n = SCR_OUT + (int64_t)RED_BLOCKS * SCR_PART + 16 + 256 + (int64_t)GRAM_LD * GRAM_LD;
  SMF_CUDA(cudaMalloc(&ctx->scratch, n * sizeof(double)));
  SMF_CUDA(cudaMemsetAsync(ctx->scratch, 0, n * sizeof(double), ctx->stream));
  ctx->scratch_n = n;
  SMF_CUDA(cudaMallocHost(&ctx->host_pinned, 1024 * sizeof(double)));
  ctx->host_pinned_n = 1024;
  return SMF_OK;
}
static double *scr_out(smf_ctx *ctx) { return ctx->scratch; }
static double *scr_partial(smf_ctx *ctx) { return ctx->scratch + SCR_OUT; }
static unsigned int *scr_counter(smf_ctx *ctx) { return (unsigned int *)(ctx->scratch + SCR_OUT + (int64_t)RED_BLOCKS * SCR_PART); }
static double **scr_ptrs(smf_ctx *ctx) { return (double **)(ctx->scratch + SCR_OUT + (int64_t)RED_BLOCKS * SCR_PART + 16); }
static double *scr_gram(smf_ctx *ctx) { return ctx->scratch + SCR_OUT + (int64_t)RED_BLOCKS * SCR_PART + 16 + 256; }

static int all_aligned16(const void *w, double *const *V, int k)
{
  uintptr_t m = (uintptr_t)w;
  for (int j = 0; j < k; ++j) m |= (uintptr_t)V[j];
  return (m & 15u) == 0 ? 1 : 0;
}
// persistent grids: never more blocks than are resident at once (a partial second wave would idle most SMs)
template <class K> static unsigned red_blocks(K kernel, int64_t n)
{
  static int resident_dev[MAX_DEVICES] = {0};
  int per_sm = 1, dev = 0, sms = 148;
  cudaGetDevice(&dev);
  int &resident = resident_dev[dev % MAX_DEVICES];
  if (resident == 0) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, VT, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    resident = per_sm * sms;
    if (resident > RED_BLOCKS) resident = RED_BLOCKS;
  }
  int64_t b = (n + 8 * VT - 1) / (8 * VT);
  if (b < 1) b = 1;
  if (b > resident) b = resident;
  return (unsigned)b;
}

// dots of k (<= MD_CHUNK) device vectors with w, plus w.w; host_out has k+1 entries; synchronises the stream
int vec_multi_dot(smf_ctx *ctx, double *const *V_host_ptrs, int k, const double *w, int64_t n, double *host_out, bool sync)
{
  if (int e = ensure_scratch(ctx)) return e;
  if (k > MD_CHUNK) return ctx->fail(SMF_ERR_INVALID, "multi_dot: too many vectors");
  if (k > 0) SMF_CUDA(cudaMemcpyAsync(scr_ptrs(ctx), V_host_ptrs, sizeof(double *) * k, cudaMemcpyHostToDevice, ctx->stream));
  { ProfScope prof(ctx, ctx->vec_bucket);
  k_multi_dot<false><<<red_blocks(k_multi_dot<false>, n), VT, 0, ctx->stream>>>((const double *const *)scr_ptrs(ctx), k, w, nullptr, n, scr_partial(ctx), scr_counter(ctx), scr_out(ctx),
                                                      ctx->md_vec2 = (all_aligned16(w, V_host_ptrs, k) && ((uintptr_t)ctx->dot_weights & 15u) == 0),
                                                      ctx->dot_weights); }
  if (int e = allreduce_sum(ctx, scr_out(ctx), k + 1)) return e;
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  if (sync) {
    SMF_CUDA(cudaMemcpyAsync(ctx->host_pinned, scr_out(ctx), sizeof(double) * (k + 1), cudaMemcpyDeviceToHost, ctx->stream));
    SMF_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j <= k; ++j) host_out[j] = ctx->host_pinned[j];
  }
  return SMF_OK;
}

// Re-orthogonalised classical Gram-Schmidt in ONE pass over the basis.  CGS2 computes h1 = V^T w, w1 = w - V h1,
// h2 = V^T w1, w2 = w1 - V h2.  Since V^T w1 = (I - G) h1 with the Gram matrix G = V^T V, the second projection is
// known without touching w1: h = h1 + h2 = h1 - E h1, E = G - I.  The column of G that belongs to the newest basis
// vector V[k-1] is measured in the same pass that computes h1 (second right-hand side of k_multi_dot), so one
// orthogonalisation reads the basis twice (dots, update) instead of four times.  On return the device holds the
// corrected coefficients where vec_multi_axpy_from_dots expects them.
__global__ void k_gram_correct(double *E, double *out, int k)
{
  __shared__ double s_h[MD_CHUNK];
  const int m = threadIdx.x, j = k - 1;
  if (m < k) {
    s_h[m] = out[m];
    const double e = out[k + 1 + m] - (m == j ? 1.0 : 0.0);
    E[(size_t)m * GRAM_LD + j] = e;
    E[(size_t)j * GRAM_LD + m] = e;
  }
  __syncthreads();
  if (m < k) {
    double c = 0.0;
    for (int l = 0; l < k; ++l) c = fma(E[(size_t)m * GRAM_LD + l], s_h[l], c);
    out[m] = s_h[m] - c;
  }
}
int vec_multi_dot_gram(smf_ctx *ctx, double *const *V_host_ptrs, int k, const double *w, int64_t n)
{
  if (int e = ensure_scratch(ctx)) return e;
  if (k < 1 || k > MD_CHUNK) return ctx->fail(SMF_ERR_INVALID, "multi_dot_gram: bad number of vectors");
  SMF_CUDA(cudaMemcpyAsync(scr_ptrs(ctx), V_host_ptrs, sizeof(double *) * k, cudaMemcpyHostToDevice, ctx->stream));
  { ProfScope prof(ctx, ctx->vec_bucket);
  k_multi_dot<true><<<red_blocks(k_multi_dot<true>, n), VT, 0, ctx->stream>>>((const double *const *)scr_ptrs(ctx), k, w, V_host_ptrs[k - 1], n, scr_partial(ctx), scr_counter(ctx), scr_out(ctx),
                                                      ctx->md_vec2 = (all_aligned16(w, V_host_ptrs, k) && ((uintptr_t)ctx->dot_weights & 15u) == 0),
                                                      ctx->dot_weights); }
  if (int e = allreduce_sum(ctx, scr_out(ctx), 2 * k + 1)) return e;
  k_gram_correct<<<1, 256, 0, ctx->stream>>>(scr_gram(ctx), scr_out(ctx), k);
  ctx->launches += 2;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}

// w -= sum_j h_j V_j with h = the first k results of the preceding vec_multi_dot (still on the device);
// returns h (k entries), the pre-orthogonalisation |w|^2 (entry k) and the new |w|^2 (entry k+1) on the host.
int vec_multi_axpy_from_dots(smf_ctx *ctx, double *w, int k, int64_t n, double *host_out)
{
  // pointer table already uploaded by vec_multi_dot
  { ProfScope prof(ctx, ctx->vec_bucket);
  k_multi_axpy<<<red_blocks(k_multi_axpy, n), VT, 0, ctx->stream>>>(w, (const double *const *)scr_ptrs(ctx), scr_out(ctx), k, n,
                                                       scr_partial(ctx), scr_counter(ctx) + 1, scr_out(ctx) + OUT_NRM,
                                                       ctx->md_vec2 && ((uintptr_t)w & 15u) == 0, ctx->dot_weights); }
  if (int e = allreduce_sum(ctx, scr_out(ctx) + OUT_NRM, 1)) return e;
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  SMF_CUDA(cudaMemcpyAsync(ctx->host_pinned, scr_out(ctx), sizeof(double) * (OUT_NRM + 1), cudaMemcpyDeviceToHost, ctx->stream));
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int j = 0; j <= k; ++j) host_out[j] = ctx->host_pinned[j];
  host_out[k + 1] = ctx->host_pinned[OUT_NRM];
  return SMF_OK;
}

// x += sum_j coef[j] V_j   (coefficients from the host)
int vec_multi_add(smf_ctx *ctx, double *x, double *const *V_host_ptrs, const double *coef_host, int k, int64_t n)
{
  if (int e = ensure_scratch(ctx)) return e;
  if (k > MD_CHUNK) return ctx->fail(SMF_ERR_INVALID, "multi_add: too many vectors");
  if (k == 0) return SMF_OK;
  for (int j = 0; j < k; ++j) ctx->host_pinned[512 + j] = -coef_host[j];
  SMF_CUDA(cudaMemcpyAsync(scr_ptrs(ctx), V_host_ptrs, sizeof(double *) * k, cudaMemcpyHostToDevice, ctx->stream));
  SMF_CUDA(cudaMemcpyAsync(scr_out(ctx), ctx->host_pinned + 512, sizeof(double) * k, cudaMemcpyHostToDevice, ctx->stream));
  { ProfScope prof(ctx, ctx->vec_bucket);
  k_multi_axpy<<<red_blocks(k_multi_axpy, n), VT, 0, ctx->stream>>>(x, (const double *const *)scr_ptrs(ctx), scr_out(ctx), k, n,
                                                       scr_partial(ctx), scr_counter(ctx) + 1, scr_out(ctx) + OUT_NRM,
                                                       all_aligned16(x, V_host_ptrs, k), nullptr); }
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  // the pinned staging area is reused by the next call: make sure the copies above have been consumed
  SMF_CUDA(cudaStreamSynchronize(ctx->stream));
  return SMF_OK;
}

int vec_dot(smf_ctx *ctx, const double *x, const double *y, int64_t n, double *host_result)
{
  if (int e = ensure_scratch(ctx)) return e;
  double *ptr = const_cast<double *>(x);
  double out[2];
  if (int e = vec_multi_dot(ctx, &ptr, 1, y, n, out, true)) return e;
  *host_result = out[0];
  return SMF_OK;
}
int vec_norm2(smf_ctx *ctx, const double *x, int64_t n, double *host_result)
{
  double out[1];
  if (int e = vec_multi_dot(ctx, nullptr, 0, x, n, out, true)) return e;
  *host_result = out[0];
  return SMF_OK;
}

}  // namespace smf

// ---- measured FP64 peak (every kernel is reported against the HBM roofline AND the FP64 pipe) ----
// 8 independent DFMA chains per thread, 4 resident blocks of 256 threads per SM; the result is written once so that the
// compiler keeps the chains.  smf_measure_fp64_peak times `reps` launches with CUDA events on the context's stream.
__global__ void __launch_bounds__(256) k_fp64_fma(double *out, double a, double b, int iters)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // (never true: a = 1 - 2^-20, b = 2^-30 keep the chains bounded)
}

extern "C" int smf_measure_fp64_peak(smf_ctx *ctx, double *tflops_out)
{
  if (!ctx || !tflops_out) return SMF_ERR_INVALID;
  int dev = 0, sms = 0;
  SMF_CUDA(cudaGetDevice(&dev));
  SMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256, iters = 2048, reps = 5;
  double *out = nullptr;
  SMF_CUDA(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaEvent_t a, b;
  SMF_CUDA(cudaEventCreate(&a)); SMF_CUDA(cudaEventCreate(&b));
  const double fa = 1.0 - 1.0 / 1048576.0, fb = 1.0 / 1073741824.0;
  k_fp64_fma<<<blocks, threads, 0, ctx->stream>>>(out, fa, fb, iters);      // warm-up
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    SMF_CUDA(cudaEventRecord(a, ctx->stream));
    k_fp64_fma<<<blocks, threads, 0, ctx->stream>>>(out, fa, fb, iters);
    SMF_CUDA(cudaEventRecord(b, ctx->stream));
    SMF_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    SMF_CUDA(cudaEventElapsedTime(&ms, a, b));
    best = ms < best ? ms : best;
    ctx->launches++;
  }
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
  const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
  *tflops_out = flops / ((double)best * 1e-3) / 1e12;
  return SMF_OK;
}
