// Matrix-free Q2-Q1 cell kernels for sm_100a (CUDA cores, FP64).
//
// Replaces the cell loops of /root/reference/source/simulator/solver/matrix_free_operators.cc
//   StokesOperator::local_apply :376-547, BTBlockOperator::local_apply :668-701,
//   MassMatrixOperator::local_apply :787-845, ABlockOperator::local_apply :974-1052,
//   compute_diagonal :861-908 / :1068-1098,
// and deal.II's FEEvaluation / MGTransferMatrixFree kernels underneath them.
//
// Design: a thread block stages the unique DoFs of a batch of 64 consecutive cells in
// shared memory (coalesced gather, constraints resolved while loading), then ONE THREAD PER
// (cell, velocity component) runs the whole sum factorisation of its 3^dim values in registers.  The only
// cross-thread traffic is the symmetric off-diagonal stress, exchanged through shared memory between the
// dim warps of the block.  Results are accumulated in the shared staging area in collision-free rounds and
// flushed once per batch: nodes private to the batch with plain stores, shared nodes with one RED each.
#include "smf_internal.h"

namespace smf {

// ---- reference-element constants ---------------------------------------------------------
#define SMF_SQ 0.7745966692414834   /* sqrt(3/5) */
__device__ constexpr double NA = 0.5 * SMF_SQ + 0.3;     // phi_0(x_0) = phi_2(x_2)
__device__ constexpr double NB = 0.4;                    // phi_1(x_0) = phi_1(x_2)
__device__ constexpr double NC = 0.3 - 0.5 * SMF_SQ;     // phi_2(x_0) = phi_0(x_2)
__device__ constexpr double M0 = 0.5 + 0.5 * SMF_SQ;     // Q1: 1 - x_0
__device__ constexpr double M1 = 0.5 - 0.5 * SMF_SQ;     // Q1: x_0
constexpr double HOST_GAUSS_A = 0.5 * SMF_SQ;            // collocation derivative = D' / (h * a)

__host__ __device__ constexpr double gw(int i) { return i == 1 ? 8.0 / 18.0 : 5.0 / 18.0; }
template <int DIM> __host__ __device__ constexpr double qweight(int q)
{ return DIM == 3 ? gw(q % 3) * gw((q / 3) % 3) * gw(q / 9) : gw(q % 3) * gw(q / 3); }
template <int DIM> __host__ __device__ constexpr int nn() { return DIM == 3 ? 27 : 9; }
template <int DIM> __host__ __device__ constexpr int np() { return DIM == 3 ? 8 : 4; }
__host__ __device__ constexpr int stride3(int d) { return d == 0 ? 1 : (d == 1 ? 3 : 9); }
__host__ __device__ constexpr int stride2(int d) { return d == 0 ? 1 : (d == 1 ? 2 : 4); }

// Visit every line of a 3^DIM tensor along direction DIR: f(i0, i1, i2) with the three register indices.
template <int DIM, int DIR, class F> __device__ __forceinline__ void for_lines(F f)
{
  constexpr int S = stride3(DIR);
#pragma unroll
  for (int n = 0; n < nn<DIM>(); ++n)
    if ((n / S) % 3 == 0) f(n, n + S, n + 2 * S);
}

// nodal Q2 -> Gauss points along DIR, in place (middle row of the interpolation matrix is (0,1,0))
template <int DIM, int DIR> __device__ __forceinline__ void interp(double (&v)[nn<DIM>()])
{
  for_lines<DIM, DIR>([&](int a, int b, int c) {
    const double t0 = NA * v[a] + NB * v[b] + NC * v[c];
    const double t2 = NC * v[a] + NB * v[b] + NA * v[c];
    v[a] = t0; v[c] = t2;
  });
}
template <int DIM> __device__ __forceinline__ void interp_all(double (&v)[nn<DIM>()])
{ interp<DIM, 0>(v); interp<DIM, 1>(v); if (DIM == 3) interp<DIM, (DIM == 3 ? 2 : 0)>(v); }

// transpose of the above (integration against the test functions)
template <int DIM, int DIR> __device__ __forceinline__ void interpT(double (&v)[nn<DIM>()])
{
  for_lines<DIM, DIR>([&](int a, int b, int c) {
    const double t0 = NA * v[a] + NC * v[c];
    const double t1 = NB * (v[a] + v[c]) + v[b];
    const double t2 = NC * v[a] + NA * v[c];
    v[a] = t0; v[b] = t1; v[c] = t2;
  });
}
template <int DIM> __device__ __forceinline__ void interpT_all(double (&v)[nn<DIM>()])
{ interpT<DIM, 0>(v); interpT<DIM, 1>(v); if (DIM == 3) interpT<DIM, (DIM == 3 ? 2 : 0)>(v); }

// g = D' u along DIR (collocation derivative times a*h): rows (-1.5,2,-.5), (-.5,0,.5), (.5,-2,1.5)
template <int DIM, int DIR> __device__ __forceinline__ void deriv(const double (&u)[nn<DIM>()], double (&g)[nn<DIM>()])
{
  for_lines<DIM, DIR>([&](int a, int b, int c) {
    g[a] = -1.5 * u[a] + 2.0 * u[b] - 0.5 * u[c];
    g[b] = 0.5 * (u[c] - u[a]);
    g[c] = 0.5 * u[a] - 2.0 * u[b] + 1.5 * u[c];
  });
}
// T += s * D'^T x along DIR
template <int DIM, int DIR> __device__ __forceinline__ void derivT_add(double (&T)[nn<DIM>()], const double (&x)[nn<DIM>()], double s)
{
  for_lines<DIM, DIR>([&](int a, int b, int c) {
    const double xa = s * x[a], xb = s * x[b], xc = s * x[c];
    T[a] += -1.5 * xa - 0.5 * xb + 0.5 * xc;
    T[b] += 2.0 * (xa - xc);
    T[c] += -0.5 * xa + 0.5 * xb + 1.5 * xc;
  });
}
template <int DIM> __device__ __forceinline__ void deriv_dir(int dir, const double (&u)[nn<DIM>()], double (&g)[nn<DIM>()])
{
  if (dir == 0) deriv<DIM, 0>(u, g);
  else if (dir == 1) deriv<DIM, 1>(u, g);
  else deriv<DIM, (DIM == 3 ? 2 : 0)>(u, g);
}
template <int DIM> __device__ __forceinline__ void derivT_add_dir(int dir, double (&T)[nn<DIM>()], const double (&x)[nn<DIM>()], double s)
{
  if (dir == 0) derivT_add<DIM, 0>(T, x, s);
  else if (dir == 1) derivT_add<DIM, 1>(T, x, s);
  else derivT_add<DIM, (DIM == 3 ? 2 : 0)>(T, x, s);
}

// Q1 nodal (2^DIM) -> Gauss points (3^DIM)
template <int DIM> __device__ __forceinline__ void q1_to_quad(const double (&p)[np<DIM>()], double (&pq)[nn<DIM>()])
{
  if (DIM == 2) {
    double t[6];   // [j][q]
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double a = p[2 * j], b = p[2 * j + 1];
      t[3 * j] = M0 * a + M1 * b; t[3 * j + 1] = 0.5 * (a + b); t[3 * j + 2] = M1 * a + M0 * b;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double a = t[q], b = t[3 + q];
      pq[q] = M0 * a + M1 * b; pq[3 + q] = 0.5 * (a + b); pq[6 + q] = M1 * a + M0 * b;
    }
  } else {
    double t1[12];  // [k][j][qx]
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const double a = p[2 * l], b = p[2 * l + 1];
      t1[3 * l] = M0 * a + M1 * b; t1[3 * l + 1] = 0.5 * (a + b); t1[3 * l + 2] = M1 * a + M0 * b;
    }
    double t2[18];  // [k][qy][qx]
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double a = t1[6 * k + q], b = t1[6 * k + 3 + q];
        t2[9 * k + q] = M0 * a + M1 * b; t2[9 * k + 3 + q] = 0.5 * (a + b); t2[9 * k + 6 + q] = M1 * a + M0 * b;
      }
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const double a = t2[q], b = t2[9 + q];
      pq[q] = M0 * a + M1 * b; pq[9 + q] = 0.5 * (a + b); pq[18 + q] = M1 * a + M0 * b;
    }
  }
}
// transpose: Gauss points -> Q1 test functions
template <int DIM> __device__ __forceinline__ void quad_to_q1T(const double (&vq)[nn<DIM>()], double (&out)[np<DIM>()])
{
  if (DIM == 2) {
    double t[6];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double a = vq[q], b = vq[3 + q], c = vq[6 + q];
      t[q] = M0 * a + 0.5 * b + M1 * c; t[3 + q] = M1 * a + 0.5 * b + M0 * c;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double a = t[3 * j], b = t[3 * j + 1], c = t[3 * j + 2];
      out[2 * j] = M0 * a + 0.5 * b + M1 * c; out[2 * j + 1] = M1 * a + 0.5 * b + M0 * c;
    }
  } else {
    double t2[18];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const double a = vq[q], b = vq[9 + q], c = vq[18 + q];
      t2[q] = M0 * a + 0.5 * b + M1 * c; t2[9 + q] = M1 * a + 0.5 * b + M0 * c;
    }
    double t1[12];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double a = t2[9 * k + q], b = t2[9 * k + 3 + q], c = t2[9 * k + 6 + q];
        t1[6 * k + q] = M0 * a + 0.5 * b + M1 * c; t1[6 * k + 3 + q] = M1 * a + 0.5 * b + M0 * c;
      }
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const double a = t1[3 * l], b = t1[3 * l + 1], c = t1[3 * l + 2];
      out[2 * l] = M0 * a + 0.5 * b + M1 * c; out[2 * l + 1] = M1 * a + 0.5 * b + M0 * c;
    }
  }
}

// scatter round of a local node: nodes of one round never alias inside a sub-batch (checked at set-up)
template <int DIM> __host__ __device__ constexpr int round_q2(int n)
{ return DIM == 3 ? ((n % 3) == 2) + 2 * (((n / 3) % 3) == 2) + 4 * ((n / 9) == 2) : ((n % 3) == 2) + 2 * ((n / 3) == 2); }

// ---------------------------------------------------------------------------------------------------------
struct CellParams {
  const int *batch_info;
  const uint32_t *node_list_u; const uint16_t *cell_loc_u;
  const uint32_t *node_list_p; const uint16_t *cell_loc_p;
  int max_nodes_u, max_nodes_p;
  const double *eta_t; int n_eta;     // lane-transposed table [batch][k][n_eta][32]
  double ee[3];                        // 1 / (h_d * a)
  double vol, ps;
  const double *src_u, *src_p;
  double *dst_u, *dst_p;
};

template <int DIM> __host__ __device__ constexpr int npair() { return DIM == 3 ? 3 : 1; }
template <int DIM> size_t vel_smem_bytes(int max_nodes_u, int max_nodes_p, int mode)
{
  size_t d = 2 * (size_t)DIM * max_nodes_u + (size_t)npair<DIM>() * nn<DIM>() * 32;
  if (mode != 0) d += 2 * (size_t)max_nodes_p + (size_t)DIM * np<DIM>() * 32;
  return d * sizeof(double);
}

// MODE 0: A block (u -> u); MODE 1: Stokes ((u,p) -> (u,p)); MODE 2: B^T (p -> u)
template <int DIM, int MODE>
__global__ void __launch_bounds__(32 * DIM, DIM == 3 ? 2 : 4) velocity_kernel(const CellParams P)
{
  constexpr int NN = nn<DIM>(), NP = np<DIM>(), NT = 32 * DIM;
  extern __shared__ double smem[];
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x;
  const int *info = P.batch_info + (size_t)b * BATCH_INFO;
  const int n_nodes_u = info[0], n_excl_u = info[1], n_nodes_p = info[2], n_excl_p = info[3], n_cells_b = info[4];
  const int MU = P.max_nodes_u, MP = P.max_nodes_p;
  double *s_src = smem;                       // [DIM][MU]
  double *s_dst = s_src + DIM * MU;           // [DIM][MU]
  double *s_ex = s_dst + DIM * MU;            // [npair][NN][32]
  double *s_srcp = s_ex + npair<DIM>() * NN * 32;   // [MP]
  double *s_dstp = s_srcp + MP;               // [MP]
  double *s_pex = s_dstp + MP;                // [DIM][NP][32]

  const uint32_t *nlu = P.node_list_u + (size_t)b * MU;
  const uint32_t *nlp = P.node_list_p + (size_t)b * MP;
  // ---- stage the batch's unique DoFs (constrained DoFs read as zero: read_dof_values semantics) ----
  for (int j = threadIdx.x; j < n_nodes_u; j += NT) {
    const uint32_t e = nlu[j];
    const uint32_t idx = e & NODE_IDX_MASK;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      if (MODE != 2) s_src[d * MU + j] = ((e >> (NODE_FLAG_SHIFT + d)) & 1u) ? 0.0 : P.src_u[idx + d];
      s_dst[d * MU + j] = 0.0;
    }
  }
  if (MODE != 0)
    for (int j = threadIdx.x; j < n_nodes_p; j += NT) {
      const uint32_t e = nlp[j];
      s_srcp[j] = (e >> 31) ? 0.0 : P.src_p[e & NODE_IDX_MASK];
      s_dstp[j] = 0.0;
    }
  __syncthreads();

  const double ec = P.ee[c];
  for (int k = 0; k < KSUB; ++k) {
    const bool active = (k * SUB + lane) < n_cells_b;
    const uint16_t *loc = P.cell_loc_u + (((size_t)b * KSUB + k) * NN) * SUB + lane;
    const uint16_t *locp = P.cell_loc_p + (((size_t)b * KSUB + k) * NP) * SUB + lane;
    const double *etab = P.eta_t + ((size_t)b * KSUB + k) * P.n_eta * SUB + lane;
    const int eq = (P.n_eta == 1) ? 0 : SUB;           // stride between quadrature points in the table
    double u[NN], T[NN];

    if (MODE != 2) {
#pragma unroll
      for (int n = 0; n < NN; ++n) u[n] = s_src[c * MU + loc[n * SUB]];
      interp_all<DIM>(u);
    }
    // ---- diagonal direction: T = D'_c^T [ 2 eta w vol e_c^2 g_cc - e_c ps w vol p ] ----
    if (MODE != 0) {
      double pn[NP], pq[NN];
#pragma unroll
      for (int i = 0; i < NP; ++i) pn[i] = s_srcp[locp[i * SUB]];
      q1_to_quad<DIM>(pn, pq);
      const double f = -ec * P.ps * P.vol;
#pragma unroll
      for (int q = 0; q < NN; ++q) T[q] = f * qweight<DIM>(q) * pq[q];
    } else {
#pragma unroll
      for (int q = 0; q < NN; ++q) T[q] = 0.0;
    }
    if (MODE != 2) {
      double g[NN];
      deriv_dir<DIM>(c, u, g);
      if (MODE == 1) {   // pressure rows: -ps (div u, q): this component's share, reduced below by warp 0
        double dq[NN], y[NP];
        const double f = -P.ps * P.vol * ec;
#pragma unroll
        for (int q = 0; q < NN; ++q) dq[q] = f * qweight<DIM>(q) * g[q];
        quad_to_q1T<DIM>(dq, y);
#pragma unroll
        for (int i = 0; i < NP; ++i) s_pex[(c * NP + i) * SUB + lane] = y[i];
      }
      const double f2 = 2.0 * ec * ec * P.vol;
#pragma unroll
      for (int q = 0; q < NN; ++q) T[q] += f2 * qweight<DIM>(q) * etab[q * eq] * g[q];
    }
    {
      double Tc[NN];
#pragma unroll
      for (int q = 0; q < NN; ++q) { Tc[q] = T[q]; T[q] = 0.0; }
      derivT_add_dir<DIM>(c, T, Tc, 1.0);
    }
    // ---- off-diagonal pairs {c,d}: X = eta w vol (e_d g_cd + e_c g_dc) is symmetric; the lower thread
    //      publishes g, the higher one forms X, uses it and publishes it back ----
    if (MODE != 2) {
#pragma unroll
      for (int d = 1; d < DIM; ++d)
        if (c < d) {
          double g[NN];
          deriv_dir<DIM>(d, u, g);
          double *slot = s_ex + (size_t)(c + d - 1) * NN * SUB + lane;
#pragma unroll
          for (int q = 0; q < NN; ++q) slot[q * SUB] = g[q];
        }
    }
    __syncthreads();
    if (MODE == 1 && c == 0) {   // pressure scatter (one warp: collision-free rounds over the 2^dim nodes)
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        double y = s_pex[i * SUB + lane];
#pragma unroll
        for (int d = 1; d < DIM; ++d) y += s_pex[(d * NP + i) * SUB + lane];
        if (active) s_dstp[locp[i * SUB]] += y;
        __syncwarp();
      }
    }
    if (MODE != 2) {
#pragma unroll
      for (int d = 0; d < DIM - 1; ++d)
        if (d < c) {
          double g[NN];
          deriv_dir<DIM>(d, u, g);           // g_cd = D'_d Uq_c
          double *slot = s_ex + (size_t)(c + d - 1) * NN * SUB + lane;
          const double ed = P.ee[d], vol = P.vol;
#pragma unroll
          for (int q = 0; q < NN; ++q) {
            const double x = vol * qweight<DIM>(q) * etab[q * eq] * (ed * g[q] + ec * slot[q * SUB]);
            slot[q * SUB] = x;
            g[q] = x;
          }
          derivT_add_dir<DIM>(d, T, g, ed);
        }
    }
    __syncthreads();
    if (MODE != 2) {
#pragma unroll
      for (int d = 1; d < DIM; ++d)
        if (c < d) {
          double x[NN];
          const double *slot = s_ex + (size_t)(c + d - 1) * NN * SUB + lane;
#pragma unroll
          for (int q = 0; q < NN; ++q) x[q] = slot[q * SUB];
          derivT_add_dir<DIM>(d, T, x, P.ee[d]);
        }
    }
    interpT_all<DIM>(T);
    // ---- accumulate into the staging area, one collision-free round at a time ----
#pragma unroll
    for (int r = 0; r < (1 << DIM); ++r) {
#pragma unroll
      for (int n = 0; n < NN; ++n)
        if (round_q2<DIM>(n) == r && active) s_dst[c * MU + loc[n * SUB]] += T[n];
      __syncwarp();
    }
  }
  __syncthreads();
  // ---- flush: private nodes with plain stores, shared nodes with one RED per DoF; constrained DoFs get the
  //      unit diagonal of MatrixFreeOperators::Base (dst = src) ----
  for (int j = threadIdx.x; j < n_nodes_u; j += NT) {
    const uint32_t e = nlu[j];
    const uint32_t idx = e & NODE_IDX_MASK;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      if ((e >> (NODE_FLAG_SHIFT + d)) & 1u) P.dst_u[idx + d] = (MODE == 2) ? 0.0 : P.src_u[idx + d];
      else if (j < n_excl_u) P.dst_u[idx + d] = s_dst[d * MU + j];
      else atomicAdd(P.dst_u + idx + d, s_dst[d * MU + j]);
    }
  }
  if (MODE == 1)
    for (int j = threadIdx.x; j < n_nodes_p; j += NT) {
      const uint32_t e = nlp[j];
      const uint32_t idx = e & NODE_IDX_MASK;
      if (e >> 31) P.dst_p[idx] = P.src_p[idx];
      else if (j < n_excl_p) P.dst_p[idx] = s_dstp[j];
      else atomicAdd(P.dst_p + idx, s_dstp[j]);
    }
}

// ---- pressure mass matrix: one thread per cell, two warps (= two sub-batches) per block -------------------
template <int DIM>
__global__ void __launch_bounds__(BATCH_CELLS) mass_kernel(const CellParams P)
{
  constexpr int NN = nn<DIM>(), NP = np<DIM>();
  extern __shared__ double smem[];
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x;
  const int *info = P.batch_info + (size_t)b * BATCH_INFO;
  const int n_nodes_p = info[2], n_excl_p = info[3], n_cells_b = info[4];
  const int MP = P.max_nodes_p;
  double *s_src = smem;            // [MP]
  double *s_dst = smem + MP;       // [KSUB][MP]
  const uint32_t *nlp = P.node_list_p + (size_t)b * MP;
  for (int j = threadIdx.x; j < n_nodes_p; j += BATCH_CELLS) {
    const uint32_t e = nlp[j];
    s_src[j] = (e >> 31) ? 0.0 : P.src_p[e & NODE_IDX_MASK];
#pragma unroll
    for (int kk = 0; kk < KSUB; ++kk) s_dst[kk * MP + j] = 0.0;
  }
  __syncthreads();
  const bool active = (k * SUB + lane) < n_cells_b;
  const uint16_t *locp = P.cell_loc_p + (((size_t)b * KSUB + k) * NP) * SUB + lane;
  const double *etab = P.eta_t + ((size_t)b * KSUB + k) * P.n_eta * SUB + lane;
  double pn[NP], pq[NN], y[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) pn[i] = s_src[locp[i * SUB]];
  q1_to_quad<DIM>(pn, pq);
  const double f = P.ps * P.ps * P.vol;
  if (P.n_eta == 1) {
    const double ce = active ? f / etab[0] : 0.0;
#pragma unroll
    for (int q = 0; q < NN; ++q) pq[q] *= ce * qweight<DIM>(q);
  } else {
#pragma unroll
    for (int q = 0; q < NN; ++q) pq[q] *= active ? (f / etab[q * SUB]) * qweight<DIM>(q) : 0.0;
  }
  quad_to_q1T<DIM>(pq, y);
#pragma unroll
  for (int i = 0; i < NP; ++i) {     // node i of different cells never alias: one round per node
    if (active) s_dst[k * MP + locp[i * SUB]] += y[i];
    __syncwarp();
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n_nodes_p; j += BATCH_CELLS) {
    const uint32_t e = nlp[j];
    const uint32_t idx = e & NODE_IDX_MASK;
    double v = s_dst[j];
#pragma unroll
    for (int kk = 1; kk < KSUB; ++kk) v += s_dst[kk * MP + j];
    if (e >> 31) P.dst_p[idx] = P.src_p[idx];
    else if (j < n_excl_p) P.dst_p[idx] = v;
    else atomicAdd(P.dst_p + idx, v);
  }
}

// ---- diagonals (MatrixFreeTools::compute_diagonal): closed form of "apply the cell operator to e_i" ---------
// A: diag(c, i) = sum_q w vol eta_q [ sum_d (d_d phi_i)^2 + (d_c phi_i)^2 ]  (- (2/3) eta (d_c phi_i)^2 if compressible)
// M: diag(i)    = sum_q w vol (ps^2 / eta_q) psi_i^2
struct DiagParams {
  int64_t n_cells;
  const uint32_t *cell_nodes_u, *cell_nodes_p;
  const double *eta; int n_eta;     // lane-transposed table [cell/32][n_eta][32]
  double ih[3]; double vol, ps; int compressible;
  double *diag_A, *diag_M;
};
__device__ __forceinline__ double q2v(int q, int i)
{ return q == 1 ? (i == 1 ? 1.0 : 0.0) : (i == 1 ? NB : ((i == 0) == (q == 0) ? NA : NC)); }
__device__ __forceinline__ double q2d(int q, int i)   // d phi_i / dx at Gauss point q (unit cell)
{
  const double x = q == 0 ? 0.5 - 0.5 * SMF_SQ : (q == 1 ? 0.5 : 0.5 + 0.5 * SMF_SQ);
  return i == 0 ? 4 * x - 3 : (i == 1 ? 4 - 8 * x : 4 * x - 1);
}
__device__ __forceinline__ double q1v(int q, int i)
{ return q == 1 ? 0.5 : ((i == 0) == (q == 0) ? M0 : M1); }

template <int DIM> __global__ void diag_kernel(const DiagParams P)
{
  constexpr int NN = nn<DIM>(), NP = np<DIM>();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t cell = t / NN;
  const int i = (int)(t % NN);
  if (cell >= P.n_cells) return;
  int id[3] = {i % 3, (i / 3) % 3, DIM == 3 ? i / 9 : 0};
  double a[3] = {0, 0, 0};
  for (int q = 0; q < NN; ++q) {
    int qd[3] = {q % 3, (q / 3) % 3, DIM == 3 ? q / 9 : 0};
    const double coef = qweight<DIM>(q) * P.vol * P.eta[((cell / SUB) * P.n_eta + (P.n_eta == 1 ? 0 : q)) * SUB + (cell % SUB)];
    for (int d = 0; d < DIM; ++d) {
      double g = P.ih[d];
      for (int e = 0; e < DIM; ++e) g *= (e == d) ? q2d(qd[e], id[e]) : q2v(qd[e], id[e]);
      a[d] += coef * g * g;
    }
  }
  double sum = 0;
  for (int d = 0; d < DIM; ++d) sum += a[d];
  const uint32_t e = P.cell_nodes_u[cell * NN + i];
  const uint32_t idx = e & NODE_IDX_MASK;
  for (int c = 0; c < DIM; ++c) {
    if ((e >> (NODE_FLAG_SHIFT + c)) & 1u) continue;
    double v = sum + a[c];
    if (P.compressible) v -= (2.0 / 3.0) * a[c];
    atomicAdd(P.diag_A + idx + c, v);
  }
  if (i < NP) {
    int jd[3] = {i % 2, (i / 2) % 2, DIM == 3 ? i / 4 : 0};
    double m = 0;
    for (int q = 0; q < NN; ++q) {
      int qd[3] = {q % 3, (q / 3) % 3, DIM == 3 ? q / 9 : 0};
      double v = 1.0;
      for (int d = 0; d < DIM; ++d) v *= q1v(qd[d], jd[d]);
      m += qweight<DIM>(q) * P.vol * (P.ps * P.ps / P.eta[((cell / SUB) * P.n_eta + (P.n_eta == 1 ? 0 : q)) * SUB + (cell % SUB)]) * v * v;
    }
    const uint32_t ep = P.cell_nodes_p[cell * NP + i];
    if (!(ep >> 31)) atomicAdd(P.diag_M + (ep & NODE_IDX_MASK), m);
  }
}

__global__ void invert_diag_kernel(double *d, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = 1.0 / d[i];
}
__global__ void set_entries_kernel(double *d, const uint32_t *idx, int64_t n, double v)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[idx[i]] = v;
}

// ---- multigrid transfer (MGTransferMatrixFree, uniform refinement) --------------------------------------
// 1D prolongation rows for the child patch: Q2 (5x3) and Q1 (3x2)
__device__ __forceinline__ double p2row(int r, int j)
{
  return r == 0 ? (j == 0 ? 1.0 : 0.0) : r == 1 ? (j == 0 ? 0.375 : (j == 1 ? 0.75 : -0.125))
       : r == 2 ? (j == 1 ? 1.0 : 0.0) : r == 3 ? (j == 0 ? -0.125 : (j == 1 ? 0.75 : 0.375)) : (j == 2 ? 1.0 : 0.0);
}
__device__ __forceinline__ double p1row(int r, int j)
{ return r == 0 ? (j == 0 ? 1.0 : 0.0) : r == 1 ? 0.5 : (j == 1 ? 1.0 : 0.0); }

struct TransferParams {
  int64_t n_parents;
  const uint32_t *patch;          // [n_parents][NPATCH] fine node | owner bit
  const uint32_t *coarse_nodes;   // [n_parents][NC] coarse node | constraint flags
  const double *src; double *dst;
  int add;
};

// one thread per (parent, fine patch entry): interpolate the coarse cell function at that patch node; only the
// owning parent writes (all parents would write the same value), so no atomics and no weights are needed.
template <int DIM, int NPD, int NCOMP> __global__ void prolongate_kernel(const TransferParams P)
{
  constexpr int NPP = 2 * NPD - 1;
  constexpr int NPATCH = DIM == 3 ? NPP * NPP * NPP : NPP * NPP;
  constexpr int NC = DIM == 3 ? NPD * NPD * NPD : NPD * NPD;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t par = t / NPATCH;
  const int s = (int)(t % NPATCH);
  if (par >= P.n_parents) return;
  const uint32_t pe = P.patch[par * NPATCH + s];
  if (!(pe >> 31)) return;
  const int sd[3] = {s % NPP, (s / NPP) % NPP, DIM == 3 ? s / (NPP * NPP) : 0};
  double acc[NCOMP];
  for (int c = 0; c < NCOMP; ++c) acc[c] = 0.0;
  for (int j = 0; j < NC; ++j) {
    const int jd[3] = {j % NPD, (j / NPD) % NPD, DIM == 3 ? j / (NPD * NPD) : 0};
    double w = 1.0;
    for (int d = 0; d < DIM; ++d) w *= (NPD == 3) ? p2row(sd[d], jd[d]) : p1row(sd[d], jd[d]);
    if (w == 0.0) continue;
    const uint32_t ce = P.coarse_nodes[par * NC + j];
    const uint32_t ci = ce & NODE_IDX_MASK;
    for (int c = 0; c < NCOMP; ++c) {
      const bool cons = (NCOMP == 1) ? (ce >> 31) : ((ce >> (NODE_FLAG_SHIFT + c)) & 1u);
      if (!cons) acc[c] += w * P.src[ci + c];
    }
  }
  const uint32_t fi = pe & NODE_IDX_MASK;
  for (int c = 0; c < NCOMP; ++c) {
    if (P.add) P.dst[fi + c] += acc[c];
    else P.dst[fi + c] = acc[c];
  }
}

// one thread per (parent, coarse node): gather the owned fine patch values with the transposed weights and add
// into the coarse vector (coarse Dirichlet DoFs stay untouched = zero).
template <int DIM, int NPD, int NCOMP> __global__ void restrict_kernel(const TransferParams P)
{
  constexpr int NPP = 2 * NPD - 1;
  constexpr int NPATCH = DIM == 3 ? NPP * NPP * NPP : NPP * NPP;
  constexpr int NC = DIM == 3 ? NPD * NPD * NPD : NPD * NPD;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t par = t / NC;
  const int j = (int)(t % NC);
  if (par >= P.n_parents) return;
  const int jd[3] = {j % NPD, (j / NPD) % NPD, DIM == 3 ? j / (NPD * NPD) : 0};
  double acc[NCOMP];
  for (int c = 0; c < NCOMP; ++c) acc[c] = 0.0;
  for (int s = 0; s < NPATCH; ++s) {
    const int sd[3] = {s % NPP, (s / NPP) % NPP, DIM == 3 ? s / (NPP * NPP) : 0};
    double w = 1.0;
    for (int d = 0; d < DIM; ++d) w *= (NPD == 3) ? p2row(sd[d], jd[d]) : p1row(sd[d], jd[d]);
    if (w == 0.0) continue;
    const uint32_t pe = P.patch[par * NPATCH + s];
    if (!(pe >> 31)) continue;
    const uint32_t fi = pe & NODE_IDX_MASK;
    for (int c = 0; c < NCOMP; ++c) acc[c] += w * P.src[fi + c];
  }
  const uint32_t ce = P.coarse_nodes[par * NC + j];
  const uint32_t ci = ce & NODE_IDX_MASK;
  for (int c = 0; c < NCOMP; ++c) {
    const bool cons = (NCOMP == 1) ? (ce >> 31) : ((ce >> (NODE_FLAG_SHIFT + c)) & 1u);
    if (!cons && acc[c] != 0.0) atomicAdd(P.dst + ci + c, acc[c]);
  }
}

// ---------------------------------------------------------------------------------------------------------
static CellParams make_params(const LevelDev &L)
{
  CellParams P{};
  P.batch_info = L.batch_info;
  P.node_list_u = L.node_list_u; P.cell_loc_u = L.cell_loc_u;
  P.node_list_p = L.node_list_p; P.cell_loc_p = L.cell_loc_p;
  P.max_nodes_u = L.max_nodes_u; P.max_nodes_p = L.max_nodes_p;
  P.eta_t = L.eta; P.n_eta = L.n_eta;
  double vol = 1.0;
  for (int d = 0; d < L.dim; ++d) { P.ee[d] = 1.0 / (L.h[d] * HOST_GAUSS_A); vol *= L.h[d]; }
  P.vol = vol; P.ps = L.ps;
  return P;
}

template <int DIM, int MODE> static int launch_velocity(smf_ctx *ctx, LevelDev &L, CellParams &P)
{
  const size_t smem = vel_smem_bytes<DIM>(L.max_nodes_u, L.max_nodes_p, MODE);
  static size_t configured = 0;
  if (smem > configured) {
    SMF_CUDA(cudaFuncSetAttribute(velocity_kernel<DIM, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const bool fine = (&L == &ctx->lev(SMF_ACTIVE)) || (&L == &ctx->levels[ctx->cfg.n_levels - 1]);
  {
    ProfScope prof(ctx, MODE == 1 ? B_STOKES : (MODE == 2 ? B_BT : (fine ? B_A_FINE : B_A_COARSE)));
    velocity_kernel<DIM, MODE><<<L.n_batches, 32 * DIM, smem, ctx->stream>>>(P);
  }
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}

static int check_level(smf_ctx *ctx, const LevelDev &L)
{
  if (L.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "level mesh not set (smf_set_level)");
  if (!L.has_cell_data) return ctx->fail(SMF_ERR_INVALID, "cell data not set (smf_set_cell_data)");
  if (L.compressible) return ctx->fail(SMF_ERR_UNSUPPORTED, "is_compressible = true is not implemented on the device path yet");
  if (!L.safe_rounds) return ctx->fail(SMF_ERR_UNSUPPORTED, "cell batches with aliasing scatter rounds (unstructured orientation) not implemented");
  return SMF_OK;
}

int launch_vmult_A(smf_ctx *ctx, LevelDev &L, double *dst, const double *src)
{
  if (int e = check_level(ctx, L)) return e;
  CellParams P = make_params(L);
  P.src_u = src; P.dst_u = dst;
  SMF_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * L.n_u, ctx->stream));
  return L.dim == 3 ? launch_velocity<3, 0>(ctx, L, P) : launch_velocity<2, 0>(ctx, L, P);
}
int launch_vmult_stokes(smf_ctx *ctx, LevelDev &L, double *dst_u, double *dst_p, const double *src_u, const double *src_p)
{
  if (int e = check_level(ctx, L)) return e;
  CellParams P = make_params(L);
  P.src_u = src_u; P.src_p = src_p; P.dst_u = dst_u; P.dst_p = dst_p;
  SMF_CUDA(cudaMemsetAsync(dst_u, 0, sizeof(double) * L.n_u, ctx->stream));
  SMF_CUDA(cudaMemsetAsync(dst_p, 0, sizeof(double) * L.n_p, ctx->stream));
  return L.dim == 3 ? launch_velocity<3, 1>(ctx, L, P) : launch_velocity<2, 1>(ctx, L, P);
}
int launch_vmult_BT(smf_ctx *ctx, LevelDev &L, double *dst_u, const double *src_p)
{
  if (int e = check_level(ctx, L)) return e;
  CellParams P = make_params(L);
  P.src_p = src_p; P.dst_u = dst_u;
  SMF_CUDA(cudaMemsetAsync(dst_u, 0, sizeof(double) * L.n_u, ctx->stream));
  return L.dim == 3 ? launch_velocity<3, 2>(ctx, L, P) : launch_velocity<2, 2>(ctx, L, P);
}
int launch_vmult_mass(smf_ctx *ctx, LevelDev &L, double *dst, const double *src)
{
  if (int e = check_level(ctx, L)) return e;
  CellParams P = make_params(L);
  P.src_p = src; P.dst_p = dst;
  SMF_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * L.n_p, ctx->stream));
  const size_t smem = (size_t)(1 + KSUB) * L.max_nodes_p * sizeof(double);
  ProfScope prof(ctx, B_MASS);
  if (L.dim == 3) mass_kernel<3><<<L.n_batches, BATCH_CELLS, smem, ctx->stream>>>(P);
  else mass_kernel<2><<<L.n_batches, BATCH_CELLS, smem, ctx->stream>>>(P);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}

int launch_diag(smf_ctx *ctx, LevelDev &L)
{
  DiagParams P{};
  P.n_cells = L.n_cells; P.cell_nodes_u = L.cell_nodes_u; P.cell_nodes_p = L.cell_nodes_p;
  P.eta = L.eta; P.n_eta = L.n_eta;
  double vol = 1.0;
  for (int d = 0; d < L.dim; ++d) { P.ih[d] = 1.0 / L.h[d]; vol *= L.h[d]; }
  P.vol = vol; P.ps = L.ps; P.compressible = L.compressible;
  P.diag_A = L.dinv_A; P.diag_M = L.dinv_M;
  SMF_CUDA(cudaMemsetAsync(L.dinv_A, 0, sizeof(double) * L.n_u, ctx->stream));
  SMF_CUDA(cudaMemsetAsync(L.dinv_M, 0, sizeof(double) * L.n_p, ctx->stream));
  const int NN = L.dim == 3 ? 27 : 9;
  const int64_t nt = L.n_cells * NN;
  const int bs = 128;
  if (L.dim == 3) diag_kernel<3><<<(unsigned)((nt + bs - 1) / bs), bs, 0, ctx->stream>>>(P);
  else diag_kernel<2><<<(unsigned)((nt + bs - 1) / bs), bs, 0, ctx->stream>>>(P);
  ctx->launches++;
  // set_constrained_entries_to_one, then invert (matrix_free_operators.cc:1085-1097)
  if (L.n_constrained_u) {
    set_entries_kernel<<<(unsigned)((L.n_constrained_u + 255) / 256), 256, 0, ctx->stream>>>(L.dinv_A, L.constrained_u, L.n_constrained_u, 1.0);
    ctx->launches++;
  }
  if (L.n_constrained_p) {
    set_entries_kernel<<<(unsigned)((L.n_constrained_p + 255) / 256), 256, 0, ctx->stream>>>(L.dinv_M, L.constrained_p, L.n_constrained_p, 1.0);
    ctx->launches++;
  }
  invert_diag_kernel<<<(unsigned)((L.n_u + 255) / 256), 256, 0, ctx->stream>>>(L.dinv_A, L.n_u);
  invert_diag_kernel<<<(unsigned)((L.n_p + 255) / 256), 256, 0, ctx->stream>>>(L.dinv_M, L.n_p);
  ctx->launches += 2;
  SMF_CUDA(cudaGetLastError());
  L.has_diag = true;
  return SMF_OK;
}

template <int DIM, int NPD, int NCOMP>
static void launch_transfer_t(smf_ctx *ctx, const TransferParams &P, bool prolong)
{
  constexpr int NPP = 2 * NPD - 1;
  constexpr int NPATCH = DIM == 3 ? NPP * NPP * NPP : NPP * NPP;
  constexpr int NC = DIM == 3 ? NPD * NPD * NPD : NPD * NPD;
  const int bs = 128;
  ProfScope prof(ctx, B_TRANSFER);
  if (prolong) {
    const int64_t nt = P.n_parents * NPATCH;
    prolongate_kernel<DIM, NPD, NCOMP><<<(unsigned)((nt + bs - 1) / bs), bs, 0, ctx->stream>>>(P);
  } else {
    const int64_t nt = P.n_parents * NC;
    restrict_kernel<DIM, NPD, NCOMP><<<(unsigned)((nt + bs - 1) / bs), bs, 0, ctx->stream>>>(P);
  }
  ctx->launches++;
}

static int launch_transfer(smf_ctx *ctx, int which, int fine_level, double *dst, const double *src, bool prolong, bool add)
{
  if (fine_level < 1 || fine_level >= ctx->cfg.n_levels || !ctx->transfers[fine_level].set)
    return ctx->fail(SMF_ERR_INVALID, "transfer not set for this level (smf_set_transfer)");
  TransferDev &T = ctx->transfers[fine_level];
  LevelDev &C = ctx->levels[fine_level - 1];
  TransferParams P{};
  P.n_parents = T.n_parents; P.src = src; P.dst = dst; P.add = add ? 1 : 0;
  const int dim = ctx->cfg.dim;
  if (which == 0) {
    P.patch = T.patch_u; P.coarse_nodes = C.cell_nodes_u;
    if (dim == 3) launch_transfer_t<3, 3, 3>(ctx, P, prolong); else launch_transfer_t<2, 3, 2>(ctx, P, prolong);
  } else {
    P.patch = T.patch_p; P.coarse_nodes = C.cell_nodes_p;
    if (dim == 3) launch_transfer_t<3, 2, 1>(ctx, P, prolong); else launch_transfer_t<2, 2, 1>(ctx, P, prolong);
  }
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}

int launch_prolongate(smf_ctx *ctx, int which, int to_level, double *dst, const double *src, bool add)
{ return launch_transfer(ctx, which, to_level, dst, src, true, add); }
int launch_restrict_add(smf_ctx *ctx, int which, int from_level, double *dst, const double *src)
{ return launch_transfer(ctx, which, from_level, dst, src, false, true); }

}  // namespace smf
