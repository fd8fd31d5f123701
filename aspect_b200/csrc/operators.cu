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
#include <cuda.h>
#include <cuda/ptx>
#include <type_traits>
#include <algorithm>

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
  const uint32_t *node_list_u; const uint16_t *node_slot_u; const uint16_t *cell_loc_u; int list_stride_u;
  const uint32_t *node_list_p; const uint16_t *cell_loc_p;
  int max_nodes_u, max_nodes_p;
  const double *eta_t; int n_eta;     // lane-transposed table [batch][n_eta][32]
  double ee[3];                        // 1 / (2 h_d a): the line derivative below returns 2 D'
  double vol, ps;
  const double *src_u, *src_p;
  double *dst_u, *dst_p;
  int plain;                           // 1: read_dof_values_plain (constrained DoFs keep their values on read; their rows
                                       //    receive nothing): the operator of correct_stokes_rhs()
};

// Loads whose issue order the compiler must keep (asm volatile): the staging code issues every index load of a thread,
// then every value load, so a block exposes two memory latencies instead of two per staged item.
__device__ __forceinline__ uint32_t ldg_u32(const uint32_t *p)
{ uint32_t v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ldg_u16(const uint16_t *p)
{ uint16_t v; asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p)); return v; }
__device__ __forceinline__ double ldg_f64(const double *p)
{ double v; asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }

// ---- line primitives (3 values along one direction) ---------------------------------------------------------
// 2 D' (a, b, c): rows (-3, 4, -1), (-1, 0, 1), (1, -4, 3)
__device__ __forceinline__ void dline(double a, double b, double c, double &ga, double &gb, double &gc)
{
  const double d1 = b - a, d2 = c - b;
  gb = c - a;
  ga = fma(4.0, d1, -gb);
  gc = fma(4.0, d2, -gb);
}
// T += s * (2 D')^T x
__device__ __forceinline__ void dTline_add(double &Ta, double &Tb, double &Tc, double xa, double xb, double xc, double s, double s3, double s4)
{
  Ta = fma(-s3, xa, Ta); Ta = fma(-s, xb, Ta); Ta = fma(s, xc, Ta);
  Tb = fma(s4, xa - xc, Tb);
  Tc = fma(-s, xa, Tc); Tc = fma(s, xb, Tc); Tc = fma(s3, xc, Tc);
}

template <int DIM> __host__ __device__ constexpr int npair() { return DIM == 3 ? 3 : 1; }
template <int DIM> __host__ __device__ constexpr int pair_slot(int lo, int hi) { return DIM == 3 ? lo + hi - 1 : 0; }

template <int DIM> size_t vel_smem_bytes(int max_nodes_u, int max_nodes_p, int mode)
{
  size_t b = ((size_t)DIM * max_nodes_u + (size_t)npair<DIM>() * nn<DIM>() * SUB) * sizeof(double) + (size_t)nn<DIM>() * SUB * sizeof(uint16_t);
  if (mode != 0) b += ((size_t)max_nodes_p + (size_t)DIM * np<DIM>() * SUB) * sizeof(double) + (size_t)np<DIM>() * SUB * sizeof(uint16_t);
  return b;
}

// The per-direction pieces are templates on the direction so that every register index is a compile-time constant.
template <int DIM, int DIR, bool PERQ>
__device__ __forceinline__ void diag_dir(const double (&u)[nn<DIM>()], double (&T)[nn<DIM>()], double f2, const double *etab, bool add)
{
  // T (+)= (2D')^T [ f2 w_q eta_q (2D' u) ] along DIR, line by line
  for_lines<DIM, DIR>([&](int a, int b, int c) {
    double ga, gb, gc;
    dline(u[a], u[b], u[c], ga, gb, gc);
    if (PERQ) { ga *= f2 * qweight<DIM>(a) * etab[a * SUB]; gb *= f2 * qweight<DIM>(b) * etab[b * SUB]; gc *= f2 * qweight<DIM>(c) * etab[c * SUB]; }
    else { ga *= f2 * qweight<DIM>(a); gb *= f2 * qweight<DIM>(b); gc *= f2 * qweight<DIM>(c); }
    double ta = fma(-3.0, ga, gc - gb), tb = 4.0 * (ga - gc), tc = fma(3.0, gc, gb - ga);
    if (add) { T[a] += ta; T[b] += tb; T[c] += tc; }
    else { T[a] = ta; T[b] = tb; T[c] = tc; }
  });
}
// publish e * (2D'_DIR u) for the partner thread
template <int DIM, int DIR>
__device__ __forceinline__ void publish_dir(const double (&u)[nn<DIM>()], double *slot, double e)
{
  for_lines<DIM, DIR>([&](int a, int b, int c) {
    double ga, gb, gc;
    dline(u[a], u[b], u[c], ga, gb, gc);
    slot[a * SUB] = e * ga; slot[b * SUB] = e * gb; slot[c * SUB] = e * gc;
  });
}
// upper thread of a pair: X = coef_q (e_d 2D'_DIR u + partner), publish X, T += e_d (2D'_DIR)^T X
template <int DIM, int DIR, bool PERQ>
__device__ __forceinline__ void upper_dir(const double (&u)[nn<DIM>()], double (&T)[nn<DIM>()], double *slot, double e, double cvol, const double *etab)
{
  const double e3 = 3.0 * e, e4 = 4.0 * e;
  for_lines<DIM, DIR>([&](int a, int b, int c) {
    double ga, gb, gc;
    dline(u[a], u[b], u[c], ga, gb, gc);
    double xa = fma(e, ga, slot[a * SUB]), xb = fma(e, gb, slot[b * SUB]), xc = fma(e, gc, slot[c * SUB]);
    if (PERQ) { xa *= cvol * qweight<DIM>(a) * etab[a * SUB]; xb *= cvol * qweight<DIM>(b) * etab[b * SUB]; xc *= cvol * qweight<DIM>(c) * etab[c * SUB]; }
    else { xa *= cvol * qweight<DIM>(a); xb *= cvol * qweight<DIM>(b); xc *= cvol * qweight<DIM>(c); }
    slot[a * SUB] = xa; slot[b * SUB] = xb; slot[c * SUB] = xc;
    dTline_add(T[a], T[b], T[c], xa, xb, xc, e, e3, e4);
  });
}
// lower thread of a pair: T += e_d (2D'_DIR)^T X
template <int DIM, int DIR>
__device__ __forceinline__ void lower_dir(double (&T)[nn<DIM>()], const double *slot, double e)
{
  const double e3 = 3.0 * e, e4 = 4.0 * e;
  for_lines<DIM, DIR>([&](int a, int b, int c) {
    dTline_add(T[a], T[b], T[c], slot[a * SUB], slot[b * SUB], slot[c * SUB], e, e3, e4);
  });
}
#define SMF_DIR_SWITCH(dir, CALL0, CALL1, CALL2) \
  do { if ((dir) == 0) { CALL0; } else if ((dir) == 1) { CALL1; } else if (DIM == 3) { CALL2; } } while (0)

// MODE 0: A block (u -> u); MODE 1: Stokes ((u,p) -> (u,p)); MODE 2: B^T (p -> u).  PERQ: one viscosity per
// quadrature point instead of one per cell.  One block = one batch of 32 cells (a Morton 4x4x2 / 8x4 brick);
// warp c of the block owns velocity component c of those cells, one cell per lane.
// compile-time part of the structured-brick slot of local node n (smf_internal.h); the lane part is lb[variant]
__host__ __device__ constexpr int sbn_variant(int n) { return sb_variant((n % 3) & 1, ((n / 3) % 3) & 1); }
__host__ __device__ constexpr int sbn_const(int n)
{ return sb_class_base(sbn_variant(n), (n / 9) & 1) + sb_offset(sbn_variant(n), (n % 3) >> 1, ((n / 3) % 3) >> 1, (n / 9) >> 1); }

template <int DIM, int MODE, bool PERQ, bool STRUCT>
__global__ void __launch_bounds__(32 * DIM, DIM == 3 ? 4 : 8) velocity_kernel(const CellParams P)
{
  static_assert(!STRUCT || DIM == 3, "structured bricks are a 3D layout");
  constexpr int NN = nn<DIM>(), NP = np<DIM>(), NT = 32 * DIM, D2 = DIM == 3 ? 2 : 0;
  extern __shared__ double smem[];
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x;
  const int *info = P.batch_info + (size_t)b * BATCH_INFO;
  // (a structured brick always has 9 x 9 x 5 nodes and 32 cells: no dependent load in front of the staging loads)
  const int n_nodes_u = STRUCT ? 405 : info[0], n_excl_u = STRUCT ? 0 : info[1], n_cells_b = STRUCT ? SUB : info[4];
  const int n_nodes_p = MODE != 0 ? info[2] : 0, n_excl_p = MODE != 0 ? info[3] : 0;
  const int MU = STRUCT ? SB_SLOTS : P.max_nodes_u, MP = P.max_nodes_p;
  double *s_stage = smem;                              // [DIM][MU]: source values, later the accumulated result
  double *s_ex = s_stage + DIM * MU;                   // [npair][NN][32]
  double *s_stagep = s_ex + npair<DIM>() * NN * SUB;   // [MP]           (MODE != 0)
  double *s_pex = s_stagep + (MODE != 0 ? MP : 0);     // [DIM][NP][32]  (MODE != 0)
  uint16_t *s_loc = reinterpret_cast<uint16_t *>(s_pex + (MODE != 0 ? DIM * NP * SUB : 0));   // [NN][32]
  uint16_t *s_locp = s_loc + NN * SUB;                 // [NP][32]       (MODE != 0)

  const uint32_t *nlu = P.node_list_u + (size_t)b * P.list_stride_u;
  const uint16_t *nsl = P.node_slot_u + (size_t)b * P.list_stride_u;   // (structured bricks only)
  const uint32_t *nlp = P.node_list_p + (size_t)b * MP;
  // ---- stage the batch's unique DoFs (constrained DoFs read as zero: read_dof_values semantics) and its slot map ----
  {
    if (!STRUCT) {
      const uint32_t *gl = reinterpret_cast<const uint32_t *>(P.cell_loc_u + (size_t)b * NN * SUB);
      uint32_t *sl = reinterpret_cast<uint32_t *>(s_loc);
      for (int j = threadIdx.x; j < NN * SUB / 2; j += NT) sl[j] = gl[j];
    }
    if (MODE != 0) {
      const uint32_t *glp = reinterpret_cast<const uint32_t *>(P.cell_loc_p + (size_t)b * NP * SUB);
      uint32_t *slp = reinterpret_cast<uint32_t *>(s_locp);
      for (int j = threadIdx.x; j < NP * SUB / 2; j += NT) slp[j] = glp[j];
    }
  }
  // One lane per (node, component), nodes in ascending DoF order: consecutive lanes read consecutive doubles wherever
  // the batch's DoFs are contiguous in the vector.  All index loads of a thread first, then all value loads: two
  // exposed memory latencies per block.
  constexpr int SU = DIM == 3 ? 13 : 5;      // ceil(DIM * nodes of a structured brick / threads): 3 * 405 / 96, 2 * 153 / 64
  const int n_items = DIM * n_nodes_u;
  if (MODE != 2)
    for (int t0 = threadIdx.x; t0 < n_items; t0 += NT * SU) {
      uint32_t e[SU];
      int sl[SU];
      double v[SU];
#pragma unroll
      for (int i = 0; i < SU; ++i) {       // (items past the end repeat the last one: no predicated loads)
        const int j = min(t0 + i * NT, n_items - 1) / DIM;
        e[i] = ldg_u32(nlu + j);
        sl[i] = STRUCT ? (int)ldg_u16(nsl + j) : j;
      }
#pragma unroll
      for (int i = 0; i < SU; ++i) {
        const int t = min(t0 + i * NT, n_items - 1), d = t - (t / DIM) * DIM;
        const bool cons = (e[i] >> (NODE_FLAG_SHIFT + d)) & 1u;
        v[i] = ldg_f64(P.src_u + (e[i] & NODE_IDX_MASK) + d);      // constrained DoFs are valid addresses, too
        if (cons && !P.plain) v[i] = 0.0;
      }
#pragma unroll
      for (int i = 0; i < SU; ++i) {
        const int t = t0 + i * NT, d = t - (t / DIM) * DIM;
        if (t < n_items) s_stage[d * MU + sl[i]] = v[i];
      }
    }
  if (MODE != 0)
    for (int j = threadIdx.x; j < n_nodes_p; j += NT) {
      const uint32_t e = nlp[j];
      s_stagep[j] = ((e >> 31) && !P.plain) ? 0.0 : P.src_p[e & NODE_IDX_MASK];
    }
  const bool active = lane < n_cells_b;
  const double *etab = P.eta_t + (size_t)b * P.n_eta * SUB + lane;
  const double eta0 = PERQ ? 1.0 : etab[0];
  const double ec = P.ee[c];
  __syncthreads();

  // staging slot of local node n of this lane's cell
  int lb[4] = {0, 0, 0, 0};
  if (STRUCT) {
    const int x = sb_x(lane), y = sb_y(lane), z = sb_z(lane);
    #pragma unroll
    for (int v = 0; v < 4; ++v) lb[v] = x + sb_sy(v) * y + sb_sz(v) * z;
  }
  double *my_stage = s_stage + c * MU;
  const auto slot = [&](int n) -> int { return STRUCT ? sbn_const(n) + lb[sbn_variant(n)] : (int)s_loc[n * SUB + lane]; };

  double u[NN], T[NN];
  if (MODE != 2) {
#pragma unroll
    for (int n = 0; n < NN; ++n) u[n] = my_stage[slot(n)];
    interp_all<DIM>(u);
  }
  // ---- pressure contribution to the velocity rows: T = (2D'_c)^T [ -e_c ps vol w_q p_q ] ----
  if (MODE != 0) {
    double pn[NP], pq[NN];
#pragma unroll
    for (int i = 0; i < NP; ++i) pn[i] = s_stagep[s_locp[i * SUB + lane]];
    q1_to_quad<DIM>(pn, pq);
    const double f = -ec * P.ps * P.vol;
#pragma unroll
    for (int q = 0; q < NN; ++q) pq[q] *= f * qweight<DIM>(q);
    const auto pT = [&](auto dirc) {
      constexpr int DIR = decltype(dirc)::value;
      for_lines<DIM, DIR>([&](int a, int b2, int c2) {
        T[a] = fma(-3.0, pq[a], pq[c2] - pq[b2]); T[b2] = 4.0 * (pq[a] - pq[c2]); T[c2] = fma(3.0, pq[c2], pq[b2] - pq[a]);
      });
    };
    SMF_DIR_SWITCH(c, pT(std::integral_constant<int, 0>()), pT(std::integral_constant<int, 1>()), pT(std::integral_constant<int, D2>()));
  }
  if (MODE == 1) {   // pressure rows: -ps (div u, q): this component's share, summed below by warp 0
    double dq[NN], y[NP];
    const double f = -P.ps * P.vol * ec;
    const auto dv = [&](auto dirc) {
      constexpr int DIR = decltype(dirc)::value;
      for_lines<DIM, DIR>([&](int a, int b2, int c2) {
        dline(u[a], u[b2], u[c2], dq[a], dq[b2], dq[c2]);
        dq[a] *= f * qweight<DIM>(a); dq[b2] *= f * qweight<DIM>(b2); dq[c2] *= f * qweight<DIM>(c2);
      });
    };
    SMF_DIR_SWITCH(c, dv(std::integral_constant<int, 0>()), dv(std::integral_constant<int, 1>()), dv(std::integral_constant<int, D2>()));
    quad_to_q1T<DIM>(dq, y);
#pragma unroll
    for (int i = 0; i < NP; ++i) s_pex[(c * NP + i) * SUB + lane] = y[i];
  }
  if (MODE != 2) {
    // ---- diagonal direction: T (+)= (2D'_c)^T [ 2 eta w vol e_c^2 (2D'_c u) ] ----
    const double f2 = 2.0 * ec * ec * P.vol * eta0;
    SMF_DIR_SWITCH(c, (diag_dir<DIM, 0, PERQ>(u, T, f2, etab, MODE != 0)), (diag_dir<DIM, 1, PERQ>(u, T, f2, etab, MODE != 0)),
                   (diag_dir<DIM, D2, PERQ>(u, T, f2, etab, MODE != 0)));
    // ---- off-diagonal pairs {c,d}: X = eta w vol (e_d g_cd + e_c g_dc) is symmetric; the lower thread publishes
    //      its scaled derivative, the higher one forms X, uses it and publishes it back ----
    if (c == 0) {
      publish_dir<DIM, 1>(u, s_ex + (size_t)pair_slot<DIM>(0, 1) * NN * SUB + lane, P.ee[1]);
      if (DIM == 3) publish_dir<DIM, D2>(u, s_ex + (size_t)pair_slot<DIM>(0, 2) * NN * SUB + lane, P.ee[2]);
    } else if (DIM == 3 && c == 1) {
      publish_dir<DIM, D2>(u, s_ex + (size_t)pair_slot<DIM>(1, 2) * NN * SUB + lane, P.ee[2]);
    }
  }
  __syncthreads();       // every gather from the staging areas is done: they become the accumulation buffers
  for (int j = threadIdx.x; j < DIM * MU; j += NT) s_stage[j] = 0.0;
  if (MODE == 1)
    for (int j = threadIdx.x; j < MP; j += NT) s_stagep[j] = 0.0;
  if (MODE != 2) {
    const double cvol = P.vol * eta0;
    if (c == 1) {
      upper_dir<DIM, 0, PERQ>(u, T, s_ex + (size_t)pair_slot<DIM>(0, 1) * NN * SUB + lane, P.ee[0], cvol, etab);
    } else if (DIM == 3 && c == 2) {
      upper_dir<DIM, 0, PERQ>(u, T, s_ex + (size_t)pair_slot<DIM>(0, 2) * NN * SUB + lane, P.ee[0], cvol, etab);
      upper_dir<DIM, 1, PERQ>(u, T, s_ex + (size_t)pair_slot<DIM>(1, 2) * NN * SUB + lane, P.ee[1], cvol, etab);
    }
    __syncthreads();
    if (c == 0) {
      lower_dir<DIM, 1>(T, s_ex + (size_t)pair_slot<DIM>(0, 1) * NN * SUB + lane, P.ee[1]);
      if (DIM == 3) lower_dir<DIM, D2>(T, s_ex + (size_t)pair_slot<DIM>(0, 2) * NN * SUB + lane, P.ee[2]);
    } else if (DIM == 3 && c == 1) {
      lower_dir<DIM, D2>(T, s_ex + (size_t)pair_slot<DIM>(1, 2) * NN * SUB + lane, P.ee[2]);
    }
  } else {
    __syncthreads();     // (zeroing of s_stage finished)
  }
  if (MODE == 1 && c == DIM - 1) {   // pressure rows: sum the components' shares, scatter in collision-free rounds (one warp)
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      double y = s_pex[i * SUB + lane];
#pragma unroll
      for (int d = 1; d < DIM; ++d) y += s_pex[(d * NP + i) * SUB + lane];
      if (active) s_stagep[s_locp[i * SUB + lane]] += y;
      __syncwarp();
    }
  }
  interpT_all<DIM>(T);
  // ---- accumulate into the staging area, one collision-free round at a time ----
#pragma unroll
  for (int r = 0; r < (1 << DIM); ++r) {
#pragma unroll
    for (int n = 0; n < NN; ++n)
      if (round_q2<DIM>(n) == r && active) my_stage[slot(n)] += T[n];
    __syncwarp();
  }
  __syncthreads();
  // ---- flush: private nodes with plain stores, shared nodes with one RED per DoF; constrained DoFs get the
  //      unit diagonal of MatrixFreeOperators::Base (dst = src) ----
  for (int t0 = threadIdx.x; t0 < n_items; t0 += NT * SU) {
    uint32_t e[SU];
    int sl[SU];
#pragma unroll
    for (int i = 0; i < SU; ++i) {
      const int j = min(t0 + i * NT, n_items - 1) / DIM;
      e[i] = ldg_u32(nlu + j);
      sl[i] = STRUCT ? (int)ldg_u16(nsl + j) : j;
    }
#pragma unroll
    for (int i = 0; i < SU; ++i) {
      const int t = t0 + i * NT, j = t / DIM, d = t - j * DIM;
      if (t >= n_items) continue;
      const uint32_t idx = e[i] & NODE_IDX_MASK;
      const bool excl = STRUCT ? !(e[i] & NODE_SHARED_BIT) : (j < n_excl_u);
      if ((e[i] >> (NODE_FLAG_SHIFT + d)) & 1u) P.dst_u[idx + d] = (MODE == 2 || P.plain) ? 0.0 : P.src_u[idx + d];
      else if (excl) P.dst_u[idx + d] = s_stage[d * MU + sl[i]];
      else atomicAdd(P.dst_u + idx + d, s_stage[d * MU + sl[i]]);
    }
  }
  if (MODE == 1)
    for (int j = threadIdx.x; j < n_nodes_p; j += NT) {
      const uint32_t e = nlp[j];
      const uint32_t idx = e & NODE_IDX_MASK;
      if (e >> 31) P.dst_p[idx] = P.plain ? 0.0 : P.src_p[idx];
      else if (j < n_excl_p) P.dst_p[idx] = s_stagep[j];
      else atomicAdd(P.dst_p + idx, s_stagep[j]);
    }
}

// ---- structured bricks, persistent and software-pipelined (3D, A block and Stokes) -------------------------------
// Same arithmetic as velocity_kernel<3, MODE, PERQ, true>, organised so that no global-memory latency is exposed:
//  * the grid is one wave of resident blocks (4 per SM); block k walks bricks k, k + G, k + 2G, ...;
//  * while brick i is computed, cp.async copies the node list of brick i+1 (16-byte chunks) and then its DoF values
//    (8-byte copies straight into the staging slots; constrained DoFs arrive as zero through the src-size operand);
//  * two staging buffers: sA receives source values, sB accumulates results.  Accumulation needs no zero fill: the
//    rounds run in ascending order of the node's low-face pattern, so for every lattice point the first cell to touch
//    it is known per lane (first-touch mask) and stores instead of adding;
//  * the flush reads the node list from shared memory (no dependent global load in front of the stores / REDs);
//  * ROTATED FRAMES: warp c keeps its 27 values in a register frame whose direction r is the actual direction
//    (c + r) mod 3.  The operator is invariant under cyclic permutation of the axes, so in these frames all three
//    warps execute the SAME instructions (diagonal stress along register direction 0, pair {c, c+1} along 1, pair
//    {c-1, c} along 2); only the gathers / scatters between shared memory and registers differ per warp.  This cuts
//    the instruction footprint of the kernel by almost 3x (the one-brick-per-block kernel stalls on instruction
//    fetch).  A point with frame coordinates (q0, q1, q2) of warp c+1 is the point (q2, q0, q1) of warp c.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8_zfill(uint32_t dst, const void *src, int src_bytes)
{ asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{ asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// low-face pattern of (actual) local node n: bit d set when the node lies on the cell's low face in direction d
__host__ __device__ constexpr int lowface(int n) { return ((n % 3) == 0) + 2 * (((n / 3) % 3) == 0) + 4 * ((n / 9) == 0); }
// actual lexicographic index (base B digits, x fastest) of frame index m of component C
template <int C, int B> __host__ __device__ constexpr int rot_actual(int m)
{
  const int q0 = m % B, q1 = (m / B) % B, q2 = m / (B * B);
  return C == 0 ? q0 + B * q1 + B * B * q2 : (C == 1 ? q2 + B * q0 + B * B * q1 : q1 + B * q2 + B * B * q0);
}
// index in the frame of warp c-1 of the point with index m in the frame of warp c
__host__ __device__ constexpr int xperm(int m) { return (m / 9) + 3 * (m % 3) + 9 * ((m / 3) % 3); }

// one DoF of a brick's flush without branches: plain store (private to the brick), RED (shared with other bricks) or nothing
__device__ __forceinline__ void flush_dof(double *addr, double v, bool store, bool red)
{
  asm volatile("{\n\t.reg .pred ps, pr;\n\tsetp.ne.u32 ps, %2, 0;\n\tsetp.ne.u32 pr, %3, 0;\n\t"
               "@ps st.global.f64 [%0], %1;\n\t@pr red.global.add.f64 [%0], %1;\n\t}"
               ::"l"(addr), "d"(v), "r"((unsigned)store), "r"((unsigned)red) : "memory");
}

constexpr int SBK_THREADS = 96;
constexpr int SBK_SU = 13;            // ceil(3 * 405 / 96) staged items per thread
constexpr size_t sbk_smem_bytes(int mode)
{
  return (size_t)(2 * 3 * SB_SLOTS + 3 * 27 * SUB) * sizeof(double) + 2 * (size_t)(mode == 0 ? SB_LIST_BYTES : SB_BLOCK_BYTES)
         + (mode == 0 ? 0 : SB_PLIST * sizeof(double));
}

// viscosity at frame point m: es[r] = SUB * 3^(actual direction of register direction r)
struct EtaRot {
  const double *etab; int es0, es1, es2;
  __device__ __forceinline__ double operator()(int m) const { return etab[(m % 3) * es0 + ((m / 3) % 3) * es1 + (m / 9) * es2]; }
};
// T (+)= (2D'_0)^T [ f2 w_q eta_q (2D'_0 u) ]
template <bool PERQ> __device__ __forceinline__ void sb_diag(const double (&u)[27], double (&T)[27], double f2, const EtaRot &eta, bool add)
{
  for_lines<3, 0>([&](int a, int b, int c) {
    double ga, gb, gc;
    dline(u[a], u[b], u[c], ga, gb, gc);
    if (PERQ) { ga *= f2 * qweight<3>(a) * eta(a); gb *= f2 * qweight<3>(b) * eta(b); gc *= f2 * qweight<3>(c) * eta(c); }
    else { ga *= f2 * qweight<3>(a); gb *= f2 * qweight<3>(b); gc *= f2 * qweight<3>(c); }
    const double ta = fma(-3.0, ga, gc - gb), tb = 4.0 * (ga - gc), tc = fma(3.0, gc, gb - ga);
    if (add) { T[a] += ta; T[b] += tb; T[c] += tc; }
    else { T[a] = ta; T[b] = tb; T[c] = tc; }
  });
}
// pair {c, c+1}, lower thread: publish e1 (2D'_1 u) in its own frame
__device__ __forceinline__ void sb_publish(const double (&u)[27], double *buf, double e1)
{
  for_lines<3, 1>([&](int a, int b, int c) {
    double ga, gb, gc;
    dline(u[a], u[b], u[c], ga, gb, gc);
    buf[a * SUB] = e1 * ga; buf[b * SUB] = e1 * gb; buf[c * SUB] = e1 * gc;
  });
}
// pair {c-1, c}, upper thread: X = coef_q (e2 2D'_2 u + partner), published in the partner's frame; T += e2 (2D'_2)^T X
template <bool PERQ> __device__ __forceinline__ void sb_upper(const double (&u)[27], double (&T)[27], double *buf, double e2, double cvol, const EtaRot &eta)
{
  const double e3 = 3.0 * e2, e4 = 4.0 * e2;
  for_lines<3, 2>([&](int a, int b, int c) {
    double ga, gb, gc;
    dline(u[a], u[b], u[c], ga, gb, gc);
    double xa = fma(e2, ga, buf[xperm(a) * SUB]), xb = fma(e2, gb, buf[xperm(b) * SUB]), xc = fma(e2, gc, buf[xperm(c) * SUB]);
    if (PERQ) { xa *= cvol * qweight<3>(a) * eta(a); xb *= cvol * qweight<3>(b) * eta(b); xc *= cvol * qweight<3>(c) * eta(c); }
    else { xa *= cvol * qweight<3>(a); xb *= cvol * qweight<3>(b); xc *= cvol * qweight<3>(c); }
    buf[xperm(a) * SUB] = xa; buf[xperm(b) * SUB] = xb; buf[xperm(c) * SUB] = xc;
    dTline_add(T[a], T[b], T[c], xa, xb, xc, e2, e3, e4);
  });
}
// pair {c, c+1}, lower thread: T += e1 (2D'_1)^T X
__device__ __forceinline__ void sb_lower(double (&T)[27], const double *buf, double e1)
{
  const double e3 = 3.0 * e1, e4 = 4.0 * e1;
  for_lines<3, 1>([&](int a, int b, int c) { dTline_add(T[a], T[b], T[c], buf[a * SUB], buf[b * SUB], buf[c * SUB], e1, e3, e4); });
}
// registers <-> staging slots of this lane's cell, per component frame
template <int C> __device__ __forceinline__ void sb_gather(double (&u)[27], const double *stage, const int (&lb)[4])
{
#pragma unroll
  for (int m = 0; m < 27; ++m) { const int n = rot_actual<C, 3>(m); u[m] = stage[sbn_const(n) + lb[sbn_variant(n)]]; }
}
template <int C> __device__ __forceinline__ void sb_accumulate(double (&T)[27], double *stage, const int (&lb)[4], unsigned first_mask)
{
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const bool first = (first_mask >> r) & 1u;
    if (r != 0) {       // (all loads of a round before its stores: the slots of a round are distinct)
#pragma unroll
      for (int m = 0; m < 27; ++m) {
        const int n = rot_actual<C, 3>(m);
        if (lowface(n) == r) { const double old = first ? 0.0 : stage[sbn_const(n) + lb[sbn_variant(n)]]; T[m] += old; }
      }
    }
#pragma unroll
    for (int m = 0; m < 27; ++m) {
      const int n = rot_actual<C, 3>(m);
      if (lowface(n) == r) stage[sbn_const(n) + lb[sbn_variant(n)]] = T[m];
    }
    __syncwarp();
  }
}
template <int C> __device__ __forceinline__ void sb_gather_p(double (&pn)[8], const double *sPA, int x, int y, int z)
{
#pragma unroll
  for (int m = 0; m < 8; ++m) { const int n = rot_actual<C, 2>(m); pn[m] = sPA[(x + (n & 1)) + 5 * (y + ((n >> 1) & 1)) + 25 * (z + (n >> 2))]; }
}
template <int C> __device__ __forceinline__ void sb_park_p(const double (&yv)[8], double *park)
{
#pragma unroll
  for (int m = 0; m < 8; ++m) park[rot_actual<C, 2>(m) * SUB] = yv[m];
}

template <int MODE, bool PERQ>
__global__ void __launch_bounds__(SBK_THREADS, 4) velocity_sb_kernel(const CellParams P, const unsigned char *lists, int n_batches)
{
  static_assert(MODE >= 0 && MODE <= 2, "A block, Stokes operator, B^T");
  constexpr int DIM = 3, NN = 27, NP = 8, MU = SB_SLOTS;
  constexpr int LBYTES = MODE == 0 ? SB_LIST_BYTES : SB_BLOCK_BYTES;     // bytes of a brick's lists that this mode needs
  extern __shared__ __align__(16) double smem[];
  double *sA = smem;                       // [3][MU] source values of the current brick (cp.async target)
  double *sB = sA + 3 * MU;                // [3][MU] accumulated results of the current brick
  double *s_ex = sB + 3 * MU;              // [3][NN][32] pair buffers, indexed by the lower component of the pair
  unsigned char *sI = reinterpret_cast<unsigned char *>(s_ex + 3 * NN * SUB);   // [2][LBYTES] node lists
  double *sPA = reinterpret_cast<double *>(sI + 2 * LBYTES);                    // [SB_PLIST] pressure values (MODE 1)
  const int tid = threadIdx.x, c = tid >> 5, lane = tid & 31, G = gridDim.x;
  int b = blockIdx.x;
  if (b >= n_batches) return;

  // lane constants: slot offsets of this lane's cell and its first-touch mask
  const int x = sb_x(lane), y = sb_y(lane), z = sb_z(lane);
  int lb[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) lb[v] = x + sb_sy(v) * y + sb_sz(v) * z;
  unsigned first_mask = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if ((!(r & 1) || x == 0) && (!(r & 2) || y == 0) && (!(r & 4) || z == 0)) first_mask |= 1u << r;
  const int j0 = tid / 3, d0 = tid - 3 * j0;      // staged item k of this thread: node j0 + 32 k, component d0
  const int c1 = c == 2 ? 0 : c + 1, c2 = c == 0 ? 2 : c - 1;      // actual directions of register directions 1, 2
  const double e0 = P.ee[c], e1 = P.ee[c1], e2 = P.ee[c2];
  double *ex_lo = s_ex + (size_t)c * NN * SUB + lane;      // pair {c, c+1}: this thread is the lower one
  double *ex_up = s_ex + (size_t)c2 * NN * SUB + lane;     // pair {c-1, c}: this thread is the upper one
  double *myA = sA + c * MU, *myB = sB + c * MU;
  EtaRot eta;
  eta.es0 = SUB * stride3(c); eta.es1 = SUB * stride3(c1); eta.es2 = SUB * stride3(c2);

  const auto load_list = [&](unsigned char *dst, int brick) {
    const unsigned char *src = lists + (size_t)brick * SB_BLOCK_BYTES;
    for (int ch = tid; ch < LBYTES / 16; ch += SBK_THREADS) cp_async16(smem_u32(dst + 16 * ch), src + 16 * ch);
  };
  const auto issue_values = [&](const unsigned char *I) {
    const uint32_t *le = reinterpret_cast<const uint32_t *>(I);
    const uint16_t *ls = reinterpret_cast<const uint16_t *>(I + SB_LIST * 4);
    const uint32_t base = smem_u32(sA + d0 * MU);
    const double *srcd = P.src_u + d0;
    // (sub-batches: the list reads of a batch are in flight before its first copy is issued; u and T are live here)
    constexpr int SUBB = 4;
    if (MODE != 2) {
      // 405 nodes = 12 full rounds of 32 (no bounds checks) + 21 nodes in a 13th round
#pragma unroll 1
      for (int k0 = 0; k0 < 12; k0 += SUBB) {
        uint32_t e[SUBB], sl[SUBB];
#pragma unroll
        for (int k = 0; k < SUBB; ++k) {
          const int j = j0 + 32 * (k0 + k);
          e[k] = le[j]; sl[k] = ls[j];
        }
#pragma unroll
        for (int k = 0; k < SUBB; ++k) {
          const bool cons = ((e[k] >> (NODE_FLAG_SHIFT + d0)) & 1u) && !P.plain;
          cp_async8_zfill(base + 8u * sl[k], srcd + (e[k] & NODE_IDX_MASK), cons ? 0 : 8);
        }
      }
      if (j0 < 405 - 384) {
        const uint32_t e = le[j0 + 384], sl = ls[j0 + 384];
        const bool cons = ((e >> (NODE_FLAG_SHIFT + d0)) & 1u) && !P.plain;
        cp_async8_zfill(base + 8u * sl, srcd + (e & NODE_IDX_MASK), cons ? 0 : 8);
      }
    }
    if (MODE != 0 && tid < 75) {      // pressure values in lattice order
      const uint32_t e = reinterpret_cast<const uint32_t *>(I + SB_LIST_BYTES)[tid];
      cp_async8_zfill(smem_u32(sPA + tid), P.src_p + (e & NODE_IDX_MASK), ((e >> 31) && !P.plain) ? 0 : 8);
    }
  };

  // prologue: node list and values of the first brick
  load_list(sI, b);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  issue_values(sI);
  cp_async_commit();
  double eta_next = (PERQ || MODE == 2) ? 1.0 : P.eta_t[(size_t)b * P.n_eta * SUB + lane];

  for (int it = 0; b < n_batches; b += G, ++it) {
    unsigned char *I_cur = sI + (it & 1) * LBYTES, *I_nxt = sI + ((it + 1) & 1) * LBYTES;
    const int bn = b + G;
    cp_async_wait_all();
    __syncthreads();                        // values of this brick are in sA; the previous flush is done with sB / I_nxt
    if (bn < n_batches) load_list(I_nxt, bn);
    cp_async_commit();
    const double eta0 = eta_next;
    eta.etab = P.eta_t + (size_t)b * P.n_eta * SUB + lane;

    double u[NN], T[NN];
    if (MODE != 2) {
      if (c == 0) sb_gather<0>(u, myA, lb); else if (c == 1) sb_gather<1>(u, myA, lb); else sb_gather<2>(u, myA, lb);
      interp_all<DIM>(u);
    }
    if (MODE != 0) {
      // pressure contribution to the velocity rows: T = (2D'_0)^T [ -e_c ps vol w_q p_q ]
      {
        double pn[NP], pq[NN];
        if (c == 0) sb_gather_p<0>(pn, sPA, x, y, z); else if (c == 1) sb_gather_p<1>(pn, sPA, x, y, z); else sb_gather_p<2>(pn, sPA, x, y, z);
        q1_to_quad<DIM>(pn, pq);
        const double f = -e0 * P.ps * P.vol;
#pragma unroll
        for (int q = 0; q < NN; ++q) pq[q] *= f * qweight<DIM>(q);
        for_lines<DIM, 0>([&](int a, int b2, int c2_) {
          T[a] = fma(-3.0, pq[a], pq[c2_] - pq[b2]); T[b2] = 4.0 * (pq[a] - pq[c2_]); T[c2_] = fma(3.0, pq[c2_], pq[b2] - pq[a]);
        });
      }
      // pressure rows -ps (div u, q): this component's share, parked in sB (free until the accumulation) for the
      // lattice-wise sum after the next barrier
      if (MODE == 1) {
        double dq[NN], yv[NP];
        const double f = -P.ps * P.vol * e0;
        for_lines<DIM, 0>([&](int a, int b2, int c2_) {
          dline(u[a], u[b2], u[c2_], dq[a], dq[b2], dq[c2_]);
          dq[a] *= f * qweight<DIM>(a); dq[b2] *= f * qweight<DIM>(b2); dq[c2_] *= f * qweight<DIM>(c2_);
        });
        quad_to_q1T<DIM>(dq, yv);
        const int pc = x + 4 * y + 16 * z;      // lexicographic cell number within the brick
        if (c == 0) sb_park_p<0>(yv, myB + pc); else if (c == 1) sb_park_p<1>(yv, myB + pc); else sb_park_p<2>(yv, myB + pc);
      }
    }
    if (MODE != 2) {
      // diagonal direction: T (+)= (2D'_0)^T [ 2 eta w vol e_c^2 (2D'_0 u) ]
      sb_diag<PERQ>(u, T, 2.0 * e0 * e0 * P.vol * eta0, eta, MODE == 1);
      // pair {c, c+1}: publish e_{c+1} (2D'_1 u)
      sb_publish(u, ex_lo, e1);
    }
    cp_async_wait_all();                    // (this thread's chunks of the next node list)
    __syncthreads();                        // publishes visible; every gather from sA is done; next list visible
    if (bn < n_batches) {
      issue_values(I_nxt);
      if (!PERQ && MODE != 2) eta_next = P.eta_t[(size_t)bn * P.n_eta * SUB + lane];
    }
    cp_async_commit();
    if (MODE == 1 && tid < 75) {
      // pressure rows: lattice node tid sums the shares of the (up to 8) cells around it, all components, in a fixed
      // order, and goes straight to global memory (private node: store, shared: RED, constrained: dst = src).  The
      // shares are parked by LEXICOGRAPHIC cell number (x + 4 y + 16 z, not by lane), so the eight reads are one base
      // address plus compile-time offsets: with lane-ordered parking the kernel kept four loop-invariant addresses,
      // spilled them, and the reload in front of the first read held 7 % of the stall samples.
      const int X = tid % 5, Y = (tid / 5) % 5, Z = tid / 25;
      const double *pk = sB + (X + 4 * Y + 16 * Z);
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int cx = X - i, cy = Y - j, cz = Z - k;
            if (cx >= 0 && cx < 4 && cy >= 0 && cy < 4 && cz >= 0 && cz < 2) {
                      const int off = (i + 2 * j + 4 * k) * SUB - (i + 4 * j + 16 * k);
              acc += pk[off] + pk[MU + off] + pk[2 * MU + off];
            }
          }
      const uint32_t e = reinterpret_cast<const uint32_t *>(I_cur + SB_LIST_BYTES)[tid];
      const uint32_t idx = e & NODE_IDX_MASK;
      if (e >> 31) P.dst_p[idx] = P.plain ? 0.0 : P.src_p[idx];
      else if (!(e & NODE_SHARED_BIT)) P.dst_p[idx] = acc;
      else atomicAdd(P.dst_p + idx, acc);
    }
    // pair {c-1, c}: X = eta w vol (e_{c-1} 2D'_2 u + partner), T += e_{c-1} (2D'_2)^T X, X published
    if (MODE != 2) {
      sb_upper<PERQ>(u, T, ex_up, e2, P.vol * eta0, eta);
      __syncthreads();
      sb_lower(T, ex_lo, e1);
    }
    interpT_all<DIM>(T);
    // accumulate into sB in ascending order of the low-face pattern; the first cell to touch a lattice point stores
    if (c == 0) sb_accumulate<0>(T, myB, lb, first_mask); else if (c == 1) sb_accumulate<1>(T, myB, lb, first_mask); else sb_accumulate<2>(T, myB, lb, first_mask);
    __syncthreads();
    // flush: private nodes with plain stores, shared nodes with one RED per DoF; constrained DoFs are skipped (dst is
    // zero there on entry; dst = src of MatrixFreeOperators::Base is one small kernel over the constrained list after
    // this one -- inside this kernel it was a dependent global load per boundary brick, 7 % of the stall samples)
    {
      const uint32_t *le = reinterpret_cast<const uint32_t *>(I_cur);
      const uint16_t *ls = reinterpret_cast<const uint16_t *>(I_cur + SB_LIST * 4);
      const double *sBd = sB + d0 * MU;
      double *dstd = P.dst_u + d0;
      const int n_private = (int)le[405];
      // u and T are dead here: all list reads of the thread first, then all value reads, then straight-line predicated
      // stores / REDs (flush_dof) -- the branchy version (store / RED / skip per item, three sub-batches) held 18 % of the
      // kernel's stall samples on its BSSY / BRA / BSYNC sequences
      constexpr int NF = 12;
      uint32_t e[NF], sl[NF];
      double v[NF];
#pragma unroll
      for (int k = 0; k < NF; ++k) { const int j = j0 + 32 * k; e[k] = le[j]; sl[k] = ls[j]; }
#pragma unroll
      for (int k = 0; k < NF; ++k) v[k] = sBd[sl[k]];
#pragma unroll
      for (int k = 0; k < NF; ++k) {
        const int j = j0 + 32 * k;
        const bool free_dof = !((e[k] >> (NODE_FLAG_SHIFT + d0)) & 1u);
        flush_dof(dstd + (e[k] & NODE_IDX_MASK), v[k], free_dof && j < n_private, free_dof && j >= n_private);
      }
      if (j0 < 405 - 384) {
        const int j = j0 + 384;
        const uint32_t e1 = le[j];
        const double v1 = sBd[ls[j]];
        const bool free_dof = !((e1 >> (NODE_FLAG_SHIFT + d0)) & 1u);
        flush_dof(dstd + (e1 & NODE_IDX_MASK), v1, free_dof && j < n_private, free_dof && j >= n_private);
      }
    }
  }
}

// ---- structured bricks on LATTICE-ORDERED vectors: the brick is a box of the tensor [component][Z][Y][X] and arrives by TMA ----
// (SMF_OPT_LATTICE_VELOCITY; set-up and rationale: velocity_lattice.cu.)  Same thread mapping and arithmetic as
// velocity_sb_kernel<0>: one block per brick, warp c owns component c, lane = cell.  What changes is how the 9 x 9 x 5 lattice gets
// into shared memory and back:
//   * the vectors are stored as [component][x parity][Z][Y][X / 2] (even and odd lattice columns apart: TMA has no element stride
//     in the innermost dimension), and thread 0 issues six cp.async.bulk.tensor loads per brick (3 components x 2 parities: boxes
//     of 6 x 10 x 5 doubles), completion on an mbarrier; no index lists, no per-thread copies;
//   * in shared memory a row is 6 doubles (5 even-x or 4 odd-x points + padding) and a plane 10 rows (9 + 1): the slot of local
//     node (i, j, k) of the cell in lane (cx, cy, cz) is const(i, j, k) + cx + 12 cy + 120 cz, and the 16 lanes of a half-warp
//     (x0 y0 z0 x1 of the Morton lane order) fall into 16 different 8-byte banks: cx + 12 y0 + 8 z0 (mod 16);
//   * the flush computes its global addresses from the brick's lattice origin: rows of 9 consecutive doubles, plain stores for
//     the 7 x 7 x 3 interior, one RED per DoF on the brick's faces, nothing on constrained boundary faces.
constexpr int VL_ROW = 6, VL_ROWS = 10, VL_PLANE = VL_ROW * VL_ROWS;       // doubles per staged row / rows per plane / plane pitch
constexpr int VL_BOX = VL_PLANE * 5;                                        // one (component, x-parity) box: 300 doubles = 2,400 bytes
constexpr int VL_BOX_PITCH = 304;                                           // boxes start 128-byte aligned
constexpr int VL_COMP = 2 * VL_BOX_PITCH, VL_STAGE = 3 * VL_COMP;           // per component / per buffer (1,824 doubles)
__host__ __device__ constexpr int vl_const(int n)
{ return ((n / 9) * VL_ROWS + ((n / 3) % 3)) * VL_ROW + ((n % 3) >> 1) + (((n % 3) & 1) ? VL_BOX_PITCH : 0); }
constexpr size_t vl_smem_bytes() { return (size_t)(2 * VL_STAGE + 3 * 27 * SUB) * sizeof(double) + 128; }

template <int C> __device__ __forceinline__ void vl_gather(double (&u)[27], const double *stage)
{
#pragma unroll
  for (int m = 0; m < 27; ++m) u[m] = stage[vl_const(rot_actual<C, 3>(m))];
}
template <int C> __device__ __forceinline__ void vl_accumulate(double (&T)[27], double *stage, unsigned first_mask)
{
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const bool first = (first_mask >> r) & 1u;
    if (r != 0) {
#pragma unroll
      for (int m = 0; m < 27; ++m) {
        const int n = rot_actual<C, 3>(m);
        if (lowface(n) == r) { const double old = first ? 0.0 : stage[vl_const(n)]; T[m] += old; }
      }
    }
#pragma unroll
    for (int m = 0; m < 27; ++m) {
      const int n = rot_actual<C, 3>(m);
      if (lowface(n) == r) stage[vl_const(n)] = T[m];
    }
    __syncwarp();
  }
}

struct VLParams {
  const int *brick_org;      // [n_batches][4]
  double *dst;               // lattice-ordered result (zero on entry)
  int Xp, ny, nz, cons_mask, plain;
};

template <bool PERQ>
__global__ void __launch_bounds__(SBK_THREADS, 4) velocity_lat_kernel(const __grid_constant__ CUtensorMap tmap, const CellParams P, const VLParams V, int n_batches)
{
  constexpr int NN = 27;
  extern __shared__ __align__(128) unsigned char vl_raw[];
  // (the dynamic shared-memory window is 16-byte aligned by the ABI; the TMA destination wants 128)
  double *sA = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(vl_raw) + 127) & ~uintptr_t(127));      // [3][2][VL_BOX_PITCH] values of the current brick
  double *sB = sA + VL_STAGE;                 // accumulated results
  double *s_ex = sB + VL_STAGE;               // [3][NN][32] pair buffers
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, c = tid >> 5, lane = tid & 31, G = gridDim.x;
  int b = blockIdx.x;
  if (b >= n_batches) return;
  const int x = sb_x(lane), y = sb_y(lane), z = sb_z(lane);
  const int lb = x + 2 * VL_ROW * y + 2 * VL_PLANE * z;
  unsigned first_mask = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if ((!(r & 1) || x == 0) && (!(r & 2) || y == 0) && (!(r & 4) || z == 0)) first_mask |= 1u << r;
  const int c1 = c == 2 ? 0 : c + 1, c2 = c == 0 ? 2 : c - 1;
  const double e0 = P.ee[c], e1 = P.ee[c1], e2 = P.ee[c2];
  double *ex_lo = s_ex + (size_t)c * NN * SUB + lane;
  double *ex_up = s_ex + (size_t)c2 * NN * SUB + lane;
  double *myA = sA + c * VL_COMP + lb, *myB = sB + c * VL_COMP + lb;
  EtaRot eta;
  eta.es0 = SUB * stride3(c); eta.es1 = SUB * stride3(c1); eta.es2 = SUB * stride3(c2);

  const int frow = tid / 9, fx = tid - 9 * frow;      // flush: first row and column of this thread
  constexpr uint32_t STAGE_BYTES = 6 * VL_BOX * sizeof(double);
  const auto issue = [&](int brick) {      // (thread 0) arm the barrier, start the six box loads of `brick`
    const int4 o = *reinterpret_cast<const int4 *>(V.brick_org + 4 * (size_t)brick);
    cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared, &bar, STAGE_BYTES);
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int32_t crd[4] = {o.x >> 1, o.y, o.z, 2 * d + par};
        cuda::ptx::cp_async_bulk_tensor(cuda::ptx::space_cluster, cuda::ptx::space_global, sA + d * VL_COMP + par * VL_BOX_PITCH, &tmap, crd, &bar);
      }
  };
  if (tid == 0) {
    cuda::ptx::mbarrier_init(&bar, 1);
    cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);
  }
  __syncthreads();
  if (tid == 0) issue(b);
  double eta_next = PERQ ? 1.0 : P.eta_t[(size_t)b * P.n_eta * SUB + lane];
  uint32_t parity = 0;

  for (; b < n_batches; b += G) {
    const int bn = b + G;
    const int4 org = *reinterpret_cast<const int4 *>(V.brick_org + 4 * (size_t)b);
    while (!cuda::ptx::mbarrier_try_wait_parity(&bar, parity)) {}
    parity ^= 1;
    const double eta0 = eta_next;
    eta.etab = P.eta_t + (size_t)b * P.n_eta * SUB + lane;
    // read_dof_values: constrained DoFs enter as zero -- only bricks on a constrained boundary face have any
    if (org.w && V.cons_mask && !V.plain) {
      for (int f = 0; f < 6; ++f) {
        const int d = f >> 1, side = f & 1;
        if (!((org.w >> (side ? 3 + d : d)) & 1)) continue;
        const int cm = (V.cons_mask >> (3 * f)) & 7;
        if (!cm) continue;
        const int n1 = d == 0 ? 9 : 9, n2 = d == 2 ? 9 : 5;      // extents of the two other directions: (Y, Z) / (X, Z) / (X, Y)
        for (int t = tid; t < 3 * n1 * n2; t += SBK_THREADS) {
          const int comp = t / (n1 * n2), r = t - comp * n1 * n2, a = r % n1, bb = r / n1;
          if (!((cm >> comp) & 1)) continue;
          const int X = d == 0 ? (side ? 8 : 0) : a, Y = d == 1 ? (side ? 8 : 0) : (d == 0 ? a : bb), Z = d == 2 ? (side ? 4 : 0) : bb;
          sA[comp * VL_COMP + ((X & 1) ? VL_BOX_PITCH : 0) + (Z * VL_ROWS + Y) * VL_ROW + (X >> 1)] = 0.0;
        }
      }
      __syncthreads();
    }
    double u[NN], T[NN];
    if (c == 0) vl_gather<0>(u, myA); else if (c == 1) vl_gather<1>(u, myA); else vl_gather<2>(u, myA);
    interp_all<3>(u);
    sb_diag<PERQ>(u, T, 2.0 * e0 * e0 * P.vol * eta0, eta, false);
    sb_publish(u, ex_lo, e1);
    __syncthreads();                        // publishes visible; every gather from sA is done
    if (bn < n_batches) {
      if (tid == 0) issue(bn);
      if (!PERQ) eta_next = P.eta_t[(size_t)bn * P.n_eta * SUB + lane];
    }
    sb_upper<PERQ>(u, T, ex_up, e2, P.vol * eta0, eta);
    __syncthreads();
    sb_lower(T, ex_lo, e1);
    interpT_all<3>(T);
    if (c == 0) vl_accumulate<0>(T, myB, first_mask); else if (c == 1) vl_accumulate<1>(T, myB, first_mask); else vl_accumulate<2>(T, myB, first_mask);
    __syncthreads();
    // flush: thread (tid / 9, tid % 9) owns column X = tid % 9 of every 10th row (component, Z, Y) of the 135 rows; the row
    // counters advance by additions only (no division in the loop: the first version decoded a flat item index and spent
    // more instructions here than on the arithmetic of the cell)
    if (tid < 90) {
      const int X = fx, par = X & 1;
      int cons_x = 0;
      if (org.w) {
        if (X == 0 && (org.w & 1)) cons_x |= V.cons_mask;            // faces 0..5: (d, side) = (0,0) (0,1) (1,0) (1,1) (2,0) (2,1)
        if (X == 8 && (org.w & 8)) cons_x |= V.cons_mask >> 3;
      }
      const bool priv_x = X > 0 && X < 8;
      // running 32-bit offsets (a level has fewer than 2^28 DoFs per family): additions only
      const int sY = V.Xp, sZ = V.ny * V.Xp, cpl = V.nz * sZ;      // row / plane / (component, parity) array pitch of the result
      int goff = par * cpl + (org.z + (frow >= 9 ? 1 : 0)) * sZ + (org.y + (frow >= 9 ? frow - 9 : frow)) * sY + ((org.x + X) >> 1);
      int soff = par * VL_BOX_PITCH + (X >> 1) + (frow >= 9 ? VL_PLANE + (frow - 9) * VL_ROW : frow * VL_ROW);
      int comp = 0, Z = frow >= 9 ? 1 : 0, Y = frow >= 9 ? frow - 9 : frow;      // row = frow, frow + 10, ... = comp * 45 + Z * 9 + Y
#pragma unroll 1
      for (int row = frow; row < 135; row += 10) {
        const double v = sB[soff];
        int cons = cons_x;
        if (org.w) {
          if (Y == 0 && (org.w & 2)) cons |= V.cons_mask >> 6;
          if (Y == 8 && (org.w & 16)) cons |= V.cons_mask >> 9;
          if (Z == 0 && (org.w & 4)) cons |= V.cons_mask >> 12;
          if (Z == 4 && (org.w & 32)) cons |= V.cons_mask >> 15;
        }
        if (!((cons >> comp) & 1)) {      // constrained rows are skipped (dst = src there is the list kernel after this one)
          if (priv_x && (unsigned)(Y - 1) < 7u && (unsigned)(Z - 1) < 3u) V.dst[goff] = v; else atomicAdd(V.dst + goff, v);
        }
        Y += 1; Z += 1; goff += sY + sZ; soff += VL_ROW + VL_PLANE;      // + 10 rows = + 1 plane + 1 row
        if (Y >= 9) { Y -= 9; Z += 1; goff += sZ - 9 * sY; soff += VL_PLANE - 9 * VL_ROW; }
        if (Z >= 5) { Z -= 5; comp += 1; goff += 2 * cpl - 5 * sZ; soff += VL_COMP - 5 * VL_PLANE; }
      }
    }
  }
}

// host numbering <-> lattice order (thread per lattice node; the three components of a node are consecutive DoFs)
__global__ void vl_to_lattice(double *lat, const double *vec, const uint32_t *lat2dof, int nx, int ny, int nz, int Xp)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, n = (int64_t)nx * ny * nz;
  if (i >= n) return;
  const int X = (int)(i % nx);
  const int64_t row = i / nx, cpl = (int64_t)nz * ny * Xp;
  const uint32_t d = lat2dof[i];
#pragma unroll
  for (int c = 0; c < 3; ++c) lat[(2 * c + (X & 1)) * cpl + row * Xp + (X >> 1)] = vec[d + c];
}
__global__ void vl_from_lattice(double *vec, const double *lat, const uint32_t *lat2dof, int nx, int ny, int nz, int Xp)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, n = (int64_t)nx * ny * nz;
  if (i >= n) return;
  const int X = (int)(i % nx);
  const int64_t row = i / nx, cpl = (int64_t)nz * ny * Xp;
  const uint32_t d = lat2dof[i];
#pragma unroll
  for (int c = 0; c < 3; ++c) vec[d + c] = lat[(2 * c + (X & 1)) * cpl + row * Xp + (X >> 1)];
}

// ---- structured bricks, ONE WARP PER BRICK, all three velocity components of a cell in one thread -----------------
// velocity_sb_kernel above is bound by the shared-memory pipe (69 % of its wavefront peak at 38 % FP64): half of its
// shared-memory traffic is the exchange of the symmetric off-diagonal stress between the three component warps, and its
// three warps meet at four block-wide barriers per brick.  Here lane l owns cell l of the brick with ALL 81 values in
// registers: the exchange disappears -- X_cd = eta w (d_d u_c + d_c u_d) is formed in registers, once along the d-lines
// for test component c and once along the c-lines for test component d (+29 % FP64 instructions, the price for not
// holding two 27-value accumulators next to the 81 values) -- and a warp depends on nobody: no __syncthreads, only
// __syncwarp.  Per cell the shared-memory traffic is 81 gathers + 81 read-modify-writes + staging / flush (0.85 kB per
// cell-component instead of 1.7 kB).  Each warp is its own block, persistent over the bricks b, b + G, ...; two
// staging buffers per warp ([3][SB_SLOTS] doubles each: the values of brick i+1 arrive by cp.async while brick i is
// computed; the buffer of brick i turns into its accumulation buffer once the gathers are done, first-touch stores
// instead of a zero fill as above); node lists double-buffered the same way.
constexpr size_t sbw_smem_bytes(int mode)
{ return (size_t)(2 * 3 * SB_SLOTS) * sizeof(double) + 2 * (size_t)(mode == 0 ? SB_LIST_BYTES : SB_BLOCK_BYTES) + (mode == 0 ? 0 : 2 * SB_PLIST * sizeof(double)); }

// value of (2 D'_DIR w) at point n, from the line of w through n along DIR
template <int DIR> __device__ __forceinline__ double dpoint(const double (&w)[27], int n)
{
  constexpr int S = stride3(DIR);
  const int pos = (n / S) % 3, base = n - pos * S;
  const double a = w[base], b = w[base + S], c = w[base + 2 * S];
  return pos == 0 ? fma(4.0, b - a, a - c) : (pos == 1 ? c - a : fma(4.0, c - b, a - c));
}
// T (+)= e_D (2D'_D)^T [ cvol w_q eta_q (e_D 2D'_D uc + e_C 2D'_C ud) ] : the off-diagonal pair (C, D) seen from test component C
template <int C, int D, bool PERQ>
__device__ __forceinline__ void sbw_offdiag(const double (&uc)[27], const double (&ud)[27], double (&T)[27], double eC, double eD, double cvol, const double *etab)
{
  const double e3 = 3.0 * eD, e4 = 4.0 * eD;
  for_lines<3, D>([&](int a, int b, int c) {
    double ga, gb, gc;
    dline(uc[a], uc[b], uc[c], ga, gb, gc);
    double xa = fma(eD, ga, eC * dpoint<C>(ud, a)), xb = fma(eD, gb, eC * dpoint<C>(ud, b)), xc = fma(eD, gc, eC * dpoint<C>(ud, c));
    if (PERQ) { xa *= cvol * qweight<3>(a) * etab[a * SUB]; xb *= cvol * qweight<3>(b) * etab[b * SUB]; xc *= cvol * qweight<3>(c) * etab[c * SUB]; }
    else { xa *= cvol * qweight<3>(a); xb *= cvol * qweight<3>(b); xc *= cvol * qweight<3>(c); }
    dTline_add(T[a], T[b], T[c], xa, xb, xc, eD, e3, e4);
  });
}
// rows of test component C: pressure term (MODE 1), diagonal direction, the two off-diagonal pairs, transposed interpolation
template <int C, int MODE, bool PERQ>
__device__ __forceinline__ void sbw_rows(const double (&u0)[27], const double (&u1)[27], const double (&u2)[27], double (&T)[27],
                                         const double (&ee)[3], double vol, double eta0, const double *etab, const double (&pn)[8], double ps)
{
  constexpr int D1 = (C + 1) % 3, D2 = (C + 2) % 3;
  const double (&uc)[27] = C == 0 ? u0 : (C == 1 ? u1 : u2);
  const double (&ud1)[27] = D1 == 0 ? u0 : (D1 == 1 ? u1 : u2);
  const double (&ud2)[27] = D2 == 0 ? u0 : (D2 == 1 ? u1 : u2);
  if (MODE == 1) {
    double pq[27];
    q1_to_quad<3>(pn, pq);
    const double f = -ee[C] * ps * vol;
#pragma unroll
    for (int q = 0; q < 27; ++q) pq[q] *= f * qweight<3>(q);
    for_lines<3, C>([&](int a, int b, int c) {
      T[a] = fma(-3.0, pq[a], pq[c] - pq[b]); T[b] = 4.0 * (pq[a] - pq[c]); T[c] = fma(3.0, pq[c], pq[b] - pq[a]);
    });
  }
  {
    const double f2 = 2.0 * ee[C] * ee[C] * vol * eta0;
    for_lines<3, C>([&](int a, int b, int c) {
      double ga, gb, gc;
      dline(uc[a], uc[b], uc[c], ga, gb, gc);
      if (PERQ) { ga *= f2 * qweight<3>(a) * etab[a * SUB]; gb *= f2 * qweight<3>(b) * etab[b * SUB]; gc *= f2 * qweight<3>(c) * etab[c * SUB]; }
      else { ga *= f2 * qweight<3>(a); gb *= f2 * qweight<3>(b); gc *= f2 * qweight<3>(c); }
      const double ta = fma(-3.0, ga, gc - gb), tb = 4.0 * (ga - gc), tc = fma(3.0, gc, gb - ga);
      if (MODE == 1) { T[a] += ta; T[b] += tb; T[c] += tc; }
      else { T[a] = ta; T[b] = tb; T[c] = tc; }
    });
  }
  sbw_offdiag<C, D1, PERQ>(uc, ud1, T, ee[C], ee[D1], vol * eta0, etab);
  sbw_offdiag<C, D2, PERQ>(uc, ud2, T, ee[C], ee[D2], vol * eta0, etab);
  interpT_all<3>(T);
}

template <int MODE, bool PERQ>
__global__ void __launch_bounds__(32, 6) velocity_sbw_kernel(const CellParams P, const unsigned char *lists, int n_batches)
{
  static_assert(MODE == 0 || MODE == 1, "A block, Stokes operator");
  constexpr int MU = SB_SLOTS;
  constexpr int LBYTES = MODE == 0 ? SB_LIST_BYTES : SB_BLOCK_BYTES;
  extern __shared__ __align__(16) double smem[];
  double *buf = smem;                                                            // [2][3][MU]
  unsigned char *sI = reinterpret_cast<unsigned char *>(buf + 2 * 3 * MU);       // [2][LBYTES]
  double *sP = reinterpret_cast<double *>(sI + 2 * LBYTES);                      // [2][SB_PLIST] pressure values (MODE 1)
  const int lane = threadIdx.x, G = gridDim.x;
  int b = blockIdx.x;
  if (b >= n_batches) return;
  const int x = sb_x(lane), y = sb_y(lane), z = sb_z(lane);
  int lb[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) lb[v] = x + sb_sy(v) * y + sb_sz(v) * z;
  unsigned first_mask = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if ((!(r & 1) || x == 0) && (!(r & 2) || y == 0) && (!(r & 4) || z == 0)) first_mask |= 1u << r;
  const double ee[3] = {P.ee[0], P.ee[1], P.ee[2]};

  const auto load_list = [&](unsigned char *dst, int brick) {
    const unsigned char *src = lists + (size_t)brick * SB_BLOCK_BYTES;
    for (int ch = lane; ch < LBYTES / 16; ch += 32) cp_async16(smem_u32(dst + 16 * ch), src + 16 * ch);
  };
  // values of one brick: lane l stages the nodes l, l + 32, ... of the list (ascending DoF order), three components each
  const auto issue_values = [&](const unsigned char *I, double *dstbuf, double *dstp) {
    const uint32_t *le = reinterpret_cast<const uint32_t *>(I);
    const uint16_t *ls = reinterpret_cast<const uint16_t *>(I + SB_LIST * 4);
    const uint32_t base = smem_u32(dstbuf);
    constexpr int SUBB = 4;
#pragma unroll 1
    for (int k0 = 0; k0 < 12; k0 += SUBB) {
      uint32_t e[SUBB], sl[SUBB];
#pragma unroll
      for (int k = 0; k < SUBB; ++k) { const int j = lane + 32 * (k0 + k); e[k] = le[j]; sl[k] = ls[j]; }
#pragma unroll
      for (int k = 0; k < SUBB; ++k) {
        const double *src = P.src_u + (e[k] & NODE_IDX_MASK);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const bool cons = ((e[k] >> (NODE_FLAG_SHIFT + d)) & 1u) && !P.plain;
          cp_async8_zfill(base + 8u * (sl[k] + d * MU), src + d, cons ? 0 : 8);
        }
      }
    }
    if (lane < 405 - 384) {
      const uint32_t e = le[lane + 384], sl = ls[lane + 384];
      const double *src = P.src_u + (e & NODE_IDX_MASK);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const bool cons = ((e >> (NODE_FLAG_SHIFT + d)) & 1u) && !P.plain;
        cp_async8_zfill(base + 8u * (sl + d * MU), src + d, cons ? 0 : 8);
      }
    }
    if (MODE == 1) {
      const uint32_t *lp = reinterpret_cast<const uint32_t *>(I + SB_LIST_BYTES);
      for (int j = lane; j < 75; j += 32) {
        const uint32_t e = lp[j];
        cp_async8_zfill(smem_u32(dstp + j), P.src_p + (e & NODE_IDX_MASK), ((e >> 31) && !P.plain) ? 0 : 8);
      }
    }
  };

  load_list(sI, b);
  cp_async_commit();
  cp_async_wait_all();
  __syncwarp();
  issue_values(sI, buf, sP);
  cp_async_commit();

  for (int it = 0; b < n_batches; b += G, ++it) {
    const int cur = it & 1, nxt = cur ^ 1;
    double *B = buf + cur * 3 * MU;
    unsigned char *I_cur = sI + cur * LBYTES, *I_nxt = sI + nxt * LBYTES;
    const int bn = b + G;
    cp_async_wait_all();
    __syncwarp();                       // values of this brick are in B; the flush of the previous brick is done with the other buffer
    if (bn < n_batches) load_list(I_nxt, bn);
    cp_async_commit();
    const double *etab = P.eta_t + (size_t)b * P.n_eta * SUB + lane;
    const double eta0 = PERQ ? 1.0 : etab[0];

    double u0[27], u1[27], u2[27];
    sb_gather<0>(u0, B, lb); sb_gather<0>(u1, B + MU, lb); sb_gather<0>(u2, B + 2 * MU, lb);
    double pn[8];
    if (MODE == 1) sb_gather_p<0>(pn, sP + cur * SB_PLIST, x, y, z);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) pn[i] = 0.0;
    }
    cp_async_wait_all();                // (this lane's chunks of the next node list)
    __syncwarp();                       // every gather from B is done; the next list is visible
    if (bn < n_batches) issue_values(I_nxt, buf + nxt * 3 * MU, sP + nxt * SB_PLIST);
    cp_async_commit();
    interp_all<3>(u0); interp_all<3>(u1); interp_all<3>(u2);

    if (MODE == 1) {
      // pressure rows -ps (div u, q): this cell's share is parked in B (free until the accumulation), summed lattice node by
      // lattice node in a fixed order and sent straight to global memory
      double dq[27], yv[8];
      const double f = -P.ps * P.vol;
      for_lines<3, 0>([&](int a, int b2, int c) { double ga, gb, gc; dline(u0[a], u0[b2], u0[c], ga, gb, gc); dq[a] = ee[0] * ga; dq[b2] = ee[0] * gb; dq[c] = ee[0] * gc; });
      for_lines<3, 1>([&](int a, int b2, int c) { double ga, gb, gc; dline(u1[a], u1[b2], u1[c], ga, gb, gc); dq[a] = fma(ee[1], ga, dq[a]); dq[b2] = fma(ee[1], gb, dq[b2]); dq[c] = fma(ee[1], gc, dq[c]); });
      for_lines<3, 2>([&](int a, int b2, int c) { double ga, gb, gc; dline(u2[a], u2[b2], u2[c], ga, gb, gc); dq[a] = fma(ee[2], ga, dq[a]); dq[b2] = fma(ee[2], gb, dq[b2]); dq[c] = fma(ee[2], gc, dq[c]); });
#pragma unroll
      for (int q = 0; q < 27; ++q) dq[q] *= f * qweight<3>(q);
      quad_to_q1T<3>(dq, yv);
#pragma unroll
      for (int m = 0; m < 8; ++m) B[m * SUB + lane] = yv[m];
      __syncwarp();
      const uint32_t *lp = reinterpret_cast<const uint32_t *>(I_cur + SB_LIST_BYTES);
      for (int t = lane; t < 75; t += 32) {
        const int X = t % 5, Y = (t / 5) % 5, Z = t / 25;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int cx = X - i, cy = Y - j, cz = Z - k;
              if (cx >= 0 && cx < 4 && cy >= 0 && cy < 4 && cz >= 0 && cz < 2) {
                const int ln = (cx & 1) | ((cy & 1) << 1) | (cz << 2) | ((cx >> 1) << 3) | ((cy >> 1) << 4);
                acc += B[(i + 2 * j + 4 * k) * SUB + ln];
              }
            }
        const uint32_t e = lp[t];
        const uint32_t idx = e & NODE_IDX_MASK;
        if (e >> 31) P.dst_p[idx] = P.plain ? 0.0 : P.src_p[idx];
        else if (!(e & NODE_SHARED_BIT)) P.dst_p[idx] = acc;
        else atomicAdd(P.dst_p + idx, acc);
      }
      __syncwarp();
    }

    {
      double T[27];
      sbw_rows<0, MODE, PERQ>(u0, u1, u2, T, ee, P.vol, eta0, etab, pn, P.ps);
      sb_accumulate<0>(T, B, lb, first_mask);
      sbw_rows<1, MODE, PERQ>(u0, u1, u2, T, ee, P.vol, eta0, etab, pn, P.ps);
      sb_accumulate<0>(T, B + MU, lb, first_mask);
      sbw_rows<2, MODE, PERQ>(u0, u1, u2, T, ee, P.vol, eta0, etab, pn, P.ps);
      sb_accumulate<0>(T, B + 2 * MU, lb, first_mask);
    }
    __syncwarp();
    // flush: private nodes with plain stores, shared nodes with one RED per DoF; constrained DoFs are skipped (dst = src
    // on those rows is the small list kernel after this one)
    {
      const uint32_t *le = reinterpret_cast<const uint32_t *>(I_cur);
      const uint16_t *ls = reinterpret_cast<const uint16_t *>(I_cur + SB_LIST * 4);
      const int n_private = (int)le[405];
#pragma unroll 1
      for (int k = 0; k < 13; ++k) {
        const int j = lane + 32 * k;
        if (j < 405) {
          const uint32_t e = le[j], sl = ls[j];
          double *dst = P.dst_u + (e & NODE_IDX_MASK);
          const double v0 = B[sl], v1 = B[MU + sl], v2 = B[2 * MU + sl];
          const bool priv = j < n_private;
          if (!((e >> NODE_FLAG_SHIFT) & 1u)) { if (priv) dst[0] = v0; else atomicAdd(dst, v0); }
          if (!((e >> (NODE_FLAG_SHIFT + 1)) & 1u)) { if (priv) dst[1] = v1; else atomicAdd(dst + 1, v1); }
          if (!((e >> (NODE_FLAG_SHIFT + 2)) & 1u)) { if (priv) dst[2] = v2; else atomicAdd(dst + 2, v2); }
        }
      }
    }
  }
}

// ---- pressure mass matrix: one warp per batch (one cell per lane), MASS_WARPS independent warps per block --------
// With one viscosity per cell the cell matrix is (ps^2 vol / eta) M1 x M1 x M1 with the exact 1D Q1 mass matrix
// M1 = [2 1; 1 2] / 6 (the 3-point Gauss rule integrates it exactly), applied as 2x2 sweeps on the 2^dim nodal
// values; with one viscosity per quadrature point the Q1 values go through the quadrature points.
constexpr int MASS_WARPS = 4;
constexpr int MASS_SU = 3;     // staging rounds unrolled together: ceil(75 / 32) nodes of a structured brick per lane
template <int DIM, bool PERQ>
__global__ void __launch_bounds__(32 * MASS_WARPS, PERQ ? 3 : 8) mass_kernel(const CellParams P, int n_batches)
{
  constexpr int NN = nn<DIM>(), NP = np<DIM>();
  extern __shared__ double smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x * MASS_WARPS + w;
  if (b >= n_batches) return;
  const int *info = P.batch_info + (size_t)b * BATCH_INFO;
  const int n_nodes_p = info[2], n_excl_p = info[3], n_cells_b = info[4];
  const int MP = P.max_nodes_p;
  double *s_src = smem + (size_t)w * 2 * MP;   // [MP]
  double *s_dst = s_src + MP;                  // [MP]
  const uint32_t *nlp = P.node_list_p + (size_t)b * MP;
  const uint16_t *locp = P.cell_loc_p + (size_t)b * NP * SUB + lane;
  uint16_t loc[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) loc[i] = locp[i * SUB];
  const double *etab = P.eta_t + (size_t)b * P.n_eta * SUB + lane;
  const double eta0 = PERQ ? 1.0 : etab[0];
  for (int j0 = lane; j0 < n_nodes_p; j0 += 32 * MASS_SU) {
    uint32_t e[MASS_SU];
    double v[MASS_SU];
#pragma unroll
    for (int i = 0; i < MASS_SU; ++i) e[i] = (j0 + 32 * i < n_nodes_p) ? nlp[j0 + 32 * i] : 0xffffffffu;
#pragma unroll
    for (int i = 0; i < MASS_SU; ++i) v[i] = (e[i] >> 31) ? 0.0 : P.src_p[e[i] & NODE_IDX_MASK];
#pragma unroll
    for (int i = 0; i < MASS_SU; ++i)
      if (j0 + 32 * i < n_nodes_p) { s_src[j0 + 32 * i] = v[i]; s_dst[j0 + 32 * i] = 0.0; }
  }
  __syncwarp();
  const bool active = lane < n_cells_b;
  double y[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) y[i] = s_src[loc[i]];
  const double f = P.ps * P.ps * P.vol;
  if (!PERQ) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const int S = 1 << d;
#pragma unroll
      for (int i = 0; i < NP; ++i)
        if (!(i & S)) {
          const double a = y[i], c = y[i + S];
          y[i] = fma(2.0, a, c); y[i + S] = fma(2.0, c, a);
        }
    }
    const double ce = f / (eta0 * (DIM == 3 ? 216.0 : 36.0));
#pragma unroll
    for (int i = 0; i < NP; ++i) y[i] *= ce;
  } else {
    double pn[NP], pq[NN];
#pragma unroll
    for (int i = 0; i < NP; ++i) pn[i] = y[i];
    q1_to_quad<DIM>(pn, pq);
#pragma unroll
    for (int q = 0; q < NN; ++q) pq[q] *= active ? (f / etab[q * SUB]) * qweight<DIM>(q) : 0.0;
    quad_to_q1T<DIM>(pq, y);
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) {     // node i of different cells never alias: one round per node
    if (active) s_dst[loc[i]] += y[i];
    __syncwarp();
  }
  for (int j = lane; j < n_nodes_p; j += 32) {
    const uint32_t e = nlp[j];
    const uint32_t idx = e & NODE_IDX_MASK;
    if (e >> 31) P.dst_p[idx] = P.src_p[idx];
    else if (j < n_excl_p) P.dst_p[idx] = s_dst[j];
    else atomicAdd(P.dst_p + idx, s_dst[j]);
  }
}

// ---- pressure mass matrix on structured bricks: one warp per brick, no slot tables, no accumulation rounds -------
// The warp stages the 5 x 5 x 3 pressure lattice of its brick (node list in lattice order, part of the brick's block
// in LevelDev::sb_lists), every lane applies the cell matrix to the 8 values of its cell and publishes the 8 results;
// then lattice node t sums the results of the (up to 8) cells around it in a fixed order and writes once: private
// node plain store, shared node one RED, constrained node dst = src.
constexpr int MSB_WARPS = 8;
template <bool PERQ>
__global__ void __launch_bounds__(32 * MSB_WARPS) mass_sb_kernel(const CellParams P, const unsigned char *lists, int n_batches)
{
  constexpr int NN = 27, NP = 8;
  __shared__ double s_x[MSB_WARPS][SB_PLIST];
  __shared__ double s_y[MSB_WARPS][NP * SUB];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x * MSB_WARPS + w;
  if (b >= n_batches) return;
  const uint32_t *pl = reinterpret_cast<const uint32_t *>(lists + (size_t)b * SB_BLOCK_BYTES + SB_LIST_BYTES);
  uint32_t e[3];
  double v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) e[k] = pl[min(lane + 32 * k, 74)];
  const double *etab = P.eta_t + (size_t)b * P.n_eta * SUB + lane;
  const double eta0 = PERQ ? 1.0 : etab[0];
#pragma unroll
  for (int k = 0; k < 3; ++k) v[k] = P.src_p[e[k] & NODE_IDX_MASK];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (lane + 32 * k < 75) s_x[w][lane + 32 * k] = (e[k] >> 31) ? 0.0 : v[k];
  __syncwarp();
  const int x = sb_x(lane), y = sb_y(lane), z = sb_z(lane);
  double yv[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) yv[i] = s_x[w][(x + (i & 1)) + 5 * (y + ((i >> 1) & 1)) + 25 * (z + (i >> 2))];
  const double f = P.ps * P.ps * P.vol;
  if (!PERQ) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int S = 1 << d;
#pragma unroll
      for (int i = 0; i < NP; ++i)
        if (!(i & S)) {
          const double a = yv[i], c = yv[i + S];
          yv[i] = fma(2.0, a, c); yv[i + S] = fma(2.0, c, a);
        }
    }
    const double ce = f / (eta0 * 216.0);
#pragma unroll
    for (int i = 0; i < NP; ++i) yv[i] *= ce;
  } else {
    double pn[NP], pq[NN];
#pragma unroll
    for (int i = 0; i < NP; ++i) pn[i] = yv[i];
    q1_to_quad<3>(pn, pq);
#pragma unroll
    for (int q = 0; q < NN; ++q) pq[q] *= (f / etab[q * SUB]) * qweight<3>(q);
    quad_to_q1T<3>(pq, yv);
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) s_y[w][i * SUB + lane] = yv[i];
  __syncwarp();
#pragma unroll
  for (int k3 = 0; k3 < 3; ++k3) {
    const int t = lane + 32 * k3;
    if (t >= 75) break;
    const int X = t % 5, Y = (t / 5) % 5, Z = t / 25;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int cx = X - i, cy = Y - j, cz = Z - k;
          if (cx >= 0 && cx < 4 && cy >= 0 && cy < 4 && cz >= 0 && cz < 2) {
            const int ln = (cx & 1) | ((cy & 1) << 1) | (cz << 2) | ((cx >> 1) << 3) | ((cy >> 1) << 4);
            acc += s_y[w][(i + 2 * j + 4 * k) * SUB + ln];
          }
        }
    const uint32_t idx = e[k3] & NODE_IDX_MASK;
    if (e[k3] >> 31) P.dst_p[idx] = v[k3];
    else if (!(e[k3] & NODE_SHARED_BIT)) P.dst_p[idx] = acc;
    else atomicAdd(P.dst_p + idx, acc);
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
  const double *geo;                // curved cells: blk_index table of the geometry (NULL = Cartesian)
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
    constexpr int GR = NN * (DIM * DIM + 1);
    const double *gq = P.geo ? P.geo + blk_index(cell, q * (DIM * DIM + 1), GR) : nullptr;      // entry k of the point: gq[k * GEO_BLOCK]
    const double JxW = gq ? gq[DIM * DIM * GEO_BLOCK] : qweight<DIM>(q) * P.vol;
    const double coef = JxW * P.eta[((cell / SUB) * P.n_eta + (P.n_eta == 1 ? 0 : q)) * SUB + (cell % SUB)];
    double ref[DIM];
    for (int r = 0; r < DIM; ++r) {
      double g = 1.0;
      for (int e = 0; e < DIM; ++e) g *= (e == r) ? q2d(qd[e], id[e]) : q2v(qd[e], id[e]);
      ref[r] = g;
    }
    for (int d = 0; d < DIM; ++d) {
      double g = 0.0;
      if (gq) { for (int r = 0; r < DIM; ++r) g = fma(ref[r], gq[(r * DIM + d) * GEO_BLOCK], g); }
      else g = ref[d] * P.ih[d];
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
      const double JxW = P.geo ? P.geo[blk_index(cell, q * (DIM * DIM + 1) + DIM * DIM, NN * (DIM * DIM + 1))] : qweight<DIM>(q) * P.vol;
      m += JxW * (P.ps * P.ps / P.eta[((cell / SUB) * P.n_eta + (P.n_eta == 1 ? 0 : q)) * SUB + (cell % SUB)]) * v * v;
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

// Sparse 1D prolongation rows: patch position r touches coarse nodes [first, first + count) with weights w[]
struct Row1D { int first, count; double w0, w1, w2; __device__ __forceinline__ double w(int k) const { return k == 0 ? w0 : (k == 1 ? w1 : w2); } };
template <int NPD> __device__ __forceinline__ Row1D prow(int r)
{
  if (NPD == 3) {
    if ((r & 1) == 0) return {r >> 1, 1, 1.0, 0.0, 0.0};
    if (r == 1) return {0, 3, 0.375, 0.75, -0.125};
    return {0, 3, -0.125, 0.75, 0.375};
  }
  if ((r & 1) == 0) return {r >> 1, 1, 1.0, 0.0, 0.0};
  return {0, 2, 0.5, 0.5, 0.0};
}
// transposed: coarse node j receives from patch positions pos[0..count) with weights w[]
struct Col1D {
  int count, p0, p1, p2; double w0, w1, w2;
  __device__ __forceinline__ double w(int k) const { return k == 0 ? w0 : (k == 1 ? w1 : w2); }
  __device__ __forceinline__ int pos(int k) const { return k == 0 ? p0 : (k == 1 ? p1 : p2); }
};
template <int NPD> __device__ __forceinline__ Col1D pcol(int j)
{
  if (NPD == 3) {
    if (j == 0) return {3, 0, 1, 3, 1.0, 0.375, -0.125};
    if (j == 1) return {3, 1, 2, 3, 0.75, 1.0, 0.75};
    return {3, 1, 3, 4, -0.125, 0.375, 1.0};
  }
  if (j == 0) return {2, 0, 1, 0, 1.0, 0.5, 0.0};
  return {2, 1, 2, 0, 0.5, 1.0, 0.0};
}

// One warp per parent cell.  Prolongation: the lanes stage the parent's coarse values (constrained ones as zero) in
// shared memory, then every lane evaluates the coarse function at its patch nodes with the sparse tensor-product
// rows; only the parent that owns a fine node writes it (all parents would write the same value), so no atomics
// and no valence weights are needed.
constexpr int TR_WARPS = 4;
template <int DIM, int NPD, int NCOMP> __global__ void __launch_bounds__(32 * TR_WARPS) prolongate_kernel(const TransferParams P)
{
  constexpr int NPP = 2 * NPD - 1;
  constexpr int NPATCH = DIM == 3 ? NPP * NPP * NPP : NPP * NPP;
  constexpr int NC = DIM == 3 ? NPD * NPD * NPD : NPD * NPD;
  __shared__ double s_c[TR_WARPS][NCOMP][NC];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t par = (int64_t)blockIdx.x * TR_WARPS + w;
  if (par >= P.n_parents) return;
  if (lane < NC) {
    const uint32_t ce = P.coarse_nodes[par * NC + lane];
    const uint32_t ci = ce & NODE_IDX_MASK;
#pragma unroll
    for (int c = 0; c < NCOMP; ++c) {
      const bool cons = (NCOMP == 1) ? (ce >> 31) : ((ce >> (NODE_FLAG_SHIFT + c)) & 1u);
      s_c[w][c][lane] = cons ? 0.0 : P.src[ci + c];
    }
  }
  __syncwarp();
  constexpr int ROUNDS = (NPATCH + 31) / 32;
  uint32_t pe[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) pe[r] = (lane + 32 * r < NPATCH) ? P.patch[par * NPATCH + lane + 32 * r] : 0u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    if (!(pe[r] >> 31)) continue;
    const int s = lane + 32 * r;
    const Row1D rx = prow<NPD>(s % NPP), ry = prow<NPD>((s / NPP) % NPP);
    const Row1D rz = DIM == 3 ? prow<NPD>(s / (NPP * NPP)) : Row1D{0, 1, 1.0, 0.0, 0.0};
    double acc[NCOMP];
#pragma unroll
    for (int c = 0; c < NCOMP; ++c) acc[c] = 0.0;
    for (int kz = 0; kz < rz.count; ++kz)
      for (int ky = 0; ky < ry.count; ++ky) {
        const double wzy = rz.w(kz) * ry.w(ky);
        const int base = (DIM == 3 ? (rz.first + kz) * NPD * NPD : 0) + (ry.first + ky) * NPD + rx.first;
        for (int kx = 0; kx < rx.count; ++kx) {
          const double wt = wzy * rx.w(kx);
#pragma unroll
          for (int c = 0; c < NCOMP; ++c) acc[c] = fma(wt, s_c[w][c][base + kx], acc[c]);
        }
      }
    const uint32_t fi = pe[r] & NODE_IDX_MASK;
#pragma unroll
    for (int c = 0; c < NCOMP; ++c) {
      if (P.add) P.dst[fi + c] += acc[c];
      else P.dst[fi + c] = acc[c];
    }
  }
}

// Restriction (transpose): the lanes stage the fine patch values this parent owns, then lane j < NC gathers them with
// the transposed rows and adds into coarse node j (coarse Dirichlet DoFs stay untouched = zero).
template <int DIM, int NPD, int NCOMP> __global__ void __launch_bounds__(32 * TR_WARPS) restrict_kernel(const TransferParams P)
{
  constexpr int NPP = 2 * NPD - 1;
  constexpr int NPATCH = DIM == 3 ? NPP * NPP * NPP : NPP * NPP;
  constexpr int NC = DIM == 3 ? NPD * NPD * NPD : NPD * NPD;
  __shared__ double s_f[TR_WARPS][NCOMP][NPATCH];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t par = (int64_t)blockIdx.x * TR_WARPS + w;
  if (par >= P.n_parents) return;
  constexpr int ROUNDS = (NPATCH + 31) / 32;
  uint32_t pe[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) pe[r] = (lane + 32 * r < NPATCH) ? P.patch[par * NPATCH + lane + 32 * r] : 0u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r)
    if (lane + 32 * r < NPATCH) {
      const bool own = (pe[r] >> 30) & 1u;
      const uint32_t fi = pe[r] & NODE_IDX_MASK;
#pragma unroll
      for (int c = 0; c < NCOMP; ++c) s_f[w][c][lane + 32 * r] = own ? P.src[fi + c] : 0.0;
    }
  __syncwarp();
  if (lane >= NC) return;
  const int j = lane;
  const uint32_t ce = P.coarse_nodes[par * NC + j];
  const Col1D cx = pcol<NPD>(j % NPD), cy = pcol<NPD>((j / NPD) % NPD);
  const Col1D cz = DIM == 3 ? pcol<NPD>(j / (NPD * NPD)) : Col1D{1, 0, 0, 0, 1.0, 0.0, 0.0};
  double acc[NCOMP];
#pragma unroll
  for (int c = 0; c < NCOMP; ++c) acc[c] = 0.0;
  for (int kz = 0; kz < cz.count; ++kz)
    for (int ky = 0; ky < cy.count; ++ky) {
      const double wzy = cz.w(kz) * cy.w(ky);
      const int base = (DIM == 3 ? cz.pos(kz) * NPP * NPP : 0) + cy.pos(ky) * NPP;
      for (int kx = 0; kx < cx.count; ++kx) {
        const double wt = wzy * cx.w(kx);
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) acc[c] = fma(wt, s_f[w][c][base + cx.pos(kx)], acc[c]);
      }
    }
  const uint32_t ci = ce & NODE_IDX_MASK;
#pragma unroll
  for (int c = 0; c < NCOMP; ++c) {
    const bool cons = (NCOMP == 1) ? (ce >> 31) : ((ce >> (NODE_FLAG_SHIFT + c)) & 1u);
    if (!cons && acc[c] != 0.0) atomicAdd(P.dst + ci + c, acc[c]);
  }
}

// ---- velocity transfers on structured fine levels: one block per FINE brick -----------------------------------
// A 4x4x2 brick of fine cells lies under 2x2x1 parents: the 5x5x3 coarse node lattice is interpolated to the brick's
// 9x9x5 fine lattice by three 1D sweeps in shared memory (two-cell macro line: coarse node I sits at fine position
// 2 I; fine position X takes the Q2 row p2row(X - 2 I0, .) on coarse nodes I0 .. I0 + 2, I0 = min(2 (X >> 2), NC - 3)),
// and the fine side goes through the brick's own node list (ascending DoF order: coalesced) with one owner bit per
// node, so every fine DoF is written (prolongation) / read (restriction) exactly once and the fine side needs no
// atomics.  Restriction is the exact transpose; coarse nodes are shared between bricks, so they take one RED each.
constexpr int TSB_THREADS = 96;
// 1D Q2 prolongation on a two-cell macro line (5 coarse -> 9 fine), on one cell (3 -> 5), and their transposes
__device__ __forceinline__ void pro9(const double (&c)[5], double (&f)[9])
{
  f[0] = c[0]; f[2] = c[1]; f[4] = c[2]; f[6] = c[3]; f[8] = c[4];
  f[1] = 0.375 * c[0] + 0.75 * c[1] - 0.125 * c[2]; f[3] = -0.125 * c[0] + 0.75 * c[1] + 0.375 * c[2];
  f[5] = 0.375 * c[2] + 0.75 * c[3] - 0.125 * c[4]; f[7] = -0.125 * c[2] + 0.75 * c[3] + 0.375 * c[4];
}
__device__ __forceinline__ void pro5(const double (&c)[3], double (&f)[5])
{
  f[0] = c[0]; f[2] = c[1]; f[4] = c[2];
  f[1] = 0.375 * c[0] + 0.75 * c[1] - 0.125 * c[2]; f[3] = -0.125 * c[0] + 0.75 * c[1] + 0.375 * c[2];
}
__device__ __forceinline__ void res9(const double (&f)[9], double (&c)[5])
{
  c[0] = f[0] + 0.375 * f[1] - 0.125 * f[3];
  c[1] = 0.75 * f[1] + f[2] + 0.75 * f[3];
  c[2] = -0.125 * f[1] + 0.375 * f[3] + f[4] + 0.375 * f[5] - 0.125 * f[7];
  c[3] = 0.75 * f[5] + f[6] + 0.75 * f[7];
  c[4] = -0.125 * f[5] + 0.375 * f[7] + f[8];
}
__device__ __forceinline__ void res5(const double (&f)[5], double (&c)[3])
{
  c[0] = f[0] + 0.375 * f[1] - 0.125 * f[3];
  c[1] = 0.75 * f[1] + f[2] + 0.75 * f[3];
  c[2] = -0.125 * f[1] + 0.375 * f[3] + f[4];
}
// staging slots of the five lattice points (X, Y, Z = 0..4) of one z-line
__device__ __forceinline__ void zline_slots(int X, int Y, int (&sl)[5])
{
  const int v = (X & 1) + 2 * (Y & 1);
  constexpr int B00 = sb_class_base(0, 0), B10 = sb_class_base(1, 0), B20 = sb_class_base(2, 0), B30 = sb_class_base(3, 0);
  constexpr int B01 = sb_class_base(0, 1), B11 = sb_class_base(1, 1), B21 = sb_class_base(2, 1), B31 = sb_class_base(3, 1);
  const int be = v == 0 ? B00 : (v == 1 ? B10 : (v == 2 ? B20 : B30)), bo = v == 0 ? B01 : (v == 1 ? B11 : (v == 2 ? B21 : B31));
  const int sy = (v & 1) ? 4 : 8, sz = v == 0 ? 52 : (v == 2 ? 36 : 24);
  const int o = (X >> 1) + sy * (Y >> 1);
  sl[0] = be + o; sl[2] = be + o + sz; sl[4] = be + o + 2 * sz;
  sl[1] = bo + o; sl[3] = bo + o + sz;
}

struct TransferSbParams {
  const unsigned char *tb;       // [n_bricks][TB_BYTES]
  const unsigned char *lists;    // fine level's sb_lists [n_bricks][SB_BLOCK_BYTES]
  const double *src; double *dst;
  int add;
  const double *minus;           // restriction only: the restricted vector is src - minus (the V-cycle's residual b - A x, never stored); may be null
};

// One thread per line of each sweep; all weights and loop bounds are compile-time constants.
__global__ void __launch_bounds__(TSB_THREADS) prolongate_sb_kernel(const TransferSbParams P)
{
  __shared__ __align__(16) unsigned char s_tb[TB_BYTES];
  __shared__ __align__(16) unsigned char s_l[SB_LIST_BYTES];
  __shared__ double sC[3][TB_COARSE], A1[3][3 * 5 * 9], A2[3][3 * 9 * 9], sO[3][SB_SLOTS];
  const int tid = threadIdx.x, b = blockIdx.x;
  {
    const uint4 *g1 = reinterpret_cast<const uint4 *>(P.tb + (size_t)b * TB_BYTES);
    const uint4 *g2 = reinterpret_cast<const uint4 *>(P.lists + (size_t)b * SB_BLOCK_BYTES);
    for (int i = tid; i < TB_BYTES / 16; i += TSB_THREADS) reinterpret_cast<uint4 *>(s_tb)[i] = g1[i];
    for (int i = tid; i < SB_LIST_BYTES / 16; i += TSB_THREADS) reinterpret_cast<uint4 *>(s_l)[i] = g2[i];
  }
  __syncthreads();
  const uint32_t *cl = reinterpret_cast<const uint32_t *>(s_tb);
  const unsigned char *ob = s_tb + TB_COARSE * 4;
  const int j0 = tid / 3, d0 = tid - 3 * j0;
  for (int j = j0; j < 75; j += 32) {                  // coarse values, constrained ones as zero
    const uint32_t e = cl[j];
    sC[d0][j] = ((e >> (NODE_FLAG_SHIFT + d0)) & 1u) ? 0.0 : P.src[(e & NODE_IDX_MASK) + d0];
  }
  __syncthreads();
  if (tid < 45) {                                      // x: lines (d, K, J): 5 coarse -> 9 fine
    const int d = tid / 15, jk = tid - 15 * d;
    double c[5], f[9];
#pragma unroll
    for (int i = 0; i < 5; ++i) c[i] = sC[d][jk * 5 + i];
    pro9(c, f);
#pragma unroll
    for (int i = 0; i < 9; ++i) A1[d][jk * 9 + i] = f[i];
  }
  __syncthreads();
  if (tid < 81) {                                      // y: lines (d, K, X)
    const int d = tid / 27, r = tid - 27 * d, K = r / 9, X = r - 9 * K;
    double c[5], f[9];
#pragma unroll
    for (int i = 0; i < 5; ++i) c[i] = A1[d][(K * 5 + i) * 9 + X];
    pro9(c, f);
#pragma unroll
    for (int i = 0; i < 9; ++i) A2[d][(K * 9 + i) * 9 + X] = f[i];
  }
  __syncthreads();
  for (int t = tid; t < 243; t += TSB_THREADS) {       // z: lines (d, Y, X): 3 -> 5, into the staging slots
    const int d = t / 81, yx = t - 81 * d, Y = yx / 9, X = yx - 9 * Y;
    double c[3], f[5];
    int sl[5];
#pragma unroll
    for (int i = 0; i < 3; ++i) c[i] = A2[d][i * 81 + yx];
    pro5(c, f);
    zline_slots(X, Y, sl);
#pragma unroll
    for (int i = 0; i < 5; ++i) sO[d][sl[i]] = f[i];
  }
  __syncthreads();
  const uint32_t *le = reinterpret_cast<const uint32_t *>(s_l);
  const uint16_t *ls = reinterpret_cast<const uint16_t *>(s_l + SB_LIST * 4);
  // x += P x_c is a read-modify-write of the fine vector: all loads of a thread first (13 independent ones in flight), then the
  // stores -- one load -> add -> store chain per loop trip exposed the memory latency 13 times per block
  {
    constexpr int NIT = (405 + 31) / 32;
    uint32_t idx[NIT];
    double v[NIT];
    bool w[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int j = j0 + 32 * k;
      w[k] = j < 405 && (ob[min(j, 404)] & 1);
      idx[k] = w[k] ? (le[j] & NODE_IDX_MASK) + d0 : 0u;
      v[k] = w[k] ? sO[d0][ls[j]] : 0.0;
    }
    if (P.add) {
#pragma unroll
      for (int k = 0; k < NIT; ++k) if (w[k]) v[k] += P.dst[idx[k]];
    }
#pragma unroll
    for (int k = 0; k < NIT; ++k) if (w[k]) P.dst[idx[k]] = v[k];
  }
}

__global__ void __launch_bounds__(TSB_THREADS) restrict_sb_kernel(const TransferSbParams P)
{
  __shared__ __align__(16) unsigned char s_tb[TB_BYTES];
  __shared__ __align__(16) unsigned char s_l[SB_LIST_BYTES];
  __shared__ double sO[3][SB_SLOTS], B2[3][3 * 9 * 9], B1[3][3 * 5 * 9];
  const int tid = threadIdx.x, b = blockIdx.x;
  {
    const uint4 *g1 = reinterpret_cast<const uint4 *>(P.tb + (size_t)b * TB_BYTES);
    const uint4 *g2 = reinterpret_cast<const uint4 *>(P.lists + (size_t)b * SB_BLOCK_BYTES);
    for (int i = tid; i < TB_BYTES / 16; i += TSB_THREADS) reinterpret_cast<uint4 *>(s_tb)[i] = g1[i];
    for (int i = tid; i < SB_LIST_BYTES / 16; i += TSB_THREADS) reinterpret_cast<uint4 *>(s_l)[i] = g2[i];
  }
  __syncthreads();
  const uint32_t *cl = reinterpret_cast<const uint32_t *>(s_tb);
  const unsigned char *ob = s_tb + TB_COARSE * 4;
  const uint32_t *le = reinterpret_cast<const uint32_t *>(s_l);
  const uint16_t *ls = reinterpret_cast<const uint16_t *>(s_l + SB_LIST * 4);
  const int j0 = tid / 3, d0 = tid - 3 * j0;
  {
    constexpr int NIT = (405 + 31) / 32;      // all loads of a thread in flight before the first shared-memory store
    double v[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int j = j0 + 32 * k;
      v[k] = 0.0;
      if (j < 405 && (ob[j] & 2)) {
        const uint32_t idx = (le[j] & NODE_IDX_MASK) + d0;
        v[k] = P.minus ? P.src[idx] - P.minus[idx] : P.src[idx];
      }
    }
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int j = j0 + 32 * k;
      if (j < 405) sO[d0][ls[j]] = v[k];
    }
  }
  __syncthreads();
  for (int t = tid; t < 243; t += TSB_THREADS) {       // z^T: lines (d, Y, X): 5 -> 3
    const int d = t / 81, yx = t - 81 * d, Y = yx / 9, X = yx - 9 * Y;
    double f[5], c[3];
    int sl[5];
    zline_slots(X, Y, sl);
#pragma unroll
    for (int i = 0; i < 5; ++i) f[i] = sO[d][sl[i]];
    res5(f, c);
#pragma unroll
    for (int i = 0; i < 3; ++i) B2[d][i * 81 + yx] = c[i];
  }
  __syncthreads();
  if (tid < 81) {                                      // y^T: lines (d, K, X): 9 -> 5
    const int d = tid / 27, r = tid - 27 * d, K = r / 9, X = r - 9 * K;
    double f[9], c[5];
#pragma unroll
    for (int i = 0; i < 9; ++i) f[i] = B2[d][(K * 9 + i) * 9 + X];
    res9(f, c);
#pragma unroll
    for (int i = 0; i < 5; ++i) B1[d][(K * 5 + i) * 9 + X] = c[i];
  }
  __syncthreads();
  if (tid < 45) {                                      // x^T: lines (d, K, J): 9 -> 5 coarse nodes, one RED per DoF
    const int d = tid / 15, jk = tid - 15 * d;
    double f[9], c[5];
#pragma unroll
    for (int i = 0; i < 9; ++i) f[i] = B1[d][jk * 9 + i];
    res9(f, c);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const uint32_t e = cl[jk * 5 + i];
      if (!((e >> (NODE_FLAG_SHIFT + d)) & 1u) && c[i] != 0.0) atomicAdd(P.dst + (e & NODE_IDX_MASK) + d, c[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
static CellParams make_params(const LevelDev &L)
{
  CellParams P{};
  P.batch_info = L.batch_info;
  P.node_list_u = L.node_list_u; P.node_slot_u = L.node_slot_u; P.cell_loc_u = L.cell_loc_u; P.list_stride_u = L.list_stride_u;
  P.node_list_p = L.node_list_p; P.cell_loc_p = L.cell_loc_p;
  P.max_nodes_u = L.max_nodes_u; P.max_nodes_p = L.max_nodes_p;
  P.eta_t = L.eta; P.n_eta = L.n_eta;
  double vol = 1.0;
  for (int d = 0; d < L.dim; ++d) { P.ee[d] = 0.5 / (L.h[d] * HOST_GAUSS_A); vol *= L.h[d]; }
  P.vol = vol; P.ps = L.ps;
  return P;
}

template <int DIM, int MODE, bool PERQ, bool STRUCT> static int launch_velocity_t(smf_ctx *ctx, LevelDev &L, CellParams &P)
{
  const size_t smem = vel_smem_bytes<DIM>(L.max_nodes_u, L.max_nodes_p, MODE);
  static size_t configured_dev[MAX_DEVICES] = {0};      // (function attributes are per device; a process may hold one context per device)
  size_t &configured = configured_dev[ctx->cfg.device % MAX_DEVICES];
  if (smem > configured) {
    SMF_CUDA(cudaFuncSetAttribute(velocity_kernel<DIM, MODE, PERQ, STRUCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const bool fine = (&L == &ctx->lev(SMF_ACTIVE)) || (&L == &ctx->levels[ctx->cfg.n_levels - 1]);
  {
    ProfScope prof(ctx, MODE == 1 ? B_STOKES : (MODE == 2 ? B_BT : (fine ? B_A_FINE : B_A_COARSE)));
    velocity_kernel<DIM, MODE, PERQ, STRUCT><<<L.n_batches, 32 * DIM, smem, ctx->stream>>>(P);
  }
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}
__global__ void copy_entries_kernel(double *dst, const double *src, const uint32_t *idx, int64_t n);

template <int MODE, bool PERQ> static int launch_velocity_sb(smf_ctx *ctx, LevelDev &L, CellParams &P)
{
  constexpr size_t smem = sbk_smem_bytes(MODE);
  static int resident_dev[MAX_DEVICES] = {0};
  int &resident = resident_dev[ctx->cfg.device % MAX_DEVICES];
  if (resident == 0) {
    SMF_CUDA(cudaFuncSetAttribute(velocity_sb_kernel<MODE, PERQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, sms = 148;
    SMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->cfg.device));
    SMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, velocity_sb_kernel<MODE, PERQ>, SBK_THREADS, smem));
    if (per_sm < 1) return ctx->fail(SMF_ERR_CUDA, "velocity_sb_kernel does not fit on an SM");
    resident = per_sm * sms;
  }
  const bool fine = (&L == &ctx->lev(SMF_ACTIVE)) || (&L == &ctx->levels[ctx->cfg.n_levels - 1]);
  {
    ProfScope prof(ctx, MODE == 1 ? B_STOKES : (MODE == 2 ? B_BT : (fine ? B_A_FINE : B_A_COARSE)));
    int grid = std::min(L.n_batches, resident);
    if (ctx->max_grid_blocks > 0) grid = std::min(grid, ctx->max_grid_blocks);
    velocity_sb_kernel<MODE, PERQ><<<grid, SBK_THREADS, smem, ctx->stream>>>(P, L.sb_lists, L.n_batches);
    if (MODE != 2 && !P.plain && L.n_constrained_u) {      // unit diagonal on the constrained rows: dst = src
      copy_entries_kernel<<<(unsigned)((L.n_constrained_u + 255) / 256), 256, 0, ctx->stream>>>(P.dst_u, P.src_u, L.constrained_u, L.n_constrained_u);
      ctx->launches++;
    }
  }
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}
template <int MODE, bool PERQ> static int launch_velocity_sbw(smf_ctx *ctx, LevelDev &L, CellParams &P)
{
  constexpr size_t smem = sbw_smem_bytes(MODE);
  static int resident_dev[MAX_DEVICES] = {0};
  int &resident = resident_dev[ctx->cfg.device % MAX_DEVICES];
  if (resident == 0) {
    SMF_CUDA(cudaFuncSetAttribute(velocity_sbw_kernel<MODE, PERQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, sms = 148;
    SMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->cfg.device));
    SMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, velocity_sbw_kernel<MODE, PERQ>, 32, smem));
    if (per_sm < 1) return ctx->fail(SMF_ERR_CUDA, "velocity_sbw_kernel does not fit on an SM");
    resident = per_sm * sms;
  }
  const bool fine = (&L == &ctx->lev(SMF_ACTIVE)) || (&L == &ctx->levels[ctx->cfg.n_levels - 1]);
  {
    ProfScope prof(ctx, MODE == 1 ? B_STOKES : (fine ? B_A_FINE : B_A_COARSE));
    int grid = std::min(L.n_batches, resident);
    if (ctx->max_grid_blocks > 0) grid = std::min(grid, ctx->max_grid_blocks);
    velocity_sbw_kernel<MODE, PERQ><<<grid, 32, smem, ctx->stream>>>(P, L.sb_lists, L.n_batches);
    if (!P.plain && L.n_constrained_u) {      // unit diagonal on the constrained rows: dst = src
      copy_entries_kernel<<<(unsigned)((L.n_constrained_u + 255) / 256), 256, 0, ctx->stream>>>(P.dst_u, P.src_u, L.constrained_u, L.n_constrained_u);
      ctx->launches++;
    }
  }
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}
template <int DIM, int MODE> static int launch_velocity(smf_ctx *ctx, LevelDev &L, CellParams &P)
{
  if (DIM == 3 && MODE != 2 && L.structured && ctx->use_pipelined && ctx->use_warp_brick && L.sb_lists) {
    if (L.n_eta != 1) return launch_velocity_sbw<MODE == 2 ? 0 : MODE, true>(ctx, L, P);
    return launch_velocity_sbw<MODE == 2 ? 0 : MODE, false>(ctx, L, P);
  }
  if (DIM == 3 && L.structured && ctx->use_pipelined && L.sb_lists) {
    if (MODE != 2 && L.n_eta != 1) return launch_velocity_sb<MODE, true>(ctx, L, P);
    return launch_velocity_sb<MODE, false>(ctx, L, P);
  }
  if (DIM == 3 && L.structured) {
    if (MODE != 2 && L.n_eta != 1) return launch_velocity_t<3, MODE, true, true>(ctx, L, P);
    return launch_velocity_t<3, MODE, false, true>(ctx, L, P);
  }
  if (MODE != 2 && L.n_eta != 1) return launch_velocity_t<DIM, MODE, true, false>(ctx, L, P);
  return launch_velocity_t<DIM, MODE, false, false>(ctx, L, P);
}


// ---- general cell operation: every branch of StokesOperator::local_apply ---------------------------------------
// Compressible / prescribed-dilation / Newton-derivative terms (matrix_free_operators.cc:400-446, :464-538) and the
// deviatoric A block (:999-1005) couple all components of the symmetric gradient at a quadrature point, which the
// component-per-warp kernels above do not exchange.  This kernel is the straightforward restatement -- one thread per
// cell, the cell's values in local memory, basis functions evaluated on the fly -- used only when one of those
// flags is set (no BASELINE configuration sets them); it is correct for every flag combination and not tuned.
struct GeneralParams {
  int64_t n_cells;
  const uint32_t *cell_nodes_u, *cell_nodes_p;
  const double *src_u, *src_p;
  double *dst_u, *dst_p;
  const double *eta; int n_eta;
  double ih[3], vol, ps;
  int mode;                 // 0 = A block, 1 = Stokes operator, 2 = B^T
  const double *geo;        // curved cells: rows q * (dim*dim + 1) + k of a blk_index table: inverse Jacobian (d xi_r / d x_d, r-major), JxW; NULL = Cartesian
  const uint32_t *cell_nodes_u_t, *cell_nodes_p_t;   // curved_cell_kernel: transposed maps (blk_index tables, rows = nodes)
  int compressible, newton, symmetrize, average, dilation, plain;
  const double *eps, *deta, *detadp, *dil_lhs, *ddil, *ddildp;     // [cell][q][n_sym] / [cell][q]
};

template <int DIM> __device__ __forceinline__ double sym_contract(const double *t, const double (&g)[DIM][DIM])
{
  // t: unrolled symmetric tensor (00, 11, (22,) 01, (02, 12)); g: symmetric DIM x DIM
  if (DIM == 2) return t[0] * g[0][0] + t[1] * g[1][1] + 2.0 * t[2] * g[0][1];
  return t[0] * g[0][0] + t[1] * g[1][1] + t[2] * g[2][2] + 2.0 * (t[3] * g[0][1] + t[4] * g[0][2] + t[5] * g[1][2]);
}
template <int DIM> __device__ __forceinline__ double sym_entry(const double *t, int i, int j)
{
  if (i == j) return t[i];
  if (DIM == 2) return t[2];
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  return lo == 0 ? (hi == 1 ? t[3] : t[4]) : t[5];
}
template <int DIM> __device__ __forceinline__ double sym_trace(const double *t) { return DIM == 2 ? t[0] + t[1] : t[0] + t[1] + t[2]; }

template <int DIM> __global__ void __launch_bounds__(64) general_cell_kernel(const GeneralParams P)
{
  constexpr int NN = nn<DIM>(), NP = np<DIM>(), NS = DIM * (DIM + 1) / 2;
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= P.n_cells) return;
  double u[DIM][NN], T[DIM][NN], pn[NP], Tp[NP];
  for (int i = 0; i < NN; ++i) {
    const uint32_t e = P.cell_nodes_u[cell * NN + i];
    const uint32_t idx = e & NODE_IDX_MASK;
    for (int c = 0; c < DIM; ++c) {
      const bool cons = ((e >> (NODE_FLAG_SHIFT + c)) & 1u) && !P.plain;
      u[c][i] = (cons || P.mode == 2) ? 0.0 : P.src_u[idx + c];
      T[c][i] = 0.0;
    }
  }
  // inverse Jacobian and JxW at quadrature point q (Cartesian: diag(1/h), w_q vol)
  auto geometry = [&](int q, double (&Ji)[DIM][DIM], double &JxW) {
    if (P.geo) {
      const double *g = P.geo + blk_index(cell, q * (DIM * DIM + 1), NN * (DIM * DIM + 1));
      for (int r = 0; r < DIM; ++r) for (int d = 0; d < DIM; ++d) Ji[r][d] = g[(r * DIM + d) * GEO_BLOCK];
      JxW = g[DIM * DIM * GEO_BLOCK];
    } else {
      for (int r = 0; r < DIM; ++r) for (int d = 0; d < DIM; ++d) Ji[r][d] = r == d ? P.ih[d] : 0.0;
      JxW = qweight<DIM>(q) * P.vol;
    }
  };
  // real-space gradient of the scalar Q2 shape function i at point q
  auto shape_grad = [&](const int (&qd)[3], int i, const double (&Ji)[DIM][DIM], double (&dphi)[DIM]) {
    int id[3] = {i % 3, (i / 3) % 3, DIM == 3 ? i / 9 : 0};
    double ref[DIM];
    for (int r = 0; r < DIM; ++r) {
      double v = 1.0;
      for (int e = 0; e < DIM; ++e) v *= (e == r) ? q2d(qd[e], id[e]) : q2v(qd[e], id[e]);
      ref[r] = v;
    }
    for (int d = 0; d < DIM; ++d) {
      double v = 0.0;
      for (int r = 0; r < DIM; ++r) v = fma(ref[r], Ji[r][d], v);
      dphi[d] = v;
    }
  };
  for (int a = 0; a < NP; ++a) {
    pn[a] = 0.0; Tp[a] = 0.0;
    if (P.mode != 0) {
      const uint32_t e = P.cell_nodes_p[cell * NP + a];
      pn[a] = ((e >> 31) && !P.plain) ? 0.0 : P.src_p[e & NODE_IDX_MASK];
    }
  }
  const bool deviatoric = P.compressible || P.dilation;
  const bool newton = P.mode == 1 && P.newton;
  // symmetric gradient, divergence and pressure at quadrature point q
  auto evaluate = [&](int q, const double (&Ji)[DIM][DIM], double (&sg)[DIM][DIM], double &div, double &valp) {
    int qd[3] = {q % 3, (q / 3) % 3, DIM == 3 ? q / 9 : 0};
    double g[DIM][DIM];
    for (int c = 0; c < DIM; ++c) for (int d = 0; d < DIM; ++d) g[c][d] = 0.0;
    if (P.mode != 2)
      for (int i = 0; i < NN; ++i) {
        double dphi[DIM];
        shape_grad(qd, i, Ji, dphi);
        for (int c = 0; c < DIM; ++c) for (int d = 0; d < DIM; ++d) g[c][d] = fma(u[c][i], dphi[d], g[c][d]);
      }
    div = 0.0;
    for (int c = 0; c < DIM; ++c) { div += g[c][c]; for (int d = 0; d < DIM; ++d) sg[c][d] = 0.5 * (g[c][d] + g[d][c]); }
    valp = 0.0;
    if (P.mode != 0)
      for (int a = 0; a < NP; ++a) {
        int ad[3] = {a % 2, (a / 2) % 2, DIM == 3 ? a / 4 : 0};
        double v = pn[a];
        for (int d = 0; d < DIM; ++d) v *= q1v(qd[d], ad[d]);
        valp += v;
      }
  };
  // averaged Newton factors (:400-446)
  double deta_deps_times_sym_grad_u = 0.0, deta_dp_times_p = 0.0, eps_times_sym_grad_u_JxW = 0.0, theta_times_div_u_JxW = 0.0;
  if (newton && P.average)
    for (int q = 0; q < NN; ++q) {
      double sg[DIM][DIM], div, valp, Ji[DIM][DIM], JxW;
      geometry(q, Ji, JxW);
      evaluate(q, Ji, sg, div, valp);
      const int64_t tq = cell * NN + q;
      deta_deps_times_sym_grad_u += sym_contract<DIM>(P.deta + tq * NS, sg);
      deta_dp_times_p += P.detadp[tq] * valp;
      if (P.symmetrize) {
        eps_times_sym_grad_u_JxW += sym_contract<DIM>(P.eps + tq * NS, sg) * JxW;
        if (deviatoric) theta_times_div_u_JxW += sym_trace<DIM>(P.eps + tq * NS) * div * JxW;
      }
    }
  for (int q = 0; q < NN; ++q) {
    double sg[DIM][DIM], div_u, val_p, Ji[DIM][DIM], JxW;
    geometry(q, Ji, JxW);
    evaluate(q, Ji, sg, div_u, val_p);
    const int64_t tq = cell * NN + q;
    const double viscosity_x_2 = P.mode == 2 ? 0.0 : 2.0 * P.eta[((cell / SUB) * P.n_eta + (P.n_eta == 1 ? 0 : q)) * SUB + (cell % SUB)];
    double pressure_terms = -P.ps * div_u;
    double vt[DIM][DIM];
    for (int c = 0; c < DIM; ++c) for (int d = 0; d < DIM; ++d) vt[c][d] = viscosity_x_2 * sg[c][d];
    if (P.mode == 2) for (int d = 0; d < DIM; ++d) vt[d][d] -= P.ps * val_p;
    if (P.mode == 1) {
      if (P.dilation) pressure_terms -= P.ps * P.ps * P.dil_lhs[tq] * val_p;
      for (int d = 0; d < DIM; ++d) vt[d][d] -= P.ps * val_p;
    }
    if (deviatoric) for (int d = 0; d < DIM; ++d) vt[d][d] -= viscosity_x_2 / 3.0 * div_u;
    if (newton) {
      const double *eps = P.eps + tq * NS, *deta = P.deta + tq * NS;
      double eps_times_sym_grad_u = 0.0, theta_times_div_u = 0.0;
      if (P.average) {
        if (P.symmetrize) eps_times_sym_grad_u = eps_times_sym_grad_u_JxW / JxW;
      } else {
        deta_deps_times_sym_grad_u = sym_contract<DIM>(deta, sg);
        deta_dp_times_p = P.detadp[tq] * val_p;
        if (P.symmetrize) eps_times_sym_grad_u = sym_contract<DIM>(eps, sg);
      }
      for (int c = 0; c < DIM; ++c)
        for (int d = 0; d < DIM; ++d) {
          const double e_cd = sym_entry<DIM>(eps, c, d), n_cd = sym_entry<DIM>(deta, c, d);
          vt[c][d] += (P.symmetrize ? e_cd * deta_deps_times_sym_grad_u + n_cd * eps_times_sym_grad_u : 2.0 * e_cd * deta_deps_times_sym_grad_u)
                      + 2.0 * e_cd * deta_dp_times_p;
        }
      if (deviatoric) {
        const double tr = sym_trace<DIM>(eps);
        if (P.symmetrize) {
          theta_times_div_u = P.average ? theta_times_div_u_JxW / JxW : tr * div_u;
          for (int c = 0; c < DIM; ++c) for (int d = 0; d < DIM; ++d) vt[c][d] -= (1.0 / 3.0) * sym_entry<DIM>(deta, c, d) * theta_times_div_u;
          for (int d = 0; d < DIM; ++d) vt[d][d] -= (1.0 / 3.0) * tr * deta_deps_times_sym_grad_u;
        } else {
          for (int d = 0; d < DIM; ++d) vt[d][d] -= (2.0 / 3.0) * tr * deta_deps_times_sym_grad_u;
        }
        if (P.dilation)
          pressure_terms += (sym_contract<DIM>(P.ddil + tq * NS, sg) + P.ddildp[tq] * P.ps * val_p) * P.ps;
      }
    }
    // submit_symmetric_gradient / submit_value, integrate
    int qd[3] = {q % 3, (q / 3) % 3, DIM == 3 ? q / 9 : 0};
    for (int c = 0; c < DIM; ++c) for (int d = 0; d < DIM; ++d) vt[c][d] *= JxW;
    for (int i = 0; i < NN; ++i) {
      double dphi[DIM];
      shape_grad(qd, i, Ji, dphi);
      for (int c = 0; c < DIM; ++c) {
        double acc = T[c][i];
        for (int d = 0; d < DIM; ++d) acc = fma(vt[c][d], dphi[d], acc);
        T[c][i] = acc;
      }
    }
    if (P.mode == 1)
      for (int a = 0; a < NP; ++a) {
        int ad[3] = {a % 2, (a / 2) % 2, DIM == 3 ? a / 4 : 0};
        double v = pressure_terms * JxW;
        for (int d = 0; d < DIM; ++d) v *= q1v(qd[d], ad[d]);
        Tp[a] += v;
      }
  }
  // distribute_local_to_global: nothing is added to constrained rows
  for (int i = 0; i < NN; ++i) {
    const uint32_t e = P.cell_nodes_u[cell * NN + i];
    const uint32_t idx = e & NODE_IDX_MASK;
    for (int c = 0; c < DIM; ++c)
      if (!((e >> (NODE_FLAG_SHIFT + c)) & 1u)) atomicAdd(P.dst_u + idx + c, T[c][i]);
  }
  if (P.mode == 1)
    for (int a = 0; a < NP; ++a) {
      const uint32_t e = P.cell_nodes_p[cell * NP + a];
      if (!(e >> 31)) atomicAdd(P.dst_p + (e & NODE_IDX_MASK), Tp[a]);
    }
}

// ---- curved cells (geometry = 1): sum factorisation with per-quadrature-point Jacobians ------------------------
// BASELINE config 3 (spherical shell, MappingQ(4)): the inverse Jacobian is a full matrix that changes from point to point,
// so all components of the gradient meet at every quadrature point (grad_x u = grad_xi u J^-1) and the component-per-warp
// split of the Cartesian kernels does not apply.  One thread per cell: the values of the three components at the 27 Gauss
// points live in the thread's own column of shared memory (lane-fastest, conflict-free, no barriers), the 81 accumulated
// test-side values in registers; per point 4 small (3 x 3) contractions: reference gradient by collocation
// (D'/a along the three lines through the point), J^-1, the stress, J^-T, and the transposed collocation derivative.
// HBM per cell: 27 x 10 doubles of geometry (2,160 B; a thread block reads one contiguous 138 KB chunk, blk_index) next to
// the ~400 B of vector data, i.e. the kernel is bound by the geometry stream (G term).  The node lists of
// unstructured meshes have no collision-free rounds, results go to global memory with one RED per DoF.
constexpr int CV_THREADS = 64;
// collocation derivative at Gauss point i from the values at the Gauss points j of the line, per unit reference length
__host__ __device__ constexpr double dpc(int i, int j)
{
  return (i == 0 ? (j == 0 ? -1.5 : (j == 1 ? 2.0 : -0.5)) : (i == 1 ? (j == 0 ? -0.5 : (j == 1 ? 0.0 : 0.5)) : (j == 0 ? 0.5 : (j == 1 ? -2.0 : 1.5))))
         / HOST_GAUSS_A;
}
template <int MODE, bool PERQ>
__global__ void __launch_bounds__(CV_THREADS, 4) curved_cell_kernel(const GeneralParams P)
{
  static_assert(MODE >= 0 && MODE <= 2, "A block, Stokes operator, B^T");
  static_assert(CV_THREADS == GEO_BLOCK, "one thread block reads one block of the per-cell tables");
  constexpr int NN = 27, NP = 8, G = 10;
  extern __shared__ __align__(16) double cv_smem[];
  const int64_t cell = (int64_t)blockIdx.x * CV_THREADS + threadIdx.x;
  if (cell >= P.n_cells) return;                      // (no barriers below: every thread works in its own column)
  double *U = cv_smem + threadIdx.x;                  // U[(c * 27 + q) * CV_THREADS]
  const uint32_t *nodes_u = P.cell_nodes_u_t + blk_index(cell, 0, NN);   // transposed maps: node i at nodes_u[i * GEO_BLOCK]
  const uint32_t *nodes_p = P.cell_nodes_p_t + blk_index(cell, 0, NP);
  if (MODE != 2) {
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      double v[NN];
#pragma unroll
      for (int i = 0; i < NN; ++i) {
        const uint32_t e = nodes_u[i * GEO_BLOCK];
        const bool cons = ((e >> (NODE_FLAG_SHIFT + c)) & 1u) && !P.plain;
        v[i] = cons ? 0.0 : P.src_u[(e & NODE_IDX_MASK) + c];
      }
      interp_all<3>(v);
#pragma unroll
      for (int q = 0; q < NN; ++q) U[(c * NN + q) * CV_THREADS] = v[q];
    }
  }
  // (volatile: every use below is a fresh 8-byte shared-memory load -- otherwise the compiler keeps all 81 values in
  //  registers next to the 81 accumulators and spills; the loads are conflict-free and cheap)
  const volatile double *Uv = U;
  double pn[NP], Tp[NP];
#pragma unroll
  for (int a = 0; a < NP; ++a) {
    pn[a] = 0.0; Tp[a] = 0.0;
    if (MODE != 0) {
      const uint32_t e = nodes_p[a * GEO_BLOCK];
      pn[a] = ((e >> 31) && !P.plain) ? 0.0 : P.src_p[e & NODE_IDX_MASK];
    }
  }
  double T[3][NN];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int i = 0; i < NN; ++i) T[c][i] = 0.0;
  const double *geo = P.geo + blk_index(cell, 0, NN * G);               // row q * 10 + k at geo[row * GEO_BLOCK]
  const double *etab = P.eta + (cell / SUB) * P.n_eta * SUB + (cell % SUB);
  const double eta_cell = (MODE != 2 && !PERQ) ? etab[0] : 1.0;
  // geometry of the next point travels while the current one is being worked on
  double gn[G];
#pragma unroll
  for (int k = 0; k < G; ++k) gn[k] = geo[k * GEO_BLOCK];
#pragma unroll
  for (int q = 0; q < NN; ++q) {
    const int qd[3] = {q % 3, (q / 3) % 3, q / 9};
    double Ji[3][3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Ji[k / 3][k % 3] = gn[k];
    const double JxW = gn[9];
    if (q + 1 < NN) {
#pragma unroll
      for (int k = 0; k < G; ++k) gn[k] = geo[((q + 1) * G + k) * GEO_BLOCK];
    }
    double S[3][3];                                   // stress times JxW
    double div = 0.0;
    if (MODE != 2) {
      double Gx[3][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double u0 = Uv[(c * NN + q) * CV_THREADS];
        double gh[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int st = stride3(r), base = q - qd[r] * st;
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (dpc(qd[r], j) != 0.0) acc = fma(dpc(qd[r], j), j == qd[r] ? u0 : Uv[(c * NN + base + j * st) * CV_THREADS], acc);
          gh[r] = acc;
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) Gx[c][d] = gh[0] * Ji[0][d] + gh[1] * Ji[1][d] + gh[2] * Ji[2][d];
      }
      div = Gx[0][0] + Gx[1][1] + Gx[2][2];
      const double ew = (PERQ ? etab[q * SUB] : eta_cell) * JxW;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d) S[c][d] = ew * (Gx[c][d] + Gx[d][c]);
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d) S[c][d] = 0.0;
    }
    if (MODE != 0) {
      double valp = 0.0;
#pragma unroll
      for (int a = 0; a < NP; ++a) valp = fma(pn[a], q1v(qd[0], a & 1) * q1v(qd[1], (a >> 1) & 1) * q1v(qd[2], a >> 2), valp);
      const double pw = P.ps * valp * JxW;
      S[0][0] -= pw; S[1][1] -= pw; S[2][2] -= pw;
      if (MODE == 1) {
        const double pt = -P.ps * div * JxW;
#pragma unroll
        for (int a = 0; a < NP; ++a) Tp[a] = fma(pt, q1v(qd[0], a & 1) * q1v(qd[1], (a >> 1) & 1) * q1v(qd[2], a >> 2), Tp[a]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double sh = S[c][0] * Ji[r][0] + S[c][1] * Ji[r][1] + S[c][2] * Ji[r][2];
        const int st = stride3(r), base = q - qd[r] * st;
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (dpc(qd[r], j) != 0.0) T[c][base + j * st] = fma(dpc(qd[r], j), sh, T[c][base + j * st]);
      }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    interpT_all<3>(T[c]);
#pragma unroll
    for (int i = 0; i < NN; ++i) {
      const uint32_t e = nodes_u[i * GEO_BLOCK];
      if (!((e >> (NODE_FLAG_SHIFT + c)) & 1u)) atomicAdd(P.dst_u + (e & NODE_IDX_MASK) + c, T[c][i]);
    }
  }
  if (MODE == 1) {
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      const uint32_t e = nodes_p[a * GEO_BLOCK];
      if (!(e >> 31)) atomicAdd(P.dst_p + (e & NODE_IDX_MASK), Tp[a]);
    }
  }
}

// Same kernel with the geometry stream taken off the threads: the 64 cells of a block own one contiguous chunk of the
// geometry table (blk_index), so thread 0 bulk-copies the 5,120 bytes of a quadrature point into a shared-memory stage
// (cp.async.bulk = the TMA engine, completion on an mbarrier) CV_STAGES points ahead; no registers are tied up while the
// data travels and the loads do not depend on the warps' progress.
constexpr int CV_STAGES = 2;
template <int MODE, bool PERQ>
__global__ void __launch_bounds__(CV_THREADS, 4) curved_cell_kernel_tma(const GeneralParams P)
{
  static_assert(MODE >= 0 && MODE <= 2, "A block, Stokes operator, B^T");
  static_assert(CV_THREADS == GEO_BLOCK, "one thread block reads one block of the per-cell tables");
  constexpr int NN = 27, NP = 8, G = 10;
  extern __shared__ __align__(128) double cv_smem[];
  // [CV_STAGES][10][64] geometry of one quadrature point of the block's 64 cells per stage (bulk-copied, 5,120 B each),
  // then the threads' columns of point values; one mbarrier per stage
  double *stage = cv_smem;
  double *U = cv_smem + CV_STAGES * G * CV_THREADS + threadIdx.x;          // U[(c * 27 + q) * CV_THREADS]
  __shared__ __align__(8) uint64_t bar[CV_STAGES];
  const int tid = threadIdx.x;
  const int64_t cell0 = (int64_t)blockIdx.x * CV_THREADS, cell = cell0 + tid;
  const bool active = cell < P.n_cells;               // (all threads stay: the block synchronises once per point)
  const int64_t cc = active ? cell : P.n_cells - 1;   // idle threads of the last block repeat its last cell and write nothing
  const double *geo_blk = P.geo + blk_index(cell0, 0, NN * G);             // the block's contiguous chunk: point q at + q * 640
  constexpr uint32_t STAGE_BYTES = G * CV_THREADS * sizeof(double);
  const auto issue = [&](int q) {                     // (thread 0) arm the stage's barrier and start the bulk copy of point q
    const int st = q % CV_STAGES;
    cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared, &bar[st], STAGE_BYTES);
    cuda::ptx::cp_async_bulk(cuda::ptx::space_cluster, cuda::ptx::space_global, stage + st * G * CV_THREADS,
                             geo_blk + (int64_t)q * G * CV_THREADS, STAGE_BYTES, &bar[st]);
  };
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < CV_STAGES; ++st) cuda::ptx::mbarrier_init(&bar[st], 1);
    cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);                // the initialised barriers are visible to the copy engine
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < CV_STAGES; ++q) issue(q);
  }
  const uint32_t *nodes_u = P.cell_nodes_u_t + blk_index(cc, 0, NN);     // transposed maps: node i at nodes_u[i * GEO_BLOCK]
  const uint32_t *nodes_p = P.cell_nodes_p_t + blk_index(cc, 0, NP);
  if (MODE != 2) {
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      double v[NN];
#pragma unroll
      for (int i = 0; i < NN; ++i) {
        const uint32_t e = nodes_u[i * GEO_BLOCK];
        const bool cons = ((e >> (NODE_FLAG_SHIFT + c)) & 1u) && !P.plain;
        v[i] = cons ? 0.0 : P.src_u[(e & NODE_IDX_MASK) + c];      // (idle threads: the last cell's DoFs, valid addresses)
      }
      interp_all<3>(v);
#pragma unroll
      for (int q = 0; q < NN; ++q) U[(c * NN + q) * CV_THREADS] = v[q];
    }
  }
  // (volatile: every use below is a fresh 8-byte shared-memory load -- otherwise the compiler keeps all 81 values in
  //  registers next to the 81 accumulators and spills; the loads are conflict-free and cheap)
  const volatile double *Uv = U;
  double pn[NP], Tp[NP];
#pragma unroll
  for (int a = 0; a < NP; ++a) {
    pn[a] = 0.0; Tp[a] = 0.0;
    if (MODE != 0) {
      const uint32_t e = nodes_p[a * GEO_BLOCK];
      pn[a] = ((e >> 31) && !P.plain) ? 0.0 : P.src_p[e & NODE_IDX_MASK];
    }
  }
  double T[3][NN];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int i = 0; i < NN; ++i) T[c][i] = 0.0;
  const double *etab = P.eta + (cc / SUB) * P.n_eta * SUB + (cc % SUB);
  const double eta_cell = (MODE != 2 && !PERQ) ? etab[0] : 1.0;
#pragma unroll
  for (int q = 0; q < NN; ++q) {
    const int qd[3] = {q % 3, (q / 3) % 3, q / 9};
    // wait for the point's geometry (stage q % CV_STAGES, its (q / CV_STAGES)-th use), take this thread's 10 values, hand the
    // stage back: the copies of the next CV_STAGES - 1 points are in flight meanwhile
    {
      const uint32_t parity = (q / CV_STAGES) & 1;
      while (!cuda::ptx::mbarrier_try_wait_parity(&bar[q % CV_STAGES], parity)) {}
    }
    const double *sg = stage + (q % CV_STAGES) * G * CV_THREADS + tid;
    double Ji[3][3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Ji[k / 3][k % 3] = sg[k * CV_THREADS];
    const double JxW = sg[9 * CV_THREADS];
    __syncthreads();
    if (tid == 0 && q + CV_STAGES < NN) issue(q + CV_STAGES);
    double S[3][3];                                   // stress times JxW
    double div = 0.0;
    if (MODE != 2) {
      double Gx[3][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double u0 = Uv[(c * NN + q) * CV_THREADS];
        double gh[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int st = stride3(r), base = q - qd[r] * st;
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (dpc(qd[r], j) != 0.0) acc = fma(dpc(qd[r], j), j == qd[r] ? u0 : Uv[(c * NN + base + j * st) * CV_THREADS], acc);
          gh[r] = acc;
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) Gx[c][d] = gh[0] * Ji[0][d] + gh[1] * Ji[1][d] + gh[2] * Ji[2][d];
      }
      div = Gx[0][0] + Gx[1][1] + Gx[2][2];
      const double ew = (PERQ ? etab[q * SUB] : eta_cell) * JxW;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d) S[c][d] = ew * (Gx[c][d] + Gx[d][c]);
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d) S[c][d] = 0.0;
    }
    if (MODE != 0) {
      double valp = 0.0;
#pragma unroll
      for (int a = 0; a < NP; ++a) valp = fma(pn[a], q1v(qd[0], a & 1) * q1v(qd[1], (a >> 1) & 1) * q1v(qd[2], a >> 2), valp);
      const double pw = P.ps * valp * JxW;
      S[0][0] -= pw; S[1][1] -= pw; S[2][2] -= pw;
      if (MODE == 1) {
        const double pt = -P.ps * div * JxW;
#pragma unroll
        for (int a = 0; a < NP; ++a) Tp[a] = fma(pt, q1v(qd[0], a & 1) * q1v(qd[1], (a >> 1) & 1) * q1v(qd[2], a >> 2), Tp[a]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double sh = S[c][0] * Ji[r][0] + S[c][1] * Ji[r][1] + S[c][2] * Ji[r][2];
        const int st = stride3(r), base = q - qd[r] * st;
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (dpc(qd[r], j) != 0.0) T[c][base + j * st] = fma(dpc(qd[r], j), sh, T[c][base + j * st]);
      }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    interpT_all<3>(T[c]);
#pragma unroll
    for (int i = 0; i < NN; ++i) {
      const uint32_t e = nodes_u[i * GEO_BLOCK];
      if (active && !((e >> (NODE_FLAG_SHIFT + c)) & 1u)) atomicAdd(P.dst_u + (e & NODE_IDX_MASK) + c, T[c][i]);
    }
  }
  if (MODE == 1) {
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      const uint32_t e = nodes_p[a * GEO_BLOCK];
      if (active && !(e >> 31)) atomicAdd(P.dst_p + (e & NODE_IDX_MASK), Tp[a]);
    }
  }
}

// pressure mass matrix on curved cells (or any mesh without batch structure): one thread per cell
template <int DIM, bool PERQ>
__global__ void __launch_bounds__(128) curved_mass_kernel(const GeneralParams P)
{
  constexpr int NN = nn<DIM>(), NP = np<DIM>(), G = DIM * DIM + 1;
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= P.n_cells) return;
  double pn[NP], pq[NN], out[NP];
#pragma unroll
  for (int a = 0; a < NP; ++a) {
    const uint32_t e = P.cell_nodes_p[cell * NP + a];
    pn[a] = (e >> 31) ? 0.0 : P.src_p[e & NODE_IDX_MASK];
  }
  q1_to_quad<DIM>(pn, pq);
  const double *etab = P.eta + (cell / SUB) * P.n_eta * SUB + (cell % SUB);
  const double ps2 = P.ps * P.ps;
#pragma unroll
  for (int q = 0; q < NN; ++q) {
    const double JxW = P.geo ? P.geo[blk_index(cell, q * G + DIM * DIM, NN * G)] : qweight<DIM>(q) * P.vol;
    pq[q] *= JxW * (ps2 / etab[PERQ ? q * SUB : 0]);
  }
  quad_to_q1T<DIM>(pq, out);
#pragma unroll
  for (int a = 0; a < NP; ++a) {
    const uint32_t e = P.cell_nodes_p[cell * NP + a];
    if (!(e >> 31)) atomicAdd(P.dst_p + (e & NODE_IDX_MASK), out[a]);
  }
}
// dst[c] = src[c] on constrained DoFs (unit diagonal of MatrixFreeOperators::Base)
__global__ void copy_entries_kernel(double *dst, const double *src, const uint32_t *idx, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = src[idx[i]];
}

static bool needs_general(const LevelDev &L, int mode)
{ return L.geo || L.compressible || L.nd.dilation || (mode == 1 && L.nd.newton); }

// dst must be zeroed by the caller
static int launch_general(smf_ctx *ctx, LevelDev &L, int mode, double *dst_u, double *dst_p, const double *src_u, const double *src_p, bool plain)
{
  GeneralParams P{};
  P.n_cells = L.n_cells; P.cell_nodes_u = L.cell_nodes_u; P.cell_nodes_p = L.cell_nodes_p;
  P.src_u = src_u; P.src_p = src_p; P.dst_u = dst_u; P.dst_p = dst_p;
  P.eta = L.eta; P.n_eta = L.n_eta;
  double vol = 1.0;
  for (int d = 0; d < L.dim; ++d) { P.ih[d] = 1.0 / L.h[d]; vol *= L.h[d]; }
  P.vol = vol; P.ps = L.ps; P.mode = mode; P.plain = plain ? 1 : 0; P.geo = L.geo;
  P.compressible = L.compressible; P.newton = L.nd.newton; P.symmetrize = L.nd.symmetrize; P.average = L.nd.average; P.dilation = L.nd.dilation;
  P.eps = L.nd.eps; P.deta = L.nd.deta; P.detadp = L.nd.detadp; P.dil_lhs = L.nd.dil_lhs; P.ddil = L.nd.ddil; P.ddildp = L.nd.ddildp;
  const bool fine = (&L == &ctx->lev(SMF_ACTIVE)) || (&L == &ctx->levels[ctx->cfg.n_levels - 1]);
  // curved cells without Newton / compressible / dilation terms: the sum-factorised kernel
  const bool fast = L.geo && L.dim == 3 && ctx->use_curved_fast && !L.compressible && !L.nd.dilation && !(mode == 1 && L.nd.newton);
  {
    ProfScope prof(ctx, mode == 1 ? B_STOKES : (mode == 2 ? B_BT : (fine ? B_A_FINE : B_A_COARSE)));
    if (fast) {
      const unsigned grid = (unsigned)((L.n_cells + CV_THREADS - 1) / CV_THREADS);
      const size_t smem = sizeof(double) * 81 * CV_THREADS;
      const bool perq = L.n_eta != 1;
      P.cell_nodes_u_t = L.cell_nodes_u_t; P.cell_nodes_p_t = L.cell_nodes_p_t;
      if (ctx->use_curved_tma) {
        const size_t smem_tma = smem + sizeof(double) * CV_STAGES * 10 * CV_THREADS;
        static bool configured_dev[MAX_DEVICES] = {false};
        bool &configured = configured_dev[ctx->cfg.device % MAX_DEVICES];
        if (!configured) {
          SMF_CUDA(cudaFuncSetAttribute(curved_cell_kernel_tma<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
          SMF_CUDA(cudaFuncSetAttribute(curved_cell_kernel_tma<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
          SMF_CUDA(cudaFuncSetAttribute(curved_cell_kernel_tma<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
          SMF_CUDA(cudaFuncSetAttribute(curved_cell_kernel_tma<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
          SMF_CUDA(cudaFuncSetAttribute(curved_cell_kernel_tma<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
          configured = true;
        }
        if (mode == 0) { if (perq) curved_cell_kernel_tma<0, true><<<grid, CV_THREADS, smem_tma, ctx->stream>>>(P); else curved_cell_kernel_tma<0, false><<<grid, CV_THREADS, smem_tma, ctx->stream>>>(P); }
        else if (mode == 1) { if (perq) curved_cell_kernel_tma<1, true><<<grid, CV_THREADS, smem_tma, ctx->stream>>>(P); else curved_cell_kernel_tma<1, false><<<grid, CV_THREADS, smem_tma, ctx->stream>>>(P); }
        else curved_cell_kernel_tma<2, false><<<grid, CV_THREADS, smem_tma, ctx->stream>>>(P);
      } else
      if (mode == 0) { if (perq) curved_cell_kernel<0, true><<<grid, CV_THREADS, smem, ctx->stream>>>(P); else curved_cell_kernel<0, false><<<grid, CV_THREADS, smem, ctx->stream>>>(P); }
      else if (mode == 1) { if (perq) curved_cell_kernel<1, true><<<grid, CV_THREADS, smem, ctx->stream>>>(P); else curved_cell_kernel<1, false><<<grid, CV_THREADS, smem, ctx->stream>>>(P); }
      else curved_cell_kernel<2, false><<<grid, CV_THREADS, smem, ctx->stream>>>(P);
    } else {
      const unsigned grid = (unsigned)((L.n_cells + 63) / 64);
      if (L.dim == 3) general_cell_kernel<3><<<grid, 64, 0, ctx->stream>>>(P);
      else general_cell_kernel<2><<<grid, 64, 0, ctx->stream>>>(P);
    }
  }
  ctx->launches++;
  if (!plain && mode != 2) {
    if (L.n_constrained_u) {
      copy_entries_kernel<<<(unsigned)((L.n_constrained_u + 255) / 256), 256, 0, ctx->stream>>>(dst_u, src_u, L.constrained_u, L.n_constrained_u);
      ctx->launches++;
    }
    if (mode == 1 && L.n_constrained_p) {
      copy_entries_kernel<<<(unsigned)((L.n_constrained_p + 255) / 256), 256, 0, ctx->stream>>>(dst_p, src_p, L.constrained_p, L.n_constrained_p);
      ctx->launches++;
    }
  }
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}


// ---- Stokes right-hand side on the device -------------------------------------------------------
// f_i = sum_q force_q . phi_i JxW_q with force = density * gravity from the plugins (assemblers/stokes.cc:387-396),
// one thread per (cell, node); constrained rows receive nothing (the reference condenses them away, assembly.cc:761-889)
template <int DIM> __global__ void rhs_kernel(const double *force_q, const uint32_t *cell_nodes_u, double *rhs_u, int64_t n_cells, double vol, const double *geo)
{
  constexpr int NN = nn<DIM>();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t cell = t / NN;
  const int i = (int)(t % NN);
  if (cell >= n_cells) return;
  int id[3] = {i % 3, (i / 3) % 3, DIM == 3 ? i / 9 : 0};
  double acc[DIM];
  for (int c = 0; c < DIM; ++c) acc[c] = 0.0;
  for (int q = 0; q < NN; ++q) {
    int qd[3] = {q % 3, (q / 3) % 3, DIM == 3 ? q / 9 : 0};
    double phi = geo ? geo[blk_index(cell, q * (DIM * DIM + 1) + DIM * DIM, NN * (DIM * DIM + 1))] : qweight<DIM>(q) * vol;
    for (int d = 0; d < DIM; ++d) phi *= q2v(qd[d], id[d]);
    const double *f = force_q + (cell * NN + q) * DIM;
    for (int c = 0; c < DIM; ++c) acc[c] = fma(f[c], phi, acc[c]);
  }
  const uint32_t e = cell_nodes_u[cell * NN + i];
  const uint32_t idx = e & NODE_IDX_MASK;
  for (int c = 0; c < DIM; ++c)
    if (!((e >> (NODE_FLAG_SHIFT + c)) & 1u) && acc[c] != 0.0) atomicAdd(rhs_u + idx + c, acc[c]);
}

int launch_rhs(smf_ctx *ctx, LevelDev &L, const double *force_q_dev, double *rhs_u)
{
  double vol = 1.0;
  for (int d = 0; d < L.dim; ++d) vol *= L.h[d];
  const int NN = L.dim == 3 ? 27 : 9;
  const int64_t nt = L.n_cells * NN;
  SMF_CUDA(cudaMemsetAsync(rhs_u, 0, sizeof(double) * L.n_u, ctx->stream));
  if (L.dim == 3) rhs_kernel<3><<<(unsigned)((nt + 127) / 128), 128, 0, ctx->stream>>>(force_q_dev, L.cell_nodes_u, rhs_u, L.n_cells, vol, L.geo);
  else rhs_kernel<2><<<(unsigned)((nt + 127) / 128), 128, 0, ctx->stream>>>(force_q_dev, L.cell_nodes_u, rhs_u, L.n_cells, vol, L.geo);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return exchange_add(ctx, L.ex_u, rhs_u);        // compress(add) of the per-rank partial sums
}

static int check_level(smf_ctx *ctx, const LevelDev &L)
{
  if (L.n_cells == 0) return ctx->fail(SMF_ERR_INVALID, "level mesh not set (smf_set_level)");
  if (!L.has_cell_data) return ctx->fail(SMF_ERR_INVALID, "cell data not set (smf_set_cell_data)");
  if (!L.safe_rounds && !L.geo) return ctx->fail(SMF_ERR_UNSUPPORTED, "cell batches with aliasing scatter rounds (unstructured orientation) need geometry = 1 (the per-cell kernels)");
  return SMF_OK;
}

// vec[idx[i]] = value (AffineConstraints::set_zero on a device vector)
int launch_set_entries(smf_ctx *ctx, double *vec, const uint32_t *idx, int64_t n, double value)
{
  if (n <= 0) return SMF_OK;
  set_entries_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(vec, idx, n, value);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

// A block through the lattice mirror: dst (zero on entry, host numbering) = A src on the unconstrained rows
int launch_velocity_lat(smf_ctx *ctx, LevelDev &L, double *dst, const double *src)
{
  VLat &V = *L.vlat;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return ctx->fail(SMF_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap tmap;
  const cuuint64_t dims[4] = {(cuuint64_t)V.Xp, (cuuint64_t)V.ny, (cuuint64_t)V.nz, 6};
  const cuuint64_t strides[3] = {(cuuint64_t)V.Xp * 8, (cuuint64_t)V.Xp * V.ny * 8, (cuuint64_t)V.Xp * V.ny * V.nz * 8};
  const cuuint32_t box[4] = {VL_ROW, VL_ROWS, 5, 1}, estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, V.src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ctx->fail(SMF_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  const unsigned nb = (unsigned)((V.n_nodes + 255) / 256);
  {
    ProfScope prof(ctx, B_VEC);
    vl_to_lattice<<<nb, 256, 0, ctx->stream>>>(V.src, src, V.lat2dof, V.nx, V.ny, V.nz, V.Xp);
    SMF_CUDA(cudaMemsetAsync(V.dst, 0, sizeof(double) * V.n_lat, ctx->stream));
    ctx->launches++;
  }
  CellParams P = make_params(L);
  VLParams Q{V.brick_org, V.dst, V.Xp, V.ny, V.nz, V.cons_mask, 0};
  constexpr size_t smem = vl_smem_bytes();
  static int resident_dev[MAX_DEVICES] = {0};
  int &resident = resident_dev[ctx->cfg.device % MAX_DEVICES];
  if (resident == 0) {
    SMF_CUDA(cudaFuncSetAttribute(velocity_lat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, sms = 148;
    SMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->cfg.device));
    SMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, velocity_lat_kernel<false>, SBK_THREADS, smem));
    if (per_sm < 1) return ctx->fail(SMF_ERR_CUDA, "velocity_lat_kernel does not fit on an SM");
    resident = per_sm * sms;
  }
  const bool fine = (&L == &ctx->lev(SMF_ACTIVE)) || (&L == &ctx->levels[ctx->cfg.n_levels - 1]);
  {
    ProfScope prof(ctx, fine ? B_A_FINE : B_A_COARSE);
    int grid = std::min(L.n_batches, resident);
    if (ctx->max_grid_blocks > 0) grid = std::min(grid, ctx->max_grid_blocks);
    velocity_lat_kernel<false><<<grid, SBK_THREADS, smem, ctx->stream>>>(tmap, P, Q, L.n_batches);
    ctx->launches++;
  }
  {
    ProfScope prof(ctx, B_VEC);
    vl_from_lattice<<<nb, 256, 0, ctx->stream>>>(dst, V.dst, V.lat2dof, V.nx, V.ny, V.nz, V.Xp);
    ctx->launches++;
    if (L.n_constrained_u) {      // unit diagonal on the constrained rows: dst = src
      copy_entries_kernel<<<(unsigned)((L.n_constrained_u + 255) / 256), 256, 0, ctx->stream>>>(dst, src, L.constrained_u, L.n_constrained_u);
      ctx->launches++;
    }
  }
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}

int launch_vmult_A(smf_ctx *ctx, LevelDev &L, double *dst, const double *src, bool dst_is_zero)
{
  if (int e = check_level(ctx, L)) return e;
  CellParams P = make_params(L);
  P.src_u = src; P.dst_u = dst;
  if (!dst_is_zero) SMF_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * L.n_u, ctx->stream));
  if (int e = hanging_pre(ctx, L.hang_u, src)) return e;              // read_dof_values resolves the hanging-node lines
  if (ctx->use_vlat && !L.vlat && L.structured && L.has_cell_data) { if (int e = vlat_prepare(ctx, L)) return e; }
  if (needs_general(L, 0)) { if (int e = launch_general(ctx, L, 0, dst, nullptr, src, nullptr, false)) return e; }
  else if (ctx->use_vlat && L.vlat && L.n_eta == 1) { if (int e = launch_velocity_lat(ctx, L, dst, src)) return e; }      // experimental: TMA-staged bricks
  else if (int e = L.dim == 3 ? launch_velocity<3, 0>(ctx, L, P) : launch_velocity<2, 0>(ctx, L, P)) return e;
  if (int e = hanging_post(ctx, L.hang_u, dst, src, true)) return e;  // distribute_local_to_global + unit diagonal
  return exchange_add(ctx, L.ex_u, dst);
}
int launch_vmult_stokes(smf_ctx *ctx, LevelDev &L, double *dst_u, double *dst_p, const double *src_u, const double *src_p, bool plain)
{
  if (int e = check_level(ctx, L)) return e;
  CellParams P = make_params(L);
  P.src_u = src_u; P.src_p = src_p; P.dst_u = dst_u; P.dst_p = dst_p;
  P.plain = plain ? 1 : 0;
  SMF_CUDA(cudaMemsetAsync(dst_u, 0, sizeof(double) * L.n_u, ctx->stream));
  SMF_CUDA(cudaMemsetAsync(dst_p, 0, sizeof(double) * L.n_p, ctx->stream));
  if (plain && (L.hang_u.n || L.hang_p.n))
    return ctx->fail(SMF_ERR_UNSUPPORTED, "correct_stokes_rhs (read_dof_values_plain) on a mesh with hanging nodes: inhomogeneous hanging-node constraints are not implemented");
  if (int e = hanging_pre(ctx, L.hang_u, src_u)) return e;
  if (int e = hanging_pre(ctx, L.hang_p, src_p)) return e;
  if (needs_general(L, 1)) { if (int e = launch_general(ctx, L, 1, dst_u, dst_p, src_u, src_p, plain)) return e; }
  else if (int e = L.dim == 3 ? launch_velocity<3, 1>(ctx, L, P) : launch_velocity<2, 1>(ctx, L, P)) return e;
  if (int e = free_surface_faces_apply(ctx, L, dst_u, src_u, plain)) return e;      // local_apply_boundary_face (:563-602)
  if (int e = hanging_post(ctx, L.hang_u, dst_u, src_u, true)) return e;
  if (int e = hanging_post(ctx, L.hang_p, dst_p, src_p, true)) return e;
  if (int e = exchange_add(ctx, L.ex_u, dst_u)) return e;
  return exchange_add(ctx, L.ex_p, dst_p);
}
int launch_vmult_BT(smf_ctx *ctx, LevelDev &L, double *dst_u, const double *src_p)
{
  if (int e = check_level(ctx, L)) return e;
  CellParams P = make_params(L);
  P.src_p = src_p; P.dst_u = dst_u;
  SMF_CUDA(cudaMemsetAsync(dst_u, 0, sizeof(double) * L.n_u, ctx->stream));
  if (int e = hanging_pre(ctx, L.hang_p, src_p)) return e;
  if (L.geo) { if (int e = launch_general(ctx, L, 2, dst_u, nullptr, nullptr, src_p, false)) return e; }
  else if (int e = L.dim == 3 ? launch_velocity<3, 2>(ctx, L, P) : launch_velocity<2, 2>(ctx, L, P)) return e;
  if (int e = hanging_restore(ctx, L.hang_p, src_p)) return e;
  if (int e = hanging_post(ctx, L.hang_u, dst_u, nullptr, false)) return e;      // (the velocity source block of B^T is zero)
  return exchange_add(ctx, L.ex_u, dst_u);
}
int launch_vmult_mass(smf_ctx *ctx, LevelDev &L, double *dst, const double *src, bool dst_is_zero)
{
  if (int e = check_level(ctx, L)) return e;
  CellParams P = make_params(L);
  P.src_p = src; P.dst_p = dst;
  if (!dst_is_zero) SMF_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * L.n_p, ctx->stream));
  if (int e = hanging_pre(ctx, L.hang_p, src)) return e;
  const size_t smem = (size_t)MASS_WARPS * 2 * L.max_nodes_p * sizeof(double);
  const unsigned grid = (unsigned)((L.n_batches + MASS_WARPS - 1) / MASS_WARPS);
  if (L.geo) {
    ProfScope prof(ctx, B_MASS);
    GeneralParams Q{};
    Q.n_cells = L.n_cells; Q.cell_nodes_p = L.cell_nodes_p; Q.src_p = src; Q.dst_p = dst; Q.eta = L.eta; Q.n_eta = L.n_eta; Q.ps = L.ps; Q.geo = L.geo;
    const unsigned g = (unsigned)((L.n_cells + 127) / 128);
    if (L.dim == 3) { if (L.n_eta == 1) curved_mass_kernel<3, false><<<g, 128, 0, ctx->stream>>>(Q); else curved_mass_kernel<3, true><<<g, 128, 0, ctx->stream>>>(Q); }
    else { if (L.n_eta == 1) curved_mass_kernel<2, false><<<g, 128, 0, ctx->stream>>>(Q); else curved_mass_kernel<2, true><<<g, 128, 0, ctx->stream>>>(Q); }
    if (L.n_constrained_p) {
      copy_entries_kernel<<<(unsigned)((L.n_constrained_p + 255) / 256), 256, 0, ctx->stream>>>(dst, src, L.constrained_p, L.n_constrained_p);
      ctx->launches++;
    }
  } else if (L.dim == 3 && L.structured && L.sb_lists && ctx->use_pipelined) {
    ProfScope prof(ctx, B_MASS);
    const unsigned g = (unsigned)((L.n_batches + MSB_WARPS - 1) / MSB_WARPS);
    if (L.n_eta == 1) mass_sb_kernel<false><<<g, 32 * MSB_WARPS, 0, ctx->stream>>>(P, L.sb_lists, L.n_batches);
    else mass_sb_kernel<true><<<g, 32 * MSB_WARPS, 0, ctx->stream>>>(P, L.sb_lists, L.n_batches);
  } else {
    ProfScope prof(ctx, B_MASS);
    if (L.dim == 3) {
      if (L.n_eta == 1) mass_kernel<3, false><<<grid, 32 * MASS_WARPS, smem, ctx->stream>>>(P, L.n_batches);
      else mass_kernel<3, true><<<grid, 32 * MASS_WARPS, smem, ctx->stream>>>(P, L.n_batches);
    } else {
      if (L.n_eta == 1) mass_kernel<2, false><<<grid, 32 * MASS_WARPS, smem, ctx->stream>>>(P, L.n_batches);
      else mass_kernel<2, true><<<grid, 32 * MASS_WARPS, smem, ctx->stream>>>(P, L.n_batches);
    }
  }
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  if (int e = hanging_post(ctx, L.hang_p, dst, src, true)) return e;
  return exchange_add(ctx, L.ex_p, dst);
}

int launch_diag(smf_ctx *ctx, LevelDev &L)
{
  DiagParams P{};
  P.n_cells = L.n_cells; P.cell_nodes_u = L.cell_nodes_u; P.cell_nodes_p = L.cell_nodes_p;
  P.eta = L.eta; P.n_eta = L.n_eta;
  double vol = 1.0;
  for (int d = 0; d < L.dim; ++d) { P.ih[d] = 1.0 / L.h[d]; vol *= L.h[d]; }
  P.vol = vol; P.ps = L.ps; P.compressible = L.compressible || L.nd.dilation;   // deviatoric A block (:999-1005)
  P.diag_A = L.dinv_A; P.diag_M = L.dinv_M; P.geo = L.geo;
  SMF_CUDA(cudaMemsetAsync(L.dinv_A, 0, sizeof(double) * L.n_u, ctx->stream));
  SMF_CUDA(cudaMemsetAsync(L.dinv_M, 0, sizeof(double) * L.n_p, ctx->stream));
  const int NN = L.dim == 3 ? 27 : 9;
  const int64_t nt = L.n_cells * NN;
  const int bs = 128;
  if (L.dim == 3) diag_kernel<3><<<(unsigned)((nt + bs - 1) / bs), bs, 0, ctx->stream>>>(P);
  else diag_kernel<2><<<(unsigned)((nt + bs - 1) / bs), bs, 0, ctx->stream>>>(P);
  ctx->launches++;
  // compress(add) of the per-rank partial diagonals
  if (int e = exchange_add(ctx, L.ex_u, L.dinv_A)) return e;
  if (int e = exchange_add(ctx, L.ex_p, L.dinv_M)) return e;
  // set_constrained_entries_to_one, then invert (matrix_free_operators.cc:1085-1097)
  if (L.n_constrained_u) {
    set_entries_kernel<<<(unsigned)((L.n_constrained_u + 255) / 256), 256, 0, ctx->stream>>>(L.dinv_A, L.constrained_u, L.n_constrained_u, 1.0);
    ctx->launches++;
  }
  if (L.n_constrained_p) {
    set_entries_kernel<<<(unsigned)((L.n_constrained_p + 255) / 256), 256, 0, ctx->stream>>>(L.dinv_M, L.constrained_p, L.n_constrained_p, 1.0);
    ctx->launches++;
  }
  // refinement-edge rows of a level operator and hanging rows of the active operator carry a unit diagonal, too
  // (A_l = (I - E) A_bc (I - E) + E; the diagonal of C^T A C on the masters of hanging nodes is not corrected: the active
  // diagonal is not used by the solve)
  if (L.edge_u.n) { set_entries_kernel<<<(unsigned)((L.edge_u.n + 255) / 256), 256, 0, ctx->stream>>>(L.dinv_A, L.edge_u.idx, L.edge_u.n, 1.0); ctx->launches++; }
  if (L.edge_p.n) { set_entries_kernel<<<(unsigned)((L.edge_p.n + 255) / 256), 256, 0, ctx->stream>>>(L.dinv_M, L.edge_p.idx, L.edge_p.n, 1.0); ctx->launches++; }
  if (L.hang_u.n) { set_entries_kernel<<<(unsigned)((L.hang_u.n + 255) / 256), 256, 0, ctx->stream>>>(L.dinv_A, L.hang_u.dof, L.hang_u.n, 1.0); ctx->launches++; }
  if (L.hang_p.n) { set_entries_kernel<<<(unsigned)((L.hang_p.n + 255) / 256), 256, 0, ctx->stream>>>(L.dinv_M, L.hang_p.dof, L.hang_p.n, 1.0); ctx->launches++; }
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
  const unsigned grid = (unsigned)((P.n_parents + TR_WARPS - 1) / TR_WARPS);
  ProfScope prof(ctx, B_TRANSFER);
  if (prolong) prolongate_kernel<DIM, NPD, NCOMP><<<grid, 32 * TR_WARPS, 0, ctx->stream>>>(P);
  else restrict_kernel<DIM, NPD, NCOMP><<<grid, 32 * TR_WARPS, 0, ctx->stream>>>(P);
  ctx->launches++;
}

static bool transfer_is_brickwise(smf_ctx *ctx, int which, int fine_level)
{
  return which == 0 && ctx->cfg.dim == 3 && fine_level >= 1 && fine_level < ctx->cfg.n_levels && ctx->transfers[fine_level].tb_u &&
         ctx->levels[fine_level].sb_lists && ctx->use_pipelined;
}
static int launch_transfer(smf_ctx *ctx, int which, int fine_level, double *dst, const double *src, bool prolong, bool add, const double *minus = nullptr)
{
  if (fine_level < 1 || fine_level >= ctx->cfg.n_levels || !ctx->transfers[fine_level].set)
    return ctx->fail(SMF_ERR_INVALID, "transfer not set for this level (smf_set_transfer)");
  TransferDev &T = ctx->transfers[fine_level];
  LevelDev &C = ctx->levels[fine_level - 1];
  TransferParams P{};
  P.n_parents = T.n_parents; P.src = src; P.dst = dst; P.add = add ? 1 : 0;
  const int dim = ctx->cfg.dim;
  LevelDev &F = ctx->levels[fine_level];
  if (which == 0 && dim == 3 && T.tb_u && F.sb_lists && ctx->use_pipelined) {
    TransferSbParams Q{T.tb_u, F.sb_lists, src, dst, add ? 1 : 0, prolong ? nullptr : minus};
    ProfScope prof(ctx, B_TRANSFER);
    if (prolong) prolongate_sb_kernel<<<F.n_batches, TSB_THREADS, 0, ctx->stream>>>(Q);
    else restrict_sb_kernel<<<F.n_batches, TSB_THREADS, 0, ctx->stream>>>(Q);
    ctx->launches++;
  } else if (which == 0) {
    P.patch = T.patch_u; P.coarse_nodes = C.cell_nodes_u;
    if (dim == 3) launch_transfer_t<3, 3, 3>(ctx, P, prolong); else launch_transfer_t<2, 3, 2>(ctx, P, prolong);
  } else {
    P.patch = T.patch_p; P.coarse_nodes = C.cell_nodes_p;
    if (dim == 3) launch_transfer_t<3, 2, 1>(ctx, P, prolong); else launch_transfer_t<2, 2, 1>(ctx, P, prolong);
  }
  SMF_CUDA(cudaGetLastError());
  // restriction: the coarse vector holds per-rank partial sums on the subdomain interfaces
  if (!prolong) return exchange_add(ctx, which == 0 ? C.ex_u : C.ex_p, dst);
  return SMF_OK;
}

int launch_prolongate(smf_ctx *ctx, int which, int to_level, double *dst, const double *src, bool add)
{
  // constraint lines on the COARSE level (no-normal-flux on a curved boundary: user constraints of MGConstrainedDoFs):
  // MGTransferMatrixFree::prolongate distributes them on its source vector before it interpolates; the source comes back
  // unchanged (the entries are parked and restored)
  ConstraintLines *H = nullptr;
  if (to_level >= 1 && to_level < ctx->cfg.n_levels) {
    LevelDev &C = ctx->levels[to_level - 1];
    H = which == 0 ? &C.hang_u : &C.hang_p;
    if (!H->n) H = nullptr;
  }
  if (H) { if (int e = hanging_pre(ctx, *H, src)) return e; }
  if (int e = launch_transfer(ctx, which, to_level, dst, src, true, add)) return e;
  if (H) return hanging_restore(ctx, *H, src);
  return SMF_OK;
}
int launch_restrict_add(smf_ctx *ctx, int which, int from_level, double *dst, const double *src)
{ return launch_transfer(ctx, which, from_level, dst, src, false, true); }
// dst += R (b - t): the residual of the V-cycle restricted without being stored (brick-wise velocity transfers); elsewhere
// t <- b - t first.  Returns through *fused whether t was left untouched.
int launch_restrict_add_residual(smf_ctx *ctx, int which, int from_level, double *dst, const double *b, double *t, int64_t n_fine)
{
  if (transfer_is_brickwise(ctx, which, from_level)) return launch_transfer(ctx, which, from_level, dst, b, false, true, t);
  if (int e = vec_residual(ctx, t, b, n_fine)) return e;
  return launch_transfer(ctx, which, from_level, dst, t, false, true);
}

}  // namespace smf
