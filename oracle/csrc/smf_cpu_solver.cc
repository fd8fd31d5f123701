/* CPU implementation of the reference's block-GMG Stokes solve (ORACLE / TEST INFRASTRUCTURE ONLY).
 *
 * This is the "one thread per core, SIMD cell batches, whole solver loop in C++" restatement BASELINE.md section 3 asks
 * for: it is what bench.py times as `--impl reference` / `cpu_baseline` (kind "port": deal.II / ASPECT cannot be built
 * in this image) and what the GPU parity tests compare with at sizes the numpy-driven oracle (oracle/solver.py) cannot
 * reach (BASELINE cfg 2 = nsinker refinement 5, cfg 1 = SolCx refinement 6).  Nothing under aspect_b200/ links or loads it.
 *
 * What it follows (all paths relative to /root/reference):
 *   cell kernels      source/simulator/solver/matrix_free_operators.cc  StokesOperator::local_apply :376-547 (no Newton
 *                     terms), BTBlockOperator :668-701, MassMatrixOperator :787-845, ABlockOperator :974-1052,
 *                     compute_diagonal :861-908, :1068-1098 -- executed the way deal.II's FEEvaluation does: sum
 *                     factorisation with the cells of a SIMD batch in the lanes of a vector (VectorizedArray<double>),
 *                     cell batches distributed over the threads by a block colouring (MatrixFree partition_color).
 *   solve()           source/simulator/solver/stokes_matrix_free_local_smoothing.cc :758-1313: Chebyshev smoothers
 *                     :774-825, eigenvalue estimation :834-849, coarse solver = smoother :855-860, V-cycles :893-916,
 *                     solver tolerance :977-1041, right-preconditioned GMRES :1277-1301.
 *   preconditioner    include/aspect/simulator/solver/block_stokes_preconditioner.h :220-255 (cheap variant).
 *   deal.II pieces    PreconditionChebyshev, Multigrid/PreconditionMG, MGTransferMatrixFree, SolverGMRES as recalled in
 *                     SURVEY.md App. A and restated line by line in oracle/solver.py, which this file repeats in C++
 *                     (tests/test_cpuref.py checks both against each other and against the reference's goldens).
 *
 * Cartesian cells only (per-level cell size h[dim]); Q2 velocity / Q1 pressure; 3-point Gauss rule per direction.
 * Build: oracle/Makefile, twice (SMFC_W = 4 for AVX2, 8 for AVX-512); oracle/cpuref.py loads the one the host supports.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef SMFC_W
#define SMFC_W 4
#endif
namespace {
constexpr int W = SMFC_W;
typedef double vd __attribute__((vector_size(8 * SMFC_W)));

double N2[3][3], D2c[3][3], N1[3][2], GW[3];
bool tables_ready = false;
const double P2[5][3] = {{1, 0, 0}, {0.375, 0.75, -0.125}, {0, 1, 0}, {-0.125, 0.75, 0.375}, {0, 0, 1}};
const double P1[3][2] = {{1, 0}, {0.5, 0.5}, {0, 1}};

void init_tables()
{
  if (tables_ready) return;
  const double s = std::sqrt(3.0 / 5.0);
  const double x[3] = {0.5 - 0.5 * s, 0.5, 0.5 + 0.5 * s};
  GW[0] = 5.0 / 18.0; GW[1] = 8.0 / 18.0; GW[2] = 5.0 / 18.0;
  for (int q = 0; q < 3; ++q) {
    N2[q][0] = (1 - x[q]) * (1 - 2 * x[q]); N2[q][1] = 4 * x[q] * (1 - x[q]); N2[q][2] = x[q] * (2 * x[q] - 1);
    N1[q][0] = 1 - x[q]; N1[q][1] = x[q];
  }
  for (int q = 0; q < 3; ++q)          /* derivative at Gauss point q of the Lagrange polynomial through Gauss point j */
    for (int j = 0; j < 3; ++j) {
      double sum = 0;
      for (int m = 0; m < 3; ++m) {
        if (m == j) continue;
        double prod = 1.0 / (x[j] - x[m]);
        for (int l = 0; l < 3; ++l)
          if (l != j && l != m) prod *= (x[q] - x[l]) / (x[j] - x[l]);
        sum += prod;
      }
      D2c[q][j] = sum;
    }
  tables_ready = true;
}

constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }

/* out[.., o, ..] (+)= sum_m M(o, m) in[.., m, ..] along direction DIR of an I0 x I1 x I2 tensor (x fastest).
 * M is [NO][NI] row-major; T = true uses the transpose of an [NI][NO] matrix. */
template <int I0, int I1, int I2, int DIR, int NO, bool T, bool ADD, typename V>
inline void sweep(const double *M, const V *in, V *out)
{
  constexpr int NI = DIR == 0 ? I0 : DIR == 1 ? I1 : I2;
  constexpr int O0 = DIR == 0 ? NO : I0, O1 = DIR == 1 ? NO : I1, O2 = DIR == 2 ? NO : I2;
  constexpr int sI = DIR == 0 ? 1 : DIR == 1 ? I0 : I0 * I1;
  constexpr int sO = DIR == 0 ? 1 : DIR == 1 ? O0 : O0 * O1;
  for (int k = 0; k < (DIR == 2 ? 1 : O2); ++k)
    for (int j = 0; j < (DIR == 1 ? 1 : O1); ++j)
      for (int i = 0; i < (DIR == 0 ? 1 : O0); ++i) {
        const int bi = i + j * I0 + k * I0 * I1, bo = i + j * O0 + k * O0 * O1;
        for (int o = 0; o < NO; ++o) {
          V s = (T ? M[o] : M[o * NI]) * in[bi];
          for (int m = 1; m < NI; ++m) s += (T ? M[m * NO + o] : M[o * NI + m]) * in[bi + m * sI];
          if (ADD) out[bo + o * sO] += s; else out[bo + o * sO] = s;
        }
      }
}

template <int DIM, typename V> inline void q2_to_quad(const V *u, V *uq)
{
  constexpr int Z = DIM == 3 ? 3 : 1;
  V t[27];
  sweep<3, 3, Z, 0, 3, false, false>(&N2[0][0], u, uq);
  sweep<3, 3, Z, 1, 3, false, false>(&N2[0][0], uq, t);
  if constexpr (DIM == 2) { for (int i = 0; i < 9; ++i) uq[i] = t[i]; }
  else sweep<3, 3, Z, 2, 3, false, false>(&N2[0][0], t, uq);
}
template <int DIM, typename V> inline void quad_to_q2_T(const V *vq, V *v)
{
  constexpr int Z = DIM == 3 ? 3 : 1;
  V t[27];
  sweep<3, 3, Z, 0, 3, true, false>(&N2[0][0], vq, v);
  sweep<3, 3, Z, 1, 3, true, false>(&N2[0][0], v, t);
  if constexpr (DIM == 2) { for (int i = 0; i < 9; ++i) v[i] = t[i]; }
  else sweep<3, 3, Z, 2, 3, true, false>(&N2[0][0], t, v);
}
template <int DIM, typename V> inline void q1_to_quad(const V *p, V *pq)
{
  V t[27], t2[27];
  if constexpr (DIM == 2) { sweep<2, 2, 1, 0, 3, false, false>(&N1[0][0], p, t); sweep<3, 2, 1, 1, 3, false, false>(&N1[0][0], t, pq); (void)t2; }
  else {
    sweep<2, 2, 2, 0, 3, false, false>(&N1[0][0], p, t); sweep<3, 2, 2, 1, 3, false, false>(&N1[0][0], t, t2);
    sweep<3, 3, 2, 2, 3, false, false>(&N1[0][0], t2, pq);
  }
}
template <int DIM, typename V> inline void quad_to_q1_T(const V *pq, V *p)
{
  V t[27], t2[27];
  if constexpr (DIM == 2) { sweep<3, 3, 1, 0, 2, true, false>(&N1[0][0], pq, t); sweep<2, 3, 1, 1, 2, true, false>(&N1[0][0], t, p); (void)t2; }
  else {
    sweep<3, 3, 3, 0, 2, true, false>(&N1[0][0], pq, t); sweep<2, 3, 3, 1, 2, true, false>(&N1[0][0], t, t2);
    sweep<2, 2, 3, 2, 2, true, false>(&N1[0][0], t2, p);
  }
}
template <int DIM, int DIR, typename V> inline void colloc_grad(const V *uq, V *g)
{ sweep<3, 3, DIM == 3 ? 3 : 1, DIR, 3, false, false>(&D2c[0][0], uq, g); }
template <int DIM, int DIR, typename V> inline void colloc_grad_T_add(const V *s, V *acc)
{ sweep<3, 3, DIM == 3 ? 3 : 1, DIR, 3, true, true>(&D2c[0][0], s, acc); }

template <int DIM> inline void quad_weights(const double *h, double *w)
{
  double vol = 1; for (int d = 0; d < DIM; ++d) vol *= h[d];
  for (int q = 0; q < ipow(3, DIM); ++q) {
    double ww = vol; int r = q;
    for (int d = 0; d < DIM; ++d) { ww *= GW[r % 3]; r /= 3; }
    w[q] = ww;
  }
}

/* velocity part of the cell kernel (matrix_free_operators.cc:376-547 without the Newton branches): nodal u
 * (component-major), pressure at the quadrature points pq (or nullptr) -> out_u and, if divq != nullptr, div u. */
template <int DIM, typename V>
inline void cell_velocity(const double *h, const double *w, const V *eta, int n_eta, bool compressible, const V *u, const V *pq,
                          double ps, bool do_viscous, V *out_u, V *divq)
{
  constexpr int NN = ipow(3, DIM), NQ = NN;
  V G[DIM][DIM][27];
  if (do_viscous || divq)
    for (int c = 0; c < DIM; ++c) {
      V uq[27]; q2_to_quad<DIM>(u + c * NN, uq);
      colloc_grad<DIM, 0>(uq, G[c][0]);
      colloc_grad<DIM, 1>(uq, G[c][1]);
      if constexpr (DIM == 3) colloc_grad<DIM, 2>(uq, G[c][2]);
      for (int d = 0; d < DIM; ++d) { const double ih = 1.0 / h[d]; for (int q = 0; q < NQ; ++q) G[c][d][q] *= ih; }
    }
  V S[DIM][DIM][27];
  for (int q = 0; q < NQ; ++q) {
    const V e2 = 2.0 * eta[n_eta == 1 ? 0 : q];
    V div = e2 - e2;
    if (do_viscous || divq) for (int c = 0; c < DIM; ++c) div += G[c][c][q];
    if (divq) divq[q] = div;
    for (int c = 0; c < DIM; ++c)
      for (int d = 0; d < DIM; ++d) {
        V s = e2 - e2;
        if (do_viscous) s = e2 * 0.5 * (G[c][d][q] + G[d][c][q]);
        if (c == d) {
          if (pq) s -= ps * pq[q];
          if (do_viscous && compressible) s -= e2 / 3.0 * div;
        }
        S[c][d][q] = s * (w[q] / h[d]);
      }
  }
  for (int c = 0; c < DIM; ++c) {
    V acc[27];
    for (int q = 0; q < NQ; ++q) acc[q] = S[0][0][q] - S[0][0][q];
    colloc_grad_T_add<DIM, 0>(S[c][0], acc);
    colloc_grad_T_add<DIM, 1>(S[c][1], acc);
    if constexpr (DIM == 3) colloc_grad_T_add<DIM, 2>(S[c][2], acc);
    quad_to_q2_T<DIM>(acc, out_u + c * NN);
  }
}

/* ---- greedy colouring of blocks of items by the nodes they touch ------------------------------------------ */
struct Coloring {
  std::vector<int64_t> block_begin;     /* item range of block b: [block_begin[b], block_begin[b+1]) */
  std::vector<int64_t> order;           /* block ids sorted by colour */
  std::vector<int64_t> color_ptr;
  template <typename F> void build(int64_t n_items, int64_t block, int64_t n_nodes, int n_idx, F idx_of)
  {
    const int64_t nb = (n_items + block - 1) / block;
    block_begin.resize(nb + 1);
    for (int64_t b = 0; b <= nb; ++b) block_begin[b] = std::min(b * block, n_items);
    std::vector<uint64_t> nodemask(n_nodes, 0);
    std::vector<int> color(nb);
    int n_colors = 0;
    for (int64_t b = 0; b < nb; ++b) {
      uint64_t m = 0;
      for (int64_t it = block_begin[b]; it < block_begin[b + 1]; ++it)
        for (int k = 0; k < n_idx; ++k) { const int64_t node = idx_of(it, k); if (node >= 0) m |= nodemask[node]; }
      int c = 0;
      while (c < 63 && ((m >> c) & 1)) ++c;
      color[b] = c; n_colors = std::max(n_colors, c + 1);
      for (int64_t it = block_begin[b]; it < block_begin[b + 1]; ++it)
        for (int k = 0; k < n_idx; ++k) { const int64_t node = idx_of(it, k); if (node >= 0) nodemask[node] |= uint64_t(1) << c; }
    }
    color_ptr.assign(n_colors + 1, 0);
    for (int64_t b = 0; b < nb; ++b) color_ptr[color[b] + 1]++;
    for (int c = 0; c < n_colors; ++c) color_ptr[c + 1] += color_ptr[c];
    order.resize(nb);
    std::vector<int64_t> pos(color_ptr.begin(), color_ptr.end() - 1);
    for (int64_t b = 0; b < nb; ++b) order[pos[color[b]]++] = b;
  }
  template <typename F> void for_each_block(F body) const
  {
    for (size_t c = 0; c + 1 < color_ptr.size(); ++c) {
      const int64_t b0 = color_ptr[c], b1 = color_ptr[c + 1];
      #pragma omp parallel for schedule(static) if (b1 - b0 >= 4)
      for (int64_t i = b0; i < b1; ++i) {
        const int64_t b = order[i];
        body(block_begin[b], block_begin[b + 1]);
      }
    }
  }
};

struct Level {
  int dim = 0; int64_t n_cells = 0, n_u = 0, n_p = 0, n_batches = 0;
  double h[3] = {1, 1, 1}, ps = 1.0; bool compressible = false; int n_eta = 1;
  std::vector<int32_t> bu, bp;          /* [batch][node][W], -1: constrained DoF (reads 0, not written) or padding lane */
  std::vector<double> beta;             /* [batch][n_eta][W] */
  std::vector<int32_t> con_u;           /* constrained velocity DoFs (unit diagonal) */
  Coloring col;
  std::vector<double> dinv_A, dinv_M;
};

struct Transfer {
  int dim = 0, npd = 0, ncomp = 0; int64_t n_items = 0, n_c = 0, n_f = 0;
  std::vector<int32_t> patch;           /* [item][(2 npd - 1)^dim] fine DoFs */
  std::vector<int32_t> coarse;          /* [item][npd^dim], -1 = constrained coarse DoF */
  std::vector<double> weights;          /* 1 / valence per fine DoF */
  Coloring col;
};

struct Cheb { double theta = 1, delta = 0, lam_min = 1, lam_max = 1; int degree = 1, cg_its = 0; };

struct Solver {
  int dim = 0, n_levels = 0;
  std::vector<Level> lv; Level act;
  std::vector<Transfer> tu, tp;
  std::vector<Cheb> chA, chS;
  /* scratch vectors per level and family (0 = u, 1 = p): x, t, xold, defect, residual */
  std::vector<std::vector<double>> sx[2], st[2], so[2], sd[2], sr[2];
  int64_t n_prec = 0;
  double t_vA = 0, t_vS = 0, t_op = 0, t_ortho = 0, t_loop = 0;
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/* ---- vector helpers ---------------------------------------------------------------------------------------- */
inline double dot(const double *a, const double *b, int64_t n)
{
  double s = 0;
  #pragma omp parallel for reduction(+ : s) schedule(static) if (n > 20000)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
inline void set_zero(double *a, int64_t n)
{
  #pragma omp parallel for schedule(static) if (n > 20000)
  for (int64_t i = 0; i < n; ++i) a[i] = 0.0;
}

/* classical Gram-Schmidt with re-orthogonalisation (the numerics of deal.II 9.6's delayed classical Gram-Schmidt): vv is
 * orthogonalised against the dim basis vectors V[0..dim), twice; h accumulates the coefficients.  The basis streams from
 * memory once per pass and direction (dot, then update), chunked so that the piece of vv stays in cache. */
void cgs2(const double *V, int64_t n, int dim, double *vv, double *h, double *hh)
{
  for (int pass = 0; pass < 2; ++pass) {
    std::fill(hh, hh + dim, 0.0);
    const int64_t chunk = 4096, nchunks = (n + chunk - 1) / chunk;
    #pragma omp parallel
    {
      std::vector<double> loc(dim, 0.0);
      #pragma omp for schedule(static) nowait
      for (int64_t cidx = 0; cidx < nchunks; ++cidx) {
        const int64_t a = cidx * chunk, b = std::min(n, a + chunk);
        for (int i = 0; i < dim; ++i) {
          const double *vi = &V[(size_t)i * n]; double s = 0;
          for (int64_t k = a; k < b; ++k) s += vi[k] * vv[k];
          loc[i] += s;
        }
      }
      #pragma omp critical
      for (int i = 0; i < dim; ++i) hh[i] += loc[i];
    }
    #pragma omp parallel for schedule(static)
    for (int64_t cidx = 0; cidx < nchunks; ++cidx) {
      const int64_t a = cidx * chunk, b = std::min(n, a + chunk);
      for (int i = 0; i < dim; ++i) {
        const double *vi = &V[(size_t)i * n]; const double c = hh[i];
        for (int64_t k = a; k < b; ++k) vv[k] -= c * vi[k];
      }
    }
    for (int i = 0; i < dim; ++i) h[i] += hh[i];
  }
}

/* ---- operators --------------------------------------------------------------------------------------------- */
template <int DIM> inline void gather(const double *x, const int32_t *idx, int n, vd *loc)
{
  for (int k = 0; k < n; ++k)
    for (int l = 0; l < W; ++l) { const int32_t i = idx[k * W + l]; loc[k][l] = i >= 0 ? x[i] : 0.0; }
}
template <int DIM> inline void scatter_add(double *y, const int32_t *idx, int n, const vd *loc)
{
  for (int k = 0; k < n; ++k)
    for (int l = 0; l < W; ++l) { const int32_t i = idx[k * W + l]; if (i >= 0) y[i] += loc[k][l]; }
}

enum Mode { MODE_A = 0, MODE_STOKES = 1, MODE_BT = 2 };

template <int DIM, int MODE>
void apply_velocity(const Level &L, const double *src_u, const double *src_p, double *dst_u, double *dst_p)
{
  constexpr int NN = ipow(3, DIM), NU = DIM * NN, NP = 1 << DIM, NQ = NN;
  double w[27]; quad_weights<DIM>(L.h, w);
  set_zero(dst_u, L.n_u);
  if (MODE == MODE_STOKES) set_zero(dst_p, L.n_p);
  L.col.for_each_block([&](int64_t b0, int64_t b1) {
    for (int64_t b = b0; b < b1; ++b) {
      vd u[NU], out[NU], p[NP], pq[27], divq[27], outp[NP];
      vd eta[27]; std::memcpy(eta, L.beta.data() + b * L.n_eta * W, sizeof(double) * L.n_eta * W);
      if (MODE != MODE_BT) gather<DIM>(src_u, L.bu.data() + b * NU * W, NU, u);
      if (MODE != MODE_A) { gather<DIM>(src_p, L.bp.data() + b * NP * W, NP, p); q1_to_quad<DIM>(p, pq); }
      cell_velocity<DIM, vd>(L.h, w, eta, L.n_eta, L.compressible, u, MODE == MODE_A ? nullptr : pq, L.ps, MODE != MODE_BT, out,
                             MODE == MODE_STOKES ? divq : nullptr);
      scatter_add<DIM>(dst_u, L.bu.data() + b * NU * W, NU, out);
      if (MODE == MODE_STOKES) {
        for (int q = 0; q < NQ; ++q) divq[q] *= -L.ps * w[q];
        quad_to_q1_T<DIM>(divq, outp);
        scatter_add<DIM>(dst_p, L.bp.data() + b * NP * W, NP, outp);
      }
    }
  });
  if (MODE != MODE_BT) {          /* unit diagonal on constrained rows (MatrixFreeOperators::Base::vmult) */
    const int64_t nc = (int64_t)L.con_u.size();
    for (int64_t i = 0; i < nc; ++i) dst_u[L.con_u[i]] = src_u[L.con_u[i]];
  }
}

template <int DIM> void apply_mass(const Level &L, const double *src, double *dst)
{
  constexpr int NP = 1 << DIM, NQ = ipow(3, DIM);
  double w[27]; quad_weights<DIM>(L.h, w);
  set_zero(dst, L.n_p);
  L.col.for_each_block([&](int64_t b0, int64_t b1) {
    for (int64_t b = b0; b < b1; ++b) {
      vd p[NP], pq[27], outp[NP];
      vd eta[27]; std::memcpy(eta, L.beta.data() + b * L.n_eta * W, sizeof(double) * L.n_eta * W);
      gather<DIM>(src, L.bp.data() + b * NP * W, NP, p);
      q1_to_quad<DIM>(p, pq);
      for (int q = 0; q < NQ; ++q) pq[q] *= (L.ps * L.ps * w[q]) / eta[L.n_eta == 1 ? 0 : q];
      quad_to_q1_T<DIM>(pq, outp);
      scatter_add<DIM>(dst, L.bp.data() + b * NP * W, NP, outp);
    }
  });
}

template <int DIM> void compute_diagonals(Level &L)
{
  constexpr int NN = ipow(3, DIM), NU = DIM * NN, NP = 1 << DIM, NQ = NN;
  double w[27]; quad_weights<DIM>(L.h, w);
  std::vector<double> dA(L.n_u, 0.0), dM(L.n_p, 0.0);
  L.col.for_each_block([&](int64_t b0, int64_t b1) {
    for (int64_t b = b0; b < b1; ++b) {
      vd u[NU], out[NU], d[NU], p[NP], pq[27], outp[NP], dp[NP];
      vd eta[27]; std::memcpy(eta, L.beta.data() + b * L.n_eta * W, sizeof(double) * L.n_eta * W);
      const vd zero = eta[0] - eta[0], one = zero + 1.0;
      for (int i = 0; i < NU; ++i) {
        for (int k = 0; k < NU; ++k) u[k] = zero;
        u[i] = one;
        cell_velocity<DIM, vd>(L.h, w, eta, L.n_eta, L.compressible, u, nullptr, 0.0, true, out, nullptr);
        d[i] = out[i];
      }
      scatter_add<DIM>(dA.data(), L.bu.data() + b * NU * W, NU, d);
      for (int i = 0; i < NP; ++i) {
        for (int k = 0; k < NP; ++k) p[k] = zero;
        p[i] = one;
        q1_to_quad<DIM>(p, pq);
        for (int q = 0; q < NQ; ++q) pq[q] *= (L.ps * L.ps * w[q]) / eta[L.n_eta == 1 ? 0 : q];
        quad_to_q1_T<DIM>(pq, outp);
        dp[i] = outp[i];
      }
      scatter_add<DIM>(dM.data(), L.bp.data() + b * NP * W, NP, dp);
    }
  });
  for (int32_t i : L.con_u) dA[i] = 1.0;
  L.dinv_A.resize(L.n_u); L.dinv_M.resize(L.n_p);
  for (int64_t i = 0; i < L.n_u; ++i) L.dinv_A[i] = 1.0 / dA[i];
  for (int64_t i = 0; i < L.n_p; ++i) L.dinv_M[i] = 1.0 / dM[i];
}

void vmult_A(const Level &L, const double *src, double *dst)
{ if (L.dim == 2) apply_velocity<2, MODE_A>(L, src, nullptr, dst, nullptr); else apply_velocity<3, MODE_A>(L, src, nullptr, dst, nullptr); }
void vmult_M(const Level &L, const double *src, double *dst)
{ if (L.dim == 2) apply_mass<2>(L, src, dst); else apply_mass<3>(L, src, dst); }
void vmult_stokes(const Level &L, const double *su, const double *sp, double *du, double *dp)
{ if (L.dim == 2) apply_velocity<2, MODE_STOKES>(L, su, sp, du, dp); else apply_velocity<3, MODE_STOKES>(L, su, sp, du, dp); }
void vmult_BT(const Level &L, const double *sp, double *du)
{ if (L.dim == 2) apply_velocity<2, MODE_BT>(L, nullptr, sp, du, nullptr); else apply_velocity<3, MODE_BT>(L, nullptr, sp, du, nullptr); }

/* ---- transfers (MGTransferMatrixFree on a uniformly refined mesh) ------------------------- */
template <int DIM, int NPD> void transfer_apply(const Transfer &T, bool restrict_, const double *src, double *dst)
{
  constexpr int NP5 = 2 * NPD - 1, NN = ipow(NPD, DIM), NPATCH = ipow(NP5, DIM);
  const double *P = NPD == 3 ? &P2[0][0] : &P1[0][0];
  T.col.for_each_block([&](int64_t i0, int64_t i1) {
    for (int64_t it = i0; it < i1; ++it) {
      const int32_t *patch = T.patch.data() + it * NPATCH, *cd = T.coarse.data() + it * NN;
      double t[125], t2[125];
      if (!restrict_) {
        for (int k = 0; k < NN; ++k) t[k] = cd[k] >= 0 ? src[cd[k]] : 0.0;
        if constexpr (DIM == 2) {
          sweep<NPD, NPD, 1, 0, NP5, false, false>(P, t, t2); sweep<NP5, NPD, 1, 1, NP5, false, false>(P, t2, t);
        } else {
          sweep<NPD, NPD, NPD, 0, NP5, false, false>(P, t, t2); sweep<NP5, NPD, NPD, 1, NP5, false, false>(P, t2, t);
          sweep<NP5, NP5, NPD, 2, NP5, false, false>(P, t, t2); std::memcpy(t, t2, sizeof(double) * NPATCH);
        }
        for (int k = 0; k < NPATCH; ++k) dst[patch[k]] += T.weights[patch[k]] * t[k];
      } else {
        for (int k = 0; k < NPATCH; ++k) t[k] = T.weights[patch[k]] * src[patch[k]];
        if constexpr (DIM == 2) {
          sweep<NP5, NP5, 1, 0, NPD, true, false>(P, t, t2); sweep<NPD, NP5, 1, 1, NPD, true, false>(P, t2, t);
        } else {
          sweep<NP5, NP5, NP5, 0, NPD, true, false>(P, t, t2); sweep<NPD, NP5, NP5, 1, NPD, true, false>(P, t2, t);
          sweep<NPD, NPD, NP5, 2, NPD, true, false>(P, t, t2); std::memcpy(t, t2, sizeof(double) * NN);
        }
        for (int k = 0; k < NN; ++k) if (cd[k] >= 0) dst[cd[k]] += t[k];
      }
    }
  });
}
void transfer(const Transfer &T, bool restrict_, const double *src, double *dst)
{
  set_zero(dst, restrict_ ? T.n_c : T.n_f);
  if (T.dim == 2) { if (T.npd == 3) transfer_apply<2, 3>(T, restrict_, src, dst); else transfer_apply<2, 2>(T, restrict_, src, dst); }
  else { if (T.npd == 3) transfer_apply<3, 3>(T, restrict_, src, dst); else transfer_apply<3, 2>(T, restrict_, src, dst); }
}

/* ---- eigenvalues of a symmetric tridiagonal matrix: bisection on the Sturm count ---------------------------- */
int sturm_count(const std::vector<double> &d, const std::vector<double> &e, double x)
{
  int cnt = 0; double q = 1.0;
  for (size_t i = 0; i < d.size(); ++i) {
    q = d[i] - x - (i > 0 ? e[i - 1] * e[i - 1] / q : 0.0);
    if (q == 0.0) q = 1e-300;
    if (q < 0) ++cnt;
  }
  return cnt;
}
double tridiag_eig(const std::vector<double> &d, const std::vector<double> &e, int k)     /* k-th smallest, 0-based */
{
  double lo = 1e300, hi = -1e300;
  for (size_t i = 0; i < d.size(); ++i) {
    const double r = (i > 0 ? std::fabs(e[i - 1]) : 0.0) + (i + 1 < d.size() ? std::fabs(e[i]) : 0.0);
    lo = std::min(lo, d[i] - r); hi = std::max(hi, d[i] + r);
  }
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid == lo || mid == hi) break;
    if (sturm_count(d, e, mid) > k) hi = mid; else lo = mid;
  }
  return 0.5 * (lo + hi);
}

/* ---- PreconditionChebyshev (deal.II 9.6), App. A.2 ---------------------------------------------------------- */
template <typename Apply>
void estimate_eigenvalues(Cheb &c, Apply apply, const double *dinv, int64_t n, int degree, double smoothing_range, int eig_its,
                          double eig_res = 1e-7)
{
  std::vector<double> b(n), x(n, 0.0), r(n), z(n), p(n), v(n);
  double mean = 0;
  for (int64_t i = 0; i < n; ++i) { b[i] = double(i % 11); mean += b[i]; }      /* set_initial_guess: global index mod 11 */
  mean /= double(n);
  for (int64_t i = 0; i < n; ++i) { b[i] -= mean; r[i] = b[i]; }
  double res = std::sqrt(dot(r.data(), r.data(), n));
  std::vector<double> diag, off;
  int it = 0; double lam_min = 1.0, lam_max = 1.0;
  if (!(res <= eig_res || eig_its == 0)) {
    double rz_old = 0, alpha_old = 0, beta = 0;
    while (true) {
      ++it;
      #pragma omp parallel for schedule(static) if (n > 20000)
      for (int64_t i = 0; i < n; ++i) z[i] = dinv[i] * r[i];
      const double rz = dot(r.data(), z.data(), n);
      if (it > 1) {
        beta = rz / rz_old;
        #pragma omp parallel for schedule(static) if (n > 20000)
        for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
      } else p = z;
      apply(p.data(), v.data());
      const double alpha = rz / dot(p.data(), v.data(), n);
      #pragma omp parallel for schedule(static) if (n > 20000)
      for (int64_t i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * v[i]; }
      res = std::sqrt(dot(r.data(), r.data(), n));
      if (it > 1) { diag.push_back(1.0 / alpha + beta / alpha_old); off.push_back(std::sqrt(beta) / alpha_old); }
      else diag.push_back(1.0 / alpha);
      rz_old = rz; alpha_old = alpha;
      if (it >= eig_its || res <= eig_res) break;
    }
    lam_min = tridiag_eig(diag, off, 0);
    lam_max = tridiag_eig(diag, off, int(diag.size()) - 1);
  }
  c.cg_its = it; c.lam_min = lam_min; c.lam_max = lam_max; c.degree = degree;
  const double lmax = 1.2 * lam_max;
  const double alpha = smoothing_range > 1.0 ? lmax / smoothing_range : std::min(0.9 * lmax, lam_min);
  c.delta = 0.5 * (lmax - alpha);
  c.theta = 0.5 * (lmax + alpha);
}

/* x = x_in + ... one Chebyshev sweep; from_zero: x = p(A) b with a zero start (degree-1 matvecs), else degree matvecs.
 * x, t, xold are level scratch vectors; result in x. */
template <typename Apply>
void cheb_apply(const Cheb &c, Apply apply, const double *dinv, int64_t n, bool from_zero, double *x, const double *b, double *t, double *xold)
{
  const double theta = c.theta, delta = c.delta;
  if (from_zero) {
    #pragma omp parallel for schedule(static) if (n > 20000)
    for (int64_t i = 0; i < n; ++i) { x[i] = dinv[i] * b[i] / theta; xold[i] = 0.0; }
  } else {
    apply(x, t);
    #pragma omp parallel for schedule(static) if (n > 20000)
    for (int64_t i = 0; i < n; ++i) { const double xi = x[i]; x[i] = xi + dinv[i] * (b[i] - t[i]) / theta; xold[i] = xi; }
  }
  if (c.degree < 2 || std::fabs(delta) < 1e-40) return;
  double rhok = delta / theta; const double sigma = theta / delta;
  for (int k = 0; k < c.degree - 1; ++k) {
    apply(x, t);
    const double rhokp = 1.0 / (2.0 * sigma - rhok);
    const double f1 = rhokp * rhok, f2 = 2.0 * rhokp / delta;
    rhok = rhokp;
    #pragma omp parallel for schedule(static) if (n > 20000)
    for (int64_t i = 0; i < n; ++i) {
      const double xi = x[i];
      x[i] = (1.0 + f1) * xi - f1 * xold[i] + f2 * dinv[i] * (b[i] - t[i]);
      xold[i] = xi;
    }
  }
}

/* ---- V-cycle (Multigrid + SmootherRelaxation + MGCoarseGridApplySmoother), App. A.3 -------------------------- */
void vcycle_level(Solver &S, int fam, int l, int steps)
{
  /* in: S.sd[fam][l] = defect; out: S.sx[fam][l] */
  Level &L = S.lv[l];
  const int64_t n = fam == 0 ? L.n_u : L.n_p;
  const double *dinv = fam == 0 ? L.dinv_A.data() : L.dinv_M.data();
  const Cheb &c = fam == 0 ? S.chA[l] : S.chS[l];
  auto apply = [&](const double *a, double *b) { if (fam == 0) vmult_A(L, a, b); else vmult_M(L, a, b); };
  double *x = S.sx[fam][l].data(), *d = S.sd[fam][l].data(), *t = S.st[fam][l].data(), *xo = S.so[fam][l].data();
  /* SmootherRelaxation::apply: first step from zero, then steps-1 sweeps */
  cheb_apply(c, apply, dinv, n, true, x, d, t, xo);
  for (int s = 1; s < steps; ++s) cheb_apply(c, apply, dinv, n, false, x, d, t, xo);
  if (l == 0) return;
  apply(x, t);
  double *r = S.sr[fam][l].data();
  #pragma omp parallel for schedule(static) if (n > 20000)
  for (int64_t i = 0; i < n; ++i) r[i] = d[i] - t[i];
  const Transfer &T = fam == 0 ? S.tu[l] : S.tp[l];
  transfer(T, true, r, S.sd[fam][l - 1].data());
  vcycle_level(S, fam, l - 1, steps);
  transfer(T, false, S.sx[fam][l - 1].data(), t);
  #pragma omp parallel for schedule(static) if (n > 20000)
  for (int64_t i = 0; i < n; ++i) x[i] += t[i];
  for (int s = 0; s < steps; ++s) cheb_apply(c, apply, dinv, n, false, x, d, t, xo);
}

void vcycle(Solver &S, int fam, const double *src, double *dst)
{
  const int l = S.n_levels - 1;
  const int64_t n = fam == 0 ? S.lv[l].n_u : S.lv[l].n_p;
  std::memcpy(S.sd[fam][l].data(), src, sizeof(double) * n);
  vcycle_level(S, fam, l, fam == 0 ? 1 : 4);
  std::memcpy(dst, S.sx[fam][l].data(), sizeof(double) * n);
}

/* BlockSchurPreconditioner::vmult, cheap variant (block_stokes_preconditioner.h:220-255) */
void precondition(Solver &S, const double *r, double *z, double *tmp_u)
{
  const int64_t nu = S.act.n_u, np = S.act.n_p;
  ++S.n_prec;
  double t0 = now();
  vcycle(S, 1, r + nu, z + nu);
  S.t_vS += now() - t0;
  for (int64_t i = 0; i < np; ++i) z[nu + i] = -z[nu + i];
  vmult_BT(S.act, z + nu, tmp_u);
  #pragma omp parallel for schedule(static) if (nu > 20000)
  for (int64_t i = 0; i < nu; ++i) tmp_u[i] = r[i] - tmp_u[i];
  t0 = now();
  vcycle(S, 0, tmp_u, z);
  S.t_vA += now() - t0;
}

void setup_smoothers(Solver &S)
{
  S.chA.assign(S.n_levels, Cheb()); S.chS.assign(S.n_levels, Cheb());
  for (int l = 0; l < S.n_levels; ++l) {
    Level &L = S.lv[l];
    auto aA = [&](const double *a, double *b) { vmult_A(L, a, b); };
    auto aM = [&](const double *a, double *b) { vmult_M(L, a, b); };
    if (l > 0) {
      estimate_eigenvalues(S.chA[l], aA, L.dinv_A.data(), L.n_u, 4, 15.0, 10);
      estimate_eigenvalues(S.chS[l], aM, L.dinv_M.data(), L.n_p, 4, 15.0, 10);
    } else {
      estimate_eigenvalues(S.chA[l], aA, L.dinv_A.data(), L.n_u, 8, 1e-3, 100);
      estimate_eigenvalues(S.chS[l], aM, L.dinv_M.data(), L.n_p, 8, 1e-3, 100);
    }
  }
}

template <int DIM> void fill_level(Level &L, int64_t n_cells, int64_t n_u, int64_t n_p, const int64_t *cell_u, const int64_t *cell_p,
                                   const double *h, const double *eta, int n_eta, double ps, int compressible, const int64_t *con, int64_t n_con)
{
  constexpr int NN = ipow(3, DIM), NU = DIM * NN, NP = 1 << DIM;
  L.dim = DIM; L.n_cells = n_cells; L.n_u = n_u; L.n_p = n_p; L.n_eta = n_eta; L.ps = ps; L.compressible = compressible != 0;
  for (int d = 0; d < DIM; ++d) L.h[d] = h[d];
  std::vector<uint8_t> mask(n_u, 0);
  L.con_u.resize(n_con);
  for (int64_t i = 0; i < n_con; ++i) { mask[con[i]] = 1; L.con_u[i] = (int32_t)con[i]; }
  L.n_batches = (n_cells + W - 1) / W;
  L.bu.assign(L.n_batches * NU * W, -1); L.bp.assign(L.n_batches * NP * W, -1); L.beta.assign(L.n_batches * n_eta * W, 1.0);
  for (int64_t c = 0; c < n_cells; ++c) {
    const int64_t b = c / W; const int l = int(c % W);
    for (int k = 0; k < NU; ++k) { const int64_t i = cell_u[c * NU + k]; L.bu[(b * NU + k) * W + l] = mask[i] ? -1 : (int32_t)i; }
    for (int k = 0; k < NP; ++k) L.bp[(b * NP + k) * W + l] = (int32_t)cell_p[c * NP + k];
    for (int k = 0; k < n_eta; ++k) L.beta[(b * n_eta + k) * W + l] = eta[c * n_eta + k];
  }
  /* blocks of 64 cells (a 4x4x4 brick in Morton order), smaller on small levels so that every colour has work for all threads */
  int64_t cells_per_block = 64;
  while (cells_per_block > W && n_cells / cells_per_block < 256) cells_per_block /= 2;
  const int64_t bpb = std::max<int64_t>(1, cells_per_block / W);
  /* colour by ALL velocity nodes a batch touches (component 0; the unmasked map) */
  L.col.build(L.n_batches, bpb, n_u / DIM, NN * W, [&](int64_t b, int k) -> int64_t {
    const int64_t c = b * W + k % W; if (c >= n_cells) return -1;
    return cell_u[c * NU + k / W] / DIM; });
  compute_diagonals<DIM>(L);
}

template <int DIM, int NPD> void fill_transfer(Transfer &T, int ncomp, int64_t n_par, const int64_t *children, const int64_t *cell_c,
                                               const int64_t *cell_f, int64_t n_c, int64_t n_f, const int64_t *con_c, int64_t n_con)
{
  constexpr int NP5 = 2 * NPD - 1, NN = ipow(NPD, DIM), NPATCH = ipow(NP5, DIM);
  const int nloc = ncomp * NN;
  T.dim = DIM; T.npd = NPD; T.ncomp = ncomp; T.n_c = n_c; T.n_f = n_f; T.n_items = n_par * ncomp;
  std::vector<uint8_t> mask(n_c, 0);
  for (int64_t i = 0; i < n_con; ++i) mask[con_c[i]] = 1;
  T.patch.resize(T.n_items * NPATCH); T.coarse.resize(T.n_items * NN);
  std::vector<double> val(n_f, 0.0);
  for (int64_t par = 0; par < n_par; ++par)
    for (int c = 0; c < ncomp; ++c) {
      const int64_t it = par * ncomp + c;
      int32_t *patch = T.patch.data() + it * NPATCH;
      for (int k = 0; k < (1 << DIM); ++k) {
        const int64_t child = children[par * (1 << DIM) + k];
        for (int a = 0; a < NN; ++a) {
          int r = a, pi = 0, mul = 1;
          for (int d = 0; d < DIM; ++d) { const int loc = r % NPD; r /= NPD; pi += (loc + (NPD - 1) * ((k >> d) & 1)) * mul; mul *= NP5; }
          patch[pi] = (int32_t)cell_f[child * nloc + c * NN + a];
        }
      }
      for (int a = 0; a < NN; ++a) { const int64_t i = cell_c[par * nloc + c * NN + a]; T.coarse[it * NN + a] = mask[i] ? -1 : (int32_t)i; }
      for (int k = 0; k < NPATCH; ++k) val[patch[k]] += 1.0;
    }
  T.weights.resize(n_f);
  for (int64_t i = 0; i < n_f; ++i) T.weights[i] = 1.0 / std::max(val[i], 1.0);
  int64_t block = 64 * ncomp;
  while (block > ncomp && T.n_items / block < 256) block /= 2;
  T.col.build(T.n_items, std::max<int64_t>(block, 1), n_f, NPATCH, [&](int64_t it, int k) -> int64_t { return T.patch[it * NPATCH + k]; });
}

} // namespace

extern "C" {

struct smfc_result { int64_t iterations, n_prec; double residual, tolerance, initial_residual, seconds, setup_seconds,
                     seconds_vcycle_A, seconds_vcycle_S, seconds_stokes, seconds_ortho, seconds_loop; int converged; };

void *smfc_create(int dim, int n_levels)
{
  init_tables();
  Solver *S = new Solver; S->dim = dim; S->n_levels = n_levels;
  S->lv.resize(n_levels); S->tu.resize(n_levels); S->tp.resize(n_levels);
  return S;
}
void smfc_destroy(void *h) { delete static_cast<Solver *>(h); }
int smfc_width(void) { return W; }
int smfc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void smfc_set_num_threads(int n)
{
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* level < 0: the active level (Stokes / B^T operators, its own viscosity table and n_eta) */
void smfc_set_level(void *hd, int level, int64_t n_cells, int64_t n_u, int64_t n_p, const int64_t *cell_u, const int64_t *cell_p,
                    const double *h, const double *eta, int n_eta, double ps, int compressible, const int64_t *con, int64_t n_con)
{
  Solver &S = *static_cast<Solver *>(hd);
  Level &L = level < 0 ? S.act : S.lv[level];
  if (S.dim == 2) fill_level<2>(L, n_cells, n_u, n_p, cell_u, cell_p, h, eta, n_eta, ps, compressible, con, n_con);
  else fill_level<3>(L, n_cells, n_u, n_p, cell_u, cell_p, h, eta, n_eta, ps, compressible, con, n_con);
}

/* transfer between level-1 and level for family 0 (velocity) / 1 (pressure) */
void smfc_set_transfer(void *hd, int level, int family, int64_t n_par, const int64_t *children, const int64_t *cell_c, const int64_t *cell_f,
                       int64_t n_c, int64_t n_f, const int64_t *con_c, int64_t n_con)
{
  Solver &S = *static_cast<Solver *>(hd);
  if (family == 0) {
    if (S.dim == 2) fill_transfer<2, 3>(S.tu[level], 2, n_par, children, cell_c, cell_f, n_c, n_f, con_c, n_con);
    else fill_transfer<3, 3>(S.tu[level], 3, n_par, children, cell_c, cell_f, n_c, n_f, con_c, n_con);
  } else {
    if (S.dim == 2) fill_transfer<2, 2>(S.tp[level], 1, n_par, children, cell_c, cell_f, n_c, n_f, nullptr, 0);
    else fill_transfer<3, 2>(S.tp[level], 1, n_par, children, cell_c, cell_f, n_c, n_f, nullptr, 0);
  }
}

void smfc_finalize(void *hd)
{
  Solver &S = *static_cast<Solver *>(hd);
  for (int f = 0; f < 2; ++f) {
    S.sx[f].resize(S.n_levels); S.st[f].resize(S.n_levels); S.so[f].resize(S.n_levels); S.sd[f].resize(S.n_levels); S.sr[f].resize(S.n_levels);
    for (int l = 0; l < S.n_levels; ++l) {
      const int64_t n = f == 0 ? S.lv[l].n_u : S.lv[l].n_p;
      S.sx[f][l].assign(n, 0.0); S.st[f][l].assign(n, 0.0); S.so[f][l].assign(n, 0.0); S.sd[f][l].assign(n, 0.0); S.sr[f][l].assign(n, 0.0);
    }
  }
}

void smfc_vmult_A(void *hd, int level, const double *src, double *dst) { Solver &S = *static_cast<Solver *>(hd); vmult_A(level < 0 ? S.act : S.lv[level], src, dst); }
void smfc_vmult_mass(void *hd, int level, const double *src, double *dst) { Solver &S = *static_cast<Solver *>(hd); vmult_M(level < 0 ? S.act : S.lv[level], src, dst); }
void smfc_vmult_stokes(void *hd, const double *su, const double *sp, double *du, double *dp) { vmult_stokes(static_cast<Solver *>(hd)->act, su, sp, du, dp); }
void smfc_vmult_BT(void *hd, const double *sp, double *du) { vmult_BT(static_cast<Solver *>(hd)->act, sp, du); }
void smfc_dinv(void *hd, int level, int family, double *out)
{
  Solver &S = *static_cast<Solver *>(hd);
  const std::vector<double> &d = family == 0 ? S.lv[level].dinv_A : S.lv[level].dinv_M;
  std::memcpy(out, d.data(), sizeof(double) * d.size());
}
void smfc_transfer(void *hd, int level, int family, int restrict_, const double *src, double *dst)
{ Solver &S = *static_cast<Solver *>(hd); transfer(family == 0 ? S.tu[level] : S.tp[level], restrict_ != 0, src, dst); }
void smfc_setup_smoothers(void *hd) { setup_smoothers(*static_cast<Solver *>(hd)); }
void smfc_smoother_info(void *hd, int level, int family, double *out4)
{
  Solver &S = *static_cast<Solver *>(hd); const Cheb &c = family == 0 ? S.chA[level] : S.chS[level];
  out4[0] = c.lam_min; out4[1] = c.lam_max; out4[2] = double(c.cg_its); out4[3] = c.theta;
}
void smfc_vcycle(void *hd, int family, const double *src, double *dst) { vcycle(*static_cast<Solver *>(hd), family, src, dst); }
void smfc_precondition(void *hd, const double *r, double *z)
{
  Solver &S = *static_cast<Solver *>(hd);
  std::vector<double> tmp(S.act.n_u);
  precondition(S, r, z, tmp.data());
}

/* Seconds of one CGS2 orthogonalisation (the code of the solve) of a block vector against a basis of k vectors: the part
 * of a GMRES iteration whose cost grows with the iteration number.  bench.py uses it at k = (iterations + 1) / 2, the mean
 * basis length of a cycle, when a complete solve at refinement 7 is too long to repeat for every step. */
double smfc_time_ortho(void *hd, int k, int reps)
{
  Solver &S = *static_cast<Solver *>(hd);
  const int64_t n = S.act.n_u + S.act.n_p;
  std::unique_ptr<double[]> V(new double[(size_t)k * n]), vv(new double[n]);
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)k * n; ++i) V[i] = 1e-3 * double((i * 2654435761u) % 1000) - 0.5;
  std::vector<double> h(k + 1), hh(k + 1);
  double best = 1e300;
  for (int r = 0; r < reps; ++r) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) vv[i] = 1.0 + 1e-3 * double(i % 7);
    std::fill(h.begin(), h.end(), 0.0);
    const double t0 = now();
    cgs2(V.get(), n, k, vv.get(), h.data(), hh.data());
    best = std::min(best, now() - t0);
  }
  return best;
}

/* StokesMatrixFreeHandler::solve(), cheap phase: smoother set-up + eigenvalue estimation, solver tolerance
 * (local_smoothing.cc:1016-1025), right-preconditioned GMRES(restart) with CGS2 (oracle/solver.py: gmres_right). */
int smfc_solve(void *hd, const double *rhs, double *x, double tol_rel, int restart, int max_its, smfc_result *out)
{
  Solver &S = *static_cast<Solver *>(hd);
  const int64_t nu = S.act.n_u, np = S.act.n_p, n = nu + np;
  const double t0 = now();
  setup_smoothers(S);
  const double t1 = now();
  S.n_prec = 0;
  std::vector<double> tmp_u(nu), p(n), vv(n), z(n);
  S.t_vA = S.t_vS = S.t_op = S.t_ortho = S.t_loop = 0;
  auto op = [&](const double *a, double *b) { const double t = now(); vmult_stokes(S.act, a, a + nu, b, b + nu); S.t_op += now() - t; };
  /* tolerance: velocity part of the guess zeroed */
  {
    std::vector<double> g(x, x + n);
    std::fill(g.begin(), g.begin() + nu, 0.0);
    op(g.data(), vv.data());
    double ru = 0, rp = 0;
    for (int64_t i = 0; i < nu; ++i) { const double d = rhs[i] - vv[i]; ru += d * d; }
    for (int64_t i = 0; i < np; ++i) rp += rhs[nu + i] * rhs[nu + i];
    out->tolerance = tol_rel * std::sqrt(ru + rp);
  }
  const double tol = out->tolerance;
  /* basis: allocated untouched so that only the vectors a cycle really builds occupy memory */
  std::unique_ptr<double[]> Vmem(new double[(size_t)(restart + 1) * n]);
  double *V = Vmem.get();
  std::vector<double> H((size_t)(restart + 1) * restart), gamma(restart + 1), ci(restart), si(restart), h(restart + 1), hh(restart + 1);
  int64_t its = 0; double rho = 0; int state = 0;     /* 0 iterate, 1 success, 2 failure */
  bool first = true;
  while (state == 0) {
    op(x, p.data());
    #pragma omp parallel for schedule(static) if (n > 20000)
    for (int64_t i = 0; i < n; ++i) p[i] = rhs[i] - p[i];
    rho = std::sqrt(dot(p.data(), p.data(), n));
    if (first) { out->initial_residual = rho; first = false; }
    if (rho <= tol) { state = 1; break; }
    if (its >= max_its) { state = 2; break; }
    #pragma omp parallel for schedule(static) if (n > 20000)
    for (int64_t i = 0; i < n; ++i) V[i] = p[i] / rho;
    std::fill(gamma.begin(), gamma.end(), 0.0); gamma[0] = rho;
    std::fill(H.begin(), H.end(), 0.0);
    int dim = 0;
    for (int j = 0; j < restart && state == 0; ++j) {
      const double t_j = now();
      ++its;
      precondition(S, &V[(size_t)j * n], z.data(), tmp_u.data());
      op(z.data(), vv.data());
      dim = j + 1;
      std::fill(h.begin(), h.end(), 0.0);
      const double t_o = now();
      cgs2(V, n, dim, vv.data(), h.data(), hh.data());
      const double nrm = std::sqrt(dot(vv.data(), vv.data(), n));
      h[dim] = nrm;
      double *vn = &V[(size_t)(j + 1) * n];
      #pragma omp parallel for schedule(static) if (n > 20000)
      for (int64_t i = 0; i < n; ++i) vn[i] = nrm != 0 ? vv[i] / nrm : vv[i];
      S.t_ortho += now() - t_o;
      for (int i = 0; i < j; ++i) {
        const double t = h[i];
        h[i] = ci[i] * t + si[i] * h[i + 1];
        h[i + 1] = -si[i] * t + ci[i] * h[i + 1];
      }
      const double s = std::hypot(h[j], h[j + 1]);
      ci[j] = h[j] / s; si[j] = h[j + 1] / s;
      h[j] = s;
      gamma[j + 1] = -si[j] * gamma[j];
      gamma[j] *= ci[j];
      for (int i = 0; i < dim; ++i) H[(size_t)i * restart + j] = h[i];
      rho = std::fabs(gamma[dim]);
      if (rho <= tol) state = 1;
      else if (its >= max_its) state = 2;
      S.t_loop += now() - t_j;
    }
    /* back substitution H[:dim,:dim] y = gamma[:dim] */
    std::vector<double> y(dim);
    for (int i = dim - 1; i >= 0; --i) {
      double s = gamma[i];
      for (int k = i + 1; k < dim; ++k) s -= H[(size_t)i * restart + k] * y[k];
      y[i] = s / H[(size_t)i * restart + i];
    }
    #pragma omp parallel for schedule(static) if (n > 20000)
    for (int64_t k = 0; k < n; ++k) { double s = 0; for (int i = 0; i < dim; ++i) s += y[i] * V[(size_t)i * n + k]; p[k] = s; }
    precondition(S, p.data(), z.data(), tmp_u.data());
    #pragma omp parallel for schedule(static) if (n > 20000)
    for (int64_t k = 0; k < n; ++k) x[k] += z[k];
  }
  out->iterations = its; out->n_prec = S.n_prec; out->residual = rho; out->converged = state == 1;
  out->seconds = now() - t0; out->setup_seconds = t1 - t0;
  out->seconds_vcycle_A = S.t_vA; out->seconds_vcycle_S = S.t_vS; out->seconds_stokes = S.t_op; out->seconds_ortho = S.t_ortho; out->seconds_loop = S.t_loop;
  return 0;
}

} // extern "C"
