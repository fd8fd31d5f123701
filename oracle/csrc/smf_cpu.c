/* CPU restatement of the matrix-free Stokes operators (ORACLE / TEST INFRASTRUCTURE ONLY).
 *
 * Restates, in plain C with FEEvaluation-style sum factorisation, the cell kernels of
 *   /root/reference/source/simulator/solver/matrix_free_operators.cc
 *     StokesOperator::local_apply        :376-547   (incompressible/compressible, no Newton terms)
 *     BTBlockOperator::local_apply       :668-701
 *     MassMatrixOperator::local_apply    :787-845
 *     ABlockOperator::local_apply        :974-1052
 *     compute_diagonal (A, mass)         :1068-1098, :861-908
 * plus the deal.II semantics of MatrixFreeOperators::Base::vmult (dst = 0; constrained DoFs
 * read 0 / are not written / get a unit diagonal) and of
 * MGTransferMatrixFree prolongate / restrict_and_add on uniformly refined meshes (App. A.3).
 * It is used by tests/ as the checker and by bench.py as the `cpu_baseline` /
 * `--impl reference` arm.  Nothing in aspect_b200/ links or loads this file.
 *
 * Cells are Cartesian (per-level cell size h[dim]); Q2 velocity / Q1 pressure; 3-point Gauss
 * rule per direction.  Loops over cells are OpenMP-parallel over colours (cells of one colour
 * share no DoF), mirroring one-thread-per-core execution of deal.II's cell_loop.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double N2[3][3], D2c[3][3], N1[3][2], GW[3];
static int tables_ready = 0;

static void init_tables(void)
{
  if (tables_ready) return;
  const double s = sqrt(3.0 / 5.0);
  const double x[3] = {0.5 - 0.5 * s, 0.5, 0.5 + 0.5 * s};
  double D2[3][3];
  GW[0] = 5.0 / 18.0; GW[1] = 8.0 / 18.0; GW[2] = 5.0 / 18.0;
  for (int q = 0; q < 3; ++q) {
    N2[q][0] = (1 - x[q]) * (1 - 2 * x[q]); N2[q][1] = 4 * x[q] * (1 - x[q]); N2[q][2] = x[q] * (2 * x[q] - 1);
    D2[q][0] = 4 * x[q] - 3; D2[q][1] = 4 - 8 * x[q]; D2[q][2] = 4 * x[q] - 1;
    N1[q][0] = 1 - x[q]; N1[q][1] = x[q];
  }
  /* collocation derivative: derivative at Gauss point q of the Lagrange polynomial through Gauss point j */
  for (int q = 0; q < 3; ++q)
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
  tables_ready = 1;
}

/* out[..o..] = sum_i M[o][i] in[..i..] along direction dir of a (dim)-tensor; sizes given per direction. */
static void sweep(const double *M, int mo, int mi, int transpose, const double *in, double *out, int dir, int dim, const int *sz_in)
{
  int so[3] = {1, 1, 1}, si[3] = {1, 1, 1}, szo[3] = {1, 1, 1}, szi[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d) { szi[d] = sz_in[d]; szo[d] = (d == dir) ? mo : sz_in[d]; }
  si[1] = szi[0]; si[2] = szi[0] * szi[1];
  so[1] = szo[0]; so[2] = szo[0] * szo[1];
  for (int k = 0; k < szo[2]; ++k)
    for (int j = 0; j < szo[1]; ++j)
      for (int i = 0; i < szo[0]; ++i) {
        int idx[3] = {i, j, k};
        int o = idx[dir];
        double s = 0;
        for (int m = 0; m < mi; ++m) {
          idx[dir] = m;
          const double c = transpose ? M[m * mo + o] : M[o * mi + m];
          s += c * in[idx[0] * si[0] + idx[1] * si[1] + idx[2] * si[2]];
        }
        out[i * so[0] + j * so[1] + k * so[2]] = s;
      }
}

/* nodal Q2 values (3^dim) -> values at Gauss points (3^dim) */
static void q2_to_quad(int dim, const double *u, double *uq)
{
  double t[27], tmp[27]; int sz[3] = {3, 3, 3};
  memcpy(t, u, sizeof(double) * 27);
  for (int d = 0; d < dim; ++d) { sweep(&N2[0][0], 3, 3, 0, t, tmp, d, dim, sz); memcpy(t, tmp, sizeof(double) * 27); }
  memcpy(uq, t, sizeof(double) * 27);
}
static void quad_to_q2_T(int dim, const double *vq, double *v)
{
  double t[27], tmp[27]; int sz[3] = {3, 3, 3};
  memcpy(t, vq, sizeof(double) * 27);
  for (int d = 0; d < dim; ++d) { sweep(&N2[0][0], 3, 3, 1, t, tmp, d, dim, sz); memcpy(t, tmp, sizeof(double) * 27); }
  memcpy(v, t, sizeof(double) * 27);
}
/* nodal Q1 (2^dim) -> quad (3^dim) */
static void q1_to_quad(int dim, const double *p, double *pq)
{
  double t[27], tmp[27]; int sz[3] = {2, 2, 2};
  memcpy(t, p, sizeof(double) * (1 << dim));
  for (int d = 0; d < dim; ++d) { sweep(&N1[0][0], 3, 2, 0, t, tmp, d, dim, sz); sz[d] = 3; memcpy(t, tmp, sizeof(double) * 27); }
  memcpy(pq, t, sizeof(double) * 27);
}
static void quad_to_q1_T(int dim, const double *pq, double *p)
{
  double t[27], tmp[27]; int sz[3] = {3, 3, 3};
  memcpy(t, pq, sizeof(double) * 27);
  for (int d = 0; d < dim; ++d) { sweep(&N1[0][0], 2, 3, 1, t, tmp, d, dim, sz); sz[d] = 2; memcpy(t, tmp, sizeof(double) * 27); }
  memcpy(p, t, sizeof(double) * (1 << dim));
}
static void colloc_grad(int dim, const double *uq, int dir, double *g)
{ int sz[3] = {3, 3, 3}; sweep(&D2c[0][0], 3, 3, 0, uq, g, dir, dim, sz); }
static void colloc_grad_T_add(int dim, const double *s, int dir, double *acc)
{
  int sz[3] = {3, 3, 3}; double t[27];
  sweep(&D2c[0][0], 3, 3, 1, s, t, dir, dim, sz);
  for (int i = 0; i < 27; ++i) acc[i] += t[i];
}

static int ipow(int b, int e) { int r = 1; while (e--) r *= b; return r; }

static void quad_weights(int dim, const double *h, double *w)
{
  const int nq = ipow(3, dim);
  double vol = 1; for (int d = 0; d < dim; ++d) vol *= h[d];
  for (int q = 0; q < nq; ++q) {
    double ww = vol; int r = q;
    for (int d = 0; d < dim; ++d) { ww *= GW[r % 3]; r /= 3; }
    w[q] = ww;
  }
}

/* The velocity part of the cell kernel: given nodal u (component-major, dim*3^dim) and pressure values at
 * quadrature points pq (or NULL), compute out_u (dim*3^dim) and, if divq != NULL, div u at the quadrature points. */
static void cell_velocity(int dim, const double *h, const double *eta, int n_eta, int compressible,
                          const double *u, const double *pq, double ps, int do_viscous,
                          double *out_u, double *divq)
{
  const int nn = ipow(3, dim), nq = nn;
  double w[27]; quad_weights(dim, h, w);
  double G[3][3][27];             /* G[c][d][q] = d u_c / d x_d */
  if (do_viscous || divq)
    for (int c = 0; c < dim; ++c) {
      double uq[27]; q2_to_quad(dim, u + c * nn, uq);
      for (int d = 0; d < dim; ++d) {
        colloc_grad(dim, uq, d, G[c][d]);
        for (int q = 0; q < nq; ++q) G[c][d][q] /= h[d];
      }
    }
  double S[3][3][27];             /* stress to be tested with grad v */
  for (int q = 0; q < nq; ++q) {
    const double e2 = 2.0 * eta[n_eta == 1 ? 0 : q];
    double div = 0;
    if (do_viscous || divq) for (int c = 0; c < dim; ++c) div += G[c][c][q];
    if (divq) divq[q] = div;
    for (int c = 0; c < dim; ++c)
      for (int d = 0; d < dim; ++d) {
        double s = 0;
        if (do_viscous) s = e2 * 0.5 * (G[c][d][q] + G[d][c][q]);
        if (c == d) {
          if (pq) s -= ps * pq[q];
          if (do_viscous && compressible) s -= e2 / 3.0 * div;
        }
        S[c][d][q] = s * w[q] / h[d];
      }
  }
  for (int c = 0; c < dim; ++c) {
    double acc[27]; memset(acc, 0, sizeof acc);
    for (int d = 0; d < dim; ++d) colloc_grad_T_add(dim, S[c][d], d, acc);
    quad_to_q2_T(dim, acc, out_u + c * nn);
  }
}

typedef struct {
  int dim; int64_t n_cells;
  const int64_t *cell_u, *cell_p;     /* [cell][dim*3^dim], [cell][2^dim] */
  const double *h;                    /* dim */
  const double *eta; int n_eta;       /* [cell][n_eta] */
  double ps; int compressible;
  const uint8_t *mask_u;              /* 1 = constrained velocity DoF (may be NULL) */
  const int64_t *order; const int64_t *color_ptr; int n_colors;   /* cells sorted by colour */
} smf_cpu_level;

static void gather(const double *x, const int64_t *idx, int n, const uint8_t *mask, double *loc)
{ for (int i = 0; i < n; ++i) loc[i] = (mask && mask[idx[i]]) ? 0.0 : x[idx[i]]; }
static void scatter_add(double *y, const int64_t *idx, int n, const uint8_t *mask, const double *loc)
{ for (int i = 0; i < n; ++i) if (!(mask && mask[idx[i]])) y[idx[i]] += loc[i]; }

#define FOR_COLORED_CELLS(L, body)                                         \
  for (int col = 0; col < (L)->n_colors; ++col) {                          \
    const int64_t b = (L)->color_ptr[col], e = (L)->color_ptr[col + 1];    \
    _Pragma("omp parallel for schedule(static)")                           \
    for (int64_t ii = b; ii < e; ++ii) {                                   \
      const int64_t cell = (L)->order[ii];                                 \
      body                                                                 \
    }                                                                      \
  }

static void unit_diag(const smf_cpu_level *L, int64_t n_u, const double *src, double *dst)
{
  if (!L->mask_u) return;
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_u; ++i) if (L->mask_u[i]) dst[i] += src[i];
}

void smf_cpu_vmult_A(const smf_cpu_level *L, int64_t n_u, const double *src, double *dst)
{
  init_tables();
  const int dim = L->dim, nu = dim * ipow(3, dim);
  memset(dst, 0, sizeof(double) * n_u);
  FOR_COLORED_CELLS(L, {
    double u[81]; double out[81];
    gather(src, L->cell_u + cell * nu, nu, L->mask_u, u);
    cell_velocity(dim, L->h, L->eta + cell * L->n_eta, L->n_eta, L->compressible, u, NULL, 0.0, 1, out, NULL);
    scatter_add(dst, L->cell_u + cell * nu, nu, L->mask_u, out);
  })
  unit_diag(L, n_u, src, dst);
}

void smf_cpu_vmult_stokes(const smf_cpu_level *L, int64_t n_u, int64_t n_p,
                          const double *src_u, const double *src_p, double *dst_u, double *dst_p)
{
  init_tables();
  const int dim = L->dim, nu = dim * ipow(3, dim), np = 1 << dim, nq = ipow(3, dim);
  memset(dst_u, 0, sizeof(double) * n_u);
  memset(dst_p, 0, sizeof(double) * n_p);
  FOR_COLORED_CELLS(L, {
    double u[81]; double out[81]; double p[8]; double pq[27]; double divq[27]; double outp[8];
    gather(src_u, L->cell_u + cell * nu, nu, L->mask_u, u);
    gather(src_p, L->cell_p + cell * np, np, NULL, p);
    q1_to_quad(dim, p, pq);
    cell_velocity(dim, L->h, L->eta + cell * L->n_eta, L->n_eta, L->compressible, u, pq, L->ps, 1, out, divq);
    double w[27]; quad_weights(dim, L->h, w);
    for (int q = 0; q < nq; ++q) divq[q] *= -L->ps * w[q];
    quad_to_q1_T(dim, divq, outp);
    scatter_add(dst_u, L->cell_u + cell * nu, nu, L->mask_u, out);
    scatter_add(dst_p, L->cell_p + cell * np, np, NULL, outp);
  })
  unit_diag(L, n_u, src_u, dst_u);
}

void smf_cpu_vmult_BT(const smf_cpu_level *L, int64_t n_u, const double *src_p, double *dst_u)
{
  init_tables();
  const int dim = L->dim, nu = dim * ipow(3, dim), np = 1 << dim;
  memset(dst_u, 0, sizeof(double) * n_u);
  FOR_COLORED_CELLS(L, {
    double u[81]; double out[81]; double p[8]; double pq[27];
    memset(u, 0, sizeof u);
    gather(src_p, L->cell_p + cell * np, np, NULL, p);
    q1_to_quad(dim, p, pq);
    cell_velocity(dim, L->h, L->eta + cell * L->n_eta, L->n_eta, 0, u, pq, L->ps, 0, out, NULL);
    scatter_add(dst_u, L->cell_u + cell * nu, nu, L->mask_u, out);
  })
  /* Base::vmult adds src to dst at constrained DoFs of the block vector; BT's velocity source block is unused
     by the cell kernel, and the driver only ever calls it with a zero velocity block (block_stokes_preconditioner.h:245). */
}

void smf_cpu_vmult_mass(const smf_cpu_level *L, int64_t n_p, const double *src, double *dst)
{
  init_tables();
  const int dim = L->dim, np = 1 << dim, nq = ipow(3, dim);
  memset(dst, 0, sizeof(double) * n_p);
  FOR_COLORED_CELLS(L, {
    double p[8]; double pq[27]; double outp[8]; double w[27];
    gather(src, L->cell_p + cell * np, np, NULL, p);
    q1_to_quad(dim, p, pq);
    quad_weights(dim, L->h, w);
    const double *eta = L->eta + cell * L->n_eta;
    for (int q = 0; q < nq; ++q) pq[q] *= (L->ps * L->ps / eta[L->n_eta == 1 ? 0 : q]) * w[q];
    quad_to_q1_T(dim, pq, outp);
    scatter_add(dst, L->cell_p + cell * np, np, NULL, outp);
  })
}

/* MatrixFreeTools::compute_diagonal: apply the cell operator to every local unit vector (constrained
 * local DoFs contribute nothing), constrained entries := 1. */
void smf_cpu_diag_A(const smf_cpu_level *L, int64_t n_u, double *diag)
{
  init_tables();
  const int dim = L->dim, nu = dim * ipow(3, dim);
  memset(diag, 0, sizeof(double) * n_u);
  FOR_COLORED_CELLS(L, {
    double u[81]; double out[81]; double d[81];
    for (int i = 0; i < nu; ++i) {
      memset(u, 0, sizeof u); u[i] = 1.0;
      cell_velocity(dim, L->h, L->eta + cell * L->n_eta, L->n_eta, L->compressible, u, NULL, 0.0, 1, out, NULL);
      d[i] = out[i];
    }
    scatter_add(diag, L->cell_u + cell * nu, nu, L->mask_u, d);
  })
  if (L->mask_u) for (int64_t i = 0; i < n_u; ++i) if (L->mask_u[i]) diag[i] = 1.0;
}

void smf_cpu_diag_mass(const smf_cpu_level *L, int64_t n_p, double *diag)
{
  init_tables();
  const int dim = L->dim, np = 1 << dim, nq = ipow(3, dim);
  memset(diag, 0, sizeof(double) * n_p);
  FOR_COLORED_CELLS(L, {
    double p[8]; double pq[27]; double outp[8]; double d[8]; double w[27];
    quad_weights(dim, L->h, w);
    const double *eta = L->eta + cell * L->n_eta;
    for (int i = 0; i < np; ++i) {
      memset(p, 0, sizeof p); p[i] = 1.0;
      q1_to_quad(dim, p, pq);
      for (int q = 0; q < nq; ++q) pq[q] *= (L->ps * L->ps / eta[L->n_eta == 1 ? 0 : q]) * w[q];
      quad_to_q1_T(dim, pq, outp);
      d[i] = outp[i];
    }
    scatter_add(diag, L->cell_p + cell * np, np, NULL, d);
  })
}

/* ---- MGTransferMatrixFree on a uniformly refined mesh -------------------------------------------
 * children[parent][2^dim] = fine cell ids; cell maps of both levels; npd = nodes per direction (3 or 2);
 * ncomp components (component-major cell maps).  weights[fine dof] = 1/valence.
 * prolongate: fine = P coarse (coarse Dirichlet DoFs read as zero); restrict: coarse += P^T fine, then
 * coarse Dirichlet DoFs zeroed by the caller-provided mask. */
static const double P2[5][3] = {{1, 0, 0}, {0.375, 0.75, -0.125}, {0, 1, 0}, {-0.125, 0.75, 0.375}, {0, 0, 1}};
static const double P1[3][2] = {{1, 0}, {0.5, 0.5}, {0, 1}};

static void patch_index(int dim, int npd, int64_t parent, const int64_t *children, const int64_t *cell_f, int ncomp, int comp, int64_t *patch)
{
  const int np = 2 * npd - 1, nn = ipow(npd, dim), nloc = ncomp * nn;
  for (int k = 0; k < (1 << dim); ++k) {
    const int64_t child = children[parent * (1 << dim) + k];
    for (int a = 0; a < nn; ++a) {
      int r = a, pi = 0, mul = 1;
      for (int d = 0; d < dim; ++d) { int loc = r % npd; r /= npd; pi += (loc + (npd - 1) * ((k >> d) & 1)) * mul; mul *= np; }
      patch[pi] = cell_f[child * nloc + comp * nn + a];
    }
  }
}

void smf_cpu_prolongate(int dim, int npd, int ncomp, int64_t n_parents, const int64_t *children,
                        const int64_t *cell_c, const int64_t *cell_f, const double *weights,
                        const uint8_t *mask_c, int64_t n_fine, const double *src_c, double *dst_f)
{
  const int np = 2 * npd - 1, nn = ipow(npd, dim), npatch = ipow(np, dim), nloc = ncomp * nn;
  const double *P = (npd == 3) ? &P2[0][0] : &P1[0][0];
  memset(dst_f, 0, sizeof(double) * n_fine);
  /* serial over parents (deterministic; this is test infrastructure) */
  for (int64_t par = 0; par < n_parents; ++par)
    for (int c = 0; c < ncomp; ++c) {
      double t[125], tmp[125]; int64_t patch[125]; int sz[3] = {npd, npd, npd};
      gather(src_c, cell_c + par * nloc + c * nn, nn, mask_c, t);
      for (int d = 0; d < dim; ++d) { sweep(P, np, npd, 0, t, tmp, d, dim, sz); sz[d] = np; memcpy(t, tmp, sizeof(double) * npatch); }
      patch_index(dim, npd, par, children, cell_f, ncomp, c, patch);
      for (int i = 0; i < npatch; ++i) dst_f[patch[i]] += weights[patch[i]] * t[i];
    }
}

void smf_cpu_restrict_add(int dim, int npd, int ncomp, int64_t n_parents, const int64_t *children,
                          const int64_t *cell_c, const int64_t *cell_f, const double *weights,
                          const uint8_t *mask_c, const double *src_f, double *dst_c)
{
  const int np = 2 * npd - 1, nn = ipow(npd, dim), npatch = ipow(np, dim), nloc = ncomp * nn;
  const double *P = (npd == 3) ? &P2[0][0] : &P1[0][0];
  for (int64_t par = 0; par < n_parents; ++par)
    for (int c = 0; c < ncomp; ++c) {
      double t[125], tmp[125]; int64_t patch[125]; int sz[3] = {np, np, np};
      patch_index(dim, npd, par, children, cell_f, ncomp, c, patch);
      for (int i = 0; i < npatch; ++i) t[i] = weights[patch[i]] * src_f[patch[i]];
      for (int d = 0; d < dim; ++d) { sweep(P, npd, np, 1, t, tmp, d, dim, sz); sz[d] = npd; memcpy(t, tmp, sizeof(double) * npatch); }
      scatter_add(dst_c, cell_c + par * nloc + c * nn, nn, mask_c, t);
    }
}

int smf_cpu_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
