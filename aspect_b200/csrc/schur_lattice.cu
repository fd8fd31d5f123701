// Lattice path of the Schur-complement V-cycle (prec_Schur: /root/reference/source/simulator/solver/
// stokes_matrix_free_local_smoothing.cc:803-849 smoothers, :893-916 PreconditionMG; operator = MassMatrixOperator,
// matrix_free_operators.cc:787-845).
//
// The Schur V-cycle applies the pressure mass matrix 32 times per level and cycle (4 + 4 Chebyshev(4) sweeps); with
// cell-batch kernels every application is a gather / scatter with REDs on the 88 % of a brick's pressure nodes that it
// shares with other bricks, followed by a separate vector update: at refinement 7 that was 256 launch-bound kernels
// and 3.4 ms per cycle for vectors that fit the L2 cache many times over.
//
// On a level whose cells form a Cartesian box (every uniformly refined level of a box mesh; detected here from the
// cell -> DoF maps alone, no numbering is assumed) the pressure nodes are an nx x ny x nz lattice.  The V-cycle is
// self-contained -- only its defect comes in and its result goes out -- so inside it all level vectors live in lattice
// order, and
//   * the mass matrix with one viscosity per cell, sum_cells (ps^2 vol / eta_c) M1 x M1 x M1, becomes a node-centred
//     27-point stencil: thread (X, Y) marches along Z, keeps the x-y partial sums of three node planes in registers
//     (9 loads + 4 coefficient loads per node instead of 27 + 8), sums in a fixed order and writes each node once;
//   * the Chebyshev update xo <- x + f1 (x - xo) + f2 D^-1 (b - M x) and the residual b - M x are the epilogue of
//     that kernel (PreconditionChebyshev's vector_updates) -- one launch per smoothing step, no
//     zero fill, no REDs;
//   * the transfers are the trilinear stencils of the Q1 embedding (MGTransferMatrixFree for FE_Q(1)): fine node
//     (X, Y, Z) <- coarse (X/2, Y/2, Z/2) with weights 1 / 1/2 per direction, restriction its transpose;
//   * the whole cycle is one CUDA graph.
// On several ranks every rank's box is its own lattice; nodes shared with other ranks keep partial sums, go through
// the level's halo exchange (the same lists, mapped to lattice indices) and are finished by a small list kernel.
#include "smf_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

using namespace smf;

namespace smf {

#define TRY(x) do { int e__ = (x); if (e__) return e__; } while (0)

struct LatDims { int nx, ny, nz; };
// coarse side of a transfer: extent of the region under the fine lattice, pitches and origin of the array it lives in
// (the rank's own coarse lattice: pitch = extent, origin 0; a replicated global lattice: the rank's box inside it)
struct LatBox { int nx, ny, nz, px, py, ox, oy, oz; };
__host__ __device__ inline int64_t lb_index(const LatBox &b, int x, int y, int z)
{ return ((int64_t)(z + b.oz) * b.py + (y + b.oy)) * b.px + (x + b.ox); }

// ---- kernels ---------------------------------------------------------------------------------------------------
__global__ void k_lat_coef(double *coef, const double *eta_t, const uint32_t *cell2lat, int64_t n_cells, double scale)
{
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_cells) coef[cell2lat[c]] = scale / eta_t[c];      // (n_eta = 1: the lane-transposed table is the plain cell order)
}
__global__ void k_lat_scatter(double *lat, const double *vec, const uint32_t *to_lat, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lat[to_lat[i]] = vec[i];
}
__global__ void k_lat_gather(double *vec, const double *lat, const uint32_t *to_lat, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vec[i] = lat[to_lat[i]];
}

// replicated levels: this rank's coefficients into the global table (the other ranks' cells stay zero until the all-reduce)
__global__ void k_lat_coef_global(double *gcoef, const double *coef, LatDims cells, LatBox g)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)cells.nx * cells.ny * cells.nz) return;
  const int X = (int)(i % cells.nx), Y = (int)((i / cells.nx) % cells.ny), Z = (int)(i / ((int64_t)cells.nx * cells.ny));
  gcoef[lb_index(g, X, Y, Z)] = coef[i];
}
// inverse diagonal of the lattice mass matrix in closed form: 1 / (8 x sum of the coefficients of the cells around the node)
__global__ void k_lat_dinv(double *dinv, const double *coef, LatDims d)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)d.nx * d.ny * d.nz) return;
  const int X = (int)(i % d.nx), Y = (int)((i / d.nx) % d.ny), Z = (int)(i / ((int64_t)d.nx * d.ny));
  const int cxn = d.nx - 1, cyn = d.ny - 1;
  double ks = 0.0;
  for (int cz = max(Z - 1, 0); cz <= min(Z, d.nz - 2); ++cz)
    for (int cy = max(Y - 1, 0); cy <= min(Y, cyn - 1); ++cy)
      for (int cx = max(X - 1, 0); cx <= min(X, cxn - 1); ++cx) ks += coef[((int64_t)cz * cyn + cy) * cxn + cx];
  dinv[i] = 1.0 / (8.0 * ks);
}

// MODE 0: t = M x;  1: xo <- x + f1 (x - xo | x) + f2 dinv (b - M x);  2: t = b - M x.
// Nodes with mask bit 0 (shared with another rank) get t = (this rank's partial) M x in every mode.
struct LatStep {
  LatDims d;
  const double *x, *coef, *dinv, *b;
  double *xo, *t;
  const unsigned char *mask;
  double f1, f2;
  int use_xo, zpt;
};
constexpr int LAT_THREADS = 128;
// The march is a chain of dependent memory round trips (ncu of the first version: long scoreboard 10.7 per issue, nothing
// else above 25 %), so everything iteration Z + 1 needs -- the nine values of node plane Z + 2, four coefficients, b, xo,
// the mask -- is loaded at the top of iteration Z, before anything of iteration Z is consumed: one round trip per node.
template <int MODE>
__global__ void __launch_bounds__(LAT_THREADS, 4) k_lat_step(const LatStep P)
{
  const int nx = P.d.nx, ny = P.d.ny, nz = P.d.nz, cx_n = nx - 1, cy_n = ny - 1;
  const int col = blockIdx.x * LAT_THREADS + threadIdx.x;
  if (col >= nx * ny) return;
  const double *__restrict__ xs = P.x;
  const double *__restrict__ cf = P.coef;
  const double *__restrict__ bs = P.b;
  const unsigned char *__restrict__ mk = P.mask;
  const int Y = col / nx, X = col - Y * nx;
  const int Z0 = blockIdx.y * P.zpt, Z1 = min(Z0 + P.zpt, nz);
  // clamped neighbour columns; cells outside the box have coefficient zero
  const int Xm = max(X - 1, 0), Xp = min(X + 1, nx - 1), Ym = max(Y - 1, 0), Yp = min(Y + 1, ny - 1);
  const bool cxm = X > 0, cxp = X < nx - 1, cym = Y > 0, cyp = Y < ny - 1;
  // (32-bit indices throughout: a level has fewer than 2^28 pressure DoFs; the first version spent 160 of its 290 instructions
  //  per node on 64-bit address arithmetic and ran at 56 % issue-slot utilisation)
  const int pl = nx * ny, cpl = cx_n * cy_n;
  const int o_mm = Xm + nx * Ym, o_0m = X + nx * Ym, o_pm = Xp + nx * Ym;
  const int o_m0 = Xm + nx * Y, o_00 = X + nx * Y, o_p0 = Xp + nx * Y;
  const int o_mp = Xm + nx * Yp, o_0p = X + nx * Yp, o_pp = Xp + nx * Yp;
  // cells (X-1|X, Y-1|Y): indices in a cell plane (clamped; masked by the booleans above)
  const int c_mm = max(X - 1, 0) + cx_n * max(Y - 1, 0), c_pm = min(X, cx_n - 1) + cx_n * max(Y - 1, 0);
  const int c_mp = max(X - 1, 0) + cx_n * min(Y, cy_n - 1), c_pp = min(X, cx_n - 1) + cx_n * min(Y, cy_n - 1);
  const auto load_plane = [&](int Z, double (&a)[9]) {
    long long zo = (long long)pl * min(max(Z, 0), nz - 1);
    asm volatile("" : "+l"(zo));     // (opaque plane offset: the plane pointer is formed once, every load is one IMAD.WIDE + LDG instead of a 64-bit add chain)
    const double *__restrict__ p = xs + zo;
    a[0] = p[o_mm]; a[1] = p[o_0m]; a[2] = p[o_pm]; a[3] = p[o_m0]; a[4] = p[o_00]; a[5] = p[o_p0]; a[6] = p[o_mp]; a[7] = p[o_0p]; a[8] = p[o_pp];
  };
  // q[k]: x-y partial sums of a node plane for cell column k (mm, pm, mp, pp), weights 2 towards (X, Y)
  const auto plane_sums = [&](const double (&a)[9], double (&q)[4]) {
    const double rm_lo = fma(2.0, a[1], a[0]), rm_hi = fma(2.0, a[1], a[2]);      // row Y-1: cell columns X-1 / X
    const double r0_lo = fma(2.0, a[4], a[3]), r0_hi = fma(2.0, a[4], a[5]);      // row Y
    const double rp_lo = fma(2.0, a[7], a[6]), rp_hi = fma(2.0, a[7], a[8]);      // row Y+1
    q[0] = fma(2.0, r0_lo, rm_lo); q[1] = fma(2.0, r0_hi, rm_hi);
    q[2] = fma(2.0, r0_lo, rp_lo); q[3] = fma(2.0, r0_hi, rp_hi);
  };
  const auto load_coefs = [&](int cz, double (&k)[4]) {
    const bool in = cz >= 0 && cz < nz - 1;
    long long co = (long long)cpl * min(max(cz, 0), nz - 2);
    asm volatile("" : "+l"(co));
    const double *__restrict__ c = cf + co;
    k[0] = (in && cxm && cym) ? c[c_mm] : 0.0; k[1] = (in && cxp && cym) ? c[c_pm] : 0.0;
    k[2] = (in && cxm && cyp) ? c[c_mp] : 0.0; k[3] = (in && cxp && cyp) ? c[c_pp] : 0.0;
  };
  // q[s][k]: plane sums of three consecutive node planes, kc[s][k]: coefficients of three consecutive cell layers; the roles
  // (previous / current / next) rotate through the slots with the phase of the step, so the march is unrolled three times and no
  // register is ever copied (the rotating copies were 40 of the ~240 instructions per node)
  double a0[9], a1[9], an[9], q[3][4], kc[3][4];
  load_plane(Z0 - 1, a0); load_plane(Z0, a1); load_plane(Z0 + 1, an);
  load_coefs(Z0 - 1, kc[0]); load_coefs(Z0, kc[1]);
  int i = o_00 + pl * Z0;
  double bn = MODE != 0 ? bs[i] : 0.0, xon = (MODE == 1 && P.use_xo) ? P.xo[i] : 0.0;
  unsigned char mn = mk ? mk[i] : 0;
  plane_sums(a0, q[0]); plane_sums(a1, q[1]);
  double xc1 = a1[4];
  // one node: phase PH = (Z - Z0) mod 3; slots: planes Z-1, Z, Z+1 -> q[PH], q[PH+1], q[PH+2]; layers Z-1, Z -> kc[PH], kc[PH+1];
  // the layer prefetched for the next node goes to kc[PH+2]
  const auto step = [&](auto phase, int Z) {
    constexpr int PH = decltype(phase)::value, S0 = PH % 3, S1 = (PH + 1) % 3, S2 = (PH + 2) % 3;
    plane_sums(an, q[S2]);
    const double xc2 = an[4], bv = bn, xov = xon;
    const bool shared = mn & 1;
    if (Z + 1 < Z1) {      // start the loads of the next node before anything of this one is consumed
      load_plane(Z + 2, an);
      load_coefs(Z + 1, kc[S2]);
      if (MODE != 0) bn = bs[i + pl];
      if (MODE == 1 && P.use_xo) xon = P.xo[i + pl];
      if (mk) mn = mk[i + pl];
    }
    double acc = 0.0, ksum = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc = fma(kc[S0][k], fma(2.0, q[S1][k], q[S0][k]), acc);      // cell layer Z-1: node planes Z-1, Z
      acc = fma(kc[S1][k], fma(2.0, q[S1][k], q[S2][k]), acc);      // cell layer Z:   node planes Z, Z+1
      ksum += kc[S0][k] + kc[S1][k];
    }
    if (MODE == 0 || shared) P.t[i] = acc;
    else if (MODE == 2) P.t[i] = bv - acc;
    else {
      // diagonal of the mass matrix = 8 x (sum of the coefficients of the cells around the node): no D^-1 stream
      // (MassMatrixOperator::compute_diagonal, matrix_free_operators.cc:861-908, in closed form); reciprocal + multiply:
      // a quarter of the instructions of an IEEE division
      double v = xc1 + P.f2 * (bv - acc) * __drcp_rn(8.0 * ksum);
      if (P.use_xo) v += P.f1 * (xc1 - xov);
      else v += P.f1 * xc1;
      P.xo[i] = v;
    }
    xc1 = xc2;
    i += pl;
  };
  for (int Z = Z0; Z < Z1; Z += 3) {
    step(std::integral_constant<int, 0>(), Z);
    if (Z + 1 < Z1) step(std::integral_constant<int, 1>(), Z + 1);
    if (Z + 2 < Z1) step(std::integral_constant<int, 2>(), Z + 2);
  }
}
// the interface nodes after the exchange: t holds the complete M x
template <int MODE>
__global__ void k_lat_finish(const uint32_t *nodes, int64_t n, const double *x, double *xo, const double *dinv, const double *b, double *t,
                             double f1, double f2, int use_xo)
{
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t i = nodes[j];
  if (MODE == 2) t[i] = b[i] - t[i];
  else {
    const double xi = x[i];
    double v = xi + f2 * dinv[i] * (b[i] - t[i]);
    if (use_xo) v += f1 * (xi - xo[i]);
    else v += f1 * xi;
    xo[i] = v;
  }
}
// coarse defect = P^T (fine residual): coarse node (X, Y, Z) gathers the 27 fine nodes around (2X, 2Y, 2Z);
// fine nodes owned by a lower rank (mask bit 1) are counted there
__global__ void k_lat_restrict(double *coarse, const double *fine, const unsigned char *fine_mask, LatBox c, LatDims f)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)c.nx * c.ny * c.nz) return;
  const int X = (int)(i % c.nx), Y = (int)((i / c.nx) % c.ny), Z = (int)(i / ((int64_t)c.nx * c.ny));
  double acc = 0.0;
  for (int dz = -1; dz <= 1; ++dz) {
    const int fz = 2 * Z + dz;
    if (fz < 0 || fz >= f.nz) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int fy = 2 * Y + dy;
      if (fy < 0 || fy >= f.ny) continue;
      const double wyz = (dz ? 0.5 : 1.0) * (dy ? 0.5 : 1.0);
      const int64_t row = ((int64_t)fz * f.ny + fy) * f.nx;
      double r = 0.0;
      for (int dx = -1; dx <= 1; ++dx) {
        const int fx = 2 * X + dx;
        if (fx < 0 || fx >= f.nx) continue;
        if (fine_mask && (fine_mask[row + fx] & 2)) continue;
        r = fma(dx ? 0.5 : 1.0, fine[row + fx], r);
      }
      acc = fma(wyz, r, acc);
    }
  }
  coarse[lb_index(c, X, Y, Z)] = acc;
}
// fine += P coarse
__global__ void k_lat_prolong_add(double *fine, const double *coarse, LatBox c, LatDims f)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)f.nx * f.ny * f.nz) return;
  const int X = (int)(i % f.nx), Y = (int)((i / f.nx) % f.ny), Z = (int)(i / ((int64_t)f.nx * f.ny));
  const int x0 = X >> 1, x1 = (X + 1) >> 1, y0 = Y >> 1, y1 = (Y + 1) >> 1, z0 = Z >> 1, z1 = (Z + 1) >> 1;
  const auto at = [&](int a, int b, int cc) { return coarse[lb_index(c, a, b, cc)]; };
  const auto linex = [&](int b, int cc) { return x0 == x1 ? at(x0, b, cc) : 0.5 * (at(x0, b, cc) + at(x1, b, cc)); };
  const auto faceyz = [&](int cc) { return y0 == y1 ? linex(y0, cc) : 0.5 * (linex(y0, cc) + linex(y1, cc)); };
  const double v = z0 == z1 ? faceyz(z0) : 0.5 * (faceyz(z0) + faceyz(z1));
  fine[i] += v;
}

// ---- coarse tail: levels 0..top of the cycle in ONE thread block -------------------------------------------------------
// Below ~1,000 nodes a level's 35 launches are pure latency (three dependent trips to L2 per launch).  Here the level
// vectors, coefficients and inverse diagonals of levels 0..top live in shared memory and block barriers replace the kernel
// boundaries; same smoothers, same transfers, one graph node instead of ~35 per level.
constexpr int LT_MAX_LEVELS = 6;
constexpr int LT_THREADS = 512;
struct LatTailLevel { int nx, ny, nz, off; double theta, delta; int degree; const double *coef; };      // off: start of the level's block in shared memory (doubles)
struct LatTail { int n_levels, steps; LatTailLevel lv[LT_MAX_LEVELS]; const double *b_in; double *x_out; };
__host__ __device__ inline int lt_nodes(const LatTailLevel &L) { return L.nx * L.ny * L.nz; }
__host__ __device__ inline int lt_cells(const LatTailLevel &L) { return (L.nx - 1) * (L.ny - 1) * (L.nz - 1); }
__host__ __device__ inline int lt_doubles(const LatTailLevel &L) { return 4 * lt_nodes(L) + lt_cells(L); }      // x, xo, b, dinv, coef

// xo <- x + f1 (x - xo | x) + f2 dinv (b - M x) on one level in shared memory (mode 1), or xo <- b - M x (mode 2)
__device__ void lt_step(const LatTailLevel &L, const double *x, double *xo, const double *b, const double *dinv, const double *coef,
                        double f1, double f2, int use_xo, int mode)
{
  const int nx = L.nx, ny = L.ny, nz = L.nz, n = nx * ny * nz, cxn = nx - 1, cyn = ny - 1;
  for (int i = threadIdx.x; i < n; i += LT_THREADS) {
    const int X = i % nx, Y = (i / nx) % ny, Z = i / (nx * ny);
    // the formulation of k_lat_step without the march: x-y partial sums of the three node planes (clamped indices, all
    // loads independent), cells outside the box have coefficient zero
    const int Xm = max(X - 1, 0), Xp = min(X + 1, nx - 1), Ym = max(Y - 1, 0), Yp = min(Y + 1, ny - 1);
    const bool cxm = X > 0, cxp = X < nx - 1, cym = Y > 0, cyp = Y < ny - 1;
    double q[3][4];
#pragma unroll
    for (int dz = 0; dz < 3; ++dz) {
      const double *p = x + min(max(Z + dz - 1, 0), nz - 1) * nx * ny;
      const double a_mm = p[Xm + nx * Ym], a_0m = p[X + nx * Ym], a_pm = p[Xp + nx * Ym];
      const double a_m0 = p[Xm + nx * Y], a_00 = p[X + nx * Y], a_p0 = p[Xp + nx * Y];
      const double a_mp = p[Xm + nx * Yp], a_0p = p[X + nx * Yp], a_pp = p[Xp + nx * Yp];
      const double rm_lo = fma(2.0, a_0m, a_mm), rm_hi = fma(2.0, a_0m, a_pm);
      const double r0_lo = fma(2.0, a_00, a_m0), r0_hi = fma(2.0, a_00, a_p0);
      const double rp_lo = fma(2.0, a_0p, a_mp), rp_hi = fma(2.0, a_0p, a_pp);
      q[dz][0] = fma(2.0, r0_lo, rm_lo); q[dz][1] = fma(2.0, r0_hi, rm_hi);
      q[dz][2] = fma(2.0, r0_lo, rp_lo); q[dz][3] = fma(2.0, r0_hi, rp_hi);
    }
    double k[2][4];
#pragma unroll
    for (int dz = 0; dz < 2; ++dz) {
      const int cz = Z + dz - 1;
      const bool in = cz >= 0 && cz < nz - 1;
      const double *c = coef + min(max(cz, 0), nz - 2) * cxn * cyn;
      k[dz][0] = (in && cxm && cym) ? c[max(X - 1, 0) + cxn * max(Y - 1, 0)] : 0.0;
      k[dz][1] = (in && cxp && cym) ? c[min(X, cxn - 1) + cxn * max(Y - 1, 0)] : 0.0;
      k[dz][2] = (in && cxm && cyp) ? c[max(X - 1, 0) + cxn * min(Y, cyn - 1)] : 0.0;
      k[dz][3] = (in && cxp && cyp) ? c[min(X, cxn - 1) + cxn * min(Y, cyn - 1)] : 0.0;
    }
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc = fma(k[0][j], fma(2.0, q[1][j], q[0][j]), acc);
      acc = fma(k[1][j], fma(2.0, q[1][j], q[2][j]), acc);
    }
    if (mode == 2) xo[i] = b[i] - acc;
    else {
      const double xi = x[i];
      double v = xi + f2 * dinv[i] * (b[i] - acc);
      v += use_xo ? f1 * (xi - xo[i]) : f1 * xi;
      xo[i] = v;
    }
  }
  __syncthreads();
}
__device__ void lt_cheb_tail(const LatTailLevel &L, double *&x, double *&xo, const double *b, const double *dinv, const double *coef, bool have_old)
{
  if (L.degree < 2 || fabs(L.delta) < 1e-40) return;
  double rhok = L.delta / L.theta;
  const double sigma = L.theta / L.delta;
  for (int k = 0; k < L.degree - 1; ++k) {
    const double rhokp = 1.0 / (2.0 * sigma - rhok);
    const double f1 = rhokp * rhok, f2 = 2.0 * rhokp / L.delta;
    rhok = rhokp;
    lt_step(L, x, xo, b, dinv, coef, f1, f2, have_old, 1);
    double *t = x; x = xo; xo = t;
    have_old = true;
  }
}
__device__ void lt_smooth(const LatTailLevel &L, double *&x, double *&xo, const double *b, const double *dinv, const double *coef, int steps, bool zero_start)
{
  const int n = lt_nodes(L);
  int s = 0;
  if (zero_start) {      // PreconditionChebyshev::vmult
    const double it = 1.0 / L.theta;
    for (int i = threadIdx.x; i < n; i += LT_THREADS) x[i] = it * dinv[i] * b[i];
    __syncthreads();
    lt_cheb_tail(L, x, xo, b, dinv, coef, false);
    s = 1;
  }
  for (; s < steps; ++s) {      // PreconditionChebyshev::step
    lt_step(L, x, xo, b, dinv, coef, 0.0, 1.0 / L.theta, 0, 1);
    double *t = x; x = xo; xo = t;
    lt_cheb_tail(L, x, xo, b, dinv, coef, true);
  }
}
__global__ void __launch_bounds__(LT_THREADS) k_lat_tail(const LatTail P)
{
  extern __shared__ double lt_smem[];
  __shared__ double *s_res[LT_MAX_LEVELS];
  const int top = P.n_levels - 1;
  // load: coefficients of every level, defect of the top level; inverse diagonals in closed form
  for (int l = 0; l <= top; ++l) {
    const LatTailLevel &L = P.lv[l];
    const int n = lt_nodes(L), nc = lt_cells(L);
    double *base = lt_smem + L.off, *coef = base + 4 * n;
    for (int i = threadIdx.x; i < nc; i += LT_THREADS) coef[i] = L.coef[i];
    if (l == top) for (int i = threadIdx.x; i < n; i += LT_THREADS) base[2 * n + i] = P.b_in[i];
  }
  __syncthreads();
  for (int l = 0; l <= top; ++l) {
    const LatTailLevel &L = P.lv[l];
    const int nx = L.nx, ny = L.ny, nz = L.nz, n = nx * ny * nz, cxn = nx - 1, cyn = ny - 1;
    double *base = lt_smem + L.off, *dinv = base + 3 * n;
    const double *coef = base + 4 * n;
    for (int i = threadIdx.x; i < n; i += LT_THREADS) {
      const int X = i % nx, Y = (i / nx) % ny, Z = i / (nx * ny);
      double ks = 0.0;
      for (int cz = max(Z - 1, 0); cz <= min(Z, nz - 2); ++cz)
        for (int cy = max(Y - 1, 0); cy <= min(Y, cyn - 1); ++cy)
          for (int cx = max(X - 1, 0); cx <= min(X, cxn - 1); ++cx) ks += coef[(cz * cyn + cy) * cxn + cx];
      dinv[i] = 1.0 / (8.0 * ks);
    }
  }
  __syncthreads();
  // down: pre-smoothing (on level 0: the coarse solver), residual, restriction
  for (int l = top; l >= 0; --l) {
    const LatTailLevel &L = P.lv[l];
    const int n = lt_nodes(L);
    double *base = lt_smem + L.off, *x = base, *xo = base + n, *b = base + 2 * n, *dinv = base + 3 * n, *coef = base + 4 * n;
    lt_smooth(L, x, xo, b, dinv, coef, P.steps, true);
    if (l > 0) {
      lt_step(L, x, xo, b, dinv, coef, 0.0, 0.0, 0, 2);       // xo = defect - M x (xo is free between the sweeps)
      const LatTailLevel &C = P.lv[l - 1];
      const int cn = lt_nodes(C);
      double *cb = lt_smem + C.off + 2 * cn;
      for (int i = threadIdx.x; i < cn; i += LT_THREADS) {
        const int X = i % C.nx, Y = (i / C.nx) % C.ny, Z = i / (C.nx * C.ny);
        double acc = 0.0;
        for (int dz = -1; dz <= 1; ++dz) {
          const int fz = 2 * Z + dz;
          if (fz < 0 || fz >= L.nz) continue;
          for (int dy = -1; dy <= 1; ++dy) {
            const int fy = 2 * Y + dy;
            if (fy < 0 || fy >= L.ny) continue;
            const double wyz = (dz ? 0.5 : 1.0) * (dy ? 0.5 : 1.0);
            double r = 0.0;
            for (int dx = -1; dx <= 1; ++dx) {
              const int fx = 2 * X + dx;
              if (fx < 0 || fx >= L.nx) continue;
              r = fma(dx ? 0.5 : 1.0, xo[(fz * L.ny + fy) * L.nx + fx], r);
            }
            acc = fma(wyz, r, acc);
          }
        }
        cb[i] = acc;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) s_res[l] = x;
    __syncthreads();
  }
  // up: prolongation, post-smoothing
  for (int l = 1; l <= top; ++l) {
    const LatTailLevel &L = P.lv[l], &C = P.lv[l - 1];
    const int n = lt_nodes(L);
    double *base = lt_smem + L.off, *x = s_res[l], *xo = (x == base ? base + n : base), *b = base + 2 * n, *dinv = base + 3 * n, *coef = base + 4 * n;
    const double *cx_ = s_res[l - 1];
    for (int i = threadIdx.x; i < n; i += LT_THREADS) {
      const int X = i % L.nx, Y = (i / L.nx) % L.ny, Z = i / (L.nx * L.ny);
      const int x0 = X >> 1, x1 = (X + 1) >> 1, y0 = Y >> 1, y1 = (Y + 1) >> 1, z0 = Z >> 1, z1 = (Z + 1) >> 1;
      const auto at = [&](int a, int bb, int cc) { return cx_[(cc * C.ny + bb) * C.nx + a]; };
      const auto linex = [&](int bb, int cc) { return x0 == x1 ? at(x0, bb, cc) : 0.5 * (at(x0, bb, cc) + at(x1, bb, cc)); };
      const auto faceyz = [&](int cc) { return y0 == y1 ? linex(y0, cc) : 0.5 * (linex(y0, cc) + linex(y1, cc)); };
      x[i] += z0 == z1 ? faceyz(z0) : 0.5 * (faceyz(z0) + faceyz(z1));
    }
    __syncthreads();
    lt_smooth(L, x, xo, b, dinv, coef, P.steps, false);
    if (threadIdx.x == 0) s_res[l] = x;
    __syncthreads();
  }
  const double *res = s_res[top];
  for (int i = threadIdx.x; i < lt_nodes(P.lv[top]); i += LT_THREADS) P.x_out[i] = res[i];
}

// ---- host: lattice detection -----------------------------------------------------------------------------------
// Coordinates of the pressure nodes from the cell -> vertex maps alone: local vertices (i, i + 2^d) of a cell are
// neighbours along direction d; walking those chains from the nodes without a predecessor numbers every direction.
struct LatHost {
  int nx = 0, ny = 0, nz = 0;
  std::vector<uint32_t> dof2lat, cell2lat;
  std::vector<int> cell_xyz;       // [n_cells][3]
};
static bool detect_lattice(const std::vector<uint32_t> &nodes_p, int64_t n_cells, int64_t n_p, LatHost &H)
{
  const uint32_t NONE = 0xffffffffu;
  std::vector<int> coord[3];
  for (int d = 0; d < 3; ++d) {
    std::vector<uint32_t> nxt((size_t)n_p, NONE);
    std::vector<unsigned char> has_prev((size_t)n_p, 0);
    const int S = 1 << d;
    for (int64_t c = 0; c < n_cells; ++c)
      for (int i = 0; i < 8; ++i)
        if (!(i & S)) {
          const uint32_t a = nodes_p[c * 8 + i] & NODE_IDX_MASK, b = nodes_p[c * 8 + i + S] & NODE_IDX_MASK;
          if (a == b) return false;
          if (nxt[a] == NONE) nxt[a] = b;
          else if (nxt[a] != b) return false;
          has_prev[b] = 1;
        }
    coord[d].assign((size_t)n_p, -1);
    for (int64_t v = 0; v < n_p; ++v)
      if (!has_prev[v]) {
        int k = 0;
        for (uint32_t u = (uint32_t)v; u != NONE; u = nxt[u]) {
          if (coord[d][u] >= 0 || k > n_p) return false;
          coord[d][u] = k++;
        }
      }
    for (int64_t v = 0; v < n_p; ++v) if (coord[d][v] < 0) return false;      // (a node on a cycle, or touched by no cell)
  }
  int mx[3] = {0, 0, 0};
  for (int d = 0; d < 3; ++d) for (int64_t v = 0; v < n_p; ++v) mx[d] = std::max(mx[d], coord[d][v]);
  H.nx = mx[0] + 1; H.ny = mx[1] + 1; H.nz = mx[2] + 1;
  if ((int64_t)H.nx * H.ny * H.nz != n_p || (int64_t)(H.nx - 1) * (H.ny - 1) * (H.nz - 1) != n_cells) return false;
  H.dof2lat.assign((size_t)n_p, 0);
  std::vector<unsigned char> seen((size_t)n_p, 0);
  for (int64_t v = 0; v < n_p; ++v) {
    const int64_t li = coord[0][v] + (int64_t)H.nx * (coord[1][v] + (int64_t)H.ny * coord[2][v]);
    if (seen[li]) return false;
    seen[li] = 1;
    H.dof2lat[v] = (uint32_t)li;
  }
  H.cell2lat.assign((size_t)n_cells, 0);
  H.cell_xyz.assign((size_t)n_cells * 3, 0);
  std::vector<unsigned char> cseen((size_t)n_cells, 0);
  for (int64_t c = 0; c < n_cells; ++c) {
    const uint32_t v0 = nodes_p[c * 8] & NODE_IDX_MASK;
    const int cx = coord[0][v0], cy = coord[1][v0], cz = coord[2][v0];
    for (int i = 0; i < 8; ++i) {
      const uint32_t v = nodes_p[c * 8 + i] & NODE_IDX_MASK;
      if (coord[0][v] != cx + (i & 1) || coord[1][v] != cy + ((i >> 1) & 1) || coord[2][v] != cz + (i >> 2)) return false;
    }
    const int64_t li = cx + (int64_t)(H.nx - 1) * (cy + (int64_t)(H.ny - 1) * cz);
    if (cseen[li]) return false;
    cseen[li] = 1;
    H.cell2lat[c] = (uint32_t)li;
    H.cell_xyz[c * 3] = cx; H.cell_xyz[c * 3 + 1] = cy; H.cell_xyz[c * 3 + 2] = cz;
  }
  return true;
}

// cell coordinates of a box level (used by velocity_lattice.cu): dims = pressure nodes per direction
bool lattice_cell_coords(const std::vector<uint32_t> &nodes_p, int64_t n_cells, int64_t n_p, int dims[3], std::vector<int> &cell_xyz)
{
  LatHost H;
  if (nodes_p.size() != (size_t)n_cells * 8 || !detect_lattice(nodes_p, n_cells, n_p, H)) return false;
  dims[0] = H.nx; dims[1] = H.ny; dims[2] = H.nz;
  cell_xyz.swap(H.cell_xyz);
  return true;
}

template <class T> static int up(smf_ctx *ctx, T **dev, const std::vector<T> &h)
{
  SMF_CUDA(cudaMalloc(dev, sizeof(T) * std::max<size_t>(h.size(), 1)));
  if (!h.empty()) SMF_CUDA(cudaMemcpy(*dev, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  ctx->bytes += sizeof(T) * h.size();
  return SMF_OK;
}

void lattice_free(smf_ctx *ctx)
{
  replica_free(ctx);             // (built on the lattice's knowledge of the rank's box)
  ctx->replica_tried = false;
  for (auto &L : ctx->lat) {
    cudaFree(L.dof2lat); cudaFree(L.cell2lat); cudaFree(L.coef); cudaFree(L.dinv);
    cudaFree(L.x); cudaFree(L.xo); cudaFree(L.b); cudaFree(L.t); cudaFree(L.mask); cudaFree(L.if_nodes);
    cudaFree(L.ex.send_idx); cudaFree(L.ex.if_idx);
  }
  ctx->lat.clear();
  for (auto &L : ctx->glat) { cudaFree(L.coef); cudaFree(L.dinv); cudaFree(L.x); cudaFree(L.xo); cudaFree(L.b); cudaFree(L.t); }
  ctx->glat.clear();
  ctx->lat_repl = -1;
  cudaFree(ctx->lat_act); ctx->lat_act = nullptr;
  for (void *g : ctx->graph_L_seg) cudaGraphExecDestroy((cudaGraphExec_t)g);
  ctx->graph_L_seg.clear();
  ctx->graph_L_stale = true;
  ctx->lat_state = 0;
}

// local part of the detection: is every level of this rank a box lattice, consistent from level to level?
static bool lattice_detect_local(smf_ctx *ctx, std::vector<LatHost> &H)
{
  const int nl = ctx->cfg.n_levels;
  if (ctx->cfg.dim != 3 || ctx->adaptive) return false;
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  if (A.hang_p.n || A.n_p != ctx->levels[nl - 1].n_p) return false;
  for (int l = 0; l < nl; ++l) {
    const LevelDev &L = ctx->levels[l];
    if (L.geo || L.n_constrained_p || L.edge_p.n || L.copy_p.n || L.h_nodes_p.size() != (size_t)L.n_cells * 8) return false;
    if (ctx->world > 1 && !L.ex_p.set) return false;
    if (!detect_lattice(L.h_nodes_p, L.n_cells, L.n_p, H[l])) return false;
    if (l > 0) {
      // the children of cell (cx, cy, cz) are the cells (2 cx + i, 2 cy + j, 2 cz + k) in deal.II's child order i + 2 j + 4 k
      const std::vector<int32_t> &ch = ctx->transfers[l].h_children;
      const LatHost &C = H[l - 1], &F = H[l];
      if (ch.size() != (size_t)ctx->levels[l - 1].n_cells * 8) return false;
      if (F.nx != 2 * C.nx - 1 || F.ny != 2 * C.ny - 1 || F.nz != 2 * C.nz - 1) return false;
      for (int64_t p = 0; p < ctx->levels[l - 1].n_cells; ++p)
        for (int k = 0; k < 8; ++k) {
          const int32_t fc = ch[p * 8 + k];
          if (fc < 0 || fc >= L.n_cells) return false;
          if (F.cell_xyz[(size_t)fc * 3] != 2 * C.cell_xyz[p * 3] + (k & 1) || F.cell_xyz[(size_t)fc * 3 + 1] != 2 * C.cell_xyz[p * 3 + 1] + ((k >> 1) & 1) ||
              F.cell_xyz[(size_t)fc * 3 + 2] != 2 * C.cell_xyz[p * 3 + 2] + (k >> 2))
            return false;
        }
    }
  }
  return true;
}

static int lattice_detect(smf_ctx *ctx)
{
  const int nl = ctx->cfg.n_levels;
  ctx->lat_state = -1;
  std::vector<LatHost> H(nl);
  bool ok = lattice_detect_local(ctx, H);
  if (ctx->world > 1) {      // the path changes the sequence of halo exchanges: all ranks or none
    double v = ok ? 1.0 : 0.0, *d = nullptr;
    SMF_CUDA(cudaMalloc(&d, sizeof(double)));
    SMF_CUDA(cudaMemcpyAsync(d, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    TRY(allreduce_sum(ctx, d, 1));
    SMF_CUDA(cudaMemcpyAsync(&v, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SMF_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    ok = v == (double)ctx->world;
  }
  if (!ok) return SMF_OK;
  ctx->lat.assign(nl, LatLevel());
  for (int l = 0; l < nl; ++l) {
    LevelDev &L = ctx->levels[l];
    LatLevel &T = ctx->lat[l];
    T.nx = H[l].nx; T.ny = H[l].ny; T.nz = H[l].nz; T.n = L.n_p;
    TRY(up(ctx, &T.dof2lat, H[l].dof2lat));
    TRY(up(ctx, &T.cell2lat, H[l].cell2lat));
    SMF_CUDA(cudaMalloc(&T.coef, sizeof(double) * L.n_cells));
    SMF_CUDA(cudaMalloc(&T.dinv, sizeof(double) * T.n));
    SMF_CUDA(cudaMalloc(&T.x, sizeof(double) * T.n)); SMF_CUDA(cudaMalloc(&T.xo, sizeof(double) * T.n));
    SMF_CUDA(cudaMalloc(&T.b, sizeof(double) * T.n)); SMF_CUDA(cudaMalloc(&T.t, sizeof(double) * T.n));
    ctx->bytes += sizeof(double) * (L.n_cells + 5 * T.n);
    if (ctx->world > 1) {
      // interface nodes of this rank's box: mask + list, and the level's exchange lists in lattice indices
      const ExchangeDev &E = L.ex_p;
      std::vector<unsigned char> mask((size_t)T.n, 0);
      std::vector<uint32_t> send((size_t)E.n_send), ifx((size_t)E.n_if), ifn((size_t)E.n_if);
      if (E.n_send) SMF_CUDA(cudaMemcpy(send.data(), E.send_idx, sizeof(uint32_t) * E.n_send, cudaMemcpyDeviceToHost));
      if (E.n_if) SMF_CUDA(cudaMemcpy(ifx.data(), E.if_idx, sizeof(uint32_t) * E.n_if, cudaMemcpyDeviceToHost));
      for (auto &e : send) e = H[l].dof2lat[e & NODE_IDX_MASK] | (e & ~NODE_IDX_MASK);
      for (size_t i = 0; i < ifx.size(); ++i) {
        const uint32_t li = H[l].dof2lat[ifx[i] & NODE_IDX_MASK];
        ifn[i] = li;
        mask[li] |= 1;
        ifx[i] = li | (ifx[i] & ~NODE_IDX_MASK);
      }
      for (int64_t v = 0; v < L.n_p; ++v)
        if (!L.h_not_owned_p.empty() && L.h_not_owned_p[v]) mask[H[l].dof2lat[v]] |= 2;
      TRY(up(ctx, &T.mask, mask));
      TRY(up(ctx, &T.if_nodes, ifn));
      T.n_if = E.n_if;
      T.ex = E;                                  // shallow twin: buffers, tables and the peer-memory state are shared ...
      T.ex.send_idx = nullptr; T.ex.if_idx = nullptr;
      TRY(up(ctx, &T.ex.send_idx, send));        // ... the two index lists are its own
      TRY(up(ctx, &T.ex.if_idx, ifx));
    }
  }
  if (ctx->world > 1) {
    // Where does this rank's box sit in the global lattice of every level?  The partition-independent DoF numbers of
    // smf_set_exchange tell, if they are lexicographic in that lattice (checked, not assumed): gid = g0 + X + gnx (Y + gny Z).
    // The decision to replicate coarse levels is collective: one all-reduce carries every rank's answer.
    std::vector<double> mine((size_t)4 * nl * ctx->world, 0.0);
    for (int l = 0; l < nl; ++l) {
      LatLevel &T = ctx->lat[l];
      const std::vector<int64_t> &gid = ctx->levels[l].ex_p.h_gid;
      bool ok = gid.size() == (size_t)T.n && T.nx >= 2 && T.ny >= 2 && T.nz >= 2;
      int64_t g0 = 0, sy = 0, sz = 0;
      if (ok) {
        std::vector<int64_t> lg((size_t)T.n);
        for (int64_t v = 0; v < T.n; ++v) lg[H[l].dof2lat[v]] = gid[v];
        g0 = lg[0]; sy = lg[T.nx] - g0; sz = lg[(size_t)T.nx * T.ny] - g0;
        ok = sy >= T.nx && sz >= sy * T.ny && sz % sy == 0;
        for (int64_t z = 0; ok && z < T.nz; ++z)
          for (int64_t y = 0; ok && y < T.ny; ++y)
            for (int64_t x = 0; x < T.nx; ++x)
              if (lg[(z * T.ny + y) * T.nx + x] != g0 + x + sy * y + sz * z) { ok = false; break; }
      }
      if (ok) {
        T.gnx = (int)sy; T.gny = (int)(sz / sy);
        T.oz = (int)(g0 / sz); T.oy = (int)((g0 % sz) / sy); T.ox = (int)(g0 % sy);
        ok = T.ox + T.nx <= T.gnx && T.oy + T.ny <= T.gny;
      }
      double *m = mine.data() + ((size_t)ctx->rank * nl + l) * 4;
      m[0] = ok ? 1.0 : 0.0; m[1] = T.gnx; m[2] = T.gny; m[3] = T.oz + T.nz;
    }
    double *d = nullptr;
    SMF_CUDA(cudaMalloc(&d, sizeof(double) * mine.size()));
    SMF_CUDA(cudaMemcpyAsync(d, mine.data(), sizeof(double) * mine.size(), cudaMemcpyHostToDevice, ctx->stream));
    TRY(allreduce_sum(ctx, d, (int)mine.size()));
    SMF_CUDA(cudaMemcpyAsync(mine.data(), d, sizeof(double) * mine.size(), cudaMemcpyDeviceToHost, ctx->stream));
    SMF_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    const char *env = getenv("SMF_LAT_REPLICATE");
    const int64_t max_nodes = env ? atoll(env) : 2300000;      // largest global level every rank computes redundantly
    for (int l = 0; l < nl - 1; ++l) {
      bool ok = true;
      int gnz = 0;
      for (int r = 0; r < ctx->world; ++r) {
        const double *m = mine.data() + ((size_t)r * nl + l) * 4;
        ok = ok && m[0] == 1.0 && (int)m[1] == ctx->lat[l].gnx && (int)m[2] == ctx->lat[l].gny;
        gnz = std::max(gnz, (int)m[3]);
      }
      ctx->lat[l].gnz = gnz;
      if (!ok || (int64_t)ctx->lat[l].gnx * ctx->lat[l].gny * gnz > max_nodes) break;
      ctx->lat_repl = l;
    }
    ctx->glat.assign(ctx->lat_repl + 1, LatLevel());
    for (int l = 0; l <= ctx->lat_repl; ++l) {
      LatLevel &G = ctx->glat[l];
      const LatLevel &T = ctx->lat[l];
      G.nx = T.gnx; G.ny = T.gny; G.nz = T.gnz; G.n = (int64_t)G.nx * G.ny * G.nz;
      const int64_t nc = (int64_t)(G.nx - 1) * (G.ny - 1) * (G.nz - 1);
      SMF_CUDA(cudaMalloc(&G.coef, sizeof(double) * nc)); SMF_CUDA(cudaMalloc(&G.dinv, sizeof(double) * G.n));
      SMF_CUDA(cudaMalloc(&G.x, sizeof(double) * G.n)); SMF_CUDA(cudaMalloc(&G.xo, sizeof(double) * G.n));
      SMF_CUDA(cudaMalloc(&G.b, sizeof(double) * G.n)); SMF_CUDA(cudaMalloc(&G.t, sizeof(double) * G.n));
      ctx->bytes += sizeof(double) * (nc + 5 * G.n);
    }
  }
  {
    LevelDev &A = ctx->lev(SMF_ACTIVE);
    // active pressure DoF -> lattice index on the finest level (copy_to_mg / copy_from_mg composed with the permutation)
    std::vector<uint32_t> act((size_t)A.n_p);
    std::vector<uint32_t> map;
    if (ctx->copy_p) { map.resize((size_t)A.n_p); SMF_CUDA(cudaMemcpy(map.data(), ctx->copy_p, sizeof(uint32_t) * A.n_p, cudaMemcpyDeviceToHost)); }
    for (int64_t i = 0; i < A.n_p; ++i) act[i] = H[nl - 1].dof2lat[map.empty() ? (uint32_t)i : map[i]];
    TRY(up(ctx, &ctx->lat_act, act));
  }
  ctx->lat_state = 1;
  return SMF_OK;
}

// called by smf_setup_smoothers: detection once per mesh, coefficient tables and inverse diagonals every time
int lattice_prepare(smf_ctx *ctx)
{
  if (!ctx->use_lattice) return SMF_OK;
  if (ctx->lat_state == 0) {
    const int e = lattice_detect(ctx);
    if (e) { lattice_free(ctx); ctx->lat_state = -1; return e; }
    if (ctx->lat_state != 1) { const int s = ctx->lat_state; lattice_free(ctx); ctx->lat_state = s; }
  }
  if (ctx->lat_state != 1) return SMF_OK;
  for (int l = 0; l < ctx->cfg.n_levels; ++l) {
    const LevelDev &L = ctx->levels[l];
    LatLevel &T = ctx->lat[l];
    if (L.n_eta != 1) { lattice_free(ctx); ctx->lat_state = -1; return SMF_OK; }      // (one viscosity per quadrature point: cell-batch path)
    const double scale = L.ps * L.ps * L.h[0] * L.h[1] * L.h[2] / 216.0;
    k_lat_coef<<<(unsigned)((L.n_cells + 255) / 256), 256, 0, ctx->stream>>>(T.coef, L.eta, T.cell2lat, L.n_cells, scale);
    k_lat_scatter<<<(unsigned)((T.n + 255) / 256), 256, 0, ctx->stream>>>(T.dinv, L.dinv_M, T.dof2lat, T.n);
    ctx->launches += 2;
  }
  for (int l = 0; l <= ctx->lat_repl; ++l) {
    LatLevel &G = ctx->glat[l];
    const LatLevel &T = ctx->lat[l];
    const int64_t nc = (int64_t)(G.nx - 1) * (G.ny - 1) * (G.nz - 1), lc = (int64_t)(T.nx - 1) * (T.ny - 1) * (T.nz - 1);
    SMF_CUDA(cudaMemsetAsync(G.coef, 0, sizeof(double) * nc, ctx->stream));
    k_lat_coef_global<<<(unsigned)((lc + 255) / 256), 256, 0, ctx->stream>>>(G.coef, T.coef, LatDims{T.nx - 1, T.ny - 1, T.nz - 1},
                                                                        LatBox{T.nx - 1, T.ny - 1, T.nz - 1, G.nx - 1, G.ny - 1, T.ox, T.oy, T.oz});
    TRY(allreduce_sum(ctx, G.coef, (int)nc));       // every cell belongs to one rank: the sum is a gather, bitwise equal everywhere
    k_lat_dinv<<<(unsigned)((G.n + 255) / 256), 256, 0, ctx->stream>>>(G.dinv, G.coef, LatDims{G.nx, G.ny, G.nz});
    ctx->launches += 2;
  }
  SMF_CUDA(cudaGetLastError());
  ctx->graph_L_stale = true;
  // coarse tail: the levels whose vectors fit one block's shared memory (no exchanges inside a block: one rank, or replicated levels)
  ctx->lat_tail = -1;
  if (ctx->use_coarse_tail) {
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    size_t doubles = 0;
    for (int l = 0; l < ctx->cfg.n_levels - 1 && l < LT_MAX_LEVELS; ++l) {
      if (ctx->world > 1 && l > ctx->lat_repl) break;
      const LatLevel &T = ctx->world > 1 ? ctx->glat[l] : ctx->lat[l];
      if (T.n > ctx->tail_max_cells * 2) break;
      doubles += (size_t)4 * T.n + (size_t)(T.nx - 1) * (T.ny - 1) * (T.nz - 1);
      if (doubles * 8 + 1024 > (size_t)max_smem) break;
      ctx->lat_tail = l;
    }
  }
  return SMF_OK;
}

// the lattice the cycle works on at level l: the rank's own box, or the replicated global lattice
static LatLevel &lat_level(smf_ctx *ctx, int l) { return l <= ctx->lat_repl ? ctx->glat[l] : ctx->lat[l]; }

static int lat_launch_tail(smf_ctx *ctx, int top)
{
  LatTail P{};
  P.n_levels = top + 1; P.steps = 4;
  int off = 0;
  for (int l = 0; l <= top; ++l) {
    const LatLevel &T = lat_level(ctx, l);
    const Smoother &sm = ctx->smoother_S[l];
    P.lv[l] = LatTailLevel{T.nx, T.ny, T.nz, off, sm.theta, sm.delta, sm.degree, T.coef};
    off += lt_doubles(P.lv[l]);
  }
  P.b_in = lat_level(ctx, top).b; P.x_out = lat_level(ctx, top).x;
  const size_t bytes = (size_t)off * sizeof(double);
  static size_t configured_dev[MAX_DEVICES] = {0};
  size_t &configured = configured_dev[ctx->cfg.device % MAX_DEVICES];
  if (bytes > configured) { SMF_CUDA(cudaFuncSetAttribute(k_lat_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); configured = bytes; }
  ProfScope prof(ctx, B_MASS);
  k_lat_tail<<<1, LT_THREADS, bytes, ctx->stream>>>(P);
  ctx->launches++;
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}

// ---- the V-cycle -------------------------------------------------------------------------------------------------
static int lat_zpt(const LatLevel &T)
{
  static const int forced = getenv("SMF_LAT_ZPT") ? atoi(getenv("SMF_LAT_ZPT")) : 0;
  if (forced > 0) return forced;
  // nodes per thread along z: long marches on the large levels (fewer plane loads), short ones where threads are scarce
  const int64_t cols = (int64_t)T.nx * T.ny;
  int zpt = 1;
  while (zpt < 8 && cols * ((T.nz + 2 * zpt - 1) / (2 * zpt)) >= 148 * 1024) zpt *= 2;
  return zpt;
}
template <int MODE>
static int lat_apply(smf_ctx *ctx, int l, const double *x, double *xo, const double *b, double *t, double f1, double f2, bool use_xo)
{
  LatLevel &T = lat_level(ctx, l);
  LatStep P{};
  P.d = LatDims{T.nx, T.ny, T.nz};
  P.x = x; P.coef = T.coef; P.dinv = T.dinv; P.b = b; P.xo = xo; P.t = t; P.mask = T.mask;
  P.f1 = f1; P.f2 = f2; P.use_xo = use_xo ? 1 : 0; P.zpt = lat_zpt(T);
  const dim3 grid((unsigned)(((int64_t)T.nx * T.ny + LAT_THREADS - 1) / LAT_THREADS), (unsigned)((T.nz + P.zpt - 1) / P.zpt));
  {
    ProfScope prof(ctx, B_MASS);
    k_lat_step<MODE><<<grid, LAT_THREADS, 0, ctx->stream>>>(P);
    ctx->launches++;
  }
  if (T.mask && ctx->world > 1) {
    TRY(exchange_add(ctx, T.ex, t));
    if (MODE != 0 && T.n_if) {
      ProfScope prof(ctx, B_SMOOTH_VEC);
      k_lat_finish<MODE><<<(unsigned)((T.n_if + 127) / 128), 128, 0, ctx->stream>>>(T.if_nodes, T.n_if, x, xo, T.dinv, b, t, f1, f2, use_xo ? 1 : 0);
      ctx->launches++;
    }
  }
  return SMF_OK;
}

// Chebyshev iterations after the first update (solver.cu: cheb_tail); x = current iterate, xo = previous
static int lat_cheb_tail(smf_ctx *ctx, int l, double *&x, double *&xo, const double *b, bool have_old)
{
  const Smoother &sm = ctx->smoother_S[l];
  if (sm.degree < 2 || std::fabs(sm.delta) < 1e-40) return SMF_OK;
  double rhok = sm.delta / sm.theta;
  const double sigma = sm.theta / sm.delta;
  for (int k = 0; k < sm.degree - 1; ++k) {
    const double rhokp = 1.0 / (2.0 * sigma - rhok);
    const double f1 = rhokp * rhok, f2 = 2.0 * rhokp / sm.delta;
    rhok = rhokp;
    TRY(lat_apply<1>(ctx, l, x, xo, b, lat_level(ctx, l).t, f1, f2, have_old));
    std::swap(x, xo);
    have_old = true;
  }
  return SMF_OK;
}
static int lat_cheb_vmult(smf_ctx *ctx, int l, double *&x, double *&xo, const double *b)
{
  VecBucket vb(ctx, B_SMOOTH_VEC);
  TRY(cheb_first(ctx, x, lat_level(ctx, l).dinv, b, 1.0 / ctx->smoother_S[l].theta, lat_level(ctx, l).n));
  return lat_cheb_tail(ctx, l, x, xo, b, false);
}
static int lat_cheb_step(smf_ctx *ctx, int l, double *&x, double *&xo, const double *b)
{
  TRY(lat_apply<1>(ctx, l, x, xo, b, lat_level(ctx, l).t, 0.0, 1.0 / ctx->smoother_S[l].theta, false));
  std::swap(x, xo);
  return lat_cheb_tail(ctx, l, x, xo, b, true);
}
// Multigrid::level_v_step (solver.cu: level_v_step) with 4 smoothing steps on either side; the result ends in res
// The cycle is captured into CUDA graphs; the one NCCL all-reduce of a cycle on several ranks (defect of the finest replicated
// level) stays outside: the capture is cut there into two segments and the replay runs segment, all-reduce, segment.
static int lat_cut_capture(smf_ctx *ctx, std::vector<cudaGraph_t> *segs)
{
  cudaGraph_t g = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
  if (ce != cudaSuccess) return ctx->fail(SMF_ERR_CUDA, std::string("cudaStreamEndCapture (lattice V-cycle): ") + cudaGetErrorString(ce));
  segs->push_back(g);
  return SMF_OK;
}

static int lat_level_step(smf_ctx *ctx, int l, double *&res)
{
  LatLevel &T = lat_level(ctx, l);
  if (l == ctx->lat_tail) { res = T.x; return lat_launch_tail(ctx, l); }
  const int steps = 4;
  double *x = T.x, *xo = T.xo;
  TRY(lat_cheb_vmult(ctx, l, x, xo, T.b));
  for (int s = 1; s < steps; ++s) TRY(lat_cheb_step(ctx, l, x, xo, T.b));
  if (l > 0) {
    LatLevel &C = lat_level(ctx, l - 1);
    const bool into_global = l - 1 <= ctx->lat_repl && l > ctx->lat_repl;      // distributed level -> replicated level
    const LatLevel &CL = ctx->lat[l - 1];                                       // (the rank's own coarse box: extent of the region under T)
    const LatBox cbox = into_global ? LatBox{CL.nx, CL.ny, CL.nz, C.nx, C.ny, CL.ox, CL.oy, CL.oz} : LatBox{C.nx, C.ny, C.nz, C.nx, C.ny, 0, 0, 0};
    TRY(lat_apply<2>(ctx, l, x, nullptr, T.b, T.t, 0.0, 0.0, false));      // t = defect - M x
    {
      ProfScope prof(ctx, B_TRANSFER);
      if (into_global) SMF_CUDA(cudaMemsetAsync(C.b, 0, sizeof(double) * C.n, ctx->stream));
      const int64_t nc = (int64_t)cbox.nx * cbox.ny * cbox.nz;
      k_lat_restrict<<<(unsigned)((nc + 127) / 128), 128, 0, ctx->stream>>>(C.b, T.t, T.mask, cbox, LatDims{T.nx, T.ny, T.nz});
      ctx->launches++;
    }
    if (into_global) {
      // sum of the ranks' partial defects (interface nodes) = gather of the rest: every rank now holds the global coarse defect
      if (ctx->lat_capturing) {
        TRY(lat_cut_capture(ctx, (std::vector<cudaGraph_t> *)ctx->lat_capture_segs));
        SMF_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
      } else TRY(allreduce_sum(ctx, C.b, (int)C.n));
    } else if (C.mask && ctx->world > 1) TRY(exchange_add(ctx, C.ex, C.b));
    double *cres = nullptr;
    TRY(lat_level_step(ctx, l - 1, cres));
    {
      ProfScope prof(ctx, B_TRANSFER);
      k_lat_prolong_add<<<(unsigned)((T.n + 255) / 256), 256, 0, ctx->stream>>>(x, cres, cbox, LatDims{T.nx, T.ny, T.nz});
      ctx->launches++;
    }
    for (int s = 0; s < steps; ++s) TRY(lat_cheb_step(ctx, l, x, xo, T.b));
  }
  res = x;
  return SMF_OK;
}

int lattice_vcycle(smf_ctx *ctx, double *dst, const double *src)
{
  const int top = ctx->cfg.n_levels - 1;
  LatLevel &T = ctx->lat[top];
  LevelDev &A = ctx->lev(SMF_ACTIVE);
  {
    ProfScope prof(ctx, B_VEC);
    k_lat_scatter<<<(unsigned)((A.n_p + 255) / 256), 256, 0, ctx->stream>>>(T.b, src, ctx->lat_act, A.n_p);
    ctx->launches++;
  }
  double *res = nullptr;
  const bool graph = ctx->use_graphs && ctx->prof_mode == 0;
  if (!graph) TRY(lat_level_step(ctx, top, res));
  else {
    if (ctx->graph_L_stale) {
      std::vector<cudaGraph_t> segs;
      ctx->lat_capture_segs = &segs;
      ctx->lat_capturing = true;
      SMF_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
      const int64_t l0 = ctx->launches, x0 = ctx->exchanges;
      int rc = lat_level_step(ctx, top, res);
      const int rc2 = lat_cut_capture(ctx, &segs);
      ctx->lat_capturing = false;
      ctx->lat_capture_segs = nullptr;
      if (!rc) rc = rc2;
      if (rc) { for (cudaGraph_t g : segs) if (g) cudaGraphDestroy(g); return rc; }
      ctx->graph_L_launches = ctx->launches - l0; ctx->launches = l0;
      ctx->graph_L_exchanges = ctx->exchanges - x0; ctx->exchanges = x0;
      if (ctx->graph_L_seg.size() != segs.size()) {
        for (void *g : ctx->graph_L_seg) cudaGraphExecDestroy((cudaGraphExec_t)g);
        ctx->graph_L_seg.assign(segs.size(), nullptr);
      }
      for (size_t k = 0; k < segs.size(); ++k) {
        bool updated = false;
        if (ctx->graph_L_seg[k]) {
          cudaGraphExecUpdateResultInfo info;
          updated = cudaGraphExecUpdate((cudaGraphExec_t)ctx->graph_L_seg[k], segs[k], &info) == cudaSuccess;
          if (!updated) { cudaGetLastError(); cudaGraphExecDestroy((cudaGraphExec_t)ctx->graph_L_seg[k]); ctx->graph_L_seg[k] = nullptr; }
        }
        if (!updated) {
          cudaGraphExec_t ge = nullptr;
          const cudaError_t ie = cudaGraphInstantiate(&ge, segs[k], 0);
          if (ie != cudaSuccess) { for (cudaGraph_t g : segs) cudaGraphDestroy(g); return ctx->fail(SMF_ERR_CUDA, std::string("cudaGraphInstantiate (lattice V-cycle): ") + cudaGetErrorString(ie)); }
          ctx->graph_L_seg[k] = ge;
        }
        cudaGraphDestroy(segs[k]);
      }
      ctx->graph_L_stale = false;
      ctx->lat_res = res;
    }
    for (size_t k = 0; k < ctx->graph_L_seg.size(); ++k) {
      SMF_CUDA(cudaGraphLaunch((cudaGraphExec_t)ctx->graph_L_seg[k], ctx->stream));
      if (k + 1 < ctx->graph_L_seg.size()) {      // the cut: all-reduce of the defect of the finest replicated level
        LatLevel &C = ctx->glat[ctx->lat_repl];
        TRY(allreduce_sum(ctx, C.b, (int)C.n));
      }
    }
    ctx->launches += ctx->graph_L_launches;
    ctx->exchanges += ctx->graph_L_exchanges;
    res = ctx->lat_res;
  }
  {
    ProfScope prof(ctx, B_VEC);
    k_lat_gather<<<(unsigned)((A.n_p + 255) / 256), 256, 0, ctx->stream>>>(dst, res, ctx->lat_act, A.n_p);
    ctx->launches++;
  }
  SMF_CUDA(cudaGetLastError());
  return SMF_OK;
}

}  // namespace smf
