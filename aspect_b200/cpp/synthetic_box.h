// Synthetic box driver in C++: the arrays ASPECT's handler would pull out of deal.II for a uniformly refined box with one
// coarse cell (source/geometry_model/box.cc:66-74), by the numbering rules of SURVEY App. A.4 -- cells in Morton order
// (deal.II child order), DoFs by first touch over that sequence and, inside a cell, FE-local order (vertices, lines, quads,
// hex interior) with the vector components of a node consecutive; cell->DoF maps lexicographic and component-major as
// MatrixFree stores them.  Same rules as aspect_b200/synthetic.py (the Python driver of the tests), written with plain loops.
#ifndef ASPECT_B200_SYNTHETIC_BOX_H
#define ASPECT_B200_SYNTHETIC_BOX_H

#include <array>
#include <cmath>
#include <string>
#include <vector>

#include "stokes_matrix_free_b200.h"

namespace aspect_b200
{
  namespace internal
  {
    inline std::vector<std::array<int, 3>> lex_nodes(int npd, int dim)
    {
      std::vector<std::array<int, 3>> out;
      const int n = dim == 3 ? npd * npd * npd : npd * npd;
      for (int a = 0; a < n; ++a) out.push_back({a % npd, (a / npd) % npd, dim == 3 ? a / (npd * npd) : 0});
      return out;
    }
    // deal.II FE_Q local ordering as lexicographic multi-indices (degree 1 or 2)
    inline std::vector<std::array<int, 3>> fe_q_local_order(int degree, int dim)
    {
      const int hi = degree, m = 1;
      std::vector<std::array<int, 3>> o;
      for (int v = 0; v < (1 << dim); ++v) o.push_back({hi * (v & 1), hi * ((v >> 1) & 1), dim == 3 ? hi * ((v >> 2) & 1) : 0});
      if (degree == 1) return o;
      if (dim == 2)
        {
          o.push_back({0, m, 0}); o.push_back({hi, m, 0}); o.push_back({m, 0, 0}); o.push_back({m, hi, 0}); o.push_back({m, m, 0});
          return o;
        }
      for (int z : {0, hi}) { o.push_back({0, m, z}); o.push_back({hi, m, z}); o.push_back({m, 0, z}); o.push_back({m, hi, z}); }
      o.push_back({0, 0, m}); o.push_back({hi, 0, m}); o.push_back({0, hi, m}); o.push_back({hi, hi, m});
      o.push_back({0, m, m}); o.push_back({hi, m, m}); o.push_back({m, 0, m}); o.push_back({m, hi, m}); o.push_back({m, m, 0}); o.push_back({m, m, hi});
      o.push_back({m, m, m});
      return o;
    }
    inline std::array<std::int64_t, 3> morton_cell(std::int64_t m, int level, int dim)
    {
      std::array<std::int64_t, 3> ijk{0, 0, 0};
      for (int b = 0; b < level; ++b)
        for (int d = 0; d < dim; ++d) ijk[d] |= ((m >> (dim * b + d)) & 1) << b;
      return ijk;
    }
  }

  /** one level of the box; bc: "zero" (all components on every face) or "tangential" (normal component on each face) */
  inline LevelArrays make_box_level(int dim, int level, const std::string &bc)
  {
    using namespace internal;
    LevelArrays L;
    const std::int64_t n = std::int64_t(1) << level;
    L.n_cells = 1;
    for (int d = 0; d < dim; ++d) L.n_cells *= n;
    const int nn2 = dim == 3 ? 27 : 9, nn1 = dim == 3 ? 8 : 4;
    const std::int64_t g2 = 2 * n + 1, g1 = n + 1;
    const auto ord2 = fe_q_local_order(2, dim), ord1 = fe_q_local_order(1, dim);
    const auto lex2 = lex_nodes(3, dim), lex1 = lex_nodes(2, dim);
    std::vector<std::int64_t> num2((std::size_t)(dim == 3 ? g2 * g2 * g2 : g2 * g2), -1), num1((std::size_t)(dim == 3 ? g1 * g1 * g1 : g1 * g1), -1);
    std::vector<std::array<std::int64_t, 3>> pos2;
    std::int64_t n2 = 0, n1 = 0;
    L.cell_dofs_u.resize((std::size_t)L.n_cells * dim * nn2);
    L.cell_dofs_p.resize((std::size_t)L.n_cells * nn1);
    for (std::int64_t m = 0; m < L.n_cells; ++m)
      {
        const auto ijk = morton_cell(m, level, dim);
        auto key = [&](const std::array<int, 3> &loc, int per, std::int64_t g)
        { return (per * ijk[0] + loc[0]) + g * ((per * ijk[1] + loc[1]) + g * (dim == 3 ? per * ijk[2] + loc[2] : 0)); };
        for (const auto &loc : ord2)
          {
            const std::int64_t k = key(loc, 2, g2);
            if (num2[(std::size_t)k] < 0) { num2[(std::size_t)k] = n2++; pos2.push_back({2 * ijk[0] + loc[0], 2 * ijk[1] + loc[1], dim == 3 ? 2 * ijk[2] + loc[2] : 0}); }
          }
        for (int a = 0; a < nn2; ++a)
          for (int c = 0; c < dim; ++c)
            L.cell_dofs_u[(std::size_t)m * dim * nn2 + c * nn2 + a] = (std::uint32_t)(dim * num2[(std::size_t)key(lex2[a], 2, g2)] + c);
        for (const auto &loc : ord1)
          {
            const std::int64_t k = key(loc, 1, g1);
            if (num1[(std::size_t)k] < 0) num1[(std::size_t)k] = n1++;
          }
        for (int a = 0; a < nn1; ++a) L.cell_dofs_p[(std::size_t)m * nn1 + a] = (std::uint32_t)num1[(std::size_t)key(lex1[a], 1, g1)];
      }
    L.n_dofs_u = dim * n2; L.n_dofs_p = n1;
    for (std::int64_t k = 0; k < n2; ++k)
      for (int c = 0; c < dim; ++c)
        {
          bool any = false;
          for (int d = 0; d < dim; ++d) any = any || pos2[(std::size_t)k][d] == 0 || pos2[(std::size_t)k][d] == 2 * n;
          const bool own = pos2[(std::size_t)k][c] == 0 || pos2[(std::size_t)k][c] == 2 * n;
          if ((bc == "zero" && any) || (bc == "tangential" && own)) L.constrained_u.push_back((std::uint32_t)(dim * k + c));
        }
    L.cell_size.assign((std::size_t)dim, 1.0 / (double)n);
    return L;
  }

  inline HierarchyArrays make_box_hierarchy(int dim, int refinement, const std::string &bc)
  {
    HierarchyArrays h;
    h.dim = dim;
    h.children.resize((std::size_t)refinement + 1);
    for (int l = 0; l <= refinement; ++l)
      {
        h.levels.push_back(make_box_level(dim, l, bc));
        if (l > 0)
          {
            const std::int64_t npar = h.levels[(std::size_t)l - 1].n_cells;
            const int nch = 1 << dim;
            h.children[(std::size_t)l].resize((std::size_t)npar * nch);
            for (std::int64_t p = 0; p < npar; ++p)
              for (int k = 0; k < nch; ++k) h.children[(std::size_t)l][(std::size_t)p * nch + k] = (std::int32_t)(p * nch + k);    // Morton order
          }
      }
    h.active = h.levels.back();
    return h;
  }

  /** physical Gauss points of every cell of a box level: [n_cells][3^dim][dim], q lexicographic (x fastest) */
  inline std::vector<double> box_quadrature_points(int dim, int level)
  {
    const double s = std::sqrt(3.0 / 5.0), gx[3] = {0.5 - 0.5 * s, 0.5, 0.5 + 0.5 * s};
    const std::int64_t n = std::int64_t(1) << level;
    std::int64_t n_cells = 1;
    for (int d = 0; d < dim; ++d) n_cells *= n;
    const int nq = dim == 3 ? 27 : 9;
    std::vector<double> x((std::size_t)n_cells * nq * dim);
    for (std::int64_t m = 0; m < n_cells; ++m)
      {
        const auto ijk = internal::morton_cell(m, level, dim);
        for (int q = 0; q < nq; ++q)
          {
            const int qd[3] = {q % 3, (q / 3) % 3, q / 9};
            for (int d = 0; d < dim; ++d) x[((std::size_t)m * nq + q) * dim + d] = ((double)ijk[d] + gx[qd[d]]) / (double)n;
          }
      }
    return x;
  }
}  // namespace aspect_b200
#endif
