// First-touch numbering of a box lattice in plain loops (host only): the slow part of the synthetic driver
// (aspect_b200/synthetic.py: _first_touch) for meshes of millions of cells.  Same rule as there and as
// aspect_b200/cpp/synthetic_box.h: walk the cells in the given order, inside a cell the FE-local order (vertices, lines, quads,
// hex interior); a lattice node gets the next number when it is touched first.
//   g++ -O2 -shared -fPIC -o aspect_b200/libbox_maps.so aspect_b200/cpp/box_maps.cc
#include <cstdint>
#include <vector>

extern "C" {

// cells [n_cells][dim] lattice positions of the cells; per = npd - 1 (2 for Q2, 1 for Q1); grid[dim] nodes per direction;
// local_order / lex_order [npd^dim][dim] local node multi-indices in FE-local and in lexicographic order.
// Out: cell_nodes [n_cells][npd^dim] (lexicographic), node_grid [n_nodes][dim].  Returns the number of nodes (-1: bad input).
int64_t box_first_touch(int dim, int64_t n_cells, const int64_t *cells, int per, const int64_t *grid, int n_local,
                        const int64_t *local_order, const int64_t *lex_order, int64_t *cell_nodes, int64_t *node_grid, int64_t node_capacity)
{
  if (dim < 2 || dim > 3) return -1;
  int64_t total = 1, mult[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d) { mult[d] = total; total *= grid[d]; }
  std::vector<int32_t> num((size_t)total, -1);
  int64_t n_nodes = 0;
  for (int64_t c = 0; c < n_cells; ++c) {
    int64_t base = 0;
    for (int d = 0; d < dim; ++d) base += per * cells[c * dim + d] * mult[d];
    for (int a = 0; a < n_local; ++a) {
      int64_t key = base;
      for (int d = 0; d < dim; ++d) key += local_order[a * dim + d] * mult[d];
      if (key < 0 || key >= total) return -1;
      if (num[(size_t)key] < 0) {
        if (n_nodes >= node_capacity) return -1;
        num[(size_t)key] = (int32_t)n_nodes;
        int64_t r = key;
        for (int d = 0; d < dim; ++d) { node_grid[n_nodes * dim + d] = r % grid[d]; r /= grid[d]; }
        ++n_nodes;
      }
    }
    for (int a = 0; a < n_local; ++a) {
      int64_t key = base;
      for (int d = 0; d < dim; ++d) key += lex_order[a * dim + d] * mult[d];
      cell_nodes[c * n_local + a] = num[(size_t)key];
    }
  }
  return n_nodes;
}

}  // extern "C"
