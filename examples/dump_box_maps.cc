// Writes the index maps of the C++ box driver (aspect_b200/cpp/synthetic_box.h) as raw little-endian arrays, for the test
// that compares them bit by bit with the Python driver:  dump_box_maps <dim> <R> <bc> <out file>
#include <cstdio>
#include <cstdlib>

#include "synthetic_box.h"

template <class T> static void put(std::FILE *f, const std::vector<T> &v)
{
  const std::int64_t n = (std::int64_t)v.size();
  std::fwrite(&n, sizeof n, 1, f);
  if (n) std::fwrite(v.data(), sizeof(T), (std::size_t)n, f);
}

int main(int argc, char **argv)
{
  if (argc < 5) return 1;
  const auto h = aspect_b200::make_box_hierarchy(std::atoi(argv[1]), std::atoi(argv[2]), argv[3]);
  std::FILE *f = std::fopen(argv[4], "wb");
  if (!f) return 1;
  for (std::size_t l = 0; l < h.levels.size(); ++l)
    {
      const auto &L = h.levels[l];
      const std::int64_t sizes[3] = {L.n_cells, L.n_dofs_u, L.n_dofs_p};
      std::fwrite(sizes, sizeof(std::int64_t), 3, f);
      put(f, L.cell_dofs_u); put(f, L.cell_dofs_p); put(f, L.constrained_u); put(f, h.children[l]);
    }
  std::fclose(f);
  return 0;
}
