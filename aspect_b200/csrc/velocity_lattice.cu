// Lattice-ordered mirror of a structured box level (SMF_OPT_LATTICE_VELOCITY, experimental): host-side set-up for
// velocity_lat_kernel (operators.cu), the A-block kernel that stages its bricks with TMA box loads.
//
// Why: velocity_sb_kernel is bound by the shared-memory pipe (ncu: LSU wavefronts 72 %), and more than a third of its
// wavefronts is not arithmetic traffic but the list-driven staging of a brick's 9 x 9 x 5 node lattice (index + slot loads,
// 8-byte cp.async into a parity-class layout, bank conflicts of the list order) and its flush.  If the level vectors are
// stored in lattice order, [component][x parity][Z][Y][X / 2], a brick is six BOXES of that tensor: the TMA engine copies it without a single
// load/store instruction of the SM, and the node lists (0.36 GB of DRAM traffic per application at refinement 7) disappear.
// The gathers stay conflict-free because even and odd lattice columns live in separate arrays (in global memory too: TMA
// has no element stride in the innermost dimension), rows padded to 6, planes to 60 doubles.
//
// The numbering of the vectors is an input of the C ABI (MatrixFree's), so the mirror is internal: this
// file derives the lattice from the cell -> DoF maps (no numbering assumed), checks that the level's constraints are whole
// boundary faces per component, and builds the permutation.  The experiment applies the kernel between two permutation
// passes (smf_vmult_A on a structured level); the kernel itself is timed separately (smf_profile_*).
#include "smf_internal.h"
#include <algorithm>

using namespace smf;

namespace smf {

void vlat_free(LevelDev &L)
{
  if (!L.vlat) return;
  cudaFree(L.vlat->lat2dof); cudaFree(L.vlat->brick_org); cudaFree(L.vlat->src); cudaFree(L.vlat->dst);
  delete L.vlat;
  L.vlat = nullptr;
}

int vlat_prepare(smf_ctx *ctx, LevelDev &L)
{
  if (L.vlat || !L.structured || L.dim != 3 || L.geo || L.hang_u.n || L.edge_u.n || L.n_eta != 1) return SMF_OK;
  int pd[3];
  std::vector<int> cxyz;
  if (!lattice_cell_coords(L.h_nodes_p, L.n_cells, L.n_p, pd, cxyz)) return SMF_OK;
  const int nx = 2 * pd[0] - 1, ny = 2 * pd[1] - 1, nz = 2 * pd[2] - 1;
  const int64_t n_nodes = (int64_t)nx * ny * nz;
  if (n_nodes * 3 != L.n_u) return SMF_OK;
  if ((int64_t)6 * nz * ny * ((nx + 1) / 2 + 2) >= (int64_t)1 << 31) return SMF_OK;      // (the flush of the kernel runs on 32-bit offsets)
  // node (X, Y, Z) -> first DoF; every cell must agree on the nodes it shares
  const uint32_t NONE = 0xffffffffu;
  std::vector<uint32_t> lat2dof((size_t)n_nodes, NONE);
  std::vector<uint32_t> flags((size_t)n_nodes, 0);
  for (int64_t c = 0; c < L.n_cells; ++c)
    for (int n = 0; n < 27; ++n) {
      const uint32_t e = L.h_nodes_u[(size_t)c * 27 + n];
      const int X = 2 * cxyz[c * 3] + n % 3, Y = 2 * cxyz[c * 3 + 1] + (n / 3) % 3, Z = 2 * cxyz[c * 3 + 2] + n / 9;
      const int64_t li = ((int64_t)Z * ny + Y) * nx + X;
      if (lat2dof[li] == NONE) { lat2dof[li] = e & NODE_IDX_MASK; flags[li] = e >> NODE_FLAG_SHIFT; }
      else if (lat2dof[li] != (e & NODE_IDX_MASK)) return SMF_OK;
    }
  for (int64_t i = 0; i < n_nodes; ++i) if (lat2dof[i] == NONE) return SMF_OK;
  // constraints: whole boundary faces per component?
  int mask = 0;
  {
    // a face/component pair is constrained if the face's centre node is
    const int mid[3] = {nx / 2, ny / 2, nz / 2}, hi[3] = {nx - 1, ny - 1, nz - 1};
    for (int d = 0; d < 3; ++d)
      for (int s = 0; s < 2; ++s) {
        int p[3] = {mid[0], mid[1], mid[2]};
        p[d] = s ? hi[d] : 0;
        const uint32_t f = flags[((int64_t)p[2] * ny + p[1]) * nx + p[0]];
        for (int c = 0; c < 3; ++c) if ((f >> c) & 1u) mask |= 1 << (3 * (2 * d + s) + c);
      }
    for (int Z = 0; Z < nz; ++Z)
      for (int Y = 0; Y < ny; ++Y)
        for (int X = 0; X < nx; ++X) {
          uint32_t expect = 0;
          const int p[3] = {X, Y, Z};
          for (int d = 0; d < 3; ++d) {
            if (p[d] == 0) expect |= (mask >> (3 * (2 * d))) & 7u;
            if (p[d] == hi[d]) expect |= (mask >> (3 * (2 * d + 1))) & 7u;
          }
          if (flags[((int64_t)Z * ny + Y) * nx + X] != expect) return SMF_OK;      // (constraints that are not whole faces: list kernel)
        }
  }
  // bricks: 32 consecutive cells = a 4 x 4 x 2 box in the lane order of the brick kernels
  std::vector<int> org((size_t)L.n_batches * 4, 0);
  if ((int64_t)L.n_batches * 32 != L.n_cells) return SMF_OK;
  for (int b = 0; b < L.n_batches; ++b) {
    const int *c0 = &cxyz[(size_t)b * 32 * 3];
    for (int l = 0; l < 32; ++l) {
      const int *c = &cxyz[((size_t)b * 32 + l) * 3];
      if (c[0] != c0[0] + sb_x(l) || c[1] != c0[1] + sb_y(l) || c[2] != c0[2] + sb_z(l)) return SMF_OK;
    }
    int f = 0;
    const int lo[3] = {2 * c0[0], 2 * c0[1], 2 * c0[2]}, ext[3] = {8, 8, 4}, hi[3] = {nx - 1, ny - 1, nz - 1};
    for (int d = 0; d < 3; ++d) {
      if (lo[d] == 0) f |= 1 << d;
      if (lo[d] + ext[d] == hi[d]) f |= 1 << (3 + d);
    }
    org[b * 4] = lo[0]; org[b * 4 + 1] = lo[1]; org[b * 4 + 2] = lo[2]; org[b * 4 + 3] = f;
  }
  VLat *V = new VLat();
  V->nx = nx; V->ny = ny; V->nz = nz; V->Xp = ((nx + 1) / 2 + 1) & ~1;      // row pitch of one parity array (even: rows start 16-byte aligned)
  V->n_nodes = n_nodes; V->n_lat = (int64_t)6 * nz * ny * V->Xp;
  V->cons_mask = mask;
  L.vlat = V;
  SMF_CUDA(cudaMalloc(&V->lat2dof, sizeof(uint32_t) * n_nodes));
  SMF_CUDA(cudaMemcpy(V->lat2dof, lat2dof.data(), sizeof(uint32_t) * n_nodes, cudaMemcpyHostToDevice));
  SMF_CUDA(cudaMalloc(&V->brick_org, sizeof(int) * org.size()));
  SMF_CUDA(cudaMemcpy(V->brick_org, org.data(), sizeof(int) * org.size(), cudaMemcpyHostToDevice));
  SMF_CUDA(cudaMalloc(&V->src, sizeof(double) * V->n_lat));
  SMF_CUDA(cudaMalloc(&V->dst, sizeof(double) * V->n_lat));
  SMF_CUDA(cudaMemset(V->src, 0, sizeof(double) * V->n_lat));      // (the pad column stays zero)
  ctx->bytes += sizeof(uint32_t) * n_nodes + sizeof(int) * org.size() + 2 * sizeof(double) * V->n_lat;
  return SMF_OK;
}

}  // namespace smf
