// pnp-b200 — mesh partitioner of the multi-GPU path (north_star: "rows are partitioned
// across the 8 GPUs of one box by mesh-graph partition, with halo exchange"). Host code, C ABI
// (include/pnpb200/pnpb200.h: pnpb200_partition_*). Two partitioners, one output format:
//   * pnpb200_partition_box  — dolfin::BoxMesh (src/domain.cpp:306-308) cut into px x py x pz bricks of
//     vertex planes, in closed form: a rank only ever generates its own brick + ghost layer, so the
//     406 M-dof weak-scaling mesh is never materialised anywhere;
//   * pnpb200_partition_mesh — recursive coordinate bisection (RCB) of the vertices of ANY conforming
//     tetrahedral mesh (the adaptively refined meshes of src/mesh_refiner.cpp:73-153); every rank
//     runs the same deterministic bisection on the global arrays and keeps its own part.
// Output for rank r (what pnpb200_create + pnpb200_set_partition consume):
//   local vertices = owned vertices (ascending global id), then ghost vertices grouped by owner rank
//     (ascending rank, ascending global id inside a group);
//   local cells    = every cell that touches an owned vertex (ghost-layer cells are duplicated on the
//     neighbour, so assembly needs no communication beyond the halo of the solution vector);
//   halo lists     = per neighbour k (ascending rank): the owned vertices k holds as ghosts (ascending
//     global id = the order of k's ghost group, so both sides agree without a handshake) and the
//     local index range of the ghosts received from k.
#include "pnpb200/pnpb200.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <numeric>
#include <stdexcept>
#include <vector>

struct pnpb200_partition {
  int n_own = 0, n_loc = 0, n_cells = 0;
  std::vector<double> xyz;            // 3 * n_loc
  std::vector<int> cells;             // 4 * n_cells, local vertex ids, each cell ascending in GLOBAL id (UFC order)
  std::vector<long long> gids;        // n_loc
  std::vector<int> nbr, send_ptr, send_idx, recv_ptr;
  std::vector<int> color;             // n_loc (box: (i + j + k) mod 4 of the GLOBAL lattice; else empty)
  std::vector<int> bc_vertices;       // box: local ids of the vertices on the x-faces (Dirichlet set of the benchmarks)
};

namespace {

typedef long long i64;

std::vector<int> plane_split(int planes, int parts)
{
  std::vector<int> p(parts + 1);
  for (int r = 0; r <= parts; ++r) p[r] = (int)std::llround((double)r * planes / parts);
  return p;
}

// Shared core. `enumerate(fn)` calls fn(v, all_owned) for every candidate cell (4 global vertex ids;
// all_owned = the caller already knows that all four vertices belong to this rank — the interior
// of a brick — so the cell needs no ownership test). It is called twice: once to find the ghost
// vertices / send lists and to count, once to fill the local cell array. owner(g) = owning rank of
// a vertex that occurs in a candidate cell, local_owned(g) = local index of an OWNED vertex.
template <typename Enum, typename OwnerFn, typename LocalFn, typename CoordFn>
void build_partition(pnpb200_partition& P, int rank, Enum enumerate, OwnerFn owner, LocalFn local_owned, CoordFn coord)
{
  // P.gids[0 .. n_own) = the owned vertices (ascending global id), set by the caller
  P.n_own = (int)P.gids.size();
  std::vector<std::pair<int, i64>> ghosts, sends;
  size_t nkept = 0;
  enumerate([&](const i64* v, bool all_owned) {
    if (all_owned) { ++nkept; return; }
    int o[4]; bool touch = false;
    for (int r = 0; r < 4; ++r) { o[r] = owner(v[r]); touch = touch || o[r] == rank; }
    if (!touch) return;
    ++nkept;
    for (int r = 0; r < 4; ++r) {
      if (o[r] != rank) { ghosts.emplace_back(o[r], v[r]); continue; }
      // an owned vertex goes to rank k iff it shares a cell with a vertex owned by k: that cell touches
      // a k-owned vertex, so it is one of k's local cells and this vertex one of k's ghosts
      for (int q = 0; q < 4; ++q) if (o[q] != rank) sends.emplace_back(o[q], v[r]);
    }
    if (ghosts.size() > (1u << 22)) { std::sort(ghosts.begin(), ghosts.end()); ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end()); }
    if (sends.size() > (1u << 22)) { std::sort(sends.begin(), sends.end()); sends.erase(std::unique(sends.begin(), sends.end()), sends.end()); }
  });
  std::sort(ghosts.begin(), ghosts.end()); ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
  std::sort(sends.begin(), sends.end()); sends.erase(std::unique(sends.begin(), sends.end()), sends.end());
  if (nkept > 0x7fffffffull / 4) throw std::runtime_error("partition: more than 2^31 / 4 local cells");
  P.n_loc = P.n_own + (int)ghosts.size();
  P.n_cells = (int)nkept;
  for (auto& og : ghosts) P.gids.push_back(og.second);
  P.nbr.clear(); P.recv_ptr.assign(1, P.n_own);
  for (size_t q = 0; q < ghosts.size(); ++q)
    if (q == 0 || ghosts[q].first != ghosts[q - 1].first) { if (q > 0) P.recv_ptr.push_back(P.n_own + (int)q); P.nbr.push_back(ghosts[q].first); }
  if (!ghosts.empty()) P.recv_ptr.push_back(P.n_loc);
  std::vector<std::pair<i64, int>> glook(ghosts.size());      // ghost gid -> local id
  for (size_t q = 0; q < ghosts.size(); ++q) glook[q] = {ghosts[q].second, P.n_own + (int)q};
  std::sort(glook.begin(), glook.end());
  auto local_ghost = [&](i64 g) -> int {
    auto jt = std::lower_bound(glook.begin(), glook.end(), std::make_pair(g, -1));
    if (jt == glook.end() || jt->first != g) throw std::runtime_error("partition: vertex of a local cell is neither owned nor ghost");
    return jt->second;
  };
  P.cells.resize(4 * nkept);
  size_t w = 0;
  enumerate([&](const i64* vin, bool all_owned) {
    i64 v[4] = {vin[0], vin[1], vin[2], vin[3]};
    int o[4] = {rank, rank, rank, rank};
    if (!all_owned) {
      bool touch = false;
      for (int r = 0; r < 4; ++r) { o[r] = owner(v[r]); touch = touch || o[r] == rank; }
      if (!touch) return;
    }
    // UFC order: ascending GLOBAL vertex ids
    for (int a = 1; a < 4; ++a) for (int b = a; b > 0 && v[b] < v[b - 1]; --b) { std::swap(v[b], v[b - 1]); std::swap(o[b], o[b - 1]); }
    for (int r = 0; r < 4; ++r) P.cells[w++] = o[r] == rank ? local_owned(v[r]) : local_ghost(v[r]);
  });
  if (w != 4 * nkept) throw std::runtime_error("partition: cell enumeration is not repeatable");
  // the neighbour relation is symmetric (a shared cell has vertices of both ranks): send groups follow P.nbr
  P.send_ptr.assign(1, 0); P.send_idx.clear();
  size_t q = 0;
  for (size_t k = 0; k < P.nbr.size(); ++k) {
    if (q < sends.size() && sends[q].first < P.nbr[k]) throw std::runtime_error("partition: asymmetric neighbour relation");
    while (q < sends.size() && sends[q].first == P.nbr[k]) { P.send_idx.push_back(local_owned(sends[q].second)); ++q; }
    P.send_ptr.push_back((int)P.send_idx.size());
  }
  if (q != sends.size()) throw std::runtime_error("partition: asymmetric neighbour relation");
  P.xyz.resize(3 * (size_t)P.n_loc);
  for (int l = 0; l < P.n_loc; ++l) coord(P.gids[l], &P.xyz[3 * (size_t)l]);
}

template <typename F>
int guarded(F&& f)
{
  try { f(); return FASP_SUCCESS; }
  catch (const std::bad_alloc&) { fprintf(stderr, "### pnpb200 ERROR: partition: out of host memory\n"); return ERROR_ALLOC_MEM; }
  catch (const std::exception& e) { fprintf(stderr, "### pnpb200 ERROR: %s\n", e.what()); return ERROR_INPUT_PAR; }
}

} // namespace

extern "C" {

int pnpb200_partition_box(int nx, int ny, int nz, double lx, double ly, double lz, int px, int py, int pz, int rank,
                          pnpb200_partition** out)
{
  if (!out) return ERROR_INPUT_PAR;
  *out = nullptr;
  if (nx < 1 || ny < 1 || nz < 1 || px < 1 || py < 1 || pz < 1 || rank < 0 || rank >= px * py * pz) return ERROR_INPUT_PAR;
  if (px > nx + 1 || py > ny + 1 || pz > nz + 1) return ERROR_INPUT_PAR;      // every rank needs at least one vertex plane per direction
  pnpb200_partition* P = new pnpb200_partition;
  const int rc = guarded([&]() {
    const std::vector<int> sx = plane_split(nx + 1, px), sy = plane_split(ny + 1, py), sz = plane_split(nz + 1, pz);
    // rank = rx + px * (ry + py * rz): x fastest, like the vertex numbering
    const int rx = rank % px, ry = (rank / px) % py, rz = rank / (px * py);
    const int x0 = sx[rx], x1 = sx[rx + 1], y0 = sy[ry], y1 = sy[ry + 1], z0 = sz[rz], z1 = sz[rz + 1];
    const i64 NX = nx + 1, NY = ny + 1;
    auto part_of = [](const std::vector<int>& s, int i) { return (int)(std::upper_bound(s.begin(), s.end(), i) - s.begin()) - 1; };
    auto owner = [&](i64 g) -> int {
      const int i = (int)(g % NX), j = (int)((g / NX) % NY), k = (int)(g / (NX * NY));
      return part_of(sx, i) + px * (part_of(sy, j) + py * part_of(sz, k));
    };
    auto coord = [&](i64 g, double* x) {      // dolfin::BoxMesh: a + i * (b - a) / n, a = -l/2, b = +l/2
      const i64 i = g % NX, j = (g / NX) % NY, k = g / (NX * NY);
      x[0] = -lx / 2 + (double)i * (lx / 2 - (-lx / 2)) / nx;
      x[1] = -ly / 2 + (double)j * (ly / 2 - (-ly / 2)) / ny;
      x[2] = -lz / 2 + (double)k * (lz / 2 - (-lz / 2)) / nz;
    };
    P->gids.reserve((size_t)(x1 - x0) * (y1 - y0) * (z1 - z0));
    for (int k = z0; k < z1; ++k) for (int j = y0; j < y1; ++j) for (int i = x0; i < x1; ++i) P->gids.push_back(i + NX * (j + NY * (i64)k));
    const i64 bx = x1 - x0, by = y1 - y0;
    auto local_owned = [&](i64 g) -> int {
      const i64 i = g % NX, j = (g / NX) % NY, k = g / (NX * NY);
      return (int)((i - x0) + bx * ((j - y0) + by * (k - z0)));
    };
    // candidate cubes: every cube with a corner in the owned brick; the 6 tetrahedra of a cube around
    // the diagonal v0-v7. Cubes whose 8 corners are all owned need no ownership test.
    const int cx0 = std::max(x0 - 1, 0), cx1 = std::min(x1, nx), cy0 = std::max(y0 - 1, 0), cy1 = std::min(y1, ny),
              cz0 = std::max(z0 - 1, 0), cz1 = std::min(z1, nz);
    static const int T[6][4] = {{0, 1, 3, 7}, {0, 1, 7, 5}, {0, 5, 7, 4}, {0, 3, 2, 7}, {0, 6, 4, 7}, {0, 2, 6, 7}};
    auto enumerate = [&](auto&& fn) {
      for (int k = cz0; k < cz1; ++k) for (int j = cy0; j < cy1; ++j) for (int i = cx0; i < cx1; ++i) {
        const i64 v0 = i + NX * (j + NY * (i64)k);
        const i64 v[8] = {v0, v0 + 1, v0 + NX, v0 + 1 + NX, v0 + NX * NY, v0 + 1 + NX * NY, v0 + NX + NX * NY, v0 + 1 + NX + NX * NY};
        const bool inside = i >= x0 && i + 1 < x1 && j >= y0 && j + 1 < y1 && k >= z0 && k + 1 < z1;
        for (int t = 0; t < 6; ++t) { const i64 c[4] = {v[T[t][0]], v[T[t][1]], v[T[t][2]], v[T[t][3]]}; fn(c, inside); }
      }
    };
    build_partition(*P, rank, enumerate, owner, local_owned, coord);
    P->color.resize(P->n_loc);
    for (int l = 0; l < P->n_loc; ++l) {
      const i64 g = P->gids[l];
      const i64 i = g % NX, j = (g / NX) % NY, k = g / (NX * NY);
      P->color[l] = (int)((i + j + k) % 4);
      if (i == 0 || i == nx) P->bc_vertices.push_back(l);
    }
  });
  if (rc != FASP_SUCCESS) { delete P; return rc; }
  *out = P;
  return FASP_SUCCESS;
}

int pnpb200_partition_mesh(int nv, const double* xyz, int nc, const int* cells, int nranks, int rank, pnpb200_partition** out)
{
  if (!out) return ERROR_INPUT_PAR;
  *out = nullptr;
  if (nv < 4 || nc < 1 || !xyz || !cells || nranks < 1 || rank < 0 || rank >= nranks || nranks > nv) return ERROR_INPUT_PAR;
  pnpb200_partition* P = new pnpb200_partition;
  const int rc = guarded([&]() {
    // recursive coordinate bisection: split the longest axis of the part's bounding box so that the two
    // halves get vertex counts proportional to the ranks they will hold; ties broken by vertex id
    std::vector<int> own(nv, 0), idx(nv);
    std::iota(idx.begin(), idx.end(), 0);
    struct Job { int lo, hi, r0, nr; };
    std::vector<Job> stack{{0, nv, 0, nranks}};
    while (!stack.empty()) {
      const Job jb = stack.back(); stack.pop_back();
      if (jb.nr == 1) { for (int q = jb.lo; q < jb.hi; ++q) own[idx[q]] = jb.r0; continue; }
      double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
      for (int q = jb.lo; q < jb.hi; ++q) for (int d = 0; d < 3; ++d) {
        const double c = xyz[3 * (size_t)idx[q] + d];
        mn[d] = std::min(mn[d], c); mx[d] = std::max(mx[d], c);
      }
      int ax = 0;
      for (int d = 1; d < 3; ++d) if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
      const int n1 = jb.nr / 2;
      const int cut = jb.lo + (int)std::llround((double)(jb.hi - jb.lo) * n1 / jb.nr);
      std::nth_element(idx.begin() + jb.lo, idx.begin() + cut, idx.begin() + jb.hi, [&](int a, int b) {
        const double ca = xyz[3 * (size_t)a + ax], cb = xyz[3 * (size_t)b + ax];
        return ca < cb || (ca == cb && a < b);
      });
      stack.push_back({jb.lo, cut, jb.r0, n1});
      stack.push_back({cut, jb.hi, jb.r0 + n1, jb.nr - n1});
    }
    std::vector<int> g2l(nv, -1);
    for (int v = 0; v < nv; ++v) if (own[v] == rank) { g2l[v] = (int)P->gids.size(); P->gids.push_back(v); }
    if (P->gids.empty()) throw std::runtime_error("partition: a rank received no vertex");
    for (size_t e = 0; e < 4 * (size_t)nc; ++e) if (cells[e] < 0 || cells[e] >= nv) throw std::runtime_error("partition: cell vertex out of range");
    auto enumerate = [&](auto&& fn) {
      for (int c = 0; c < nc; ++c) {
        const i64 v[4] = {cells[4 * (size_t)c], cells[4 * (size_t)c + 1], cells[4 * (size_t)c + 2], cells[4 * (size_t)c + 3]};
        fn(v, false);
      }
    };
    auto owner = [&](i64 g) -> int { return own[(size_t)g]; };
    auto local_owned = [&](i64 g) -> int { return g2l[(size_t)g]; };
    auto coord = [&](i64 g, double* x) { for (int d = 0; d < 3; ++d) x[d] = xyz[3 * (size_t)g + d]; };
    build_partition(*P, rank, enumerate, owner, local_owned, coord);
  });
  if (rc != FASP_SUCCESS) { delete P; return rc; }
  *out = P;
  return FASP_SUCCESS;
}

int pnpb200_partition_sizes(const pnpb200_partition* P, int* n_own, int* n_loc, int* n_cells, int* n_neighbors, int* n_send, int* n_bc)
{
  if (!P) return ERROR_INPUT_PAR;
  if (n_own) *n_own = P->n_own;
  if (n_loc) *n_loc = P->n_loc;
  if (n_cells) *n_cells = P->n_cells;
  if (n_neighbors) *n_neighbors = (int)P->nbr.size();
  if (n_send) *n_send = (int)P->send_idx.size();
  if (n_bc) *n_bc = (int)P->bc_vertices.size();
  return FASP_SUCCESS;
}

int pnpb200_partition_arrays(const pnpb200_partition* P, const double** xyz, const int** cells, const long long** gids,
                             const int** neighbor_ranks, const int** send_ptr, const int** send_idx, const int** recv_ptr,
                             const int** color, const int** bc_vertices)
{
  if (!P) return ERROR_INPUT_PAR;
  if (xyz) *xyz = P->xyz.data();
  if (cells) *cells = P->cells.data();
  if (gids) *gids = P->gids.data();
  if (neighbor_ranks) *neighbor_ranks = P->nbr.data();
  if (send_ptr) *send_ptr = P->send_ptr.data();
  if (send_idx) *send_idx = P->send_idx.data();
  if (recv_ptr) *recv_ptr = P->recv_ptr.data();
  if (color) *color = P->color.empty() ? nullptr : P->color.data();
  if (bc_vertices) *bc_vertices = P->bc_vertices.data();
  return FASP_SUCCESS;
}

void pnpb200_partition_free(pnpb200_partition* P) { delete P; }

}
