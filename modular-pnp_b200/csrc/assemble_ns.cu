// pnp-b200 — device assembly of the PNP-Stokes Newton system (include/pnpb200/pnpb200_ns.h).
// Replaces, for benchmarks/pnp_stokes, what PDE::setup_linear_algebra (src/pde.cpp:542-555) does through
// dolfin::assemble with the generated kernels of vector_linear_pnp_ns_forms.h (cell + interior-facet integrals of a and
// L) followed by DirichletBC::apply: the mixed-space matrix and right-hand side that Linear_PNP_NS::
// setup_fasp_linear_algebra (linear_pnp_ns.cpp:137-176) then cuts into the 2 x 2 block system.
//   host, once per mesh : UFC vertex ordering, facet numbering (dolfin::Mesh::init(2)), vertex -> cell adjacency, the
//                         CSR pattern of the mixed space (cell couplings + the RT-RT couplings of the facet macro
//                         elements)
//   device, per Newton step : ns_moments_kernel (thread per cell), ns_rows_kernel (thread per matrix row; ns_forms.cuh)
#include "common.cuh"
#include "ns_forms.cuh"
#include "pnpb200/pnpb200_ns.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>

using namespace pnpb200;

struct pnpb200_ns_system {
  int device = 0;
  cudaStream_t stream = nullptr;
  int nv = 0, nc = 0, nf = 0, nif = 0, ndof = 0, nnz = 0;
  std::vector<int> cells, cell_facets, facet_cells, ia, ja;      // host copies (topology is also handed to the caller)
  DevBuf<double> xyz, cc, uu, pp, mom, val, rhs, partial;
  DevBuf<int> d_cells, d_cell_facets, d_facet_cells, vcell_ptr, vcell, d_ia, d_ja;
  DevBuf<unsigned char> bc;
  bool has_bc = false, assembled = false;
  double par[8] = {1.0, 1.0, 1.0, 1.0, -1.0, 0.1, 1.0, 1.0};     // benchmarks/pnp_stokes/main.cpp:123-132
  NsView view() const
  {
    NsView S;
    S.nv = nv; S.nc = nc; S.nf = nf; S.ndof = ndof;
    S.xyz = xyz.p; S.cells = d_cells.p; S.cell_facets = d_cell_facets.p; S.facet_cells = d_facet_cells.p;
    S.vcell_ptr = vcell_ptr.p; S.vcell = vcell.p; S.cc = cc.p; S.uu = uu.p; S.pp = pp.p;
    for (int i = 0; i < 8; ++i) S.par[i] = par[i];
    S.mom = mom.p; S.ia = d_ia.p; S.ja = d_ja.p; S.val = val.p; S.rhs = rhs.p; S.bc = has_bc ? bc.p : nullptr;
    return S;
  }
};

namespace {

thread_local std::string g_ns_error;

template <typename F>
int guarded(F&& f)
{
  try { return f(); }
  catch (const Error& e) { g_ns_error = e.what(); fprintf(stderr, "### pnpb200 ERROR: %s\n", e.what()); return e.code; }
  catch (const std::exception& e) { g_ns_error = e.what(); fprintf(stderr, "### pnpb200 ERROR: %s\n", e.what()); return ERROR_UNKNOWN; }
}

__global__ void __launch_bounds__(128) ns_moments_kernel(NsView S)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < S.nc) ns_cell_moments(S, c);
}

__global__ void __launch_bounds__(128) ns_rows_kernel(NsView S, bool matrix)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < S.ndof) ns_assemble_row(S, r, matrix);
}

// per-block partial sums of squares and maxima of |b| (fixed tree: reproducible); the host adds the few hundred partials
__global__ void __launch_bounds__(256) ns_norm_kernel(int n, const double* b, double* partial)
{
  __shared__ double s2[256], sm[256];
  double a2 = 0.0, am = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { a2 += b[i] * b[i]; am = fmax(am, fabs(b[i])); }
  s2[threadIdx.x] = a2; sm[threadIdx.x] = am;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { s2[threadIdx.x] += s2[threadIdx.x + o]; sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s2[0]; partial[2 * blockIdx.x + 1] = sm[0]; }
}

__global__ void ns_update_kernel(int nv3, int nf, int nc, const double* du, double alpha, double* cc, double* uu, double* pp)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nv3) cc[i] += alpha * du[i];
  else if (i < nv3 + nf) uu[i - nv3] += alpha * du[i];
  else if (i < nv3 + nf + nc) pp[i - nv3 - nf] += alpha * du[i];
}

void build_topology(pnpb200_ns_system& s, const double* xyz, const int* cells_in)
{
  const int nc = s.nc, nv = s.nv;
  s.cells.assign(cells_in, cells_in + 4 * (size_t)nc);
  for (int c = 0; c < nc; ++c) {
    int* v = &s.cells[4 * (size_t)c];
    std::sort(v, v + 4);
    if (v[0] < 0 || v[3] >= nv || v[0] == v[1] || v[1] == v[2] || v[2] == v[3]) throw Error(ERROR_INPUT_PAR, "pnpb200_ns_create: bad cell");
  }
  // facets: facet i of a cell is opposite its local vertex i; numbered in lexicographic order of the vertex triples
  struct Face { std::array<int, 3> key; int cell, loc; };
  std::vector<Face> faces(4 * (size_t)nc);
  for (int c = 0; c < nc; ++c) for (int i = 0; i < 4; ++i) {
    Face& f = faces[4 * (size_t)c + i]; int n = 0;
    for (int j = 0; j < 4; ++j) if (j != i) f.key[n++] = s.cells[4 * (size_t)c + j];
    f.cell = c; f.loc = i;
  }
  std::sort(faces.begin(), faces.end(), [](const Face& a, const Face& b) { return a.key != b.key ? a.key < b.key : a.cell < b.cell; });
  s.cell_facets.assign(4 * (size_t)nc, -1);
  s.facet_cells.clear();
  int nf = 0, nif = 0;
  for (size_t p = 0; p < faces.size();) {
    size_t e = p + 1;
    while (e < faces.size() && faces[e].key == faces[p].key) ++e;
    if (e - p > 2) throw Error(ERROR_INPUT_PAR, "pnpb200_ns_create: a facet is shared by more than two cells");
    s.facet_cells.push_back(faces[p].cell);
    s.facet_cells.push_back(e - p == 2 ? faces[p + 1].cell : -1);
    for (size_t t = p; t < e; ++t) s.cell_facets[4 * (size_t)faces[t].cell + faces[t].loc] = nf;
    if (e - p == 2) ++nif;
    ++nf; p = e;
  }
  s.nf = nf; s.nif = nif;
  s.ndof = 3 * nv + nf + nc;
  // vertex -> cells
  std::vector<int> ptr(nv + 1, 0), adj(4 * (size_t)nc);
  for (size_t i = 0; i < s.cells.size(); ++i) ++ptr[s.cells[i] + 1];
  for (int v = 0; v < nv; ++v) ptr[v + 1] += ptr[v];
  { std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (int c = 0; c < nc; ++c) for (int i = 0; i < 4; ++i) adj[fill[s.cells[4 * (size_t)c + i]]++] = c; }
  // pattern of the mixed space
  const int o_u = 3 * nv, o_p = 3 * nv + nf;
  s.ia.assign(s.ndof + 1, 0);
  std::vector<std::vector<int>> rows(s.ndof);
  std::vector<int> cols;
  for (int v = 0; v < nv; ++v) {                     // P1 rows: P1 and RT dofs of the cells around the vertex
    cols.clear();
    for (int p = ptr[v]; p < ptr[v + 1]; ++p) {
      const int c = adj[p];
      for (int j = 0; j < 4; ++j) {
        const int w = s.cells[4 * (size_t)c + j];
        cols.push_back(3 * w); cols.push_back(3 * w + 1); cols.push_back(3 * w + 2);
        cols.push_back(o_u + s.cell_facets[4 * (size_t)c + j]);
      }
    }
    std::sort(cols.begin(), cols.end()); cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    for (int k = 0; k < 3; ++k) rows[3 * v + k] = cols;
  }
  auto cell_dofs = [&](int c, std::vector<int>& out) {
    for (int j = 0; j < 4; ++j) {
      const int w = s.cells[4 * (size_t)c + j];
      out.push_back(3 * w); out.push_back(3 * w + 1); out.push_back(3 * w + 2);
      out.push_back(o_u + s.cell_facets[4 * (size_t)c + j]);
    }
    out.push_back(o_p + c);
  };
  for (int f = 0; f < nf; ++f) {                     // velocity rows: all dofs of the 1-2 cells + RT dofs of their facet neighbours
    cols.clear();
    for (int t = 0; t < 2; ++t) {
      const int c = s.facet_cells[2 * (size_t)f + t];
      if (c < 0) continue;
      cell_dofs(c, cols);
      for (int j = 0; j < 4; ++j) {
        const int F = s.cell_facets[4 * (size_t)c + j];
        for (int t2 = 0; t2 < 2; ++t2) {
          const int c2 = s.facet_cells[2 * (size_t)F + t2];
          if (c2 < 0 || c2 == c) continue;
          for (int jj = 0; jj < 4; ++jj) cols.push_back(o_u + s.cell_facets[4 * (size_t)c2 + jj]);
        }
      }
    }
    std::sort(cols.begin(), cols.end()); cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    rows[o_u + f] = cols;
  }
  for (int c = 0; c < nc; ++c) {                     // pressure rows: the cell's RT dofs and the (zero) diagonal
    cols.clear();
    for (int j = 0; j < 4; ++j) cols.push_back(o_u + s.cell_facets[4 * (size_t)c + j]);
    cols.push_back(o_p + c);
    std::sort(cols.begin(), cols.end());
    rows[o_p + c] = cols;
  }
  size_t nnz = 0;
  for (int r = 0; r < s.ndof; ++r) { nnz += rows[r].size(); s.ia[r + 1] = (int)nnz; }
  if (nnz > 2147483647u) throw Error(ERROR_INPUT_PAR, "pnpb200_ns_create: more than 2^31 matrix entries");
  s.nnz = (int)nnz;
  s.ja.resize(nnz);
  for (int r = 0; r < s.ndof; ++r) std::copy(rows[r].begin(), rows[r].end(), s.ja.begin() + s.ia[r]);

  cudaStream_t st = s.stream;
  auto up_i = [&](DevBuf<int>& d, const std::vector<int>& h) { d.alloc(h.size(), st); d.upload(h.data(), h.size()); };
  s.xyz.alloc(3 * (size_t)nv, st); s.xyz.upload(xyz, 3 * (size_t)nv);
  up_i(s.d_cells, s.cells); up_i(s.d_cell_facets, s.cell_facets); up_i(s.d_facet_cells, s.facet_cells);
  up_i(s.vcell_ptr, ptr); up_i(s.vcell, adj); up_i(s.d_ia, s.ia); up_i(s.d_ja, s.ja);
  s.cc.alloc(3 * (size_t)nv, st); s.uu.alloc(nf, st); s.pp.alloc(nc, st);
  s.cc.zero(); s.uu.zero(); s.pp.zero();
  s.mom.alloc((size_t)NS_MOM * nc, st); s.val.alloc(nnz, st); s.rhs.alloc(s.ndof, st);
  s.bc.alloc(s.ndof, st); s.bc.zero();
  s.partial.alloc(2 * 512, st);
  PNP_CUDA(cudaStreamSynchronize(st));               // the host vectors above go out of scope
}

void run_assembly(pnpb200_ns_system& s, bool matrix)
{
  PNP_CUDA(cudaSetDevice(s.device));
  const NsView S = s.view();
  ns_moments_kernel<<<cdiv(s.nc, 128), 128, 0, s.stream>>>(S);
  PNP_LAUNCH_CHECK(); PNP_COUNT_LAUNCH();
  ns_rows_kernel<<<cdiv(s.ndof, 128), 128, 0, s.stream>>>(S, matrix);
  PNP_LAUNCH_CHECK(); PNP_COUNT_LAUNCH();
}

}  // namespace

extern "C" {

int pnpb200_ns_create(pnpb200_ns_system** out, int n_vertices, const double* xyz, int n_cells, const int* cells)
{
  if (!out) return ERROR_INPUT_PAR;
  *out = nullptr;
  return guarded([&]() -> int {
    if (!xyz || !cells || n_vertices < 4 || n_cells < 1) throw Error(ERROR_INPUT_PAR, "pnpb200_ns_create: empty mesh");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      throw Error(ERROR_DEVICE, "no CUDA device available (pnp-b200 has no CPU fallback)");
    auto* s = new pnpb200_ns_system;
    try {
      PNP_CUDA(cudaGetDevice(&s->device));
      PNP_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
      s->nv = n_vertices; s->nc = n_cells;
      build_topology(*s, xyz, cells);
    } catch (...) { if (s->stream) cudaStreamDestroy(s->stream); delete s; throw; }
    *out = s;
    return FASP_SUCCESS;
  });
}

void pnpb200_ns_destroy(pnpb200_ns_system* s)
{
  if (!s) return;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  cudaStream_t st = s->stream;
  delete s;
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
}

int pnpb200_ns_sizes(const pnpb200_ns_system* s, int* n_facets, int* n_interior_facets, int* n_dofs, int* nnz)
{
  if (!s) return ERROR_INPUT_PAR;
  if (n_facets) *n_facets = s->nf;
  if (n_interior_facets) *n_interior_facets = s->nif;
  if (n_dofs) *n_dofs = s->ndof;
  if (nnz) *nnz = s->nnz;
  return FASP_SUCCESS;
}

int pnpb200_ns_topology(const pnpb200_ns_system* s, int* cells_sorted, int* cell_facets, int* facet_cells)
{
  if (!s) return ERROR_INPUT_PAR;
  if (cells_sorted) std::memcpy(cells_sorted, s->cells.data(), s->cells.size() * sizeof(int));
  if (cell_facets) std::memcpy(cell_facets, s->cell_facets.data(), s->cell_facets.size() * sizeof(int));
  if (facet_cells) std::memcpy(facet_cells, s->facet_cells.data(), s->facet_cells.size() * sizeof(int));
  return FASP_SUCCESS;
}

int pnpb200_ns_set_parameters(pnpb200_ns_system* s, const double par8[8])
{
  if (!s || !par8) return ERROR_INPUT_PAR;
  for (int i = 0; i < 8; ++i) if (!std::isfinite(par8[i])) return ERROR_INPUT_PAR;
  std::memcpy(s->par, par8, sizeof s->par);
  s->assembled = false;
  return FASP_SUCCESS;
}

int pnpb200_ns_set_dirichlet(pnpb200_ns_system* s, int n_bc, const int* bc_dofs)
{
  if (!s || n_bc < 0 || (n_bc > 0 && !bc_dofs)) return ERROR_INPUT_PAR;
  return guarded([&]() -> int {
    PNP_CUDA(cudaSetDevice(s->device));
    std::vector<unsigned char> flags(s->ndof, 0);
    for (int i = 0; i < n_bc; ++i) {
      if (bc_dofs[i] < 0 || bc_dofs[i] >= s->ndof) throw Error(ERROR_INPUT_PAR, "pnpb200_ns_set_dirichlet: dof out of range");
      flags[bc_dofs[i]] = 1;
    }
    s->bc.upload(flags.data(), flags.size());
    PNP_CUDA(cudaStreamSynchronize(s->stream));
    s->has_bc = n_bc > 0; s->assembled = false;
    return FASP_SUCCESS;
  });
}

int pnpb200_ns_set_solution(pnpb200_ns_system* s, const double* cc, const double* uu, const double* pp)
{
  if (!s || !cc || !uu || !pp) return ERROR_INPUT_PAR;
  return guarded([&]() -> int {
    PNP_CUDA(cudaSetDevice(s->device));
    s->cc.upload(cc, 3 * (size_t)s->nv); s->uu.upload(uu, s->nf); s->pp.upload(pp, s->nc);
    PNP_CUDA(cudaStreamSynchronize(s->stream));
    s->assembled = false;
    return FASP_SUCCESS;
  });
}

int pnpb200_ns_get_solution(pnpb200_ns_system* s, double* cc, double* uu, double* pp)
{
  if (!s) return ERROR_INPUT_PAR;
  return guarded([&]() -> int {
    PNP_CUDA(cudaSetDevice(s->device));
    if (cc) s->cc.download(cc, 3 * (size_t)s->nv);
    if (uu) s->uu.download(uu, s->nf);
    if (pp) s->pp.download(pp, s->nc);
    return FASP_SUCCESS;
  });
}

int pnpb200_ns_assemble(pnpb200_ns_system* s)
{
  if (!s) return ERROR_INPUT_PAR;
  return guarded([&]() -> int { run_assembly(*s, true); s->assembled = true; return FASP_SUCCESS; });
}

int pnpb200_ns_get_system(pnpb200_ns_system* s, dCSRmat* A, dvector* b)
{
  if (!s || (!A && !b)) return ERROR_INPUT_PAR;
  return guarded([&]() -> int {
    if (!s->assembled) throw Error(ERROR_INPUT_PAR, "pnpb200_ns_get_system: call pnpb200_ns_assemble first");
    PNP_CUDA(cudaSetDevice(s->device));
    if (A) {
      if (A->row != s->ndof || A->col != s->ndof || A->nnz != s->nnz || !A->IA || !A->JA || !A->val)
        throw Error(ERROR_INPUT_PAR, "pnpb200_ns_get_system: A must be allocated with the sizes of pnpb200_ns_sizes");
      std::memcpy(A->IA, s->ia.data(), (s->ndof + 1) * sizeof(int));
      std::memcpy(A->JA, s->ja.data(), (size_t)s->nnz * sizeof(int));
      s->val.download(A->val, s->nnz);
    }
    if (b) {
      if (b->row != s->ndof || !b->val) throw Error(ERROR_INPUT_PAR, "pnpb200_ns_get_system: b must have n_dofs entries");
      s->rhs.download(b->val, s->ndof);
    }
    return FASP_SUCCESS;
  });
}

int pnpb200_ns_residual(pnpb200_ns_system* s, double* l2, double* linf)
{
  if (!s) return ERROR_INPUT_PAR;
  return guarded([&]() -> int {
    run_assembly(*s, false);                         // L only; Dirichlet rows zero (src/pde.cpp:377-397)
    const int nb = std::min(512, cdiv(s->ndof, 256));
    ns_norm_kernel<<<nb, 256, 0, s->stream>>>(s->ndof, s->rhs.p, s->partial.p);
    PNP_LAUNCH_CHECK(); PNP_COUNT_LAUNCH();
    const std::vector<double> part = s->partial.to_host(2 * (size_t)nb);
    double a2 = 0.0, am = 0.0;
    for (int i = 0; i < nb; ++i) { a2 += part[2 * i]; am = std::max(am, part[2 * i + 1]); }
    if (l2) *l2 = std::sqrt(a2);
    if (linf) *linf = am;
    return FASP_SUCCESS;
  });
}

int pnpb200_ns_add_update(pnpb200_ns_system* s, const double* du, double alpha)
{
  if (!s || !du) return ERROR_INPUT_PAR;
  return guarded([&]() -> int {
    PNP_CUDA(cudaSetDevice(s->device));
    DevBuf<double> d; d.alloc(s->ndof, s->stream); d.upload(du, s->ndof);
    ns_update_kernel<<<cdiv(s->ndof, 256), 256, 0, s->stream>>>(3 * s->nv, s->nf, s->nc, d.p, alpha, s->cc.p, s->uu.p, s->pp.p);
    PNP_LAUNCH_CHECK(); PNP_COUNT_LAUNCH();
    PNP_CUDA(cudaStreamSynchronize(s->stream));
    s->assembled = false;
    return FASP_SUCCESS;
  });
}

}  // extern "C"
