// TEST INFRASTRUCTURE ONLY (oracle/) — see pnp_oracle.h.
// CPU restatement of the DOLFIN slice the reference's hot path drives [EXT: DOLFIN 2016.x,
// not in /root/reference]: BoxMesh (src/domain.cpp:306-308), blocked P1^3 dofs, cell-based
// sparsity, cell-loop assembly (src/pde.cpp:542-555, :377-397), DirichletBC::apply
// (src/pde.cpp:548-551), the EAFE overwrite (benchmarks/pnp_exact_solution/linear_pnp.cpp:195-285,
// BC re-apply :78-80), the zero-row fix (src/pde.cpp:596-621) and CSR->BSR
// (fasp_format_dcsr_dbsr, linear_pnp.cpp:85).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>
#include "pnp_oracle.h"
#include "oracle_internal.hpp"

using namespace orc;

extern "C" {

// ---- BoxMesh [EXT DOLFIN generation/BoxMesh.cpp]: lattice vertices, 6 tets per cube around
// the v0-v7 diagonal, cell vertices sorted ascending (UFC order) -------------------------------
void orc_box_mesh_counts(int nx, int ny, int nz, int* nv, int* nc)
{
  *nv = (nx + 1) * (ny + 1) * (nz + 1);
  *nc = 6 * nx * ny * nz;
}

void orc_box_mesh(int nx, int ny, int nz, double lx, double ly, double lz, double* xyz, int* cells)
{
  // src/domain.cpp:306-307: p0 = -L/2, p1 = +L/2
  const double a = -lx / 2, b = lx / 2, c = -ly / 2, d = ly / 2, e = -lz / 2, f = lz / 2;
  size_t v = 0;
  for (int iz = 0; iz <= nz; ++iz) {
    const double z = e + ((double)iz) * (f - e) / (double)nz;
    for (int iy = 0; iy <= ny; ++iy) {
      const double y = c + ((double)iy) * (d - c) / (double)ny;
      for (int ix = 0; ix <= nx; ++ix) {
        const double x = a + ((double)ix) * (b - a) / (double)nx;
        xyz[3 * v] = x; xyz[3 * v + 1] = y; xyz[3 * v + 2] = z; ++v;
      }
    }
  }
  size_t cidx = 0;
  for (int iz = 0; iz < nz; ++iz)
    for (int iy = 0; iy < ny; ++iy)
      for (int ix = 0; ix < nx; ++ix) {
        const int v0 = iz * (nx + 1) * (ny + 1) + iy * (nx + 1) + ix;
        const int v1 = v0 + 1, v2 = v0 + (nx + 1), v3 = v1 + (nx + 1);
        const int v4 = v0 + (nx + 1) * (ny + 1), v5 = v1 + (nx + 1) * (ny + 1);
        const int v6 = v2 + (nx + 1) * (ny + 1), v7 = v3 + (nx + 1) * (ny + 1);
        const int t[6][4] = {{v0, v1, v3, v7}, {v0, v1, v7, v5}, {v0, v5, v7, v4},
                             {v0, v3, v2, v7}, {v0, v6, v4, v7}, {v0, v2, v6, v7}};
        for (int k = 0; k < 6; ++k) {
          int s[4] = {t[k][0], t[k][1], t[k][2], t[k][3]};
          std::sort(s, s + 4);
          for (int r = 0; r < 4; ++r) cells[4 * cidx + r] = s[r];
          ++cidx;
        }
      }
}

// ---- Dirichlet vertex set: src/pde.cpp:67-87 (bounding box, eps = 0.1*rmin),
// src/dirichlet.cpp:24-37 (x[d] < min+eps or > max-eps, on_boundary), topological
// DirichletBC [EXT]: dofs of exterior facets whose vertices all satisfy the predicate. ---------
int orc_dirichlet_vertices(int nv, const double* xyz, int nc, const int* cells, int coord,
                           unsigned char* flags)
{
  double mn[3] = {1e20, 1e20, 1e20}, mx[3] = {-1e20, -1e20, -1e20};
  for (int v = 0; v < nv; ++v)
    for (int d = 0; d < 3; ++d) {
      mn[d] = std::min(mn[d], xyz[3 * v + d]);
      mx[d] = std::max(mx[d], xyz[3 * v + d]);
    }
  // rmin = smallest inradius = 3 V / (sum of face areas)
  double rmin = 1e300;
  for (int c = 0; c < nc; ++c) {
    const int* cv = cells + 4 * c;
    double p[4][3];
    for (int r = 0; r < 4; ++r) for (int d = 0; d < 3; ++d) p[r][d] = xyz[3 * cv[r] + d];
    auto sub = [](const double* a, const double* b, double* o) { for (int d = 0; d < 3; ++d) o[d] = a[d] - b[d]; };
    auto cross = [](const double* a, const double* b, double* o) {
      o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
    double e1[3], e2[3], e3[3], cr[3];
    sub(p[1], p[0], e1); sub(p[2], p[0], e2); sub(p[3], p[0], e3);
    cross(e1, e2, cr);
    const double vol = std::fabs(cr[0] * e3[0] + cr[1] * e3[1] + cr[2] * e3[2]) / 6.0;
    double area = 0.0;
    const int fc[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
    for (int f = 0; f < 4; ++f) {
      double u[3], w[3], n[3];
      sub(p[fc[f][1]], p[fc[f][0]], u); sub(p[fc[f][2]], p[fc[f][0]], w); cross(u, w, n);
      area += 0.5 * std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    }
    rmin = std::min(rmin, 3.0 * vol / area);
  }
  const double eps = 0.1 * rmin;
  // exterior facets = faces that belong to exactly one cell
  std::map<std::array<int, 3>, int> faces;
  const int fc[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
  for (int c = 0; c < nc; ++c)
    for (int f = 0; f < 4; ++f) {
      std::array<int, 3> key = {cells[4 * c + fc[f][0]], cells[4 * c + fc[f][1]], cells[4 * c + fc[f][2]]};
      std::sort(key.begin(), key.end());
      faces[key] += 1;
    }
  std::memset(flags, 0, nv);
  auto inside = [&](int v) {
    const double x = xyz[3 * v + coord];
    return x < mn[coord] + eps || x > mx[coord] - eps;
  };
  for (auto& kv : faces) {
    if (kv.second != 1) continue;
    const auto& k = kv.first;
    if (inside(k[0]) && inside(k[1]) && inside(k[2])) flags[k[0]] = flags[k[1]] = flags[k[2]] = 1;
  }
  int n = 0;
  for (int v = 0; v < nv; ++v) n += flags[v];
  return n;
}

// ---- sparsity: all vertex pairs sharing a cell, sorted, diagonal included [EXT DOLFIN
// SparsityPatternBuilder]; with blocked dofs the scalar pattern is its dense 3x3 expansion ------
static void vertex_adjacency(int nv, int nc, const int* cells, std::vector<std::vector<int>>& adj)
{
  adj.assign(nv, {});
  for (int c = 0; c < nc; ++c)
    for (int r = 0; r < 4; ++r)
      for (int s = 0; s < 4; ++s) adj[cells[4 * c + r]].push_back(cells[4 * c + s]);
  for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
}

int orc_block_pattern_count(int nv, int nc, const int* cells, int* IA)
{
  std::vector<std::vector<int>> adj;
  vertex_adjacency(nv, nc, cells, adj);
  IA[0] = 0;
  for (int v = 0; v < nv; ++v) IA[v + 1] = IA[v] + (int)adj[v].size();
  return IA[nv];
}

void orc_block_pattern_fill(int nv, int nc, const int* cells, const int* IA, int* JA)
{
  std::vector<std::vector<int>> adj;
  vertex_adjacency(nv, nc, cells, adj);
  for (int v = 0; v < nv; ++v) std::copy(adj[v].begin(), adj[v].end(), JA + IA[v]);
}

static inline int find_slot(const int* IA, const int* JA, int i, int j)
{
  const int* b = JA + IA[i];
  const int* e = JA + IA[i + 1];
  const int* p = std::lower_bound(b, e, j);
  return (int)(p - JA);
}

static void gather_cell(const int* cv, const double* xyz, const double* uu, const double* eps,
                        const double* fixed, const double* diff, const double* val,
                        const double* reac, double* x12, double* uu12, double* eps4, double* fc4,
                        double* d12, double* q12, double* r12)
{
  for (int r = 0; r < 4; ++r) {
    const int v = cv[r];
    for (int d = 0; d < 3; ++d) x12[3 * r + d] = xyz[3 * v + d];
    eps4[r] = eps[v];
    if (fc4) fc4[r] = fixed[v];
    for (int k = 0; k < 3; ++k) {
      uu12[4 * k + r] = uu[3 * v + k];
      d12[4 * k + r] = diff[3 * v + k];
      q12[4 * k + r] = val[3 * v + k];
      if (r12) r12[4 * k + r] = reac[3 * v + k];
    }
  }
}

void orc_eafe_matrix(int nv, const double* xyz, int nc, const int* cells, const int* IA,
                     const int* JA, const double* eta, const double* alpha, const double* beta,
                     const double* gamma, double* val)
{
  std::fill(val, val + IA[nv], 0.0);
  for (int c = 0; c < nc; ++c) {
    const int* cv = cells + 4 * c;
    double x12[12], e4[4], a4[4], b4[4], g4[4], A[16];
    for (int r = 0; r < 4; ++r) {
      for (int d = 0; d < 3; ++d) x12[3 * r + d] = xyz[3 * cv[r] + d];
      e4[r] = eta[cv[r]]; a4[r] = alpha[cv[r]]; b4[r] = beta[cv[r]]; g4[r] = gamma ? gamma[cv[r]] : 0.0;
    }
    kernels.eafe(A, e4, a4, b4, g4, x12);
    for (int r = 0; r < 4; ++r)
      for (int s = 0; s < 4; ++s) val[find_slot(IA, JA, cv[r], cv[s])] += A[4 * r + s];
  }
}

static void apply_bc_rows(int nv, const int* IA, const int* JA, const unsigned char* bcflag,
                          double* bsr)
{
  // DirichletBC::apply(A) [EXT]: row zeroed (pattern kept), unit diagonal; columns untouched
  for (int i = 0; i < nv; ++i)
    for (int k = 0; k < 3; ++k) {
      if (!bcflag[3 * i + k]) continue;
      for (int p = IA[i]; p < IA[i + 1]; ++p) {
        double* blk = bsr + 9 * (size_t)p;
        for (int c = 0; c < 3; ++c) blk[3 * k + c] = 0.0;
        if (JA[p] == i) blk[3 * k + k] = 1.0;
      }
    }
}

void orc_assemble(int nv, const double* xyz, int nc, const int* cells, const int* IA, const int* JA,
                  const double* uu, const double* eps, const double* fixed, const double* diff,
                  const double* val, const double* reac, const unsigned char* bcflag, int use_eafe,
                  double* bsr, double* b)
{
  const size_t nnzb = IA[nv];
  std::fill(bsr, bsr + 9 * nnzb, 0.0);
  std::fill(b, b + 3 * (size_t)nv, 0.0);
  // PDE::setup_linear_algebra (src/pde.cpp:542-555): assemble a and L cell by cell
  for (int c = 0; c < nc; ++c) {
    const int* cv = cells + 4 * c;
    double x12[12], uu12[12], e4[4], f4[4], d12[12], q12[12], r12[12], A[144], L[12];
    gather_cell(cv, xyz, uu, eps, fixed, diff, val, reac, x12, uu12, e4, f4, d12, q12, r12);
    kernels.jac(A, uu12, e4, d12, q12, x12);
    kernels.res(L, uu12, e4, f4, d12, q12, r12, x12);
    for (int r = 0; r < 4; ++r) {
      for (int s = 0; s < 4; ++s) {
        double* blk = bsr + 9 * (size_t)find_slot(IA, JA, cv[r], cv[s]);
        for (int kr = 0; kr < 3; ++kr)
          for (int kc = 0; kc < 3; ++kc) blk[3 * kr + kc] += A[(4 * kr + r) * 12 + 4 * kc + s];
      }
      for (int k = 0; k < 3; ++k) b[3 * cv[r] + k] += L[4 * k + r];
    }
  }
  apply_bc_rows(nv, IA, JA, bcflag, bsr);
  for (size_t i = 0; i < 3 * (size_t)nv; ++i) if (bcflag[i]) b[i] = 0.0;

  if (use_eafe) {
    // Linear_PNP::apply_eafe (linear_pnp.cpp:195-285): per ion k, eta = u_k,
    // beta = u_k + q_k*phi (q_k = valency function value at dofs 0..2, :220-224, q_0 forced 0),
    // alpha = D_k, gamma = 0; every (3i+k,3j+k) entry on the vertex pattern is overwritten.
    std::vector<double> eta(nv), beta(nv), alpha(nv), gamma(nv, 0.0), ev(nnzb);
    for (int k = 1; k < 3; ++k) {
      const double qk = val[k];
      for (int v = 0; v < nv; ++v) {
        eta[v] = uu[3 * v + k];
        beta[v] = uu[3 * v + k];
        if (qk != 0.0) { const double phi = uu[3 * v] * qk; beta[v] = beta[v] + phi; }
        alpha[v] = diff[3 * v + k];
      }
      orc_eafe_matrix(nv, xyz, nc, cells, IA, JA, eta.data(), alpha.data(), beta.data(), gamma.data(), ev.data());
      for (size_t p = 0; p < nnzb; ++p) bsr[9 * p + 3 * k + k] = ev[p];
    }
    apply_bc_rows(nv, IA, JA, bcflag, bsr);   // linear_pnp.cpp:78-80
  }
  // zero-row fix, src/pde.cpp:596-621
  for (int i = 0; i < nv; ++i)
    for (int k = 0; k < 3; ++k) {
      bool nonzero = false; double* diag = nullptr;
      for (int p = IA[i]; p < IA[i + 1]; ++p) {
        double* blk = bsr + 9 * (size_t)p;
        for (int c = 0; c < 3; ++c) if (blk[3 * k + c] != 0.0) nonzero = true;
        if (JA[p] == i) diag = &blk[3 * k + k];
      }
      if (!nonzero && diag) *diag = 1.0;
    }
}

void orc_residual(int nv, const double* xyz, int nc, const int* cells, const double* uu,
                  const double* eps, const double* fixed, const double* diff, const double* val,
                  const double* reac, const unsigned char* bcflag, double* b, double* l2, double* linf)
{
  const size_t N = 3 * (size_t)nv;
  std::fill(b, b + N, 0.0);
  for (int c = 0; c < nc; ++c) {
    const int* cv = cells + 4 * c;
    double x12[12], uu12[12], e4[4], f4[4], d12[12], q12[12], r12[12], L[12];
    gather_cell(cv, xyz, uu, eps, fixed, diff, val, reac, x12, uu12, e4, f4, d12, q12, r12);
    kernels.res(L, uu12, e4, f4, d12, q12, r12, x12);
    for (int r = 0; r < 4; ++r)
      for (int k = 0; k < 3; ++k) b[3 * cv[r] + k] += L[4 * k + r];
  }
  double s = 0.0, m = 0.0;
  for (size_t i = 0; i < N; ++i) {
    if (bcflag[i]) b[i] = 0.0;
    s += b[i] * b[i];
    m = std::max(m, std::fabs(b[i]));
  }
  if (l2) *l2 = std::sqrt(s);
  if (linf) *linf = m;   // src/pde.cpp:386-390: max(max, -min)
}

// ---- Error::compute_l2_error / compute_h1_error restated (src/error.cpp:58-103; forms
// include/L2Error.h, include/SemiH1error.h: int e^2 dx and int |grad e|^2 dx per component) ------
void orc_error_norms(int nv, const double* xyz, int nc, const int* cells, const double* uu,
                     const double* exact, double* l2, double* h1)
{
  (void)nv;
  double a = 0.0, b = 0.0;
  for (int c = 0; c < nc; ++c) {
    const int* cv = cells + 4 * c;
    double x12[12], g[4][3], det;
    for (int r = 0; r < 4; ++r) for (int d = 0; d < 3; ++d) x12[3 * r + d] = xyz[3 * cv[r] + d];
    tet_geometry(x12, g, &det);
    for (int k = 0; k < 3; ++k) {
      double e[4], gr[3] = {0, 0, 0};
      for (int r = 0; r < 4; ++r) {
        e[r] = uu[3 * cv[r] + k] - exact[3 * cv[r] + k];
        for (int d = 0; d < 3; ++d) gr[d] += e[r] * g[r][d];
      }
      for (int r = 0; r < 4; ++r) for (int s = 0; s < 4; ++s) a += det * e[r] * e[s] * (r == s ? 1.0 / 60.0 : 1.0 / 120.0);
      b += det / 6.0 * (gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
    }
  }
  *l2 = std::sqrt(a);
  *h1 = std::sqrt(a + b);
}

// ---- Mesh_Refiner::compute_entropy_error_vector restated (src/mesh_refiner.cpp:212-271) with the
// bindings of benchmarks/pnp_refine_exact/main.cpp:260-266,318-338: per ion species k = 1, 2 the
// entropy potential is u_k + q_k phi, the log weight u_k, the diffusivity D_k; the DG0 values of the
// species are added up. eta[nc].
void orc_entropy_indicator(int nv, const double* xyz, int nc, const int* cells, const double* uu,
                           const double* diff, const double* val, double* eta)
{
  (void)nv;
  for (int c = 0; c < nc; ++c) {
    const int* cv = cells + 4 * c;
    double x12[12];
    for (int r = 0; r < 4; ++r) for (int d = 0; d < 3; ++d) x12[3 * r + d] = xyz[3 * cv[r] + d];
    double e = 0.0;
    for (int k = 1; k < 3; ++k) {
      double mu[4], lw[4], D[4];
      for (int r = 0; r < 4; ++r) {
        const size_t o = 3 * (size_t)cv[r];
        lw[r] = uu[o + k]; mu[r] = uu[o + k] + val[o + k] * uu[o]; D[r] = diff[o + k];
      }
      e += marker_element(mu, lw, D, x12);
    }
    eta[c] = e;
  }
}

// ---- fasp_format_dcsr_dbsr restated for nb = 3 [EXT, SURVEY App. B] -------------------------
void orc_csr_to_bsr3_count(int n, const int* IA, const int* JA, int* bIA)
{
  const int nb = n / 3;
  bIA[0] = 0;
  for (int I = 0; I < nb; ++I) {
    std::vector<int> cols;
    for (int r = 3 * I; r < 3 * I + 3; ++r)
      for (int p = IA[r]; p < IA[r + 1]; ++p) cols.push_back(JA[p] / 3);
    std::sort(cols.begin(), cols.end());
    cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    bIA[I + 1] = bIA[I] + (int)cols.size();
  }
}

void orc_csr_to_bsr3_fill(int n, const int* IA, const int* JA, const double* val, const int* bIA,
                          int* bJA, double* bval)
{
  const int nb = n / 3;
  std::fill(bval, bval + 9 * (size_t)bIA[nb], 0.0);
  for (int I = 0; I < nb; ++I) {
    std::vector<int> cols;
    for (int r = 3 * I; r < 3 * I + 3; ++r)
      for (int p = IA[r]; p < IA[r + 1]; ++p) cols.push_back(JA[p] / 3);
    std::sort(cols.begin(), cols.end());
    cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    std::copy(cols.begin(), cols.end(), bJA + bIA[I]);
    for (int r = 3 * I; r < 3 * I + 3; ++r)
      for (int p = IA[r]; p < IA[r + 1]; ++p) {
        const int slot = bIA[I] + (int)(std::lower_bound(cols.begin(), cols.end(), JA[p] / 3) - cols.begin());
        bval[9 * (size_t)slot + 3 * (r - 3 * I) + JA[p] % 3] = val[p];
      }
  }
}

// ---- pnp_exact_solution physics (benchmarks/pnp_exact_solution/main.cpp:24-135, :293-302) ----
void orc_exact_problem(int nv, const double* xyz, double* eps, double* fixed, double* diff,
                       double* val, double* reac, double* uu_init, double* uu_exact)
{
  const double pi = 3.14159265359;           // main.cpp:24 (truncated on purpose)
  const double E = 1.0, cref = 1.0, perm = 1.0;
  double xmin = 1e20, xmax = -1e20;
  for (int v = 0; v < nv; ++v) { xmin = std::min(xmin, xyz[3 * v]); xmax = std::max(xmax, xyz[3 * v]); }
  auto sol = [&](double x, double* s) {
    s[0] = x * E + std::sin(pi * x); s[1] = std::log(cref) - x * x * E; s[2] = std::log(cref) + x * x * E; };
  auto dsol = [&](double x, double* s) { s[0] = E + pi * std::cos(pi * x); s[1] = -2.0 * x * E; s[2] = 2.0 * x * E; };
  auto ddsol = [&](double x, double* s) { s[0] = -(pi * pi) * std::sin(pi * x); s[1] = -2.0 * E; s[2] = 2.0 * E; };
  // Dirichlet values at the two x-faces (main.cpp:296-300 uses exact_solution(-1), (+1))
  double left[3], right[3];
  sol(-1.0, left); sol(+1.0, right);
  const double dist = 1.0 / (xmax - xmin);   // src/dirichlet.cpp:53
  for (int v = 0; v < nv; ++v) {
    const double x = xyz[3 * v];
    double s[3], ds[3], dds[3];
    sol(x, s); dsol(x, ds); ddsol(x, dds);
    eps[v] = perm;
    fixed[v] = -perm * dds[0] - (std::exp(s[1]) - std::exp(s[2]));           // main.cpp:61-66
    const double D[3] = {0.0, 1.0, 1.0};                                     // main.cpp:53-59
    for (int k = 0; k < 3; ++k) diff[3 * v + k] = D[k];
    val[3 * v] = 0.0; val[3 * v + 1] = 1.0; val[3 * v + 2] = -1.0;           // main.cpp:128-134
    reac[3 * v] = 0.0;                                                       // main.cpp:68-78
    reac[3 * v + 1] = -D[1] * std::exp(s[1]) * (ds[1] * (ds[1] + ds[0]) + (dds[1] + dds[0]));
    reac[3 * v + 2] = -D[2] * std::exp(s[2]) * (ds[2] * (ds[2] - ds[0]) + (dds[2] - dds[0]));
    for (int k = 0; k < 3; ++k) {
      // Linear_Function::eval, src/dirichlet.cpp:73-90
      double u = left[k] * (xmax - x) * dist;
      u += right[k] * (x - xmin) * dist;
      uu_init[3 * v + k] = u;
      if (uu_exact) uu_exact[3 * v + k] = s[k];
    }
  }
}

}
